#!/usr/bin/env python
"""Time the unmodified reference on this box's host cores (one core per run).  Examples:
    python tools/time_reference.py --agents 3000 --ticks 480            # BASELINE config 2 in full
    python tools/time_reference.py --agents 10000 --ticks 24            # a prefix of config 3 / 5 size
Prints one JSON line per run."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--agents", type=int, default=3000)
    ap.add_argument("--ticks", type=int, default=48)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--max-seconds", type=float, default=None)
    a = ap.parse_args()
    import warnings
    warnings.filterwarnings("ignore")
    from oracle import ref_timing
    r = ref_timing.time_reference(a.agents, a.ticks, a.seed, max_seconds=a.max_seconds)
    r["cores"] = 1
    r["host_cores_available"] = len(os.sched_getaffinity(0))
    print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
