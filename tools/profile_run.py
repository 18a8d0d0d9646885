#!/usr/bin/env python
"""Minimal driver for ncu / timing: upload the bench workload, run `--ticks` ticks for `--replicas` replicas."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--replicas", type=int, default=16)
    ap.add_argument("--ticks", type=int, default=24)
    ap.add_argument("--tick0", type=int, default=None)
    ap.add_argument("--warm-ticks", type=int, default=0, help="ticks run before the measured ones (state ages)")
    ap.add_argument("--timing", action="store_true")
    args = ap.parse_args()
    import bench
    import pyproton_oc_b200 as pkg
    name, st, prm, tick0, mult = bench.load_workload()
    if args.tick0 is not None:
        tick0 = args.tick0
    n, stat, crim = pkg.CrimeEngine.capacity_for([st], 160000)
    E = pkg.CrimeEngine(args.replicas, n, stat, crim)
    E.set_params(pkg.make_params(prm))
    for r in range(args.replicas):
        E.upload_state(r, st)
        E.set_crime_multiplier(mult, r)
    keys = [r * 7919 + 17 for r in range(args.replicas)]
    if args.warm_ticks:
        E.run_ticks(tick0, args.warm_ticks, keys, reporters=False)
        tick0 += args.warm_ticks
    E.sync()
    if args.timing:
        E.set_timing(True)
        E.timers_reset()
    t = time.perf_counter()
    rep = E.run_ticks(tick0, args.ticks, keys, reporters=True)
    E.sync()
    dt = time.perf_counter() - t
    print("workload %s  replicas %d  ticks %d..%d  wall %.1f ms  (%.1f us/tick)" % (
        name, args.replicas, tick0, tick0 + args.ticks - 1, dt * 1e3, dt * 1e6 / args.ticks))
    print("crimes %d  oc members %d  emb evals(last tick) %d" % (rep[0, -1, 0], rep[0, -1, 7], E.counters()[0][0, 13]))
    print("stats", E.stats())
    if args.timing:
        ms, nl = E.timers()
        for k in ms:
            if nl[k]:
                print("  %-13s %4d launches  %8.3f ms  %7.1f us/launch" % (k, nl[k], ms[k], 1e3 * ms[k] / nl[k]))


if __name__ == "__main__":
    main()
