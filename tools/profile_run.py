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
    ap.add_argument("--repeat", type=int, default=1, help="repeat upload+run and report the fastest")
    ap.add_argument("--phases", type=int, default=0, help="device phases of step() to run (CrimeEngine.PHASE_*: 127 = the seven of every tick, 255 = + the yearly labour-status block)")
    ap.add_argument("--synthetic", type=int, default=0, help="use a synthetic population of this many agents (config 4) instead of the bench workload")
    args = ap.parse_args()
    import bench
    import pyproton_oc_b200 as pkg
    cfg = bench.homogeneous_config(1)
    name = cfg["name"]
    st, prm, tick0, mult = cfg["entries"][0]
    if args.synthetic:
        import util as U
        from pyproton_oc_b200 import synth
        tpl = U.pre_state(U.load_fixture(U.GOLDEN + "/state_n10000_s42_baseline.npz"))
        st = synth.synthetic_state(args.synthetic, seed=42, template=tpl, num_oc_persons=30)
        name, prm, tick0, mult = "synthetic_n%d_s42" % args.synthetic, {}, 1, 0.0
    if args.tick0 is not None:
        tick0 = args.tick0
    n, stat, crim = pkg.CrimeEngine.capacity_for([st], 16 * len(st["alive"]))
    if args.phases:
        n, stat = n + 3000, stat + 100000
    E = pkg.CrimeEngine(args.replicas, n, stat, crim)
    E.set_params(pkg.make_params(prm))
    if args.phases:
        E.set_demography(pkg.make_demography())
        if args.phases & E.PHASE_LABOUR:
            E.set_labour(pkg.make_labour())
        E.set_phases(args.phases)
    keys = [r * 7919 + 17 for r in range(args.replicas)]

    def upload(r):
        E.upload_state(r, st)
        E.set_crime_multiplier(mult, r)
        if (args.phases & E.PHASE_LABOUR) and "max_education_level" in st:
            E.upload_education(r, st["max_education_level"], st["school_level"])

    best = None
    for rep_i in range(args.repeat - 1):
        for r in range(args.replicas):
            upload(r)
        E.sync(); E.marker(0)
        E.run_ticks(tick0, args.ticks, keys, reporters=False)
        E.marker(1); E.sync()
        ms = E.marker_elapsed_ms(0, 1)
        best = ms if best is None else min(best, ms)
    if best is not None:
        print("device time, best of %d: %.2f ms (%.1f us/tick, %.1f M agent-ticks/s)" % (
            args.repeat - 1, best, 1e3 * best / args.ticks, 1e-6 * int(st["alive"].sum()) * args.ticks * args.replicas / (best / 1e3)))
    for r in range(args.replicas):
        upload(r)
    if args.warm_ticks:
        E.run_ticks(tick0, args.warm_ticks, keys, reporters=False)
        tick0 += args.warm_ticks
        E.stats(reset=True)   # the counters printed below cover the measured ticks only
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
