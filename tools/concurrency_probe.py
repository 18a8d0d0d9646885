#!/usr/bin/env python
"""Probe: do two independent engines (two streams, two host threads) overlap on the GPU?"""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
import pyproton_oc_b200 as pkg

name, st, prm, tick0, mult = bench.load_workload()
n, stat, crim = pkg.CrimeEngine.capacity_for([st], 160000)
T = int(sys.argv[2]) if len(sys.argv) > 2 else 96
NE = int(sys.argv[1]) if len(sys.argv) > 1 else 2
Rtot = 128
engines = []
for e in range(NE):
    E = pkg.CrimeEngine(Rtot // NE, n, stat, crim)
    E.set_params(pkg.make_params(prm))
    engines.append(E)

def prep(E):
    for r in range(E.R):
        E.upload_state(r, st); E.set_crime_multiplier(mult, r)
    E.sync()

def run(E, i):
    E.run_ticks(tick0, T, [i * 1000 + r for r in range(E.R)], reporters=False)
    E.sync()

for rep in range(3):
    for E in engines: prep(E)
    t = time.perf_counter()
    th = [threading.Thread(target=run, args=(E, i)) for i, E in enumerate(engines)]
    for x in th: x.start()
    for x in th: x.join()
    dt = time.perf_counter() - t
    print("%d engines x %d replicas, %d ticks: %.1f ms  (%.1f us/tick, %.0f M agent-ticks/s)" % (
        NE, Rtot // NE, T, dt * 1e3, dt * 1e6 / T, 1e-6 * int(st["alive"].sum()) * T * Rtot / dt))
