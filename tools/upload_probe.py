"""Where the end-to-end overhead goes: cProfile of uploading R replicas from pinned host memory and reading results."""
import argparse, cProfile, pstats, sys, os, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, pyproton_oc_b200 as pkg
ap = argparse.ArgumentParser(); ap.add_argument("--replicas", type=int, default=128); a = ap.parse_args()
name, st, prm, tick0, mult = bench.load_workload()
n, stat, crim = pkg.CrimeEngine.capacity_for([st], 160000)
E = pkg.CrimeEngine(a.replicas, n, stat, crim, device=0); E.set_params(pkg.make_params(prm))
pinned = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in st.items()}
st_pin = {k: v.numpy() for k, v in pinned.items()}
def up():
    for r in range(a.replicas):
        E.upload_state(r, st_pin); E.set_crime_multiplier(mult, r)
    E.sync()
up()
t0 = time.perf_counter(); up(); print("upload %d replicas: %.1f ms" % (a.replicas, 1e3 * (time.perf_counter() - t0)))
pr = cProfile.Profile(); pr.enable(); up(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
t0 = time.perf_counter(); cols = [E.download_agents(r) for r in range(a.replicas)]; print("download agents: %.1f ms" % (1e3 * (time.perf_counter() - t0)))
for k, v in st.items():
    print(k, v.dtype, v.shape, end="; ")

# split of download_agents: the C call alone vs the Python wrapper (allocation of the 21 output columns)
import ctypes as C
from pyproton_oc_b200.engine import AgentCols, STATE_KEYS, _DTYPES
n = E.n_agents[0]
out = {key: np.empty(n, dtype=_DTYPES[f]) for f, key in STATE_KEYS.items()}
cols = AgentCols()
for f, key in STATE_KEYS.items():
    setattr(cols, f, out[key].ctypes.data)
t0 = time.perf_counter()
for r in range(a.replicas):
    E.L.poc_download_agents(E.h, r, C.byref(cols))
print("\nC call only, reused output arrays: %.1f ms" % (1e3 * (time.perf_counter() - t0)))
t0 = time.perf_counter()
for r in range(a.replicas):
    o2 = {key: np.empty(n, dtype=_DTYPES[f]) for f, key in STATE_KEYS.items()}
print("allocating the output arrays: %.1f ms" % (1e3 * (time.perf_counter() - t0)))
