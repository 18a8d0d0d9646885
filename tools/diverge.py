"""Find the first tick at which the CUDA path and the C oracle differ (debug aid; uses oracle/ as the checker)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util as U
import pyproton_oc_b200 as P
from oracle import oracle_c as O
from pyproton_oc_b200.model import ProtonOC

name = sys.argv[1] if len(sys.argv) > 1 else "state_n3000_s42_baseline"
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 480
fx = U.load_fixture(U.GOLDEN + "/%s.npz" % name)
st = U.pre_state(fx)
m = ProtonOC(seed=1); m.choose_intervention_setting()
prm = {k: getattr(m, k) for k in P.HOT_PATH_DEFAULTS}
key = 20261020
S = O.OracleState(st); OP = O.make_params(prm)
n, stat, crim = P.CrimeEngine.capacity_for([st], 400000)
E = P.CrimeEngine(1, n, stat, crim)
E.set_params(P.make_params(prm)); E.upload_state(0, st)
mult = 0.0
COLS = ["oc_member", "num_crimes_committed", "prisoner", "criminal_tendency", "sentence_countdown", "new_recruit", "has_job", "has_school", "job_level"]
for t in range(ticks):
    mult, tot = S.run_ticks(OP, 1 + t, 1, key, mult)
    E.run_ticks(1 + t, 1, [key])
    g = E.download_state(0); o = S.cols()
    bad = [k for k in COLS if not np.array_equal(g[k], o[k])]
    rp, col, co = S.layer(6)
    if not (np.array_equal(g["rowptr_6"], rp) and np.array_equal(g["col_6"], col) and np.array_equal(g["co_off"], co)):
        bad.append("criminal layer")
    if bad:
        print("tick", 1 + t, "differs in", bad)
        for k in bad:
            if k in g:
                d = np.nonzero(g[k] != o[k])[0]
                print(" ", k, "n diff", len(d), "first", d[:8], "gpu", g[k][d[:8]], "oracle", o[k][d[:8]])
        cn, _ = E.counters(); print("counters", cn[0].tolist())
        break
else:
    print("no divergence in", ticks, "ticks")
