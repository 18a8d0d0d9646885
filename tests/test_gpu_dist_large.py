"""North-star check 2 at the sizes the path is benchmarked at, on the device: 24 replicas x 3,000 agents x 96 ticks and 14
replicas x 10,000 agents x 36 ticks from the reference's own setup states against as many continuations of those states by
the UNMODIFIED reference with exactly the crime path active (tests/golden/dist_n3000_s42.npz, dist_n10000_s42.npz; same keys
and same checks as the CPU restatement's half, tests/test_oracle_dist_large.py)."""
import numpy as np
import pytest

import util as U
from test_oracle_dist_large import CASES, CHECK_COLS, check_against_reference

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,key0", CASES)
def test_device_whole_runs_in_distribution_against_the_reference(name, key0):
    import pyproton_oc_b200 as P
    fx = U.load_fixture(U.GOLDEN + "/%s.npz" % name)
    st, prm = U.pre_state(fx), U.fixture_params(fx)
    runs, ticks = fx["refhot_number_crimes"].shape
    n, stat, crim = P.CrimeEngine.capacity_for([st], 16 * len(st["alive"]))
    with P.CrimeEngine(runs, n, stat, crim) as E:
        E.set_params(P.make_params(prm))
        for r in range(runs):
            E.upload_state(r, st)
            E.set_crime_multiplier(float(fx["crime_multiplier"]), r)
        rep = E.run_ticks(1, ticks, [key0 + r for r in range(runs)])
        cn, _ = E.counters()
    assert (cn[:, 9] == 0).all()
    names = list(P.REPORTER_NAMES)
    mine = {c: rep[:, :, names.index(c)] for c in ["number_crimes"] + CHECK_COLS}
    check_against_reference(fx, mine, ticks)
