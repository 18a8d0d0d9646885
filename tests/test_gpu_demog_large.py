"""The eight device phases of step() at 10,000 agents on the device against 14 continuations of the reference's own setup state
by the UNMODIFIED reference with exactly these phases active (tests/golden/distdemog4_n10000_s42.npz; same keys and checks as
tests/test_oracle_demog_large.py, the CPU restatement's half)."""
import numpy as np
import pytest

import util as U
from test_oracle_demog_large import CHECKS, COLS, KEY0, NAME, check_against_reference, hists_of

pytestmark = pytest.mark.gpu


def test_device_eight_phases_at_10k_in_distribution_against_the_reference():
    import pyproton_oc_b200 as P
    fx = U.load_fixture(U.GOLDEN + "/%s.npz" % NAME)
    st, prm = U.pre_state(fx), U.fixture_params(fx)
    runs, ticks = fx["refdemog_number_crimes"].shape
    n, stat, crim = P.CrimeEngine.capacity_for([st], 16 * len(st["alive"]))
    with P.CrimeEngine(runs, n + 3000, stat + 100000, crim) as E:
        E.set_params(P.make_params(prm))
        E.set_demography(P.make_demography())
        E.set_labour(P.make_labour())
        E.set_phases(E.PHASE_EVERY)
        for r in range(runs):
            E.upload_state(r, st)
            E.upload_children(r, st["number_of_children"])
            E.upload_education(r, st["max_education_level"], st["school_level"])
        rep = E.run_ticks(1, ticks, [KEY0 + r for r in range(runs)])
        cn, _ = E.counters()
        hists = []
        for r in range(runs):
            hists.append(hists_of(E.download_agents(r), E.download_education(r)[1]))
    assert (cn[:, 9] == 0).all()
    names = list(P.REPORTER_NAMES)
    mine = {c: rep[:, list(CHECKS), names.index(c)] for c in COLS}
    check_against_reference(fx, mine, np.array(hists))
