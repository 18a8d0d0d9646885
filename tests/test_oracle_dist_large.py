"""North-star check 2 at the sizes the path is benchmarked at: whole-run outputs of the crime path against the UNMODIFIED
reference, in distribution over seeds.  tests/golden/dist_n3000_s42.npz (24 runs x 96 ticks) and dist_n10000_s42.npz (14 runs x
36 ticks) hold continuations of ONE reference setup state by the reference's own code for exactly this path (every other phase
of step() a no-op; oracle/make_states.py disthot); dist_n1000_s42_oc (28 x 60 ticks, 30 OC members among 1,000 agents: the
OC-initiated branch with its embeddedness-weighted candidates every few ticks) and dist_n3000_s42_sample0 (14 x 36 ticks, the
samples/sample_json.json variant with facilitator repression and 24 OC members) do the same for two other parameter sets.  Here the CPU restatement runs the same number of continuations with its own
Philox keys; tests/test_gpu_dist_large.py runs the same keys on the device, where every run equals the restatement's bit for bit
(tests/test_gpu_parity.py), so the two files pass or fail together."""
import numpy as np
import pytest
from scipy.stats import ks_2samp

import util as U
from oracle import oracle_c as O

CASES = [("dist_n3000_s42", 7000), ("dist_n10000_s42", 7100), ("dist_n1000_s42_oc", 7200), ("dist_n3000_s42_sample0", 7300)]
CHECK_COLS = ["current_oc_members", "people_jailed", "current_prisoners", "tot_criminal_link", "number_crimes_committed_of_persons",
              "big_crime_from_small_fish", "criminal_tendency_mean", "num_crime_committed_sd"]


def check_against_reference(fx, mine, ticks):
    """mine[col]: [runs][ticks].  Per-tick crime counts are round(expected) per cell: the same set of values while the
    population is the same; every other column KS-matched at three ticks."""
    ref = {c: fx["refhot_" + c] for c in ["number_crimes"] + CHECK_COLS}
    assert np.array_equal(np.unique(np.diff(mine["number_crimes"], axis=1)), np.unique(np.diff(ref["number_crimes"], axis=1)))
    assert ks_2samp(np.diff(mine["number_crimes"], axis=1).ravel(), np.diff(ref["number_crimes"], axis=1).ravel()).pvalue > 0.01
    assert np.array_equal(mine["number_crimes"][:, 0], ref["number_crimes"][:, 0][:len(mine["number_crimes"])])
    for col in CHECK_COLS:
        for t in (11, ticks // 2 + 5, ticks - 1):
            a, b = mine[col][:, t], ref[col][:, t]
            # a column without spread is a constant of the state, not a sample: compare the value.  criminal_tendency_mean is
            # one whenever nobody is born or dies -- the epsilon pass (model.py:1183-1186) pins every cell's mean to its c, so
            # the population mean is sum(|cell| * c) / N in every run and what differs between runs (and between np.mean and
            # the reporter kernel's sum) is rounding noise in the last bits, which a rank test would mistake for a signal
            if max(np.ptp(a), np.ptp(b)) <= 1e-9 * max(abs(a.mean()), abs(b.mean())):
                assert a.mean() == pytest.approx(b.mean(), rel=1e-9), (col, t)
                continue
            p = ks_2samp(a, b).pvalue
            assert p > 0.01, (col, t, p, a.mean(), b.mean())
    # means of the two head-line columns within a few per cent at the end of the run
    for col, tol in (("tot_criminal_link", 0.05), ("number_crimes_committed_of_persons", 0.02)):
        a, b = mine[col][:, -1].mean(), ref[col][:, -1].mean()
        assert abs(a - b) <= tol * b, (col, a, b)


def oracle_columns(fx, key0):
    st, prm = U.pre_state(fx), U.fixture_params(fx)
    runs, ticks = fx["refhot_number_crimes"].shape
    P = O.make_params(prm)
    mine = {c: np.zeros((runs, ticks)) for c in ["number_crimes"] + CHECK_COLS}
    for r in range(runs):
        S = O.OracleState(st)
        mult, crimes, jailed, small_fish = float(fx["crime_multiplier"]), 0, 0, 0
        for t in range(ticks):
            S.begin_tick(1 + t)
            if (1 + t) % 12 == 0 or t == 0:
                S.tendency(P)
                mult = S.crime_multiplier(P)
            res = S.commit_crimes(P, 1 + t, mult, philox_key=key0 + r)
            o = res["out"]
            crimes += o["d_number_crimes"]; jailed += o["d_people_jailed"]; small_fish += o["d_big_crime_from_small_fish"]
            S.end_tick()
            c = S.cols()
            al = c["alive"] > 0
            mine["number_crimes"][r, t] = crimes
            mine["people_jailed"][r, t] = jailed
            mine["big_crime_from_small_fish"][r, t] = small_fish
            mine["current_oc_members"][r, t] = c["oc_member"][al].sum()
            mine["current_prisoners"][r, t] = c["prisoner"][al].sum()
            mine["tot_criminal_link"][r, t] = np.diff(S.layer(6)[0])[al].sum()
            mine["number_crimes_committed_of_persons"][r, t] = c["num_crimes_committed"][al].sum()
            mine["criminal_tendency_mean"][r, t] = c["criminal_tendency"][al].mean()
            mine["num_crime_committed_sd"][r, t] = np.std(c["num_crimes_committed"][al])
    return mine, ticks


@pytest.mark.parametrize("name,key0", CASES)
def test_oracle_whole_runs_in_distribution_against_the_reference(name, key0):
    fx = U.load_fixture(U.GOLDEN + "/%s.npz" % name)
    mine, ticks = oracle_columns(fx, key0)
    check_against_reference(fx, mine, ticks)
