"""The eight step() phases that exist on the device, at 10,000 agents, against the UNMODIFIED reference in distribution:
tests/golden/distdemog4_n10000_s42.npz holds 14 continuations (36 ticks) of the reference's own 10,000-agent setup state by the
reference with graduate_and_enter_jobmarket, update_unemployment_status, wedding, retire_persons, make_baby, remove_excess_*,
make_friends, make_people_die and the crime path active and every other phase a no-op (oracle/make_states.py distdemogpar).
The CPU restatement runs as many continuations with its own keys; tests/test_gpu_demog_large.py runs the same keys on the device,
where every run equals the restatement's bit for bit (tests/test_gpu_labour.py), so the two files pass or fail together."""
import numpy as np
from scipy.stats import ks_2samp

import util as U
from oracle import oracle_c as O

NAME, KEY0 = "distdemog4_n10000_s42", 9600
CHECKS = (11, 23, 35)
COLS = ["current_num_persons", "tot_friendship_link", "tot_household_link", "tot_partner_link", "tot_professional_link",
        "tot_offspring_link", "employed", "number_crimes", "current_oc_members", "number_weddings", "number_born",
        "number_deceased", "education_level_mean", "number_students"]
LAYER_OF = {"tot_friendship_link": 5, "tot_household_link": 4, "tot_partner_link": 3, "tot_professional_link": 7,
            "tot_offspring_link": 1}


def hists_of(c, school_level):
    al = c["alive"] > 0
    return np.array([np.bincount(c["job_level"][al], minlength=8)[:8], np.bincount(c["education_level"][al], minlength=8)[:8],
                     np.bincount(c["wealth_level"][al], minlength=8)[:8],
                     np.bincount(np.where(c["has_school"][al] > 0, school_level[al], 0), minlength=8)[:8]], dtype=np.float64)


# Two columns are NOT matched in distribution at this size, and the tests say by how much.  At 10,000 agents the reference's setup
# leaves 282 agents over the cap of 30 professional links (employers of up to 248 people; at 1,000 agents the largest employer
# has 12 and nobody is over the cap).  The reference walks them one after the other, so an agent whose colleague has just dropped
# their link finds itself with less to drop (model.py:646-659); in the batch form (oracle/hotpath.c remove_excess) every agent
# sheds its whole surplus as counted before the phase, both ends of a pair can each drop other links, and the first tick removes
# about 900 directed entries (6 %) more.  The gap does not accumulate (933, 837, 769 entries at ticks 12, 24, 36).  Colleagues
# are candidates of make_friends, so the friendship layer follows at 1 %.  Everything else is KS-matched.
BOUNDED = {"tot_professional_link": (-0.08, 0.01), "tot_friendship_link": (-0.02, 0.005)}


def check_against_reference(fx, mine, hists):
    """mine[col]: [runs][len(CHECKS)] at the ticks CHECKS; hists: [runs][4][8] after the last tick."""
    for col, (lo, hi) in BOUNDED.items():
        ref = fx["refdemog_" + col][:, list(CHECKS)].mean(0)
        gap = (mine[col].mean(0) - ref) / ref
        assert (gap >= lo).all() and (gap <= hi).all(), (col, gap)
        assert abs(gap[-1]) <= abs(gap[0]) + 0.002, (col, gap)   # a one-off of the first tick, not a drift
    for col in COLS:
        if col in BOUNDED:
            continue
        ref = fx["refdemog_" + col]
        for ci, t in enumerate(CHECKS):
            a, b = mine[col][:, ci], ref[:, t]
            p = ks_2samp(a, b).pvalue
            assert p > 0.005, (col, t, p, a.mean(), b.mean())
        if ref[:, -1].mean() > 20:   # means at the end of the run: within four standard errors of the difference plus 2 %
            a, b = mine[col][:, -1], ref[:, -1]
            se = np.sqrt(a.var() / len(a) + b.var() / len(b))
            assert abs(a.mean() - b.mean()) <= 4 * se + 0.02 * b.mean(), (col, a.mean(), b.mean(), se)
    ref_h = fx["refdemog_final_hists"]
    for i, nm in enumerate(["job_level", "education_level", "wealth_level", "school_level"]):
        for v in range(8):
            a, b = hists[:, i, v], ref_h[:, i, v]
            if a.mean() < 3 and b.mean() < 3:
                continue
            se = np.sqrt(a.var() / len(a) + b.var() / len(b))
            assert abs(a.mean() - b.mean()) <= 4 * se + 0.02 * b.mean() + 1.0, (nm, v, a.mean(), b.mean(), se)


def test_oracle_eight_phases_at_10k_in_distribution_against_the_reference():
    fx = U.load_fixture(U.GOLDEN + "/%s.npz" % NAME)
    st, prm = U.pre_state(fx), U.fixture_params(fx)
    runs, ticks = fx["refdemog_number_crimes"].shape
    assert ticks == CHECKS[-1] + 1
    P, D, L = O.make_params(prm), O.make_demog(), O.make_labour()
    mine = {c: np.zeros((runs, len(CHECKS))) for c in COLS}
    hists = []
    for r in range(runs):
        S = O.OracleState(st)
        S.reserve(len(st["alive"]) + 3000)
        S.set_children(st["number_of_children"])
        S.set_labour(L)
        mult, crimes, done = 0.0, 0, 0
        for ci, te in enumerate(CHECKS):
            mult, tot = S.run_ticks_phases(P, D, 255, done + 1, te + 1 - done, KEY0 + r, mult)
            crimes += int(tot[0])
            done = te + 1
            c = S.cols()
            al = c["alive"] > 0
            dc = S.demog_counters()
            vals = {"current_num_persons": al.sum(), "employed": (c["has_job"][al] > 0).sum(), "number_crimes": crimes,
                    "current_oc_members": c["oc_member"][al].sum(), "number_weddings": dc["number_weddings"],
                    "number_born": dc["number_born"], "number_deceased": dc["number_deceased"],
                    "education_level_mean": c["education_level"][al].mean(), "number_students": (c["has_school"][al] > 0).sum()}
            for k, l in LAYER_OF.items():
                vals[k] = np.diff(S.layer(l)[0])[al].sum()
            for k in COLS:
                mine[k][r, ci] = vals[k]
        hists.append(hists_of(c, S.education()[1]))
    check_against_reference(fx, mine, np.array(hists))


def test_the_deviation_is_the_batch_order_of_remove_excess():
    """The diagnosis behind BOUNDED, checked: with the two remove_excess phases run in the reference's order (agents one after the
    other, every drop applied before the next agent counts its links; diagnostic bit 8 of orc_run_ticks_phases, no kernel follows
    it) the restatement's professional and friendship layers are KS-matched to the reference's at 10,000 agents too."""
    fx = U.load_fixture(U.GOLDEN + "/%s.npz" % NAME)
    st, prm = U.pre_state(fx), U.fixture_params(fx)
    runs = fx["refdemog_number_crimes"].shape[0]
    P, D, L = O.make_params(prm), O.make_demog(), O.make_labour()
    mine = {c: np.zeros((runs, len(CHECKS))) for c in BOUNDED}
    for r in range(runs):
        S = O.OracleState(st)
        S.reserve(len(st["alive"]) + 3000)
        S.set_children(st["number_of_children"])
        S.set_labour(L)
        mult, done = 0.0, 0
        for ci, te in enumerate(CHECKS):
            mult, _ = S.run_ticks_phases(P, D, 255 | 256, done + 1, te + 1 - done, KEY0 + r, mult)
            done = te + 1
            al = S.cols()["alive"] > 0
            for c in BOUNDED:
                mine[c][r, ci] = np.diff(S.layer(LAYER_OF[c])[0])[al].sum()
    for c in BOUNDED:
        ref = fx["refdemog_" + c]
        for ci, t in enumerate(CHECKS):
            p = ks_2samp(mine[c][:, ci], ref[:, t]).pvalue
            assert p > 0.01, (c, t, p, mine[c][:, ci].mean(), ref[:, t].mean())
        assert abs(mine[c][:, -1].mean() - ref[:, -1].mean()) <= 0.012 * ref[:, -1].mean(), c


def test_rounds_of_non_adjacent_agents_reproduce_the_sequential_order():
    """The parallel form of the reference's order proposed in DESIGN.md section 9 (an over-cap agent acts in the round in which no
    lower-slot neighbour of the layer is still over its cap and waiting), against the agent-by-agent walk, in the restatement:
    the same state, layer for layer.  51 rounds on the first tick after the 10,000-agent setup (the longest chain inside the big
    employers), none on the ticks after it."""
    fx = U.load_fixture(U.GOLDEN + "/%s.npz" % NAME)
    st = U.pre_state(fx)
    A, B = O.OracleState(st), O.OracleState(st)
    rounds = []
    for t in (1, 2, 3):
        A.begin_tick(t); B.begin_tick(t)
        A.remove_excess_diagnostic(0, t, 5)
        rounds.append(B.remove_excess_diagnostic(1, t, 5))
        for l in (5, 7):
            for x, y in zip(A.layer(l)[:2], B.layer(l)[:2]):
                assert np.array_equal(x, y), (t, l)
    assert 10 < rounds[0] < 250 and rounds[1] == 0 and rounds[2] == 0
    deg = np.diff(A.layer(7)[0])
    assert deg.max() <= 30 and len(A.layer(7)[1]) < len(st["col_7"])
