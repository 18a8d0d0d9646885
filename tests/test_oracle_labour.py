"""The CPU restatement of the yearly labour-status block (oracle/hotpath.c orc_labour_yearly: graduate_and_enter_jobmarket,
model.py:1351-1380, and the update_unemployment_status loop of step(), model.py:252-256) against the UNMODIFIED reference, in
distribution: tests/golden/distdemog4_n1000_s42.npz holds 30 continuations (60 ticks = six yearly ticks) of one reference setup
state by the reference with graduate_and_enter_jobmarket, the seven per-tick phases and the crime path active
(oracle/make_states.py distdemog4; find_job, migrants and the interventions are no-ops there, as they are on the device).
The kernel is then checked bit for bit against this restatement (tests/test_gpu_labour.py)."""
import ctypes as C

import numpy as np
import pytest
from scipy.stats import ks_2samp

import util as U
from oracle import oracle_c as O


@pytest.fixture(scope="module")
def fx():
    return U.load_fixture(U.GOLDEN + "/distdemog4_n1000_s42.npz")


def _hists(c, school_level):
    al = c["alive"] > 0
    return np.array([np.bincount(c["job_level"][al], minlength=8)[:8], np.bincount(c["education_level"][al], minlength=8)[:8],
                     np.bincount(c["wealth_level"][al], minlength=8)[:8],
                     np.bincount(np.where(c["has_school"][al] > 0, school_level[al], 0), minlength=8)[:8]], dtype=np.float64)


def test_labour_tables_match_the_reference_layout():
    """make_labour: the product struct and the oracle's mirror are the same bytes; every choice CDF ends at 1; the tables are the
    reference's (schools.csv levels, 4 work statuses, 5 wealth levels, ages 15..64, the five range ages)."""
    import pyproton_oc_b200 as P
    a, b = P.make_labour(), O.make_labour()
    assert C.sizeof(a) == C.sizeof(b) == O.lib().orc_sizeof_labour()
    assert bytes(a) == bytes(b)
    assert a.n_levels == 4 and a.primary_age == 6 and list(a.end_age)[:5] == [0, 10, 13, 18, 25]
    for lv in range(1, 5):
        for m in (0, 1):
            assert a.work_n[lv][m] == 4 and a.work_cdf[lv][m][3] == 1.0 and list(a.work_val[lv][m])[:4] == [1, 2, 3, 4]
            assert a.wealth_n[lv][m] == 5 and a.wealth_cdf[lv][m][4] == 1.0 and list(a.wealth_val[lv][m])[:5] == [1, 2, 3, 4, 5]
    assert a.work_n[0][0] == 0 and a.wealth_n[5][1] == 0   # no rows: education level 0 never graduates, work status 5 has no rates
    for m in (0, 1):
        assert [x for x in range(128) if a.range_age[m][x]] == [15, 25, 35, 45, 55]
        assert all(a.inactive[m][x] < 0 for x in list(range(15)) + list(range(65, 128)))
        assert all(0 < a.inactive[m][x] < 1 for x in range(15, 65))
    # a female primary-school leaver is unemployed (rate 1.0 on work status 1): the CDF puts every u in [0, 1) on row 0
    assert a.work_cdf[1][0][0] == 1.0


def test_labour_tables_are_validated():
    import pyproton_oc_b200 as P
    with pytest.raises(ValueError):
        P.make_labour(education_levels={1: (6, 10), 3: (14, 18)})           # levels must be numbered 1..n
    with pytest.raises(ValueError):
        P.make_labour(work_status={9: {True: [[1, 1.0]], False: [[1, 1.0]]}})   # education level past the table
    lb = P.make_labour(education_levels={1: (5, 9), 2: (10, 14)}, range_ages={True: [20], False: [30]})
    assert lb.n_levels == 2 and lb.primary_age == 5 and list(lb.end_age)[:3] == [0, 9, 14]
    assert lb.range_age[1][20] == 1 and lb.range_age[0][30] == 1 and lb.range_age[0][20] == 0


def test_labour_block_in_distribution_against_the_reference(fx):
    st, prm = U.pre_state(fx), U.fixture_params(fx)
    runs, ticks = fx["refdemog_number_crimes"].shape
    assert ticks == 60 and "max_education_level" in st
    P, D, L = O.make_params(prm), O.make_demog(), O.make_labour()
    mine_h, edu_mean, students, born, employed, prof = [], [], [], [], [], []
    n_runs = 16
    for j in range(n_runs):
        S = O.OracleState(st)
        S.reserve(len(st["alive"]) + 600)
        S.set_children(st["number_of_children"])
        S.set_labour(L)
        S.run_ticks_phases(P, D, 255, 1, ticks, 8800 + j, float(fx["crime_multiplier"]))
        c = S.cols()
        al = c["alive"] > 0
        mine_h.append(_hists(c, S.education()[1]))
        edu_mean.append(c["education_level"][al].mean())
        students.append(int((c["has_school"][al] > 0).sum()))
        born.append(S.demog_counters()["number_born"])
        employed.append(int((c["has_job"][al] > 0).sum()))
        prof.append(int(np.diff(S.layer(7)[0])[al].sum()))
    mine_h, ref_h = np.array(mine_h), fx["refdemog_final_hists"]
    # people hold jobs in this fixture (find_job is neutralised after setup, not before): retire_persons and
    # remove_excess_professional_links move these two columns
    assert fx["refdemog_employed"][:, 0].mean() > 100 and fx["refdemog_tot_professional_link"][:, 0].mean() > 300
    assert ks_2samp(employed, fx["refdemog_employed"][:, -1]).pvalue > 0.005
    assert ks_2samp(prof, fx["refdemog_tot_professional_link"][:, -1]).pvalue > 0.005
    assert ks_2samp(edu_mean, fx["refdemog_education_level_mean"][:, -1]).pvalue > 0.005
    # (make_people_die takes the dead out of their school, model.py:1597-1598: number_students counts the living on both sides)
    assert abs(np.mean(students) - fx["refdemog_number_students"][:, -1].mean()) <= 4.0
    assert ks_2samp(born, fx["refdemog_number_born"][:, -1]).pvalue > 0.005
    names = ["job_level", "education_level", "wealth_level", "school_level"]
    for i, nm in enumerate(names):
        for v in range(8):
            a, b = mine_h[:, i, v], ref_h[:, i, v]
            if a.mean() < 3 and b.mean() < 3:
                continue
            se = np.sqrt(a.var() / len(a) + b.var() / len(b))
            assert abs(a.mean() - b.mean()) <= 4 * se + 0.03 * b.mean() + 1.0, (nm, v, a.mean(), b.mean(), se)
    # and the block matters: without it (phases 127) the children born or of age 6 during the run stay at education level 0
    S = O.OracleState(st)
    S.reserve(len(st["alive"]) + 600)
    S.set_children(st["number_of_children"])
    S.run_ticks_phases(P, D, 127, 1, ticks, 8800, float(fx["crime_multiplier"]))
    c = S.cols()
    h0 = np.bincount(c["education_level"][c["alive"] > 0], minlength=8)
    assert h0[0] > ref_h[:, 1, 0].mean() + 25


def test_labour_block_invariants(fx):
    """Properties that hold whatever the draws: a student sits in a school whose age range (plus the year of grace) holds him;
    graduates carry the level of the school they left; nobody passes max_education_level; the block
    changes nothing but the four attributes and the school level; a phase without tables is refused."""
    st, prm = U.pre_state(fx), U.fixture_params(fx)
    P, D, L = O.make_params(prm), O.make_demog(), O.make_labour()
    S = O.OracleState(st)
    with pytest.raises(RuntimeError):
        S.labour_yearly(P, D, 12, 1)
    S.set_labour(L)
    S.reserve(len(st["alive"]) + 100)
    n_grad_total = 0
    for year in range(1, 9):
        t = 12 * year
        S.begin_tick(t)
        before, (me0, sl0) = S.cols(), S.education()
        layers0 = [S.layer(l) for l in range(9)]
        ngrad, ndraw = S.labour_yearly(P, D, t, 99)
        n_grad_total += ngrad
        c, (me, sl) = S.cols(), S.education()
        assert np.array_equal(me, me0)
        for k in c:
            if k not in ("job_level", "education_level", "wealth_level", "has_school"):
                assert np.array_equal(c[k], before[k]), k
        for l in range(9):
            for x, y in zip(S.layer(l), layers0[l]):
                assert (x is None and y is None) or np.array_equal(x, y)
        al = c["alive"] > 0
        stu = al & (c["has_school"] > 0)
        assert (sl[stu] >= 1).all() and (sl[stu] <= 4).all()
        lo = np.array([0, 6, 11, 14, 19])[sl[stu]]
        hi = np.array([0, 10, 13, 18, 25])[sl[stu]]
        # the setup puts students of every age of a level's range into it; afterwards each leaves at escape age + 1
        assert (c["age"][stu] >= lo).all() and (c["age"][stu] <= hi).all()
        assert (sl[stu] <= np.maximum(me[stu], 1)).all()
        changed = al & (c["education_level"] != before["education_level"])
        assert (c["education_level"][changed] == sl0[changed]).all() and changed.sum() == ngrad
        assert (c["education_level"][al] >= before["education_level"][al]).all()
        drew = al & (before["job_level"] < 2) & ((t - c["birth_tick"]) % 12 == 0) & np.isin(c["age"], [15, 25, 35, 45, 55]) & ~changed
        assert ndraw >= drew.sum() and (c["job_level"][drew] < 2).all()
    assert n_grad_total > 100
