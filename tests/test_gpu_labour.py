"""The yearly labour-status block on the device (k_labour, poc_demog.cuh: graduate_and_enter_jobmarket, model.py:1351-1380, and
the update_unemployment_status loop of step(), model.py:252-256):
  * bit for bit against the CPU restatement (oracle/hotpath.c orc_labour_yearly), year by year with births in between (newborns
    inherit max_education_level and reach school age) and over whole runs with all eight device phases;
  * in distribution against the reference itself: tests/golden/distdemog4_n1000_s42.npz, 30 continuations of one setup state by
    the UNMODIFIED reference with exactly these phases active (tests/test_oracle_labour.py has the CPU half of this)."""
import numpy as np
import pytest

import util as U
from test_gpu_demog import assert_equal, children_of, engine_for

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    import pyproton_oc_b200 as pkg
    return pkg


@pytest.fixture(scope="module")
def O():
    from oracle import oracle_c
    return oracle_c


def assert_education_equal(E, r, S, what):
    me, sl = E.download_education(r)
    ome, osl = S.education()
    assert np.array_equal(me, ome), what + ": max_education_level"
    assert np.array_equal(sl, osl), what + ": school_level"


@pytest.mark.parametrize("name", ["distdemog4_n1000_s42", "tick_n1000_s5_late_t252", "tick_n3000_s42_t24"])
def test_labour_block_matches_the_oracle_year_by_year(P, O, name):
    """Nine yearly ticks; between them births (the newborns of the first years reach primary age), deaths and a crime tick (arrests
    take students out of school one-sidedly).  The older fixtures carry no education columns: everybody's max_education_level is
    0 there, i.e. every student leaves for the job market at the end of the school he is in."""
    fx = U.load_fixture(U.GOLDEN + "/%s.npz" % name)
    st, prm = U.pre_state(fx), U.fixture_params(fx)
    key = 616161
    S = O.OracleState(st)
    S.reserve(len(st["alive"]) + 600)
    S.set_children(children_of(st))
    S.set_labour(O.make_labour())
    OP, OD = O.make_params(prm), O.make_demog()
    mult = float(fx["crime_multiplier"])
    with engine_for(P, [st]) as E:
        E.set_params(P.make_params(prm))
        E.set_demography(P.make_demography())
        E.set_labour(P.make_labour())
        E.upload_state(0, st)
        E.upload_children(0, children_of(st))
        E.set_crime_multiplier(mult)
        if "max_education_level" in st:
            E.upload_education(0, st["max_education_level"], st["school_level"])
        t0 = int(fx["tick"])
        t0 += (-t0) % 12
        grads = draws = 0
        for year in range(9):
            t = t0 + 12 * year
            S.begin_tick(t); E.begin_tick(t)
            ng, nd = S.labour_yearly(OP, OD, t, key); E.run_phase(E.PHASE_LABOUR, t, [key])
            grads += ng; draws += nd
            assert_equal(E, 0, S, "%s labour block, year %d (%d graduations, %d status draws)" % (name, year, ng, nd))
            assert_education_equal(E, 0, S, "%s year %d" % (name, year))
            for k in range(3):   # births and deaths of three of the year's ticks
                S.make_baby(OP, OD, t + k, key); E.run_phase(E.PHASE_BABY, t + k, [key])
                S.make_people_die(OP, OD, t + k, key); E.run_phase(E.PHASE_DIE, t + k, [key])
            if year in (1, 4):    # crime ticks: get_caught clears has_school and leaves the school level behind
                mult, _ = S.run_ticks(OP, t + 3, 2, key, mult); E.run_ticks(t + 3, 2, [key])
            assert_equal(E, 0, S, "%s after the year's births and deaths, year %d" % (name, year))
            assert_education_equal(E, 0, S, "%s newborns, year %d" % (name, year))
        cn, _ = E.counters()
        assert cn[0, 9] == 0
        # (without education columns only the children who enrol at six during the run ever graduate)
        assert grads > (50 if "max_education_level" in st else 0) and draws > 20, (grads, draws)


def test_runs_with_all_eight_phases_match_the_oracle(P, O):
    fx = U.load_fixture(U.GOLDEN + "/distdemog4_n1000_s42.npz")
    st, prm = U.pre_state(fx), U.fixture_params(fx)
    ticks, keys = 75, [291, 292]
    mult0 = float(fx["crime_multiplier"])
    with engine_for(P, [st, st]) as E:
        E.set_params(P.make_params(prm))
        E.set_demography(P.make_demography())
        E.set_labour(P.make_labour())
        E.set_phases(E.PHASE_EVERY)
        for r in range(2):
            E.upload_state(r, st)
            E.upload_children(r, children_of(st))
            E.upload_education(r, st["max_education_level"], st["school_level"])
            E.set_crime_multiplier(mult0, r)
        rep = E.run_ticks(1, ticks, keys)
        cn, _ = E.counters()
        assert (cn[:, 9] == 0).all()
        names = list(P.REPORTER_NAMES)
        for r in range(2):
            S = O.OracleState(st)
            S.reserve(len(st["alive"]) + 600)
            S.set_children(children_of(st))
            S.set_labour(O.make_labour())
            mult, tot = S.run_ticks_phases(O.make_params(prm), O.make_demog(), 255, 1, ticks, keys[r], mult0)
            assert_equal(E, r, S, "replica %d after %d ticks with all eight phases" % (r, ticks))
            assert_education_equal(E, r, S, "replica %d" % r)
            c = S.cols()
            al = c["alive"] > 0
            assert rep[r, -1, 0] == tot[0]
            assert rep[r, -1, names.index("number_students")] == int((c["has_school"][al] > 0).sum())
            assert rep[r, -1, names.index("education_level_mean")] == pytest.approx(c["education_level"][al].mean(), rel=1e-12)
        # the labour block runs on the yearly ticks only, and it matters: the same run without it ends elsewhere
        E.set_phases(E.PHASE_ALL)
        E.upload_state(0, st)
        E.upload_children(0, children_of(st))
        E.upload_education(0, st["max_education_level"], st["school_level"])
        E.set_crime_multiplier(mult0, 0)
        rep7 = E.run_ticks(1, ticks, keys)
        assert rep7[0, -1, names.index("education_level_mean")] < rep[0, -1, names.index("education_level_mean")]


def test_labour_phase_needs_its_tables(P):
    fx = U.load_fixture(U.GOLDEN + "/tick_n100_s42_t12.npz")
    st = U.pre_state(fx)
    with engine_for(P, [st]) as E:
        E.set_params(P.make_params(U.fixture_params(fx)))
        E.set_demography(P.make_demography())
        E.upload_state(0, st)
        with pytest.raises(P.PocError):
            E.set_phases(E.PHASE_EVERY)
        with pytest.raises(P.PocError):
            E.set_phases(512)
        lb = P.make_labour()
        lb.work_val[2][1][0] = 9   # a work status that would index past the wealth table
        with pytest.raises(P.PocError):
            E.set_labour(lb)
        E.set_labour(P.make_labour())
        E.set_phases(E.PHASE_EVERY)
        me, sl = E.download_education(0)
        assert not me.any() and not sl.any()   # never uploaded: zeros
        with pytest.raises(P.PocError):
            E.upload_education(0, np.full(len(me), 16, np.int8), sl)


def test_eight_phases_in_distribution_against_the_reference(P):
    """30 replicas against 30 continuations by the unmodified reference with graduate_and_enter_jobmarket, its own
    update_unemployment_status loop and the seven per-tick phases active (distdemog4)."""
    from scipy.stats import ks_2samp
    fx = U.load_fixture(U.GOLDEN + "/distdemog4_n1000_s42.npz")
    st, prm = U.pre_state(fx), U.fixture_params(fx)
    runs, ticks = fx["refdemog_number_crimes"].shape
    n, stat, crim = P.CrimeEngine.capacity_for([st], 60000)
    with P.CrimeEngine(runs, n + 400, stat + 24000, crim) as E:
        E.set_params(P.make_params(prm))
        E.set_demography(P.make_demography())
        E.set_labour(P.make_labour())
        E.set_phases(E.PHASE_EVERY)
        for r in range(runs):
            E.upload_state(r, st)
            E.upload_children(r, st["number_of_children"])
            E.upload_education(r, st["max_education_level"], st["school_level"])
        rep = E.run_ticks(1, ticks, [9400 + r for r in range(runs)])
        cn, _ = E.counters()
        hists = []
        for r in range(runs):
            c = E.download_agents(r)
            al = c["alive"] > 0
            sl = E.download_education(r)[1]
            hists.append([np.bincount(c["job_level"][al], minlength=8)[:8], np.bincount(c["education_level"][al], minlength=8)[:8],
                          np.bincount(c["wealth_level"][al], minlength=8)[:8],
                          np.bincount(np.where(c["has_school"][al] > 0, sl[al], 0), minlength=8)[:8]])
    assert (cn[:, 9] == 0).all()
    names = list(P.REPORTER_NAMES)
    for col in ["education_level_mean", "current_num_persons", "tot_friendship_link", "number_crimes", "current_oc_members",
                "number_weddings", "employed", "tot_professional_link"]:   # the last two: retire_persons, remove_excess_professional_links
        mine = rep[:, :, names.index(col)]
        ref = fx["refdemog_" + col]
        for t in (11, 35, ticks - 1):
            p = ks_2samp(mine[:, t], ref[:, t]).pvalue
            assert p > 0.005, (col, t, p, mine[:, t].mean(), ref[:, t].mean())
    # (make_people_die takes the dead out of their school, model.py:1597-1598: number_students counts the living on both sides)
    stu = rep[:, -1, names.index("number_students")]
    assert abs(stu.mean() - fx["refdemog_number_students"][:, -1].mean()) <= 4.0
    mine_h, ref_h = np.array(hists, dtype=np.float64), fx["refdemog_final_hists"]
    for i, nm in enumerate(["job_level", "education_level", "wealth_level", "school_level"]):
        for v in range(8):
            a, b = mine_h[:, i, v], ref_h[:, i, v]
            if a.mean() < 3 and b.mean() < 3:
                continue
            se = np.sqrt(a.var() / len(a) + b.var() / len(b))
            assert abs(a.mean() - b.mean()) <= 4 * se + 0.03 * b.mean() + 1.0, (nm, v, a.mean(), b.mean(), se)


def test_full_length_eight_phase_run_of_the_bench_state_matches_the_oracle(P, O):
    """The run bench.py times as `with_step_phases`, one replica of it: the reference's 10,000-agent state at tick 13 (education
    columns from the _edu sidecar), 480 ticks with all eight device phases, against the CPU restatement -- whole state."""
    fx = U.load_fixture(U.GOLDEN + "/state_n10000_s42_t13.npz")
    st, prm = U.pre_state(fx), U.fixture_params(fx)
    st.update(U.pre_state(U.load_fixture(U.GOLDEN + "/state_n10000_s42_t13_edu.npz")))
    tick0, mult0, key, ticks = int(fx["tick"]), float(fx["crime_multiplier"]), 17, 480
    n, stat, crim = P.CrimeEngine.capacity_for([st], 160000)
    with P.CrimeEngine(1, n + 3000, stat + 100000, crim) as E:
        E.set_params(P.make_params(prm))
        E.set_demography(P.make_demography())
        E.set_labour(P.make_labour())
        E.set_phases(E.PHASE_EVERY)
        E.upload_state(0, st)
        E.upload_education(0, st["max_education_level"], st["school_level"])
        E.set_crime_multiplier(mult0)
        rep = E.run_ticks(tick0, ticks, [key])
        cn, _ = E.counters()
        assert cn[0, 9] == 0
        S = O.OracleState(st)
        S.reserve(len(st["alive"]) + 3000)
        S.set_labour(O.make_labour())
        mult, tot = S.run_ticks_phases(O.make_params(prm), O.make_demog(), 255, tick0, ticks, key, mult0)
        assert_equal(E, 0, S, "10,000 agents, 480 ticks, eight phases")
        assert_education_equal(E, 0, S, "10,000 agents, 480 ticks")
        dc = S.demog_counters()
        assert rep[0, -1, 0] == tot[0] and cn[0, 16] == dc["number_deceased"] and cn[0, 17] == dc["number_born"]
        assert dc["number_deceased"] > 5000 and dc["number_born"] > 500   # four decades: most of the initial population is gone
