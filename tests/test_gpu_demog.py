"""SURVEY.md 8(f) rows on the device: make_baby, make_friends, make_people_die (poc_demog.cuh).
  * bit for bit against the CPU restatement of the same batch semantics (oracle/hotpath.c), phase by phase and over
    whole runs with the crime path in between;
  * in distribution against the reference itself: tests/golden/distdemog_* holds 30 continuations of one reference
    setup state by the UNMODIFIED reference with exactly these phases (and the crime path) active."""
import numpy as np
import pytest

import util as U

pytestmark = pytest.mark.gpu

COLS = ["alive", "male", "oc_member", "facilitator", "prisoner", "has_job", "has_school", "target_of_intervention",
        "job_level", "education_level", "wealth_level", "age", "birth_tick", "num_crimes_committed",
        "num_crimes_committed_this_tick", "father", "new_recruit", "propensity", "criminal_tendency", "sentence_countdown"]


@pytest.fixture(scope="module")
def P():
    import pyproton_oc_b200 as pkg
    return pkg


@pytest.fixture(scope="module")
def O():
    from oracle import oracle_c
    return oracle_c


def engine_for(P, states, extra_agents=600, extra_static=40000, crim_headroom=60000):
    n, stat, crim = P.CrimeEngine.capacity_for(states, crim_headroom)
    return P.CrimeEngine(len(states), n + extra_agents, stat + extra_static, crim)


def children_of(st):
    return st["number_of_children"] if "number_of_children" in st else np.diff(st["rowptr_1"]).astype(np.int32)


def assert_equal(E, r, S, what):
    got, oc = E.download_state(r), S.cols()
    assert len(got["alive"]) == len(oc["alive"]), what + ": agent slots"
    for k in COLS:
        assert np.array_equal(got[k], oc[k]), "%s: column %s" % (what, k)
    for l in range(9):
        rp, col, co = S.layer(l)
        assert np.array_equal(got["rowptr_%d" % l], rp), "%s: layer %d rowptr" % (what, l)
        assert np.array_equal(got["col_%d" % l], col), "%s: layer %d col" % (what, l)
        if l == 6:
            assert np.array_equal(got["co_off"], co), what
    assert np.array_equal(E.download_children(r), S.children()), what + ": number_of_children"


@pytest.mark.parametrize("name", ["tick_n1000_s42_t12", "tick_n1000_s5_late_t252", "tick_n1000_s11_oc200_t15",
                                  "tick_n3000_s42_t24", "tick_n1000_s7_disruptive_strong_t25"])
def test_each_phase_matches_the_oracle(P, O, name):
    """One phase at a time on reference states (late ones carry one-sided links and dead-but-referenced agents): the
    whole state -- columns, all nine layers, children counts, new slots -- equals the oracle's after every phase."""
    fx = U.load_fixture(U.GOLDEN + "/%s.npz" % name)
    st, prm = U.pre_state(fx), U.fixture_params(fx)
    tick, key = int(fx["tick"]), 424242
    S = O.OracleState(st)
    S.reserve(len(st["alive"]) + 600)
    S.set_children(children_of(st))
    OP, OD = O.make_params(prm), O.make_demog()
    with engine_for(P, [st]) as E:
        E.set_params(P.make_params(prm))
        E.set_demography(P.make_demography())
        E.upload_state(0, st)
        E.upload_children(0, children_of(st))
        for rep in range(6):   # several rounds so that the phases see each other's edits
            t = tick + rep
            S.begin_tick(t); E.begin_tick(t)
            nb = S.make_baby(OP, OD, t, key); E.run_phase(E.PHASE_BABY, t, [key])
            assert_equal(E, 0, S, "%s make_baby round %d (%d births)" % (name, rep, nb))
            nf = S.make_friends(OP, OD, t, key); E.run_phase(E.PHASE_FRIENDS, t, [key])
            assert_equal(E, 0, S, "%s make_friends round %d (%d friendships)" % (name, rep, nf))
            nd = S.make_people_die(OP, OD, t, key); E.run_phase(E.PHASE_DIE, t, [key])
            assert_equal(E, 0, S, "%s make_people_die round %d (%d deaths)" % (name, rep, nd))
        cn, _ = E.counters()
        assert cn[0, 9] == 0
        dc = S.demog_counters()
        assert cn[0, 16] == dc["number_deceased"] and cn[0, 17] == dc["number_born"]
        # the tot_<layer>_link reporters (kept incrementally on the device) equal the layer sizes over the living
        row = dict(zip(P.REPORTER_NAMES, E.report()[0]))
        al = S.cols()["alive"] > 0
        for l, lname in enumerate(P.LAYER_NAMES):
            rp, _, _ = S.layer(l)
            assert row["tot_%s_link" % lname] == int(np.diff(rp)[al].sum()), lname
        assert row["current_num_persons"] == int(al.sum())


@pytest.mark.parametrize("name,ticks", [("tick_n1000_s42_t12", 60), ("tick_n1000_s11_oc200_t3", 36), ("tick_n3000_s42_t1", 30)])
def test_runs_with_all_phases_match_the_oracle(P, O, name, ticks):
    """poc_run_ticks with the three phases switched on (crime path, make_baby, make_friends, prisoner countdown,
    make_people_die per tick, one CUDA graph) equals the oracle's run tick for tick, two replicas with different keys."""
    fx = U.load_fixture(U.GOLDEN + "/%s.npz" % name)
    st, prm = U.pre_state(fx), U.fixture_params(fx)
    tick0, keys = int(fx["tick"]), [77, 78]
    mult0 = float(fx["crime_multiplier"])
    with engine_for(P, [st, st]) as E:
        E.set_params(P.make_params(prm))
        E.set_demography(P.make_demography())
        E.set_phases(7)
        for r in range(2):
            E.upload_state(r, st)
            E.upload_children(r, children_of(st))
            E.set_crime_multiplier(mult0, r)
        rep = E.run_ticks(tick0, ticks, keys)
        cn, _ = E.counters()
        assert (cn[:, 9] == 0).all()
        for r in range(2):
            S = O.OracleState(st)
            S.reserve(len(st["alive"]) + 600)
            S.set_children(children_of(st))
            mult, tot = S.run_ticks_phases(O.make_params(prm), O.make_demog(), 7, tick0, ticks, keys[r], mult0)
            assert_equal(E, r, S, "%s replica %d after %d ticks" % (name, r, ticks))
            assert rep[r, -1, 0] == tot[0] and rep[r, -1, 15] == mult
            al = S.cols()["alive"] > 0
            assert rep[r, -1, 14] == int(al.sum())
            assert rep[r, -1, 30] == int(np.diff(S.layer(5)[0])[al].sum())   # tot_friendship_link


def _with_hub(st, layer, hub, members):
    st = {k: v.copy() for k, v in st.items()}
    n = len(st["alive"])
    rp, col = st["rowptr_%d" % layer], st["col_%d" % layer]
    src = np.repeat(np.arange(n), np.diff(rp))
    pairs = set(zip(src.tolist(), col.tolist())) | {(hub, o) for o in members} | {(o, hub) for o in members}
    key = np.array(sorted(a * n + b for a, b in pairs), dtype=np.int64)
    newrp = np.zeros(n + 1, np.int64)
    np.add.at(newrp, key // n + 1, 1)
    st["rowptr_%d" % layer], st["col_%d" % layer] = np.cumsum(newrp).astype(np.int32), (key % n).astype(np.int32)
    return st


@pytest.mark.parametrize("name", ["tick_n1000_s42_t12", "tick_n1000_s5_late_t252", "tick_n3000_s42_t24",
                                  "tick_n1000_s7_disruptive_strong_t25"])
def test_retire_and_remove_excess_match_the_oracle(P, O, name):
    """retire_persons, remove_excess_friends, remove_excess_professional_links one at a time, interleaved with the crime path
    (which reads the one-sided flags and upper-half pointers these phases have to keep) and the other phases: the whole state
    equals the oracle's after every phase.  Two hubs put agents over both limits."""
    fx = U.load_fixture(U.GOLDEN + "/%s.npz" % name)
    st, prm = U.pre_state(fx), U.fixture_params(fx)
    alive = np.flatnonzero(st["alive"] > 0)
    st = _with_hub(st, 5, int(alive[3]), [int(x) for x in alive[50:300]])
    st = _with_hub(st, 7, int(alive[7]), [int(x) for x in alive[320:400]])
    tick, key = int(fx["tick"]), 515151
    S = O.OracleState(st)
    S.reserve(len(st["alive"]) + 600)
    S.set_children(children_of(st))
    OP, OD = O.make_params(prm), O.make_demog()
    with engine_for(P, [st]) as E:
        E.set_params(P.make_params(prm))
        E.set_demography(P.make_demography())
        E.upload_state(0, st)
        E.upload_children(0, children_of(st))
        for rep in range(4):
            t = tick + rep
            S.begin_tick(t); E.begin_tick(t)
            nr = S.retire_persons(OD); E.run_phase(E.PHASE_RETIRE, t, [key])
            assert_equal(E, 0, S, "%s retire_persons round %d (%d retired)" % (name, rep, nr))
            nb = S.make_baby(OP, OD, t, key); E.run_phase(E.PHASE_BABY, t, [key])
            assert_equal(E, 0, S, "%s make_baby round %d (%d births)" % (name, rep, nb))
            nx = S.remove_excess_friends(t, key); E.run_phase(E.PHASE_XFRIENDS, t, [key])
            assert_equal(E, 0, S, "%s remove_excess_friends round %d (%d removals)" % (name, rep, nx))
            assert rep > 0 or nx > 0
            np_ = S.remove_excess_professional(t, key); E.run_phase(E.PHASE_XPROF, t, [key])
            assert_equal(E, 0, S, "%s remove_excess_professional_links round %d (%d removals)" % (name, rep, np_))
            assert rep > 0 or np_ > 0
            S.make_friends(OP, OD, t, key); E.run_phase(E.PHASE_FRIENDS, t, [key])
            assert_equal(E, 0, S, "%s make_friends round %d" % (name, rep))
            S.make_people_die(OP, OD, t, key); E.run_phase(E.PHASE_DIE, t, [key])
            assert_equal(E, 0, S, "%s make_people_die round %d" % (name, rep))
        cn, _ = E.counters()
        assert cn[0, 9] == 0
        row = dict(zip(P.REPORTER_NAMES, E.report()[0]))
        c = S.cols()
        al = c["alive"] > 0
        for l, lname in enumerate(P.LAYER_NAMES):
            rp, _, _ = S.layer(l)
            assert row["tot_%s_link" % lname] == int(np.diff(rp)[al].sum()), lname
        assert row["employed"] == int(c["has_job"][al].sum())


@pytest.mark.parametrize("name,ticks", [("tick_n1000_s42_t12", 60), ("tick_n1000_s11_oc200_t3", 36), ("tick_n3000_s42_t1", 30)])
def test_runs_with_six_phases_match_the_oracle(P, O, name, ticks):
    """poc_run_ticks with six phases on (all but wedding) equals the oracle's run: the crime path's traversals (embeddedness
    reads every undirected link once, through the one-sided flags) see the same network tick after tick."""
    fx = U.load_fixture(U.GOLDEN + "/%s.npz" % name)
    st, prm = U.pre_state(fx), U.fixture_params(fx)
    alive = np.flatnonzero(st["alive"] > 0)
    st = _with_hub(st, 5, int(alive[3]), [int(x) for x in alive[50:300]])
    st = _with_hub(st, 7, int(alive[7]), [int(x) for x in alive[320:400]])
    tick0, keys = int(fx["tick"]), [91, 92]
    mult0 = float(fx["crime_multiplier"])
    with engine_for(P, [st, st]) as E:
        E.set_params(P.make_params(prm))
        E.set_demography(P.make_demography())
        E.set_phases(63)   # every phase but wedding (the fixture / oracle call of this test)
        for r in range(2):
            E.upload_state(r, st)
            E.upload_children(r, children_of(st))
            E.set_crime_multiplier(mult0, r)
        rep = E.run_ticks(tick0, ticks, keys)
        cn, _ = E.counters()
        assert (cn[:, 9] == 0).all()
        for r in range(2):
            S = O.OracleState(st)
            S.reserve(len(st["alive"]) + 600)
            S.set_children(children_of(st))
            mult, tot = S.run_ticks_phases(O.make_params(prm), O.make_demog(), 63, tick0, ticks, keys[r], mult0)
            assert_equal(E, r, S, "%s replica %d after %d ticks" % (name, r, ticks))
            assert rep[r, -1, 0] == tot[0] and rep[r, -1, 15] == mult
            c = S.cols()
            al = c["alive"] > 0
            assert rep[r, -1, 22] == int(c["has_job"][al].sum())                   # employed
            assert rep[r, -1, 31] == int(np.diff(S.layer(7)[0])[al].sum())         # tot_professional_link


@pytest.mark.parametrize("name", ["tick_n1000_s42_t12", "tick_n1000_s5_late_t252", "tick_n3000_s42_t24",
                                  "tick_n1000_s7_disruptive_strong_t25", "state_n3000_s42_baseline"])
def test_wedding_matches_the_oracle(P, O, name):
    """wedding (model.py:368-402) on the device against the oracle's restatement, phase by phase, interleaved with the crime
    path's state changes of the other phases: whole state equal after every call (the couples' new household / partner links
    go through the insert pass, the households they leave are cleared in place)."""
    fx = U.load_fixture(U.GOLDEN + "/%s.npz" % name)
    st, prm = U.pre_state(fx), U.fixture_params(fx)
    tick, key = int(fx["tick"]), 616161
    S = O.OracleState(st)
    S.reserve(len(st["alive"]) + 600)
    S.set_children(children_of(st))
    OP = O.make_params(prm)
    # a high marriage rate so that every tick has several weddings (and egos that find nobody)
    OD, PD = O.make_demog(), P.make_demography(number_weddings_mean=60.0)
    OD.weddings_mean = 60.0
    total = 0
    with engine_for(P, [st]) as E:
        E.set_params(P.make_params(prm))
        E.set_demography(PD)
        E.upload_state(0, st)
        E.upload_children(0, children_of(st))
        for rep in range(6):
            t = tick + rep
            S.begin_tick(t); E.begin_tick(t)
            nw = S.wedding(OP, OD, t, key); E.run_phase(E.PHASE_WEDDING, t, [key])
            total += nw
            assert_equal(E, 0, S, "%s wedding round %d (%d weddings)" % (name, rep, nw))
            S.make_baby(OP, OD, t, key); E.run_phase(E.PHASE_BABY, t, [key])
            S.make_friends(OP, OD, t, key); E.run_phase(E.PHASE_FRIENDS, t, [key])
            S.make_people_die(OP, OD, t, key); E.run_phase(E.PHASE_DIE, t, [key])
            assert_equal(E, 0, S, "%s after the other phases, round %d" % (name, rep))
        cn, _ = E.counters()
        assert cn[0, 9] == 0 and cn[0, 19] == S.demog_counters()["number_weddings"] == total and total > 0
        row = dict(zip(P.REPORTER_NAMES, E.report()[0]))
        al = S.cols()["alive"] > 0
        for l, lname in enumerate(P.LAYER_NAMES):
            rp, _, _ = S.layer(l)
            assert row["tot_%s_link" % lname] == int(np.diff(rp)[al].sum()), lname
        assert row["number_weddings"] == total


@pytest.mark.parametrize("name,ticks", [("tick_n1000_s42_t12", 60), ("tick_n3000_s42_t1", 30)])
def test_runs_with_all_seven_phases_match_the_oracle(P, O, name, ticks):
    fx = U.load_fixture(U.GOLDEN + "/%s.npz" % name)
    st, prm = U.pre_state(fx), U.fixture_params(fx)
    tick0, keys = int(fx["tick"]), [191, 192]
    mult0 = float(fx["crime_multiplier"])
    with engine_for(P, [st, st]) as E:
        E.set_params(P.make_params(prm))
        E.set_demography(P.make_demography())
        E.set_phases(E.PHASE_ALL)
        for r in range(2):
            E.upload_state(r, st)
            E.upload_children(r, children_of(st))
            E.set_crime_multiplier(mult0, r)
        rep = E.run_ticks(tick0, ticks, keys)
        cn, _ = E.counters()
        assert (cn[:, 9] == 0).all()
        for r in range(2):
            S = O.OracleState(st)
            S.reserve(len(st["alive"]) + 600)
            S.set_children(children_of(st))
            mult, tot = S.run_ticks_phases(O.make_params(prm), O.make_demog(), 127, tick0, ticks, keys[r], mult0)
            assert_equal(E, r, S, "%s replica %d after %d ticks" % (name, r, ticks))
            assert rep[r, -1, 0] == tot[0] and rep[r, -1, 42] == S.demog_counters()["number_weddings"]


def test_seven_phases_in_distribution_against_the_reference(P):
    """30 replicas against 30 continuations by the unmodified reference with wedding and the six other phases active
    (distdemog3)."""
    from scipy.stats import ks_2samp
    fx = U.load_fixture(U.GOLDEN + "/distdemog3_n1000_s42.npz")
    st, prm = U.pre_state(fx), U.fixture_params(fx)
    runs, ticks = fx["refdemog_number_crimes"].shape
    n, stat, crim = P.CrimeEngine.capacity_for([st], 60000)
    with P.CrimeEngine(runs, n + 300, stat + 20000, crim) as E:
        E.set_params(P.make_params(prm))
        E.set_demography(P.make_demography())
        E.set_phases(E.PHASE_ALL)
        for r in range(runs):
            E.upload_state(r, st)
            E.upload_children(r, st["number_of_children"])
        rep = E.run_ticks(1, ticks, [9300 + r for r in range(runs)])
        cn, _ = E.counters()
    assert (cn[:, 9] == 0).all()
    names = list(P.REPORTER_NAMES)
    for col in ["number_weddings", "tot_partner_link", "tot_household_link", "current_num_persons", "tot_friendship_link",
                "number_crimes", "current_oc_members"]:
        mine = rep[:, :, names.index(col)]
        ref = fx["refdemog_" + col]
        for t in (5, 23, ticks - 1):
            p = ks_2samp(mine[:, t], ref[:, t]).pvalue
            assert p > 0.005, (col, t, p, mine[:, t].mean(), ref[:, t].mean())


def test_six_phases_in_distribution_against_the_reference(P):
    """30 replicas against 30 continuations by the unmodified reference with the six phases active (distdemog2).  Nobody holds
    a job in this fixture (tests/golden/README.md): `employed` and `tot_professional_link` are 0 on both sides here; the runs
    with jobs are distdemog4 (tests/test_gpu_labour.py, tests/test_gpu_demog_large.py)."""
    from scipy.stats import ks_2samp
    fx = U.load_fixture(U.GOLDEN + "/distdemog2_n1000_s42.npz")
    st, prm = U.pre_state(fx), U.fixture_params(fx)
    runs, ticks = fx["refdemog_number_crimes"].shape
    n, stat, crim = P.CrimeEngine.capacity_for([st], 60000)
    with P.CrimeEngine(runs, n + 300, stat + 20000, crim) as E:
        E.set_params(P.make_params(prm))
        E.set_demography(P.make_demography())
        E.set_phases(63)   # every phase but wedding (the fixture / oracle call of this test)
        for r in range(runs):
            E.upload_state(r, st)
            E.upload_children(r, st["number_of_children"])
        rep = E.run_ticks(1, ticks, [9100 + r for r in range(runs)])
        cn, _ = E.counters()
    assert (cn[:, 9] == 0).all()
    names = list(P.REPORTER_NAMES)
    for col in ["current_num_persons", "tot_friendship_link", "tot_household_link", "tot_professional_link", "employed",
                "number_crimes", "tot_criminal_link", "current_oc_members"]:
        mine = rep[:, :, names.index(col)]
        ref = fx["refdemog_" + col]
        for t in (5, 23, ticks - 1):
            p = ks_2samp(mine[:, t], ref[:, t]).pvalue
            assert p > 0.005, (col, t, p, mine[:, t].mean(), ref[:, t].mean())
        if ref[:, -1].mean() > 20:
            assert abs(mine[:, -1].mean() - ref[:, -1].mean()) <= 0.12 * ref[:, -1].mean(), (col, mine[:, -1].mean(), ref[:, -1].mean())


def test_phases_in_distribution_against_the_reference(P):
    """30 replicas from the reference's 1,000-agent setup state against 30 continuations of that state by the
    unmodified reference with the same phases active (oracle/make_states.py distdemog): two-sample KS per column at
    three ticks, and the means within a few per cent."""
    from scipy.stats import ks_2samp
    fx = U.load_fixture(U.GOLDEN + "/distdemog_n1000_s42.npz")
    st, prm = U.pre_state(fx), U.fixture_params(fx)
    runs, ticks = fx["refdemog_number_crimes"].shape
    n, stat, crim = P.CrimeEngine.capacity_for([st], 60000)
    with P.CrimeEngine(runs, n + 300, stat + 20000, crim) as E:
        E.set_params(P.make_params(prm))
        E.set_demography(P.make_demography())
        E.set_phases(7)
        for r in range(runs):
            E.upload_state(r, st)
            E.upload_children(r, st["number_of_children"])
        rep = E.run_ticks(1, ticks, [9000 + r for r in range(runs)])
        cn, _ = E.counters()
    assert (cn[:, 9] == 0).all()
    names = list(P.REPORTER_NAMES)
    for col in ["current_num_persons", "tot_friendship_link", "tot_household_link", "tot_sibling_link", "tot_offspring_link",
                "tot_parent_link", "number_crimes", "tot_criminal_link", "current_oc_members"]:
        mine = rep[:, :, names.index(col)]
        ref = fx["refdemog_" + col]
        for t in (5, 23, ticks - 1):
            p = ks_2samp(mine[:, t], ref[:, t]).pvalue
            assert p > 0.005, (col, t, p, mine[:, t].mean(), ref[:, t].mean())
        if ref[:, -1].mean() > 20:
            assert abs(mine[:, -1].mean() - ref[:, -1].mean()) <= 0.12 * ref[:, -1].mean(), (col, mine[:, -1].mean(), ref[:, -1].mean())
    assert ks_2samp(cn[:, 16], fx["refdemog_number_deceased"][:, -1]).pvalue > 0.005
    assert ks_2samp(cn[:, 17], fx["refdemog_number_born"][:, -1]).pvalue > 0.005
