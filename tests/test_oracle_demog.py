"""The CPU restatement of the SURVEY.md 8(f) phases (oracle/hotpath.c: orc_make_baby, orc_make_friends,
orc_make_people_die) against the UNMODIFIED reference, in distribution: tests/golden/distdemog_n1000_s42.npz holds 30
continuations of one reference setup state by the reference with exactly these phases and the crime path active
(oracle/make_states.py distdemog).  The kernels are then checked bit for bit against this restatement (-m gpu)."""
import numpy as np
from scipy.stats import ks_2samp

import util as U
from oracle import oracle_c as O


def test_demography_phases_in_distribution_against_the_reference():
    fx = U.load_fixture(U.GOLDEN + "/distdemog_n1000_s42.npz")
    st, prm = U.pre_state(fx), U.fixture_params(fx)
    runs, ticks = fx["refdemog_number_crimes"].shape
    P, D = O.make_params(prm), O.make_demog()
    names = ["current_num_persons", "tot_friendship_link", "tot_household_link", "tot_sibling_link", "tot_offspring_link",
             "number_deceased", "number_born", "number_crimes", "tot_criminal_link", "current_oc_members", "facilitators"]
    layer_of = {"tot_friendship_link": 5, "tot_household_link": 4, "tot_sibling_link": 0, "tot_offspring_link": 1, "tot_criminal_link": 6}
    checks = (5, 23, ticks - 1)
    cols = {k: np.zeros((runs, len(checks))) for k in names}
    for j in range(runs):
        S = O.OracleState(st)
        S.reserve(len(st["alive"]) + 300)
        S.set_children(st["number_of_children"])
        mult, crimes, t_done = 0.0, 0, 0
        for ci, t_end in enumerate(checks):
            mult, tot = S.run_ticks_phases(P, D, 7, t_done + 1, t_end + 1 - t_done, 7000 + j, mult)
            crimes += int(tot[0])
            t_done = t_end + 1
            c = S.cols()
            al = c["alive"] > 0
            dc = S.demog_counters()
            vals = {"current_num_persons": al.sum(), "number_deceased": dc["number_deceased"], "number_born": dc["number_born"],
                    "number_crimes": crimes, "current_oc_members": c["oc_member"][al].sum(), "facilitators": c["facilitator"][al].sum()}
            for k, l in layer_of.items():
                vals[k] = np.diff(S.layer(l)[0])[al].sum()
            for k in names:
                cols[k][j, ci] = vals[k]
    for k in names:
        ref = fx["refdemog_" + k]
        for ci, t in enumerate(checks):
            p = ks_2samp(cols[k][:, ci], ref[:, t]).pvalue
            assert p > 0.005, (k, t, p, cols[k][:, ci].mean(), ref[:, t].mean())
        if ref[:, -1].mean() > 20:
            assert abs(cols[k][:, -1].mean() - ref[:, -1].mean()) <= 0.12 * ref[:, -1].mean(), k


def test_phase_invariants():
    """Size-independent properties: friendship stays symmetric, the dead lose their incoming links from the living but
    keep their own sets, a newborn is linked to mother, father, siblings and household, slots only ever grow."""
    fx = U.load_fixture(U.GOLDEN + "/state_n3000_s42_baseline.npz")
    st = U.pre_state(fx)
    P, D = O.make_params(U.fixture_params(fx)), O.make_demog()
    S = O.OracleState(st)
    S.reserve(len(st["alive"]) + 500)
    n0 = S.n
    for t in range(1, 25):
        S.begin_tick(t)
        alive_before = S.cols()["alive"].copy()
        nb = S.make_baby(P, D, t, 5)
        c = S.cols()
        for b in range(S.n - nb, S.n):
            par = S.layer(2)
            parents = par[1][par[0][b]:par[0][b + 1]]
            assert 1 <= len(parents) <= 2 and c["birth_tick"][b] == t and c["alive"][b] == 1
            assert c["father"][b] in (-1, *parents.tolist())
        S.make_friends(P, D, t, 5)
        rp, col, _ = S.layer(5)
        src = np.repeat(np.arange(S.n), np.diff(rp))
        pairs = set(zip(src.tolist(), col.tolist()))
        nd = S.make_people_die(P, D, t, 5)
        c2 = S.cols()
        dead_now = np.flatnonzero((alive_before[:len(c2["alive"])] > 0) & (c2["alive"][:len(alive_before)] == 0))
        assert len(dead_now) == nd
        for l in range(9):
            rp, col, _ = S.layer(l)
            src = np.repeat(np.arange(S.n), np.diff(rp))
            living_src = c2["alive"][src] > 0
            # nobody alive keeps a link to somebody who died this tick and had them as a neighbour
            hit = np.isin(col, dead_now) & living_src
            for u, v in zip(src[hit].tolist(), col[hit].tolist()):
                rpv = S.layer(l)[0]
                out_of_dead = set()
                for l2 in range(9):
                    r2, c2l, _ = S.layer(l2)
                    out_of_dead.update(c2l[r2[v]:r2[v + 1]].tolist())
                assert u not in out_of_dead, (t, l, u, v)
    assert S.n >= n0 and S.demog_counters()["number_born"] == S.n - n0
    # friendship symmetric among pairs created by make_friends (one-sided links exist only where get_caught cleared a side)
    assert len(pairs) > 0


def test_six_phases_in_distribution_against_the_reference():
    """retire_persons, remove_excess_friends, remove_excess_professional_links added to the three above: the oracle's runs
    against 30 continuations by the unmodified reference with exactly these six phases and the crime path active
    (oracle/make_states.py distdemog2), KS per column at three ticks."""
    fx = U.load_fixture(U.GOLDEN + "/distdemog2_n1000_s42.npz")
    st, prm = U.pre_state(fx), U.fixture_params(fx)
    runs, ticks = fx["refdemog_number_crimes"].shape
    P, D = O.make_params(prm), O.make_demog()
    layer_of = {"tot_friendship_link": 5, "tot_household_link": 4, "tot_professional_link": 7, "tot_criminal_link": 6}
    names = ["current_num_persons", "employed", "number_crimes", "current_oc_members"] + list(layer_of)
    checks = (5, 23, ticks - 1)
    cols = {k: np.zeros((runs, len(checks))) for k in names}
    for j in range(runs):
        S = O.OracleState(st)
        S.reserve(len(st["alive"]) + 300)
        S.set_children(st["number_of_children"])
        mult, crimes, t_done = 0.0, 0, 0
        for ci, t_end in enumerate(checks):
            mult, tot = S.run_ticks_phases(P, D, 63, t_done + 1, t_end + 1 - t_done, 7100 + j, mult)
            crimes += int(tot[0])
            t_done = t_end + 1
            c = S.cols()
            al = c["alive"] > 0
            vals = {"current_num_persons": al.sum(), "employed": c["has_job"][al].sum(), "number_crimes": crimes,
                    "current_oc_members": c["oc_member"][al].sum()}
            for k, l in layer_of.items():
                vals[k] = np.diff(S.layer(l)[0])[al].sum()
            for k in names:
                cols[k][j, ci] = vals[k]
    for k in names:
        ref = fx["refdemog_" + k]
        for ci, t in enumerate(checks):
            p = ks_2samp(cols[k][:, ci], ref[:, t]).pvalue
            assert p > 0.005, (k, t, p, cols[k][:, ci].mean(), ref[:, t].mean())
        if ref[:, -1].mean() > 20:
            assert abs(cols[k][:, -1].mean() - ref[:, -1].mean()) <= 0.12 * ref[:, -1].mean(), (k, cols[k][:, -1].mean(), ref[:, -1].mean())


def _with_hub(st, layer, hub, members):
    """the state with `hub` linked to every agent of `members` in `layer` (both directions)"""
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


def test_retire_and_remove_excess_invariants():
    """retire_persons: nobody of retirement age keeps a job or an own professional set, the colleagues keep theirs (one-sided,
    model.py:1550); remove_excess_*: afterwards no living agent is over its limit, exactly the surplus went where nobody else
    cut, and the layer stays symmetric."""
    fx = U.load_fixture(U.GOLDEN + "/state_n3000_s42_baseline.npz")
    st = U.pre_state(fx)
    n = len(st["alive"])
    alive = np.flatnonzero(st["alive"] > 0)
    hub_f, hub_p = int(alive[10]), int(alive[20])
    st = _with_hub(st, 5, hub_f, [int(x) for x in alive[100:400]])     # 300 friends: over any Dunbar number
    st = _with_hub(st, 7, hub_p, [int(x) for x in alive[500:580]])     # 80 colleagues: over 30
    D = O.make_demog()
    S = O.OracleState(st)
    S.begin_tick(12)   # the first birthday tick of a setup state (birth_tick = -12 * age, entities.py:271): the 64-year-olds turn 65
    c0 = S.cols()
    prof0 = S.layer(7)
    old = np.flatnonzero((c0["alive"] > 0) & (c0["age"] >= 65) & (c0["has_job"] > 0))
    nr = S.retire_persons(D)
    assert nr == len(old) and nr > 0
    c1, prof1 = S.cols(), S.layer(7)
    assert not c1["has_job"][old].any() and (np.diff(prof1[0])[old] == 0).all()
    src = np.repeat(np.arange(n), np.diff(prof1[0]))
    still = np.isin(prof1[1], old) & ~np.isin(src, old)
    assert still.sum() == (np.isin(prof0[1], old) & ~np.isin(np.repeat(np.arange(n), np.diff(prof0[0])), old)).sum()   # colleagues untouched
    assert S.retire_persons(D) == 0   # idempotent
    for layer, fn, lim in ((5, S.remove_excess_friends, lambda age: 150 - abs(int(age) - 30)), (7, S.remove_excess_professional, lambda age: 30)):
        before = np.diff(S.layer(layer)[0])
        over = [a for a in alive.tolist() if before[a] > lim(c1["age"][a])]
        assert over, layer
        nrem = fn(1, 99)
        rp, col, _ = S.layer(layer)
        after = np.diff(rp)
        assert nrem == 2 * sum(before[a] - lim(c1["age"][a]) for a in over) or len(over) > 1
        for a in alive.tolist():
            assert after[a] <= max(lim(c1["age"][a]), 0) or before[a] <= lim(c1["age"][a]), (layer, a, after[a])
        srcs = np.repeat(np.arange(n), after)
        pairs = set(zip(srcs.tolist(), col.tolist()))
        hub = hub_f if layer == 5 else hub_p
        assert all((v, u) in pairs for (u, v) in pairs if u == hub)
        assert fn(2, 99) == 0   # nobody is over the limit any more


def test_wedding_in_distribution_against_the_reference_and_invariants():
    """orc_wedding (model.py:368-402 restated): 30 oracle runs with all seven phases against 30 continuations by the unmodified
    reference with exactly these phases active (distdemog3): number_weddings and the partner / household totals, KS at three
    ticks.  Invariants of every wedding: both were marriable (25 < age < 55, no partner), of different sex, linked as partner
    and household both ways afterwards, and no longer in the household of anybody they left."""
    fx = U.load_fixture(U.GOLDEN + "/distdemog3_n1000_s42.npz")
    st, prm = U.pre_state(fx), U.fixture_params(fx)
    runs, ticks = fx["refdemog_number_crimes"].shape
    P, D = O.make_params(prm), O.make_demog()
    names = ["number_weddings", "tot_partner_link", "tot_household_link", "current_num_persons", "number_crimes"]
    checks = (5, 23, ticks - 1)
    cols = {k: np.zeros((runs, len(checks))) for k in names}
    for j in range(runs):
        S = O.OracleState(st)
        S.reserve(len(st["alive"]) + 300)
        S.set_children(st["number_of_children"])
        mult, crimes, t_done = 0.0, 0, 0
        for ci, t_end in enumerate(checks):
            mult, tot = S.run_ticks_phases(P, D, 127, t_done + 1, t_end + 1 - t_done, 7300 + j, mult)
            crimes += int(tot[0])
            t_done = t_end + 1
            c = S.cols()
            al = c["alive"] > 0
            vals = {"number_weddings": S.demog_counters()["number_weddings"], "tot_partner_link": np.diff(S.layer(3)[0])[al].sum(),
                    "tot_household_link": np.diff(S.layer(4)[0])[al].sum(), "current_num_persons": al.sum(), "number_crimes": crimes}
            for k in names:
                cols[k][j, ci] = vals[k]
    for k in names:
        ref = fx["refdemog_" + k]
        for ci, t in enumerate(checks):
            p = ks_2samp(cols[k][:, ci], ref[:, t]).pvalue
            assert p > 0.005, (k, t, p, cols[k][:, ci].mean(), ref[:, t].mean())
    # one phase call on the 3,000-agent setup state, looked at closely
    fx = U.load_fixture(U.GOLDEN + "/state_n3000_s42_baseline.npz")
    st = U.pre_state(fx)
    n = len(st["alive"])
    S = O.OracleState(st)
    total = 0
    for t in range(1, 13):
        S.begin_tick(t)
        c0, part0, hh0 = S.cols(), S.layer(3), S.layer(4)
        nw = S.wedding(O.make_params(U.fixture_params(fx)), D, t, 31)
        total += nw
        part1, hh1 = S.layer(3), S.layer(4)
        newly = [a for a in range(n) if part1[0][a + 1] - part1[0][a] > part0[0][a + 1] - part0[0][a]]
        assert len(newly) == 2 * nw
        for a in newly:
            b = int(part1[1][part1[0][a]])
            assert part0[0][a + 1] == part0[0][a] and 25 < c0["age"][a] < 55 and c0["male"][a] != c0["male"][b]
            assert b in part1[1][part1[0][b]:part1[0][b + 1]] or a in part1[1][part1[0][b]:part1[0][b + 1]]
            hh_a = set(hh1[1][hh1[0][a]:hh1[0][a + 1]].tolist())
            assert b in hh_a
            for m in hh0[1][hh0[0][a]:hh0[0][a + 1]].tolist():   # the household it left (reciprocal links only)
                if m != b and a in hh0[1][hh0[0][m]:hh0[0][m + 1]].tolist():
                    assert m not in hh_a and a not in hh1[1][hh1[0][m]:hh1[0][m + 1]].tolist()
    assert total > 0
