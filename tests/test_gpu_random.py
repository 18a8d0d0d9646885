"""Randomised GPU-vs-oracle parity on states the reference fixtures do not reach: asymmetric layers, dead agents
that are still referenced, dense OC cliques that overflow the staged-edge and shared-memory capacities of the
embeddedness kernel (its global-memory fall-back paths), large radii, big groups, all arrest regimes."""
import numpy as np
import pytest

import util as U

pytestmark = pytest.mark.gpu

COLS = ["oc_member", "prisoner", "has_job", "has_school", "job_level", "num_crimes_committed",
        "num_crimes_committed_this_tick", "new_recruit", "criminal_tendency", "sentence_countdown", "age"]


def random_state(rng, n, mean_deg, oc_frac, clique=0, asym=0.05, dead=0.02):
    pairs = {l: [] for l in U.LAYERS}
    directed = {"offspring": []}
    for l, share in [("friendship", 0.45), ("household", 0.15), ("professional", 0.12), ("school", 0.12),
                     ("sibling", 0.05), ("partner", 0.03)]:
        m = int(n * mean_deg * share / 2)
        a, b = rng.integers(0, n, m), rng.integers(0, n, m)
        pairs[l] = [(int(x), int(y)) for x, y in zip(a, b) if x != y]
    m = int(n * mean_deg * 0.04)
    directed["offspring"] = [(int(x), int(y)) for x, y in zip(rng.integers(0, n, m), rng.integers(0, n, m)) if x != y]
    oc = (rng.random(n) < oc_frac)
    members = np.flatnonzero(oc)
    if clique:
        members = rng.choice(n, size=clique, replace=False)
        oc[:] = False
        oc[members] = True
    crim = [(int(a), int(b)) for i, a in enumerate(members) for b in members[i + 1:]]
    extra = int(n * 0.05)
    crim += [(int(x), int(y)) for x, y in zip(rng.integers(0, n, extra), rng.integers(0, n, extra)) if x != y]
    pairs["criminal"] = crim
    co = {p: int(rng.integers(1, 6)) for p in crim}
    age = rng.integers(0, 90, n)
    st = U.build_state(
        n, pairs, directed, co_off=co, male=rng.random(n) < 0.5, age=age, birth_tick=-age * 12 - rng.integers(0, 12, n),
        education_level=rng.integers(0, 5, n), wealth_level=rng.integers(1, 6, n), job_level=rng.integers(0, 4, n),
        propensity=np.exp(1 + 0.25 * rng.standard_normal(n)), oc_member=oc, facilitator=(rng.random(n) < 0.05) & ~oc,
        alive=rng.random(n) >= dead, has_job=rng.random(n) < 0.4, has_school=rng.random(n) < 0.15,
        num_crimes_committed=rng.integers(0, 3, n), prisoner=rng.random(n) < 0.03,
        sentence_countdown=rng.integers(1, 24, n).astype(float), target_of_intervention=rng.random(n) < 0.1)
    st["sentence_countdown"] = np.where(st["prisoner"] > 0, st["sentence_countdown"], 0.0)
    # one-sided clears (get_caught / retire_persons, Q11): drop a few agents' own professional / school rows
    for l in (7, 8):
        rp, col = st["rowptr_%d" % l], st["col_%d" % l]
        keep = np.ones(len(col), bool)
        for a in np.flatnonzero(rng.random(n) < asym):
            keep[rp[a]:rp[a + 1]] = False
        src = np.repeat(np.arange(n), np.diff(rp))[keep]
        newrp = np.zeros(n + 1, np.int32)
        np.add.at(newrp, src + 1, 1)
        st["rowptr_%d" % l], st["col_%d" % l] = np.cumsum(newrp).astype(np.int32), col[keep]
    return st


def run_both(P, O, st, prm, ticks, key, co_offenders=None, crim_headroom=200000):
    S = O.OracleState(st)
    mult, totals = S.run_ticks(O.make_params(prm, co_offenders=co_offenders), 1, ticks, key, 0.0)
    n, stat, crim = P.CrimeEngine.capacity_for([st], crim_headroom)
    with P.CrimeEngine(1, n, stat, crim) as E:
        E.set_params(P.make_params(prm, co_offenders=co_offenders))
        E.upload_state(0, st)
        rep = E.run_ticks(1, ticks, [key])
        after = E.download_state(0)
        cn, _ = E.counters()
    assert cn[0, 9] == 0, "device error code %d" % cn[0, 9]
    oc = S.cols()
    for k in COLS:
        assert np.array_equal(after[k], oc[k]), k
    for l in (6, 7, 8):
        rp, col, co = S.layer(l)
        assert np.array_equal(after["rowptr_%d" % l], rp) and np.array_equal(after["col_%d" % l], col), l
        if l == 6:
            assert np.array_equal(after["co_off"], co)
    assert rep[0, -1, 0] == totals[0] and rep[0, -1, 15] == mult
    return rep, totals


@pytest.mark.parametrize("seed", range(8))
def test_random_states(seed):
    import pyproton_oc_b200 as P
    from oracle import oracle_c as O
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(200, 1500))
    st = random_state(rng, n, mean_deg=float(rng.uniform(4, 22)), oc_frac=float(rng.uniform(0.01, 0.15)))
    prm = {"max_accomplice_radius": int(rng.integers(1, 5)), "oc_embeddedness_radius": int(rng.integers(1, 4)),
           "threshold_use_facilitators": int(rng.integers(2, 5)), "number_arrests_per_year": int(rng.integers(20, 2000)),
           "number_crimes_yearly_per10k": int(rng.integers(1500, 6000)),
           "facilitator_repression": bool(seed % 3 == 1), "oc_boss_repression": bool(seed % 3 == 2),
           "ticks_between_intervention": 1, "intervention_start": 2, "punishment_length": 1}
    co = [(1, 0.5), (2, 0.2), (3, 0.1), (5, 0.1), (7, 0.1)] if seed % 2 else None
    rep, totals = run_both(P, O, st, prm, ticks=14, key=77 + seed, co_offenders=co)
    assert totals[0] > 0


@pytest.mark.parametrize("emb_g", [None, "128", "256"])
def test_dense_clique_overflows_staged_edges(emb_g, monkeypatch):
    """150-member OC clique + co-offence counts above 255: more induced entries than the staged-edge capacity and
    multiplicities that need the two-slot entry form.  With small blocks (128 / 256 threads per evaluation, the shapes
    large batches run) the evaluations that do not fit are handed to the large-block launch."""
    if emb_g:
        monkeypatch.setenv("POC_EMB_G", emb_g)
    import pyproton_oc_b200 as P
    from oracle import oracle_c as O
    rng = np.random.default_rng(5)
    st = random_state(rng, 900, mean_deg=10, oc_frac=0.0, clique=150)
    st["co_off"][:] = np.where(rng.random(len(st["co_off"])) < 0.02, 300, st["co_off"])
    # symmetrise the co-offence counts (Person.num_co_offenses is kept symmetric by commit_crime)
    n = len(st["alive"])
    src = np.repeat(np.arange(n), np.diff(st["rowptr_6"]))
    d = {}
    for a, b, c in zip(src.tolist(), st["col_6"].tolist(), st["co_off"].tolist()):
        d[(min(a, b), max(a, b))] = max(d.get((min(a, b), max(a, b)), 0), c)
    st["co_off"] = np.array([d[(min(a, b), max(a, b))] for a, b in zip(src.tolist(), st["col_6"].tolist())], np.int32)
    run_both(P, O, st, {"number_crimes_yearly_per10k": 4000}, ticks=6, key=9)


def hub_state(rng, n=6000, hub=17, nfriends=2600):
    st = random_state(rng, n, mean_deg=6, oc_frac=0.02)
    others = [i for i in range(n) if i != hub][:nfriends]
    rp, col = st["rowptr_5"], st["col_5"]
    src = np.repeat(np.arange(n), np.diff(rp))
    pairs = set(zip(src.tolist(), col.tolist())) | {(hub, o) for o in others} | {(o, hub) for o in others}
    key = np.array(sorted(a * n + b for a, b in pairs), dtype=np.int64)
    s, dd = key // n, key % n
    newrp = np.zeros(n + 1, np.int32)
    np.add.at(newrp, s + 1, 1)
    st["rowptr_5"], st["col_5"] = np.cumsum(newrp).astype(np.int32), dd.astype(np.int32)
    st["oc_member"][hub] = 1
    st["alive"][hub] = 1
    return st


def test_large_blocks_take_what_small_blocks_cannot_in_concurrent_groups(monkeypatch):
    """256 replicas of the hub state in four replica groups on their own streams, 128 threads per evaluation: the small
    blocks hand the hub's neighbourhood to the large-block launch (second size class), whose scratch slots must not be
    shared with the launches of the other groups running at the same time.  Sampled replicas equal their single runs."""
    monkeypatch.setenv("POC_GROUPS", "4")
    monkeypatch.setenv("POC_EMB_G", "128")
    import pyproton_oc_b200 as P
    st = hub_state(np.random.default_rng(6), n=3000, nfriends=1400)
    prm = P.make_params({"number_crimes_yearly_per10k": 3000, "oc_embeddedness_radius": 2})
    R, ticks = 256, 3
    keys = [31 + 5 * r for r in range(R)]
    n, stat, crim = P.CrimeEngine.capacity_for([st], 200000)
    with P.CrimeEngine(R, n, stat, crim) as E:
        E.set_params(prm)
        for r in range(R):
            E.upload_state(r, st, wait=False)
        E.run_ticks(1, ticks, keys)
        got = {r: E.download_state(r) for r in (0, 25, 100, 200, 255)}
        cn, _ = E.counters()
    assert (cn[:, 9] == 0).all()
    monkeypatch.delenv("POC_GROUPS")
    monkeypatch.delenv("POC_EMB_G")
    for r, g in got.items():
        with P.CrimeEngine(1, n, stat, crim) as E1:
            E1.set_params(prm)
            E1.upload_state(0, st)
            E1.run_ticks(1, ticks, [keys[r]])
            want = E1.download_state(0)
        for k in COLS:
            assert np.array_equal(g[k], want[k]), (r, k)
        assert np.array_equal(g["col_6"], want["col_6"]) and np.array_equal(g["co_off"], want["co_off"]), r


@pytest.mark.parametrize("emb_g", [None, "128"])
def test_huge_ball_uses_global_distance_array(emb_g, monkeypatch):
    """A hub with 2600 friends: balls larger than the shared-memory distance array (and than the 12-bit rank
    fields when the radius is 2)."""
    if emb_g:
        monkeypatch.setenv("POC_EMB_G", emb_g)
    import pyproton_oc_b200 as P
    from oracle import oracle_c as O
    rng = np.random.default_rng(6)
    n = 6000
    st = random_state(rng, n, mean_deg=6, oc_frac=0.02)
    hub = 17
    others = [i for i in range(n) if i != hub][:2600]
    links = {l: [] for l in U.LAYERS}
    # rebuild with the hub's friendships added
    rp, col = st["rowptr_5"], st["col_5"]
    src = np.repeat(np.arange(n), np.diff(rp))
    pairs = set(zip(src.tolist(), col.tolist())) | {(hub, o) for o in others} | {(o, hub) for o in others}
    key = np.array(sorted(a * n + b for a, b in pairs), dtype=np.int64)
    s, dd = key // n, key % n
    newrp = np.zeros(n + 1, np.int32)
    np.add.at(newrp, s + 1, 1)
    st["rowptr_5"], st["col_5"] = np.cumsum(newrp).astype(np.int32), dd.astype(np.int32)
    st["oc_member"][hub] = 1
    st["alive"][hub] = 1
    run_both(P, O, st, {"number_crimes_yearly_per10k": 3000, "oc_embeddedness_radius": 2}, ticks=3, key=3,
             crim_headroom=400000)
