"""The reference-side binding (pyproton_oc_b200.refbind) on the UNMODIFIED reference: export of the Person graph,
the path on an executor, import of what the path changed.  Build container only (needs /root/reference).

CPU tests use an oracle-backed stand-in executor (the tier allows the oracle as executor in tests); the GPU test at the
bottom drives the same binding through CrimeEngine."""
import numpy as np
import pytest

import util as U

pytestmark = pytest.mark.reference


class OracleExecutor:
    """Stand-in with the executor interface of refbind.EngineExecutor, computing on oracle/hotpath.c."""

    def __init__(self):
        from oracle import oracle_c as O
        self.O = O

    def _state(self, st, params, tables):
        c_range, co, conv = tables
        return self.O.OracleState(st), self.O.make_params(params, c_range=c_range, co_offenders=co, conviction=conv)

    def run_tendency(self, st, params, tables):
        S, P = self._state(st, params, tables)
        S.tendency(P)
        return S.cols()["criminal_tendency"], S.crime_multiplier(P)

    def run_multiplier(self, st, params, tables):
        S, P = self._state(st, params, tables)
        return S.crime_multiplier(P)

    def run_commit(self, st, params, tables, tick, mult, replay, key):
        S, P = self._state(st, params, tables)
        res = S.commit_crimes(P, tick, mult, replay, key)
        after = S.cols()
        for l in (6, 7, 8):
            rp, col, co = S.layer(l)
            after["rowptr_%d" % l], after["col_%d" % l] = rp, col
            if l == 6:
                after["co_off"] = co
        return res["out"], {"groups": res["groups"], "arrested": res["arrested"], "initiator": res["initiator"], "k": res["k"]}, after


def _ref():
    from oracle import ref_harness as H
    return H, H.load_reference()[0]


COMPARE_COLS = ["alive", "oc_member", "prisoner", "num_crimes_committed", "num_crimes_committed_this_tick", "new_recruit",
                "job_level", "has_job", "has_school", "sentence_countdown", "criminal_tendency", "age"]


def _assert_same(sa, sb, what):
    assert np.array_equal(sa["unique_id"], sb["unique_id"]), what
    for k in COMPARE_COLS:
        assert np.array_equal(sa[k], sb[k]), "%s: column %s" % (what, k)
    for l in range(9):
        assert np.array_equal(sa["rowptr_%d" % l], sb["rowptr_%d" % l]), "%s: layer %d" % (what, l)
        assert np.array_equal(sa["col_%d" % l], sb["col_%d" % l]), "%s: layer %d" % (what, l)
    assert np.array_equal(sa["co_off"], sb["co_off"]), what


def test_export_matches_the_fixture_exporter():
    """refbind.export_state (product code) and oracle/ref_harness.export_state (what recorded the fixtures) agree."""
    from pyproton_oc_b200 import refbind
    H, ref_model = _ref()
    m = ref_model.ProtonOC(seed=3)
    m.setup(300)
    for _ in range(14):
        m.step()
    a, agents = refbind.export_state(m)
    b = H.export_state(m)
    assert set(a) == set(b)
    for k in b:
        assert np.array_equal(a[k], b[k]), k
    assert [p.unique_id for p in agents] == a["unique_id"].tolist()


@pytest.mark.parametrize("seed,n,overrides", [(6, 400, None), (10, 500, {"num_oc_persons": 400, "num_oc_families": 6}),
                                              (11, 400, {"intervention": "facilitators"}),
                                              (13, 400, {"intervention": "disruptive-strong", "num_oc_persons": 300, "num_oc_families": 6})])
def test_bound_model_in_lockstep_with_the_reference(seed, n, overrides):
    """Two reference models with one seed.  A runs unmodified (its commit_crimes recorded); B has the four hot-path
    methods bound to the executor and is fed A's uniforms (replay) tick by tick.  After every full step() the two
    Person graphs must be identical -- criminal links and co-offence counts, crime counters, OC membership, arrests with
    their job / school bookkeeping, the counters, and the tendencies of the yearly pass -- for two simulated years,
    unless a crime hit a tie between equal sort keys (CPython set order), where the comparison stops."""
    from pyproton_oc_b200 import refbind
    H, ref_model = _ref()
    A, B = ref_model.ProtonOC(seed=seed), ref_model.ProtonOC(seed=seed)
    for m in (A, B):
        if overrides:
            m.override_dict(overrides)
        m.setup(n)
    bind = refbind.bind(B, OracleExecutor())
    _assert_same(H.export_state(A), H.export_state(B), "after setup")   # (tick 1 runs the bound yearly tendency pass)
    ticks_equal = crimes = multi = 0
    for t in range(24):
        h = H.step_with_recording(A)
        pre = h["pre"]
        slot_of = {int(u): j for j, u in enumerate(pre["unique_id"])}
        rp = bind.replay = H.replay_from_record(h["rec"], slot_of)
        bind.rng_state_after = h["rng_state_after"]
        B.step()
        out, rec = bind.last
        ref_groups = [[slot_of[u] for u in c["group"]] for c in h["rec"]["crimes"]]
        same = all(g.tolist() == U.canonical_group(r) for g, r in zip(rec["groups"], ref_groups)) and len(ref_groups) == len(rec["groups"])
        crimes += len(ref_groups)
        multi += sum(len(g) > 1 for g in ref_groups)
        if not same:
            break       # a tie flipped a pick: B is no longer A's twin
        # arrests index the flattened criminals list: comparable when its order is the canonical one at those positions
        flat_ref = [m_ for g in ref_groups for m_ in g]
        flat_can = [m_ for g in ref_groups for m_ in U.canonical_group(g)]
        if any(flat_ref[i] != flat_can[i] for i in rp["arrest_idx"].tolist()):
            break
        _assert_same(H.export_state(A), H.export_state(B), "after tick %d" % (t + 1))
        for k in ["number_crimes", "crime_size_fails", "facilitator_crimes", "facilitator_fails", "big_crime_from_small_fish",
                  "number_offspring_recruited_this_tick", "people_jailed", "number_law_interventions_this_tick",
                  "crime_multiplier", "current_oc_members", "tot_criminal_link", "number_crimes_committed_of_persons"]:
            assert getattr(A, k) == getattr(B, k), (t, k)
        assert A.random.bit_generator.state == B.random.bit_generator.state
        ticks_equal += 1
    assert ticks_equal >= 12 and crimes > 50, (ticks_equal, crimes, multi)


def test_patched_class_reproduces_the_reference_in_distribution():
    """patch_reference on the CLASS, production random numbers (Philox): over a handful of seeds the per-run totals of
    crimes, criminal links, OC members and arrests of the patched reference are statistically indistinguishable from the
    unmodified reference's (two-sample KS test per column, and means within 15 %)."""
    from scipy.stats import ks_2samp
    from pyproton_oc_b200 import refbind
    H, ref_model = _ref()
    n, ticks, seeds = 300, 24, list(range(20, 30))

    def run(patched):
        rows = []
        restore = refbind.patch_reference(ref_model.ProtonOC, OracleExecutor) if patched else None
        try:
            for s in seeds:
                m = ref_model.ProtonOC(seed=s)
                m.override_dict({"num_oc_persons": 200, "num_oc_families": 4})
                m.setup(n)
                for _ in range(ticks):
                    m.step()
                rows.append((m.number_crimes, m.tot_criminal_link, m.current_oc_members, m.people_jailed,
                             m.number_crimes_committed_of_persons))
        finally:
            if restore:
                restore()
        return np.array(rows, dtype=np.float64)
    a, b = run(False), run(True)
    for c, name in enumerate(["number_crimes", "tot_criminal_link", "current_oc_members", "people_jailed", "participations"]):
        p = ks_2samp(a[:, c], b[:, c]).pvalue
        assert p > 0.01, (name, p, a[:, c], b[:, c])
        if a[:, c].mean() > 5:
            assert abs(a[:, c].mean() - b[:, c].mean()) <= 0.15 * a[:, c].mean(), (name, a[:, c].mean(), b[:, c].mean())


@pytest.mark.gpu
def test_bound_model_on_the_gpu():
    """The same binding through CrimeEngine (libprotonoc_b200.so): one reference model with the hot path on the B200,
    in lockstep with an unmodified twin under replay for a year."""
    from pyproton_oc_b200 import refbind
    H, ref_model = _ref()
    A, B = ref_model.ProtonOC(seed=6), ref_model.ProtonOC(seed=6)
    for m in (A, B):
        m.setup(400)
    bind = refbind.bind(B)
    ok = 0
    for t in range(12):
        h = H.step_with_recording(A)
        slot_of = {int(u): j for j, u in enumerate(h["pre"]["unique_id"])}
        rp = bind.replay = H.replay_from_record(h["rec"], slot_of)
        bind.rng_state_after = h["rng_state_after"]
        B.step()
        out, rec = bind.last
        ref_groups = [[slot_of[u] for u in c["group"]] for c in h["rec"]["crimes"]]
        if not all(g.tolist() == U.canonical_group(r) for g, r in zip(rec["groups"], ref_groups)):
            break
        flat_ref = [m_ for g in ref_groups for m_ in g]
        flat_can = [m_ for g in ref_groups for m_ in U.canonical_group(g)]
        if any(flat_ref[i] != flat_can[i] for i in rp["arrest_idx"].tolist()):
            break
        _assert_same(H.export_state(A), H.export_state(B), "after tick %d" % (t + 1))
        ok += 1
    bind.executor.close()
    assert ok >= 6
