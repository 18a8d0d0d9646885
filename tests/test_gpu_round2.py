"""Round-2 GPU checks: the reporter kernel against values the reference collected, packed uploads, and the k_draw
window mode on mixed-size / mostly-dead batches."""
import numpy as np
import pytest

import util as U

pytestmark = pytest.mark.gpu

# device reporter column -> attribute ProtonOC.calculate_fast_reporters sets (model.py:1687-1735)
FAST = ["current_oc_members", "current_prisoners", "tot_criminal_link", "number_crimes_committed_of_persons",
        "crimes_committed_by_oc_this_tick", "criminal_tendency_mean", "criminal_tencency_sd", "current_num_persons",
        "age_mean", "age_sd", "education_level_mean", "education_level_sd", "num_crime_committed_mean",
        "num_crime_committed_sd", "employed", "facilitators", "number_students", "tot_sibling_link", "tot_offspring_link",
        "tot_parent_link", "tot_partner_link", "tot_household_link", "tot_friendship_link", "tot_professional_link",
        "tot_school_link"]
FLOATS = {"criminal_tendency_mean", "criminal_tencency_sd", "age_mean", "age_sd", "education_level_mean",
          "education_level_sd", "num_crime_committed_mean", "num_crime_committed_sd"}


@pytest.fixture(scope="module")
def P():
    import pyproton_oc_b200 as pkg
    return pkg


def engine_for(P, states, crim_headroom=20000):
    n, stat, crim = P.CrimeEngine.capacity_for(states, crim_headroom)
    return P.CrimeEngine(len(states), n, stat, crim)


@pytest.mark.parametrize("path", U.golden_files("rep"), ids=lambda p: p.split("/")[-1][:-4])
def test_report_kernel_against_reference_values(P, path):
    """k_report (poc_report) on the state a reference step() ended in, against the values the reference's own
    calculate_fast_reporters computed from it (fixtures rep_*: recorded right after model.py:286): integers exact,
    means / standard deviations within 1e-12 relative."""
    fx = U.load_fixture(path)
    st = U.pre_state(fx)
    with engine_for(P, [st]) as E:
        E.set_params(P.make_params(U.fixture_params(fx)))
        E.upload_state(0, st)
        row = dict(zip(P.REPORTER_NAMES, E.report()[0]))
    for k in FAST:
        want = float(fx["rep_" + k])
        if k in FLOATS:
            assert abs(row[k] - want) <= 1e-12 * max(abs(want), 1e-300), (k, row[k], want)
        else:
            assert row[k] == want, (k, row[k], want)


def test_packed_upload_equals_column_upload(P):
    fx = U.load_fixture(U.GOLDEN + "/tick_n3000_s19_oc300_big_t2.npz")
    st, prm = U.pre_state(fx), U.fixture_params(fx)
    packed = P.CrimeEngine.pack_state(st, pinned=True)
    n, stat, crim = P.CrimeEngine.capacity_for([st], 60000)
    res = []
    for mode in ("columns", "packed"):
        with P.CrimeEngine(3, n + 100, stat + 1000, crim) as E:
            E.set_params(P.make_params(prm))
            for r in range(3):
                if mode == "columns":
                    E.upload_state(r, st, wait=False)
                else:
                    E.upload_packed(r, packed)
            E.sync()
            first = E.download_state(1)
            rep = E.run_ticks(2, 10, [7, 8, 9])
            res.append((first, rep, E.download_state(2)))
    for k in res[0][0]:
        assert np.array_equal(res[0][0][k], res[1][0][k]), k
        assert np.array_equal(res[0][2][k], res[1][2][k]), k
    for k in st:
        if k not in ("unique_id", "retired"):
            assert np.array_equal(res[1][0][k], st[k]), k       # round trip through the packed block
    assert np.array_equal(res[0][1], res[1][1])


def test_packed_upload_rejects_bad_blocks(P):
    fx = U.load_fixture(U.GOLDEN + "/tick_n100_s42_t1.npz")
    st = U.pre_state(fx)
    packed = P.CrimeEngine.pack_state(st)
    with P.CrimeEngine(1, 50, 100000, 1000) as E:             # too few agent slots
        with pytest.raises(P.PocError):
            E.upload_packed(0, packed)
    with engine_for(P, [st]) as E:
        bad = P.CrimeEngine.pack_state(st)
        bad.blob[0] ^= 0xff                                    # magic
        with pytest.raises(P.PocError):
            E.upload_packed(0, bad)


def test_draw_window_mode_mixed_sizes_and_mostly_dead(P):
    """with a ctx sized for > 20k slots k_draw runs its chains over shared-memory windows; a replica
    whose own alive count would fit the whole-cell layout (a smaller population in the same batch, or a large state
    that is mostly dead slots) must take the windowed path too instead of writing behind the weights."""
    from oracle import oracle_c as O
    from pyproton_oc_b200 import synth
    tpl = U.pre_state(U.load_fixture(U.GOLDEN + "/state_n10000_s42_baseline.npz"))
    big = synth.synthetic_state(24000, seed=5, template=tpl, num_oc_persons=30)
    dead = {k: v.copy() for k, v in big.items()}
    rng = np.random.default_rng(3)
    kill = rng.random(len(dead["alive"])) < 0.8
    kill &= dead["oc_member"] == 0
    dead["alive"][kill] = 0                                     # dead agents stay as referenced slots (entities.py:652-661)
    small = U.pre_state(U.load_fixture(U.GOLDEN + "/tick_n1000_s11_oc200_t3.npz"))
    states, keys, ticks = [big, dead, small], [11, 12, 13], 4
    n, stat, crim = P.CrimeEngine.capacity_for(states, 200000)
    with P.CrimeEngine(3, n, stat, crim) as E:
        E.set_params(P.make_params({}))
        for r, st in enumerate(states):
            E.upload_state(r, st)
        E.run_ticks(1, ticks, keys)
        got = [E.download_state(r) for r in range(3)]
        cn, _ = E.counters()
    assert (cn[:, 9] == 0).all()
    for r, st in enumerate(states):
        S = O.OracleState(st)
        S.run_ticks(O.make_params({}), 1, ticks, keys[r], 0.0)
        oc = S.cols()
        for k in ["oc_member", "prisoner", "num_crimes_committed", "criminal_tendency", "new_recruit"]:
            assert np.array_equal(got[r][k], oc[k]), (r, k)
        rp, col, co = S.layer(6)
        assert np.array_equal(got[r]["col_6"], col) and np.array_equal(got[r]["co_off"], co), r


def test_download_agents_many_equals_per_replica_downloads(P):
    """poc_download_agents_many (four staging blocks in flight) returns what poc_download_agents returns replica by replica:
    nine replicas of two different sizes after a few ticks."""
    fa = U.load_fixture(U.GOLDEN + "/tick_n1000_s42_t12.npz")
    fb = U.load_fixture(U.GOLDEN + "/tick_n100_s42_t1.npz")
    sa, sb = U.pre_state(fa), U.pre_state(fb)
    states = [sa, sb, sa, sa, sb, sa, sb, sb, sa]
    n, stat, crim = P.CrimeEngine.capacity_for(states, 40000)
    with P.CrimeEngine(len(states), n, stat, crim) as E:
        E.set_params(P.make_params(U.fixture_params(fa)))
        for r, st in enumerate(states):
            E.upload_state(r, st)
        E.run_ticks(12, 5, [100 + r for r in range(len(states))])
        one = [E.download_agents(r) for r in range(len(states))]
        outs = [E.agent_buffers(len(st["alive"])) for st in states]
        E.download_agents_many(0, outs)
        part = [E.agent_buffers(len(st["alive"])) for st in states[3:8]]
        E.download_agents_many(3, part)
    for r in range(len(states)):
        for k in one[r]:
            assert np.array_equal(one[r][k], outs[r][k]), (r, k)
    for i, r in enumerate(range(3, 8)):
        for k in one[r]:
            assert np.array_equal(one[r][k], part[i][k]), (r, k)
