"""CPU-only checks of the host side: the C-ABI library exports what the header declares, the drop-in
ProtonOC keeps the reference's parameter API and exception types, run lists of the override mode, the
synthetic state generator, and the parameter block (CDFs as the reference builds them)."""
import ctypes as C
import json
import os
import re
import sys

import numpy as np
import pytest

import util as U

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from pyproton_oc_b200 import _cabi
    header = open(_cabi.HEADER).read()
    declared = set(re.findall(r"\b(poc_[a-z_0-9]+)\s*\(", header))
    assert len(declared) >= 30
    lib = C.CDLL(_cabi.build())
    for name in sorted(declared):
        assert hasattr(lib, name), "libprotonoc_b200.so lacks %s" % name
    assert declared == set(_cabi.SYMBOLS), declared ^ set(_cabi.SYMBOLS)
    L = _cabi.lib()
    assert L.poc_sizeof_params() == C.sizeof(_cabi.Params)


def test_no_cpu_fallback():
    import pyproton_oc_b200 as P
    from pyproton_oc_b200 import _cabi
    if _cabi.lib().poc_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(P.PocError):
        P.CrimeEngine(1, 10, 10, 10)
    m = P.ProtonOC(seed=1)
    with pytest.raises(P.PocError):
        m.setup(50)


def test_parameter_api_matches_reference_contract():
    import pyproton_oc_b200 as P
    m = P.ProtonOC(seed=7)
    assert repr(m) == "PROTON-OC MODEL, seed: 7"
    for name in P.free_parameters:
        assert hasattr(m, name)
    m.set_param("max_accomplice_radius", 4)
    assert m.max_accomplice_radius == 4
    with pytest.raises(AttributeError):
        m.set_param("not_a_param", 1)
    with pytest.raises(ValueError):
        m.set_param("max_accomplice_radius", 2.5)
    with pytest.raises(Exception, match="is not a model attribute"):
        m.override_dict({"bogus": 1})
    m.override_dict({"intervention": "facilitators", "num_oc_persons": 80, "repetitions": 3})
    m.choose_intervention_setting()
    assert m.facilitator_repression is True and m.facilitator_repression_multiplier == 3
    assert m.ticks_between_intervention == 1 and m.intervention_end == 9999
    assert m.tick == 1 and m.ticks_per_year == 12 and m.initial_agents == 1000 and m.num_ticks == 480
    m.tick = 13
    assert m.intervention_is_on()


def test_params_block_equals_oracle_block():
    import pyproton_oc_b200 as P
    from oracle import oracle_c as O
    over = {"number_arrests_per_year": 20, "max_accomplice_radius": 3, "facilitator_repression": True}
    a, b = P.make_params(over), O.make_params(over)
    assert bytes(a) == bytes(b)


SAMPLE_JSON = {"0": {"intervention": "facilitators", "num_oc_persons": 80, "number_arrests_per_year": 20,
                     "num_oc_families": 16, "initial_agents": 10000},
               "1": {"intervention": "baseline", "num_oc_persons": 20, "number_arrests_per_year": 24,
                     "num_oc_families": 4, "initial_agents": 10000},
               "2": {"intervention": "students", "num_oc_persons": 20, "number_arrests_per_year": 28,
                     "num_oc_families": 4, "initial_agents": 10000}}


def test_override_mode_run_lists(tmp_path):
    from pyproton_oc_b200 import ensemble
    p = tmp_path / "sample_json.json"
    p.write_text(json.dumps(SAMPLE_JSON))
    om = ensemble.OverrideMode(str(p), str(tmp_path), None, False, 42, 4, False)
    assert [a["filename"] for a in om.args] == ["sample_json0", "sample_json1", "sample_json2"]
    for a in om.args:
        assert set(a) == {"seed", "source_override", "save_path", "filename", "verbose", "snapshot", "alldata"}
        assert a["seed"] == 42 and a["verbose"] is False
    flat = tmp_path / "flat.json"
    flat.write_text(json.dumps({"num_oc_persons": 50}))
    om = ensemble.OverrideMode(str(flat), str(tmp_path), None, False, None, None, False)
    assert len(om.args) == 1 and om.args[0]["filename"] == "flat" and om.args[0]["verbose"] is True
    bad = tmp_path / "x.txt"
    bad.write_text("x")
    with pytest.raises(Exception, match="not a valid file"):
        ensemble.OverrideMode(str(bad), str(tmp_path), None, False, 1, None, False)
    sweep = ensemble.seed_sweep(4, [SAMPLE_JSON["0"], SAMPLE_JSON["1"]])
    assert len(sweep) == 8 and sorted({a["seed"] for a in sweep}) == [0, 1, 2, 3]


def test_partition_is_a_balanced_cover():
    from pyproton_oc_b200.ensemble import partition
    for n in (0, 1, 7, 1024, 3072):
        for world in (1, 2, 4, 8):
            parts = [partition(n, world, r) for r in range(world)]
            assert sorted(i for p in parts for i in p) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_synthetic_state_is_well_formed():
    from pyproton_oc_b200 import synth
    tpl = U.pre_state(U.load_fixture(os.path.join(U.GOLDEN, "state_n3000_s42_baseline.npz")))
    st = synth.synthetic_state(5000, seed=5, template=tpl, num_oc_persons=30)
    n = 5000
    assert int(st["oc_member"].sum()) == 15
    for l in range(9):
        rp, col = st["rowptr_%d" % l], st["col_%d" % l]
        assert rp[0] == 0 and rp[-1] == len(col) and (np.diff(rp) >= 0).all()
        src = np.repeat(np.arange(n), np.diff(rp))
        assert (col != src).all() and col.min(initial=0) >= 0 and col.max(initial=0) < n
        key = src.astype(np.int64) * n + col
        assert (np.diff(key) > 0).all()                      # rows ascending, no duplicates
        if l not in (1, 2):                                   # all layers but offspring/parent are symmetric
            assert set(key.tolist()) == set((col.astype(np.int64) * n + src).tolist())
    # offspring is the transpose of parent
    o = set((np.repeat(np.arange(n), np.diff(st["rowptr_1"])).astype(np.int64) * n + st["col_1"]).tolist())
    p = set((st["col_2"].astype(np.int64) * n + np.repeat(np.arange(n), np.diff(st["rowptr_2"]))).tolist())
    assert o == p
    assert len(st["co_off"]) == len(st["col_6"]) and (st["co_off"] == 1).all()
    # the oracle accepts it and runs
    from oracle import oracle_c as O
    S = O.OracleState(st)
    mult, totals = S.run_ticks(O.make_params({}), 1, 13, 5, 0.0)
    assert totals[0] > 0 and mult > 0


def test_person_api_builds_the_same_state_as_the_fixture_helper():
    """The Person link makers (entities.py:133-224 mirror) edit the host-side multiplex net; the resulting upload
    state equals the one tests/util.py builds directly for the reference's micro net (test_protonoc.py:577-625).
    Pure host logic: no engine is created before the first hot-path call."""
    import pyproton_oc_b200 as P
    import util as U
    model = P.ProtonOC(seed=2)
    for i in range(18):
        a = P.Person(model)
        model.schedule.add(a)
        a.gender_is_male = i % 2 == 0
        a.age = i * 149 % 45 + 10
        a.education_level = i * 149 % 4
        a.criminal_tendency = i * 149 % 100 / 100
        a.wealth_level = i * 149 % 4
        a.job_level = i * 149 % 4
        a.propensity = i * 149 % 100 / 100
    ag = model.schedule.agents
    for x, y in [(0, 1), (7, 8)]:
        ag[x].make_partner_link(ag[y])
    for x, y in [(0, 2), (6, 5), (8, 9), (13, 15)]:
        ag[x].make_friendship_link(ag[y])
    for x, y in [(0, 3), (10, 9), (10, 11), (11, 12), (12, 13), (12, 14), (12, 16), (14, 15)]:
        ag[x].make_professional_link(ag[y])
    for x, y in [(0, 4), (10, 13)]:
        ag[x].make_school_link(ag[y])
    ag[0].add_criminal_link(ag[5])
    ag[0].num_co_offenses[ag[5]] = 5
    ag[5].num_co_offenses[ag[0]] = 5
    ag[5].make_parent_offsprings_link(ag[0])
    ag[7].make_parent_offsprings_link(ag[9])
    ag[14].make_household_link([ag[17]])
    ag[16].make_household_link([ag[12]])
    ag[16].make_household_link([ag[17]])
    ag[5].oc_member = True
    ag[17].oc_member = True
    ag[9].facilitator = True
    ag[3].facilitator = True
    assert model._engine is None
    got, want = model.state_dict(), U.micro_net_state()
    for k in ["male", "age", "education_level", "wealth_level", "job_level", "criminal_tendency", "propensity",
              "oc_member", "facilitator", "alive"]:
        assert np.array_equal(np.asarray(got[k]).astype(np.float64), np.asarray(want[k]).astype(np.float64)), k
    for li in range(9):
        assert np.array_equal(got["rowptr_%d" % li], want["rowptr_%d" % li]), li
        assert np.array_equal(got["col_%d" % li], want["col_%d" % li]), li
    assert np.array_equal(got["co_off"], want["co_off"])
    # the reference's parent link is one-directional offspring + reverse parent (entities.py:206-215)
    assert ag[5].neighbors["offspring"] == {ag[0]} and ag[0].neighbors["parent"] == {ag[5]}


def test_reference_arm_prints_exactly_one_json_line():
    """bench.py contract: one JSON line on stdout (library chatter goes to stderr), with the reference-arm keys."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--ticks", "12"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "agent-ticks/s" and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": "agent-ticks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # under torchrun only rank 0 works and prints
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--ticks", "12"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
