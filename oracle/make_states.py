"""Export full-size reference setup states (the output of the reference's own ProtonOC.setup, model.py:1003-1049)
as the workload inputs for bench.py and the full-size tests.  TEST INFRASTRUCTURE, build container only.

    python -m oracle.make_states            # baseline 3k + 10k, and the three samples/sample_json.json variants
"""
import json
import os
import sys
import time

import numpy as np

from . import ref_harness as H
from .make_golden import OUT, PARAM_KEYS


def export_setup(name, n_agents, seed, overrides=None, steps=0):
    """steps > 0: run that many reference step()s after setup and export the state the next tick starts from."""
    ref_model, _, _ = H.load_reference()
    model = ref_model.ProtonOC(seed=seed)
    if overrides:
        model.override_dict(overrides)
    t = time.time()
    model.setup(n_agents)
    for _ in range(steps):
        model.step()
    st = H.export_state(model)
    d = {"pre_" + k: v for k, v in st.items()}
    for k in PARAM_KEYS:
        d["param_" + k] = np.float64(getattr(model, k))
    d["tick"] = np.int32(model.tick)
    d["crime_multiplier"] = np.float64(model.crime_multiplier)
    d["overrides_json"] = np.array(json.dumps(overrides or {}))
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **d)
    print("wrote %s (%.1f KB, setup %.1fs, %d OC, %d facilitators)" % (
        path, os.path.getsize(path) / 1024, time.time() - t, int(st["oc_member"].sum()), int(st["facilitator"].sum())))


def export_education_sidecar(name, n_agents, seed, steps):
    """The two columns the yearly labour-status block reads (max_education_level, school_level) for a state fixture exported
    before they were part of export_state: the reference run is repeated (it is bit-reproducible), every other column is
    checked against the fixture, and the two columns go to <name>_edu.npz."""
    ref_model, _, _ = H.load_reference()
    model = ref_model.ProtonOC(seed=seed)
    model.setup(n_agents)
    for _ in range(steps):
        model.step()
    st = H.export_state(model)
    old = np.load(os.path.join(OUT, name + ".npz"))
    for k, v in st.items():
        if "pre_" + k not in old.files:   # columns added to export_state after the fixture was written
            continue
        assert np.array_equal(old["pre_" + k], v), k
    path = os.path.join(OUT, name + "_edu.npz")
    np.savez_compressed(path, pre_max_education_level=st["max_education_level"], pre_school_level=st["school_level"])
    print("wrote", path, os.path.getsize(path))


def export_sweep_pool(variant: str, seeds):
    """Config 5 (override-mode ensemble): the reference's own setup states at 10,000 agents for the three
    samples/sample_json.json variants x a pool of seeds -- what bench.py --sweep draws its 1,024 replicas from."""
    with open(os.path.join("/root/reference", "samples", "sample_json.json")) as f:
        sj = json.load(f)
    ov = dict(sj[variant])
    n = ov.pop("initial_agents")
    for seed in seeds:
        export_setup("pool_n%d_s%d_sample%s" % (n, seed, variant), n, seed, ov, 0)


def main(argv):
    if len(argv) > 1 and argv[1] == "distdemog":
        export_distribution_demog()
        return
    if len(argv) > 1 and argv[1] == "distdemog2":   # + retire_persons, remove_excess_friends / _professional_links
        export_distribution_demog("distdemog2_n1000_s42", keep=DEMOG_KEEP2)
        return
    if len(argv) > 1 and argv[1] == "distdemog3":   # + wedding
        export_distribution_demog("distdemog3_n1000_s42", keep=DEMOG_KEEP2 + ["wedding"])
        return
    if len(argv) > 2 and argv[1] == "pool":   # python -m oracle.make_states pool <variant> <seed> [<seed> ...]
        export_sweep_pool(argv[2], [int(x) for x in argv[3:]])
        return
    if len(argv) > 1 and argv[1] == "disthot":   # python -m oracle.make_states disthot <agents> <runs> <ticks> <workers> [oc | sample0]
        kind = argv[6] if len(argv) > 6 else ""
        ov = {"": None,
              # many OC members: OC-initiated group crimes, embeddedness-weighted candidates and recruitment every tick
              "oc": {"num_oc_persons": 300, "num_oc_families": 24},
              # samples/sample_json.json variant "0": facilitator repression, 80 OC members
              "sample0": {"intervention": "facilitators", "num_oc_persons": 80, "number_arrests_per_year": 20, "num_oc_families": 16}}[kind]
        export_distribution_hot(int(argv[2]), int(argv[3]), int(argv[4]), int(argv[5]), overrides=ov, tag="_" + kind if kind else "")
        return
    if len(argv) > 1 and argv[1] == "distdemogpar":   # python -m oracle.make_states distdemogpar <agents> <runs> <ticks> <workers>  (1000 30 60 / 10000 14 36)
        export_distribution_demog_parallel("distdemog4_n%d_s42" % int(argv[2]), int(argv[2]), int(argv[3]), int(argv[4]), int(argv[5]))
        return
    if len(argv) > 1 and argv[1] == "edu":   # python -m oracle.make_states edu state_n10000_s42_t13 10000 42 12
        export_education_sidecar(argv[2], int(argv[3]), int(argv[4]), int(argv[5]))
        return
    if len(argv) > 1 and argv[1] == "distdemog4":   # + graduate_and_enter_jobmarket (update_unemployment_status always ran)
        export_distribution_demog("distdemog4_n1000_s42", ticks=60, keep=DEMOG_KEEP2 + ["wedding", "graduate_and_enter_jobmarket"],
                                  cols=DEMOG_COLS + LABOUR_COLS, hists=True)
        return
    only = argv[1] if len(argv) > 1 else None
    cases = [("state_n3000_s42_baseline", 3000, 42, None), ("state_n10000_s42_baseline", 10000, 42, None)]
    with open(os.path.join(H.REFERENCE_ROOT, "samples", "sample_json.json")) as f:
        sj = json.load(f)
    for key, ov in sj.items():
        ov = dict(ov)
        n = ov.pop("initial_agents")
        cases.append(("state_n%d_s42_sample%s" % (n, key), n, 42, ov))
    cases = [c + (0,) for c in cases]
    # the 10k state a quarter into the second year (friendship layer grown to ~8.9 per agent): bench workload
    cases.append(("state_n10000_s42_t13", 10000, 42, None, 12))
    # four years in: arrests, recruitment, deaths and births of the reference itself behind the exported state
    cases.append(("state_n10000_s42_t49", 10000, 42, None, 48))
    for name, n, seed, ov, steps in cases:
        if only and not name.startswith(only):
            continue
        export_setup(name, n, seed, ov, steps)
    if not only or "dist_n1000_s42".startswith(only):
        export_distribution()




OUT_OF_SCOPE_PHASES = ["wedding", "retire_persons", "make_baby", "remove_excess_friends",
                       "remove_excess_professional_links", "make_friends", "make_people_die",
                       "graduate_and_enter_jobmarket", "let_migrants_in", "return_kids"]


DEMOG_KEEP = ["make_baby", "make_friends", "make_people_die"]
DEMOG_KEEP2 = DEMOG_KEEP + ["retire_persons", "remove_excess_friends", "remove_excess_professional_links"]
DEMOG_COLS = ["number_crimes", "current_oc_members", "people_jailed", "tot_criminal_link", "tot_friendship_link",
              "tot_household_link", "tot_sibling_link", "tot_offspring_link", "tot_parent_link", "number_deceased",
              "number_born", "current_num_persons", "facilitators", "criminal_tendency_mean", "age_mean",
              "tot_professional_link", "tot_partner_link", "tot_school_link", "employed", "number_weddings"]


LABOUR_COLS = ["education_level_mean", "education_level_sd", "number_students"]


def _hists(model):
    """job / education / wealth / school level counts over the living (0..7 each), and the living students"""
    ag = model.schedule.agents
    h = np.zeros((4, 8), dtype=np.float64)
    for a in ag:
        h[0, int(a.job_level)] += 1
        h[1, int(a.education_level)] += 1
        h[2, int(a.wealth_level)] += 1
        h[3, int(a.my_school.diploma_level) if a.my_school is not None else 0] += 1
    return h


def export_distribution_demog(name="distdemog_n1000_s42", n_agents=1000, seed=42, n_runs=30, ticks=48, keep=None, cols=None,
                              hists=False):
    """Whole-run distribution fixture for the SURVEY.md 8(f) rows built on the device: the reference continued from ONE
    setup state with `n_runs` RNG streams, with its own make_baby / make_friends / make_people_die and the crime path
    active and every other phase of step() replaced by a no-op (`refdemog_*`).  Per tick: DEMOG_COLS."""
    ref_model, ref_entities, _ = H.load_reference()
    find_job = ref_entities.Person.find_job
    cols = cols or DEMOG_COLS
    d = {}
    rows_all, hist_all = [], []
    for j in range(n_runs):
        ref_entities.Person.find_job = find_job   # setup hands out the jobs through it (model.py:1045-1047)
        model = ref_model.ProtonOC(seed=seed)
        model.setup(n_agents)
        if keep and "retire_persons" in keep:
            # the yearly hiring loop is written out inside step() (model.py:257-269), not a phase that can be switched off: its
            # find_job would refill the jobs the retired leave and add professional links; with find_job a no-op AFTER setup the
            # loop does nothing.  (distdemog2 / distdemog3 were recorded with the no-op in place during setup as well: nobody has
            # a job in those two fixtures, so they say nothing about the professional layer; distdemog4 does.)
            ref_entities.Person.find_job = lambda self: None
        if not d:
            st = H.export_state(model)
            d.update({"pre_" + k: v for k, v in st.items()})
            for k in PARAM_KEYS:
                d["param_" + k] = np.float64(getattr(model, k))
            d["tick"] = np.int32(model.tick)
            d["crime_multiplier"] = np.float64(model.crime_multiplier)
        model.random = np.random.default_rng(1000 + j)
        for ph in OUT_OF_SCOPE_PHASES:
            if ph not in (keep or DEMOG_KEEP):
                setattr(model, ph, lambda *a, **k: None)
        rows = []
        for _ in range(ticks):
            model.step()
            rows.append([float(getattr(model, c)) for c in cols])
        rows_all.append(np.array(rows, dtype=np.float64))
        if hists:
            hist_all.append(_hists(model))
        print("refdemog run", j, rows_all[-1][-1][:12], flush=True)
    arr = np.array(rows_all)
    for i, col in enumerate(cols):
        d["refdemog_" + col] = arr[:, :, i]
    if hists:   # [run][job, education, wealth, school level][value] after the last tick
        d["refdemog_final_hists"] = np.array(hist_all)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **d)
    print("wrote", path)


def export_distribution(name="dist_n1000_s42", n_agents=1000, seed=42, n_runs=40, ticks=36):
    """Whole-run distribution fixture (north-star check 2).  The reference is continued from ONE setup state
    with `n_runs` different RNG streams, twice: `ref_*` = the full step(); `refhot_*` = step() with the phases
    that are out of scope here (OUT_OF_SCOPE_PHASES) replaced by no-ops, i.e. the reference's own code for
    exactly the path this repo implements.  Per tick: number_crimes, current_oc_members, people_jailed,
    current_prisoners, tot_criminal_link."""
    ref_model, _, _ = H.load_reference()
    d = {}
    for tag, hot_only in (("ref", False), ("refhot", True)):
        rows_all = []
        for j in range(n_runs):
            model = ref_model.ProtonOC(seed=seed)
            model.setup(n_agents)
            if not d:
                st = H.export_state(model)
                d.update({"pre_" + k: v for k, v in st.items()})
                for k in PARAM_KEYS:
                    d["param_" + k] = np.float64(getattr(model, k))
                d["tick"] = np.int32(model.tick)
                d["crime_multiplier"] = np.float64(model.crime_multiplier)
            model.random = np.random.default_rng(1000 + j)
            if hot_only:
                for ph in OUT_OF_SCOPE_PHASES:
                    setattr(model, ph, lambda *a, **k: None)
            rows = []
            for _ in range(ticks):
                model.step()
                rows.append((model.number_crimes, model.current_oc_members, model.people_jailed,
                             model.current_prisoners, model.tot_criminal_link))
            rows_all.append(np.array(rows, dtype=np.float64))
            print(tag, "run", j, rows_all[-1][-1])
        arr = np.array(rows_all)
        for i, col in enumerate(["number_crimes", "current_oc_members", "people_jailed", "current_prisoners",
                                 "tot_criminal_link"]):
            d["%s_%s" % (tag, col)] = arr[:, :, i]
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **d)
    print("wrote", path)


HOT_COLS = ["number_crimes", "current_oc_members", "people_jailed", "current_prisoners", "tot_criminal_link",
            "number_crimes_committed_of_persons", "crimes_committed_by_oc_this_tick", "big_crime_from_small_fish",
            "criminal_tendency_mean", "num_crime_committed_sd"]


def _hot_run(job):
    """One continuation of the setup state by the reference with the out-of-scope phases as no-ops (worker process)."""
    n_agents, seed, j, ticks, want_state, overrides = job
    ref_model, _, _ = H.load_reference()
    model = ref_model.ProtonOC(seed=seed)
    if overrides:
        model.override_dict(overrides)
    model.setup(n_agents)
    head = None
    if want_state:
        head = {"pre_" + k: v for k, v in H.export_state(model).items()}
        for k in PARAM_KEYS:
            head["param_" + k] = np.float64(getattr(model, k))
        head["tick"] = np.int32(model.tick)
        head["crime_multiplier"] = np.float64(model.crime_multiplier)
    model.random = np.random.default_rng(1000 + j)
    for ph in OUT_OF_SCOPE_PHASES:
        setattr(model, ph, lambda *a, **k: None)
    rows = []
    for _ in range(ticks):
        model.step()
        rows.append([float(getattr(model, c)) for c in HOT_COLS])
    print("refhot run", j, rows[-1][:5], flush=True)
    return j, np.array(rows, dtype=np.float64), head


def _demog_run(job):
    """One continuation with the phases in `keep` (and the crime path) active, every other phase a no-op (worker process)."""
    n_agents, seed, j, ticks, want_state, keep, cols = job
    ref_model, ref_entities, _ = H.load_reference()
    # a pool worker serves several runs: put the reference's own find_job back for the setup (which hands out the jobs through
    # it), switch it off AFTERWARDS (see export_distribution_demog)
    if not hasattr(ref_entities.Person, "_find_job_of_the_reference"):
        ref_entities.Person._find_job_of_the_reference = ref_entities.Person.find_job
    ref_entities.Person.find_job = ref_entities.Person._find_job_of_the_reference
    model = ref_model.ProtonOC(seed=seed)
    model.setup(n_agents)
    ref_entities.Person.find_job = lambda self: None
    head = None
    if want_state:
        head = {"pre_" + k: v for k, v in H.export_state(model).items()}
        for k in PARAM_KEYS:
            head["param_" + k] = np.float64(getattr(model, k))
        head["tick"] = np.int32(model.tick)
        head["crime_multiplier"] = np.float64(model.crime_multiplier)
    model.random = np.random.default_rng(1000 + j)
    for ph in OUT_OF_SCOPE_PHASES:
        if ph not in keep:
            setattr(model, ph, lambda *a, **k: None)
    rows = []
    for _ in range(ticks):
        model.step()
        rows.append([float(getattr(model, c)) for c in cols])
    print("refdemog run", j, rows[-1][:12], flush=True)
    return j, np.array(rows, dtype=np.float64), _hists(model), head


def export_distribution_demog_parallel(name, n_agents, n_runs, ticks, workers, seed=42):
    """export_distribution_demog with all eight device phases kept, one process per run (10,000-agent reference ticks take seconds)."""
    import multiprocessing as mp
    keep = DEMOG_KEEP2 + ["wedding", "graduate_and_enter_jobmarket"]
    cols = DEMOG_COLS + LABOUR_COLS
    jobs = [(n_agents, seed, j, ticks, j == 0, keep, cols) for j in range(n_runs)]
    with mp.get_context("fork").Pool(workers) as pool:
        res = sorted(pool.map(_demog_run, jobs, chunksize=1), key=lambda r: r[0])
    d = dict(res[0][3])
    arr = np.array([r[1] for r in res])
    for i, col in enumerate(cols):
        d["refdemog_" + col] = arr[:, :, i]
    d["refdemog_final_hists"] = np.array([r[2] for r in res])
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **d)
    print("wrote", path, os.path.getsize(path))


def export_distribution_hot(n_agents, n_runs, ticks, workers, seed=42, overrides=None, tag=""):
    """Whole-run distribution fixture at the sizes the path is benchmarked at (north-star check 2): `n_runs` continuations of
    ONE reference setup state by the reference's own code for exactly the crime path (`refhot_*`, like export_distribution),
    one process per run because a 10,000-agent tick of the reference takes seconds."""
    import multiprocessing as mp
    jobs = [(n_agents, seed, j, ticks, j == 0, overrides) for j in range(n_runs)]
    with mp.get_context("fork").Pool(workers) as pool:
        res = sorted(pool.map(_hot_run, jobs, chunksize=1), key=lambda r: r[0])
    d = dict(res[0][2])
    arr = np.array([r[1] for r in res])
    for i, col in enumerate(HOT_COLS):
        d["refhot_" + col] = arr[:, :, i]
    d["overrides_json"] = np.array(json.dumps(overrides or {}))
    path = os.path.join(OUT, "dist_n%d_s%d%s.npz" % (n_agents, seed, tag))
    np.savez_compressed(path, **d)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main(sys.argv)
