#!/usr/bin/env python
"""bench.py -- agent-ticks/s of the crime / recruitment hot path (BASELINE.json metric) on B200.

One "step" = the hot path of one complete run (num_ticks ticks of
{age update; yearly tendency + multiplier; reset_oc_embeddedness + commit_crimes; prisoner countdown},
the hot-path subset of ProtonOC.step, model.py:226-288) for a batch of `replicas` independent simulations per GPU
(the override-mode ensemble; each replica has its own Philox key = seed).  With --gpus N (torchrun, one rank per GPU)
every rank runs its own batch: weak scaling, no data-path collective, the only exchange is the final gather of the
reporter rows.

--config selects the BASELINE.json configuration (default 5 = the headline):
  1  ProtonOC().run(100, 480), the reference's own setup state (seed 42)                 single run
  2  baseline parameters, 3,000 agents x 480 ticks                                         single run
  3  the three samples/sample_json.json variants, 10,000 agents x 480 ticks                3 runs, own parameters
  4  synthetic 100,000-agent population, 480 ticks                                         single run
  5  override-mode ensemble: `--replicas` (1,024) runs per GPU drawn from the three sample_json variants x a pool of
     eight reference setup states per variant (tests/golden/pool_*), distinct Philox seeds; the homogeneous batch of
     round 1 (one state, one parameter set) is measured beside it (`homogeneous`).

  value     whole-job agent-ticks/s, state resident in HBM when the timed region starts (CUDA events)
  e2e       same metric through the public API with HOST state: per step upload of every replica's packed state from
            pinned host memory, the run, download of the reporter rows and of the agent columns
  roofline  dominant kernel class (and `roofline_search` for the accomplice search): algorithmic bytes /
            CUDA-event launch time vs the measured HBM peak
  with_step_phases  the headline batch again with the eight phases of step() that exist on the device switched on (the yearly
            labour-status block; wedding, retire_persons, make_baby, remove_excess_*, make_friends, make_people_die every tick);
            measured last, a failure there is reported in the line instead of losing the figures above
  cpu_baseline  the UNMODIFIED reference (baseline/_ref through oracle/refshim) on one host core, a bounded sample,
            with the C restatement (oracle/hotpath.c) beside it (`port`)

--impl reference times the CPU implementation of the path on all host cores, one run per process like the reference's
ProcessPoolExecutor (run_modes.py:221-232): the C restatement on the bench workload (the Python reference needs hours
for it), plus a bounded sample of the reference itself.
"""
import argparse
import glob
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SAMPLE_OVERRIDES = {   # samples/sample_json.json of the reference (initial_agents 10000 in all three)
    "0": {"intervention": "facilitators", "num_oc_persons": 80, "number_arrests_per_year": 20, "num_oc_families": 16},
    "1": {"intervention": "baseline", "num_oc_persons": 20, "number_arrests_per_year": 24, "num_oc_families": 4},
    "2": {"intervention": "students", "num_oc_persons": 20, "number_arrests_per_year": 28, "num_oc_families": 4}}


def _fixture(name):
    import util as U
    return U.load_fixture(os.path.join(U.GOLDEN, name + ".npz"))


def variant_params(overrides):
    """The hot-path parameters a ProtonOC with these overrides runs with (interventions expanded, model.py:1244-1348)."""
    import pyproton_oc_b200 as pkg
    m = pkg.ProtonOC(seed=1)
    m.override_dict(overrides)
    m.choose_intervention_setting()
    return {k: getattr(m, k) for k in pkg.HOT_PATH_DEFAULTS}


def load_config(cfg, replicas):
    """-> dict(name, entries=[(state, params, tick0, mult)], assign=[entry index per replica], n_agents)"""
    import util as U
    if cfg == 1:
        fx = _fixture("tick_n100_s42_t1")
        ent = [(U.pre_state(fx), U.fixture_params(fx), 1, 0.0)]
        return dict(name="config 1: tick_n100_s42_t1 (reference setup, 100 agents)", entries=ent, assign=[0])
    if cfg == 2:
        fx = _fixture("state_n3000_s42_baseline")
        ent = [(U.pre_state(fx), U.fixture_params(fx), 1, 0.0)]
        return dict(name="config 2: state_n3000_s42_baseline (reference setup, 3,000 agents)", entries=ent, assign=[0])
    if cfg == 3:
        ent = []
        for v in "012":
            fx = _fixture("state_n10000_s42_sample" + v)
            ent.append((U.pre_state(fx), variant_params(SAMPLE_OVERRIDES[v]), 1, 0.0))
        return dict(name="config 3: samples/sample_json.json runs 0/1/2 (reference setups, 10,000 agents)", entries=ent,
                    assign=[0, 1, 2])
    if cfg == 4:
        from pyproton_oc_b200 import synth
        tpl = U.pre_state(_fixture("state_n10000_s42_baseline"))
        st = synth.synthetic_state(100000, seed=42, template=tpl, num_oc_persons=30)
        return dict(name="config 4: synthetic_n100000_s42", entries=[(st, {}, 1, 0.0)], assign=[0])
    # config 5
    pool = sorted(glob.glob(os.path.join(U.GOLDEN, "pool_n10000_s*_sample*.npz")))
    if not pool:
        return None
    ent, by_variant = [], {"0": [], "1": [], "2": []}
    for p in pool:
        v = p[-5]
        fx = U.load_fixture(p)
        by_variant[v].append(len(ent))
        ent.append((U.pre_state(fx), variant_params(SAMPLE_OVERRIDES[v]), 1, 0.0))
    assign = []
    for r in range(replicas):   # replica r: variant r % 3, state (r // 3) % pool size of that variant
        lst = by_variant["012"[r % 3]]
        assign.append(lst[(r // 3) % len(lst)])
    return dict(name="config 5 sweep: samples/sample_json.json variants 0/1/2 x %d reference setup states each "
                     "(tests/golden/pool_*), 10,000 agents" % min(len(v) for v in by_variant.values()),
                entries=ent, assign=assign)


def homogeneous_config(replicas):
    import util as U
    for name in ["state_n10000_s42_t13", "state_n10000_s42_baseline"]:
        path = os.path.join(U.GOLDEN, name + ".npz")
        if os.path.exists(path):
            fx = U.load_fixture(path)
            st = U.pre_state(fx)
            edu = os.path.join(U.GOLDEN, name + "_edu.npz")   # the two columns the yearly labour-status block reads
            if "max_education_level" not in st and os.path.exists(edu):
                st.update(U.pre_state(U.load_fixture(edu)))
            ent = [(st, U.fixture_params(fx), int(fx["tick"]), float(fx["crime_multiplier"]))]
            return dict(name="%s (one reference state, one parameter set)" % name, entries=ent, assign=[0] * replicas)
    raise SystemExit("no workload state under tests/golden")


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index=0):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self.gpu = gpu_index
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0])); self.max_mhz = float(f[1])
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self.t.join(timeout=6)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------- CPU arms
def _oracle_run(args):
    st, prm, tick0, ticks, key, mult = args
    from oracle import oracle_c as O
    S = O.OracleState(st)
    P = O.make_params(prm)
    t = time.perf_counter()
    _, totals = S.run_ticks(P, tick0, ticks, key, mult)
    return time.perf_counter() - t, totals.tolist()


def reference_sample(n_agents=3000, ticks=24, max_seconds=40.0):
    """A bounded sample of the UNMODIFIED reference on one host core (None when no reference tree travels with the repo)."""
    try:
        from oracle import ref_harness, ref_timing
        if not ref_harness.reference_available():
            return None
        import warnings
        warnings.filterwarnings("ignore")
        return ref_timing.time_reference(n_agents, ticks, seed=42, max_seconds=max_seconds)
    except Exception as e:   # the reference's own dependencies (numba, networkx, pandas) missing on the box
        return {"error": repr(e)}


def cpu_baseline(ent, ticks, n_agents):
    """Both CPU figures, one host core each: the reference itself (bounded sample: its setup at 3,000 agents + 24 ticks;
    a 10,000-agent tick takes it 15-20 s) and the C restatement on one full run of the bench workload."""
    st, prm, tick0, mult = ent
    dt, totals = _oracle_run((st, prm, tick0, ticks, 1, mult))
    port = {"value": n_agents * ticks / dt, "unit": "agent-ticks/s", "cores": 1, "kind": "port",
            "sample": "oracle/hotpath.c, one full run of the workload state: %d agents x %d ticks (%.1f s), %d crimes, "
                      "%d embeddedness evaluations" % (n_agents, ticks, dt, totals[0], totals[3])}
    rs = reference_sample()
    if not rs or "error" in rs:
        port["reference_sample"] = rs
        return port
    return {"value": rs["agent_ticks_per_s_hot_path"], "unit": "agent-ticks/s", "cores": 1, "kind": "reference",
            "sample": "unmodified reference (baseline/_ref via oracle/refshim), ProtonOC(seed=42): setup(%d) then %d ticks in "
                      "%.1f s; value = agents x ticks / time inside reset_oc_embeddedness + commit_crimes + yearly "
                      "tendency / multiplier (%.0f %% of the step); its rate falls with the population (profiles/: "
                      "%s)" % (rs["n_agents"], rs["ticks"], rs["step_s"], 100 * rs["hot_share"], REF_10K_NOTE),
            "whole_step": rs["agent_ticks_per_s_whole_step"], "hot_share": rs["hot_share"], "setup_s": rs["setup_s"],
            "port": port}


REF_10K_NOTE = ("profiles/r02_reference_times_gpu_box.jsonl, same box class, one core: 100 agents 258 k, 3,000 agents x 480 ticks "
                "19.7 k, 10,000 agents x 24 ticks 1,558 agent-ticks/s on the hot path")


def reference_arm(args):
    """CPU implementation of the path on all host cores: one oracle run per process."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import oracle_c as O
    O.lib()   # built and loaded in this process too (the workers are forks of it)
    cfg = homogeneous_config(1)
    st, prm, tick0, mult = cfg["entries"][0]
    n_agents = int(st["alive"].sum())
    cores = len(os.sched_getaffinity(0))
    procs = max(1, min(cores, 64))
    ticks = args.ticks
    jobs = [(st, prm, tick0, ticks, 1000 + i, mult) for i in range(procs)]
    times = []
    warm = min(args.warmup, 1)
    with mp.get_context("fork").Pool(procs) as pool:
        for i in range(warm + args.steps):
            t = time.perf_counter()
            pool.map(_oracle_run, jobs)
            dt = time.perf_counter() - t
            if i >= warm:
                times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = n_agents * ticks * procs / (ms / 1e3)
    line = {"impl": "reference", "metric": "agent-ticks/s", "value": value, "unit": "agent-ticks/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s: %d agents x %d ticks, hot path only, %d concurrent runs" % (cfg["name"], n_agents, ticks, procs)},
            "cpu_baseline": {"value": value, "unit": "agent-ticks/s", "cores": procs, "kind": "port",
                             "sample": "oracle/hotpath.c (C restatement of the path), one run per process on %d host "
                                       "cores, %d steps; the Python reference itself: see reference_sample" % (procs, args.steps)},
            "e2e": {"value": value, "unit": "agent-ticks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    line["reference_sample"] = reference_sample()
    emit(line)


# ---------------------------------------------------------------------------------------------- GPU arm
_REAL_STDOUT = None


def quiet_stdout():
    """The contract is ONE JSON line on stdout: anything a library prints there while the bench runs (NCCL's
    version banner, for one) is sent to stderr instead."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(line), flush=True)


def measured_traffic():
    """DRAM bytes per embeddedness evaluation from the committed ncu --set full capture of the final build
    (profiles/r02_traffic.json, written by tools/ncu_traffic.py from the .ncu-rep); None when absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", type=int, default=5, choices=[1, 2, 3, 4, 5])
    ap.add_argument("--replicas", type=int, default=None, help="independent runs per GPU in one batch (config 5: 1024)")
    ap.add_argument("--ticks", type=int, default=480)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-sweep", action="store_true", help="config 5: skip the parameter-sweep batch measured beside the headline")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
        return
    import torch
    import pyproton_oc_b200 as pkg

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)

    T = args.ticks
    R = args.replicas if args.replicas is not None else (1024 if args.config == 5 else {1: 1, 2: 1, 3: 3, 4: 1}[args.config])
    # config 5: the headline workload stays the homogeneous batch of round 1 (one reference state at tick 13, one parameter
    # set: comparable across rounds and with the --impl reference arm); the parameter sweep over the reference's own setup
    # states is measured beside it (`sweep`)
    if args.config == 5:
        cfg, sweep_cfg = homogeneous_config(R), (None if args.no_sweep else load_config(5, R))
    else:
        cfg, sweep_cfg = load_config(args.config, R), None
    R = len(cfg["assign"])
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def barrier(E):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        E.sync()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    class Workload:
        """One batch on one engine: packed pinned states, per-replica parameters and keys."""

        def __init__(self, c, E=None, phases=0):
            self.c = c
            self.phases = phases
            ents, self.assign = c["entries"], c["assign"]
            self.R = len(self.assign)
            self.alive = [int(e[0]["alive"].sum()) for e in ents]
            self.agent_ticks = sum(self.alive[a] for a in self.assign) * T      # per step on this rank
            states = [e[0] for e in ents]
            head = 160000 if max(len(s["alive"]) for s in states) <= 20000 else 16 * max(len(s["alive"]) for s in states)
            self.cap = pkg.CrimeEngine.capacity_for(states, head)
            if phases:   # births add agent slots and links on the device
                self.cap = (self.cap[0] + 3000, self.cap[1] + 100000, self.cap[2])
            self.E = E if E is not None else pkg.CrimeEngine(self.R, *self.cap, device=local)
            if phases:
                self.E.set_demography(pkg.make_demography())
                if phases & pkg.CrimeEngine.PHASE_LABOUR:
                    self.E.set_labour(pkg.make_labour())
                self.E.set_phases(phases)
            self.packed = [pkg.CrimeEngine.pack_state(s, pinned=True) for s in states]
            self.params = [pkg.make_params(e[1]) for e in ents]
            self.tick0 = ents[0][2]
            assert all(e[2] == self.tick0 for e in ents)
            self.keys = [(rank * self.R + r) * 7919 + 17 for r in range(self.R)]
            self.h2d = sum(self.packed[a].nbytes for a in self.assign)
            same = all(a == self.assign[0] for a in self.assign)
            if same:
                self.E.set_params(self.params[self.assign[0]])
            else:
                for r, a in enumerate(self.assign):
                    self.E.set_params(self.params[a], replica=r)

        def upload_all(self):
            E, ents = self.E, self.c["entries"]
            for r, a in enumerate(self.assign):
                E.upload_packed(r, self.packed[a])     # queued; E.sync() / the run's own synchronisation wait for them
                if ents[a][3]:
                    E.set_crime_multiplier(ents[a][3], r)
            if self.phases & pkg.CrimeEngine.PHASE_LABOUR:   # outside every timed region (state resident when the clock starts)
                E.sync()
                for r, a in enumerate(self.assign):
                    if "max_education_level" in ents[a][0]:
                        E.upload_education(r, ents[a][0]["max_education_level"], ents[a][0]["school_level"])

        def time_resident(self, warmup, steps, clk=True):
            E, times, launches = self.E, [], 0
            for i in range(warmup + steps):
                self.upload_all()
                flush.fill_(i & 0xff)          # L2 flush between iterations (and the batch state exceeds L2 anyway)
                barrier(E)
                E.timers_reset()
                E.marker(0)
                E.run_ticks(self.tick0, T, self.keys, reporters=False)
                E.marker(1)
                E.sync()
                barrier(E)
                if i >= warmup:
                    times.append(E.marker_elapsed_ms(0, 1))
                    launches = E.launch_count()
            ms = max_over_ranks(float(np.mean(times)))
            return ms, self.agent_ticks * world / (ms / 1e3), launches

    def e2e_pass(W, steps):
        """Through the public API with HOST buffers: upload of every replica, run, download of reporter rows + agent columns."""
        E, Rw = W.E, W.R
        e2e_times, d2h = [], 0
        host_rep = np.zeros((Rw, T, len(pkg.REPORTER_NAMES)), np.float64)
        host_cols = [E.agent_buffers(len(W.c["entries"][a][0]["alive"])) for a in W.assign]   # result buffers allocated (and touched) once
        if dist is not None:
            # multi-GPU: the reporter rows go from every rank's device memory straight into the NCCL gather (NVLink) and
            # reach the host once, on rank 0, in a pinned buffer
            shape = (Rw, T, len(pkg.REPORTER_NAMES))
            gathered = [torch.empty(shape, dtype=torch.float64, device=dev) for _ in range(world)] if rank == 0 else None
            host_all = torch.empty((world,) + shape, dtype=torch.float64).pin_memory() if rank == 0 else None
        for i in range(1 + steps):
            barrier(E)
            t0 = time.perf_counter()
            W.upload_all()
            if dist is None:
                rep = E.run_ticks(W.tick0, T, W.keys, reporters=True, out=host_rep)
                rep_bytes = rep.nbytes
            else:
                rows = E.run_ticks_device(W.tick0, T, W.keys)
                dist.gather(rows.contiguous(), gathered, dst=0)   # the ensemble's only exchange
                rep_bytes = 0
                if rank == 0:
                    for w_ in range(world):
                        host_all[w_].copy_(gathered[w_], non_blocking=True)
                    torch.cuda.synchronize()
                    rep_bytes = host_all.numel() * 8
            cols = E.download_agents_many(0, host_cols)   # every replica's agent columns, one call (copies pipelined)
            E.sync()
            dt = time.perf_counter() - t0
            d2h = rep_bytes + sum(sum(v.nbytes for v in c.values()) for c in cols)
            if i >= 1:
                e2e_times.append(dt * 1e3)
        e2e_ms = max_over_ranks(float(np.mean(e2e_times)))
        # where the end-to-end time goes (one extra step with a synchronisation after every part; single GPU only)
        breakdown = None
        if dist is None:
            t0 = time.perf_counter(); W.upload_all(); E.sync()
            t1 = time.perf_counter(); E.run_ticks(W.tick0, T, W.keys, reporters=True, out=host_rep)
            t2 = time.perf_counter()
            E.download_agents_many(0, host_cols)
            E.sync()
            t3 = time.perf_counter()
            breakdown = {"upload": 1e3 * (t1 - t0), "run_and_reporter_rows": 1e3 * (t2 - t1), "agent_columns": 1e3 * (t3 - t2)}
        return {"value": W.agent_ticks * world / (e2e_ms / 1e3), "unit": "agent-ticks/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(W.h2d), "d2h_bytes_per_step": int(d2h), "breakdown_ms": breakdown}

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic = measured_traffic()

    def kernel_pass(W):
        """Per-kernel-class device time (a separate pass with event pairs around every launch) and the rooflines."""
        E = W.E
        W.upload_all()
        E.stats(reset=True)
        E.set_timing(True)
        E.timers_reset()
        E.run_ticks(W.tick0, T, W.keys, reporters=False)
        tm_ms, tm_n = E.timers()
        E.set_timing(False)
        cn, edges = E.counters()
        per_rep = E.stats_per_replica()
        stats = {k: int(v.sum()) for k, v in per_rep.items()}
        total_ms = sum(tm_ms.values())
        top = max(tm_ms, key=tm_ms.get)

        def roof(cls, alg_bytes, units, unit_name, kname):
            t_ms, nl = tm_ms[cls], max(tm_n[cls], 1)
            ach = alg_bytes / (t_ms / 1e3) / 1e9 if t_ms > 0 else None
            per_unit = (traffic or {}).get(kname, {}).get("dram_bytes_per_unit")
            return {"bound": "hbm", "kernel": kname, "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                    "frac": (ach / hbm_peak) if ach else None,
                    # dram__bytes_read.sum + dram__bytes_write.sum of the committed ncu --set full capture of this build,
                    # per unit of work, times this run's units per launch
                    "traffic": per_unit * units / nl if per_unit is not None else None,
                    "traffic_source": (traffic or {}).get(kname, {}).get("source"),
                    "avg_launch_ms": t_ms / nl, "launches": tm_n[cls], "algorithmic_bytes_per_launch": alg_bytes / nl,
                    "units_per_launch": units / nl, "unit_of_work": unit_name, "bytes_per_unit": alg_bytes / max(units, 1),
                    "share_of_step": t_ms / total_ms if total_ms else None,
                    "peak_source": "MEASURED_PEAKS.json (copy bandwidth, of measured)" if peaks else "fallback 6650 GB/s (of fallback)"}
        r1 = roof("embeddedness", stats["emb_bytes"], stats["emb_requests"], "oc_embeddedness evaluation", "k_emb3")
        r1["note"] = ("latency / instruction-issue bound, not bandwidth bound: one evaluation is a BFS, a per-row gather pass and a "
                      "shortest-path fixed point on a few hundred nodes; DESIGN.md section 5 gives the measured ceilings (issue slots, "
                      "shared-memory wavefronts) and profiles/README.md the ncu captures")
        r2 = roof("search", stats["search_bytes"], stats["search_crimes"], "accomplice search (crime with k >= 1)", "k_search")
        out = {"kernel_ms": {k: round(v, 3) for k, v in tm_ms.items()}, "kernel_launches": tm_n,
               "kernel_share_top": {"class": top, "share": tm_ms[top] / total_ms if total_ms else None},
               "traversed_edges_per_s": float(edges.sum()) / (total_ms / 1e3) if total_ms else None,
               "roofline": r1, "roofline_search": r2, "work": stats}
        return out, per_rep

    W = Workload(cfg)
    E = W.E
    n_agents = W.alive[W.assign[0]]

    # ---- resident-state throughput (value)
    with ClockSampler(local) as clk:
        ms, value, launches = W.time_resident(args.warmup, args.steps)
        clocks = clk.summary()
    e2e = e2e_pass(W, args.steps)
    kp, _ = kernel_pass(W)
    n_entries = len(cfg["entries"])
    E.close()
    del W

    # ---- config 5: the parameter sweep beside the headline
    sweep = None
    if sweep_cfg is not None:
        SW = Workload(sweep_cfg)
        sms, sval, _ = SW.time_resident(2, max(1, min(args.steps, 2)))
        se2e = e2e_pass(SW, 1)
        skp, per_rep = kernel_pass(SW)
        variants = {}
        for v in range(3):   # replica r runs variant r % 3
            sel = np.arange(SW.R) % 3 == v
            variants[str(v)] = {"overrides": SAMPLE_OVERRIDES[str(v)], "replicas": int(sel.sum()),
                                "embeddedness_evaluations_per_tick": float(per_rep["emb_requests"][sel].sum()) / max(int(sel.sum()), 1) / T,
                                "bytes_per_evaluation": float(per_rep["emb_bytes"][sel].sum()) / max(float(per_rep["emb_requests"][sel].sum()), 1.0),
                                "accomplice_searches_per_tick": float(per_rep["search_crimes"][sel].sum()) / max(int(sel.sum()), 1) / T}
        sweep = {"workload": sweep_cfg["name"], "distinct_initial_states": len(sweep_cfg["entries"]), "value": sval,
                 "ms_per_step": sms, "runs_per_hour": SW.R * world * 3600.0 / (sms / 1e3), "e2e": se2e, "variants": variants,
                 "kernel_ms": skp["kernel_ms"], "roofline": skp["roofline"], "roofline_search": skp["roofline_search"]}
        SW.E.close()
        del SW

    # ---- single-run latency (batch of one): the "one 10k-agent run" figure
    single = None
    if rank == 0:
        ent = cfg["entries"][0]
        one = dict(name="single", entries=[ent], assign=[0])
        S1 = Workload(one)
        ts = []
        for i in range(3):
            S1.upload_all()
            S1.E.sync(); S1.E.marker(0)
            S1.E.run_ticks(S1.tick0, T, [S1.keys[0]], reporters=False)
            S1.E.marker(1); S1.E.sync()
            ts.append(S1.E.marker_elapsed_ms(0, 1))
        a1 = S1.alive[0]
        single = {"agent_ticks_per_s": a1 * T / (min(ts[1:]) / 1e3), "ms_per_run": min(ts[1:]),
                  "us_per_tick": 1e3 * min(ts[1:]) / T, "agents": a1}
        S1.E.close()

    # ---- config 5 with the step() phases that run on the device switched on: the yearly labour-status block (graduate_and_
    # enter_jobmarket, update_unemployment_status) and, every tick, wedding, retire_persons, make_baby, remove_excess_friends /
    # _professional_links, make_friends, make_people_die: births and deaths keep the age structure and the criminal layer from
    # drifting the way the crime path alone lets them over 480 ticks.  Measured last and contained: a failure here is reported in
    # the line, the headline figures above are already taken.
    demog = None
    if args.config == 5 and not args.no_sweep:
        try:
            dcfg = homogeneous_config(R)
            has_edu = "max_education_level" in dcfg["entries"][0][0]
            DW = Workload(dcfg, phases=pkg.CrimeEngine.PHASE_EVERY if has_edu else pkg.CrimeEngine.PHASE_ALL)
            dms, dval, dl = DW.time_resident(1, max(1, min(args.steps, 2)))
            cn0, _ = DW.E.counters()
            dkp, _ = kernel_pass(DW)
            cn, _ = DW.E.counters()
            cn = cn - cn0 * (cn0 > 0) * np.isin(np.arange(cn.shape[1]), [0, 16, 17, 19])[None, :]   # cumulative counters: this pass only
            cols0 = DW.E.download_agents(0)
            al0 = cols0["alive"] > 0
            demog = {"phases": ("graduate_and_enter_jobmarket + update_unemployment_status on the yearly ticks (model.py:251-256); " if has_edu else "") +
                               "wedding, retire_persons, make_baby, remove_excess_friends, remove_excess_professional_links, "
                               "make_friends, make_people_die (model.py:271-284) every tick, on the device",
                     "phase_mask": int(DW.phases),
                     "value": dval, "ms_per_step": dms, "gpu_launches": int(dl), "kernel_ms": dkp["kernel_ms"],
                     "roofline": dkp["roofline"], "errors": int((cn[:, 9] != 0).sum()),
                     "born_per_run": float(cn[:, 17].mean()), "deceased_per_run": float(cn[:, 16].mean()),
                     "crimes_per_run": float(cn[:, 0].mean()), "weddings_per_run": float(cn[:, 19].mean()),
                     "education_level_mean_after_run": float(cols0["education_level"][al0].mean()),
                     "students_after_run": int((cols0["has_school"][al0] > 0).sum()),
                     "note": "value counts the agents of the initial state x ticks, like the headline"}
            DW.E.close()
            del DW
        except Exception as ex:   # noqa: BLE001
            if dist is not None:   # the other ranks wait in this block's collectives: fail the job rather than hang it
                raise
            demog = {"error": "%s: %s" % (type(ex).__name__, ex)}

    if rank == 0:
        line = {
            "metric": "agent-ticks/s", "value": value, "unit": "agent-ticks/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s x %d ticks, crime/recruitment hot path only, %d replicas per GPU (distinct Philox "
                                   "seeds)" % (cfg["name"], T, R),
                       "baseline_config": args.config, "replicas_per_gpu": R, "ticks": T, "agents": n_agents,
                       "distinct_initial_states": n_entries,
                       "l2": "L2 flushed between steps; batch state (%.0f MB) %s L2" % (R * 1.8 * n_agents / 1e4, "exceeds" if R * n_agents > 7e5 else "fits")},
            "runs_per_hour": R * world * 3600.0 / (ms / 1e3),
            "sweep": sweep,
            "with_step_phases": demog,
            "single_run": single,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        line.update(kp)
        if not args.no_cpu:
            hent = cfg["entries"][0]
            line["cpu_baseline"] = cpu_baseline(hent, T, int(hent[0]["alive"].sum()))
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
