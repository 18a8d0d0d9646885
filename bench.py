#!/usr/bin/env python
"""bench.py -- agent-ticks/s of the crime / recruitment hot path (BASELINE.json metric) on B200.

One "step" = the hot path of one complete run (num_ticks ticks of
{age update; yearly tendency + multiplier; reset_oc_embeddedness + commit_crimes; prisoner countdown},
the hot-path subset of ProtonOC.step, model.py:226-288) for a batch of `replicas` independent 10k-agent
simulations per GPU (the override-mode ensemble; each replica has its own Philox key = seed).  With
--gpus N (torchrun, one rank per GPU) every rank runs its own batch: weak scaling, no data-path collective,
the only exchange is the final gather of the reporter rows.

  value     whole-job agent-ticks/s, state resident in HBM when the timed region starts (CUDA events)
  e2e       same metric through the public API with HOST state: per step upload of every replica's
            state from pinned host memory, the run, download of the reporter rows and of the agent columns
  roofline  dominant kernel class: algorithmic bytes / CUDA-event launch time vs measured HBM peak
  cpu_baseline  the oracle port (oracle/hotpath.c), 1 host core, one full run of the same workload

--impl reference times the CPU implementation of the path (the oracle port; the reference itself is pure
Python and cannot travel to the GPU box) on all host cores, one process per run like the reference's
ProcessPoolExecutor (run_modes.py:221-232).
"""
import argparse
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

WORKLOADS = ["state_n10000_s42_t13", "state_n10000_s42_baseline"]


def load_workload():
    import util as U
    for name in WORKLOADS:
        path = os.path.join(U.GOLDEN, name + ".npz")
        if os.path.exists(path):
            fx = U.load_fixture(path)
            return name, U.pre_state(fx), U.fixture_params(fx), int(fx["tick"]), float(fx["crime_multiplier"])
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


def cpu_baseline(st, prm, tick0, ticks, mult, n_agents):
    dt, totals = _oracle_run((st, prm, tick0, ticks, 1, mult))
    return {"value": n_agents * ticks / dt, "unit": "agent-ticks/s", "cores": 1, "kind": "port",
            "sample": "oracle/hotpath.c, one full run: %d agents x %d ticks (%.1f s), %d crimes, %d embeddedness evaluations"
                      % (n_agents, ticks, dt, totals[0], totals[3])}


def reference_arm(args):
    """CPU implementation of the path on all host cores: one oracle run per process."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    name, st, prm, tick0, mult = load_workload()
    n_agents = int(st["alive"].sum())
    cores = len(os.sched_getaffinity(0))
    procs = max(1, min(cores, 64))
    ticks = args.ticks
    jobs = [(st, prm, tick0, ticks, 1000 + i, mult) for i in range(procs)]
    times = []
    with mp.get_context("fork").Pool(procs) as pool:
        for i in range(min(args.warmup, 1) + args.steps):
            t = time.perf_counter()
            pool.map(_oracle_run, jobs)
            dt = time.perf_counter() - t
            if i >= min(args.warmup, 1):
                times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = n_agents * ticks * procs / (ms / 1e3)
    line = {"impl": "reference", "metric": "agent-ticks/s", "value": value, "unit": "agent-ticks/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s: %d agents x %d ticks, hot path only, %d concurrent runs" % (name, n_agents, ticks, procs)},
            "cpu_baseline": {"value": value, "unit": "agent-ticks/s", "cores": procs, "kind": "port",
                             "sample": "oracle/hotpath.c (C restatement; the Python reference cannot travel), one run "
                                       "per process on %d host cores, %d steps" % (procs, args.steps)},
            "e2e": {"value": value, "unit": "agent-ticks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ---------------------------------------------------------------------------------------------- GPU arm
def state_bytes(st):
    return int(sum(v.nbytes for k, v in st.items() if k != "unique_id" and k != "retired"))




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


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--replicas", type=int, default=1024, help="independent runs per GPU in one batch")
    ap.add_argument("--ticks", type=int, default=480)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
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

    name, st, prm, tick0, mult = load_workload()
    n_agents = int(st["alive"].sum())
    R, T = args.replicas, args.ticks
    n, stat, crim = pkg.CrimeEngine.capacity_for([st], 160000)
    E = pkg.CrimeEngine(R, n, stat, crim, device=local)
    P = pkg.make_params(prm)
    E.set_params(P)
    keys = [(rank * R + r) * 7919 + 17 for r in range(R)]
    # pinned host copy of the state (PyTorch is the pinned-memory allocator here, nothing more)
    pinned = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in st.items()}
    st_pin = {k: v.numpy() for k, v in pinned.items()}
    h2d = state_bytes(st) * R
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def upload_all():
        for r in range(R):
            E.upload_state(r, st_pin, wait=False)     # queued; E.sync() / the run's own synchronisation wait for them
            E.set_crime_multiplier(mult, r)

    def barrier():
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

    # ---- resident-state throughput (value)
    times = []
    launches = 0
    with ClockSampler(local) as clk:
        for i in range(args.warmup + args.steps):
            upload_all()
            flush.fill_(i & 0xff)          # L2 flush between iterations (and the batch state exceeds L2 anyway)
            barrier()
            E.timers_reset()
            E.marker(0)
            E.run_ticks(tick0, T, keys, reporters=False)
            E.marker(1)
            E.sync()
            barrier()
            if i >= args.warmup:
                times.append(E.marker_elapsed_ms(0, 1))
                launches = E.launch_count()
        clocks = clk.summary()
    ms = max_over_ranks(float(np.mean(times)))
    value = n_agents * T * R * world / (ms / 1e3)

    # ---- end to end through the public API with host buffers
    e2e_times = []
    d2h = 0
    host_rep = np.zeros((R, T, len(pkg.REPORTER_NAMES)), np.float64)
    host_cols = [E.agent_buffers(len(st["alive"])) for _ in range(R)]   # result buffers allocated (and touched) once
    if dist is not None:
        # multi-GPU: the reporter rows go from every rank's device memory straight into the NCCL gather (NVLink) and
        # reach the host once, on rank 0, in a pinned buffer
        shape = (R, T, len(pkg.REPORTER_NAMES))
        gathered = [torch.empty(shape, dtype=torch.float64, device=dev) for _ in range(world)] if rank == 0 else None
        host_all = torch.empty((world,) + shape, dtype=torch.float64).pin_memory() if rank == 0 else None
    for i in range(1 + args.steps):
        barrier()
        t0 = time.perf_counter()
        upload_all()
        if dist is None:
            rep = E.run_ticks(tick0, T, keys, reporters=True, out=host_rep)
            rep_bytes = rep.nbytes
        else:
            rows = E.run_ticks_device(tick0, T, keys)
            dist.gather(rows.contiguous(), gathered, dst=0)   # the ensemble's only exchange
            rep_bytes = 0
            if rank == 0:
                for w_ in range(world):
                    host_all[w_].copy_(gathered[w_], non_blocking=True)
                torch.cuda.synchronize()
                rep_bytes = host_all.numel() * 8
        cols = [E.download_agents(r, out=host_cols[r]) for r in range(R)]
        E.sync()
        dt = time.perf_counter() - t0
        d2h = rep_bytes + sum(sum(v.nbytes for v in c.values()) for c in cols)
        if i >= 1:
            e2e_times.append(dt * 1e3)
    e2e_ms = max_over_ranks(float(np.mean(e2e_times)))
    e2e_value = n_agents * T * R * world / (e2e_ms / 1e3)

    # ---- per-kernel-class device time (separate pass with event pairs around every launch)
    upload_all()
    E.stats(reset=True)
    E.set_timing(True)
    E.timers_reset()
    E.run_ticks(tick0, T, keys, reporters=False)
    tm_ms, tm_n = E.timers()
    E.set_timing(False)
    cn, edges = E.counters()
    stats = E.stats()
    total_ms = sum(tm_ms.values())
    top = max(tm_ms, key=tm_ms.get)
    # algorithmic bytes of the traversal kernels, counted by the kernels themselves
    alg = {"embeddedness": stats["emb_bytes"], "search": stats["search_bytes"]}
    roof_class = top if top in alg else "embeddedness"
    roof_ms = tm_ms[roof_class]
    roof_launch_ms = roof_ms / max(tm_n[roof_class], 1)
    achieved = alg[roof_class] / (roof_ms / 1e3) / 1e9 if roof_ms > 0 else None

    # ---- single-run latency (batch of one): the "one 10k-agent run" figure
    single = None
    if rank == 0:
        E1 = pkg.CrimeEngine(1, n, stat, crim, device=local)
        E1.set_params(P)
        ts = []
        for i in range(3):
            E1.upload_state(0, st_pin); E1.set_crime_multiplier(mult, 0)
            E1.sync(); E1.marker(0)
            E1.run_ticks(tick0, T, [keys[0]], reporters=False)
            E1.marker(1); E1.sync()
            ts.append(E1.marker_elapsed_ms(0, 1))
        single = {"agent_ticks_per_s": n_agents * T / (min(ts[1:]) / 1e3), "ms_per_run": min(ts[1:]),
                  "us_per_tick": 1e3 * min(ts[1:]) / T}
        E1.close()

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        line = {
            "metric": "agent-ticks/s", "value": value, "unit": "agent-ticks/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s: %d agents x %d ticks, crime/recruitment hot path only, %d replicas per GPU "
                                   "(distinct Philox seeds, same initial state)" % (name, n_agents, T, R),
                       "replicas_per_gpu": R, "ticks": T, "agents": n_agents,
                       "l2": "L2 flushed between steps; batch state (%.0f MB) exceeds L2" % (R * 1.8)},
            "runs_per_hour": R * world * 3600.0 / (ms / 1e3),
            "single_run": single,
            "e2e": {"value": e2e_value, "unit": "agent-ticks/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "kernel_ms": {k: round(v, 3) for k, v in tm_ms.items()},
            "kernel_launches": tm_n,
            "kernel_share_top": {"class": top, "share": tm_ms[top] / total_ms if total_ms else None},
            "traversed_edges_per_s": float(edges.sum()) / (total_ms / 1e3) if total_ms else None,
            "roofline": {"bound": "hbm", "kernel": "k_" + roof_class, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": (achieved / hbm_peak) if achieved else None,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel
                         # (profiles/r01_05_*: 42.8 MB for a launch of 5,760 evaluations), scaled to this run's
                         # evaluations per launch
                         "traffic": (298.3e6 / 12957.0) * (stats["emb_requests"] / max(tm_n[roof_class], 1)) if roof_class == "embeddedness" else None,
                         "traffic_source": "ncu capture profiles/r01_06_ncu_full_k_embeddedness_512_raw.csv (dram read + write of one launch / its evaluations = 23 KB per evaluation) x evaluations per launch",
                         "avg_launch_ms": roof_launch_ms, "launches": tm_n[roof_class],
                         "algorithmic_bytes_per_launch": alg[roof_class] / max(tm_n[roof_class], 1),
                         "units_per_launch": (stats["emb_requests"] if roof_class == "embeddedness" else stats["search_crimes"]) / max(tm_n[roof_class], 1),
                         "bytes_per_unit": alg[roof_class] / max(stats["emb_requests"] if roof_class == "embeddedness" else stats["search_crimes"], 1),
                         "peak_source": "MEASURED_PEAKS.json (copy bandwidth, of measured)" if peaks else "fallback 6650 GB/s",
                         "note": "instruction-issue bound (sm__issue_active 74 %), DRAM traffic below the algorithmic bytes (L2 reuse across the balls of a replica); see DESIGN.md section 5 and profiles/README.md"},
            "work": stats,
        }
        if not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(st, prm, tick0, T, mult, n_agents)
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
