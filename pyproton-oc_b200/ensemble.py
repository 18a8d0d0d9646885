"""Run modes of the reference (protonoc/run_modes.py) over the batched CUDA engine.

`BaseMode` / `OverrideMode` keep the reference's constructor arguments and the run-argument dictionaries
(keys seed, source_override, save_path, filename, verbose, snapshot, alldata; run_modes.py:63-69, 168-174,
190-204).  The reference fans runs out over a ProcessPoolExecutor, one process per run
(run_modes.py:221-232); here the run list is partitioned over the ranks of a torch.distributed job (one rank
per GPU), every rank executes its share in batches of `replicas_per_batch` replicas that share each kernel
launch, and the per-tick reporter rows are gathered on rank 0.  Runs are independent, so there is no
data-path collective: the gather of the outputs is the only exchange.
"""
from __future__ import annotations

import json
import os
import pickle
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import pandas as pd

from .engine import REPORTER_NAMES
from .model import ProtonOC, parse_behaviorspace_xml


# ---------------------------------------------------------------------------------------------- run lists
def runs_from_json(source: str, randomstate, save_path, snapshot, alldata, parallel) -> List[Dict]:
    """run_modes.py:177-205"""
    with open(source, "rb") as f:
        js = json.load(f)
    name = os.path.basename(source).replace(".json", "")
    first = js[list(js.keys())[0]]
    items = [(name + str(k), v) for k, v in js.items()] if type(first) == dict else [(name, js)]
    return [{"seed": randomstate, "source_override": ov, "save_path": save_path, "filename": fn,
             "verbose": False if parallel else True, "snapshot": snapshot, "alldata": alldata} for fn, ov in items]


def runs_from_xml(xml_file: str, randomstate, save_path, snapshot, alldata, parallel) -> List[Dict]:
    """run_modes.py:138-175"""
    from xml.dom import minidom
    name = os.path.basename(xml_file).replace(".xml", "")
    run = parse_behaviorspace_xml(xml_file)
    reps = int(minidom.parse(xml_file).getElementsByTagName("experiment")[0].attributes["repetitions"].value)
    return [{"seed": randomstate, "source_override": run, "save_path": save_path, "filename": name + "_run" + str(r + 1),
             "verbose": False if parallel else True, "snapshot": snapshot, "alldata": alldata} for r in range(reps)]


def seed_sweep(n_seeds: int, overrides: Sequence[Dict], save_path=None, prefix="run") -> List[Dict]:
    """The ensemble of BASELINE.json config 5: seeds 0..n_seeds-1 x a list of override dicts."""
    out = []
    for oi, ov in enumerate(overrides):
        for s in range(n_seeds):
            out.append({"seed": s, "source_override": dict(ov), "save_path": save_path,
                        "filename": "%s_o%d_s%d" % (prefix, oi, s), "verbose": False, "snapshot": None, "alldata": False})
    return out


def partition(n_items: int, world: int, rank: int) -> List[int]:
    """Round-robin: item i belongs to rank i % world (balanced to within one run)."""
    return list(range(rank, n_items, world))


# ---------------------------------------------------------------------------------------------- execution
def gpu_batch_runner(device: int = 0, template: Optional[Dict[str, np.ndarray]] = None) -> Callable:
    """Returns runner(batch of run-arg dicts) -> float64 array [len(batch), num_ticks + 1, len(REPORTER_NAMES)]
    executing the whole batch as one CrimeEngine (one replica per run).  Row 0 of a run is the state after setup
    (the row the reference's DataCollector collects at model.py:1042), row t the state after tick t."""
    from .engine import CrimeEngine

    def runner(batch: List[Dict]) -> np.ndarray:
        models, states = [], []
        for a in batch:
            m = ProtonOC(seed=a["seed"] if a["seed"] is not None else int.from_bytes(os.urandom(4), "little"))
            if a.get("source_override"):
                m.override_dict(a["source_override"])
            m.choose_intervention_setting()
            if a.get("initial_state") is not None:
                m.load_state(a["initial_state"])
            else:
                from . import synth
                m.load_state(synth.synthetic_state(m.initial_agents, seed=m.seed, template=template,
                                                   num_oc_persons=m.num_oc_persons,
                                                   likelihood_of_facilitators=m.likelihood_of_facilitators,
                                                   nat_propensity_m=m.nat_propensity_m,
                                                   nat_propensity_sigma=m.nat_propensity_sigma))
            models.append(m)
            states.append(m.state_dict())
        ticks = {m.num_ticks for m in models}
        if len(ticks) != 1:
            raise ValueError("all runs of one batch must have the same num_ticks")
        T = ticks.pop()
        n, stat, crim = CrimeEngine.capacity_for(states, headroom_criminal=max(200000, 16 * max(len(s["alive"]) for s in states)))
        phases = models[0].device_phases
        if any(m.device_phases != phases for m in models):
            raise ValueError("all runs of one batch must run the same device phases")
        # births add agent slots and links on the device (the same head room as a single run, ProtonOC._engine_ready)
        grow = (max(n // 2, 2000), max(stat, 1)) if phases else (0, 0)
        with CrimeEngine(len(batch), n + grow[0], 2 * max(stat, 1) + grow[1], crim, device=device) as E:
            from .params import make_demography
            m0 = models[0]   # the demography tables are per engine: one batch, one (nat_propensity, retirement_age) setting
            E.set_demography(make_demography(m0.nat_propensity_m, m0.nat_propensity_sigma, retirement_age=m0.retirement_age,
                                             number_weddings_mean=m0.number_weddings_mean))
            for r, (m, st) in enumerate(zip(models, states)):
                E.set_params(m._params(), replica=r)
                E.upload_state(r, st, wait=False)
            E.sync()   # input errors of the queued uploads surface here
            E.set_labour(m0._labour())   # likewise one set of labour-status tables per engine
            for r, st in enumerate(states):
                E.upload_children(r, st["number_of_children"])
                E.upload_education(r, st["max_education_level"], st["school_level"])
            E.set_phases(phases)
            E.crime_multiplier()   # ProtonOC.setup: model.py:1029-1030
            E.tendency()
            rep = np.zeros((len(batch), T + 1, len(REPORTER_NAMES)), np.float64)
            rep[:, 0] = E.report()
            rep[:, 1:] = E.run_ticks(1, T, [m._philox_key for m in models], reporters=True)
            cn, _ = E.counters()
            if cn[:, 9].any():
                raise RuntimeError("crime path error codes %s" % cn[:, 9].tolist())
        return rep

    return runner


def run_ensemble(args: List[Dict], runner: Callable, replicas_per_batch: int = 64, dist=None,
                 device=None) -> Optional[np.ndarray]:
    """Partition `args` over the ranks of `dist` (torch.distributed, already initialised, or None), run this
    rank's share in batches, gather [n_runs, T, K] on rank 0 (returns None on the other ranks)."""
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    mine = partition(len(args), world, rank)
    outs = []
    for b in range(0, len(mine), replicas_per_batch):
        idx = mine[b:b + replicas_per_batch]
        outs.append(runner([args[i] for i in idx]))
    local = np.concatenate(outs, axis=0) if outs else None
    if dist is None:
        return local
    import torch
    shape = torch.zeros(3, dtype=torch.int64)
    if local is not None:
        shape = torch.tensor(local.shape[0:3], dtype=torch.int64)
    if device is None:   # NCCL moves device tensors only; gloo (the CPU tests) host tensors
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    dev = device
    # ranks may hold different run counts: exchange the shapes, pad to the maximum, gather, trim
    shapes = [torch.zeros(3, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(shapes, shape.to(dev))
    T, K = int(max(s[1] for s in shapes)), int(max(s[2] for s in shapes))
    nmax = int(max(s[0] for s in shapes))
    buf = torch.zeros((nmax, T, K), dtype=torch.float64, device=dev)
    if local is not None:
        buf[:local.shape[0]] = torch.from_numpy(local).to(dev)
    gathered = [torch.zeros_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, gathered, dst=0)
    if rank != 0:
        return None
    full = np.zeros((len(args), T, K), np.float64)
    for r in range(world):
        idx = partition(len(args), world, r)
        full[idx] = gathered[r][:len(idx)].cpu().numpy()
    return full


def reporter_frame(rows: np.ndarray) -> pd.DataFrame:
    return pd.DataFrame(rows, columns=REPORTER_NAMES)


# ---------------------------------------------------------------------------------------------- run modes
class BaseMode:
    """run_modes.py:39-88"""

    def __init__(self, name, save_path, snapshot, alldata, randomstate):
        self.save_path, self.alldata, self.name = save_path, alldata, name
        self.snapshot_active, self.randomstate = snapshot, randomstate
        self.args: List[Dict] = []

    def run(self) -> None:
        self.args.append({"seed": self.randomstate, "source_override": None, "save_path": self.save_path,
                          "filename": self.name, "verbose": True, "snapshot": self.snapshot_active,
                          "alldata": self.alldata})
        self._single_run(self.args[0])

    def _single_run(self, args: Dict):
        if args["seed"] is None:
            args["seed"] = int.from_bytes(os.urandom(4), "little")
        model = ProtonOC(seed=args["seed"], collect_agents=args["alldata"])
        if args["source_override"] is not None:
            model.override_dict(args["source_override"])
        model.run(verbose=args["verbose"])
        return model.save_data(save_dir=args["save_path"], name=args["filename"], snapshot=args["snapshot"])


class OverrideMode(BaseMode):
    """run_modes.py:91-232.  `parallel` (the reference's process count) selects batched execution: all runs of
    this rank go through one engine per `parallel` replicas instead of one process per run."""

    def __init__(self, source_path, save_path, snapshot, alldata, randomstate, parallel, merge=False):
        super().__init__(None, save_path, snapshot, alldata, randomstate)
        self.source_path, self.parallel, self.merge = source_path, parallel, merge
        self.args = [run for sub in self.detect_files() for run in sub]

    def detect_files(self):
        plans = []
        if os.path.isfile(self.source_path):
            if self.source_path.endswith(".json"):
                plans.append(self.get_from_json(self.source_path))
            elif self.source_path.endswith(".xml"):
                plans.append(self.get_from_xml(self.source_path))
            else:
                raise Exception(self.source_path + " is not a valid file. \n Source file should be .json or .xml")
        else:
            for entry in os.scandir(self.source_path):
                if entry.path.endswith(".xml"):
                    plans.append(self.get_from_xml(entry.path))
                elif entry.path.endswith(".json"):
                    plans.append(self.get_from_json(entry.path))
        return plans

    def get_from_json(self, source):
        return runs_from_json(source, self.randomstate, self.save_path, self.snapshot_active, self.alldata, self.parallel)

    def get_from_xml(self, xml_file):
        return runs_from_xml(xml_file, self.randomstate, self.save_path, self.snapshot_active, self.alldata, self.parallel)

    def run(self, dist=None, device: int = 0, template=None) -> None:
        if self.parallel is None:
            for a in self.args:
                self._single_run(a)
            return
        # a run without a seed draws one (run_modes.py:80-81); here before the batch, so that every rank and the
        # frames written below see the same value
        seeds = [a["seed"] if a["seed"] is not None else int.from_bytes(os.urandom(4), "little") for a in self.args]
        if dist is not None:
            box = [seeds]
            dist.broadcast_object_list(box, src=0)
            seeds = box[0]
        for a, sd in zip(self.args, seeds):
            a["seed"] = sd
        rep = run_ensemble(self.args, gpu_batch_runner(device, template), replicas_per_batch=int(self.parallel), dist=dist)
        if rep is None:
            return
        for a, rows in zip(self.args, rep):
            # the same pickle as _single_run writes: the reference's (T+1) x 80 model-vars frame (extra.py:246-274)
            m = ProtonOC(seed=a["seed"], collect_agents=False)
            if a["source_override"] is not None:
                m.override_dict(a["source_override"])
            m.choose_intervention_setting()
            df = m.frame_from_device_rows(rows)
            if a["snapshot"]:
                df = df.iloc[list(a["snapshot"])]
            with open(os.path.join(a["save_path"], a["filename"] + ".pkl"), "wb") as f:
                pickle.dump(df, f)
