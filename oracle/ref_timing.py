"""Time the UNMODIFIED reference (LABSS/PyPROTON-OC, through oracle/refshim) on this machine's host cores.

TEST / BENCH INFRASTRUCTURE: used by bench.py's cpu_baseline leg and by tools/time_reference.py.  One process = one core
(the reference is single-threaded; its only parallelism is one process per run, run_modes.py:221-232).

    whole step  = ProtonOC.step() (model.py:226-288), setup excluded          -> the drop-in metric of SURVEY.md 8(d)
    hot path    = reset_oc_embeddedness + commit_crimes + the yearly calculate_criminal_tendency /
                  calculate_crime_multiplier (model.py:272-273, 249-250)       -> the part this repository replaces
"""
from __future__ import annotations

import time

from . import ref_harness as H

HOT = ["reset_oc_embeddedness", "commit_crimes", "calculate_criminal_tendency", "calculate_crime_multiplier"]


def time_reference(n_agents: int, ticks: int, seed: int = 42, overrides=None, max_seconds: float = None) -> dict:
    ref_model, _, _ = H.load_reference()
    m = ref_model.ProtonOC(seed=seed)
    if overrides:
        m.override_dict(overrides)
    t0 = time.perf_counter()
    m.setup(n_agents)
    setup_s = time.perf_counter() - t0
    acc = {"hot": 0.0}

    def wrap(name):
        orig = getattr(m, name)

        def f(*a, **k):
            t = time.perf_counter()
            try:
                return orig(*a, **k)
            finally:
                acc["hot"] += time.perf_counter() - t
        setattr(m, name, f)
    for name in HOT:
        wrap(name)
    t0 = time.perf_counter()
    done = 0
    for _ in range(ticks):
        m.step()
        done += 1
        if max_seconds is not None and time.perf_counter() - t0 > max_seconds:
            break
    step_s = time.perf_counter() - t0
    return {"n_agents": n_agents, "ticks": done, "seed": seed, "setup_s": setup_s, "step_s": step_s, "hot_s": acc["hot"],
            "agent_ticks_per_s_whole_step": n_agents * done / step_s,
            "agent_ticks_per_s_hot_path": n_agents * done / max(acc["hot"], 1e-9),
            "hot_share": acc["hot"] / step_s, "number_crimes": int(m.number_crimes),
            "current_oc_members": int(m.current_oc_members), "reference_root": H.REFERENCE_ROOT}
