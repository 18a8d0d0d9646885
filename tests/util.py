"""Shared helpers for the parity tests: fixture loading, hand-built states, tie analysis."""
from __future__ import annotations

import glob
import os
from typing import Dict, Iterable, List, Tuple

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
LAYERS = ["sibling", "offspring", "parent", "partner", "household",
          "friendship", "criminal", "professional", "school"]
PARAM_KEYS = ["nat_propensity_m", "nat_propensity_sigma", "nat_propensity_threshold",
              "number_crimes_yearly_per10k", "number_arrests_per_year", "punishment_length",
              "threshold_use_facilitators", "facilitator_repression_multiplier", "good_guy_threshold",
              "max_accomplice_radius", "oc_embeddedness_radius", "facilitator_repression",
              "oc_boss_repression", "ticks_between_intervention", "intervention_start", "intervention_end",
              "this_is_a_big_crime", "ticks_per_year"]


def golden_files(kind: str) -> List[str]:
    return sorted(glob.glob(os.path.join(GOLDEN, kind + "_*.npz")))


def load_fixture(path: str) -> Dict[str, np.ndarray]:
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


def pre_state(fx: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    return {k[4:]: v for k, v in fx.items() if k.startswith("pre_")}


def fixture_params(fx) -> dict:
    d = {k: float(fx["param_" + k]) for k in PARAM_KEYS}
    for k in ["facilitator_repression", "oc_boss_repression"]:
        d[k] = bool(d[k])
    return d


def fixture_replay(fx) -> dict:
    return {k[len("replay_"):]: (float(v) if v.ndim == 0 else v) for k, v in fx.items() if k.startswith("replay_")}


def fixture_co_offenders(fx):
    return [(int(r[0]), float(r[1])) for r in fx["num_co_offenders_dist"]]


def ref_groups(fx) -> List[np.ndarray]:
    off, mem = fx["group_off"], fx["group_mem"]
    return [mem[off[i]:off[i + 1]] for i in range(len(off) - 1)]


def ref_keys(fx, crime: int) -> Dict[int, float]:
    a, b = fx["key_off"][crime], fx["key_off"][crime + 1]
    return dict(zip(fx["key_slot"][a:b].tolist(), fx["key_val"][a:b].tolist()))


def ref_rings(fx, crime: int) -> List[Tuple[int, np.ndarray]]:
    out = []
    for r in range(fx["ring_off"][crime], fx["ring_off"][crime + 1]):
        out.append((int(fx["ring_d"][r]), fx["ring_mem"][fx["ring_mem_off"][r]:fx["ring_mem_off"][r + 1]]))
    return out


def canonical_group(g: Iterable[int]) -> List[int]:
    g = list(int(x) for x in g)
    return sorted(g[:-1]) + [g[-1]]


def tie_explains(fx, crime: int, mine: Iterable[int], rel: float = 0.0) -> bool:
    """True when `mine` and the reference's group for this crime differ only by a permutation among ranked
    candidates whose sort keys are equal (`rel` > 0: equal within that relative tolerance, used for
    OC initiators whose keys carry a 1/dist sum that the reference adds up in set order).  The ranked accomplices are
    the members with a recorded key; a member without one is the facilitator drawn after the rings
    (entities.py:439-445), whose pool -- the facilitators near the accomplices -- follows from who was picked, so a
    tie among the ranked ones may add, drop or change it."""
    ref = [int(x) for x in ref_groups(fx)[crime]]
    mine = [int(x) for x in mine]
    if ref[-1] != mine[-1]:
        return False
    keys = ref_keys(fx, crime)
    ra, ma = [m for m in ref[:-1] if m in keys], [m for m in mine[:-1] if m in keys]
    if len(ref) - 1 - len(ra) > 1 or len(mine) - 1 - len(ma) > 1:
        return False
    fac = pre_state_facilitators(fx)
    if any(not fac[m] for m in ref[:-1] + mine[:-1] if m not in keys):
        return False
    if sorted(ra) == sorted(ma):
        return False    # the ranked picks agree, so the rest must as well: not a tie
    # a facilitator among the ranked picks raises the target by one (Q6): compare the common prefix of the sorted keys
    ka, kb = sorted(keys[m] for m in ra), sorted(keys[m] for m in ma)
    if len(ka) != len(kb) and not any(fac[m] for m in ra + ma):
        return False
    return all(abs(x - y) <= rel * max(abs(x), abs(y)) for x, y in zip(ka, kb))


def pre_state_facilitators(fx):
    return fx["pre_facilitator"]


# ---------------------------------------------------------------- hand-built states
def build_state(n: int, links: Dict[str, List[Tuple[int, int]]], directed: Dict[str, List[Tuple[int, int]]] = None,
                co_off: Dict[Tuple[int, int], int] = None, **cols) -> Dict[str, np.ndarray]:
    """State dict from undirected `links` per layer name (parent/offspring via
    `directed['offspring'] = [(parent, child)]`, mirrored into 'parent')."""
    adj = {l: [set() for _ in range(n)] for l in LAYERS}
    for l, pairs in (links or {}).items():
        for a, b in pairs:
            adj[l][a].add(b)
            adj[l][b].add(a)
    for l, pairs in (directed or {}).items():
        for a, b in pairs:
            adj[l][a].add(b)
            if l == "offspring":
                adj["parent"][b].add(a)
    st = {
        "unique_id": np.arange(n, dtype=np.int64), "alive": np.ones(n, np.uint8), "age": np.full(n, 30, np.int32),
        "birth_tick": np.full(n, -360, np.int32), "male": np.ones(n, np.uint8), "job_level": np.zeros(n, np.int8),
        "education_level": np.zeros(n, np.int8), "wealth_level": np.ones(n, np.int8),
        "propensity": np.ones(n, np.float64), "oc_member": np.zeros(n, np.uint8),
        "facilitator": np.zeros(n, np.uint8), "prisoner": np.zeros(n, np.uint8), "retired": np.zeros(n, np.uint8),
        "has_job": np.zeros(n, np.uint8), "has_school": np.zeros(n, np.uint8),
        "target_of_intervention": np.zeros(n, np.uint8), "num_crimes_committed": np.zeros(n, np.int32),
        "num_crimes_committed_this_tick": np.zeros(n, np.int32), "criminal_tendency": np.zeros(n, np.float64),
        "sentence_countdown": np.zeros(n, np.float64), "new_recruit": np.full(n, -2, np.int32),
        "arrest_weight": np.zeros(n, np.float64), "father": np.full(n, -1, np.int32)}
    for k, v in cols.items():
        st[k] = np.asarray(v, dtype=st[k].dtype)
    for li, l in enumerate(LAYERS):
        rowptr = np.zeros(n + 1, np.int32)
        col: List[int] = []
        for a in range(n):
            col.extend(sorted(adj[l][a]))
            rowptr[a + 1] = len(col)
        st["rowptr_%d" % li] = rowptr
        st["col_%d" % li] = np.array(col, dtype=np.int32)
    cnt = []
    for a in range(n):
        for b in sorted(adj["criminal"][a]):
            cnt.append((co_off or {}).get((a, b), (co_off or {}).get((b, a), 0)))
    st["co_off"] = np.array(cnt, dtype=np.int32)
    return st


def micro_net_state() -> Dict[str, np.ndarray]:
    """The 18-agent multiplex net of the reference's test_micro_net
    (protonoc/test/test_protonoc.py:577-625)."""
    n = 18
    i = np.arange(n)
    links = {"partner": [(0, 1), (7, 8)], "friendship": [(0, 2), (6, 5), (8, 9), (13, 15)],
             "professional": [(0, 3), (10, 9), (10, 11), (11, 12), (12, 13), (12, 14), (12, 16), (14, 15)],
             "school": [(0, 4), (10, 13)], "criminal": [(0, 5)],
             "household": [(14, 17), (16, 12), (16, 17)]}
    directed = {"offspring": [(5, 0), (7, 9)]}
    oc = np.zeros(n, np.uint8); oc[[5, 17]] = 1
    fac = np.zeros(n, np.uint8); fac[[9, 3]] = 1
    return build_state(
        n, links, directed, co_off={(0, 5): 5},
        male=(i % 2 == 0), age=i * 149 % 45 + 10, education_level=i * 149 % 4, wealth_level=i * 149 % 4,
        job_level=i * 149 % 4, criminal_tendency=(i * 149 % 100) / 100, propensity=(i * 149 % 100) / 100,
        oc_member=oc, facilitator=fac)
