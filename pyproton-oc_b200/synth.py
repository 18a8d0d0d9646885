"""Synthetic initial states for the crime path (SURVEY.md §8d config 4 and the ensemble).

The reference's population synthesis (ProtonOC.setup, model.py:1003-1141) is out of scope and super-linear in
the number of agents; for sizes it cannot produce (100k agents) or for ensembles that need many distinct
initial states, this module writes a state directly in the SoA + 9-layer CSR form the engine uploads.

It is NOT a restatement of the reference's setup.  Agent attributes are bootstrapped from a template state
(normally the reference's own 10k setup state, so the joint age/sex/job/education/wealth distribution is the
reference's), and the nine layers are generated with per-layer mean out-degrees matched to the reference at 10k
agents (friendship 8.9, school 3.3, household 2.2, professional 1.9, sibling 0.47,
offspring = parent^T 0.50, partner 0.38) and the OC members form a complete criminal graph with
num_co_offenses = 1 (model.py:701-706)."""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

# inputs/palermo/data/household_size_dist.csv (sizes 1..9)
HOUSEHOLD_SIZE_P = np.array([0.2820801951962931, 0.25447100342132306, 0.2052870794062523, 0.1803756912881339,
                             0.05575717457983921, 0.015016875704757909, 0.004548935695290582,
                             0.0018280409943003605, 6.35003713809599e-4])
LAYERS = ["sibling", "offspring", "parent", "partner", "household", "friendship", "criminal", "professional", "school"]


def _csr(n: int, src: np.ndarray, dst: np.ndarray):
    """Directed CSR with ascending, de-duplicated rows and no self loops."""
    keep = src != dst
    src, dst = src[keep].astype(np.int64), dst[keep].astype(np.int64)
    key = np.unique(src * n + dst)
    s, d = key // n, key % n
    rowptr = np.zeros(n + 1, np.int32)
    np.add.at(rowptr, s + 1, 1)
    return np.cumsum(rowptr, dtype=np.int64).astype(np.int32), d.astype(np.int32)


def _sym(a, b):
    return np.concatenate([a, b]), np.concatenate([b, a])


def _groups(members: np.ndarray, sizes: np.ndarray):
    """Split `members` into consecutive groups with the given sizes; returns group id per member."""
    gid = np.repeat(np.arange(len(sizes)), sizes)[:len(members)]
    return gid


def _clique_edges(members: np.ndarray, gid: np.ndarray, max_links: Optional[int], rng):
    """All ordered pairs inside each group (optionally a random subset of at most max_links partners per member)."""
    order = np.argsort(gid, kind="stable")
    m, g = members[order], gid[order]
    starts = np.flatnonzero(np.r_[True, g[1:] != g[:-1]])
    sizes = np.diff(np.r_[starts, len(g)])
    src, dst = [], []
    for size in np.unique(sizes):
        if size < 2:
            continue
        blocks = starts[sizes == size]
        idx = blocks[:, None] + np.arange(size)[None, :]          # [nb, size]
        mem = m[idx]
        if max_links is None or size - 1 <= max_links:
            a = np.repeat(mem, size, axis=1)
            b = np.tile(mem, (1, size))
            src.append(a.ravel()); dst.append(b.ravel())
        else:
            for _ in range(max_links // 2 + 1):
                perm = np.argsort(rng.random(mem.shape), axis=1)
                b = np.take_along_axis(mem, perm, axis=1)
                src.append(mem.ravel()); dst.append(b.ravel())
    if not src:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    return np.concatenate(src), np.concatenate(dst)


def synthetic_state(n_agents: int, seed: int = 42, template: Optional[Dict[str, np.ndarray]] = None,
                    num_oc_persons: int = 30, likelihood_of_facilitators: float = 0.005,
                    nat_propensity_m: float = 1.0, nat_propensity_sigma: float = 0.25) -> Dict[str, np.ndarray]:
    rng = np.random.default_rng(seed)
    n = int(n_agents)
    # ---- attributes
    if template is not None:
        alive_t = np.flatnonzero(template["alive"])
        pick = alive_t[rng.integers(0, len(alive_t), n)]
        age = template["age"][pick].astype(np.int32)
        male = template["male"][pick].astype(np.uint8)
        job = template["job_level"][pick].astype(np.int8)
        edu = template["education_level"][pick].astype(np.int8)
        wealth = template["wealth_level"][pick].astype(np.int8)
        has_job = template["has_job"][pick].astype(np.uint8)
        has_school = template["has_school"][pick].astype(np.uint8)
        retired = template["retired"][pick].astype(np.uint8) if "retired" in template else (age >= 65).astype(np.uint8)
    else:
        age = np.minimum(rng.triangular(0, 35, 100, n), 99).astype(np.int32)
        male = (rng.random(n) < 0.485).astype(np.uint8)
        adult = (age >= 18) & (age < 65)
        has_school = ((age >= 6) & (age < 19)).astype(np.uint8)
        has_job = (adult & (rng.random(n) < 0.42)).astype(np.uint8)
        job = np.where(has_job > 0, rng.integers(2, 5, n), np.where(adult & (rng.random(n) < 0.2), 1, 0)).astype(np.int8)
        edu = np.where(age < 6, 0, np.minimum(rng.integers(0, 5, n), (age // 6))).astype(np.int8)
        wealth = rng.integers(1, 6, n).astype(np.int8)
        retired = (age >= 65).astype(np.uint8)
    # what graduate_and_enter_jobmarket reads (model.py:1351-1380): the level of the school an agent attends and how far it may go
    if template is not None and "max_education_level" in template:
        school_level = np.where(has_school > 0, template["school_level"][pick], 0).astype(np.int8)
        max_edu = template["max_education_level"][pick].astype(np.int8)
    else:
        from . import tables
        school_level = np.zeros(n, np.int8)
        for lv, (a0, a1) in tables.EDUCATION_LEVELS.items():
            school_level[(has_school > 0) & (age >= a0) & (age <= a1 + 1)] = lv
        has_school = (school_level > 0).astype(np.uint8)   # a student attends a school of some level
        max_edu = np.maximum(np.maximum(edu, school_level), rng.integers(1, len(tables.EDUCATION_LEVELS) + 1, n)).astype(np.int8)
    order = np.argsort(rng.random(n))  # slots carry no meaning beyond tie-breaks
    propensity = np.exp(nat_propensity_m + nat_propensity_sigma * rng.standard_normal(n))  # model.py:1216
    birth_tick = (0 - age * 12).astype(np.int32)                                            # entities.py:271
    # ---- households: consecutive blocks of a random permutation, sizes from the Palermo table
    sizes = rng.choice(np.arange(1, 10), size=n, p=HOUSEHOLD_SIZE_P / HOUSEHOLD_SIZE_P.sum())
    sizes = sizes[:np.searchsorted(np.cumsum(sizes), n) + 1]
    hh = _groups(order, sizes)
    hs, hd = _clique_edges(order, hh, None, rng)
    # family roles inside a household: the two oldest adults of opposite sex are partners, members at least
    # 15 years younger than the oldest are offspring of the adults, offspring of one household are siblings
    oldest = np.zeros(hh.max() + 1, np.int32)
    np.maximum.at(oldest, hh, age[order])
    is_child = age[order] <= oldest[hh] - 15
    cs, cd = hs, hd  # all ordered pairs (a, b) within households
    inv = np.empty(n, np.int64); inv[order] = np.arange(n)
    child_a, child_b = is_child[inv[cs]], is_child[inv[cd]]
    sib_s, sib_d = cs[child_a & child_b], cd[child_a & child_b]
    par_mask = (~child_a) & child_b & (age[cs] >= 18)
    off_s, off_d = cs[par_mask], cd[par_mask]                       # parent -> child
    adult_pair = (~child_a) & (~child_b) & (male[cs] != male[cd]) & (age[cs] >= 18) & (age[cd] >= 18) & \
                 (np.abs(age[cs] - age[cd]) <= 15)
    # keep at most one partner per agent: the first listed
    ps, pd = cs[adult_pair], cd[adult_pair]
    if len(ps):
        first = np.unique(ps, return_index=True)[1]
        ps, pd = ps[first], pd[first]
        ok = np.zeros(n, np.int64) - 1
        ok[ps] = pd
        mutual = ok[pd] == ps
        ps, pd = ps[mutual], pd[mutual]
    # ---- friendship: Watts-Strogatz ring (k = 8, rewiring 0.1) over a random ring order (model.py:720-760)
    ring = np.argsort(rng.random(n))
    pos = np.arange(n)
    fs, fd = [], []
    for off in range(1, 5):
        a, b = ring[pos], ring[(pos + off) % n]
        rew = rng.random(n) < 0.1
        b = np.where(rew, ring[rng.integers(0, n, n)], b)
        fs.append(a); fd.append(b)
    fs, fd = _sym(np.concatenate(fs), np.concatenate(fd))
    # ---- school: classes of 28 same-age students, all class-mates linked (model.py:980-1001)
    students = np.flatnonzero(has_school)
    students = students[np.argsort(age[students], kind="stable")]
    sg = np.arange(len(students)) // 28
    ss, sd = _clique_edges(students, sg, None, rng)
    ss, sd = _sym(ss, sd)
    # ---- professional: employers with heavy-tailed sizes, up to 20 colleagues (model.py:1115-1141)
    workers = np.flatnonzero(has_job)
    workers = workers[np.argsort(rng.random(len(workers)))]
    emp_sizes = np.minimum(np.maximum((rng.pareto(1.2, len(workers)) * 2 + 1).astype(np.int64), 1), 200)
    emp_sizes = emp_sizes[:np.searchsorted(np.cumsum(emp_sizes), len(workers)) + 1]
    wg = _groups(workers, emp_sizes)
    ws, wd = _clique_edges(workers, wg, 20, rng)
    ws, wd = _sym(ws, wd)
    # ---- OC members: complete criminal graph, one co-offence each (model.py:669-706)
    n_oc = int(np.ceil(num_oc_persons * n / 10000))
    adults_m = np.flatnonzero((age >= 18) & (age < 60) & (male > 0))
    oc_ids = rng.choice(adults_m, size=min(n_oc, len(adults_m)), replace=False)
    oc = np.zeros(n, np.uint8); oc[oc_ids] = 1
    a = np.repeat(oc_ids, len(oc_ids)); b = np.tile(oc_ids, len(oc_ids))
    fac = ((oc == 0) & (age > 18) & (rng.random(n) < likelihood_of_facilitators)).astype(np.uint8)  # model.py:353
    st: Dict[str, np.ndarray] = {
        "unique_id": np.arange(n, dtype=np.int64), "alive": np.ones(n, np.uint8), "age": age, "birth_tick": birth_tick,
        "male": male, "job_level": job, "education_level": edu, "wealth_level": wealth, "propensity": propensity,
        "oc_member": oc, "facilitator": fac, "prisoner": np.zeros(n, np.uint8), "retired": retired,
        "has_job": has_job, "has_school": has_school, "target_of_intervention": np.zeros(n, np.uint8),
        "num_crimes_committed": np.zeros(n, np.int32), "num_crimes_committed_this_tick": np.zeros(n, np.int32),
        "criminal_tendency": np.zeros(n, np.float64), "sentence_countdown": np.zeros(n, np.float64),
        "new_recruit": np.full(n, -2, np.int32), "arrest_weight": np.zeros(n, np.float64),
        "max_education_level": max_edu, "school_level": school_level}
    father = np.full(n, -1, np.int32)
    fm = male[off_s] > 0
    father[off_d[fm]] = off_s[fm]
    st["father"] = father
    layers = {"sibling": (sib_s, sib_d), "offspring": (off_s, off_d), "parent": (off_d, off_s),
              "partner": _sym(ps, pd) if len(ps) else (ps, pd), "household": (hs, hd), "friendship": (fs, fd),
              "criminal": (a, b), "professional": (ws, wd), "school": (ss, sd)}
    for li, name in enumerate(LAYERS):
        s, d = layers[name]
        st["rowptr_%d" % li], st["col_%d" % li] = _csr(n, np.asarray(s), np.asarray(d))
    st["co_off"] = np.ones(len(st["col_6"]), np.int32)
    return st


def layer_degrees(st: Dict[str, np.ndarray]) -> Dict[str, float]:
    n = len(st["alive"])
    return {name: len(st["col_%d" % li]) / n for li, name in enumerate(LAYERS)}
