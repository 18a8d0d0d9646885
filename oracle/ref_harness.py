"""Harness that runs the UNMODIFIED reference (PyPROTON-OC) as the parity oracle.

TEST INFRASTRUCTURE ONLY.  Works only where `/root/reference` exists (the build
container); nothing on the GPU box imports this module.  It is used by
`oracle/make_golden.py` to generate the committed fixtures under
`tests/golden/` and by the container-only cross-checks in `tests/`.

What it does
  * loads the reference through the Mesa/prettytable stand-ins in
    `oracle/refshim/` (reference import quirk: model.py:35-36 uses top-level
    `from entities import ...`, so `protonoc/simulator` itself goes on sys.path);
  * `export_state(model)` flattens the Person object graph into the SoA + 9
    directed CSR layers that `include/protonoc_b200.h` uploads
    (layer order = Person.network_names, entities.py:38-47);
  * `record_commit_crimes(model)` runs the reference's own
    `reset_oc_embeddedness(); commit_crimes()` (model.py:272-273) while
    recording every random draw (raw uniforms recovered from a cloned PCG64),
    every crime (initiator, k, ring sets, candidate keys, group) and every
    oc_embeddedness evaluation (accumulated value as the reference computed it
    plus the value on a fresh meta_graph).
"""
from __future__ import annotations

import copy
import os
import sys
from typing import Dict, List

import numpy as np

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_reference() -> str:
    """The reference tree: $PROTONOC_REFERENCE, /root/reference (build container), or the unmodified install under
    baseline/_ref (`pip install --no-deps --target baseline/_ref /root/reference`, done by __graft_entry__.build();
    git-ignored, travels to the GPU box with the snapshot)."""
    cands = [os.environ.get("PROTONOC_REFERENCE"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "protonoc", "simulator", "model.py")):
            return c
    return "/root/reference"


REFERENCE_ROOT = _find_reference()
_SIM = os.path.join(REFERENCE_ROOT, "protonoc", "simulator")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refshim")

LAYERS = ["sibling", "offspring", "parent", "partner", "household",
          "friendship", "criminal", "professional", "school"]


def reference_available() -> bool:
    return os.path.isfile(os.path.join(_SIM, "model.py"))


def load_reference():
    """Import the reference's `model`, `entities`, `extra` modules."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True
    for p in (_SIM, _SHIM):
        if p not in sys.path:
            sys.path.insert(0, p)
    import warnings
    warnings.filterwarnings("ignore")
    import model as ref_model      # noqa: E402
    import entities as ref_entities  # noqa: E402
    import extra as ref_extra      # noqa: E402
    return ref_model, ref_entities, ref_extra


# --------------------------------------------------------------------------
# state export
# --------------------------------------------------------------------------

def _closure(model) -> list:
    """Scheduled agents plus every Person still reachable through a neighbour
    set or a `father` pointer (dead agents stay referenced when a link was
    cleared one-sidedly before death: entities.py:601-602, 652-661)."""
    sched = model.schedule.agents
    seen = {a.unique_id: a for a in sched}
    stack = list(sched)
    while stack:
        a = stack.pop()
        nxt = []
        for net in LAYERS:
            nxt.extend(a.neighbors[net])
        if a.father is not None:
            nxt.append(a.father)
        for b in nxt:
            if b.unique_id not in seen:
                seen[b.unique_id] = b
                stack.append(b)
    return [seen[k] for k in sorted(seen)]


def export_state(model, extra_agents=()) -> Dict[str, np.ndarray]:
    """Flatten the reference model into SoA columns + 9 directed CSR layers.

    Slot order is ascending unique_id, which for scheduled agents equals the
    Mesa schedule order the reference iterates cells in (model.py:1430).
    """
    alive_ids = {a.unique_id for a in model.schedule.agents}
    agents = _closure(model)
    known = {a.unique_id for a in agents}
    for a in extra_agents:
        if a.unique_id not in known:
            agents.append(a)
            known.add(a.unique_id)
    agents.sort(key=lambda a: a.unique_id)
    # extra (unscheduled) agents may reference further unscheduled ones
    grew = True
    while grew:
        grew = False
        for a in list(agents):
            for net in LAYERS:
                for b in a.neighbors[net]:
                    if b.unique_id not in known:
                        agents.append(b)
                        known.add(b.unique_id)
                        grew = True
        agents.sort(key=lambda a: a.unique_id)
    n = len(agents)
    slot = {a.unique_id: i for i, a in enumerate(agents)}
    st: Dict[str, np.ndarray] = {}
    st["unique_id"] = np.array([a.unique_id for a in agents], dtype=np.int64)
    st["alive"] = np.array([a.unique_id in alive_ids for a in agents], dtype=np.uint8)
    st["age"] = np.array([int(a.age) for a in agents], dtype=np.int32)
    st["birth_tick"] = np.array([int(a.birth_tick) for a in agents], dtype=np.int32)
    st["male"] = np.array([bool(a.gender_is_male) for a in agents], dtype=np.uint8)
    st["job_level"] = np.array([int(a.job_level) for a in agents], dtype=np.int8)
    st["education_level"] = np.array([int(a.education_level) for a in agents], dtype=np.int8)
    st["wealth_level"] = np.array([int(a.wealth_level) for a in agents], dtype=np.int8)
    st["propensity"] = np.array([float(a.propensity) for a in agents], dtype=np.float64)
    st["oc_member"] = np.array([bool(a.oc_member) for a in agents], dtype=np.uint8)
    st["facilitator"] = np.array([bool(a.facilitator) for a in agents], dtype=np.uint8)
    st["prisoner"] = np.array([bool(a.prisoner) for a in agents], dtype=np.uint8)
    st["retired"] = np.array([bool(a.retired) for a in agents], dtype=np.uint8)
    st["has_job"] = np.array([a.my_job is not None for a in agents], dtype=np.uint8)
    st["has_school"] = np.array([a.my_school is not None for a in agents], dtype=np.uint8)
    # graduate_and_enter_jobmarket (model.py:1351-1380) reads these two; a school's identity never matters to it
    st["max_education_level"] = np.array([int(a.max_education_level) for a in agents], dtype=np.int8)
    st["school_level"] = np.array([int(a.my_school.diploma_level) if a.my_school is not None else 0 for a in agents],
                                  dtype=np.int8)
    st["target_of_intervention"] = np.array([bool(a.target_of_intervention) for a in agents], dtype=np.uint8)
    st["num_crimes_committed"] = np.array([int(a.num_crimes_committed) for a in agents], dtype=np.int32)
    st["num_crimes_committed_this_tick"] = np.array(
        [int(a.num_crimes_committed_this_tick) for a in agents], dtype=np.int32)
    st["criminal_tendency"] = np.array([float(a.criminal_tendency) for a in agents], dtype=np.float64)
    st["sentence_countdown"] = np.array([float(a.sentence_countdown) for a in agents], dtype=np.float64)
    st["new_recruit"] = np.array([int(a.new_recruit) for a in agents], dtype=np.int32)
    st["arrest_weight"] = np.array([float(a.arrest_weight) for a in agents], dtype=np.float64)
    st["father"] = np.array([slot[a.father.unique_id] if a.father is not None else -1
                             for a in agents], dtype=np.int32)
    st["number_of_children"] = np.array([int(a.number_of_children) for a in agents], dtype=np.int32)
    for li, net in enumerate(LAYERS):
        rowptr = np.zeros(n + 1, dtype=np.int32)
        cols: List[int] = []
        cnt: List[int] = []
        for i, a in enumerate(agents):
            row = sorted(slot[b.unique_id] for b in a.neighbors[net])
            cols.extend(row)
            if net == "criminal":
                inv = {slot[b.unique_id]: b for b in a.neighbors[net]}
                # a criminal neighbour without a num_co_offenses entry counts 0
                # (model.py:1531-1533)
                cnt.extend(int(a.num_co_offenses.get(inv[c], 0)) for c in row)
            rowptr[i + 1] = len(cols)
        st["rowptr_%d" % li] = rowptr
        st["col_%d" % li] = np.array(cols, dtype=np.int32)
        if net == "criminal":
            st["co_off"] = np.array(cnt, dtype=np.int32)
    return st


def model_scalars(model) -> Dict[str, float]:
    names = ["tick", "crime_multiplier", "number_crimes", "crime_size_fails", "facilitator_crimes",
             "facilitator_fails", "big_crime_from_small_fish", "number_offspring_recruited_this_tick",
             "number_protected_recruited_this_tick", "people_jailed",
             "number_law_interventions_this_tick", "max_accomplice_radius", "oc_embeddedness_radius",
             "threshold_use_facilitators", "number_crimes_yearly_per10k", "number_arrests_per_year",
             "punishment_length", "facilitator_repression", "facilitator_repression_multiplier",
             "oc_boss_repression", "nat_propensity_m", "nat_propensity_sigma",
             "nat_propensity_threshold", "ticks_between_intervention", "intervention_start",
             "intervention_end", "this_is_a_big_crime", "good_guy_threshold", "ticks_per_year"]
    return {k: float(getattr(model, k)) for k in names}


# --------------------------------------------------------------------------
# RNG recording
# --------------------------------------------------------------------------

class RecordingRNG:
    """Proxy around the model's numpy Generator that records, for every call on
    the hot path, the raw uniform doubles it consumed.  The uniforms are
    recovered by replaying the call's documented algorithm on a cloned PCG64
    and checking that result and post-state agree."""

    def __init__(self, gen: np.random.Generator):
        self._gen = gen
        self.log: List[dict] = []

    def _clone(self) -> np.random.Generator:
        bg = type(self._gen.bit_generator)()
        bg.state = copy.deepcopy(self._gen.bit_generator.state)
        return np.random.Generator(bg)

    def random(self, *a, **kw):
        r = self._gen.random(*a, **kw)
        self.log.append({"op": "random", "u": np.atleast_1d(np.asarray(r, dtype=np.float64)).copy()})
        return r

    def choice(self, a, size=None, replace=True, p=None, **kw):
        clone = self._clone()
        res = self._gen.choice(a, size=size, replace=replace, p=p, **kw)
        n_pop = len(a)
        rec = {"op": "choice", "n_pop": n_pop, "size": size, "replace": replace}
        if p is not None and not replace:
            p_arr = np.array(p, dtype=np.float64)
            sz = int(size)
            us: List[float] = []
            found: List[int] = []
            pw = p_arr.copy()
            while len(found) < sz:
                x = clone.random((sz - len(found),))
                us.extend(x.tolist())
                if found:
                    pw[np.array(found)] = 0
                cdf = np.cumsum(pw)
                cdf /= cdf[-1]
                new = cdf.searchsorted(x, side="right")
                _, ui = np.unique(new, return_index=True)
                ui.sort()
                found.extend(new.take(ui).tolist())
            rec["u"] = np.array(us, dtype=np.float64)
            rec["idx"] = np.array(found, dtype=np.int64)
            rec["p"] = p_arr
            assert clone.bit_generator.state == self._gen.bit_generator.state, "choice emulation drifted"
        elif p is None and size is None:
            idx = int(clone.integers(0, n_pop))
            rec["idx"] = np.array([idx], dtype=np.int64)
            rec["res_id"] = int(getattr(res, "unique_id", -1))
            assert clone.bit_generator.state == self._gen.bit_generator.state, "choice emulation drifted"
        else:
            rec["unmodelled"] = True
        self.log.append(rec)
        return res

    def __getattr__(self, name):
        return getattr(self._gen, name)


# --------------------------------------------------------------------------
# recording one commit_crimes call
# --------------------------------------------------------------------------

def fresh_embeddedness(agent, nx, emb_fn=None) -> float:
    """Reference oc_embeddedness evaluated on a fresh meta_graph (H1)."""
    model = agent.model
    saved_graph, saved_cache = model.meta_graph, agent.cached_oc_embeddedness
    model.meta_graph = nx.Graph()
    agent.cached_oc_embeddedness = None
    try:
        v = emb_fn(agent) if emb_fn is not None else agent.oc_embeddedness()
    finally:
        model.meta_graph = saved_graph
        agent.cached_oc_embeddedness = saved_cache
    return float(v)


def record_commit_crimes(model) -> dict:
    """Run the reference's reset_oc_embeddedness + commit_crimes, recording
    everything a replay needs.  The model is advanced exactly as the reference
    would advance it (same RNG stream position afterwards)."""
    ref_model, ref_entities, ref_extra = load_reference()
    import networkx as nx
    Person = ref_entities.Person
    rec: dict = {"crimes": [], "emb": [], "arrested": [], "emb_over": []}
    rng = RecordingRNG(model.random)
    real_rng = model.random
    model.random = rng

    orig_find = Person.find_accomplices
    orig_emb = Person.oc_embeddedness
    orig_atd = Person.agents_at_distance
    orig_cw = Person.candidates_weight
    orig_caught = Person.get_caught
    orig_uml = model.update_meta_links
    cur: dict = {}
    fresh_flag = {"on": False}

    def update_meta_links(agents):
        """model.py:1515-1536 unchanged; afterwards note, for every directed entry a->b inside the ball whose stored
        meta-edge weight is NOT 1/w(a->b), that the other direction's multiplicity won (nx.Graph.add_edge keeps the
        direction the set iteration visited last): the one place where the reference's own value depends on
        CPython set order.  Recorded as (evaluation index, a, b, multiplicity in use)."""
        orig_uml(agents)
        if fresh_flag["on"]:
            return
        j = len(rec["emb"])
        G = model.meta_graph
        for a in agents:
            for b in a.agents_in_radius(1):
                if b is a or b not in agents:
                    continue
                w = 0
                for net in Person.network_names:
                    if b in a.neighbors.get(net):
                        if net == "criminal":
                            if b in a.num_co_offenses:
                                w += a.num_co_offenses[b]
                        else:
                            w += 1
                stored = G[a.unique_id][b.unique_id]["weight"]
                if stored != 1 / w:
                    rec["emb_over"].append((j, a.unique_id, b.unique_id, int(round(1 / stored))))

    def find_accomplices(self, n):
        cur.clear()
        cur.update({"initiator": self.unique_id, "k": int(n), "rings": [], "keys": {},
                    "log0": len(rng.log), "oc": bool(self.oc_member),
                    "fails0": (model.crime_size_fails, model.facilitator_crimes, model.facilitator_fails)})
        cur["active"] = True
        out = orig_find(self, n)
        cur["active"] = False
        cur["group"] = [a.unique_id for a in out]
        cur["fac_pick"] = -1
        for e in rng.log[cur["log0"]:]:
            if e["op"] == "choice" and e.get("size") is None:
                cur["fac_idx"] = int(e["idx"][0])
                cur["fac_npop"] = int(e["n_pop"])
        cur["d_fails"] = (model.crime_size_fails - cur["fails0"][0],
                          model.facilitator_crimes - cur["fails0"][1],
                          model.facilitator_fails - cur["fails0"][2])
        rec["crimes"].append(copy.deepcopy({k: v for k, v in cur.items() if k not in ("active",)}))
        return out

    def agents_at_distance(self, d, *a, **kw):
        out = orig_atd(self, d, *a, **kw)
        if cur.get("active") and self.unique_id == cur.get("initiator"):
            cur["rings"].append((int(d), sorted(x.unique_id for x in out)))
        return out

    def candidates_weight(self, agent):
        w = orig_cw(self, agent)
        if cur.get("active"):
            cur["keys"][agent.unique_id] = float(w)
        return w

    def oc_embeddedness(self):
        miss = self.cached_oc_embeddedness is None
        if miss:
            fresh_flag["on"] = True
            try:
                fresh = fresh_embeddedness(self, nx, orig_emb)
            finally:
                fresh_flag["on"] = False
        v = orig_emb(self)
        if miss:
            rec["emb"].append((self.unique_id, float(v), fresh))
        return v

    def get_caught(self):
        rec["arrested"].append(self.unique_id)
        return orig_caught(self)

    Person.find_accomplices = find_accomplices
    Person.oc_embeddedness = oc_embeddedness
    Person.agents_at_distance = agents_at_distance
    Person.candidates_weight = candidates_weight
    Person.get_caught = get_caught
    model.update_meta_links = update_meta_links
    try:
        model.reset_oc_embeddedness()
        model.commit_crimes()
    finally:
        Person.find_accomplices = orig_find
        Person.oc_embeddedness = orig_emb
        Person.agents_at_distance = orig_atd
        Person.candidates_weight = orig_cw
        Person.get_caught = orig_caught
        del model.update_meta_links
        model.random = real_rng
    rec["rng_log"] = rng.log
    return rec


def step_with_recording(model) -> dict:
    """One reference `step()` (model.py:226-288) whose hot path is recorded.
    Returns {"pre", "pre_scalars", "rec", "post", "post_scalars"} where pre/post
    are `export_state` snapshots taken immediately around
    reset_oc_embeddedness + commit_crimes."""
    holder: dict = {}
    orig_reset, orig_commit = model.reset_oc_embeddedness, model.commit_crimes

    def _noop():
        return None

    def _commit():
        model.reset_oc_embeddedness = orig_reset
        model.commit_crimes = orig_commit
        holder["pre"] = export_state(model)
        holder["pre_scalars"] = model_scalars(model)
        holder["n_alive"] = len(model.schedule.agents)
        holder["intervention_on"] = bool(model.intervention_is_on())
        holder["rec"] = record_commit_crimes(model)
        holder["rng_state_after"] = copy.deepcopy(model.random.bit_generator.state)
        holder["post"] = export_state(model)
        holder["post_scalars"] = model_scalars(model)

    model.reset_oc_embeddedness = _noop
    model.commit_crimes = _commit
    try:
        model.step()
    finally:
        model.reset_oc_embeddedness = orig_reset
        model.commit_crimes = orig_commit
    holder["end"] = export_state(model)          # the state calculate_fast_reporters (model.py:286) has just summarised
    holder["end_reporters"] = fast_reporters(model)
    return holder


FAST_REPORTERS = ["number_crimes_committed_of_persons", "current_oc_members", "current_num_persons",
                  "criminal_tendency_mean", "criminal_tencency_sd", "age_mean", "age_sd", "education_level_mean",
                  "education_level_sd", "num_crime_committed_mean", "num_crime_committed_sd",
                  "crimes_committed_by_oc_this_tick", "current_prisoners", "employed", "facilitators",
                  "tot_friendship_link", "tot_household_link", "tot_partner_link", "tot_offspring_link",
                  "tot_criminal_link", "tot_school_link", "tot_professional_link", "tot_sibling_link",
                  "tot_parent_link", "number_students", "number_crimes", "crime_size_fails", "facilitator_crimes",
                  "facilitator_fails", "big_crime_from_small_fish", "number_offspring_recruited_this_tick",
                  "number_protected_recruited_this_tick", "people_jailed", "number_law_interventions_this_tick",
                  "crime_multiplier", "tick"]


def fast_reporters(model) -> Dict[str, float]:
    """The values ProtonOC.calculate_fast_reporters (model.py:1687-1735) and the hot-path counters hold right now."""
    return {k: float(getattr(model, k)) for k in FAST_REPORTERS}


def replay_from_record(rec: dict, slot_of: Dict[int, int]) -> dict:
    """Turn the RNG log of one recorded commit_crimes into the replay arrays the
    C-ABI / oracle take.  `slot_of` maps unique_id -> slot."""
    log = rec["rng_log"]
    n = len(rec["crimes"])
    u_init, u_size, fac = np.zeros(n), np.zeros(n), np.full(n, -1, dtype=np.int32)
    i = 0
    for c in range(n):
        u_init[c] = log[i]["u"][0]; i += 1
        u_size[c] = log[i]["u"][0]; i += 1
        if i < len(log) and log[i]["op"] == "choice" and log[i].get("size") is None:
            fac[c] = slot_of[log[i]["res_id"]]; i += 1
    out = {"u_init": u_init, "u_size": u_size, "fac_pick": fac, "u_arrest_count": 0.0,
           "u_arrest": np.zeros(0), "u_sentence": np.zeros(0), "arrest_idx": np.zeros(0, dtype=np.int32)}
    if i < len(log):
        assert log[i]["op"] == "random"
        out["u_arrest_count"] = float(log[i]["u"][0]); i += 1
        out["u_arrest"] = np.asarray(log[i]["u"], dtype=np.float64)
        out["arrest_idx"] = np.asarray(log[i]["idx"], dtype=np.int32); i += 1
        sent = []
        while i < len(log):
            sent.append(log[i]["u"][0]); i += 1
        out["u_sentence"] = np.array(sent, dtype=np.float64)
    return out
