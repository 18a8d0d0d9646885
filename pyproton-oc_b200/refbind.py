"""Reference-side binding: run the crime / recruitment hot path of an UNMODIFIED LABSS/PyPROTON-OC model on the B200.

The reference has no plugin layer; its boundary for this path is four methods of `ProtonOC`
(`/root/reference/protonoc/simulator/model.py`):

    calculate_criminal_tendency   model.py:1160-1186      -> poc_tendency
    calculate_crime_multiplier    model.py:1144-1157      -> poc_crime_multiplier
    reset_oc_embeddedness         model.py:709-718        -> (the per-tick cache is a device epoch)
    commit_crimes                 model.py:1417-1485      -> poc_commit_crimes

`patch_reference(ProtonOC)` replaces them on the reference class (or `bind(model)` on one instance); everything else of
the reference -- setup, demography, labour market, interventions, data collection -- keeps running as it is.  Per call
the `Person` object graph is flattened into the structure-of-arrays + nine directed CSR layers the C ABI uploads
(`export_state`), the device runs the path, and `import_state` writes back exactly what the reference's own code would
have changed: criminal links and co-offence counts (model.py:1487-1504, entities.py:215-224), crime counters, OC
membership (model.py:1452-1460), arrests with `get_caught`'s bookkeeping (entities.py:582-603) and the model counters.

Executors.  The binding talks to an executor object with two methods, `run_tendency` and `run_commit`;
`EngineExecutor` (the default) drives `CrimeEngine`, i.e. libprotonoc_b200.so -- there is no CPU fallback.  The CPU test
suite passes its own oracle-backed stand-in to exercise export / import without a GPU.

Random numbers.  Production draws come from Philox keyed by (model seed, tick), so a patched model does not consume
`model.random` inside commit_crimes: whole runs agree with the reference in distribution, not draw for draw.  For
bit-exact checks `binding.replay` takes the uniforms recorded from a reference run (include/protonoc_b200.h poc_replay)
and `binding.rng_state_after` the generator state the reference was left in.
"""
from __future__ import annotations

import types
from typing import Dict, List, Optional, Sequence

import numpy as np

LAYERS = ["sibling", "offspring", "parent", "partner", "household", "friendship", "criminal", "professional", "school"]
PARAM_NAMES = ["nat_propensity_m", "nat_propensity_sigma", "nat_propensity_threshold", "number_crimes_yearly_per10k",
               "number_arrests_per_year", "punishment_length", "threshold_use_facilitators",
               "facilitator_repression_multiplier", "good_guy_threshold", "max_accomplice_radius",
               "oc_embeddedness_radius", "facilitator_repression", "oc_boss_repression", "ticks_between_intervention",
               "intervention_start", "intervention_end", "this_is_a_big_crime", "ticks_per_year"]


# ------------------------------------------------------------------------------------------------ export
def _closure(model) -> list:
    """Scheduled agents plus every Person still reachable through a neighbour set or a `father` pointer: dead agents
    stay referenced where a link was cleared one-sidedly (entities.py:601-602, 652-661) and the reference keeps
    traversing them, so they get slots (alive = 0)."""
    sched = model.schedule.agents
    seen = {a.unique_id: a for a in sched}
    stack = list(sched)
    while stack:
        a = stack.pop()
        for net in LAYERS:
            for b in a.neighbors[net]:
                if b.unique_id not in seen:
                    seen[b.unique_id] = b
                    stack.append(b)
        f = a.father
        if f is not None and f.unique_id not in seen:
            seen[f.unique_id] = f
            stack.append(f)
    return [seen[k] for k in sorted(seen)]


def export_state(model):
    """(state, agents): SoA columns + 9 directed CSR layers (layer order = Person.network_names, entities.py:38-47) and
    the Person behind every slot.  Slot order is ascending unique_id = the Mesa schedule order in which the reference
    walks a cell (model.py:1430-1432)."""
    agents = _closure(model)
    n = len(agents)
    alive_ids = {a.unique_id for a in model.schedule.agents}
    slot = {a.unique_id: i for i, a in enumerate(agents)}
    col = lambda f, dt: np.fromiter((f(a) for a in agents), dtype=dt, count=n)  # noqa: E731
    st: Dict[str, np.ndarray] = {
        "unique_id": col(lambda a: a.unique_id, np.int64),
        "alive": col(lambda a: a.unique_id in alive_ids, np.uint8),
        "age": col(lambda a: int(a.age), np.int32),
        "birth_tick": col(lambda a: int(a.birth_tick), np.int32),
        "male": col(lambda a: bool(a.gender_is_male), np.uint8),
        "job_level": col(lambda a: int(a.job_level), np.int8),
        "education_level": col(lambda a: int(a.education_level), np.int8),
        "wealth_level": col(lambda a: int(a.wealth_level), np.int8),
        "propensity": col(lambda a: float(a.propensity), np.float64),
        "oc_member": col(lambda a: bool(a.oc_member), np.uint8),
        "facilitator": col(lambda a: bool(a.facilitator), np.uint8),
        "prisoner": col(lambda a: bool(a.prisoner), np.uint8),
        "retired": col(lambda a: bool(a.retired), np.uint8),
        "has_job": col(lambda a: a.my_job is not None, np.uint8),
        "has_school": col(lambda a: a.my_school is not None, np.uint8),
        "max_education_level": col(lambda a: int(a.max_education_level), np.int8),
        "school_level": col(lambda a: int(a.my_school.diploma_level) if a.my_school is not None else 0, np.int8),
        "target_of_intervention": col(lambda a: bool(a.target_of_intervention), np.uint8),
        "num_crimes_committed": col(lambda a: int(a.num_crimes_committed), np.int32),
        "num_crimes_committed_this_tick": col(lambda a: int(a.num_crimes_committed_this_tick), np.int32),
        "criminal_tendency": col(lambda a: float(a.criminal_tendency), np.float64),
        "sentence_countdown": col(lambda a: float(a.sentence_countdown), np.float64),
        "new_recruit": col(lambda a: int(a.new_recruit), np.int32),
        "arrest_weight": col(lambda a: float(a.arrest_weight), np.float64),
        "father": col(lambda a: slot[a.father.unique_id] if a.father is not None else -1, np.int32),
        "number_of_children": col(lambda a: int(a.number_of_children), np.int32),
    }
    for li, net in enumerate(LAYERS):
        rowptr = np.zeros(n + 1, dtype=np.int32)
        cols: List[int] = []
        cnt: List[int] = []
        for i, a in enumerate(agents):
            nb = a.neighbors[net]
            if nb:
                if net == "criminal":
                    # a criminal neighbour without a num_co_offenses entry counts 0 (model.py:1531-1533)
                    row = sorted((slot[b.unique_id], int(a.num_co_offenses.get(b, 0))) for b in nb)
                    cols.extend(r[0] for r in row)
                    cnt.extend(r[1] for r in row)
                else:
                    cols.extend(sorted(slot[b.unique_id] for b in nb))
            rowptr[i + 1] = len(cols)
        st["rowptr_%d" % li] = rowptr
        st["col_%d" % li] = np.asarray(cols, dtype=np.int32)
        if net == "criminal":
            st["co_off"] = np.asarray(cnt, dtype=np.int32)
    return st, agents


def model_tables(model):
    """The three tables of the path in the layout make_params takes: c_range_by_age_and_sex (model.py:177-178),
    num_co_offenders_dist (:188-190), male / female_punishment_length (:170-174)."""
    c_range = [(int(bool(k[0])), int(k[1]), int(v[0]), float(v[1])) for k, v in model.c_range_by_age_and_sex]
    co = [(int(r[0]), float(r[1])) for r in model.num_co_offenders_dist]
    conv = [(float(f[0]), float(f[1]), float(m[1])) for f, m in zip(model.female_punishment_length, model.male_punishment_length)]
    return c_range, co, conv


def model_params(model) -> dict:
    return {k: getattr(model, k) for k in PARAM_NAMES}


# ------------------------------------------------------------------------------------------------ import
def import_state(model, agents: Sequence, before: Dict[str, np.ndarray], after: Dict[str, np.ndarray], out: dict,
                 rec: dict) -> None:
    """Write one device-side commit_crimes back into the Person graph: what model.py:1450-1485 mutates, nothing else."""
    members = np.unique(np.concatenate(rec["groups"])) if len(rec["groups"]) else np.zeros(0, np.int64)
    for i in members.tolist():
        a = agents[i]
        a.num_crimes_committed = int(after["num_crimes_committed"][i])
        a.num_crimes_committed_this_tick = int(after["num_crimes_committed_this_tick"][i])
        a.arrest_weight = float(after["arrest_weight"][i])
        if after["oc_member"][i] and not before["oc_member"][i]:
            a.oc_member = True
            a.new_recruit = int(after["new_recruit"][i])
    # criminal rows of the members of multi-member groups (commit_crime, model.py:1487-1504): links are symmetric,
    # every other row is untouched
    rp, cl, co = after["rowptr_6"], after["col_6"], after["co_off"]
    touched = {int(m) for g in rec["groups"] if len(g) > 1 for m in g}
    for i in touched:
        a = agents[i]
        nb = a.neighbors["criminal"]
        for e in range(rp[i], rp[i + 1]):
            b = agents[int(cl[e])]
            nb.add(b)
            a.num_co_offenses[b] = int(co[e])
    # arrests: Person.get_caught (entities.py:582-603) minus its sentence draw, which the device already made
    for i in rec["arrested"].tolist():
        a = agents[int(i)]
        model.number_law_interventions_this_tick += 1
        model.people_jailed += 1
        a.prisoner = True
        a.sentence_countdown = float(after["sentence_countdown"][i])
        if a.my_job:
            a.my_job.my_worker = None
            a.my_job = None
            a.job_level = 1
        if a.my_school:
            a.leave_school()
        a.neighbors.get("professional").clear()
        a.neighbors.get("school").clear()
    model.number_crimes += int(out["d_number_crimes"])
    model.crime_size_fails += int(out["d_crime_size_fails"])
    model.facilitator_crimes += int(out["d_facilitator_crimes"])
    model.facilitator_fails += int(out["d_facilitator_fails"])
    model.big_crime_from_small_fish += int(out["d_big_crime_from_small_fish"])
    model.number_offspring_recruited_this_tick += int(out["d_offspring_recruited"])
    model.number_protected_recruited_this_tick += int(out["d_protected_recruited"])


# ------------------------------------------------------------------------------------------------ executors
class EngineExecutor:
    """The B200 executor: one CrimeEngine (libprotonoc_b200.so), re-created when the population outgrows it."""

    def __init__(self, device: int = 0, headroom: float = 1.5):
        self.device, self.headroom, self.E = device, headroom, None

    def _engine(self, st, params, tables):
        from .engine import CrimeEngine
        from .params import make_params
        n, stat, crim = CrimeEngine.capacity_for([st], 16 * len(st["alive"]))
        E = self.E
        if E is None or n > E.max_agents or stat > E.max_static or crim > E.max_criminal:
            if E is not None:
                E.close()
            h = self.headroom
            E = self.E = CrimeEngine(1, int(n * h) + 64, int(stat * h) + 1024, int(crim * h) + 1024, device=self.device)
        c_range, co, conv = tables
        E.set_params(make_params(params, c_range=c_range, co_offenders=co, conviction=conv))
        E.upload_state(0, st)
        return E

    def run_tendency(self, st, params, tables):
        E = self._engine(st, params, tables)
        E.tendency()
        mult = float(E.crime_multiplier()[0])
        return E.download_agents(0)["criminal_tendency"], mult

    def run_multiplier(self, st, params, tables):
        return float(self._engine(st, params, tables).crime_multiplier()[0])

    def run_commit(self, st, params, tables, tick, mult, replay, key):
        E = self._engine(st, params, tables)
        E.set_crime_multiplier(mult)
        outs, rec = E.commit_crimes(tick, [replay] if replay is not None else None, philox_keys=[key], records=True)
        return outs[0], rec, E.download_state(0)

    def close(self):
        if self.E is not None:
            self.E.close()
            self.E = None


class HotPathBinding:
    def __init__(self, model, executor=None):
        self.model = model
        self.executor = executor if executor is not None else EngineExecutor()
        self.replay: Optional[dict] = None          # recorded uniforms for the next commit_crimes (bit-exact checks)
        self.rng_state_after: Optional[dict] = None  # generator state to leave model.random in after it
        self.last = None                            # (out, rec) of the last commit_crimes

    def _key(self) -> int:
        return (int(self.model.seed) * 1000003 + 12345) & 0xFFFFFFFFFFFFFFFF

    def calculate_criminal_tendency(self):
        m = self.model
        st, agents = export_state(m)
        tend, _ = self.executor.run_tendency(st, model_params(m), model_tables(m))
        alive = st["alive"].astype(bool)
        for a, t, al in zip(agents, tend.tolist(), alive.tolist()):
            if al:
                a.criminal_tendency = t

    def calculate_crime_multiplier(self):
        m = self.model
        st, _ = export_state(m)
        m.crime_multiplier = self.executor.run_multiplier(st, model_params(m), model_tables(m))

    def commit_crimes(self):
        m = self.model
        st, agents = export_state(m)
        out, rec, after = self.executor.run_commit(st, model_params(m), model_tables(m), int(m.tick),
                                                   float(m.crime_multiplier), self.replay, self._key())
        if out["error"]:
            raise RuntimeError("crime path reported error code %d (a state the reference itself fails on)" % out["error"])
        import_state(m, agents, st, after, out, rec)
        self.last = (out, rec)
        self.replay = None
        if self.rng_state_after is not None:
            m.random.bit_generator.state = self.rng_state_after
            self.rng_state_after = None


def bind(model, executor=None) -> HotPathBinding:
    """Route the four hot-path methods of ONE reference model instance to the executor (default: the GPU)."""
    b = HotPathBinding(model, executor)
    model._poc_binding = b
    model.calculate_criminal_tendency = types.MethodType(lambda self: self._poc_binding.calculate_criminal_tendency(), model)
    model.calculate_crime_multiplier = types.MethodType(lambda self: self._poc_binding.calculate_crime_multiplier(), model)
    model.commit_crimes = types.MethodType(lambda self: self._poc_binding.commit_crimes(), model)
    return b


def patch_reference(ProtonOC, executor_factory=None):
    """Replace the hot-path methods on the reference CLASS: every model created afterwards runs them on the device.
    Returns a function that restores the original methods."""
    saved = {k: getattr(ProtonOC, k) for k in ("calculate_criminal_tendency", "calculate_crime_multiplier", "commit_crimes")}

    def _binding(self) -> HotPathBinding:
        b = self.__dict__.get("_poc_binding")
        if b is None:
            b = HotPathBinding(self, executor_factory() if executor_factory else None)
            self.__dict__["_poc_binding"] = b
        return b

    ProtonOC.calculate_criminal_tendency = lambda self: _binding(self).calculate_criminal_tendency()
    ProtonOC.calculate_crime_multiplier = lambda self: _binding(self).calculate_crime_multiplier()
    ProtonOC.commit_crimes = lambda self: _binding(self).commit_crimes()

    def restore():
        for k, v in saved.items():
            setattr(ProtonOC, k, v)
    return restore
