"""Drop-in `ProtonOC` for the crime / recruitment hot path (reference: protonoc/simulator/model.py).

Same constructor, attributes, parameter-override methods and run()/setup()/step()/save_data() signatures as
the reference class (model.py:44-312, 1604-1763) and the same exception types; `step()` routes the ★ rows of
SURVEY.md §8a -- calculate_criminal_tendency, calculate_crime_multiplier, reset_oc_embeddedness and
commit_crimes -- to the CUDA engine.  `Person` objects are thin handles over the structure-of-arrays state
that keep the per-agent methods the reference's tests call (agents_at_distance, agents_in_radius,
social_proximity, oc_embeddedness, find_accomplices, the link makers), each a batch-of-one call of the
corresponding stage hook of the C ABI.

Out of scope for this round: the reference's population synthesis and the
demographic / friendship / labour-market phases of step().  setup() therefore builds its population with
`synth.synthetic_state` (or takes a state exported from the reference via `load_state`), and step() advances
only the hot path plus the bookkeeping it depends on (ageing, this-tick counters, prisoner countdown).
`ProtonOC.out_of_scope_phases` lists what step() does not do.
"""
from __future__ import annotations

import json
import os
import pickle
import sys
from typing import Any, Dict, Iterable, List, Optional, Set, Union
from xml.dom import minidom

import numpy as np
import pandas as pd

from . import synth, tables
from .engine import LAYER_NAMES, REPORTER_NAMES, CrimeEngine
from .params import make_demography, make_labour, make_params

free_parameters = ["migration_on", "initial_agents", "num_ticks", "intervention",
                   "max_accomplice_radius", "number_arrests_per_year",
                   "number_crimes_yearly_per10k", "ticks_between_intervention",
                   "intervention_start", "intervention_end", "num_oc_persons",
                   "num_oc_families", "education_modifier", "retirement_age",
                   "unemployment_multiplier", "nat_propensity_m",
                   "nat_propensity_sigma", "nat_propensity_threshold",
                   "facilitator_repression", "facilitator_repression_multiplier",
                   "likelihood_of_facilitators", "targets_addressed_percent",
                   "threshold_use_facilitators", "oc_embeddedness_radius",
                   "oc_boss_repression", "punishment_length", "constant_population"]

# the DataCollector columns of the reference (extra.py:246-274), in order
MODEL_REPORTERS = ["seed", "family_intervention", "social_support", "welfare_support",
                   "this_is_a_big_crime", "good_guy_threshold", "number_deceased",
                   "facilitator_fails", "facilitator_crimes", "crime_size_fails",
                   "number_born", "number_migrants", "number_weddings",
                   "number_weddings_mean", "number_law_interventions_this_tick",
                   "correction_for_non_facilitators", "number_protected_recruited_this_tick",
                   "people_jailed", "number_offspring_recruited_this_tick", "number_crimes",
                   "crime_multiplier", "kids_intervention_counter", "big_crime_from_small_fish",
                   "arrest_rate", "migration_on", "initial_agents", "intervention",
                   "max_accomplice_radius", "number_arrests_per_year", "ticks_per_year",
                   "num_ticks", "tick", "ticks_between_intervention", "intervention_start",
                   "intervention_end", "num_oc_persons", "num_oc_families",
                   "education_modifier", "retirement_age", "unemployment_multiplier",
                   "nat_propensity_m", "nat_propensity_sigma", "nat_propensity_threshold",
                   "facilitator_repression", "facilitator_repression_multiplier",
                   "percentage_of_facilitators", "targets_addressed_percent",
                   "threshold_use_facilitators", "oc_embeddedness_radius",
                   "oc_boss_repression", "punishment_length",
                   "constant_population", "number_crimes_yearly_per10k",
                   "number_crimes_committed_of_persons", "current_oc_members",
                   "current_num_persons", "criminal_tendency_mean", "criminal_tencency_sd",
                   "age_mean", "age_sd", "education_level_mean", "education_level_sd",
                   "num_crime_committed_mean", "num_crime_committed_sd",
                   "crimes_committed_by_oc_this_tick", "current_prisoners", "employed",
                   "facilitators", "tot_friendship_link", "tot_household_link",
                   "tot_partner_link", "tot_offspring_link", "tot_criminal_link",
                   "tot_school_link", "tot_professional_link", "tot_sibling_link",
                   "tot_parent_link", "number_students", "number_jobs",
                   "likelihood_of_facilitators"]

# the nine intervention presets, only the fields the hot path reads (model.py:1244-1348)
_PRESETS = {
    "baseline": dict(oc_boss_repression=False, facilitator_repression=False, ticks_between_intervention=12),
    "preventive": dict(oc_boss_repression=False, facilitator_repression=False, ticks_between_intervention=1),
    "preventive-strong": dict(oc_boss_repression=False, facilitator_repression=False, ticks_between_intervention=1),
    "disruptive": dict(oc_boss_repression=False, facilitator_repression=False, ticks_between_intervention=1),
    "disruptive-strong": dict(oc_boss_repression=True, facilitator_repression=False, ticks_between_intervention=1),
    "students": dict(oc_boss_repression=False, facilitator_repression=False, ticks_between_intervention=12),
    "students-strong": dict(oc_boss_repression=False, facilitator_repression=False, ticks_between_intervention=1),
    "facilitators": dict(oc_boss_repression=False, facilitator_repression=True, ticks_between_intervention=1,
                         facilitator_repression_multiplier=3),
    "facilitators-strong": dict(oc_boss_repression=False, facilitator_repression=True, ticks_between_intervention=1,
                                facilitator_repression_multiplier=20),
}

_COLS = {"alive": np.uint8, "age": np.int32, "birth_tick": np.int32, "male": np.uint8, "job_level": np.int8,
         "education_level": np.int8, "wealth_level": np.int8, "propensity": np.float64, "oc_member": np.uint8,
         "facilitator": np.uint8, "prisoner": np.uint8, "retired": np.uint8, "has_job": np.uint8,
         "has_school": np.uint8, "target_of_intervention": np.uint8, "num_crimes_committed": np.int32,
         "num_crimes_committed_this_tick": np.int32, "criminal_tendency": np.float64,
         "sentence_countdown": np.float64, "new_recruit": np.int32, "arrest_weight": np.float64, "father": np.int32,
         "number_of_children": np.int32,
         # graduate_and_enter_jobmarket (model.py:1351-1380): Person.max_education_level, diploma level of Person.my_school
         "max_education_level": np.int8, "school_level": np.int8}
_COL_DEFAULTS = {"alive": 1, "wealth_level": 1, "new_recruit": -2, "father": -1}
# Person attribute -> state column
_ATTR = {"gender_is_male": "male", "num_crimes_committed": "num_crimes_committed",
         "num_crimes_committed_this_tick": "num_crimes_committed_this_tick"}


class _CoOffenses:
    """dict-like view of Person.num_co_offenses (entities.py:83) keyed by Person."""

    def __init__(self, model, slot):
        self._m, self._s = model, slot

    def __getitem__(self, other):
        self._m._pull()
        return self._m._co[(self._s, other.slot)]

    def __setitem__(self, other, value):
        self._m._pull()
        self._m._co[(self._s, other.slot)] = int(value)
        self._m._dirty = True

    def __contains__(self, other):
        self._m._pull()
        return (self._s, other.slot) in self._m._co


class Person:
    """Handle on one agent slot.  `Person(model)` appends a new agent like the reference constructor does."""
    network_names: List[str] = list(LAYER_NAMES)

    def __init__(self, model: "ProtonOC", _slot: Optional[int] = None):
        object.__setattr__(self, "model", model)
        if _slot is None:
            _slot = model._append_agent()
        object.__setattr__(self, "slot", _slot)
        object.__setattr__(self, "unique_id", _slot)

    # -- identity
    def __hash__(self):
        return self.slot

    def __eq__(self, other):
        return isinstance(other, Person) and other.slot == self.slot and other.model is self.model

    def __repr__(self):
        return "Agent: " + str(self.unique_id)

    # -- attributes live in the SoA state
    def __getattr__(self, name):
        m = object.__getattribute__(self, "model")
        col = _ATTR.get(name, name)
        if col in _COLS:
            m._pull()
            v = m._st[col][self.slot]
            if _COLS[col] is np.uint8:
                return bool(v)
            return v.item()
        if name == "num_co_offenses":
            return _CoOffenses(m, self.slot)
        if name == "neighbors":
            m._pull()
            return {l: {Person(m, s) for s in m._adj[li][self.slot]} for li, l in enumerate(LAYER_NAMES)}
        if name == "father":
            m._pull()
            f = int(m._st["father"][self.slot])
            return Person(m, f) if f >= 0 else None
        raise AttributeError(name)

    def __setattr__(self, name, value):
        m = self.model
        col = _ATTR.get(name, name)
        if col == "father":
            value = value.slot if isinstance(value, Person) else -1
        if col in _COLS:
            m._pull()
            m._st[col][self.slot] = value
            m._dirty = True
            return
        object.__setattr__(self, name, value)

    # -- link makers (entities.py:133-224)
    def _link(self, layer, other, both=True):
        m = self.model
        m._pull()
        li = LAYER_NAMES.index(layer)
        others = other if isinstance(other, list) else [other]
        for o in others:
            if o.slot == self.slot:
                continue
            m._adj[li][self.slot].add(o.slot)
            if both:
                m._adj[li][o.slot].add(self.slot)
        m._dirty = True

    def make_friendship_link(self, asker): self._link("friendship", asker)
    def make_professional_link(self, asker): self._link("professional", asker)
    def add_sibling_link(self, targets): self._link("sibling", targets)
    def make_household_link(self, targets): self._link("household", targets)
    def make_partner_link(self, asker): self._link("partner", asker)
    def make_school_link(self, asker): self._link("school", asker)

    def make_parent_offsprings_link(self, asker):
        for o in (asker if isinstance(asker, list) else [asker]):
            self._link("offspring", o, both=False)
            o._link("parent", self, both=False)

    def add_criminal_link(self, asker):
        self._link("criminal", asker)
        self.model._co[(self.slot, asker.slot)] = 0
        self.model._co[(asker.slot, self.slot)] = 0

    # -- the hot-path methods, each a batch-of-one device call
    def _wrap(self, slots: Iterable[int]) -> Set["Person"]:
        return {Person(self.model, int(s)) for s in slots}

    def agents_at_distance(self, d: int) -> Set["Person"]:
        """entities.py:506-524"""
        return self._wrap(self.model._engine_ready().rings([self.slot], d, True)[0])

    def agents_in_radius(self, d: int) -> Set["Person"]:
        """entities.py:482-504"""
        return self._wrap(self.model._engine_ready().rings([self.slot], d, False)[0])

    def social_proximity(self, target: "Person") -> float:
        """entities.py:674-689"""
        return float(self.model._engine_ready().social_proximity([self.slot], [target.slot])[0])

    def oc_embeddedness(self) -> float:
        """entities.py:526-543"""
        return float(self.model._engine_ready().embeddedness([self.slot])[0])

    def candidates_weight(self, agent: "Person") -> float:
        """entities.py:457-467"""
        return float(self.model._engine_ready().candidate_weights([self.slot], [agent.slot])[0])

    def find_accomplices(self, n_of_accomplices: int) -> List["Person"]:
        """entities.py:413-455 (counters crime_size_fails / facilitator_crimes / facilitator_fails updated)"""
        m = self.model
        u = float(m.random.random())
        groups, fails = m._engine_ready().find_accomplices([self.slot], [int(n_of_accomplices)], u_fac=[u])
        m.crime_size_fails += int(fails[0][0]); m.facilitator_crimes += int(fails[0][1]); m.facilitator_fails += int(fails[0][2])
        return [Person(m, int(s)) for s in groups[0]]

    def isneighbor(self, other: "Person") -> bool:
        self.model._pull()
        return any(other.slot in self.model._adj[li][self.slot] for li in range(len(LAYER_NAMES)))


class _Schedule:
    """Stand-in for the Mesa BaseScheduler the reference uses as an ordered agent container (model.py:74)."""

    def __init__(self, model):
        self._m = model
        self.steps = 0

    @property
    def agents(self) -> List[Person]:
        self._m._pull()
        return [Person(self._m, int(s)) for s in np.flatnonzero(self._m._st["alive"])]

    def add(self, agent: Person):
        agent.alive = True

    def remove(self, agent: Person):
        agent.alive = False

    def step(self):
        self.steps += 1


class _DataCollector:
    """model_reporters by attribute name, like mesa.datacollection.DataCollector as the reference uses it."""

    def __init__(self, names):
        self.model_vars = {k: [] for k in names}

    def collect(self, model):
        for k in self.model_vars:
            self.model_vars[k].append(getattr(model, k, None))

    def get_model_vars_dataframe(self):
        return pd.DataFrame(self.model_vars)


class ProtonOC:
    out_of_scope_phases = ["interventions (family / social / welfare)", "find_job", "let_migrants_in", "return_kids"]
    # the phases of step() besides the crime path that run on the device (CrimeEngine.PHASE_*): on the yearly
    # ticks graduate_and_enter_jobmarket and the update_unemployment_status loop (PHASE_LABOUR, model.py:251-256); every tick
    # wedding, retire_persons, make_baby, remove_excess_friends, remove_excess_professional_links, make_friends,
    # make_people_die, in the reference's order (model.py:271-284).  0 = the crime path alone.
    device_phases: int = CrimeEngine.PHASE_EVERY

    def __init__(self, seed: int = int.from_bytes(os.urandom(4), sys.byteorder), collect_agents: bool = False) -> None:
        self.seed: int = seed
        self.random = np.random.default_rng(seed=seed)
        self.verbose = False
        self.datacollector: Optional[_DataCollector] = None
        self.ticks_per_year: int = 12
        self.tick: int = 1
        self.collect_agents = collect_agents
        self.schedule = _Schedule(self)
        # intervention
        self.family_intervention = None
        self.social_support = None
        self.welfare_support = None
        # outputs (model.py:82-104)
        self.this_is_a_big_crime: int = 3
        self.good_guy_threshold: float = 0.6
        self.number_deceased: float = 0
        self.facilitator_fails: float = 0
        self.facilitator_crimes: float = 0
        self.crime_size_fails: float = 0
        self.number_born: int = 0
        self.number_migrants: int = 0
        self.number_weddings: int = 0
        self.number_weddings_mean = tables.NUMBER_WEDDINGS_MEAN   # marriages_stats.csv (model.py:184-187)
        self.number_weddings_sd: float = tables.NUMBER_WEDDINGS_SD
        self.number_law_interventions_this_tick: int = 0
        self.correction_for_non_facilitators: float = 0
        self.number_protected_recruited_this_tick: int = 0
        self.number_offspring_recruited_this_tick: int = 0
        self.people_jailed: int = 0
        self.number_crimes: int = 0
        self.crime_multiplier: float = 0
        self.kids_intervention_counter: int = 0
        self.big_crime_from_small_fish: int = 0
        self.arrest_rate: Union[int, float] = 0
        self.city = None
        # free parameters (model.py:108-134)
        self.migration_on: bool = True
        self.initial_agents: int = 1000
        self.num_ticks: int = 480
        self.intervention: str = "baseline"
        self.max_accomplice_radius: int = 2
        self.number_arrests_per_year: int = 30
        self.number_crimes_yearly_per10k: int = 2000
        self.ticks_between_intervention: int = 12
        self.intervention_start: int = 13
        self.intervention_end: int = 999
        self.num_oc_persons: int = 30
        self.num_oc_families: int = 8
        self.education_modifier: float = 1.0
        self.retirement_age: int = 65
        self.unemployment_multiplier: Union[str, int] = "base"
        self.nat_propensity_m: float = 1.0
        self.nat_propensity_sigma: float = 0.25
        self.nat_propensity_threshold: float = 1.0
        self.facilitator_repression: bool = False
        self.facilitator_repression_multiplier: float = 2.0
        self.likelihood_of_facilitators: float = 0.005
        self.targets_addressed_percent: float = 10
        self.threshold_use_facilitators: float = 4
        self.oc_embeddedness_radius: int = 2
        self.oc_boss_repression: bool = False
        self.punishment_length: int = 1
        self.constant_population: bool = False
        # tables on the hot path (model.py:170-190)
        self.c_range_by_age_and_sex = [[[bool(m), a0], [a1, c]] for m, a0, a1, c in tables.C_RANGE_BY_AGE_AND_SEX]
        self.num_co_offenders_dist = [list(r) for r in tables.NUM_CO_OFFENDERS_DIST]
        self.male_punishment_length = [[r[0], r[2]] for r in tables.CONVICTION_LENGTH]
        self.female_punishment_length = [[r[0], r[1]] for r in tables.CONVICTION_LENGTH]
        # tables of the yearly labour-status block (model.py:166-169, 180-183; education_levels: enter and escape age per level,
        # model.py:953-965 without the school counts)
        self.education_levels: Dict = {k: [v[0], v[1]] for k, v in tables.EDUCATION_LEVELS.items()}
        self.work_status_by_edu_lvl = {k: {m: [list(r) for r in rows] for m, rows in d.items()}
                                       for k, d in tables.WORK_STATUS_BY_EDU_LVL.items()}
        self.wealth_quintile_by_work_status = {k: {m: [list(r) for r in rows] for m, rows in d.items()}
                                               for k, d in tables.WEALTH_QUINTILE_BY_WORK_STATUS.items()}
        self.labour_status_by_age_and_sex = {m: dict(d) for m, d in tables.LABOUR_STATUS_BY_AGE_AND_SEX.items()}
        self.labour_status_range = {m: dict(d) for m, d in tables.LABOUR_STATUS_RANGE.items()}
        # ---- everything below is implementation state (not model attributes)
        self._st: Dict[str, np.ndarray] = {k: np.zeros(0, dt) for k, dt in _COLS.items()}
        self._adj: List[List[set]] = [[] for _ in LAYER_NAMES]
        self._co: Dict[tuple, int] = {}
        self._engine: Optional[CrimeEngine] = None
        self._dirty = True         # host state changed since the last upload
        self._host_stale = False   # device state changed since the last download
        self._device = 0
        self._philox_key = int(seed) & 0xFFFFFFFFFFFFFFFF

    def __repr__(self):
        return "PROTON-OC MODEL, seed: {}".format(str(self.seed))

    # ------------------------------------------------------------------ parameter API (model.py:1635-1763)
    def _public(self):
        return {k for k in self.__dict__ if not k.startswith("_")}

    def set_param(self, param_name: str, value: Any) -> None:
        if param_name not in self._public():
            raise AttributeError(param_name + "attribute is not a valid parameter, to find out "
                                              "the valid parameters: ProtonOC.overview().")
        if type(value) != type(getattr(self, param_name)):
            raise ValueError("param " + param_name + " accepts parameters of type " + type(
                getattr(self, param_name)).__name__ + " instead provided " + type(value).__name__)
        setattr(self, param_name, value)

    def override_dict(self, dictionary) -> None:
        for key, value in dictionary.items():
            if key not in self._public() and key != "repetitions":
                raise Exception("{} is not a model attribute".format(key))
            setattr(self, key, value)

    def override_json(self, json_file_path: str) -> None:
        with open(json_file_path) as f:
            self.override_dict(json.load(f))

    def override_xml(self, xml_file: str) -> None:
        for key, value in parse_behaviorspace_xml(xml_file).items():
            if key not in self._public():
                raise Exception("{} is not a model attribute".format(key))
            setattr(self, key, value)

    def overview(self) -> None:
        width = max(len(p) for p in free_parameters)
        print("%-*s | value" % (width, "free parameter name"))
        for par in free_parameters:
            print("%-*s | %s" % (width, par, getattr(self, par)))

    def intervention_is_on(self) -> bool:
        return self.tick % self.ticks_between_intervention == 0 and self.intervention_start <= self.tick < self.intervention_end

    def choose_intervention_setting(self) -> None:
        """model.py:1244-1348, restricted to the attributes the hot path reads."""
        preset = _PRESETS.get(self.intervention)
        if preset is None:
            return
        for k, v in preset.items():
            setattr(self, k, v)
        self.intervention_start, self.intervention_end = 13, 9999

    # ------------------------------------------------------------------ host state
    def _append_agent(self) -> int:
        self._pull()
        slot = len(self._st["alive"])
        for k, dt in _COLS.items():
            self._st[k] = np.append(self._st[k], np.array([_COL_DEFAULTS.get(k, 0)], dtype=dt))
        self._st["birth_tick"][slot] = self.tick
        self._st["propensity"][slot] = np.exp(self.nat_propensity_m + self.nat_propensity_sigma * self.random.normal())
        self._st["male"][slot] = bool(self.random.choice([True, False]))
        for li in range(len(LAYER_NAMES)):
            self._adj[li].append(set())
        self._dirty = True
        return slot

    def load_state(self, st: Dict[str, np.ndarray]) -> None:
        """Adopt an exported state (the format of tests/golden and CrimeEngine.download_state)."""
        n = len(st["alive"])
        self._st = {k: np.array(st[k], dtype=dt) if k in st else np.full(n, _COL_DEFAULTS.get(k, 0), dt)
                    for k, dt in _COLS.items()}
        if "number_of_children" not in st:   # Person.number_of_children (entities.py:75): the offspring links when not exported
            self._st["number_of_children"] = np.diff(st["rowptr_1"]).astype(np.int32)
        self._adj = []
        for li in range(len(LAYER_NAMES)):
            rp, col = st["rowptr_%d" % li], st["col_%d" % li]
            self._adj.append([set(col[rp[a]:rp[a + 1]].tolist()) for a in range(n)])
        rp, col = st["rowptr_6"], st["col_6"]
        self._co = {}
        for a in range(n):
            for e in range(rp[a], rp[a + 1]):
                self._co[(a, int(col[e]))] = int(st["co_off"][e])
        self._dirty, self._host_stale = True, False

    def state_dict(self) -> Dict[str, np.ndarray]:
        self._pull()
        n = len(self._st["alive"])
        st = {k: v.copy() for k, v in self._st.items()}
        st["unique_id"] = np.arange(n, dtype=np.int64)
        for li in range(len(LAYER_NAMES)):
            rowptr = np.zeros(n + 1, np.int32)
            cols: List[int] = []
            for a in range(n):
                cols.extend(sorted(self._adj[li][a]))
                rowptr[a + 1] = len(cols)
            st["rowptr_%d" % li], st["col_%d" % li] = rowptr, np.array(cols, dtype=np.int32)
        rp, col = st["rowptr_6"], st["col_6"]
        st["co_off"] = np.array([self._co.get((a, int(col[e])), 0) for a in range(n) for e in range(rp[a], rp[a + 1])],
                                dtype=np.int32)
        return st

    def _params(self):
        return make_params(self, c_range=[(int(c[0][0]), c[0][1], c[1][0], c[1][1]) for c in self.c_range_by_age_and_sex],
                           co_offenders=self.num_co_offenders_dist,
                           conviction=[(m[0], f[1], m[1]) for m, f in zip(self.male_punishment_length, self.female_punishment_length)])

    def _labour(self):
        """Tables of the yearly labour-status block from the model's own attributes (model.py:166-169, 180-183, 953-965)."""
        return make_labour(education_levels={int(k): (int(v[0]), int(v[1])) for k, v in self.education_levels.items()},
                           work_status=self.work_status_by_edu_lvl, wealth_quintile=self.wealth_quintile_by_work_status,
                           labour_status=self.labour_status_by_age_and_sex,
                           range_ages={k: sorted(v) for k, v in self.labour_status_range.items()})

    def _engine_ready(self) -> CrimeEngine:
        """Upload the host state if it changed; (re)create the engine when capacities are exceeded."""
        if self._dirty or self._engine is None:
            st = self.state_dict()
            n, stat, crim = CrimeEngine.capacity_for([st], headroom_criminal=max(200000, 16 * len(st["alive"])))
            e = self._engine
            if e is None or n > e.max_agents or stat > getattr(e, "_stat_cap", 0) or crim > getattr(e, "_crim_cap", 0):
                if e is not None:
                    e.close()
                # births add agent slots and links on the device: room for them (a run that outgrows it raises, error 20-22)
                grow = (max(n // 2, 2000), max(stat, 1)) if self.device_phases else (0, 0)
                e = CrimeEngine(1, max(n, 1) + grow[0], max(stat, 1) * 2 + grow[1], crim, device=self._device)
                e._stat_cap, e._crim_cap = max(stat, 1) * 2, crim
                self._engine = e
            e.set_params(self._params())
            e.set_demography(make_demography(self.nat_propensity_m, self.nat_propensity_sigma, retirement_age=self.retirement_age,
                                             number_weddings_mean=getattr(self, "number_weddings_mean", None)))
            e.set_labour(self._labour())
            e.upload_state(0, st)
            if "number_of_children" in st:
                e.upload_children(0, st["number_of_children"])
            e.upload_education(0, st["max_education_level"], st["school_level"])
            e.set_crime_multiplier(float(self.crime_multiplier))
            self._dirty = False
        else:
            self._engine.set_params(self._params())
        return self._engine

    def _pull(self) -> None:
        """Refresh the host copy after device steps."""
        if not self._host_stale or self._engine is None:
            return
        self._host_stale = False
        st = self._engine.download_state(0)
        n = len(st["alive"])
        retired = np.zeros(n, self._st["retired"].dtype)   # newborns (slots added by make_baby on the device) are not retired
        retired[:len(self._st["retired"])] = self._st["retired"][:n]
        st["retired"] = retired | ((st["age"] >= self.retirement_age) & (st["alive"] > 0)).astype(retired.dtype)
        st["number_of_children"] = self._engine.download_children(0)
        st["max_education_level"], st["school_level"] = self._engine.download_education(0)
        self.load_state(st)
        self._dirty = False

    # ------------------------------------------------------------------ the hot path as model methods
    def calculate_criminal_tendency(self) -> None:
        """model.py:1160-1186"""
        e = self._engine_ready()
        e.tendency()
        self._host_stale = True

    def calculate_crime_multiplier(self) -> None:
        """model.py:1144-1157"""
        self.crime_multiplier = float(self._engine_ready().crime_multiplier()[0])

    def reset_oc_embeddedness(self) -> None:
        """model.py:709-718: the per-tick cache is an epoch stamp on the device, bumped by commit_crimes."""
        return None

    def commit_crimes(self) -> None:
        """model.py:1417-1485"""
        e = self._engine_ready()
        outs, _ = e.commit_crimes(self.tick, philox_keys=[self._philox_key])
        o = outs[0]
        if o["error"]:
            raise RuntimeError("crime path error code %d at tick %d" % (o["error"], self.tick))
        self.number_crimes += o["d_number_crimes"]
        self.crime_size_fails += o["d_crime_size_fails"]
        self.facilitator_crimes += o["d_facilitator_crimes"]
        self.facilitator_fails += o["d_facilitator_fails"]
        self.big_crime_from_small_fish += o["d_big_crime_from_small_fish"]
        self.number_offspring_recruited_this_tick += o["d_offspring_recruited"]
        self.number_protected_recruited_this_tick += o["d_protected_recruited"]
        self.people_jailed += o["d_people_jailed"]
        self.number_law_interventions_this_tick += o["d_people_jailed"]
        self._host_stale = True

    def calculate_arrest_rate(self) -> None:
        self.arrest_rate = self.number_arrests_per_year / self.ticks_per_year / self.number_crimes_yearly_per10k / 10000 * self.initial_agents

    _FLOAT_REPORTERS = {"criminal_tendency_mean", "criminal_tencency_sd", "crime_multiplier", "age_mean", "age_sd",
                        "education_level_mean", "education_level_sd", "num_crime_committed_mean",
                        "num_crime_committed_sd"}
    _DEVICE_ONLY = {"tick", "recruited_total", "crimes_this_tick", "embeddedness_evaluations_this_tick",
                    "error_code", "crime_multiplier", "number_crimes", "crime_size_fails", "facilitator_crimes",
                    "facilitator_fails", "big_crime_from_small_fish", "number_offspring_recruited_this_tick",
                    "people_jailed", "number_protected_recruited_this_tick", "number_law_interventions_this_tick"}

    def calculate_fast_reporters(self, from_device: bool = True) -> None:
        """model.py:1687-1735.  With the state on the device the indicators come from the reporter kernel
        (poc_report); from_device=False recomputes them on the host from the downloaded columns."""
        if from_device and self._engine is not None and not self._dirty:
            row = self._engine.report()[0]
            for name, v in zip(REPORTER_NAMES, row):
                if name not in self._DEVICE_ONLY:   # the cumulative counters are kept by commit_crimes() here
                    setattr(self, name, float(v) if name in self._FLOAT_REPORTERS else int(v))
            self.number_jobs = self.employed
            return
        self._pull()
        st = self._st
        alive = st["alive"].astype(bool)
        deg = [np.array([len(s) for s in self._adj[li]], dtype=np.int64) for li in range(len(LAYER_NAMES))]
        self.number_crimes_committed_of_persons = int(st["num_crimes_committed"][alive].sum())
        self.current_oc_members = int(st["oc_member"][alive].sum())
        self.current_num_persons = int(alive.sum())
        for name, col in [("criminal_tendency", "criminal_tendency"), ("age", "age"),
                          ("education_level", "education_level"), ("num_crime_committed", "num_crimes_committed")]:
            v = st[col][alive].astype(np.float64)
            setattr(self, name + "_mean", float(np.mean(v)) if len(v) else float("nan"))
            setattr(self, ("criminal_tencency" if name == "criminal_tendency" else name) + "_sd",
                    float(np.std(v)) if len(v) else float("nan"))
        self.crimes_committed_by_oc_this_tick = int(st["num_crimes_committed_this_tick"][alive & (st["oc_member"] > 0)].sum())
        self.current_prisoners = int(st["prisoner"][alive].sum())
        self.employed = int(st["has_job"][alive].sum())
        self.facilitators = int(st["facilitator"][alive].sum())
        for li, l in enumerate(LAYER_NAMES):
            setattr(self, "tot_%s_link" % l, int(deg[li][alive].sum()) if len(alive) else 0)
        self.number_students = int(st["has_school"][alive].sum())
        self.number_jobs = int(st["has_job"][alive].sum())

    # ------------------------------------------------------------------ setup / step / run
    def setup(self, n_agents: Optional[int] = None, template: Optional[Dict[str, np.ndarray]] = None) -> None:
        self.choose_intervention_setting()
        if n_agents is not None:
            self.initial_agents = n_agents
        if len(self._st["alive"]) == 0:
            self.load_state(synth.synthetic_state(self.initial_agents, seed=self.seed, template=template,
                                                  num_oc_persons=self.num_oc_persons,
                                                  likelihood_of_facilitators=self.likelihood_of_facilitators,
                                                  nat_propensity_m=self.nat_propensity_m,
                                                  nat_propensity_sigma=self.nat_propensity_sigma))
        self.calculate_crime_multiplier()
        self.calculate_criminal_tendency()
        self.calculate_arrest_rate()
        self.datacollector = _DataCollector(MODEL_REPORTERS)
        self.calculate_fast_reporters()
        self.datacollector.collect(self)

    def step(self) -> None:
        """The hot-path subset of model.py:226-288."""
        e = self._engine_ready()
        e.begin_tick(self.tick)
        self._host_stale = True
        self.number_law_interventions_this_tick = 0
        if (self.tick % self.ticks_per_year) == 0 or self.tick == 1:
            self.calculate_criminal_tendency()
            self.calculate_crime_multiplier()
            if self.device_phases & CrimeEngine.PHASE_LABOUR:                                          # model.py:251-256
                self._engine_ready().run_phase(CrimeEngine.PHASE_LABOUR, self.tick, [self._philox_key])
                self._host_stale = True
        if self.device_phases & CrimeEngine.PHASE_WEDDING:                                             # model.py:271
            self._engine_ready().run_phase(CrimeEngine.PHASE_WEDDING, self.tick, [self._philox_key])
            self._host_stale = True
        self.reset_oc_embeddedness()
        self.commit_crimes()
        e = self._engine_ready()
        ph = self.device_phases
        for p in (e.PHASE_RETIRE, e.PHASE_BABY, e.PHASE_XFRIENDS, e.PHASE_XPROF, e.PHASE_FRIENDS):   # model.py:274-278
            if ph & p:
                e.run_phase(p, self.tick, [self._philox_key])
        e.end_tick()
        if ph & e.PHASE_DIE:                                                                          # model.py:284
            e.run_phase(e.PHASE_DIE, self.tick, [self._philox_key])
        self._host_stale = True
        self.schedule.step()
        self.calculate_fast_reporters()
        self.datacollector.collect(self)
        self.tick += 1

    def run(self, n_agents: Optional[int] = None, num_ticks: Optional[int] = None, verbose: bool = False,
            fast: bool = True) -> None:
        """model.py:290-312.  fast=True keeps the whole tick loop on the device (poc_run_ticks: no host
        synchronisation inside) and fills the per-tick reporter rows from the device-side reporter kernel;
        fast=False calls step() per tick like the reference (state downloaded every tick)."""
        if n_agents is None:
            n_agents = self.initial_agents
        if num_ticks is not None:
            self.num_ticks = num_ticks
        self.verbose = verbose
        self.setup(n_agents=n_agents)
        if not fast:
            for _tick in range(1, self.num_ticks + 1):
                self.step()
            return
        e = self._engine_ready()
        e.set_phases(self.device_phases)
        rep = e.run_ticks(self.tick, self.num_ticks, [self._philox_key], reporters=True)[0]
        e.set_phases(0)
        self._host_stale = True
        cn, _ = e.counters()
        if cn[0, 9]:
            raise RuntimeError("crime path error code %d" % cn[0, 9])
        for t in range(self.num_ticks):
            self._apply_device_row(rep[t])   # every calculate_fast_reporters column comes from k_report
            self.schedule.step()
            self.datacollector.collect(self)
            self.tick += 1

    _ROW_SKIP = {"tick", "recruited_total", "crimes_this_tick", "embeddedness_evaluations_this_tick", "error_code"}

    def _apply_device_row(self, row) -> None:
        for name, v in zip(REPORTER_NAMES, row):
            if name not in self._ROW_SKIP:
                setattr(self, name, float(v) if name in self._FLOAT_REPORTERS else int(v))
        self.number_jobs = self.employed

    def frame_from_device_rows(self, rows) -> "pd.DataFrame":
        """The (T+1) x 80 model-vars frame of the reference's DataCollector (extra.py:246-274; row 0 collected after
        setup, model.py:1042, row t after step t, :287) from device reporter rows [T+1][len(REPORTER_NAMES)] whose row 0
        is the state after setup: what a batched run (override mode on the GPU) pickles, so that it has the columns,
        the row count and the snapshot indices of a single run."""
        self.datacollector = _DataCollector(MODEL_REPORTERS)
        self.calculate_arrest_rate()   # ProtonOC.setup (model.py:1031)
        self.tick = 1
        for i, row in enumerate(rows):
            self._apply_device_row(row)
            if i > 0:
                self.tick = i
            self.datacollector.collect(self)
        self.tick = len(rows)
        return self.datacollector.get_model_vars_dataframe()

    def save_data(self, save_dir: str, name: str, snapshot: tuple):
        model_data = self.datacollector.get_model_vars_dataframe()
        if snapshot:
            model_data = model_data.iloc[list(snapshot)]
        path = os.path.join(save_dir, name + ".pkl")
        with open(path, "wb") as f:
            pickle.dump(model_data, f)
        return "Saved: {}".format(path)


def standardize_value(value: str):
    """extra.standardize_value (extra.py:318-338)"""
    from ast import literal_eval
    if any(ch.isdigit() for ch in value):
        try:
            val = literal_eval(value)
        except Exception:
            return value
        if isinstance(val, float) and val.is_integer():
            return int(val)
        return val
    if value == "\"none\"":
        return None
    if value == "true":
        return True
    if value == "false":
        return False
    if "palermo" in value:
        return "palermo"
    if "eindhoven" in value:
        return "eindhoven"
    return value.replace("\"", "")


def parse_behaviorspace_xml(xml_file: str) -> Dict[str, Any]:
    """NetLogo BehaviorSpace xml -> attribute dict (model.py:1635-1659 / run_modes.py:138-164)."""
    map_attr = {"education_rate": "education_modifier", "data_folder": "city", "[num_oc_persons]": "num_oc_persons",
                "num_persons": "initial_agents", "percentage_of_facilitators": "likelihood_of_facilitators"}
    doc = minidom.parse(xml_file)
    run = {"num_ticks": standardize_value(doc.getElementsByTagName("timeLimit")[0].attributes["steps"].value)}
    for par in doc.getElementsByTagName("enumeratedValueSet"):
        attribute = par.attributes["variable"].value.replace("-", "_").replace("?", "").lower()
        if attribute in ("output", "oc_members_scrutinize"):
            continue
        attribute = map_attr.get(attribute, attribute)
        run[attribute] = standardize_value(par.getElementsByTagName("value")[0].attributes["value"].value)
    return run
