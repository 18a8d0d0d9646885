"""ctypes front-end of oracle/hotpath.c (the CPU restatement).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package never
imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_hotpath.so")
MAX_CELLS, MAX_TAB, MAX_GROUP, NL = 32, 16, 32, 9

# Palermo tables on the hot path, restated from the reference's input data
# (protonoc/simulator/inputs/palermo/data/crime_rate_by_gender_and_age_range.csv,
#  inputs/general/data/num_co_offenders_dist.csv, inputs/palermo/data/conviction_length.csv).
# Kept separately from the product's copy so the oracle stays self-contained.
# NOTE the values are what the reference actually computes with: it loads the CSVs through
# pandas.read_csv's default (fast, not correctly rounded, 15-significant-digit) float parser, so e.g.
# "0.00220888483881099" becomes 0.0022088848388109.  Dumped with repr() from the loaded reference model.
C_RANGE = [  # (male, age_from, age_to, c) in file order
    (0, 0, 13, 0.000419095530846447), (1, 0, 13, 0.0022088848388109),
    (0, 14, 17, 0.0223485080056817), (1, 14, 17, 0.15015920337183),
    (0, 18, 24, 0.0510827740184912), (1, 18, 24, 0.301852378661966),
    (0, 25, 34, 0.0634378338266472), (1, 25, 34, 0.303561858487422),
    (0, 35, 44, 0.0642953324827881), (1, 35, 44, 0.275109864781256),
    (0, 45, 54, 0.0488535842667994), (1, 45, 54, 0.19961940993774),
    (0, 55, 64, 0.0308128546507127), (1, 55, 64, 0.126843825050472),
    (0, 65, 200, 0.0111212661509359), (1, 65, 200, 0.0536705333692047)]
CO_OFFENDERS = [(1, 0.8744024339984563), (2, 0.0960396945979063), (3, 0.0207191652676337),
                (4, 0.0055633579042241), (5, 0.0018980490549422), (6, 0.0013772991768372)]
CONVICTION = [  # (months, F, M)
    (1, 0.135167397354478, 0.078947866071776), (3, 0.19269400131653, 0.131143083933566),
    (6, 0.262419887899263, 0.236163183846043), (12, 0.205487284788606, 0.238697628478702),
    (24, 0.140485495612569, 0.175982293376721), (36, 0.032234256152626, 0.059220489881712),
    (60, 0.021495330654007, 0.049035337226683), (120, 0.008189481085251, 0.022837548624071),
    (240, 0.00182686513667, 0.007972568560726)]


class Params(C.Structure):
    _fields_ = [
        ("n_cells", C.c_int32), ("cell_male", C.c_int32 * MAX_CELLS), ("cell_age_from", C.c_int32 * MAX_CELLS),
        ("cell_age_to", C.c_int32 * MAX_CELLS), ("cell_c", C.c_double * MAX_CELLS),
        ("n_sizes", C.c_int32), ("size_n", C.c_int32 * MAX_TAB), ("size_cdf", C.c_double * MAX_TAB),
        ("n_sent", C.c_int32), ("sent_months", C.c_double * MAX_TAB), ("sent_cdf_m", C.c_double * MAX_TAB),
        ("sent_cdf_f", C.c_double * MAX_TAB),
        ("propensity_threshold", C.c_double), ("crimes_yearly_per10k", C.c_double),
        ("arrests_per_year", C.c_double), ("punishment_length", C.c_double),
        ("threshold_use_facilitators", C.c_double), ("facilitator_repression_multiplier", C.c_double),
        ("good_guy_threshold", C.c_double),
        ("max_accomplice_radius", C.c_int32), ("oc_embeddedness_radius", C.c_int32),
        ("facilitator_repression", C.c_int32), ("oc_boss_repression", C.c_int32),
        ("ticks_between_intervention", C.c_int32), ("intervention_start", C.c_int32),
        ("intervention_end", C.c_int32), ("this_is_a_big_crime", C.c_int32),
        ("ticks_per_year", C.c_int32), ("_pad", C.c_int32)]


class Cols(C.Structure):
    _u8 = ["alive", "male", "oc_member", "facilitator", "prisoner", "has_job", "has_school", "target"]
    _i8 = ["job_level", "education_level", "wealth_level"]
    _i32 = ["age", "birth_tick", "num_crimes", "num_crimes_tick", "father", "new_recruit"]
    _f64 = ["propensity", "tendency", "sentence", "arrest_weight"]
    _fields_ = ([(n, C.c_void_p) for n in _u8] + [(n, C.c_void_p) for n in _i8] +
                [(n, C.c_void_p) for n in _i32] + [(n, C.c_void_p) for n in _f64])


# names used in exported state dicts (oracle/ref_harness.export_state) -> Cols fields
STATE_KEYS = {"alive": "alive", "male": "male", "oc_member": "oc_member", "facilitator": "facilitator",
              "prisoner": "prisoner", "has_job": "has_job", "has_school": "has_school",
              "target": "target_of_intervention", "job_level": "job_level",
              "education_level": "education_level", "wealth_level": "wealth_level", "age": "age",
              "birth_tick": "birth_tick", "num_crimes": "num_crimes_committed",
              "num_crimes_tick": "num_crimes_committed_this_tick", "father": "father",
              "new_recruit": "new_recruit", "propensity": "propensity", "tendency": "criminal_tendency",
              "sentence": "sentence_countdown", "arrest_weight": "arrest_weight"}
_DTYPES = {**{k: np.uint8 for k in Cols._u8}, **{k: np.int8 for k in Cols._i8},
           **{k: np.int32 for k in Cols._i32}, **{k: np.float64 for k in Cols._f64}}


class Replay(C.Structure):
    _fields_ = [("n_crimes", C.c_int32), ("u_init", C.c_void_p), ("u_size", C.c_void_p), ("fac_pick", C.c_void_p),
                ("u_arrest_count", C.c_double), ("n_u_arrest", C.c_int32), ("u_arrest", C.c_void_p),
                ("n_u_sentence", C.c_int32), ("u_sentence", C.c_void_p),
                ("n_arrest_idx", C.c_int32), ("arrest_idx", C.c_void_p)]


class TickOut(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ["n_crimes", "n_members", "n_arrests", "n_emb_evals", "d_number_crimes", "d_crime_size_fails",
                 "d_facilitator_crimes", "d_facilitator_fails", "d_big_crime_from_small_fish",
                 "d_offspring_recruited", "d_protected_recruited", "d_people_jailed", "n_recruited", "error"]
                ] + [("traversed_edges", C.c_int64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class Demog(C.Structure):
    """Mirror of orc_demog (hotpath.c)."""
    _fields_ = [("mort", (C.c_double * 128) * 2), ("fert", (C.c_double * 64) * 3), ("poisson_cdf", C.c_double * 32),
                ("prop_q", C.c_double * 1024), ("mort_max_age", C.c_int32), ("n_prop", C.c_int32),
                ("retirement_age", C.c_int32), ("_pad", C.c_int32), ("weddings_mean", C.c_double)]


def make_demog(nat_propensity_m: float = 1.0, nat_propensity_sigma: float = 0.25, mortality=None, fertility=None) -> Demog:
    """mortality: (female[120], male[120]) yearly probabilities; fertility: [3][64]; defaults = the Palermo tables."""
    from pyproton_oc_b200 import tables
    fem, mal = mortality if mortality is not None else (tables.MORTALITY_FEMALE, tables.MORTALITY_MALE)
    fert = fertility if fertility is not None else tables.FERTILITY
    d = Demog()
    for a in range(128):
        d.mort[0][a] = float(fem[a]) if a < len(fem) else 1.0
        d.mort[1][a] = float(mal[a]) if a < len(mal) else 1.0
    for k in range(3):
        for a in range(64):
            d.fert[k][a] = float(fert[k][a])
    for i, v in enumerate(tables.poisson_cdf(3.0, 32)):
        d.poisson_cdf[i] = v
    for i, v in enumerate(tables.lognormal_quantiles(nat_propensity_m, nat_propensity_sigma, 1024)):
        d.prop_q[i] = v
    d.mort_max_age, d.n_prop = len(fem) - 1, 1024
    d.retirement_age = 65
    d.weddings_mean = tables.NUMBER_WEDDINGS_MEAN
    return d


class Labour(C.Structure):
    """Mirror of orc_labour (hotpath.c); same layout as poc_labour."""
    _fields_ = [("n_levels", C.c_int32), ("primary_age", C.c_int32), ("end_age", C.c_int32 * 8),
                ("work_n", (C.c_int32 * 2) * 8), ("work_val", ((C.c_int32 * 8) * 2) * 8), ("work_cdf", ((C.c_double * 8) * 2) * 8),
                ("wealth_n", (C.c_int32 * 2) * 8), ("wealth_val", ((C.c_int32 * 8) * 2) * 8),
                ("wealth_cdf", ((C.c_double * 8) * 2) * 8),
                ("inactive", (C.c_double * 128) * 2), ("range_age", (C.c_uint8 * 128) * 2)]


def make_labour(**tables_) -> Labour:
    """The tables of the yearly labour-status block; defaults = the Palermo tables (pyproton_oc_b200.tables)."""
    from pyproton_oc_b200.params import fill_labour
    return fill_labour(Labour(), **tables_)


_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "hotpath.c")):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle_hotpath.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int32, C.POINTER(Cols), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_get_cols.argtypes = [C.c_void_p, C.POINTER(Cols)]
        L.orc_layer_entries.restype = C.c_int64
        L.orc_layer_entries.argtypes = [C.c_void_p, C.c_int]
        L.orc_get_layer.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_begin_tick.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        L.orc_end_tick.argtypes = [C.c_void_p]
        L.orc_tendency.argtypes = [C.c_void_p, C.POINTER(Params)]
        L.orc_crime_multiplier.restype = C.c_double
        L.orc_crime_multiplier.argtypes = [C.c_void_p, C.POINTER(Params)]
        L.orc_ring.restype = C.c_int32
        L.orc_ring.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
        L.orc_social_proximity.restype = C.c_double
        L.orc_social_proximity.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        L.orc_embeddedness.restype = C.c_double
        L.orc_embeddedness.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        L.orc_embeddedness_ex.restype = C.c_double
        L.orc_embeddedness_ex.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_reset_embeddedness.argtypes = [C.c_void_p]
        L.orc_embeddedness_accumulated.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        L.orc_preload_embeddedness.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        L.orc_candidate_weight.restype = C.c_double
        L.orc_candidate_weight.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
        L.orc_find_accomplices.restype = C.c_int32
        L.orc_find_accomplices.argtypes = [C.c_void_p, C.POINTER(Params), C.c_int32, C.c_int32, C.c_int32, C.c_double,
                                           C.c_void_p, C.c_void_p]
        L.orc_commit_crimes.restype = C.c_int32
        L.orc_commit_crimes.argtypes = [C.c_void_p, C.POINTER(Params), C.c_int32, C.c_double, C.POINTER(Replay),
                                        C.c_uint64, C.POINTER(TickOut), C.c_int32, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_run_ticks.restype = C.c_int32
        L.orc_run_ticks.argtypes = [C.c_void_p, C.POINTER(Params), C.c_int32, C.c_int32, C.c_uint64,
                                    C.POINTER(C.c_double), C.c_void_p]
        L.orc_philox4x32_10.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_reserve.argtypes = [C.c_void_p, C.c_int32]
        L.orc_n_agents.restype = C.c_int32
        L.orc_n_agents.argtypes = [C.c_void_p]
        L.orc_get_children.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_set_children.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_demog_counters.argtypes = [C.c_void_p, C.c_void_p]
        for fn in ("orc_make_baby", "orc_make_friends", "orc_make_people_die", "orc_wedding"):
            getattr(L, fn).restype = C.c_int32
            getattr(L, fn).argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(Demog), C.c_int32, C.c_uint64]
        L.orc_retire_persons.restype = C.c_int32
        L.orc_retire_persons.argtypes = [C.c_void_p, C.POINTER(Demog)]
        for fn in (L.orc_remove_excess_friends, L.orc_remove_excess_professional):
            fn.restype = C.c_int32
            fn.argtypes = [C.c_void_p, C.c_int32, C.c_uint64]
        L.orc_sizeof_labour.restype = C.c_int32
        L.orc_set_labour.argtypes = [C.c_void_p, C.POINTER(Labour)]
        L.orc_get_education.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_set_education.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_labour_yearly.restype = C.c_int32
        L.orc_labour_yearly.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(Demog), C.c_int32, C.c_uint64]
        assert L.orc_sizeof_labour() == C.sizeof(Labour)
        L.orc_remove_excess_diagnostic.restype = C.c_int32
        L.orc_remove_excess_diagnostic.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_uint64]
        L.orc_run_ticks_phases.restype = C.c_int32
        L.orc_run_ticks_phases.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(Demog), C.c_int32, C.c_int32, C.c_int32,
                                           C.c_uint64, C.POINTER(C.c_double), C.c_void_p]
        L.orc_sizeof_demog.restype = C.c_int32
        assert L.orc_sizeof_demog() == C.sizeof(Demog), (L.orc_sizeof_demog(), C.sizeof(Demog))
        L.orc_sizeof_params.restype = C.c_int32
        L.orc_sizeof_tick_out.restype = C.c_int32
        assert L.orc_sizeof_params() == C.sizeof(Params), (L.orc_sizeof_params(), C.sizeof(Params))
        assert L.orc_sizeof_tick_out() == C.sizeof(TickOut)
        _lib = L
    return _lib


def numpy_choice_cdf(weights) -> np.ndarray:
    """CDF exactly as extra.weighted_n_of + Generator.choice build it
    (extra.py:114-128; numpy: cumsum(p); cdf /= cdf[-1])."""
    p = [float(x) for x in weights]
    sump = sum(p)
    p = np.array([i / sump for i in p], dtype=np.float64)
    cdf = np.cumsum(p)
    cdf /= cdf[-1]
    return cdf


def propensity_threshold(m: float, sigma: float, thr: float) -> float:
    """entities.py:350-353 (Q4), evaluated once on the host."""
    return float(np.exp(m - sigma ** 2 / 2) + thr * np.sqrt(np.exp(sigma) ** 2 - 1) * np.exp(m + sigma ** 2 / 2))


DEFAULTS = dict(nat_propensity_m=1.0, nat_propensity_sigma=0.25, nat_propensity_threshold=1.0,
                number_crimes_yearly_per10k=2000, number_arrests_per_year=30, punishment_length=1,
                threshold_use_facilitators=4, facilitator_repression_multiplier=2.0, good_guy_threshold=0.6,
                max_accomplice_radius=2, oc_embeddedness_radius=2, facilitator_repression=False,
                oc_boss_repression=False, ticks_between_intervention=12, intervention_start=13,
                intervention_end=999, this_is_a_big_crime=3, ticks_per_year=12)


def make_params(overrides: Optional[Dict] = None, c_range=None, co_offenders=None, conviction=None) -> Params:
    d = dict(DEFAULTS)
    if overrides:
        d.update({k: v for k, v in overrides.items() if k in d})
    c_range = c_range or C_RANGE
    co_offenders = co_offenders or CO_OFFENDERS
    conviction = conviction or CONVICTION
    p = Params()
    p.n_cells = len(c_range)
    for i, (male, a0, a1, c) in enumerate(c_range):
        p.cell_male[i], p.cell_age_from[i], p.cell_age_to[i], p.cell_c[i] = int(male), int(a0), int(a1), float(c)
    p.n_sizes = len(co_offenders)
    cdf = numpy_choice_cdf([r[-1] for r in co_offenders])
    for i, r in enumerate(co_offenders):
        p.size_n[i], p.size_cdf[i] = int(r[0]), float(cdf[i])
    p.n_sent = len(conviction)
    cf, cm = numpy_choice_cdf([r[1] for r in conviction]), numpy_choice_cdf([r[2] for r in conviction])
    for i, r in enumerate(conviction):
        p.sent_months[i], p.sent_cdf_f[i], p.sent_cdf_m[i] = float(r[0]), float(cf[i]), float(cm[i])
    p.propensity_threshold = propensity_threshold(d["nat_propensity_m"], d["nat_propensity_sigma"],
                                                  d["nat_propensity_threshold"])
    p.crimes_yearly_per10k = float(d["number_crimes_yearly_per10k"])
    p.arrests_per_year = float(d["number_arrests_per_year"])
    p.punishment_length = float(d["punishment_length"])
    p.threshold_use_facilitators = float(d["threshold_use_facilitators"])
    p.facilitator_repression_multiplier = float(d["facilitator_repression_multiplier"])
    p.good_guy_threshold = float(d["good_guy_threshold"])
    for k in ["max_accomplice_radius", "oc_embeddedness_radius", "facilitator_repression", "oc_boss_repression",
              "ticks_between_intervention", "intervention_start", "intervention_end", "this_is_a_big_crime",
              "ticks_per_year"]:
        setattr(p, k, int(d[k]))
    return p


class OracleState:
    """Owns one oracle state built from an exported-state dict."""

    def __init__(self, st: Dict[str, np.ndarray]):
        self.L = lib()
        self.n = int(len(st["alive"]))
        keep = []
        cols = Cols()
        for f, key in STATE_KEYS.items():
            arr = np.ascontiguousarray(st[key], dtype=_DTYPES[f])
            keep.append(arr)
            setattr(cols, f, arr.ctypes.data)
        rp = (C.c_void_p * NL)()
        cp = (C.c_void_p * NL)()
        for l in range(NL):
            r = np.ascontiguousarray(st["rowptr_%d" % l], dtype=np.int32)
            c = np.ascontiguousarray(st["col_%d" % l], dtype=np.int32)
            keep += [r, c]
            rp[l], cp[l] = r.ctypes.data, c.ctypes.data
        co = np.ascontiguousarray(st["co_off"], dtype=np.int32)
        self.h = self.L.orc_create(self.n, C.byref(cols), rp, cp, co.ctypes.data)
        if "max_education_level" in st:
            self.set_education(st["max_education_level"], st["school_level"])

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- demography
    def reserve(self, cap: int):
        self.L.orc_reserve(self.h, int(cap))

    def _refresh_n(self):
        self.n = int(self.L.orc_n_agents(self.h))

    def children(self) -> np.ndarray:
        self._refresh_n()
        out = np.zeros(self.n, np.int32)
        self.L.orc_get_children(self.h, out.ctypes.data)
        return out

    def set_children(self, arr):
        a = np.ascontiguousarray(arr, dtype=np.int32)
        assert len(a) == self.n
        self.L.orc_set_children(self.h, a.ctypes.data)

    def demog_counters(self):
        out = np.zeros(3, np.int64)
        self.L.orc_demog_counters(self.h, out.ctypes.data)
        return {"number_deceased": int(out[0]), "number_born": int(out[1]), "number_weddings": int(out[2])}

    def make_baby(self, p, dm, tick, key):
        r = int(self.L.orc_make_baby(self.h, C.byref(p), C.byref(dm), tick, key))
        self._refresh_n()
        return r

    def make_friends(self, p, dm, tick, key):
        return int(self.L.orc_make_friends(self.h, C.byref(p), C.byref(dm), tick, key))

    def make_people_die(self, p, dm, tick, key):
        return int(self.L.orc_make_people_die(self.h, C.byref(p), C.byref(dm), tick, key))

    def wedding(self, p, dm, tick, key):
        return int(self.L.orc_wedding(self.h, C.byref(p), C.byref(dm), tick, key))

    def retire_persons(self, dm):
        return int(self.L.orc_retire_persons(self.h, C.byref(dm)))

    def remove_excess_friends(self, tick, key):
        return int(self.L.orc_remove_excess_friends(self.h, tick, key))

    def remove_excess_professional(self, tick, key):
        return int(self.L.orc_remove_excess_professional(self.h, tick, key))

    # ---- the yearly labour-status block (model.py:251-256)
    def set_labour(self, lb):
        self.L.orc_set_labour(self.h, C.byref(lb))

    def set_education(self, max_edu, school_level):
        self._refresh_n()
        a = np.ascontiguousarray(max_edu, dtype=np.int8)
        b = np.ascontiguousarray(school_level, dtype=np.int8)
        assert len(a) == self.n and len(b) == self.n
        self.L.orc_set_education(self.h, a.ctypes.data, b.ctypes.data)

    def education(self):
        self._refresh_n()
        a, b = np.empty(self.n, np.int8), np.empty(self.n, np.int8)
        self.L.orc_get_education(self.h, a.ctypes.data, b.ctypes.data)
        return a, b

    def labour_yearly(self, p, dm, tick, key):
        """graduate_and_enter_jobmarket + update_unemployment_status: (graduations, status draws)"""
        r = int(self.L.orc_labour_yearly(self.h, C.byref(p), C.byref(dm), tick, key))
        if r < 0:
            raise RuntimeError("oracle: set_labour first")
        return r % 65536, r // 65536

    def remove_excess_diagnostic(self, mode: int, tick: int, key: int) -> int:
        """Both remove_excess phases in the reference's sequential order (mode 0) or as the rounds of its parallel form (mode 1;
        returns the number of rounds).  Diagnostic: no kernel follows either."""
        return int(self.L.orc_remove_excess_diagnostic(self.h, mode, tick, key))

    def run_ticks_phases(self, p, dm, phases: int, tick0: int, n_ticks: int, philox_key: int, crime_multiplier: float = 0.0):
        """Hot path + the demography phases in `phases` (1 make_baby, 2 make_friends, 4 make_people_die, 8 retire_persons,
        16 remove_excess_friends, 32 remove_excess_professional_links, 64 wedding, 128 the yearly labour-status block)."""
        mult = C.c_double(crime_multiplier)
        totals = np.zeros(6, dtype=np.int64)
        rc = self.L.orc_run_ticks_phases(self.h, C.byref(p), C.byref(dm), phases, tick0, n_ticks, philox_key, C.byref(mult),
                                         totals.ctypes.data)
        if rc:
            raise RuntimeError("oracle: set_labour first" if rc == -2 else "oracle: slot capacity exhausted (reserve more)")
        self._refresh_n()
        return float(mult.value), totals

    # ---- read back
    def cols(self) -> Dict[str, np.ndarray]:
        self._refresh_n()
        out, cols = {}, Cols()
        for f, key in STATE_KEYS.items():
            out[key] = np.empty(self.n, dtype=_DTYPES[f])
            setattr(cols, f, out[key].ctypes.data)
        self.L.orc_get_cols(self.h, C.byref(cols))
        return out

    def layer(self, l: int):
        self._refresh_n()
        e = int(self.L.orc_layer_entries(self.h, l))
        rowptr = np.empty(self.n + 1, dtype=np.int32)
        col = np.empty(max(e, 1), dtype=np.int32)
        co = np.empty(max(e, 1), dtype=np.int32)
        self.L.orc_get_layer(self.h, l, rowptr.ctypes.data, col.ctypes.data, co.ctypes.data if l == 6 else None)
        return rowptr, col[:e], (co[:e] if l == 6 else None)

    # ---- stages
    def begin_tick(self, tick, tpy=12):
        self.L.orc_begin_tick(self.h, tick, tpy)

    def end_tick(self):
        self.L.orc_end_tick(self.h)

    def tendency(self, p: Params):
        self.L.orc_tendency(self.h, C.byref(p))

    def crime_multiplier(self, p: Params) -> float:
        return float(self.L.orc_crime_multiplier(self.h, C.byref(p)))

    def ring(self, src: int, d: int, exact: bool) -> np.ndarray:
        out = np.empty(self.n + 1, dtype=np.int32)
        m = self.L.orc_ring(self.h, src, d, 1 if exact else 0, out.ctypes.data)
        return out[:m].copy()

    def social_proximity(self, a: int, b: int) -> float:
        return float(self.L.orc_social_proximity(self.h, a, b))

    def embeddedness(self, a: int, radius: int = 2) -> float:
        return float(self.L.orc_embeddedness(self.h, a, radius))

    def embeddedness_ex(self, a: int, radius: int = 2):
        ball = np.empty(self.n + 1, dtype=np.int32)
        dist = np.empty(self.n + 1, dtype=np.float64)
        nb, nasym = C.c_int32(0), C.c_int32(0)
        v = float(self.L.orc_embeddedness_ex(self.h, a, radius, ball.ctypes.data, dist.ctypes.data, C.byref(nb),
                                             C.byref(nasym)))
        return v, ball[:nb.value].copy(), dist[:nb.value].copy(), int(nasym.value)

    def embeddedness_accumulated(self, slots, radius: int = 2, over=None) -> np.ndarray:
        """H1 exact mode: the evaluations `slots`, in this order, on the accumulating meta graph (model.py:1515-1536).
        over: int32 [m, 4] rows (evaluation index, a, b, multiplicity in use), see hotpath.c."""
        slots = np.ascontiguousarray(slots, dtype=np.int32)
        ov = np.zeros((0, 4), np.int32) if over is None else np.ascontiguousarray(over, dtype=np.int32).reshape(-1, 4)
        ov = np.ascontiguousarray(ov[np.argsort(ov[:, 0], kind="stable")])
        out = np.zeros(len(slots), np.float64)
        self.L.orc_embeddedness_accumulated(self.h, len(slots), slots.ctypes.data, radius, len(ov), ov.ctypes.data, out.ctypes.data)
        return out

    def preload_embeddedness(self, slots, vals):
        slots = np.ascontiguousarray(slots, dtype=np.int32)
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        assert len(slots) == len(vals)
        self.L.orc_preload_embeddedness(self.h, len(slots), slots.ctypes.data, vals.ctypes.data)

    def reset_embeddedness(self):
        self.L.orc_reset_embeddedness(self.h)

    def candidate_weight(self, self_slot: int, cand: int, radius: int = 2) -> float:
        return float(self.L.orc_candidate_weight(self.h, self_slot, cand, radius))

    def find_accomplices(self, p: Params, self_slot: int, k: int, fac_pick: int = -1, u_fac: float = 0.0):
        group = np.empty(MAX_GROUP, dtype=np.int32)
        fails = np.zeros(3, dtype=np.int32)
        g = self.L.orc_find_accomplices(self.h, C.byref(p), self_slot, k, fac_pick, u_fac,
                                        group.ctypes.data, fails.ctypes.data)
        if g < 0:
            raise RuntimeError("oracle find_accomplices error %d" % g)
        return group[:g].copy(), fails

    def commit_crimes(self, p: Params, tick: int, crime_multiplier: float, replay: Optional[dict] = None,
                      philox_key: int = 0, rec_cap: int = 1 << 16):
        out = TickOut()
        ini = np.full(rec_cap, -1, dtype=np.int32)
        kk = np.full(rec_cap, -1, dtype=np.int32)
        goff = np.zeros(rec_cap + 1, dtype=np.int32)
        gmem = np.full(rec_cap * 8, -1, dtype=np.int32)
        arrested = np.full(4096, -1, dtype=np.int32)
        rp, keep = None, []
        if replay is not None:
            rp = Replay()

            def arr(key, dt):
                a = np.ascontiguousarray(replay.get(key, np.zeros(0)), dtype=dt)
                keep.append(a)
                return a
            ui, us = arr("u_init", np.float64), arr("u_size", np.float64)
            fp = np.ascontiguousarray(replay.get("fac_pick", np.full(len(ui), -1)), dtype=np.int32)
            ua, use = arr("u_arrest", np.float64), arr("u_sentence", np.float64)
            keep.append(fp)
            rp.n_crimes, rp.u_init, rp.u_size, rp.fac_pick = len(ui), ui.ctypes.data, us.ctypes.data, fp.ctypes.data
            rp.u_arrest_count = float(replay.get("u_arrest_count", 0.0))
            rp.n_u_arrest, rp.u_arrest = len(ua), ua.ctypes.data
            rp.n_u_sentence, rp.u_sentence = len(use), use.ctypes.data
            if replay.get("arrest_idx") is not None:
                ai = arr("arrest_idx", np.int32)
                rp.n_arrest_idx, rp.arrest_idx = len(ai), ai.ctypes.data
            else:
                rp.n_arrest_idx, rp.arrest_idx = -1, None
        self.L.orc_commit_crimes(self.h, C.byref(p), tick, crime_multiplier, C.byref(rp) if rp else None,
                                 philox_key, C.byref(out), rec_cap, ini.ctypes.data, kk.ctypes.data,
                                 goff.ctypes.data, gmem.ctypes.data, arrested.ctypes.data)
        ng = min(out.n_crimes, rec_cap)
        groups = [gmem[goff[g]:goff[g + 1]].copy() for g in range(ng)]
        return {"out": out.as_dict(), "initiator": ini[:ng].copy(), "k": kk[:ng].copy(), "groups": groups,
                "arrested": arrested[:out.n_arrests].copy()}

    def run_ticks(self, p: Params, tick0: int, n_ticks: int, philox_key: int, crime_multiplier: float = 0.0):
        mult = C.c_double(crime_multiplier)
        totals = np.zeros(6, dtype=np.int64)
        self.L.orc_run_ticks(self.h, C.byref(p), tick0, n_ticks, philox_key, C.byref(mult), totals.ctypes.data)
        return float(mult.value), totals


def philox4x32_10(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32_10(c.ctypes.data, k.ctypes.data, o.ctypes.data)
    return o
