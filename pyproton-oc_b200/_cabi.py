"""ctypes binding of libprotonoc_b200.so (include/protonoc_b200.h).

There is no CPU fallback: if the library is missing it is built with nvcc for sm_100a, and creating a
context without a CUDA device raises."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libprotonoc_b200.so")
SRC_DIR = os.path.join(_HERE, "csrc")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "protonoc_b200.h")

NUM_LAYERS, MAX_CELLS, MAX_TAB, MAX_GROUP, N_REPORTERS, CN_COUNT = 9, 32, 16, 32, 43, 24
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--fmad=false", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


class PocError(RuntimeError):
    pass


class Demography(C.Structure):
    """poc_demography (include/protonoc_b200.h)."""
    _fields_ = [("mort", (C.c_double * 128) * 2), ("fert", (C.c_double * 64) * 3), ("poisson_cdf", C.c_double * 32),
                ("prop_q", C.c_double * 1024), ("mort_max_age", C.c_int32), ("n_prop", C.c_int32),
                ("retirement_age", C.c_int32), ("_pad", C.c_int32), ("weddings_mean", C.c_double)]


class Labour(C.Structure):
    """poc_labour (include/protonoc_b200.h): the tables of the yearly labour-status block (model.py:251-256, 1351-1380)."""
    _fields_ = [("n_levels", C.c_int32), ("primary_age", C.c_int32), ("end_age", C.c_int32 * 8),
                ("work_n", (C.c_int32 * 2) * 8), ("work_val", ((C.c_int32 * 8) * 2) * 8), ("work_cdf", ((C.c_double * 8) * 2) * 8),
                ("wealth_n", (C.c_int32 * 2) * 8), ("wealth_val", ((C.c_int32 * 8) * 2) * 8),
                ("wealth_cdf", ((C.c_double * 8) * 2) * 8),
                ("inactive", (C.c_double * 128) * 2), ("range_age", (C.c_uint8 * 128) * 2)]


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


COLS_U8 = ["alive", "male", "oc_member", "facilitator", "prisoner", "has_job", "has_school", "target"]
COLS_I8 = ["job_level", "education_level", "wealth_level"]
COLS_I32 = ["age", "birth_tick", "num_crimes", "num_crimes_tick", "father", "new_recruit"]
COLS_F64 = ["propensity", "tendency", "sentence", "arrest_weight"]


class AgentCols(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in COLS_U8 + COLS_I8 + COLS_I32 + COLS_F64]


class Replay(C.Structure):
    _fields_ = [("n_crimes", C.c_int32), ("u_init", C.c_void_p), ("u_size", C.c_void_p), ("fac_pick", C.c_void_p),
                ("u_arrest_count", C.c_double), ("n_u_arrest", C.c_int32), ("u_arrest", C.c_void_p),
                ("n_u_sentence", C.c_int32), ("u_sentence", C.c_void_p),
                ("n_arrest_idx", C.c_int32), ("arrest_idx", C.c_void_p)]


TICK_OUT_FIELDS = ["n_crimes", "n_members", "n_arrests", "n_emb_evals", "d_number_crimes", "d_crime_size_fails",
                   "d_facilitator_crimes", "d_facilitator_fails", "d_big_crime_from_small_fish",
                   "d_offspring_recruited", "d_protected_recruited", "d_people_jailed", "n_recruited", "error"]


class TickOut(C.Structure):
    _fields_ = [(n, C.c_int32) for n in TICK_OUT_FIELDS] + [("traversed_edges", C.c_int64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class TickRecords(C.Structure):
    _fields_ = [("cap_crimes", C.c_int32), ("cap_members", C.c_int32), ("cap_arrests", C.c_int32),
                ("initiator", C.c_void_p), ("k", C.c_void_p), ("group_off", C.c_void_p),
                ("group_mem", C.c_void_p), ("arrested", C.c_void_p)]


# every symbol include/protonoc_b200.h declares: (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "poc_device_count": (C.c_int, []),
    "poc_ctx_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.POINTER(_P)]),
    "poc_ctx_destroy": (C.c_int, [_P]),
    "poc_last_error": (C.c_char_p, [_P]),
    "poc_sync": (C.c_int, [_P]),
    "poc_set_timing": (C.c_int, [_P, C.c_int]),
    "poc_timers_reset": (C.c_int, [_P]),
    "poc_timers_read": (C.c_int, [_P, _P, _P]),
    "poc_launch_count": (C.c_longlong, [_P]),
    "poc_read_counters": (C.c_int, [_P, _P, _P]),
    "poc_read_stats": (C.c_int, [_P, _P, C.c_int]),
    "poc_marker": (C.c_int, [_P, C.c_int]),
    "poc_marker_elapsed_ms": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "poc_set_params": (C.c_int, [_P, C.c_int, C.POINTER(Params)]),
    "poc_upload_agents": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(AgentCols)]),
    "poc_upload_layer": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P]),
    "poc_build_graph": (C.c_int, [_P, C.c_int]),
    "poc_build_graph_async": (C.c_int, [_P, C.c_int]),
    "poc_download_agents": (C.c_int, [_P, C.c_int, C.POINTER(AgentCols)]),
    "poc_download_agents_many": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "poc_layer_entries": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_int64)]),
    "poc_download_layer": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P]),
    "poc_begin_tick": (C.c_int, [_P, C.c_int]),
    "poc_end_tick": (C.c_int, [_P]),
    "poc_tendency": (C.c_int, [_P]),
    "poc_crime_multiplier": (C.c_int, [_P, _P]),
    "poc_set_crime_multiplier": (C.c_int, [_P, C.c_int, C.c_double]),
    "poc_commit_crimes": (C.c_int, [_P, C.c_int, _P, _P, _P, _P]),
    "poc_run_ticks": (C.c_int, [_P, C.c_int, C.c_int, _P, _P]),
    "poc_report": (C.c_int, [_P, _P]),
    "poc_run_ticks_device": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P]),
    "poc_rings": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_int, C.c_int, _P, _P, C.c_int64]),
    "poc_social_proximity": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P]),
    "poc_embeddedness": (C.c_int, [_P, C.c_int, C.c_int, _P, _P]),
    "poc_set_demography": (C.c_int, [_P, _P]),
    "poc_set_phases": (C.c_int, [_P, C.c_int]),
    "poc_run_phase": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "poc_upload_children": (C.c_int, [_P, C.c_int, _P]),
    "poc_download_children": (C.c_int, [_P, C.c_int, _P]),
    "poc_set_labour": (C.c_int, [_P, _P]),
    "poc_upload_education": (C.c_int, [_P, C.c_int, _P, _P]),
    "poc_download_education": (C.c_int, [_P, C.c_int, _P, _P]),
    "poc_sizeof_labour": (C.c_int, []),
    "poc_n_agents": (C.c_int, [_P, C.c_int, _P]),
    "poc_packed_size": (C.c_int64, [C.c_int, _P]),
    "poc_pack_state": (C.c_int, [C.c_int, _P, _P, _P, _P, _P, C.c_int64]),
    "poc_upload_packed": (C.c_int, [_P, C.c_int, _P, C.c_int64]),
    "poc_embeddedness_accumulated": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_int, _P, _P]),
    "poc_preload_embeddedness": (C.c_int, [_P, C.c_int, C.c_int, _P, _P]),
    "poc_candidate_weights": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P]),
    "poc_find_accomplices": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, _P]),
    "poc_sizeof_params": (C.c_int, []),
    "poc_sizeof_tick_out": (C.c_int, []),
}

_lib = None


def sources():
    import glob
    return sorted(glob.glob(os.path.join(SRC_DIR, "*.cu")) + glob.glob(os.path.join(SRC_DIR, "*.cuh"))) + [HEADER]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    stale = force or not os.path.exists(LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in sources())
    if stale:
        cmd = ["nvcc"] + NVCC_FLAGS + ["-o", LIB_PATH, os.path.join(SRC_DIR, "poc_api.cu")]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        # Only a MISSING library is built here.  A stale one is rebuilt by build() (__graft_entry__.build(), which compares
        # modification times against every CUDA source): doing that on import would let the ranks of a torchrun job, whose
        # snapshot copies carry arbitrary mtimes, compile over each other's shared object.
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError here means the library does not match the header
            fn.restype, fn.argtypes = res, args
        if L.poc_sizeof_params() != C.sizeof(Params) or L.poc_sizeof_tick_out() != C.sizeof(TickOut) or \
                L.poc_sizeof_labour() != C.sizeof(Labour):
            raise PocError("ctypes structs out of sync with libprotonoc_b200.so")
        _lib = L
    return _lib


def check(rc: int, ctx=None):
    if rc != 0:
        msg = lib().poc_last_error(ctx)
        raise PocError("libprotonoc_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))
