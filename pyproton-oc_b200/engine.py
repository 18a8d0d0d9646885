"""CrimeEngine: the device-resident crime path of one batch of replicas, over the C ABI.

State dictionaries use the reference's attribute names (entities.py:49-83) plus `rowptr_<l>` / `col_<l>`
for the nine directed layers in Person.network_names order and `co_off` for Person.num_co_offenses."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _cabi
from ._cabi import (COLS_F64, COLS_I8, COLS_I32, COLS_U8, MAX_GROUP, N_REPORTERS, NUM_LAYERS, AgentCols, Params,
                    PocError, Replay, TickOut, TickRecords, check)

LAYER_NAMES = ["sibling", "offspring", "parent", "partner", "household",
               "friendship", "criminal", "professional", "school"]
CRIMINAL = 6
# C-ABI column -> state-dict key
STATE_KEYS = {"alive": "alive", "male": "male", "oc_member": "oc_member", "facilitator": "facilitator",
              "prisoner": "prisoner", "has_job": "has_job", "has_school": "has_school",
              "target": "target_of_intervention", "job_level": "job_level",
              "education_level": "education_level", "wealth_level": "wealth_level", "age": "age",
              "birth_tick": "birth_tick", "num_crimes": "num_crimes_committed",
              "num_crimes_tick": "num_crimes_committed_this_tick", "father": "father",
              "new_recruit": "new_recruit", "propensity": "propensity", "tendency": "criminal_tendency",
              "sentence": "sentence_countdown", "arrest_weight": "arrest_weight"}
_DTYPES = {**{k: np.uint8 for k in COLS_U8}, **{k: np.int8 for k in COLS_I8},
           **{k: np.int32 for k in COLS_I32}, **{k: np.float64 for k in COLS_F64}}
REPORTER_NAMES = ["number_crimes", "crime_size_fails", "facilitator_crimes", "facilitator_fails",
                  "big_crime_from_small_fish", "number_offspring_recruited_this_tick", "people_jailed",
                  "current_oc_members", "current_prisoners", "tot_criminal_link",
                  "number_crimes_committed_of_persons", "crimes_committed_by_oc_this_tick",
                  "criminal_tendency_mean", "criminal_tencency_sd", "current_num_persons", "crime_multiplier",
                  "age_mean", "age_sd", "education_level_mean", "education_level_sd", "num_crime_committed_mean",
                  "num_crime_committed_sd", "employed", "facilitators", "number_students", "tot_sibling_link",
                  "tot_offspring_link", "tot_parent_link", "tot_partner_link", "tot_household_link",
                  "tot_friendship_link", "tot_professional_link", "tot_school_link",
                  "number_protected_recruited_this_tick", "number_law_interventions_this_tick", "tick",
                  "recruited_total", "crimes_this_tick", "embeddedness_evaluations_this_tick", "error_code",
                  "number_born", "number_deceased", "number_weddings"]
TIMER_CLASSES = ["cells", "tendency", "draw", "search", "embeddedness", "commit", "arrests", "other"]


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


class _PackedState:
    __slots__ = ("blob", "holder", "n")

    def __init__(self, blob, holder, n):
        self.blob, self.holder, self.n = blob, holder, n

    @property
    def nbytes(self) -> int:
        return int(self.blob.nbytes)


class CrimeEngine:
    def __init__(self, n_replicas: int, max_agents: int, max_static_entries: int, max_criminal_entries: int,
                 device: int = 0):
        self.L = _cabi.lib()
        if self.L.poc_device_count() <= 0:
            raise PocError("no CUDA device: the crime path has no CPU fallback")
        h = C.c_void_p()
        check(self.L.poc_ctx_create(device, n_replicas, max_agents, int(max_static_entries),
                                    int(max_criminal_entries), C.byref(h)))
        self.h = h
        self.R, self.max_agents = n_replicas, max_agents
        self.max_static, self.max_criminal = int(max_static_entries), int(max_criminal_entries)
        self.n_agents = [0] * n_replicas
        self.device = device
        self._keepalive = []   # host arrays of uploads queued with wait=False, released at sync()

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        if getattr(self, "h", None):
            self.L.poc_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        check(rc, self.h)

    def sync(self):
        self._ck(self.L.poc_sync(self.h))
        self._keepalive.clear()

    @staticmethod
    def capacity_for(states: Sequence[Dict[str, np.ndarray]], headroom_criminal: int = 0):
        """(max_agents, max_static_entries, max_criminal_entries) that fit all the given states."""
        n = max(len(s["alive"]) for s in states)
        stat = max(sum(len(s["col_%d" % l]) for l in range(NUM_LAYERS) if l != CRIMINAL) for s in states)
        crim = max(len(s["col_%d" % CRIMINAL]) for s in states)
        return n, stat, crim + headroom_criminal

    # ------------------------------------------------------------------ state in / out
    def set_params(self, p: Params, replica: int = -1):
        self._ck(self.L.poc_set_params(self.h, replica, C.byref(p)))

    def upload_state(self, replica: int, st: Dict[str, np.ndarray], wait: bool = True):
        """wait=False queues the upload without synchronising (input errors then surface at sync())."""
        n = int(len(st["alive"]))
        cols, keep = AgentCols(), []
        for f, key in STATE_KEYS.items():
            arr = np.ascontiguousarray(st[key], dtype=_DTYPES[f])
            keep.append(arr)
            setattr(cols, f, arr.ctypes.data)
        self._ck(self.L.poc_upload_agents(self.h, replica, n, C.byref(cols)))
        self.n_agents[replica] = n
        for l in range(NUM_LAYERS):
            rp, col = _i32(st["rowptr_%d" % l]), _i32(st["col_%d" % l])
            co = _i32(st["co_off"]) if l == CRIMINAL else None
            keep += [rp, col, co]  # the uploads are asynchronous until poc_build_graph returns
            self._ck(self.L.poc_upload_layer(self.h, replica, l, rp.ctypes.data, col.ctypes.data if len(col) else None,
                                             co.ctypes.data if (co is not None and len(co)) else None))
        if wait:
            self._ck(self.L.poc_build_graph(self.h, replica))
        else:
            self._ck(self.L.poc_build_graph_async(self.h, replica))
            self._keepalive.append(keep)

    @staticmethod
    def pack_state(st: Dict[str, np.ndarray], pinned: bool = False):
        """One contiguous block holding the whole state (poc_pack_state): build it once per distinct initial state and
        hand it to upload_packed for every replica that starts from it.  pinned=True returns a pinned torch uint8
        tensor's numpy view (asynchronous copies); otherwise a plain numpy array."""
        L = _cabi.lib()
        n = int(len(st["alive"]))
        ne = np.array([len(st["col_%d" % l]) for l in range(NUM_LAYERS)], dtype=np.int64)
        size = int(L.poc_packed_size(n, ne.ctypes.data))
        if pinned:
            import torch
            holder = torch.empty(size, dtype=torch.uint8).pin_memory()
            blob = holder.numpy()
        else:
            holder = blob = np.empty(size, np.uint8)
        cols, keep = AgentCols(), []
        for f, key in STATE_KEYS.items():
            arr = np.ascontiguousarray(st[key], dtype=_DTYPES[f])
            keep.append(arr)
            setattr(cols, f, arr.ctypes.data)
        rps = (C.c_void_p * NUM_LAYERS)()
        cls = (C.c_void_p * NUM_LAYERS)()
        for l in range(NUM_LAYERS):
            rp, col = _i32(st["rowptr_%d" % l]), _i32(st["col_%d" % l])
            keep += [rp, col]
            rps[l], cls[l] = rp.ctypes.data, (col.ctypes.data if len(col) else None)
        co = _i32(st["co_off"])
        rc = L.poc_pack_state(n, C.byref(cols), rps, cls, co.ctypes.data if len(co) else None, blob.ctypes.data, size)
        if rc:
            raise PocError("poc_pack_state failed (%d)" % rc)
        blob_obj = _PackedState(blob, holder, n)
        return blob_obj

    def upload_packed(self, replica: int, packed: "_PackedState"):
        """Queue the upload of a packed state (one copy + the unpack / merge kernels); errors surface at sync()."""
        self._ck(self.L.poc_upload_packed(self.h, replica, packed.blob.ctypes.data, packed.blob.nbytes))
        self.n_agents[replica] = packed.n
        self._keepalive.append(packed)

    def download_agents(self, replica: int = 0, out: Optional[Dict[str, np.ndarray]] = None) -> Dict[str, np.ndarray]:
        """Agent columns of one replica.  `out` (a dict from agent_buffers()) is filled in place when given: writing
        into already-touched host memory is about five times faster than into freshly allocated arrays."""
        n = self.n_agents[replica]
        if out is None:
            out = self.agent_buffers(n)
        cols = AgentCols()
        for f, key in STATE_KEYS.items():
            arr = out[key]
            if arr.dtype != _DTYPES[f] or len(arr) < n or not arr.flags.c_contiguous:
                raise ValueError("download_agents: out[%r] must be a contiguous %s array of at least %d" % (key, _DTYPES[f], n))
            setattr(cols, f, arr.ctypes.data)
        self._ck(self.L.poc_download_agents(self.h, replica, C.byref(cols)))
        return out

    def download_agents_many(self, replica0: int, outs: Sequence[Dict[str, np.ndarray]]) -> Sequence[Dict[str, np.ndarray]]:
        """Agent columns of the replicas replica0 .. replica0 + len(outs) - 1 into the given buffers (agent_buffers()) with
        one call: the copies of the next replicas run while the columns of the one before are handed out."""
        arr = (AgentCols * len(outs))()
        for i, out in enumerate(outs):
            n = self.n_agents[replica0 + i]
            for f, key in STATE_KEYS.items():
                a = out[key]
                if a.dtype != _DTYPES[f] or len(a) < n or not a.flags.c_contiguous:
                    raise ValueError("download_agents_many: out[%r] must be a contiguous %s array of at least %d" % (key, _DTYPES[f], n))
                setattr(arr[i], f, a.ctypes.data)
        self._ck(self.L.poc_download_agents_many(self.h, int(replica0), len(outs), C.byref(arr)))
        return outs

    @staticmethod
    def agent_buffers(n: int) -> Dict[str, np.ndarray]:
        return {key: np.zeros(n, dtype=_DTYPES[f]) for f, key in STATE_KEYS.items()}

    def download_layer(self, replica: int, layer: int):
        n = self.n_agents[replica]
        e = C.c_int64(0)
        self._ck(self.L.poc_layer_entries(self.h, replica, layer, C.byref(e)))
        rowptr = np.empty(n + 1, np.int32)
        col = np.empty(max(e.value, 1), np.int32)
        co = np.empty(max(e.value, 1), np.int32)
        self._ck(self.L.poc_download_layer(self.h, replica, layer, rowptr.ctypes.data, col.ctypes.data,
                                           co.ctypes.data if layer == CRIMINAL else None))
        return rowptr, col[:e.value], (co[:e.value] if layer == CRIMINAL else None)

    def download_state(self, replica: int = 0) -> Dict[str, np.ndarray]:
        st = self.download_agents(replica)
        for l in range(NUM_LAYERS):
            rp, col, co = self.download_layer(replica, l)
            st["rowptr_%d" % l], st["col_%d" % l] = rp, col
            if l == CRIMINAL:
                st["co_off"] = co
        return st

    # ------------------------------------------------------------------ demography phases
    PHASE_BABY, PHASE_FRIENDS, PHASE_DIE, PHASE_RETIRE, PHASE_XFRIENDS, PHASE_XPROF, PHASE_WEDDING, PHASE_ALL = 1, 2, 4, 8, 16, 32, 64, 127
    # the yearly labour-status block (graduate_and_enter_jobmarket + update_unemployment_status, model.py:251-256): needs set_labour
    PHASE_LABOUR, PHASE_EVERY = 128, 255

    def set_demography(self, d):
        self._ck(self.L.poc_set_demography(self.h, C.byref(d)))

    def set_phases(self, mask: int):
        """Which of make_baby (1) / make_friends (2) / make_people_die (4) / retire_persons (8) / remove_excess_friends (16) /
        remove_excess_professional_links (32) / wedding (64) run_ticks runs every tick and whether it runs the labour-status block
        (128) on the yearly ticks, in the order of ProtonOC.step (default 0)."""
        self._ck(self.L.poc_set_phases(self.h, int(mask)))
        self._phases = int(mask)

    def set_labour(self, lb):
        """Tables of PHASE_LABOUR (params.make_labour)."""
        self._ck(self.L.poc_set_labour(self.h, C.byref(lb)))

    def upload_education(self, replica: int, max_education_level, school_level):
        """Person.max_education_level and the diploma level of Person.my_school (0 = none) per agent slot; without this call
        both are 0 (nobody continues past the school he is in)."""
        a = np.ascontiguousarray(max_education_level, dtype=np.int8)
        b = np.ascontiguousarray(school_level, dtype=np.int8)
        assert len(a) == self.n_agents[replica] and len(b) == self.n_agents[replica]
        self._ck(self.L.poc_upload_education(self.h, replica, a.ctypes.data, b.ctypes.data))

    def download_education(self, replica: int = 0):
        a, b = np.zeros(self.n_agents[replica], np.int8), np.zeros(self.n_agents[replica], np.int8)
        self._ck(self.L.poc_download_education(self.h, replica, a.ctypes.data, b.ctypes.data))
        return a, b

    def _refresh_counts(self):
        v = C.c_int32(0)
        for r in range(self.R):
            self._ck(self.L.poc_n_agents(self.h, r, C.byref(v)))
            self.n_agents[r] = int(v.value)

    def run_phase(self, phase: int, tick: int, philox_keys: Sequence[int]):
        keys = np.ascontiguousarray(philox_keys, dtype=np.uint64)
        assert len(keys) == self.R
        self._ck(self.L.poc_run_phase(self.h, int(phase), int(tick), keys.ctypes.data))
        self._refresh_counts()

    def upload_children(self, replica: int, n_children):
        a = _i32(n_children)
        assert len(a) == self.n_agents[replica]
        self._ck(self.L.poc_upload_children(self.h, replica, a.ctypes.data))

    def download_children(self, replica: int = 0) -> np.ndarray:
        out = np.zeros(self.n_agents[replica], np.int32)
        self._ck(self.L.poc_download_children(self.h, replica, out.ctypes.data))
        return out

    # ------------------------------------------------------------------ hot path
    def begin_tick(self, tick: int):
        self._ck(self.L.poc_begin_tick(self.h, tick))

    def end_tick(self):
        self._ck(self.L.poc_end_tick(self.h))

    def tendency(self):
        self._ck(self.L.poc_tendency(self.h))

    def crime_multiplier(self) -> np.ndarray:
        out = np.zeros(self.R, np.float64)
        self._ck(self.L.poc_crime_multiplier(self.h, out.ctypes.data))
        return out

    def set_crime_multiplier(self, value: float, replica: int = 0):
        self._ck(self.L.poc_set_crime_multiplier(self.h, replica, float(value)))

    def commit_crimes(self, tick: int, replays: Optional[Sequence[Optional[dict]]] = None,
                      philox_keys: Optional[Sequence[int]] = None, records: bool = False):
        """reset_oc_embeddedness + commit_crimes for every replica.  Returns (list of per-replica counter
        dicts, records-of-replica-0 or None)."""
        keep = []
        rp_ptrs = None
        if replays is not None:
            rp_ptrs = (C.c_void_p * self.R)()
            for r in range(self.R):
                d = replays[r] if r < len(replays) else None
                if d is None:
                    rp_ptrs[r] = None
                    continue
                rp = Replay()

                def arr(key, dt, default):
                    a = np.ascontiguousarray(d.get(key, default), dtype=dt)
                    keep.append(a)
                    return a
                ui, us = arr("u_init", np.float64, np.zeros(0)), arr("u_size", np.float64, np.zeros(0))
                fp = arr("fac_pick", np.int32, np.full(len(ui), -1))
                ua, use = arr("u_arrest", np.float64, np.zeros(0)), arr("u_sentence", np.float64, np.zeros(0))
                rp.n_crimes, rp.u_init, rp.u_size, rp.fac_pick = len(ui), ui.ctypes.data, us.ctypes.data, fp.ctypes.data
                rp.u_arrest_count = float(d.get("u_arrest_count", 0.0))
                rp.n_u_arrest, rp.u_arrest = len(ua), ua.ctypes.data
                rp.n_u_sentence, rp.u_sentence = len(use), use.ctypes.data
                if d.get("arrest_idx") is not None:
                    ai = arr("arrest_idx", np.int32, np.zeros(0))
                    rp.n_arrest_idx, rp.arrest_idx = len(ai), ai.ctypes.data
                else:
                    rp.n_arrest_idx, rp.arrest_idx = -1, None
                keep.append(rp)
                rp_ptrs[r] = C.addressof(rp)
        keys = np.ascontiguousarray(philox_keys if philox_keys is not None else np.zeros(self.R), dtype=np.uint64)
        outs = (TickOut * self.R)()
        rec, bufs = None, None
        if records:
            n = self.max_agents
            ccap, mcap = n // 16 + 64, (n // 16 + 64) * 4 + 256
            bufs = {"initiator": np.full(ccap, -1, np.int32), "k": np.full(ccap, -1, np.int32),
                    "group_off": np.zeros(ccap + 1, np.int32), "group_mem": np.full(mcap, -1, np.int32),
                    "arrested": np.full(mcap, -1, np.int32)}
            rec = TickRecords(ccap, mcap, mcap, *(bufs[k].ctypes.data for k in
                                                  ["initiator", "k", "group_off", "group_mem", "arrested"]))
        self._ck(self.L.poc_commit_crimes(self.h, tick, rp_ptrs, keys.ctypes.data, outs,
                                          C.byref(rec) if rec is not None else None))
        res = [outs[r].as_dict() for r in range(self.R)]
        recs = None
        if records:
            ng, na = res[0]["n_crimes"], res[0]["n_arrests"]
            off = bufs["group_off"]
            recs = {"initiator": bufs["initiator"][:ng].copy(), "k": bufs["k"][:ng].copy(),
                    "groups": [bufs["group_mem"][off[g]:off[g + 1]].copy() for g in range(ng)],
                    "arrested": bufs["arrested"][:na].copy()}
        return res, recs

    def run_ticks(self, tick0: int, n_ticks: int, philox_keys: Sequence[int], reporters: bool = True,
                  out: Optional[np.ndarray] = None):
        """n_ticks ticks of the hot path for every replica, no host synchronisation inside.  reporters=True returns
        the [R, n_ticks, len(REPORTER_NAMES)] rows (into `out` when given: a float64 array of that shape)."""
        keys = np.ascontiguousarray(philox_keys, dtype=np.uint64)
        assert len(keys) == self.R
        rep = None
        if reporters:
            rep = out if out is not None else np.zeros((self.R, n_ticks, N_REPORTERS), np.float64)
            if rep.shape != (self.R, n_ticks, N_REPORTERS) or rep.dtype != np.float64 or not rep.flags.c_contiguous:
                raise ValueError("run_ticks: out must be a contiguous float64 array of shape (R, n_ticks, %d)" % N_REPORTERS)
        self._ck(self.L.poc_run_ticks(self.h, tick0, n_ticks, keys.ctypes.data,
                                      rep.ctypes.data if rep is not None else None))
        if getattr(self, "_phases", 0) & 1:
            self._refresh_counts()   # births added agent slots on the device
        return rep

    def run_ticks_device(self, tick0: int, n_ticks: int, philox_keys: Sequence[int]):
        """run_ticks with the reporter rows left on the device: returns a torch tensor [R, n_ticks, len(REPORTER_NAMES)]
        that aliases ctx memory (valid until the next run) — what the multi-GPU ensemble hands to the NCCL gather."""
        import torch
        keys = np.ascontiguousarray(philox_keys, dtype=np.uint64)
        assert len(keys) == self.R
        ptr, tcap = C.c_void_p(), C.c_int(0)
        self._ck(self.L.poc_run_ticks_device(self.h, tick0, n_ticks, keys.ctypes.data, C.byref(ptr), C.byref(tcap)))

        class _Rows:   # the CUDA array interface is how torch adopts foreign device memory without a copy
            __cuda_array_interface__ = {"shape": (self.R, tcap.value, N_REPORTERS), "typestr": "<f8", "data": (ptr.value, False),
                                        "version": 3, "strides": None}
        t = torch.as_tensor(_Rows(), device=torch.device("cuda", self.device))
        return t[:, :n_ticks, :]

    def report(self) -> np.ndarray:
        """calculate_fast_reporters for the current device state: [R, len(REPORTER_NAMES)]."""
        out = np.zeros((self.R, N_REPORTERS), np.float64)
        self._ck(self.L.poc_report(self.h, out.ctypes.data))
        return out

    def counters(self):
        cn = np.zeros((self.R, _cabi.CN_COUNT), np.int32)
        ed = np.zeros(self.R, np.uint64)
        self._ck(self.L.poc_read_counters(self.h, cn.ctypes.data, ed.ctypes.data))
        return cn, ed

    def stats(self, reset: bool = False) -> Dict[str, int]:
        """Algorithmic-work counters summed over replicas."""
        a = np.zeros((self.R, 16), np.uint64)
        self._ck(self.L.poc_read_stats(self.h, a.ctypes.data, 1 if reset else 0))
        s = a.sum(axis=0)
        names = ["emb_requests", "emb_full", "emb_bytes", "emb_nodes", "search_crimes", "search_bytes",
                 "search_candidates", "spare"] + ["phase%d" % i for i in range(8)]
        return {k: int(v) for k, v in zip(names, s)}

    def stats_per_replica(self) -> Dict[str, np.ndarray]:
        """The same counters per replica: name -> uint64 [R]."""
        a = np.zeros((self.R, 16), np.uint64)
        self._ck(self.L.poc_read_stats(self.h, a.ctypes.data, 0))
        names = ["emb_requests", "emb_full", "emb_bytes", "emb_nodes", "search_crimes", "search_bytes",
                 "search_candidates", "spare"] + ["phase%d" % i for i in range(8)]
        return {k: a[:, i].copy() for i, k in enumerate(names)}

    # ------------------------------------------------------------------ timers
    def set_timing(self, on: bool):
        self._ck(self.L.poc_set_timing(self.h, 1 if on else 0))

    def timers_reset(self):
        self._ck(self.L.poc_timers_reset(self.h))

    def timers(self):
        ms, n = np.zeros(8, np.float64), np.zeros(8, np.int64)
        self._ck(self.L.poc_timers_read(self.h, ms.ctypes.data, n.ctypes.data))
        return dict(zip(TIMER_CLASSES, ms.tolist())), dict(zip(TIMER_CLASSES, n.tolist()))

    def marker(self, slot: int):
        self._ck(self.L.poc_marker(self.h, slot))

    def marker_elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_double(0)
        self._ck(self.L.poc_marker_elapsed_ms(self.h, a, b, C.byref(ms)))
        return float(ms.value)

    def launch_count(self) -> int:
        return int(self.L.poc_launch_count(self.h))

    # ------------------------------------------------------------------ stage hooks
    def rings(self, sources: Sequence[int], d: int, exact: bool, replica: int = 0) -> List[np.ndarray]:
        src = _i32(sources)
        off = np.zeros(len(src) + 1, np.int32)
        cap = max(1, len(src) * self.n_agents[replica])
        mem = np.empty(cap, np.int32)
        self._ck(self.L.poc_rings(self.h, replica, len(src), src.ctypes.data, d, 1 if exact else 0,
                                  off.ctypes.data, mem.ctypes.data, cap))
        return [mem[off[i]:off[i + 1]].copy() for i in range(len(src))]

    def social_proximity(self, a: Sequence[int], b: Sequence[int], replica: int = 0) -> np.ndarray:
        a, b = _i32(a), _i32(b)
        out = np.zeros(len(a), np.float64)
        self._ck(self.L.poc_social_proximity(self.h, replica, len(a), a.ctypes.data, b.ctypes.data, out.ctypes.data))
        return out

    def embeddedness(self, agents: Sequence[int], replica: int = 0) -> np.ndarray:
        a = _i32(agents)
        out = np.zeros(len(a), np.float64)
        self._ck(self.L.poc_embeddedness(self.h, replica, len(a), a.ctypes.data, out.ctypes.data))
        return out

    def embeddedness_accumulated(self, agents: Sequence[int], over=None, replica: int = 0) -> np.ndarray:
        """Test hook: the evaluations `agents`, in this order, on the accumulating tick-global meta graph,
        exactly as one reference commit_crimes computes them (model.py:1515-1536).  over: int32 [m, 4] rows
        (evaluation index, a, b, multiplicity in use).  The values become the cache of the next commit_crimes."""
        a = _i32(agents)
        ov = np.zeros((0, 4), np.int32) if over is None else np.ascontiguousarray(over, dtype=np.int32).reshape(-1, 4)
        out = np.zeros(len(a), np.float64)
        self._ck(self.L.poc_embeddedness_accumulated(self.h, replica, len(a), a.ctypes.data, len(ov),
                                                     ov.ctypes.data if len(ov) else None, out.ctypes.data))
        return out

    def preload_embeddedness(self, agents: Sequence[int], values, replica: int = 0):
        """Test hook: hand the next commit_crimes its per-tick oc_embeddedness cache (entities.py:531)."""
        a = _i32(agents)
        v = np.ascontiguousarray(values, dtype=np.float64)
        assert len(a) == len(v)
        self._ck(self.L.poc_preload_embeddedness(self.h, replica, len(a), a.ctypes.data, v.ctypes.data))

    def candidate_weights(self, selfs: Sequence[int], cands: Sequence[int], replica: int = 0) -> np.ndarray:
        a, b = _i32(selfs), _i32(cands)
        out = np.zeros(len(a), np.float64)
        self._ck(self.L.poc_candidate_weights(self.h, replica, len(a), a.ctypes.data, b.ctypes.data, out.ctypes.data))
        return out

    def find_accomplices(self, initiators: Sequence[int], ks: Sequence[int], fac_pick=None, u_fac=None,
                         replica: int = 0):
        a, k = _i32(initiators), _i32(ks)
        n = len(a)
        groups = np.full((max(n, 1), MAX_GROUP), -1, np.int32)
        glen = np.zeros(max(n, 1), np.int32)
        fails = np.zeros((max(n, 1), 3), np.int32)
        fp = _i32(fac_pick) if fac_pick is not None else None
        uf = np.ascontiguousarray(u_fac, dtype=np.float64) if u_fac is not None else None
        self._ck(self.L.poc_find_accomplices(self.h, replica, n, a.ctypes.data, k.ctypes.data,
                                             fp.ctypes.data if fp is not None else None,
                                             uf.ctypes.data if uf is not None else None,
                                             groups.ctypes.data, glen.ctypes.data, fails.ctypes.data))
        return [groups[i, :glen[i]].copy() for i in range(n)], fails[:n]
