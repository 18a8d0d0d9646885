// poc_device.cuh -- device-side data model and block-cooperative primitives (sm_100a).
//
// Data layout in HBM (per ctx = batch of R replicas, every array is [R][capacity], replica-major):
//   attr      u32   age | flags<<8 | job<<16 | edu<<20 | wealth<<24   (one 4-byte gather scores a candidate)
//   tendency  f64, propensity f64, sentence f64, arrest_w f64
//   birth, ncrimes, ncrimes_tick, father, new_recruit   i32
//   u_rowptr  i32 [Ncap+1], u_ent u32 [ucap]: the eight non-criminal layers merged into one directed CSR,
//             entry = neighbour slot (22 bit) | one-sided flag (bit 22) | also-criminal flag (bit 23) | layer mask (8 bit) << 24 ;
//             rows ascending
//   c_rowptr  i32 [Ncap+1], c_ent int2 [ccap]: criminal layer {neighbour, num_co_offenses | flags | static mask << 24}, rows
//             ascending.  The flag and the mask copy make the multiplicity of an edge (model.py:1528-1535) a single
//             load from either row: no row intersection on the traversal paths.
// Reference mapping: Person attributes entities.py:49-83, the nine neighbour sets entities.py:38-47.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/protonoc_b200.h"

#define FULL_MASK 0xffffffffu

// ---- packed attribute word ------------------------------------------------------------------------
enum : uint32_t {
    F_ALIVE = 1u << 8, F_MALE = 1u << 9, F_OC = 1u << 10, F_FAC = 1u << 11, F_PRISONER = 1u << 12,
    F_HASJOB = 1u << 13, F_HASSCHOOL = 1u << 14, F_TARGET = 1u << 15,
    F_DYING = 1u << 31   /* set by make_people_die for the length of that phase only */
};
__host__ __device__ __forceinline__ int attr_age(uint32_t a) { return (int)(a & 0xffu); }
__host__ __device__ __forceinline__ int attr_job(uint32_t a) { return (int)((a >> 16) & 0xfu); }
__host__ __device__ __forceinline__ int attr_edu(uint32_t a) { return (int)((a >> 20) & 0xfu); }
__host__ __device__ __forceinline__ int attr_wealth(uint32_t a) { return (int)((a >> 24) & 0xfu); }

// layer -> bit of the 8-bit mask in the multiplex entry (criminal is stored separately)
__host__ __device__ __forceinline__ int layer_bit(int layer) { return layer < POC_CRIMINAL ? layer : layer - 1; }
#define MB_SIB 0x01u
#define MB_OFF 0x02u
#define MB_PAR 0x04u
#define MB_PARTNER 0x08u
#define MB_HH 0x10u
#define MB_FRIEND 0x20u
#define MB_PROF 0x40u
#define MB_SCHOOL 0x80u
#define ENT_COL(e) ((int)((e) & 0x003fffffu))
// ENT_ASYM on the entry u -> v: the reverse entry v -> u does not carry the same multiplicity (it is missing, has another
// number of layers, or is weighed through a criminal entry while this one is not).  Links are made in pairs, so the flag
// is rare (one-sided clears of get_caught, entities.py:601-602, and Person.die, :652-661); a traversal that wants every
// undirected meta edge ONCE (poc_embed3.cuh) takes an entry when v > u or the flag is set.  The flag may be set where it
// is not needed (never the other way round): it is a performance hint, not state; downloads do not see it.
// u_mid[u] = index of the first entry of row u such a traversal must read: the first entry with a neighbour slot > u
// (rows ascend), or the row's begin when an entry before that carries the flag.
#define ENT_ASYM 0x00400000u
#define ENT_CRIM 0x00800000u /* the same neighbour also has an entry in the criminal row */
#define ENT_MASK(e) ((e) >> 24)
// criminal entry .y = num_co_offenses (22 bit) | one-sided flag (bit 22) | removed flag (bit 23) | copy of the static layer mask of the same neighbour << 24
// A removed entry (Person.die took the owner's dead neighbour out of its criminal set, entities.py:652-661) stays in the
// row as a tombstone with count 0 and mask 0 until the row is laid out again: traversals skip it, weights see 0.
#define CE_CNT(y) ((int)((unsigned)(y) & 0x003fffffu))
#define CE_ASYM 0x00400000 /* like ENT_ASYM: the reverse criminal entry is missing, removed or weighs differently */
#define CE_DEAD 0x00800000
#define CE_IS_DEAD(y) (((unsigned)(y) & 0x00800000u) != 0u)
#define CE_SMASK(y) ((unsigned)(y) >> 24)

// counters per replica (ints), cumulative unless noted
enum {
    CN_NUMBER_CRIMES = 0, CN_SIZE_FAILS, CN_FAC_CRIMES, CN_FAC_FAILS, CN_BIG_CRIME, CN_OFFSPRING_RECRUITED,
    CN_PROTECTED_RECRUITED, CN_PEOPLE_JAILED, CN_RECRUITED, CN_ERROR, CN_TICK_CRIMES, CN_TICK_MEMBERS,
    CN_TICK_ARRESTS, CN_TICK_EMB, CN_TICK_RECRUITED, CN_N_ALIVE, CN_DECEASED, CN_BORN, CN_FRIENDSHIPS, CN_WEDDINGS, CN_COUNT = 24
};

// algorithmic-work counters per replica (what an ideal implementation must move once)
enum { ST_EMB_REQUESTS = 0, ST_EMB_FULL, ST_EMB_BYTES, ST_EMB_NODES, ST_SEARCH_CRIMES, ST_SEARCH_BYTES,
       ST_SEARCH_CANDIDATES, ST_SPARE, ST_PH0 = 8, ST_COUNT = 16 };  // ST_PH0..: per-phase cycles (POC_PHASES builds)

// per replica-group tick state kept on the device so that a whole tick is a parameter-free launch sequence
// (CUDA graph replay): current tick, embeddedness-cache epoch, reporter row
struct TickInfo { int tick, epoch, rep, pad; };

struct DevCtx {
    int R, Ncap, Ccap, Mcap, Pcap;  // replicas, agent / crime / member / pair capacities per replica
    int ecap;                       // staged-edge capacity per embeddedness block (entries)
    int r0, Rg, grp;                // the replica range [r0, r0 + Rg) this launch serves (one range per stream)
    TickInfo *ti;                   // [8] tick state per replica group
    long long ucap, ccap;           // static / criminal entry capacity per replica
    int *n_agents;                  // [R]
    // agents
    uint32_t *attr;
    int32_t *birth, *ncrimes, *ncrimes_tick, *father, *new_recruit;
    double *propensity, *tendency, *sentence, *arrest_w;
    // graph
    int32_t *u_rowptr; uint32_t *u_ent;
    int32_t *u_rowptr_tmp; uint32_t *u_ent_tmp;   // second multiplex buffer: make_baby lays the rows out again (k_multiplex_insert)
    int32_t *u_mid, *u_mid_tmp;     // [R][Ncap] per row: first entry a once-per-pair traversal has to look at (see ENT_ASYM)
    int *u_par;                     // [R] which multiplex CSR is current
    int32_t *c_rowptr; int2 *c_ent;
    int32_t *c_rowptr_tmp; int2 *c_ent_tmp; int32_t *c_addlen;
    int *c_par;                     // [R] which criminal CSR is current: 0 = c_rowptr/c_ent, 1 = the _tmp pair (k_commit
                                    // writes the other one and flips this, so nothing is copied back)
    int32_t *pair_new_ex;           // [R][Pcap+1] exclusive count of new links before each unique pair
    int32_t *ins_row;               // [R][Pcap] index of the first unique pair of every row that receives pairs
    int32_t *ins_rowid, *ins_before;  // [R][Pcap], [R][Pcap+1] that row's id / new links in the receiving rows before it
    // parameters / scalars
    poc_params *params;             // [R]
    double *crime_mult;             // [R]
    int *counters;                  // [R][CN_COUNT]
    int *prev_counters;             // [R][CN_COUNT] snapshot taken by k_draw at the start of commit_crimes
    unsigned long long *edges;      // [R] traversed multiplex entries (bench metric)
    unsigned long long *stats;      // [R][ST_COUNT] algorithmic-work counters, see enum below
    long long *layer_tot;           // [R][8] directed entries per non-criminal layer over alive agents (tot_*_link)
    // cells
    int32_t *cell_members;          // [R][Ncap] ascending slot inside each cell
    int2 *pw_leaf; double *pw_sum; int pw_cap;   // [R][pw_cap] pieces of the pairwise sums in k_tendency_eps
    int32_t *cell_off;              // [R][POC_MAX_CELLS+1]
    int32_t *cell_strict;           // [R][POC_MAX_CELLS] counts with age_from < age (Q8)
    double *cdf;                    // [R][Ncap] the tick's per-cell CDFs in cell_members order; kept from tick to tick:
    int *cdf_valid, *cdf_nz;        // [R], [R][POC_MAX_CELLS] k_draw builds them again only when a tendency, a cell's member
                                    // list or the crime multiplier changed since (k_cells, k_tendency_pre, uploads clear the flag)
    // crimes of the tick
    int32_t *n_crimes;              // [R]
    int32_t *crime_init, *crime_k, *crime_flags;  // [R][Ccap]; flags bit0 = started by OC member
    int32_t *cr_target, *cr_nacc, *cr_d, *cr_state, *cr_fails;  // [R][Ccap]
    int32_t *cr_acc;                // [R][Ccap][POC_MAX_GROUP]
    int32_t *search_list, *n_search;  // [R][Ccap], [R]
    // embeddedness requests
    int32_t *emb_list;              // [2][R][Ncap]
    int32_t *n_emb;                 // [2][R]
    int32_t *emb_big, *n_emb_big;   // [R][big_cap], [R] evaluations the small blocks hand to the large-block launch (poc_embed3.cuh)
    int big_cap;
    int emb_flags;                  // A/B switches (POC_EMB_FLAGS): 1 = poc_embed3.cuh takes criminal entries from both rows,
                                    // 2 = k_search stages the ring in shared memory
    int32_t *emb_stamp;             // [R][Ncap]
    double *emb_val;                // [R][Ncap]
    // groups of the tick, canonical
    int32_t *group_off, *group_mem; // [R][Ccap+1], [R][Mcap]
    unsigned long long *pairs;      // [R][Pcap]
    int32_t *pair_cnt;              // [R][Pcap]
    int32_t *n_pairs;               // [R]
    int32_t *arrested, *n_arrested; // [R][Mcap], [R]
    double *aw, *ap, *acdf;         // [R][Mcap] arrest sampler scratch
    // replay (device copies), NULL pointers => Philox
    double *rp_u_init, *rp_u_size; int32_t *rp_fac_pick;   // [R][Ccap]
    double *rp_u_arrest, *rp_u_sentence; int32_t *rp_arrest_idx;  // [R][Mcap]
    int32_t *rp_hdr;                // [R][8]: use_replay, n_crimes, n_u_arrest, n_u_sentence, n_arrest_idx
    double *rp_u_arrest_count;      // [R]
    unsigned long long *philox_key; // [R]
    // per-block scratch (block slot = blockIdx.y * gridDim.x + blockIdx.x)
    int32_t *scr_list;              // [slots][Ncap+1]
    double *scr_key;                // [slots][Ncap+1]
    unsigned long long *scr_dist;   // [slots][Ncap+1]
    unsigned *scr_edges;            // [slots][ecap] staged induced entries of one embeddedness evaluation
    const double *invw;             // [256] 1.0 / i (meta-edge weight of multiplicity i, model.py:1536)
    // demography (make_baby, make_friends, make_people_die)
    const struct DevDemog *demog;   // tables, one set for the ctx
    int32_t *n_children;            // [R][Ncap] Person.number_of_children
    int32_t *ncrim_dead;            // [R][Ncap] tombstones in the agent's criminal row (tot_criminal_link counts live entries)
    int32_t *dm_list;               // [R][Ncap] mothers of the tick / agents dying in the tick
    int32_t *dm_n;                  // [R][4] their numbers; [2] = queued multiplex insertions / friendships
    unsigned long long *dm_pairs;   // [R][DPcap] directed entries to add: u << 40 | v << 8 | layer mask
    unsigned long long *dm_defer;   // [R][1024] friendships whose reverse entry has no slot yet
    int DPcap;
    // the yearly labour-status block (model.py:251-256): tables, one set for the ctx; per agent max_education_level (low nibble)
    // and the diploma level of my_school (high nibble, read only while F_HASSCHOOL is set)
    const struct DevLabour *labour;
    uint8_t *edu2;                  // [R][Ncap]
    int scr_slots;
    int slot_stride;                // scratch slot of block bx of replica r = r * slot_stride + bx in EVERY traversal kernel: the
                                    // replica groups run concurrently on their own streams and must not meet in a slot
    double *reporters;              // [R][Tcap][POC_N_REPORTERS]
    int Tcap;
};

struct DevDemog {
    double mort[2][128];      // yearly death probability by [male][age]; past the table 1 (entities.py:615-626)
    double fert[3][64];       // yearly fertility by [min(children, 2)][age] (entities.py:605-613)
    double poisson_cdf[32];   // CDF of Poisson(3) (model.py:619)
    double prop_q[1024];      // quantiles of the newborns' propensity lognormal (entities.py:71)
    int mort_max_age, n_prop;
    int retirement_age, _pad;
    double weddings_mean;     // number_weddings_mean (model.py:184-186)
};

struct DevLabour {            // = poc_labour (include/protonoc_b200.h)
    int n_levels, primary_age;
    int end_age[8];
    int work_n[8][2]; int work_val[8][2][8]; double work_cdf[8][2][8];
    int wealth_n[8][2]; int wealth_val[8][2][8]; double wealth_cdf[8][2][8];
    double inactive[2][128];
    uint8_t range_age[2][128];
};

// ---- Philox4x32-10 (Salmon et al. SC'11) -----------------------------------------------------------
enum { ST_CRIME = 0, ST_FAC = 1, ST_ARREST_N = 2, ST_ARREST = 3, ST_SENT = 4, ST_DIE = 5, ST_FACREP = 6, ST_BABY = 7,
       ST_BABYATTR = 8, ST_FRIEND_N = 9, ST_FRIEND = 10, ST_XFRIEND = 11, ST_XPROF = 12, ST_WED_N = 13, ST_WED_EGO = 14, ST_WED_PICK = 15,
       ST_GRAD = 16, ST_UNEMP = 17 };
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// two doubles in [0,1) with 53 random bits each (numpy's random() mapping)
__device__ __forceinline__ void philox_u2(unsigned long long key, int tick, int stage, int idx, double &u0, double &u1) {
    uint32_t o[4];
    philox4x32_10((uint32_t)tick, (uint32_t)stage, (uint32_t)idx, 0u, (uint32_t)key, (uint32_t)(key >> 32), o);
    unsigned long long a = ((unsigned long long)o[1] << 32) | o[0], b = ((unsigned long long)o[3] << 32) | o[2];
    u0 = __dmul_rn((double)(a >> 11), 1.0 / 9007199254740992.0);
    u1 = __dmul_rn((double)(b >> 11), 1.0 / 9007199254740992.0);
}

// one uniform from the counter (tick, stage, idx, sub): phases that draw per (agent, entry)
__device__ __forceinline__ double philox_u1s(unsigned long long key, int tick, int stage, int idx, int sub) {
    uint32_t o[4];
    philox4x32_10((uint32_t)tick, (uint32_t)stage, (uint32_t)idx, (uint32_t)sub, (uint32_t)key, (uint32_t)(key >> 32), o);
    unsigned long long a = ((unsigned long long)o[1] << 32) | o[0];
    return __dmul_rn((double)(a >> 11), 1.0 / 9007199254740992.0);
}

// searchsorted(cdf, u, side='right')
__device__ __forceinline__ int ss_right(const double *cdf, int n, double u) {
    int lo = 0, hi = n;
    while (lo < hi) { int m = (lo + hi) >> 1; if (cdf[m] <= u) lo = m + 1; else hi = m; }
    return lo;
}

// ---- block primitives ------------------------------------------------------------------------------
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

// exclusive scan of one int per thread across the block; returns the exclusive prefix, *total = block sum.
// smem `ws` needs 33 ints.  All threads must call.
__device__ __forceinline__ int block_exscan(int v, int *ws, int *total) {
    int lane = lane_id(), w = warp_id(), nw = (blockDim.x + 31) >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(FULL_MASK, x, o); if (lane >= o) x += y; }
    if (lane == 31) ws[w] = x;
    __syncthreads();
    if (w == 0) {
        int s = (lane < nw) ? ws[lane] : 0, t = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(FULL_MASK, t, o); if (lane >= o) t += y; }
        ws[lane] = t - s;
        if (lane == 31) ws[32] = t;
    }
    __syncthreads();
    int res = ws[w] + x - v;
    if (total) *total = ws[32];
    __syncthreads();
    return res;
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}

// Order-defined float64 running sum, executed redundantly by all 32 lanes of one warp: loads and stores are
// coalesced, every lane carries the same accumulator, and the only serial part is the dependent DADD chain.
// Adds v (lane l holds element base + l; lanes past the end must pass 0.0, acc + 0.0 == acc) left to right;
// returns this lane's inclusive prefix, leaves the running total in acc.
__device__ __forceinline__ double warp_chain32(double v, double &acc) {
    double mine = 0.0;
    int lane = lane_id();
#pragma unroll
    for (int j = 0; j < 32; j++) {
        double vj = __shfl_sync(FULL_MASK, v, j);
        acc = acc + vj;
        if (lane == j) mine = acc;
    }
    return mine;
}

// ---- multiplex traversal ---------------------------------------------------------------------------
struct Graph {
    const int32_t *u_rowptr; const uint32_t *u_ent;
    int32_t *u_mid;
    const int32_t *c_rowptr; const int2 *c_ent;
};
__device__ __forceinline__ Graph graph_of(const DevCtx &C, int r) {
    Graph g;
    int upar = C.u_par[r];
    g.u_rowptr = (upar ? C.u_rowptr_tmp : C.u_rowptr) + (size_t)r * (C.Ncap + 1);
    g.u_ent = (upar ? C.u_ent_tmp : C.u_ent) + (size_t)r * C.ucap;
    g.u_mid = (upar ? C.u_mid_tmp : C.u_mid) + (size_t)r * C.Ncap;
    int par = C.c_par[r];
    g.c_rowptr = (par ? C.c_rowptr_tmp : C.c_rowptr) + (size_t)r * (C.Ncap + 1);
    g.c_ent = (par ? C.c_ent_tmp : C.c_ent) + (size_t)r * C.ccap;
    return g;
}

// Block-cooperative frontier BFS over the union of the nine OUT-neighbour sets
// (Person._agents_in_radius, entities.py:469-480; ring peeling entities.py:482-524).
// list[0..nsrc) holds the sources (already written by the caller); level L members are appended at
// lvl_off[L]..lvl_off[L+1] (lvl_off[0] = 0, lvl_off[1] = nsrc).  One warp expands one frontier node;
// newly discovered neighbours are compacted with ballot/popc and a single atomicAdd per warp.
// bitmap: ceil(N/32) words of shared memory, cleared here.  Returns the total list length.
__device__ int bfs_levels(const Graph &g, int nwords, unsigned *bitmap, int *list, int nsrc, int depth, int *lvl_off,
                          int *s_count, unsigned long long &edges, unsigned long long &rowbytes) {
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) bitmap[i] = 0u;
    __syncthreads();
    for (int i = threadIdx.x; i < nsrc; i += blockDim.x) { int v = list[i]; atomicOr(&bitmap[v >> 5], 1u << (v & 31)); }
    if (threadIdx.x == 0) { *s_count = nsrc; lvl_off[0] = 0; lvl_off[1] = nsrc; }
    __syncthreads();
    int lane = lane_id(), w = warp_id(), nw = blockDim.x >> 5;
    for (int lvl = 1; lvl <= depth; lvl++) {
        int fb = lvl_off[lvl - 1], fe = lvl_off[lvl];
        // frontier node fb + lane * nw + w (+ 32 * nw per sweep) belongs to lane `lane` of warp `w`; the rows of a
        // warp's up-to-32 nodes are flattened over its lanes (exclusive scan of the row lengths + shuffle search)
        for (int f0 = fb; f0 < fe; f0 += 32 * nw) {
            int fi = f0 + lane * nw + w;
            int b0 = 0, len0 = 0, b1 = 0, len1 = 0;
            if (fi < fe) {
                int u = list[fi];
                b0 = g.u_rowptr[u]; len0 = g.u_rowptr[u + 1] - b0;
                b1 = g.c_rowptr[u]; len1 = g.c_rowptr[u + 1] - b1;
                rowbytes += 16ull + 4ull * len0 + 8ull * len1;
            }
            int len = len0 + len1, incl = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(FULL_MASK, incl, o); if (lane >= o) incl += y; }
            int excl = incl - len, tot = __shfl_sync(FULL_MASK, incl, 31);
            if (lane == 0) edges += (unsigned long long)tot;
            for (int base = 0; base < tot; base += 32) {
                int f = base + lane, t = 0;
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) { int e = __shfl_sync(FULL_MASK, excl, t + s); if (e <= f) t += s; }
                int tb0 = __shfl_sync(FULL_MASK, b0, t), tl0 = __shfl_sync(FULL_MASK, len0, t);
                int tb1 = __shfl_sync(FULL_MASK, b1, t), tex = __shfl_sync(FULL_MASK, excl, t);
                int v = -1;
                if (f < tot) {
                    int j = f - tex;
                    if (j < tl0) { uint32_t ent = g.u_ent[tb0 + j]; if (ENT_MASK(ent)) v = ENT_COL(ent); }
                    else { int2 ce = g.c_ent[tb1 + (j - tl0)]; if (!CE_IS_DEAD(ce.y)) v = ce.x; }
                }
                bool isnew = false;
                if (v >= 0) {
                    unsigned bit = 1u << (v & 31);
                    unsigned old = atomicOr(&bitmap[v >> 5], bit);
                    isnew = !(old & bit);
                }
                unsigned m = __ballot_sync(FULL_MASK, isnew);
                if (m) {
                    int pos = 0;
                    if (lane == 0) pos = atomicAdd(s_count, __popc(m));
                    pos = __shfl_sync(FULL_MASK, pos, 0);
                    if (isnew) list[pos + __popc(m & ((1u << lane) - 1))] = v;
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) lvl_off[lvl + 1] = *s_count;
        __syncthreads();
    }
    return lvl_off[depth + 1];
}

// number of friendship entries shared test helper: does row `c` contain a friendship neighbour whose bit is
// set in `fbits` (bitmap of the initiator's friends)?  Warp-cooperative; all lanes return the result.
__device__ __forceinline__ bool warp_friend_intersect(const Graph &g, int c, const unsigned *fbits) {
    int b = g.u_rowptr[c], e = g.u_rowptr[c + 1];
    bool hit = false;
    for (int i = b + lane_id(); i < e; i += 32) {
        uint32_t ent = g.u_ent[i];
        if (ENT_MASK(ent) & MB_FRIEND) { int v = ENT_COL(ent); if (fbits[v >> 5] & (1u << (v & 31))) hit = true; }
    }
    return __any_sync(FULL_MASK, hit);
}

// Person.social_proximity (entities.py:674-689) given both packed attribute words and the friendship test.
// fp64, no contraction: ((((0 + ageTerm) + sex) + wealth) + edu) + friends, then / 5.
__device__ __forceinline__ double social_proximity(uint32_t self_attr, uint32_t tgt_attr, bool friends_shared) {
    double da = fabs((double)attr_age(tgt_attr) - (double)attr_age(self_attr));
    double total = 0.0;
    total = __dadd_rn(total, (da > 18.0) ? 0.0 : __dsub_rn(1.0, __ddiv_rn(da, 18.0)));
    total = __dadd_rn(total, (((self_attr ^ tgt_attr) & F_MALE) == 0) ? 1.0 : 0.0);
    total = __dadd_rn(total, (attr_wealth(self_attr) == attr_wealth(tgt_attr)) ? 1.0 : 0.0);
    total = __dadd_rn(total, (attr_edu(self_attr) == attr_edu(tgt_attr)) ? 1.0 : 0.0);
    total = __dadd_rn(total, friends_shared ? 1.0 : 0.0);
    return __ddiv_rn(total, 5.0);
}

// binary search helpers on sorted rows
__device__ __forceinline__ int crim_find(const Graph &g, int u, int v) {  // returns count or -1
    int lo = g.c_rowptr[u], hi = g.c_rowptr[u + 1], end = hi;
    while (lo < hi) { int m = (lo + hi) >> 1; if (g.c_ent[m].x < v) lo = m + 1; else hi = m; }
    if (lo < end) { int2 e = g.c_ent[lo]; if (e.x == v && !CE_IS_DEAD(e.y)) return CE_CNT(e.y); }
    return -1;
}
__device__ __forceinline__ int crim_find_idx(const Graph &g, int u, int v) {  // entry index (tombstones included) or -1
    int lo = g.c_rowptr[u], hi = g.c_rowptr[u + 1], end = hi;
    while (lo < hi) { int m = (lo + hi) >> 1; if (g.c_ent[m].x < v) lo = m + 1; else hi = m; }
    return (lo < end && g.c_ent[lo].x == v) ? lo : -1;
}
__device__ __forceinline__ int static_find_idx(const Graph &g, int u, int v) {  // entry index of v in u's row or -1
    int lo = g.u_rowptr[u], hi = g.u_rowptr[u + 1], end = hi;
    while (lo < hi) { int m = (lo + hi) >> 1; if (ENT_COL(g.u_ent[m]) < v) lo = m + 1; else hi = m; }
    return (lo < end && ENT_COL(g.u_ent[lo]) == v) ? lo : -1;
}
// ENT_ASYM of the entry u -> v (value `e`), from the rows as they are now
__device__ __forceinline__ bool ent_is_asym(const Graph &g, int u, uint32_t e) {
    if (e & ENT_CRIM) return false;   // weighed at its criminal entry, which every traversal takes from both sides
    const int ri = static_find_idx(g, ENT_COL(e), u);
    if (ri < 0) return true;
    const uint32_t re = g.u_ent[ri];
    return (re & ENT_CRIM) || __popc(ENT_MASK(re)) != __popc(ENT_MASK(e));
}
// CE_ASYM of the criminal entry u -> v (value y), from the rows as they are now
__device__ __forceinline__ bool crim_is_asym(const Graph &g, int u, int v, int y) {
    const int ri = crim_find_idx(g, v, u);
    if (ri < 0) return true;
    const int ry = g.c_ent[ri].y;
    return CE_IS_DEAD(ry) || CE_CNT(ry) + __popc(CE_SMASK(ry)) != CE_CNT(y) + __popc(CE_SMASK(y));
}
// flag entry idx of row u; a flagged entry in the lower part of the row makes the traversal read the whole row
__device__ __forceinline__ void ent_flag_asym(const Graph &g, int u, int idx) {
    uint32_t *ent = const_cast<uint32_t *>(g.u_ent);
    const uint32_t old = atomicOr(&ent[idx], ENT_ASYM);
    if (ENT_COL(old) < u) g.u_mid[u] = g.u_rowptr[u];
}
// after the masks of the pair (u, v) changed: set the flag on both entries if they no longer mirror each other.  Only
// ever sets (a stale flag costs a little time, a missing one would drop a meta edge).
__device__ __forceinline__ void ent_asym_update(const Graph &g, int u, int v) {
    const int i1 = static_find_idx(g, u, v), i2 = static_find_idx(g, v, u);
    if (i1 >= 0 && ent_is_asym(g, u, g.u_ent[i1])) ent_flag_asym(g, u, i1);
    if (i2 >= 0 && ent_is_asym(g, v, g.u_ent[i2])) ent_flag_asym(g, v, i2);
}
// u_mid of row a from the row as it is (one thread)
__device__ __forceinline__ int row_mid(const Graph &g, int a) {
    const int b = g.u_rowptr[a], e = g.u_rowptr[a + 1];
    int i = b;
    bool low_flag = false;
    for (; i < e; i++) { const uint32_t en = g.u_ent[i]; if (ENT_COL(en) > a) break; low_flag |= (en & ENT_ASYM) != 0; }
    return low_flag ? b : i;
}
__device__ __forceinline__ uint32_t static_find_mask(const Graph &g, int u, int v) {  // layer mask of v in u's row or 0
    int lo = g.u_rowptr[u], hi = g.u_rowptr[u + 1], end = hi;
    while (lo < hi) { int m = (lo + hi) >> 1; if (ENT_COL(g.u_ent[m]) < v) lo = m + 1; else hi = m; }
    if (lo < end) { uint32_t e = g.u_ent[lo]; if (ENT_COL(e) == v) return ENT_MASK(e); }
    return 0u;
}
