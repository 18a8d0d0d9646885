// poc_embed2.cuh -- Person.oc_embeddedness on the device (A7), second formulation.
//
// Reference: entities.py:526-558 + ProtonOC.update_meta_links, model.py:1515-1536, evaluated on the caller's own
// ball.  One block of G threads per evaluation.  What changed against poc_embed.cuh (kept for A/B
// runs, POC_EMB_V=1) and why (profiles/r02_*: the first formulation issued ~69 k warp instructions per evaluation):
//   * the staging pass walks the ball's rows LEVEL BY LEVEL (self, ring 1, ring 2, ...: the BFS list is already in
//     that order) and relaxes every induced entry the moment it is staged, so when the pass ends the distances
//     have already travelled outwards once: the first ~3 Jacobi rounds of the old scheme disappear.  What is left
//     is a short tail of plain sweeps over the staged entries until one of them changes nothing (= the least
//     fixed point = Dijkstra's distances; the same float64 sums along the same paths, so the result is
//     bit-identical to poc_embed.cuh and to oracle/hotpath.c).
//   * everything the hot loops touch is in SHARED memory at compile-time offsets: {bitmap word, rank prefix} pairs
//     (one 8-byte load answers "in the ball?" and "which rank?"), distances, the staged entries, the 1/w table.
//     The staged entries are per-warp regions (a warp appends with a register counter: no atomics, no broadcast)
//     and each warp later sweeps its own region; only the tail of an unusually large ball spills to the per-block
//     global scratch.  The first formulation wrote 157 MB of staging scratch per launch to DRAM and paid for
//     generic addressing (its distance array could be global) in every access.
//   * small frontiers are walked by as many warps as they can feed (ring 0 is one row), not by all of them.
// Distances are float64 bit patterns (positive doubles order like unsigned integers): atomicMin in shared memory.
#pragma once
#include "poc_device.cuh"
#include "poc_embed.cuh"

struct Emb2Shared {
    unsigned long long num, den;
    unsigned long long st_bytes, st_edges, st_nodes;   // algorithmic-work counters of the block (flushed once at the end)
    long long ph[8];                                    // phase cycle counters (thread 0)
    long long t0;
    unsigned ev_bytes, ev_edges;                        // ... of the evaluation in hand
    int count, over, flag_oc, nreq, nevals, any;
    int lvl_off[MAX_RADIUS + 3];
    int ws[33];
};

// per-evaluation phase cycle counters (thread 0 of the block): stats[ST_PH0 + k]
enum { PH_BFS = 0, PH_PREFIX, PH_STAGE, PH_SWEEP, PH_REDUCE, PH_NSWEEPS, PH_SPILL, PH_FALLBACK };

template <int G> struct Emb2Layout {
    static constexpr int NW = G / 32;
    static constexpr int OFF_ROWTAB = (int)((sizeof(Emb2Shared) + 15) & ~(size_t)15);
    static constexpr int OFF_INVW = OFF_ROWTAB + NW * 32 * (int)sizeof(int4);
    static constexpr int OFF_BP = OFF_INVW + 256 * (int)sizeof(double);
};
static inline size_t emb2_fixed_bytes(int G, int Ncap) {
    size_t wcap = (size_t)(Ncap + 31) / 32;
    size_t off_bp = ((sizeof(Emb2Shared) + 15) & ~(size_t)15) + (size_t)(G / 32) * 32 * sizeof(int4) + 256 * sizeof(double);
    return off_bp + ((wcap * 8 + 15) & ~(size_t)15);
}

// Shared-memory accesses of the window loops through explicit 32-bit shared-space addresses.  Written out because the
// compiler otherwise re-derives the block's shared window base and every run-time offset (S2R / S2UR / ULEA chains,
// 20-30 instructions per window in profiles/r02_*) at each access instead of keeping one address register.
__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ int4 lds_i4(unsigned a) { int4 v; asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ uint2 lds_u2(unsigned a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ unsigned lds_u32(unsigned a) { unsigned v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ double lds_f64(unsigned a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_u32(unsigned a, unsigned v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_f64(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void sts_u64(unsigned a, unsigned long long v) { asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ void sts_i4(unsigned a, int4 v) { asm volatile("st.shared.v4.s32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
__device__ __forceinline__ void smin_f64(unsigned a, double v) {   // positive doubles order like their bit patterns
    asm volatile("red.shared.min.u64 [%0], %1;" ::"r"(a), "l"((unsigned long long)__double_as_longlong(v)) : "memory");
}
#define OPAQUE(x) asm volatile("" : "+r"(x))

// Rows of list[lb..le) flattened over the lanes of the participating warps.  BODY(f-th flattened entry) is written
// inline twice below (BFS and staging differ in what they do with an entry), so this is only the shared prologue:
// row bounds -> exclusive scan -> compacted row table {multiplex base - excl, split, criminal base - split, payload}.
#define EMB2_ROWS_PROLOGUE(PAYLOAD)                                                                                   \
    int b0 = 0, len0 = 0, b1 = 0, len1 = 0, pay = 0;                                                                   \
    if (i < le) {                                                                                                      \
        int u = list[i];                                                                                               \
        b0 = g.u_rowptr[u]; len0 = g.u_rowptr[u + 1] - b0;                                                             \
        b1 = g.c_rowptr[u]; len1 = g.c_rowptr[u + 1] - b1;                                                             \
        pay = PAYLOAD;                                                                                                 \
    }                                                                                                                  \
    int len = len0 + len1, incl = len;                                                                                 \
    _Pragma("unroll") for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(FULL_MASK, incl, o); if (lane >= o) incl += y; } \
    const int excl = incl - len, tot = __shfl_sync(FULL_MASK, incl, 31);                                               \
    if (tot == 0) continue;                                                                                            \
    {   /* algorithmic bytes of these rows: 16 B of row pointers + 4 B per multiplex entry + 8 B per criminal entry */  \
        const unsigned c1 = __reduce_add_sync(FULL_MASK, (unsigned)len1), nr = __popc(__ballot_sync(FULL_MASK, i < le)); \
        if (lane == 0) { atomicAdd(&S->ev_edges, (unsigned)tot); atomicAdd(&S->ev_bytes, 16u * nr + 4u * (unsigned)tot + 4u * c1); } \
    }                                                                                                                  \
    {                                                                                                                  \
        unsigned nem = __ballot_sync(FULL_MASK, len > 0);                                                              \
        if (len > 0) sts_i4(a_rowtab + 16u * __popc(nem & ltmask), make_int4(b0 - excl, excl + len0, b1 - excl - len0, pay)); \
    }                                                                                                                  \
    __syncwarp();

template <int G, int MB>
__global__ void __launch_bounds__(G, MB) k_emb2(DevCtx C, int parity, int smem_bytes) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typedef Emb2Layout<G> L;
    constexpr int NW = L::NW;
    if ((int)blockIdx.x >= C.n_emb[parity * C.R + C.r0 + blockIdx.y]) return;  // one block per request; the rest leave at once
    const int r = C.r0 + blockIdx.y, n = C.n_agents[r], nwords = (n + 31) >> 5, wcap = (C.Ncap + 31) >> 5;
    const int gt = threadIdx.x, w = gt >> 5;
    int lane = threadIdx.x & 31;
    unsigned ltmask = (1u << lane) - 1u;
    OPAQUE(lane); OPAQUE(ltmask);
    Emb2Shared *S = (Emb2Shared *)smem_raw;
    double *invw = (double *)(smem_raw + L::OFF_INVW);
    uint2 *bp = (uint2 *)(smem_raw + L::OFF_BP);          // {bitmap word, ball members in the words before it}
    unsigned a_base = smem_addr(smem_raw);
    OPAQUE(a_base);
    unsigned a_rowtab = a_base + L::OFF_ROWTAB + w * 512;  // this warp's 32-row slab
    OPAQUE(a_rowtab);
    const unsigned a_bp = a_base + L::OFF_BP, a_invw = a_base + L::OFF_INVW;
    int off_dyn = L::OFF_BP + ((wcap * 8 + 15) & ~15);
    OPAQUE(off_dyn);
    const int dyn_bytes = smem_bytes - off_dyn;
    const int slot = r * C.slot_stride + blockIdx.x;   // (poc_device.cuh: one stride for every traversal kernel)
    int *list = C.scr_list + (size_t)slot * (C.Ncap + 1);
    const poc_params &P = C.params[r];
    Graph g = graph_of(C, r);
    // opaque copies: under the register cap the compiler otherwise re-derives these 64-bit bases from the kernel
    // parameters (eight instructions) at every use inside the window loops
    asm volatile("" : "+l"(g.u_ent)); asm volatile("" : "+l"(g.c_ent));
    const uint32_t *attr = RP(C.attr, r, C.Ncap);
    const int32_t *reqs = C.emb_list + ((size_t)parity * C.R + r) * C.Ncap;
    const int nreqs = C.n_emb[parity * C.R + r];
    const int R = P.oc_embeddedness_radius;
    if (gt < 8) S->ph[gt] = 0;
    if (gt == 8) { S->st_bytes = 0ull; S->st_edges = 0ull; S->st_nodes = 0ull; S->nreq = 0; S->nevals = 0; }
    for (int i = gt; i < 256; i += G) invw[i] = C.invw[i];
#define PHASE(k) { long long t1 = clock64(); S->ph[k] += t1 - S->t0; S->t0 = t1; }
    for (int item = blockIdx.x; item < nreqs; item += gridDim.x) {
        const int self = reqs[item];
        // ---- radius-R ball: frontier BFS over the multiplex + criminal rows, visited bits in bp[].x
        for (int i = gt; i < nwords; i += G) bp[i] = make_uint2(0u, 0u);
        if (gt == 0) {
            list[0] = self; S->flag_oc = 0; S->over = 0; S->nreq++; S->count = 1; S->lvl_off[0] = 0; S->lvl_off[1] = 1;
            S->t0 = clock64(); S->ev_bytes = 0u; S->ev_edges = 0u;
        }
        __syncthreads();
        if (gt == 0) bp[self >> 5].x = 1u << (self & 31);
        __syncthreads();
        for (int lvl = 1; lvl <= R; lvl++) {
            const int lb = S->lvl_off[lvl - 1], le = S->lvl_off[lvl];
            const int nwp = min(NW, (le - lb + 7) >> 3);   // warps that get at least ~8 rows each
            if (w < nwp) {
                for (int i0 = lb; i0 < le; i0 += nwp * 32) {
                    const int i = i0 + lane * nwp + w;
                    EMB2_ROWS_PROLOGUE(0)
                    int rows_before = 0;
                    for (int fb = 0; fb < tot; fb += 32) {
                        const int f = fb + lane;
                        const unsigned d = (unsigned)(excl - fb);
                        const unsigned starts = __reduce_or_sync(FULL_MASK, (len > 0 && d < 32u) ? (1u << d) : 0u);
                        const int k = rows_before + __popc(starts & (0xffffffffu >> (31 - lane))) - 1;
                        rows_before += __popc(starts);
                        int v = -1;
                        if (f < tot) {
                            const int4 rt = lds_i4(a_rowtab + 16u * k);
                            if (f < rt.y) { uint32_t ent = __ldg(&g.u_ent[rt.x + f]); if (ENT_MASK(ent)) v = ENT_COL(ent); }
                            else { const int2 ce = __ldg(&g.c_ent[rt.z + f]); if (!CE_IS_DEAD(ce.y)) v = ce.x; }
                        }
                        bool isnew = false;
                        if (v >= 0) {
                            const unsigned bit = 1u << (v & 31);
                            isnew = !(atomicOr(&bp[v >> 5].x, bit) & bit);
                        }
                        const unsigned m = __ballot_sync(FULL_MASK, isnew);
                        if (m) {
                            int pos = 0;
                            if (lane == 0) pos = atomicAdd(&S->count, __popc(m));
                            pos = __shfl_sync(FULL_MASK, pos, 0);
                            if (isnew) list[pos + __popc(m & ltmask)] = v;
                        }
                    }
                    __syncwarp();   // the next batch overwrites the row table
                }
            }
            __syncthreads();
            if (gt == 0) S->lvl_off[lvl + 1] = S->count;
            __syncthreads();
        }
        const int total = S->lvl_off[R + 1];
        {
            bool anyoc = false;
            for (int i = 1 + gt; i < total; i += G) anyoc |= (attr[list[i]] & F_OC) != 0;
            if (anyoc) S->flag_oc = 1;
        }
        __syncthreads();
        if (gt == 0) { S->st_nodes += total; PHASE(PH_BFS) }
        if (!S->flag_oc) {  // no OC member in the ball: only the rows the BFS expanded were needed (+ OC flags, result)
            if (gt == 0) {
                S->st_bytes += S->ev_bytes + 4ull * total + 8ull; S->st_edges += S->ev_edges;
                RP(C.emb_val, r, C.Ncap)[self] = 0.0;
            }
            __syncthreads();
            continue;
        }
        if (gt == 0) { S->st_edges += S->ev_edges; S->ev_edges = 0u; S->ev_bytes = 4u * total + 8u; }  // the staging pass counts every ball row once
        // ---- per-word exclusive popcount prefix -> dense local index rank(v), ascending slot order.  Each warp scans a
        // contiguous chunk of words; the chunk totals are combined through shared memory.
        {
            const int chunk = (((nwords + NW - 1) / NW) + 31) & ~31;
            const int wb = w * chunk, we = min(nwords, wb + chunk);
            int run = 0;
            for (int base = wb; base < we; base += 32) {
                const int i = base + lane;
                int c = (i < we) ? __popc(bp[i].x) : 0, x = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(FULL_MASK, x, o); if (lane >= o) x += y; }
                if (i < we) bp[i].y = (unsigned)(run + x - c);
                run += __shfl_sync(FULL_MASK, x, 31);
            }
            if (lane == 0) S->ws[w] = run;
            __syncthreads();
            int carry = 0;
            for (int q = 0; q < w; q++) carry += S->ws[q];
            if (carry) for (int i = wb + lane; i < we; i += 32) bp[i].y += (unsigned)carry;
        }
        __syncthreads();
        if (gt == 0) { S->nevals++; PHASE(PH_PREFIX) }
#define RANK(v) ((int)(bp[(v) >> 5].y + __popc(bp[(v) >> 5].x & ((1u << ((v) & 31)) - 1u))))
        // ---- split of the dynamic region: distances of the ball first, one region of staged entries per warp behind them
        const int dist_bytes = (total * 8 + 15) & ~15;
        bool staged = total <= RCAP && dist_bytes + NW * 256 <= dyn_bytes;
        // a staged entry is rank(u) | rank(v) << rb | multiplicity << 2 rb: balls of up to 1,024 nodes (nearly all at 10,000
        // agents) leave 12 bits for the multiplicity -- co-offence counts of a long-lived OC pair pass 255 late in a run --
        // larger ones 8
        const int rb = (total <= 1024) ? 10 : 12;
        const unsigned rmask = (1u << rb) - 1u;
        const int wmax = (1 << (32 - 2 * rb)) - 1;
        int rounds = 0;
        if (staged) {
            unsigned long long *dist = (unsigned long long *)(smem_raw + off_dyn);
            int ecap_w = ((dyn_bytes - dist_bytes) >> 2) / NW;          // staged entries per warp in shared memory
            unsigned a_dist = a_base + off_dyn, a_sedge = a_dist + dist_bytes + 4u * (w * ecap_w);
            OPAQUE(ecap_w); OPAQUE(a_dist); OPAQUE(a_sedge);
            const int gcap_w = C.ecap / NW;
            unsigned *gedge = C.scr_edges + (size_t)slot * C.ecap + (size_t)w * gcap_w;
            for (int i = gt; i < total; i += G) dist[i] = DIST_INF;
            __syncthreads();
            if (gt == 0) dist[RANK(self)] = 0ull;
            __syncthreads();
            int wcnt = 0;   // entries this warp has staged (warp-uniform)
            // Level by level (list[lvl_off[k] .. lvl_off[k+1]) is ring k, ring 0 = self): each participating warp takes 32
            // rows (interleaved, so the OC clique spreads over the warps), flattens them over its lanes, appends the in-ball
            // entries to its own region and relaxes each of them both ways on the spot.
            for (int lvl = 0; lvl <= R; lvl++) {
                const int lb = S->lvl_off[lvl], le = S->lvl_off[lvl + 1];
                const int nwp = min(NW, (le - lb + 7) >> 3);
                if (w < nwp) {
                    for (int i0 = lb; i0 < le; i0 += nwp * 32) {
                        const int i = i0 + lane * nwp + w;
                        EMB2_ROWS_PROLOGUE(RANK(u))
                        int rows_before = 0;
                        for (int fb = 0; fb < tot; fb += 32) {
                            const int f = fb + lane;
                            const unsigned d = (unsigned)(excl - fb);
                            const unsigned starts = __reduce_or_sync(FULL_MASK, (len > 0 && d < 32u) ? (1u << d) : 0u);
                            const int k = rows_before + __popc(starts & (0xffffffffu >> (31 - lane))) - 1;
                            rows_before += __popc(starts);
                            bool ok = false;
                            int v = 0, wgt = 0, tru = 0;
                            uint2 bw = make_uint2(0u, 0u);
                            if (f < tot) {
                                const int4 rt = lds_i4(a_rowtab + 16u * k);
                                tru = rt.w;
                                if (f < rt.y) {
                                    const uint32_t ent = __ldg(&g.u_ent[rt.x + f]);
                                    v = ENT_COL(ent); wgt = __popc(ENT_MASK(ent));
                                    bw = lds_u2(a_bp + 8u * (v >> 5));
                                    // a neighbour that also has a criminal entry is weighed there (mask copy + co-offences)
                                    ok = wgt && !(ent & ENT_CRIM) && ((bw.x >> (v & 31)) & 1u);
                                } else {
                                    const int2 ce = __ldg(&g.c_ent[rt.z + f]);
                                    v = ce.x; wgt = CE_CNT(ce.y) + __popc(CE_SMASK(ce.y));
                                    bw = lds_u2(a_bp + 8u * (v >> 5));
                                    ok = wgt > 0 && ((bw.x >> (v & 31)) & 1u);  // 0: the reference divides by zero here
                                    if (ok && wgt > wmax) { S->over = 1; ok = false; }
                                }
                            }
                            const unsigned m = __ballot_sync(FULL_MASK, ok);
                            if (ok) {
                                const int rv = (int)bw.y + __popc(bw.x & ((1u << (v & 31)) - 1u));
                                const int pos = wcnt + __popc(m & ltmask);
                                const unsigned pk = (unsigned)tru | ((unsigned)rv << rb) | ((unsigned)wgt << (2 * rb));
                                if (pos < ecap_w) sts_u32(a_sedge + 4u * pos, pk);
                                else if (pos - ecap_w < gcap_w) gedge[pos - ecap_w] = pk;
                                else S->over = 1;
                                const double wt = (wgt < 256) ? lds_f64(a_invw + 8u * wgt) : 1.0 / (double)wgt;
                                const double du = lds_f64(a_dist + 8u * tru), dv = lds_f64(a_dist + 8u * rv);
                                const double nv = du + wt, nu = dv + wt;   // undirected meta edge (CANON: each direction relaxes with its own weight)
                                if (nv < dv) smin_f64(a_dist + 8u * rv, nv);
                                if (nu < du) smin_f64(a_dist + 8u * tru, nu);
                            }
                            wcnt += __popc(m);
                        }
                        __syncwarp();   // the next batch overwrites the row table
                    }
                }
                __syncthreads();
            }
            if (S->over) staged = false;
            if (gt == 0) PHASE(PH_STAGE)
            if (staged) {
                const int ns = min(wcnt, ecap_w);
                if (lane == 0 && wcnt > ecap_w) S->ph[PH_SPILL]++;
                for (;;) {   // plain sweeps (every warp over the entries it staged) until one changes nothing
                    bool changed = false;
                    for (int e = lane; e < wcnt; e += 32) {
                        const unsigned pk = (e < ns) ? lds_u32(a_sedge + 4u * e) : gedge[e - ns];
                        const int ru = pk & rmask, rv = (pk >> rb) & rmask, wg = pk >> (2 * rb);
                        const double wt = (wg < 256) ? lds_f64(a_invw + 8u * wg) : 1.0 / (double)wg;
                        const double du = lds_f64(a_dist + 8u * ru), dv = lds_f64(a_dist + 8u * rv);
                        const double nv = du + wt, nu = dv + wt;
                        if (nv < dv) { smin_f64(a_dist + 8u * rv, nv); changed = true; }
                        if (nu < du) { smin_f64(a_dist + 8u * ru, nu); changed = true; }
                    }
                    rounds++;
                    if (!__syncthreads_or(changed) || rounds > total + 2) break;
                }
                if (gt == 0) { PHASE(PH_SWEEP) S->ph[PH_NSWEEPS] += rounds; S->num = 0ull; S->den = 0ull; }
                __syncthreads();
                unsigned long long num = 0, den = 0;
                for (int i = 1 + gt; i < total; i += G) {
                    const int v = list[i];
                    const double dv = __longlong_as_double((long long)dist[RANK(v)]);
                    const long long q = __double2ll_rn((1.0 / dv) * EMB_SCALE);
                    den += (unsigned long long)q;
                    if (attr[v] & F_OC) num += (unsigned long long)q;
                }
                num = warp_sum_u64(num); den = warp_sum_u64(den);
                if (lane == 0 && den) { atomicAdd(&S->num, num); atomicAdd(&S->den, den); }
            }
        }
        if (!staged) {
            // very large ball (or a multiplicity beyond the 8-bit field): distances in global scratch, relaxation straight
            // from the rows, one warp per node per round
            unsigned long long *gdist = C.scr_dist + (size_t)slot * (C.Ncap + 1);
            if (gt == 0) { S->ph[PH_FALLBACK]++; S->ev_bytes = 4u * total + 8u; }
            for (int i = gt; i < total; i += G) gdist[i] = DIST_INF;
            __syncthreads();
            if (gt == 0) gdist[RANK(self)] = 0ull;
            __syncthreads();
            for (;;) {
                bool changed = false;
                for (int i = w; i < total; i += NW) {
                    const int u = list[i], ru = RANK(u);
                    const int b0 = g.u_rowptr[u], e0 = g.u_rowptr[u + 1], b1 = g.c_rowptr[u], e1 = g.c_rowptr[u + 1];
                    const int len0 = e0 - b0, len = len0 + (e1 - b1);
                    for (int j = lane; j < len; j += 32) {
                        int v, wgt;
                        uint2 bw;
                        if (j < len0) {
                            const uint32_t ent = g.u_ent[b0 + j];
                            v = ENT_COL(ent); wgt = __popc(ENT_MASK(ent));
                            if (wgt == 0 || (ent & ENT_CRIM)) continue;  // cleared by an arrest (Q11) / weighed below
                            bw = bp[v >> 5];
                            if (!((bw.x >> (v & 31)) & 1u)) continue;
                        } else {
                            const int2 ce = g.c_ent[b1 + (j - len0)];
                            v = ce.x; wgt = CE_CNT(ce.y) + __popc(CE_SMASK(ce.y));
                            bw = bp[v >> 5];
                            if (wgt <= 0 || !((bw.x >> (v & 31)) & 1u)) continue;       // the reference divides by zero at 0
                        }
                        const int rv = (int)bw.y + __popc(bw.x & ((1u << (v & 31)) - 1u));
                        const double wt = 1.0 / (double)wgt;
                        const double du = __longlong_as_double((long long)gdist[ru]), dv = __longlong_as_double((long long)gdist[rv]);
                        const double nv = du + wt, nu = dv + wt;
                        if (nv < dv) { atomicMin(&gdist[rv], (unsigned long long)__double_as_longlong(nv)); changed = true; }
                        if (nu < du) { atomicMin(&gdist[ru], (unsigned long long)__double_as_longlong(nu)); changed = true; }
                    }
                    if (lane == 0) { atomicAdd(&S->ev_edges, (unsigned)len); if (rounds == 0) atomicAdd(&S->ev_bytes, 16u + 4u * len0 + 8u * (len - len0)); }
                }
                rounds++;
                if (!__syncthreads_or(changed) || rounds > total + 2) break;  // the bound cannot be hit with positive weights
            }
            if (gt == 0) { PHASE(PH_SWEEP) S->ph[PH_NSWEEPS] += rounds; S->num = 0ull; S->den = 0ull; }
            __syncthreads();
            unsigned long long num = 0, den = 0;
            for (int i = 1 + gt; i < total; i += G) {
                const int v = list[i];
                const double dv = __longlong_as_double((long long)gdist[RANK(v)]);
                const long long q = __double2ll_rn((1.0 / dv) * EMB_SCALE);
                den += (unsigned long long)q;
                if (attr[v] & F_OC) num += (unsigned long long)q;
            }
            num = warp_sum_u64(num); den = warp_sum_u64(den);
            if (lane == 0 && den) { atomicAdd(&S->num, num); atomicAdd(&S->den, den); }
        }
        __syncthreads();
        if (gt == 0) {
            RP(C.emb_val, r, C.Ncap)[self] = (double)(long long)S->num / (double)(long long)S->den;
            S->st_bytes += S->ev_bytes; S->st_edges += S->ev_edges;
            PHASE(PH_REDUCE)
        }
        __syncthreads();
    }
#undef RANK
#undef PHASE
    if (gt == 0) {
        unsigned long long *stq = RP(C.stats, r, ST_COUNT);
        if (S->st_edges) atomicAdd(&C.edges[r], S->st_edges);
        if (S->st_bytes) atomicAdd(&stq[ST_EMB_BYTES], S->st_bytes);
        if (S->st_nodes) atomicAdd(&stq[ST_EMB_NODES], S->st_nodes);
        atomicAdd(&stq[ST_EMB_REQUESTS], (unsigned long long)S->nreq); atomicAdd(&stq[ST_EMB_FULL], (unsigned long long)S->nevals);
        for (int k = 0; k < 8; k++) if (S->ph[k]) atomicAdd(&stq[ST_PH0 + k], (unsigned long long)S->ph[k]);
        if (S->nevals) atomicAdd(&RP(C.counters, r, CN_COUNT)[CN_TICK_EMB], S->nevals);
    }
}
