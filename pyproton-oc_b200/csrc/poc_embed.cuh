// poc_embed.cuh -- Person.oc_embeddedness on the device (A7).
//
// Reference: entities.py:526-558 + ProtonOC.update_meta_links, model.py:1515-1536, evaluated on the caller's
// own ball.  One GROUP of G threads per evaluation:
//   G = 256 (one group = one block)  small batches: lowest latency per evaluation
//   G = 32  (one group = one warp, eight evaluations per block, no block barrier at all)  large batches:
//           the evaluation is a chain of dependent gathers, so throughput comes from evaluations in flight
// Steps: radius-R ball by frontier BFS over the multiplex + criminal rows (visited bitmap in shared memory,
// ballot/popc compaction); dense local index = rank of the slot in the bitmap (per-word popcount prefix);
// one pass over the ball's rows appends the induced directed entries rank(u) | rank(v) << 12 | multiplicity << 24
// to a per-group edge list (L1/L2-resident scratch); SSSP by iterated relaxation with 64-bit atomicMin on the
// distance bit patterns, gated by "moved last round" bitmaps, until the fixed point Dijkstra reaches;
// sum(1/dist) over OC members / over the ball in 2^-34 fixed point (integer reductions: order independent).
#pragma once
#include "poc_device.cuh"

#define EMB_SCALE 17179869184.0 /* 2^34 */
#define DIST_INF 0x7ff0000000000000ull
#define RCAP 4096  /* ball size up to which ranks fit the 12-bit fields of a staged edge */
/* staged induced entries per evaluation live in per-block global scratch of C.ecap entries: 64 k (256 KB) up to 30,000
 * agents, 256 k above (the 300-member OC clique of the 100,000-agent configuration alone is 90 k entries) */
#define MAX_RADIUS 16

template <int G> struct EmbCfg;
template <> struct EmbCfg<256> { static constexpr int DCAP = 2048; };  // distances in shared memory up to this ball size
template <> struct EmbCfg<1024> { static constexpr int DCAP = 4096; };  // = RCAP: every staged ball keeps its distances in shared memory
template <> struct EmbCfg<512> { static constexpr int DCAP = 2048; };
template <> struct EmbCfg<128> { static constexpr int DCAP = 1024; };
template <> struct EmbCfg<64> { static constexpr int DCAP = 1024; };
template <> struct EmbCfg<32> { static constexpr int DCAP = 512; };

struct EmbShared {
    unsigned long long num, den;
    int count, ne, over, flag_oc, flag_more, carry;
    int lvl_off[MAX_RADIUS + 3];
    int ws[33];
    unsigned chg[2 * (RCAP >> 5)];
};

template <int G> __device__ __forceinline__ void gsync() {
    if (G == 32) __syncwarp(); else __syncthreads();
}

// exclusive scan of one int per thread across the group; total returned to every thread
template <int G> __device__ __forceinline__ int group_exscan(int v, int *ws, int &total) {
    int lane = threadIdx.x & 31, x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(FULL_MASK, x, o); if (lane >= o) x += y; }
    if (G == 32) { total = __shfl_sync(FULL_MASK, x, 31); return x - v; }
    int w = (threadIdx.x % G) >> 5;
    constexpr int NW = G / 32;
    if (lane == 31) ws[w] = x;
    __syncthreads();
    if (w == 0) {
        int s = (lane < NW) ? ws[lane] : 0, t = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(FULL_MASK, t, o); if (lane >= o) t += y; }
        ws[lane] = t - s;
        if (lane == 31) ws[32] = t;
    }
    __syncthreads();
    int res = ws[w] + x - v;
    total = ws[32];
    __syncthreads();
    return res;
}

// Frontier BFS for one group (same algorithm as bfs_levels in poc_device.cuh, group-relative).
template <int G>
__device__ int bfs_group(const Graph &g, int nwords, unsigned *bitmap, int *list, int depth, EmbShared *, int &s_count, int *s_lvl_off,
                         unsigned long long &edges, unsigned long long &rowbytes) {
    int gt = threadIdx.x % G, lane = threadIdx.x & 31, w = gt >> 5;
    constexpr int NW = G / 32;
    for (int i = gt; i < nwords; i += G) bitmap[i] = 0u;
    gsync<G>();
    if (gt == 0) { int v = list[0]; bitmap[v >> 5] = 1u << (v & 31); s_count = 1; s_lvl_off[0] = 0; s_lvl_off[1] = 1; }
    gsync<G>();
    for (int lvl = 1; lvl <= depth; lvl++) {
        int fb = s_lvl_off[lvl - 1], fe = s_lvl_off[lvl];
        for (int f0 = fb; f0 < fe; f0 += 32 * NW) {
            int fi = f0 + lane * NW + w;
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
                    if (lane == 0) pos = atomicAdd(&s_count, __popc(m));
                    pos = __shfl_sync(FULL_MASK, pos, 0);
                    if (isnew) list[pos + __popc(m & ((1u << lane) - 1))] = v;
                }
            }
        }
        gsync<G>();
        if (gt == 0) s_lvl_off[lvl + 1] = s_count;
        gsync<G>();
    }
    return s_lvl_off[depth + 1];
}

template <int G>
__global__ void __launch_bounds__((G == 32) ? 256 : G, (G == 32) ? 4 : (G == 1024 ? 2 : (G == 512 ? 4 : (G == 256 ? 8 : (G == 128 ? 16 : 16))))) k_embeddedness(DevCtx C, int parity) {
    // G == 32: eight warp-sized groups per 256-thread block; otherwise the block is the group
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int GPB = (G == 32) ? 8 : 1, DCAP = EmbCfg<G>::DCAP, NW = G / 32;
    if ((int)blockIdx.x * GPB >= C.n_emb[parity * C.R + C.r0 + blockIdx.y]) return;  // one block per request; the rest leave at once
    int r = C.r0 + blockIdx.y, n = C.n_agents[r], nwords = (n + 31) >> 5, wcap = (C.Ncap + 31) >> 5;
    int gid = threadIdx.x / G, gt = threadIdx.x % G, lane = threadIdx.x & 31, w = gt >> 5;
    // shared memory: [invw 256 f64] then per group [dist DCAP u64][EmbShared][bitmap wcap][wprefix wcap]
    double *invw = (double *)smem_raw;
    size_t per_group = ((size_t)DCAP * 8 + ((sizeof(EmbShared) + 7) & ~(size_t)7) + 2 * (((size_t)wcap * 4 + 7) & ~(size_t)7) + 15) & ~(size_t)15;
    unsigned char *gbase = smem_raw + 256 * 8 + (size_t)gid * per_group;
    unsigned long long *s_dist = (unsigned long long *)gbase;
    int4 *rowtab = (int4 *)gbase + w * 32;   // staging pass only: 32 rows x {entry address bases, split, rank} per warp
    EmbShared *S = (EmbShared *)(gbase + (size_t)DCAP * 8);
    unsigned *bitmap = (unsigned *)((unsigned char *)S + ((sizeof(EmbShared) + 7) & ~(size_t)7));
    int *wprefix = (int *)((unsigned char *)bitmap + (((size_t)wcap * 4 + 7) & ~(size_t)7));
    int slot = r * C.slot_stride + blockIdx.x * GPB + gid;   // (poc_device.cuh: one stride for every traversal kernel)
    int *list = C.scr_list + (size_t)slot * (C.Ncap + 1);
    unsigned long long *gdist = C.scr_dist + (size_t)slot * (C.Ncap + 1);
    unsigned *edgebuf = C.scr_edges + (size_t)slot * C.ecap;
    const poc_params &P = C.params[r];
    Graph g = graph_of(C, r);
    const uint32_t *attr = RP(C.attr, r, C.Ncap);
    double *embv = RP(C.emb_val, r, C.Ncap);
    const int32_t *reqs = C.emb_list + ((size_t)parity * C.R + r) * C.Ncap;
    int nreqs = C.n_emb[parity * C.R + r];
    int R = P.oc_embeddedness_radius;
    unsigned long long edges = 0, bfsbytes = 0, abytes = 0, anodes = 0;
    int nevals = 0, nreq = 0;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) invw[i] = (i > 0) ? 1.0 / (double)i : 0.0;
    __syncthreads();
#define RANK(v) (wprefix[(v) >> 5] + __popc(bitmap[(v) >> 5] & ((1u << ((v) & 31)) - 1)))
#define INBALL(v) (bitmap[(v) >> 5] & (1u << ((v) & 31)))
    for (int item = blockIdx.x * GPB + gid; item < nreqs; item += gridDim.x * GPB) {
        int self = reqs[item];
        if (gt == 0) { list[0] = self; S->flag_oc = 0; S->ne = 0; S->over = 0; }
        gsync<G>();
        bfsbytes = 0;
        unsigned long long ib = 0;  // algorithmic bytes of this evaluation
        int total = bfs_group<G>(g, nwords, bitmap, list, R, S, S->count, S->lvl_off, edges, bfsbytes);
        nreq++;
        bool anyoc = false;
        for (int i = 1 + gt; i < total; i += G) anyoc |= (attr[list[i]] & F_OC) != 0;
        if (anyoc) S->flag_oc = 1;
        gsync<G>();
        if (gt == 0) { ib += 4ull * total + 8ull; anodes += total; }  // OC flag of every ball node + the result
        if (!S->flag_oc) {  // no OC member in the ball: only the rows the BFS expanded were needed
            abytes += ib + bfsbytes;
            if (gt == 0) embv[self] = 0.0;
            gsync<G>();
            continue;
        }
        nevals++;
        // per-word exclusive popcount prefix -> dense local index rank(v), ascending slot order
        {
            int carry = 0;
            for (int base = 0; base < nwords; base += G) {
                int i = base + gt, v = (i < nwords) ? __popc(bitmap[i]) : 0, tot;
                int ex = group_exscan<G>(v, S->ws, tot);
                if (i < nwords) wprefix[i] = carry + ex;
                carry += tot;
            }
            gsync<G>();
        }
        bool staged = total <= RCAP;
        if (staged) {
            // one pass over the ball's rows: each warp takes 32 nodes (interleaved, so the OC clique spreads over
            // the warps), flattens their rows over its lanes, and appends the in-ball entries
            for (int i0 = 0; i0 < total; i0 += NW * 32) {
                int i = i0 + lane * NW + w;
                int b0 = 0, len0 = 0, b1 = 0, len1 = 0, ru = 0;
                if (i < total) {
                    int u = list[i];
                    b0 = g.u_rowptr[u]; len0 = g.u_rowptr[u + 1] - b0;
                    b1 = g.c_rowptr[u]; len1 = g.c_rowptr[u + 1] - b1;
                    ru = RANK(u);
                    ib += 16ull + 4ull * len0 + 8ull * len1;  // every ball row once
                }
                int len = len0 + len1, incl = len;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(FULL_MASK, incl, o); if (lane >= o) incl += y; }
                int excl = incl - len, tot = __shfl_sync(FULL_MASK, incl, 31);
                if (lane == 0) edges += (unsigned long long)tot;
                // flattened index f of node t: multiplex entry (b0 - excl) + f while f < excl + len0, else criminal
                // entry (b1 - excl - len0) + f
                int adr0 = b0 - excl, split = excl + len0, adr1 = b1 - split;
                // The non-empty rows of the batch go, compacted, into a per-warp table (in the distance array's shared
                // memory, which is idle until the relaxation).  A window of 32 flattened positions then finds its rows
                // from one OR-reduction of the row-start bits and a popcount instead of a shuffle binary search.
                unsigned nem = __ballot_sync(FULL_MASK, len > 0);
                if (len > 0) rowtab[__popc(nem & ((1u << lane) - 1))] = make_int4(adr0, split, adr1, ru);
                __syncwarp();
                int rows_before = 0;
                for (int fb = 0; fb < tot; fb += 32) {
                    int f = fb + lane;
                    unsigned d = (unsigned)(excl - fb);
                    unsigned starts = __reduce_or_sync(FULL_MASK, (len > 0 && d < 32u) ? (1u << d) : 0u);
                    int k = rows_before + __popc(starts & (0xffffffffu >> (31 - lane))) - 1;
                    rows_before += __popc(starts);
                    int4 rt = rowtab[f < tot ? k : 0];
                    int ta0 = rt.x, tsp = rt.y, ta1 = rt.z, tru = rt.w;
                    bool ok = false;
                    int v = 0, wgt = 0;
                    if (f < tot) {
                        if (f < tsp) {
                            uint32_t ent = g.u_ent[ta0 + f];
                            v = ENT_COL(ent); wgt = __popc(ENT_MASK(ent));
                            // a neighbour that also has a criminal entry is weighed there (mask copy + co-offences)
                            ok = wgt && !(ent & ENT_CRIM) && INBALL(v);
                        } else {
                            int2 ce = g.c_ent[ta1 + f];
                            v = ce.x; wgt = CE_CNT(ce.y) + __popc(CE_SMASK(ce.y));
                            ok = wgt > 0 && INBALL(v);  // 0: the reference divides by zero here
                            if (ok && wgt > 255) { S->over = 1; ok = false; }
                        }
                    }
                    unsigned m = __ballot_sync(FULL_MASK, ok);
                    if (m) {
                        int basep = 0;
                        if (lane == 0) basep = atomicAdd(&S->ne, __popc(m));
                        basep = __shfl_sync(FULL_MASK, basep, 0);
                        if (ok) {
                            int pos = basep + __popc(m & ((1u << lane) - 1));
                            if (pos < C.ecap) edgebuf[pos] = (unsigned)tru | ((unsigned)RANK(v) << 12) | ((unsigned)wgt << 24);
                            else S->over = 1;
                        }
                    }
                }
                __syncwarp();   // the next batch overwrites the row table
            }
            gsync<G>();
            if (S->over) { staged = false; ib = (gt == 0) ? 4ull * total + 8ull : 0ull; }
        }
        unsigned long long *dist = (total <= DCAP) ? s_dist : gdist;
        for (int i = gt; i < total; i += G) dist[i] = DIST_INF;
        for (int i = gt; i < 2 * (RCAP >> 5); i += G) S->chg[i] = 0u;
        gsync<G>();
        if (gt == 0) { int rs = RANK(self); dist[rs] = 0ull; S->flag_more = staged ? 0 : 1; if (rs < RCAP) S->chg[rs >> 5] = 1u << (rs & 31); }
        gsync<G>();
        int rounds = 0;
        if (staged) {
            // An entry can only improve something if one of its end points moved in the previous round: two
            // bitmaps (moved last round / this round) gate the relaxations.
            int nedge = S->ne, cw = (total + 31) >> 5;
            unsigned *chg0 = S->chg, *chg1 = S->chg + (RCAP >> 5);
            // two barriers per round: a round that moved anything writes its number to flag_more, and everybody goes
            // on while flag_more has reached the round just finished (a faster thread can only have raised it)
            for (int round = 1;; round++) {
                bool changed = false;
                for (int e = gt; e < nedge; e += G) {
                    unsigned ed = edgebuf[e];
                    int ru = ed & 0xfffu, rv = (ed >> 12) & 0xfffu;
                    // moved last round, or already this round (racy read: can only let more entries through)
                    bool cu = (chg0[ru >> 5] | chg1[ru >> 5]) & (1u << (ru & 31)), cv = (chg0[rv >> 5] | chg1[rv >> 5]) & (1u << (rv & 31));
                    if (!(cu | cv)) continue;
                    double wt = invw[ed >> 24];
                    double du = __longlong_as_double((long long)dist[ru]), dv = __longlong_as_double((long long)dist[rv]);
                    // undirected meta edge (CANON: each direction relaxes with its own weight = max multiplicity)
                    double nv = du + wt, nu = dv + wt;
                    if (cu && nv < dv) {
                        atomicMin(&dist[rv], (unsigned long long)__double_as_longlong(nv));
                        atomicOr(&chg1[rv >> 5], 1u << (rv & 31)); changed = true;
                    }
                    if (cv && nu < du) {
                        atomicMin(&dist[ru], (unsigned long long)__double_as_longlong(nu));
                        atomicOr(&chg1[ru >> 5], 1u << (ru & 31)); changed = true;
                    }
                }
                if (changed) S->flag_more = round;
                gsync<G>();
                for (int i = gt; i < cw; i += G) chg0[i] = 0u;  // becomes next round's "this round" map
                unsigned *tmp = chg0; chg0 = chg1; chg1 = tmp;
                bool more = S->flag_more >= round;
                rounds = round;
                gsync<G>();
                if (!more || round > total + 2) break;
            }
        } else {
            // very large ball: relax straight from the rows in global memory, one warp per node per round
            while (S->flag_more) {
                gsync<G>();
                if (gt == 0) S->flag_more = 0;
                gsync<G>();
                bool changed = false;
                for (int i = w; i < total; i += NW) {
                    int u = list[i], ru = RANK(u);
                    int b0 = g.u_rowptr[u], e0 = g.u_rowptr[u + 1], b1 = g.c_rowptr[u], e1 = g.c_rowptr[u + 1];
                    int len0 = e0 - b0, len = len0 + (e1 - b1);
                    for (int j = lane; j < len; j += 32) {
                        int v, wgt;
                        if (j < len0) {
                            uint32_t ent = g.u_ent[b0 + j];
                            v = ENT_COL(ent); wgt = __popc(ENT_MASK(ent));
                            if (wgt == 0 || (ent & ENT_CRIM)) continue;  // cleared by an arrest (Q11) / weighed below
                            if (!INBALL(v)) continue;
                        } else {
                            int2 ce = g.c_ent[b1 + (j - len0)];
                            v = ce.x; wgt = CE_CNT(ce.y) + __popc(CE_SMASK(ce.y));
                            if (wgt <= 0 || !INBALL(v)) continue;       // the reference divides by zero at 0
                        }
                        int rv = RANK(v);
                        double wt = 1.0 / (double)wgt;
                        double du = __longlong_as_double((long long)dist[ru]), dv = __longlong_as_double((long long)dist[rv]);
                        double nv = du + wt, nu = dv + wt;
                        if (nv < dv) { atomicMin(&dist[rv], (unsigned long long)__double_as_longlong(nv)); changed = true; }
                        if (nu < du) { atomicMin(&dist[ru], (unsigned long long)__double_as_longlong(nu)); changed = true; }
                    }
                    if (lane == 0) { edges += (unsigned long long)len; if (rounds == 0) ib += 16ull + 4ull * len0 + 8ull * (len - len0); }
                }
                if (changed) S->flag_more = 1;
                gsync<G>();
                if (++rounds > total + 2) break;  // cannot happen with positive weights
            }
        }
        if (gt == 0) { S->num = 0ull; S->den = 0ull; }
        gsync<G>();
        unsigned long long num = 0, den = 0;
        for (int i = 1 + gt; i < total; i += G) {
            int v = list[i];
            double dv = __longlong_as_double((long long)dist[RANK(v)]);
            long long q = __double2ll_rn((1.0 / dv) * EMB_SCALE);
            den += (unsigned long long)q;
            if (attr[v] & F_OC) num += (unsigned long long)q;
        }
        num = warp_sum_u64(num); den = warp_sum_u64(den);
        if (G == 32) { if (lane == 0) { S->num = num; S->den = den; } }
        else if (lane == 0) { atomicAdd(&S->num, num); atomicAdd(&S->den, den); }
        gsync<G>();
        if (gt == 0) embv[self] = (double)(long long)S->num / (double)(long long)S->den;
        abytes += ib;
        gsync<G>();
    }
#undef RANK
#undef INBALL
    edges = warp_sum_u64(edges); abytes = warp_sum_u64(abytes); anodes = warp_sum_u64(anodes);
    unsigned long long *stq = RP(C.stats, r, ST_COUNT);
    if (lane == 0) {
        if (edges) atomicAdd(&C.edges[r], edges);
        if (abytes) atomicAdd(&stq[ST_EMB_BYTES], abytes);
        if (anodes) atomicAdd(&stq[ST_EMB_NODES], anodes);
    }
    if (gt == 0 && nreq) { atomicAdd(&stq[ST_EMB_REQUESTS], (unsigned long long)nreq); atomicAdd(&stq[ST_EMB_FULL], (unsigned long long)nevals); }
    if (gt == 0 && nevals) atomicAdd(&RP(C.counters, r, CN_COUNT)[CN_TICK_EMB], nevals);
}

static inline size_t emb_smem_bytes(int G, int Ncap) {
    size_t wcap = (size_t)(Ncap + 31) / 32;
    size_t dcap = (G == 32) ? EmbCfg<32>::DCAP : (G == 1024 ? EmbCfg<1024>::DCAP : (G >= 256 ? EmbCfg<256>::DCAP : EmbCfg<128>::DCAP));
    size_t per_group = (dcap * 8 + ((sizeof(EmbShared) + 7) & ~(size_t)7) + 2 * ((wcap * 4 + 7) & ~(size_t)7) + 15) & ~(size_t)15;
    return 256 * 8 + (size_t)((G == 32) ? 8 : 1) * per_group;
}
