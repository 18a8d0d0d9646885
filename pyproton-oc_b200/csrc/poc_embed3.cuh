// poc_embed3.cuh -- Person.oc_embeddedness on the device (A7), third formulation (the default).
//
// Reference: entities.py:526-558 + ProtonOC.update_meta_links, model.py:1515-1536, evaluated on the caller's own
// ball.  One block of G threads per evaluation, level-ordered stage-and-relax in shared memory like
// poc_embed2.cuh (kept for A/B runs, POC_EMB_V=2); what changed and why (profiles/r02b_*: 41 k warp instructions per
// evaluation, 60 % of them in the staging loop, a third of the resident warps parked at block barriers):
//   * every undirected meta edge is staged and relaxed ONCE.  The reference's meta graph is undirected, and links are
//     made in pairs: the entries u -> v and v -> u nearly always carry the same multiplicity.  The staging pass takes a
//     multiplex entry only from the row of the lower slot (one compare, before the in-ball lookup) unless the entry is
//     flagged ENT_ASYM (poc_device.cuh: its reverse is missing or differs -- then both are taken, each with its own
//     weight, exactly as before).  Half of the in-ball entries leave the loop after ten instructions and the sweeps run
//     over half as many staged entries.  The relaxations that remain are the same set of float64 sums: distances, and
//     so the result, are bit-identical to poc_embed2.cuh, poc_embed.cuh and oracle/hotpath.c.
//   * multiplex rows and criminal rows are walked by separate, uniform loops (a warp whose 32 entries straddled both
//     kinds used to execute both bodies).
//   * 13 block barriers per evaluation instead of 19: per-level append counters (nobody waits for thread 0 to publish a
//     level), the OC test folded into a voting barrier, the source's distance set by the warp that stages its row;
//     small BFS levels are spread over all warps (one row each) because their length is latency, not work.
#pragma once
#include "poc_embed2.cuh"

struct Emb3Shared {
    unsigned long long num, den;
    unsigned long long st_bytes, st_edges, st_nodes;   // algorithmic-work counters of the block (flushed once at the end)
    long long ph[8];                                    // phase cycle counters (thread 0)
    long long t0;
    unsigned ev_bytes, ev_edges;                        // ... of the evaluation in hand
    int over, nreq, nevals;
    int cnt[MAX_RADIUS + 2];                            // members appended at each BFS level
    int ws[33];
};

template <int G> struct Emb3Layout {
    static constexpr int NW = G / 32;
    static constexpr int OFF_TAB = (int)((sizeof(Emb3Shared) + 15) & ~(size_t)15);   // per warp: 32 rows x int4
    static constexpr int OFF_TABU = OFF_TAB + NW * 32 * (int)sizeof(int4);           // per warp: 32 slots
    static constexpr int OFF_INVW = OFF_TABU + NW * 32 * (int)sizeof(int);
    static constexpr int OFF_BP = OFF_INVW + 128 * (int)sizeof(double);
};
static inline size_t emb3_fixed_bytes(int G, int Ncap) {
    size_t wcap = (size_t)(Ncap + 31) / 32;
    size_t off_bp = ((sizeof(Emb3Shared) + 15) & ~(size_t)15) + (size_t)(G / 32) * 32 * (sizeof(int4) + sizeof(int)) + 128 * sizeof(double);
    return off_bp + ((wcap * 8 + 15) & ~(size_t)15);
}

// lane `lane` holds row `lane` of the batch: exclusive scan of the row lengths, total in every lane
#define EMB3_SCAN(LEN, EXCL, TOT)                                                                                      \
    int EXCL, TOT;                                                                                                     \
    {                                                                                                                  \
        int incl_ = (LEN);                                                                                             \
        _Pragma("unroll") for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(FULL_MASK, incl_, o); if (lane >= o) incl_ += y; } \
        EXCL = incl_ - (LEN); TOT = __shfl_sync(FULL_MASK, incl_, 31);                                                 \
    }
// which compacted non-empty row the flattened position fb + lane belongs to
#define EMB3_ROW_OF(LEN, EXCL, K)                                                                                      \
    int K;                                                                                                             \
    {                                                                                                                  \
        const unsigned d_ = (unsigned)((EXCL) - fb);                                                                   \
        const unsigned starts_ = __reduce_or_sync(FULL_MASK, ((LEN) > 0 && d_ < 32u) ? (1u << d_) : 0u);               \
        K = rows_before + __popc(starts_ & (0xffffffffu >> (31 - lane))) - 1;                                          \
        rows_before += __popc(starts_);                                                                                \
    }

template <int G, int MB>
__global__ void __launch_bounds__(G, MB) k_emb3(DevCtx C, int parity, int smem_bytes, int mode) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typedef Emb3Layout<G> L;
    constexpr int NW = L::NW;
    // mode 0: the tick's request list; an evaluation whose ball does not fit this block's shared memory is handed to the
    //         second launch (mode 1: few large blocks over C.emb_big) instead of walking the rows from global memory
    // mode 2: the request list, everything evaluated here (small batches run large blocks in the first place)
    const int nreqs = (mode == 1) ? C.n_emb_big[C.r0 + blockIdx.y] : C.n_emb[parity * C.R + C.r0 + blockIdx.y];
    if ((int)blockIdx.x >= nreqs) return;  // one block per request; the rest leave at once
    const int r = C.r0 + blockIdx.y, n = C.n_agents[r], nwords = (n + 31) >> 5, wcap = (C.Ncap + 31) >> 5;
    const int gt = threadIdx.x, w = gt >> 5;
    int lane = threadIdx.x & 31;
    unsigned ltmask = (1u << lane) - 1u;
    OPAQUE(lane); OPAQUE(ltmask);
    Emb3Shared *S = (Emb3Shared *)smem_raw;
    double *invw = (double *)(smem_raw + L::OFF_INVW);
    uint2 *bp = (uint2 *)(smem_raw + L::OFF_BP);          // {bitmap word, ball members in the words before it}
    unsigned a_base = smem_addr(smem_raw);
    OPAQUE(a_base);
    unsigned a_tab = a_base + L::OFF_TAB + w * 512;       // this warp's 32-row slab
    OPAQUE(a_tab);
    const unsigned a_tabu = a_base + L::OFF_TABU + w * 128;
    const unsigned a_bp = a_base + L::OFF_BP, a_invw = a_base + L::OFF_INVW;
    int off_dyn = L::OFF_BP + ((wcap * 8 + 15) & ~15);
    OPAQUE(off_dyn);
    const int dyn_bytes = smem_bytes - off_dyn;
    const int slot = r * C.slot_stride + blockIdx.x;   // (poc_device.cuh: one stride for every traversal kernel)
    int *list = C.scr_list + (size_t)slot * (C.Ncap + 1);
    const poc_params &P = C.params[r];
    Graph g = graph_of(C, r);
    asm volatile("" : "+l"(g.u_ent)); asm volatile("" : "+l"(g.c_ent));
    const uint32_t *attr = RP(C.attr, r, C.Ncap);
    const int32_t *reqs = (mode == 1) ? C.emb_big + (size_t)r * C.big_cap : C.emb_list + ((size_t)parity * C.R + r) * C.Ncap;
    const int R = P.oc_embeddedness_radius;
    if (gt < 8) S->ph[gt] = 0;
    if (gt == 8) { S->st_bytes = 0ull; S->st_edges = 0ull; S->st_nodes = 0ull; S->nreq = 0; S->nevals = 0; }
    for (int i = gt; i < 128; i += G) invw[i] = C.invw[i];
#define PHASE(k) { long long t1 = clock64(); S->ph[k] += t1 - S->t0; S->t0 = t1; }
    for (int item = blockIdx.x; item < nreqs; item += gridDim.x) {
        const int self = reqs[item];
        // ---- radius-R ball: frontier BFS over the multiplex + criminal rows, visited bits in bp[].x
        for (int i = gt; i < nwords; i += G) bp[i] = make_uint2((i == (self >> 5)) ? (1u << (self & 31)) : 0u, 0u);
        if (gt < MAX_RADIUS + 2) S->cnt[gt] = 0;
        if (gt == 0) {
            list[0] = self; S->over = 0; S->nreq++; S->num = 0ull; S->den = 0ull;
            S->t0 = clock64(); S->ev_bytes = 0u; S->ev_edges = 0u;
        }
        __syncthreads();
        // ring k = list[off_k .. off_k + cnt[k]), ring 0 = self; S->cnt[k] = members of ring k (k >= 1)
        int lb = 0, le = 1;
#pragma unroll 1
        for (int lvl = 1; lvl <= R; lvl++) {
            const int nwp = min(NW, le - lb);   // a level this small is a chain of dependent loads: one row per warp
            if (w < nwp) {
                for (int i0 = lb; i0 < le; i0 += nwp * 32) {
                    const int i = i0 + lane * nwp + w;
                    int b0 = 0, e0 = 0, b1 = 0, len1 = 0;
                    if (i < le) {
                        const int u = list[i];
                        b0 = g.u_rowptr[u]; e0 = g.u_rowptr[u + 1];
                        b1 = g.c_rowptr[u]; len1 = g.c_rowptr[u + 1] - b1;
                    }
                    const int len0 = e0 - b0;
                    {   // algorithmic bytes of these rows: 16 B of row pointers + 4 B per multiplex entry + 8 B per criminal entry
                        const unsigned t0_ = __reduce_add_sync(FULL_MASK, (unsigned)len0), c1 = __reduce_add_sync(FULL_MASK, (unsigned)len1);
                        const unsigned nr = __popc(__ballot_sync(FULL_MASK, i < le));
                        if (lane == 0) { atomicAdd(&S->ev_edges, t0_ + c1); atomicAdd(&S->ev_bytes, 16u * nr + 4u * t0_ + 8u * c1); }
                    }
#define EMB3_VISIT(V)                                                                                                  \
                    {                                                                                                  \
                        bool isnew_ = false;                                                                           \
                        if ((V) >= 0) {                                                                                \
                            const unsigned bit_ = 1u << ((V) & 31);                                                    \
                            isnew_ = !(atomicOr(&bp[(V) >> 5].x, bit_) & bit_);                                        \
                        }                                                                                              \
                        const unsigned m_ = __ballot_sync(FULL_MASK, isnew_);                                          \
                        if (m_) {                                                                                      \
                            int pos_ = 0;                                                                              \
                            if (lane == 0) pos_ = atomicAdd(&S->cnt[lvl], __popc(m_));                                 \
                            pos_ = __shfl_sync(FULL_MASK, pos_, 0);                                                    \
                            if (isnew_) list[le + pos_ + __popc(m_ & ltmask)] = (V);                                   \
                        }                                                                                              \
                    }
                    {   // multiplex parts, four entries (one aligned 16-byte load) per lane and step
                        const int qb = b0 >> 2, lenq = (len0 > 0) ? ((e0 + 3) >> 2) - qb : 0;
                        EMB3_SCAN(lenq, excl, tot)
                        const unsigned nem = __ballot_sync(FULL_MASK, lenq > 0);
                        if (lenq > 0) sts_i4(a_tab + 16u * __popc(nem & ltmask), make_int4(qb - excl, b0, e0, 0));
                        __syncwarp();
                        int rows_before = 0;
                        for (int fb = 0; fb < tot; fb += 32) {
                            const int f = fb + lane;
                            EMB3_ROW_OF(lenq, excl, k)
                            uint4 q = make_uint4(0u, 0u, 0u, 0u);   // mask 0: not an entry
                            if (f < tot) {
                                const int4 rt = lds_i4(a_tab + 16u * k);
                                const int qi = rt.x + f, p0 = qi << 2;
                                q = __ldg(reinterpret_cast<const uint4 *>(g.u_ent) + qi);
                                if (p0 < rt.y || p0 >= rt.z) q.x = 0u;   // the quad's ends may belong to the neighbouring rows
                                if (p0 + 1 < rt.y || p0 + 1 >= rt.z) q.y = 0u;
                                if (p0 + 2 < rt.y || p0 + 2 >= rt.z) q.z = 0u;
                                if (p0 + 3 < rt.y || p0 + 3 >= rt.z) q.w = 0u;
                            }
#pragma unroll 1
                            for (int c = 0; c < 4; c++) {
                                const int v = ENT_MASK(q.x) ? ENT_COL(q.x) : -1;
                                EMB3_VISIT(v)
                                q.x = q.y; q.y = q.z; q.z = q.w;
                            }
                        }
                        __syncwarp();   // the criminal parts overwrite the row table
                    }
                    {   // criminal parts
                        EMB3_SCAN(len1, excl, tot)
                        if (tot) {
                            const unsigned nem = __ballot_sync(FULL_MASK, len1 > 0);
                            if (len1 > 0) sts_i4(a_tab + 16u * __popc(nem & ltmask), make_int4(b1 - excl, 0, 0, 0));
                            __syncwarp();
                            int rows_before = 0;
                            for (int fb = 0; fb < tot; fb += 32) {
                                const int f = fb + lane;
                                EMB3_ROW_OF(len1, excl, k)
                                int v = -1;
                                if (f < tot) {
                                    const int4 rt = lds_i4(a_tab + 16u * k);
                                    const int2 ce = __ldg(&g.c_ent[rt.x + f]);
                                    if (!CE_IS_DEAD(ce.y)) v = ce.x;
                                }
                                EMB3_VISIT(v)
                            }
                            __syncwarp();   // the next batch overwrites the row table
                        }
                    }
#undef EMB3_VISIT
                }
            }
            __syncthreads();
            lb = le; le += S->cnt[lvl];   // nobody touches this level's counter again
        }
        const int total = le;
        bool anyoc = false;
        for (int i = 1 + gt; i < total; i += G) anyoc |= (attr[list[i]] & F_OC) != 0;
        anyoc = __syncthreads_or(anyoc);
        if (gt == 0) { S->st_nodes += total; PHASE(PH_BFS) }
        if (!anyoc) {  // no OC member in the ball: only the rows the BFS expanded were needed (+ OC flags, result)
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
        const int dist_bytes = (total * 8 + 15) & ~15;
        unsigned long long *dist = (unsigned long long *)(smem_raw + off_dyn);
        bool staged = total <= RCAP && dist_bytes + NW * 256 <= dyn_bytes;
        // (block-uniform) hand the evaluation to the large-block launch; `continue`s when the list took it
#define EMB3_DEFER(UNDO_EVAL)                                                                                          \
        {                                                                                                              \
            if (gt == 0) {                                                                                             \
                const int q_ = atomicAdd(&C.n_emb_big[r], 1);                                                          \
                S->ws[32] = q_;                                                                                        \
                if (q_ < C.big_cap) { (C.emb_big + (size_t)r * C.big_cap)[q_] = self; S->nreq--; S->st_nodes -= total; S->nevals -= (UNDO_EVAL); } \
            }                                                                                                          \
            __syncthreads();                                                                                           \
            const bool deferred_ = S->ws[32] < C.big_cap;                                                              \
            __syncthreads();                                                                                           \
            if (deferred_) continue;                                                                                   \
        }
        if (!staged && mode == 0) EMB3_DEFER(0)
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
            if (staged) for (int i = gt; i < total; i += G) dist[i] = DIST_INF;
            __syncthreads();
            int carry = 0;
            for (int q = 0; q < w; q++) carry += S->ws[q];
            if (carry) for (int i = wb + lane; i < we; i += 32) bp[i].y += (unsigned)carry;
        }
        __syncthreads();
        if (gt == 0) { S->nevals++; PHASE(PH_PREFIX) }
#define RANK(v) ((int)(bp[(v) >> 5].y + __popc(bp[(v) >> 5].x & ((1u << ((v) & 31)) - 1u))))
        int rounds = 0;
        // a staged entry is rank(u) | rank(v) << 12 | multiplicity << 24.  Multiplicities from 254 on (co-offence counts of a
        // long-lived OC pair pass 255 late in a run) take two slots: the entry with 255 in the field, then 254 << 24 | count
        constexpr unsigned RB = 12, RMASK = (1u << RB) - 1u;
        if (staged) {
            // ---- split of the dynamic region: distances of the ball first, one region of staged entries per warp behind them
            int ecap_w = ((dyn_bytes - dist_bytes) >> 2) / NW;          // staged entries per warp in shared memory
            unsigned a_dist = a_base + off_dyn, a_sedge = a_dist + dist_bytes + 4u * (w * ecap_w);
            OPAQUE(ecap_w); OPAQUE(a_dist); OPAQUE(a_sedge);
            const int gcap_w = C.ecap / NW;
            unsigned *gedge = C.scr_edges + (size_t)slot * C.ecap + (size_t)w * gcap_w;
            int wcnt = 0, wrel = 0;   // entries this warp has staged / has relaxed once (warp-uniform)
            // one staged entry: compacted over the lanes and appended to the warp's region.  The relaxations run over the
            // appended entries afterwards, 32 at a time: a lane that drops an entry here (half of them, see the header)
            // would otherwise idle through its neighbours' relaxation code
#define EMB3_TAKE(OK, RU, RV, WI)                                                                                      \
            {                                                                                                          \
                const unsigned m_ = __ballot_sync(FULL_MASK, OK);                                                      \
                if (OK) {                                                                                              \
                    const int pos_ = wcnt + __popc(m_ & ltmask);                                                       \
                    const unsigned pk_ = (unsigned)(RU) | ((unsigned)(RV) << RB) | ((unsigned)(WI) << (2 * RB));       \
                    if (pos_ < ecap_w) sts_u32(a_sedge + 4u * pos_, pk_);                                              \
                    else if (pos_ - ecap_w < gcap_w) gedge[pos_ - ecap_w] = pk_;                                       \
                    else S->over = 1;                                                                                  \
                }                                                                                                      \
                wcnt += __popc(m_);                                                                                    \
            }
            // the same for a criminal entry, whose multiplicity may need the second slot
#define EMB3_TAKE_C(OK, RU, RV, WGT)                                                                                   \
            {                                                                                                          \
                const bool big_ = (OK) && (WGT) >= 254;                                                                \
                const unsigned m_ = __ballot_sync(FULL_MASK, OK), mb_ = __ballot_sync(FULL_MASK, big_);                \
                if (OK) {                                                                                              \
                    const int pos_ = wcnt + __popc(m_ & ltmask) + __popc(mb_ & ltmask);                                \
                    const unsigned pk_ = (unsigned)(RU) | ((unsigned)(RV) << RB) | ((big_ ? 255u : (unsigned)(WGT)) << (2 * RB)); \
                    if (pos_ + (big_ ? 1 : 0) < ecap_w) {                                                              \
                        sts_u32(a_sedge + 4u * pos_, pk_);                                                             \
                        if (big_) sts_u32(a_sedge + 4u * pos_ + 4u, (254u << 24) | (unsigned)(WGT));                   \
                    } else if (pos_ >= ecap_w && pos_ + (big_ ? 1 : 0) - ecap_w < gcap_w) {                            \
                        gedge[pos_ - ecap_w] = pk_;                                                                    \
                        if (big_) gedge[pos_ + 1 - ecap_w] = (254u << 24) | (unsigned)(WGT);                           \
                    } else if (pos_ < ecap_w && gcap_w > 0) {   /* the pair straddles the two regions */               \
                        sts_u32(a_sedge + 4u * pos_, pk_); gedge[0] = (254u << 24) | (unsigned)(WGT);                  \
                    } else S->over = 1;                                                                                \
                }                                                                                                      \
                wcnt += __popc(m_) + __popc(mb_);                                                                      \
            }
            // relax the warp's staged entries [FROM, TO) both ways (undirected meta edge); CHG: any distance moved
#define EMB3_RELAX(FROM, TO, CHG, REV)                                                                                 \
            for (int t_ = (FROM) + lane; t_ < (TO); t_ += 32) {                                                        \
                const int e_ = (REV) ? (FROM) + (TO) - 1 - t_ : t_;                                                    \
                const unsigned pk_ = (e_ < ecap_w) ? lds_u32(a_sedge + 4u * e_) : gedge[e_ - ecap_w];                  \
                const unsigned ru_ = pk_ & RMASK, rv_ = (pk_ >> RB) & RMASK;                                           \
                unsigned wg_ = pk_ >> (2 * RB);                                                                        \
                double wt_;                                                                                            \
                if (wg_ < 128u) wt_ = lds_f64(a_invw + 8u * wg_);   /* nearly every entry */                           \
                else {                                                                                                 \
                    if (wg_ == 254u) continue;   /* second slot of the entry before it */                              \
                    if (wg_ == 255u) wg_ = ((e_ + 1 < ecap_w) ? lds_u32(a_sedge + 4u * (e_ + 1)) : gedge[e_ + 1 - ecap_w]) & 0xffffffu; \
                    wt_ = 1.0 / (double)wg_;                                                                           \
                }                                                                                                      \
                const double du_ = lds_f64(a_dist + 8u * ru_), dv_ = lds_f64(a_dist + 8u * rv_);                       \
                const double nv_ = du_ + wt_, nu_ = dv_ + wt_;                                                         \
                if (nv_ < dv_) { smin_f64(a_dist + 8u * rv_, nv_); CHG = true; }                                       \
                if (nu_ < du_) { smin_f64(a_dist + 8u * ru_, nu_); CHG = true; }                                       \
            }
            // Level by level: each participating warp takes 32 rows (interleaved, so the OC clique spreads over the warps),
            // flattens their multiplex parts over its lanes, then their criminal parts.
            lb = 0; le = 0;
#pragma unroll 1
            for (int lvl = 0; lvl <= R; lvl++) {
                lb = le; le += (lvl == 0) ? 1 : S->cnt[lvl];
                const int nwp = min(NW, (le - lb + 3) >> 2);
                if (lvl == 0 && gt == 0) sts_u64(a_dist + 8u * RANK(self), 0ull);   // warp 0 stages the source's row: program order
                if (w < nwp) {
                    if (lvl == 0) __syncwarp();
                    for (int i0 = lb; i0 < le; i0 += nwp * 32) {
                        const int i = i0 + lane * nwp + w;
                        int u = 0, b0 = 0, e0 = 0, b1 = 0, len1 = 0, ru = 0, rowlen = 0;
                        if (i < le) {
                            u = list[i];
                            e0 = g.u_rowptr[u + 1]; rowlen = e0 - g.u_rowptr[u];
                            b0 = g.u_mid[u];   // the entries before it point to lower slots, whose rows hold the same pairs
                            b1 = g.c_rowptr[u]; len1 = g.c_rowptr[u + 1] - b1;
                            ru = RANK(u);
                        }
                        const int len0 = e0 - b0;
                        {   // (the algorithmic bytes count the whole rows, SURVEY.md 8(d), although half of each is read)
                            const unsigned t0_ = __reduce_add_sync(FULL_MASK, (unsigned)rowlen), c1 = __reduce_add_sync(FULL_MASK, (unsigned)len1);
                            const unsigned nr = __popc(__ballot_sync(FULL_MASK, i < le));
                            if (lane == 0) { atomicAdd(&S->ev_edges, t0_ + c1); atomicAdd(&S->ev_bytes, 16u * nr + 4u * t0_ + 8u * c1); }
                        }
                        // -- multiplex parts, four entries (one aligned 16-byte load) per lane and step:
                        //    {first quad - first flattened position, row begin, row end, rank} + the slot in a second table
                        {
                            const int qb = b0 >> 2, lenq = (len0 > 0) ? ((e0 + 3) >> 2) - qb : 0;
                            EMB3_SCAN(lenq, excl, tot)
                            const unsigned nem = __ballot_sync(FULL_MASK, lenq > 0);
                            if (lenq > 0) {
                                const unsigned kk = __popc(nem & ltmask);
                                sts_i4(a_tab + 16u * kk, make_int4(qb - excl, b0, e0, ru));
                                sts_u32(a_tabu + 4u * kk, (unsigned)u);
                            }
                            __syncwarp();
                            int rows_before = 0;
                            for (int fb = 0; fb < tot; fb += 32) {
                                const int f = fb + lane;
                                EMB3_ROW_OF(lenq, excl, k)
                                uint4 q = make_uint4(0u, 0u, 0u, 0u);   // mask 0: not an entry
                                int tru = 0, uu = 0;
                                if (f < tot) {
                                    const int4 rt = lds_i4(a_tab + 16u * k);
                                    uu = (int)lds_u32(a_tabu + 4u * k);
                                    const int qi = rt.x + f, p0 = qi << 2;
                                    q = __ldg(reinterpret_cast<const uint4 *>(g.u_ent) + qi);
                                    if (p0 < rt.y || p0 >= rt.z) q.x = 0u;   // the quad's ends may belong to the neighbouring rows
                                    if (p0 + 1 < rt.y || p0 + 1 >= rt.z) q.y = 0u;
                                    if (p0 + 2 < rt.y || p0 + 2 >= rt.z) q.z = 0u;
                                    if (p0 + 3 < rt.y || p0 + 3 >= rt.z) q.w = 0u;
                                    tru = rt.w;
                                }
                                // one meta edge per pair: from the lower slot's row, unless the reverse entry differs.  A
                                // neighbour that also has a criminal entry is weighed there (mask copy + co-offences)
#define EMB3_ENTRY(ENT)                                                                                                \
                                {                                                                                      \
                                    bool ok_ = false;                                                                  \
                                    int rv_ = 0;                                                                       \
                                    const int v_ = ENT_COL(ENT);                                                       \
                                    if (ENT_MASK(ENT) && !((ENT) & ENT_CRIM) && (v_ > uu || ((ENT) & ENT_ASYM))) {     \
                                        const uint2 bw_ = lds_u2(a_bp + 8u * (v_ >> 5));                               \
                                        if ((bw_.x >> (v_ & 31)) & 1u) {                                               \
                                            ok_ = true;                                                                \
                                            rv_ = (int)bw_.y + __popc(bw_.x & ((1u << (v_ & 31)) - 1u));               \
                                        }                                                                              \
                                    }                                                                                  \
                                    EMB3_TAKE(ok_, tru, rv_, __popc(ENT_MASK(ENT)))                                    \
                                }
                                // (a rolled loop over the quad: four copies of the body cost more in instruction-cache misses --
                                //  1.2 stalled warps per issue in profiles/r02i -- than the three moves per step cost here)
#pragma unroll 1
                                for (int c = 0; c < 4; c++) {
                                    EMB3_ENTRY(q.x)
                                    q.x = q.y; q.y = q.z; q.z = q.w;
                                }
#undef EMB3_ENTRY
                            }
                            __syncwarp();   // the criminal parts overwrite the row table
                        }
                        // -- criminal parts (few rows have one): weight = co-offences + layers of the mask copy
                        {
                            EMB3_SCAN(len1, excl, tot)
                            if (tot) {
                                const unsigned nem = __ballot_sync(FULL_MASK, len1 > 0);
                                if (len1 > 0) sts_i4(a_tab + 16u * __popc(nem & ltmask), make_int4(b1 - excl, u, ru, 0));
                                __syncwarp();
                                int rows_before = 0;
                                for (int fb = 0; fb < tot; fb += 32) {
                                    const int f = fb + lane;
                                    EMB3_ROW_OF(len1, excl, k)
                                    bool ok = false;
                                    int tru = 0, rv = 0, wi = 0;
                                    if (f < tot) {
                                        const int4 rt = lds_i4(a_tab + 16u * k);
                                        const int2 ce = __ldg(&g.c_ent[rt.x + f]);
                                        const int v = ce.x, wgt = CE_CNT(ce.y) + __popc(CE_SMASK(ce.y));
                                        const uint2 bw = lds_u2(a_bp + 8u * (v >> 5));
                                        // once per pair like the multiplex entries: from the lower slot's row unless flagged
                                        if (wgt > 0 && (v > rt.y || (ce.y & CE_ASYM) || (C.emb_flags & 1)) && ((bw.x >> (v & 31)) & 1u)) {   // 0: the reference divides by zero here
                                            ok = true; tru = rt.z; wi = wgt;
                                            rv = (int)bw.y + __popc(bw.x & ((1u << (v & 31)) - 1u));
                                            if (wgt > 0xffffff) { S->over = 1; ok = false; }
                                        }
                                    }
                                    EMB3_TAKE_C(ok, tru, rv, wi)
                                }
                                __syncwarp();
                            }
                        }
                        // -- the entries of these rows travel outwards once before the next level is staged
                        if (!S->over) {
                            bool chg_ = false;
                            __syncwarp();
                            EMB3_RELAX(wrel, wcnt, chg_, false)
                            (void)chg_;
                        }
                        wrel = wcnt;
                    }
                }
                __syncthreads();
            }
#undef EMB3_TAKE
#undef EMB3_TAKE_C
            if (S->over) staged = false;
            if (gt == 0) PHASE(PH_STAGE)
            if (!staged && mode == 0) EMB3_DEFER(1)
            if (staged) {
                if (lane == 0 && wcnt > ecap_w) S->ph[PH_SPILL]++;
                for (;;) {   // plain sweeps (every warp over the entries it staged) until one changes nothing
                    bool changed = false;
                    EMB3_RELAX(0, wcnt, changed, false)   // (alternating the direction of the passes measured slower: 3.8 passes per evaluation against 3.5)
                    rounds++;
                    if (!__syncthreads_or(changed) || rounds > total + 2) break;
                }
                if (gt == 0) { PHASE(PH_SWEEP) S->ph[PH_NSWEEPS] += rounds; }
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
#undef EMB3_RELAX
        }
        if (!staged) {
            // very large ball, or more large multiplicities than the table holds: distances in global scratch, relaxation
            // straight from the rows (every entry, both ways), one warp per node per round
            unsigned long long *gdist = C.scr_dist + (size_t)slot * (C.Ncap + 1);
            __syncthreads();
            if (gt == 0) { S->ph[PH_FALLBACK]++; S->ev_bytes = 4u * total + 8u; S->num = 0ull; S->den = 0ull; }
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
            if (gt == 0) { PHASE(PH_SWEEP) S->ph[PH_NSWEEPS] += rounds; }
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
#undef EMB3_DEFER
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
