// poc_kernels.cuh -- the hot-path kernels (sm_100a).  Included once by poc_api.cu.
// Every kernel is batched: blockIdx.y (or blockIdx.x for the one-block-per-replica kernels) selects the
// replica, so one launch serves the whole ensemble batch.
#pragma once
#include "poc_device.cuh"

#define RP(ptr, r, stride) ((ptr) + (size_t)(r) * (size_t)(stride))
#include "poc_embed.cuh"
#include "poc_embed2.cuh"
#include "poc_embed3.cuh"
#include "poc_embed_acc.cuh"
#include "poc_demog.cuh"

// ============================================================================ state I/O
struct StageCols {
    uint8_t *alive, *male, *oc_member, *facilitator, *prisoner, *has_job, *has_school, *target;
    int8_t *job_level, *education_level, *wealth_level;
    int32_t *age, *birth_tick, *num_crimes, *num_crimes_tick, *father, *new_recruit;
    double *propensity, *tendency, *sentence, *arrest_weight;
};

__global__ void k_pack_agents(DevCtx C, int r, int n, StageCols S, int *err) {
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a == 0) { C.n_agents[r] = n; C.u_par[r] = 0; C.cdf_valid[r] = 0; }   // an upload lands in the first multiplex buffer
    if (a >= n) return;
    int age = S.age[a], job = S.job_level[a], edu = S.education_level[a], w = S.wealth_level[a];
    if (age < 0 || job < 0 || job > 15 || edu < 0 || edu > 15 || w < 0 || w > 15) atomicExch(err, 1);
    if (age > 255) age = 255;
    uint32_t x = (uint32_t)age | ((uint32_t)(job & 15) << 16) | ((uint32_t)(edu & 15) << 20) | ((uint32_t)(w & 15) << 24);
    if (S.alive[a]) x |= F_ALIVE;
    if (S.male[a]) x |= F_MALE;
    if (S.oc_member[a]) x |= F_OC;
    if (S.facilitator[a]) x |= F_FAC;
    if (S.prisoner[a]) x |= F_PRISONER;
    if (S.has_job[a]) x |= F_HASJOB;
    if (S.has_school[a]) x |= F_HASSCHOOL;
    if (S.target[a]) x |= F_TARGET;
    size_t o = (size_t)r * C.Ncap + a;
    C.attr[o] = x;
    C.birth[o] = S.birth_tick[a]; C.ncrimes[o] = S.num_crimes[a]; C.ncrimes_tick[o] = S.num_crimes_tick[a];
    C.father[o] = S.father[a]; C.new_recruit[o] = S.new_recruit[a];
    C.propensity[o] = S.propensity[a]; C.tendency[o] = S.tendency[a]; C.sentence[o] = S.sentence[a];
    C.arrest_w[o] = S.arrest_weight[a];
    C.emb_stamp[o] = 0;
    C.ncrim_dead[o] = 0;
}

__global__ void k_unpack_agents(DevCtx C, int r, int n, StageCols S) {
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    size_t o = (size_t)r * C.Ncap + a;
    uint32_t x = C.attr[o];
    S.age[a] = attr_age(x); S.job_level[a] = (int8_t)attr_job(x); S.education_level[a] = (int8_t)attr_edu(x);
    S.wealth_level[a] = (int8_t)attr_wealth(x);
    S.alive[a] = (x & F_ALIVE) != 0; S.male[a] = (x & F_MALE) != 0; S.oc_member[a] = (x & F_OC) != 0;
    S.facilitator[a] = (x & F_FAC) != 0; S.prisoner[a] = (x & F_PRISONER) != 0; S.has_job[a] = (x & F_HASJOB) != 0;
    S.has_school[a] = (x & F_HASSCHOOL) != 0; S.target[a] = (x & F_TARGET) != 0;
    S.birth_tick[a] = C.birth[o]; S.num_crimes[a] = C.ncrimes[o]; S.num_crimes_tick[a] = C.ncrimes_tick[o];
    S.father[a] = C.father[o]; S.new_recruit[a] = C.new_recruit[o];
    S.propensity[a] = C.propensity[o]; S.tendency[a] = C.tendency[o]; S.sentence[a] = C.sentence[o];
    S.arrest_weight[a] = C.arrest_w[o];
}

// The eight non-criminal layers arrive as separate CSRs (rows ascending); one thread k-way-merges the
// rows of one agent into packed {slot, layer mask} entries.  fill = 0: count only.
struct StageLayers { const int32_t *rowptr[8]; const int32_t *col[8]; };
__global__ void k_union_rows(DevCtx C, int r, int n, StageLayers L, int *lens, int fill, int *err) {
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    int cur[8], end[8];
#pragma unroll
    for (int l = 0; l < 8; l++) { cur[l] = L.rowptr[l][a]; end[l] = L.rowptr[l][a + 1]; }
    int out = fill ? RP(C.u_rowptr, r, C.Ncap + 1)[a] : 0, cnt = 0, prev = -1;
    uint32_t *ent = RP(C.u_ent, r, C.ucap);
    for (;;) {
        int mn = 0x7fffffff;
#pragma unroll
        for (int l = 0; l < 8; l++) if (cur[l] < end[l]) { int v = L.col[l][cur[l]]; if (v < mn) mn = v; }
        if (mn == 0x7fffffff) break;
        if (mn <= prev || mn < 0 || mn >= n || mn == a) atomicExch(err, 2);  // unsorted row / bad slot / self loop
        prev = mn;
        uint32_t mask = 0;
#pragma unroll
        for (int l = 0; l < 8; l++) if (cur[l] < end[l] && L.col[l][cur[l]] == mn) { mask |= 1u << l; cur[l]++; }
        if (fill) ent[out + cnt] = (uint32_t)mn | (mask << 24);
        cnt++;
    }
    if (!fill) lens[a] = cnt;
}

// single-block exclusive scan: out[0..n] (out[n] = total)
__global__ void __launch_bounds__(1024) k_exscan_i32(const int *in, int *out, int n) {
    __shared__ int ws[33];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += blockDim.x) {
        int i = base + threadIdx.x, v = (i < n) ? in[i] : 0, tot;
        int ex = block_exscan(v, ws, &tot);
        if (i < n) out[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry;
}

__global__ void k_pack_criminal(DevCtx C, int r, int n, const int32_t *rowptr, const int32_t *col, const int32_t *co, int *err) {
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a > n) return;
    RP(C.c_rowptr, r, C.Ncap + 1)[a] = rowptr[a];
    if (a == n) { C.c_par[r] = 0; return; }   // an upload always lands in the first of the two criminal CSR buffers
    int2 *ent = RP(C.c_ent, r, C.ccap);
    int prev = -1;
    for (int e = rowptr[a]; e < rowptr[a + 1]; e++) {
        int v = col[e];
        if (v <= prev || v < 0 || v >= n || v == a) atomicExch(err, 3);
        prev = v;
        ent[e] = make_int2(v, co ? co[e] : 0);
    }
}

// After both the multiplex CSR and the criminal CSR of a replica are in place: copy the static layer mask of the
// same neighbour into every criminal entry and flag that static entry (see poc_device.cuh).
__global__ void k_link_criminal(DevCtx C, int r, int n) {
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    Graph g = graph_of(C, r);
    uint32_t *uent = const_cast<uint32_t *>(g.u_ent);
    int2 *cent = RP(C.c_ent, r, C.ccap);
    for (int e = g.c_rowptr[a]; e < g.c_rowptr[a + 1]; e++) {
        int2 ce = cent[e];
        int idx = static_find_idx(g, a, ce.x);
        unsigned smask = 0;
        if (idx >= 0) { smask = ENT_MASK(uent[idx]); uent[idx] |= ENT_CRIM; }
        cent[e].y = CE_CNT(ce.y) | (int)(smask << 24);
    }
}

// ENT_ASYM of every multiplex entry of an uploaded state (poc_device.cuh); warp per agent, lane per entry
__global__ void k_mark_asym(DevCtx C, int r, int n) {
    const int a = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (a >= n) return;
    Graph g = graph_of(C, r);
    uint32_t *uent = const_cast<uint32_t *>(g.u_ent);
    const int b = g.u_rowptr[a], end = g.u_rowptr[a + 1];
    int lower = 0;
    bool low_flag = false;
    for (int e0 = b; e0 < end; e0 += 32) {
        const int e = e0 + lane_id();
        bool below = false, flag = false;
        if (e < end) {
            const uint32_t en = uent[e] & ~ENT_ASYM;
            flag = ent_is_asym(g, a, en);
            uent[e] = en | (flag ? ENT_ASYM : 0u);
            below = ENT_COL(en) < a;
        }
        lower += __popc(__ballot_sync(FULL_MASK, below));
        low_flag |= __any_sync(FULL_MASK, below && flag);
    }
    if (lane_id() == 0) g.u_mid[a] = low_flag ? b : b + lower;
    int2 *cent = const_cast<int2 *>(g.c_ent);
    for (int e = g.c_rowptr[a] + lane_id(); e < g.c_rowptr[a + 1]; e += 32) {
        const int2 ce = cent[e];
        if (CE_IS_DEAD(ce.y)) continue;
        const int y = ce.y & ~CE_ASYM;
        cent[e].y = y | (crim_is_asym(g, a, ce.x, y) ? CE_ASYM : 0);
    }
}

// tot_<layer>_link of calculate_fast_reporters (model.py:1715-1732) for the eight static layers: computed once here
// and kept current by get_caught (the only place on the path that removes entries)
__global__ void k_layer_totals(DevCtx C, int r, int n) {
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (a < n) {
        Graph g = graph_of(C, r);
        for (int e = g.u_rowptr[a]; e < g.u_rowptr[a + 1]; e++) {
            uint32_t m = ENT_MASK(g.u_ent[e]);
#pragma unroll
            for (int l = 0; l < 8; l++) cnt[l] += (m >> l) & 1;
        }
        // Person.number_of_children is not among the uploaded columns: it starts from the offspring links
        // (poc_upload_children replaces it)
        RP(C.n_children, r, C.Ncap)[a] = cnt[1];
        if (!(RP(C.attr, r, C.Ncap)[a] & F_ALIVE)) for (int l = 0; l < 8; l++) cnt[l] = 0;
    }
#pragma unroll
    for (int l = 0; l < 8; l++) {
        int v = cnt[l];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
        if (lane_id() == 0 && v) atomicAdd((unsigned long long *)&RP(C.layer_tot, r, 8)[l], (unsigned long long)v);
    }
}

// ============================================================================ tick bookkeeping
// model.py:236-238 (+ extra.py:359-361): per scheduled agent zero the this-tick counter and recompute age.
__global__ void k_begin_tick(DevCtx C) {
    int tick = C.ti[C.grp].tick;
    int r = C.r0 + blockIdx.y, a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= C.n_agents[r]) return;
    size_t o = (size_t)r * C.Ncap + a;
    uint32_t x = C.attr[o];
    if (!(x & F_ALIVE)) return;
    const int tpy = 12;   // extra._age hard-codes 12 (extra.py:359-361), whatever ticks_per_year is set to
    int dt = tick - C.birth[o];
    int age = (dt >= 0) ? dt / tpy : -((-dt + tpy - 1) / tpy);
    age = age < 0 ? 0 : (age > 255 ? 255 : age);
    C.attr[o] = (x & ~0xffu) | (uint32_t)age;
    C.ncrimes_tick[o] = 0;
}

// model.py:279-283
__global__ void k_end_tick(DevCtx C, int advance) {
    // last kernel of the tick for this replica group: move the device-side tick state on
    if (advance && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) { TickInfo &t = C.ti[C.grp]; t.tick++; t.epoch++; t.rep++; }
    int r = C.r0 + blockIdx.y, a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= C.n_agents[r]) return;
    size_t o = (size_t)r * C.Ncap + a;
    uint32_t x = C.attr[o];
    if ((x & (F_ALIVE | F_PRISONER)) != (F_ALIVE | F_PRISONER)) return;
    double s = C.sentence[o] - 1.0;
    C.sentence[o] = s;
    if (s == 0.0) C.attr[o] = x & ~F_PRISONER;
}

// Age/sex cell partition (model.py:1172-1173, 1430-1432): stable (ascending slot) member lists per cell,
// the strict counts calculate_crime_multiplier uses (model.py:1152, Q8) and, when `yearly`, the multiplier.
// One block per replica.
// Every thread owns a contiguous slice of the slots and counts its members per cell into its own column of a
// shared-memory table cnt[cell][thread]; warp `c` then turns row c into exclusive prefixes (a 1024-element scan
// per cell, all cells in parallel), which are exactly the stable write offsets of the second pass.
// Dynamic shared memory: n_cells_max * blockDim.x ints.
__device__ __forceinline__ bool is_yearly(const DevCtx &C, const poc_params &P) {
    int tick = C.ti[C.grp].tick;
    return (tick % P.ticks_per_year) == 0 || tick == 1;
}
// mode: 0 partition only, 1 also the crime multiplier, 2 the multiplier on yearly ticks (model.py:248)
__global__ void __launch_bounds__(1024) k_cells(DevCtx C, int mode, int begin) {
    extern __shared__ int cnt_tab[];  // [nc][blockDim.x]
    int r = C.r0 + blockIdx.x, n = C.n_agents[r], tid = threadIdx.x, lane = lane_id(), w = warp_id(), nw = blockDim.x >> 5;
    const poc_params &P = C.params[r];
    bool yearly = mode == 1 || (mode == 2 && is_yearly(C, P));
    int nc = P.n_cells, NT = blockDim.x;
    __shared__ unsigned char lut[512], luts[512];
    __shared__ int base[POC_MAX_CELLS + 1], tot[POC_MAX_CELLS], adj[POC_MAX_CELLS], s_alive;
    for (int i = tid; i < 512; i += NT) {
        int male = i >> 8, age = i & 255, c = 255, cs = 255;
        for (int k = nc - 1; k >= 0; k--)
            if (P.cell_male[k] == male) {
                if (P.cell_age_from[k] <= age && age <= P.cell_age_to[k]) c = k;
                if (P.cell_age_from[k] < age && age <= P.cell_age_to[k]) cs = k;
            }
        lut[i] = (unsigned char)c; luts[i] = (unsigned char)cs;
    }
    for (int k = 0; k < nc; k++) cnt_tab[k * NT + tid] = 0;
    if (tid < POC_MAX_CELLS) adj[tid] = 0;
    if (tid == 0) s_alive = 0;
    __syncthreads();
    uint32_t *attr = RP(C.attr, r, C.Ncap);
    int chunk = (n + NT - 1) / NT;
    int a_begin = min(n, tid * chunk), a_end = min(n, a_begin + chunk);
    int alive_cnt = 0;
    const int tick_now = C.ti[C.grp].tick, tpy = 12;   // extra._age hard-codes 12 (extra.py:359-361)
    for (int a = a_begin; a < a_end; a++) {
        uint32_t x = attr[a];
        if (!(x & F_ALIVE)) continue;
        if (begin) {  // model.py:236-238, extra.py:359-361
            int dt = tick_now - RP(C.birth, r, C.Ncap)[a];
            int age = (dt >= 0) ? dt / tpy : -((-dt + tpy - 1) / tpy);
            age = age < 0 ? 0 : (age > 255 ? 255 : age);
            x = (x & ~0xffu) | (uint32_t)age;
            attr[a] = x;
            RP(C.ncrimes_tick, r, C.Ncap)[a] = 0;
        }
        alive_cnt++;
        int i = ((x & F_MALE) ? 256 : 0) | attr_age(x);
        int c = lut[i], cs = luts[i];
        if (c != 255) cnt_tab[c * NT + tid]++;
        if (cs != c) {  // the strict membership of calculate_crime_multiplier (Q8) differs: record the difference
            if (c != 255) atomicSub(&adj[c], 1);
            if (cs != 255) atomicAdd(&adj[cs], 1);
        }
    }
    for (int o = 16; o > 0; o >>= 1) alive_cnt += __shfl_xor_sync(FULL_MASK, alive_cnt, o);
    if (lane == 0) atomicAdd(&s_alive, alive_cnt);
    __syncthreads();
    // warp c: exclusive scan of row c (one element per thread of the block), 32 elements per step
    for (int c = w; c < nc; c += nw) {
        int carry = 0;
        for (int base_t = 0; base_t < NT; base_t += 32) {
            int v = cnt_tab[c * NT + base_t + lane], x = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(FULL_MASK, x, o); if (lane >= o) x += y; }
            cnt_tab[c * NT + base_t + lane] = carry + x - v;
            carry += __shfl_sync(FULL_MASK, x, 31);
        }
        if (lane == 0) tot[c] = carry;
    }
    __syncthreads();
    if (tid == 0) {
        int s = 0;
        for (int k = 0; k < nc; k++) { base[k] = s; s += tot[k]; }
        base[nc] = s;
        int32_t *co = RP(C.cell_off, r, POC_MAX_CELLS + 1);
        bool moved = false;
        for (int k = 0; k <= nc; k++) { moved |= co[k] != base[k]; co[k] = base[k]; }
        if (moved || yearly) C.cdf_valid[r] = 0;   // k_draw keeps its CDFs while cells, tendencies and multiplier stay as they are
        for (int k = 0; k < nc; k++) RP(C.cell_strict, r, POC_MAX_CELLS)[k] = tot[k] + adj[k];
        RP(C.counters, r, CN_COUNT)[CN_N_ALIVE] = s_alive;
        if (yearly) {  // model.py:1150-1157, left-to-right float64
            double total = 0.0;
            for (int k = 0; k < nc; k++) total += P.cell_c[k] * (double)(tot[k] + adj[k]);
            C.crime_mult[r] = P.crimes_yearly_per10k / 10000.0 * (double)s_alive / total;
        }
    }
    __syncthreads();
    int32_t *members = RP(C.cell_members, r, C.Ncap);
    bool changed = false;
    for (int a = a_begin; a < a_end; a++) {
        uint32_t x = attr[a];
        if (!(x & F_ALIVE)) continue;
        int c = lut[((x & F_MALE) ? 256 : 0) | attr_age(x)];
        if (c == 255) continue;
        const int pos = base[c] + cnt_tab[c * NT + tid]++;
        if (members[pos] != a) { members[pos] = a; changed = true; }
    }
    if (__syncthreads_or(changed) && tid == 0) C.cdf_valid[r] = 0;
}

// ============================================================================ A1 criminal tendency
// Person.update_criminal_tendency, entities.py:337-367 (quirks Q2, Q3 kept).
// Eight lanes share one agent: the row is read coalesced and the five neighbour counts (integers, so the
// order of the partial sums does not matter) are combined with three shuffles.
// The grid is a few blocks per replica striding over the agents, so that the eleven launches a year that
// only find out it is not a yearly tick cost one small wave.
__global__ void __launch_bounds__(256) k_tendency_pre(DevCtx C, int mode) {
    int r = C.r0 + blockIdx.y, sub = threadIdx.x & 7;
    if (mode == 2 && !is_yearly(C, C.params[r])) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) C.cdf_valid[r] = 0;   // the crime draw's CDFs are built from the tendencies
    const poc_params &P = C.params[r];
    Graph g = graph_of(C, r);
    const int32_t *ncr = RP(C.ncrimes, r, C.Ncap);
    int n = C.n_agents[r];
    // four agents per lane group and trip, so that four chains of dependent loads (row bounds -> entry ->
    // neighbour's crime count) are in flight per lane
    for (int a0 = blockIdx.x * 128; a0 < n; a0 += gridDim.x * 128) {   // block-uniform trip count
        int av[4], e0[4], e1[4], fam4[4], famc4[4], nf4[4], np4[4], pc4[4];
        uint32_t xv[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            av[q] = a0 + q * 32 + (threadIdx.x >> 3);
            bool in = av[q] < n;
            xv[q] = in ? C.attr[(size_t)r * C.Ncap + av[q]] : 0u;
            e0[q] = in ? g.u_rowptr[av[q]] + sub : 0; e1[q] = in ? g.u_rowptr[av[q] + 1] : 0;
            fam4[q] = famc4[q] = nf4[q] = np4[q] = pc4[q] = 0;
        }
#pragma unroll
        for (int q = 0; q < 4; q++) if (!(xv[q] & F_ALIVE)) e1[q] = 0;
        while (e0[0] < e1[0] || e0[1] < e1[1] || e0[2] < e1[2] || e0[3] < e1[3]) {
            uint32_t ent[4];
#pragma unroll
            for (int q = 0; q < 4; q++) ent[q] = (e0[q] < e1[q]) ? g.u_ent[e0[q]] : 0u;
            int nc4[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                uint32_t m = ENT_MASK(ent[q]);
                nc4[q] = (m & (MB_SIB | MB_OFF | MB_PARTNER | MB_PROF)) ? ncr[ENT_COL(ent[q])] : 0;
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                uint32_t m = ENT_MASK(ent[q]);   // a padding entry (0) has an empty mask
                int f = __popc(m & (MB_SIB | MB_OFF | MB_PARTNER));
                bool crim = nc4[q] > 0;
                fam4[q] += f; if (crim) famc4[q] += f;
                if (m & MB_FRIEND) nf4[q]++;
                if (m & MB_PROF) { np4[q]++; if (crim) pc4[q]++; }
                e0[q] += 8;
            }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
#pragma unroll
            for (int s = 1; s < 8; s <<= 1) {
                fam4[q] += __shfl_xor_sync(FULL_MASK, fam4[q], s); famc4[q] += __shfl_xor_sync(FULL_MASK, famc4[q], s);
                nf4[q] += __shfl_xor_sync(FULL_MASK, nf4[q], s); np4[q] += __shfl_xor_sync(FULL_MASK, np4[q], s);
                pc4[q] += __shfl_xor_sync(FULL_MASK, pc4[q], s);
            }
        }
        // lane q of the group finishes agent q
        if (sub >= 4) continue;
        int a = a0 + sub * 32 + (threadIdx.x >> 3);
        int fam = 0, famc = 0, nf = 0, np = 0, pc = 0;
        uint32_t x = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) if (sub == q) { fam = fam4[q]; famc = famc4[q]; nf = nf4[q]; np = np4[q]; pc = pc4[q]; x = xv[q]; }
        if (!(x & F_ALIVE)) continue;
        size_t o = (size_t)r * C.Ncap + a;
        int age = attr_age(x), male = (x & F_MALE) ? 1 : 0, cell = -1;
        for (int k = 0; k < P.n_cells; k++)
            if (P.cell_male[k] == male && P.cell_age_from[k] <= age && age <= P.cell_age_to[k]) { cell = k; break; }
        if (cell < 0) continue;
        double t = P.cell_c[cell];
        t *= (attr_job(x) == 1) ? 1.30 : 1.0;
        t *= (attr_edu(x) >= 2) ? 0.94 : 1.0;
        t *= (C.propensity[o] > P.propensity_threshold) ? 1.97 : 1.0;
        t *= 1.62;
        t *= (fam > 0 && ((double)famc / (double)fam) > 0.5) ? 1.45 : 1.0;
        bool neigh = nf > 0;
        if (!neigh && np > 0) neigh = ((double)pc / (double)(nf + np)) > 0.5;
        t *= neigh ? 1.81 : 1.0;
        t *= (x & F_OC) ? 4.50 : 1.0;
        C.tendency[o] = t;
    }
}

// numpy's pairwise add.reduce over tend[idx[0..n)] (what np.mean uses at model.py:1182) splits the range in halves
// (left half rounded down to a multiple of 8) until a piece has <= 128 elements, sums each piece with eight strided
// accumulators, and adds the pieces up the same tree.  The pieces are independent, so the tree is walked three
// times: pw_split lists the pieces, eight lanes per piece sum them, pw_combine adds the sums in tree order.
// (explicit stacks instead of recursion: the depth is log2(n / 64) + 1 <= 24)
__device__ void pw_split(int start, int n, int2 *leaf, int &cur) {
    int ss[24], sn[24], sp = 0;
    ss[0] = start; sn[0] = n; sp = 1;
    while (sp) {
        int s = ss[--sp], m = sn[sp];
        while (m > 128) {                 // descend to the left, remember the right half
            int n2 = m / 2;
            n2 -= n2 % 8;
            ss[sp] = s + n2; sn[sp] = m - n2; sp++;
            m = n2;
        }
        leaf[cur++] = make_int2(s, m);
    }
}
__device__ double pw_combine(int n, const double *sum, int &cur) {
    int fn[24], fs[24], sp = 0;          // frame: piece length, state (0 = nothing done, 1 = left done)
    double fl[24], ret = 0.0;
    fn[0] = n; fs[0] = 0; sp = 1;
    while (sp) {
        int m = fn[sp - 1];
        if (m <= 128) { ret = sum[cur++]; sp--; continue; }
        int n2 = m / 2;
        n2 -= n2 % 8;
        if (fs[sp - 1] == 0) { fs[sp - 1] = 1; fn[sp] = n2; fs[sp] = 0; sp++; }
        else if (fs[sp - 1] == 1) { fl[sp - 1] = ret; fs[sp - 1] = 2; fn[sp] = m - n2; fs[sp] = 0; sp++; }
        else { ret = fl[sp - 1] + ret; sp--; }
    }
    return ret;
}
// one piece (n <= 128) by the eight lanes of a group; the result is valid in every lane of the group
__device__ __forceinline__ double pw_piece8(const double *t, const int *idx, int n, int sub) {
    // every lane of the warp runs the shuffles (groups with different n share a warp)
    int full = (n >= 8) ? n - (n % 8) : 0;
    double q = 0.0;
    if (full) {
        q = t[idx[sub]];
        for (int i = 8; i < full; i += 8) q += t[idx[i + sub]];
    }
    // ((q0+q1)+(q2+q3))+((q4+q5)+(q6+q7)): each level adds two operands, so the exchange order within a pair is free
    q += __shfl_xor_sync(FULL_MASK, q, 1);
    q += __shfl_xor_sync(FULL_MASK, q, 2);
    q += __shfl_xor_sync(FULL_MASK, q, 4);
    if (!full) q = 0.0;                             // n < 8: plain left-to-right sum from 0.0
    for (int i = full; i < n; i++) q += t[idx[i]];
    return q;
}

// epsilon re-centring per cell, model.py:1180-1184.  One block per replica.
__global__ void __launch_bounds__(1024) k_tendency_eps(DevCtx C, int mode) {
    int r = C.r0 + blockIdx.x, tid = threadIdx.x;
    const poc_params &P = C.params[r];
    if (mode == 2 && !is_yearly(C, P)) return;
    __shared__ double eps[POC_MAX_CELLS];
    __shared__ int off[POC_MAX_CELLS + 1], lbase[POC_MAX_CELLS + 1];
    const int32_t *members = RP(C.cell_members, r, C.Ncap);
    double *tend = RP(C.tendency, r, C.Ncap);
    int2 *leaf = RP(C.pw_leaf, r, C.pw_cap);
    double *lsum = RP(C.pw_sum, r, C.pw_cap);
    int nc = P.n_cells;
    if (tid <= nc) off[tid] = RP(C.cell_off, r, POC_MAX_CELLS + 1)[tid];
    __syncthreads();
    if (tid == 0) {  // a cell of m members has at most m / 64 + 1 pieces
        int s = 0;
        for (int k = 0; k < nc; k++) { lbase[k] = s; s += (off[k + 1] - off[k]) / 64 + 1; }
        lbase[nc] = s;
    }
    __syncthreads();
    if (tid < nc) {
        int cur = lbase[tid], m = off[tid + 1] - off[tid];
        if (m > 0) pw_split(off[tid], m, leaf, cur);
        for (; cur < lbase[tid + 1]; cur++) leaf[cur] = make_int2(0, 0);
    }
    __syncthreads();
    int nl = lbase[nc], sub = tid & 7;
    for (int j0 = 0; j0 < nl; j0 += blockDim.x >> 3) {   // warp-uniform trip count: the shuffles stay convergent
        int j = j0 + (tid >> 3);
        int2 lf = (j < nl) ? leaf[j] : make_int2(0, 0);
        double s = pw_piece8(tend, members + lf.x, lf.y, sub);
        if (j < nl && sub == 0) lsum[j] = s;
    }
    __syncthreads();
    if (tid < nc) {
        int cur = lbase[tid], m = off[tid + 1] - off[tid];
        eps[tid] = (m > 0) ? P.cell_c[tid] - pw_combine(m, lsum, cur) / (double)m : 0.0;
    }
    __syncthreads();
    int total = off[nc];
    for (int i = tid; i < total; i += blockDim.x) {
        int k = 0;
        while (i >= off[k + 1]) k++;
        tend[members[i]] += eps[k];
    }
}

// ============================================================================ A3 crime draw
// model.py:1427-1440 + extra.py:102-149.  One block per replica.  The CDF is built exactly as
// weighted_n_of + Generator.choice build it: optional min-shift, left-to-right sum, p/sum, sequential
// cumsum, divide by the last element; one lane per cell runs the two order-defined chains.
__global__ void __launch_bounds__(1024) k_draw(DevCtx C, int smem_doubles, int allow_fit) {
    extern __shared__ double wt_sm[];
    int tick = C.ti[C.grp].tick;
    int r = C.r0 + blockIdx.x, tid = threadIdx.x, lane = lane_id(), w = warp_id(), nw = blockDim.x >> 5;
    const poc_params &P = C.params[r];
    int nc = P.n_cells;
    __shared__ int off[POC_MAX_CELLS + 1], ncr[POC_MAX_CELLS], coff[POC_MAX_CELLS + 1], nz[POC_MAX_CELLS];
    __shared__ double mn[POC_MAX_CELLS], last[POC_MAX_CELLS];
    __shared__ int neg[POC_MAX_CELLS], nzc[POC_MAX_CELLS];
    __shared__ double s_sum[POC_MAX_CELLS];
    const int32_t *members = RP(C.cell_members, r, C.Ncap);
    const double *tend = RP(C.tendency, r, C.Ncap);
    double *cdf = RP(C.cdf, r, C.Ncap);
    int *cn = RP(C.counters, r, CN_COUNT);
    if (tid <= nc) off[tid] = RP(C.cell_off, r, POC_MAX_CELLS + 1)[tid];
    // counter snapshot for the per-tick deltas (read before thread 0 updates the cumulative counters below)
    if (tid >= 64 && tid < 64 + CN_COUNT) RP(C.prev_counters, r, CN_COUNT)[tid - 64] = cn[tid - 64];
    __syncthreads();
    if (tid < nc) {
        int m = off[tid + 1] - off[tid];
        double target = P.cell_c[tid] * (double)m / (double)P.ticks_per_year * C.crime_mult[r];
        int k = (int)rint(target);  // np.round: half to even (Q1)
        ncr[tid] = k > 0 ? k : 0;
    }
    __syncthreads();
    if (tid == 0) {
        int s = 0;
        for (int k = 0; k < nc; k++) { coff[k] = s; s += ncr[k]; }
        coff[nc] = s;
        if (s > C.Ccap) { atomicExch(&cn[CN_ERROR], 10); s = C.Ccap; for (int k = 0; k <= nc; k++) if (coff[k] > s) coff[k] = s; }
        C.n_crimes[r] = s;
        C.n_search[r] = 0;
        C.n_pairs[r] = 0;
        cn[CN_TICK_CRIMES] = s; cn[CN_TICK_MEMBERS] = 0; cn[CN_TICK_ARRESTS] = 0; cn[CN_TICK_EMB] = 0; cn[CN_TICK_RECRUITED] = 0;
        cn[CN_NUMBER_CRIMES] += s;
        C.n_emb[r] = 0; C.n_emb[C.R + r] = 0;
        C.n_arrested[r] = 0;
    }
    // The cell-ordered weights live in shared memory when the population fits (every pass below is then on chip);
    // larger populations work in the global cdf array.  In shared memory every cell starts at a multiple of eight
    // and is padded with zeros to a multiple of eight, so the serial chains below run without per-element
    // predicates (acc + 0.0 == acc: the weights are >= 0 once the min-shift is applied).
    __shared__ int poff[POC_MAX_CELLS + 1];
    if (tid == 0) {
        int s = 0;
        for (int k = 0; k < nc; k++) { poff[k] = s; s += (off[k + 1] - off[k] + 7) & ~7; }
        poff[nc] = s;
    }
    __syncthreads();
    int total = off[nc], ptotal = poff[nc];
    // The CDFs depend on the tendencies (yearly), the cells' member lists (birthdays, deaths) and, through the crimes per
    // cell, the multiplier (yearly): on most ticks nothing of that moved and last tick's CDFs, kept in the global cdf array,
    // are the ones this tick would build (the same inputs through the same chains): the draws read them there.
    const bool cached = C.cdf_valid[r] != 0;
    if (cached && tid < nc) nz[tid] = RP(C.cdf_nz, r, POC_MAX_CELLS)[tid];
    // allow_fit = 0: the launch reserved the chain windows only (populations too large for the whole-cell layout), not the
    // cell-of-eight bytes behind the weights, whatever this replica's own alive count happens to be
    const bool fit = allow_fit && ptotal <= smem_doubles;
    unsigned char *cellof8 = (unsigned char *)(wt_sm + smem_doubles);   // cell of every block of eight positions
    const int *woff = (fit && !cached) ? poff : off;
    double *wt = (fit && !cached) ? wt_sm : cdf;
    if (!cached) {
    if (fit) {
        for (int k = w; k < nc; k += nw)
            for (int q = (poff[k] >> 3) + lane; q < (poff[k + 1] >> 3); q += 32) cellof8[q] = (unsigned char)k;
        __syncthreads();
        for (int q0 = tid; q0 < ptotal; q0 += 8 * blockDim.x) {  // eight independent gathers in flight per thread
            int id[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                int q = q0 + u * blockDim.x;
                id[u] = -1;
                if (q < ptotal) { int k = cellof8[q >> 3], i = q - poff[k]; if (i < off[k + 1] - off[k]) id[u] = members[off[k] + i]; }
            }
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) v[u] = (id[u] >= 0) ? tend[id[u]] : 0.0;
#pragma unroll
            for (int u = 0; u < 8; u++) { int q = q0 + u * blockDim.x; if (q < ptotal) wt_sm[q] = v[u]; }
        }
    } else {
#pragma unroll 4
        for (int i = tid; i < total; i += blockDim.x) cdf[i] = tend[members[i]];
    }
    __syncthreads();
    for (int k = w; k < nc; k += nw) {   // per-cell min / any-negative
        double m = INFINITY;
        bool ng = false;
        int len = off[k + 1] - off[k];
        const double *row = wt + woff[k];
        for (int base = 0; base < len; base += 256) {   // eight independent loads in flight per lane
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) { int i = base + u * 32 + lane; v[u] = (i < len) ? row[i] : INFINITY; }
#pragma unroll
            for (int u = 0; u < 8; u++) { m = fmin(m, v[u]); ng |= (v[u] < 0.0); }
        }
        for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(FULL_MASK, m, o));
        ng = __any_sync(FULL_MASK, ng);
        if (lane == 0) { mn[k] = m; neg[k] = ng; }
    }
    __syncthreads();
    if (fit) {
        // The two order-defined chains (sum, then cumulative sum of p/sum) are serial within a cell and independent
        // across cells: lane k of warp 0 walks cell k in shared memory (one LDS + one DADD per element, the next
        // eight values in flight); shift, non-zero count and the divisions stay parallel.
        if (tid < nc) { nzc[tid] = 0; if (ncr[tid] == 0) { nz[tid] = 1; last[tid] = 1.0; } }
        __syncthreads();
        for (int q0 = 0; q0 < ptotal; q0 += blockDim.x) {     // block-uniform trip count (ballot below)
            int q = q0 + tid, k = -1;
            double x = 0.0;
            if (q < ptotal) {
                k = cellof8[q >> 3];
                if (ncr[k] > 0) { x = wt_sm[q]; if (neg[k] && q - poff[k] < off[k + 1] - off[k]) { x = x - mn[k]; wt_sm[q] = x; } }
            }
            // eight consecutive positions belong to one cell: one shared-memory add per group of eight lanes
            unsigned nzb = (__ballot_sync(FULL_MASK, x != 0.0) >> (lane & 24)) & 0xffu;
            if ((lane & 7) == 0 && nzb) atomicAdd(&nzc[k], __popc(nzb));
        }
        __syncthreads();
        bool act = lane < nc && ncr[lane] > 0;
        int len8 = act ? poff[lane + 1] - poff[lane] : 0, maxlen = len8;
        for (int o = 16; o > 0; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(FULL_MASK, maxlen, o));
        double *src = wt_sm + (act ? poff[lane] : 0);
        if (w == 0) {
            double acc = 0.0, xa[8], xb[8];
#pragma unroll
            for (int q = 0; q < 8; q++) xa[q] = (len8 > 0) ? src[q] : 0.0;
            for (int j = 0; j < maxlen; j += 16) {     // two halves in flight: one is summed while the other loads
                bool pb = j + 8 < len8, pa = j + 16 < len8;
#pragma unroll
                for (int q = 0; q < 8; q++) xb[q] = pb ? src[j + 8 + q] : 0.0;
#pragma unroll
                for (int q = 0; q < 8; q++) acc = acc + xa[q];
#pragma unroll
                for (int q = 0; q < 8; q++) xa[q] = pa ? src[j + 16 + q] : 0.0;
#pragma unroll
                for (int q = 0; q < 8; q++) acc = acc + xb[q];
            }
            s_sum[lane] = acc;
        }
        __syncthreads();
        for (int q = tid; q < ptotal; q += blockDim.x) {
            int k = cellof8[q >> 3];
            if (ncr[k] > 0 && s_sum[k] != 0.0) wt_sm[q] = wt_sm[q] / s_sum[k];
        }
        __syncthreads();
        if (w == 0) {
            double acc = 0.0, xa[8], xb[8];
            if (act && s_sum[lane] == 0.0) len8 = 0;
#pragma unroll
            for (int q = 0; q < 8; q++) xa[q] = (len8 > 0) ? src[q] : 0.0;
            for (int j = 0; j < maxlen; j += 16) {
                bool sa = j < len8, pb = j + 8 < len8, pa = j + 16 < len8;
#pragma unroll
                for (int q = 0; q < 8; q++) xb[q] = pb ? src[j + 8 + q] : 0.0;
#pragma unroll
                for (int q = 0; q < 8; q++) { acc = acc + xa[q]; if (sa) src[j + q] = acc; }
#pragma unroll
                for (int q = 0; q < 8; q++) xa[q] = pa ? src[j + 16 + q] : 0.0;
#pragma unroll
                for (int q = 0; q < 8; q++) { acc = acc + xb[q]; if (pb) src[j + 8 + q] = acc; }
            }
            if (act) { nz[lane] = (s_sum[lane] == 0.0) ? 0 : nzc[lane]; last[lane] = acc; }
        }
        __syncthreads();
        for (int q = tid; q < ptotal; q += blockDim.x) {
            int k = cellof8[q >> 3];
            if (ncr[k] > 0 && nz[k] > 0) wt_sm[q] = wt_sm[q] / last[k];
        }
    } else {
        // Population too large for shared memory: the weights stay in the global cdf array and the two chains run
        // over a shared-memory window of W elements per cell (all threads fill the window with coalesced loads, lane
        // k of warp 0 walks cell k's part, the accumulators carry over from window to window).
        if (tid < nc) { nzc[tid] = 0; if (ncr[tid] == 0) { nz[tid] = 1; last[tid] = 1.0; } }
        __syncthreads();
        for (int k = w; k < nc; k += nw) {   // min-shift and non-zero count, warp per cell
            if (ncr[k] == 0) continue;
            int len = off[k + 1] - off[k], cnt = 0;
            double *row = cdf + off[k];
            for (int base = 0; base < len; base += 128) {
                double x[4];
#pragma unroll
                for (int u = 0; u < 4; u++) { int i = base + u * 32 + lane; x[u] = (i < len) ? row[i] : 0.0; }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    int i = base + u * 32 + lane;
                    if (neg[k] && i < len) { x[u] = x[u] - mn[k]; row[i] = x[u]; }
                    cnt += __popc(__ballot_sync(FULL_MASK, x[u] != 0.0));
                }
            }
            if (lane == 0) nzc[k] = cnt;
        }
        __syncthreads();
        const int W = min(1024, (smem_doubles / max(nc, 1) - 1) & ~7);   // window per cell, a multiple of eight
        const int WS = W + 1;                                             // row stride: odd, so the lanes of the chain warp hit different banks
        bool act = lane < nc && ncr[lane] > 0;
        int len = act ? off[lane + 1] - off[lane] : 0, maxlen = len;
        for (int o = 16; o > 0; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(FULL_MASK, maxlen, o));
        double *src = wt_sm + lane * WS;
        for (int pass = 0; pass < 2; pass++) {
            double acc = 0.0;
            if (pass == 1 && act && s_sum[lane] == 0.0) len = 0;
            for (int win = 0; win < maxlen; win += W) {
                // fill: cell k's elements [win, win + W), zero behind its end (rounded up to the chain's step of 16)
                for (int k = 0; k < nc; k++) {
                    if (ncr[k] == 0 || (pass == 1 && s_sum[k] == 0.0)) continue;
                    int lk = off[k + 1] - off[k] - win;
                    if (lk <= 0) continue;
                    int fill = min(W, (lk + 15) & ~15);
                    double s = (pass == 1) ? s_sum[k] : 1.0;
                    for (int j = tid; j < fill; j += blockDim.x) {
                        double x = (j < lk) ? cdf[off[k] + win + j] : 0.0;
                        wt_sm[k * WS + j] = (pass == 1) ? x / s : x;
                    }
                }
                __syncthreads();
                if (w == 0) {
                    int n8 = min(W, (max(len - win, 0) + 7) & ~7);   // this lane's part of the window, padded
                    int wmax = n8;
                    for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(FULL_MASK, wmax, o));
                    double xa[8], xb[8];
#pragma unroll
                    for (int q = 0; q < 8; q++) xa[q] = (n8 > 0) ? src[q] : 0.0;
                    for (int j = 0; j < wmax; j += 16) {
                        bool sa = j < n8, pb = j + 8 < n8, pa = j + 16 < n8;
#pragma unroll
                        for (int q = 0; q < 8; q++) xb[q] = pb ? src[j + 8 + q] : 0.0;
#pragma unroll
                        for (int q = 0; q < 8; q++) { acc = acc + xa[q]; if (pass == 1 && sa) src[j + q] = acc; }
#pragma unroll
                        for (int q = 0; q < 8; q++) xa[q] = pa ? src[j + 16 + q] : 0.0;
#pragma unroll
                        for (int q = 0; q < 8; q++) { acc = acc + xb[q]; if (pass == 1 && pb) src[j + 8 + q] = acc; }
                    }
                }
                __syncthreads();
                if (pass == 1) {   // write the cumulative sums of this window back
                    for (int k = 0; k < nc; k++) {
                        if (ncr[k] == 0 || s_sum[k] == 0.0) continue;
                        int lk = min(W, off[k + 1] - off[k] - win);
                        for (int j = tid; j < lk; j += blockDim.x) cdf[off[k] + win + j] = wt_sm[k * WS + j];
                    }
                    __syncthreads();
                }
            }
            if (w == 0) {
                if (pass == 0) s_sum[lane] = acc;
                else if (act) { nz[lane] = (s_sum[lane] == 0.0) ? 0 : nzc[lane]; last[lane] = acc; }
            }
            __syncthreads();
        }
        for (int k = 0; k < nc; k++) {
            if (!(ncr[k] > 0 && nz[k] > 0)) continue;
            double l = last[k];
            for (int i = off[k] + tid; i < off[k + 1]; i += blockDim.x) cdf[i] = cdf[i] / l;
        }
    }
    __syncthreads();
    // keep what was built for the ticks to come (cell_members order, without the padding of the shared-memory layout)
    if (fit)
        for (int q = tid; q < ptotal; q += blockDim.x) {
            const int k = cellof8[q >> 3], i = q - poff[k];
            if (i < off[k + 1] - off[k]) cdf[off[k] + i] = wt_sm[q];
        }
    if (tid < nc) RP(C.cdf_nz, r, POC_MAX_CELLS)[tid] = nz[tid];
    if (tid == 0) C.cdf_valid[r] = 1;
    }   // !cached
    __syncthreads();
    // the draws
    int ncrimes = coff[nc];
    const int32_t *hdr = RP(C.rp_hdr, r, 8);
    bool replay = hdr[0] != 0;
    const uint32_t *attr = RP(C.attr, r, C.Ncap);
    for (int j = tid; j < ncrimes; j += blockDim.x) {
        int k = 0;
        while (j >= coff[k + 1]) k++;
        double u1, u2;
        if (replay) {
            if (j < hdr[1]) { u1 = RP(C.rp_u_init, r, C.Ccap)[j]; u2 = RP(C.rp_u_size, r, C.Ccap)[j]; }
            else { u1 = 0.0; u2 = 0.0; atomicExch(&cn[CN_ERROR], 3); }
        } else philox_u2(C.philox_key[r], tick, ST_CRIME, j, u1, u2);
        size_t o = (size_t)r * C.Ccap + j;
        if (nz[k] < 1) {  // the reference raises IndexError here (weighted_one_of on all-zero weights)
            atomicExch(&cn[CN_ERROR], 1);
            C.crime_init[o] = -1; C.crime_k[o] = 0; C.crime_flags[o] = 0; C.cr_nacc[o] = 0; C.cr_state[o] = 4;
            continue;
        }
        int m = off[k + 1] - off[k];
        int ii = ss_right(wt + woff[k], m, u1);
        if (ii >= m) ii = m - 1;
        int init = members[off[k] + ii];
        int si = ss_right(P.size_cdf, P.n_sizes, u2);
        if (si >= P.n_sizes) si = P.n_sizes - 1;
        int kk = P.size_n[si] - 1;
        uint32_t ia = attr[init];
        C.crime_init[o] = init; C.crime_k[o] = kk; C.crime_flags[o] = (ia & F_OC) ? 1 : 0;
        // find_accomplices prologue, entities.py:419-426
        int target = kk, fn = 0;
        if (kk > 0) { fn = ((double)kk >= P.threshold_use_facilitators) && !(ia & F_FAC); if (fn) target -= 1; }
        C.cr_target[o] = target; C.cr_nacc[o] = 0; C.cr_d[o] = 1; C.cr_fails[o] = 0;
        C.cr_state[o] = (kk > 0) ? (fn ? 2 : 0) : 4;
        if (kk > 0) { int pos = atomicAdd(&C.n_search[r], 1); RP(C.search_list, r, C.Ccap)[pos] = j; }
    }
}

// ============================================================================ A4-A6 accomplice search
// Person.find_accomplices, entities.py:413-455.  One block per crime with k >= 1.
// cr_state bits: 1 = ring cr_d built and its embeddedness requested (OC initiators), 2 = facilitator still
// needed, 4 = done.  Non-OC initiators run to completion in one call; an OC initiator needs
// oc_embeddedness of every ring member (candidates_weight, entities.py:465-466), so its block requests the
// evaluations, returns, and resumes in the next round after k_embeddedness has run.
struct ArgBest { double key; int slot; int idx; };
__device__ __forceinline__ bool better(double k1, int s1, double k2, int s2) { return (k1 < k2) || (k1 == k2 && s1 < s2); }

#define SEARCH_SCAND 512   /* ring members staged in shared memory for scoring and selection (larger rings stay in scratch) */
template <bool STAGE>
__device__ void search_crime(const DevCtx &C, int r, int c, int tick, int epoch, int parity, unsigned *bitmap,
                             unsigned *fbits, int *list, double *keys, int *s_slot, double *s_key,
                             unsigned long long &edges, unsigned long long &abytes, unsigned long long &ncand) {
    __shared__ int s_count, lvl_off[MAX_RADIUS + 3];
    __shared__ int s_st[6];  // nacc, target, mrem, facneeded, pickidx, stop
    __shared__ double wkey[32];
    __shared__ int wslot[32], widx[32];
    int n = C.n_agents[r], nwords = (n + 31) >> 5, tid = threadIdx.x, lane = lane_id(), w = warp_id(), nw = blockDim.x >> 5;
    const poc_params &P = C.params[r];
    Graph g = graph_of(C, r);
    size_t o = (size_t)r * C.Ccap + c;
    const uint32_t *attr = RP(C.attr, r, C.Ncap);
    const double *tend = RP(C.tendency, r, C.Ncap);
    const double *embv = RP(C.emb_val, r, C.Ncap);
    int *cn = RP(C.counters, r, CN_COUNT);
    int *acc = C.cr_acc + o * POC_MAX_GROUP;
    int init = C.crime_init[o];
    uint32_t self_attr = attr[init];
    bool is_oc = (C.crime_flags[o] & 1) != 0;
    int state = C.cr_state[o], d = C.cr_d[o];
    int maxrad = P.max_accomplice_radius;
    if (tid == 0) { s_st[0] = C.cr_nacc[o]; s_st[1] = C.cr_target[o]; s_st[3] = (state >> 1) & 1; s_st[5] = 0; }
    // bitmap of the initiator's friends (entities.py:687-688)
    for (int i = tid; i < nwords; i += blockDim.x) fbits[i] = 0u;
    __syncthreads();
    for (int e = g.u_rowptr[init] + tid; e < g.u_rowptr[init + 1]; e += blockDim.x) {
        uint32_t ent = g.u_ent[e];
        if (ENT_MASK(ent) & MB_FRIEND) { int v = ENT_COL(ent); atomicOr(&fbits[v >> 5], 1u << (v & 31)); }
    }
    __syncthreads();
    bool awaiting = state & 1;
    while (s_st[0] < s_st[1] && d <= maxrad && !s_st[5]) {
        if (tid == 0) list[0] = init;
        __syncthreads();
        bfs_levels(g, nwords, bitmap, list, 1, d, lvl_off, &s_count, edges, abytes);
        int rb = lvl_off[d], m = lvl_off[d + 1] - rb;
        if (is_oc && !awaiting) {
            // request oc_embeddedness for every ring member not yet evaluated this tick (cache: entities.py:531)
            int32_t *stamp = RP(C.emb_stamp, r, C.Ncap);
            int32_t *elist = C.emb_list + ((size_t)parity * C.R + r) * C.Ncap;
            int *ecount = &C.n_emb[parity * C.R + r];
            for (int i = tid; i < m; i += blockDim.x) {
                int v = list[rb + i];
                if (atomicExch(&stamp[v], epoch) != epoch) elist[atomicAdd(ecount, 1)] = v;
            }
            if (tid == 0) { C.cr_nacc[o] = s_st[0]; C.cr_target[o] = s_st[1]; C.cr_d[o] = d; C.cr_state[o] = 1 | (s_st[3] << 1); }
            return;
        }
        awaiting = false;
        // the ring -- the adjacency this search works on -- can be staged in shared memory (north-star (b)): slots and keys
        // are read once per pick by the whole block; a ring over SEARCH_SCAND members stays in the block's scratch.
        // Measured at 1,024 replicas: 178 ms per 480 ticks with the staging against 168 without (the ring is read a handful
        // of times and was L1-resident already; the copy, its barrier and the smaller L1 cost more): off unless POC_EMB_FLAGS & 2
        // (a template switch: the unstaged kernel keeps plain global loads, k_search<true> is launched under the flag)
        const bool staged = STAGE && m <= SEARCH_SCAND;
        int *cl = staged ? s_slot : list + rb;
        double *ck = staged ? s_key : keys;
        if (staged) {
            for (int i = tid; i < m; i += blockDim.x) s_slot[i] = list[rb + i];
            __syncthreads();
        }
        // candidate keys: warp per candidate
        for (int i = w; i < m; i += nw) {
            int cand = cl[i];
            uint32_t ca = attr[cand];
            bool shared = warp_friend_intersect(g, cand, fbits);
            if (lane == 0) {
                abytes += 20ull + (is_oc ? 8ull : 0ull) + 4ull * (unsigned)(g.u_rowptr[cand + 1] - g.u_rowptr[cand]);
                ncand++;
                double prox = social_proximity(self_attr, ca, shared);
                double key = is_oc ? -1.0 * ((prox * embv[cand]) * tend[cand]) : prox * tend[cand];
                ck[i] = key;
            }
        }
        if (tid == 0) s_st[2] = m;
        __syncthreads();
        // take the smallest keys first (ascending stable sort, entities.py:429-437; ties: ascending slot)
        while (s_st[0] < s_st[1] && s_st[2] > 0 && !s_st[5]) {
            int mrem = s_st[2];
            double bk = INFINITY; int bs = 0x7fffffff, bi = -1;
            for (int i = tid; i < mrem; i += blockDim.x) {
                double k = ck[i]; int s = cl[i];
                if (bi < 0 || better(k, s, bk, bs)) { bk = k; bs = s; bi = i; }
            }
            for (int of = 16; of > 0; of >>= 1) {
                double k2 = __shfl_xor_sync(FULL_MASK, bk, of); int s2 = __shfl_xor_sync(FULL_MASK, bs, of), i2 = __shfl_xor_sync(FULL_MASK, bi, of);
                if (i2 >= 0 && (bi < 0 || better(k2, s2, bk, bs))) { bk = k2; bs = s2; bi = i2; }
            }
            if (lane == 0) { wkey[w] = bk; wslot[w] = bs; widx[w] = bi; }
            __syncthreads();
            if (tid == 0) {
                for (int ww = 1; ww < nw; ww++)
                    if (widx[ww] >= 0 && (bi < 0 || better(wkey[ww], wslot[ww], bk, bs))) { bk = wkey[ww]; bs = wslot[ww]; bi = widx[ww]; }
                int na = s_st[0];
                if (na >= POC_MAX_GROUP - 1) { atomicExch(&cn[CN_ERROR], 2); s_st[5] = 1; }
                else {
                    acc[na] = bs; s_st[0] = na + 1;
                    if (attr[bs] & F_FAC) { s_st[1] += 1; s_st[3] = 0; }  // Q6
                    cl[bi] = cl[mrem - 1]; ck[bi] = ck[mrem - 1]; s_st[2] = mrem - 1;
                }
            }
            __syncthreads();
        }
        d++;
    }
    // ---- epilogue, entities.py:439-455
    int nacc = s_st[0];
    if (s_st[3] && nacc > 0) {
        for (int i = tid; i < nacc; i += blockDim.x) list[i] = acc[i];
        __syncthreads();
        int total = bfs_levels(g, nwords, bitmap, list, nacc, maxrad, lvl_off, &s_count, edges, abytes);
        if (tid == 0) abytes += 4ull * (total - nacc);  // facilitator flag of every node reached
        // facilitators within max_accomplice_radius of any accomplice
        int *avail = (int *)keys;
        if (tid == 0) s_count = 0;
        __syncthreads();
        for (int i = nacc + tid; i < total; i += blockDim.x) { int v = list[i]; if (attr[v] & F_FAC) avail[atomicAdd(&s_count, 1)] = v; }
        __syncthreads();
        int na = s_count;
        if (na > 0) {
            const int32_t *hdr = RP(C.rp_hdr, r, 8);
            int pick = -1;
            if (hdr[0] && RP(C.rp_fac_pick, r, C.Ccap)[c] >= 0) pick = RP(C.rp_fac_pick, r, C.Ccap)[c];
            if (pick < 0) {
                double u = 0.0, u1;
                if (hdr[0] == 2) u = RP(C.rp_u_init, r, C.Ccap)[c];  // stage-hook mode: u_fac supplied here
                else if (!hdr[0]) philox_u2(C.philox_key[r], tick, ST_FAC, c, u, u1);
                int ix = (int)floor(u * (double)na);
                if (ix >= na) ix = na - 1;
                if (tid == 0) s_st[4] = -1;
                __syncthreads();
                for (int i = tid; i < na; i += blockDim.x) {  // ix-th smallest slot (CANON: ascending slot list)
                    int v = avail[i], rank = 0;
                    for (int j = 0; j < na; j++) rank += (avail[j] < v);
                    if (rank == ix) s_st[4] = v;
                }
                __syncthreads();
                pick = s_st[4];
            }
            if (tid == 0 && pick >= 0) {
                bool dup = false;
                for (int i = 0; i < nacc; i++) dup |= (acc[i] == pick);
                if (!dup && nacc < POC_MAX_GROUP - 1) { acc[nacc] = pick; s_st[0] = nacc + 1; }
            }
            __syncthreads();
            nacc = s_st[0];
        }
    }
    if (tid == 0) {
        int target = s_st[1], fails = 0;
        if (nacc < target) { atomicAdd(&cn[CN_SIZE_FAILS], 1); fails |= 1; }
        if ((double)target >= P.threshold_use_facilitators) {
            bool anyf = (self_attr & F_FAC) != 0;
            for (int i = 0; i < nacc; i++) anyf |= (attr[acc[i]] & F_FAC) != 0;
            if (anyf) { atomicAdd(&cn[CN_FAC_CRIMES], 1); fails |= 2; } else { atomicAdd(&cn[CN_FAC_FAILS], 1); fails |= 4; }
        }
        C.cr_nacc[o] = nacc; C.cr_target[o] = target; C.cr_d[o] = d; C.cr_state[o] = 4; C.cr_fails[o] = fails;
    }
}

template <bool STAGE>
__global__ void __launch_bounds__(512, 4) k_search(DevCtx C, int parity) {
    int tick = C.ti[C.grp].tick, epoch = C.ti[C.grp].epoch;
    extern __shared__ unsigned smem_u[];
    int r = C.r0 + blockIdx.y;
    int wcap = (C.Ncap + 31) >> 5;
    unsigned *bitmap = smem_u, *fbits = smem_u + wcap;
    double *s_key = (double *)(smem_u + ((2 * wcap + 1) & ~1));   // 8-byte aligned behind the two bitmaps
    int *s_slot = (int *)(s_key + SEARCH_SCAND);
    // the BFS list shares its slots with the embeddedness kernels of the other replica groups (one stride for all,
    // poc_device.cuh); the keys are this kernel's alone
    int *list = C.scr_list + (size_t)(r * C.slot_stride + blockIdx.x) * (C.Ncap + 1);
    double *keys = C.scr_key + (size_t)(r * gridDim.x + blockIdx.x) * (C.Ncap + 1);
    unsigned long long edges = 0, abytes = 0, ncand = 0, ncrimes = 0;
    // the next round's request list must start empty; nobody appends to it during this launch
    if (blockIdx.x == 0 && threadIdx.x == 0) { C.n_emb[(parity ^ 1) * C.R + r] = 0; C.n_emb_big[r] = 0; }
    int ns = C.n_search[r];
    for (int item = blockIdx.x; item < ns; item += gridDim.x) {
        int c = RP(C.search_list, r, C.Ccap)[item];
        if (C.cr_state[(size_t)r * C.Ccap + c] & 4) continue;
        search_crime<STAGE>(C, r, c, tick, epoch, parity, bitmap, fbits, list, keys, s_slot, s_key, edges, abytes, ncand);
        if (threadIdx.x == 0) { ncrimes++; abytes += 8ull + 4ull * POC_MAX_GROUP; }  // friends row pointer + group out
        __syncthreads();
    }
    edges = warp_sum_u64(edges); abytes = warp_sum_u64(abytes); ncand = warp_sum_u64(ncand);
    if (lane_id() == 0) {
        unsigned long long *stq = RP(C.stats, r, ST_COUNT);
        if (edges) atomicAdd(&C.edges[r], edges);
        if (abytes) atomicAdd(&stq[ST_SEARCH_BYTES], abytes);
        if (ncand) atomicAdd(&stq[ST_SEARCH_CANDIDATES], ncand);
        if (ncrimes) atomicAdd(&stq[ST_SEARCH_CRIMES], ncrimes);
    }
}

// ============================================================================ A7 OC embeddedness: poc_embed.cuh

// ============================================================================ A8 commit: groups, counters, link merge
// ProtonOC.commit_crime, model.py:1487-1504 + Person.add_criminal_link, entities.py:215-224, for all groups
// of the tick at once: canonical groups, per-agent crime counters, ordered co-offender pairs -> bitonic sort
// -> dedup with multiplicity -> merge into the criminal CSR (new links enter with 0 and every co-offence
// adds 1).  One block per replica.
__global__ void __launch_bounds__(1024) k_commit(DevCtx C) {
    int r = C.r0 + blockIdx.x, tid = threadIdx.x, n = C.n_agents[r];
    __shared__ int ws[33];
    __shared__ int carry, s_np;
    const poc_params &P = C.params[r];
    int *cn = RP(C.counters, r, CN_COUNT);
    int ncr = C.n_crimes[r];
    int32_t *goff = RP(C.group_off, r, C.Ccap + 1), *gmem = RP(C.group_mem, r, C.Mcap);
    unsigned long long *pairs = RP(C.pairs, r, C.Pcap);
    int32_t *pcnt = RP(C.pair_cnt, r, C.Pcap);
    const double *tend = RP(C.tendency, r, C.Ncap);
    int32_t *ncrimes = RP(C.ncrimes, r, C.Ncap), *ncrimes_tick = RP(C.ncrimes_tick, r, C.Ncap);
    Graph g = graph_of(C, r);
    // ---- group offsets (flattened criminals list, model.py:1461)
    if (tid == 0) { carry = 0; s_np = 0; }
    __syncthreads();
    for (int base = 0; base < ncr; base += blockDim.x) {
        int c = base + tid, gs = 0, tot;
        if (c < ncr) { size_t o = (size_t)r * C.Ccap + c; gs = (C.crime_init[o] >= 0) ? C.cr_nacc[o] + 1 : 0; }
        int ex = block_exscan(gs, ws, &tot);
        if (c < ncr) goff[c] = carry + ex;
        __syncthreads();
        if (tid == 0) carry += tot;
        __syncthreads();
    }
    if (tid == 0) { goff[ncr] = carry; cn[CN_TICK_MEMBERS] = carry; if (carry > C.Mcap) atomicExch(&cn[CN_ERROR], 11); }
    __syncthreads();
    if (carry > C.Mcap) return;
    // ---- canonical groups (accomplices ascending slot, initiator last), counters, ordered pairs
    for (int c = tid; c < ncr; c += blockDim.x) {
        size_t o = (size_t)r * C.Ccap + c;
        int init = C.crime_init[o];
        if (init < 0) continue;
        int na = C.cr_nacc[o];
        int grp[POC_MAX_GROUP];
        const int *acc = C.cr_acc + o * POC_MAX_GROUP;
        for (int i = 0; i < na; i++) {  // insertion sort
            int v = acc[i], j = i;
            while (j > 0 && grp[j - 1] > v) { grp[j] = grp[j - 1]; j--; }
            grp[j] = v;
        }
        grp[na] = init;
        int gs = na + 1, b = goff[c];
        for (int i = 0; i < gs; i++) { gmem[b + i] = grp[i]; atomicAdd(&ncrimes[grp[i]], 1); atomicAdd(&ncrimes_tick[grp[i]], 1); }
        if (gs > P.this_is_a_big_crime && tend[init] < P.good_guy_threshold) atomicAdd(&cn[CN_BIG_CRIME], 1);
        if (gs > 1) {
            int pos = atomicAdd(&s_np, gs * (gs - 1));
            if (pos + gs * (gs - 1) <= C.Pcap) {
                for (int i = 0; i < gs; i++)
                    for (int j = 0; j < gs; j++)
                        if (i != j) pairs[pos++] = ((unsigned long long)(unsigned)grp[i] << 32) | (unsigned)grp[j];
            }
        }
    }
    __syncthreads();
    int np = s_np;
    if (np > C.Pcap) { if (tid == 0) atomicExch(&cn[CN_ERROR], 12); return; }
    // the criminal CSR is double-buffered: read the current one (g), write the other, flip C.c_par[r] at the end
    int par = C.c_par[r];
    int32_t *rp_new = (par ? C.c_rowptr : C.c_rowptr_tmp) + (size_t)r * (C.Ncap + 1);
    int2 *ent_new = (par ? C.c_ent : C.c_ent_tmp) + (size_t)r * C.ccap;
    int32_t *pnew = RP(C.pair_new_ex, r, C.Pcap + 1), *ins = RP(C.ins_row, r, C.Pcap);
    int32_t *ins_start = RP(C.c_addlen, r, C.Ncap);   // first old entry of every receiving row (at most n of them)
    if (np == 0) { if (tid == 0) C.n_pairs[r] = 0; return; }  // no multi-member group: criminal layer unchanged
    // ---- bitonic sort of the ordered pairs
    int np2 = 1;
    while (np2 < np) np2 <<= 1;
    for (int i = np + tid; i < np2; i += blockDim.x) pairs[i] = ~0ull;
    __syncthreads();
    for (int k = 2; k <= np2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < np2; i += blockDim.x) {
                int ixj = i ^ j;
                if (ixj > i) {
                    unsigned long long a = pairs[i], b = pairs[ixj];
                    bool up = (i & k) == 0;
                    if ((a > b) == up) { pairs[i] = b; pairs[ixj] = a; }
                }
            }
            __syncthreads();
        }
    // ---- dedup with multiplicity (in place; unique pair u gets pcnt[u] = occurrences)
    for (int i = tid; i < np; i += blockDim.x) pcnt[i] = 0;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < np; base += blockDim.x) {
        int i = base + tid;
        bool valid = i < np;
        unsigned long long me = valid ? pairs[i] : 0ull;
        bool head = valid && (i == 0 || pairs[i - 1] != me);
        int tot, ex = block_exscan(head ? 1 : 0, ws, &tot);
        int run = carry + ex + (head ? 0 : -1);
        __syncthreads();
        if (valid) { atomicAdd(&pcnt[run], 1); if (head) pairs[run] = me; }
        __syncthreads();
        if (tid == 0) carry += tot;
        __syncthreads();
    }
    int nu = carry;
    if (tid == 0) C.n_pairs[r] = nu;
    // ---- resolve each unique pair against the old rows: final count, and whether the link is new
    for (int i = tid; i < nu; i += blockDim.x) {
        int a = (int)(pairs[i] >> 32), b = (int)(pairs[i] & 0xffffffffu);
        int cab = crim_find(g, a, b), cba = crim_find(g, b, a);
        int old = (cab >= 0 && cba >= 0) ? cab : 0;  // CANON: a half-present link is re-created with 0
        pcnt[i] = (old + pcnt[i]) | (cab < 0 ? 0x40000000 : 0);
    }
    __syncthreads();
    // ---- rows that receive pairs (the unique pairs are sorted by row), and the number of new links before each pair
    __shared__ int s_nins;
    if (tid == 0) { carry = 0; s_nins = 0; }
    __syncthreads();
    for (int base = 0; base < nu; base += blockDim.x) {
        int i = base + tid;
        bool valid = i < nu;
        int isnew = (valid && (pcnt[i] & 0x40000000)) ? 1 : 0;
        bool head = valid && (i == 0 || (int)(pairs[i - 1] >> 32) != (int)(pairs[i] >> 32));
        int tot, ex = block_exscan(isnew, ws, &tot);
        if (valid) pnew[i] = carry + ex;
        int toth, exh = block_exscan(head ? 1 : 0, ws, &toth);
        if (head) { ins[s_nins + exh] = i; ins_start[s_nins + exh] = g.c_rowptr[(int)(pairs[i] >> 32)]; }
        __syncthreads();
        if (tid == 0) { carry += tot; s_nins += toth; }
        __syncthreads();
    }
    int nins = s_nins, tot_add = carry;
    if (tid == 0) pnew[nu] = tot_add;
    int e_old = g.c_rowptr[n], tot_new = e_old + tot_add;
    if (tot_new > C.ccap) { if (tid == 0) atomicExch(&cn[CN_ERROR], 13); return; }
    __syncthreads();
    // The receiving rows (first old entry, row id, new links before the row) are what every row and every entry
    // is located against below: they go to shared memory when there are at most 1024 of them.
    __shared__ int sm_start[1024], sm_row[1024], sm_before[1025];
    int32_t *ins_rowid = RP(C.ins_rowid, r, C.Pcap), *ins_before = RP(C.ins_before, r, C.Pcap + 1);
    for (int k = tid; k <= nins; k += blockDim.x) {
        int before = (k < nins) ? pnew[ins[k]] : tot_add;
        int row = (k < nins) ? (int)(pairs[ins[k]] >> 32) : n;
        ins_before[k] = before;
        if (k < nins) ins_rowid[k] = row;
        if (nins <= 1024) { sm_before[k] = before; if (k < nins) { sm_row[k] = row; sm_start[k] = ins_start[k]; } }
    }
    __syncthreads();
    const int *q_start = (nins <= 1024) ? sm_start : ins_start, *q_row = (nins <= 1024) ? sm_row : ins_rowid;
    const int *q_before = (nins <= 1024) ? sm_before : ins_before;
    // ---- new row pointers: a row moves right by the number of new links in the receiving rows before it
    {   // (a thread's rows ascend: the number of receiving rows before them only grows, so it is carried along instead of
        //  searched for every row -- the search was a third of this kernel, profiles/r03b_commit_*)
        int lo = 0;   // receiving rows with id < a
        for (int a = tid; a <= n; a += blockDim.x) {
            while (lo < nins && q_row[lo] < a) lo++;
            rp_new[a] = g.c_rowptr[a] + q_before[lo];
        }
    }
    // ---- entries of rows without pairs move as they are: one thread per old entry
    {
        int lo = 0;   // receiving rows whose first old entry is <= e (carried along like above)
        for (int e = tid; e < e_old; e += blockDim.x) {
            while (lo < nins && q_start[lo] <= e) lo++;
            if (lo > 0 && e < g.c_rowptr[q_row[lo - 1] + 1]) continue;     // inside a receiving row: merged below
            ent_new[e + q_before[lo]] = g.c_ent[e];
        }
    }
    // ---- receiving rows: one warp per row.  An old entry moves right by the new links of this row that sort
    // before it (and takes the new count if a pair names it); a new link goes behind the old entries that sort
    // before it and the new links before it.
    uint32_t *uent_rw = const_cast<uint32_t *>(g.u_ent);
    int lane = lane_id(), w = warp_id(), nw = blockDim.x >> 5;
    for (int k = w; k < nins; k += nw) {
        int pb = ins[k], pe = (k + 1 < nins) ? ins[k + 1] : nu;
        int a = q_row[k], ob = g.c_rowptr[a], oe = g.c_rowptr[a + 1], nb = ob + q_before[k];
        for (int i = ob + lane; i < oe; i += 32) {
            int2 ce = g.c_ent[i];
            int shift = 0;
            for (int j = pb; j < pe; j++) {
                int b = (int)(pairs[j] & 0xffffffffu), pc = pcnt[j];
                if (b < ce.x) shift += (pc & 0x40000000) ? 1 : 0;
                else if (b == ce.x) ce.y = (pc & 0x003fffff) | (ce.y & CE_ASYM) | (int)(CE_SMASK(ce.y) << 24);
            }
            ent_new[nb + (i - ob) + shift] = ce;
        }
        for (int j = pb + lane; j < pe; j += 32) {
            int pc = pcnt[j];
            if (!(pc & 0x40000000)) continue;
            int b = (int)(pairs[j] & 0xffffffffu);
            int lo = ob, hi = oe;   // old entries that sort before b
            while (lo < hi) { int m = (lo + hi) >> 1; if (g.c_ent[m].x < b) lo = m + 1; else hi = m; }
            int idx = static_find_idx(g, a, b);  // new link: copy the static mask of the pair and flag the static entry
            unsigned smask = 0;
            if (idx >= 0) { smask = ENT_MASK(uent_rw[idx]); atomicOr(&uent_rw[idx], ENT_CRIM); }
            // both rows of a new link see the same two static masks: one-sided (CE_ASYM) when their layer counts differ
            const int asym = (__popc(static_find_mask(g, b, a)) != __popc(smask)) ? CE_ASYM : 0;
            ent_new[nb + (lo - ob) + (pnew[j] - pnew[pb])] = make_int2(b, (pc & 0x003fffff) | asym | (int)(smask << 24));
        }
    }
    __syncthreads();
    if (tid == 0) C.c_par[r] = par ^ 1;   // publish: the other buffer is the criminal layer from here on
}

// ============================================================================ A9-A11 recruitment and arrests
// model.py:1452-1485, extra.py:102-132, 223-237, entities.py:573-603.  Order-defined and tiny (a few hundred
// list entries): one thread per replica follows the reference's sequence exactly.
__device__ double pw_sum_plain(const double *a, int n) {
    if (n < 8) { double r = 0.0; for (int i = 0; i < n; i++) r += a[i]; return r; }
    if (n <= 128) {
        double q[8];
        for (int j = 0; j < 8; j++) q[j] = a[j];
        int i;
        for (i = 8; i < n - (n % 8); i += 8) for (int j = 0; j < 8; j++) q[j] += a[i + j];
        double res = ((q[0] + q[1]) + (q[2] + q[3])) + ((q[4] + q[5]) + (q[6] + q[7]));
        for (; i < n; i++) res += a[i];
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return pw_sum_plain(a, n2) + pw_sum_plain(a + n2, n - n2);
}

// warp-cooperative version of weighted_cdf: coalesced passes, the two order-defined sums by warp_chain32
__device__ int warp_weighted_cdf(const double *w, int n, double *p, double *cdf) {
    int lane = lane_id();
    bool neg = false;
    double mn = INFINITY;
    for (int i = lane; i < n; i += 32) { double v = w[i]; neg |= (v < 0.0); mn = fmin(mn, v); }
    neg = __any_sync(FULL_MASK, neg);
    for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(FULL_MASK, mn, o));
    double sump = 0.0;
    int nz = 0;
    for (int base = 0; base < n; base += 32) {
        int i = base + lane;
        double v = 0.0;
        if (i < n) { v = neg ? (w[i] - mn) : w[i]; p[i] = v; }
        nz += __popc(__ballot_sync(FULL_MASK, v != 0.0));
        warp_chain32(v, sump);
    }
    if (sump == 0.0) return 0;
    double acc = 0.0;
    for (int base = 0; base < n; base += 32) {
        int i = base + lane;
        double q = (i < n) ? p[i] / sump : 0.0;
        double mine = warp_chain32(q, acc);
        if (i < n) { p[i] = q; cdf[i] = mine; }
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) cdf[i] = cdf[i] / acc;
    __syncwarp();
    return nz;
}

// one warp per replica; lane 0 walks the strictly sequential parts (recruitment order, the sampler, get_caught)
__device__ void recruit_arrest_warp(const DevCtx &C, int r, int tick) {
    int lane = lane_id();
    const poc_params &P = C.params[r];
    int *cn = RP(C.counters, r, CN_COUNT);
    int ncr = C.n_crimes[r];
    const int32_t *goff = RP(C.group_off, r, C.Ccap + 1), *gmem = RP(C.group_mem, r, C.Mcap);
    const int32_t *cflags = RP(C.crime_flags, r, C.Ccap);
    uint32_t *attr = RP(C.attr, r, C.Ncap);
    const int32_t *father = RP(C.father, r, C.Ncap);
    int32_t *new_recruit = RP(C.new_recruit, r, C.Ncap);
    double *arrest_w = RP(C.arrest_w, r, C.Ncap);
    if (cn[CN_ERROR] >= 11 && cn[CN_ERROR] <= 13) return;
    int nm = goff[ncr];
    // A9 recruitment, sequential over the OC-started groups (a father recruited earlier in the loop counts,
    // model.py:1456-1458); the groups are found 32 at a time
    int nrec = 0, noff = 0, nprot = 0;
    for (int base = 0; base < ncr; base += 32) {
        int g = base + lane;
        unsigned m = __ballot_sync(FULL_MASK, g < ncr && (cflags[g] & 1));
        while (m) {
            int gi = base + (__ffs(m) - 1);
            m &= m - 1;
            if (lane == 0) {
                for (int i = goff[gi]; i < goff[gi + 1]; i++) {
                    int a = gmem[i];
                    uint32_t x = attr[a];
                    if (x & F_OC) continue;
                    new_recruit[a] = tick;
                    attr[a] = x | F_OC;
                    nrec++;
                    int f = father[a];
                    if (f >= 0 && (attr[f] & F_OC)) noff++;
                    if (x & F_TARGET) nprot++;
                }
            }
            __syncwarp();
        }
    }
    if (lane == 0) {
        cn[CN_OFFSPRING_RECRUITED] += noff; cn[CN_PROTECTED_RECRUITED] += nprot;
        cn[CN_RECRUITED] += nrec; cn[CN_TICK_RECRUITED] = nrec;
    }
    __syncwarp();
    if (nm <= 0) return;
    // A10 arrest weights
    Graph g = graph_of(C, r);
    double *aw = RP(C.aw, r, C.Mcap), *ap = RP(C.ap, r, C.Mcap), *acdf = RP(C.acdf, r, C.Mcap);
    bool ion = (tick % P.ticks_between_intervention == 0) && P.intervention_start <= tick && tick < P.intervention_end;
    if (ion && P.facilitator_repression) {
        for (int i = lane; i < nm; i += 32) aw[i] = (attr[gmem[i]] & F_FAC) ? P.facilitator_repression_multiplier : 1.0;
    } else {
        bool anyoc = false;
        if (ion && P.oc_boss_repression) {
            for (int i = lane; i < nm; i += 32) anyoc |= (attr[gmem[i]] & F_OC) != 0;
            anyoc = __any_sync(FULL_MASK, anyoc);
        }
        if (anyoc) {
            // extra.calculate_oc_status over the OC members of the list (duplicates included, list order kept)
            int noc = 0;
            for (int base = 0; base < nm; base += 32) {
                int i = base + lane;
                bool isoc = false;
                double pos = 0.0;
                int a = -1;
                if (i < nm) {
                    a = gmem[i];
                    isoc = (attr[a] & F_OC) != 0;
                    if (isoc) {  // Person.calculate_oc_member_position, entities.py:573-580
                        int nn = 0, sumco = 0, cntoc = 0;
                        for (int e = g.u_rowptr[a]; e < g.u_rowptr[a + 1]; e++) {
                            uint32_t ent = g.u_ent[e];
                            if (ENT_MASK(ent) && (attr[ENT_COL(ent)] & F_OC)) nn++;
                        }
                        for (int e = g.c_rowptr[a]; e < g.c_rowptr[a + 1]; e++) {
                            int2 ce = g.c_ent[e];
                            if (CE_IS_DEAD(ce.y)) continue;
                            if (attr[ce.x] & F_OC) { sumco += CE_CNT(ce.y); cntoc++; if (!CE_SMASK(ce.y)) nn++; }
                        }
                        pos = (double)(nn + sumco - cntoc);
                    }
                }
                unsigned m = __ballot_sync(FULL_MASK, isoc);
                if (isoc) { ap[noc + __popc(m & ((1u << lane) - 1))] = pos; arrest_w[a] = pos; }
                else if (i < nm) arrest_w[a] = 1.0;
                noc += __popc(m);
            }
            __syncwarp();
            double mn = INFINITY;
            for (int i = lane; i < noc; i += 32) mn = fmin(mn, ap[i]);
            for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(FULL_MASK, mn, o));
            for (int i = lane; i < noc; i += 32) ap[i] = ap[i] - mn;
            __syncwarp();
            if (lane == 0) {
                double div = pw_sum_plain(ap, noc) / (double)noc;
                // third loop of calculate_oc_status in list order: a duplicated agent is normalised once per
                // occurrence, from its already-normalised value
                for (int i = 0; i < nm; i++) {
                    int a = gmem[i];
                    if (!(attr[a] & F_OC)) continue;
                    arrest_w[a] = (div > 0.0) ? (arrest_w[a] - mn) / div : 1.0;
                }
            }
            __syncwarp();
            for (int i = lane; i < nm; i += 32) aw[i] = arrest_w[gmem[i]];
        } else {
            for (int i = lane; i < nm; i += 32) aw[i] = 1.0;
        }
    }
    __syncwarp();
    for (int i = lane; i < nm; i += 32) arrest_w[gmem[i]] = aw[i];   // Person.arrest_weight (duplicates agree)
    double arrest_mod = P.arrests_per_year / (double)P.ticks_per_year / 10000.0 * (double)cn[CN_N_ALIVE];
    const int32_t *hdr = RP(C.rp_hdr, r, 8);
    bool replay = hdr[0] == 1;
    double ua, ub;
    if (replay) ua = C.rp_u_arrest_count[r]; else philox_u2(C.philox_key[r], tick, ST_ARREST_N, 0, ua, ub);
    int target = (ua < (arrest_mod - floor(arrest_mod))) ? (int)floor(arrest_mod + 1.0) : 0;  // Q7
    if (target <= 0) { if (lane == 0) { C.n_arrested[r] = 0; cn[CN_TICK_ARRESTS] = 0; } return; }
    int nz = warp_weighted_cdf(aw, nm, ap, acdf);
    if (nz < target) target = nz;
    if (lane != 0) return;
    int32_t *found = RP(C.arrested, r, C.Mcap);
    int nfound = 0;
    if (replay && hdr[4] >= 0) {
        const int32_t *ai = RP(C.rp_arrest_idx, r, C.Mcap);
        for (int i = 0; i < hdr[4] && i < target; i++) found[nfound++] = ai[i];
    } else {
        int ucur = 0;
        while (nfound < target) {  // numpy Generator.choice(replace=False, p) (R3)
            int need = target - nfound;
            if (nfound > 0) {
                for (int i = 0; i < nfound; i++) ap[found[i]] = 0.0;
                double acc = 0.0;
                for (int i = 0; i < nm; i++) { acc += ap[i]; acdf[i] = acc; }
                double last = acdf[nm - 1];
                for (int i = 0; i < nm; i++) acdf[i] /= last;
            }
            int base = nfound;
            for (int j = 0; j < need; j++) {
                double u, u2;
                if (replay) { if (ucur >= hdr[2]) { cn[CN_ERROR] = 4; u = 0.0; } else u = RP(C.rp_u_arrest, r, C.Mcap)[ucur]; }
                else philox_u2(C.philox_key[r], tick, ST_ARREST, ucur, u, u2);
                ucur++;
                int ix = ss_right(acdf, nm, u);
                if (ix >= nm) ix = nm - 1;
                bool dup = false;
                for (int q = base; q < nfound; q++) dup |= (found[q] == ix);
                if (!dup) found[nfound++] = ix;
            }
        }
    }
    // A11 get_caught
    double *sentence = RP(C.sentence, r, C.Ncap);
    uint32_t *uent = const_cast<uint32_t *>(g.u_ent);
    int2 *cent_rw = const_cast<int2 *>(g.c_ent);
    for (int i = 0; i < nfound; i++) {
        int a = gmem[found[i]];
        found[i] = a;  // report slots
        cn[CN_PEOPLE_JAILED] += 1;
        uint32_t x = attr[a] | F_PRISONER;
        double us, us2;
        if (replay) { if (i >= hdr[3]) { cn[CN_ERROR] = 5; us = 0.0; } else us = RP(C.rp_u_sentence, r, C.Mcap)[i]; }
        else philox_u2(C.philox_key[r], tick, ST_SENT, i, us, us2);
        int si = ss_right((x & F_MALE) ? P.sent_cdf_m : P.sent_cdf_f, P.n_sent, us);
        if (si >= P.n_sent) si = P.n_sent - 1;
        sentence[a] = P.sent_months[si] * P.punishment_length;
        if (x & F_HASJOB) x = (x & ~(F_HASJOB | (0xfu << 16))) | (1u << 16);
        x &= ~F_HASSCHOOL;
        attr[a] = x;
        int nprof = 0, nschool = 0;
        bool low_flag = false;
        for (int e = g.u_rowptr[a]; e < g.u_rowptr[a + 1]; e++) {  // Q11: own sets only
            uint32_t en = uent[e];
            nprof += (ENT_MASK(en) & MB_PROF) != 0; nschool += (ENT_MASK(en) & MB_SCHOOL) != 0;
            if (ENT_MASK(en) & (MB_PROF | MB_SCHOOL)) {
                // the pair loses a layer on this side only: it no longer mirrors itself (poc_device.cuh: ENT_ASYM).  The
                // reverse entries are flagged by the whole block afterwards (k_recruit_arrest)
                en = (en & ~((MB_PROF | MB_SCHOOL) << 24)) | ENT_ASYM;
                low_flag |= ENT_COL(en) < a;
                uent[e] = en;
            }
        }
        if (low_flag) g.u_mid[a] = g.u_rowptr[a];
        if (x & F_ALIVE) { RP(C.layer_tot, r, 8)[6] -= nprof; RP(C.layer_tot, r, 8)[7] -= nschool; }
        for (int e = g.c_rowptr[a]; e < g.c_rowptr[a + 1]; e++) {
            const int y = cent_rw[e].y;
            if (CE_SMASK(y) & (MB_PROF | MB_SCHOOL)) cent_rw[e].y = (y & ~(int)((MB_PROF | MB_SCHOOL) << 24)) | CE_ASYM;
        }
    }
    C.n_arrested[r] = nfound;
    cn[CN_TICK_ARRESTS] = nfound;
}

// end_tick: 0 = arrests only; 1 = also the prisoner countdown of model.py:279-283 (k_end_tick fused);
// 2 = also advance the replica group's device-side tick state once the last block of the launch is through
__global__ void __launch_bounds__(256, 4) k_recruit_arrest(DevCtx C, int end_tick) {
    int tick = C.ti[C.grp].tick;
    int r = C.r0 + blockIdx.x;
    if (warp_id() == 0) recruit_arrest_warp(C, r, tick);
    __syncthreads();
    {   // get_caught flagged the entries it changed in the arrested agents' rows: flag their reverse entries (warp per agent)
        Graph g = graph_of(C, r);
        const int na = min(C.n_arrested[r], C.Mcap);
        const int32_t *arr = RP(C.arrested, r, C.Mcap);
        for (int i = warp_id(); i < na; i += (int)(blockDim.x >> 5)) {
            const int a = arr[i];
            for (int e = g.u_rowptr[a] + lane_id(); e < g.u_rowptr[a + 1]; e += 32) {
                const uint32_t en = g.u_ent[e];
                if (!(en & ENT_ASYM)) continue;
                const int ri = static_find_idx(g, ENT_COL(en), a);
                if (ri >= 0) ent_flag_asym(g, ENT_COL(en), ri);
            }
            int2 *cent = const_cast<int2 *>(g.c_ent);
            for (int e = g.c_rowptr[a] + lane_id(); e < g.c_rowptr[a + 1]; e += 32) {
                const int2 ce = cent[e];
                if (!(ce.y & CE_ASYM) || CE_IS_DEAD(ce.y)) continue;
                const int ri = crim_find_idx(g, ce.x, a);
                if (ri >= 0) atomicOr(&cent[ri].y, CE_ASYM);
            }
        }
    }
    if (!end_tick) return;
    __syncthreads();
    int n = C.n_agents[r];
    uint32_t *attr = RP(C.attr, r, C.Ncap);
    double *sentence = RP(C.sentence, r, C.Ncap);
    for (int a = threadIdx.x; a < n; a += blockDim.x) {
        uint32_t x = attr[a];
        if ((x & (F_ALIVE | F_PRISONER)) != (F_ALIVE | F_PRISONER)) continue;
        double s = sentence[a] - 1.0;
        sentence[a] = s;
        if (s == 0.0) attr[a] = x & ~F_PRISONER;
    }
    if (end_tick == 2) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            TickInfo &t = C.ti[C.grp];
            if (atomicAdd(&t.pad, 1) == (int)gridDim.x - 1) { t.pad = 0; t.tick++; t.epoch++; t.rep++; __threadfence(); }
        }
    }
}

// ============================================================================ reporters (calculate_fast_reporters subset)
// model.py:1687-1735, the columns the hot path feeds.  One block per replica.
// mode 0/1: write the row (1: the tick state has already moved on); 2: write the row, then the last block moves the
// replica group's tick state on; 3: only move the tick state on
__global__ void __launch_bounds__(512, 3) k_report(DevCtx C, int mode) {
    int t = C.ti[C.grp].rep - (mode == 1 ? 1 : 0);
    if (mode == 3) {
        if (threadIdx.x == 0) { TickInfo &ti = C.ti[C.grp]; if (atomicAdd(&ti.pad, 1) == (int)gridDim.x - 1) { ti.pad = 0; ti.tick++; ti.epoch++; ti.rep++; } }
        return;
    }
    int r = C.r0 + blockIdx.x, n = C.n_agents[r], tid = threadIdx.x;
    const uint32_t *attr = RP(C.attr, r, C.Ncap);
    const int32_t *ncr = RP(C.ncrimes, r, C.Ncap), *nct = RP(C.ncrimes_tick, r, C.Ncap);
    const int32_t *crp = graph_of(C, r).c_rowptr;
    const int32_t *ncd = RP(C.ncrim_dead, r, C.Ncap);
    const double *tend = RP(C.tendency, r, C.Ncap);
    enum { I_OC = 0, I_PRIS, I_LINK, I_SUMCR, I_OCOC, I_ALIVE, I_AGE, I_AGE2, I_EDU, I_EDU2, I_CR2, I_EMPL, I_FAC, I_STUD, NI };
    long long vi[NI];
#pragma unroll
    for (int k = 0; k < NI; k++) vi[k] = 0;
    double vd[2] = { 0.0, 0.0 };
    for (int a = tid; a < n; a += blockDim.x) {
        uint32_t x = attr[a];
        if (!(x & F_ALIVE)) continue;
        long long age = attr_age(x), edu = attr_edu(x), c = ncr[a];
        vi[I_ALIVE]++;
        if (x & F_OC) { vi[I_OC]++; vi[I_OCOC] += nct[a]; }
        vi[I_PRIS] += (x & F_PRISONER) != 0; vi[I_EMPL] += (x & F_HASJOB) != 0; vi[I_FAC] += (x & F_FAC) != 0;
        vi[I_STUD] += (x & F_HASSCHOOL) != 0;
        vi[I_LINK] += crp[a + 1] - crp[a] - ncd[a];
        vi[I_SUMCR] += c; vi[I_CR2] += c * c; vi[I_AGE] += age; vi[I_AGE2] += age * age; vi[I_EDU] += edu; vi[I_EDU2] += edu * edu;
        double tv = tend[a];
        vd[0] += tv; vd[1] += tv * tv;
    }
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < NI; k++) vi[k] += __shfl_xor_sync(FULL_MASK, vi[k], o);
        for (int k = 0; k < 2; k++) vd[k] += __shfl_xor_sync(FULL_MASK, vd[k], o);
    }
    __shared__ long long wi[16][NI];
    __shared__ double wd[16][2];
    __shared__ double s_mean[4], s_ss[4];   // tendency, age, education level, crimes committed: mean and sum of squared deviations
    __shared__ double wq[16][4];
    if (lane_id() == 0) { for (int k = 0; k < NI; k++) wi[warp_id()][k] = vi[k]; for (int k = 0; k < 2; k++) wd[warp_id()][k] = vd[k]; }
    __syncthreads();
    const int nw = blockDim.x >> 5;
    if (tid == 0) {
        for (int ww = 1; ww < nw; ww++) { for (int k = 0; k < NI; k++) vi[k] += wi[ww][k]; for (int k = 0; k < 2; k++) vd[k] += wd[ww][k]; }
        double nn = (double)vi[I_ALIVE], inv = nn > 0 ? 1.0 / nn : 0.0;
        s_mean[0] = vd[0] * inv; s_mean[1] = (double)vi[I_AGE] * inv; s_mean[2] = (double)vi[I_EDU] * inv; s_mean[3] = (double)vi[I_SUMCR] * inv;
        s_ss[0] = s_ss[1] = s_ss[2] = s_ss[3] = 0.0;
    }
    __syncthreads();
    {   // second pass: np.std is sqrt(mean((x - mean)^2)) (model.py:1698-1709); E[x^2] - mean^2 would cancel
        double q[4] = { 0.0, 0.0, 0.0, 0.0 };
        const double m0 = s_mean[0], m1 = s_mean[1], m2 = s_mean[2], m3 = s_mean[3];
        for (int a = tid; a < n; a += blockDim.x) {
            uint32_t x = attr[a];
            if (!(x & F_ALIVE)) continue;
            double d0 = tend[a] - m0, d1 = (double)attr_age(x) - m1, d2 = (double)attr_edu(x) - m2, d3 = (double)ncr[a] - m3;
            q[0] += d0 * d0; q[1] += d1 * d1; q[2] += d2 * d2; q[3] += d3 * d3;
        }
        for (int o = 16; o > 0; o >>= 1)
            for (int k = 0; k < 4; k++) q[k] += __shfl_xor_sync(FULL_MASK, q[k], o);
        if (lane_id() == 0) for (int k = 0; k < 4; k++) wq[warp_id()][k] = q[k];
    }
    __syncthreads();
    if (tid == 0) {
        // warp partials added in warp order: the row is the same bits run after run (a floating-point atomicAdd is not)
        for (int ww = 0; ww < nw; ww++) for (int k = 0; k < 4; k++) s_ss[k] += wq[ww][k];
        const int *cn = RP(C.counters, r, CN_COUNT);
        const int *pv = RP(C.prev_counters, r, CN_COUNT);
        const long long *lt = RP(C.layer_tot, r, 8);
        double *row = C.reporters + ((size_t)r * C.Tcap + t) * POC_N_REPORTERS;
        double nn = (double)vi[I_ALIVE], inv = nn > 0 ? 1.0 / nn : 0.0;
        auto sd = [&](int k) { double v = s_ss[k] * inv; return v > 0 ? sqrt(v) : 0.0; };
        row[0] = cn[CN_NUMBER_CRIMES]; row[1] = cn[CN_SIZE_FAILS]; row[2] = cn[CN_FAC_CRIMES]; row[3] = cn[CN_FAC_FAILS];
        row[4] = cn[CN_BIG_CRIME]; row[5] = cn[CN_OFFSPRING_RECRUITED]; row[6] = cn[CN_PEOPLE_JAILED];
        row[7] = (double)vi[I_OC]; row[8] = (double)vi[I_PRIS]; row[9] = (double)vi[I_LINK]; row[10] = (double)vi[I_SUMCR];
        row[11] = (double)vi[I_OCOC]; row[12] = s_mean[0]; row[13] = sd(0); row[14] = nn;
        row[15] = C.crime_mult[r];
        row[16] = s_mean[1]; row[17] = sd(1);
        row[18] = s_mean[2]; row[19] = sd(2);
        row[20] = s_mean[3]; row[21] = sd(3);
        row[22] = (double)vi[I_EMPL]; row[23] = (double)vi[I_FAC]; row[24] = (double)vi[I_STUD];
        // sibling, offspring, parent, partner, household, friendship, professional, school
        for (int l = 0; l < 8; l++) row[25 + l] = (double)lt[l];
        row[33] = cn[CN_PROTECTED_RECRUITED]; row[34] = cn[CN_PEOPLE_JAILED] - pv[CN_PEOPLE_JAILED];
        row[35] = C.ti[C.grp].tick; row[36] = cn[CN_RECRUITED]; row[37] = cn[CN_TICK_CRIMES]; row[38] = cn[CN_TICK_EMB]; row[39] = cn[CN_ERROR];
        row[40] = cn[CN_BORN]; row[41] = cn[CN_DECEASED]; row[42] = cn[CN_WEDDINGS];   // the step phases, when they run on the device
        if (mode == 2) {
            __threadfence();
            TickInfo &ti = C.ti[C.grp];
            if (atomicAdd(&ti.pad, 1) == (int)gridDim.x - 1) { ti.pad = 0; ti.tick++; ti.epoch++; ti.rep++; }
        }
    }
}

// ============================================================================ stage hooks
__global__ void __launch_bounds__(256) k_rings_hook(DevCtx C, int r, int nsrc, const int32_t *src, int d, int exact,
                                                    int32_t *out_cnt, int32_t *out_mem, long long cap_per_src) {
    extern __shared__ unsigned smem_u[];
    __shared__ int s_count, lvl_off[MAX_RADIUS + 3];
    int n = C.n_agents[r], nwords = (n + 31) >> 5;
    int *list = C.scr_list + (size_t)blockIdx.x * (C.Ncap + 1);
    Graph g = graph_of(C, r);
    unsigned long long edges = 0, rb = 0;
    for (int s = blockIdx.x; s < nsrc; s += gridDim.x) {
        if (threadIdx.x == 0) list[0] = src[s];
        __syncthreads();
        int total = bfs_levels(g, nwords, smem_u, list, 1, d, lvl_off, &s_count, edges, rb);
        int b = exact ? lvl_off[d] : 1, m = total - b;
        if (threadIdx.x == 0) out_cnt[s] = m;
        for (int i = threadIdx.x; i < m && i < cap_per_src; i += blockDim.x) out_mem[(size_t)s * cap_per_src + i] = list[b + i];
        __syncthreads();
    }
}

// warp per pair: social proximity, and candidate weight (needs emb_val for OC `self`)
__global__ void k_pairs_hook(DevCtx C, int r, int npairs, const int32_t *sa, const int32_t *sb, double *out, int mode) {
    int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = lane_id();
    if (gw >= npairs) return;
    Graph g = graph_of(C, r);
    const uint32_t *attr = RP(C.attr, r, C.Ncap);
    int a = sa[gw], b = sb[gw];
    // friendship intersection by merging the two sorted rows: lane-strided over a's friends, binary search in b
    bool hit = false;
    for (int e = g.u_rowptr[a] + lane; e < g.u_rowptr[a + 1]; e += 32) {
        uint32_t ent = g.u_ent[e];
        if ((ENT_MASK(ent) & MB_FRIEND) && (static_find_mask(g, b, ENT_COL(ent)) & MB_FRIEND)) hit = true;
    }
    hit = __any_sync(FULL_MASK, hit);
    if (lane == 0) {
        double prox = social_proximity(attr[a], attr[b], hit);
        if (mode == 0) out[gw] = prox;
        else {
            double t = RP(C.tendency, r, C.Ncap)[b];
            out[gw] = (attr[a] & F_OC) ? -1.0 * ((prox * RP(C.emb_val, r, C.Ncap)[b]) * t) : prox * t;
        }
    }
}

// stage hook for Person.find_accomplices: n independent searches (crime slots 0..n-1 of replica r)
__global__ void k_hook_init_crimes(DevCtx C, int r, int n, const int32_t *init, const int32_t *k) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j == 0) { C.n_crimes[r] = n; C.n_search[r] = 0; C.n_emb[r] = 0; C.n_emb[C.R + r] = 0; }
    if (j >= n) return;
    const poc_params &P = C.params[r];
    size_t o = (size_t)r * C.Ccap + j;
    uint32_t ia = RP(C.attr, r, C.Ncap)[init[j]];
    int kk = k[j];
    C.crime_init[o] = init[j]; C.crime_k[o] = kk; C.crime_flags[o] = (ia & F_OC) ? 1 : 0;
    int target = kk, fn = 0;
    if (kk > 0) { fn = ((double)kk >= P.threshold_use_facilitators) && !(ia & F_FAC); if (fn) target -= 1; }
    C.cr_target[o] = target; C.cr_nacc[o] = 0; C.cr_d[o] = 1; C.cr_fails[o] = 0;
    C.cr_state[o] = (kk > 0) ? (fn ? 2 : 0) : 4;
}
__global__ void k_hook_fill_search(DevCtx C, int r, int n) {  // deterministic search list 0..n-1 with k > 0
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int m = 0;
        for (int j = 0; j < n; j++) if (C.crime_k[(size_t)r * C.Ccap + j] > 0) RP(C.search_list, r, C.Ccap)[m++] = j;
        C.n_search[r] = m;
    }
}
__global__ void k_snapshot_counters(DevCtx C, int *prev) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < C.Rg * CN_COUNT) prev[C.r0 * CN_COUNT + i] = C.counters[C.r0 * CN_COUNT + i];
}
