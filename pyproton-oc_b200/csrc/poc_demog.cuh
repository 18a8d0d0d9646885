// poc_demog.cuh -- SURVEY.md 8(f): the three phases of ProtonOC.step that write the multiplex network, on the device.
//
//   make_baby          model.py:1553-1573 (non-constant population), entities.py:605-613, 628-650
//   make_friends       model.py:610-626, entities.py:246-254, 674-689, extra.py:102-132
//   make_people_die    model.py:1575-1602, entities.py:615-626, 652-661
//
// The reference walks the agents one after the other on one shared PCG64 stream; what runs here is the BATCH semantics
// oracle/hotpath.c restates (every agent of a phase decides on the state the phase started from, with its own Philox
// draws; the edits are applied together), which the kernels follow bit for bit.  Parity with the reference itself is
// distributional (tests: KS against reference runs with exactly these phases active).
//
// Data: deaths clear layer masks / leave tombstones in criminal rows (entries never move); friendships set a mask bit on
// entries that already exist (a potential friend is a relative, schoolmate or colleague: already a neighbour); births
// need new entries, so the multiplex rows are laid out again into the second buffer (k_multiplex_insert: sort the new
// directed entries -> merge, like k_commit does for the criminal layer).
#pragma once
#include "poc_device.cuh"

#define DM_KEY(u, v, m) (((unsigned long long)(unsigned)(u) << 40) | ((unsigned long long)(unsigned)(v) << 8) | (unsigned long long)(m))
#define DM_U(k) ((int)(((k) >> 40) & 0x7fffffu))
#define DM_V(k) ((int)(((k) >> 8) & 0x7fffffu))
#define DM_M(k) ((unsigned)((k) & 0xffu))
#define FRIEND_CAND_MAX 64   /* CANON: the 64 lowest slots among an agent's potential friends are considered */

// a criminal entry carries a copy of the static layer mask of the same neighbour (poc_device.cuh): keep it current
__device__ __forceinline__ void crim_mask_or(const Graph &g, int u, int v, unsigned bits) {
    const int ci = crim_find_idx(g, u, v);
    if (ci >= 0 && !CE_IS_DEAD(g.c_ent[ci].y)) {
        // the pair may no longer weigh the same from both rows: both entries stand for themselves from now on (CE_ASYM)
        atomicOr(&const_cast<int2 *>(g.c_ent)[ci].y, (int)(bits << 24) | CE_ASYM);
        const int cj = crim_find_idx(g, v, u);
        if (cj >= 0) atomicOr(&const_cast<int2 *>(g.c_ent)[cj].y, CE_ASYM);
    }
}

// ---------------------------------------------------------------------------------------------- make_baby
// one block per replica: eligible mothers draw, the births are listed in ascending slot order of the mother, every birth
// fills its slot and queues its directed entries
__global__ void __launch_bounds__(1024) k_make_baby(DevCtx C) {
    const int tick = C.ti[C.grp].tick;
    const int r = C.r0 + blockIdx.x, tid = threadIdx.x, n = C.n_agents[r];
    __shared__ int ws[33];
    __shared__ int carry, s_np;
    const poc_params &P = C.params[r];
    const DevDemog &D = *C.demog;
    uint32_t *attr = RP(C.attr, r, C.Ncap);
    int32_t *nch = RP(C.n_children, r, C.Ncap), *list = RP(C.dm_list, r, C.Ncap);
    int *cn = RP(C.counters, r, CN_COUNT);
    Graph g = graph_of(C, r);
    const unsigned long long key = C.philox_key[r];
    if (tid == 0) { carry = 0; s_np = 0; }
    __syncthreads();
    for (int base = 0; base < n; base += blockDim.x) {
        const int a = base + tid;
        bool born = false;
        if (a < n) {
            const uint32_t x = attr[a];
            const int age = attr_age(x);
            if ((x & F_ALIVE) && !(x & F_MALE) && age >= 14 && age <= 50) {
                double u, u2;
                philox_u2(key, tick, ST_BABY, a, u, u2);
                const int nc = min(nch[a], 2);
                const double pf = (age < 64) ? D.fert[nc][age] / (double)P.ticks_per_year : 0.0;
                born = u < pf;
            }
        }
        int tot, ex = block_exscan(born ? 1 : 0, ws, &tot);
        if (born && carry + ex < C.Ncap) list[carry + ex] = a;
        __syncthreads();
        if (tid == 0) carry += tot;
        __syncthreads();
    }
    const int nb = carry;
    if (tid == 0) { RP(C.dm_n, r, 4)[0] = nb; RP(C.dm_n, r, 4)[2] = 0; if (nb) C.cdf_valid[r] = 0; }
    if (nb == 0) return;
    if (n + nb > C.Ncap) { if (tid == 0) { atomicExch(&cn[CN_ERROR], 20); RP(C.dm_n, r, 4)[0] = 0; } return; }
    unsigned long long *pairs = RP(C.dm_pairs, r, C.DPcap);
    for (int i = tid; i < nb; i += blockDim.x) {
        const int m = list[i], b = n + i;
        const size_t o = (size_t)r * C.Ncap + b;
        double u0, u1;
        philox_u2(key, tick, ST_BABYATTR, m, u0, u1);
        int qi = (int)floor(u1 * (double)D.n_prop);
        if (qi >= D.n_prop) qi = D.n_prop - 1;
        const uint32_t xm = attr[m];
        // Person.__init__ (entities.py:49-83) + init_baby: age 0, the mother's wealth level, everything else at its default
        attr[b] = F_ALIVE | (u0 < 0.5 ? F_MALE : 0u) | ((uint32_t)attr_wealth(xm) << 24);
        C.birth[o] = tick; C.ncrimes[o] = 0; C.ncrimes_tick[o] = 0; C.new_recruit[o] = -2;
        C.propensity[o] = D.prop_q[qi]; C.tendency[o] = 0.0; C.sentence[o] = 0.0; C.arrest_w[o] = 0.0;
        C.emb_stamp[o] = 0; C.ncrim_dead[o] = 0; nch[b] = 0;
        // the mother's row as the phase found it: children so far (siblings), partner (father), household
        int f = -1, cnt = 0;
        const int e0 = g.u_rowptr[m], e1 = g.u_rowptr[m + 1];
        for (int e = e0; e < e1; e++) {
            const uint32_t ent = g.u_ent[e];
            const unsigned mk = ENT_MASK(ent);
            if ((mk & MB_PARTNER) && f < 0) f = ENT_COL(ent);   // CANON: the partner with the smallest slot
            cnt += ((mk & MB_OFF) ? 2 : 0) + ((mk & MB_HH) ? 2 : 0);
        }
        cnt += 2 + (f >= 0 ? 2 : 0);
        C.father[o] = f;
        {   // max_education_level of the father, else of the mother (entities.py:646-649); not at school
            uint8_t *e2 = RP(C.edu2, r, C.Ncap);
            e2[b] = e2[f >= 0 ? f : m] & 15u;
        }
        int pos = atomicAdd(&s_np, cnt);
        if (pos + cnt <= C.DPcap) {
            for (int e = e0; e < e1; e++) {
                const uint32_t ent = g.u_ent[e];
                const unsigned mk = ENT_MASK(ent);
                const int v = ENT_COL(ent);
                if (mk & MB_OFF) { pairs[pos++] = DM_KEY(b, v, MB_SIB); pairs[pos++] = DM_KEY(v, b, MB_SIB); }
                if (mk & MB_HH) { pairs[pos++] = DM_KEY(b, v, MB_HH); pairs[pos++] = DM_KEY(v, b, MB_HH); }
            }
            pairs[pos++] = DM_KEY(m, b, MB_OFF); pairs[pos++] = DM_KEY(b, m, MB_PAR);
            if (f >= 0) { pairs[pos++] = DM_KEY(f, b, MB_OFF); pairs[pos++] = DM_KEY(b, f, MB_PAR); }
        }
        nch[m] += 1;
    }
    __syncthreads();
    if (tid == 0) {
        if (s_np > C.DPcap) { atomicExch(&cn[CN_ERROR], 21); RP(C.dm_n, r, 4)[0] = 0; return; }
        RP(C.dm_n, r, 4)[2] = s_np;
        cn[CN_BORN] += nb;
    }
}

// The queued directed entries {u, v, layer bit} go into the multiplex rows: an entry that exists gets the bit, a new one
// is inserted in slot order.  Everything is laid out again in the other buffer (rows keep no slack); one block per replica.
// n_new = number of agent slots afterwards (newborn rows start empty).
__global__ void __launch_bounds__(1024) k_multiplex_insert(DevCtx C) {
    const int r = C.r0 + blockIdx.x, tid = threadIdx.x;
    int *dmn = RP(C.dm_n, r, 4);
    const int nb = dmn[0], np = dmn[2];
    if (nb == 0 && np == 0) return;
    __shared__ int ws[33];
    __shared__ int carry, s_nins;
    int *cn = RP(C.counters, r, CN_COUNT);
    const int n_old = C.n_agents[r], n_new = n_old + nb;
    unsigned long long *pairs = RP(C.dm_pairs, r, C.DPcap);
    Graph g = graph_of(C, r);
    uint32_t *ent_old = const_cast<uint32_t *>(g.u_ent);
    const uint32_t *attr = RP(C.attr, r, C.Ncap);
    long long *ltot = RP(C.layer_tot, r, 8);
    // ---- bitonic sort by (u, v)
    int np2 = 1;
    while (np2 < np) np2 <<= 1;
    if (np2 > C.DPcap) { if (tid == 0) atomicExch(&cn[CN_ERROR], 21); return; }
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
    // ---- one entry per (u, v): OR the layer bits of a run into its first element, compact
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < np; base += blockDim.x) {
        const int i = base + tid;
        const bool valid = i < np;
        unsigned long long me = valid ? pairs[i] : 0ull;
        const bool head = valid && (i == 0 || (pairs[i - 1] >> 8) != (me >> 8));
        if (head) for (int j = i + 1; j < np && (pairs[j] >> 8) == (me >> 8); j++) me |= pairs[j] & 0xffull;
        int tot, ex = block_exscan(head ? 1 : 0, ws, &tot);
        __syncthreads();
        if (head) pairs[carry + ex] = me;
        __syncthreads();
        if (tid == 0) carry += tot;
        __syncthreads();
    }
    const int nu = carry;
    // ---- an entry that exists already takes the bits in place (the copy below carries it over); the rest are new
    int32_t *isnew_ex = RP(C.pair_new_ex, r, C.Pcap + 1), *ins = RP(C.ins_row, r, C.Pcap);
    int32_t *ins_rowid = RP(C.ins_rowid, r, C.Pcap), *ins_before = RP(C.ins_before, r, C.Pcap + 1);
    if (nu > C.Pcap) { if (tid == 0) atomicExch(&cn[CN_ERROR], 21); return; }
    for (int i = tid; i < nu; i += blockDim.x) {
        const unsigned long long k = pairs[i];
        const int u = DM_U(k), v = DM_V(k);
        const unsigned m = DM_M(k);
        const int idx = (u < n_old) ? static_find_idx(g, u, v) : -1;
        const bool ualive = (attr[u] & F_ALIVE) != 0;
        if (idx >= 0) {
            const unsigned old = atomicOr(&ent_old[idx], m << 24);
            if (old & ENT_CRIM) crim_mask_or(g, u, v, m);
            if (ualive) for (int l = 0; l < 8; l++) if (((m >> l) & 1u) && !((old >> (24 + l)) & 1u)) atomicAdd((unsigned long long *)&ltot[l], 1ull);
        } else {
            if (ualive) for (int l = 0; l < 8; l++) if ((m >> l) & 1u) atomicAdd((unsigned long long *)&ltot[l], 1ull);
            pairs[i] = k | (1ull << 63);   // new entry
        }
    }
    if (tid == 0) { carry = 0; s_nins = 0; }
    __syncthreads();
    // number of new entries before each pair; first new pair of every row that receives any
    for (int base = 0; base < nu; base += blockDim.x) {
        const int i = base + tid;
        const bool valid = i < nu;
        const int isnew = (valid && (pairs[i] >> 63)) ? 1 : 0;
        bool head = isnew != 0;
        for (int j = i - 1; head && j >= 0 && DM_U(pairs[j]) == DM_U(pairs[i]); j--) if (pairs[j] >> 63) head = false;
        int tot, ex = block_exscan(isnew, ws, &tot);
        if (valid) isnew_ex[i] = carry + ex;
        int toth, exh = block_exscan(head ? 1 : 0, ws, &toth);
        if (head) ins[s_nins + exh] = i;
        __syncthreads();
        if (tid == 0) { carry += tot; s_nins += toth; }
        __syncthreads();
    }
    const int nins = s_nins, tot_add = carry;
    if (tid == 0) isnew_ex[nu] = tot_add;
    const int e_old = g.u_rowptr[n_old];
    if ((long long)e_old + tot_add > C.ucap) { if (tid == 0) atomicExch(&cn[CN_ERROR], 22); return; }
    __syncthreads();
    // receiving rows: id, new entries in the receiving rows before it
    for (int k = tid; k <= nins; k += blockDim.x) {
        ins_before[k] = (k < nins) ? isnew_ex[ins[k]] : tot_add;
        if (k < nins) ins_rowid[k] = DM_U(pairs[ins[k]]);
    }
    __syncthreads();
    const int upar = C.u_par[r];
    int32_t *rp_new = (upar ? C.u_rowptr : C.u_rowptr_tmp) + (size_t)r * (C.Ncap + 1);
    uint32_t *ent_new = (upar ? C.u_ent : C.u_ent_tmp) + (size_t)r * C.ucap;
    // ---- new row pointers (rows of newborns start behind the old rows)
    for (int a = tid; a <= n_new; a += blockDim.x) {
        int lo = 0, hi = nins;   // receiving rows with id < a
        while (lo < hi) { int m = (lo + hi) >> 1; if (ins_rowid[m] < a) lo = m + 1; else hi = m; }
        rp_new[a] = ((a <= n_old) ? g.u_rowptr[a] : e_old) + ins_before[lo];
    }
    // ---- rows without new entries move as they are: the stretch of rows between two receiving rows shifts by one amount,
    // so it is copied entry-parallel (coalesced); receiving rows are merged by one warp each
    const int lane = lane_id(), w = warp_id(), nw = blockDim.x >> 5;
    for (int k = 0; k <= nins; k++) {
        const int ra = (k == 0) ? 0 : ins_rowid[k - 1] + 1, rb = (k < nins) ? min(ins_rowid[k], n_old) : n_old;   // rows [ra, rb)
        if (ra >= rb || ra >= n_old) continue;
        const int b = g.u_rowptr[ra], e = g.u_rowptr[rb], sh = ins_before[k];
        for (int i = b + tid; i < e; i += blockDim.x) ent_new[i + sh] = ent_old[i];
    }
    for (int k = w; k < nins; k += nw) {
        const int a = ins_rowid[k];
        const int ob = (a < n_old) ? g.u_rowptr[a] : e_old, oe = (a < n_old) ? g.u_rowptr[a + 1] : e_old, nbase = ob + ins_before[k];
        // the pairs of row a: from ins[k] to the end of the row's run
        const int pb = ins[k];
        int pe = pb;
        while (pe < nu && DM_U(pairs[pe]) == a) pe++;
        for (int i = ob + lane; i < oe; i += 32) {
            const uint32_t en = ent_old[i];
            int shift = 0;
            for (int j = pb; j < pe; j++) if ((pairs[j] >> 63) && DM_V(pairs[j]) < ENT_COL(en)) shift++;
            ent_new[nbase + (i - ob) + shift] = en;
        }
        for (int j = pb + lane; j < pe; j += 32) {
            if (!(pairs[j] >> 63)) continue;
            const unsigned long long kk = pairs[j];
            const int v = DM_V(kk);
            int lo = ob, hi = oe;   // old entries that sort before v
            while (lo < hi) { int m = (lo + hi) >> 1; if (ENT_COL(ent_old[m]) < v) lo = m + 1; else hi = m; }
            // a criminal entry towards v may exist already (the flag lives in the static entry)
            const unsigned crim = (a < n_old && crim_find(g, a, v) >= 0) ? ENT_CRIM : 0u;
            if (crim) crim_mask_or(g, a, v, DM_M(kk));
            ent_new[nbase + (lo - ob) + (isnew_ex[j] - isnew_ex[pb])] = (uint32_t)v | crim | (DM_M(kk) << 24);
        }
    }
    // newborns have empty criminal rows
    {
        int32_t *crp = const_cast<int32_t *>(g.c_rowptr);
        const int ctot = crp[n_old];
        for (int a = n_old + 1 + tid; a <= n_new; a += blockDim.x) crp[a] = ctot;
    }
    __syncthreads();
    if (tid == 0) { C.u_par[r] = upar ^ 1; C.n_agents[r] = n_new; dmn[0] = 0; dmn[2] = 0; }
    __syncthreads();
    // the pairs that changed may no longer mirror their reverse entries (poc_device.cuh: ENT_ASYM); rows as laid out now
    {
        Graph gn = g;
        gn.u_rowptr = rp_new; gn.u_ent = ent_new;
        gn.u_mid = (upar ? C.u_mid : C.u_mid_tmp) + (size_t)r * C.Ncap;
        // every row moved: a row without new entries keeps its pointer relative to its begin, a receiving row is read again
        for (int a = tid; a < n_new; a += blockDim.x) {
            int lo = 0, hi = nins;
            while (lo < hi) { int m = (lo + hi) >> 1; if (ins_rowid[m] < a) lo = m + 1; else hi = m; }
            const bool receiving = (lo < nins && ins_rowid[lo] == a) || a >= n_old;
            gn.u_mid[a] = receiving ? row_mid(gn, a) : rp_new[a] + (g.u_mid[a] - g.u_rowptr[a]);
        }
        __syncthreads();
        for (int i = tid; i < nu; i += blockDim.x) ent_asym_update(gn, DM_U(pairs[i]), DM_V(pairs[i]));
    }
}

// ---------------------------------------------------------------------------------------------- make_friends
// social_proximity's last term: does any friend of a also count c among its friends?  Both rows ascending.
__device__ __forceinline__ bool friends_shared_rows(const Graph &g, int a, int c) {
    int i = g.u_rowptr[a], ie = g.u_rowptr[a + 1], j = g.u_rowptr[c], je = g.u_rowptr[c + 1];
    while (i < ie && j < je) {
        const uint32_t x = g.u_ent[i], y = g.u_ent[j];
        if (!(ENT_MASK(x) & MB_FRIEND)) { i++; continue; }
        if (!(ENT_MASK(y) & MB_FRIEND)) { j++; continue; }
        const int cx = ENT_COL(x), cy = ENT_COL(y);
        if (cx == cy) return true;
        if (cx < cy) i++; else j++;
    }
    return false;
}

// pass 1 (thread per agent, decisions on the state the phase started from): the chosen pairs are queued
__global__ void __launch_bounds__(128) k_make_friends(DevCtx C) {
    const int tick = C.ti[C.grp].tick;
    const int r = C.r0 + blockIdx.y, a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= C.n_agents[r]) return;
    const uint32_t *attr = RP(C.attr, r, C.Ncap);
    const uint32_t xa = attr[a];
    if (!(xa & F_ALIVE)) return;
    const DevDemog &D = *C.demog;
    Graph g = graph_of(C, r);
    const unsigned long long key = C.philox_key[r];
    int cand[FRIEND_CAND_MAX];
    double w[FRIEND_CAND_MAX], cdf[FRIEND_CAND_MAX];
    int m = 0;
    const unsigned src = MB_SIB | MB_OFF | MB_PARTNER | MB_SCHOOL | MB_PROF;
    for (int e = g.u_rowptr[a]; e < g.u_rowptr[a + 1] && m < FRIEND_CAND_MAX; e++) {
        const uint32_t ent = g.u_ent[e];
        const unsigned mk = ENT_MASK(ent);
        if ((mk & src) && !(mk & MB_FRIEND)) cand[m++] = ENT_COL(ent);
    }
    if (m == 0) return;
    double u, u2;
    philox_u2(key, tick, ST_FRIEND_N, a, u, u2);
    int want = 0;
    while (want < 31 && u >= D.poisson_cdf[want]) want++;
    if (want > m) want = m;
    if (want == 0) return;
    // extra.weighted_n_of(n, candidates, social_proximity): left-to-right sum, p = w / sum, numpy's no-replacement sampler
    int nz = 0;
    double sump = 0.0;
    for (int i = 0; i < m; i++) w[i] = social_proximity(xa, attr[cand[i]], friends_shared_rows(g, a, cand[i]));
    for (int i = 0; i < m; i++) { sump = sump + w[i]; if (w[i] != 0.0) nz++; }
    if (nz < want) want = nz;
    if (sump == 0.0) return;
    for (int i = 0; i < m; i++) w[i] = w[i] / sump;
    int found[32], nf = 0, ucur = 0;
    while (nf < want) {
        for (int i = 0; i < nf; i++) w[found[i]] = 0.0;
        double acc = 0.0;
        for (int i = 0; i < m; i++) { acc = acc + w[i]; cdf[i] = acc; }
        for (int i = 0; i < m; i++) cdf[i] = cdf[i] / acc;
        const int need = want - nf, base = nf;
        for (int j = 0; j < need; j++) {
            double uu, uu2;
            philox_u2(key, tick, ST_FRIEND, a * 64 + (ucur & 63), uu, uu2);
            ucur++;
            int ix = ss_right(cdf, m, uu);
            if (ix >= m) ix = m - 1;
            bool dup = false;
            for (int q = base; q < nf; q++) dup |= (found[q] == ix);
            if (!dup) found[nf++] = ix;
        }
        if (ucur > 4096) break;
    }
    if (nf == 0) return;
    int *dmn = RP(C.dm_n, r, 4);
    const int pos = atomicAdd(&dmn[3], nf);
    unsigned long long *pairs = RP(C.dm_pairs, r, C.DPcap);
    if (pos + nf > C.DPcap) { atomicExch(&RP(C.counters, r, CN_COUNT)[CN_ERROR], 23); return; }
    for (int i = 0; i < nf; i++) pairs[pos + i] = DM_KEY(a, cand[found[i]], MB_FRIEND);
}

// pass 2: the friendship bit on both entries of every queued pair (Person.make_friendship_link, entities.py:133-140)
__global__ void __launch_bounds__(256) k_apply_friends(DevCtx C) {
    const int r = C.r0 + blockIdx.y;
    int *dmn = RP(C.dm_n, r, 4);
    const int np = min(dmn[3], C.DPcap);
    const unsigned long long *pairs = RP(C.dm_pairs, r, C.DPcap);
    Graph g = graph_of(C, r);
    uint32_t *ent = const_cast<uint32_t *>(g.u_ent);
    const uint32_t *attr = RP(C.attr, r, C.Ncap);
    long long *ltot = RP(C.layer_tot, r, 8);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < np; i += gridDim.x * blockDim.x) {
        const int a = DM_U(pairs[i]), c = DM_V(pairs[i]);
        const int e1 = static_find_idx(g, a, c), e2 = static_find_idx(g, c, a);
        unsigned m1 = 0u, m2 = 0u;   // the two entries as this phase leaves them (only friendship bits are set here)
        if (e1 >= 0) {
            const unsigned old = atomicOr(&ent[e1], MB_FRIEND << 24);
            m1 = old | (MB_FRIEND << 24);
            if (old & ENT_CRIM) crim_mask_or(g, a, c, MB_FRIEND);
            if (!(old & (MB_FRIEND << 24)) && (attr[a] & F_ALIVE)) atomicAdd((unsigned long long *)&ltot[5], 1ull);
        }
        if (e2 >= 0) {
            const unsigned old = atomicOr(&ent[e2], MB_FRIEND << 24);
            m2 = old | (MB_FRIEND << 24);
            if (old & ENT_CRIM) crim_mask_or(g, c, a, MB_FRIEND);
            if (!(old & (MB_FRIEND << 24)) && (attr[c] & F_ALIVE)) atomicAdd((unsigned long long *)&ltot[5], 1ull);
        }
        // ENT_ASYM (poc_device.cuh) when the two no longer mirror each other
        if (e1 < 0 || e2 < 0 || ((m1 ^ m2) & ENT_CRIM) || __popc(ENT_MASK(m1)) != __popc(ENT_MASK(m2))) {
            if (e1 >= 0) ent_flag_asym(g, a, e1);
            if (e2 >= 0) ent_flag_asym(g, c, e2);
        }
        if (e2 < 0) {
            // c's row has no slot for a (a one-sided link in the uploaded state: c cleared its own sets when it was
            // arrested before the upload, entities.py:601-602): the entry is created by the insert pass that follows
            const int q = atomicAdd(&dmn[2], 1);
            if (q < 1024) RP(C.dm_defer, r, 1024)[q] = DM_KEY(c, a, MB_FRIEND); else atomicExch(&RP(C.counters, r, CN_COUNT)[CN_ERROR], 24);
        }
    }
}
__global__ void k_friends_done(DevCtx C) {   // after pass 2: count, reset the queue, hand deferred entries to the insert pass
    const int r = C.r0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= C.r0 + C.Rg) return;
    int *dmn = RP(C.dm_n, r, 4);
    RP(C.counters, r, CN_COUNT)[CN_FRIENDSHIPS] += dmn[3];
    dmn[3] = 0;
    const int nd = min(dmn[2], 1024);
    dmn[2] = nd;
    unsigned long long *pairs = RP(C.dm_pairs, r, C.DPcap);
    for (int i = 0; i < nd; i++) pairs[i] = RP(C.dm_defer, r, 1024)[i];
}

// ---------------------------------------------------------------------------------------------- make_people_die
// pass 1 (thread per agent): who dies this tick
__global__ void __launch_bounds__(256) k_die_mark(DevCtx C) {
    const int tick = C.ti[C.grp].tick;
    const int r = C.r0 + blockIdx.y, a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= C.n_agents[r]) return;
    uint32_t *attr = RP(C.attr, r, C.Ncap);
    const uint32_t x = attr[a];
    if (!(x & F_ALIVE)) return;
    const DevDemog &D = *C.demog;
    double u, u2;
    philox_u2(C.philox_key[r], tick, ST_DIE, a, u, u2);
    const int age = attr_age(x);
    const double pm = (age <= D.mort_max_age && age < 128) ? D.mort[(x & F_MALE) ? 1 : 0][age] / (double)C.params[r].ticks_per_year : 1.0;
    if (u < pm || age > 119) {
        attr[a] = x | F_DYING;
        RP(C.dm_list, r, C.Ncap)[atomicAdd(&RP(C.dm_n, r, 4)[1], 1)] = a;
    }
}

// pass 2 (block per replica): facilitator replacement in ascending order of the dying agents, then Person.die
__global__ void __launch_bounds__(256) k_die_apply(DevCtx C) {
    const int tick = C.ti[C.grp].tick;
    const int r = C.r0 + blockIdx.x, tid = threadIdx.x, lane = lane_id(), w = warp_id(), nw = blockDim.x >> 5;
    int *dmn = RP(C.dm_n, r, 4);
    const int nd = dmn[1];
    if (nd == 0) return;
    const int n = C.n_agents[r];
    uint32_t *attr = RP(C.attr, r, C.Ncap);
    int32_t *list = RP(C.dm_list, r, C.Ncap);
    int32_t *ncd = RP(C.ncrim_dead, r, C.Ncap);
    long long *ltot = RP(C.layer_tot, r, 8);
    Graph g = graph_of(C, r);
    uint32_t *uent = const_cast<uint32_t *>(g.u_ent);
    int2 *cent = const_cast<int2 *>(g.c_ent);
    if (tid == 0) {
        // dying facilitators, ascending slot: each is replaced by a random adult who is neither facilitator nor OC member
        // (model.py:1590-1594): uniform over the slots by rejection
        for (int guard = 0; guard < nd; guard++) {
            int best = -1;
            for (int i = 0; i < nd; i++) { int a = list[i]; if ((attr[a] & F_FAC) && !(attr[a] & (1u << 30)) && (best < 0 || a < best)) best = a; }
            if (best < 0) break;
            attr[best] |= (1u << 30);   // handled (cleared below together with F_DYING)
            for (int j = 0; j < 64; j++) {
                double u, u2;
                philox_u2(C.philox_key[r], tick, ST_FACREP, best * 64 + j, u, u2);
                int c = (int)floor(u * (double)n);
                if (c >= n) c = n - 1;
                const uint32_t xc = attr[c];
                if ((xc & F_ALIVE) && !(xc & F_FAC) && !(xc & F_OC) && attr_age(xc) > 18) { attr[c] = xc | F_FAC; break; }
            }
        }
    }
    __syncthreads();
    // Person.die: the dying agent leaves the neighbour sets of its own out-neighbours, in all nine networks; its own
    // sets stay.  Rows of agents that die in the same tick are left alone (batch rule).  One warp per dying agent.
    for (int i = w; i < nd; i += nw) {
        const int a = list[i];
        const int b0 = g.u_rowptr[a], e0 = g.u_rowptr[a + 1], b1 = g.c_rowptr[a], e1 = g.c_rowptr[a + 1];
        const int len0 = e0 - b0, len = len0 + (e1 - b1);
        int own[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int j = lane; j < len; j += 32) {
            int v = -1;
            if (j < len0) {
                const uint32_t ent = uent[b0 + j];
                const unsigned mk = ENT_MASK(ent);
#pragma unroll
                for (int l = 0; l < 8; l++) own[l] += (mk >> l) & 1u;
                if (mk) v = ENT_COL(ent);
            } else {
                const int2 ce = cent[b1 + (j - len0)];
                if (!CE_IS_DEAD(ce.y)) v = ce.x;
            }
            if (v < 0 || v == a || (attr[v] & F_DYING)) continue;
            if (j < len0) ent_flag_asym(g, a, b0 + j);   // the reverse entry is emptied below: this one stands alone
            else atomicOr(&cent[b1 + (j - len0)].y, CE_ASYM);
            const bool valive = (attr[v] & F_ALIVE) != 0;
            const int idx = static_find_idx(g, v, a);
            if (idx >= 0) {
                const unsigned old = atomicAnd(&uent[idx], 0x00ffffffu);
                if (valive) for (int l = 0; l < 8; l++) if ((old >> (24 + l)) & 1u) atomicAdd((unsigned long long *)&ltot[l], (unsigned long long)-1ll);
            }
            const int cidx = crim_find_idx(g, v, a);
            if (cidx >= 0) {
                const int oldy = atomicExch(&cent[cidx].y, CE_DEAD);
                if (!CE_IS_DEAD(oldy)) atomicAdd(&ncd[v], 1);
            }
        }
        // the dying agent's own rows no longer count towards the tot_<layer>_link reporters (scheduled agents only)
#pragma unroll
        for (int l = 0; l < 8; l++) {
            int s = own[l];
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL_MASK, s, o);
            if (lane == 0 && s) atomicAdd((unsigned long long *)&ltot[l], (unsigned long long)(-(long long)s));
        }
    }
    __syncthreads();
    for (int i = tid; i < nd; i += blockDim.x) { const int a = list[i]; attr[a] &= ~(F_ALIVE | F_DYING | (1u << 30)); }
    if (tid == 0) { RP(C.counters, r, CN_COUNT)[CN_DECEASED] += nd; dmn[1] = 0; }
}

// ---------------------------------------------------------------------------------------------- the next rows of 8(f)
// a layer bit leaves the criminal entry's copy of the static mask too; the pair stands for itself from then on
__device__ __forceinline__ void crim_mask_clear(const Graph &g, int u, int v, unsigned bits) {
    const int ci = crim_find_idx(g, u, v);
    if (ci >= 0 && !CE_IS_DEAD(g.c_ent[ci].y)) {
        int2 *cent = const_cast<int2 *>(g.c_ent);
        atomicAnd(&cent[ci].y, ~(int)(bits << 24));
        atomicOr(&cent[ci].y, CE_ASYM);
        const int cj = crim_find_idx(g, v, u);
        if (cj >= 0) atomicOr(&cent[cj].y, CE_ASYM);
    }
}

// retire_persons, model.py:1538-1551 (oracle/hotpath.c orc_retire_persons): thread per agent, no draws
__global__ void __launch_bounds__(256) k_retire(DevCtx C) {
    const int r = C.r0 + blockIdx.y, a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= C.n_agents[r]) return;
    uint32_t *attr = RP(C.attr, r, C.Ncap);
    const uint32_t x = attr[a];
    if (!(x & F_ALIVE) || !(x & F_HASJOB) || attr_age(x) < C.demog->retirement_age) return;
    attr[a] = x & ~F_HASJOB;
    Graph g = graph_of(C, r);
    uint32_t *uent = const_cast<uint32_t *>(g.u_ent);
    int nprof = 0;
    for (int e = g.u_rowptr[a]; e < g.u_rowptr[a + 1]; e++) {   // the OWN professional set only
        const uint32_t en = uent[e];
        if (!(ENT_MASK(en) & MB_PROF)) continue;
        nprof++;
        atomicAnd(&uent[e], ~(MB_PROF << 24));
        if (en & ENT_CRIM) crim_mask_clear(g, a, ENT_COL(en), MB_PROF);
        ent_flag_asym(g, a, e);   // one-sided: neither entry of the pair mirrors the other now
        const int ri = static_find_idx(g, ENT_COL(en), a);
        if (ri >= 0) ent_flag_asym(g, ENT_COL(en), ri);
    }
    if (nprof) atomicAdd((unsigned long long *)&RP(C.layer_tot, r, 8)[6], (unsigned long long)(-(long long)nprof));
}

// ---------------------------------------------------------------------------------------------- the yearly labour-status block
// graduate_and_enter_jobmarket (model.py:1351-1380; enroll_to_school / leave_school, entities.py:292-308, 389-395) and the
// update_unemployment_status loop written out in step() (model.py:252-256, entities.py:397-411) = oracle/hotpath.c
// orc_labour_yearly: thread per agent, the agent's own attributes only (no links are made or cleared, and which school of a level
// enroll_to_school picks matters to nothing on this path).  The launch sequence of a tick carries no arguments, so the kernel is
// launched every tick and leaves at once unless the tick is one of the reference's yearly ones (force: the stage hook).
__device__ __forceinline__ int pick_pair(const int *val, const double *cdf, int n, double u) {   // extra.pick_from_pair_list
    const int i = ss_right(cdf, n, u);
    return val[i < n ? i : n - 1];
}
__global__ void __launch_bounds__(256) k_labour(DevCtx C, int force) {
    const int r = C.r0 + blockIdx.y;
    const int tick = C.ti[C.grp].tick;
    const poc_params &P = C.params[r];
    if (!force && !((tick % P.ticks_per_year) == 0 || tick == 1)) return;   // model.py:248
    const int n = C.n_agents[r];
    uint32_t *attr = RP(C.attr, r, C.Ncap);
    uint8_t *e2 = RP(C.edu2, r, C.Ncap);
    const int32_t *birth = RP(C.birth, r, C.Ncap);
    const DevLabour &L = *C.labour;
    const unsigned long long key = C.philox_key[r];
    const int retirement_age = C.demog->retirement_age;
    // a few blocks per replica stride over the agents: eleven launches of twelve find nothing to do and should cost nothing
    for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < n; a += gridDim.x * blockDim.x) {
        const uint32_t x = attr[a];
        if (!(x & F_ALIVE)) continue;
        const unsigned ed = e2[a];
        const int age = attr_age(x), male = (x & F_MALE) ? 1 : 0, maxedu = (int)(ed & 15u);
        int edu = attr_edu(x), job = attr_job(x), wl = attr_wealth(x), slv = (int)(ed >> 4);
        bool school = (x & F_HASSCHOOL) != 0;
        // model.py:1356-1360: children of primary age who never went to school enrol at level 1
        if (edu == 0 && age == L.primary_age && !school) { school = true; slv = 1; }
        // :1361-1380: who is one year past the last age of the own school graduates
        const int lv = school ? slv : 0;
        if (lv >= 1 && lv <= L.n_levels && age == L.end_age[lv] + 1) {
            school = false; slv = 0;   // leave_school
            edu = lv;
            if (lv + 1 <= L.n_levels && lv + 1 <= maxedu) { school = true; slv = lv + 1; }
            else {
                double u0, u1;
                philox_u2(key, tick, ST_GRAD, a, u0, u1);
                const double u2 = philox_u1s(key, tick, ST_GRAD, a, 1);
                if (L.work_n[lv][male] > 0) job = pick_pair(L.work_val[lv][male], L.work_cdf[lv][male], L.work_n[lv][male], u0);
                if (job >= 0 && job < 8 && L.wealth_n[job][male] > 0)
                    wl = pick_pair(L.wealth_val[job][male], L.wealth_cdf[job][male], L.wealth_n[job][male], u1);
                if (age > 14 && age < retirement_age && job == 1 && age < 128 && u2 < L.inactive[male][age]) job = 0;
            }
        }
        // model.py:252-256: the not employed draw their status again when they enter an age range of labour_status_range
        if (job < 2 && ((tick - birth[a]) % P.ticks_per_year) == 0 && age < 128 && L.range_age[male][age]) {
            double u, u_;
            philox_u2(key, tick, ST_UNEMP, a, u, u_);
            job = (u < L.inactive[male][age]) ? 0 : 1;   // entities.py:410-411
        }
        const uint32_t y = (x & ~(F_HASSCHOOL | 0x0fff0000u)) | (school ? F_HASSCHOOL : 0u) | ((uint32_t)(job & 15) << 16) |
                           ((uint32_t)(edu & 15) << 20) | ((uint32_t)(wl & 15) << 24);
        if (y != x) attr[a] = y;
        const unsigned ed2 = (unsigned)maxedu | ((unsigned)slv << 4);
        if (ed2 != ed) e2[a] = (uint8_t)ed2;
    }
}

// remove_excess_friends / remove_excess_professional_links, pass 1 (thread per agent, decisions on the state the phase
// started from): selection sampling over the row in ascending slot order, the chosen links are queued
__global__ void __launch_bounds__(128) k_remove_excess(DevCtx C, unsigned bit, int stage, int dunbar) {
    const int tick = C.ti[C.grp].tick;
    const int r = C.r0 + blockIdx.y, a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= C.n_agents[r]) return;
    const uint32_t xa = RP(C.attr, r, C.Ncap)[a];
    if (!(xa & F_ALIVE)) return;
    Graph g = graph_of(C, r);
    const int b = g.u_rowptr[a], e = g.u_rowptr[a + 1];
    const int age = attr_age(xa), limit = dunbar ? 150 - abs(age - 30) : 30;
    if (e - b <= limit) return;   // not even the whole row is over the limit: nearly every agent leaves here
    int len = 0;
    for (int i = b; i < e; i++) len += (ENT_MASK(g.u_ent[i]) & bit) != 0;
    if (len <= limit) return;
    int need = len - limit, k = 0;
    const unsigned long long key = C.philox_key[r];
    int *dmn = RP(C.dm_n, r, 4);
    unsigned long long *pairs = RP(C.dm_pairs, r, C.DPcap);
    for (int i = b; i < e && need > 0; i++) {
        const uint32_t en = g.u_ent[i];
        if (!(ENT_MASK(en) & bit)) continue;
        const double u = philox_u1s(key, tick, stage, a, k);
        if (u * (double)(len - k) < (double)need) {
            const int pos = atomicAdd(&dmn[3], 1);
            if (pos < C.DPcap) pairs[pos] = DM_KEY(a, ENT_COL(en), bit); else atomicExch(&RP(C.counters, r, CN_COUNT)[CN_ERROR], 25);
            need--;
        }
        k++;
    }
}

// pass 2: the bit leaves both entries of every queued link (Person.remove_link, entities.py:226-237)
__global__ void __launch_bounds__(256) k_apply_removals(DevCtx C, unsigned bit, int layer) {
    const int r = C.r0 + blockIdx.y;
    int *dmn = RP(C.dm_n, r, 4);
    const int np = min(dmn[3], C.DPcap);
    const unsigned long long *pairs = RP(C.dm_pairs, r, C.DPcap);
    Graph g = graph_of(C, r);
    uint32_t *ent = const_cast<uint32_t *>(g.u_ent);
    const uint32_t *attr = RP(C.attr, r, C.Ncap);
    long long *ltot = RP(C.layer_tot, r, 8);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < np; i += gridDim.x * blockDim.x) {
        const int a = DM_U(pairs[i]), c = DM_V(pairs[i]);
        const int e1 = static_find_idx(g, a, c), e2 = static_find_idx(g, c, a);
        if (e1 >= 0) {
            const unsigned old = atomicAnd(&ent[e1], ~(bit << 24));
            if (old & (bit << 24)) {
                if (attr[a] & F_ALIVE) atomicAdd((unsigned long long *)&ltot[layer], (unsigned long long)-1ll);
                if (old & ENT_CRIM) crim_mask_clear(g, a, c, bit);
            }
        }
        if (e2 >= 0) {
            const unsigned old = atomicAnd(&ent[e2], ~(bit << 24));
            if (old & (bit << 24)) {
                if (attr[c] & F_ALIVE) atomicAdd((unsigned long long *)&ltot[layer], (unsigned long long)-1ll);
                if (old & ENT_CRIM) crim_mask_clear(g, c, a, bit);
            }
        }
        // a link that was one-sided before may leave the pair unbalanced (or balance it: the flag only ever gets set)
        ent_asym_update(g, a, c);
    }
}
__global__ void k_removals_done(DevCtx C) {
    const int r = C.r0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (r < C.r0 + C.Rg) RP(C.dm_n, r, 4)[3] = 0;
}

// ---------------------------------------------------------------------------------------------- wedding, model.py:368-402
// oracle/hotpath.c orc_wedding states the semantics (sequential by nature: the marriable list shrinks with every ego); one
// block per replica walks it with three bitmaps in shared memory (marriable, visited, pool).  Links that go are cleared in
// place; the new household / partner links are queued for k_multiplex_insert (a couple need not have been neighbours).
__device__ __forceinline__ double det_exp_neg(double x) {   // exp(-x) by fixed arithmetic: the same bits as the oracle's
    const double y = x * (1.0 / 1024.0);
    double t = 1.0;
    for (int k = 14; k >= 1; k--) t = 1.0 - y * t / (double)k;
    for (int i = 0; i < 10; i++) t = t * t;
    return t;
}
// the j-th set bit (ascending) of a bitmap; one warp, every lane returns the slot (-1: fewer bits)
__device__ int warp_select_bit(const unsigned *bm, int nwords, int j) {
    const int lane = lane_id();
    int base = 0;
    for (int w0 = 0; w0 < nwords; w0 += 32) {
        const int i = w0 + lane;
        const unsigned wd = (i < nwords) ? bm[i] : 0u;
        const int c = __popc(wd);
        int x = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(FULL_MASK, x, o); if (lane >= o) x += y; }
        const int tot = __shfl_sync(FULL_MASK, x, 31);
        if (j < base + tot) {
            const bool mine = j >= base + x - c && j < base + x;
            const unsigned m = __ballot_sync(FULL_MASK, mine);
            int res = -1;
            if (mine) {
                unsigned t = wd;
                for (int q = j - (base + x - c); q > 0; q--) t &= t - 1;
                res = i * 32 + (__ffs(t) - 1);
            }
            return __shfl_sync(FULL_MASK, res, __ffs(m) - 1);
        }
        base += tot;
    }
    return -1;
}

__global__ void __launch_bounds__(256) k_wedding(DevCtx C) {
    extern __shared__ unsigned wsm[];
    const int tick = C.ti[C.grp].tick;
    const int r = C.r0 + blockIdx.x, tid = threadIdx.x, lane = lane_id(), w = warp_id(), nw = blockDim.x >> 5;
    const int n = C.n_agents[r], nwords = (n + 31) >> 5, wcap = (C.Ncap + 31) >> 5;
    unsigned *marr = wsm, *seen = wsm + wcap, *pool = wsm + 2 * wcap;
    __shared__ int s_M, s_num, s_alive, s_ego, s_partner, s_np, s_cnt, s_done, s_npairs;
    const poc_params &P = C.params[r];
    const DevDemog &D = *C.demog;
    const uint32_t *attr = RP(C.attr, r, C.Ncap);
    Graph g = graph_of(C, r);
    uint32_t *uent = const_cast<uint32_t *>(g.u_ent);
    int32_t *list = RP(C.dm_list, r, C.Ncap);
    long long *ltot = RP(C.layer_tot, r, 8);
    const unsigned long long key = C.philox_key[r];
    int *dmn = RP(C.dm_n, r, 4);
    unsigned long long *pairs = RP(C.dm_pairs, r, C.DPcap);
    if (tid == 0) { s_M = 0; s_alive = 0; s_done = 0; s_npairs = 0; }
    __syncthreads();
    // ---- the marriable: alive, 25 < age < 55, no partner (a bit per agent, ascending slot order = the list order)
    for (int base = 0; base < nwords * 32; base += blockDim.x) {
        const int a = base + tid;
        bool al = false, mr = false;
        if (a < n) {
            const uint32_t x = attr[a];
            al = (x & F_ALIVE) != 0;
            const int age = attr_age(x);
            if (al && age > 25 && age < 55) {
                mr = true;
                for (int e = g.u_rowptr[a]; e < g.u_rowptr[a + 1]; e++) if (ENT_MASK(uent[e]) & MB_PARTNER) { mr = false; break; }
            }
        }
        const unsigned bm = __ballot_sync(FULL_MASK, mr), ba = __ballot_sync(FULL_MASK, al);
        if (lane == 0) {
            if ((a >> 5) < nwords) marr[a >> 5] = bm;
            if (bm) atomicAdd(&s_M, __popc(bm));
            if (ba) atomicAdd(&s_alive, __popc(ba));
        }
    }
    __syncthreads();
    if (tid == 0) {   // numpy's poisson by inversion of the CDF with one uniform (oracle: the same chain)
        const double lam = D.weddings_mean * (double)s_alive / 1000.0 / 12.0;
        double u, u2;
        philox_u2(key, tick, ST_WED_N, 0, u, u2);
        int num = 0;
        double pk = det_exp_neg(lam), cdf = pk;
        while (u >= cdf && num < 1000) { num++; pk = pk * lam / (double)num; cdf = cdf + pk; }
        s_num = num;
    }
    __syncthreads();
    const int radius = P.max_accomplice_radius;
    for (int it = 0; s_num > 0 && s_M > 1; it++) {
        if (w == 0) {
            double u, u2;
            philox_u2(key, tick, ST_WED_EGO, it, u, u2);
            int j = (int)floor(u * (double)s_M);
            if (j >= s_M) j = s_M - 1;
            const int ego = warp_select_bit(marr, nwords, j);
            if (lane == 0) s_ego = ego;
        }
        for (int i = tid; i < nwords; i += blockDim.x) pool[i] = 0u;
        __syncthreads();
        const int ego = s_ego;
        const uint32_t xe = attr[ego];
        // ---- neighbors_range(friendship, radius) | neighbors_range(professional, radius): each network walked on its own
        for (int q = 0; q < 2; q++) {
            const unsigned lbit = q ? MB_PROF : MB_FRIEND;
            for (int i = tid; i < nwords; i += blockDim.x) seen[i] = 0u;
            __syncthreads();
            if (tid == 0) { list[0] = ego; seen[ego >> 5] = 1u << (ego & 31); s_cnt = 1; }
            __syncthreads();
            int fb = 0, fe = 1;
            for (int d = 0; d < radius && fb < fe; d++) {
                for (int i = fb + w; i < fe; i += nw) {
                    const int u = list[i];
                    for (int e = g.u_rowptr[u] + lane; e < g.u_rowptr[u + 1]; e += 32) {
                        const uint32_t en = uent[e];
                        if (!(ENT_MASK(en) & lbit)) continue;
                        const int v = ENT_COL(en);
                        const unsigned bit = 1u << (v & 31);
                        if (!(atomicOr(&seen[v >> 5], bit) & bit)) { list[atomicAdd(&s_cnt, 1)] = v; atomicOr(&pool[v >> 5], bit); }
                    }
                }
                __syncthreads();
                fb = fe; fe = s_cnt;
                __syncthreads();
            }
        }
        // ---- the pool: marriable, other sex, (age - ego.age) < 8, not sibling / offspring of the ego, ego not its offspring
        if (tid == 0) s_np = 0;
        __syncthreads();
        int cnt = 0;
        for (int wi = tid; wi < nwords; wi += blockDim.x) {
            unsigned pw = pool[wi] & marr[wi], keep = 0u;
            while (pw) {
                const int b = __ffs(pw) - 1;
                pw &= pw - 1;
                const int v = wi * 32 + b;
                const uint32_t xv = attr[v];
                bool ok = v != ego && ((xv ^ xe) & F_MALE) && (attr_age(xv) - attr_age(xe)) < 8;
                if (ok && (static_find_mask(g, ego, v) & (MB_SIB | MB_OFF))) ok = false;
                if (ok && (static_find_mask(g, v, ego) & MB_OFF)) ok = false;
                if (ok) keep |= 1u << b;
            }
            pool[wi] = keep;
            cnt += __popc(keep);
        }
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(FULL_MASK, cnt, o);
        if (lane == 0 && cnt) atomicAdd(&s_np, cnt);
        __syncthreads();
        if (w == 0) {
            int partner = -1;
            if (s_np > 0) {   // a uniform pick: extra.wedding_proximity_with always returns equal weights (see the oracle)
                double u, u2;
                philox_u2(key, tick, ST_WED_PICK, it, u, u2);
                int idx = (int)floor(u * (double)s_np);
                if (idx >= s_np) idx = s_np - 1;
                partner = warp_select_bit(pool, nwords, idx);
            }
            if (lane == 0) s_partner = partner;
        }
        __syncthreads();
        const int partner = s_partner;
        if (partner >= 0) {
            // ---- both leave their households: reciprocal links only (remove_from_household, entities.py:378-387); warp 0 / 1
            if (w < 2) {
                const int x = w ? partner : ego;
                for (int e = g.u_rowptr[x] + lane; e < g.u_rowptr[x + 1]; e += 32) {
                    const uint32_t en = uent[e];
                    if (!(ENT_MASK(en) & MB_HH)) continue;
                    const int m = ENT_COL(en);
                    const int ri = static_find_idx(g, m, x);
                    if (ri < 0) continue;
                    // (the two warps may meet on the couple's own link: whoever clears an entry accounts for it)
                    if (!(uent[ri] & (MB_HH << 24)) && !(m == (w ? ego : partner))) continue;
                    const unsigned o1 = atomicAnd(&uent[e], ~(MB_HH << 24)), o2 = atomicAnd(&uent[ri], ~(MB_HH << 24));
                    if ((o1 & (MB_HH << 24)) && (attr[x] & F_ALIVE)) atomicAdd((unsigned long long *)&ltot[4], (unsigned long long)-1ll);
                    if ((o2 & (MB_HH << 24)) && (attr[m] & F_ALIVE)) atomicAdd((unsigned long long *)&ltot[4], (unsigned long long)-1ll);
                    if ((o1 & (MB_HH << 24)) && (o1 & ENT_CRIM)) crim_mask_clear(g, x, m, MB_HH);
                    if ((o2 & (MB_HH << 24)) && (o2 & ENT_CRIM)) crim_mask_clear(g, m, x, MB_HH);
                    ent_asym_update(g, x, m);
                }
            }
            if (tid == 0) {
                const int q = s_npairs;
                if (q + 2 <= C.DPcap) {
                    pairs[q] = DM_KEY(ego, partner, MB_HH | MB_PARTNER); pairs[q + 1] = DM_KEY(partner, ego, MB_HH | MB_PARTNER);
                    s_npairs = q + 2;
                } else atomicExch(&RP(C.counters, r, CN_COUNT)[CN_ERROR], 26);
                marr[partner >> 5] &= ~(1u << (partner & 31));
                s_M -= 1; s_num -= 1; s_done += 1;
            }
        }
        __syncthreads();
        if (tid == 0) { marr[ego >> 5] &= ~(1u << (ego & 31)); s_M -= 1; }
        __syncthreads();
    }
    if (tid == 0) {
        dmn[0] = 0; dmn[2] = s_npairs;   // the insert pass that follows adds the couples' household + partner links
        RP(C.counters, r, CN_COUNT)[CN_WEDDINGS] += s_done;
    }
}
