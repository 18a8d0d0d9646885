// poc_api.cu -- C ABI (include/protonoc_b200.h) over the kernels in poc_kernels.cuh.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 --fmad=false -shared -Xcompiler -fPIC
//        (--fmad=false: candidate keys, tendencies and CDFs must equal CPython's left-to-right float64
//         arithmetic bit for bit, so no multiply-add contraction anywhere.)
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "poc_kernels.cuh"

#define N_TIMERS 8
enum { TM_CELLS = 0, TM_TENDENCY, TM_DRAW, TM_SEARCH, TM_EMB, TM_COMMIT, TM_ARREST, TM_OTHER };
#define EV_POOL 32768

#define POC_STAGE_COLS(X) \
    X(propensity, double) X(tendency, double) X(sentence, double) X(arrest_weight, double) \
    X(age, int32_t) X(birth_tick, int32_t) X(num_crimes, int32_t) X(num_crimes_tick, int32_t) X(father, int32_t) X(new_recruit, int32_t) \
    X(alive, uint8_t) X(male, uint8_t) X(oc_member, uint8_t) X(facilitator, uint8_t) X(prisoner, uint8_t) \
    X(has_job, uint8_t) X(has_school, uint8_t) X(target, uint8_t) \
    X(job_level, int8_t) X(education_level, int8_t) X(wealth_level, int8_t)

struct poc_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    DevCtx d{};
    std::vector<void *> allocs;
    std::vector<poc_params> h_params;
    std::vector<int> h_nagents;
    std::string err;
    int sticky = 0;
    int epoch = 0;
    int gx = 4;            // blocks per replica of the search grid (and of the block-per-evaluation embeddedness grid)
    int emb_g = 256;       // threads per embeddedness evaluation: 256 (block) for small batches, 32 (warp) for large ones
    int emb_gx = 4;        // blocks per replica of the embeddedness grid
    size_t smem_search = 0, smem_emb = 0, smem_emb2 = 0, smem_emb_big = 0;
    int emb_blocks = 8;    // resident blocks per SM the second formulation is compiled / sized for
    int emb_v = 3;         // embeddedness formulation: 3 = poc_embed3.cuh (each meta edge once), 2 = poc_embed2.cuh, 1 = poc_embed.cuh
    // staging
    StageCols stage{};
    uint8_t *stage_dev = nullptr, *stage_host = nullptr;   // one block for all staging columns + its pinned mirror
    uint8_t *stage_host2[2] = {nullptr, nullptr};          // upload mirrors, used alternately
    cudaEvent_t stage_ev2[2] = {nullptr, nullptr}; bool stage_busy2[2] = {false, false}; int stage_next = 0;
    uint8_t *many_dev[4] = {}, *many_host[4] = {};         // poc_download_agents_many: four staging blocks in flight
    cudaEvent_t many_ev[4] = {};
    size_t stage_bytes = 0, stage_off[24] = {};
    int cells_threads = 1024;       // k_cells block size (its count table is n_cells x threads ints of shared memory)
    int search_threads = 256;       // k_search block size (one block per crime)
    int commit_threads = 1024;      // k_commit block size, same reasoning
    int draw_threads = 1024;        // k_draw block size: one block per SM is latency-bound, several need fewer registers per block
    int32_t *st_rowptr[POC_NUM_LAYERS] = {};
    int32_t *st_col = nullptr;      // all staged layers back to back
    int32_t *st_co = nullptr;
    long long st_off[POC_NUM_LAYERS + 1] = {};
    int st_idx[POC_NUM_LAYERS] = {};  // layer -> position in the staging area
    int st_have = 0;                // bitmask of layers staged for st_replica
    int st_replica = -1;
    int *d_err = nullptr;
    int *d_prev_counters = nullptr;
    // packed uploads (poc_upload_packed): two device staging blobs used alternately
    uint8_t *blob_dev[2] = {nullptr, nullptr}; size_t blob_cap = 0;
    cudaEvent_t blob_ev[2] = {nullptr, nullptr}; int blob_next = 0;
    // instantiated tick graphs of poc_run_ticks, kept across calls
    struct GraphKey { int ticks, report, n, groups, radius, Tcap, phases; const void *rows; bool operator==(const GraphKey &o) const { return ticks == o.ticks && report == o.report && n == o.n && groups == o.groups && radius == o.radius && Tcap == o.Tcap && phases == o.phases && rows == o.rows; } };
    std::vector<std::pair<GraphKey, cudaGraphExec_t>> graphs;
    std::vector<long long> graph_launches;
    int32_t *d_hook_a = nullptr, *d_hook_b = nullptr; double *d_hook_out = nullptr; long long hook_cap = 0;
    // timing
    bool timing = false;
    std::vector<cudaEvent_t> ev;
    std::vector<int> ev_class;
    int ev_used = 0;
    double tm_ms[N_TIMERS] = {};
    long long tm_n[N_TIMERS] = {};
    long long launches = 0;
    cudaEvent_t marker[16] = {};
    // replica groups of poc_run_ticks: each group's tick pipeline runs on its own stream so that the
    // latency-bound one-block-per-replica kernels of one group overlap the traversal kernels of another
    int ngroups = 1;
    cudaStream_t gstream[8] = {};
    cudaEvent_t gev[9] = {};
    cudaEvent_t gstag[8] = {};
    int graph_ticks = 8;
    TickInfo *d_ti = nullptr;
    bool use_graph = true;
    int phases = 0;                 // demography phases poc_run_ticks runs per tick (poc_set_phases)
    DevDemog *d_demog = nullptr;
    bool demog_set = false;
    DevLabour *d_labour = nullptr;
    bool labour_set = false;
};

static thread_local std::string g_err;

static int fail(poc_ctx *c, int code, const std::string &msg) {
    if (c) { c->err = msg; if (code == POC_ERR_CUDA) c->sticky = code; }
    g_err = msg;
    return code;
}
#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t _e = (call);                                                                         \
        if (_e != cudaSuccess) return fail(ctx, POC_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e)); \
    } while (0)
#define CHECK_CTX()                                                   \
    do {                                                              \
        if (!ctx) return POC_ERR_ARG;                                 \
        if (ctx->sticky) return ctx->sticky;                          \
        cudaSetDevice(ctx->device);                                   \
    } while (0)

template <typename T>
static int dalloc(poc_ctx *ctx, T **p, size_t count) {
    void *q = nullptr;
    size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    CU(cudaMalloc(&q, bytes));
    CU(cudaMemsetAsync(q, 0, bytes, ctx->stream));
    ctx->allocs.push_back(q);
    *p = (T *)q;
    return 0;
}
#define DA(ptr, count)                                 \
    do {                                               \
        int _r = dalloc(ctx, &(ptr), (size_t)(count)); \
        if (_r) return _r;                             \
    } while (0)

struct Timed {
    poc_ctx *c; int cls; int idx; cudaStream_t s;
    Timed(poc_ctx *c_, int cls_, cudaStream_t s_ = nullptr) : c(c_), cls(cls_), idx(-1), s(s_ ? s_ : c_->stream) {
        c->launches++;
        if (c->timing && c->ev_used + 2 <= (int)c->ev.size()) {
            idx = c->ev_used; c->ev_used += 2;
            c->ev_class[idx] = cls;
            cudaEventRecord(c->ev[idx], s);
        }
    }
    ~Timed() { if (idx >= 0) cudaEventRecord(c->ev[idx + 1], s); }
};

extern "C" {

int poc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

const char *poc_last_error(poc_ctx *ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }
int poc_sizeof_params(void) { return (int)sizeof(poc_params); }
int poc_sizeof_tick_out(void) { return (int)sizeof(poc_tick_out); }

static int ctx_init(poc_ctx *ctx, int n_replicas, int max_agents, int64_t max_static_entries, int64_t max_criminal_entries);

int poc_ctx_create(int device, int n_replicas, int max_agents, int64_t max_static_entries,
                   int64_t max_criminal_entries, poc_ctx **out) {
    if (!out || n_replicas < 1 || max_agents < 1 || max_agents >= (1 << 22) || max_static_entries < 0 || max_criminal_entries < 0)
        return fail(nullptr, POC_ERR_ARG, "poc_ctx_create: bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(nullptr, POC_ERR_NOGPU, "poc_ctx_create: no CUDA device (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(nullptr, POC_ERR_ARG, "poc_ctx_create: bad device index");
    poc_ctx *ctx = new poc_ctx();
    ctx->device = device;
    cudaSetDevice(device);
    int rc = ctx_init(ctx, n_replicas, max_agents, max_static_entries, max_criminal_entries);
    if (rc) {   // release whatever was created; the message survives in the library-wide slot
        if (!ctx->err.empty()) g_err = ctx->err;
        poc_ctx_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return POC_OK;
}

static int ctx_init(poc_ctx *ctx, int n_replicas, int max_agents, int64_t max_static_entries, int64_t max_criminal_entries) {
    cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) return fail(nullptr, POC_ERR_CUDA, cudaGetErrorString(e));
    DevCtx &d = ctx->d;
    int R = n_replicas, N = max_agents;
    d.R = R; d.Ncap = N; d.r0 = 0; d.Rg = R;
    d.Ccap = N / 16 + 64;
    d.Mcap = d.Ccap * 4 + 256;
    int pc = 16384;
    while (pc < d.Ccap * 8) pc <<= 1;
    d.Pcap = pc;
    d.ucap = (std::max<long long>(max_static_entries, 1) + 3) & ~3LL;   // rows are read 16 bytes at a time: replica bases stay aligned
    d.ccap = std::max<long long>(max_criminal_entries, 1);
    d.Tcap = 1;
    // one block per accomplice search / embeddedness evaluation: at 10k agents a replica issues about 25 searches
    // and 20-50 evaluations per tick, so 32-64 blocks per replica let every block take at most one or two items
    // (blocks without work exit at once); 6 blocks of 256 threads are resident per SM
    ctx->gx = (R >= 256) ? 16 : (R >= 64) ? 32 : 64;   // search at 1,024 replicas: 16 / 32 / 64 blocks per replica -> 168 / 174 / 197 ms per 480 ticks
    // a small batch of large populations has hundreds of evaluations per replica and tick: enough blocks for every SM
    if (N > 30000 && R * ctx->gx < 2 * 148) ctx->gx = (2 * 148 + R - 1) / R;
    // embeddedness: a chain of dependent gathers per evaluation, so large batches run one evaluation per warp
    // (about 24 resident per SM) and small ones one per block
    // threads per evaluation (one block each), measured us per tick at 1 / 128 / 512 replicas (profiles/README.md):
    // 1024 -> 198 / - / -, 512 -> - / 688 / 2041, 256 -> - / 699 / 1914, 128 -> - / 806 / 1969, 64 -> - / 1055 / 2389
    ctx->emb_g = (R < 8) ? 1024 : (R < 256 ? 512 : 256);
    if (const char *e = getenv("POC_EMB_G")) { int g = atoi(e); if (g == 32 || g == 64 || g == 128 || g == 256 || g == 512 || g == 1024) ctx->emb_g = g; }  // experiments
    if (const char *e = getenv("POC_GX")) ctx->gx = std::max(1, atoi(e));
    ctx->draw_threads = (R > 148) ? 512 : 1024;   // measured at 512 replicas: 256 / 512 / 1024 threads -> 177 / 143 / 213 us
    ctx->commit_threads = (R > 512) ? 256 : (R > 148) ? 512 : 1024;   // 1,024 replicas: 256 / 512 / 1024 threads -> 137 / 147 / 169 ms per 480 ticks
    ctx->search_threads = (R < 8) ? 512 : (R < 256 ? 256 : 128);   // measured us per tick at 1 / 128 / 512 replicas: 128 threads - / 650 / 1723, 256: 175 / 641 / 1753, 512: 173 / 672 / 1916
    if (const char *e = getenv("POC_SEARCH_T")) { int t = atoi(e); if (t == 128 || t == 256 || t == 512) ctx->search_threads = t; }
    ctx->cells_threads = 1024;   // measured: 512 threads is no faster at 512 replicas (59.6 vs 52 us)
    if (const char *e = getenv("POC_CELLS_T")) { int t = atoi(e); if (t == 256 || t == 512 || t == 1024) ctx->cells_threads = t; }
    if (const char *e = getenv("POC_COMMIT_T")) { int t = atoi(e); if (t == 256 || t == 512 || t == 1024) ctx->commit_threads = t; }
    if (const char *e = getenv("POC_DRAW_T")) { int t = atoi(e); if (t == 256 || t == 384 || t == 512 || t == 1024) ctx->draw_threads = t; }
    // embeddedness at 1,024 replicas: 16 / 32 / 64 / 96 / 128 blocks per replica -> 737 / 648 / 620 / 623 / 635 ms per 480 ticks
    // (late in a run a replica asks for 40-100 evaluations per launch; a block that gets a second one is the tail)
    ctx->emb_gx = (ctx->emb_g == 32) ? std::max(1, std::min(16, (148 * 24 + R * 8 - 1) / (R * 8))) : std::max(ctx->gx, 64);
    if (const char *e = getenv("POC_EMB_GX")) ctx->emb_gx = std::max(1, atoi(e));
    int emb_slots = R * ctx->emb_gx * (ctx->emb_g == 32 ? 8 : 1);
    d.scr_slots = std::max(R * ctx->gx, emb_slots);
    d.slot_stride = std::max(ctx->gx, ctx->emb_gx * (ctx->emb_g == 32 ? 8 : 1));
    size_t wcap = (size_t)(N + 31) / 32;
    {   // bitmaps (+ the staged ring when that A/B switch is on: poc_kernels.cuh search_crime)
        const char *e = getenv("POC_EMB_FLAGS");
        const bool stage_ring = e && (atoi(e) & 2);
        ctx->smem_search = ((2 * wcap + 1) & ~(size_t)1) * sizeof(unsigned) + (stage_ring ? SEARCH_SCAND * (sizeof(double) + sizeof(int)) : 0);
    }
    ctx->smem_emb = emb_smem_bytes(ctx->emb_g, N);
    if (ctx->smem_emb + 2 * 1024 > 220 * 1024) return fail(nullptr, POC_ERR_CAPACITY, "max_agents too large for the shared-memory bitmaps");
    if (const char *e = getenv("POC_EMB_V")) { int v = atoi(e); if (v >= 1 && v <= 3) ctx->emb_v = v; }
    // third formulation, large batches: 128 threads per evaluation (641 us per launch at 1,024 replicas against 730 with 256:
    // fewer warps of a block wait at its barriers while one of them walks a short level; profiles/r02h_ab.log)
    if (ctx->emb_v == 3 && R >= 256 && !getenv("POC_EMB_G")) ctx->emb_g = 128;
    if (ctx->emb_g == 32) ctx->emb_v = 1;   // the warp-per-evaluation variant exists in the first formulation only
    {
        // second formulation: the block's dynamic shared memory is whatever its share of the SM allows (227 KB per SM,
        // 1 KB reserved per block) at the residency its block size permits; the kernel splits it into distances + staged entries
        int g = ctx->emb_g, blocks = (g == 1024) ? 2 : (g == 512) ? 4 : (g == 256) ? 8 : (ctx->emb_v == 3 ? 14 : 12);
        if (g == 64) { g = ctx->emb_g = 128; ctx->smem_emb = emb_smem_bytes(128, N); }
        if (const char *e = getenv("POC_EMB_BLOCKS")) blocks = std::max(1, std::min(16, atoi(e)));
        size_t fixed = (ctx->emb_v == 3) ? emb3_fixed_bytes(g, N) : emb2_fixed_bytes(g, N);
        size_t share = ((size_t)227 * 1024 / blocks - 1024) & ~(size_t)15;
        while (blocks > 1 && share < fixed + 8 * 1024) { blocks--; share = ((size_t)227 * 1024 / blocks - 1024) & ~(size_t)15; }
        if (share < fixed + 1024) return fail(nullptr, POC_ERR_CAPACITY, "max_agents too large for the shared-memory bitmaps");
        ctx->smem_emb2 = share;
        ctx->emb_blocks = blocks;
        ctx->smem_emb_big = ((size_t)227 * 1024 / 2 - 1024) & ~(size_t)15;   // the large-block launch: two blocks of 1,024 threads per SM
    }
    size_t RN = (size_t)R * N;
    DA(d.n_agents, R);
    DA(d.attr, RN); DA(d.birth, RN); DA(d.ncrimes, RN); DA(d.ncrimes_tick, RN); DA(d.father, RN); DA(d.new_recruit, RN);
    DA(d.propensity, RN); DA(d.tendency, RN); DA(d.sentence, RN); DA(d.arrest_w, RN);
    DA(d.u_rowptr, (size_t)R * (N + 1)); DA(d.u_ent, (size_t)R * d.ucap);
    DA(d.c_rowptr, (size_t)R * (N + 1)); DA(d.c_ent, (size_t)R * d.ccap);
    DA(d.c_rowptr_tmp, (size_t)R * (N + 1)); DA(d.c_ent_tmp, (size_t)R * d.ccap); DA(d.c_addlen, RN);
    DA(d.c_par, R);
    DA(d.u_rowptr_tmp, (size_t)R * (N + 1)); DA(d.u_ent_tmp, (size_t)R * d.ucap); DA(d.u_par, R);
    DA(d.u_mid, RN); DA(d.u_mid_tmp, RN);
    DA(d.n_children, RN); DA(d.ncrim_dead, RN); DA(d.dm_list, RN); DA(d.dm_n, (size_t)R * 4);
    d.DPcap = std::max(4 * N, 16384);
    DA(d.dm_pairs, (size_t)R * d.DPcap); DA(d.dm_defer, (size_t)R * 1024);
    DA(ctx->d_demog, 1); d.demog = ctx->d_demog;
    DA(ctx->d_labour, 1); d.labour = ctx->d_labour;
    DA(d.edu2, RN);
    DA(d.params, R); DA(d.crime_mult, R); DA(d.counters, (size_t)R * CN_COUNT); DA(d.edges, R); DA(d.stats, (size_t)R * ST_COUNT); DA(d.layer_tot, (size_t)R * 8);
    DA(d.cell_members, RN); DA(d.cell_off, (size_t)R * (POC_MAX_CELLS + 1)); DA(d.cell_strict, (size_t)R * POC_MAX_CELLS);
    DA(d.cdf, RN); DA(d.cdf_valid, R); DA(d.cdf_nz, (size_t)R * POC_MAX_CELLS);
    d.pw_cap = N / 64 + POC_MAX_CELLS + 8;
    DA(d.pw_leaf, (size_t)R * d.pw_cap); DA(d.pw_sum, (size_t)R * d.pw_cap);
    size_t RC = (size_t)R * d.Ccap, RM = (size_t)R * d.Mcap;
    DA(d.n_crimes, R); DA(d.crime_init, RC); DA(d.crime_k, RC); DA(d.crime_flags, RC);
    DA(d.cr_target, RC); DA(d.cr_nacc, RC); DA(d.cr_d, RC); DA(d.cr_state, RC); DA(d.cr_fails, RC);
    DA(d.cr_acc, RC * POC_MAX_GROUP); DA(d.search_list, RC); DA(d.n_search, R);
    DA(d.emb_list, 2 * RN); DA(d.n_emb, 2 * R); DA(d.emb_stamp, RN); DA(d.emb_val, RN);
    d.big_cap = std::min(N, 4096);
    d.emb_flags = 0;
    if (const char *e = getenv("POC_EMB_FLAGS")) d.emb_flags = atoi(e);
    DA(d.emb_big, (size_t)R * d.big_cap); DA(d.n_emb_big, R);
    DA(d.group_off, (size_t)R * (d.Ccap + 1)); DA(d.group_mem, RM);
    DA(d.pairs, (size_t)R * d.Pcap); DA(d.pair_cnt, (size_t)R * d.Pcap); DA(d.n_pairs, R);
    DA(d.pair_new_ex, (size_t)R * (d.Pcap + 1)); DA(d.ins_row, (size_t)R * d.Pcap);
    DA(d.ins_rowid, (size_t)R * d.Pcap); DA(d.ins_before, (size_t)R * (d.Pcap + 1));
    DA(d.arrested, RM); DA(d.n_arrested, R); DA(d.aw, RM); DA(d.ap, RM); DA(d.acdf, RM);
    DA(d.rp_u_init, RC); DA(d.rp_u_size, RC); DA(d.rp_fac_pick, RC);
    DA(d.rp_u_arrest, RM); DA(d.rp_u_sentence, RM); DA(d.rp_arrest_idx, RM);
    DA(d.rp_hdr, (size_t)R * 8); DA(d.rp_u_arrest_count, R); DA(d.philox_key, R);
    DA(d.scr_list, (size_t)d.scr_slots * (N + 1)); DA(d.scr_key, (size_t)R * ctx->gx * (N + 1));
    DA(d.scr_dist, (size_t)emb_slots * (N + 1));
    d.ecap = (N <= 30000) ? 65536 : 262144;
    DA(d.scr_edges, (size_t)emb_slots * d.ecap);
    DA(d.reporters, (size_t)R * d.Tcap * POC_N_REPORTERS);
    {
        double *iw = nullptr, h[256];
        DA(iw, 256);
        h[0] = 0.0;
        for (int i = 1; i < 256; i++) h[i] = 1.0 / (double)i;
        CU(cudaMemcpyAsync(iw, h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        d.invw = iw;
    }
    // staging
    // the 21 staging columns share one device block (and one pinned host mirror for downloads), 8-byte columns first
    StageCols &s = ctx->stage;
    {
        size_t o = 0, k = 0;
#define SZ(f, T) ctx->stage_off[k++] = o; o += (sizeof(T) * (size_t)N + 15) & ~(size_t)15;
        POC_STAGE_COLS(SZ)
#undef SZ
        ctx->stage_bytes = o;
        DA(ctx->stage_dev, o);
        for (int b = 0; b < 2; b++) {
            cudaEventCreateWithFlags(&ctx->stage_ev2[b], cudaEventDisableTiming);
            if (cudaHostAlloc((void **)&ctx->stage_host2[b], o, cudaHostAllocDefault) != cudaSuccess) return fail(nullptr, POC_ERR_CUDA, "cudaHostAlloc failed");
        }
        if (cudaHostAlloc((void **)&ctx->stage_host, o, cudaHostAllocDefault) != cudaSuccess) return fail(nullptr, POC_ERR_CUDA, "cudaHostAlloc failed");
        k = 0;
#define PT(f, T) s.f = (T *)(ctx->stage_dev + ctx->stage_off[k++]);
        POC_STAGE_COLS(PT)
#undef PT
    }
    for (int l = 0; l < POC_NUM_LAYERS; l++) DA(ctx->st_rowptr[l], N + 1);
    DA(ctx->st_col, d.ucap + d.ccap); DA(ctx->st_co, d.ccap);
    DA(ctx->d_err, 1); DA(ctx->d_prev_counters, (size_t)R * CN_COUNT);
    d.prev_counters = ctx->d_prev_counters;
    ctx->h_params.resize(R);
    ctx->h_nagents.assign(R, 0);
    ctx->ev.resize(EV_POOL); ctx->ev_class.assign(EV_POOL, 0);
    for (auto &evx : ctx->ev) cudaEventCreate(&evx);
    for (auto &m : ctx->marker) cudaEventCreate(&m);
    // replica groups on separate streams: the one-block-per-replica kernels of one group overlap the traversal kernels
    // of the other.  Measured (profiles/README.md): 512 replicas 851 -> 787 ms per 480-tick run with 2 groups (4: 799),
    // 128 replicas 641 -> 626 us per tick: two groups from 256 replicas on
    ctx->ngroups = (R >= 1024) ? 4 : (R >= 256) ? 2 : 1;   // 1,024 replicas: 2 groups 1,409 ms per 480 ticks, 4 groups 1,362
    if (const char *e = getenv("POC_GROUPS")) ctx->ngroups = std::max(1, std::min(8, std::min(R, atoi(e))));
    for (int g = 0; g < ctx->ngroups; g++) cudaStreamCreateWithFlags(&ctx->gstream[g], cudaStreamNonBlocking);
    for (auto &e : ctx->gev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    for (auto &e : ctx->gstag) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    if (const char *e = getenv("POC_GRAPH_TICKS")) ctx->graph_ticks = std::max(1, atoi(e));
    d.r0 = 0; d.Rg = R; d.grp = 0;
    DA(ctx->d_ti, 8); d.ti = ctx->d_ti;
    if (const char *e = getenv("POC_GRAPH")) ctx->use_graph = atoi(e) != 0;
    cudaFuncSetAttribute(k_search<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_search);
    cudaFuncSetAttribute(k_search<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_search);
    cudaFuncSetAttribute(k_cells, cudaFuncAttributeMaxDynamicSharedMemorySize, POC_MAX_CELLS * 1024 * (int)sizeof(int));
    cudaFuncSetAttribute(k_embeddedness<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)emb_smem_bytes(32, N));
    cudaFuncSetAttribute(k_embeddedness<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)emb_smem_bytes(64, N));
    cudaFuncSetAttribute(k_embeddedness<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)emb_smem_bytes(512, N));
    cudaFuncSetAttribute(k_embeddedness<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)emb_smem_bytes(1024, N));
    cudaFuncSetAttribute(k_embeddedness<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)emb_smem_bytes(128, N));
    cudaFuncSetAttribute(k_embeddedness<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)emb_smem_bytes(256, N));
#define EMB2_ATTR(G, MB) cudaFuncSetAttribute(k_emb2<G, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_emb2)
    EMB2_ATTR(128, 12); EMB2_ATTR(256, 8); EMB2_ATTR(256, 6); EMB2_ATTR(256, 5); EMB2_ATTR(512, 4); EMB2_ATTR(512, 3); EMB2_ATTR(1024, 2); EMB2_ATTR(1024, 1);
#undef EMB2_ATTR
#define EMB3_ATTR(G, MB) cudaFuncSetAttribute(k_emb3<G, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_emb2)
    EMB3_ATTR(128, 16); EMB3_ATTR(128, 14); EMB3_ATTR(128, 12); EMB3_ATTR(256, 8); EMB3_ATTR(256, 6); EMB3_ATTR(512, 4);
    cudaFuncSetAttribute(k_emb3<1024, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(ctx->smem_emb2, ctx->smem_emb_big));
#undef EMB3_ATTR
    cudaFuncSetAttribute(k_draw, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaFuncSetAttribute(k_rings_hook, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_search);
    cudaFuncSetAttribute(k_wedding, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(3 * wcap * sizeof(unsigned)));
    cudaFuncSetAttribute(k_emb_acc_mark, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_search);
    cudaFuncSetAttribute(k_emb_acc_eval, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_search);
    CU(cudaStreamSynchronize(ctx->stream));
    return POC_OK;
}

int poc_ctx_destroy(poc_ctx *ctx) {
    if (!ctx) return POC_ERR_ARG;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (auto &ge : ctx->graphs) if (ge.second) cudaGraphExecDestroy(ge.second);
    for (void *p : ctx->allocs) cudaFree(p);
    for (int b = 0; b < 2; b++) if (ctx->blob_ev[b]) cudaEventDestroy(ctx->blob_ev[b]);
    if (ctx->stage_host) cudaFreeHost(ctx->stage_host);
    for (int b = 0; b < 4; b++) { if (ctx->many_host[b]) cudaFreeHost(ctx->many_host[b]); if (ctx->many_ev[b]) cudaEventDestroy(ctx->many_ev[b]); }
    for (int b = 0; b < 2; b++) { if (ctx->stage_host2[b]) cudaFreeHost(ctx->stage_host2[b]); if (ctx->stage_ev2[b]) cudaEventDestroy(ctx->stage_ev2[b]); }
    // (a ctx whose creation failed half-way comes through here too: every handle may still be null)
    for (auto &evx : ctx->ev) if (evx) cudaEventDestroy(evx);
    for (auto &m : ctx->marker) if (m) cudaEventDestroy(m);
    for (int g = 0; g < 8; g++) if (ctx->gstream[g]) cudaStreamDestroy(ctx->gstream[g]);
    for (auto &e : ctx->gev) if (e) cudaEventDestroy(e);
    for (auto &e : ctx->gstag) if (e) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    cudaGetLastError();   // leave no stale error behind for the next ctx of this thread
    delete ctx;
    return POC_OK;
}

static int read_err(poc_ctx *ctx, const char *what);
int poc_sync(poc_ctx *ctx) {
    CHECK_CTX();
    return read_err(ctx, "poc_sync");  // also reports invalid input of uploads finished with poc_build_graph_async
}

int poc_set_timing(poc_ctx *ctx, int on) { CHECK_CTX(); ctx->timing = on != 0; return POC_OK; }
int poc_timers_reset(poc_ctx *ctx) {
    CHECK_CTX();
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->ev_used = 0; ctx->launches = 0;
    for (int i = 0; i < N_TIMERS; i++) { ctx->tm_ms[i] = 0; ctx->tm_n[i] = 0; }
    return POC_OK;
}
int poc_timers_read(poc_ctx *ctx, double *ms_out, int64_t *launches_out) {
    CHECK_CTX();
    CU(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i + 1 < ctx->ev_used; i += 2) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ctx->ev[i], ctx->ev[i + 1]) == cudaSuccess) { ctx->tm_ms[ctx->ev_class[i]] += ms; ctx->tm_n[ctx->ev_class[i]]++; }
    }
    ctx->ev_used = 0;
    for (int i = 0; i < N_TIMERS; i++) { if (ms_out) ms_out[i] = ctx->tm_ms[i]; if (launches_out) launches_out[i] = ctx->tm_n[i]; }
    return POC_OK;
}
long long poc_launch_count(poc_ctx *ctx) { return ctx ? ctx->launches : 0; }
int poc_marker(poc_ctx *ctx, int slot) {
    CHECK_CTX();
    if (slot < 0 || slot >= 16) return fail(ctx, POC_ERR_ARG, "poc_marker: slot out of range");
    CU(cudaEventRecord(ctx->marker[slot], ctx->stream));
    return POC_OK;
}
int poc_marker_elapsed_ms(poc_ctx *ctx, int from, int to, double *ms_out) {
    CHECK_CTX();
    if (from < 0 || from >= 16 || to < 0 || to >= 16 || !ms_out) return fail(ctx, POC_ERR_ARG, "poc_marker_elapsed_ms: bad argument");
    CU(cudaEventSynchronize(ctx->marker[to]));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, ctx->marker[from], ctx->marker[to]));
    *ms_out = ms;
    return POC_OK;
}

int poc_set_params(poc_ctx *ctx, int replica, const poc_params *p) {
    CHECK_CTX();
    if (!p || replica >= ctx->d.R) return fail(ctx, POC_ERR_ARG, "poc_set_params: bad argument");
    if (p->n_cells < 1 || p->n_cells > POC_MAX_CELLS || p->n_sizes < 1 || p->n_sizes > POC_MAX_TAB || p->n_sent < 1 ||
        p->n_sent > POC_MAX_TAB || p->max_accomplice_radius < 1 || p->max_accomplice_radius > MAX_RADIUS ||
        p->oc_embeddedness_radius < 1 || p->oc_embeddedness_radius > MAX_RADIUS || p->ticks_per_year < 1 ||
        p->ticks_between_intervention < 1)
        return fail(ctx, POC_ERR_ARG, "poc_set_params: table sizes / radii out of range");
    for (int a = 0; a < p->n_cells; a++)
        for (int b = a + 1; b < p->n_cells; b++)
            if (p->cell_male[a] == p->cell_male[b] && p->cell_age_from[a] <= p->cell_age_to[b] && p->cell_age_from[b] <= p->cell_age_to[a])
                return fail(ctx, POC_ERR_ARG, "poc_set_params: age/sex cells overlap");
    int r0 = replica < 0 ? 0 : replica, r1 = replica < 0 ? ctx->d.R : replica + 1;
    for (int r = r0; r < r1; r++) {
        ctx->h_params[r] = *p;
        CU(cudaMemcpyAsync(ctx->d.params + r, p, sizeof(poc_params), cudaMemcpyHostToDevice, ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));  // *p may be a temporary
    return POC_OK;
}

static int read_err(poc_ctx *ctx, const char *what) {
    int h = 0;
    CU(cudaMemcpyAsync(&h, ctx->d_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (h) {
        CU(cudaMemsetAsync(ctx->d_err, 0, sizeof(int), ctx->stream));
        return fail(ctx, POC_ERR_ARG, std::string(what) + ": invalid input (code " + std::to_string(h) +
                                          ": 1 attribute out of range, 2/3 row not ascending / bad slot / self loop)");
    }
    return POC_OK;
}

int poc_upload_agents(poc_ctx *ctx, int replica, int n, const poc_agent_cols *c) {
    CHECK_CTX();
    if (!c || replica < 0 || replica >= ctx->d.R || n < 0 || n > ctx->d.Ncap) return fail(ctx, POC_ERR_ARG, "poc_upload_agents: bad argument");
    StageCols &s = ctx->stage;
    cudaStream_t st = ctx->stream;
    // the columns are gathered in a pinned mirror of the staging block and go up in one copy; two mirrors alternate,
    // so the host packs the next replica while the device still consumes the previous one
    int mb = ctx->stage_next; ctx->stage_next ^= 1;
    if (ctx->stage_busy2[mb]) { CU(cudaEventSynchronize(ctx->stage_ev2[mb])); ctx->stage_busy2[mb] = false; }
    uint8_t *mirror = ctx->stage_host2[mb];
    {
        size_t k = 0;
#define UP(f, T) { if (!c->f && n) return fail(ctx, POC_ERR_ARG, "poc_upload_agents: column " #f " is NULL"); \
                   if (n) memcpy(mirror + ctx->stage_off[k], c->f, sizeof(T) * (size_t)n); k++; }
        POC_STAGE_COLS(UP)
#undef UP
    }
    CU(cudaMemcpyAsync(ctx->stage_dev, mirror, ctx->stage_bytes, cudaMemcpyHostToDevice, st));
    CU(cudaEventRecord(ctx->stage_ev2[mb], st)); ctx->stage_busy2[mb] = true;
    CU(cudaMemcpyAsync(ctx->d.n_agents + replica, &n, sizeof(int), cudaMemcpyHostToDevice, st));
    ctx->h_nagents[replica] = n;
    if (n) k_pack_agents<<<(n + 255) / 256, 256, 0, st>>>(ctx->d, replica, n, s, ctx->d_err);
    CU(cudaGetLastError());
    // the staging columns are reused by the next upload: order it behind this pack kernel (same stream), and
    // report invalid input at the next synchronising call (poc_build_graph / poc_sync)
    return POC_OK;
}

int poc_upload_layer(poc_ctx *ctx, int replica, int layer, const int32_t *rowptr, const int32_t *col, const int32_t *co) {
    CHECK_CTX();
    if (!rowptr || replica < 0 || replica >= ctx->d.R || layer < 0 || layer >= POC_NUM_LAYERS) return fail(ctx, POC_ERR_ARG, "poc_upload_layer: bad argument");
    int n = ctx->h_nagents[replica];
    if (ctx->st_replica != replica) { ctx->st_replica = replica; ctx->st_have = 0; ctx->st_off[0] = 0; }
    long long ne = rowptr[n];
    if (ne < 0 || (ne > 0 && !col)) return fail(ctx, POC_ERR_ARG, "poc_upload_layer: bad rowptr/col");
    cudaStream_t st = ctx->stream;
    if (layer == POC_CRIMINAL) {
        if (ne > ctx->d.ccap) return fail(ctx, POC_ERR_CAPACITY, "poc_upload_layer: criminal layer exceeds max_criminal_entries");
        int32_t *dcol = ctx->st_col + ctx->d.ucap;
        CU(cudaMemcpyAsync(ctx->st_rowptr[layer], rowptr, sizeof(int32_t) * (size_t)(n + 1), cudaMemcpyHostToDevice, st));
        if (ne) CU(cudaMemcpyAsync(dcol, col, sizeof(int32_t) * (size_t)ne, cudaMemcpyHostToDevice, st));
        if (ne && co) CU(cudaMemcpyAsync(ctx->st_co, co, sizeof(int32_t) * (size_t)ne, cudaMemcpyHostToDevice, st));
        k_pack_criminal<<<(n + 256) / 256, 256, 0, st>>>(ctx->d, replica, n, ctx->st_rowptr[layer], dcol, co ? ctx->st_co : nullptr, ctx->d_err);
        CU(cudaGetLastError());
        ctx->st_have |= 1 << layer;
        return POC_OK;
    }
    // static layers are staged back to back in upload order; poc_build_graph merges them
    int idx = 0;
    for (int l = 0; l < POC_NUM_LAYERS; l++) if (l != POC_CRIMINAL && (ctx->st_have & (1 << l))) idx++;
    if (ctx->st_have & (1 << layer)) return fail(ctx, POC_ERR_STATE, "poc_upload_layer: layer uploaded twice before poc_build_graph");
    long long off = ctx->st_off[idx];
    if (off + ne > ctx->d.ucap) return fail(ctx, POC_ERR_CAPACITY, "poc_upload_layer: static layers exceed max_static_entries");
    CU(cudaMemcpyAsync(ctx->st_rowptr[layer], rowptr, sizeof(int32_t) * (size_t)(n + 1), cudaMemcpyHostToDevice, st));
    if (ne) CU(cudaMemcpyAsync(ctx->st_col + off, col, sizeof(int32_t) * (size_t)ne, cudaMemcpyHostToDevice, st));
    ctx->st_off[idx + 1] = off + ne;
    ctx->st_have |= 1 << layer;
    ctx->st_idx[layer] = idx;
    return POC_OK;  // asynchronous: the caller's buffers must stay valid until poc_build_graph returns
}

static int build_union(poc_ctx *ctx, int replica, int n, const StageLayers &L);
static int build_graph(poc_ctx *ctx, int replica, bool wait) {
    CHECK_CTX();
    if (replica < 0 || replica >= ctx->d.R || ctx->st_replica != replica) return fail(ctx, POC_ERR_STATE, "poc_build_graph: upload the layers of this replica first");
    int all = (1 << POC_NUM_LAYERS) - 1;
    if ((ctx->st_have & all) != all) return fail(ctx, POC_ERR_STATE, "poc_build_graph: all nine layers must be uploaded");
    int n = ctx->h_nagents[replica];
    StageLayers L;
    for (int l = 0, b = 0; l < POC_NUM_LAYERS; l++) {
        if (l == POC_CRIMINAL) continue;
        int idx = ctx->st_idx[l];
        L.rowptr[b] = ctx->st_rowptr[l];
        L.col[b] = ctx->st_col + ctx->st_off[idx];
        b++;
    }
    int rc = build_union(ctx, replica, n, L);
    if (rc) return rc;
    ctx->st_have = 0; ctx->st_replica = -1; ctx->st_off[0] = 0;
    return wait ? read_err(ctx, "poc_build_graph") : POC_OK;
}

// the eight static layers (staged per-layer CSRs on the device) -> merged multiplex rows, criminal flags, layer totals
static int build_union(poc_ctx *ctx, int replica, int n, const StageLayers &L) {
    cudaStream_t st = ctx->stream;
    int *lens = ctx->d.c_addlen + (size_t)replica * ctx->d.Ncap;  // scratch
    int32_t *urp = ctx->d.u_rowptr + (size_t)replica * (ctx->d.Ncap + 1);
    if (n) {
        k_union_rows<<<(n + 127) / 128, 128, 0, st>>>(ctx->d, replica, n, L, lens, 0, ctx->d_err);
        k_exscan_i32<<<1, 1024, 0, st>>>(lens, urp, n);
        k_union_rows<<<(n + 127) / 128, 128, 0, st>>>(ctx->d, replica, n, L, lens, 1, ctx->d_err);
        k_link_criminal<<<(n + 127) / 128, 128, 0, st>>>(ctx->d, replica, n);
        k_mark_asym<<<(n + 3) / 4, 128, 0, st>>>(ctx->d, replica, n);
        CU(cudaMemsetAsync(ctx->d.layer_tot + (size_t)replica * 8, 0, sizeof(long long) * 8, st));
        k_layer_totals<<<(n + 127) / 128, 128, 0, st>>>(ctx->d, replica, n);
    } else CU(cudaMemsetAsync(urp, 0, sizeof(int32_t), st));
    CU(cudaGetLastError());
    return POC_OK;
}

// ---- packed state: one contiguous host block per state (header, the 21 agent columns, nine {rowptr, col} pairs, the
// co-offence counts), built once by the caller and uploaded with ONE copy per replica.  An ensemble uploads hundreds
// of replicas: the ~20 copies and the host-side column gathering of the call-per-array path were the end-to-end cost.
static size_t al16(size_t x) { return (x + 15) & ~(size_t)15; }
int64_t poc_packed_size(int n, const int64_t *n_entries /* [9] */) {
    if (n < 0 || !n_entries) return -1;
    size_t o = 256;
#define SZ(f, T) o += al16(sizeof(T) * (size_t)n);
    POC_STAGE_COLS(SZ)
#undef SZ
    for (int l = 0; l < POC_NUM_LAYERS; l++) { if (n_entries[l] < 0) return -1; o += al16(sizeof(int32_t) * (size_t)(n + 1)) + al16(sizeof(int32_t) * (size_t)n_entries[l]); }
    o += al16(sizeof(int32_t) * (size_t)n_entries[POC_CRIMINAL]);
    return (int64_t)o;
}

int poc_pack_state(int n, const poc_agent_cols *c, const int32_t *const *rowptr, const int32_t *const *col, const int32_t *co, void *blob, int64_t blob_bytes) {
    if (n < 0 || !c || !rowptr || !col || !blob) return POC_ERR_ARG;
    int64_t ne[POC_NUM_LAYERS];
    for (int l = 0; l < POC_NUM_LAYERS; l++) { if (!rowptr[l]) return POC_ERR_ARG; ne[l] = rowptr[l][n]; if (ne[l] < 0 || (ne[l] && !col[l])) return POC_ERR_ARG; }
    if (poc_packed_size(n, ne) > blob_bytes) return POC_ERR_CAPACITY;
    uint8_t *b = (uint8_t *)blob;
    int64_t *h = (int64_t *)b;
    memset(b, 0, 256);
    h[0] = 0x504f4331;  // "POC1"
    h[1] = n;
    for (int l = 0; l < POC_NUM_LAYERS; l++) h[2 + l] = ne[l];
    size_t o = 256;
#define PK(f, T) { if (!c->f && n) return POC_ERR_ARG; if (n) memcpy(b + o, c->f, sizeof(T) * (size_t)n); o += al16(sizeof(T) * (size_t)n); }
    POC_STAGE_COLS(PK)
#undef PK
    for (int l = 0; l < POC_NUM_LAYERS; l++) {
        memcpy(b + o, rowptr[l], sizeof(int32_t) * (size_t)(n + 1)); o += al16(sizeof(int32_t) * (size_t)(n + 1));
        if (ne[l]) memcpy(b + o, col[l], sizeof(int32_t) * (size_t)ne[l]);
        o += al16(sizeof(int32_t) * (size_t)ne[l]);
    }
    if (ne[POC_CRIMINAL]) { if (co) memcpy(b + o, co, sizeof(int32_t) * (size_t)ne[POC_CRIMINAL]); else memset(b + o, 0, sizeof(int32_t) * (size_t)ne[POC_CRIMINAL]); }
    return POC_OK;
}

int poc_upload_packed(poc_ctx *ctx, int replica, const void *blob, int64_t blob_bytes) {
    CHECK_CTX();
    DevCtx &d = ctx->d;
    if (!blob || blob_bytes < 256 || replica < 0 || replica >= d.R) return fail(ctx, POC_ERR_ARG, "poc_upload_packed: bad argument");
    const int64_t *h = (const int64_t *)blob;
    if (h[0] != 0x504f4331) return fail(ctx, POC_ERR_ARG, "poc_upload_packed: not a packed state");
    int n = (int)h[1];
    int64_t ne[POC_NUM_LAYERS], stat = 0;
    for (int l = 0; l < POC_NUM_LAYERS; l++) { ne[l] = h[2 + l]; if (l != POC_CRIMINAL) stat += ne[l]; }
    if (n < 0 || n > d.Ncap) return fail(ctx, POC_ERR_CAPACITY, "poc_upload_packed: more agents than max_agents");
    if (stat > d.ucap) return fail(ctx, POC_ERR_CAPACITY, "poc_upload_packed: static layers exceed max_static_entries");
    if (ne[POC_CRIMINAL] > d.ccap) return fail(ctx, POC_ERR_CAPACITY, "poc_upload_packed: criminal layer exceeds max_criminal_entries");
    int64_t need = poc_packed_size(n, ne);
    if (need < 0 || need > blob_bytes) return fail(ctx, POC_ERR_ARG, "poc_upload_packed: truncated block");
    if (!ctx->blob_cap) {   // sized once for the ctx capacity
        int64_t cap_ne[POC_NUM_LAYERS];
        for (int l = 0; l < POC_NUM_LAYERS; l++) cap_ne[l] = 0;
        cap_ne[0] = d.ucap; cap_ne[POC_CRIMINAL] = d.ccap;
        ctx->blob_cap = (size_t)poc_packed_size(d.Ncap, cap_ne) + 16 * POC_NUM_LAYERS;
        for (int b = 0; b < 2; b++) { DA(ctx->blob_dev[b], ctx->blob_cap); cudaEventCreateWithFlags(&ctx->blob_ev[b], cudaEventDisableTiming); }
    }
    cudaStream_t st = ctx->stream;
    int bi = ctx->blob_next; ctx->blob_next ^= 1;
    uint8_t *dev = ctx->blob_dev[bi];
    // the stream itself orders the reuse of a staging blob behind the kernels that read it (everything is on ctx->stream)
    CU(cudaMemcpyAsync(dev, blob, (size_t)need, cudaMemcpyHostToDevice, st));
    StageCols sc;
    size_t o = 256;
#define PT(f, T) sc.f = (T *)(dev + o); o += al16(sizeof(T) * (size_t)n);
    POC_STAGE_COLS(PT)
#undef PT
    StageLayers L;
    const int32_t *crp = nullptr, *ccol = nullptr;
    for (int l = 0, b = 0; l < POC_NUM_LAYERS; l++) {
        const int32_t *rp = (const int32_t *)(dev + o); o += al16(sizeof(int32_t) * (size_t)(n + 1));
        const int32_t *cl = (const int32_t *)(dev + o); o += al16(sizeof(int32_t) * (size_t)ne[l]);
        if (l == POC_CRIMINAL) { crp = rp; ccol = cl; } else { L.rowptr[b] = rp; L.col[b] = cl; b++; }
    }
    const int32_t *cco = (const int32_t *)(dev + o);
    ctx->h_nagents[replica] = n;
    k_pack_agents<<<(std::max(n, 1) + 255) / 256, 256, 0, st>>>(d, replica, n, sc, ctx->d_err);
    k_pack_criminal<<<(n + 256) / 256, 256, 0, st>>>(d, replica, n, crp, ccol, cco, ctx->d_err);
    CU(cudaGetLastError());
    return build_union(ctx, replica, n, L);   // asynchronous: poc_sync reports what the device-side input checks found
}

int poc_build_graph(poc_ctx *ctx, int replica) { return build_graph(ctx, replica, true); }
int poc_build_graph_async(poc_ctx *ctx, int replica) { return build_graph(ctx, replica, false); }

int poc_download_agents(poc_ctx *ctx, int replica, poc_agent_cols *c) {
    CHECK_CTX();
    if (!c || replica < 0 || replica >= ctx->d.R) return fail(ctx, POC_ERR_ARG, "poc_download_agents: bad argument");
    int n = ctx->h_nagents[replica];
    StageCols &s = ctx->stage;
    cudaStream_t st = ctx->stream;
    if (n) k_unpack_agents<<<(n + 255) / 256, 256, 0, st>>>(ctx->d, replica, n, s);
    CU(cudaGetLastError());
    // one device-to-host copy of the whole staging block into the pinned mirror, then plain memcpy per column
    CU(cudaMemcpyAsync(ctx->stage_host, ctx->stage_dev, ctx->stage_bytes, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    {
        size_t k = 0;
#define DN(f, T) { if (c->f && n) memcpy(c->f, ctx->stage_host + ctx->stage_off[k], sizeof(T) * (size_t)n); k++; }
        POC_STAGE_COLS(DN)
#undef DN
    }
    return POC_OK;
}

int poc_download_agents_many(poc_ctx *ctx, int replica0, int count, const poc_agent_cols *cols) {
    CHECK_CTX();
    if (!cols || replica0 < 0 || count < 0 || replica0 + count > ctx->d.R) return fail(ctx, POC_ERR_ARG, "poc_download_agents_many: bad argument");
    const int K = 4;
    cudaStream_t st = ctx->stream;
    if (!ctx->many_dev[0])
        for (int b = 0; b < K; b++) {
            DA(ctx->many_dev[b], ctx->stage_bytes);
            if (cudaHostAlloc((void **)&ctx->many_host[b], ctx->stage_bytes, cudaHostAllocDefault) != cudaSuccess) return fail(ctx, POC_ERR_CUDA, "cudaHostAlloc failed");
            CU(cudaEventCreateWithFlags(&ctx->many_ev[b], cudaEventDisableTiming));
        }
    auto hand_out = [&](int i) -> int {   // replica replica0 + i has arrived in its pinned block
        const int b = i % K, n = ctx->h_nagents[replica0 + i];
        CU(cudaEventSynchronize(ctx->many_ev[b]));
        const poc_agent_cols *c = &cols[i];
        size_t k = 0;
#define DN(f, T) { if (c->f && n) memcpy(c->f, ctx->many_host[b] + ctx->stage_off[k], sizeof(T) * (size_t)n); k++; }
        POC_STAGE_COLS(DN)
#undef DN
        return POC_OK;
    };
    for (int i = 0; i < count; i++) {
        const int b = i % K, n = ctx->h_nagents[replica0 + i];
        if (i >= K) { int rc = hand_out(i - K); if (rc) return rc; }
        StageCols s;
        size_t k = 0;
#define PT(f, T) s.f = (T *)(ctx->many_dev[b] + ctx->stage_off[k++]);
        POC_STAGE_COLS(PT)
#undef PT
        if (n) k_unpack_agents<<<(n + 255) / 256, 256, 0, st>>>(ctx->d, replica0 + i, n, s);
        CU(cudaMemcpyAsync(ctx->many_host[b], ctx->many_dev[b], ctx->stage_bytes, cudaMemcpyDeviceToHost, st));
        CU(cudaEventRecord(ctx->many_ev[b], st));
    }
    CU(cudaGetLastError());
    for (int i = std::max(0, count - K); i < count; i++) { int rc = hand_out(i); if (rc) return rc; }
    return POC_OK;
}

// layers come back through the host: state I/O, not the hot path
static int fetch_graph(poc_ctx *ctx, int replica, std::vector<int32_t> &urp, std::vector<uint32_t> &uent,
                       std::vector<int32_t> &crp, std::vector<int2> &cent) {
    int n = ctx->h_nagents[replica];
    DevCtx &d = ctx->d;
    urp.resize(n + 1); crp.resize(n + 1);
    int par = 0, upar = 0;
    CU(cudaMemcpyAsync(&par, d.c_par + replica, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(&upar, d.u_par + replica, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    const int32_t *c_rowptr = par ? d.c_rowptr_tmp : d.c_rowptr;
    const int2 *c_ent = par ? d.c_ent_tmp : d.c_ent;
    const int32_t *u_rowptr = upar ? d.u_rowptr_tmp : d.u_rowptr;
    const uint32_t *u_ent = upar ? d.u_ent_tmp : d.u_ent;
    CU(cudaMemcpyAsync(urp.data(), u_rowptr + (size_t)replica * (d.Ncap + 1), sizeof(int32_t) * (n + 1), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(crp.data(), c_rowptr + (size_t)replica * (d.Ncap + 1), sizeof(int32_t) * (n + 1), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    uent.resize(urp[n]); cent.resize(crp[n]);
    if (urp[n]) CU(cudaMemcpyAsync(uent.data(), u_ent + (size_t)replica * d.ucap, sizeof(uint32_t) * urp[n], cudaMemcpyDeviceToHost, ctx->stream));
    if (crp[n]) CU(cudaMemcpyAsync(cent.data(), c_ent + (size_t)replica * d.ccap, sizeof(int2) * crp[n], cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return POC_OK;
}

int poc_layer_entries(poc_ctx *ctx, int replica, int layer, int64_t *out) {
    CHECK_CTX();
    if (!out || replica < 0 || replica >= ctx->d.R || layer < 0 || layer >= POC_NUM_LAYERS) return fail(ctx, POC_ERR_ARG, "poc_layer_entries: bad argument");
    std::vector<int32_t> urp, crp; std::vector<uint32_t> uent; std::vector<int2> cent;
    int rc = fetch_graph(ctx, replica, urp, uent, crp, cent);
    if (rc) return rc;
    if (layer == POC_CRIMINAL) { int64_t k = 0; for (const int2 &e : cent) if (!CE_IS_DEAD(e.y)) k++; *out = k; return POC_OK; }
    uint32_t bit = 1u << layer_bit(layer);
    int64_t k = 0;
    for (uint32_t e : uent) if (ENT_MASK(e) & bit) k++;
    *out = k;
    return POC_OK;
}

int poc_download_layer(poc_ctx *ctx, int replica, int layer, int32_t *rowptr, int32_t *col, int32_t *co) {
    CHECK_CTX();
    if (!rowptr || replica < 0 || replica >= ctx->d.R || layer < 0 || layer >= POC_NUM_LAYERS) return fail(ctx, POC_ERR_ARG, "poc_download_layer: bad argument");
    std::vector<int32_t> urp, crp; std::vector<uint32_t> uent; std::vector<int2> cent;
    int rc = fetch_graph(ctx, replica, urp, uent, crp, cent);
    if (rc) return rc;
    int n = ctx->h_nagents[replica];
    int32_t p = 0;
    for (int a = 0; a < n; a++) {
        rowptr[a] = p;
        if (layer == POC_CRIMINAL) {
            for (int e = crp[a]; e < crp[a + 1]; e++) { if (CE_IS_DEAD(cent[e].y)) continue; if (col) col[p] = cent[e].x; if (co) co[p] = CE_CNT(cent[e].y); p++; }
        } else {
            uint32_t bit = 1u << layer_bit(layer);
            for (int e = urp[a]; e < urp[a + 1]; e++) if (ENT_MASK(uent[e]) & bit) { if (col) col[p] = ENT_COL(uent[e]); p++; }
        }
    }
    rowptr[n] = p;
    return POC_OK;
}

// ------------------------------------------------------------------------------------------------ hot path
static int max_n(poc_ctx *ctx) {
    if (ctx->phases & POC_PHASE_BABY) return ctx->d.Ncap;   // births add slots on the device: size the grids for the capacity
    int m = 0;
    for (int v : ctx->h_nagents) m = std::max(m, v);
    return m;
}
static int max_radius(poc_ctx *ctx) { int m = 1; for (auto &p : ctx->h_params) m = std::max(m, (int)p.max_accomplice_radius); return m; }

static int set_tick(poc_ctx *ctx, int grp, int tick, int epoch, int rep, cudaStream_t s) {
    TickInfo t{tick, epoch, rep, 0};
    CU(cudaMemcpyAsync(ctx->d_ti + grp, &t, sizeof(t), cudaMemcpyHostToDevice, s));  // pageable source: staged before return
    return POC_OK;
}

int poc_begin_tick(poc_ctx *ctx, int tick) {
    CHECK_CTX();
    int n = max_n(ctx);
    if (!n) return POC_OK;
    { int rc = set_tick(ctx, 0, tick, ctx->epoch, 0, ctx->stream); if (rc) return rc; }
    { Timed t(ctx, TM_OTHER); k_begin_tick<<<dim3((n + 255) / 256, ctx->d.R), 256, 0, ctx->stream>>>(ctx->d); }
    CU(cudaGetLastError());
    return POC_OK;
}

int poc_end_tick(poc_ctx *ctx) {
    CHECK_CTX();
    int n = max_n(ctx);
    if (!n) return POC_OK;
    { Timed t(ctx, TM_OTHER); k_end_tick<<<dim3((n + 255) / 256, ctx->d.R), 256, 0, ctx->stream>>>(ctx->d, 0); }
    CU(cudaGetLastError());
    return POC_OK;
}

static int launch_cells(poc_ctx *ctx, const DevCtx &D, cudaStream_t s, int yearly, int begin = 0) {
    int nc = 1;
    for (auto &hp : ctx->h_params) nc = std::max(nc, (int)hp.n_cells);
    { Timed t(ctx, TM_CELLS, s); k_cells<<<D.Rg, ctx->cells_threads, (size_t)nc * ctx->cells_threads * sizeof(int), s>>>(D, yearly, begin); }
    CU(cudaGetLastError());
    return POC_OK;
}

static int launch_tendency(poc_ctx *ctx, const DevCtx &D, cudaStream_t s, int mode) {
    int n = max_n(ctx);
    if (!n) return POC_OK;
    { Timed t(ctx, TM_TENDENCY, s); k_tendency_pre<<<dim3(std::max(1, std::min((n + 127) / 128, (148 * 16 + D.Rg - 1) / D.Rg)), D.Rg), 256, 0, s>>>(D, mode); }
    { Timed t(ctx, TM_TENDENCY, s); k_tendency_eps<<<D.Rg, 1024, 0, s>>>(D, mode); }
    CU(cudaGetLastError());
    return POC_OK;
}

int poc_tendency(poc_ctx *ctx) {
    CHECK_CTX();
    int rc = launch_cells(ctx, ctx->d, ctx->stream, 0);
    if (rc) return rc;
    return launch_tendency(ctx, ctx->d, ctx->stream, 1);
}

int poc_crime_multiplier(poc_ctx *ctx, double *out) {
    CHECK_CTX();
    int rc = launch_cells(ctx, ctx->d, ctx->stream, 1);
    if (rc) return rc;
    if (out) {
        CU(cudaMemcpyAsync(out, ctx->d.crime_mult, sizeof(double) * ctx->d.R, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return POC_OK;
}

int poc_set_crime_multiplier(poc_ctx *ctx, int replica, double v) {
    CHECK_CTX();
    if (replica < 0 || replica >= ctx->d.R) return fail(ctx, POC_ERR_ARG, "poc_set_crime_multiplier: bad replica");
    // pageable source: staged before the call returns, no synchronisation needed
    CU(cudaMemcpyAsync(ctx->d.crime_mult + replica, &v, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(ctx->d.cdf_valid + replica, 0, sizeof(int), ctx->stream));   // the crimes per cell follow the multiplier
    return POC_OK;
}

static void launch_emb(poc_ctx *ctx, const DevCtx &D, cudaStream_t s, int parity) {
    dim3 grid(ctx->emb_gx, D.Rg);
    if (ctx->emb_v == 3) {
        int sm = (int)ctx->smem_emb2, mb = ctx->emb_blocks;
        if (ctx->emb_g == 128 || ctx->emb_g == 256) {
            // small blocks for the evaluations that fit them, then a few large blocks for the rest (mode 0 / mode 1)
            if (ctx->emb_g == 128) {
                // register budget by residency: 32 at 16 blocks per SM, 36 at 14, 40 at 12 (POC_EMB_REGS picks one for A/B runs)
                int rb = mb >= 16 ? 16 : 12;
                if (const char *e = getenv("POC_EMB_REGS")) rb = atoi(e);
                if (rb >= 16) k_emb3<128, 16><<<grid, 128, sm, s>>>(D, parity, sm, 0);
                else if (rb >= 14) k_emb3<128, 14><<<grid, 128, sm, s>>>(D, parity, sm, 0);
                else k_emb3<128, 12><<<grid, 128, sm, s>>>(D, parity, sm, 0);
            }
            else if (mb >= 8) k_emb3<256, 8><<<grid, 256, sm, s>>>(D, parity, sm, 0);
            else k_emb3<256, 6><<<grid, 256, sm, s>>>(D, parity, sm, 0);
            ctx->launches++;
            k_emb3<1024, 2><<<dim3(std::min(4, ctx->emb_gx), D.Rg), 1024, (int)ctx->smem_emb_big, s>>>(D, parity, (int)ctx->smem_emb_big, 1);
        }
        else if (ctx->emb_g == 512) k_emb3<512, 4><<<grid, 512, sm, s>>>(D, parity, sm, 2);
        else k_emb3<1024, 2><<<grid, 1024, sm, s>>>(D, parity, sm, 2);
        return;
    }
    if (ctx->emb_v == 2) {
        int sm = (int)ctx->smem_emb2, mb = ctx->emb_blocks;
        if (ctx->emb_g == 128) k_emb2<128, 12><<<grid, 128, sm, s>>>(D, parity, sm);
        else if (ctx->emb_g == 512) { if (mb >= 4) k_emb2<512, 4><<<grid, 512, sm, s>>>(D, parity, sm); else k_emb2<512, 3><<<grid, 512, sm, s>>>(D, parity, sm); }
        else if (ctx->emb_g == 1024) { if (mb >= 2) k_emb2<1024, 2><<<grid, 1024, sm, s>>>(D, parity, sm); else k_emb2<1024, 1><<<grid, 1024, sm, s>>>(D, parity, sm); }
        else if (mb >= 8) k_emb2<256, 8><<<grid, 256, sm, s>>>(D, parity, sm);
        else if (mb >= 6) k_emb2<256, 6><<<grid, 256, sm, s>>>(D, parity, sm);
        else k_emb2<256, 5><<<grid, 256, sm, s>>>(D, parity, sm);
        return;
    }
    if (ctx->emb_g == 32) k_embeddedness<32><<<grid, 256, ctx->smem_emb, s>>>(D, parity);
    else if (ctx->emb_g == 64) k_embeddedness<64><<<grid, 64, ctx->smem_emb, s>>>(D, parity);
    else if (ctx->emb_g == 128) k_embeddedness<128><<<grid, 128, ctx->smem_emb, s>>>(D, parity);
    else if (ctx->emb_g == 512) k_embeddedness<512><<<grid, 512, ctx->smem_emb, s>>>(D, parity);
    else if (ctx->emb_g == 1024) k_embeddedness<1024><<<grid, 1024, ctx->smem_emb, s>>>(D, parity);
    else k_embeddedness<256><<<grid, 256, ctx->smem_emb, s>>>(D, parity);
}

// search / embeddedness rounds shared by commit_crimes and the find_accomplices hook
static int launch_search_rounds(poc_ctx *ctx, const DevCtx &D, cudaStream_t s) {
    int rounds = max_radius(ctx);
    dim3 grid(ctx->gx, D.Rg);
    for (int rd = 0; rd <= rounds; rd++) {
        {
            Timed t(ctx, TM_SEARCH, s);
            if (D.emb_flags & 2) k_search<true><<<grid, ctx->search_threads, ctx->smem_search, s>>>(D, rd & 1);
            else k_search<false><<<grid, ctx->search_threads, ctx->smem_search, s>>>(D, rd & 1);
        }
        if (rd < rounds) { Timed t(ctx, TM_EMB, s); launch_emb(ctx, D, s, rd & 1); }
    }
    CU(cudaGetLastError());
    return POC_OK;
}

// one replica range, one stream: reset_oc_embeddedness + commit_crimes.  ctx->epoch must have been bumped.
static int launch_commit(poc_ctx *ctx, const DevCtx &D, cudaStream_t s, bool cells_done, bool report, cudaEvent_t after_draw = nullptr,
                         int end_tick = 0) {
    if (!cells_done) { int rc = launch_cells(ctx, D, s, 0); if (rc) return rc; }
    {
        // weights of one replica in shared memory when they fit (k_draw falls back to the global cdf array otherwise):
        // cells padded to multiples of eight doubles, plus one byte per eight positions
        int nd = max_n(ctx) + 8 * POC_MAX_CELLS;
        size_t draw_smem = (size_t)nd * sizeof(double) + nd / 8 + 16;
        int allow_fit = 1;
        if (draw_smem > 160 * 1024) { nd = 16384; draw_smem = (size_t)nd * sizeof(double) + 16; allow_fit = 0; }   // window mode
        Timed t(ctx, TM_DRAW, s);
        k_draw<<<D.Rg, ctx->draw_threads, draw_smem, s>>>(D, nd, allow_fit);
    }
    if (after_draw) CU(cudaEventRecord(after_draw, s));
    int rc = launch_search_rounds(ctx, D, s);
    if (rc) return rc;
    { Timed t(ctx, TM_COMMIT, s); k_commit<<<D.Rg, ctx->commit_threads, 0, s>>>(D); }
    { Timed t(ctx, TM_ARREST, s); k_recruit_arrest<<<D.Rg, 256, 0, s>>>(D, end_tick); }
    // the reporter row of the tick is written after the end-of-tick bookkeeping (model.py:286-287)
    if (report) { Timed t(ctx, TM_OTHER, s); k_report<<<D.Rg, 512, 0, s>>>(D, end_tick == 2 ? 1 : 0); }
    CU(cudaGetLastError());
    return POC_OK;
}

// the demography phases of one tick for one replica range (kernels in poc_demog.cuh); n = agent slots the grids cover
static int launch_phases(poc_ctx *ctx, const DevCtx &D, cudaStream_t s, int mask, int n, int force_labour = 0) {
    if (!mask || !n) return POC_OK;
    if (!ctx->demog_set) return fail(ctx, POC_ERR_STATE, "demography phases requested before poc_set_demography");
    if (mask & POC_PHASE_LABOUR) {   // model.py:251-256 (after the tendencies and the multiplier of a yearly tick)
        if (!ctx->labour_set) return fail(ctx, POC_ERR_STATE, "POC_PHASE_LABOUR requested before poc_set_labour");
        Timed t(ctx, TM_OTHER, s);
        k_labour<<<dim3(std::min((n + 255) / 256, 4), D.Rg), 256, 0, s>>>(D, force_labour);
    }
    if (mask & POC_PHASE_WEDDING) {   // model.py:271 (before the crimes of the tick)
        Timed t(ctx, TM_OTHER, s);
        k_wedding<<<D.Rg, 256, 3 * (((size_t)D.Ncap + 31) / 32) * sizeof(unsigned), s>>>(D);
        k_multiplex_insert<<<D.Rg, 1024, 0, s>>>(D);
    }
    if (mask & POC_PHASE_RETIRE) {   // model.py:274
        Timed t(ctx, TM_OTHER, s);
        k_retire<<<dim3((n + 255) / 256, D.Rg), 256, 0, s>>>(D);
    }
    if (mask & POC_PHASE_BABY) {
        Timed t(ctx, TM_OTHER, s);
        k_make_baby<<<D.Rg, 1024, 0, s>>>(D);
        k_multiplex_insert<<<D.Rg, 1024, 0, s>>>(D);
    }
    if (mask & POC_PHASE_XFRIENDS) {   // model.py:276
        Timed t(ctx, TM_OTHER, s);
        k_remove_excess<<<dim3((n + 127) / 128, D.Rg), 128, 0, s>>>(D, MB_FRIEND, ST_XFRIEND, 1);
        k_apply_removals<<<dim3(8, D.Rg), 256, 0, s>>>(D, MB_FRIEND, 5);
        k_removals_done<<<(D.Rg + 127) / 128, 128, 0, s>>>(D);
    }
    if (mask & POC_PHASE_XPROF) {   // model.py:277
        Timed t(ctx, TM_OTHER, s);
        k_remove_excess<<<dim3((n + 127) / 128, D.Rg), 128, 0, s>>>(D, MB_PROF, ST_XPROF, 0);
        k_apply_removals<<<dim3(8, D.Rg), 256, 0, s>>>(D, MB_PROF, 6);
        k_removals_done<<<(D.Rg + 127) / 128, 128, 0, s>>>(D);
    }
    if (mask & POC_PHASE_FRIENDS) {
        Timed t(ctx, TM_OTHER, s);
        k_make_friends<<<dim3((n + 127) / 128, D.Rg), 128, 0, s>>>(D);
        k_apply_friends<<<dim3(8, D.Rg), 256, 0, s>>>(D);
        k_friends_done<<<(D.Rg + 127) / 128, 128, 0, s>>>(D);
        k_multiplex_insert<<<D.Rg, 1024, 0, s>>>(D);   // entries whose reverse slot was missing (rare; leaves at once otherwise)
    }
    if (mask & POC_PHASE_DIE) {
        Timed t(ctx, TM_OTHER, s);
        k_die_mark<<<dim3((n + 255) / 256, D.Rg), 256, 0, s>>>(D);
        k_die_apply<<<D.Rg, 256, 0, s>>>(D);
    }
    CU(cudaGetLastError());
    return POC_OK;
}

static int refresh_counts(poc_ctx *ctx) {   // births add agent slots on the device
    CU(cudaMemcpyAsync(ctx->h_nagents.data(), ctx->d.n_agents, sizeof(int) * ctx->d.R, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return POC_OK;
}

int poc_set_demography(poc_ctx *ctx, const poc_demography *dm) {
    CHECK_CTX();
    static_assert(sizeof(poc_demography) == sizeof(DevDemog), "poc_demography / DevDemog out of sync");
    if (!dm || dm->n_prop < 1 || dm->n_prop > 1024 || dm->mort_max_age < 0 || dm->mort_max_age > 127 || dm->retirement_age < 0) return fail(ctx, POC_ERR_ARG, "poc_set_demography: bad argument");
    CU(cudaMemcpyAsync(ctx->d_demog, dm, sizeof(DevDemog), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->demog_set = true;
    return POC_OK;
}

int poc_set_phases(poc_ctx *ctx, int mask) {
    CHECK_CTX();
    if (mask & ~POC_PHASE_EVERY) return fail(ctx, POC_ERR_ARG, "poc_set_phases: unknown phase");
    if (mask && !ctx->demog_set) return fail(ctx, POC_ERR_STATE, "poc_set_phases: poc_set_demography first");
    if ((mask & POC_PHASE_LABOUR) && !ctx->labour_set) return fail(ctx, POC_ERR_STATE, "poc_set_phases: poc_set_labour first");
    ctx->phases = mask;
    return POC_OK;
}

int poc_run_phase(poc_ctx *ctx, int phase, int tick, const uint64_t *keys) {
    CHECK_CTX();
    if (phase <= 0 || (phase & ~POC_PHASE_EVERY) || (phase & (phase - 1))) return fail(ctx, POC_ERR_ARG, "poc_run_phase: one phase at a time");
    if (keys) CU(cudaMemcpyAsync(ctx->d.philox_key, keys, sizeof(uint64_t) * ctx->d.R, cudaMemcpyHostToDevice, ctx->stream));
    TickInfo t{tick, ctx->epoch, 0, 0};
    CU(cudaMemcpyAsync(ctx->d_ti, &t, sizeof(t), cudaMemcpyHostToDevice, ctx->stream));
    int rc = launch_phases(ctx, ctx->d, ctx->stream, phase, ctx->d.Ncap, 1);   // the hook runs the labour block whatever the tick
    if (rc) return rc;
    return refresh_counts(ctx);
}

int poc_sizeof_labour(void) { return (int)sizeof(poc_labour); }

int poc_set_labour(poc_ctx *ctx, const poc_labour *lb) {
    CHECK_CTX();
    static_assert(sizeof(poc_labour) == sizeof(DevLabour), "poc_labour / DevLabour out of sync");
    if (!lb || lb->n_levels < 1 || lb->n_levels > 7) return fail(ctx, POC_ERR_ARG, "poc_set_labour: bad argument");
    for (int k = 0; k < 8; k++)
        for (int m = 0; m < 2; m++) {
            if (lb->work_n[k][m] < 0 || lb->work_n[k][m] > 8 || lb->wealth_n[k][m] < 0 || lb->wealth_n[k][m] > 8)
                return fail(ctx, POC_ERR_ARG, "poc_set_labour: table too long");
            for (int i = 0; i < lb->work_n[k][m]; i++)      // a work status indexes the wealth table, both go into 4-bit fields
                if (lb->work_val[k][m][i] < 0 || lb->work_val[k][m][i] > 7) return fail(ctx, POC_ERR_ARG, "poc_set_labour: work status out of range");
            for (int i = 0; i < lb->wealth_n[k][m]; i++)
                if (lb->wealth_val[k][m][i] < 0 || lb->wealth_val[k][m][i] > 15) return fail(ctx, POC_ERR_ARG, "poc_set_labour: wealth level out of range");
        }
    CU(cudaMemcpyAsync(ctx->d_labour, lb, sizeof(DevLabour), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->labour_set = true;
    return POC_OK;
}

int poc_upload_education(poc_ctx *ctx, int replica, const int8_t *max_edu, const int8_t *school_level) {
    CHECK_CTX();
    if (!max_edu || !school_level || replica < 0 || replica >= ctx->d.R) return fail(ctx, POC_ERR_ARG, "poc_upload_education: bad argument");
    const int n = ctx->h_nagents[replica];
    std::vector<uint8_t> h((size_t)n);
    for (int a = 0; a < n; a++) {
        if (max_edu[a] < 0 || max_edu[a] > 15 || school_level[a] < 0 || school_level[a] > 15) return fail(ctx, POC_ERR_ARG, "poc_upload_education: level out of range");
        h[a] = (uint8_t)(max_edu[a] | (school_level[a] << 4));
    }
    CU(cudaMemcpyAsync(ctx->d.edu2 + (size_t)replica * ctx->d.Ncap, h.data(), (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return POC_OK;
}

int poc_download_education(poc_ctx *ctx, int replica, int8_t *max_edu, int8_t *school_level) {
    CHECK_CTX();
    if (!max_edu || !school_level || replica < 0 || replica >= ctx->d.R) return fail(ctx, POC_ERR_ARG, "poc_download_education: bad argument");
    const int n = ctx->h_nagents[replica];
    std::vector<uint8_t> h((size_t)n);
    CU(cudaMemcpyAsync(h.data(), ctx->d.edu2 + (size_t)replica * ctx->d.Ncap, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (int a = 0; a < n; a++) { max_edu[a] = (int8_t)(h[a] & 15); school_level[a] = (int8_t)(h[a] >> 4); }
    return POC_OK;
}

int poc_upload_children(poc_ctx *ctx, int replica, const int32_t *nc) {
    CHECK_CTX();
    if (!nc || replica < 0 || replica >= ctx->d.R) return fail(ctx, POC_ERR_ARG, "poc_upload_children: bad argument");
    CU(cudaMemcpyAsync(ctx->d.n_children + (size_t)replica * ctx->d.Ncap, nc, sizeof(int32_t) * ctx->h_nagents[replica], cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return POC_OK;
}

int poc_download_children(poc_ctx *ctx, int replica, int32_t *nc) {
    CHECK_CTX();
    if (!nc || replica < 0 || replica >= ctx->d.R) return fail(ctx, POC_ERR_ARG, "poc_download_children: bad argument");
    CU(cudaMemcpyAsync(nc, ctx->d.n_children + (size_t)replica * ctx->d.Ncap, sizeof(int32_t) * ctx->h_nagents[replica], cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return POC_OK;
}

int poc_n_agents(poc_ctx *ctx, int replica, int32_t *out) {
    CHECK_CTX();
    if (!out || replica < 0 || replica >= ctx->d.R) return fail(ctx, POC_ERR_ARG, "poc_n_agents: bad argument");
    *out = ctx->h_nagents[replica];
    return POC_OK;
}

static int upload_replay(poc_ctx *ctx, const poc_replay *const *replay, const uint64_t *keys) {
    DevCtx &d = ctx->d;
    cudaStream_t st = ctx->stream;
    std::vector<int32_t> hdr((size_t)d.R * 8, 0);
    std::vector<double> uac(d.R, 0.0);
    for (int r = 0; r < d.R; r++) {
        const poc_replay *rp = replay ? replay[r] : nullptr;
        if (!rp) continue;
        if (rp->n_crimes > d.Ccap || rp->n_u_arrest > d.Mcap || rp->n_u_sentence > d.Mcap || rp->n_arrest_idx > d.Mcap)
            return fail(ctx, POC_ERR_CAPACITY, "replay arrays exceed capacity");
        int32_t *h = &hdr[(size_t)r * 8];
        h[0] = 1; h[1] = rp->n_crimes; h[2] = rp->n_u_arrest; h[3] = rp->n_u_sentence;
        h[4] = (rp->arrest_idx && rp->n_arrest_idx >= 0) ? rp->n_arrest_idx : -1;
        uac[r] = rp->u_arrest_count;
        if (rp->n_crimes) {
            CU(cudaMemcpyAsync(d.rp_u_init + (size_t)r * d.Ccap, rp->u_init, sizeof(double) * rp->n_crimes, cudaMemcpyHostToDevice, st));
            CU(cudaMemcpyAsync(d.rp_u_size + (size_t)r * d.Ccap, rp->u_size, sizeof(double) * rp->n_crimes, cudaMemcpyHostToDevice, st));
            if (rp->fac_pick) CU(cudaMemcpyAsync(d.rp_fac_pick + (size_t)r * d.Ccap, rp->fac_pick, sizeof(int32_t) * rp->n_crimes, cudaMemcpyHostToDevice, st));
            else CU(cudaMemsetAsync(d.rp_fac_pick + (size_t)r * d.Ccap, 0xff, sizeof(int32_t) * rp->n_crimes, st));
        }
        if (rp->n_u_arrest) CU(cudaMemcpyAsync(d.rp_u_arrest + (size_t)r * d.Mcap, rp->u_arrest, sizeof(double) * rp->n_u_arrest, cudaMemcpyHostToDevice, st));
        if (rp->n_u_sentence) CU(cudaMemcpyAsync(d.rp_u_sentence + (size_t)r * d.Mcap, rp->u_sentence, sizeof(double) * rp->n_u_sentence, cudaMemcpyHostToDevice, st));
        if (h[4] > 0) CU(cudaMemcpyAsync(d.rp_arrest_idx + (size_t)r * d.Mcap, rp->arrest_idx, sizeof(int32_t) * h[4], cudaMemcpyHostToDevice, st));
    }
    CU(cudaMemcpyAsync(d.rp_hdr, hdr.data(), sizeof(int32_t) * hdr.size(), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d.rp_u_arrest_count, uac.data(), sizeof(double) * d.R, cudaMemcpyHostToDevice, st));
    if (keys) CU(cudaMemcpyAsync(d.philox_key, keys, sizeof(uint64_t) * d.R, cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));  // hdr / uac are locals
    return POC_OK;
}

static int read_tick_out(poc_ctx *ctx, poc_tick_out *out, poc_tick_records *rec) {
    DevCtx &d = ctx->d;
    std::vector<int> cn((size_t)d.R * CN_COUNT), prev((size_t)d.R * CN_COUNT);
    std::vector<unsigned long long> edges(d.R);
    CU(cudaMemcpyAsync(cn.data(), d.counters, sizeof(int) * cn.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(prev.data(), ctx->d_prev_counters, sizeof(int) * prev.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(edges.data(), d.edges, sizeof(unsigned long long) * d.R, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemsetAsync(d.edges, 0, sizeof(unsigned long long) * d.R, ctx->stream));
    for (int r = 0; r < d.R && out; r++) {
        const int *c = &cn[(size_t)r * CN_COUNT], *p = &prev[(size_t)r * CN_COUNT];
        poc_tick_out &o = out[r];
        o.n_crimes = c[CN_TICK_CRIMES]; o.n_members = c[CN_TICK_MEMBERS]; o.n_arrests = c[CN_TICK_ARRESTS];
        o.n_emb_evals = c[CN_TICK_EMB];
        o.d_number_crimes = c[CN_NUMBER_CRIMES] - p[CN_NUMBER_CRIMES];
        o.d_crime_size_fails = c[CN_SIZE_FAILS] - p[CN_SIZE_FAILS];
        o.d_facilitator_crimes = c[CN_FAC_CRIMES] - p[CN_FAC_CRIMES];
        o.d_facilitator_fails = c[CN_FAC_FAILS] - p[CN_FAC_FAILS];
        o.d_big_crime_from_small_fish = c[CN_BIG_CRIME] - p[CN_BIG_CRIME];
        o.d_offspring_recruited = c[CN_OFFSPRING_RECRUITED] - p[CN_OFFSPRING_RECRUITED];
        o.d_protected_recruited = c[CN_PROTECTED_RECRUITED] - p[CN_PROTECTED_RECRUITED];
        o.d_people_jailed = c[CN_PEOPLE_JAILED] - p[CN_PEOPLE_JAILED];
        o.n_recruited = c[CN_TICK_RECRUITED]; o.error = c[CN_ERROR];
        o.traversed_edges = (int64_t)edges[r];
    }
    if (rec) {
        int ncr = cn[CN_TICK_CRIMES], nm = cn[CN_TICK_MEMBERS], na = cn[CN_TICK_ARRESTS];
        if (ncr > rec->cap_crimes || nm > rec->cap_members || na > rec->cap_arrests) return fail(ctx, POC_ERR_CAPACITY, "poc_tick_records buffers too small");
        cudaStream_t st = ctx->stream;
        if (ncr && rec->initiator) CU(cudaMemcpyAsync(rec->initiator, d.crime_init, sizeof(int32_t) * ncr, cudaMemcpyDeviceToHost, st));
        if (ncr && rec->k) CU(cudaMemcpyAsync(rec->k, d.crime_k, sizeof(int32_t) * ncr, cudaMemcpyDeviceToHost, st));
        if (rec->group_off) CU(cudaMemcpyAsync(rec->group_off, d.group_off, sizeof(int32_t) * (ncr + 1), cudaMemcpyDeviceToHost, st));
        if (nm && rec->group_mem) CU(cudaMemcpyAsync(rec->group_mem, d.group_mem, sizeof(int32_t) * nm, cudaMemcpyDeviceToHost, st));
        if (na && rec->arrested) CU(cudaMemcpyAsync(rec->arrested, d.arrested, sizeof(int32_t) * na, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    return POC_OK;
}

int poc_commit_crimes(poc_ctx *ctx, int tick, const poc_replay *const *replay, const uint64_t *keys, poc_tick_out *out,
                      poc_tick_records *rec) {
    CHECK_CTX();
    int rc = upload_replay(ctx, replay, keys);
    if (rc) return rc;
    ctx->epoch++;
    rc = set_tick(ctx, 0, tick, ctx->epoch, 0, ctx->stream);
    if (rc) return rc;
    rc = launch_commit(ctx, ctx->d, ctx->stream, false, false);
    if (rc) return rc;
    if (out || rec) return read_tick_out(ctx, out, rec);
    return POC_OK;
}

static int run_ticks_impl(poc_ctx *ctx, int tick0, int n_ticks, const uint64_t *keys, double *reporters_out, bool report) {
    DevCtx &d = ctx->d;
    if (n_ticks < 0) return fail(ctx, POC_ERR_ARG, "poc_run_ticks: n_ticks < 0");
    int rc = upload_replay(ctx, nullptr, keys);
    if (rc) return rc;
    if (report && n_ticks > d.Tcap) {
        double *p = nullptr;
        DA(p, (size_t)d.R * n_ticks * POC_N_REPORTERS);
        d.reporters = p; d.Tcap = n_ticks;
    }
    int n = max_n(ctx);
    int G = ctx->timing ? 1 : ctx->ngroups;   // per-kernel timing wants the kernels alone on the GPU
    for (int g = 0; g < G; g++) { rc = set_tick(ctx, g, tick0, ctx->epoch + 1, 0, ctx->stream); if (rc) return rc; }
    ctx->epoch += n_ticks;
    // one tick of every replica group; each group's pipeline on its own stream, forked from / joined to the
    // main stream.  Kernels take the tick from the device-side TickInfo, so the sequence has no per-tick
    // arguments and can be captured once and replayed as a CUDA graph.
    // K ticks per replica group as one independent chain per stream.  Group g starts its chain when group g-1
    // has finished its first crime draw, so the groups run out of phase and the latency-bound one-block-per-
    // replica kernels of one group overlap the traversal kernels of the others.
    auto enqueue_ticks = [&](int K) -> int {
        CU(cudaEventRecord(ctx->gev[8], ctx->stream));
        for (int g = 0; g < G; g++) {
            DevCtx D = d;
            D.r0 = (int)((long long)d.R * g / G);
            D.Rg = (int)((long long)d.R * (g + 1) / G) - D.r0;
            D.grp = g;
            if (D.Rg <= 0) continue;
            cudaStream_t s = ctx->gstream[g];
            CU(cudaStreamWaitEvent(s, ctx->gev[8], 0));
            if (g > 0) CU(cudaStreamWaitEvent(s, ctx->gstag[g - 1], 0));
            for (int k = 0; k < K; k++) {
                // (fusing begin/end-of-tick bookkeeping into the one-block-per-replica kernels measured slower:
                //  they are grid-wide passes, and the single blocks are the critical path)
                if (n) { Timed t(ctx, TM_OTHER, s); k_begin_tick<<<dim3((n + 255) / 256, D.Rg), 256, 0, s>>>(D); }
                int rc2 = launch_cells(ctx, D, s, 2, 0);
                if (rc2) return rc2;
                rc2 = launch_tendency(ctx, D, s, 2);
                if (rc2) return rc2;
                rc2 = launch_phases(ctx, D, s, ctx->phases & POC_PHASE_LABOUR, n);    // model.py:251-256
                if (rc2) return rc2;
                rc2 = launch_phases(ctx, D, s, ctx->phases & POC_PHASE_WEDDING, n);   // model.py:271
                if (rc2) return rc2;
                rc2 = launch_commit(ctx, D, s, true, false, (k == 0 && g + 1 < G) ? ctx->gstag[g] : nullptr, 0);
                if (rc2) return rc2;
                // model.py:274-278: retire_persons, make_baby, remove_excess_friends / _professional_links, make_friends
                rc2 = launch_phases(ctx, D, s, ctx->phases & ~(POC_PHASE_DIE | POC_PHASE_WEDDING | POC_PHASE_LABOUR), n);
                if (rc2) return rc2;
                if (n) { Timed t(ctx, TM_OTHER, s); k_end_tick<<<dim3((n + 255) / 256, D.Rg), 256, 0, s>>>(D, 0); }
                rc2 = launch_phases(ctx, D, s, ctx->phases & POC_PHASE_DIE, n);   // model.py:284
                if (rc2) return rc2;
                // reporter row after the end-of-tick bookkeeping (model.py:286-287); it also moves the tick state on
                { Timed t(ctx, TM_OTHER, s); k_report<<<D.Rg, 512, 0, s>>>(D, report ? 2 : 3); }
            }
            CU(cudaEventRecord(ctx->gev[g], s));
            CU(cudaStreamWaitEvent(ctx->stream, ctx->gev[g], 0));
        }
        return POC_OK;
    };
    if (ctx->use_graph && !ctx->timing && n_ticks > 1 && n > 0) {
        int K = std::min(ctx->graph_ticks, n_ticks);
        int done = 0;
        while (done < n_ticks) {
            int kk = std::min(K, n_ticks - done);   // the tail gets its own (smaller) graph
            int reps = (kk == K) ? (n_ticks - done) / K : 1;
            // instantiated graphs are kept across calls: an ensemble driver calls poc_run_ticks once per step with the
            // same shape, and capture + instantiation of ~450 nodes costs more than replaying them once
            int nc_max = 1;
            for (auto &hp : ctx->h_params) nc_max = std::max(nc_max, (int)hp.n_cells);
            // everything the captured launch shapes depend on
            poc_ctx::GraphKey key{kk, report ? 1 : 0, n, G, max_radius(ctx) * 1024 + nc_max, d.Tcap, ctx->phases, (const void *)d.reporters};
            cudaGraphExec_t exec = nullptr;
            long long per_graph = 0;
            for (size_t gi = 0; gi < ctx->graphs.size(); gi++)
                if (ctx->graphs[gi].first == key) { exec = ctx->graphs[gi].second; per_graph = ctx->graph_launches[gi]; }
            if (!exec) {
                cudaGraph_t graph = nullptr;
                long long launches_before = ctx->launches;
                CU(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
                rc = enqueue_ticks(kk);
                cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
                if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
                if (ce != cudaSuccess) return fail(ctx, POC_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
                cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
                cudaGraphDestroy(graph);
                if (ie != cudaSuccess) return fail(ctx, POC_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie));
                per_graph = ctx->launches - launches_before;
                ctx->launches = launches_before;
                if (ctx->graphs.size() >= 16) { cudaGraphExecDestroy(ctx->graphs[0].second); ctx->graphs.erase(ctx->graphs.begin()); ctx->graph_launches.erase(ctx->graph_launches.begin()); }
                ctx->graphs.push_back({key, exec});
                ctx->graph_launches.push_back(per_graph);
            }
            for (int i = 0; i < reps; i++) {
                cudaError_t le = cudaGraphLaunch(exec, ctx->stream);
                if (le != cudaSuccess) return fail(ctx, POC_ERR_CUDA, std::string("cudaGraphLaunch: ") + cudaGetErrorString(le));
            }
            ctx->launches += per_graph * reps;
            done += kk * reps;
        }
    } else {
        rc = enqueue_ticks(n_ticks);
        if (rc) return rc;
    }
    CU(cudaGetLastError());
    if (ctx->phases & POC_PHASE_BABY) { rc = refresh_counts(ctx); if (rc) return rc; }
    if (reporters_out && n_ticks) {
        // device layout [R][Tcap][K] -> host [R][n_ticks][K]
        for (int r = 0; r < d.R; r++)
            CU(cudaMemcpyAsync(reporters_out + (size_t)r * n_ticks * POC_N_REPORTERS, d.reporters + (size_t)r * d.Tcap * POC_N_REPORTERS,
                               sizeof(double) * (size_t)n_ticks * POC_N_REPORTERS, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return POC_OK;
}

int poc_run_ticks(poc_ctx *ctx, int tick0, int n_ticks, const uint64_t *keys, double *reporters_out) {
    CHECK_CTX();
    return run_ticks_impl(ctx, tick0, n_ticks, keys, reporters_out, reporters_out != nullptr);
}

int poc_run_ticks_device(poc_ctx *ctx, int tick0, int n_ticks, const uint64_t *keys, double **rows_dev, int *ticks_per_replica) {
    CHECK_CTX();
    if (!rows_dev || !ticks_per_replica) return fail(ctx, POC_ERR_ARG, "poc_run_ticks_device: NULL output");
    int rc = run_ticks_impl(ctx, tick0, n_ticks, keys, nullptr, true);
    if (rc) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    *rows_dev = ctx->d.reporters; *ticks_per_replica = ctx->d.Tcap;
    return POC_OK;
}

int poc_report(poc_ctx *ctx, double *out /* [R][POC_N_REPORTERS] */) {
    CHECK_CTX();
    DevCtx &d = ctx->d;
    if (!out) return fail(ctx, POC_ERR_ARG, "poc_report: out is NULL");
    int rep0 = 0;
    // the row index lives in the tick state of group 0; write row 0 without touching the tick
    CU(cudaMemcpyAsync(&ctx->d_ti[0].rep, &rep0, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    { Timed t(ctx, TM_OTHER); k_report<<<d.R, 512, 0, ctx->stream>>>(d, 0); }
    CU(cudaGetLastError());
    for (int r = 0; r < d.R; r++)
        CU(cudaMemcpyAsync(out + (size_t)r * POC_N_REPORTERS, d.reporters + (size_t)r * d.Tcap * POC_N_REPORTERS,
                           sizeof(double) * POC_N_REPORTERS, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return POC_OK;
}

int poc_read_stats(poc_ctx *ctx, uint64_t *stats_out /* [R][16] */, int reset) {
    CHECK_CTX();
    DevCtx &d = ctx->d;
    if (stats_out) CU(cudaMemcpyAsync(stats_out, d.stats, sizeof(unsigned long long) * (size_t)d.R * ST_COUNT, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (reset) {
        CU(cudaMemsetAsync(d.stats, 0, sizeof(unsigned long long) * (size_t)d.R * ST_COUNT, ctx->stream));
        CU(cudaMemsetAsync(d.edges, 0, sizeof(unsigned long long) * (size_t)d.R, ctx->stream));
    }
    return POC_OK;
}

int poc_read_counters(poc_ctx *ctx, int32_t *counters_out /* [R][24] */, uint64_t *edges_out /* [R] */) {
    CHECK_CTX();
    DevCtx &d = ctx->d;
    if (counters_out) CU(cudaMemcpyAsync(counters_out, d.counters, sizeof(int) * (size_t)d.R * CN_COUNT, cudaMemcpyDeviceToHost, ctx->stream));
    if (edges_out) CU(cudaMemcpyAsync(edges_out, d.edges, sizeof(unsigned long long) * d.R, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return POC_OK;
}

// ------------------------------------------------------------------------------------------------ stage hooks
static int hook_bufs(poc_ctx *ctx, long long n) {
    if (n <= ctx->hook_cap) return POC_OK;
    long long cap = std::max<long long>(n, 1024);
    DA(ctx->d_hook_a, cap); DA(ctx->d_hook_b, cap); DA(ctx->d_hook_out, cap);
    ctx->hook_cap = cap;
    return POC_OK;
}

int poc_rings(poc_ctx *ctx, int replica, int n_src, const int32_t *src, int dd, int exact, int32_t *out_off, int32_t *out_mem,
              int64_t cap_mem) {
    CHECK_CTX();
    if (replica < 0 || replica >= ctx->d.R || n_src < 0 || !src || !out_off || dd < 1 || dd > MAX_RADIUS) return fail(ctx, POC_ERR_ARG, "poc_rings: bad argument");
    int n = ctx->h_nagents[replica];
    for (int i = 0; i < n_src; i++) if (src[i] < 0 || src[i] >= n) return fail(ctx, POC_ERR_ARG, "poc_rings: bad source slot");
    int rc = hook_bufs(ctx, std::max<long long>(n_src, 1));
    if (rc) return rc;
    int32_t *d_mem = nullptr;
    long long per = n;  // a ring cannot exceed n members
    CU(cudaMalloc((void **)&d_mem, sizeof(int32_t) * std::max<size_t>((size_t)n_src * per, 1)));
    cudaStream_t st = ctx->stream;
    CU(cudaMemcpyAsync(ctx->d_hook_a, src, sizeof(int32_t) * n_src, cudaMemcpyHostToDevice, st));
    int grid = std::min(std::max(n_src, 1), ctx->d.scr_slots);
    k_rings_hook<<<grid, 256, ctx->smem_search, st>>>(ctx->d, replica, n_src, ctx->d_hook_a, dd, exact, ctx->d_hook_b, d_mem, per);
    std::vector<int32_t> cnt(n_src), mem((size_t)n_src * per);
    cudaError_t e1 = cudaMemcpyAsync(cnt.data(), ctx->d_hook_b, sizeof(int32_t) * n_src, cudaMemcpyDeviceToHost, st);
    cudaError_t e2 = cudaMemcpyAsync(mem.data(), d_mem, sizeof(int32_t) * mem.size(), cudaMemcpyDeviceToHost, st);
    cudaError_t e3 = cudaStreamSynchronize(st);
    cudaFree(d_mem);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) return fail(ctx, POC_ERR_CUDA, "poc_rings: CUDA failure");
    int64_t p = 0;
    for (int s = 0; s < n_src; s++) {
        out_off[s] = (int32_t)p;
        if (p + cnt[s] > cap_mem) return fail(ctx, POC_ERR_CAPACITY, "poc_rings: out_mem too small");
        std::sort(mem.begin() + (size_t)s * per, mem.begin() + (size_t)s * per + cnt[s]);
        if (out_mem) std::copy(mem.begin() + (size_t)s * per, mem.begin() + (size_t)s * per + cnt[s], out_mem + p);
        p += cnt[s];
    }
    out_off[n_src] = (int32_t)p;
    return POC_OK;
}

static int pairs_hook(poc_ctx *ctx, int replica, int n_pairs, const int32_t *a, const int32_t *b, double *out, int mode) {
    int n = ctx->h_nagents[replica];
    for (int i = 0; i < n_pairs; i++) if (a[i] < 0 || a[i] >= n || b[i] < 0 || b[i] >= n) return fail(ctx, POC_ERR_ARG, "bad slot in pair list");
    int rc = hook_bufs(ctx, std::max(n_pairs, 1));
    if (rc) return rc;
    cudaStream_t st = ctx->stream;
    CU(cudaMemcpyAsync(ctx->d_hook_a, a, sizeof(int32_t) * n_pairs, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->d_hook_b, b, sizeof(int32_t) * n_pairs, cudaMemcpyHostToDevice, st));
    if (n_pairs) k_pairs_hook<<<(n_pairs * 32 + 255) / 256, 256, 0, st>>>(ctx->d, replica, n_pairs, ctx->d_hook_a, ctx->d_hook_b, ctx->d_hook_out, mode);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, ctx->d_hook_out, sizeof(double) * n_pairs, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return POC_OK;
}

int poc_social_proximity(poc_ctx *ctx, int replica, int n_pairs, const int32_t *a, const int32_t *b, double *out) {
    CHECK_CTX();
    if (replica < 0 || replica >= ctx->d.R || n_pairs < 0 || !a || !b || !out) return fail(ctx, POC_ERR_ARG, "poc_social_proximity: bad argument");
    return pairs_hook(ctx, replica, n_pairs, a, b, out, 0);
}

static int run_emb(poc_ctx *ctx, int replica, int n, const int32_t *agents) {
    DevCtx &d = ctx->d;
    if (n > d.Ncap) return fail(ctx, POC_ERR_CAPACITY, "too many embeddedness requests");
    cudaStream_t st = ctx->stream;
    std::vector<int32_t> zero(2 * d.R, 0);
    zero[replica] = n;
    CU(cudaMemcpyAsync(d.n_emb, zero.data(), sizeof(int32_t) * zero.size(), cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(d.n_emb_big, 0, sizeof(int32_t) * d.R, st));
    if (n) CU(cudaMemcpyAsync(d.emb_list + (size_t)replica * d.Ncap, agents, sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
    { Timed t(ctx, TM_EMB); launch_emb(ctx, ctx->d, ctx->stream, 0); }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));
    return POC_OK;
}

int poc_embeddedness(poc_ctx *ctx, int replica, int n, const int32_t *agents, double *out) {
    CHECK_CTX();
    if (replica < 0 || replica >= ctx->d.R || n < 0 || !agents || !out) return fail(ctx, POC_ERR_ARG, "poc_embeddedness: bad argument");
    int na = ctx->h_nagents[replica];
    for (int i = 0; i < n; i++) if (agents[i] < 0 || agents[i] >= na) return fail(ctx, POC_ERR_ARG, "poc_embeddedness: bad slot");
    int rc = run_emb(ctx, replica, n, agents);
    if (rc) return rc;
    std::vector<double> all(na);
    CU(cudaMemcpyAsync(all.data(), ctx->d.emb_val + (size_t)replica * ctx->d.Ncap, sizeof(double) * na, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < n; i++) out[i] = all[agents[i]];
    return POC_OK;
}

int poc_embeddedness_accumulated(poc_ctx *ctx, int replica, int n, const int32_t *agents, int n_over, const int32_t *over, double *out) {
    CHECK_CTX();
    DevCtx &d = ctx->d;
    if (replica < 0 || replica >= d.R || n < 0 || n > d.Ncap || (n && (!agents || !out)) || n_over < 0 || (n_over && !over))
        return fail(ctx, POC_ERR_ARG, "poc_embeddedness_accumulated: bad argument");
    int na = ctx->h_nagents[replica];
    for (int i = 0; i < n; i++) if (agents[i] < 0 || agents[i] >= na) return fail(ctx, POC_ERR_ARG, "poc_embeddedness_accumulated: bad slot");
    if (!n) return POC_OK;
    // rows normalised to {lo, hi, evaluation, multiplicity}, sorted by (lo, hi, evaluation)
    std::vector<int4> ov(n_over);
    for (int i = 0; i < n_over; i++) {
        int j = over[4 * i], a = over[4 * i + 1], b = over[4 * i + 2], m = over[4 * i + 3];
        if (j < 0 || j >= n || a < 0 || a >= na || b < 0 || b >= na || m < 1) return fail(ctx, POC_ERR_ARG, "poc_embeddedness_accumulated: bad override row");
        ov[i] = make_int4(std::min(a, b), std::max(a, b), j, m);
    }
    std::sort(ov.begin(), ov.end(), [](const int4 &x, const int4 &y) { return x.x != y.x ? x.x < y.x : (x.y != y.y ? x.y < y.y : x.z < y.z); });
    int EW = (n + 31) / 32;
    int32_t *d_slots = nullptr; unsigned *d_bits = nullptr; int *d_has = nullptr; int4 *d_over = nullptr; double *d_out = nullptr;
    cudaStream_t st = ctx->stream;
    cudaError_t e = cudaMalloc((void **)&d_slots, sizeof(int32_t) * n);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_bits, sizeof(unsigned) * (size_t)std::max(na, 1) * EW);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_has, sizeof(int) * n);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_over, sizeof(int4) * std::max(n_over, 1));
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_out, sizeof(double) * n);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_bits, 0, sizeof(unsigned) * (size_t)std::max(na, 1) * EW, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_slots, agents, sizeof(int32_t) * n, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && n_over) e = cudaMemcpyAsync(d_over, ov.data(), sizeof(int4) * n_over, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        int grid = std::min(n, d.scr_slots);
        k_emb_acc_mark<<<grid, 256, ctx->smem_search, st>>>(d, replica, n, d_slots, d_bits, EW, d_has);
        k_emb_acc_eval<<<grid, 256, ctx->smem_search, st>>>(d, replica, n, d_slots, d_bits, EW, d_has, n_over, d_over, d_out, ctx->epoch + 1);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, sizeof(double) * n, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_slots); cudaFree(d_bits); cudaFree(d_has); cudaFree(d_over); cudaFree(d_out);
    if (e != cudaSuccess) return fail(ctx, POC_ERR_CUDA, std::string("poc_embeddedness_accumulated: ") + cudaGetErrorString(e));
    return POC_OK;
}

int poc_preload_embeddedness(poc_ctx *ctx, int replica, int n, const int32_t *agents, const double *values) {
    CHECK_CTX();
    DevCtx &d = ctx->d;
    if (replica < 0 || replica >= d.R || n < 0 || (n && (!agents || !values))) return fail(ctx, POC_ERR_ARG, "poc_preload_embeddedness: bad argument");
    int na = ctx->h_nagents[replica];
    for (int i = 0; i < n; i++) if (agents[i] < 0 || agents[i] >= na) return fail(ctx, POC_ERR_ARG, "poc_preload_embeddedness: bad slot");
    if (!n) return POC_OK;
    int32_t *d_slots = nullptr; double *d_vals = nullptr;
    cudaStream_t st = ctx->stream;
    cudaError_t e = cudaMalloc((void **)&d_slots, sizeof(int32_t) * n);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_vals, sizeof(double) * n);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_slots, agents, sizeof(int32_t) * n, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_vals, values, sizeof(double) * n, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) { k_emb_preload<<<(n + 255) / 256, 256, 0, st>>>(d, replica, n, d_slots, d_vals, ctx->epoch + 1); e = cudaGetLastError(); }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_slots); cudaFree(d_vals);
    if (e != cudaSuccess) return fail(ctx, POC_ERR_CUDA, std::string("poc_preload_embeddedness: ") + cudaGetErrorString(e));
    return POC_OK;
}

int poc_candidate_weights(poc_ctx *ctx, int replica, int n_pairs, const int32_t *self, const int32_t *cand, double *out) {
    CHECK_CTX();
    if (replica < 0 || replica >= ctx->d.R || n_pairs < 0 || !self || !cand || !out) return fail(ctx, POC_ERR_ARG, "poc_candidate_weights: bad argument");
    int na = ctx->h_nagents[replica];
    for (int i = 0; i < n_pairs; i++) if (cand[i] < 0 || cand[i] >= na) return fail(ctx, POC_ERR_ARG, "poc_candidate_weights: bad slot");
    std::vector<int32_t> uniq(cand, cand + n_pairs);
    std::sort(uniq.begin(), uniq.end());
    uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
    int rc = run_emb(ctx, replica, (int)uniq.size(), uniq.data());
    if (rc) return rc;
    return pairs_hook(ctx, replica, n_pairs, self, cand, out, 1);
}

int poc_find_accomplices(poc_ctx *ctx, int replica, int n, const int32_t *initiator, const int32_t *k, const int32_t *fac_pick,
                         const double *u_fac, int32_t *groups, int32_t *group_len, int32_t *d_fails) {
    CHECK_CTX();
    DevCtx &d = ctx->d;
    if (replica < 0 || replica >= d.R || n < 0 || n > d.Ccap || !initiator || !k || !groups || !group_len) return fail(ctx, POC_ERR_ARG, "poc_find_accomplices: bad argument");
    int na = ctx->h_nagents[replica];
    for (int i = 0; i < n; i++) if (initiator[i] < 0 || initiator[i] >= na || k[i] < 0) return fail(ctx, POC_ERR_ARG, "poc_find_accomplices: bad initiator / k");
    int rc = hook_bufs(ctx, std::max(n, 1));
    if (rc) return rc;
    cudaStream_t st = ctx->stream;
    std::vector<int32_t> hdr((size_t)d.R * 8, 0);
    hdr[(size_t)replica * 8] = 2;  // hook mode: rp_u_init carries u_fac, rp_fac_pick the recorded pick
    std::vector<double> uf(n, 0.0);
    std::vector<int32_t> fp(n, -1);
    if (u_fac) std::copy(u_fac, u_fac + n, uf.begin());
    if (fac_pick) std::copy(fac_pick, fac_pick + n, fp.begin());
    CU(cudaMemcpyAsync(d.rp_hdr, hdr.data(), sizeof(int32_t) * hdr.size(), cudaMemcpyHostToDevice, st));
    if (n) {
        CU(cudaMemcpyAsync(d.rp_u_init + (size_t)replica * d.Ccap, uf.data(), sizeof(double) * n, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d.rp_fac_pick + (size_t)replica * d.Ccap, fp.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(ctx->d_hook_a, initiator, sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(ctx->d_hook_b, k, sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
    }
    // other replicas must stay idle during the hook
    std::vector<int32_t> zeros(d.R, 0);
    CU(cudaMemcpyAsync(d.n_search, zeros.data(), sizeof(int32_t) * d.R, cudaMemcpyHostToDevice, st));
    std::vector<int> cn_before((size_t)d.R * CN_COUNT);
    CU(cudaMemcpyAsync(cn_before.data(), d.counters, sizeof(int) * cn_before.size(), cudaMemcpyDeviceToHost, st));
    ctx->epoch++;
    rc = set_tick(ctx, 0, 0, ctx->epoch, 0, st);
    if (rc) return rc;
    k_hook_init_crimes<<<(std::max(n, 1) + 255) / 256, 256, 0, st>>>(d, replica, n, ctx->d_hook_a, ctx->d_hook_b);
    k_hook_fill_search<<<1, 32, 0, st>>>(d, replica, n);
    CU(cudaGetLastError());
    rc = launch_search_rounds(ctx, ctx->d, ctx->stream);
    if (rc) return rc;
    std::vector<int32_t> acc((size_t)n * POC_MAX_GROUP), nacc(n), fails(n);
    if (n) {
        CU(cudaMemcpyAsync(acc.data(), d.cr_acc + (size_t)replica * d.Ccap * POC_MAX_GROUP, sizeof(int32_t) * acc.size(), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(nacc.data(), d.cr_nacc + (size_t)replica * d.Ccap, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(fails.data(), d.cr_fails + (size_t)replica * d.Ccap, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
    }
    CU(cudaStreamSynchronize(st));
    // the hook must not disturb the model counters
    CU(cudaMemcpyAsync(d.counters, cn_before.data(), sizeof(int) * cn_before.size(), cudaMemcpyHostToDevice, st));
    std::fill(hdr.begin(), hdr.end(), 0);
    CU(cudaMemcpyAsync(d.rp_hdr, hdr.data(), sizeof(int32_t) * hdr.size(), cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));
    for (int i = 0; i < n; i++) {
        int g = k[i] > 0 ? nacc[i] : 0;
        for (int j = 0; j < g; j++) groups[(size_t)i * POC_MAX_GROUP + j] = acc[(size_t)i * POC_MAX_GROUP + j];
        groups[(size_t)i * POC_MAX_GROUP + g] = initiator[i];
        group_len[i] = g + 1;
        if (d_fails) { d_fails[3 * i] = fails[i] & 1; d_fails[3 * i + 1] = (fails[i] >> 1) & 1; d_fails[3 * i + 2] = (fails[i] >> 2) & 1; }
    }
    return POC_OK;
}

}  // extern "C"
