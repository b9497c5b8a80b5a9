// orx_api.cu -- the C ABI of include/orx.h: context, tables, launches.  Plain CUDA runtime, no torch.
//
// There is no CPU fallback anywhere in this file: every entry point either runs its CUDA kernels or
// returns ORX_E_CUDA / ORX_E_INVALID.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <string>
#include <vector>

#define ORX_BLOCK_THREADS 128
#ifndef ORX_MIN_BLOCKS
#define ORX_MIN_BLOCKS 4   /* bound seen by the front end; ptxas gets the tight one (build.py) */
#endif

#include "../../include/orx.h"
#include "orx_game.cuh"
#include "orx_lanes.cuh"
#include "orx_mapimage.h"

using namespace orx;

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int set_err(int code, const char *fmt, const char *a = "", const char *b = "")
{
    snprintf(g_err, sizeof g_err, fmt, a, b);
    return code;
}
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return set_err(ORX_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

extern "C" const char *orx_last_error(void) { return g_err; }
// for the other translation units of the library (orx_tree.cpp)
__attribute__((visibility("hidden"))) void orx_internal_set_error(const char *text) { snprintf(g_err, sizeof g_err, "%s", text ? text : ""); }

// ---------------------------------------------------------------------------------------------
// BattleSim outcome tables (O/BattleSim.java:190-286), built on the device at context creation.
// One block per (a, d).  generateProbabilities' scatter sweep is restated as a gather: a cell's
// contributions are added in the reference's source order (remA ascending, remD ascending), each
// as an unfused float multiply followed by an unfused float add, so every float equals Java's.
// ---------------------------------------------------------------------------------------------
__constant__ float c_single[3][2][3] = {
    { { 0.4167f, 0.5833f, 0.f }, { 0.2546f, 0.7454f, 0.f } },          // 1 v 1, 1 v 2   (O/BattleSim.java:15-19)
    { { 0.5787f, 0.4213f, 0.f }, { 0.2276f, 0.3241f, 0.4483f } },     // 2 v 1, 2 v 2   (:21-26)
    { { 0.6597f, 0.3403f, 0.f }, { 0.3717f, 0.3358f, 0.2926f } },     // 3 v 1, 3 v 2   (:28-33)
};

__device__ __forceinline__ float table_contrib(const float *last, int A, int D, int sa, int sd, int ta, int td)
{
    if (sa < 1 || sd < 1 || sa > A || sd > D) return 0.f;
    int pa = sa < 3 ? sa : 3, pd = sd < 2 ? sd : 2, n = pa < pd ? pa : pd;
    int x = sa - ta, y = sd - td;
    if (x < 0 || y < 0 || x + y != n) return 0.f;
    float v = last[sa * (D + 1) + sd];
    if (v == 0.f) return 0.f;
    return __fmul_rn(v, c_single[pa - 1][pd - 1][x]);
}

__global__ void __launch_bounds__(256) build_tables_kernel(uint32_t *meta, uint64_t *rows, float *probs)
{
    extern __shared__ float tsm[];
    const int A = blockIdx.x / ORX_TABLE_LIMIT + 1, D = blockIdx.x % ORX_TABLE_LIMIT + 1;
    const int RL = D + 1, cells = (A + 1) * RL;
    float *last = tsm, *next = tsm + cells;
    for (int c = threadIdx.x; c < cells; c += blockDim.x) last[c] = 0.f;
    __syncthreads();
    if (threadIdx.x == 0) last[A * RL + D] = 1.0f;
    __syncthreads();
    for (;;) {
        int transient = 0;
        for (int c = threadIdx.x; c < cells; c += blockDim.x) {
            int ta = c / RL, td = c % RL;
            float acc = 0.f;
            if (ta == 0 || td == 0) acc = last[c];  // an absorbing cell keeps its mass (added first: it is its own first source)
            else if (last[c] != 0.f) transient = 1;
            // sources in (a ascending, d ascending) order
            acc = __fadd_rn(acc, table_contrib(last, A, D, ta, td + 1, ta, td));
            acc = __fadd_rn(acc, table_contrib(last, A, D, ta, td + 2, ta, td));
            acc = __fadd_rn(acc, table_contrib(last, A, D, ta + 1, td, ta, td));
            acc = __fadd_rn(acc, table_contrib(last, A, D, ta + 1, td + 1, ta, td));
            acc = __fadd_rn(acc, table_contrib(last, A, D, ta + 2, td, ta, td));
            next[c] = acc;
        }
        int any = __syncthreads_or(transient);
        float *t = last; last = next; next = t;
        if (!any) break;
    }
    // candidates in (remA asc, remD asc) order: (0, q) for q <= D, then (q - D, 0)
    __shared__ float s_p[2 * ORX_TABLE_LIMIT + 2];
    __shared__ uint8_t s_code[2 * ORX_TABLE_LIMIT + 2];
    __shared__ int s_n;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    const int ncand = A + D + 1;
    float *cand_p = next;  // scratch (the other sweep buffer)
    for (int q = threadIdx.x; q < ncand; q += blockDim.x) cand_p[q] = q <= D ? last[0 * RL + q] : last[(q - D) * RL + 0];
    __syncthreads();
    for (int q = threadIdx.x; q < ncand; q += blockDim.x) {
        float p = cand_p[q];
        if (p != 0.f) {
            int rank = 0;  // Collections.sort(reverseOrder()): descending probability, stable
            for (int j = 0; j < ncand; j++) {
                float pj = cand_p[j];
                if (pj != 0.f && (pj > p || (pj == p && j < q))) rank++;
            }
            s_p[rank] = p;
            s_code[rank] = q <= D ? (uint8_t)q : (uint8_t)(0x80 | (q - D));  // survivors; bit 7 = invading (remD == 0)
            atomicAdd(&s_n, 1);
        }
    }
    __syncthreads();
    // row base: tables are laid out with capacity a + d each, (a, d) row-major
    if (threadIdx.x == 0) {
        uint32_t off = 0;
        for (int a = 1; a < A; a++) off += (uint32_t)(ORX_TABLE_LIMIT * a + ORX_TABLE_LIMIT * (ORX_TABLE_LIMIT + 1) / 2);
        for (int d = 1; d < D; d++) off += (uint32_t)(A + d);
        int n = s_n;
        double total = 0;
        for (int i = 0; i < n; i++) {
            total = __dadd_rn(total, (double)s_p[i]);                // "total += ao.probability" (O/BattleSim.java:132)
            double scaled = __dmul_rn(total, 9007199254740992.0);    // exact: power of two
            unsigned long long thr = (unsigned long long)scaled;     // r <= total  <=>  k <= floor(total * 2^53)
            if (thr > 9007199254740991ull) thr = 9007199254740991ull;
            rows[off + i] = (thr << 11) | (unsigned long long)s_code[i];
            probs[off + i] = s_p[i];
        }
        meta[(A - 1) * ORX_TABLE_LIMIT + (D - 1)] = (off << 8) | (uint32_t)n;
    }
}

static const int kTableRows = 2 * ORX_TABLE_LIMIT * (ORX_TABLE_LIMIT * (ORX_TABLE_LIMIT + 1) / 2);  // sum of a + d

// ---------------------------------------------------------------------------------------------
// small kernels behind the BattleSim / stream / card entry points
// ---------------------------------------------------------------------------------------------
__global__ void battles_kernel(const __grid_constant__ DevMap dm, int a, int d, int n, unsigned long long seed,
                               unsigned long long first_game, int individual, int32_t *al_out, int32_t *dl_out)
{
    __shared__ __align__(16) uint32_t sm[4][128 + 16];   // 16 header words, then the 128-word block
    int wid = threadIdx.x >> 5;
    long long i = (long long)blockIdx.x * 4 + wid;
    if (i >= n) return;
    Stream<false> s;
    s.init(to_saddr(sm[wid] + 16), to_saddr(sm[wid]), seed, first_game + (unsigned long long)i, nullptr, 0);
    int al = 0, dl = 0;
    uint32_t rk[20];
#pragma unroll
    for (int r = 0; r < 10; r++) { rk[2 * r] = (uint32_t)seed + (uint32_t)r * 0x9E3779B9u; rk[2 * r + 1] = (uint32_t)(seed >> 32) + (uint32_t)r * 0xBB67AE85u; }
    if (individual) {
        if (a < 1 || a > 3 || d < 1 || d > 2) s.flags |= ORX_FLAG_ERROR;
        else al = single_roll(dm, a, d, s.next_k53());
    } else if (a >= 1 && d >= 1) {
        simulate_battle(dm, s, rk, a, d, al, dl);
    }
    if (lane_id() == 0) { al_out[i] = al; if (dl_out) dl_out[i] = dl; }
}

template <bool REPLAY>
__global__ void draws_kernel(unsigned long long seed, unsigned long long game, const uint32_t *words, uint32_t n_words,
                             const int32_t *ops, int n_ops, double *out, uint32_t *pos_flags)
{
    __shared__ __align__(16) uint32_t sm[128 + 16];
    Stream<REPLAY> s;
    s.init(to_saddr(sm + 16), to_saddr(sm), seed, game, words, n_words);
    for (int i = 0; i < n_ops; i++) {
        int op = ops[i];
        double v;
        if (op > 0) v = (double)s.next_int(op);
        else if (op == 0) v = s.next_double();
        else if (op == -1) v = next_gaussian(s);
        else v = (double)s.word();
        if (lane_id() == 0) out[i] = v;
    }
    if (lane_id() == 0) { pos_flags[0] = s.pos(); pos_flags[1] = s.flags; }
}

__global__ void card_value_kernel(const __grid_constant__ DevMap dm, int pos, int32_t *out)
{
    uint32_t flags = 0;
    int v = card_cash_value(dm, pos, flags);
    out[0] = v; out[1] = (int32_t)flags;
}

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
struct OrxCtx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    DevMap dm;
    int nwt = 2;
    int sm_count = 0;
    int blocks_per_sm[4] = { 0, 0, 0, 0 };  // [replay + 2 * modelling]
    bool modelling = false;           // orx_set_sim_agents: the MODEL kernels run, the slice carries moveable armies
    // the lane-per-game kernel (orx_lanes.cuh): all-SimRandom, counter streams, <= 64 territories, large batches
    int lane_blocks_per_sm = 0;
    size_t lane_smem_bytes = 0;
    unsigned long long lane_min_total = 0;   // batches at least this large go to the lane kernel
    int kernel_choice = 0;                   // ORX_KERNEL: 0 auto, 1 always warp-per-game, 2 lane-per-game whenever eligible
    LaneTune lane_tune;
    int last_kernel = 0;                     // 1 warp-per-game, 2 lane-per-game (orx_last_kernel_kind)
    int blocks_cap = 0;               // ORX_BLOCKS_PER_SM tuning knob
    size_t smem_bytes = 0;
    uint8_t *d_blob = nullptr;
    int32_t *d_bonus_table = nullptr, *d_card_table = nullptr;
    uint32_t *d_meta = nullptr;
    uint64_t *d_rows = nullptr;
    float *d_probs = nullptr;
    unsigned long long *d_ticket = nullptr;
    // staging for the host-buffer entry points (grown on demand)
    void *d_leaves = nullptr; size_t cap_leaves = 0;
    void *d_agg = nullptr; size_t cap_agg = 0;
    void *d_games = nullptr; size_t cap_games = 0;
    void *d_words = nullptr; size_t cap_words = 0;
    bool have_time = false;
    uint64_t launches = 0;
    // two independent in-flight batches for orx_rollout_begin / orx_rollout_end
    struct AsyncSlot {
        cudaStream_t stream = nullptr;
        unsigned long long *d_ticket = nullptr;
        void *d_leaves = nullptr; size_t cap_leaves = 0;
        void *d_agg = nullptr; size_t cap_agg = 0;
        void *d_games = nullptr; size_t cap_games = 0;
        int n_leaves = 0, ppl = 0;
        bool want_games = false, busy = false;
    } aslot[2];
};

static int ensure(void **p, size_t *cap, size_t need)
{
    if (need <= *cap) return ORX_OK;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    CK(cudaMalloc(p, need));
    *cap = need;
    return ORX_OK;
}

typedef void (*RolloutKernel)(const DevMap, const RolloutArgs);

// the eight instantiations: territories <= 64 or more, counter or replay stream, all-SimRandom or modelled opponents
static RolloutKernel kernel_for(int nwt, bool replay, bool model)
{
    if (nwt == 2) {
        if (model) return replay ? rollout_kernel<2, true, true> : rollout_kernel<2, false, true>;
        return replay ? rollout_kernel<2, true, false> : rollout_kernel<2, false, false>;
    }
    if (model) return replay ? rollout_kernel<8, true, true> : rollout_kernel<8, false, true>;
    return replay ? rollout_kernel<8, true, false> : rollout_kernel<8, false, false>;
}

static cudaError_t configure_kernel(RolloutKernel k, size_t smem, int *blocks_per_sm)
{
    cudaError_t e = cudaFuncSetAttribute((const void *)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, (const void *)k, ORX_BLOCK_THREADS, smem);
}

// per-warp slice of dynamic shared memory; the moveable-armies array only exists while opponent modelling is on
static void layout_slice(OrxCtx *c, bool with_moveable)
{
    DevMap &dm = c->dm;
    int off = ORX_WARP_HDR_BYTES + ORX_WARP_ACC_BYTES;   // header words, then the warp's running OrxLeafResult sums
    dm.w_rng = off;    off += 512;
    dm.w_armies = off; off += 4 * dm.NP;
    dm.w_owner = off;  off += dm.NP;
    dm.w_alist = off;  off += dm.cap;
    dm.w_cards = off;  off += (dm.P + 1) * dm.cap;
    off = (off + 3) & ~3;
    dm.w_move = off;   if (with_moveable) off += 4 * dm.NP;
    dm.w_bytes = (off + 15) & ~15;
    c->smem_bytes = (size_t)dm.blob_bytes + (size_t)(ORX_BLOCK_THREADS / 32) * (size_t)dm.w_bytes;
}

// occupancy of the two kernels (counter / replay) this context launches with its current slice size
static int configure_context(OrxCtx *c)
{
    for (int replay = 0; replay < 2; replay++) {
        int idx = replay + 2 * (int)c->modelling;
        CK(configure_kernel(kernel_for(c->nwt, replay != 0, c->modelling), c->smem_bytes, &c->blocks_per_sm[idx]));
        if (c->blocks_cap >= 1 && c->blocks_per_sm[idx] > c->blocks_cap) c->blocks_per_sm[idx] = c->blocks_cap;
        if (c->blocks_per_sm[idx] < 1) return set_err(ORX_E_CUDA, "rollout kernel does not fit on an SM");
    }
    return ORX_OK;
}

// the lane-per-game kernel: shared memory per block, residency, scheduler tuning, and the batch size from which it
// is used (below that the warp-per-game kernel fills the machine better: a batch needs 32 games per warp here)
static int configure_lanes(OrxCtx *c)
{
    c->lane_blocks_per_sm = 0;
    if (c->dm.N > LN_MAXN) return ORX_OK;
    const LaneLayout LL = lane_layout(c->dm.N, c->dm.P);
    c->lane_smem_bytes = (size_t)c->dm.blob_bytes + (size_t)(ORX_LANE_THREADS / 32) * (size_t)LL.words * 128u;
    CK(cudaFuncSetAttribute((const void *)rollout_lanes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->lane_smem_bytes));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->lane_blocks_per_sm, (const void *)rollout_lanes_kernel, ORX_LANE_THREADS, c->lane_smem_bytes));
    if (const char *lim = getenv("ORX_LANE_BLOCKS_PER_SM")) { int v = atoi(lim); if (v >= 1 && v < c->lane_blocks_per_sm) c->lane_blocks_per_sm = v; }
    LaneTune &t = c->lane_tune;
    memset(&t, 0, sizeof t);
    t.thr_new = 4; t.thr_turn = 8; t.thr_sel = 8; t.thr_fin = 8;
    t.age_new = 64; t.age_turn = 32; t.age_sel = 8; t.age_fin = 8;
    t.coop_min = 48;
    if (const char *tv = getenv("ORX_LANE_TUNE")) {   // "thr_new,thr_turn,thr_sel,thr_fin,age_new,age_turn,age_sel,age_fin,coop_min"
        int v[9], n = sscanf(tv, "%d,%d,%d,%d,%d,%d,%d,%d,%d", v, v + 1, v + 2, v + 3, v + 4, v + 5, v + 6, v + 7, v + 8);
        int32_t *dst[9] = { &t.thr_new, &t.thr_turn, &t.thr_sel, &t.thr_fin, &t.age_new, &t.age_turn, &t.age_sel, &t.age_fin, &t.coop_min };
        for (int i = 0; i < n && i < 9; i++) *dst[i] = v[i];
    }
    // The lane kernel is opt-in: measured on B200 it executes 30 % fewer warp instructions per playout than the
    // warp-per-game kernel but keeps the issue slots only 28 % busy (instruction-fetch stalls, 14 of 32 lanes active,
    // 16 warps per SM), so it is the slower of the two (profiles/r2d_*).  ORX_KERNEL=lanes forces it;
    // ORX_LANE_MIN_TOTAL=n uses it for batches of at least n playouts.
    c->lane_min_total = ~0ull;
    if (const char *k = getenv("ORX_KERNEL")) c->kernel_choice = !strcmp(k, "warp") ? 1 : (!strcmp(k, "lanes") ? 2 : 0);
    if (const char *mt = getenv("ORX_LANE_MIN_TOTAL")) c->lane_min_total = strtoull(mt, nullptr, 10);
    return ORX_OK;
}

static bool use_lane_kernel(const OrxCtx *c, bool replay, const RolloutArgs &ar)
{
    if (replay || c->modelling || c->lane_blocks_per_sm < 1 || c->kernel_choice == 1) return false;
    if (ar.total >= (1ull << 32)) return false;
    return c->kernel_choice == 2 || ar.total >= c->lane_min_total;
}

extern "C" int orx_create(const OrxMap *map, const OrxRules *rules, int device, OrxCtx **out)
{
    if (!map || !rules || !out) return set_err(ORX_E_INVALID, "null argument");
    *out = nullptr;
    MapImage im;
    if (const char *why = build_map_image(map, rules, im)) return set_err(ORX_E_INVALID, "%s", why);

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return set_err(ORX_E_CUDA, "no CUDA device: this engine has no CPU path");
    if (device < 0 || device >= ndev) return set_err(ORX_E_INVALID, "device index out of range");
    CK(cudaSetDevice(device));

    OrxCtx *c = new OrxCtx();
    c->device = device;
    c->dm = im.dm;
    c->nwt = im.nwt;
    DevMap &dm = c->dm;
    const std::vector<uint8_t> &blob = im.blob;
    layout_slice(c, false);
    dm.model_limit = 0; dm.agents = 0u;

#define CKC(call)                                                                                         \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) { set_err(ORX_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); orx_destroy(c); return ORX_E_CUDA; } \
    } while (0)

    CKC(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CKC(cudaEventCreate(&c->ev0));
    CKC(cudaEventCreate(&c->ev1));
    cudaDeviceProp prop;
    CKC(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;

    {   // floor(2^32 / b) + 1 (Stream::next_int)
        std::vector<uint32_t> magic(ORX_MAGIC_N, 0u);
        for (uint32_t b = 2; b < ORX_MAGIC_N; b++) magic[b] = (uint32_t)((1ull << 32) / b) + 1u;
        CKC(cudaMemcpyToSymbol(c_magic, magic.data(), sizeof(uint32_t) * ORX_MAGIC_N));
    }
    CKC(cudaMalloc(&c->d_blob, blob.size()));
    CKC(cudaMemcpyAsync(c->d_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice, c->stream));
    dm.blob = c->d_blob;
    if (!im.bonus_table.empty()) {
        CKC(cudaMalloc(&c->d_bonus_table, im.bonus_table.size() * 4));
        CKC(cudaMemcpy(c->d_bonus_table, im.bonus_table.data(), im.bonus_table.size() * 4, cudaMemcpyHostToDevice));
        dm.bonus_table = c->d_bonus_table;
    }
    if (!im.card_table.empty()) {
        CKC(cudaMalloc(&c->d_card_table, im.card_table.size() * 4));
        CKC(cudaMemcpy(c->d_card_table, im.card_table.data(), im.card_table.size() * 4, cudaMemcpyHostToDevice));
        dm.card_table = c->d_card_table;
    }
    CKC(cudaMalloc(&c->d_meta, sizeof(uint32_t) * ORX_TABLE_LIMIT * ORX_TABLE_LIMIT));
    CKC(cudaMalloc(&c->d_rows, sizeof(uint64_t) * kTableRows));
    CKC(cudaMalloc(&c->d_probs, sizeof(float) * kTableRows));
    CKC(cudaMalloc(&c->d_ticket, sizeof(unsigned long long)));
    dm.bt_meta = c->d_meta; dm.bt_rows = c->d_rows;
    {
        size_t tsm = 2 * sizeof(float) * (ORX_TABLE_LIMIT + 1) * (ORX_TABLE_LIMIT + 1);
        CKC(cudaFuncSetAttribute(build_tables_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm));
        CKC(cudaMemsetAsync(c->d_rows, 0, sizeof(uint64_t) * kTableRows, c->stream));   // unused rows of a block stay zero
        build_tables_kernel<<<ORX_TABLE_LIMIT * ORX_TABLE_LIMIT, 256, tsm, c->stream>>>(c->d_meta, c->d_rows, c->d_probs);
        CKC(cudaGetLastError());
        c->launches++;
    }
    if (const char *lim = getenv("ORX_BLOCKS_PER_SM")) c->blocks_cap = atoi(lim);  // tuning knob: cap the resident blocks per SM
    if (configure_context(c) != ORX_OK) { orx_destroy(c); return ORX_E_CUDA; }
    if (configure_lanes(c) != ORX_OK) { orx_destroy(c); return ORX_E_CUDA; }
    CKC(cudaStreamSynchronize(c->stream));
#undef CKC
    *out = c;
    return ORX_OK;
}

extern "C" void orx_destroy(OrxCtx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    cudaFree(c->d_blob); cudaFree(c->d_bonus_table); cudaFree(c->d_card_table); cudaFree(c->d_meta);
    cudaFree(c->d_rows); cudaFree(c->d_probs); cudaFree(c->d_ticket);
    cudaFree(c->d_leaves); cudaFree(c->d_agg); cudaFree(c->d_games); cudaFree(c->d_words);
    for (auto &a : c->aslot) {
        if (a.stream) { cudaStreamSynchronize(a.stream); cudaStreamDestroy(a.stream); }
        cudaFree(a.d_ticket); cudaFree(a.d_leaves); cudaFree(a.d_agg); cudaFree(a.d_games);
    }
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

// new ModellingBoard(bs, sas, cm, cid, modellingTurnLimit) (ModellingBoard.java:13-19)
extern "C" int orx_set_sim_agents(OrxCtx *c, const int32_t *agent_ids, int modelling_turn_limit)
{
    if (!c) return set_err(ORX_E_INVALID, "null context");
    CK(cudaSetDevice(c->device));
    for (int k = 0; k < 2; k++) if (c->aslot[k].busy) return set_err(ORX_E_INVALID, "a rollout of this context is in flight");
    CK(cudaStreamSynchronize(c->stream));
    uint32_t packed = 0;
    bool on = agent_ids && modelling_turn_limit > 0;
    if (on) {
        bool any = false;
        for (int p = 0; p < c->dm.P; p++) {
            if (agent_ids[p] < 0 || agent_ids[p] >= ORX_AGENT_COUNT) return set_err(ORX_E_INVALID, "unknown agent id (see ORX_AGENT_* in orx_state.h)");
            packed |= (uint32_t)agent_ids[p] << (3 * p);
            any = any || agent_ids[p] != ORX_AGENT_RANDOM;
        }
        on = any;  // an all-SimRandom line-up behind ModellingBoard is the plain board (ModellingBoard.java:29-47)
    }
    c->modelling = on;
    c->dm.agents = on ? packed : 0u;
    c->dm.model_limit = on ? modelling_turn_limit : 0;
    layout_slice(c, on);
    return configure_context(c);
}

extern "C" int orx_last_kernel_kind(const OrxCtx *c) { return c ? c->last_kernel : 0; }

extern "C" int orx_kernel_info(const OrxCtx *c, uint64_t total_playouts, int32_t *kind, int32_t *blocks_per_sm,
                               int32_t *threads_per_block, int32_t *smem_bytes_per_block)
{
    if (!c) return set_err(ORX_E_INVALID, "null context");
    RolloutArgs ar;
    memset(&ar, 0, sizeof ar);
    ar.total = total_playouts;
    const bool lanes = use_lane_kernel(c, false, ar);
    if (kind) *kind = lanes ? 2 : 1;
    if (blocks_per_sm) *blocks_per_sm = lanes ? c->lane_blocks_per_sm : c->blocks_per_sm[2 * (int)c->modelling];
    if (threads_per_block) *threads_per_block = lanes ? ORX_LANE_THREADS : ORX_BLOCK_THREADS;
    if (smem_bytes_per_block) *smem_bytes_per_block = (int32_t)(lanes ? c->lane_smem_bytes : c->smem_bytes);
    return ORX_OK;
}
extern "C" int orx_leaf_stride(const OrxCtx *c) { return c ? c->dm.leaf_stride : ORX_E_INVALID; }
extern "C" void *orx_stream(OrxCtx *c) { return c ? (void *)c->stream : nullptr; }
extern "C" uint64_t orx_kernel_launches(const OrxCtx *c) { return c ? c->launches : 0; }

extern "C" int orx_sync(OrxCtx *c)
{
    if (!c) return set_err(ORX_E_INVALID, "null context");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    return ORX_OK;
}

extern "C" float orx_last_kernel_ms(OrxCtx *c)
{
    if (!c || !c->have_time) return -1.f;
    cudaSetDevice(c->device);
    if (cudaEventSynchronize(c->ev1) != cudaSuccess) return -1.f;
    float ms = -1.f;
    if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) != cudaSuccess) return -1.f;
    return ms;
}

// The opt-in dynamic shared-memory size is an attribute of the kernel FUNCTION, not of a context: two contexts with
// different maps (different sizes) driven from different host threads must not interleave "set attribute" and "launch".
static std::mutex g_launch_mutex;

// one launch of the rollout kernel on `stream` (default: the context's stream, bracketed by the timing events)
static int launch_rollout(OrxCtx *c, bool replay, const RolloutArgs &ar, cudaStream_t stream = nullptr)
{
    const bool timed = stream == nullptr;
    if (!stream) stream = c->stream;
    CK(cudaMemsetAsync(ar.ticket, 0, sizeof(unsigned long long), stream));
    std::lock_guard<std::mutex> launch_lock(g_launch_mutex);
    if (use_lane_kernel(c, replay, ar)) {
        LaneArgs la;
        memset(&la, 0, sizeof la);
        la.leaves = ar.leaves; la.n_leaves = ar.n_leaves; la.playouts_per_leaf = ar.playouts_per_leaf; la.total = ar.total;
        la.seed = ar.seed; la.first_game = ar.first_game; la.agg = ar.agg; la.games = ar.games; la.ticket = ar.ticket;
        for (int r = 0; r < 10; r++) {
            la.rk[2 * r] = (uint32_t)ar.seed + (uint32_t)r * 0x9E3779B9u;
            la.rk[2 * r + 1] = (uint32_t)(ar.seed >> 32) + (uint32_t)r * 0xBB67AE85u;
        }
        la.tune = c->lane_tune;
        unsigned long long blocks = (unsigned long long)c->sm_count * (unsigned long long)c->lane_blocks_per_sm;
        const unsigned long long need = (ar.total + ORX_LANE_THREADS - 1) / ORX_LANE_THREADS;
        if (need < blocks) blocks = need;
        if (blocks == 0) blocks = 1;
        {
            cudaError_t e = cudaFuncSetAttribute((const void *)rollout_lanes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->lane_smem_bytes);
            if (e != cudaSuccess) return set_err(ORX_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        }
        if (timed) CK(cudaEventRecord(c->ev0, stream));
        rollout_lanes_kernel<<<dim3((unsigned)blocks), dim3(ORX_LANE_THREADS), c->lane_smem_bytes, stream>>>(c->dm, la);
        CK(cudaGetLastError());
        if (timed) { CK(cudaEventRecord(c->ev1, stream)); c->have_time = true; }
        c->launches++;
        c->last_kernel = 2;
        return ORX_OK;
    }
    c->last_kernel = 1;
    RolloutArgs wa = ar;
    for (int r = 0; r < 10; r++) {
        wa.rk[2 * r] = (uint32_t)ar.seed + (uint32_t)r * 0x9E3779B9u;
        wa.rk[2 * r + 1] = (uint32_t)(ar.seed >> 32) + (uint32_t)r * 0xBB67AE85u;
    }
    const RolloutKernel kern = kernel_for(c->nwt, replay, c->modelling);
    unsigned long long warps = (unsigned long long)c->sm_count * c->blocks_per_sm[(int)replay + 2 * (int)c->modelling] * (ORX_BLOCK_THREADS / 32);
    unsigned long long need = (ar.total + (ORX_BLOCK_THREADS / 32) - 1) / (ORX_BLOCK_THREADS / 32);
    unsigned long long blocks = warps / (ORX_BLOCK_THREADS / 32);
    if (need < blocks) blocks = need;
    if (blocks == 0) blocks = 1;
    // the opt-in shared-memory size is per function, not per context: re-assert it for this context's map
    {
        cudaError_t e = cudaFuncSetAttribute((const void *)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_bytes);
        if (e != cudaSuccess) return set_err(ORX_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    if (timed) CK(cudaEventRecord(c->ev0, stream));
    dim3 grid((unsigned)blocks), block(ORX_BLOCK_THREADS);
    kern<<<grid, block, c->smem_bytes, stream>>>(c->dm, wa);
    CK(cudaGetLastError());
    if (timed) { CK(cudaEventRecord(c->ev1, stream)); c->have_time = true; }
    c->launches++;
    return ORX_OK;
}

extern "C" int orx_rollout_batch_device(OrxCtx *c, const void *leaves_dev, int n_leaves, int playouts_per_leaf,
                                        uint64_t seed, uint64_t first_game, OrxLeafResult *agg_dev, OrxGameResult *games_dev)
{
    if (!c || !leaves_dev || n_leaves < 0 || playouts_per_leaf < 0) return set_err(ORX_E_INVALID, "bad argument");
    if (!agg_dev && !games_dev) return set_err(ORX_E_INVALID, "no output buffer");
    CK(cudaSetDevice(c->device));
    if (agg_dev) CK(cudaMemsetAsync(agg_dev, 0, sizeof(OrxLeafResult) * (size_t)n_leaves, c->stream));
    RolloutArgs ar;
    memset(&ar, 0, sizeof ar);
    ar.leaves = (const uint8_t *)leaves_dev; ar.n_leaves = n_leaves; ar.playouts_per_leaf = playouts_per_leaf;
    ar.total = (unsigned long long)n_leaves * (unsigned long long)playouts_per_leaf;
    ar.seed = seed; ar.first_game = first_game; ar.agg = agg_dev; ar.games = games_dev; ar.ticket = c->d_ticket;
    if (ar.total == 0) return ORX_OK;
    return launch_rollout(c, false, ar);
}

extern "C" int orx_rollout_batch(OrxCtx *c, const void *leaves, int n_leaves, int playouts_per_leaf, uint64_t seed,
                                 uint64_t first_game, OrxLeafResult *agg, OrxGameResult *games)
{
    if (!c || !leaves || n_leaves < 0 || playouts_per_leaf < 0) return set_err(ORX_E_INVALID, "bad argument");
    if (!agg && !games) return set_err(ORX_E_INVALID, "no output buffer");
    CK(cudaSetDevice(c->device));
    size_t total = (size_t)n_leaves * (size_t)playouts_per_leaf;
    size_t lbytes = (size_t)n_leaves * (size_t)c->dm.leaf_stride;
    int rc;
    if ((rc = ensure(&c->d_leaves, &c->cap_leaves, lbytes ? lbytes : 16)) != ORX_OK) return rc;
    if ((rc = ensure(&c->d_agg, &c->cap_agg, sizeof(OrxLeafResult) * (size_t)(n_leaves ? n_leaves : 1))) != ORX_OK) return rc;
    if (games && (rc = ensure(&c->d_games, &c->cap_games, sizeof(OrxGameResult) * (total ? total : 1))) != ORX_OK) return rc;
    CK(cudaMemcpyAsync(c->d_leaves, leaves, lbytes, cudaMemcpyHostToDevice, c->stream));
    rc = orx_rollout_batch_device(c, c->d_leaves, n_leaves, playouts_per_leaf, seed, first_game,
                                  (OrxLeafResult *)c->d_agg, games ? (OrxGameResult *)c->d_games : nullptr);
    if (rc != ORX_OK) return rc;
    if (agg) CK(cudaMemcpyAsync(agg, c->d_agg, sizeof(OrxLeafResult) * (size_t)n_leaves, cudaMemcpyDeviceToHost, c->stream));
    if (games) CK(cudaMemcpyAsync(games, c->d_games, sizeof(OrxGameResult) * total, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return ORX_OK;
}

extern "C" int orx_rollout_begin(OrxCtx *c, int slot, const void *leaves, int n_leaves, int playouts_per_leaf, uint64_t seed,
                                 uint64_t first_game, int want_games)
{
    if (!c || slot < 0 || slot > 1 || !leaves || n_leaves < 1 || playouts_per_leaf < 1) return set_err(ORX_E_INVALID, "bad argument");
    OrxCtx::AsyncSlot &a = c->aslot[slot];
    if (a.busy) return set_err(ORX_E_INVALID, "slot still in flight: call orx_rollout_end first");
    CK(cudaSetDevice(c->device));
    if (!a.stream) {
        CK(cudaStreamCreateWithFlags(&a.stream, cudaStreamNonBlocking));
        CK(cudaMalloc(&a.d_ticket, sizeof(unsigned long long)));
    }
    size_t total = (size_t)n_leaves * (size_t)playouts_per_leaf, lbytes = (size_t)n_leaves * (size_t)c->dm.leaf_stride;
    int rc;
    if ((rc = ensure(&a.d_leaves, &a.cap_leaves, lbytes)) != ORX_OK) return rc;
    if ((rc = ensure(&a.d_agg, &a.cap_agg, sizeof(OrxLeafResult) * (size_t)n_leaves)) != ORX_OK) return rc;
    if (want_games && (rc = ensure(&a.d_games, &a.cap_games, sizeof(OrxGameResult) * total)) != ORX_OK) return rc;
    CK(cudaMemcpyAsync(a.d_leaves, leaves, lbytes, cudaMemcpyHostToDevice, a.stream));
    CK(cudaMemsetAsync(a.d_agg, 0, sizeof(OrxLeafResult) * (size_t)n_leaves, a.stream));
    RolloutArgs ar;
    memset(&ar, 0, sizeof ar);
    ar.leaves = (const uint8_t *)a.d_leaves; ar.n_leaves = n_leaves; ar.playouts_per_leaf = playouts_per_leaf; ar.total = total;
    ar.seed = seed; ar.first_game = first_game; ar.agg = (OrxLeafResult *)a.d_agg;
    ar.games = want_games ? (OrxGameResult *)a.d_games : nullptr; ar.ticket = a.d_ticket;
    if ((rc = launch_rollout(c, false, ar, a.stream)) != ORX_OK) return rc;
    a.n_leaves = n_leaves; a.ppl = playouts_per_leaf; a.want_games = want_games != 0; a.busy = true;
    return ORX_OK;
}

extern "C" int orx_rollout_end(OrxCtx *c, int slot, OrxLeafResult *agg, OrxGameResult *games)
{
    if (!c || slot < 0 || slot > 1) return set_err(ORX_E_INVALID, "bad argument");
    OrxCtx::AsyncSlot &a = c->aslot[slot];
    if (!a.busy) return set_err(ORX_E_INVALID, "slot is not in flight");
    CK(cudaSetDevice(c->device));
    if (agg) CK(cudaMemcpyAsync(agg, a.d_agg, sizeof(OrxLeafResult) * (size_t)a.n_leaves, cudaMemcpyDeviceToHost, a.stream));
    if (games && a.want_games)
        CK(cudaMemcpyAsync(games, a.d_games, sizeof(OrxGameResult) * (size_t)a.n_leaves * (size_t)a.ppl, cudaMemcpyDeviceToHost, a.stream));
    cudaError_t e = cudaStreamSynchronize(a.stream);
    a.busy = false;
    if (e != cudaSuccess) return set_err(ORX_E_CUDA, "cudaStreamSynchronize: %s", cudaGetErrorString(e));
    return ORX_OK;
}

extern "C" int orx_rollout_replay(OrxCtx *c, const void *leaves, int n_games, const uint32_t *words, uint32_t words_per_game,
                                  OrxGameResult *games)
{
    if (!c || !leaves || !games || n_games < 0 || (!words && words_per_game)) return set_err(ORX_E_INVALID, "bad argument");
    CK(cudaSetDevice(c->device));
    size_t lbytes = (size_t)n_games * (size_t)c->dm.leaf_stride, wbytes = (size_t)n_games * words_per_game * 4;
    int rc;
    if ((rc = ensure(&c->d_leaves, &c->cap_leaves, lbytes ? lbytes : 16)) != ORX_OK) return rc;
    if ((rc = ensure(&c->d_words, &c->cap_words, wbytes ? wbytes : 16)) != ORX_OK) return rc;
    if ((rc = ensure(&c->d_games, &c->cap_games, sizeof(OrxGameResult) * (size_t)(n_games ? n_games : 1))) != ORX_OK) return rc;
    if (n_games == 0) return ORX_OK;
    CK(cudaMemcpyAsync(c->d_leaves, leaves, lbytes, cudaMemcpyHostToDevice, c->stream));
    if (wbytes) CK(cudaMemcpyAsync(c->d_words, words, wbytes, cudaMemcpyHostToDevice, c->stream));
    RolloutArgs ar;
    memset(&ar, 0, sizeof ar);
    ar.leaves = (const uint8_t *)c->d_leaves; ar.n_leaves = n_games; ar.playouts_per_leaf = 1; ar.total = (unsigned long long)n_games;
    ar.games = (OrxGameResult *)c->d_games; ar.ticket = c->d_ticket;
    ar.words = (const uint32_t *)c->d_words; ar.words_per_game = words_per_game;
    if ((rc = launch_rollout(c, true, ar)) != ORX_OK) return rc;
    CK(cudaMemcpyAsync(games, c->d_games, sizeof(OrxGameResult) * (size_t)n_games, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return ORX_OK;
}

extern "C" int orx_get_attack_outcomes(OrxCtx *c, int a, int d, float *probability, int32_t *attacker_losses,
                                       int32_t *defender_losses, int32_t *invading, int cap)
{
    if (!c || a < 1 || d < 1 || a > ORX_TABLE_LIMIT || d > ORX_TABLE_LIMIT || cap < 0) return set_err(ORX_E_INVALID, "bad argument");
    CK(cudaSetDevice(c->device));
    uint32_t meta;
    CK(cudaMemcpy(&meta, c->d_meta + (a - 1) * ORX_TABLE_LIMIT + (d - 1), 4, cudaMemcpyDeviceToHost));
    int n = (int)(meta & 0xFFu);
    std::vector<uint64_t> rows((size_t)(n ? n : 1));
    std::vector<float> pr((size_t)(n ? n : 1));
    if (n) {
        CK(cudaMemcpy(rows.data(), c->d_rows + (meta >> 8), sizeof(uint64_t) * (size_t)n, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(pr.data(), c->d_probs + (meta >> 8), sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost));
    }
    for (int i = 0; i < n && i < cap; i++) {
        int code = (int)(rows[i] & 0xFFu), rem = code & 0x7F, inv = code >> 7;
        if (probability) probability[i] = pr[i];
        if (attacker_losses) attacker_losses[i] = inv ? a - rem : a;
        if (defender_losses) defender_losses[i] = inv ? d : d - rem;
        if (invading) invading[i] = inv;
    }
    return n;
}

static int run_battles(OrxCtx *c, int a, int d, int n, uint64_t seed, uint64_t first_game, int individual, int32_t *al, int32_t *dl)
{
    if (!c || n < 0 || !al) return set_err(ORX_E_INVALID, "bad argument");
    if (n == 0) return ORX_OK;
    CK(cudaSetDevice(c->device));
    int rc;
    if ((rc = ensure(&c->d_games, &c->cap_games, sizeof(int32_t) * 2 * (size_t)n)) != ORX_OK) return rc;
    int32_t *d_al = (int32_t *)c->d_games, *d_dl = d_al + n;
    battles_kernel<<<(n + 3) / 4, 128, 0, c->stream>>>(c->dm, a, d, n, seed, first_game, individual, d_al, dl ? d_dl : nullptr);
    CK(cudaGetLastError());
    c->launches++;
    CK(cudaMemcpyAsync(al, d_al, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    if (dl) CK(cudaMemcpyAsync(dl, d_dl, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return ORX_OK;
}

extern "C" int orx_simulate_battles(OrxCtx *c, int a, int d, int n, uint64_t seed, uint64_t first_game, int32_t *al, int32_t *dl)
{
    if (!dl) return set_err(ORX_E_INVALID, "bad argument");
    return run_battles(c, a, d, n, seed, first_game, 0, al, dl);
}

extern "C" int orx_simulate_individual(OrxCtx *c, int a, int d, int n, uint64_t seed, uint64_t first_game, int32_t *al)
{
    if (a < 1 || a > 3 || d < 1 || d > 2) return set_err(ORX_E_INVALID, "simulateIndividualBattle needs 1..3 attackers and 1..2 defenders");
    return run_battles(c, a, d, n, seed, first_game, 1, al, nullptr);
}

extern "C" int orx_stream_draws(OrxCtx *c, uint64_t seed, uint64_t game, const uint32_t *words, uint32_t n_words,
                                const int32_t *ops, int n_ops, double *out, uint32_t *pos_out, uint32_t *flags_out)
{
    if (!c || !ops || !out || n_ops < 0) return set_err(ORX_E_INVALID, "bad argument");
    CK(cudaSetDevice(c->device));
    size_t need = 8 * (size_t)n_ops + 4 * (size_t)n_ops + 4 * (size_t)n_words + 64;
    int rc;
    if ((rc = ensure(&c->d_words, &c->cap_words, need)) != ORX_OK) return rc;
    uint8_t *base = (uint8_t *)c->d_words;
    double *d_out = (double *)base;
    uint32_t *d_pf = (uint32_t *)(base + 8 * (size_t)n_ops);
    int32_t *d_ops = (int32_t *)(base + 8 * (size_t)n_ops + 16);
    uint32_t *d_w = (uint32_t *)(base + 8 * (size_t)n_ops + 16 + 4 * (size_t)n_ops);
    CK(cudaMemcpyAsync(d_ops, ops, 4 * (size_t)n_ops, cudaMemcpyHostToDevice, c->stream));
    if (words) {
        if (n_words) CK(cudaMemcpyAsync(d_w, words, 4 * (size_t)n_words, cudaMemcpyHostToDevice, c->stream));
        draws_kernel<true><<<1, 32, 0, c->stream>>>(seed, game, d_w, n_words, d_ops, n_ops, d_out, d_pf);
    } else {
        draws_kernel<false><<<1, 32, 0, c->stream>>>(seed, game, nullptr, 0, d_ops, n_ops, d_out, d_pf);
    }
    CK(cudaGetLastError());
    c->launches++;
    uint32_t pf[2];
    CK(cudaMemcpyAsync(out, d_out, 8 * (size_t)n_ops, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(pf, d_pf, 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (pos_out) *pos_out = pf[0];
    if (flags_out) *flags_out = pf[1];
    return ORX_OK;
}

extern "C" int orx_card_cash_value(OrxCtx *c, int pos, int32_t *value)
{
    if (!c || !value) return set_err(ORX_E_INVALID, "bad argument");
    CK(cudaSetDevice(c->device));
    int rc;
    if ((rc = ensure(&c->d_agg, &c->cap_agg, 64)) != ORX_OK) return rc;
    card_value_kernel<<<1, 1, 0, c->stream>>>(c->dm, pos, (int32_t *)c->d_agg);
    CK(cudaGetLastError());
    c->launches++;
    int32_t v[2];
    CK(cudaMemcpyAsync(v, c->d_agg, 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (v[1]) return set_err(ORX_E_INVALID, "progression position outside the device table");
    *value = v[0];
    return ORX_OK;
}
