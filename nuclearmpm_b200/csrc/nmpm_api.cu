// C-ABI implementation (include/nmpm.h): owns the device particle store, the grid, the stream and
// the per-step launch sequence.  No CPU fallback: without a usable CUDA device every compute entry
// point returns NMPM_ERR_NO_DEVICE / NMPM_ERR_CUDA.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <array>
#include <unordered_map>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/nmpm.h"
#include "nmpm_kernels.cuh"
#include "nmpm_p2g_cell.cuh"
#include "nmpm_fused.cuh"
#include "nmpm_sort.cuh"

using namespace nmpm;

#define NMPM_STR2(x) #x
#define NMPM_STR(x) NMPM_STR2(x)

namespace {
thread_local std::string g_create_error;

__global__ void __launch_bounds__(256) k_add_planes(float4* __restrict__ dst, const float4* __restrict__ src, size_t count) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    float4 a = dst[i];
    const float4 b = src[i];
    a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
    dst[i] = a;
}
}  // namespace

// Ring of node boxes.  Three are needed (previous / current / next); four keeps the ring index in phase with the
// sort cadence (1, 2, 4, 8), so that the CUDA-graph cache of step_once sees 2*sort_every host states instead of 6x.
constexpr int kBoxRing = 4;

struct nmpm_sim {
    int dim = 0, model = 0, res = 0;
    size_t n = 0, cells = 0;  // n = live particles (slots [0,n) of store[cur] after a G2P)
    size_t cap = 0;           // allocated particle slots
    size_t n_store = 0;       // slab mode: entries in store[cur] incl. migrated-away ("gone") and just-received ones
    size_t n_gone = 0;        // slab mode: entries of store[cur] whose key is kKeyGone
    bool slab = false;
    MaterialParams P{};
    nmpm_options opt{};
    float E = 0, nu = 0, gravity = 0;
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;

    ParticleStore store[2]{};
    int cur = 0;
    float4* grid = nullptr;
    // Fused G2P+P2G (nmpm_fused.cuh; 3D, single GPU): the G2P of step n scatters step n+1 into `grid_alt`; the next step's
    // P2G phase swaps the two buffers.  Invariant: outside [fused launch, next P2G phase] `grid_alt` is all zero.
    float4* grid_alt = nullptr;
    int fuse = 0;            // 0 off, 1 on (not on the first step after an upload: that state may be replaced again), 2 always
    bool p2g_ahead = false;  // grid_alt holds the P2G of store[cur] (the coming step) over box[box_cur]
    int grid_sel = 0;        // which of the two buffers `grid` is (CUDA-graph key)
    int fused_minb = NMPM_FUSED_MINB;  // CTAs per SM the fused kernel is compiled for (env NMPM_FUSED_MINB: experiments)
    // Batch of independent 2D scenes stacked along x (nmpm_create_batch*): MaterialParams::scenes / scene_of / lame
    int scenes = 1;
    unsigned short* d_scene_of = nullptr;
    float2* d_lame = nullptr;
    std::vector<float2> lame_host;
    // Active node tiles (nmpm_kernels.cuh: k_tile_decide, k_mark_tiles, k_tiles3; 3D, single GPU): one flag bit per
    // 4^3-node tile, raised from the cell keys of the coming step, consumed by grid_op and the clear.  A ring like the
    // node boxes and indexed like them: tile_ring[b] / d_tile_want[b] belong to the positions box[b] bounds.  The mode is
    // decided on the device per step: d_tile_want[b] says whether the flags of slot b were raised; every tile / box kernel
    // checks it and returns at once if it is not its turn.
    uint32_t* tile_ring[kBoxRing] = {nullptr, nullptr, nullptr, nullptr};
    int* d_tile_want = nullptr;  // kBoxRing ints
    bool tiles = false;          // the arrays above exist (policy is not "never")
    int tiles_u = 4;             // flagged tiles in flight per warp of k_tiles3 (env NMPM_TILES_U: experiments)
    int tile_policy = 0;         // 0 adaptive (k_tile_decide), 1 never, 2 always — nmpm_options.tiles after the grid-size rule
    // node boxes (GridBox, device): box[box_cur] bounds the particles of the current step, box[(box_cur+3)%4]
    // the nodes the previous P2G wrote (cleared at the start of the next one), box[(box_cur+1)%4] is being
    // built by the G2P in flight
    GridBox* d_box = nullptr;
    int* d_box_partial = nullptr;  // one partial box (8 ints) per G2P warp
    int box_cur = 0;
    std::vector<std::pair<int, int>> dirty_planes;  // slab mode: node planes written by nmpm_grid_add_planes
    std::vector<std::array<int, 5>> dirty_rects;    // slab mode: in-plane rectangles written by the native ghost exchange
    struct nmpm_slab_comm* sc = nullptr;            // native slab step (nmpm_slab_comm.inl)
    bool box_valid = false;   // box[box_cur] describes store[cur]
    bool local_reorder = true;  // in-place G2P re-groups each warp's 32 slots by cell key (NMPM_LOCAL_REORDER=0: off)
    CUtensorMap grid_map{};     // 3D: the grid as a rank-4 tensor {4 floats, z, y, x} with one x-plane window box (G2P)
    bool g2p_window = false;    // 3D: G2P stages its node window through the TMA (NMPM_G2P_WINDOW=0: off)
    bool g2p_pipe = false;      // 3D: ... in the persistent software-pipelined kernel (k_g2p_pipe)
    // device-driven slab step (nmpm_slab_comm.inl): the true slot / gone counts live in d_ctr, `n_store` is only an upper
    // bound for launch sizes, every slot beyond the true count (and every migrated-away slot) carries kKeyGone in keys_a
    bool dev_counts = false;
    int* d_ctr = nullptr;
    bool grid_valid = false;  // false until the first p2g: the reference's grid() is empty (src/solver.cpp:52-57)

    SortWorkspace sort;
    int tiles_per_axis = 0, key_bits = 0;
    bool keys_valid = false;        // sort.keys_a holds the cell keys of store[cur] (written by the last G2P)
    const uint32_t* perm = nullptr; // this step's sorted permutation (slot of the i-th sorted particle), or null

    int* d_error = nullptr;
    int* h_error = nullptr;  // pinned
    bool error_latched = false;

    float* staging = nullptr;  // device scratch for import/export
    size_t staging_bytes = 0;
    // pipelined host I/O (nmpm_upload_particles_async / nmpm_download_particles_async): two copy streams and double-
    // buffered device staging, so that the H2D copy of step k+1 and the D2H copy of step k-1 overlap step k
    struct AsyncIo {
        cudaStream_t s_in = nullptr, s_out = nullptr;
        float* in[2] = {nullptr, nullptr};
        float* out[2] = {nullptr, nullptr};
        size_t bytes = 0;
        cudaEvent_t in_ready[2] = {}, in_free[2] = {}, out_ready[2] = {}, out_free[2] = {};
        unsigned k_in = 0, k_out = 0;
        bool ok = false;
    } io;

    long long steps_done = 0;
    long long launches = 0;
    int phase_next = 0;

    // CUDA-graph cache: one captured step per host-side step state (store parity, sort phase, keys_valid)
    struct StepGraph {
        cudaGraphExec_t exec = nullptr;
        int cur_after = 0;
        int box_after = 0;
        bool keys_valid_after = false;
        bool ahead_after = false;
        int grid_sel_after = 0;
        int launches = 0;
    };
    std::unordered_map<int, StepGraph> graphs;
    bool graphs_ok = true;

    bool timing = false;
    cudaEvent_t ev[6]{};
    float t_ms[NMPM_T_COUNT]{};
    int t_steps = 0;

    std::string last_error;
};

static void slab_comm_free(nmpm_sim* h);

#define CUDA_TRY(h, expr)                                                                              \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess) {                                                                       \
            char _buf[512];                                                                            \
            snprintf(_buf, sizeof(_buf), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                     __LINE__);                                                                        \
            if (h) (h)->last_error = _buf;                                                             \
            else                                                                                       \
                g_create_error = _buf;                                                                 \
            return NMPM_ERR_CUDA;                                                                      \
        }                                                                                              \
    } while (0)

#define NMPM_DISPATCH_DIM(h, CALL)  \
    do {                            \
        if ((h)->dim == 2) {        \
            constexpr int D = 2;    \
            CALL;                   \
        } else {                    \
            constexpr int D = 3;    \
            CALL;                   \
        }                           \
    } while (0)

#define NMPM_DISPATCH(h, CALL)                  \
    do {                                        \
        if ((h)->dim == 2) {                    \
            constexpr int D = 2;                \
            if ((h)->model == 0) {              \
                constexpr int MODEL = 0;        \
                CALL;                           \
            } else if ((h)->model == 1) {       \
                constexpr int MODEL = 1;        \
                CALL;                           \
            } else {                            \
                constexpr int MODEL = 2;        \
                CALL;                           \
            }                                   \
        } else {                                \
            constexpr int D = 3;                \
            if ((h)->model == 0) {              \
                constexpr int MODEL = 0;        \
                CALL;                           \
            } else if ((h)->model == 1) {       \
                constexpr int MODEL = 1;        \
                CALL;                           \
            } else {                            \
                constexpr int MODEL = 2;        \
                CALL;                           \
            }                                   \
        }                                       \
    } while (0)

static inline unsigned blocks_for(size_t n, int threads) { return (unsigned) ((n + threads - 1) / threads); }

static int ensure_staging(nmpm_sim* h, size_t bytes) {
    if (h->staging_bytes >= bytes) return NMPM_OK;
    if (h->staging) CUDA_TRY(h, cudaFree(h->staging));
    h->staging = nullptr;
    h->staging_bytes = 0;
    CUDA_TRY(h, cudaMalloc(&h->staging, bytes));
    h->staging_bytes = bytes;
    return NMPM_OK;
}

static int alloc_store(nmpm_sim* h, ParticleStore& S) {
    const size_t n = h->cap ? h->cap : 1;
    const int nq = (h->dim == 3) ? 6 : 3;
    for (int k = 0; k < nq; ++k) CUDA_TRY(h, cudaMalloc(&S.q[k], n * sizeof(float4)));
    CUDA_TRY(h, cudaMalloc(&S.s, n * sizeof(float)));
    CUDA_TRY(h, cudaMalloc(&S.mv, n * sizeof(float2)));
    CUDA_TRY(h, cudaMalloc(&S.id, n * sizeof(uint32_t)));
    return NMPM_OK;
}

static void free_store(ParticleStore& S) {
    for (int k = 0; k < 6; ++k)
        if (S.q[k]) cudaFree(S.q[k]);
    if (S.s) cudaFree(S.s);
    if (S.mv) cudaFree(S.mv);
    if (S.id) cudaFree(S.id);
    S = ParticleStore{};
}

// constants exactly as the reference ctor computes them (src/nclr.h:74-78, :285, :325)
static void fill_params(nmpm_sim* h, float dt, float E, float nu, float gravity) {
    MaterialParams& P = h->P;
    P.res = h->res;
    P.n1 = h->res + 1;
    P.dt = dt;
    P.dx = (float) (1.0 / h->res);
    P.inv_dx = 1 / P.dx;
    P.mu_0 = E / (2 * (1 + nu));
    P.lambda_0 = E * nu / ((1 + nu) * (1 - 2 * nu));
    P.Dinv = 4 * P.inv_dx * P.inv_dx;
    P.vmax = (float) ((double) P.dx * 0.9 / (double) dt);
    P.dt_gravity = dt * gravity;
}

// The dense grid as a TMA tensor: rank 4 {4 floats of a node, z, y, x}, box = one x-plane of the G2P node window
// (nmpm_kernels.cuh: kWinY x kWinZ nodes).  Out-of-range box parts (beyond node res) are filled with zeros.
static int make_grid_map(nmpm_sim* h) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
        qres != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return NMPM_ERR_CUDA;
    }
    const cuuint64_t n1 = (cuuint64_t) h->res + 1;
    const cuuint64_t dims[4] = {4, n1, n1, n1};
    const cuuint64_t strides[3] = {16, 16 * n1, 16 * n1 * n1};  // bytes, dims 1..3
    const cuuint32_t box[4] = {4, (cuuint32_t) kWinZ, (cuuint32_t) kWinY, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = reinterpret_cast<EncodeFn>(fn)(&h->grid_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, h->grid, dims, strides, box,
                                                      estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? NMPM_OK : NMPM_ERR_CUDA;
}

// Batch of independent 2D scenes (nmpm_create_batch / nmpm_create_batch_aos)
struct BatchSpec {
    int nscenes;
    const size_t* counts;  // particles per scene; the particle arrays are the scenes' concatenated
    const float* E;
    const float* nu;
};

static void lame_of(float E, float nu, float* mu_0, float* lambda_0) {  // src/nclr.h:76-77, in float like the reference
    *mu_0 = E / (2 * (1 + nu));
    *lambda_0 = E * nu / ((1 + nu) * (1 - 2 * nu));
}

static int create_common(int dim, int model, int res, float dt, float E, float nu, float gravity, size_t n,
                         const nmpm_options* opt, nmpm_sim** out, const BatchSpec* batch = nullptr) {
    if (!out) return NMPM_ERR_INVALID;
    *out = nullptr;
    if ((dim != 2 && dim != 3) || model < 0 || model > 2 || res < 4 || n > 0xFFFFFFF0ull) {
        g_create_error = "invalid argument (dim must be 2 or 3, model 0..2, res >= 4)";
        return NMPM_ERR_INVALID;
    }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        g_create_error = "no CUDA device: libnmpm has no CPU fallback";
        return NMPM_ERR_NO_DEVICE;
    }
    nmpm_sim* h = new (std::nothrow) nmpm_sim();
    if (!h) return NMPM_ERR_INVALID;
    nmpm_default_options(&h->opt);
    if (opt) h->opt = *opt;
    if (h->opt.device < 0 || h->opt.device >= count) {
        g_create_error = "invalid device ordinal";
        delete h;
        return NMPM_ERR_INVALID;
    }
    h->dim = dim, h->model = model, h->res = res, h->n = n;
    if (batch) {
        size_t total = 0;
        for (int k = 0; k < batch->nscenes; ++k) total += batch->counts[k];
        if (dim != 2 || batch->nscenes < 1 || batch->nscenes > 65535 || total != n || h->opt.slab_x1 > 0 ||
            h->opt.sort_every < 1 || h->opt.p2g_variant == 1) {
            g_create_error = "batch: 2D scenes only (1..65535 of them), counts must add up to n, not a slab, with binning";
            delete h;
            return NMPM_ERR_INVALID;
        }
        h->scenes = batch->nscenes;
    }
    if (const char* lr = std::getenv("NMPM_LOCAL_REORDER")) h->local_reorder = (*lr != '0');
    h->n_store = n;
    h->slab = h->opt.slab_x1 > 0;
    h->cap = n;
    if (h->slab) {
        if (h->opt.slab_x0 < 0 || h->opt.slab_x1 <= h->opt.slab_x0 || h->opt.capacity < 0) {
            g_create_error = "invalid slab range / capacity";
            delete h;
            return NMPM_ERR_INVALID;
        }
        if ((size_t) h->opt.capacity > n) h->cap = (size_t) h->opt.capacity;
        h->opt.use_graph = 0;   // particle counts change every step
        if (h->opt.sort_every < 1) h->opt.sort_every = 1;  // migrants are compacted away by the sorts
    }
    h->E = E, h->nu = nu, h->gravity = gravity;
    h->device = h->opt.device;
    fill_params(h, dt, E, nu, gravity);
    const size_t n1 = (size_t) res + 1;
    h->cells = (dim == 3) ? n1 * n1 * n1 : n1 * n1 * (size_t) h->scenes;  // batch: the scenes' grids stacked along x
    if (h->cells >= 0x7FFFFFFFull) {  // node indices travel as 32-bit signed ints in the P2G packets
        g_create_error = "grid too large: (res+1)^dim must be below 2^31";
        delete h;
        return NMPM_ERR_INVALID;
    }
    h->tiles_per_axis = (int) ((n1 + (1 << kTileBits) - 1) >> kTileBits);
    {
        size_t tiles = 1;
        for (int d = 0; d < dim; ++d) tiles *= (size_t) h->tiles_per_axis;
        if (h->scenes > 1)  // node rows along x: scenes * n1
            tiles = ((n1 * (size_t) h->scenes + (1 << kTileBits) - 1) >> kTileBits) * (size_t) h->tiles_per_axis;
        int bits = 0;
        while ((1ull << bits) < tiles) ++bits;
        h->key_bits = bits + dim * kTileBits;
        if (h->key_bits > 32) {
            g_create_error = "grid too large for 32-bit cell keys";
            delete h;
            return NMPM_ERR_INVALID;
        }
    }
    *out = h;
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->own_stream = true;
    if (int rc = alloc_store(h, h->store[0])) return rc;
    if (int rc = alloc_store(h, h->store[1])) return rc;
    CUDA_TRY(h, cudaMalloc(&h->grid, h->cells * sizeof(float4)));
    CUDA_TRY(h, cudaMemset(h->grid, 0, h->cells * sizeof(float4)));  // the only dense clear; afterwards box by box
    if (dim == 3) {
        // auto = global gather: measured on cfg4, the staged window is 5 % slower (profiles/r02c)
        int mode = h->opt.g2p_window;  // 0 auto, 1 global gather, 2 one-shot window, 3 pipelined window
        if (const char* gw = std::getenv("NMPM_G2P_WINDOW")) mode = (*gw >= '0' && *gw <= '3') ? (*gw - '0') + 1 : mode;
        if (mode == 0) mode = 1;
        h->g2p_window = mode >= 2 && make_grid_map(h) == NMPM_OK;  // no tensor map: plain global gather
        h->g2p_pipe = h->g2p_window && mode == 3;
    }
    {   // fused G2P+P2G: 3D, single GPU, binned, global gather
        int mode = h->opt.fuse;  // 0 auto, 1 off, 2 on, 3 always
        if (const char* fz = std::getenv("NMPM_FUSE")) mode = (*fz >= '0' && *fz <= '2') ? (*fz - '0') + 1 : mode;
        if (mode == 0) mode = 2;
        h->fuse = (dim == 3 && !h->slab && h->opt.sort_every > 0 && !h->g2p_window && mode >= 2) ? mode - 1 : 0;
        if (const char* mb = std::getenv("NMPM_FUSED_MINB")) h->fused_minb = std::atoi(mb);
        if (h->fuse) {
            CUDA_TRY(h, cudaMalloc(&h->grid_alt, h->cells * sizeof(float4)));
            CUDA_TRY(h, cudaMemset(h->grid_alt, 0, h->cells * sizeof(float4)));
        }
    }
    if (dim == 3 && !h->slab) {
        h->tile_policy = h->opt.tiles;
        if (const char* tl = std::getenv("NMPM_TILES")) h->tile_policy = (*tl == '0') ? 1 : (*tl == '2') ? 2 : (*tl == '3') ? 3 : 0;
        // adaptive: only where the dense grid is large enough for its bounding box to cost anything (cfg2 / snow128 at
        // res 128 clear + update their whole 2 M-node grid in 0.03 ms; the five extra near-empty launches per step of the
        // tile machinery cost a launch-bound 262 k-particle scene 8 %)
        if (h->tile_policy == 0 && h->cells < ((size_t) 8 << 20)) h->tile_policy = 1;
        if (h->tile_policy == 3) h->tile_policy = 0;  // adaptive whatever the grid size (tests)
        if (h->tile_policy != 1) {
            const size_t T = (n1 + 3) / 4, bytes = ((T * T * T + 31) / 32 + 16) * sizeof(uint32_t);
            for (int k = 0; k < kBoxRing; ++k) {
                CUDA_TRY(h, cudaMalloc(&h->tile_ring[k], bytes));
                CUDA_TRY(h, cudaMemset(h->tile_ring[k], 0, bytes));
            }
            h->tiles = true;
            if (const char* tu = std::getenv("NMPM_TILES_U")) h->tiles_u = std::atoi(tu);
            CUDA_TRY(h, cudaMalloc(&h->d_tile_want, kBoxRing * sizeof(int)));
            CUDA_TRY(h, cudaMemset(h->d_tile_want, 0, kBoxRing * sizeof(int)));
        }
    }
    CUDA_TRY(h, cudaMalloc(&h->d_box, kBoxRing * sizeof(GridBox)));
    CUDA_TRY(h, cudaMalloc(&h->d_box_partial, ((h->cap + 127) / 128 * 4 + 4) * 8 * sizeof(int)));
    for (int k = 0; k < kBoxRing; ++k) k_box_reset<<<1, 32, 0, h->stream>>>(h->d_box + k);
    CUDA_TRY(h, cudaMalloc(&h->d_error, 2 * sizeof(int)));  // [0] this step, [1] found ahead by the fused scatter
    CUDA_TRY(h, cudaMemset(h->d_error, 0, 2 * sizeof(int)));
    CUDA_TRY(h, cudaMallocHost(&h->h_error, sizeof(int)));
    *h->h_error = 0;
    // sort workspace
    const size_t nn = h->cap ? h->cap : 1;
    h->sort.ntiles = (uint32_t) ((nn + kSortTile - 1) / kSortTile);
    if (const char* sm = std::getenv("NMPM_SORT_MINB")) h->sort.scatter_minb = std::atoi(sm);
    CUDA_TRY(h, cudaMalloc(&h->sort.keys_a, nn * sizeof(uint32_t)));
    CUDA_TRY(h, cudaMalloc(&h->sort.keys_b, nn * sizeof(uint32_t)));
    CUDA_TRY(h, cudaMalloc(&h->sort.vals_a, nn * sizeof(uint32_t)));
    CUDA_TRY(h, cudaMalloc(&h->sort.vals_b, nn * sizeof(uint32_t)));
    CUDA_TRY(h, cudaMalloc(&h->sort.hist, (size_t) kRadix * h->sort.ntiles * sizeof(uint32_t)));
    CUDA_TRY(h, cudaMalloc(&h->sort.digit_total, kRadix * sizeof(uint32_t)));
    CUDA_TRY(h, cudaMalloc(&h->sort.digit_base, kRadix * sizeof(uint32_t)));
    CUDA_TRY(h, cudaMalloc(&h->sort.done_counter, sizeof(unsigned int)));
    CUDA_TRY(h, cudaMemset(h->sort.done_counter, 0, sizeof(unsigned int)));
    if (h->slab) {  // unused slots read as "migrated away" until a particle is appended there
        CUDA_TRY(h, cudaMemset(h->sort.keys_a, 0xFF, nn * sizeof(uint32_t)));
        CUDA_TRY(h, cudaMemset(h->sort.keys_b, 0xFF, nn * sizeof(uint32_t)));
        CUDA_TRY(h, cudaMalloc(&h->d_ctr, 4 * sizeof(int)));
        const int ctr0[4] = {(int) n, 0, 0, 0};
        CUDA_TRY(h, cudaMemcpy(h->d_ctr, ctr0, sizeof(ctr0), cudaMemcpyHostToDevice));
    }
    if (batch && h->scenes >= 1) {
        std::vector<unsigned short> scene_of(n ? n : 1);
        h->lame_host.resize((size_t) h->scenes);
        size_t i = 0;
        for (int k = 0; k < h->scenes; ++k) {
            for (size_t c = 0; c < batch->counts[k]; ++c) scene_of[i++] = (unsigned short) k;
            lame_of(batch->E[k], batch->nu[k], &h->lame_host[k].x, &h->lame_host[k].y);
        }
        CUDA_TRY(h, cudaMalloc(&h->d_scene_of, scene_of.size() * sizeof(unsigned short)));
        CUDA_TRY(h, cudaMemcpy(h->d_scene_of, scene_of.data(), scene_of.size() * sizeof(unsigned short), cudaMemcpyHostToDevice));
        CUDA_TRY(h, cudaMalloc(&h->d_lame, (size_t) h->scenes * sizeof(float2)));
        CUDA_TRY(h, cudaMemcpy(h->d_lame, h->lame_host.data(), (size_t) h->scenes * sizeof(float2), cudaMemcpyHostToDevice));
        h->P.mu_0 = h->lame_host[0].x, h->P.lambda_0 = h->lame_host[0].y;
        if (h->scenes > 1) {
            h->P.scenes = h->scenes;
            h->P.scene_of = h->d_scene_of;
            h->P.lame = h->d_lame;
        }
    }
    for (auto& e : h->ev) CUDA_TRY(h, cudaEventCreate(&e));
    return NMPM_OK;
}

extern "C" {

void nmpm_default_options(nmpm_options* opt) {
    if (!opt) return;
    std::memset(opt, 0, sizeof(*opt));
    opt->device = 0;
    opt->sort_every = 4;  // measured optimum on the 3D snow scenes (profiles/r01_sort_cadence.md)
    opt->p2g_variant = 0;
    opt->use_graph = 1;
    opt->slab_x0 = 0;
    opt->slab_x1 = 0;  // 0 = not a slab
    opt->capacity = 0;
}

const char* nmpm_build_info(void) {
    return "libnmpm " __DATE__ " " __TIME__ " sm_100a nvcc " NMPM_STR(__CUDACC_VER_MAJOR__) "." NMPM_STR(__CUDACC_VER_MINOR__);
}

const char* nmpm_last_error(nmpm_handle h) { return h ? h->last_error.c_str() : g_create_error.c_str(); }

void nmpm_destroy(nmpm_handle h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    slab_comm_free(h);
    free_store(h->store[0]);
    free_store(h->store[1]);
    if (h->grid) cudaFree(h->grid);
    if (h->grid_alt) cudaFree(h->grid_alt);
    if (h->d_scene_of) cudaFree(h->d_scene_of);
    if (h->d_lame) cudaFree(h->d_lame);
    for (auto* t : h->tile_ring)
        if (t) cudaFree(t);
    if (h->d_tile_want) cudaFree(h->d_tile_want);
    if (h->d_box) cudaFree(h->d_box);
    if (h->d_box_partial) cudaFree(h->d_box_partial);
    if (h->d_error) cudaFree(h->d_error);
    if (h->d_ctr) cudaFree(h->d_ctr);
    if (h->h_error) cudaFreeHost(h->h_error);
    if (h->staging) cudaFree(h->staging);
    if (h->io.ok) {
        cudaStreamSynchronize(h->io.s_in), cudaStreamSynchronize(h->io.s_out);
        for (int b = 0; b < 2; ++b) {
            cudaFree(h->io.in[b]), cudaFree(h->io.out[b]);
            cudaEventDestroy(h->io.in_ready[b]), cudaEventDestroy(h->io.in_free[b]);
            cudaEventDestroy(h->io.out_ready[b]), cudaEventDestroy(h->io.out_free[b]);
        }
        cudaStreamDestroy(h->io.s_in), cudaStreamDestroy(h->io.s_out);
    }
    cudaFree(h->sort.keys_a);
    cudaFree(h->sort.keys_b);
    cudaFree(h->sort.vals_a);
    cudaFree(h->sort.vals_b);
    cudaFree(h->sort.hist);
    cudaFree(h->sort.digit_total);
    cudaFree(h->sort.digit_base);
    cudaFree(h->sort.done_counter);
    for (auto& e : h->ev)
        if (e) cudaEventDestroy(e);
    for (auto& kv : h->graphs)
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    cudaGetLastError();
    delete h;
}

static int create_soa(int dim, int model, int res, float dt, float E, float nu, float gravity, size_t n, const float* x,
                      const float* v, const float* F, const float* C, const float* Jp, const float* mass,
                      const float* volume, const nmpm_options* opt, nmpm_handle* out, const BatchSpec* batch) {
    nmpm_sim* h = nullptr;
    int rc = create_common(dim, model, res, dt, E, nu, gravity, n, opt, &h, batch);
    if (rc != NMPM_OK) {
        if (h) {
            g_create_error = h->last_error;
            nmpm_destroy(h);
            if (out) *out = nullptr;
        }
        return rc;
    }
    if (n && !x) {
        g_create_error = "x must not be NULL";
        nmpm_destroy(h);
        *out = nullptr;
        return NMPM_ERR_INVALID;
    }
    if (n) {
        // stage the interchange arrays on the device, then pack into the SoA store (K5)
        const size_t D = (size_t) dim;
        const size_t words = n * (2 * D + 2 * D * D + 3);
        rc = ensure_staging(h, words * sizeof(float));
        if (rc == NMPM_OK) {
            float* p = h->staging;
            auto up = [&](const float* src, size_t cnt) -> float* {
                if (!src) return nullptr;
                float* dst = p;
                p += cnt;
                if (cudaMemcpyAsync(dst, src, cnt * sizeof(float), cudaMemcpyHostToDevice, h->stream) != cudaSuccess)
                    rc = NMPM_ERR_CUDA;
                return dst;
            };
            float* dx = up(x, n * D);
            float* dv = up(v, n * D);
            float* dF = up(F, n * D * D);
            float* dC = up(C, n * D * D);
            float* dJ = up(Jp, n);
            float* dm = up(mass, n);
            float* dvol = up(volume, n);
            if (rc == NMPM_OK) {
                NMPM_DISPATCH_DIM(h, (k_import_soa<D><<<blocks_for(n, 256), 256, 0, h->stream>>>(
                                         dx, dv, dF, dC, dJ, dm, dvol, (uint32_t) n, h->store[0], 0)));
                h->launches++;
                if (cudaStreamSynchronize(h->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess)
                    rc = NMPM_ERR_CUDA;
            }
        }
        if (rc != NMPM_OK) {
            g_create_error = h->last_error.empty() ? std::string("particle upload failed: ") +
                                                         cudaGetErrorString(cudaGetLastError())
                                                   : h->last_error;
            nmpm_destroy(h);
            *out = nullptr;
            return rc;
        }
    }
    *out = h;
    return NMPM_OK;
}

int nmpm_create(int dim, int model, int res, float dt, float E, float nu, float gravity, size_t n, const float* x,
                const float* v, const float* F, const float* C, const float* Jp, const float* mass,
                const float* volume, const nmpm_options* opt, nmpm_handle* out) {
    return create_soa(dim, model, res, dt, E, nu, gravity, n, x, v, F, C, Jp, mass, volume, opt, out, nullptr);
}

static bool batch_args_ok(int nscenes, const size_t* counts, const float* E, const float* nu, nmpm_handle* out) {
    if (nscenes >= 1 && counts && E && nu) return true;
    g_create_error = "batch: nscenes >= 1 and counts, E, nu must not be NULL";
    if (out) *out = nullptr;
    return false;
}

int nmpm_create_batch(int model, int res, float dt, float gravity, int nscenes, const size_t* counts, const float* E,
                      const float* nu, const float* x, const float* v, const float* F, const float* C, const float* Jp,
                      const float* mass, const float* volume, const nmpm_options* opt, nmpm_handle* out) {
    if (!batch_args_ok(nscenes, counts, E, nu, out)) return NMPM_ERR_INVALID;
    size_t n = 0;
    for (int k = 0; k < nscenes; ++k) n += counts[k];
    const BatchSpec b{nscenes, counts, E, nu};
    return create_soa(2, model, res, dt, E[0], nu[0], gravity, n, x, v, F, C, Jp, mass, volume, opt, out, &b);
}

static int create_aos(int dim, int model, int res, float dt, float E, float nu, float gravity, size_t n,
                      const void* particles_aos, size_t stride, const nmpm_options* opt, nmpm_handle* out,
                      const BatchSpec* batch);

int nmpm_create_aos(int dim, int model, int res, float dt, float E, float nu, float gravity, size_t n,
                    const void* particles_aos, size_t stride, const nmpm_options* opt, nmpm_handle* out) {
    return create_aos(dim, model, res, dt, E, nu, gravity, n, particles_aos, stride, opt, out, nullptr);
}

int nmpm_create_batch_aos(int model, int res, float dt, float gravity, int nscenes, const size_t* counts, const float* E,
                          const float* nu, const void* particles_aos, size_t stride, const nmpm_options* opt,
                          nmpm_handle* out) {
    if (!batch_args_ok(nscenes, counts, E, nu, out)) return NMPM_ERR_INVALID;
    size_t n = 0;
    for (int k = 0; k < nscenes; ++k) n += counts[k];
    const BatchSpec b{nscenes, counts, E, nu};
    return create_aos(2, model, res, dt, E[0], nu[0], gravity, n, particles_aos, stride, opt, out, &b);
}

int nmpm_num_scenes(nmpm_handle h) { return h ? h->scenes : 0; }

int nmpm_batch_lame(nmpm_handle h, float* mu_0, float* lambda_0) {
    if (!h) return NMPM_ERR_INVALID;
    for (int k = 0; k < h->scenes; ++k) {
        const float2 l = h->lame_host.empty() ? make_float2(h->P.mu_0, h->P.lambda_0) : h->lame_host[(size_t) k];
        if (mu_0) mu_0[k] = l.x;
        if (lambda_0) lambda_0[k] = l.y;
    }
    return NMPM_OK;
}

static int create_aos(int dim, int model, int res, float dt, float E, float nu, float gravity, size_t n,
                      const void* particles_aos, size_t stride, const nmpm_options* opt, nmpm_handle* out,
                      const BatchSpec* batch) {
    const size_t rec = (dim == 3) ? 112 : 64;
    if ((dim == 2 || dim == 3) && (stride < rec || stride % 4 != 0 || (n && !particles_aos))) {
        g_create_error = "AoS stride must be >= sizeof(Particle<dim>) (64 B in 2D, 112 B in 3D) and a multiple of 4";
        if (out) *out = nullptr;
        return NMPM_ERR_INVALID;
    }
    nmpm_sim* h = nullptr;
    int rc = create_common(dim, model, res, dt, E, nu, gravity, n, opt, &h, batch);
    if (rc == NMPM_OK && n) {
        rc = ensure_staging(h, n * stride);
        if (rc == NMPM_OK) {
            cudaError_t e = cudaMemcpyAsync(h->staging, particles_aos, n * stride, cudaMemcpyHostToDevice, h->stream);
            if (e == cudaSuccess) {
                NMPM_DISPATCH_DIM(h, (k_import_aos<D><<<blocks_for(n, 256), 256, 0, h->stream>>>(
                                         h->staging, stride / 4, (uint32_t) n, h->store[0])));
                h->launches++;
                e = cudaStreamSynchronize(h->stream);
            }
            if (e != cudaSuccess) {
                h->last_error = std::string("AoS upload failed: ") + cudaGetErrorString(e);
                rc = NMPM_ERR_CUDA;
            }
        }
    }
    if (rc != NMPM_OK) {
        if (h) {
            g_create_error = h->last_error;
            nmpm_destroy(h);
        }
        if (out) *out = nullptr;
        return rc;
    }
    *out = h;
    return NMPM_OK;
}

size_t nmpm_num_particles(nmpm_handle h) {
    if (!h) return 0;
    if (h->dev_counts) {  // device-driven slab: the host only holds a bound; read the counters (synchronises)
        long long np = 0;
        return nmpm_slab_counts(h, &np, nullptr) == NMPM_OK ? (size_t) np : 0;
    }
    return h->n_store - h->n_gone;
}
size_t nmpm_grid_cells(nmpm_handle h) { return h ? h->cells : 0; }
size_t nmpm_num_slots(nmpm_handle h) { return h ? h->n_store : 0; }
int nmpm_key_tile_bits(nmpm_handle) { return kTileBits; }
long long nmpm_launch_count(nmpm_handle h) { return h ? h->launches : 0; }
int nmpm_fused(nmpm_handle h) { return h ? h->fuse : 0; }
int nmpm_tiles_active(nmpm_handle h) {  // synchronises: test / diagnostic hook
    if (!h || !h->tiles) return 0;
    int w = 0;
    cudaSetDevice(h->device);
    if (cudaStreamSynchronize(h->stream) != cudaSuccess ||
        cudaMemcpy(&w, h->d_tile_want + h->box_cur, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return w;
}

int nmpm_lame(nmpm_handle h, float* mu_0, float* lambda_0) {
    if (!h) return NMPM_ERR_INVALID;
    if (mu_0) *mu_0 = h->P.mu_0;
    if (lambda_0) *lambda_0 = h->P.lambda_0;
    return NMPM_OK;
}

int nmpm_set_stream(nmpm_handle h, void* cuda_stream) {
    if (!h) return NMPM_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (h->own_stream) cudaStreamDestroy(h->stream);
    h->stream = (cudaStream_t) cuda_stream;
    h->own_stream = false;
    return NMPM_OK;
}
void* nmpm_get_stream(nmpm_handle h) { return h ? (void*) h->stream : nullptr; }

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// the step
// ---------------------------------------------------------------------------------------------
// Active node tiles of the positions whose cell keys are `keys` (flags of ring slot `box`); see k_mark_tiles.
// (Flags left in the slot by a replaced state are only a superset.)
static void mark_tiles(nmpm_sim* h, const uint32_t* keys, uint32_t n, int box) {
    if (!h->tiles) return;
    // decide for these positions (hysteresis against the previous slot's decision), then raise the flags if wanted
    k_tile_decide<<<1, 32, 0, h->stream>>>(h->d_box + box, n, h->P.n1, h->d_tile_want + (box + kBoxRing - 1) % kBoxRing,
                                           h->d_tile_want + box, h->tile_policy);
    h->launches++;
    if (n == 0) return;
    const unsigned per_block = kMarkWarps * kMarkKeysPerWarp;
    k_mark_tiles<<<(n + per_block - 1) / per_block, kMarkWarps * 32, 0, h->stream>>>(keys, n, h->tile_ring[box],
                                                                                    (h->P.n1 + 3) >> 2, h->d_tile_want + box);
    h->launches++;
}

// K0: (keys from the previous G2P, or a key pass) + radix sort.  The reorder itself is fused into the
// readers: P2G and G2P index the store through h->perm, and G2P writes the other store in sorted order.
static int do_sort(nmpm_sim* h) {
    h->perm = nullptr;
    if (h->n_store == 0) return NMPM_OK;
    const uint32_t n = (uint32_t) h->n_store;
    ParticleStore& S = h->store[h->cur];
    if (!h->keys_valid) {
        // (first step / after an upload) the key pass also builds the node box of the current positions
        k_box_reset<<<1, 32, 0, h->stream>>>(h->d_box + h->box_cur);
        NMPM_DISPATCH_DIM(h, (k_cell_keys<D><<<blocks_for(n, 256), 256, 0, h->stream>>>(
                                 S, n, h->P, h->tiles_per_axis, h->sort.keys_a, nullptr, h->d_error,
                                 h->d_box + h->box_cur)));
        mark_tiles(h, h->sort.keys_a, n, h->box_cur);
        h->launches += 2;
        h->box_valid = true;
    }
    uint32_t *ks = nullptr, *perm = nullptr;
    h->launches += radix_sort_pairs(h->sort, n, h->key_bits, h->stream, &ks, &perm);
    h->keys_valid = false;
    h->perm = perm;
    if (h->dev_counts) {
        // the sorted order ends with the migrated-away and the unused slots; they are skipped through their sorted keys,
        // which must sit in keys_a (the buffer G2P writes the next keys into, slot by slot)
        if (ks != h->sort.keys_a)
            CUDA_TRY(h, cudaMemcpyAsync(h->sort.keys_a, ks, (size_t) n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, h->stream));
        k_ctr_after_sort<<<1, 32, 0, h->stream>>>(h->d_ctr);
        h->launches++;
        h->n = h->n_store;
        return NMPM_OK;
    }
    // slab mode: particles that migrated away carry kKeyGone and sort last; they drop out here
    h->n = h->n_store - h->n_gone;
    h->n_gone = 0;
    return NMPM_OK;
}

constexpr int kBoxBlocks = 148 * 8;  // grid-stride kernels over a node box: 8 CTAs of 256 threads per SM

// make box[box_cur] valid when no key pass ran for the current positions (sort_every == 0 on the first step)
static int ensure_box(nmpm_sim* h) {
    if (h->box_valid || h->n_store == 0) return NMPM_OK;
    const uint32_t n = (uint32_t) h->n_store;
    k_box_reset<<<1, 32, 0, h->stream>>>(h->d_box + h->box_cur);
    NMPM_DISPATCH_DIM(h, (k_cell_keys<D><<<blocks_for(n, 256), 256, 0, h->stream>>>(
                             h->store[h->cur], n, h->P, h->tiles_per_axis, h->sort.keys_b, nullptr, h->d_error,
                             h->d_box + h->box_cur)));
    mark_tiles(h, h->sort.keys_b, n, h->box_cur);
    h->launches += 2;
    h->box_valid = true;
    return NMPM_OK;
}

// zero what the P2G of the positions of box `box` scattered into `grid`: their flagged tiles (lowering the flags), or the
// node box itself
static void clear_grid(nmpm_sim* h, float4* grid, int box) {
    const int* want = h->tiles ? h->d_tile_want + box : nullptr;
    if (want) {
        if (h->tiles_u == 8) k_tiles3<0, 8><<<kBoxBlocks, 256, 0, h->stream>>>(grid, h->tile_ring[box], h->P, want);
        else if (h->tiles_u == 2) k_tiles3<0, 2><<<kBoxBlocks, 256, 0, h->stream>>>(grid, h->tile_ring[box], h->P, want);
        else k_tiles3<0, 4><<<kBoxBlocks, 256, 0, h->stream>>>(grid, h->tile_ring[box], h->P, want);
        h->launches++;
    }
    NMPM_DISPATCH_DIM(h, (k_clear_box<D><<<kBoxBlocks, 256, 0, h->stream>>>(grid, h->d_box + box, h->P.n1,
                                                                            h->scenes > 1 ? h->scenes * h->P.n1 : 0, want)));
    h->launches++;
}

// An out-of-grid particle scatters to clamped nodes that neither its key (kKeyOutOfGrid) nor, in tile mode, any flag
// accounts for.  The step is flagged and the state is void, but the caller may upload a repaired state into the same
// handle: start it from dense zero grids and lowered flags.
static void reset_grids_after_error(nmpm_sim* h) {
    if (!h->tiles || !(h->error_latched || (h->h_error && *((volatile int*) h->h_error) != 0))) return;
    const size_t T = ((size_t) h->P.n1 + 3) / 4, words = (T * T * T + 31) / 32;
    cudaMemsetAsync(h->grid, 0, h->cells * sizeof(float4), h->stream);
    if (h->grid_alt) cudaMemsetAsync(h->grid_alt, 0, h->cells * sizeof(float4), h->stream);
    for (auto* t : h->tile_ring) cudaMemsetAsync(t, 0, words * sizeof(uint32_t), h->stream);
    cudaMemsetAsync(h->d_tile_want, 0, kBoxRing * sizeof(int), h->stream);
    for (int k = 0; k < kBoxRing; ++k) k_box_reset<<<1, 32, 0, h->stream>>>(h->d_box + k);
    h->p2g_ahead = false;
    h->box_valid = false;
}

// The particle state is about to be replaced (upload): the sums the last fused G2P scattered ahead into grid_alt belong to
// the old state.  They cover box[box_cur]; zero them (before the key pass of the new state resets that box).
static void discard_p2g_ahead(nmpm_sim* h) {
    if (!h->p2g_ahead) return;
    clear_grid(h, h->grid_alt, h->box_cur);
    h->p2g_ahead = false;
}

static int do_p2g(nmpm_sim* h) {
    if (h->p2g_ahead) {
        // the previous step's fused G2P already scattered this step into grid_alt: swap the buffers, take over what its
        // scatter half found, and zero the grid of the previous step (now grid_alt) over the box its P2G wrote
        std::swap(h->grid, h->grid_alt);
        h->grid_sel ^= 1;
        h->p2g_ahead = false;
        k_promote_error<<<1, 32, 0, h->stream>>>(h->d_error);
        h->launches++;
        clear_grid(h, h->grid_alt, (h->box_cur + kBoxRing - 1) % kBoxRing);
        h->grid_valid = true;
        return NMPM_OK;
    }
    if (int rc = ensure_box(h)) return rc;
    // K1: clear what the previous P2G (and, for a slab, the neighbours' ghost planes) wrote
    clear_grid(h, h->grid, (h->box_cur + kBoxRing - 1) % kBoxRing);
    for (const auto& pr : h->dirty_planes)
        CUDA_TRY(h, cudaMemsetAsync(nmpm_grid_plane_ptr(h, pr.first), 0, (size_t) pr.second * nmpm_grid_plane_bytes(h),
                                    h->stream));
    h->dirty_planes.clear();
    for (const auto& r : h->dirty_rects) {
        const size_t nodes = (size_t) 2 * r[2] * r[4];
        k_rect<2><<<blocks_for(nodes, 256), 256, 0, h->stream>>>(h->grid, nmpm_grid_plane_bytes(h) / sizeof(float4), h->P.n1, r[0],
                                                                 2, r[1], r[2], r[3], r[4], nullptr);
        h->launches++;
    }
    h->dirty_rects.clear();
    h->grid_valid = true;
    if (h->n == 0) return NMPM_OK;
    const uint32_t n = (uint32_t) h->n;
    ParticleStore& S = h->store[h->cur];
    // slab mode, step without a sort: slots of migrated-away particles are still in the store
    const uint32_t* gone_keys = (h->dev_counts || (h->slab && !h->perm && h->n_gone)) ? h->sort.keys_a : nullptr;
    int variant = h->opt.p2g_variant;
    // auto: per-particle reductions without binning; with binning the per-column-packet kernel, on three streams per
    // warp once the scene is large enough for the reductions to miss L2 (measured: profiles/r01g)
    if (variant == 0) variant = (h->opt.sort_every > 0) ? ((h->dim == 3 && h->n >= ((size_t) 8 << 20)) ? 4 : 3) : 1;
    if (variant == 4) {
        NMPM_DISPATCH(h, (launch_p2g_streams<D, MODEL>(S, h->perm, n, h->P, h->grid, h->d_error, gone_keys, h->stream)));
    } else if (variant > 40 && variant < 50) {  // tests / experiments: 4C = C chunks of 32 slots per warp
        NMPM_DISPATCH(h, (launch_p2g_streams<D, MODEL>(S, h->perm, n, h->P, h->grid, h->d_error, gone_keys, h->stream,
                                                       variant - 40)));
    } else if (variant == 3) {
        NMPM_DISPATCH(h, (launch_p2g_cols<D, MODEL>(S, h->perm, n, h->P, h->grid, h->d_error, gone_keys, h->stream)));
    } else if (variant == 2) {
        NMPM_DISPATCH(h, (launch_p2g_cell<D, MODEL>(S, h->perm, n, h->P, h->grid, h->d_error, gone_keys, h->stream)));
    } else {
        NMPM_DISPATCH(h, (k_p2g_scatter<D, MODEL><<<blocks_for(n, 128), 128, 0, h->stream>>>(S, h->perm, n, h->P,
                                                                                            h->grid, h->d_error, gone_keys, 0u)));
    }
    h->launches++;
    return NMPM_OK;
}

static int do_grid_op(nmpm_sim* h) {
    const int* want = h->tiles ? h->d_tile_want + h->box_cur : nullptr;
    if (want) {
        if (h->tiles_u == 8) k_tiles3<1, 8><<<kBoxBlocks, 256, 0, h->stream>>>(h->grid, h->tile_ring[h->box_cur], h->P, want);
        else if (h->tiles_u == 2) k_tiles3<1, 2><<<kBoxBlocks, 256, 0, h->stream>>>(h->grid, h->tile_ring[h->box_cur], h->P, want);
        else k_tiles3<1, 4><<<kBoxBlocks, 256, 0, h->stream>>>(h->grid, h->tile_ring[h->box_cur], h->P, want);
        h->launches++;
    }
    NMPM_DISPATCH_DIM(h, (k_grid_op<D><<<kBoxBlocks, 256, 0, h->stream>>>(h->grid, h->d_box + h->box_cur, h->P, want)));
    h->launches++;
    return NMPM_OK;
}

static int do_g2p(nmpm_sim* h, const MigrateArgs& mig = MigrateArgs{0, 0, nullptr, nullptr, 0, nullptr}) {
    if (h->n == 0) {
        h->n_store = 0;
        h->box_cur = (h->box_cur + 1) % kBoxRing;
        k_box_reset<<<1, 32, 0, h->stream>>>(h->d_box + h->box_cur);
        if (h->tiles) CUDA_TRY(h, cudaMemsetAsync(h->d_tile_want + h->box_cur, 0, sizeof(int), h->stream));
        // an empty slab: its key array is all "gone" marks, and stays aligned with the slots as particles arrive —
        // no key pass may ever run over the (uninitialised) store
        if (h->slab) h->keys_valid = true;
        return NMPM_OK;
    }
    const uint32_t n = (uint32_t) h->n;
    ParticleStore& S = h->store[h->cur];
    ParticleStore& T = h->perm ? h->store[h->cur ^ 1] : S;
    // emit the next step's keys only if the next step sorts
    const bool next_sorts = h->opt.sort_every > 0 && ((h->steps_done + 1) % h->opt.sort_every) == 0;
    // a slab keeps the key array aligned with the slots at all times: it carries the "migrated away" marks
    // (and in tile mode: the active node tiles of the coming step are raised from them, mark_tiles below)
    uint32_t* keys_out = (next_sorts || h->slab || h->tiles) ? h->sort.keys_a : nullptr;
    const uint32_t* gone_keys = (h->dev_counts || (h->slab && !h->perm && h->n_gone)) ? h->sort.keys_a : nullptr;
    const int box_next = (h->box_cur + 1) % kBoxRing;
    k_box_reset<<<1, 32, 0, h->stream>>>(h->d_box + box_next);
    // fused G2P+P2G: not on the first step after an upload (fuse == 1) — that state is likely to be replaced again
    const bool fused = h->fuse && !mig.left && (h->fuse == 2 || h->steps_done > 0);
    if (fused) {
        if (h->model == 0)
            launch_g2p_p2g<0>(S, T, h->perm, n, h->P, h->grid, h->grid_alt, keys_out, h->tiles_per_axis, h->d_error,
                              h->d_box_partial, h->local_reorder ? 1 : 0, h->stream, h->fused_minb);
        else if (h->model == 1)
            launch_g2p_p2g<1>(S, T, h->perm, n, h->P, h->grid, h->grid_alt, keys_out, h->tiles_per_axis, h->d_error,
                              h->d_box_partial, h->local_reorder ? 1 : 0, h->stream, h->fused_minb);
        else
            launch_g2p_p2g<2>(S, T, h->perm, n, h->P, h->grid, h->grid_alt, keys_out, h->tiles_per_axis, h->d_error,
                              h->d_box_partial, h->local_reorder ? 1 : 0, h->stream, h->fused_minb);
        h->p2g_ahead = true;
    } else if (h->g2p_window && h->g2p_pipe) {  // 3D only: persistent CTAs, pipelined rows (cp.async) and node windows (TMA)
        const unsigned chunks = blocks_for(n, 128);
        const unsigned grid_ctas = chunks < (unsigned) (148 * NMPM_G2P_PIPE_MINB) ? chunks : (unsigned) (148 * NMPM_G2P_PIPE_MINB);
        if (h->model == 0)
            k_g2p_pipe<0><<<grid_ctas, 128, 0, h->stream>>>(S, T, h->perm, n, h->P, h->grid, keys_out, h->tiles_per_axis, h->d_error,
                                                            mig, h->d_box_partial, gone_keys, h->local_reorder ? 1 : 0, h->grid_map);
        else if (h->model == 1)
            k_g2p_pipe<1><<<grid_ctas, 128, 0, h->stream>>>(S, T, h->perm, n, h->P, h->grid, keys_out, h->tiles_per_axis, h->d_error,
                                                            mig, h->d_box_partial, gone_keys, h->local_reorder ? 1 : 0, h->grid_map);
        else
            k_g2p_pipe<2><<<grid_ctas, 128, 0, h->stream>>>(S, T, h->perm, n, h->P, h->grid, keys_out, h->tiles_per_axis, h->d_error,
                                                            mig, h->d_box_partial, gone_keys, h->local_reorder ? 1 : 0, h->grid_map);
    } else if (h->g2p_window) {  // 3D only (set at creation): one-shot CTAs with a TMA-staged node window
        constexpr int D = 3;
        if (h->model == 0)
            k_g2p_gather<D, 0, true><<<blocks_for(n, 128), 128, 0, h->stream>>>(S, T, h->perm, n, h->P, h->grid, keys_out,
                                                                               h->tiles_per_axis, h->d_error, mig, h->d_box_partial,
                                                                               gone_keys, h->local_reorder ? 1 : 0, h->grid_map);
        else if (h->model == 1)
            k_g2p_gather<D, 1, true><<<blocks_for(n, 128), 128, 0, h->stream>>>(S, T, h->perm, n, h->P, h->grid, keys_out,
                                                                               h->tiles_per_axis, h->d_error, mig, h->d_box_partial,
                                                                               gone_keys, h->local_reorder ? 1 : 0, h->grid_map);
        else
            k_g2p_gather<D, 2, true><<<blocks_for(n, 128), 128, 0, h->stream>>>(S, T, h->perm, n, h->P, h->grid, keys_out,
                                                                               h->tiles_per_axis, h->d_error, mig, h->d_box_partial,
                                                                               gone_keys, h->local_reorder ? 1 : 0, h->grid_map);
    } else {
        NMPM_DISPATCH(h, (k_g2p_gather<D, MODEL, false><<<blocks_for(n, 128), 128, 0, h->stream>>>(
                             S, T, h->perm, n, h->P, h->grid, keys_out, h->tiles_per_axis, h->d_error, mig,
                             h->d_box_partial, gone_keys, h->local_reorder ? 1 : 0, h->grid_map)));
    }
    {   // warps of the launch above (128-thread CTAs): every one of them wrote a partial box
        const uint32_t nwarps = blocks_for(n, 128) * 4u - ((blocks_for(n, 128) * 128u - n) / 32u);
        const unsigned rb = nwarps / 1024 + 1 < 148u ? nwarps / 1024 + 1 : 148u;
        k_box_reduce<<<rb, 256, 0, h->stream>>>(h->d_box_partial, nwarps, h->d_box + box_next);
    }
    h->launches += 3;
    mark_tiles(h, h->sort.keys_a, n, box_next);
    h->box_cur = box_next;
    if (h->perm) h->cur ^= 1;
    h->perm = nullptr;
    h->keys_valid = keys_out != nullptr;
    h->n_store = h->n;  // a sorted write compacted the store; an in-place step keeps its slots (incl. gone ones)
    return NMPM_OK;
}

static int run_phase(nmpm_sim* h, int phase) {
    int rc = NMPM_OK;
    if (phase == NMPM_PHASE_P2G) {
        const bool sort_now = h->opt.sort_every > 0 && (h->steps_done % h->opt.sort_every) == 0;
        if (h->timing) cudaEventRecord(h->ev[0], h->stream);
        if (sort_now) rc = do_sort(h);
        if (h->timing) cudaEventRecord(h->ev[1], h->stream);
        if (rc == NMPM_OK) rc = do_p2g(h);
        if (h->timing) cudaEventRecord(h->ev[2], h->stream);
    } else if (phase == NMPM_PHASE_GRID_OP) {
        rc = do_grid_op(h);
        if (h->timing) cudaEventRecord(h->ev[3], h->stream);
    } else {
        rc = do_g2p(h);
        if (h->timing) {
            cudaEventRecord(h->ev[4], h->stream);
            cudaEventSynchronize(h->ev[4]);
            float ms = 0;
            cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
            h->t_ms[NMPM_T_SORT] += ms;
            cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]);
            h->t_ms[NMPM_T_P2G] += ms;
            cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]);
            h->t_ms[NMPM_T_GRID] += ms;
            cudaEventElapsedTime(&ms, h->ev[3], h->ev[4]);
            h->t_ms[NMPM_T_G2P] += ms;
            h->t_steps++;
        }
        h->steps_done++;
    }
    if (rc == NMPM_OK) {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            h->last_error = std::string("kernel launch failed: ") + cudaGetErrorString(e);
            rc = NMPM_ERR_CUDA;
        }
    }
    return rc;
}

// non-blocking look at the error flag copied back by earlier calls
static int poll_error(nmpm_sim* h) {
    if (h->error_latched || (h->h_error && *((volatile int*) h->h_error) != 0)) {
        h->error_latched = true;
        h->last_error = "a particle's stencil left the grid (reference: std::out_of_range from vector::at, "
                        "src/nclr.h:163)";
        return NMPM_ERR_OUT_OF_GRID;
    }
    return NMPM_OK;
}

static int sync_and_check(nmpm_sim* h) {
    CUDA_TRY(h, cudaMemcpyAsync(h->h_error, h->d_error, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (h->io.ok) {  // pipelined host I/O in flight
        CUDA_TRY(h, cudaStreamSynchronize(h->io.s_in));
        CUDA_TRY(h, cudaStreamSynchronize(h->io.s_out));
    }
    return poll_error(h);
}

extern "C" {

// `count` whole steps.  With opt.use_graph the launch sequence of each distinct host-side step state is captured once
// into a CUDA graph and replayed afterwards (one launch per step instead of ~12) — and, for runs of steps, a whole CYCLE
// of host states (sort cadence x store parity x node-box ring: 8 steps at the default cadence) is captured into ONE graph,
// so a small scene costs the host one launch per 8 steps (config 5: thousands of steps of 1 250-particle scenes).
static int graph_steps(nmpm_sim* h, int count) {
    const bool can_graph = h->opt.use_graph && h->graphs_ok && !h->timing && h->phase_next == 0 && h->n > 0;
    if (!can_graph) {
        for (int k = 0; k < count; ++k) {
            for (int ph = h->phase_next; ph < 3; ++ph)
                if (int rc = run_phase(h, ph)) return rc;
            h->phase_next = 0;
        }
        return NMPM_OK;
    }
    const int se = h->opt.sort_every;
    const int box0 = h->box_cur;
    const bool boxv0 = h->box_valid;
    const bool ahead0 = h->p2g_ahead;
    const int gsel0 = h->grid_sel;
    const int key = h->cur | (h->keys_valid ? 2 : 0) | (box0 << 2) | (boxv0 ? 16 : 0) |
                    ((se > 0 ? (int) (h->steps_done % se) : 0) << 5) | (count << 12) | (ahead0 ? 1 << 20 : 0) | (gsel0 << 21) |
                    ((h->fuse == 1 && h->steps_done == 0) ? 1 << 22 : 0);
    auto it = h->graphs.find(key);
    if (it == h->graphs.end()) {
        if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            cudaGetLastError();
            h->graphs_ok = false;  // e.g. the legacy default stream: fall back to plain launches
            return graph_steps(h, count);
        }
        const long long l0 = h->launches, s0 = h->steps_done;
        int rc = NMPM_OK;
        for (int k = 0; k < count && rc == NMPM_OK; ++k)
            for (int ph = 0; ph < 3 && rc == NMPM_OK; ++ph) rc = run_phase(h, ph);
        cudaGraph_t g = nullptr;
        cudaError_t e = cudaStreamEndCapture(h->stream, &g);
        nmpm_sim::StepGraph sg;
        if (rc == NMPM_OK && e == cudaSuccess && g) e = cudaGraphInstantiate(&sg.exec, g, 0);
        if (g) cudaGraphDestroy(g);
        if (rc != NMPM_OK || e != cudaSuccess || !sg.exec) {
            cudaGetLastError();
            h->graphs_ok = false;
            h->last_error = "CUDA graph capture failed; falling back to plain launches";
            // host-side state already advanced by the capture pass: undo and run the steps for real
            h->steps_done = s0;
            h->launches = l0;
            h->cur = key & 1;
            h->keys_valid = (key & 2) != 0;
            h->box_cur = box0;
            h->box_valid = boxv0;
            h->perm = nullptr;
            if (h->grid_sel != gsel0) std::swap(h->grid, h->grid_alt), h->grid_sel = gsel0;
            h->p2g_ahead = ahead0;
            return graph_steps(h, count);
        }
        sg.cur_after = h->cur;
        sg.keys_valid_after = h->keys_valid;
        sg.box_after = h->box_cur;
        sg.ahead_after = h->p2g_ahead;
        sg.grid_sel_after = h->grid_sel;
        sg.launches = (int) (h->launches - l0);
        it = h->graphs.emplace(key, sg).first;
        CUDA_TRY(h, cudaGraphLaunch(it->second.exec, h->stream));
        return NMPM_OK;  // host state was advanced while capturing
    }
    CUDA_TRY(h, cudaGraphLaunch(it->second.exec, h->stream));
    h->cur = it->second.cur_after;
    h->keys_valid = it->second.keys_valid_after;
    h->box_cur = it->second.box_after;
    h->box_valid = true;
    h->perm = nullptr;
    h->grid_valid = true;
    h->p2g_ahead = it->second.ahead_after;
    if (h->grid_sel != it->second.grid_sel_after) std::swap(h->grid, h->grid_alt), h->grid_sel = it->second.grid_sel_after;
    h->steps_done += count;
    h->launches += it->second.launches;
    return NMPM_OK;
}

// host states repeat with this period (sort cadence x store parity, node-box ring)
static int graph_cycle(const nmpm_sim* h) {
    const int se = h->opt.sort_every;
    int c = se > 0 ? 2 * se : 1;
    while (c % kBoxRing) c += (se > 0 ? 2 * se : 1);
    return c;
}

static int not_for_slabs(nmpm_sim* h, const char* what) {
    if (h->slab) {
        h->last_error = std::string(what) + ": not available on a slab sim (use the nmpm_slab_* calls)";
        return NMPM_ERR_INVALID;
    }
    return NMPM_OK;
}

int nmpm_advance(nmpm_handle h, int nsteps) {
    if (!h || nsteps < 0) return NMPM_ERR_INVALID;
    if (int rc = not_for_slabs(h, "nmpm_advance")) return rc;
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (int rc = poll_error(h)) return rc;
    const int cyc = graph_cycle(h);
    for (int s = 0; s < nsteps;) {
        // whole cycles in one graph once the state is periodic (keys and box valid: i.e. not the very first step)
        const int k = (nsteps - s >= cyc && cyc <= 64 && h->steps_done >= cyc && h->steps_done % cyc == 0) ? cyc : 1;
        if (int rc = graph_steps(h, k)) return rc;
        s += k;
    }
    CUDA_TRY(h, cudaMemcpyAsync(h->h_error, h->d_error, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    return NMPM_OK;
}

int nmpm_phase(nmpm_handle h, int phase) {
    if (!h || phase < 0 || phase > 2) return NMPM_ERR_INVALID;
    if (int rc = not_for_slabs(h, "nmpm_phase")) return rc;
    if (phase != h->phase_next) {
        h->last_error = "nmpm_phase: phases must be called in order p2g, grid_op, g2p";
        return NMPM_ERR_INVALID;
    }
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (int rc = run_phase(h, phase)) return rc;
    h->phase_next = (phase + 1) % 3;
    return sync_and_check(h);
}

int nmpm_synchronize(nmpm_handle h) {
    if (!h) return NMPM_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    return sync_and_check(h);
}

int nmpm_download_particles(nmpm_handle h, float* x, float* v, float* F, float* C, float* Jp) {
    if (!h) return NMPM_ERR_INVALID;
    if (int rc = not_for_slabs(h, "nmpm_download_particles")) return rc;
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t n = h->n, D = (size_t) h->dim;
    if (n) {
        const size_t words = n * (2 * D + 2 * D * D + 1);
        if (int rc = ensure_staging(h, words * sizeof(float))) return rc;
        float* p = h->staging;
        float* dx = x ? p : nullptr;
        p += x ? n * D : 0;
        float* dv = v ? p : nullptr;
        p += v ? n * D : 0;
        float* dF = F ? p : nullptr;
        p += F ? n * D * D : 0;
        float* dC = C ? p : nullptr;
        p += C ? n * D * D : 0;
        float* dJ = Jp ? p : nullptr;
        NMPM_DISPATCH_DIM(h, (k_export_soa<D><<<blocks_for(n, 256), 256, 0, h->stream>>>(h->store[h->cur], (uint32_t) n,
                                                                                         dx, dv, dF, dC, dJ, nullptr, nullptr)));
        h->launches++;
        if (x) CUDA_TRY(h, cudaMemcpyAsync(x, dx, n * D * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        if (v) CUDA_TRY(h, cudaMemcpyAsync(v, dv, n * D * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        if (F) CUDA_TRY(h, cudaMemcpyAsync(F, dF, n * D * D * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        if (C) CUDA_TRY(h, cudaMemcpyAsync(C, dC, n * D * D * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        if (Jp) CUDA_TRY(h, cudaMemcpyAsync(Jp, dJ, n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    }
    return sync_and_check(h);
}

int nmpm_download_positions(nmpm_handle h, float* x) { return nmpm_download_particles(h, x, nullptr, nullptr, nullptr, nullptr); }

int nmpm_download_particles_aos(nmpm_handle h, void* particles_aos, size_t stride) {
    if (!h || !particles_aos) return NMPM_ERR_INVALID;
    const size_t rec = (h->dim == 3) ? 112 : 64;
    if (stride < rec || stride % 4) return NMPM_ERR_INVALID;
    if (int rc = not_for_slabs(h, "nmpm_download_particles_aos")) return rc;
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t n = h->n;
    if (n) {
        if (int rc = ensure_staging(h, n * stride)) return rc;
        // mass / volume / colour never change on the device: stage the caller's records so that the
        // untouched words survive the round trip
        CUDA_TRY(h, cudaMemcpyAsync(h->staging, particles_aos, n * stride, cudaMemcpyHostToDevice, h->stream));
        NMPM_DISPATCH_DIM(h, (k_export_aos<D><<<blocks_for(n, 256), 256, 0, h->stream>>>(h->store[h->cur], (uint32_t) n,
                                                                                         h->staging, stride / 4)));
        h->launches++;
        CUDA_TRY(h, cudaMemcpyAsync(particles_aos, h->staging, n * stride, cudaMemcpyDeviceToHost, h->stream));
    }
    return sync_and_check(h);
}

int nmpm_download_grid(nmpm_handle h, float* gv, float* gm, size_t* cells_out) {
    if (!h) return NMPM_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (!h->grid_valid) {
        if (cells_out) *cells_out = 0;
        return sync_and_check(h);
    }
    const size_t cells = h->cells, D = (size_t) h->dim;
    if (int rc = ensure_staging(h, cells * (D + 1) * sizeof(float))) return rc;
    float* dgv = h->staging;
    float* dgm = h->staging + cells * D;
    NMPM_DISPATCH_DIM(h, (k_export_grid<D><<<blocks_for(cells, 256), 256, 0, h->stream>>>(h->grid, cells, dgv, dgm,
                                                                                          nullptr, 0)));
    h->launches++;
    if (gv) CUDA_TRY(h, cudaMemcpyAsync(gv, dgv, cells * D * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (gm) CUDA_TRY(h, cudaMemcpyAsync(gm, dgm, cells * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (cells_out) *cells_out = cells;
    return sync_and_check(h);
}

int nmpm_download_grid_aos(nmpm_handle h, void* cells_aos, size_t stride, size_t* cells_out) {
    if (!h || !cells_aos) return NMPM_ERR_INVALID;
    const size_t rec = (size_t) (h->dim + 1) * 4;
    if (stride < rec || stride % 4) return NMPM_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (!h->grid_valid) {
        if (cells_out) *cells_out = 0;
        return sync_and_check(h);
    }
    const size_t cells = h->cells;
    if (int rc = ensure_staging(h, cells * stride)) return rc;
    if (stride != rec) CUDA_TRY(h, cudaMemsetAsync(h->staging, 0, cells * stride, h->stream));
    NMPM_DISPATCH_DIM(h, (k_export_grid<D><<<blocks_for(cells, 256), 256, 0, h->stream>>>(h->grid, cells, nullptr,
                                                                                          nullptr, h->staging,
                                                                                          stride / 4)));
    h->launches++;
    CUDA_TRY(h, cudaMemcpyAsync(cells_aos, h->staging, cells * stride, cudaMemcpyDeviceToHost, h->stream));
    if (cells_out) *cells_out = cells;
    return sync_and_check(h);
}

int nmpm_upload_particles(nmpm_handle h, const float* x, const float* v, const float* F, const float* C,
                          const float* Jp) {
    if (!h || (h->n && !x)) return NMPM_ERR_INVALID;
    if (int rc = not_for_slabs(h, "nmpm_upload_particles")) return rc;
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t n = h->n, D = (size_t) h->dim;
    if (n == 0) return NMPM_OK;
    const size_t words = n * (2 * D + 2 * D * D + 1);
    if (int rc = ensure_staging(h, words * sizeof(float))) return rc;
    float* p = h->staging;
    cudaError_t e = cudaSuccess;
    auto up = [&](const float* src, size_t cnt) -> float* {
        if (!src) return nullptr;
        float* dst = p;
        p += cnt;
        cudaError_t ee = cudaMemcpyAsync(dst, src, cnt * sizeof(float), cudaMemcpyHostToDevice, h->stream);
        if (ee != cudaSuccess) e = ee;
        return dst;
    };
    float* dx = up(x, n * D);
    float* dv = up(v, n * D);
    float* dF = up(F, n * D * D);
    float* dC = up(C, n * D * D);
    float* dJ = up(Jp, n);
    CUDA_TRY(h, e);
    ParticleStore& S = h->store[h->cur];
    ParticleStore& T = h->store[h->cur ^ 1];
    reset_grids_after_error(h);
    discard_p2g_ahead(h);
    if (h->phase_next != 0) {
        // Upload in the middle of a step (after nmpm_phase(P2G) or (GRID_OP)): the aborted step's P2G has already
        // written the nodes of box[box_cur], and the next step only clears box[box_cur-1] (which that P2G cleared
        // itself) before the key pass rebuilds box[box_cur].  Zero those nodes now, or later steps would accumulate on
        // top of the stale sums.
        clear_grid(h, h->grid, h->box_cur);
    }
    // slots go back to input order: carry mass/volume over through id
    NMPM_DISPATCH_DIM(h, (k_restore_constants<D><<<blocks_for(n, 256), 256, 0, h->stream>>>(S, T, (uint32_t) n)));
    NMPM_DISPATCH_DIM(h, (k_import_soa<D><<<blocks_for(n, 256), 256, 0, h->stream>>>(dx, dv, dF, dC, dJ, nullptr,
                                                                                     nullptr, (uint32_t) n, T, 1)));
    h->launches += 2;
    h->cur ^= 1;
    h->phase_next = 0;
    h->keys_valid = false;
    h->box_valid = false;  // rebuilt by the next key pass; the nodes of the last P2G stay scheduled for clearing
    h->perm = nullptr;
    // the sort cadence restarts so that the next step re-bins the new state
    h->steps_done = 0;
    CUDA_TRY(h, cudaMemsetAsync(h->d_error, 0, 2 * sizeof(int), h->stream));
    // the host copy of the flag is reset only after the stream has drained: an error-flag copy of an earlier
    // nmpm_advance may still be in flight and would otherwise re-raise the old state's error for the new one
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    *h->h_error = 0;
    h->error_latched = false;
    return NMPM_OK;
}

static int async_io_init(nmpm_sim* h) {
    if (h->io.ok) return NMPM_OK;
    const size_t D = (size_t) h->dim;
    h->io.bytes = (h->n ? h->n : 1) * (2 * D + 2 * D * D + 1) * sizeof(float);
    CUDA_TRY(h, cudaStreamCreateWithFlags(&h->io.s_in, cudaStreamNonBlocking));
    CUDA_TRY(h, cudaStreamCreateWithFlags(&h->io.s_out, cudaStreamNonBlocking));
    for (int b = 0; b < 2; ++b) {
        CUDA_TRY(h, cudaMalloc(&h->io.in[b], h->io.bytes));
        CUDA_TRY(h, cudaMalloc(&h->io.out[b], h->io.bytes));
        CUDA_TRY(h, cudaEventCreateWithFlags(&h->io.in_ready[b], cudaEventDisableTiming));
        CUDA_TRY(h, cudaEventCreateWithFlags(&h->io.in_free[b], cudaEventDisableTiming));
        CUDA_TRY(h, cudaEventCreateWithFlags(&h->io.out_ready[b], cudaEventDisableTiming));
        CUDA_TRY(h, cudaEventCreateWithFlags(&h->io.out_free[b], cudaEventDisableTiming));
    }
    h->io.ok = true;
    return NMPM_OK;
}

// Pipelined form of nmpm_upload_particles: returns as soon as the copies are enqueued.  The host arrays (pinned memory
// for real overlap) must stay untouched until nmpm_synchronize or until two more async uploads have been issued.
int nmpm_upload_particles_async(nmpm_handle h, const float* x, const float* v, const float* F, const float* C, const float* Jp) {
    if (!h || (h->n && !x)) return NMPM_ERR_INVALID;
    if (int rc = not_for_slabs(h, "nmpm_upload_particles_async")) return rc;
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t n = h->n, D = (size_t) h->dim;
    if (n == 0) return NMPM_OK;
    if (int rc = async_io_init(h)) return rc;
    const int b = (int) (h->io.k_in++ & 1u);
    CUDA_TRY(h, cudaStreamWaitEvent(h->io.s_in, h->io.in_free[b], 0));  // the import that last read this buffer is done
    float* p = h->io.in[b];
    cudaError_t e = cudaSuccess;
    auto up = [&](const float* src, size_t cnt) -> float* {
        if (!src) return nullptr;
        float* dst = p;
        p += cnt;
        cudaError_t ee = cudaMemcpyAsync(dst, src, cnt * sizeof(float), cudaMemcpyHostToDevice, h->io.s_in);
        if (ee != cudaSuccess) e = ee;
        return dst;
    };
    float* dx = up(x, n * D);
    float* dv = up(v, n * D);
    float* dF = up(F, n * D * D);
    float* dC = up(C, n * D * D);
    float* dJ = up(Jp, n);
    CUDA_TRY(h, e);
    CUDA_TRY(h, cudaEventRecord(h->io.in_ready[b], h->io.s_in));
    CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->io.in_ready[b], 0));
    ParticleStore& S = h->store[h->cur];
    ParticleStore& T = h->store[h->cur ^ 1];
    reset_grids_after_error(h);
    discard_p2g_ahead(h);
    if (h->phase_next != 0) clear_grid(h, h->grid, h->box_cur);  // see nmpm_upload_particles
    NMPM_DISPATCH_DIM(h, (k_restore_constants<D><<<blocks_for(n, 256), 256, 0, h->stream>>>(S, T, (uint32_t) n)));
    NMPM_DISPATCH_DIM(h, (k_import_soa<D><<<blocks_for(n, 256), 256, 0, h->stream>>>(dx, dv, dF, dC, dJ, nullptr, nullptr,
                                                                                     (uint32_t) n, T, 1)));
    h->launches += 2;
    CUDA_TRY(h, cudaEventRecord(h->io.in_free[b], h->stream));
    h->cur ^= 1;
    h->phase_next = 0;
    h->keys_valid = false;
    h->box_valid = false;
    h->perm = nullptr;
    h->steps_done = 0;
    // an error latched by earlier steps stays latched (no host wait here); the device flag starts clean for the new state
    CUDA_TRY(h, cudaMemsetAsync(h->d_error, 0, 2 * sizeof(int), h->stream));
    CUDA_TRY(h, cudaGetLastError());
    return NMPM_OK;
}

// Pipelined form of nmpm_download_particles: the state as of the steps enqueued so far, in input order; the host arrays
// are valid after nmpm_synchronize (or once two more async downloads have completed).
int nmpm_download_particles_async(nmpm_handle h, float* x, float* v, float* F, float* C, float* Jp) {
    if (!h) return NMPM_ERR_INVALID;
    if (int rc = not_for_slabs(h, "nmpm_download_particles_async")) return rc;
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t n = h->n, D = (size_t) h->dim;
    if (n == 0) return NMPM_OK;
    if (int rc = async_io_init(h)) return rc;
    const int b = (int) (h->io.k_out++ & 1u);
    CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->io.out_free[b], 0));  // the D2H copy that last read this buffer is done
    float* p = h->io.out[b];
    float* dx = x ? p : nullptr;
    p += x ? n * D : 0;
    float* dv = v ? p : nullptr;
    p += v ? n * D : 0;
    float* dF = F ? p : nullptr;
    p += F ? n * D * D : 0;
    float* dC = C ? p : nullptr;
    p += C ? n * D * D : 0;
    float* dJ = Jp ? p : nullptr;
    NMPM_DISPATCH_DIM(h, (k_export_soa<D><<<blocks_for(n, 256), 256, 0, h->stream>>>(h->store[h->cur], (uint32_t) n, dx, dv, dF, dC,
                                                                                     dJ, nullptr, nullptr)));
    h->launches++;
    CUDA_TRY(h, cudaEventRecord(h->io.out_ready[b], h->stream));
    CUDA_TRY(h, cudaStreamWaitEvent(h->io.s_out, h->io.out_ready[b], 0));
    if (x) CUDA_TRY(h, cudaMemcpyAsync(x, dx, n * D * sizeof(float), cudaMemcpyDeviceToHost, h->io.s_out));
    if (v) CUDA_TRY(h, cudaMemcpyAsync(v, dv, n * D * sizeof(float), cudaMemcpyDeviceToHost, h->io.s_out));
    if (F) CUDA_TRY(h, cudaMemcpyAsync(F, dF, n * D * D * sizeof(float), cudaMemcpyDeviceToHost, h->io.s_out));
    if (C) CUDA_TRY(h, cudaMemcpyAsync(C, dC, n * D * D * sizeof(float), cudaMemcpyDeviceToHost, h->io.s_out));
    if (Jp) CUDA_TRY(h, cudaMemcpyAsync(Jp, dJ, n * sizeof(float), cudaMemcpyDeviceToHost, h->io.s_out));
    CUDA_TRY(h, cudaEventRecord(h->io.out_free[b], h->io.s_out));
    CUDA_TRY(h, cudaGetLastError());
    return NMPM_OK;
}

int nmpm_sort_debug(nmpm_handle h, int32_t* base, uint32_t* keys_unsorted, uint32_t* keys_sorted, uint32_t* perm,
                    uint32_t* ids) {
    if (!h) return NMPM_ERR_INVALID;
    if (h->phase_next != 0) {
        h->last_error = "nmpm_sort_debug: not allowed in the middle of a step";
        return NMPM_ERR_INVALID;
    }
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t n = h->n, D = (size_t) h->dim;
    if (n == 0) return NMPM_OK;
    h->keys_valid = false;  // the sort below consumes the key buffer
    if (int rc = ensure_staging(h, n * D * sizeof(int32_t))) return rc;
    ParticleStore& S = h->store[h->cur];
    int32_t* dbase = (int32_t*) h->staging;
    NMPM_DISPATCH_DIM(h, (k_cell_keys<D><<<blocks_for(n, 256), 256, 0, h->stream>>>(
                             S, (uint32_t) n, h->P, h->tiles_per_axis, h->sort.keys_a, dbase, h->d_error, nullptr)));
    h->launches++;
    if (base) CUDA_TRY(h, cudaMemcpyAsync(base, dbase, n * D * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    if (keys_unsorted)
        CUDA_TRY(h, cudaMemcpyAsync(keys_unsorted, h->sort.keys_a, n * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                                    h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    uint32_t *ks = nullptr, *pm = nullptr;
    h->launches += radix_sort_pairs(h->sort, (uint32_t) n, h->key_bits, h->stream, &ks, &pm);
    if (keys_sorted)
        CUDA_TRY(h, cudaMemcpyAsync(keys_sorted, ks, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    if (perm) CUDA_TRY(h, cudaMemcpyAsync(perm, pm, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    if (ids) CUDA_TRY(h, cudaMemcpyAsync(ids, S.id, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    CUDA_TRY(h, cudaGetLastError());
    return NMPM_OK;
}

int nmpm_affine_debug(nmpm_handle h, float* A_out) {
    if (!h || !A_out) return NMPM_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t n = h->n, D = (size_t) h->dim;
    if (n == 0) return NMPM_OK;
    if (int rc = ensure_staging(h, n * D * D * sizeof(float))) return rc;
    NMPM_DISPATCH(h, (k_affine_debug<D, MODEL><<<blocks_for(n, 128), 128, 0, h->stream>>>(h->store[h->cur], (uint32_t) n,
                                                                                         h->P, h->staging)));
    h->launches++;
    CUDA_TRY(h, cudaMemcpyAsync(A_out, h->staging, n * D * D * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    CUDA_TRY(h, cudaGetLastError());
    return NMPM_OK;
}

static int svd_polar_batch(int dim, size_t count, const float* A, float* U, float* sig, float* V, float* R, int device,
                           float* G = nullptr, float lo = 0.0f, float hi = 0.0f) {
    if ((dim != 2 && dim != 3) || (count && !A)) return NMPM_ERR_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        g_create_error = "no CUDA device: libnmpm has no CPU fallback";
        return NMPM_ERR_NO_DEVICE;
    }
    if (count == 0) return NMPM_OK;
    nmpm_sim* h = nullptr;
    CUDA_TRY(h, cudaSetDevice(device));
    const size_t w = count * (size_t) (dim * dim);
    float* d = nullptr;
    CUDA_TRY(h, cudaMalloc(&d, 6 * w * sizeof(float)));
    cudaError_t e = cudaMemcpy(d, A, w * sizeof(float), cudaMemcpyHostToDevice);
    float *dU = U ? d + w : nullptr, *dS = d + 2 * w, *dV = d + 3 * w, *dR = R ? d + 4 * w : nullptr,
          *dG = G ? d + 5 * w : nullptr;
    if (e == cudaSuccess) {
        if (dim == 2) k_svd_batch<2><<<blocks_for(count, 128), 128>>>(d, count, dU, dS, dV, dR, dG, lo, hi);
        else
            k_svd_batch<3><<<blocks_for(count, 128), 128>>>(d, count, dU, dS, dV, dR, dG, lo, hi);
        e = cudaDeviceSynchronize();
    }
    if (e == cudaSuccess && U) e = cudaMemcpy(U, dU, w * sizeof(float), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && sig) e = cudaMemcpy(sig, dS, w * sizeof(float), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && V) e = cudaMemcpy(V, dV, w * sizeof(float), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && R) e = cudaMemcpy(R, dR, w * sizeof(float), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && G) e = cudaMemcpy(G, dG, w * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(d);
    CUDA_TRY(h, e);
    return NMPM_OK;
}

int nmpm_svd_batch(int dim, size_t count, const float* A, float* U, float* sig, float* V, int device) {
    if (!U || !sig || !V) return NMPM_ERR_INVALID;
    return svd_polar_batch(dim, count, A, U, sig, V, nullptr, device);
}
int nmpm_polar_batch(int dim, size_t count, const float* A, float* R, int device) {
    if (!R) return NMPM_ERR_INVALID;
    return svd_polar_batch(dim, count, A, nullptr, nullptr, nullptr, R, device);
}

int nmpm_snow_project_batch(int dim, size_t count, const float* A, float lo, float hi, float* G, int device) {
    if (!G) return NMPM_ERR_INVALID;
    return svd_polar_batch(dim, count, A, nullptr, nullptr, nullptr, nullptr, device, G, lo, hi);
}

int nmpm_timing_enable(nmpm_handle h, int on) {
    if (!h) return NMPM_ERR_INVALID;
    h->timing = on != 0;
    return NMPM_OK;
}
int nmpm_timing_read(nmpm_handle h, float* ms, int* steps, int reset) {
    if (!h) return NMPM_ERR_INVALID;
    if (ms) std::memcpy(ms, h->t_ms, sizeof(h->t_ms));
    if (steps) *steps = h->t_steps;
    if (reset) {
        std::memset(h->t_ms, 0, sizeof(h->t_ms));
        h->t_steps = 0;
    }
    return NMPM_OK;
}


// ---- multi-GPU slab plumbing -------------------------------------------------
size_t nmpm_grid_plane_bytes(nmpm_handle h) {
    if (!h) return 0;
    const size_t n1 = (size_t) h->res + 1;
    return (h->dim == 3 ? n1 * n1 : n1) * sizeof(float4);
}
void* nmpm_grid_plane_ptr(nmpm_handle h, int x_plane) {
    if (!h || x_plane < 0 || x_plane > h->res) return nullptr;
    return (void*) ((char*) h->grid + (size_t) x_plane * nmpm_grid_plane_bytes(h));
}
int nmpm_grid_add_planes(nmpm_handle h, int x_plane, int planes, const void* device_src) {
    if (!h || !device_src || planes < 0 || x_plane < 0 || x_plane + planes > h->res + 1) return NMPM_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t count = (size_t) planes * nmpm_grid_plane_bytes(h) / sizeof(float4);
    if (count == 0) return NMPM_OK;
    k_add_planes<<<blocks_for(count, 256), 256, 0, h->stream>>>((float4*) nmpm_grid_plane_ptr(h, x_plane),
                                                              (const float4*) device_src, count);
    h->launches++;
    h->dirty_planes.emplace_back(x_plane, planes);  // cleared wholesale at the next step (beyond the node box)
    CUDA_TRY(h, cudaGetLastError());
    return NMPM_OK;
}
size_t nmpm_migrate_record_bytes(nmpm_handle h) {
    if (!h) return 0;
    return (size_t) (2 * h->dim + 2 * h->dim * h->dim + 4) * sizeof(float);
}

static int slab_check(nmpm_sim* h, const char* what) {
    if (!h) return NMPM_ERR_INVALID;
    if (!h->slab) {
        h->last_error = std::string(what) + ": the sim was not created as a slab (nmpm_options.slab_x1 == 0)";
        return NMPM_ERR_INVALID;
    }
    return NMPM_OK;
}

int nmpm_slab_p2g(nmpm_handle h) {
    if (int rc = slab_check(h, "nmpm_slab_p2g")) return rc;
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (int rc = poll_error(h)) return rc;
    if (h->phase_next != 0) {
        h->last_error = "nmpm_slab_p2g: called in the middle of a step";
        return NMPM_ERR_INVALID;
    }
    if (h->timing) cudaEventRecord(h->ev[0], h->stream);
    if (h->steps_done % h->opt.sort_every == 0 || !h->keys_valid) {
        if (int rc = do_sort(h)) return rc;
    } else {
        h->perm = nullptr;
        h->n = h->n_store;  // in place: every slot, the gone ones are skipped on the device
    }
    if (h->timing) cudaEventRecord(h->ev[1], h->stream);
    if (int rc = do_p2g(h)) return rc;
    if (h->timing) cudaEventRecord(h->ev[2], h->stream);
    h->phase_next = 1;
    CUDA_TRY(h, cudaGetLastError());
    return NMPM_OK;
}

}  // extern "C"

extern "C" {

int nmpm_slab_grid_g2p(nmpm_handle h, void* send_left, void* send_right, size_t cap_records, int* d_counts) {
    if (int rc = slab_check(h, "nmpm_slab_grid_g2p")) return rc;
    if (!send_left || !send_right || !d_counts || cap_records > 0xFFFFFFF0ull) return NMPM_ERR_INVALID;
    if (h->phase_next != 1) {
        h->last_error = "nmpm_slab_grid_g2p: nmpm_slab_p2g must come first";
        return NMPM_ERR_INVALID;
    }
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaMemsetAsync(d_counts, 0, 4 * sizeof(int), h->stream));
    if (h->timing) cudaEventRecord(h->ev[5], h->stream);
    if (int rc = do_grid_op(h)) return rc;
    if (h->timing) cudaEventRecord(h->ev[3], h->stream);
    MigrateArgs mig{h->opt.slab_x0, h->opt.slab_x1, (float*) send_left, (float*) send_right, (uint32_t) cap_records,
                    d_counts, (uint32_t) cap_records};
    if (int rc = do_g2p(h, mig)) return rc;
    if (h->timing) {  // per-phase device times of this slab (the ghost exchange between P2G and grid_op is not included)
        cudaEventRecord(h->ev[4], h->stream);
        cudaEventSynchronize(h->ev[4]);
        float ms = 0;
        cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
        h->t_ms[NMPM_T_SORT] += ms;
        cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]);
        h->t_ms[NMPM_T_P2G] += ms;
        cudaEventElapsedTime(&ms, h->ev[5], h->ev[3]);
        h->t_ms[NMPM_T_GRID] += ms;
        cudaEventElapsedTime(&ms, h->ev[3], h->ev[4]);
        h->t_ms[NMPM_T_G2P] += ms;
        h->t_steps++;
    }
    h->steps_done++;
    h->phase_next = 0;
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaMemcpyAsync(h->h_error, h->d_error, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    return NMPM_OK;
}

int nmpm_slab_unpack(nmpm_handle h, const void* recv_left, size_t n_from_left, const void* recv_right,
                     size_t n_from_right, size_t n_sent) {
    if (int rc = slab_check(h, "nmpm_slab_unpack")) return rc;
    if ((n_from_left && !recv_left) || (n_from_right && !recv_right) || n_sent > h->n_store) return NMPM_ERR_INVALID;
    if (h->phase_next != 0) {
        h->last_error = "nmpm_slab_unpack: must follow nmpm_slab_grid_g2p";
        return NMPM_ERR_INVALID;
    }
    if (h->n_store + n_from_left + n_from_right > h->cap) {
        h->last_error = "nmpm_slab_unpack: particle capacity exceeded (raise nmpm_options.capacity)";
        return NMPM_ERR_INVALID;
    }
    CUDA_TRY(h, cudaSetDevice(h->device));
    ParticleStore& T = h->store[h->cur];
    const void* src[2] = {recv_left, recv_right};
    const size_t cnt[2] = {n_from_left, n_from_right};
    for (int s = 0; s < 2; ++s) {
        if (!cnt[s]) continue;
        NMPM_DISPATCH_DIM(h, (k_unpack_records<D><<<blocks_for(cnt[s], 256), 256, 0, h->stream>>>(
                                 (const float*) src[s], (uint32_t) cnt[s], (uint32_t) h->n_store, T, h->P,
                                 h->tiles_per_axis, h->sort.keys_a, h->d_box + h->box_cur)));
        h->launches++;
        h->n_store += cnt[s];
    }
    h->n_gone += n_sent;  // they stay in the store (marked in the key array) until the next sort drops them
    CUDA_TRY(h, cudaGetLastError());
    return NMPM_OK;
}

int nmpm_slab_set_range(nmpm_handle h, int slab_x0, int slab_x1) {
    if (int rc = slab_check(h, "nmpm_slab_set_range")) return rc;
    if (slab_x0 < 0 || slab_x1 <= slab_x0 || h->phase_next != 0) return NMPM_ERR_INVALID;
    // particles keep their current owner until the next G2P hands them over (the node box follows the
    // particles, not the ownership range, so nothing else changes)
    h->opt.slab_x0 = slab_x0;
    h->opt.slab_x1 = slab_x1;
    return NMPM_OK;
}

int nmpm_slab_histogram(nmpm_handle h, int* d_hist) {
    if (!h || !d_hist) return NMPM_ERR_INVALID;
    if (h->phase_next != 0) {
        h->last_error = "nmpm_slab_histogram: call between steps";
        return NMPM_ERR_INVALID;
    }
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (h->n_store == 0) return NMPM_OK;
    NMPM_DISPATCH_DIM(h, (k_histogram_x<D><<<blocks_for(h->n_store, 256), 256, 0, h->stream>>>(
                             h->store[h->cur], (uint32_t) h->n_store, h->P, d_hist,
                             (h->n_gone || h->dev_counts) ? h->sort.keys_a : nullptr)));
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return NMPM_OK;
}

int nmpm_set_ids(nmpm_handle h, const uint32_t* ids) {
    if (!h || (h->n && !ids)) return NMPM_ERR_INVALID;
    if (h->steps_done != 0 || h->n_gone != 0 || h->n_store != h->n) {
        h->last_error = "nmpm_set_ids: only right after creation";
        return NMPM_ERR_INVALID;
    }
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (h->n == 0) return NMPM_OK;
    if (int rc = ensure_staging(h, h->n * sizeof(uint32_t))) return rc;
    CUDA_TRY(h, cudaMemcpyAsync(h->staging, ids, h->n * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    NMPM_DISPATCH_DIM(h, (k_set_ids<D><<<blocks_for(h->n, 256), 256, 0, h->stream>>>(h->store[h->cur], (uint32_t) h->n,
                                                                                     (const uint32_t*) h->staging)));
    h->launches++;
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return NMPM_OK;
}

int nmpm_download_particles_slots(nmpm_handle h, float* x, float* v, float* F, float* C, float* Jp, uint32_t* ids) {
    if (!h || !ids) return NMPM_ERR_INVALID;
    if (h->phase_next != 0) {
        h->last_error = "nmpm_download_particles_slots: not in the middle of a step";
        return NMPM_ERR_INVALID;
    }
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t n = h->n_store, D = (size_t) h->dim;
    if (n) {
        const size_t words = n * (2 * D + 2 * D * D + 2);
        if (int rc = ensure_staging(h, words * sizeof(float))) return rc;
        float* p = h->staging;
        float* dx = p;
        p += n * D;
        float* dv = p;
        p += n * D;
        float* dF = p;
        p += n * D * D;
        float* dC = p;
        p += n * D * D;
        float* dJ = p;
        p += n;
        uint32_t* did = (uint32_t*) p;
        NMPM_DISPATCH_DIM(h, (k_export_soa<D><<<blocks_for(n, 256), 256, 0, h->stream>>>(h->store[h->cur], (uint32_t) n,
                                                                                         dx, dv, dF, dC, dJ, did,
                                                                                         (h->n_gone || h->dev_counts) ? h->sort.keys_a : nullptr)));
        h->launches++;
        if (x) CUDA_TRY(h, cudaMemcpyAsync(x, dx, n * D * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        if (v) CUDA_TRY(h, cudaMemcpyAsync(v, dv, n * D * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        if (F) CUDA_TRY(h, cudaMemcpyAsync(F, dF, n * D * D * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        if (C) CUDA_TRY(h, cudaMemcpyAsync(C, dC, n * D * D * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        if (Jp) CUDA_TRY(h, cudaMemcpyAsync(Jp, dJ, n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaMemcpyAsync(ids, did, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    }
    return sync_and_check(h);
}

}  // extern "C"

#include "nmpm_slab_comm.inl"
