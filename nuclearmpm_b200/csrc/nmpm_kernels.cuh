// Step kernels (sm_100a): binning keys, reorder, P2G, grid update, G2P, import/export.
// Reference semantics: src/nclr.h:104-165 (p2g), :263-310 (grid_op), :167-261 (g2p).
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#include "nmpm_math.cuh"
#include "nmpm_store.cuh"

// occupancy targets (CTAs of 128 threads per SM) of the two particle kernels; see DESIGN.md §5
#ifndef NMPM_G2P_MINB
#define NMPM_G2P_MINB 8
#endif
#ifndef NMPM_P2G_MINB
#define NMPM_P2G_MINB 8
#endif
#ifndef NMPM_G2P_PIPE_MINB
#define NMPM_G2P_PIPE_MINB 6   // persistent pipelined G2P: 80 registers (the window posting sits in the middle of the live state)
#endif

namespace nmpm {

// Grid node = float4 {momentum/velocity xyz, mass} (2D: {x, y, mass, 0}); dense (res+1)^dim array in
// the reference's index order, x slowest (src/nclr.h:141-142,152).  One float4 per node makes the
// P2G scatter a single vector reduction (RED.E.ADD.F32x4) and the G2P gather a single LDG.128.
__device__ __forceinline__ float2 splat2g(float a) { return make_float2(a, a); }

template <int D>
__device__ __forceinline__ float4 node_pack(const float (&mom)[D], float m) {
    if constexpr (D == 3) return make_float4(mom[0], mom[1], mom[2], m);
    return make_float4(mom[0], mom[1], m, 0.0f);
}

__device__ __forceinline__ void red_add_f32x4(float4* addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

// Keys above every valid cell key.  A particle whose stencil left the grid (the step is flagged, Q5)
// sorts after all valid particles; a particle that migrated to a neighbouring slab sorts last of all,
// so that the first n_live entries of the sorted order are exactly the particles that are still here.
constexpr uint32_t kKeyOutOfGrid = 0xFFFFFFFEu;
constexpr uint32_t kKeyGone = 0xFFFFFFFFu;

// Slab migration (multi-GPU, include/nmpm.h): G2P hands particles whose next base.x is outside
// [x0, x1) to the left / right send buffer as Particle<dim> AoS records.  left == nullptr: disabled.
struct MigrateArgs {
    int x0, x1;
    float* left;
    float* right;
    uint32_t cap;     // records the left buffer holds
    int* counts;      // {n_left, n_right, n_kept, overflow}
    uint32_t cap_right = 0;  // records the right buffer holds
};

template <int D>
struct RecordTraits {
    static constexpr int WORDS = 2 * D + 2 * D * D + 4;  // 16 (2D) / 28 (3D): sizeof(Particle<dim>) / 4
};

template <int D>
__device__ __forceinline__ void write_record(float* __restrict__ r, const PState<D>& p, float2 mv, uint32_t id) {
    float w[RecordTraits<D>::WORDS];
#pragma unroll
    for (int d = 0; d < D; ++d) w[d] = p.x[d], w[D + d] = p.v[d];
#pragma unroll
    for (int k = 0; k < D * D; ++k) w[2 * D + k] = p.F.m[k], w[2 * D + D * D + k] = p.C.m[k];
    w[2 * D + 2 * D * D] = p.Jp;
    w[2 * D + 2 * D * D + 1] = mv.x;
    w[2 * D + 2 * D * D + 2] = mv.y;
    w[2 * D + 2 * D * D + 3] = __uint_as_float(id);
    float4* r4 = reinterpret_cast<float4*>(r);
#pragma unroll
    for (int k = 0; k < RecordTraits<D>::WORDS / 4; ++k) r4[k] = make_float4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
}

// Bounding box of the particles' stencil bases (inclusive, base coordinates).  P2G only writes nodes
// [lo, hi+2] per axis, so the grid clear and the grid update run over that box instead of the dense
// (res+1)^dim array (cfg4: 1.6 % of the 513^3 nodes are ever touched while the cube is compact).  Built on
// the device by whoever produces the next step's positions (key pass, G2P, slab unpack); never read by
// the host.  Empty box: lo = INT_MAX, hi = INT_MIN.
struct GridBox {
    int lo[3];
    int hi[3];
    int pad[2];
};

template <int D>
__device__ __forceinline__ void box_update(GridBox* __restrict__ box, const int (&b)[D], bool valid) {
    const unsigned mask = __activemask();
    const int leader = __ffs(mask) - 1;
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const int mn = __reduce_min_sync(mask, valid ? b[d] : 0x7fffffff);
        const int mx = __reduce_max_sync(mask, valid ? b[d] : (int) 0x80000000);
        if (lane == leader) {
            // monotone bounds: a stale read can only cause a redundant atomic, never a missed one
            if (mn < *((volatile int*) &box->lo[d])) atomicMin(&box->lo[d], mn);
            if (mx > *((volatile int*) &box->hi[d])) atomicMax(&box->hi[d], mx);
        }
    }
}

// G2P flavour: no atomics and no shared address at all — every warp leaves its own partial box
// (8 ints, two 16-byte stores) and k_box_reduce folds the partials afterwards.  (Guarded atomics on one
// GridBox cost +40 % G2P time: half a million same-address L2 round trips at the tail of every warp.)
// `live` = ballot of the lanes that own a particle, taken before any lane left the kernel.
template <int D>
__device__ __forceinline__ void box_partial_write(int* __restrict__ partial, unsigned live, const int (&b)[D], bool valid,
                                                  uint32_t warp_slot) {
    __syncwarp(live);
    int mn[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, mx[3] = {(int) 0x80000000, (int) 0x80000000, (int) 0x80000000};
#pragma unroll
    for (int d = 0; d < D; ++d) {
        mn[d] = __reduce_min_sync(live, valid ? b[d] : 0x7fffffff);
        mx[d] = __reduce_max_sync(live, valid ? b[d] : (int) 0x80000000);
    }
    if ((threadIdx.x & 31) == __ffs(live) - 1) {
        int4* out = reinterpret_cast<int4*>(partial + (size_t) warp_slot * 8);
        out[0] = make_int4(mn[0], mn[1], mn[2], mx[0]);
        out[1] = make_int4(mx[1], mx[2], 0, 0);
    }
}

// folds `count` per-warp partial boxes into *box (which must have been reset)
__global__ void __launch_bounds__(256) k_box_reduce(const int* __restrict__ partial, uint32_t count, GridBox* __restrict__ box) {
    int mn[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, mx[3] = {(int) 0x80000000, (int) 0x80000000, (int) 0x80000000};
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < count; w += gridDim.x * blockDim.x) {
        const int4 a = __ldg(reinterpret_cast<const int4*>(partial) + 2 * (size_t) w);
        const int4 c = __ldg(reinterpret_cast<const int4*>(partial) + 2 * (size_t) w + 1);
        mn[0] = min(mn[0], a.x), mn[1] = min(mn[1], a.y), mn[2] = min(mn[2], a.z);
        mx[0] = max(mx[0], a.w), mx[1] = max(mx[1], c.x), mx[2] = max(mx[2], c.y);
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        mn[d] = __reduce_min_sync(0xffffffffu, mn[d]);
        mx[d] = __reduce_max_sync(0xffffffffu, mx[d]);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            if (mn[d] != 0x7fffffff) atomicMin(&box->lo[d], mn[d]);
            if (mx[d] != (int) 0x80000000) atomicMax(&box->hi[d], mx[d]);
        }
    }
}

__global__ void k_box_reset(GridBox* __restrict__ box) {
    if (threadIdx.x < 3) {
        box->lo[threadIdx.x] = 0x7fffffff;
        box->hi[threadIdx.x] = (int) 0x80000000;
    }
}

// node box [lo, hi+2] clipped to the grid; returns the number of nodes (0 for an empty box)
// `n1x`: node rows along x (scenes * n1 for a batch of stacked 2D scenes; 0 = n1)
template <int D>
__device__ __forceinline__ uint32_t box_extent(const GridBox* __restrict__ box, int n1, int (&lo)[D], uint32_t (&ext)[D],
                                               int n1x = 0) {
    uint32_t vol = 1;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const int l = max(box->lo[d], 0);
        const int h = min(box->hi[d] + 2, ((d == 0 && n1x > 0) ? n1x : n1) - 1);
        if (h < l) return 0;
        lo[d] = l;
        ext[d] = (uint32_t) (h - l + 1);
        vol *= ext[d];
    }
    return vol;
}

template <int D>
__device__ __forceinline__ size_t box_node(uint32_t i, const int (&lo)[D], const uint32_t (&ext)[D], int n1, int (&c)[D]) {
    if constexpr (D == 3) {
        const uint32_t q = i / ext[2];
        c[2] = lo[2] + (int) (i - q * ext[2]);
        const uint32_t q2 = q / ext[1];
        c[1] = lo[1] + (int) (q - q2 * ext[1]);
        c[0] = lo[0] + (int) q2;
        return ((size_t) c[0] * n1 + c[1]) * n1 + c[2];
    } else {
        const uint32_t q = i / ext[1];
        c[1] = lo[1] + (int) (i - q * ext[1]);
        c[0] = lo[0] + (int) q;
        return (size_t) c[0] * n1 + c[1];
    }
}

// ---- K1: clear the nodes the previous P2G wrote (grid-stride over the previous box) ----------
template <int D>
__global__ void __launch_bounds__(256) k_clear_box(float4* __restrict__ grid, const GridBox* __restrict__ box, int n1,
                                                   int n1x = 0, const int* __restrict__ tiles_instead = nullptr) {
    if (tiles_instead && *tiles_instead) return;  // k_tiles3<0> clears these positions' nodes
    int lo[D], c[D];
    uint32_t ext[D];
    const uint32_t vol = box_extent<D>(box, n1, lo, ext, n1x);
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < vol; i += stride)
        grid[box_node<D>(i, lo, ext, n1, c)] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
}

// ---- active node tiles (3D, single GPU) --------------------------------------------------------------------------
// The node box above is exact for a compact block, but the 3D snow scenes turn into thin sheets and droplets that fill
// their bounding box very sparsely (cfg4 at step 300: the box holds 95 M nodes, 9 M of them non-zero, tools/
// grid_occupancy.py) while clear + grid_op stream over all of it.  Tile mode: the dense grid is cut into tiles of
// 4^3 nodes with one flag bit each; the flags of the tiles the stencils cover are raised together with the node box,
// grid_op and the clear visit only flagged tiles (the clear also lowers the flags).  The flag array of the whole 513^3
// grid is 268 KB, so it is simply scanned in full: no compaction, no box.
// A stencil with base b covers nodes b..b+2 per axis, i.e. the tile of b and, where (b & 3) >= 2, the next one.
__device__ __forceinline__ int node_tiles_per_axis(int n1) { return (n1 + 3) >> 2; }

// Raised from the cell KEYS of the coming step (G2P and the key pass write one per particle: key = tile << 6 | cell in
// tile, and the tile of a stencil's base node IS the cell tile), by a kernel of its own right after them.  Raising a
// flag is a test-then-set on a word that every warp of the same tile wants to touch — measured on cfg4 (gpurun
// r2u-r3c), extra time per step when it is done inside the particle kernels:
//   one flag BYTE per tile, plain stores from the P2G run flushes ............ P2G +0.34 ms  (partial-sector writes)
//   ... test through L1 before the store ...................................... P2G +0.27 ms
//   bytes, in G2P per warp and tile (match + OR of the corner sets) ........... G2P +0.54 ms  (4 M one-byte stores)
//   bits, test through L1 + RED.OR, per warp and tile (match) ................. G2P +0.11 ms, fused kernel +0.37 ms
//   bits, RED.OR without a test, lanes de-duplicated against their left neighbour only .. G2P +3.1 ms
//   bits, test at the L2 (ld.global.cg) + RED.OR, per warp and tile ........... fused +1.3 ms
// so here the global traffic is cut at the source: a warp walks 1 024 consecutive (cell-sorted) keys and collects
// {tile, corners} in a private 32-entry table in shared memory; what it holds at the end goes to memory once (test through
// L1, then a 32-bit RED.OR).  (x & 3) >= 2 is bit 1 of the 2-bit cell coordinate: key bits 5, 3, 1.
// The kernel streams 4 B per particle.
// Tile mode is decided on the DEVICE, per step, from the node box of the coming step's positions (so that a host that
// enqueues thousands of steps ahead still gets the switch at the right step): tiles pay when the box is mostly empty —
// wanted above 1.5 box nodes per particle (a compact 8-per-cell block has ~0.4, cfg4 at step 150 has 2.2 and 76 % of its
// box empty), dropped again below 1.0.  want[b] belongs to ring slot b like box[b] and the flag array b; every consumer
// (k_mark_tiles, k_tiles3 / k_clear_box, k_grid_op) reads it and returns at once if it is not its turn.
__global__ void k_tile_decide(const GridBox* __restrict__ box, uint32_t n, int n1, const int* __restrict__ want_prev,
                              int* __restrict__ want_out, int policy) {
    if (threadIdx.x != 0) return;
    int w = policy == 2 ? 1 : 0;
    if (policy == 0) {
        float vol = 1.0f;
        for (int d = 0; d < 3; ++d) {
            const int lo = max(box->lo[d], 0), hi = min(box->hi[d] + 2, n1 - 1);
            vol *= hi >= lo ? (float) (hi - lo + 1) : 0.0f;
        }
        const float per_particle = n ? vol / (float) n : 0.0f;
        w = *want_prev;
        if (!w && per_particle > 1.5f) w = 1;
        else if (w && per_particle < 1.0f)
            w = 0;
    }
    *want_out = w;
}

constexpr int kMarkWarps = 8, kMarkKeysPerWarp = 1024;
__device__ __forceinline__ void tile_raise(uint32_t* __restrict__ flags, uint32_t lo, unsigned corners, int T) {
    unsigned up = 0u;  // all tests first (through L1: a stale 0 only costs a redundant reduction), then the reductions
#pragma unroll
    for (int c = 0; c < 8; ++c)
        if ((corners >> c) & 1u) {
            const uint32_t t = lo + (uint32_t) (((c & 1) ? T * T : 0) + ((c & 2) ? T : 0) + ((c & 4) ? 1 : 0));
            up |= ((__ldca(flags + (t >> 5)) >> (t & 31u)) & 1u) << c;
        }
    corners &= ~up;
#pragma unroll
    for (int c = 0; c < 8; ++c)
        if ((corners >> c) & 1u) {
            const uint32_t t = lo + (uint32_t) (((c & 1) ? T * T : 0) + ((c & 2) ? T : 0) + ((c & 4) ? 1 : 0));
            atomicOr(flags + (t >> 5), 1u << (t & 31u));
        }
}
__global__ void __launch_bounds__(kMarkWarps * 32) k_mark_tiles(const uint32_t* __restrict__ keys, uint32_t n,
                                                                uint32_t* __restrict__ flags, int T,
                                                                const int* __restrict__ want) {
    if (*want == 0) return;  // the coming step works on its node box (k_tile_decide)
    // per warp: 32 entries {tile, corners to raise}; an entry is claimed by the first tile that hashes to it and never
    // evicted (a tile that finds its entry taken goes to memory directly), the warp flushes its entries once at the end —
    // so the walk itself waits for no global memory access except the key loads
    __shared__ unsigned int tag[kMarkWarps][32], mask[kMarkWarps][32];
    constexpr unsigned kEmpty = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    tag[warp][lane] = kEmpty, mask[warp][lane] = 0u;
    __syncwarp();
    const uint32_t first = (blockIdx.x * kMarkWarps + warp) * (uint32_t) kMarkKeysPerWarp;
    constexpr int B = 8;  // keys per lane in flight
    for (uint32_t base = first; base < first + kMarkKeysPerWarp && base < n; base += 32u * B) {
        uint32_t kk[B];
#pragma unroll
        for (int j = 0; j < B; ++j) {
            const uint32_t i = base + (uint32_t) (j * 32 + lane);
            kk[j] = (i < n) ? __ldg(keys + i) : 0xFFFFFFFFu;
        }
#pragma unroll
        for (int j = 0; j < B; ++j) {
            const uint32_t key = kk[j];
            const bool valid = key < 0xFFFFFFFEu;  // not "out of grid" / "migrated away"
            const uint32_t lo = key >> 6;
            // a lane whose left neighbour starts in the same tile with the same corner set has nothing to add
            const uint32_t pkey = __shfl_up_sync(0xffffffffu, key, 1);
            if (valid && (lane == 0 || (pkey >> 6) != lo || ((pkey ^ key) & 0x2Au) != 0u)) {
                unsigned cm = 1u;  // bit c: corner tile c = (c&1 ? +x) (c&2 ? +y) (c&4 ? +z) is covered by this stencil
                if (key & 0x20u) cm |= cm << 1;  // (x & 3) >= 2
                if (key & 0x08u) cm |= cm << 2;  // (y & 3) >= 2
                if (key & 0x02u) cm |= cm << 4;  // (z & 3) >= 2
                const unsigned slot = (lo ^ (lo >> 5)) & 31u;
                const unsigned seen = atomicCAS(&tag[warp][slot], kEmpty, lo);
                if (seen == kEmpty || seen == lo) atomicOr(&mask[warp][slot], cm);
                else
                    tile_raise(flags, lo, cm, T);
            }
        }
    }
    __syncwarp();
    const unsigned mine = mask[warp][lane];
    if (mine) tile_raise(flags, tag[warp][lane], mine, T);
}

// One warp per flag word (32 consecutive tiles): the warp visits every flagged tile — 16 (x,y) rows of 4 z-contiguous
// nodes, a lane pair per row, two nodes per lane (z and z + 2).  OP 0: zero the nodes and lower the flags; OP 1: grid_op.
template <int D>
__device__ __forceinline__ bool grid_op_value(float4& g, const int (&c)[D], const MaterialParams& P);
template <int OP, int U = 4>  // U: flagged tiles in flight per warp (their loads are issued together)
__global__ void __launch_bounds__(256) k_tiles3(float4* __restrict__ grid, uint32_t* __restrict__ flags, MaterialParams P,
                                                const int* __restrict__ want) {
    if (*want == 0) return;  // these positions were not flagged: the node-box kernel does the work
    const int n1 = P.n1, T = node_tiles_per_axis(n1);
    const uint32_t nwords = ((uint32_t) (T * T * T) + 31u) >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    // lane -> (row, half): a lane pair covers two z-adjacent nodes, so that every load / store instruction of the warp
    // moves whole 32-byte sectors (half-filled sectors are read-modify-write at the L2)
    const int row = lane >> 1, half = lane & 1;
    for (uint32_t w = warp; w < nwords; w += nwarps) {
        unsigned active = flags[w];
        if (active == 0u) continue;
        if (OP == 0 && lane == 0) flags[w] = 0u;
        const uint32_t t0 = w << 5;
        while (active) {
            float4* node[U];
            int c[U][3];
            bool ok[U][2];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                node[u] = nullptr;
                ok[u][0] = ok[u][1] = false;
                if (active) {
                    const uint32_t t = t0 + (uint32_t) (__ffs(active) - 1);
                    active &= active - 1u;
                    const uint32_t q = t / (uint32_t) T;
                    const int tz = (int) (t - q * (uint32_t) T), ty = (int) (q % (uint32_t) T), tx = (int) (q / (uint32_t) T);
                    c[u][0] = 4 * tx + (row >> 2), c[u][1] = 4 * ty + (row & 3), c[u][2] = 4 * tz + half;  // and z + 2
                    if (c[u][0] < n1 && c[u][1] < n1) {
                        node[u] = grid + ((size_t) (c[u][0] * n1 + c[u][1]) * n1 + c[u][2]);
                        ok[u][0] = c[u][2] < n1, ok[u][1] = c[u][2] + 2 < n1;
                    }
                }
            }
            if constexpr (OP == 0) {
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int e = 0; e < 2; ++e)
                        if (ok[u][e]) node[u][2 * e] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            } else {
                float4 g[U][2];
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int e = 0; e < 2; ++e) g[u][e] = ok[u][e] ? node[u][2 * e] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        if (!ok[u][e]) continue;
                        const int cc[3] = {c[u][0], c[u][1], c[u][2] + 2 * e};
                        if (grid_op_value<3>(g[u][e], cc, P)) node[u][2 * e] = g[u][e];
                    }
            }
        }
    }
}

// Batch of stacked 2D scenes (MaterialParams::scenes): scene of the particle in slot `slot`, and the node-row offset
// that turns its base.x into the stacked grid's.  Single scene / 3D: 0.
template <int D>
__device__ __forceinline__ int scene_of_slot(const ParticleStore& S, uint32_t slot, const MaterialParams& P) {
    if constexpr (D == 2) {
        if (P.scenes > 1) return (int) P.scene_of[S.id[slot]];
    }
    return 0;
}
__device__ __forceinline__ MaterialParams scene_params(const MaterialParams& P, int scene) {
    MaterialParams Q = P;
    if (P.scenes > 1) {
        const float2 l = __ldg(P.lame + scene);
        Q.mu_0 = l.x, Q.lambda_0 = l.y;
    }
    return Q;
}

// ---- K0a: cell keys from current positions -------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) k_cell_keys(ParticleStore S, uint32_t n, MaterialParams P, int tiles_per_axis,
                                                   uint32_t* __restrict__ keys, int32_t* __restrict__ base_out,
                                                   int* __restrict__ error_flag, GridBox* __restrict__ box) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x[D];
    load_position<D>(S, i, x);
    int b[D];
    bool bad = false;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const Stencil1 s = stencil_axis(x[d], P.inv_dx, P.res);
        b[d] = s.base;
        if (base_out) base_out[(size_t) i * D + d] = b[d];
        if (!s.ok) bad = true;  // stencil left [0,res] (Q5) or x is not finite
    }
    const int xoff = scene_of_slot<D>(S, i, P) * P.n1;
    if (bad) {
        atomicOr(error_flag, 1);
        keys[i] = kKeyOutOfGrid;
    } else {
        b[0] += xoff;
        keys[i] = cell_key<D>(b, tiles_per_axis);
        b[0] -= xoff;
    }
    if (box) {  // flagged particles still scatter (to clamped nodes, see stencil_of): keep those nodes in the box
#pragma unroll
        for (int d = 0; d < D; ++d) b[d] = min(max(b[d], 0), P.res - 2);
        b[0] += xoff;
        box_update<D>(box, b, true);
    }
}

// ---- K0b: gather every particle array into sorted order -------------------------------------
template <int D>
__global__ void __launch_bounds__(256) k_reorder(ParticleStore src, ParticleStore dst, const uint32_t* __restrict__ perm,
                                                 uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t j = perm[i];
#pragma unroll
    for (int k = 0; k < StoreTraits<D>::NQ; ++k) dst.q[k][i] = ldg4(src.q[k] + j);
    dst.s[i] = __ldg(src.s + j);
    dst.mv[i] = __ldg(src.mv + j);
    dst.id[i] = __ldg(src.id + j);
}

// `xoff` (batch of stacked 2D scenes): node rows of the scenes before this particle's; added to base[0] AFTER the
// in-grid test and the clamp, which are per scene
template <int D>
__device__ __forceinline__ bool stencil_of(const float (&x)[D], const MaterialParams& P, int (&base)[D], float (&fx)[D],
                                           float (&w)[D][3], int xoff = 0) {
    bool ok = true;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const Stencil1 s = stencil_axis(x[d], P.inv_dx, P.res);
        base[d] = s.base;
        fx[d] = s.fx;
        w[d][0] = s.w[0], w[d][1] = s.w[1], w[d][2] = s.w[2];
        if (!s.ok) {
            ok = false;
            base[d] = min(max(s.base, 0), P.res - 2);  // keep every access in bounds; the step is flagged
        }
    }
    base[0] += xoff;
    return ok;
}

// ---- K2 (variant 1): one thread per particle, 3^D vector reductions --------------------------
template <int D, int MODEL>
__global__ void __launch_bounds__(128) k_p2g_scatter(ParticleStore S, const uint32_t* __restrict__ perm, uint32_t n,
                                                     MaterialParams P, float4* __restrict__ grid,
                                                     int* __restrict__ error_flag, const uint32_t* __restrict__ gone_keys,
                                                     uint32_t first) {
    // slots [first, n): `first` > 0 scatters only the tail of the store (slab mode: particles received after the
    // P2G of the resident ones was already issued)
    const uint32_t i = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (gone_keys && __ldg(gone_keys + i) == kKeyGone) return;  // slab mode between sorts: migrated away
    PState<D> p;
    load_for_p2g<D>(S, perm ? __ldg(perm + i) : i, p);
    int base[D];
    float fx[D], w[D][3];
    if (!stencil_of<D>(p.x, P, base, fx, w)) atomicOr(error_flag, 1);
    const Mat<D> A = affine_matrix<D, MODEL>(p.F, p.C, p.Jp, p.mass, p.volume, P);
    float mv[D];
#pragma unroll
    for (int d = 0; d < D; ++d) mv[d] = p.v[d] * p.mass;
    const int n1 = P.n1;
#pragma unroll
    for (int ii = 0; ii < 3; ++ii)
#pragma unroll
        for (int jj = 0; jj < 3; ++jj) {
            if constexpr (D == 3) {
#pragma unroll
                for (int kk = 0; kk < 3; ++kk) {
                    const float dpos[3] = {((float) ii - fx[0]) * P.dx, ((float) jj - fx[1]) * P.dx,
                                           ((float) kk - fx[2]) * P.dx};
                    const float weight = w[0][ii] * w[1][jj] * w[2][kk];
                    float mom[3];
#pragma unroll
                    for (int r = 0; r < 3; ++r)
                        mom[r] = weight * (mv[r] + (A(r, 0) * dpos[0] + A(r, 1) * dpos[1] + A(r, 2) * dpos[2]));
                    const size_t node = ((size_t) (base[0] + ii) * n1 + (base[1] + jj)) * n1 + (base[2] + kk);
                    red_add_f32x4(grid + node, node_pack<3>(mom, weight * p.mass));
                }
            } else {
                const float dpos[2] = {((float) ii - fx[0]) * P.dx, ((float) jj - fx[1]) * P.dx};
                const float weight = w[0][ii] * w[1][jj];
                float mom[2];
#pragma unroll
                for (int r = 0; r < 2; ++r) mom[r] = weight * (mv[r] + (A(r, 0) * dpos[0] + A(r, 1) * dpos[1]));
                const size_t node = (size_t) (base[0] + ii) * n1 + (base[1] + jj);
                red_add_f32x4(grid + node, node_pack<2>(mom, weight * p.mass));
            }
        }
}

// ---- K3: grid update (src/nclr.h:263-310) ---------------------------------------------------
// One thread per node, float4 in / float4 out.  Normalise by mass, gravity on y (Q7), clamp to
// ±0.9 dx/dt, then the sticky 3-node walls which zero the WHOLE node incl. its mass (Q6).
// the update of one node value; false: the node stays as it is (nothing to write)
template <int D>
__device__ __forceinline__ bool grid_op_value(float4& g, const int (&c)[D], const MaterialParams& P) {
    // untouched node (all zero): normalisation is skipped (mass == 0) and the sticky walls only act on
    // non-zero velocities, so the node stays as it is
    if (g.x == 0.0f && g.y == 0.0f && g.z == 0.0f && g.w == 0.0f) return false;
    float vel[D];
    float m;
    if constexpr (D == 3) {
        vel[0] = g.x, vel[1] = g.y, vel[2] = g.z, m = g.w;
    } else {
        vel[0] = g.x, vel[1] = g.y, m = g.z;
    }
    const int n1 = P.n1;
    if (m > 0.0f) {
#pragma unroll
        for (int d = 0; d < D; ++d) vel[d] = vel[d] / m;
        vel[1] += P.dt_gravity;
#pragma unroll
        for (int d = 0; d < D; ++d) vel[d] = clampf(vel[d], -P.vmax, P.vmax);
    }
#pragma unroll
    for (int d = 0; d < D; ++d) {
        int cd = c[d];
        if (D == 2 && d == 0 && P.scenes > 1) cd -= (cd / n1) * n1;  // stacked scenes: the walls are per scene
        if ((cd < 3 && vel[d] < 0.0f) || (cd >= n1 - 3 && vel[d] > 0.0f)) {
#pragma unroll
            for (int e = 0; e < D; ++e) vel[e] = 0.0f;
            m = 0.0f;
        }
    }
    g = node_pack<D>(vel, m);
    return true;
}
template <int D>
__device__ __forceinline__ void grid_op_node(float4* __restrict__ cell, const int (&c)[D], const MaterialParams& P) {
    float4 g = *cell;
    if (grid_op_value<D>(g, c, P)) *cell = g;
}

// grid-stride over the node box of the current particles (see GridBox)
template <int D>
__global__ void __launch_bounds__(256) k_grid_op(float4* __restrict__ grid, const GridBox* __restrict__ box,
                                                 MaterialParams P, const int* __restrict__ tiles_instead = nullptr) {
    if (tiles_instead && *tiles_instead) return;  // k_tiles3<1> updates these positions' nodes
    int lo[D], c[D];
    uint32_t ext[D];
    const uint32_t vol = box_extent<D>(box, P.n1, lo, ext, P.scenes > 1 ? P.scenes * P.n1 : 0);
    const uint32_t stride = gridDim.x * blockDim.x;
    constexpr int U = 4;  // nodes in flight per thread (the loop is a chain of dependent load -> store otherwise)
    for (uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < vol; i0 += U * stride) {
        float4* cell[U];
        float4 g[U];
        int cc[U][D];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t i = i0 + (uint32_t) u * stride;
            cell[u] = nullptr;
            if (i < vol) {
                cell[u] = grid + box_node<D>(i, lo, ext, P.n1, c);
#pragma unroll
                for (int d = 0; d < D; ++d) cc[u][d] = c[d];
                g[u] = *cell[u];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (cell[u] && grid_op_value<D>(g[u], cc[u], P)) *cell[u] = g[u];
    }
}

// ---- K4: G2P (src/nclr.h:167-261) -------------------------------------------------------------
template <int D, int MODEL>
__device__ __forceinline__ void g2p_update(PState<D>& p, const Mat<D>& Cn, const float (&vn)[D], const MaterialParams& P) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
        p.v[d] = vn[d];
        p.x[d] = fmaf(P.dt, vn[d], p.x[d]);  // advection (src/nclr.h:229)
    }
    p.C = Cn;
    // F' = (diag<dim>(1) + dt*C) * F   (Q1: the "identity" has a zero (2,2) entry in 3D)
    Mat<D> M;
#pragma unroll
    for (int k = 0; k < D * D; ++k) M.m[k] = P.dt * Cn.m[k];
    M(0, 0) += 1.0f;
    M(1, 1) += 1.0f;
    Mat<D> Fn = mat_mul<D>(M, p.F);
    if constexpr (MODEL == 1) {  // jelly (src/nclr.h:232-234)
        p.F = Fn;
    } else if constexpr (MODEL == 0) {  // snow plasticity (src/nclr.h:239-250)
        const float old_J = det(Fn);
        Fn = snow_project(Fn, 0.975f, 1.0045f);  // U clamp(sig) V^T
        p.Jp = clampf(p.Jp * old_J / det(Fn), 0.6f, 20.0f);
        p.F = Fn;
    } else {  // liquid (src/nclr.h:252-258): F = diag<dim>(1) with F(0,0) = J = prod(sig)
        // prod(sig) needs no SVD: in 3D the reference's sign fix makes det U = det V = +1, so
        // sig0 sig1 sig2 = det(F'); in 2D the fix is a no-op (Q3), all sig >= 0, so sig0 sig1 = |det F'|.
        // (The reference multiplies the float singular values in double, Q8; the difference to the
        // fp32 determinant is a few ulp of |F'|^dim, far inside the F tolerance.)
        float J = det(Fn);
        if constexpr (D == 2) J = fabsf(J);
#pragma unroll
        for (int k = 0; k < D * D; ++k) p.F.m[k] = 0.0f;
        p.F(1, 1) = 1.0f;
        p.F(0, 0) = J;
    }
}

// ---- async copy plumbing (sm_90+ PTX): mbarrier + TMA tensor loads into shared memory --------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");  // visible to the async proxy
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0u;
}
// one box of a rank-4 tensor map (coordinates fastest axis first) -> shared memory, completion on `bar`
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// 16-byte asynchronous global -> shared copies (LDGSTS), one commit group per thread
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Node window of the TMA-staged G2P: up to kWinX x-planes of kWinY x kWinZ nodes (float4), one TMA box per plane
// (tensor-map box {4 floats, kWinZ, kWinY, 1}); planes sit kWinPitch nodes apart (128-byte aligned for the TMA).
// kWinZ = 13 (odd): rows of a plane start 13 nodes apart, which spreads the 16-byte bank groups of neighbouring rows.
constexpr int kWinX = 8, kWinY = 12, kWinZ = 13;
constexpr int kWinPitch = 160;                           // >= kWinY * kWinZ = 156, multiple of 8 nodes (128 B)
constexpr uint32_t kWinPlaneBytes = kWinY * kWinZ * 16;  // bytes one TMA box delivers

// ---- K4 helpers --------------------------------------------------------------------------------
// Where the 27 stencil nodes of a particle come from: the dense grid in global memory (three adjacent float4
// nodes per (i,j) row), or a node window staged in shared memory by the TMA (nmpm_g2p_tile.cuh) whose strides
// are compile-time constants, so that all 27 loads are one base register plus an immediate offset.
struct GlobalNodes {
    const float4* gp;  // node (base.x, base.y, base.z)
    int plane, n1;
    __device__ __forceinline__ float4 load(int ii, int jj, int kk) const { return __ldg(gp + (ii * plane + jj * n1) + kk); }
};
template <int PLANE, int ROW>
struct WindowNodes {
    const float4* sp;  // node (base - window origin) of the shared-memory window
    __device__ __forceinline__ float4 load(int ii, int jj, int kk) const { return sp[ii * PLANE + jj * ROW + kk]; }
};

// Gather with the stencil offsets centred on the middle node:  o = ijk - 1 in {-1,0,+1},
//   v   = sum w g
//   B_c = sum w g o_c                  (the o_c = 0 terms vanish at compile time)
//   C   = 4 inv_dx (B - v (fx-1)^T)    == sum 4 inv_dx (w g) (ijk - fx)^T   (src/nclr.h:206,223)
// Sum factorisation: w = wx_i wy_j wz_k is separable, so the 27-node sums are three nested 3-term sums — along z
// per (i,j) row, along y per i, along x — and the weight products are never formed: 126 packed instructions
// instead of ~270.  Packed pairs: (x,y) components, and (plain, z-offset-weighted) sums of the z component.
template <class Nodes>
__device__ __forceinline__ void g2p_gather3(const Nodes& nodes, const float (&w)[3][3], const float (&fx)[3],
                                            float four_inv_dx, float (&vn)[3], Mat<3>& Cn) {
    const float2 wz0 = splat2g(w[2][0]), wz1 = splat2g(w[2][1]), wz2 = splat2g(w[2][2]), nwz0 = splat2g(-w[2][0]);
    const float2 wzp0 = make_float2(w[2][0], -w[2][0]), wzp1 = make_float2(w[2][1], 0.0f), wzp2 = wz2;
    float2 v01, vzBzz, Bz01, By01, Bx01;  // v.xy | (v.z, B_z.z) | B_z.xy | B_y.xy | B_x.xy
    float Byz, Bxz;
#pragma unroll
    for (int ii = 0; ii < 3; ++ii) {
        float2 t01, tzz, tz01, ty01;  // sums over (j,k) of this i: t.xy | (t.z, t_z.z) | t_z.xy | t_y.xy
        float tyz;
#pragma unroll
        for (int jj = 0; jj < 3; ++jj) {
            const float4 g0 = nodes.load(ii, jj, 0), g1 = nodes.load(ii, jj, 1), g2 = nodes.load(ii, jj, 2);
            float2 s01 = __fmul2_rn(wz0, make_float2(g0.x, g0.y));
            s01 = __ffma2_rn(wz1, make_float2(g1.x, g1.y), s01);
            s01 = __ffma2_rn(wz2, make_float2(g2.x, g2.y), s01);
            float2 sz01 = __fmul2_rn(wz2, make_float2(g2.x, g2.y));
            sz01 = __ffma2_rn(nwz0, make_float2(g0.x, g0.y), sz01);
            float2 szz = __fmul2_rn(wzp0, splat2g(g0.z));  // (sum wz g.z, sum wz o_z g.z)
            szz = __ffma2_rn(wzp1, splat2g(g1.z), szz);
            szz = __ffma2_rn(wzp2, splat2g(g2.z), szz);
            const float2 wy = splat2g(w[1][jj]);
            if (jj == 0) {
                t01 = __fmul2_rn(wy, s01), tzz = __fmul2_rn(wy, szz), tz01 = __fmul2_rn(wy, sz01);
                ty01 = __fmul2_rn(splat2g(-w[1][0]), s01);
                tyz = -w[1][0] * szz.x;
            } else {
                t01 = __ffma2_rn(wy, s01, t01), tzz = __ffma2_rn(wy, szz, tzz), tz01 = __ffma2_rn(wy, sz01, tz01);
                if (jj == 2) {
                    ty01 = __ffma2_rn(wy, s01, ty01);
                    tyz = fmaf(w[1][2], szz.x, tyz);
                }
            }
        }
        const float2 wx = splat2g(w[0][ii]);
        if (ii == 0) {
            v01 = __fmul2_rn(wx, t01), vzBzz = __fmul2_rn(wx, tzz), Bz01 = __fmul2_rn(wx, tz01);
            By01 = __fmul2_rn(wx, ty01), Byz = w[0][0] * tyz;
            Bx01 = __fmul2_rn(splat2g(-w[0][0]), t01), Bxz = -w[0][0] * tzz.x;
        } else {
            v01 = __ffma2_rn(wx, t01, v01), vzBzz = __ffma2_rn(wx, tzz, vzBzz), Bz01 = __ffma2_rn(wx, tz01, Bz01);
            By01 = __ffma2_rn(wx, ty01, By01), Byz = fmaf(w[0][ii], tyz, Byz);
            if (ii == 2) Bx01 = __ffma2_rn(wx, t01, Bx01), Bxz = fmaf(w[0][2], tzz.x, Bxz);
        }
    }
    vn[0] = v01.x, vn[1] = v01.y, vn[2] = vzBzz.x;
    const float Bc[3][3] = {{Bx01.x, Bx01.y, Bxz}, {By01.x, By01.y, Byz}, {Bz01.x, Bz01.y, vzBzz.y}};  // Bc[c][r]
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float fc = fx[c] - 1.0f;
#pragma unroll
        for (int r = 0; r < 3; ++r) Cn(r, c) = four_inv_dx * fmaf(-vn[r], fc, Bc[c][r]);
    }
}

// Everything after the gather: F/Jp/C update, the next step's cell key, the warp-local re-grouping, the stores,
// slab migration and the warp's partial node box.
template <int D, int MODEL>
__device__ __forceinline__ void g2p_finish(PState<D>& p, const Mat<D>& Cn, const float (&vn)[D], const ParticleStore& S,
                                           const ParticleStore& T, const uint32_t* __restrict__ perm, uint32_t src, uint32_t i,
                                           unsigned live, const MaterialParams& P, uint32_t* __restrict__ keys_out,
                                           int tiles_per_axis, const MigrateArgs& mig, int* __restrict__ box_partial,
                                           const uint32_t* __restrict__ gone_keys, int local_reorder, int xoff = 0) {
    g2p_update<D, MODEL>(p, Cn, vn, P);
    // warp-uniform (kernel argument).  Slabs re-group as well: ranks are mapped onto the slots of the warp's LIVE lanes
    // (slots whose particle migrated away keep their "gone" mark and stay where they are).
    const bool reorder = local_reorder != 0;
    {   // bin the advected particle for the NEXT step: cell key, slab owner, node box
        int b[D];
        bool bad = false;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const Stencil1 s = stencil_axis(p.x[d], P.inv_dx, P.res);
            b[d] = s.base;
            bad = bad || !s.ok;
        }
        b[0] += xoff;  // batch of stacked 2D scenes: node rows of the scenes before this one (else 0)
        // an out-of-grid position is flagged by the next step's P2G/G2P (that is when the reference throws)
        uint32_t key = bad ? kKeyOutOfGrid : cell_key<D>(b, tiles_per_axis);
        b[0] -= xoff;
        // Single GPU: the warp writes its 32 particles back GROUPED BY THEIR NEW CELL instead of slot by slot.  Between
        // two radix sorts the particles of a cell drift into 2-3 neighbouring cells and interleave (A B A A C B ...): P2G
        // then meets one run per fragment, and every run costs 27 lane-reductions.  Re-grouping inside the warp's own 32
        // slots needs no extra pass and no shared memory (on a re-binned step it is free of extra traffic as well).
        uint32_t dst = i;
        bool moved = false;
        // nothing to do for a warp whose keys are already non-decreasing along the lanes (calm scenes, fresh sorts)
        if (reorder) {
            const int lane = threadIdx.x & 31;
            const uint32_t prev = __shfl_up_sync(live, key, 1);
            moved = __any_sync(live, lane > 0 && ((live >> (lane - 1)) & 1u) && prev > key);
        }
        // mass/volume and id travel with the particle whenever it changes slot (re-binned write, migration, re-grouping).
        // (Loading them at the top of the kernel to hide their latency costs more in register pressure than it saves.)
        float2 mv = make_float2(0.0f, 0.0f);
        uint32_t pid = 0;
        if (perm || mig.left || moved) {
            mv = S.mv[src];   // coherent loads: the re-grouping below rewrites these arrays in place
            pid = S.id[src];
        }
        if (moved) {
            const int lane = threadIdx.x & 31;
            const unsigned peers = __match_any_sync(live, key);  // lanes whose particle lands in the same cell
            // cells in the order of their first particle: `before` = lanes whose cell's first lane precedes mine's,
            // counted with five ballots (one per bit of the 5-bit first-lane index, most significant first)
            const int leader = __ffs(peers) - 1;
            unsigned lt = 0u, eq = live;
#pragma unroll
            for (int bit = 4; bit >= 0; --bit) {
                const unsigned ones = __ballot_sync(live, (leader >> bit) & 1);
                if ((leader >> bit) & 1) {
                    lt |= eq & ~ones;
                    eq &= ones;
                } else {
                    eq &= ~ones;
                }
            }
            const int rank = __popc(lt) + __popc(peers & ((1u << lane) - 1u));
            // the rank-th live slot of the warp (all 32 lanes live: the rank itself)
            dst = (i - lane) + (uint32_t) ((live == 0xffffffffu) ? rank : (int) __fns(live, 0, rank + 1));
            __syncwarp(live);  // T == S: every lane has read its old slot (state, mass/volume, id) before any is overwritten
        }
        store_state<D>(T, dst, p);
        if (perm || moved) {
            T.mv[dst] = mv;
            T.id[dst] = pid;
        }
        bool gone = false;
        if (mig.left && !bad && (b[0] < mig.x0 || b[0] >= mig.x1)) {
            const int side = (b[0] < mig.x0) ? 0 : 1;
            const uint32_t slot = (uint32_t) atomicAdd(mig.counts + side, 1);
            if (slot < (side ? mig.cap_right : mig.cap)) {
                write_record<D>((side ? mig.right : mig.left) + (size_t) slot * RecordTraits<D>::WORDS, p, mv, pid);
                key = kKeyGone;
                gone = true;
            } else {
                atomicExch(mig.counts + 3, 1);  // send buffer overflow: fatal for the caller
            }
        }
        (void) gone;
        if (keys_out) keys_out[dst] = key;
#pragma unroll
        for (int d = 0; d < D; ++d) b[d] = min(max(b[d], 0), P.res - 2);
        b[0] += xoff;
        // migrants stay in the sender's box: the box table of the slab protocol must cover them until they are unpacked
        box_partial_write<D>(box_partial, live, b, true, i >> 5);
    }
}

// `perm` (nullable) fuses the re-binning reorder into this kernel: thread i reads slot perm[i] of `S`
// and writes slot i of `T` (T may equal S when perm is null), so after the kernel `T` is in cell-sorted
// order without a separate gather pass.  `keys_out` (nullable) receives the NEXT step's cell key of
// the advected particle, so the next step's sort starts without a key pass.
//
// WINDOW (3D): the grid nodes the CTA's 128 particles touch are staged in shared memory by the TMA.  Slots are
// cell-sorted (and stay grouped between two sorts: a particle moves < 1 cell per step), so the stencil bases of a CTA
// span a few cells per axis: the CTA reduces the bounding box of its bases, one thread issues one
// cp.async.bulk.tensor box per x-plane of the node window onto an mbarrier, the F rows of the particles are loaded
// while the boxes are in flight, and the 27-node gather reads shared memory (one base register + immediate offsets,
// 29-cycle latency) instead of 27 dependent-latency L1/L2 loads.  A CTA whose box does not fit the window (sparse or
// dispersed particles, a chunk that straddles the end of a tile row) gathers from global memory as before.
template <int D, int MODEL, bool WINDOW = false>
__global__ void __launch_bounds__(128, NMPM_G2P_MINB) k_g2p_gather(ParticleStore S, ParticleStore T, const uint32_t* __restrict__ perm,
                                                    uint32_t n, MaterialParams P, const float4* __restrict__ grid,
                                                    uint32_t* __restrict__ keys_out, int tiles_per_axis,
                                                    int* __restrict__ error_flag, MigrateArgs mig,
                                                    int* __restrict__ box_partial, const uint32_t* __restrict__ gone_keys,
                                                    int local_reorder, const __grid_constant__ CUtensorMap grid_map) {
    static_assert(!WINDOW || D == 3, "the TMA node window is a 3D path");
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    // slab mode between sorts: a slot whose particle migrated away is skipped (its key stays kKeyGone)
    const bool mine = i < n && !(gone_keys && __ldg(gone_keys + i) == kKeyGone);
    const unsigned live = __ballot_sync(0xffffffffu, mine);
    if (live == 0u) {  // nobody here owns a particle: leave an empty partial box (k_box_reduce reads every warp's slot)
        if ((threadIdx.x & 31) == 0) {
            int4* out = reinterpret_cast<int4*>(box_partial + (size_t) ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 8);
            out[0] = make_int4(0x7fffffff, 0x7fffffff, 0x7fffffff, (int) 0x80000000);
            out[1] = make_int4((int) 0x80000000, (int) 0x80000000, 0, 0);
        }
        if constexpr (!WINDOW) return;
    }
    if constexpr (!WINDOW) {
        if (!mine) return;
    }
    const uint32_t src = (mine && perm) ? __ldg(perm + i) : i;
    PState<D> p;
    int base[D];
    float fx[D], w[D][3];
    int xoff = 0;                                       // batch of stacked 2D scenes: node-row offset of this particle's scene
    [[maybe_unused]] bool in_window = false;            // WINDOW: the CTA's node box fits the window and has landed
    [[maybe_unused]] const float4* win_node = nullptr;  // WINDOW: this particle's base node inside the window
    if constexpr (WINDOW) {
        __shared__ __align__(128) float4 win[kWinX * kWinPitch];
        __shared__ __align__(8) uint64_t mbar;
        __shared__ int red[4][8];
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (threadIdx.x == 0) mbar_init(&mbar, 1);
        // positions first: the window origin depends on them
        float4 a0 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (mine) a0 = ld4(S.q[0] + src);
        p.x[0] = a0.x, p.x[1] = a0.y, p.x[2] = a0.z, p.Jp = a0.w;
        bool ok = stencil_of<D>(p.x, P, base, fx, w);
        if (mine && !ok) atomicOr(error_flag, 1);
        int mn[3], mx[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            mn[d] = __reduce_min_sync(0xffffffffu, mine ? base[d] : 0x7fffffff);
            mx[d] = __reduce_max_sync(0xffffffffu, mine ? base[d] : (int) 0x80000000);
        }
        if (lane == 0) {
            *reinterpret_cast<int4*>(&red[warp][0]) = make_int4(mn[0], mn[1], mn[2], mx[0]);
            *reinterpret_cast<int2*>(&red[warp][4]) = make_int2(mx[1], mx[2]);
        }
        __syncthreads();
#pragma unroll
        for (int wi = 0; wi < 4; ++wi) {
            const int4 a = *reinterpret_cast<const int4*>(&red[wi][0]);
            const int2 b = *reinterpret_cast<const int2*>(&red[wi][4]);
            mn[0] = min(mn[0], a.x), mn[1] = min(mn[1], a.y), mn[2] = min(mn[2], a.z);
            mx[0] = max(mx[0], a.w), mx[1] = max(mx[1], b.x), mx[2] = max(mx[2], b.y);
        }
        // node extents: bases lo..hi, stencil reaches +2 (CTA-uniform; an empty CTA has hi < lo and fails the test)
        const int ex = mx[0] - mn[0] + 3, ey = mx[1] - mn[1] + 3, ez = mx[2] - mn[2] + 3;
        in_window = ex >= 3 && ey >= 3 && ez >= 3 && ex <= kWinX && ey <= kWinY && ez <= kWinZ;
        if (in_window && threadIdx.x == 0) {
            mbar_arrive_expect_tx(&mbar, (uint32_t) ex * kWinPlaneBytes);
            for (int px = 0; px < ex; ++px) tma_load_4d(win + px * kWinPitch, &grid_map, &mbar, 0, mn[2], mn[1], mn[0] + px);
        }
        // the F rows travel while the boxes are in flight
        if (mine) {
            const float4 a1 = ld4(S.q[1] + src), a2 = ld4(S.q[2] + src), a3 = ld4(S.q[3] + src);
            p.F.m[0] = a1.x, p.F.m[1] = a1.y, p.F.m[2] = a1.z, p.F.m[3] = a1.w;
            p.F.m[4] = a2.x, p.F.m[5] = a2.y, p.F.m[6] = a2.z, p.F.m[7] = a2.w;
            p.F.m[8] = a3.x;
        }
        if (in_window) {
            // bounded wait: a box that never lands (it cannot, short of a bad descriptor) must not hang the device
            uint32_t spins = 0;
            while (!mbar_try_wait(&mbar, 0u)) {
                if (++spins > (1u << 22)) {
                    in_window = false;
                    atomicOr(error_flag, 4);
                    break;
                }
            }
        }
        if (!mine) return;
        win_node = win + ((base[0] - mn[0]) * kWinPitch + (base[1] - mn[1]) * kWinZ + (base[2] - mn[2]));
    } else {
        load_for_g2p<D>(S, src, p);
        xoff = scene_of_slot<D>(S, src, P) * P.n1;
        if (!stencil_of<D>(p.x, P, base, fx, w, xoff)) atomicOr(error_flag, 1);
    }
    // Gather with the stencil offsets centred on the middle node:  o = ijk - 1 in {-1,0,+1},
    //   v   = sum w g
    //   B_c = sum w g o_c                  (the o_c = 0 terms vanish at compile time)
    //   C   = 4 inv_dx (B - v (fx-1)^T)    == sum 4 inv_dx (w g) (ijk - fx)^T   (src/nclr.h:206,223)
    const int n1 = P.n1;
    const float four_inv_dx = 4.0f * P.inv_dx;
    float vn[D];
    Mat<D> Cn;
    if constexpr (D == 3) {
        if (WINDOW && in_window) {
            g2p_gather3(WindowNodes<kWinPitch, kWinZ>{win_node}, w, fx, four_inv_dx, vn, Cn);
        } else {
            const GlobalNodes nodes{grid + ((size_t) (base[0] * n1 + base[1]) * n1 + base[2]), n1 * n1, n1};
            g2p_gather3(nodes, w, fx, four_inv_dx, vn, Cn);
        }
    } else {
        float2 v01 = make_float2(0.0f, 0.0f), B01[D];
#pragma unroll
        for (int c = 0; c < D; ++c) B01[c] = make_float2(0.0f, 0.0f);
        const float2 plus1 = make_float2(1.0f, 1.0f), minus1 = make_float2(-1.0f, -1.0f);
#pragma unroll
        for (int ii = 0; ii < 3; ++ii)
#pragma unroll
            for (int jj = 0; jj < 3; ++jj) {
                const size_t node = (size_t) (base[0] + ii) * n1 + (base[1] + jj);
                const float4 g = ldg4(grid + node);
                const float weight = w[0][ii] * w[1][jj];
                const float2 wv01 = __fmul2_rn(make_float2(weight, weight), make_float2(g.x, g.y));
                v01 = __fadd2_rn(v01, wv01);
                const int o[2] = {ii - 1, jj - 1};
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    if (o[c] == 1) B01[c] = __ffma2_rn(wv01, plus1, B01[c]);
                    else if (o[c] == -1)
                        B01[c] = __ffma2_rn(wv01, minus1, B01[c]);
                }
            }
        vn[0] = v01.x, vn[1] = v01.y;
#pragma unroll
        for (int c = 0; c < D; ++c) {
            const float fc = fx[c] - 1.0f;
            Cn(0, c) = four_inv_dx * fmaf(-vn[0], fc, B01[c].x);
            Cn(1, c) = four_inv_dx * fmaf(-vn[1], fc, B01[c].y);
        }
    }
    g2p_finish<D, MODEL>(p, Cn, vn, S, T, perm, src, i, live, P, keys_out, tiles_per_axis, mig, box_partial, gone_keys,
                         local_reorder, xoff);
}

// ---- K4, software-pipelined (3D): persistent CTAs, asynchronous particle rows and node windows ------------------------
//
// The one-shot kernel above starts every CTA with a dependent chain — positions from DRAM, bounding box, node loads —
// and the 27 node loads of the gather are its long-scoreboard stalls (profiles/r01ze, r02c).  Here a CTA of 128 threads
// is persistent and walks chunks of 128 slots (chunk = blockIdx.x, += gridDim.x).  While chunk k is computed, everything
// chunk k+1 needs is already on its way:
//   [A] the rows q0..q3 of chunk k (x, Jp, F) are read from shared memory into registers; the rows of chunk k+1 are
//       requested with 16-byte cp.async copies (LDGSTS; through the sort permutation on re-binned steps);
//   [B] the node window of chunk k has landed (mbarrier, posted one iteration ago): 27-node gather from shared memory;
//   [C] the rows of k+1 have landed: bounding box of their stencil bases (warp reductions + one shared-memory round),
//       one elected thread issues one TMA box per x-plane of the window of chunk k+1 (the single window buffer is free:
//       every warp has passed its gather);
//   [D] F' = (I + dt C) F, snow projection / liquid reset, next cell key, re-grouping, stores — ~1000 instructions per
//       warp during which the window of k+1 travels.
// Chunks whose box exceeds the window (dispersed particles) gather from global memory; shared memory: 8 KB rows +
// 20 KB window per CTA, 8 CTAs per SM as before.
template <int MODEL>
__global__ void __launch_bounds__(128, NMPM_G2P_PIPE_MINB) k_g2p_pipe(ParticleStore S, ParticleStore T, const uint32_t* __restrict__ perm,
                                                                 uint32_t n, MaterialParams P, const float4* __restrict__ grid,
                                                                 uint32_t* __restrict__ keys_out, int tiles_per_axis,
                                                                 int* __restrict__ error_flag, MigrateArgs mig,
                                                                 int* __restrict__ box_partial, const uint32_t* __restrict__ gone_keys,
                                                                 int local_reorder, const __grid_constant__ CUtensorMap grid_map) {
    constexpr int D = 3;
    __shared__ __align__(128) float4 win[kWinX * kWinPitch];
    __shared__ __align__(16) float4 rows[4][128];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ int red[4][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t nchunks = (n + 127u) >> 7;
    if (blockIdx.x >= nchunks) return;
    if (threadIdx.x == 0) mbar_init(&mbar, 1);

    // request the rows of `chunk` for this thread's slot; returns whether the slot holds a particle of this slab
    auto request_rows = [&](uint32_t chunk, uint32_t& src_out) -> bool {
        const uint32_t i = chunk * 128u + threadIdx.x;
        const bool m = chunk < nchunks && i < n && !(gone_keys && __ldg(gone_keys + i) == kKeyGone);
        src_out = i;
        if (m) {
            const uint32_t src = perm ? __ldg(perm + i) : i;
            src_out = src;
#pragma unroll
            for (int k = 0; k < 4; ++k) cp_async16(&rows[k][threadIdx.x], S.q[k] + src);
        }
        cp_async_commit();
        return m;
    };
    // bounding box of the CTA's stencil bases -> window origin; the elected thread posts the TMA boxes.  Returns whether
    // the chunk gathers from the window (CTA-uniform).
    auto post_window = [&](bool m, const int (&base)[3], int (&origin)[3]) -> bool {
        int mn[3], mx[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            mn[d] = __reduce_min_sync(0xffffffffu, m ? base[d] : 0x7fffffff);
            mx[d] = __reduce_max_sync(0xffffffffu, m ? base[d] : (int) 0x80000000);
        }
        if (lane == 0) {
            *reinterpret_cast<int4*>(&red[warp][0]) = make_int4(mn[0], mn[1], mn[2], mx[0]);
            *reinterpret_cast<int2*>(&red[warp][4]) = make_int2(mx[1], mx[2]);
        }
        __syncthreads();  // also: every warp is past its gather from the window, and past reading `red` of the last round
#pragma unroll
        for (int wi = 0; wi < 4; ++wi) {
            const int4 a = *reinterpret_cast<const int4*>(&red[wi][0]);
            const int2 b = *reinterpret_cast<const int2*>(&red[wi][4]);
            mn[0] = min(mn[0], a.x), mn[1] = min(mn[1], a.y), mn[2] = min(mn[2], a.z);
            mx[0] = max(mx[0], a.w), mx[1] = max(mx[1], b.x), mx[2] = max(mx[2], b.y);
        }
        const int ex = mx[0] - mn[0] + 3, ey = mx[1] - mn[1] + 3, ez = mx[2] - mn[2] + 3;
        const bool fits = ex >= 3 && ey >= 3 && ez >= 3 && ex <= kWinX && ey <= kWinY && ez <= kWinZ;
        if (fits && threadIdx.x == 0) {
            mbar_arrive_expect_tx(&mbar, (uint32_t) ex * kWinPlaneBytes);
            for (int px = 0; px < ex; ++px) tma_load_4d(win + px * kWinPitch, &grid_map, &mbar, 0, mn[2], mn[1], mn[0] + px);
        }
        origin[0] = mn[0], origin[1] = mn[1], origin[2] = mn[2];
        __syncthreads();  // `red` may be rewritten by the next round only after everybody has read it
        return fits;
    };

    // ---- prologue: rows and window of the first chunk ----------------------------------------------------------
    uint32_t chunk = blockIdx.x;
    uint32_t src_next;
    bool mine_next = request_rows(chunk, src_next);
    cp_async_wait_all();
    // stencil base of the staged position (clamped like stencil_of does): all the window origin needs; the full stencil is
    // recomputed at [A] rather than carried in registers across [D]
    auto staged_base = [&](bool m, int (&b)[3]) {
        const float4 a0 = m ? rows[0][threadIdx.x] : make_float4(0.f, 0.f, 0.f, 0.f);
        const float xs[3] = {a0.x, a0.y, a0.z};
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const Stencil1 st = stencil_axis(xs[d], P.inv_dx, P.res);
            b[d] = st.ok ? st.base : min(max(st.base, 0), P.res - 2);
        }
    };
    int origin_n[3];
    bool inwin_next;
    {
        int bn[3];
        staged_base(mine_next, bn);
        inwin_next = post_window(mine_next, bn, origin_n);
    }
    uint32_t parity = 0;

    for (; chunk < nchunks; chunk += gridDim.x) {
        // ---- [A] this chunk's particle out of the staged rows; next chunk's rows requested --------------------
        const bool mine = mine_next, in_window = inwin_next;
        const uint32_t src = src_next, i = chunk * 128u + threadIdx.x;
        PState<D> p;
        int base[3] = {0, 0, 0};
        const int origin[3] = {origin_n[0], origin_n[1], origin_n[2]};
        float fx[3], w[3][3];
        if (mine) {
            const float4 a0 = rows[0][threadIdx.x], a1 = rows[1][threadIdx.x], a2 = rows[2][threadIdx.x], a3 = rows[3][threadIdx.x];
            p.x[0] = a0.x, p.x[1] = a0.y, p.x[2] = a0.z, p.Jp = a0.w;
            p.F.m[0] = a1.x, p.F.m[1] = a1.y, p.F.m[2] = a1.z, p.F.m[3] = a1.w;
            p.F.m[4] = a2.x, p.F.m[5] = a2.y, p.F.m[6] = a2.z, p.F.m[7] = a2.w;
            p.F.m[8] = a3.x;
            if (!stencil_of<D>(p.x, P, base, fx, w)) atomicOr(error_flag, 1);
        }
        const unsigned live = __ballot_sync(0xffffffffu, mine);
        const uint32_t next = chunk + gridDim.x;
        mine_next = request_rows(next, src_next);  // each thread overwrites only the row slots it has just read

        // ---- [B] gather: from the window once it has landed, else from global memory ---------------------------
        float vn[3];
        Mat<3> Cn;
        const float four_inv_dx = 4.0f * P.inv_dx;
        bool use_window = in_window;
        if (in_window) {
            uint32_t spins = 0;
            while (!mbar_try_wait(&mbar, parity)) {
                if (++spins > (1u << 22)) {  // cannot happen short of a bad descriptor: do not hang the device
                    use_window = false;
                    atomicOr(error_flag, 4);
                    break;
                }
            }
            parity ^= 1u;
        }
        if (mine) {
            if (use_window) {
                const float4* node = win + ((base[0] - origin[0]) * kWinPitch + (base[1] - origin[1]) * kWinZ + (base[2] - origin[2]));
                g2p_gather3(WindowNodes<kWinPitch, kWinZ>{node}, w, fx, four_inv_dx, vn, Cn);
            } else {
                const int n1 = P.n1;
                const GlobalNodes nodes{grid + ((size_t) (base[0] * n1 + base[1]) * n1 + base[2]), n1 * n1, n1};
                g2p_gather3(nodes, w, fx, four_inv_dx, vn, Cn);
            }
        }

        // ---- [C] next chunk: rows have landed -> stencil, bounding box, window boxes posted -----------------------
        cp_async_wait_all();
        {
            int bn[3];
            staged_base(mine_next, bn);
            inwin_next = post_window(mine_next, bn, origin_n);  // (next >= nchunks: nobody is `mine`, nothing is posted)
        }

        // ---- [D] update, re-binning, stores ---------------------------------------------------------------------------
        if (live == 0u) {
            if (lane == 0) {
                int4* out = reinterpret_cast<int4*>(box_partial + (size_t) (i >> 5) * 8);
                out[0] = make_int4(0x7fffffff, 0x7fffffff, 0x7fffffff, (int) 0x80000000);
                out[1] = make_int4((int) 0x80000000, (int) 0x80000000, 0, 0);
            }
        } else if (mine) {
            g2p_finish<D, MODEL>(p, Cn, vn, S, T, perm, src, i, live, P, keys_out, tiles_per_axis, mig, box_partial, gone_keys,
                                 local_reorder);
        }
    }
}

// slab migration, receiving side: append records to slots [first, first + count) and bin them
template <int D>
__global__ void __launch_bounds__(256) k_unpack_records(const float* __restrict__ rec, uint32_t count, uint32_t first,
                                                        ParticleStore T, MaterialParams P, int tiles_per_axis,
                                                        uint32_t* __restrict__ keys, GridBox* __restrict__ box) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    const float* r = rec + (size_t) j * RecordTraits<D>::WORDS;
    PState<D> p;
#pragma unroll
    for (int d = 0; d < D; ++d) p.x[d] = r[d], p.v[d] = r[D + d];
#pragma unroll
    for (int k = 0; k < D * D; ++k) p.F.m[k] = r[2 * D + k], p.C.m[k] = r[2 * D + D * D + k];
    p.Jp = r[2 * D + 2 * D * D];
    const uint32_t i = first + j;
    store_state<D>(T, i, p);
    T.mv[i] = make_float2(r[2 * D + 2 * D * D + 1], r[2 * D + 2 * D * D + 2]);
    T.id[i] = __float_as_uint(r[2 * D + 2 * D * D + 3]);
    int b[D];
    bool bad = false;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const Stencil1 s = stencil_axis(p.x[d], P.inv_dx, P.res);
        b[d] = s.base;
        bad = bad || !s.ok;
    }
    keys[i] = bad ? kKeyOutOfGrid : cell_key<D>(b, tiles_per_axis);
#pragma unroll
    for (int d = 0; d < D; ++d) b[d] = min(max(b[d], 0), P.res - 2);
    box_update<D>(box, b, true);
}

// ---- device-driven slab step (nmpm_slab_comm.inl): particle counts live on the device ------------------------------
// ctr[0] = slots in use in the current store (true value; the host only keeps an upper bound for launch sizes),
// ctr[1] = slots whose particle migrated away since the last sort, ctr[2] = error bits (1: store capacity exceeded)
//
// A migrant message is [header record][records ...]; header ints: [0] count, [1..3] lo, [4..6] hi of the sender's node
// box for the coming step, [7] the sender's step number.
constexpr int kHdrInts = 8;

// after G2P: headers of the two outgoing messages, gone count, read-back record.  The header pointers address the
// NEIGHBOURS' inboxes (peer memory over NVLink, or local send buffers on the NCCL path); `flag_*` (nullable) are the
// neighbours' arrival flags: the records were written by the G2P kernel before this one in stream order, the header by
// this thread, then a system-scope fence, then the flag — a neighbour that sees the flag sees all of it.
template <int D>
__global__ void k_slab_post(const int* __restrict__ counts /* n_left, n_right, n_kept, overflow */, const GridBox* __restrict__ box,
                            int* hdr_left, int* hdr_right, int* __restrict__ ctr, int step, int* __restrict__ ring,
                            int slot_bound, int* flag_left, int* flag_right) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int nl = counts[0], nr = counts[1];
    if (ctr[0] > slot_bound) ctr[2] |= 8;  // this step's launches did not cover every slot in use
    int* h[2] = {hdr_left, hdr_right};
    for (int s = 0; s < 2; ++s) {
        if (!h[s]) continue;
        volatile int* hv = h[s];
        hv[0] = s ? nr : nl;
        for (int d = 0; d < 3; ++d) hv[1 + d] = box->lo[d], hv[4 + d] = box->hi[d];
        hv[7] = step;
    }
    ctr[1] += nl + nr;
    if (counts[3]) ctr[2] |= 2;  // a send buffer overflowed
    // read-back record of this step, first part: counts + box
    for (int k = 0; k < 4; ++k) ring[k] = counts[k];
    for (int d = 0; d < 3; ++d) ring[4 + d] = box->lo[d], ring[7 + d] = box->hi[d];
    ring[10] = step, ring[11] = 0;
    if (flag_left || flag_right) {
        __threadfence_system();
        if (flag_left) *((volatile int*) flag_left) = step;
        if (flag_right) *((volatile int*) flag_right) = step;
    }
}

// receiving side of the peer-memory path: wait until both neighbours have posted step `step` into this rank's inboxes.
// Bounded: a neighbour that never arrives (it died) must not hang the device — the header count is zeroed, the error
// bit set, and the host raises the error when it reads the record.
__global__ void k_slab_wait(const int* flag_from_left, const int* flag_from_right, int step, int* hdr_from_left,
                            int* hdr_from_right, int* __restrict__ ctr) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int* f[2] = {flag_from_left, flag_from_right};
    int* hd[2] = {hdr_from_left, hdr_from_right};
    const long long t0 = clock64();
    for (int s = 0; s < 2; ++s) {
        if (!f[s]) continue;
        bool ok = false;
        while (true) {
            int v;
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(f[s]) : "memory");
            if (v >= step) {
                ok = true;
                break;
            }
            if (clock64() - t0 > 20000000000LL) break;  // ~10 s at 2 GHz
            __nanosleep(200);
        }
        if (!ok) {
            hd[s][0] = 0;
            atomicOr(ctr + 2, 4);
        }
    }
    __threadfence_system();
}

// receiving side: append the records of both neighbours behind slot ctr[0] and bin them; counts come from the headers
template <int D>
__global__ void __launch_bounds__(256) k_unpack_records2(const float* rec_left, const float* rec_right,
                                                         uint32_t cap_msg_left, uint32_t cap_msg_right, uint32_t cap_store,
                                                         ParticleStore T, MaterialParams P, int tiles_per_axis,
                                                         uint32_t* __restrict__ keys, GridBox* __restrict__ box, int* __restrict__ ctr) {
    constexpr int W = RecordTraits<D>::WORDS;  // (rec_*: no __restrict__/__ldg — written by another GPU during this launch's lifetime)
    const uint32_t cl = rec_left ? min((uint32_t) __float_as_int(rec_left[0]), cap_msg_left) : 0u;
    const uint32_t cr = rec_right ? min((uint32_t) __float_as_int(rec_right[0]), cap_msg_right) : 0u;
    const uint32_t first = (uint32_t) ctr[0];  // not modified by this kernel (k_ctr_after_unpack follows)
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < cl + cr; j += gridDim.x * blockDim.x) {
    const uint32_t i = first + j;
    if (i >= cap_store) {
        atomicOr(ctr + 2, 1);
        return;
    }
    const float* r = (j < cl) ? rec_left + (size_t) (1 + j) * W : rec_right + (size_t) (1 + j - cl) * W;
    PState<D> p;
#pragma unroll
    for (int d = 0; d < D; ++d) p.x[d] = r[d], p.v[d] = r[D + d];
#pragma unroll
    for (int k = 0; k < D * D; ++k) p.F.m[k] = r[2 * D + k], p.C.m[k] = r[2 * D + D * D + k];
    p.Jp = r[2 * D + 2 * D * D];
    store_state<D>(T, i, p);
    T.mv[i] = make_float2(r[2 * D + 2 * D * D + 1], r[2 * D + 2 * D * D + 2]);
    T.id[i] = __float_as_uint(r[2 * D + 2 * D * D + 3]);
    int b[D];
    bool bad = false;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const Stencil1 s = stencil_axis(p.x[d], P.inv_dx, P.res);
        b[d] = s.base;
        bad = bad || !s.ok;
    }
    keys[i] = bad ? kKeyOutOfGrid : cell_key<D>(b, tiles_per_axis);
#pragma unroll
    for (int d = 0; d < D; ++d) b[d] = min(max(b[d], 0), P.res - 2);
    box_update<D>(box, b, true);
    }
}

__global__ void k_ctr_after_unpack(const float* rec_left, const float* rec_right, uint32_t cap_msg_left,
                                   uint32_t cap_msg_right, uint32_t cap_store, int* __restrict__ ctr, int* __restrict__ ring) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const uint32_t cl = rec_left ? min((uint32_t) __float_as_int(rec_left[0]), cap_msg_left) : 0u;
    const uint32_t cr = rec_right ? min((uint32_t) __float_as_int(rec_right[0]), cap_msg_right) : 0u;
    const uint32_t n = min((uint32_t) ctr[0] + cl + cr, cap_store);
    ctr[0] = (int) n;
    // read-back record, second part: the two received headers and the counters
    for (int k = 0; k < kHdrInts; ++k) {
        ring[12 + k] = rec_left ? __float_as_int(rec_left[k]) : 0;
        ring[20 + k] = rec_right ? __float_as_int(rec_right[k]) : 0;
    }
    for (int k = 0; k < 4; ++k) ring[28 + k] = ctr[k];
}

// after a sort: the migrated-away slots have dropped to the end of the order and out of the count
__global__ void k_ctr_after_sort(int* __restrict__ ctr) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    ctr[0] -= ctr[1];
    ctr[1] = 0;
}

// sub-rectangle of `planes` consecutive node planes: a = slow in-plane axis (y in 3D, none in 2D),
// b = fast axis (z in 3D, y in 2D).  mode 0: buf <- grid, 1: grid += buf, 2: grid <- 0
template <int MODE>
__global__ void __launch_bounds__(256) k_rect(float4* __restrict__ grid, size_t plane_nodes, int n1, int x_plane, int planes,
                                              int a0, int na, int b0, int nb, float4* __restrict__ buf) {
    const uint32_t total = (uint32_t) planes * na * nb;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t q = i / nb, b = i - q * nb;
        const uint32_t p = q / na, a = q - p * na;
        float4* g = grid + (size_t) (x_plane + p) * plane_nodes + (size_t) (a0 + a) * n1 + (b0 + b);
        if (MODE == 0) buf[i] = *g;
        else if (MODE == 1) {
            float4 v = *g;
            const float4 w = buf[i];
            v.x += w.x, v.y += w.y, v.z += w.z, v.w += w.w;
            *g = v;
        } else
            *g = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
}

// live particles per base.x (re-balancing of the slab boundaries)
template <int D>
__global__ void __launch_bounds__(256) k_histogram_x(ParticleStore S, uint32_t n, MaterialParams P, int* __restrict__ hist,
                                                     const uint32_t* __restrict__ keys) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (keys && keys[i] == kKeyGone) return;  // migrated away: counted by its new owner
    float x[D];
    load_position<D>(S, i, x);
    const Stencil1 s = stencil_axis(x[0], P.inv_dx, P.res);
    atomicAdd(hist + min(max(s.base, 0), P.res), 1);
}

template <int D>
__global__ void __launch_bounds__(256) k_set_ids(ParticleStore S, uint32_t n, const uint32_t* __restrict__ ids) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) S.id[i] = ids[S.id[i]];
}

// ---- K5: import / export between the interchange layouts and the device store ---------------
// AoS record = the reference's Particle<dim> (src/nclr.h:20-48): x, v, F, C, Jp, mass, volume, c
//   2D: 16 words (64 B), 3D: 28 words (112 B).  stride_w = record stride in 4-byte words.
template <int D>
__global__ void __launch_bounds__(256) k_import_aos(const float* __restrict__ aos, size_t stride_w, uint32_t n,
                                                    ParticleStore S) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* r = aos + (size_t) i * stride_w;
    PState<D> p;
#pragma unroll
    for (int d = 0; d < D; ++d) p.x[d] = r[d], p.v[d] = r[D + d];
#pragma unroll
    for (int k = 0; k < D * D; ++k) p.F.m[k] = r[2 * D + k], p.C.m[k] = r[2 * D + D * D + k];
    p.Jp = r[2 * D + 2 * D * D];
    store_state<D>(S, i, p);
    S.mv[i] = make_float2(r[2 * D + 2 * D * D + 1], r[2 * D + 2 * D * D + 2]);
    S.id[i] = i;
}

// writes x, v, F, C, Jp of slot i into record id[i] (input order); mass/volume/c are left untouched
template <int D>
__global__ void __launch_bounds__(256) k_export_aos(ParticleStore S, uint32_t n, float* __restrict__ aos,
                                                    size_t stride_w) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    PState<D> p;
    load_for_p2g<D>(S, i, p);
    float* r = aos + (size_t) S.id[i] * stride_w;
#pragma unroll
    for (int d = 0; d < D; ++d) r[d] = p.x[d], r[D + d] = p.v[d];
#pragma unroll
    for (int k = 0; k < D * D; ++k) r[2 * D + k] = p.F.m[k], r[2 * D + D * D + k] = p.C.m[k];
    r[2 * D + 2 * D * D] = p.Jp;
}

// SoA interchange arrays (any may be null on import => reference defaults, src/nclr.h:46-47)
template <int D>
__global__ void __launch_bounds__(256) k_import_soa(const float* __restrict__ x, const float* __restrict__ v,
                                                    const float* __restrict__ F, const float* __restrict__ C,
                                                    const float* __restrict__ Jp, const float* __restrict__ mass,
                                                    const float* __restrict__ volume, uint32_t n, ParticleStore S,
                                                    int keep_constants) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    PState<D> p;
#pragma unroll
    for (int d = 0; d < D; ++d) p.x[d] = x[(size_t) i * D + d], p.v[d] = v ? v[(size_t) i * D + d] : 0.0f;
#pragma unroll
    for (int k = 0; k < D * D; ++k) {
        p.F.m[k] = F ? F[(size_t) i * D * D + k] : 0.0f;
        p.C.m[k] = C ? C[(size_t) i * D * D + k] : 0.0f;
    }
    if (!F) p.F(0, 0) = 1.0f, p.F(1, 1) = 1.0f;  // diag<dim>(1): Q1
    p.Jp = Jp ? Jp[i] : 1.0f;
    store_state<D>(S, i, p);
    if (!keep_constants) S.mv[i] = make_float2(mass ? mass[i] : 1.0f, volume ? volume[i] : 1.0f);
    S.id[i] = i;
}

template <int D>
__global__ void __launch_bounds__(256) k_export_soa(ParticleStore S, uint32_t n, float* __restrict__ x,
                                                    float* __restrict__ v, float* __restrict__ F, float* __restrict__ C,
                                                    float* __restrict__ Jp, uint32_t* __restrict__ ids_out,
                                                    const uint32_t* __restrict__ keys) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    PState<D> p;
    load_for_p2g<D>(S, i, p);
    // ids_out != nullptr: slot order + the id of each slot (slab download); else input order.
    // keys != nullptr: slots whose particle migrated away (kKeyGone) report id 0xFFFFFFFF.
    const size_t o = ids_out ? (size_t) i : (size_t) S.id[i];
    if (ids_out) ids_out[i] = (keys && keys[i] == kKeyGone) ? 0xFFFFFFFFu : S.id[i];
#pragma unroll
    for (int d = 0; d < D; ++d) {
        if (x) x[o * D + d] = p.x[d];
        if (v) v[o * D + d] = p.v[d];
    }
#pragma unroll
    for (int k = 0; k < D * D; ++k) {
        if (F) F[o * D * D + k] = p.F.m[k];
        if (C) C[o * D * D + k] = p.C.m[k];
    }
    if (Jp) Jp[o] = p.Jp;
}

// When slots must return to input order (upload of a new state keeps mass/volume): inverse of id
template <int D>
__global__ void __launch_bounds__(256) k_restore_constants(ParticleStore src, ParticleStore dst, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dst.mv[src.id[i]] = src.mv[i];
}

// grid export: float4 nodes -> (gv, gm) SoA or Cell<dim> AoS (src/nclr.h:50-55: velocity then mass)
template <int D>
__global__ void __launch_bounds__(256) k_export_grid(const float4* __restrict__ grid, size_t cells, float* __restrict__ gv,
                                                     float* __restrict__ gm, float* __restrict__ aos, size_t stride_w) {
    const size_t idx = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= cells) return;
    const float4 g = grid[idx];
    const float vel[3] = {g.x, g.y, g.z};
    const float m = (D == 3) ? g.w : g.z;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        if (gv) gv[idx * D + d] = vel[d];
        if (aos) aos[idx * stride_w + d] = vel[d];
    }
    if (gm) gm[idx] = m;
    if (aos) aos[idx * stride_w + D] = m;
}

// ---- unit hooks -------------------------------------------------------------------------------
template <int D>
__global__ void k_svd_batch(const float* __restrict__ A, size_t count, float* __restrict__ U, float* __restrict__ S,
                            float* __restrict__ V, float* __restrict__ R, float* __restrict__ G, float lo, float hi) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    Mat<D> a, u, v;
    float sig[D];
#pragma unroll
    for (int k = 0; k < D * D; ++k) a.m[k] = A[i * D * D + k];
    if (R) {
        const Mat<D> r = nclr_polar_R(a);
#pragma unroll
        for (int k = 0; k < D * D; ++k) R[i * D * D + k] = r.m[k];
    }
    if (G) {
        const Mat<D> g = snow_project(a, lo, hi);
#pragma unroll
        for (int k = 0; k < D * D; ++k) G[i * D * D + k] = g.m[k];
    }
    if (U) {
        nclr_svd<D>(a, u, sig, v);
#pragma unroll
        for (int k = 0; k < D * D; ++k) {
            U[i * D * D + k] = u.m[k];
            V[i * D * D + k] = v.m[k];
            S[i * D * D + k] = 0.0f;
        }
#pragma unroll
        for (int d = 0; d < D; ++d) S[i * D * D + d + d * D] = sig[d];
    }
}

template <int D, int MODEL>
__global__ void k_affine_debug(ParticleStore S, uint32_t n, MaterialParams P, float* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    PState<D> p;
    load_for_p2g<D>(S, i, p);
    const Mat<D> A = affine_matrix<D, MODEL>(p.F, p.C, p.Jp, p.mass, p.volume, P);
    const size_t o = S.id[i];
#pragma unroll
    for (int k = 0; k < D * D; ++k) out[o * D * D + k] = A.m[k];
}

}  // namespace nmpm
