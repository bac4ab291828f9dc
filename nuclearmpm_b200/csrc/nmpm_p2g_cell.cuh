// K2: P2G by warp-level segmented reduction over cell-sorted particles.  Three kernels share the scheme
// (phase A: lane = particle, packet to shared memory; phase B: lanes = stencil nodes walk the packets, accumulate a
// run of particles of one cell in registers, one vector reduction per node and run):
//   variant 2  k_p2g_cell<D,MODEL>    lane = stencil node, particle pairs in packed fp32x2   (2D default; 3D baseline)
//   variant 3  k_p2g_cols<MODEL>      three 9-lane groups, lane = stencil column (j,k), per-column packets   (3D default)
//   variant 4  k_p2g_streams<MODEL>   variant 3 over three particle streams per warp   (3D default from 8 Mi particles)
//
// Variant 2:
//
// Reference: p2g() + compute_fused_momentum() + first_piola_kirchoff_stress(), src/nclr.h:104-165,
// 313-337 — there a serial loop doing 3^dim read-modify-writes per particle.  Here:
//
//   phase A (lane = particle): each warp takes 32 consecutive slots of the cell-sorted store, loads
//     them with coalesced float4 reads, runs polar/SVD + stress in registers and leaves a folded
//     packet in shared memory:
//         b  = mass*v - dx * A*fx        (so that  mass*v + A*((ijk-fx)*dx) = b + (dx*A)*ijk )
//         A' = dx * A,  mass,  the 3x3 per-axis weights,  linear index of the stencil's base node
//   phase B (lane = stencil node, 27 of 32 lanes in 3D, 9 in 2D): the warp walks its packets in slot
//     order TWO PARTICLES AT A TIME with packed fp32x2 math (FFMA2/FMUL2, sm_100+): packets of
//     particles 2t and 2t+1 are interleaved in shared memory so one LDS.128 broadcast yields two
//     values for both particles as register pairs.  Every lane evaluates ITS node's weight and fused
//     momentum and accumulates in registers (even/odd particle sums in the two halves).  Consecutive
//     particles of the same cell form a segment; at a segment boundary each lane issues ONE vector
//     reduction (RED.E.ADD.F32x4 {px,py,pz,m}) for its node.
//
// Global atomics per particle drop from 3^dim*(dim+1) scalar (or 3^dim vector) to 3^dim/ppc vector
// reductions (ppc = particles per cell; 3.4 at 8 ppc in 3D), and no shared-memory float atomics are
// used at all — on sm_100a those are CAS loops (ATOMS.CAST.SPIN), see DESIGN.md.
#pragma once
#include "nmpm_kernels.cuh"

namespace nmpm {

constexpr int kP2GWarps = 4;
constexpr int kP2GColsMinB = 6;  // CTAs per SM of the column-lane kernels: 80 registers, no spills in the node walk
#ifndef NMPM_P2G_STREAMS_MINB
#define NMPM_P2G_STREAMS_MINB 5
#endif
constexpr int kP2GStreamsMinB = NMPM_P2G_STREAMS_MINB;

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 splat2(float a) { return make_float2(a, a); }

template <int D>
struct P2GPacket {
    // values per particle that phase B consumes through packed loads: b (D), A' (D*D), mass
    static constexpr int NV = D + D * D + 1;          // 13 (3D) / 7 (2D)
    static constexpr int NQ4 = (NV + 1) / 2;          // float4 words per particle PAIR: 7 / 4
    static constexpr int WSTRIDE = 34;                // padded row stride of the weight table (bank spread)
};

template <int D, int MODEL>
__global__ void __launch_bounds__(kP2GWarps * 32, NMPM_P2G_MINB) k_p2g_cell(ParticleStore S, const uint32_t* __restrict__ perm,
                                                             uint32_t n, MaterialParams P, float4* __restrict__ grid,
                                                             int* __restrict__ error_flag,
                                                             const uint32_t* __restrict__ gone_keys) {
    using PK = P2GPacket<D>;
    constexpr int NODES = (D == 3) ? 27 : 9;
    // pk[warp][q][t] = { val_{2q}(2t), val_{2q}(2t+1), val_{2q+1}(2t), val_{2q+1}(2t+1) }
    __shared__ float4 pk[kP2GWarps][PK::NQ4][16];
    __shared__ __align__(8) float wt[kP2GWarps][D * 3][PK::WSTRIDE];  // wt[d*3+i][slot]
    __shared__ __align__(8) int node0[kP2GWarps][32];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t first = (blockIdx.x * kP2GWarps + warp) * 32u;
    if (first >= n) return;
    const int cnt = min(32u, n - first);
    const int n1 = P.n1;

    // ---- phase A ---------------------------------------------------------------------------
    {
        float vals[PK::NV + 1];
        float w[D][3];
        int nd = -1;
#pragma unroll
        for (int k = 0; k <= PK::NV; ++k) vals[k] = 0.0f;
#pragma unroll
        for (int d = 0; d < D; ++d) w[d][0] = w[d][1] = w[d][2] = 0.0f;
        // slab mode: a slot whose particle migrated away (or that was never filled) keeps the all-zero packet and
        // node -2 (its rows are never read); -1 marks the padding slot of an odd tail
        const bool gone = lane < cnt && gone_keys && __ldg(gone_keys + first + lane) == kKeyGone;
        if (gone) nd = -2;
        if (lane < cnt && !gone) {
            PState<D> p;
            // `perm` (nullable): the store is read THROUGH the sorted permutation, no reorder pass
            const uint32_t src = perm ? __ldg(perm + first + lane) : first + lane;
            load_for_p2g<D>(S, src, p);
            int base[D];
            float fx[D];
            const int scene = scene_of_slot<D>(S, src, P);  // batch of stacked 2D scenes (else 0)
            if (!stencil_of<D>(p.x, P, base, fx, w, scene * P.n1)) atomicOr(error_flag, 1);
            const Mat<D> A = affine_matrix<D, MODEL>(p.F, p.C, p.Jp, p.mass, p.volume, (D == 2) ? scene_params(P, scene) : P);
#pragma unroll
            for (int r = 0; r < D; ++r) {
                float afx = A(r, 0) * fx[0];
#pragma unroll
                for (int k = 1; k < D; ++k) afx = fmaf(A(r, k), fx[k], afx);
                vals[r] = fmaf(-P.dx, afx, p.v[r] * p.mass);  // b_r
#pragma unroll
                for (int c = 0; c < D; ++c) vals[D + r * D + c] = P.dx * A(r, c);  // A'_rc (row-major here)
            }
            vals[D + D * D] = p.mass;
            nd = base[0] * n1 + base[1];
            if constexpr (D == 3) nd = nd * n1 + base[2];
        }
        float* pkf = reinterpret_cast<float*>(&pk[warp][0][0]);
        const int t = lane >> 1, h = lane & 1;
#pragma unroll
        for (int k = 0; k < PK::NV; ++k) pkf[((k >> 1) * 16 + t) * 4 + (k & 1) * 2 + h] = vals[k];
        if constexpr (PK::NV & 1) pkf[((PK::NV >> 1) * 16 + t) * 4 + 2 + h] = 0.0f;
#pragma unroll
        for (int d = 0; d < D; ++d)
#pragma unroll
            for (int i = 0; i < 3; ++i) wt[warp][d * 3 + i][lane] = w[d][i];
        node0[warp][lane] = nd;
    }
    __syncwarp();

    // ---- phase B ---------------------------------------------------------------------------
    if (lane >= NODES) return;
    int ijk[D];
    if constexpr (D == 3) {
        ijk[0] = lane / 9, ijk[1] = (lane / 3) % 3, ijk[2] = lane % 3;
    } else {
        ijk[0] = lane / 3, ijk[1] = lane % 3;
    }
    float2 fi[D];
    int lane_off = 0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        fi[d] = splat2((float) ijk[d]);
        lane_off = lane_off * n1 + ijk[d];
    }
    const float* wrow[D];
#pragma unroll
    for (int d = 0; d < D; ++d) wrow[d] = &wt[warp][d * 3 + ijk[d]][0];

    float2 acc[D], acc_m = splat2(0.0f);
#pragma unroll
    for (int d = 0; d < D; ++d) acc[d] = splat2(0.0f);
    int cur = node0[warp][0];

    auto flush = [&](int node) {
        float mom[D];
#pragma unroll
        for (int d = 0; d < D; ++d) mom[d] = acc[d].x + acc[d].y;
        if (node >= 0)  // a negative node is the (all-zero) run of a migrated-away slot
            red_add_f32x4(grid + (size_t) (node + lane_off), node_pack<D>(mom, acc_m.x + acc_m.y));
#pragma unroll
        for (int d = 0; d < D; ++d) acc[d] = splat2(0.0f);
        acc_m = splat2(0.0f);
    };

    const int npairs = (cnt + 1) >> 1;
    for (int t = 0; t < npairs; ++t) {
        const int2 nn = *reinterpret_cast<const int2*>(&node0[warp][2 * t]);
        // packed values for the pair (2t, 2t+1)
        float2 val[PK::NV + 1];
#pragma unroll
        for (int q = 0; q < PK::NQ4; ++q) {
            const float4 f = pk[warp][q][t];
            val[2 * q] = make_float2(f.x, f.y);
            val[2 * q + 1] = make_float2(f.z, f.w);
        }
        float2 weight = *reinterpret_cast<const float2*>(wrow[0] + 2 * t);
#pragma unroll
        for (int d = 1; d < D; ++d) weight = fmul2(weight, *reinterpret_cast<const float2*>(wrow[d] + 2 * t));
        // q_r = b_r + A'_r . ijk ; contribution = weight * q_r
        float2 q[D];
#pragma unroll
        for (int r = 0; r < D; ++r) {
            q[r] = val[r];
#pragma unroll
            for (int c = 0; c < D; ++c) q[r] = ffma2(val[D + r * D + c], fi[c], q[r]);
        }
        const float2 mass = val[D + D * D];
        const int nb = (nn.y < 0) ? nn.x : nn.y;  // odd tail: the padding slot has zero weight
        if (nn.x == cur && nb == cur) {           // warp-uniform: both particles continue the segment
#pragma unroll
            for (int r = 0; r < D; ++r) acc[r] = ffma2(weight, q[r], acc[r]);
            acc_m = ffma2(weight, mass, acc_m);
        } else {
            // a segment ends inside this pair: scalar math on the halves of the packed accumulators (no
            // re-packing of half-zero operand pairs)
            if (nn.x != cur) {
                flush(cur);
                cur = nn.x;
            }
#pragma unroll
            for (int r = 0; r < D; ++r) acc[r].x = fmaf(weight.x, q[r].x, acc[r].x);
            acc_m.x = fmaf(weight.x, mass.x, acc_m.x);
            if (nb != cur) {
                flush(cur);
                cur = nb;
            }
#pragma unroll
            for (int r = 0; r < D; ++r) acc[r].y = fmaf(weight.y, q[r].y, acc[r].y);
            acc_m.y = fmaf(weight.y, mass.y, acc_m.y);
        }
    }
    flush(cur);
}

// ---- K2 (variant 3, 3D): three 9-lane groups per warp, lane = stencil column, per-column packets --
//
// ncu on the lane = node kernel above (profiles/r01d, r01e): it is bound by the L1/shared DATA PIPE (l1tex
// data-pipe wavefronts 83-93 % of peak), not by issue slots.  A 128-bit shared load costs 4 wavefronts
// whether or not it is a broadcast, so what matters is how many BYTES each lane pulls out of shared memory
// (there: 68 B per lane per particle x 32 lanes), and a vector reduction costs ~1.3 LSU cycles per LANE.
// Here the 27 active lanes form THREE groups of 9; group g walks its own contiguous third of the warp's 32
// cell-sorted slots ([0,11) [11,22) [22,32)) one particle at a time, a lane is the (j,k) column of the
// stencil and owns its three nodes i = 0,1,2 in registers.  Phase A leaves, per particle, the quantities
// already specialised per column:
//     chunk jk (9x):  { b + A'_.1 j + A'_.2 k  (xyz),  wy[j] wz[k] }        read by ONE lane
//     chunk 9:        { A'_.0 (xyz), mass }                                  read by the 9 lanes of a group
//     chunk 10:       { wx[0], wx[1], wx[2], base node }                     read by the 9 lanes of a group
// so a lane reads 48 B per particle of its group (three LDS.128 = 12 wavefronts for 3 particles) and
// spends 16 math instructions per particle on its 3 nodes: val_i = chunk_jk + i A'_.0, w_i = wx[i] wy wz,
// acc_i += w_i (val_i, mass).  A run of particles of one cell that ends costs three vector reductions per
// lane (one per node, as before).
template <int MODEL>
__global__ void __launch_bounds__(kP2GWarps * 32, kP2GColsMinB) k_p2g_cols(ParticleStore S, const uint32_t* __restrict__ perm,
                                                                          uint32_t n, MaterialParams P, float4* __restrict__ grid,
                                                                          int* __restrict__ error_flag,
                                                                          const uint32_t* __restrict__ gone_keys) {
    constexpr int D = 3;
    constexpr int CH = 11;  // float4 chunks per particle; odd stride: conflict-free 128-bit stores (lane = slot)
    __shared__ float4 pkt[kP2GWarps][32 * CH];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t first = (blockIdx.x * kP2GWarps + warp) * 32u;
    if (first >= n) return;
    const int cnt = min(32u, n - first);
    const int n1 = P.n1;

    // ---- phase A (lane = particle) ----------------------------------------------------------
    // slab mode: a slot whose particle migrated away (or that was never filled) leaves an all-zero packet with node -1
    const bool gone = lane < cnt && gone_keys && __ldg(gone_keys + first + lane) == kKeyGone;
    if (gone_keys && !__any_sync(0xffffffffu, lane < cnt && !gone)) return;  // nothing to scatter in this warp
    if (gone) {
        float4* my = &pkt[warp][lane * CH];
#pragma unroll
        for (int q = 0; q < 10; ++q) my[q] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        my[10] = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(-1));
    } else if (lane < cnt) {
        PState<D> p;
        load_for_p2g<D>(S, perm ? __ldg(perm + first + lane) : first + lane, p);
        int base[D];
        float fx[D], w[D][3];
        if (!stencil_of<D>(p.x, P, base, fx, w)) atomicOr(error_flag, 1);
        const Mat<D> A = affine_matrix<D, MODEL>(p.F, p.C, p.Jp, p.mass, p.volume, P);
        float b[D], c0[D], c1[D], c2[D];
#pragma unroll
        for (int r = 0; r < D; ++r) {
            float afx = A(r, 0) * fx[0];
#pragma unroll
            for (int k = 1; k < D; ++k) afx = fmaf(A(r, k), fx[k], afx);
            b[r] = fmaf(-P.dx, afx, p.v[r] * p.mass);  // mass*v + A*((ijk-fx)*dx) = b + (dx*A)*ijk
            c0[r] = P.dx * A(r, 0), c1[r] = P.dx * A(r, 1), c2[r] = P.dx * A(r, 2);
        }
        float4* my = &pkt[warp][lane * CH];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float tj[D];
#pragma unroll
            for (int r = 0; r < D; ++r) tj[r] = (j == 0) ? b[r] : (j == 1) ? b[r] + c1[r] : fmaf(c1[r], 2.0f, b[r]);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                float q[D];
#pragma unroll
                for (int r = 0; r < D; ++r) q[r] = (k == 0) ? tj[r] : (k == 1) ? tj[r] + c2[r] : fmaf(c2[r], 2.0f, tj[r]);
                my[j * 3 + k] = make_float4(q[0], q[1], q[2], w[1][j] * w[2][k]);
            }
        }
        my[9] = make_float4(c0[0], c0[1], c0[2], p.mass);
        my[10] = make_float4(w[0][0], w[0][1], w[0][2], __int_as_float((base[0] * n1 + base[1]) * n1 + base[2]));
    }
    __syncwarp();

    // ---- phase B (lane = group g, stencil column (j,k); nodes i = 0,1,2 in registers) ---------
    if (lane >= 27) return;
    const int g = lane / 9, jk = lane - 9 * g, j = jk / 3, k = jk - 3 * j;
    const int s_begin = 11 * g;
    const int s_end = min((g == 2) ? 32 : s_begin + 11, cnt);
    if (s_begin >= s_end) return;
    const uint32_t plane = (uint32_t) (n1 * n1);
    const uint32_t col = (uint32_t) (j * n1 + k);  // this lane's column of the stencil, relative to the base node

    float2 acc01[3];
    float acc2[3], accm[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) acc01[i] = splat2(0.0f), acc2[i] = 0.0f, accm[i] = 0.0f;

    // one vector reduction per owned node; the accumulators are not cleared: the particle that opens the next run
    // overwrites them
    auto flush = [&](int node) {
        const uint32_t idx = (uint32_t) node + col;
#pragma unroll
        for (int i = 0; i < 3; ++i)
            red_add_f32x4(grid + (idx + (uint32_t) i * plane), make_float4(acc01[i].x, acc01[i].y, acc2[i], accm[i]));
    };

    const float4* pp = &pkt[warp][s_begin * CH];
    int cur = -1;
    for (int s = s_begin; s < s_end; ++s, pp += CH) {
        const float4 a = pp[jk], c = pp[9], x = pp[10];
        const int node = __float_as_int(x.w);
        const float w0 = x.x * a.w, w1 = x.y * a.w, w2 = x.z * a.w;
        const float2 q0 = make_float2(a.x, a.y), c01 = make_float2(c.x, c.y);
        const float2 q1 = __fadd2_rn(q0, c01), q2 = ffma2(c01, splat2(2.0f), q0);
        const float z1 = a.z + c.z, z2 = fmaf(c.z, 2.0f, a.z);
        if (node != cur) {  // a new cell starts: one vector reduction per owned node, the run restarts from this particle
            if (cur >= 0) flush(cur);
            cur = node;
            acc01[0] = fmul2(splat2(w0), q0), acc2[0] = w0 * a.z, accm[0] = w0 * c.w;
            acc01[1] = fmul2(splat2(w1), q1), acc2[1] = w1 * z1, accm[1] = w1 * c.w;
            acc01[2] = fmul2(splat2(w2), q2), acc2[2] = w2 * z2, accm[2] = w2 * c.w;
        } else {
            acc01[0] = ffma2(splat2(w0), q0, acc01[0]), acc2[0] = fmaf(w0, a.z, acc2[0]), accm[0] = fmaf(w0, c.w, accm[0]);
            acc01[1] = ffma2(splat2(w1), q1, acc01[1]), acc2[1] = fmaf(w1, z1, acc2[1]), accm[1] = fmaf(w1, c.w, accm[1]);
            acc01[2] = ffma2(splat2(w2), q2, acc01[2]), acc2[2] = fmaf(w2, z2, acc2[2]), accm[2] = fmaf(w2, c.w, accm[2]);
        }
    }
    if (cur >= 0) flush(cur);
}

// ---- K2 (variant 4, 3D): variant 3 with three particle STREAMS per warp -------------------------
//
// A vector reduction costs the SM ~1.3 cycles per LANE (REDG, B300_MICROARCH.md "Atomics"), so the 27
// lane-reductions of one flushed cell run cost as much as 35 ordinary instructions, and in variants
// 3-6 every warp pays three extra flushes because its 32 slots are cut into three group ranges.
// Here a warp owns 32*C consecutive slots, cut ONCE into three contiguous streams of 11C, 11C and 10C
// slots.  Per chunk, lanes [0,11) [11,22) [22,32) load the next 11/11/10 particles of streams 0/1/2
// (phase A), then group g walks them (phase B) with its run accumulators carried from chunk to chunk:
// the only flushes left are real cell changes plus three per warp, i.e. 3/C per 32 particles.
// Raw particle rows of one phase-A lane (108 B): loaded one chunk AHEAD, right before the node walk of the current
// chunk, so that their DRAM latency is covered by phase B instead of stalling the next phase A.
struct P2GRaw {
    float4 a0, a1, a2, a3, a4, a5;
    float c8;
    float2 mv;
};
__device__ __forceinline__ void p2g_load_raw(const ParticleStore& S, uint32_t i, P2GRaw& r) {
    r.a0 = ldg4(S.q[0] + i), r.a1 = ldg4(S.q[1] + i), r.a2 = ldg4(S.q[2] + i), r.a3 = ldg4(S.q[3] + i);
    r.a4 = ldg4(S.q[4] + i), r.a5 = ldg4(S.q[5] + i);
    r.c8 = __ldg(S.s + i);
    r.mv = __ldg(S.mv + i);
}
__device__ __forceinline__ void p2g_unpack_raw(const P2GRaw& r, PState<3>& p) {
    p.x[0] = r.a0.x, p.x[1] = r.a0.y, p.x[2] = r.a0.z, p.Jp = r.a0.w;
    p.F.m[0] = r.a1.x, p.F.m[1] = r.a1.y, p.F.m[2] = r.a1.z, p.F.m[3] = r.a1.w;
    p.F.m[4] = r.a2.x, p.F.m[5] = r.a2.y, p.F.m[6] = r.a2.z, p.F.m[7] = r.a2.w;
    p.F.m[8] = r.a3.x, p.v[0] = r.a3.y, p.v[1] = r.a3.z, p.v[2] = r.a3.w;
    p.C.m[0] = r.a4.x, p.C.m[1] = r.a4.y, p.C.m[2] = r.a4.z, p.C.m[3] = r.a4.w;
    p.C.m[4] = r.a5.x, p.C.m[5] = r.a5.y, p.C.m[6] = r.a5.z, p.C.m[7] = r.a5.w;
    p.C.m[8] = r.c8;
    p.mass = r.mv.x, p.volume = r.mv.y;
}

template <int MODEL, int MINB>
__global__ void __launch_bounds__(kP2GWarps * 32, MINB) k_p2g_streams(ParticleStore S, const uint32_t* __restrict__ perm,
                                                                             uint32_t n, MaterialParams P,
                                                                             float4* __restrict__ grid, int* __restrict__ error_flag,
                                                                             const uint32_t* __restrict__ gone_keys, int chunks) {
    constexpr int D = 3;
    constexpr int CH = 11;  // float4 chunks per particle (see variant 5)
    __shared__ float4 pkt[kP2GWarps][32 * CH];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t per_warp = 32u * (uint32_t) chunks;
    const uint32_t first = (blockIdx.x * kP2GWarps + warp) * per_warp;
    if (first >= n) return;
    const int total = (int) min(per_warp, n - first);
    const int n1 = P.n1;

    // phase A role: lane -> (stream, index in the stream's slice of this chunk)
    const int ga = (lane < 11) ? 0 : (lane < 22) ? 1 : 2;
    const int ia = lane - 11 * ga;
    const int wa = (ga == 2) ? 10 : 11;                            // particles of stream ga per chunk
    const int off_a = 11 * ga * chunks;                            // stream start: 0, 11C, 22C
    const int len_a = max(0, min(wa * chunks, total - off_a));     // stream length (short in the last warp)
    // phase B role: lane -> (stream, stencil column)
    const int g = lane / 9, jk = lane - 9 * g, j = jk / 3, k = jk - 3 * j;
    const int wb = (g == 2) ? 10 : 11;
    const int len_b = (lane < 27) ? max(0, min(wb * chunks, total - 11 * g * chunks)) : 0;
    const uint32_t plane = (uint32_t) (n1 * n1);
    const uint32_t col = (uint32_t) (j * n1 + k);  // this lane's stencil column, relative to the base node

    float2 acc01[3];
    float acc2[3], accm[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) acc01[i] = splat2(0.0f), acc2[i] = 0.0f, accm[i] = 0.0f;
    int cur = -1;

    // one vector reduction per owned node; the accumulators are NOT cleared here: the particle that opens the next
    // run overwrites them (see the walk below)
    auto flush = [&](int node) {
        const uint32_t idx = (uint32_t) node + col;  // 32-bit node indices (cells < 2^31): one IMAD.WIDE per address
#pragma unroll
        for (int i = 0; i < 3; ++i)
            red_add_f32x4(grid + (idx + (uint32_t) i * plane), make_float4(acc01[i].x, acc01[i].y, acc2[i], accm[i]));
    };

    // chunk 0's rows; afterwards every chunk's rows are requested while the previous chunk is walked
    // Slab mode: a slot whose particle migrated away, or that was never filled, carries kKeyGone; it leaves an all-zero
    // packet with node -1 (its rows are never read: they may be uninitialised memory) — a run is cut there, nothing is added.
    P2GRaw raw;
    uint32_t slot = 0;
    bool have = ia < len_a, gone = false;
    if (have) {
        slot = first + (uint32_t) (off_a + ia);
        gone = gone_keys && __ldg(gone_keys + slot) == kKeyGone;
        if (!gone) p2g_load_raw(S, perm ? __ldg(perm + slot) : slot, raw);
    }

    for (int c = 0; c < chunks; ++c) {
        // slab mode: the slots beyond the slab's particle count are all marked gone — nothing to scatter in such a chunk
        if (gone_keys && !__any_sync(0xffffffffu, have && !gone)) {
            const int pos = (c + 1) * wa + ia;
            have = (c + 1 < chunks) && pos < len_a;
            gone = false;
            if (have) {
                slot = first + (uint32_t) (off_a + pos);
                gone = __ldg(gone_keys + slot) == kKeyGone;
                if (!gone) p2g_load_raw(S, perm ? __ldg(perm + slot) : slot, raw);
            }
            continue;
        }
        // ---- phase A (lane = particle of stream ga) ------------------------------------------
        if (have && gone) {
            float4* my = &pkt[warp][lane * CH];
#pragma unroll
            for (int q = 0; q < 10; ++q) my[q] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            my[10] = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(-1));
        } else if (have) {
            PState<D> p;
            p2g_unpack_raw(raw, p);
            int base[D];
            float fx[D], w[D][3];
            if (!stencil_of<D>(p.x, P, base, fx, w)) atomicOr(error_flag, 1);
            const Mat<D> A = affine_matrix<D, MODEL>(p.F, p.C, p.Jp, p.mass, p.volume, P);
            float b[D], c0[D], c1[D], c2[D];
#pragma unroll
            for (int r = 0; r < D; ++r) {
                float afx = A(r, 0) * fx[0];
#pragma unroll
                for (int q = 1; q < D; ++q) afx = fmaf(A(r, q), fx[q], afx);
                b[r] = fmaf(-P.dx, afx, p.v[r] * p.mass);
                c0[r] = P.dx * A(r, 0), c1[r] = P.dx * A(r, 1), c2[r] = P.dx * A(r, 2);
            }
            float4* my = &pkt[warp][lane * CH];
#pragma unroll
            for (int jj = 0; jj < 3; ++jj) {
                float tj[D];
#pragma unroll
                for (int r = 0; r < D; ++r) tj[r] = (jj == 0) ? b[r] : (jj == 1) ? b[r] + c1[r] : fmaf(c1[r], 2.0f, b[r]);
#pragma unroll
                for (int kk = 0; kk < 3; ++kk) {
                    float q[D];
#pragma unroll
                    for (int r = 0; r < D; ++r) q[r] = (kk == 0) ? tj[r] : (kk == 1) ? tj[r] + c2[r] : fmaf(c2[r], 2.0f, tj[r]);
                    my[jj * 3 + kk] = make_float4(q[0], q[1], q[2], w[1][jj] * w[2][kk]);
                }
            }
            my[9] = make_float4(c0[0], c0[1], c0[2], p.mass);
            my[10] = make_float4(w[0][0], w[0][1], w[0][2], __int_as_float((base[0] * n1 + base[1]) * n1 + base[2]));
        }
        __syncwarp();

        // ---- next chunk's rows: in flight during the node walk below --------------------------------
        {
            const int pos = (c + 1) * wa + ia;
            have = (c + 1 < chunks) && pos < len_a;
            gone = false;
            if (have) {
                slot = first + (uint32_t) (off_a + pos);
                gone = gone_keys && __ldg(gone_keys + slot) == kKeyGone;
                if (!gone) p2g_load_raw(S, perm ? __ldg(perm + slot) : slot, raw);
            }
        }

        // ---- phase B (lane = stream g, stencil column (j,k); nodes i = 0,1,2 in registers) -----
        const int cnt_b = min(wb, len_b - c * wb);  // <= 0 for lanes >= 27 and for exhausted streams
        const float4* pp = &pkt[warp][(11 * g) * CH];
        for (int u = 0; u < cnt_b; ++u, pp += CH) {
            const float4 a = pp[jk], cc = pp[9], x = pp[10];
            const int node = __float_as_int(x.w);
            const float w0 = x.x * a.w, w1 = x.y * a.w, w2 = x.z * a.w;
            const float2 q0 = make_float2(a.x, a.y), c01 = make_float2(cc.x, cc.y);
            const float2 q1 = __fadd2_rn(q0, c01), q2 = ffma2(c01, splat2(2.0f), q0);
            const float z1 = a.z + cc.z, z2 = fmaf(cc.z, 2.0f, a.z);
            if (node != cur) {  // a new cell starts: one vector reduction per owned node, then the run restarts from this particle
                if (cur >= 0) flush(cur);
                cur = node;
                acc01[0] = fmul2(splat2(w0), q0), acc2[0] = w0 * a.z, accm[0] = w0 * cc.w;
                acc01[1] = fmul2(splat2(w1), q1), acc2[1] = w1 * z1, accm[1] = w1 * cc.w;
                acc01[2] = fmul2(splat2(w2), q2), acc2[2] = w2 * z2, accm[2] = w2 * cc.w;
            } else {
                acc01[0] = ffma2(splat2(w0), q0, acc01[0]), acc2[0] = fmaf(w0, a.z, acc2[0]), accm[0] = fmaf(w0, cc.w, accm[0]);
                acc01[1] = ffma2(splat2(w1), q1, acc01[1]), acc2[1] = fmaf(w1, z1, acc2[1]), accm[1] = fmaf(w1, cc.w, accm[1]);
                acc01[2] = ffma2(splat2(w2), q2, acc01[2]), acc2[2] = fmaf(w2, z2, acc2[2]), accm[2] = fmaf(w2, cc.w, accm[2]);
            }
        }
        __syncwarp();  // the packets are overwritten by the next chunk
    }
    if (cur >= 0) flush(cur);
}

// particles per warp = 32 * chunks; fewer chunks on small scenes so that the grid still fills the GPU
inline int p2g_stream_chunks(uint32_t n) {
    const uint32_t per_wave = 148u * 6u * kP2GWarps * 32u;  // slots of one resident wave at chunks = 1
    const uint32_t c = n / (2u * per_wave);
    return (int) (c < 1u ? 1u : (c > 4u ? 4u : c));
}

template <int D, int MODEL>
inline void launch_p2g_streams(const ParticleStore& S, const uint32_t* perm, uint32_t n, const MaterialParams& P,
                               float4* grid, int* error_flag, const uint32_t* gone_keys, cudaStream_t st, int chunks = 0) {
    if constexpr (D == 3) {
        if (chunks <= 0) chunks = p2g_stream_chunks(n);
        const unsigned per_block = kP2GWarps * 32 * chunks;
        const unsigned blocks = (n + per_block - 1) / per_block;
        // 5 CTAs per SM leave 96 registers per thread: the rows of the next chunk stay in registers during the node walk
        k_p2g_streams<MODEL, kP2GStreamsMinB><<<blocks, kP2GWarps * 32, 0, st>>>(S, perm, n, P, grid, error_flag, gone_keys, chunks);
    } else {  // 2D scenes are launch-bound (cfg1: 5 000 particles): the lane = node kernel stays
        const unsigned blocks = (n + kP2GWarps * 32 - 1) / (kP2GWarps * 32);
        k_p2g_cell<D, MODEL><<<blocks, kP2GWarps * 32, 0, st>>>(S, perm, n, P, grid, error_flag, gone_keys);
    }
}

template <int D, int MODEL>
inline void launch_p2g_cols(const ParticleStore& S, const uint32_t* perm, uint32_t n, const MaterialParams& P,
                            float4* grid, int* error_flag, const uint32_t* gone_keys, cudaStream_t st) {
    const unsigned blocks = (n + kP2GWarps * 32 - 1) / (kP2GWarps * 32);
    if constexpr (D == 3) {
        k_p2g_cols<MODEL><<<blocks, kP2GWarps * 32, 0, st>>>(S, perm, n, P, grid, error_flag, gone_keys);
    } else {
        k_p2g_cell<D, MODEL><<<blocks, kP2GWarps * 32, 0, st>>>(S, perm, n, P, grid, error_flag, gone_keys);
    }
}

template <int D, int MODEL>
inline void launch_p2g_cell(const ParticleStore& S, const uint32_t* perm, uint32_t n, const MaterialParams& P,
                            float4* grid, int* error_flag, const uint32_t* gone_keys, cudaStream_t st) {
    const unsigned blocks = (n + kP2GWarps * 32 - 1) / (kP2GWarps * 32);
    k_p2g_cell<D, MODEL><<<blocks, kP2GWarps * 32, 0, st>>>(S, perm, n, P, grid, error_flag, gone_keys);
}

}  // namespace nmpm
