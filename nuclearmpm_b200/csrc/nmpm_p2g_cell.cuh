// K2 (variant 2): P2G by warp-level segmented reduction over cell-sorted particles.
//
// Reference: p2g() + compute_fused_momentum() + first_piola_kirchoff_stress(), src/nclr.h:104-165,
// 313-337 — there a serial loop doing 3^dim read-modify-writes per particle.  Here:
//
//   phase A (lane = particle): each warp takes 32 consecutive slots of the cell-sorted store, loads
//     them with coalesced float4 reads, runs polar/SVD + stress in registers and leaves a 16-word
//     packet {fx, mass, mass*v, affine} + the linear index of the stencil's base node in shared memory.
//   phase B (lane = stencil node, 27 of 32 lanes in 3D): the warp walks its 32 packets in slot order;
//     every lane evaluates ITS node's weight and fused momentum (packet reads are shared-memory
//     broadcasts) and accumulates in registers.  Consecutive particles of the same cell form a
//     segment; at a segment boundary each lane issues ONE vector reduction
//     (RED.E.ADD.F32x4 {px,py,pz,m}) for its node.
//
// Global atomics per particle drop from 3^dim*(dim+1) scalar (or 3^dim vector) to 3^dim/ppc vector
// reductions (ppc = particles per cell; 3.4 at 8 ppc in 3D), and no shared-memory float atomics are
// used at all — on sm_100a those are CAS loops (ATOMS.CAST.SPIN), see DESIGN.md.
// Weights are evaluated per lane as fma(t*t, k_i, b_i) with t = fx - c_i, which is bit-identical to
// the reference's three formulas (src/nclr.h:124-127).
#pragma once
#include "nmpm_kernels.cuh"

namespace nmpm {

constexpr int kP2GWarps = 4;

template <int D, int MODEL>
__global__ void __launch_bounds__(kP2GWarps * 32) k_p2g_cell(ParticleStore S, uint32_t n, MaterialParams P,
                                                             float4* __restrict__ grid, int* __restrict__ error_flag) {
    constexpr int NODES = (D == 3) ? 27 : 9;
    constexpr int NPK = (D == 3) ? 4 : 3;  // float4 words per packet
    __shared__ float4 pk[kP2GWarps][NPK][32];
    __shared__ int node0[kP2GWarps][32];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t first = (blockIdx.x * kP2GWarps + warp) * 32u;
    if (first >= n) return;
    const int cnt = min(32u, n - first);
    const int n1 = P.n1;

    // ---- phase A ---------------------------------------------------------------------------
    if (lane < cnt) {
        PState<D> p;
        load_for_p2g<D>(S, first + lane, p);
        int base[D];
        float fx[D], w[D][3];
        if (!stencil_of<D>(p.x, P, base, fx, w)) atomicOr(error_flag, 1);
        const Mat<D> A = affine_matrix<D, MODEL>(p.F, p.C, p.Jp, p.mass, p.volume, P);
        if constexpr (D == 3) {
            pk[warp][0][lane] = make_float4(fx[0], fx[1], fx[2], p.mass);
            pk[warp][1][lane] = make_float4(p.v[0] * p.mass, p.v[1] * p.mass, p.v[2] * p.mass, A.m[0]);
            pk[warp][2][lane] = make_float4(A.m[1], A.m[2], A.m[3], A.m[4]);
            pk[warp][3][lane] = make_float4(A.m[5], A.m[6], A.m[7], A.m[8]);
            node0[warp][lane] = (base[0] * n1 + base[1]) * n1 + base[2];
        } else {
            pk[warp][0][lane] = make_float4(fx[0], fx[1], p.mass, 0.0f);
            pk[warp][1][lane] = make_float4(p.v[0] * p.mass, p.v[1] * p.mass, 0.0f, 0.0f);
            pk[warp][2][lane] = make_float4(A.m[0], A.m[1], A.m[2], A.m[3]);
            node0[warp][lane] = base[0] * n1 + base[1];
        }
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
    float fi[D], ci[D], ki[D], bi[D];
    int lane_off = 0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        fi[d] = (float) ijk[d];
        ci[d] = 1.5f - 0.5f * fi[d];                // 1.5, 1.0, 0.5
        ki[d] = (ijk[d] == 1) ? -1.0f : 0.5f;       // w1 = 0.75 - t^2 ; w0,w2 = 0.5 t^2
        bi[d] = (ijk[d] == 1) ? 0.75f : 0.0f;
        lane_off = lane_off * n1 + ijk[d];
    }

    float acc[D], acc_m = 0.0f;
#pragma unroll
    for (int d = 0; d < D; ++d) acc[d] = 0.0f;
    int cur = node0[warp][0];

    for (int j = 0; j < cnt; ++j) {
        const int nj = node0[warp][j];
        if (nj != cur) {  // warp-uniform: segment boundary
            red_add_f32x4(grid + (size_t) (cur + lane_off), node_pack<D>(acc, acc_m));
#pragma unroll
            for (int d = 0; d < D; ++d) acc[d] = 0.0f;
            acc_m = 0.0f;
            cur = nj;
        }
        float fx[D], mv[D], mass;
        Mat<D> A;
        if constexpr (D == 3) {
            const float4 a = pk[warp][0][j], b = pk[warp][1][j], c = pk[warp][2][j], e = pk[warp][3][j];
            fx[0] = a.x, fx[1] = a.y, fx[2] = a.z, mass = a.w;
            mv[0] = b.x, mv[1] = b.y, mv[2] = b.z;
            A.m[0] = b.w, A.m[1] = c.x, A.m[2] = c.y, A.m[3] = c.z, A.m[4] = c.w;
            A.m[5] = e.x, A.m[6] = e.y, A.m[7] = e.z, A.m[8] = e.w;
        } else {
            const float4 a = pk[warp][0][j], b = pk[warp][1][j], c = pk[warp][2][j];
            fx[0] = a.x, fx[1] = a.y, mass = a.z;
            mv[0] = b.x, mv[1] = b.y;
            A.m[0] = c.x, A.m[1] = c.y, A.m[2] = c.z, A.m[3] = c.w;
        }
        float weight = 1.0f, dpos[D];
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const float t = fx[d] - ci[d];
            const float wd = fmaf(t * t, ki[d], bi[d]);
            weight = (d == 0) ? wd : weight * wd;
            dpos[d] = (fi[d] - fx[d]) * P.dx;
        }
#pragma unroll
        for (int r = 0; r < D; ++r) {
            float ad = A(r, 0) * dpos[0];
#pragma unroll
            for (int k = 1; k < D; ++k) ad = fmaf(A(r, k), dpos[k], ad);
            acc[r] = fmaf(weight, mv[r] + ad, acc[r]);
        }
        acc_m = fmaf(weight, mass, acc_m);
    }
    red_add_f32x4(grid + (size_t) (cur + lane_off), node_pack<D>(acc, acc_m));
}

template <int D, int MODEL>
inline void launch_p2g_cell(const ParticleStore& S, uint32_t n, const MaterialParams& P, float4* grid, int* error_flag,
                            cudaStream_t st) {
    const unsigned blocks = (n + kP2GWarps * 32 - 1) / (kP2GWarps * 32);
    k_p2g_cell<D, MODEL><<<blocks, kP2GWarps * 32, 0, st>>>(S, n, P, grid, error_flag);
}

}  // namespace nmpm
