// K4+K2 fused (3D, single GPU): G2P of step n and P2G of step n+1 in ONE pass over the particles.
//
// Reference: advance() = p2g(); grid_op(); g2p();  (src/nclr.h:80-84).  The g2p() of step n leaves exactly the state
// the p2g() of step n+1 reads (src/nclr.h:104-165 reads x, v, F, C, Jp, mass, volume of every particle), and nothing
// happens in between.  With two grids the scatter of step n+1 can therefore run while the particle is still in
// registers: the kernel gathers from `grid` (the finished grid of step n), updates the particle (k_g2p_gather's
// body), writes the state back, and scatters mass / APIC momentum / stress of the NEW state into `grid_next`
// (k_p2g_cols' packet + node walk).  Per particle and step that removes the 108 B P2G read (a fused step moves
// 160 B per particle instead of 260), the second wait on DRAM latency, and one kernel boundary; the node walk
// (shared memory + reductions) of one warp overlaps the node gather (L1/L2 latency) of the others.
//
// The speculation is invisible at the API: grid() keeps returning the grid of the last advance() (the scatter goes to
// the OTHER buffer), the next step's P2G phase just swaps the buffers, an upload discards the speculative sums, and
// an out-of-grid position found by the scatter half is reported by the NEXT step (that is when the reference throws:
// its p2g() of step n+1, src/nclr.h:163) — error_flag[1] is promoted to error_flag[0] by that step.
//
// Order of the scatter: the slots as this G2P writes them, i.e. cell-sorted as of the last radix sort plus the
// warp-local re-grouping by the new cell (nmpm_kernels.cuh:g2p_finish) — a run of equal cells ends in one vector
// reduction per node as in k_p2g_cols.
#pragma once
#include "nmpm_kernels.cuh"
#include "nmpm_p2g_cell.cuh"

#ifndef NMPM_FUSED_MINB
#define NMPM_FUSED_MINB 7
#endif

namespace nmpm {

template <int MODEL, int MINB>
__global__ void __launch_bounds__(128, MINB) k_g2p_p2g(ParticleStore S, ParticleStore T, const uint32_t* __restrict__ perm,
                                                                 uint32_t n, MaterialParams P, const float4* __restrict__ grid,
                                                                 float4* __restrict__ grid_next, uint32_t* __restrict__ keys_out,
                                                                 int tiles_per_axis, int* __restrict__ error_flag,
                                                                 int* __restrict__ box_partial, int local_reorder) {
    constexpr int D = 3;
    constexpr int CH = 11;  // float4 chunks per particle packet (k_p2g_cols)
    __shared__ float4 pkt[4][32 * CH];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool mine = i < n;
    const unsigned live = __ballot_sync(0xffffffffu, mine);  // the low `cnt` lanes (only the last warp is short)
    if (live == 0u) {
        if (lane == 0) {
            int4* out = reinterpret_cast<int4*>(box_partial + (size_t) (i >> 5) * 8);
            out[0] = make_int4(0x7fffffff, 0x7fffffff, 0x7fffffff, (int) 0x80000000);
            out[1] = make_int4((int) 0x80000000, (int) 0x80000000, 0, 0);
        }
        return;
    }
    const int cnt = __popc(live);
    const int n1 = P.n1;

    if (mine) {
        // ---- G2P of step n (k_g2p_gather) -------------------------------------------------------------------
        const uint32_t src = perm ? __ldg(perm + i) : i;
        PState<D> p;
        load_for_g2p<D>(S, src, p);
        float vn[D];
        Mat<D> Cn;
        {
            int base[D];
            float fx[D], w[D][3];
            if (!stencil_of<D>(p.x, P, base, fx, w)) atomicOr(error_flag, 1);
            const GlobalNodes nodes{grid + ((size_t) (base[0] * n1 + base[1]) * n1 + base[2]), n1 * n1, n1};
            g2p_gather3(nodes, w, fx, 4.0f * P.inv_dx, vn, Cn);
        }
        const float2 mv = S.mv[src];  // coherent loads: the re-grouping rewrites these arrays in place
        // snow: the plasticity projection of this G2P already holds the SVD of the new F — the stress of step n+1 is built
        // from its factors (affine_matrix_snow_factors) instead of a second polar decomposition
        [[maybe_unused]] SvdFactors sf;
        [[maybe_unused]] bool have_sf = false;
        if constexpr (MODEL == 0) {
#pragma unroll
            for (int d = 0; d < D; ++d) {
                p.v[d] = vn[d];
                p.x[d] = fmaf(P.dt, vn[d], p.x[d]);  // advection (src/nclr.h:229)
            }
            p.C = Cn;
            Mat<D> M;  // F' = (diag<dim>(1) + dt*C) * F   (Q1)
#pragma unroll
            for (int q = 0; q < D * D; ++q) M.m[q] = P.dt * Cn.m[q];
            M(0, 0) += 1.0f;
            M(1, 1) += 1.0f;
            Mat<D> Fn = mat_mul<D>(M, p.F);
            const float old_J = det(Fn);
            have_sf = snow_project_factors(Fn, 0.975f, 1.0045f, p.F, sf);  // U clamp(sig) V^T (src/nclr.h:239-247)
            const float new_J = have_sf ? sf.f[0] * sf.f[1] * sf.f[2] : det(p.F);
            p.Jp = clampf(p.Jp * old_J / new_J, 0.6f, 20.0f);
        } else {
            g2p_update<D, MODEL>(p, Cn, vn, P);
        }

        // ---- bin the advected particle: cell key of step n+1, rank inside the warp ------------------------------
        int b[D];
        bool ok = true;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const Stencil1 s = stencil_axis(p.x[d], P.inv_dx, P.res);
            b[d] = s.base;
            ok = ok && s.ok;
        }
        const uint32_t key = ok ? cell_key<D>(b, tiles_per_axis) : kKeyOutOfGrid;
        if (!ok) atomicOr(error_flag + 1, 1);  // reported by the next step (the reference throws from ITS p2g)
#pragma unroll
        for (int d = 0; d < D; ++d) b[d] = min(max(b[d], 0), P.res - 2);
        bool moved = false;
        if (local_reorder) {
            const uint32_t prev = __shfl_up_sync(live, key, 1);
            moved = __any_sync(live, lane > 0 && prev > key);
        }
        uint32_t pid = 0;
        if (perm || moved) pid = S.id[src];
        int rank = lane;
        if (moved) {  // nmpm_kernels.cuh:g2p_finish — cells in the order of their first particle, five ballots
            const unsigned peers = __match_any_sync(live, key);
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
            rank = __popc(lt) + __popc(peers & ((1u << lane) - 1u));
            __syncwarp(live);  // T == S: every lane has read its old slot before any is overwritten
        }
        const uint32_t dst = (i - lane) + (uint32_t) rank;

        // ---- stress of the new state (k_p2g_cols phase A) ------------------------------------------------------
        Mat<D> A;
        if constexpr (MODEL == 0) {
            if (have_sf) A = affine_matrix_snow_factors(sf, p.C, p.Jp, mv.x, mv.y, P);
            else
                A = affine_matrix<D, MODEL>(p.F, p.C, p.Jp, mv.x, mv.y, P);
        } else {
            A = affine_matrix<D, MODEL>(p.F, p.C, p.Jp, mv.x, mv.y, P);
        }
        store_state<D>(T, dst, p);
        if (perm || moved) {
            T.mv[dst] = mv;
            T.id[dst] = pid;
        }
        if (keys_out) keys_out[dst] = key;
        box_partial_write<D>(box_partial, live, b, true, i >> 5);

        // ---- packet of step n+1's scatter, at the particle's NEW place in the warp ------------------------------
        float fx[D], w[D][3];
#pragma unroll
        for (int d = 0; d < D; ++d) {
            // same arithmetic as stencil_axis on the clamped base (a flagged particle scatters inside the grid)
            const float g = nmpm_fmul_rn(p.x[d], P.inv_dx);
            fx[d] = nmpm_fsub_rn(g, (float) b[d]);
            const float a = 1.5f - fx[d], bb = fx[d] - 1.0f, c = fx[d] - 0.5f;
            w[d][0] = 0.5f * (a * a), w[d][1] = 0.75f - (bb * bb), w[d][2] = 0.5f * (c * c);
        }
        float bv[D], c0[D], c1[D], c2[D];
#pragma unroll
        for (int r = 0; r < D; ++r) {
            float afx = A(r, 0) * fx[0];
#pragma unroll
            for (int q = 1; q < D; ++q) afx = fmaf(A(r, q), fx[q], afx);
            bv[r] = fmaf(-P.dx, afx, p.v[r] * mv.x);  // mass*v + A*((ijk-fx)*dx) = b + (dx*A)*ijk
            c0[r] = P.dx * A(r, 0), c1[r] = P.dx * A(r, 1), c2[r] = P.dx * A(r, 2);
        }
        float4* my = &pkt[warp][rank * CH];
#pragma unroll
        for (int jj = 0; jj < 3; ++jj) {
            float tj[D];
#pragma unroll
            for (int r = 0; r < D; ++r) tj[r] = (jj == 0) ? bv[r] : (jj == 1) ? bv[r] + c1[r] : fmaf(c1[r], 2.0f, bv[r]);
#pragma unroll
            for (int kk = 0; kk < 3; ++kk) {
                float q[D];
#pragma unroll
                for (int r = 0; r < D; ++r) q[r] = (kk == 0) ? tj[r] : (kk == 1) ? tj[r] + c2[r] : fmaf(c2[r], 2.0f, tj[r]);
                my[jj * 3 + kk] = make_float4(q[0], q[1], q[2], w[1][jj] * w[2][kk]);
            }
        }
        my[9] = make_float4(c0[0], c0[1], c0[2], mv.x);
        my[10] = make_float4(w[0][0], w[0][1], w[0][2], __int_as_float((b[0] * n1 + b[1]) * n1 + b[2]));
    }
    __syncwarp();

    // ---- P2G of step n+1 (k_p2g_cols phase B: lane = group g, stencil column (j,k); nodes i = 0,1,2 in registers) ----
    // (Cutting the thirds at run heads instead of at 11/22 saves two flushes per warp but costs as many issue slots as
    // it saves LSU cycles: measured +-0 on cfg4, gpurun r2s.)
    constexpr int s1 = 11, s2 = 22;
    if (lane >= 27) return;
    const int g = lane / 9, jk = lane - 9 * g, j = jk / 3, k = jk - 3 * j;
    const int s_begin = (g == 0) ? 0 : (g == 1) ? s1 : s2;
    const int s_end = min((g == 0) ? s1 : (g == 1) ? s2 : 32, cnt);
    if (s_begin >= s_end) return;
    const uint32_t plane = (uint32_t) (n1 * n1);
    const uint32_t col = (uint32_t) (j * n1 + k);

    float2 acc01[3];
    float acc2[3], accm[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) acc01[q] = splat2(0.0f), acc2[q] = 0.0f, accm[q] = 0.0f;
    auto flush = [&](int node) {
        const uint32_t idx = (uint32_t) node + col;
#pragma unroll
        for (int q = 0; q < 3; ++q)
            red_add_f32x4(grid_next + (idx + (uint32_t) q * plane), make_float4(acc01[q].x, acc01[q].y, acc2[q], accm[q]));
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
        if (node != cur) {
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

// error_flag[1] (found by the scatter half of the previous step's fused kernel) becomes this step's error
__global__ void k_promote_error(int* __restrict__ error_flag) {
    if (threadIdx.x == 0 && error_flag[1]) {
        atomicOr(error_flag, error_flag[1]);
        error_flag[1] = 0;
    }
}

template <int MODEL>
inline void launch_g2p_p2g(const ParticleStore& S, const ParticleStore& T, const uint32_t* perm, uint32_t n, const MaterialParams& P,
                           const float4* grid, float4* grid_next, uint32_t* keys_out, int tiles_per_axis, int* error_flag,
                           int* box_partial, int local_reorder, cudaStream_t st, int minb = NMPM_FUSED_MINB) {
    const unsigned blocks = (n + 127) / 128;
#define NMPM_FUSED_LAUNCH(B)                                                                                                \
    k_g2p_p2g<MODEL, B><<<blocks, 128, 0, st>>>(S, T, perm, n, P, grid, grid_next, keys_out, tiles_per_axis, error_flag, \
                                                box_partial, local_reorder)
    // CTAs per SM = register budget: 8 -> 64 registers, 7 -> 72, 6 -> 80, 5 -> 96 (NMPM_FUSED_MINB env: experiments).
    // Measured on cfg4 (fused launch): 5: 1.81 ms, 6: 1.76, 7: 1.725, 8: 1.76 — at 7 the 157 KB of packets leave the L1
    // a 92 KB carve-out instead of 60 KB at 8 (gpurun r2r, r3v)
    if (minb >= 8) NMPM_FUSED_LAUNCH(8);
    else if (minb == 7) NMPM_FUSED_LAUNCH(7);
    else if (minb == 5) NMPM_FUSED_LAUNCH(5);
    else NMPM_FUSED_LAUNCH(6);
#undef NMPM_FUSED_LAUNCH
}

}  // namespace nmpm
