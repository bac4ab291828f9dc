// Particle store (device SoA, float4-packed) and grid node layout.
//
// Replaces the reference's AoS std::vector<Particle<dim>> (src/nclr.h:20-48,101) and
// std::vector<Cell<dim>> (src/nclr.h:50-55,100).  Every array is indexed by device SLOT; slots are
// kept sorted by cell key (nmpm_sort.cuh) and `id[slot]` is the particle's input-order index, so
// particles() can always be returned in input order (src/nclr.h:86).
//
//  3D  q0={x0,x1,x2,Jp} q1={F0..F3} q2={F4..F7} q3={F8,v0,v1,v2} q4={C0..C3} q5={C4..C7} s={C8}
//  2D  q0={x0,x1,v0,v1} q1={F0..F3} q2={C0..C3}                                          s={Jp}
//  mv={mass,volume} (constant per particle), id (uint32)
// F and C keep the reference's column-major coefficient order.
//
// Bytes moved per particle: P2G reads 108 B (3D) / 60 B (2D) — exactly the algorithmic figure of
// SURVEY.md §8(d); G2P reads 64/36 B and writes 100/52 B.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "nmpm_math.cuh"

namespace nmpm {

template <int D>
struct StoreTraits;
template <>
struct StoreTraits<3> {
    static constexpr int NQ = 6;
};
template <>
struct StoreTraits<2> {
    static constexpr int NQ = 3;
};

struct ParticleStore {
    float4* q[6];
    float* s;
    float2* mv;
    uint32_t* id;
};

template <int D>
struct PState {
    float x[D], v[D];
    Mat<D> F, C;
    float Jp, mass, volume;
};

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

// ---- loads ------------------------------------------------------------------------------
template <int D>
__device__ __forceinline__ void load_for_p2g(const ParticleStore& S, size_t i, PState<D>& p);
template <>
__device__ __forceinline__ void load_for_p2g<3>(const ParticleStore& S, size_t i, PState<3>& p) {
    const float4 a0 = ldg4(S.q[0] + i), a1 = ldg4(S.q[1] + i), a2 = ldg4(S.q[2] + i), a3 = ldg4(S.q[3] + i),
                 a4 = ldg4(S.q[4] + i), a5 = ldg4(S.q[5] + i);
    const float c8 = __ldg(S.s + i);
    const float2 mv = __ldg(S.mv + i);
    p.x[0] = a0.x, p.x[1] = a0.y, p.x[2] = a0.z, p.Jp = a0.w;
    p.F.m[0] = a1.x, p.F.m[1] = a1.y, p.F.m[2] = a1.z, p.F.m[3] = a1.w;
    p.F.m[4] = a2.x, p.F.m[5] = a2.y, p.F.m[6] = a2.z, p.F.m[7] = a2.w;
    p.F.m[8] = a3.x, p.v[0] = a3.y, p.v[1] = a3.z, p.v[2] = a3.w;
    p.C.m[0] = a4.x, p.C.m[1] = a4.y, p.C.m[2] = a4.z, p.C.m[3] = a4.w;
    p.C.m[4] = a5.x, p.C.m[5] = a5.y, p.C.m[6] = a5.z, p.C.m[7] = a5.w;
    p.C.m[8] = c8;
    p.mass = mv.x, p.volume = mv.y;
}
template <>
__device__ __forceinline__ void load_for_p2g<2>(const ParticleStore& S, size_t i, PState<2>& p) {
    const float4 a0 = ldg4(S.q[0] + i), a1 = ldg4(S.q[1] + i), a2 = ldg4(S.q[2] + i);
    const float jp = __ldg(S.s + i);
    const float2 mv = __ldg(S.mv + i);
    p.x[0] = a0.x, p.x[1] = a0.y, p.v[0] = a0.z, p.v[1] = a0.w;
    p.F.m[0] = a1.x, p.F.m[1] = a1.y, p.F.m[2] = a1.z, p.F.m[3] = a1.w;
    p.C.m[0] = a2.x, p.C.m[1] = a2.y, p.C.m[2] = a2.z, p.C.m[3] = a2.w;
    p.Jp = jp, p.mass = mv.x, p.volume = mv.y;
}

// G2P needs x, F, Jp only (v and C are overwritten: src/nclr.h:185-186).  Plain (coherent) loads, not
// ld.global.nc: the in-place G2P writes other lanes' slots of these same arrays later in the kernel.
__device__ __forceinline__ float4 ld4(const float4* p) { return *p; }
template <int D>
__device__ __forceinline__ void load_for_g2p(const ParticleStore& S, size_t i, PState<D>& p);
template <>
__device__ __forceinline__ void load_for_g2p<3>(const ParticleStore& S, size_t i, PState<3>& p) {
    const float4 a0 = ld4(S.q[0] + i), a1 = ld4(S.q[1] + i), a2 = ld4(S.q[2] + i), a3 = ld4(S.q[3] + i);
    p.x[0] = a0.x, p.x[1] = a0.y, p.x[2] = a0.z, p.Jp = a0.w;
    p.F.m[0] = a1.x, p.F.m[1] = a1.y, p.F.m[2] = a1.z, p.F.m[3] = a1.w;
    p.F.m[4] = a2.x, p.F.m[5] = a2.y, p.F.m[6] = a2.z, p.F.m[7] = a2.w;
    p.F.m[8] = a3.x;
}
template <>
__device__ __forceinline__ void load_for_g2p<2>(const ParticleStore& S, size_t i, PState<2>& p) {
    const float4 a0 = ld4(S.q[0] + i), a1 = ld4(S.q[1] + i);
    p.x[0] = a0.x, p.x[1] = a0.y;
    p.F.m[0] = a1.x, p.F.m[1] = a1.y, p.F.m[2] = a1.z, p.F.m[3] = a1.w;
    p.Jp = S.s[i];
}

// ---- stores (x, v, F, C, Jp; mass/volume/id are never rewritten by the step) ---------------
template <int D>
__device__ __forceinline__ void store_state(const ParticleStore& S, size_t i, const PState<D>& p);
template <>
__device__ __forceinline__ void store_state<3>(const ParticleStore& S, size_t i, const PState<3>& p) {
    S.q[0][i] = make_float4(p.x[0], p.x[1], p.x[2], p.Jp);
    S.q[1][i] = make_float4(p.F.m[0], p.F.m[1], p.F.m[2], p.F.m[3]);
    S.q[2][i] = make_float4(p.F.m[4], p.F.m[5], p.F.m[6], p.F.m[7]);
    S.q[3][i] = make_float4(p.F.m[8], p.v[0], p.v[1], p.v[2]);
    S.q[4][i] = make_float4(p.C.m[0], p.C.m[1], p.C.m[2], p.C.m[3]);
    S.q[5][i] = make_float4(p.C.m[4], p.C.m[5], p.C.m[6], p.C.m[7]);
    S.s[i] = p.C.m[8];
}
template <>
__device__ __forceinline__ void store_state<2>(const ParticleStore& S, size_t i, const PState<2>& p) {
    S.q[0][i] = make_float4(p.x[0], p.x[1], p.v[0], p.v[1]);
    S.q[1][i] = make_float4(p.F.m[0], p.F.m[1], p.F.m[2], p.F.m[3]);
    S.q[2][i] = make_float4(p.C.m[0], p.C.m[1], p.C.m[2], p.C.m[3]);
    S.s[i] = p.Jp;
}

template <int D>
__device__ __forceinline__ void load_position(const ParticleStore& S, size_t i, float (&x)[D]) {
    const float4 a0 = ldg4(S.q[0] + i);
    x[0] = a0.x, x[1] = a0.y;
    if constexpr (D == 3) x[2] = a0.z;
}

// ---- cell key (K0) ------------------------------------------------------------------------
// Blocked key, mode 1 of oracle/nclr_oracle.h: tiles of 2^TB cells per axis, x slowest.
//   tile = ((bx>>TB)*T + (by>>TB)) [*T + (bz>>TB)],  T = ceil(n1 / 2^TB)
//   key  = tile << (D*TB) | cell-in-tile (x slowest)
constexpr int kTileBits = 2;
template <int D>
__device__ __host__ __forceinline__ uint32_t cell_key(const int (&b)[D], int tiles_per_axis) {
    constexpr uint32_t msk = (1u << kTileBits) - 1u;
    uint32_t t = (uint32_t) ((b[0] >> kTileBits) * tiles_per_axis + (b[1] >> kTileBits));
    uint32_t c = (((uint32_t) b[0] & msk) << kTileBits) | ((uint32_t) b[1] & msk);
    if constexpr (D == 3) {
        t = t * (uint32_t) tiles_per_axis + (uint32_t) (b[2] >> kTileBits);
        c = (c << kTileBits) | ((uint32_t) b[2] & msk);
    }
    return (t << (D * kTileBits)) | c;
}

}  // namespace nmpm
