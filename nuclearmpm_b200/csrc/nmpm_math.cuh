// Device math for the MLS-MPM step: fixed-size matrices in registers, determinant, the two-sided
// Jacobi SVD, polar decomposition and the fused affine (stress + APIC) matrix.
//
// Semantics follow the reference (file:line relative to the reference repo) including its quirks
// (SURVEY.md §2.3); the implementation is register-resident, fully unrolled CUDA.
//   diag<dim>      src/nclr_math.h:13-19   (Q1: only (0,0),(1,1) are set)
//   nclr_svd       src/nclr_math.h:50-74   (Eigen::JacobiSVD + det sign fix on index 2; Q3)
//   nclr_polar     src/nclr_math.h:76-98
//   stress/affine  src/nclr.h:313-337      (Q2: constant-filled volumetric term)
//   hardening      src/nclr.h:351-372      (Q8: exp in double)
#pragma once
#include <cfloat>
#include <cmath>
#include <cuda_runtime.h>

// The math below is device code.  It is ALSO compilable for the host (plain g++, see
// tests/cpp/hostmath.cpp) so that the CPU test-suite can check the algorithms against the oracle
// before any GPU time is spent; libnmpm.so never calls the host instantiations.
#if defined(__CUDACC__)
#define NMPM_HD __host__ __device__ __forceinline__
#else
#define NMPM_HD inline
#endif

namespace nmpm {

template <int D>
struct Mat {
    float m[D * D];  // column-major: (i,j) at i + j*D, like Eigen
    NMPM_HD float& operator()(int i, int j) { return m[i + j * D]; }
    NMPM_HD float operator()(int i, int j) const { return m[i + j * D]; }
};

template <int D>
NMPM_HD Mat<D> mat_mul(const Mat<D>& a, const Mat<D>& b) {
    Mat<D> r;
#pragma unroll
    for (int j = 0; j < D; ++j)
#pragma unroll
        for (int i = 0; i < D; ++i) {
            float acc = a(i, 0) * b(0, j);
#pragma unroll
            for (int k = 1; k < D; ++k) acc = fmaf(a(i, k), b(k, j), acc);
            r(i, j) = acc;
        }
    return r;
}

// a * b^T
template <int D>
NMPM_HD Mat<D> mat_mul_bt(const Mat<D>& a, const Mat<D>& b) {
    Mat<D> r;
#pragma unroll
    for (int j = 0; j < D; ++j)
#pragma unroll
        for (int i = 0; i < D; ++i) {
            float acc = a(i, 0) * b(j, 0);
#pragma unroll
            for (int k = 1; k < D; ++k) acc = fmaf(a(i, k), b(j, k), acc);
            r(i, j) = acc;
        }
    return r;
}

NMPM_HD float det(const Mat<2>& a) { return a(0, 0) * a(1, 1) - a(1, 0) * a(0, 1); }
// cofactors along row 0, the order Eigen's fixed-size determinant uses
NMPM_HD float det(const Mat<3>& a) {
    return a(0, 0) * (a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1)) - a(0, 1) * (a(1, 0) * a(2, 2) - a(1, 2) * a(2, 0)) +
           a(0, 2) * (a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0));
}

NMPM_HD float clampf(float v, float lo, float hi) { return (v < lo) ? lo : (hi < v) ? hi : v; }

// Plane rotation of two length-N "vectors" given as register arrays: x' = c x + s y ; y' = -s x + c y
#define NMPM_ROT(xv, yv, c, s)             \
    {                                      \
        const float _x = (xv), _y = (yv);  \
        (xv) = fmaf((c), _x, (s) * _y);    \
        (yv) = fmaf(-(s), _x, (c) * _y);   \
    }

// ~1 ulp reciprocal / reciprocal square root: MUFU approximation + one Newton-Raphson step, no
// slow-path branch (IEEE div/sqrt cost ~10 instructions each plus an FCHK branch, and the Jacobi
// rotation parameters are one long dependent chain of them).
// raw MUFU approximations (~2^-22 relative error); host builds use the exact value
NMPM_HD float rcp_approx(float x) {
#ifdef __CUDA_ARCH__
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}
NMPM_HD float rsqrt_approx(float x) {
#ifdef __CUDA_ARCH__
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return 1.0f / sqrtf(x);
#endif
}
NMPM_HD float sqrt_approx(float x) {
#ifdef __CUDA_ARCH__
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return sqrtf(x);
#endif
}
NMPM_HD float rcp_nr(float x) {
    const float r = rcp_approx(x);
    return fmaf(r, fmaf(-x, r, 1.0f), r);
}
NMPM_HD float rsqrt_nr(float x) {
    const float y = rsqrt_approx(x);
    const float h = 0.5f * y;
    return fmaf(h, fmaf(-x * y, y, 1.0f), y);  // y + 0.5 y (1 - x y^2)
}
#ifndef __CUDA_ARCH__
// host stand-ins for the round-to-nearest intrinsics (host test builds use -ffp-contract=off)
inline float nmpm_fmul_rn(float a, float b) { return a * b; }
inline float nmpm_fsub_rn(float a, float b) { return a - b; }
#else
__device__ __forceinline__ float nmpm_fmul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float nmpm_fsub_rn(float a, float b) { return __fsub_rn(a, b); }
#endif

// One two-sided Jacobi step on the (p,q) 2x2 block of W, accumulating U and V — the body of
// Eigen::JacobiSVD::compute's inner loop with internal::real_2x2_jacobi_svd and
// JacobiRotation::makeJacobi inlined (restated in oracle/eigen_standin/Eigen/Dense).
//
// The rotation parameters are the same closed forms written without quotients of quotients:
//   rot1 (symmetrising rotation, u = t/d):  s1 = 1/sqrt(1+u^2) = |d| h,  c1 = u/sqrt(1+u^2) = sgn(d) t h,
//        h = rsqrt(t^2 + d^2)
//   makeJacobi (tau = (x-z)/(2|y|), t = 1/(tau +- sqrt(tau^2+1))):
//        t = sgn * deno / (|x-z| + sqrt((x-z)^2 + deno^2)),  deno = 2|y|,  sgn = +1 if x-z > 0 else -1
//        c = rsqrt(t^2+1),  s = -sgn(y) t c
template <int N, int P, int Q>
NMPM_HD void jacobi_pq(Mat<N>& W, Mat<N>& U, Mat<N>& V, float& maxDiag, bool& finished) {
    const float thr = fmaxf(FLT_MIN, (2.0f * FLT_EPSILON) * maxDiag);
    if (fabsf(W(P, Q)) > thr || fabsf(W(Q, P)) > thr) {
        finished = false;
        float m00 = W(P, P), m01 = W(P, Q), m10 = W(Q, P), m11 = W(Q, Q);
        float c1 = 1.0f, s1 = 0.0f;
        const float t = m00 + m11;
        const float d = m10 - m01;
        if (fabsf(d) >= FLT_MIN) {
            const float r2 = fmaf(t, t, d * d);
            if (r2 > 1e-30f) {
                const float h = rsqrt_nr(r2);
                s1 = fabsf(d) * h;
                c1 = copysignf(t * h, t * d);
            } else {  // both tiny: the textbook form is safe from underflow of the squares
                const float u = t / d;
                const float tmp = sqrtf(fmaf(u, u, 1.0f));
                s1 = 1.0f / tmp;
                c1 = u / tmp;
            }
        }
        NMPM_ROT(m00, m10, c1, s1);
        NMPM_ROT(m01, m11, c1, s1);
        // makeJacobi(x = m00, y = m01, z = m11)
        float cr = 1.0f, sr = 0.0f;
        const float deno = 2.0f * fabsf(m01);
        if (deno >= FLT_MIN) {
            const float xz = m00 - m11;
            const float q2 = fmaf(xz, xz, deno * deno);
            float tt;
            if (q2 > 1e-30f) {
                const float root = q2 * rsqrt_nr(q2);
                tt = deno * rcp_nr(fabsf(xz) + root);
                tt = (xz > 0.0f) ? tt : -tt;
            } else {
                const float tau = xz / deno;
                const float w = sqrtf(fmaf(tau, tau, 1.0f));
                tt = 1.0f / ((tau > 0.0f) ? (tau + w) : (tau - w));
            }
            const float nn = rsqrt_nr(fmaf(tt, tt, 1.0f));
            sr = -copysignf(1.0f, m01) * tt * nn;
            cr = nn;
        }
        // j_left = rot1 * j_right^T
        const float cl = fmaf(c1, cr, s1 * sr);
        const float sl = fmaf(s1, cr, -c1 * sr);
#pragma unroll
        for (int k = 0; k < N; ++k) NMPM_ROT(W(P, k), W(Q, k), cl, sl);  // W.applyOnTheLeft(p,q,j_left)
#pragma unroll
        for (int k = 0; k < N; ++k) NMPM_ROT(U(k, P), U(k, Q), cl, sl);  // U.applyOnTheRight(p,q,j_left^T)
#pragma unroll
        for (int k = 0; k < N; ++k) NMPM_ROT(W(k, P), W(k, Q), cr, -sr);  // W.applyOnTheRight(p,q,j_right)
#pragma unroll
        for (int k = 0; k < N; ++k) NMPM_ROT(V(k, P), V(k, Q), cr, -sr);  // V.applyOnTheRight(p,q,j_right)
        maxDiag = fmaxf(maxDiag, fmaxf(fabsf(W(P, P)), fabsf(W(Q, Q))));
    }
}

#define NMPM_SWAPF(a, b) \
    {                    \
        float _t = (a);  \
        (a) = (b);       \
        (b) = _t;        \
    }

// Eigen::JacobiSVD<Matrix<float,N,N>>(a, ComputeFullU|ComputeFullV): a = U diag(sv) V^T,
// sv >= 0 descending (first maximum wins ties, zero tail left in place).
template <int N>
NMPM_HD void jacobi_svd(const Mat<N>& a, Mat<N>& U, float (&sv)[N], Mat<N>& V) {
    float scale = fabsf(a.m[0]);
#pragma unroll
    for (int k = 1; k < N * N; ++k) scale = fmaxf(scale, fabsf(a.m[k]));
    if (scale == 0.0f) scale = 1.0f;
    Mat<N> W;
    const float inv_scale = 1.0f / scale;  // one IEEE divide; W = a * (1/scale) is within 1 ulp of a / scale
#pragma unroll
    for (int k = 0; k < N * N; ++k) W.m[k] = a.m[k] * inv_scale;
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int i = 0; i < N; ++i) U(i, j) = V(i, j) = (i == j) ? 1.0f : 0.0f;
    float maxDiag = fabsf(W(0, 0));
#pragma unroll
    for (int i = 1; i < N; ++i) maxDiag = fmaxf(maxDiag, fabsf(W(i, i)));

    bool finished = false;
    // Eigen iterates until a sweep makes no rotation; converged inputs exit after one check sweep.
    // The cap only guards against non-finite input (where Eigen's comparisons are all false anyway).
    for (int sweep = 0; sweep < 32 && !finished; ++sweep) {
        finished = true;
        if constexpr (N == 2) {
            jacobi_pq<2, 1, 0>(W, U, V, maxDiag, finished);
        } else {
            jacobi_pq<3, 1, 0>(W, U, V, maxDiag, finished);
            jacobi_pq<3, 2, 0>(W, U, V, maxDiag, finished);
            jacobi_pq<3, 2, 1>(W, U, V, maxDiag, finished);
        }
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const float aii = W(i, i);
        sv[i] = fabsf(aii) * scale;
        if (aii < 0.0f) {
#pragma unroll
            for (int r = 0; r < N; ++r) U(r, i) = -U(r, i);
        }
    }
    // selection sort, descending, first maximum wins; stop at an all-zero tail
    if constexpr (N == 2) {
        if (sv[1] > sv[0]) {
            NMPM_SWAPF(sv[0], sv[1]);
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                NMPM_SWAPF(U(r, 0), U(r, 1));
                NMPM_SWAPF(V(r, 0), V(r, 1));
            }
        }
    } else {
        // i = 0: max of (sv0, sv1, sv2)
        int pos = 0;
        float best = sv[0];
        if (sv[1] > best) {
            best = sv[1];
            pos = 1;
        }
        if (sv[2] > best) {
            best = sv[2];
            pos = 2;
        }
        if (best != 0.0f) {
            if (pos == 1) {
                NMPM_SWAPF(sv[0], sv[1]);
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    NMPM_SWAPF(U(r, 0), U(r, 1));
                    NMPM_SWAPF(V(r, 0), V(r, 1));
                }
            } else if (pos == 2) {
                NMPM_SWAPF(sv[0], sv[2]);
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    NMPM_SWAPF(U(r, 0), U(r, 2));
                    NMPM_SWAPF(V(r, 0), V(r, 2));
                }
            }
            // i = 1: max of (sv1, sv2)
            if (sv[2] > sv[1]) {  // then best != 0 automatically
                NMPM_SWAPF(sv[1], sv[2]);
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    NMPM_SWAPF(U(r, 1), U(r, 2));
                    NMPM_SWAPF(V(r, 1), V(r, 2));
                }
            }
        }
    }
}

// nclr_svd<dim> (src/nclr_math.h:50-74): JacobiSVD, then force det U = det V = +1 by flipping
// column 2 and sigma_2.  In 2D the reference's hard-coded index 2 is out of bounds: no-op (Q3).
template <int D>
NMPM_HD void nclr_svd(const Mat<D>& a, Mat<D>& U, float (&sig)[D], Mat<D>& V) {
    jacobi_svd<D>(a, U, sig, V);
    if constexpr (D == 3) {
        if (det(U) < 0.0f) {
#pragma unroll
            for (int r = 0; r < 3; ++r) U(r, 2) = -U(r, 2);
            sig[2] = -sig[2];
        }
        if (det(V) < 0.0f) {
#pragma unroll
            for (int r = 0; r < 3; ++r) V(r, 2) = -V(r, 2);
            sig[2] = -sig[2];
        }
    }
}

// nclr_polar<dim> (src/nclr_math.h:76-98): rotation factor only (S is unused by every caller).
NMPM_HD Mat<2> nclr_polar_R(const Mat<2>& m) {
    const float x = m(0, 0) + m(1, 1);
    const float y = m(1, 0) - m(0, 1);
    const float scale = 1.0f / sqrtf(fmaf(x, x, y * y));
    const float c = x * scale, s = y * scale;
    Mat<2> R;
    R(0, 0) = c;
    R(0, 1) = -s;
    R(1, 0) = s;
    R(1, 1) = c;
    return R;
}
// ---------------------------------------------------------------------------------------------
// Fast 3x3 path: G = sum_i f(sigma_i) u_i v_i^T without forming U, sigma, V of the reference's
// two-sided JacobiSVD (which costs ~900 instructions per matrix and dominated both particle kernels,
// profiles/r01a_ncu_summary.md).
//
// One-sided (Hestenes) Jacobi: rotate pairs of COLUMNS of A until they are mutually orthogonal,
// accumulating the same rotations in V.  Then A V = [sigma_i u_i], i.e. column i IS sigma_i u_i, and
//     G = sum_i (f(sigma_i) / sigma_i) a_i v_i^T.
// The column of the smallest singular value is completed by a cross product, which reproduces the
// reference's det(U) = det(V) = +1 rule (src/nclr_math.h:63-71): V is a product of rotations
// (det +1), u_k := u_i x u_j for the cyclic order (i,j,k) makes det U = +1, and the signed
// sigma_k = u_k . a_k carries sign(det A) exactly like the reference's flipped sigma_2.  Exactly
// rank-2 input (3D jelly: third column of F is 0 forever, Q1) and tiny sigma_k (3D snow: F' has an
// O(dt C) third row) are therefore handled without dividing by sigma_k.
// Returns false when the matrix has (numerical) rank < 2 or is not finite: the caller falls back to
// the reference-shaped two-sided Jacobi above.
//   MODE 0: f = 1              (polar rotation R = U V^T)
//   MODE 1: f = clamp(., lo, hi)  (snow plasticity projection, src/nclr.h:239-247)
#if defined(NMPM_HOST_STATS) && !defined(__CUDA_ARCH__)
extern long nmpm_stat_rot, nmpm_stat_sweep, nmpm_stat_calls;
#define NMPM_STAT(x) (++(x))
#else
#define NMPM_STAT(x) ((void) 0)
#endif
// Packed fp32x2 helpers (FFMA2/FMUL2 on sm_100: two fp32 results per issue slot); plain pairs on the host.
NMPM_HD float2 f2(float a, float b) { return make_float2(a, b); }
NMPM_HD float2 f2_fma(float2 a, float2 b, float2 c) {
#ifdef __CUDA_ARCH__
    return __ffma2_rn(a, b, c);
#else
    return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
NMPM_HD float2 f2_mul(float2 a, float2 b) {
#ifdef __CUDA_ARCH__
    return __fmul2_rn(a, b);
#else
    return make_float2(a.x * b.x, a.y * b.y);
#endif
}

// Column j of the working pair (A V, V) as three packed words: (a0,a1) (a2,v0) (v1,v2).  A plane rotation
// of two columns is then 3 x (FMUL2, FFMA2, FMUL2, FFMA2) = 12 issue slots instead of 24 scalar ones.
struct HCol {
    float2 w[3];
};

#define NMPM_HESTENES_PAIR(P, Q)                                                                    \
    {                                                                                               \
        const float2 pr = f2_mul(col[P].w[0], col[Q].w[0]);                                         \
        const float gam = fmaf(col[P].w[1].x, col[Q].w[1].x, pr.x + pr.y);                          \
        const float gg = gam * gam, nn = fmaxf(nrm[P], floor2) * fmaxf(nrm[Q], floor2);             \
        if (gg > kTol2 * nn) {                                                                      \
            excess = fmaxf(excess, fmaf(-kSettle2, nn, gg));                                        \
            NMPM_STAT(nmpm_stat_rot);                                                               \
            const float d = nrm[Q] - nrm[P], g2 = gam + gam;                                        \
            const float r = sqrt_approx(fmaf(d, d, g2 * g2));                                       \
            float t = g2 * rcp_approx(fabsf(d) + r);                                                \
            t = (d < 0.0f) ? -t : t;                                                                \
            const float c = rsqrt_nr(fmaf(t, t, 1.0f)), sn = c * t;                                 \
            nrm[P] = fmaf(-t, gam, nrm[P]);                                                         \
            nrm[Q] = fmaf(t, gam, nrm[Q]);                                                          \
            const float2 c2 = f2(c, c), s2 = f2(sn, sn), ns2 = f2(-sn, -sn);                        \
            _Pragma("unroll") for (int e = 0; e < 3; ++e) {                                         \
                const float2 x = col[P].w[e], y = col[Q].w[e];                                      \
                col[P].w[e] = f2_fma(c2, x, f2_mul(ns2, y));                                        \
                col[Q].w[e] = f2_fma(s2, x, f2_mul(c2, y));                                         \
            }                                                                                       \
        }                                                                                           \
    }

// What the snow projection knows besides the projected matrix G = sum_j f_j u_j v_j^T: the clamped singular values and
// the RIGHT singular vectors v_j of the input.  (snow_project runs on the transpose, where they are the LEFT singular
// vectors l_j of the projected F — all the fused G2P+P2G kernel needs for the stress of the new state: with
// F = sum f_j l_j r_j^T and R = sum l_j r_j^T,  (F - R) F^T = sum (f_j - 1) f_j l_j l_j^T  and  det F = f_0 f_1 f_2.)
struct SvdFactors {
    float f[3];
    float v[3][3];  // v[j] = j-th right singular vector of the input
};

template <int MODE, bool FACTORS = false>
NMPM_HD bool svd3_recompose(const Mat<3>& A, float lo, float hi, Mat<3>& G, SvdFactors* sf = nullptr) {
    // |a_p . a_q| <= 4 eps |a_p| |a_q| counts as orthogonal
    constexpr float kTol2 = (4.0f * FLT_EPSILON) * (4.0f * FLT_EPSILON);
    // a sweep whose largest relative inner product was below 2e-4 leaves all of them below ~4e-8 (cyclic Jacobi
    // converges quadratically): no further sweep, not even a checking one
    constexpr float kSettle2 = 2.0e-4f * 2.0e-4f;
    HCol col[3];
    float nrm[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        col[j].w[0] = f2(A(0, j), A(1, j));
        col[j].w[1] = f2(A(2, j), (j == 0) ? 1.0f : 0.0f);
        col[j].w[2] = f2((j == 1) ? 1.0f : 0.0f, (j == 2) ? 1.0f : 0.0f);
        nrm[j] = fmaf(A(0, j), A(0, j), fmaf(A(1, j), A(1, j), A(2, j) * A(2, j)));
    }
    // A column that the rotations shrank far below sigma_max carries the absolute rounding noise of the large
    // ones (~eps sigma_max per entry), so its inner products cannot be driven below ~eps sigma_max |a_q|: norms
    // enter the test floored at (sigma_max/4)^2-ish.  (Its direction is not used anyway: the smallest column is
    // completed by a cross product below.)  Without the floor rank-deficient input spins to the sweep cap.
    const float floor2 = (1.0f / 32.0f) * (nrm[0] + nrm[1] + nrm[2]);
    float excess = 1.0f;  // > 0: some pair of the last sweep was still far from orthogonal
    NMPM_STAT(nmpm_stat_calls);
    for (int sweep = 0; sweep < 8 && excess > 0.0f; ++sweep) {
        excess = -1.0f;
        NMPM_STAT(nmpm_stat_sweep);
        NMPM_HESTENES_PAIR(0, 1)
        NMPM_HESTENES_PAIR(0, 2)
        NMPM_HESTENES_PAIR(1, 2)
    }
    float a[3][3], v[3][3];  // a[j] = column j of A V (= sigma_j u_j),  v[j] = column j of V
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        a[j][0] = col[j].w[0].x, a[j][1] = col[j].w[0].y, a[j][2] = col[j].w[1].x;
        v[j][0] = col[j].w[1].y, v[j][1] = col[j].w[2].x, v[j][2] = col[j].w[2].y;
    }
    // squared singular values from the final columns (the running values above only steer the sweeps)
#pragma unroll
    for (int j = 0; j < 3; ++j) nrm[j] = fmaf(a[j][0], a[j][0], fmaf(a[j][1], a[j][1], a[j][2] * a[j][2]));
    // k = column of the smallest singular value (the reference's sorted index 2); ties -> last
    const int k = (nrm[0] < nrm[1]) ? ((nrm[0] < nrm[2]) ? 0 : 2) : ((nrm[1] < nrm[2]) ? 1 : 2);
    const float big = fmaxf(nrm[0], fmaxf(nrm[1], nrm[2]));
    const float mid = (k == 0) ? fminf(nrm[1], nrm[2]) : (k == 1) ? fminf(nrm[0], nrm[2]) : fminf(nrm[0], nrm[1]);
    // rank >= 2 and finite (NaN fails every comparison); 1e-10 on the squares = 1e-5 on sigma_mid/sigma_max
    if (!(mid > 1e-10f * big) || !(big < 3.0e38f)) return false;
    float inv[3], u[3][3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        inv[j] = rsqrt_nr(fmaxf(nrm[j], 1e-37f));
#pragma unroll
        for (int e = 0; e < 3; ++e) u[j][e] = a[j][e] * inv[j];
    }
    // u_k = u_i x u_j with (i,j,k) cyclic
    float p[3], q[3];
#pragma unroll
    for (int e = 0; e < 3; ++e) {
        p[e] = (k == 0) ? u[1][e] : (k == 1) ? u[2][e] : u[0][e];
        q[e] = (k == 0) ? u[2][e] : (k == 1) ? u[0][e] : u[1][e];
    }
    const float ck[3] = {fmaf(p[1], q[2], -p[2] * q[1]), fmaf(p[2], q[0], -p[0] * q[2]), fmaf(p[0], q[1], -p[1] * q[0])};
    // G = sum_j (f_j u_j) v_j^T, rows 0/1 as one packed pair
    float2 fu01[3];
    float fu2[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const bool is_k = (j == k);
        float sg = nrm[j] * inv[j];  // sigma_j >= 0
        float uj[3] = {u[j][0], u[j][1], u[j][2]};
        if (is_k) {
            sg = fmaf(ck[0], a[j][0], fmaf(ck[1], a[j][1], ck[2] * a[j][2]));  // signed: u_k . (sigma_k u_k)
#pragma unroll
            for (int e = 0; e < 3; ++e) uj[e] = ck[e];
        }
        const float f = (MODE == 0) ? 1.0f : clampf(sg, lo, hi);
        if constexpr (FACTORS) {
            sf->f[j] = f;
#pragma unroll
            for (int e = 0; e < 3; ++e) sf->v[j][e] = v[j][e];
        }
        fu01[j] = f2_mul(f2(f, f), f2(uj[0], uj[1]));
        fu2[j] = f * uj[2];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float2 g01 = f2_mul(fu01[0], f2(v[0][c], v[0][c]));
        g01 = f2_fma(fu01[1], f2(v[1][c], v[1][c]), g01);
        g01 = f2_fma(fu01[2], f2(v[2][c], v[2][c]), g01);
        G(0, c) = g01.x;
        G(1, c) = g01.y;
        G(2, c) = fmaf(fu2[2], v[2][c], fmaf(fu2[1], v[1][c], fu2[0] * v[0][c]));
    }
    return true;
}

// rotation factor of the polar decomposition through the reference-shaped SVD (always valid)
NMPM_HD Mat<3> nclr_polar_R_jacobi(const Mat<3>& m) {
    Mat<3> U, V;
    float sig[3];
    nclr_svd<3>(m, U, sig, V);
    return mat_mul_bt<3>(U, V);
}
// Newton (Higham) iteration X <- (X + X^-T)/2 for the rotation factor of a well-conditioned matrix with
// det > 0: two steps take the singular values of snow's F (clamped to [0.975, 1.0045] by every G2P,
// src/nclr.h:241) to 1 within 1e-7; anything further from a rotation fails the test below.  ~110 instructions instead
// of ~400 for the one-sided Jacobi.  Accepted only if the result is orthogonal to 1e-6 (which bounds the
// distance to the true factor by the same amount); for det > 0 the polar rotation IS the reference's
// U V^T (no sign fix is active, src/nclr_math.h:63-71).
NMPM_HD bool polar3_newton(const Mat<3>& A, Mat<3>& R) {
    Mat<3> X = A;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        // cofactor matrix: column j = cross product of the other two columns (cyclic)
        Mat<3> Cf;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int a = (j + 1) % 3, b = (j + 2) % 3;
            Cf(0, j) = fmaf(X(1, a), X(2, b), -X(2, a) * X(1, b));
            Cf(1, j) = fmaf(X(2, a), X(0, b), -X(0, a) * X(2, b));
            Cf(2, j) = fmaf(X(0, a), X(1, b), -X(1, a) * X(0, b));
        }
        const float dt = fmaf(X(0, 0), Cf(0, 0), fmaf(X(1, 0), Cf(1, 0), X(2, 0) * Cf(2, 0)));
        if (!(dt > 0.5f && dt < 2.0f)) return false;  // not a near-rotation with det > 0 (also catches NaN)
        const float hinv = 0.5f * rcp_nr(dt);
#pragma unroll
        for (int k = 0; k < 9; ++k) X.m[k] = fmaf(hinv, Cf.m[k], 0.5f * X.m[k]);  // X^-T = cof(X)/det(X)
    }
    // accept iff X^T X = I to 1e-6
    float worst = 0.0f;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = i; j < 3; ++j) {
            const float g = fmaf(X(0, i), X(0, j), fmaf(X(1, i), X(1, j), X(2, i) * X(2, j)));
            worst = fmaxf(worst, fabsf(g - ((i == j) ? 1.0f : 0.0f)));
        }
    if (!(worst <= 1e-6f)) return false;
    R = X;
    return true;
}

NMPM_HD Mat<3> nclr_polar_R(const Mat<3>& m) {
    Mat<3> R;
    if (polar3_newton(m, R)) return R;
    if (svd3_recompose<0>(m, 0.0f, 0.0f, R)) return R;
    return nclr_polar_R_jacobi(m);
}
// snow: U clamp(sigma, lo, hi) V^T of nclr_svd(m)  (src/nclr.h:239-247)
NMPM_HD Mat<3> snow_project(const Mat<3>& m, float lo, float hi) {
    // The one-sided Jacobi runs on m^T (it orthogonalises the ROWS of m): U f(S) V^T of m^T is the transpose of
    // that of m, and the det(U) = det(V) = +1 / signed sigma_2 rule is symmetric in U and V.  Snow's
    // F' = (diag(1,1,0) + dt C) F has two near-orthonormal rows (F is within 2.5 % of a rotation) and an O(dt C)
    // third one, so F' F'^T is nearly diagonal while F'^T F' is a full rank-2 matrix: 2.0 sweeps / 3.0 rotations
    // per matrix instead of 3.1 / 6.2 (tools/hestenes_sweeps.cpp), and a warp pays the slowest lane.
    Mat<3> mt, Gt, G;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) mt(i, j) = m(j, i);
    if (svd3_recompose<1>(mt, lo, hi, Gt)) {
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) G(i, j) = Gt(j, i);
        return G;
    }
    Mat<3> U, V;
    float sig[3];
    nclr_svd<3>(m, U, sig, V);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float sj = clampf(sig[j], lo, hi);
#pragma unroll
        for (int i = 0; i < 3; ++i) U(i, j) *= sj;
    }
    return mat_mul_bt<3>(U, V);
}
// snow_project that also hands out the factors of the projected matrix (false: the fast path did not apply, `sf` is
// not filled and the caller must decompose G itself)
NMPM_HD bool snow_project_factors(const Mat<3>& m, float lo, float hi, Mat<3>& G, SvdFactors& sf) {
    Mat<3> mt, Gt;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) mt(i, j) = m(j, i);
    if (svd3_recompose<1, true>(mt, lo, hi, Gt, &sf)) {
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) G(i, j) = Gt(j, i);
        return true;
    }
    G = snow_project(m, lo, hi);
    return false;
}
NMPM_HD Mat<2> snow_project(const Mat<2>& m, float lo, float hi) {
    Mat<2> U, V;
    float sig[2];
    nclr_svd<2>(m, U, sig, V);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const float sj = clampf(sig[j], lo, hi);
#pragma unroll
        for (int i = 0; i < 2; ++i) U(i, j) *= sj;
    }
    return mat_mul_bt<2>(U, V);
}

struct MaterialParams {
    float mu_0, lambda_0;
    float dt, dx, inv_dx;
    float Dinv;       // 4*inv_dx*inv_dx            (src/nclr.h:325)
    float vmax;       // (float)(dx*0.9/dt)          (src/nclr.h:285)
    float dt_gravity; // dt*gravity                  (src/nclr.h:292)
    int res, n1;      // n1 = res+1 nodes per axis
    // Batch of independent 2D scenes advanced by the same launches (nmpm_create_batch; BASELINE config 5): the scenes'
    // grids are stacked along x (the slowest index), scene s owns node rows [s*n1, (s+1)*n1); a particle's scene comes
    // from its input index, and the Lame parameters are per scene.  scenes <= 1: a single scene, the pointers are unused.
    int scenes;
    const unsigned short* scene_of;  // scene of input index i (device)
    const float2* lame;              // {mu_0, lambda_0} of scene s (device)
};

// hardening (src/nclr.h:351-372).  The reference evaluates the snow factor exp(10(1-Jp)) in double and
// narrows it to float (Q8).  The device uses the fp32 expf (<= 2 ulp, and the fp32 argument adds
// <= 1e-7*|10(1-Jp)| <= 4e-7 relative for Jp >= 0.6): the FP64 pipe of this part is slow enough that the
// double version cost ~7 % of P2G's stall samples (profiles/r01e).  -DNMPM_EXACT_HARDENING restores it.
template <int MODEL>
NMPM_HD float hardening_e(float Jp) {
    if constexpr (MODEL == 0) {
#ifdef NMPM_EXACT_HARDENING
        return (float) exp(10.0 * (1.0 - (double) Jp));
#else
        return expf(10.0f * (1.0f - Jp));
#endif
    }
    if constexpr (MODEL == 1) return 0.3f;
    return 1.0f;
}

// first_piola_kirchoff_stress (src/nclr.h:313-337): returns -(dt*vol)*(Dinv*PF) + mass*C with
// PF = 2mu(F-R)F^T + lambda(J-1)J * ones(D,D)   (Q2)
template <int D, int MODEL>
NMPM_HD Mat<D> affine_matrix(const Mat<D>& F, const Mat<D>& C, float Jp, float mass, float volume,
                                                const MaterialParams& P) {
    const float e = hardening_e<MODEL>(Jp);
    const float mu = P.mu_0 * e, lambda = P.lambda_0 * e;
    const float J = det(F);
    const Mat<D> R = nclr_polar_R(F);
    Mat<D> lhs;
    const float two_mu = 2.0f * mu;
#pragma unroll
    for (int k = 0; k < D * D; ++k) lhs.m[k] = two_mu * (F.m[k] - R.m[k]);
    Mat<D> PF = mat_mul_bt<D>(lhs, F);
    const float cst = lambda * (J - 1.0f) * J;
    const float neg = -(P.dt * volume);
    Mat<D> A;
#pragma unroll
    for (int k = 0; k < D * D; ++k) A.m[k] = fmaf(neg, P.Dinv * (PF.m[k] + cst), mass * C.m[k]);
    return A;
}

// The same matrix for snow from the factors of F that the plasticity projection of the same G2P produced (see
// SvdFactors): PF = sum_j 2mu (f_j - 1) f_j l_j l_j^T + lambda (J - 1) J ones,  J = f_0 f_1 f_2 — no polar decomposition.
// f_j in [0.975, 1.0045] > 0, so R = sum l_j r_j^T is the rotation factor of F and det F = prod f_j (det U = det V = +1).
NMPM_HD Mat<3> affine_matrix_snow_factors(const SvdFactors& sf, const Mat<3>& C, float Jp, float mass, float volume,
                                          const MaterialParams& P) {
    const float e = hardening_e<0>(Jp);
    const float mu = P.mu_0 * e, lambda = P.lambda_0 * e;
    const float J = sf.f[0] * sf.f[1] * sf.f[2];
    const float cst = lambda * (J - 1.0f) * J;
    const float two_mu = 2.0f * mu;
    float g[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) g[j] = two_mu * (sf.f[j] - 1.0f) * sf.f[j];
    const float neg = -(P.dt * volume) * P.Dinv;
    Mat<3> A;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = r; c < 3; ++c) {
            float pf = g[0] * sf.v[0][r] * sf.v[0][c];
            pf = fmaf(g[1] * sf.v[1][r], sf.v[1][c], pf);
            pf = fmaf(g[2] * sf.v[2][r], sf.v[2][c], pf);
            const float t = neg * (pf + cst);
            A(r, c) = fmaf(mass, C(r, c), t);
            if (c != r) A(c, r) = fmaf(mass, C(c, r), t);
        }
    return A;
}

// per-axis quadratic B-spline stencil (src/nclr.h:115-127 == :172-183)
struct Stencil1 {
    int base;
    float fx;
    float w[3];
    bool ok;  // stencil nodes base..base+2 inside [0,res] and x finite
};
NMPM_HD Stencil1 stencil_axis(float x, float inv_dx, int res) {
    Stencil1 s;
    // __fmul_rn / __fsub_rn: no FMA contraction, so base and fx are bit-identical to the strict-FP
    // reference for any res, not only powers of two (Q10)
    const float g = nmpm_fmul_rn(x, inv_dx);
    const float t = nmpm_fsub_rn(g, 0.5f);
    s.base = (int) t;  // cast<int>: truncation toward zero (Q4)
    // base >= 0 && base+2 <= res, written on the float so that NaN / inf fail too.  (The reference
    // converts NaN to INT_MIN on x86 and throws from vector::at; CUDA's cvt would give 0.)
    s.ok = (t > -1.0f) && (t < (float) (res - 1));
    s.fx = nmpm_fsub_rn(g, (float) s.base);
    const float a = 1.5f - s.fx, b = s.fx - 1.0f, c = s.fx - 0.5f;
    s.w[0] = 0.5f * (a * a);
    s.w[1] = 0.75f - (b * b);
    s.w[2] = 0.5f * (c * c);
    return s;
}

}  // namespace nmpm
