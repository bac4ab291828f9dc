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
#include <cuda_runtime.h>

namespace nmpm {

template <int D>
struct Mat {
    float m[D * D];  // column-major: (i,j) at i + j*D, like Eigen
    __device__ __forceinline__ float& operator()(int i, int j) { return m[i + j * D]; }
    __device__ __forceinline__ float operator()(int i, int j) const { return m[i + j * D]; }
};

template <int D>
__device__ __forceinline__ Mat<D> mat_mul(const Mat<D>& a, const Mat<D>& b) {
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
__device__ __forceinline__ Mat<D> mat_mul_bt(const Mat<D>& a, const Mat<D>& b) {
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

__device__ __forceinline__ float det(const Mat<2>& a) { return a(0, 0) * a(1, 1) - a(1, 0) * a(0, 1); }
// cofactors along row 0, the order Eigen's fixed-size determinant uses
__device__ __forceinline__ float det(const Mat<3>& a) {
    return a(0, 0) * (a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1)) - a(0, 1) * (a(1, 0) * a(2, 2) - a(1, 2) * a(2, 0)) +
           a(0, 2) * (a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0));
}

__device__ __forceinline__ float clampf(float v, float lo, float hi) { return (v < lo) ? lo : (hi < v) ? hi : v; }

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
__device__ __forceinline__ float rcp_nr(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return fmaf(r, fmaf(-x, r, 1.0f), r);
}
__device__ __forceinline__ float rsqrt_nr(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    const float h = 0.5f * y;
    return fmaf(h, fmaf(-x * y, y, 1.0f), y);  // y + 0.5 y (1 - x y^2)
}

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
__device__ __forceinline__ void jacobi_pq(Mat<N>& W, Mat<N>& U, Mat<N>& V, float& maxDiag, bool& finished) {
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
__device__ __forceinline__ void jacobi_svd(const Mat<N>& a, Mat<N>& U, float (&sv)[N], Mat<N>& V) {
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
__device__ __forceinline__ void nclr_svd(const Mat<D>& a, Mat<D>& U, float (&sig)[D], Mat<D>& V) {
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
__device__ __forceinline__ Mat<2> nclr_polar_R(const Mat<2>& m) {
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
__device__ __forceinline__ Mat<3> nclr_polar_R(const Mat<3>& m) {
    Mat<3> U, V;
    float sig[3];
    nclr_svd<3>(m, U, sig, V);
    return mat_mul_bt<3>(U, V);
}

struct MaterialParams {
    float mu_0, lambda_0;
    float dt, dx, inv_dx;
    float Dinv;       // 4*inv_dx*inv_dx            (src/nclr.h:325)
    float vmax;       // (float)(dx*0.9/dt)          (src/nclr.h:285)
    float dt_gravity; // dt*gravity                  (src/nclr.h:292)
    int res, n1;      // n1 = res+1 nodes per axis
};

// hardening (src/nclr.h:351-372): snow exp(10(1-Jp)) evaluated in double then narrowed (Q8)
template <int MODEL>
__device__ __forceinline__ float hardening_e(float Jp) {
    if constexpr (MODEL == 0) return (float) exp(10.0 * (1.0 - (double) Jp));
    if constexpr (MODEL == 1) return 0.3f;
    return 1.0f;
}

// first_piola_kirchoff_stress (src/nclr.h:313-337): returns -(dt*vol)*(Dinv*PF) + mass*C with
// PF = 2mu(F-R)F^T + lambda(J-1)J * ones(D,D)   (Q2)
template <int D, int MODEL>
__device__ __forceinline__ Mat<D> affine_matrix(const Mat<D>& F, const Mat<D>& C, float Jp, float mass, float volume,
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

// per-axis quadratic B-spline stencil (src/nclr.h:115-127 == :172-183)
struct Stencil1 {
    int base;
    float fx;
    float w[3];
    bool ok;  // stencil nodes base..base+2 inside [0,res] and x finite
};
__device__ __forceinline__ Stencil1 stencil_axis(float x, float inv_dx, int res) {
    Stencil1 s;
    // __fmul_rn / __fsub_rn: no FMA contraction, so base and fx are bit-identical to the strict-FP
    // reference for any res, not only powers of two (Q10)
    const float g = __fmul_rn(x, inv_dx);
    const float t = __fsub_rn(g, 0.5f);
    s.base = (int) t;  // cast<int>: truncation toward zero (Q4)
    // base >= 0 && base+2 <= res, written on the float so that NaN / inf fail too.  (The reference
    // converts NaN to INT_MIN on x86 and throws from vector::at; CUDA's cvt would give 0.)
    s.ok = (t > -1.0f) && (t < (float) (res - 1));
    s.fx = __fsub_rn(g, (float) s.base);
    const float a = 1.5f - s.fx, b = s.fx - 1.0f, c = s.fx - 0.5f;
    s.w[0] = 0.5f * (a * a);
    s.w[1] = 0.75f - (b * b);
    s.w[2] = 0.5f * (c * c);
    return s;
}

}  // namespace nmpm
