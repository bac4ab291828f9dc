// TEST INFRASTRUCTURE ONLY: host instantiation of the device math header (csrc/nmpm_math.cuh) so that
// the CPU suite can check the register algorithms (one-sided Jacobi recompose, polar, snow projection,
// stress) against the oracle without a GPU.  Never linked into libnmpm.so.
#include "../../nuclearmpm_b200/csrc/nmpm_math.cuh"

using namespace nmpm;

extern "C" {
// returns the number of matrices that took the fast (one-sided Jacobi) path
long hm_polar3(const float* A, float* R, long n) {
    long fast = 0;
    for (long i = 0; i < n; ++i) {
        Mat<3> a, r;
        for (int k = 0; k < 9; ++k) a.m[k] = A[i * 9 + k];
        if (polar3_newton(a, r)) fast += 1000000;          // millions digit: Newton path
        else if (svd3_recompose<0>(a, 0.0f, 0.0f, r)) ++fast;  // units: one-sided Jacobi path
        else
            r = nclr_polar_R_jacobi(a);
        for (int k = 0; k < 9; ++k) R[i * 9 + k] = r.m[k];
    }
    return fast;
}
void hm_polar3_jacobi(const float* A, float* R, long n) {
    for (long i = 0; i < n; ++i) {
        Mat<3> a;
        for (int k = 0; k < 9; ++k) a.m[k] = A[i * 9 + k];
        const Mat<3> r = nclr_polar_R_jacobi(a);
        for (int k = 0; k < 9; ++k) R[i * 9 + k] = r.m[k];
    }
}
long hm_snow_project3(const float* A, float lo, float hi, float* G, long n) {
    long fast = 0;
    for (long i = 0; i < n; ++i) {
        Mat<3> a, g;
        for (int k = 0; k < 9; ++k) a.m[k] = A[i * 9 + k];
        if (svd3_recompose<1>(a, lo, hi, g)) ++fast;
        g = snow_project(a, lo, hi);
        for (int k = 0; k < 9; ++k) G[i * 9 + k] = g.m[k];
    }
    return fast;
}
void hm_svd3(const float* A, float* U, float* S, float* V, long n) {
    for (long i = 0; i < n; ++i) {
        Mat<3> a, u, v;
        float sig[3];
        for (int k = 0; k < 9; ++k) a.m[k] = A[i * 9 + k];
        nclr_svd<3>(a, u, sig, v);
        for (int k = 0; k < 9; ++k) U[i * 9 + k] = u.m[k], V[i * 9 + k] = v.m[k];
        for (int k = 0; k < 3; ++k) S[i * 3 + k] = sig[k];
    }
}
void hm_affine3(int model, const float* F, const float* C, const float* Jp, const float* mass, const float* volume, float mu_0,
                float lambda_0, float dt, float inv_dx, float* A, long n) {
    MaterialParams P{};
    P.mu_0 = mu_0, P.lambda_0 = lambda_0, P.dt = dt, P.inv_dx = inv_dx, P.dx = 1.0f / inv_dx;
    P.Dinv = 4 * inv_dx * inv_dx;
    for (long i = 0; i < n; ++i) {
        Mat<3> f, c, a;
        for (int k = 0; k < 9; ++k) f.m[k] = F[i * 9 + k], c.m[k] = C[i * 9 + k];
        if (model == 0) a = affine_matrix<3, 0>(f, c, Jp[i], mass[i], volume[i], P);
        else if (model == 1)
            a = affine_matrix<3, 1>(f, c, Jp[i], mass[i], volume[i], P);
        else
            a = affine_matrix<3, 2>(f, c, Jp[i], mass[i], volume[i], P);
        for (int k = 0; k < 9; ++k) A[i * 9 + k] = a.m[k];
    }
}
}
