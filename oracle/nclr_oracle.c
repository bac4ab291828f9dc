/* TEST INFRASTRUCTURE ONLY — see nclr_oracle.h for scope, pinning status and layout.
 * Every function cites the reference lines (relative to /root/reference/) it restates. */
#define _POSIX_C_SOURCE 200809L
#include "nclr_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

enum { kSnow = 0, kJelly = 1, kLiquid = 2 }; /* src/nclr.h:57-61 */
#define K_BOUNDARY 3                         /* src/nclr.h:66 */

struct nclr_oracle_sim {
    int dim, model, res;
    long n;
    float dt, dx, inv_dx, E, nu, gravity, mu_0, lambda_0;
    float *x, *v, *F, *C, *Jp, *mass, *volume;
    float *gv, *gm; /* grid momentum/velocity and mass; NULL until the first p2g (src/nclr.h:105-109) */
    long ncells;
};

static long g_oob_events = 0;
long nclr_oracle_oob_events(void) { return g_oob_events; }
void nclr_oracle_oob_reset(void) { g_oob_events = 0; }

/* ---------------------------------------------------------------- small dense helpers ----- */
#define M(a, D, i, j) ((a)[(i) + (j) * (D)])

/* Eigen fixed-size determinant (cofactors along row 0 for 3x3); src/nclr.h:246,248,318 */
static float det(const float *m, int D) {
    if (D == 2) return M(m, 2, 0, 0) * M(m, 2, 1, 1) - M(m, 2, 1, 0) * M(m, 2, 0, 1);
#define H3(a, b, c) (M(m, 3, 0, a) * (M(m, 3, 1, b) * M(m, 3, 2, c) - M(m, 3, 1, c) * M(m, 3, 2, b)))
    return H3(0, 1, 2) - H3(1, 0, 2) + H3(2, 0, 1);
#undef H3
}

/* out = a*b, k-ascending sums (Eigen coefficient-based product) */
static void matmul(const float *a, const float *b, float *out, int D) {
    float t[9];
    for (int j = 0; j < D; ++j)
        for (int i = 0; i < D; ++i) {
            float acc = M(a, D, i, 0) * M(b, D, 0, j);
            for (int k = 1; k < D; ++k) acc = acc + M(a, D, i, k) * M(b, D, k, j);
            M(t, D, i, j) = acc;
        }
    memcpy(out, t, sizeof(float) * (size_t) (D * D));
}
static void transpose(const float *a, float *out, int D) {
    float t[9];
    for (int j = 0; j < D; ++j)
        for (int i = 0; i < D; ++i) M(t, D, j, i) = M(a, D, i, j);
    memcpy(out, t, sizeof(float) * (size_t) (D * D));
}
/* diag<dim>(v): sets only (0,0) and (1,1) — Q1; src/nclr_math.h:13-19 */
static void diag_q1(float *m, int D, float v) {
    for (int k = 0; k < D * D; ++k) m[k] = 0.0f;
    M(m, D, 0, 0) = v;
    M(m, D, 1, 1) = v;
}
static float clampf(float v, float lo, float hi) { return (v < lo) ? lo : (hi < v) ? hi : v; } /* std::clamp */

/* ---------------------------------------------------------------- Jacobi SVD -------------- */
/* Eigen::JacobiSVD (square real, two-sided Jacobi) as restated in oracle/eigen_standin/Eigen/Dense;
 * call site src/nclr_math.h:57-61. */
static void rot_plane(float *x, int incx, float *y, int incy, int n, float c, float s) {
    if (c == 1.0f && s == 0.0f) return;
    for (int i = 0; i < n; ++i) {
        const float xi = x[i * incx], yi = y[i * incy];
        x[i * incx] = c * xi + s * yi;
        y[i * incy] = -s * xi + c * yi;
    }
}
static void jacobi_svd(const float *a, int N, float *U, float *sv, float *V) {
    const float precision = 2.0f * FLT_EPSILON;
    const float considerAsZero = FLT_MIN;
    float W[9];
    float scale = fabsf(a[0]);
    for (int k = 1; k < N * N; ++k)
        if (fabsf(a[k]) > scale) scale = fabsf(a[k]);
    if (scale == 0.0f) scale = 1.0f;
    for (int k = 0; k < N * N; ++k) W[k] = a[k] / scale;
    for (int j = 0; j < N; ++j)
        for (int i = 0; i < N; ++i) M(U, N, i, j) = M(V, N, i, j) = (i == j) ? 1.0f : 0.0f;
    float maxDiag = fabsf(M(W, N, 0, 0));
    for (int i = 1; i < N; ++i) maxDiag = fmaxf(maxDiag, fabsf(M(W, N, i, i)));

    int finished = 0;
    while (!finished) {
        finished = 1;
        for (int p = 1; p < N; ++p)
            for (int q = 0; q < p; ++q) {
                const float thr = fmaxf(considerAsZero, precision * maxDiag);
                if (fabsf(M(W, N, p, q)) > thr || fabsf(M(W, N, q, p)) > thr) {
                    finished = 0;
                    /* real_2x2_jacobi_svd */
                    float m[4];
                    M(m, 2, 0, 0) = M(W, N, p, p);
                    M(m, 2, 0, 1) = M(W, N, p, q);
                    M(m, 2, 1, 0) = M(W, N, q, p);
                    M(m, 2, 1, 1) = M(W, N, q, q);
                    float c1, s1;
                    const float t = M(m, 2, 0, 0) + M(m, 2, 1, 1);
                    const float d = M(m, 2, 1, 0) - M(m, 2, 0, 1);
                    if (fabsf(d) < considerAsZero) {
                        s1 = 0.0f;
                        c1 = 1.0f;
                    } else {
                        const float u = t / d;
                        const float tmp = sqrtf(1.0f + u * u);
                        s1 = 1.0f / tmp;
                        c1 = u / tmp;
                    }
                    rot_plane(&M(m, 2, 0, 0), 2, &M(m, 2, 1, 0), 2, 2, c1, s1);
                    /* makeJacobi(x=m00, y=m01, z=m11) */
                    float cr, sr;
                    {
                        const float x = M(m, 2, 0, 0), y = M(m, 2, 0, 1), z = M(m, 2, 1, 1);
                        const float deno = 2.0f * fabsf(y);
                        if (deno < FLT_MIN) {
                            cr = 1.0f;
                            sr = 0.0f;
                        } else {
                            const float tau = (x - z) / deno;
                            const float w = sqrtf(tau * tau + 1.0f);
                            float tt;
                            if (tau > 0.0f) tt = 1.0f / (tau + w);
                            else
                                tt = 1.0f / (tau - w);
                            const float sign_t = tt > 0.0f ? 1.0f : -1.0f;
                            const float nn = 1.0f / sqrtf(tt * tt + 1.0f);
                            sr = -sign_t * (y / fabsf(y)) * fabsf(tt) * nn;
                            cr = nn;
                        }
                    }
                    /* j_left = rot1 * j_right^T  (JacobiRotation product: c=c1*c2-s1*s2, s=c1*s2+s1*c2) */
                    const float c2 = cr, s2 = -sr;
                    const float cl = c1 * c2 - s1 * s2;
                    const float sl = c1 * s2 + s1 * c2;

                    rot_plane(&M(W, N, p, 0), N, &M(W, N, q, 0), N, N, cl, sl);  /* W.applyOnTheLeft(p,q,j_left) */
                    rot_plane(&M(U, N, 0, p), 1, &M(U, N, 0, q), 1, N, cl, sl);  /* U.applyOnTheRight(p,q,j_left^T) */
                    rot_plane(&M(W, N, 0, p), 1, &M(W, N, 0, q), 1, N, cr, -sr); /* W.applyOnTheRight(p,q,j_right) */
                    rot_plane(&M(V, N, 0, p), 1, &M(V, N, 0, q), 1, N, cr, -sr); /* V.applyOnTheRight(p,q,j_right) */

                    maxDiag = fmaxf(maxDiag, fmaxf(fabsf(M(W, N, p, p)), fabsf(M(W, N, q, q))));
                }
            }
    }
    for (int i = 0; i < N; ++i) {
        const float aii = M(W, N, i, i);
        sv[i] = fabsf(aii);
        if (aii < 0.0f)
            for (int r = 0; r < N; ++r) M(U, N, r, i) = -M(U, N, r, i);
    }
    for (int i = 0; i < N; ++i) sv[i] = sv[i] * scale;
    for (int i = 0; i < N; ++i) {
        int pos = 0;
        float best = sv[i];
        for (int k = 1; k < N - i; ++k)
            if (sv[i + k] > best) {
                best = sv[i + k];
                pos = k;
            }
        if (best == 0.0f) break;
        if (pos) {
            pos += i;
            float ts = sv[i];
            sv[i] = sv[pos];
            sv[pos] = ts;
            for (int r = 0; r < N; ++r) {
                float tu = M(U, N, r, i);
                M(U, N, r, i) = M(U, N, r, pos);
                M(U, N, r, pos) = tu;
                float tv = M(V, N, r, i);
                M(V, N, r, i) = M(V, N, r, pos);
                M(V, N, r, pos) = tv;
            }
        }
    }
}

/* nclr_svd<dim> — src/nclr_math.h:50-74.  The sign fix hard-codes index 2: effective for dim 3,
 * out-of-bounds for dim 2 where it is a counted no-op (Q3 policy). */
static void nclr_svd(const float *a, int D, float *U, float *sig, float *V) {
    float values[3];
    jacobi_svd(a, D, U, values, V);
    if (det(U, D) < 0.0f) {
        if (D == 3) {
            for (int r = 0; r < 3; ++r) M(U, 3, r, 2) = M(U, 3, r, 2) * -1.0f;
            values[2] = values[2] * -1.0f;
        } else {
            g_oob_events += 2;
        }
    }
    if (det(V, D) < 0.0f) {
        if (D == 3) {
            for (int r = 0; r < 3; ++r) M(V, 3, r, 2) = M(V, 3, r, 2) * -1.0f;
            values[2] = values[2] * -1.0f;
        } else {
            g_oob_events += 2;
        }
    }
    for (int j = 0; j < D; ++j)
        for (int i = 0; i < D; ++i) M(sig, D, i, j) = (i == j) ? values[i] : 0.0f; /* setIdentity + diagonal fill */
}

/* nclr_polar<dim> — src/nclr_math.h:76-98 */
static void nclr_polar(const float *m, int D, float *R, float *S) {
    if (D == 2) {
        const float x = M(m, 2, 0, 0) + M(m, 2, 1, 1);
        const float y = M(m, 2, 1, 0) - M(m, 2, 0, 1);
        const float scale = 1.0f / sqrtf(x * x + y * y);
        const float c = x * scale, s = y * scale;
        M(R, 2, 0, 0) = c;
        M(R, 2, 0, 1) = -s;
        M(R, 2, 1, 0) = s;
        M(R, 2, 1, 1) = c;
        if (S) {
            float Rt[4];
            transpose(R, Rt, 2);
            matmul(Rt, m, S, 2);
        }
    } else {
        float U[9], sig[9], V[9], Vt[9];
        nclr_svd(m, 3, U, sig, V);
        transpose(V, Vt, 3);
        matmul(U, Vt, R, 3);
        if (S) {
            float t[9];
            matmul(V, sig, t, 3);
            matmul(t, Vt, S, 3);
        }
    }
}

void nclr_oracle_svd(int dim, const float *a, float *U, float *sig, float *V) { nclr_svd(a, dim, U, sig, V); }
void nclr_oracle_polar(int dim, const float *m, float *R, float *S) { nclr_polar(m, dim, R, S); }

/* ---------------------------------------------------------------- constitutive model ------ */
/* hardening / snow_hardening / constant_hardening — src/nclr.h:351-372 (exp in double, Q8) */
static void hardening(const struct nclr_oracle_sim *s, float Jp, float *mu, float *lambda) {
    float e;
    if (s->model == kSnow) {
        const double ed = exp((double) 10.0f * (1.0 - (double) Jp));
        e = (float) ed;
    } else if (s->model == kJelly) {
        e = 0.3f;
    } else {
        e = 1.0f;
    }
    *mu = s->mu_0 * e;
    *lambda = s->lambda_0 * e;
}

/* first_piola_kirchoff_stress — src/nclr.h:313-337 */
static void affine_of(const struct nclr_oracle_sim *s, long p, float *A) {
    const int D = s->dim;
    const float *F = s->F + p * D * D, *C = s->C + p * D * D;
    float mu, lambda;
    hardening(s, s->Jp[p], &mu, &lambda);
    const float J = det(F, D);
    float r[9];
    nclr_polar(F, D, r, NULL);
    const float Dinv = 4 * s->inv_dx * s->inv_dx;
    const float two_mu = 2 * mu;
    float lhs[9], Ft[9], PF[9];
    for (int k = 0; k < D * D; ++k) lhs[k] = two_mu * (F[k] - r[k]);
    transpose(F, Ft, D);
    matmul(lhs, Ft, PF, D);
    const float cst = lambda * (J - 1) * J; /* constmat: ALL entries — Q2 */
    for (int k = 0; k < D * D; ++k) PF[k] = PF[k] + cst;
    const float neg = -(s->dt * s->volume[p]);
    for (int k = 0; k < D * D; ++k) {
        const float stress = neg * (Dinv * PF[k]);
        A[k] = stress + s->mass[p] * C[k];
    }
}
void nclr_oracle_affine(void *sv, long p, float *A) { affine_of((const struct nclr_oracle_sim *) sv, p, A); }

/* ---------------------------------------------------------------- stencil ----------------- */
/* base / fx / weights — src/nclr.h:115-127 and :172-183 (identical in both phases) */
static void stencil(const struct nclr_oracle_sim *s, const float *x, int *base, float *fx, float w[3][3]) {
    for (int d = 0; d < s->dim; ++d) {
        const float g = x[d] * s->inv_dx;
        base[d] = (int) (g - 0.5f); /* cast<int>: truncation (Q4) */
        fx[d] = g - (float) base[d];
        const float a = 1.5f - fx[d], b = fx[d] - 1.0f, c = fx[d] - 0.5f;
        w[0][d] = 0.5f * (a * a);
        w[1][d] = 0.75f - (b * b);
        w[2][d] = 0.5f * (c * c);
    }
}
/* vector::at() throws when index >= size (Q5).  A negative int index converts to a huge size_t. */
static int at_ok(const struct nclr_oracle_sim *s, long index) { return index >= 0 && index < s->ncells; }

/* ---------------------------------------------------------------- p2g --------------------- */
/* src/nclr.h:104-165 */
static int p2g(struct nclr_oracle_sim *s) {
    const int D = s->dim, n1 = s->res + 1;
    s->ncells = (D == 3) ? (long) n1 * n1 * n1 : (long) n1 * n1;
    if (!s->gv) {
        s->gv = (float *) malloc(sizeof(float) * (size_t) (s->ncells * D));
        s->gm = (float *) malloc(sizeof(float) * (size_t) s->ncells);
    }
    memset(s->gv, 0, sizeof(float) * (size_t) (s->ncells * D));
    memset(s->gm, 0, sizeof(float) * (size_t) s->ncells);

    for (long pp = 0; pp < s->n; ++pp) {
        const float *x = s->x + pp * D, *v = s->v + pp * D;
        int base[3] = {0, 0, 0};
        float fx[3] = {0, 0, 0}, w[3][3];
        stencil(s, x, base, fx, w);
        float A[9];
        affine_of(s, pp, A);
        const float mass = s->mass[pp];
        const int nk = (D == 3) ? 3 : 1;
        for (int ii = 0; ii < 3; ++ii)
            for (int jj = 0; jj < 3; ++jj)
                for (int kk = 0; kk < nk; ++kk) {
                    const int ijk[3] = {ii, jj, kk};
                    float dpos[3];
                    for (int d = 0; d < D; ++d) dpos[d] = ((float) ijk[d] - fx[d]) * s->dx;
                    float weight;
                    int index; /* int arithmetic like the reference (const auto index = int expr) */
                    if (D == 3) {
                        weight = w[ii][0] * w[jj][1] * w[kk][2];
                        index = ((base[0] + ii) * n1 * n1) + ((base[1] + jj) * n1) + (base[2] + kk);
                    } else {
                        weight = w[ii][0] * w[jj][1];
                        index = ((base[0] + ii) * n1) + (base[1] + jj);
                    }
                    if (!at_ok(s, index)) return 1;
                    /* compute_fused_momentum — src/nclr.h:160-165 */
                    for (int i = 0; i < D; ++i) {
                        const float mxv = v[i] * mass;
                        float ad = M(A, D, i, 0) * dpos[0];
                        for (int k = 1; k < D; ++k) ad = ad + M(A, D, i, k) * dpos[k];
                        s->gv[(long) index * D + i] = s->gv[(long) index * D + i] + weight * (mxv + ad);
                    }
                    s->gm[index] = s->gm[index] + weight * mass;
                }
    }
    return 0;
}

/* ---------------------------------------------------------------- grid_op ----------------- */
/* src/nclr.h:263-310 */
static void grid_op(struct nclr_oracle_sim *s) {
    const int D = s->dim, n1 = s->res + 1;
    const float allowed = (float) ((double) s->dx * 0.9 / (double) s->dt); /* :285, double then narrowed */
    const float g_dt = s->dt * s->gravity;
    for (long idx = 0; idx < s->ncells; ++idx) {
        float *vel = s->gv + idx * D;
        float *m = s->gm + idx;
        int c[3];
        if (D == 3) {
            c[0] = (int) (idx / ((long) n1 * n1));
            c[1] = (int) ((idx / n1) % n1);
            c[2] = (int) (idx % n1);
        } else {
            c[0] = (int) (idx / n1);
            c[1] = (int) (idx % n1);
            c[2] = 0;
        }
        if ((double) *m > 0.0) { /* grid_normalization :284-299 */
            for (int d = 0; d < D; ++d) vel[d] = vel[d] / *m;
            vel[1] = vel[1] + g_dt; /* gravity on component 1 (Q7) */
            for (int d = 0; d < D; ++d) vel[d] = clampf(vel[d], -allowed, allowed);
        }
        for (int d = 0; d < D; ++d) { /* sticky_boundary :301-310 (Q6) */
            const float fi = (float) c[d];
            if ((fi < K_BOUNDARY && vel[d] < 0) || (fi >= (n1) -K_BOUNDARY && vel[d] > 0)) {
                for (int e = 0; e < D; ++e) vel[e] = 0.0f;
                *m = 0.0f;
            }
        }
    }
}

/* ---------------------------------------------------------------- g2p --------------------- */
/* src/nclr.h:167-261 */
static int g2p(struct nclr_oracle_sim *s) {
    const int D = s->dim, n1 = s->res + 1;
    for (long pp = 0; pp < s->n; ++pp) {
        float *x = s->x + pp * D, *v = s->v + pp * D;
        float *F = s->F + pp * D * D, *C = s->C + pp * D * D;
        int base[3] = {0, 0, 0};
        float fx[3] = {0, 0, 0}, w[3][3];
        stencil(s, x, base, fx, w);
        for (int k = 0; k < D * D; ++k) C[k] = 0.0f;
        for (int d = 0; d < D; ++d) v[d] = 0.0f;
        const int nk = (D == 3) ? 3 : 1;
        const float four_inv_dx = 4 * s->inv_dx;
        for (int ii = 0; ii < 3; ++ii)
            for (int jj = 0; jj < 3; ++jj)
                for (int kk = 0; kk < nk; ++kk) {
                    const int ijk[3] = {ii, jj, kk};
                    float dpos[3];
                    for (int d = 0; d < D; ++d) dpos[d] = (float) ijk[d] - fx[d];
                    float weight;
                    int index;
                    if (D == 3) {
                        index = ((base[0] + ii) * n1 * n1) + ((base[1] + jj) * n1) + (base[2] + kk);
                        weight = w[ii][0] * w[jj][1] * w[kk][2];
                    } else {
                        index = ((base[0] + ii) * n1) + (base[1] + jj);
                        weight = w[ii][0] * w[jj][1];
                    }
                    if (!at_ok(s, index)) return 1;
                    const float *gv = s->gv + (long) index * D;
                    for (int i = 0; i < D; ++i) v[i] = v[i] + weight * gv[i];
                    for (int j = 0; j < D; ++j)
                        for (int i = 0; i < D; ++i) {
                            const float t = four_inv_dx * (weight * gv[i]);
                            M(C, D, i, j) = M(C, D, i, j) + t * dpos[j];
                        }
                }
        for (int d = 0; d < D; ++d) x[d] = x[d] + s->dt * v[d]; /* advection :229 */
        float Mx[9], Fn[9];
        diag_q1(Mx, D, 1.0f); /* Q1 */
        for (int k = 0; k < D * D; ++k) Mx[k] = Mx[k] + s->dt * C[k];
        matmul(Mx, F, Fn, D); /* :230 */

        if (s->model == kJelly) {
            memcpy(F, Fn, sizeof(float) * (size_t) (D * D));
        } else {
            float U[9], sig[9], V[9];
            nclr_svd(Fn, D, U, sig, V);
            if (s->model == kSnow) { /* :239-250 */
                for (int dd = 0; dd < D; ++dd)
                    M(sig, D, dd, dd) = clampf(M(sig, D, dd, dd), (float) (1.0 - 2.5e-2), (float) (1.0 + 4.5e-3));
                const float old_J = det(Fn, D);
                float t[9], Vt[9];
                matmul(U, sig, t, D);
                transpose(V, Vt, D);
                matmul(t, Vt, Fn, D);
                s->Jp[pp] = clampf(s->Jp[pp] * old_J / det(Fn, D), 0.6f, 20.0f);
                memcpy(F, Fn, sizeof(float) * (size_t) (D * D));
            } else { /* liquid :252-258 */
                double J = 1.0;
                for (int dd = 0; dd < D; ++dd) J *= (double) M(sig, D, dd, dd);
                diag_q1(F, D, 1.0f);
                M(F, D, 0, 0) = (float) J;
            }
        }
    }
    return 0;
}

/* ---------------------------------------------------------------- API --------------------- */
static float *dupf(const float *src, long count, float fill) {
    float *p = (float *) malloc(sizeof(float) * (size_t) (count > 0 ? count : 1));
    if (src) memcpy(p, src, sizeof(float) * (size_t) count);
    else
        for (long i = 0; i < count; ++i) p[i] = fill;
    return p;
}

/* ctor — src/nclr.h:74-78; Particle defaults — src/nclr.h:46-47 */
void *nclr_oracle_create(int dim, int model, int res, float dt, float E, float nu, float gravity, long n,
                         const float *x, const float *v, const float *F, const float *C, const float *Jp,
                         const float *mass, const float *volume) {
    struct nclr_oracle_sim *s = (struct nclr_oracle_sim *) calloc(1, sizeof(*s));
    s->dim = dim;
    s->model = model;
    s->res = res;
    s->n = n;
    s->dt = dt;
    s->dx = (float) (1.0 / res);
    s->inv_dx = 1 / s->dx;
    s->E = E;
    s->nu = nu;
    s->gravity = gravity;
    s->mu_0 = E / (2 * (1 + nu));
    s->lambda_0 = E * nu / ((1 + nu) * (1 - 2 * nu));
    const int D = dim;
    s->x = dupf(x, n * D, 0.0f);
    s->v = dupf(v, n * D, 0.0f);
    s->C = dupf(C, n * D * D, 0.0f);
    s->Jp = dupf(Jp, n, 1.0f);
    s->mass = dupf(mass, n, 1.0f);
    s->volume = dupf(volume, n, 1.0f);
    s->F = dupf(F, n * D * D, 0.0f);
    if (!F)
        for (long p = 0; p < n; ++p) diag_q1(s->F + p * D * D, D, 1.0f);
    return s;
}

void nclr_oracle_destroy(void *sv) {
    struct nclr_oracle_sim *s = (struct nclr_oracle_sim *) sv;
    if (!s) return;
    free(s->x);
    free(s->v);
    free(s->F);
    free(s->C);
    free(s->Jp);
    free(s->mass);
    free(s->volume);
    free(s->gv);
    free(s->gm);
    free(s);
}

int nclr_oracle_phase(void *sv, int phase) {
    struct nclr_oracle_sim *s = (struct nclr_oracle_sim *) sv;
    if (phase == 0) return p2g(s);
    if (phase == 1) {
        grid_op(s);
        return 0;
    }
    return g2p(s);
}

/* advance — src/nclr.h:80-84 */
int nclr_oracle_advance(void *sv, int nsteps) {
    struct nclr_oracle_sim *s = (struct nclr_oracle_sim *) sv;
    for (int k = 0; k < nsteps; ++k) {
        if (p2g(s)) return 1;
        grid_op(s);
        if (g2p(s)) return 1;
    }
    return 0;
}

double nclr_oracle_time_advance(void *sv, int nsteps) {
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    nclr_oracle_advance(sv, nsteps);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (double) (t1.tv_sec - t0.tv_sec) + 1e-9 * (double) (t1.tv_nsec - t0.tv_nsec);
}

long nclr_oracle_num_particles(void *sv) { return ((struct nclr_oracle_sim *) sv)->n; }

void nclr_oracle_get_particles(void *sv, float *x, float *v, float *F, float *C, float *Jp) {
    const struct nclr_oracle_sim *s = (const struct nclr_oracle_sim *) sv;
    const long n = s->n;
    const int D = s->dim;
    if (x) memcpy(x, s->x, sizeof(float) * (size_t) (n * D));
    if (v) memcpy(v, s->v, sizeof(float) * (size_t) (n * D));
    if (F) memcpy(F, s->F, sizeof(float) * (size_t) (n * D * D));
    if (C) memcpy(C, s->C, sizeof(float) * (size_t) (n * D * D));
    if (Jp) memcpy(Jp, s->Jp, sizeof(float) * (size_t) n);
}

long nclr_oracle_get_grid(void *sv, float *gv, float *gm) {
    const struct nclr_oracle_sim *s = (const struct nclr_oracle_sim *) sv;
    if (!s->gv) return 0;
    if (gv) memcpy(gv, s->gv, sizeof(float) * (size_t) (s->ncells * s->dim));
    if (gm) memcpy(gm, s->gm, sizeof(float) * (size_t) s->ncells);
    return s->ncells;
}

/* test hook for the slab (multi-GPU) protocol tests: overwrite the grid between phases */
long nclr_oracle_set_grid(void *sv, const float *gv, const float *gm) {
    struct nclr_oracle_sim *s = (struct nclr_oracle_sim *) sv;
    if (!s->gv) return 0;
    memcpy(s->gv, gv, sizeof(float) * (size_t) (s->ncells * s->dim));
    memcpy(s->gm, gm, sizeof(float) * (size_t) s->ncells);
    return s->ncells;
}

void nclr_oracle_lame(void *sv, float *mu0, float *lambda0) {
    const struct nclr_oracle_sim *s = (const struct nclr_oracle_sim *) sv;
    *mu0 = s->mu_0;
    *lambda0 = s->lambda_0;
}

/* cube<dim> — src/nclr_math.h:100-129, with Eigen's LinSpaced rule */
static float linspaced(int n, float low, float high, int i) {
    const int size1 = (n == 1) ? 1 : n - 1;
    const float step = (n == 1) ? 0.0f : (high - low) / (float) (n - 1);
    const int flip = fabsf(high) < fabsf(low);
    if (flip) return (i == 0) ? low : (high - (float) (size1 - i) * step);
    return (i == size1) ? high : (low + (float) i * step);
}
long nclr_oracle_cube(int dim, int res, float lo, float hi, float *out) {
    long cnt = 0;
    if (dim == 2) {
        for (int r = 0; r < res; ++r)
            for (int c = 0; c < res; ++c, ++cnt)
                if (out) {
                    out[2 * cnt] = linspaced(res, lo, hi, r);
                    out[2 * cnt + 1] = linspaced(res, lo, hi, c);
                }
    } else {
        for (int l = 0; l < res; ++l)
            for (int r = 0; r < res; ++r)
                for (int c = 0; c < res; ++c, ++cnt)
                    if (out) {
                        out[3 * cnt] = linspaced(res, lo, hi, l);
                        out[3 * cnt + 1] = linspaced(res, lo, hi, r);
                        out[3 * cnt + 2] = linspaced(res, lo, hi, c);
                    }
    }
    return cnt;
}

/* ---------------------------------------------------------------- binning oracle ---------- */
long nclr_oracle_cell_keys(int dim, int res, long n, const float *x, int mode, int tb, int32_t *base,
                           uint32_t *keys) {
    const float dx = (float) (1.0 / res);
    const float inv_dx = 1 / dx;
    const int n1 = res + 1;
    const int tile = 1 << tb;
    const int T = (n1 + tile - 1) / tile;
    const uint32_t msk = (uint32_t) (tile - 1);
    long bad = 0;
    for (long p = 0; p < n; ++p) {
        int b[3] = {0, 0, 0};
        int out = 0;
        for (int d = 0; d < dim; ++d) {
            b[d] = (int) (x[p * dim + d] * inv_dx - 0.5f);
            if (base) base[p * dim + d] = b[d];
            if (b[d] < 0 || b[d] + 2 > res) out = 1;
        }
        bad += out;
        if (!keys) continue;
        uint32_t key;
        if (out) {
            key = 0xFFFFFFFEu; /* sorted after every valid key (0xFFFFFFFF = migrated away, slab mode); the solver raises before using it */
        } else if (mode == 0) {
            key = (dim == 3) ? (uint32_t) ((b[0] * n1 + b[1]) * n1 + b[2]) : (uint32_t) (b[0] * n1 + b[1]);
        } else {
            uint32_t t = (uint32_t) ((b[0] >> tb) * T + (b[1] >> tb));
            uint32_t c = (((uint32_t) b[0] & msk) << tb) | ((uint32_t) b[1] & msk);
            if (dim == 3) {
                t = t * (uint32_t) T + (uint32_t) (b[2] >> tb);
                c = (c << tb) | ((uint32_t) b[2] & msk);
            }
            key = (t << (dim * tb)) | c;
        }
        keys[p] = key;
    }
    return bad;
}

void nclr_oracle_stable_sort(long n, const uint32_t *keys, uint32_t *perm) {
    /* LSD counting sort, 4 passes of 8 bits: stable by construction */
    uint32_t *a = (uint32_t *) malloc(sizeof(uint32_t) * (size_t) (n > 0 ? n : 1));
    uint32_t *b = (uint32_t *) malloc(sizeof(uint32_t) * (size_t) (n > 0 ? n : 1));
    for (long i = 0; i < n; ++i) a[i] = (uint32_t) i;
    for (int pass = 0; pass < 4; ++pass) {
        long count[257];
        memset(count, 0, sizeof(count));
        const int sh = pass * 8;
        for (long i = 0; i < n; ++i) count[((keys[a[i]] >> sh) & 0xFFu) + 1]++;
        for (int k = 0; k < 256; ++k) count[k + 1] += count[k];
        for (long i = 0; i < n; ++i) b[count[(keys[a[i]] >> sh) & 0xFFu]++] = a[i];
        uint32_t *t = a;
        a = b;
        b = t;
    }
    memcpy(perm, a, sizeof(uint32_t) * (size_t) n);
    free(a);
    free(b);
}
