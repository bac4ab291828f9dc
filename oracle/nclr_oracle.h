/* TEST INFRASTRUCTURE ONLY — CPU restatement ("port") of the reference's advance() hot path.
 *
 * Follows /root/reference/src/nclr.h:74-84,104-165,167-261,263-310,313-372 and
 * src/nclr_math.h:11-19,40-129 operation by operation (same float expression order, same
 * double-precision islands, same quirks Q1-Q9 of SURVEY.md §2.3), so that a strict-FP build
 * is BIT-EXACT with the reference header compiled against oracle/eigen_standin
 * (oracle/_ref/libnclr_ref_strict.so); tests/test_oracle.py asserts that.
 *
 * Pinning status: the reference ships no golden vectors or solver tests and its
 * arithmetic lives in Eigen3 (un-vendored, version unpinned, absent here).  The pin is therefore
 * "outputs of the reference header itself run in this container against the Eigen stand-in";
 * at the Eigen boundary (JacobiSVD, determinant, product order) parity is UNPINNED — see DESIGN.md.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libnmpm.so) never does.
 *
 * Layout: x,v n*dim; F,C n*dim*dim per-particle column-major (M(i,j) at [i+j*dim]); Jp,mass,volume n.
 * Grid node index x*n1+y (2D), (x*n1+y)*n1+z (3D), n1 = res+1.
 */
#ifndef NCLR_ORACLE_H
#define NCLR_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nclr_oracle_sim nclr_oracle_sim;

void *nclr_oracle_create(int dim, int model, int res, float dt, float E, float nu, float gravity, long n,
                         const float *x, const float *v, const float *F, const float *C, const float *Jp,
                         const float *mass, const float *volume);
void nclr_oracle_destroy(void *sim);
/* 0 = ok, 1 = a particle's stencil left the grid (reference: std::out_of_range from .at(), Q5) */
int nclr_oracle_advance(void *sim, int nsteps);
int nclr_oracle_phase(void *sim, int phase); /* 0 p2g, 1 grid_op, 2 g2p */
double nclr_oracle_time_advance(void *sim, int nsteps);
long nclr_oracle_num_particles(void *sim);
void nclr_oracle_get_particles(void *sim, float *x, float *v, float *F, float *C, float *Jp);
long nclr_oracle_get_grid(void *sim, float *gv, float *gm);
long nclr_oracle_set_grid(void *sim, const float *gv, const float *gm); /* port only: slab protocol tests */
void nclr_oracle_lame(void *sim, float *mu0, float *lambda0);
void nclr_oracle_svd(int dim, const float *a, float *U, float *sig, float *V);
void nclr_oracle_polar(int dim, const float *m, float *R, float *S);
void nclr_oracle_affine(void *sim, long p, float *A);
long nclr_oracle_cube(int dim, int res, float lo, float hi, float *out);
long nclr_oracle_oob_events(void);
void nclr_oracle_oob_reset(void);

/* ---- binning oracle (K0; nothing in the reference sorts — the key is derived from base_coord,
 *      src/nclr.h:115, and the node index formula :141-142,152) -------------------------------
 * base[d] = (int)(x[d]*inv_dx - 0.5f).  key modes:
 *   0: linear node index of base:  bx*n1+by  /  (bx*n1+by)*n1+bz
 *   1: blocked: (tile index, cell-in-tile) with tiles of 2^tb cells per axis, x slowest:
 *        tile = ((bx>>tb)*T + (by>>tb)) [*T + (bz>>tb)],  T = ceil(n1 / 2^tb)
 *        key  = tile << (dim*tb) | ((bx&m)<<(tb*(dim-1)) | (by&m)<<(tb*(dim-2)) [| bz&m])
 * returns number of particles whose 3-wide stencil leaves [0,res] on any axis (would throw). */
long nclr_oracle_cell_keys(int dim, int res, long n, const float *x, int mode, int tb, int32_t *base,
                           uint32_t *keys);
/* stable ascending sort of keys: perm[i] = original index of the i-th particle in sorted order */
void nclr_oracle_stable_sort(long n, const uint32_t *keys, uint32_t *perm);

#ifdef __cplusplus
}
#endif
#endif
