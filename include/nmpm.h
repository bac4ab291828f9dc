/* nmpm.h — C-ABI of the B200-native MLS-MPM step ("libnmpm.so").
 *
 * This is the drop-in boundary for the reference's per-step hot path.  The reference has no FFI
 * table: its boundary is the C++ class surface of src/nclr.h (MPMSimulation<dim>, src/nclr.h:63-87)
 * consumed at src/example.cpp:46,53,65,77 and src/solver.cpp:46-59,100-104,184-188.  Each entry
 * point below names the reference interface it replaces; include/nclr.h (same namespace, types and
 * signatures as the reference header) is the binding a maintainer uses — see INTEGRATION.md.
 *
 * Plain pointers and sizes only.  All `const float*` particle inputs are HOST pointers in the
 * interchange layout:  x,v : n*dim ;  F,C : n*dim*dim, per particle column-major (Eigen storage,
 * M(i,j) at [i + j*dim]) ;  Jp,mass,volume : n.  A NULL input means the reference's default
 * (src/nclr.h:46-47: v=0, F=diag<dim>(1) [Q1: diag(1,1,0) in 3D], C=0, Jp=1, mass=volume=1).
 * Grid outputs: node index x*n1+y (2D) / (x*n1+y)*n1+z (3D), n1=res+1 (src/nclr.h:141-142,152).
 *
 * There is no CPU fallback: every compute entry point fails with NMPM_ERR_NO_DEVICE / NMPM_ERR_CUDA
 * when no sm_100 device is usable.
 */
#ifndef NMPM_H
#define NMPM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NMPM_VERSION 100

typedef struct nmpm_sim *nmpm_handle;

enum nmpm_status {
    NMPM_OK = 0,
    NMPM_ERR_INVALID = 1,     /* bad argument */
    NMPM_ERR_CUDA = 2,        /* CUDA runtime error; see nmpm_last_error */
    NMPM_ERR_OUT_OF_GRID = 3, /* a particle's 3-wide stencil left [0,res]: the reference throws
                                 std::out_of_range from .at() (src/nclr.h:113,163,199; Q5) */
    NMPM_ERR_NO_DEVICE = 4
};

/* enum class MaterialModel — src/nclr.h:57-61 */
enum nmpm_model { NMPM_SNOW = 0, NMPM_JELLY = 1, NMPM_LIQUID = 2 };

/* phases of advance() — src/nclr.h:80-84 */
enum nmpm_phase { NMPM_PHASE_P2G = 0, NMPM_PHASE_GRID_OP = 1, NMPM_PHASE_G2P = 2 };

/* tunables that do not change results beyond float summation order */
typedef struct nmpm_options {
    int device;       /* CUDA device ordinal (default 0) */
    int sort_every;   /* re-bin + radix-sort particles by cell key every k steps (default 4; 0 = never).
                         Between sorts particles keep their last cell order; only float summation order changes */
    int p2g_variant;  /* 0 = auto (3D with binning: 3, or 4 from 8 Mi particles; 2D with binning: 2; no binning: 1),
                         1 = per-particle float4 REDs,
                         2 = cell-segmented, lane = stencil node,
                         3 = cell-segmented, three 9-lane groups per warp, lane = stencil column (3D; 2D runs 2),
                         4 = 3 on three particle streams per warp over 32*C slots (4C, C = 1..9: C forced) */
    int use_graph;    /* capture the step into a CUDA graph and replay it (default 1) */
    int slab_x0;      /* multi-GPU x-slab: this sim owns particles with slab_x0 <= base.x < slab_x1 */
    int slab_x1;      /* 0 (default) = not a slab: the sim owns the whole domain */
    int capacity;     /* slab mode: particle slots to allocate (>= n; room for migrants). 0 = n */
    int g2p_window;   /* 3D G2P node gather: 0 = auto, 1 = straight from global memory (L1/L2),
                         2 = node window of each 128-particle CTA staged in shared memory by the TMA
                         (cp.async.bulk.tensor + mbarrier; CTAs whose bounding box exceeds the window fall back to 1),
                         3 = 2 in persistent CTAs that request the next chunk's particle rows (cp.async) and node
                         window while the current chunk computes.  Env NMPM_G2P_WINDOW=0/1/2 (= modes 1/2/3) overrides.
                         Every mode reads the same nodes (same results).  Measured on cfg4: mode 1 is the fastest
                          and is what 0 selects */
    int fuse;         /* 3D, single GPU, with binning: G2P of step n also scatters step n+1 (P2G) into a second grid while
                         the particle is in registers (nmpm_fused.cuh) - 160 B instead of 260 B per particle and step.
                         0 = auto (on), 1 = off, 2 = on except for the first step after an upload (the uploaded state is
                         likely to be replaced again: teacher-forced loops), 3 = always.  Env NMPM_FUSE=0/1/2 (= 1/2/3)
                         overrides.  Results: same sums in a different floating-point order (like any sort cadence) */
    int tiles;        /* 3D, single GPU: grid_op and the grid clear visit only the 4^3-node tiles that stencils cover (one flag
                         bit per tile, raised together with the node box) instead of the particles' whole bounding box.
                         0 = auto: adaptive on grids of 8 Mi nodes and more (decided on the device per step: on while the
                         box holds more than ~1.5 nodes per particle, i.e. a dispersed scene), never on smaller grids,
                         1 = never, 2 = always, 3 = adaptive whatever the grid size.
                         Env NMPM_TILES=0/1/2/3 (= never / auto / always / adaptive) overrides.  Same results */
    int reserved[6];
} nmpm_options;

void nmpm_default_options(nmpm_options *opt);

/* MPMSimulation<dim>::MPMSimulation(particles, model, res, dt, E, nu, gravity) — src/nclr.h:74-78.
 * SoA host arrays in the interchange layout. */
int nmpm_create(int dim, int model, int res, float dt, float E, float nu, float gravity, size_t n,
                const float *x, const float *v, const float *F, const float *C, const float *Jp,
                const float *mass, const float *volume, const nmpm_options *opt, nmpm_handle *out);

/* Same constructor, taking the reference's AoS std::vector<Particle<dim>>::data() directly
 * (struct Particle — src/nclr.h:20-48: x, v, F, C, Jp, mass, volume, c; 64 B in 2D, 112 B in 3D).
 * `stride` = sizeof(Particle<dim>).  The int colour `c` is kept and returned by download_aos. */
int nmpm_create_aos(int dim, int model, int res, float dt, float E, float nu, float gravity, size_t n,
                    const void *particles_aos, size_t stride, const nmpm_options *opt, nmpm_handle *out);

/* A batch of independent 2D scenes behind ONE handle, advanced by the same launches (BASELINE.json config 5: 64 two-cube
 * scenes of 1 250 particles each; the reference runs one scene per process, src/solver.cpp:45-62,133-149 — a step of such a
 * scene is ~14 kernels of a few microseconds, so scene-by-scene the GPU is dispatch-bound).  The scenes share model, res, dt
 * and gravity and differ in particles and in E / nu.  Their grids are stacked along x (scene s owns node rows
 * [s*(res+1), (s+1)*(res+1)) of one tall grid), a particle's scene follows from its input index, every per-scene rule
 * (in-grid test, sticky walls, Lame parameters) is applied per scene: each scene evolves exactly as it would alone, up to
 * the order of floating-point sums.  Particle arrays = the scenes' arrays concatenated (counts[s] particles each);
 * downloads return the same layout; nmpm_download_grid* return nscenes*(res+1)^2 cells, scene by scene.
 * An out-of-grid particle in any scene fails the whole batch (NMPM_ERR_OUT_OF_GRID). */
int nmpm_create_batch(int model, int res, float dt, float gravity, int nscenes, const size_t *counts, const float *E,
                      const float *nu, const float *x, const float *v, const float *F, const float *C, const float *Jp,
                      const float *mass, const float *volume, const nmpm_options *opt, nmpm_handle *out);
int nmpm_create_batch_aos(int model, int res, float dt, float gravity, int nscenes, const size_t *counts, const float *E,
                          const float *nu, const void *particles_aos, size_t stride, const nmpm_options *opt,
                          nmpm_handle *out);
int nmpm_num_scenes(nmpm_handle h);                                   /* 1 for a plain sim */
int nmpm_batch_lame(nmpm_handle h, float *mu_0, float *lambda_0);     /* per scene (src/nclr.h:76-77), nscenes each */

void nmpm_destroy(nmpm_handle h);

/* MPMSimulation::advance() × nsteps — src/nclr.h:80-84.  Asynchronous on the sim's stream unless an
 * error must be reported; NMPM_ERR_OUT_OF_GRID is reported by the call that detects it or by the
 * next synchronising call (nmpm_synchronize / any download). */
int nmpm_advance(nmpm_handle h, int nsteps);

/* One phase of advance() (p2g / grid_op / g2p are private in the reference — src/nclr.h:104,167,263):
 * test and profiling hook; the post-P2G grid is needed for conservation checks (Q6). */
int nmpm_phase(nmpm_handle h, int phase);

int nmpm_synchronize(nmpm_handle h);

/* MPMSimulation::particles() — src/nclr.h:86.  Input order, forever.  Any output may be NULL. */
int nmpm_download_particles(nmpm_handle h, float *x, float *v, float *F, float *C, float *Jp);
int nmpm_download_particles_aos(nmpm_handle h, void *particles_aos, size_t stride);
/* positions only (src/example.cpp:77 consumes just x and c every 10th step) */
int nmpm_download_positions(nmpm_handle h, float *x);

/* MPMSimulation::grid() — src/nclr.h:87.  (res+1)^dim cells in reference index order.
 * Returns NMPM_OK and *cells_out = 0 before the first p2g (the reference's grid() is empty then,
 * src/solver.cpp:52-57).  gv: cells*dim (momentum after P2G, velocity after grid_op), gm: cells. */
int nmpm_download_grid(nmpm_handle h, float *gv, float *gm, size_t *cells_out);
/* struct Cell<dim> AoS — src/nclr.h:50-55 (12 B in 2D, 16 B in 3D) */
int nmpm_download_grid_aos(nmpm_handle h, void *cells_aos, size_t stride, size_t *cells_out);

/* Pipelined host I/O for per-step snapshot loops (the reference's solve_mpm pushes particles() every step,
 * src/solver.cpp:50-59): both calls only ENQUEUE — the upload's H2D copies run on a copy stream into double-buffered
 * device staging, the download's D2H copies on another, so the copy-in of step k+1 and the copy-out of step k-1
 * overlap step k and both PCIe directions stay busy.  Same semantics as the blocking forms below otherwise.  Host
 * arrays (pinned memory, or the copies serialise) must stay valid/untouched until nmpm_synchronize, or until two
 * further async calls of the same kind have been issued and completed.  An error is reported by nmpm_synchronize. */
int nmpm_upload_particles_async(nmpm_handle h, const float *x, const float *v, const float *F, const float *C,
                                const float *Jp);
int nmpm_download_particles_async(nmpm_handle h, float *x, float *v, float *F, float *C, float *Jp);

/* Replace the particle state (teacher-forced parity tests / resume from a snapshot: the reference
 * does this by constructing a new sim from a saved particle vector).  n must match.
 * A NULL v / F / C / Jp means "the constructor default" (v = 0, F = diag<dim>(1), C = 0, Jp = 1,
 * src/nclr.h:46-47), NOT "keep the device value"; mass and volume are always kept.  Allowed between steps
 * and in the middle of a step (after nmpm_phase): the aborted step's grid sums are discarded. */
int nmpm_upload_particles(nmpm_handle h, const float *x, const float *v, const float *F, const float *C,
                          const float *Jp);

size_t nmpm_num_particles(nmpm_handle h);
size_t nmpm_grid_cells(nmpm_handle h); /* (res+1)^dim */
/* public consts mu_0 / lambda_0 — src/nclr.h:71-72 (computed in fp32 exactly like the ctor) */
int nmpm_lame(nmpm_handle h, float *mu_0, float *lambda_0);

/* Binning debug hook (K0): recompute cell keys from the current positions and radix-sort them.
 * base: n*dim int32 in CURRENT device order; keys_sorted: n; perm: n (device slot of the i-th sorted
 * particle); ids: n original (input-order) index of each device slot.  Any output may be NULL. */
int nmpm_sort_debug(nmpm_handle h, int32_t *base, uint32_t *keys_unsorted, uint32_t *keys_sorted,
                    uint32_t *perm, uint32_t *ids);
/* key layout used by the solver: tile bits per axis (blocked key mode 1 of oracle/nclr_oracle.h) */
int nmpm_key_tile_bits(nmpm_handle h);

/* Device-side SVD / polar / stress unit hooks (nclr_svd, nclr_polar — src/nclr_math.h:50-98;
 * first_piola_kirchoff_stress — src/nclr.h:313-337).  count matrices, host pointers, column-major. */
int nmpm_svd_batch(int dim, size_t count, const float *A, float *U, float *sig, float *V, int device);
int nmpm_polar_batch(int dim, size_t count, const float *A, float *R, int device);
/* snow plasticity projection U clamp(sig, lo, hi) V^T of nclr_svd(A) — src/nclr.h:239-247 */
int nmpm_snow_project_batch(int dim, size_t count, const float *A, float lo, float hi, float *G, int device);
int nmpm_affine_debug(nmpm_handle h, float *A_out /* n*dim*dim, input order */);

/* Timing: accumulated CUDA-event milliseconds per phase since the last reset (only recorded when
 * enabled; recording breaks graph replay into per-kernel launches). */
enum nmpm_timer { NMPM_T_SORT = 0, NMPM_T_P2G = 1, NMPM_T_GRID = 2, NMPM_T_G2P = 3, NMPM_T_CLEAR = 4, NMPM_T_COUNT = 8 };
int nmpm_timing_enable(nmpm_handle h, int on);
int nmpm_timing_read(nmpm_handle h, float *ms /* NMPM_T_COUNT */, int *steps, int reset);
/* number of kernel launches issued by the library for this sim since creation */
long long nmpm_launch_count(nmpm_handle h);
/* 1 (or 2 = also right after an upload) when this sim runs the fused G2P+P2G kernel (nmpm_options.fuse), else 0: the
 * NMPM_T_G2P timer then covers G2P of step n AND the P2G scatter of step n+1, NMPM_T_P2G only the buffer swap + clear */
int nmpm_fused(nmpm_handle h);
/* 1 while the sim raises / uses active node tiles (nmpm_options.tiles; adaptive: follows the node box), else 0 */
int nmpm_tiles_active(nmpm_handle h);

/* Stream plumbing: run on an existing CUDA stream (e.g. torch.cuda.current_stream().cuda_stream). */
int nmpm_set_stream(nmpm_handle h, void *cuda_stream);
void *nmpm_get_stream(nmpm_handle h);

/* ---- Multi-GPU x-slabs -------------------------------------------------------
 * One process per GPU, one handle per slab (nmpm_options.slab_x0/x1/capacity).  Nothing in the
 * reference corresponds to this (it is a single address space, src/nclr.h:100-101); the slab step is
 * advance() (src/nclr.h:80-84) cut at the two points where neighbouring slabs must talk:
 *
 *   nmpm_slab_p2g        re-bin + sort the slab's particles, clear node planes [x0, x1+2), P2G
 *   -- exchange A: both neighbours swap their partial sums of the two shared node planes
 *      (nmpm_grid_plane_ptr to send, nmpm_grid_add_planes to add what was received) --
 *   nmpm_slab_grid_g2p   grid_op on planes [x0, x1+2) (the shared planes are updated redundantly on
 *                        both sides, bit-identically), G2P, and packing of the particles whose new
 *                        base.x left [x0, x1) into caller-provided device buffers
 *   -- exchange B: counts, then the migrant records, to the left / right neighbour --
 *   nmpm_slab_unpack     append the received records; they are binned by the next nmpm_slab_p2g
 *
 * The transport is the caller's (nuclearmpm_b200/slab.py: torch.distributed, NCCL send/recv on the
 * sim's stream).  All calls are asynchronous on the sim's stream.  A migrant record is the reference's
 * Particle<dim> AoS record (src/nclr.h:20-48: x, v, F, C, Jp, mass, volume, c) with the global particle
 * id in the colour word: 64 B in 2D, 112 B in 3D. */
size_t nmpm_migrate_record_bytes(nmpm_handle h);
/* device pointer to node plane `x_plane` of the grid (contiguous n1^(dim-1) float4 nodes) */
void *nmpm_grid_plane_ptr(nmpm_handle h, int x_plane);
size_t nmpm_grid_plane_bytes(nmpm_handle h);
/* add `planes` received node planes (device buffer) into the grid starting at x_plane */
int nmpm_grid_add_planes(nmpm_handle h, int x_plane, int planes, const void *device_src);
int nmpm_slab_p2g(nmpm_handle h);
/* send_left/send_right: device buffers of cap_records records each; d_counts: device int[4] =
 * {n_left, n_right, n_kept, overflow} written by the kernel (n_left/n_right may exceed cap_records only
 * together with overflow != 0, which is an error the caller must treat as fatal). */
int nmpm_slab_grid_g2p(nmpm_handle h, void *send_left, void *send_right, size_t cap_records, int *d_counts);
/* n_sent = records this slab packed in the last nmpm_slab_grid_g2p (n_left + n_right) */
int nmpm_slab_unpack(nmpm_handle h, const void *recv_left, size_t n_from_left, const void *recv_right,
                     size_t n_from_right, size_t n_sent);
/* move the slab's ownership range (re-balancing); particles now outside migrate at the next step */
int nmpm_slab_set_range(nmpm_handle h, int slab_x0, int slab_x1);
/* d_hist: device int[res+1], incremented by the number of live particles per base.x */
int nmpm_slab_histogram(nmpm_handle h, int *d_hist);
/* Native slab step: the protocol above driven inside the library with NCCL point-to-point calls
 * (ncclSend/ncclRecv in a group, one 12-int ncclAllGather for the migrant counts and node boxes) on the
 * sim's stream; ghost planes are exchanged as the in-plane rectangle the neighbouring slabs can touch.
 * NCCL is taken with dlopen from the libnccl.so.2 already loaded in the process (or `libnccl_path`).
 *   unique id: 128 bytes from nmpm_nccl_unique_id on rank 0, distributed by the caller;
 *   bounds: world+1 ownership boundaries (rank r owns bounds[r] <= base.x < bounds[r+1]). */
int nmpm_nccl_unique_id(void *out128, const char *libnccl_path);
int nmpm_slab_comm_init(nmpm_handle h, const void *unique_id128, int rank, int world, const int *bounds,
                        size_t cap_records, const char *libnccl_path);
/* particle count of the global simulation, the same on every rank (sizes the fixed-capacity migrant messages);
 * call right after nmpm_slab_comm_init */
int nmpm_slab_set_global_count(nmpm_handle h, size_t n_global);
int nmpm_slab_step(nmpm_handle h, int nsteps);
int nmpm_slab_set_bounds(nmpm_handle h, const int *bounds);
long long nmpm_slab_migrated(nmpm_handle h);
/* A sim driven by nmpm_slab_step keeps its particle counts on the device (the host only holds an upper bound of the
 * slots in use, so that no step waits for the host): true counts, synchronising the stream.  nmpm_num_particles
 * returns the same `particles`; nmpm_num_slots the host's bound (slots beyond the true count report id 0xFFFFFFFF
 * in nmpm_download_particles_slots, like migrated-away ones). */
int nmpm_slab_counts(nmpm_handle h, long long *particles, long long *slots_in_use);
/* global ids for the slab's particles (host array of n, creation order); particles() of the global
 * simulation is assembled from nmpm_download_particles_slots by scattering on ids */
int nmpm_set_ids(nmpm_handle h, const uint32_t *ids);
/* slot order; arrays sized for nmpm_num_slots(h) entries; slots whose particle has just migrated away
 * report id 0xFFFFFFFF (they are compacted away by the next nmpm_slab_p2g) */
int nmpm_download_particles_slots(nmpm_handle h, float *x, float *v, float *F, float *C, float *Jp, uint32_t *ids);
size_t nmpm_num_slots(nmpm_handle h);

const char *nmpm_last_error(nmpm_handle h); /* h may be NULL: last creation error */
const char *nmpm_build_info(void);          /* arch, compiler, date */

#ifdef __cplusplus
}
#endif
#endif /* NMPM_H */
