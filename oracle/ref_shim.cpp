// TEST INFRASTRUCTURE ONLY.  C-ABI shim around the UNMODIFIED reference header
// (/root/reference/src/nclr.h, included from where it lies; never copied into this repo),
// compiled against oracle/eigen_standin.  Built by oracle/Makefile into oracle/_ref/ as
//   libnclr_ref_strict.so  (-O2 -ffp-contract=off : deterministic parity oracle)
//   libnclr_ref_fast.so    (-Ofast -DNDEBUG       : CMakeLists.txt:8-10 flags, timed CPU baseline)
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load it.
//
// Interchange layout (shared with oracle/nclr_oracle.c and the product C-ABI, include/nmpm.h):
//   x,v : n*dim floats, particle-major;  F,C : n*dim*dim floats, per particle column-major
//   (Eigen storage: M(i,j) at [i + j*dim]);  Jp,mass,volume : n floats.
//   grid : node-major, index = x*n1+y (2D) / (x*n1+y)*n1+z (3D), n1=res+1 (src/nclr.h:141-142,152);
//   gv : nodes*dim floats, gm : nodes floats.
#include <algorithm>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <memory>
#include <sstream>
#include <utility>
#include <vector>

#include <Eigen/Dense>

// The phases (p2g / grid_op / g2p) and the stress helper are private in the reference class;
// the oracle needs them one at a time (post-P2G grid for conservation checks).
// All standard headers the reference pulls in are already included above, so this only
// affects the reference's own two headers.
#define private public
#include "nclr.h"
#undef private

namespace {
    struct RefSim {
        int dim;
        std::unique_ptr<nclr::MPMSimulation<2>> s2;
        std::unique_ptr<nclr::MPMSimulation<3>> s3;
    };

    template<int dim>
    std::vector<nclr::Particle<dim>> unpack(long n, const float *x, const float *v, const float *F, const float *C,
                                            const float *Jp, const float *mass, const float *volume) {
        std::vector<nclr::Particle<dim>> ps;
        ps.reserve(size_t(n));
        for (long p = 0; p < n; ++p) {
            nclr::Vector<nclr::real, dim> px;
            for (int d = 0; d < dim; ++d) px(d) = x[p * dim + d];
            nclr::Particle<dim> q(px, 0);
            if (v)
                for (int d = 0; d < dim; ++d) q.v(d) = v[p * dim + d];
            if (F)
                for (int k = 0; k < dim * dim; ++k) q.F.data()[k] = F[p * dim * dim + k];
            if (C)
                for (int k = 0; k < dim * dim; ++k) q.C.data()[k] = C[p * dim * dim + k];
            if (Jp) q.Jp = Jp[p];
            if (mass) q.mass = mass[p];
            if (volume) q.volume = volume[p];
            ps.push_back(q);
        }
        return ps;
    }

    template<int dim>
    void pack(const std::vector<nclr::Particle<dim>> &ps, float *x, float *v, float *F, float *C, float *Jp) {
        for (size_t p = 0; p < ps.size(); ++p) {
            const auto &q = ps[p];
            for (int d = 0; d < dim; ++d) {
                if (x) x[p * dim + d] = q.x(d);
                if (v) v[p * dim + d] = q.v(d);
            }
            for (int k = 0; k < dim * dim; ++k) {
                if (F) F[p * dim * dim + k] = q.F.data()[k];
                if (C) C[p * dim * dim + k] = q.C.data()[k];
            }
            if (Jp) Jp[p] = q.Jp;
        }
    }

    template<int dim>
    long grid_out(const std::vector<nclr::Cell<dim>> &cells, float *gv, float *gm) {
        for (size_t i = 0; i < cells.size(); ++i) {
            if (gv)
                for (int d = 0; d < dim; ++d) gv[i * dim + d] = cells[i].velocity(d);
            if (gm) gm[i] = cells[i].mass;
        }
        return long(cells.size());
    }
}// namespace

extern "C" {

void *nclr_ref_create(int dim, int model, int res, float dt, float E, float nu, float gravity, long n, const float *x,
                      const float *v, const float *F, const float *C, const float *Jp, const float *mass,
                      const float *volume) {
    auto *h = new RefSim();
    h->dim = dim;
    const auto mm = static_cast<nclr::MaterialModel>(model);
    if (dim == 2)
        h->s2 = std::make_unique<nclr::MPMSimulation<2>>(unpack<2>(n, x, v, F, C, Jp, mass, volume), mm, res, dt, E, nu,
                                                         gravity);
    else
        h->s3 = std::make_unique<nclr::MPMSimulation<3>>(unpack<3>(n, x, v, F, C, Jp, mass, volume), mm, res, dt, E, nu,
                                                         gravity);
    return h;
}

void nclr_ref_destroy(void *hv) { delete static_cast<RefSim *>(hv); }

// returns 0, or 1 if the reference threw std::out_of_range (Q5: particle left the grid)
int nclr_ref_advance(void *hv, int nsteps) {
    auto *h = static_cast<RefSim *>(hv);
    try {
        for (int s = 0; s < nsteps; ++s) {
            if (h->dim == 2) h->s2->advance();
            else
                h->s3->advance();
        }
    } catch (const std::out_of_range &) { return 1; }
    return 0;
}

// phase: 0 = p2g, 1 = grid_op, 2 = g2p  (src/nclr.h:80-84 runs them in this order)
int nclr_ref_phase(void *hv, int phase) {
    auto *h = static_cast<RefSim *>(hv);
    try {
        if (h->dim == 2) {
            if (phase == 0) h->s2->p2g();
            else if (phase == 1)
                h->s2->grid_op();
            else
                h->s2->g2p();
        } else {
            if (phase == 0) h->s3->p2g();
            else if (phase == 1)
                h->s3->grid_op();
            else
                h->s3->g2p();
        }
    } catch (const std::out_of_range &) { return 1; }
    return 0;
}

// seconds of steady_clock around the advance() loop only
double nclr_ref_time_advance(void *hv, int nsteps) {
    auto *h = static_cast<RefSim *>(hv);
    const auto t0 = std::chrono::steady_clock::now();
    for (int s = 0; s < nsteps; ++s) {
        if (h->dim == 2) h->s2->advance();
        else
            h->s3->advance();
    }
    const auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

long nclr_ref_num_particles(void *hv) {
    auto *h = static_cast<RefSim *>(hv);
    return h->dim == 2 ? long(h->s2->particles().size()) : long(h->s3->particles().size());
}

void nclr_ref_get_particles(void *hv, float *x, float *v, float *F, float *C, float *Jp) {
    auto *h = static_cast<RefSim *>(hv);
    if (h->dim == 2) pack<2>(h->s2->particles(), x, v, F, C, Jp);
    else
        pack<3>(h->s3->particles(), x, v, F, C, Jp);
}

// number of cells currently in grid() (0 before the first p2g — src/solver.cpp:52-57 relies on it)
long nclr_ref_get_grid(void *hv, float *gv, float *gm) {
    auto *h = static_cast<RefSim *>(hv);
    return h->dim == 2 ? grid_out<2>(h->s2->grid(), gv, gm) : grid_out<3>(h->s3->grid(), gv, gm);
}

void nclr_ref_lame(void *hv, float *mu0, float *lambda0) {
    auto *h = static_cast<RefSim *>(hv);
    *mu0 = h->dim == 2 ? h->s2->mu_0 : h->s3->mu_0;
    *lambda0 = h->dim == 2 ? h->s2->lambda_0 : h->s3->lambda_0;
}

// nclr_svd (src/nclr_math.h:50-74); all matrices column-major dim*dim
void nclr_ref_svd(int dim, const float *a, float *U, float *sig, float *V) {
    if (dim == 2) {
        nclr::Matrix<nclr::real, 2> A, u, s, v;
        std::memcpy(A.data(), a, sizeof(float) * 4);
        nclr::nclr_svd<2>(A, u, s, v);
        std::memcpy(U, u.data(), sizeof(float) * 4);
        std::memcpy(sig, s.data(), sizeof(float) * 4);
        std::memcpy(V, v.data(), sizeof(float) * 4);
    } else {
        nclr::Matrix<nclr::real, 3> A, u, s, v;
        std::memcpy(A.data(), a, sizeof(float) * 9);
        nclr::nclr_svd<3>(A, u, s, v);
        std::memcpy(U, u.data(), sizeof(float) * 9);
        std::memcpy(sig, s.data(), sizeof(float) * 9);
        std::memcpy(V, v.data(), sizeof(float) * 9);
    }
}

// nclr_polar (src/nclr_math.h:76-98)
void nclr_ref_polar(int dim, const float *m, float *R, float *S) {
    if (dim == 2) {
        nclr::Matrix<nclr::real, 2> M, r, s;
        std::memcpy(M.data(), m, sizeof(float) * 4);
        nclr::nclr_polar<2>(M, r, s);
        std::memcpy(R, r.data(), sizeof(float) * 4);
        std::memcpy(S, s.data(), sizeof(float) * 4);
    } else {
        nclr::Matrix<nclr::real, 3> M, r, s;
        std::memcpy(M.data(), m, sizeof(float) * 9);
        nclr::nclr_polar<3>(M, r, s);
        std::memcpy(R, r.data(), sizeof(float) * 9);
        std::memcpy(S, s.data(), sizeof(float) * 9);
    }
}

// first_piola_kirchoff_stress (src/nclr.h:313-337) for particle `p` of the sim: the fused affine matrix
void nclr_ref_affine(void *hv, long p, float *A) {
    auto *h = static_cast<RefSim *>(hv);
    if (h->dim == 2) {
        const auto a = h->s2->first_piola_kirchoff_stress(h->s2->particles_.at(size_t(p)));
        std::memcpy(A, a.data(), sizeof(float) * 4);
    } else {
        const auto a = h->s3->first_piola_kirchoff_stress(h->s3->particles_.at(size_t(p)));
        std::memcpy(A, a.data(), sizeof(float) * 9);
    }
}

// cube<dim>(res,min,max) (src/nclr_math.h:100-129): writes res^dim points, returns the count
long nclr_ref_cube(int dim, int res, float lo, float hi, float *out) {
    if (dim == 2) {
        const auto pts = nclr::cube<2>(res, lo, hi);
        if (out)
            for (size_t i = 0; i < pts.size(); ++i) {
                out[2 * i] = pts[i](0);
                out[2 * i + 1] = pts[i](1);
            }
        return long(pts.size());
    }
    const auto pts = nclr::cube<3>(res, lo, hi);
    if (out)
        for (size_t i = 0; i < pts.size(); ++i) {
            out[3 * i] = pts[i](0);
            out[3 * i + 1] = pts[i](1);
            out[3 * i + 2] = pts[i](2);
        }
    return long(pts.size());
}

long nclr_ref_oob_events(void) { return Eigen::standin_oob_events(); }
void nclr_ref_oob_reset(void) { Eigen::standin_oob_events() = 0; }

int nclr_ref_sizeof_particle(int dim) {
    return dim == 2 ? int(sizeof(nclr::Particle<2>)) : int(sizeof(nclr::Particle<3>));
}
int nclr_ref_sizeof_cell(int dim) { return dim == 2 ? int(sizeof(nclr::Cell<2>)) : int(sizeof(nclr::Cell<3>)); }

}// extern "C"
