// nclr.h — drop-in replacement for NuclearMPM's embeddable solver header, backed by libnmpm.so
// (hand-written sm_100a CUDA behind the C-ABI of nmpm.h).
//
// Same namespace, type names, member names, constructor and method signatures as the reference's
// src/nclr.h + src/nclr_math.h, so that src/example.cpp and src/solver.cpp of the reference compile
// against this header unchanged:
//
//   nclr::real, nclr::Vector<T,dim>, nclr::Matrix<T,dim>          src/nclr_math.h:4-11
//   nclr::diag / constmat / constvec / cube                       src/nclr_math.h:13-19,40-48,100-129
//   nclr::Particle<dim>, nclr::Cell<dim>, nclr::MaterialModel     src/nclr.h:20-61
//   nclr::MPMSimulation<dim>: ctor, advance(), particles(), grid(), mu_0, lambda_0, k* constants
//                                                                 src/nclr.h:63-87
//
// What is different from the reference:
//   * advance() enqueues the step on the GPU and returns; particles() / grid() synchronise, download
//     and return host mirrors in the reference's layout and order (input order, forever).
//   * A particle whose stencil leaves the grid makes the reference throw std::out_of_range out of
//     advance() (vector::at, src/nclr.h:163).  Here the same exception is thrown by the advance(),
//     particles() or grid() call that first observes the device flag (at the latest the next
//     synchronising call).  Define NCLR_SYNC_ERRORS to synchronise after every advance() and get the
//     reference's exact "throws from the faulting advance()" behaviour.
//   * Vector/Matrix are a small fixed-size column-major implementation (Eigen is not required).  With
//     -DNCLR_USE_EIGEN the aliases are Eigen's, exactly like the reference; sizeof(Particle<dim>) and
//     sizeof(Cell<dim>) are the same either way (64/112 B and 12/16 B).
//   * There is no CPU fallback: constructing a simulation without a usable B200 throws
//     std::runtime_error.
#pragma once

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdint>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#ifdef NCLR_USE_EIGEN
#include <Eigen/Dense>
#endif

#include "nmpm.h"

namespace nclr {
    using real = float;

#ifdef NCLR_USE_EIGEN
    template<typename T, int dim>
    using Vector = Eigen::Matrix<T, dim, 1>;
    template<typename T, int dim>
    using Matrix = Eigen::Matrix<T, dim, dim>;
#else
    // Fixed-size column-major dense matrix with the slice of Eigen's interface that callers of the
    // reference header use on solver types: (i), (i,j), x()/y()/z(), data(), rows()/cols()/size(),
    // Zero()/Constant(), transpose(), + - * with scalars and each other, operator<< (Eigen's default
    // IOFormat: stream precision, single space between columns, columns right-aligned to the widest
    // coefficient, one row per line).
    template<typename T, int R, int C>
    struct Mat {
        T m[R * C];

        Mat() = default;
        template<typename A, typename B, int RR = R, int CC = C,
                 typename = std::enable_if_t<RR * CC == 2 && std::is_arithmetic_v<A> && std::is_arithmetic_v<B>>>
        Mat(A a, B b) : m{T(a), T(b)} {}
        template<typename A, typename B, typename D, int RR = R, int CC = C,
                 typename = std::enable_if_t<RR * CC == 3 && std::is_arithmetic_v<A>>>
        Mat(A a, B b, D c) : m{T(a), T(b), T(c)} {}
        // a 2-vector built from a scalar fills both components (taichi::Vector2(0.04)-style call sites
        // never reach nclr types; kept explicit to avoid surprises)
        static Mat Zero() { return Constant(T(0)); }
        static Mat Ones() { return Constant(T(1)); }
        static Mat Constant(T v) {
            Mat r;
            for (int k = 0; k < R * C; ++k) r.m[k] = v;
            return r;
        }
        static Mat Identity() {
            Mat r = Zero();
            for (int k = 0; k < (R < C ? R : C); ++k) r(k, k) = T(1);
            return r;
        }

        static constexpr int rows() { return R; }
        static constexpr int cols() { return C; }
        static constexpr int size() { return R * C; }
        T *data() { return m; }
        const T *data() const { return m; }

        T &operator()(int i, int j) { return m[i + j * R]; }
        const T &operator()(int i, int j) const { return m[i + j * R]; }
        T &operator()(int i) { return m[i]; }
        const T &operator()(int i) const { return m[i]; }
        T &operator[](int i) { return m[i]; }
        const T &operator[](int i) const { return m[i]; }
        T &x() { return m[0]; }
        const T &x() const { return m[0]; }
        T &y() { return m[1]; }
        const T &y() const { return m[1]; }
        T &z() { return m[2]; }
        const T &z() const { return m[2]; }

        template<typename U>
        Mat<U, R, C> cast() const {
            Mat<U, R, C> r;
            for (int k = 0; k < R * C; ++k) r.m[k] = static_cast<U>(m[k]);
            return r;
        }
        Mat<T, C, R> transpose() const {
            Mat<T, C, R> r;
            for (int i = 0; i < R; ++i)
                for (int j = 0; j < C; ++j) r(j, i) = (*this)(i, j);
            return r;
        }
        T determinant() const {
            static_assert(R == C && (R == 2 || R == 3), "determinant: 2x2 or 3x3 only");
            const Mat &a = *this;
            if constexpr (R == 2) {
                return a(0, 0) * a(1, 1) - a(1, 0) * a(0, 1);
            } else {
                return a(0, 0) * (a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1)) -
                       a(0, 1) * (a(1, 0) * a(2, 2) - a(1, 2) * a(2, 0)) +
                       a(0, 2) * (a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0));
            }
        }
        T norm() const {
            T s = T(0);
            for (int k = 0; k < R * C; ++k) s += m[k] * m[k];
            return std::sqrt(s);
        }

        Mat &operator+=(const Mat &o) {
            for (int k = 0; k < R * C; ++k) m[k] += o.m[k];
            return *this;
        }
        Mat &operator-=(const Mat &o) {
            for (int k = 0; k < R * C; ++k) m[k] -= o.m[k];
            return *this;
        }
        Mat &operator*=(T s) {
            for (int k = 0; k < R * C; ++k) m[k] *= s;
            return *this;
        }
        Mat &operator/=(T s) {
            for (int k = 0; k < R * C; ++k) m[k] /= s;
            return *this;
        }
        friend Mat operator+(Mat a, const Mat &b) { return a += b; }
        friend Mat operator-(Mat a, const Mat &b) { return a -= b; }
        friend Mat operator-(Mat a) {
            for (int k = 0; k < R * C; ++k) a.m[k] = -a.m[k];
            return a;
        }
        friend Mat operator*(Mat a, T s) { return a *= s; }
        friend Mat operator*(T s, Mat a) { return a *= s; }
        friend Mat operator/(Mat a, T s) { return a /= s; }
        friend bool operator==(const Mat &a, const Mat &b) {
            for (int k = 0; k < R * C; ++k)
                if (!(a.m[k] == b.m[k])) return false;
            return true;
        }
        friend bool operator!=(const Mat &a, const Mat &b) { return !(a == b); }
    };

    template<typename T, int R, int K, int C>
    inline Mat<T, R, C> operator*(const Mat<T, R, K> &a, const Mat<T, K, C> &b) {
        Mat<T, R, C> r;
        for (int j = 0; j < C; ++j)
            for (int i = 0; i < R; ++i) {
                T acc = a(i, 0) * b(0, j);
                for (int k = 1; k < K; ++k) acc += a(i, k) * b(k, j);
                r(i, j) = acc;
            }
        return r;
    }

    // Eigen's default IOFormat (src/solver.cpp:71 streams x, v, F, C with it)
    template<typename T, int R, int C>
    inline std::ostream &operator<<(std::ostream &os, const Mat<T, R, C> &a) {
        // Fast path (the dump writers call this millions of times): a stream in the default state prints a float like
        // printf("%.*g", precision, double(x)); coefficients are formatted once into stack buffers and written
        // right-aligned to the widest one.  Any other stream state takes the generic two-pass path below.
        if constexpr (std::is_floating_point_v<T>) {
            const auto fl = os.flags();
            const bool plain = !(fl & std::ios_base::floatfield) && !(fl & (std::ios_base::showpos | std::ios_base::showpoint |
                                                                            std::ios_base::uppercase)) &&
                               (fl & std::ios_base::adjustfield) != std::ios_base::left &&
                               (fl & std::ios_base::adjustfield) != std::ios_base::internal && os.fill() == ' ' &&
                               os.width() == 0 && os.getloc() == std::locale::classic();
            if (plain) {
                char buf[R * C][40];
                int len[R * C];
                int width = 0;
                const int prec = os.precision() > 0 ? int(os.precision()) : (os.precision() == 0 ? 1 : 6);
                for (int k = 0; k < R * C; ++k) {
                    len[k] = std::snprintf(buf[k], sizeof buf[k], "%.*g", prec, double(a.m[k]));
                    width = std::max(width, len[k]);
                }
                static const char spaces[41] = "                                        ";
                for (int i = 0; i < R; ++i) {
                    if (i) os.put('\n');
                    for (int j = 0; j < C; ++j) {
                        if (j) os.put(' ');
                        const int k = i + j * R;
                        os.write(spaces, width - len[k]);
                        os.write(buf[k], len[k]);
                    }
                }
                return os;
            }
        }
        std::size_t width = 0;
        for (int k = 0; k < R * C; ++k) {
            std::ostringstream ss;
            ss.copyfmt(os);
            ss << a.m[k];
            width = std::max(width, ss.str().size());
        }
        for (int i = 0; i < R; ++i) {
            if (i) os << "\n";
            for (int j = 0; j < C; ++j) {
                if (j) os << " ";
                if (width) os.width(std::streamsize(width));
                os << a(i, j);
            }
        }
        return os;
    }

    template<typename T, int dim>
    using Vector = Mat<T, dim, 1>;
    template<typename T, int dim>
    using Matrix = Mat<T, dim, dim>;
#endif  // NCLR_USE_EIGEN

    // src/nclr_math.h:13-19 — sets (0,0) and (1,1) only; in 3D this is diag(v, v, 0)
    template<int dim>
    inline auto diag(const float value) -> Matrix<real, dim> {
        Matrix<real, dim> m = Matrix<real, dim>::Zero();
        m(0, 0) = value;
        m(1, 1) = value;
        return m;
    }
    // src/nclr_math.h:40-48
    template<int dim>
    inline auto constmat(const float value) -> Matrix<real, dim> {
        return Matrix<real, dim>::Constant(value);
    }
    template<int dim>
    inline auto constvec(const float value) -> Vector<real, dim> {
        return Vector<real, dim>::Constant(value);
    }

    // Eigen's Vector<float,-1>::LinSpaced(n, lo, hi) for floats: step = (hi-lo)/(n-1); values counted
    // from lo with the last pinned to hi, or — when |hi| < |lo| — counted back from hi with the first
    // pinned to lo.
    inline auto linspaced(int n, real lo, real hi) -> std::vector<real> {
        std::vector<real> a(size_t(n > 0 ? n : 0));
        if (n <= 0) return a;
        if (n == 1) {
            a[0] = lo;
            return a;
        }
        const real step = (hi - lo) / real(n - 1);
        if (std::fabs(hi) < std::fabs(lo)) {
            for (int i = 0; i < n; ++i) a[size_t(i)] = hi - real(n - 1 - i) * step;
            a[0] = lo;
        } else {
            for (int i = 0; i < n; ++i) a[size_t(i)] = lo + real(i) * step;
            a[size_t(n - 1)] = hi;
        }
        return a;
    }

    // src/nclr_math.h:100-129 — res^dim points, x slowest
    template<int dim>
    inline auto cube(int res, real min, real max) -> std::vector<Vector<real, dim>> {
        static_assert(dim == 2 || dim == 3, "cube<dim>: dim must be 2 or 3");
        const auto a = linspaced(res, min, max);
        std::vector<Vector<real, dim>> pts;
        pts.reserve(dim == 2 ? a.size() * a.size() : a.size() * a.size() * a.size());
        for (size_t i = 0; i < a.size(); ++i)
            for (size_t j = 0; j < a.size(); ++j) {
                if constexpr (dim == 2) {
                    pts.emplace_back(a[i], a[j]);
                } else {
                    for (size_t k = 0; k < a.size(); ++k) pts.emplace_back(a[i], a[j], a[k]);
                }
            }
        return pts;
    }

    // src/nclr.h:20-48 — member order is part of the contract: it is the AoS record libnmpm imports
    template<int dim>
    struct Particle {
        Vector<real, dim> x;   // position
        Vector<real, dim> v;   // velocity
        Matrix<real, dim> F;   // deformation gradient
        Matrix<real, dim> C;   // APIC affine momentum
        real Jp;               // plastic volume ratio
        real mass;
        real volume;
        int c;                 // colour

        Particle(Vector<real, dim> x, int c, Vector<real, dim> v = constvec<dim>(0), real mass = 1.0,
                 real volume = 1.0)
            : x(x), v(v), F(diag<dim>(1)), C(constmat<dim>(0)), Jp(1.0), mass(mass), volume(volume), c(c) {}
    };

    // src/nclr.h:50-55
    template<int dim>
    struct Cell {
        Vector<real, dim> velocity;
        real mass;
        Cell() : velocity(constvec<dim>(0)), mass(0.0) {}
    };

    static_assert(sizeof(Particle<2>) == 64 && sizeof(Particle<3>) == 112, "Particle<dim> AoS layout");
    static_assert(sizeof(Cell<2>) == 12 && sizeof(Cell<3>) == 16, "Cell<dim> AoS layout");

    // src/nclr.h:57-61
    enum class MaterialModel {
        kSnow = 0,
        kJelly,
        kLiquid,
    };

    // src/nclr.h:63-87
    template<int dim>
    class MPMSimulation {
        static_assert(dim == 2 || dim == 3, "MPMSimulation<dim>: dim must be 2 or 3");

    public:
        constexpr static int kBoundary = 3;
        constexpr static nclr::real kSnowHardening = 10.0;
        constexpr static nclr::real kJellyHardening = 0.3;
        constexpr static nclr::real kLiquidHardening = 1.0;

        const real mu_0;
        const real lambda_0;

        MPMSimulation(std::vector<Particle<dim>> particles, const MaterialModel model, int res = 64, real dt = 1e-4,
                      real E = 1e4, real nu = 0.2, real gravity = -100)
            : mu_0(E / (2 * (1 + nu))), lambda_0(E * nu / ((1 + nu) * (1 - 2 * nu))), particles_(std::move(particles)),
              res_(res) {
            const int rc = nmpm_create_aos(dim, static_cast<int>(model), res, dt, E, nu, gravity, particles_.size(),
                                           particles_.data(), sizeof(Particle<dim>), nullptr, &handle_);
            if (rc != NMPM_OK) {
                const std::string why = nmpm_last_error(nullptr);
                handle_ = nullptr;
                throw std::runtime_error("nclr::MPMSimulation: " + why);
            }
        }
        ~MPMSimulation() { nmpm_destroy(handle_); }
        MPMSimulation(const MPMSimulation &) = delete;
        MPMSimulation &operator=(const MPMSimulation &) = delete;

        auto advance() -> void { advance(1); }

        // extension: several steps per call (one CUDA-graph replay per step, no host sync in between)
        auto advance(int steps) -> void {
            particles_dirty_ = grid_dirty_ = true;
            check(nmpm_advance(handle_, steps));
#ifdef NCLR_SYNC_ERRORS
            check(nmpm_synchronize(handle_));
#endif
        }

        auto particles() const -> const std::vector<Particle<dim>> & {
            if (particles_dirty_) {
                check(nmpm_download_particles_aos(handle_, particles_.data(), sizeof(Particle<dim>)));
                particles_dirty_ = false;
            }
            return particles_;
        }

        // empty before the first advance() like the reference (src/solver.cpp:52-57 works around it)
        auto grid() const -> const std::vector<Cell<dim>> & {
            if (grid_dirty_) {
                cells_.assign(nmpm_grid_cells(handle_), Cell<dim>());
                size_t got = 0;
                check(nmpm_download_grid_aos(handle_, cells_.data(), sizeof(Cell<dim>), &got));
                cells_.resize(got);
                grid_dirty_ = false;
            }
            return cells_;
        }

        // extension: positions only, input order (what src/example.cpp:77 draws every 10th step)
        auto positions() const -> const std::vector<Vector<real, dim>> & {
            positions_.resize(particles_.size());
            check(nmpm_download_positions(handle_, reinterpret_cast<float *>(positions_.data())));
            return positions_;
        }

        // extension: block until the GPU has finished every enqueued step (throws like advance())
        auto synchronize() const -> void { check(nmpm_synchronize(handle_)); }
        auto handle() const -> nmpm_handle { return handle_; }

    private:
        auto check(int rc) const -> void {
            if (rc == NMPM_OK) return;
            const std::string why = nmpm_last_error(handle_);
            if (rc == NMPM_ERR_OUT_OF_GRID) throw std::out_of_range("nclr::MPMSimulation: " + why);
            throw std::runtime_error("nclr::MPMSimulation: " + why);
        }

        mutable std::vector<Cell<dim>> cells_;
        mutable std::vector<Particle<dim>> particles_;
        mutable std::vector<Vector<real, dim>> positions_;
        mutable bool particles_dirty_ = false;
        mutable bool grid_dirty_ = false;
        const int res_;
        nmpm_handle handle_ = nullptr;
    };

    // Extension (no counterpart in the reference, which runs one scene per process — src/solver.cpp:45-62): a batch of
    // independent 2D scenes with the same model / res / dt / gravity behind ONE simulation (nmpm_create_batch_aos), advanced
    // by the same kernel launches.  Each scene evolves as it would alone (up to the order of floating-point sums).
    class MPMBatch2D {
    public:
        MPMBatch2D(const std::vector<std::vector<Particle<2>>> &scenes, const MaterialModel model, int res, real dt,
                   const std::vector<real> &E, const std::vector<real> &nu, real gravity)
            : res_(res) {
            if (scenes.empty() || E.size() != scenes.size() || nu.size() != scenes.size())
                throw std::invalid_argument("nclr::MPMBatch2D: one E and one nu per scene");
            std::vector<size_t> counts;
            offsets_.push_back(0);
            for (const auto &sc : scenes) {
                counts.push_back(sc.size());
                particles_.insert(particles_.end(), sc.begin(), sc.end());
                offsets_.push_back(particles_.size());
            }
            const int rc = nmpm_create_batch_aos(static_cast<int>(model), res, dt, gravity, static_cast<int>(scenes.size()),
                                                 counts.data(), E.data(), nu.data(), particles_.data(), sizeof(Particle<2>),
                                                 nullptr, &handle_);
            if (rc != NMPM_OK) {
                const std::string why = nmpm_last_error(nullptr);
                handle_ = nullptr;
                throw std::runtime_error("nclr::MPMBatch2D: " + why);
            }
            mu_0_.resize(scenes.size()), lambda_0_.resize(scenes.size());
            nmpm_batch_lame(handle_, mu_0_.data(), lambda_0_.data());
        }
        ~MPMBatch2D() { nmpm_destroy(handle_); }
        MPMBatch2D(const MPMBatch2D &) = delete;
        MPMBatch2D &operator=(const MPMBatch2D &) = delete;

        auto scenes() const -> size_t { return offsets_.size() - 1; }
        auto mu_0(size_t s) const -> real { return mu_0_.at(s); }
        auto lambda_0(size_t s) const -> real { return lambda_0_.at(s); }

        auto advance(int steps = 1) -> void {
            particles_dirty_ = grid_dirty_ = true;
            check(nmpm_advance(handle_, steps));
        }
        auto synchronize() const -> void { check(nmpm_synchronize(handle_)); }

        // particles of scene s, input order (one download of the whole batch per step, shared by all scenes)
        auto particles(size_t s) const -> std::vector<Particle<2>> {
            if (particles_dirty_) {
                check(nmpm_download_particles_aos(handle_, particles_.data(), sizeof(Particle<2>)));
                particles_dirty_ = false;
            }
            return std::vector<Particle<2>>(particles_.begin() + offsets_.at(s), particles_.begin() + offsets_.at(s + 1));
        }
        // grid of scene s: (res+1)^2 cells in the reference's order; empty before the first advance()
        auto grid(size_t s) const -> std::vector<Cell<2>> {
            if (grid_dirty_) {
                cells_.assign(nmpm_grid_cells(handle_), Cell<2>());
                size_t got = 0;
                check(nmpm_download_grid_aos(handle_, cells_.data(), sizeof(Cell<2>), &got));
                cells_.resize(got);
                grid_dirty_ = false;
            }
            if (cells_.empty()) return {};
            const size_t per = size_t(res_ + 1) * size_t(res_ + 1);
            return std::vector<Cell<2>>(cells_.begin() + s * per, cells_.begin() + (s + 1) * per);
        }

    private:
        auto check(int rc) const -> void {
            if (rc == NMPM_OK) return;
            const std::string why = nmpm_last_error(handle_);
            if (rc == NMPM_ERR_OUT_OF_GRID) throw std::out_of_range("nclr::MPMBatch2D: " + why);
            throw std::runtime_error("nclr::MPMBatch2D: " + why);
        }
        mutable std::vector<Particle<2>> particles_;
        mutable std::vector<Cell<2>> cells_;
        std::vector<size_t> offsets_;
        std::vector<real> mu_0_, lambda_0_;
        mutable bool particles_dirty_ = false;
        mutable bool grid_dirty_ = false;
        const int res_;
        nmpm_handle handle_ = nullptr;
    };
}// namespace nclr
