// Host experiment: sweeps / rotations of the one-sided Jacobi (csrc/nmpm_math.cuh svd3_recompose) on snow F' = (diag(1,1,0)+dt C) F,
// column-wise vs row-wise (transposed input).  g++ -std=c++17 -O2 -mfma -ffp-contract=off -I/usr/local/cuda/include tools/hestenes_sweeps.cpp
#define NMPM_HOST_STATS
#include "../nuclearmpm_b200/csrc/nmpm_math.cuh"
#include <cstdio>
#include <random>
#include <cmath>
namespace nmpm { long nmpm_stat_rot, nmpm_stat_sweep, nmpm_stat_calls; }
using namespace nmpm;
static void qr_rot(std::mt19937& g, double q[3][3]) {
    std::normal_distribution<double> N;
    double a[3][3];
    for (auto& r : a) for (auto& x : r) x = N(g);
    // gram-schmidt columns
    for (int j = 0; j < 3; ++j) {
        for (int k = 0; k < j; ++k) { double d = 0; for (int i = 0; i < 3; ++i) d += a[i][j]*a[i][k]; for (int i = 0; i < 3; ++i) a[i][j] -= d*a[i][k]; }
        double n = 0; for (int i = 0; i < 3; ++i) n += a[i][j]*a[i][j]; n = std::sqrt(n); for (int i = 0; i < 3; ++i) a[i][j] /= n;
    }
    double det = a[0][0]*(a[1][1]*a[2][2]-a[1][2]*a[2][1]) - a[0][1]*(a[1][0]*a[2][2]-a[1][2]*a[2][0]) + a[0][2]*(a[1][0]*a[2][1]-a[1][1]*a[2][0]);
    if (det < 0) for (int i = 0; i < 3; ++i) a[i][0] = -a[i][0];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) q[i][j] = a[i][j];
}
int main() {
    std::mt19937 g(1);
    std::uniform_real_distribution<double> S(0.975, 1.0045);
    std::normal_distribution<double> N;
    for (double cs : {1.0, 100.0, 3000.0, 15000.0}) {
        long rot[2] = {0, 0}, sw[2] = {0, 0}; double maxdiff = 0; int n = 20000; long fails[2]={0,0};
        for (int it = 0; it < n; ++it) {
            double U[3][3], V[3][3]; qr_rot(g, U); qr_rot(g, V);
            double s[3] = {S(g), S(g), S(g)};
            double F[3][3] = {}, M[3][3], Fp[3][3] = {};
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) for (int k = 0; k < 3; ++k) F[i][j] += U[i][k]*s[k]*V[j][k];
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) M[i][j] = 1e-4 * cs * N(g) + ((i == j && i < 2) ? 1.0 : 0.0);
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) for (int k = 0; k < 3; ++k) Fp[i][j] += M[i][k]*F[k][j];
            Mat<3> A, At, G, Gt;
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { A(i, j) = (float) Fp[i][j]; At(j, i) = (float) Fp[i][j]; }
            nmpm_stat_rot = nmpm_stat_sweep = 0;
            if (!svd3_recompose<1>(A, 0.975f, 1.0045f, G)) fails[0]++;
            rot[0] += nmpm_stat_rot; sw[0] += nmpm_stat_sweep;
            nmpm_stat_rot = nmpm_stat_sweep = 0;
            if (!svd3_recompose<1>(At, 0.975f, 1.0045f, Gt)) fails[1]++;
            rot[1] += nmpm_stat_rot; sw[1] += nmpm_stat_sweep;
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) maxdiff = std::fmax(maxdiff, std::fabs(G(i, j) - Gt(j, i)));
        }
        printf("Cscale %7.0f: col-wise sweeps %.2f rot %.2f fail %ld | row-wise sweeps %.2f rot %.2f fail %ld | max|G-Gt^T| %.2e\n", cs,
               (double) sw[0]/n, (double) rot[0]/n, fails[0], (double) sw[1]/n, (double) rot[1]/n, fails[1], maxdiff);
    }
}
