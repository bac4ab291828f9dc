// Prints a few nclr::Vector / nclr::Matrix values with operator<< (Eigen default IOFormat rules).
#include <cstring>
#include <sstream>
#include <string>

#include "nclr.h"
int main() {
    nclr::Vector<nclr::real, 2> v(0.4f, 0.6f);
    nclr::Matrix<nclr::real, 2> m = nclr::diag<2>(1);
    nclr::Matrix<nclr::real, 2> w;
    w(0, 0) = 1.5f, w(0, 1) = -2.0f, w(1, 0) = 3.0f, w(1, 1) = 4.25f;
    nclr::Matrix<nclr::real, 3> d3 = nclr::diag<3>(1);
    nclr::Vector<nclr::real, 2> lame(3846.15384f, 5769.2307f);
    std::cout << v << "\n--\n" << m << "\n--\n" << w << "\n--\n" << d3 << "\n--\n" << lame.transpose() << "\n--\n"
              << nclr::constvec<3>(0.123456789f) << std::endl;
    // the snprintf fast path of operator<< against plain iostream formatting (what Eigen's IOFormat does), coefficient
    // by coefficient, on values of every magnitude and at two precisions
    std::uint32_t seed = 12345u;
    auto next = [&seed] {
        seed = seed * 1664525u + 1013904223u;
        return seed;
    };
    long bad = 0;
    for (int prec : {6, 9})
        for (int it = 0; it < 20000; ++it) {
            nclr::Matrix<nclr::real, 2> a;
            for (int k = 0; k < 4; ++k) {
                std::uint32_t bits = next();
                float f;
                std::memcpy(&f, &bits, 4);  // any bit pattern: denormals, inf, nan, huge, tiny
                if (it % 3 == 0) f = float(int(next() % 2001) - 1000) / 8.0f;
                if (it % 7 == 0 && k == 1) f = 0.0f;
                a.m[k] = f;
            }
            std::ostringstream fast, slow;
            fast.precision(prec), slow.precision(prec);
            fast << a;
            std::size_t width = 0;
            std::string cell[4];
            for (int k = 0; k < 4; ++k) {
                std::ostringstream ss;
                ss.precision(prec);
                ss << a.m[k];
                cell[k] = ss.str();
                width = std::max(width, cell[k].size());
            }
            for (int i = 0; i < 2; ++i) {
                if (i) slow << "\n";
                for (int j = 0; j < 2; ++j) {
                    if (j) slow << " ";
                    slow << std::string(width - cell[i + 2 * j].size(), ' ') << cell[i + 2 * j];
                }
            }
            if (fast.str() != slow.str()) ++bad;
        }
    std::cout << "fast-path mismatches: " << bad << std::endl;
    return 0;
}
