// Prints a few nclr::Vector / nclr::Matrix values with operator<< (Eigen default IOFormat rules).
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
    return 0;
}
