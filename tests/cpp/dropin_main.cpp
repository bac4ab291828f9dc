// Test driver for include/nclr.h (the drop-in C++ header): builds a cube scene exactly like the
// reference's callers do (src/example.cpp:29-46), advances, and dumps particles()/grid() as raw
// AoS records for the Python side to compare against the oracle.
//   dropin_main <dim> <model 0|1|2> <res> <cube_res> <lo> <hi> <steps> <out.bin>
#include <cstdio>
#include <cstdlib>
#include <fstream>

#include "nclr.h"

template<int dim>
int run(int model, int res, int cube_res, float lo, float hi, int steps, const char *out) {
    std::vector<nclr::Particle<dim>> particles;
    for (const auto &pos : nclr::cube<dim>(cube_res, lo, hi)) particles.emplace_back(nclr::Particle<dim>(pos, 0xED553B));
    nclr::MPMSimulation<dim> sim(particles, static_cast<nclr::MaterialModel>(model), res);
    if (!sim.grid().empty()) return 3;  // the reference's grid() is empty before the first advance()
    for (int s = 0; s < steps; ++s) sim.advance();
    const auto &ps = sim.particles();
    const auto &cells = sim.grid();
    std::ofstream f(out, std::ios::binary);
    const std::uint64_t hdr[4] = {ps.size(), sizeof(nclr::Particle<dim>), cells.size(), sizeof(nclr::Cell<dim>)};
    f.write(reinterpret_cast<const char *>(hdr), sizeof(hdr));
    f.write(reinterpret_cast<const char *>(ps.data()), std::streamsize(ps.size() * sizeof(nclr::Particle<dim>)));
    f.write(reinterpret_cast<const char *>(cells.data()), std::streamsize(cells.size() * sizeof(nclr::Cell<dim>)));
    std::printf("mu_0 %.9g lambda_0 %.9g colour %d\n", sim.mu_0, sim.lambda_0, ps.empty() ? 0 : ps[0].c);
    // positions(): the frame hand-off of src/example.cpp:56-82 — same values and order as particles()[i].x
    const auto &pos = sim.positions();
    if (pos.size() != ps.size()) return 5;
    for (std::size_t i = 0; i < ps.size(); ++i)
        for (int d = 0; d < dim; ++d)
            if (pos[i](d) != ps[i].x(d)) return 5;
    std::printf("positions ok\n");
    // out-of-grid particle => std::out_of_range, like vector::at in the reference (src/nclr.h:163)
    std::vector<nclr::Particle<dim>> bad = {nclr::Particle<dim>(nclr::constvec<dim>(0.999f), 0)};
    nclr::MPMSimulation<dim> sim2(bad, nclr::MaterialModel::kJelly, res);
    try {
        sim2.advance();
        sim2.particles();
        return 4;
    } catch (const std::out_of_range &e) { std::printf("out_of_range ok\n"); }
    return 0;
}

int main(int argc, char **argv) {
    if (argc != 9) return 2;
    const int dim = std::atoi(argv[1]), model = std::atoi(argv[2]), res = std::atoi(argv[3]), cres = std::atoi(argv[4]);
    const float lo = float(std::atof(argv[5])), hi = float(std::atof(argv[6]));
    const int steps = std::atoi(argv[7]);
    try {
        return dim == 2 ? run<2>(model, res, cres, lo, hi, steps, argv[8]) : run<3>(model, res, cres, lo, hi, steps, argv[8]);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
}
