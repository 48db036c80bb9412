// The C++ host wrapper on a QLTEN_Complex state: SetComplex + Evaluate(std::vector<std::complex<double>>).
// Reads: rows cols phys D walkers chi nsamples seed, the packed TPS as (re im) pairs and the initial configuration from stdin;
// prints Re E, Im E, energy_error, gradient_norm, accept rate.
#include <cstdio>
#include <iostream>
#include "../../include/peps_b200.hpp"

int main() {
  int rows, cols, phys, D, walkers, chi, nsamples;
  unsigned seed;
  std::cin >> rows >> cols >> phys >> D >> walkers >> chi >> nsamples >> seed;
  size_t n;
  std::cin >> n;
  std::vector<std::complex<double>> tps(n);
  for (auto &x : tps) { double a, b; std::cin >> a >> b; x = {a, b}; }
  peps_b200::MonteCarloParams mc;
  mc.num_samples = (size_t)nsamples; mc.sweeps_between_samples = 1;
  mc.initial_config.resize((size_t)rows * cols);
  for (auto &c : mc.initial_config) std::cin >> c;
  try {
    peps_b200::MCEnergyGradEvaluator ev(mc, peps_b200::BMPSTruncateParams::SVD(chi, chi, 0.0), rows, cols, phys, D, walkers,
                                        peps_b200::XXZModel{1.0, 1.0, 0.0}, seed);
    ev.batch().SetComplex();
    auto r = ev.Evaluate(tps);
    std::printf("%.17g %.17g %.17g %.17g %.17g\n", r.energy.real(), r.energy.imag(), r.energy_error, r.gradient_norm, r.accept_rates_avg[0]);
  } catch (const std::exception &e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
