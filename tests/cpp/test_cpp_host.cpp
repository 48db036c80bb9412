// Compiles the C++ host wrapper (include/peps_b200.hpp) and runs Evaluate through the C ABI.
// Reads: rows cols phys D walkers chi nsamples seed, then the packed TPS and the initial configuration from stdin;
// prints energy, energy_error, gradient_norm, accept rate.
#include <cstdio>
#include <iostream>
#include "../../include/peps_b200.hpp"

int main() {
  int rows, cols, phys, D, walkers, chi, nsamples;
  unsigned seed;
  std::cin >> rows >> cols >> phys >> D >> walkers >> chi >> nsamples >> seed;
  size_t n;
  std::cin >> n;
  std::vector<double> tps(n);
  for (auto &x : tps) std::cin >> x;
  peps_b200::MonteCarloParams mc;
  mc.num_samples = (size_t)nsamples; mc.sweeps_between_samples = 1;
  mc.initial_config.resize((size_t)rows * cols);
  for (auto &c : mc.initial_config) std::cin >> c;
  try {
    peps_b200::MCEnergyGradEvaluator ev(mc, peps_b200::BMPSTruncateParams::SVD(chi, chi, 0.0), rows, cols, phys, D, walkers,
                                        peps_b200::XXZModel{1.0, 1.0, 0.0}, seed);
    if (ev.batch().tps_size() != n) throw std::runtime_error("tps size mismatch");
    auto r = ev.Evaluate(tps);
    std::printf("%.17g %.17g %.17g %.17g\n", r.energy, r.energy_error, r.gradient_norm, r.accept_rates_avg[0]);
    auto f = peps_b200::MakeEnergyEvaluator<std::vector<double>>(
        ev, [](const std::vector<double> &s) { return s; },
        [](const std::vector<double> &g, const std::vector<double> &) { return g; });
    auto t = f(tps);
    std::printf("%.17g\n", std::get<0>(t));
    // measurement through the wrapper: total energy equals the sum of the bond energies (no pinning field here)
    auto obs = ev.batch().Measure();
    double e0 = obs.energy[0], sum = 0.0;
    for (int i = 0; i < rows * (cols - 1); ++i) sum += obs.bond_energy_h[(size_t)i];
    for (int i = 0; i < (rows - 1) * cols; ++i) sum += obs.bond_energy_v[(size_t)i];
    std::printf("%.17g %.17g\n", e0, sum);
    // seam B2 as data: the XXZ bond term probed from a reference-style EvaluateBondEnergy functor reproduces the built-in
    // solver; rescue of an illegal walker is a no-op on legal configurations
    auto e_builtin = ev.batch().EnergyAndHoles(false);
    ev.batch().SetModelTerm(peps_b200::XXZBondTerm(0, 1.0, 1.0));
    auto e_table = ev.batch().EnergyAndHoles(false);
    ev.batch().ClearModelTerms();
    double dmax = 0.0;
    for (size_t w = 0; w < e_builtin.size(); ++w) dmax = std::max(dmax, std::fabs(e_builtin[w] - e_table[w]));
    std::printf("%.17g %d\n", dmax, (int)ev.batch().EnsureConfigurationValidity().size());
    // EvaluateEnergyOnly on a fresh evaluator: the same chains as Evaluate (same seeds), so the same energy
    peps_b200::MCEnergyGradEvaluator ev2(mc, peps_b200::BMPSTruncateParams::SVD(chi, chi, 0.0), rows, cols, phys, D, walkers,
                                         peps_b200::XXZModel{1.0, 1.0, 0.0}, seed);
    auto eo = ev2.EvaluateEnergyOnly(tps);
    std::printf("%.17g %.17g\n", std::get<0>(eo), std::get<2>(eo));
    // SR through the wrapper: Evaluate with collect_sr_buffers, then CalculateNaturalGradient on the device-resident samples
    peps_b200::MCEnergyGradEvaluator ev3(mc, peps_b200::BMPSTruncateParams::SVD(chi, chi, 0.0), rows, cols, phys, D, walkers,
                                         peps_b200::XXZModel{1.0, 1.0, 0.0}, seed);
    auto r3 = ev3.Evaluate(tps, true);
    peps_b200::ConjugateGradientParams cg;
    cg.max_iter = 200; cg.relative_tolerance = 1e-10;
    auto ng = ev3.CalculateNaturalGradient(r3, 1e-3, cg);
    double ng2 = 0.0;
    for (double x : std::get<0>(ng)) ng2 += x * x;
    std::printf("%.17g %d %zu\n", ng2, std::get<1>(ng), ev3.batch().SrCount());
    // MCPEPSMeasurer on a fresh batch of the same chains
    peps_b200::MCEnergyGradEvaluator ev4(mc, peps_b200::BMPSTruncateParams::SVD(chi, chi, 0.0), rows, cols, phys, D, walkers,
                                         peps_b200::XXZModel{1.0, 1.0, 0.0}, seed);
    ev4.batch().SetTPS(tps);
    ev4.batch().InitWalkers();
    peps_b200::MonteCarloParams mmc = mc;
    mmc.is_warmed_up = true;
    peps_b200::MCPEPSMeasurer meas(ev4.batch(), mmc, true);
    auto st = meas.Execute();
    std::printf("%.17g %.17g %.17g %zu\n", st["energy"].first[0], st["energy"].second[0], st["bond_energy_h"].first[0],
                st["SpSm_cross_raw"].first.size());
  } catch (const std::exception &e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
