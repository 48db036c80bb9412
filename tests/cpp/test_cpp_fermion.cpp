// C++ driver of the fermion mode through include/peps_b200.hpp: exact summation of the spinless-fermion energy over the
// configurations read from stdin (the reference's Z2SpinlessFreeFermionTest, tests/test_algorithm/
// test_exact_summation_evaluator.cpp:428-458, with the model written against the EvaluateBondEnergy probe).
// stdin: rows cols phys D W chi t t2 V | phys_par[phys] | n_leg leg_par[n_leg] | n_tps tps[n_tps] | configs[W*rows*cols]
#include <cstdio>
#include <iostream>
#include "../../include/peps_b200.hpp"

int main() {
  int rows, cols, phys, D, W, chi;
  double t, t2, V;
  std::cin >> rows >> cols >> phys >> D >> W >> chi >> t >> t2 >> V;
  peps_b200::FermionParities par;
  par.phys_par.resize((size_t)phys);
  for (auto &x : par.phys_par) std::cin >> x;
  size_t n;
  std::cin >> n;
  par.leg_par.resize(n);
  for (auto &x : par.leg_par) std::cin >> x;
  std::cin >> n;
  std::vector<double> tps(n);
  for (auto &x : tps) std::cin >> x;
  std::vector<int32_t> cfg((size_t)W * rows * cols);
  for (auto &x : cfg) std::cin >> x;
  try {
    peps_b200::WalkerBatch b(rows, cols, phys, D, W, peps_b200::BMPSTruncateParams::SVD((size_t)chi, (size_t)chi, 1e-16));
    b.SetFermion(par);
    b.SetTPS(tps);
    b.SetModelTerm(peps_b200::SpinlessFermionBondTerm(t, V));
    if (t2 != 0.0) b.SetModelTerm(peps_b200::SpinlessFermionNNNTerm(t2));
    b.SetConfigs(cfg);
    b.InitWalkers();
    std::vector<double> e = b.EnergyAndHoles(false), a = b.Amplitudes();
    double num = 0.0, den = 0.0;
    for (int w = 0; w < W; ++w) { num += a[(size_t)w] * a[(size_t)w] * e[(size_t)w]; den += a[(size_t)w] * a[(size_t)w]; }
    std::printf("%.15e\n", num / den);
    // the evaluator built from model terms + parities: 2 samples per walker from the first configuration, seed 77
    peps_b200::MonteCarloParams mc;
    mc.num_samples = (size_t)(2 * W); mc.num_warmup_sweeps = 0; mc.sweeps_between_samples = 1;
    mc.initial_config.assign(cfg.begin(), cfg.begin() + rows * cols);
    std::vector<peps_b200::ModelTerm> terms = {peps_b200::SpinlessFermionBondTerm(t, V)};
    if (t2 != 0.0) terms.push_back(peps_b200::SpinlessFermionNNNTerm(t2));
    peps_b200::MCEnergyGradEvaluator ev(mc, peps_b200::BMPSTruncateParams::SVD((size_t)chi, (size_t)chi, 1e-16), rows, cols, phys, D, W,
                                        terms, 77u, &par);
    peps_b200::EvaluateResult r = ev.Evaluate(tps);
    std::printf("%.15e %.15e\n", r.energy, r.gradient_norm);
  } catch (const std::exception &ex) {
    std::fprintf(stderr, "error: %s\n", ex.what());
    return 1;
  }
  return 0;
}
