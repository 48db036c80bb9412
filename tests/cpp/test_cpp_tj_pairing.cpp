// C++ driver of the t-J pairing pieces through include/peps_b200.hpp: SetBondPin(tJSingletPairTerm(2, delta)) in E_loc and
// MeasureBondTerm(tJSingletPairTerm(0 / 1)) = (delta_dag, delta) on every NN bond (square_tJ_model.h:86-137, 546-602).
// stdin: rows cols D W chi t J mu delta pin_s1 pin_s2 | n_leg leg_par[n_leg] | n_tps tps[n_tps] | configs[W*rows*cols]
// stdout: W unpinned energies, W pinned energies, then delta_dag horizontal / vertical and delta horizontal / vertical arrays
#include <cstdio>
#include <iostream>
#include "../../include/peps_b200.hpp"

int main() {
  int rows, cols, D, W, chi, s1, s2;
  double t, J, mu, delta;
  std::cin >> rows >> cols >> D >> W >> chi >> t >> J >> mu >> delta >> s1 >> s2;
  peps_b200::FermionParities par;
  par.phys_par = {1, 1, 0};
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
    peps_b200::WalkerBatch b(rows, cols, 3, D, W, peps_b200::BMPSTruncateParams::SVD((size_t)chi, (size_t)chi, 0.0));
    b.SetFermion(par);
    b.SetTPS(tps);
    b.SetModelTerm(peps_b200::tJBondTerm(t, J, 0.0));
    b.SetModelTerm(peps_b200::tJOnsiteTerm(mu));
    b.SetConfigs(cfg);
    b.InitWalkers();
    for (double e : b.EnergyAndHoles(false)) std::printf("%.17g\n", e);
    b.SetBondPin(s1, s2, peps_b200::tJSingletPairTerm(2, delta));
    for (double e : b.EnergyAndHoles(false)) std::printf("%.17g\n", e);
    b.ClearBondPin();
    for (int which = 0; which < 2; ++which) {
      auto hv = b.MeasureBondTerm(peps_b200::tJSingletPairTerm(which));
      for (double x : hv.first) std::printf("%.17g\n", x);
      for (double x : hv.second) std::printf("%.17g\n", x);
    }
  } catch (const std::exception &ex) {
    std::fprintf(stderr, "error: %s\n", ex.what());
    return 1;
  }
  return 0;
}
