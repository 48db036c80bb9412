"""The example drivers run end to end on the host simulation (host logic only; the GPU twin runs the CUDA library)."""
import os
import sys

import numpy as np

import hostsim_lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "examples"))


def run_tfim_example(lib):
    """examples/tfim_vmc_optimize.py (the reference's transverse_field_ising_vmc_optimize.cpp flow: Evaluate -> natural gradient
    -> update) on a 3x3 lattice: the SR iterations lower the energy towards the exact ground state."""
    import tfim_vmc_optimize as ex
    energies, state = ex.optimize(rows=3, cols=3, D=2, chi=4, h=0.5, walkers=8, samples=160, iters=8, step=0.15, lib=lib,
                                  log=lambda *_: None)
    # exact ground-state energy of H = -sum zz - h sum x on the 3x3 open lattice by dense diagonalisation
    n = 9
    H = np.zeros((2 ** n, 2 ** n))
    idx = np.arange(2 ** n)
    bits = (idx[:, None] >> np.arange(n)[None, :]) & 1
    sz = 1.0 - 2.0 * bits
    for r in range(3):
        for c in range(3):
            s = r * 3 + c
            if c + 1 < 3:
                H[idx, idx] += -sz[:, s] * sz[:, s + 1]
            if r + 1 < 3:
                H[idx, idx] += -sz[:, s] * sz[:, s + 3]
            H[idx, idx ^ (1 << s)] += -0.5
    e0 = np.linalg.eigvalsh(H)[0]
    assert energies[-1] < energies[0] - 0.5, energies           # the optimisation works ...
    assert energies[-1] > e0 - 0.3, (energies[-1], e0)          # ... and stays variational within the statistical error
    return energies, e0


def test_tfim_vmc_optimize_example_hostsim():
    run_tfim_example(hostsim_lib.load())


import pytest


@pytest.mark.gpu
def test_tfim_vmc_optimize_example_gpu():
    from peps_b200 import _lib
    lib = _lib.load()
    assert lib.peps_backend_name() == b"cuda-sm_100a"
    run_tfim_example(lib)


def test_tfim_mc_measure_example_hostsim(tmp_path):
    """examples/tfim_mc_measure.py (the reference's transverse_field_ising_mc_measure.cpp flow): the optimised state of the
    VMC example measures an energy consistent with its last optimisation energies, E = diag - h sum sigma_x holds for the
    means, and DumpData writes the reference's stats layout."""
    import tfim_vmc_optimize as ex
    import tfim_mc_measure as mm
    lib = hostsim_lib.load()
    energies, state = ex.optimize(rows=3, cols=3, D=2, chi=4, h=0.5, walkers=8, samples=160, iters=6, step=0.15, lib=lib,
                                  log=lambda *_: None)
    r = mm.measure(state=state, rows=3, cols=3, D=2, chi=4, h=0.5, walkers=8, samples=240, out=str(tmp_path / "m"), lib=lib)
    assert set(r) == {"energy", "spin_z", "sigma_x", "SzSz_row"}
    assert abs(r["energy"][0] - energies[-1]) < 2.0 and r["energy"][0] < energies[0]
    assert np.all(r["sigma_x"][0] > 0)                                     # a positive state: <sigma_x> > 0 on every site
    for f in ("energy.csv", "sigma_x_mean.csv", "sigma_x_stderr.csv", "spin_z_mean.csv"):
        assert os.path.exists(tmp_path / "m" / "stats" / f), f


def test_heisenberg_vmc_optimize_example_hostsim():
    """examples/heisenberg_vmc_optimize.py (the flow of the reference's integration test test_square_heisenberg_obc.cpp: SR
    optimisation, then measurement) on the 2x2 simple-update fixture of K4: a few SR steps move the energy from -1.9952 towards
    the exact -2, and the measured energy of the optimised state lies between them."""
    import heisenberg_vmc_optimize as ex
    from helpers import load_golden_tps
    from peps_b200.api import SplitIndexTPS, Configuration
    lib = hostsim_lib.load()
    tps, z = load_golden_tps("heis2x2_double_su")
    e_start = float(z["exp_energy"])                                       # -1.99521278793
    energies, state = ex.optimize(SplitIndexTPS(tps), 2, 2, chi=16, walkers=16, samples=1600, iters=6, step=0.3, lib=lib,
                                  log=lambda *_: None, init=Configuration(np.array([[0, 1], [1, 0]])))
    obs = ex.measure(state, 2, 2, chi=16, walkers=16, samples=3200, lib=lib)
    e_final, err = float(obs["energy"][0]), float(obs["energy"][1])
    assert -2.0 - 4 * err - 1e-9 <= e_final < e_start - 1e-3, (e_start, energies, e_final, err)


@pytest.mark.parametrize("name,exact", [("sf2x2_t2_+0.0_double_su", -2.0), ("tj2x2_double_su", -2.9431635706137875),
                                        ("sf2x2_t2_+0.0_complex_su", -2.0), ("tj2x2_complex_su", -2.9431635706137875)])
def test_fermion_vmc_optimisation_reaches_the_exact_ground_energy_hostsim(name, exact):
    """End to end in fermion mode (the flow of the reference's integration tests test_square_nn_spinless_free_fermion.cpp /
    test_square_tj_model.cpp): SR from the reference's 2x2 simple-update fZ2 states (K8 energies -1.98218 / -2.78008) descends to
    the exact ground energies (-2 and -2.94316, the `lowest` fixtures of K8) -- the fermionic O* / gradient and the device CG
    point the right way, and the optimised FermionSplitIndexTPS measures that energy."""
    import heisenberg_vmc_optimize as ex
    from test_fermion_oracle import load_golden
    from peps_b200.api import FermionSplitIndexTPS, Configuration, TableModel
    lib = hostsim_lib.load()
    f, z = load_golden(name)
    ftps = FermionSplitIndexTPS(f.T, f.par, f.phys_par)
    if name.startswith("sf"):
        model, init = TableModel.spinless_fermion(1.0, 0.0, 0.0), [[0, 1], [1, 0]]
    else:
        model, init = TableModel.tj(1.0, 0.3, V=0.075), [[0, 2], [2, 1]]                  # SquaretJVModel(t, 0, J, J/4, mu = 0)
    init = Configuration(np.array(init))
    energies, state = ex.optimize(ftps, 2, 2, chi=8, walkers=16, samples=1600, iters=8, step=0.3, lib=lib, log=lambda *_: None,
                                  init=init, model=model)
    assert isinstance(state, FermionSplitIndexTPS)
    obs = ex.measure(state, 2, 2, chi=8, walkers=16, samples=3200, lib=lib, init=init, model=model)
    assert abs(np.imag(obs["energy"][0])) < 1e-9                    # the complex fixtures are the real states times a phase
    e, err = float(np.real(obs["energy"][0])), float(obs["energy"][1])
    start = float(z["exp_energy"])
    assert exact - 5 * err - 1e-6 <= e < start - 0.9 * (start - exact) + 5 * err, (start, energies, e, err, exact)


def test_complex_state_vmc_optimisation_hostsim():
    """The same flow on a COMPLEX state (the 2x2 K4 fixture with random phases): the complex gradient
    sum conj(E_loc) O* / N - conj(E) sum O* / N and the complex SR (real embedding of the O* samples, planar CG) drive the energy
    from -1.78 to the exact -2 and its imaginary part to zero -- an end-to-end check of the conjugation conventions."""
    import heisenberg_vmc_optimize as ex
    from helpers import load_golden_tps
    from peps_b200.api import SplitIndexTPS, Configuration
    lib = hostsim_lib.load()
    tps, _ = load_golden_tps("heis2x2_double_su")
    rng = np.random.default_rng(4)
    ctps = [[[x * np.exp(1j * 0.3 * rng.standard_normal(x.shape)) + 0.05j * rng.standard_normal(x.shape) * np.max(np.abs(x))
              for x in site] for site in row] for row in tps]
    energies, state = ex.optimize(SplitIndexTPS(ctps), 2, 2, chi=16, walkers=16, samples=1600, iters=10, step=0.3, lib=lib,
                                  log=lambda *_: None, init=Configuration(np.array([[0, 1], [1, 0]])))
    assert np.iscomplexobj(state.pack()) and energies[0] > -1.9
    obs = ex.measure(state, 2, 2, chi=16, walkers=16, samples=3200, lib=lib)
    e, err = complex(obs["energy"][0]), float(obs["energy"][1])
    assert -2.0 - 5 * err - 1e-6 <= e.real < -1.999 and abs(e.imag) < 2e-3, (energies, e, err)
