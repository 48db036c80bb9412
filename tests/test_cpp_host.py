"""The C++ host wrapper (include/peps_b200.hpp, the reference-shaped interface a C++ driver would include) compiled
with g++ against the host-simulation library must reproduce the Python mirror's Evaluate()."""
import os
import subprocess
import tempfile

import numpy as np

import hostsim_lib
from oracle import vmc
from peps_b200.api import (BMPSTruncateParams, SplitIndexTPS, MCEnergyGradEvaluator, MonteCarloParams, Configuration,
                           SquareSpinOneHalfXXZModelOBC, MCUpdateSquareNNExchange)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_cpp_wrapper_case(lib, libdir, libfile, extra_link=()):
    """Builds tests/cpp/test_cpp_host.cpp against `libfile` in `libdir` and compares its Evaluate() with the Python
    mirror driving `lib` (the ctypes binding of the same shared library)."""
    tps = SplitIndexTPS(vmc.random_tps(3, 3, 2, 2, seed=4))
    cfg = vmc.neel_config(3, 3)
    W, chi, ns, seed = 3, 4, 6, 21
    with tempfile.TemporaryDirectory() as td:
        exe = os.path.join(td, "cpp_host")
        subprocess.check_call(["g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests", "cpp", "test_cpp_host.cpp"), "-o", exe,
                               "-L" + libdir, "-l:" + libfile, "-Wl,-rpath," + libdir] + list(extra_link))
        flat = tps.pack()
        inp = f"3 3 2 2 {W} {chi} {ns} {seed}\n{flat.size}\n" + " ".join(repr(float(x)) for x in flat) + "\n" + \
              " ".join(str(int(c)) for c in cfg.ravel()) + "\n"
        out = subprocess.run([exe], input=inp, capture_output=True, text=True, check=True).stdout.split()
    (e_cpp, err_cpp, gn_cpp, acc_cpp, e2, e_meas, e_bonds, d_table, n_rescued, e_only, acc_only, ng2, ng_it, sr_n,
     m_e, m_err, m_bh0, m_npairs) = map(float, out)
    assert abs(e_only - e_cpp) < 1e-12 and abs(acc_only - acc_cpp) < 1e-12     # EvaluateEnergyOnly: same chains, no holes
    assert d_table < 1e-13 and n_rescued == 0
    assert abs(e_meas - e_bonds) < 1e-10 * max(1.0, abs(e_meas))
    mc = MonteCarloParams(num_samples=ns, num_warmup_sweeps=0, sweeps_between_samples=1, initial_config=Configuration(cfg),
                          is_warmed_up=True)
    ev = MCEnergyGradEvaluator(mc, BMPSTruncateParams.SVD(chi, chi, 0.0), tps, SquareSpinOneHalfXXZModelOBC(1, 1, 0),
                               MCUpdateSquareNNExchange(seed), walkers=W, lib=lib)
    res = ev.Evaluate()
    assert abs(e_cpp - res.energy) < 1e-12
    assert np.isfinite(e2)          # second Evaluate through the seam-B1 adapter continues the same chains
    assert abs(gn_cpp - res.gradient_norm) < 1e-12 * max(1.0, res.gradient_norm)
    assert abs(acc_cpp - res.accept_rates_avg[0]) < 1e-12
    ev2 = MCEnergyGradEvaluator(mc, BMPSTruncateParams.SVD(chi, chi, 0.0), tps, SquareSpinOneHalfXXZModelOBC(1, 1, 0),
                                MCUpdateSquareNNExchange(seed), walkers=W, lib=lib)
    # SR and the measurer through the C++ wrapper against the Python mirror on the same chains
    from peps_b200 import sr
    from peps_b200.api import MCPEPSMeasurer
    ev3 = MCEnergyGradEvaluator(mc, BMPSTruncateParams.SVD(chi, chi, 0.0), tps, SquareSpinOneHalfXXZModelOBC(1, 1, 0),
                                MCUpdateSquareNNExchange(seed), walkers=W, lib=lib)
    r3 = ev3.Evaluate(collect_sr_buffers=True)
    nat, iters, _ = ev3.CalculateNaturalGradient(r3, 1e-3, sr.ConjugateGradientParams(max_iter=200, relative_tolerance=1e-10))
    assert int(ng_it) == iters and int(sr_n) == ev3.batch.sr_count()
    assert abs(ng2 - nat.NormSquare()) < 1e-9 * nat.NormSquare()
    pm = MCPEPSMeasurer(mc, BMPSTruncateParams.SVD(chi, chi, 0.0), tps, SquareSpinOneHalfXXZModelOBC(1, 1, 0),
                        MCUpdateSquareNNExchange(seed), W, lib=lib, enable_structure_factor=True).Execute()
    assert abs(m_e - pm["energy"][0]) < 1e-12 and abs(m_err - pm["energy"][1]) < 1e-12
    assert abs(m_bh0 - pm["bond_energy_h"][0][0, 0]) < 1e-12 and int(m_npairs) == pm["SpSm_cross"][0].size
    e_py, err_py, acc_py = ev2.EvaluateEnergyOnly()
    assert abs(e_py - res.energy) < 1e-12 and abs(err_py - res.energy_error) < 1e-12 and abs(acc_py[0] - res.accept_rates_avg[0]) < 1e-12


def test_cpp_wrapper_matches_python_mirror():
    run_cpp_wrapper_case(hostsim_lib.load(), os.path.join(ROOT, "tests", "hostsim"), "libpeps_hostsim.so")


def run_cpp_complex_case(lib, libdir, libfile, extra_link=()):
    """tests/cpp/test_cpp_complex.cpp: the C++ wrapper's complex Evaluate against the Python mirror on the same library."""
    from parity_common import complex_tps
    tps = SplitIndexTPS(complex_tps(3, 3, 2, 4))
    cfg = vmc.neel_config(3, 3)
    W, chi, ns, seed = 3, 4, 6, 21
    with tempfile.TemporaryDirectory() as td:
        exe = os.path.join(td, "cpp_complex")
        subprocess.check_call(["g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests", "cpp", "test_cpp_complex.cpp"), "-o", exe,
                               "-L" + libdir, "-l:" + libfile, "-Wl,-rpath," + libdir] + list(extra_link))
        flat = tps.pack()
        inp = f"3 3 2 2 {W} {chi} {ns} {seed}\n{flat.size}\n" + " ".join(f"{float(x.real)!r} {float(x.imag)!r}" for x in flat) + "\n" + \
              " ".join(str(int(c)) for c in cfg.ravel()) + "\n"
        out = subprocess.run([exe], input=inp, capture_output=True, text=True, check=True).stdout.split()
    er, ei, err, gn, acc = map(float, out)
    mc = MonteCarloParams(num_samples=ns, num_warmup_sweeps=0, sweeps_between_samples=1, initial_config=Configuration(cfg),
                          is_warmed_up=True)
    ev = MCEnergyGradEvaluator(mc, BMPSTruncateParams.SVD(chi, chi, 0.0), tps, SquareSpinOneHalfXXZModelOBC(1, 1, 0),
                               MCUpdateSquareNNExchange(seed), walkers=W, lib=lib)
    res = ev.Evaluate()
    assert abs(complex(er, ei) - res.energy) < 1e-12 and abs(ei) > 1e-6
    assert abs(gn - res.gradient_norm) < 1e-12 * max(1.0, res.gradient_norm)
    assert abs(acc - res.accept_rates_avg[0]) < 1e-12


def test_cpp_wrapper_complex_matches_python_mirror():
    run_cpp_complex_case(hostsim_lib.load(), os.path.join(ROOT, "tests", "hostsim"), "libpeps_hostsim.so")
