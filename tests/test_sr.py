"""Stochastic reconfiguration: the device-resident O* store + S-matrix matvec (host-simulated device ops here, the
same scenario runs on the GPU in test_gpu_parity.py) against the oracle's dense S matrix and CG."""
import numpy as np
import pytest

import hostsim_lib
from oracle import vmc, sr as osr
from peps_b200 import sr
from peps_b200.api import (BMPSTruncateParams, SplitIndexTPS, MCEnergyGradEvaluator, MonteCarloParams,
                           SquareSpinOneHalfXXZModelOBC, MCUpdateSquareNNExchange)


def sr_scenario(lib, rows=3, cols=3, D=2, W=3, n=4, chi=4, diag_shift=1e-3):
    tps_l = vmc.random_tps(rows, cols, 2, D, seed=8)
    tps = SplitIndexTPS(tps_l)
    cfgs = np.stack([vmc.shuffled_half_filled_config(rows, cols, 60 + w) for w in range(W)])
    mc = MonteCarloParams(num_samples=n * W, num_warmup_sweeps=0, sweeps_between_samples=1, is_warmed_up=True)
    ev = MCEnergyGradEvaluator(mc, BMPSTruncateParams.SVD(chi, chi, 0.0), tps, SquareSpinOneHalfXXZModelOBC(1, 1, 0),
                               MCUpdateSquareNNExchange(31), walkers=W, configs=cfgs, lib=lib)
    res = ev.Evaluate(collect_sr_buffers=True)
    assert ev.batch.sr_count() == n * W and res.total_samples == n * W
    # oracle: the same chains, dense O* vectors
    oev = vmc.EnergyGradEvaluator(tps_l, vmc.XXZModel(), (chi, chi, 0.0), 1)
    ostars = []
    for w in range(W):
        rr = oev.sample_rank(vmc.Walker(tps_l, cfgs[w], (chi, chi, 0.0)), vmc.NNExchangeUpdater(31 + w), n, collect_sr=True)
        ostars += [osr.dense_ostar(s, tps_l) for s in rr["ostar_samples"]]
    obar = sum(ostars) / len(ostars)
    assert np.max(np.abs(res.Ostar_mean.pack() - obar)) < 1e-12 * np.max(np.abs(obar))
    rng = np.random.default_rng(0)
    v = rng.standard_normal(obar.size)
    smat = sr.SRSMatrix(ev.batch, res.Ostar_mean.pack(), res.total_samples, diag_shift)
    ref = osr.s_matvec(ostars, obar, v, diag_shift)
    assert np.max(np.abs(smat(v) - ref)) < 1e-11 * np.max(np.abs(ref))
    # natural gradient: product CG on the device matvec == oracle CG on the dense matvec == dense solve
    g = res.gradient.pack()
    params = sr.ConjugateGradientParams(max_iter=200, relative_tolerance=1e-10)
    nat, iters, resid = ev.CalculateNaturalGradient(res, diag_shift, params)
    x_ref, r_ref, it_ref = osr.cg(lambda x: osr.s_matvec(ostars, obar, x, diag_shift), g, np.zeros_like(g), 200, 1e-10)
    assert iters == it_ref
    assert np.max(np.abs(nat.pack() - x_ref)) < 1e-8 * np.max(np.abs(x_ref))
    o = np.stack(ostars)
    s_dense = (o - obar).T @ (o - obar) / len(ostars) + diag_shift * np.eye(obar.size)
    x_dense = np.linalg.solve(s_dense, g)
    assert np.max(np.abs(nat.pack() - x_dense)) < 1e-6 * np.max(np.abs(x_dense))
    return iters


def test_sr_matvec_and_natural_gradient_hostsim():
    assert sr_scenario(hostsim_lib.load()) > 0


def test_cg_reference_exits():
    a = np.diag([1.0, 2.0, 3.0])
    b = np.array([1.0, 1.0, 1.0])
    r = sr.conjugate_gradient(lambda v: a @ v, b, np.zeros(3), sr.ConjugateGradientParams(relative_tolerance=1e-12))
    assert r.reason == sr.CONVERGED and np.allclose(r.x, [1, 0.5, 1 / 3])
    r = sr.conjugate_gradient(lambda v: -(a @ v), b, np.zeros(3), sr.ConjugateGradientParams())
    assert r.reason == sr.INDEFINITE_MATRIX
    r = sr.conjugate_gradient(lambda v: a @ v, b, np.linalg.solve(a, b), sr.ConjugateGradientParams())
    assert r.iterations == 0 and r.reason == sr.CONVERGED


def test_complex_sr_matvec_and_natural_gradient_hostsim():
    from parity_common import run_complex_sr
    assert run_complex_sr(hostsim_lib.load()) > 0
