"""Statistical and conservation checks in the style of the reference's tests/test_algorithm/test_mc_energy_grad_evaluator.cpp
(Monte Carlo estimates vs exact summation within a few sigma) and tests/test_monte_carlo_tools/test_mc_updater_conservation.cpp
(U(1) conservation of the exchange updaters), through the C ABI on the host simulation."""
import itertools

import numpy as np
import pytest

import hostsim_lib
from helpers import load_golden_tps
from oracle import vmc
from peps_b200.api import (BMPSTruncateParams, Configuration, MCEnergyGradEvaluator, MonteCarloParams, MCUpdateSquareNNExchange,
                           MCUpdateSquareTNN3SiteExchange, SplitIndexTPS, SquareSpinOneHalfXXZModelOBC, WalkerBatch)


@pytest.fixture(scope="module")
def lib():
    return hostsim_lib.load()


def exact_energy_and_gradient(tps):
    """Exact summation over the six S_z = 0 configurations of a 2x2 state with the oracle (the reference's
    ExactSumEnergyEvaluator, tests/test_algorithm/test_exact_summation_evaluator.cpp)."""
    cfgs = [np.array(p).reshape(2, 2) for p in sorted(set(itertools.permutations([0, 0, 1, 1])))]
    model = vmc.XXZModel()
    wsum = esum = 0.0
    osum = [[[np.zeros_like(x) for x in site] for site in row] for row in tps]
    eosum = [[[np.zeros_like(x) for x in site] for site in row] for row in tps]
    for cfg in cfgs:
        w = vmc.Walker(tps, cfg, (1, 1000, 0.0))
        e, holes, _ = model.energy_and_holes(tps, w, True)
        wt = abs(w.amplitude) ** 2
        wsum += wt
        esum += wt * e
        for r in range(2):
            for c in range(2):
                s = int(cfg[r, c])
                o = holes[r][c] / w.amplitude
                osum[r][c][s] += wt * o
                eosum[r][c][s] += wt * e * o
    energy = esum / wsum
    grad = np.concatenate([(eosum[r][c][s] / wsum - energy * osum[r][c][s] / wsum).ravel() for r in range(2) for c in range(2) for s in range(2)])
    return energy, grad


def test_mc_energy_and_gradient_agree_with_exact_summation(lib):
    """2x2 Heisenberg simple-update fixture (exact energy -1.99521278793, K4): the Monte Carlo estimate of the evaluator lies
    within 4 standard errors of the exact summation and the sampled gradient converges to the exact one."""
    tps, z = load_golden_tps("heis2x2_double_su")
    e_exact, g_exact = exact_energy_and_gradient(tps)
    assert abs(e_exact - float(z["exp_energy"])) < 1e-9
    W, n = 16, 250
    D = max(max(x.shape) for row in tps for site in row for x in site)
    mc = MonteCarloParams(num_samples=W * n, num_warmup_sweeps=20, sweeps_between_samples=1,
                          initial_config=Configuration(np.array([[0, 1], [1, 0]])))
    ev = MCEnergyGradEvaluator(mc, BMPSTruncateParams.SVD(1, 1000, 0.0), SplitIndexTPS(tps), SquareSpinOneHalfXXZModelOBC(1, 1, 0),
                               MCUpdateSquareNNExchange(seed=2024), W, lib=lib)
    ev.WarmUp()
    res = ev.Evaluate()
    f = ev.batch.get_tps_flat()[0] / SplitIndexTPS(tps).pack()[0]          # WarmUp rescaled every site tensor by one factor
    assert res.energy_error < 0.02
    assert abs(res.energy - e_exact) < 4 * res.energy_error + 1e-6, (res.energy, e_exact, res.energy_error)
    g = res.gradient.pack() * f                                            # d/dT of the rescaled state = (1/f) d/dT: compare like with like
    rel = np.linalg.norm(g - g_exact) / np.linalg.norm(g_exact)
    assert rel < 0.35, rel                                                 # 4000 samples: the direction of the exact gradient
    assert float(np.dot(g, g_exact)) / (np.linalg.norm(g) * np.linalg.norm(g_exact)) > 0.9


@pytest.mark.parametrize("updater", [MCUpdateSquareNNExchange(seed=5), MCUpdateSquareTNN3SiteExchange(seed=5)])
def test_exchange_updaters_conserve_sz(lib, updater):
    """test_mc_updater_conservation.cpp: the exchange updaters keep the number of up spins of every walker."""
    rows, cols, D, W = 4, 4, 2, 4
    tps = vmc.random_tps(rows, cols, 2, D, seed=3)
    cfgs = np.stack([vmc.shuffled_half_filled_config(rows, cols, 40 + w) for w in range(W)])
    cfgs[1, 0, 0] = cfgs[1, 0, 1] = 1                                      # a walker away from half filling
    before = cfgs.sum(axis=(1, 2))
    b = WalkerBatch(rows, cols, 2, D, W, BMPSTruncateParams.SVD(4, 4, 0.0), lib=lib)
    b.set_tps(SplitIndexTPS(tps)); b.set_configs(cfgs); b.seed_rng(np.arange(W) + 1); b.set_updater(updater)
    b.init_walkers()
    moved = 0.0
    for _ in range(4):
        moved += float(np.sum(b.sweep(1)))
    after = b.get_configs().sum(axis=(1, 2))
    assert moved > 0 and np.array_equal(before, after)
    b.close()


@pytest.mark.parametrize("complex_", [False, True])
def test_split_index_tps_vector_space(complex_):
    """SplitIndexTPS as the gradient / O* / CG vector type (two_dim_tn/tps/split_index_tps.h:171-177, 370-377): + - * scalar,
    operator*(a, b) = sum conj(a) b, NormSquare, pack / unpack round trip (tests/test_2d_tn/test_split_index_tps.cpp)."""
    rng = np.random.default_rng(1)

    def rnd():
        t = vmc.random_tps(3, 4, 2, 3, seed=int(rng.integers(1 << 30)))
        if complex_:
            t = [[[x + 1j * rng.standard_normal(x.shape) for x in site] for site in row] for row in t]
        return SplitIndexTPS(t)
    a, b = rnd(), rnd()
    fa, fb = a.pack(), b.pack()
    assert np.iscomplexobj(fa) == complex_
    assert np.allclose((a + b).pack(), fa + fb) and np.allclose((a - b).pack(), fa - fb)
    s = (0.3 - 0.2j) if complex_ else 0.3
    assert np.allclose((a * s).pack(), fa * s) and np.allclose((s * a).pack(), fa * s)
    assert abs((a * b) - np.vdot(fa, fb)) < 1e-12 * abs(np.vdot(fa, fb))          # conjugates the LEFT operand
    assert abs(a.NormSquare() - np.vdot(fa, fa).real) < 1e-12 * a.NormSquare()
    back = SplitIndexTPS.unpack(fa, a)
    assert all(np.array_equal(x, y) for ra, rb in zip(a.t, back.t) for sa, sb in zip(ra, rb) for x, y in zip(sa, sb))
    assert a.rows() == 3 and a.cols() == 4 and a.PhysicalDim() == 2 and a.bond_dim() == 3
