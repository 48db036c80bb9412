"""Pins the CPU oracle against the reference's own known-answer tests
before it is trusted as the checker for the CUDA path."""
import math

import numpy as np
import pytest

from oracle.bmps import LEFT, RIGHT, UP, DOWN, HORIZONTAL, VERTICAL, truncation_dim
from oracle.contractor import BMPSContractor
from oracle.mt19937 import MT19937
from oracle import vmc
from helpers import (load_golden_tps, ising_tn, ising_exact_logZ, exact_summation, grad_norm_square,
                     weighted_probe)


def test_mt19937_known_answer():
    # ISO C++ [rand.predef]: the 10000th consecutive invocation of a default-constructed mt19937 is 4123659995.
    g = MT19937(5489)
    for _ in range(9999):
        g.next_u32()
    assert g.next_u32() == 4123659995
    # numpy's legacy RandomState uses the same init_genrand seeding; its 53-bit doubles are a different
    # construction, but the raw 32-bit stream must agree.
    g = MT19937(12345)
    rs = np.random.RandomState(12345)
    ref = rs.randint(0, 2 ** 32, size=1000, dtype=np.uint64)
    assert [g.next_u32() for _ in range(1000)] == [int(x) for x in ref]


def test_uniform01_is_generate_canonical():
    g, h = MT19937(7), MT19937(7)
    for _ in range(100):
        x0, x1 = h.next_u32(), h.next_u32()
        assert g.uniform01() == (x0 + x1 * 4294967296.0) / 18446744073709551616.0


def test_truncation_rule():
    s = np.array([1.0, 0.5, 0.1, 1e-9, 0.0])
    assert truncation_dim(s, 1, 10, 0.0) == 3          # weights below double rounding of the total are dropped down to Dmin
    assert truncation_dim(s, 5, 10, 0.0) == 5
    assert truncation_dim(s, 1, 2, 0.0) == 2
    assert truncation_dim(s, 1, 10, 1e-15) == 3
    assert truncation_dim(s, 1, 10, 0.009) == 2


@pytest.mark.parametrize("L", [6, 12])
def test_K1_ising_partition_function(L):
    """reference: tests/test_2d_tn/test_bmps_contractor.cpp:472-493, SVD(10, 30, 1e-15), tol 1e-8 on F/site."""
    beta = math.log(1 + math.sqrt(2.0)) / 2.0
    tn = ising_tn(L, beta)
    lz = ising_exact_logZ(L, beta)
    c = BMPSContractor(L, L)
    c.init(tn)
    c.set_truncate_params(10, 30, 1e-15)
    zs = []
    c.grow_bmps_for_row(tn, 2)
    c.init_bten(tn, LEFT, 2)
    c.grow_full_bten(tn, RIGHT, 2, 2, True)
    zs.append(c.trace(tn, (2, 0), HORIZONTAL))
    c.shift_bten_window(tn, RIGHT)
    zs.append(c.trace(tn, (2, 1), HORIZONTAL))
    c.init_bten(tn, LEFT, 2)
    c.grow_full_bten(tn, RIGHT, 2, 1, True)          # remain=1: the E_loc flow, enables one-site traces
    zs.append(c.replace_one_site_trace(tn, (2, 0), tn[2][0], HORIZONTAL))
    c.shift_bten_window(tn, RIGHT)
    zs.append(c.replace_one_site_trace(tn, (2, 1), tn[2][1], HORIZONTAL))
    c.grow_bmps_for_col(tn, 1)
    c.init_bten(tn, UP, 1)
    c.grow_full_bten(tn, DOWN, 1, 2, True)
    zs.append(c.trace(tn, (0, 1), VERTICAL))
    c.shift_bten_window(tn, DOWN)
    zs.append(c.trace(tn, (1, 1), VERTICAL))
    c.init_bten(tn, DOWN, 1)
    c.grow_full_bten(tn, UP, 1, 2, True)
    zs.append(c.trace(tn, (L - 2, 1), VERTICAL))
    for z in zs:
        assert abs((math.log(z) - lz) / (L * L * beta)) < 1e-8


def test_K1_three_site_traces_reproduce_partition_function():
    """ReplaceTNNSiteTrace closures of the reference's K1 list (tests/test_2d_tn/test_bmps_contractor.cpp:273-405):
    with the original tensors every closure is the partition function."""
    L = 8
    beta = math.log(1 + math.sqrt(2.0)) / 2.0
    tn = ising_tn(L, beta)
    lz = ising_exact_logZ(L, beta)
    c = BMPSContractor(L, L)
    c.init(tn)
    c.set_truncate_params(10, 30, 1e-15)
    zs = []
    c.grow_bmps_for_row(tn, 3)
    c.init_bten(tn, LEFT, 3)
    c.grow_full_bten(tn, RIGHT, 3, 3, True)
    zs.append(c.replace_tnn_site_trace(tn, (3, 0), HORIZONTAL, tn[3][0], tn[3][1], tn[3][2]))
    c.shift_bten_window(tn, RIGHT)
    zs.append(c.replace_tnn_site_trace(tn, (3, 1), HORIZONTAL, tn[3][1], tn[3][2], tn[3][3]))
    c.grow_bmps_for_col(tn, 2)
    c.init_bten(tn, UP, 2)
    c.grow_full_bten(tn, DOWN, 2, 3, True)
    zs.append(c.replace_tnn_site_trace(tn, (0, 2), VERTICAL, tn[0][2], tn[1][2], tn[2][2]))
    for z in zs:
        assert abs((math.log(z) - lz) / (L * L * beta)) < 1e-8


def test_K3_rectangular_ising_partition_function():
    """reference: OBCIsing2DZ2TenNet (tests/test_2d_tn/test_bmps_contractor.cpp:499-686): 24 rows x 10 columns at the
    critical point, SVD(1, 10, 1e-15), free energy per site to 1e-8. Dense restatement of the Z2-blocked fixture;
    closures along rows and along columns of the rectangular lattice."""
    rows, cols = 24, 10
    beta = math.log(1 + math.sqrt(2.0)) / 2.0
    tn = ising_tn(rows, beta, cols)
    lz = ising_exact_logZ(cols, beta, rows)          # transfer matrix across the short side
    c = BMPSContractor(rows, cols)
    c.init(tn)
    c.set_truncate_params(1, 10, 1e-15)
    zs = []
    c.grow_bmps_for_row(tn, 11)
    c.init_bten(tn, LEFT, 11)
    c.grow_full_bten(tn, RIGHT, 11, 2, True)
    zs.append(c.trace(tn, (11, 0), HORIZONTAL))
    c.grow_bmps_for_col(tn, 4)
    c.init_bten(tn, UP, 4)
    c.grow_full_bten(tn, DOWN, 4, 2, True)
    zs.append(c.trace(tn, (0, 4), VERTICAL))
    for z in zs:
        assert abs((math.log(z) - lz) / (rows * cols * beta)) < 1e-8


def test_K1_nnn_traces_reproduce_partition_function():
    """The NNN closures of the reference's K1 list (test_bmps_contractor.cpp:312-335): ReplaceNNNSiteTrace with the
    original tensors equals Z, before and after ShiftBTen2Window."""
    L = 8
    beta = math.log(1 + math.sqrt(2.0)) / 2.0
    tn = ising_tn(L, beta)
    lz = ising_exact_logZ(L, beta)
    c = BMPSContractor(L, L)
    c.init(tn)
    c.set_truncate_params(10, 30, 1e-15)
    c.grow_bmps_for_row(tn, 1)
    c.init_bten2(tn, LEFT, 1)
    c.grow_full_bten2(tn, RIGHT, 1, 2, True)
    zs = [c.replace_nnn_site_trace(tn, (1, 0), 1, HORIZONTAL, tn[2][0], tn[1][1]),
          c.replace_nnn_site_trace(tn, (1, 0), 0, HORIZONTAL, tn[1][0], tn[2][1])]
    c.shift_bten2_window(tn, RIGHT, 1)
    zs += [c.replace_nnn_site_trace(tn, (1, 1), 1, HORIZONTAL, tn[2][1], tn[1][2]),
           c.replace_nnn_site_trace(tn, (1, 1), 0, HORIZONTAL, tn[1][1], tn[2][2])]
    for z in zs:
        assert abs((math.log(z) - lz) / (L * L * beta)) < 1e-8


def test_j1j2_local_energy_equals_brute_force():
    tps = vmc.random_tps(3, 3, 2, 2, seed=3)
    cfg = vmc.shuffled_half_filled_config(3, 3, 5)
    trunc = (1, 1000, 0.0)
    amp = lambda c: vmc.Walker(tps, c, trunc).amplitude
    e, _, _ = vmc.XXZModel(1.0, 1.0, 0.0, 0.5, 0.5).energy_and_holes(tps, vmc.Walker(tps, cfg, trunc), True)
    psi = amp(cfg)

    def bond(s1, s2, jz, jxy):
        if cfg[s1] == cfg[s2]:
            return 0.25 * jz
        c2 = cfg.copy()
        c2[s1], c2[s2] = cfg[s2], cfg[s1]
        return -0.25 * jz + 0.5 * jxy * amp(c2) / psi

    ref = 0.0
    for r in range(3):
        for c in range(3):
            if c < 2:
                ref += bond((r, c), (r, c + 1), 1, 1)
            if r < 2:
                ref += bond((r, c), (r + 1, c), 1, 1)
            if r < 2 and c < 2:
                ref += bond((r, c), (r + 1, c + 1), 0.5, 0.5) + bond((r + 1, c), (r, c + 1), 0.5, 0.5)
    assert abs(e - ref) < 1e-12


def test_K2_punch_hole_and_invalidation():
    """reference: tests/test_2d_tn/test_bmps_contractor.cpp:407-470."""
    L = 8
    beta = math.log(1 + math.sqrt(2.0)) / 2.0
    tn = ising_tn(L, beta)
    c = BMPSContractor(L, L)
    c.init(tn)
    c.set_truncate_params(4, 10, 1e-10)
    c.grow_bmps_for_row(tn, 2)
    c.grow_full_bten(tn, LEFT, 2, 2, True)
    c.grow_full_bten(tn, RIGHT, 2, 2, True)
    val1 = c.trace(tn, (2, 0), HORIZONTAL)
    hole = c.punch_hole(tn, (2, 1), HORIZONTAL)
    tr = c.trace(tn, (2, 1), HORIZONTAL)
    assert abs(np.sum(hole * tn[2][1]) - tr) < 1e-10 * abs(tr)
    tn[2][1] = tn[2][1] * 0.5
    c.erase_envs_after_update((2, 1))
    c.grow_bmps_for_row(tn, 2)
    c.grow_full_bten(tn, LEFT, 2, 2, True)
    c.grow_full_bten(tn, RIGHT, 2, 2, True)
    val2 = c.trace(tn, (2, 0), HORIZONTAL)
    assert abs(val2 - 0.5 * val1) < 1e-10 * abs(val1)


@pytest.mark.parametrize("name,complex_", [("heis2x2_double_lowest", False), ("heis2x2_complex_lowest", True),
                                            ("heis2x2_double_su", False)])
def test_K4_K5_exact_summation_goldens(name, complex_):
    """Energies and gradient signatures asserted by tests/test_algorithm/test_exact_summation_evaluator.cpp:531-606."""
    tps, z = load_golden_tps(name)
    energy, grad = exact_summation(tps, vmc.XXZModel(1.0, 1.0, 0.0))
    assert abs(energy - float(z["exp_energy"])) < float(z["exp_energy_tol"])
    if "exp_grad_norm" in z:
        # the reference's tolerance is abs 1e-8; the oracle reproduces the quoted digits to ~1e-18 abs
        assert abs(grad_norm_square(grad) - float(z["exp_grad_norm"])) < 1e-17
        p = weighted_probe(grad, complex_)
        assert abs(np.real(p) - float(z["exp_grad_probe_re"])) < 1e-17
        assert abs(np.imag(p) - float(z["exp_grad_probe_im"])) < 1e-17


@pytest.mark.parametrize("name", ["tfim2x2_double_lowest", "tfim2x2_double_su"])
def test_K4_K5_transverse_ising_goldens(name):
    """TrivialTransverseIsingTest (tests/test_algorithm/test_exact_summation_evaluator.cpp:700-790): J = h = 1, all 16
    configurations, SVD(1, 8, 1e-16); energies -5.19991995228 (quoted) / 2x2 ED, gradient signatures of the lowest
    state."""
    tps, z = load_golden_tps(name)
    energy, grad = exact_summation(tps, vmc.TFIMModel(1.0), (1, 8, 1e-16), all_configs=True)
    assert abs(energy - float(z["exp_energy"])) < float(z["exp_energy_tol"])
    if "exp_grad_norm" in z:
        assert abs(grad_norm_square(grad) - float(z["exp_grad_norm"])) < 1e-19
        assert abs(np.real(weighted_probe(grad, False)) - float(z["exp_grad_probe_re"])) < 1e-19


def test_K9_suwa_todo_is_stationary_and_rejection_free():
    """SuwaTodoStateUpdate (suwa_todo_update.h:53-113): the update leaves pi ~ weights invariant (balance condition) and
    the heaviest state never stays put when it holds less than half of the weight."""
    from oracle.mt19937 import MT19937
    weights = [0.3, 1.7, 0.9, 0.45]
    n, trials = len(weights), 4000
    rng = MT19937(99)
    flow = np.zeros((n, n))
    for i in range(n):
        for _ in range(trials):
            flow[i, vmc.suwa_todo_state_update(i, weights, rng)] += 1.0 / trials
    pi = np.array(weights) / sum(weights)
    assert np.allclose(pi @ flow, pi, atol=0.02)
    assert flow[1, 1] < 0.05            # w_max = 1.7 < sum of the others = 1.65 + ...: rejection-free up to the surplus
    assert np.allclose(flow.sum(axis=1), 1.0)


def test_full_space_updater_samples_psi_squared():
    """MCUpdateSquareNNFullSpaceUpdateOBC on a 2x2 lattice: the visit histogram converges to |psi|^2 over all 16
    configurations (no Sz conservation)."""
    tps = vmc.random_tps(2, 2, 2, 2, seed=21)
    w = vmc.Walker(tps, vmc.neel_config(2, 2), (1, 100, 0.0))
    up = vmc.NNFullSpaceUpdater(5)
    hist = np.zeros(16)
    for _ in range(3000):
        up.sweep(tps, w)
        hist[int("".join(str(int(x)) for x in w.config.flatten()), 2)] += 1
    exact = np.zeros(16)
    for k in range(16):
        cfg = np.array([int(b) for b in format(k, "04b")]).reshape(2, 2)
        exact[k] = abs(vmc.Walker(tps, cfg, (1, 100, 0.0)).amplitude) ** 2
    exact /= exact.sum()
    assert np.max(np.abs(hist / hist.sum() - exact)) < 0.03


def test_hole_is_amplitude_derivative():
    tps = vmc.random_tps(3, 3, 2, 3, seed=3)
    cfg = vmc.neel_config(3, 3)
    w = vmc.Walker(tps, cfg, (1, 1000, 0.0))
    _, holes, psis = vmc.XXZModel().energy_and_holes(tps, w, True)
    for r in range(3):
        for c in range(3):
            assert abs(np.sum(holes[r][c] * w.tn[r][c]) - psis[r]) < 1e-12 * abs(psis[r])
    assert max(abs(p - psis[0]) for p in psis) < 1e-12 * abs(psis[0])


def test_K6_heisenberg_4x4_D8_mc_energy():
    """reference: tests/slow_tests/test_boson_mc_peps_measure.cpp:31-76 (e0_state = -9.18912, ED -9.1892).
    A short chain from the stored configurations must land within a few sigma of the state energy."""
    tps, z = load_golden_tps("heis4x4_D8_double")
    cfgs = z["configs"]
    model = vmc.XXZModel(1.0, 1.0, 0.0)
    energies = []
    for k in range(4):
        w = vmc.Walker(tps, cfgs[k], (8, 16, 1e-15))
        up = vmc.NNExchangeUpdater(100 + k)
        for _ in range(8):
            up.sweep(tps, w)
            e, _, psis = model.energy_and_holes(tps, w, False)
            energies.append(e)
            assert max(abs(p / psis[0] - 1) for p in psis) < 1e-3
    mean = float(np.mean(energies))
    err = float(np.std(energies) / math.sqrt(len(energies)))
    assert abs(mean - float(z["exp_e0_state"])) < max(5 * err, 0.05)


def test_K1_vertical_nnn_and_sqrt5_traces_reproduce_partition_function():
    """The remaining closures of the reference's K1 list (test_bmps_contractor.cpp:342-405): ReplaceSqrt5DistTwoSiteTrace
    in both link directions and both MPS orientations, and ReplaceNNNSiteTrace with VERTICAL MPS orientation, all with the
    original tensors, equal Z."""
    L = 8
    beta = math.log(1 + math.sqrt(2.0)) / 2.0
    tn = ising_tn(L, beta)
    lz = ising_exact_logZ(L, beta)
    c = BMPSContractor(L, L)
    c.init(tn)
    c.set_truncate_params(10, 30, 1e-15)
    c.grow_bmps_for_row(tn, 1)
    c.init_bten2(tn, LEFT, 1)
    c.grow_full_bten2(tn, LEFT, 1, L - 1, False)
    c.grow_full_bten2(tn, RIGHT, 1, 3, True)
    zs = [c.replace_sqrt5_dist_two_site_trace(tn, (1, 0), 1, HORIZONTAL, tn[2][0], tn[1][2]),
          c.replace_sqrt5_dist_two_site_trace(tn, (1, 1), 1, HORIZONTAL, tn[2][1], tn[1][3]),
          c.replace_sqrt5_dist_two_site_trace(tn, (1, 0), 0, HORIZONTAL, tn[1][0], tn[2][2]),
          c.replace_sqrt5_dist_two_site_trace(tn, (1, 1), 0, HORIZONTAL, tn[1][1], tn[2][3])]
    c.grow_bmps_for_col(tn, 1)
    c.grow_full_bten2(tn, DOWN, 1, 3, True)
    c.grow_full_bten2(tn, UP, 1, L - 2, True)
    zs += [c.replace_nnn_site_trace(tn, (2, 1), 1, VERTICAL, tn[3][1], tn[2][2]),
           c.replace_nnn_site_trace(tn, (2, 1), 0, VERTICAL, tn[2][1], tn[3][2]),
           c.replace_nnn_site_trace(tn, (1, 1), 1, VERTICAL, tn[2][1], tn[1][2]),
           c.replace_nnn_site_trace(tn, (1, 1), 0, VERTICAL, tn[1][1], tn[2][2]),
           c.replace_sqrt5_dist_two_site_trace(tn, (1, 1), 1, VERTICAL, tn[3][1], tn[1][2]),
           c.replace_sqrt5_dist_two_site_trace(tn, (1, 1), 0, VERTICAL, tn[1][1], tn[3][2])]
    for z in zs:
        assert abs((math.log(z) - lz) / (L * L * beta)) < 1e-8


def test_plaquette_traces_equal_amplitudes_of_exchanged_configurations():
    """On a random (asymmetric) TPS every plaquette trace must equal the amplitude evaluated from scratch on the
    configuration with the two corner spins exchanged: pins WHICH sites the replacement tensors land on."""
    rows, cols = 4, 5
    tps = vmc.random_tps(rows, cols, 2, 2, seed=12)
    cfg = vmc.shuffled_half_filled_config(rows, cols, 3)
    trunc = (1, 1000, 0.0)
    w = vmc.Walker(tps, cfg, trunc)
    c, tn = w.contractor, w.tn
    cases = [(0, 1, 1, d, o) for d in (0, 1) for o in (HORIZONTAL, VERTICAL)] + \
            [(1, 1, 1, d, o) for d in (0, 1) for o in (HORIZONTAL, VERTICAL)]
    for kind, r1, c1, d, o in cases:
        h, wd = (2, 2) if kind == 0 else ((2, 3) if o == HORIZONTAL else (3, 2))
        a, b = ((r1, c1), (r1 + h - 1, c1 + wd - 1)) if d == 0 else ((r1 + h - 1, c1), (r1, c1 + wd - 1))
        span = 2 if kind == 0 else 3
        if o == HORIZONTAL:
            c.grow_bmps_for_row(tn, r1)
            c.grow_full_bten2(tn, LEFT, r1, cols - c1, True)
            c.grow_full_bten2(tn, RIGHT, r1, c1 + span, True)
        else:
            c.grow_bmps_for_col(tn, c1)
            c.grow_full_bten2(tn, UP, c1, rows - r1, True)
            c.grow_full_bten2(tn, DOWN, c1, r1 + span, True)
        ta, tb = tps[a[0]][a[1]][int(cfg[b])], tps[b[0]][b[1]][int(cfg[a])]
        f = c.replace_nnn_site_trace if kind == 0 else c.replace_sqrt5_dist_two_site_trace
        psi = f(tn, (r1, c1), d, o, ta, tb)
        c2 = cfg.copy()
        c2[a], c2[b] = cfg[b], cfg[a]
        ref = vmc.Walker(tps, c2, trunc).amplitude
        assert abs(psi / ref - 1) < 1e-11, (kind, d, o, psi, ref)


@pytest.mark.parametrize("scheme", [1, 2])
def test_K1_variational_compression_reproduces_partition_function(scheme):
    """The reference runs K1 with the variational schemes too (test_bmps_contractor.cpp:474-485: Variational2Site /
    Variational1Site, free energy per site to 1e-8)."""
    L = 8
    beta = math.log(1 + math.sqrt(2.0)) / 2.0
    tn = ising_tn(L, beta)
    lz = ising_exact_logZ(L, beta)
    c = BMPSContractor(L, L)
    c.init(tn)
    c.set_truncate_params(10, 30, 1e-15)
    c.set_compress_scheme(scheme, 1e-13, 20)
    c.grow_bmps_for_row(tn, 3)
    c.init_bten(tn, LEFT, 3)
    c.grow_full_bten(tn, RIGHT, 3, 2, True)
    z = c.trace(tn, (3, 0), HORIZONTAL)
    assert abs((math.log(z) - lz) / (L * L * beta)) < 1e-8
    c.grow_bmps_for_col(tn, 2)
    c.init_bten(tn, UP, 2)
    c.grow_full_bten(tn, DOWN, 2, 2, True)
    z = c.trace(tn, (0, 2), VERTICAL)
    assert abs((math.log(z) - lz) / (L * L * beta)) < 1e-8


def test_K7_all_pairs_correlators_vs_exact_diagonalisation():
    """K7: the reference's 4x4 D=8 Heisenberg fixture against its exact-diagonalisation table
    (tests/test_data/ed_reference/square_heisenberg_4x4_obc_ed.json, re-packed by tests/golden/make_k7_golden.py together with
    the oracle amplitude of EVERY S_z = 0 configuration): <Sz_i Sz_j> and <S+_i S-_j + S-_i S+_j>/2 of all 120 pairs by exact
    summation over the 12870 configurations, and the energy as the sum of the nearest-neighbour S_i.S_j. The residuals are
    the variational error of the D = 8 state (energy -9.18912 vs ED -9.18921, tests/slow_tests/test_boson_mc_peps_measure.cpp:
    31-76), not contraction error: the stored amplitudes are re-derived for a sample of configurations below."""
    import os
    from helpers import load_golden_tps
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "heis4x4_D8_all_amplitudes.npz"))
    states, amps = z["states"], z["amplitudes"]
    assert len(states) == 12870
    w = amps ** 2
    norm = w.sum()
    bits = (states[:, None] >> np.arange(16)[None, :]) & 1
    sz = bits - 0.5
    index = {int(s): k for k, s in enumerate(states)}
    worst_zz = worst_pm = 0.0
    energy = 0.0
    for (i, j), ezz, epm in zip(z["pairs"], z["ed_szsz"], z["ed_spsm"]):
        zz = float(np.sum(w * sz[:, i] * sz[:, j]) / norm)
        differ = bits[:, i] != bits[:, j]
        other = np.array([index[int(s)] for s in states[differ] ^ ((1 << int(i)) | (1 << int(j)))])
        pm = float(np.sum(amps[differ] * amps[other]) / norm) / 2.0
        worst_zz, worst_pm = max(worst_zz, abs(zz - ezz)), max(worst_pm, abs(pm - epm))
        if (j == i + 1 and i % 4 != 3) or j == i + 4:                          # nearest-neighbour bond (row-major sites)
            energy += zz + pm
    assert worst_zz < 1e-3 and worst_pm < 5e-4, (worst_zz, worst_pm)
    # the state's energy: the reference quotes its Monte Carlo estimate -9.18912; always above the ED ground state
    assert abs(energy - (-9.18912)) < 1e-4 and energy > float(z["ed_energy"])
    tps, _ = load_golden_tps("heis4x4_D8_double")
    rng = np.random.default_rng(0)
    for k in rng.choice(len(states), 6, replace=False):
        cfg = ((int(states[k]) >> np.arange(16)) & 1).reshape(4, 4)
        a = vmc.Walker(tps, cfg, (8, 16, 1e-15)).amplitude
        assert abs(a - amps[k]) <= 1e-11 * np.max(np.abs(amps))
