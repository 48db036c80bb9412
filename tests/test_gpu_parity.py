"""Parity of the CUDA path (through the C ABI) against the oracle. Run on the GPU box with ``-m gpu``.

Tolerances: FP64, <= 1e-10 relative on amplitudes / local energies / gradients (BASELINE.json north_star);
sampled configurations and acceptance counts bit-exact for equal RNG streams."""
import ctypes as C
import math

import numpy as np
import pytest

from parity_common import run_pipeline_parity, run_gradient_parity
from helpers import load_golden_tps, ising_tn, ising_exact_logZ

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from peps_b200 import _lib
    l = _lib.load()
    assert l.peps_backend_name() == b"cuda-sm_100a"
    return l


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


@pytest.mark.parametrize("spec,da,db", [
    ("apb,kea->kepb", (5, 3, 7), (4, 2, 5)),
    ("kepb,epfo->kofb", (6, 3, 2, 5), (3, 2, 4, 3)),
    ("apx,xyz->apyz", (64, 8, 64), (64, 8, 64)),
    ("apyz,yfop->zafo", (16, 8, 8, 16), (8, 8, 8, 8)),
    ("zafo,zfb->aob", (64, 64, 8, 8), (64, 8, 64)),
    ("xldb,brux->ldru", (9, 3, 3, 9), (9, 3, 3, 9)),
    ("rc,rt->ct", (513, 32), (513, 100)),
    ("ab,bc->ac", (1, 1), (1, 1)),
    ("ab,bc->ac", (70, 33), (33, 130)),
    # large-tile kernel (M > 64, N > 32): the four operand-layout variants, aligned pairs and true 8-byte gathers,
    # ragged edges in M, N and K
    ("ba,bc->ac", (34, 200), (34, 130)),
    ("ba,bc->ac", (33, 71), (33, 67)),
    ("ab,cb->ac", (200, 34), (130, 34)),
    ("ab,cb->ac", (71, 35), (67, 35)),
    ("ba,cb->ac", (40, 130), (66, 40)),
    ("ab,bc->ac", (129, 64), (64, 65)),
    ("kea,eaoj->koj", (96, 4, 10), (4, 10, 6, 10)),
    ("apb,kea->ekpb", (16, 4, 18), (80, 3, 16)),
    ("ekpb,epfo->kofb", (3, 80, 4, 18), (3, 4, 5, 7)),
    ("apx,xyz->apyz", (64, 8, 64), (64, 8, 64)),
])
def test_gett_contraction_vs_numpy(lib, spec, da, db):
    W = 3
    rng = np.random.default_rng(0)
    a = rng.standard_normal((W,) + da)
    b = rng.standard_normal((W,) + db)
    ins, out = spec.split("->")
    la, lb = ins.split(",")
    ref = np.einsum(f"w{la},w{lb}->w{out}", a, b)
    c = np.empty(ref.shape)
    dda, ddb = np.array(da, dtype=np.int32), np.array(db, dtype=np.int32)
    rc = lib.peps_test_einsum(0, W, spec.encode(), _ip(dda), len(da), _ip(ddb), len(db), _dp(a), _dp(b), _dp(c))
    assert rc == 0, lib.peps_last_error(None)
    assert np.max(np.abs(c - ref)) < 1e-12 * max(1.0, np.max(np.abs(ref)))


@pytest.mark.parametrize("m,n", [(8, 64), (64, 512), (512, 512), (4096, 512), (1000, 96), (37, 5), (5, 37), (2, 2),
                                 (1, 9), (600, 600)])
def test_caqr_r_factor(lib, m, n):
    W = 2
    rng = np.random.default_rng(1)
    a = rng.standard_normal((W, m, n))
    kk = min(m, n)
    r = np.empty((W, kk, n))
    assert lib.peps_test_qr_r(0, W, m, n, _dp(a), _dp(r)) == 0, lib.peps_last_error(None)
    for w in range(W):
        rw = r[w]
        assert np.max(np.abs(np.tril(rw[:, :kk], -1))) == 0.0
        # R^T R = A^T A  (R is unique up to row signs)
        ref = a[w].T @ a[w]
        assert np.max(np.abs(rw.T @ rw - ref)) < 1e-12 * np.max(np.abs(ref))
        rr = np.linalg.qr(a[w], mode="r")
        assert np.max(np.abs(np.abs(np.diag(rw)) - np.abs(np.diag(rr)))) < 1e-11 * np.max(np.abs(rr))


@pytest.mark.parametrize("nr,nc,dmin,dmax,terr", [(64, 64, 8, 8, 0.0), (512, 512, 64, 64, 0.0), (8, 512, 8, 8, 0.0),
                                                   (512, 64, 16, 16, 0.0), (40, 72, 4, 30, 1e-6), (3, 3, 2, 2, 0.0),
                                                   (100, 37, 10, 20, 1e-3)])
def test_jacobi_truncation_vs_lapack(lib, nr, nc, dmin, dmax, terr):
    from oracle.bmps import truncation_dim
    W = 2
    rng = np.random.default_rng(2)
    # known SVD with a geometric spectrum: the truth is V, not LAPACK's answer (whose own error is eps*s_1/gap)
    th = np.empty((W, nr, nc))
    vs, ss = [], []
    k = min(nr, nc)
    for w in range(W):
        u, _ = np.linalg.qr(rng.standard_normal((nr, k)))
        v, _ = np.linalg.qr(rng.standard_normal((nc, k)))
        s = 0.93 ** np.arange(k) * (1 + 0.02 * rng.random(k))
        s = np.sort(s)[::-1]
        th[w] = (u * s) @ v.T
        vs.append(v)
        ss.append(s)
    tcap = min(dmax, nr, nc)
    b = np.empty((W, tcap, nc))
    kept = np.empty(W, dtype=np.int32)
    sweeps = C.c_int32()
    rc = lib.peps_test_truncate(0, W, nr, nc, dmin, dmax, terr, _dp(th), _dp(b), _ip(kept), C.byref(sweeps))
    assert rc == 0, lib.peps_last_error(None)
    print("jacobi sweeps", sweeps.value)
    for w in range(W):
        t = truncation_dim(ss[w], dmin, dmax, terr)
        assert kept[w] == t
        bw = b[w][:t]
        assert np.max(np.abs(bw @ bw.T - np.eye(t))) < 1e-12
        # same projector onto the kept right singular subspace; conditioning ~ eps * s_1 / gap at the cut
        p_ref = vs[w][:, :t] @ vs[w][:, :t].T
        gap = ss[w][t - 1] - (ss[w][t] if t < k else 0.0)
        assert np.max(np.abs(bw.T @ bw - p_ref)) < 50 * 2.2e-16 * ss[w][0] / gap + 1e-13
        assert t == tcap or np.max(np.abs(b[w][t:])) == 0.0


@pytest.mark.parametrize("rows,cols,D,trunc,W", [
    (4, 4, 3, (6, 6, 0.0), 4),
    (3, 5, 2, (4, 4, 0.0), 3),
    (4, 4, 3, (2, 7, 1e-6), 3),
    (2, 2, 4, (1, 100, 0.0), 2),
    (4, 4, 4, (8, 8, 0.0), 8),       # BASELINE config #1 sizes (TFIM lattice, D=4, chi=8), Heisenberg bonds
])
def test_pipeline_parity_gpu(lib, rows, cols, D, trunc, W):
    rep = run_pipeline_parity(lib, rows, cols, D, W, trunc, nsweeps=2)
    print(rep)


def test_pipeline_parity_gpu_signed(lib):
    run_pipeline_parity(lib, 4, 4, 3, 2, (9, 9, 0.0), nsweeps=2, signed=True, seed=5)


@pytest.mark.parametrize("rows,cols,D,trunc,W", [(4, 4, 3, (6, 6, 0.0), 4), (3, 5, 2, (4, 4, 0.0), 3), (6, 6, 4, (16, 16, 0.0), 2)])
def test_j1j2_pipeline_parity_gpu(lib, rows, cols, D, trunc, W):
    """J1-J2 Heisenberg (BASELINE config #3's model): NNN terms through BTen2 + ReplaceNNNSiteTrace on the GPU."""
    run_pipeline_parity(lib, rows, cols, D, W, trunc, nsweeps=1, j2=0.5)


@pytest.mark.parametrize("rows,cols,D,trunc,W", [(4, 4, 4, (2, 8, 1e-14), 6), (3, 4, 3, (5, 5, 0.0), 3)])
def test_tfim_full_space_pipeline_parity_gpu(lib, rows, cols, D, trunc, W):
    """BASELINE config #1 (4x4, D=4, D_min=2, D_max=8, trunc_err=1e-14, h=0.5, full-space updater;
    examples/transverse_field_ising_vmc_optimize.cpp:69-90): chains bit-identical to the oracle's (Suwa-Todo with the
    reference's long double arithmetic), energies / holes / amplitudes to 1e-10."""
    rep = run_pipeline_parity(lib, rows, cols, D, W, trunc, nsweeps=2, tfim_h=0.5)
    print(rep)


def test_tfim_golden_2x2_energy_gpu(lib):
    """K4 (TFIM) on the GPU: exact summation over all 16 configurations of the reference's 2x2 simple-update fixture
    reproduces the energy its test asserts (test_exact_summation_evaluator.cpp:775)."""
    from helpers import load_golden_tps
    from peps_b200.api import BMPSTruncateParams, SplitIndexTPS, WalkerBatch, TransverseFieldIsingSquareOBC
    tps, z = load_golden_tps("tfim2x2_double_su")
    cfgs = np.array([[int(b) for b in format(k, "04b")] for k in range(16)]).reshape(16, 2, 2)
    b = WalkerBatch(2, 2, 2, tps[0][0][0].shape[2], 16, BMPSTruncateParams.SVD(1, 8, 1e-16), lib=lib)
    b.set_tps(SplitIndexTPS(tps))
    b.set_configs(cfgs)
    b.set_model(TransverseFieldIsingSquareOBC(1.0))
    b.init_walkers()
    e = b.energy_and_holes(True)
    wt = b.amplitudes() ** 2
    assert abs(np.sum(wt * e) / np.sum(wt) - float(z["exp_energy"])) < 1e-10


@pytest.mark.parametrize("j2", [0.0, 0.5])
def test_measure_bond_energies_parity_gpu(lib, j2):
    """EvaluateObservables (base/square_nnn_model_measurement_solver.h:33-214) on the GPU: every bond energy
    (horizontal, vertical, both diagonals) against the oracle's measurement restatement."""
    from oracle import vmc
    from peps_b200.api import BMPSTruncateParams, SplitIndexTPS, WalkerBatch, SquareSpinOneHalfJ1J2XXZModelOBC
    rows, cols, D, W = 4, 5, 3, 3
    tps = vmc.random_tps(rows, cols, 2, D, seed=31)
    cfgs = np.stack([vmc.shuffled_half_filled_config(rows, cols, 70 + w) for w in range(W)])
    b = WalkerBatch(rows, cols, 2, D, W, BMPSTruncateParams.SVD(6, 6, 0.0), lib=lib)
    b.set_tps(SplitIndexTPS(tps))
    b.set_configs(cfgs)
    b.set_model(SquareSpinOneHalfJ1J2XXZModelOBC(1.0, 0.8, j2, 0.7 * j2, 0.3))
    b.init_walkers()
    obs = b.measure()
    model = vmc.XXZModel(1.0, 0.8, 0.3, j2, 0.7 * j2)
    for w in range(W):
        ref = model.measure(tps, vmc.Walker(tps, cfgs[w], (6, 6, 0.0)))
        for k, v in ref.items():
            assert np.allclose(obs[k][w], v, rtol=1e-10, atol=1e-12), (k, w)


def test_three_site_updater_pipeline_parity_gpu(lib):
    """MCUpdateSquareTNN3SiteExchange (square_3site_updater.h:23-160) on the GPU: chains bit-identical to the oracle's."""
    rep = run_pipeline_parity(lib, 4, 5, 3, 3, (6, 6, 0.0), nsweeps=2, three_site=True)
    print(rep)


@pytest.mark.parametrize("orient", [0, 1])
def test_three_site_trace_parity_gpu(lib, orient):
    """ReplaceTNNSiteTrace (bmps/impl/bmps_contractor_trace.h:326-420) on the GPU against amplitudes of the oracle
    evaluated from scratch on the modified configurations."""
    from oracle import vmc
    from peps_b200.api import BMPSTruncateParams, SplitIndexTPS, WalkerBatch
    rows, cols, D, W = 4, 5, 3, 3
    tps = vmc.random_tps(rows, cols, 2, D, seed=8)
    cfgs = np.stack([vmc.shuffled_half_filled_config(rows, cols, 20 + w) for w in range(W)])
    b = WalkerBatch(rows, cols, 2, D, W, BMPSTruncateParams.SVD(1, 200, 0.0), lib=lib)
    b.set_tps(SplitIndexTPS(tps))
    b.set_configs(cfgs)
    b.init_walkers()
    r0, c0 = (2, 1) if orient == 0 else (1, 3)
    sites = [(r0, c0 + k) if orient == 0 else (r0 + k, c0) for k in range(3)]
    new = np.array([[1, 0, 1], [0, 0, 1], [1, 1, 0]], dtype=np.int32)
    psi = b.probe_tnn_trace(r0, c0, orient, new)
    for w in range(W):
        cf = cfgs[w].copy()
        for k, st in enumerate(sites):
            cf[st] = new[w, k]
        ref = vmc.Walker(tps, cf, (1, 200, 0.0)).amplitude
        assert abs(psi[w] / ref - 1) < 1e-10


def test_gradient_parity_gpu(lib):
    run_gradient_parity(lib, 3, 4, 2, 3, (4, 4, 0.0), nsamples=4)


def test_golden_4x4_D8_fixture_amplitudes(lib):
    """K6 fixture (the one physical, non-synthetic state of the suite): amplitudes / energies of ALL 32 stored
    configurations at the reference's truncation (8, 16, 1e-15), to the north-star tolerance 1e-10."""
    from oracle import vmc
    from peps_b200.api import BMPSTruncateParams, SplitIndexTPS, WalkerBatch
    tps, z = load_golden_tps("heis4x4_D8_double")
    cfgs = z["configs"]
    b = WalkerBatch(4, 4, 2, 8, len(cfgs), BMPSTruncateParams.SVD(8, 16, 1e-15), lib=lib)
    b.set_tps(SplitIndexTPS(tps))
    b.set_configs(cfgs)
    b.init_walkers()
    amp = b.amplitudes()
    e = b.energy_and_holes(False)
    for w in range(len(cfgs)):
        wk = vmc.Walker(tps, cfgs[w], (8, 16, 1e-15))
        assert abs(amp[w] / wk.amplitude - 1) < 1e-10, (w, amp[w], wk.amplitude)
        ee, _, _ = vmc.XXZModel().energy_and_holes(tps, wk, False)
        assert abs(e[w] - ee) < 1e-10 * max(1, abs(ee)), (w, e[w], ee)


def test_K1_ising_partition_function_on_gpu(lib):
    """The exact-partition-function KAT through the CUDA path: a 'TPS' whose two physical slices are the Ising
    site tensor (slice 0) and zero (slice 1), configuration all 0."""
    from peps_b200.api import BMPSTruncateParams, SplitIndexTPS, WalkerBatch
    L = 8
    beta = math.log(1 + math.sqrt(2.0)) / 2.0
    tn = ising_tn(L, beta)
    tps = [[[tn[r][c], np.zeros_like(tn[r][c])] for c in range(L)] for r in range(L)]
    b = WalkerBatch(L, L, 2, 2, 1, BMPSTruncateParams.SVD(10, 30, 1e-15), lib=lib)
    b.set_tps(SplitIndexTPS(tps))
    b.set_configs(np.zeros((1, L, L), dtype=np.int32))
    b.init_walkers()
    lz = ising_exact_logZ(L, beta)
    for row in (0, 2, 5):
        z = b.probe_trace_row(row)[0]
        assert abs((math.log(z) - lz) / (L * L * beta)) < 1e-8


def test_full_size_properties(lib):
    """At a BASELINE-sized bond dimension (D=8, chi=64 on a 4x4 lattice so the oracle is not needed): closure --
    every row/column trace of the same configuration agrees, and PunchHole . site == amplitude."""
    from oracle import vmc
    from peps_b200.api import BMPSTruncateParams, SplitIndexTPS, WalkerBatch
    rows = cols = 4
    D, chi, W = 8, 64, 4
    tps = vmc.random_tps(rows, cols, 2, D, seed=11)
    cfgs = np.stack([vmc.shuffled_half_filled_config(rows, cols, 30 + w) for w in range(W)])
    b = WalkerBatch(rows, cols, 2, D, W, BMPSTruncateParams.SVD(chi, chi, 0.0), lib=lib)
    b.set_tps(SplitIndexTPS(tps))
    b.set_configs(cfgs)
    b.init_walkers()
    e, psi = b.energy_and_holes(True, True)
    assert np.max(np.abs(psi / psi[0] - 1)) < 1e-10          # D^2 = chi: no truncation on 4x4, closures agree
    holes = b.holes()
    flat_tn = np.stack([np.concatenate([tps[r][c][cfgs[w, r, c]].ravel() for r in range(rows) for c in range(cols)])
                        for w in range(W)])
    sizes = [tps[r][c][0].size for r in range(rows) for c in range(cols)]
    pos = 0
    for s, sz in enumerate(sizes):
        dot = np.sum(holes[:, pos:pos + sz] * flat_tn[:, pos:pos + sz], axis=1)
        assert np.max(np.abs(dot / psi[s // cols] - 1)) < 1e-10
        pos += sz


@pytest.mark.parametrize("L,D,chi,signed,tol", [(10, 8, 64, False, 1e-10), (8, 6, 36, True, 1e-9)])
def test_full_size_amplitude_parity(lib, L, D, chi, signed, tol):
    """BASELINE's headline size against the oracle itself (one CPU amplitude = 9 row absorptions, ~20 s each):
    amplitudes of two walkers at 10x10, D=8, chi=64 must agree to 1e-10 relative (observed 7e-15).
    The signed [-1,1) TPS is the cancellation stress variant: its boundary spectra are flat, the chi-truncated
    contraction is not a meaningful approximation of such a state (row closures of ONE configuration differ by
    factors of ~10 in the oracle too), and the conditioning of the kept subspaces degrades the agreement between any
    two correct implementations: 1e-12 at 8x8, D=6, chi=36, and no digits at 10x10, D=8, chi=64."""
    from oracle import vmc
    from peps_b200.api import BMPSTruncateParams, SplitIndexTPS, WalkerBatch
    W = 2
    tps = vmc.random_tps(L, L, 2, D, seed=20260101, signed=signed)
    cfgs = np.stack([vmc.shuffled_half_filled_config(L, L, 1000 + w) for w in range(W)])
    b = WalkerBatch(L, L, 2, D, W, BMPSTruncateParams.SVD(chi, chi, 0.0), lib=lib)
    b.set_tps(SplitIndexTPS(tps))
    b.set_configs(cfgs)
    b.init_walkers()
    amp = b.amplitudes()
    worst = 0.0
    for w in range(W):
        ref = vmc.Walker(tps, cfgs[w], (chi, chi, 0.0)).amplitude
        worst = max(worst, abs(amp[w] / ref - 1))
    print(f"{L}x{L} D{D} chi{chi} signed={signed} amplitude rel err", worst, "chain rows kept", b.stat(13), "of", b.stat(12))
    assert worst < tol
    if not signed:
        # the rank-revealing forward chain really shortened the chain here (and the amplitudes above still agree with
        # the oracle, which keeps every row)
        assert 0 < b.stat(13) < b.stat(12)


def test_config2_8x8_D6_chi36_sample_vs_oracle(lib):
    """BASELINE config #2 sizes (Heisenberg 8x8, D=6, chi=36): one sweep + E_loc + holes of two walkers vs the oracle."""
    rep = run_pipeline_parity(lib, 8, 8, 6, 2, (36, 36, 0.0), nsweeps=1, seed=20260101)
    print(rep)


def test_config5_14x14_D10_chi100_smoke(lib):
    """BASELINE config #5 sizes exercise the 16-wide CAQR panels and the 8-row Jacobi blocks (D*chi = 1000): the
    amplitude is finite, identical walkers agree bit for bit, and the row closures agree to truncation accuracy."""
    from oracle import vmc
    from peps_b200.api import BMPSTruncateParams, SplitIndexTPS, WalkerBatch
    L, D, chi, W = 14, 10, 100, 2
    tps = vmc.random_tps(L, L, 2, D, seed=5)
    cfg = vmc.shuffled_half_filled_config(L, L, 77)
    b = WalkerBatch(L, L, 2, D, W, BMPSTruncateParams.SVD(chi, chi, 0.0), lib=lib)
    b.set_tps(SplitIndexTPS(tps))
    b.set_configs(np.stack([cfg, cfg]))
    b.init_walkers()
    amp = b.amplitudes()
    assert np.all(np.isfinite(amp)) and amp[0] != 0.0 and amp[0] == amp[1]
    psi_mid = b.probe_trace_row(7)
    assert abs(psi_mid[0] / amp[0] - 1) < 1e-6


def test_sr_matvec_and_natural_gradient_gpu(lib):
    """O* sample store in HBM + sr_dots / sr_accumulate kernels + CG vs the oracle's dense S matrix."""
    from test_sr import sr_scenario
    sr_scenario(lib)


def _full_sample_case(lib, z, idx, W_extra=0):
    """Runs case `idx` of tests/golden/full_sample_10x10_D8_chi64.npz through the CUDA path: init, one MC sweep, E_loc +
    holes, O* accumulation. Returns everything the golden holds."""
    from oracle import vmc
    from peps_b200.api import BMPSTruncateParams, SplitIndexTPS, WalkerBatch, SquareSpinOneHalfJ1J2XXZModelOBC
    L, D, chi = int(z["L"]), int(z["D"]), int(z["chi"])
    tps = vmc.random_tps(L, L, 2, D, seed=int(z["tps_seed"]))
    names = [str(z[f"c{i}_name"]) for i in range(int(z["ncases"]))]
    same = [i for i in range(len(names)) if names[i] == names[idx]]
    W = len(same)
    b = WalkerBatch(L, L, 2, D, W, BMPSTruncateParams.SVD(chi, chi, 0.0), lib=lib)
    b.set_tps(SplitIndexTPS(tps))
    j2 = float(z[f"c{idx}_j2"])
    if j2 != 0.0:
        b.set_model(SquareSpinOneHalfJ1J2XXZModelOBC(1.0, 1.0, j2, j2, 0.0))
    b.set_configs(np.stack([z[f"c{i}_cfg0"] for i in same]))
    b.seed_rng(np.array([7 + int(z[f"c{i}_w"]) for i in same], dtype=np.uint32))
    b.init_walkers()
    amp0 = b.amplitudes()
    b.zero_accumulators()
    acc = b.sweep(1)
    amp1 = b.amplitudes()
    cfg1 = b.get_configs()
    e, psi = b.energy_and_holes(True, True)
    holes = b.holes()
    b.accumulate_ostar()
    osum, eosum = b.accumulators()
    return dict(tps=tps, same=same, amp0=amp0, amp1=amp1, acc=acc, cfg1=cfg1, e=e, psi=psi, holes=holes, osum=osum,
                eosum=eosum, batch=b)


@pytest.mark.parametrize("model", ["nn", "j1j2"])
def test_full_sample_at_headline_config_vs_oracle_golden(lib, model):
    """BASELINE's headline configuration (10x10, D=8, chi=64; NN Heisenberg with two walkers, J1-J2 with one): ONE FULL
    SAMPLE -- MC sweep + CalEnergyAndHoles<true> + O* accumulation -- against the oracle's result stored by
    tests/golden/make_full_sample_golden.py. Swept configurations and acceptance counts bit-identical; amplitudes,
    E_loc, the psi list, the holes (every 13th element, per-site <hole, site> and norms, 8 random projections) and the
    accumulators sum O*, sum E_loc O* (rebuilt from the oracle's holes) to <= 1e-10 relative."""
    import os
    from helpers import GOLDEN
    z = np.load(os.path.join(GOLDEN, "full_sample_10x10_D8_chi64.npz"))
    names = [str(z[f"c{i}_name"]) for i in range(int(z["ncases"]))]
    idx = names.index(model)
    r = _full_sample_case(lib, z, idx)
    tol = 1e-10
    L = int(z["L"])
    stride = int(z["stride"])
    rng = np.random.default_rng(4242)
    nh = r["holes"].shape[1]
    probes = [rng.standard_normal(nh) for _ in range(8)]
    worst = {}
    for k, i in enumerate(r["same"]):
        assert np.array_equal(r["cfg1"][k], z[f"c{i}_cfg1"]), f"walker {k}: swept configuration differs from the oracle's"
        assert r["acc"][k] == float(z[f"c{i}_accept"])
        rel = lambda a, b_: float(np.max(np.abs(np.asarray(a) - np.asarray(b_))) / np.max(np.abs(np.asarray(b_))))
        checks = {"amp0": abs(r["amp0"][k] / float(z[f"c{i}_amp0"]) - 1), "amp1": abs(r["amp1"][k] / float(z[f"c{i}_amp1"]) - 1),
                  "eloc": abs(r["e"][k] - float(z[f"c{i}_eloc"])) / max(1.0, abs(float(z[f"c{i}_eloc"]))),
                  "psi": float(np.max(np.abs(r["psi"][:, k] / z[f"c{i}_psi"] - 1))),
                  "hole_sub": rel(r["holes"][k][::stride], z[f"c{i}_hole_sub"]),
                  "hole_proj": rel([float(np.dot(p, r["holes"][k])) for p in probes], z[f"c{i}_hole_proj"])}
        # per-site <hole, site tensor> = psi and per-site norms
        tps, cfg = r["tps"], r["cfg1"][k]
        pos, dots, norms = 0, [], []
        for rr in range(L):
            for cc in range(L):
                t = tps[rr][cc][int(cfg[rr, cc])].ravel()
                h = r["holes"][k][pos:pos + t.size]
                dots.append(float(np.dot(h, t))); norms.append(float(np.linalg.norm(h)))
                pos += t.size
        checks["hole_dots"] = float(np.max(np.abs(np.array(dots) / z[f"c{i}_hole_dots"] - 1)))
        checks["hole_norms"] = float(np.max(np.abs(np.array(norms) / z[f"c{i}_hole_norms"] - 1)))
        for name, v in checks.items():
            worst[name] = max(worst.get(name, 0.0), v)
    # accumulators: sum_w O*_w and sum_w E_w O*_w with O* = hole / amplitude at the sampled slot; the oracle side is
    # rebuilt from its stored strided hole elements, the CUDA side read back from the device accumulators
    b = r["batch"]
    lay_off = []
    pos = 0
    tps = r["tps"]
    for rr in range(L):
        for cc in range(L):
            lay_off.append(pos)
            pos += 2 * tps[rr][cc][0].size
    ref_o, got_o, ref_eo, got_eo = [], [], [], []
    for k, i in enumerate(r["same"]):
        cfg = r["cfg1"][k]
        inv, el = 1.0 / float(z[f"c{i}_amp1"]), float(z[f"c{i}_eloc"])
        sub = z[f"c{i}_hole_sub"]
        hpos, sidx = 0, 0
        o_w = np.zeros(pos)
        mask = np.zeros(pos, dtype=bool)
        for rr in range(L):
            for cc in range(L):
                sz = tps[rr][cc][0].size
                s = int(cfg[rr, cc])
                first = (-hpos) % stride
                sel = np.arange(first, sz, stride)
                gl = (hpos + sel) // stride
                o_w[lay_off[sidx] + s * sz + sel] = inv * sub[gl]
                mask[lay_off[sidx] + s * sz + sel] = True
                hpos += sz; sidx += 1
        ref_o.append(o_w); ref_eo.append(el * o_w)
        got_o.append(mask)
    ref_osum, ref_eosum = sum(ref_o), sum(ref_eo)
    anymask = np.any(np.stack(got_o), axis=0)
    worst["osum"] = float(np.max(np.abs(r["osum"][anymask] - ref_osum[anymask])) / np.max(np.abs(ref_osum)))
    worst["eosum"] = float(np.max(np.abs(r["eosum"][anymask] - ref_eosum[anymask])) / np.max(np.abs(ref_eosum)))
    print(model, "full sample vs oracle golden:", {k: f"{v:.2e}" for k, v in worst.items()},
          "chain rows kept", b.stat(13), "of", b.stat(12))
    for name, v in worst.items():
        assert v < tol, (name, worst)
    assert 0 < b.stat(13) < b.stat(12)          # the rank-revealing chain was active in this run
    b.close()


def test_walker_alone_vs_inside_a_batch(lib):
    """A walker's numbers inside a batch: kept counts of the rank-revealing chain are per walker, but buffers are sized by
    the batch maximum (rounded), which changes the CAQR tree and so the rounding. Guaranteed (and asserted): the
    sampled configuration and acceptance count are identical; amplitude / E_loc agree to 1e-11 relative (observed
    ~1e-14: the differences are those of two backward-stable factorizations of the same matrix)."""
    from oracle import vmc
    from peps_b200.api import BMPSTruncateParams, SplitIndexTPS, WalkerBatch
    L, D, chi, W = 8, 6, 36, 37
    tps = vmc.random_tps(L, L, 2, D, seed=20260101)
    cfgs = np.stack([vmc.shuffled_half_filled_config(L, L, 1000 + w) for w in range(W)])
    seeds = np.arange(7, 7 + W, dtype=np.uint32)
    out = []
    for sel in (slice(0, 1), slice(0, W)):
        n = len(range(W)[sel])
        b = WalkerBatch(L, L, 2, D, n, BMPSTruncateParams.SVD(chi, chi, 0.0), lib=lib)
        b.set_tps(SplitIndexTPS(tps)); b.set_configs(cfgs[sel]); b.seed_rng(seeds[sel]); b.init_walkers()
        acc = b.sweep(1)
        e = b.energy_and_holes(True)
        out.append((b.amplitudes()[0], acc[0], e[0], b.get_configs()[0]))
        b.close()
    (a1, c1, e1, g1), (a2, c2, e2, g2) = out
    assert np.array_equal(g1, g2) and c1 == c2
    print("alone vs batch of 37: amplitude", abs(a1 / a2 - 1), "eloc", abs(e1 - e2) / max(1, abs(e1)))
    assert abs(a1 / a2 - 1) < 1e-11 and abs(e1 - e2) < 1e-11 * max(1, abs(e1))


def test_cpp_wrapper_on_cuda_library(lib):
    """include/peps_b200.hpp compiled with g++ and linked against libpeps_b200.so itself (not the host simulation):
    Evaluate() through the C++ wrapper equals the Python mirror on the same CUDA library."""
    import os
    from test_cpp_host import run_cpp_wrapper_case, ROOT
    run_cpp_wrapper_case(lib, os.path.join(ROOT, "peps_b200"), "libpeps_b200.so",
                         extra_link=["-Wl,-rpath,/usr/local/cuda/lib64", "-L/usr/local/cuda/lib64", "-lcudart"])


def test_context_is_usable_from_another_thread(lib):
    """a context created in one host thread and driven from another (and two contexts interleaved in one
    thread) must run on the context's own device / stream: every C-ABI entry binds them."""
    import threading
    from oracle import vmc
    from peps_b200.api import BMPSTruncateParams, SplitIndexTPS, WalkerBatch
    tps = vmc.random_tps(4, 4, 2, 3, seed=1)
    cfgs = np.stack([vmc.shuffled_half_filled_config(4, 4, 10 + w) for w in range(3)])

    def make():
        b = WalkerBatch(4, 4, 2, 3, 3, BMPSTruncateParams.SVD(6, 6, 0.0), lib=lib)
        b.set_tps(SplitIndexTPS(tps)); b.set_configs(cfgs); b.seed_rng(np.arange(3) + 5)
        return b
    b1, b2 = make(), make()
    b1.init_walkers()
    ref = b1.sample(1)[0]
    res = {}

    def worker():
        b2.init_walkers()
        res["e"] = b2.sample(1)[0]
    t = threading.Thread(target=worker); t.start(); t.join()
    assert np.array_equal(res["e"], ref)
    e1 = b1.sample(1)[0]; e2 = b2.sample(1)[0]      # interleaved in the main thread
    assert np.array_equal(e1, e2)
    with pytest.raises(Exception):
        b1.set_configs(np.full((3, 4, 4), 2))       # entry outside [0, phys)


NCCL_WORKER = r'''
import os, sys, pickle
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import torch, torch.distributed as dist
from oracle import vmc
from peps_b200 import sr
from peps_b200.api import (BMPSTruncateParams, SplitIndexTPS, MCEnergyGradEvaluator, MonteCarloParams,
                           SquareSpinOneHalfXXZModelOBC, MCUpdateSquareNNExchange)
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
from parity_common import complex_tps
tps = SplitIndexTPS(complex_tps(4, 4, 3, 4) if {cx!r} else vmc.random_tps(4, 4, 2, 3, seed=4))
W = 3
cfgs = np.stack([vmc.shuffled_half_filled_config(4, 4, 50 + rank * W + w) for w in range(W)])
mc = MonteCarloParams(num_samples=4 * W * world, num_warmup_sweeps=0, sweeps_between_samples=1, is_warmed_up=True)
ev = MCEnergyGradEvaluator(mc, BMPSTruncateParams.SVD(6, 6, 0.0), tps, SquareSpinOneHalfXXZModelOBC(1, 1, 0),
                           MCUpdateSquareNNExchange(9), walkers=W, configs=cfgs, device=rank, dist=dist, rank=rank, world_size=world)
res = ev.Evaluate(collect_sr_buffers=True)
nat, iters, resid = ev.CalculateNaturalGradient(res, 1e-3, sr.ConjugateGradientParams(max_iter=300, relative_tolerance=1e-10))
if rank == 0:
    pickle.dump(dict(energy=res.energy, grad=res.gradient.pack(), es=res.energy_samples, nat=nat.pack(), iters=iters), open({out!r}, "wb"))
dist.destroy_process_group()
'''


@pytest.mark.parametrize("cx", [False, True])
def test_two_rank_nccl_evaluator_and_sr(lib, cx):
    """MCEnergyGradEvaluator + SR natural gradient over two GPUs (mc_energy_grad_evaluator.h:205-310): accumulators
    all-reduced by NCCL on peps_ostar_sum_device() pointers, the CG matvec output all-reduced on its device pointer.
    Equals one GPU holding all six walkers. cx: a complex state (both planes of an accumulator / CG vector in one
    all-reduce). Skipped below two GPUs."""
    import os, pickle, subprocess, sys, tempfile
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from oracle import vmc
    from peps_b200 import sr
    from peps_b200.api import (BMPSTruncateParams, SplitIndexTPS, MCEnergyGradEvaluator, MonteCarloParams,
                               SquareSpinOneHalfXXZModelOBC, MCUpdateSquareNNExchange)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with tempfile.TemporaryDirectory() as td:
        out, script = os.path.join(td, "res.pkl"), os.path.join(td, "worker.py")
        open(script, "w").write(NCCL_WORKER.format(root=root, out=out, cx=cx))
        env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29712" if cx else "29711", WORLD_SIZE="2")
        procs = [subprocess.Popen([sys.executable, script], env=dict(env, RANK=str(r))) for r in range(2)]
        for p in procs:
            assert p.wait(timeout=600) == 0
        two = pickle.load(open(out, "rb"))
    from parity_common import complex_tps
    tps = SplitIndexTPS(complex_tps(4, 4, 3, 4) if cx else vmc.random_tps(4, 4, 2, 3, seed=4))
    cfgs = np.stack([vmc.shuffled_half_filled_config(4, 4, 50 + w) for w in range(6)])
    mc = MonteCarloParams(num_samples=24, num_warmup_sweeps=0, sweeps_between_samples=1, is_warmed_up=True)
    ev = MCEnergyGradEvaluator(mc, BMPSTruncateParams.SVD(6, 6, 0.0), tps, SquareSpinOneHalfXXZModelOBC(1, 1, 0),
                               MCUpdateSquareNNExchange(9), walkers=6, configs=cfgs, lib=lib)
    one = ev.Evaluate(collect_sr_buffers=True)
    nat1, it1, _ = ev.CalculateNaturalGradient(one, 1e-3, sr.ConjugateGradientParams(max_iter=300, relative_tolerance=1e-10))
    assert np.allclose(two["es"], one.energy_samples, rtol=1e-11, atol=1e-12)
    assert abs(two["energy"] - one.energy) < 1e-11 * max(1, abs(one.energy))
    g1 = one.gradient.pack()
    assert np.max(np.abs(two["grad"] - g1)) < 1e-10 * np.max(np.abs(g1))
    assert np.max(np.abs(two["nat"] - nat1.pack())) < 1e-7 * np.max(np.abs(nat1.pack()))


def test_plaquette_traces_gpu(lib):
    """ReplaceNNNSiteTrace with VERTICAL MPS orientation and ReplaceSqrt5DistTwoSiteTrace (trace.h:282-324, 426-536) on
    the GPU: equal to the oracle amplitude of the exchanged configuration."""
    from parity_common import run_plaquette_trace_parity
    print("plaquette traces worst rel err", run_plaquette_trace_parity(lib))


def test_table_model_gpu(lib):
    """Seam B2 as data on the GPU: XXZ / J1-J2 and TFIM tables reproduce the built-in solvers; a spin-1 model with no
    engine branch matches the oracle's generic restatement."""
    from parity_common import run_table_model_parity
    print("table model worst rel err", run_table_model_parity(lib))


def test_structure_factor_gpu(lib):
    """MeasureStructureFactor (all-pairs S+S- by excited-state propagation, structure_factor_measurement_mixin.h:89-228)
    on the GPU against the oracle, with and without truncation of the propagated boundary."""
    from parity_common import run_structure_factor_parity
    print("structure factor worst rel err", run_structure_factor_parity(lib), run_structure_factor_parity(lib, 4, 4, 3, 2, chi=5))


@pytest.mark.parametrize("scheme", [1, 2])
def test_variational_compression_gpu(lib, scheme):
    """VARIATION2Site / VARIATION1Site boundary compression (bmps_impl.h:864-1172) on the GPU vs the oracle."""
    from parity_common import run_variational_parity
    print("variational scheme", scheme, "worst rel err", run_variational_parity(lib, scheme))


@pytest.mark.parametrize("complex_", [False, True])
def test_tfim_measurement_parity_gpu(lib, complex_):
    """sigma_x per site (peps_measure_site_term), energy, spin_z, SzSz_row of the TFIM measurement solver on the CUDA path."""
    from parity_common import run_tfim_measure_parity
    run_tfim_measure_parity(lib, complex_)
