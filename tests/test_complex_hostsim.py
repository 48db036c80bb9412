"""Complex (QLTEN_Complex) states through the C ABI on the test-only host simulation of the device ops, against the
oracle (pinned on the reference's complex goldens by tests/test_oracle_kat.py)."""
import numpy as np
import pytest

import hostsim_lib
from parity_common import run_complex_pipeline_parity


@pytest.fixture(scope="module")
def lib():
    return hostsim_lib.load()


@pytest.mark.parametrize("rows,cols,D,trunc", [
    (2, 2, 3, (1, 100, 0.0)),
    (3, 3, 2, (4, 4, 0.0)),
    (4, 4, 3, (6, 6, 0.0)),
    (3, 4, 2, (2, 4, 1e-8)),
])
def test_complex_pipeline_parity_hostsim(lib, rows, cols, D, trunc):
    run_complex_pipeline_parity(lib, rows, cols, D, 2, trunc, nsweeps=2)


def test_complex_j1j2_pipeline_parity_hostsim(lib):
    """J1-J2 (BTen2 / NNN traces) on a complex state."""
    run_complex_pipeline_parity(lib, 3, 4, 2, 2, (4, 4, 0.0), nsweeps=1, j2=0.5)


def test_k5_complex_golden_through_abi(lib):
    from parity_common import run_complex_k5_golden
    run_complex_k5_golden(lib)


def test_complex_mode_guards(lib):
    from peps_b200.api import BMPSTruncateParams, WalkerBatch, PepsError, TableModel, FermionSplitIndexTPS
    b = WalkerBatch(3, 3, 2, 2, 1, BMPSTruncateParams.SVD(4, 4, 0.0), lib=lib)
    b.set_complex()
    b.close()
    b = WalkerBatch(3, 3, 2, 2, 1, BMPSTruncateParams.SVD(4, 4, 0.0), lib=lib)
    b.set_tps(np.zeros(b.tps_size))
    with pytest.raises(PepsError):                       # the switch comes before the first state
        b.set_complex()
    b.close()


def test_complex_tall_chain_matrices_are_reduced_by_chunks(lib, monkeypatch):
    """The real embeddings of the complex chain matrices exceed the two-stage CAQR height at the headline size; they are
    reduced by row chunks first. Forced here at a small size (PEPS_QR_MAX_ROWS=64): results unchanged."""
    monkeypatch.setenv("PEPS_QR_MAX_ROWS", "64")
    run_complex_pipeline_parity(lib, 4, 4, 3, 2, (6, 6, 0.0), nsweeps=1)


def test_complex_evaluator_matches_oracle_chain(lib):
    """MCEnergyGradEvaluator on a complex state (seam B1): energy and gradient = sum conj(E_loc) O* / N - conj(E) sum O* / N
    from the oracle chain's samples."""
    from parity_common import complex_tps, flat_holes
    from oracle import vmc
    from peps_b200.api import (BMPSTruncateParams, SplitIndexTPS, MCEnergyGradEvaluator, MonteCarloParams, Configuration,
                               SquareSpinOneHalfXXZModelOBC, MCUpdateSquareNNExchange, combine_energy_bins)
    rows, cols, D, W, n, trunc = 3, 3, 2, 2, 3, (4, 4, 0.0)
    tps = complex_tps(rows, cols, D, 9)
    cfgs = np.stack([vmc.shuffled_half_filled_config(rows, cols, 70 + w) for w in range(W)])
    ev = MCEnergyGradEvaluator(MonteCarloParams(n * W, 0, 1, Configuration(cfgs[0]), True), BMPSTruncateParams.SVD(*trunc),
                               SplitIndexTPS(tps), SquareSpinOneHalfXXZModelOBC(1, 1, 0), MCUpdateSquareNNExchange(seed=55), W,
                               configs=cfgs, lib=lib)
    res = ev.Evaluate()
    model = vmc.XXZModel()
    es = np.zeros((W, n), dtype=complex)
    osum = np.zeros(ev.batch.tps_size, dtype=complex)
    eosum = np.zeros(ev.batch.tps_size, dtype=complex)
    for w in range(W):
        wk = vmc.Walker(tps, cfgs[w], trunc)
        up = vmc.NNExchangeUpdater(55 + w)
        for k in range(n):
            up.sweep(tps, wk)
            e, holes, _ = model.energy_and_holes(tps, wk, True)
            es[w, k] = e
            ost = flat_holes(holes, rows, cols) * np.conj(1.0 / wk.amplitude)
            off = hoff = 0
            for r in range(rows):
                for c in range(cols):
                    sz = tps[r][c][0].size
                    s_ = int(wk.config[r, c])
                    osum[off + s_ * sz: off + (s_ + 1) * sz] += ost[hoff:hoff + sz]
                    eosum[off + s_ * sz: off + (s_ + 1) * sz] += np.conj(e) * ost[hoff:hoff + sz]
                    off += 2 * sz
                    hoff += sz
    energy = complex(combine_energy_bins(es.real)[0], combine_energy_bins(es.imag)[0])
    grad = (eosum - np.conj(energy) * osum) / (n * W)
    assert abs(res.energy - energy) < 1e-10
    got = np.concatenate([x.ravel() for row in res.gradient.t for site in row for x in site])
    assert np.max(np.abs(got - grad)) <= 1e-9 * max(1.0, np.max(np.abs(grad)))


def test_complex_tfim_full_space_pipeline_parity_hostsim(lib):
    """Transverse-field Ising + full-space (Suwa-Todo) updater on a complex state (BASELINE config #1's flow in QLTEN_Complex)."""
    run_complex_pipeline_parity(lib, 3, 3, 2, 2, (4, 4, 0.0), nsweeps=2, tfim_h=0.7)


def test_complex_three_site_updater_pipeline_parity_hostsim(lib):
    run_complex_pipeline_parity(lib, 3, 4, 2, 2, (4, 4, 0.0), nsweeps=2, three_site=True)


@pytest.mark.parametrize("table,j2", [("xxz", 0.0), ("xxz", 0.4), ("spin1", 0.0)])
def test_complex_table_model_pipeline_parity_hostsim(lib, table, j2):
    """Seam B2 as data on complex states: XXZ / J1-J2 as tables against the oracle's XXZ solver, and the phys = 3 spin-1
    model (NN + NNN + on-site off-diagonal terms) with the full-space updater against the oracle's generic traversal."""
    run_complex_pipeline_parity(lib, 3, 3, 2, 2, (4, 4, 0.0), nsweeps=1, j2=j2, table=table)


@pytest.mark.parametrize("j2", [0.0, 0.5])
def test_complex_measure_parity_hostsim(lib, j2):
    from parity_common import run_complex_measure_parity
    run_complex_measure_parity(lib, j2)


def test_complex_structure_factor_hostsim(lib):
    from parity_common import run_structure_factor_parity
    run_structure_factor_parity(lib, complex_=True)


@pytest.mark.parametrize("scheme", [1, 2])
def test_complex_variational_compression_hostsim(lib, scheme):
    """VARIATION2Site / VARIATION1Site on a complex state: the environments contract the conjugate of the result tensors
    (res_dag, bmps_impl.h:885-947); amplitudes against the oracle restatement with the same number of sweeps."""
    from parity_common import run_variational_parity
    run_variational_parity(lib, scheme, complex_=True)
