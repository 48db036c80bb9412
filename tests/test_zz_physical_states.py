"""Physical (non-synthetic) states of the reference through the product path. (i) K7 golden amplitudes
(tests/golden/heis4x4_D8_all_amplitudes.npz: the oracle amplitude of every S_z = 0 configuration of the reference's 4x4 D=8
Heisenberg fixture, whose correlators match the reference's ED table in tests/test_oracle_kat.py): a random sample of
configurations through the C ABI at the reference's truncation (8, 16, 1e-15). (ii) The t-J iPEPS unit cell tiled to an OBC
lattice in fermion mode. The GPU twins run last in the suite."""
import os

import numpy as np
import pytest

import hostsim_lib
from helpers import load_golden_tps
from peps_b200.api import BMPSTruncateParams, SplitIndexTPS, WalkerBatch


def run_k7_amplitudes(lib, nsample, seed=3):
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "heis4x4_D8_all_amplitudes.npz"))
    states, amps = z["states"], z["amplitudes"]
    tps, _ = load_golden_tps("heis4x4_D8_double")
    pick = np.random.default_rng(seed).choice(len(states), nsample, replace=False)
    cfgs = ((states[pick][:, None] >> np.arange(16)[None, :]) & 1).reshape(nsample, 4, 4).astype(np.int32)
    b = WalkerBatch(4, 4, 2, 8, nsample, BMPSTruncateParams.SVD(8, 16, 1e-15), lib=lib)
    b.set_tps(SplitIndexTPS(tps))
    b.set_configs(cfgs)
    b.init_walkers()
    got = b.amplitudes()
    b.close()
    scale = np.max(np.abs(amps))
    worst = float(np.max(np.abs(got - amps[pick])) / scale)
    assert worst < 1e-10, worst
    return worst


def test_k7_fixture_amplitudes_hostsim():
    run_k7_amplitudes(hostsim_lib.load(), 48)


@pytest.mark.gpu
def test_k7_fixture_amplitudes_gpu():
    from peps_b200 import _lib
    lib = _lib.load()
    assert lib.peps_backend_name() == b"cuda-sm_100a"
    run_k7_amplitudes(lib, 512)


# ---- the reference's physical t-J iPEPS tensors (tests/golden/ipeps_tj_ab.npz) in fermion mode -------------------------------
def run_physical_tj_parity(lib, rows, cols, trunc, W=2):
    """Sweeps + E_loc + O* of the t-J model on an OBC tiling of the reference's iPEPS unit cell (a physical fermionic state with a
    realistic boundary spectrum, not a random one) through the C ABI against oracle/fermion.py: chains bit-identical,
    amplitudes / E_loc / O* to 1e-10."""
    from parity_common import run_fermion_pipeline_parity, ipeps_tj_state, doped_tj_configs
    f = ipeps_tj_state(rows, cols)
    cfgs = doped_tj_configs(rows, cols, W, seed=3)
    return run_fermion_pipeline_parity(lib, rows, cols, 4, W, trunc, model="tj", nsweeps=2, state=(f, cfgs))


@pytest.mark.parametrize("rows,cols,trunc", [(6, 6, (8, 24, 1e-12)), (6, 8, (16, 32, 0.0))])
def test_physical_tj_ipeps_state_parity_hostsim(rows, cols, trunc):
    run_physical_tj_parity(hostsim_lib.load(), rows, cols, trunc)


@pytest.mark.gpu
@pytest.mark.parametrize("rows,cols,trunc", [(6, 8, (16, 32, 0.0))])
def test_physical_tj_ipeps_state_parity_gpu(rows, cols, trunc):
    from peps_b200 import _lib
    lib = _lib.load()
    assert lib.peps_backend_name() == b"cuda-sm_100a"
    run_physical_tj_parity(lib, rows, cols, trunc, W=3)
