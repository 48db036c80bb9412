"""K7 golden amplitudes (tests/golden/heis4x4_D8_all_amplitudes.npz: the oracle amplitude of every S_z = 0 configuration of the
reference's 4x4 D=8 Heisenberg fixture, whose correlators match the reference's ED table in tests/test_oracle_kat.py) against
the product path: a random sample of configurations through the C ABI at the reference's truncation (8, 16, 1e-15)."""
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
