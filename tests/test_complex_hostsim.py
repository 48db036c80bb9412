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
    with pytest.raises(PepsError):
        b.set_fermion(FermionSplitIndexTPS.random(3, 3, 2, 1))
    b.close()


def test_complex_tall_chain_matrices_are_reduced_by_chunks(lib, monkeypatch):
    """The real embeddings of the complex chain matrices exceed the two-stage CAQR height at the headline size; they are
    reduced by row chunks first. Forced here at a small size (PEPS_QR_MAX_ROWS=64): results unchanged."""
    monkeypatch.setenv("PEPS_QR_MAX_ROWS", "64")
    run_complex_pipeline_parity(lib, 4, 4, 3, 2, (6, 6, 0.0), nsweeps=1)
