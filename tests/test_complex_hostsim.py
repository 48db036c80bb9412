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
