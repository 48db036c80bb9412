"""Complex (QLTEN_Complex) states on the CUDA path through the C ABI against the oracle (pinned on the reference's complex
goldens). Run on the GPU box with ``-m gpu``. Tolerance 1e-10 relative; configurations / acceptance counts bit-exact."""
import numpy as np
import pytest

from parity_common import run_complex_pipeline_parity, run_complex_k5_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from peps_b200 import _lib
    l = _lib.load()
    assert l.peps_backend_name() == b"cuda-sm_100a"
    return l


@pytest.mark.parametrize("rows,cols,D,trunc,W", [
    (2, 2, 3, (1, 100, 0.0), 2),
    (4, 4, 3, (6, 6, 0.0), 4),
    (3, 5, 2, (4, 4, 0.0), 3),
    (4, 4, 4, (2, 8, 1e-8), 3),
    (6, 6, 4, (16, 16, 0.0), 2),
])
def test_complex_pipeline_parity_gpu(lib, rows, cols, D, trunc, W):
    run_complex_pipeline_parity(lib, rows, cols, D, W, trunc, nsweeps=2)


def test_complex_j1j2_pipeline_parity_gpu(lib):
    run_complex_pipeline_parity(lib, 4, 4, 3, 3, (6, 6, 0.0), nsweeps=1, j2=0.5)


def test_k5_complex_golden_gpu(lib):
    run_complex_k5_golden(lib)


def test_complex_tfim_full_space_pipeline_parity_gpu(lib):
    """BASELINE config #1's flow (TFIM + full-space Suwa-Todo updater) on a complex state."""
    run_complex_pipeline_parity(lib, 4, 4, 3, 3, (6, 6, 0.0), nsweeps=2, tfim_h=0.7)


def test_complex_three_site_updater_pipeline_parity_gpu(lib):
    run_complex_pipeline_parity(lib, 4, 4, 3, 3, (6, 6, 0.0), nsweeps=2, three_site=True)


@pytest.mark.parametrize("table,j2", [("xxz", 0.4), ("spin1", 0.0)])
def test_complex_table_model_pipeline_parity_gpu(lib, table, j2):
    run_complex_pipeline_parity(lib, 3, 4, 2, 3, (4, 4, 0.0), nsweeps=1, j2=j2, table=table)


def test_complex_measure_parity_gpu(lib):
    from parity_common import run_complex_measure_parity
    run_complex_measure_parity(lib, 0.5)


def test_complex_sr_matvec_and_natural_gradient_gpu(lib):
    """The O* store as the real embedding of complex samples: matvec vs the dense Hermitian S of the oracle chain, CG vs the
    dense solve, on the CUDA kernels."""
    from parity_common import run_complex_sr
    run_complex_sr(lib)


def test_cpp_wrapper_complex_on_cuda_library(lib):
    import os
    from test_cpp_host import run_cpp_complex_case, ROOT
    run_cpp_complex_case(lib, os.path.join(ROOT, "peps_b200"), "libpeps_b200.so",
                         extra_link=["-Wl,-rpath,/usr/local/cuda/lib64", "-L/usr/local/cuda/lib64", "-lcudart"])


def test_complex_structure_factor_gpu(lib):
    from parity_common import run_structure_factor_parity
    run_structure_factor_parity(lib, complex_=True)


@pytest.mark.parametrize("scheme", [1, 2])
def test_complex_variational_compression_gpu(lib, scheme):
    from parity_common import run_variational_parity
    run_variational_parity(lib, scheme, complex_=True)
