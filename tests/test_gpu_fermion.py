"""Fermion mode (fZ2-graded tensors, BASELINE config #4) of the CUDA path through the C ABI against oracle/fermion.py,
which is pinned by the reference's K8 energies (tests/test_fermion_oracle.py). Run on the GPU box with ``-m gpu``.
Tolerance 1e-10 relative on |amplitudes|, E_loc and O*; configurations and acceptance counts bit-exact."""
import numpy as np
import pytest

from parity_common import run_fermion_pipeline_parity, fermion_configs
from peps_b200.api import BMPSTruncateParams, FermionSplitIndexTPS, TableModel, WalkerBatch
from test_fermion_oracle import load_golden, perms

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from peps_b200 import _lib
    l = _lib.load()
    assert l.peps_backend_name() == b"cuda-sm_100a"
    return l


@pytest.mark.parametrize("rows,cols,D,trunc,W", [
    (4, 4, 4, (8, 8, 0.0), 5),
    (3, 5, 2, (4, 4, 0.0), 3),
    (4, 4, 2, (2, 6, 1e-8), 4),
    (2, 2, 4, (1, 100, 0.0), 2),
    (6, 6, 4, (16, 16, 0.0), 2),
])
def test_spinless_fermion_pipeline_parity_gpu(lib, rows, cols, D, trunc, W):
    run_fermion_pipeline_parity(lib, rows, cols, D, W, trunc, model="spinless", nsweeps=2)


def test_tj_pipeline_parity_gpu(lib):
    run_fermion_pipeline_parity(lib, 4, 4, 4, 4, (8, 8, 0.0), model="tj", nsweeps=2)
    run_fermion_pipeline_parity(lib, 4, 4, 2, 3, (4, 4, 0.0), model="tj_nnn", nsweeps=2, t2=0.45)


@pytest.mark.parametrize("t2", [2.1, 0.0, -2.5])
def test_k8_spinless_fermion_energy_gpu(lib, t2):
    """The reference's own known answer through the CUDA path: exact summation over the 6 half-filled configurations of the
    2x2 simple-update fixture gives -4.1879072654 / -1.98218053854 / -4.98966397657
    (tests/test_algorithm/test_exact_summation_evaluator.cpp:428-458)."""
    f, z = load_golden(f"sf2x2_t2_{t2:+.1f}_double_su")
    cfgs = np.stack(perms([0, 0, 1, 1], 2, 2))
    ftps = FermionSplitIndexTPS(f.T, f.par, f.phys_par)
    b = WalkerBatch(2, 2, 2, ftps.bond_dim(), len(cfgs), BMPSTruncateParams.SVD(8, 8, 1e-16), lib=lib)
    b.set_fermion(ftps)
    b.set_tps(ftps)
    b.set_model(TableModel.spinless_fermion(1.0, t2, 0.0))
    b.set_configs(cfgs)
    b.init_walkers()
    e = b.energy_and_holes(False)
    w = b.amplitudes() ** 2
    assert abs(float(np.sum(w * e) / np.sum(w)) - float(z["exp_energy"])) < 1e-9
    b.close()


def test_k8_tj_energy_gpu(lib):
    f, z = load_golden("tj2x2_double_su")
    cfgs = np.stack(perms([0, 1, 2, 2], 2, 2))
    ftps = FermionSplitIndexTPS(f.T, f.par, f.phys_par)
    b = WalkerBatch(2, 2, 3, ftps.bond_dim(), len(cfgs), BMPSTruncateParams.SVD(4, 4, 0.0), lib=lib)
    b.set_fermion(ftps)
    b.set_tps(ftps)
    b.set_model(TableModel.tj(1.0, 0.3, V=0.075))
    b.set_configs(cfgs)
    b.init_walkers()
    e = b.energy_and_holes(False)
    w = b.amplitudes() ** 2
    assert abs(float(np.sum(w * e) / np.sum(w)) - float(z["exp_energy"])) < 1e-9       # -2.78008187385
    b.close()


def test_config4_8x8_D8_chi64_fermion_properties(lib):
    """BASELINE config #4 sizes (8x8, D = 8 with even / odd blocks of 4, chi = 64): size-independent properties.
    (i) the hole contracted with the site tensor rebuilds psi at every site: sum(O* . conj-free T) == 1;
    (ii) psi of the row machinery and of the column machinery agree in magnitude (psi list);
    (iii) a sweep keeps the fermion number and the chains move; (iv) E_loc is finite and walker-dependent."""
    rows = cols = 8
    D, chi, W = 8, 64, 4
    ftps = FermionSplitIndexTPS.random(rows, cols, D, 20260104)
    cfgs = fermion_configs(rows, cols, W, 2)
    b = WalkerBatch(rows, cols, 2, D, W, BMPSTruncateParams.SVD(chi, chi, 0.0), lib=lib)
    b.set_fermion(ftps)
    b.set_tps(ftps)
    b.set_model(TableModel.spinless_fermion(1.0, 0.0, 0.5))
    b.set_configs(cfgs)
    b.seed_rng(np.arange(40, 40 + W))
    b.init_walkers()
    acc = b.sweep(1)
    assert np.all(acc > 0)
    c = b.get_configs()
    assert np.all((1 - c).reshape(W, -1).sum(1) == 32)
    e, psi = b.energy_and_holes(True, True)
    assert np.all(np.isfinite(e)) and np.ptp(e) > 0
    amp = b.amplitudes()
    # the truncated boundary-MPS contraction of a random signed state agrees between paths only roughly; magnitudes
    # of the same path (rows among themselves) are the meaningful closure
    rowpsi = np.abs(psi[:rows])
    assert np.all(rowpsi > 0)
    holes = b.holes()
    flat = ftps.pack()
    off = 0
    hoff = 0
    for r in range(rows):
        for cc in range(cols):
            sz = ftps.t[r][cc][0].size
            for w in range(W):
                s = int(c[w, r, cc])
                t = flat[off + s * sz: off + (s + 1) * sz]
                ostar = holes[w, hoff:hoff + sz] / amp[w]
                assert abs(float(np.dot(ostar, t)) - 1.0) < 1e-9, (r, cc, w)
            off += 2 * sz
            hoff += sz
    b.close()


def test_cpp_wrapper_fermion_on_cuda_library(lib):
    """The C++ wrapper (SetFermion + probe-built SquareSpinlessFermion terms) linked against libpeps_b200.so reproduces
    the reference's golden energy of the 2x2 simple-update fixture (t2 = -2.5)."""
    import os
    from test_fermion_hostsim import run_cpp_fermion_case, check_cpp_fermion_evaluator
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    e, gn, ftps, cfgs = run_cpp_fermion_case(os.path.join(root, "peps_b200"), "libpeps_b200.so",
                                             extra_link=["-Wl,-rpath,/usr/local/cuda/lib64", "-L/usr/local/cuda/lib64", "-lcudart"])
    check_cpp_fermion_evaluator(lib, e, gn, ftps, cfgs)


def test_tj_jastrow_dressed_pipeline_parity_gpu(lib):
    """Jastrow-dressed t-J sampling (MCUpdateSquareNNExchangeJastrowDressedTJ + dressed solver) on the CUDA path."""
    run_fermion_pipeline_parity(lib, 4, 4, 4, 4, (8, 8, 0.0), model="tj", nsweeps=2, jastrow=True)


@pytest.mark.parametrize("sectors", ["1", "0"])
def test_sector_truncation_gpu(lib, monkeypatch, sectors):
    """Block-Jacobi truncation of two-sector (block-diagonal up to permutations) Theta matrices at config #4's shape
    (512 x 512, chi = 64), with and without the sector regrouping / cross-sector pair skipping."""
    from parity_common import run_sector_truncation_case
    monkeypatch.setenv("PEPS_SMALL_SVD", "0")
    monkeypatch.setenv("PEPS_Z2_SECTORS", sectors)
    run_sector_truncation_case(lib, nr=160, nc=192, t=24, W=3)
    run_sector_truncation_case(lib, nr=512, nc=512, t=64, W=2, seed=9)
    run_sector_truncation_case(lib, nr=70, nc=81, t=10, W=2, seed=3)        # odd width: scalar (non-TMA) tile path, odd windows
    run_sector_truncation_case(lib, nr=41, nc=130, t=12, W=2, seed=5)       # wide matrix


@pytest.mark.parametrize("sectors", ["1", "0"])
def test_fermion_pipeline_block_jacobi_path_gpu(lib, monkeypatch, sectors):
    """Full pipeline vs the oracle with the truncations forced onto the block-Jacobi path (PEPS_SMALL_SVD=0), with and
    without the sector regrouping / column windows: configurations bit-identical, E_loc and O* to 1e-10."""
    monkeypatch.setenv("PEPS_SMALL_SVD", "0")
    monkeypatch.setenv("PEPS_Z2_SECTORS", sectors)
    run_fermion_pipeline_parity(lib, 6, 6, 4, 2, (16, 16, 0.0), model="spinless", nsweeps=1)


@pytest.mark.parametrize("model,rows,cols,D,trunc,W", [
    ("spinless", 4, 4, 4, (8, 8, 0.0), 3),
    ("spinless", 4, 4, 2, (2, 6, 1e-8), 3),
    ("tj_nnn", 3, 4, 2, (4, 4, 0.0), 3),
    ("spinless", 6, 6, 4, (16, 16, 0.0), 2),
    ("tj", 6, 6, 4, (8, 16, 1e-9), 2),
])
def test_complex_fermion_pipeline_parity_gpu(lib, model, rows, cols, D, trunc, W):
    """fZ2 tensors with complex entries: dressed planes, complex E_loc and O* on the CUDA path against the oracle."""
    run_fermion_pipeline_parity(lib, rows, cols, D, W, trunc, model=model, nsweeps=2, complex_=True)


def test_complex_tj_jastrow_dressed_pipeline_parity_gpu(lib):
    run_fermion_pipeline_parity(lib, 4, 4, 2, 3, (4, 4, 0.0), model="tj", nsweeps=2, jastrow=True, complex_=True)


def test_k8_complex_goldens_gpu(lib):
    """The reference's complex fZ2 known answers (spinless fermions t2 = 2.1 / 0 / -2.5, t-J; su and lowest states) through
    the CUDA path."""
    from parity_common import run_complex_k8_goldens
    run_complex_k8_goldens(lib)


@pytest.mark.parametrize("complex_", [False, True])
def test_tj_singlet_pair_pinning_and_sc_bond_singlet_gpu(lib, complex_):
    """Singlet-pair pinning field in E_loc and the SC_bond_singlet observable (pair creation / annihilation targets) on the
    CUDA path against the oracle."""
    from parity_common import run_tj_pairing_parity
    run_tj_pairing_parity(lib, complex_=complex_)


def test_boson_bond_observable_gpu(lib):
    from parity_common import run_boson_bond_observable_parity
    run_boson_bond_observable_parity(lib)


@pytest.mark.parametrize("updater,model,complex_", [("full_space", "tj", False), ("three_site", "tj", False),
                                                    ("three_site", "tj_nnn", True)])
def test_fermion_full_space_and_three_site_updaters_gpu(lib, updater, model, complex_):
    run_fermion_pipeline_parity(lib, 4, 4, 2, 3, (4, 4, 0.0), model=model, nsweeps=2, updater=updater, complex_=complex_)
