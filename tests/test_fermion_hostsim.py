"""Fermion mode of the engine (fZ2-graded tensors as sign-dressed dense tensors) through the C ABI on the test-only host
simulation of the device ops, against oracle/fermion.py (which is pinned by the reference's K8 energies)."""
import itertools
import os
import numpy as np
import pytest

import hostsim_lib
from parity_common import run_fermion_pipeline_parity
from peps_b200.api import (BMPSTruncateParams, FermionSplitIndexTPS, TableModel, WalkerBatch, MCEnergyGradEvaluator,
                           MonteCarloParams, Configuration, MCUpdateSquareNNExchange, PepsError)
from test_fermion_oracle import load_golden, perms


@pytest.fixture(scope="module")
def lib():
    return hostsim_lib.load()


@pytest.mark.parametrize("rows,cols,D,trunc", [
    (4, 4, 4, (8, 8, 0.0)),
    (3, 5, 2, (4, 4, 0.0)),
    (4, 4, 2, (2, 6, 1e-8)),
    (2, 2, 4, (1, 100, 0.0)),
])
def test_spinless_fermion_pipeline_parity_hostsim(lib, rows, cols, D, trunc):
    run_fermion_pipeline_parity(lib, rows, cols, D, 3, trunc, model="spinless", nsweeps=2)


def test_tj_pipeline_parity_hostsim(lib):
    run_fermion_pipeline_parity(lib, 4, 4, 4, 3, (8, 8, 0.0), model="tj", nsweeps=2)


def test_tj_nnn_hopping_pipeline_parity_hostsim(lib):
    """SquaretJNNNModel: t2 hopping on both diagonals through the BTen2 plaquette traces with the fermionic hop rules."""
    run_fermion_pipeline_parity(lib, 3, 4, 2, 3, (4, 4, 0.0), model="tj_nnn", nsweeps=2, t2=0.45)


@pytest.mark.parametrize("t2", [2.1, -2.5])
def test_k8_energy_through_abi(lib, t2):
    """K8 through the C ABI: exact summation over the 6 half-filled configurations of the 2x2 simple-update fixture."""
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


def test_fermion_mode_rejects_unsupported(lib):
    ftps = FermionSplitIndexTPS.random(3, 3, 2, 1)
    b = WalkerBatch(3, 3, 2, 2, 2, BMPSTruncateParams.SVD(4, 4, 0.0), lib=lib)
    b.set_fermion(ftps)
    with pytest.raises(PepsError):
        b.set_fermion(ftps)                                       # once only
    pair = np.zeros((4, 4)); pair[0, 3] = pair[3, 0] = 1.0        # pair creation: parities of both sites flip, no hop
    b.set_model(TableModel(2, pair))                              # both flip -> allowed structure (moves counted as hop)
    bad = np.zeros((4, 4)); bad[0, 1] = bad[1, 0] = 1.0           # one site parity flips
    with pytest.raises(PepsError):
        b.set_model(TableModel(2, bad))
    b.close()


def test_evaluator_gradient_fermion(lib):
    """MCEnergyGradEvaluator on a FermionSplitIndexTPS: energy and gradient equal the oracle chain's accumulation."""
    from oracle import fermion as F
    rows, cols, D, W, n = 3, 4, 2, 2, 3
    trunc = (4, 4, 0.0)
    f = F.FermionTPS.random(rows, cols, D, 21)
    ftps = FermionSplitIndexTPS(f.T, f.par, f.phys_par)
    from parity_common import fermion_configs
    cfgs = fermion_configs(rows, cols, W, 2)
    ev = MCEnergyGradEvaluator(MonteCarloParams(n * W, 0, 1, Configuration(cfgs[0]), True), BMPSTruncateParams.SVD(*trunc), ftps,
                               TableModel.spinless_fermion(1.0, 0.5, 0.2), MCUpdateSquareNNExchange(seed=77), W, configs=cfgs, lib=lib)
    res = ev.Evaluate(ftps)
    omodel = F.SpinlessFermionModel(1.0, 0.5, 0.2)
    es = []
    osum = [[[np.zeros_like(x) for x in site] for site in row] for row in f.T]
    eosum = [[[np.zeros_like(x) for x in site] for site in row] for row in f.T]
    for w in range(W):
        wk = F.FermionWalker(f, cfgs[w], trunc)
        up = F.FermionNNExchangeUpdater(77 + w)
        for _ in range(n):
            up.sweep(wk)
            e, ost, _ = omodel.energy_and_holes(wk, True)
            es.append((w, e))
            for r in range(rows):
                for c in range(cols):
                    s = int(wk.config[r, c])
                    osum[r][c][s] += ost[r][c]
                    eosum[r][c][s] += e * ost[r][c]
    N = n * W
    from peps_b200.api import combine_energy_bins
    emean, _ = combine_energy_bins(np.array([e for _, e in es]).reshape(W, n))
    assert abs(res.energy - emean) < 1e-10
    grad = np.concatenate([(eosum[r][c][s] / N - emean * osum[r][c][s] / N).ravel()
                           for r in range(rows) for c in range(cols) for s in range(2)])
    got = res.gradient.pack()
    assert np.max(np.abs(got - grad)) <= 1e-9 * max(1.0, np.max(np.abs(grad)))
    assert isinstance(res.gradient, FermionSplitIndexTPS)


def run_cpp_fermion_case(libdir, libfile, extra_link=()):
    """tests/cpp/test_cpp_fermion.cpp (the C++ wrapper: SetFermion + probe-built SpinlessFermion terms) on the 2x2
    simple-update fixture must print the reference's golden energy -4.98966397657 (t2 = -2.5)."""
    import subprocess
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    f, z = load_golden("sf2x2_t2_-2.5_double_su")
    cfgs = np.stack(perms([0, 0, 1, 1], 2, 2))
    ftps = FermionSplitIndexTPS(f.T, f.par, f.phys_par)
    flat, lp = ftps.pack(), ftps.leg_par_flat()
    with tempfile.TemporaryDirectory() as td:
        exe = os.path.join(td, "cpp_fermion")
        subprocess.check_call(["g++", "-std=c++17", "-O1", os.path.join(root, "tests", "cpp", "test_cpp_fermion.cpp"), "-o", exe,
                               "-L" + libdir, "-l:" + libfile, "-Wl,-rpath," + libdir] + list(extra_link))
        inp = (f"2 2 2 {ftps.bond_dim()} {len(cfgs)} 8 1.0 -2.5 0.0\n" + " ".join(map(str, ftps.phys_par)) + f"\n{lp.size} "
               + " ".join(map(str, lp)) + f"\n{flat.size} " + " ".join(repr(float(x)) for x in flat) + "\n"
               + " ".join(str(int(c)) for c in cfgs.ravel()) + "\n")
        out = subprocess.run([exe], input=inp, capture_output=True, text=True, check=True).stdout.split()
    assert abs(float(out[0]) - float(z["exp_energy"])) < 1e-9
    return float(out[1]), float(out[2]), ftps, cfgs


def check_cpp_fermion_evaluator(lib, e_cpp, gn_cpp, ftps, cfgs):
    """The C++ MCEnergyGradEvaluator built from model terms + parities equals the Python mirror on the same library."""
    W = len(cfgs)
    ev = MCEnergyGradEvaluator(MonteCarloParams(2 * W, 0, 1, Configuration(cfgs[0]), True), BMPSTruncateParams.SVD(8, 8, 1e-16), ftps,
                               TableModel.spinless_fermion(1.0, -2.5, 0.0), MCUpdateSquareNNExchange(seed=77), W, lib=lib)
    res = ev.Evaluate(ftps)
    assert abs(res.energy - e_cpp) < 1e-11 * max(1.0, abs(e_cpp))
    assert abs(res.gradient_norm - gn_cpp) < 1e-10 * max(1.0, gn_cpp)


def test_cpp_wrapper_fermion_golden():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = hostsim_lib.load()
    e, gn, ftps, cfgs = run_cpp_fermion_case(os.path.join(root, "tests", "hostsim"), "libpeps_hostsim.so")
    check_cpp_fermion_evaluator(lib, e, gn, ftps, cfgs)


def test_fermion_measure_bond_energies(lib):
    """peps_measure in fermion mode: the recorded bond energies add up to E_loc and 'charge' is the density."""
    from parity_common import fermion_configs
    rows, cols, D, W = 3, 4, 2, 3
    ftps = FermionSplitIndexTPS.random(rows, cols, D, 8)
    b = WalkerBatch(rows, cols, 2, D, W, BMPSTruncateParams.SVD(4, 4, 0.0), lib=lib)
    b.set_fermion(ftps)
    b.set_tps(ftps)
    b.set_model(TableModel.spinless_fermion(1.0, 0.4, 0.7))
    cfgs = fermion_configs(rows, cols, W, 2)
    b.set_configs(cfgs)
    b.init_walkers()
    e = b.energy_and_holes(False)
    m = b.measure()
    tot = m["bond_energy_h"].reshape(W, -1).sum(1) + m["bond_energy_v"].reshape(W, -1).sum(1) \
        + m["bond_energy_dr"].reshape(W, -1).sum(1) + m["bond_energy_ur"].reshape(W, -1).sum(1)
    assert np.allclose(tot, e, rtol=1e-12, atol=1e-12) and np.allclose(m["energy"], e, rtol=1e-12, atol=1e-12)
    assert np.array_equal(m["charge"], 1.0 - cfgs)
    assert np.any(m["bond_energy_dr"] != 0) and np.any(m["bond_energy_ur"] != 0)
    b.close()


def test_tj_jastrow_dressed_pipeline_parity_hostsim(lib):
    """MCUpdateSquareNNExchangeJastrowDressedTJ + the Jastrow-dressed t-J solver (square_nn_updater.h:380-438,
    square_tJ_model.h:352-410): chains bit-identical to the oracle, E_loc and O* to 1e-10."""
    run_fermion_pipeline_parity(lib, 4, 4, 2, 3, (4, 4, 0.0), model="tj", nsweeps=2, jastrow=True)
    run_fermion_pipeline_parity(lib, 3, 4, 2, 2, (4, 4, 0.0), model="spinless", nsweeps=2, jastrow=True)


def test_mcpeps_measurer_on_fermion_state(lib):
    """MCPEPSMeasurer (config #4: sampling + measurement) on an fZ2 state: energy / charge / bond-energy statistics."""
    from peps_b200.api import MCPEPSMeasurer
    from parity_common import fermion_configs
    rows, cols, D, W = 3, 4, 2, 4
    ftps = FermionSplitIndexTPS.random(rows, cols, D, 31)
    cfg = fermion_configs(rows, cols, 1, 2)[0]
    mc = MonteCarloParams(8, 1, 1, Configuration(cfg), False)
    m = MCPEPSMeasurer(mc, BMPSTruncateParams.SVD(4, 4, 0.0), ftps, TableModel.spinless_fermion(1.0, 0.3, 0.5),
                       MCUpdateSquareNNExchange(seed=5), W, lib=lib)
    out = m.Execute()
    assert set(out) == {"energy", "charge", "bond_energy_h", "bond_energy_v", "bond_energy_dr", "bond_energy_ur"}
    assert abs(out["charge"][0].sum() - 6.0) < 1e-12              # the exchange updater conserves the fermion number
    tot = out["bond_energy_h"][0].sum() + out["bond_energy_v"][0].sum() + out["bond_energy_dr"][0].sum() + out["bond_energy_ur"][0].sum()
    assert abs(tot - out["energy"][0]) < 1e-10


@pytest.mark.parametrize("sectors", ["1", "0"])
def test_block_jacobi_path_with_z2_sectors(lib, monkeypatch, sectors):
    """The full-rank truncation path (block Jacobi instead of the small SVD) in fermion mode: rows regrouped by parity
    sector, cross-sector block pairs skipped (PEPS_Z2_SECTORS=1, default) vs. the plain schedule; both match the oracle."""
    monkeypatch.setenv("PEPS_SMALL_SVD", "0")
    monkeypatch.setenv("PEPS_Z2_SECTORS", sectors)
    run_fermion_pipeline_parity(lib, 4, 4, 4, 2, (16, 16, 0.0), model="spinless", nsweeps=1)


@pytest.mark.parametrize("sectors", ["1", "0"])
def test_sector_truncation_hostsim(lib, monkeypatch, sectors):
    from parity_common import run_sector_truncation_case
    monkeypatch.setenv("PEPS_SMALL_SVD", "0")
    monkeypatch.setenv("PEPS_Z2_SECTORS", sectors)
    run_sector_truncation_case(lib, nr=96, nc=112, t=16, W=2)


def test_fermion_pipeline_block_jacobi_path_hostsim(lib, monkeypatch):
    """Same scenario as the GPU test of that name, smaller: the sector path inside the full pipeline vs the oracle."""
    monkeypatch.setenv("PEPS_SMALL_SVD", "0")
    run_fermion_pipeline_parity(lib, 5, 5, 4, 1, (16, 16, 0.0), model="spinless", nsweeps=1, check_holes=False)


def test_fermion_sr_natural_gradient(lib):
    """SR on an fZ2 state (the O* store holds the physical O* = Pi(R*) in the uploaded tensor entries, so the S matrix and
    the CG need no parity wrapping: docs/dev/design/math/fermion-vmc-implementation.md section 3): the device matvec /
    natural gradient against the dense S built from the oracle chain's O* samples."""
    from oracle import fermion as F
    from peps_b200 import sr
    from parity_common import fermion_configs
    rows, cols, D, W, n, trunc, shift = 3, 3, 2, 3, 4, (4, 4, 0.0), 1e-3
    f = F.FermionTPS.random(rows, cols, D, 13)
    ftps = FermionSplitIndexTPS(f.T, f.par, f.phys_par)
    cfgs = fermion_configs(rows, cols, W, 2)
    ev = MCEnergyGradEvaluator(MonteCarloParams(n * W, 0, 1, Configuration(cfgs[0]), True), BMPSTruncateParams.SVD(*trunc), ftps,
                               TableModel.spinless_fermion(1.0, 0.5, 0.2), MCUpdateSquareNNExchange(seed=41), W, configs=cfgs, lib=lib)
    res = ev.Evaluate(collect_sr_buffers=True)
    assert ev.batch.sr_count() == n * W
    omodel = F.SpinlessFermionModel(1.0, 0.5, 0.2)
    ostars = []
    for w in range(W):
        wk = F.FermionWalker(f, cfgs[w], trunc)
        up = F.FermionNNExchangeUpdater(41 + w)
        for _ in range(n):
            up.sweep(wk)
            _, ost, _ = omodel.energy_and_holes(wk, True)
            dense = [[[np.zeros_like(x) for x in site] for site in row] for row in f.T]
            for r in range(rows):
                for c in range(cols):
                    dense[r][c][int(wk.config[r, c])] = ost[r][c]
            ostars.append(np.concatenate([x.ravel() for row in dense for site in row for x in site]))
    o = np.stack(ostars)
    obar = o.mean(0)
    assert np.max(np.abs(res.Ostar_mean.pack() - obar)) < 1e-10 * np.max(np.abs(obar))
    g = res.gradient.pack()
    nat, iters, _ = ev.CalculateNaturalGradient(res, shift, sr.ConjugateGradientParams(max_iter=300, relative_tolerance=1e-11))
    s_dense = (o - obar).T @ (o - obar) / len(ostars) + shift * np.eye(obar.size)
    x_dense = np.linalg.solve(s_dense, g)
    assert np.max(np.abs(nat.pack() - x_dense)) < 1e-6 * np.max(np.abs(x_dense))


def test_set_fermion_after_set_tps(lib):
    """Order independence: a state uploaded before peps_set_fermion is dressed when the mode is switched on."""
    from parity_common import fermion_configs
    ftps = FermionSplitIndexTPS.random(3, 3, 2, 3)
    cfgs = fermion_configs(3, 3, 2, 2)
    amps = []
    for order in (0, 1):
        b = WalkerBatch(3, 3, 2, 2, 2, BMPSTruncateParams.SVD(4, 4, 0.0), lib=lib)
        if order == 0:
            b.set_fermion(ftps); b.set_tps(ftps)
        else:
            b.set_tps(ftps); b.set_fermion(ftps)
        b.set_configs(cfgs)
        b.init_walkers()
        amps.append(b.amplitudes())
        b.close()
    assert np.array_equal(amps[0], amps[1]) and np.all(amps[0] != 0)


@pytest.mark.parametrize("model,rows,cols,D,trunc", [
    ("spinless", 3, 3, 2, (4, 4, 0.0)),
    ("spinless", 3, 4, 3, (2, 6, 1e-9)),
    ("tj", 3, 3, 2, (4, 4, 0.0)),
    ("tj_nnn", 3, 3, 2, (4, 4, 0.0)),
])
def test_complex_fermion_pipeline_parity_hostsim(lib, model, rows, cols, D, trunc):
    """fZ2 tensors with complex entries (QLTEN_Complex + fZ2QN, the reference's *_complex fermion fixtures): the dressed
    planes, complex E_loc = ... conj(psi_ex / psi) and O* = conj(d psi / d T) / conj(psi_site) against oracle/fermion.py,
    which is pinned on the complex K8 goldens."""
    run_fermion_pipeline_parity(lib, rows, cols, D, 2, trunc, model=model, nsweeps=2, complex_=True)


def test_complex_tj_jastrow_dressed_pipeline_parity_hostsim(lib):
    run_fermion_pipeline_parity(lib, 3, 3, 2, 2, (4, 4, 0.0), model="tj", nsweeps=2, jastrow=True, complex_=True)


def test_k8_complex_goldens_through_abi(lib):
    from parity_common import run_complex_k8_goldens
    run_complex_k8_goldens(lib)


@pytest.mark.parametrize("complex_", [False, True])
def test_tj_singlet_pair_pinning_and_sc_bond_singlet_hostsim(lib, complex_):
    from parity_common import run_tj_pairing_parity
    run_tj_pairing_parity(lib, complex_=complex_)


def test_boson_bond_observable_hostsim(lib):
    from parity_common import run_boson_bond_observable_parity
    run_boson_bond_observable_parity(lib)


@pytest.mark.parametrize("updater,model,complex_", [("full_space", "tj", False), ("three_site", "tj", False),
                                                    ("full_space", "spinless", True), ("three_site", "tj_nnn", True)])
def test_fermion_full_space_and_three_site_updaters_hostsim(lib, updater, model, complex_):
    """The multi-state updaters on fZ2 tensors ("work for both fermion and boson", square_nn_updater.h:251): chains
    bit-identical to the oracle's restatement, E_loc / O* after every sweep."""
    run_fermion_pipeline_parity(lib, 3, 4, 2, 3, (4, 4, 0.0), model=model, nsweeps=2, updater=updater, complex_=complex_)


def test_evaluator_sr_and_measurer_on_complex_fermion_state(lib):
    """MCEnergyGradEvaluator (+ SR natural gradient, MCPEPSMeasurer) on a complex FermionSplitIndexTPS: energy and gradient
    = sum conj(E_loc) O* / N - conj(E) sum O* / N of the oracle chain."""
    from oracle import fermion as F
    from parity_common import fermion_configs
    from peps_b200 import sr
    from peps_b200.api import combine_energy_bins, MCPEPSMeasurer
    rows, cols, D, W, n = 3, 4, 2, 2, 3
    trunc = (4, 4, 0.0)
    f = F.FermionTPS.random(rows, cols, D, 21, complex_=True)
    ftps = FermionSplitIndexTPS(f.T, f.par, f.phys_par)
    cfgs = fermion_configs(rows, cols, W, 2)
    model = TableModel.spinless_fermion(1.0, 0.5, 0.2)
    ev = MCEnergyGradEvaluator(MonteCarloParams(n * W, 0, 1, Configuration(cfgs[0]), True), BMPSTruncateParams.SVD(*trunc), ftps,
                               model, MCUpdateSquareNNExchange(seed=77), W, configs=cfgs, lib=lib)
    res = ev.Evaluate(ftps, collect_sr_buffers=True)
    omodel = F.SpinlessFermionModel(1.0, 0.5, 0.2)
    es = []
    osum = [[[np.zeros_like(x) for x in site] for site in row] for row in f.T]
    eosum = [[[np.zeros_like(x) for x in site] for site in row] for row in f.T]
    for w in range(W):
        wk = F.FermionWalker(f, cfgs[w], trunc)
        up = F.FermionNNExchangeUpdater(77 + w)
        for _ in range(n):
            up.sweep(wk)
            e, ost, _ = omodel.energy_and_holes(wk, True)
            es.append(e)
            for r in range(rows):
                for c in range(cols):
                    s = int(wk.config[r, c])
                    osum[r][c][s] += ost[r][c]
                    eosum[r][c][s] += np.conj(e) * ost[r][c]
    N = n * W
    es = np.array(es).reshape(W, n)
    emean = complex(combine_energy_bins(es.real)[0], combine_energy_bins(es.imag)[0])
    assert abs(res.energy - emean) < 1e-10
    grad = np.concatenate([(eosum[r][c][s] / N - np.conj(emean) * osum[r][c][s] / N).ravel()
                           for r in range(rows) for c in range(cols) for s in range(2)])
    got = res.gradient.pack()
    assert np.iscomplexobj(got) and np.max(np.abs(got - grad)) <= 1e-9 * max(1.0, np.max(np.abs(grad)))
    assert isinstance(res.gradient, FermionSplitIndexTPS)
    nat, iters, resid = ev.CalculateNaturalGradient(res, 1e-3, sr.ConjugateGradientParams(max_iter=200, relative_tolerance=1e-8))
    assert iters > 0 and np.iscomplexobj(nat.pack())
    obs = MCPEPSMeasurer(MonteCarloParams(4, 0, 1, Configuration(cfgs[0]), True), BMPSTruncateParams.SVD(*trunc), ftps, model,
                         MCUpdateSquareNNExchange(seed=5), 2, lib=lib).Execute()
    assert np.iscomplexobj(obs["bond_energy_h"][0]) and obs["charge"][0].shape == (rows, cols)


def run_cpp_tj_pairing_case(lib, libdir, libfile, extra_link=()):
    """tests/cpp/test_cpp_tj_pairing.cpp: SetBondPin / MeasureBondTerm with the probe-built tJSingletPairTerm tables of the C++
    wrapper against the Python mirror (TableModel.SetSingletPairPinningField / measure_bond_observable) on the same library."""
    import subprocess
    import tempfile
    from oracle import fermion as F
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rows, cols, D, W, chi, delta = 3, 3, 2, 3, 4, 0.07
    f = F.FermionTPS.random(rows, cols, D, 23, phys_par=(1, 1, 0))
    ftps = FermionSplitIndexTPS(f.T, f.par, (1, 1, 0))
    rng = np.random.default_rng(8)
    cfgs = []
    while len(cfgs) < W:
        c = rng.integers(0, 3, size=(rows, cols))
        if f.parities(c).sum() % 2 == 0:
            cfgs.append(c)
    cfgs = np.stack(cfgs)
    flat, lp = ftps.pack(), ftps.leg_par_flat()
    pin = ((1, 0), (1, 1))
    with tempfile.TemporaryDirectory() as td:
        exe = os.path.join(td, "cpp_tj")
        subprocess.check_call(["g++", "-std=c++17", "-O1", os.path.join(root, "tests", "cpp", "test_cpp_tj_pairing.cpp"), "-o", exe,
                               "-L" + libdir, "-l:" + libfile, "-Wl,-rpath," + libdir] + list(extra_link))
        inp = (f"{rows} {cols} {D} {W} {chi} 1.0 0.3 0.2 {delta} {pin[0][0] * cols + pin[0][1]} {pin[1][0] * cols + pin[1][1]}\n"
               + f"{lp.size} " + " ".join(map(str, lp)) + f"\n{flat.size} " + " ".join(repr(float(x)) for x in flat) + "\n"
               + " ".join(str(int(c)) for c in cfgs.ravel()) + "\n")
        out = np.array(subprocess.run([exe], input=inp, capture_output=True, text=True, check=True).stdout.split(), dtype=float)
    b = WalkerBatch(rows, cols, 3, D, W, BMPSTruncateParams.SVD(chi, chi, 0.0), lib=lib)
    b.set_fermion(ftps); b.set_tps(ftps); b.set_configs(cfgs); b.init_walkers()
    b.set_model(TableModel.tj(1.0, 0.3, mu=0.2))
    e0 = b.energy_and_holes(False)
    b.set_model(TableModel.tj(1.0, 0.3, mu=0.2).SetSingletPairPinningField(pin[1], pin[0], delta))
    e1 = b.energy_and_holes(False)
    dd, d = TableModel.tj_singlet_pair_tables()
    parts = [e0, e1]
    for H in (dd, d):
        h, v = b.measure_bond_observable(H)
        parts += [h.ravel(), v.ravel()]
    ref = np.concatenate(parts)
    assert out.shape == ref.shape and np.max(np.abs(out - ref)) < 1e-12 * max(1.0, np.max(np.abs(ref)))
    assert np.max(np.abs(e1 - e0)) > 0 or np.max(np.abs(ref[2 * W:])) > 0
    b.close()


def test_cpp_wrapper_tj_pairing(lib):
    run_cpp_tj_pairing_case(lib, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"), "libpeps_hostsim.so")
