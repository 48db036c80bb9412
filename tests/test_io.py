"""Product-side readers / writers of the reference's on-disk formats."""
import os

import numpy as np
import pytest

from helpers import load_golden_tps
from oracle import vmc
from peps_b200 import io as pio
from peps_b200.api import SplitIndexTPS, Configuration

REF = "/root/reference/tests"


def test_tps_and_configuration_round_trip(tmp_path):
    tps = SplitIndexTPS(vmc.random_tps(3, 4, 2, 3, seed=1))
    pio.dump_tps(tps, str(tmp_path / "tps"))
    back = pio.load_tps(str(tmp_path / "tps"))
    assert np.array_equal(back.pack(), tps.pack())
    cfg = Configuration(vmc.shuffled_half_filled_config(3, 4, 9))
    pio.dump_configuration(cfg, str(tmp_path / "configuration0"))
    assert pio.load_configuration(str(tmp_path / "configuration0")) == cfg


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
def test_reads_reference_fixture_and_rewrites_it_byte_for_byte(tmp_path):
    d = os.path.join(REF, "slow_tests/test_data/tps_square_heisenberg4x4D8Double")
    tps = pio.load_tps(d)
    gold, z = load_golden_tps("heis4x4_D8_double")
    assert np.array_equal(tps.pack(), SplitIndexTPS(gold).pack())
    assert np.array_equal(pio.load_configuration(os.path.join(d, "configuration3")).data, z["configs"][3])
    # with the fixture as header template the writer reproduces the reference's files exactly
    pio.dump_tps(tps, str(tmp_path / "out"), template_dir=d)
    for name in ("tps_ten1_1_0.qlten", "tps_ten0_0_1.qlten", "tps_ten3_3_0.qlten"):
        assert open(os.path.join(d, name), "rb").read() == open(tmp_path / "out" / name, "rb").read()


def test_read_qlten_checks_the_payload(tmp_path):
    a = np.arange(24, dtype=np.float64).reshape(2, 3, 4)
    p = str(tmp_path / "t.qlten")
    pio.write_qlten(p, a)
    assert np.array_equal(pio.read_qlten(p), a)
    raw = open(p, "rb").read()
    open(p, "wb").write(raw[:-17])                       # truncated payload
    with pytest.raises(ValueError):
        pio.read_qlten(p)
    with pytest.raises(ValueError):                      # a real file read as complex
        open(p, "wb").write(raw)
        pio.read_qlten(p, dtype=np.complex128)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
def test_complex_fixture_is_detected():
    d = os.path.join(REF, "slow_tests/test_data/tps_square_heisenberg4x4D8Complex")
    t = pio.read_qlten(os.path.join(d, "tps_ten1_1_0.qlten"))
    assert t.dtype == np.complex128 and t.shape == (8, 8, 8, 8)
    with pytest.raises(ValueError):
        pio.read_qlten(os.path.join(d, "tps_ten1_1_0.qlten"), dtype=np.float64)


def test_measurement_stats_dump_layout(tmp_path):
    """MCPEPSMeasurer::DumpData layout (monte_carlo_peps_measurer_impl.h:266-330, 544-620): 2-D observables as
    <key>_mean.csv / <key>_stderr.csv (17 significant digits, scientific), flat ones as index,mean,stderr."""
    from peps_b200.api import dump_measurement_stats
    res = {"spin_z": (np.array([[0.5, -0.25], [0.125, 1 / 3]]), np.array([[0.1, 0.2], [0.3, 0.4]])),
           "energy": (np.array(-1.9952127879312345), np.array(1e-3))}
    dump_measurement_stats(res, str(tmp_path / "out"), {"lx": 2, "ly": 2})
    rows = open(tmp_path / "out" / "stats" / "spin_z_mean.csv").read().splitlines()
    assert rows[0] == "5.0000000000000000e-01,-2.5000000000000000e-01" and float(rows[1].split(",")[1]) == 1 / 3
    flat = open(tmp_path / "out" / "stats" / "energy.csv").read().splitlines()
    assert flat[0] == "index,mean,stderr" and flat[1].startswith("0,-1.9952127879312345e+00,")
    meta = open(tmp_path / "out" / "metadata.txt").read()
    assert meta.startswith("format_version 1\n") and "lx 2" in meta


def test_fermion_tps_loader_matches_repacked_fixture():
    """peps_b200.io.load_fermion_tps on the reference's fZ2 fixture == the committed re-packed golden (needs the
    reference tree: skipped on the GPU box)."""
    import os
    import pytest
    d = "/root/reference/tests/test_data/spinless_fermion_tps_t2_2.100000_double_from_simple_update"
    if not os.path.isdir(d):
        pytest.skip("reference fixtures not present")
    from peps_b200.io import load_fermion_tps
    from test_fermion_oracle import load_golden
    f = load_fermion_tps(d)
    g, _ = load_golden("sf2x2_t2_+2.1_double_su")
    assert f.phys_par == g.phys_par == (1, 0)
    for r in range(2):
        for c in range(2):
            for s in range(2):
                assert np.array_equal(f.t[r][c][s], g.T[r][c][s])
            for k in range(4):
                assert np.array_equal(f.par[r][c][k], g.par[r][c][k])


@pytest.mark.parametrize("name", ["spinless_fermion_tps_t2_2.100000_double_from_simple_update",
                                  "spinless_fermion_tps_t2_-2.500000_complexlowest",
                                  "tj_model_tps_double_from_simple_update"])
def test_fermion_tps_writer_reproduces_reference_files(tmp_path, name):
    """dump_fermion_tps with the fixture as header template rewrites every fZ2 file of the reference byte for byte (double
    and complex); an edited state reloads with the edit."""
    d = os.path.join(REF, "test_data", name)
    if not os.path.isdir(d):
        pytest.skip("reference fixtures not present")
    f = pio.load_fermion_tps(d)
    assert np.iscomplexobj(f.t[0][0][0]) == ("complex" in name)
    pio.dump_fermion_tps(f, str(tmp_path / "out"), d)
    for fn in sorted(os.listdir(d)):
        if fn.endswith(".qlten") and fn.startswith("tps_ten"):
            assert open(os.path.join(d, fn), "rb").read() == open(tmp_path / "out" / fn, "rb").read(), fn
    g = f * 0.5
    g = type(f)(g.t, f.par, f.phys_par)
    pio.dump_fermion_tps(g, str(tmp_path / "half"), d)
    back = pio.load_fermion_tps(str(tmp_path / "half"))
    assert np.array_equal(back.pack(), 0.5 * f.pack())
    bad = type(f)([[[x + 1.0 for x in site] for site in row] for row in f.t], f.par, f.phys_par)   # parity-violating entries
    with pytest.raises(ValueError):
        pio.dump_fermion_tps(bad, str(tmp_path / "bad"), d)


def test_measurement_stats_dump_complex_values(tmp_path):
    """QLTEN_Complex measurement runs: ToCsvString(std::complex<double>) streams (re,im) (monte_carlo_peps_measurer_impl.h:31-36)."""
    from peps_b200.api import dump_measurement_stats
    res = {"bond_energy_h": (np.array([[0.5 + 0.25j, -1.0j]]), np.array([[0.1, 0.2]])), "energy": (np.array(-2.0 + 1e-3j), np.array(1e-4))}
    dump_measurement_stats(res, str(tmp_path / "out"))
    row = open(tmp_path / "out" / "stats" / "bond_energy_h_mean.csv").read().splitlines()[0]
    assert row == "(5.0000000000000000e-01,2.5000000000000000e-01),(-0.0000000000000000e+00,-1.0000000000000000e+00)" or \
        row == "(5.0000000000000000e-01,2.5000000000000000e-01),(0.0000000000000000e+00,-1.0000000000000000e+00)"
    flat = open(tmp_path / "out" / "stats" / "energy.csv").read().splitlines()
    assert flat[1].startswith("0,(-2.0000000000000000e+00,1.0000000000000000e-03),")
