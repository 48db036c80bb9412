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
