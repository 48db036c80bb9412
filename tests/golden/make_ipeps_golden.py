"""Re-packs the reference's t-J iPEPS unit-cell tensors (tests/test_data/ipeps_tJ_t{a,b}_doping0.125.qlten, fZ2 block-sparse,
legs (L, D, R, U, phys) with phys = (up, down, empty)) as dense arrays + leg parities for the physical-fermion-state tests
(the reference tiles them to a 20 x 24 OBC lattice in tests/test_2d_tn/test_bmps_contractor.cpp:688-847).
Run in the build container (needs /root/reference): python tests/golden/make_ipeps_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from peps_b200.io import read_qlten_fz2      # noqa: E402

REF = "/root/reference/tests/test_data/"

if __name__ == "__main__":
    ta, par, dirs = read_qlten_fz2(REF + "ipeps_tJ_ta_doping0.125.qlten")
    tb, par_b, dirs_b = read_qlten_fz2(REF + "ipeps_tJ_tb_doping0.125.qlten")
    assert dirs == dirs_b == [-1, 1, 1, -1, -1] and all((a == b).all() for a, b in zip(par, par_b))
    np.savez_compressed(os.path.join(HERE, "ipeps_tj_ab.npz"), ta=ta, tb=tb, leg_par=np.stack(par[:4]), phys_par=par[4])
    print("stored", ta.shape, [list(map(int, p)) for p in par])
