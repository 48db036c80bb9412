"""Generates the committed golden fixtures under tests/golden/ from the reference's own test data.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
The ``.qlten`` payloads are re-packed as ``.npz`` (dense arrays, legs (L, D, R, U)); the expected
numbers stored beside them are the constants the reference's tests assert
(tests/test_algorithm/test_exact_summation_evaluator.cpp:578-606, 700-790,
 tests/slow_tests/test_boson_mc_peps_measure.cpp:31-76).
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.qlten_io import load_tps_dir, load_configuration  # noqa: E402

REF = "/root/reference/tests"


def pack(tps):
    out = {}
    for r, row in enumerate(tps):
        for c, site in enumerate(row):
            for s, t in enumerate(site):
                out[f"t_{r}_{c}_{s}"] = t
    return out


def main():
    two_by_two = [
        ("heis2x2_double_lowest", "test_data/heisenberg_tps_doublelowest", False,
         dict(energy=-2.0, energy_tol=1e-6, grad_norm=2.69115141087757e-08, grad_probe_re=8.719330571244627e-09,
              grad_probe_im=0.0)),
        ("heis2x2_complex_lowest", "test_data/heisenberg_tps_complexlowest", True,
         dict(energy=-2.0, energy_tol=1e-6, grad_norm=2.277663798157925e-08, grad_probe_re=7.37963070602707e-09,
              grad_probe_im=1.708247848617349e-10)),
        ("heis2x2_double_su", "test_data/heisenberg_tps_double_from_simple_update", False,
         dict(energy=-1.99521278793, energy_tol=1e-10)),
        # TrivialTransverseIsingTest (test_exact_summation_evaluator.cpp:620-790), J = h = 1, all 16 configurations
        ("tfim2x2_double_lowest", "test_data/transverse_ising_tps_doublelowest", False,
         dict(energy=-5.226251859505504, energy_tol=1e-7, grad_norm=1.290630314256308e-10,
              grad_probe_re=4.081475798300479e-11, grad_probe_im=0.0)),
        ("tfim2x2_double_su", "test_data/transverse_ising_tps_double_from_simple_update", False,
         dict(energy=-5.19991995228, energy_tol=1e-10)),
    ]
    for name, rel, cx, exp in two_by_two:
        tps = load_tps_dir(os.path.join(REF, rel), 2, 2, 2, cx)
        np.savez(os.path.join(HERE, name + ".npz"), rows=2, cols=2, phys=2, **pack(tps),
                 **{"exp_" + k: v for k, v in exp.items()})
    d = os.path.join(REF, "slow_tests/test_data/tps_square_heisenberg4x4D8Double")
    tps = load_tps_dir(d, 4, 4, 2, False)
    cfgs = np.stack([load_configuration(os.path.join(d, f"configuration{i}")) for i in range(32)])
    np.savez_compressed(os.path.join(HERE, "heis4x4_D8_double.npz"), rows=4, cols=4, phys=2, configs=cfgs,
                        exp_e0_state=-9.18912, exp_E0_ED=-9.189207065192933, **pack(tps))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
