"""Re-packs the reference's fZ2 (fermionic) 2x2 fixtures as dense ``.npz`` files under tests/golden/.

Run in the build container (needs /root/reference):  python tests/golden/make_fermion_golden.py
Stored per fixture: ``t_{r}_{c}_{s}`` dense (L, D, R, U) arrays (the dim-1 parity leg dropped), ``par_{r}_{c}_{k}``
the parity of every index value of leg k in (L, D, R, U), ``phys_par`` the parity of each physical state, and the
constants the reference's tests assert (tests/test_algorithm/test_exact_summation_evaluator.cpp:137-151, 353-458,
807, 905-990).
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.fermion import FermionTPS  # noqa: E402

REF = "/root/reference/tests/test_data"


def pack(f):
    out = dict(rows=f.rows, cols=f.cols, phys=f.phys, phys_par=np.array(f.phys_par))
    for r in range(f.rows):
        for c in range(f.cols):
            for s in range(f.phys):
                out[f"t_{r}_{c}_{s}"] = f.T[r][c][s]
            for k in range(4):
                out[f"par_{r}_{c}_{k}"] = f.par[r][c][k]
    return out


def free_fermion_2x2(t, t2):
    """Calculate2x2OBCSpinlessFreeFermionEnergy (test_exact_summation_evaluator.cpp:137-151)."""
    e = [-2.0 * t * np.cos(k) - t2 * np.cos(2.0 * k) for k in (0.0, np.pi / 2, np.pi, 1.5 * np.pi)]
    return float(sum(x for x in e if x < 0))


def main():
    su = {2.1: -4.1879072654, 0.0: -1.98218053854, -2.5: -4.98966397657}
    gnorm = {2.1: 1.152438300797112e-12, 0.0: 2.184991439005157e-17, -2.5: 1.928582676855609e-10}
    gprobe = {2.1: 3.683776728378202e-13, 0.0: 7.407222090395872e-18, -2.5: 6.551465833461134e-11}
    gnorm_c = {2.1: 8.774999537624253e-15, 0.0: 3.080265654142141e-17, -2.5: 1.489500558224182e-15}
    for t2 in (2.1, 0.0, -2.5):
        for cx in (False, True):
            ty = "complex" if cx else "double"
            f = FermionTPS.load(os.path.join(REF, f"spinless_fermion_tps_t2_{t2:.6f}_{ty}lowest"), cx)
            np.savez(os.path.join(HERE, f"sf2x2_t2_{t2:+.1f}_{ty}_lowest.npz"), **pack(f), t=1.0, t2=t2,
                     exp_energy=free_fermion_2x2(1.0, t2), exp_energy_tol=1e-6,
                     exp_grad_norm=(gnorm_c if cx else gnorm)[t2], exp_grad_probe_re=(np.nan if cx else gprobe[t2]))
            f = FermionTPS.load(os.path.join(REF, f"spinless_fermion_tps_t2_{t2:.6f}_{ty}_from_simple_update"), cx)
            np.savez(os.path.join(HERE, f"sf2x2_t2_{t2:+.1f}_{ty}_su.npz"), **pack(f), t=1.0, t2=t2,
                     exp_energy=su[t2], exp_energy_tol=1e-9)
    for cx in (False, True):
        ty = "complex" if cx else "double"
        f = FermionTPS.load(os.path.join(REF, f"tj_model_tps_{ty}lowest"), cx)
        np.savez(os.path.join(HERE, f"tj2x2_{ty}_lowest.npz"), **pack(f), t=1.0, J=0.3, mu=0.0,
                 exp_energy=-2.9431635706137875, exp_energy_tol=1e-6)
        f = FermionTPS.load(os.path.join(REF, f"tj_model_tps_{ty}_from_simple_update"), cx)
        np.savez(os.path.join(HERE, f"tj2x2_{ty}_su.npz"), **pack(f), t=1.0, J=0.3, mu=0.0,
                 exp_energy=-2.78008187385, exp_energy_tol=1e-9)
    print("fermion golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
