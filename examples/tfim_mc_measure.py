"""Monte Carlo measurement of the transverse-field Ising model on the B200 path -- the flow of the reference's
examples/transverse_field_ising_mc_measure.cpp: MCPEPSMeasurer(state, mc_params, peps_params, model).Execute() + DumpData:
energy, spin_z, sigma_x (per site) and SzSz_row, written as stats/<key>_mean.csv / _stderr.csv like the reference.

Usage on a B200:  python examples/tfim_mc_measure.py [--samples 512] [--out ./tfim_measure]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from peps_b200.api import (BMPSTruncateParams, Configuration, MCPEPSMeasurer, MonteCarloParams,   # noqa: E402
                           MCUpdateSquareNNFullSpaceUpdate, TransverseFieldIsingSquareOBC)
from tfim_vmc_optimize import random_state                                                          # noqa: E402


def measure(state=None, rows=4, cols=4, D=4, chi=8, h=0.5, walkers=32, samples=512, seed=1, out=None, lib=None):
    state = state if state is not None else random_state(rows, cols, D, seed)
    init = Configuration(rows, cols).Random([rows * cols // 2, rows * cols - rows * cols // 2], seed=seed)
    mc = MonteCarloParams(num_samples=samples, num_warmup_sweeps=20, sweeps_between_samples=2, initial_config=init)
    m = MCPEPSMeasurer(mc, BMPSTruncateParams.SVD(2, chi, 1e-15), state, TransverseFieldIsingSquareOBC(h),
                       MCUpdateSquareNNFullSpaceUpdate(seed=seed), walkers, lib=lib)
    res = m.Execute()
    if out:
        m.DumpData(out)
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=512)
    ap.add_argument("--walkers", type=int, default=32)
    ap.add_argument("--out", default="./tfim_measure")
    a = ap.parse_args()
    r = measure(samples=a.samples, walkers=a.walkers, out=a.out)
    print(f"energy = {np.real(r['energy'][0]):+.8f} +- {r['energy'][1]:.2e};  <sigma_x> = {np.real(np.mean(r['sigma_x'][0])):.5f};"
          f"  stats written to {a.out}/stats")
