"""VMC optimisation of the transverse-field Ising model on the B200 path -- the flow of the reference's
examples/transverse_field_ising_vmc_optimize.cpp (BASELINE config #1: 4x4, D = 4, chi = 8, SR):

    state -> MCEnergyGradEvaluator.Evaluate(state, collect_sr_buffers=True)   (walker-batched sampling, E_loc, O*)
          -> CalculateNaturalGradient(result, diag_shift, cg_params)          (S-matrix-free CG on the device)
          -> state -= step * natural_gradient                                  (the optimizer's update rule, host side)

The optimizer algebra itself (SGD / SR step, learning rate) is ordinary host code on packed vectors; everything on the
sampling path runs through libpeps_b200. Usage on a B200:  python examples/tfim_vmc_optimize.py [--iters 20]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from peps_b200 import sr                                                        # noqa: E402
from peps_b200.api import (BMPSTruncateParams, Configuration, MCEnergyGradEvaluator, MonteCarloParams,   # noqa: E402
                           MCUpdateSquareNNFullSpaceUpdate, SplitIndexTPS, TransverseFieldIsingSquareOBC)


def random_state(rows, cols, D, seed):
    """A positive random SplitIndexTPS (the reference example loads a simple-update PEPS instead)."""
    rng = np.random.default_rng(seed)
    t = [[[rng.random((D if c > 0 else 1, D if r < rows - 1 else 1, D if c < cols - 1 else 1, D if r > 0 else 1)) + 0.1
           for _ in range(2)] for c in range(cols)] for r in range(rows)]
    return SplitIndexTPS(t)


def optimize(rows=4, cols=4, D=4, chi=8, h=0.5, walkers=32, samples=512, iters=20, step=0.1, diag_shift=1e-3, seed=1, lib=None,
             log=print):
    state = random_state(rows, cols, D, seed)
    init = Configuration(rows, cols).Random([rows * cols // 2, rows * cols - rows * cols // 2], seed=seed)
    mc = MonteCarloParams(num_samples=samples, num_warmup_sweeps=20, sweeps_between_samples=2, initial_config=init)
    ev = MCEnergyGradEvaluator(mc, BMPSTruncateParams.SVD(2, chi, 1e-15), state, TransverseFieldIsingSquareOBC(h),
                               MCUpdateSquareNNFullSpaceUpdate(seed=seed), walkers, lib=lib)   # TFIM does not conserve Sz
    ev.EnsureConfigurationValidity()
    ev.WarmUp()
    state = ev.state                                       # rescaled by NormalizeStateOrder1
    cg = sr.ConjugateGradientParams(max_iter=100, relative_tolerance=1e-5, residual_recompute_interval=20)
    energies = []
    for it in range(iters):
        res = ev.Evaluate(state, collect_sr_buffers=True)
        nat, cg_iters, resid = ev.CalculateNaturalGradient(res, diag_shift, cg)
        state = state - nat * step                         # StochasticReconfigurationUpdate: state += -lr * x
        energies.append(float(np.real(res.energy)))
        log(f"iter {it:3d}  E = {res.energy:+.8f} +- {res.energy_error:.2e}  |grad|^2 = {res.gradient_norm:.3e}  "
            f"CG {cg_iters} its  accept {res.accept_rates_avg[0]:.2f}")
    return energies, state


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--walkers", type=int, default=32)
    ap.add_argument("--samples", type=int, default=512)
    a = ap.parse_args()
    optimize(iters=a.iters, walkers=a.walkers, samples=a.samples)
