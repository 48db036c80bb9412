"""Stochastic reconfiguration on top of the device-resident O* sample store.

Host-side restatement (plumbing; O(P) vector algebra once per CG iteration) of
  * ConjugateGradientSolver        utility/conjugate_gradient_solver.h:181-276 (serial form; with several GPUs every rank
                                   runs the identical vector updates on the all-reduced matvec, so the reference's
                                   master/slave broadcast of v, :355-611, is not needed)
  * ConjugateGradientParams        optimizer/optimizer_params.h:50-57
  * SRSMatrix::operator*           optimizer/stochastic_reconfiguration_smatrix.h:45-91 (the sum over samples runs in the
                                   CUDA kernels sr_dots / sr_accumulate; normalisation + diag_shift here)
  * Optimizer::CalculateNaturalGradient  optimizer/optimizer_impl.h:1031-1089
"""
import math
from dataclasses import dataclass

import numpy as np

CONVERGED, MAX_ITERATIONS, STAGNATED, INDEFINITE_MATRIX, NUMERICAL_BREAKDOWN = range(5)
REASONS = ("kConverged", "kMaxIterations", "kStagnated", "kIndefiniteMatrix", "kNumericalBreakdown")


@dataclass
class ConjugateGradientParams:
    max_iter: int = 100
    relative_tolerance: float = 1e-4
    absolute_tolerance: float = 0.0
    residual_recompute_interval: int = 20
    orthogonality_threshold: float = 0.5


@dataclass
class CGResult:
    x: np.ndarray
    residual_norm: float
    iterations: int
    reason: int


def conjugate_gradient(matvec, b, x0, params: ConjugateGradientParams) -> CGResult:
    """Line-by-line restatement of the reference's serial solver (best-iterate tracking, stagnation / NaN /
    indefiniteness exits, orthogonality restart, periodic residual recomputation)."""
    nrm2 = lambda v: float(np.dot(v, v))
    rhs_norm_sq = nrm2(b)
    tol_sq = max(params.relative_tolerance ** 2 * rhs_norm_sq, params.absolute_tolerance ** 2)
    r = b - matvec(x0)
    r_norm_sq = nrm2(r)
    if r_norm_sq <= tol_sq:
        return CGResult(x0, math.sqrt(r_norm_sq), 0, CONVERGED)
    p, x, best_x = r.copy(), x0.copy(), x0.copy()
    best_sq = r_norm_sq
    r_prev = r.copy()
    rkp1 = r_norm_sq
    stagnation = 0
    eps = np.finfo(np.float64).eps
    for k in range(params.max_iter):
        rk = rkp1
        ap = matvec(p)
        pap = float(np.dot(p, ap))
        if not (math.isfinite(pap) and pap > 0.0):
            return CGResult(best_x, math.sqrt(best_sq), k, INDEFINITE_MATRIX)
        alpha = rk / pap
        x = x + alpha * p
        if alpha * alpha * nrm2(p) < eps * eps * nrm2(x):
            stagnation += 1
            if stagnation >= 3:
                return CGResult(best_x, math.sqrt(best_sq), k + 1, STAGNATED)
        else:
            stagnation = 0
        ri = params.residual_recompute_interval
        if ri > 0 and (k % ri) == ri - 1:
            r = b - matvec(x)
        else:
            r = r - alpha * ap
        rkp1 = nrm2(r)
        if not math.isfinite(rkp1):
            return CGResult(best_x, math.sqrt(best_sq), k + 1, NUMERICAL_BREAKDOWN)
        if rkp1 < best_sq:
            best_x, best_sq = x.copy(), rkp1
        if rkp1 <= tol_sq:
            return CGResult(x, math.sqrt(rkp1), k + 1, CONVERGED)
        if k > 0:
            if abs(float(np.dot(r_prev, r))) > params.orthogonality_threshold * rkp1:
                p = r.copy()
                r_prev = r.copy()
                continue
        r_prev = r.copy()
        beta = rkp1 / rk
        if not math.isfinite(beta):
            return CGResult(best_x, math.sqrt(best_sq), k + 1, NUMERICAL_BREAKDOWN)
        p = r + beta * p
    return CGResult(best_x, math.sqrt(best_sq), params.max_iter, MAX_ITERATIONS)


class SRSMatrix:
    """S v = (1/N_total) sum_i (O*_i . v - Obar . v) O*_i + diag_shift v with the samples resident in HBM."""

    def __init__(self, batch, ostar_mean_flat, total_samples, diag_shift=0.0, allreduce=None):
        self.batch, self.mean, self.n, self.diag_shift, self.allreduce = batch, ostar_mean_flat, total_samples, diag_shift, allreduce

    def __call__(self, v):
        mean_dot_v = float(np.dot(self.mean, v))
        out = self.batch.sr_matvec(v, mean_dot_v)
        if self.allreduce is not None:
            out = self.allreduce(out)
        out = out * (1.0 / float(self.n))
        if self.diag_shift != 0.0:
            out = out + self.diag_shift * v
        return out


def calculate_natural_gradient(batch, gradient_flat, ostar_mean_flat, total_samples, diag_shift, cg_params, init_guess=None,
                               allreduce=None):
    """Optimizer::CalculateNaturalGradient: solves (S + diag_shift) x = gradient; raises on indefinite / breakdown like
    the reference, returns the best iterate on non-convergence."""
    s = SRSMatrix(batch, ostar_mean_flat, total_samples, diag_shift, allreduce)
    x0 = np.zeros_like(gradient_flat) if init_guess is None else init_guess
    res = conjugate_gradient(s, gradient_flat, x0, cg_params)
    if res.reason in (INDEFINITE_MATRIX, NUMERICAL_BREAKDOWN):
        raise RuntimeError(f"CG solver terminated: {REASONS[res.reason]} iterations={res.iterations} residual_norm={res.residual_norm}")
    return res
