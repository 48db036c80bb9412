"""SR restatement for the oracle: dense O* vectors, dense S, the reference's CG.  Test infrastructure.

  * SRSMatrix::operator*      optimizer/stochastic_reconfiguration_smatrix.h:45-91
  * ConjugateGradientSolver   utility/conjugate_gradient_solver.h:181-276
"""
import math
import numpy as np


def dense_ostar(sample, tps):
    """Flattens one O* sample {(r,c): (phys index, tensor)} into the packed-TPS layout (zeros elsewhere)."""
    parts = []
    for r, row in enumerate(tps):
        for c, site in enumerate(row):
            b, t = sample[(r, c)]
            for s in range(len(site)):
                parts.append(t.ravel() if s == b else np.zeros(site[s].size))
    return np.concatenate(parts)


def s_matvec(ostars, obar, v, diag_shift):
    mean_dot_v = np.dot(obar, v)
    res = np.zeros_like(v)
    for o in ostars:                                   # :60-65
        res = res + (np.dot(o, v) - mean_dot_v) * o
    res = res * (1.0 / len(ostars))                    # :66
    return res + diag_shift * v                        # :86-88


def cg(matvec, b, x0, max_iter=100, rel_tol=1e-4, abs_tol=0.0, recompute=20, ortho=0.5):
    rhs = float(b @ b)
    tol_sq = max(rel_tol * rel_tol * rhs, abs_tol * abs_tol)
    r = b - matvec(x0)
    rr = float(r @ r)
    if rr <= tol_sq:
        return x0, math.sqrt(rr), 0
    p, x, best_x, best = r.copy(), x0.copy(), x0.copy(), rr
    r_prev, rkp1, stag = r.copy(), rr, 0
    eps = np.finfo(float).eps
    for k in range(max_iter):
        rk = rkp1
        ap = matvec(p)
        pap = float(p @ ap)
        if not (math.isfinite(pap) and pap > 0):
            return best_x, math.sqrt(best), k
        alpha = rk / pap
        x = x + alpha * p
        if alpha * alpha * float(p @ p) < eps * eps * float(x @ x):
            stag += 1
            if stag >= 3:
                return best_x, math.sqrt(best), k + 1
        else:
            stag = 0
        r = b - matvec(x) if (recompute > 0 and k % recompute == recompute - 1) else r - alpha * ap
        rkp1 = float(r @ r)
        if rkp1 < best:
            best_x, best = x.copy(), rkp1
        if rkp1 <= tol_sq:
            return x, math.sqrt(rkp1), k + 1
        if k > 0 and abs(float(r_prev @ r)) > ortho * rkp1:
            p, r_prev = r.copy(), r.copy()
            continue
        r_prev = r.copy()
        p = r + (rkp1 / rk) * p
    return best_x, math.sqrt(best), max_iter
