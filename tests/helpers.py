"""Shared helpers for the test-suite (oracle-side; never imported by peps_b200)."""
import itertools
import math
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden_tps(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    rows, cols, phys = int(z["rows"]), int(z["cols"]), int(z["phys"])
    tps = [[[z[f"t_{r}_{c}_{s}"] for s in range(phys)] for c in range(cols)] for r in range(rows)]
    return tps, z


def ising_tn(L, beta, cols=None):
    """OBC Ising partition function as a 2D tensor network: one Boltzmann matrix per bond, placed on
    the L and U legs of each site (same network as tests/test_2d_tn/test_bmps_contractor.cpp:153-270
    of the reference up to the gauge of where the bond matrices sit). L rows x cols columns (square by default)."""
    B = np.array([[math.exp(beta), math.exp(-beta)], [math.exp(-beta), math.exp(beta)]])
    rows, cols = L, (L if cols is None else cols)
    tn = []
    for r in range(rows):
        row = []
        for c in range(cols):
            dl, dd = (1 if c == 0 else 2), (1 if r == rows - 1 else 2)
            dr, du = (1 if c == cols - 1 else 2), (1 if r == 0 else 2)
            T = np.zeros((dl, dd, dr, du))
            for s in range(2):
                for l in range(dl):
                    for u in range(du):
                        wl = B[l, s] if c > 0 else 1.0
                        wu = B[u, s] if r > 0 else 1.0
                        T[l, s if dd == 2 else 0, s if dr == 2 else 0, u] += wl * wu
            row.append(T)
        tn.append(row)
    return tn


def ising_exact_logZ(L, beta, length=None):
    """Exact OBC partition function by transfer matrix (reference: test_bmps_contractor.cpp:27-126): strips of width
    L, `length` of them (square by default)."""
    n = 1 << L
    length = L if length is None else length

    def chain(cfg):
        return sum(1 if ((cfg >> i) & 1) == ((cfg >> (i + 1)) & 1) else -1 for i in range(L - 1))

    ch = np.array([chain(c) for c in range(n)], dtype=float)
    idx = np.arange(n)
    pop = np.array([bin(v).count("1") for v in range(n)])
    lad = L - 2 * pop[idx[:, None] ^ idx[None, :]]
    Tm = np.exp(beta * (lad + 0.5 * ch[:, None] + 0.5 * ch[None, :]))
    b = np.exp(beta * 0.5 * ch)
    v, logscale = b.copy(), 0.0
    for _ in range(length - 1):
        v = v @ Tm
        s = v.max()
        v /= s
        logscale += math.log(s)
    return math.log(v @ b) + logscale


def exact_summation(tps, model, trunc=(1, 1000, 0.0), all_configs=False):
    """ExactSumEnergyEvaluatorMPI restated on top of the oracle walker
    (algorithm/vmc_update/exact_summation_energy_evaluator.h:190-295)."""
    from oracle.vmc import Walker, tps_like_zeros
    rows, cols = len(tps), len(tps[0])
    s_o, s_eo, wsum, esum = tps_like_zeros(tps), tps_like_zeros(tps), 0.0, 0.0
    for bits in itertools.product([0, 1], repeat=rows * cols):
        if not all_configs and sum(bits) != rows * cols // 2:
            continue
        cfg = np.array(bits).reshape(rows, cols)
        w = Walker(tps, cfg, trunc)
        e, holes, _ = model.energy_and_holes(tps, w, True)
        wt = abs(w.amplitude) ** 2
        wsum += wt
        esum += wt * e
        for r in range(rows):
            for c in range(cols):
                b = cfg[r, c]
                inc = w.amplitude * holes[r][c]
                s_o[r][c][b] = s_o[r][c][b] + inc
                s_eo[r][c][b] = s_eo[r][c][b] + np.conj(e) * inc
    energy = esum / wsum
    grad = [[[(s_eo[r][c][s] - np.conj(energy) * s_o[r][c][s]) / wsum for s in range(len(tps[r][c]))]
             for c in range(cols)] for r in range(rows)]
    return energy, grad


def grad_norm_square(g):
    return sum(float(np.sum(np.abs(t) ** 2)) for row in g for site in row for t in site)


def weighted_probe(g, complex_):
    """WeightedProbeInnerProduct (tests/test_algorithm/test_exact_summation_evaluator.cpp:49-70)."""
    tot = 0.0
    for r, row in enumerate(g):
        for c, site in enumerate(row):
            for i, t in enumerate(site):
                base = 0.012 * ((r + 1) * 11 + (c + 1) * 5 + (i + 1) * 2)
                if complex_:
                    base = complex(base, 0.0025 * ((r + 1) + (i + 1)))
                tot = tot + np.sum(np.conj(t) * (t * base))
    return tot
