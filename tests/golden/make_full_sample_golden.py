"""Generates tests/golden/full_sample_10x10_D8_chi64.npz: ONE full VMC sample (MC sweep + CalEnergyAndHoles<true>) of
the oracle at BASELINE's headline configuration, which the GPU parity test then compares the CUDA path against
without re-running 3-4 minutes of numpy per walker on the GPU box.

    python tests/golden/make_full_sample_golden.py            # ~6 min on 3 cores

Inputs are the bench's own synthetic inputs (bench.py: TPS seed 20260101, configurations shuffled_half_filled_config(
1000 + w), updater seeds 7 + w). Two walkers with the NN Heisenberg model and one walker with the J1-J2 model (j2 =
0.5, BASELINE config #3). Stored per walker: the initial and the swept configuration, the acceptance count, the
amplitudes before / after the sweep, E_loc, the psi list of the energy solver, and of the 278 784 hole elements every
13th one, plus per-site <hole, site tensor> (= psi), per-site hole norms and 8 fixed random projections of the holes
(the full tensors would be 2.2 MB per walker).
"""
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

L, D, CHI = 10, 8, 64
TPS_SEED, CFG_SEED0, RNG_SEED0 = 20260101, 1000, 7
STRIDE = 13
CASES = [("nn", 0, 0.0), ("nn", 1, 0.0), ("j1j2", 0, 0.5)]


def hole_summaries(flat, nproj=8):
    rng = np.random.default_rng(4242)
    proj = np.array([float(np.dot(rng.standard_normal(flat.size), flat)) for _ in range(nproj)])
    return flat[::STRIDE].copy(), proj


def run(case):
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=2)
    except Exception:
        pass
    from oracle import vmc
    name, w, j2 = case
    tps = vmc.random_tps(L, L, 2, D, seed=TPS_SEED)
    cfg0 = vmc.shuffled_half_filled_config(L, L, CFG_SEED0 + w)
    wk = vmc.Walker(tps, cfg0, (CHI, CHI, 0.0))
    amp0 = wk.amplitude
    up = vmc.NNExchangeUpdater(RNG_SEED0 + w)
    acc = up.sweep(tps, wk)[0]
    model = vmc.XXZModel(1.0, 1.0, 0.0, j2, j2)
    e, holes, psi = model.energy_and_holes(tps, wk, True)
    flat = np.concatenate([holes[r][c].ravel() for r in range(L) for c in range(L)])
    sub, proj = hole_summaries(flat)
    dots = np.array([float(np.sum(holes[r][c] * tps[r][c][int(wk.config[r, c])])) for r in range(L) for c in range(L)])
    norms = np.array([float(np.linalg.norm(holes[r][c])) for r in range(L) for c in range(L)])
    return dict(name=name, w=w, j2=j2, cfg0=np.array(cfg0), cfg1=np.array(wk.config), accept=acc, amp0=amp0,
                amp1=wk.amplitude, eloc=e, psi=np.array(psi), hole_sub=sub, hole_proj=proj, hole_dots=dots, hole_norms=norms)


if __name__ == "__main__":
    with mp.get_context("spawn").Pool(len(CASES)) as pool:
        res = pool.map(run, CASES)
    out = {"L": L, "D": D, "chi": CHI, "tps_seed": TPS_SEED, "stride": STRIDE, "ncases": len(res)}
    for i, r in enumerate(res):
        for k, v in r.items():
            out[f"c{i}_{k}"] = v
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "full_sample_10x10_D8_chi64.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: (v["eloc"], v["accept"]) for k, v in zip(range(len(res)), res)})
