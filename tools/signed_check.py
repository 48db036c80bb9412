"""CUDA vs oracle amplitude parity for the signed-TPS stress variant at a given size (development aid)."""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from peps_b200.api import BMPSTruncateParams, SplitIndexTPS, WalkerBatch
from oracle import vmc
ap = argparse.ArgumentParser()
ap.add_argument("--L", type=int, default=8); ap.add_argument("--D", type=int, default=6); ap.add_argument("--chi", type=int, default=36)
ap.add_argument("--defl", type=float, default=-1.0); ap.add_argument("--tol", type=float, default=1e-14); ap.add_argument("--maxsweeps", type=int, default=40)
a = ap.parse_args()
L, D, chi = a.L, a.D, a.chi
for signed in (False, True):
    tps = vmc.random_tps(L, L, 2, D, seed=20260101, signed=signed)
    W = 2
    cfgs = np.stack([vmc.shuffled_half_filled_config(L, L, 1000 + w) for w in range(W)])
    b = WalkerBatch(L, L, 2, D, W, BMPSTruncateParams.SVD(chi, chi, 0.0))
    if a.defl >= 0: b.set_deflation(a.defl)
    b.set_jacobi(a.tol, 1, a.maxsweeps)
    b.set_tps(SplitIndexTPS(tps)); b.set_configs(cfgs); b.init_walkers()
    amp = b.amplitudes()
    psi = [b.probe_trace_row(r) for r in (0, L // 2, L - 2)]
    ref = np.array([vmc.Walker(tps, cfgs[w], (chi, chi, 0.0)).amplitude for w in range(W)])
    print("signed" if signed else "pos", "amp", amp, "ref", ref, "rel", np.abs(amp / ref - 1), "jsweeps/jcalls", b.stat(3), b.stat(4),
          "closure spread", [float(np.max(np.abs(p / amp - 1))) for p in psi], flush=True)
