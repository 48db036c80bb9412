"""Ad-hoc timing of the hot path at a given size (development aid)."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from peps_b200.api import BMPSTruncateParams, SplitIndexTPS, WalkerBatch
from oracle import vmc

ap = argparse.ArgumentParser()
ap.add_argument("--L", type=int, default=10)
ap.add_argument("--D", type=int, default=8)
ap.add_argument("--chi", type=int, default=64)
ap.add_argument("--walkers", type=int, default=8)
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--inner", type=int, default=1)
ap.add_argument("--profile", type=int, default=1)
ap.add_argument("--j2", type=float, default=0.0)
a = ap.parse_args()
L, D, chi, W = a.L, a.D, a.chi, a.walkers
tps = vmc.random_tps(L, L, 2, D, seed=20260101)
cfgs = np.stack([vmc.shuffled_half_filled_config(L, L, 1000 + w) for w in range(W)])
b = WalkerBatch(L, L, 2, D, W, BMPSTruncateParams.SVD(chi, chi, 0.0))
b.set_jacobi(1e-14, a.inner, 60); b.profile_enable(bool(a.profile))
b.set_tps(SplitIndexTPS(tps)); b.set_configs(cfgs); b.seed_rng(np.arange(W) + 7)
if a.j2 != 0.0:
    from peps_b200.api import SquareSpinOneHalfJ1J2XXZModelOBC
    b.set_model(SquareSpinOneHalfJ1J2XXZModelOBC(1.0, 1.0, a.j2, a.j2, 0.0))
def tm(f, name):
    b.sync(); t = time.time(); r = f(); b.sync(); dt = time.time() - t
    if a.profile:
        pr = b.profile_get(True)
        print("   " + "  ".join(f"{k}: {v['ms']:.0f} ms / {v['launches']} launches / {v['flops']/max(v['ms'],1e-9)/1e9:.2f} TF/s" for k, v in pr.items()))
    print(f"{name}: {dt:.3f} s  stats absorb={b.stat(0)} bten={b.stat(1)} trace={b.stat(2)} jsweeps={b.stat(3)} jcalls={b.stat(4)} qr={b.stat(5)} launches={b.stat(6)} rows_in={b.stat(8)} rows_kept={b.stat(9)} jrounds={b.stat(10)} chain_rows={b.stat(13)}/{b.stat(12)} pool={b.stat(7)/2**30:.2f} GiB", flush=True)
    return r
tm(b.init_walkers, "init_walkers")
amp = b.amplitudes(); print("amp", amp[:4])
f = b.normalize_state_order1(); print("site factor", f)
for r in range(a.reps):
    acc = tm(lambda: b.sweep(1), "sweep")
    print("accept", acc[:4])
    e = tm(lambda: b.energy_and_holes(True), "energy+holes")
    print("eloc", e[:4])
