"""Ad-hoc timing of the complex path (development aid)."""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from peps_b200.api import BMPSTruncateParams, SplitIndexTPS, WalkerBatch
from oracle import vmc
from parity_common import complex_tps
ap = argparse.ArgumentParser()
ap.add_argument("--L", type=int, default=8); ap.add_argument("--D", type=int, default=6)
ap.add_argument("--chi", type=int, default=36); ap.add_argument("--walkers", type=int, default=37)
a = ap.parse_args()
L, D, chi, W = a.L, a.D, a.chi, a.walkers
tps = complex_tps(L, L, D, 20260101)
cfgs = np.stack([vmc.shuffled_half_filled_config(L, L, 1000 + w) for w in range(W)])
b = WalkerBatch(L, L, 2, D, W, BMPSTruncateParams.SVD(chi, chi, 0.0))
b.set_complex(); b.set_tps(SplitIndexTPS(tps)); b.set_configs(cfgs); b.seed_rng(np.arange(W) + 7); b.profile_enable(True)
def tm(f, name):
    b.sync(); t = time.time(); r = f(); b.sync(); dt = time.time() - t
    pr = b.profile_get(True)
    print("   " + "  ".join(f"{k}: {v['ms']:.0f} ms / {v['launches']}" for k, v in pr.items()))
    print(f"{name}: {dt:.3f} s rows_in={b.stat(8)} rows_kept={b.stat(9)} chain_rows={b.stat(13)}/{b.stat(12)}", flush=True)
    return r
tm(b.init_walkers, "init_walkers")
print("amp", b.amplitudes_c()[:2])
tm(lambda: b.sweep(1), "sweep")
tm(lambda: b.energy_and_holes(True), "energy+holes")
print("eloc", b.eloc_c()[:2], "samples/s", W / 1.0)
