"""Ad-hoc per-class timing of the fermion path at BASELINE config #4 sizes (development aid)."""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from peps_b200.api import BMPSTruncateParams, FermionSplitIndexTPS, TableModel, WalkerBatch
from oracle import vmc
ap = argparse.ArgumentParser()
ap.add_argument("--L", type=int, default=8); ap.add_argument("--D", type=int, default=8)
ap.add_argument("--chi", type=int, default=64); ap.add_argument("--walkers", type=int, default=74)
a = ap.parse_args()
L, D, chi, W = a.L, a.D, a.chi, a.walkers
ftps = FermionSplitIndexTPS.random(L, L, D, 20260101)
cfgs = np.stack([vmc.shuffled_half_filled_config(L, L, 1000 + w) for w in range(W)])
b = WalkerBatch(L, L, 2, D, W, BMPSTruncateParams.SVD(chi, chi, 0.0))
b.set_fermion(ftps); b.set_tps(ftps); b.set_model(TableModel.spinless_fermion(1.0, 0.0, 0.0))
b.set_configs(cfgs); b.seed_rng(np.arange(W) + 7); b.profile_enable(True)
def tm(f, name):
    b.sync(); t = time.time(); r = f(); b.sync(); dt = time.time() - t
    pr = b.profile_get(True)
    print("   " + "  ".join(f"{k}: {v['ms']:.0f} ms / {v['launches']} / {v['flops']/max(v['ms'],1e-9)/1e9:.2f} TF/s" for k, v in pr.items()))
    print(f"{name}: {dt:.3f} s rows_in={b.stat(8)} rows_kept={b.stat(9)} chain_rows={b.stat(13)}/{b.stat(12)} jcalls={b.stat(4)} small_svd={b.stat(14)} jrounds={b.stat(10)}", flush=True)
    return r
tm(b.init_walkers, "init_walkers")
b.normalize_state_order1()
tm(lambda: b.sweep(1), "sweep")
tm(lambda: b.energy_and_holes(True), "energy+holes")
