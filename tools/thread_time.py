"""Times T host threads, each driving its own WalkerBatch (own CUDA stream) on the same GPU."""
import argparse, os, sys, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from peps_b200.api import BMPSTruncateParams, SplitIndexTPS, WalkerBatch
from oracle import vmc
ap = argparse.ArgumentParser()
ap.add_argument("--L", type=int, default=10); ap.add_argument("--D", type=int, default=8); ap.add_argument("--chi", type=int, default=64)
ap.add_argument("--walkers", type=int, default=64); ap.add_argument("--threads", type=int, default=2); ap.add_argument("--steps", type=int, default=2)
a = ap.parse_args()
L, D, chi = a.L, a.D, a.chi
tps = SplitIndexTPS(vmc.random_tps(L, L, 2, D, seed=20260101))
Wt = a.walkers // a.threads
res = [None] * a.threads
bar = threading.Barrier(a.threads + 1)
def work(i):
    cfgs = np.stack([vmc.shuffled_half_filled_config(L, L, 1000 + i * Wt + w) for w in range(Wt)])
    b = WalkerBatch(L, L, 2, D, Wt, BMPSTruncateParams.SVD(chi, chi, 0.0))
    b.set_tps(tps); b.set_configs(cfgs); b.seed_rng(np.arange(Wt) + 7 + i * Wt); b.init_walkers(); b.normalize_state_order1()
    b.zero_accumulators(); b.sample(1)
    bar.wait()
    for _ in range(a.steps):
        e, _ = b.sample(1)
    b.sync()
    bar.wait()
    res[i] = float(np.mean(e))
ths = [threading.Thread(target=work, args=(i,)) for i in range(a.threads)]
[t.start() for t in ths]
bar.wait(); t0 = time.time(); bar.wait(); dt = time.time() - t0
[t.join() for t in ths]
print(f"threads={a.threads} walkers={a.walkers} steps={a.steps}: {dt:.2f} s -> {a.walkers * a.steps / dt:.2f} samples/s  mean_eloc={res}")
