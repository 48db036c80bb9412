"""N>1 path on CPU: two gloo ranks, each with its own walkers (host-simulated device ops), must produce the same
energy and gradient as one rank holding all walkers -- the all-reduce of [sum O*, sum E_loc O*] and the all-gather of
energies replace the reference's MPI gathers (mc_energy_grad_evaluator.h:292-310)."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, pickle
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import torch.distributed as dist
import hostsim_lib
from oracle import vmc
from peps_b200.api import (BMPSTruncateParams, SplitIndexTPS, MCEnergyGradEvaluator, MonteCarloParams, Configuration,
                           SquareSpinOneHalfXXZModelOBC, MCUpdateSquareNNExchange)
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
from parity_common import complex_tps
tps = SplitIndexTPS(complex_tps(3, 3, 2, 4) if {cx!r} else vmc.random_tps(3, 3, 2, 2, seed=4))
W = 2
cfgs = np.stack([vmc.shuffled_half_filled_config(3, 3, 50 + rank * W + w) for w in range(W)])
mc = MonteCarloParams(num_samples=3 * W * world, num_warmup_sweeps=0, sweeps_between_samples=1, is_warmed_up=True)
ev = MCEnergyGradEvaluator(mc, BMPSTruncateParams.SVD(4, 4, 0.0), tps, SquareSpinOneHalfXXZModelOBC(1, 1, 0),
                           MCUpdateSquareNNExchange(9), walkers=W, configs=cfgs, lib=hostsim_lib.load(), dist=dist,
                           rank=rank, world_size=world)
res = ev.Evaluate(collect_sr_buffers=True)
from peps_b200 import sr
nat, iters, resid = ev.CalculateNaturalGradient(res, 1e-3, sr.ConjugateGradientParams(max_iter=200, relative_tolerance=1e-10))
if rank == 0:
    pickle.dump(dict(energy=res.energy, err=res.energy_error, grad=res.gradient.pack(), es=res.energy_samples,
                     nat=nat.pack(), iters=iters), open({out!r}, "wb"))
dist.destroy_process_group()
'''


import pytest


@pytest.mark.parametrize("cx", [False, True])
def test_two_rank_gloo_matches_single_rank(cx):
    """cx: a complex state -- planar accumulators / CG vectors through the same all-reduces."""
    import pickle
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import hostsim_lib
    from oracle import vmc
    from peps_b200.api import (BMPSTruncateParams, SplitIndexTPS, MCEnergyGradEvaluator, MonteCarloParams,
                               SquareSpinOneHalfXXZModelOBC, MCUpdateSquareNNExchange)
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "res.pkl")
        script = os.path.join(td, "worker.py")
        open(script, "w").write(WORKER.format(root=ROOT, out=out, cx=cx))
        env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29612" if cx else "29611", WORLD_SIZE="2")
        procs = [subprocess.Popen([sys.executable, script], env=dict(env, RANK=str(r))) for r in range(2)]
        for p in procs:
            assert p.wait(timeout=300) == 0
        two = pickle.load(open(out, "rb"))
    # single rank with the same four walkers (seeds 9..12, configurations 50..53)
    from parity_common import complex_tps
    tps = SplitIndexTPS(complex_tps(3, 3, 2, 4) if cx else vmc.random_tps(3, 3, 2, 2, seed=4))
    cfgs = np.stack([vmc.shuffled_half_filled_config(3, 3, 50 + w) for w in range(4)])
    mc = MonteCarloParams(num_samples=12, num_warmup_sweeps=0, sweeps_between_samples=1, is_warmed_up=True)
    ev = MCEnergyGradEvaluator(mc, BMPSTruncateParams.SVD(4, 4, 0.0), tps, SquareSpinOneHalfXXZModelOBC(1, 1, 0),
                               MCUpdateSquareNNExchange(9), walkers=4, configs=cfgs, lib=hostsim_lib.load())
    one = ev.Evaluate(collect_sr_buffers=True)
    from peps_b200 import sr
    nat1, it1, _ = ev.CalculateNaturalGradient(one, 1e-3, sr.ConjugateGradientParams(max_iter=200, relative_tolerance=1e-10))
    # SR across ranks: every rank holds half of the O* samples, the matvec output is all-reduced through the
    # peps_allreduce_fn callback on the engine's buffer (NCCL on the device pointer on GPUs)
    assert np.max(np.abs(two["nat"] - nat1.pack())) < 1e-8 * np.max(np.abs(nat1.pack()))
    assert abs(two["iters"] - it1) <= 1
    assert two["es"].shape == (4, 3)
    assert np.allclose(two["es"], one.energy_samples, rtol=0, atol=1e-12)
    assert abs(two["energy"] - one.energy) < 1e-12
    assert np.max(np.abs(two["grad"] - one.gradient.pack())) < 1e-12 * max(1.0, np.max(np.abs(one.gradient.pack())))
