#!/usr/bin/env python
"""bench.py -- VMC samples/s of the walker-batched sampling hot path (BASELINE.json metric).

One "step" = one VMC sample for every walker on every GPU: `sweeps_between_samples` NN-exchange Metropolis sweeps +
CalEnergyAndHoles<true> + O* accumulation (the loop body of the reference's MCEnergyGradEvaluator::Evaluate,
algorithm/vmc_update/mc_energy_grad_evaluator.h:205-282). Workload: BASELINE.json's headline configuration,
10x10 Heisenberg, D=8, chi=64 (Dmin=Dmax=chi, trunc_err=0), synthetic random positive TPS.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--walkers W_per_gpu] [--impl ours|reference]

Prints ONE JSON line on rank 0. `value` is timed with CUDA events on the library's stream with the TPS, the
configurations and the RNG state resident in HBM; `e2e` goes through the public evaluator-style call with host
buffers (TPS upload from pinned memory, walker refresh, sample, download of energies and both accumulators).
The `--impl reference` arm times the oracle port of the reference's CPU algorithm (oracle/) on the host cores.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (L, D, chi)
    "heisenberg_10x10_D8_chi64": (10, 8, 64),
    "heisenberg_8x8_D6_chi36": (8, 6, 36),
    "heisenberg_4x4_D4_chi8": (4, 4, 8),
}
TPS_SEED = 20260101
CFG_SEED0 = 1000
RNG_SEED0 = 7

# algorithmic flop model per sample per walker (factorisation counts are
# LAPACK model counts, not executed counts)
MODEL_FLOPS = {"heisenberg_10x10_D8_chi64": dict(gemm=2.96e11, qr=1.18e12, svd=9.80e11, total=2.45e12),
               "heisenberg_8x8_D6_chi36": dict(gemm=1.08e10, qr=2.88e10, svd=3.18e10, total=7.1e10),
               "heisenberg_4x4_D4_chi8": dict(gemm=2.43e6, qr=1.23e6, svd=1.88e6, total=5.5e6)}


SIGNED_TPS = False


def make_inputs(L, D, n_walkers, first_walker):
    from oracle import vmc
    tps = vmc.random_tps(L, L, 2, D, seed=TPS_SEED, signed=SIGNED_TPS)
    cfgs = np.stack([vmc.shuffled_half_filled_config(L, L, CFG_SEED0 + first_walker + w) for w in range(n_walkers)])
    seeds = np.arange(RNG_SEED0 + first_walker, RNG_SEED0 + first_walker + n_walkers, dtype=np.uint32)
    return tps, cfgs, seeds


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm on the host cores
# ---------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """One process == one Markov chain with one BLAS thread (the reference runs one chain per MPI rank with
    hp_numeric::SetTensorManipulationThreads(1)). Times a bounded slice of one sample and extrapolates by the
    reference's own operation counts: a sample is 8 full boundary-MPS growths (4 in the sweep, 4 in E_loc) plus
    2 x 2L row/column passes of BTen / trace / hole work."""
    L, D, chi, walker, n_rows = args
    try:
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(limits=1)
    except Exception:
        limiter = None
    from oracle import vmc
    from oracle.bmps import LEFT, RIGHT, UP, HORIZONTAL
    tps, cfgs, _ = make_inputs(L, D, 1, walker)
    t0 = time.time()
    w = vmc.Walker.__new__(vmc.Walker)
    w.config = np.array(cfgs[0], dtype=np.int64)
    w.rows = w.cols = L
    w.trunc = (chi, chi, 0.0)
    w.tn = vmc.project(tps, w.config)
    from oracle.contractor import BMPSContractor
    w.contractor = BMPSContractor(L, L)
    w.contractor.init(w.tn)
    w.contractor.set_truncate_params(chi, chi, 0.0)
    c = w.contractor
    t0 = time.time()
    c.generate_bmps_approach(w.tn, UP)             # one full growth of the DOWN stack: (L-1) MultiplyMPO
    t_grow = time.time() - t0
    t1 = time.time()
    rows_done = 0
    model = vmc.XXZModel()
    for row in range(min(n_rows, L)):              # E_loc-style row passes (BTen growth, traces, holes)
        c.init_bten(w.tn, LEFT, row)
        c.grow_full_bten(w.tn, RIGHT, row, 1, True)
        psi = c.trace(w.tn, (row, 0), HORIZONTAL)
        for col in range(L):
            c.punch_hole(w.tn, (row, col), HORIZONTAL)
            if col < L - 1:
                s1, s2 = (row, col), (row, col + 1)
                model.bond_energy(s1, s2, int(w.config[s1]), int(w.config[s2]), HORIZONTAL, w, tps, 1.0 / psi)
                c.shift_bten_window(w.tn, RIGHT)
        rows_done += 1
        if row < L - 1:
            break                                  # later rows need shifted BMPS windows; one row is the slice
    t_row = (time.time() - t1) / max(rows_done, 1)
    t_sample = 8.0 * t_grow + 4.0 * L * t_row      # 2L passes with holes-class work + 2L lighter passes, bounded above
    return dict(t_grow=t_grow, t_row=t_row, t_sample=t_sample)


def cpu_samples_per_s(L, D, chi, cores):
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(L, D, chi, 100000 + i, 1) for i in range(cores)])
    t_sample = float(np.mean([r["t_sample"] for r in res]))
    return cores / t_sample, dict(t_grow=float(np.mean([r["t_grow"] for r in res])),
                                  t_row=float(np.mean([r["t_row"] for r in res])), t_sample=t_sample)


def run_reference_arm(args, rank):
    if rank != 0:
        return
    L, D, chi = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    vals = []
    detail = None
    for _ in range(max(1, min(args.steps, 2))):
        v, detail = cpu_samples_per_s(L, D, chi, cores)
        vals.append(v)
    value = float(np.mean(vals))
    sample = ("per process: one full boundary-MPS growth ((L-1) MultiplyMPO) + one row pass (BTen growth, traces, holes); "
              "sample time = 8 growths + 4L row passes; one chain and one BLAS thread per core")
    line = {"impl": "reference", "metric": "vmc_samples_per_s", "value": value, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / value if value > 0 else None,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "lattice": f"{L}x{L}", "D": D, "chi": chi,
                       "trunc": "Dmin=Dmax=chi, trunc_err=0", "model": "Heisenberg NN (XXZ jz=jxy=1)"},
            "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample,
                             "detail": detail},
            "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------
# clocks sampler
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.lines, self.proc, self.device = [], None, device

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measure_fp64_peak(torch, n=4096, reps=6):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    (a @ b); torch.cuda.synchronize()
    best = 0.0
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); (a @ b); e.record(); torch.cuda.synchronize()
        best = max(best, 2 * n ** 3 / (s.elapsed_time(e) * 1e-3) / 1e12)
    del a, b
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--walkers", type=int, default=148, help="walkers (Markov chains) per GPU")
    ap.add_argument("--j2", type=float, default=0.0, help="next-nearest-neighbour coupling (J1-J2 model, BASELINE config #3); 0 = NN Heisenberg headline")
    ap.add_argument("--workload", default="heisenberg_10x10_D8_chi64", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--signed", action="store_true", help="uniform [-1,1) TPS entries (cancellation stress variant of SURVEY.md 8d.1)")
    ap.add_argument("--e2e-steps", type=int, default=1)
    ap.add_argument("--streams", type=int, default=4, help="host threads / CUDA streams sharing the walkers of a GPU")
    args = ap.parse_args()

    global SIGNED_TPS
    SIGNED_TPS = args.signed
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from peps_b200.api import BMPSTruncateParams, SplitIndexTPS, WalkerBatch
    import queue
    L, D, chi = WORKLOADS[args.workload]
    W = args.walkers
    S = max(1, min(args.streams, W))
    while W % S:
        S -= 1
    Ws = W // S
    tps, cfgs, seeds = make_inputs(L, D, W, rank * W)
    sit = SplitIndexTPS(tps)
    dev = torch.device("cuda", local_rank)

    # S host threads, each owning one context (= one CUDA stream) with W/S walkers: kernels of different contexts
    # overlap on the GPU and fill the tails of partially occupied launches.
    class Lane(threading.Thread):
        def __init__(self, i):
            super().__init__(daemon=True)
            self.i, self.q, self.r, self.b = i, queue.Queue(), queue.Queue(), None
            self.start()

        def run(self):
            while True:
                fn = self.q.get()
                if fn is None:
                    return
                try:
                    self.r.put(("ok", fn(self)))
                except Exception as exc:      # surface worker failures on the main thread
                    self.r.put(("err", exc))

    lanes = [Lane(i) for i in range(S)]

    def on_all(fn):
        for ln in lanes:
            ln.q.put(fn)
        out = []
        for ln in lanes:
            st, v = ln.r.get()
            if st == "err":
                raise v
            out.append(v)
        return out

    def setup(ln):
        sl = slice(ln.i * Ws, (ln.i + 1) * Ws)
        ln.b = WalkerBatch(L, L, 2, D, Ws, BMPSTruncateParams.SVD(chi, chi, 0.0), device=local_rank)
        ln.b.set_tps(sit)
        if args.j2 != 0.0:
            from peps_b200.api import SquareSpinOneHalfJ1J2XXZModelOBC
            ln.b.set_model(SquareSpinOneHalfJ1J2XXZModelOBC(1.0, 1.0, args.j2, args.j2, 0.0))
        ln.b.set_configs(cfgs[sl])
        ln.b.seed_rng(seeds[sl])
        ln.b.init_walkers()
        return float(np.max(np.abs(ln.b.amplitudes())))

    mx = max(on_all(setup))
    if world > 1:
        t = torch.tensor([mx], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        mx = float(t.item())
    on_all(lambda ln: ln.b.normalize_state_order1(mx))     # MonteCarloEngine::NormalizeStateOrder1

    def barrier():
        on_all(lambda ln: ln.b.sync())
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # accumulators as torch views for the NCCL all-reduce of [sum O*, sum E_loc O*]
    n_par = lanes[0].b.tps_size

    class _Cai:
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3}

    views = []
    for ln in lanes:
        p_o, p_eo = ln.b.accumulator_device_ptrs()
        views.append((torch.as_tensor(_Cai(p_o, n_par), device=dev), torch.as_tensor(_Cai(p_eo, n_par), device=dev)))

    on_all(lambda ln: ln.b.zero_accumulators())
    for _ in range(args.warmup):
        on_all(lambda ln: ln.b.sample(1))
    barrier()
    launches0 = sum(on_all(lambda ln: ln.b.stat(6)))
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()                                      # device idle on both sides of the timed region

    def timed(ln):
        last = None
        for _ in range(args.steps):
            last, _ = ln.b.sample(1)
        ln.b.sync()
        return last

    energies = np.concatenate(on_all(timed))
    if world > 1:                                     # gradient reduction of the iteration (NCCL over NVLink)
        acc_o = torch.stack([v[0] for v in views]).sum(0)
        acc_eo = torch.stack([v[1] for v in views]).sum(0)
        dist.all_reduce(acc_o)
        dist.all_reduce(acc_eo)
    torch.cuda.synchronize()
    ev1.record()
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    clk = clocks.stop() if rank == 0 else None
    launches = sum(on_all(lambda ln: ln.b.stat(6))) - launches0
    # one more step of the same loop with every launch bracketed by a CUDA-event pair on the launching stream:
    # per-kernel-class device time and useful flops for the roofline (kept out of `value`: the 2 x ~100k event
    # records per step cost about 10 % of a step)

    def profiled(ln):
        ln.b.profile_enable(True)
        ln.b.profile_get(True)
        ln.b.sample(1)
        pr = ln.b.profile_get(True)
        ln.b.profile_enable(False)
        return pr

    prs = []
    for ln in lanes:                                   # one lane at a time: per-launch durations free of overlap
        ln.q.put(profiled)
        st, v = ln.r.get()
        if st == "err":
            raise v
        prs.append(v)
    prof = {k: {f: sum(p[k][f] for p in prs) for f in ("ms", "launches", "flops")} for k in prs[0]}
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    total_samples = W * world * args.steps
    value = total_samples / (elapsed_ms * 1e-3)

    # ---- end-to-end through the evaluator-style public call with host buffers
    flat = torch.from_numpy(sit.pack()).pin_memory()
    h2d = flat.numel() * 8 * S
    d2h = (2 * n_par * 8 + 2 * Ws * 8) * S

    def e2e_step(ln):
        for _ in range(args.e2e_steps):
            ln.b.set_tps(flat.numpy())                # state fan-out (mc_energy_grad_evaluator.h:161)
            ln.b.init_walkers()                       # RefreshWavefunctionComponent (:164)
            ln.b.zero_accumulators()
            e, acc = ln.b.sample(1)
            osum, eosum = ln.b.accumulators()
        return None

    barrier()
    t0 = time.perf_counter()
    on_all(e2e_step)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = W * world * args.e2e_steps / e2e_s if args.e2e_steps else 0.0

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel
    dom = max(prof.items(), key=lambda kv: kv[1]["ms"])
    dom_name, dom_p = dom
    fp64_peak = measure_fp64_peak(torch)
    achieved = dom_p["flops"] / max(dom_p["ms"], 1e-9) / 1e9            # TFLOP/s of useful FP64 work in that kernel class
    step_ms = elapsed_ms / args.steps
    mf = MODEL_FLOPS[args.workload]
    # dram__bytes_read + write of ONE captured launch (ncu --set full, W=16; profiles/r1_ncu_*_final.txt / _k2.txt). For the
    # trailing update that launch (1664 CTAs: 16 walkers x 8 row blocks x 13 column tiles) moves 436 MB algorithmically
    # (C tile in and out) + 17 MB of reflectors: measured 399 MB, no wasted re-reads.
    traffic = {"jacobi_round": 2.0e6, "apply_reflector": 399.3e6, "panel_qr": 1.4e6}.get(dom_name)
    roofline = {"kernel": dom_name, "bound": "tensor", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": achieved / fp64_peak,
                "peak_source": "cuBLAS DGEMM 4096^3 through torch.matmul, measured in this run (MEASURED_PEAKS.json carries no FP64 figure)",
                "traffic": traffic, "measured_on": "one extra step of the timed loop with per-launch CUDA-event pairs (see bench.py)",
                "launches": dom_p["launches"], "avg_launch_us": 1e3 * dom_p["ms"] / max(dom_p["launches"], 1),
                "share_of_step": dom_p["ms"] / max(sum(v["ms"] for v in prof.values()), 1e-9),
                "per_class_ms": {k: round(v["ms"], 1) for k, v in prof.items()},
                "per_class_tflops": {k: (v["flops"] / max(v["ms"], 1e-9) / 1e9) for k, v in prof.items()},
                "whole_path": {"model_flops_per_sample": mf["total"], "model_tflops": value * mf["total"] / world / 1e12,
                               "frac_of_fp64_peak": value * mf["total"] / world / 1e12 / fp64_peak,
                               "contraction_model_tflops": value * mf["gemm"] / world / 1e12}}

    cpu = None
    if not args.no_cpu_baseline and world == 1:      # the CPU arm is timed at N=1 only (rank 0, all host cores)
        cores = os.cpu_count() or 1
        v, detail = cpu_samples_per_s(L, D, chi, cores)
        cpu = {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
               "sample": ("oracle port, one chain + one BLAS thread per core: one full boundary-MPS growth + one row pass timed "
                          "per process, sample = 8 growths + 4L row passes"), "detail": detail}

    line = {"metric": "vmc_samples_per_s", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "lattice": f"{L}x{L}", "D": D, "chi": chi, "walkers_per_gpu": W, "streams_per_gpu": S,
                       "trunc": "Dmin=Dmax=chi, trunc_err=0",
                       "model": "Heisenberg NN (XXZ jz=jxy=1)" if args.j2 == 0.0 else f"J1-J2 Heisenberg (j2={args.j2})",
                       "sweeps_between_samples": 1, "tps": ("uniform[-1,1)" if args.signed else "uniform[0,1)") + f" seed {TPS_SEED}, NormalizeAllSite + order-1 rescale",
                       "l2": "per-step working set (walkers x ~60 MB of BMPS stacks + scratch) exceeds the 126 MB L2",
                       "parallelism": f"walkers sharded over {world} GPU(s); NCCL all-reduce of the two accumulators",
                       "algorithm": "reference call sequence; exact boundary-MPS memo (6(L-1) absorptions per sample instead of "
                                    "8(L-1), bit-identical); R-only QR chain with rows below 1e-13 of the largest dropped; "
                                    "column-sorted preconditioning QR + block Jacobi truncation (DESIGN.md section 2)"},
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": args.e2e_steps, "call": "set_tps + init_walkers + sample + accumulators (Evaluate with 1 sample per walker)"},
            "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "cpu_baseline": cpu,
            "mean_eloc": float(np.mean(energies))}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
