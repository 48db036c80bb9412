"""Runs the stand-alone batched einsum entry point on one contraction (used under ncu)."""
import argparse, ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from peps_b200 import _lib
ap = argparse.ArgumentParser()
ap.add_argument("--spec", default="kea,eaoj->koj")
ap.add_argument("--da", default="512,8,64"); ap.add_argument("--db", default="8,64,8,64")
ap.add_argument("--W", type=int, default=32); ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
lib = _lib.load()
da = np.array([int(x) for x in a.da.split(",")], dtype=np.int32); db = np.array([int(x) for x in a.db.split(",")], dtype=np.int32)
rng = np.random.default_rng(0)
A = rng.standard_normal((a.W,) + tuple(da)); B = rng.standard_normal((a.W,) + tuple(db))
ins, out = a.spec.split("->"); la, lb = ins.split(",")
ref = np.einsum(f"w{la},w{lb}->w{out}", A, B)
Cc = np.empty(ref.shape)
dp = lambda x: x.ctypes.data_as(C.POINTER(C.c_double)); ip = lambda x: x.ctypes.data_as(C.POINTER(C.c_int32))
for _ in range(a.reps):
    t0 = time.time()
    rc = lib.peps_test_einsum(0, a.W, a.spec.encode(), ip(da), len(da), ip(db), len(db), dp(A), dp(B), dp(Cc))
    print("rc", rc, "err", np.max(np.abs(Cc - ref)), "time", time.time() - t0)
