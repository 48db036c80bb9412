"""Runs the stand-alone truncation entry point (CAQR + block Jacobi) on synthetic Theta matrices; used under ncu."""
import argparse, ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from peps_b200 import _lib
ap = argparse.ArgumentParser()
ap.add_argument("--nr", type=int, default=512); ap.add_argument("--nc", type=int, default=512)
ap.add_argument("--t", type=int, default=64); ap.add_argument("--W", type=int, default=8)
ap.add_argument("--decay", type=float, default=0.93)
a = ap.parse_args()
lib = _lib.load()
rng = np.random.default_rng(0)
k = min(a.nr, a.nc)
th = np.empty((a.W, a.nr, a.nc))
for w in range(a.W):
    u, _ = np.linalg.qr(rng.standard_normal((a.nr, k)))
    v, _ = np.linalg.qr(rng.standard_normal((a.nc, k)))
    th[w] = (u * a.decay ** np.arange(k)) @ v.T
b = np.empty((a.W, a.t, a.nc)); kept = np.empty(a.W, dtype=np.int32); sw = C.c_int32()
dp = lambda x: x.ctypes.data_as(C.POINTER(C.c_double))
t0 = time.time()
rc = lib.peps_test_truncate(0, a.W, a.nr, a.nc, a.t, a.t, 0.0, dp(th), dp(b), kept.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(sw))
print("rc", rc, "sweeps", sw.value, "time", time.time() - t0)
