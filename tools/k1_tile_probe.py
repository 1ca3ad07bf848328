#!/usr/bin/env python3
"""K1 under sustained load (interleaved with K3, clocks settled): the pair kernel against tile-kernel configurations
KSN_K1_TILE="W,C,S" (warps, modes per lane, stages)."""
import ctypes as C, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kspace_neutrinos_b200 import capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
cfgs = sys.argv[2:] or ["auto"]
nslab = int(os.environ.get("KSN_PROBE_NSLAB", n))          # planes of the slab held by this GPU (N=4096: 512 = one of 8 GPUs)
nrbins = int(os.environ.get("KSN_PROBE_NRBINS", n // 2)); nel = nslab * n * (n // 2 + 1)
L = capi.lib(); capi.check(L.ksn_init(-1)); L.ksn_set_quiet(1)
ptr = C.c_void_p(); capi.check(L.ksn_device_malloc(C.byref(ptr), nel * 16))
capi.check(L.ksn_fill_synthetic_grid(ptr, 8, n, 0, nslab, 1, -1.0))
thr = C.POINTER(C.c_uint)(); iw = capi.c_double_p(); L.ksn_bin_tables(n, nrbins, C.byref(thr), C.byref(iw))
power, keff = np.zeros(nrbins), np.zeros(nrbins); count = np.zeros(nrbins, dtype=np.int64); m2 = C.c_double()
dp = lambda a: a.ctypes.data_as(capi.c_double_p)
logkk = np.log(np.geomspace(1.0, n * 0.86, nrbins) * 2 * np.pi / 512000.0); ratio = np.linspace(0.9, 0.1, nrbins)
L.ksn_timing_enable(1); t = capi.Timing()
def k1():
    L.ksn_timing_reset()
    capi.check(L.ksn_powerspectrum_sums(ptr, 8, n, nrbins, 0, nslab, thr, iw, dp(power), dp(keff), count.ctypes.data_as(capi.c_longlong_p), C.byref(m2)))
    L.ksn_timing_get(C.byref(t)); return t.k1_ms
def k3():
    L.ksn_timing_reset()
    capi.check(L.ksn_scale_modes(ptr, 8, n, 0, nslab, 512000.0, dp(logkk), dp(ratio), nrbins, 0.0))   # norm 0: the grid stays put
    L.ksn_timing_get(C.byref(t)); return t.k3_ms
os.environ["KSN_K1_PAIR"] = "1"
k1()
for _ in range(6): k3(); k1()          # settle the clocks
ts = []
for _ in range(6): k3(); ts.append(k1())
ref = power.copy()
print(f"pair kernel: median {np.median(ts):.2f} ms  min {min(ts):.2f}  ({nel*16/np.median(ts)/1e6:.0f} GB/s)", flush=True)
del os.environ["KSN_K1_PAIR"]
for cfg in cfgs:
    if cfg == "auto": os.environ.pop("KSN_K1_TILE", None)
    else: os.environ["KSN_K1_TILE"] = cfg
    ts = []
    for _ in range(6):
        k3(); ts.append(k1())
    p = power.copy()
    k1(); same = np.array_equal(p, power)
    print(f"tile {cfg}: median {np.median(ts):.2f} ms  min {min(ts):.2f}  ({nel*16/np.median(ts)/1e6:.0f} GB/s)  max|dP/P| vs pair {np.nanmax(np.abs(p[ref != 0]/ref[ref != 0]-1)):.1e}  bitwise repeatable {same}", flush=True)
