#!/usr/bin/env python3
"""PMGRID = 4096 slab (BASELINE config 5, the share of one of 8 GPUs by default) under sustained load: K1 with and
without the bin window (KSN_K1_WIN), K3 with rows cut into bulk-copy pieces against the plain-load kernel (KSN_K3_NOSPLIT)."""
import ctypes as C, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kspace_neutrinos_b200 import capi
n = 4096
nslab = int(sys.argv[1]) if len(sys.argv) > 1 else 512
nrbins = n // 2; nel = nslab * n * (n // 2 + 1)
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
k1()
for _ in range(4): k3(); k1()          # geometry cache, clocks settled
ref = None
for win in ("0", "1"):               # bin window off / on
    os.environ["KSN_K1_WIN"] = win
    ts = []
    for _ in range(6): k3(); ts.append(k1())
    p = power.copy()
    if ref is None: ref = p
    print(f"K1 KSN_K1_WIN={win}: median {np.median(ts):.2f} ms  min {min(ts):.2f}  ({nel*16/np.median(ts)/1e6:.0f} GB/s)  {L.ksn_last_k1_kernel().decode()}  "
          f"max|dP/P| vs window off {np.nanmax(np.abs(p[ref != 0]/ref[ref != 0]-1)):.1e}", flush=True)
for nosplit in (True, False):
    if nosplit: os.environ["KSN_K3_NOSPLIT"] = "1"
    else: os.environ.pop("KSN_K3_NOSPLIT", None)
    ts = []
    for _ in range(6): k1(); ts.append(k3())
    print(f"K3 {'plain loads' if nosplit else 'bulk-copy row pieces'}: median {np.median(ts):.2f} ms  min {min(ts):.2f}  ({nel*32/np.median(ts)/1e6:.0f} GB/s)", flush=True)
