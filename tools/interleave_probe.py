#!/usr/bin/env python3
"""Does K1 slow down when it alternates with K3 (sustained power) compared with K1 alone?"""
import ctypes as C, sys, os, subprocess, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kspace_neutrinos_b200 import capi
n = 2048; nrbins = n // 2; nel = n * n * (n // 2 + 1)
L = capi.lib(); capi.check(L.ksn_init(-1)); L.ksn_set_quiet(1)
ptr = C.c_void_p(); capi.check(L.ksn_device_malloc(C.byref(ptr), nel * 16))
capi.check(L.ksn_fill_synthetic_grid(ptr, 8, n, 0, n, 1, -1.0))
thr = C.POINTER(C.c_uint)(); iw = capi.c_double_p(); L.ksn_bin_tables(n, nrbins, C.byref(thr), C.byref(iw))
power, keff = np.zeros(nrbins), np.zeros(nrbins); count = np.zeros(nrbins, dtype=np.int64); m2 = C.c_double()
dp = lambda a: a.ctypes.data_as(capi.c_double_p)
logkk = np.log(np.geomspace(1.0, n * 0.86, nrbins) * 2 * np.pi / 512000.0); ratio = np.linspace(0.9, 0.1, nrbins)
rows = []
def sample():
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
    for line in p.stdout: rows.append((time.perf_counter(), line.strip()))
threading.Thread(target=sample, daemon=True).start()
time.sleep(1.5)
L.ksn_timing_enable(1); t = capi.Timing()
def k1():
    L.ksn_timing_reset()
    capi.check(L.ksn_powerspectrum_sums(ptr, 8, n, nrbins, 0, n, thr, iw, dp(power), dp(keff), count.ctypes.data_as(capi.c_longlong_p), C.byref(m2)))
    L.ksn_timing_get(C.byref(t)); return t.k1_ms
def k3():
    L.ksn_timing_reset()
    capi.check(L.ksn_scale_modes(ptr, 8, n, 0, n, 512000.0, dp(logkk), dp(ratio), nrbins, 0.01))
    L.ksn_timing_get(C.byref(t)); return t.k3_ms
k1(); k1()
t0 = time.perf_counter(); a = [k1() for _ in range(12)]; t1 = time.perf_counter()
print("K1 alone x12      :", " ".join(f"{x:.2f}" for x in a))
b = []
for _ in range(12): b.append((k3(), k1()))
t2 = time.perf_counter()
print("K3,K1 interleaved :", " ".join(f"{x:.2f}/{y:.2f}" for x, y in b))
c = [k3() for _ in range(12)]
print("K3 alone x12      :", " ".join(f"{x:.2f}" for x in c))
for lo, hi, name in ((t0, t1, "K1 alone"), (t1, t2, "interleaved")):
    sel = [r for ts, r in rows if lo <= ts <= hi]
    print(name, "clock samples:", sel[:12])
