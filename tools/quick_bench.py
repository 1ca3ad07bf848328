#!/usr/bin/env python3
"""Developer micro-benchmark (not the contract bench): device-resident synthetic grid, K1 and K3
timed alone with CUDA events inside the library (ksn_timing_*)."""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from kspace_neutrinos_b200 import capi  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    rb = int(sys.argv[3]) if len(sys.argv) > 3 else 8            # bytes per real: 8 (double grid) or 4 (float grid)
    planes = int(sys.argv[4]) if len(sys.argv) > 4 and int(sys.argv[4]) > 0 else n   # a slab of the grid (same kernels: they go by width)
    greens = len(sys.argv) > 5 and sys.argv[5] == "greens"       # also time K3 fused with the PM Green's function
    L = capi.lib()
    capi.check(L.ksn_init(-1))
    L.ksn_set_quiet(1)
    nrbins = n // 2
    nel = planes * n * (n // 2 + 1)
    ptr = C.c_void_p()
    capi.check(L.ksn_device_malloc(C.byref(ptr), nel * 2 * rb))
    capi.check(L.ksn_fill_synthetic_grid(ptr, rb, n, 0, planes, 20261017, -1.0))
    thr = C.POINTER(C.c_uint)()
    iw = capi.c_double_p()
    L.ksn_bin_tables(n, nrbins, C.byref(thr), C.byref(iw))
    power, keff = np.zeros(nrbins), np.zeros(nrbins)
    count = np.zeros(nrbins, dtype=np.int64)
    m2 = C.c_double()
    dp = lambda a: a.ctypes.data_as(capi.c_double_p)
    L.ksn_timing_enable(1)
    t = capi.Timing()
    for label in ("K1 full (first call, geometry)", "K1 fast"):
        for r in range(reps if label == "K1 fast" else 1):
            L.ksn_timing_reset()
            capi.check(L.ksn_powerspectrum_sums(ptr, rb, n, nrbins, 0, planes, thr, iw, dp(power), dp(keff),
                                                count.ctypes.data_as(capi.c_longlong_p), C.byref(m2)))
            L.ksn_timing_get(C.byref(t))
            print(f"{label}: k1 {t.k1_ms:.3f} ms  reduce {t.k1_reduce_ms:.3f} ms  -> {nel * 2 * rb / t.k1_ms / 1e6:.1f} GB/s  [{L.ksn_last_k1_kernel().decode()}]", flush=True)
    logkk = np.log(np.geomspace(1.0, n * 0.86, nrbins) * 2 * np.pi / 512000.0)   # log-spaced knots like keff of log-k bins
    ratio = np.linspace(0.9, 0.1, nrbins)
    for r in range(reps):
        L.ksn_timing_reset()
        capi.check(L.ksn_scale_modes(ptr, rb, n, 0, planes, 512000.0, dp(logkk), dp(ratio), nrbins, 0.01))
        L.ksn_timing_get(C.byref(t))
        print(f"K3: {t.k3_ms:.3f} ms -> {nel * 4 * rb / t.k3_ms / 1e6:.1f} GB/s  [{L.ksn_last_k3_kernel().decode()}]", flush=True)
    if greens:
        asmth2 = (2 * np.pi * 1.25 / n) ** 2
        for r in range(reps):
            if r == reps // 2:
                capi.check(L.ksn_fill_synthetic_grid(ptr, rb, n, 0, planes, 20261017, -1.0))     # before the values underflow
            L.ksn_timing_reset()
            capi.check(L.ksn_scale_modes_greens(ptr, rb, n, 0, planes, 512000.0, dp(logkk), dp(ratio), nrbins, 0.01, iw, asmth2))
            L.ksn_timing_get(C.byref(t))
            print(f"K3+Greens: {t.k3_ms:.3f} ms -> {nel * 4 * rb / t.k3_ms / 1e6:.1f} GB/s  [{L.ksn_last_k3_kernel().decode()}]", flush=True)
    print("sum count", count.sum(), n ** 3 - 1)


if __name__ == "__main__":
    main()
