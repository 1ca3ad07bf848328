#!/usr/bin/env python3
"""Where the K2 kernel's critical path -- the CTA of the deepest k bin -- spends its cycles.  Needs the developer build
    make -C kspace_neutrinos_b200 clean; make -C kspace_neutrinos_b200 EXTRA_NVFLAGS=-DKSN_K2_TRACE
(csrc/k2_delta_nu.cu, g_k2_trace); runs tools/k2_bench.py's set-up and prints the counters of its last step.
usage: k2_trace.py [nk [hybrid [nondegenerate]]]"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kspace_neutrinos_b200 import capi  # noqa: E402

L = capi.lib()
if not hasattr(L, "ksn_k2_trace"):
    sys.exit("this libkspace_neutrinos_b200.so was built without -DKSN_K2_TRACE")
if len(sys.argv) < 2:
    sys.argv += ["788", "1"]
capi.check(L.ksn_init(-1))
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "k2_bench.py")).read()
src = src.replace("    L.ksn_timing_reset()\n", "    L.ksn_timing_reset(); L.ksn_k2_trace(None, 1)\n", 1)
exec(compile(src, "k2_bench.py", "exec"))
z = (C.c_ulonglong * 16)()
L.ksn_k2_trace(z, 0)
v = list(z)
print(f"deepest bin's CTA: set-up {v[3]} cycles, {v[2]} passes through the integrand {v[0]} cycles ({v[0] // max(1, v[2])} each), "
      f"replay of QAG's loop {v[1]} cycles, whole CTA {v[4]} cycles = {(v[6] - v[5]) / 1e3:.1f} us")
print(f"kernel, first CTA's start to last CTA's end: {(v[7] - v[8]) / 1e3:.1f} us; the deepest bin's CTA runs {(v[5] - v[8]) / 1e3:.1f} .. {(v[6] - v[8]) / 1e3:.1f} us")
