#!/usr/bin/env python3
"""Developer benchmark of the K2 phase alone (no grid): nk bins, ~99 stored rows, hybrid on/off."""
import ctypes as C
import sys
import os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kspace_neutrinos_b200 import capi  # noqa: E402
from tests import refs  # noqa: E402

nk = int(sys.argv[1]) if len(sys.argv) > 1 else 880
hybrid = int(sys.argv[2]) if len(sys.argv) > 2 else 1
masses = (0.1, 0.1, 0.1) if len(sys.argv) <= 3 else (0.2, 0.1, 0.3)
L = capi.lib()
capi.check(L.ksn_init(-1))
L.ksn_set_quiet(1)
om = refs.make_omnu(L, masses)
if hybrid:
    L.init_hybrid_nu(C.byref(om.hybnu), (C.c_double * 3)(*masses), 500.0, 2.99792458e10 / 1e5, 0.333, om.kBtnu)
refs.set_background(L, om)
tr = refs.load_transfer(L)
kk = np.geomspace(2 * np.pi / refs.BOX * 1.01, 2 * np.pi / refs.BOX * 1700, nk)
dcdm = 1e5 * (kk / kk[0]) ** -0.8
d = refs.new_delta_tot(L, om, nk)
L.delta_tot_init(C.byref(d), nk, refs.dptr(kk), refs.dptr(dcdm), C.byref(tr), 0.01)
rows = 98
for i in range(1, rows):
    d.scalefact[i] = np.log(0.01 * (i + 1))
    for k in range(nk):
        d.delta_tot[k][i] = d.delta_tot[k][0] * (i + 1)
d.ia = rows
out = np.zeros(nk)
L.ksn_timing_enable(1)
t = capi.Timing()
a = 0.98
for it in range(6):
    a += 0.001
    L.ksn_timing_reset()
    L.get_delta_nu_update(C.byref(d), a, nk, refs.dptr(kk), refs.dptr(dcdm), refs.dptr(out), C.byref(tr))
    L.ksn_timing_get(C.byref(t))
    print(f"K2 phase: {t.k2_ms:.3f} ms  launches {t.launches}  Na={d.ia + 1}  evals {L.ksn_last_k2_evals()}  -> {(L.ksn_last_k2_evals() - 61 * 16 * (d.ia + 1)) / 61 / nk:.1f} GK61 passes per k, max {L.ksn_last_k2_max_passes()}, slowest bin: {L.ksn_last_k2_max_trips()} passes through the integrand")
