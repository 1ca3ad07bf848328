#!/usr/bin/env python3
"""Developer benchmark of the whole PM step (add_nu_power_to_rhogrid) on ONE GPU for a double OR a float grid, device
resident, with the library's per-phase timers -- bench.py (the contract bench) is fixed to double grids.
   python tools/step_bench.py PMGRID real_bytes [steps] [hybrid]
Same set-up as bench.py: 3 x 0.1 eV, 98-row history, a advances by less than 0.009 per step (steady state)."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kspace_neutrinos_b200 import capi, host  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
rb = int(sys.argv[2]) if len(sys.argv) > 2 else 8
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
hybrid = int(sys.argv[4]) if len(sys.argv) > 4 else 1
L = capi.lib()
capi.check(L.ksn_init(-1))
transfer = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ics_transfer_99.dat")
cosmo = host.Cosmology(transfer_file=transfer, mnu=(0.1, 0.1, 0.1), hybrid_neutrinos_on=hybrid)
sim = host.KspaceNeutrinos(cosmo, n)
slab = host.Slab(0, n)
grid = host.DeviceGrid(n, slab, rb)
grid.fill_synthetic()
sim.add_nu_power_to_rhogrid(cosmo.time_transfer, grid.ptr, slab, rb)
sim.seed_history(98)
a = 0.98
da = min(0.001, (0.9995 - a) / (steps + 5))
for _ in range(3):
    a += da
    sim.add_nu_power_to_rhogrid(a, grid.ptr, slab, rb)
L.ksn_timing_enable(1)
L.ksn_timing_reset()
capi.check(L.ksn_device_synchronize())
t0 = time.perf_counter()
for _ in range(steps):
    a += da
    sim.add_nu_power_to_rhogrid(a, grid.ptr, slab, rb)
capi.check(L.ksn_device_synchronize())
wall = (time.perf_counter() - t0) * 1e3 / steps
t = capi.Timing()
L.ksn_timing_get(C.byref(t))
modes = n * n * (n // 2 + 1)
k1, k2, k3 = t.k1_ms / steps, t.k2_ms / steps, t.k3_ms / steps
print(f"PMGRID {n} {'double' if rb == 8 else 'float'} grid ({modes * 2 * rb / 1e9:.1f} GB): {wall:.3f} ms per step = {modes / wall / 1e6:.1f} Gmodes/s = "
      f"{6 * rb * modes / wall / 1e6:.0f} GB/s of algorithmic traffic;  K1 {k1:.3f} ms ({2 * rb * modes / k1 / 1e6:.0f} GB/s)  "
      f"K2 {k2:.3f} ms  K3 {k3:.3f} ms ({4 * rb * modes / k3 / 1e6:.0f} GB/s)  [{L.ksn_last_k1_kernel().decode()}]  [{L.ksn_last_k3_kernel().decode()}]")
