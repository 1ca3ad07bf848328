"""Debug: the hundred-step test body; per step the worst delta_nu deviation from the oracle, where, and evaluation counts."""
import sys; sys.path.insert(0, '.')
import ctypes as C, numpy as np
from tests import refs
from kspace_neutrinos_b200 import capi
L = capi.lib(); capi.check(L.ksn_init(-1)); L.ksn_set_quiet(1)
L.ksn_last_k2_evals.restype = C.c_ulonglong
o = refs.orc()
n = 32
for hybrid, masses in [(False, (0.1, 0.1, 0.1)), (True, (0.2, 0.1, 0.3))]:
    g = refs.random_grid(n, seed=77)
    refs.init_module(L, n, masses=masses, hybrid=hybrid)
    dt = capi.global_delta_tot_table()
    m = refs.orc_module(n, masses=masses, hybrid=hybrid)
    dev = refs.DeviceBuffer(L, g)
    cur = g.copy()
    times = [0.01]
    for i in range(1, 95):
        times.append(0.01 + 0.0105 * i)
        if i % 6 == 2: times.append(0.01 + 0.0105 * i + 0.004)
    for a in times:
        L.add_nu_power_to_rhogrid_f64(a, refs.BOX, dev.ptr, n, 0, n, 0)
        ev = L.ksn_last_k2_evals()
        rc = o.orc_add_nu_power_to_rhogrid(C.byref(m), a, refs.BOX, cur.ctypes.data_as(C.c_void_p), 1, n, 0, n)
        got = np.array([dt.delta_nu_last[i] for i in range(dt.nk)])
        want = np.array([m.dtot.delta_nu_last[i] for i in range(dt.nk)])
        d = np.abs(got / want - 1)
        # history rows
        na = dt.ia
        hist = 0.0
        for k in range(dt.nk):
            for r in range(na):
                x, y = dt.delta_tot[k][r], m.dtot.delta_tot[k * m.dtot.namax + r]
                hist = max(hist, abs(x / y - 1))
        print(f"a={a:.4f} ia={dt.ia} worst={d.max():.3e} at k#{d.argmax()} hist={hist:.3e} evals dev={ev} orc={m.dtot.n_evals} diff={ev - m.dtot.n_evals}", flush=True)
    dev.free()
print("done")
