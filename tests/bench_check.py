"""CHECKER for bench.py (test infrastructure; calls oracle/): one extra PM step OUTSIDE the timed region whose results are
compared with the CPU restatement of the reference on the same inputs, so that the JSON line of a run also says that the
kernels it timed -- the instantiations chosen by the benchmark's own grid width and history depth -- returned the
reference's numbers.

* K3 (interface_gadget.c:163-188): a few planes of the device-resident slab are copied out before and after the step; the
  expected planes are orc_scale_modes(before) with exactly the table the pass used (ksn_last_k3_table).
* K2 (delta_tot_table.c:507-611): the integrator's stored rows, the provisional row of the step and delta_nu_last are read
  from the module; orc_get_delta_nu_combined on a copy of that state must return delta_nu_last.
* K1 (powerspectrum.c:56-89; one rank only -- the entry all-reduces): the bin sums of the sample planes through
  ksn_powerspectrum_sums, first sweep (three-sum kernel) and cached sweep (tile kernel), against orc_powerspectrum_sums.
Tolerances are BASELINE.json's: counts exact, everything else 1e-10 relative.
"""
import ctypes as C

import numpy as np


def _planes(L, grid_ptr, n, first_plane_local, count):
    """planes [first, first+count) of a device-resident double slab -> host array"""
    from kspace_neutrinos_b200 import capi
    plane_bytes = n * (n // 2 + 1) * 16
    out = np.empty((count, n, n // 2 + 1, 2))
    capi.check(L.ksn_memcpy_d2h(out.ctypes.data_as(C.c_void_p), C.c_void_p(grid_ptr.value + first_plane_local * plane_bytes), out.nbytes), "d2h")
    return out


def _rel(got, want):
    live = want != 0
    if not live.any():
        return 0.0
    return float(np.max(np.abs(got[live] / want[live] - 1)))


def check_step(L, sim, grid, slab, n, a, masses, hybrid, world, depth=2):
    from kspace_neutrinos_b200 import capi
    from tests import refs
    o = refs.orc()
    res = {"checker": "oracle/libksn_oracle.so (C restatement of the reference, pinned to the reference sources in tests/)",
           "tolerance": "mode counts exact; P(k) sums, delta_nu, corrected modes 1e-10 relative"}
    depth = min(depth, slab.count)
    ranges = [(0, depth)] if slab.count < 2 * depth else [(0, depth), (slab.count - depth, depth)]
    before = [_planes(L, grid.ptr, n, p0, c) for p0, c in ranges]
    st = sim.state
    ia0 = st.ia
    sim.add_nu_power_to_rhogrid(a, grid.ptr, slab)
    after = [_planes(L, grid.ptr, n, p0, c) for p0, c in ranges]
    res["k3_kernel"] = L.ksn_last_k3_kernel().decode()
    res["k1_kernel"] = L.ksn_last_k1_kernel().decode()
    # ---- K3
    lk, rt = capi.c_double_p(), capi.c_double_p()
    nb, norm, box = C.c_int(), C.c_double(), C.c_double()
    capi.check(L.ksn_last_k3_table(C.byref(lk), C.byref(rt), C.byref(nb), C.byref(norm), C.byref(box)), "ksn_last_k3_table")
    logkk = np.array(lk[:nb.value])
    ratio = np.array(rt[:nb.value])
    k3 = 0.0
    changed = False
    for (p0, c), b, g in zip(ranges, before, after):
        want = b.copy()
        o.orc_scale_modes(want.ctypes.data_as(C.c_void_p), 1, n, slab.start + p0, c, box.value, refs.dptr(logkk), refs.dptr(ratio), nb.value, norm.value)
        k3 = max(k3, _rel(g, want))
        changed = changed or not np.array_equal(want, b)
    res["k3_max_rel"] = k3
    res["k3_planes"] = [[slab.start + p0, c] for p0, c in ranges]
    res["k3_factor_not_trivial"] = bool(changed)
    # ---- K2
    if st.ia == ia0:                       # the step's row was not kept: rows [0, ia) + the provisional row at ia are intact
        nk, rows = st.nk, st.ia + 1
        oc = refs.orc_cosmo(masses, a0=st.TimeTransfer, hybrid=bool(hybrid))
        od = refs.OrcDtot()
        o.orc_dtot_alloc(C.byref(od), st.nk_allocated, st.TimeTransfer, 1.0, refs.OMEGA0, C.byref(oc), refs.UNIT_TIME, refs.UNIT_LENGTH)
        od.nk = nk
        od.ia = rows
        od.init_done = 1
        for i in range(rows):
            od.scalefact[i] = st.scalefact[i]
        for k in range(nk):
            row = st.delta_tot[k]
            base = k * od.namax
            for i in range(rows):
                od.delta_tot[base + i] = row[i]
            od.delta_nu_init[k] = st.delta_nu_init[k]
            od.wavenum[k] = st.wavenum[k]
        wav = np.array([st.wavenum[k] for k in range(nk)])
        want = np.zeros(nk)
        o.orc_get_delta_nu_combined(C.byref(od), a, refs.dptr(wav), refs.dptr(want))
        got = sim.delta_nu_last()
        res["k2_max_rel"] = _rel(got, want)
        res["k2_bins"], res["k2_rows"] = int(nk), int(rows)
        res["k2_deepest_bin_rule_applications"] = int(L.ksn_last_k2_max_passes())
        o.orc_dtot_free(C.byref(od))
    else:
        res["k2_max_rel"] = None
        res["k2_note"] = "the step's row was kept (a advanced by >= 0.009): provisional row overwritten, K2 not re-derived here"
    # ---- K1 (single rank: ksn_powerspectrum_sums all-reduces)
    if world == 1:
        thr = C.POINTER(C.c_uint)()
        iw = capi.c_double_p()
        L.ksn_bin_tables(n, n // 2, C.byref(thr), C.byref(iw))
        nrb = n // 2
        plane_bytes = n * (n // 2 + 1) * 16
        worst_p, counts_ok, names = 0.0, True, []
        for (p0, c), g in zip(ranges, after):
            wp, wk = np.zeros(nrb), np.zeros(nrb)
            wc = np.zeros(nrb, dtype=np.int64)
            wm = C.c_double()
            o.orc_powerspectrum_sums(n, g.ctypes.data_as(C.c_void_p), 1, nrb, slab.start + p0, c, refs.dptr(wp), refs.dptr(wk),
                                     wc.ctypes.data_as(capi.c_longlong_p), C.byref(wm))
            for sweep in range(2):             # first sweep of this geometry: three-sum kernel; second: the tile kernel
                gp, gk = np.zeros(nrb), np.zeros(nrb)
                gc = np.zeros(nrb, dtype=np.int64)
                gm = C.c_double()
                capi.check(L.ksn_powerspectrum_sums(C.c_void_p(grid.ptr.value + p0 * plane_bytes), 8, n, nrb, slab.start + p0, c, thr, iw,
                                                    refs.dptr(gp), refs.dptr(gk), gc.ctypes.data_as(capi.c_longlong_p), C.byref(gm)), "ksn_powerspectrum_sums")
                names.append(L.ksn_last_k1_kernel().decode())
                counts_ok = counts_ok and bool(np.array_equal(gc, wc)) and gm.value == wm.value
                worst_p = max(worst_p, _rel(gp, wp), _rel(gk, wk))
        res["k1_counts_equal"] = counts_ok
        res["k1_max_rel"] = worst_p
        res["k1_sample_kernels"] = sorted(set(names))
    else:
        res["k1_counts_equal"] = None
        res["k1_note"] = "not re-derived on more than one rank (the entry all-reduces); see tests/test_multi_gpu.py, tests/test_fullsize_parity_gpu.py"
    ok = k3 <= 1e-10 and changed
    if res.get("k2_max_rel") is not None:
        ok = ok and res["k2_max_rel"] <= 1e-10
    if world == 1:
        ok = ok and res["k1_counts_equal"] and res["k1_max_rel"] <= 1e-10
    res["ok"] = bool(ok)
    return res
