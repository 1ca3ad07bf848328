"""How K3's passes evaluate a _delta_pow table is decided on the host from bounds over the table (csrc/k3_scale.cu:
k3_build_table) -- series length of ln(1+u), whether a float grid may use the all-float factor, from which k^2 on rows take
the branch-free path, how fine the lookup cells are.  ksn_k3_table_plan exports the decisions; pinned here without a GPU."""
import ctypes as C

import numpy as np
import pytest

from tests import refs
from tests.test_promoted_kernels_gpu import _smooth_table, _table


def plan(L, n, tab):
    logkk, ratio, norm = tab
    v = [C.c_int(), C.c_int(), C.c_uint(), C.c_int(), C.c_int()]
    assert L.ksn_k3_table_plan(n, refs.BOX, refs.dptr(logkk), refs.dptr(ratio), len(logkk), norm, *[C.byref(x) for x in v]) == 0
    return dict(zip(("series", "f32_ok", "k2_narrow", "cells", "multi"), (x.value for x in v)))


def test_smooth_step_like_tables(ksn):
    """One knot per P(k) bin, ratio falling smoothly with k, small norm: what a PM step produces."""
    for n in (2048, 4096):
        p = plan(ksn, n, _smooth_table(n, refs.BOX))
        assert p == {"series": 5, "f32_ok": 1, "k2_narrow": 2, "cells": n, "multi": 0}, (n, p)
    p = plan(ksn, 1024, _smooth_table(1024, refs.BOX))            # bins of 1024 are twice as wide in log k: u^5 is not enough
    assert (p["series"], p["f32_ok"], p["k2_narrow"]) == (9, 1, 2)
    # noisier ratio (2 % per bin, as for bins of few modes): still the float factor, but the full series
    p = plan(ksn, 2048, _smooth_table(2048, refs.BOX, jitter=0.02))
    assert (p["series"], p["f32_ok"]) == (9, 1)


def test_rough_tables_keep_the_exact_paths(ksn):
    """40 random knots with ratios jumping by tens of per cent: wide segments (log1p), steep slopes (double factor)."""
    for n in (64, 1024, 4096):
        p = plan(ksn, n, _table(n, refs.BOX, nk=min(40, n // 2)))
        assert p["series"] == 9 and p["f32_ok"] == 0 and p["multi"] == 0
        # only rows beyond the last wide segment are branch-free; the first knot is not among them
        assert p["k2_narrow"] > 100


def test_large_norm_forbids_the_float_factor(ksn):
    logkk, ratio, norm = _smooth_table(2048, refs.BOX)
    assert plan(ksn, 2048, (logkk, ratio, 0.5))["f32_ok"] == 0     # |norm ratio| up to 0.45: not "1 + small"


def test_coincident_knots_switch_the_segment_search_to_a_loop(ksn):
    logkk, ratio, norm = _table(128, refs.BOX, nk=30)
    logkk = np.sort(np.concatenate([logkk, logkk[[5, 11]] + np.array([1e-9, 1e-12])]))
    ratio = np.linspace(0.2, 0.8, len(logkk))
    p = plan(ksn, 128, (logkk, ratio, norm))
    assert p["multi"] == 1 and p["cells"] == 16384 and p["k2_narrow"] == 0xFFFFFFFF


def test_bad_tables_are_rejected(ksn):
    logkk, ratio, norm = _table(64, refs.BOX, nk=10)
    assert ksn.ksn_k3_table_plan(64, refs.BOX, refs.dptr(logkk[::-1].copy()), refs.dptr(ratio), 10, norm, None, None, None, None, None) != 0
    assert ksn.ksn_k3_table_plan(64, refs.BOX, refs.dptr(logkk), refs.dptr(ratio), 1, norm, None, None, None, None, None) != 0
