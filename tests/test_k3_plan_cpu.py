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


def _hash(L, n, box, logkk, ratio, norm, fresh):
    h = C.c_ulonglong()
    assert L.ksn_k3_table_hash(n, box, refs.dptr(logkk), refs.dptr(ratio), len(logkk), norm, fresh, C.byref(h)) == 0
    return h.value


def test_table_built_from_cached_knot_geometry_equals_a_fresh_build(ksn):
    """The knots of the table are log(keff) -- the same every PM step -- so what depends on them alone (k^2 of the knots,
    reciprocals, integer thresholds, lookup cells) is kept between builds and only the ratio-dependent half is redone.
    The bytes that go to the device, the kernel parameters and the decisions must not depend on where the knot half came
    from: fresh, cached, cached after other knots were seen in between, cached with a different ratio / norm."""
    rng = np.random.default_rng(7)
    tabs = []
    for n, nb in ((2048, 788), (4096, 1700), (256, 109), (64, 29)):
        lk = np.log(np.geomspace(1.0, n * 0.86, nb) * 2 * np.pi / refs.BOX)
        tabs.append((n, refs.BOX, lk, np.linspace(0.9, 0.1, nb), 0.01))
        lk2 = np.sort(lk + rng.normal(0, 0.3 * np.diff(lk).min(), nb))
        tabs.append((n, refs.BOX, lk2, rng.uniform(0.05, 1.0, nb), 0.02))
        tabs.append((n, refs.BOX / 2, lk2 + 0.3, np.exp(-np.linspace(0, 5, nb)), 0.3))
    lk = np.log(np.geomspace(1.0, 100, 40) * 2 * np.pi / refs.BOX)
    lk[20] = lk[19] + 1e-7                                        # two knots in one lookup cell
    tabs.append((256, refs.BOX, lk, np.linspace(0.5, 0.2, 40), 0.01))
    fresh = [_hash(ksn, *t, 1) for t in tabs]
    assert len(set(fresh)) == len(tabs)
    for t, want in zip(tabs, fresh):                              # cold, then twice from the cache
        assert [_hash(ksn, *t, f) for f in (1, 0, 0)] == [want] * 3
    assert [_hash(ksn, *t, 0) for t in tabs] == fresh             # other knots in between: the cache is replaced, not mixed
    n, box, lk, ratio, norm = tabs[0]
    for step in range(4):                                         # a run of PM steps: same knots, new ratios and norm
        r, nm = ratio * (1 + 0.07 * step), norm * (1 + 0.5 * step)
        assert _hash(ksn, n, box, lk, r, nm, 0) == _hash(ksn, n, box, lk, r, nm, 1)
    assert ksn.ksn_k3_table_hash(64, refs.BOX, refs.dptr(lk[::-1].copy()), refs.dptr(ratio), 10, norm, 0, C.byref(C.c_ulonglong())) != 0
