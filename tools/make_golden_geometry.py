"""TEST INFRASTRUCTURE: writes tests/golden/geometry_counts.npz -- per-bin mode counts and sum(|k|) of
total_powerspectrum (powerspectrum.c:33-89) over the WHOLE grid at PMGRID = 1024, 2048, 4096 with nrbins = PMGRID/2 (the
BASELINE.json sizes the reference cannot sweep in test time).  The numbers come from tools/gen_geometry_golden.c; before
anything is stored, that generator is checked against the compiled reference itself (oracle/_ref/libksref_double.so, run
on a constant grid -- counts and keff do not depend on the data) at every size the reference can do here, including 192
(where the corner mode sits on a bin edge to within an ulp).  Run in the container that has /root/reference:
    python tools/make_golden_geometry.py"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import refs  # noqa: E402


def generate(exe, n, nrbins):
    out = subprocess.run([exe, str(n), str(nrbins)], capture_output=True, text=True, check=True).stdout.split("\n")
    rows = [l.split() for l in out if l.strip()]
    assert len(rows) == nrbins
    return np.array([int(r[1]) for r in rows], dtype=np.int64), np.array([float(r[2]) for r in rows])


def main():
    ref = refs.ref_lib(True)
    assert ref is not None, "needs oracle/_ref (the reference's sources compiled here)"
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "gen")
        subprocess.run(["gcc", "-O2", "-o", exe, os.path.join(ROOT, "tools", "gen_geometry_golden.c"), "-lm"], check=True)
        for n, nrbins in ((4, 15), (16, 8), (64, 32), (96, 48), (128, 64), (128, 200), (192, 96), (256, 128)):
            cnt, ksum = generate(exe, n, nrbins)
            assert cnt.sum() == n ** 3 - 1
            g = np.ones((n, n, n // 2 + 1, 2))
            r_n, _, r_c, r_k = refs.total_powerspectrum(ref, g, nrbins)
            keep = cnt > 0
            assert r_n == np.count_nonzero(keep), (n, nrbins)
            assert np.array_equal(r_c[:r_n], cnt[keep]), (n, nrbins)
            np.testing.assert_allclose(r_k[:r_n], ksum[keep] / cnt[keep], rtol=5e-12)   # the reference adds up to 1e7 terms per bin in double, in loop order
            print(f"generator == reference at PMGRID {n}, nrbins {nrbins}: {r_n} non-empty bins")
        store = {}
        for n in (1024, 2048, 4096):
            cnt, ksum = generate(exe, n, n // 2)
            assert cnt.sum() == n ** 3 - 1
            store[f"count_{n}"] = cnt
            store[f"keffsum_{n}"] = ksum
            print(f"PMGRID {n}: {np.count_nonzero(cnt)} non-empty bins of {n // 2}")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "geometry_counts.npz"), **store)


if __name__ == "__main__":
    main()
