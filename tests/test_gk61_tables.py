"""The 61-point Gauss-Kronrod constants (oracle/gk61_tables.h for the checker, csrc/ksn_gk61_tables.h for K2) are generated
by the same script (tools/derive_gk61.py), so agreement between them proves nothing.  Pinned here to INDEPENDENT sources:
scipy's Gauss-Legendre nodes/weights for the embedded 30-point rule, the defining properties of the Kronrod extension
(weights sum to 2, symmetric rule exact for every polynomial up to degree 3*30+1 = 91, checked in exact rational arithmetic
on the even monomials), and the digits of the QUADPACK dqk61 listing quoted in SURVEY.md Appendix B.1."""
import os
import re
from fractions import Fraction

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADERS = [os.path.join(ROOT, "oracle", "gk61_tables.h"), os.path.join(ROOT, "kspace_neutrinos_b200", "csrc", "ksn_gk61_tables.h")]


def parse(path):
    txt = open(path).read()
    out = {}
    for name in ("KSN_XGK61_INIT", "KSN_WGK61_INIT", "KSN_WG30_INIT"):
        m = re.search(name + r"\s*\{(.*?)\}", txt, re.S)
        assert m, (path, name)
        out[name] = [s.strip() for s in m.group(1).replace("\\", " ").split(",") if s.strip()]
    return out["KSN_XGK61_INIT"], out["KSN_WGK61_INIT"], out["KSN_WG30_INIT"]


@pytest.mark.parametrize("path", HEADERS)
def test_gauss_half_equals_scipy_gauss_legendre_30(path):
    from scipy.special import roots_legendre
    xs, ws, wg = parse(path)
    assert (len(xs), len(ws), len(wg)) == (31, 31, 15)
    x, w = roots_legendre(30)
    pos = np.argsort(-x)[:15]                                    # the 15 positive nodes, descending
    np.testing.assert_allclose([float(xs[j]) for j in range(1, 30, 2)], x[pos], rtol=0, atol=3e-16)
    np.testing.assert_allclose([float(v) for v in wg], w[pos], rtol=2e-12, atol=0)       # (scipy's own weights are good to ~5e-13)
    # the textbook weight formula w_i = 2 / ((1 - x_i^2) P_30'(x_i)^2), in exact rational arithmetic on the table's digits
    for j, wj in zip(range(1, 30, 2), wg):
        xq = Fraction(xs[j])
        p0, p1 = Fraction(1), xq
        for k in range(2, 31):
            p0, p1 = p1, ((2 * k - 1) * xq * p1 - (k - 1) * p0) / k
        assert abs(p1) < Fraction(1, 10 ** 36)                                           # a root of P_30
        dp = 30 * (xq * p1 - p0) / (xq * xq - 1)
        assert abs(2 / ((1 - xq * xq) * dp * dp) - Fraction(wj)) < Fraction(1, 10 ** 36)


@pytest.mark.parametrize("path", HEADERS)
def test_kronrod_rule_is_exact_to_degree_91(path):
    xs, ws, _ = parse(path)
    X = [Fraction(s) for s in xs]                                # the decimal strings as exact rationals (40 digits)
    Wt = [Fraction(s) for s in ws]
    assert X[30] == 0 and all(X[j] > X[j + 1] for j in range(30))
    assert abs(2 * sum(Wt[:30]) + Wt[30] - 2) < Fraction(1, 10 ** 36)
    for deg in range(0, 92, 2):                                  # odd monomials vanish by symmetry
        got = 2 * sum(wj * xj ** deg for wj, xj in zip(Wt[:30], X[:30])) + (Wt[30] if deg == 0 else 0)
        assert abs(got - Fraction(2, deg + 1)) < Fraction(1, 10 ** 38), deg          # (the digits' own truncation: ~1e-42)
    # ... and NOT beyond: at the next even monomial the rule's true error appears (2.7e-31, eleven orders above the above)
    deg = 92
    got = 2 * sum(wj * xj ** deg for wj, xj in zip(Wt[:30], X[:30]))
    assert abs(got - Fraction(2, deg + 1)) > Fraction(1, 10 ** 33)


@pytest.mark.parametrize("path", HEADERS)
def test_first_and_last_entries_match_the_quadpack_listing(path):
    xs, ws, wg = parse(path)
    # QUADPACK dqk61 (as quoted with 33 digits in SURVEY.md Appendix B.1)
    assert xs[0].startswith("0.99948441005049063757132589570581")
    assert xs[29].startswith("0.051471842555317695833025213166")
    assert ws[0].startswith("0.0013890136986770076245515912267")
    assert ws[30].startswith("0.051494729429451567558340433647")
    assert wg[0].startswith("0.0079681924961666056154658834746")
    assert wg[14].startswith("0.10285265289355884034128563670")
