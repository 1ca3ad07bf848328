/* Host numerics (see ksn_numeric.h).  Written for this project from the published QUADPACK /
 * GSL algorithms; constants from tools/derive_gk61.py. */
#include "ksn_numeric.h"
#include "../csrc/ksn_gk61_tables.h"
#include <float.h>
#include <math.h>
#include <stdlib.h>

static const double XGK[31] = KSN_XGK61_INIT;
static const double WGK[31] = KSN_WGK61_INIT;
static const double WG[15] = KSN_WG30_INIT;

struct gk_out { double integral, err, resabs, resasc; };

static double gk_error(double raw, double resabs, double resasc)
{
    double err = fabs(raw);
    if (resasc != 0 && err != 0) {
        const double s = pow(200 * err / resasc, 1.5);
        err = s < 1 ? resasc * s : resasc;
    }
    if (resabs > DBL_MIN / (50 * DBL_EPSILON)) {
        const double floor_err = 50 * DBL_EPSILON * resabs;
        if (floor_err > err) err = floor_err;
    }
    return err;
}

/* one application of the 61-point rule on [a,b]; sums accumulated in QUADPACK order */
static struct gk_out gk61(ksn_integrand f, void *ctx, double a, double b)
{
    double lo[30], hi[30];
    const double mid = 0.5 * (a + b), hl = 0.5 * (b - a);
    const double fc = f(mid, ctx);
    double gauss = 0, kron = fc * WGK[30], sabs = fabs(kron);
    for (int g = 0; g < 15; g++) {                 /* Gauss nodes: odd indices */
        const int j = 2 * g + 1;
        const double dx = hl * XGK[j];
        lo[j] = f(mid - dx, ctx);
        hi[j] = f(mid + dx, ctx);
        const double s = lo[j] + hi[j];
        gauss += WG[g] * s;
        kron += WGK[j] * s;
        sabs += WGK[j] * (fabs(lo[j]) + fabs(hi[j]));
    }
    for (int j = 0; j < 30; j += 2) {              /* Kronrod-only nodes: even indices */
        const double dx = hl * XGK[j];
        lo[j] = f(mid - dx, ctx);
        hi[j] = f(mid + dx, ctx);
        kron += WGK[j] * (lo[j] + hi[j]);
        sabs += WGK[j] * (fabs(lo[j]) + fabs(hi[j]));
    }
    const double mean = 0.5 * kron;
    double sasc = WGK[30] * fabs(fc - mean);
    for (int j = 0; j < 30; j++) sasc += WGK[j] * (fabs(lo[j] - mean) + fabs(hi[j] - mean));
    struct gk_out o;
    o.integral = kron * hl;
    o.resabs = sabs * fabs(hl);
    o.resasc = sasc * fabs(hl);
    o.err = gk_error((kron - gauss) * hl, o.resabs, o.resasc);
    return o;
}

struct piece { double a, b, val, err; };

int ksn_qag61(ksn_integrand f, void *ctx, double a, double b, double epsabs, double epsrel, int limit,
              double *result, double *abserr)
{
    struct piece list[KSN_QAG_MAX_INTERVALS];
    *result = 0;
    *abserr = 0;
    if (limit > KSN_QAG_MAX_INTERVALS || limit < 1) return KSN_Q_EINVAL;
    if (epsabs <= 0 && (epsrel < 50 * DBL_EPSILON || epsrel < 0.5e-28)) return KSN_Q_EBADTOL;
    const struct gk_out first = gk61(f, ctx, a, b);
    double tol = fmax(epsabs, epsrel * fabs(first.integral));
    *result = first.integral;
    *abserr = first.err;
    if (first.err <= 50 * DBL_EPSILON * first.resabs && first.err > tol) return KSN_Q_EROUND;
    if ((first.err <= tol && first.err != first.resasc) || first.err == 0.0) return KSN_Q_OK;
    if (limit == 1) return KSN_Q_EMAXITER;
    list[0].a = a; list[0].b = b; list[0].val = first.integral; list[0].err = first.err;
    int n = 1, iter = 1, worst = 0, noise1 = 0, noise2 = 0, trouble = 0;
    double area = first.integral, errsum = first.err;
    do {
        const struct piece p = list[worst];
        const double m = 0.5 * (p.a + p.b);
        const struct gk_out l = gk61(f, ctx, p.a, m), r = gk61(f, ctx, m, p.b);
        const double both = l.integral + r.integral, eboth = l.err + r.err;
        errsum += eboth - p.err;
        area += both - p.val;
        if (l.resasc != l.err && r.resasc != r.err) {
            if (fabs(p.val - both) <= 1.0e-5 * fabs(both) && eboth >= 0.99 * p.err) noise1++;
            if (iter >= 10 && eboth > p.err) noise2++;
        }
        tol = fmax(epsabs, epsrel * fabs(area));
        if (errsum > tol) {
            if (noise1 >= 6 || noise2 >= 20) trouble = 2;
            const double tiny = (1 + 100 * DBL_EPSILON) * (fabs(m) + 1000 * DBL_MIN);
            if (fabs(p.a) <= tiny && fabs(p.b) <= tiny) trouble = 3;
        }
        /* the half with the larger error keeps the slot; the other one is appended */
        const struct piece L = { p.a, m, l.integral, l.err }, R = { m, p.b, r.integral, r.err };
        if (r.err > l.err) { list[worst] = R; list[n] = L; } else { list[worst] = L; list[n] = R; }
        n++;
        worst = 0;
        for (int i = 1; i < n; i++) if (list[i].err > list[worst].err) worst = i;
        iter++;
    } while (iter < limit && !trouble && errsum > tol);
    double s = 0;
    for (int i = 0; i < n; i++) s += list[i].val;
    *result = s;
    *abserr = errsum;
    if (errsum <= tol) return KSN_Q_OK;
    if (trouble == 2) return KSN_Q_EROUND;
    if (trouble == 3) return KSN_Q_ESING;
    if (iter == limit) return KSN_Q_EMAXITER;
    return KSN_Q_EFAILED;
}

void ksn_cspline_natural(const double *x, const double *y, int n, double *c)
{
    c[0] = 0;
    c[n - 1] = 0;
    const int m = n - 2;                 /* interior unknowns c[1..n-2] */
    if (m < 1) return;
    double *piv = malloc(sizeof(double) * m), *mul = malloc(sizeof(double) * m);
    double *u = c + 1;
    for (int i = 0; i < m; i++) {
        const double h0 = x[i + 1] - x[i], h1 = x[i + 2] - x[i + 1];
        u[i] = 3.0 * ((y[i + 2] - y[i + 1]) * (h1 != 0 ? 1.0 / h1 : 0.0) - (y[i + 1] - y[i]) * (h0 != 0 ? 1.0 / h0 : 0.0));
    }
    if (m == 1) {
        u[0] /= 2.0 * ((x[2] - x[1]) + (x[1] - x[0]));
    } else {
        /* symmetric tridiagonal L D L^T; diagonal 2(h_{i+1}+h_i), off-diagonal h_{i+1} */
        piv[0] = 2.0 * ((x[2] - x[1]) + (x[1] - x[0]));
        mul[0] = (x[2] - x[1]) / piv[0];
        for (int i = 1; i < m; i++) {
            const double d = 2.0 * ((x[i + 2] - x[i + 1]) + (x[i + 1] - x[i]));
            piv[i] = d - (x[i + 1] - x[i]) * mul[i - 1];
            if (i < m - 1) mul[i] = (x[i + 2] - x[i + 1]) / piv[i];
        }
        for (int i = 1; i < m; i++) u[i] -= mul[i - 1] * u[i - 1];
        for (int i = 0; i < m; i++) u[i] /= piv[i];
        for (int i = m - 2; i >= 0; i--) u[i] -= mul[i] * u[i + 1];
    }
    free(piv);
    free(mul);
}

int ksn_locate(const double *x, int n, double xq, int *hint)
{
    int lo = 0, hi = n - 1;
    if (hint) {
        const int h = *hint;
        if (h >= 0 && h < n - 1) {
            if (xq >= x[h] && xq < x[h + 1]) return h;
            if (xq < x[h]) hi = h; else lo = h;
        }
    }
    while (hi > lo + 1) {
        const int mid = (lo + hi) / 2;
        if (x[mid] > xq) hi = mid; else lo = mid;
    }
    if (hint) *hint = lo;
    return lo;
}

double ksn_cspline_eval(const double *x, const double *y, const double *c, int n, double xq, int *hint)
{
    const int i = ksn_locate(x, n, xq, hint);
    const double dx = x[i + 1] - x[i], dy = y[i + 1] - y[i], t = xq - x[i];
    const double b = dy / dx - dx * (c[i + 1] + 2.0 * c[i]) / 3.0;
    const double d = (c[i + 1] - c[i]) / (3.0 * dx);
    return y[i] + t * (b + t * (c[i] + t * d));
}

double ksn_linear_eval(const double *x, const double *y, int n, double xq, int *hint)
{
    const int i = ksn_locate(x, n, xq, hint);
    return y[i] + (xq - x[i]) / (x[i + 1] - x[i]) * (y[i + 1] - y[i]);
}
