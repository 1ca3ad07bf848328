// K2 -- the linear-response integral for delta_nu(k)  (sm_100a).
//
// Replaces the body of get_delta_nu, delta_tot_table.c:522-599: the table of Nfs = 16*Na
// free-streaming lengths (one adaptive Gauss-Kronrod quadrature each, :574-584), the natural
// cubic splines over it and over every delta_tot(k, a) history (:559-596), and the per-k
// adaptive 61-point Gauss-Kronrod integral of get_delta_nu_int (:492-500, :597).
//
// Latency-bound, not bandwidth-bound: the whole state is < 2 MB.  One CTA of 128 threads per
// (k bin, mass species): the 2 x 61 abscissae of the two halves of the interval being bisected
// are evaluated by 122 lanes at once; Kronrod/Gauss/|f|/|f-mean| sums are fixed-shape shuffle
// trees (deterministic); the QAG interval list (limit 200, as the reference's GSL_VAL) lives in
// shared memory and follows GSL's qag.c decision logic step by step, so bisection decisions
// agree with the CPU oracle except on exact floating-point ties.
//
// hubble_function(a) is a HOST callback in the reference (gadget_defines.h:11).  The device
// integrand reads 1/(a H(a)) from a table sampled once on a uniform log-a grid
// (ksn_set_background) with 4-point Lagrange interpolation (relative error ~1e-16 at the
// default 16384 points, far inside the 1e-10 parity budget).
#include "ksn_internal.cuh"
#include <cooperative_groups.h>
#include "ksn_gk61_tables.h"
#include "ksn_qag_spec.h"

#include <algorithm>
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace ksn {
namespace cg = cooperative_groups;

__constant__ double c_xgk[31] = KSN_XGK61_INIT;
__constant__ double c_wgk[31] = KSN_WGK61_INIT;
__constant__ double c_wg[15] = KSN_WG30_INIT;

constexpr int K2_THREADS = 128;
constexpr int QAG_LIMIT = 200;      // GSL_VAL, kspace_neutrino_const.h:19
#ifndef KSN_K2_SPEC_DEFAULT
#define KSN_K2_SPEC_DEFAULT 0       // see k2_spec_width()
#endif

enum { Q_OK = 0, Q_EROUND = 18, Q_ESING = 21, Q_EMAXITER = 11, Q_EFAILED = 5 };

// A host Hubble rate is not smooth everywhere: the reference's own Omega_nu(a) switches from a spline table to a series at
// a = 100 kT_nu/m_nu (omega_nu_single.c:180-199), which leaves a kink in H(a).  A 4-point stencil that straddles a kink is
// only first-order accurate there (measured: 1e-10 pointwise, enough to flip adaptive-quadrature decisions against the
// CPU path).  ksn_set_background() therefore checks every cell of the uniform table against the host function and covers
// the cells that fail with PATCHES: sub-tables BG_SUB times finer, searched first.
constexpr int BG_MAX_PATCH = 4;
constexpr int BG_SUB = 4096;            // sub-cells per coarse cell inside a patch
constexpr int BG_MAX_PATCH_CELLS = 48;  // coarse cells covered by all patches together

struct BgPatch {
    double lo, hi;     // covers lo <= x < hi
    double x0, inv_h;  // sub-node j sits at x0 + j / inv_h; x0 = lo - one sub-cell
    int off, n;        // first sub-node in g[], count
};

struct BgTable {
    const double *g;   // 1/(a H(a)) at x_i = lo + i*h, followed by the patch sub-tables
    int n;
    double lo, h, inv_h;
    int npatch;
    BgPatch patch[BG_MAX_PATCH];
};

__device__ __forceinline__ double bg_lagrange(const double *__restrict__ g, int n, double s)
{
    int i = (int) floor(s);
    i = max(1, min(i, n - 3));
    const double u = s - i, um1 = u - 1.0, up1 = u + 1.0, um2 = u - 2.0;
    const double w0 = -u * um1 * um2 * (1.0 / 6.0);
    const double w1 = up1 * um1 * um2 * 0.5;
    const double w2 = -up1 * u * um2 * 0.5;
    const double w3 = up1 * u * um1 * (1.0 / 6.0);
    return w0 * __ldg(g + i - 1) + w1 * __ldg(g + i) + w2 * __ldg(g + i + 1) + w3 * __ldg(g + i + 2);
}

__device__ __forceinline__ double bg_eval(const BgTable &t, double x)
{
#pragma unroll
    for (int p = 0; p < BG_MAX_PATCH; p++)
        if (p < t.npatch && x >= t.patch[p].lo && x < t.patch[p].hi)
            return bg_lagrange(t.g + t.patch[p].off, t.patch[p].n, (x - t.patch[p].x0) * t.patch[p].inv_h);
    return bg_lagrange(t.g, t.n, (x - t.lo) * t.inv_h);
}

// ---------------------------------------------------------------- specialJ (delta_tot_table.c:417-463)
__device__ __forceinline__ double specialJ_fit_d(double x)
{
    if (x <= 0.) return 1.;
    const double x2 = x * x, x4 = x2 * x2, x8 = x4 * x4;
    return (1. + 0.0168 * x2 + 0.0407 * x4) / (1. + 2.1734 * x2 + 1.6787 * exp(4.1811 * log(x)) + 0.1467 * x8);
}

__device__ __forceinline__ double bessel_j0_d(double x)
{
    const double ax = fabs(x);
    if (ax < 0.5) {
        const double y = x * x;
        return 1.0 + y * (-1.0 / 6.0 + y * (1.0 / 120.0 + y * (-1.0 / 5040.0 + y * (1.0 / 362880.0 + y * (-1.0 / 39916800.0 + y * (1.0 / 6227020800.0))))));
    }
    return sin(x) / x;
}

// 1/d for a d inside the float range (here d = n^2 + x^2 >= 1 and |qc x| >= 0.5): float seed and two Newton steps, good to
// about an ulp, in 4 FP64 instructions instead of the ~8 plus a slow-path test of a correctly rounded 1.0/d.  d beyond
// the float range gives 0 (the term it scales is then below 1e-76 of the leading one), never a NaN.
__device__ __forceinline__ double fast_rcp_d(double d)
{
    float s;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"((float) d));
    double r = (double) s;
    r = fma(r, fma(-d, r, 1.0), r);
    r = fma(r, fma(-d, r, 1.0), r);
    return r;
}

// Truncated Fermi-Dirac transform (delta_tot_table.c:431-454):
//   J = sum_{n=1}^{19} -(-1)^n e^{-n qc} II(x,qc,n) / (n^2+x^2)^2 / (1.5 zeta(3) (1 - nufrac_low)),
//   II = (n^2 + n^3 qc + n qc x^2 - x^2) qc j0(qc x) + (2n + n^2 qc + qc x^2) cos(qc x).
// With D = n^2 + x^2 the two brackets are (n qc - 1) D + 2 n^2 and 2n + qc D, so a term is
//   e_n [ ((n qc - 1) qc j0 + qc cos) / D + 2n (n qc j0 + cos) / D^2 ]
// -- the same number (to rounding; the K2 kernel is bound by FP64 issue, and this form needs 7 FP64 instructions per term
// besides the reciprocal instead of 12).  jt: per-species table (jfrac_table_init): jt[2(n-1)] = (-1)^(n+1) e^{-n qc},
// jt[2(n-1)+1] = n qc - 1, jt[38] = 1 / (1.5 zeta(3) (1 - nufrac_low)).
constexpr int JT_DOUBLES = 40;
__device__ __forceinline__ void jfrac_table_init(double *jt, double qc, double nufrac_low)
{
    if (threadIdx.x < 19) {
        const int n = threadIdx.x + 1;
        jt[2 * (n - 1)] = ((n & 1) ? 1.0 : -1.0) * exp(-(double) n * qc);
        jt[2 * (n - 1) + 1] = (double) n * qc - 1.0;
    }
    if (threadIdx.x == 19) jt[38] = 1.0 / (1.5 * 1.202056903159594 * (1 - nufrac_low));
}

__device__ double Jfrac_high_d(double x, double qc, const double *__restrict__ jt)
{
    const double arg = qc * x, x2 = x * x;
    double sn, cs;
    sincos(arg, &sn, &cs);
    double j0;                                   // gsl_sf_bessel_j0: Taylor series below 0.5, sin(x)/x above
    if (fabs(arg) < 0.5) {
        const double y = arg * arg;
        j0 = 1.0 + y * (-1.0 / 6.0 + y * (1.0 / 120.0 + y * (-1.0 / 5040.0 + y * (1.0 / 362880.0 + y * (-1.0 / 39916800.0 + y * (1.0 / 6227020800.0))))));
    } else {
        j0 = sn * fast_rcp_d(arg);
    }
    const double qj = qc * j0, qcs = qc * cs;
    double integ = 0;
#pragma unroll
    for (int n = 1; n < 20; n++) {
        const double dn = (double) n;
        const double r = fast_rcp_d(dn * dn + x2);
        const double P = fma(jt[2 * n - 1], qj, qcs);
        const double Q = (2 * dn) * fma(dn, qj, cs);
        integ = fma(jt[2 * n - 2] * r, fma(Q, r, P), integ);
    }
    return integ * jt[38];
}

__device__ __forceinline__ double specialJ_d(double x, double qc, const double *__restrict__ jt)
{
    return qc > 0 ? Jfrac_high_d(x, qc, jt) : specialJ_fit_d(x);
}

// ---------------------------------------------------------------- natural cubic spline (GSL cspline.c)
// interval index with xa[i] <= x < xa[i+1] (last interval for x == xa[n-1])
__device__ __forceinline__ int bsearch_d(const double *xa, int n, double x)
{
    int lo = 0, hi = n - 1;
    while (hi > lo + 1) {
        const int m = (hi + lo) >> 1;
        if (xa[m] > x) hi = m; else lo = m;
    }
    return lo;
}

__device__ __forceinline__ double cspline_eval_d(const double *xa, const double *ya, const double *ca, int i, double x)
{
    const double x_lo = xa[i], x_hi = xa[i + 1], y_lo = ya[i], y_hi = ya[i + 1];
    const double dx = x_hi - x_lo, dy = y_hi - y_lo, delx = x - x_lo;
    const double c_i = ca[i], c_ip1 = ca[i + 1];
    const double b_i = (dy / dx) - dx * (c_ip1 + 2.0 * c_i) / 3.0;
    const double d_i = (c_ip1 - c_i) / (3.0 * dx);
    return y_lo + delx * (b_i + delx * (c_i + delx * d_i));
}

// Polynomial coefficients of segment i exactly as gsl cspline_eval derives them at every call; done once instead.
__device__ __forceinline__ void cspline_segment(const double *xa, const double *ya, const double *ca, int i, double &b, double &d)
{
    const double dx = xa[i + 1] - xa[i], dy = ya[i + 1] - ya[i];
    b = (dy / dx) - dx * (ca[i + 1] + 2.0 * ca[i]) / 3.0;
    d = (ca[i + 1] - ca[i]) / (3.0 * dx);
}

// LDL^T factors of the natural-spline system for knots xa[0..n): alpha[i], gamma[i], i < n-2
// (GSL linalg/tridiag.c solve_tridiag).  One thread.
__device__ void spline_factor_seq(const double *xa, int n, double *alpha, double *gamma)
{
    const int N = n - 2;
    if (N < 1) return;
    alpha[0] = 2.0 * ((xa[2] - xa[1]) + (xa[1] - xa[0]));
    if (N == 1) return;
    gamma[0] = (xa[2] - xa[1]) / alpha[0];
    for (int i = 1; i < N - 1; i++) {
        const double diag = 2.0 * ((xa[i + 2] - xa[i + 1]) + (xa[i + 1] - xa[i]));
        const double off_prev = xa[i + 1] - xa[i];
        alpha[i] = diag - off_prev * gamma[i - 1];
        gamma[i] = (xa[i + 2] - xa[i + 1]) / alpha[i];
    }
    {
        const int i = N - 1;
        const double diag = 2.0 * ((xa[i + 2] - xa[i + 1]) + (xa[i + 1] - xa[i]));
        alpha[i] = diag - (xa[i + 1] - xa[i]) * gamma[i - 1];
    }
}

__device__ __forceinline__ double spline_rhs(const double *xa, const double *ya, int i)
{
    const double h_i = xa[i + 1] - xa[i], h_ip1 = xa[i + 2] - xa[i + 1];
    const double g_i = (h_i != 0.0) ? 1.0 / h_i : 0.0, g_ip1 = (h_ip1 != 0.0) ? 1.0 / h_ip1 : 0.0;
    return 3.0 * ((ya[i + 2] - ya[i + 1]) * g_ip1 - (ya[i + 1] - ya[i]) * g_i);
}

// Solve for c[0..n) given factors; rhs in c[1..n-2] on entry (in place).  One thread.
__device__ void spline_solve_seq(int n, const double *alpha, const double *gamma, double *c)
{
    const int N = n - 2;
    c[0] = 0.0;
    c[n - 1] = 0.0;
    if (N < 1) return;
    double *z = c + 1;
    if (N == 1) { z[0] = z[0] / alpha[0]; return; }
    for (int i = 1; i < N; i++) z[i] = z[i] - gamma[i - 1] * z[i - 1];
    for (int i = 0; i < N; i++) z[i] = z[i] / alpha[i];
    for (int i = N - 2; i >= 0; i--) z[i] = z[i] - gamma[i] * z[i + 1];
}

// The same solve by a whole CTA, for the per-bin delta_tot spline at the head of the K2 kernels (it sits on the critical path
// of every bin): identical operations in identical order -- the two recurrences stay on one thread, but with the factors
// staged in shared memory (`ga`, n doubles of scratch) so that no iteration waits for a global load, and the divisions in
// between, which are independent, are dealt to all threads.  rhs in c[1..n-2] on entry; ends with a barrier.
__device__ __forceinline__ void spline_solve_cta(int n, const double *__restrict__ alpha, const double *__restrict__ gamma,
                                                 double *__restrict__ c, double *__restrict__ ga)
{
    const int N = n - 2, tid = threadIdx.x, T = blockDim.x;
    if (N < 2) {
        if (tid == 0) spline_solve_seq(n, alpha, gamma, c);
        __syncthreads();
        return;
    }
    double *z = c + 1;
    for (int i = tid; i < N - 1; i += T) ga[i] = gamma[i];
    if (tid == 0) { c[0] = 0.0; c[n - 1] = 0.0; }
    __syncthreads();
    if (tid == 0) {
        double prev = z[0];
#pragma unroll 8
        for (int i = 1; i < N; i++) { prev = z[i] - ga[i - 1] * prev; z[i] = prev; }
    }
    __syncthreads();
    for (int i = tid; i < N; i += T) z[i] = z[i] / alpha[i];
    __syncthreads();
    if (tid == 0) {
        double next = z[N - 1];
#pragma unroll 8
        for (int i = N - 2; i >= 0; i--) { next = z[i] - ga[i] * next; z[i] = next; }
    }
    __syncthreads();
}

// ---------------------------------------------------------------- block-wide QK61 / QAG
struct QkOut { double result, abserr, resabs, resasc; };

struct QagShared {
    double a[QAG_LIMIT], b[QAG_LIMIT], r[QAG_LIMIT], e[QAG_LIMIT];
    double red[4][4];
    int sel;
};

__device__ __forceinline__ double rescale_error_d(double err, double result_abs, double result_asc)
{
    err = fabs(err);
    if (result_asc != 0 && err != 0) {
        const double t = 200 * err / result_asc;
        const double scale = t * sqrt(t);                 // pow(t, 1.5)
        err = scale < 1 ? result_asc * scale : result_asc;
    }
    if (result_abs > DBL_MIN / (50 * DBL_EPSILON)) {
        const double min_err = 50 * DBL_EPSILON * result_abs;
        if (min_err > err) err = min_err;
    }
    return err;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// Two 61-point rules at once: threads 0..63 take [a1,b1], threads 64..127 take [a2,b2].
// Every thread returns both results (uniform).  Must be called by all K2_THREADS threads.
template <class F>
__device__ void qk61_pair(const F &f, double a1, double b1, double a2, double b2, bool two,
                          QagShared &S, QkOut &o1, QkOut &o2)
{
    const int tid = threadIdx.x, half = tid >> 6, n = tid & 63, warp = tid >> 5;
    const double a = half ? a2 : a1, b = half ? b2 : b1;
    const double center = 0.5 * (a + b), half_length = 0.5 * (b - a);
    const bool active = n < 61 && (two || half == 0);
    const int j = n <= 30 ? n : 60 - n;
    double fv = 0.0, wk = 0.0, wgs = 0.0;
    if (active) {
        const double absc = half_length * c_xgk[j];
        fv = f(n <= 30 ? center - absc : center + absc);
        wk = c_wgk[j];
        wgs = (j & 1) ? c_wg[j >> 1] : 0.0;
    }
    const double sk = warp_sum(wk * fv), sg = warp_sum(wgs * fv), sa = warp_sum(wk * fabs(fv));
    if ((tid & 31) == 0) { S.red[warp][0] = sk; S.red[warp][1] = sg; S.red[warp][2] = sa; }
    __syncthreads();
    const double kron[2] = { S.red[0][0] + S.red[1][0], S.red[2][0] + S.red[3][0] };
    const double gaus[2] = { S.red[0][1] + S.red[1][1], S.red[2][1] + S.red[3][1] };
    const double rabs[2] = { S.red[0][2] + S.red[1][2], S.red[2][2] + S.red[3][2] };
    const double mean = kron[half] * 0.5;
    const double sc = warp_sum(active ? wk * fabs(fv - mean) : 0.0);
    if ((tid & 31) == 0) S.red[warp][3] = sc;
    __syncthreads();
    const double rasc[2] = { S.red[0][3] + S.red[1][3], S.red[2][3] + S.red[3][3] };
    __syncthreads();   // S.red is reused by the next call
    const double hl1 = 0.5 * (b1 - a1), hl2 = 0.5 * (b2 - a2);
    o1.result = kron[0] * hl1;
    o1.resabs = rabs[0] * fabs(hl1);
    o1.resasc = rasc[0] * fabs(hl1);
    o1.abserr = rescale_error_d((kron[0] - gaus[0]) * hl1, o1.resabs, o1.resasc);
    o2.result = kron[1] * hl2;
    o2.resabs = rabs[1] * fabs(hl2);
    o2.resasc = rasc[1] * fabs(hl2);
    o2.abserr = rescale_error_d((kron[1] - gaus[1]) * hl2, o2.resabs, o2.resasc);
}

__device__ __forceinline__ bool subinterval_too_small_d(double a1, double a2, double b2)
{
    const double tmp = (1 + 100 * DBL_EPSILON) * (fabs(a2) + 1000 * DBL_MIN);
    return fabs(a1) <= tmp && fabs(b2) <= tmp;
}

// gsl_integration_qag (key 6) for one integrand, executed by the whole CTA; all threads return the
// same values.  passes (optional) counts 61-point rule applications.
template <class F>
__device__ int qag61_block(const F &f, double a, double b, double epsabs, double epsrel, int limit,
                           QagShared &S, double *result, double *abserr, unsigned *passes)
{
    QkOut q0, qd;
    qk61_pair(f, a, b, a, b, false, S, q0, qd);
    unsigned np = 1;
    double tolerance = fmax(epsabs, epsrel * fabs(q0.result));
    const double round_off = 50 * DBL_EPSILON * q0.resabs;
    *result = q0.result;
    *abserr = q0.abserr;
    if (passes) *passes = np;
    if (q0.abserr <= round_off && q0.abserr > tolerance) return Q_EROUND;
    if ((q0.abserr <= tolerance && q0.abserr != q0.resasc) || q0.abserr == 0.0) return Q_OK;
    if (limit == 1) return Q_EMAXITER;
    if (threadIdx.x == 0) { S.a[0] = a; S.b[0] = b; S.r[0] = q0.result; S.e[0] = q0.abserr; S.sel = 0; }
    __syncthreads();
    double area = q0.result, errsum = q0.abserr;
    int size = 1, iteration = 1, roundoff_type1 = 0, roundoff_type2 = 0, error_type = 0;
    do {
        const int i = S.sel;
        const double a_i = S.a[i], b_i = S.b[i], r_i = S.r[i], e_i = S.e[i];
        const double a1 = a_i, b1 = 0.5 * (a_i + b_i), a2 = b1, b2 = b_i;
        QkOut q1, q2;
        qk61_pair(f, a1, b1, a2, b2, true, S, q1, q2);
        np += 2;
        const double area12 = q1.result + q2.result, error12 = q1.abserr + q2.abserr;
        errsum += (error12 - e_i);
        area += area12 - r_i;
        if (q1.resasc != q1.abserr && q2.resasc != q2.abserr) {
            const double delta = r_i - area12;
            if (fabs(delta) <= 1.0e-5 * fabs(area12) && error12 >= 0.99 * e_i) roundoff_type1++;
            if (iteration >= 10 && error12 > e_i) roundoff_type2++;
        }
        tolerance = fmax(epsabs, epsrel * fabs(area));
        if (errsum > tolerance) {
            if (roundoff_type1 >= 6 || roundoff_type2 >= 20) error_type = 2;
            if (subinterval_too_small_d(a1, a2, b2)) error_type = 3;
        }
        if (threadIdx.x == 0) {
            // GSL workspace update(): the half with the larger error keeps slot i, the other is appended
            if (q2.abserr > q1.abserr) {
                S.a[i] = a2; S.r[i] = q2.result; S.e[i] = q2.abserr;
                S.a[size] = a1; S.b[size] = b1; S.r[size] = q1.result; S.e[size] = q1.abserr;
            } else {
                S.b[i] = b1; S.r[i] = q1.result; S.e[i] = q1.abserr;
                S.a[size] = a2; S.b[size] = b2; S.r[size] = q2.result; S.e[size] = q2.abserr;
            }
            int best = 0;
            double emax = S.e[0];
            for (int k = 1; k <= size; k++) if (S.e[k] > emax) { emax = S.e[k]; best = k; }
            S.sel = best;
        }
        size++;
        iteration++;
        __syncthreads();
    } while (iteration < limit && !error_type && errsum > tolerance);
    double s = 0;
    for (int k = 0; k < size; k++) s += S.r[k];
    __syncthreads();
    *result = s;
    *abserr = errsum;
    if (passes) *passes = np;
    if (errsum <= tolerance) return Q_OK;
    if (error_type == 2) return Q_EROUND;
    if (error_type == 3) return Q_ESING;
    if (iteration == limit) return Q_EMAXITER;
    return Q_EFAILED;
}

// ---------------------------------------------------------------- free-streaming table
struct FsIntegrand {
    BgTable bg;
    // fslength_int, delta_tot_table.c:378-383: 1/a/(a H(a))
    __device__ double operator()(double loga) const { return bg_eval(bg, loga) * exp(-loga); }
};

// One CTA per requested lower limit: out[i] = light * int_{logai[i]}^{logaf} (delta_tot_table.c:394-407)
__global__ void __launch_bounds__(K2_THREADS)
fslength_kernel(BgTable bg, const double *__restrict__ logai, int n, double logaf, double light,
                double *__restrict__ out, int *__restrict__ status, unsigned long long *evals)
{
    __shared__ QagShared S;
    const int i = blockIdx.x;
    if (i >= n) return;
    const double lo = logai[i];
    if (lo >= logaf) {
        if (threadIdx.x == 0) { out[i] = 0.0; status[i] = Q_OK; }
        return;
    }
    FsIntegrand f{ bg };
    double res, err;
    unsigned passes = 0;
    const int st = qag61_block(f, lo, logaf, 0.0, 1e-6, QAG_LIMIT, S, &res, &err, &passes);
    if (threadIdx.x == 0) {
        out[i] = light * res;
        status[i] = st;
        if (evals) atomicAdd(evals, 61ull * passes);
    }
}

// knots of the free-streaming spline, delta_tot_table.c:582
__global__ void fs_knots_kernel(double loga0, double loga, int Nfs, double *__restrict__ fsscales)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < Nfs) fsscales[i] = loga0 + i * (loga - loga0) / (Nfs - 1.);
}

// Block 0: natural-spline coefficients of the free-streaming table (Nfs = 16 Na ~ 1600 knots).  The knots are
// equally spaced, so every recurrence of the L D L^T solve (pivots, forward and backward substitution) forgets its
// starting value geometrically (factor <= 0.27 per step): each thread owns a short chunk and warms up FS_WARM steps
// before it -- the values it stores equal the sequential solve's to ~1e-37, without 1600 dependent steps.
// Block 1, thread 0: L D L^T factors for the (unequally spaced, <= ~100) delta_tot knots, sequentially.
constexpr int FS_WARM = 64;
__global__ void __launch_bounds__(256)
k2_prep_splines_kernel(const double *__restrict__ fsscales, const double *__restrict__ fslengths, int Nfs,
                       double *__restrict__ fs_c, double *__restrict__ fs_b, double *__restrict__ fs_d,
                       double *__restrict__ alpha, double *__restrict__ gamma,
                       const double *__restrict__ scalefact, int Na, double *__restrict__ dt_alpha, double *__restrict__ dt_gamma)
{
    if (blockIdx.x == 1) {
        if (threadIdx.x == 0 && Na > 2) spline_factor_seq(scalefact, Na, dt_alpha, dt_gamma);
        return;
    }
    const int M = Nfs - 2;                       // interior unknowns c[1..Nfs-2]
    const int T = blockDim.x, t = threadIdx.x;
    const int chunk = (M + T - 1) / T;
    if (chunk > 16) {                            // far more knots than any namax gives: plain sequential solve
        if (t == 0) {
            spline_factor_seq(fsscales, Nfs, alpha, gamma);
            for (int i = 0; i < M; i++) fs_c[i + 1] = spline_rhs(fsscales, fslengths, i);
            spline_solve_seq(Nfs, alpha, gamma, fs_c);
            for (int i = 0; i < Nfs - 1; i++) cspline_segment(fsscales, fslengths, fs_c, i, fs_b[i], fs_d[i]);
        }
        return;
    }
    const int lo = min(M, t * chunk), hi = min(M, lo + chunk);
    double *z = fs_c + 1;
    auto diag = [&](int i) { return 2.0 * ((fsscales[i + 2] - fsscales[i + 1]) + (fsscales[i + 1] - fsscales[i])); };
    auto off = [&](int i) { return fsscales[i + 2] - fsscales[i + 1]; };
    if (t == 0) { fs_c[0] = 0.0; fs_c[Nfs - 1] = 0.0; }
    // pivots alpha_i = diag_i - off_{i-1} gamma_{i-1}, gamma_i = off_i / alpha_i
    if (lo < hi) {
        int s = max(0, lo - FS_WARM);
        double a = diag(s), g = off(s) / a;
        for (int i = s + 1; i < hi; i++) {
            if (i - 1 >= lo) { alpha[i - 1] = a; gamma[i - 1] = g; }
            a = diag(i) - off(i - 1) * g;
            g = off(i) / a;
        }
        alpha[hi - 1] = a; gamma[hi - 1] = g;
    }
    __syncthreads();
    // forward substitution z_i = rhs_i - gamma_{i-1} z_{i-1}
    double zkeep[16];
    if (lo < hi) {
        int s = max(0, lo - FS_WARM);
        double v = spline_rhs(fsscales, fslengths, s);
        for (int i = s + 1; i < hi; i++) {
            if (i - 1 >= lo) zkeep[i - 1 - lo] = v;
            v = spline_rhs(fsscales, fslengths, i) - gamma[i - 1] * v;
        }
        zkeep[hi - 1 - lo] = v;
    }
    __syncthreads();
    for (int i = lo; i < hi; i++) z[i] = zkeep[i - lo] / alpha[i];       // D^-1
    __syncthreads();
    // backward substitution x_i = c_i - gamma_i x_{i+1}
    if (lo < hi) {
        int e = min(M - 1, hi - 1 + FS_WARM);
        double v = z[e];
        for (int i = e - 1; i >= lo; i--) {
            if (i + 1 < hi) zkeep[i + 1 - lo] = v;
            v = z[i] - gamma[i] * v;
        }
        zkeep[0] = v;
    }
    __syncthreads();
    for (int i = lo; i < hi; i++) z[i] = zkeep[i - lo];
    __syncthreads();
    for (int i = t; i < Nfs - 1; i += T) cspline_segment(fsscales, fslengths, fs_c, i, fs_b[i], fs_d[i]);
}

// ---------------------------------------------------------------- the per-k integral
struct K2Dev {
    int nk, Na, namax, Nfs;
    int k_first;            // this launch integrates bins [k_first, k_first + gridDim.x)
    double loga0, loga, light, delta_nu_prefac, deriv_prefac, nufrac_low0;
    double mnubykT[3], qc[3], relerr[3];
    int integrate[3];
    const double *scalefact, *delta_tot, *wavenum, *delta_nu_init;   // device
    const double *fsscales, *fslengths, *fs_c, *fs_b, *fs_d, *dt_alpha, *dt_gamma; // device
    double *out;
    int *status;
    unsigned long long *evals;
    BgTable bg;
};

struct DeltaNuIntegrand {
    // get_delta_nu_int, delta_tot_table.c:492-500
    const K2Dev *P;
    const double *sx, *sy, *sc, *sb, *sd;   // shared: knots, delta_tot row, spline c and per-segment b, d
    double k, mnubykT, qc, fs_x0, fs_inv_dx;
    const double *jt;             // shared: the species' table for Jfrac_high_d (jfrac_table_init)
    __device__ double operator()(double logai) const
    {
        const K2Dev &p = *P;
        // free-streaming length: near-uniform knots -> direct index, then fix against the stored knots
        int i = (int) ((logai - fs_x0) * fs_inv_dx);
        i = max(0, min(i, p.Nfs - 2));
        while (i > 0 && logai < __ldg(p.fsscales + i)) i--;
        while (i < p.Nfs - 2 && logai >= __ldg(p.fsscales + i + 1)) i++;
        const double dfs = logai - __ldg(p.fsscales + i);
        const double fsl = __ldg(p.fslengths + i) + dfs * (__ldg(p.fs_b + i) + dfs * (__ldg(p.fs_c + i) + dfs * __ldg(p.fs_d + i)));
        double dtot;
        if (p.Na > 2) {
            const int m = bsearch_d(sx, p.Na, logai);
            const double dt = logai - sx[m];
            dtot = sy[m] + dt * (sb[m] + dt * (sc[m] + dt * sd[m]));
        } else {
            dtot = sy[0] + (logai - sx[0]) / (sx[1] - sx[0]) * (sy[1] - sy[0]);
        }
        const double specJ = specialJ_d(k * fsl / mnubykT, qc, jt);
        return fsl * bg_eval(p.bg, logai) * specJ * dtot;
    }
};

__global__ void __launch_bounds__(K2_THREADS, 6)
k2_delta_nu_kernel(const __grid_constant__ K2Dev p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ QagShared S;
    double *sx = (double *) smem_raw, *sy = sx + p.Na, *sc = sy + p.Na, *sb = sc + p.Na, *sd = sb + p.Na;
    __shared__ double jt_s[JT_DOUBLES];
    const int ik = p.k_first + blockIdx.x, sp = blockIdx.y;
    jfrac_table_init(jt_s, p.qc[sp], p.nufrac_low0);
    for (int i = threadIdx.x; i < p.Na; i += blockDim.x) {
        sx[i] = p.scalefact[i];
        sy[i] = p.delta_tot[(size_t) ik * p.namax + i];
    }
    __syncthreads();
    const double k = p.wavenum[ik], mnubykT = p.mnubykT[sp];
    // initial-condition term, delta_tot_table.c:525-535 (always the untruncated fit: qc is still 0 there)
    const double fsl_A0a = p.fslengths[0];
    const double specJ0 = specialJ_fit_d(k * fsl_A0a / (mnubykT > 0 ? mnubykT : 1));
    double dnu = specJ0 * p.delta_nu_init[ik] * (1. + p.deriv_prefac * fsl_A0a);
    int st = Q_OK;
    unsigned passes = 0;
    if (p.integrate[sp]) {
        if (p.Na > 2) {
            for (int i = threadIdx.x; i < p.Na - 2; i += blockDim.x) sc[i + 1] = spline_rhs(sx, sy, i);
            __syncthreads();
            spline_solve_cta(p.Na, p.dt_alpha, p.dt_gamma, sc, sb);          // (sb is free until the segments are formed)
            for (int i = threadIdx.x; i < p.Na - 1; i += blockDim.x) cspline_segment(sx, sy, sc, i, sb[i], sd[i]);
            __syncthreads();
        }
        DeltaNuIntegrand f;
        f.P = &p; f.sx = sx; f.sy = sy; f.sc = sc; f.sb = sb; f.sd = sd;
        f.k = k; f.mnubykT = mnubykT; f.qc = p.qc[sp]; f.jt = jt_s;
        f.fs_x0 = p.loga0;
        f.fs_inv_dx = (p.Nfs - 1.) / (p.loga - p.loga0);
        double res, err;
        st = qag61_block(f, p.loga0, p.loga, 0.0, p.relerr[sp], QAG_LIMIT, S, &res, &err, &passes);
        dnu += p.delta_nu_prefac * res;
    }
    if (threadIdx.x == 0) {
        p.out[(size_t) sp * p.nk + ik] = dnu;
        // low byte: GSL-style status; bits 8-19: 61-point rule applications; bits 20+: passes through the integrand
        p.status[(size_t) sp * p.nk + ik] = st | ((int) passes << 8) | ((int) ((passes + 1) / 2) << 20);
        if (p.evals && passes) atomicAdd(p.evals, 61ull * passes);
    }
}

// ---------------------------------------------------------------- the per-k integral, bisections integrated ahead of time
// K2's time is the critical path of its deepest bin: up to 57 SEQUENTIAL bisections of ~5 us each with hybrid neutrinos.
// A bisection's values depend only on the interval, so a CTA of 128 M threads integrates the halves of the M worst
// intervals at once and warp 0 then replays QAG's loop over the cached results (ksn_qag_spec.h): same decisions, same
// sums in the same order, hence bit-identical output -- with up to M loop trips per pass through the integrand.
template <int M>
struct SpecShared {
    QagSpecList L;
    double red[4 * M][4];      // per warp: Kronrod, Gauss, |f|, |f - mean| partial sums
    double q0[4];              // the rule on the whole interval: result, abserr, resabs, resasc
    double result, abserr;
    int sel[M];                // slots whose halves the next pass integrates
    int nsel, done, status;
    unsigned passes, rules;    // rule applications of the sequential algorithm / actually computed
    unsigned trips;            // passes through the integrand (the CTA's critical path)
};

// 2M groups of 64 threads: group g integrates half (g & 1) of slot sel[g >> 1]; `whole`: group 0 integrates slot 0 itself
// and groups 1 and 2 its two halves (the whole interval is bisected in all but trivial cases: one trip saved per bin).
// Same arithmetic, same reduction tree as qk61_pair.  Must be called by all 128 M threads; ends with a barrier.
template <int M, class F>
__device__ void qk61_groups(const F &f, SpecShared<M> &S, bool whole)
{
    const int tid = threadIdx.x, g = tid >> 6, n = tid & 63, warp = tid >> 5;
    const int which = whole ? 0 : g >> 1, child = whole ? g - 1 : g & 1;
    const bool gactive = whole ? g <= 2 : which < S.nsel;
    const bool parent = whole && g == 0;
    int slot = 0;
    double a = 0.0, b = 0.0;
    if (gactive) {
        slot = whole ? 0 : S.sel[which];
        const double a_i = S.L.a[slot], b_i = S.L.b[slot];
        if (parent) {
            a = a_i; b = b_i;
        } else {
            const double mid = 0.5 * (a_i + b_i);
            a = child ? mid : a_i;
            b = child ? b_i : mid;
        }
    }
    const double center = 0.5 * (a + b), half_length = 0.5 * (b - a);
    const bool active = gactive && n < 61;
    const int j = n <= 30 ? n : 60 - n;
    double fv = 0.0, wk = 0.0, wgs = 0.0;
    if (active) {
        const double absc = half_length * c_xgk[j];
        fv = f(n <= 30 ? center - absc : center + absc);
        wk = c_wgk[j];
        wgs = (j & 1) ? c_wg[j >> 1] : 0.0;
    }
    const double sk = warp_sum(wk * fv), sg = warp_sum(wgs * fv), sa = warp_sum(wk * fabs(fv));
    if ((tid & 31) == 0) { S.red[warp][0] = sk; S.red[warp][1] = sg; S.red[warp][2] = sa; }
    __syncthreads();
    const double kron = S.red[2 * g][0] + S.red[2 * g + 1][0];
    const double gaus = S.red[2 * g][1] + S.red[2 * g + 1][1];
    const double rabs = S.red[2 * g][2] + S.red[2 * g + 1][2];
    const double mean = kron * 0.5;
    const double sc = warp_sum(active ? wk * fabs(fv - mean) : 0.0);
    if ((tid & 31) == 0) S.red[warp][3] = sc;
    __syncthreads();
    if (gactive && n == 0) {
        const double rasc = S.red[2 * g][3] + S.red[2 * g + 1][3];
        QkOut o;
        o.result = kron * half_length;
        o.resabs = rabs * fabs(half_length);
        o.resasc = rasc * fabs(half_length);
        o.abserr = rescale_error_d((kron - gaus) * half_length, o.resabs, o.resasc);
        if (parent) {
            S.q0[0] = o.result; S.q0[1] = o.abserr; S.q0[2] = o.resabs; S.q0[3] = o.resasc;
        } else {
            S.L.cr[2 * slot + child] = o.result;
            S.L.ce[2 * slot + child] = o.abserr;
            S.L.cf[2 * slot + child] = o.resasc != o.abserr;
        }
    }
    __syncthreads();
}

// Slot with the largest error estimate, the lowest index among equals (what a sequential scan with '>' finds).
// skip != nullptr: only slots with skip[k] == 0 and e[k] > floor; -1 if there is none.  One warp, all lanes.
__device__ __forceinline__ int warp_argmax_err(const double *e, const unsigned char *skip, double floor_, int size, int lane)
{
    int best = -1;
    double bv = 0.0;
    for (int k = lane; k < size; k += 32) {
        if (skip && (skip[k] || !(e[k] > floor_))) continue;
        const double v = e[k];
        if (best < 0 || v > bv) { bv = v; best = k; }
    }
    // Error estimates are non-negative, so they order like their bit patterns: the warp's maximum in two 32-bit redux steps
    // (high word, then the low word among the lanes that hold it), the lowest index among its holders in a third -- three
    // instructions instead of five rounds of two 64-bit shuffles each, on the serial path of every trip.
    const unsigned long long key = best >= 0 ? (unsigned long long) __double_as_longlong(bv) + 1ull : 0ull;   // 0: no candidate
    const unsigned hi = (unsigned) (key >> 32), lo = (unsigned) key;
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    const unsigned cand = (best >= 0 && hi == mh && lo == ml) ? (unsigned) best : 0xffffffffu;
    const unsigned win = __reduce_min_sync(0xffffffffu, cand);
    return win == 0xffffffffu ? -1 : (int) win;
}

#ifdef KSN_K2_TRACE
// Developer build (make EXTRA_NVFLAGS=-DKSN_K2_TRACE; tools/k2_trace.py): where the CTA of the deepest bin -- the kernel's
// critical path -- spends its cycles.  [0] integrand passes, [1] replay of QAG's loop, [2] passes, [3] set-up before the
// quadrature, [4] whole CTA, [5..6] its start/end (globaltimer ns), [7] end of the last CTA, [8] start of the first.
__device__ unsigned long long g_k2_trace[16];
#define K2_TRACE(...) __VA_ARGS__
#else
#define K2_TRACE(...)
#endif

// gsl_integration_qag (key 6) by a CTA of 128 M threads; all threads return the same values.
template <int M, class F>
__device__ int qag61_spec(const F &f, double a, double b, double epsabs, double epsrel, int limit,
                          SpecShared<M> &S, double *result, double *abserr, unsigned *passes, unsigned *rules, unsigned *trips)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { S.L.a[0] = a; S.L.b[0] = b; S.nsel = 0; S.done = 0; }
    __syncthreads();
    qk61_groups<M>(f, S, true);           // the whole interval and, ahead of time, its two halves
    QagSpecState s;                       // lives in lane 0 of warp 0
    unsigned nrules = 3, ntrips = 1;
    int size = 1;                         // uniform in warp 0
    if (warp == 0) {
        if (lane == 0) {
            int st = QAGS_OK;
            if (qags_begin(s, S.L, a, b, epsabs, epsrel, limit, S.q0[0], S.q0[1], S.q0[2], S.q0[3], &st)) {
                S.result = S.q0[0]; S.abserr = S.q0[1]; S.status = st; S.passes = 1; S.rules = 3; S.trips = 1; S.done = 1;
            } else {
                S.L.cached[0] = 1; S.sel[0] = 0; S.nsel = 0;
            }
        }
    }
    bool have_children = true;            // uniform in the CTA: the halves the replay starts with are in the cache already
    K2_TRACE(const bool tr = blockIdx.x == 0 && blockIdx.y == 0 && tid == 0; long long t0 = clock64(), t_eval = 0, t_replay = 0, n_eval = 0;)
    for (;;) {
        __syncthreads();                  // selection (or the verdict) of warp 0 is visible
        K2_TRACE(if (tr) { const long long t = clock64(); t_replay += t - t0; t0 = t; })
        if (S.done) break;
        if (!have_children) {
            qk61_groups<M>(f, S, false);
            K2_TRACE(if (tr) { const long long t = clock64(); t_eval += t - t0; t0 = t; n_eval++; })
        }
        if (warp != 0) { have_children = false; continue; }
        if (!have_children) { nrules += 2 * S.nsel; ntrips++; }
        have_children = false;
        // replay QAG's loop over the cached halves
        for (;;) {
            const int i = warp_argmax_err(S.L.e, nullptr, 0.0, size, lane);
            __syncwarp();                 // every lane has read the list before lane 0 changes it
            int act = 1;                  // 0: one trip made, go on; 1: slot i must be integrated first; 2: finished
            double floor_ = 0.0;
            if (lane == 0) {
                if (S.L.cached[i]) act = qags_apply(s, S.L, i) ? 0 : 2;
                else floor_ = qags_spec_threshold(s);
            }
            act = __shfl_sync(0xffffffffu, act, 0);
            __syncwarp();                 // list updates of lane 0 are visible to the scans below
            if (act != 1) size++;
            if (act == 0) continue;
            if (act == 2) {
                if (lane == 0) {
                    double res, err;
                    S.status = qags_finish(s, S.L, &res, &err);
                    S.result = res; S.abserr = err; S.passes = s.passes; S.rules = nrules; S.trips = ntrips; S.done = 1;
                }
                break;
            }
            // next pass: slot i, and the worst of the other slots the loop is likely to reach
            floor_ = __shfl_sync(0xffffffffu, floor_, 0);
            if (lane == 0) { S.L.cached[i] = 1; S.sel[0] = i; }
            int ns = 1;
            for (; ns < M; ns++) {
                __syncwarp();
                const int jn = warp_argmax_err(S.L.e, S.L.cached, floor_, size, lane);
                __syncwarp();
                if (jn < 0) break;
                if (lane == 0) { S.L.cached[jn] = 1; S.sel[ns] = jn; }
            }
            if (lane == 0) S.nsel = ns;
            break;
        }
    }
    K2_TRACE(if (tr) { g_k2_trace[0] = t_eval; g_k2_trace[1] = t_replay; g_k2_trace[2] = n_eval; })
    *result = S.result;
    *abserr = S.abserr;
    if (passes) *passes = S.passes;
    if (rules) *rules = S.rules;
    if (trips) *trips = S.trips;
    const int st = S.status;
    __syncthreads();
    return st;
}

template <int M, int MINB>
__global__ void __launch_bounds__(K2_THREADS * M, MINB)
k2_delta_nu_spec_kernel(const __grid_constant__ K2Dev p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ SpecShared<M> S;
    double *sx = (double *) smem_raw, *sy = sx + p.Na, *sc = sy + p.Na, *sb = sc + p.Na, *sd = sb + p.Na;
    __shared__ double jt_s[JT_DOUBLES];
    const int ik = p.k_first + (int) (gridDim.x - 1 - blockIdx.x);      // deepest (highest-k) bins are scheduled first
    const int sp = blockIdx.y;
    K2_TRACE(const long long tk0 = clock64(); unsigned long long gt0; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt0));)
    jfrac_table_init(jt_s, p.qc[sp], p.nufrac_low0);
    for (int i = threadIdx.x; i < p.Na; i += blockDim.x) {
        sx[i] = p.scalefact[i];
        sy[i] = p.delta_tot[(size_t) ik * p.namax + i];
    }
    __syncthreads();
    const double k = p.wavenum[ik], mnubykT = p.mnubykT[sp];
    const double fsl_A0a = p.fslengths[0];
    const double specJ0 = specialJ_fit_d(k * fsl_A0a / (mnubykT > 0 ? mnubykT : 1));
    double dnu = specJ0 * p.delta_nu_init[ik] * (1. + p.deriv_prefac * fsl_A0a);
    int st = Q_OK;
    unsigned passes = 0, rules = 0, trips = 0;
    if (p.integrate[sp]) {
        if (p.Na > 2) {
            for (int i = threadIdx.x; i < p.Na - 2; i += blockDim.x) sc[i + 1] = spline_rhs(sx, sy, i);
            __syncthreads();
            spline_solve_cta(p.Na, p.dt_alpha, p.dt_gamma, sc, sb);          // (sb is free until the segments are formed)
            for (int i = threadIdx.x; i < p.Na - 1; i += blockDim.x) cspline_segment(sx, sy, sc, i, sb[i], sd[i]);
            __syncthreads();
        }
        DeltaNuIntegrand f;
        f.P = &p; f.sx = sx; f.sy = sy; f.sc = sc; f.sb = sb; f.sd = sd;
        f.k = k; f.mnubykT = mnubykT; f.qc = p.qc[sp]; f.jt = jt_s;
        f.fs_x0 = p.loga0;
        f.fs_inv_dx = (p.Nfs - 1.) / (p.loga - p.loga0);
        double res, err;
        K2_TRACE(if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) g_k2_trace[3] = clock64() - tk0;)
        st = qag61_spec<M>(f, p.loga0, p.loga, 0.0, p.relerr[sp], QAG_LIMIT, S, &res, &err, &passes, &rules, &trips);
        dnu += p.delta_nu_prefac * res;
    }
    if (threadIdx.x == 0) {
#ifdef KSN_K2_TRACE
        unsigned long long gt1;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt1));
        if (blockIdx.x == 0 && blockIdx.y == 0) { g_k2_trace[4] = clock64() - tk0; g_k2_trace[5] = gt0; g_k2_trace[6] = gt1; }
        atomicMax(&g_k2_trace[7], gt1);
        atomicMin(&g_k2_trace[8], gt0);
#endif
        p.out[(size_t) sp * p.nk + ik] = dnu;
        // bits 8-19: the SEQUENTIAL algorithm's rule count, as k2_delta_nu_kernel reports it; bits 20+: passes through the integrand
        p.status[(size_t) sp * p.nk + ik] = st | ((int) passes << 8) | ((int) trips << 20);
        if (p.evals && rules) atomicAdd(p.evals, 61ull * rules);          // integrand evaluations actually made
    }
}

static BgPatch g_bg_patch[BG_MAX_PATCH];
static int g_bg_npatch = 0, g_bg_flagged = 0;

static BgTable bg_table()
{
    Ctx &c = ctx();
    BgTable t;
    t.g = c.d_bg; t.n = c.bg_n; t.lo = c.bg_lo; t.h = c.bg_h; t.inv_h = 1.0 / c.bg_h;
    t.npatch = g_bg_npatch;
    for (int p = 0; p < BG_MAX_PATCH; p++) t.patch[p] = g_bg_patch[p];
    return t;
}

// carve n doubles out of a bump allocator over c.d_k2
struct Bump {
    char *base; size_t off = 0;
    template <typename T> T *take(size_t n) { off = (off + 15) & ~(size_t) 15; T *p = (T *) (base + off); off += n * sizeof(T); return p; }
};

}  // namespace ksn

using namespace ksn;

// How many intervals a K2 CTA bisects per pass through the integrand (k2_delta_nu_spec_kernel<M>); 1 = the plain
// sequential kernel.  KSN_K2_SPEC overrides the default.
// KSN_K2_SPEC_DEFAULT 0 = by regime (k2_spec_auto).
static int k2_spec_width(void)
{
    const char *env = getenv("KSN_K2_SPEC");
    const int m = env ? atoi(env) : KSN_K2_SPEC_DEFAULT;
    return m >= 0 && m <= 4 ? m : 1;
}
extern "C" int ksn_k2_spec_width(void) { return k2_spec_width(); }
#ifdef KSN_K2_TRACE
extern "C" int ksn_k2_trace(unsigned long long *out16, int reset)
{
    if (out16) cudaMemcpyFromSymbol(out16, ksn::g_k2_trace, 16 * sizeof(unsigned long long));
    if (reset) { unsigned long long z[16] = {}; z[8] = ~0ull; cudaMemcpyToSymbol(ksn::g_k2_trace, z, sizeof z); }
    return 0;
}
#endif

// Measured (profiles/r1_k2_widths.txt, nk = 788, Na = 99): integrating ahead pays where the kernel's time is the critical
// path of a few deep bins -- hybrid neutrinos, one species: 57 sequential bisections on the highest-k bins -- and costs
// where the launch is throughput-bound (no hybrid cut: <= 8 bisections per bin; three species: 3 nk CTAs).
static int k2_spec_auto(int nspecies_launched, const double *qc)
{
    bool hybrid = false;
    for (int s = 0; s < nspecies_launched; s++) hybrid |= qc[s] > 0;
    return hybrid && nspecies_launched == 1 ? 3 : 1;
}

static unsigned long long g_last_evals = 0;
static int g_last_prefetch_hit = 0;
extern "C" int ksn_last_k2_prefetch_used(void) { return g_last_prefetch_hit; }
static unsigned g_max_passes = 0, g_max_trips = 0;
extern "C" unsigned ksn_last_k2_max_passes(void) { return g_max_passes; }
extern "C" unsigned ksn_last_k2_max_trips(void) { return g_max_trips; }
extern "C" unsigned long long ksn_last_k2_evals(void) { return g_last_evals; }

// host mirror of bg_lagrange (same formula), used to validate the table against the host function
static double bg_lagrange_host(const double *g, int n, double s)
{
    int i = (int) floor(s);
    i = i < 1 ? 1 : (i > n - 3 ? n - 3 : i);
    const double u = s - i, um1 = u - 1.0, up1 = u + 1.0, um2 = u - 2.0;
    return -u * um1 * um2 * (1.0 / 6.0) * g[i - 1] + up1 * um1 * um2 * 0.5 * g[i] - up1 * u * um2 * 0.5 * g[i + 1] + up1 * u * um1 * (1.0 / 6.0) * g[i + 2];
}

extern "C" int ksn_background_loaded(void) { return ctx().inited && ctx().d_bg != nullptr; }

extern "C" int ksn_background_info(int *npatch, int *flagged_cells)
{
    if (npatch) *npatch = g_bg_npatch;
    if (flagged_cells) *flagged_cells = g_bg_flagged;
    return KSN_OK;
}

extern "C" int ksn_set_background(ksn_hubble_fn hub, void *user, double loga_lo, double loga_hi, int n)
{
    int rc = ensure_init();
    if (rc) return rc;
    if (!hub || !(loga_hi > loga_lo) || n < 16) return set_error(KSN_EINVAL, "ksn_set_background: bad arguments");
    Ctx &c = ctx();
    const double h = (loga_hi - loga_lo) / (n - 1);
    auto sample = [&](double x) { const double a = exp(x); return 1.0 / (a * hub(a, user)); };
    const size_t cap = (size_t) n + (size_t) BG_MAX_PATCH_CELLS * BG_SUB + 4 * BG_MAX_PATCH;
    double *tab = (double *) malloc(sizeof(double) * cap);
    unsigned char *bad = (unsigned char *) calloc(n, 1);
    if (!tab || !bad) { free(tab); free(bad); return set_error(KSN_ENOMEM, "ksn_set_background: out of host memory"); }
    for (int i = 0; i < n; i++) tab[i] = sample(loga_lo + i * h);
    // validate every cell whose stencil is centred (the integrand never leaves [lo+2h, hi-2h]) at its midpoint
    const bool patches_on = !getenv("KSN_BG_NOPATCH");
    int flagged = 0;
    for (int i = 1; patches_on && i < n - 2; i++) {
        const double want = sample(loga_lo + (i + 0.5) * h), got = bg_lagrange_host(tab, n, i + 0.5);
        // 1e-13: above the ~1e-14 the interpolant loses at the knots of a cubic-spline Omega_nu table (third-derivative
        // jumps), far below the ~6e-9 it loses at a slope discontinuity
        if (!(fabs(got - want) <= 1e-13 * fabs(want))) { bad[i] = 1; flagged++; }
    }
    // cells next to a failing one share its stencil nodes: cover them too, then merge into ranges
    int npatch = 0, cells = 0;
    size_t used = n;
    bool overflow = false;
    if (flagged) {
        unsigned char *cover = (unsigned char *) calloc(n, 1);
        for (int i = 1; i < n - 2; i++)
            if (bad[i]) for (int d = -2; d <= 2; d++) if (i + d >= 1 && i + d < n - 2) cover[i + d] = 1;
        for (int i = 1; i < n - 2 && !overflow; i++) {
            if (!cover[i]) continue;
            int j = i;
            while (j + 1 < n - 2 && cover[j + 1]) j++;
            const int nc = j - i + 1;
            if (npatch == BG_MAX_PATCH || cells + nc > BG_MAX_PATCH_CELLS) { overflow = true; break; }
            BgPatch &P = g_bg_patch[npatch];
            const double hs = h / BG_SUB;
            P.lo = loga_lo + i * h; P.hi = loga_lo + (j + 1) * h;
            P.x0 = P.lo - hs; P.inv_h = 1.0 / hs;
            P.off = (int) used; P.n = nc * BG_SUB + 3;
            for (int k = 0; k < P.n; k++) tab[used + k] = sample(P.x0 + k * hs);
            used += P.n;
            cells += nc;
            npatch++;
            i = j;
        }
        free(cover);
    }
    free(bad);
    if (overflow) { npatch = 0; used = n; }   // a host function that is rough everywhere: plain table (first-order near its kinks)
    g_bg_npatch = npatch;
    g_bg_flagged = flagged;
    if (c.d_bg) { cudaFree(c.d_bg); c.d_bg = nullptr; }
    cudaError_t e = cudaMalloc((void **) &c.d_bg, sizeof(double) * used);
    if (e == cudaSuccess) e = cudaMemcpy(c.d_bg, tab, sizeof(double) * used, cudaMemcpyHostToDevice);
    free(tab);
    KSN_CUDA(e);
    c.bg_n = n; c.bg_lo = loga_lo; c.bg_hi = loga_hi; c.bg_h = h;
    return KSN_OK;
}

extern "C" int ksn_fslength_device(const double *logai, int n, double logaf, double light, double *out)
{
    int rc = ensure_init();
    if (rc) return rc;
    Ctx &c = ctx();
    if (!c.d_bg) return set_error(KSN_EINVAL, "ksn_fslength_device: call ksn_set_background first");
    if (!logai || !out || n < 1) return set_error(KSN_EINVAL, "ksn_fslength_device: bad arguments");
    for (int i = 0; i < n; i++)
        if (logai[i] < c.bg_lo + 2 * c.bg_h || logaf > c.bg_hi - 2 * c.bg_h)
            return set_error(KSN_EINVAL, "ksn_fslength_device: range outside the background table");
    rc = ensure_device_buffer((void **) &c.d_k2, &c.k2_cap, (size_t) n * 24 + 64);
    if (rc) return rc;
    Bump bp{ (char *) c.d_k2 };
    double *d_in = bp.take<double>(n), *d_out = bp.take<double>(n);
    int *d_st = bp.take<int>(n);
    KSN_CUDA(cudaMemcpyAsync(d_in, logai, sizeof(double) * n, cudaMemcpyHostToDevice, c.stream));
    fslength_kernel<<<n, K2_THREADS, 0, c.stream>>>(bg_table(), d_in, n, logaf, light, d_out, d_st, nullptr);
    c.launches++;
    KSN_CUDA(cudaGetLastError());
    int *st = (int *) malloc(sizeof(int) * n);
    KSN_CUDA(cudaMemcpyAsync(out, d_out, sizeof(double) * n, cudaMemcpyDeviceToHost, c.stream));
    KSN_CUDA(cudaMemcpyAsync(st, d_st, sizeof(int) * n, cudaMemcpyDeviceToHost, c.stream));
    KSN_CUDA(cudaStreamSynchronize(c.stream));
    for (int i = 0; i < n; i++)
        if (st[i]) { rc = set_error(KSN_EQUAD, "fslength quadrature %d failed with GSL-style code %d", i, st[i]); break; }
    free(st);
    return rc;
}

// ---------------------------------------------------------------- the part of K2 that depends on `a` and the stored knots only
// fs_knots + fslength (16 Na quadratures of 1/(a^2 H)) + the spline solves of k2_prep_splines need neither the power
// spectrum of this step nor the delta_tot rows: a host that knows the scale factor of the coming step (the PM hook does:
// it is an argument of add_nu_power_to_rhogrid) can have them computed on a side stream WHILE K1 sweeps the grid.
// ksn_delta_nu_integrate uses the prefetched tables if -- and only if -- every input they depend on is bit-for-bit what
// it is called with (a, TimeTransfer, light, the Na knots, the background table); otherwise it computes them itself.
struct K2Prefetch {
    bool valid = false;          // tables for (a, a0, light, sf[0..Na)) are on the device (or on their way: `done`)
    bool pending = false;        // a request has been recorded and not launched yet
    double a = 0, a0 = 0, light = 0;
    int Na = 0, namax = 0;
    double *sf = nullptr; int sf_cap = 0;     // host copy of the knots the tables were built for
    const double *bg = nullptr; int bg_n = 0; double bg_lo = 0, bg_h = 0; int bg_npatch = 0;
    double *d_buf = nullptr; size_t cap = 0;  // device: knots | fsscales | fslengths | fs_c | sa | sg | fs_b | fs_d | dta | dtg | evals | status
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr, main_idle = nullptr;
    double *h_sf = nullptr; size_t h_cap = 0; // pinned staging of the knots
};
static K2Prefetch g_pre;

struct K2PreLayout { double *knots, *fsscales, *fslengths, *fsc, *sa, *sg, *fsb, *fsd, *dta, *dtg; unsigned long long *evals; int *status; size_t bytes; };
static K2PreLayout k2_pre_layout(double *base, int Na)
{
    const size_t Nfs = (size_t) 16 * Na;
    Bump bp{ (char *) base };
    K2PreLayout l;
    l.knots = bp.take<double>(Na);
    l.fsscales = bp.take<double>(Nfs); l.fslengths = bp.take<double>(Nfs); l.fsc = bp.take<double>(Nfs);
    l.sa = bp.take<double>(Nfs); l.sg = bp.take<double>(Nfs); l.fsb = bp.take<double>(Nfs); l.fsd = bp.take<double>(Nfs);
    l.dta = bp.take<double>(Na); l.dtg = bp.take<double>(Na);
    l.evals = bp.take<unsigned long long>(1);
    l.status = bp.take<int>(Nfs);
    l.bytes = bp.off + 64;
    return l;
}

namespace ksn {
void k2_prefetch_shutdown()
{
    K2Prefetch &q = g_pre;
    if (q.d_buf) cudaFree(q.d_buf);
    if (q.h_sf) cudaFreeHost(q.h_sf);
    if (q.done) cudaEventDestroy(q.done);
    if (q.main_idle) cudaEventDestroy(q.main_idle);
    if (q.stream) cudaStreamDestroy(q.stream);
    free(q.sf);
    q = K2Prefetch();
}
}  // namespace ksn

// The request is only RECORDED here; the kernels are launched by k2_prefetch_launch_pending() -- which K1's launcher calls
// right after the sweep is in flight (so the host time of these launches does not delay K1), and ksn_delta_nu_integrate
// itself if nobody has by then.
extern "C" int ksn_delta_nu_prefetch(double a, double TimeTransfer, double light, const double *scalefact, int Na, int namax)
{
    int rc = ensure_init();
    if (rc) return rc;
    Ctx &c = ctx();
    K2Prefetch &q = g_pre;
    q.valid = false;
    q.pending = false;
    if (!scalefact || Na < 1 || namax < Na || !(a > 0) || !(TimeTransfer > 0)) return set_error(KSN_EINVAL, "ksn_delta_nu_prefetch: bad arguments");
    if (!c.d_bg) return set_error(KSN_EINVAL, "ksn_delta_nu_prefetch: call ksn_set_background first");
    const double loga0 = log(TimeTransfer), loga = log(a);
    if (loga0 < c.bg_lo + 2 * c.bg_h || loga > c.bg_hi - 2 * c.bg_h) return set_error(KSN_EINVAL, "ksn_delta_nu_prefetch: outside the background table");
    if (q.sf_cap < namax) {
        free(q.sf);
        q.sf = (double *) malloc(sizeof(double) * namax);
        q.sf_cap = q.sf ? namax : 0;
        if (!q.sf) return set_error(KSN_ENOMEM, "ksn_delta_nu_prefetch: out of host memory");
    }
    memcpy(q.sf, scalefact, sizeof(double) * Na);
    q.a = a; q.a0 = TimeTransfer; q.light = light; q.Na = Na; q.namax = namax;
    q.pending = true;
    return KSN_OK;
}

namespace ksn {
int k2_prefetch_launch_pending()
{
    K2Prefetch &q = g_pre;
    if (!q.pending) return KSN_OK;
    q.pending = false;
    Ctx &c = ctx();
    if (!c.d_bg) return KSN_OK;
    const int Na = q.Na, namax = q.namax;
    const double loga0 = log(q.a0), loga = log(q.a);
    if (!q.stream) {
        KSN_CUDA(cudaStreamCreateWithFlags(&q.stream, cudaStreamNonBlocking));
        KSN_CUDA(cudaEventCreateWithFlags(&q.done, cudaEventDisableTiming));
        KSN_CUDA(cudaEventCreateWithFlags(&q.main_idle, cudaEventDisableTiming));
    }
    // sized for a full history the first time (no reallocation during a run: see ksn_delta_nu_integrate)
    const K2PreLayout full = k2_pre_layout(nullptr, namax);
    if (full.bytes > q.cap) {
        KSN_CUDA(cudaStreamSynchronize(q.stream));
        if (q.d_buf) cudaFree(q.d_buf);
        q.d_buf = nullptr; q.cap = 0;
        KSN_CUDA(cudaMalloc((void **) &q.d_buf, full.bytes));
        q.cap = full.bytes;
    }
    if ((size_t) namax * sizeof(double) > q.h_cap) {
        KSN_CUDA(cudaStreamSynchronize(q.stream));
        if (q.h_sf) cudaFreeHost(q.h_sf);
        q.h_sf = nullptr; q.h_cap = 0;
        KSN_CUDA(cudaHostAlloc((void **) &q.h_sf, (size_t) namax * sizeof(double), cudaHostAllocDefault));
        q.h_cap = (size_t) namax * sizeof(double);
    }
    // The previous tables may still be read by a K2 launch on the main stream: they were, at the latest, when that call
    // returned (it synchronises), so only the prefetch stream's own previous work is waited for (it finished long ago).
    KSN_CUDA(cudaStreamSynchronize(q.stream));
    memcpy(q.h_sf, q.sf, sizeof(double) * Na);
    const K2PreLayout l = k2_pre_layout(q.d_buf, Na);
    const int Nfs = 16 * Na;
    KSN_CUDA(cudaMemcpyAsync(l.knots, q.h_sf, sizeof(double) * Na, cudaMemcpyHostToDevice, q.stream));
    KSN_CUDA(cudaMemsetAsync(l.evals, 0, sizeof(unsigned long long) + 8 + (size_t) Nfs * sizeof(int), q.stream));   // evals | status (adjacent)
    const BgTable bg = bg_table();
    fs_knots_kernel<<<(Nfs + 127) / 128, 128, 0, q.stream>>>(loga0, loga, Nfs, l.fsscales);
    fslength_kernel<<<Nfs, K2_THREADS, 0, q.stream>>>(bg, l.fsscales, Nfs, loga, q.light, l.fslengths, l.status, l.evals);
    k2_prep_splines_kernel<<<2, 256, 0, q.stream>>>(l.fsscales, l.fslengths, Nfs, l.fsc, l.fsb, l.fsd, l.sa, l.sg, l.knots, Na, l.dta, l.dtg);
    c.launches += 3;
    KSN_CUDA(cudaGetLastError());
    KSN_CUDA(cudaEventRecord(q.done, q.stream));
    q.bg = c.d_bg; q.bg_n = c.bg_n; q.bg_lo = c.bg_lo; q.bg_h = c.bg_h; q.bg_npatch = g_bg_npatch;
    q.valid = true;
    return KSN_OK;
}
}  // namespace ksn

extern "C" int ksn_delta_nu_integrate(const ksn_delta_nu_args *A, double *out, unsigned long long *n_evals)
{
    int rc = ensure_init();
    if (rc) return rc;
    Ctx &c = ctx();
    if (!A || !out || A->nk < 1 || A->Na < 1 || A->namax < A->Na || A->nspecies < 1 || A->nspecies > 3 ||
        !A->scalefact || !A->delta_tot || !A->wavenum || !A->delta_nu_init || !(A->a > 0) || !(A->TimeTransfer > 0))
        return set_error(KSN_EINVAL, "ksn_delta_nu_integrate: bad arguments");
    if (!c.d_bg) return set_error(KSN_EINVAL, "ksn_delta_nu_integrate: call ksn_set_background first");
    const int nk = A->nk, Na = A->Na, ns = A->nspecies, Nfs = 16 * Na, namax = A->namax;
    const double loga0 = log(A->TimeTransfer), loga = log(A->a);
    if (loga0 < c.bg_lo + 2 * c.bg_h || loga > c.bg_hi - 2 * c.bg_h)
        return set_error(KSN_EINVAL, "ksn_delta_nu_integrate: [%g,%g] outside the background table [%g,%g]", loga0, loga, c.bg_lo, c.bg_hi);
    bool any_integral = false;
    for (int s = 0; s < ns; s++) any_integral |= A->integrate[s] != 0;

    // Were the a-only tables prefetched for exactly these inputs (ksn_delta_nu_prefetch)?
    K2Prefetch &q = g_pre;
    if (q.pending) { rc = k2_prefetch_launch_pending(); if (rc) return rc; }      // (nobody launched the request: do it now)
    const bool hit = q.valid && q.a == A->a && q.a0 == A->TimeTransfer && q.light == A->light && q.Na == Na &&
                     q.bg == c.d_bg && q.bg_n == c.bg_n && q.bg_lo == c.bg_lo && q.bg_h == c.bg_h && q.bg_npatch == g_bg_npatch &&
                     memcmp(q.sf, A->scalefact, sizeof(double) * Na) == 0;
    q.valid = false;                     // single use either way
    g_last_prefetch_hit = hit ? 1 : 0;

    // Device block: [scalefact: namax slots | delta_tot: nk rows of namax | wavenum nk | delta_nu_init nk], the reference's own
    // layout of the first two (delta_tot_table.c:40-46), so that a table whose block is page-locked (the host layer
    // registers it, src/deltatot.c) goes up in ONE DMA without a staging copy.  Sized for a full history (three species)
    // the first time, so that no step of a run ever reallocates: cudaFree / cudaMalloc / cudaHostAlloc cost ~0.1 s each
    // once peer access between the GPUs of the box is on (measured: one 104 ms step when the 100th row arrived).
    const size_t n_blk = (size_t) namax + (size_t) nk * namax, n_in = n_blk + 2 * (size_t) nk;
    const size_t n_res = (size_t) 3 * nk + 8 + ((size_t) 3 * nk + 16 * (size_t) namax + 8) / 2;     // out | evals | status, in doubles
    {
        const size_t cNfs = 16 * (size_t) namax;
        const size_t dev_doubles = n_in + 3 * cNfs + 4 * cNfs + 2 * (size_t) namax + n_res + 32;
        rc = ensure_device_buffer((void **) &c.d_k2, &c.k2_cap, dev_doubles * sizeof(double) + 256);
        if (rc) return rc;
        rc = ensure_pinned_buffer((void **) &c.h_k2, &c.h_k2_cap, (n_in + n_res + 16) * sizeof(double));
        if (rc) return rc;
    }
    KSN_CUDA(cudaStreamSynchronize(c.stream));
    const bool contiguous = A->delta_tot == A->scalefact + namax;
    const bool direct = contiguous && host_range_registered(A->scalefact, n_blk * sizeof(double));
    double *h = c.h_k2;
    if (!direct) {
        memcpy(h, A->scalefact, sizeof(double) * Na);
        memcpy(h + namax, A->delta_tot, sizeof(double) * (size_t) nk * namax);
    }
    memcpy(h + n_blk, A->wavenum, sizeof(double) * nk);
    memcpy(h + n_blk + nk, A->delta_nu_init, sizeof(double) * nk);

    Bump bp{ (char *) c.d_k2 };
    double *d_in = bp.take<double>(n_in);
    double *d_fsscales = bp.take<double>(Nfs), *d_fslengths = bp.take<double>(Nfs), *d_fsc = bp.take<double>(Nfs);
    double *d_sa = bp.take<double>(Nfs), *d_sg = bp.take<double>(Nfs);
    double *d_fsb = bp.take<double>(Nfs), *d_fsd = bp.take<double>(Nfs);
    double *d_dta = bp.take<double>(Na), *d_dtg = bp.take<double>(Na);
    // results, one contiguous span (one memset, one copy back): delta_nu | evaluations | per-bin status | fslength status
    double *d_out = bp.take<double>((size_t) ns * nk);
    unsigned long long *d_evals = bp.take<unsigned long long>(1);
    int *d_status = bp.take<int>((size_t) ns * nk + Nfs);
    const size_t res_bytes = (size_t) ((char *) (d_status + (size_t) ns * nk + Nfs) - (char *) d_out);
    if (bp.off > c.k2_cap || res_bytes > n_res * sizeof(double)) return set_error(KSN_ENOMEM, "K2 workspace accounting error");
    int *d_fs_status = d_status + (size_t) ns * nk;
    if (hit) {
        const K2PreLayout l = k2_pre_layout(q.d_buf, Na);
        d_fsscales = l.fsscales; d_fslengths = l.fslengths; d_fsc = l.fsc; d_fsb = l.fsb; d_fsd = l.fsd; d_dta = l.dta; d_dtg = l.dtg;
    }

    phase_begin(PH_K2);
    if (direct) KSN_CUDA(cudaMemcpyAsync(d_in, A->scalefact, n_blk * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    else KSN_CUDA(cudaMemcpyAsync(d_in, h, n_blk * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    KSN_CUDA(cudaMemcpyAsync(d_in + n_blk, h + n_blk, 2 * (size_t) nk * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    KSN_CUDA(cudaMemsetAsync(d_evals, 0, (size_t) ((char *) (d_status + (size_t) ns * nk + Nfs) - (char *) d_evals), c.stream));
    const BgTable bg = bg_table();
    if (hit) {
        // tables, their quadrature status and evaluation count come from the prefetch (side stream, beside K1)
        KSN_CUDA(cudaStreamWaitEvent(c.stream, q.done, 0));
        const K2PreLayout l = k2_pre_layout(q.d_buf, Na);
        KSN_CUDA(cudaMemcpyAsync(d_fs_status, l.status, (size_t) Nfs * sizeof(int), cudaMemcpyDeviceToDevice, c.stream));
        KSN_CUDA(cudaMemcpyAsync(d_evals, l.evals, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, c.stream));
    } else {
        // The fs table is needed for the initial-condition term too (its knot 0 is fslength(log a0, log a)).
        fs_knots_kernel<<<(Nfs + 127) / 128, 128, 0, c.stream>>>(loga0, loga, Nfs, d_fsscales);
        fslength_kernel<<<any_integral ? Nfs : 1, K2_THREADS, 0, c.stream>>>(bg, d_fsscales, any_integral ? Nfs : 1, loga, A->light,
                                                                            d_fslengths, d_fs_status, d_evals);
        c.launches += 2;
        if (any_integral) {
            k2_prep_splines_kernel<<<2, 256, 0, c.stream>>>(d_fsscales, d_fslengths, Nfs, d_fsc, d_fsb, d_fsd, d_sa, d_sg, d_in, Na, d_dta, d_dtg);
            c.launches++;
        }
    }
    const int k_first = 0, k_count = nk;
    K2Dev p;
    p.nk = nk; p.Na = Na; p.namax = A->namax; p.Nfs = Nfs; p.k_first = k_first;
    p.loga0 = loga0; p.loga = loga; p.light = A->light; p.delta_nu_prefac = A->delta_nu_prefac;
    p.deriv_prefac = A->deriv_prefac; p.nufrac_low0 = A->nufrac_low0;
    for (int s = 0; s < 3; s++) {
        p.mnubykT[s] = s < ns ? A->mnubykT[s] : 0; p.qc[s] = s < ns ? A->qc[s] : 0;
        p.relerr[s] = s < ns ? A->relerr[s] : 1e-6; p.integrate[s] = s < ns ? A->integrate[s] : 0;
    }
    p.scalefact = d_in; p.delta_tot = d_in + namax; p.wavenum = d_in + n_blk;
    p.delta_nu_init = p.wavenum + nk;
    p.fsscales = d_fsscales; p.fslengths = d_fslengths; p.fs_c = d_fsc; p.fs_b = d_fsb; p.fs_d = d_fsd; p.dt_alpha = d_dta; p.dt_gamma = d_dtg;
    p.out = d_out; p.status = d_status; p.evals = d_evals; p.bg = bg;
    if (k_count > 0) {
        const dim3 grid(k_count, ns);
        const size_t smem = 5 * (size_t) Na * sizeof(double);
        int width = k2_spec_width();
        if (width == 0) width = k2_spec_auto(ns, p.qc);
        switch (width) {
        case 2: k2_delta_nu_spec_kernel<2, 3><<<grid, 2 * K2_THREADS, smem, c.stream>>>(p); break;
        case 3: k2_delta_nu_spec_kernel<3, 2><<<grid, 3 * K2_THREADS, smem, c.stream>>>(p); break;
        case 4: k2_delta_nu_spec_kernel<4, 2><<<grid, 4 * K2_THREADS, smem, c.stream>>>(p); break;
        default: k2_delta_nu_kernel<<<grid, K2_THREADS, smem, c.stream>>>(p); break;
        }
    }
    c.launches++;
    KSN_CUDA(cudaGetLastError());
    double *h_out = h + n_in + (16 - (n_in & 15)) % 16;                 // the result span, laid out as on the device
    int *h_status = (int *) ((char *) h_out + ((char *) d_status - (char *) d_out));
    unsigned long long *h_evals = (unsigned long long *) ((char *) h_out + ((char *) d_evals - (char *) d_out));
    KSN_CUDA(cudaMemcpyAsync(h_out, d_out, res_bytes, cudaMemcpyDeviceToHost, c.stream));
    phase_end(PH_K2);
    KSN_CUDA(cudaStreamSynchronize(c.stream));
    phase_collect();
    memcpy(out, h_out, sizeof(double) * ns * nk);
    g_last_evals = *h_evals;
    if (n_evals) *n_evals = *h_evals;
    g_max_passes = 0;
    g_max_trips = 0;
    for (size_t i = 0; i < (size_t) ns * nk + Nfs; i++) {
        const int st = i < (size_t) ns * nk ? (h_status[i] & 0xff) : h_status[i];
        if (i < (size_t) ns * nk) {
            g_max_passes = std::max(g_max_passes, ((unsigned) h_status[i] >> 8) & 0xfffu);
            g_max_trips = std::max(g_max_trips, (unsigned) h_status[i] >> 20);
        }
        if (st)
            return set_error(KSN_EQUAD, "quadrature %zu (%s) failed with GSL-style code %d at a=%g",
                             i, i < (size_t) ns * nk ? "delta_nu" : "fslength", st, A->a);
    }
    return KSN_OK;
}
