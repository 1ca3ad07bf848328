// Bookkeeping of QUADPACK's adaptive bisection (gsl_integration_qag, GSL integration/qag.c; the reference calls it at
// delta_tot_table.c:404,597) written so that the expensive part -- the two 61-point rules on the halves of an
// interval -- may be computed AHEAD of the moment the sequential algorithm asks for it.
//
// QAG is a strictly sequential loop: pick the interval with the largest error estimate, bisect it, update the sums,
// test for termination.  Its critical path on the GPU is (number of bisections) x (latency of one rule application).
// The values a bisection produces depend only on the interval's end points, never on when it is bisected, so the
// kernel evaluates the halves of the M worst intervals at once ("speculation"), keeps the results with the interval
// ("cached children"), and then REPLAYS the sequential loop step by step with qags_apply(), stopping at the first
// interval whose children are not known yet.  Decisions, sums and their order are exactly the sequential ones, hence
// so is the result -- bit for bit; only the number of trips through the integrand shrinks.
//
// Plain C++ on purpose (no CUDA types): the same code is compiled by nvcc into the kernel and by g++ into the CPU test
// that replays it against the sequential loop (tests/test_qag_spec_cpu.py).
#ifndef KSN_QAG_SPEC_H
#define KSN_QAG_SPEC_H

#include <float.h>
#include <math.h>

#if defined(__CUDACC__)
#define KSN_HD __host__ __device__ __forceinline__
#else
#define KSN_HD inline
#endif

namespace ksn {

constexpr int QAGS_LIMIT = 200;          // GSL_VAL, kspace_neutrino_const.h:19

enum { QAGS_OK = 0, QAGS_EFAILED = 5, QAGS_EMAXITER = 11, QAGS_EROUND = 18, QAGS_ESING = 21 };

struct QagSpecList {                     // shared memory on the device
    double a[QAGS_LIMIT], b[QAGS_LIMIT], r[QAGS_LIMIT], e[QAGS_LIMIT];   // GSL workspace alist/blist/rlist/elist
    double cr[2 * QAGS_LIMIT], ce[2 * QAGS_LIMIT];                        // children of slot i: result/abserr of [a,mid] at 2i, of [mid,b] at 2i+1
    unsigned char cf[2 * QAGS_LIMIT];                                     // child's resasc != abserr (qag.c's round-off test needs it)
    unsigned char cached[QAGS_LIMIT];                                     // children of slot i are known (or being computed)
};

struct QagSpecState {
    double area, errsum, tolerance, epsabs, epsrel;
    int size, iteration, limit, roundoff_type1, roundoff_type2, error_type;
    unsigned passes;                     // 61-point rule applications the SEQUENTIAL algorithm has made so far
};

KSN_HD bool qags_too_small(double a1, double a2, double b2)
{
    const double tmp = (1 + 100 * DBL_EPSILON) * (fabs(a2) + 1000 * DBL_MIN);
    return fabs(a1) <= tmp && fabs(b2) <= tmp;
}

// After the rule on the whole interval [a,b]: either the integral is finished (returns true, *status set) or the list
// starts with that one interval.
KSN_HD bool qags_begin(QagSpecState &s, QagSpecList &L, double a, double b, double epsabs, double epsrel, int limit,
                       double result, double abserr, double resabs, double resasc, int *status)
{
    s.epsabs = epsabs; s.epsrel = epsrel; s.limit = limit;
    s.area = result; s.errsum = abserr;
    s.tolerance = fmax(epsabs, epsrel * fabs(result));
    s.size = 1; s.iteration = 1; s.roundoff_type1 = 0; s.roundoff_type2 = 0; s.error_type = 0;
    s.passes = 1;
    const double round_off = 50 * DBL_EPSILON * resabs;
    if (abserr <= round_off && abserr > s.tolerance) { *status = QAGS_EROUND; return true; }
    if ((abserr <= s.tolerance && abserr != resasc) || abserr == 0.0) { *status = QAGS_OK; return true; }
    if (limit == 1) { *status = QAGS_EMAXITER; return true; }
    L.a[0] = a; L.b[0] = b; L.r[0] = result; L.e[0] = abserr; L.cached[0] = 0;
    return false;
}

// One trip through qag.c's loop for slot i (the interval with the largest error), whose halves are in the cache.
// Returns true while the loop goes on.
KSN_HD bool qags_apply(QagSpecState &s, QagSpecList &L, int i)
{
    const double a_i = L.a[i], b_i = L.b[i], r_i = L.r[i], e_i = L.e[i];
    const double a1 = a_i, b1 = 0.5 * (a_i + b_i), a2 = b1, b2 = b_i;
    const double r1 = L.cr[2 * i], e1 = L.ce[2 * i], r2 = L.cr[2 * i + 1], e2 = L.ce[2 * i + 1];
    s.passes += 2;
    const double area12 = r1 + r2, error12 = e1 + e2;
    s.errsum += (error12 - e_i);
    s.area += area12 - r_i;
    if (L.cf[2 * i] && L.cf[2 * i + 1]) {
        const double delta = r_i - area12;
        if (fabs(delta) <= 1.0e-5 * fabs(area12) && error12 >= 0.99 * e_i) s.roundoff_type1++;
        if (s.iteration >= 10 && error12 > e_i) s.roundoff_type2++;
    }
    s.tolerance = fmax(s.epsabs, s.epsrel * fabs(s.area));
    if (s.errsum > s.tolerance) {
        if (s.roundoff_type1 >= 6 || s.roundoff_type2 >= 20) s.error_type = 2;
        if (qags_too_small(a1, a2, b2)) s.error_type = 3;
    }
    // GSL workspace update(): the half with the larger error keeps slot i, the other one is appended
    const int n = s.size;
    if (e2 > e1) {
        L.a[i] = a2; L.r[i] = r2; L.e[i] = e2;
        L.a[n] = a1; L.b[n] = b1; L.r[n] = r1; L.e[n] = e1;
    } else {
        L.b[i] = b1; L.r[i] = r1; L.e[i] = e1;
        L.a[n] = a2; L.b[n] = b2; L.r[n] = r2; L.e[n] = e2;
    }
    L.cached[i] = 0;
    L.cached[n] = 0;
    s.size++;
    s.iteration++;
    return s.iteration < s.limit && !s.error_type && s.errsum > s.tolerance;
}

// The loop has ended: result summed in storage order (as qag.c does), GSL-style status.
KSN_HD int qags_finish(const QagSpecState &s, const QagSpecList &L, double *result, double *abserr)
{
    double sum = 0;
    for (int k = 0; k < s.size; k++) sum += L.r[k];
    *result = sum;
    *abserr = s.errsum;
    if (s.errsum <= s.tolerance) return QAGS_OK;
    if (s.error_type == 2) return QAGS_EROUND;
    if (s.error_type == 3) return QAGS_ESING;
    if (s.iteration == s.limit) return QAGS_EMAXITER;
    return QAGS_EFAILED;
}

// An interval is worth integrating ahead of time only if the loop is likely to come to it: were every error estimate
// below tolerance/size the loop would have stopped already.  (A heuristic: it changes how much work is wasted, never
// the result.)
KSN_HD double qags_spec_threshold(const QagSpecState &s) { return s.tolerance / (4.0 * s.size); }

}  // namespace ksn

#endif
