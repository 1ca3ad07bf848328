// TEST PROGRAM (CPU).  Replays the bookkeeping K2's speculative kernel uses (kspace_neutrinos_b200/csrc/ksn_qag_spec.h,
// the same header nvcc compiles into the kernel) against (1) a plain sequential restatement of QAG's loop and
// (2) the oracle's mini-GSL gsl_integration_qag, on a family of integrands: result, error estimate, status and the
// count of rule applications must be IDENTICAL (bitwise) whatever the speculation width M; and prints how many passes
// through the integrand (the critical path on the GPU) each width needs.
//
// Build: g++ -O2 -I kspace_neutrinos_b200/csrc -I oracle/shim tests/qag_spec_host.cpp oracle/mini_gsl.c -lm
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#include "ksn_qag_spec.h"

extern "C" {
#include <gsl/gsl_errno.h>
#include <gsl/gsl_integration.h>
}

static void quiet_handler(const char *, const char *, int, int) {}

using namespace ksn;

typedef double (*fn_t)(double, void *);

struct Case { const char *name; fn_t f; double w, a, b, epsrel; };

static double f_smooth(double x, void *) { return exp(x) * sin(3 * x); }
static double f_cos(double x, void *p) { return cos(*(double *) p * x); }
static double f_peak(double x, void *p) { const double w = *(double *) p; return 1.0 / ((x + 1.3) * (x + 1.3) + w); }
static double f_sqrt(double x, void *) { return sqrt(fabs(x + 0.7)); }
// like the hybrid-neutrino integrand: an envelope times an oscillation whose frequency grows towards early times
static double f_chirp(double x, void *p) { const double w = *(double *) p; return exp(1.5 * x) * cos(w * (exp(-0.5 * x) - 1.0)) / (1 + 0.1 * w * w * exp(-x) * 1e-4); }
static double f_log(double x, void *) { return log(fabs(x + 2.0) + 1e-300); }

static void rule(fn_t f, void *p, double a, double b, double *res, double *err, double *rabs, double *rasc)
{
    gsl_function F;
    F.function = f;
    F.params = p;
    gsl_integration_qk61(&F, a, b, res, err, rabs, rasc);
}

struct Out { double result, abserr; int status; unsigned passes, rules, trips; };

// the loop exactly as k2_delta_nu_kernel's qag61_block runs it (largest error, lowest index among equals)
static Out sequential(fn_t f, void *p, double a, double b, double epsabs, double epsrel, int limit)
{
    static QagSpecList L;
    QagSpecState s;
    Out o;
    double r, e, ra, rs;
    rule(f, p, a, b, &r, &e, &ra, &rs);
    o.rules = 1; o.trips = 1;
    int st;
    if (qags_begin(s, L, a, b, epsabs, epsrel, limit, r, e, ra, rs, &st)) { o.result = r; o.abserr = e; o.status = st; o.passes = 1; return o; }
    for (;;) {
        int i = 0;
        for (int k = 1; k < s.size; k++) if (L.e[k] > L.e[i]) i = k;
        const double mid = 0.5 * (L.a[i] + L.b[i]);
        for (int c = 0; c < 2; c++) {
            rule(f, p, c ? mid : L.a[i], c ? L.b[i] : mid, &r, &e, &ra, &rs);
            L.cr[2 * i + c] = r; L.ce[2 * i + c] = e; L.cf[2 * i + c] = rs != e;
        }
        o.rules += 2; o.trips++;
        if (!qags_apply(s, L, i)) break;
    }
    o.status = qags_finish(s, L, &o.result, &o.abserr);
    o.passes = s.passes;
    return o;
}

// the speculative schedule of qag61_spec<M>: M slots per trip through the integrand, then replay
static Out speculative(int M, fn_t f, void *p, double a, double b, double epsabs, double epsrel, int limit)
{
    static QagSpecList L;
    QagSpecState s;
    Out o;
    double r, e, ra, rs;
    rule(f, p, a, b, &r, &e, &ra, &rs);
    o.rules = 1; o.trips = 1;
    int st;
    if (qags_begin(s, L, a, b, epsabs, epsrel, limit, r, e, ra, rs, &st)) { o.result = r; o.abserr = e; o.status = st; o.passes = 1; return o; }
    int sel[8], nsel = 1;
    sel[0] = 0; L.cached[0] = 1;
    for (;;) {
        for (int m = 0; m < nsel; m++) {
            const int i = sel[m];
            const double mid = 0.5 * (L.a[i] + L.b[i]);
            for (int c = 0; c < 2; c++) {
                rule(f, p, c ? mid : L.a[i], c ? L.b[i] : mid, &r, &e, &ra, &rs);
                L.cr[2 * i + c] = r; L.ce[2 * i + c] = e; L.cf[2 * i + c] = rs != e;
            }
        }
        o.rules += 2 * nsel; o.trips++;
        bool done = false;
        int i;
        for (;;) {
            i = 0;
            for (int k = 1; k < s.size; k++) if (L.e[k] > L.e[i]) i = k;
            if (!L.cached[i]) break;
            if (!qags_apply(s, L, i)) { done = true; break; }
        }
        if (done) break;
        const double floor_ = qags_spec_threshold(s);
        L.cached[i] = 1; sel[0] = i; nsel = 1;
        for (; nsel < M; nsel++) {
            int j = -1;
            for (int k = 0; k < s.size; k++) {
                if (L.cached[k] || !(L.e[k] > floor_)) continue;
                if (j < 0 || L.e[k] > L.e[j]) j = k;
            }
            if (j < 0) break;
            L.cached[j] = 1; sel[nsel] = j;
        }
    }
    o.status = qags_finish(s, L, &o.result, &o.abserr);
    o.passes = s.passes;
    return o;
}

int main(void)
{
    gsl_set_error_handler(quiet_handler);
    static double w[] = { 0, 50, 300, 2000, 1e-4, 1e-8, 0, 30, 400, 3000, 0 };
    const Case cases[] = {
        { "smooth", f_smooth, 0, -4.6, 0.0, 1e-6 },       { "cos50", f_cos, 50, -4.6, 0.0, 1e-6 },
        { "cos300", f_cos, 300, -4.6, 0.0, 1e-6 },        { "cos2000", f_cos, 2000, -4.6, 0.0, 1e-6 },
        { "peak1e-4", f_peak, 1e-4, -4.6, 0.0, 1e-6 },    { "peak1e-8", f_peak, 1e-8, -4.6, 0.0, 1e-6 },
        { "sqrt", f_sqrt, 0, -4.6, 0.0, 1e-6 },           { "chirp30", f_chirp, 30, -4.6, 0.0, 1e-6 },
        { "chirp400", f_chirp, 400, -4.6, 0.0, 1e-6 },    { "chirp3000", f_chirp, 3000, -4.6, 0.0, 1e-4 },
        { "log-sing", f_log, 0, -4.6, 0.0, 1e-9 },        { "cos2000-tight", f_cos, 2000, -4.6, 0.0, 1e-12 },
        { "chirp400-loose", f_chirp, 400, -4.6, 0.0, 3e-5 },
    };
    (void) w;
    int bad = 0;
    gsl_integration_workspace *ws = gsl_integration_workspace_alloc(QAGS_LIMIT);
    for (size_t c = 0; c < sizeof cases / sizeof cases[0]; c++) {
        Case cs = cases[c];
        const Out q = sequential(cs.f, &cs.w, cs.a, cs.b, 0.0, cs.epsrel, QAGS_LIMIT);
        gsl_function F;
        F.function = cs.f;
        F.params = &cs.w;
        double gr = 0, ge = 0;
        const int gst = gsl_integration_qag(&F, cs.a, cs.b, 0.0, cs.epsrel, QAGS_LIMIT, GSL_INTEG_GAUSS61, ws, &gr, &ge);
        const bool gsl_same = memcmp(&gr, &q.result, 8) == 0 && memcmp(&ge, &q.abserr, 8) == 0 && gst == q.status;
        printf("%-16s seq: result %.17g err %.3g status %d passes %u trips %u  mini-gsl %s\n", cs.name, q.result, q.abserr, q.status,
               q.passes, q.trips, gsl_same ? "identical" : "DIFFERS");
        if (!gsl_same) { printf("   mini-gsl: %.17g %.3g status %d\n", gr, ge, gst); bad++; }
        for (int M = 1; M <= 4; M++) {
            const Out s = speculative(M, cs.f, &cs.w, cs.a, cs.b, 0.0, cs.epsrel, QAGS_LIMIT);
            const bool same = memcmp(&s.result, &q.result, 8) == 0 && memcmp(&s.abserr, &q.abserr, 8) == 0 && s.status == q.status && s.passes == q.passes;
            printf("   M=%d: trips %3u  rules %3u (wasted %u)  %s\n", M, s.trips, s.rules, s.rules - q.passes, same ? "identical" : "DIFFERS");
            if (!same) bad++;
            if (M == 1 && (s.trips != q.trips || s.rules != q.rules)) { printf("   M=1 must be the sequential schedule\n"); bad++; }
        }
    }
    gsl_integration_workspace_free(ws);
    printf(bad ? "FAILED (%d)\n" : "ALL IDENTICAL\n", bad);
    return bad != 0;
}
