/* Host-side numerics of the B200 build: adaptive 61-point Gauss-Kronrod quadrature and 1-D
 * interpolation.  These replace the reference's GSL calls that stay on the host (init-time
 * Omega_nu tables, nufrac_low, the public fslength(), the transfer-function spline, the
 * delta_pow lookup).  The per-step integrals run on the GPU (csrc/k2_delta_nu.cu). */
#ifndef KSN_NUMERIC_H
#define KSN_NUMERIC_H
#include <stddef.h>

typedef double (*ksn_integrand)(double x, void *ctx);

/* status codes mirror the GSL errno values the reference's handler would report */
enum { KSN_Q_OK = 0, KSN_Q_EINVAL = 4, KSN_Q_EFAILED = 5, KSN_Q_EMAXITER = 11, KSN_Q_EBADTOL = 13, KSN_Q_EROUND = 18, KSN_Q_ESING = 21 };

/* QUADPACK QAG with the 61-point rule, same acceptance logic as gsl_integration_qag(key=6).
 * limit <= KSN_QAG_MAX_INTERVALS.  Returns a KSN_Q_* status; *result is set either way. */
#define KSN_QAG_MAX_INTERVALS 200
int ksn_qag61(ksn_integrand f, void *ctx, double a, double b, double epsabs, double epsrel, int limit,
              double *result, double *abserr);

/* Natural cubic spline: second-derivative coefficients c[0..n) for knots (x,y). n >= 3. */
void ksn_cspline_natural(const double *x, const double *y, int n, double *c);
/* Evaluate on the interval containing xq; *hint (may be NULL) caches the last interval. */
double ksn_cspline_eval(const double *x, const double *y, const double *c, int n, double xq, int *hint);
double ksn_linear_eval(const double *x, const double *y, int n, double xq, int *hint);
/* index i with x[i] <= xq < x[i+1]; the last interval for xq == x[n-1] */
int ksn_locate(const double *x, int n, double xq, int *hint);

/* private interpolator objects behind the gsl_interp* / gsl_interp_accel* struct members */
struct ksn_interp_s { int n; int cubic; double *c; };
struct ksn_accel_s { int hint; };
#endif
