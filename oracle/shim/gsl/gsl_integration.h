/* TEST INFRASTRUCTURE (oracle): mini-GSL adaptive Gauss-Kronrod quadrature (QAG).
 * Restates the published GSL 2.x / QUADPACK algorithm (integration/qag.c, qk.c, qk61.c,
 * qpsrt.c, util.c); GSL itself is neither in this image nor vendored by the reference
 * (README.txt:44-48 names "GSL" without a version). */
#ifndef KSN_MINIGSL_INTEGRATION_H
#define KSN_MINIGSL_INTEGRATION_H
#include <stddef.h>
typedef struct { double (*function)(double x, void *params); void *params; } gsl_function;
#define GSL_FN_EVAL(F, x) (*((F)->function))(x, (F)->params)
typedef struct {
    size_t limit, size, nrmax, i, maximum_level;
    double *alist, *blist, *rlist, *elist;
    size_t *order, *level;
} gsl_integration_workspace;
enum { GSL_INTEG_GAUSS15 = 1, GSL_INTEG_GAUSS21 = 2, GSL_INTEG_GAUSS31 = 3,
       GSL_INTEG_GAUSS41 = 4, GSL_INTEG_GAUSS51 = 5, GSL_INTEG_GAUSS61 = 6 };
gsl_integration_workspace *gsl_integration_workspace_alloc(const size_t n);
void gsl_integration_workspace_free(gsl_integration_workspace *w);
int gsl_integration_qag(const gsl_function *f, double a, double b, double epsabs, double epsrel,
                        size_t limit, int key, gsl_integration_workspace *workspace,
                        double *result, double *abserr);
void gsl_integration_qk61(const gsl_function *f, double a, double b,
                          double *result, double *abserr, double *resabs, double *resasc);
/* oracle instrumentation: number of integrand evaluations since process start */
extern unsigned long long ksn_minigsl_nevals;
#endif
