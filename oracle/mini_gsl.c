/* TEST INFRASTRUCTURE (oracle) -- never linked into the product.
 *
 * mini-GSL: the subset of the GNU Scientific Library the reference calls, restated from
 * the published GSL 2.x algorithms so that the reference's own .c files compile
 * unmodified (oracle/_ref) and so that oracle/ksn_oracle.c can restate the path.
 * Third-party module restated: GSL (version unpinned by the reference, README.txt:44-48;
 * source absent from /root/reference).  Call sites this serves:
 *   gsl_integration_qag(key 6)   delta_tot_table.c:404,597  omega_nu_single.c:148,224
 *   gsl_interp_{cspline,linear}  delta_tot_table.c:110-116,495-496,559-596  delta_pow.c:14-16,35
 *                                omega_nu_single.c:136-151,199
 *   gsl_sf_bessel_j0             delta_tot_table.c:433
 *   gsl_set_error_handler        delta_tot_table.c:101
 * Pinned by the reference's known answers through oracle/run_ref_tests.sh (test_fslength
 * 1e-5, test_nufrac_low 1e-5, test_omega_nu_single_exact 1e-6, ...).  Agreement with a
 * real GSL build below ~1e-6 is UNPINNED: nothing in the reference's tests constrains it.
 */
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "gk61_tables.h"
#include "shim/gsl/gsl_errno.h"
#include "shim/gsl/gsl_integration.h"
#include "shim/gsl/gsl_interp.h"
#include "shim/gsl/gsl_sf_bessel.h"

/* ---------------------------------------------------------------- errors */
static gsl_error_handler_t *the_handler = NULL;

gsl_error_handler_t *gsl_set_error_handler(gsl_error_handler_t *h)
{
    gsl_error_handler_t *old = the_handler;
    the_handler = h;
    return old;
}

void gsl_error(const char *reason, const char *file, int line, int gsl_errno)
{
    if (the_handler) { the_handler(reason, file, line, gsl_errno); return; }
    fprintf(stderr, "mini-gsl: %s:%d: ERROR: %s (errno %d)\nDefault GSL error handler invoked.\n", file, line, reason, gsl_errno);
    fflush(NULL);
    abort();
}

/* ---------------------------------------------------------------- QK61 */
static const double xgk[31] = KSN_XGK61_INIT;
static const double wgk[31] = KSN_WGK61_INIT;
static const double wg[15] = KSN_WG30_INIT;
unsigned long long ksn_minigsl_nevals = 0;

static double rescale_error(double err, const double result_abs, const double result_asc)
{
    err = fabs(err);
    if (result_asc != 0 && err != 0) {
        const double scale = pow((200 * err / result_asc), 1.5);
        err = scale < 1 ? result_asc * scale : result_asc;
    }
    if (result_abs > DBL_MIN / (50 * DBL_EPSILON)) {
        const double min_err = 50 * DBL_EPSILON * result_abs;
        if (min_err > err) err = min_err;
    }
    return err;
}

/* gsl_integration_qk with n = 31: same evaluation and accumulation order as GSL's qk.c
 * (centre; Gauss nodes = odd indices; then the Kronrod-only even indices). */
void gsl_integration_qk61(const gsl_function *f, double a, double b,
                          double *result, double *abserr, double *resabs, double *resasc)
{
    enum { n = 31 };
    double fv1[n], fv2[n];
    const double center = 0.5 * (a + b);
    const double half_length = 0.5 * (b - a);
    const double abs_half_length = fabs(half_length);
    const double f_center = GSL_FN_EVAL(f, center);
    double result_gauss = 0;
    double result_kronrod = f_center * wgk[n - 1];
    double result_abs = fabs(result_kronrod);
    double result_asc, mean, err;
    int j;
    ksn_minigsl_nevals += 61;
    for (j = 0; j < (n - 1) / 2; j++) {
        const int jtw = j * 2 + 1;
        const double abscissa = half_length * xgk[jtw];
        const double fval1 = GSL_FN_EVAL(f, center - abscissa);
        const double fval2 = GSL_FN_EVAL(f, center + abscissa);
        const double fsum = fval1 + fval2;
        fv1[jtw] = fval1;
        fv2[jtw] = fval2;
        result_gauss += wg[j] * fsum;
        result_kronrod += wgk[jtw] * fsum;
        result_abs += wgk[jtw] * (fabs(fval1) + fabs(fval2));
    }
    for (j = 0; j < n / 2; j++) {
        const int jtwm1 = j * 2;
        const double abscissa = half_length * xgk[jtwm1];
        const double fval1 = GSL_FN_EVAL(f, center - abscissa);
        const double fval2 = GSL_FN_EVAL(f, center + abscissa);
        fv1[jtwm1] = fval1;
        fv2[jtwm1] = fval2;
        result_kronrod += wgk[jtwm1] * (fval1 + fval2);
        result_abs += wgk[jtwm1] * (fabs(fval1) + fabs(fval2));
    }
    mean = result_kronrod * 0.5;
    result_asc = wgk[n - 1] * fabs(f_center - mean);
    for (j = 0; j < n - 1; j++)
        result_asc += wgk[j] * (fabs(fv1[j] - mean) + fabs(fv2[j] - mean));
    err = (result_kronrod - result_gauss) * half_length;
    result_kronrod *= half_length;
    result_abs *= abs_half_length;
    result_asc *= abs_half_length;
    *result = result_kronrod;
    *resabs = result_abs;
    *resasc = result_asc;
    *abserr = rescale_error(err, result_abs, result_asc);
}

/* ---------------------------------------------------------------- QAG workspace */
gsl_integration_workspace *gsl_integration_workspace_alloc(const size_t n)
{
    gsl_integration_workspace *w;
    if (n == 0) GSL_ERROR_VAL("workspace length n must be positive integer", GSL_EDOM, 0);
    w = malloc(sizeof(*w));
    if (!w) return NULL;
    w->alist = malloc(n * sizeof(double));
    w->blist = malloc(n * sizeof(double));
    w->rlist = malloc(n * sizeof(double));
    w->elist = malloc(n * sizeof(double));
    w->order = malloc(n * sizeof(size_t));
    w->level = malloc(n * sizeof(size_t));
    w->size = 0; w->limit = n; w->maximum_level = 0; w->nrmax = 0; w->i = 0;
    return w;
}

void gsl_integration_workspace_free(gsl_integration_workspace *w)
{
    if (!w) return;
    free(w->level); free(w->order); free(w->elist); free(w->rlist); free(w->blist); free(w->alist);
    free(w);
}

/* keep `order` sorted by decreasing error estimate (QUADPACK dqpsrt as in GSL qpsrt.c) */
static void qpsrt(gsl_integration_workspace *w)
{
    const size_t last = w->size - 1;
    const size_t limit = w->limit;
    double *elist = w->elist;
    size_t *order = w->order;
    double errmax, errmin;
    int i, k, top;
    size_t i_nrmax = w->nrmax;
    size_t i_maxerr = order[i_nrmax];
    if (last < 2) {
        order[0] = 0;
        order[1] = 1;
        w->i = i_maxerr;
        return;
    }
    errmax = elist[i_maxerr];
    while (i_nrmax > 0 && errmax > elist[order[i_nrmax - 1]]) {
        order[i_nrmax] = order[i_nrmax - 1];
        i_nrmax--;
    }
    if (last < (limit / 2 + 2)) top = (int) last;
    else top = (int) (limit - last + 1);
    i = (int) i_nrmax + 1;
    while (i < top && errmax < elist[order[i]]) {
        order[i - 1] = order[i];
        i++;
    }
    order[i - 1] = i_maxerr;
    errmin = elist[last];
    k = top - 1;
    while (k > i - 2 && errmin >= elist[order[k]]) {
        order[k + 1] = order[k];
        k--;
    }
    order[k + 1] = last;
    i_maxerr = order[i_nrmax];
    w->i = i_maxerr;
    w->nrmax = i_nrmax;
}

static void ws_update(gsl_integration_workspace *w, double a1, double b1, double area1, double error1,
                      double a2, double b2, double area2, double error2)
{
    const size_t i_max = w->i;
    const size_t i_new = w->size;
    const size_t new_level = w->level[i_max] + 1;
    if (error2 > error1) {
        w->alist[i_max] = a2;      /* blist[i_max] is already b2 */
        w->rlist[i_max] = area2; w->elist[i_max] = error2; w->level[i_max] = new_level;
        w->alist[i_new] = a1; w->blist[i_new] = b1;
        w->rlist[i_new] = area1; w->elist[i_new] = error1; w->level[i_new] = new_level;
    } else {
        w->blist[i_max] = b1;      /* alist[i_max] is already a1 */
        w->rlist[i_max] = area1; w->elist[i_max] = error1; w->level[i_max] = new_level;
        w->alist[i_new] = a2; w->blist[i_new] = b2;
        w->rlist[i_new] = area2; w->elist[i_new] = error2; w->level[i_new] = new_level;
    }
    w->size++;
    if (new_level > w->maximum_level) w->maximum_level = new_level;
    qpsrt(w);
}

static int subinterval_too_small(double a1, double a2, double b2)
{
    const double e = DBL_EPSILON, u = DBL_MIN;
    const double tmp = (1 + 100 * e) * (fabs(a2) + 1000 * u);
    return fabs(a1) <= tmp && fabs(b2) <= tmp;
}

int gsl_integration_qag(const gsl_function *f, double a, double b, double epsabs, double epsrel,
                        size_t limit, int key, gsl_integration_workspace *w,
                        double *result, double *abserr)
{
    double area, errsum, result0, abserr0, resabs0, resasc0, tolerance, round_off;
    size_t iteration = 0, i;
    int roundoff_type1 = 0, roundoff_type2 = 0, error_type = 0;
    if (key != GSL_INTEG_GAUSS61) {
        /* the reference only ever passes 6 (GSL maps key<1 -> 15pt and key>6 -> 61pt) */
        if (key < 6) GSL_ERROR("mini-gsl implements only the 61 point rule", GSL_EINVAL);
    }
    w->size = 0; w->nrmax = 0; w->i = 0; w->maximum_level = 0;
    w->alist[0] = a; w->blist[0] = b; w->rlist[0] = 0; w->elist[0] = 0; w->order[0] = 0; w->level[0] = 0;
    *result = 0;
    *abserr = 0;
    if (limit > w->limit) GSL_ERROR("iteration limit exceeds available workspace", GSL_EINVAL);
    if (epsabs <= 0 && (epsrel < 50 * DBL_EPSILON || epsrel < 0.5e-28))
        GSL_ERROR("tolerance cannot be achieved with given epsabs and epsrel", GSL_EBADTOL);
    gsl_integration_qk61(f, a, b, &result0, &abserr0, &resabs0, &resasc0);
    w->size = 1; w->rlist[0] = result0; w->elist[0] = abserr0;
    tolerance = fmax(epsabs, epsrel * fabs(result0));
    round_off = 50 * DBL_EPSILON * resabs0;
    if (abserr0 <= round_off && abserr0 > tolerance) {
        *result = result0; *abserr = abserr0;
        GSL_ERROR("cannot reach tolerance because of roundoff error on first attempt", GSL_EROUND);
    } else if ((abserr0 <= tolerance && abserr0 != resasc0) || abserr0 == 0.0) {
        *result = result0; *abserr = abserr0;
        return GSL_SUCCESS;
    } else if (limit == 1) {
        *result = result0; *abserr = abserr0;
        GSL_ERROR("a maximum of one iteration was insufficient", GSL_EMAXITER);
    }
    area = result0;
    errsum = abserr0;
    iteration = 1;
    do {
        double a1, b1, a2, b2, a_i, b_i, r_i, e_i;
        double area1 = 0, area2 = 0, area12 = 0, error1 = 0, error2 = 0, error12 = 0;
        double resasc1, resasc2, resabs1, resabs2;
        a_i = w->alist[w->i]; b_i = w->blist[w->i]; r_i = w->rlist[w->i]; e_i = w->elist[w->i];
        a1 = a_i; b1 = 0.5 * (a_i + b_i); a2 = b1; b2 = b_i;
        gsl_integration_qk61(f, a1, b1, &area1, &error1, &resabs1, &resasc1);
        gsl_integration_qk61(f, a2, b2, &area2, &error2, &resabs2, &resasc2);
        area12 = area1 + area2;
        error12 = error1 + error2;
        errsum += (error12 - e_i);
        area += area12 - r_i;
        if (resasc1 != error1 && resasc2 != error2) {
            const double delta = r_i - area12;
            if (fabs(delta) <= 1.0e-5 * fabs(area12) && error12 >= 0.99 * e_i) roundoff_type1++;
            if (iteration >= 10 && error12 > e_i) roundoff_type2++;
        }
        tolerance = fmax(epsabs, epsrel * fabs(area));
        if (errsum > tolerance) {
            if (roundoff_type1 >= 6 || roundoff_type2 >= 20) error_type = 2;
            if (subinterval_too_small(a1, a2, b2)) error_type = 3;
        }
        ws_update(w, a1, b1, area1, error1, a2, b2, area2, error2);
        iteration++;
    } while (iteration < limit && !error_type && errsum > tolerance);
    {
        double s = 0;
        for (i = 0; i < w->size; i++) s += w->rlist[i];
        *result = s;
    }
    *abserr = errsum;
    if (errsum <= tolerance) return GSL_SUCCESS;
    else if (error_type == 2) GSL_ERROR("roundoff error prevents tolerance from being achieved", GSL_EROUND);
    else if (error_type == 3) GSL_ERROR("bad integrand behavior found in the integration interval", GSL_ESING);
    else if (iteration == limit) GSL_ERROR("maximum number of subdivisions reached", GSL_EMAXITER);
    else GSL_ERROR("could not integrate function", GSL_EFAILED);
}

/* ---------------------------------------------------------------- interpolation */
struct gsl_interp_type_s { const char *name; unsigned min_size; int cubic; };
static const struct gsl_interp_type_s linear_type = { "linear", 2, 0 };
static const struct gsl_interp_type_s cspline_type = { "cspline", 3, 1 };
const gsl_interp_type *gsl_interp_linear = &linear_type;
const gsl_interp_type *gsl_interp_cspline = &cspline_type;

typedef struct { double *c, *g, *diag, *offdiag; } cspline_state;

gsl_interp_accel *gsl_interp_accel_alloc(void)
{
    gsl_interp_accel *a = malloc(sizeof(*a));
    if (a) { a->cache = 0; a->hit_count = 0; a->miss_count = 0; }
    return a;
}

void gsl_interp_accel_free(gsl_interp_accel *a) { free(a); }

size_t gsl_interp_bsearch(const double x_array[], double x, size_t index_lo, size_t index_hi)
{
    size_t ilo = index_lo, ihi = index_hi;
    while (ihi > ilo + 1) {
        const size_t i = (ihi + ilo) / 2;
        if (x_array[i] > x) ihi = i; else ilo = i;
    }
    return ilo;
}

size_t gsl_interp_accel_find(gsl_interp_accel *a, const double xa[], size_t len, double x)
{
    const size_t x_index = a->cache;
    if (x < xa[x_index]) {
        a->miss_count++;
        a->cache = gsl_interp_bsearch(xa, x, 0, x_index);
    } else if (x >= xa[x_index + 1]) {
        a->miss_count++;
        a->cache = gsl_interp_bsearch(xa, x, x_index, len - 1);
    } else {
        a->hit_count++;
    }
    return a->cache;
}

gsl_interp *gsl_interp_alloc(const gsl_interp_type *T, size_t size)
{
    gsl_interp *interp;
    if (size < T->min_size) GSL_ERROR_VAL("insufficient number of points for interpolation type", GSL_EINVAL, NULL);
    interp = malloc(sizeof(*interp));
    if (!interp) return NULL;
    interp->type = T;
    interp->size = size;
    interp->state = NULL;
    if (T->cubic) {
        cspline_state *s = malloc(sizeof(*s));
        s->c = malloc(size * sizeof(double));
        s->g = malloc(size * sizeof(double));
        s->diag = malloc(size * sizeof(double));
        s->offdiag = malloc(size * sizeof(double));
        interp->state = s;
    }
    return interp;
}

void gsl_interp_free(gsl_interp *interp)
{
    if (!interp) return;
    if (interp->state) {
        cspline_state *s = interp->state;
        free(s->c); free(s->g); free(s->diag); free(s->offdiag); free(s);
    }
    free(interp);
}

/* symmetric positive-definite tridiagonal solve, A = L D L^T (GSL linalg/tridiag.c) */
static int solve_symm_tridiag(const double diag[], const double offdiag[], const double b[], double x[], size_t N)
{
    int status = GSL_SUCCESS;
    double *gamma = malloc(N * sizeof(double));
    double *alpha = malloc(N * sizeof(double));
    double *c = malloc(N * sizeof(double));
    double *z = malloc(N * sizeof(double));
    size_t i, j;
    alpha[0] = diag[0];
    gamma[0] = offdiag[0] / alpha[0];
    if (alpha[0] == 0) status = GSL_EZERODIV;
    for (i = 1; i < N - 1; i++) {
        alpha[i] = diag[i] - offdiag[i - 1] * gamma[i - 1];
        gamma[i] = offdiag[i] / alpha[i];
        if (alpha[i] == 0) status = GSL_EZERODIV;
    }
    if (N > 1) alpha[N - 1] = diag[N - 1] - offdiag[N - 2] * gamma[N - 2];
    z[0] = b[0];
    for (i = 1; i < N; i++) z[i] = b[i] - gamma[i - 1] * z[i - 1];
    for (i = 0; i < N; i++) c[i] = z[i] / alpha[i];
    x[N - 1] = c[N - 1];
    if (N >= 2) for (i = N - 2, j = 0; j <= N - 2; j++, i--) x[i] = c[i] - gamma[i] * x[i + 1];
    free(z); free(c); free(alpha); free(gamma);
    if (status == GSL_EZERODIV) GSL_ERROR("matrix must be positive definite", status);
    return status;
}

static int cspline_init(cspline_state *state, const double xa[], const double ya[], size_t size)
{
    size_t i;
    const size_t num_points = size;
    const size_t max_index = num_points - 1;
    const size_t sys_size = max_index - 1;
    state->c[0] = 0.0;
    state->c[max_index] = 0.0;
    for (i = 0; i < sys_size; i++) {
        const double h_i = xa[i + 1] - xa[i];
        const double h_ip1 = xa[i + 2] - xa[i + 1];
        const double ydiff_i = ya[i + 1] - ya[i];
        const double ydiff_ip1 = ya[i + 2] - ya[i + 1];
        const double g_i = (h_i != 0.0) ? 1.0 / h_i : 0.0;
        const double g_ip1 = (h_ip1 != 0.0) ? 1.0 / h_ip1 : 0.0;
        state->offdiag[i] = h_ip1;
        state->diag[i] = 2.0 * (h_ip1 + h_i);
        state->g[i] = 3.0 * (ydiff_ip1 * g_ip1 - ydiff_i * g_i);
    }
    if (sys_size == 1) {
        state->c[1] = state->g[0] / state->diag[0];
        return GSL_SUCCESS;
    }
    return solve_symm_tridiag(state->diag, state->offdiag, state->g, state->c + 1, sys_size);
}

int gsl_interp_init(gsl_interp *interp, const double xa[], const double ya[], size_t size)
{
    size_t i;
    if (size != interp->size) GSL_ERROR("data must match size of interpolation object", GSL_EINVAL);
    for (i = 1; i < size; i++)
        if (!(xa[i - 1] < xa[i])) GSL_ERROR("x values must be strictly increasing", GSL_EINVAL);
    interp->xmin = xa[0];
    interp->xmax = xa[size - 1];
    if (interp->type->cubic) return cspline_init(interp->state, xa, ya, size);
    return GSL_SUCCESS;
}

double gsl_interp_eval(const gsl_interp *interp, const double xa[], const double ya[], double x, gsl_interp_accel *a)
{
    size_t index;
    double x_lo, x_hi, y_lo, y_hi, dx, dy;
    if (x < interp->xmin || x > interp->xmax) GSL_ERROR_VAL("interpolation error", GSL_EDOM, NAN);
    if (a) index = gsl_interp_accel_find(a, xa, interp->size, x);
    else index = gsl_interp_bsearch(xa, x, 0, interp->size - 1);
    x_lo = xa[index]; x_hi = xa[index + 1];
    y_lo = ya[index]; y_hi = ya[index + 1];
    dx = x_hi - x_lo;
    dy = y_hi - y_lo;
    if (!(dx > 0.0)) GSL_ERROR_VAL("interpolation error", GSL_EINVAL, 0.0);
    if (interp->type->cubic) {
        const cspline_state *s = interp->state;
        const double delx = x - x_lo;
        const double c_i = s->c[index];
        const double c_ip1 = s->c[index + 1];
        const double b_i = (dy / dx) - dx * (c_ip1 + 2.0 * c_i) / 3.0;
        const double d_i = (c_ip1 - c_i) / (3.0 * dx);
        return y_lo + delx * (b_i + delx * (c_i + delx * d_i));
    }
    return y_lo + (x - x_lo) / dx * dy;
}

/* ---------------------------------------------------------------- special functions */
double gsl_sf_bessel_j0(const double x)
{
    const double ax = fabs(x);
    if (ax < 0.5) {
        const double y = x * x;
        const double c1 = -1.0 / 6.0, c2 = 1.0 / 120.0, c3 = -1.0 / 5040.0, c4 = 1.0 / 362880.0,
                     c5 = -1.0 / 39916800.0, c6 = 1.0 / 6227020800.0;
        return 1.0 + y * (c1 + y * (c2 + y * (c3 + y * (c4 + y * (c5 + y * c6)))));
    }
    return sin(x) / x;
}
