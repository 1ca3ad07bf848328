/* TEST INFRASTRUCTURE -- CPU oracle (see ksn_oracle.h).  Serial, simple, slow on purpose.
 * Each block cites the reference lines it restates (paths under /root/reference).
 * Numerics come from oracle/mini_gsl.c (GSL restated).  Never linked into the product. */
#include "ksn_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "shim/gsl/gsl_integration.h"
#include "shim/gsl/gsl_interp.h"
#include "shim/gsl/gsl_sf_bessel.h"

#define LIGHT_CGS 2.99792458e10
#define BOLTZ_EVK 8.61734e-5
#define HUBBLE_CGS 3.24077929e-18
#define T_NU_OVER_T_CMB (pow(4 / 11., 1 / 3.) * 1.00328)
#define QAG_WS 200

/* =============================================================================== Omega_nu(a) */
static double ev4_to_gcm3(void)                        /* omega_nu_single.c:100-114 */
{
    double f = 4 * M_PI * 2;
    const double chbar = 1. / (2 * M_PI * LIGHT_CGS * 6.582119e-16);
    f *= chbar * chbar * chbar;
    f *= 1.60217646e-12 / LIGHT_CGS / LIGHT_CGS;
    return f;
}

static double rho_integrand(double q, void *p)         /* omega_nu_single.c:89-96 */
{
    const double amnu = ((double *) p)[0], kT = ((double *) p)[1];
    return q * q * sqrt(q * q + amnu * amnu) / (exp(q / kT) + 1);
}

static void species_init(orc_species *s, double a0, double mnu, double kBtnu)   /* omega_nu_single.c:117-153 */
{
    const double x0 = log(a0) - log(1.2), x1 = log(100 * kBtnu / mnu) + log(1.2);
    s->mnu = mnu;
    s->tabulated = 0;
    s->spline = s->acc = NULL;
    if (mnu < 1e-6 * kBtnu || x1 < x0) return;
    gsl_integration_workspace *w = gsl_integration_workspace_alloc(QAG_WS);
    gsl_function F;
    F.function = rho_integrand;
    for (int i = 0; i < 200; i++) {
        double par[2], err;
        s->loga[i] = x0 + i * (x1 - x0) / (200 - 1);
        par[0] = mnu * exp(s->loga[i]);
        par[1] = kBtnu;
        F.params = par;
        gsl_integration_qag(&F, 0, 500 * kBtnu, 0, 1e-9, QAG_WS, 6, w, &s->rho[i], &err);
        s->rho[i] = s->rho[i] / pow(exp(s->loga[i]), 4) * ev4_to_gcm3();
    }
    gsl_integration_workspace_free(w);
    gsl_interp *sp = gsl_interp_alloc(gsl_interp_cspline, 200);
    gsl_interp_init(sp, s->loga, s->rho, 200);
    s->spline = sp;
    s->acc = gsl_interp_accel_alloc();
    s->tabulated = 1;
}

static double rho_nonrel(double a, double kT, double amnu, double r2)           /* omega_nu_single.c:156-160 */
{
    return amnu * (kT * kT * kT) / (a * a * a * a) *
           (1.5 * 1.202056903159594 + r2 * 45. / 4. * 1.0369277551433704 + 2835. / 32. * r2 * r2 * 1.0083492773819229 +
            80325 / 32. * r2 * r2 * r2 * 1.0020083928260826) * ev4_to_gcm3();
}

static double rho_rel(double a, double kT) { return 7 * pow(M_PI * kT / a, 4) / 120. * ev4_to_gcm3(); }   /* :164-167 */

static double species_rho(const orc_species *s, double a, double kT)            /* omega_nu_single.c:171-202 */
{
    const double amnu = a * s->mnu, r2 = kT * kT / amnu / amnu;
    if (100 * 100 * r2 < 1) return rho_nonrel(a, kT, amnu, r2);
    if (amnu < 1e-6 * kT) return rho_rel(a, kT);
    const double la = log(a);
    if (!s->tabulated || la < s->loga[0]) return amnu < 1e-4 * kT ? rho_rel(a, kT) : rho_nonrel(a, kT, amnu, r2);
    return gsl_interp_eval(s->spline, s->loga, s->rho, la, s->acc);
}

void orc_cosmo_init(orc_cosmo *c, const double mnu[3], double a0, double h, double tcmb0)   /* omega_nu_single.c:16-51 */
{
    memset(c, 0, sizeof *c);
    c->tcmb0 = tcmb0;
    c->kBtnu = BOLTZ_EVK * T_NU_OVER_T_CMB * tcmb0;
    c->rhocrit = (3 * HUBBLE_CGS * h * HUBBLE_CGS * h) / (8 * M_PI * 6.67408e-8);
    for (int i = 0; i < ORC_NSPECIES; i++) {
        int j;
        c->degeneracy[i] = 0;
        for (j = 0; j < i; j++)
            if (fabs(mnu[i] - mnu[j]) < 1e-6) { c->degeneracy[j] += 1; break; }
        if (j == i) c->degeneracy[i] = 1;
    }
    for (int i = 0; i < ORC_NSPECIES; i++)
        if (c->degeneracy[i]) species_init(&c->sp[i], a0, mnu[i], c->kBtnu);
}

static double fd_kernel(double x, void *p) { (void) p; return x * x / (exp(x) + 1); }      /* omega_nu_single.c:207-210 */

double orc_nufrac_low(double qc)                                                           /* omega_nu_single.c:214-228 */
{
    gsl_integration_workspace *w = gsl_integration_workspace_alloc(100);
    gsl_function F = { fd_kernel, NULL };
    double v, e;
    gsl_integration_qag(&F, 0, qc, 0, 1e-6, 100, 6, w, &v, &e);
    gsl_integration_workspace_free(w);
    return v / (1.5 * 1.202056903159594);
}

void orc_cosmo_hybrid(orc_cosmo *c, const double mnu[3], double vcrit_kms, double nu_crit_time)   /* :230-240, interface_common.c:80-81 */
{
    const double light = LIGHT_CGS / 1e5;
    c->hybrid_on = 1;
    c->nu_crit_time = nu_crit_time;
    c->vcrit = vcrit_kms / light;
    for (int i = 0; i < ORC_NSPECIES; i++) c->nufrac_low[i] = orc_nufrac_low(mnu[i] * vcrit_kms / light / c->kBtnu);
}

double orc_particle_nu_fraction(const orc_cosmo *c, double a, int i)                       /* omega_nu_single.c:246-257 */
{
    if (!c->hybrid_on) return 0;
    return a > c->nu_crit_time ? c->nufrac_low[i] : 0;
}

double orc_omega_nu(const orc_cosmo *c, double a)                                          /* omega_nu_single.c:55-65 */
{
    double rho = 0;
    for (int i = 0; i < ORC_NSPECIES; i++)
        if (c->degeneracy[i] > 0) rho += c->degeneracy[i] * species_rho(&c->sp[i], a, c->kBtnu);
    return rho / c->rhocrit;
}

double orc_omega_nu_nopart(const orc_cosmo *c, double a)                                   /* omega_nu_single.c:69-74 */
{
    return orc_omega_nu(c, a) - orc_omega_nu(c, 1) * orc_particle_nu_fraction(c, a, 0) / (a * a * a);
}

double orc_omegag(const orc_cosmo *c, double a)                                            /* omega_nu_single.c:77-81 */
{
    return 4 * 5.670373e-5 / (LIGHT_CGS * LIGHT_CGS * LIGHT_CGS) * pow(c->tcmb0, 4) / c->rhocrit / pow(a, 4);
}

double orc_omega_nu_single(const orc_cosmo *c, double a, int i)                            /* omega_nu_single.c:262-279 */
{
    if (c->degeneracy[i] == 0)
        for (int j = i; j >= 0; j--)
            if (c->degeneracy[j]) { i = j; break; }
    double now = species_rho(&c->sp[i], a, c->kBtnu) / c->rhocrit;
    double part = species_rho(&c->sp[i], 1, c->kBtnu) / c->rhocrit;
    part *= orc_particle_nu_fraction(c, a, i) / (a * a * a);
    return now - part;
}

void orc_cosmo_background(orc_cosmo *c, double Omega0, double UnitTime_in_s)               /* delta_tot_table_test.c:25-31 */
{
    c->Omega_nonu = Omega0 - orc_omega_nu(c, 1);
    c->OmegaLambda = 1 - Omega0;
    c->Hubble_internal = HUBBLE_CGS * UnitTime_in_s;
}

double orc_hubble(const orc_cosmo *c, double a)                                            /* delta_tot_table_test.c:33-45 */
{
    double om = c->Omega_nonu / pow(a, 3) + c->OmegaLambda;
    om += orc_omega_nu(c, a);
    om += orc_omegag(c, a);
    return c->Hubble_internal * sqrt(om);
}

/* =============================================================================== K1 */
static double inv_window_1d(int k, int n)                                                  /* powerspectrum.c:8-12 */
{
    return k ? M_PI * k / (n * sin(M_PI * k / (double) n)) : 1.0;
}

static int kval(long long i, int n) { return i <= n / 2 ? (int) i : (int) (i - n); }       /* powerspectrum.c:27 */

void orc_powerspectrum_sums(int dims, const void *grid, int is_double, int nrbins, long long startslab, long long nslab,
                            double *power_sum, double *keff_sum, long long *count, double *total_mass2)
{
    /* powerspectrum.c:36-89.  The bin expression floor(binsperunit*log(kk)), binsperunit = (nrbins-1)/log(sqrt(3)*dims/2.0)
     * (:40,67) is written here as the reference's own -ffast-math build (Makefile:2) evaluates it -- log(sqrt(x)) folded
     * into 0.5*log(x), sqrt(3)*dims/2.0 into dims*(sqrt(3)/2); seen in the disassembly of oracle/_ref -- because the corner
     * mode (N/2,N/2,N/2) sits on the last bin edge to within one ulp and the two forms split at PMGRID=192. */
    const double halfbinsperunit = 0.5 * ((nrbins - 1) / log(dims * 0.8660254037844386));
    const int nzc = dims / 2 + 1;
    const double *gd = grid;
    const float *gf = grid;
    memset(power_sum, 0, sizeof(double) * nrbins);
    memset(keff_sum, 0, sizeof(double) * nrbins);
    memset(count, 0, sizeof(long long) * nrbins);
    *total_mass2 = 0;
    if (startslab == 0 && nslab > 0) {
        const double re = is_double ? gd[0] : gf[0], im = is_double ? gd[1] : gf[1];
        *total_mass2 = is_double ? re * re + im * im : (double) ((float) re * (float) re + (float) im * (float) im);
    }
    for (long long i = startslab; i < startslab + nslab; i++)
        for (int j = 0; j < dims; j++)
            for (int k = 0; k < nzc; k++) {
                const int ki = kval(i, dims), kj = kval(j, dims);
                const double kk = sqrt((double) ki * ki + (double) kj * kj + (double) k * k);
                if (!(kk > 0)) continue;
                const size_t idx = (size_t) ((i - startslab) * dims + j) * nzc + k;
                const int b = (int) floor(halfbinsperunit * log((double) ki * ki + (double) kj * kj + (double) k * k));
                const int mult = (k == 0 || k == dims / 2) ? 1 : 2;
                double mod2, win;
                if (is_double) {
                    const double w3 = inv_window_1d(ki, dims) * inv_window_1d(kj, dims) * inv_window_1d(k, dims);
                    const double w = w3 * w3;                  /* invwindow(): pow(.,2), :23 */
                    win = w * w;                               /* pow(invwindow,2), :68 */
                    mod2 = gd[2 * idx] * gd[2 * idx] + gd[2 * idx + 1] * gd[2 * idx + 1];
                } else {
                    /* fftw_real = float: the 1-D windows and their product are narrowed to float (:8,20-23) */
                    const float a = (float) inv_window_1d(ki, dims), bb = (float) inv_window_1d(kj, dims), c = (float) inv_window_1d(k, dims);
                    const float w3 = a * bb * c;
                    const float w = (float) ((double) w3 * (double) w3);
                    win = (double) w * (double) w;
                    mod2 = (double) (gf[2 * idx] * gf[2 * idx] + gf[2 * idx + 1] * gf[2 * idx + 1]);
                }
                power_sum[b] += mult * mod2 * win;
                keff_sum[b] += mult * kk;
                count[b] += mult;
            }
}

int orc_powerspectrum_finish(int nrbins, double total_mass2, double *power, long long *count, double *keffs)
{
    /* powerspectrum.c:98-116 */
    int nz = 0;
    for (int i = 0; i < nrbins; i++) {
        power[i] /= total_mass2;
        if (count[i]) { keffs[i] /= count[i]; power[i] /= count[i]; }
    }
    for (int i = 0; i < nrbins; i++)
        if (count[i]) {
            power[nz] = power[i]; keffs[nz] = keffs[i]; count[nz] = count[i];
            nz++;
        }
    return nz;
}

int orc_total_powerspectrum(int dims, const void *grid, int is_double, int nrbins, long long startslab, long long nslab,
                            double *power, long long *count, double *keffs)
{
    double m2;
    orc_powerspectrum_sums(dims, grid, is_double, nrbins, startslab, nslab, power, keffs, count, &m2);
    return orc_powerspectrum_finish(nrbins, m2, power, count, keffs);
}

/* =============================================================================== K3 */
double orc_dnudcdm(const double *logkk, const double *ratio, int nbins, double norm, double x)   /* delta_pow.c:19-37 */
{
    if (x < logkk[0]) x = logkk[0];
    if (x > logkk[nbins - 1]) x = logkk[nbins - 1];
    const size_t i = gsl_interp_bsearch(logkk, x, 0, nbins - 1);
    const double y = ratio[i] + (x - logkk[i]) / (logkk[i + 1] - logkk[i]) * (ratio[i + 1] - ratio[i]);
    return norm * y;
}

void orc_scale_modes(void *grid, int is_double, int dims, long long startslab, long long nslab, double box,
                     const double *logkk, const double *ratio, int nbins, double norm)
{
    /* interface_gadget.c:163-188 */
    const int nzc = dims / 2 + 1;
    double *gd = grid;
    float *gf = grid;
    for (long long y = startslab; y < startslab + nslab; y++)
        for (int x = 0; x < dims; x++)
            for (int z = 0; z < nzc; z++) {
                const double kx = x > dims / 2 ? x - dims : x, ky = y > dims / 2 ? y - dims : y, kz = z;
                double k2 = kx * kx + ky * ky + kz * kz;
                if (k2 <= 0) continue;
                k2 = log(sqrt(k2) * 2 * M_PI / box);
                const double smth = 1 + orc_dnudcdm(logkk, ratio, nbins, norm, k2);
                const size_t ip = ((size_t) (y - startslab) * dims + x) * nzc + z;
                if (is_double) { gd[2 * ip] *= smth; gd[2 * ip + 1] *= smth; }
                else { gf[2 * ip] = (float) (gf[2 * ip] * smth); gf[2 * ip + 1] = (float) (gf[2 * ip + 1] * smth); }
            }
}

/* =============================================================================== integrator */
void orc_dtot_alloc(orc_dtot *d, int nk, double TimeTransfer, double TimeMax, double Omega0, const orc_cosmo *c,
                    double UnitTime_in_s, double UnitLength_in_cm)
{
    /* delta_tot_table.c:25-60 */
    memset(d, 0, sizeof *d);
    d->nk = d->nk_allocated = nk;
    d->TimeTransfer = TimeTransfer;
    d->namax = ceil(100 * (TimeMax - TimeTransfer)) + 2;
    d->scalefact = calloc(d->namax, sizeof(double));
    d->delta_tot = calloc((size_t) d->namax * nk, sizeof(double));
    d->delta_nu_init = calloc(nk, sizeof(double));
    d->delta_nu_last = calloc(nk, sizeof(double));
    d->wavenum = calloc(nk, sizeof(double));
    d->cosmo = c;
    d->light = LIGHT_CGS * UnitTime_in_s / UnitLength_in_cm;
    d->delta_nu_prefac = 1.5 * Omega0 * HUBBLE_CGS * HUBBLE_CGS * pow(UnitTime_in_s, 2) / d->light;
    d->Omeganonu = Omega0 - orc_omega_nu(c, 1);
}

void orc_dtot_free(orc_dtot *d)
{
    free(d->scalefact); free(d->delta_tot); free(d->delta_nu_init); free(d->delta_nu_last); free(d->wavenum);
}

int orc_dtot_read(orc_dtot *d, const char *path)                                           /* delta_tot_table.c:254-301 */
{
    FILE *fd = fopen(path, "r");
    if (!fd) return 0;
    int row;
    for (row = 0; row < d->namax; row++) {
        double s;
        if (fscanf(fd, "# %lg ", &s) != 1) break;
        d->scalefact[row] = s;
        for (int k = 0; k < d->nk; k++)
            if (fscanf(fd, "%lg ", &d->delta_tot[(size_t) k * d->namax + row]) != 1) {
                if (row != 0) { fclose(fd); return -2006; }
                d->nk = k;
                break;
            }
    }
    fclose(fd);
    if (fabs(d->scalefact[0] - log(d->TimeTransfer)) > 1e-4) return -2007;
    if (row > 0) d->ia = row;
    return row;
}

int orc_transfer_read(const char *path, double box, double UnitLength_in_cm, double InputUnit_in_cm, double **logk, double **tnu)
{
    /* transfer_init.c:9-83 */
    const double scale = InputUnit_in_cm / UnitLength_in_cm, kmin = M_PI / box * scale;
    char line[1000];
    int n = 0, cap = 1024;
    FILE *fd = fopen(path, "r");
    if (!fd) return -2019;
    *logk = malloc(sizeof(double) * cap);
    *tnu = malloc(sizeof(double) * cap);
    while (fgets(line, sizeof line, fd)) {
        double k, c2, c3, c4, c5, t_nu, c7, t_nonu;
        if (line[0] == '#') continue;
        if (sscanf(line, " %lg %lg %lg %lg %lg %lg %lg %lg", &k, &c2, &c3, &c4, &c5, &t_nu, &c7, &t_nonu) != 8) break;
        if (!(k > kmin)) continue;
        if (n == cap) { cap *= 2; *logk = realloc(*logk, sizeof(double) * cap); *tnu = realloc(*tnu, sizeof(double) * cap); }
        (*tnu)[n] = t_nu / t_nonu;
        (*logk)[n] = log(k / scale);
        n++;
    }
    fclose(fd);
    return n;
}

static double total_delta(double dnu, double dcdm, double OmegaNua3, double Omeganonu, double Omeganu1, double partnu)
{
    /* get_delta_tot, delta_tot_table.c:613-617 */
    const double fcdm = 1 - OmegaNua3 / (Omeganonu + Omeganu1);
    return fcdm * (dcdm + dnu * OmegaNua3 / (Omeganonu + Omeganu1 * partnu));
}

void orc_update_delta_tot(orc_dtot *d, double a, const double *delta_cdm, const double *delta_nu, int overwrite)
{
    /* delta_tot_table.c:177-191 */
    const double OmegaNua3 = orc_omega_nu_nopart(d->cosmo, a) * pow(a, 3);
    const double OmegaNu1 = orc_omega_nu(d->cosmo, 1);
    const double partnu = orc_particle_nu_fraction(d->cosmo, a, 0);
    if (!overwrite) d->ia++;
    d->scalefact[d->ia - 1] = log(a);
    for (int k = 0; k < d->nk; k++)
        d->delta_tot[(size_t) k * d->namax + d->ia - 1] = total_delta(delta_nu[k], delta_cdm[k], OmegaNua3, d->Omeganonu, OmegaNu1, partnu);
}

/* ---- free-streaming length and J ---- */
static double fsl_integrand(double loga, void *p)                                          /* delta_tot_table.c:378-383 */
{
    const double a = exp(loga);
    return 1. / a / (a * orc_hubble(p, a));
}

double orc_fslength(const orc_cosmo *c, double logai, double logaf, double light)          /* delta_tot_table.c:394-407 */
{
    double v, e;
    if (logai >= logaf) return 0;
    gsl_integration_workspace *w = gsl_integration_workspace_alloc(QAG_WS);
    gsl_function F = { fsl_integrand, (void *) c };
    gsl_integration_qag(&F, logai, logaf, 0, 1e-6, QAG_WS, 6, w, &v, &e);
    gsl_integration_workspace_free(w);
    return light * v;
}

static double J_fit(double x)                                                              /* delta_tot_table.c:417-428 */
{
    if (x <= 0.) return 1.;
    const double x2 = x * x, x4 = x2 * x2, x8 = x4 * x4;
    return (1. + 0.0168 * x2 + 0.0407 * x4) / (1. + 2.1734 * x2 + 1.6787 * exp(4.1811 * log(x)) + 0.1467 * x8);
}

static double J_high(double x, double qc, double nufrac_low)                               /* delta_tot_table.c:431-454 */
{
    double integ = 0;
    for (int n = 1; n < 20; n++) {
        const double II = (n * n + n * n * n * qc + n * qc * x * x - x * x) * qc * gsl_sf_bessel_j0(qc * x) + (2 * n + n * n * qc + qc * x * x) * cos(qc * x);
        integ += -1 * pow((-1), n) * exp(-n * qc) / (n * n + x * x) / (n * n + x * x) * II;
    }
    return integ / (1.5 * 1.202056903159594 * (1 - nufrac_low));
}

double orc_specialJ(double x, double qc, double nufrac_low)                                /* delta_tot_table.c:457-463 */
{
    return qc > 0 ? J_high(x, qc, nufrac_low) : J_fit(x);
}

struct dnu_par {
    const orc_cosmo *c;
    double k, mnubykT, qc, nufrac_low;
    gsl_interp *sp, *fs_sp;
    gsl_interp_accel *acc, *fs_acc;
    const double *fsl, *fsx, *dt, *x;
    unsigned long long *evals;
};

static double dnu_integrand(double logai, void *vp)                                        /* delta_tot_table.c:492-500 */
{
    struct dnu_par *p = vp;
    const double fsl = gsl_interp_eval(p->fs_sp, p->fsx, p->fsl, logai, p->fs_acc);
    const double dtot = gsl_interp_eval(p->sp, p->x, p->dt, logai, p->acc);
    const double J = orc_specialJ(p->k * fsl / p->mnubykT, p->qc, p->nufrac_low);
    const double ai = exp(logai);
    (*p->evals)++;
    return fsl / (ai * orc_hubble(p->c, ai)) * J * dtot;
}

void orc_get_delta_nu(orc_dtot *d, double a, const double *wavenum, double *out, double mnu)
{
    /* delta_tot_table.c:507-611 */
    const orc_cosmo *c = d->cosmo;
    const int Na = d->ia;
    const double mnubykT = mnu / c->kBtnu;
    double qc = 0, relerr = 1e-6;
    const double loga0 = log(d->TimeTransfer), loga = log(a);
    const double fsl_A0a = orc_fslength(c, loga0, loga, d->light);
    const double deriv_prefac = d->TimeTransfer * (orc_hubble(c, d->TimeTransfer) / d->light) * d->TimeTransfer;
    d->n_evals = 0;
    for (int k = 0; k < d->nk; k++) {
        const double J = orc_specialJ(wavenum[k] * fsl_A0a / (mnubykT > 0 ? mnubykT : 1), qc, c->nufrac_low[0]);
        out[k] = J * d->delta_nu_init[k] * (1. + deriv_prefac * fsl_A0a);
    }
    const double partnu = orc_particle_nu_fraction(c, a, 0);
    if (partnu > 0) {
        if (1 - partnu < 1e-3) return;
        qc = c->vcrit * mnubykT;
        relerr /= (1. + 1e-5 - partnu);
    }
    if (!(Na > 1 && mnubykT > 0)) return;
    struct dnu_par p;
    const int Nfs = Na * 16;
    double *fsl = malloc(sizeof(double) * Nfs), *fsx = malloc(sizeof(double) * Nfs);
    for (int i = 0; i < Nfs; i++) {
        fsx[i] = loga0 + i * (loga - loga0) / (Nfs - 1.);
        fsl[i] = orc_fslength(c, fsx[i], loga, d->light);
    }
    p.c = c; p.mnubykT = mnubykT; p.qc = qc; p.nufrac_low = c->nufrac_low[0];
    p.fsl = fsl; p.fsx = fsx; p.x = d->scalefact; p.evals = &d->n_evals;
    p.acc = gsl_interp_accel_alloc();
    p.fs_acc = gsl_interp_accel_alloc();
    p.sp = gsl_interp_alloc(Na > 2 ? gsl_interp_cspline : gsl_interp_linear, Na);
    p.fs_sp = gsl_interp_alloc(gsl_interp_cspline, Nfs);
    gsl_interp_init(p.fs_sp, fsx, fsl, Nfs);
    gsl_integration_workspace *w = gsl_integration_workspace_alloc(QAG_WS);
    gsl_function F = { dnu_integrand, &p };
    for (int k = 0; k < d->nk; k++) {
        double v, e;
        p.k = wavenum[k];
        p.dt = d->delta_tot + (size_t) k * d->namax;
        gsl_interp_init(p.sp, p.x, p.dt, Na);
        gsl_integration_qag(&F, loga0, loga, 0, relerr, QAG_WS, 6, w, &v, &e);
        out[k] += d->delta_nu_prefac * v;
    }
    gsl_integration_workspace_free(w);
    gsl_interp_free(p.sp); gsl_interp_free(p.fs_sp);
    gsl_interp_accel_free(p.acc); gsl_interp_accel_free(p.fs_acc);
    free(fsl); free(fsx);
}

void orc_get_delta_nu_combined(orc_dtot *d, double a, const double *wavenum, double *out)
{
    /* delta_tot_table.c:153-172 */
    const double tot = orc_omega_nu_nopart(d->cosmo, a);
    double *one = malloc(sizeof(double) * d->nk);
    unsigned long long ev = 0;
    memset(out, 0, sizeof(double) * d->nk);
    for (int m = 0; m < ORC_NSPECIES; m++) {
        if (d->cosmo->degeneracy[m] <= 0) continue;
        const double om = d->cosmo->degeneracy[m] * orc_omega_nu_single(d->cosmo, a, m);
        orc_get_delta_nu(d, a, wavenum, one, d->cosmo->sp[m].mnu);
        ev += d->n_evals;
        for (int k = 0; k < d->nk; k++) out[k] += one[k] * om / tot;
    }
    d->n_evals = ev;
    free(one);
}

void orc_dtot_init(orc_dtot *d, int nk, const double *wavenum, const double *delta_cdm, const double *t_logk, const double *t_tnu, int nt, double Time)
{
    /* delta_tot_table.c:79-149 */
    for (int i = 0; i < d->ia; i++)
        if (log(Time) <= d->scalefact[i]) { d->ia = i; break; }
    d->nk = nk;
    const double OmegaNua3 = orc_omega_nu_nopart(d->cosmo, d->TimeTransfer) * pow(d->TimeTransfer, 3);
    const double OmegaNu1 = orc_omega_nu(d->cosmo, 1);
    gsl_interp *sp = gsl_interp_alloc(gsl_interp_cspline, nt);
    gsl_interp_accel *acc = gsl_interp_accel_alloc();
    gsl_interp_init(sp, t_logk, t_tnu, nt);
    for (int k = 0; k < nk; k++) {
        const double T = gsl_interp_eval(sp, t_logk, t_tnu, log(wavenum[k]), acc);
        const double OmegaMa = d->Omeganonu + OmegaNua3;
        if (d->ia == 0) {
            const double partnu = orc_particle_nu_fraction(d->cosmo, d->TimeTransfer, 0);
            d->delta_tot[(size_t) k * d->namax] = total_delta(delta_cdm[k] * T, delta_cdm[k], OmegaNua3, d->Omeganonu, OmegaNu1, partnu);
        }
        d->delta_nu_init[k] = d->delta_tot[(size_t) k * d->namax] * OmegaMa / (OmegaMa - OmegaNua3 + T * OmegaNua3) * fabs(T);
        d->wavenum[k] = wavenum[k];
    }
    gsl_interp_accel_free(acc);
    gsl_interp_free(sp);
    if (d->ia == 0) { d->scalefact[0] = log(d->TimeTransfer); d->ia = 1; }
    orc_get_delta_nu_combined(d, exp(d->scalefact[d->ia - 1]), wavenum, d->delta_nu_last);
    d->init_done = 1;
}

int orc_get_delta_nu_update(orc_dtot *d, double a, int nk, const double *keff, const double *delta_cdm, double *delta_nu,
                            const double *t_logk, const double *t_tnu, int nt)
{
    /* delta_tot_table.c:193-250 */
    if (!d->init_done) orc_dtot_init(d, nk, keff, delta_cdm, t_logk, t_tnu, nt, a);
    if (nk != d->nk) return 2002;
    if (d->nk < 2) return 2003;
    if (log(a) - d->scalefact[d->ia - 1] < 1e-6) {
        for (int k = 0; k < d->nk; k++) delta_nu[k] = d->delta_nu_last[k];
        return 0;
    }
    orc_update_delta_tot(d, a, delta_cdm, d->delta_nu_last, 0);
    orc_get_delta_nu_combined(d, a, keff, delta_nu);
    for (int k = 0; k < d->nk; k++) d->delta_nu_last[k] = delta_nu[k];
    if (a >= exp(d->scalefact[d->ia - 2]) + 0.009) orc_update_delta_tot(d, a, delta_cdm, delta_nu, 1);
    else d->ia--;
    for (int k = 0; k < d->nk; k++) {
        if (isnan(delta_nu[k])) return 2004;
        if (delta_nu[k] < 0) delta_nu[k] = 0;
    }
    return 0;
}

/* =============================================================================== whole step */
int orc_module_init(orc_module *m, int nk_in, const double mnu[3], int hybrid_on, double vcrit, double nu_crit_time,
                    const char *transfer_file, double TimeTransfer, double box, double UnitTime_in_s, double UnitLength_in_cm,
                    double InputUnit_in_cm, double Omega0, double hubble_param, double tcmb0, double TimeMax)
{
    /* interface_common.c:183-188 then :77-102 */
    memset(m, 0, sizeof *m);
    orc_cosmo_init(&m->cosmo, mnu, TimeTransfer, hubble_param, tcmb0);
    if (hybrid_on) orc_cosmo_hybrid(&m->cosmo, mnu, vcrit, nu_crit_time);
    orc_cosmo_background(&m->cosmo, Omega0, UnitTime_in_s);
    m->nt = orc_transfer_read(transfer_file, box, UnitLength_in_cm, InputUnit_in_cm, &m->t_logk, &m->t_tnu);
    if (m->nt < 0) return -m->nt;
    orc_dtot_alloc(&m->dtot, nk_in, TimeTransfer, TimeMax, Omega0, &m->cosmo, UnitTime_in_s, UnitLength_in_cm);
    m->scratch = calloc(3 * (size_t) nk_in, sizeof(double));
    return 0;
}

int orc_add_nu_power_to_rhogrid(orc_module *m, double Time, double box, void *grid, int is_double, int pmgrid,
                                long long slabstart, long long nslab)
{
    const int nka = m->dtot.nk_allocated;
    double *dcdm = m->scratch, *dnu = dcdm + nka, *keff = dcdm + 2 * nka;
    long long *count = malloc(sizeof(long long) * nka);
    /* interface_gadget.c:92-101 */
    const int nk = orc_total_powerspectrum(pmgrid, grid, is_double, nka, slabstart, nslab, dcdm, count, keff);
    free(count);
    const double scale = pow(box, -3);
    for (int i = 0; i < nk; i++) {
        dcdm[i] = sqrt(dcdm[i] / scale);
        keff[i] *= (2 * M_PI / box);
    }
    /* interface_common.c:125-148 */
    const int rc = orc_get_delta_nu_update(&m->dtot, Time, nk, keff, dcdm, dnu, m->t_logk, m->t_tnu, m->nt);
    if (rc) return rc;
    for (int i = 0; i < nk; i++) {
        keff[i] = log(keff[i]);
        dcdm[i] = dnu[i] / dcdm[i];
    }
    const double nop = orc_omega_nu_nopart(&m->cosmo, Time);
    const double hyb = orc_omega_nu(&m->cosmo, Time) - nop;
    const double prefac = nop / (m->dtot.Omeganonu / pow(Time, 3) + hyb);
    m->last_prefac = prefac;
    m->last_nk = nk;
    /* interface_gadget.c:163-188 */
    orc_scale_modes(grid, is_double, pmgrid, slabstart, nslab, box, keff, dcdm, nk, prefac);
    return 0;
}
