/* Omega_nu(a) for up to three neutrino species (host scalars; SURVEY component 6).
 * Behaviour follows omega_nu_single.c of the reference line by line in meaning:
 *   init_omega_nu :16-51, get_omega_nu :55-65, get_omega_nu_nopart :69-74, get_omegag :77-81,
 *   rho_nu_init :117-153, rho_nu :171-202, nufrac_low :214-228, init_hybrid_nu :230-240,
 *   particle_nu_fraction :246-257, omega_nu_single :262-279.
 * Quadrature and splines are the in-tree ones (ksn_numeric.c), not GSL. */
#include <math.h>
#include <string.h>
#include "ksn_host.h"

#define HBAR_EVS 6.582119e-16
#define STEFAN_BOLTZMANN 5.670373e-5
#define GRAVITY 6.67408e-8
#define NRHOTAB 200
#define NU_SW 100
#define ZETA3 1.202056903159594
#define ZETA5 1.0369277551433704
#define ZETA7 1.0083492773819229
#define ZETA9 1.0020083928260826

/* (eV/c)^4 -> g/cm^3 for one species incl. antineutrinos, omega_nu_single.c:100-114 */
static double rho_unit(void)
{
    const double inv_hc = 1. / (2 * M_PI * LIGHTCGS * HBAR_EVS);
    double u = 4 * M_PI * 2;
    u *= inv_hc * inv_hc * inv_hc;
    u *= 1.60217646e-12 / LIGHTCGS / LIGHTCGS;
    return u;
}

struct fd_ctx { double amnu, kT; };

static double energy_density_kernel(double q, void *vp)
{
    const struct fd_ctx *p = vp;
    return q * q * sqrt(q * q + p->amnu * p->amnu) / (exp(q / p->kT) + 1);
}

/* The reference leaves these two external although no header declares them, and its own
 * omega_nu_single_test.c:68-85 links them to integrate the density exactly; params = {a m_nu, kT}
 * (omega_nu_single.c:89-96,100-115). */
double rho_nu_int(double q, void *params)
{
    const double *p = params;
    const struct fd_ctx c = { p[0], p[1] };
    return energy_density_kernel(q, (void *) &c);
}

double get_rho_nu_conversion(void)
{
    return rho_unit();
}

void rho_nu_init(_rho_nu_single *tab, double a0, const double mnu, const double HubbleParam, const double kBtnu)
{
    (void) HubbleParam;
    const double x_first = log(a0) - log(1.2);
    const double x_last = log(NU_SW * kBtnu / mnu) + log(1.2);
    tab->mnu = mnu;
    if (mnu < 1e-6 * kBtnu || x_last < x_first) return;      /* analytic limits suffice */
    tab->loga = mymalloc("rho_nu_table", 2 * NRHOTAB * sizeof(double));
    tab->acc = mymalloc("rho_nu_acc", sizeof(struct ksn_accel_s));
    tab->interp = mymalloc("rho_nu_interp", sizeof(struct ksn_interp_s));
    if (!tab->loga || !tab->acc || !tab->interp) terminate(2035, "Could not initialise tables for neutrino matter density\n");
    tab->rhonu = tab->loga + NRHOTAB;
    tab->interp->c = mymalloc("rho_nu_spline", NRHOTAB * sizeof(double));
    tab->interp->n = NRHOTAB;
    tab->interp->cubic = 1;
    tab->acc->hint = 0;
    for (int i = 0; i < NRHOTAB; i++) {
        struct fd_ctx p;
        double err, val;
        tab->loga[i] = x_first + i * (x_last - x_first) / (NRHOTAB - 1);
        p.amnu = mnu * exp(tab->loga[i]);
        p.kT = kBtnu;
        const int st = ksn_qag61(energy_density_kernel, &p, 0, 500 * kBtnu, 0, 1e-9, GSL_VAL, &val, &err);
        if (st) terminate(2001, "GSL_ERROR in rho_nu_init: quadrature status %d\n", st);
        tab->rhonu[i] = val / pow(exp(tab->loga[i]), 4) * rho_unit();
    }
    ksn_cspline_natural(tab->loga, tab->rhonu, NRHOTAB, tab->interp->c);
}

static double rho_nonrel(double a, double kT, double amnu, double r2)
{
    return amnu * (kT * kT * kT) / (a * a * a * a) *
           (1.5 * ZETA3 + r2 * 45. / 4. * ZETA5 + 2835. / 32. * r2 * r2 * ZETA7 + 80325 / 32. * r2 * r2 * r2 * ZETA9) * rho_unit();
}

static double rho_rel(double a, double kT)
{
    return 7 * pow(M_PI * kT / a, 4) / 120. * rho_unit();
}

double rho_nu(_rho_nu_single *tab, const double a, const double kT)
{
    const double amnu = a * tab->mnu;
    const double r2 = kT * kT / amnu / amnu;
    if (NU_SW * NU_SW * r2 < 1) return rho_nonrel(a, kT, amnu, r2);
    if (amnu < 1e-6 * kT) return rho_rel(a, kT);
    const double loga = log(a);
    if (!tab->loga || loga < tab->loga[0])
        return amnu < 1e-4 * kT ? rho_rel(a, kT) : rho_nonrel(a, kT, amnu, r2);
    if (loga > tab->loga[NRHOTAB - 1]) terminate(2001, "GSL_ERROR in rho_nu: interpolation error at a=%g\n", a);
    return ksn_cspline_eval(tab->loga, tab->rhonu, tab->interp->c, NRHOTAB, loga, &tab->acc->hint);
}

void init_omega_nu(_omega_nu *omnu, const double MNu[], const double a0, const double HubbleParam, const double tcmb0)
{
    omnu->hybnu.enabled = 0;
    omnu->tcmb0 = tcmb0;
    omnu->kBtnu = BOLEVK * TNUCMB * tcmb0;
    omnu->rhocrit = (3 * HUBBLE * HubbleParam * HUBBLE * HubbleParam) / (8 * M_PI * GRAVITY);
    /* species whose masses agree to FLOAT_ACC are folded into the first of them */
    for (int i = 0; i < NUSPECIES; i++) {
        int first = i;
        for (int j = 0; j < i; j++)
            if (fabs(MNu[i] - MNu[j]) < FLOAT_ACC) { first = j; break; }
        omnu->nu_degeneracies[i] = 0;
        omnu->nu_degeneracies[first] += 1;
    }
    for (int i = 0; i < NUSPECIES; i++) {
        omnu->RhoNuTab[i] = NULL;
        if (!omnu->nu_degeneracies[i]) continue;
        omnu->RhoNuTab[i] = mymalloc("RhoNuTab", sizeof(_rho_nu_single));
        memset(omnu->RhoNuTab[i], 0, sizeof(_rho_nu_single));
        rho_nu_init(omnu->RhoNuTab[i], a0, MNu[i], HubbleParam, omnu->kBtnu);
    }
    ksn_invalidate_background();
}

double get_omega_nu(const _omega_nu *const omnu, const double a)
{
    double rho = 0;
    for (int i = 0; i < NUSPECIES; i++)
        if (omnu->nu_degeneracies[i] > 0) rho += omnu->nu_degeneracies[i] * rho_nu(omnu->RhoNuTab[i], a, omnu->kBtnu);
    return rho / omnu->rhocrit;
}

double get_omega_nu_nopart(const _omega_nu *const omnu, const double a)
{
    const double all = get_omega_nu(omnu, a);
    const double in_particles = get_omega_nu(omnu, 1) * particle_nu_fraction(&omnu->hybnu, a, 0) / (a * a * a);
    return all - in_particles;
}

double get_omegag(const _omega_nu *const omnu, const double a)
{
    const double og0 = 4 * STEFAN_BOLTZMANN / (LIGHTCGS * LIGHTCGS * LIGHTCGS) * pow(omnu->tcmb0, 4) / omnu->rhocrit;
    return og0 / pow(a, 4);
}

static double fermi_dirac_q2(double x, void *unused)
{
    (void) unused;
    return x * x / (exp(x) + 1);
}

double nufrac_low(const double qc)
{
    double val, err;
    const int st = ksn_qag61(fermi_dirac_q2, NULL, 0, qc, 0, 1e-6, 100, &val, &err);
    if (st) terminate(2001, "GSL_ERROR in nufrac_low: quadrature status %d\n", st);
    return val / (1.5 * ZETA3);
}

void init_hybrid_nu(_hybrid_nu *const hybnu, const double mnu[], const double vcrit, const double light, const double nu_crit_time, const double kBtnu)
{
    hybnu->enabled = 1;
    hybnu->nu_crit_time = nu_crit_time;
    hybnu->vcrit = vcrit / light;
    for (int i = 0; i < NUSPECIES; i++) hybnu->nufrac_low[i] = nufrac_low(mnu[i] * vcrit / light / kBtnu);
    ksn_invalidate_background();
}

double particle_nu_fraction(const _hybrid_nu *const hybnu, const double a, const int i)
{
    if (!hybnu->enabled) return 0;
    return a > hybnu->nu_crit_time ? hybnu->nufrac_low[i] : 0;
}

double omega_nu_single(const _omega_nu *const omnu, const double a, int i)
{
    /* a species merged into a lower index reads that index's table */
    if (omnu->nu_degeneracies[i] == 0)
        for (int j = i; j >= 0; j--)
            if (omnu->nu_degeneracies[j]) { i = j; break; }
    const double now = rho_nu(omnu->RhoNuTab[i], a, omnu->kBtnu) / omnu->rhocrit;
    const double today = rho_nu(omnu->RhoNuTab[i], 1, omnu->kBtnu) / omnu->rhocrit;
    return now - today * particle_nu_fraction(&omnu->hybnu, a, i) / (a * a * a);
}
