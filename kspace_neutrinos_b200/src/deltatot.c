/* Integrator state for the linear-response neutrino method and its per-step update
 * (host state machine; the integral itself is kernel K2 on the GPU).
 *
 * Mirrors delta_tot_table.c of the reference in behaviour, entry point by entry point:
 *   allocate_delta_tot_table :25-60   delta_tot_init :79-149     get_delta_nu_combined :153-172
 *   update_delta_tot :177-191         get_delta_nu_update :193-250 read_all_nu_state :254-301
 *   save_delta_tot :304-327           save_all_nu_state :331-352  save_nu_power :356-374
 *   fslength :394-407                 specialJ :457-463           get_delta_nu :507-611
 *   get_delta_tot :613-617
 * Differences, all deliberate (SURVEY Appendix A "quirks"): the row index is checked against
 * namax before a row is appended (the reference would write past the block), and the leaked
 * GSL objects of the reference do not exist here. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>
#include "ksn_host.h"

void allocate_delta_tot_table(_delta_tot_table *d_tot, const int nk_in, const double TimeTransfer, const double TimeMax, const double Omega0, const _omega_nu *const omnu, const double UnitTime_in_s, const double UnitLength_in_cm, int debug)
{
    d_tot->nk_allocated = nk_in;
    d_tot->nk = nk_in;
    d_tot->TimeTransfer = TimeTransfer;
    d_tot->namax = ceil(100 * (TimeMax - TimeTransfer)) + 2;
    d_tot->ia = 0;
    d_tot->delta_tot_init_done = 0;
    /* one block: scalefact[namax] followed by nk_in rows of namax values (k-major, a contiguous);
     * the device copy in K2 relies on exactly this layout */
    d_tot->delta_tot = (double **) mymalloc("kspace_delta_tot", nk_in * sizeof(double *));
    d_tot->scalefact = (double *) mymalloc("kspace_scalefact", d_tot->namax * (nk_in + 1) * sizeof(double));
    for (int k = 0; k < nk_in; k++) d_tot->delta_tot[k] = d_tot->scalefact + (size_t) d_tot->namax * (k + 1);
    /* page-lock the history block where there is a device: K2 then copies it up in one DMA, without a staging copy
     * (unregistered again in free_delta_tot_table; a failure only means the staged path is taken) */
    if (ksn_device_available()) ksn_host_register(d_tot->scalefact, (size_t) d_tot->namax * (nk_in + 1) * sizeof(double));
    d_tot->delta_nu_init = (double *) mymalloc("kspace_delta_nu_init", 3 * nk_in * sizeof(double));
    d_tot->delta_nu_last = d_tot->delta_nu_init + nk_in;
    d_tot->wavenum = d_tot->delta_nu_init + 2 * nk_in;
    d_tot->omnu = omnu;
    d_tot->light = LIGHTCGS * UnitTime_in_s / UnitLength_in_cm;
    d_tot->delta_nu_prefac = 1.5 * Omega0 * HUBBLE * HUBBLE * pow(UnitTime_in_s, 2) / d_tot->light;
    d_tot->Omeganonu = Omega0 - get_omega_nu(omnu, 1);
    d_tot->debug = debug;
}

void free_delta_tot_table(_delta_tot_table *d_tot)
{
    if (ksn_device_available()) ksn_host_unregister(d_tot->scalefact);
    myfree(d_tot->delta_tot);
    myfree(d_tot->scalefact);
    myfree(d_tot->delta_nu_init);
}

double get_delta_tot(const double delta_nu_curr, const double delta_cdm_curr, const double OmegaNua3, const double Omeganonu, const double Omeganu1, const double particle_nu_fraction)
{
    const double fcdm = 1 - OmegaNua3 / (Omeganonu + Omeganu1);
    return fcdm * (delta_cdm_curr + delta_nu_curr * OmegaNua3 / (Omeganonu + Omeganu1 * particle_nu_fraction));
}

void delta_tot_init(_delta_tot_table *const d_tot, const int nk_in, const double wavenum[], const double delta_cdm_curr[], const _transfer_init_table *const t_init, const double Time)
{
    const double logtime = log(Time);
    /* forget rows (read from a restart file) that lie at or after the current time */
    for (int i = 0; i < d_tot->ia; i++)
        if (logtime <= d_tot->scalefact[i]) { d_tot->ia = i; break; }
    if (Time > d_tot->TimeTransfer + 0.01 && d_tot->ia == 0)
        terminate(2023, "Did not read delta_tot from resume file, but we probably should have\n");
    if (Time < d_tot->TimeTransfer - 1e-5)
        terminate(2024, "Trying to compute delta_tot at a=%g < Transfer time of %g\n", Time, d_tot->TimeTransfer);
    if (nk_in > d_tot->nk_allocated)
        terminate(2011, "input power of %d is longer than memory of %d\n", nk_in, d_tot->nk_allocated);
    d_tot->nk = nk_in;
    const int nt = t_init->NPowerTable;
    if (log(wavenum[d_tot->nk - 1]) > t_init->logk[nt - 1])
        terminate(2, "Want k = %g but maximum in CAMB table is %g\n", wavenum[d_tot->nk - 1], exp(t_init->logk[nt - 1]));
    const double OmegaNua3 = get_omega_nu_nopart(d_tot->omnu, d_tot->TimeTransfer) * pow(d_tot->TimeTransfer, 3);
    const double OmegaNu1 = get_omega_nu(d_tot->omnu, 1);
    const double OmegaMa = d_tot->Omeganonu + OmegaNua3;
    const int fresh = d_tot->ia == 0;
    const double partnu = particle_nu_fraction(&d_tot->omnu->hybnu, d_tot->TimeTransfer, 0);
    /* natural cubic spline of T_nu/T_nonu against log k */
    double *c = mymalloc("transfer_spline", nt * sizeof(double));
    ksn_cspline_natural(t_init->logk, t_init->T_nu, nt, c);
    int hint = 0;
    for (int ik = 0; ik < d_tot->nk; ik++) {
        const double lk = log(wavenum[ik]);
        if (lk < t_init->logk[0] || lk > t_init->logk[nt - 1])
            terminate(2001, "GSL_ERROR in delta_tot_init: interpolation error, k=%g outside the transfer table\n", wavenum[ik]);
        const double T = ksn_cspline_eval(t_init->logk, t_init->T_nu, c, nt, lk, &hint);
        if (fresh)
            d_tot->delta_tot[ik][0] = get_delta_tot(delta_cdm_curr[ik] * T, delta_cdm_curr[ik], OmegaNua3, d_tot->Omeganonu, OmegaNu1, partnu);
        d_tot->delta_nu_init[ik] = d_tot->delta_tot[ik][0] * OmegaMa / (OmegaMa - OmegaNua3 + T * OmegaNua3) * fabs(T);
        d_tot->wavenum[ik] = wavenum[ik];
    }
    myfree(c);
    if (fresh) {
        d_tot->scalefact[0] = log(d_tot->TimeTransfer);
        d_tot->ia = 1;
    }
    if (d_tot->ThisTask == 0 && d_tot->debug) save_all_nu_state(d_tot, NULL);
    get_delta_nu_combined(d_tot, exp(d_tot->scalefact[d_tot->ia - 1]), wavenum, d_tot->delta_nu_last);
    d_tot->delta_tot_init_done = 1;
}

void update_delta_tot(_delta_tot_table *const d_tot, const double a, const double delta_cdm_curr[], const double delta_nu_curr[], const int overwrite)
{
    const double OmegaNua3 = get_omega_nu_nopart(d_tot->omnu, a) * pow(a, 3);
    const double OmegaNu1 = get_omega_nu(d_tot->omnu, 1);
    const double partnu = particle_nu_fraction(&d_tot->omnu->hybnu, a, 0);
    if (!overwrite) {
        if (d_tot->ia >= d_tot->namax)
            terminate(2036, "delta_tot table is full: %d rows stored, namax = %d (a=%g)\n", d_tot->ia, d_tot->namax, a);
        d_tot->ia++;
    }
    const int row = d_tot->ia - 1;
    d_tot->scalefact[row] = log(a);
    for (int ik = 0; ik < d_tot->nk; ik++)
        d_tot->delta_tot[ik][row] = get_delta_tot(delta_nu_curr[ik], delta_cdm_curr[ik], OmegaNua3, d_tot->Omeganonu, OmegaNu1, partnu);
}

void get_delta_nu_update(_delta_tot_table *const d_tot, const double a, const int nk_in, const double keff[], const double delta_cdm_curr[], double delta_nu_curr[], _transfer_init_table *transfer_init)
{
    if (!d_tot->delta_tot_init_done) delta_tot_init(d_tot, nk_in, keff, delta_cdm_curr, transfer_init, a);
    if (!d_tot->delta_tot_init_done) terminate(2001, "Should have called delta_tot_init first\n");
    if (nk_in != d_tot->nk) terminate(2002, "Number of kbins %d != stored delta_tot %d\n", nk_in, d_tot->nk);
    if (d_tot->nk < 2) terminate(2003, "Number of kbins is unreasonably small: %d\n", d_tot->nk);
    const int nk = d_tot->nk;
    /* a repeated call at the same scale factor returns the stored answer */
    if (log(a) - d_tot->scalefact[d_tot->ia - 1] < FLOAT_ACC) {
        memcpy(delta_nu_curr, d_tot->delta_nu_last, nk * sizeof(double));
        return;
    }
    /* provisional row from the previous delta_nu, integrate, then keep or drop the row */
    update_delta_tot(d_tot, a, delta_cdm_curr, d_tot->delta_nu_last, 0);
    get_delta_nu_combined(d_tot, a, keff, delta_nu_curr);
    memcpy(d_tot->delta_nu_last, delta_nu_curr, nk * sizeof(double));
    if (a >= exp(d_tot->scalefact[d_tot->ia - 2]) + 0.009) {
        update_delta_tot(d_tot, a, delta_cdm_curr, delta_nu_curr, 1);
        if (d_tot->ThisTask == 0 && d_tot->debug) save_delta_tot(d_tot, d_tot->ia - 1, NULL);
    } else {
        d_tot->ia--;
    }
    for (int ik = 0; ik < nk; ik++) {
        if (isnan(delta_nu_curr[ik]))
            terminate(2004, "delta_nu_curr=%g i=%d delta_cdm_curr=%g kk=%g\n", delta_nu_curr[ik], ik, delta_cdm_curr[ik], keff[ik]);
        if (delta_nu_curr[ik] < 0) delta_nu_curr[ik] = 0;
    }
}

void read_all_nu_state(_delta_tot_table *const d_tot, char *savefile)
{
    const char *path = savefile ? savefile : "delta_tot_nu.txt";
    FILE *fd = fopen(path, "r");
    if (!fd) return;
    int row;
    for (row = 0; row < d_tot->namax; row++) {
        double loga;
        if (fscanf(fd, "# %lg ", &loga) != 1) break;
        d_tot->scalefact[row] = loga;
        for (int ik = 0; ik < d_tot->nk; ik++) {
            if (fscanf(fd, "%lg ", &d_tot->delta_tot[ik][row]) == 1) continue;
            /* a short first line defines nk; a short later line is an error */
            if (row != 0)
                terminate(2006, "Expected %d k values, got %d for delta_tot in %s; a=%g\n", d_tot->nk, ik, path, exp(d_tot->scalefact[row]));
            d_tot->nk = ik;
            break;
        }
    }
    if (fabs(d_tot->scalefact[0] - log(d_tot->TimeTransfer)) > 1e-4)
        terminate(2007, "%s starts wih a=%g, transfer function is at a=%g\n", path, exp(d_tot->scalefact[0]), d_tot->TimeTransfer);
    if (row > 0) d_tot->ia = row;
    if (d_tot->debug) message(1, "Read %d stored power spectra from %s\n", row, path);
    fclose(fd);
}

void save_delta_tot(const _delta_tot_table *const d_tot, const int iia, char *savefile)
{
    const char *path = savefile ? savefile : "delta_tot_nu.txt";
    FILE *fd = fopen(path, "a");
    if (!fd) terminate(2012, "Could not open %s for writing!\n", path);
    fprintf(fd, "# %le ", d_tot->scalefact[iia]);
    for (int i = 0; i < d_tot->nk; i++) fprintf(fd, "%le ", d_tot->delta_tot[i][iia]);
    fprintf(fd, "\n");
    fclose(fd);
}

void save_all_nu_state(const _delta_tot_table *const d_tot, char *savefile)
{
    char *path = savefile ? savefile : "delta_tot_nu.txt";
    if (access(path, F_OK) != -1) {
        /* keep the previous file as <name>.bak */
        const size_t len = strlen(path) + 6;
        char *bak = mymalloc("filename2", len);
        if (bak) {
            snprintf(bak, len, "%s.bak", path);
            rename(path, bak);
            myfree(bak);
        }
    }
    for (int row = 0; row < d_tot->ia; row++) save_delta_tot(d_tot, row, path);
}

int save_nu_power(const _delta_tot_table *const d_tot, const double Time, const int snapnum, const char *OutputDir)
{
    char fname[1000];
    snprintf(fname, sizeof fname, "%s/powerspec_nu_%03d.txt", OutputDir, snapnum);
    FILE *fd = fopen(fname, "w");
    if (!fd) {
        fprintf(stderr, "can't open file `%s` for writing\n", fname);
        return -1;
    }
    fprintf(fd, "# k P_nu(k)\n");
    fprintf(fd, "# a = %g\n", Time);
    fprintf(fd, "# nbins = %d\n", d_tot->nk);
    for (int i = 0; i < d_tot->nk; i++)
        fprintf(fd, "%g %g\n", d_tot->wavenum[i], d_tot->delta_nu_last[i] * d_tot->delta_nu_last[i]);
    fclose(fd);
    return 0;
}

/* ---- scalar pieces kept on the host for API users and tests ------------------------------ */
static double inv_a2H(double loga, void *unused)
{
    (void) unused;
    const double a = exp(loga);
    return 1. / a / (a * hubble_function(a));
}

double fslength(const double logai, const double logaf, const double light)
{
    double val, err;
    if (logai >= logaf) return 0;
    const int st = ksn_qag61(inv_a2H, NULL, logai, logaf, 0, 1e-6, GSL_VAL, &val, &err);
    if (st) terminate(2001, "GSL_ERROR in fslength: quadrature status %d\n", st);
    return light * val;
}

static double J_fit(double x)
{
    if (x <= 0.) return 1.;
    const double x2 = x * x, x4 = x2 * x2, x8 = x4 * x4;
    return (1. + 0.0168 * x2 + 0.0407 * x4) / (1. + 2.1734 * x2 + 1.6787 * exp(4.1811 * log(x)) + 0.1467 * x8);
}

static double sph_j0(double x)
{
    if (fabs(x) < 0.5) {
        const double y = x * x;
        return 1.0 + y * (-1.0 / 6.0 + y * (1.0 / 120.0 + y * (-1.0 / 5040.0 + y * (1.0 / 362880.0 + y * (-1.0 / 39916800.0 + y * (1.0 / 6227020800.0))))));
    }
    return sin(x) / x;
}

/* truncated Fermi-Dirac transform for q > qc (hybrid neutrinos), delta_tot_table.c:431-454 */
static double J_truncated(double x, double qc, double frac_low)
{
    const double j0 = sph_j0(qc * x), cs = cos(qc * x), x2 = x * x;
    double sum = 0;
    for (int n = 1; n < 20; n++) {
        const double dn = n, n2 = dn * dn;
        const double poly = (n2 + n2 * dn * qc + dn * qc * x2 - x2) * qc * j0 + (2 * dn + n2 * qc + qc * x2) * cs;
        sum += ((n & 1) ? 1.0 : -1.0) * exp(-dn * qc) / (n2 + x2) / (n2 + x2) * poly;
    }
    return sum / (1.5 * 1.202056903159594 * (1 - frac_low));
}

double specialJ(const double x, const double qc, const double nufrac_low)
{
    return qc > 0 ? J_truncated(x, qc, nufrac_low) : J_fit(x);
}

/* ---- device dispatch ------------------------------------------------------------------------ */
static double hubble_cb(double a, void *unused) { (void) unused; return hubble_function(a); }

static struct { int valid; double a_lo, a_hi, probe_a[3], probe_h[3]; } bgc;

void ksn_invalidate_background(void) { bgc.valid = 0; }

void ksn_ensure_background(double a_lo, double a_hi)
{
    /* (the device table dies with ksn_shutdown -- also when the context moves to another device: ask the device layer) */
    if (bgc.valid && ksn_background_loaded() && a_lo >= bgc.a_lo && a_hi <= bgc.a_hi) {
        int same = 1;
        for (int i = 0; i < 3; i++) same &= hubble_function(bgc.probe_a[i]) == bgc.probe_h[i];
        if (same) return;
    }
    /* margin of a few table cells on both sides for the 4-point stencil */
    const double lo = fmin(a_lo, bgc.valid ? bgc.a_lo : a_lo), hi = fmax(a_hi, bgc.valid ? bgc.a_hi : a_hi);
    const double xlo = log(lo) - 0.01, xhi = log(hi) + 0.01;
    const char *env = getenv("KSN_BG_POINTS");                  /* experiment knob: table resolution */
    const int npts = env && atoi(env) >= 64 ? atoi(env) : 16384;
    const int rc = ksn_set_background(hubble_cb, NULL, xlo, xhi, npts);
    if (rc) ksn_fatal_device(rc, "ksn_set_background");
    bgc.valid = 1;
    bgc.a_lo = lo;
    bgc.a_hi = hi;
    bgc.probe_a[0] = lo; bgc.probe_a[1] = sqrt(lo * hi); bgc.probe_a[2] = hi;
    for (int i = 0; i < 3; i++) bgc.probe_h[i] = hubble_function(bgc.probe_a[i]);
}

/* The step's scale factor is known before its power spectrum is: have the a-only tables of K2 computed on a side stream
 * while K1 sweeps the grid (ksn_delta_nu_prefetch).  Predicts what get_delta_nu_update will ask for -- the stored rows plus
 * the provisional row at log(a) -- and is simply not used if the prediction is off (first step, repeated a, full table). */
void ksn_prefetch_delta_nu(const _delta_tot_table *const d_tot, const double a)
{
    if (!d_tot->delta_tot_init_done || d_tot->ia < 2 || d_tot->ia >= d_tot->namax || d_tot->namax > 4096) return;
    if (log(a) - d_tot->scalefact[d_tot->ia - 1] < FLOAT_ACC) return;          /* same a: the stored answer is returned */
    double sf[4096];
    memcpy(sf, d_tot->scalefact, sizeof(double) * d_tot->ia);
    sf[d_tot->ia] = log(a);
    const double a_hi = fmax(a, d_tot->TimeTransfer + (d_tot->namax - 2) / 100.);
    ksn_ensure_background(fmin(d_tot->TimeTransfer, a), a_hi);                  /* (the range delta_nu_on_device asks for) */
    ksn_delta_nu_prefetch(a, d_tot->TimeTransfer, d_tot->light, sf, d_tot->ia + 1, d_tot->namax);   /* failure: K2 computes them itself */
}

/* integrate `ns` species with masses mnu[] in one launch; out is species-major [ns][nk] */
static void delta_nu_on_device(const _delta_tot_table *const d_tot, const double a, const double wavenum[], int ns, const double mnu[], double *out)
{
    ksn_delta_nu_args A;
    memset(&A, 0, sizeof A);
    const _hybrid_nu *hyb = &d_tot->omnu->hybnu;
    const double partnu = particle_nu_fraction(hyb, a, 0);
    A.nk = d_tot->nk;
    A.Na = d_tot->ia;
    A.namax = d_tot->namax;
    A.nspecies = ns;
    A.a = a;
    A.TimeTransfer = d_tot->TimeTransfer;
    A.light = d_tot->light;
    A.delta_nu_prefac = d_tot->delta_nu_prefac;
    A.deriv_prefac = d_tot->TimeTransfer * (hubble_function(d_tot->TimeTransfer) / d_tot->light) * d_tot->TimeTransfer;
    A.nufrac_low0 = hyb->nufrac_low[0];
    for (int s = 0; s < ns; s++) {
        const double mnubykT = mnu[s] / d_tot->omnu->kBtnu;
        A.mnubykT[s] = mnubykT;
        A.relerr[s] = 1e-6;
        A.qc[s] = 0;
        A.integrate[s] = d_tot->ia > 1 && mnubykT > 0;
        if (partnu > 0) {
            if (1 - partnu < 1e-3) A.integrate[s] = 0;        /* everything is in particles */
            A.qc[s] = hyb->vcrit * mnubykT;
            A.relerr[s] /= (1. + 1e-5 - partnu);
        }
    }
    A.scalefact = d_tot->scalefact;
    A.delta_tot = d_tot->delta_tot[0];
    A.wavenum = wavenum;
    A.delta_nu_init = d_tot->delta_nu_init;
    if (d_tot->delta_tot[0] != d_tot->scalefact + d_tot->namax)
        terminate(3003, "delta_tot table is not in the contiguous layout of allocate_delta_tot_table\n");
    const double a_hi = fmax(a, d_tot->TimeTransfer + (d_tot->namax - 2) / 100.);
    ksn_ensure_background(fmin(d_tot->TimeTransfer, a), a_hi);
    if (d_tot->debug)
        message(0, "Start get_delta_nu: a=%g Na =%d wavenum[0]=%g delta_tot[0]=%g m_nu=%g\n", a, d_tot->ia, wavenum[0], d_tot->delta_tot[0][d_tot->ia - 1], mnu[0]);
    const int rc = ksn_delta_nu_integrate(&A, out, NULL);
    if (rc) ksn_fatal_device(rc, "get_delta_nu");
}

void get_delta_nu(const _delta_tot_table *const d_tot, const double a, const double wavenum[], double delta_nu_curr[], const double mnu)
{
    delta_nu_on_device(d_tot, a, wavenum, 1, &mnu, delta_nu_curr);
    if (d_tot->debug)
        for (int i = 0; i < 3; i++) message(0, "k %g d_nu %g\n", wavenum[d_tot->nk / 8 * i], delta_nu_curr[d_tot->nk / 8 * i]);
}

void get_delta_nu_combined(const _delta_tot_table *const d_tot, const double a, const double wavenum[], double delta_nu_curr[])
{
    const double Omega_nu_tot = get_omega_nu_nopart(d_tot->omnu, a);
    const int nk = d_tot->nk;
    double mnu[NUSPECIES], weight[NUSPECIES];
    int ns = 0;
    for (int mi = 0; mi < NUSPECIES; mi++) {
        if (d_tot->omnu->nu_degeneracies[mi] <= 0) continue;
        mnu[ns] = d_tot->omnu->RhoNuTab[mi]->mnu;
        weight[ns] = d_tot->omnu->nu_degeneracies[mi] * omega_nu_single(d_tot->omnu, a, mi);
        ns++;
    }
    memset(delta_nu_curr, 0, nk * sizeof(double));
    if (ns == 0) return;
    double *single = mymalloc("delta_nu_single", (size_t) ns * nk * sizeof(double));
    delta_nu_on_device(d_tot, a, wavenum, ns, mnu, single);
    for (int s = 0; s < ns; s++)
        for (int ik = 0; ik < nk; ik++) delta_nu_curr[ik] += single[(size_t) s * nk + ik] * weight[s] / Omega_nu_tot;
    myfree(single);
}
