/* TEST INFRASTRUCTURE -- CPU oracle for the per-PM-step k-space hot path.
 *
 * A plain-C restatement of the reference algorithm (sbird/kspace-neutrinos), each function citing
 * the reference file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's CPU
 * baseline may call it; the product (kspace_neutrinos_b200/) never links or imports it.
 *
 * Pinning: checked against the reference's own known answers (tests/test_oracle.py replays
 * powerspectrum_test.c, delta_tot_table_test.c, omega_nu_single_test.c, transfer_init_test.c values)
 * and, where /root/reference is present, against the reference sources compiled unmodified
 * (oracle/_ref) to ~1e-13.  Its quadrature/interpolation is oracle/mini_gsl.c, a restatement of the
 * GSL algorithms (GSL itself is absent; agreement with a real GSL build below ~1e-6 is unpinned).
 */
#ifndef KSN_ORACLE_H
#define KSN_ORACLE_H
#include <stddef.h>

#define ORC_NSPECIES 3

/* ---- Omega_nu(a)  (omega_nu_single.c) ---- */
typedef struct orc_species {
    double mnu;
    int tabulated;
    double loga[200], rho[200];
    void *spline, *acc;
} orc_species;

typedef struct orc_cosmo {
    orc_species sp[ORC_NSPECIES];
    int degeneracy[ORC_NSPECIES];
    double rhocrit, kBtnu, tcmb0;
    /* hybrid neutrinos */
    int hybrid_on;
    double nufrac_low[ORC_NSPECIES], nu_crit_time, vcrit;
    /* background used by orc_hubble (delta_tot_table_test.c:25-45) */
    double Omega_nonu, OmegaLambda, Hubble_internal;
} orc_cosmo;

void orc_cosmo_init(orc_cosmo *c, const double mnu[3], double a0, double hubble_param, double tcmb0);
void orc_cosmo_hybrid(orc_cosmo *c, const double mnu[3], double vcrit_kms, double nu_crit_time);
void orc_cosmo_background(orc_cosmo *c, double Omega0, double UnitTime_in_s);
double orc_omega_nu(const orc_cosmo *c, double a);
double orc_omega_nu_nopart(const orc_cosmo *c, double a);
double orc_omega_nu_single(const orc_cosmo *c, double a, int i);
double orc_omegag(const orc_cosmo *c, double a);
double orc_particle_nu_fraction(const orc_cosmo *c, double a, int i);
double orc_nufrac_low(double qc);
double orc_hubble(const orc_cosmo *c, double a);

/* ---- K1: powerspectrum.c:33-117 ---- */
/* raw per-slab sums (what each MPI rank holds before powerspectrum.c:91) */
void orc_powerspectrum_sums(int dims, const void *grid, int is_double, int nrbins, long long startslab, long long nslab,
                            double *power_sum, double *keff_sum, long long *count, double *total_mass2);
/* powerspectrum.c:96-116 */
int orc_powerspectrum_finish(int nrbins, double total_mass2, double *power, long long *count, double *keffs);
int orc_total_powerspectrum(int dims, const void *grid, int is_double, int nrbins, long long startslab, long long nslab,
                            double *power, long long *count, double *keffs);

/* ---- K3: interface_gadget.c:163-188 with delta_pow.c:19-37 ---- */
double orc_dnudcdm(const double *logkk, const double *ratio, int nbins, double norm, double logk);
void orc_scale_modes(void *grid, int is_double, int dims, long long startslab, long long nslab, double box,
                     const double *logkk, const double *ratio, int nbins, double norm);

/* ---- the Green's-function multiply that follows the hook in GADGET-2's pmforce_periodic (gadget2_greens.c) ---- */
void orc_gadget2_greens(void *grid, int is_double, int PMGRID, long long slabstart_y, long long nslab_y, double asmth2);

/* ---- K2 + state machine: delta_tot_table.c ---- */
typedef struct orc_dtot {
    int nk, nk_allocated, namax, ia, init_done;
    double delta_nu_prefac, Omeganonu, light, TimeTransfer;
    double *scalefact;       /* namax */
    double *delta_tot;       /* nk_allocated rows of namax */
    double *delta_nu_init, *delta_nu_last, *wavenum;
    const orc_cosmo *cosmo;
    unsigned long long n_evals;   /* integrand evaluations of the last orc_get_delta_nu */
} orc_dtot;

void orc_dtot_alloc(orc_dtot *d, int nk, double TimeTransfer, double TimeMax, double Omega0, const orc_cosmo *c,
                    double UnitTime_in_s, double UnitLength_in_cm);
void orc_dtot_free(orc_dtot *d);
int orc_dtot_read(orc_dtot *d, const char *path);
/* transfer table: logk[], T_nu/T_nonu[] of length nt (transfer_init.c) */
int orc_transfer_read(const char *path, double box, double UnitLength_in_cm, double InputUnit_in_cm, double **logk, double **tnu);
void orc_dtot_init(orc_dtot *d, int nk, const double *wavenum, const double *delta_cdm, const double *t_logk, const double *t_tnu, int nt, double Time);
double orc_fslength(const orc_cosmo *c, double logai, double logaf, double light);
double orc_specialJ(double x, double qc, double nufrac_low);
void orc_get_delta_nu(orc_dtot *d, double a, const double *wavenum, double *out, double mnu);
void orc_get_delta_nu_combined(orc_dtot *d, double a, const double *wavenum, double *out);
void orc_update_delta_tot(orc_dtot *d, double a, const double *delta_cdm, const double *delta_nu, int overwrite);
/* returns 0, or the reference's terminate() code */
int orc_get_delta_nu_update(orc_dtot *d, double a, int nk, const double *keff, const double *delta_cdm, double *delta_nu,
                            const double *t_logk, const double *t_tnu, int nt);

/* ---- one whole step: interface_gadget.c:75-102,158-194 + interface_common.c:125-148 ---- */
typedef struct orc_module {
    orc_cosmo cosmo;
    orc_dtot dtot;
    double *t_logk, *t_tnu;
    int nt;
    double *scratch;          /* 3 * nk_allocated */
    double last_prefac;
    int last_nk;
} orc_module;

int orc_module_init(orc_module *m, int nk_in, const double mnu[3], int hybrid_on, double vcrit, double nu_crit_time,
                    const char *transfer_file, double TimeTransfer, double box, double UnitTime_in_s, double UnitLength_in_cm,
                    double InputUnit_in_cm, double Omega0, double hubble_param, double tcmb0, double TimeMax);
int orc_add_nu_power_to_rhogrid(orc_module *m, double Time, double box, void *grid, int is_double, int pmgrid,
                                long long slabstart, long long nslab);
#endif
