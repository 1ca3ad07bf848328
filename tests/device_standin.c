/* TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT, never linked into libkspace_neutrinos_b200.so.
 *
 * A CPU stand-in for the device entry points the product's HOST layer calls (include/ksn_b200.h:
 * ksn_set_background, ksn_delta_nu_integrate, ksn_powerspectrum_sums, ksn_step_staged, ksn_step_staged_greens,
 * ksn_last_error, and the ksn_comm_* / ksn_init family the -DKSN_HAVE_MPI build uses to pick its collective), written on top of the CPU oracle (oracle/ksn_oracle.c, oracle/mini_gsl.c).  CPU tests link it with
 * kspace_neutrinos_b200/src/ *.c -- instead of the CUDA objects -- so that the host layer's own logic (the per-step
 * state machine of get_delta_nu_update, the table layout, save/resume files, the glue either side of the kernels) can be
 * run in a container with no GPU, e.g. under the reference's own cmocka programs (tests/test_reference_programs.py).
 * What it says about the kernels: nothing; those are checked on the GPU against the same oracle (tests -m gpu).
 *
 * Contract followed for the integral (ksn_b200.h, ksn_delta_nu_args; reference delta_tot_table.c:507-611): per species,
 * the initial-condition term with qc = 0, then -- if integrate[s] -- QAG(key 6, epsrel = relerr[s]) of
 * fs/(a H) * J(k fs / mnubykT, qc[s]) * delta_tot(log a) over [log a0, log a], times delta_nu_prefac. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <gsl/gsl_integration.h>
#include <gsl/gsl_interp.h>
#include "ksn_b200.h"
#include "ksn_oracle.h"

#define WS 200

static char errbuf[256] = "";
const char *ksn_last_error(void) { return errbuf; }

/* There is no device here: everything that needs one says so (the host layer's backend bootstrap under -DKSN_HAVE_MPI
 * must then fall back, collectively, to the host call-back).  KSN_STANDIN_P2P=1 lets the peer-memory entry points
 * "succeed" instead -- the sums then travel through the MPI shim -- so that the bootstrap's handshake (handles gathered,
 * trial sum, all-ranks verdicts) can be walked through on the CPU; KSN_STANDIN_P2P_FAIL_RANK=r makes rank r alone fail. */
#ifdef KSN_HAVE_MPI
#include <mpi.h>
#endif
static ksn_allreduce_fn comm_fn;
static void *comm_user;
static int comm_n = 1, comm_r = 0, fake_p2p = 0, exported = 0;
static int nodev(const char *what) { snprintf(errbuf, sizeof errbuf, "%s: no CUDA device (CPU stand-in)", what); return KSN_ENODEV; }
static int standin_p2p(void)
{
#ifdef KSN_HAVE_MPI
    const char *e = getenv("KSN_STANDIN_P2P"), *f = getenv("KSN_STANDIN_P2P_FAIL_RANK");
    int rank = 0;
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    return e && atoi(e) && !(f && atoi(f) == rank);
#else
    return 0;
#endif
}
int ksn_init(int device) { (void) device; return nodev("ksn_init"); }
int ksn_device(void) { return -1; }
int ksn_device_available(void) { return standin_p2p(); }
int ksn_device_count(void) { return standin_p2p() ? 1 : 0; }
int ksn_comm_single(void) { comm_fn = NULL; comm_user = NULL; comm_n = 1; comm_r = 0; fake_p2p = 0; return KSN_OK; }
int ksn_comm_nccl_unique_id(void *id128) { (void) id128; return nodev("ksn_comm_nccl_unique_id"); }
int ksn_comm_nccl_init(const void *id128, int nranks, int rank) { (void) id128; (void) nranks; (void) rank; return nodev("ksn_comm_nccl_init"); }
int ksn_comm_p2p_export(void *handle64)
{
    if (!standin_p2p()) return nodev("ksn_comm_p2p_export");
    memset(handle64, 0, 64);
    strcpy((char *) handle64, "standin mailbox");
    exported = 1;
    return KSN_OK;
}
int ksn_comm_p2p_init(const void *handles, int nranks, int rank)
{
    if (!exported) { snprintf(errbuf, sizeof errbuf, "ksn_comm_p2p_init: call ksn_comm_p2p_export first"); return KSN_EINVAL; }
    for (int r = 0; r < nranks; r++)
        if (strcmp((const char *) handles + 64 * r, "standin mailbox")) { snprintf(errbuf, sizeof errbuf, "ksn_comm_p2p_init: bad handle of rank %d", r); return KSN_ECOMM; }
    ksn_comm_single();
    comm_n = nranks; comm_r = rank; fake_p2p = 1;
    return KSN_OK;
}
int ksn_comm_host_callback(ksn_allreduce_fn fn, void *user, int nranks, int rank) { ksn_comm_single(); comm_fn = fn; comm_user = user; comm_n = nranks; comm_r = rank; return KSN_OK; }
int ksn_comm_allreduce_host(double *buf, size_t n)
{
    if (comm_n <= 1) return KSN_OK;
#ifdef KSN_HAVE_MPI
    if (fake_p2p) return MPI_Allreduce(MPI_IN_PLACE, buf, (int) n, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD) == MPI_SUCCESS ? KSN_OK : KSN_ECOMM;
#endif
    return comm_fn ? comm_fn(buf, n, comm_user) : KSN_ECOMM;
}
int ksn_comm_size(void) { return comm_n; }
int ksn_comm_rank(void) { return comm_r; }
/* which backend the host layer ended up with (tests): 0 none, 1 host call-back, 2 the pretended peer memory */
int ksn_standin_backend(void) { return comm_n <= 1 ? 0 : fake_p2p ? 2 : 1; }

static ksn_hubble_fn bg_hub;
static void *bg_user;
static double bg_lo, bg_hi;

int ksn_set_background(ksn_hubble_fn hub, void *user, double loga_lo, double loga_hi, int n)
{
    if (!hub || !(loga_hi > loga_lo) || n < 4) { snprintf(errbuf, sizeof errbuf, "ksn_set_background: bad arguments"); return KSN_EINVAL; }
    bg_hub = hub; bg_user = user; bg_lo = loga_lo; bg_hi = loga_hi;
    return KSN_OK;
}

int ksn_background_loaded(void) { return bg_hub != NULL; }
/* no device here: nothing to page-lock, nothing to prefetch */
int ksn_host_register(void *ptr, size_t bytes) { (void) ptr; (void) bytes; return KSN_ENODEV; }
int ksn_host_unregister(void *ptr) { (void) ptr; return KSN_ENODEV; }
int ksn_delta_nu_prefetch(double a, double a0, double light, const double *sf, int Na, int namax)
{
    (void) a; (void) a0; (void) light; (void) sf; (void) Na; (void) namax;
    return KSN_ENODEV;
}

static double inv_a2H(double loga, void *unused)                                   /* delta_tot_table.c:378-384 */
{
    (void) unused;
    const double a = exp(loga);
    return 1. / (a * a * bg_hub(a, bg_user));
}

static double fsl(double logai, double logaf, double light)                        /* delta_tot_table.c:395-407 */
{
    double v, e;
    if (logai >= logaf) return 0;
    gsl_integration_workspace *w = gsl_integration_workspace_alloc(WS);
    gsl_function F = { inv_a2H, NULL };
    gsl_integration_qag(&F, logai, logaf, 0, 1e-6, WS, 6, w, &v, &e);
    gsl_integration_workspace_free(w);
    return light * v;
}

struct par {
    double k, mnubykT, qc, nufrac_low;
    gsl_interp *sp, *fs_sp;
    gsl_interp_accel *acc, *fs_acc;
    const double *fsv, *fsx, *dt, *x;
    unsigned long long evals;
};

static double integrand(double logai, void *vp)                                    /* delta_tot_table.c:492-500 */
{
    struct par *p = vp;
    const double f = gsl_interp_eval(p->fs_sp, p->fsx, p->fsv, logai, p->fs_acc);
    const double dtot = gsl_interp_eval(p->sp, p->x, p->dt, logai, p->acc);
    const double ai = exp(logai);
    p->evals++;
    return f / (ai * bg_hub(ai, bg_user)) * orc_specialJ(p->k * f / p->mnubykT, p->qc, p->nufrac_low) * dtot;
}

int ksn_delta_nu_integrate(const ksn_delta_nu_args *A, double *out, unsigned long long *n_evals)
{
    if (!bg_hub) { snprintf(errbuf, sizeof errbuf, "ksn_delta_nu_integrate: no background table"); return KSN_EINVAL; }
    const double loga0 = log(A->TimeTransfer), loga = log(A->a);
    if (loga0 < bg_lo || loga > bg_hi) { snprintf(errbuf, sizeof errbuf, "ksn_delta_nu_integrate: a outside the background table"); return KSN_EINVAL; }
    const double fsl_A0a = fsl(loga0, loga, A->light);
    unsigned long long ev = 0;
    for (int s = 0; s < A->nspecies; s++) {
        double *o = out + (size_t) s * A->nk;
        const double m = A->mnubykT[s];
        for (int k = 0; k < A->nk; k++)
            o[k] = orc_specialJ(A->wavenum[k] * fsl_A0a / (m > 0 ? m : 1), 0, A->nufrac_low0) * A->delta_nu_init[k] * (1. + A->deriv_prefac * fsl_A0a);
        if (!A->integrate[s]) continue;
        const int Na = A->Na, Nfs = 16 * Na;
        struct par p;
        double *fsv = malloc(sizeof(double) * Nfs), *fsx = malloc(sizeof(double) * Nfs);
        for (int i = 0; i < Nfs; i++) {
            fsx[i] = loga0 + i * (loga - loga0) / (Nfs - 1.);
            fsv[i] = fsl(fsx[i], loga, A->light);
        }
        p.mnubykT = m; p.qc = A->qc[s]; p.nufrac_low = A->nufrac_low0; p.fsv = fsv; p.fsx = fsx; p.x = A->scalefact; p.evals = 0;
        p.acc = gsl_interp_accel_alloc();
        p.fs_acc = gsl_interp_accel_alloc();
        p.sp = gsl_interp_alloc(Na > 2 ? gsl_interp_cspline : gsl_interp_linear, Na);
        p.fs_sp = gsl_interp_alloc(gsl_interp_cspline, Nfs);
        gsl_interp_init(p.fs_sp, fsx, fsv, Nfs);
        gsl_integration_workspace *w = gsl_integration_workspace_alloc(WS);
        gsl_function F = { integrand, &p };
        for (int k = 0; k < A->nk; k++) {
            double v, e;
            p.k = A->wavenum[k];
            p.dt = A->delta_tot + (size_t) k * A->namax;
            gsl_interp_init(p.sp, p.x, p.dt, Na);
            gsl_integration_qag(&F, loga0, loga, 0, A->relerr[s], WS, 6, w, &v, &e);
            o[k] += A->delta_nu_prefac * v;
        }
        ev += p.evals;
        gsl_integration_workspace_free(w);
        gsl_interp_free(p.sp); gsl_interp_free(p.fs_sp);
        gsl_interp_accel_free(p.acc); gsl_interp_accel_free(p.fs_acc);
        free(fsv); free(fsx);
    }
    if (n_evals) *n_evals = ev;
    return KSN_OK;
}

int ksn_powerspectrum_sums(const void *grid, int real_bytes, int dims, int nrbins, long long startslab, long long nslab,
                           const unsigned int *thresholds, const double *invwin,
                           double *power_sum, double *keff_sum, long long *count, double *total_mass2)
{
    (void) thresholds; (void) invwin;            /* the oracle bins with log() as powerspectrum.c:33-89 does */
    if (real_bytes != 4 && real_bytes != 8) { snprintf(errbuf, sizeof errbuf, "ksn_powerspectrum_sums: real_bytes = %d", real_bytes); return KSN_EINVAL; }
    orc_powerspectrum_sums(dims, grid, real_bytes == 8, nrbins, startslab, nslab, power_sum, keff_sum, count, total_mass2);
    if (comm_n > 1) {                            /* powerspectrum.c:91-95: the sums over all ranks (counts are exact in a double) */
        double *b = malloc(sizeof(double) * (3 * (size_t) nrbins + 1));
        for (int i = 0; i < nrbins; i++) { b[i] = power_sum[i]; b[nrbins + i] = keff_sum[i]; b[2 * nrbins + i] = (double) count[i]; }
        b[3 * nrbins] = *total_mass2;
        const int rc = ksn_comm_allreduce_host(b, 3 * (size_t) nrbins + 1);
        for (int i = 0; i < nrbins; i++) { power_sum[i] = b[i]; keff_sum[i] = b[nrbins + i]; count[i] = (long long) b[2 * nrbins + i]; }
        *total_mass2 = b[3 * nrbins];
        free(b);
        if (rc) { snprintf(errbuf, sizeof errbuf, "ksn_powerspectrum_sums: all-reduce call-back failed (%d)", rc); return KSN_EINVAL; }
    }
    return KSN_OK;
}

int ksn_step_staged(void *hgrid, int real_bytes, int dims, int nrbins, long long startslab, long long nslab,
                    const unsigned int *thresholds, const double *invwin, double boxsize, ksn_between_fn between, void *user)
{
    double *ps = calloc(2 * (size_t) nrbins, sizeof(double)), *ks = ps + nrbins, tm2 = 0, norm = 0;
    long long *cnt = calloc(nrbins, sizeof(long long));
    const double *logkk = NULL, *ratio = NULL;
    int nbins = 0;
    int rc = ksn_powerspectrum_sums(hgrid, real_bytes, dims, nrbins, startslab, nslab, thresholds, invwin, ps, ks, cnt, &tm2);
    if (!rc && between(user, ps, ks, cnt, tm2, &logkk, &ratio, &nbins, &norm)) rc = KSN_EINVAL;
    if (!rc) orc_scale_modes(hgrid, real_bytes == 8, dims, startslab, nslab, boxsize, logkk, ratio, nbins, norm);
    free(ps); free(cnt);
    return rc;
}

int ksn_step_staged_greens(void *hgrid, int real_bytes, int dims, int nrbins, long long startslab, long long nslab,
                           const unsigned int *thresholds, const double *invwin, double boxsize, ksn_between_fn between, void *user, double asmth2)
{
    (void) hgrid; (void) real_bytes; (void) dims; (void) nrbins; (void) startslab; (void) nslab; (void) thresholds; (void) invwin;
    (void) boxsize; (void) between; (void) user; (void) asmth2;
    snprintf(errbuf, sizeof errbuf, "ksn_step_staged_greens: not part of the CPU stand-in");
    return KSN_EINVAL;
}
