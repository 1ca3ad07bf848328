/* FFTW2-slab interface (replaces interface_gadget.c): add_nu_power_to_rhogrid is the per-PM-step
 * entry a Gadget-style code calls right after its forward r2c FFT.
 *
 * One call = K1 (bin |delta(k)|^2 on the GPU, cross-rank sum) -> host glue (units, integrator
 * state machine; the integral is K2 on the GPU) -> K3 (scale every mode on the GPU).  For a
 * host-resident grid the slab is uploaded once, stays in HBM between K1 and K3, and is copied
 * back chunk by chunk as K3 finishes with it. */
#include <math.h>
#include <stdio.h>
#include <string.h>
#include "ksn_host.h"

double *delta_cdm_last;
static _delta_pow d_pow;

int set_kspace_vars(char tag[][50], void *addr[], int id[], int nt)
{
    static const struct { const char *name; int type; } names[] = {
        { "KspaceTransferFunction", STRING }, { "TimeTransfer", REAL }, { "InputSpectrum_UnitLength_in_cm", REAL },
        { "MNue", REAL }, { "MNum", REAL }, { "MNut", REAL },
        { "HybridNeutrinosOn", INT }, { "Vcrit", REAL }, { "NuPartTime", REAL } };
    void *targets[] = { kspace_params.KspaceTransferFunction, &kspace_params.TimeTransfer, &kspace_params.InputSpectrum_UnitLength_in_cm,
                        &kspace_params.MNu[0], &kspace_params.MNu[1], &kspace_params.MNu[2],
                        &kspace_params.hybrid_neutrinos_on, &kspace_params.vcrit, &kspace_params.nu_crit_time };
    for (size_t i = 0; i < sizeof names / sizeof names[0]; i++, nt++) {
        strcpy(tag[nt], names[i].name);
        addr[nt] = targets[i];
        id[nt] = names[i].type;
    }
    return nt;
}

struct step_ctx { double Time, BoxSize; int nk_allocated; };

/* Runs between K1 and K3: interface_gadget.c:92-101 followed by interface_common.c:125-148. */
static int between_passes(void *user, const double *power_sum, const double *keff_sum, const long long *count, double total_mass2,
                          const double **logkk, const double **ratio, int *nbins, double *norm)
{
    const struct step_ctx *s = user;
    const int nk_allocated = s->nk_allocated;
    double *delta_nu_curr = delta_cdm_curr + nk_allocated;
    double *keff = delta_cdm_curr + 2 * nk_allocated;
    long long *cnt = mymalloc("temp_modecount", nk_allocated * sizeof(long long));
    static int last_cap = 0;
    if (!delta_cdm_last || last_cap < nk_allocated) {
        /* the reference sizes this once (interface_gadget.c:85-86); re-size when the module is
         * re-initialised with more bins */
        delta_cdm_last = mymalloc("delta_cdm", nk_allocated * sizeof(double));
        last_cap = nk_allocated;
    }
    if (!cnt || !delta_cdm_last) terminate(1, "Could not allocate temporary memory for power spectra\n");
    memcpy(delta_cdm_curr, power_sum, nk_allocated * sizeof(double));
    memcpy(keff, keff_sum, nk_allocated * sizeof(double));
    memcpy(cnt, count, nk_allocated * sizeof(long long));
    const int nk_in = ksn_finish_powerspectrum(nk_allocated, total_mass2, delta_cdm_curr, cnt, keff);
    myfree(cnt);
    const double scale = pow(s->BoxSize, -3);
    for (int i = 0; i < nk_in; i++) {
        delta_cdm_curr[i] = sqrt(delta_cdm_curr[i] / scale);
        delta_cdm_last[i] = delta_cdm_curr[i];
        keff[i] *= (2 * M_PI / s->BoxSize);
    }
    d_pow = compute_neutrino_power_internal(s->Time, keff, delta_cdm_curr, delta_nu_curr, nk_in);
    for (int i = 0; i < d_pow.nbins; i++)
        if (isnan(d_pow.delta_ratio[i]) || isnan(d_pow.norm)) terminate(5, "delta_nu or delta_cdm is nan\n");
    *logkk = d_pow.logkk;
    *ratio = d_pow.delta_ratio;
    *nbins = d_pow.nbins;
    *norm = d_pow.norm;
    return 0;
}

/* asmth2 < 0: the reference's step.  asmth2 >= 0: the same step with the PM Green's function fused into K3. */
static void add_nu_power_any(int real_bytes, const double Time, const double BoxSize, void *grid, const int pmgrid, int slabstart_y, int nslab_y, double asmth2)
{
    const unsigned int *thr;
    const double *iw;
    struct step_ctx s = { Time, BoxSize, delta_tot_table.nk_allocated };
    if (ksn_bin_tables(pmgrid, s.nk_allocated, &thr, &iw)) terminate(1, "Could not allocate temporary memory for power spectra\n");
    ksn_prefetch_delta_nu(&delta_tot_table, Time);
    const int rc = asmth2 < 0 ? ksn_step_staged(grid, real_bytes, pmgrid, s.nk_allocated, slabstart_y, nslab_y, thr, iw, BoxSize, between_passes, &s)
                              : ksn_step_staged_greens(grid, real_bytes, pmgrid, s.nk_allocated, slabstart_y, nslab_y, thr, iw, BoxSize, between_passes, &s, asmth2);
    if (rc) ksn_fatal_device(rc, "add_nu_power_to_rhogrid");
    message(0, "Done adding neutrinos to grid on all processors\n");
    free_d_pow(&d_pow);
}

void add_nu_power_to_rhogrid_f64(const double Time, const double BoxSize, void *grid, const int pmgrid, int slabstart_y, int nslab_y, MPI_Comm comm)
{
    ksn_bind_comm(comm);
    add_nu_power_any(8, Time, BoxSize, grid, pmgrid, slabstart_y, nslab_y, -1.0);
}

void add_nu_power_to_rhogrid_f32(const double Time, const double BoxSize, void *grid, const int pmgrid, int slabstart_y, int nslab_y, MPI_Comm comm)
{
    ksn_bind_comm(comm);
    add_nu_power_any(4, Time, BoxSize, grid, pmgrid, slabstart_y, nslab_y, -1.0);
}

/* Extension: add_nu_power_to_rhogrid followed, in the same pass over the grid, by the multiplication
 * with the periodic Green's function and CIC deconvolution that Gadget-2's pmforce_periodic performs next
 * (pm_periodic.c, the loop after the hook of gadget-2/0002 patch:116-125).  asmth2 = (2 pi Asmth / BoxSize)^2, as there. */
void add_nu_power_and_greens_to_rhogrid_f64(const double Time, const double BoxSize, void *grid, const int pmgrid, int slabstart_y, int nslab_y, const double asmth2, MPI_Comm comm)
{
    if (!(asmth2 >= 0)) terminate(1, "add_nu_power_and_greens_to_rhogrid: asmth2 = %g\n", asmth2);
    ksn_bind_comm(comm);
    add_nu_power_any(8, Time, BoxSize, grid, pmgrid, slabstart_y, nslab_y, asmth2);
}

void add_nu_power_and_greens_to_rhogrid_f32(const double Time, const double BoxSize, void *grid, const int pmgrid, int slabstart_y, int nslab_y, const double asmth2, MPI_Comm comm)
{
    if (!(asmth2 >= 0)) terminate(1, "add_nu_power_and_greens_to_rhogrid: asmth2 = %g\n", asmth2);
    ksn_bind_comm(comm);
    add_nu_power_any(4, Time, BoxSize, grid, pmgrid, slabstart_y, nslab_y, asmth2);
}

/* No-neutrino P(k) path (interface_gadget.c:114-144): fills the module's d_pow for save_total_power. */
static void total_power_any(int real_bytes, const double BoxSize, void *grid, const int pmgrid, int slabstart_y, int nslab_y, MPI_Comm comm)
{
    const int nb = pmgrid / 2;
    if (!delta_cdm_curr) {
        /* four arrays of pmgrid/2: delta_cdm | delta_nu | keff | delta_cdm_last.  (The reference
         * computes the last offset as 3/2*pmgrid in integer arithmetic, which aliases keff.) */
        delta_cdm_curr = mymalloc("temp_power_spectrum", 4 * nb * sizeof(double));
        delta_cdm_last = delta_cdm_curr + 3 * nb;
    }
    double *delta_nu_curr = delta_cdm_curr + nb;
    double *keff = delta_cdm_curr + 2 * nb;
    long long *count = mymalloc("temp_modecount", nb * sizeof(long long));
    if (!count) terminate(1, "Could not allocate temporary memory for power spectra\n");
    const int nk_in = real_bytes == 8 ? total_powerspectrum_f64(pmgrid, grid, nb, slabstart_y, nslab_y, delta_cdm_curr, count, keff, comm)
                                      : total_powerspectrum_f32(pmgrid, grid, nb, slabstart_y, nslab_y, delta_cdm_curr, count, keff, comm);
    myfree(count);
    const double scale = pow(BoxSize, -3);
    for (int i = 0; i < nk_in; i++) {
        delta_cdm_curr[i] = sqrt(delta_cdm_curr[i] / scale);
        if (delta_cdm_last) delta_cdm_last[i] = delta_cdm_curr[i];
        delta_nu_curr[i] = 0;
        keff[i] = log(keff[i] * 2 * M_PI / BoxSize);
    }
    d_pow.delta_ratio = delta_nu_curr;
    d_pow.logkk = keff;
    d_pow.nbins = nk_in;
    d_pow.norm = 0;
}

void compute_total_power_spectrum_f64(const double Time, const double BoxSize, void *grid, const int pmgrid, int slabstart_y, int nslab_y, MPI_Comm comm)
{
    (void) Time;
    total_power_any(8, BoxSize, grid, pmgrid, slabstart_y, nslab_y, comm);
}

void compute_total_power_spectrum_f32(const double Time, const double BoxSize, void *grid, const int pmgrid, int slabstart_y, int nslab_y, MPI_Comm comm)
{
    (void) Time;
    total_power_any(4, BoxSize, grid, pmgrid, slabstart_y, nslab_y, comm);
}

int save_total_power(const double Time, const int snapnum, const char *OutputDir)
{
    char fname[1000];
    snprintf(fname, sizeof fname, "%s/powerspec_tot_%03d.txt", OutputDir, snapnum);
    FILE *fd = fopen(fname, "w");
    if (!fd) {
        fprintf(stderr, "can't open file `%s` for writing\n", fname);
        return -1;
    }
    const int with_nu = delta_tot_table.delta_tot_init_done;
    const double OmegaNua3 = with_nu ? OmegaNu_nopart(Time) * pow(Time, 3) : 0;
    const double OmegaNu1 = with_nu ? OmegaNu(1) : 0;
    const double partnu = with_nu ? particle_nu_fraction(&delta_tot_table.omnu->hybnu, Time, 0) : 0;
    fprintf(fd, "# k P_nu(k)\n");
    fprintf(fd, "# a = %g\n", Time);
    fprintf(fd, "# nbins = %d\n", d_pow.nbins);
    for (int i = 0; i < d_pow.nbins; i++) {
        const double dt = with_nu ? get_delta_tot(delta_tot_table.delta_nu_last[i], delta_cdm_last[i], OmegaNua3, delta_tot_table.Omeganonu, OmegaNu1, partnu)
                                  : delta_cdm_curr[i];
        fprintf(fd, "%g %g\n", exp(d_pow.logkk[i]), dt * dt);
    }
    fclose(fd);
    return 0;
}

/* link-level drop-in names (interface_gadget.h:38,48), bound to this library's default precision */
#ifdef KSN_DEFAULT_F32
void add_nu_power_to_rhogrid(const double Time, const double BoxSize, void *grid, const int pmgrid, int slabstart_y, int nslab_y, MPI_Comm comm)
    __attribute__((alias("add_nu_power_to_rhogrid_f32")));
void compute_total_power_spectrum(const double Time, const double BoxSize, void *grid, const int pmgrid, int slabstart_y, int nslab_y, MPI_Comm comm)
    __attribute__((alias("compute_total_power_spectrum_f32")));
#else
void add_nu_power_to_rhogrid(const double Time, const double BoxSize, void *grid, const int pmgrid, int slabstart_y, int nslab_y, MPI_Comm comm)
    __attribute__((alias("add_nu_power_to_rhogrid_f64")));
void compute_total_power_spectrum(const double Time, const double BoxSize, void *grid, const int pmgrid, int slabstart_y, int nslab_y, MPI_Comm comm)
    __attribute__((alias("compute_total_power_spectrum_f64")));
#endif
