/* TEST/BENCH INFRASTRUCTURE: times the REFERENCE's own CPU implementation of the per-PM-step hot
 * path (add_nu_power_to_rhogrid, interface_gadget.c:158-194) on this host's cores.
 *
 * Linked against the reference sources compiled unmodified from /root/reference plus the shims
 * (oracle/Makefile target _ref/ref_bench).  R forked ranks (mini-MPI) each own a slab that starts
 * where their slab of the full PMGRID^3 grid starts and is P planes deep: P = N/R is the whole
 * slab (nothing extrapolated), a smaller P a bounded sample of it (both grid passes are linear in
 * the number of modes).  The linear-response integral is replicated on every rank exactly as in
 * the reference and is always timed in full, at a ~100-row stored history.  Phase boundaries come
 * from the reference's own progress messages (ref_host.c records their time).
 *
 * The grid is the one bench.py's GPU arm fills its slabs with (ksn_fill_synthetic_grid,
 * csrc/ksn_device.cu): the same counter-based generator keyed by (seed, global mode index), restated
 * in oracle/synthetic_grid.h -- a Gaussian field with P(k) ~ k^slope, element (0,0,0) = N^3.  (Equal values up to the last
 * bits of libm's log/sincos against the device's.)
 *
 * usage: ref_bench N P R STEPS HYBRID TRANSFER_FILE [WARMUP [BUDGET_S [m0 m1 m2]]]
 *   WARMUP   untimed steps before the STEPS timed ones (default 1)
 *   BUDGET_S > 0: wall-clock budget for the WARMUP + STEPS steps.  The initialising call at a = TimeTransfer (same grid
 *            passes, a trivial integral) shows what a step costs: if WARMUP + STEPS of them would not fit, the planes per
 *            rank are cut -- for EVERY step, warm-up included, and the module is initialised again on the shallower slabs:
 *            the set of non-empty k bins depends on the slab depth, and the reference terminates when it changes between
 *            steps (delta_tot_table.c:198, "Number of kbins ... != stored delta_tot").  The JSON says how many planes
 *            the steps used.
 * prints one JSON line on rank 0.
 */
#include "synthetic_grid.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "shim/mpi.h"
#include "interface_gadget.h"
#include "delta_tot_table.h"
#include "omega_nu_single.h"

extern int ThisTask;
extern int ksn_ref_quiet;
extern double ksn_ref_t_mass, ksn_ref_t_nupower;      /* wall-clock stamps taken in message() */
extern _delta_tot_table delta_tot_table;
void ksn_ref_set_background(const _omega_nu *omnu, double Omega0, double UnitTime_in_s);

static double now(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

#define SEED 20261017ull     /* host.DeviceGrid.fill_synthetic's default */
#define SLOPE (-1.0)

int main(int argc, char **argv)
{
    if (argc < 7) { fprintf(stderr, "usage: %s N P R STEPS HYBRID TRANSFER_FILE [WARMUP [BUDGET_S [m0 m1 m2]]]\n", argv[0]); return 2; }
    const int N = atoi(argv[1]), R = atoi(argv[3]), steps = atoi(argv[4]), hybrid = atoi(argv[5]);
    int P = atoi(argv[2]);
    const char *transfer = argv[6];
    const int warmup = argc > 7 ? atoi(argv[7]) : 1;
    const double budget = argc > 8 ? atof(argv[8]) : 0;
    double mnu[3] = { 0.1, 0.1, 0.1 };
    if (argc > 11) for (int i = 0; i < 3; i++) mnu[i] = atof(argv[9 + i]);
    const double UL = 3.085678e21, UT = UL / 1e5, BOX = 512000, OMEGA0 = 0.2793;
    const int L = N / 2 + 1;
    const int total = warmup + steps;
    double *times = ksn_minimpi_shared_alloc(sizeof(double) * 4 * (total + 1));
    int *planes = ksn_minimpi_shared_alloc(sizeof(int) * (total + 2));
    const int rank = ksn_minimpi_fork(R);
    ThisTask = rank;
    ksn_ref_quiet = 2;                  /* record time stamps, print nothing */
    strncpy(kspace_params.KspaceTransferFunction, transfer, 499);
    kspace_params.TimeTransfer = 0.01;
    kspace_params.InputSpectrum_UnitLength_in_cm = UL * 1e3;
    for (int i = 0; i < 3; i++) kspace_params.MNu[i] = mnu[i];
    kspace_params.hybrid_neutrinos_on = hybrid;
    kspace_params.vcrit = 500;
    kspace_params.nu_crit_time = 0.333;
    InitOmegaNu(0.7, 2.7255, MPI_COMM_WORLD);
    static _omega_nu om;
    init_omega_nu(&om, kspace_params.MNu, 0.01, 0.7, 2.7255);
    ksn_ref_set_background(&om, OMEGA0, UT);
    allocate_kspace_memory(N / 2, rank, BOX, UT, UL, OMEGA0, NULL, 1.0, MPI_COMM_WORLD);

    const long long startslab = (long long) rank * (N / R);
    const size_t nel = (size_t) P * N * L;
    fftw_complex *grid = malloc(nel * sizeof(fftw_complex));
    if (!grid) { fprintf(stderr, "rank %d: cannot allocate %zu bytes\n", rank, nel * sizeof(fftw_complex)); ksn_minimpi_exit(1); return 1; }
    const double tg0 = now();
    orc_fill_synthetic_slab((double *) grid, N, startslab, P, SEED, SLOPE);
    const double t_gen = now() - tg0;

    /* first call: delta_tot_init at a = TimeTransfer (untimed) -- and the yardstick for the budget */
    {
        MPI_Barrier(MPI_COMM_WORLD);
        const double t0 = now();
        add_nu_power_to_rhogrid(0.01, BOX, grid, N, (int) startslab, P, MPI_COMM_WORLD);
        const double t_init = now() - t0;
        if (rank == 0) {
            planes[total] = P;
            /* a full-history step adds the integral (~0.3 s on these cores at nk ~ 800) to the grid passes */
            if (budget > 0 && (t_init + 0.4) * total > budget) {
                int p2 = (int) (P * (budget / total - 0.4) / t_init);
                if (p2 < 1) p2 = 1;
                if (p2 < P) planes[total] = p2;
            }
        }
        MPI_Barrier(MPI_COMM_WORLD);
        if (planes[total] < P) {
            P = planes[total];
            delta_tot_table.delta_tot_init_done = 0;        /* initialise again on the slabs the steps will see */
            delta_tot_table.ia = 0;
            add_nu_power_to_rhogrid(0.01, BOX, grid, N, (int) startslab, P, MPI_COMM_WORLD);
        }
    }
    /* long stored history without stepping 100 times: rows at a = 0.01 ... 0.98, delta_tot ~ a */
    {
        const int nk = delta_tot_table.nk, ia = 98;
        double *sf = malloc(sizeof(double) * ia), *dt = malloc(sizeof(double) * (size_t) nk * ia);
        for (int i = 0; i < ia; i++) sf[i] = i == 0 ? delta_tot_table.scalefact[0] : log(0.01 * (i + 1));
        for (int k = 0; k < nk; k++)
            for (int i = 0; i < ia; i++) dt[(size_t) k * ia + i] = delta_tot_table.delta_tot[k][0] * exp(sf[i] - sf[0]);
        set_nu_state(sf, dt, nk, ia, MPI_COMM_WORLD);
        free(sf); free(dt);
    }
    /* every step advances a by less than 0.009 (the row is integrated in full but not kept: steady state), and a stays < 1 */
    const double a0 = 0.98, da = fmin(0.001, (0.9995 - a0) / (total + 1));
    for (int s = 0; s < total; s++) {
        const double a = a0 + da * (s + 1);
        MPI_Barrier(MPI_COMM_WORLD);
        const double t0 = now();
        add_nu_power_to_rhogrid(a, BOX, grid, N, (int) startslab, P, MPI_COMM_WORLD);
        const double t1 = now();
        if (rank == 0) {
            times[4 * s + 0] = t1 - t0;
            times[4 * s + 1] = ksn_ref_t_mass - t0;                    /* K1 loop + all-reduce */
            times[4 * s + 2] = ksn_ref_t_nupower - ksn_ref_t_mass;     /* integral */
            times[4 * s + 3] = t1 - ksn_ref_t_nupower;                 /* scaling loop + barrier */
            planes[s] = P;
        }
        MPI_Barrier(MPI_COMM_WORLD);
    }
    if (rank == 0) {
        printf("{\"N\": %d, \"P\": %d, \"P_full\": %d, \"R\": %d, \"nk\": %d, \"Na\": %d, \"warmup\": %d, \"gen_s\": %.3f, \"same_generator_as_gpu_arm\": true, \"steps\": [",
               N, planes[total - 1], N / R, R, delta_tot_table.nk, delta_tot_table.ia + 1, warmup, t_gen);
        for (int s = warmup; s < total; s++)
            printf("%s{\"total\": %.6f, \"k1\": %.6f, \"integral\": %.6f, \"k3\": %.6f, \"planes\": %d}", s > warmup ? ", " : "",
                   times[4 * s], times[4 * s + 1], times[4 * s + 2], times[4 * s + 3], planes[s]);
        printf("]}\n");
    }
    ksn_minimpi_exit(0);
    return 0;
}
