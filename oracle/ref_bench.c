/* TEST/BENCH INFRASTRUCTURE: times the REFERENCE's own CPU implementation of the per-PM-step hot
 * path (add_nu_power_to_rhogrid, interface_gadget.c:158-194) on this host's cores.
 *
 * Linked against the reference sources compiled unmodified from /root/reference plus the shims
 * (oracle/Makefile target _ref/ref_bench).  R forked ranks (mini-MPI) each own a slab that starts
 * where their slab of the full PMGRID^3 grid would start, but only P planes deep (bounded sample:
 * both grid passes are linear in the number of modes); the linear-response integral is replicated on
 * every rank exactly as in the reference and is timed in full, at a ~100-row stored history.
 * Phase boundaries come from the reference's own progress messages (ref_host.c records their time).
 *
 * usage: ref_bench N P R STEPS HYBRID TRANSFER_FILE
 * prints one JSON line on rank 0.
 */
#define _GNU_SOURCE
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

static unsigned long long mix(unsigned long long x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

int main(int argc, char **argv)
{
    if (argc < 7) { fprintf(stderr, "usage: %s N P R STEPS HYBRID TRANSFER_FILE\n", argv[0]); return 2; }
    const int N = atoi(argv[1]), P = atoi(argv[2]), R = atoi(argv[3]), steps = atoi(argv[4]), hybrid = atoi(argv[5]);
    const char *transfer = argv[6];
    const double UL = 3.085678e21, UT = UL / 1e5, BOX = 512000, OMEGA0 = 0.2793;
    const int L = N / 2 + 1;
    double *times = ksn_minimpi_shared_alloc(sizeof(double) * 4 * (steps + 1));
    const int rank = ksn_minimpi_fork(R);
    ThisTask = rank;
    ksn_ref_quiet = 2;                  /* record time stamps, print nothing */
    strncpy(kspace_params.KspaceTransferFunction, transfer, 499);
    kspace_params.TimeTransfer = 0.01;
    kspace_params.InputSpectrum_UnitLength_in_cm = UL * 1e3;
    kspace_params.MNu[0] = kspace_params.MNu[1] = kspace_params.MNu[2] = 0.1;
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
    for (size_t e = 0; e < nel; e++) {
        const long long row = e / L, i = startslab + row / N;
        const int z = (int) (e - row * L), j = (int) (row % N);
        const double ki = i <= N / 2 ? i : i - N, kj = j <= N / 2 ? j : j - N;
        const double k2 = ki * ki + kj * kj + (double) z * z;
        const unsigned long long h = mix(20261017ull ^ mix((unsigned long long) ((i * N + j) * L + z)));
        const double amp = k2 > 0 ? pow(k2, -0.25) : 0;       /* P(k) ~ 1/k */
        grid[e].re = amp * (((h >> 11) & 0xfffff) / 524288.0 - 1.0);
        grid[e].im = amp * (((h >> 31) & 0xfffff) / 524288.0 - 1.0);
    }
    if (rank == 0) { grid[0].re = (double) N * N * N; grid[0].im = 0; }

    /* first call: delta_tot_init at a = TimeTransfer (untimed) */
    add_nu_power_to_rhogrid(0.01, BOX, grid, N, (int) startslab, P, MPI_COMM_WORLD);
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
    for (int s = 0; s <= steps; s++) {          /* s = 0 is a warm-up */
        const double a = 0.98 + 0.001 * (s + 1);
        MPI_Barrier(MPI_COMM_WORLD);
        const double t0 = now();
        add_nu_power_to_rhogrid(a, BOX, grid, N, (int) startslab, P, MPI_COMM_WORLD);
        const double t1 = now();
        if (rank == 0) {
            times[4 * s + 0] = t1 - t0;
            times[4 * s + 1] = ksn_ref_t_mass - t0;                    /* K1 loop + all-reduce */
            times[4 * s + 2] = ksn_ref_t_nupower - ksn_ref_t_mass;     /* integral */
            times[4 * s + 3] = t1 - ksn_ref_t_nupower;                 /* scaling loop + barrier */
        }
    }
    if (rank == 0) {
        printf("{\"N\": %d, \"P\": %d, \"R\": %d, \"nk\": %d, \"Na\": %d, \"steps\": [", N, P, R, delta_tot_table.nk, delta_tot_table.ia + 1);
        for (int s = 1; s <= steps; s++)
            printf("%s{\"total\": %.6f, \"k1\": %.6f, \"integral\": %.6f, \"k3\": %.6f}", s > 1 ? ", " : "",
                   times[4 * s], times[4 * s + 1], times[4 * s + 2], times[4 * s + 3]);
        printf("]}\n");
    }
    ksn_minimpi_exit(0);
    return 0;
}
