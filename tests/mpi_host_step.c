/* TEST PROGRAM (CPU, no GPU needed).  The product's host C layer built with -DKSN_HAVE_MPI and linked with the test-only
 * CPU stand-in for the device entry points (tests/device_standin.c) takes a few PM steps on a small grid split into
 * x-slabs over R ranks of the oracle's fork-based mini-MPI (oracle/mini_mpi.c): every rank calls
 * add_nu_power_to_rhogrid(Time, BoxSize, its slab, pmgrid, slabstart, nslab, comm) as a Gadget PM routine would
 * (interface_gadget.h:38), the bin sums go through the communicator bound at InitOmegaNu.  The corrected slabs and
 * delta_nu_last are written to <out>; the caller compares R = 1 with R = 2, 3 (uneven split).
 *   usage: mpi_host_step <transfer_file> <nranks> <out>                                                    */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mpi.h>
#include "ksn_host.h"

int ksn_minimpi_fork(int nranks);
void ksn_minimpi_exit(int code);
void *ksn_minimpi_shared_alloc(size_t bytes);

#define N 24

/* test-only stand-in (tests/device_standin.c) reports which backend the bootstrap chose; absent when the program is linked
 * to the real library (the multi-GPU run), where ksn_comm_size() has to do */
int ksn_standin_backend(void) __attribute__((weak));

int main(int argc, char **argv)
{
    if (argc < 4) { fprintf(stderr, "usage: %s transfer_file nranks out\n", argv[0]); return 2; }
    const int R = atoi(argv[2]);
    const size_t plane = (size_t) N * (N / 2 + 1), total = N * plane;
    /* the whole grid lives in memory every rank sees, each rank touches only its slab */
    fftw_complex *grid = ksn_minimpi_shared_alloc(total * sizeof(fftw_complex));
    double *verdict = ksn_minimpi_shared_alloc(sizeof(double) * 64);
    unsigned long long s = 88172645463325252ull;
    for (size_t i = 0; i < total; i++) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;                                   /* xorshift64 */
        const double k = 1 + (double) ((i / plane) % (N / 2)) + (double) (i % (N / 2 + 1));
        grid[i].re = ((double) (s >> 11) / 9007199254740992.0 - 0.5) / k;
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        grid[i].im = ((double) (s >> 11) / 9007199254740992.0 - 0.5) / k;
    }
    grid[0].re = (double) N * N * N; grid[0].im = 0;
    const int rank = ksn_minimpi_fork(R);
    ksn_set_quiet(1);
    memset(&kspace_params, 0, sizeof kspace_params);
    if (rank == 0) {
        strncpy(kspace_params.KspaceTransferFunction, argv[1], sizeof kspace_params.KspaceTransferFunction - 1);
        kspace_params.TimeTransfer = 0.01;
        kspace_params.InputSpectrum_UnitLength_in_cm = 3.085678e24;
        kspace_params.MNu[0] = 0.2; kspace_params.MNu[1] = 0.1; kspace_params.MNu[2] = 0.3;
    }
    const double UnitLength = 3.085678e21, Box = 512000.;
    InitOmegaNu(0.7, 2.7255, MPI_COMM_WORLD);
    allocate_kspace_memory(N / 2, rank, Box, UnitLength / 1e5, UnitLength, 0.2793, NULL, 1.0, MPI_COMM_WORLD);
    ksn_set_default_hubble(NULL, 0.2793, UnitLength / 1e5);
    /* uneven split: rank r owns [r N / R, (r + 1) N / R) rounded down, the last rank the rest */
    const int start = (int) ((long long) rank * N / R), end = rank == R - 1 ? N : (int) ((long long) (rank + 1) * N / R);
    const double times[] = { 0.01, 0.02, 0.0205, 0.05, 0.2 };
    for (size_t t = 0; t < sizeof times / sizeof times[0]; t++)
        add_nu_power_to_rhogrid_f64(times[t], Box, grid + start * plane, N, start, end - start, MPI_COMM_WORLD);
    /* every rank must hold the same integrator state */
    double chk = 0;
    for (int k = 0; k < delta_tot_table.nk; k++) chk += delta_tot_table.delta_nu_last[k] * (k + 1);
    verdict[rank] = chk;
    MPI_Barrier(MPI_COMM_WORLD);
    int bad = 0;
    for (int r = 0; r < R; r++) if (verdict[r] != verdict[0]) bad = 1;
    /* which collective the bootstrap (iface_common.c: bind_comm) settled on -- it must be the same one on every rank */
    verdict[32 + rank] = ksn_standin_backend ? ksn_standin_backend() : (ksn_comm_size() > 1 ? 100 + ksn_comm_size() : 0);
    MPI_Barrier(MPI_COMM_WORLD);
    for (int r = 0; r < R; r++) if (verdict[32 + r] != verdict[32]) bad = 1;
    if (rank == 0) printf("BACKEND %d\n", (int) verdict[32]);
    if (rank == 0) {
        FILE *f = fopen(argv[3], "wb");
        if (!f) { perror(argv[3]); bad = 1; }
        else {
            const int hdr[3] = { N, delta_tot_table.nk, delta_tot_table.ia };
            fwrite(hdr, sizeof hdr, 1, f);
            fwrite(delta_tot_table.delta_nu_last, sizeof(double), delta_tot_table.nk, f);
            fwrite(grid, sizeof(fftw_complex), total, f);
            fclose(f);
        }
        printf(bad ? "MPI HOST STEP FAILED\n" : "MPI HOST STEP OK\n");
    }
    ksn_minimpi_exit(bad);
    return bad;
}
