/* TEST PROGRAM (CPU, no GPU needed).  A host on the KSPACE_NEUTRINOS_2-off path of the Gadget-2 patches (0004:
 * pmforce_periodic -> compute_total_power_spectrum(..., MPI_COMM_WORLD) and nothing else) never calls InitOmegaNu or
 * allocate_kspace_memory, so the communicator reaches the library only as the last argument of the call itself
 * (interface_gadget.c:114, powerspectrum.c:33,91-95).  Built with -DKSN_HAVE_MPI over the mini-MPI and the test-only CPU
 * stand-in for the device entry points: R ranks call total_powerspectrum on their x-slabs; every rank must get the
 * GLOBAL spectrum (same counts, finite power on the ranks that do not hold plane 0).
 *   usage: mpi_total_power <nranks> <out>                                                                          */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mpi.h>
#include "ksn_host.h"

int ksn_minimpi_fork(int nranks);
void ksn_minimpi_exit(int code);
void *ksn_minimpi_shared_alloc(size_t bytes);

#define N 20
#define NB (N / 2)

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s nranks out\n", argv[0]); return 2; }
    const int R = atoi(argv[1]);
    const size_t plane = (size_t) N * (N / 2 + 1), total = N * plane;
    fftw_complex *grid = ksn_minimpi_shared_alloc(total * sizeof(fftw_complex));
    double *all = ksn_minimpi_shared_alloc(sizeof(double) * 16 * (3 * NB + 1));
    unsigned long long s = 1442695040888963407ull;
    for (size_t i = 0; i < total; i++) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        grid[i].re = (double) (s >> 11) / 9007199254740992.0 - 0.5;
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        grid[i].im = (double) (s >> 11) / 9007199254740992.0 - 0.5;
    }
    grid[0].re = (double) N * N * N; grid[0].im = 0;
    const int rank = ksn_minimpi_fork(R);
    ksn_set_quiet(1);
    const int start = (int) ((long long) rank * N / R), end = rank == R - 1 ? N : (int) ((long long) (rank + 1) * N / R);
    double power[NB], keffs[NB];
    long long count[NB];
    const int nret = total_powerspectrum_f64(N, grid + start * plane, NB, start, end - start, power, count, keffs, MPI_COMM_WORLD);
    double *mine = all + (size_t) rank * (3 * NB + 1);
    mine[0] = nret;
    for (int i = 0; i < NB; i++) { mine[1 + i] = i < nret ? power[i] : 0; mine[1 + NB + i] = i < nret ? keffs[i] : 0; mine[1 + 2 * NB + i] = i < nret ? (double) count[i] : 0; }
    /* the no-neutrino path of the PM hook, twice (the second call reuses the cached geometry) */
    compute_total_power_spectrum_f64(0.5, 512000., grid + start * plane, N, start, end - start, MPI_COMM_WORLD);
    compute_total_power_spectrum_f64(0.5, 512000., grid + start * plane, N, start, end - start, MPI_COMM_WORLD);
    MPI_Barrier(MPI_COMM_WORLD);
    int bad = 0;
    for (int r = 1; r < R; r++)
        if (memcmp(all, all + (size_t) r * (3 * NB + 1), sizeof(double) * (3 * NB + 1))) bad = 1;
    for (int i = 0; i < nret; i++) if (!isfinite(power[i]) || !(power[i] > 0)) bad = 1;
    if (rank == 0) {
        FILE *f = fopen(argv[2], "wb");
        if (!f) { perror(argv[2]); bad = 1; }
        else { fwrite(all, sizeof(double), 3 * NB + 1, f); fclose(f); }
        printf(bad ? "MPI TOTAL POWER FAILED\n" : "MPI TOTAL POWER OK\n");
    }
    ksn_minimpi_exit(bad);
    return bad;
}
