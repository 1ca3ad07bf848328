/* TEST INFRASTRUCTURE: the REFERENCE's own total_powerspectrum (powerspectrum.c:33-117) and add_nu_power_to_rhogrid
 * (interface_gadget.c:158-194) on a handful of planes of a FULL-WIDTH grid (PMGRID 1024 / 2048 / 4096), so that the kernels
 * the benchmark times can be compared with the reference on the same bytes in seconds of CPU time.
 *
 * Linked against the reference sources compiled unmodified from /root/reference plus the shims (oracle/Makefile target
 * _ref/ref_slabs).  R forked mini-MPI ranks each own ONE contiguous range of planes; the ranges need not tile the grid
 * (the reference never checks that they do: every rank just passes its startslab/nslab, and the bin sums are all-reduced),
 * so rank 0 can hold planes [0, p) -- it must hold plane 0, whose first element is the total mass -- while another rank
 * holds planes around the Nyquist index.
 *
 *   ref_slabs N hybrid m0 m1 m2 TRANSFER IN OUT  R start_0 n_0 ... start_{R-1} n_{R-1}  T a_0 ... a_{T-1}
 *
 * IN:  the slabs in rank order, raw fftw_complex (double; float in the build without -DDOUBLEPRECISION_FFTW,
 *      _ref/ref_slabs_single), n_r * N * (N/2+1) elements each.
 * OUT: (rank 0) int32 {N, nret, nk, ia, T}; double power[nret], keffs[nret], count[nret] of total_powerspectrum on the
 *      input (nrbins = N/2, as compute_neutrino_power_spectrum passes it); then per time step delta_nu_last[nk_t] preceded
 *      by a double nk_t; OUT.grid: the slabs after the T calls of add_nu_power_to_rhogrid, rank order.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "shim/mpi.h"
#include "interface_gadget.h"
#include "delta_tot_table.h"
#include "omega_nu_single.h"
#include "powerspectrum.h"

extern int ThisTask;
extern int ksn_ref_quiet;
extern _delta_tot_table delta_tot_table;
void ksn_ref_set_background(const _omega_nu *omnu, double Omega0, double UnitTime_in_s);
int ksn_minimpi_fork(int nranks);
void ksn_minimpi_exit(int code);
void *ksn_minimpi_shared_alloc(size_t bytes);

int main(int argc, char **argv)
{
    if (argc < 13) { fprintf(stderr, "usage: %s N hybrid m0 m1 m2 TRANSFER IN OUT R start n ... T a ...\n", argv[0]); return 2; }
    const int N = atoi(argv[1]), hybrid = atoi(argv[2]);
    const double m[3] = { atof(argv[3]), atof(argv[4]), atof(argv[5]) };
    const char *transfer = argv[6], *in = argv[7], *out = argv[8];
    const int R = atoi(argv[9]);
    if (R < 1 || R > 16 || argc < 10 + 2 * R + 2) { fprintf(stderr, "bad rank list\n"); return 2; }
    long long start[16], cnt[16], off[17];
    off[0] = 0;
    const size_t plane = (size_t) N * (N / 2 + 1);
    for (int r = 0; r < R; r++) {
        start[r] = atoll(argv[10 + 2 * r]);
        cnt[r] = atoll(argv[11 + 2 * r]);
        off[r + 1] = off[r] + cnt[r];
    }
    const int T = atoi(argv[10 + 2 * R]);
    if (T < 1 || argc < 11 + 2 * R + T) { fprintf(stderr, "bad time list\n"); return 2; }
    const double UL = 3.085678e21, UT = UL / 1e5, BOX = 512000, OMEGA0 = 0.2793;
    const int rank = ksn_minimpi_fork(R);
    ThisTask = rank;
    ksn_ref_quiet = 1;
    strncpy(kspace_params.KspaceTransferFunction, transfer, 499);
    kspace_params.TimeTransfer = 0.01;
    kspace_params.InputSpectrum_UnitLength_in_cm = UL * 1e3;
    for (int i = 0; i < 3; i++) kspace_params.MNu[i] = m[i];
    kspace_params.hybrid_neutrinos_on = hybrid;
    kspace_params.vcrit = 500;
    kspace_params.nu_crit_time = 0.333;
    InitOmegaNu(0.7, 2.7255, MPI_COMM_WORLD);
    static _omega_nu om;
    init_omega_nu(&om, kspace_params.MNu, 0.01, 0.7, 2.7255);
    ksn_ref_set_background(&om, OMEGA0, UT);
    allocate_kspace_memory(N / 2, rank, BOX, UT, UL, OMEGA0, NULL, 1.0, MPI_COMM_WORLD);

    const size_t nel = (size_t) cnt[rank] * plane;
    fftw_complex *grid = malloc((nel ? nel : 1) * sizeof(fftw_complex));
    FILE *f = fopen(in, "rb");
    if (!grid || !f) { fprintf(stderr, "rank %d: cannot read %s\n", rank, in); ksn_minimpi_exit(1); return 1; }
    fseeko(f, (off_t) ((size_t) off[rank] * plane * sizeof(fftw_complex)), SEEK_SET);
    if (fread(grid, sizeof(fftw_complex), nel, f) != nel) { fprintf(stderr, "rank %d: short read\n", rank); ksn_minimpi_exit(1); return 1; }
    fclose(f);

    const int nb = N / 2;
    double *power = malloc(sizeof(double) * nb), *keffs = malloc(sizeof(double) * nb);
    long long *count = malloc(sizeof(long long) * nb);
    const int nret = total_powerspectrum(N, grid, nb, (int) start[rank], (int) cnt[rank], power, count, keffs, MPI_COMM_WORLD);
    FILE *fo = NULL;
    if (rank == 0) {
        fo = fopen(out, "wb");
        if (!fo) { perror(out); ksn_minimpi_exit(1); return 1; }
        const int hdr[5] = { N, nret, 0, 0, T };
        fwrite(hdr, sizeof hdr, 1, fo);
        fwrite(power, sizeof(double), nret, fo);
        fwrite(keffs, sizeof(double), nret, fo);
        for (int i = 0; i < nret; i++) { const double c = (double) count[i]; fwrite(&c, sizeof c, 1, fo); }
    }
    for (int t = 0; t < T; t++) {
        const double a = atof(argv[11 + 2 * R + t]);
        add_nu_power_to_rhogrid(a, BOX, grid, N, (int) start[rank], (int) cnt[rank], MPI_COMM_WORLD);
        if (rank == 0) {
            const double nk = delta_tot_table.nk;
            fwrite(&nk, sizeof nk, 1, fo);
            fwrite(delta_tot_table.delta_nu_last, sizeof(double), delta_tot_table.nk, fo);
        }
    }
    if (rank == 0) {
        const int tail[2] = { delta_tot_table.nk, delta_tot_table.ia };
        fseek(fo, 2 * sizeof(int), SEEK_SET);
        fwrite(tail, sizeof tail, 1, fo);
        fclose(fo);
    }
    /* every rank writes its slab at its own offset of OUT.grid (rank 0 creates the file first) */
    char gname[1024];
    snprintf(gname, sizeof gname, "%s.grid", out);
    if (rank == 0) { FILE *g = fopen(gname, "wb"); if (g) fclose(g); }
    MPI_Barrier(MPI_COMM_WORLD);
    FILE *g = fopen(gname, "r+b");
    if (!g) { perror(gname); ksn_minimpi_exit(1); return 1; }
    fseeko(g, (off_t) ((size_t) off[rank] * plane * sizeof(fftw_complex)), SEEK_SET);
    fwrite(grid, sizeof(fftw_complex), nel, g);
    fclose(g);
    MPI_Barrier(MPI_COMM_WORLD);
    ksn_minimpi_exit(0);
    return 0;
}
