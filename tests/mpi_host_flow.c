/* TEST PROGRAM (CPU, no GPU needed).  The product's host C layer built with -DKSN_HAVE_MPI -- the build an MPI PM code
 * uses -- run as TWO ranks over the oracle's fork-based mini-MPI (oracle/mini_mpi.c, test infrastructure):
 *   - only rank 0 knows the parameter block and can read the transfer file; after InitOmegaNu / allocate_kspace_memory
 *     (interface_common.c:54-75 broadcasts) rank 1 must hold the same parameters and the same transfer table;
 *   - the communicator passed in must have been bound as the library's collective (host call-back -> MPI_Allreduce),
 *     so that ksn_comm_allreduce_host sums over the ranks.
 * Exit code 0 on every rank = pass. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mpi.h>
#include "ksn_host.h"

int ksn_minimpi_fork(int nranks);
void ksn_minimpi_exit(int code);

#define CHECK(cond) do { if (!(cond)) { fprintf(stderr, "rank %d: check failed: %s (line %d)\n", rank, #cond, __LINE__); bad = 1; } } while (0)

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s transfer_file\n", argv[0]); return 2; }
    const int rank = ksn_minimpi_fork(2);
    int bad = 0;
    ksn_set_quiet(1);
    memset(&kspace_params, 0, sizeof kspace_params);
    if (rank == 0) {                                   /* the other rank gets everything by broadcast */
        strncpy(kspace_params.KspaceTransferFunction, argv[1], sizeof kspace_params.KspaceTransferFunction - 1);
        kspace_params.TimeTransfer = 0.01;
        kspace_params.InputSpectrum_UnitLength_in_cm = 3.085678e24;
        kspace_params.MNu[0] = kspace_params.MNu[1] = kspace_params.MNu[2] = 0.15;
    } else {
        strcpy(kspace_params.KspaceTransferFunction, "/nonexistent/only-rank-0-reads-the-file");
    }
    InitOmegaNu(0.7, 2.7255, MPI_COMM_WORLD);
    CHECK(kspace_params.TimeTransfer == 0.01 && kspace_params.MNu[2] == 0.15);
    CHECK(strcmp(kspace_params.KspaceTransferFunction, argv[1]) == 0);
    CHECK(ksn_comm_size() == 2 && ksn_comm_rank() == rank);
    const double UnitLength = 3.085678e21;
    allocate_kspace_memory(32, rank, 512000., UnitLength / 1e5, UnitLength, 0.2793, NULL, 1.0, MPI_COMM_WORLD);
    _transfer_init_table *t = ksn_global_transfer();
    CHECK(t->NPowerTable == 271);                      /* transfer_init_test.c: the fixture's table length */
    /* identical tables on both ranks: sum of (rank ? -x : x) over ranks must vanish element by element */
    double *diff = malloc(sizeof(double) * 2 * t->NPowerTable);
    for (int i = 0; i < 2 * t->NPowerTable; i++) diff[i] = (rank ? -1.0 : 1.0) * t->logk[i];
    CHECK(ksn_comm_allreduce_host(diff, 2 * (size_t) t->NPowerTable) == 0);
    for (int i = 0; i < 2 * t->NPowerTable; i++) if (diff[i] != 0.0) { CHECK(diff[i] == 0.0); break; }
    CHECK(fabs(t->T_nu[0] - 0.5084792) < 1e-6);        /* known value of the fixture */
    /* the collective really sums over ranks */
    double v[3] = { 1.0 + rank, 10.0 * (1 + rank), -2.5 };
    CHECK(ksn_comm_allreduce_host(v, 3) == 0);
    CHECK(v[0] == 3.0 && v[1] == 30.0 && v[2] == -5.0);
    CHECK(delta_tot_table.ThisTask == rank && delta_tot_table.nk_allocated == 32);
    free(diff);
    /* fold rank 1's verdict into rank 0's exit code */
    double verdict = bad;
    ksn_comm_allreduce_host(&verdict, 1);
    if (rank == 0) printf(verdict == 0.0 ? "MPI HOST FLOW OK\n" : "MPI HOST FLOW FAILED\n");
    ksn_minimpi_exit(verdict != 0.0);
    return verdict != 0.0;
}
