/* Generic interface + module-global state (replaces interface_common.c).  Everything the host
 * N-body code calls that is independent of the FFT layout: InitOmegaNu, allocate_kspace_memory,
 * OmegaNu, compute_neutrino_power_from_cdm, integrator state get/set/save.
 *
 * Multi-rank note: the reference broadcasts the parameter block, the transfer table and the
 * restart state from rank 0 with MPI_Bcast (interface_common.c:54-75,186).  With KSN_HAVE_MPI
 * the same broadcasts are issued on the host's communicator; without MPI every rank reads the
 * (small) input files itself, which yields the same state on every rank. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "ksn_host.h"

struct __kspace_params kspace_params;
static _transfer_init_table transfer_init;
_delta_tot_table delta_tot_table;
static _omega_nu omeganu_table;
/* three arrays of nk_allocated doubles: delta_cdm | delta_nu | keff  (interface_common.c:24-26,99) */
double *delta_cdm_curr;

_omega_nu *ksn_global_omnu(void) { return &omeganu_table; }
_transfer_init_table *ksn_global_transfer(void) { return &transfer_init; }

double OmegaNu(double a) { return get_omega_nu(&omeganu_table, a); }
double OmegaNu_nopart(double a) { return get_omega_nu_nopart(&omeganu_table, a); }

#ifdef KSN_HAVE_MPI
#include <unistd.h>
static int mpi_allreduce_cb(double *buf, size_t n, void *user)
{
    return MPI_Allreduce(MPI_IN_PLACE, buf, (int) n, MPI_DOUBLE, MPI_SUM, *(MPI_Comm *) user) != MPI_SUCCESS;
}
static MPI_Comm the_comm;

/* collective: did it work on EVERY rank?  (a backend is kept only then -- half the ranks on another one would hang) */
static int all_ok(MPI_Comm comm, int ok)
{
    int bad = !ok, nbad = 0;
    MPI_Allreduce(&bad, &nbad, 1, MPI_INT, MPI_SUM, comm);
    return nbad == 0;
}

/* The ranks of this rank's box, by host name: its index among them, and whether the communicator lives on one box. */
static void box_layout(MPI_Comm comm, int rank, int size, int *local_rank, int *one_box)
{
    char mine[64], *all = malloc((size_t) 64 * size);
    if (!all) terminate(1, "Could not allocate temporary memory for the host names of %d ranks\n", size);
    memset(mine, 0, sizeof mine);
    gethostname(mine, sizeof mine - 1);
    MPI_Allgather(mine, 64, MPI_BYTE, all, 64, MPI_BYTE, comm);
    *local_rank = 0;
    *one_box = 1;
    for (int r = 0; r < size; r++) {
        const int same = !memcmp(all + (size_t) 64 * r, mine, 64);
        if (same && r < rank) (*local_rank)++;
        if (!same) *one_box = 0;
    }
    free(all);
}

/* The one exchange step of the path (powerspectrum.c:91-95) on the communicator the host passes in.  One MPI rank per
 * GPU; the first call picks, collectively, the first of these that works on every rank ($KSN_COMM = p2p | nccl | mpi
 * starts further down the list):
 *   p2p   all ranks on one box and able to map each other's HBM: the sum happens inside the final-reduce kernel over
 *         NVLink peer memory (handles exchanged with MPI_Allgather, a trial sum checked on every rank);
 *   nccl  ncclAllReduce on a communicator bootstrapped with MPI_Bcast of the unique id;
 *   mpi   the sums go to the host and through MPI_Allreduce, as the reference does.
 * A host that has already bound a backend itself (ksn_comm_*) is left alone. */
void ksn_bind_comm(MPI_Comm comm)
{
    int rank, size;
    the_comm = comm;
    MPI_Comm_rank(comm, &rank);
    MPI_Comm_size(comm, &size);
    if (size <= 1 || ksn_comm_size() != 1) return;
    const char *want = getenv("KSN_COMM");
    const int from_p2p = !want || !strcmp(want, "p2p"), from_nccl = from_p2p || !strcmp(want, "nccl");
    int local_rank, one_box;
    box_layout(comm, rank, size, &local_rank, &one_box);
    /* unless the host chose a device (ksn_init, $KSN_DEVICE, $LOCAL_RANK): rank i of a box takes GPU i */
    const int ndev = ksn_device_count();
    if (ndev > 0 && ksn_device() < 0 && !getenv("KSN_DEVICE") && !getenv("LOCAL_RANK")) ksn_init(local_rank % ndev);
    if (from_p2p && one_box && size <= 16) {
        unsigned char mine[64], *all = malloc((size_t) 64 * size);
        int ok = all != NULL && ndev > 0 && ksn_comm_p2p_export(mine) == 0;
        if (all_ok(comm, ok)) {
            MPI_Allgather(mine, 64, MPI_BYTE, all, 64, MPI_BYTE, comm);
            ok = ksn_comm_p2p_init(all, size, rank) == 0;
            if (all_ok(comm, ok)) {                   /* (also the barrier: every mailbox is mapped before a round starts) */
                double v[8];
                for (int i = 0; i < 8; i++) v[i] = (rank + 1) + 100. * i;
                ok = ksn_comm_allreduce_host(v, 8) == 0;
                for (int i = 0; i < 8; i++) ok = ok && v[i] == size * (size + 1) / 2. + 100. * i * size;
                if (all_ok(comm, ok)) {
                    free(all);
                    message(0, "kspace-neutrinos: bin sums are reduced through NVLink peer memory (%d ranks)\n", size);
                    return;
                }
            }
        }
        free(all);
        ksn_comm_single();                            /* every rank forgets the half-built backend */
    }
    if (from_nccl) {
        unsigned char id[128], probe[128];
        /* every rank: is there a device, can libnccl be loaded here?  (a rank that failed later would leave the others
         * waiting inside ncclCommInitRank) */
        int ok = ndev > 0 && ksn_comm_nccl_unique_id(rank == 0 ? id : probe) == 0;
        if (all_ok(comm, ok)) {
            MPI_Bcast(id, 128, MPI_BYTE, 0, comm);
            ok = ksn_comm_nccl_init(id, size, rank) == 0;
            if (all_ok(comm, ok)) {
                message(0, "kspace-neutrinos: bin sums are reduced with NCCL (%d ranks)\n", size);
                return;
            }
            ksn_comm_single();
        }
    }
    ksn_comm_host_callback(mpi_allreduce_cb, &the_comm, size, rank);
    message(0, "kspace-neutrinos: bin sums are reduced with MPI_Allreduce on the host (%d ranks)\n", size);
}
#else
void ksn_bind_comm(MPI_Comm comm) { (void) comm; }
#endif

void InitOmegaNu(const double HubbleParam, const double tcmb0, MPI_Comm MYMPI_COMM_WORLD)
{
    ksn_bind_comm(MYMPI_COMM_WORLD);
#ifdef KSN_HAVE_MPI
    MPI_Bcast(&kspace_params, sizeof(kspace_params), MPI_BYTE, 0, MYMPI_COMM_WORLD);
#endif
    init_omega_nu(&omeganu_table, kspace_params.MNu, kspace_params.TimeTransfer, HubbleParam, tcmb0);
}

void allocate_kspace_memory(const int nk_in, const int ThisTask, const double BoxSize, const double UnitTime_in_s, const double UnitLength_in_cm, const double Omega0, char *snapdir, const double TimeMax, MPI_Comm MYMPI_COMM_WORLD)
{
    ksn_bind_comm(MYMPI_COMM_WORLD);
    /* vcrit is in km/s, so the speed of light goes in km/s as well */
    if (kspace_params.hybrid_neutrinos_on)
        init_hybrid_nu(&omeganu_table.hybnu, kspace_params.MNu, kspace_params.vcrit, LIGHTCGS / 1e5, kspace_params.nu_crit_time, omeganu_table.kBtnu);
#ifdef KSN_HAVE_MPI
    if (ThisTask == 0)
        allocate_transfer_init_table(&transfer_init, BoxSize, UnitLength_in_cm, kspace_params.InputSpectrum_UnitLength_in_cm, kspace_params.KspaceTransferFunction);
    MPI_Bcast(&transfer_init.NPowerTable, 1, MPI_INT, 0, MYMPI_COMM_WORLD);
    if (ThisTask != 0) transfer_init.logk = (double *) mymalloc("Transfer_functions", 2 * transfer_init.NPowerTable * sizeof(double));
    transfer_init.T_nu = transfer_init.logk + transfer_init.NPowerTable;
    MPI_Bcast(transfer_init.logk, 2 * transfer_init.NPowerTable, MPI_DOUBLE, 0, MYMPI_COMM_WORLD);
#else
    allocate_transfer_init_table(&transfer_init, BoxSize, UnitLength_in_cm, kspace_params.InputSpectrum_UnitLength_in_cm, kspace_params.KspaceTransferFunction);
#endif
    delta_tot_table.ThisTask = ThisTask;
    allocate_delta_tot_table(&delta_tot_table, nk_in, kspace_params.TimeTransfer, TimeMax, Omega0, &omeganu_table, UnitTime_in_s, UnitLength_in_cm, 0);
#ifdef KSN_HAVE_MPI
    if (ThisTask == 0 && snapdir != NULL) read_all_nu_state(&delta_tot_table, snapdir);
    MPI_Bcast(&delta_tot_table.ia, 1, MPI_INT, 0, MYMPI_COMM_WORLD);
    if (delta_tot_table.ia > 0) {
        MPI_Bcast(&delta_tot_table.nk, 1, MPI_INT, 0, MYMPI_COMM_WORLD);
        MPI_Bcast(delta_tot_table.scalefact, delta_tot_table.namax * (nk_in + 1), MPI_DOUBLE, 0, MYMPI_COMM_WORLD);
    }
#else
    if (snapdir != NULL) read_all_nu_state(&delta_tot_table, snapdir);
#endif
    delta_cdm_curr = mymalloc("temp_power_spectrum", 3 * nk_in * sizeof(double));
    if (!delta_cdm_curr) terminate(2018, "Could not allocate temporary memory for power spectra\n");
    ksn_invalidate_background();
}

void save_nu_state(char *savefile)
{
    if (delta_tot_table.ThisTask == 0) save_all_nu_state(&delta_tot_table, savefile);
}

int save_neutrino_power(const double Time, const int snapnum, const char *OutputDir)
{
    if (delta_tot_table.ThisTask != 0) return 0;
    return save_nu_power(&delta_tot_table, Time, snapnum, OutputDir);
}

_delta_pow compute_neutrino_power_internal(const double Time, double *keff, double *delta_cdm, double *delta_nu, const int nk_nonzero)
{
    _delta_pow d_pow;
    get_delta_nu_update(&delta_tot_table, Time, nk_nonzero, keff, delta_cdm, delta_nu, &transfer_init);
    message(0, "Done getting neutrino power: nk= %d, k = %g, delta_nu = %g, delta_cdm = %g,\n", nk_nonzero, keff[1], delta_nu[1], delta_cdm[1]);
    /* the table interpolates delta_nu/delta_cdm in log k; both conversions are in place.  keff is the mean |k| of the
     * modes of a bin: the same numbers step after step while the slab geometry stands, so the logarithms of the last call
     * are kept and reused when the input compares equal (nk libm calls less between K2 and K3 of every step) */
    {
        static double *seen = NULL, *logs = NULL;
        static int cap = 0, n_seen = 0;
        if (n_seen == nk_nonzero && nk_nonzero > 0 && memcmp(seen, keff, sizeof(double) * nk_nonzero) == 0) {
            memcpy(keff, logs, sizeof(double) * nk_nonzero);
        } else {
            if (cap < nk_nonzero) {
                free(seen); free(logs);
                seen = malloc(sizeof(double) * nk_nonzero);
                logs = malloc(sizeof(double) * nk_nonzero);
                cap = (seen && logs) ? nk_nonzero : 0;
            }
            n_seen = 0;
            if (cap >= nk_nonzero) memcpy(seen, keff, sizeof(double) * nk_nonzero);
            for (int i = 0; i < nk_nonzero; i++) keff[i] = log(keff[i]);
            if (cap >= nk_nonzero) { memcpy(logs, keff, sizeof(double) * nk_nonzero); n_seen = nk_nonzero; }
        }
    }
    for (int i = 0; i < nk_nonzero; i++) delta_cdm[i] = delta_nu[i] / delta_cdm[i];
    /* analytic neutrino mass over mass carried by particles (hybrid particles included) */
    const double OmegaNu_nop = get_omega_nu_nopart(&omeganu_table, Time);
    const double omega_hybrid = get_omega_nu(&omeganu_table, Time) - OmegaNu_nop;
    const double kspace_prefac = OmegaNu_nop / (delta_tot_table.Omeganonu / pow(Time, 3) + omega_hybrid);
    init_delta_pow(&d_pow, keff, delta_cdm, nk_nonzero, kspace_prefac);
    return d_pow;
}

_delta_pow compute_neutrino_power_from_cdm(const double Time, const double keff_in[], const double P_cdm[], const long int Nmodes[], const int nk_in, MPI_Comm MYMPI_COMM_WORLD)
{
    (void) MYMPI_COMM_WORLD;
    double *delta_nu = delta_cdm_curr + nk_in;
    double *keff = delta_cdm_curr + 2 * nk_in;
    int kept = 0;
    for (int i = 0; i < nk_in; i++) {
        if (Nmodes[i] == 0) continue;
        delta_cdm_curr[kept] = sqrt(P_cdm[i]);
        keff[kept] = keff_in[i];
        kept++;
    }
    return compute_neutrino_power_internal(Time, keff, delta_cdm_curr, delta_nu, kept);
}

void get_nu_state(double **scalefact, double **delta_tot, size_t *nk, size_t *ia)
{
    *nk = delta_tot_table.nk;
    *ia = delta_tot_table.ia;
    *scalefact = mymalloc("tmp_scales", (*ia) * sizeof(double));
    *delta_tot = mymalloc("tmp_delta", (*nk) * (*ia) * sizeof(double));
    for (size_t i = 0; i < *ia; i++) (*scalefact)[i] = delta_tot_table.scalefact[i];
    for (size_t k = 0; k < *nk; k++)
        for (size_t i = 0; i < *ia; i++) (*delta_tot)[k * (*ia) + i] = delta_tot_table.delta_tot[k][i];
}

void set_nu_state(double *scalefact, double *delta_tot, const size_t nk, const size_t ia, MPI_Comm MYMPI_COMM_WORLD)
{
    (void) MYMPI_COMM_WORLD;
    delta_tot_table.nk = nk;
    delta_tot_table.ia = ia;
    for (size_t i = 0; i < ia; i++) delta_tot_table.scalefact[i] = scalefact[i];
    for (size_t k = 0; k < nk; k++)
        for (size_t i = 0; i < ia; i++) delta_tot_table.delta_tot[k][i] = delta_tot[k * ia + i];
#ifdef KSN_HAVE_MPI
    MPI_Bcast(&delta_tot_table.ia, 1, MPI_INT, 0, MYMPI_COMM_WORLD);
    if (delta_tot_table.ia > 0) {
        MPI_Bcast(&delta_tot_table.nk, 1, MPI_INT, 0, MYMPI_COMM_WORLD);
        MPI_Bcast(delta_tot_table.scalefact, delta_tot_table.namax * (delta_tot_table.nk_allocated + 1), MPI_DOUBLE, 0, MYMPI_COMM_WORLD);
    }
#endif
}

int particle_nu_active(double a)
{
    return particle_nu_fraction(&omeganu_table.hybnu, a, 0) != 0.;
}
