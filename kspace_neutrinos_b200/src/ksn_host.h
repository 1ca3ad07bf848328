/* Internal host-side helpers shared by the C translation units (not part of the public API). */
#ifndef KSN_HOST_H
#define KSN_HOST_H
#define KSN_NO_TYPED_MACROS
#include "kspace_neutrinos.h"
#include "ksn_b200.h"
#include "ksn_numeric.h"

/* module-global state (the reference keeps the same objects in interface_common.c:17-26) */
extern _delta_tot_table delta_tot_table;
extern double *delta_cdm_curr;
_omega_nu *ksn_global_omnu(void);
_transfer_init_table *ksn_global_transfer(void);

/* weak host fallbacks may be configured for stand-alone use (tests, bench) */
void ksn_set_default_hubble(const _omega_nu *omnu, double Omega0, double UnitTime_in_s);
void ksn_set_quiet(int quiet);

/* map a device-layer failure onto the reference's error convention: terminate(code, ...) */
void ksn_fatal_device(int rc, const char *where);

/* K1 geometry tables for (dims, nrbins): cached */
int ksn_bin_tables(int dims, int nrbins, const unsigned int **thresholds, const double **invwin);

/* make sure the device table of 1/(aH) covers [a_lo, a_hi] and matches the current hubble_function */
void ksn_ensure_background(double a_lo, double a_hi);
void ksn_invalidate_background(void);
/* start K2's a-only tables for the step at scale factor a on a side stream (beside K1) */
void ksn_prefetch_delta_nu(const _delta_tot_table *const d_tot, const double a);

/* The collective for the bin sums on the communicator a reference-named entry was handed (-DKSN_HAVE_MPI: the first call
 * with more than one rank picks the backend collectively, later calls return at once; without MPI the handle is an int and
 * the host binds a backend through ksn_comm_* itself).  Every entry that ends in the reference's MPI_Allreduce calls
 * (powerspectrum.c:91-95) calls it, so that a host which never goes through InitOmegaNu -- gadget-2 patch 0004 with
 * KSPACE_NEUTRINOS_2 off: pmforce_periodic -> compute_total_power_spectrum -- still gets global sums. */
void ksn_bind_comm(MPI_Comm comm);

/* glue used by both interface files */
_delta_pow compute_neutrino_power_internal(const double Time, double *keff, double *delta_cdm_curr, double *delta_nu_curr, const int nk_nonzero);
int ksn_finish_powerspectrum(int nrbins, double total_mass2, double *power, long long *count, double *keffs);
#endif
