/* Weak fallbacks for the symbols a host N-body code provides (gadget_defines.h:11,24-32), so
 * that the shared library also loads stand-alone (ctypes, bench, tests).  A host that defines
 * hubble_function/terminate/message/mymalloc_fullinfo/myfree_fullinfo overrides these at link
 * time.  The fallback Hubble rate is the flat LCDM + massive-neutrino + photon background the
 * reference's tests use (delta_tot_table_test.c:25-45); it must be configured first. */
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include "ksn_host.h"

static const _omega_nu *bg_omnu;
static double bg_nonu, bg_lambda, bg_h0;
static int quiet = 0;

void ksn_set_quiet(int q) { quiet = q; }

void ksn_set_default_hubble(const _omega_nu *omnu, double Omega0, double UnitTime_in_s)
{
    bg_omnu = omnu ? omnu : ksn_global_omnu();
    bg_nonu = Omega0 - get_omega_nu(bg_omnu, 1);
    bg_lambda = 1 - Omega0;
    bg_h0 = HUBBLE * UnitTime_in_s;
    ksn_invalidate_background();
}

__attribute__((weak)) double hubble_function(double a)
{
    if (!bg_omnu) terminate(1, "hubble_function: no host Hubble rate linked and ksn_set_default_hubble() not called\n");
    const double om = bg_nonu / (a * a * a) + bg_lambda + get_omega_nu(bg_omnu, a) + get_omegag(bg_omnu, a);
    return bg_h0 * sqrt(om);
}

__attribute__((weak)) void terminate(int ierr, const char *fmt, ...)
{
    va_list va;
    va_start(va, fmt);
    vfprintf(stderr, fmt, va);
    va_end(va);
    fflush(NULL);
    exit(ierr);
}

__attribute__((weak)) void message(int ierr, const char *fmt, ...)
{
    if (quiet) return;
    if (ierr > 0 || delta_tot_table.ThisTask == 0) {
        va_list va;
        va_start(va, fmt);
        vprintf(fmt, va);
        va_end(va);
    }
}

__attribute__((weak)) void *mymalloc_fullinfo(const char *string, size_t size, const char *func, const char *file, int line)
{
    (void) string; (void) func; (void) file; (void) line;
    return malloc(size);
}

__attribute__((weak)) void myfree_fullinfo(void *ptr, const char *func, const char *file, int line)
{
    (void) func; (void) file; (void) line;
    free(ptr);
}

void ksn_fatal_device(int rc, const char *where)
{
    /* new codes beyond the reference's list: 3001 no device, 3002 CUDA/comm failure;
     * quadrature failures keep the reference's GSL-handler code 2001 (delta_tot_table.c:70-73) */
    if (rc == KSN_EQUAD) terminate(2001, "GSL_ERROR in %s: %s\n", where, ksn_last_error());
    if (rc == KSN_ENODEV) terminate(3001, "%s: no usable B200 device: %s\n", where, ksn_last_error());
    terminate(3002, "%s: device layer failed (%d): %s\n", where, rc, ksn_last_error());
}
