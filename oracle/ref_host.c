/* TEST INFRASTRUCTURE (oracle): the callbacks the reference imports from its host N-body
 * code (gadget_defines.h:11,24-32), for the oracle/_ref build of the reference sources.
 * hubble_function is the flat LCDM + massive neutrinos + photons background the
 * reference's own tests use (delta_tot_table_test.c:25-45), restated. */
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include "omega_nu_single.h"
#include "gadget_defines.h"

#include <string.h>
#include <time.h>

int ThisTask = 0;
int ksn_ref_quiet = 1;     /* 0: print, 1: silent, 2: silent but record when the phase messages arrive */
double ksn_ref_t_mass, ksn_ref_t_nupower;

static double stamp(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static const _omega_nu *bg_omnu;
static double bg_Omega_nonu, bg_OmegaLambda, bg_Hubble;

void ksn_ref_set_background(const _omega_nu *omnu, double Omega0, double UnitTime_in_s)
{
    bg_omnu = omnu;
    bg_Omega_nonu = Omega0 - get_omega_nu(omnu, 1);
    bg_OmegaLambda = 1 - Omega0;
    bg_Hubble = HUBBLE * UnitTime_in_s;
}

double hubble_function(double a)
{
    if (!bg_omnu) terminate(1, "ksn_ref_set_background was not called\n");
    double omega_tot = bg_Omega_nonu / pow(a, 3) + bg_OmegaLambda;
    omega_tot += get_omega_nu(bg_omnu, a);
    omega_tot += get_omegag(bg_omnu, a);
    return bg_Hubble * sqrt(omega_tot);
}

void terminate(int ierr, const char *fmt, ...)
{
    va_list va;
    va_start(va, fmt);
    vfprintf(stderr, fmt, va);
    va_end(va);
    fflush(NULL);
    exit(ierr);
}

void message(int ierr, const char *fmt, ...)
{
    if (ksn_ref_quiet == 2) {
        /* powerspectrum.c:96 and interface_common.c:130 mark the ends of the K1 and integral phases */
        if (!strncmp(fmt, "Total powerspectrum mass", 24)) ksn_ref_t_mass = stamp();
        else if (!strncmp(fmt, "Done getting neutrino power", 27)) ksn_ref_t_nupower = stamp();
    }
    if (ksn_ref_quiet) return;
    if (ierr > 0 || ThisTask == 0) {
        va_list va;
        va_start(va, fmt);
        vprintf(fmt, va);
        va_end(va);
    }
}

void *mymalloc_fullinfo(const char *string, size_t size, const char *func, const char *file, int line)
{
    (void) string; (void) func; (void) file; (void) line;
    return malloc(size);
}

void myfree_fullinfo(void *ptr, const char *func, const char *file, int line)
{
    (void) func; (void) file; (void) line;
    free(ptr);
}
