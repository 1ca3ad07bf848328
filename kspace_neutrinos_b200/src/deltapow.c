/* delta_nu/delta_cdm lookup table (host side of delta_pow.c:6-43).  The same table is what
 * kernel K3 stages in shared memory; this scalar version serves hosts that apply it themselves
 * (compute_neutrino_power_from_cdm) and the tests. */
#include <math.h>
#include <stdio.h>
#include "ksn_host.h"

void init_delta_pow(_delta_pow *d_pow, double logkk[], double delta_ratio[], int nbins, double norm)
{
    d_pow->logkk = logkk;
    d_pow->delta_ratio = delta_ratio;
    d_pow->nbins = nbins;
    d_pow->norm = norm;
    d_pow->acc = mymalloc("d_pow_acc", sizeof(struct ksn_accel_s));
    d_pow->spline = mymalloc("d_pow_interp", sizeof(struct ksn_interp_s));
    d_pow->acc->hint = 0;
    d_pow->spline->n = nbins;
    d_pow->spline->cubic = 0;
    d_pow->spline->c = NULL;
}

double get_dnudcdm_powerspec(_delta_pow *d_pow, double kk)
{
    const double first = d_pow->logkk[0], last = d_pow->logkk[d_pow->nbins - 1];
    if (kk < first) {
        /* modes just beyond the box size: hold P(k) constant; complain when far outside */
        if (kk < first - log(2)) fprintf(stderr, "trying to extract a k= %g < min stored = %g \n", kk, first);
        kk = first;
    }
    if (kk > last) kk = last;
    return d_pow->norm * ksn_linear_eval(d_pow->logkk, d_pow->delta_ratio, d_pow->nbins, kk, &d_pow->acc->hint);
}

void free_d_pow(_delta_pow *d_pow)
{
    myfree(d_pow->spline);
    myfree(d_pow->acc);
    d_pow->spline = NULL;
    d_pow->acc = NULL;
}
