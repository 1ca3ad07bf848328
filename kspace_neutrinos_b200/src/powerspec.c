/* total_powerspectrum for FFTW2-style slab-decomposed r2c grids: host wrapper around kernel K1.
 * Replaces powerspectrum.c:33-117.  The sweep and the cross-rank sum run on the GPU
 * (ksn_powerspectrum_sums); what stays here is what the reference does after its
 * MPI_Allreduce calls: normalisation by |F(0,0,0)|^2 and the mode count, and removal of empty
 * bins (powerspectrum.c:96-116), plus the host-libm bin-threshold table that makes the mode
 * counts bit-exact. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "ksn_host.h"

static struct { int dims, nrbins; unsigned int *thr; double *iw; } tabs;

/* The reference's bin of a mode with squared integer wave number k2: floor(binsperunit*log(sqrt(k2))) with
 * binsperunit = (nrbins-1)/log(sqrt(3)*dims/2.0) (powerspectrum.c:40,67) -- evaluated the way the reference's own build
 * evaluates it.  Its Makefile compiles with -ffast-math (Makefile:2), under which gcc folds log(sqrt(x)) into 0.5*log(x)
 * and sqrt(3)*dims/2.0 into dims*(sqrt(3)/2) (checked in the disassembly of oracle/_ref: log() is called on k2 and the
 * product is taken with a pre-halved binsperunit).  The form matters exactly where a mode sits on a bin edge: the corner
 * mode (N/2,N/2,N/2) has binsperunit*log(kk) == nrbins-1 up to one ulp, and at PMGRID=192 the literal form puts it in the
 * last bin while the reference's binary (and this table) put it one below.  -DKSN_LITERAL_BIN_FORMULA restores the
 * literal evaluation. */
static double ref_binsperunit(int dims, int nrbins)
{
#ifdef KSN_LITERAL_BIN_FORMULA
    return (nrbins - 1) / log(sqrt(3) * dims / 2.0);
#else
    return (nrbins - 1) / log(dims * 0.8660254037844386);      /* sqrt(3)/2, correctly rounded */
#endif
}

static double ref_bin(double binsperunit, unsigned int k2)
{
#ifdef KSN_LITERAL_BIN_FORMULA
    return floor(binsperunit * log(sqrt((double) k2)));
#else
    return floor((0.5 * binsperunit) * log((double) k2));
#endif
}

int ksn_bin_tables(int dims, int nrbins, const unsigned int **thresholds, const double **invwin)
{
    if (tabs.dims != dims || tabs.nrbins != nrbins) {
        free(tabs.thr);
        free(tabs.iw);
        tabs.thr = malloc(sizeof(unsigned int) * nrbins);
        tabs.iw = malloc(sizeof(double) * (dims / 2 + 1));
        if (!tabs.thr || !tabs.iw) return -1;
        const double binsperunit = ref_binsperunit(dims, nrbins);
        const unsigned int k2max = 3u * (unsigned int) (dims / 2) * (unsigned int) (dims / 2);
        tabs.thr[0] = 0;
        for (int b = 1; b < nrbins; b++) {
            /* smallest k2 whose bin is >= b: start from the analytic inverse, then walk with the
             * exact expression (monotone in k2) */
            double guess = exp(2.0 * b / binsperunit);
            unsigned int k = guess >= (double) k2max + 1 ? k2max + 1 : (unsigned int) guess;
            if (k < 1) k = 1;
            while (k > 1 && ref_bin(binsperunit, k - 1) >= b) k--;
            while (k <= k2max && ref_bin(binsperunit, k) < b) k++;
            tabs.thr[b] = k;
        }
        /* 1-D inverse CIC window, powerspectrum.c:8-12 */
        tabs.iw[0] = 1.0;
        for (int q = 1; q <= dims / 2; q++) tabs.iw[q] = M_PI * q / (dims * sin(M_PI * q / (double) dims));
        tabs.dims = dims;
        tabs.nrbins = nrbins;
    }
    *thresholds = tabs.thr;
    *invwin = tabs.iw;
    return 0;
}

/* powerspectrum.c:96-116: normalise, then squeeze out empty bins; returns the number kept */
int ksn_finish_powerspectrum(int nrbins, double total_mass2, double *power, long long *count, double *keffs)
{
    message(0, "Total powerspectrum mass: %g\n", sqrt(total_mass2));
    for (int i = 0; i < nrbins; i++) {
        power[i] /= total_mass2;
        if (count[i]) {
            keffs[i] /= count[i];
            power[i] /= count[i];
        }
    }
    int kept = 0;
    for (int i = 0; i < nrbins; i++) {
        if (!count[i]) continue;
        if (kept < i) {
            power[kept] = power[i];
            keffs[kept] = keffs[i];
            count[kept] = count[i];
        }
        kept++;
    }
    return kept;
}

static int total_powerspectrum_any(int real_bytes, const int dims, void *outfield, const int nrbins, const int startslab, const int nslab, double *power, long long int *count, double *keffs)
{
    const unsigned int *thr;
    const double *iw;
    double total_mass2 = 0;
    if (ksn_bin_tables(dims, nrbins, &thr, &iw)) terminate(1, "Could not allocate temporary memory for power spectra\n");
    const int rc = ksn_powerspectrum_sums(outfield, real_bytes, dims, nrbins, startslab, nslab, thr, iw, power, keffs, count, &total_mass2);
    if (rc) ksn_fatal_device(rc, "total_powerspectrum");
    return ksn_finish_powerspectrum(nrbins, total_mass2, power, count, keffs);
}

int total_powerspectrum_f64(const int dims, void *outfield, const int nrbins, const int startslab, const int nslab, double *power, long long int *count, double *keffs, const MPI_Comm comm)
{
    ksn_bind_comm(comm);
    return total_powerspectrum_any(8, dims, outfield, nrbins, startslab, nslab, power, count, keffs);
}

int total_powerspectrum_f32(const int dims, void *outfield, const int nrbins, const int startslab, const int nslab, double *power, long long int *count, double *keffs, const MPI_Comm comm)
{
    ksn_bind_comm(comm);
    return total_powerspectrum_any(4, dims, outfield, nrbins, startslab, nslab, power, count, keffs);
}

/* link-level drop-in name (powerspectrum.h:30): bound to the grid precision this library was built for */
#ifdef KSN_DEFAULT_F32
int total_powerspectrum(const int dims, void *outfield, const int nrbins, const int startslab, const int nslab, double *power, long long int *count, double *keffs, const MPI_Comm comm)
    __attribute__((alias("total_powerspectrum_f32")));
#else
int total_powerspectrum(const int dims, void *outfield, const int nrbins, const int startslab, const int nslab, double *power, long long int *count, double *keffs, const MPI_Comm comm)
    __attribute__((alias("total_powerspectrum_f64")));
#endif
