/* TEST INFRASTRUCTURE: golden mode counts and sum(|k|) per bin of total_powerspectrum (powerspectrum.c:33-89) for the whole
 * grid at a given PMGRID -- the part of the reference's output that does not depend on the data -- at sizes where the
 * reference itself cannot be run in test time (1024^3 ... 4096^3).
 *
 * Two independent steps:
 *   1. r[k2] = number of stored modes (i, j, k <= N/2) with kx^2+ky^2+kz^2 = k2, weighted with the reference's Hermitian
 *      multiplicity (1 for k = 0 and k = N/2, else 2; :63-87): pure integer arithmetic over one octant of (i, j), using
 *      that index i and N-i carry the same kx^2.
 *   2. bin(k2) by the reference's own expression as its -ffast-math build evaluates it, floor((binsperunit/2) log(k2)),
 *      binsperunit = (nrbins-1)/log(N sqrt(3)/2) (:40,67; see oracle/ksn_oracle.c), with this machine's libm -- the same
 *      one the reference would link.
 * Output (stdout): one line per bin "b count keffsum", keffsum = sum m*sqrt(k2) in long double.
 * Checked against the compiled reference at small sizes by tools/make_golden_geometry.py before its output is stored.
 *   gcc -O2 -o gen_geometry_golden tools/gen_geometry_golden.c -lm ; ./gen_geometry_golden N nrbins */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s N nrbins\n", argv[0]); return 2; }
    const int N = atoi(argv[1]), nrbins = atoi(argv[2]), H = N / 2;
    const long long k2max = 3LL * H * H;
    long long *r = calloc((size_t) k2max + 1, sizeof(long long));
    if (!r) { fprintf(stderr, "out of memory\n"); return 1; }
    for (int a = 0; a <= H; a++) {
        const long long wa = (a == 0 || a == H) ? 1 : 2;            /* indices i with kx^2 = a^2 */
        for (int b = 0; b <= H; b++) {
            const long long wab = wa * ((b == 0 || b == H) ? 1 : 2);
            long long *row = r + (long long) a * a + (long long) b * b;
            for (int c = 0; c <= H; c++) row[(long long) c * c] += wab * ((c == 0 || c == H) ? 1 : 2);
        }
    }
    const double halfbinsperunit = 0.5 * ((nrbins - 1) / log(N * 0.8660254037844386));
    long long *count = calloc(nrbins, sizeof(long long));
    long double *keff = calloc(nrbins, sizeof(long double));
    for (long long k2 = 1; k2 <= k2max; k2++) {
        if (!r[k2]) continue;
        const int bin = (int) floor(halfbinsperunit * log((double) k2));
        if (bin < 0 || bin >= nrbins) { fprintf(stderr, "k2 = %lld falls in bin %d\n", k2, bin); return 1; }
        count[bin] += r[k2];
        keff[bin] += (long double) r[k2] * sqrtl((long double) k2);
    }
    for (int b = 0; b < nrbins; b++) printf("%d %lld %.21Lg\n", b, count[b], keff[b]);
    return 0;
}
