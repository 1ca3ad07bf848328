/* TEST/BENCH INFRASTRUCTURE: CPU restatement of the synthetic k-space Gaussian field of ksn_fill_synthetic_grid
 * (kspace_neutrinos_b200/csrc/ksn_device.cu: fill_synthetic_kernel) -- element (i,j,z) = sigma(|k|) (n1, n2), n ~ N(0,1)
 * from a counter-based generator keyed by (seed, global mode index), sigma^2 ~ |k|^slope, element (0,0,0) = (N^3, 0) --
 * so that the reference's CPU arm (oracle/ref_bench.c) is timed on the field the GPU arm is timed on.  Equal to the device
 * values up to the last bits of libm's log/exp/sincos against the device's (tests/test_k1_gpu.py pins it to 1e-13). */
#ifndef KSN_SYNTHETIC_GRID_H
#define KSN_SYNTHETIC_GRID_H
#define _GNU_SOURCE
#include <math.h>
#include <stddef.h>

static inline unsigned long long orc_mix64(unsigned long long x)      /* splitmix64 finaliser */
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

/* planes [startslab, startslab+nslab) of the N^3 grid into g (2 doubles per mode) */
static inline void orc_fill_synthetic_slab(double *g, int N, long long startslab, long long nslab, unsigned long long seed, double slope)
{
    const int L = N / 2 + 1;
    const size_t nel = (size_t) nslab * N * L;
    for (size_t e = 0; e < nel; e++) {
        const long long row = (long long) (e / L), i = startslab + row / N;
        const int z = (int) (e - (size_t) row * L), j = (int) (row % N);
        const unsigned long long gidx = (unsigned long long) ((i * N + j) * L + z);
        const double ki = i <= N / 2 ? (double) i : (double) (i - N), kj = j <= N / 2 ? j : j - N;
        const double k2 = ki * ki + kj * kj + (double) z * z;
        if (gidx == 0) { g[2 * e] = (double) N * N * N; g[2 * e + 1] = 0; continue; }
        const unsigned long long h1 = orc_mix64(seed ^ orc_mix64(2 * gidx)), h2 = orc_mix64(seed ^ orc_mix64(2 * gidx + 1));
        const double u1 = ((double) (h1 >> 11) + 0.5) * (1.0 / 9007199254740992.0);
        const double u2 = ((double) (h2 >> 11) + 0.5) * (1.0 / 9007199254740992.0);
        const double r = sqrt(-2.0 * log(u1)) * exp(0.25 * slope * log(k2));
        double s, c;
        sincos(2.0 * M_PI * u2, &s, &c);
        g[2 * e] = r * c;
        g[2 * e + 1] = r * s;
    }
}
#endif
