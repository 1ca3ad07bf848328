/* TEST INFRASTRUCTURE -- CPU checker for the Green's-function multiply that follows the neutrino hook in a Gadget-2 PM step
 * (SURVEY.md 8f row 1).  Only tests/ may call it; the product never links it.
 *
 * What it restates: GADGET-2, public release 2.0.7 (V. Springel, MNRAS 364, 1105 (2005); the code base the reference's
 * patches in gadget-2/ apply to), file pm_periodic.c, function pmforce_periodic(), the loop headed
 *     "multiply with Green's function for the potential"
 * that starts right after  rfftwnd_mpi(fft_forward_plan, 1, rhogrid, workspace, FFTW_TRANSPOSED_ORDER);  -- i.e. right
 * after the line where the reference inserts add_nu_power_to_rhogrid (/root/reference/gadget-2/
 * 0002-Add-hooks-to-kspace-neutrino-code.patch:114-125 shows both lines as diff context, pm_periodic.c:374-380 there).
 * GADGET-2 itself is NOT in /root/reference, so this file is written from the published source; the reference's own
 * scaling loop (interface_gadget.c:163-188) is that same loop with another factor -- same loop nest, same
 * kx/ky/kz/k2/smth/ip names, same index expression -- which is the anchor available inside the repository.
 * Pinning status: "restated from the published third-party source, no golden vector in the reference" (DESIGN.md 7).
 *
 * The published loop, statement for statement (x, y, z: grid indices; the slab index y is the slowest one because of
 * FFTW_TRANSPOSED_ORDER; asmth2 = (2 pi Asmth / BoxSize)^2 in units of the fundamental mode squared):
 *     k{x,y,z} = index > PMGRID/2 ? index - PMGRID : index;        k2 = kx^2 + ky^2 + kz^2
 *     if (k2 > 0) {
 *         smth = -exp(-k2 * asmth2) / k2;
 *         f{x,y,z} = 1, or sin(pi k / PMGRID) / (pi k / PMGRID) where k != 0          (CIC deconvolution, applied twice:
 *         ff = 1 / (fx * fy * fz);   smth *= ff * ff * ff * ff;                        mass assignment + interpolation)
 *         fft_of_rhogrid[ip].re *= smth;   fft_of_rhogrid[ip].im *= smth;
 *     }
 *     after the loop:  if (slabstart_y == 0) fft_of_rhogrid[0].re = fft_of_rhogrid[0].im = 0.0;
 */
#define _GNU_SOURCE
#include <math.h>
#include <stddef.h>

#define GREENS_LOOP(REAL)                                                                                             \
    REAL *g = grid;                                                                                                   \
    for (long long y = slabstart_y; y < slabstart_y + nslab_y; y++)                                                   \
        for (int x = 0; x < PMGRID; x++)                                                                              \
            for (int z = 0; z < PMGRID / 2 + 1; z++) {                                                                \
                const int kx = x > PMGRID / 2 ? x - PMGRID : x;                                                       \
                const int ky = y > PMGRID / 2 ? (int) (y - PMGRID) : (int) y;                                         \
                const int kz = z > PMGRID / 2 ? z - PMGRID : z;                                                       \
                const double k2 = (double) kx * kx + (double) ky * ky + (double) kz * kz;                             \
                if (k2 > 0) {                                                                                         \
                    double smth = -exp(-k2 * asmth2) / k2;                                                            \
                    double fx = 1, fy = 1, fz = 1;                                                                    \
                    if (kx != 0) { fx = (M_PI * kx) / PMGRID; fx = sin(fx) / fx; }                                    \
                    if (ky != 0) { fy = (M_PI * ky) / PMGRID; fy = sin(fy) / fy; }                                    \
                    if (kz != 0) { fz = (M_PI * kz) / PMGRID; fz = sin(fz) / fz; }                                    \
                    const double ff = 1 / (fx * fy * fz);                                                             \
                    smth *= ff * ff * ff * ff;                                                                        \
                    const size_t ip = (size_t) PMGRID * (PMGRID / 2 + 1) * (size_t) (y - slabstart_y) + (size_t) (PMGRID / 2 + 1) * x + z; \
                    g[2 * ip] *= smth;                                                                                \
                    g[2 * ip + 1] *= smth;                                                                            \
                }                                                                                                     \
            }                                                                                                         \
    if (slabstart_y == 0 && nslab_y > 0) g[0] = g[1] = 0.0;

/* grid: nslab_y * PMGRID * (PMGRID/2+1) complex values (fftw_real pairs), modified in place */
void orc_gadget2_greens(void *grid, int is_double, int PMGRID, long long slabstart_y, long long nslab_y, double asmth2)
{
    if (is_double) { GREENS_LOOP(double) } else { GREENS_LOOP(float) }
}

/* the synthetic field of the benchmark, CPU restatement (synthetic_grid.h) */
#include "synthetic_grid.h"
void orc_fill_synthetic_grid(double *g, int N, long long startslab, long long nslab, unsigned long long seed, double slope)
{
    orc_fill_synthetic_slab(g, N, startslab, nslab, seed, slope);
}
