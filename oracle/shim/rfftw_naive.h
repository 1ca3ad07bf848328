/* TEST INFRASTRUCTURE (oracle): the four rfftwnd calls powerspectrum_test.c:33-46 makes,
 * as an O(n^2) DFT good for the 4^3 known-answer grid only. */
#ifndef KSN_ORACLE_RFFTW_NAIVE_H
#define KSN_ORACLE_RFFTW_NAIVE_H
#include "fftw_types.h"
#include <math.h>
#include <stdlib.h>
#define FFTW_FORWARD (-1)
#define FFTW_ESTIMATE 0
#define FFTW_IN_PLACE 8
typedef struct { int nx, ny, nz; } *rfftwnd_plan;
static inline rfftwnd_plan rfftw3d_create_plan(int nx, int ny, int nz, int dir, int flags)
{
    (void) dir; (void) flags;
    rfftwnd_plan p = malloc(sizeof(*p));
    p->nx = nx; p->ny = ny; p->nz = nz;
    return p;
}
static inline void rfftwnd_destroy_plan(rfftwnd_plan p) { free(p); }
/* in-place layout: real input padded to 2*(nz/2+1) along the last axis */
static inline void rfftwnd_one_real_to_complex(rfftwnd_plan p, fftw_real *in, fftw_complex *out)
{
    const int nx = p->nx, ny = p->ny, nz = p->nz, nzc = nz / 2 + 1, pad = 2 * nzc;
    double *tmp = malloc(sizeof(double) * nx * ny * nz);
    for (int x = 0; x < nx; x++) for (int y = 0; y < ny; y++) for (int z = 0; z < nz; z++)
        tmp[(x * ny + y) * nz + z] = in[(x * ny + y) * pad + z];
    for (int a = 0; a < nx; a++) for (int b = 0; b < ny; b++) for (int c = 0; c < nzc; c++) {
        double re = 0, im = 0;
        for (int x = 0; x < nx; x++) for (int y = 0; y < ny; y++) for (int z = 0; z < nz; z++) {
            const double ph = -2 * M_PI * ((double) a * x / nx + (double) b * y / ny + (double) c * z / nz);
            re += tmp[(x * ny + y) * nz + z] * cos(ph);
            im += tmp[(x * ny + y) * nz + z] * sin(ph);
        }
        out[(a * ny + b) * nzc + c].re = (fftw_real) re;
        out[(a * ny + b) * nzc + c].im = (fftw_real) im;
    }
    free(tmp);
}
#endif
