/* TEST INFRASTRUCTURE (oracle): FFTW2 type names only (powerspectrum.h:5-14 includes
 * <dfftw.h>/<sfftw.h> just for fftw_real / fftw_complex). */
#ifndef KSN_ORACLE_FFTW_TYPES_H
#define KSN_ORACLE_FFTW_TYPES_H
#ifndef FFTW_H
#define FFTW_H              /* FFTW2's own include guard: lets other headers see that the types exist */
#endif
#ifndef KSN_HAVE_FFTW_TYPES /* the product header's stand-in (include/kspace_neutrinos.h) got there first */
#ifdef DOUBLEPRECISION_FFTW
typedef double fftw_real;
#else
typedef float fftw_real;
#endif
typedef struct { fftw_real re, im; } fftw_complex;
#endif
#endif
