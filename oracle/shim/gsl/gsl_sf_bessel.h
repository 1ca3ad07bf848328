/* TEST INFRASTRUCTURE (oracle): mini-GSL spherical Bessel j0 (GSL specfunc/bessel_j.c). */
#ifndef KSN_MINIGSL_SF_BESSEL_H
#define KSN_MINIGSL_SF_BESSEL_H
double gsl_sf_bessel_j0(const double x);
#endif
