/* TEST INFRASTRUCTURE (oracle): mini-GSL 1-D interpolation (linear, natural cubic spline).
 * Restates GSL 2.x interpolation/{interp.c,accel.c,bsearch.c,linear.c,cspline.c} and
 * linalg/tridiag.c (symmetric solver). */
#ifndef KSN_MINIGSL_INTERP_H
#define KSN_MINIGSL_INTERP_H
#include <stddef.h>
typedef struct { size_t cache, miss_count, hit_count; } gsl_interp_accel;
typedef struct gsl_interp_type_s gsl_interp_type;
typedef struct { const gsl_interp_type *type; double xmin, xmax; size_t size; void *state; } gsl_interp;
extern const gsl_interp_type *gsl_interp_linear;
extern const gsl_interp_type *gsl_interp_cspline;
gsl_interp_accel *gsl_interp_accel_alloc(void);
void gsl_interp_accel_free(gsl_interp_accel *a);
size_t gsl_interp_accel_find(gsl_interp_accel *a, const double x_array[], size_t size, double x);
size_t gsl_interp_bsearch(const double x_array[], double x, size_t index_lo, size_t index_hi);
gsl_interp *gsl_interp_alloc(const gsl_interp_type *T, size_t n);
int gsl_interp_init(gsl_interp *obj, const double xa[], const double ya[], size_t size);
double gsl_interp_eval(const gsl_interp *obj, const double xa[], const double ya[], double x, gsl_interp_accel *a);
void gsl_interp_free(gsl_interp *interp);
#endif
