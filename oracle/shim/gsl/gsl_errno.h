/* TEST INFRASTRUCTURE (oracle): mini-GSL error interface (see oracle/mini_gsl.c). */
#ifndef KSN_MINIGSL_ERRNO_H
#define KSN_MINIGSL_ERRNO_H
enum { GSL_SUCCESS = 0, GSL_FAILURE = -1, GSL_EDOM = 1, GSL_EINVAL = 4, GSL_EFAILED = 5,
       GSL_EMAXITER = 11, GSL_EZERODIV = 12, GSL_EBADTOL = 13, GSL_EROUND = 18, GSL_ESING = 21 };
typedef void gsl_error_handler_t(const char *reason, const char *file, int line, int gsl_errno);
gsl_error_handler_t *gsl_set_error_handler(gsl_error_handler_t *new_handler);
void gsl_error(const char *reason, const char *file, int line, int gsl_errno);
#define GSL_ERROR(reason, gsl_errno) do { gsl_error(reason, __FILE__, __LINE__, gsl_errno); return gsl_errno; } while (0)
#define GSL_ERROR_VAL(reason, gsl_errno, value) do { gsl_error(reason, __FILE__, __LINE__, gsl_errno); return value; } while (0)
#endif
