/* kspace_neutrinos.h -- the complete reference-compatible C API of the B200 build.
 *
 * Every declaration below keeps the name, argument order, argument meaning and error
 * behaviour of the sbird/kspace-neutrinos symbol it replaces; the citation after each
 * one is the reference declaration (file:line under /root/reference).  Structs keep the
 * reference's member names, order and types because host codes and the reference's own
 * tests read and write them directly.  The per-header forwarding files next to this one
 * (interface_gadget.h, powerspectrum.h, ...) exist so that a PM code keeps its
 * "#include "kspace-neutrinos/interface_gadget.h"" line unchanged.
 *
 * Grid precision is a build-time choice exactly as in the reference
 * (powerspectrum.h:6-14): define DOUBLEPRECISION_FFTW for a double grid.  Both
 * precisions are compiled into the shared library; the typed entry points are
 * total_powerspectrum / add_nu_power_to_rhogrid / compute_total_power_spectrum
 * (bound to the _f64 or _f32 symbol by the macros at the end of this file).
 */
#ifndef KSPACE_NEUTRINOS_B200_API_H
#define KSPACE_NEUTRINOS_B200_API_H

#include <stddef.h>
#include "ksn_mpi.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ constants
 * kspace_neutrino_const.h:9-19, gadget_defines.h:8,38-40 */
#define NUSPECIES 3
#define LIGHTCGS 2.99792458e10
#define BOLEVK 8.61734e-5
#define FLOAT_ACC 1e-6
#define GSL_VAL 200
#ifndef HUBBLE
#define HUBBLE 3.24077929e-18
#endif
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
/* T_nu / T_cmb, omega_nu_single.h:17 */
#define TNUCMB (pow(4 / 11., 1 / 3.) * 1.00328)
/* parameter-file type tags used by set_kspace_vars, gadget_defines.h:38-40 */
#ifndef KSN_NO_GADGET_TAGS
#ifndef REAL
#define REAL 1
#endif
#ifndef STRING
#define STRING 2
#endif
#ifndef INT
#define INT 3
#endif
#endif

/* ------------------------------------------------------------------ grid element
 * powerspectrum.h:5-14 takes fftw_complex from FFTW2's headers (<fftw.h>, <dfftw.h> or <sfftw.h>
 * by NOTYPEPREFIX_FFTW / DOUBLEPRECISION_FFTW).  Do the same where the host's include path has
 * them, so that a source which includes FFTW *after* this header (powerspectrum_test.c:7-15)
 * still sees one definition; otherwise provide the identical layout. */
#if !defined(FFTW_H) && !defined(KSN_HAVE_FFTW_TYPES) && defined(__has_include)
#if defined(NOTYPEPREFIX_FFTW)
#if __has_include(<fftw.h>)
#include <fftw.h>
#endif
#elif defined(DOUBLEPRECISION_FFTW)
#if __has_include(<dfftw.h>)
#include <dfftw.h>
#endif
#else
#if __has_include(<sfftw.h>)
#include <sfftw.h>
#endif
#endif
#endif
#if !defined(FFTW_H) && !defined(KSN_HAVE_FFTW_TYPES)
#define KSN_HAVE_FFTW_TYPES
#ifdef DOUBLEPRECISION_FFTW
typedef double fftw_real;
#else
typedef float fftw_real;
#endif
typedef struct { fftw_real re, im; } fftw_complex;
#endif

/* Interpolator members: the reference structs hold gsl_interp* / gsl_interp_accel*
 * (delta_pow.h:13-14, omega_nu_single.h:24-25).  This build has no GSL; the members
 * keep their names and pointer size and point at private tables. */
#ifndef __GSL_INTERP_H__
typedef struct ksn_interp_s gsl_interp;
typedef struct ksn_accel_s gsl_interp_accel;
#endif

/* ------------------------------------------------------------------ imports
 * What the host N-body code must provide (gadget_defines.h:11,24-32).  The shared
 * library carries weak fallbacks (malloc/free/printf/exit and a flat-LCDM Hubble rate,
 * see ksn_set_default_hubble) so that it also loads stand-alone. */
double hubble_function(double a);
void terminate(int ierr, const char *fmt, ...);
void message(int ierr, const char *fmt, ...);
void *mymalloc_fullinfo(const char *string, size_t size, const char *func, const char *file, int line);
void myfree_fullinfo(void *ptr, const char *func, const char *file, int line);
#ifndef mymalloc
#define mymalloc(x, y) mymalloc_fullinfo(x, y, __func__, __FILE__, __LINE__)
#define myfree(x) myfree_fullinfo(x, __func__, __FILE__, __LINE__)
#endif

/* ------------------------------------------------------------------ Omega_nu(a)
 * omega_nu_single.h:21-29,51-64,94-108 */
struct _rho_nu_single {
    double *loga;
    double *rhonu;
    gsl_interp *interp;
    gsl_interp_accel *acc;
    double mnu;
};
typedef struct _rho_nu_single _rho_nu_single;

struct _hybrid_nu {
    int enabled;
    double nufrac_low[NUSPECIES];
    double nu_crit_time;
    double vcrit;
};
typedef struct _hybrid_nu _hybrid_nu;

struct _omega_nu {
    _rho_nu_single *RhoNuTab[NUSPECIES];
    int nu_degeneracies[NUSPECIES];
    double rhocrit;
    double kBtnu;
    double tcmb0;
    _hybrid_nu hybnu;
};
typedef struct _omega_nu _omega_nu;

void rho_nu_init(_rho_nu_single *rho_nu_tab, double a0, const double mnu, const double HubbleParam, const double kBtnu); /* omega_nu_single.h:37 */
double rho_nu(_rho_nu_single *rho_nu_tab, const double a, const double kT);                     /* :45 */
void init_hybrid_nu(_hybrid_nu *const hybnu, const double mnu[], const double vcrit, const double light, const double nu_crit_time, const double kBtnu); /* :76 */
double particle_nu_fraction(const _hybrid_nu *const hybnu, const double a, int i);              /* :85 */
double nufrac_low(const double qc);                                                             /* :88 */
double rho_nu_int(double q, void *params);      /* omega_nu_single.c:89 (undeclared there; omega_nu_single_test.c:72); params = {a*mnu, kT} */
double get_rho_nu_conversion(void);              /* omega_nu_single.c:100 (omega_nu_single_test.c:68) */
void init_omega_nu(_omega_nu *const omnu, const double MNu[], const double a0, const double HubbleParam, const double tcmb0); /* :116 */
double get_omega_nu(const _omega_nu *const omnu, const double a);                               /* :119 */
double get_omega_nu_nopart(const _omega_nu *const omnu, const double a);                        /* :122 */
double get_omegag(const _omega_nu *const omnu, const double a);                                 /* :125 */
double omega_nu_single(const _omega_nu *const rho_nu_tab, const double a, const int i);         /* :131 */

/* ------------------------------------------------------------------ CAMB transfer table
 * transfer_init.h:12-18,28,31 */
struct _transfer_init_table {
    int NPowerTable;
    double *logk;
    double *T_nu;
};
typedef struct _transfer_init_table _transfer_init_table;
void allocate_transfer_init_table(_transfer_init_table *t_init, const double BoxSize, const double UnitLength_in_cm, const double InputSpectrum_UnitLength_in_cm, const char *KspaceTransferFunction);
void free_transfer_init_table(_transfer_init_table *t_init);

/* ------------------------------------------------------------------ delta_nu/delta_cdm table
 * delta_pow.h:10-18,30,39,42 */
struct _delta_pow {
    double *logkk;
    double *delta_ratio;
    gsl_interp *spline;
    gsl_interp_accel *acc;
    int nbins;
    double norm;
};
typedef struct _delta_pow _delta_pow;
void init_delta_pow(_delta_pow *d_pow, double logkk[], double delta_ratio[], int nbins, double norm);
double get_dnudcdm_powerspec(_delta_pow *d_pow, double kk);
void free_d_pow(_delta_pow *d_pow);

/* ------------------------------------------------------------------ integrator state
 * delta_tot_table.h:17-58 */
struct _delta_tot_table {
    int nk;
    int nk_allocated;
    int namax;
    int ia;
    int ThisTask;
    double delta_nu_prefac;
    int delta_tot_init_done;
    int debug;
    double **delta_tot;
    double *scalefact;
    double *delta_nu_init;
    double *delta_nu_last;
    double *wavenum;
    const _omega_nu *omnu;
    double Omeganonu;
    double light;
    double TimeTransfer;
};
typedef struct _delta_tot_table _delta_tot_table;

void allocate_delta_tot_table(_delta_tot_table *d_tot, const int nk_in, const double TimeTransfer, const double TimeMax, const double Omega0, const _omega_nu *const omnu, const double UnitTime_in_s, const double UnitLength_in_cm, int debug); /* delta_tot_table.h:70 */
void free_delta_tot_table(_delta_tot_table *d_tot);                                             /* :73 */
void delta_tot_init(_delta_tot_table *const d_tot, const int nk_in, const double wavenum[], const double delta_cdm_curr[], const _transfer_init_table *const t_init, const double Time); /* :85 */
void update_delta_tot(_delta_tot_table *const d_tot, const double a, const double delta_cdm_curr[], const double delta_nu_curr[], const int overwrite); /* :90 */
/* The per-step integrator entry; the linear-response integral inside runs on the GPU. */
void get_delta_nu_update(_delta_tot_table *const d_tot, const double a, const int nk_in, const double keff[], const double P_cdm_curr[], double delta_nu_curr[], _transfer_init_table *transfer_init); /* :104 */
void get_delta_nu(const _delta_tot_table *const d_tot, const double a, const double wavenum[], double delta_nu_curr[], const double mnu); /* :113 */
void get_delta_nu_combined(const _delta_tot_table *const d_tot, const double a, const double wavenum[], double delta_nu_curr[]); /* :117 */
void save_delta_tot(const _delta_tot_table *const d_tot, const int iia, char *savedir);         /* :120 */
void save_all_nu_state(const _delta_tot_table *const d_tot, char *savedir);                     /* :123 */
int save_nu_power(const _delta_tot_table *const d_tot, const double Time, const int snapnum, const char *OutputDir); /* :131 */
void read_all_nu_state(_delta_tot_table *const d_tot, char *savedir);                           /* :135 */
double specialJ(const double x, const double vcmnubylight, const double nufrac_low);            /* :138 */
double fslength(const double logai, const double logaf, const double light);                    /* :149 */
double get_delta_tot(const double delta_nu_curr, const double delta_cdm_curr, const double OmegaNua3, const double Omeganonu, const double Omeganu1, const double partnu); /* :155 */

/* ------------------------------------------------------------------ module parameters
 * interface_common.h:12-30 (the host's parameter reader writes into this object) */
extern struct __kspace_params {
    char KspaceTransferFunction[500];
    double TimeTransfer;
    double InputSpectrum_UnitLength_in_cm;
    double MNu[NUSPECIES];
    int hybrid_neutrinos_on;
    double vcrit;
    double nu_crit_time;
} kspace_params;

/* ------------------------------------------------------------------ generic interface
 * interface_common.h:37-121 */
double OmegaNu(double a);                                                                       /* :37 */
double OmegaNu_nopart(double a);                                                                /* :44 */
void InitOmegaNu(const double HubbleParam, const double tcmb0, MPI_Comm MYMPI_COMM_WORLD);      /* :46 */
void allocate_kspace_memory(const int nk_in, const int ThisTask, const double BoxSize, const double UnitTime_in_s, const double UnitLength_in_cm, const double Omega0, char *snapdir, const double TimeMax, MPI_Comm MYMPI_COMM_WORLD); /* :66 */
_delta_pow compute_neutrino_power_from_cdm(const double Time, const double keff_in[], const double P_cdm[], const long int Nmodes[], const int nk_in, MPI_Comm MYMPI_COMM_WORLD); /* :80 */
void save_nu_state(char *savedir);                                                              /* :86 */
void get_nu_state(double **scalefact, double **delta_tot, size_t *nk, size_t *ia);              /* :97 */
void set_nu_state(double *scalefact, double *delta_tot, const size_t nk, const size_t ia, MPI_Comm MYMPI_COMM_WORLD); /* :108 */
int save_neutrino_power(const double Time, const int snapnum, const char *OutputDir);           /* :117 */
int particle_nu_active(double a);                                                               /* :121 */

/* ------------------------------------------------------------------ FFTW2-slab interface
 * interface_gadget.h:38-57 and powerspectrum.h:30.  `slabstart`/`nslab` are the range of
 * the slowest grid index this rank owns; the grid is nslab x pmgrid x (pmgrid/2+1)
 * complex values, row-major, caller-owned, modified in place.  The pointer may be host
 * memory (staged through the GPU) or device/managed memory (used in place). */
int set_kspace_vars(char tag[][50], void *addr[], int id[], int nt);                            /* interface_gadget.h:45 */
/* the reference picks the neutrino-aware variant with -DKSPACE_NEUTRINOS_2 (interface_gadget.c:199-219); this library
 * picks it at run time: once the integrator has been initialised by a PM step */
int save_total_power(const double Time, const int snapnum, const char *OutputDir);              /* :57 */

int total_powerspectrum_f64(const int dims, void *outfield, const int nrbins, const int startslab, const int nslab, double *power, long long int *count, double *keffs, const MPI_Comm MYMPI_COMM_WORLD);
int total_powerspectrum_f32(const int dims, void *outfield, const int nrbins, const int startslab, const int nslab, double *power, long long int *count, double *keffs, const MPI_Comm MYMPI_COMM_WORLD);
void add_nu_power_to_rhogrid_f64(const double Time, const double BoxSize, void *fft_of_rhogrid, const int pmgrid, int slabstart_y, int nslab_y, MPI_Comm MYMPI_COMM_WORLD);
void add_nu_power_to_rhogrid_f32(const double Time, const double BoxSize, void *fft_of_rhogrid, const int pmgrid, int slabstart_y, int nslab_y, MPI_Comm MYMPI_COMM_WORLD);
/* Extension, not in the reference: the neutrino correction and the PM Green's function that follows it in
 * pmforce_periodic (gadget-2/0002 patch:116-125 context), in one pass over the grid.  asmth2 = (2 pi Asmth / BoxSize)^2. */
void add_nu_power_and_greens_to_rhogrid_f64(const double Time, const double BoxSize, void *fft_of_rhogrid, const int pmgrid, int slabstart_y, int nslab_y, const double asmth2, MPI_Comm MYMPI_COMM_WORLD);
void add_nu_power_and_greens_to_rhogrid_f32(const double Time, const double BoxSize, void *fft_of_rhogrid, const int pmgrid, int slabstart_y, int nslab_y, const double asmth2, MPI_Comm MYMPI_COMM_WORLD);
void compute_total_power_spectrum_f64(const double Time, const double BoxSize, void *fft_of_rhogrid, const int pmgrid, int slabstart_y, int nslab_y, MPI_Comm MYMPI_COMM_WORLD);
void compute_total_power_spectrum_f32(const double Time, const double BoxSize, void *fft_of_rhogrid, const int pmgrid, int slabstart_y, int nslab_y, MPI_Comm MYMPI_COMM_WORLD);

#ifndef KSN_NO_TYPED_MACROS
#ifdef DOUBLEPRECISION_FFTW
/* powerspectrum.h:30 */
#define total_powerspectrum(dims, outfield, nrbins, startslab, nslab, power, count, keffs, comm) \
    total_powerspectrum_f64(dims, (fftw_complex *) (outfield), nrbins, startslab, nslab, power, count, keffs, comm)
/* interface_gadget.h:38 */
#define add_nu_power_to_rhogrid(Time, BoxSize, grid, pmgrid, slabstart_y, nslab_y, comm) \
    add_nu_power_to_rhogrid_f64(Time, BoxSize, (fftw_complex *) (grid), pmgrid, slabstart_y, nslab_y, comm)
/* interface_gadget.h:48 */
#define compute_total_power_spectrum(Time, BoxSize, grid, pmgrid, slabstart_y, nslab_y, comm) \
    compute_total_power_spectrum_f64(Time, BoxSize, (fftw_complex *) (grid), pmgrid, slabstart_y, nslab_y, comm)
#define add_nu_power_and_greens_to_rhogrid(Time, BoxSize, grid, pmgrid, slabstart_y, nslab_y, asmth2, comm) \
    add_nu_power_and_greens_to_rhogrid_f64(Time, BoxSize, (fftw_complex *) (grid), pmgrid, slabstart_y, nslab_y, asmth2, comm)
#else
#define total_powerspectrum(dims, outfield, nrbins, startslab, nslab, power, count, keffs, comm) \
    total_powerspectrum_f32(dims, (fftw_complex *) (outfield), nrbins, startslab, nslab, power, count, keffs, comm)
#define add_nu_power_to_rhogrid(Time, BoxSize, grid, pmgrid, slabstart_y, nslab_y, comm) \
    add_nu_power_to_rhogrid_f32(Time, BoxSize, (fftw_complex *) (grid), pmgrid, slabstart_y, nslab_y, comm)
#define compute_total_power_spectrum(Time, BoxSize, grid, pmgrid, slabstart_y, nslab_y, comm) \
    compute_total_power_spectrum_f32(Time, BoxSize, (fftw_complex *) (grid), pmgrid, slabstart_y, nslab_y, comm)
#define add_nu_power_and_greens_to_rhogrid(Time, BoxSize, grid, pmgrid, slabstart_y, nslab_y, asmth2, comm) \
    add_nu_power_and_greens_to_rhogrid_f32(Time, BoxSize, (fftw_complex *) (grid), pmgrid, slabstart_y, nslab_y, asmth2, comm)
#endif
#endif

#ifdef __cplusplus
}
#endif
#endif
