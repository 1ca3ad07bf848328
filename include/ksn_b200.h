/* ksn_b200.h -- the thin C-ABI between the host C code and the CUDA (sm_100a) kernels.
 *
 * Plain pointers and sizes only.  Host C (kspace_neutrinos_b200/src) calls these; so can
 * any FFI (ctypes stubs in kspace_neutrinos_b200/capi.py, cgo/JNI sketches in
 * INTEGRATION.md).  Every function returns 0 on success or a negative KSN_E* code and
 * records a message retrievable with ksn_last_error(); nothing here calls terminate().
 * There is no CPU fallback: without a usable CUDA device every compute entry fails with
 * KSN_ENODEV.
 *
 * Reference loops replaced (file:line under /root/reference):
 *   ksn_powerspectrum_sums  <- powerspectrum.c:56-95   (hot loop 1 + the 4 MPI_Allreduce)
 *   ksn_delta_nu_integrate  <- delta_tot_table.c:522-599 (fslength table + per-k QAG)
 *   ksn_scale_modes         <- interface_gadget.c:163-188 (hot loop 3)
 */
#ifndef KSN_B200_H
#define KSN_B200_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

enum {
    KSN_OK = 0,
    KSN_ENODEV = -1,     /* no CUDA device / driver */
    KSN_ECUDA = -2,      /* a CUDA runtime call failed */
    KSN_EINVAL = -3,     /* bad argument */
    KSN_ENOMEM = -4,     /* device or pinned allocation failed */
    KSN_ECOMM = -5,      /* collective backend failed */
    KSN_EQUAD = -6       /* quadrature did not converge (GSL would have raised) */
};

/* ---- lifecycle ---------------------------------------------------------------- */
/* Bind this process to a CUDA device (default: env KSN_DEVICE, else LOCAL_RANK, else 0).
 * Called implicitly by the first compute entry. */
int ksn_init(int device);
void ksn_shutdown(void);
const char *ksn_last_error(void);
/* 1 if a CUDA device is usable from this process, else 0 (never fails). */
int ksn_device_available(void);
/* number of CUDA devices visible to this process (0 if there is none or no driver; never fails) */
int ksn_device_count(void);
/* device selected by ksn_init, -1 before */
int ksn_device(void);

/* ---- collective backend --------------------------------------------------------- */
/* One rank (default). */
int ksn_comm_single(void);
/* NCCL over NVLink/NVSwitch, one process per GPU: rank 0 calls ksn_comm_nccl_unique_id,
 * ships the 128 bytes to the other ranks by any means (torch.distributed, MPI_Bcast, a
 * file), then every rank calls ksn_comm_nccl_init. */
int ksn_comm_nccl_unique_id(void *id128);
int ksn_comm_nccl_init(const void *id128, int nranks, int rank);
/* Peer memory over NVLink/NVSwitch, one process per GPU of ONE box: the cross-rank sum of the bin sums is done INSIDE
 * the kernel that produces them (every rank stores its values into mailboxes in all ranks' HBM, waits for the others'
 * flags and adds the contributions in rank order: no collective launch, bit-identical sums on every rank).  Every rank
 * calls ksn_comm_p2p_export (allocates its mailbox, returns the 64-byte CUDA-IPC handle), the host gathers the handles
 * in rank order by any means, then every rank calls ksn_comm_p2p_init with all nranks * 64 bytes (at most 16 ranks). */
int ksn_comm_p2p_export(void *handle64);
int ksn_comm_p2p_init(const void *handles, int nranks, int rank);
/* Host all-reduce callback: buf (n doubles, host memory) must be summed over ranks in
 * place.  This is what an MPI host uses (MPI_Allreduce on its communicator). */
typedef int (*ksn_allreduce_fn)(double *buf, size_t n, void *user);
int ksn_comm_host_callback(ksn_allreduce_fn fn, void *user, int nranks, int rank);
/* Sum n host doubles over the ranks of the active backend, in place (identity for one rank). */
int ksn_comm_allreduce_host(double *buf, size_t n);
int ksn_comm_rank(void);
int ksn_comm_size(void);

/* ---- memory helpers (for harnesses that want the grid resident in HBM) ------------ */
int ksn_device_malloc(void **ptr, size_t bytes);
int ksn_device_free(void *ptr);
int ksn_host_alloc_pinned(void **ptr, size_t bytes);
int ksn_host_free_pinned(void *ptr);
/* Page-lock a host grid the caller owns so that staged copies run at full PCIe speed.  The caller
 * must ksn_host_unregister() it before freeing it.  Without this, host grids are copied as
 * pageable memory (correct, slower). */
int ksn_host_register(void *ptr, size_t bytes);
int ksn_host_unregister(void *ptr);
int ksn_memcpy_h2d(void *dst, const void *src, size_t bytes);
int ksn_memcpy_d2h(void *dst, const void *src, size_t bytes);
int ksn_memcpy_d2d(void *dst, const void *src, size_t bytes);
int ksn_device_synchronize(void);
/* 1 = device or managed pointer, 0 = host pointer */
int ksn_pointer_is_device(const void *ptr);
/* Fill a device slab with the synthetic k-space Gaussian field used by bench.py and the
 * parity tests: element (i,j,k) = sigma(|k|) * (n1, n2), n ~ N(0,1) from a counter-based
 * generator keyed by (seed, global mode index), so any slab split yields the same values.
 * sigma^2 ~ |k|^slope (slope 0 = white); element (0,0,0) = (dims^3, 0). */
int ksn_fill_synthetic_grid(void *dgrid, int real_bytes, int dims, long long startslab, long long nslab,
                            unsigned long long seed, double slope);

/* ---- K1: power-spectrum bin sums (powerspectrum.c:33-95) --------------------------- */
/* thresholds[b] = smallest integer k^2 whose reference bin index
 * floor(binsperunit*log(sqrt(k^2))) is >= b, b = 0..nrbins-1, computed by the caller
 * with the host libm (bit-exact mode counts depend on it); invwin[q] = pi q/(dims
 * sin(pi q/dims)), q = 0..dims/2, invwin[0] = 1.
 * Outputs are the sums over ALL ranks of the active communicator (length nrbins each):
 * power_sum[b] = sum m*|F|^2*W, keff_sum[b] = sum m*|k|, count[b] = sum m, and
 * *total_mass2 = |F(0,0,0)|^2.  keff_sum and count depend only on the geometry
 * (dims, nrbins, slab) and are cached between calls unless KSN_NO_GEOM_CACHE=1.
 * grid: device pointer (used in place) or host pointer (staged through HBM in chunks). */
int ksn_powerspectrum_sums(const void *grid, int real_bytes, int dims, int nrbins,
                           long long startslab, long long nslab,
                           const unsigned int *thresholds, const double *invwin,
                           double *power_sum, double *keff_sum, long long *count, double *total_mass2);

/* ---- K3: scale every mode by 1 + norm*interp(log k) (interface_gadget.c:163-188) ---- */
/* logkk/ratio: the _delta_pow table (delta_pow.c:19-37: clamped, piecewise linear in
 * log k); a mode with integer wave vector n has k = |n| * 2 pi / boxsize. */
int ksn_scale_modes(void *grid, int real_bytes, int dims, long long startslab, long long nslab,
                    double boxsize, const double *logkk, const double *ratio, int nbins, double norm);

/* Fused hot path for a host-resident grid: upload once, K1, callback, K3, download.
 * between(): called after the bin sums are known; must fill the _delta_pow table. */
typedef int (*ksn_between_fn)(void *user, const double *power_sum, const double *keff_sum,
                              const long long *count, double total_mass2,
                              const double **logkk, const double **ratio, int *nbins, double *norm);
int ksn_step_staged(void *hgrid, int real_bytes, int dims, int nrbins, long long startslab, long long nslab,
                    const unsigned int *thresholds, const double *invwin, double boxsize,
                    ksn_between_fn between, void *user);

/* ---- K3 fused with the PM Green's function -------------------------------------
 * A Gadget-style PM step multiplies the same grid, right after the neutrino correction, by the periodic Green's
 * function with CIC deconvolution (Gadget-2 pm_periodic.c, the loop following the hook of gadget-2/0002
 * patch:116-125):  G(k) = -exp(-k2*asmth2)/k2 * (iwx iwy iwz)^4, F(0,0,0) = 0, k2 in grid units, iw(q) = 1/sinc(pi q/N),
 * asmth2 = (2 pi Asmth / BoxSize)^2.  These entries apply (1 + norm*interp) * G in ONE pass, which saves the host's own
 * read-modify-write traversal (32 B per mode).  invwin: iw[0..dims/2], as for ksn_powerspectrum_sums. */
int ksn_scale_modes_greens(void *grid, int real_bytes, int dims, long long startslab, long long nslab, double boxsize,
                           const double *logkk, const double *ratio, int nbins, double norm,
                           const double *invwin, double asmth2);
int ksn_step_staged_greens(void *hgrid, int real_bytes, int dims, int nrbins, long long startslab, long long nslab,
                           const unsigned int *thresholds, const double *invwin, double boxsize,
                           ksn_between_fn between, void *user, double asmth2);


/* ---- K2: linear-response integral (delta_tot_table.c:507-611) ----------------------- */
typedef struct ksn_delta_nu_args {
    int nk;                    /* k bins */
    int Na;                    /* stored rows incl. the guess row (d_tot->ia) */
    int namax;                 /* row stride of delta_tot */
    int nspecies;              /* independent mass species integrated in this launch (<=3) */
    double a;                  /* current scale factor */
    double TimeTransfer;       /* a0 */
    double light;              /* c in internal units */
    double delta_nu_prefac;    /* 1.5 Omega0 H0^2 / c */
    double deriv_prefac;       /* a0^2 H(a0) / c  (delta_tot_table.c:524) */
    double mnubykT[3];         /* m_nu / (k_B T_nu) per species */
    double qc[3];              /* hybrid momentum cut per species, 0 = none (:545) */
    double nufrac_low0;        /* hybnu.nufrac_low[0] (:533,569) */
    double relerr[3];          /* QAG epsrel per species (:518,547) */
    int integrate[3];          /* 0: initial-condition term only (:543-544,551) */
    const double *scalefact;   /* host, Na log-a knots */
    const double *delta_tot;   /* host, nk rows of stride namax */
    const double *wavenum;     /* host, nk */
    const double *delta_nu_init; /* host, nk */
} ksn_delta_nu_args;
/* out: nspecies*nk doubles, species-major.  n_evals (may be NULL): integrand evaluations. */
int ksn_delta_nu_integrate(const ksn_delta_nu_args *args, double *out, unsigned long long *n_evals);

/* Optional: start the part of the integral that depends only on the scale factor and the stored knots -- the free-streaming
 * table (16 Na quadratures of 1/(a^2 H)) and the spline factorisations -- on a side stream, so that it runs beside K1
 * instead of between K1 and K3.  scalefact: the Na knots (log a) the coming ksn_delta_nu_integrate call will pass,
 * i.e. the stored rows followed by log(a) of the step; namax: row capacity (sizes the buffers once).  The integrate call
 * uses the result only if a, TimeTransfer, light, every knot and the background table are bit for bit what it is
 * called with; otherwise it computes the tables itself.  Never needed for correctness. */
int ksn_delta_nu_prefetch(double a, double TimeTransfer, double light, const double *scalefact, int Na, int namax);
/* 1 if the most recent ksn_delta_nu_integrate call took its a-only tables from a prefetch */
int ksn_last_k2_prefetch_used(void);

/* the tile shape K1's tile kernel takes for (dims, nrbins) on a device with smem_budget bytes of opt-in shared memory per
 * CTA and `sms` SMs (host arithmetic only): warps per CTA, modes per lane, stages, tiles per row, first bin kept in shared
 * memory (0 = all).  Returns 0 when no tile shape fits. */
int ksn_k1_tile_plan(int dims, int nrbins, size_t smem_budget, int sms, int *warps, int *chunk, int *stages, int *tiles_per_row, int *hot_lo);
/* the same for a grid of real_bytes-wide reals (8: as above; 4: float rows, KSN_K1_F32_TILE) */
int ksn_k1_tile_plan_ex(int dims, int nrbins, size_t smem_budget, int sms, int real_bytes, int *warps, int *chunk, int *stages, int *tiles_per_row, int *hot_lo);
/* which K1 kernel (and tile configuration) the most recent power-spectrum sweep launched */
const char *ksn_last_k1_kernel(void);
/* which K3 kernel the most recent scaling pass launched (template instance, rows per CTA / pieces per row) */
const char *ksn_last_k3_kernel(void);
/* the _delta_pow table (and norm, box size) the most recent scaling pass used: pointers into a host copy owned by the
 * library, valid until the next pass.  For checkers: a pass can then be compared with the CPU restatement of
 * interface_gadget.c:163-188 on exactly the table it applied. */
int ksn_last_k3_table(const double **logkk, const double **ratio, int *nbins, double *norm, double *boxsize);
/* how the scaling passes would evaluate this table (host arithmetic only, no device needed): series = 5 or 9 terms of
 * ln(1+u) in the double passes; f32_ok = 1 if a float grid's pass may evaluate the factor in float; k2_narrow = the k^2 from
 * which on every table segment is narrow, i.e. rows with kx^2 + ky^2 >= k2_narrow take the branch-free path (0xffffffff:
 * none); cells = lookup cells in log2(k^2); multi = 1 if some cell holds more than one knot.  Any output may be NULL. */
int ksn_k3_table_plan(int dims, double boxsize, const double *logkk, const double *ratio, int nbins, double norm,
                      int *series, int *f32_ok, unsigned *k2_narrow, int *cells, int *multi);
/* FNV-1a hash of everything the host builds for this table -- the words uploaded to the device, the kernels' parameters,
 * the decisions above.  The part that depends on the knots alone is kept from one build to the next (the knots are
 * log(keff): the same every step); fresh != 0 forgets it first, so a test can pin cached == fresh, bit for bit. */
int ksn_k3_table_hash(int dims, double boxsize, const double *logkk, const double *ratio, int nbins, double norm,
                      int fresh, unsigned long long *hash);
/* integrand evaluations of the most recent ksn_delta_nu_integrate call (fslength table included) */
unsigned long long ksn_last_k2_evals(void);
/* largest number of 61-point rule applications any single k bin needed in that call (the kernel's critical path) */
unsigned ksn_last_k2_max_passes(void);
/* passes through the integrand the slowest k bin of that call made (= bisections + 1 for QAG's sequential loop; fewer
 * when bisections are integrated ahead of time, see ksn_k2_spec_width) */
unsigned ksn_last_k2_max_trips(void);
/* how many intervals a K2 CTA bisects per pass through the integrand: 1 = QAG's sequential loop, 2..4 = the halves of
 * the 2..4 worst intervals are integrated at once and the loop is replayed over the cached results (bit-identical
 * output, shorter critical path); 0 = chosen per call by regime (hybrid neutrinos with one species: 3, otherwise 1).
 * Environment KSN_K2_SPEC overrides the built-in default (0). */
int ksn_k2_spec_width(void);

/* Tabulate 1/(a H(a)) for the device integrand.  hub(a, user) is called on the host at
 * n uniform points in log a over [loga_lo, loga_hi]. */
typedef double (*ksn_hubble_fn)(double a, void *user);
int ksn_set_background(ksn_hubble_fn hub, void *user, double loga_lo, double loga_hi, int n);
/* how the last table came out: cells whose 4-point interpolant missed the host function (kinks in H(a), e.g. the
 * spline/series switch of Omega_nu, omega_nu_single.c:180-199) and the refined patches that cover them */
int ksn_background_info(int *npatch, int *flagged_cells);
/* 1 while a table from ksn_set_background lives on the device (0 after ksn_shutdown: callers that cache "already set" must ask) */
int ksn_background_loaded(void);
/* device fslength (same table), for tests: light * int_{logai}^{logaf} dloga /(a^2 H) */
int ksn_fslength_device(const double *logai, int n, double logaf, double light, double *out);

/* ---- slab-decomposed FFT of the PM grid, device resident -----------------------------------------
 * Replaces, for a GPU-resident PM code, the transform in front of the hook in Gadget-2's pmforce_periodic
 * (gadget-2/0002 patch:116):  rfftwnd_mpi(fft_forward_plan, 1, rhogrid, workspace, FFTW_TRANSPOSED_ORDER).
 *   real space (in):  x-slabs  rho[x - slabstart_x][y][z], rows padded to 2 (N/2+1) doubles (FFTW's in-place r2c layout)
 *   k space (out):    y-slabs  F[y - slabstart_y][x][kz], kz = 0..N/2 complex -- the slab add_nu_power_to_rhogrid takes
 * Slabs are FFTW2's: contiguous, as even as possible, the first ranks take the extra planes.  Unnormalised, like FFTW.
 * 2-D r2c per x plane (cuFFT, loaded with dlopen) -> ONE kernel that transposes x <-> y while it writes every row into
 * the k-space slab of the GPU that owns it, over NVLink peer memory (no pack / all-to-all / unpack passes) -> 1-D c2c
 * along x.  More than one rank needs the peer-memory collective (ksn_comm_p2p_init) on the same ranks: its flag protocol
 * fences the exchange.  Call order: ksn_fft_plan, allocate both buffers with ksn_device_malloc (sizes: ksn_fft_layout),
 * ksn_fft_export on every rank, gather the 128-byte handles in rank order, ksn_fft_attach; then any number of
 * ksn_fft_forward / ksn_fft_inverse (collective, in place: the input buffer is overwritten). */
int ksn_fft_plan(int dims, int nranks, int rank);
int ksn_fft_layout(long long *slabstart_x, long long *nslab_x, long long *slabstart_y, long long *nslab_y,
                   size_t *real_bytes_padded, size_t *kspace_bytes);
int ksn_fft_export(void *d_kspace, void *d_real, void *handles128);
int ksn_fft_attach(void *d_kspace, void *d_real, const void *handles);
int ksn_fft_forward(void *d_real, void *d_kspace);
int ksn_fft_inverse(void *d_kspace, void *d_real);
/* stages of this rank's most recent transform (ms).  Forward: wait for the peers, 2-D pass with the transpose + exchange
 * behind it, closing fence, 1-D pass; inverse: wait, 1-D pass with the exchange behind it, fence, 2-D pass */
int ksn_fft_timing(float *ms4);
void ksn_fft_destroy(void);

/* ---- introspection for bench.py --------------------------------------------------- */
typedef struct ksn_timing {
    float k1_ms, k1_reduce_ms, comm_ms, k2_ms, k3_ms, h2d_ms, d2h_ms;
    unsigned long long launches;   /* kernels launched since ksn_timing_reset */
} ksn_timing;
int ksn_timing_enable(int on);
int ksn_timing_reset(void);
int ksn_timing_get(ksn_timing *out);
/* CUDA stream all kernels are launched on (cudaStream_t as void*). */
void *ksn_stream(void);

#ifdef __cplusplus
}
#endif
#endif
