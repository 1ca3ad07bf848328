// K1 -- power-spectrum bin sums over the slab-decomposed r2c grid (sm_100a).
//
// Replaces hot loop 1 of the reference, powerspectrum.c:56-89, and the four
// MPI_Allreduce calls at powerspectrum.c:91-95.  HBM-bound: 16 B (double grid) read per
// stored mode, nothing written but O(nrbins) partials.
//
// Layout and schedule
//  * The local slab is swept as a flat array of nslab*N*(N/2+1) complex values in
//    warp tiles of 2048 consecutive elements (32 KiB for double): every warp load is one
//    fully coalesced 512-byte request starting on a 512-byte boundary, independent of the
//    odd row length N/2+1.  Tiles are dealt round-robin to CTAs (persistent grid of one
//    CTA per SM), so the sweep is a pure function of (grid size, slab) -> deterministic.
//  * Everything geometric is a function of the integer k^2 = kx^2+ky^2+kz^2.  The bin is
//    estimated with one MUFU log2 and corrected against an integer threshold table that the
//    host built with its own libm from the reference expression
//    floor(binsperunit*log(sqrt(k2))) (powerspectrum.c:40,67,75,83) -> mode counts are
//    bit-exact by construction.  Thresholds and the 1-D inverse CIC window live in shared
//    memory.
//  * Along z the bin index is monotone, so the 32 lanes of a warp hold a few runs of equal
//    bins.  A segmented warp scan (shuffles) sums each run; the run's last lane adds it to
//    the warp's PRIVATE bin array in shared memory -- no atomics, fixed order.
//  * CTA epilogue: warp arrays summed in warp order -> partial[cta][bin] in global memory;
//    a second tiny kernel sums the CTAs in CTA order.  Bit-reproducible run to run.
//  * keff and count sums do not depend on the data; the FULL variant computes them (first
//    call per geometry), the fast variant bins the power only.
#include "ksn_internal.cuh"

#include <math.h>
#include <stdlib.h>
#include <string.h>

namespace ksn {

constexpr int K1_TILE_ITERS = 64;                 // warp iterations per tile
constexpr int K1_TILE = 32 * K1_TILE_ITERS;       // elements per warp tile
constexpr int K1_UNROLL = 8;                      // independent 16-byte loads in flight per lane
constexpr int K1_MAX_WARPS = 16;

template <typename real> struct Cplx;
template <> struct __align__(16) Cplx<double> { double re, im; };
template <> struct __align__(8) Cplx<float> { float re, im; };

// streaming loads: read once, do not pollute L1
__device__ __forceinline__ Cplx<double> ld_stream(const Cplx<double> *p)
{
    Cplx<double> v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.re), "=d"(v.im) : "l"(p));
    return v;
}
__device__ __forceinline__ Cplx<float> ld_stream(const Cplx<float> *p)
{
    Cplx<float> v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.re), "=f"(v.im) : "l"(p));
    return v;
}

// |F|^2 * W with the reference's operation order and (for the float grid) its roundings:
// invwindow() narrows to fftw_real, powerspectrum.c:8-24,68.
__device__ __forceinline__ double mode_power(Cplx<double> v, double wxy, double wz)
{
    const double t = wxy * wz;
    const double w = t * t;
    return (v.re * v.re + v.im * v.im) * (w * w);
}
__device__ __forceinline__ double mode_power(Cplx<float> v, double wxy, double wz)
{
    const float t = (float) wxy * (float) wz;    // wxy already holds float(iwx)*float(iwy)
    const float w = (float) ((double) t * (double) t);
    return (double) (v.re * v.re + v.im * v.im) * ((double) w * (double) w);
}

struct RowState {
    int z, j, c;        // position in the row, row index, ki^2+kj^2
    long long pl;       // plane index relative to plane0
    double wxy;         // iw(ki)*iw(kj)
};

template <typename real>
__device__ __forceinline__ void row_constants(RowState &r, int N, long long plane0, const double *iw_s)
{
    const long long gi = plane0 + r.pl;
    const int ki = gi <= N / 2 ? (int) gi : (int) (gi - N);
    const int kj = r.j <= N / 2 ? r.j : r.j - N;
    r.c = ki * ki + kj * kj;
    const double a = iw_s[ki < 0 ? -ki : ki], b = iw_s[kj < 0 ? -kj : kj];
    if (sizeof(real) == 4) r.wxy = (double) ((float) a * (float) b);
    else r.wxy = a * b;
}

// Segmented inclusive scan over runs of equal `bin` (lanes with the same bin are adjacent).
// On return the LAST lane of each run holds the run's sum.  NV values share the shuffles' control.
template <int NV>
__device__ __forceinline__ unsigned segmented_sum(int bin, double (&v)[NV], int lane)
{
    const unsigned full = 0xffffffffu;
    const int prev = __shfl_up_sync(full, bin, 1);
    const bool head = lane == 0 || bin != prev;
    const unsigned heads = __ballot_sync(full, head);
    const int start = 31 - __clz(heads & (full >> (31 - lane)));   // first lane of my run
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const bool take = lane - d >= start;
#pragma unroll
        for (int i = 0; i < NV; i++) {
            const double up = __shfl_up_sync(full, v[i], d);
            if (take) v[i] += up;
        }
    }
    return (heads >> 1) | 0x80000000u;   // lanes that end a run
}

template <typename real, bool FULL>
__global__ void __launch_bounds__(K1_MAX_WARPS * 32, 1)
k1_bin_kernel(const Cplx<real> *__restrict__ grid, long long nelem, int N, int nrbins, long long plane0,
              float binscale, const unsigned *__restrict__ thr, const double *__restrict__ iw,
              double *__restrict__ partial, int accumulate)
{
    constexpr int NV = FULL ? 3 : 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int L = N / 2 + 1;
    const int nyq = N / 2;
    const int nwarps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *iw_s = (double *) smem_raw;                       // L
    double *bins_s = iw_s + L;                                // nwarps * NV * nrbins
    unsigned *thr_s = (unsigned *) (bins_s + (size_t) nwarps * NV * nrbins);   // nrbins + 1

    for (int i = threadIdx.x; i < L; i += blockDim.x) iw_s[i] = iw[i];
    for (int i = threadIdx.x; i <= nrbins; i += blockDim.x) thr_s[i] = i < nrbins ? thr[i] : 0xffffffffu;
    for (int i = threadIdx.x; i < nwarps * NV * nrbins; i += blockDim.x) bins_s[i] = 0.0;
    __syncthreads();

    double *mybins = bins_s + (size_t) warp * NV * nrbins;
    const long long ntiles = (nelem + K1_TILE - 1) / K1_TILE;
    // tile t -> CTA (t / nwarps) % gridDim.x, warp t % nwarps
    for (long long t = (long long) blockIdx.x * nwarps + warp; t < ntiles; t += (long long) gridDim.x * nwarps) {
        const long long e0 = t * K1_TILE + lane;
        RowState r;
        {
            const long long row = e0 / L;
            r.z = (int) (e0 - row * L);
            r.j = (int) (row % N);
            r.pl = row / N;
            row_constants<real>(r, N, plane0, iw_s);
        }
#pragma unroll 1
        for (int it = 0; it < K1_TILE_ITERS; it += K1_UNROLL) {
            Cplx<real> v[K1_UNROLL];
#pragma unroll
            for (int u = 0; u < K1_UNROLL; u++) {
                const long long e = e0 + (long long) (it + u) * 32;
                if (e < nelem) v[u] = ld_stream(grid + e);
                else { v[u].re = 0; v[u].im = 0; }
            }
#pragma unroll
            for (int u = 0; u < K1_UNROLL; u++) {
                const long long e = e0 + (long long) (it + u) * 32;
                const int k2 = r.c + r.z * r.z;
                const bool live = e < nelem && k2 > 0;
                int b = 0;
                double val[NV];
                if (live) {
                    b = (int) (binscale * __log2f((float) k2));
                    b = max(0, min(b, nrbins - 1));
                    // exact correction against the host-built integer thresholds
                    while (b > 0 && (unsigned) k2 < thr_s[b]) b--;
                    while ((unsigned) k2 >= thr_s[b + 1]) b++;
                    const double mult = (r.z == 0 || r.z == nyq) ? 1.0 : 2.0;
                    val[0] = mode_power(v[u], r.wxy, iw_s[r.z]) * mult;
                    if (FULL) { val[1] = sqrt((double) k2) * mult; val[2] = mult; }
                } else {
                    b = e < nelem ? 0 : 0x7fffffff;
#pragma unroll
                    for (int i = 0; i < NV; i++) val[i] = 0.0;
                }
                const unsigned tails = segmented_sum<NV>(b, val, lane);
                // Bins are monotone along a row, so the runs of one warp step carry distinct bins --
                // unless a new row starts inside this step; then (and only then) fall back to atomics.
                const bool wrapped = __any_sync(0xffffffffu, r.z < lane);
                if (((tails >> lane) & 1u) && b != 0x7fffffff) {
                    if (wrapped) {
#pragma unroll
                        for (int i = 0; i < NV; i++) atomicAdd(&mybins[i * nrbins + b], val[i]);
                    } else {
#pragma unroll
                        for (int i = 0; i < NV; i++) mybins[i * nrbins + b] += val[i];
                    }
                }
                __syncwarp();
                // advance 32 elements along the flat index
                r.z += 32;
                if (r.z >= L) {
                    do {
                        r.z -= L;
                        if (++r.j == N) { r.j = 0; r.pl++; }
                    } while (r.z >= L);
                    row_constants<real>(r, N, plane0, iw_s);
                }
            }
        }
    }
    __syncthreads();
    // CTA epilogue: fixed warp order
    double *out = partial + (size_t) blockIdx.x * NV * nrbins;
    for (int i = threadIdx.x; i < NV * nrbins; i += blockDim.x) {
        double s = 0.0;
        for (int w = 0; w < nwarps; w++) s += bins_s[(size_t) w * NV * nrbins + i];
        out[i] = accumulate ? out[i] + s : s;
    }
}

// Sum the per-CTA partials in CTA order into the reduce buffer
//   red = [ power(nrbins) | total_mass2 | keff(nrbins) | count(nrbins) ]
template <typename real>
__global__ void k1_final_kernel(const double *__restrict__ partial, int ctas, int nv, int nrbins,
                                const Cplx<real> *origin, double *__restrict__ red)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nv * nrbins) {
        double s = 0.0;
        for (int c = 0; c < ctas; c++) s += partial[(size_t) c * nv * nrbins + i];
        const int which = i / nrbins, b = i - which * nrbins;
        red[which == 0 ? b : which * nrbins + 1 + b] = s;
    }
    if (i == 0) {
        double m2 = 0.0;
        if (origin) {   // powerspectrum.c:45-47: only the rank holding plane 0
            const double re = (double) origin->re, im = (double) origin->im;
            m2 = re * re + im * im;
        }
        red[nrbins] = m2;
    }
}

static size_t k1_smem_bytes(int dims, int nrbins, int nwarps, int nv)
{
    return (size_t) (dims / 2 + 1) * 8 + (size_t) nwarps * nv * nrbins * 8 + (size_t) (nrbins + 1) * 4 + 16;
}

template <typename real, bool FULL>
static int k1_launch_t(const void *dgrid, int dims, int nrbins, long long plane0, long long nplanes,
                       bool accumulate, int *ctas_out, int *stride_out)
{
    Ctx &c = ctx();
    constexpr int NV = FULL ? 3 : 1;
    const int L = dims / 2 + 1;
    const long long nelem = nplanes * dims * L;
    int nwarps = K1_MAX_WARPS;
    while (nwarps > 1 && k1_smem_bytes(dims, nrbins, nwarps, NV) > c.smem_optin) nwarps--;
    const size_t smem = k1_smem_bytes(dims, nrbins, nwarps, NV);
    if (smem > c.smem_optin)
        return set_error(KSN_EINVAL, "K1: nrbins=%d with dims=%d needs %zu B of shared memory (> %zu)", nrbins, dims, smem, c.smem_optin);
    const int ctas = c.num_sms;
    int rc = ensure_device_buffer((void **) &c.d_partial, &c.partial_cap, (size_t) ctas * 3 * nrbins * sizeof(double));
    if (rc) return rc;
    const double binsperunit = (nrbins - 1) / log(sqrt(3.0) * dims / 2.0);
    const float binscale = (float) (binsperunit * 0.5 * M_LN2);
    auto launch = [&](auto kern) -> int {
        KSN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        kern<<<ctas, nwarps * 32, smem, c.stream>>>((const Cplx<real> *) dgrid, nelem, dims, nrbins, plane0, binscale,
                                                    c.d_thr, c.d_iw, c.d_partial, accumulate ? 1 : 0);
        c.launches++;
        KSN_CUDA(cudaGetLastError());
        return KSN_OK;
    };
    rc = launch(k1_bin_kernel<real, FULL>);
    if (rc) return rc;
    *ctas_out = ctas;
    *stride_out = NV * nrbins;
    return KSN_OK;
}

int k1_launch(const void *dgrid, int real_bytes, int dims, int nrbins, long long plane0_global, long long nplanes,
              bool full, bool accumulate, int *ctas_out, int *stride_out)
{
    if (real_bytes == 8)
        return full ? k1_launch_t<double, true>(dgrid, dims, nrbins, plane0_global, nplanes, accumulate, ctas_out, stride_out)
                    : k1_launch_t<double, false>(dgrid, dims, nrbins, plane0_global, nplanes, accumulate, ctas_out, stride_out);
    return full ? k1_launch_t<float, true>(dgrid, dims, nrbins, plane0_global, nplanes, accumulate, ctas_out, stride_out)
                : k1_launch_t<float, false>(dgrid, dims, nrbins, plane0_global, nplanes, accumulate, ctas_out, stride_out);
}

int k1_finish(int real_bytes, int dims, int nrbins, bool full, int ctas, int stride, const void *origin_elem)
{
    (void) dims; (void) stride;
    Ctx &c = ctx();
    const int nv = full ? 3 : 1;
    const int threads = 128, blocks = (nv * nrbins + threads - 1) / threads;
    if (real_bytes == 8)
        k1_final_kernel<double><<<blocks, threads, 0, c.stream>>>(c.d_partial, ctas, nv, nrbins, (const Cplx<double> *) origin_elem, c.d_red);
    else
        k1_final_kernel<float><<<blocks, threads, 0, c.stream>>>(c.d_partial, ctas, nv, nrbins, (const Cplx<float> *) origin_elem, c.d_red);
    c.launches++;
    KSN_CUDA(cudaGetLastError());
    return KSN_OK;
}

static int k1_upload_tables(int dims, int nrbins, const unsigned int *thresholds, const double *invwin)
{
    Ctx &c = ctx();
    const int L = dims / 2 + 1;
    int rc = ensure_device_buffer((void **) &c.d_thr, &c.thr_cap, (size_t) nrbins * sizeof(unsigned));
    if (rc) return rc;
    rc = ensure_device_buffer((void **) &c.d_iw, &c.iw_cap, (size_t) L * sizeof(double));
    if (rc) return rc;
    KSN_CUDA(cudaMemcpyAsync(c.d_thr, thresholds, (size_t) nrbins * sizeof(unsigned), cudaMemcpyHostToDevice, c.stream));
    KSN_CUDA(cudaMemcpyAsync(c.d_iw, invwin, (size_t) L * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    rc = ensure_device_buffer((void **) &c.d_red, &c.red_cap, (size_t) (3 * nrbins + 1) * sizeof(double));
    if (rc) return rc;
    rc = ensure_pinned_buffer((void **) &c.h_red, &c.h_red_cap, (size_t) (3 * nrbins + 1) * sizeof(double));
    if (rc) return rc;
    return KSN_OK;
}

// Plane-chunked staging of a host-resident slab into c.d_stage, K1 on each chunk as it lands.
static int k1_over_host_grid(const void *hgrid, int real_bytes, int dims, int nrbins, long long startslab, long long nslab,
                             bool full, int *ctas, int *stride)
{
    Ctx &c = ctx();
    const size_t plane_bytes = (size_t) dims * (dims / 2 + 1) * 2 * real_bytes;
    const size_t total = plane_bytes * (size_t) nslab;
    int rc = ensure_device_buffer(&c.d_stage, &c.stage_cap, total);
    if (rc) return set_error(KSN_ENOMEM, "staging a %zu-byte host grid needs as much free HBM (streaming path not built yet)", total);
    ensure_host_pinned(hgrid, total);
    long long chunk = (long long) ((256ull << 20) / plane_bytes);
    if (chunk < 1) chunk = 1;
    const int nchunks = (int) ((nslab + chunk - 1) / chunk);
    cudaEvent_t *evs = (cudaEvent_t *) malloc(sizeof(cudaEvent_t) * nchunks);
    for (int i = 0; i < nchunks; i++) cudaEventCreateWithFlags(&evs[i], cudaEventDisableTiming);
    // make sure the staging buffer is not still being read by a previous step
    KSN_CUDA(cudaStreamSynchronize(c.stream));
    phase_begin(PH_H2D);
    for (int i = 0; i < nchunks; i++) {
        const long long p0 = i * chunk, np = (p0 + chunk <= nslab) ? chunk : nslab - p0;
        cudaMemcpyAsync((char *) c.d_stage + p0 * plane_bytes, (const char *) hgrid + p0 * plane_bytes,
                        np * plane_bytes, cudaMemcpyHostToDevice, c.copy_stream);
        cudaEventRecord(evs[i], c.copy_stream);
    }
    phase_end(PH_H2D);
    phase_begin(PH_K1);
    for (int i = 0; i < nchunks && !rc; i++) {
        const long long p0 = i * chunk, np = (p0 + chunk <= nslab) ? chunk : nslab - p0;
        cudaStreamWaitEvent(c.stream, evs[i], 0);
        rc = k1_launch((char *) c.d_stage + p0 * plane_bytes, real_bytes, dims, nrbins, startslab + p0, np, full, i > 0, ctas, stride);
    }
    phase_end(PH_K1);
    for (int i = 0; i < nchunks; i++) cudaEventDestroy(evs[i]);
    free(evs);
    if (rc) return rc;
    KSN_CUDA(cudaGetLastError());
    return KSN_OK;
}

// Shared by ksn_powerspectrum_sums and ksn_step_staged.  dgrid_or_null: device-resident slab, else staged from hgrid.
int k1_sums(const void *dgrid, const void *hgrid, int real_bytes, int dims, int nrbins, long long startslab, long long nslab,
            const unsigned int *thresholds, const double *invwin,
            double *power_sum, double *keff_sum, long long *count, double *total_mass2)
{
    Ctx &c = ctx();
    int rc = k1_upload_tables(dims, nrbins, thresholds, invwin);
    if (rc) return rc;
    const bool cache_ok = !getenv("KSN_NO_GEOM_CACHE");
    const bool have_geom = cache_ok && c.geom.valid && c.geom.dims == dims && c.geom.nrbins == nrbins &&
                           c.geom.startslab == startslab && c.geom.nslab == nslab && c.geom.epoch == c.comm_epoch;
    const bool full = !have_geom;
    int ctas = 0, stride = 0;
    const void *origin = nullptr;
    if (nslab > 0) {
        if (dgrid) {
            phase_begin(PH_K1);
            rc = k1_launch(dgrid, real_bytes, dims, nrbins, startslab, nslab, full, false, &ctas, &stride);
            phase_end(PH_K1);
            origin = dgrid;
        } else {
            rc = k1_over_host_grid(hgrid, real_bytes, dims, nrbins, startslab, nslab, full, &ctas, &stride);
            origin = c.d_stage;
        }
        if (rc) return rc;
    } else {
        // a rank that owns no planes still takes part in the collective with zeros
        rc = ensure_device_buffer((void **) &c.d_partial, &c.partial_cap, (size_t) 3 * nrbins * sizeof(double));
        if (rc) return rc;
        KSN_CUDA(cudaMemsetAsync(c.d_partial, 0, (size_t) 3 * nrbins * sizeof(double), c.stream));
        ctas = 1;
    }
    phase_begin(PH_K1RED);
    rc = k1_finish(real_bytes, dims, nrbins, full, ctas, stride, (startslab == 0 && nslab > 0) ? origin : nullptr);
    phase_end(PH_K1RED);
    if (rc) return rc;
    const size_t nred = full ? (size_t) 3 * nrbins + 1 : (size_t) nrbins + 1;
    rc = allreduce_to_host(c.d_red, c.h_red, nred);
    if (rc) return rc;
    phase_collect();
    if (full) {
        if (c.geom.cap < (size_t) nrbins) {
            free(c.geom.keff); free(c.geom.count);
            c.geom.keff = (double *) malloc(sizeof(double) * nrbins);
            c.geom.count = (long long *) malloc(sizeof(long long) * nrbins);
            c.geom.cap = nrbins;
        }
        for (int b = 0; b < nrbins; b++) {
            c.geom.keff[b] = c.h_red[nrbins + 1 + b];
            c.geom.count[b] = (long long) llrint(c.h_red[2 * nrbins + 1 + b]);   // exact: < 2^53
        }
        c.geom.valid = true; c.geom.dims = dims; c.geom.nrbins = nrbins;
        c.geom.startslab = startslab; c.geom.nslab = nslab; c.geom.epoch = c.comm_epoch;
    }
    memcpy(power_sum, c.h_red, sizeof(double) * nrbins);
    *total_mass2 = c.h_red[nrbins];
    memcpy(keff_sum, c.geom.keff, sizeof(double) * nrbins);
    memcpy(count, c.geom.count, sizeof(long long) * nrbins);
    return KSN_OK;
}

}  // namespace ksn

using namespace ksn;

extern "C" int ksn_powerspectrum_sums(const void *grid, int real_bytes, int dims, int nrbins,
                                      long long startslab, long long nslab,
                                      const unsigned int *thresholds, const double *invwin,
                                      double *power_sum, double *keff_sum, long long *count, double *total_mass2)
{
    int rc = ensure_init();
    if (rc) return rc;
    if ((!grid && nslab > 0) || (real_bytes != 4 && real_bytes != 8) || dims < 2 || (dims & 1) || nrbins < 2 || nslab < 0 ||
        startslab < 0 || startslab + nslab > dims || !thresholds || !invwin || !power_sum || !keff_sum || !count || !total_mass2)
        return set_error(KSN_EINVAL, "ksn_powerspectrum_sums: bad arguments (dims=%d nrbins=%d slab=[%lld,+%lld))", dims, nrbins, startslab, nslab);
    if ((double) dims * dims * 0.75 > 2.0e9) return set_error(KSN_EINVAL, "dims=%d: k^2 does not fit 31 bits", dims);
    const bool on_device = nslab > 0 && ksn_pointer_is_device(grid);
    return k1_sums(on_device ? grid : nullptr, on_device ? nullptr : grid, real_bytes, dims, nrbins, startslab, nslab,
                   thresholds, invwin, power_sum, keff_sum, count, total_mass2);
}
