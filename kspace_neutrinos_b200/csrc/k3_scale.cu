// K3 -- multiply every stored Fourier mode by 1 + norm * interp(log k)  (sm_100a).
//
// Replaces hot loop 3 of the reference, interface_gadget.c:163-188, including
// get_dnudcdm_powerspec (delta_pow.c:19-37: clamp log k into the table, piecewise-linear
// interpolation, times norm).  HBM-bound read-modify-write: 16 B in + 16 B out per mode
// for a double grid.
//
// The factor depends only on the integer k^2.  Per segment i between table knots the host
// precomputes, in units of the grid's integer wave numbers,
//     K2_i   = (exp(logkk_i) * box / 2pi)^2          knot position in k^2
//     A_i    = 1 + norm * ratio_i
//     B_i    = norm * (ratio_{i+1}-ratio_i) / (2 (logkk_{i+1}-logkk_i))      (B_last = 0)
// so that   smth(k2) = A_i + B_i * ln(k2 / K2_i)     for K2_i <= k2 < K2_{i+1},
// identical to the reference's 1 + norm*(y_lo + (x-x_lo)/dx*dy) with x = log(sqrt(k2) 2pi/box).
// ln(k2/K2_i) = log1p(u), u = k2/K2_i - 1; bins are narrow (u < 2%) for all but a handful of
// low-k modes, so a degree-8 series is exact to < 1e-15 there and the slow log1p() is taken
// only when u is large.  The segment is found from a 4096-cell lookup in log2(k2) (one MUFU
// log2) followed by a short walk, all in shared memory.
#include "ksn_internal.cuh"

#include <math.h>
#include <stdlib.h>

namespace ksn {

constexpr int K3_CELLS = 4096;
constexpr int K3_TILE_ITERS = 32;
constexpr int K3_TILE = 32 * K3_TILE_ITERS;
constexpr int K3_UNROLL = 8;
constexpr int K3_THREADS = 512;

template <typename real> struct C2;
template <> struct __align__(16) C2<double> { double re, im; };
template <> struct __align__(8) C2<float> { float re, im; };

__device__ __forceinline__ C2<double> ld_cs(const C2<double> *p)
{
    C2<double> v;
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.re), "=d"(v.im) : "l"(p));
    return v;
}
__device__ __forceinline__ C2<float> ld_cs(const C2<float> *p)
{
    C2<float> v;
    asm volatile("ld.global.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.re), "=f"(v.im) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_cs(C2<double> *p, C2<double> v)
{
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.re), "d"(v.im) : "memory");
}
__device__ __forceinline__ void st_cs(C2<float> *p, C2<float> v)
{
    asm volatile("st.global.cs.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.re), "f"(v.im) : "memory");
}

// table layout in global/shared memory (doubles): K2[n] | invK2[n] | A[n] | B[n] | then ushort cell[K3_CELLS]
struct K3Params {
    int n;            // knots
    float cell_lo;    // log2(K2_0)
    float cell_scale; // cells per unit log2
};

template <typename real>
__global__ void __launch_bounds__(K3_THREADS, 1)
k3_scale_kernel(C2<real> *__restrict__ grid, long long nelem, int N, long long plane0,
                const double *__restrict__ tab, K3Params prm)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = prm.n;
    double *K2_s = (double *) smem_raw;
    double *inv_s = K2_s + n;
    double *A_s = inv_s + n;
    double *B_s = A_s + n;
    unsigned short *cell_s = (unsigned short *) (B_s + n);
    for (int i = threadIdx.x; i < 4 * n; i += blockDim.x) K2_s[i] = tab[i];
    const unsigned short *cell_g = (const unsigned short *) (tab + 4 * n);
    for (int i = threadIdx.x; i < K3_CELLS; i += blockDim.x) cell_s[i] = cell_g[i];
    __syncthreads();

    const int L = N / 2 + 1;
    const int nwarps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long ntiles = (nelem + K3_TILE - 1) / K3_TILE;
    for (long long t = (long long) blockIdx.x * nwarps + warp; t < ntiles; t += (long long) gridDim.x * nwarps) {
        const long long e0 = t * K3_TILE + lane;
        int z, j, c;
        long long pl;
        {
            const long long row = e0 / L;
            z = (int) (e0 - row * L);
            j = (int) (row % N);
            pl = row / N;
        }
        auto rowc = [&]() {
            const long long gi = plane0 + pl;
            const int ki = gi <= N / 2 ? (int) gi : (int) (gi - N);
            const int kj = j <= N / 2 ? j : j - N;
            c = ki * ki + kj * kj;
        };
        rowc();
#pragma unroll 1
        for (int it = 0; it < K3_TILE_ITERS; it += K3_UNROLL) {
            C2<real> v[K3_UNROLL];
#pragma unroll
            for (int u = 0; u < K3_UNROLL; u++) {
                const long long e = e0 + (long long) (it + u) * 32;
                if (e < nelem) v[u] = ld_cs(grid + e);
            }
#pragma unroll
            for (int u = 0; u < K3_UNROLL; u++) {
                const long long e = e0 + (long long) (it + u) * 32;
                const int k2i = c + z * z;
                if (e < nelem && k2i > 0) {
                    const double k2 = (double) k2i;
                    int cell = (int) ((__log2f((float) k2i) - prm.cell_lo) * prm.cell_scale);
                    cell = max(0, min(cell, K3_CELLS - 1));
                    int s = cell_s[cell];
                    while (s > 0 && k2 < K2_s[s]) s--;
                    while (s + 1 < n && k2 >= K2_s[s + 1]) s++;
                    double uu = fma(k2, inv_s[s], -1.0);
                    if (uu < 0.0) uu = 0.0;                  // below the first knot: clamp (delta_pow.c:24-31)
                    double lg;
                    if (uu < 0.03125) {
                        // log1p(u), |u| < 2^-5: alternating series to u^9, error < 3e-16
                        lg = uu * (1.0 + uu * (-0.5 + uu * (1.0 / 3 + uu * (-0.25 + uu * (0.2 + uu * (-1.0 / 6 + uu * (1.0 / 7 + uu * (-0.125 + uu * (1.0 / 9)))))))));
                    } else {
                        lg = log1p(uu);
                    }
                    const double smth = fma(B_s[s], lg, A_s[s]);
                    C2<real> o;
                    o.re = (real) ((double) v[u].re * smth);
                    o.im = (real) ((double) v[u].im * smth);
                    st_cs(grid + e, o);
                }
                z += 32;
                if (z >= L) {
                    do {
                        z -= L;
                        if (++j == N) { j = 0; pl++; }
                    } while (z >= L);
                    rowc();
                }
            }
        }
    }
}

static size_t k3_tab_doubles(int n) { return (size_t) 4 * n + (K3_CELLS * sizeof(unsigned short) + 7) / 8; }
static K3Params g_k3prm;

int k3_upload_table(int dims, double boxsize, const double *logkk, const double *ratio, int nbins, double norm)
{
    (void) dims;
    Ctx &c = ctx();
    const size_t nd = k3_tab_doubles(nbins);
    int rc = ensure_device_buffer((void **) &c.d_k3tab, &c.k3tab_cap, nd * sizeof(double));
    if (rc) return rc;
    rc = ensure_pinned_buffer((void **) &c.h_k3tab, &c.h_k3tab_cap, nd * sizeof(double));
    if (rc) return rc;
    // the pinned table may still be in flight from the previous step
    KSN_CUDA(cudaStreamSynchronize(c.stream));
    double *K2 = c.h_k3tab, *inv = K2 + nbins, *A = inv + nbins, *B = A + nbins;
    unsigned short *cell = (unsigned short *) (B + nbins);
    const double unit = boxsize / (2 * M_PI);
    for (int i = 0; i < nbins; i++) {
        const double kg = exp(logkk[i]) * unit;     // knot in integer-wave-number units
        K2[i] = kg * kg;
        inv[i] = 1.0 / K2[i];
        A[i] = 1.0 + norm * ratio[i];
        B[i] = (i + 1 < nbins) ? norm * (ratio[i + 1] - ratio[i]) / (2.0 * (logkk[i + 1] - logkk[i])) : 0.0;
    }
    const double lo = log2(K2[0]), hi = log2(K2[nbins - 1]);
    const double scale = hi > lo ? K3_CELLS / (hi - lo) : 0.0;
    int s = 0;
    for (int k = 0; k < K3_CELLS; k++) {
        // last knot at or below the lower edge of cell k (minus a float-rounding guard)
        const double edge = exp2(lo + (k - 0.01) / (scale > 0 ? scale : 1.0));
        while (s + 1 < nbins && K2[s + 1] <= edge) s++;
        cell[k] = (unsigned short) s;
    }
    g_k3prm.n = nbins;
    g_k3prm.cell_lo = (float) lo;
    g_k3prm.cell_scale = (float) scale;
    KSN_CUDA(cudaMemcpyAsync(c.d_k3tab, c.h_k3tab, nd * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    return KSN_OK;
}

int k3_launch(void *dgrid, int real_bytes, int dims, long long plane0_global, long long nplanes, int nknots)
{
    Ctx &c = ctx();
    const long long nelem = nplanes * dims * (dims / 2 + 1);
    if (nelem == 0) return KSN_OK;
    const size_t smem = k3_tab_doubles(nknots) * sizeof(double) + 16;
    if (smem > c.smem_optin) return set_error(KSN_EINVAL, "K3: %d knots need %zu B of shared memory", nknots, smem);
    // two resident CTAs per SM when the table is small enough
    const int per_sm = (2 * (smem + 1024) <= c.smem_optin) ? 2 : 1;
    const long long ntiles = (nelem + K3_TILE - 1) / K3_TILE;
    long long want = (ntiles + (K3_THREADS / 32) - 1) / (K3_THREADS / 32);
    int ctas = (int) (want < (long long) c.num_sms * per_sm ? want : (long long) c.num_sms * per_sm);
    if (ctas < 1) ctas = 1;
    if (real_bytes == 8) {
        KSN_CUDA(cudaFuncSetAttribute(k3_scale_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        k3_scale_kernel<double><<<ctas, K3_THREADS, smem, c.stream>>>((C2<double> *) dgrid, nelem, dims, plane0_global, c.d_k3tab, g_k3prm);
    } else {
        KSN_CUDA(cudaFuncSetAttribute(k3_scale_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        k3_scale_kernel<float><<<ctas, K3_THREADS, smem, c.stream>>>((C2<float> *) dgrid, nelem, dims, plane0_global, c.d_k3tab, g_k3prm);
    }
    c.launches++;
    KSN_CUDA(cudaGetLastError());
    return KSN_OK;
}

// K3 over a slab that already sits in c.d_stage (uploaded by K1's staged path), chunk by chunk, each
// chunk copied back to the host buffer as soon as it is scaled.
int k3_over_staged_grid(void *hgrid, int real_bytes, int dims, long long startslab, long long nslab, int nknots)
{
    Ctx &c = ctx();
    const size_t plane_bytes = (size_t) dims * (dims / 2 + 1) * 2 * real_bytes;
    long long chunk = (long long) ((256ull << 20) / plane_bytes);
    if (chunk < 1) chunk = 1;
    const int nchunks = (int) ((nslab + chunk - 1) / chunk);
    cudaEvent_t *evs = (cudaEvent_t *) malloc(sizeof(cudaEvent_t) * nchunks);
    for (int i = 0; i < nchunks; i++) cudaEventCreateWithFlags(&evs[i], cudaEventDisableTiming);
    int rc = KSN_OK;
    phase_begin(PH_K3);
    for (int i = 0; i < nchunks && !rc; i++) {
        const long long p0 = i * chunk, np = (p0 + chunk <= nslab) ? chunk : nslab - p0;
        rc = k3_launch((char *) c.d_stage + p0 * plane_bytes, real_bytes, dims, startslab + p0, np, nknots);
        cudaEventRecord(evs[i], c.stream);
    }
    phase_end(PH_K3);
    phase_begin(PH_D2H);
    for (int i = 0; i < nchunks && !rc; i++) {
        const long long p0 = i * chunk, np = (p0 + chunk <= nslab) ? chunk : nslab - p0;
        cudaStreamWaitEvent(c.copy_stream, evs[i], 0);
        cudaMemcpyAsync((char *) hgrid + p0 * plane_bytes, (char *) c.d_stage + p0 * plane_bytes, np * plane_bytes,
                        cudaMemcpyDeviceToHost, c.copy_stream);
    }
    phase_end(PH_D2H);
    cudaError_t e1 = cudaStreamSynchronize(c.copy_stream), e2 = cudaStreamSynchronize(c.stream);
    for (int i = 0; i < nchunks; i++) cudaEventDestroy(evs[i]);
    free(evs);
    if (rc) return rc;
    KSN_CUDA(e1);
    KSN_CUDA(e2);
    phase_collect();
    return KSN_OK;
}

int k1_sums(const void *dgrid, const void *hgrid, int real_bytes, int dims, int nrbins, long long startslab, long long nslab,
            const unsigned int *thresholds, const double *invwin,
            double *power_sum, double *keff_sum, long long *count, double *total_mass2);

}  // namespace ksn

using namespace ksn;

static int check_table(const double *logkk, const double *ratio, int nbins, double norm)
{
    if (!logkk || !ratio || nbins < 2 || nbins > 65535) return set_error(KSN_EINVAL, "K3: bad table (nbins=%d)", nbins);
    for (int i = 1; i < nbins; i++)
        if (!(logkk[i] > logkk[i - 1])) return set_error(KSN_EINVAL, "K3: logkk must increase strictly (i=%d)", i);
    (void) norm;
    return KSN_OK;
}

extern "C" int ksn_scale_modes(void *grid, int real_bytes, int dims, long long startslab, long long nslab,
                               double boxsize, const double *logkk, const double *ratio, int nbins, double norm)
{
    int rc = ensure_init();
    if (rc) return rc;
    if ((!grid && nslab > 0) || (real_bytes != 4 && real_bytes != 8) || dims < 2 || (dims & 1) || nslab < 0 || startslab < 0 ||
        startslab + nslab > dims || !(boxsize > 0))
        return set_error(KSN_EINVAL, "ksn_scale_modes: bad arguments");
    rc = check_table(logkk, ratio, nbins, norm);
    if (rc) return rc;
    if (nslab == 0) return KSN_OK;
    Ctx &c = ctx();
    rc = k3_upload_table(dims, boxsize, logkk, ratio, nbins, norm);
    if (rc) return rc;
    if (ksn_pointer_is_device(grid)) {
        phase_begin(PH_K3);
        rc = k3_launch(grid, real_bytes, dims, startslab, nslab, nbins);
        phase_end(PH_K3);
        if (rc) return rc;
        KSN_CUDA(cudaStreamSynchronize(c.stream));
        phase_collect();
        return KSN_OK;
    }
    // host grid: upload, scale, download
    const size_t total = (size_t) nslab * dims * (dims / 2 + 1) * 2 * real_bytes;
    rc = ensure_device_buffer(&c.d_stage, &c.stage_cap, total);
    if (rc) return set_error(KSN_ENOMEM, "staging a %zu-byte host grid needs as much free HBM", total);
    ensure_host_pinned(grid, total);
    phase_begin(PH_H2D);
    KSN_CUDA(cudaMemcpyAsync(c.d_stage, grid, total, cudaMemcpyHostToDevice, c.copy_stream));
    phase_end(PH_H2D);
    KSN_CUDA(cudaStreamSynchronize(c.copy_stream));
    return k3_over_staged_grid(grid, real_bytes, dims, startslab, nslab, nbins);
}

extern "C" int ksn_step_staged(void *hgrid, int real_bytes, int dims, int nrbins, long long startslab, long long nslab,
                               const unsigned int *thresholds, const double *invwin, double boxsize,
                               ksn_between_fn between, void *user)
{
    int rc = ensure_init();
    if (rc) return rc;
    if ((!hgrid && nslab > 0) || !between || !thresholds || !invwin || (real_bytes != 4 && real_bytes != 8) || dims < 2 || (dims & 1) ||
        nrbins < 2 || nslab < 0 || startslab < 0 || startslab + nslab > dims || !(boxsize > 0))
        return set_error(KSN_EINVAL, "ksn_step_staged: bad arguments");
    Ctx &c = ctx();
    const bool on_device = nslab > 0 && ksn_pointer_is_device(hgrid);
    // geometry sums come back in temporaries sized nrbins
    double *power = (double *) malloc(sizeof(double) * nrbins * 2);
    long long *count = (long long *) malloc(sizeof(long long) * nrbins);
    double *keff = power + nrbins, mass2 = 0;
    rc = k1_sums(on_device ? hgrid : nullptr, on_device ? nullptr : hgrid, real_bytes, dims, nrbins, startslab, nslab,
                 thresholds, invwin, power, keff, count, &mass2);
    const double *logkk = nullptr, *ratio = nullptr;
    int nbins = 0;
    double norm = 0;
    if (!rc) {
        int brc = between(user, power, keff, count, mass2, &logkk, &ratio, &nbins, &norm);
        if (brc) rc = set_error(KSN_EINVAL, "ksn_step_staged: callback failed (%d)", brc);
    }
    free(power);
    free(count);
    if (rc) return rc;
    rc = check_table(logkk, ratio, nbins, norm);
    if (rc) return rc;
    if (nslab == 0) return KSN_OK;
    rc = k3_upload_table(dims, boxsize, logkk, ratio, nbins, norm);
    if (rc) return rc;
    if (on_device) {
        phase_begin(PH_K3);
        rc = k3_launch(hgrid, real_bytes, dims, startslab, nslab, nbins);
        phase_end(PH_K3);
        if (rc) return rc;
        KSN_CUDA(cudaStreamSynchronize(c.stream));
        phase_collect();
        return KSN_OK;
    }
    return k3_over_staged_grid(hgrid, real_bytes, dims, startslab, nslab, nbins);
}
