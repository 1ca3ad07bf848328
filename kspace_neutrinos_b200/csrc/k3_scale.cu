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
// low-k modes, so a short series is exact to < 1e-16 there and the slow log1p() is taken
// only when u is large.  The segment comes from a cell lookup in log2(k2) (one MUFU log2, cells
// sized by the host so that no cell holds two knots) plus one compare, all in shared memory.
// Sweep: one warp per row, 32 consecutive z per step, rows dealt round-robin to a persistent grid.
#include "ksn_internal.cuh"
#include <vector>

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <type_traits>

namespace ksn {

constexpr int K3_MAX_CELLS = 16384;
constexpr int K3_THREADS = 256;

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

__device__ __forceinline__ float fast_log2(float x)
{
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// per-knot record
struct __align__(16) K3Seg { double K2, inv, A, B; };
// the same in float, for float grids whose table is smooth enough (see k3_upload_table): delta = A - 1 = norm * ratio
struct __align__(16) K3SegF { float inv, A1, B, pad; };

// Optional second factor: the periodic PM Green's function with CIC deconvolution that a Gadget-style PM step applies
// to the same grid right after the neutrino correction (Gadget-2 pm_periodic.c, the loop that follows the hook of
// gadget-2/0002 patch:116-125):  G(k) = -exp(-k2*asmth2)/k2 * (iwx iwy iwz)^4,  G(0) = 0,  iw(q) = 1/sinc(pi q/N).
// Fusing it into K3 saves the host's own read-modify-write pass over the grid (32 B per mode).
struct K3Greens {
    int on;
    double asmth2;
    const double *iw;   // device: 1-D inverse CIC window, q = 0..N/2 (the table K1 uses)
    const double *gz;   // device: exp(-z^2 asmth2) * iw[z]^4, z = 0..N/2
};

// exp(-k2 a) = exp(-(kx^2+ky^2) a) * exp(-kz^2 a): a per-row scalar (with the x-y window) times a per-z table, so that a
// mode costs one table read, two multiplies and a reciprocal of the integer k2 (correctly rounded float seed, relative
// error 2^-24, + one Newton step: 2^-48 = 3.6e-15; FP64 issue is what the fused pass is short of).
__device__ __forceinline__ double k3_greens_row(const K3Greens &gr, int ki, int kj)
{
    const double a = __ldg(gr.iw + (ki < 0 ? -ki : ki)), b = __ldg(gr.iw + (kj < 0 ? -kj : kj));
    const double q = a * b, q2 = q * q;
    return -exp(-(double) (ki * ki + kj * kj) * gr.asmth2) * (q2 * q2);
}

__device__ __forceinline__ double k3_greens(const K3Greens &gr, int k2i, double rowfac, int z)
{
    const double k2 = (double) k2i;
    double r = (double) __frcp_rn((float) k2i);
    r = r * fma(-k2, r, 2.0);
    return rowfac * __ldg(gr.gz + z) * r;
}

struct K3Params {
    int n;            // knots
    int cells;        // lookup cells in log2(k2); the host sizes them so that no cell holds two knots
    float cell_lo;    // log2(K2_0)
    float cell_scale; // cells per unit log2
    int off_kthr;     // table offsets in units of 4 bytes from the table base: integer knot thresholds, float records
    int off_segf;
    int multi;        // a cell may hold several knots (knots closer than the finest cells): loop in the segment search
    unsigned k2_narrow; // from this k2 on every segment is narrow (u < 2^-5 throughout) and at or above the first knot:
                        // a grid row with kx^2 + ky^2 >= k2_narrow needs no clamp, no log1p and no bounds on the cell index
    float cell_off;   // -cell_lo * cell_scale
};

// How the factor 1 + norm*interp is evaluated (chosen per table by the host, k3_upload_table):
//   FM_D9   double, ln(1+u) as a series to u^9 (truncation < 1e-16 for u < 2^-5): any table
//   FM_D5   double, series to u^5: tables whose narrow segments satisfy |B| u_max^6 / 6 <= 1e-14 -- four FP64
//           instructions less per mode, the same factor to 1e-14
//   FM_F32  FLOAT grids only: delta = factor - 1 entirely in float and x*factor as fmaf(x, delta, x), for tables with
//           |B| (1 + 3 u) + |norm*ratio| <= 1/8, where the float evaluation of delta is off by < 2^-26 of the factor
//           (a quarter of a float ulp of the product) -- no FP64 instruction left in the float pass
enum { FM_D9 = 0, FM_D5 = 1, FM_F32 = 2 };

// The lookup tables live in GLOBAL memory and are read through L1 (__ldg): they are identical for every CTA, a few
// tens of KB, and the streaming grid accesses bypass L1 (L1::no_allocate), so after the first CTAs of a launch every
// lookup is an L1 hit -- without the per-CTA staging cost that would forbid short-lived CTAs.
// Segment of k2: a cell lookup in log2(k2) (one MUFU) gives the last knot at or below the cell's lower edge; the cell
// holds at most one more knot (host guarantee), found by one INTEGER compare -- k2 >= K2_{s+1} <=> k2 >= ceil(K2_{s+1})
// for integer k2.
__device__ __forceinline__ int k3_segment(int k2i, const unsigned short *__restrict__ cellv, const unsigned *__restrict__ kthr, const K3Params &prm)
{
    int cell = (int) ((fast_log2((float) k2i) - prm.cell_lo) * prm.cell_scale);
    cell = max(0, min(cell, prm.cells - 1));
    int s = __ldg(cellv + cell);
    s += ((unsigned) k2i >= __ldg(kthr + s + 1));
    if (prm.multi) while ((unsigned) k2i >= __ldg(kthr + s + 1)) s++;        // (kthr ends in 0xffffffff: terminates)
    return s;
}

template <int FM>
__device__ __forceinline__ double k3_factor(int k2i, const double *__restrict__ tab, const K3Params &prm)
{
    const K3Seg *seg = (const K3Seg *) tab;
    const unsigned short *cellv = (const unsigned short *) (seg + prm.n + 1);
    const unsigned *kthr = (const unsigned *) tab + prm.off_kthr;
    const int s = k3_segment(k2i, cellv, kthr, prm);
    const double2 ki = __ldg((const double2 *) &seg[s].K2);    // {K2, 1/K2}
    const double2 ab = __ldg((const double2 *) &seg[s].A);     // {A, B}
    const double uu = fmax(fma((double) k2i, ki.y, -1.0), 0.0); // below the first knot: clamp (delta_pow.c:24-31)
    double lg;
    if (uu < 0.03125) {
        // log1p(u), u < 2^-5: alternating series
        if (FM == FM_D5) lg = uu * (1.0 + uu * (-0.5 + uu * (1.0 / 3 + uu * (-0.25 + uu * 0.2))));
        else lg = uu * (1.0 + uu * (-0.5 + uu * (1.0 / 3 + uu * (-0.25 + uu * (0.2 + uu * (-1.0 / 6 + uu * (1.0 / 7 + uu * (-0.125 + uu * (1.0 / 9)))))))));
    } else {
        lg = log1p(uu);
    }
    return fma(ab.y, lg, ab.x);
}

// The same factor for a mode the host has vouched for (k2 >= prm.k2_narrow): no clamp below the first knot, no wide
// segment (so no log1p), the cell index needs no bounds -- a straight line of code, so that the compiler can interleave
// the dependent FP64 chains of a thread's nine modes.
template <int FM>
__device__ __forceinline__ double k3_factor_narrow(int k2i, const double *__restrict__ tab, const K3Params &prm)
{
    const K3Seg *seg = (const K3Seg *) tab;
    const unsigned short *cellv = (const unsigned short *) (seg + prm.n + 1);
    const unsigned *kthr = (const unsigned *) tab + prm.off_kthr;
    const int cell = (int) fmaf(fast_log2((float) k2i), prm.cell_scale, prm.cell_off);
    int s = __ldg(cellv + cell);
    s += ((unsigned) k2i >= __ldg(kthr + s + 1));
    const double inv = __ldg(&seg[s].inv);
    const double2 ab = __ldg((const double2 *) &seg[s].A);     // {A, B}
    const double uu = fma((double) k2i, inv, -1.0);
    double lg;
    if (FM == FM_D5) lg = uu * (1.0 + uu * (-0.5 + uu * (1.0 / 3 + uu * (-0.25 + uu * 0.2))));
    else lg = uu * (1.0 + uu * (-0.5 + uu * (1.0 / 3 + uu * (-0.25 + uu * (0.2 + uu * (-1.0 / 6 + uu * (1.0 / 7 + uu * (-0.125 + uu * (1.0 / 9)))))))));
    return fma(ab.y, lg, ab.x);
}

// FM_F32: factor - 1 in float
__device__ __forceinline__ float k3_delta_f32(int k2i, const double *__restrict__ tab, const K3Params &prm)
{
    const K3Seg *seg = (const K3Seg *) tab;
    const unsigned short *cellv = (const unsigned short *) (seg + prm.n + 1);
    const unsigned *kthr = (const unsigned *) tab + prm.off_kthr;
    const K3SegF *segf = (const K3SegF *) ((const unsigned *) tab + prm.off_segf);
    const int s = k3_segment(k2i, cellv, kthr, prm);
    const float4 r = __ldg((const float4 *) (segf + s));       // {1/K2, A - 1, B, -}
    const float uu = fmaxf(fmaf((float) k2i, r.x, -1.0f), 0.0f);  // k2 < 2^24 is exact in float
    float lg;
    if (uu < 0.03125f) lg = uu * (1.0f + uu * (-0.5f + uu * (1.0f / 3 + uu * (-0.25f + uu * 0.2f))));
    else lg = log1pf(uu);
    return fmaf(r.z, lg, r.y);
}

__device__ __forceinline__ float k3_delta_f32_narrow(int k2i, const double *__restrict__ tab, const K3Params &prm)
{
    const K3Seg *seg = (const K3Seg *) tab;
    const unsigned short *cellv = (const unsigned short *) (seg + prm.n + 1);
    const unsigned *kthr = (const unsigned *) tab + prm.off_kthr;
    const K3SegF *segf = (const K3SegF *) ((const unsigned *) tab + prm.off_segf);
    const int cell = (int) fmaf(fast_log2((float) k2i), prm.cell_scale, prm.cell_off);
    int s = __ldg(cellv + cell);
    s += ((unsigned) k2i >= __ldg(kthr + s + 1));
    const float4 r = __ldg((const float4 *) (segf + s));
    const float uu = fmaf((float) k2i, r.x, -1.0f);
    const float lg = uu * (1.0f + uu * (-0.5f + uu * (1.0f / 3 + uu * (-0.25f + uu * 0.2f))));
    return fmaf(r.z, lg, r.y);
}

// one mode times its factor, in the grid's precision rules: fftw_real *= double (interface_gadget.c:185-186)
template <typename real> __device__ __forceinline__ void k3_apply(C2<real> &v, double smth)
{
    v.re = (real) ((double) v.re * smth);
    v.im = (real) ((double) v.im * smth);
}
__device__ __forceinline__ void k3_apply_delta(C2<float> &v, float d)
{
    v.re = fmaf(v.re, d, v.re);
    v.im = fmaf(v.im, d, v.im);
}

// Sweep: SHORT-LIVED CTAs, one per block of `rows_per_cta` consecutive rows (one row of N/2+1 modes for large grids):
// every thread issues all its loads up front, the CTA's accesses form one contiguous range, and the hardware block
// scheduler balances the SMs.  Measured on this B200 (tools/bw_probe.cu) this pattern sustains 6.9 TB/s for an
// in-place scale against 6.1 TB/s for a persistent loop.  Plain loads and stores: the fall-back for slabs the bulk-copy
// kernels below cannot take (a base pointer off the 16-byte granule, an odd mode total of a float slab) and the
// comparison kernel of the tests (KSN_K3_NOTMA).
template <typename real, int U, int FM>
__global__ void __launch_bounds__(K3_THREADS)
k3_scale_kernel(C2<real> *__restrict__ grid, int nrows, int rows_per_cta, int N, long long plane0,
                const double *__restrict__ tab, const K3Params prm, const K3Greens gr)
{
    const int L = N / 2 + 1;
    const int row0 = blockIdx.x * rows_per_cta;
    const int nr = min(rows_per_cta, nrows - row0);
    const int nel = nr * L;                                      // modes in this CTA's block (contiguous)
    C2<real> *base = grid + (size_t) row0 * L;
    for (int e0 = threadIdx.x; e0 < nel; e0 += K3_THREADS * U) {
        C2<real> v[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int e = e0 + K3_THREADS * u;
            if (e < nel) v[u] = ld_cs(base + e);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int e = e0 + K3_THREADS * u;
            if (e >= nel) continue;
            const int rl = rows_per_cta == 1 ? 0 : e / L;
            const int z = e - rl * L;
            const int r = row0 + rl;
            const int pl = r / N, j = r - pl * N;
            const long long gi = plane0 + pl;
            const int ki = gi <= N / 2 ? (int) gi : (int) (gi - N);
            const int kj = j <= N / 2 ? j : j - N;
            const int k2i = ki * ki + kj * kj + z * z;
            if (k2i == 0 && !gr.on) continue;                    // F(0,0,0) is skipped (interface_gadget.c:174)
            C2<real> o = v[u];
            if constexpr (FM == FM_F32) {
                k3_apply_delta(o, k3_delta_f32(k2i, tab, prm));          // (never launched with the Green's function on)
            } else {
                double smth = k2i > 0 ? k3_factor<FM>(k2i, tab, prm) : 0.0;
                if (gr.on && k2i > 0) smth *= k3_greens(gr, k2i, k3_greens_row(gr, ki, kj), z);   // the mean is zeroed
                k3_apply<real>(o, smth);
            }
            st_cs(base + e, o);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Bulk-copy kernels (default): the CTA's piece of the grid is brought into shared memory by ONE bulk asynchronous copy
// (cp.async.bulk, completion on an mbarrier) issued before anything else; while it is in flight every thread computes
// the scale factors of its modes (they depend on geometry only), then scales in shared memory and one bulk store
// writes the piece back.  In-flight bytes are bounded by shared memory (~16 KB per CTA, 7-8 CTAs per SM), not by
// registers.
constexpr int K3_EPT = 9;                  // modes per thread
constexpr int K3_TMA_THREADS = 128;        // <= 1152 modes (one 2048^3 row) per CTA: many small CTAs keep L1 for the tables

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned) __cvta_generic_to_shared(p); }

// the three phases shared by the bulk-copy kernels
__device__ __forceinline__ void k3_bulk_load(void *buf, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(buf)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void k3_bulk_wait(unsigned long long *bar)
{
    unsigned done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void k3_bulk_store(void *dst, const void *buf, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(buf)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory must outlive the read
}

// One grid ROW per CTA (PMGRID 1152 ... 2302), or -- SPLIT -- one piece of a row that does not fit one CTA's 1152 modes
// (PMGRID = 4096: two pieces of 1025 and 1024 modes): everything but z is CTA-uniform.
template <bool SPLIT, int FM>
__global__ void __launch_bounds__(K3_TMA_THREADS, 8)
k3_scale_row_kernel(C2<double> *__restrict__ grid, int N, long long plane0,
                    const double *__restrict__ tab, const K3Params prm, const K3Greens gr, int segs, int seg_len)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bar;
    C2<double> *buf = (C2<double> *) smem_raw;
    const int L = N / 2 + 1;
    int row0, nel, z0 = 0;
    if (SPLIT) {
        row0 = (int) (blockIdx.x / (unsigned) segs);
        z0 = (int) (blockIdx.x - (unsigned) row0 * (unsigned) segs) * seg_len;
        nel = min(seg_len, L - z0);
    } else {
        row0 = blockIdx.x;
        nel = L;
    }
    const unsigned bytes = (unsigned) nel * (unsigned) sizeof(C2<double>);
    C2<double> *base = grid + (size_t) row0 * L + z0;
    if (threadIdx.x == 0) k3_bulk_load(buf, base, bytes, &bar);
    // factors while the copy is in flight
    const int pl = row0 / N, j = row0 - pl * N;
    const long long gi = plane0 + pl;
    const int ki = gi <= N / 2 ? (int) gi : (int) (gi - N);
    const int kj = j <= N / 2 ? j : j - N;
    const int c0 = ki * ki + kj * kj;
    const double w0 = gr.on ? k3_greens_row(gr, ki, kj) : 1.0;   // -exp(-(kx^2+ky^2) asmth2) (iwx iwy)^4 of this row
    double smth[K3_EPT];
    if ((unsigned) c0 + (unsigned) z0 * (unsigned) z0 >= prm.k2_narrow) {
        // every mode of this piece sits in a narrow segment at or above the first knot (all rows but the few around the
        // k_x = k_y = 0 axis): branch-free factors, whole iterations without a bounds check
        const int nfull = nel / K3_TMA_THREADS;
#pragma unroll
        for (int k = 0; k < K3_EPT; k++) {
            const int e = threadIdx.x + K3_TMA_THREADS * k;
            smth[k] = 1.0;
            if (k < nfull || e < nel) {
                const int z = z0 + e, k2i = c0 + z * z;
                smth[k] = k3_factor_narrow<FM>(k2i, tab, prm);
                if (gr.on) smth[k] *= k3_greens(gr, k2i, w0, z);
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < K3_EPT; k++) {
            const int e = threadIdx.x + K3_TMA_THREADS * k;
            smth[k] = 1.0;
            if (e < nel) {
                const int z = z0 + e, k2i = c0 + z * z;
                if (k2i > 0) smth[k] = k3_factor<FM>(k2i, tab, prm);                            // F(0,0,0) keeps factor 1 ...
                if (gr.on) smth[k] = k2i > 0 ? smth[k] * k3_greens(gr, k2i, w0, z) : 0.0;       // ... or is zeroed with the potential
            }
        }
    }
    __syncthreads();                       // the barrier was initialised before anyone polls it
    k3_bulk_wait(&bar);
#pragma unroll
    for (int k = 0; k < K3_EPT; k++) {
        const int e = threadIdx.x + K3_TMA_THREADS * k;
        if (e < nel) {
            C2<double> v = buf[e];
            k3_apply<double>(v, smth[k]);
            buf[e] = v;
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the bulk store
    __syncthreads();
    if (threadIdx.x == 0) k3_bulk_store(base, buf, bytes);
}

// FLAT chunks: the slab cut into pieces of an even number of modes, ignoring row boundaries; a thread finds row and z of
// its modes from the chunk's first mode (one division per thread, none per mode).  Serves
//  * double grids whose rows are short enough for several to share a CTA (PMGRID <= 1150: whole rows per chunk,
//    128 threads; 6.6 TB/s at 1024^3 against 5.3 for a kernel that divides per mode), and
//  * float grids (256 threads): a float row is (N/2+1)*8 bytes -- 8200 at PMGRID = 2048 -- so every other row starts
//    8 bytes off the 16-byte granule bulk copies need; chunks of an even mode count (two whole rows, 16400 B, at 2048) do not.
constexpr int K3_FLAT_THREADS = 256;
constexpr int K3_FLAT_MAX_ROWS = 64;       // rows a chunk may touch (the launcher falls back to plain loads beyond)

template <typename real, int THREADS, int FM>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS)
k3_scale_flat_kernel(C2<real> *__restrict__ grid, long long total, int chunk, int N, long long plane0,
                     const double *__restrict__ tab, const K3Params prm, const K3Greens gr)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bar;
    C2<real> *buf = (C2<real> *) smem_raw;
    const int L = N / 2 + 1;
    const long long e0 = (long long) blockIdx.x * chunk;           // first mode of this CTA's chunk (slab-relative, even)
    const int nel = (int) min((long long) chunk, total - e0);      // even: chunk and total are
    const unsigned bytes = (unsigned) nel * (unsigned) sizeof(C2<real>);
    C2<real> *base = grid + e0;
    if (threadIdx.x == 0) k3_bulk_load(buf, base, bytes, &bar);
    // (plane, row-in-plane, z) of the chunk's first mode
    const long long r0 = e0 / L;
    const int zb = (int) (e0 - r0 * L);
    const long long pl0 = r0 / N;
    const int j0 = (int) (r0 - pl0 * N);
    using fac_t = typename std::conditional<FM == FM_F32, float, double>::type;
    fac_t smth[K3_EPT];
    // per row of the chunk: kx^2 + ky^2 and the packed wave numbers (a chunk holds whole rows plus, at most, two cut ones)
    __shared__ int c_s[K3_FLAT_MAX_ROWS], kk_s[K3_FLAT_MAX_ROWS];
    __shared__ int fast_s;
    const int rows = (zb + nel + L - 1) / L;
    if (threadIdx.x == 0) fast_s = 1;
    __syncthreads();
    if ((int) threadIdx.x < rows) {
        int j = j0 + (int) threadIdx.x;
        long long gi = plane0 + pl0;
        if (j >= N) { const int q = j / N; j -= q * N; gi += q; }
        const int ki = gi <= N / 2 ? (int) gi : (int) (gi - N);
        const int kj = j <= N / 2 ? j : j - N;
        c_s[threadIdx.x] = ki * ki + kj * kj;
        kk_s[threadIdx.x] = (ki << 16) | (kj & 0xffff);
        if ((unsigned) (ki * ki + kj * kj) < prm.k2_narrow) fast_s = 0;        // (a benign race: everybody writes 0)
    }
    __syncthreads();
    // row of a mode: floor((zb + e) / L) by a multiply (exact for (zb + e) * L < 2^32: both are below 2^15 here)
    const unsigned magic = 0xffffffffu / (unsigned) L + 1u;
    if (fast_s) {
        // every row of this chunk lies where all segments are narrow (all but the rows around the k_x = k_y = 0 axis)
        const int nfull = nel / THREADS;
#pragma unroll
        for (int k = 0; k < K3_EPT; k++) {
            const int e = threadIdx.x + THREADS * k;
            smth[k] = FM == FM_F32 ? (fac_t) 0 : (fac_t) 1;
            if (k < nfull || e < nel) {
                const int zz = zb + e, rl = (int) __umulhi((unsigned) zz, magic), z = zz - rl * L;
                const int k2i = c_s[rl] + z * z;
                if constexpr (FM == FM_F32) {
                    smth[k] = k3_delta_f32_narrow(k2i, tab, prm);
                } else {
                    smth[k] = k3_factor_narrow<FM>(k2i, tab, prm);
                    if (gr.on) { const int kk = kk_s[rl]; smth[k] *= k3_greens(gr, k2i, k3_greens_row(gr, kk >> 16, (int) (short) (kk & 0xffff)), z); }
                }
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < K3_EPT; k++) {
            const int e = threadIdx.x + THREADS * k;
            smth[k] = FM == FM_F32 ? (fac_t) 0 : (fac_t) 1;            // FM_F32 holds factor - 1
            if (e < nel) {
                const int zz = zb + e, rl = (int) __umulhi((unsigned) zz, magic), z = zz - rl * L;
                const int k2i = c_s[rl] + z * z;
                if constexpr (FM == FM_F32) {
                    if (k2i > 0) smth[k] = k3_delta_f32(k2i, tab, prm);       // (never launched with the Green's function on)
                } else {
                    if (k2i > 0) smth[k] = k3_factor<FM>(k2i, tab, prm);                                         // F(0,0,0) keeps factor 1 ...
                    if (gr.on) {                                                                                 // ... or is zeroed
                        const int kk = kk_s[rl];
                        smth[k] = k2i > 0 ? smth[k] * k3_greens(gr, k2i, k3_greens_row(gr, kk >> 16, (int) (short) (kk & 0xffff)), z) : 0.0;
                    }
                }
            }
        }
    }
    __syncthreads();                       // the barrier was initialised before anyone polls it
    k3_bulk_wait(&bar);
#pragma unroll
    for (int k = 0; k < K3_EPT; k++) {
        const int e = threadIdx.x + THREADS * k;
        if (e < nel) {
            C2<real> v = buf[e];
            if constexpr (FM == FM_F32) k3_apply_delta(v, smth[k]); else k3_apply<real>(v, smth[k]);
            buf[e] = v;
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the bulk store
    __syncthreads();
    if (threadIdx.x == 0) k3_bulk_store(base, buf, bytes);
}

// host copy of the table of the most recent scaling pass (ksn_last_k3_table: bench.py and the tests check a pass against
// the CPU restatement with exactly the table the pass used)
static struct { double *logkk, *ratio; int nbins, cap; double norm, boxsize; } g_k3_tab_copy = { nullptr, nullptr, 0, 0, 0.0, 0.0 };

// table layout (4-byte words from the base): K3Seg[n+1] | u16 cell[cells] | u32 kthr[n+2] | pad to 16 B | K3SegF[n+1]
static size_t k3_tab_words(int n, int cells, int *off_kthr, int *off_segf)
{
    size_t w = (size_t) 8 * (n + 1) + (size_t) cells / 2;
    if (off_kthr) *off_kthr = (int) w;
    w += (size_t) n + 2;
    w = (w + 3) & ~(size_t) 3;
    if (off_segf) *off_segf = (int) w;
    return w + (size_t) 4 * (n + 1);
}
static size_t k3_tab_doubles(int n, int cells) { return (k3_tab_words(n, cells, nullptr, nullptr) + 1) / 2; }
static K3Params g_k3prm;
static int g_k3_fm_double = FM_D9;         // series length the double passes may use with the current table
static bool g_k3_f32_ok = false;           // the current table qualifies for the all-float pass over float grids
static bool g_k3_multi = false;            // some lookup cell holds more than one knot: the segment search loops

// Build the device table for (logkk, ratio, norm) in the host buffer `hbuf` (k3_tab_doubles(nbins, K3_MAX_CELLS) doubles)
// and decide how the passes may evaluate it; pure host arithmetic, no device involved (ksn_k3_table_plan exports the
// decisions so that a CPU test can pin them).  Returns the doubles to upload.
// Everything in the table that depends on the knots alone.  The knots are log(keff) of the non-empty bins: the same numbers
// step after step for as long as the slab geometry stands, while the ratios change every step -- and this half is the one
// with a libm call and a division per knot (35 of the 50 us the whole build took on the host, between K2 and K3 of every
// step).  Kept from one build to the next, compared by content.
struct K3Knots {
    bool valid = false;
    int dims = 0, nbins = 0, cells = 0;
    double boxsize = 0, lo = 0, scale = 0;
    bool multi = false;
    unsigned k2_narrow = 0xffffffffu;      // (without the KSN_K3_NOFAST / multi override)
    std::vector<double> logkk, K2, inv, den, umax;
    std::vector<unsigned> kthr;            // nbins + 2
    std::vector<unsigned short> cell;
};
static K3Knots g_k3_knots;
static void k3_forget_knots() { g_k3_knots.valid = false; }

static void k3_build_knots(K3Knots &kn, int dims, double boxsize, const double *logkk, int nbins)
{
    kn.valid = false;
    kn.dims = dims; kn.nbins = nbins; kn.boxsize = boxsize;
    kn.logkk.assign(logkk, logkk + nbins);
    kn.K2.resize(nbins); kn.inv.resize(nbins); kn.den.resize(nbins); kn.umax.resize(nbins); kn.kthr.resize(nbins + 2);
    const double unit = boxsize / (2 * M_PI);
    double min_gap = 1e300;
    for (int i = 0; i < nbins; i++) {
        const double kg = exp(logkk[i]) * unit;     // knot in integer-wave-number units
        kn.K2[i] = kg * kg;
        kn.inv[i] = 1.0 / kn.K2[i];
        kn.den[i] = (i + 1 < nbins) ? 2.0 * (logkk[i + 1] - logkk[i]) : 0.0;
        if (i > 0) min_gap = fmin(min_gap, 2.0 * (logkk[i] - logkk[i - 1]) / M_LN2);   // gap in log2(k2)
    }
    for (int i = 0; i < nbins; i++) kn.umax[i] = i + 1 < nbins ? kn.K2[i + 1] * kn.inv[i] - 1.0 : 0.0;
    // the cells run from the first knot to the largest k^2 of the grid (or the last knot, whichever is larger), so that a
    // mode at or above the first knot needs no bound on its cell index
    const double k2max = 3.0 * (double) (dims / 2) * (double) (dims / 2);
    const double lunit = log(unit);
    auto lg2K2 = [&](int i) { return 2.0 * (logkk[i] + lunit) * (1.0 / M_LN2); };     // log2(K2_i) without a libm call per knot
    const double lo = lg2K2(0), hi = fmax(lg2K2(nbins - 1), log2(fmax(k2max, 2.0)) + 1e-3);
    // cells narrow enough that (with the float-rounding guard) no cell sees two knots; where the knots are closer than the
    // finest cells (keff values are data-dependent means: nothing forbids it, and gsl_interp has no such limit) the
    // segment search steps over the extra knots in a loop instead
    int cells = 1024;
    while (cells < K3_MAX_CELLS && (hi - lo) / cells * 1.05 >= min_gap) cells *= 2;
    kn.multi = (hi - lo) / cells * 1.05 >= min_gap;
    const double scale = cells / (hi - lo);
    kn.cells = cells; kn.lo = lo; kn.scale = scale;
    kn.cell.resize(cells);
    // cell[k] = last knot at or below the lower edge of cell k, minus a guard (0.02 cell) for the float log2 on the
    // device.  One log2 per knot: knot i starts to count from cell ceil((log2 K2_i - lo) * scale + 0.02).
    {
        int k = 0;
        for (int i = 1; i < nbins; i++) {
            int first = (int) ceil((lg2K2(i) - lo) * scale + 0.02);
            if (first > cells) first = cells;
            for (; k < first; k++) kn.cell[k] = (unsigned short) (i - 1);
        }
        for (; k < cells; k++) kn.cell[k] = (unsigned short) (nbins - 1);
    }
    // integer knot thresholds: for integer k2,  k2 >= K2_i  <=>  k2 >= ceil(K2_i)
    for (int i = 0; i < nbins; i++) kn.kthr[i] = kn.K2[i] >= 4294967295.0 ? 0xffffffffu : (unsigned) ceil(kn.K2[i]);
    kn.kthr[nbins] = kn.kthr[nbins + 1] = 0xffffffffu;
    {
        // first knot from which on all segments are narrow (the last one, clamped above, has B = 0: any u will do)
        int j = 0;
        for (int i = 0; i + 1 < nbins; i++)
            if (!(kn.umax[i] < 0.03125)) j = i + 1;
        kn.k2_narrow = kn.kthr[j];
    }
    kn.valid = true;
}

// Build the device table for (logkk, ratio, norm) in the host buffer `hbuf` (k3_tab_doubles(nbins, K3_MAX_CELLS) doubles)
// and decide how the passes may evaluate it; pure host arithmetic, no device involved (ksn_k3_table_plan exports the
// decisions so that a CPU test can pin them, ksn_k3_table_hash the bytes).  Returns the doubles to upload.
static size_t k3_build_table(void *hbuf, int dims, double boxsize, const double *logkk, const double *ratio, int nbins, double norm)
{
    K3Knots &kn = g_k3_knots;
    if (!(kn.valid && kn.dims == dims && kn.nbins == nbins && kn.boxsize == boxsize &&
          memcmp(kn.logkk.data(), logkk, sizeof(double) * nbins) == 0))
        k3_build_knots(kn, dims, boxsize, logkk, nbins);
    K3Seg *seg = (K3Seg *) hbuf;
    const int cells = kn.cells;
    int off_kthr, off_segf;
    k3_tab_words(nbins, cells, &off_kthr, &off_segf);
    K3SegF *segf = (K3SegF *) ((unsigned *) hbuf + off_segf);
    // how much arithmetic the table needs.  Narrow segments (u = k2/K2_i - 1 < 2^-5 over the whole segment) use a series
    // for ln(1+u): to u^5 where |B| u_max^6 / 6 <= 1e-14 on every one of them (four orders below the 1e-10 bar), else to u^9.  The all-float pass over a
    // float grid: error of delta = factor - 1 about 2^-23 (|B| (1 + 3 u) + |delta|), wanted below 2^-26 (a quarter of a float ulp of the product).
    double worst_d5 = 0, worst_f32 = 0;
    for (int i = 0; i < nbins; i++) {
        seg[i].K2 = kn.K2[i];
        seg[i].inv = kn.inv[i];
        seg[i].A = 1.0 + norm * ratio[i];
        seg[i].B = (i + 1 < nbins) ? norm * (ratio[i + 1] - ratio[i]) / kn.den[i] : 0.0;
        const double umax = kn.umax[i];
        const double ab = fabs(seg[i].B);
        const double uc = umax < 0.03125 ? umax : 0.03125, uc2 = uc * uc;   // (of a wide segment: the part below the log1p switch)
        // (plain comparisons where fmax / fmin would be calls into libm on this path; same values, NaNs included)
        const double d5 = ab * (uc2 * uc2 * uc2) * (1.0 / 6.0);
        const double f32 = ab * (1.0 + 3.0 * (umax < 1.0 ? umax : 1.0)) + fabs(seg[i].A - 1.0) + ab * (umax < 1e30 ? umax : 1e30);   // (ln(1+u) <= u)
        worst_d5 = d5 > worst_d5 ? d5 : worst_d5;
        worst_f32 = f32 > worst_f32 ? f32 : worst_f32;
        segf[i].inv = (float) seg[i].inv; segf[i].A1 = (float) (norm * ratio[i]); segf[i].B = (float) seg[i].B; segf[i].pad = 0;
    }
    seg[nbins].K2 = INFINITY; seg[nbins].inv = 0; seg[nbins].A = seg[nbins - 1].A; seg[nbins].B = 0;
    segf[nbins] = segf[nbins - 1]; segf[nbins].B = 0; segf[nbins].inv = 0;
    memcpy(seg + nbins + 1, kn.cell.data(), sizeof(unsigned short) * cells);
    memcpy((unsigned *) hbuf + off_kthr, kn.kthr.data(), sizeof(unsigned) * (nbins + 2));
    if (g_k3_tab_copy.cap < nbins) {
        free(g_k3_tab_copy.logkk); free(g_k3_tab_copy.ratio);
        g_k3_tab_copy.logkk = (double *) malloc(sizeof(double) * nbins);
        g_k3_tab_copy.ratio = (double *) malloc(sizeof(double) * nbins);
        g_k3_tab_copy.cap = (g_k3_tab_copy.logkk && g_k3_tab_copy.ratio) ? nbins : 0;
    }
    g_k3_tab_copy.nbins = 0;
    if (g_k3_tab_copy.cap >= nbins) {
        memcpy(g_k3_tab_copy.logkk, logkk, sizeof(double) * nbins);
        memcpy(g_k3_tab_copy.ratio, ratio, sizeof(double) * nbins);
        g_k3_tab_copy.nbins = nbins; g_k3_tab_copy.norm = norm; g_k3_tab_copy.boxsize = boxsize;
    }
    g_k3_multi = kn.multi;
    g_k3_fm_double = worst_d5 <= 1e-14 ? FM_D5 : FM_D9;
    g_k3_f32_ok = worst_f32 <= 1.0 / 8 && 3.0 * (double) dims * dims / 4 < 16777216.0 && seg[0].K2 > 1e-30;
    g_k3prm.n = nbins;
    g_k3prm.cells = cells;
    g_k3prm.cell_lo = (float) kn.lo;
    g_k3prm.cell_scale = (float) kn.scale;
    g_k3prm.off_kthr = off_kthr;
    g_k3prm.off_segf = off_segf;
    g_k3prm.multi = g_k3_multi ? 1 : 0;
    g_k3prm.cell_off = (float) (-kn.lo * kn.scale);
    g_k3prm.k2_narrow = g_k3_multi || getenv("KSN_K3_NOFAST") ? 0xffffffffu : kn.k2_narrow;
    return k3_tab_doubles(nbins, cells);
}

int k3_upload_table(int dims, double boxsize, const double *logkk, const double *ratio, int nbins, double norm)
{
    Ctx &c = ctx();
    const size_t nd_max = k3_tab_doubles(nbins, K3_MAX_CELLS);
    int rc = ensure_device_buffer((void **) &c.d_k3tab, &c.k3tab_cap, nd_max * sizeof(double));
    if (rc) return rc;
    rc = ensure_pinned_buffer((void **) &c.h_k3tab, &c.h_k3tab_cap, nd_max * sizeof(double));
    if (rc) return rc;
    // the pinned table may still be in flight from the previous step
    KSN_CUDA(cudaStreamSynchronize(c.stream));
    const size_t nd = k3_build_table(c.h_k3tab, dims, boxsize, logkk, ratio, nbins, norm);
    KSN_CUDA(cudaMemcpyAsync(c.d_k3tab, c.h_k3tab, nd * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    return KSN_OK;
}

static K3Greens g_k3greens = { 0, 0.0, nullptr, nullptr };
static char g_k3_last[160] = "none";

// Switch the fused Green's function on (invwin: host table iw[0..N/2], as K1 takes it) or off (invwin == nullptr).
static int k3_set_greens(int dims, const double *invwin, double asmth2)
{
    g_k3greens.on = 0;
    if (!invwin) return KSN_OK;
    Ctx &c = ctx();
    const size_t L = (size_t) dims / 2 + 1;
    int rc = ensure_device_buffer((void **) &c.d_iw, &c.iw_cap, (2 * L + K1_WZ_PAD) * sizeof(double));   // K1's layout: iw | z weights
    if (rc) return rc;
    rc = ensure_device_buffer((void **) &c.d_gz, &c.gz_cap, L * sizeof(double));
    if (rc) return rc;
    double *gz = (double *) malloc(L * sizeof(double));
    if (!gz) return set_error(KSN_ENOMEM, "K3: out of host memory");
    for (size_t z = 0; z < L; z++) {
        const double w2 = invwin[z] * invwin[z];
        gz[z] = exp(-(double) (z * z) * asmth2) * (w2 * w2);
    }
    k1_tables_invalidate();                 // c.d_iw is K1's table buffer: it must re-upload its own next time
    cudaError_t e = cudaMemcpyAsync(c.d_iw, invwin, L * sizeof(double), cudaMemcpyHostToDevice, c.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(c.d_gz, gz, L * sizeof(double), cudaMemcpyHostToDevice, c.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c.stream);       // gz is about to be freed
    free(gz);
    KSN_CUDA(e);
    g_k3greens.on = 1;
    g_k3greens.asmth2 = asmth2;
    g_k3greens.iw = c.d_iw;
    g_k3greens.gz = c.d_gz;
    return KSN_OK;
}

int k3_launch(void *dgrid, int real_bytes, int dims, long long plane0_global, long long nplanes, int nknots)
{
    (void) nknots;
    Ctx &c = ctx();
    const K3Greens gr = g_k3greens;
    if (nplanes * dims > 0x7fffffffLL) return set_error(KSN_EINVAL, "K3: %lld rows in one slab", nplanes * dims);
    const int nrows = (int) (nplanes * dims);
    if (nrows == 0) return KSN_OK;
    constexpr int U = 4;
    const int L = dims / 2 + 1;
    const long long total = (long long) nrows * L;
    const int cap = K3_TMA_THREADS * K3_EPT;                    // modes one 128-thread CTA holds
    const int segs = (L + cap - 1) / cap;                       // pieces per row when a row is longer than that (PMGRID > 2302)
    // KSN_K3_EXACT=1 (tests): always the 9-term double factor, whatever the table would allow
    const bool exact = getenv("KSN_K3_EXACT") != nullptr;
    // (the all-float factor only without the Green's function, whose range is not that of "1 + small")
    const int fm = exact ? FM_D9 : (real_bytes == 4 && g_k3_f32_ok && !gr.on) ? FM_F32 : g_k3_fm_double;
    static const char *const fm_name[] = { "series to u^9", "series to u^5", "factor in float" };
    const char *gname = gr.on ? ", Green's function fused" : "";
    const bool aligned = ((uintptr_t) dgrid & 15) == 0;
    const bool bulk = !getenv("KSN_K3_NOTMA") && aligned;       // bulk copies need 16-byte granules
    auto shared_cfg = [&](auto kern, size_t smem) -> int {
        return func_attributes((const void *) kern, smem, 58);      // carve-out ~132 of 228 KB: L1 keeps the tables
    };
    auto done = [&]() -> int {
        c.launches++;
        KSN_CUDA(cudaGetLastError());
        return KSN_OK;
    };
    if (real_bytes == 8 && bulk && (segs == 1 || (!getenv("KSN_K3_NOSPLIT") && (long long) nrows * segs <= 0x7fffffffLL))) {
        // ~16 KB of grid per CTA: one row at PMGRID = 2048 (7 CTAs per SM inside a 132 KB carve-out), half a row at 4096,
        // several whole rows (as one flat chunk) where they are shorter
        const size_t row_bytes = (size_t) L * 16;
        int rpc = segs > 1 ? 1 : cap / L;
        while (rpc > 1 && rpc * row_bytes > 18432) rpc--;
        rpc = min(rpc, K3_FLAT_MAX_ROWS - 2);                      // (tiny grids: the kernel's per-row table is that long)
        if (segs == 1 && rpc > 1) {
            const int chunk = rpc * L;
            const size_t smem = (size_t) chunk * 16 + 128;
            const long long nct = (total + chunk - 1) / chunk;
            auto go = [&](auto kern) -> int {
                int rc = shared_cfg(kern, smem);
                if (rc) return rc;
                kern<<<(unsigned) nct, K3_TMA_THREADS, smem, c.stream>>>((C2<double> *) dgrid, total, chunk, dims, plane0_global, c.d_k3tab, g_k3prm, gr);
                return KSN_OK;
            };
            const int rc = fm == FM_D5 ? go(k3_scale_flat_kernel<double, K3_TMA_THREADS, FM_D5>) : go(k3_scale_flat_kernel<double, K3_TMA_THREADS, FM_D9>);
            if (rc) return rc;
            snprintf(g_k3_last, sizeof g_k3_last, "k3_scale_flat_kernel<double> (%d rows = %d modes per CTA, %s%s)", rpc, chunk, fm_name[fm], gname);
            return done();
        }
        const int seg_len = (L + segs - 1) / segs;
        const size_t smem = (size_t) (segs > 1 ? seg_len : L) * 16 + 128;
        const int nct = nrows * segs;
        auto go = [&](auto kern) -> int {
            int rc = shared_cfg(kern, smem);
            if (rc) return rc;
            kern<<<nct, K3_TMA_THREADS, smem, c.stream>>>((C2<double> *) dgrid, dims, plane0_global, c.d_k3tab, g_k3prm, gr, segs, seg_len);
            return KSN_OK;
        };
        const int rc = segs > 1 ? (fm == FM_D5 ? go(k3_scale_row_kernel<true, FM_D5>) : go(k3_scale_row_kernel<true, FM_D9>))
                                : (fm == FM_D5 ? go(k3_scale_row_kernel<false, FM_D5>) : go(k3_scale_row_kernel<false, FM_D9>));
        if (rc) return rc;
        snprintf(g_k3_last, sizeof g_k3_last, "k3_scale_row_kernel<%s> (one row per CTA, %d piece%s per row, %s%s)", segs > 1 ? "split" : "whole",
                 segs, segs > 1 ? "s" : "", fm_name[fm], gname);
        return done();
    }
    if (real_bytes == 4 && bulk && total % 2 == 0) {
        const int capf = K3_FLAT_THREADS * K3_EPT, pair = 2 * L;
        // whole row pairs (~16-18 KB of them) where a pair fits one CTA, else the largest even piece
        const int chunk = pair <= capf ? pair * max(1, min(min(capf / pair, (K3_FLAT_MAX_ROWS - 2) / 2), (int) (18432 / ((size_t) pair * 8)))) : (capf & ~1);
        const long long nct = (total + chunk - 1) / chunk;
        if (nct <= 0x7fffffffLL && chunk / L + 2 <= K3_FLAT_MAX_ROWS) {
            const size_t smem = (size_t) chunk * 8 + 128;
            auto go = [&](auto kern) -> int {
                int rc = shared_cfg(kern, smem);
                if (rc) return rc;
                kern<<<(unsigned) nct, K3_FLAT_THREADS, smem, c.stream>>>((C2<float> *) dgrid, total, chunk, dims, plane0_global, c.d_k3tab, g_k3prm, gr);
                return KSN_OK;
            };
            const int rc = fm == FM_F32 ? go(k3_scale_flat_kernel<float, K3_FLAT_THREADS, FM_F32>)
                         : fm == FM_D5 ? go(k3_scale_flat_kernel<float, K3_FLAT_THREADS, FM_D5>) : go(k3_scale_flat_kernel<float, K3_FLAT_THREADS, FM_D9>);
            if (rc) return rc;
            snprintf(g_k3_last, sizeof g_k3_last, "k3_scale_flat_kernel<float> (%d modes per CTA, %s%s)", chunk, fm_name[fm], gname);
            return done();
        }
    }
    // plain loads and stores, one CTA per ~1024 modes: a single row for large grids, several short rows otherwise
    const int rows_per_cta = L >= K3_THREADS * U ? 1 : (K3_THREADS * U) / L;
    const int ctas = (nrows + rows_per_cta - 1) / rows_per_cta;
    auto go = [&](auto kern, auto *g) {
        kern<<<ctas, K3_THREADS, 0, c.stream>>>(g, nrows, rows_per_cta, dims, plane0_global, c.d_k3tab, g_k3prm, gr);
    };
    if (real_bytes == 8) {
        if (fm == FM_D5) go(k3_scale_kernel<double, U, FM_D5>, (C2<double> *) dgrid); else go(k3_scale_kernel<double, U, FM_D9>, (C2<double> *) dgrid);
    } else {
        if (fm == FM_F32) go(k3_scale_kernel<float, U, FM_F32>, (C2<float> *) dgrid);
        else if (fm == FM_D5) go(k3_scale_kernel<float, U, FM_D5>, (C2<float> *) dgrid);
        else go(k3_scale_kernel<float, U, FM_D9>, (C2<float> *) dgrid);
    }
    snprintf(g_k3_last, sizeof g_k3_last, "k3_scale_kernel<%s> (plain loads, %d row%s per CTA, %s%s)", real_bytes == 8 ? "double" : "float", rows_per_cta,
             rows_per_cta > 1 ? "s" : "", fm_name[fm], gname);
    return done();
}

// K3 over a host-resident slab, chunk by chunk, each chunk copied back to the host buffer as soon as it is scaled.
// resident: the slab already sits in c.d_stage (uploaded by K1's staged path).  Otherwise (streaming plan) every chunk is
// uploaded again into the ring; uploads run on copy_stream and downloads on copy_stream2, so they overlap on the full-duplex
// PCIe link and the pass costs about one transfer time, like the resident one.
int k3_over_staged_grid(void *hgrid, int real_bytes, int dims, long long startslab, long long nslab, int nknots, bool resident)
{
    Ctx &c = ctx();
    StagePlan pl;
    int rc = stage_plan(real_bytes, dims, nslab, &pl);
    if (rc) return rc;
    if (resident && pl.streaming) return set_error(KSN_EINVAL, "K3: staging plan changed between the passes");
    const bool ring = !resident && pl.streaming;
    const int nev = ring ? STAGE_RING : pl.nchunks;
    cudaEvent_t *landed = (cudaEvent_t *) malloc(sizeof(cudaEvent_t) * nev), *scaled = (cudaEvent_t *) malloc(sizeof(cudaEvent_t) * nev),
                *sent = (cudaEvent_t *) malloc(sizeof(cudaEvent_t) * nev);
    for (int i = 0; i < nev; i++) {
        cudaEventCreateWithFlags(&landed[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&scaled[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&sent[i], cudaEventDisableTiming);
    }
    cudaStream_t down = resident ? c.copy_stream : c.copy_stream2;   // uploads and downloads overlap when both are in this pass
    if (!resident) {
        ensure_host_pinned(hgrid, pl.total);
        KSN_CUDA(cudaStreamSynchronize(c.stream));      // nothing of a previous call may still read the staging buffer
    }
    phase_begin(PH_K3);
    for (int i = 0; i < pl.nchunks && !rc; i++) {
        const long long p0 = i * pl.chunk, np = (p0 + pl.chunk <= nslab) ? pl.chunk : nslab - p0;
        const int s = ring ? i % STAGE_RING : i;
        char *slot = (char *) c.d_stage + (ring ? (size_t) s * pl.chunk : (size_t) p0) * pl.plane_bytes;
        if (!resident) {
            if (ring && i >= STAGE_RING) cudaStreamWaitEvent(c.copy_stream, sent[s], 0);
            cudaMemcpyAsync(slot, (const char *) hgrid + p0 * pl.plane_bytes, np * pl.plane_bytes, cudaMemcpyHostToDevice, c.copy_stream);
            cudaEventRecord(landed[s], c.copy_stream);
            cudaStreamWaitEvent(c.stream, landed[s], 0);
        }
        rc = k3_launch(slot, real_bytes, dims, startslab + p0, np, nknots);
        cudaEventRecord(scaled[s], c.stream);
        cudaStreamWaitEvent(down, scaled[s], 0);
        cudaMemcpyAsync((char *) hgrid + p0 * pl.plane_bytes, slot, np * pl.plane_bytes, cudaMemcpyDeviceToHost, down);
        cudaEventRecord(sent[s], down);
    }
    phase_end(PH_K3);
    cudaError_t e1 = cudaStreamSynchronize(c.copy_stream), e2 = cudaStreamSynchronize(c.stream);
    cudaError_t e3 = resident ? cudaSuccess : cudaStreamSynchronize(c.copy_stream2);
    for (int i = 0; i < nev; i++) { cudaEventDestroy(landed[i]); cudaEventDestroy(scaled[i]); cudaEventDestroy(sent[i]); }
    free(landed); free(scaled); free(sent);
    if (rc) return rc;
    KSN_CUDA(e1);
    KSN_CUDA(e2);
    KSN_CUDA(e3);
    phase_collect();
    return KSN_OK;
}

int k1_sums(const void *dgrid, const void *hgrid, int real_bytes, int dims, int nrbins, long long startslab, long long nslab,
            const unsigned int *thresholds, const double *invwin,
            double *power_sum, double *keff_sum, long long *count, double *total_mass2);

}  // namespace ksn

using namespace ksn;

extern "C" const char *ksn_last_k3_kernel(void) { return g_k3_last; }

// What k3_build_table decides for a table (host arithmetic only, no device needed): series = 5 or 9 (terms of ln(1+u) the
// double passes use), f32_ok = 1 if float grids may evaluate the factor in float, k2_narrow = the k^2 from which on every
// segment is narrow (rows at or above it take the branch-free path; 0xffffffff: none), cells = lookup cells in log2(k^2),
// multi = 1 if some cell holds more than one knot (the segment search loops).
extern "C" int ksn_k3_table_plan(int dims, double boxsize, const double *logkk, const double *ratio, int nbins, double norm,
                                 int *series, int *f32_ok, unsigned *k2_narrow, int *cells, int *multi)
{
    if (!logkk || !ratio || nbins < 2 || nbins > 65535 || dims < 2 || !(boxsize > 0)) return KSN_EINVAL;
    for (int i = 1; i < nbins; i++) if (!(logkk[i] > logkk[i - 1])) return KSN_EINVAL;
    void *buf = malloc(k3_tab_doubles(nbins, K3_MAX_CELLS) * sizeof(double));
    if (!buf) return KSN_ENOMEM;
    k3_build_table(buf, dims, boxsize, logkk, ratio, nbins, norm);
    free(buf);
    if (series) *series = g_k3_fm_double == FM_D5 ? 5 : 9;
    if (f32_ok) *f32_ok = g_k3_f32_ok ? 1 : 0;
    if (k2_narrow) *k2_narrow = g_k3prm.k2_narrow;
    if (cells) *cells = g_k3prm.cells;
    if (multi) *multi = g_k3prm.multi;
    return KSN_OK;
}

// FNV-1a hash of everything k3_build_table produces for a table -- the words that go to the device, the kernel parameters,
// the decisions -- so that a CPU test can pin that a build from the cached knot geometry equals a fresh one bit for bit.
// fresh != 0: forget the cached knot geometry first.
extern "C" int ksn_k3_table_hash(int dims, double boxsize, const double *logkk, const double *ratio, int nbins, double norm,
                                 int fresh, unsigned long long *hash)
{
    if (!hash || !logkk || !ratio || nbins < 2 || nbins > 65535 || dims < 2 || !(boxsize > 0)) return KSN_EINVAL;
    for (int i = 1; i < nbins; i++) if (!(logkk[i] > logkk[i - 1])) return KSN_EINVAL;
    const size_t cap = k3_tab_doubles(nbins, K3_MAX_CELLS) * sizeof(double);
    unsigned char *buf = (unsigned char *) calloc(cap, 1);
    if (!buf) return KSN_ENOMEM;
    if (fresh) k3_forget_knots();
    const size_t nd = k3_build_table(buf, dims, boxsize, logkk, ratio, nbins, norm);
    unsigned long long h = 1469598103934665603ull;
    auto eat = [&](const void *p, size_t n) { for (size_t i = 0; i < n; i++) { h ^= ((const unsigned char *) p)[i]; h *= 1099511628211ull; } };
    // (the table has no holes except the two bytes of padding per K3SegF and the alignment words, zeroed by calloc)
    eat(buf, nd * sizeof(double));
    eat(&g_k3prm.n, sizeof(int)); eat(&g_k3prm.cells, sizeof(int)); eat(&g_k3prm.cell_lo, sizeof(float)); eat(&g_k3prm.cell_scale, sizeof(float));
    eat(&g_k3prm.off_kthr, sizeof(int)); eat(&g_k3prm.off_segf, sizeof(int)); eat(&g_k3prm.multi, sizeof(int));
    eat(&g_k3prm.k2_narrow, sizeof(unsigned)); eat(&g_k3prm.cell_off, sizeof(float));
    const int dec[2] = { g_k3_fm_double, g_k3_f32_ok ? 1 : 0 };
    eat(dec, sizeof dec);
    free(buf);
    *hash = h;
    return KSN_OK;
}

extern "C" int ksn_last_k3_table(const double **logkk, const double **ratio, int *nbins, double *norm, double *boxsize)
{
    if (!g_k3_tab_copy.nbins) return set_error(KSN_EINVAL, "ksn_last_k3_table: no scaling pass yet");
    if (logkk) *logkk = g_k3_tab_copy.logkk;
    if (ratio) *ratio = g_k3_tab_copy.ratio;
    if (nbins) *nbins = g_k3_tab_copy.nbins;
    if (norm) *norm = g_k3_tab_copy.norm;
    if (boxsize) *boxsize = g_k3_tab_copy.boxsize;
    return KSN_OK;
}

static int check_table(const double *logkk, const double *ratio, int nbins, double norm)
{
    if (!logkk || !ratio || nbins < 2 || nbins > 65535) return set_error(KSN_EINVAL, "K3: bad table (nbins=%d)", nbins);
    for (int i = 1; i < nbins; i++)
        if (!(logkk[i] > logkk[i - 1])) return set_error(KSN_EINVAL, "K3: logkk must increase strictly (i=%d)", i);
    (void) norm;
    return KSN_OK;
}

static int scale_modes_impl(void *grid, int real_bytes, int dims, long long startslab, long long nslab,
                            double boxsize, const double *logkk, const double *ratio, int nbins, double norm,
                            const double *invwin, double asmth2)
{
    int rc = ensure_init();
    if (rc) return rc;
    if ((!grid && nslab > 0) || (real_bytes != 4 && real_bytes != 8) || dims < 2 || nslab < 0 || startslab < 0 ||
        startslab + nslab > dims || !(boxsize > 0))
        return set_error(KSN_EINVAL, "ksn_scale_modes: bad arguments");
    rc = check_table(logkk, ratio, nbins, norm);
    if (rc) return rc;
    if (nslab == 0) return KSN_OK;
    Ctx &c = ctx();
    rc = k3_upload_table(dims, boxsize, logkk, ratio, nbins, norm);
    if (rc) return rc;
    rc = k3_set_greens(dims, invwin, asmth2);
    if (rc) return rc;
    struct GreensOff { ~GreensOff() { g_k3greens.on = 0; } } greens_off;     // whatever happens below, the switch does not outlive this call
    if (ksn_pointer_is_device(grid)) {
        phase_begin(PH_K3);
        rc = k3_launch(grid, real_bytes, dims, startslab, nslab, nbins);
        phase_end(PH_K3);
        if (rc) return rc;
        KSN_CUDA(cudaStreamSynchronize(c.stream));
        phase_collect();
        return KSN_OK;
    }
    // host grid: upload, scale, download -- chunk by chunk
    return k3_over_staged_grid(grid, real_bytes, dims, startslab, nslab, nbins, false);
}

extern "C" int ksn_scale_modes(void *grid, int real_bytes, int dims, long long startslab, long long nslab,
                               double boxsize, const double *logkk, const double *ratio, int nbins, double norm)
{
    return scale_modes_impl(grid, real_bytes, dims, startslab, nslab, boxsize, logkk, ratio, nbins, norm, nullptr, 0.0);
}

extern "C" int ksn_scale_modes_greens(void *grid, int real_bytes, int dims, long long startslab, long long nslab,
                                      double boxsize, const double *logkk, const double *ratio, int nbins, double norm,
                                      const double *invwin, double asmth2)
{
    if (!invwin || !(asmth2 >= 0)) return set_error(KSN_EINVAL, "ksn_scale_modes_greens: bad arguments");
    return scale_modes_impl(grid, real_bytes, dims, startslab, nslab, boxsize, logkk, ratio, nbins, norm, invwin, asmth2);
}

static int step_staged_impl(void *hgrid, int real_bytes, int dims, int nrbins, long long startslab, long long nslab,
                            const unsigned int *thresholds, const double *invwin, double boxsize,
                            ksn_between_fn between, void *user, bool greens, double asmth2)
{
    int rc = ensure_init();
    if (rc) return rc;
    if ((!hgrid && nslab > 0) || !between || !thresholds || !invwin || (real_bytes != 4 && real_bytes != 8) || dims < 2 ||
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
    rc = k3_set_greens(dims, greens ? invwin : nullptr, asmth2);
    if (rc) return rc;
    struct GreensOff { ~GreensOff() { g_k3greens.on = 0; } } greens_off;
    if (on_device) {
        phase_begin(PH_K3);
        rc = k3_launch(hgrid, real_bytes, dims, startslab, nslab, nbins);
        phase_end(PH_K3);
        if (rc) return rc;
        KSN_CUDA(cudaStreamSynchronize(c.stream));
        phase_collect();
        return KSN_OK;
    }
    return k3_over_staged_grid(hgrid, real_bytes, dims, startslab, nslab, nbins, !c.stage_streaming);
}

extern "C" int ksn_step_staged(void *hgrid, int real_bytes, int dims, int nrbins, long long startslab, long long nslab,
                               const unsigned int *thresholds, const double *invwin, double boxsize,
                               ksn_between_fn between, void *user)
{
    return step_staged_impl(hgrid, real_bytes, dims, nrbins, startslab, nslab, thresholds, invwin, boxsize, between, user, false, 0.0);
}

extern "C" int ksn_step_staged_greens(void *hgrid, int real_bytes, int dims, int nrbins, long long startslab, long long nslab,
                                      const unsigned int *thresholds, const double *invwin, double boxsize,
                                      ksn_between_fn between, void *user, double asmth2)
{
    if (!(asmth2 >= 0)) return set_error(KSN_EINVAL, "ksn_step_staged_greens: bad arguments");
    return step_staged_impl(hgrid, real_bytes, dims, nrbins, startslab, nslab, thresholds, invwin, boxsize, between, user, true, asmth2);
}
