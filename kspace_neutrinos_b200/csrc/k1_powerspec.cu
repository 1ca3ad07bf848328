// K1 -- power-spectrum bin sums over the slab-decomposed r2c grid (sm_100a).
//
// Replaces hot loop 1 of the reference, powerspectrum.c:56-89, and the four
// MPI_Allreduce calls at powerspectrum.c:91-95.  HBM-bound: 16 B (double grid) read per
// stored mode, nothing written but O(nrbins) partials.
//
// Layout and schedule
//  * The local slab is swept row by row (a row = the N/2+1 complex values of one (i,j)):
//    one warp per row, 32 consecutive z per warp step, so every warp load is one coalesced
//    512-byte request and all row geometry (ki^2+kj^2, the x-y window product) is warp-uniform.
//    Rows are dealt round-robin to the warps of a persistent grid (one CTA per SM), so the
//    sweep is a pure function of (grid size, slab) -> deterministic.
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
#include "ksn_p2p.cuh"

#include <math.h>
#include <type_traits>
#include <algorithm>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace ksn {

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

// bare MUFU.LG2 (no denormal fix-up: arguments are integers >= 1, or 0 -> -inf)
__device__ __forceinline__ float fast_log2(float x)
{
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
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

// Segmented inclusive scan over runs of equal `bin` (lanes with the same bin are adjacent).
// On return the LAST lane of each run holds the run's sum.  The conditional add is an FMA with an
// exact 0.0/1.0 mask (x*1+v and x*0+v round like v+x and v), cheaper than add+select.
template <int NV>
__device__ __forceinline__ unsigned segmented_sum(int bin, double (&v)[NV], int lane, unsigned le_mask)
{
    const unsigned full = 0xffffffffu;
    const int prev = __shfl_up_sync(full, bin, 1);            // lane 0 receives its own bin
    const unsigned heads = __ballot_sync(full, bin != prev) | 1u;
    const int dist = lane - (31 - __clz(heads & le_mask));    // distance to the first lane of my run
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double m = __hiloint2double(dist >= d ? 0x3ff00000 : 0, 0);
#pragma unroll
        for (int i = 0; i < NV; i++) v[i] = fma(__shfl_up_sync(full, v[i], d), m, v[i]);
    }
    return (heads >> 1) | 0x80000000u;   // lanes that end a run
}

// One group of U warp steps (32 consecutive z each) of one row.  MASKED: steps may run past the row end.
template <typename real, bool FULL, int U, bool MASKED>
__device__ __forceinline__ void k1_group(const Cplx<real> *__restrict__ rowptr, int z0, int L, int nyq, int c, double wxy,
                                         float binscale, int nrbins, const uint2 *thr_s, const double *iw_s,
                                         double *mybins, int lane, bool origin_row)
{
    constexpr int NV = FULL ? 3 : 1;
    const unsigned eq_mask = 1u << lane, le_mask = 0xffffffffu >> (31 - lane);
    Cplx<real> v[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        const int z = z0 + 32 * u;
        if (!MASKED || z < L) v[u] = ld_stream(rowptr + z);
        else { v[u].re = 0; v[u].im = 0; }
    }
    if (origin_row && z0 == 0) { v[0].re = 0; v[0].im = 0; }   // F(0,0,0): the mean, not a mode (powerspectrum.c:65)
    int bin[U];
    double val[U][NV];
    // Phase A: bin + weighted power per element (independent across u)
#pragma unroll
    for (int u = 0; u < U; u++) {
        const int z = MASKED ? min(z0 + 32 * u, L - 1) : z0 + 32 * u;
        const int k2 = c + z * z;
        // MUFU estimate of floor(binsperunit*log(sqrt(k2))), exact after one step against the host thresholds
        int b = max((int) (binscale * fast_log2((float) k2)), 0);   // <= true bin + 1 <= nrbins: table has nrbins+1 pairs
        const uint2 t = thr_s[b];                       // {thr[b], thr[b+1]}
        b += ((unsigned) k2 >= t.y) - ((unsigned) k2 < t.x);
        // iw_s[z] carries the multiplicity: (2^(1/4) iw)^4 = 2 iw^4 for 0 < z < N/2 (kz=0 and Nyquist are not doubled)
        val[u][0] = mode_power(v[u], wxy, iw_s[z]);
        if (FULL) {
            const double m = (z == 0 || z == nyq) ? 1.0 : 2.0;
            val[u][1] = k2 > 0 ? sqrt((double) k2) * m : 0.0;
            val[u][2] = k2 > 0 ? m : 0.0;
        }
        if (MASKED && z0 + 32 * u >= L) {
            b = 0x7fffffff;
#pragma unroll
            for (int i = 0; i < NV; i++) val[u][i] = 0.0;
        }
        bin[u] = b;
    }
    // Phase B: the U segmented scans are independent -> their shuffle chains interleave
    unsigned tails[U];
#pragma unroll
    for (int u = 0; u < U; u++) tails[u] = segmented_sum<NV>(bin[u], val[u], lane, le_mask);
    // Phase C: run tails go to the warp-private bins in element order.  Bins are monotone along a row,
    // so the tails of one step hit distinct addresses: plain read-modify-write, no atomics.
#pragma unroll
    for (int u = 0; u < U; u++) {
        if ((tails[u] & eq_mask) && (!MASKED || bin[u] != 0x7fffffff)) {
#pragma unroll
            for (int i = 0; i < NV; i++) mybins[i * nrbins + bin[u]] += val[u][i];
        }
        __syncwarp();
    }
}

// Fast-path weight: the CIC deconvolution weight (iwx iwy iwz)^4 (powerspectrum.c:20-24,68) times the Hermitian
// multiplicity factorises into a per-row scalar (iwx iwy)^4 and a per-z table m_z iwz^4 (m = 1 for kz = 0 and the
// Nyquist plane, 2 otherwise): one multiply per mode; differs from the reference's rounding order by ~4e-16.
template <typename real>
__device__ __forceinline__ double pair_power(real re, real im, double w)
{
    const double a = (double) re, b = (double) im;
    return fma(b, b, a * a) * w;
}

// ------------------------------------------------------------------------------------------------
// Fast path (power only): every lane owns TWO consecutive modes, fetched with one 256-bit load
// (sm_100a LDG.256), so a warp step covers 64 modes with one segmented scan.  The two modes of a
// lane are combined first; a lane whose pair straddles a bin edge closes the lower run itself.
template <typename real> struct Pair;
template <> struct __align__(32) Pair<double> { double re0, im0, re1, im1; };
template <> struct __align__(16) Pair<float> { float re0, im0, re1, im1; };

__device__ __forceinline__ Pair<double> ld_pair(const Pair<double> *p)
{
    Pair<double> v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0, %1, %2, %3}, [%4];"
                 : "=d"(v.re0), "=d"(v.im0), "=d"(v.re1), "=d"(v.im1) : "l"(p));
    return v;
}
__device__ __forceinline__ Pair<float> ld_pair(const Pair<float> *p)
{
    Pair<float> v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.re0), "=f"(v.im0), "=f"(v.re1), "=f"(v.im1) : "l"(p));
    return v;
}

__device__ __forceinline__ int bin_of(int k2, float binscale, const uint2 *thr_s)
{
    int b = max((int) (binscale * fast_log2((float) k2)), 0);   // within one bin of the truth
    const uint2 t = thr_s[b];                                    // {thr[b], thr[b+1]}
    return b + ((unsigned) k2 >= t.y) - ((unsigned) k2 < t.x);
}

// A group = U warp steps of 32 pairs each, starting at pair index q0 (this lane's first pair).  zfirst = z of pair
// 0's first element (0 or 1, whichever makes the pairs 32-byte aligned in this row).  The work is split in three
// so that the row loop can issue the NEXT group's loads between phase A (which consumes the loaded registers) and
// phases B/C (shuffle scans + bin updates, ~60 % of the instructions): loads stay in flight behind the compute.
template <typename real, int U, bool MASKED>
__device__ __forceinline__ void pair_load(Pair<real> (&v)[U], const Cplx<real> *__restrict__ rowptr, int q0, int npairs, int zfirst)
{
#pragma unroll
    for (int u = 0; u < U; u++) {
        const int q = q0 + 32 * u;
        if (!MASKED || q < npairs) v[u] = ld_pair((const Pair<real> *) (rowptr + zfirst + 2 * q));
        else { v[u].re0 = 0; v[u].im0 = 0; v[u].re1 = 0; v[u].im1 = 0; }
    }
}

// Phase A: bins and weighted powers of both modes of every pair; x = the lane's contribution to the run that
// contains its second mode.
template <typename real, int U, bool MASKED>
__device__ __forceinline__ void pair_phaseA(Pair<real> (&v)[U], int q0, int npairs, int zfirst, int c, double wxy, float binscale,
                                            const uint2 *thr_s, const double *iw_s, bool origin_row,
                                            int (&ba)[U], int (&bb)[U], double (&pa)[U], double (&x)[U])
{
    if (origin_row && zfirst == 0 && q0 == 0) { v[0].re0 = 0; v[0].im0 = 0; }   // F(0,0,0) is not a mode
#pragma unroll
    for (int u = 0; u < U; u++) {
        const int q = MASKED ? min(q0 + 32 * u, npairs - 1) : q0 + 32 * u;
        const int z = zfirst + 2 * q;
        const int k2a = c + z * z, k2b = k2a + 2 * z + 1;
        ba[u] = bin_of(k2a, binscale, thr_s);
        bb[u] = bin_of(k2b, binscale, thr_s);
        pa[u] = pair_power<real>(v[u].re0, v[u].im0, wxy * iw_s[z]);
        const double pb = pair_power<real>(v[u].re1, v[u].im1, wxy * iw_s[z + 1]);
        if (MASKED && q0 + 32 * u >= npairs) { ba[u] = 0x7fffffff; bb[u] = 0x7fffffff; }
        x[u] = fma(pa[u], __hiloint2double(ba[u] == bb[u] ? 0x3ff00000 : 0, 0), pb);
    }
}

// Phases B and C: one segmented scan per step over the lanes' run contributions, then the run totals go to the
// warp-private bins (distinct bins within a step because bins are monotone along a row -> plain read-modify-write).
template <int U, bool MASKED>
__device__ __forceinline__ void pair_phaseBC(const int (&ba)[U], const int (&bb)[U], const double (&pa)[U], double (&x)[U],
                                             double *mybins, int lane)
{
    const unsigned full = 0xffffffffu, le_mask = full >> (31 - lane), eq_mask = 1u << lane;
    unsigned tails[U];
    double carry[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        const int prevkey = __shfl_up_sync(full, bb[u], 1);
        const bool cont = lane > 0 && ba[u] == prevkey;          // my first mode continues the previous lane's run
        const unsigned contmask = __ballot_sync(full, cont);
        const unsigned heads = __ballot_sync(full, !cont || ba[u] != bb[u]) | 1u;
        const int dist = lane - (31 - __clz(heads & le_mask));
        double sx = x[u];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
            sx = fma(__shfl_up_sync(full, sx, d), __hiloint2double(dist >= d ? 0x3ff00000 : 0, 0), sx);
        const double prev_sum = __shfl_up_sync(full, sx, 1);
        carry[u] = fma(prev_sum, __hiloint2double(cont ? 0x3ff00000 : 0, 0), pa[u]);   // closes the lower run if the pair is split
        x[u] = sx;
        tails[u] = ~(contmask >> 1) | 0x80000000u;               // lane l ends a run iff lane l+1 does not continue it
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
        if (!MASKED || ba[u] != 0x7fffffff) {
            if (ba[u] != bb[u]) mybins[ba[u]] += carry[u];
            if (tails[u] & eq_mask) mybins[bb[u]] += x[u];
        }
        __syncwarp();
    }
}

// one mode handled by lane 0 alone (the unpaired first or last element of a row)
template <typename real>
__device__ __forceinline__ void k1_single(const Cplx<real> *__restrict__ rowptr, int z, int c, double wxy, float binscale,
                                          const uint2 *thr_s, const double *iw_s, double *mybins, int lane)
{
    if (lane == 0) {
        const int k2 = c + z * z;
        if (k2 > 0) {
            const Cplx<real> e = ld_stream(rowptr + z);
            mybins[bin_of(k2, binscale, thr_s)] += pair_power<real>(e.re, e.im, wxy * iw_s[z]);
        }
    }
    __syncwarp();
}

template <typename real, int MAXW, int U>
__global__ void __launch_bounds__(MAXW * 32, 1)
k1_pair_kernel(const Cplx<real> *__restrict__ grid, int nrows, int N, int nrbins, long long plane0,
               float binscale, const unsigned *__restrict__ thr, const double *__restrict__ iw,
               double *__restrict__ partial, int accumulate)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int L = N / 2 + 1;
    const int nyq = N / 2;
    const int nwarps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *iw_s = (double *) smem_raw;                       // 2L: z table (with multiplicity) | plain table
    double *bins_s = iw_s + 2 * L;                            // nwarps * nrbins
    uint2 *thr_s = (uint2 *) (bins_s + (size_t) nwarps * nrbins);   // nrbins+1 pairs {thr[b], thr[b+1]}
    for (int i = threadIdx.x; i < L; i += blockDim.x) iw_s[L + i] = iw[i];              // plain 1-D window: x/y factors
    for (int i = threadIdx.x; i < L; i += blockDim.x) {                                 // m_z * iwz^4
        const double w2 = iw[i] * iw[i];
        iw_s[i] = ((i == 0 || i == nyq) ? 1.0 : 2.0) * (w2 * w2);
    }
    for (int i = threadIdx.x; i <= nrbins; i += blockDim.x)
        thr_s[i] = make_uint2(i < nrbins ? thr[i] : 0xffffffffu, i + 1 < nrbins ? thr[i + 1] : 0xffffffffu);
    for (int i = threadIdx.x; i < nwarps * nrbins; i += blockDim.x) bins_s[i] = 0.0;
    __syncthreads();

    double *mybins = bins_s + (size_t) warp * nrbins;
    // the slab base may itself sit on an odd 16-byte boundary
    const int base_odd = (int) (((size_t) grid / sizeof(Cplx<real>)) & 1);
    const int stride = gridDim.x * nwarps;

    struct Row { const Cplx<real> *ptr; int c, zfirst, npairs, nfull; double wxy; };
    auto setup = [&](int r) {
        Row R;
        const int pl = r / N, j = r - pl * N;
        const long long gi = plane0 + pl;
        const int ki = gi <= N / 2 ? (int) gi : (int) (gi - N);
        const int kj = j <= N / 2 ? j : j - N;
        R.c = ki * ki + kj * kj;
        const double a = iw_s[L + (ki < 0 ? -ki : ki)], b = iw_s[L + (kj < 0 ? -kj : kj)];
        const double t = a * b, t2 = t * t;
        R.wxy = t2 * t2;                            // (iwx iwy)^4
        R.ptr = grid + (size_t) r * L;
        // pairs must start on a 2-element boundary of the slab: skip z=0 when the row starts odd
        R.zfirst = (int) ((((size_t) r * L) + base_odd) & 1);
        R.npairs = (L - R.zfirst) >> 1;
        R.nfull = (R.npairs / 32) / U;              // groups of U steps that lie entirely inside the row
        return R;
    };

    int r = blockIdx.x * nwarps + warp;
    Pair<real> v[U];
    Row cur;
    if (r < nrows) {
        cur = setup(r);
        if (cur.nfull > 0) pair_load<real, U, false>(v, cur.ptr, lane, cur.npairs, cur.zfirst);
    }
    while (r < nrows) {
        // the unpaired first / last mode of the row
        if (cur.zfirst) k1_single<real>(cur.ptr, 0, cur.c, cur.wxy, binscale, thr_s, iw_s, mybins, lane);
        if ((L - cur.zfirst) & 1) k1_single<real>(cur.ptr, L - 1, cur.c, cur.wxy, binscale, thr_s, iw_s, mybins, lane);
        const int rn = r + stride;
        Row nxt = cur;
        int ba[U], bb[U];
        double pa[U], x[U];
#pragma unroll 1
        for (int g = 0; g < cur.nfull; g++) {
            pair_phaseA<real, U, false>(v, lane + g * 32 * U, cur.npairs, cur.zfirst, cur.c, cur.wxy, binscale, thr_s, iw_s, cur.c == 0, ba, bb, pa, x);
            // v is free again: put the next group's loads in flight before the shuffle-heavy phases
            if (g + 1 < cur.nfull) {
                pair_load<real, U, false>(v, cur.ptr, lane + (g + 1) * 32 * U, cur.npairs, cur.zfirst);
            } else if (rn < nrows) {
                nxt = setup(rn);
                if (nxt.nfull > 0) pair_load<real, U, false>(v, nxt.ptr, lane, nxt.npairs, nxt.zfirst);
            }
            pair_phaseBC<U, false>(ba, bb, pa, x, mybins, lane);
        }
        if (cur.nfull == 0 && rn < nrows) {          // rows shorter than one full group: nothing was prefetched
            nxt = setup(rn);
            if (nxt.nfull > 0) pair_load<real, U, false>(v, nxt.ptr, lane, nxt.npairs, nxt.zfirst);
        }
        // the rest of the row goes one (masked) step at a time
        const int nsteps = (cur.npairs + 31) / 32;
#pragma unroll 1
        for (int st = cur.nfull * U; st < nsteps; st++) {
            Pair<real> v1[1];
            int ba1[1], bb1[1];
            double pa1[1], x1[1];
            pair_load<real, 1, true>(v1, cur.ptr, lane + st * 32, cur.npairs, cur.zfirst);
            pair_phaseA<real, 1, true>(v1, lane + st * 32, cur.npairs, cur.zfirst, cur.c, cur.wxy, binscale, thr_s, iw_s, cur.c == 0, ba1, bb1, pa1, x1);
            pair_phaseBC<1, true>(ba1, bb1, pa1, x1, mybins, lane);
        }
        r = rn;
        cur = nxt;
    }
    __syncthreads();
    double *out = partial + (size_t) blockIdx.x * nrbins;
    for (int i = threadIdx.x; i < nrbins; i += blockDim.x) {
        double s = 0.0;
        for (int w = 0; w < nwarps; w++) s += bins_s[(size_t) w * nrbins + i];
        out[i] = accumulate ? out[i] + s : s;
    }
}


// ------------------------------------------------------------------------------------------------
// Tile path (default for power-only sweeps of a double grid).  The pair kernel above is ISSUE-bound (ncu: 67 warp
// instructions per 32 modes, most of them the per-step shuffle scans).  Here the work is laid out so that little but the
// arithmetic is left:
//  * a row is cut into tiles of 32*C modes (C = 4q+1); one TMA bulk copy (cp.async.bulk, completion on an mbarrier)
//    brings a tile into the warp's own shared-memory ring (S stages per warp) -- no load instructions, bytes in flight
//    bounded by shared memory, not registers;
//  * lane l walks the C CONSECUTIVE modes [l*C, (l+1)*C) of the tile (odd C => the 32 lanes' 16-byte reads fall in
//    distinct bank groups): k^2 advances by an integer add, the bin changes only when k^2 crosses the next host-built
//    threshold, and a run of equal bins is a private FMA chain.  The z-window weights of a lane's chunk are the same for
//    every row, so (C known at compile time) they stay in registers;
//  * FAST tiles (one mode crosses at most one threshold; no lane closes more than three runs -- all but the low-k corner):
//    the walk has no branch and no dependent load; a finished run is pushed by one predicated store onto a small queue
//    (run i of a lane is bin fbin+i);  GENERAL tiles update the bins from inside the walk;
//  * after the walk the queues are drained into the warp-private bin array.  Bins are monotone along a row, so the
//    queued runs of different lanes hit disjoint bins (plain read-modify-write, fixed order), except the FIRST run of a
//    lane, which may share its bin with the runs of the lanes before it: those 32 partial sums are combined by one
//    segmented warp scan per tile.
#ifndef KSN_K1_WIN_DEFAULT
#define KSN_K1_WIN_DEFAULT 1              // see k1_tile_config()
#endif
constexpr int K1T_MAXW = 16;
constexpr int K1T_QRUNS = 8;              // queue slots per lane: up to seven closed runs and the open one

__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    unsigned done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

// queue accesses bypass the compiler's alias analysis on purpose (no "memory" clobber): the queue is a region of its
// own, and the walk's loads must stay free to move ahead of these stores
__device__ __forceinline__ void queue_push_if(bool ch, unsigned addr, double v)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t@p st.shared.f64 [%1], %2;\n\t}" ::"r"((unsigned) ch), "r"(addr), "d"(v));
}
__device__ __forceinline__ void queue_put(unsigned addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v)); }
__device__ __forceinline__ double queue_get(unsigned addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

// CT > 0: modes per lane known at compile time (walk fully unrolled, weights in registers); CT == 0: run-time C.
// WIN (bin window): at PMGRID = 4096 a warp's private copy of all 2048 bins (16 KB) leaves shared memory for six warps
// only.  Bins are logarithmic, so all but a sliver of the rows (those within ~60 grid units of the k_x = k_y = 0 axis at
// 4096) touch only the upper part of the bin range: with WIN a warp keeps the bins >= hot_lo in shared memory and the
// rarely used ones below in a global-memory array of its own (`cold`, read and written through L1 by this warp only, in
// the same fixed order), which makes room for eight warps again.
//
// real = float: a float row is (N/2+1)*8 bytes, so every other row -- and with it every tile of that row -- starts 8 bytes off the
// 16-byte granule of a bulk copy.  Such a tile is copied from one mode earlier (`sh` = 1) and to an even mode count; the
// lanes read their chunks `sh` modes into the stage, and the one or two foreign modes at the ends are either never read
// or walked with weight 0 like everything past a row's end.  |F|^2 is formed in float as the reference's float build does
// (powerspectrum.c:68 with fftw_real = float), the window stays the separable double one (float-grid tolerance 1e-5).
template <typename real, int CT, int WIN>
__global__ void __launch_bounds__((CT == 5 || CT == 9 ? K1T_MAXW : 8) * 32, 1)
k1_tile_kernel(const Cplx<real> *__restrict__ grid, int nrows, int N, int nrbins, long long plane0, float binscale,
               const unsigned *__restrict__ thr, const double *__restrict__ iw, double *__restrict__ partial, int accumulate,
               int Crt, int T, int S, int stage_bytes, unsigned k2_single, int log2N, int hot_lo, double *__restrict__ cold)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int C = CT > 0 ? CT : Crt;
    const int L = N / 2 + 1;
    const int W = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int TE = 32 * C;                                     // modes per tile
    const double *__restrict__ iwz_g = iw + L;                 // m_z * iwz^4 (Hermitian multiplicity folded in), zeros past the row end
    unsigned char *sp = smem_raw;
    unsigned char *stages = sp;                   sp += (size_t) W * S * stage_bytes;
    const int nhot = WIN ? nrbins - hot_lo : nrbins;           // bins a warp keeps in shared memory
    double *bins_s = (double *) sp;               sp += (size_t) W * nhot * sizeof(double);
    double *queue_s = (double *) sp;              sp += (size_t) W * K1T_QRUNS * 32 * sizeof(double);
    unsigned long long *bars = (unsigned long long *) sp;  sp += (size_t) W * S * sizeof(unsigned long long);
    unsigned *thr_s = (unsigned *) sp;                      // nrbins + 3, padded with "never"
    for (int i = threadIdx.x; i < nrbins + 3; i += blockDim.x) thr_s[i] = i < nrbins ? thr[i] : 0xffffffffu;
    for (int i = threadIdx.x; i < W * nhot; i += blockDim.x) bins_s[i] = 0.0;
    double *mycold = WIN ? cold + ((size_t) blockIdx.x * W + warp) * hot_lo : nullptr;
    if (WIN) for (int i = lane; i < hot_lo; i += 32) mycold[i] = 0.0;
    // the stages start as zeros: modes past a row's end are walked like any other (weight 0), so they must be finite
    for (int i = threadIdx.x; i < W * S * stage_bytes / 16; i += blockDim.x) ((double2 *) stages)[i] = make_double2(0.0, 0.0);
    if (lane == 0) {
        for (int s = 0; s < S; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bars + warp * S + s)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the zeros above, before any bulk copy lands on them
    __syncthreads();

    const long long first = (long long) blockIdx.x * W + warp, stride = (long long) gridDim.x * W;
    const int dr = (int) (stride / T), dt = (int) (stride - (long long) dr * T);     // a step of `stride` tiles in (row, tile-of-row)
    unsigned char *mystage = stages + (size_t) warp * S * stage_bytes;
    const unsigned bar0 = smem_addr(bars + warp * S);
    double *mybins = bins_s + (size_t) warp * nhot;
    auto binp = [&](int idx) -> double * {          // where this warp keeps bin idx
        if (WIN) return idx >= hot_lo ? mybins + (idx - hot_lo) : mycold + idx;
        return mybins + idx;
    };
    const unsigned q0 = smem_addr(queue_s + (size_t) warp * K1T_QRUNS * 32 + lane);   // run i of this lane: q0 + 256 i
    const unsigned le_mask = 0xffffffffu >> (31 - lane);

    auto issue = [&](int r, int t, int s) {         // lane 0: bulk copy of tile t of row r into stage s
        const int z0 = t * TE, len = min(TE, L - z0);
        unsigned bytes = (unsigned) len * 16u;
        const unsigned bar = bar0 + 8u * s, dst = smem_addr(mystage + (size_t) s * stage_bytes);
        const Cplx<real> *src = grid + (size_t) r * L + z0;
        if constexpr (sizeof(real) == 4) {
            const unsigned sh = (unsigned) (((size_t) r * L + z0) & 1);    // slab-relative mode index of the tile's first mode: odd?
            bytes = (((unsigned) len + sh + 1u) & ~1u) * 8u;               // (the host took this kernel only for an even mode total)
            src -= sh;
        }
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
    };
    // consumer position (r, t) and producer position (ri, ti) = S tiles ahead
    int r = (int) (first / T), t = (int) (first - (long long) r * T);
    int ri = r, ti = t;
    for (int s = 0; s < S; s++) {
        if (lane == 0 && ri < nrows) issue(ri, ti, s);
        ri += dr; ti += dt;
        if (ti >= T) { ti -= T; ri++; }
    }

    // A lane's FIRST run may share its bin with the runs of the lanes before it (everything else hits bins no other lane
    // touches, because bins are monotone along a row): one segmented scan over the 32 first runs per tile.  The scan of
    // tile k is issued together with the set-up arithmetic of tile k+1, so that the two latency chains overlap.
    int p_fbin = 0x7fffffff;               // previous tile's first-run bin / sum / row factor (none yet)
    double p_facc = 0.0, p_wxy = 0.0;
    auto merge_first_runs = [&]() {
        double v1[1] = { p_facc };
        const unsigned tails = segmented_sum<1>(p_fbin, v1, lane, le_mask);
        if (((tails >> lane) & 1u) && p_fbin != 0x7fffffff) { double *bp = binp(p_fbin); *bp = fma(v1[0], p_wxy, *bp); }
        __syncwarp();
    };
    double wreg[CT > 0 ? CT : 1];
    int t_loaded = -1;
    unsigned phases = 0;
    int s = 0;
    while (r < nrows) {
        // row geometry (warp-uniform); the two window loads are consumed only when the runs are drained
        const int pl = log2N >= 0 ? r >> log2N : r / N, j = r - pl * N;
        const long long gi = plane0 + pl;
        const int ki = gi <= N / 2 ? (int) gi : (int) (gi - N);
        const int kj = j <= N / 2 ? j : j - N;
        const int c = ki * ki + kj * kj;
        const double wa = __ldg(iw + (ki < 0 ? -ki : ki)), wb = __ldg(iw + (kj < 0 ? -kj : kj));
        const int z0 = t * TE, zl = z0 + lane * C;             // first z of the tile / of this lane's chunk
        unsigned k2 = (unsigned) c + (unsigned) zl * (unsigned) zl, dz = 2u * (unsigned) zl + 1u;
        int b = max((int) (binscale * fast_log2((float) k2)), 0);      // within one bin of the truth
        b = min(b, nrbins - 1);
        b += (k2 >= thr_s[b + 1]) - (k2 < thr_s[b]);
        const int fbin = b;
        unsigned nxt = thr_s[b + 1];
        // bin of the chunk's last mode: a lane that would close more than three runs cannot use the register window
        const unsigned k2e = (unsigned) c + (unsigned) (zl + C - 1) * (unsigned) (zl + C - 1);
        int be = min(max((int) (binscale * fast_log2((float) k2e)), 0), nrbins - 1);
        be += (k2e >= thr_s[be + 1]) - (k2e < thr_s[be]);
        const int spread = __reduce_max_sync(0xffffffffu, be - fbin);      // most runs any lane will close in this tile
        const bool single = (unsigned) c + (unsigned) z0 * (unsigned) z0 >= k2_single;
        merge_first_runs();                                            // of the previous tile
        const Cplx<real> *chunk = (const Cplx<real> *) (mystage + (size_t) s * stage_bytes) + lane * C;
        if constexpr (sizeof(real) == 4) chunk += ((size_t) r * L + z0) & 1;     // the copy started one mode early
        const double *wz = iwz_g + zl;
        if (CT > 0 && t != t_loaded) {             // (a warp keeps the same tile-of-row whenever T divides its stride)
#pragma unroll
            for (int e = 0; e < (CT > 0 ? CT : 1); e++) wreg[e] = __ldg(wz + e);
            t_loaded = t;
        }
        const double worigin = (c == 0 && zl == 0) ? 0.0 : 1.0;        // F(0,0,0) is the mean, not a mode: weight 0
        mbar_wait(bar0 + 8u * s, (phases >> s) & 1u);
        phases ^= 1u << s;
        const double q1 = wa * wb, q2 = q1 * q1, wxy = q2 * q2;        // (iwx iwy)^4, applied once per run

        double acc = 0.0, facc;
        // FAST tiles: one mode crosses at most one bin threshold (so run i of a lane is bin fbin+i) and no lane closes more
        // than R-1 runs (so the thresholds it will meet sit in R-1 registers): the walk has no branch and no dependent
        // load; a closed run is pushed onto the lane's queue by one predicated store.  R = 4 for most tiles, 8 near the
        // k_x = k_y = 0 axis where bins are narrow along z.
        auto fast_walk = [&](auto RT, auto bp_of) {
            constexpr int R = decltype(RT)::value;
            unsigned nx[R - 1];
            nx[0] = nxt;
#pragma unroll
            for (int i = 1; i < R - 1; i++) nx[i] = thr_s[min(b + 1 + i, nrbins + 2)];
            unsigned qa = q0;
            auto step = [&](const Cplx<real> v, double w) {
                double pp;
                if constexpr (sizeof(real) == 8) pp = fma(v.im, v.im, v.re * v.re); else pp = (double) fmaf(v.im, v.im, v.re * v.re);
                const bool ch = k2 >= nx[0];
                queue_push_if(ch, qa, acc);
                qa += ch ? 256u : 0u;
#pragma unroll
                for (int i = 0; i < R - 2; i++) nx[i] = ch ? nx[i + 1] : nx[i];
                nx[R - 2] = ch ? 0xffffffffu : nx[R - 2];
                acc = ch ? 0.0 : acc;
                acc = fma(pp, w, acc);
                k2 += dz;
                dz += 2u;
            };
            if (CT > 0) {
                step(chunk[0], wreg[0] * worigin);
#pragma unroll
                for (int e = 1; e < (CT > 0 ? CT : 1); e++) step(chunk[e], wreg[e]);
            } else {
                Cplx<real> vc[4];
                double wc[4];
#pragma unroll
                for (int u = 0; u < 4; u++) { vc[u] = chunk[u]; wc[u] = __ldg(wz + u); }
                wc[0] *= worigin;
#pragma unroll 2
                for (int e = 0; e + 4 < C; e += 4) {           // C = 4q+1: q blocks of four, then one mode
                    Cplx<real> vn[4];
                    double wn[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) { vn[u] = chunk[e + 4 + u]; wn[u] = __ldg(wz + e + 4 + u); }
#pragma unroll
                    for (int u = 0; u < 4; u++) step(vc[u], wc[u]);
#pragma unroll
                    for (int u = 0; u < 4; u++) { vc[u] = vn[u]; wc[u] = wn[u]; }
                }
                step(vc[0], wc[0]);
            }
            queue_put(qa, acc);                                 // the open run
            // the stage is consumed: put the next bulk copy in flight before the bookkeeping below
            __syncwarp();
            if (lane == 0 && ri < nrows) issue(ri, ti, s);
            // drain: run i >= 1 of this lane is bin fbin+i, and nobody else's (loads first, then the updates)
            const int closed = (int) ((qa - q0) >> 8);
            double g[R - 1], m[R - 1];
#pragma unroll
            for (int i = 1; i < R; i++)
                if (i <= closed) { g[i - 1] = queue_get(q0 + 256u * i); m[i - 1] = *bp_of(fbin + i); }
#pragma unroll
            for (int i = 1; i < R; i++)
                if (i <= closed) *bp_of(fbin + i) = fma(g[i - 1], wxy, m[i - 1]);
            facc = queue_get(q0);
        };
        {
            if (single && spread <= 3) {
                fast_walk(std::integral_constant<int, 4>(), binp);
            } else if (single && spread <= K1T_QRUNS - 1) {
                fast_walk(std::integral_constant<int, K1T_QRUNS>(), binp);
            } else {
                // GENERAL tile (the low-k corner, where bins are narrower than a step in k^2): finished runs go straight to the bins
                facc = 0.0;
                auto step = [&](const Cplx<real> v, double w) {
                    double pp;
                    if constexpr (sizeof(real) == 8) pp = fma(v.im, v.im, v.re * v.re); else pp = (double) fmaf(v.im, v.im, v.re * v.re);
                    if (k2 >= nxt) {
                        if (b == fbin) facc = acc; else { double *bp = binp(b); *bp = fma(acc, wxy, *bp); }
                        acc = 0.0;
                        do { b++; nxt = thr_s[b + 1]; } while (k2 >= nxt);
                    }
                    acc = fma(pp, w, acc);
                    k2 += dz;
                    dz += 2u;
                };
                step(chunk[0], __ldg(wz) * worigin);
#pragma unroll 1
                for (int e = 1; e < C; e++) step(chunk[e], __ldg(wz + e));
                if (b == fbin) facc = acc; else { double *bp = binp(b); *bp = fma(acc, wxy, *bp); }
                __syncwarp();
                if (lane == 0 && ri < nrows) issue(ri, ti, s);
            }
        }
        __syncwarp();                                                  // every lane is done with the bins
        p_fbin = fbin; p_facc = facc; p_wxy = wxy;                     // merged at the top of the next tile
        ri += dr; ti += dt;
        if (ti >= T) { ti -= T; ri++; }
        r += dr; t += dt;
        if (t >= T) { t -= T; r++; }
        s = s + 1 == S ? 0 : s + 1;
    }
    merge_first_runs();                                                // of the last tile
    __syncthreads();
    double *out = partial + (size_t) blockIdx.x * nrbins;
    for (int i = threadIdx.x; i < nrbins; i += blockDim.x) {
        double sum = 0.0;
        for (int w = 0; w < W; w++)
            sum += !WIN ? bins_s[(size_t) w * nrbins + i]
                        : (i >= hot_lo ? bins_s[(size_t) w * nhot + (i - hot_lo)] : cold[((size_t) blockIdx.x * W + w) * hot_lo + i]);
        out[i] = accumulate ? out[i] + sum : sum;
    }
}

static unsigned g_k1_k2_single = 0xffffffffu;
static char g_k1_last[192] = "none";

struct K1TileCfg { int W, C, S, T, stage_bytes, hot_lo; size_t smem, inflight; };

// nhot: bins a warp keeps in shared memory (all nrbins of them without the bin window)
static size_t k1_tile_smem(int L, int nrbins, int nhot, int W, int C, int S, int *T_out, int *stage_out, int esz = 16)
{
    const int TE = 32 * C, T = (L + TE - 1) / TE;
    const int stage = (int) ((((size_t) TE + 8) * esz + 127) & ~(size_t) 127);      // esz: bytes per mode (16 double, 8 float)
    if (T_out) *T_out = T;
    if (stage_out) *stage_out = stage;
    return (size_t) W * S * stage + (size_t) W * nhot * 8 + (size_t) W * K1T_QRUNS * 32 * 8 + (size_t) W * S * 8 +
           (size_t) (nrbins + 3 + K1T_QRUNS) * 4 + 128;
}

static int k1_tile_max_warps(int C) { return (C == 5 || C == 9) ? K1T_MAXW : 8; }      // the kernels' launch bounds

// Pick (warps, chunk, stages).  The walk is bound by per-warp latency, not by issue slots or (yet) by HBM, so throughput
// goes as warps x C/(C+10) (the per-tile bookkeeping costs about ten modes' worth) -- fitted to probes at 1024^3, 2048^3
// and 4096^3 slabs (tools/k1_tile_probe.py); shared memory (bins: 8 nrbins bytes per warp) decides how many warps fit.
// KSN_K1_TILE="W,C,S" overrides.
static bool k1_tile_config(int L, int nrbins, size_t budget, int ctas, K1TileCfg *best, int esz = 16)
{
    best->W = 0;
    best->hot_lo = 0;
    double best_score = -1;
    int fw = 0, fc = 0, fs = 0;
    const char *env = getenv("KSN_K1_TILE");
    if (env && sscanf(env, "%d,%d,%d", &fw, &fc, &fs) != 3) fw = 0;
    // bin window: "0" off, "1" where all the bins do not fit for eight warps, "2" (tests) always, with a quarter of the bins
    const char *wenv = getenv("KSN_K1_WIN");
    const int window = wenv ? atoi(wenv) : KSN_K1_WIN_DEFAULT;
    for (int C = 1; C <= 65; C += 4)
        for (int W = 4; W <= k1_tile_max_warps(C); W++)
            for (int S = 1; S <= 4; S++) {
                if (fw ? (W != fw || C != fc || S != fs) : S < 2) continue;
                int T, stage;
                size_t smem = k1_tile_smem(L, nrbins, nrbins, W, C, S, &T, &stage, esz);
                const bool compile_time = C == 5 || C == 9 || C == 13 || C == 17 || (esz == 8 && C == 33);   // 33: a whole 2048 float row is one tile
                int hot_lo = 0;
                if (smem > budget || window == 2) {
                    // bin window (k1_tile_kernel<CT, true>): only where all the bins fit for fewer than eight warps, only
                    // up to eight warps, and with at least a quarter of the bins (the upper e-fold and more) in shared memory
                    if (!window || !compile_time || C < 9 || C > 17 || W > 8) continue;
                    const size_t rest = k1_tile_smem(L, nrbins, 0, W, C, S, nullptr, nullptr, esz);
                    if (rest >= budget) continue;
                    int nhot = (int) ((budget - rest) / ((size_t) W * 8)) & ~31;
                    if (window == 2) nhot = std::min(nhot, std::max(64, (nrbins / 4) & ~31));
                    if (nhot < nrbins / 4 || nhot < 64 || nhot >= nrbins) continue;
                    hot_lo = nrbins - nhot;
                    smem = k1_tile_smem(L, nrbins, nhot, W, C, S, nullptr, nullptr, esz);
                }
                const int TE = 32 * C;
                if (T > 1 && C < 5) continue;                             // tiny tiles only for tiny rows
                const double eff = (double) L / ((double) T * TE);        // lane slots that carry a mode
                double score = W * (C / (C + 10.0)) * eff * (compile_time ? 1.0 : 0.88) * (1.0 + 0.01 * S) * (hot_lo ? 0.99 : 1.0);
                // float rows: half the bytes per mode for the same instructions -- the walk is bound by issue slots, not by
                // per-warp latency, so warps beyond eight buy nothing and the per-tile bookkeeping is what counts (measured
                // at 2048^3: 8 warps x 33 modes, a whole row per tile, 5.57 TB/s; 15 x 9: 4.10; 8 x 17: 4.52; 12 x 9: 3.76)
                if (esz == 8) score = std::min(W, 8) * (C / (C + 10.0)) * eff * (compile_time ? 1.0 : 0.88) * (1.0 + 0.01 * S) * (hot_lo ? 0.99 : 1.0) - 0.001 * W;
                if (((long long) ctas * W) % T) score *= 0.97;            // the warp's tile-of-row changes every tile: weights reloaded
                if (score > best_score) {
                    best_score = score;
                    best->W = W; best->C = C; best->S = S; best->T = T; best->stage_bytes = stage; best->smem = smem;
                    best->hot_lo = hot_lo;
                    best->inflight = (size_t) W * S * TE * esz;
                }
            }
    return best->W > 0;
}

}  // namespace ksn

// Which tile shape the K1 tile kernel would take for (dims, nrbins) on a device with `smem_budget` bytes of opt-in shared
// memory per CTA and `sms` SMs -- pure host arithmetic (no device needed), exported so that the choice can be pinned by a
// CPU test.  Returns 1 and fills warps / modes per lane / stages / tiles per row / first bin kept in shared memory
// (0 = all of them), or 0 when no tile shape fits (the scan-based kernel runs then).
extern "C" int ksn_k1_tile_plan_ex(int dims, int nrbins, size_t smem_budget, int sms, int real_bytes, int *warps, int *chunk, int *stages, int *tiles_per_row, int *hot_lo)
{
    ksn::K1TileCfg tc;
    if (dims < 2 || nrbins < 1 || (real_bytes != 4 && real_bytes != 8) ||
        !ksn::k1_tile_config(dims / 2 + 1, nrbins, smem_budget, sms, &tc, 2 * real_bytes)) return 0;
    if (warps) *warps = tc.W;
    if (chunk) *chunk = tc.C;
    if (stages) *stages = tc.S;
    if (tiles_per_row) *tiles_per_row = tc.T;
    if (hot_lo) *hot_lo = tc.hot_lo;
    return 1;
}

extern "C" int ksn_k1_tile_plan(int dims, int nrbins, size_t smem_budget, int sms, int *warps, int *chunk, int *stages, int *tiles_per_row, int *hot_lo)
{
    return ksn_k1_tile_plan_ex(dims, nrbins, smem_budget, sms, 8, warps, chunk, stages, tiles_per_row, hot_lo);
}

namespace ksn {

template <typename real, bool FULL>
__global__ void __launch_bounds__(K1_MAX_WARPS * 32, 1)
k1_bin_kernel(const Cplx<real> *__restrict__ grid, int nrows, int N, int nrbins, long long plane0,
              float binscale, const unsigned *__restrict__ thr, const double *__restrict__ iw,
              double *__restrict__ partial, int accumulate)
{
    constexpr int NV = FULL ? 3 : 1;
    constexpr int U = FULL ? 4 : K1_UNROLL;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int L = N / 2 + 1;
    const int nyq = N / 2;
    const int nwarps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *iw_s = (double *) smem_raw;                       // 2L: z table (with multiplicity) | plain table
    double *bins_s = iw_s + 2 * L;                            // nwarps * NV * nrbins
    uint2 *thr_s = (uint2 *) (bins_s + (size_t) nwarps * NV * nrbins);   // nrbins+1 pairs {thr[b], thr[b+1]}

    // z-window with the Hermitian multiplicity folded in as a fourth root (the weight is the 4th power)
    const double root2_4 = sizeof(real) == 4 ? (double) 1.189207115002721f : 1.189207115002721;
    for (int i = threadIdx.x; i < L; i += blockDim.x) iw_s[L + i] = iw[i];          // plain table: x/y factors
    for (int i = threadIdx.x; i < L; i += blockDim.x) iw_s[i] = (i == 0 || i == nyq) ? iw[i] : iw[i] * root2_4;
    for (int i = threadIdx.x; i <= nrbins; i += blockDim.x)
        thr_s[i] = make_uint2(i < nrbins ? thr[i] : 0xffffffffu, i + 1 < nrbins ? thr[i + 1] : 0xffffffffu);
    for (int i = threadIdx.x; i < nwarps * NV * nrbins; i += blockDim.x) bins_s[i] = 0.0;
    __syncthreads();

    double *mybins = bins_s + (size_t) warp * NV * nrbins;
    const int nfull_groups = (L / 32) / U;                     // groups of U steps that lie entirely inside the row
    const int nsteps = (L + 31) / 32;                          // the rest of the row goes one (masked) step at a time
    // rows are dealt to warps round-robin: the 16 warps of a CTA sweep 16 consecutive rows
    for (int r = blockIdx.x * nwarps + warp; r < nrows; r += gridDim.x * nwarps) {
        const int pl = r / N, j = r - pl * N;
        const long long gi = plane0 + pl;
        const int ki = gi <= N / 2 ? (int) gi : (int) (gi - N);
        const int kj = j <= N / 2 ? j : j - N;
        const int c = ki * ki + kj * kj;
        double wxy;
        {
            const double a = iw_s[L + (ki < 0 ? -ki : ki)], b = iw_s[L + (kj < 0 ? -kj : kj)];
            wxy = sizeof(real) == 4 ? (double) ((float) a * (float) b) : a * b;
        }
        const Cplx<real> *rowptr = grid + (size_t) r * L;
        const bool origin_row = c == 0;
        int g = 0;
#pragma unroll 1
        for (; g < nfull_groups; g++)
            k1_group<real, FULL, U, false>(rowptr, lane + g * 32 * U, L, nyq, c, wxy, binscale, nrbins, thr_s, iw_s, mybins, lane, origin_row);
#pragma unroll 1
        for (int st = g * U; st < nsteps; st++)
            k1_group<real, FULL, 1, true>(rowptr, lane + st * 32, L, nyq, c, wxy, binscale, nrbins, thr_s, iw_s, mybins, lane, origin_row);
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
    // one warp per output value: lane l adds the partials of CTAs l, l+32, ... in that order, then a fixed xor tree
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i < nv * nrbins) {
        double s = 0.0;
        for (int c = lane; c < ctas; c += 32) s += partial[(size_t) c * nv * nrbins + i];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        const int which = i / nrbins, b = i - which * nrbins;
        if (lane == 0) red[which == 0 ? b : which * nrbins + 1 + b] = s;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double m2 = 0.0;
        if (origin) {   // powerspectrum.c:45-47: only the rank holding plane 0
            const double re = (double) origin->re, im = (double) origin->im;
            m2 = re * re + im * im;
        }
        red[nrbins] = m2;
    }
}

// The same reduction FUSED with the cross-rank sum (peer-memory backend, ksn_p2p.cuh): every value goes straight from the
// warp that produced it into all ranks' mailboxes over NVLink, so the transfer overlaps the rest of the reduction; the
// last block to finish raises this rank's flags, waits for the other ranks' and adds the R contributions in rank order.
// One kernel instead of k1_final_kernel + ncclAllReduce; red = the GLOBAL sums, bit-identical on every rank.
template <typename real>
__global__ void __launch_bounds__(256)
k1_final_p2p_kernel(const double *__restrict__ partial, int ctas, int nv, int nrbins,
                    const Cplx<real> *origin, double *__restrict__ red, const P2PDev p)
{
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i < nv * nrbins) {
        double s = 0.0;
        for (int c = lane; c < ctas; c += 32) s += partial[(size_t) c * nv * nrbins + i];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        const int which = i / nrbins, b = i - which * nrbins;
        if (lane == 0) p2p_push(p, which == 0 ? b : which * nrbins + 1 + b, s);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double m2 = 0.0;
        if (origin) {   // powerspectrum.c:45-47: only the rank holding plane 0
            const double re = (double) origin->re, im = (double) origin->im;
            m2 = re * re + im * im;
        }
        p2p_push(p, nrbins, m2);
    }
    if (!p2p_last_block(p)) return;
    p2p_finish(p, (size_t) nv * nrbins + 1, red);
}

static size_t k1_smem_bytes(int dims, int nrbins, int nwarps, int nv)
{
    return (size_t) (dims / 2 + 1) * 16 + (size_t) nwarps * nv * nrbins * 8 + (size_t) (nrbins + 1) * 8 + 16;
}

template <typename real, bool FULL>
static int k1_launch_t(const void *dgrid, int dims, int nrbins, long long plane0, long long nplanes,
                       bool accumulate, int *ctas_out, int *stride_out)
{
    Ctx &c = ctx();
    constexpr int NV = FULL ? 3 : 1;
    if (nplanes * dims > 0x7fffffffLL) return set_error(KSN_EINVAL, "K1: %lld rows in one slab", nplanes * dims);
    const int nrows = (int) (nplanes * dims);
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
        { const int rca = func_attributes((const void *) kern, smem, -1); if (rca) return rca; }
        kern<<<ctas, nwarps * 32, smem, c.stream>>>((const Cplx<real> *) dgrid, nrows, dims, nrbins, plane0, binscale,
                                                    c.d_thr, c.d_iw, c.d_partial, accumulate ? 1 : 0);
        c.launches++;
        KSN_CUDA(cudaGetLastError());
        return KSN_OK;
    };
    K1TileCfg tc;
    // float rows through the tile kernel need an even number of modes in the slab (the last bulk copy is rounded up to a
    // mode pair); the scan-based kernel takes the rest
    const bool tile_ok = sizeof(real) == 8 || ((long long) nrows * (dims / 2 + 1)) % 2 == 0;
    if (!FULL && tile_ok && !getenv("KSN_K1_PAIR") && !getenv("KSN_K1_NOPAIR") && ((uintptr_t) dgrid & 15) == 0 &&
        k1_tile_config(dims / 2 + 1, nrbins, c.smem_optin, c.num_sms, &tc, (int) (2 * sizeof(real)))) {
        int log2N = -1;
        if ((dims & (dims - 1)) == 0) { log2N = 0; while ((1 << log2N) < dims) log2N++; }
        if (tc.hot_lo) {                          // per-warp bins below the window (bin window only)
            rc = ensure_device_buffer((void **) &c.d_cold, &c.cold_cap, (size_t) ctas * tc.W * tc.hot_lo * sizeof(double));
            if (rc) return rc;
        }
        auto go = [&](auto kern) -> int {
            { const int rca = func_attributes((const void *) kern, tc.smem, -1); if (rca) return rca; }
            kern<<<ctas, tc.W * 32, tc.smem, c.stream>>>((const Cplx<real> *) dgrid, nrows, dims, nrbins, plane0, binscale,
                                                         c.d_thr, c.d_iw, c.d_partial, accumulate ? 1 : 0, tc.C, tc.T, tc.S, tc.stage_bytes,
                                                         g_k1_k2_single, log2N, tc.hot_lo, c.d_cold);
            return KSN_OK;
        };
        int rct;
        switch (tc.hot_lo ? -tc.C : tc.C) {
        case 5: rct = go(k1_tile_kernel<real, 5, 0>); break;
        case 9: rct = go(k1_tile_kernel<real, 9, 0>); break;
        case 13: rct = go(k1_tile_kernel<real, 13, 0>); break;
        case 17: rct = go(k1_tile_kernel<real, 17, 0>); break;
        case 33:
            if constexpr (sizeof(real) == 4) rct = go(k1_tile_kernel<real, 33, 0>);
            else rct = go(k1_tile_kernel<real, 0, 0>);
            break;
        case -9: rct = go(k1_tile_kernel<real, 9, 1>); break;
        case -13: rct = go(k1_tile_kernel<real, 13, 1>); break;
        case -17: rct = go(k1_tile_kernel<real, 17, 1>); break;
        default: rct = go(k1_tile_kernel<real, 0, 0>); break;
        }
        if (rct) return rct;
        c.launches++;
        KSN_CUDA(cudaGetLastError());
        if (tc.hot_lo)
            snprintf(g_k1_last, sizeof g_k1_last, "k1_tile_kernel%s (%d warps x %d modes per lane, %d stages, %d tiles per row, bins >= %d of %d in shared memory)",
                     sizeof(real) == 4 ? "<float>" : "", tc.W, tc.C, tc.S, tc.T, tc.hot_lo, nrbins);
        else
            snprintf(g_k1_last, sizeof g_k1_last, "k1_tile_kernel%s (%d warps x %d modes per lane, %d stages, %d tiles per row)",
                     sizeof(real) == 4 ? "<float>" : "", tc.W, tc.C, tc.S, tc.T);
        *ctas_out = ctas;
        *stride_out = NV * nrbins;
        return KSN_OK;
    }
    snprintf(g_k1_last, sizeof g_k1_last, "%s<%s>", FULL || getenv("KSN_K1_NOPAIR") ? "k1_bin_kernel" : "k1_pair_kernel", sizeof(real) == 8 ? "double" : "float");
    if (FULL || getenv("KSN_K1_NOPAIR")) rc = launch(k1_bin_kernel<real, FULL>);
    else rc = launch(k1_pair_kernel<real, 16, 4>);
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

int k1_finish(int real_bytes, int dims, int nrbins, bool full, int ctas, int stride, const void *origin_elem, bool fuse_p2p)
{
    (void) dims; (void) stride;
    Ctx &c = ctx();
    const int nv = full ? 3 : 1;
    const int threads = 256, blocks = (nv * nrbins * 32 + threads - 1) / threads;
    if (fuse_p2p) {
        P2PDev dev;
        int rc = p2p_next_round(&dev);
        if (rc) return rc;
        if (real_bytes == 8)
            k1_final_p2p_kernel<double><<<blocks, threads, 0, c.stream>>>(c.d_partial, ctas, nv, nrbins, (const Cplx<double> *) origin_elem, c.d_red, dev);
        else
            k1_final_p2p_kernel<float><<<blocks, threads, 0, c.stream>>>(c.d_partial, ctas, nv, nrbins, (const Cplx<float> *) origin_elem, c.d_red, dev);
        c.launches++;
        KSN_CUDA(cudaGetLastError());
        return KSN_OK;
    }
    if (real_bytes == 8)
        k1_final_kernel<double><<<blocks, threads, 0, c.stream>>>(c.d_partial, ctas, nv, nrbins, (const Cplx<double> *) origin_elem, c.d_red);
    else
        k1_final_kernel<float><<<blocks, threads, 0, c.stream>>>(c.d_partial, ctas, nv, nrbins, (const Cplx<float> *) origin_elem, c.d_red);
    c.launches++;
    KSN_CUDA(cudaGetLastError());
    return KSN_OK;
}

// The tables depend on (dims, nrbins) only, and a PM run passes the same ones every step: remember what sits in
// c.d_thr / c.d_iw (by content hash, not by host address) and skip the three uploads and the stream sync they cost.
struct K1TabKey { bool valid; int dims, nrbins; unsigned long long h_thr, h_iw; const void *d_thr, *d_iw; };
static K1TabKey g_k1_tab = { false, 0, 0, 0, 0, nullptr, nullptr };
void k1_tables_invalidate() { g_k1_tab.valid = false; }

static unsigned long long hash_words(const void *p, size_t bytes)
{
    const unsigned char *b = (const unsigned char *) p;
    unsigned long long h = 0x9E3779B97F4A7C15ull ^ bytes;
    size_t i = 0;
    for (; i + 8 <= bytes; i += 8) {
        unsigned long long w;
        memcpy(&w, b + i, 8);
        h = (h ^ w) * 0x9E3779B97F4A7C15ull;
        h ^= h >> 29;
    }
    for (; i < bytes; i++) { h = (h ^ b[i]) * 0x100000001B3ull; }
    return h;
}

static int k1_upload_tables(int dims, int nrbins, const unsigned int *thresholds, const double *invwin)
{
    // From which k^2 on can one step along z (k^2 -> k^2 + 2z+1 <= k^2 + 2 sqrt(k^2) + 1) cross at most ONE bin threshold?
    // J = one past the last pair of thresholds closer than such a step; a tile that starts at or above thr[J] is "fast".
    {
        int J = 0;
        for (int j = 0; j + 1 < nrbins; j++)
            if ((double) thresholds[j + 1] - (double) thresholds[j] < 2.0 * sqrt((double) thresholds[j + 1]) + 2.0) J = j + 2;
        g_k1_k2_single = J < nrbins ? thresholds[J] : 0xffffffffu;
    }
    Ctx &c = ctx();
    const int L = dims / 2 + 1;
    int rc = ensure_device_buffer((void **) &c.d_thr, &c.thr_cap, (size_t) nrbins * sizeof(unsigned));
    if (rc) return rc;
    // d_iw = [ iw[0..L) | m_z iw[z]^4 for z in [0, L), then K1_WZ_PAD zeros (lanes walk past the row end with weight 0) ]
    rc = ensure_device_buffer((void **) &c.d_iw, &c.iw_cap, (size_t) (2 * L + K1_WZ_PAD) * sizeof(double));
    if (rc) return rc;
    rc = ensure_device_buffer((void **) &c.d_red, &c.red_cap, (size_t) (3 * nrbins + 1) * sizeof(double));
    if (rc) return rc;
    rc = ensure_pinned_buffer((void **) &c.h_red, &c.h_red_cap, (size_t) (3 * nrbins + 1) * sizeof(double));
    if (rc) return rc;
    const unsigned long long h_thr = hash_words(thresholds, (size_t) nrbins * sizeof(unsigned));
    const unsigned long long h_iw = hash_words(invwin, (size_t) L * sizeof(double));
    if (g_k1_tab.valid && g_k1_tab.dims == dims && g_k1_tab.nrbins == nrbins && g_k1_tab.h_thr == h_thr && g_k1_tab.h_iw == h_iw &&
        g_k1_tab.d_thr == c.d_thr && g_k1_tab.d_iw == c.d_iw)
        return KSN_OK;
    g_k1_tab.valid = false;
    {
        double *wz = (double *) calloc((size_t) L + K1_WZ_PAD, sizeof(double));
        if (!wz) return set_error(KSN_ENOMEM, "K1: out of host memory");
        for (int z = 0; z < L; z++) {
            const double w2 = invwin[z] * invwin[z];
            wz[z] = ((z == 0 || z == dims / 2) ? 1.0 : 2.0) * (w2 * w2);      // Hermitian multiplicity folded in
        }
        const cudaError_t e = cudaMemcpyAsync(c.d_iw + L, wz, ((size_t) L + K1_WZ_PAD) * sizeof(double), cudaMemcpyHostToDevice, c.stream);
        if (e == cudaSuccess) cudaStreamSynchronize(c.stream);
        free(wz);
        KSN_CUDA(e);
    }
    KSN_CUDA(cudaMemcpyAsync(c.d_thr, thresholds, (size_t) nrbins * sizeof(unsigned), cudaMemcpyHostToDevice, c.stream));
    KSN_CUDA(cudaMemcpyAsync(c.d_iw, invwin, (size_t) L * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    g_k1_tab = { true, dims, nrbins, h_thr, h_iw, c.d_thr, c.d_iw };
    return KSN_OK;
}

// Plane-chunked staging of a host-resident slab into c.d_stage, K1 on each chunk as it lands.  Resident plan: chunk i
// lives at its own offset and stays for K3.  Streaming plan (the slab does not fit): chunk i lands in ring slot i % 3,
// which is free again once K1 has read it.
static int k1_over_host_grid(const void *hgrid, int real_bytes, int dims, int nrbins, long long startslab, long long nslab,
                             bool full, int *ctas, int *stride, const void **origin)
{
    Ctx &c = ctx();
    StagePlan pl;
    int rc = stage_plan(real_bytes, dims, nslab, &pl);
    if (rc) return rc;
    c.stage_streaming = pl.streaming;
    ensure_host_pinned(hgrid, pl.total);
    const int nev = pl.streaming ? STAGE_RING : pl.nchunks;
    cudaEvent_t *landed = (cudaEvent_t *) malloc(sizeof(cudaEvent_t) * nev), *read = (cudaEvent_t *) malloc(sizeof(cudaEvent_t) * nev);
    for (int i = 0; i < nev; i++) {
        cudaEventCreateWithFlags(&landed[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&read[i], cudaEventDisableTiming);
    }
    // make sure the staging buffer is not still being read by a previous step
    KSN_CUDA(cudaStreamSynchronize(c.stream));
    *origin = pl.streaming ? c.d_origin : c.d_stage;
    phase_begin(PH_H2D);
    phase_begin(PH_K1);
    for (int i = 0; i < pl.nchunks && !rc; i++) {
        const long long p0 = i * pl.chunk, np = (p0 + pl.chunk <= nslab) ? pl.chunk : nslab - p0;
        const int s = pl.streaming ? i % STAGE_RING : i;
        char *slot = (char *) c.d_stage + (pl.streaming ? (size_t) s * pl.chunk : (size_t) p0) * pl.plane_bytes;
        if (pl.streaming && i >= STAGE_RING) cudaStreamWaitEvent(c.copy_stream, read[s], 0);
        cudaMemcpyAsync(slot, (const char *) hgrid + p0 * pl.plane_bytes, np * pl.plane_bytes, cudaMemcpyHostToDevice, c.copy_stream);
        cudaEventRecord(landed[s], c.copy_stream);
        cudaStreamWaitEvent(c.stream, landed[s], 0);
        if (pl.streaming && i == 0) cudaMemcpyAsync(c.d_origin, slot, 2 * real_bytes, cudaMemcpyDeviceToDevice, c.stream);
        rc = k1_launch(slot, real_bytes, dims, nrbins, startslab + p0, np, full, i > 0, ctas, stride);
        cudaEventRecord(read[s], c.stream);
    }
    phase_end(PH_H2D);
    phase_end(PH_K1);
    for (int i = 0; i < nev; i++) { cudaEventDestroy(landed[i]); cudaEventDestroy(read[i]); }
    free(landed);
    free(read);
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
    // keff and count sums depend on the thresholds' CONTENT too (a C-ABI caller may pass other bin edges for the same shape)
    const unsigned long long h_thr = g_k1_tab.h_thr;
    const bool have_geom = cache_ok && c.geom.valid && c.geom.dims == dims && c.geom.nrbins == nrbins && c.geom.h_thr == h_thr &&
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
            rc = k1_over_host_grid(hgrid, real_bytes, dims, nrbins, startslab, nslab, full, &ctas, &stride, &origin);
        }
        if (rc) return rc;
        // K1 is in flight: now is the time for the side-stream work that wants to run beside it (K2's a-only tables)
        rc = k2_prefetch_launch_pending();
        if (rc) return rc;
    } else {
        // a rank that owns no planes still takes part in the collective with zeros
        rc = ensure_device_buffer((void **) &c.d_partial, &c.partial_cap, (size_t) 3 * nrbins * sizeof(double));
        if (rc) return rc;
        KSN_CUDA(cudaMemsetAsync(c.d_partial, 0, (size_t) 3 * nrbins * sizeof(double), c.stream));
        ctas = 1;
    }
    const size_t nred = full ? (size_t) 3 * nrbins + 1 : (size_t) nrbins + 1;
    // peer-memory backend: the cross-rank sum happens inside the final-reduce kernel (KSN_P2P_UNFUSED=1: as a kernel of
    // its own behind it, for comparison)
    const bool fuse = c.comm_kind == COMM_P2P && c.nranks > 1 && nred <= KSN_P2P_SLOT && !getenv("KSN_P2P_UNFUSED");
    phase_begin(PH_K1RED);
    rc = k1_finish(real_bytes, dims, nrbins, full, ctas, stride, (startslab == 0 && nslab > 0) ? origin : nullptr, fuse);
    phase_end(PH_K1RED);
    if (rc) return rc;
    rc = fuse ? reduced_to_host(c.d_red, c.h_red, nred) : allreduce_to_host(c.d_red, c.h_red, nred);
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
        c.geom.startslab = startslab; c.geom.nslab = nslab; c.geom.epoch = c.comm_epoch; c.geom.h_thr = h_thr;
    }
    memcpy(power_sum, c.h_red, sizeof(double) * nrbins);
    *total_mass2 = c.h_red[nrbins];
    memcpy(keff_sum, c.geom.keff, sizeof(double) * nrbins);
    memcpy(count, c.geom.count, sizeof(long long) * nrbins);
    return KSN_OK;
}

}  // namespace ksn

using namespace ksn;

extern "C" const char *ksn_last_k1_kernel(void) { return g_k1_last; }

extern "C" int ksn_powerspectrum_sums(const void *grid, int real_bytes, int dims, int nrbins,
                                      long long startslab, long long nslab,
                                      const unsigned int *thresholds, const double *invwin,
                                      double *power_sum, double *keff_sum, long long *count, double *total_mass2)
{
    int rc = ensure_init();
    if (rc) return rc;
    if ((!grid && nslab > 0) || (real_bytes != 4 && real_bytes != 8) || dims < 2 || nrbins < 2 || nslab < 0 ||
        startslab < 0 || startslab + nslab > dims || !thresholds || !invwin || !power_sum || !keff_sum || !count || !total_mass2)
        return set_error(KSN_EINVAL, "ksn_powerspectrum_sums: bad arguments (dims=%d nrbins=%d slab=[%lld,+%lld))", dims, nrbins, startslab, nslab);
    if ((double) dims * dims * 0.75 > 2.0e9) return set_error(KSN_EINVAL, "dims=%d: k^2 does not fit 31 bits", dims);
    const bool on_device = nslab > 0 && ksn_pointer_is_device(grid);
    return k1_sums(on_device ? grid : nullptr, on_device ? nullptr : grid, real_bytes, dims, nrbins, startslab, nslab,
                   thresholds, invwin, power_sum, keff_sum, count, total_mass2);
}
