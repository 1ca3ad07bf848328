// Developer probe: what can plain streaming kernels reach on this B200?  (read-only / in-place scale,
// 128- and 256-bit accesses, flat grid-stride vs the row-round-robin order K1/K3 use)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
struct __align__(32) D4 { double a, b, c, d; };
struct __align__(16) D2 { double a, b; };

__device__ __forceinline__ D4 ld256(const D4 *p) { D4 v; asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.a), "=d"(v.b), "=d"(v.c), "=d"(v.d) : "l"(p)); return v; }
__device__ __forceinline__ D4 ld256rw(const D4 *p) { D4 v; asm volatile("ld.global.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.a), "=d"(v.b), "=d"(v.c), "=d"(v.d) : "l"(p)); return v; }
__device__ __forceinline__ void st256(D4 *p, D4 v) { asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.a), "d"(v.b), "d"(v.c), "d"(v.d) : "memory"); }
__device__ __forceinline__ D2 ld128(const D2 *p) { D2 v; asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.a), "=d"(v.b) : "l"(p)); return v; }
__device__ __forceinline__ D2 ld128rw(const D2 *p) { D2 v; asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.a), "=d"(v.b) : "l"(p)); return v; }
__device__ __forceinline__ void st128(D2 *p, D2 v) { asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.a), "d"(v.b) : "memory"); }

template <int U> __global__ void read128(const D2 *p, size_t n, double *out)
{
    double s = 0;
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < n; i += U * stride) {
        D2 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) v[u] = ld128(p + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; u++) s += v[u].a + v[u].b;
    }
    if (s == 123.456) out[0] = s;
}
template <int U> __global__ void read256(const D4 *p, size_t n, double *out)
{
    double s = 0;
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < n; i += U * stride) {
        D4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) v[u] = ld256(p + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; u++) s += v[u].a + v[u].b + v[u].c + v[u].d;
    }
    if (s == 123.456) out[0] = s;
}
template <int U> __global__ void scale128(D2 *p, size_t n, double f)
{
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < n; i += U * stride) {
        D2 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) v[u] = ld128rw(p + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; u++) { v[u].a *= f; v[u].b *= f; st128(p + i + u * stride, v[u]); }
    }
}
template <int U> __global__ void scale256(D4 *p, size_t n, double f)
{
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < n; i += U * stride) {
        D4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) v[u] = ld256rw(p + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; u++) { v[u].a *= f; v[u].b *= f; v[u].c *= f; v[u].d *= f; st256(p + i + u * stride, v[u]); }
    }
}
// K1/K3 order: one warp per row of `rowlen` 16-byte elements, rows round-robin over all warps
template <int U> __global__ void scale128_rows(D2 *p, int nrows, int rowlen, double f)
{
    const int nw = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = blockIdx.x * nw + w; r < nrows; r += gridDim.x * nw) {
        D2 *row = p + (size_t) r * rowlen;
        for (int z0 = lane; z0 + 32 * (U - 1) < rowlen; z0 += 32 * U) {
            D2 v[U];
#pragma unroll
            for (int u = 0; u < U; u++) v[u] = ld128rw(row + z0 + 32 * u);
#pragma unroll
            for (int u = 0; u < U; u++) { v[u].a *= f; v[u].b *= f; st128(row + z0 + 32 * u, v[u]); }
        }
    }
}
template <int U> __global__ void read128_rows(const D2 *p, int nrows, int rowlen, double *out)
{
    double s = 0;
    const int nw = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = blockIdx.x * nw + w; r < nrows; r += gridDim.x * nw) {
        const D2 *row = p + (size_t) r * rowlen;
        for (int z0 = lane; z0 + 32 * (U - 1) < rowlen; z0 += 32 * U) {
            D2 v[U];
#pragma unroll
            for (int u = 0; u < U; u++) v[u] = ld128(row + z0 + 32 * u);
#pragma unroll
            for (int u = 0; u < U; u++) s += v[u].a + v[u].b;
        }
    }
    if (s == 123.456) out[0] = s;
}

__device__ __forceinline__ void pf_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// rows, NOT persistent: CTA b handles rows [b*RPC, (b+1)*RPC), one warp per row at a time
template <int U> __global__ void scale128_rows_np(D2 *p, int nrows, int rowlen, int rpc, double f)
{
    const int nw = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r1 = min(nrows, (blockIdx.x + 1) * rpc);
    for (int r = blockIdx.x * rpc + w; r < r1; r += nw) {
        D2 *row = p + (size_t) r * rowlen;
        for (int z0 = lane; z0 + 32 * (U - 1) < rowlen; z0 += 32 * U) {
            D2 v[U];
#pragma unroll
            for (int u = 0; u < U; u++) v[u] = ld128rw(row + z0 + 32 * u);
#pragma unroll
            for (int u = 0; u < U; u++) { v[u].a *= f; v[u].b *= f; st128(row + z0 + 32 * u, v[u]); }
        }
    }
}
// persistent rows with an L2 prefetch DIST groups ahead
template <int U, int DIST> __global__ void scale128_rows_pf(D2 *p, int nrows, int rowlen, double f)
{
    const int nw = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = blockIdx.x * nw + w; r < nrows; r += gridDim.x * nw) {
        D2 *row = p + (size_t) r * rowlen;
        for (int z0 = lane; z0 + 32 * (U - 1) < rowlen; z0 += 32 * U) {
            D2 v[U];
            if ((lane & 7) == 0) {
#pragma unroll
                for (int u = 0; u < U; u++) pf_l2(row + z0 + 32 * (u + U * DIST));
            }
#pragma unroll
            for (int u = 0; u < U; u++) v[u] = ld128rw(row + z0 + 32 * u);
#pragma unroll
            for (int u = 0; u < U; u++) { v[u].a *= f; v[u].b *= f; st128(row + z0 + 32 * u, v[u]); }
        }
    }
}
template <int U, int DIST> __global__ void read128_rows_pf(const D2 *p, int nrows, int rowlen, double *out)
{
    double s = 0;
    const int nw = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = blockIdx.x * nw + w; r < nrows; r += gridDim.x * nw) {
        const D2 *row = p + (size_t) r * rowlen;
        for (int z0 = lane; z0 + 32 * (U - 1) < rowlen; z0 += 32 * U) {
            D2 v[U];
            if ((lane & 7) == 0) {
#pragma unroll
                for (int u = 0; u < U; u++) pf_l2(row + z0 + 32 * (u + U * DIST));
            }
#pragma unroll
            for (int u = 0; u < U; u++) v[u] = ld128(row + z0 + 32 * u);
#pragma unroll
            for (int u = 0; u < U; u++) s += v[u].a + v[u].b;
        }
    }
    if (s == 123.456) out[0] = s;
}

// CTA-cooperative rows: the whole CTA sweeps RPI consecutive rows per iteration; warp w takes the w-th
// 32*U-element segment of the (contiguous) RPI-row block -> CTA-wide accesses are contiguous
template <int U> __global__ void scale128_ctarows(D2 *p, size_t nelem, double f, int persistent)
{
    const size_t per_cta = (size_t) blockDim.x * U;            // elements per CTA iteration
    const size_t ntiles = (nelem + per_cta - 1) / per_cta;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (size_t t = blockIdx.x; t < ntiles; t += persistent ? gridDim.x : ntiles) {
        const size_t base = t * per_cta + (size_t) w * 32 * U + lane;
        D2 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) if (base + 32 * u < nelem) v[u] = ld128rw(p + base + 32 * u);
#pragma unroll
        for (int u = 0; u < U; u++) if (base + 32 * u < nelem) { v[u].a *= f; v[u].b *= f; st128(p + base + 32 * u, v[u]); }
    }
}
template <int U> __global__ void read128_ctarows(const D2 *p, size_t nelem, double *out, int persistent)
{
    double s = 0;
    const size_t per_cta = (size_t) blockDim.x * U;
    const size_t ntiles = (nelem + per_cta - 1) / per_cta;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (size_t t = blockIdx.x; t < ntiles; t += persistent ? gridDim.x : ntiles) {
        const size_t base = t * per_cta + (size_t) w * 32 * U + lane;
        D2 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) if (base + 32 * u < nelem) v[u] = ld128(p + base + 32 * u); else { v[u].a = 0; v[u].b = 0; }
#pragma unroll
        for (int u = 0; u < U; u++) s += v[u].a + v[u].b;
    }
    if (s == 123.456) out[0] = s;
}

template <class F> static void timeit(const char *name, double bytes, F f)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f();
    float best = 1e30f;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    printf("%-44s %8.3f ms  %7.1f GB/s %s\n", name, best, bytes / best / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
    fflush(stdout);
}

int main(int argc, char **argv)
{
    const int N = argc > 1 ? atoi(argv[1]) : 1024;
    const int L = N / 2 + 1;
    const size_t n = (size_t) N * N * L;     // 16-byte elements
    const double bytes = 16.0 * n;
    void *p; double *out;
    if (cudaMalloc(&p, n * 16) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMalloc(&out, 8);
    cudaMemset(p, 0, n * 16);
    printf("N=%d  %.2f GB\n", N, bytes / 1e9);
    for (int cta_per_sm : {1, 2, 4}) {
        const int g = 148 * cta_per_sm, t = cta_per_sm == 1 ? 512 : 256;
        char nm[96];
        snprintf(nm, 96, "read128 flat U8 grid=%d x %d", g, t); timeit(nm, bytes, [&] { read128<8><<<g, t>>>((D2 *) p, n, out); });
        snprintf(nm, 96, "read256 flat U4 grid=%d x %d", g, t); timeit(nm, bytes, [&] { read256<4><<<g, t>>>((D4 *) p, n / 2, out); });
        snprintf(nm, 96, "read128 rows U8 grid=%d x %d", g, t); timeit(nm, bytes, [&] { read128_rows<8><<<g, t>>>((D2 *) p, N * N, L, out); });
        snprintf(nm, 96, "scale128 flat U4 grid=%d x %d", g, t); timeit(nm, 2 * bytes, [&] { scale128<4><<<g, t>>>((D2 *) p, n, 1.0); });
        snprintf(nm, 96, "scale128 flat U8 grid=%d x %d", g, t); timeit(nm, 2 * bytes, [&] { scale128<8><<<g, t>>>((D2 *) p, n, 1.0); });
        snprintf(nm, 96, "scale256 flat U4 grid=%d x %d", g, t); timeit(nm, 2 * bytes, [&] { scale256<4><<<g, t>>>((D4 *) p, n / 2, 1.0); });
        snprintf(nm, 96, "scale128 rows U4 grid=%d x %d", g, t); timeit(nm, 2 * bytes, [&] { scale128_rows<4><<<g, t>>>((D2 *) p, N * N, L, 1.0); });
        snprintf(nm, 96, "scale128 rows U8 grid=%d x %d", g, t); timeit(nm, 2 * bytes, [&] { scale128_rows<8><<<g, t>>>((D2 *) p, N * N, L, 1.0); });
    }
    timeit("read128 rows U8 pf1 148x512", bytes, [&] { read128_rows_pf<8, 1><<<148, 512>>>((D2 *) p, N * N, L, out); });
    timeit("read128 rows U8 pf2 148x512", bytes, [&] { read128_rows_pf<8, 2><<<148, 512>>>((D2 *) p, N * N, L, out); });
    timeit("read128 rows U8 pf4 148x512", bytes, [&] { read128_rows_pf<8, 4><<<148, 512>>>((D2 *) p, N * N, L, out); });
    timeit("scale128 rows U4 pf2 592x256", 2 * bytes, [&] { scale128_rows_pf<4, 2><<<592, 256>>>((D2 *) p, N * N, L, 1.0); });
    timeit("scale128 rows U4 pf4 592x256", 2 * bytes, [&] { scale128_rows_pf<4, 4><<<592, 256>>>((D2 *) p, N * N, L, 1.0); });
    timeit("scale128 cta-contig U4 persistent 592x256", 2 * bytes, [&] { scale128_ctarows<4><<<592, 256>>>((D2 *) p, n, 1.0, 1); });
    timeit("scale128 cta-contig U8 persistent 296x256", 2 * bytes, [&] { scale128_ctarows<8><<<296, 256>>>((D2 *) p, n, 1.0, 1); });
    timeit("scale128 cta-contig U4 persistent 1184x256", 2 * bytes, [&] { scale128_ctarows<4><<<1184, 256>>>((D2 *) p, n, 1.0, 1); });
    timeit("scale128 cta-contig U4 one-tile-per-CTA x256", 2 * bytes, [&] { scale128_ctarows<4><<<(unsigned) ((n + 1023) / 1024), 256>>>((D2 *) p, n, 1.0, 0); });
    timeit("read128 cta-contig U8 persistent 148x512", bytes, [&] { read128_ctarows<8><<<148, 512>>>((D2 *) p, n, out, 1); });
    timeit("read128 cta-contig U8 persistent 296x512", bytes, [&] { read128_ctarows<8><<<296, 512>>>((D2 *) p, n, out, 1); });
    timeit("read128 cta-contig U8 one-tile-per-CTA x512", bytes, [&] { read128_ctarows<8><<<(unsigned) ((n + 4095) / 4096), 512>>>((D2 *) p, n, out, 0); });
    for (int rpc : {8, 16, 64, 256}) {
        char nm[96];
        snprintf(nm, 96, "scale128 rows U4 non-persistent rpc=%d x256", rpc);
        timeit(nm, 2 * bytes, [&] { scale128_rows_np<4><<<(N * N + rpc - 1) / rpc, 256>>>((D2 *) p, N * N, L, rpc, 1.0); });
    }
    // big grids (not persistent): one CTA per 64 KB
    {
        const int t = 256; const size_t per = 4096; const int g = (int) ((n + per - 1) / per);
        timeit("scale128 flat U4 grid=n/4096 x 256", 2 * bytes, [&] { scale128<4><<<g, t>>>((D2 *) p, n, 1.0); });
        timeit("read128 flat U8 grid=n/4096 x 256", bytes, [&] { read128<8><<<g / 2, t>>>((D2 *) p, n, out); });
    }
    void *q;
    if (cudaMalloc(&q, n * 16) == cudaSuccess) {
        timeit("cudaMemcpy D2D", 2 * bytes, [&] { cudaMemcpyAsync(q, p, n * 16, cudaMemcpyDeviceToDevice); });
        cudaFree(q);
    }
    return 0;
}
