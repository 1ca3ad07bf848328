// Slab-decomposed r2c / c2r FFT of the PM density grid, device resident, producing exactly the layout the hook is handed
// in a Gadget-2 PM step:
//
//     rfftwnd_mpi(fft_forward_plan, 1, rhogrid, workspace, FFTW_TRANSPOSED_ORDER);     (gadget-2/0002 patch:116)
//     add_nu_power_to_rhogrid(Time, BoxSize, fft_of_rhogrid, PMGRID, slabstart_y, nslab_y, comm);
//
// i.e. in:  real space, x-slabs:   rho[x - slabstart_x][y][z], z padded to 2 (N/2+1) reals per row (FFTW's in-place r2c
//           layout, pm_periodic.c's rhogrid);
//      out: k space, y-slabs, "transposed order":  F[y - slabstart_y][x][kz], kz = 0..N/2 -- the slab K1 and K3 sweep.
//
// Three steps, all on the device:
//  1. 2-D r2c over (y, z) of every local x plane, in place in the padded grid (cuFFT, batched; loaded with dlopen like
//     NCCL, so the library still loads where cuFFT is absent);
//  2. the global transpose x <-> y.  Every row of N/2+1 complex values moves as a whole: row (x, y) of rank s goes to
//     row (y - ystart_r, x) of the rank r that owns y.  ONE kernel does the transposition AND the exchange: it reads the
//     rank's own rows and writes them straight into the k-space slabs of the other GPUs, which are mapped here through
//     CUDA IPC (peer memory over NVLink / NVSwitch) -- no pack / all-to-all / unpack passes, each byte is read once and
//     written once, at its final address.  Two rounds of the peer-memory flag protocol (ksn_p2p.cuh) fence it: nobody
//     writes into a slab its owner may still be reading, nobody reads its slab before every peer's rows have landed;
//  3. 1-D c2c along x (stride N/2+1) of every (y, kz) column of the received slab, in place (cuFFT, one call per y plane).
// The stages overlap: the exchange of a batch of planes runs on a stream of its own while cuFFT transforms the next batch
// (the 2-D pass is HBM-bound, the exchange NVLink-bound), and the per-plane 1-D calls are dealt to four streams.
// The inverse runs the same steps backwards (c2c inverse, the mirrored exchange, 2-D c2r); unnormalised like FFTW.
// One rank: the same kernels with every row local.
#include "ksn_internal.cuh"
#include "ksn_p2p.cuh"

#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

namespace ksn {

// ---------------------------------------------------------------- cuFFT through dlopen
typedef int cufft_handle;
enum { CUFFT_D2Z_ = 0x6a, CUFFT_Z2D_ = 0x6c, CUFFT_Z2Z_ = 0x69, CUFFT_FWD = -1, CUFFT_INV = 1 };
struct CufftApi {
    void *h = nullptr;
    int (*Create)(cufft_handle *) = nullptr;
    int (*Destroy)(cufft_handle) = nullptr;
    int (*SetAutoAllocation)(cufft_handle, int) = nullptr;
    int (*MakePlanMany)(cufft_handle, int, int *, int *, int, int, int *, int, int, int, int, size_t *) = nullptr;
    int (*SetWorkArea)(cufft_handle, void *) = nullptr;
    int (*SetStream)(cufft_handle, cudaStream_t) = nullptr;
    int (*ExecD2Z)(cufft_handle, double *, void *) = nullptr;
    int (*ExecZ2D)(cufft_handle, void *, double *) = nullptr;
    int (*ExecZ2Z)(cufft_handle, void *, void *, int) = nullptr;
};
static CufftApi g_cufft;

static int load_cufft()
{
    if (g_cufft.h) return KSN_OK;
    const char *names[] = { getenv("KSN_CUFFT_LIB"), "libcufft.so.11", "libcufft.so.12", "libcufft.so" };
    void *h = nullptr;
    for (const char *n : names) { if (n && *n && (h = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL))) break; }
    for (const char *n : names) { if (h) break; if (n && *n) h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); }
    if (!h) return set_error(KSN_ENODEV, "cannot load libcufft: %s", dlerror());
    CufftApi a;
    a.Create = (decltype(a.Create)) dlsym(h, "cufftCreate");
    a.Destroy = (decltype(a.Destroy)) dlsym(h, "cufftDestroy");
    a.SetAutoAllocation = (decltype(a.SetAutoAllocation)) dlsym(h, "cufftSetAutoAllocation");
    a.MakePlanMany = (decltype(a.MakePlanMany)) dlsym(h, "cufftMakePlanMany");
    a.SetWorkArea = (decltype(a.SetWorkArea)) dlsym(h, "cufftSetWorkArea");
    a.SetStream = (decltype(a.SetStream)) dlsym(h, "cufftSetStream");
    a.ExecD2Z = (decltype(a.ExecD2Z)) dlsym(h, "cufftExecD2Z");
    a.ExecZ2D = (decltype(a.ExecZ2D)) dlsym(h, "cufftExecZ2D");
    a.ExecZ2Z = (decltype(a.ExecZ2Z)) dlsym(h, "cufftExecZ2Z");
    if (!a.Create || !a.Destroy || !a.SetAutoAllocation || !a.MakePlanMany || !a.SetWorkArea || !a.SetStream || !a.ExecD2Z || !a.ExecZ2D || !a.ExecZ2Z)
        return set_error(KSN_ENODEV, "libcufft lacks an expected symbol");
    a.h = h;
    g_cufft = a;
    return KSN_OK;
}
#define KSN_FFT(call) do { int _e = (call); if (_e) return set_error(KSN_ECUDA, "%s failed with cuFFT code %d", #call, _e); } while (0)

// ---------------------------------------------------------------- state
struct FftState {
    bool planned = false;
    int N = 0, R = 1, rank = 0;
    long long xs[KSN_P2P_MAX_RANKS + 1] = {}, ys[KSN_P2P_MAX_RANKS + 1] = {};   // FFTW-style partitions of x and y planes
    static constexpr int K1D = 4;         // streams (and plans) the per-plane 1-D transforms are dealt to
    static constexpr int G1D = 8;         // y planes per 1-D batch (the unit the inverse hands to the exchange)
    cufft_handle p2d_f = 0, p2d_i = 0, p1d[K1D] = {};
    bool have2d = false, have1d = false;
    int batch2d = 0;
    void *work = nullptr; size_t work_bytes = 0;
    void *work1d[K1D] = {};
    cudaStream_t xstream = nullptr, s1d[K1D] = {};    // exchange stream, 1-D streams
    cudaEvent_t *evp = nullptr; int nevp = 0;         // event pool: one per batch in flight
    // the k-space slabs of all ranks as mapped in this process (slab[rank] = the pointer the caller registered)
    void *slab[KSN_P2P_MAX_RANKS] = {};
    bool opened[KSN_P2P_MAX_RANKS] = {};
    // ... and the real-space grids, for the inverse
    void *real[KSN_P2P_MAX_RANKS] = {};
    bool ropened[KSN_P2P_MAX_RANKS] = {};
    double *d_token = nullptr;            // one double: the payload of the barrier rounds
    cudaEvent_t ev[5] = {};               // stage boundaries of the last transform (ksn_fft_timing)
    bool have_ev = false;
    float stage_ms[4] = {};               // wait for the peers | 2-D pass with the exchange behind it | closing fence | 1-D pass
};
static FftState g_fft;

struct ExchangePlan {
    double2 *peer[KSN_P2P_MAX_RANKS];     // destination base of every rank
    long long lo[KSN_P2P_MAX_RANKS + 1];  // partition of the index that selects the destination rank
    int R, N, L;
    long long own0, nown;                 // this rank's planes of the OTHER index
};

// Forward: src = this rank's [x - x0][y][kz] rows; row (x, y) -> rank r = owner(y), row (y - ys[r], x).
// Inverse: src = this rank's [y - y0][x][kz] rows; row (y, x) -> rank r = owner(x), row (x - xs[r], y).
// Both are "row (a, b) -> owner(b): row (b - lo[owner], a)": one kernel, one CTA per row, 16-byte loads and stores
// (a row is N/2+1 double2: 16.4 KB at PMGRID 2048, contiguous on both sides).
__global__ void __launch_bounds__(256)
fft_exchange_kernel(const double2 *__restrict__ src, const ExchangePlan p)
{
    const long long row = blockIdx.x;                    // a_local * N + b
    const long long al = row / p.N;
    const int b = (int) (row - al * p.N);
    int r = 0;
    while (b >= p.lo[r + 1]) r++;                        // owner of b (R <= 16)
    const long long a = p.own0 + al;
    const double2 *s = src + row * p.L;
    double2 *d = p.peer[r] + ((b - p.lo[r]) * (long long) p.N + a) * p.L;
    for (int i = threadIdx.x; i < p.L; i += blockDim.x) {
        double2 v;
        asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(s + i));
        asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(d + i), "d"(v.x), "d"(v.y) : "memory");
    }
}

static void partition(int n, int R, long long *lo)
{
    // FFTW2-style slabs: contiguous, as even as possible, the first ranks take the extra planes (host.slab_partition)
    const int base = n / R, extra = n % R;
    lo[0] = 0;
    for (int r = 0; r < R; r++) lo[r + 1] = lo[r] + base + (r < extra ? 1 : 0);
}

static void fft_release_plans()
{
    FftState &f = g_fft;
    if (f.have2d && g_cufft.Destroy) { g_cufft.Destroy(f.p2d_f); g_cufft.Destroy(f.p2d_i); }
    if (f.have1d && g_cufft.Destroy) for (int k = 0; k < FftState::K1D; k++) g_cufft.Destroy(f.p1d[k]);
    f.have2d = f.have1d = false;
    if (f.work) cudaFree(f.work);
    f.work = nullptr; f.work_bytes = 0;
    for (int k = 0; k < FftState::K1D; k++) { if (f.work1d[k]) cudaFree(f.work1d[k]); f.work1d[k] = nullptr; }
}

static void fft_unmap()
{
    FftState &f = g_fft;
    for (int r = 0; r < KSN_P2P_MAX_RANKS; r++) {
        if (f.opened[r] && f.slab[r]) cudaIpcCloseMemHandle(f.slab[r]);
        if (f.ropened[r] && f.real[r]) cudaIpcCloseMemHandle(f.real[r]);
        f.slab[r] = f.real[r] = nullptr;
        f.opened[r] = f.ropened[r] = false;
    }
}

static void fft_mark(int i)
{
    FftState &f = g_fft;
    if (!f.have_ev) {
        for (auto &e : f.ev) cudaEventCreate(&e);
        f.have_ev = true;
    }
    cudaEventRecord(f.ev[i], ctx().stream);
}

static void fft_collect(bool inverse)
{
    FftState &f = g_fft;
    float t[4] = {};
    for (int i = 0; i < 4; i++) cudaEventElapsedTime(&t[i], f.ev[i], f.ev[i + 1]);
    (void) inverse;
    for (int i = 0; i < 4; i++) f.stage_ms[i] = t[i];
}

void fft_shutdown()
{
    fft_release_plans();
    fft_unmap();
    if (g_fft.have_ev) for (auto &e : g_fft.ev) cudaEventDestroy(e);
    for (int i = 0; i < g_fft.nevp; i++) cudaEventDestroy(g_fft.evp[i]);
    free(g_fft.evp);
    if (g_fft.xstream) cudaStreamDestroy(g_fft.xstream);
    for (auto &st : g_fft.s1d) if (st) cudaStreamDestroy(st);
    if (g_fft.d_token) cudaFree(g_fft.d_token);
    g_fft = FftState();
}

// a barrier over the ranks of the peer-memory backend, in stream order: everything this rank enqueued before it is
// visible to a peer once that peer's barrier kernel has returned (release / acquire at system scope, ksn_p2p.cuh)
static int fft_barrier()
{
    FftState &f = g_fft;
    if (f.R == 1) return KSN_OK;
    if (!p2p_active() || ctx().nranks != f.R) return set_error(KSN_ECOMM, "ksn_fft: %d ranks need the peer-memory backend (ksn_comm_p2p_init) on the same ranks", f.R);
    return p2p_allreduce_device(f.d_token, 1);
}

}  // namespace ksn

using namespace ksn;

extern "C" int ksn_fft_plan(int dims, int nranks, int rank)
{
    int rc = ensure_init();
    if (rc) return rc;
    if (dims < 2 || nranks < 1 || nranks > KSN_P2P_MAX_RANKS || rank < 0 || rank >= nranks || nranks > dims)
        return set_error(KSN_EINVAL, "ksn_fft_plan: bad arguments (dims=%d, rank %d of %d)", dims, rank, nranks);
    rc = load_cufft();
    if (rc) return rc;
    FftState &f = g_fft;
    Ctx &c = ctx();
    fft_release_plans();
    fft_unmap();
    f.planned = false;
    f.N = dims; f.R = nranks; f.rank = rank;
    partition(dims, nranks, f.xs);
    partition(dims, nranks, f.ys);
    const int N = dims, L = N / 2 + 1;
    const long long nx = f.xs[rank + 1] - f.xs[rank], ny = f.ys[rank + 1] - f.ys[rank];
    if (!f.d_token) { KSN_CUDA(cudaMalloc((void **) &f.d_token, 64)); KSN_CUDA(cudaMemset(f.d_token, 0, 64)); }
    size_t ws_max = 0;
    if (nx > 0) {
        // planes per call: the largest divisor of nx whose planes stay below ~1 GB (bounds cuFFT's work area)
        int B = 1;
        const size_t plane_bytes = (size_t) N * L * 16;
        for (int b = 1; b <= nx; b++) if (nx % b == 0 && (size_t) b * plane_bytes <= ((size_t) 1 << 30)) B = b;
        f.batch2d = B;
        int n2[2] = { N, N }, rembed[2] = { N, 2 * L }, cembed[2] = { N, L };
        size_t ws = 0;
        KSN_FFT(g_cufft.Create(&f.p2d_f));
        KSN_FFT(g_cufft.Create(&f.p2d_i));
        f.have2d = true;
        KSN_FFT(g_cufft.SetAutoAllocation(f.p2d_f, 0));
        KSN_FFT(g_cufft.SetAutoAllocation(f.p2d_i, 0));
        KSN_FFT(g_cufft.MakePlanMany(f.p2d_f, 2, n2, rembed, 1, N * 2 * L, cembed, 1, N * L, CUFFT_D2Z_, B, &ws));
        ws_max = ws;
        KSN_FFT(g_cufft.MakePlanMany(f.p2d_i, 2, n2, cembed, 1, N * L, rembed, 1, N * 2 * L, CUFFT_Z2D_, B, &ws));
        if (ws > ws_max) ws_max = ws;
        KSN_FFT(g_cufft.SetStream(f.p2d_f, c.stream));
        KSN_FFT(g_cufft.SetStream(f.p2d_i, c.stream));
    }
    if (!f.xstream) {
        // the exchange runs BEHIND cuFFT's batches, whose kernels fill the device: give its CTAs the first free slots
        int lo_pri = 0, hi_pri = 0;
        KSN_CUDA(cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri));
        KSN_CUDA(cudaStreamCreateWithPriority(&f.xstream, cudaStreamNonBlocking, hi_pri));
    }
    for (auto &st : f.s1d) if (!st) KSN_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    if (ny > 0) {
        int n1[1] = { N }, embed[1] = { N };
        for (int k = 0; k < FftState::K1D; k++) {
            size_t ws = 0;
            KSN_FFT(g_cufft.Create(&f.p1d[k]));
            if (k == 0) f.have1d = true;
            KSN_FFT(g_cufft.SetAutoAllocation(f.p1d[k], 0));
            // along x inside one y plane [x][kz]: element stride L, the L columns one after the other
            KSN_FFT(g_cufft.MakePlanMany(f.p1d[k], 1, n1, embed, L, 1, embed, L, 1, CUFFT_Z2Z_, L, &ws));
            if (ws) { KSN_CUDA(cudaMalloc(&f.work1d[k], ws)); KSN_FFT(g_cufft.SetWorkArea(f.p1d[k], f.work1d[k])); }
            KSN_FFT(g_cufft.SetStream(f.p1d[k], f.s1d[k]));
        }
    }
    if (ws_max) {
        KSN_CUDA(cudaMalloc(&f.work, ws_max));
        f.work_bytes = ws_max;
        if (f.have2d) { KSN_FFT(g_cufft.SetWorkArea(f.p2d_f, f.work)); KSN_FFT(g_cufft.SetWorkArea(f.p2d_i, f.work)); }
    }
    {
        const long long nb2 = nx > 0 ? nx / f.batch2d + 1 : 1, nb1 = ny / FftState::G1D + 2;
        const int want = (int) (nb2 > nb1 ? nb2 : nb1) + FftState::K1D + 4;
        if (want > f.nevp) {
            cudaEvent_t *np = (cudaEvent_t *) realloc(f.evp, sizeof(cudaEvent_t) * want);
            if (!np) return set_error(KSN_ENOMEM, "ksn_fft_plan: out of host memory");
            f.evp = np;
            for (int i = f.nevp; i < want; i++) KSN_CUDA(cudaEventCreateWithFlags(&f.evp[i], cudaEventDisableTiming));
            f.nevp = want;
        }
    }
    f.planned = true;
    return KSN_OK;
}

extern "C" int ksn_fft_layout(long long *slabstart_x, long long *nslab_x, long long *slabstart_y, long long *nslab_y,
                              size_t *real_bytes_padded, size_t *kspace_bytes)
{
    const FftState &f = g_fft;
    if (!f.planned) return set_error(KSN_EINVAL, "ksn_fft_layout: call ksn_fft_plan first");
    const long long nx = f.xs[f.rank + 1] - f.xs[f.rank], ny = f.ys[f.rank + 1] - f.ys[f.rank];
    const size_t L = (size_t) f.N / 2 + 1;
    if (slabstart_x) *slabstart_x = f.xs[f.rank];
    if (nslab_x) *nslab_x = nx;
    if (slabstart_y) *slabstart_y = f.ys[f.rank];
    if (nslab_y) *nslab_y = ny;
    if (real_bytes_padded) *real_bytes_padded = (size_t) nx * f.N * 2 * L * sizeof(double);
    if (kspace_bytes) *kspace_bytes = (size_t) ny * f.N * L * 2 * sizeof(double);
    return KSN_OK;
}

// CUDA-IPC handles of this rank's two buffers (128 bytes: k-space slab | real-space grid); the host gathers them in rank
// order by any means and hands all of them to ksn_fft_attach.  The buffers must be whole cudaMalloc allocations.
extern "C" int ksn_fft_export(void *d_kspace, void *d_real, void *handles128)
{
    int rc = ensure_init();
    if (rc) return rc;
    if (!d_kspace || !d_real || !handles128) return set_error(KSN_EINVAL, "ksn_fft_export: null argument");
    cudaIpcMemHandle_t h[2];
    KSN_CUDA(cudaIpcGetMemHandle(&h[0], d_kspace));
    KSN_CUDA(cudaIpcGetMemHandle(&h[1], d_real));
    memcpy(handles128, h, sizeof h);
    return KSN_OK;
}

// handles: nranks * 128 bytes in rank order (ignored for one rank).  d_kspace / d_real: this rank's own buffers.
extern "C" int ksn_fft_attach(void *d_kspace, void *d_real, const void *handles)
{
    FftState &f = g_fft;
    if (!f.planned) return set_error(KSN_EINVAL, "ksn_fft_attach: call ksn_fft_plan first");
    if (!d_kspace || !d_real || (f.R > 1 && !handles)) return set_error(KSN_EINVAL, "ksn_fft_attach: null argument");
    fft_unmap();
    for (int r = 0; r < f.R; r++) {
        if (r == f.rank) { f.slab[r] = d_kspace; f.real[r] = d_real; continue; }
        cudaIpcMemHandle_t h[2];
        memcpy(h, (const char *) handles + (size_t) r * sizeof h, sizeof h);
        cudaError_t e = cudaIpcOpenMemHandle(&f.slab[r], h[0], cudaIpcMemLazyEnablePeerAccess);
        if (e == cudaSuccess) { f.opened[r] = true; e = cudaIpcOpenMemHandle(&f.real[r], h[1], cudaIpcMemLazyEnablePeerAccess); }
        if (e != cudaSuccess) {
            cudaGetLastError();
            fft_unmap();
            return set_error(KSN_ECOMM, "ksn_fft_attach: cudaIpcOpenMemHandle(rank %d): %s (peer access between the GPUs of one box is required)", r, cudaGetErrorString(e));
        }
        f.ropened[r] = true;
    }
    return KSN_OK;
}

static int fft_check(const char *who, void *d_real, void *d_kspace)
{
    const FftState &f = g_fft;
    if (!f.planned) return set_error(KSN_EINVAL, "%s: call ksn_fft_plan first", who);
    if (f.slab[f.rank] != d_kspace || f.real[f.rank] != d_real)
        return set_error(KSN_EINVAL, "%s: these are not the buffers ksn_fft_attach was given", who);
    return KSN_OK;
}

// rows [p0, p0 + np) x N of `src` (complex view, L per row) to their owners, on f.xstream
static int fft_exchange(const void *src, void *const *dest, const long long *lo, long long own0, long long p0, long long np)
{
    FftState &f = g_fft;
    const int N = f.N, L = N / 2 + 1;
    if (np <= 0) return KSN_OK;
    if (np * N > 0x7fffffffLL) return set_error(KSN_EINVAL, "ksn_fft: %lld rows in one exchange batch", np * N);
    ExchangePlan p;
    for (int r = 0; r < KSN_P2P_MAX_RANKS; r++) p.peer[r] = (double2 *) dest[r];
    for (int r = 0; r <= f.R; r++) p.lo[r] = lo[r];
    p.R = f.R; p.N = N; p.L = L; p.own0 = own0 + p0; p.nown = np;
    fft_exchange_kernel<<<(unsigned) (np * N), 256, 0, f.xstream>>>((const double2 *) src + (size_t) p0 * N * L, p);
    ctx().launches++;
    KSN_CUDA(cudaGetLastError());
    return KSN_OK;
}

// the per-plane 1-D transforms of planes [0, ny), batch j (G1D planes) on stream j % K1D; batch j's event is evp[j]
static int fft_1d_batches(void *d_kspace, long long ny, int dir, cudaEvent_t after)
{
    FftState &f = g_fft;
    const int N = f.N, L = N / 2 + 1;
    for (int k = 0; k < FftState::K1D; k++) KSN_CUDA(cudaStreamWaitEvent(f.s1d[k], after, 0));
    int j = 0;
    for (long long y0 = 0; y0 < ny; y0 += FftState::G1D, j++) {
        const int k = j % FftState::K1D;
        const long long y1 = y0 + FftState::G1D < ny ? y0 + FftState::G1D : ny;
        for (long long y = y0; y < y1; y++) {
            void *plane = (char *) d_kspace + (size_t) y * N * L * 16;
            KSN_FFT(g_cufft.ExecZ2Z(f.p1d[k], plane, plane, dir));
        }
        KSN_CUDA(cudaEventRecord(f.evp[j], f.s1d[k]));
    }
    return KSN_OK;
}

extern "C" int ksn_fft_forward(void *d_real, void *d_kspace)
{
    int rc = fft_check("ksn_fft_forward", d_real, d_kspace);
    if (rc) return rc;
    FftState &f = g_fft;
    Ctx &c = ctx();
    const int N = f.N, L = N / 2 + 1;
    const long long x0 = f.xs[f.rank], nx = f.xs[f.rank + 1] - x0, ny = f.ys[f.rank + 1] - f.ys[f.rank];
    cudaEvent_t *ev = f.evp, go = f.evp[f.nevp - 1], joined = f.evp[f.nevp - 2];
    // nobody writes into a slab its owner may still be using (every rank is past whatever it did with its k-space slab)
    fft_mark(0);
    rc = fft_barrier();
    if (rc) return rc;
    fft_mark(1);
    KSN_CUDA(cudaEventRecord(go, c.stream));
    KSN_CUDA(cudaStreamWaitEvent(f.xstream, go, 0));
    // 1. 2-D r2c of the local x planes, batch by batch, in place in the padded grid; 2. behind each batch, on the exchange
    // stream, its rows go to the k-space slabs of their owners (transposed)
    int j = 0;
    for (long long p = 0; p < nx; p += f.batch2d, j++) {
        double *plane = (double *) d_real + (size_t) p * N * 2 * L;
        KSN_FFT(g_cufft.ExecD2Z(f.p2d_f, plane, plane));
        KSN_CUDA(cudaEventRecord(ev[j], c.stream));
        KSN_CUDA(cudaStreamWaitEvent(f.xstream, ev[j], 0));
        rc = fft_exchange(d_real, f.slab, f.ys, x0, p, f.batch2d);
        if (rc) return rc;
    }
    KSN_CUDA(cudaEventRecord(joined, f.xstream));
    KSN_CUDA(cudaStreamWaitEvent(c.stream, joined, 0));
    // ... and nobody reads its slab before every peer's rows have landed
    fft_mark(2);
    rc = fft_barrier();
    if (rc) return rc;
    fft_mark(3);
    // 3. 1-D c2c along x of every column of the received slab
    KSN_CUDA(cudaEventRecord(go, c.stream));
    rc = fft_1d_batches(d_kspace, ny, CUFFT_FWD, go);
    if (rc) return rc;
    for (int k = 0; k < FftState::K1D; k++) {
        KSN_CUDA(cudaEventRecord(joined, f.s1d[k]));
        KSN_CUDA(cudaStreamWaitEvent(c.stream, joined, 0));
    }
    fft_mark(4);
    KSN_CUDA(cudaStreamSynchronize(c.stream));
    fft_collect(false);
    if (f.R > 1) { rc = p2p_status_async(); if (!rc) { KSN_CUDA(cudaStreamSynchronize(c.stream)); rc = p2p_status_result(); } }
    return rc;
}

extern "C" int ksn_fft_inverse(void *d_kspace, void *d_real)
{
    int rc = fft_check("ksn_fft_inverse", d_real, d_kspace);
    if (rc) return rc;
    FftState &f = g_fft;
    Ctx &c = ctx();
    const int N = f.N, L = N / 2 + 1;
    const long long y0 = f.ys[f.rank], ny = f.ys[f.rank + 1] - y0, nx = f.xs[f.rank + 1] - f.xs[f.rank];
    cudaEvent_t *ev = f.evp, go = f.evp[f.nevp - 1], joined = f.evp[f.nevp - 2];
    fft_mark(0);
    rc = fft_barrier();                   // every rank is past whatever it did with its real-space grid
    if (rc) return rc;
    fft_mark(1);
    KSN_CUDA(cudaEventRecord(go, c.stream));
    // 1-D inverse along x, batch by batch on the 1-D streams; behind each batch its rows go back to the x-slabs of their owners
    rc = fft_1d_batches(d_kspace, ny, CUFFT_INV, go);
    if (rc) return rc;
    int j = 0;
    for (long long y = 0; y < ny; y += FftState::G1D, j++) {
        KSN_CUDA(cudaStreamWaitEvent(f.xstream, ev[j], 0));
        rc = fft_exchange(d_kspace, f.real, f.xs, y0, y, y + FftState::G1D < ny ? FftState::G1D : ny - y);
        if (rc) return rc;
    }
    if (ny == 0) KSN_CUDA(cudaStreamWaitEvent(f.xstream, go, 0));
    KSN_CUDA(cudaEventRecord(joined, f.xstream));
    KSN_CUDA(cudaStreamWaitEvent(c.stream, joined, 0));
    fft_mark(2);
    rc = fft_barrier();
    if (rc) return rc;
    fft_mark(3);
    for (long long p = 0; p < nx; p += f.batch2d) {
        double *plane = (double *) d_real + (size_t) p * N * 2 * L;
        KSN_FFT(g_cufft.ExecZ2D(f.p2d_i, plane, plane));
    }
    fft_mark(4);
    KSN_CUDA(cudaStreamSynchronize(c.stream));
    fft_collect(true);
    if (f.R > 1) { rc = p2p_status_async(); if (!rc) { KSN_CUDA(cudaStreamSynchronize(c.stream)); rc = p2p_status_result(); } }
    return rc;
}

// stages of the most recent transform on this rank, in ms.  Forward: waiting for the peers | 2-D pass with the transpose +
// exchange behind it | closing fence | 1-D pass.  Inverse: waiting | 1-D pass with the exchange behind it | fence | 2-D pass
extern "C" int ksn_fft_timing(float *ms4)
{
    if (!ms4 || !g_fft.have_ev) return set_error(KSN_EINVAL, "ksn_fft_timing: no transform yet");
    for (int i = 0; i < 4; i++) ms4[i] = g_fft.stage_ms[i];
    return KSN_OK;
}

extern "C" void ksn_fft_destroy(void) { fft_shutdown(); }
