// Device context, memory helpers, collective backends and timing for libkspace_neutrinos_b200.
// C-ABI: include/ksn_b200.h.  No CPU fallback lives here: when CUDA is unusable every
// compute entry reports KSN_ENODEV.
#include "ksn_internal.cuh"
#include "ksn_p2p.cuh"

#include <dlfcn.h>
#include <nvtx3/nvToolsExt.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <mutex>

namespace ksn {

static Ctx g_ctx;
static char g_err[1024] = "";
static std::mutex g_mu;

Ctx &ctx() { return g_ctx; }

int set_error(int code, const char *fmt, ...)
{
    va_list va;
    va_start(va, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, va);
    va_end(va);
    if (getenv("KSN_VERBOSE")) fprintf(stderr, "[ksn] error %d: %s\n", code, g_err);
    return code;
}

int check_cuda(cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return KSN_OK;
    int code = (e == cudaErrorMemoryAllocation) ? KSN_ENOMEM
             : (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? KSN_ENODEV : KSN_ECUDA;
    return set_error(code, "%s: %s", what, cudaGetErrorString(e));
}

static int pick_device()
{
    const char *e = getenv("KSN_DEVICE");
    if (e && *e) return atoi(e);
    e = getenv("LOCAL_RANK");
    if (e && *e) {
        int n = 0;
        if (cudaGetDeviceCount(&n) == cudaSuccess && n > 0) return atoi(e) % n;
    }
    return 0;
}

static int do_init(int device)
{
    Ctx &c = g_ctx;
    if (c.inited) return KSN_OK;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return set_error(KSN_ENODEV, "no CUDA device visible (%s); this library has no CPU path",
                         e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0) device = pick_device();
    if (device >= n) return set_error(KSN_EINVAL, "device %d requested but only %d visible", device, n);
    KSN_CUDA(cudaSetDevice(device));
    cudaDeviceProp p;
    KSN_CUDA(cudaGetDeviceProperties(&p, device));
    if (p.major < 10)
        return set_error(KSN_ENODEV, "device %d is sm_%d%d, this build carries sm_100a code only; this library has no CPU path", device, p.major, p.minor);
    c.device = device;
    c.num_sms = p.multiProcessorCount;
    c.smem_optin = p.sharedMemPerBlockOptin;
    KSN_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    KSN_CUDA(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < PH_COUNT; i++)
        for (int j = 0; j < 2; j++) KSN_CUDA(cudaEventCreate(&c.ev[i][j]));
    c.inited = true;
    return KSN_OK;
}

int ensure_init()
{
    if (g_ctx.inited) {
        // another library (torch) may have switched the current device on this thread
        int cur = -1;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != g_ctx.device) cudaSetDevice(g_ctx.device);
        return KSN_OK;
    }
    std::lock_guard<std::mutex> lk(g_mu);
    return do_init(-1);
}

int ensure_device_buffer(void **p, size_t *cap, size_t bytes)
{
    if (*p && *cap >= bytes) return KSN_OK;
    if (*p) { cudaFree(*p); *p = nullptr; *cap = 0; }
    KSN_CUDA(cudaMalloc(p, bytes));
    *cap = bytes;
    return KSN_OK;
}

int ensure_pinned_buffer(void **p, size_t *cap, size_t bytes)
{
    if (*p && *cap >= bytes) return KSN_OK;
    if (*p) { cudaFreeHost(*p); *p = nullptr; *cap = 0; }
    KSN_CUDA(cudaHostAlloc(p, bytes, cudaHostAllocDefault));
    *cap = bytes;
    return KSN_OK;
}

// ---------------------------------------------------------------- staging plan for host-resident slabs
// Resident when the whole slab fits in free HBM (with 1 GB to spare): uploaded once, used by K1 and K3, downloaded once.
// Otherwise STREAMING: a ring of STAGE_RING chunks; K1 consumes the chunks as they land, K3 uploads them a second time
// and sends each back as soon as it is scaled (the second upload overlaps the downloads: PCIe is full duplex).
// KSN_STAGE_CHUNK_MB (default 256) sets the chunk size, KSN_STAGE_MAX_MB caps what may stay resident (tests).
int stage_plan(int real_bytes, int dims, long long nslab, StagePlan *plan)
{
    Ctx &c = g_ctx;
    plan->plane_bytes = (size_t) dims * (dims / 2 + 1) * 2 * real_bytes;
    plan->total = plan->plane_bytes * (size_t) nslab;
    const char *e = getenv("KSN_STAGE_CHUNK_MB");
    const size_t chunk_bytes = (size_t) (e && atoi(e) > 0 ? atoi(e) : 256) << 20;
    plan->chunk = (long long) (chunk_bytes / plan->plane_bytes);
    if (plan->chunk < 1) plan->chunk = 1;
    plan->nchunks = (int) ((nslab + plan->chunk - 1) / plan->chunk);
    size_t freeb = 0, totb = 0;
    if (cudaMemGetInfo(&freeb, &totb) != cudaSuccess) { cudaGetLastError(); freeb = 0; }
    e = getenv("KSN_STAGE_MAX_MB");
    const size_t cap = e ? (size_t) atoll(e) << 20 : (size_t) -1;
    const bool fits = plan->total <= cap && (c.stage_cap >= plan->total || plan->total + ((size_t) 1 << 30) <= freeb + c.stage_cap);
    plan->streaming = !fits && plan->nchunks > STAGE_RING;
    const size_t need = plan->streaming ? (size_t) STAGE_RING * plan->chunk * plan->plane_bytes : plan->total;
    int rc = ensure_device_buffer(&c.d_stage, &c.stage_cap, need);
    if (rc) return set_error(KSN_ENOMEM, "staging a host grid needs %zu bytes of free HBM", need);
    if (!c.copy_stream2) KSN_CUDA(cudaStreamCreateWithFlags(&c.copy_stream2, cudaStreamNonBlocking));
    if (!c.d_origin) KSN_CUDA(cudaMalloc(&c.d_origin, 16));
    return KSN_OK;
}

// ---------------------------------------------------------------- timing
// Every phase is also an NVTX range (host side, header-only NVTX 3: a few ns when no tool is attached), so that a
// profiler can be told to look at one phase only (ncu --nvtx --nvtx-include "ksn/K3/").  Phases overlap (uploads run
// beside K1), hence start/end ranges rather than push/pop.
static const char *const kPhaseName[PH_COUNT] = { "K1", "K1 reduce", "collective", "K2", "K3", "H2D", "D2H" };
static nvtxRangeId_t g_range[PH_COUNT];
static bool g_range_open[PH_COUNT];
static nvtxDomainHandle_t nvtx_domain()
{
    static nvtxDomainHandle_t d = nvtxDomainCreateA("ksn");
    return d;
}
static void range_begin(Phase p)
{
    if (g_range_open[p]) nvtxDomainRangeEnd(nvtx_domain(), g_range[p]);
    nvtxEventAttributes_t a = {};
    a.version = NVTX_VERSION;
    a.size = NVTX_EVENT_ATTRIB_STRUCT_SIZE;
    a.messageType = NVTX_MESSAGE_TYPE_ASCII;
    a.message.ascii = kPhaseName[p];
    g_range[p] = nvtxDomainRangeStartEx(nvtx_domain(), &a);
    g_range_open[p] = true;
}
static void range_end(Phase p)
{
    if (g_range_open[p]) nvtxDomainRangeEnd(nvtx_domain(), g_range[p]);
    g_range_open[p] = false;
}

void phase_begin(Phase p)
{
    Ctx &c = g_ctx;
    range_begin(p);
    if (!c.timing) return;
    if (c.ev_used[p]) {   // fold the previous interval of this phase before reusing its events
        float ms = 0;
        if (cudaEventSynchronize(c.ev[p][1]) == cudaSuccess && cudaEventElapsedTime(&ms, c.ev[p][0], c.ev[p][1]) == cudaSuccess)
            c.acc_ms[p] += ms;
        c.ev_used[p] = false;
    }
    cudaEventRecord(c.ev[p][0], (p == PH_H2D || p == PH_D2H) ? c.copy_stream : c.stream);
}

void phase_end(Phase p)
{
    Ctx &c = g_ctx;
    range_end(p);
    if (!c.timing) return;
    cudaEventRecord(c.ev[p][1], (p == PH_H2D || p == PH_D2H) ? c.copy_stream : c.stream);
    c.ev_used[p] = true;
}

void phase_collect()
{
    Ctx &c = g_ctx;
    if (!c.timing) return;
    for (int p = 0; p < PH_COUNT; p++) {
        if (!c.ev_used[p]) continue;
        float ms = 0;
        if (cudaEventSynchronize(c.ev[p][1]) == cudaSuccess && cudaEventElapsedTime(&ms, c.ev[p][0], c.ev[p][1]) == cudaSuccess)
            c.acc_ms[p] += ms;
        c.ev_used[p] = false;
    }
}

// ---------------------------------------------------------------- NCCL through dlopen
// NCCL is loaded lazily so that the library also loads on hosts without it (the no-GPU
// build container, single-GPU hosts).  Inside a torch process the already-loaded
// libnccl.so.2 is reused.
typedef struct { char internal[128]; } nccl_uid;
struct NcclApi {
    void *h = nullptr;
    int (*GetUniqueId)(nccl_uid *) = nullptr;
    int (*CommInitRank)(ncclComm **, int, nccl_uid, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm *, cudaStream_t) = nullptr;
    int (*CommDestroy)(ncclComm *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static int load_nccl()
{
    if (g_nccl.h) return KSN_OK;
    const char *names[] = { getenv("KSN_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
    void *h = nullptr;
    for (const char *n : names) {
        if (!n || !*n) continue;
        h = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        if (h) break;
    }
    for (const char *n : names) {
        if (h) break;
        if (!n || !*n) continue;
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    }
    if (!h) return set_error(KSN_ECOMM, "cannot load libnccl: %s", dlerror());
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId)) dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank)) dlsym(h, "ncclCommInitRank");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce)) dlsym(h, "ncclAllReduce");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy)) dlsym(h, "ncclCommDestroy");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString)) dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy)
        return set_error(KSN_ECOMM, "libnccl lacks an expected symbol");
    g_nccl.h = h;
    return KSN_OK;
}

static int nccl_fail(int rc, const char *what)
{
    return set_error(KSN_ECOMM, "%s: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "nccl error");
}

static void drop_comm(bool keep_p2p = false)
{
    Ctx &c = g_ctx;
    if (!keep_p2p) p2p_drop();
    if (c.nccl && g_nccl.CommDestroy) g_nccl.CommDestroy(c.nccl);
    c.nccl = nullptr;
    c.cb = nullptr;
    c.cb_user = nullptr;
    c.comm_kind = COMM_SINGLE;
    c.rank = 0;
    c.nranks = 1;
    c.comm_epoch++;
}

void drop_comm_backend(bool keep_p2p) { drop_comm(keep_p2p); }

// The producing kernel has already summed d_buf over the ranks through peer memory (k1_final_p2p_kernel): fetch it.
int reduced_to_host(double *d_buf, double *h_buf, size_t n)
{
    Ctx &c = g_ctx;
    KSN_CUDA(cudaMemcpyAsync(h_buf, d_buf, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    int rc = p2p_status_async();
    if (rc) return rc;
    KSN_CUDA(cudaStreamSynchronize(c.stream));
    return p2p_status_result();
}

int allreduce_to_host(double *d_buf, double *h_buf, size_t n)
{
    Ctx &c = g_ctx;
    if (c.comm_kind == COMM_P2P && c.nranks > 1) {
        phase_begin(PH_COMM);
        int rc = p2p_allreduce_device(d_buf, n);
        phase_end(PH_COMM);
        if (rc) return rc;
        KSN_CUDA(cudaMemcpyAsync(h_buf, d_buf, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        rc = p2p_status_async();
        if (rc) return rc;
        KSN_CUDA(cudaStreamSynchronize(c.stream));
        return p2p_status_result();
    }
    if (c.comm_kind == COMM_NCCL && c.nranks > 1) {
        phase_begin(PH_COMM);
        const int ncclDouble = 8, ncclSum = 0;   // ncclFloat64 / ncclSum enum values (nccl.h)
        int rc = g_nccl.AllReduce(d_buf, d_buf, n, ncclDouble, ncclSum, c.nccl, c.stream);
        phase_end(PH_COMM);
        if (rc) return nccl_fail(rc, "ncclAllReduce");
    }
    KSN_CUDA(cudaMemcpyAsync(h_buf, d_buf, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    KSN_CUDA(cudaStreamSynchronize(c.stream));
    if (c.comm_kind == COMM_HOSTCB && c.nranks > 1) {
        int rc = c.cb(h_buf, n, c.cb_user);
        if (rc) return set_error(KSN_ECOMM, "host all-reduce callback returned %d", rc);
    }
    return KSN_OK;
}

// ---------------------------------------------------------------- pinned-host registry
static std::map<uintptr_t, size_t> g_registered;

int ensure_host_pinned(const void *p, size_t bytes)
{
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    if (at.type == cudaMemoryTypeHost) return 1;
    if (at.type != cudaMemoryTypeUnregistered) return 0;
    // Page-locking memory we do not own is only safe when the owner promises to unregister it
    // before freeing it (a freed-and-reused address would otherwise DMA from stale pages):
    // hosts opt in with ksn_host_register(); KSN_HOST_REGISTER=1 turns it on for every staged grid.
    if (!getenv("KSN_HOST_REGISTER")) return 0;
    auto it = g_registered.find((uintptr_t) p);
    if (it != g_registered.end() && it->second >= bytes) return 1;
    if (it != g_registered.end()) { cudaHostUnregister((void *) p); g_registered.erase(it); }
    e = cudaHostRegister((void *) p, bytes, cudaHostRegisterDefault);
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    g_registered[(uintptr_t) p] = bytes;
    return 1;
}

// cudaFuncSetAttribute once per (kernel, value): the launchers run every PM step, and the attribute calls are a few
// microseconds each of host time between two kernels (forgotten at ksn_shutdown: a new context starts from defaults)
static std::map<const void *, std::pair<size_t, int>> g_func_attr;
int func_attributes(const void *kern, size_t dyn_smem, int carveout)
{
    auto it = g_func_attr.find(kern);
    if (it != g_func_attr.end() && it->second.first == dyn_smem && it->second.second == carveout) return KSN_OK;
    KSN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) dyn_smem));
    if (carveout >= 0) KSN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carveout));
    g_func_attr[kern] = std::make_pair(dyn_smem, carveout);
    return KSN_OK;
}

// was exactly this range page-locked through ksn_host_register (and not unregistered since)?
bool host_range_registered(const void *p, size_t bytes)
{
    auto it = g_registered.find((uintptr_t) p);
    return it != g_registered.end() && it->second >= bytes;
}

// ---------------------------------------------------------------- synthetic grid
__device__ __forceinline__ unsigned long long mix64(unsigned long long x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

template <typename real>
__global__ void fill_synthetic_kernel(real *g, int N, long long plane0, long long nelem,
                                      unsigned long long seed, double slope)
{
    const int L = N / 2 + 1;
    const long long stride = (long long) gridDim.x * blockDim.x;
    for (long long e = (long long) blockIdx.x * blockDim.x + threadIdx.x; e < nelem; e += stride) {
        const long long row = e / L;
        const int z = (int) (e - row * L);
        const int j = (int) (row % N);
        const long long i = plane0 + row / N;
        const long long gidx = (i * N + j) * L + z;          // global mode index: partition independent
        const int ki = i <= N / 2 ? (int) i : (int) (i - N);
        const int kj = j <= N / 2 ? j : j - N;
        const double k2 = (double) ki * ki + (double) kj * kj + (double) z * z;
        double re, im;
        if (gidx == 0) {
            re = (double) N * N * N;
            im = 0;
        } else {
            const unsigned long long h1 = mix64(seed ^ mix64(2 * (unsigned long long) gidx));
            const unsigned long long h2 = mix64(seed ^ mix64(2 * (unsigned long long) gidx + 1));
            const double u1 = ((double) (h1 >> 11) + 0.5) * (1.0 / 9007199254740992.0);
            const double u2 = ((double) (h2 >> 11) + 0.5) * (1.0 / 9007199254740992.0);
            const double r = sqrt(-2.0 * log(u1)) * exp(0.25 * slope * log(k2));
            double s, c;
            sincospi(2.0 * u2, &s, &c);
            re = r * c;
            im = r * s;
        }
        g[2 * e] = (real) re;
        g[2 * e + 1] = (real) im;
    }
}

}  // namespace ksn

using namespace ksn;

extern "C" {

int ksn_init(int device)
{
    std::lock_guard<std::mutex> lk(g_mu);
    return do_init(device);
}

void ksn_shutdown(void)
{
    Ctx &c = g_ctx;
    if (!c.inited) return;
    cudaSetDevice(c.device);
    cudaDeviceSynchronize();
    fft_shutdown();
    k2_prefetch_shutdown();
    g_func_attr.clear();
    drop_comm();
    k1_tables_invalidate();
    for (auto &kv : g_registered) cudaHostUnregister((void *) kv.first);
    g_registered.clear();
    cudaFree(c.d_partial); cudaFree(c.d_red); cudaFreeHost(c.h_red); cudaFree(c.d_thr); cudaFree(c.d_iw); cudaFree(c.d_cold);
    cudaFree(c.d_k3tab); cudaFreeHost(c.h_k3tab); cudaFree(c.d_gz); cudaFree(c.d_stage); cudaFree(c.d_bg);
    cudaFree(c.d_k2); cudaFreeHost(c.h_k2);
    free(c.geom.keff); free(c.geom.count);
    for (int i = 0; i < PH_COUNT; i++) for (int j = 0; j < 2; j++) cudaEventDestroy(c.ev[i][j]);
    cudaStreamDestroy(c.stream);
    cudaStreamDestroy(c.copy_stream);
    if (c.copy_stream2) cudaStreamDestroy(c.copy_stream2);
    cudaFree(c.d_origin);
    c = Ctx();
}

const char *ksn_last_error(void) { return g_err; }

int ksn_device_available(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    return n > 0;
}

int ksn_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int ksn_device(void) { return g_ctx.inited ? g_ctx.device : -1; }

int ksn_comm_single(void) { drop_comm(); return KSN_OK; }

int ksn_comm_nccl_unique_id(void *id128)
{
    if (!id128) return set_error(KSN_EINVAL, "null id buffer");
    int rc = load_nccl();
    if (rc) return rc;
    nccl_uid id;
    int n = g_nccl.GetUniqueId(&id);
    if (n) return nccl_fail(n, "ncclGetUniqueId");
    memcpy(id128, &id, sizeof(id));
    return KSN_OK;
}

int ksn_comm_nccl_init(const void *id128, int nranks, int rank)
{
    if (!id128 || nranks < 1 || rank < 0 || rank >= nranks) return set_error(KSN_EINVAL, "bad NCCL rank/size %d/%d", rank, nranks);
    int rc = ensure_init();
    if (rc) return rc;
    rc = load_nccl();
    if (rc) return rc;
    drop_comm();
    nccl_uid id;
    memcpy(&id, id128, sizeof(id));
    ncclComm *comm = nullptr;
    int n = g_nccl.CommInitRank(&comm, nranks, id, rank);
    if (n) return nccl_fail(n, "ncclCommInitRank");
    Ctx &c = g_ctx;
    c.nccl = comm;
    c.comm_kind = COMM_NCCL;
    c.rank = rank;
    c.nranks = nranks;
    c.comm_epoch++;
    return KSN_OK;
}

int ksn_comm_host_callback(ksn_allreduce_fn fn, void *user, int nranks, int rank)
{
    if (!fn || nranks < 1 || rank < 0 || rank >= nranks) return set_error(KSN_EINVAL, "bad host all-reduce callback setup");
    drop_comm();
    Ctx &c = g_ctx;
    c.cb = fn;
    c.cb_user = user;
    c.comm_kind = COMM_HOSTCB;
    c.rank = rank;
    c.nranks = nranks;
    c.comm_epoch++;
    return KSN_OK;
}

int ksn_comm_allreduce_host(double *buf, size_t n)
{
    Ctx &c = g_ctx;
    if (!buf && n) return set_error(KSN_EINVAL, "ksn_comm_allreduce_host: null buffer");
    if (c.nranks == 1 || n == 0) return KSN_OK;
    if (c.comm_kind == COMM_HOSTCB) {
        int rc = c.cb(buf, n, c.cb_user);
        return rc ? set_error(KSN_ECOMM, "host all-reduce callback returned %d", rc) : KSN_OK;
    }
    // NCCL backend: bounce through the device reduce buffer
    int rc = ensure_init();
    if (rc) return rc;
    rc = ensure_device_buffer((void **) &c.d_red, &c.red_cap, n * sizeof(double));
    if (rc) return rc;
    rc = ensure_pinned_buffer((void **) &c.h_red, &c.h_red_cap, n * sizeof(double));
    if (rc) return rc;
    KSN_CUDA(cudaMemcpyAsync(c.d_red, buf, n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    rc = allreduce_to_host(c.d_red, c.h_red, n);
    if (rc) return rc;
    memcpy(buf, c.h_red, n * sizeof(double));
    return KSN_OK;
}

int ksn_comm_rank(void) { return g_ctx.rank; }
int ksn_comm_size(void) { return g_ctx.nranks; }

int ksn_device_malloc(void **ptr, size_t bytes)
{
    int rc = ensure_init();
    if (rc) return rc;
    KSN_CUDA(cudaMalloc(ptr, bytes));
    return KSN_OK;
}

int ksn_device_free(void *ptr)
{
    int rc = ensure_init();
    if (rc) return rc;
    KSN_CUDA(cudaFree(ptr));
    return KSN_OK;
}

int ksn_host_alloc_pinned(void **ptr, size_t bytes)
{
    int rc = ensure_init();
    if (rc) return rc;
    KSN_CUDA(cudaHostAlloc(ptr, bytes, cudaHostAllocDefault));
    return KSN_OK;
}

int ksn_host_free_pinned(void *ptr)
{
    int rc = ensure_init();
    if (rc) return rc;
    KSN_CUDA(cudaFreeHost(ptr));
    return KSN_OK;
}

static int copy_sync(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind)
{
    int rc = ensure_init();
    if (rc) return rc;
    KSN_CUDA(cudaMemcpyAsync(dst, src, bytes, kind, g_ctx.stream));
    KSN_CUDA(cudaStreamSynchronize(g_ctx.stream));
    return KSN_OK;
}

int ksn_host_register(void *ptr, size_t bytes)
{
    int rc = ensure_init();
    if (rc) return rc;
    auto it = g_registered.find((uintptr_t) ptr);
    if (it != g_registered.end()) { cudaHostUnregister(ptr); g_registered.erase(it); }
    KSN_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
    g_registered[(uintptr_t) ptr] = bytes;
    return KSN_OK;
}

int ksn_host_unregister(void *ptr)
{
    int rc = ensure_init();
    if (rc) return rc;
    auto it = g_registered.find((uintptr_t) ptr);
    if (it == g_registered.end()) return set_error(KSN_EINVAL, "ksn_host_unregister: pointer was not registered here");
    KSN_CUDA(cudaHostUnregister(ptr));
    g_registered.erase(it);
    return KSN_OK;
}

int ksn_memcpy_h2d(void *dst, const void *src, size_t bytes) { return copy_sync(dst, src, bytes, cudaMemcpyHostToDevice); }
int ksn_memcpy_d2h(void *dst, const void *src, size_t bytes) { return copy_sync(dst, src, bytes, cudaMemcpyDeviceToHost); }
int ksn_memcpy_d2d(void *dst, const void *src, size_t bytes) { return copy_sync(dst, src, bytes, cudaMemcpyDeviceToDevice); }

int ksn_device_synchronize(void)
{
    int rc = ensure_init();
    if (rc) return rc;
    KSN_CUDA(cudaStreamSynchronize(g_ctx.stream));
    KSN_CUDA(cudaStreamSynchronize(g_ctx.copy_stream));
    return KSN_OK;
}

int ksn_pointer_is_device(const void *ptr)
{
    if (ensure_init()) return 0;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) { cudaGetLastError(); return 0; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

int ksn_fill_synthetic_grid(void *dgrid, int real_bytes, int dims, long long startslab, long long nslab,
                            unsigned long long seed, double slope)
{
    int rc = ensure_init();
    if (rc) return rc;
    if (!dgrid || dims < 2 || (real_bytes != 4 && real_bytes != 8) || nslab < 0)
        return set_error(KSN_EINVAL, "ksn_fill_synthetic_grid: bad arguments");
    const long long nelem = nslab * dims * (dims / 2 + 1);
    if (nelem == 0) return KSN_OK;
    const int threads = 256;
    const int blocks = (int) fmin((double) ((nelem + threads - 1) / threads), (double) g_ctx.num_sms * 16);
    if (real_bytes == 8)
        fill_synthetic_kernel<double><<<blocks, threads, 0, g_ctx.stream>>>((double *) dgrid, dims, startslab, nelem, seed, slope);
    else
        fill_synthetic_kernel<float><<<blocks, threads, 0, g_ctx.stream>>>((float *) dgrid, dims, startslab, nelem, seed, slope);
    KSN_CUDA(cudaGetLastError());
    KSN_CUDA(cudaStreamSynchronize(g_ctx.stream));
    return KSN_OK;
}

int ksn_timing_enable(int on)
{
    int rc = ensure_init();
    if (rc) return rc;
    g_ctx.timing = on != 0;
    return KSN_OK;
}

int ksn_timing_reset(void)
{
    Ctx &c = g_ctx;
    phase_collect();
    for (int p = 0; p < PH_COUNT; p++) c.acc_ms[p] = 0;
    c.launches = 0;
    return KSN_OK;
}

int ksn_timing_get(ksn_timing *out)
{
    if (!out) return set_error(KSN_EINVAL, "null timing struct");
    Ctx &c = g_ctx;
    if (c.inited) { cudaStreamSynchronize(c.stream); cudaStreamSynchronize(c.copy_stream); }
    phase_collect();
    out->k1_ms = c.acc_ms[PH_K1];
    out->k1_reduce_ms = c.acc_ms[PH_K1RED];
    out->comm_ms = c.acc_ms[PH_COMM];
    out->k2_ms = c.acc_ms[PH_K2];
    out->k3_ms = c.acc_ms[PH_K3];
    out->h2d_ms = c.acc_ms[PH_H2D];
    out->d2h_ms = c.acc_ms[PH_D2H];
    out->launches = c.launches;
    return KSN_OK;
}

void *ksn_stream(void) { return ensure_init() ? nullptr : (void *) g_ctx.stream; }

}  // extern "C"
