// Peer-memory collective backend (COMM_P2P): mailbox allocation, CUDA-IPC exchange, and the stand-alone all-reduce kernel
// for callers that have no producing kernel to fuse it into.  Protocol and device code: ksn_p2p.cuh.
#include "ksn_internal.cuh"
#include "ksn_p2p.cuh"

#include <string.h>

namespace ksn {

struct P2PState {
    bool active = false;
    int R = 1, rank = 0;
    double *own = nullptr;                       // this rank's mailbox (cudaMalloc)
    double *box[KSN_P2P_MAX_RANKS] = {};         // all mailboxes as mapped here
    bool opened[KSN_P2P_MAX_RANKS] = {};         // box[r] came from cudaIpcOpenMemHandle
    unsigned *d_counter = nullptr;
    double *d_status = nullptr;                  // written by a kernel only when a wait timed out
    double *h_status = nullptr;                  // pinned mirror, fetched with the results
    unsigned long long seq = 0;
};
static P2PState g_p2p;

static size_t mailbox_bytes()
{
    return (size_t) 2 * KSN_P2P_MAX_RANKS * KSN_P2P_SLOT * sizeof(double) + (size_t) 2 * KSN_P2P_MAX_RANKS * sizeof(unsigned long long);
}

bool p2p_active() { return g_p2p.active && g_p2p.R > 1; }

void p2p_drop()
{
    P2PState &s = g_p2p;
    for (int r = 0; r < KSN_P2P_MAX_RANKS; r++) {
        if (s.opened[r] && s.box[r]) cudaIpcCloseMemHandle(s.box[r]);
        s.box[r] = nullptr;
        s.opened[r] = false;
    }
    if (s.own) cudaFree(s.own);
    if (s.d_counter) cudaFree(s.d_counter);
    if (s.h_status) cudaFreeHost(s.h_status);
    s = P2PState();
}

int p2p_next_round(P2PDev *dev)
{
    P2PState &s = g_p2p;
    if (!p2p_active()) return set_error(KSN_ECOMM, "peer-memory backend is not initialised");
    s.seq++;
    for (int r = 0; r < KSN_P2P_MAX_RANKS; r++) dev->box[r] = s.box[r];
    dev->R = s.R;
    dev->rank = s.rank;
    dev->seq = s.seq;
    dev->counter = s.d_counter;
    dev->status = s.d_status;
    return KSN_OK;
}

// Stand-alone round, one block: push n values of buf, raise the flags, wait, sum in rank order back into buf.
__global__ void __launch_bounds__(1024)
p2p_allreduce_kernel(const P2PDev p, double *__restrict__ buf, size_t n)
{
    for (size_t j = threadIdx.x; j < n; j += blockDim.x) p2p_push(p, j, buf[j]);
    __syncthreads();
    p2p_finish(p, n, buf);
}

int p2p_allreduce_device(double *d_buf, size_t n)
{
    Ctx &c = ctx();
    for (size_t done = 0; done < n; done += KSN_P2P_SLOT) {       // longer vectors go in rounds of one slot each
        const size_t m = n - done < KSN_P2P_SLOT ? n - done : KSN_P2P_SLOT;
        P2PDev dev;
        int rc = p2p_next_round(&dev);
        if (rc) return rc;
        p2p_allreduce_kernel<<<1, 1024, 0, c.stream>>>(dev, d_buf + done, m);
        c.launches++;
        KSN_CUDA(cudaGetLastError());
    }
    return KSN_OK;
}

// The status word travels with the results: enqueue its copy before the stream is synchronised, look at it afterwards.
int p2p_status_async()
{
    if (!p2p_active()) return KSN_OK;
    KSN_CUDA(cudaMemcpyAsync(g_p2p.h_status, g_p2p.d_status, sizeof(double), cudaMemcpyDeviceToHost, ctx().stream));
    return KSN_OK;
}

int p2p_status_result()
{
    if (!p2p_active() || *g_p2p.h_status == 0.0) return KSN_OK;
    *g_p2p.h_status = 0.0;
    cudaMemset(g_p2p.d_status, 0, sizeof(double));
    return set_error(KSN_ECOMM, "peer-memory all-reduce: a rank did not arrive within the time-out");
}

}  // namespace ksn

using namespace ksn;

extern "C" int ksn_comm_p2p_export(void *handle64)
{
    if (!handle64) return set_error(KSN_EINVAL, "ksn_comm_p2p_export: null handle buffer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    int rc = ensure_init();
    if (rc) return rc;
    P2PState &s = g_p2p;
    // always a fresh, zeroed mailbox: the handle is about to be handed out, so nobody can be writing into it yet (a
    // mailbox that had been in use would still hold the flags of its old rounds)
    if (ctx().comm_kind == COMM_P2P) drop_comm_backend(false); else p2p_drop();
    KSN_CUDA(cudaMalloc((void **) &s.own, mailbox_bytes()));
    KSN_CUDA(cudaMemset(s.own, 0, mailbox_bytes()));
    KSN_CUDA(cudaMalloc((void **) &s.d_counter, 64));
    KSN_CUDA(cudaMemset(s.d_counter, 0, 64));
    s.d_status = (double *) (s.d_counter + 2);              // 8-byte aligned word of the same allocation
    KSN_CUDA(cudaHostAlloc((void **) &s.h_status, sizeof(double), cudaHostAllocDefault));
    *s.h_status = 0.0;
    KSN_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    KSN_CUDA(cudaIpcGetMemHandle(&h, s.own));
    memcpy(handle64, &h, sizeof h);
    return KSN_OK;
}

extern "C" int ksn_comm_p2p_init(const void *handles, int nranks, int rank)
{
    if (!handles || nranks < 1 || nranks > KSN_P2P_MAX_RANKS || rank < 0 || rank >= nranks)
        return set_error(KSN_EINVAL, "ksn_comm_p2p_init: bad rank/size %d/%d (at most %d ranks)", rank, nranks, KSN_P2P_MAX_RANKS);
    int rc = ensure_init();
    if (rc) return rc;
    P2PState &s = g_p2p;
    if (!s.own) return set_error(KSN_EINVAL, "ksn_comm_p2p_init: call ksn_comm_p2p_export first");
    drop_comm_backend(true);                   // NCCL communicator / host callback, if any; the mailbox stays
    for (int r = 0; r < KSN_P2P_MAX_RANKS; r++) {          // a second init re-maps
        if (s.opened[r] && s.box[r]) cudaIpcCloseMemHandle(s.box[r]);
        s.box[r] = nullptr;
        s.opened[r] = false;
    }
    for (int r = 0; r < nranks; r++) {
        if (r == rank) { s.box[r] = s.own; s.opened[r] = false; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *) handles + (size_t) r * sizeof h, sizeof h);
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return set_error(KSN_ECOMM, "cudaIpcOpenMemHandle(rank %d): %s (peer access between the GPUs of one box is required)", r, cudaGetErrorString(e));
        }
        s.box[r] = (double *) p;
        s.opened[r] = true;
    }
    s.R = nranks;
    s.rank = rank;
    s.seq = 0;
    s.active = true;
    Ctx &c = ctx();
    c.comm_kind = COMM_P2P;
    c.rank = rank;
    c.nranks = nranks;
    c.comm_epoch++;
    return KSN_OK;
}
