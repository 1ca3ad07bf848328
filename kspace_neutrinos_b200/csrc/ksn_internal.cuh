// Internal declarations shared by the CUDA translation units of libkspace_neutrinos_b200.
// Nothing here is part of the C-ABI (see include/ksn_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stddef.h>
#include <stdint.h>
#include "ksn_b200.h"

struct ncclComm;

namespace ksn {

// c.d_iw holds iw[0..L) followed by K1's z-weight table of L entries and this many zeros
constexpr int K1_WZ_PAD = 32 * 65 + 16;

enum CommKind { COMM_SINGLE = 0, COMM_NCCL = 1, COMM_HOSTCB = 2, COMM_P2P = 3 };

// Events used for the optional per-phase timing (ksn_timing_*).
enum Phase { PH_K1 = 0, PH_K1RED, PH_COMM, PH_K2, PH_K3, PH_H2D, PH_D2H, PH_COUNT };

struct Ctx {
    bool inited = false;
    int device = -1;
    int num_sms = 0;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;       // every kernel is launched here
    cudaStream_t copy_stream = nullptr;  // staged H2D / D2H
    // collective backend
    int comm_kind = COMM_SINGLE;
    int rank = 0, nranks = 1;
    ncclComm *nccl = nullptr;
    ksn_allreduce_fn cb = nullptr;
    void *cb_user = nullptr;
    unsigned long long comm_epoch = 0;   // bumped whenever the backend changes (invalidates geometry cache)
    // K1 workspace
    double *d_partial = nullptr; size_t partial_cap = 0;   // [ctas][stride]
    double *d_red = nullptr;     size_t red_cap = 0;       // power | mass2 | keff | count
    double *h_red = nullptr;     size_t h_red_cap = 0;     // pinned mirror of d_red
    unsigned int *d_thr = nullptr; double *d_iw = nullptr; size_t thr_cap = 0, iw_cap = 0;
    double *d_cold = nullptr; size_t cold_cap = 0;         // per-warp bins below the bin window (k1_tile_kernel<.., true>)
    // geometry cache (global sums over all ranks)
    struct { bool valid = false; int dims = 0, nrbins = 0; long long startslab = 0, nslab = 0;
             unsigned long long epoch = 0, h_thr = 0; double *keff = nullptr; long long *count = nullptr; size_t cap = 0; } geom;
    // K3 workspace
    double *d_k3tab = nullptr; size_t k3tab_cap = 0; double *h_k3tab = nullptr; size_t h_k3tab_cap = 0;
    double *d_gz = nullptr; size_t gz_cap = 0;             // z factor of the fused Green's function (k3_set_greens)
    // staging buffer for host-resident grids
    void *d_stage = nullptr; size_t stage_cap = 0;
    bool stage_streaming = false;        // the slab did not fit: d_stage is a ring of chunks, K3 re-uploads (stage_plan)
    cudaStream_t copy_stream2 = nullptr; // D2H of the streaming path (H2D keeps copy_stream: PCIe is full duplex)
    void *d_origin = nullptr;            // copy of the slab's first element (total_mass2) when the ring recycles chunk 0
    // background table 1/(aH)
    double *d_bg = nullptr; int bg_n = 0; double bg_lo = 0, bg_hi = 0, bg_h = 0;
    // K2 workspace
    double *d_k2 = nullptr; size_t k2_cap = 0; double *h_k2 = nullptr; size_t h_k2_cap = 0;
    // timing
    bool timing = false;
    cudaEvent_t ev[PH_COUNT][2] = {};
    bool ev_used[PH_COUNT] = {};
    float acc_ms[PH_COUNT] = {};
    unsigned long long launches = 0;
};

Ctx &ctx();
int ensure_init();
int set_error(int code, const char *fmt, ...);
int check_cuda(cudaError_t e, const char *what);
int ensure_device_buffer(void **p, size_t *cap, size_t bytes);
int ensure_pinned_buffer(void **p, size_t *cap, size_t bytes);
void phase_begin(Phase p);
void phase_end(Phase p);
void phase_collect();   // after a stream sync: fold event times into acc_ms

#define KSN_CUDA(call) do { int _rc = ::ksn::check_cuda((call), #call); if (_rc) return _rc; } while (0)

// Forget the active collective backend (keep_p2p: all but the peer-memory mailboxes, which ksn_comm_p2p_init is setting up).
void drop_comm_backend(bool keep_p2p);

// Device allreduce-or-host-callback of n doubles living at d_buf (device) with pinned mirror h_buf.
// On return h_buf holds the global sums (and the stream is synchronized).
int allreduce_to_host(double *d_buf, double *h_buf, size_t n);
int reduced_to_host(double *d_buf, double *h_buf, size_t n);   // d_buf is already the global sum (fused peer-memory reduce)

// Host pointer -> pinned (registers once and remembers).  Returns 1 if the range is pinned afterwards.
int ensure_host_pinned(const void *p, size_t bytes);

// How a host-resident slab goes through HBM: whole (it stays resident between K1 and K3) or as a ring of chunks.
constexpr int STAGE_RING = 3;
struct StagePlan { size_t plane_bytes, total; long long chunk; int nchunks; bool streaming; };
int stage_plan(int real_bytes, int dims, long long nslab, StagePlan *plan);   // allocates c.d_stage accordingly

// launchers (one per kernel family)
int k1_launch(const void *dgrid, int real_bytes, int dims, int nrbins, long long plane0_global, long long nplanes,
              bool full, bool accumulate, int *ctas_out, int *stride_out);
int k1_finish(int real_bytes, int dims, int nrbins, bool full, int ctas, int stride, const void *origin_elem, bool fuse_p2p);
void k1_tables_invalidate();   // somebody else wrote c.d_iw / c.d_thr, or the context is gone
void fft_shutdown();           // plans, work area and peer mappings of ksn_fft_* (ksn_fft.cu)
void k2_prefetch_shutdown();   // side stream and tables of ksn_delta_nu_prefetch (k2_delta_nu.cu)
int k2_prefetch_launch_pending();   // launch a recorded prefetch request on its side stream (called once K1 is in flight)
bool host_range_registered(const void *p, size_t bytes);   // page-locked through ksn_host_register?
int func_attributes(const void *kern, size_t dyn_smem, int carveout);   // cudaFuncSetAttribute, once per value (carveout < 0: leave)
int k3_launch(void *dgrid, int real_bytes, int dims, long long plane0_global, long long nplanes, int nknots);
int k3_upload_table(int dims, double boxsize, const double *logkk, const double *ratio, int nbins, double norm);

}  // namespace ksn
