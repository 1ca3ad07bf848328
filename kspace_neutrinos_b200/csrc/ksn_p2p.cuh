// Peer-memory all-reduce of the bin sums (the path's one exchange step, powerspectrum.c:91-95's four MPI_Allreduce),
// written INTO the kernel that produces them instead of being a library call behind it.
//
// One process per GPU.  Every rank owns a MAILBOX in its HBM, exported to the other processes of the box through CUDA
// IPC (ksn_comm_p2p_export / ksn_comm_p2p_init) and mapped by them over NVLink / NVSwitch:
//
//     mailbox = data[2 parities][R source ranks][slot doubles] | flags[2][R] (u64)
//
// Round `seq` (parity seq & 1): each rank stores its values into slot [its own rank] of EVERY rank's mailbox as soon as
// they are computed (peer stores, st.relaxed.sys), then -- once all of them are issued and fenced -- raises flag
// [its own rank] = seq in every mailbox (st.release.sys).  A rank that has seen all R flags of its own mailbox
// (ld.acquire.sys) adds the R slots in RANK ORDER: every rank adds the same numbers in the same order, so the sums are
// bit-identical across ranks and from run to run (ncclAllReduce promises neither).  Two parities suffice: nobody can
// start round seq+2 before every rank has finished reading round seq (it needs their seq+1 flags for that).
// Cost: one NVLink store latency plus one flag latency (a few us) instead of a collective launch (~80 us measured for
// the 8-GPU ncclAllReduce of 8 KB behind k1_final_kernel).
#pragma once
#include <stddef.h>

namespace ksn {

constexpr int KSN_P2P_MAX_RANKS = 16;
constexpr size_t KSN_P2P_SLOT = 16384;          // doubles per (parity, source) slot: 3 * nrbins + 1 up to PMGRID = 8192
// ~20 s at 1.9 GHz: a peer that never arrives is an error, not a hang -- but ranks of a real run may reach the step seconds
// apart (page-locking a host slab, a snapshot being written), and that must not be one
constexpr long long KSN_P2P_TIMEOUT_CYCLES = 40000000000LL;

struct P2PDev {
    double *box[KSN_P2P_MAX_RANKS];   // every rank's mailbox as mapped in this process (box[rank] is this rank's own)
    int R, rank;
    unsigned long long seq;           // round number (>= 1)
    unsigned *counter;                // blocks of the producing kernel that have pushed (local, zero between launches)
    double *status;                   // set non-zero on a time-out (local)
};

#ifdef __CUDACC__
__device__ __forceinline__ size_t p2p_data_off(const P2PDev &p, int src) { return ((size_t) (p.seq & 1) * p.R + src) * KSN_P2P_SLOT; }
__device__ __forceinline__ unsigned long long *p2p_flag(const P2PDev &p, double *box, int src)
{
    return (unsigned long long *) (box + 2 * (size_t) p.R * KSN_P2P_SLOT) + (p.seq & 1) * p.R + src;
}

// this rank's value j -> slot [rank] of every mailbox
__device__ __forceinline__ void p2p_push(const P2PDev &p, size_t j, double v)
{
    const size_t off = p2p_data_off(p, p.rank) + j;
    for (int r = 0; r < p.R; r++) asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p.box[r] + off), "d"(v) : "memory");
}

// Called by every thread of a block once it has issued its pushes.  Returns true in exactly one block of the grid: the
// last one to get here, at which point all pushes of this rank are ordered before whatever that block does next.
__device__ __forceinline__ bool p2p_last_block(const P2PDev &p)
{
    __shared__ int last_s;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        last_s = atomicAdd(p.counter, 1u) == gridDim.x - 1;
        __threadfence_system();
        if (last_s) *p.counter = 0;           // the next launch on this stream starts from zero
    }
    __syncthreads();
    return last_s != 0;
}

// One block: raise this rank's flag everywhere, wait for everybody's, out[j] = sum over ranks (in rank order) of value j.
__device__ __forceinline__ void p2p_finish(const P2PDev &p, size_t n, double *__restrict__ out)
{
    __shared__ int ok_s;
    const int t = threadIdx.x;
    if (t == 0) ok_s = 1;
    __syncthreads();
    if (t < p.R) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p2p_flag(p, p.box[t], p.rank)), "l"(p.seq) : "memory");
        const unsigned long long *mine = p2p_flag(p, p.box[p.rank], t);
        const long long t0 = clock64();
        for (;;) {
            unsigned long long f;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(f) : "l"(mine) : "memory");
            if (f >= p.seq) break;
            if (clock64() - t0 > KSN_P2P_TIMEOUT_CYCLES) { ok_s = 0; break; }
            __nanosleep(64);
        }
    }
    __syncthreads();
    if (!ok_s) {
        if (t == 0) *p.status = 1.0;          // (the host clears it when it reports the error)
        return;
    }
    const double *own = p.box[p.rank];
    for (size_t j = t; j < n; j += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < p.R; r++) {
            double v;
            asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(own + p2p_data_off(p, r) + j) : "memory");
            s += v;
        }
        out[j] = s;
    }
}
#endif

// host side (ksn_p2p.cu)
bool p2p_active();
int p2p_next_round(P2PDev *dev);                               // fills *dev for the next collective round
int p2p_allreduce_device(double *d_buf, size_t n);             // generic: d_buf <- sum over ranks, on the library's stream
int p2p_status_async();                                        // enqueue the copy of the time-out word (before the stream sync)
int p2p_status_result();                                       // ... and look at it (after)
void p2p_drop();

}  // namespace ksn
