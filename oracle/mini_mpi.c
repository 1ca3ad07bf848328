/* TEST INFRASTRUCTURE (oracle): fork + shared-memory implementation of the handful of
 * MPI calls the reference makes.  Not part of the product; the product never links it.
 * One process per rank, a process-shared barrier and a bounce buffer in MAP_SHARED
 * memory.  Reductions are summed in rank order by every rank, so all ranks hold
 * bit-identical results (like MPI_Allreduce on most implementations). */
#define _GNU_SOURCE
#include "shim/mpi.h"
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#define BOUNCE_BYTES (8u << 20)

struct world {
    pthread_barrier_t bar;
    int size;
    pid_t pids[256];
    _Alignas(64) unsigned char bounce[];      /* size * BOUNCE_BYTES; aligned for the typed reads of MPI_Allreduce */
};

static struct world *W;
static int my_rank = 0, my_size = 1;

static size_t tsize(MPI_Datatype t)
{
    switch (t) {
    case MPI_INT: return sizeof(int);
    case MPI_DOUBLE: return sizeof(double);
    case MPI_LONG_LONG_INT: return sizeof(long long);
    default: return 1;
    }
}

void *ksn_minimpi_shared_alloc(size_t bytes)
{
    void *p = mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    return p == MAP_FAILED ? NULL : p;
}

int ksn_minimpi_fork(int nranks)
{
    if (nranks <= 1) return 0;
    if (nranks > 256) nranks = 256;
    W = ksn_minimpi_shared_alloc(sizeof(struct world) + (size_t) nranks * BOUNCE_BYTES);
    if (!W) { perror("mmap"); exit(1); }
    pthread_barrierattr_t at;
    pthread_barrierattr_init(&at);
    pthread_barrierattr_setpshared(&at, PTHREAD_PROCESS_SHARED);
    pthread_barrier_init(&W->bar, &at, nranks);
    W->size = nranks;
    my_size = nranks;
    fflush(NULL);
    for (int r = 1; r < nranks; r++) {
        pid_t p = fork();
        if (p < 0) { perror("fork"); exit(1); }
        if (p == 0) { my_rank = r; return r; }
        W->pids[r] = p;
    }
    my_rank = 0;
    return 0;
}

void ksn_minimpi_exit(int code)
{
    fflush(NULL);
    if (my_rank != 0) _exit(code);
    for (int r = 1; r < my_size; r++) {
        int st;
        waitpid(W->pids[r], &st, 0);
    }
}

int MPI_Init(int *argc, char ***argv) { (void) argc; (void) argv; return MPI_SUCCESS; }
int MPI_Finalize(void) { return MPI_SUCCESS; }
int MPI_Abort(MPI_Comm comm, int code) { (void) comm; _exit(code); }
int MPI_Comm_rank(MPI_Comm comm, int *rank) { (void) comm; *rank = my_rank; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm comm, int *size) { (void) comm; *size = my_size; return MPI_SUCCESS; }

int MPI_Barrier(MPI_Comm comm)
{
    (void) comm;
    if (my_size > 1) pthread_barrier_wait(&W->bar);
    return MPI_SUCCESS;
}

int MPI_Bcast(void *buf, int count, MPI_Datatype type, int root, MPI_Comm comm)
{
    (void) comm;
    if (my_size == 1) return MPI_SUCCESS;
    size_t left = (size_t) count * tsize(type), off = 0;
    while (left) {
        size_t n = left < BOUNCE_BYTES ? left : BOUNCE_BYTES;
        if (my_rank == root) memcpy(W->bounce, (char *) buf + off, n);
        pthread_barrier_wait(&W->bar);
        if (my_rank != root) memcpy((char *) buf + off, W->bounce, n);
        pthread_barrier_wait(&W->bar);
        left -= n; off += n;
    }
    return MPI_SUCCESS;
}

int MPI_Allgather(const void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf, int recvcount, MPI_Datatype recvtype, MPI_Comm comm)
{
    (void) comm; (void) recvcount; (void) recvtype;
    const size_t n = (size_t) sendcount * tsize(sendtype);
    if (n > BOUNCE_BYTES) { fprintf(stderr, "mini_mpi: MPI_Allgather of %zu bytes per rank\n", n); _exit(99); }
    if (my_size == 1) {
        if (sendbuf != recvbuf) memcpy(recvbuf, sendbuf, n);
        return MPI_SUCCESS;
    }
    memcpy(W->bounce + (size_t) my_rank * BOUNCE_BYTES, sendbuf, n);
    pthread_barrier_wait(&W->bar);
    for (int r = 0; r < my_size; r++) memcpy((char *) recvbuf + (size_t) r * n, W->bounce + (size_t) r * BOUNCE_BYTES, n);
    pthread_barrier_wait(&W->bar);
    return MPI_SUCCESS;
}

int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm)
{
    (void) comm; (void) op;
    const size_t es = tsize(type);
    if (sendbuf == MPI_IN_PLACE) sendbuf = recvbuf;
    if (my_size == 1) {
        if (sendbuf != recvbuf) memcpy(recvbuf, sendbuf, (size_t) count * es);
        return MPI_SUCCESS;
    }
    const int per = (int) (BOUNCE_BYTES / es);
    for (int done = 0; done < count; done += per) {
        const int n = count - done < per ? count - done : per;
        memcpy(W->bounce + (size_t) my_rank * BOUNCE_BYTES, (const char *) sendbuf + (size_t) done * es, (size_t) n * es);
        pthread_barrier_wait(&W->bar);
        for (int i = 0; i < n; i++) {
            if (type == MPI_DOUBLE) {
                double s = 0;
                for (int r = 0; r < my_size; r++) s += ((double *) (W->bounce + (size_t) r * BOUNCE_BYTES))[i];
                ((double *) recvbuf)[done + i] = s;
            } else if (type == MPI_LONG_LONG_INT) {
                long long s = 0;
                for (int r = 0; r < my_size; r++) s += ((long long *) (W->bounce + (size_t) r * BOUNCE_BYTES))[i];
                ((long long *) recvbuf)[done + i] = s;
            } else if (type == MPI_INT) {
                int s = 0;
                for (int r = 0; r < my_size; r++) s += ((int *) (W->bounce + (size_t) r * BOUNCE_BYTES))[i];
                ((int *) recvbuf)[done + i] = s;
            } else {
                fprintf(stderr, "mini_mpi: unsupported reduce type %d\n", type);
                _exit(99);
            }
        }
        pthread_barrier_wait(&W->bar);
    }
    return MPI_SUCCESS;
}
