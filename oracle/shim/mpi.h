/* TEST INFRASTRUCTURE (oracle): minimal MPI stand-in so the reference sources under
 * /root/reference compile unmodified in an image with no MPI.  Implemented in
 * oracle/mini_mpi.c: rank/size come from ksn_minimpi_fork(); with one rank every
 * collective degenerates to a copy.  Only the calls the reference makes are declared
 * (powerspectrum.c:91-95, interface_common.c:56-73,186, interface_gadget.c:189), plus MPI_Allgather for the
 * product's own host layer (kspace_neutrinos_b200/src/iface_common.c: backend bootstrap). */
#ifndef KSN_ORACLE_MPI_SHIM_H
#define KSN_ORACLE_MPI_SHIM_H
#include <stddef.h>
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
#define MPI_COMM_WORLD 0
#define MPI_INT 1
#define MPI_DOUBLE 2
#define MPI_LONG_LONG_INT 3
#define MPI_BYTE 4
#define MPI_SUM 1
#define MPI_IN_PLACE ((void *) 1)
#define MPI_SUCCESS 0
int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm comm, int code);
int MPI_Barrier(MPI_Comm comm);
int MPI_Bcast(void *buf, int count, MPI_Datatype type, int root, MPI_Comm comm);
int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm);
int MPI_Allgather(const void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf, int recvcount, MPI_Datatype recvtype, MPI_Comm comm);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_size(MPI_Comm comm, int *size);
/* oracle-only: fork nranks-1 children sharing a MAP_SHARED scratch; returns this
 * process's rank.  Children must finish with ksn_minimpi_exit(). */
int ksn_minimpi_fork(int nranks);
void ksn_minimpi_exit(int code);
/* shared anonymous memory visible to every rank forked later */
void *ksn_minimpi_shared_alloc(size_t bytes);
#endif
