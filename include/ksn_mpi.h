/* MPI_Comm for the drop-in boundary.
 *
 * The reference passes an MPI communicator through its public calls
 * (interface_common.h:46,66,80,108; interface_gadget.h:38,48; powerspectrum.h:30).
 * A host that has MPI builds this library with -DKSN_HAVE_MPI and gets the real type;
 * the bin sums are then all-reduced with MPI_Allreduce on that communicator.
 * Without MPI (this image has none) MPI_Comm is an int handle and the collective is
 * whichever backend ksn_comm_*() in ksn_b200.h configured: none (one rank), NCCL over
 * NVLink (one process per GPU), or a host all-reduce callback. */
#ifndef KSN_MPI_H
#define KSN_MPI_H
#ifdef KSN_HAVE_MPI
#include <mpi.h>
#else
#ifndef MPI_COMM_WORLD
typedef int MPI_Comm;
#define MPI_COMM_WORLD ((MPI_Comm) 0)
#endif
#endif
#endif
