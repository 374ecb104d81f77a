/*
 * Single-rank stand-in for <mpi.h>, for machines without an MPI installation: used by the driver build
 * (tdvmc_b200/host/driver/Makefile, one process driving one GPU) and by the CPU-oracle build of the unmodified
 * reference (oracle/ref_build/Makefile).  With a real MPI, build with MPI_INC= CXX=mpicxx instead.
 *
 * The reference (mathiasgartner/TDVMC) is an MPI program whose only data-path
 * collectives are MPI_Reduce(SUM)->root and MPI_Bcast (src/MPIMethods.h:132-206,
 * 329-362; src/TDVMC.cpp:506-512).  With exactly one rank every collective is a
 * copy (send and receive buffers are always distinct in the reference) or a
 * no-op, which is what this header provides so that the reference sources
 * compile with plain g++ where no MPI installation exists.
 *
 * The reference's MPIMethods.h relies on <mpi.h> transitively pulling in a few
 * standard headers (cout, map, strcpy, invalid_argument), hence the includes.
 */
#ifndef TDVMC_MPI_SINGLE_RANK_H
#define TDVMC_MPI_SINGLE_RANK_H

#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <stdexcept>
#include <string>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;

enum { MPI_COMM_WORLD = 0 };
enum { MPI_SUCCESS = 0 };
enum { MPI_MAX_PROCESSOR_NAME = 256 };
/* datatype handle == size in bytes, so a collective is memcpy(count * handle) */
enum { MPI_CHAR = 1, MPI_INT = 4, MPI_DOUBLE = 8, MPI_LONG_LONG_INT = 8 + 1024 };
enum { MPI_SUM = 1, MPI_MIN = 2, MPI_MAX = 3 };

static inline size_t tdvmc_shim_bytes(int count, MPI_Datatype t)
{
    return (size_t)count * (size_t)(t >= 1024 ? t - 1024 : t);
}

static inline int MPI_Init(int*, char***) { return MPI_SUCCESS; }
static inline int MPI_Finalize() { return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm, int* rank) { *rank = 0; return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm, int* size) { *size = 1; return MPI_SUCCESS; }
static inline int MPI_Barrier(MPI_Comm) { return MPI_SUCCESS; }
static inline int MPI_Abort(MPI_Comm, int code) { std::exit(code); return MPI_SUCCESS; }
static inline int MPI_Get_processor_name(char* name, int* len)
{
    std::strcpy(name, "serial-shim");
    *len = (int)std::strlen(name);
    return MPI_SUCCESS;
}
static inline int MPI_Bcast(void*, int, MPI_Datatype, int, MPI_Comm) { return MPI_SUCCESS; }
static inline int MPI_Reduce(const void* send, void* recv, int count, MPI_Datatype t, MPI_Op, int, MPI_Comm)
{
    if (send != recv) std::memcpy(recv, send, tdvmc_shim_bytes(count, t));
    return MPI_SUCCESS;
}
static inline int MPI_Allreduce(const void* send, void* recv, int count, MPI_Datatype t, MPI_Op, MPI_Comm)
{
    if (send != recv) std::memcpy(recv, send, tdvmc_shim_bytes(count, t));
    return MPI_SUCCESS;
}
static inline int MPI_Gather(const void* send, int scount, MPI_Datatype st, void* recv, int, MPI_Datatype, int, MPI_Comm)
{
    if (send != recv) std::memcpy(recv, send, tdvmc_shim_bytes(scount, st));
    return MPI_SUCCESS;
}

#endif
