// nccl_dyn.h -- NCCL resolved at run time (dlopen), so that a single-GPU host never needs it and a host that already
// loaded an NCCL (e.g. the one bundled with PyTorch) shares that copy instead of pulling a second one into the process.
// Only the handful of entry points the side-halo exchange needs (replacing MPI_ISEND/IRECV/WaitAll of mpi/mpi.f90:244-445
// and the MPI_ALLREDUCE of calctimestep.f90:221).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stddef.h>

namespace fx {

typedef struct ncclComm* ncclComm_t;
struct ncclUniqueIdDyn {
  char internal[128];
};
enum { NCCL_SUCCESS = 0 };
enum { NCCL_DOUBLE = 8 };  // ncclFloat64
enum { NCCL_MIN = 3 };     // ncclMin

struct Nccl {
  void* lib = nullptr;
  int (*GetUniqueId)(ncclUniqueIdDyn*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueIdDyn, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;

  // returns NULL on success, else a static error text
  const char* load() {
    if (lib) return nullptr;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) return "cannot dlopen libnccl.so.2";
#define FX_SYM(field, name)                                  \
  *(void**)(&field) = dlsym(lib, name);                      \
  if (!field) return "libnccl is missing the symbol " name;
    FX_SYM(GetUniqueId, "ncclGetUniqueId")
    FX_SYM(CommInitRank, "ncclCommInitRank")
    FX_SYM(CommDestroy, "ncclCommDestroy")
    FX_SYM(Send, "ncclSend")
    FX_SYM(Recv, "ncclRecv")
    FX_SYM(AllReduce, "ncclAllReduce")
    FX_SYM(GroupStart, "ncclGroupStart")
    FX_SYM(GroupEnd, "ncclGroupEnd")
    FX_SYM(GetErrorString, "ncclGetErrorString")
#undef FX_SYM
    return nullptr;
  }
};

inline Nccl& nccl() {
  static Nccl n;
  return n;
}

}  // namespace fx
