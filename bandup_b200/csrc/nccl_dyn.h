// nccl_dyn.h -- the five NCCL entry points the delta_N gather needs, resolved at run time with
// dlopen so that libbandup_gpu.so carries no link-time dependency on a particular libnccl
// (in a torch process the already-loaded libnccl.so.2 is picked up; stand-alone hosts find the
// system one, or the path in $BANDUP_GPU_NCCL_LIB).  Prototypes follow nccl.h (NCCL 2.x ABI).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace bandup {

struct NcclUniqueId { char internal[128]; };  // NCCL_UNIQUE_ID_BYTES
typedef void* NcclComm;
enum { kNcclSuccess = 0, kNcclInt64 = 4, kNcclFloat64 = 8 };

struct NcclApi {
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
  const char* why = "";
};

const NcclApi& nccl_api();  // loads on first use

}  // namespace bandup
