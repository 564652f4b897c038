// comm.cu -- the one collective of the path: every rank's disjoint delta_N rows gathered with
// ncclAllGather over NVLink (band_unfolding_routines_mod.f90:317-324: every pc-kpt folds onto
// exactly one SC-kpt, so rows of different ranks never overlap and nothing is reduced).
#include <dlfcn.h>

#include <cstdlib>

#include "nccl_dyn.h"

namespace bandup {

const NcclApi& nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  void* h = nullptr;
  const char* env = std::getenv("BANDUP_GPU_NCCL_LIB");
  if (env && *env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // already in the process (torch)
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    api.why = "libnccl.so.2 not found (set BANDUP_GPU_NCCL_LIB)";
    return api;
  }
  api.GetUniqueId = reinterpret_cast<int (*)(NcclUniqueId*)>(dlsym(h, "ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<int (*)(NcclComm*, int, NcclUniqueId, int)>(dlsym(h, "ncclCommInitRank"));
  api.CommDestroy = reinterpret_cast<int (*)(NcclComm)>(dlsym(h, "ncclCommDestroy"));
  api.AllGather = reinterpret_cast<int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t)>(
      dlsym(h, "ncclAllGather"));
  api.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(h, "ncclGetErrorString"));
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.GetErrorString;
  if (!api.ok) api.why = "libnccl is missing an expected symbol";
  return api;
}

}  // namespace bandup
