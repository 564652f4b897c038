// comm.cu -- the one collective of the path: every rank's disjoint delta_N rows gathered with
// ncclAllGather over NVLink (band_unfolding_routines_mod.f90:317-324: every pc-kpt folds onto
// exactly one SC-kpt, so rows of different ranks never overlap and nothing is reduced).
#include <dlfcn.h>

#include <cstdlib>

#include <algorithm>
#include <cstring>

#include "ctx.cuh"

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

using namespace bandup;

extern "C" {

// ---- the delta_N gather (one process per GPU) ---------------------------------

int bandup_gpu_comm_unique_id(void* id) {
  const NcclApi& n = nccl_api();
  if (!n.ok) return fail(nullptr, BANDUP_GPU_ERR_NCCL, n.why);
  if (!id) return fail(nullptr, BANDUP_GPU_ERR_INVALID_ARG, "id is NULL");
  NcclUniqueId u;
  const int rc = n.GetUniqueId(&u);
  if (rc != kNcclSuccess) return fail(nullptr, BANDUP_GPU_ERR_NCCL, std::string("ncclGetUniqueId: ") + n.GetErrorString(rc));
  std::memcpy(id, u.internal, sizeof(u.internal));
  return BANDUP_GPU_OK;
}

int bandup_gpu_comm_init(int handle, int n_ranks, int rank, const void* id) {
  Ctx* c = get(handle);
  if (!c) return fail(nullptr, BANDUP_GPU_ERR_BAD_HANDLE, "bad handle");
  const NcclApi& n = nccl_api();
  if (!n.ok) return fail(c, BANDUP_GPU_ERR_NCCL, n.why);
  if (!id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(c, BANDUP_GPU_ERR_INVALID_ARG, "bad rank / id");
  if (c->comm) return fail(c, BANDUP_GPU_ERR_INVALID_ARG, "communicator already initialised");
  cudaSetDevice(c->device);
  NcclUniqueId u;
  std::memcpy(u.internal, id, sizeof(u.internal));
  const int rc = n.CommInitRank(&c->comm, n_ranks, u, rank);
  if (rc != kNcclSuccess) {
    c->comm = nullptr;
    return fail(c, BANDUP_GPU_ERR_NCCL, std::string("ncclCommInitRank: ") + n.GetErrorString(rc));
  }
  c->comm_ranks = n_ranks;
  c->comm_rank = rank;
  return BANDUP_GPU_OK;
}

int bandup_gpu_comm_finalize(int handle) {
  Ctx* c = get(handle);
  if (!c) return fail(nullptr, BANDUP_GPU_ERR_BAD_HANDLE, "bad handle");
  if (c->comm) {
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->compute);
    nccl_api().CommDestroy(c->comm);
    c->comm = nullptr;
  }
  c->comm_ranks = 0;
  c->comm_rank = -1;
  return BANDUP_GPU_OK;
}

int bandup_gpu_gather_dN(int handle, const double* dN_local, const int64_t* row_ids_local, int64_t n_local,
                         double* dN_all, int64_t* row_ids_all, int64_t capacity_rows, int64_t* n_total) {
  Ctx* c = get(handle);
  if (!c) return fail(nullptr, BANDUP_GPU_ERR_BAD_HANDLE, "bad handle");
  if (!c->grid_set) return fail(c, BANDUP_GPU_ERR_NOT_CONFIGURED, "call bandup_gpu_set_grid first");
  return bandup_gpu_gather_rows(handle, c->nener, dN_local, row_ids_local, n_local, dN_all, row_ids_all,
                                capacity_rows, n_total);
}

int bandup_gpu_allgather_device(int handle, void* stream, const double* d_send, double* d_recv, int64_t count) {
  Ctx* c = get(handle);
  if (!c) return fail(nullptr, BANDUP_GPU_ERR_BAD_HANDLE, "bad handle");
  if (!c->comm) return fail(c, BANDUP_GPU_ERR_NOT_CONFIGURED, "call bandup_gpu_comm_init first");
  if (!d_send || !d_recv || count < 1) return fail(c, BANDUP_GPU_ERR_INVALID_ARG, "bad allgather arguments");
  const NcclApi& n = nccl_api();
  cudaSetDevice(c->device);
  const int rc = n.AllGather(d_send, d_recv, (size_t)count, kNcclFloat64, c->comm, static_cast<cudaStream_t>(stream));
  if (rc != kNcclSuccess) return fail(c, BANDUP_GPU_ERR_NCCL, std::string("ncclAllGather: ") + n.GetErrorString(rc));
  return BANDUP_GPU_OK;
}

int bandup_gpu_gather_rows(int handle, int row_len, const double* dN_local, const int64_t* row_ids_local,
                           int64_t n_local, double* dN_all, int64_t* row_ids_all, int64_t capacity_rows,
                           int64_t* n_total) {
  Ctx* c = get(handle);
  if (!c) return fail(nullptr, BANDUP_GPU_ERR_BAD_HANDLE, "bad handle");
  if (!c->comm) return fail(c, BANDUP_GPU_ERR_NOT_CONFIGURED, "call bandup_gpu_comm_init first");
  if (row_len < 1 || n_local < 0 || (n_local > 0 && (!dN_local || !row_ids_local)) || !n_total)
    return fail(c, BANDUP_GPU_ERR_INVALID_ARG, "bad gather arguments");
  const NcclApi& n = nccl_api();
  const int R = c->comm_ranks, nener = row_len;
  cudaSetDevice(c->device);
  cudaStream_t s = c->compute;
  // 1. how many rows everyone has
  CU(c->comm_send.reserve(sizeof(long long)));
  CU(c->comm_recv.reserve(sizeof(long long) * (size_t)R));
  const long long mine = n_local;
  CU(cudaMemcpyAsync(c->comm_send.p, &mine, sizeof(mine), cudaMemcpyHostToDevice, s));
  int rc = n.AllGather(c->comm_send.p, c->comm_recv.p, 1, kNcclInt64, c->comm, s);
  if (rc != kNcclSuccess) return fail(c, BANDUP_GPU_ERR_NCCL, std::string("ncclAllGather: ") + n.GetErrorString(rc));
  std::vector<long long> counts((size_t)R);
  CU(cudaMemcpyAsync(counts.data(), c->comm_recv.p, sizeof(long long) * (size_t)R, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  long long n_max = 1, total = 0;
  for (long long v : counts) {
    if (v > n_max) n_max = v;
    total += v;
  }
  *n_total = total;
  // count query: every rank sees the same NULL-ness of the outputs (same call on every rank), so
  // all of them stop here together, before any payload moves
  if (!dN_all || !row_ids_all) return BANDUP_GPU_OK;
  if (capacity_rows < total) return fail(c, BANDUP_GPU_ERR_CAPACITY, "capacity_rows too small: need " + std::to_string(total));
  // 2. one padded payload per rank: [row id (exact in a double below 2^53), nener values] per row
  const size_t row = (size_t)nener + 1, per_rank = row * (size_t)n_max;
  std::vector<double> send(per_rank, 0.0);
  for (long long i = 0; i < n_local; ++i) {
    send[(size_t)i * row] = (double)row_ids_local[i];
    std::memcpy(&send[(size_t)i * row + 1], dN_local + (size_t)i * nener, sizeof(double) * (size_t)nener);
  }
  CU(c->comm_send.reserve(sizeof(double) * per_rank));
  CU(c->comm_recv.reserve(sizeof(double) * per_rank * (size_t)R));
  CU(cudaMemcpyAsync(c->comm_send.p, send.data(), sizeof(double) * per_rank, cudaMemcpyHostToDevice, s));
  rc = n.AllGather(c->comm_send.p, c->comm_recv.p, per_rank, kNcclFloat64, c->comm, s);
  if (rc != kNcclSuccess) return fail(c, BANDUP_GPU_ERR_NCCL, std::string("ncclAllGather: ") + n.GetErrorString(rc));
  std::vector<double> recv(per_rank * (size_t)R);
  CU(cudaMemcpyAsync(recv.data(), c->comm_recv.p, sizeof(double) * recv.size(), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  // 3. rows in ascending global row id; a duplicate id means a pc-kpt was unfolded twice
  std::vector<std::pair<long long, const double*>> rows;
  rows.reserve((size_t)total);
  for (int r = 0; r < R; ++r)
    for (long long i = 0; i < counts[(size_t)r]; ++i) {
      const double* p = &recv[(size_t)r * per_rank + (size_t)i * row];
      rows.push_back({(long long)p[0], p + 1});
    }
  std::stable_sort(rows.begin(), rows.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
  for (size_t i = 0; i < rows.size(); ++i) {
    if (i > 0 && rows[i].first == rows[i - 1].first)
      return fail(c, BANDUP_GPU_ERR_INVALID_ARG,
                  "pc-kpt row " + std::to_string(rows[i].first) + " was produced by more than one rank");
    row_ids_all[i] = rows[i].first;
    std::memcpy(dN_all + i * (size_t)nener, rows[i].second, sizeof(double) * (size_t)nener);
  }
  return BANDUP_GPU_OK;
}

}  // extern "C"
