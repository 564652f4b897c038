// ctx.cuh -- the per-device context behind a handle of the C ABI and the small helpers every
// translation unit of libbandup_gpu.so shares.  Internal; the public surface is
// include/bandup_gpu.h.
#pragma once
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "kernels.cuh"
#include "nccl_dyn.h"

namespace bandup {


constexpr int kDepth = 2;          // k-points in flight (double buffering)
constexpr int kMaxHandles = 64;
constexpr double kEpsDp = 2.220446049250313e-16;  // epsilon(1.0_dp)
constexpr double kMaxTolVecEq = 1e-3;             // max_tol_for_vec_equality

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    // grow geometrically so k-points of slightly different n_pw do not realloc
    size_t want = bytes + bytes / 16 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

struct PinBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 16 + 256;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

struct RhoEntry {  // one stored element of an unfolding-density operator (reference order)
  int32_t fold, iener, m1, m2;
  float re, im;
};

constexpr int kDescRing = 8;
struct DescSlot {  // pinned + device staging of one batch's descriptor block (KptGeom[], ReplItem[], ...)
  char* h = nullptr;
  char* d = nullptr;
  size_t cap = 0;
  size_t used = 0;                // bytes of the block it holds (a repeated batch re-uses it without a copy)
  cudaEvent_t copied = nullptr;   // its host->device copy has finished: h may be rewritten
  cudaEvent_t done = nullptr;     // the kernels that read d have finished: d may be rewritten
  cudaStream_t stream = nullptr;  // the stream of its last use
  bool valid = false;
};

// Host threads that copy pieces of a pageable block into the page-locked staging ring
// (upload_block, stage_pipeline.h).  They live as long as the context and sleep on a condition
// variable between blocks; one job at a time (the ABI is single-caller per handle).
class StagePool;

struct Flight {  // one SC k-point in flight
  bool busy = false;      // submitted, not fetched yet
  bool resident = false;  // device buffers still hold this k-point (for the rho calls)
  int ikpt = -1;
  int n_spinor = 0, n_pw = 0, n_bands = 0, n_fold = 0, n_unique = 0, layout = 0;
  long long stride_bytes = 0;
  bool dp = false;  // uploaded as complex128
  std::vector<int> unique_of;  // pc-kpt (user fold) -> distinct coset
  DevBuf coeffs, gfrac, energies, weights, dN, nsel, norms, slot, selected, workspace;
  PinBuf h_weights, h_dN, h_nsel, h_norms;
  cudaEvent_t uploaded = nullptr, done = nullptr;
  // per-unique-fold results (alias weights/dN/nsel unless two pc-kpts share a coset)
  const double* d_dn_unique = nullptr;
  const int32_t* d_ns_unique = nullptr;
  const double* d_w_unique = nullptr;
  // projected variant
  bool rho_ready = false;
  DevBuf sel_off, csel, pair_count, pair_off, pairs, overflow, trace_out;
  std::vector<long long> h_pair_off;
  std::vector<RhoEntry> rho_entries;
  long long rho_energies = 0;
};

struct Ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t copy = nullptr, compute = nullptr, aux = nullptr;  // aux: device plane-wave lists
  cudaEvent_t aux_done = nullptr;
  bool cells_set = false, grid_set = false;
  double B_sc[9] = {0}, B_inv[9] = {0};  // row-major, row = vector
  int32_t M[9] = {0};                    // row-major int_M
  CosetCommon coset{};
  DescSlot desc_ring[kDescRing];
  int desc_next = 0, desc_last = -1;
  StagePool* stage_pool = nullptr;
  bool dp_coeffs = false;  // kind_cplx_coeffs = dp: every coefficient block is complex128
  int use_pdl = 3;  // programmatic dependent launch: bit 0 = K2 behind K1, bit 1 = K3 behind K2
  int k2_variant = 0;  // 0 auto, 1 shared-memory bins, 2 register accumulators (when <= 4 cosets)
  std::vector<double> grid;
  DevBuf d_grid;
  NcclComm comm = nullptr;  // delta_N gather across ranks (one process per GPU)
  int comm_ranks = 0, comm_rank = -1;
  DevBuf comm_send, comm_recv;
  DevBuf gvec_scratch;  // per-CTA counts/offsets of the device plane-wave enumeration
  // pageable caller memory goes to the device through these page-locked chunks (upload_block)
  static constexpr int kStageBufs = 4;
  PinBuf stage[kStageBufs];
  cudaEvent_t stage_free[kStageBufs] = {};
  int stage_next = 0;
  DevBuf avg_in, avg_out, avg_tmp;  // symmetry averaging (K8/K9) staging
  DevBuf d_proj_dual, d_proj, d_ao_mask, d_proj_ener;  // inputs of the orbital-contribution matrix
  DevBuf d_operator;  // current general operator O [n_bands][n_bands] complex128
  int operator_bands = 0;
  int nener = 0;
  double sigma = kEpsDp;
  Flight flights[kDepth];
  int next_flight = 0;
  int chunk_vecs = 0, ctas_per_sm = 0, n_stages = 0;
  unsigned profiling = 0;  // bit k: launches of kernel id k are bracketed by events
  bool perform_unfold = true;  // args%perform_unfold; false == -dont_unfold
  std::vector<cudaEvent_t> event_pool;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending[BANDUP_GPU_K_COUNT];
  double total_ms[BANDUP_GPU_K_COUNT] = {0};
  int64_t count[BANDUP_GPU_K_COUNT] = {0};
  int64_t launches = 0;
  std::string err = "";
};


extern std::mutex g_mutex;
extern Ctx* g_ctx[kMaxHandles];
extern std::string g_global_err;

Ctx* get(int h);
int fail(Ctx* c, int code, const std::string& msg);
int cuda_fail(Ctx* c, cudaError_t e, const char* where);

#define CU(call)                                         \
  do {                                                   \
    cudaError_t e_ = (call);                             \
    if (e_ != cudaSuccess) return cuda_fail(c, e_, #call); \
  } while (0)

cudaEvent_t take_event(Ctx* c);
void resolve_pending(Ctx* c);
inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

struct KernelScope {  // brackets one launch with events when profiling is on
  Ctx* c;
  int which;
  cudaStream_t s;
  cudaEvent_t a = nullptr, b = nullptr;
  KernelScope(Ctx* c_, int which_, cudaStream_t s_) : c(c_), which(which_), s(s_) {
    c->launches++;
    if (c->profiling >> which & 1u) {
      a = take_event(c);
      b = take_event(c);
      cudaEventRecord(a, s);
    }
  }
  ~KernelScope() {
    if (a) {
      cudaEventRecord(b, s);
      c->pending[which].push_back({a, b});
    }
  }
};

// Host -> device copy on the copy stream; pageable sources go through the pinned staging ring.
int upload_block(Ctx* c, void* dst, const void* src, size_t bytes);
void stage_pool_destroy(Ctx* c);

// The flight that still holds SC-kpt `ikpt` after it has been fetched (rho, Pauli vectors, the
// spectral function work on it); error when it is still in flight or its buffers were reused.
int need_resident(Ctx* c, int ikpt, Flight** out);

}  // namespace bandup
