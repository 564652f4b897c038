// coset_classify.cu -- K1: integer coset classification of the SC plane waves.
//
// Replaces the floating-point sweep of select_coeffs_to_calc_spectral_weights
// (reference band_unfolding_routines_mod.f90:586-600), which calls vec_in_latt
// (crystals_mod.f90:204-255) once per plane wave and per folding pc-kpt.  Here
// every plane wave's Miller triple is mapped ONCE to its coset key and compared
// with the keys of all pc-kpts folding onto the current SC-kpt; the result is a
// one-byte slot id per plane wave that K2 consumes (written twice: slot[] and the
// one-element-shifted slot_shift[], see kernels.cuh).  Pure integer arithmetic:
// bit-exact by construction.  HBM-trivial: 12 B read + 2 B written per wave.
#include "kernels.cuh"

namespace bandup {

__device__ __forceinline__ int pos_mod(long long v, long long D) {
  long long r = v % D;
  return (int)(r < 0 ? r + D : r);
}

__global__ void __launch_bounds__(256)
k1_coset_classify(const int32_t* __restrict__ g_frac, int n_pw, CosetParams cp,
                  uint8_t* __restrict__ slot, uint8_t* __restrict__ slot_shift,
                  int32_t* __restrict__ n_selected) {
  __shared__ int s_count[kMaxFolds];
  for (int f = threadIdx.x; f < cp.n_fold; f += blockDim.x) s_count[f] = 0;
  __syncthreads();

  const int stride = gridDim.x * blockDim.x;
  for (int pw = blockIdx.x * blockDim.x + threadIdx.x; pw < n_pw; pw += stride) {
    const long long g1 = g_frac[3 * pw], g2 = g_frac[3 * pw + 1], g3 = g_frac[3 * pw + 2];
    const int k0 = pos_mod(g1 * cp.adjT[0] + g2 * cp.adjT[3] + g3 * cp.adjT[6], cp.D);
    const int k1 = pos_mod(g1 * cp.adjT[1] + g2 * cp.adjT[4] + g3 * cp.adjT[7], cp.D);
    const int k2 = pos_mod(g1 * cp.adjT[2] + g2 * cp.adjT[5] + g3 * cp.adjT[8], cp.D);
    int s = kNoFold;
    for (int f = cp.n_fold - 1; f >= 0; --f)  // descending so the FIRST match wins
      if (cp.fold_key[f][0] == k0 && cp.fold_key[f][1] == k1 && cp.fold_key[f][2] == k2) s = f;
    slot[pw] = (uint8_t)s;
    if (pw > 0) slot_shift[pw - 1] = (uint8_t)s;
    if (pw == n_pw - 1) slot_shift[pw] = kNoFold;
    if (s != kNoFold) atomicAdd(&s_count[s], 1);
  }
  __syncthreads();
  for (int f = threadIdx.x; f < cp.n_fold; f += blockDim.x)
    if (s_count[f]) atomicAdd(&n_selected[f], s_count[f]);
}

void launch_coset_classify(const int32_t* d_g_frac, int n_pw, const CosetParams& cp,
                           uint8_t* d_slot, uint8_t* d_slot_shift, int32_t* d_n_selected,
                           cudaStream_t s) {
  cudaMemsetAsync(d_n_selected, 0, sizeof(int32_t) * cp.n_fold, s);
  if (n_pw <= 0) return;
  int blocks = (n_pw + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  k1_coset_classify<<<blocks, 256, 0, s>>>(d_g_frac, n_pw, cp, d_slot, d_slot_shift, d_n_selected);
}

// One CTA per fold walks the slot table in order and writes the 1-based indices
// of its plane waves: the same list the reference builds with append() under
// `omp critical` (band_unfolding_routines_mod.f90:596-598), here always ascending.
__global__ void __launch_bounds__(1024)
k1_selected_lists(const uint8_t* __restrict__ slot, int n_pw, int n_fold,
                  const int32_t* __restrict__ n_selected, int32_t* __restrict__ selected) {
  const int f = blockIdx.x;
  __shared__ int s_warp[32];
  __shared__ int s_base;
  if (threadIdx.x == 0) {
    int off = 0;
    for (int i = 0; i < f; ++i) off += n_selected[i];
    s_base = off;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int start = 0; start < n_pw; start += blockDim.x) {
    const int pw = start + threadIdx.x;
    const bool hit = pw < n_pw && slot[pw] == f;
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < nwarp; ++w) {
      const int c = s_warp[w];
      if (w < warp) before += c;
      total += c;
    }
    if (hit) selected[s_base + before + __popc(bal & ((1u << lane) - 1u))] = pw + 1;
    __syncthreads();
    if (threadIdx.x == 0) s_base += total;
    __syncthreads();
  }
}

void launch_selected_lists(const uint8_t* d_slot, int n_pw, int n_fold,
                           const int32_t* d_n_selected, int32_t* d_selected,
                           cudaStream_t s) {
  if (n_fold <= 0 || n_pw <= 0) return;
  k1_selected_lists<<<n_fold, 1024, 0, s>>>(d_slot, n_pw, n_fold, d_n_selected, d_selected);
}

}  // namespace bandup
