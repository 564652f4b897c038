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

// grid (CTAs per k-point <= kK1MaxBlocks, n_kpts).  Every CTA leaves its per-fold plane-wave
// counts in block_counts[blockIdx.x][*]; K3 adds the rows up (fixed order, no atomics on
// global memory, nothing to zero beforehand).
__global__ void __launch_bounds__(256)
k1_coset_classify(const KptGeom* __restrict__ geoms, CosetCommon cc) {
  // K2 is launched with programmatic stream serialisation: its producer may start prefetching
  // coefficients now; its consumers wait for this grid to finish before they read slot[]
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const KptGeom& g = geoms[blockIdx.y];
  __shared__ int s_count[kMaxFolds];
  __shared__ int s_key[kMaxFolds][3];
  const int n_fold = g.n_fold;
  for (int f = threadIdx.x; f < n_fold; f += blockDim.x) {
    s_count[f] = 0;
    s_key[f][0] = g.fold_key[f][0];
    s_key[f][1] = g.fold_key[f][1];
    s_key[f][2] = g.fold_key[f][2];
  }
  __syncthreads();

  const int n_pw = g.n_pw;
  const int32_t* __restrict__ g_frac = g.g_frac;
  uint8_t* __restrict__ slot = g.slot;
  uint8_t* __restrict__ slot_shift = g.slot_shift;
  const int stride = gridDim.x * blockDim.x;
  // |det M| < 2^15 (always, in practice): everything fits 32-bit unsigned arithmetic once the
  // Miller indices and the adjugate are reduced mod D (3 * (D-1)^2 < 2^32); same residues.
  const bool small = cc.D < 32768;
  const int D32 = (int)cc.D;
  const unsigned* a32 = cc.adj32;
  for (int pw = blockIdx.x * blockDim.x + threadIdx.x; pw < n_pw; pw += stride) {
    int k0, k1, k2;
    if (small) {
      int r1 = g_frac[3 * pw] % D32, r2 = g_frac[3 * pw + 1] % D32, r3 = g_frac[3 * pw + 2] % D32;
      const unsigned u1 = (unsigned)(r1 < 0 ? r1 + D32 : r1), u2 = (unsigned)(r2 < 0 ? r2 + D32 : r2),
                     u3 = (unsigned)(r3 < 0 ? r3 + D32 : r3);
      k0 = (int)((u1 * a32[0] + u2 * a32[3] + u3 * a32[6]) % (unsigned)D32);
      k1 = (int)((u1 * a32[1] + u2 * a32[4] + u3 * a32[7]) % (unsigned)D32);
      k2 = (int)((u1 * a32[2] + u2 * a32[5] + u3 * a32[8]) % (unsigned)D32);
    } else {
      const long long g1 = g_frac[3 * pw], g2 = g_frac[3 * pw + 1], g3 = g_frac[3 * pw + 2];
      k0 = pos_mod(g1 * cc.adjT[0] + g2 * cc.adjT[3] + g3 * cc.adjT[6], cc.D);
      k1 = pos_mod(g1 * cc.adjT[1] + g2 * cc.adjT[4] + g3 * cc.adjT[7], cc.D);
      k2 = pos_mod(g1 * cc.adjT[2] + g2 * cc.adjT[5] + g3 * cc.adjT[8], cc.D);
    }
    int q = kNoFold;
    for (int f = n_fold - 1; f >= 0; --f)  // folds are distinct cosets: at most one matches
      if (s_key[f][0] == k0 && s_key[f][1] == k1 && s_key[f][2] == k2) q = f;
    slot[pw] = (uint8_t)q;
    if (pw > 0) slot_shift[pw - 1] = (uint8_t)q;
    if (pw == n_pw - 1) slot_shift[pw] = kNoFold;
    if (q != kNoFold) atomicAdd(&s_count[q], 1);  // shared-memory integer count: order-free
  }
  __syncthreads();
  for (int f = threadIdx.x; f < n_fold; f += blockDim.x)
    g.block_counts[blockIdx.x * n_fold + f] = s_count[f];
}

// About four plane waves per thread: a 20k-wave k-point takes 20 CTAs, not 128 mostly idle ones
// (a batch of 99 such k-points used to launch 12672 CTAs: 25 us for 2 M plane waves).
int coset_classify_blocks(int max_n_pw) {
  const int b = (max_n_pw + 256 * 4 - 1) / (256 * 4);
  return b < 1 ? 1 : b > kK1MaxBlocks ? kK1MaxBlocks : b;
}

void launch_coset_classify(const KptGeom* d_geoms, int n_kpts, int max_n_pw, const CosetCommon& cc,
                           cudaStream_t s) {
  if (n_kpts <= 0 || max_n_pw <= 0) return;
  // KptGeom::k1_blocks rows of counts per k-point (K3 adds them up): the same number for every
  // k-point of the batch, sized for its largest one
  dim3 grid((unsigned)coset_classify_blocks(max_n_pw), (unsigned)n_kpts);
  k1_coset_classify<<<grid, 256, 0, s>>>(d_geoms, cc);
}

// One CTA per fold walks the slot table in order and writes the 1-based indices
// of its plane waves: the same list the reference builds with append() under
// `omp critical` (band_unfolding_routines_mod.f90:596-598), here always ascending.
// Fold f lives in the table of pass f / kMaxFolds under the id f % kMaxFolds.
__global__ void __launch_bounds__(1024)
k1_selected_lists(const uint8_t* __restrict__ slot_tables, size_t slot_pitch, int n_pw, int n_fold,
                  const int32_t* __restrict__ n_selected, int32_t* __restrict__ selected) {
  const int f_global = blockIdx.x;
  const uint8_t* __restrict__ slot = slot_tables + slot_pitch * (size_t)(f_global / kMaxFolds);
  const int f = f_global % kMaxFolds;
  __shared__ int s_warp[32];
  __shared__ int s_base;
  if (threadIdx.x == 0) {
    int off = 0;
    for (int i = 0; i < f_global; ++i) off += n_selected[i];
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

// One CTA per (fold, k-point): the rows of a pc-kpt that shares its coset with an earlier one are
// copies of that one's (replaces 3 * n_fold small device-to-device copies per k-point).
__global__ void __launch_bounds__(256)
k1_replicate_rows(const ReplItem* __restrict__ items) {
  const ReplItem it = items[blockIdx.y];
  const int f = blockIdx.x;
  if (f >= it.n_fold) return;
  const int u = (reinterpret_cast<const int*>(items + gridDim.y) + it.uo_off)[f];
  const double* __restrict__ ws = it.w_src + (long long)u * it.n_bands;
  double* __restrict__ wd = it.w_dst + (long long)f * it.n_bands;
  for (int i = threadIdx.x; i < it.n_bands; i += blockDim.x) wd[i] = ws[i];
  const double* __restrict__ ds = it.dn_src + (long long)u * it.nener;
  double* __restrict__ dd = it.dn_dst + (long long)f * it.nener;
  for (int i = threadIdx.x; i < it.nener; i += blockDim.x) dd[i] = ds[i];
  if (threadIdx.x == 0) it.ns_dst[f] = it.ns_src[u];
}

void launch_replicate_rows(const ReplItem* d_items, int n_items, int max_n_fold, cudaStream_t s) {
  if (n_items <= 0 || max_n_fold <= 0) return;
  k1_replicate_rows<<<dim3((unsigned)max_n_fold, (unsigned)n_items), 256, 0, s>>>(d_items);
}

void launch_selected_lists(const uint8_t* d_slot, size_t slot_pitch, int n_pw, int n_fold,
                           const int32_t* d_n_selected, int32_t* d_selected,
                           cudaStream_t s) {
  if (n_fold <= 0 || n_pw <= 0) return;
  k1_selected_lists<<<n_fold, 1024, 0, s>>>(d_slot, slot_pitch, n_pw, n_fold, d_n_selected, d_selected);
}

}  // namespace bandup
