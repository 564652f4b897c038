// symm_average.cu -- K8/K9: symmetry averaging over the needed directions of a pc-BZ direction,
// the post-loop step of the unfolding path (get_delta_Ns_for_output,
// band_unfolding_routines_mod.f90:789-1043).
//
//   K8   dense columns (delta_N, spin_proj_perp, spin_proj_para, sigma): out = sum_d w_d x_d in dp,
//        accumulated in direction order with separate multiply and add, so the result is bit-equal
//        to the reference's `avrgd = avrgd + weight*x` loop.  One pass over n_dirs * n_values
//        doubles, HBM-bound, grid-stride.
//   K9   sparse unfolding-density operators (:857-973).  The reference looks every (m1,m2) of the
//        nb x nb matrix up in two unsorted lists for every (direction, energy): O(nb^2 * nnz).
//        Here every direction's entries arrive sorted by key = (pckpt, iener, m1, m2) -- the
//        reference's own append order -- so "is this element already in the averaged operator"
//        is a binary search in the earlier directions, the averaged value is a walk through the
//        later ones, and the output slot follows from one exclusive scan: no sort, no atomics,
//        and the same element order as the reference's appends.
#include "kernels.cuh"

namespace bandup {

namespace {

__global__ void __launch_bounds__(256)
k8_average(const double* __restrict__ in, const double* __restrict__ w, int n_dirs, long long n_values,
           double* __restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_values; r += stride) {
    double acc = 0.0;
    for (int d = 0; d < n_dirs; ++d) acc = __dadd_rn(acc, __dmul_rn(__ldg(w + d), __ldg(in + d * n_values + r)));
    out[r] = acc;
  }
}

__device__ __forceinline__ int dir_of(const long long* __restrict__ dir_off, int n_dirs, long long i) {
  int lo = 0, hi = n_dirs;  // last d with dir_off[d] <= i
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (dir_off[mid] <= i) lo = mid; else hi = mid;
  }
  return lo;
}

// first index in [lo, hi) whose key is >= k
__device__ __forceinline__ long long lower_bound(const long long* __restrict__ key, long long lo, long long hi,
                                                 long long k) {
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (key[mid] < k) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// key[i] = group * nb^2 + pair; err: 1 = index out of range, 2 = a direction is not sorted/unique
__global__ void __launch_bounds__(256)
k9_keys(const int32_t* __restrict__ pck, const int32_t* __restrict__ ien, const int32_t* __restrict__ m1,
        const int32_t* __restrict__ m2, long long n, int n_pckpts, int nener, int n_bands,
        long long* __restrict__ key, int* __restrict__ err) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int k = pck[i], e = ien[i] - 1, a = m1[i] - 1, b = m2[i] - 1;
  if (k < 0 || k >= n_pckpts || e < 0 || e >= nener || a < 0 || a >= n_bands || b < 0 || b >= n_bands) {
    atomicOr(err, 1);
    key[i] = -1;
    return;
  }
  key[i] = ((long long)k * nener + e) * ((long long)n_bands * n_bands) + (long long)a * n_bands + b;
}

// flag[i] = 1 when entry i creates an element of the averaged operator: its energy survives the
// averaged-delta_N cut and no earlier direction stores the same (pckpt, iener, m1, m2)
__global__ void __launch_bounds__(256)
k9_first(const long long* __restrict__ key, long long n, const long long* __restrict__ dir_off, int n_dirs,
         long long pairs_per_group, const double* __restrict__ avg_dN, double min_dN, int* __restrict__ flag,
         int* __restrict__ err) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long k = key[i];
  const int d = dir_of(dir_off, n_dirs, i);
  if (i > dir_off[d] && !(key[i - 1] < k)) atomicOr(err, 2);
  int first = 0;
  if (k >= 0 && !(avg_dN[k / pairs_per_group] < min_dN)) {  // `if(symm_avg_dNs(iener) < 1E-3) cycle`
    first = 1;
    for (int e = 0; e < d && first; ++e) {
      const long long lo = dir_off[e], hi = dir_off[e + 1];
      const long long j = lower_bound(key, lo, hi, k);
      if (j < hi && key[j] == k) first = 0;
    }
  }
  flag[i] = first;
}

// one thread per created element: its slot in the reference's order and its value
__global__ void __launch_bounds__(256)
k9_merge(const long long* __restrict__ key, long long n, const long long* __restrict__ dir_off, int n_dirs,
         long long pairs_per_group, int nener, int n_bands, const double* __restrict__ w,
         const float* __restrict__ re, const float* __restrict__ im, const int* __restrict__ flag,
         const long long* __restrict__ scan, int32_t* __restrict__ o_pck, int32_t* __restrict__ o_ien,
         int32_t* __restrict__ o_m1, int32_t* __restrict__ o_m2, float* __restrict__ o_re,
         float* __restrict__ o_im) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !flag[i]) return;
  const long long k = key[i];
  const long long group = k / pairs_per_group, pair = k - group * pairs_per_group;
  const long long g_lo = group * pairs_per_group, g_hi = g_lo + pairs_per_group;
  const int d = dir_of(dir_off, n_dirs, i);
  // slot: created elements of earlier groups (all directions), then this group's elements of the
  // earlier directions, then the ones before i in its own direction
  long long slot = scan[i] - scan[dir_off[d]];
  // value: cmplx(weight*rho, kind=sp), then sp(dp(avg) + weight*rho) for every later direction
  float ar = __double2float_rn(__dmul_rn(w[d], (double)re[i]));
  float ai = __double2float_rn(__dmul_rn(w[d], (double)im[i]));
  for (int e = 0; e < n_dirs; ++e) {
    if (e == d) continue;
    const long long lo = dir_off[e], hi = dir_off[e + 1];
    if (e < d) {
      slot += scan[lower_bound(key, lo, hi, g_hi)] - scan[lo];
    } else {
      slot += scan[lower_bound(key, lo, hi, g_lo)] - scan[lo];
      const long long j = lower_bound(key, lo, hi, k);
      if (j < hi && key[j] == k) {
        ar = __double2float_rn(__dadd_rn((double)ar, __dmul_rn(w[e], (double)re[j])));
        ai = __double2float_rn(__dadd_rn((double)ai, __dmul_rn(w[e], (double)im[j])));
      }
    }
  }
  o_pck[slot] = (int32_t)(group / nener);
  o_ien[slot] = (int32_t)(group % nener) + 1;
  o_m1[slot] = (int32_t)(pair / n_bands) + 1;
  o_m2[slot] = (int32_t)(pair % n_bands) + 1;
  o_re[slot] = ar;
  o_im[slot] = ai;
}

}  // namespace

void launch_symm_average(const double* d_in, const double* d_w, int n_dirs, long long n_values, double* d_out,
                         int sm_count, cudaStream_t s) {
  if (n_values <= 0) return;
  long long blocks = (n_values + 255) / 256;
  const long long cap = (long long)sm_count * 8;
  if (blocks > cap) blocks = cap;
  k8_average<<<(unsigned)blocks, 256, 0, s>>>(d_in, d_w, n_dirs, n_values, d_out);
}

void launch_symm_average_rho_plan(const int32_t* d_pck, const int32_t* d_ien, const int32_t* d_m1,
                                  const int32_t* d_m2, long long n, const long long* d_dir_off, int n_dirs,
                                  int n_pckpts, int nener, int n_bands, const double* d_avg_dN, double min_dN,
                                  long long* d_key, int* d_flag, long long* d_scan, int* d_err, cudaStream_t s) {
  const unsigned blocks = (unsigned)((n + 255) / 256);
  k9_keys<<<blocks, 256, 0, s>>>(d_pck, d_ien, d_m1, d_m2, n, n_pckpts, nener, n_bands, d_key, d_err);
  k9_first<<<blocks, 256, 0, s>>>(d_key, n, d_dir_off, n_dirs, (long long)n_bands * n_bands, d_avg_dN, min_dN,
                                  d_flag, d_err);
  launch_scan(d_flag, (int)n, d_scan, s);
}

void launch_symm_average_rho_merge(const long long* d_key, long long n, const long long* d_dir_off, int n_dirs,
                                   int nener, int n_bands, const double* d_w, const float* d_re,
                                   const float* d_im, const int* d_flag, const long long* d_scan,
                                   int32_t* o_pck, int32_t* o_ien, int32_t* o_m1, int32_t* o_m2, float* o_re,
                                   float* o_im, cudaStream_t s) {
  const unsigned blocks = (unsigned)((n + 255) / 256);
  k9_merge<<<blocks, 256, 0, s>>>(d_key, n, d_dir_off, n_dirs, (long long)n_bands * n_bands, nener, n_bands,
                                  d_w, d_re, d_im, d_flag, d_scan, o_pck, o_ien, o_m1, o_m2, o_re, o_im);
}

}  // namespace bandup
