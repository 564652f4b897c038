// delta_n.cu -- K3: energy binning of the spectral weights into delta_N(k, E).
//
// Replaces calc_delta_N_pckpt (reference band_unfolding_routines_mod.f90:719-757)
// as driven by get_delta_Ns_for_EBS (:760-786):
//   delta_N(E_i) = sum_m P_m * lambda(E_i; E_m, delta_e, sigma)
//   lambda = 1/2 [erf((E_i + delta_e/2 - E_m)/(sqrt2 sigma))
//                - erf((E_i - delta_e/2 - E_m)/(sqrt2 sigma))]      (:701-716, math_mod.f90:181-193)
// The default sigma = epsilon(1.0_dp) turns lambda into a box with the value
// 1/2 exactly on a bin edge; those edge decisions are double-precision
// comparisons of `E_i -/+ 0.5*delta_e` against E_m, so the kernel evaluates the
// SAME expressions with explicitly rounded (non-fused) operations instead of
// computing a bin index.  The reference evaluates two erf per (E_i, band) pair;
// here a pair is skipped only when both arguments lie beyond +/-6 on the same
// side, where both erf are exactly +/-1 in double precision and the term is
// exactly zero, so skipping changes nothing.
//
// One warp per (energy bin, fold); lanes stride over the bands and a shuffle
// tree combines them (fixed order: deterministic).  Work is tiny next to K2.
#include "kernels.cuh"

namespace bandup {

namespace {

__device__ __forceinline__ double lambda_box_erf(double pc_ener, double sc_ener,
                                                 double half_de, double sqrt2_sigma) {
  const double lower = __dsub_rn(pc_ener, half_de);
  const double upper = __dadd_rn(pc_ener, half_de);
  const double x1 = __ddiv_rn(__dsub_rn(lower, sc_ener), sqrt2_sigma);
  const double x2 = __ddiv_rn(__dsub_rn(upper, sc_ener), sqrt2_sigma);
  if (x1 > 6.0 || x2 < -6.0) return 0.0;  // erf(x1) == erf(x2) exactly
  return __dmul_rn(0.5, __dsub_rn(erf(x2), erf(x1)));
}

__global__ void __launch_bounds__(128)
k3_delta_n(const double* __restrict__ grid, int nener, double sigma,
           const double* __restrict__ band_energies, const double* __restrict__ weights,
           int n_bands, int n_fold, double* __restrict__ dN) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (warp_global >= (long long)nener * n_fold) return;
  const int f = (int)(warp_global / nener);
  const int ie = (int)(warp_global - (long long)f * nener);
  // delta_e = energies(2) - energies(1)  (:735)
  const double delta_e = __dsub_rn(grid[1], grid[0]);
  const double half_de = __dmul_rn(0.5, delta_e);
  const double sqrt2_sigma = __dmul_rn(sqrt(2.0), sigma);
  const double E = grid[ie];
  const double* w = weights + (long long)f * n_bands;
  double acc = 0.0;
  for (int m = lane; m < n_bands; m += 32) {
    const double lam = lambda_box_erf(E, band_energies[m], half_de, sqrt2_sigma);
    if (lam != 0.0) acc = __dadd_rn(acc, __dmul_rn(w[m], lam));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, o));
  if (lane == 0) dN[(long long)f * nener + ie] = acc;
}

}  // namespace

void launch_delta_n(const double* d_grid, int nener, double sigma,
                    const double* d_band_energies, const double* d_weights,
                    int n_bands, int n_fold, double* d_dN, cudaStream_t s) {
  const long long warps = (long long)nener * n_fold;
  if (warps <= 0) return;
  const int warps_per_block = 4;
  k3_delta_n<<<(unsigned)((warps + warps_per_block - 1) / warps_per_block), 128, 0, s>>>(
      d_grid, nener, sigma, d_band_energies, d_weights, n_bands, n_fold, d_dN);
}

}  // namespace bandup
