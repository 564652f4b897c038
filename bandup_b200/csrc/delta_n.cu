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
// exactly zero, so skipping changes nothing.  A subtract-and-compare prefilter
// with a conservative guard band (|E_m - E_i| > delta_e/2 + 8 sqrt2 sigma + 1e-9
// relative slack) rejects almost every pair before any division or erf.
//
// One warp per (energy bin, fold), eight bins of a fold per CTA; the band energies are
// staged through shared memory in tiles (a dependent global load per band made the first
// version latency-bound: 38 us for 3000 bands), lanes stride over the bands and a shuffle
// tree combines them (fixed order: deterministic).  Work is tiny next to K2.
// Candidate bands are compacted first and their lambdas evaluated one per lane (see the kernel).
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

constexpr int kK3Threads = 256;   // 8 warps = 8 energy bins of one fold per CTA
constexpr int kK3BandTile = 2048;  // band energies staged in shared memory per pass

// grid (ceil(nener/8), n_kpts): one warp per energy bin, ALL folds of the k-point.  Which bands
// can contribute to a bin, and lambda itself, depend on the bin and the band energy only -- not on
// the fold -- so the scan and the erf evaluations are done once and reused for every fold.
__global__ void __launch_bounds__(kK3Threads)
k3_delta_n(const KptGeom* __restrict__ geoms, const double* __restrict__ grid, int nener, double sigma) {
  __shared__ double s_ener[kK3BandTile];
  __shared__ int s_hit[kK3Threads / 32][64];
  __shared__ double s_acc[kK3Threads / 32][kMaxFolds];
  const KptGeom& g = geoms[blockIdx.y];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_bands = g.n_bands, n_fold = g.n_fold;
  const double* __restrict__ band_energies = g.energies;
  const double* __restrict__ weights = g.weights;
  const int ie = blockIdx.x * (kK3Threads / 32) + warp;
  const bool live = ie < nener;
  // delta_e = energies(2) - energies(1)  (:735)
  const double delta_e = __dsub_rn(grid[1], grid[0]);
  const double half_de = __dmul_rn(0.5, delta_e);
  const double sqrt2_sigma = __dmul_rn(sqrt(2.0), sigma);
  const double E = live ? grid[ie] : 0.0;
  const double reach = half_de + 8.0 * sqrt2_sigma;
  const double guard = reach + 1e-9 * (fabs(E) + reach);
  int* hits = s_hit[warp];
  double* acc = s_acc[warp];
  for (int f = lane; f < n_fold; f += 32) acc[f] = 0.0;
  __syncwarp();
  int n_hit = 0;  // warp-uniform
  // The few bands that can contribute (|E_m - E_i| within the guard band) are first collected,
  // in ascending band order, then lambda (two erf, two divisions) is evaluated for up to 32 of
  // them at once, one per lane; every fold then takes its sum_m P_m(f) * lambda_m with a
  // shuffle tree (fixed order: deterministic).
  auto flush = [&](int count) {
    double lam = 0.0;
    int m = 0;
    if (lane < count) {
      m = hits[lane];
      lam = lambda_box_erf(E, band_energies[m], half_de, sqrt2_sigma);
    }
    if (__ballot_sync(0xffffffffu, lam != 0.0) != 0u) {
      for (int f = 0; f < n_fold; ++f) {
        double v = (lam != 0.0) ? __dmul_rn(weights[(long long)f * n_bands + m], lam) : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
        if (lane == 0) acc[f] = __dadd_rn(acc[f], v);
      }
    }
    __syncwarp();
  };
  for (int m0 = 0; m0 < n_bands; m0 += kK3BandTile) {
    const int nt = min(kK3BandTile, n_bands - m0);
    __syncthreads();
    for (int i = threadIdx.x; i < nt; i += kK3Threads) s_ener[i] = band_energies[m0 + i];
    __syncthreads();
    if (!live) continue;
    for (int i0 = 0; i0 < nt; i0 += 32) {
      const int i = i0 + lane;
      const bool hit = i < nt && !(fabs(s_ener[i] - E) > guard);
      const unsigned bal = __ballot_sync(0xffffffffu, hit);
      if (bal == 0u) continue;
      if (hit) hits[n_hit + __popc(bal & ((1u << lane) - 1u))] = m0 + i;
      n_hit += __popc(bal);
      __syncwarp();
      if (n_hit >= 32) {
        flush(32);
        if (lane < n_hit - 32) hits[lane] = hits[32 + lane];
        n_hit -= 32;
        __syncwarp();
      }
    }
  }
  if (!live) return;
  if (n_hit > 0) flush(n_hit);
  for (int f = lane; f < n_fold; f += 32) g.dN[(long long)f * nener + ie] = acc[f];
}

// A(k, E_i) = sum_m P_m * delta(E_m - E_i), delta = Gaussian of standard deviation sigma
// (calc_spectral_function, band_unfolding_routines_mod.f90:654-698, with delta, math_mod.f90:159-178;
// dead code in the reference build -- calc_spec_func_explicitly = .FALSE.,
// constants_and_types_mod.f90:75 -- offered here on request).  One warp per (energy, fold).
__global__ void __launch_bounds__(kK3Threads)
k3_spectral_function(const double* __restrict__ grid, int nener, double stddev,
                     const double* __restrict__ band_energies, const double* __restrict__ weights,
                     int n_bands, int n_fold, double* __restrict__ sf) {
  const int lane = threadIdx.x & 31;
  const long long wg = (long long)blockIdx.x * (kK3Threads / 32) + (threadIdx.x >> 5);
  if (wg >= (long long)nener * n_fold) return;
  const int f = (int)(wg / nener), ie = (int)(wg - (long long)f * nener);
  const double E = grid[ie];
  const double twopi = 2.0 * (4.0 * atan(1.0));
  const double denom = __dmul_rn(stddev, sqrt(twopi));
  const double* w = weights + (long long)f * n_bands;
  double acc = 0.0;
  for (int m = lane; m < n_bands; m += 32) {
    const double x = __ddiv_rn(__dsub_rn(band_energies[m], E), stddev);
    const double a = __ddiv_rn(__dmul_rn(x, x), 2.0);
    if (a > 746.0) continue;  // exp(-a) is exactly 0 in double precision
    acc = __dadd_rn(acc, __dmul_rn(w[m], __ddiv_rn(exp(-a), denom)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, o));
  if (lane == 0) sf[(long long)f * nener + ie] = acc;
}

}  // namespace

void launch_delta_n(const KptGeom* d_geoms, int n_kpts, int max_n_fold, const double* d_grid,
                    int nener, double sigma, cudaStream_t s) {
  if (nener <= 0 || max_n_fold <= 0 || n_kpts <= 0) return;
  const int bins_per_cta = kK3Threads / 32;
  dim3 grid((unsigned)((nener + bins_per_cta - 1) / bins_per_cta), (unsigned)n_kpts);
  k3_delta_n<<<grid, kK3Threads, 0, s>>>(d_geoms, d_grid, nener, sigma);
}

void launch_spectral_function(const double* d_grid, int nener, double stddev, const double* d_band_energies,
                              const double* d_weights, int n_bands, int n_fold, double* d_sf,
                              cudaStream_t s) {
  const long long warps = (long long)nener * n_fold;
  if (warps <= 0) return;
  const int wpb = kK3Threads / 32;
  k3_spectral_function<<<(unsigned)((warps + wpb - 1) / wpb), kK3Threads, 0, s>>>(
      d_grid, nener, stddev, d_band_energies, d_weights, n_bands, n_fold, d_sf);
}

}  // namespace bandup
