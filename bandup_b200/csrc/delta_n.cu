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
// One warp per energy bin (all folds), eight consecutive bins per CTA; candidate bands are
// compacted first (CTA window, then the warp's own bin) and their lambdas evaluated one per lane,
// a shuffle tree combines them (fixed order: deterministic).  Work is tiny next to K2.
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

// The finalisation of K2's per-tile partials, partials[band][chunk][0 = norm, 1 + fold]:
// s = sum_chunks norm, S = sum_chunks fold (fixed chunk order: bit-reproducible), P = S / s.
__device__ __forceinline__ double band_norm(const double* __restrict__ p, int n_chunks, int nslots) {
  double s = 0.0;
#pragma unroll 4
  for (int c = 0; c < n_chunks; ++c) s += p[(long long)c * nslots];
  return s;
}
__device__ __forceinline__ double band_weight(const double* __restrict__ p, int n_chunks, int nslots, int q,
                                              double s, int perform_unfold) {
  double S = 0.0;
#pragma unroll 4
  for (int c = 0; c < n_chunks; ++c) S += p[(long long)c * nslots + q];
  // io_routines_mod.f90:1325: bands with s <= epsilon(1.0_dp) stay unscaled
  const double P = (s > 2.220446049250313e-16) ? S / s : S;
  // -dont_unfold: spectral_weight = 1.0_dp (band_unfolding_routines_mod.f90:1346-1351)
  return perform_unfold ? P : 1.0;
}

constexpr int kK3Threads = 256;    // 8 warps
constexpr int kK3Warps = kK3Threads / 32;
constexpr int kK3MaxBinsPerWarp = 4;  // consecutive energy bins a warp may own
constexpr int kK3BandTile = 2048;  // bands examined per pass (8 per thread)
constexpr int kK3PerThread = kK3BandTile / kK3Threads;
static_assert(kK3PerThread * kK3Warps == 64, "the count scan below handles two entries per lane of one warp");

// grid (ceil(nener / (8 B)), n_kpts): every warp owns B consecutive energy bins (B = bins_per_warp,
// 1..4, chosen by the launcher so that a batch of many small k-points does not drown in CTAs), for
// ALL folds of the k-point.  Which bands can contribute to a bin, and lambda itself, depend on the
// bin and the band energy only -- not on the fold -- so the scan and the erf evaluations are done
// once and reused for every fold.
//
// Two-level candidate search.  The CTA's 8 B bins span a narrow energy window; the whole CTA
// first compacts, in ascending band order, the bands within reach of that window (8 independent
// coalesced loads per thread, one ballot each, one 64-entry scan) -- typically a few dozen of
// thousands -- and each warp then tests only those against its own bins, collecting (bin, band)
// pairs; lambda is evaluated for up to 32 pairs at once, one per lane, and the pairs of one bin
// are combined with a shuffle tree (fixed order: deterministic).
//
// Nothing up to here reads what K2 wrote: launched with programmatic stream serialisation the CTA
// does its search while K2 still runs and waits (griddepcontrol.wait) only before it touches the
// partials.  The kernel also finishes K2 (what used to be a launch of its own): every CTA turns
// its slice of the (band, slot) partials into the spectral weights P = S/s and the band norms,
// CTA 0 of a k-point adds up K1's per-CTA plane-wave counts.  The bins do not wait for that: a bin
// needs P of its few candidate bands only and forms it from the same partials with the same
// arithmetic.
__global__ void __launch_bounds__(kK3Threads, 4)
k3_finalize_delta_n(const KptGeom* __restrict__ geoms, const double* __restrict__ grid, int nener, double sigma,
                    int bins_per_warp) {
  __shared__ double s_ener[kK3BandTile];  // candidates of the CTA window: energy ...
  __shared__ int s_band[kK3BandTile];     // ... and band index, ascending
  __shared__ int s_cnt[kK3PerThread * kK3Warps], s_off[kK3PerThread * kK3Warps + 1];
  __shared__ int s_hit[kK3Warps][64];
  __shared__ double s_hit_e[kK3Warps][64];
  __shared__ unsigned char s_hit_b[kK3Warps][64];
  __shared__ double s_acc[kK3Warps][kK3MaxBinsPerWarp][kMaxFolds];
  const KptGeom& g = geoms[blockIdx.y];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_bands = g.n_bands, n_fold = g.n_fold;
  const double* __restrict__ band_energies = g.energies;
  const double* __restrict__ partials = g.partials;
  const int n_chunks = g.n_parts, nslots = n_fold + 1, perform_unfold = g.perform_unfold;  // n_chunks: partial rows per band
  const int B = bins_per_warp;
  const int ie_first = blockIdx.x * (kK3Warps * B);
  const int ie_last = min(ie_first + kK3Warps * B, nener) - 1;
  const int ie0 = ie_first + warp * B;                 // this warp's first bin
  const int n_my = max(0, min(B, nener - ie0));        // and how many it owns
  // delta_e = energies(2) - energies(1)  (:735)
  const double g0 = grid[0], g1 = grid[1], Ea = grid[ie_first], Eb = grid[ie_last];
  const double delta_e = __dsub_rn(g1, g0);
  const double half_de = __dmul_rn(0.5, delta_e);
  const double sqrt2_sigma = __dmul_rn(sqrt(2.0), sigma);
  const double reach = half_de + 8.0 * sqrt2_sigma;
  // CTA window: a superset of every bin's |E_m - E_i| <= guard (the grid is monotonic)
  const double w_lo = fmin(Ea, Eb), w_hi = fmax(Ea, Eb);
  const double w_guard = 1.000001 * (reach + 1e-9 * (fmax(fabs(w_lo), fabs(w_hi)) + reach));
  int* hits = s_hit[warp];
  double* hits_e = s_hit_e[warp];
  unsigned char* hits_b = s_hit_b[warp];
  for (int i = lane; i < kK3MaxBinsPerWarp * kMaxFolds; i += 32) (&s_acc[warp][0][0])[i] = 0.0;
  __syncwarp();
  int n_hit = 0;  // warp-uniform
  bool waited = false;
  // The few (bin, band) pairs that can contribute (|E_m - E_i| within the guard band) are collected
  // bin by bin in ascending band order, then lambda (two erf, two divisions) is evaluated for up to
  // 32 of them at once, one per lane; every (bin, fold) then takes its sum_m P_m(f) * lambda_m with a
  // shuffle tree (fixed order: deterministic).
  auto flush = [&](int count) {
    if (!waited) {  // the first touch of K2's partials
      asm volatile("griddepcontrol.wait;" ::: "memory");
      waited = true;
    }
    double lam = 0.0;
    int m = 0, b = 0;
    if (lane < count) {
      m = hits[lane];
      b = hits_b[lane];
      lam = lambda_box_erf(grid[ie0 + b], hits_e[lane], half_de, sqrt2_sigma);
    }
    if (__ballot_sync(0xffffffffu, lam != 0.0) != 0u) {
      const double* p = partials + (long long)m * n_chunks * nslots;
      const double s = lam != 0.0 ? band_norm(p, n_chunks, nslots) : 0.0;
      for (int f0 = 0; f0 < n_fold; f0 += 4) {  // the partial loads of four folds in flight together
        double w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          w[k] = (lam != 0.0 && f0 + k < n_fold) ? band_weight(p, n_chunks, nslots, f0 + k + 1, s, perform_unfold) : 0.0;
        for (int bb = 0; bb < n_my; ++bb) {
          if (__ballot_sync(0xffffffffu, lam != 0.0 && b == bb) == 0u) continue;  // nothing for this bin
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (f0 + k >= n_fold) break;
            double v = (lam != 0.0 && b == bb) ? __dmul_rn(w[k], lam) : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
            if (lane == 0) s_acc[warp][bb][f0 + k] = __dadd_rn(s_acc[warp][bb][f0 + k], v);
          }
        }
      }
    }
    __syncwarp();
  };
  for (int m0 = 0; m0 < n_bands; m0 += kK3BandTile) {
    const int nt = min(kK3BandTile, n_bands - m0);
    // (a) CTA-level ordered compaction of the window's candidates
    double e[kK3PerThread];
    unsigned bal[kK3PerThread];
#pragma unroll
    for (int j = 0; j < kK3PerThread; ++j) {
      const int i = j * kK3Threads + threadIdx.x;
      e[j] = i < nt ? band_energies[m0 + i] : 0.0;
    }
    __syncthreads();  // the previous tile's candidates are no longer read
#pragma unroll
    for (int j = 0; j < kK3PerThread; ++j) {
      const int i = j * kK3Threads + threadIdx.x;
      const bool cand = i < nt && !(e[j] < w_lo - w_guard) && !(e[j] > w_hi + w_guard);
      bal[j] = __ballot_sync(0xffffffffu, cand);
      if (lane == 0) s_cnt[j * kK3Warps + warp] = __popc(bal[j]);
    }
    __syncthreads();
    if (warp == 0) {  // exclusive scan of the 64 (pass, warp) counts, in band order
      constexpr int n = kK3PerThread * kK3Warps;
      const int a = s_cnt[2 * lane], b2 = s_cnt[2 * lane + 1];
      int x = a + b2;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
      }
      s_off[2 * lane] = x - a - b2;
      s_off[2 * lane + 1] = x - b2;
      if (lane == 31) s_off[n] = x;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kK3PerThread; ++j) {
      if (bal[j] >> lane & 1u) {
        const int pos = s_off[j * kK3Warps + warp] + __popc(bal[j] & ((1u << lane) - 1u));
        s_ener[pos] = e[j];
        s_band[pos] = m0 + j * kK3Threads + threadIdx.x;
      }
    }
    __syncthreads();
    const int nc = s_off[kK3PerThread * kK3Warps];
    // (b) this warp's bins against the window's candidates
    for (int bb = 0; bb < n_my; ++bb) {
      const double E = grid[ie0 + bb];
      const double guard = reach + 1e-9 * (fabs(E) + reach);
      for (int i0 = 0; i0 < nc; i0 += 32) {
        const int i = i0 + lane;
        const double em = i < nc ? s_ener[i] : 0.0;
        const bool hit = i < nc && !(fabs(em - E) > guard);
        const unsigned hb = __ballot_sync(0xffffffffu, hit);
        if (hb == 0u) continue;
        if (hit) {
          const int at = n_hit + __popc(hb & ((1u << lane) - 1u));
          hits[at] = s_band[i];
          hits_e[at] = em;
          hits_b[at] = (unsigned char)bb;
        }
        n_hit += __popc(hb);
        __syncwarp();
        if (n_hit >= 32) {
          flush(32);
          if (lane < n_hit - 32) {
            hits[lane] = hits[32 + lane];
            hits_e[lane] = hits_e[32 + lane];
            hits_b[lane] = hits_b[32 + lane];
          }
          n_hit -= 32;
          __syncwarp();
        }
      }
    }
  }
  if (n_hit > 0) flush(n_hit);
  for (int i = lane; i < n_my * n_fold; i += 32) {
    const int bb = i / n_fold, f = i - bb * n_fold;
    g.dN[(long long)f * nener + ie0 + bb] = s_acc[warp][bb][f];
  }
  if (!waited) asm volatile("griddepcontrol.wait;" ::: "memory");
  {  // ---- finalisation of K2: this CTA's slice of the (band, slot) pairs -------------------
    const long long items = (long long)n_bands * nslots;
    const long long per_cta = (items + gridDim.x - 1) / gridDim.x;
    const long long hi = min(items, per_cta * (blockIdx.x + 1));
    for (long long idx = per_cta * blockIdx.x + threadIdx.x; idx < hi; idx += kK3Threads) {
      const int band = (int)(idx / nslots), q = (int)(idx - (long long)band * nslots);
      const double* p = partials + (long long)band * n_chunks * nslots;
      const double s = band_norm(p, n_chunks, nslots);
      if (q == 0) {
        if (g.norms) g.norms[band] = s;
      } else {
        g.weights[(long long)(q - 1) * n_bands + band] = band_weight(p, n_chunks, nslots, q, s, perform_unfold);
      }
    }
    if (blockIdx.x == 0) {  // size(selected_coeff_indices) of every fold: K1's per-CTA counts
      for (int f = warp; f < n_fold; f += kK3Warps) {
        int n = 0;
        for (int bk = lane; bk < g.k1_blocks; bk += 32) n += g.block_counts[bk * n_fold + f];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
        if (lane == 0) g.n_selected[f] = n;
      }
    }
  }
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
                    int nener, double sigma, int sm_count, bool pdl, cudaStream_t s) {
  if (nener <= 0 || max_n_fold <= 0 || n_kpts <= 0) return;  // set_grid guarantees nener >= 1
  // One bin per warp while the CTAs fit the machine about once (4 resident per SM); a batch of many
  // k-points gets up to four bins per warp instead of several waves of short-lived CTAs, each of
  // which would pay the candidate search and its barriers again (Si 3x3x3: 99 k-points x 63 CTAs).
  const long long ctas_b1 = (long long)n_kpts * ((nener + kK3Warps - 1) / kK3Warps);
  const long long resident = 4LL * (sm_count > 0 ? sm_count : 148);
  int bpw = (int)((ctas_b1 + resident - 1) / resident);
  if (bpw < 1) bpw = 1;
  if (bpw > kK3MaxBinsPerWarp) bpw = kK3MaxBinsPerWarp;
  const int bins_per_cta = kK3Warps * bpw;
  dim3 grid((unsigned)((nener + bins_per_cta - 1) / bins_per_cta), (unsigned)n_kpts);
  launch_kernel(k3_finalize_delta_n, grid, dim3(kK3Threads), 0, s, pdl, d_geoms, d_grid, nener, sigma, bpw);
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
