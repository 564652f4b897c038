// rho.cu -- K4/K5: the projected variant (`bandup unfold --orbitals` -> `projected-unfold`).
//
// K4a k4_rho_plan     which (fold, energy) get an unfolding-density operator and how many
//                     band pairs each couples  (band_unfolding_routines_mod.f90:1381-1395,
//                     calc_rho :1096-1106)
// K4b k4_gather       the selected plane-wave coefficients of every band, renormalised and
//                     rounded to complex64 exactly like the reference's in-place rescale
//                     (io_routines_mod.f90:1325-1328); replaces the pf_coeffs_m1/m2 copies
//                     calc_rho makes per band pair (:1112-1128)
// K4c k4_rho_pairs    rho(m1,m2) = prefactor * sum_spinor dot_product(C_m1[sel], C_m2[sel])
//                     (:1107-1134), one warp per band pair, upper triangle
// K5a k5_orb_contrib  orbital-contribution matrix O (orbital_contributions.py:226-362)
// K5b k5_trace        clip(dN * Re Tr(rho.O), 0, dN)  (unfolding_density_op.py:121-141,
//                     orbital_contributions.py:444-458)
//
// None of this is bandwidth-critical: with the default box lambda only the few bands inside
// one energy bin couple, and only n_pw/det(M) plane waves per band are touched.
#include "kernels.cuh"

namespace bandup {

namespace {

constexpr double kEps = 2.220446049250313e-16;  // epsilon(1.0_dp)
constexpr double kMinPrefactor = 1e-5;          // calc_rho :1076

// lambda (:701-716) with the same explicitly rounded operations as K3.
__device__ __forceinline__ double lambda_dev(double pc_ener, double sc_ener, double half_de,
                                             double sqrt2_sigma) {
  const double lower = __dsub_rn(pc_ener, half_de);
  const double upper = __dadd_rn(pc_ener, half_de);
  const double x1 = __ddiv_rn(__dsub_rn(lower, sc_ener), sqrt2_sigma);
  const double x2 = __ddiv_rn(__dsub_rn(upper, sc_ener), sqrt2_sigma);
  if (x1 > 6.0 || x2 < -6.0) return 0.0;
  return __dmul_rn(0.5, __dsub_rn(erf(x2), erf(x1)));
}

struct ActiveList {  // bands with lambda >= 1e-5 at one energy, ascending band index
  int* band;
  double* lam;
};

// Block-wide ordered compaction of the active bands of energy E into shared memory.
// Returns the number found (may exceed cap: the caller flags the overflow).
__device__ int build_active(const double* __restrict__ band_energies, int n_bands, double E,
                            double half_de, double sqrt2_sigma, int cap, ActiveList out,
                            int* s_scratch /*[blockDim/32 + 1]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const double reach = half_de + 8.0 * sqrt2_sigma;
  const double guard = reach + 1e-9 * (fabs(E) + reach);
  int base = 0;
  for (int m0 = 0; m0 < n_bands; m0 += blockDim.x) {
    const int m = m0 + threadIdx.x;
    double lam = 0.0;
    if (m < n_bands) {
      const double Em = band_energies[m];
      if (fabs(Em - E) <= guard) lam = lambda_dev(E, Em, half_de, sqrt2_sigma);
    }
    const bool hit = lam >= kMinPrefactor;  // `if(lamb_ee1 < min_prefactor) cycle`
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) s_scratch[warp] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < nwarp; ++w) {
      const int c = s_scratch[w];
      if (w < warp) before += c;
      total += c;
    }
    if (hit) {
      const int pos = base + before + __popc(bal & ((1u << lane) - 1u));
      if (pos < cap) {
        out.band[pos] = m;
        out.lam[pos] = lam;
      }
    }
    base += total;
    __syncthreads();
  }
  return base;
}

// number of b >= a with (lam_a * lam_b) / dN >= 1e-5, for row a
// perform_unfold == 0: the -dont_unfold branch of calc_rho (:1136-1154), whose m1 loop takes ALL bands
__device__ __forceinline__ int row_pairs(const ActiveList& act, int n_act, int a, double dN,
                                         int n_bands, int perform_unfold) {
  if (perform_unfold && act.band[a] >= n_bands - 1) return 0;  // `do m1=1, n_SC_bands - 1` (:1096)
  int c = 0;
  for (int b = a; b < n_act; ++b)
    if (__ddiv_rn(__dmul_rn(act.lam[a], act.lam[b]), dN) >= kMinPrefactor) ++c;
  return c;
}

__global__ void __launch_bounds__(256)
k4_rho_plan(const double* __restrict__ grid, int nener, double sigma,
            const double* __restrict__ band_energies, int n_bands,
            const double* __restrict__ dN_all, double min_dN, int cap, int perform_unfold,
            int* __restrict__ pair_count /*[n_fold][nener]*/, int* __restrict__ overflow) {
  extern __shared__ unsigned char smem[];
  ActiveList act;
  act.lam = reinterpret_cast<double*>(smem);
  act.band = reinterpret_cast<int*>(act.lam + cap);
  int* s_scratch = act.band + cap;
  __shared__ int s_total;
  const int f = blockIdx.x / nener, ie = blockIdx.x - f * nener;
  const double dN = dN_all[(long long)f * nener + ie];
  if (!(dN >= min_dN)) {  // "Can hardly be considered a band otherwise" (:1388)
    if (threadIdx.x == 0) pair_count[blockIdx.x] = 0;
    return;
  }
  const double delta_e = __dsub_rn(grid[1], grid[0]);
  const double half_de = __dmul_rn(0.5, delta_e);
  const double sqrt2_sigma = __dmul_rn(sqrt(2.0), sigma);
  const int n_act = build_active(band_energies, n_bands, grid[ie], half_de, sqrt2_sigma, cap, act, s_scratch);
  if (n_act > cap) {
    if (threadIdx.x == 0) {
      atomicExch(overflow, 1);
      pair_count[blockIdx.x] = 0;
    }
    return;
  }
  if (threadIdx.x == 0) s_total = 0;
  __syncthreads();
  int mine = 0;
  for (int a = threadIdx.x; a < n_act; a += blockDim.x) mine += row_pairs(act, n_act, a, dN, n_bands, perform_unfold);
  if (mine) atomicAdd(&s_total, mine);  // integer sum: order-independent
  __syncthreads();
  if (threadIdx.x == 0) pair_count[blockIdx.x] = s_total;
}

// exclusive prefix sum of n ints by ONE block; total -> out[n]
__global__ void __launch_bounds__(1024)
k4_scan(const int* __restrict__ in, int n, long long* __restrict__ out) {
  __shared__ long long s_warp[32];
  __shared__ long long s_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += 1024) {
    const int i = i0 + threadIdx.x;
    long long v = i < n ? in[i] : 0, x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    long long before = 0, total = 0;
    for (int w = 0; w < 32; ++w) {
      if (w < warp) before += s_warp[w];
      total += s_warp[w];
    }
    if (i < n) out[i] = s_base + before + x - v;
    __syncthreads();
    if (threadIdx.x == 0) s_base += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) out[n] = s_base;
}

template <int MODE>
__device__ __forceinline__ long long elem_of(int pw, int is, int n_pw) {
  if (MODE == 0) return pw;
  if (MODE == 1) return 2LL * pw + is;
  return (long long)is * n_pw + pw;
}

// csel layout per fold u: [band][spinor][j], j < n_sel(u); folds concatenated by sel_off
template <int MODE, typename T>  // T = float2 (complex64) or double2 (complex128 build, rounded here)
__global__ void __launch_bounds__(256)
k4_gather(const char* __restrict__ coeffs, long long stride_bytes, int n_pw, int n_spinor,
          int n_bands, const int32_t* __restrict__ lists, const long long* __restrict__ sel_off,
          const double* __restrict__ norms, float2* __restrict__ csel) {
  const int band = blockIdx.x, u = blockIdx.y;
  const long long off = sel_off[u];
  const int n_sel = (int)(sel_off[u + 1] - off);
  const T* row = reinterpret_cast<const T*>(coeffs + (long long)band * stride_bytes);
  const double s = norms[band];
  // inner_prod > epsilon: coeffs = (1/sqrt(|inner_prod|)) * coeffs, rounded back to complex(sp)
  const double fac = s > kEps ? 1.0 / sqrt(fabs(s)) : 1.0;
  float2* dst = csel + (off * n_bands + (long long)band * n_sel) * n_spinor;
  for (int t = threadIdx.x; t < n_sel * n_spinor; t += blockDim.x) {
    const int is = t / n_sel, j = t - is * n_sel;
    const int pw = lists[off + j] - 1;  // lists are 1-based like selected_coeff_indices
    const T c = row[elem_of<MODE>(pw, is, n_pw)];
    float2 r;
    if (s > kEps) {
      r.x = (float)(fac * (double)c.x);
      r.y = (float)(fac * (double)c.y);
    } else {
      r.x = (float)c.x;
      r.y = (float)c.y;
    }
    dst[t] = r;
  }
}

__global__ void __launch_bounds__(256)
k4_rho_pairs(const double* __restrict__ grid, int nener, double sigma,
             const double* __restrict__ band_energies, int n_bands,
             const double* __restrict__ dN_all, int cap, int n_spinor, int perform_unfold,
             const long long* __restrict__ sel_off, const float2* __restrict__ csel,
             const int* __restrict__ pair_count, const long long* __restrict__ pair_off,
             RhoPair* __restrict__ pairs) {
  extern __shared__ unsigned char smem[];
  ActiveList act;
  act.lam = reinterpret_cast<double*>(smem);
  act.band = reinterpret_cast<int*>(act.lam + cap);
  int* row_off = act.band + cap;       // [cap + 1]
  int* s_scratch = row_off + cap + 1;  // [blockDim/32 + 1]
  if (pair_count[blockIdx.x] == 0) return;
  const int f = blockIdx.x / nener, ie = blockIdx.x - f * nener;
  const double dN = dN_all[(long long)f * nener + ie];
  const double delta_e = __dsub_rn(grid[1], grid[0]);
  const double half_de = __dmul_rn(0.5, delta_e);
  const double sqrt2_sigma = __dmul_rn(sqrt(2.0), sigma);
  const int n_act = build_active(band_energies, n_bands, grid[ie], half_de, sqrt2_sigma, cap, act, s_scratch);
  for (int a = threadIdx.x; a < n_act; a += blockDim.x)
    row_off[a + 1] = row_pairs(act, n_act, a, dN, n_bands, perform_unfold);
  __syncthreads();
  if (threadIdx.x == 0) {
    row_off[0] = 0;
    for (int a = 0; a < n_act; ++a) row_off[a + 1] += row_off[a];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const long long off = sel_off[f];
  const int n_sel = (int)(sel_off[f + 1] - off);
  const int len = n_sel * n_spinor;
  const float2* base = csel + off * n_bands * n_spinor;
  RhoPair* out = pairs + pair_off[blockIdx.x];
  for (int a = warp; a < n_act; a += nwarp) {
    if (row_off[a + 1] == row_off[a]) continue;
    const int m1 = act.band[a];
    const float2* c1 = base + (long long)m1 * len;
    int k = row_off[a];
    for (int b = a; b < n_act; ++b) {
      const double pref = __ddiv_rn(__dmul_rn(act.lam[a], act.lam[b]), dN);
      if (pref < kMinPrefactor) continue;
      const int m2 = act.band[b];
      if (!perform_unfold) {
        // -dont_unfold (:1136-1154): rho(m1,m2) = 1 for every coupled pair, then the whole matrix is
        // divided by the number of bands with lambda >= 1e-5 -- the active list; complex(sp)/real(dp)
        // is formed in dp and stored to sp.  With the box lambda the full matrix is symmetric, so
        // the upper triangle kept here expands to it.
        if (lane == 0) {
          RhoPair p;
          p.m1 = m1;
          p.m2 = m2;
          p.re = (float)(1.0 / (double)n_act);
          p.im = 0.0f;
          out[k] = p;
        }
        ++k;
        continue;
      }
      const float2* c2 = base + (long long)m2 * len;
      // dot_product(a, b) = sum conj(a) * b, accumulated in fp64 (the reference accumulates in
      // complex(sp); the fp64 sum is the better-conditioned of the two)
      double re = 0.0, im = 0.0;
      for (int t = lane; t < len; t += 32) {
        const float2 x = c1[t], y = c2[t];
        re += (double)x.x * (double)y.x + (double)x.y * (double)y.y;
        im += (double)x.x * (double)y.y - (double)x.y * (double)y.x;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        re += __shfl_xor_sync(0xffffffffu, re, o);
        im += __shfl_xor_sync(0xffffffffu, im, o);
      }
      if (lane == 0) {
        RhoPair p;
        p.m1 = m1;
        p.m2 = m2;
        p.re = (float)(pref * re);  // rho(m1,m2) = prefactor * rho(m1,m2), stored as complex(sp)
        p.im = (m1 == m2) ? 0.0f : (float)(pref * im);
        out[k] = p;
      }
      ++k;
    }
  }
}

// O(m1,m2) = sum_{a in mask} dual[m1,a] * proj[a,m2] for |E_m2 - E_m1| <= max_dE, zeroed when
// |O| < min_abs (orbital_contributions.py:337-353; use_dual=True, force_hermitian=False)
__global__ void __launch_bounds__(256)
k5_orb_contrib(int n_bands, int n_ao, const double2* __restrict__ dual /*[nb][n_ao]*/,
               const double2* __restrict__ proj /*[n_ao][nb]*/, const uint8_t* __restrict__ mask,
               const double* __restrict__ ener, double max_dE, double min_abs,
               double2* __restrict__ O /*[nb][nb]*/) {
  const int m1 = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const double e1 = ener[m1];
  for (int m2 = warp; m2 < n_bands; m2 += nwarp) {
    double re = 0.0, im = 0.0;
    const bool in_window = !(fabs(__dsub_rn(ener[m2], e1)) > max_dE);
    if (in_window) {
      for (int a = lane; a < n_ao; a += 32) {
        if (!mask[a]) continue;
        const double2 d = dual[(long long)m1 * n_ao + a];
        const double2 p = proj[(long long)a * n_bands + m2];
        re += d.x * p.x - d.y * p.y;
        im += d.x * p.y + d.y * p.x;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        re += __shfl_xor_sync(0xffffffffu, re, o);
        im += __shfl_xor_sync(0xffffffffu, im, o);
      }
      if (hypot(re, im) < min_abs) re = im = 0.0;  // `if(abs(contr) < 1E-2): continue`
    }
    if (lane == 0) O[(long long)m1 * n_bands + m2] = make_double2(re, im);
  }
}

// out[f][ie] = clip(dN * Re Tr(rho.O), 0, dN); rho Hermitian-completed from its upper
// triangle with a real diagonal, entries below min_abs_rho dropped (the rho file keeps
// |rho| >= 1e-4, io_routines_mod.f90:1861-1862).  One warp per (fold, energy).
__global__ void __launch_bounds__(128)
k5_trace(int n_cells, int n_bands, const double* __restrict__ dN_all,
         const long long* __restrict__ pair_off, const RhoPair* __restrict__ pairs,
         const double2* __restrict__ O, double min_abs_rho, int clip, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int cell = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (cell >= n_cells) return;
  const long long p0 = pair_off[cell], p1 = pair_off[cell + 1];
  double re = 0.0;
  for (long long p = p0 + lane; p < p1; p += 32) {
    const RhoPair r = pairs[p];
    if (hypotf(r.re, r.im) < (float)min_abs_rho) continue;
    if (r.m1 == r.m2) {
      re += (double)r.re * O[(long long)r.m1 * n_bands + r.m1].x;
    } else {
      const double2 o21 = O[(long long)r.m2 * n_bands + r.m1];  // rho(m1,m2) * O(m2,m1)
      const double2 o12 = O[(long long)r.m1 * n_bands + r.m2];  // conj(rho(m1,m2)) * O(m1,m2)
      re += (double)r.re * o21.x - (double)r.im * o21.y;
      re += (double)r.re * o12.x + (double)r.im * o12.y;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) re += __shfl_xor_sync(0xffffffffu, re, o);
  if (lane == 0) {
    const double N = dN_all[cell];
    double v = (p1 > p0) ? N * re : 0.0;
    if (clip && p1 > p0) v = fmin(fmax(v, 0.0), N);
    out[cell] = v;
  }
}

// K6a -- Pauli matrix elements between the two bands of every rho pair over ALL plane waves
// (get_SC_pauli_matrix_elmnts, band_unfolding_routines_mod.f90:1164-1263), contracted on the
// spot with rho: contrib[p][i] = the share of pair p in Re Tr(rho . sigma_i) (:1468-1476 with
// trace_AB_A_is_rho_B_is_cplx, math_mod.f90:321-341):
//   sigma_x(m2,m1) = <m2 a|m1 b> + <m2 b|m1 a>,  sigma_y = i(<m2 b|m1 a> - <m2 a|m1 b>),
//   sigma_z = <m2 a|m1 a> - <m2 b|m1 b>,   <u|v> = sum_G conj(u) v  (renormalised coefficients)
//   upper element rho(m1,m2), m1 < m2:  rho12 sigma(m2,m1) + conj(rho12) conj(sigma(m2,m1))
// Only stored elements (|rho| >= 1e-5) enter, as in the reference's rhos(:).  One CTA per pair.
template <int MODE, typename T>  // 1: [pw][spinor] interleaved, 2: [spinor][pw] blocked; T = float2 | double2
__global__ void __launch_bounds__(256)
k6_pauli_pairs(const char* __restrict__ coeffs, long long stride_bytes, int n_pw,
               const double* __restrict__ norms, const RhoPair* __restrict__ pairs,
               double* __restrict__ contrib) {
  __shared__ double s_red[8][8];
  const RhoPair pr = pairs[blockIdx.x];
  double* out = contrib + 3LL * blockIdx.x;
  if (hypotf(pr.re, pr.im) < 1e-5f) {  // not stored by perform_unfolding (:1418)
    if (threadIdx.x < 3) out[threadIdx.x] = 0.0;
    return;
  }
  const T* r1 = reinterpret_cast<const T*>(coeffs + (long long)pr.m1 * stride_bytes);
  const T* r2 = reinterpret_cast<const T*>(coeffs + (long long)pr.m2 * stride_bytes);
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // ab, ba, aa, bb (re, im each)
  for (int pw = threadIdx.x; pw < n_pw; pw += blockDim.x) {
    const long long ia = (MODE == 1) ? 2LL * pw : pw, ib = (MODE == 1) ? 2LL * pw + 1 : (long long)n_pw + pw;
    const T xa = r2[ia], xb = r2[ib], ya = r1[ia], yb = r1[ib];  // x: band m2 (bra), y: band m1 (ket)
    const double xar = xa.x, xai = xa.y, xbr = xb.x, xbi = xb.y, yar = ya.x, yai = ya.y, ybr = yb.x, ybi = yb.y;
    acc[0] += xar * ybr + xai * ybi;  acc[1] += xar * ybi - xai * ybr;   // conj(x_a) y_b
    acc[2] += xbr * yar + xbi * yai;  acc[3] += xbr * yai - xbi * yar;   // conj(x_b) y_a
    acc[4] += xar * yar + xai * yai;  acc[5] += xar * yai - xai * yar;   // conj(x_a) y_a
    acc[6] += xbr * ybr + xbi * ybi;  acc[7] += xbr * ybi - xbi * ybr;   // conj(x_b) y_b
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    double v = acc[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_red[warp][q] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t[8];
    for (int q = 0; q < 8; ++q) {
      double v = 0.0;
      for (int w = 0; w < 8; ++w) v += s_red[w][q];
      t[q] = v;
    }
    const double s1 = norms[pr.m1], s2 = norms[pr.m2];
    const double f = (s1 > kEps ? 1.0 / sqrt(fabs(s1)) : 1.0) * (s2 > kEps ? 1.0 / sqrt(fabs(s2)) : 1.0);
    const double sxr = f * (t[0] + t[2]), sxi = f * (t[1] + t[3]);
    const double syr = -f * (t[3] - t[1]), syi = f * (t[2] - t[0]);  // i * (ba - ab)
    const double szr = f * (t[4] - t[6]), szi = f * (t[5] - t[7]);
    const double rr = pr.re, ri = pr.im;
    const double w = (pr.m1 == pr.m2) ? 1.0 : 2.0;
    out[0] = w * (rr * sxr - ri * sxi);
    out[1] = w * (rr * syr - ri * syi);
    out[2] = w * (rr * szr - ri * szi);
  }
}

// K6b -- sigma[cell][i] = sum over the cell's pairs, in pair order (deterministic).
__global__ void __launch_bounds__(128)
k6_pauli_cells(int n_cells, const long long* __restrict__ pair_off, const double* __restrict__ contrib,
               double* __restrict__ sigma) {
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= n_cells) return;
  double a = 0.0, b = 0.0, c = 0.0;
  for (long long p = pair_off[cell]; p < pair_off[cell + 1]; ++p) {
    a += contrib[3 * p];
    b += contrib[3 * p + 1];
    c += contrib[3 * p + 2];
  }
  sigma[3LL * cell] = a;
  sigma[3LL * cell + 1] = b;
  sigma[3LL * cell + 2] = c;
}

}  // namespace

size_t rho_plan_smem(int cap) { return (size_t)cap * 12 + 64 * sizeof(int); }
size_t rho_pairs_smem(int cap) { return (size_t)cap * 12 + ((size_t)cap + 1) * 4 + 64 * sizeof(int); }

int rho_active_cap(int n_bands) {
  // active bands per energy are staged in shared memory: 16 B each (+ row offsets)
  const int hard = 12000;
  return n_bands < hard ? (n_bands < 32 ? 32 : n_bands) : hard;
}

void launch_rho_plan(const double* d_grid, int nener, double sigma, const double* d_band_energies,
                     int n_bands, int n_fold, const double* d_dN, double min_dN, int cap, bool perform_unfold,
                     int* d_pair_count, long long* d_pair_off, int* d_overflow, cudaStream_t s) {
  const int n = n_fold * nener;
  if (n <= 0) return;
  const size_t smem = rho_plan_smem(cap);
  cudaFuncSetAttribute(k4_rho_plan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k4_rho_plan<<<n, 256, smem, s>>>(d_grid, nener, sigma, d_band_energies, n_bands, d_dN, min_dN, cap,
                                   perform_unfold ? 1 : 0, d_pair_count, d_overflow);
  k4_scan<<<1, 1024, 0, s>>>(d_pair_count, n, d_pair_off);
}

void launch_scan(const int* d_in, int n, long long* d_out, cudaStream_t s) {
  k4_scan<<<1, 1024, 0, s>>>(d_in, n, d_out);
}

void launch_rho_gather(const void* d_coeffs, long long stride_bytes, int n_pw, int n_spinor,
                       int n_bands, int layout, int n_fold, const int32_t* d_lists,
                       const long long* d_sel_off, const double* d_norms, void* d_csel, bool dp,
                       cudaStream_t s) {
  if (n_bands <= 0 || n_fold <= 0) return;
  dim3 g((unsigned)n_bands, (unsigned)n_fold);
  const char* c = static_cast<const char*>(d_coeffs);
  float2* o = static_cast<float2*>(d_csel);
  const int mode = n_spinor == 1 ? 0 : layout == BANDUP_GPU_LAYOUT_FORTRAN_INMEM ? 1 : 2;
#define BANDUP_GATHER(M, T) k4_gather<M, T><<<g, 256, 0, s>>>(c, stride_bytes, n_pw, n_spinor, n_bands, d_lists, d_sel_off, d_norms, o)
  if (dp) {
    if (mode == 0) BANDUP_GATHER(0, double2); else if (mode == 1) BANDUP_GATHER(1, double2); else BANDUP_GATHER(2, double2);
  } else {
    if (mode == 0) BANDUP_GATHER(0, float2); else if (mode == 1) BANDUP_GATHER(1, float2); else BANDUP_GATHER(2, float2);
  }
#undef BANDUP_GATHER
}

void launch_rho_pairs(const double* d_grid, int nener, double sigma, const double* d_band_energies,
                      int n_bands, int n_fold, const double* d_dN, int cap, int n_spinor, bool perform_unfold,
                      const long long* d_sel_off, const void* d_csel, const int* d_pair_count,
                      const long long* d_pair_off, RhoPair* d_pairs, cudaStream_t s) {
  const int n = n_fold * nener;
  if (n <= 0) return;
  const size_t smem = rho_pairs_smem(cap);
  cudaFuncSetAttribute(k4_rho_pairs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k4_rho_pairs<<<n, 256, smem, s>>>(d_grid, nener, sigma, d_band_energies, n_bands, d_dN, cap, n_spinor,
                                    perform_unfold ? 1 : 0, d_sel_off, static_cast<const float2*>(d_csel), d_pair_count,
                                    d_pair_off, d_pairs);
}

void launch_orb_contrib(int n_bands, int n_ao, const void* d_dual, const void* d_proj,
                        const uint8_t* d_mask, const double* d_ener, double max_dE, double min_abs,
                        void* d_O, cudaStream_t s) {
  if (n_bands <= 0) return;
  k5_orb_contrib<<<n_bands, 256, 0, s>>>(n_bands, n_ao, static_cast<const double2*>(d_dual),
                                         static_cast<const double2*>(d_proj), d_mask, d_ener, max_dE,
                                         min_abs, static_cast<double2*>(d_O));
}

void launch_pauli(const void* d_coeffs, long long stride_bytes, int n_pw, int layout, const double* d_norms,
                  const RhoPair* d_pairs, long long n_pairs, int n_cells, const long long* d_pair_off,
                  double* d_contrib, double* d_sigma, bool dp, cudaStream_t s) {
  if (n_cells <= 0) return;
  const char* c = static_cast<const char*>(d_coeffs);
  if (n_pairs > 0) {
    const bool inter = layout == BANDUP_GPU_LAYOUT_FORTRAN_INMEM;
    if (dp && inter)
      k6_pauli_pairs<1, double2><<<(unsigned)n_pairs, 256, 0, s>>>(c, stride_bytes, n_pw, d_norms, d_pairs, d_contrib);
    else if (dp)
      k6_pauli_pairs<2, double2><<<(unsigned)n_pairs, 256, 0, s>>>(c, stride_bytes, n_pw, d_norms, d_pairs, d_contrib);
    else if (inter)
      k6_pauli_pairs<1, float2><<<(unsigned)n_pairs, 256, 0, s>>>(c, stride_bytes, n_pw, d_norms, d_pairs, d_contrib);
    else
      k6_pauli_pairs<2, float2><<<(unsigned)n_pairs, 256, 0, s>>>(c, stride_bytes, n_pw, d_norms, d_pairs, d_contrib);
  }
  k6_pauli_cells<<<(n_cells + 127) / 128, 128, 0, s>>>(n_cells, d_pair_off, d_contrib, d_sigma);
}

void launch_trace(int n_cells, int n_bands, const double* d_dN, const long long* d_pair_off,
                  const RhoPair* d_pairs, const void* d_O, double min_abs_rho, int clip, double* d_out,
                  cudaStream_t s) {
  if (n_cells <= 0) return;
  k5_trace<<<(n_cells + 3) / 4, 128, 0, s>>>(n_cells, n_bands, d_dN, d_pair_off, d_pairs,
                                             static_cast<const double2*>(d_O), min_abs_rho, clip, d_out);
}

}  // namespace bandup
