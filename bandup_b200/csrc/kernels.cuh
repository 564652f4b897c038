// kernels.cuh -- launch interfaces of the sm_100a kernels behind the C ABI.
// Everything here is internal; the public surface is include/bandup_gpu.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bandup_gpu.h"

namespace bandup {

constexpr int kMaxFolds = BANDUP_GPU_MAX_FOLDS;
constexpr uint8_t kNoFold = 255;  // slot id of a plane wave no requested pc-kpt needs

// Integer data of the SC<->pc relation, by value in kernel parameters.
// key(g) = (g . adjT) mod D component-wise; two SC reciprocal-lattice vectors
// differ by a pc reciprocal-lattice vector iff their keys are equal
// (pc vectors in SC Miller coordinates are the rows of transpose(M), so
// g = n.M^T has integer n iff g.adj(M^T) = 0 mod det M).
struct CosetParams {
  long long adjT[9];        // adjugate of transpose(M), row-major
  long long D;              // |det M|
  int n_fold;
  int fold_key[kMaxFolds][3];  // key(g0) of every requested pc-kpt
};

// K1 -- coset classification (replaces the vec_in_latt sweep of
// select_coeffs_to_calc_spectral_weights, band_unfolding_routines_mod.f90:586-600).
// slot[pw] = index of the first fold whose coset contains G(pw), else kNoFold;
// n_selected[f] = number of plane waves of fold f (must be zeroed by the caller,
// launch_coset_classify does it).
void launch_coset_classify(const int32_t* d_g_frac, int n_pw, const CosetParams& cp,
                           uint8_t* d_slot, int32_t* d_n_selected, cudaStream_t s);

// Ordered compaction of the 1-based selected indices of every fold (CSR).
// d_offsets[f] = exclusive prefix of n_selected, computed on device.
void launch_selected_lists(const uint8_t* d_slot, int n_pw, int n_fold,
                           const int32_t* d_n_selected, int32_t* d_selected,
                           cudaStream_t s);

struct WeightsGeom {
  const char* coeffs;      // device, first band record
  long long stride_bytes;  // distance between band records
  long long row_elems;     // n_pw * n_spinor complex64 per band
  int n_bands, n_pw, n_spinor, layout;
  const uint8_t* slot;     // [n_pw]
  int n_fold;
  int chunk_vecs;          // 16-byte vectors per tile
  int n_chunks;            // tiles per band
  double* partials;        // [n_bands][n_chunks][n_fold+1]; [..][0] = norm part
};

int weights_num_chunks(long long row_elems, int chunk_vecs);
size_t weights_partials_bytes(int n_bands, int n_chunks, int n_fold);
int weights_default_chunk_vecs();

// K2 -- the roofline kernel: one pass over every complex64 coefficient of the
// k-point; per (band, chunk) tile the band-norm part and the per-fold masked
// |C|^2 sums.  Fuses io_routines_mod.f90:1317-1330 (norm) with
// band_unfolding_routines_mod.f90:638-644 and :1337-1345 (weights, spinor sum).
void launch_weights_reduce(const WeightsGeom& g, int ctas_per_sm, int sm_count,
                           cudaStream_t s);

// K2b -- deterministic tile combine: s = sum_c norm part, S_f = sum_c fold part,
// weights[f][b] = s > eps ? S_f / s : S_f (the reference leaves a band with
// s <= epsilon unscaled, io_routines_mod.f90:1325).
void launch_weights_finalize(const double* d_partials, int n_bands, int n_chunks,
                             int n_fold, double* d_weights, double* d_norms,
                             cudaStream_t s);

// K3 -- delta_N(k, E_i) = sum_m P_m * lambda(E_i; E_m, delta_e, sigma)
// (calc_delta_N_pckpt, band_unfolding_routines_mod.f90:719-757, with lambda :701-716
// and integral_delta_x_minus_x0, math_mod.f90:181-193).
void launch_delta_n(const double* d_grid, int nener, double sigma,
                    const double* d_band_energies, const double* d_weights,
                    int n_bands, int n_fold, double* d_dN, cudaStream_t s);

void launch_synth_fill(void* d_coeffs, long long row_elems, int n_bands,
                       long long stride_bytes, uint64_t seed, uint64_t ikpt,
                       cudaStream_t s);

}  // namespace bandup
