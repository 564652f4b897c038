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
struct CosetCommon {
  long long adjT[9];   // adjugate of transpose(M), row-major
  long long D;         // |det M|
  unsigned adj32[9];   // adjT mod D in [0, D), valid when D < 2^15 (32-bit fast path of K1)
};

constexpr int kK1MaxBlocks = 128;  // CTAs of K1 per SC-kpt (each leaves one row of counts)

// One SC k-point of a batch: the device-resident descriptor every kernel indexes with
// blockIdx.y (K1, K3) or finds from the global tile index (K2).  "fold" here means a
// DISTINCT coset among the pc-kpts folding onto the SC-kpt (the host replicates rows of
// pc-kpts that share one).
struct KptGeom {
  const char* coeffs;         // first band record
  long long stride_bytes;     // distance between band records
  long long row_elems;        // n_pw * n_spinor coefficients per band
  int n_bands, n_pw, n_spinor, layout;
  int dp;                     // 1: complex128 coefficients (16 bytes per element)
  int nreg;                   // K2 accumulators of this k-point: 0 = shared-memory bins, 1/2/4 = registers
  const int32_t* g_frac;      // [n_pw][3]
  const double* energies;     // [n_bands]
  uint8_t* slot;              // [n_pw], 2-byte aligned: fold id of every plane wave, kNoFold = none
  uint8_t* slot_shift;        // [n_pw], slot_shift[i] = slot[i+1]
  int* block_counts;          // [k1_blocks][n_fold] plane waves per fold seen by each K1 CTA
  int k1_blocks;              // CTAs K1 runs per k-point in this batch (<= kK1MaxBlocks)
  int pad_;
  int32_t* n_selected;        // [n_fold]
  int n_fold;
  int chunk_vecs;             // 16-byte vectors per K2 tile (a multiple of the 1024-vector stage)
  int n_chunks;               // tiles per band
  int n_parts;                // partial rows per band: n_chunks (bins) or n_chunks * 8 warps (registers)
  int perform_unfold;         // 0: -dont_unfold, every spectral weight is 1 (:1346-1351)
  long long tile_begin;       // global index of this k-point's first K2 tile
  double* partials;           // [n_bands][n_parts][n_fold+1]; [..][0] = norm part
  double* weights;            // [n_fold][n_bands]
  double* norms;              // [n_bands] or null
  double* dN;                 // [n_fold][nener]
  int fold_key[kMaxFolds][3]; // key(g0) of every fold
};

// K1 -- coset classification (replaces the vec_in_latt sweep of
// select_coeffs_to_calc_spectral_weights, band_unfolding_routines_mod.f90:586-600).
// slot[pw] = index of the fold whose coset contains G(pw), else kNoFold; slot_shift lets K2
// fetch the two slot ids of a 16-byte vector with one aligned 16-bit load whatever the row's
// alignment.  Per-CTA fold counts go to block_counts (summed by K3: no atomics, no memset).
// CTAs per k-point for a batch whose largest k-point has max_n_pw plane waves (four per thread)
int coset_classify_blocks(int max_n_pw);
void launch_coset_classify(const KptGeom* d_geoms, int n_kpts, int max_n_pw, const CosetCommon& cc,
                           cudaStream_t s);

// Rows of pc-kpts that share a coset: dst row f <- unique row unique_of[f] (weights, delta_N,
// n_selected) for every listed SC-kpt in one launch.
struct ReplItem {
  double* w_dst;
  const double* w_src;
  double* dn_dst;
  const double* dn_src;
  int32_t* ns_dst;
  const int32_t* ns_src;
  int uo_off;  // unique_of[n_fold] starts this many ints behind the end of the item array
  int n_fold, n_bands, nener;
};
void launch_replicate_rows(const ReplItem* d_items, int n_items, int max_n_fold, cudaStream_t s);

// Ordered compaction of the 1-based selected indices of every fold (CSR).  d_slot: one table per
// pass of kMaxFolds folds, slot_pitch bytes apart.
void launch_selected_lists(const uint8_t* d_slot, size_t slot_pitch, int n_pw, int n_fold,
                           const int32_t* d_n_selected, int32_t* d_selected,
                           cudaStream_t s);

// Launch with or without programmatic stream serialisation (PDL): with it the kernel may start
// while the previous kernel of the stream still runs and must call griddepcontrol.wait before it
// touches anything that kernel writes.  K1 -> K2 -> K3 of a batch are chained this way.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                 bool pdl, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

int weights_num_chunks(long long row_elems, int chunk_vecs);
// K2 accumulator layout of a k-point: 0 = shared-memory bins, 1/2/4 = that many register
// accumulators (variant: 0 automatic by n_fold / |det M|, 1 bins always, 2 registers whenever n_fold <= 4)
int weights_pick_nreg(int variant, int n_fold, long long det_m);
int weights_partials_sub(int nreg);
size_t weights_partials_bytes(int n_bands, int n_chunks, int n_fold);
// requested > 0: rounded up to whole stages; 0: chosen from the block shape.  row_elems here and
// in weights_num_chunks counts 8-byte units (complex64 elements; a complex128 element counts twice)
int weights_pick_chunk_vecs(int requested, long long row_elems, int n_bands, int sm_count, int n_spinor);
int weights_max_folds_supported();

// K2 -- the roofline kernel: one pass over every complex64 coefficient of every k-point of
// the batch; per (band, chunk) tile the band-norm part and the per-fold masked |C|^2 sums.
// Fuses io_routines_mod.f90:1317-1330 (norm) with band_unfolding_routines_mod.f90:638-644 and
// :1337-1345 (weights, spinor sum).  All k-points of a batch share n_spinor and layout.
// bins_rows: 1 + the largest n_fold among the k-points with nreg == 0 (0 when there is none)
void launch_weights_reduce(const KptGeom* d_geoms, int n_kpts, long long total_tiles, int bins_rows,
                           int n_spinor, int layout, bool dp, int ctas_per_sm, int n_stages, int sm_count,
                           bool pdl, cudaStream_t s);

// K3 -- first the deterministic tile combine of K2's partials (s = sum_c norm part,
// S_f = sum_c fold part, weights[f][b] = s > eps ? S_f / s : S_f -- the reference leaves a band
// with s <= epsilon unscaled, io_routines_mod.f90:1325 -- and n_selected = sum of K1's CTA
// counts), then delta_N(k, E_i) = sum_m P_m * lambda(E_i; E_m, delta_e, sigma)
// (calc_delta_N_pckpt, band_unfolding_routines_mod.f90:719-757, with lambda :701-716
// and integral_delta_x_minus_x0, math_mod.f90:181-193).
void launch_delta_n(const KptGeom* d_geoms, int n_kpts, int max_n_fold, const double* d_grid,
                    int nener, double sigma, int sm_count, bool pdl, cudaStream_t s);

// K0 -- plane-wave list of an SC k-point regenerated on the device in the WAVECAR reader's order
// (read_wavecar_mod.f90:213-304).  count: per-CTA counts + exclusive scan (total at
// d_block_off[gvec_num_blocks]); write: ordered compaction into d_g_frac [n][3].
long long gvec_num_blocks(const int* nb);
void launch_gvec_count(const double* kfrac, const double* B_rows, double c, double ecut, const int* nb,
                       int* d_block_count, long long* d_block_off, cudaStream_t s);
void launch_gvec_write(const double* kfrac, const double* B_rows, double c, double ecut, const int* nb,
                       const long long* d_block_off, int32_t* d_g_frac, long long capacity, cudaStream_t s);

// K7 -- geometric unfolding relations: first (SC-kpt, equivalent point) every pc-kpt folds onto,
// with the reference's floating-point vec_in_latt (band_unfolding_routines_mod.f90:306-429).
void launch_fold_map(const double* latt, const double* aux, double tol, int n_pck, const double* d_pck, int n_sck,
                     int n_skip, const int* d_eqv_off, const double* d_eqv, int* d_fold_sck, int* d_fold_eqv,
                     int* d_n_folding, cudaStream_t s);

// A(k,E) with a Gaussian delta (calc_spectral_function, :654-698); weights per distinct fold.
void launch_spectral_function(const double* d_grid, int nener, double stddev, const double* d_band_energies,
                              const double* d_weights, int n_bands, int n_fold, double* d_sf,
                              cudaStream_t s);

// ---- projected variant (rho.cu) ----------------------------------------------
struct RhoPair {  // one upper-triangle element of an unfolding-density operator
  int m1, m2;     // 0-based band indices, m1 <= m2
  float re, im;
};

int rho_active_cap(int n_bands);
// K4a + scan: pair_count[f*nener + ie], pair_off = exclusive prefix (n+1 entries).
void launch_rho_plan(const double* d_grid, int nener, double sigma, const double* d_band_energies,
                     int n_bands, int n_fold, const double* d_dN, double min_dN, int cap, bool perform_unfold,
                     int* d_pair_count, long long* d_pair_off, int* d_overflow, cudaStream_t s);
void launch_scan(const int* d_in, int n, long long* d_out, cudaStream_t s);
void launch_rho_gather(const void* d_coeffs, long long stride_bytes, int n_pw, int n_spinor,
                       int n_bands, int layout, int n_fold, const int32_t* d_lists,
                       const long long* d_sel_off, const double* d_norms, void* d_csel, bool dp,
                       cudaStream_t s);
void launch_rho_pairs(const double* d_grid, int nener, double sigma, const double* d_band_energies,
                      int n_bands, int n_fold, const double* d_dN, int cap, int n_spinor, bool perform_unfold,
                      const long long* d_sel_off, const void* d_csel, const int* d_pair_count,
                      const long long* d_pair_off, RhoPair* d_pairs, cudaStream_t s);
void launch_orb_contrib(int n_bands, int n_ao, const void* d_dual, const void* d_proj,
                        const uint8_t* d_mask, const double* d_ener, double max_dE, double min_abs,
                        void* d_O, cudaStream_t s);
// K6: expectation values of the Pauli matrices, Re Tr(rho . sigma_i), spinor wavefunctions only
void launch_pauli(const void* d_coeffs, long long stride_bytes, int n_pw, int layout, const double* d_norms,
                  const RhoPair* d_pairs, long long n_pairs, int n_cells, const long long* d_pair_off,
                  double* d_contrib, double* d_sigma, bool dp, cudaStream_t s);
void launch_trace(int n_cells, int n_bands, const double* d_dN, const long long* d_pair_off,
                  const RhoPair* d_pairs, const void* d_O, double min_abs_rho, int clip, double* d_out,
                  cudaStream_t s);

// ---- symmetry averaging over the needed directions (symm_average.cu) ----------
// K8: out[r] = sum_d w[d] * in[d][r], direction order, dp (get_delta_Ns_for_output :789-855, :975-1043)
void launch_symm_average(const double* d_in, const double* d_w, int n_dirs, long long n_values, double* d_out,
                         int sm_count, cudaStream_t s);
// K9: symmetry-averaged unfolding-density operators (:857-973); plan = keys + first flags + scan
// (d_scan has n+1 entries, d_scan[n] = number of averaged elements), merge = slots + values.
void launch_symm_average_rho_plan(const int32_t* d_pck, const int32_t* d_ien, const int32_t* d_m1,
                                  const int32_t* d_m2, long long n, const long long* d_dir_off, int n_dirs,
                                  int n_pckpts, int nener, int n_bands, const double* d_avg_dN, double min_dN,
                                  long long* d_key, int* d_flag, long long* d_scan, int* d_err, cudaStream_t s);
void launch_symm_average_rho_merge(const long long* d_key, long long n, const long long* d_dir_off, int n_dirs,
                                   int nener, int n_bands, const double* d_w, const float* d_re,
                                   const float* d_im, const int* d_flag, const long long* d_scan,
                                   int32_t* o_pck, int32_t* o_ien, int32_t* o_m1, int32_t* o_m2, float* o_re,
                                   float* o_im, cudaStream_t s);

void launch_synth_fill(void* d_coeffs, long long row_elems, int n_bands,
                       long long stride_bytes, uint64_t seed, uint64_t ikpt,
                       cudaStream_t s);

}  // namespace bandup
