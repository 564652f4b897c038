// fold_map.cu -- K7: geometric unfolding relations, "which SC-kpt does each pc-kpt fold onto".
//
// Replaces the search loop of get_GUR_not_public (reference
// band_unfolding_routines_mod.f90:306-429): for every pc-kpt, the SC-kpts of the list are
// visited in order (the first n_sckpts_to_skip ignored) and, for each, the points equivalent to
// it under the SC's symmetry operations; the pc-kpt folds onto the first one for which
//     vec_in_latt(pc_kpt - eqv_point, B_matrix_SC, tol)        (crystals_mod.f90:204-255)
// holds.  The symmetry-equivalent points themselves come from the reference's spglib-based
// machinery (symmetry_mod.f90:311-484), which stays on the host; without symmetry each SC-kpt is
// its own only equivalent.  vec_in_latt is evaluated with the reference's floating-point recipe
// (coordinates through inverse(transpose(latt)), mod(.,1), comparison with the 27 combinations
// {0,1,-1}^3 . latt within tol per component and in norm, math_mod.f90:234-282), explicitly
// rounded, so that the map equals the CPU one decision for decision.  One thread per pc-kpt.
#include "kernels.cuh"

namespace bandup {

namespace {

struct FoldMapParams {
  double latt[9];  // B_matrix_SC, row i = vector i
  double aux[9];   // inverse(transpose(latt))
  double tol;
};

__device__ bool vec_in_latt_dev(const FoldMapParams& p, const double* v) {
  double r[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double f = __dadd_rn(__dadd_rn(__dmul_rn(p.aux[3 * i], v[0]), __dmul_rn(p.aux[3 * i + 1], v[1])),
                               __dmul_rn(p.aux[3 * i + 2], v[2]));
    r[i] = fmod(f, 1.0);  // Fortran mod(): sign of the dividend
  }
  double red[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) red[j] = __dadd_rn(red[j], __dmul_rn(r[i], p.latt[3 * i + j]));
  const double seq[3] = {0.0, 1.0, -1.0};
  for (int a3 = 0; a3 < 3; ++a3)
    for (int a2 = 0; a2 < 3; ++a2)
      for (int a1 = 0; a1 < 3; ++a1) {
        bool ok = true;
        double d2 = 0.0;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const double g = __dadd_rn(__dadd_rn(__dmul_rn(seq[a1], p.latt[j]), __dmul_rn(seq[a2], p.latt[3 + j])),
                                     __dmul_rn(seq[a3], p.latt[6 + j]));
          const double d = __dsub_rn(g, red[j]);
          if (fabs(d) > p.tol) ok = false;
          d2 = __dadd_rn(d2, __dmul_rn(d, d));
        }
        if (ok && !(__dsqrt_rn(d2) > p.tol)) return true;
      }
  return false;
}

__global__ void __launch_bounds__(128)
k7_fold_map(FoldMapParams p, int n_pck, const double* __restrict__ pck, int n_sck, int n_skip,
            const int* __restrict__ eqv_off, const double* __restrict__ eqv, int* __restrict__ fold_sck,
            int* __restrict__ fold_eqv, int* __restrict__ n_folding) {
  const int ik = blockIdx.x * blockDim.x + threadIdx.x;
  if (ik >= n_pck) return;
  const double k[3] = {pck[3 * ik], pck[3 * ik + 1], pck[3 * ik + 2]};
  int got_s = -1, got_e = -1;
  for (int s = n_skip; s < n_sck && got_s < 0; ++s)
    for (int e = eqv_off[s]; e < eqv_off[s + 1]; ++e) {
      const double v[3] = {__dsub_rn(k[0], eqv[3 * e]), __dsub_rn(k[1], eqv[3 * e + 1]), __dsub_rn(k[2], eqv[3 * e + 2])};
      if (vec_in_latt_dev(p, v)) {
        got_s = s;
        got_e = e;
        break;
      }
    }
  fold_sck[ik] = got_s;
  fold_eqv[ik] = got_e;
  if (got_s >= 0) atomicAdd(n_folding, 1);  // an integer count: order-free
}

}  // namespace

void launch_fold_map(const double* latt, const double* aux, double tol, int n_pck, const double* d_pck, int n_sck,
                     int n_skip, const int* d_eqv_off, const double* d_eqv, int* d_fold_sck, int* d_fold_eqv,
                     int* d_n_folding, cudaStream_t s) {
  if (n_pck <= 0) return;
  FoldMapParams p;
  for (int i = 0; i < 9; ++i) {
    p.latt[i] = latt[i];
    p.aux[i] = aux[i];
  }
  p.tol = tol;
  cudaMemsetAsync(d_n_folding, 0, sizeof(int), s);
  k7_fold_map<<<(n_pck + 127) / 128, 128, 0, s>>>(p, n_pck, d_pck, n_sck, n_skip, d_eqv_off, d_eqv, d_fold_sck,
                                                  d_fold_eqv, d_n_folding);
}

}  // namespace bandup
