// capi_projected.cu -- C ABI, part 2: what works on an SC-kpt that has been fetched and is still
// resident on the device -- the unfolding-density operator rho (K4), the orbital-contribution
// matrix and N Re Tr(rho O) (K5), spinor Pauli vectors (K6), the Gaussian spectral function.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "ctx.cuh"

namespace bandup {

namespace {

Flight* resident_flight(Ctx* c, int ikpt) {
  for (auto& f : c->flights)
    if (!f.busy && f.resident && f.ikpt == ikpt) return &f;
  return nullptr;
}

}  // namespace

int need_resident(Ctx* c, int ikpt, Flight** out) {
  for (auto& f : c->flights)
    if (f.busy && f.ikpt == ikpt)
      return fail(c, BANDUP_GPU_ERR_INVALID_ARG, "call bandup_gpu_fetch_kpt for this k-point first");
  *out = resident_flight(c, ikpt);
  if (!*out)
    return fail(c, BANDUP_GPU_ERR_UNKNOWN_KPT,
                "k-point " + std::to_string(ikpt) + " is no longer resident on the device");
  return BANDUP_GPU_OK;
}

}  // namespace bandup

using namespace bandup;

extern "C" {

int bandup_gpu_rho_count(int handle, int ikpt, double min_dN, int64_t* n_energies,
                         int64_t* n_entries) {
  Ctx* c = get(handle);
  if (!c) return fail(nullptr, BANDUP_GPU_ERR_BAD_HANDLE, "bad handle");
  if (!(min_dN >= kEpsDp))
    return fail(c, BANDUP_GPU_ERR_DELTA_N_ZERO,
                "ERROR (calc_rho): delta_N = 0. (min_dN must be >= epsilon(1.0_dp); the reference uses 1e-3)");
  Flight* fl = nullptr;
  int rc = need_resident(c, ikpt, &fl);
  if (rc) return rc;
  cudaSetDevice(c->device);
  cudaStream_t s = c->compute;
  const int nu = fl->n_unique, nener = c->nener, nb = fl->n_bands;
  const int cells = nu * nener;
  const int cap = rho_active_cap(nb);
  fl->rho_ready = false;
  CU(fl->pair_count.reserve(sizeof(int) * (size_t)cells));
  CU(fl->pair_off.reserve(sizeof(long long) * ((size_t)cells + 1)));
  CU(fl->overflow.reserve(sizeof(int)));
  CU(fl->sel_off.reserve(sizeof(long long) * ((size_t)nu + 1)));
  CU(cudaMemsetAsync(fl->overflow.p, 0, sizeof(int), s));
  const double* d_grid = static_cast<const double*>(c->d_grid.p);
  const double* d_ener = static_cast<const double*>(fl->energies.p);
  {
    KernelScope k(c, BANDUP_GPU_K_RHO, s);
    // perform_unfolding calls calc_rho WITHOUT std_dev (:1404-1407): lambda is the box
    // (sigma = epsilon) here whatever smearing the delta_N grid was given
    launch_rho_plan(d_grid, nener, kEpsDp, d_ener, nb, nu, fl->d_dn_unique, min_dN, cap, c->perform_unfold,
                    static_cast<int*>(fl->pair_count.p), static_cast<long long*>(fl->pair_off.p),
                    static_cast<int*>(fl->overflow.p), s);
    c->launches++;  // plan + scan
  }
  // ordered selected lists of every unique fold + their offsets
  std::vector<int32_t> h_ns((size_t)nu);
  CU(cudaMemcpyAsync(h_ns.data(), fl->d_ns_unique, sizeof(int32_t) * (size_t)nu, cudaMemcpyDeviceToHost, s));
  fl->h_pair_off.assign((size_t)cells + 1, 0);
  CU(cudaMemcpyAsync(fl->h_pair_off.data(), fl->pair_off.p, sizeof(long long) * ((size_t)cells + 1),
                     cudaMemcpyDeviceToHost, s));
  int h_overflow = 0;
  CU(cudaMemcpyAsync(&h_overflow, fl->overflow.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  if (h_overflow)
    return fail(c, BANDUP_GPU_ERR_CAPACITY,
                "more than " + std::to_string(cap) + " bands couple at one energy (smearing too wide)");
  long long total_sel = 0;
  for (int u = 0; u < nu; ++u) total_sel += h_ns[u];
  const long long total_pairs = fl->h_pair_off[(size_t)cells];
  fl->rho_entries.clear();
  fl->rho_energies = 0;
  for (int cell = 0; cell < cells; ++cell)
    if (fl->h_pair_off[cell + 1] > fl->h_pair_off[cell]) fl->rho_energies++;
  std::vector<RhoPair> h_pairs((size_t)total_pairs);
  if (total_pairs > 0) {
    CU(fl->selected.reserve(sizeof(int32_t) * (size_t)(total_sel > 0 ? total_sel : 1)));
    CU(fl->csel.reserve(sizeof(float2) * (size_t)(total_sel > 0 ? total_sel : 1) * (size_t)nb * (size_t)fl->n_spinor));
    CU(fl->pairs.reserve(sizeof(RhoPair) * (size_t)total_pairs));
    {
      KernelScope k(c, BANDUP_GPU_K_RHO, s);
      launch_scan(fl->d_ns_unique, nu, static_cast<long long*>(fl->sel_off.p), s);
    }
    {
      KernelScope k(c, BANDUP_GPU_K_COSET, s);
      launch_selected_lists(static_cast<const uint8_t*>(fl->slot.p), (size_t)fl->n_pw, fl->n_pw, nu, fl->d_ns_unique,
                            static_cast<int32_t*>(fl->selected.p), s);
    }
    {
      KernelScope k(c, BANDUP_GPU_K_RHO, s);
      launch_rho_gather(fl->coeffs.p, fl->stride_bytes, fl->n_pw, fl->n_spinor, nb, fl->layout, nu,
                        static_cast<const int32_t*>(fl->selected.p),
                        static_cast<const long long*>(fl->sel_off.p),
                        static_cast<const double*>(fl->norms.p), fl->csel.p, fl->dp, s);
    }
    {
      KernelScope k(c, BANDUP_GPU_K_RHO, s);
      launch_rho_pairs(d_grid, nener, kEpsDp, d_ener, nb, nu, fl->d_dn_unique, cap, fl->n_spinor, c->perform_unfold,
                       static_cast<const long long*>(fl->sel_off.p), fl->csel.p,
                       static_cast<const int*>(fl->pair_count.p),
                       static_cast<const long long*>(fl->pair_off.p), static_cast<RhoPair*>(fl->pairs.p), s);
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(h_pairs.data(), fl->pairs.p, sizeof(RhoPair) * (size_t)total_pairs,
                       cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
  }
  // Expand to the list perform_unfolding appends (:1416-1425): full matrix, |rho| >= 1e-5
  // (single-precision abs), m1 outer / m2 inner, for every pc-kpt (folds that share a coset
  // share the operator), energies ascending.
  for (int f = 0; f < fl->n_fold; ++f) {
    const int u = fl->unique_of[f];
    for (int ie = 0; ie < nener; ++ie) {
      const long long p0 = fl->h_pair_off[(size_t)u * nener + ie], p1 = fl->h_pair_off[(size_t)u * nener + ie + 1];
      if (p1 == p0) continue;
      const size_t first = fl->rho_entries.size();
      for (long long p = p0; p < p1; ++p) {
        const RhoPair& r = h_pairs[(size_t)p];
        if (hypotf(r.re, r.im) < 1e-5f) continue;
        fl->rho_entries.push_back({f, ie + 1, r.m1 + 1, r.m2 + 1, r.re, r.im});
        if (r.m1 != r.m2) fl->rho_entries.push_back({f, ie + 1, r.m2 + 1, r.m1 + 1, r.re, -r.im});
      }
      std::sort(fl->rho_entries.begin() + (long)first, fl->rho_entries.end(),
                [](const RhoEntry& a, const RhoEntry& b) { return a.m1 != b.m1 ? a.m1 < b.m1 : a.m2 < b.m2; });
    }
  }
  fl->rho_ready = true;
  if (n_energies) {
    long long ne = 0;
    for (int f = 0; f < fl->n_fold; ++f) {
      const int u = fl->unique_of[f];
      for (int ie = 0; ie < nener; ++ie)
        if (fl->h_pair_off[(size_t)u * nener + ie + 1] > fl->h_pair_off[(size_t)u * nener + ie]) ++ne;
    }
    *n_energies = ne;
  }
  if (n_entries) *n_entries = (int64_t)fl->rho_entries.size();
  return BANDUP_GPU_OK;
}

int bandup_gpu_fetch_rho(int handle, int ikpt, int64_t capacity, int32_t* fold, int32_t* iener,
                         int32_t* m1, int32_t* m2, float* rho_re, float* rho_im) {
  Ctx* c = get(handle);
  if (!c) return fail(nullptr, BANDUP_GPU_ERR_BAD_HANDLE, "bad handle");
  Flight* fl = nullptr;
  int rc = need_resident(c, ikpt, &fl);
  if (rc) return rc;
  if (!fl->rho_ready) return fail(c, BANDUP_GPU_ERR_NOT_CONFIGURED, "call bandup_gpu_rho_count first");
  const size_t n = fl->rho_entries.size();
  if ((int64_t)n > capacity)
    return fail(c, BANDUP_GPU_ERR_CAPACITY, "capacity too small: need " + std::to_string(n));
  for (size_t i = 0; i < n; ++i) {
    const RhoEntry& e = fl->rho_entries[i];
    if (fold) fold[i] = e.fold;
    if (iener) iener[i] = e.iener;
    if (m1) m1[i] = e.m1;
    if (m2) m2[i] = e.m2;
    if (rho_re) rho_re[i] = e.re;
    if (rho_im) rho_im[i] = e.im;
  }
  return BANDUP_GPU_OK;
}

int bandup_gpu_set_perform_unfold(int handle, int perform_unfold) {
  Ctx* c = get(handle);
  if (!c) return fail(nullptr, BANDUP_GPU_ERR_BAD_HANDLE, "bad handle");
  c->perform_unfold = perform_unfold != 0;
  return BANDUP_GPU_OK;
}

int bandup_gpu_spectral_function(int handle, int ikpt, double std_dev, double* sf) {
  Ctx* c = get(handle);
  if (!c) return fail(nullptr, BANDUP_GPU_ERR_BAD_HANDLE, "bad handle");
  Flight* fl = nullptr;
  int rc = need_resident(c, ikpt, &fl);
  if (rc) return rc;
  if (!sf) return fail(c, BANDUP_GPU_ERR_INVALID_ARG, "sf is NULL");
  const double sigma = std_dev > 0.0 ? std_dev : 0.025;  // calc_spectral_function's default (:671)
  cudaSetDevice(c->device);
  cudaStream_t s = c->compute;
  const int nener = c->nener, nu = fl->n_unique;
  CU(fl->trace_out.reserve(sizeof(double) * (size_t)nu * nener));
  {
    KernelScope k(c, BANDUP_GPU_K_DELTA_N, s);
    launch_spectral_function(static_cast<const double*>(c->d_grid.p), nener, sigma,
                             static_cast<const double*>(fl->energies.p), fl->d_w_unique, fl->n_bands, nu,
                             static_cast<double*>(fl->trace_out.p), s);
  }
  CU(cudaGetLastError());
  std::vector<double> h((size_t)nu * nener);
  CU(cudaMemcpyAsync(h.data(), fl->trace_out.p, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  for (int f = 0; f < fl->n_fold; ++f)
    std::memcpy(sf + (size_t)f * nener, h.data() + (size_t)fl->unique_of[f] * nener, sizeof(double) * (size_t)nener);
  return BANDUP_GPU_OK;
}

int bandup_gpu_pauli_vectors(int handle, int ikpt, double* sigma) {
  Ctx* c = get(handle);
  if (!c) return fail(nullptr, BANDUP_GPU_ERR_BAD_HANDLE, "bad handle");
  Flight* fl = nullptr;
  int rc = need_resident(c, ikpt, &fl);
  if (rc) return rc;
  if (!fl->rho_ready) return fail(c, BANDUP_GPU_ERR_NOT_CONFIGURED, "call bandup_gpu_rho_count first");
  if (fl->n_spinor != 2) return fail(c, BANDUP_GPU_ERR_INVALID_ARG, "Pauli vectors need a spinor wavefunction");
  if (!sigma) return fail(c, BANDUP_GPU_ERR_INVALID_ARG, "sigma is NULL");
  cudaSetDevice(c->device);
  cudaStream_t s = c->compute;
  const int nener = c->nener, cells = fl->n_unique * nener;
  const long long n_pairs = fl->h_pair_off.empty() ? 0 : fl->h_pair_off[(size_t)cells];
  DevBuf contrib;
  CU(fl->trace_out.reserve(sizeof(double) * 3 * (size_t)cells));
  cudaError_t e = contrib.reserve(sizeof(double) * 3 * (size_t)(n_pairs > 0 ? n_pairs : 1));
  if (e != cudaSuccess) return cuda_fail(c, e, "pauli scratch");
  {
    KernelScope k(c, BANDUP_GPU_K_RHO, s);
    launch_pauli(fl->coeffs.p, fl->stride_bytes, fl->n_pw, fl->layout, static_cast<const double*>(fl->norms.p),
                 static_cast<const RhoPair*>(fl->pairs.p), n_pairs, cells,
                 static_cast<const long long*>(fl->pair_off.p), static_cast<double*>(contrib.p),
                 static_cast<double*>(fl->trace_out.p), fl->dp, s);
  }
  std::vector<double> h(3 * (size_t)cells);
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(h.data(), fl->trace_out.p, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  contrib.release();
  if (e != cudaSuccess) return cuda_fail(c, e, "bandup_gpu_pauli_vectors");
  for (int f = 0; f < fl->n_fold; ++f)
    std::memcpy(sigma + 3 * (size_t)f * nener, h.data() + 3 * (size_t)fl->unique_of[f] * nener,
                sizeof(double) * 3 * (size_t)nener);
  return BANDUP_GPU_OK;
}

int bandup_gpu_set_operator(int handle, int n_bands, const double* O) {
  Ctx* c = get(handle);
  if (!c) return fail(nullptr, BANDUP_GPU_ERR_BAD_HANDLE, "bad handle");
  if (n_bands < 1 || !O) return fail(c, BANDUP_GPU_ERR_INVALID_ARG, "bad operator");
  cudaSetDevice(c->device);
  const size_t bytes = sizeof(double) * 2 * (size_t)n_bands * (size_t)n_bands;
  CU(cudaStreamSynchronize(c->compute));
  CU(c->d_operator.reserve(bytes));
  CU(cudaMemcpy(c->d_operator.p, O, bytes, cudaMemcpyHostToDevice));
  c->operator_bands = n_bands;
  return BANDUP_GPU_OK;
}

int bandup_gpu_orb_contrib_matrix(int handle, int n_bands, int n_ao, const double* dual,
                                  const double* proj, const uint8_t* ao_mask,
                                  const double* band_energies, double max_band_dE, double min_abs,
                                  double* O_out) {
  Ctx* c = get(handle);
  if (!c) return fail(nullptr, BANDUP_GPU_ERR_BAD_HANDLE, "bad handle");
  if (n_bands < 1 || n_ao < 1 || !dual || !proj || !ao_mask || !band_energies)
    return fail(c, BANDUP_GPU_ERR_INVALID_ARG, "bad orbital-projection arguments");
  cudaSetDevice(c->device);
  cudaStream_t s = c->compute;
  const size_t mat = sizeof(double) * 2 * (size_t)n_bands * (size_t)n_ao;
  const size_t obytes = sizeof(double) * 2 * (size_t)n_bands * (size_t)n_bands;
  // device copies of the projections live in the context: one k-point after another reuses them
  DevBuf &d_dual = c->d_proj_dual, &d_proj = c->d_proj, &d_mask = c->d_ao_mask, &d_ener = c->d_proj_ener;
  cudaError_t e = cudaSuccess;
  if ((e = cudaStreamSynchronize(s)) == cudaSuccess && (e = d_dual.reserve(mat)) == cudaSuccess &&
      (e = d_proj.reserve(mat)) == cudaSuccess && (e = d_mask.reserve((size_t)n_ao)) == cudaSuccess &&
      (e = d_ener.reserve(sizeof(double) * (size_t)n_bands)) == cudaSuccess &&
      (e = c->d_operator.reserve(obytes)) == cudaSuccess) {
    // the two projection matrices are the bulk (C4: 2 x 15 MB) and usually pageable NumPy memory
    int rc = upload_block(c, d_dual.p, dual, mat);
    if (!rc) rc = upload_block(c, d_proj.p, proj, mat);
    if (rc) return rc;
    cudaMemcpyAsync(d_mask.p, ao_mask, (size_t)n_ao, cudaMemcpyHostToDevice, c->copy);
    cudaMemcpyAsync(d_ener.p, band_energies, sizeof(double) * (size_t)n_bands, cudaMemcpyHostToDevice, c->copy);
    if ((e = cudaStreamSynchronize(c->copy)) != cudaSuccess) return cuda_fail(c, e, "uploading the projections");
    {
      KernelScope k(c, BANDUP_GPU_K_RHO, s);
      launch_orb_contrib(n_bands, n_ao, d_dual.p, d_proj.p, static_cast<const uint8_t*>(d_mask.p),
                         static_cast<const double*>(d_ener.p), max_band_dE, min_abs, c->d_operator.p, s);
    }
    e = cudaGetLastError();
    if (e == cudaSuccess && O_out) e = cudaMemcpyAsync(O_out, c->d_operator.p, obytes, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  }
  if (e != cudaSuccess) return cuda_fail(c, e, "bandup_gpu_orb_contrib_matrix");
  c->operator_bands = n_bands;
  return BANDUP_GPU_OK;
}

int bandup_gpu_unfold_operator(int handle, int ikpt, double min_abs_rho, int clip, double* out) {
  Ctx* c = get(handle);
  if (!c) return fail(nullptr, BANDUP_GPU_ERR_BAD_HANDLE, "bad handle");
  Flight* fl = nullptr;
  int rc = need_resident(c, ikpt, &fl);
  if (rc) return rc;
  if (!fl->rho_ready) return fail(c, BANDUP_GPU_ERR_NOT_CONFIGURED, "call bandup_gpu_rho_count first");
  if (!out) return fail(c, BANDUP_GPU_ERR_INVALID_ARG, "out is NULL");
  if (c->operator_bands != fl->n_bands)
    return fail(c, BANDUP_GPU_ERR_INVALID_ARG, "Incompatible matrix shapes! (operator vs n_bands)");
  cudaSetDevice(c->device);
  cudaStream_t s = c->compute;
  const int nener = c->nener, cells = fl->n_unique * nener;
  CU(fl->trace_out.reserve(sizeof(double) * (size_t)cells));
  {
    KernelScope k(c, BANDUP_GPU_K_RHO, s);
    launch_trace(cells, fl->n_bands, fl->d_dn_unique, static_cast<const long long*>(fl->pair_off.p),
                 static_cast<const RhoPair*>(fl->pairs.p), c->d_operator.p, min_abs_rho, clip,
                 static_cast<double*>(fl->trace_out.p), s);
  }
  CU(cudaGetLastError());
  std::vector<double> h((size_t)cells);
  CU(cudaMemcpyAsync(h.data(), fl->trace_out.p, sizeof(double) * (size_t)cells, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  for (int f = 0; f < fl->n_fold; ++f)
    std::memcpy(out + (size_t)f * nener, h.data() + (size_t)fl->unique_of[f] * nener, sizeof(double) * (size_t)nener);
  return BANDUP_GPU_OK;
}

}  // extern "C"
