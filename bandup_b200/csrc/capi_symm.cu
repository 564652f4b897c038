// capi_symm.cu -- C ABI, part 3: symmetry averaging over the needed directions (K8/K9), the step
// between the SC-kpt loop and the output writers (get_delta_Ns_for_output,
// band_unfolding_routines_mod.f90:789-1043).
#include <cmath>

#include "ctx.cuh"

using namespace bandup;

extern "C" {

int bandup_gpu_symm_average(int handle, int n_dirs, int64_t n_values, const double* weights, const double* in,
                            double* out) {
  Ctx* c = get(handle);
  if (!c) return fail(nullptr, BANDUP_GPU_ERR_BAD_HANDLE, "bad handle");
  if (n_dirs < 1 || n_values < 0 || !weights || (n_values > 0 && (!in || !out)))
    return fail(c, BANDUP_GPU_ERR_INVALID_ARG, "symm_average: n_dirs >= 1, non-null arrays");
  if (n_values == 0) return BANDUP_GPU_OK;
  cudaSetDevice(c->device);
  cudaStream_t s = c->compute;
  const size_t in_bytes = sizeof(double) * (size_t)n_dirs * (size_t)n_values;
  const size_t w_off = align256(in_bytes);
  CU(c->avg_in.reserve(w_off + sizeof(double) * (size_t)n_dirs));
  CU(c->avg_out.reserve(sizeof(double) * (size_t)n_values));
  char* base = static_cast<char*>(c->avg_in.p);
  // the columns are usually pageable host memory: through the pinned staging ring (copy stream)
  CU(cudaStreamSynchronize(s));
  int rc = upload_block(c, base, in, in_bytes);
  if (rc) return rc;
  CU(cudaMemcpyAsync(base + w_off, weights, sizeof(double) * (size_t)n_dirs, cudaMemcpyHostToDevice, c->copy));
  CU(cudaStreamSynchronize(c->copy));
  {
    KernelScope k(c, BANDUP_GPU_K_SYMM_AVG, s);
    launch_symm_average(reinterpret_cast<const double*>(base), reinterpret_cast<const double*>(base + w_off),
                        n_dirs, n_values, static_cast<double*>(c->avg_out.p), c->sm_count, s);
  }
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out, c->avg_out.p, sizeof(double) * (size_t)n_values, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return BANDUP_GPU_OK;
}

int bandup_gpu_symm_average_rho(int handle, int n_dirs, int n_pckpts, int nener, int n_bands,
                                const double* weights, const double* avg_dN, double min_dN,
                                const int64_t* dir_offsets, const int32_t* pckpt, const int32_t* iener,
                                const int32_t* m1, const int32_t* m2, const float* rho_re, const float* rho_im,
                                int64_t capacity, int64_t* n_out, int32_t* pckpt_out, int32_t* iener_out,
                                int32_t* m1_out, int32_t* m2_out, float* rho_re_out, float* rho_im_out) {
  Ctx* c = get(handle);
  if (!c) return fail(nullptr, BANDUP_GPU_ERR_BAD_HANDLE, "bad handle");
  if (n_dirs < 1 || n_pckpts < 1 || nener < 1 || n_bands < 1 || !weights || !avg_dN || !dir_offsets || !n_out)
    return fail(c, BANDUP_GPU_ERR_INVALID_ARG, "symm_average_rho: bad sizes or null arrays");
  if (dir_offsets[0] != 0) return fail(c, BANDUP_GPU_ERR_INVALID_ARG, "symm_average_rho: dir_offsets[0] must be 0");
  for (int d = 0; d < n_dirs; ++d)
    if (dir_offsets[d + 1] < dir_offsets[d])
      return fail(c, BANDUP_GPU_ERR_INVALID_ARG, "symm_average_rho: dir_offsets must not decrease");
  const int64_t n = dir_offsets[n_dirs];
  if (n >= (int64_t(1) << 31)) return fail(c, BANDUP_GPU_ERR_INVALID_ARG, "symm_average_rho: too many entries");
  if ((double)n_pckpts * nener * (double)n_bands * n_bands >= 9.0e18)
    return fail(c, BANDUP_GPU_ERR_INVALID_ARG, "symm_average_rho: key space exceeds 63 bits");
  *n_out = 0;
  if (n == 0) return BANDUP_GPU_OK;
  if (!pckpt || !iener || !m1 || !m2 || !rho_re || !rho_im)
    return fail(c, BANDUP_GPU_ERR_INVALID_ARG, "symm_average_rho: null entry arrays");
  cudaSetDevice(c->device);
  cudaStream_t s = c->compute;
  const size_t N = (size_t)n, groups = (size_t)n_pckpts * (size_t)nener;
  // staging: [pckpt | iener | m1 | m2 | re | im] (4 B each), dir_off, weights, avg_dN
  size_t off = 0;
  auto take = [&off](size_t bytes) { size_t o = off; off = align256(off + bytes); return o; };
  const size_t o_pck = take(4 * N), o_ien = take(4 * N), o_m1 = take(4 * N), o_m2 = take(4 * N),
               o_re = take(4 * N), o_im = take(4 * N), o_off = take(8 * ((size_t)n_dirs + 1)),
               o_w = take(8 * (size_t)n_dirs), o_dn = take(8 * groups);
  CU(c->avg_in.reserve(off));
  char* in = static_cast<char*>(c->avg_in.p);
  off = 0;
  const size_t t_key = take(8 * N), t_flag = take(4 * N), t_scan = take(8 * (N + 1)), t_err = take(4);
  CU(c->avg_tmp.reserve(off));
  char* tmp = static_cast<char*>(c->avg_tmp.p);
  const void* srcs[6] = {pckpt, iener, m1, m2, rho_re, rho_im};
  const size_t dsts[6] = {o_pck, o_ien, o_m1, o_m2, o_re, o_im};
  CU(cudaStreamSynchronize(s));
  for (int a = 0; a < 6; ++a) {  // pageable arrays: through the pinned staging ring (copy stream)
    const int rc = upload_block(c, in + dsts[a], srcs[a], 4 * N);
    if (rc) return rc;
  }
  CU(cudaStreamSynchronize(c->copy));
  static_assert(sizeof(long long) == sizeof(int64_t), "dir_offsets are staged as long long");
  CU(cudaMemcpyAsync(in + o_off, dir_offsets, 8 * ((size_t)n_dirs + 1), cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(in + o_w, weights, 8 * (size_t)n_dirs, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(in + o_dn, avg_dN, 8 * groups, cudaMemcpyHostToDevice, s));
  CU(cudaMemsetAsync(tmp + t_err, 0, 4, s));
  long long* d_key = reinterpret_cast<long long*>(tmp + t_key);
  int* d_flag = reinterpret_cast<int*>(tmp + t_flag);
  long long* d_scan = reinterpret_cast<long long*>(tmp + t_scan);
  const long long* d_off = reinterpret_cast<const long long*>(in + o_off);
  {
    KernelScope k(c, BANDUP_GPU_K_SYMM_AVG, s);
    launch_symm_average_rho_plan(reinterpret_cast<const int32_t*>(in + o_pck), reinterpret_cast<const int32_t*>(in + o_ien),
                                 reinterpret_cast<const int32_t*>(in + o_m1), reinterpret_cast<const int32_t*>(in + o_m2),
                                 n, d_off, n_dirs, n_pckpts, nener, n_bands,
                                 reinterpret_cast<const double*>(in + o_dn), min_dN, d_key, d_flag, d_scan,
                                 reinterpret_cast<int*>(tmp + t_err), s);
    c->launches += 2;  // keys, first, scan
  }
  CU(cudaGetLastError());
  long long h_total = 0;
  int h_err = 0;
  CU(cudaMemcpyAsync(&h_total, d_scan + N, sizeof(long long), cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(&h_err, tmp + t_err, sizeof(int), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  if (h_err & 1) return fail(c, BANDUP_GPU_ERR_INVALID_ARG, "symm_average_rho: pckpt / iener / band index out of range");
  if (h_err & 2)
    return fail(c, BANDUP_GPU_ERR_INVALID_ARG,
                "symm_average_rho: a direction's entries are not in (pckpt, iener, m1, m2) order or repeat an element");
  *n_out = h_total;
  if (h_total == 0) return BANDUP_GPU_OK;
  if (h_total > capacity)
    return fail(c, BANDUP_GPU_ERR_INVALID_ARG,
                "symm_average_rho: capacity " + std::to_string(capacity) + " < " + std::to_string(h_total) + " averaged entries");
  if (!pckpt_out || !iener_out || !m1_out || !m2_out || !rho_re_out || !rho_im_out)
    return fail(c, BANDUP_GPU_ERR_INVALID_ARG, "symm_average_rho: null output arrays");
  const size_t T = (size_t)h_total;
  off = 0;
  const size_t r_pck = take(4 * T), r_ien = take(4 * T), r_m1 = take(4 * T), r_m2 = take(4 * T),
               r_re = take(4 * T), r_im = take(4 * T);
  CU(c->avg_out.reserve(off));
  char* o = static_cast<char*>(c->avg_out.p);
  {
    KernelScope k(c, BANDUP_GPU_K_SYMM_AVG, s);
    launch_symm_average_rho_merge(d_key, n, d_off, n_dirs, nener, n_bands, reinterpret_cast<const double*>(in + o_w),
                                  reinterpret_cast<const float*>(in + o_re), reinterpret_cast<const float*>(in + o_im),
                                  d_flag, d_scan, reinterpret_cast<int32_t*>(o + r_pck),
                                  reinterpret_cast<int32_t*>(o + r_ien), reinterpret_cast<int32_t*>(o + r_m1),
                                  reinterpret_cast<int32_t*>(o + r_m2), reinterpret_cast<float*>(o + r_re),
                                  reinterpret_cast<float*>(o + r_im), s);
  }
  CU(cudaGetLastError());
  void* outs[6] = {pckpt_out, iener_out, m1_out, m2_out, rho_re_out, rho_im_out};
  const size_t rs[6] = {r_pck, r_ien, r_m1, r_m2, r_re, r_im};
  for (int a = 0; a < 6; ++a) CU(cudaMemcpyAsync(outs[a], o + rs[a], 4 * T, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return BANDUP_GPU_OK;
}

}  // extern "C"
