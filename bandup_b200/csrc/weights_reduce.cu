// weights_reduce.cu -- K2: fused band-norm + masked |C|^2 segmented reduction.
//
// The one kernel of the path that touches every plane-wave coefficient.  It
// replaces three reference loops with a single streaming pass:
//   * the band renormalisation of read_wavefunction
//     (io_routines_mod.f90:1317-1330): s_m = sum_{spinor,G} |C_m(G)|^2
//   * spectral_weight_for_coeff (band_unfolding_routines_mod.f90:638-644) and its
//     spinor sum (:1337-1345): S_{m,f} = sum_{G in coset f} |C_m(G)|^2
// for ALL pc-kpts f folding onto the SC-kpt at once.  P = S/s is formed by K2b.
//
// Roofline: HBM.  Algorithmic bytes = n_bands * n_pw * n_spinor * 8 (each
// complex64 read exactly once) + the one-byte slot table (L2-resident, re-read
// per band row from L1/L2, not HBM).  Arithmetic intensity ~0.5 flop/B: no
// tensor-core work here.
//
// Shape: persistent grid (sm_count * ctas_per_sm CTAs) walking (chunk, band)
// tiles chunk-major so the CTAs resident at one time share the same stretch of
// the slot table.  Per tile each thread issues U independent 128-bit streaming
// loads (ld.global.nc.L1::no_allocate) before consuming any of them.  |C|^2 is
// formed in fp32 (as the reference's abs() is), the norm is accumulated in
// short fp32 runs promoted to fp64, selected plane waves go straight to
// per-thread fp64 bins in shared memory (conflict-free: bin f of thread t lives
// at bins[f*THREADS + t]); warp shuffles + one shared-memory hop reduce the tile.
// Fully deterministic: no atomics, fixed combination order.
#include "kernels.cuh"

namespace bandup {

namespace {

constexpr int kThreads = 256;
constexpr int kUnroll = 8;  // 128-bit loads in flight per thread

__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// plane-wave index of row element e
template <int MODE>  // 0: one component; 1: [pw][spinor] interleaved; 2: [spinor][pw] blocked
__device__ __forceinline__ int pw_of(long long e, int n_pw) {
  if (MODE == 0) return (int)e;
  if (MODE == 1) return (int)(e >> 1);
  return (int)(e >= n_pw ? e - n_pw : e);
}

template <int MODE>
__global__ void __launch_bounds__(kThreads)
k2_weights_reduce(WeightsGeom g) {
  extern __shared__ double smem[];
  double* bins = smem;                                  // [n_fold][kThreads]
  double* wpart = smem + (size_t)g.n_fold * kThreads;   // [n_fold+1][kThreads/32]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nslots = g.n_fold + 1;
  const long long total_tiles = (long long)g.n_bands * g.n_chunks;
  const uint8_t* __restrict__ slot = g.slot;

  for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int chunk = (int)(tile / g.n_bands);
    const int band = (int)(tile - (long long)chunk * g.n_bands);
    const char* row = g.coeffs + (long long)band * g.stride_bytes;
    // first row element whose address is 16-byte aligned (rows are 8-byte aligned)
    const int e_al = (int)((reinterpret_cast<unsigned long long>(row) >> 3) & 1ull);
    const long long nvb = (g.row_elems - e_al) >> 1;  // whole 16-byte vectors in the row
    const float4* vec = reinterpret_cast<const float4*>(row + 8 * e_al);
    const long long j0 = (long long)chunk * g.chunk_vecs;
    long long j1 = j0 + g.chunk_vecs;
    if (j1 > nvb) j1 = nvb;

    for (int f = 0; f < g.n_fold; ++f) bins[f * kThreads + tid] = 0.0;
    double nrm = 0.0;

    for (long long jb = j0 + tid; jb < j1; jb += (long long)kThreads * kUnroll) {
      float4 v[kUnroll];
      uint8_t sa[kUnroll], sb[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const long long j = jb + (long long)u * kThreads;
        if (j < j1) {
          v[u] = ld_stream_f4(vec + j);
          const long long e = e_al + 2 * j;
          sa[u] = __ldg(slot + pw_of<MODE>(e, g.n_pw));
          sb[u] = (MODE == 1 && e_al == 0) ? sa[u] : __ldg(slot + pw_of<MODE>(e + 1, g.n_pw));
        } else {
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          sa[u] = kNoFold;
          sb[u] = kNoFold;
        }
      }
      float run = 0.f;
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const float pa = fmaf(v[u].y, v[u].y, v[u].x * v[u].x);
        const float pb = fmaf(v[u].w, v[u].w, v[u].z * v[u].z);
        run += pa;
        run += pb;
        if (sa[u] < g.n_fold) bins[sa[u] * kThreads + tid] += (double)pa;
        if (sb[u] < g.n_fold) bins[sb[u] * kThreads + tid] += (double)pb;
      }
      nrm += (double)run;
    }

    // row head / tail elements that do not fill an aligned vector
    if (tid == 0) {
      const float2* el = reinterpret_cast<const float2*>(row);
      if (chunk == 0 && e_al == 1 && g.row_elems > 0) {
        const float2 c = el[0];
        const float p = fmaf(c.y, c.y, c.x * c.x);
        nrm += (double)p;
        const uint8_t s = slot[pw_of<MODE>(0, g.n_pw)];
        if (s < g.n_fold) bins[s * kThreads] += (double)p;
      }
      const long long e_tail = e_al + 2 * nvb;
      if (chunk == g.n_chunks - 1 && e_tail < g.row_elems) {
        const float2 c = el[e_tail];
        const float p = fmaf(c.y, c.y, c.x * c.x);
        nrm += (double)p;
        const uint8_t s = slot[pw_of<MODE>(e_tail, g.n_pw)];
        if (s < g.n_fold) bins[s * kThreads] += (double)p;
      }
    }

    // tile reduction: warp shuffles, then the warps' partials through smem
    {
      const double w = warp_sum(nrm);
      if (lane == 0) wpart[warp] = w;
    }
    for (int f = 0; f < g.n_fold; ++f) {
      const double w = warp_sum(bins[f * kThreads + tid]);
      if (lane == 0) wpart[(f + 1) * (kThreads / 32) + warp] = w;
    }
    __syncthreads();
    double* out = g.partials + ((long long)band * g.n_chunks + chunk) * nslots;
    for (int q = tid; q < nslots; q += kThreads) {
      double acc = 0.0;
#pragma unroll
      for (int w = 0; w < kThreads / 32; ++w) acc += wpart[q * (kThreads / 32) + w];
      out[q] = acc;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
k2b_weights_finalize(const double* __restrict__ partials, int n_bands, int n_chunks,
                     int n_fold, double* __restrict__ weights, double* __restrict__ norms) {
  const int nslots = n_fold + 1;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n_bands * nslots) return;
  const int band = (int)(idx / nslots);
  const int q = (int)(idx - (long long)band * nslots);
  const double* p = partials + (long long)band * n_chunks * nslots;
  double s = 0.0;
  for (int c = 0; c < n_chunks; ++c) s += p[(long long)c * nslots];
  if (q == 0) {
    if (norms) norms[band] = s;
    return;
  }
  double S = 0.0;
  for (int c = 0; c < n_chunks; ++c) S += p[(long long)c * nslots + q];
  // io_routines_mod.f90:1325: bands with s <= epsilon(1.0_dp) stay unscaled
  weights[(long long)(q - 1) * n_bands + band] = (s > 2.220446049250313e-16) ? S / s : S;
}

size_t k2_smem_bytes(int n_fold) {
  return sizeof(double) * ((size_t)n_fold * kThreads + (size_t)(n_fold + 1) * (kThreads / 32));
}

template <int MODE>
void launch_mode(const WeightsGeom& g, int ctas_per_sm, int sm_count, cudaStream_t s) {
  const size_t smem = k2_smem_bytes(g.n_fold);
  if (smem > 48 * 1024)  // opt in to the large carve-out (attribute is per device)
    cudaFuncSetAttribute(k2_weights_reduce<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(227 * 1024));
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k2_weights_reduce<MODE>, kThreads, smem);
  if (occ < 1) occ = 1;
  if (ctas_per_sm > 0 && ctas_per_sm < occ) occ = ctas_per_sm;
  long long tiles = (long long)g.n_bands * g.n_chunks;
  long long grid = (long long)sm_count * occ;
  if (grid > tiles) grid = tiles;
  if (grid < 1) return;
  k2_weights_reduce<MODE><<<(unsigned)grid, kThreads, smem, s>>>(g);
}

}  // namespace

int weights_default_chunk_vecs() { return kThreads * kUnroll * 2; }  // 64 KiB tiles

int weights_num_chunks(long long row_elems, int chunk_vecs) {
  const long long nv = row_elems / 2;
  long long n = (nv + chunk_vecs - 1) / chunk_vecs;
  return (int)(n < 1 ? 1 : n);
}

size_t weights_partials_bytes(int n_bands, int n_chunks, int n_fold) {
  return sizeof(double) * (size_t)n_bands * (size_t)n_chunks * (size_t)(n_fold + 1);
}

void launch_weights_reduce(const WeightsGeom& g, int ctas_per_sm, int sm_count,
                           cudaStream_t s) {
  if (g.n_bands <= 0) return;
  if (g.n_spinor == 1)
    launch_mode<0>(g, ctas_per_sm, sm_count, s);
  else if (g.layout == BANDUP_GPU_LAYOUT_FORTRAN_INMEM)
    launch_mode<1>(g, ctas_per_sm, sm_count, s);
  else
    launch_mode<2>(g, ctas_per_sm, sm_count, s);
}

void launch_weights_finalize(const double* d_partials, int n_bands, int n_chunks,
                             int n_fold, double* d_weights, double* d_norms,
                             cudaStream_t s) {
  const long long n = (long long)n_bands * (n_fold + 1);
  if (n <= 0) return;
  k2b_weights_finalize<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(
      d_partials, n_bands, n_chunks, n_fold, d_weights, d_norms);
}

}  // namespace bandup
