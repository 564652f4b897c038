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
// complex64 read exactly once) + the slot table (L2/L1-resident, re-read per
// band row from cache, not HBM).  Arithmetic intensity ~0.5 flop/B: no
// tensor-core work here.
//
// Shape (sm_100a): a persistent, warp-specialised grid.  Each CTA owns a ring of
// 16 KB shared-memory stages.  One producer thread walks the CTA's (chunk, band)
// tiles and keeps the ring full with 1-D bulk async copies (cp.async.bulk
// global -> shared, completion on an mbarrier, L2 evict-first since every byte is
// used once); it runs ahead ACROSS tile boundaries, so the bytes in flight per SM
// (stages x 16 KB x resident CTAs) never drain while the consumers reduce a tile.
// Eight consumer warps wait on a stage's `full` barrier, read it with
// conflict-free 128-bit LDS, form |C|^2 in fp32 (as the reference's abs() is),
// widen every term to fp64 and add it to the band-norm accumulator and, for the
// few selected plane waves, to per-thread fp64 bins in shared memory
// (conflict-free: bin f of thread t lives at bins[f*256 + t]); then release the
// stage through its `empty` barrier.  Warp shuffles + one shared-memory hop
// reduce the tile.  Fully deterministic: no atomics, fixed combination order.
#include "kernels.cuh"

namespace bandup {

namespace {

constexpr int kConsumers = 256;               // 8 consumer warps
constexpr int kThreads = kConsumers + 32;     // + 1 producer warp
constexpr int kStageVecs = 1024;              // 16-byte vectors per stage (16 KB)
constexpr int kStageBytes = kStageVecs * 16;
constexpr int kVecPerThread = kStageVecs / kConsumers;
constexpr int kMaxStages = 12;

// ---- PTX helpers: mbarrier + bulk async copy (TMA engine, 1-D) --------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar,
                                         uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void consumer_sync() {  // named barrier of the 8 consumer warps
  asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory");
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

struct TileGeom {  // where one (chunk, band) tile lies
  const char* row;     // first byte of the band record
  int e_al;            // 1 when the row starts 8 bytes before a 16-byte boundary
  long long nvb;       // whole aligned 16-byte vectors in the row
  long long j0, j1;    // vector range [j0, j1) of this tile
  int band, chunk;
};

__device__ __forceinline__ TileGeom tile_geom(const WeightsGeom& g, long long tile) {
  TileGeom t;
  t.chunk = (int)(tile / g.n_bands);
  t.band = (int)(tile - (long long)t.chunk * g.n_bands);
  t.row = g.coeffs + (long long)t.band * g.stride_bytes;
  t.e_al = (int)((reinterpret_cast<unsigned long long>(t.row) >> 3) & 1ull);
  t.nvb = (g.row_elems - t.e_al) >> 1;
  t.j0 = (long long)t.chunk * g.chunk_vecs;
  t.j1 = t.j0 + g.chunk_vecs;
  if (t.j1 > t.nvb) t.j1 = t.nvb;
  if (t.j0 > t.j1) t.j0 = t.j1;
  return t;
}

// slot ids of the two row elements (e, e+1) held by one 16-byte vector
template <int MODE>
__device__ __forceinline__ void slot_pair(const WeightsGeom& g, int e_al, long long J, unsigned& sa,
                                          unsigned& sb) {
  if (MODE == 0) {
    // slot[] for rows that start on a 16-byte boundary, slot_shift[] (= slot[i+1]) otherwise:
    // either way ONE aligned 16-bit load
    const uint16_t* tab = reinterpret_cast<const uint16_t*>(e_al ? g.slot_shift : g.slot);
    const unsigned s = __ldg(tab + J);
    sa = s & 0xffu;
    sb = s >> 8;
  } else if (MODE == 1 && e_al == 0) {
    sa = sb = __ldg(g.slot + J);
  } else {
    const long long e = e_al + 2 * J;
    sa = __ldg(g.slot + pw_of<MODE>(e, g.n_pw));
    sb = __ldg(g.slot + pw_of<MODE>(e + 1, g.n_pw));
  }
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 3)
k2_weights_reduce(WeightsGeom g) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // layout: [n_stages][16 KB] | bins [n_fold][256] f64 | wpart [(n_fold+1)][8] f64 | barriers
  float4* stage_base = reinterpret_cast<float4*>(smem_raw);
  double* bins = reinterpret_cast<double*>(smem_raw + (size_t)g.n_stages * kStageBytes);
  double* wpart = bins + (size_t)g.n_fold * kConsumers;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(wpart + (size_t)(g.n_fold + 1) * (kConsumers / 32));
  uint64_t* empty_bar = full_bar + kMaxStages;

  const int tid = threadIdx.x;
  const long long total_tiles = (long long)g.n_bands * g.n_chunks;
  if (tid == 0) {
    for (int s = 0; s < g.n_stages; ++s) {
      mbar_init(full_bar + s, 1);                  // producer's expect_tx arrive
      mbar_init(empty_bar + s, kConsumers / 32);   // one arrive per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (tid >= kConsumers) {
    // ------------------------------ producer ------------------------------
    if (tid == kConsumers) {
      const uint64_t policy = l2_evict_first_policy();
      int s = 0;
      uint32_t phase = 0;
      for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileGeom t = tile_geom(g, tile);
        const char* src = t.row + 8 * t.e_al;
        for (long long j = t.j0; j < t.j1; j += kStageVecs) {
          const long long nv = (t.j1 - j < kStageVecs) ? (t.j1 - j) : kStageVecs;
          mbar_wait(empty_bar + s, phase ^ 1u);
          mbar_expect_tx(full_bar + s, (uint32_t)(nv * 16));
          bulk_g2s(stage_base + (size_t)s * kStageVecs, src + j * 16, (uint32_t)(nv * 16), full_bar + s,
                   policy);
          if (++s == g.n_stages) {
            s = 0;
            phase ^= 1u;
          }
        }
      }
    }
    return;
  }

  // -------------------------------- consumers --------------------------------
  const int lane = tid & 31, warp = tid >> 5;
  const int nslots = g.n_fold + 1;
  const unsigned nf = (unsigned)g.n_fold;
  int s = 0;
  uint32_t phase = 0;
  for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const TileGeom t = tile_geom(g, tile);
    for (int f = 0; f < g.n_fold; ++f) bins[f * kConsumers + tid] = 0.0;
    double nrm = 0.0;

    for (long long j = t.j0; j < t.j1; j += kStageVecs) {
      const int nv = (int)((t.j1 - j < kStageVecs) ? (t.j1 - j) : kStageVecs);
      const float4* sv = stage_base + (size_t)s * kStageVecs;
      float4 v[kVecPerThread];
      unsigned sa[kVecPerThread], sb[kVecPerThread];
      // slot ids first: they come from L1/L2 and overlap the wait for the stage
#pragma unroll
      for (int u = 0; u < kVecPerThread; ++u) {
        const int i = tid + u * kConsumers;
        if (i < nv) {
          slot_pair<MODE>(g, t.e_al, j + i, sa[u], sb[u]);
        } else {
          sa[u] = sb[u] = kNoFold;
        }
      }
      mbar_wait(full_bar + s, phase);
#pragma unroll
      for (int u = 0; u < kVecPerThread; ++u) {
        const int i = tid + u * kConsumers;
        v[u] = (i < nv) ? sv[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      // all of this warp's reads of the stage are done: hand it back to the producer
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_bar + s);
      if (++s == g.n_stages) {
        s = 0;
        phase ^= 1u;
      }
#pragma unroll
      for (int u = 0; u < kVecPerThread; ++u) {
        const float pa = fmaf(v[u].y, v[u].y, v[u].x * v[u].x);
        const float pb = fmaf(v[u].w, v[u].w, v[u].z * v[u].z);
        // the SAME fp32 |C|^2 feeds the norm and the coset bins, each widened to
        // fp64 before it is added: sum over all cosets of S == s to fp64 rounding
        nrm += (double)pa;
        nrm += (double)pb;
        if (sa[u] < nf) bins[sa[u] * kConsumers + tid] += (double)pa;
        if (sb[u] < nf) bins[sb[u] * kConsumers + tid] += (double)pb;
      }
    }

    // row head / tail elements that do not fill an aligned vector
    if (tid == 0) {
      const float2* el = reinterpret_cast<const float2*>(t.row);
      if (t.chunk == 0 && t.e_al == 1 && g.row_elems > 0) {
        const float2 c = el[0];
        const float p = fmaf(c.y, c.y, c.x * c.x);
        nrm += (double)p;
        const unsigned q = g.slot[pw_of<MODE>(0, g.n_pw)];
        if (q < nf) bins[q * kConsumers] += (double)p;
      }
      const long long e_tail = t.e_al + 2 * t.nvb;
      if (t.chunk == g.n_chunks - 1 && e_tail < g.row_elems) {
        const float2 c = el[e_tail];
        const float p = fmaf(c.y, c.y, c.x * c.x);
        nrm += (double)p;
        const unsigned q = g.slot[pw_of<MODE>(e_tail, g.n_pw)];
        if (q < nf) bins[q * kConsumers] += (double)p;
      }
    }

    // tile reduction: warp shuffles, then the warps' partials through smem
    {
      const double w = warp_sum(nrm);
      if (lane == 0) wpart[warp] = w;
    }
    for (int f = 0; f < g.n_fold; ++f) {
      const double w = warp_sum(bins[f * kConsumers + tid]);
      if (lane == 0) wpart[(f + 1) * (kConsumers / 32) + warp] = w;
    }
    consumer_sync();
    double* out = g.partials + ((long long)t.band * g.n_chunks + t.chunk) * nslots;
    for (int q = tid; q < nslots; q += kConsumers) {
      double acc = 0.0;
#pragma unroll
      for (int w = 0; w < kConsumers / 32; ++w) acc += wpart[q * (kConsumers / 32) + w];
      out[q] = acc;
    }
    consumer_sync();
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

constexpr size_t kSmemBudget = 227 * 1024;

size_t k2_fixed_smem(int n_fold) {  // everything but the stage ring
  return sizeof(double) * ((size_t)n_fold * kConsumers + (size_t)(n_fold + 1) * (kConsumers / 32)) +
         sizeof(uint64_t) * 2 * kMaxStages + 128;
}

template <int MODE>
void launch_mode(WeightsGeom g, int ctas_per_sm, int n_stages, int sm_count, cudaStream_t s) {
  const size_t fixed = k2_fixed_smem(g.n_fold);
  // default: 3 CTAs per SM with 4 stages each (192 KB of loads in flight per SM; the B200
  // sweep in profiles/ has 3x4 and 2x4 within 1 % of each other and ahead of deeper rings);
  // many folds eat shared memory for the bins, so the ring shrinks before occupancy does
  if (ctas_per_sm <= 0) ctas_per_sm = 3;
  if (n_stages <= 0) n_stages = 4;
  if (n_stages > kMaxStages) n_stages = kMaxStages;
  while (ctas_per_sm > 1 && (fixed + 2 * (size_t)kStageBytes + 1024) * ctas_per_sm > kSmemBudget) --ctas_per_sm;
  const size_t per_cta = kSmemBudget / ctas_per_sm - 1024;  // 1 KB per CTA is reserved by the driver
  while (n_stages > 2 && fixed + (size_t)n_stages * kStageBytes > per_cta) --n_stages;
  g.n_stages = n_stages;
  const size_t smem = fixed + (size_t)n_stages * kStageBytes;
  cudaFuncSetAttribute(k2_weights_reduce<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k2_weights_reduce<MODE>, kThreads, smem);
  if (occ < 1) occ = 1;
  if (occ > ctas_per_sm) occ = ctas_per_sm;
  long long tiles = (long long)g.n_bands * g.n_chunks;
  long long grid = (long long)sm_count * occ;
  if (grid > tiles) grid = tiles;
  if (grid < 1) return;
  k2_weights_reduce<MODE><<<(unsigned)grid, kThreads, smem, s>>>(g);
}

}  // namespace

// Tile size.  Large tiles amortise the per-tile reduction (measured on B200, 3000 x 100k
// block: 64 KB tiles 5.9 TB/s, 128 KB tiles 6.7 TB/s); small blocks need enough tiles to
// balance the persistent grid, so the tile shrinks (down to one 16 KB stage) until there are
// at least 6 tiles per CTA.
int weights_pick_chunk_vecs(int requested, long long row_elems, int n_bands, int sm_count) {
  if (requested > 0) return ((requested + kStageVecs - 1) / kStageVecs) * kStageVecs;
  const long long nv = row_elems / 2;
  const long long want_tiles = 6LL * sm_count * 3;
  int cv = 8 * kStageVecs;
  while (cv > kStageVecs && (long long)n_bands * ((nv + cv - 1) / cv) < want_tiles) cv >>= 1;
  return cv;
}

int weights_num_chunks(long long row_elems, int chunk_vecs) {
  const long long nv = row_elems / 2;
  long long n = (nv + chunk_vecs - 1) / chunk_vecs;
  return (int)(n < 1 ? 1 : n);
}

size_t weights_partials_bytes(int n_bands, int n_chunks, int n_fold) {
  return sizeof(double) * (size_t)n_bands * (size_t)n_chunks * (size_t)(n_fold + 1);
}

int weights_max_folds_supported() {
  // bins + 2 stages must fit one CTA
  int nf = (int)((kSmemBudget - 1024 - 2 * kStageBytes - sizeof(uint64_t) * 2 * kMaxStages - 128 -
                  sizeof(double) * (kConsumers / 32)) /
                 (sizeof(double) * (kConsumers + kConsumers / 32)));
  return nf;
}

void launch_weights_reduce(const WeightsGeom& g, int ctas_per_sm, int n_stages, int sm_count,
                           cudaStream_t s) {
  if (g.n_bands <= 0) return;
  if (g.n_spinor == 1)
    launch_mode<0>(g, ctas_per_sm, n_stages, sm_count, s);
  else if (g.layout == BANDUP_GPU_LAYOUT_FORTRAN_INMEM)
    launch_mode<1>(g, ctas_per_sm, n_stages, sm_count, s);
  else
    launch_mode<2>(g, ctas_per_sm, n_stages, sm_count, s);
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
