// weights_reduce.cu -- K2: fused band-norm + masked |C|^2 segmented reduction.
//
// The one kernel of the path that touches every plane-wave coefficient.  It
// replaces three reference loops with a single streaming pass:
//   * the band renormalisation of read_wavefunction
//     (io_routines_mod.f90:1317-1330): s_m = sum_{spinor,G} |C_m(G)|^2
//   * spectral_weight_for_coeff (band_unfolding_routines_mod.f90:638-644) and its
//     spinor sum (:1337-1345): S_{m,f} = sum_{G in coset f} |C_m(G)|^2
// for ALL pc-kpts f folding onto the SC-kpt at once.  P = S/s is formed by K3's finalisation phase.
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
//
// Two accumulator layouts (template NREG).  NREG = 0: the shared-memory bins above, for up to
// kMaxFolds distinct cosets; a real branch per vector, so it is cheapest when selected plane
// waves are rare (det M in the hundreds).  NREG = 1/2/4: at most NREG distinct cosets, the
// common case (one to three pc-kpts fold onto an SC-kpt) -- every fold has a REGISTER
// accumulator fed by predicated adds (no branch, no shared-memory read-modify-write: with a
// small det M or spinor rows nearly every warp used to take the branch), and the tile epilogue
// is warp-local (shuffles, one partial per warp: partials[band][chunk][warp][slot]), so the
// eight consumer warps never meet at a CTA barrier and drift freely across tile boundaries.
//
// Launched with programmatic stream serialisation behind K1: the producer starts filling the
// ring while K1 still classifies (the coefficients do not depend on K1); the consumers wait for
// K1's slot tables at griddepcontrol.wait.
#include <cmath>

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

struct TileGeom {  // where one (k-point, chunk, band) tile lies
  const KptGeom* g;
  const char* row;     // first byte of the band record
  int e_al;            // 1 when the row starts 8 bytes before a 16-byte boundary
  long long nvb;       // whole aligned 16-byte vectors in the row
  long long j0, j1;    // vector range [j0, j1) of this tile
  int band, chunk;
};

// Tiles are numbered k-point-major, then chunk-major, then band: the CTAs resident together
// work on the same stretch of one k-point's slot table.  `k` is a cursor: a CTA's tile indices
// only grow, so the k-point is found by stepping forward.
__device__ __forceinline__ TileGeom tile_geom(const KptGeom* __restrict__ geoms, int n_kpts, int& k,
                                              long long tile) {
  while (k + 1 < n_kpts && tile >= geoms[k + 1].tile_begin) ++k;
  const KptGeom* g = geoms + k;
  const long long local = tile - g->tile_begin;
  TileGeom t;
  t.g = g;
  t.chunk = (int)(local / g->n_bands);
  t.band = (int)(local - (long long)t.chunk * g->n_bands);
  t.row = g->coeffs + (long long)t.band * g->stride_bytes;
  // complex128 rows: one element per 16-byte vector, always aligned (the API insists on it)
  t.e_al = g->dp ? 0 : (int)((reinterpret_cast<unsigned long long>(t.row) >> 3) & 1ull);
  t.nvb = g->dp ? g->row_elems : (g->row_elems - t.e_al) >> 1;
  t.j0 = (long long)t.chunk * g->chunk_vecs;
  t.j1 = t.j0 + g->chunk_vecs;
  if (t.j1 > t.nvb) t.j1 = t.nvb;
  if (t.j0 > t.j1) t.j0 = t.j1;
  return t;
}

// Where the consumers find the slot ids of a tile's vectors (hoisted out of the stage loop:
// the descriptor lives in global memory and the mbarrier asm statements are memory clobbers).
struct SlotSrc {
  const uint16_t* tab16;  // MODE 0: slot[] (16-byte aligned rows) or slot_shift[] (rows off by 8 bytes)
  const uint8_t* slot;
  int n_pw;
  int e_al;
};

template <int MODE>
__device__ __forceinline__ SlotSrc make_slot_src(const KptGeom& g, int e_al) {
  SlotSrc ss;
  ss.slot = g.slot;
  ss.tab16 = reinterpret_cast<const uint16_t*>(e_al ? g.slot_shift : g.slot);
  ss.n_pw = g.n_pw;
  ss.e_al = e_al;
  return ss;
}

// slot ids of the two row elements (e, e+1) held by 16-byte vector J of the row, packed
// sa | sb << 8 (0xFFFF = neither is needed by any fold)
template <int MODE>
__device__ __forceinline__ unsigned slot_pair(const SlotSrc& ss, int J) {
  if (MODE == 0) {
    // either way ONE aligned 16-bit load
    return __ldg(ss.tab16 + J);
  } else if (MODE == 1 && ss.e_al == 0) {
    const unsigned q = __ldg(ss.slot + J);
    return q | (q << 8);
  } else {
    const long long e = ss.e_al + 2LL * J;
    const unsigned sa = __ldg(ss.slot + pw_of<MODE>(e, ss.n_pw));
    const unsigned sb = __ldg(ss.slot + pw_of<MODE>(e + 1, ss.n_pw));
    return sa | (sb << 8);
  }
}

// One 16 KB stage: slot ids (L1/L2) are requested before the wait so they overlap it; the
// stage is handed back to the producer as soon as its four vectors sit in registers.
// FULL = all 1024 vectors present (no bounds checks in the hot path).
template <int MODE, bool FULL>
__device__ __forceinline__ void consume_stage(const float4* __restrict__ sv, uint64_t* full, uint64_t* empty,
                                              uint32_t phase, const SlotSrc& ss, int j, int nv, int tid,
                                              int lane, unsigned nf, double* __restrict__ bins,
                                              double& nrm) {
  unsigned sl[kVecPerThread];
  float4 v[kVecPerThread];
#pragma unroll
  for (int u = 0; u < kVecPerThread; ++u) {
    const int i = tid + u * kConsumers;
    sl[u] = (FULL || i < nv) ? slot_pair<MODE>(ss, j + i) : 0xFFFFu;
  }
  mbar_wait(full, phase);
#pragma unroll
  for (int u = 0; u < kVecPerThread; ++u) {
    const int i = tid + u * kConsumers;
    v[u] = (FULL || i < nv) ? sv[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncwarp();
  if (lane == 0) mbar_arrive(empty);

  // |C|^2 in fp32 (the reference's abs() of a complex(sp) is fp32).  The band norm takes the
  // eight terms of this stage as one short fp32 run promoted to fp64 (relative error of a run
  // ~1e-7, of the whole norm ~1e-9); the coset bins take every selected term widened to fp64.
  float p[2 * kVecPerThread];
  float run = 0.f;
  unsigned all = 0xFFFFu;
#pragma unroll
  for (int u = 0; u < kVecPerThread; ++u) {
    p[2 * u] = fmaf(v[u].y, v[u].y, v[u].x * v[u].x);
    p[2 * u + 1] = fmaf(v[u].w, v[u].w, v[u].z * v[u].z);
    run += p[2 * u];
    run += p[2 * u + 1];
    all &= sl[u];
  }
  nrm += (double)run;
  if (all != 0xFFFFu) {  // some element of this thread's vectors belongs to a requested coset
#pragma unroll
    for (int u = 0; u < kVecPerThread; ++u) {
      if (sl[u] != 0xFFFFu) {  // a real branch per vector: with det(M) in the hundreds most warps skip it
        const unsigned sa = sl[u] & 0xFFu, sb = sl[u] >> 8;
        if (MODE == 1 && sa == sb) {  // the two spinor components of ONE plane wave: one update
          bins[sa * kConsumers + tid] += (double)p[2 * u] + (double)p[2 * u + 1];
        } else {
          if (sa < nf) bins[sa * kConsumers + tid] += (double)p[2 * u];
          if (sb < nf) bins[sb * kConsumers + tid] += (double)p[2 * u + 1];
        }
      }
    }
  }
}

// The same stage with register accumulators: acc[f] of this thread takes the |C|^2 of the
// elements whose slot id is f.  The eight terms of a stage are combined per fold as one short
// fp32 run (almost always a single non-zero term: exact) and widened to fp64 once per stage.
// PAIRED: spinor rows [pw][spinor] on a 16-byte boundary -- a vector is the two components of
// ONE plane wave and carries one slot id.
template <int MODE, int NREG, bool FULL, bool PAIRED>
__device__ __forceinline__ void consume_stage_reg(const float4* __restrict__ sv, uint64_t* full, uint64_t* empty,
                                                  uint32_t phase, const SlotSrc& ss, int j, int nv, int tid,
                                                  int lane, double (&acc)[NREG], double& nrm) {
  unsigned sl[kVecPerThread];
  float4 v[kVecPerThread];
#pragma unroll
  for (int u = 0; u < kVecPerThread; ++u) {
    const int i = tid + u * kConsumers;
    if (PAIRED)
      sl[u] = (FULL || i < nv) ? (unsigned)__ldg(ss.slot + j + i) : 0xFFu;
    else
      sl[u] = (FULL || i < nv) ? slot_pair<MODE>(ss, j + i) : 0xFFFFu;
  }
  mbar_wait(full, phase);
#pragma unroll
  for (int u = 0; u < kVecPerThread; ++u) {
    const int i = tid + u * kConsumers;
    v[u] = (FULL || i < nv) ? sv[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncwarp();
  if (lane == 0) mbar_arrive(empty);

  float run = 0.f;
  float r[NREG];
#pragma unroll
  for (int f = 0; f < NREG; ++f) r[f] = 0.f;
#pragma unroll
  for (int u = 0; u < kVecPerThread; ++u) {
    const float pa = fmaf(v[u].y, v[u].y, v[u].x * v[u].x);
    const float pb = fmaf(v[u].w, v[u].w, v[u].z * v[u].z);
    run += pa;
    run += pb;
    if (PAIRED) {
#pragma unroll
      for (int f = 0; f < NREG; ++f) {
        r[f] += (sl[u] == (unsigned)f) ? pa : 0.f;
        r[f] += (sl[u] == (unsigned)f) ? pb : 0.f;
      }
    } else {
      const unsigned sa = sl[u] & 0xFFu, sb = sl[u] >> 8;
#pragma unroll
      for (int f = 0; f < NREG; ++f) {
        r[f] += (sa == (unsigned)f) ? pa : 0.f;
        r[f] += (sb == (unsigned)f) ? pb : 0.f;
      }
    }
  }
  nrm += (double)run;
#pragma unroll
  for (int f = 0; f < NREG; ++f) acc[f] += (double)r[f];
}

// The same stage for complex128 coefficients (a `kind_cplx_coeffs = dp` build of the reference,
// constants_and_types_mod.f90:60): one element per 16-byte vector, |C|^2 and every sum in fp64.
// NREG > 0: register accumulators; NREG == 0: the per-thread shared-memory bins.
template <int MODE, int NREG, bool FULL>
__device__ __forceinline__ void consume_stage_dp(const float4* __restrict__ sv, uint64_t* full, uint64_t* empty,
                                                 uint32_t phase, const SlotSrc& ss, int j, int nv, int tid,
                                                 int lane, unsigned nf, double* __restrict__ bins,
                                                 double (&acc)[NREG > 0 ? NREG : 1], double& nrm) {
  unsigned sl[kVecPerThread];
  double2 v[kVecPerThread];
#pragma unroll
  for (int u = 0; u < kVecPerThread; ++u) {
    const int i = tid + u * kConsumers;
    sl[u] = (FULL || i < nv) ? (unsigned)__ldg(ss.slot + pw_of<MODE>(j + i, ss.n_pw)) : 0xFFu;
  }
  mbar_wait(full, phase);
  const double2* __restrict__ dv = reinterpret_cast<const double2*>(sv);
#pragma unroll
  for (int u = 0; u < kVecPerThread; ++u) {
    const int i = tid + u * kConsumers;
    v[u] = (FULL || i < nv) ? dv[i] : make_double2(0.0, 0.0);
  }
  __syncwarp();
  if (lane == 0) mbar_arrive(empty);
#pragma unroll
  for (int u = 0; u < kVecPerThread; ++u) {
    const double p = fma(v[u].y, v[u].y, v[u].x * v[u].x);
    nrm += p;
    if (NREG > 0) {
#pragma unroll
      for (int f = 0; f < NREG; ++f) acc[f] += (sl[u] == (unsigned)f) ? p : 0.0;
    } else if (sl[u] < nf) {
      bins[sl[u] * kConsumers + tid] += p;
    }
  }
}

__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait_primary() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// |C|^2 of the row head / tail element that does not fill an aligned 16-byte vector
__device__ __forceinline__ float lone_element(const float2* el, long long e) {
  const float2 c = el[e];
  return fmaf(c.y, c.y, c.x * c.x);
}

struct Ring {  // a consumer's view of the CTA's stage ring
  float4* stage_base;
  uint64_t* full_bar;
  uint64_t* empty_bar;
  int n_stages;
  int s;
  uint32_t phase;
  __device__ __forceinline__ void advance() {
    if (++s == n_stages) {
      s = 0;
      phase ^= 1u;
    }
  }
  __device__ __forceinline__ const float4* stage() const { return stage_base + (size_t)s * kStageVecs; }
};

struct LoneElems {  // the row head / tail elements of a tile (thread 0 only), complex64 rows only
  float p[2];
  unsigned q[2];
};

// One tile with register accumulators (at most NREG distinct cosets) and a warp-local epilogue:
// one partial per warp and slot, partials[band][chunk][warp][slot] -- the eight consumer warps never
// meet at a CTA barrier.
template <int MODE, int NREG, bool DP>
__device__ __forceinline__ void tile_registers(const TileGeom& t, const SlotSrc& ss, Ring& r, const LoneElems& lone,
                                               double nrm, int tid, int lane, int warp) {
  const KptGeom& g = *t.g;
  const int n_fold = g.n_fold, nslots = n_fold + 1;
  const int j0 = (int)t.j0, j1 = (int)t.j1;  // vectors per row < 2^31
  double racc[NREG];
#pragma unroll
  for (int f = 0; f < NREG; ++f)
    racc[f] = (lone.q[0] == (unsigned)f ? (double)lone.p[0] : 0.0) + (lone.q[1] == (unsigned)f ? (double)lone.p[1] : 0.0);
  int j = j0;
  if constexpr (DP) {
    for (; j + kStageVecs <= j1; j += kStageVecs) {
      consume_stage_dp<MODE, NREG, true>(r.stage(), r.full_bar + r.s, r.empty_bar + r.s, r.phase, ss, j, kStageVecs, tid,
                                         lane, 0u, nullptr, racc, nrm);
      r.advance();
    }
    if (j < j1) {
      consume_stage_dp<MODE, NREG, false>(r.stage(), r.full_bar + r.s, r.empty_bar + r.s, r.phase, ss, j, j1 - j, tid,
                                          lane, 0u, nullptr, racc, nrm);
      r.advance();
    }
  } else if (MODE == 1 && t.e_al == 0) {  // spinor rows [pw][spinor] on a 16-byte boundary
    for (; j + kStageVecs <= j1; j += kStageVecs) {
      consume_stage_reg<MODE, NREG, true, true>(r.stage(), r.full_bar + r.s, r.empty_bar + r.s, r.phase, ss, j,
                                                kStageVecs, tid, lane, racc, nrm);
      r.advance();
    }
    if (j < j1) {
      consume_stage_reg<MODE, NREG, false, true>(r.stage(), r.full_bar + r.s, r.empty_bar + r.s, r.phase, ss, j, j1 - j,
                                                 tid, lane, racc, nrm);
      r.advance();
    }
  } else {
    for (; j + kStageVecs <= j1; j += kStageVecs) {
      consume_stage_reg<MODE, NREG, true, false>(r.stage(), r.full_bar + r.s, r.empty_bar + r.s, r.phase, ss, j,
                                                 kStageVecs, tid, lane, racc, nrm);
      r.advance();
    }
    if (j < j1) {
      consume_stage_reg<MODE, NREG, false, false>(r.stage(), r.full_bar + r.s, r.empty_bar + r.s, r.phase, ss, j, j1 - j,
                                                  tid, lane, racc, nrm);
      r.advance();
    }
  }
  // fixed shuffle tree => bit-reproducible; K3 adds the warps' partials up in warp order
  double* out = g.partials + (((long long)t.band * g.n_chunks + t.chunk) * (kConsumers / 32) + warp) * nslots;
  const double v0 = warp_sum(nrm);
  if (lane == 0) out[0] = v0;
#pragma unroll
  for (int f = 0; f < NREG; ++f) {
    if (f < n_fold) {  // warp-uniform
      const double vf = warp_sum(racc[f]);
      if (lane == 0) out[f + 1] = vf;
    }
  }
}

// One tile with the per-thread shared-memory bins (up to kMaxFolds distinct cosets) and the
// CTA-level epilogue: partials[band][chunk][slot].
template <int MODE, bool DP>
__device__ __forceinline__ void tile_bins(const TileGeom& t, const SlotSrc& ss, Ring& r, const LoneElems& lone,
                                          double nrm, double* __restrict__ acc, int tid, int lane, int warp) {
  const KptGeom& g = *t.g;
  const int n_fold = g.n_fold, nslots = n_fold + 1;
  const unsigned nf = (unsigned)n_fold;
  const int j0 = (int)t.j0, j1 = (int)t.j1;
  double* bins = acc + kConsumers;
  for (int f = 0; f < n_fold; ++f) bins[f * kConsumers + tid] = 0.0;
  if (!DP && tid == 0) {
    if (lone.q[0] < nf) bins[lone.q[0] * kConsumers] += (double)lone.p[0];
    if (lone.q[1] < nf) bins[lone.q[1] * kConsumers] += (double)lone.p[1];
  }
  int j = j0;
  if constexpr (DP) {
    double unused[1] = {0.0};
    for (; j + kStageVecs <= j1; j += kStageVecs) {
      consume_stage_dp<MODE, 0, true>(r.stage(), r.full_bar + r.s, r.empty_bar + r.s, r.phase, ss, j, kStageVecs, tid,
                                      lane, nf, bins, unused, nrm);
      r.advance();
    }
    if (j < j1) {
      consume_stage_dp<MODE, 0, false>(r.stage(), r.full_bar + r.s, r.empty_bar + r.s, r.phase, ss, j, j1 - j, tid, lane,
                                       nf, bins, unused, nrm);
      r.advance();
    }
  } else {
    for (; j + kStageVecs <= j1; j += kStageVecs) {
      consume_stage<MODE, true>(r.stage(), r.full_bar + r.s, r.empty_bar + r.s, r.phase, ss, j, kStageVecs, tid, lane,
                                nf, bins, nrm);
      r.advance();
    }
    if (j < j1) {  // last, partial stage of a row
      consume_stage<MODE, false>(r.stage(), r.full_bar + r.s, r.empty_bar + r.s, r.phase, ss, j, j1 - j, tid, lane, nf,
                                 bins, nrm);
      r.advance();
    }
  }
  // tile reduction through shared memory: thread partials are already there for the folds;
  // warp q adds up row q (256 values: 8 per lane, conflict-free, then a shuffle tree).
  // Fixed order => bit-reproducible.
  acc[tid] = nrm;
  consumer_sync();
  double* out = g.partials + ((long long)t.band * g.n_chunks + t.chunk) * nslots;
  for (int q = warp; q < nslots; q += kConsumers / 32) {
    const double* row = acc + q * kConsumers + lane;
    double a = 0.0;
#pragma unroll
    for (int i = 0; i < kConsumers / 32; ++i) a += row[32 * i];
    a = warp_sum(a);
    if (lane == 0) out[q] = a;
  }
  consumer_sync();
}

// bins_rows: rows of the shared-memory accumulator block = 1 + the largest n_fold among the
// k-points that use the bins (0 when every k-point of the batch uses registers).
template <int MODE, bool DP>
__global__ void __launch_bounds__(kThreads, 2)
k2_weights_reduce(const KptGeom* __restrict__ geoms, int n_kpts, long long total_tiles, int bins_rows,
                  int n_stages) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // layout: [n_stages][16 KB] | acc [bins_rows][256] f64 (row 0: norm, row f+1: fold f) | barriers
  float4* stage_base = reinterpret_cast<float4*>(smem_raw);
  double* acc = reinterpret_cast<double*>(smem_raw + (size_t)n_stages * kStageBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(acc + (size_t)bins_rows * kConsumers);
  uint64_t* empty_bar = full_bar + kMaxStages;

  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(full_bar + s, 1);                  // producer's expect_tx arrive
      mbar_init(empty_bar + s, kConsumers / 32);   // one arrive per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  pdl_launch_dependents();  // K3 may be scheduled behind this grid; it waits for it to finish

  if (tid >= kConsumers) {
    // ------------------------------ producer ------------------------------
    // reads the descriptors (uploaded before K1) and the coefficients only: nothing K1 writes
    if (tid == kConsumers) {
      const uint64_t policy = l2_evict_first_policy();
      int s = 0, k = 0;
      uint32_t phase = 0;
      for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileGeom t = tile_geom(geoms, n_kpts, k, tile);
        const char* src = t.row + 8 * t.e_al;
        for (long long j = t.j0; j < t.j1; j += kStageVecs) {
          const long long nv = (t.j1 - j < kStageVecs) ? (t.j1 - j) : kStageVecs;
          mbar_wait(empty_bar + s, phase ^ 1u);
          mbar_expect_tx(full_bar + s, (uint32_t)(nv * 16));
          bulk_g2s(stage_base + (size_t)s * kStageVecs, src + j * 16, (uint32_t)(nv * 16), full_bar + s,
                   policy);
          if (++s == n_stages) {
            s = 0;
            phase ^= 1u;
          }
        }
      }
    }
    return;
  }

  // -------------------------------- consumers --------------------------------
  pdl_wait_primary();  // K1's slot tables
  const int lane = tid & 31, warp = tid >> 5;
  Ring ring{stage_base, full_bar, empty_bar, n_stages, 0, 0u};
  int k = 0;
  for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const TileGeom t = tile_geom(geoms, n_kpts, k, tile);
    const KptGeom& g = *t.g;
    const SlotSrc ss = make_slot_src<MODE>(g, t.e_al);
    double nrm = 0.0;
    // the row head / tail elements that do not fill an aligned vector fall to thread 0
    LoneElems lone{{0.f, 0.f}, {kNoFold, kNoFold}};
    if (!DP && tid == 0) {
      const float2* el = reinterpret_cast<const float2*>(t.row);
      const long long row_elems = g.row_elems, e_tail = t.e_al + 2 * t.nvb;
      if (t.chunk == 0 && t.e_al == 1 && row_elems > 0) {
        lone.p[0] = lone_element(el, 0);
        lone.q[0] = ss.slot[pw_of<MODE>(0, ss.n_pw)];
      }
      if (t.chunk == g.n_chunks - 1 && e_tail < row_elems) {
        lone.p[1] = lone_element(el, e_tail);
        lone.q[1] = ss.slot[pw_of<MODE>(e_tail, ss.n_pw)];
      }
      nrm = (double)lone.p[0] + (double)lone.p[1];
    }
    // the accumulator layout is a property of the k-point (host: weights_pick_nreg)
    switch (g.nreg) {
      case 1: tile_registers<MODE, 1, DP>(t, ss, ring, lone, nrm, tid, lane, warp); break;
      case 2: tile_registers<MODE, 2, DP>(t, ss, ring, lone, nrm, tid, lane, warp); break;
      case 4: tile_registers<MODE, 4, DP>(t, ss, ring, lone, nrm, tid, lane, warp); break;
      default: tile_bins<MODE, DP>(t, ss, ring, lone, nrm, acc, tid, lane, warp); break;
    }
  }
}

constexpr size_t kSmemBudget = 227 * 1024;

size_t k2_fixed_smem(int bins_rows) {  // everything but the stage ring
  return sizeof(double) * (size_t)bins_rows * kConsumers + sizeof(uint64_t) * 2 * kMaxStages + 128;
}

template <int MODE, bool DP>
void launch_mode(const KptGeom* d_geoms, int n_kpts, long long total_tiles, int bins_rows,
                 int ctas_per_sm, int n_stages, int sm_count, bool pdl, cudaStream_t s) {
  const size_t fixed = k2_fixed_smem(bins_rows);
  // default: 2 CTAs per SM with 4 stages each (128 KB of loads in flight per SM; the round-2 sweep
  // under the power cap, profiles/r02_ring_sweep.md, has 2x4 ahead of 2x3 by 1-2 % on the small and
  // medium cells and level on the 1000-atom cell, one CTA per SM with 6-12 stages 5-20 % behind);
  // the fourth stage is free: it fits the shared memory the CTA claims anyway (see below).  Many
  // folds eat shared memory for the bins, so the ring shrinks before occupancy does
  if (ctas_per_sm <= 0) ctas_per_sm = 2;
  if (ctas_per_sm > 2) ctas_per_sm = 2;  // the kernel is compiled for two resident CTAs per SM
  if (n_stages <= 0) n_stages = 4;
  if (n_stages > kMaxStages) n_stages = kMaxStages;
  while (ctas_per_sm > 1 && (fixed + 2 * (size_t)kStageBytes + 1024) * ctas_per_sm > kSmemBudget) --ctas_per_sm;
  const size_t per_cta = kSmemBudget / ctas_per_sm - 1024;  // 1 KB per CTA is reserved by the driver
  while (n_stages > 2 && fixed + (size_t)n_stages * kStageBytes > per_cta) --n_stages;
  size_t smem = fixed + (size_t)n_stages * kStageBytes;
  // The persistent grid hands every CTA the same share of the tiles, so the CTAs must be spread
  // evenly: ctas_per_sm on every SM, never three here and one there.  When the kernel is launched
  // behind K1 with programmatic dependent launch, K1's last CTAs still occupy some SMs and the
  // block scheduler would pack the others.  Claiming shared memory so that exactly ctas_per_sm fit
  // makes the extra CTAs wait for their SM instead.
  const size_t claim = kSmemBudget / (size_t)(ctas_per_sm + 1) + 1024;
  if (smem < claim) smem = claim < per_cta ? claim : per_cta;
  auto kernel = k2_weights_reduce<MODE, DP>;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kThreads, smem);
  if (occ < 1) occ = 1;
  if (occ > ctas_per_sm) occ = ctas_per_sm;
  long long grid = (long long)sm_count * occ;
  if (grid > total_tiles) grid = total_tiles;
  if (grid < 1) return;
  launch_kernel(kernel, dim3((unsigned)grid), dim3(kThreads), smem, s, pdl, d_geoms, n_kpts, total_tiles, bins_rows,
                n_stages);
}

template <bool DP>
void launch_precision(const KptGeom* d_geoms, int n_kpts, long long total_tiles, int bins_rows, int n_spinor,
                      int layout, int ctas_per_sm, int n_stages, int sm_count, bool pdl, cudaStream_t s) {
  if (n_spinor == 1)
    launch_mode<0, DP>(d_geoms, n_kpts, total_tiles, bins_rows, ctas_per_sm, n_stages, sm_count, pdl, s);
  else if (layout == BANDUP_GPU_LAYOUT_FORTRAN_INMEM)
    launch_mode<1, DP>(d_geoms, n_kpts, total_tiles, bins_rows, ctas_per_sm, n_stages, sm_count, pdl, s);
  else
    launch_mode<2, DP>(d_geoms, n_kpts, total_tiles, bins_rows, ctas_per_sm, n_stages, sm_count, pdl, s);
}

}  // namespace

// Tile size, from a small cost model fitted to B200 sweeps (profiles/r01_tile_sweep.md): a CTA of
// the full persistent grid streams ~24 GB/s (7.2 TB/s / 296), a tile costs its bytes at that rate
// plus the part of its epilogue (shared-memory reduction, partials, two barriers: ~1.2 us) that the
// stage ring does not hide.  Large tiles amortise
// the epilogue (3000 x 100k block: 64 KB tiles 5.9 TB/s, 128 KB 6.7, 256 KB 7.25, 512 KB 7.33); with few rows
// the number of ROUNDS of the grid is what counts (graphene, 512 rows of 184 KB: one tile per row
// 25.5 us, six 32 KB tiles per row 31.0 us).  The candidate with the lowest estimate wins.
int weights_pick_chunk_vecs(int requested, long long row_elems, int n_bands, int sm_count, int n_spinor) {
  if (requested > 0) return ((requested + kStageVecs - 1) / kStageVecs) * kStageVecs;
  const double nv = (double)(row_elems / 2 > 0 ? row_elems / 2 : 1);
  const double grid = 2.0 * sm_count, rows = n_bands > 0 ? n_bands : 1;
  const double t_epi = 0.5e-6, full_rate = 7.2e12 / grid;  // what the stage ring does not hide of the epilogue
  // 512 KB tiles pay off for scalar rows (C5 +1.1 %); spinor rows measured best at 256 KB
  const int cv_max = (n_spinor == 1 ? 32 : 16) * kStageVecs;
  int best = cv_max;
  double best_t = 1e30;
  for (int cv = cv_max; cv >= kStageVecs; cv >>= 1) {
    const double chunks = std::ceil(nv / cv), tiles = rows * chunks;
    const double active = tiles < grid ? tiles : grid;
    double rate = 7.2e12 / active;  // fewer CTAs than the grid: each gets more, up to what one ring sustains
    if (rate > 48e9) rate = 48e9;
    if (rate < full_rate) rate = full_rate;
    const double tile_vecs = nv < cv ? nv : cv;
    double t;
    if (tiles >= 4.0 * grid)
      t = rows * (nv * 16.0 / rate + chunks * t_epi) / grid + (tile_vecs * 16.0 / rate + t_epi);
    else
      t = std::ceil(tiles / grid) * (tile_vecs * 16.0 / rate + t_epi);
    if (t < best_t * 0.999) {  // ties go to the larger tile
      best_t = t;
      best = cv;
    }
  }
  return best;
}

int weights_num_chunks(long long row_elems, int chunk_vecs) {
  const long long nv = row_elems / 2;
  long long n = (nv + chunk_vecs - 1) / chunk_vecs;
  return (int)(n < 1 ? 1 : n);
}

// Accumulator layout of one k-point (pass): 0 = per-thread shared-memory bins (any number of
// cosets; a real branch per vector, cheapest when selected plane waves are rare), 1/2/4 = that many
// register accumulators with predicated adds (no branch, no shared-memory read-modify-write;
// every element costs a compare and an add per accumulator, so it pays when selected plane waves
// are frequent: n_fold / |det M| >= 1/128 -- graphene 4x4, the spinor 8x8 cell, Si 3x3x3, not the
// 500-fold Si-Ge cell with its three pc-kpts per SC-kpt; profiles/r02_k2_variants.md).  variant: 0 automatic, 1 bins always, 2 registers whenever n_fold <= 4.
int weights_pick_nreg(int variant, int n_fold, long long det_m) {
  if (variant == 1 || n_fold > 4 || n_fold < 1) return 0;
  if (variant == 0 && (long long)n_fold * 128 < det_m) return 0;
  return n_fold <= 1 ? 1 : n_fold <= 2 ? 2 : 4;
}

// partials per (band, chunk): one row of (n_fold + 1) per CTA (bins) or per consumer warp (registers)
int weights_partials_sub(int nreg) { return nreg > 0 ? kConsumers / 32 : 1; }

size_t weights_partials_bytes(int n_bands, int n_chunks, int n_fold) {
  // room for either layout: n_fold + 1 slots once, or eight warp rows of at most 4 + 1 slots
  const size_t per = (size_t)(n_fold + 1) > 40 ? (size_t)(n_fold + 1) : 40;
  return sizeof(double) * (size_t)n_bands * (size_t)n_chunks * per;
}

int weights_max_folds_supported() {
  // bins + 2 stages must fit one CTA
  int nf = (int)((kSmemBudget - 1024 - 2 * kStageBytes - sizeof(uint64_t) * 2 * kMaxStages - 128) /
                 (sizeof(double) * kConsumers)) - 1;
  return nf;
}

void launch_weights_reduce(const KptGeom* d_geoms, int n_kpts, long long total_tiles, int bins_rows,
                           int n_spinor, int layout, bool dp, int ctas_per_sm, int n_stages, int sm_count,
                           bool pdl, cudaStream_t s) {
  if (n_kpts <= 0 || total_tiles <= 0) return;
  if (dp)
    launch_precision<true>(d_geoms, n_kpts, total_tiles, bins_rows, n_spinor, layout, ctas_per_sm, n_stages, sm_count,
                           pdl, s);
  else
    launch_precision<false>(d_geoms, n_kpts, total_tiles, bins_rows, n_spinor, layout, ctas_per_sm, n_stages, sm_count,
                            pdl, s);
}

}  // namespace bandup
