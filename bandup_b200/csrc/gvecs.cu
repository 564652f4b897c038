// gvecs.cu -- K0: the plane-wave list of an SC k-point, regenerated on the device.
//
// A WAVECAR does not store its G vectors; the reference's reader regenerates them from the
// k-point, ENCUT and the lattice (n_Gvecs_2b_generated + populate_wf_with_G_vec_coords,
// vasp_related_modules/read_wavecar_mod.f90:213-304): loops ig3 (slowest) -> ig2 -> ig1
// (fastest), each running 0..nbmax then -nbmax..-1, keeping G when |(K+G).B|^2 / c < ENCUT.
// Here every candidate gets one thread (same fp64 expression, explicitly rounded operations in
// the reference's order, no FMA), and an ordered compaction (per-CTA counts -> scan -> write)
// reproduces the reader's enumeration order exactly, so the list lines up with the coefficient
// records byte for byte.
#include "kernels.cuh"

namespace bandup {

namespace {

struct GvecParams {
  double k[3];
  double b[9];   // SC reciprocal lattice, row i = b_i
  double c;      // 2m/hbar^2
  double ecut;
  int nb[3];
  long long total;  // (2nb1+1)(2nb2+1)(2nb3+1)
};

__device__ __forceinline__ bool gvec_keep(const GvecParams& p, long long t, int& p1, int& p2, int& p3) {
  const int n1 = 2 * p.nb[0] + 1, n2 = 2 * p.nb[1] + 1;
  const int i1 = (int)(t % n1);
  const long long r = t / n1;
  const int i2 = (int)(r % n2);
  const int i3 = (int)(r / n2);
  p1 = i1 > p.nb[0] ? i1 - 2 * p.nb[0] - 1 : i1;
  p2 = i2 > p.nb[1] ? i2 - 2 * p.nb[1] - 1 : i2;
  p3 = i3 > p.nb[2] ? i3 - 2 * p.nb[2] - 1 : i3;
  const double f1 = __dadd_rn(p.k[0], (double)p1), f2 = __dadd_rn(p.k[1], (double)p2),
               f3 = __dadd_rn(p.k[2], (double)p3);
  double s[3];
#pragma unroll
  for (int j = 0; j < 3; ++j)
    s[j] = __dadd_rn(__dadd_rn(__dmul_rn(f1, p.b[j]), __dmul_rn(f2, p.b[3 + j])), __dmul_rn(f3, p.b[6 + j]));
  const double g2 = __dadd_rn(__dadd_rn(__dmul_rn(s[0], s[0]), __dmul_rn(s[1], s[1])), __dmul_rn(s[2], s[2]));
  const double gtot = __dsqrt_rn(g2);
  const double etot = __ddiv_rn(__dmul_rn(gtot, gtot), p.c);
  return etot < p.ecut;
}

__global__ void __launch_bounds__(256)
k0_gvec_count(GvecParams p, int* __restrict__ block_count) {
  __shared__ int s_warp[8];
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int a, b, c;
  const bool keep = t < p.total && gvec_keep(p, t, a, b, c);
  const unsigned bal = __ballot_sync(0xffffffffu, keep);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = __popc(bal);
  __syncthreads();
  if (threadIdx.x == 0) {
    int n = 0;
    for (int w = 0; w < 8; ++w) n += s_warp[w];
    block_count[blockIdx.x] = n;
  }
}

__global__ void __launch_bounds__(256)
k0_gvec_write(GvecParams p, const long long* __restrict__ block_off, int32_t* __restrict__ g_frac,
              long long capacity) {
  __shared__ int s_warp[8];
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int p1 = 0, p2 = 0, p3 = 0;
  const bool keep = t < p.total && gvec_keep(p, t, p1, p2, p3);
  const unsigned bal = __ballot_sync(0xffffffffu, keep);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) s_warp[warp] = __popc(bal);
  __syncthreads();
  if (!keep) return;
  int before = 0;
  for (int w = 0; w < warp; ++w) before += s_warp[w];
  const long long pos = block_off[blockIdx.x] + before + __popc(bal & ((1u << lane) - 1u));
  if (pos < capacity) {
    g_frac[3 * pos] = p1;
    g_frac[3 * pos + 1] = p2;
    g_frac[3 * pos + 2] = p3;
  }
}

GvecParams make_params(const double* kfrac, const double* B_rows, double c, double ecut, const int* nb) {
  GvecParams p;
  for (int i = 0; i < 3; ++i) p.k[i] = kfrac[i];
  for (int i = 0; i < 9; ++i) p.b[i] = B_rows[i];
  p.c = c;
  p.ecut = ecut;
  for (int i = 0; i < 3; ++i) p.nb[i] = nb[i];
  p.total = (long long)(2 * nb[0] + 1) * (2 * nb[1] + 1) * (2 * nb[2] + 1);
  return p;
}

}  // namespace

long long gvec_num_blocks(const int* nb) {
  const long long total = (long long)(2 * nb[0] + 1) * (2 * nb[1] + 1) * (2 * nb[2] + 1);
  return (total + 255) / 256;
}

void launch_gvec_count(const double* kfrac, const double* B_rows, double c, double ecut, const int* nb,
                       int* d_block_count, long long* d_block_off, cudaStream_t s) {
  const GvecParams p = make_params(kfrac, B_rows, c, ecut, nb);
  const long long blocks = gvec_num_blocks(nb);
  k0_gvec_count<<<(unsigned)blocks, 256, 0, s>>>(p, d_block_count);
  launch_scan(d_block_count, (int)blocks, d_block_off, s);
}

void launch_gvec_write(const double* kfrac, const double* B_rows, double c, double ecut, const int* nb,
                       const long long* d_block_off, int32_t* d_g_frac, long long capacity, cudaStream_t s) {
  const GvecParams p = make_params(kfrac, B_rows, c, ecut, nb);
  k0_gvec_write<<<(unsigned)gvec_num_blocks(nb), 256, 0, s>>>(p, d_block_off, d_g_frac, capacity);
}

}  // namespace bandup
