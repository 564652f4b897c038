// synth.cu -- on-device generator of synthetic plane-wave coefficients.
//
// The big configurations (1000-atom Si-Ge: 2.4 GB per SC k-point, 480 GB per
// job) cannot be shipped as WAVECAR files, so benchmarks and full-size tests
// materialise any (seed, ikpt, band, element) coefficient from a counter-based
// hash.  The host twin, bit for bit, is bo_synth_fill in oracle/bandup_oracle.c
// (used only by tests and the CPU baseline).
#include "kernels.cuh"

namespace bandup {

namespace {

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__global__ void __launch_bounds__(256)
k_synth_fill(char* __restrict__ coeffs, long long row_elems, int band0, int n_bands,
             long long stride_bytes, uint64_t seed, uint64_t ikpt) {
  const int band = band0 + blockIdx.y;
  if (band >= n_bands) return;
  float2* row = reinterpret_cast<float2*>(coeffs + (long long)band * stride_bytes);
  const uint64_t key = splitmix64(seed ^ (ikpt * 0xD1B54A32D192ED03ull) ^
                                  ((uint64_t)band * 0x8CB92BA72F3D8DD7ull));
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < row_elems; e += stride) {
    const uint64_t h = splitmix64(key + (uint64_t)e);
    const float re = (float)((int32_t)(h & 0xFFFFFFFFu)) * (1.0f / 2147483648.0f);
    const float im = (float)((int32_t)(h >> 32)) * (1.0f / 2147483648.0f);
    const float env = 1.0f / (float)(1 + (int)(e % 7));
    row[e] = make_float2(re * env, im * env);
  }
}

}  // namespace

void launch_synth_fill(void* d_coeffs, long long row_elems, int n_bands,
                       long long stride_bytes, uint64_t seed, uint64_t ikpt,
                       cudaStream_t s) {
  if (n_bands <= 0 || row_elems <= 0) return;
  long long bx = (row_elems + 256 * 8 - 1) / (256 * 8);
  if (bx > 64) bx = 64;
  // gridDim.y is limited to 65535
  for (int b0 = 0; b0 < n_bands; b0 += 32768) {
    const int nb = n_bands - b0 < 32768 ? n_bands - b0 : 32768;
    dim3 grid((unsigned)bx, (unsigned)nb);
    k_synth_fill<<<grid, 256, 0, s>>>(static_cast<char*>(d_coeffs), row_elems, b0, n_bands,
                                      stride_bytes, seed, ikpt);
  }
}

}  // namespace bandup
