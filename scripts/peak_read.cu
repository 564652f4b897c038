// peak_read.cu -- how fast can a B200 READ HBM?  (a) 128-bit ld.global.nc, 8 in flight per thread;
// (b) cp.async.bulk ring with consumers that only release the stages.  Not product code: the
// ceiling K2 is compared with.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o peak_read peak_read.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) k_ldg(const float4* __restrict__ p, size_t n, float* out) {
  float acc = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 7 * stride < n; i += 8 * stride) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                   : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(p + i + u * stride));
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
  }
  if (acc == 123.456f) *out = acc;
}

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int STAGES, int STAGE_BYTES>
__global__ void __launch_bounds__(288) k_bulk(const char* __restrict__ p, size_t bytes, float* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + (size_t)STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(full + s)), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(empty + s)), "r"(8));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const size_t n_tiles = bytes / STAGE_BYTES;
  auto wait = [](uint64_t* b, uint32_t par) {
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n"
                 ::"r"(s32(b)), "r"(par) : "memory");
  };
  if (tid >= 256) {
    if (tid == 256) {
      uint64_t pol;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
      int s = 0; uint32_t ph = 0;
      for (size_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        wait(empty + s, ph ^ 1u);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(full + s)), "r"(STAGE_BYTES) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                     ::"r"(s32(sm + (size_t)s * STAGE_BYTES)), "l"(p + t * STAGE_BYTES), "r"(STAGE_BYTES), "r"(s32(full + s)), "l"(pol) : "memory");
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
    }
    return;
  }
  int s = 0; uint32_t ph = 0; float acc = 0.f;
  for (size_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    wait(full + s, ph);
    acc += reinterpret_cast<const float*>(sm + (size_t)s * STAGE_BYTES)[tid];
    __syncwarp();
    if ((tid & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(empty + s)) : "memory");
    if (++s == STAGES) { s = 0; ph ^= 1u; }
  }
  if (acc == 123.456f) *out = acc;
}

template <typename F> float time_ms(F f, int reps) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); f(); cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps;
}

int main(int argc, char** argv) {
  // reps per configuration: 5 = a burst of ~7 ms; 400 = ~0.5 s, long enough to run into the board's power cap
  // like bench.py's timed region does
  const int reps = argc > 1 ? atoi(argv[1]) : 5;
  const size_t bytes = (size_t)9600 << 20;  // ~10 GB >> L2
  char* d; float* out;
  cudaMalloc(&d, bytes); cudaMalloc(&out, 4); cudaMemset(d, 1, bytes);
  for (int cta = 2; cta <= 8; cta *= 2) {
    float ms = time_ms([&] { k_ldg<<<148 * cta, 256>>>((const float4*)d, bytes / 16, out); }, reps);
    printf("ldg.128 x8  ctas/SM=%d  %.1f GB/s\n", cta, bytes / ms / 1e6);
  }
  auto run_bulk = [&](auto kern, int stages, int sb, int cta) {
    size_t smem = (size_t)stages * sb + 16 * stages + 128;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    float ms = time_ms([&] { kern<<<148 * cta, 288, smem>>>(d, bytes, out); }, reps);
    printf("bulk ring   ctas/SM=%d stages=%d stage=%d KB  %.1f GB/s   (%s)\n", cta, stages, sb >> 10, bytes / ms / 1e6,
           cudaGetErrorString(cudaGetLastError()));
  };
  run_bulk(k_bulk<3, 16384>, 3, 16384, 2);
  run_bulk(k_bulk<4, 16384>, 4, 16384, 2);
  run_bulk(k_bulk<4, 16384>, 4, 16384, 3);
  run_bulk(k_bulk<6, 16384>, 6, 16384, 2);
  run_bulk(k_bulk<3, 32768>, 3, 32768, 2);
  run_bulk(k_bulk<2, 65536>, 2, 65536, 1);
  run_bulk(k_bulk<3, 65536>, 3, 65536, 1);
  run_bulk(k_bulk<6, 32768>, 6, 32768, 1);
  run_bulk(k_bulk<12, 16384>, 12, 16384, 1);
  return 0;
}
