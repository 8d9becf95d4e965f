// probe_sm_ingest.cu -- how fast can ONE CTA per SM pull a contiguous stream out of HBM/L2?
//   mode 0: 224 threads, 256-bit loads, D units of 4 loads per thread in flight (register prefetch, like the tcgen05 producers)
//   mode 1: one warp issues 1-D bulk (TMA) copies of CH bytes into a ring of S stages; 7 warps consume (read) them
// Every CTA streams its own contiguous region of `per_cta` bytes (persistent-grid pattern of conv_tc / wgrad_tc).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 scripts/probe_sm_ingest.cu -o scripts/probe_sm_ingest.bin
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ldg256(const float* p, float4& a, float4& b) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int D>
__global__ void __launch_bounds__(224, 1) ldg_stream(const float* base, long long per_cta_floats, float* out) {
  const float* p = base + (long long)blockIdx.x * per_cta_floats;
  const int tid = threadIdx.x;
  const long long unit = 224LL * 4 * 8;                 // floats per unit: 224 threads x 4 loads x 8 floats (28 KB)
  const long long nunits = per_cta_floats / unit;
  float4 v[D][8];
  float acc = 0.f;
  auto issue = [&](long long u, float4 (&r)[8]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) ldg256(p + u * unit + (long long)(k * 224 + tid) * 8, r[2 * k], r[2 * k + 1]);
  };
#pragma unroll
  for (int d = 0; d < D; ++d) if (d < nunits) issue(d, v[d]);
  for (long long u = 0; u < nunits; u += D) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      if (u + d < nunits) {
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += v[d][k].x + v[d][k].y + v[d][k].z + v[d][k].w;
        if (u + d + D < nunits) issue(u + d + D, v[d]);
      }
    }
  }
  if (acc == 123.456f) out[0] = acc;
}

template <int S, int CH>
__global__ void __launch_bounds__(256, 1) tma_stream(const float* base, long long per_cta_floats, float* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);          // full[S], free[S]
  unsigned char* ring = smem + 1024;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const char* src = reinterpret_cast<const char*>(base + (long long)blockIdx.x * per_cta_floats);
  constexpr int STAGE = 8 * CH;                                   // 8 chunks per stage
  const long long nunits = per_cta_floats * 4 / STAGE;
  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bars + s)), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bars + S + s)), "r"(7));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto wait = [&](uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
      asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                   : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
  };
  float acc = 0.f;
  if (warp == 7) {
    for (long long u = 0; u < nunits; ++u) {
      const int s = (int)(u % S);
      if (u >= S) wait(smem_u32(bars + S + s), (uint32_t)(((u / S) - 1) & 1));
      if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bars + s)), "r"(STAGE) : "memory");
      __syncwarp();
      if (lane < 8)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(ring + (size_t)s * STAGE + lane * CH)), "l"(src + u * STAGE + (long long)lane * CH), "r"(CH),
                       "r"(smem_u32(bars + s)) : "memory");
    }
  } else {
    for (long long u = 0; u < nunits; ++u) {
      const int s = (int)(u % S);
      wait(smem_u32(bars + s), (uint32_t)((u / S) & 1));
      const float4* r = reinterpret_cast<const float4*>(ring + (size_t)s * STAGE);
      for (int i = tid; i < STAGE / 16; i += 224) { const float4 v = r[i]; acc += v.x + v.y + v.z + v.w; }
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bars + S + s)) : "memory");
    }
  }
  if (acc == 123.456f) out[0] = acc;
}

template <typename F> float timeit(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
  const int ctas = 148;
  const long long per_cta = 16LL << 20;                         // 16 MB per CTA: 2.4 GB total (>> L2)
  float* d; cudaMalloc(&d, per_cta * ctas + (1 << 20)); cudaMemset(d, 0, per_cta * ctas);
  float* out; cudaMalloc(&out, 16);
  const long long pf = per_cta / 4;
  const double gb = (double)per_cta * ctas / 1e9;
  auto rep = [&](const char* name, float ms) { printf("%-46s %8.3f ms  %7.1f GB/s total  %6.1f GB/s per SM\n", name, ms, gb / (ms / 1e3), gb / (ms / 1e3) / ctas); };
  rep("ldg256, 1 unit (28 KB) in flight per SM", timeit([&] { ldg_stream<1><<<ctas, 224>>>(d, pf, out); }));
  rep("ldg256, 2 units in flight", timeit([&] { ldg_stream<2><<<ctas, 224>>>(d, pf, out); }));
  rep("ldg256, 4 units in flight", timeit([&] { ldg_stream<4><<<ctas, 224>>>(d, pf, out); }));
  cudaFuncSetAttribute(tma_stream<2, 5376>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(tma_stream<3, 5376>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(tma_stream<4, 5376>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(tma_stream<4, 2048>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  rep("bulk copies 8 x 5376 B per stage, 2 stages", timeit([&] { tma_stream<2, 5376><<<ctas, 256, 1024 + 2 * 8 * 5376>>>(d, pf, out); }));
  rep("bulk copies 8 x 5376 B per stage, 3 stages", timeit([&] { tma_stream<3, 5376><<<ctas, 256, 1024 + 3 * 8 * 5376>>>(d, pf, out); }));
  rep("bulk copies 8 x 5376 B per stage, 4 stages", timeit([&] { tma_stream<4, 5376><<<ctas, 256, 1024 + 4 * 8 * 5376>>>(d, pf, out); }));
  rep("bulk copies 8 x 2048 B per stage, 4 stages", timeit([&] { tma_stream<4, 2048><<<ctas, 256, 1024 + 4 * 8 * 2048>>>(d, pf, out); }));
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
