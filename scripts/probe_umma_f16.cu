// probe_umma_f16.cu -- decodes how tcgen05.mma kind::f16 addresses an MN-major shared-memory A operand.
// A region (32 KB = 16384 halfs) is filled with its own half index (two runs: low 10 bits / high bits, both exactly
// representable in fp16); B (K-major, no swizzle) is a 16 x 16 identity selector, so D[m][n] = A[m][k = n] = the index read.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I exploring_meta_b200/csrc scripts/probe_umma_f16.cu -o /tmp/probe_f16
#include <cstdio>
#include <vector>
#include <cuda_fp16.h>
#include "tc.cuh"
using namespace xm;

struct Cfg { uint32_t lbo, sbo, start_off, layout; int part; };

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

__global__ void probe(Cfg c, float* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __half* A = reinterpret_cast<__half*>(smem);               // 32 KB region = 16384 halfs
  __half* B = A + 16384;                                     // K-major identity: [kgroup 2][n 16][8]
  uint64_t* bar = reinterpret_cast<uint64_t*>(B + 2 * 16 * 8);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 16384; i += blockDim.x) A[i] = __float2half(c.part ? (float)(i >> 10) : (float)(i & 1023));
  for (int i = tid; i < 256; i += blockDim.x) {
    const int kg = i / 128, n = (i / 8) % 16, e = i % 8, k = kg * 8 + e;
    B[i] = __float2half((n == k) ? 1.f : 0.f);
  }
  if (tid == 0) { mbar_init(smem_u32(bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(32) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async(); tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = *slot;
  if (tid == 0) {
    uint64_t ad = umma_desc(smem_u32(A) + c.start_off, c.lbo, c.sbo);
    ad |= (uint64_t)(c.layout & 7) << 61;
    const uint64_t bd = umma_desc(smem_u32(B), 256u, 128u);      // K-major: n rows 16 B apart, k groups 16*16 B apart
    // kind::f16, fp16 inputs, fp32 accumulate, A MN-major (bit 15), B K-major, N = 16, M = 128
    const uint32_t idesc = (1u << 4) | (1u << 15) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
    umma_f16(tm, ad, bd, idesc, 0u);
    umma_commit(smem_u32(bar));
  }
  mbar_wait(smem_u32(bar), 0);
  tc_fence_after();
  float v[32];
  tmem_ld32(tm + ((uint32_t)(warp * 32) << 16), v);
  for (int n = 0; n < 16; ++n) out[(warp * 32 + (tid & 31)) * 16 + n] = v[n];
  tc_fence_before(); __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(32) : "memory");
}

int main() {
  std::vector<Cfg> cfgs = {
    {128, 2048, 0, 0},     // no swizzle: LBO 128 (k atoms), SBO 2048 (mn groups of 8)
    {2048, 128, 0, 0},     // roles swapped
    {64, 512, 0, 4},       // SWIZZLE_64B: rows of 64 B = 32 mn elements?
    {4096, 512, 0, 4},
    {64, 512, 64, 4},      // start + 1 row
    {64, 512, 128, 4},     // start + 2 rows
    {4096, 1024, 0, 2},    // SWIZZLE_128B
    {128, 1024, 0, 2},
    {4096, 256, 0, 6},     // SWIZZLE_32B
    {4096, 512, 0, 1},     // 128B_BASE32B
  };
  float* d; cudaMalloc(&d, 128 * 16 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (auto c : cfgs) {
    std::vector<float> lo(2048), hi(2048);
    c.part = 0; probe<<<1, 128, 32768 + 1024 + 64>>>(c, d); cudaMemcpy(lo.data(), d, 8192, cudaMemcpyDeviceToHost);
    c.part = 1; probe<<<1, 128, 32768 + 1024 + 64>>>(c, d); cudaMemcpy(hi.data(), d, 8192, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaDeviceSynchronize();
    printf("cfg lbo=%u sbo=%u start=%u layout=%u : %s\n", c.lbo, c.sbo, c.start_off, c.layout, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    const int rows[] = {0, 1, 7, 8, 16, 24, 31, 32, 33, 64, 96};
    for (int m : rows) {
      printf("  m=%3d :", m);
      for (int n = 0; n < 16; ++n) printf(" %5d", (int)(hi[m * 16 + n] * 1024 + lo[m * 16 + n]));
      printf("   (half index of A[m][k=0..15])\n");
    }
  }
  return 0;
}
