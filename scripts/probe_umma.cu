// probe_umma.cu -- decodes how tcgen05.mma addresses a no-swizzle shared-memory operand.
// A region is filled with its own float index (split in two exactly-representable parts over two runs);
// B (K-major, validated layout) is an identity selector, so D[m][n] = A[m][k = n] = the index that was read.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I exploring_meta_b200/csrc scripts/probe_umma.cu -o /tmp/probe_umma
#include <cstdio>
#include <vector>
#include "tc.cuh"
using namespace xm;

struct Cfg { int a_mn; uint32_t lbo, sbo; uint32_t start_off; uint32_t layout, base_off; int part; };

__global__ void probe(Cfg c, float* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* A = reinterpret_cast<float*>(smem);                 // 64 KB region = 16384 floats
  float* B = A + 16384;                                      // K-major identity: [kgroup 2][n 32][4]
  uint64_t* bar = reinterpret_cast<uint64_t*>(B + 2 * 32 * 4);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 16384; i += blockDim.x) A[i] = c.part ? (float)(i >> 10) : (float)(i & 1023);
  for (int i = tid; i < 256; i += blockDim.x) {
    const int kg = i / 128, n = (i / 4) % 32, e = i % 4, k = kg * 4 + e;
    B[i] = (n == k) ? 1.f : 0.f;
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
    ad |= (uint64_t)(c.base_off & 7) << 49;
    ad |= (uint64_t)(c.layout & 7) << 61;
    const uint64_t bd = umma_desc(smem_u32(B), 512u, 128u);
    umma_tf32(tm, ad, bd, umma_idesc_tf32(128, 32, c.a_mn, 0), 0u);
    umma_commit(smem_u32(bar));
  }
  mbar_wait(smem_u32(bar), 0);
  tc_fence_after();
  float v[32];
  tmem_ld32(tm + ((uint32_t)(warp * 32) << 16), v);
  for (int n = 0; n < 8; ++n) out[(warp * 32 + (tid & 31)) * 8 + n] = v[n];
  tc_fence_before(); __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(32) : "memory");
}

int main() {
  std::vector<Cfg> cfgs = {
    {1, 4096, 512, 0, 1, 0},       // SWIZZLE_128B_BASE32B, dense 128 B rows
    {1, 4096, 512, 128, 1, 0},     // start + 1 row
    {1, 4096, 512, 128, 1, 1},     // start + 1 row, base_offset 1
    {1, 4096, 512, 256, 1, 0},     // start + 2 rows
    {1, 4096, 512, 640, 1, 0},     // start + 5 rows
    {1, 0, 512, 0, 1, 0},          // LBO = 0
    {1, 16, 512, 0, 1, 0},         // LBO = 16
  };
  float* d; cudaMalloc(&d, 128 * 8 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (auto c : cfgs) {
    std::vector<float> lo(1024), hi(1024);
    c.part = 0; probe<<<1, 128, 16384 * 4 + 1024 + 64>>>(c, d); cudaMemcpy(lo.data(), d, 4096, cudaMemcpyDeviceToHost);
    c.part = 1; probe<<<1, 128, 16384 * 4 + 1024 + 64>>>(c, d); cudaMemcpy(hi.data(), d, 4096, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaDeviceSynchronize();
    printf("cfg a_mn=%d lbo=%u sbo=%u start=%u layout=%u base_off=%u : %s\n", c.a_mn, c.lbo, c.sbo, c.start_off, c.layout, c.base_off, cudaGetErrorString(e));
    const int rows[] = {0, 1, 8, 16, 24, 31, 32, 64, 96};
    for (int m : rows) {
      printf("  m=%3d :", m);
      for (int n = 0; n < 8; ++n) printf(" %6d", (int)(hi[m * 8 + n] * 1024 + lo[m * 8 + n]));
      printf("   (float index of A[m][k=0..7])\n");
    }
  }
  return 0;
}
