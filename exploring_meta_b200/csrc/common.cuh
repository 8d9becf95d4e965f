// common.cuh -- shared helpers of libxmeta (sm_100a).  See include/xmeta.h for the ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>
#include "../../include/xmeta.h"

namespace xm {

// ---- error slot + launch counter (the only global mutable state) --------------------------------
extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

inline int launched(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail((int)e, "%s: %s", what, cudaGetErrorString(e));
  return 0;
}

#define XM_REQUIRE(cond, ...) do { if (!(cond)) return xm::fail(-1, __VA_ARGS__); } while (0)
#define XM_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return xm::fail((int)e_, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)

inline int geom_ok(const XmBlockGeom& g) {
  if (g.tasks <= 0 || g.n <= 0 || g.cin <= 0 || g.cout <= 0 || g.hin <= 0 || g.win <= 0) return 0;
  if (g.stride != 1 && g.stride != 2) return 0;
  int hz = (g.hin + 2 - 3) / g.stride + 1, wz = (g.win + 2 - 3) / g.stride + 1;
  if (g.hz != hz || g.wz != wz) return 0;
  if (g.pool) { if (g.hp != g.hz / 2 || g.wp != g.wz / 2 || g.hp <= 0 || g.wp <= 0) return 0; }
  else if (g.hp != g.hz || g.wp != g.wz) return 0;
  return 1;
}

int num_sms();
// CTAs of `kernel` (threads per CTA, dynamic shared memory) that are resident on the whole GPU at once (SMs x
// occupancy): grids are sized to at most this, in one wave -- a grid a few CTAs larger runs a second wave.
int wave_ctas(const void* kernel, int threads, size_t smem);

// ---- device helpers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t f2tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

// D(16x8) += A(16x8, row) * B(8x8, col), TF32 inputs, fp32 accumulate.
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// Error-compensated TF32 split: x ~= hi + lo with hi = rna_tf32(x), lo = x - hi (exact in fp32; the
// tensor core reads the top 19 bits of lo).  a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = f2tf32(x);
  lo = __float_as_uint(x - __uint_as_float(hi));
}

// Packed fp32 FMA (sm_100 FFMA2): (d0, d1) += a * (b0, b1) with the scalar a broadcast -- one issue slot for two
// FMAs; the plain 3-register FFMA issues at half rate on Blackwell, so FMA-bound CUDA-core loops use this.
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %2};\n\tmov.b64 rb, {%3, %4};\n\tmov.b64 rc, {%0, %1};\n\t"
      "fma.rn.f32x2 rc, ra, rb, rc;\n\tmov.b64 {%0, %1}, rc;\n\t}"
      : "+f"(d0), "+f"(d1) : "f"(a), "f"(b0), "f"(b1));
}

// 4-byte asynchronous global -> shared copy (LDGSTS); src_bytes = 0 writes zero (out-of-range halo elements).
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace xm
