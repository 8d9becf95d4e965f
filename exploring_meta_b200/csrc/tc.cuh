// tc.cuh -- tcgen05 / TMEM / mbarrier PTX wrappers shared by the sm_100a tensor-core kernels.
#pragma once
#include "common.cuh"

namespace xm {

// ---- PTX wrappers ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Shared-memory matrix descriptor, canonical layouts without swizzle (16-byte units T = 4 fp32):
//   K-major : element (row, k) at start + (row/8)*SBO + (row%8)*16 + (k/4)*LBO + (k%4)*4 bytes
//   MN-major: element (mn,  k) at start + (mn/4)*SBO  + (mn%4)*4   + (k%8)*16  + (k/8)*LBO bytes
//   MN-major, layout 1 (SWIZZLE_128B_BASE32B -- the only MN-major form the hardware accepts for tf32; decoded
//   with scripts/probe_umma.cu): rows of 128 B = 32 mn-elements of one k; element (mn, k) at
//   start + (mn/32)*LBO + (k/4)*SBO + (k%4)*128 + (((mn%32)/8) ^ (row & 3))*32 + (mn%8)*4, where row =
//   absolute shared address >> 7 -- the XOR uses ABSOLUTE address bits, so a start address shifted by whole
//   rows addresses the same swizzled data.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout_type = 0) {
  uint64_t d = (uint64_t)(layout_type & 7) << 61;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
  return d;                                     // base_offset = 0, layout_type = SWIZZLE_NONE
}

// Instruction descriptor: kind::tf32 inputs, fp32 accumulate; a_mn / b_mn = 1 for MN-major operands.
__host__ __device__ constexpr uint32_t umma_idesc_tf32(uint32_t M, uint32_t N, uint32_t a_mn, uint32_t b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// One elected lane of a converged warp (the MMA / commit instructions are issued from warp-uniform code so
// that the compiler emits them without a per-thread serialisation loop).
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, %1;\n\t"
      "@px mov.s32 %0, 1;\n\t}"
      : "+r"(pred) : "r"(0xFFFFFFFFu));
  return pred != 0;
}

// Same instruction with the descriptors given as (lo, hi) 32-bit halves: only the start-address field in
// the low word changes between the MMAs of a tile, so the issuing thread spends one integer add per operand.
__device__ __forceinline__ void umma_tf32_lh(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                             uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}
__host__ __device__ constexpr uint32_t umma_desc_hi(uint32_t sbo_bytes, uint32_t layout_type = 0) {
  return ((sbo_bytes >> 4) & 0x3FFF) | (1u << 14) | ((layout_type & 7) << 29);
}
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr >> 4) & 0x3FFF) | (((lbo_bytes >> 4) & 0x3FFF) << 16);
}

__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }


// ---- flattened padded-pixel position arithmetic ----------------------------------------------------------------
// Exact unsigned division by a run-time constant d >= 2 for numerators < 2^31: q / d == umulhi(q, mul) >> sh
// (round-up magic number; verified exhaustively on the host side of the tests' shapes).  The producers map
// hundreds of staged rows per tile from a position to an address; two 2-instruction divisions per row keep every
// row's computation independent (instruction-level parallelism) instead of a serial carry chain.
struct FastDiv { uint32_t mul, sh; };
inline FastDiv make_fastdiv(uint32_t d) {
  uint32_t s = 0;
  while ((1ull << s) < d) ++s;                      // ceil(log2(d)), d >= 2
  FastDiv f;
  f.mul = (uint32_t)(((1ull << (31 + s)) / d) + 1ull);
  f.sh = s - 1;
  return f;
}
__device__ __forceinline__ uint32_t fastdiv(uint32_t q, FastDiv d) { return __umulhi(q, d.mul) >> d.sh; }

// Position sequence of one task: q = (img, r, c) with r in [0, H], c in [0, W]; row 0 and column W are padding.
struct PosMap {
  int n, H, W, Wp, HpWp;
  FastDiv dimg, drow;
};
inline PosMap make_posmap(int n, int H, int W) {
  PosMap m;
  m.n = n; m.H = H; m.W = W; m.Wp = W + 1; m.HpWp = (H + 1) * (W + 1);
  m.dimg = make_fastdiv((uint32_t)m.HpWp);
  m.drow = make_fastdiv((uint32_t)m.Wp);
  return m;
}
// Pixel index (img*H + y)*W + x of position q inside the task's [n][H][W] tensor, or -1 for padding / out of range.
__device__ __forceinline__ int pos_to_pixel(const PosMap& m, int q) {
  const uint32_t qq = (uint32_t)max(q, 0);
  const int img = (int)fastdiv(qq, m.dimg);
  const int rem = (int)qq - img * m.HpWp;
  const int r = (int)fastdiv((uint32_t)rem, m.drow);
  const int c = rem - r * m.Wp;
  const bool ok = q >= 0 && img < m.n && r >= 1 && c < m.W;
  return ok ? (img * m.H + r - 1) * m.W + c : -1;
}

// 256-bit read-only global load (sm_100: LDG.E.256): halves the number of load requests of the staging loops.
__device__ __forceinline__ void ldg256(const float* p, float4& a, float4& b) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}

// x ~= hi + lo with hi = x rounded to TF32 (nearest, ties away -- the same value cvt.rna.tf32.f32 produces, computed
// with two full-rate integer ops instead of the conversion pipe) and lo = x - hi (exact in fp32).
__device__ __forceinline__ void split_tf32_fast(const float4& v, float4& h, float4& l) {
  h.x = __uint_as_float((__float_as_uint(v.x) + 0x1000u) & 0xFFFFE000u); l.x = v.x - h.x;
  h.y = __uint_as_float((__float_as_uint(v.y) + 0x1000u) & 0xFFFFE000u); l.y = v.y - h.y;
  h.z = __uint_as_float((__float_as_uint(v.z) + 0x1000u) & 0xFFFFE000u); l.z = v.z - h.z;
  h.w = __uint_as_float((__float_as_uint(v.w) + 0x1000u) & 0xFFFFE000u); l.w = v.w - h.w;
}

}  // namespace xm
