// wgrad_tc.cu -- tcgen05 / TMEM implementation of xm_wgrad for the 32-channel stride-1 layers.
//
// gW[kh][kw][ci][co] = sum over positions q of g[q][co] * x[q + (kh-1)*Wp + (kw-1)][ci], on the same flattened
// padded position sequence as conv_tc.cu.  The reduction dimension of the GEMM is the POSITION axis, so both
// operands are MN-major: position rows of 128 B (= the 32 channels of one position, i.e. the natural NHWC row)
// in the 128B-swizzle / 32B-base layout (the swizzle is a function of absolute address bits, so whole-row
// shifts of a descriptor start address or of LBO address the same swizzled data).
//
// ALL NINE TAPS COME OUT OF ONE MMA (kind::tf32, M = 128, N = 96, K = 8 position rows).  With the substitution
// q'' = q + (kh-1)*Wp the sum is  sum_q'' x[q'' + kw - 1][ci] * g[q'' - (kh-1)*Wp][co]  and
//   * A = x tile: the four 32-lane groups of M are the x rows shifted by LBO = one position row, i.e. kw = 0, 1, 2
//     (group 3 is unused);
//   * B = g halo: the three 32-column groups of N are the g rows shifted by LBO = Wp position rows, i.e. kernel
//     rows kh = 2, 1, 0;
//   D[(kw, ci)][(2-kh, co)] += sum_k x[row k0 + k + kw][ci] * g[row k0 + k + (2-kh)*Wp][co].
// A k-step therefore costs 3 MMAs (the 3xTF32 expansion terms) and 7 KB of shared-memory operand reads instead of
// 27 MMAs / 135 KB for the one-tap-per-MMA form.  TMEM lane quarter kw / column block 2-kh holds tap (kh, kw),
// lane = ci, column = co.  Two accumulator sets (2 x 96 columns) let the drain of tile i overlap the MMAs of tile
// i+1.  Accumulators are drained after every tile into fp32 registers with round-to-nearest adds (the tensor
// core's own accumulation truncates), and each CTA writes one partial block per task split;
// wgrad_reduce_kernel (wgrad.cu) reduces the splits in double and applies the fused SGD / outer-recursion
// epilogue.
//
// FP16-split variant (template parameter F16, xm_set_precision(2)) -- where the split pays: this kernel is bound by
// shared-memory traffic (producer stores 88 KB + MMA operand reads 336 KB per tile against 128 B/clk; drain is light).
//   * operands: position rows of 64 B (32 channels of fp16) in the SWIZZLE_64B MN-major layout, decoded on hardware
//     with scripts/probe_umma_f16.cu: element (mn, k) at start + (mn/32)*LBO + (k/8)*SBO + (k%8)*64 +
//     (((mn%32)/8) ^ ((row >> 1) & 3))*16 + (mn%8)*2 with row = ABSOLUTE shared address >> 6 -- so, exactly as in the
//     TF32 layout, LBO = one position row makes the M groups the tile shifted by kw rows, LBO = Wp rows makes the N
//     groups the halo shifted by kernel rows, and a start address shifted by whole rows addresses the same data;
//   * x * s = hi + lo with hi = fp16(x * s), lo = fp16(x * s - hi): a power-of-two scale s per staged tile and operand
//     (from the tile's absolute maximum: producer warps post their maxima and arrive on an mbarrier BEFORE waiting
//     for the stage) brings the maximum to [2^14, 2^15), so lo is a normal fp16 for every element within 2^17 of the
//     maximum and the absolute error of the rest is 2^-39 of it -- no separate correction accumulator is needed;
//   * K = 16 position rows per MMA: 8 K steps x 3 terms = 24 MMAs per (tile, pair) unit instead of 48; half the staged
//     bytes; every unit is drained on its own (its scale is its own) with the power-of-two un-scaling in the drain.
// Measured (scripts/diag_wgrad.py, 42x42, 32 tasks): accuracy against fp64 4.9e-7 (3xTF32: 8.8e-7); 193 us against 197 us.
// With the global loads compiled out the FP16 kernel takes 83 us and the TF32 one 145 us: the split halves the
// tensor / shared-memory side, but what bounds the kernel is the producers' global -> register -> convert -> shared
// path (address arithmetic, 256-bit loads, conversions in 7 warps).  A per-thread cp.async ring three units deep
// (more bytes in flight, raw fp32 rows in shared memory) was 25 % SLOWER -- the extra shared-memory round trip costs
// more than the latency it hides.  Bulk (TMA) row copies issued by one loader warp into the same ring (XM_WT_TMA=1:
// cp.async.bulk per image row, padding positions zeroed by the loader, producers convert shared -> shared) are correct
// and slower still (231 us): per unit the producers then wait 2300 cycles for the copies although they are issued
// three units ahead -- the memory system delivers ~1.9 TB/s to these 148 persistent streams whatever the request depth.
#include <cuda_fp16.h>
#include <stdlib.h>
#include "tc.cuh"

namespace xm {

extern int g_precise;                        // conv.cu: 2 selects the FP16-split variant

constexpr int WT_DRAINERS = 128;             // warps 0-3: warp kw < 3 drains TMEM lane quarter kw (warp 3 idles)
constexpr int WT_PRODUCERS = 224;            // warps 4-10 (default build: 7 producer warps)
constexpr int WT_THREADS = WT_DRAINERS + WT_PRODUCERS + 32;   // + the MMA-issuing warp 11 (12 warps: 168 registers)
constexpr int WT_NPW_TMA = 11;               // producer warps of the bulk-copy (TMA) variant: converters only, 40 registers of state
constexpr int WT_TMEM_COLS = 256;            // 2 sets x 3 accumulators x 32 columns (192) -> next power of two
constexpr uint32_t WT_IDESC = umma_idesc_tf32(128, 96, 1, 1);   // A and B MN-major, N = 3 kernel rows x 32 cout
// kind::f16: fp16 inputs, fp32 accumulate, A and B MN-major
constexpr uint32_t WT_IDESC_F16 = (1u << 4) | (1u << 15) | (1u << 16) | ((96u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void umma_f16_lh(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                            uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}
// 1-D bulk (TMA) copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// power-of-two scale 2^k that brings a maximum magnitude m into [2^14, 2^15); k = 0 for m = 0
__device__ __forceinline__ int wt_scale_exp(float m) {
  if (!(m > 0.f)) return 0;
  const int k = 14 - (int)((__float_as_uint(m) >> 23) & 0xffu) + 127;
  return max(-100, min(100, k));
}
__device__ __forceinline__ float wt_exp2i(int k) { return __uint_as_float((uint32_t)(127 + k) << 23); }
__device__ __forceinline__ float wt_absmax(const float4& v) {
  return fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
}
// 8 floats -> 8 fp16 hi + 8 fp16 lo (16 bytes each).  hi is rounded to 11 significant bits in fp32 first (two integer
// ops, like split_tf32_fast), so the packing conversion is exact for normal fp16 values and lo = x*s - hi is exact.
__device__ __forceinline__ void wt_split_f16(const float4& a, const float4& b, float sc, uint4& hi, uint4& lo) {
  const float x[8] = {a.x * sc, a.y * sc, a.z * sc, a.w * sc, b.x * sc, b.y * sc, b.z * sc, b.w * sc};
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float h0 = __uint_as_float((__float_as_uint(x[2 * i]) + 0x1000u) & 0xFFFFE000u);
    const float h1 = __uint_as_float((__float_as_uint(x[2 * i + 1]) + 0x1000u) & 0xFFFFE000u);
    const __half2 hh = __floats2half2_rn(h0, h1);
    const __half2 ll = __floats2half2_rn(x[2 * i] - h0, x[2 * i + 1] - h1);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    l[i] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

struct WgradTcK {
  int tasks, n, H, W, Hp, Wp, Q;
  PosMap pm;
  int tiles_per_task, splits, npairs, ctas;
  int Rx, Rg, xbuf, gbuf;             // staged x / g rows per tile, bytes of one x / g buffer (hi or lo)
  int off_x0, off_x1, off_g0, off_g1, off_bar;   // shared-memory byte offsets
  int off_raw, raw_stage, raw_stages;            // FP16 variant: ring of raw fp32 tiles landed by bulk (TMA) copies; 0 stages = off
  const float* x[2];
  const float* g[2];
  float* partial;                     // [task][split][9*32][32]
  int x_cs, x_co, g_cs, g_co;         // floats per position of x / g and the first channel of this launch's 32-channel block
};

template <bool F16, int NPW>
__global__ void __launch_bounds__(WT_DRAINERS + NPW * 32 + 32, 1) wgrad_tc_kernel(const WgradTcK p) {
  constexpr int NPROD = NPW * 32, NTHREADS = WT_DRAINERS + NPROD + 32, MMA_WARP = 4 + NPW;
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int xset = p.xbuf, gset = p.gbuf;
  constexpr int ROWB = F16 ? 64 : 128;                 // bytes of one staged position row
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  // FP16 variant: per-unit maxima of the x tile / g halo (float bits), the unit's combined scale exponent, and the
  // "every producer warp has posted its maxima" barriers
  uint32_t* smaxx = reinterpret_cast<uint32_t*>(bars + 10);      // [4]
  uint32_t* smaxg = smaxx + 4;                                   // [4]
  int* uexp = reinterpret_cast<int*>(smaxg + 4);                 // [4]
  const uint32_t bar_max = smem_u32(bars + 16);                  // [2]
  const uint32_t bar_rfull = smem_u32(bars + 18), bar_rfree = smem_u32(bars + 21);   // [3] each: raw ring (TMA) full / free
  const int NR = p.raw_stages;
  const uint32_t bar_full = smem_u32(bars), bar_sfree = smem_u32(bars + 2), bar_tfull = smem_u32(bars + 4),
                 bar_tfree = smem_u32(bars + 6);

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_full + 8 * s, NPROD);
      mbar_init(bar_sfree + 8 * s, 1);
      mbar_init(bar_tfull + 8 * s, 1);
      mbar_init(bar_tfree + 8 * s, 96);
      if (F16) mbar_init(bar_max + 8 * s, NPW);
    }
    if (F16) {
      for (int i = 0; i < 4; ++i) { smaxx[i] = 0u; smaxg[i] = 0u; }
      for (int i = 0; i < 3; ++i) { mbar_init(bar_rfull + 8 * i, 1); mbar_init(bar_rfree + 8 * i, NPW); }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(WT_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // rows past Rx (read only by the unused 4th lane group) must hold finite values
  for (int i = tid; i < (xset - p.Rx * ROWB) / 16; i += NTHREADS) {
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const size_t o = (size_t)p.Rx * ROWB + (size_t)i * 16;
    *reinterpret_cast<float4*>(smem + p.off_x0 + o) = zero;
    *reinterpret_cast<float4*>(smem + p.off_x0 + xset + o) = zero;
    *reinterpret_cast<float4*>(smem + p.off_x1 + o) = zero;
    *reinterpret_cast<float4*>(smem + p.off_x1 + xset + o) = zero;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // persistent CTA over the flattened (task, tile) list: CTA c owns tiles [c*G/n, (c+1)*G/n) -- every SM gets the
  // same share whatever the task count.  The pipeline streams straight through task boundaries; only the drain
  // flushes its register accumulators when the task changes.  Partial slot of (CTA c, task t) = c - first CTA of t.
  const long long G = (long long)p.tasks * p.tiles_per_task;
  const int g_lo = (int)(G * blockIdx.x / gridDim.x), g_hi = (int)(G * (blockIdx.x + 1) / gridDim.x);
  const int ntiles = g_hi - g_lo;
  const int nunits = ntiles * p.npairs;

  if (warp < 4) {
    // ===================================== accumulator drain ============================================
    // warp kw (0..2) owns TMEM lane quarter kw: lane = cin, column block j = 2 - kh of 32 cout columns.  The tile's
    // accumulators are added (round-to-nearest) into fp32 master accumulators held in registers.
    if (warp < 3) {
      float macc[3][32];
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int c = 0; c < 32; ++c) macc[a][c] = 0.f;
      auto flush = [&](int task) {
        const int first = (int)((((long long)task * p.tiles_per_task + 1) * gridDim.x + G - 1) / G) - 1;   // first CTA of the task
        float* P = p.partial + ((long long)task * p.splits + ((int)blockIdx.x - first)) * 9 * 32 * 32;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          float4* dst = reinterpret_cast<float4*>(P + (((2 - j) * 3 + warp) * 32 + lane) * 32);   // tap (kh = 2-j, kw)
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            dst[c] = make_float4(macc[j][4 * c], macc[j][4 * c + 1], macc[j][4 * c + 2], macc[j][4 * c + 3]);
            macc[j][4 * c] = macc[j][4 * c + 1] = macc[j][4 * c + 2] = macc[j][4 * c + 3] = 0.f;
          }
        }
      };
      int cur_task = g_lo / p.tiles_per_task;
      if (F16) {
        // every (tile, pair) unit has its own power-of-two scale: drained on its own, un-scaled while adding
#ifdef XM_TC_TIMING
        long long tw = 0, td = 0, t0;
#endif
        for (int u = 0; u < nunits; ++u) {
          const int set = u & 1, it = u / p.npairs;
          const int t_task = (g_lo + it) / p.tiles_per_task;
          if (t_task != cur_task) { flush(cur_task); cur_task = t_task; }
#ifdef XM_TC_TIMING
          t0 = clock64();
#endif
          mbar_wait(bar_tfull + 8 * set, (u >> 1) & 1);
#ifdef XM_TC_TIMING
          tw += clock64() - t0; t0 = clock64();
#endif
          tc_fence_after();
          const float unscale = wt_exp2i(-uexp[u & 3]);
          const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(set * 96);
          float v[32];
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            tmem_ld32(taddr + (uint32_t)(kh * 32), v);
#pragma unroll
            for (int c = 0; c < 32; ++c) macc[kh][c] = fmaf(v[c], unscale, macc[kh][c]);
          }
          tc_fence_before();
          mbar_arrive(bar_tfree + 8 * set);
#ifdef XM_TC_TIMING
          td += clock64() - t0;
#endif
        }
#ifdef XM_TC_TIMING
        if (blockIdx.x == 0 && tid == 0) printf("drain: per unit wait-tfull %lld drain %lld (units %d)\n", tw / nunits, td / nunits, nunits);
#endif
      } else
      for (int it = 0; it < ntiles; ++it) {
        const int set = it & 1;
        const int t_task = (g_lo + it) / p.tiles_per_task;
        if (t_task != cur_task) { flush(cur_task); cur_task = t_task; }
        mbar_wait(bar_tfull + 8 * set, (it >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(set * 96);
        float v[32];
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          tmem_ld32(taddr + (uint32_t)(kh * 32), v);
#pragma unroll
          for (int c = 0; c < 32; ++c) macc[kh][c] += v[c];
        }
        tc_fence_before();
        mbar_arrive(bar_tfree + 8 * set);
      }
      if (ntiles > 0) flush(cur_task);
    } else if (F16 && NR > 0) {
      // ===================================== bulk-copy loader (warp 3) =====================================
      // Positions differ from pixels only by the padding column of every row and the padding row of every image, so
      // the rows a unit needs are a handful of contiguous pixel runs: one 1-D bulk copy (TMA) per image row lands them
      // as raw fp32 position rows in a ring NR units deep; the padding positions of the slot are zeroed by this warp.
      // The producer warps then only convert shared -> shared: no position arithmetic, no global loads in 224 threads.
      const int Qtask = p.Q;
      for (int u = 0; u < nunits; ++u) {
        const int st = u % NR;
        if (u >= NR) mbar_wait(bar_rfree + 8 * st, ((u / NR) - 1) & 1);
        const int it = u / p.npairs, pair = u - it * p.npairs;
        const int gtile = g_lo + it, task = gtile / p.tiles_per_task;
        const int q0 = (gtile - task * p.tiles_per_task) * 128;
        const float* srcs[2] = {p.x[pair] + (long long)task * p.n * p.H * p.W * 32, p.g[pair] + (long long)task * p.n * p.H * p.W * 32};
        unsigned char* raw = smem + p.off_raw + (size_t)st * p.raw_stage;
        const float* my_src[2] = {nullptr, nullptr};
        uint32_t my_dst[2] = {0u, 0u}, my_bytes[2] = {0u, 0u};
#pragma unroll
        for (int seg = 0; seg < 2; ++seg) {
          const int base_q = seg == 0 ? q0 - 1 : q0 - p.Wp, R = seg == 0 ? p.Rx : p.Rg;
          unsigned char* dst0 = raw + (seg == 0 ? 0 : (size_t)p.Rx * 128);
          for (int j = lane; j < R; j += 32)                       // zero rows at padding / out-of-range positions
            if (pos_to_pixel(p.pm, base_q + j) < 0) {
              float4* z = reinterpret_cast<float4*>(dst0 + (size_t)j * 128);
#pragma unroll
              for (int c = 0; c < 8; ++c) z[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          const int qa = max(base_q, 0), qb = min(base_q + R, Qtask);
          if (qa < qb) {
            const int ra = (int)fastdiv((uint32_t)qa, p.pm.drow), rb = (int)fastdiv((uint32_t)(qb - 1), p.pm.drow);
            const int rr = ra + lane;                                // global row index img * Hp + r
            if (rr <= rb) {
              const int rowq = rr * p.Wp;
              const int img = (int)fastdiv((uint32_t)rowq, p.pm.dimg), r = rr - img * p.Hp;
              const int c_lo = max(qa, rowq) - rowq, c_hi = min(min(qb, rowq + p.W) - rowq, p.W);
              if (r >= 1 && c_hi > c_lo) {
                my_src[seg] = srcs[seg] + ((long long)(img * p.H + r - 1) * p.W + c_lo) * 32;
                my_dst[seg] = smem_u32(dst0 + (size_t)(rowq + c_lo - base_q) * 128);
                my_bytes[seg] = (uint32_t)(c_hi - c_lo) * 128u;
              }
            }
          }
        }
        uint32_t total = my_bytes[0] + my_bytes[1];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
        fence_proxy_async();                                       // generic zero stores / earlier reads before async-proxy writes
        __syncwarp();
        if (lane == 0) {
          if (total) mbar_arrive_expect_tx(bar_rfull + 8 * st, total);
          else mbar_arrive(bar_rfull + 8 * st);
        }
        __syncwarp();
#pragma unroll
        for (int seg = 0; seg < 2; ++seg)
          if (my_bytes[seg]) bulk_g2s(my_dst[seg], my_src[seg], my_bytes[seg], bar_rfull + 8 * st);
      }
    }
  } else if (warp < MMA_WARP) {
    // ========================================= producers =============================================
    // A thread moves 8 channels (one 32 B swizzle chunk) of a row with one 256-bit load; rows of a unit:
    //   x tile : j = jrow + PR*k (k < 3, j < Rx = 131)  <-> position q0 - 1 + j
    //   g halo : i = jrow + PR*k (k < 4, i < Rg = 128 + 2*Wp)  <-> position q0 - Wp + i   (zero at padding positions)
    // Register-level software pipeline: the loads of unit u+1 are issued before unit u is converted and
    // stored (two register sets used alternately).
    constexpr int PR = NPROD / 4;
    constexpr int KX = (131 + PR - 1) / PR, KG = (224 + PR - 1) / PR;   // row passes per thread: x tile (131 rows), g halo (<= 224)
    const int ptid = tid - WT_DRAINERS;
    const int c8 = ptid & 3, jrow = ptid >> 2;
    struct Regs { float4 x[2 * KX]; float4 g[2 * KG]; };
#ifdef XM_TC_TIMING
    long long g_wsf = 0, g_max = 0, g_mw = 0, g_conv = 0, g_fence = 0;
#endif
    auto issue = [&](int u, Regs& r) {
      const int it = u / p.npairs, pair = u - it * p.npairs;
      const int gtile = g_lo + it, task = gtile / p.tiles_per_task;
      const int q0 = (gtile - task * p.tiles_per_task) * 128;
      const float* X = p.x[pair] + (long long)task * p.n * p.H * p.W * p.x_cs + p.x_co + c8 * 8;
      const float* G = p.g[pair] + (long long)task * p.n * p.H * p.W * p.g_cs + p.g_co + c8 * 8;
#pragma unroll
      for (int k = 0; k < KX; ++k) {
        const int j = jrow + PR * k;
        const int px = j < p.Rx ? pos_to_pixel(p.pm, q0 - 1 + j) : -1;
        r.x[2 * k] = r.x[2 * k + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (px >= 0) ldg256(X + (long long)px * p.x_cs, r.x[2 * k], r.x[2 * k + 1]);
      }
#pragma unroll
      for (int k = 0; k < KG; ++k) {
        const int i = jrow + PR * k;
        const int px = i < p.Rg ? pos_to_pixel(p.pm, q0 - p.Wp + i) : -1;
        r.g[2 * k] = r.g[2 * k + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (px >= 0) ldg256(G + (long long)px * p.g_cs, r.g[2 * k], r.g[2 * k + 1]);
      }
    };
    auto store = [&](int u, const Regs& r) {
      const int s = u & 1;
#ifdef XM_TC_TIMING
      long long tp = clock64();
#endif
      if (F16) {
        float mx = 0.f, mg = 0.f;
#pragma unroll
        for (int k = 0; k < 2 * KX; ++k) mx = fmaxf(mx, wt_absmax(r.x[k]));
#pragma unroll
        for (int k = 0; k < 2 * KG; ++k) mg = fmaxf(mg, wt_absmax(r.g[k]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
          mg = fmaxf(mg, __shfl_xor_sync(0xffffffffu, mg, o));
        }
        if (lane == 0) {
          atomicMax(&smaxx[u & 3], __float_as_uint(mx));
          atomicMax(&smaxg[u & 3], __float_as_uint(mg));
          mbar_arrive(bar_max + 8 * s);
        }
      }
#ifdef XM_TC_TIMING
      const long long ts0 = clock64();
      g_max += ts0 - tp;
#endif
      if (u >= 2) mbar_wait(bar_sfree + 8 * s, ((u - 2) >> 1) & 1);     // MMAs of unit u-2 have read stage s
#ifdef XM_TC_TIMING
      g_wsf += clock64() - ts0; tp = clock64();
#endif
      unsigned char* xhi = smem + (s ? p.off_x1 : p.off_x0);
      unsigned char* ghi = smem + (s ? p.off_g1 : p.off_g0);
      if (F16) {
        mbar_wait(bar_max + 8 * s, (u >> 1) & 1);
#ifdef XM_TC_TIMING
        g_mw += clock64() - tp; tp = clock64();
#endif
        const int kx = wt_scale_exp(__uint_as_float(smaxx[u & 3])), kg = wt_scale_exp(__uint_as_float(smaxg[u & 3]));
        // slots of unit u + 2 were last read for unit u - 2: every producer has since passed two of these waits
        if (ptid == 0) { uexp[u & 3] = kx + kg; smaxx[(u + 2) & 3] = 0u; smaxg[(u + 2) & 3] = 0u; }
        const float sx = wt_exp2i(kx), sg = wt_exp2i(kg);
#pragma unroll
        for (int k = 0; k < KX; ++k) {
          const int j = jrow + PR * k;
          if (j < p.Rx) {
            const size_t o = (size_t)j * 64 + (size_t)((c8 ^ ((j >> 1) & 3)) * 16);   // 16 B chunk swizzled with the row pair
            uint4 h, l;
            wt_split_f16(r.x[2 * k], r.x[2 * k + 1], sx, h, l);
            *reinterpret_cast<uint4*>(xhi + o) = h;
            *reinterpret_cast<uint4*>(xhi + xset + o) = l;
          }
        }
#pragma unroll
        for (int k = 0; k < KG; ++k) {
          const int i = jrow + PR * k;
          if (i < p.Rg) {
            const size_t o = (size_t)i * 64 + (size_t)((c8 ^ ((i >> 1) & 3)) * 16);
            uint4 h, l;
            wt_split_f16(r.g[2 * k], r.g[2 * k + 1], sg, h, l);
            *reinterpret_cast<uint4*>(ghi + o) = h;
            *reinterpret_cast<uint4*>(ghi + gset + o) = l;
          }
        }
#ifdef XM_TC_TIMING
        g_conv += clock64() - tp; tp = clock64();
#endif
        fence_proxy_async();
        mbar_arrive(bar_full + 8 * s);
#ifdef XM_TC_TIMING
        g_fence += clock64() - tp;
#endif
        return;
      }
#pragma unroll
      for (int k = 0; k < KX; ++k) {
        const int j = jrow + PR * k;
        if (j < p.Rx) {
          const size_t o = (size_t)j * 128 + (size_t)((c8 ^ (j & 3)) * 32);   // 32 B chunk swizzled with the row
          float4 h, l;
          split_tf32_fast(r.x[2 * k], h, l);
          *reinterpret_cast<float4*>(xhi + o) = h;
          *reinterpret_cast<float4*>(xhi + xset + o) = l;
          split_tf32_fast(r.x[2 * k + 1], h, l);
          *reinterpret_cast<float4*>(xhi + o + 16) = h;
          *reinterpret_cast<float4*>(xhi + xset + o + 16) = l;
        }
      }
#pragma unroll
      for (int k = 0; k < KG; ++k) {
        const int i = jrow + PR * k;
        if (i < p.Rg) {
          const size_t o = (size_t)i * 128 + (size_t)((c8 ^ (i & 3)) * 32);
          float4 h, l;
          split_tf32_fast(r.g[2 * k], h, l);
          *reinterpret_cast<float4*>(ghi + o) = h;
          *reinterpret_cast<float4*>(ghi + gset + o) = l;
          split_tf32_fast(r.g[2 * k + 1], h, l);
          *reinterpret_cast<float4*>(ghi + o + 16) = h;
          *reinterpret_cast<float4*>(ghi + gset + o + 16) = l;
        }
      }
      fence_proxy_async();
      mbar_arrive(bar_full + 8 * s);
    };
    if (F16 && NR > 0) {
      Regs r;
#ifdef XM_TC_TIMING
      long long t_rf = 0, t_all = 0, t0, t1;
#endif
      for (int u = 0; u < nunits; ++u) {
        const int st = u % NR;
#ifdef XM_TC_TIMING
        t0 = clock64();
#endif
        mbar_wait(bar_rfull + 8 * st, (u / NR) & 1);               // the unit's rows have landed
#ifdef XM_TC_TIMING
        t1 = clock64(); t_rf += t1 - t0;
#endif
        const unsigned char* raw = smem + p.off_raw + (size_t)st * p.raw_stage;
        const unsigned char* rawg = raw + (size_t)p.Rx * 128;
#pragma unroll
        for (int k = 0; k < KX; ++k) {
          const int j = jrow + PR * k;
          r.x[2 * k] = r.x[2 * k + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (j < p.Rx) {
            r.x[2 * k] = *reinterpret_cast<const float4*>(raw + (size_t)j * 128 + c8 * 32);
            r.x[2 * k + 1] = *reinterpret_cast<const float4*>(raw + (size_t)j * 128 + c8 * 32 + 16);
          }
        }
#pragma unroll
        for (int k = 0; k < KG; ++k) {
          const int i = jrow + PR * k;
          r.g[2 * k] = r.g[2 * k + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (i < p.Rg) {
            r.g[2 * k] = *reinterpret_cast<const float4*>(rawg + (size_t)i * 128 + c8 * 32);
            r.g[2 * k + 1] = *reinterpret_cast<const float4*>(rawg + (size_t)i * 128 + c8 * 32 + 16);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_rfree + 8 * st);            // this warp's rows are in registers
        store(u, r);
#ifdef XM_TC_TIMING
        t_all += clock64() - t1;
#endif
      }
#ifdef XM_TC_TIMING
      if (blockIdx.x == 0 && (tid == WT_DRAINERS || tid == WT_DRAINERS + 200))
        printf("producer tid %d: per unit wait-raw-full %lld | rest %lld = max+post %lld wait-sfree %lld wait-max %lld convert+store %lld fence+arrive %lld (+ raw reads)\n",
               tid, t_rf / nunits, t_all / nunits, g_max / nunits, g_wsf / nunits, g_mw / nunits, g_conv / nunits, g_fence / nunits);
#endif
    } else {
    Regs ra, rb;
    if (nunits > 0) issue(0, ra);
    for (int u = 0; u < nunits; u += 2) {
      if (u + 1 < nunits) issue(u + 1, rb);
      store(u, ra);
      if (u + 1 < nunits) {
        if (u + 2 < nunits) issue(u + 2, ra);
        store(u + 1, rb);
      }
    }
    }
  } else {
    // ======================================= MMA issuer =================================================
#ifdef XM_TC_TIMING
    long long m_wf = 0, m_wt = 0, m_is = 0, mt0;
    unsigned long long ns0, ns1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns0));
    const long long c_begin = clock64();
#endif
    for (int u = 0; u < nunits; ++u) {
      const int s = u & 1, it = u / p.npairs, pair = u - it * p.npairs;
      const int set = F16 ? (u & 1) : (it & 1);
#ifdef XM_TC_TIMING
      mt0 = clock64();
#endif
      mbar_wait(bar_full + 8 * s, (u >> 1) & 1);
#ifdef XM_TC_TIMING
      m_wf += clock64() - mt0; mt0 = clock64();
#endif
      if (F16) {
        if (u >= 2) mbar_wait(bar_tfree + 8 * set, ((u - 2) >> 1) & 1);
#ifdef XM_TC_TIMING
        m_wt += clock64() - mt0; mt0 = clock64();
#endif
      } else if (pair == 0 && it >= 2) {
        mbar_wait(bar_tfree + 8 * set, ((it - 2) >> 1) & 1);
      }
      tc_fence_after();
      if (F16) {
        if (elect_one_sync()) {
          // rows of 64 B, SWIZZLE_64B (layout type 4), SBO = 8 rows; A: LBO = one row, B: LBO = Wp rows; K = 16 rows
          const uint32_t x_hi0 = umma_desc_lo(smem_u32(smem + (s ? p.off_x1 : p.off_x0)), 64u);
          const uint32_t x_lo0 = x_hi0 + (uint32_t)(xset >> 4);
          const uint32_t g_hi0 = umma_desc_lo(smem_u32(smem + (s ? p.off_g1 : p.off_g0)), (uint32_t)p.Wp * 64u);
          const uint32_t g_lo0 = g_hi0 + (uint32_t)(gset >> 4);
          constexpr uint32_t dhi = umma_desc_hi(512u, 4u);
          const uint32_t d = tmem_base + (uint32_t)(set * 96);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint32_t ko = (uint32_t)ks * 64u;                   // 16 position rows of 64 B per K step
            umma_f16_lh(d, x_lo0 + ko, dhi, g_hi0 + ko, dhi, WT_IDESC_F16, ks == 0 ? 0u : 1u);
            umma_f16_lh(d, x_hi0 + ko, dhi, g_lo0 + ko, dhi, WT_IDESC_F16, 1u);
            umma_f16_lh(d, x_hi0 + ko, dhi, g_hi0 + ko, dhi, WT_IDESC_F16, 1u);
          }
          umma_commit(bar_sfree + 8 * s);
          umma_commit(bar_tfull + 8 * set);
        }
        __syncwarp();
#ifdef XM_TC_TIMING
        m_is += clock64() - mt0;
        if (u == nunits - 1 && blockIdx.x == 0 && lane == 0) {
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1));
          printf("mma: per unit wait-full %lld wait-tfree %lld issue %lld; %lld cycles in %llu ns = %.3f GHz, %d units\n", m_wf / nunits,
                 m_wt / nunits, m_is / nunits, clock64() - c_begin, ns1 - ns0, (double)(clock64() - c_begin) / (double)(ns1 - ns0), nunits);
        }
#endif
        continue;
      }
      if (elect_one_sync()) {
        // descriptor low words (start address | LBO); per MMA only the start field moves.  A = x tile with
        // LBO = one position row (lane group kw = tile shifted by kw rows), B = g halo with LBO = Wp position rows
        // (column group j = halo shifted by j*Wp rows = kernel row 2 - j).
        const uint32_t x_hi0 = umma_desc_lo(smem_u32(smem + (s ? p.off_x1 : p.off_x0)), 128u);
        const uint32_t x_lo0 = x_hi0 + (uint32_t)(xset >> 4);
        const uint32_t g_hi0 = umma_desc_lo(smem_u32(smem + (s ? p.off_g1 : p.off_g0)), (uint32_t)p.Wp * 128u);
        const uint32_t g_lo0 = g_hi0 + (uint32_t)(gset >> 4);
        constexpr uint32_t dhi = umma_desc_hi(512u, 1u);            // SBO 512 B, 128B-swizzle / 32B-base layout
        const uint32_t d = tmem_base + (uint32_t)(set * 96);
#pragma unroll 4
        for (int ks = 0; ks < 16; ++ks) {
          const uint32_t fresh = (pair == 0 && ks == 0) ? 0u : 1u;
          const uint32_t ko = (uint32_t)ks * 64u;                   // 8 position rows of 128 B per K step
          umma_tf32_lh(d, x_lo0 + ko, dhi, g_hi0 + ko, dhi, WT_IDESC, fresh);
          umma_tf32_lh(d, x_hi0 + ko, dhi, g_lo0 + ko, dhi, WT_IDESC, 1u);
          umma_tf32_lh(d, x_hi0 + ko, dhi, g_hi0 + ko, dhi, WT_IDESC, 1u);
        }
        umma_commit(bar_sfree + 8 * s);
        if (pair == p.npairs - 1) umma_commit(bar_tfull + 8 * set);
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(WT_TMEM_COLS) : "memory");
  }
}

static bool wgrad_tc_layout(const XmBlockGeom& g, WgradTcK& p, size_t& smem, bool f16 = (g_precise == 2)) {
  if (g.cin != g.cout || (g.cout != 32 && g.cout != 64) || (g.stride != 1 && g.stride != 2)) return false;
  p.n = g.n; p.H = g.hin; p.W = g.win; p.Hp = g.hin + 1; p.Wp = g.win + 1;
  p.Q = g.n * p.Hp * p.Wp;
  p.pm = make_posmap(g.n, g.hin, g.win);
  p.tiles_per_task = (p.Q + 127) / 128;
  p.Rx = 128 + 3;                                    // x tile + the three kw shifts
  p.Rg = 128 + 2 * p.Wp;                             // g halo: kernel rows shift it by 0, Wp, 2*Wp rows
  const int rowb = f16 ? 64 : 128, ralign = f16 ? 15 : 7;
  p.xbuf = ((p.Rx + 1 + ralign) & ~ralign) * rowb;   // + the row the unused lane group reads; whole 1 KB units
  p.gbuf = ((p.Rg + ralign) & ~ralign) * rowb;       // keep every buffer 1024 B aligned
  p.off_x0 = 0;
  p.off_g0 = p.off_x0 + 2 * p.xbuf;
  p.off_x1 = p.off_g0 + 2 * p.gbuf;
  p.off_g1 = p.off_x1 + 2 * p.xbuf;
  p.off_bar = p.off_g1 + 2 * p.gbuf;
  const int bar_bytes = 8 * 8 + 16 + 12 * 4 + 16 + 2 * 8 + 6 * 8;
  smem = (size_t)p.off_bar + bar_bytes;
  p.off_raw = 0; p.raw_stage = 0; p.raw_stages = 0;
  // Bulk-copy (TMA) loader: opt-in (XM_WT_TMA=1).  Correct, but measured SLOWER than the register-prefetch producers
  // (42x42, 32 tasks: 231 us against 193 us): the kernel is bound by how fast the memory system feeds 148 persistent
  // CTAs (~1.9 TB/s), not by request latency or producer instructions, and the raw ring adds shared-memory traffic.
  static const bool want_tma = getenv("XM_WT_TMA") && atoi(getenv("XM_WT_TMA")) != 0;
  if (f16 && want_tma && g.cout == 32) {      // contiguous 128 B position rows: as many raw stages (2..3) as fit
    const int off_raw = (p.off_bar + bar_bytes + 127) & ~127, stage = (p.Rx + p.Rg) * 128;
    int stages = (227 * 1024 - off_raw) / stage;
    if (stages > 3) stages = 3;
    if (stages >= 2) {
      p.off_raw = off_raw; p.raw_stage = stage; p.raw_stages = stages;
      smem = (size_t)off_raw + (size_t)stages * stage;
    }
  }
  // partial slots per task = the most CTAs of the persistent grid that can touch one task
  const long long total = (long long)g.tasks * p.tiles_per_task;
  p.ctas = (int)(total < num_sms() ? total : num_sms());
  p.splits = (p.ctas + g.tasks - 1) / g.tasks + 1;
  return smem <= 227 * 1024 && p.Rg <= WT_PRODUCERS;     // a producer thread stages <= 4 g rows per unit
}

// partial-buffer floats needed by the tcgen05 path for this geometry (0 if the shape is not covered)
long long wgrad_tc_partial_floats(const XmBlockGeom& g) {
  WgradTcK p{};
  size_t smem;
  if (!wgrad_tc_layout(g, p, smem)) return 0;
  const int blocks = g.cout / 32;                      // 64-channel layers: one launch per (cout block, cin block)
  long long floats = (long long)blocks * blocks * g.tasks * p.splits * 9 * 32 * 32;
  if (g.stride == 2) floats += 2LL * g.tasks * g.n * g.hin * g.win * g.cout;   // zero-inserted cotangents (two pairs)
  return floats;
}

// (conv_tc.cu) full[img][y][x] = (x, y even) ? src[img][y/2][x/2] : 0
int launch_upsample2(const float* src, float* full, long long imgs, int H, int W, int hz, int wz, int C, cudaStream_t stream);

// Returns the number of partial slots per task in a->partial (>0) when handled (slot j of task t is written by CTA
// first(t) + j of the persistent grid; wgrad_reduce_kernel derives each task's slot count from *ctas_out and
// *tiles_out), 0 when not covered, <0 on error.
int wgrad_tc_try(const XmWgradArgs* a, cudaStream_t stream, int* rc_out, int* ctas_out, int* tiles_out) {
  const XmBlockGeom& g = a->g;
  *rc_out = 0;
  if (a->src_nchw) return 0;
  WgradTcK p{};
  size_t smem;
  if (!wgrad_tc_layout(g, p, smem)) return 0;
  p.tasks = g.tasks;
  p.npairs = a->x2 ? 2 : 1;
  p.x[0] = a->x1; p.g[0] = a->g1; p.x[1] = a->x2; p.g[1] = a->g2;
  p.partial = a->partial;
  const bool tma = p.raw_stages > 0;
  auto kern = g_precise == 2 ? (tma ? wgrad_tc_kernel<true, WT_NPW_TMA> : wgrad_tc_kernel<true, 7>) : wgrad_tc_kernel<false, 7>;
  const int nthreads = WT_DRAINERS + (tma ? WT_NPW_TMA : 7) * 32 + 32;
  {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { *rc_out = fail((int)e, "cudaFuncSetAttribute(wgrad_tc): %s", cudaGetErrorString(e)); return -1; }
  }
  // block (cb, ib) of a wide layer: gW[32 cb .. +31][32 ib .. +31] from g's channel block cb and x's channel block ib,
  // into partial region cb * blocks + ib
  const int blocks = g.cout / 32;
  p.x_cs = p.g_cs = g.cout;
  if (g.stride == 2) {
    // stride-2 weight gradient = stride-1 weight gradient against the cotangent with zeros inserted between its
    // elements (full input resolution); the copies live behind the partial blocks in the scratch buffer
    float* up = a->partial + (long long)blocks * blocks * g.tasks * p.splits * 9 * 32 * 32;
    const long long imgs = (long long)g.tasks * g.n, per = imgs * g.hin * g.win * g.cout;
    for (int pair = 0; pair < p.npairs; ++pair) {
      if (int rc = launch_upsample2(p.g[pair], up + pair * per, imgs, g.hin, g.win, g.hz, g.wz, g.cout, stream)) { *rc_out = rc; return -1; }
      p.g[pair] = up + pair * per;
    }
  }
  for (int cb = 0; cb < blocks; ++cb)
    for (int ib = 0; ib < blocks; ++ib) {
      p.g_co = 32 * cb; p.x_co = 32 * ib;
      p.partial = a->partial + (long long)(cb * blocks + ib) * g.tasks * p.splits * 9 * 32 * 32;
      kern<<<p.ctas, nthreads, smem, stream>>>(p);
      if (int rc = launched("xm_wgrad(tcgen05)")) { *rc_out = rc; return -1; }
    }
  *ctas_out = p.ctas;
  *tiles_out = p.tiles_per_task;
  return p.splits;
}

}  // namespace xm
