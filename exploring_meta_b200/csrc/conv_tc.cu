// conv_tc.cu -- tcgen05 / TMEM implementation of xm_conv for the stride-1 layers with 32 output channels:
// the 32 -> 32 layers (forward, data-gradient and their tangent versions) and the image layer (cin <= 4, NCHW
// source).  Together these are every convolution of the Mini-ImageNet network.
//
// Formulation ("flattened padded pixels"): the n images of a task are laid out as ONE 1-D sequence of
// positions q = (img, r, c), r in [0, H], c in [0, W], where row r = 0 and column c = W are zero padding
// shared between neighbouring rows / images (Hp = H+1, Wp = W+1).  For output position q, tap (kh, kw) of the
// 3x3 pad-1 stencil reads position q + (kh-1)*Wp + (kw-1): every tap is the SAME 1-D array shifted by a
// constant.  A GEMM tile is therefore 128 consecutive positions (M = 128 rows of one tcgen05.mma), whatever
// the map size -- 84x84, 42x42, 21x21, 10x10 and 5x5 maps all fill tiles equally well -- and the A operand of
// tap (kh, kw) is the staged halo [q0 - Wp - 1, q0 + 128 + Wp + 1) read at row offset kh*Wp + kw.
//
// Shared-memory operand layout (UMMA canonical K-major, no swizzle): channel-group planes
//     A[c4 = ch/4][row][4 ch]   (16 B per row and plane; rows 16 B apart => SBO = 128 B, LBO = plane stride)
// so a shift by s positions is a shift of the descriptor start address by 16*s bytes -- no im2col copy, no
// per-tap restaging: each input element is written to shared memory once and read by 9 taps x 3 passes.
// Weights sit in the same layout B[tap][c4][cout][4] and stay resident for the CTA's lifetime.
// Image layer: one plane (3 channels + a zero), and one K = 8 MMA covers TWO taps -- the second K group is
// simply the same plane at LBO = (offset of tap t+1 - offset of tap t) * 16 B.
//
// Precision: parity is stated in fp32, so every product is evaluated as a 3-term TF32 expansion
// (a_hi*b_hi + a_hi*b_lo + a_lo*b_hi; hi = rna_tf32(x), lo = x - hi): the producer warps split the activations
// while staging.  The three kw taps are stacked in N (columns kw*32 + cout): per kernel row kh and channel octet the
// MMA thread issues three tcgen05.mma (kind::tf32, M = 128, N = 96, K = 8) -- 36 per tile -- into ONE 96-column TMEM
// accumulator.  The tensor core truncates when it adds, so the 24 small correction products are issued first (they
// sum among themselves) and the 12 hi*hi products go on top: the corrections lose at most one rounding.
// (Stacking the hi and lo weights in N = 192 -- 24 MMAs per tile into two accumulators -- is equivalent in time:
// an MMA's floor is proportional to N)
//
// FP16-split variant (template parameter F16, xm_set_precision(2)): the same 3-term expansion on kind::f16 -- twice the
// tensor rate of kind::tf32 and half the shared-memory bytes per element, the two things that bound this kernel.
// x * s = hi + lo * 2^-11 with hi = fp16(x * s), lo = fp16((x * s - hi) * 2^11): 22 significant bits, like the TF32 pair.
// fp16 has 5 exponent bits, so every staged tile is scaled by a power of two s = 2^k chosen from the tile's own
// absolute maximum (producer warps: register max -> shared atomicMax -> named barrier), the task's weight block by
// its own; the drain multiplies by 2^-(ka + kb) (exact).  Elements 2^-28 below the tile maximum lose relative
// precision -- invisible in sums that the large elements dominate.  Operand planes hold 8 channels (16 B of fp16), so
// a shift by one position is still a 16-byte shift of the descriptor start address.  The taps are NOT stacked in N here:
// nine taps x two K = 16 steps x three terms = 54 MMAs (M = 128, N = 32) accumulate into the same 32 columns, so
// accumulator row i is output position q0 + i and the drain -- which bounds the TF32 kernel (XM_TC_TIMING: ~7000 cycles of
// TMEM loads, kw shift-add shuffles, boundary exchange, stores and statistics per tile and drain group against ~2300
// cycles of MMAs) -- loses its cross-lane shift-add, the group barrier and the boundary fix-up.
//
// Pipeline (warp-specialised, 1 CTA per SM, 512 threads): warps 0-6 stage tiles (global -> TF32 split -> smem), one
// elected lane of warp 7 issues the MMAs, warps 8-11 and 12-15 are two drain groups working on alternate tiles
// (warp w reads TMEM lane quarter w%4: tcgen05.ld -> kw shift-add across lanes -> the warp's 32 finished rows staged in
// 4 KB of its own shared memory -> read back transposed: 8 lanes per row for coalesced NHWC stores, lane = channel for
// the BatchNorm column sums).  Two shared-memory stages and four TMEM accumulator sets; mbarrier full/free handshakes;
// tcgen05.commit signals completion.  Measured (XM_TC_TIMING, scripts/gpu_tc_timing.sh, scripts/gpu_conv_elim.sh;
// DESIGN.md section 8): every role alone needs 115-140 us of the 42x42 forward's 166 us.
// 64-channel and stride-2 layers reuse this kernel through channel-block / full-resolution passes (conv_tc_try).
#include <cuda_fp16.h>
#include "tc.cuh"

#ifndef XM_TC_GROUPS
#define XM_TC_GROUPS 2
#endif

namespace xm {

extern int g_precise;                 // conv.cu: 0 = 1xTF32, 1 = 3xTF32, 2 = 3xFP16-split on the tcgen05 conv kernel

constexpr int TC_PRODUCERS = 224;     // warps 0-6; warp 7 issues the MMAs
constexpr int TC_DRAINERS = 128;      // per drain group (warps 8-11, 12-15)
constexpr int TC_GROUPS = XM_TC_GROUPS;          // drain groups: group g drains tiles g, g + G, ...: a tile's drain is
                                      // latency-bound (TMEM loads, shuffles, exchange) and ~1.5x the MMA time of a tile,
                                      // so two tiles are drained concurrently (16 warps: 128 registers each)
constexpr int TC_THREADS = TC_PRODUCERS + 32 + TC_GROUPS * TC_DRAINERS;
constexpr int TC_STG_BYTES = 32 * 128 + 32 * 8;   // per drain WARP: its 32 accumulator rows (128 B each, chunks XOR-swizzled) + 32 output offsets
constexpr int TC_TILE = 126;          // outputs per tile: 128 accumulator rows minus the two shifted-out rows
constexpr int TC_SETS = 4;            // TMEM accumulator sets: tile it -> set it % 4 (drained by group it % G; G divides 4)
constexpr int TC_TMEM_COLS = 512;     // 4 sets x 96 columns = 384 -> next power of two
static_assert(TC_SETS % TC_GROUPS == 0, "a TMEM set must always be drained by the same group");
constexpr uint32_t TC_IDESC = umma_idesc_tf32(128, 96, 0, 0);   // A and B K-major, N = 3 taps x 32 channels
// kind::f16: fp16 inputs (a_format = b_format = 0), fp32 accumulate
constexpr uint32_t TC_IDESC_F16 = (1u << 4) | ((32u >> 3) << 17) | ((128u >> 4) << 24);   // M = 128, N = 32 (one tap)
constexpr float LO_SCALE = 2048.f, LO_UNSCALE = 1.f / 2048.f;

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
// power-of-two scale 2^k that brings a maximum magnitude m into [2^14, 2^15) (fp16 overflows at 65520); k = 0 for m = 0
__device__ __forceinline__ int f16_scale_exp(float m) {
  if (!(m > 0.f)) return 0;
  int k = 14 - (int)((__float_as_uint(m) >> 23) & 0xffu) + 127;
  return max(-100, min(100, k));
}
__device__ __forceinline__ float exp2i(int k) { return __uint_as_float((uint32_t)(127 + k) << 23); }
// 8 floats -> 8 fp16 hi + 8 fp16 scaled residuals, packed as two 16-byte vectors
__device__ __forceinline__ void split_f16(const float4& a, const float4& b, float sc, uint4& hi, uint4& lo) {
  const float x[8] = {a.x * sc, a.y * sc, a.z * sc, a.w * sc, b.x * sc, b.y * sc, b.z * sc, b.w * sc};
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 hh = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
    const float2 back = __half22float2(hh);
    const __half2 ll = __floats2half2_rn((x[2 * i] - back.x) * LO_SCALE, (x[2 * i + 1] - back.y) * LO_SCALE);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    l[i] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// 16-byte store, or (accumulate) a vector reduction into global memory: out += v, rounded like an fp32 add.
__device__ __forceinline__ void put4(float* dst, const float4 v, int accumulate) {
  if (accumulate)
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  else
    *reinterpret_cast<float4*>(dst) = v;
}

struct ConvTcK {
  int tasks, n, H, W, Hp, Wp;        // source == output spatial dims (stride 1)
  PosMap pm;
  int Q;                             // positions per task = n*Hp*Wp
  int tiles_per_task;
  int R, plane_bytes;                // staged rows per tile, bytes per channel-group plane
  int wmode;                         // 0 forward, 1 data-gradient (transposed weights, flipped taps)
  int stat_mode, accumulate;
  int ctas_per_task;                 // > 0: CTAs never cross a task boundary (small maps)
  const float* src; const float* w; long long wstride;
  float* out; const float* aux; double* stats;
  // channel blocking (64-channel layers run as 2 x 2 passes over 32-channel blocks): floats per position of the
  // source / output tensors, first channel of this pass' block in each, and the weight block W[w_ao + a][w_bo + b]
  // inside the task's [*][w_cin][3][3] tensor
  int src_cs, src_co, out_cs, out_co, w_cin, w_ao, w_bo;
};

template <bool F16>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const ConvTcK p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef XM_TC_TIMING
  // phase timestamps of CTA 0 / thread 0 (ns, %globaltimer), printed once at the end (a printf costs tens of microseconds)
  unsigned long long ph_t[8]; int ph_n = 0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ph_t[0]));
#define XM_PH(name) do { if (ph_n < 7) { ++ph_n; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ph_t[ph_n])); } } while (0)
#else
#define XM_PH(name) do { } while (0)
#endif
  constexpr int NPLANES = F16 ? 4 : 8;            // channel-group planes of the A operand (16 B per row and plane)
  constexpr int BFLOATS = F16 ? 3 * 4 * 96 * 4    // B[kh][c8][n = kw*32 + cout][8 halfs] (4-byte words)
                              : 3 * 8 * 96 * 4;   // B[kh][c4][n = kw*32 + cout][4]
  const int plane = p.plane_bytes, set_bytes = NPLANES * plane;   // one hi (or lo) set

  float* Bhi = reinterpret_cast<float*>(smem);
  float* Blo = Bhi + BFLOATS;
  unsigned char* Abase = smem + 2 * BFLOATS * 4;               // stage s: hi at s*2*set, lo at (s*2+1)*set
  uint64_t* bars = reinterpret_cast<uint64_t*>(Abase + 4 * set_bytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  float* xch = reinterpret_cast<float*>(bars + 10);            // [group][tile parity][half][4 warps][3][16] boundary rows
  // FP16 variant: per-tile absolute maxima (float bits) and scale exponents, weight-block maximum / exponent
  uint32_t* smax = reinterpret_cast<uint32_t*>(xch + TC_GROUPS * 2 * 2 * 4 * 3 * 16);   // [4]
  int* a_exp = reinterpret_cast<int*>(smax + 4);                         // [4]
  uint32_t* wmax = reinterpret_cast<uint32_t*>(a_exp + 4);               // [1]
  int* b_exp = reinterpret_cast<int*>(wmax + 1);                         // [1]
  uint64_t* mxbars = reinterpret_cast<uint64_t*>(smax + 12);             // [2] "every producer warp has posted its tile maximum"
  const uint32_t bar_max = smem_u32(mxbars);
  // "accumulators of a tile complete": ONE BARRIER PER TMEM SET (tile it -> set it % 4, phase it / 4).  A set is always
  // drained by the same group (G divides 4), which therefore sees every phase of the set's barrier in order, and the
  // barrier cannot run two phases ahead of it (the MMAs of tile it + 4 need the set that tile it's drain releases) -- so
  // the parity wait is unambiguous.  Four sets let the MMAs run up to three tiles ahead of a drain group.
  const uint32_t bar_tfull = smem_u32(mxbars + 2);
  unsigned char* stg_base = reinterpret_cast<unsigned char*>(mxbars + 2 + TC_SETS);   // 16-byte aligned
  const uint32_t bar_full = smem_u32(bars), bar_sfree = smem_u32(bars + 2), bar_tfree = smem_u32(bars + 4);   // [4]

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_full + 8 * s, TC_PRODUCERS);
      mbar_init(bar_sfree + 8 * s, 1);
      if (F16) mbar_init(bar_max + 8 * s, TC_PRODUCERS / 32);
    }
    for (int g = 0; g < TC_SETS; ++g) { mbar_init(bar_tfull + 8 * g, 1); mbar_init(bar_tfree + 8 * g, TC_DRAINERS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 7) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // ---- CTA -> tiles.  Large maps: persistent grid over the flattened (task, tile) list, CTA c owns tiles
  // [c*G/n, (c+1)*G/n), i.e. one or two SEGMENTS of consecutive tiles of one task each (the per-task weights are
  // re-staged at the task boundary) -- every SM gets the same share whatever the task count (32 tasks on 148 SMs: 4 CTAs
  // per task leave 20 SMs idle).  Small maps (few tiles per CTA): whole CTAs per task (p.ctas_per_task > 0).
  int g_lo, g_hi;
  if (p.ctas_per_task > 0) {
    const int t = blockIdx.x / p.ctas_per_task, c = blockIdx.x - t * p.ctas_per_task;
    g_lo = t * p.tiles_per_task + (int)((long long)p.tiles_per_task * c / p.ctas_per_task);
    g_hi = t * p.tiles_per_task + (int)((long long)p.tiles_per_task * (c + 1) / p.ctas_per_task);
  } else {
    const long long G = (long long)p.tasks * p.tiles_per_task;
    g_lo = (int)(G * blockIdx.x / gridDim.x);
    g_hi = (int)(G * (blockIdx.x + 1) / gridDim.x);
  }
  uint32_t tmem_base = 0;
  XM_PH("barriers + tmem alloc issued");
  for (int gt = g_lo, seg = 0; gt < g_hi; ++seg) {
  const int task = gt / p.tiles_per_task, tile0 = gt - task * p.tiles_per_task;
  const int ntiles = min(g_hi - gt, p.tiles_per_task - tile0);   // my tiles of this task: tile0 .. tile0 + ntiles - 1
  gt += ntiles;
  if (seg > 0) {
    // every role is done with the previous segment (the drainers waited for its last accumulators, which also
    // completes every MMA that read the weights / stages): barriers restart from phase 0
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      for (int s = 0; s < 2; ++s) {
        mbar_init(bar_full + 8 * s, TC_PRODUCERS);
        mbar_init(bar_sfree + 8 * s, 1);
        if (F16) mbar_init(bar_max + 8 * s, TC_PRODUCERS / 32);
      }
      for (int g = 0; g < TC_SETS; ++g) { mbar_init(bar_tfull + 8 * g, 1); mbar_init(bar_tfree + 8 * g, TC_DRAINERS); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  }
  // ---- resident weights, split into TF32 hi / lo: B[kh][k/4][n = kw*32 + out channel][k%4] ---------------
  {
    const float* W = p.w + (long long)task * p.wstride;       // [co][ci][3][3]
    float wscale = 1.f;
    if (F16) {                                                // power-of-two scale from the block's absolute maximum
      if (tid == 0) { *wmax = 0u; smax[0] = smax[1] = smax[2] = smax[3] = 0u; }
      __syncthreads();
      float m = 0.f;
      for (int i = tid; i < 32 * 32 * 9; i += TC_THREADS) {
        const int tap = i % 9, b = (i / 9) % 32, a = i / (9 * 32);
        m = fmaxf(m, fabsf(__ldg(W + ((long long)(p.w_ao + a) * p.w_cin + p.w_bo + b) * 9 + tap)));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      if (lane == 0) atomicMax(wmax, __float_as_uint(m));
      __syncthreads();
      const int kb = f16_scale_exp(__uint_as_float(*wmax));
      if (tid == 0) *b_exp = kb;
      wscale = exp2i(kb);
    }
    // Threads walk the DESTINATION linearly (consecutive lanes -> consecutive shared-memory words: conflict-free stores;
    // walking the source instead puts the 32 lanes of a store into 4 banks) and gather the source element through L2.
    const float* Wb = W + ((long long)p.w_ao * p.w_cin + p.w_bo) * 9;      // W[w_ao + a][w_bo + b][tap] = Wb[(a * w_cin + b) * 9 + tap]
    if (F16) {
      // B[tap][k/8][n][k%8] halfs, two consecutive k per 4-byte store
      for (int j = tid; j < 32 * 32 * 9 / 2; j += TC_THREADS) {
        const int k2 = j & 3, n = (j >> 2) & 31, g = j >> 7;                // g = tap * 4 + k/8
        const int t2 = g >> 2, k = (g & 3) * 8 + 2 * k2;
        int aa, bb, tap;
        if (p.wmode == 0) { aa = n; bb = k; tap = t2; } else { bb = n; aa = k; tap = 8 - t2; }
        const int step = p.wmode == 0 ? 9 : p.w_cin * 9;                     // k -> k + 1 in the source
        const float* src = Wb + ((long long)aa * p.w_cin + bb) * 9 + tap;
        const float x0 = __ldg(src) * wscale, x1 = __ldg(src + step) * wscale;
        const __half2 h = __floats2half2_rn(x0, x1);
        const float2 back = __half22float2(h);
        reinterpret_cast<__half2*>(Bhi)[j] = h;
        reinterpret_cast<__half2*>(Blo)[j] = __floats2half2_rn((x0 - back.x) * LO_SCALE, (x1 - back.y) * LO_SCALE);
      }
    } else {
      // B[kh][k/4][n = kw*32 + out channel][k%4] floats
#pragma unroll 9
      for (int j = tid; j < 32 * 32 * 9; j += TC_THREADS) {
        const int k4 = j & 3, n96 = (j >> 2) % 96, g = j / 384;             // g = kh * 8 + k/4
        const int kh = g >> 3, k = (g & 7) * 4 + k4, kw = n96 >> 5, n = n96 & 31, t2 = kh * 3 + kw;
        int aa, bb, tap;
        if (p.wmode == 0) { aa = n; bb = k; tap = t2; }                     // forward: n = cout, k = cin
        else { bb = n; aa = k; tap = 8 - t2; }                              // dgrad: n = cin (output), k = cout, flipped taps
        const float v = __ldg(Wb + ((long long)aa * p.w_cin + bb) * 9 + tap);
        const float hi = __uint_as_float(f2tf32(v));
        Bhi[j] = hi;
        Blo[j] = v - hi;
      }
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  tmem_base = *tmem_slot;
  XM_PH("weights staged, roles start");

  if (warp < 7) {
    // ========================================= producers =============================================
    // staged row j of a tile <-> position q0 - 1 - Wp + j  (q0 = first output position of the tile)
    // Register-level software pipeline: the global loads of tile it+1 are issued BEFORE tile it is converted
    // and stored, so the HBM/L2 latency overlaps the shared-memory phase and the wait for the stage.
    constexpr int PR = TC_PRODUCERS / 4;                       // rows staged per pass: a thread moves 8 channels
    const int c8 = tid & 3, jrow = tid >> 2;                   // channel octet (two planes), first row (0..PR-1)
    const float* S = p.src + (long long)task * p.n * p.H * p.W * p.src_cs + p.src_co + c8 * 8;
    auto issue_loads = [&](int it, float4 (&v)[8]) {
      const int qbase = (tile0 + it) * TC_TILE - p.Wp - 1 + jrow;
#pragma unroll
      for (int u = 0; u < 4; ++u) {                              // 4 independent rows: no carried state
        const int j = jrow + PR * u;
        const int px = j < p.R ? pos_to_pixel(p.pm, qbase + PR * u) : -1;
        v[2 * u] = v[2 * u + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
#ifndef XM_TC_NOPROD
        if (px >= 0) ldg256(S + (long long)px * p.src_cs, v[2 * u], v[2 * u + 1]);
#endif
      }
    };
#ifdef XM_TC_TIMING
    long long t_wait = 0, t_load = 0, t0;
#endif
    auto store_tile = [&](int it, const float4 (&v)[8]) {
      const int s = it & 1;
#ifdef XM_TC_TIMING
      t0 = clock64();
#endif
      if (F16) {
        // tile scale: maximum magnitude over everything the producers stage for this tile.  Each warp posts its maximum
        // and arrives on an mbarrier BEFORE waiting for the stage, so the handshake hides behind that wait.
        float m = 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u)
          m = fmaxf(m, fmaxf(fmaxf(fabsf(v[u].x), fabsf(v[u].y)), fmaxf(fabsf(v[u].z), fabsf(v[u].w))));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) {
          atomicMax(&smax[it & 3], __float_as_uint(m));
          mbar_arrive(bar_max + 8 * s);
        }
      }
      if (it >= 2) mbar_wait(bar_sfree + 8 * s, ((it - 2) >> 1) & 1);   // MMAs of tile it-2 have read stage s
#ifdef XM_TC_TIMING
      t_wait += clock64() - t0; t0 = clock64();
#endif
      if (F16) {
        mbar_wait(bar_max + 8 * s, (it >> 1) & 1);
        const int ka = f16_scale_exp(__uint_as_float(smax[it & 3]));
        // slot (it + 2) & 3 was last read for tile it - 2: every producer has since passed two of these waits
        if (tid == 0) { a_exp[it & 3] = ka; smax[(it + 2) & 3] = 0u; }
        const float sc = exp2i(ka);
        unsigned char* hi = Abase + (size_t)(2 * s) * set_bytes + (size_t)c8 * plane;     // plane = channel octet
        unsigned char* lo = hi + set_bytes;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = jrow + PR * u;
          if (j < p.R) {
            uint4 h, l;
            split_f16(v[2 * u], v[2 * u + 1], sc, h, l);
            *reinterpret_cast<uint4*>(hi + (size_t)j * 16) = h;
            *reinterpret_cast<uint4*>(lo + (size_t)j * 16) = l;
          }
        }
        fence_proxy_async();
        mbar_arrive(bar_full + 8 * s);
#ifdef XM_TC_TIMING
        t_load += clock64() - t0;
#endif
        return;
      }
      unsigned char* hi = Abase + (size_t)(2 * s) * set_bytes + (size_t)(2 * c8) * plane;
      unsigned char* lo = hi + set_bytes;
#ifndef XM_TC_NOPROD
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = jrow + PR * u;
        if (j < p.R) {
          float4 h, l;
          split_tf32_fast(v[2 * u], h, l);
          *reinterpret_cast<float4*>(hi + (size_t)j * 16) = h;
          *reinterpret_cast<float4*>(lo + (size_t)j * 16) = l;
          split_tf32_fast(v[2 * u + 1], h, l);
          *reinterpret_cast<float4*>(hi + plane + (size_t)j * 16) = h;
          *reinterpret_cast<float4*>(lo + plane + (size_t)j * 16) = l;
        }
      }
#endif
      fence_proxy_async();
      mbar_arrive(bar_full + 8 * s);
#ifdef XM_TC_TIMING
      t_load += clock64() - t0;
#endif
    };
    // two register buffers used alternately (loop unrolled by two: no register moves that would wait on loads)
    float4 va[8], vb[8];
    if (ntiles > 0) issue_loads(0, va);
#ifdef XM_TC_TIMING
    long long t_iss = 0, t1;
#define XM_T_ISS(stmt) do { t1 = clock64(); stmt; t_iss += clock64() - t1; } while (0)
#else
#define XM_T_ISS(stmt) do { stmt; } while (0)
#endif
    for (int it = 0; it < ntiles; it += 2) {
      if (it + 1 < ntiles) XM_T_ISS(issue_loads(it + 1, vb));
      store_tile(it, va);
      if (it == 0) XM_PH("first tile staged");
      if (it + 1 < ntiles) {
        if (it + 2 < ntiles) XM_T_ISS(issue_loads(it + 2, va));
        store_tile(it + 1, vb);
      }
    }
#ifdef XM_TC_TIMING
    if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0)
      printf("producer: tiles %d wait %lld store %lld issue-loads %lld (per tile %lld / %lld / %lld)\n", ntiles, t_wait,
             t_load, t_iss, t_wait / ntiles, t_load / ntiles, t_iss / ntiles);
#endif
  } else if (warp >= 8) {
    // ========================================== drainers =============================================
    // Accumulator row i (TMEM lane) holds, for position q0 - 1 + i, the three kw partial sums
    // D[i][kw*32 + c]; output position q0 + i = D[i][kw=0] + D[i+1][kw=1] + D[i+2][kw=2]: the kw = 1 / 2 blocks
    // come from the next two lanes (shuffles; the last two lanes of a warp take them from the next warp through
    // a small shared-memory exchange).  Rows 126 and 127 of a tile produce no output.
    const int quarter = warp & 3;                              // TMEM lane quarter of this warp (= warp id % 4)
    const int group = (warp - 8) >> 2;
    const int row = quarter * 32 + lane;
    // BatchNorm statistics: after each tile the warp's 32 staged rows are summed per channel (lane l = channel l: 2
    // doubles of state instead of 64 per-thread fp32 accumulators -- this is what lets 16 warps fit the register file)
    double dsum = 0.0, dsq = 0.0;
    // per-warp staging of the finished rows: stores and statistics read it back TRANSPOSED (8 lanes per row for the
    // stores: 128 contiguous bytes per row and instruction; lane = channel for the column sums) -- no shuffles
    float4* wst4 = reinterpret_cast<float4*>(stg_base + (size_t)(warp - 8) * TC_STG_BYTES);
    const float* wst = reinterpret_cast<const float*>(wst4);
    long long* wso = reinterpret_cast<long long*>(wst4 + 32 * 8);
#ifdef XM_TC_TIMING
    long long t_wait = 0, t_tmem = 0, t_rest = 0, t0, t_bar = 0, t_store = 0, t_stat = 0;
#endif
    for (int it = group; it < ntiles; it += TC_GROUPS) {
      (void)0;
#ifdef XM_TC_TIMING
      t0 = clock64();
#endif
      const int q = (tile0 + it) * TC_TILE + row;
      const int px = row < TC_TILE ? pos_to_pixel(p.pm, q) : -1;
      const bool valid = px >= 0;
      const long long o = ((long long)task * p.n * p.H * p.W + (valid ? px : 0)) * p.out_cs + p.out_co;
      // tangent statistics multiply by one 128 B aux row per accumulator row: pull it into L1 while the warp waits
      // for the accumulators, so the loads in the epilogue do not expose HBM latency inside the drain
      if (p.stat_mode == XM_STAT_SUM_AUX && valid) asm volatile("prefetch.global.L1 [%0];" ::"l"(p.aux + o));
      const int s4 = it & (TC_SETS - 1);
      mbar_wait(bar_tfull + 8 * s4, (it / TC_SETS) & 1);
#ifdef XM_TC_TIMING
      t_wait += clock64() - t0; t0 = clock64();
#endif
      tc_fence_after();
      const int xslot = group * 2 + ((it / TC_GROUPS) & 1);       // boundary-row exchange: two alternating slots per group
      const float unscale = F16 ? exp2i(-(a_exp[it & 3] + *b_exp)) : 1.f;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(s4 * (F16 ? 64 : 96));
      // ---- TMEM phase: both 16-column halves -> registers (kw shifts applied), then the set is released --------
      auto load_half = [&](int half, float (&acc)[16]) {
        // the three kw blocks of this half: all loads in flight, one wait; then kw = 1, kw = 2 and the thread's own kw = 0
        float* xb = xch + (((xslot * 2 + half) * 4 + quarter) * 3) * 16;
        uint32_t r[3][16];
        tmem_ld16_nowait(taddr + 32 + half * 16, r[1]);
        tmem_ld16_nowait(taddr + 64 + half * 16, r[2]);
        tmem_ld16_nowait(taddr + half * 16, r[0]);
        tmem_ld_wait();
#pragma unroll
        for (int blk = 1; blk <= 3; ++blk) {
          const int kw = blk % 3;                                  // 1, 2, 0
          float v[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) v[k] = __uint_as_float(r[kw][k]);
          if (kw == 1) {
            // boundary rows for the previous warp: lane 0 publishes its kw = 1 block
            if (lane == 0) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                reinterpret_cast<float4*>(xb)[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              const float t = __shfl_down_sync(0xffffffffu, v[k], 1);
              acc[k] = lane < 31 ? t : 0.f;
            }
          } else if (kw == 2) {
            if (lane < 2) {                                         // lanes 0 and 1 publish their kw = 2 blocks
#pragma unroll
              for (int k = 0; k < 4; ++k)
                reinterpret_cast<float4*>(xb + 16 + lane * 16)[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              const float t = __shfl_down_sync(0xffffffffu, v[k], 2);
              acc[k] += lane < 30 ? t : 0.f;
            }
          } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) acc[k] += v[k];
          }
        }
      };
      float acc2[2][16];
      if (F16) {
        // unstacked taps: accumulator row i IS output position q0 + i -- columns [0, 32) hi*hi, [32, 64) corrections
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t r1[16], r2[16];
          tmem_ld16_nowait(taddr + 32 + half * 16, r1);
          tmem_ld16_nowait(taddr + half * 16, r2);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 16; ++k)
            acc2[half][k] = fmaf(__uint_as_float(r1[k]), LO_UNSCALE, __uint_as_float(r2[k])) * unscale;
        }
      } else {
        load_half(0, acc2[0]);
        load_half(1, acc2[1]);
      }
      tc_fence_before();
      mbar_arrive(bar_tfree + 8 * s4);                           // this TMEM set may be overwritten
#ifdef XM_TC_TIMING
      t_tmem += clock64() - t0; t0 = clock64();
#endif
      // ---- epilogue: rows -> per-warp staging -> coalesced stores (or vector reductions) and the statistics --------
      // Rows 30 and 31 of a warp still miss the kw = 1 / kw = 2 terms that sit in the NEXT warp's lanes 0 and 1 (published
      // in xch above).  They are staged as they are; the readers below add the missing terms to those two rows after the
      // group barrier -- no separate fix-up round trip through shared memory.
      __syncwarp();                                              // the previous tile's reads of the staging are complete
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float* a4 = &acc2[k >> 2][(k & 3) * 4];
        wst4[lane * 8 + (k ^ (lane & 7))] = valid ? make_float4(a4[0], a4[1], a4[2], a4[3]) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      wso[lane] = valid ? o : -1ll;
      // the group's four warps: boundary rows published (and this warp's staging complete)
      if (!F16) asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory");
      else __syncwarp();
#ifdef XM_TC_TIMING
      t_bar += clock64() - t0; t0 = clock64();
#endif
      // xn = next warp's published rows of a half: [0..15] lane 0 kw=1, [16..31] lane 0 kw=2, [32..47] lane 1 kw=2
      //   row 31 += lane 0's kw=1 block + lane 1's kw=2 block;   row 30 += lane 0's kw=2 block
      const float* xn0 = xch + (((xslot * 2 + 0) * 4 + ((quarter + 1) & 3)) * 3) * 16;
      const float* xn1 = xch + (((xslot * 2 + 1) * 4 + ((quarter + 1) & 3)) * 3) * 16;
#ifndef XM_TC_NOSTORE
      {
        const int c = lane & 7, rsub = lane >> 3;
        float4 v[8];
        long long orow[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {                            // all reads first: one shared-memory round trip
          const int r = 4 * k + rsub;
          v[k] = wst4[r * 8 + (c ^ (r & 7))];
          orow[k] = wso[r];
        }
        if (!F16 && rsub >= 2) {                                 // k = 7: rows 30 (rsub 2) and 31 (rsub 3)
          const float4* xn = reinterpret_cast<const float4*>(c < 4 ? xn0 : xn1) + (c & 3);
          const float4 e = xn[rsub == 3 ? 0 : 4];
          const float4 f = rsub == 3 ? xn[8] : make_float4(0.f, 0.f, 0.f, 0.f);
          v[7].x += e.x + f.x; v[7].y += e.y + f.y; v[7].z += e.z + f.z; v[7].w += e.w + f.w;
        }
        // second (src, w) pair of a call: add onto the first pass' output with fire-and-forget vector reductions
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (orow[k] >= 0) put4(p.out + orow[k] + c * 4, v[k], p.accumulate);
      }
#endif
#ifdef XM_TC_TIMING
      t_store += clock64() - t0; t0 = clock64();
#endif
      if (p.stat_mode) {
        // column sums over the warp's 32 rows (rows that produce no output were staged as zeros): lane = channel
        float s1 = 0.f, s2 = 0.f;
        const int cch = lane >> 2, cin4 = lane & 3;
        float c30 = 0.f, c31 = 0.f;                              // the boundary terms of rows 30 and 31 for this channel
        if (!F16) {
          const float* xn = (lane < 16 ? xn0 : xn1) + (lane & 15);
          const long long o30 = wso[30], o31 = wso[31];
          c30 = o30 >= 0 ? xn[16] : 0.f;
          c31 = o31 >= 0 ? xn[0] + xn[32] : 0.f;
        }
        if (p.stat_mode == XM_STAT_SUM_SQ) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float z = wst[j * 32 + ((cch ^ (j & 7)) << 2) + cin4];
            if (j == 30) z += c30;
            if (j == 31) z += c31;
            s1 += z;
            s2 = fmaf(z, z, s2);
          }
        } else {
#pragma unroll
          for (int j0 = 0; j0 < 32; j0 += 8) {
            float z[8], av[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const long long orow = wso[j0 + j];
              av[j] = __ldg(p.aux + (orow >= 0 ? orow : 0ll) + lane);
              z[j] = wst[(j0 + j) * 32 + ((cch ^ ((j0 + j) & 7)) << 2) + cin4];
              if (j0 + j == 30) z[j] += c30;
              if (j0 + j == 31) z[j] += c31;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) { s1 += z[j]; s2 = fmaf(z[j], av[j], s2); }
          }
        }
        dsum += (double)s1;
        dsq += (double)s2;
      }
#ifdef XM_TC_TIMING
      t_stat += clock64() - t0;
#endif
    }
#ifdef XM_TC_TIMING
    t_rest = clock64() - t0;
    if (blockIdx.x == 0 && blockIdx.y == 0 && (tid == TC_PRODUCERS + 32 || tid == TC_PRODUCERS + 32 + 127)) {
      const int nd = (ntiles - group + TC_GROUPS - 1) / TC_GROUPS;
      printf("drainer tid %d: per DRAINED tile: wait %lld tmem %lld staging+group-barrier %lld stores %lld stats %lld\n", tid,
             t_wait / nd, t_tmem / nd, t_bar / nd, t_store / nd, t_stat / nd);
    }
#endif
    if (p.stat_mode) {
      // lane l holds channel l's sums over this warp's rows of all its tiles
      atomicAdd(&p.stats[((long long)task * 2) * 32 + lane], dsum);
      atomicAdd(&p.stats[((long long)task * 2 + 1) * 32 + lane], dsq);
    }
  } else {
    // ======================================= MMA issuer =================================================
    // per tile: for every kernel row kh and channel octet ks, ONE A tile (the halo at row offset kh*Wp) against
    // the [8 x 96] weight slab of the three kw taps; 3 expansion terms -> 36 MMAs (M=128, N=96, K=8).
#ifdef XM_TC_TIMING
    long long t_wf = 0, t_wt = 0, t_issue = 0, t0;
    unsigned long long ns0, ns1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns0));
    const long long c_begin = clock64();
#endif
    for (int it = 0; it < ntiles; ++it) {
      const int s = it & 1;
#ifdef XM_TC_TIMING
      t0 = clock64();
#endif
      mbar_wait(bar_full + 8 * s, (it >> 1) & 1);
#ifdef XM_TC_TIMING
      t_wf += clock64() - t0; t0 = clock64();
#endif
      const int s4 = it & (TC_SETS - 1);
      if (it >= TC_SETS) mbar_wait(bar_tfree + 8 * s4, (it / TC_SETS - 1) & 1);
#ifdef XM_TC_TIMING
      t_wt += clock64() - t0; t0 = clock64();
#endif
      tc_fence_after();
      if (elect_one_sync()) {
        // descriptor low words (start address | LBO) of the stage; per MMA only the start field moves
        const uint32_t a_base = smem_u32(Abase + (size_t)(2 * s) * set_bytes);
        const uint32_t b_hi0 = umma_desc_lo(smem_u32(Bhi), 96u * 16u), b_lo0 = umma_desc_lo(smem_u32(Blo), 96u * 16u);
        constexpr uint32_t dhi = umma_desc_hi(128u);
        const uint32_t d0 = tmem_base + (uint32_t)(s4 * 96);
        const uint32_t a_hi0 = umma_desc_lo(a_base, (uint32_t)plane);
        const uint32_t a_lo0 = a_hi0 + (uint32_t)(set_bytes >> 4);
        const uint32_t kstep = (uint32_t)(2 * plane) >> 4;          // two channel-group planes per K = 8
        if (F16) {
          // one MMA per (tap, 16-channel K step, expansion term): M = 128, N = 32, K = 16.  The tap's A operand is the
          // staged halo at row offset kh*Wp + kw, so all nine taps accumulate into the SAME 32 columns: accumulator row i
          // is output position q0 + i, and the drain needs no cross-lane kw shift-add (54 MMAs, 2 x 32 TMEM columns).
          const uint32_t bf_hi0 = umma_desc_lo(smem_u32(Bhi), 32u * 16u), bf_lo0 = umma_desc_lo(smem_u32(Blo), 32u * 16u);
          const uint32_t df = tmem_base + (uint32_t)(s4 * 64);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const uint32_t shift = (uint32_t)((tap / 3) * p.Wp + (tap % 3));
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {                          // K = 16 channels = two 8-channel planes
              const uint32_t ao = shift + (uint32_t)ks * kstep;
              const uint32_t bo = (uint32_t)((tap * 4 + 2 * ks) * 32);
              const uint32_t acc = (uint32_t)((tap | ks) != 0);
              umma_f16_lh(df + 32, a_lo0 + ao, dhi, bf_hi0 + bo, dhi, TC_IDESC_F16, acc);
              umma_f16_lh(df + 32, a_hi0 + ao, dhi, bf_lo0 + bo, dhi, TC_IDESC_F16, 1u);
              umma_f16_lh(df, a_hi0 + ao, dhi, bf_hi0 + bo, dhi, TC_IDESC_F16, acc);
            }
          }
        } else {
          // ONE accumulator per tile (96 columns: four TMEM sets instead of two).  The tensor core truncates when it adds,
          // so the 24 small correction products are summed FIRST, among themselves, and the 12 hi*hi products go on top:
          // the corrections then lose at most the one rounding that the drain's separate add used to cost.
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            const uint32_t shift = (uint32_t)(kh * p.Wp);           // 16 B units (one staged row = 16 B per plane)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t ao = shift + (uint32_t)ks * kstep;
              const uint32_t bo = (uint32_t)((kh * 8 + 2 * ks) * 96);   // 16 B units, compile-time
              umma_tf32_lh(d0, a_lo0 + ao, dhi, b_hi0 + bo, dhi, TC_IDESC, (uint32_t)((kh | ks) != 0));
              umma_tf32_lh(d0, a_hi0 + ao, dhi, b_lo0 + bo, dhi, TC_IDESC, 1u);
            }
          }
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            const uint32_t shift = (uint32_t)(kh * p.Wp);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
#ifdef XM_TC_NOMMA
              if (kh | ks) continue;
#endif
              umma_tf32_lh(d0, a_hi0 + shift + (uint32_t)ks * kstep, dhi, b_hi0 + (uint32_t)((kh * 8 + 2 * ks) * 96), dhi, TC_IDESC, 1u);
            }
          }
        }
        umma_commit(bar_sfree + 8 * s);     // shared-memory stage s consumed
        umma_commit(bar_tfull + 8 * s4);    // accumulators of this tile complete
      }
      __syncwarp();
#ifdef XM_TC_TIMING
      t_issue += clock64() - t0;
#endif
    }
#ifdef XM_TC_TIMING
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1));
    if (blockIdx.x == 0 && blockIdx.y == 0 && lane == 0)
      printf("mma: wait-full %lld wait-tfree %lld issue %lld (per tile %lld / %lld / %lld); %lld cycles in %llu ns = %.3f GHz\n",
             t_wf, t_wt, t_issue, t_wf / ntiles, t_wt / ntiles, t_issue / ntiles, clock64() - c_begin, ns1 - ns0,
             (double)(clock64() - c_begin) / (double)(ns1 - ns0));
#endif
  }

  tc_fence_before();
  __syncthreads();
  XM_PH("segment done (all roles)");
  }   // segments

  tc_fence_before();
  __syncthreads();
  if (warp == 7) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
  }
  XM_PH("kernel end");
#ifdef XM_TC_TIMING
  if (blockIdx.x == 0 && tid == 0) {
    printf("phases (ns since entry): alloc issued, [weights staged, first tile staged, segment done]*, end:");
    for (int i = 1; i <= ph_n; ++i) printf(" %llu", ph_t[i] - ph_t[0]);
    printf("\n");
  }
#endif
}

static size_t conv_tc_smem(int Wp, int& R, int& plane_bytes, bool f16) {
  R = 128 + 2 * Wp;
  const int rpad = R | 1;                 // odd row count per plane: conflict-free 16 B stores across planes
  plane_bytes = rpad * 16;
  const int nplanes = f16 ? 4 : 8, bwords = f16 ? 3 * 4 * 96 * 4 : 3 * 8 * 96 * 4;
  return (size_t)2 * bwords * 4 + (size_t)4 * nplanes * plane_bytes + 10 * 8 + (size_t)TC_GROUPS * 2 * 2 * 4 * 3 * 16 * 4 + 16 * 4 + 2 * 8 +
         TC_SETS * 8 + (size_t)TC_GROUPS * 4 * TC_STG_BYTES;
}

// Returns 1 if the call was handled by the tcgen05 path, 0 if the shape is not covered (caller falls back),
// <0 / >0 on error like every entry point.
// {sum v, sum v^2} per (task, channel) of an NHWC tensor, in double: BatchNorm statistics of the 64-channel layers,
// whose output is accumulated over two input-channel passes (sum of squares is not additive over passes).
// aux != NULL: {sum v, sum v*aux} (tangent statistics).
__global__ void __launch_bounds__(256) chan_stats_kernel(const float* __restrict__ z, long long pixels, int C, double* stats,
                                                         const float* __restrict__ aux) {
  extern __shared__ double cs_sh[];                       // [2][C]
  const int task = blockIdx.y, c4n = C / 4, c4 = threadIdx.x % c4n, slot = threadIdx.x / c4n, slots = blockDim.x / c4n;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) cs_sh[i] = 0.0;
  __syncthreads();
  double s[4] = {0.0, 0.0, 0.0, 0.0}, q[4] = {0.0, 0.0, 0.0, 0.0};
  float fs[4] = {0.f, 0.f, 0.f, 0.f}, fq[4] = {0.f, 0.f, 0.f, 0.f};
  const float4* Z = reinterpret_cast<const float4*>(z + (long long)task * pixels * C) + c4;
  const float4* A = aux ? reinterpret_cast<const float4*>(aux + (long long)task * pixels * C) + c4 : Z;
  int run = 0;
  if (slot < slots)
    for (long long px = (long long)blockIdx.x * slots + slot; px < pixels; px += (long long)gridDim.x * slots) {
      const float4 v = __ldg(Z + px * c4n), w = __ldg(A + px * c4n);
      fs[0] += v.x; fs[1] += v.y; fs[2] += v.z; fs[3] += v.w;
      fq[0] = fmaf(v.x, w.x, fq[0]); fq[1] = fmaf(v.y, w.y, fq[1]); fq[2] = fmaf(v.z, w.z, fq[2]); fq[3] = fmaf(v.w, w.w, fq[3]);
      if ((++run & 15) == 0)
#pragma unroll
        for (int k = 0; k < 4; ++k) { s[k] += (double)fs[k]; q[k] += (double)fq[k]; fs[k] = fq[k] = 0.f; }
    }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    atomicAdd(&cs_sh[4 * c4 + k], s[k] + (double)fs[k]);
    atomicAdd(&cs_sh[C + 4 * c4 + k], q[k] + (double)fq[k]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(&stats[(long long)task * 2 * C + i], cs_sh[i]);
}

// Stride-2 layers on the stride-1 kernels: conv_s2(x)[r][c] = conv_s1(x)[2r][2c] (forward: full-resolution pass into
// the workspace, then sub-sampling), and the data gradient of conv_s2 is the stride-1 data gradient of the cotangent
// with zeros inserted between its elements.  4x the arithmetic of a native stride-2 kernel, on a path ~40x faster.
__global__ void subsample2_kernel(const float4* __restrict__ full, float4* __restrict__ out, long long imgs, int H, int W,
                                  int hz, int wz, int c4n) {
  const long long total = imgs * hz * wz * c4n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4n);
    long long r = i / c4n;
    const int x = (int)(r % wz); r /= wz;
    const int y = (int)(r % hz);
    const long long img = r / hz;
    out[i] = __ldg(full + ((img * H + 2 * y) * W + 2 * x) * c4n + c);
  }
}
__global__ void upsample2_kernel(const float4* __restrict__ src, float4* __restrict__ full, long long imgs, int H, int W,
                                 int hz, int wz, int c4n) {
  const long long total = imgs * H * W * c4n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4n);
    long long r = i / c4n;
    const int x = (int)(r % W); r /= W;
    const int y = (int)(r % H);
    const long long img = r / H;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!(x & 1) && !(y & 1) && (y >> 1) < hz && (x >> 1) < wz) v = __ldg(src + ((img * hz + (y >> 1)) * wz + (x >> 1)) * c4n + c);
    full[i] = v;
  }
}

// host-side launcher shared with wgrad_tc.cu
int launch_upsample2(const float* src, float* full, long long imgs, int H, int W, int hz, int wz, int C, cudaStream_t stream) {
  const long long total = imgs * H * W * (C / 4);
  const int nb = (int)((total + 255) / 256 < 65535 ? (total + 255) / 256 : 65535);
  upsample2_kernel<<<nb, 256, 0, stream>>>(reinterpret_cast<const float4*>(src), reinterpret_cast<float4*>(full), imgs, H, W,
                                           hz, wz, C / 4);
  return launched("zero insertion (stride 2)");
}

static bool conv_tc_covers(const XmBlockGeom& g) {
  return g.cin == g.cout && (g.cout == 32 || g.cout == 64) && (g.stride == 1 || g.stride == 2);
}
long long conv_tc_workspace_floats(const XmBlockGeom& g) {
  if (!conv_tc_covers(g) || g.stride != 2) return 0;
  return (long long)g.tasks * g.n * g.hin * g.win * g.cout;       // one full-resolution tensor
}

// Returns 1 if the call was handled by the tcgen05 path, 0 if the shape is not covered (caller falls back),
// <0 / >0 on error like every entry point.
int conv_tc_try(const XmConvArgs* a, cudaStream_t stream) {
  const XmBlockGeom& g = a->g;
  if (a->src_nchw || !conv_tc_covers(g)) return 0;
  const int blocks = g.cout / 32;                            // 32-channel blocks per side
  const bool s2 = g.stride == 2;
  if (s2 && (!a->workspace || a->workspace_bytes < conv_tc_workspace_floats(g) * 4)) return 0;
  // statistics in the drain need ONE pass per output element and (sum of squares) one (src, w) pair; otherwise a
  // streaming pass over the finished output computes them
  const bool multi = blocks > 1 || s2;
  if (!multi && a->src2 && a->stat_mode == XM_STAT_SUM_SQ) return 0;
  int R, plane_bytes;
  const bool f16 = g_precise == 2;
  const size_t smem = conv_tc_smem(g.win + 1, R, plane_bytes, f16);
  auto kern = f16 ? conv_tc_kernel<true> : conv_tc_kernel<false>;
  if (smem > 227 * 1024 || R > TC_PRODUCERS) return 0;      // each producer thread stages <= 4 rows per tile
  ConvTcK p{};
  p.tasks = g.tasks; p.n = g.n; p.H = g.hin; p.W = g.win; p.Hp = g.hin + 1; p.Wp = g.win + 1;
  p.Q = g.n * p.Hp * p.Wp;
  p.pm = make_posmap(g.n, g.hin, g.win);
  p.tiles_per_task = (p.Q + TC_TILE - 1) / TC_TILE;
  p.R = R; p.plane_bytes = plane_bytes;
  p.wmode = a->mode == XM_CONV_FWD ? 0 : 1;
  p.out = a->out; p.aux = a->aux; p.stats = a->stats;
  p.src_cs = p.out_cs = p.w_cin = g.cout;
  // persistent grid when a CTA gets enough tiles to amortise re-staging the weights at a task boundary
  const long long total_tiles = (long long)g.tasks * p.tiles_per_task;
  int ctas = (int)(total_tiles < num_sms() ? total_tiles : num_sms());
  p.ctas_per_task = 0;
  if (total_tiles / ctas < 24 || g.tasks >= num_sms()) {
    int per_task = num_sms() / g.tasks;
    if (per_task < 1) per_task = 1;
    if (per_task > p.tiles_per_task) per_task = p.tiles_per_task;
    p.ctas_per_task = per_task;
    ctas = per_task * g.tasks;
  }
  dim3 grid(ctas);
  XM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  if (a->stat_mode) XM_CUDA(cudaMemsetAsync(a->stats, 0, (size_t)g.tasks * 2 * g.cout * sizeof(double), stream));
  const int npairs = a->src2 ? 2 : 1;
  if (!multi) {
    for (int pair = 0; pair < npairs; ++pair) {
      p.src = pair ? a->src2 : a->src1;
      p.w = pair ? a->w2 : a->w1;
      p.wstride = pair ? a->w2_task_stride : a->w1_task_stride;
      p.accumulate = pair;                                   // second pair adds onto the first pass' output
      p.stat_mode = a->stat_mode;                            // {sum v, sum v*aux} are linear: each pass adds its share
      kern<<<grid, TC_THREADS, smem, stream>>>(p);
      if (int rc = launched("xm_conv(tcgen05)")) return rc;
    }
    return 1;
  }
  // Channel-block passes (64 -> 64: output block ob = sum over source blocks sb of the 32 -> 32 convolution with weight
  // block W[co block][ci block]) and / or stride 2.  Every pass after the first into an output block accumulates with
  // vector red.add.
  const long long imgs = (long long)g.tasks * g.n;
  const int c4n = g.cout / 4;
  const bool fwd = p.wmode == 0;
  p.stat_mode = XM_STAT_NONE;
  p.out = (s2 && fwd) ? a->workspace : a->out;             // stride-2 forward: full-resolution result first
  for (int pair = 0; pair < npairs; ++pair) {
    const float* src = pair ? a->src2 : a->src1;
    if (s2 && !fwd) {                                      // stride-2 data gradient: zero-inserted cotangent
      if (int rc = launch_upsample2(src, a->workspace, imgs, g.hin, g.win, g.hz, g.wz, g.cout, stream)) return rc;
      src = a->workspace;
    }
    p.src = src;
    p.w = pair ? a->w2 : a->w1;
    p.wstride = pair ? a->w2_task_stride : a->w1_task_stride;
    for (int ob = 0; ob < blocks; ++ob)
      for (int sb = 0; sb < blocks; ++sb) {
        p.out_co = 32 * ob; p.src_co = 32 * sb;
        p.w_ao = 32 * (fwd ? ob : sb);                     // weight rows = output channels of the FORWARD conv
        p.w_bo = 32 * (fwd ? sb : ob);
        p.accumulate = (pair > 0 || sb > 0) ? 1 : 0;
        kern<<<grid, TC_THREADS, smem, stream>>>(p);
        if (int rc = launched("xm_conv(tcgen05, channel block)")) return rc;
      }
  }
  long long out_pixels = (long long)g.n * g.hin * g.win;
  if (s2 && fwd) {
    const long long total = imgs * g.hz * g.wz * c4n;
    subsample2_kernel<<<(int)((total + 255) / 256 < 65535 ? (total + 255) / 256 : 65535), 256, 0, stream>>>(
        reinterpret_cast<const float4*>(a->workspace), reinterpret_cast<float4*>(a->out), imgs, g.hin, g.win, g.hz, g.wz, c4n);
    if (int rc = launched("xm_conv(sub-sampling)")) return rc;
    out_pixels = (long long)g.n * g.hz * g.wz;
  }
  if (a->stat_mode) {
    int nb = (2 * num_sms() + g.tasks - 1) / g.tasks;
    if (nb < 1) nb = 1;
    chan_stats_kernel<<<dim3(nb, g.tasks), 256, 2 * g.cout * sizeof(double), stream>>>(
        a->out, out_pixels, g.cout, a->stats, a->stat_mode == XM_STAT_SUM_AUX ? a->aux : nullptr);
    if (int rc = launched("xm_conv(statistics)")) return rc;
  }
  return 1;
}

}  // namespace xm
