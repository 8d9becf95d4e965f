// wgrad_tc.cu -- tcgen05 / TMEM implementation of xm_wgrad for the 32-channel stride-1 layers.
//
// gW[tap][ci][co] = sum over positions q of g[q][co] * x[q + delta_tap][ci], on the same flattened
// padded position sequence as conv_tc.cu (g is staged as zero at padding positions).  The reduction
// dimension of the GEMM is the POSITION axis, so both operands are MN-major: position rows of 128 B
// (= the 32 channels of one position, i.e. the natural NHWC row) in the 128B-swizzle/32B-base layout,
// (the swizzle is a function of absolute address bits, so whole-row shifts of a descriptor start address or
// of LBO address the same swizzled data).  A = X halo (M = 4 shifted copies x cin, K = 8 position rows)
// and B = G (N = cout, K = 8 consecutive position rows).  The M = 128 rows of one tcgen05.mma (kind::tf32,
// N = 32, K = 8) are FOUR 32-channel groups whose shared-memory stride is LBO; with LBO = 128 B = one position
// row, group j is the x halo shifted by j positions -- i.e. the kernel-column taps kw = 0, 1, 2 (group 3 is
// unused) of one kernel row kh come out of ONE instruction:
//     D_kh[(kw, ci)][co] += sum_k x[row k0 + k + kh*Wp + kw][ci] * g[row k0 + k][co]
// so a k-step costs 3 MMAs per expansion term instead of 9.  TMEM lane quarter kw holds tap (kh, kw), lane = ci,
// column = co.  Two accumulator sets (2 x 3 x 32 columns) let the drain of tile i overlap the MMAs of tile i+1.
// Accumulators are drained after every tile into fp32 registers with round-to-nearest adds (the tensor
// core's own accumulation truncates), and each CTA writes one partial block per task split;
// wgrad_reduce_kernel (wgrad.cu) reduces the splits in double and applies the fused SGD / outer-recursion
// epilogue.
#include "tc.cuh"

namespace xm {

constexpr int WT_WORKERS = 256;
constexpr int WT_THREADS = WT_WORKERS + 32;
constexpr int WT_TMEM_COLS = 256;            // 2 sets x 3 accumulators x 32 columns (192) -> next power of two
constexpr uint32_t WT_IDESC = umma_idesc_tf32(128, 32, 1, 1);   // A and B MN-major

struct WgradTcK {
  int tasks, n, H, W, Hp, Wp, Q;
  int tiles_per_task, splits, npairs;
  int R, xbuf;                        // staged x rows per tile, bytes of one x buffer (hi or lo)
  int off_x0, off_x1, off_g0, off_g1, off_bar;   // shared-memory byte offsets
  const float* x[2];
  const float* g[2];
  float* partial;                     // [task][split][9*32][32]
};

__global__ void __launch_bounds__(WT_THREADS, 1) wgrad_tc_kernel(const WgradTcK p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int task = blockIdx.y, split = blockIdx.x;
  const int xset = p.xbuf;
  constexpr int gset = 128 * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const uint32_t bar_full = smem_u32(bars), bar_sfree = smem_u32(bars + 2), bar_tfull = smem_u32(bars + 4),
                 bar_tfree = smem_u32(bars + 6);

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_full + 8 * s, WT_WORKERS);
      mbar_init(bar_sfree + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_tfull + 8 * s, 1);
      mbar_init(bar_tfree + 8 * s, 96);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == WT_WORKERS / 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(WT_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // rows past R (read only by the unused 4th lane group) must hold finite values
  for (int i = tid; i < (xset - p.R * 128) / 16; i += WT_THREADS) {
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const size_t o = (size_t)p.R * 128 + (size_t)i * 16;
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

  const int ntiles = (p.tiles_per_task - split + p.splits - 1) / p.splits;
  const int nunits = ntiles * p.npairs;
  const int HpWp = p.Hp * p.Wp;

  if (warp < WT_WORKERS / 32) {
    // =============================== producers (+ warps 0-3: accumulator drain) =========================
    const int c4 = tid & 7, jrow = tid >> 3;
    float macc[3][32];                      // fp32 master accumulators: taps (kh, kw = warp), lane = ci, [co]
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int c = 0; c < 32; ++c) macc[a][c] = 0.f;

    auto split_store = [](const float4& v, unsigned char* hi, unsigned char* lo) {
      float4 h, l;
      h.x = __uint_as_float(f2tf32(v.x)); l.x = v.x - h.x;
      h.y = __uint_as_float(f2tf32(v.y)); l.y = v.y - h.y;
      h.z = __uint_as_float(f2tf32(v.z)); l.z = v.z - h.z;
      h.w = __uint_as_float(f2tf32(v.w)); l.w = v.w - h.w;
      *reinterpret_cast<float4*>(hi) = h;
      *reinterpret_cast<float4*>(lo) = l;
    };

    auto stage = [&](int u) {
      const int s = u & 1, it = u / p.npairs, pair = u - it * p.npairs;
      const int q0 = (split + it * p.splits) * 128;
      const float* X = p.x[pair] + (long long)task * p.n * p.H * p.W * 32;
      const float* G = p.g[pair] + (long long)task * p.n * p.H * p.W * 32;
      // ---- x halo: rows j <-> positions q0 - Wp - 1 + j ------------------------------------------------
      {
        unsigned char* hi = smem + (s ? p.off_x1 : p.off_x0) + (c4 & 1) * 16;
        unsigned char* lo = hi + xset;
        const int ch = c4 >> 1;               // 32 B chunk of the 128 B row; swizzled with the row index
        int q = q0 - p.Wp - 1 + jrow;
        int img, r, c;
        if (q >= 0) { img = q / HpWp; const int rem = q - img * HpWp; r = rem / p.Wp; c = rem - r * p.Wp; }
        else { img = -1; r = p.Hp - 1; c = q + p.Wp; if (c < 0) { c += p.Wp; r -= 1; } }
        for (int j0 = jrow; j0 < p.R; j0 += 128) {
          float4 v[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int j = j0 + 32 * k;
            v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j < p.R && img >= 0 && img < p.n && r >= 1 && c < p.W)
              v[k] = __ldg(reinterpret_cast<const float4*>(X + (((long long)img * p.H + (r - 1)) * p.W + c) * 32) + c4);
            c += 32;
            while (c >= p.Wp) { c -= p.Wp; r += 1; }
            while (r >= p.Hp) { r -= p.Hp; img += 1; }
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int j = j0 + 32 * k;
            if (j < p.R) {
              const size_t o = (size_t)j * 128 + (size_t)((ch ^ (j & 3)) * 32);
              split_store(v[k], hi + o, lo + o);
            }
          }
        }
      }
      // ---- g tile: rows i <-> positions q0 + i, zero at padding positions ----------------------------------
      {
        unsigned char* hi = smem + (s ? p.off_g1 : p.off_g0) + (c4 & 1) * 16;
        unsigned char* lo = hi + gset;
        const int ch = c4 >> 1;
        int q = q0 + jrow;
        int img = q / HpWp;
        const int rem = q - img * HpWp;
        int r = rem / p.Wp, c = rem - r * p.Wp;
        float4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (img < p.n && r >= 1 && c < p.W)
            v[k] = __ldg(reinterpret_cast<const float4*>(G + (((long long)img * p.H + (r - 1)) * p.W + c) * 32) + c4);
          c += 32;
          while (c >= p.Wp) { c -= p.Wp; r += 1; }
          while (r >= p.Hp) { r -= p.Hp; img += 1; }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = jrow + 32 * k;
          const size_t o = (size_t)i * 128 + (size_t)((ch ^ (i & 3)) * 32);
          split_store(v[k], hi + o, lo + o);
        }
      }
      fence_proxy_async();
      mbar_arrive(bar_full + 8 * s);
    };

    if (nunits > 0) stage(0);
    for (int u = 0; u < nunits; ++u) {
      if (u + 1 < nunits) {
        if (u >= 1) mbar_wait(bar_sfree + 8 * ((u + 1) & 1), ((u - 1) >> 1) & 1);
        stage(u + 1);
      }
      const int it = u / p.npairs;
      if (warp < 3 && (u - it * p.npairs) == p.npairs - 1) {
        // ---- drain the tile's accumulators: lane quarter = kw, lane = cin, 32 cout columns per kernel row ----
        const int set = it & 1;
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
    }
    if (warp < 3) {
      float* P = p.partial + ((long long)task * p.splits + split) * 9 * 32 * 32;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        float4* dst = reinterpret_cast<float4*>(P + ((kh * 3 + warp) * 32 + lane) * 32);
#pragma unroll
        for (int c = 0; c < 8; ++c)
          dst[c] = make_float4(macc[kh][4 * c], macc[kh][4 * c + 1], macc[kh][4 * c + 2], macc[kh][4 * c + 3]);
      }
    }
  } else {
    // ======================================= MMA issuer =================================================
    for (int u = 0; u < nunits; ++u) {
      const int s = u & 1, it = u / p.npairs, pair = u - it * p.npairs;
      const int set = it & 1;
      mbar_wait(bar_full + 8 * s, (u >> 1) & 1);
      if (pair == 0 && it >= 2) mbar_wait(bar_tfree + 8 * set, ((it - 2) >> 1) & 1);
      tc_fence_after();
      if (elect_one_sync()) {
        // descriptor low words (start address | LBO); per MMA only the start field moves.  A = x halo with
        // LBO = one position row (lane group j = halo shifted by j rows), B = g tile.
        const uint32_t x_hi0 = umma_desc_lo(smem_u32(smem + (s ? p.off_x1 : p.off_x0)), 128u);
        const uint32_t x_lo0 = x_hi0 + (uint32_t)(xset >> 4);
        const uint32_t g_hi0 = umma_desc_lo(smem_u32(smem + (s ? p.off_g1 : p.off_g0)), 0u);
        const uint32_t g_lo0 = g_hi0 + (uint32_t)(gset >> 4);
        constexpr uint32_t dhi = umma_desc_hi(512u, 1u);            // SBO 512 B, 128B-swizzle / 32B-base layout
        const uint32_t d0 = tmem_base + (uint32_t)(set * 96);
        const uint32_t rowoff = (uint32_t)p.Wp * 8u;                // one kernel row = Wp position rows (16 B units)
        for (int ks = 0; ks < 16; ++ks) {
          const uint32_t fresh = (pair == 0 && ks == 0) ? 0u : 1u;
          const uint32_t ko = (uint32_t)ks * 64u;                   // 8 position rows of 128 B per K step
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            const uint32_t ao = ko + (uint32_t)kh * rowoff;
            const uint32_t d = d0 + (uint32_t)(kh * 32);
            umma_tf32_lh(d, x_lo0 + ao, dhi, g_hi0 + ko, dhi, WT_IDESC, fresh);
            umma_tf32_lh(d, x_hi0 + ao, dhi, g_lo0 + ko, dhi, WT_IDESC, 1u);
            umma_tf32_lh(d, x_hi0 + ao, dhi, g_hi0 + ko, dhi, WT_IDESC, 1u);
          }
        }
        umma_commit(bar_sfree + 8 * s);
        if (pair == p.npairs - 1) umma_commit(bar_tfull + 8 * set);
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == WT_WORKERS / 32) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(WT_TMEM_COLS) : "memory");
  }
}

static bool wgrad_tc_layout(const XmBlockGeom& g, WgradTcK& p, size_t& smem) {
  if (g.cin != 32 || g.cout != 32 || g.stride != 1) return false;
  p.n = g.n; p.H = g.hin; p.W = g.win; p.Hp = g.hin + 1; p.Wp = g.win + 1;
  p.Q = g.n * p.Hp * p.Wp;
  p.tiles_per_task = (p.Q + 127) / 128;
  p.R = 128 + 2 * p.Wp + 2;
  p.xbuf = ((p.R + 1 + 7) & ~7) * 128;               // + the row the unused lane group reads; whole 1 KB units
                                                     // keep every buffer 1024 B aligned
  const int gbuf = 128 * 128;
  p.off_x0 = 0;
  p.off_g0 = p.off_x0 + 2 * p.xbuf;
  p.off_x1 = p.off_g0 + 2 * gbuf;
  p.off_g1 = p.off_x1 + 2 * p.xbuf;
  p.off_bar = p.off_g1 + 2 * gbuf;
  smem = (size_t)p.off_bar + 8 * 8 + 16;
  int splits = num_sms() / g.tasks;
  if (splits < 1) splits = 1;
  if (splits > p.tiles_per_task) splits = p.tiles_per_task;
  p.splits = splits;
  return smem <= 227 * 1024;
}

// partial-buffer floats needed by the tcgen05 path for this geometry (0 if the shape is not covered)
long long wgrad_tc_partial_floats(const XmBlockGeom& g) {
  WgradTcK p{};
  size_t smem;
  if (!wgrad_tc_layout(g, p, smem)) return 0;
  return (long long)g.tasks * p.splits * 9 * 32 * 32;
}

// Returns the number of splits written to a->partial (>0) when handled, 0 when not covered, <0 on error.
int wgrad_tc_try(const XmWgradArgs* a, cudaStream_t stream, int* rc_out) {
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
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { *rc_out = fail((int)e, "cudaFuncSetAttribute(wgrad_tc): %s", cudaGetErrorString(e)); return -1; }
    attr_set = true;
  }
  dim3 grid(p.splits, g.tasks);
  wgrad_tc_kernel<<<grid, WT_THREADS, smem, stream>>>(p);
  if (int rc = launched("xm_wgrad(tcgen05)")) { *rc_out = rc; return -1; }
  return p.splits;
}

}  // namespace xm
