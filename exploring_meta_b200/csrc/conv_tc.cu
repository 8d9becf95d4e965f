// conv_tc.cu -- tcgen05 / TMEM implementation of xm_conv for the 32-channel stride-1 layers (the bulk of
// the Mini-ImageNet network: layers 2-4 forward, data-gradient, and their tangent versions).
//
// Formulation ("flattened padded pixels"): the n images of a task are laid out as ONE 1-D sequence of
// positions q = (img, r, c), r in [0, H], c in [0, W], where row r = 0 and column c = W are zero padding
// shared between neighbouring rows / images (Hp = H+1, Wp = W+1).  For output position q, tap (kh, kw) of the
// 3x3 pad-1 stencil reads position q + (kh-1)*Wp + (kw-1): every tap is the SAME 1-D array shifted by a
// constant.  A GEMM tile is therefore 128 consecutive positions (M = 128 rows of one tcgen05.mma), whatever
// the map size -- 42x42, 21x21, 10x10 and 5x5 maps all fill tiles equally well -- and the A operand of tap
// (kh, kw) is the staged halo [q0 - Wp - 1, q0 + 128 + Wp + 1) read at row offset kh*Wp + kw.
//
// Shared-memory operand layout (UMMA canonical K-major, no swizzle): channel-group planes
//     A[c4 = ch/4][row][4 ch]   (16 B per row and plane; rows 16 B apart => SBO = 128 B, LBO = plane stride)
// so a shift by s positions is a shift of the descriptor start address by 16*s bytes -- no im2col copy, no
// per-tap restaging: each input element is written to shared memory once and read by 9 taps x 3 passes.
// Weights sit in the same layout B[tap][c4][cout][4] and stay resident for the CTA's lifetime.
//
// Precision: parity is stated in fp32, so every product is evaluated as a 3-term TF32 expansion
// (a_hi*b_hi + a_lo*b_hi + a_hi*b_lo; hi = rna_tf32(x), lo = x - hi): the producer warps split the
// activations while staging, the MMA thread issues 3 tcgen05.mma (kind::tf32, M=128, N=32, K=8) per K step.
// The tensor core adds into fp32 accumulators with truncation, so long accumulation chains drift; the tile
// therefore uses FOUR TMEM accumulators -- one per kernel row kh for the hi*hi terms (12 MMAs each) and one for
// all small correction terms -- which the epilogue sums with round-to-nearest adds.
//
// Pipeline (warp-specialised, 1 CTA per SM, 160 threads): warps 0-3 stage tile i+1 (global -> split -> smem)
// and run the epilogue of tile i (tcgen05.ld -> NHWC store + BatchNorm statistics); lane 0 of warp 4 issues
// the MMAs of tile i+1 meanwhile.  Two shared-memory stages and two TMEM accumulator sets, mbarrier
// full/free handshakes, tcgen05.commit for completion.
#include "common.cuh"

namespace xm {

constexpr int TC_THREADS = 160;
constexpr int TC_WORKERS = 128;
constexpr int TC_TMEM_COLS = 256;     // 2 stages x 4 accumulators x 32 columns

struct ConvTcK {
  int tasks, n, H, W, Hp, Wp;        // source == output spatial dims (stride 1)
  int Q;                             // positions per task = n*Hp*Wp
  int tiles_per_task;
  int R, plane_bytes;                // staged rows per tile, bytes per channel-group plane
  int wmode;                         // 0 forward, 1 data-gradient (transposed weights, flipped taps)
  int stat_mode, accumulate;
  const float* src; const float* w; long long wstride;
  float* out; const float* aux; double* stats;
};

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

// Shared-memory matrix descriptor, canonical K-major layout without swizzle:
// element (row, k) at start + (row/8)*SBO + (row%8)*16 + (k/4)*LBO + (k%4)*4 bytes.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
  return d;                                     // base_offset = 0, layout_type = SWIZZLE_NONE
}

// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 32
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(TC_IDESC), "r"(accumulate) : "memory");
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

// ---- kernel ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const ConvTcK p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int task = blockIdx.y;
  const int plane = p.plane_bytes, set_bytes = 8 * plane;     // one hi (or lo) set of 8 channel-group planes

  float* Bhi = reinterpret_cast<float*>(smem);                 // [9][8][32][4]
  float* Blo = Bhi + 9 * 8 * 32 * 4;
  unsigned char* Abase = smem + 2 * 9 * 8 * 32 * 4 * 4;        // stage s: hi at s*2*set, lo at (s*2+1)*set
  uint64_t* bars = reinterpret_cast<uint64_t*>(Abase + 4 * set_bytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const uint32_t bar_full = smem_u32(bars), bar_sfree = smem_u32(bars + 2), bar_tfull = smem_u32(bars + 4),
                 bar_tfree = smem_u32(bars + 6);

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_full + 8 * s, TC_WORKERS);
      mbar_init(bar_sfree + 8 * s, 1);
      mbar_init(bar_tfull + 8 * s, 1);
      mbar_init(bar_tfree + 8 * s, TC_WORKERS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // ---- resident weights, split into TF32 hi / lo ---------------------------------------------------------
  {
    const float* W = p.w + (long long)task * p.wstride;       // [co][ci][3][3]
    for (int i = tid; i < 32 * 32 * 9; i += TC_THREADS) {
      const int tap = i % 9, b = (i / 9) % 32, a = i / (9 * 32);   // element W[a][b][tap]
      const float v = __ldg(W + i);
      int n, k, t2;
      if (p.wmode == 0) { n = a; k = b; t2 = tap; }           // forward: n = cout, k = cin
      else { n = b; k = a; t2 = 8 - tap; }                    // dgrad: n = cin (output), k = cout, flipped taps
      const int idx = ((t2 * 8 + (k >> 2)) * 32 + n) * 4 + (k & 3);
      const float hi = __uint_as_float(f2tf32(v));
      Bhi[idx] = hi;
      Blo[idx] = v - hi;
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int ntiles = (p.tiles_per_task - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // my tiles
  const int HpWp = p.Hp * p.Wp;

  if (warp < 4) {
    // =============================== producer + epilogue warps ======================================
    const int c4 = tid & 7, jrow = tid >> 3;
    const float* S = p.src + (long long)task * p.n * p.H * p.W * 32;
    float ssum[32], ssq[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) ssum[c] = ssq[c] = 0.f;

    auto stage = [&](int it) {
      const int s = it & 1;
      const int q0 = ((int)blockIdx.x + it * (int)gridDim.x) * 128;
      unsigned char* hi = Abase + (size_t)(2 * s) * set_bytes + (size_t)c4 * plane;
      unsigned char* lo = hi + set_bytes;
      const int qb = q0 - p.Wp - 1;
      for (int j0 = jrow; j0 < p.R; j0 += 64) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = j0 + 16 * u;
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          const int q = qb + j;
          if (j < p.R && q >= 0 && q < p.Q) {
            const int img = q / HpWp, rem = q - img * HpWp;
            const int r = rem / p.Wp, c = rem - r * p.Wp;
            if (r >= 1 && c < p.W)
              v[u] = __ldg(reinterpret_cast<const float4*>(S + (((long long)img * p.H + (r - 1)) * p.W + c) * 32) + c4);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = j0 + 16 * u;
          if (j < p.R) {
            float4 h, l;
            h.x = __uint_as_float(f2tf32(v[u].x)); l.x = v[u].x - h.x;
            h.y = __uint_as_float(f2tf32(v[u].y)); l.y = v[u].y - h.y;
            h.z = __uint_as_float(f2tf32(v[u].z)); l.z = v[u].z - h.z;
            h.w = __uint_as_float(f2tf32(v[u].w)); l.w = v[u].w - h.w;
            *reinterpret_cast<float4*>(hi + (size_t)j * 16) = h;
            *reinterpret_cast<float4*>(lo + (size_t)j * 16) = l;
          }
        }
      }
      fence_proxy_async();
      mbar_arrive(bar_full + 8 * s);
    };

    if (ntiles > 0) stage(0);
    for (int it = 0; it < ntiles; ++it) {
      const int s = it & 1;
      if (it + 1 < ntiles) {
        // stage s^1 was last read by the MMAs of tile it-1
        if (it >= 1) mbar_wait(bar_sfree + 8 * (s ^ 1), ((it - 1) >> 1) & 1);
        stage(it + 1);
      }
      mbar_wait(bar_tfull + 8 * s, (it >> 1) & 1);
      tc_fence_after();
      // ---- epilogue: row (32*warp + lane) of the tile --------------------------------------------------
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(s * 128);
      float v[32], t[32];
      tmem_ld32(taddr + 96, v);                      // correction terms
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        tmem_ld32(taddr + 32 * a, t);                // hi*hi terms of kernel row a
#pragma unroll
        for (int c = 0; c < 32; ++c) v[c] += t[c];
      }
      tc_fence_before();
      mbar_arrive(bar_tfree + 8 * s);                // TMEM set s may be overwritten

      const int q = ((int)blockIdx.x + it * (int)gridDim.x) * 128 + warp * 32 + lane;
      if (q < p.Q) {
        const int img = q / HpWp, rem = q - img * HpWp;
        const int r = rem / p.Wp, c = rem - r * p.Wp;
        if (r >= 1 && c < p.W) {
          const long long o = ((((long long)task * p.n + img) * p.H + (r - 1)) * p.W + c) * 32;
          float4* dst = reinterpret_cast<float4*>(p.out + o);
          if (p.accumulate) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float4 old = dst[k];
              v[4 * k] += old.x; v[4 * k + 1] += old.y; v[4 * k + 2] += old.z; v[4 * k + 3] += old.w;
            }
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) dst[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
          if (p.stat_mode == XM_STAT_SUM_SQ) {
#pragma unroll
            for (int c2 = 0; c2 < 32; ++c2) { ssum[c2] += v[c2]; ssq[c2] = fmaf(v[c2], v[c2], ssq[c2]); }
          } else if (p.stat_mode == XM_STAT_SUM_AUX) {
            const float4* ax = reinterpret_cast<const float4*>(p.aux + o);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float4 a4 = __ldg(ax + k);
              ssum[4 * k] += v[4 * k]; ssum[4 * k + 1] += v[4 * k + 1];
              ssum[4 * k + 2] += v[4 * k + 2]; ssum[4 * k + 3] += v[4 * k + 3];
              ssq[4 * k] = fmaf(v[4 * k], a4.x, ssq[4 * k]); ssq[4 * k + 1] = fmaf(v[4 * k + 1], a4.y, ssq[4 * k + 1]);
              ssq[4 * k + 2] = fmaf(v[4 * k + 2], a4.z, ssq[4 * k + 2]); ssq[4 * k + 3] = fmaf(v[4 * k + 3], a4.w, ssq[4 * k + 3]);
            }
          }
        }
      }
    }
    if (p.stat_mode) {
      // per-thread fp32 partials (<= a few hundred terms each) -> double across the CTA -> global atomics
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const double a = warp_sum((double)ssum[c]);
        const double b = warp_sum((double)ssq[c]);
        if (lane == 0) {
          atomicAdd(&p.stats[((long long)task * 2) * 32 + c], a);
          atomicAdd(&p.stats[((long long)task * 2 + 1) * 32 + c], b);
        }
      }
    }
  } else {
    // ======================================= MMA issuer =================================================
    for (int it = 0; it < ntiles; ++it) {
      const int s = it & 1;
      mbar_wait(bar_full + 8 * s, (it >> 1) & 1);
      if (it >= 2) mbar_wait(bar_tfree + 8 * s, ((it - 2) >> 1) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_hi = smem_u32(Abase + (size_t)(2 * s) * set_bytes), a_lo = a_hi + set_bytes;
        const uint32_t b_hi = smem_u32(Bhi), b_lo = smem_u32(Blo);
        const uint32_t d0 = tmem_base + (uint32_t)(s * 128);
        for (int kh = 0; kh < 3; ++kh) {
          for (int kw = 0; kw < 3; ++kw) {
            const uint32_t shift = (uint32_t)(kh * p.Wp + kw) * 16u;
            const uint32_t boff = (uint32_t)((kh * 3 + kw) * 8 * 32 * 16);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t aoff = (uint32_t)(2 * ks) * (uint32_t)plane + shift;
              const uint32_t bo = boff + (uint32_t)(2 * ks * 32 * 16);
              const uint64_t ah = umma_desc(a_hi + aoff, (uint32_t)plane, 128u);
              const uint64_t al = umma_desc(a_lo + aoff, (uint32_t)plane, 128u);
              const uint64_t bh = umma_desc(b_hi + bo, 512u, 128u);
              const uint64_t bl = umma_desc(b_lo + bo, 512u, 128u);
              const uint32_t first_corr = (kh | kw | ks) != 0;
              umma_tf32(d0 + 96, al, bh, first_corr);
              umma_tf32(d0 + 96, ah, bl, 1u);
              umma_tf32(d0 + 32 * kh, ah, bh, (uint32_t)((kw | ks) != 0));
            }
          }
        }
        umma_commit(bar_sfree + 8 * s);     // shared-memory stage s consumed
        umma_commit(bar_tfull + 8 * s);     // accumulators of this tile complete
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
  }
}

static size_t conv_tc_smem(int Wp, int& R, int& plane_bytes) {
  R = 128 + 2 * Wp + 2;
  const int rpad = R | 1;                 // odd row count per plane: conflict-free 16 B stores across planes
  plane_bytes = rpad * 16;
  return (size_t)2 * 9 * 8 * 32 * 4 * 4 + (size_t)4 * 8 * plane_bytes + 8 * 8 + 16;
}

// Returns 1 if the call was handled by the tcgen05 path, 0 if the shape is not covered (caller falls back),
// <0 / >0 on error like every entry point.
int conv_tc_try(const XmConvArgs* a, cudaStream_t stream) {
  const XmBlockGeom& g = a->g;
  if (g.cin != 32 || g.cout != 32 || g.stride != 1 || a->src_nchw) return 0;
  int R, plane_bytes;
  const size_t smem = conv_tc_smem(g.win + 1, R, plane_bytes);
  if (smem > 227 * 1024) return 0;
  ConvTcK p{};
  p.tasks = g.tasks; p.n = g.n; p.H = g.hin; p.W = g.win; p.Hp = g.hin + 1; p.Wp = g.win + 1;
  p.Q = g.n * p.Hp * p.Wp;
  p.tiles_per_task = (p.Q + 127) / 128;
  p.R = R; p.plane_bytes = plane_bytes;
  p.wmode = a->mode == XM_CONV_FWD ? 0 : 1;
  p.out = a->out; p.aux = a->aux; p.stats = a->stats;
  int per_task = num_sms() / g.tasks;
  if (per_task < 1) per_task = 1;
  if (per_task > p.tiles_per_task) per_task = p.tiles_per_task;
  dim3 grid(per_task, g.tasks);
  static bool attr_set = false;
  if (!attr_set) {
    XM_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  if (a->stat_mode) XM_CUDA(cudaMemsetAsync(a->stats, 0, (size_t)g.tasks * 2 * 32 * sizeof(double), stream));
  const int npairs = a->src2 ? 2 : 1;
  for (int pair = 0; pair < npairs; ++pair) {
    p.src = pair ? a->src2 : a->src1;
    p.w = pair ? a->w2 : a->w1;
    p.wstride = pair ? a->w2_task_stride : a->w1_task_stride;
    p.accumulate = pair;                                   // second pair adds onto the first pass' output
    p.stat_mode = (pair == npairs - 1) ? a->stat_mode : 0; // statistics of the final values only
    conv_tc_kernel<<<grid, TC_THREADS, smem, stream>>>(p);
    if (int rc = launched("xm_conv(tcgen05)")) return rc;
  }
  return 1;
}

}  // namespace xm
