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
// (a_hi*b_hi + a_lo*b_hi + a_hi*b_lo; hi = rna_tf32(x), lo = x - hi): the producer warps split the
// activations while staging, the MMA thread issues 3 tcgen05.mma (kind::tf32, M=128, N=32, K=8) per K step.
// The tensor core adds into fp32 accumulators with truncation, so long accumulation chains drift; the tile
// therefore uses FOUR TMEM accumulators -- one per kernel row kh for the hi*hi terms (12 MMAs each) and one for
// all small correction terms -- which the epilogue sums with round-to-nearest adds.
//
// Pipeline (warp-specialised, 1 CTA per SM, 416 threads): warps 0-7 stage tiles (global -> TF32 split -> smem),
// warps 8-11 drain accumulators (tcgen05.ld -> NHWC store + BatchNorm statistics; warp w reads TMEM lane
// quarter w%4), one elected lane of warp 12 issues the MMAs.  Two shared-memory stages
// and two TMEM accumulator sets; mbarrier full/free handshakes; tcgen05.commit signals completion.
#include "tc.cuh"

namespace xm {

constexpr int TC_PRODUCERS = 256;     // warps 0-7
constexpr int TC_DRAINERS = 128;      // warps 8-11
constexpr int TC_THREADS = TC_PRODUCERS + TC_DRAINERS + 32;   // + the MMA-issuing warp
constexpr int TC_TMEM_COLS = 256;     // 2 stages x 4 accumulators x 32 columns
constexpr uint32_t TC_IDESC = umma_idesc_tf32(128, 32, 0, 0);   // A and B K-major

struct ConvTcK {
  int tasks, n, H, W, Hp, Wp;        // source == output spatial dims (stride 1)
  int Q;                             // positions per task = n*Hp*Wp
  int tiles_per_task;
  int R, plane_bytes;                // staged rows per tile, bytes per channel-group plane
  int wmode;                         // 0 forward, 1 data-gradient (transposed weights, flipped taps)
  int stat_mode, accumulate;
  int cin, row0, row_step, rows_per_task;   // image layer: source = user images [task][row][c][H][W]
  const float* src; const float* w; long long wstride;
  float* out; const float* aux; double* stats;
};

template <bool IMG>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const ConvTcK p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int task = blockIdx.y;
  constexpr int NPLANES = IMG ? 1 : 8;            // channel-group planes of the A operand
  constexpr int BSLOTS = IMG ? 10 : 72;           // [slot][cout 32][4]: image layer 9 taps + a zero slot
  const int plane = p.plane_bytes, set_bytes = NPLANES * plane;   // one hi (or lo) set

  float* Bhi = reinterpret_cast<float*>(smem);
  float* Blo = Bhi + BSLOTS * 32 * 4;
  unsigned char* Abase = smem + 2 * BSLOTS * 32 * 4 * 4;       // stage s: hi at s*2*set, lo at (s*2+1)*set
  uint64_t* bars = reinterpret_cast<uint64_t*>(Abase + 4 * set_bytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const uint32_t bar_full = smem_u32(bars), bar_sfree = smem_u32(bars + 2), bar_tfull = smem_u32(bars + 4),
                 bar_tfree = smem_u32(bars + 6);

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_full + 8 * s, TC_PRODUCERS);
      mbar_init(bar_sfree + 8 * s, 1);
      mbar_init(bar_tfull + 8 * s, 1);
      mbar_init(bar_tfree + 8 * s, TC_DRAINERS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 12) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // ---- resident weights, split into TF32 hi / lo ---------------------------------------------------------
  {
    const float* W = p.w + (long long)task * p.wstride;       // [co][ci][3][3]
    if (IMG) {
      for (int i = tid; i < BSLOTS * 32 * 4; i += TC_THREADS) { Bhi[i] = 0.f; Blo[i] = 0.f; }
      __syncthreads();
      for (int i = tid; i < 32 * p.cin * 9; i += TC_THREADS) {
        const int tap = i % 9, ci = (i / 9) % p.cin, co = i / (9 * p.cin);
        const float v = __ldg(W + i);
        const int idx = (tap * 32 + co) * 4 + ci;
        const float hi = __uint_as_float(f2tf32(v));
        Bhi[idx] = hi;
        Blo[idx] = v - hi;
      }
    } else {
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
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int ntiles = (p.tiles_per_task - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // my tiles
  const int HpWp = p.Hp * p.Wp;

  if (warp < 8) {
    // ========================================= producers =============================================
    for (int it = 0; it < ntiles; ++it) {
      const int s = it & 1;
      if (it >= 2) mbar_wait(bar_sfree + 8 * s, ((it - 2) >> 1) & 1);   // MMAs of tile it-2 have read stage s
      const int q0 = ((int)blockIdx.x + it * (int)gridDim.x) * 128;
      if (IMG) {
        // one plane: row j <-> position q0 - Wp - 1 + j, 4 floats = (c0, c1, c2, 0) gathered from NCHW planes
        unsigned char* hi = Abase + (size_t)(2 * s) * set_bytes;
        unsigned char* lo = hi + set_bytes;
        for (int j = tid; j < p.R; j += TC_PRODUCERS) {
          const int q = q0 - p.Wp - 1 + j;
          float v[4] = {0.f, 0.f, 0.f, 0.f};
          if (q >= 0 && q < p.Q) {
            const int img = q / HpWp, rem = q - img * HpWp;
            const int r = rem / p.Wp, c = rem - r * p.Wp;
            if (r >= 1 && c < p.W) {
              const float* X = p.src + (((long long)task * p.rows_per_task + p.row0 + (long long)img * p.row_step) * p.cin)
                                       * p.H * p.W + (long long)(r - 1) * p.W + c;
#pragma unroll
              for (int ch = 0; ch < 4; ++ch)
                if (ch < p.cin) v[ch] = __ldg(X + (long long)ch * p.H * p.W);
            }
          }
          float4 h, l;
          h.x = __uint_as_float(f2tf32(v[0])); l.x = v[0] - h.x;
          h.y = __uint_as_float(f2tf32(v[1])); l.y = v[1] - h.y;
          h.z = __uint_as_float(f2tf32(v[2])); l.z = v[2] - h.z;
          h.w = __uint_as_float(f2tf32(v[3])); l.w = v[3] - h.w;
          *reinterpret_cast<float4*>(hi + (size_t)j * 16) = h;
          *reinterpret_cast<float4*>(lo + (size_t)j * 16) = l;
        }
      } else {
        const int c4 = tid & 7, jrow = tid >> 3;                 // channel group, first row (0..31)
        const float* S = p.src + (long long)task * p.n * p.H * p.W * 32;
        unsigned char* hi = Abase + (size_t)(2 * s) * set_bytes + (size_t)c4 * plane;
        unsigned char* lo = hi + set_bytes;
        // position of this thread's first row, then advanced incrementally by 32 positions per row step
        int q = q0 - p.Wp - 1 + jrow;
        int img, r, c;
        if (q >= 0) { img = q / HpWp; const int rem = q - img * HpWp; r = rem / p.Wp; c = rem - r * p.Wp; }
        else { img = -1; r = p.Hp - 1; c = q + p.Wp; if (c < 0) { c += p.Wp; r -= 1; } }   // q >= -Wp-1
        for (int j0 = jrow; j0 < p.R; j0 += 256) {
          float4 v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int j = j0 + 32 * u;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j < p.R && img >= 0 && img < p.n && r >= 1 && c < p.W)
              v[u] = __ldg(reinterpret_cast<const float4*>(S + (((long long)img * p.H + (r - 1)) * p.W + c) * 32) + c4);
            c += 32;
            while (c >= p.Wp) { c -= p.Wp; r += 1; }
            while (r >= p.Hp) { r -= p.Hp; img += 1; }
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int j = j0 + 32 * u;
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
      }
      fence_proxy_async();
      mbar_arrive(bar_full + 8 * s);
    }
  } else if (warp < 12) {
    // ========================================== drainers =============================================
    const int quarter = warp & 3;                              // TMEM lane quarter of this warp
    float ssum[32], ssq[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) ssum[c] = ssq[c] = 0.f;
    for (int it = 0; it < ntiles; ++it) {
      const int s = it & 1;
      mbar_wait(bar_tfull + 8 * s, (it >> 1) & 1);
      tc_fence_after();
      // row (32*quarter + lane) of the tile
      const int q = ((int)blockIdx.x + it * (int)gridDim.x) * 128 + quarter * 32 + lane;
      bool valid = false;
      long long o = 0;
      if (q < p.Q) {
        const int img = q / HpWp, rem = q - img * HpWp;
        const int r = rem / p.Wp, c = rem - r * p.Wp;
        valid = r >= 1 && c < p.W;
        o = ((((long long)task * p.n + img) * p.H + (r - 1)) * p.W + c) * 32;
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(s * 128);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float v[16];
        uint32_t rr[16];
        tmem_ld16_nowait(taddr + 96 + half * 16, rr);            // correction terms first (small)
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = __uint_as_float(rr[k]);
#pragma unroll
        for (int a = 0; a < (IMG ? 1 : 3); ++a) {                // hi*hi terms of kernel rows 0..2
          tmem_ld16_nowait(taddr + 32 * a + half * 16, rr);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 16; ++k) v[k] += __uint_as_float(rr[k]);
        }
        if (half == 1) {
          tc_fence_before();
          mbar_arrive(bar_tfree + 8 * s);                        // TMEM set s may be overwritten
        }
        if (valid) {
          float4* dst = reinterpret_cast<float4*>(p.out + o + half * 16);
          if (p.accumulate) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float4 old = dst[k];
              v[4 * k] += old.x; v[4 * k + 1] += old.y; v[4 * k + 2] += old.z; v[4 * k + 3] += old.w;
            }
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) dst[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
          if (p.stat_mode == XM_STAT_SUM_SQ) {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              ssum[half * 16 + k] += v[k];
              ssq[half * 16 + k] = fmaf(v[k], v[k], ssq[half * 16 + k]);
            }
          } else if (p.stat_mode == XM_STAT_SUM_AUX) {
            const float4* ax = reinterpret_cast<const float4*>(p.aux + o + half * 16);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float4 a4 = __ldg(ax + k);
              const int b = half * 16 + 4 * k;
              ssum[b] += v[4 * k]; ssum[b + 1] += v[4 * k + 1]; ssum[b + 2] += v[4 * k + 2]; ssum[b + 3] += v[4 * k + 3];
              ssq[b] = fmaf(v[4 * k], a4.x, ssq[b]); ssq[b + 1] = fmaf(v[4 * k + 1], a4.y, ssq[b + 1]);
              ssq[b + 2] = fmaf(v[4 * k + 2], a4.z, ssq[b + 2]); ssq[b + 3] = fmaf(v[4 * k + 3], a4.w, ssq[b + 3]);
            }
          }
        }
      }
    }
    if (p.stat_mode) {
      // per-thread fp32 partials (<= a few hundred terms each) -> double across the warp -> global atomics
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
      if (elect_one_sync()) {
        // descriptor low words (start address | LBO) of the stage; per MMA only the start field moves
        const uint32_t a_base = smem_u32(Abase + (size_t)(2 * s) * set_bytes);
        const uint32_t b_hi0 = umma_desc_lo(smem_u32(Bhi), 512u), b_lo0 = umma_desc_lo(smem_u32(Blo), 512u);
        constexpr uint32_t dhi = umma_desc_hi(128u);
        const uint32_t d0 = tmem_base + (uint32_t)(s * 128);
        if (IMG) {
          // K = 8 = two taps x (3 channels + zero); tap pairs (0,1) (2,3) (4,5) (6,7) (8, zero slot)
#pragma unroll
          for (int pr = 0; pr < 5; ++pr) {
            const int t0 = 2 * pr, t1 = 2 * pr + 1;
            const uint32_t off0 = (uint32_t)((t0 / 3) * p.Wp + (t0 % 3));
            const uint32_t off1 = pr < 4 ? (uint32_t)((t1 / 3) * p.Wp + (t1 % 3)) : off0;   // zero slot: LBO = 0
            const uint32_t a_hi = umma_desc_lo(a_base + off0 * 16u, (off1 - off0) * 16u);
            const uint32_t a_lo = a_hi + (uint32_t)(set_bytes >> 4);
            const uint32_t bo = (uint32_t)(t0 * 32);                       // 16 B units: slot t0, next slot at LBO
            umma_tf32_lh(d0 + 96, a_lo, dhi, b_hi0 + bo, dhi, TC_IDESC, (uint32_t)(pr != 0));
            umma_tf32_lh(d0 + 96, a_hi, dhi, b_lo0 + bo, dhi, TC_IDESC, 1u);
            umma_tf32_lh(d0, a_hi, dhi, b_hi0 + bo, dhi, TC_IDESC, (uint32_t)(pr != 0));
          }
        } else {
          const uint32_t a_hi0 = umma_desc_lo(a_base, (uint32_t)plane);
          const uint32_t a_lo0 = a_hi0 + (uint32_t)(set_bytes >> 4);
          const uint32_t kstep = (uint32_t)(2 * plane) >> 4;          // two channel-group planes per K = 8
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              const uint32_t shift = (uint32_t)(kh * p.Wp + kw);      // 16 B units
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                const uint32_t ao = shift + (uint32_t)ks * kstep;
                const uint32_t bo = (uint32_t)(((kh * 3 + kw) * 8 + 2 * ks) * 32);   // 16 B units, compile-time
                umma_tf32_lh(d0 + 96, a_lo0 + ao, dhi, b_hi0 + bo, dhi, TC_IDESC, (uint32_t)((kh | kw | ks) != 0));
                umma_tf32_lh(d0 + 96, a_hi0 + ao, dhi, b_lo0 + bo, dhi, TC_IDESC, 1u);
                umma_tf32_lh(d0 + 32 * kh, a_hi0 + ao, dhi, b_hi0 + bo, dhi, TC_IDESC, (uint32_t)((kw | ks) != 0));
              }
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
  if (warp == 12) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
  }
}

static size_t conv_tc_smem(bool img, int Wp, int& R, int& plane_bytes) {
  R = 128 + 2 * Wp + 2;
  const int rpad = R | 1;                 // odd row count per plane: conflict-free 16 B stores across planes
  plane_bytes = rpad * 16;
  const int nplanes = img ? 1 : 8, bslots = img ? 10 : 72;
  return (size_t)2 * bslots * 32 * 4 * 4 + (size_t)4 * nplanes * plane_bytes + 8 * 8 + 16;
}

// Returns 1 if the call was handled by the tcgen05 path, 0 if the shape is not covered (caller falls back),
// <0 / >0 on error like every entry point.
int conv_tc_try(const XmConvArgs* a, cudaStream_t stream) {
  const XmBlockGeom& g = a->g;
  const bool img = a->src_nchw != 0;
  if (g.cout != 32 || g.stride != 1) return 0;
  if (img) { if (g.cin > 4 || a->mode != XM_CONV_FWD || a->src2) return 0; }
  else if (g.cin != 32) return 0;
  int R, plane_bytes;
  const size_t smem = conv_tc_smem(img, g.win + 1, R, plane_bytes);
  if (smem > 227 * 1024) return 0;
  ConvTcK p{};
  p.tasks = g.tasks; p.n = g.n; p.H = g.hin; p.W = g.win; p.Hp = g.hin + 1; p.Wp = g.win + 1;
  p.Q = g.n * p.Hp * p.Wp;
  p.tiles_per_task = (p.Q + 127) / 128;
  p.R = R; p.plane_bytes = plane_bytes;
  p.wmode = a->mode == XM_CONV_FWD ? 0 : 1;
  p.cin = g.cin; p.row0 = a->row0; p.row_step = a->row_step; p.rows_per_task = a->rows_per_task;
  p.out = a->out; p.aux = a->aux; p.stats = a->stats;
  int per_task = num_sms() / g.tasks;
  if (per_task < 1) per_task = 1;
  if (per_task > p.tiles_per_task) per_task = p.tiles_per_task;
  dim3 grid(per_task, g.tasks);
  static bool attr_set = false;
  if (!attr_set) {
    XM_CUDA(cudaFuncSetAttribute(conv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    XM_CUDA(cudaFuncSetAttribute(conv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
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
    if (img) conv_tc_kernel<true><<<grid, TC_THREADS, smem, stream>>>(p);
    else conv_tc_kernel<false><<<grid, TC_THREADS, smem, stream>>>(p);
    if (int rc = launched("xm_conv(tcgen05)")) return rc;
  }
  return 1;
}

}  // namespace xm
