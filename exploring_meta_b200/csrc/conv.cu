// conv.cu -- xm_conv: task-batched 3x3 pad-1 convolution (forward and data-gradient) as an implicit
// GEMM on the tensor cores with per-task weights.
//
// GEMM view per task: M = n*oh*ow output pixels, N = output channels, K = 9 * source channels
// (x2 when a second (src, w) pair is given: zdot = conv(xdot, W) + conv(x, Wdot)).
//   * A CTA owns one task and one 32-wide slice of the output channels; its weights (all K rows)
//     are transposed into shared memory ONCE and stay resident while the CTA walks over its share
//     of the task's pixel tiles (persistent loop) -- per-task weights are never re-read per tile.
//   * A pixel tile is TI images x TH x TW pixels = 128 GEMM rows (TW, TH, TI powers of two chosen on
//     the host per layer so that small maps -- 10x10, 7x7, 2x2 -- pack several images per tile).
//     The source halo of the tile is staged in shared memory once per 32-channel chunk; the nine
//     taps are shifted views of it (offset table), so every source element is read from L2/HBM once.
//   * Contraction: mma.sync m16n8k8 TF32 with error-compensated operand splitting (3 MMAs per
//     product, fp32 accumulate) -> fp32-level accuracy, which the parity contract is stated in.
//   * Epilogue: NHWC store + per-(task, channel) batch statistics in double ({sum, sum sq} for the
//     BatchNorm that follows, or {sum, sum v*aux} for the tangent pass), block-reduced, one double
//     atomic per channel per CTA.
// Reference ops replaced: aten::conv2d (core_functions/vision_models.py:189), the dgrad half of
// convolution_backward and the conv terms of _convolution_double_backward (vision/maml_vision.py:112).
#include "tile.cuh"

namespace xm {

constexpr int CONV_THREADS = 128;
constexpr int WSTR = 40;           // smem weight row stride (32 + 8): conflict-free B fragments

struct ConvK {
  TileGeo t;
  int tasks, oc;                   // output channels
  int wmode;                       // 0: forward weights, 1: data-gradient (transposed, flipped taps)
  int wt_cin, wt_cout;             // PyTorch weight tensor [wt_cout][wt_cin][3][3]
  int nchunks, npairs, kseg;
  int resident;                    // 1: all K rows of the weights stay in shared memory for the CTA's lifetime
  int stat_mode;
  const float* src[2];
  const float* w[2];
  long long wstride[2];
  float* out;
  const float* aux;
  double* stats;
};

template <int PRECISE>
__global__ void __launch_bounds__(CONV_THREADS)
conv_kernel(const ConvK p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const TileGeo& tg = p.t;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int task = blockIdx.y;
  const int co0 = blockIdx.z * 32;
  const int ncols = min(32, p.oc - co0);
  const int nseg = p.npairs * p.nchunks;
  const int TW = 1 << tg.tw_log, TH = 1 << tg.th_log;
  const int halo_px = (1 << tg.ti_log) * tg.halo_h * tg.halo_w;

  double* sstat = reinterpret_cast<double*>(smem_raw);            // [2][32]
  float* ws = reinterpret_cast<float*>(sstat + 64);               // [nseg*kseg][WSTR]
  float* halo = ws + (size_t)(p.resident ? nseg : 1) * p.kseg * WSTR;   // [halo_px][cstride]  (16B aligned)
  int* offtab = reinterpret_cast<int*>(halo + (size_t)halo_px * tg.cstride);  // [kseg]

  // ---- weights: ws[(seg, tap, cl)][co], transposed from PyTorch's [co][ci][3][3] -----------------
  auto load_weights = [&](int seg, float* wseg) {
    const int pair = seg / p.nchunks, chunk = seg - pair * p.nchunks;
    const int c0 = chunk * 32, cc = min(32, tg.sc - c0);
    const float* W = p.w[pair] + (long long)task * p.wstride[pair];
    if (p.wmode == 0) {
      // forward: source channel = conv cin, output channel = conv cout
      const int per_co = cc * 9;
      for (int i = tid; i < ncols * per_co; i += CONV_THREADS) {
        const int col = i / per_co, r = i - col * per_co;
        const int cl = r / 9, tap = r - cl * 9;
        wseg[(tap * cc + cl) * WSTR + col] = __ldg(W + ((long long)(co0 + col) * p.wt_cin + c0) * 9 + r);
      }
    } else {
      // data gradient: source channel = conv cout (chunked), output channel = conv cin, taps flipped
      const int per_src = ncols * 9;
      for (int i = tid; i < cc * per_src; i += CONV_THREADS) {
        const int cl = i / per_src, r = i - cl * per_src;
        const int col = r / 9, tap = r - col * 9;
        wseg[((8 - tap) * cc + cl) * WSTR + col] =
            __ldg(W + ((long long)(c0 + cl) * p.wt_cin + co0 + col) * 9 + tap);
      }
    }
  };
  for (int i = tid; i < (p.resident ? nseg : 1) * p.kseg * WSTR; i += CONV_THREADS) ws[i] = 0.f;
  if (tid < 64) sstat[tid] = 0.0;
  __syncthreads();
  if (p.resident)
    for (int seg = 0; seg < nseg; ++seg) load_weights(seg, ws + (size_t)seg * p.kseg * WSTR);

  double st[4][2][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) st[a][b][0] = st[a][b][1] = 0.0;

  const long long out_task = (long long)task * tg.n * tg.oh * tg.ow * p.oc;
  int last_cc = -1;

  // per-lane pixel bases: rows g and g+8 of the two 16-row m-tiles of this warp (tile-invariant)
  int pbase[2][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) pbase[mt][hr] = pixel_base(tg, warp * 32 + mt * 16 + hr * 8 + g);

  for (int tile = blockIdx.x; tile < tg.tiles_per_task; tile += gridDim.x) {
    int i0, h0, w0;
    tile_origin(tg, tile, i0, h0, w0);

    float acc[2][4][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;

    for (int seg = 0; seg < nseg; ++seg) {
      const int pair = seg / p.nchunks, chunk = seg - pair * p.nchunks;
      const int c0 = chunk * 32, cc = min(32, tg.sc - c0);
      __syncthreads();                       // previous consumers of halo/offtab are done
      if (cc != last_cc) {
        build_offtab(tg, offtab, p.kseg, cc, tid, CONV_THREADS);
        last_cc = cc;
      }
      stage_halo(tg, p.src[pair], task, i0, h0, w0, c0, cc, halo, tid, CONV_THREADS);
      if (!p.resident) {                     // weights streamed per segment (very wide layers only)
        if (cc < 32) {
          for (int i = tid; i < p.kseg * WSTR; i += CONV_THREADS) ws[i] = 0.f;
          __syncthreads();
        }
        load_weights(seg, ws);
      }
      __syncthreads();

      // ---- tensor-core contraction over this segment's K rows -----------------------------------
      // The tensor core adds into its fp32 accumulator with truncation, so a long chain of MMAs into
      // one accumulator drifts (~1e-7 * chain length, biased).  Chains are therefore kept to one
      // 32-row K chunk (4 k-steps, small correction terms first) in `cacc`, which is then folded into
      // the running sum `acc` with ordinary round-to-nearest fp32 adds.
      const float* wseg = p.resident ? ws + (size_t)seg * p.kseg * WSTR : ws;
      const int ksteps = (9 * cc + 7) >> 3;
      for (int kc = 0; kc < ksteps; kc += 4) {
        float cacc[2][4][4];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int c = 0; c < 4; ++c) cacc[a][b][c] = 0.f;
        const int kend = min(kc + 4, ksteps);
        for (int ks = kc; ks < kend; ++ks) {
          const int k0 = ks * 8;
          const int o0 = offtab[k0 + t], o1 = offtab[k0 + t + 4];
          uint32_t ah[2][4], al[2][4];
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            const float a0 = halo[pbase[mt][0] + o0], a1 = halo[pbase[mt][1] + o0];
            const float a2 = halo[pbase[mt][0] + o1], a3 = halo[pbase[mt][1] + o1];
            if (PRECISE) {
              split_tf32(a0, ah[mt][0], al[mt][0]); split_tf32(a1, ah[mt][1], al[mt][1]);
              split_tf32(a2, ah[mt][2], al[mt][2]); split_tf32(a3, ah[mt][3], al[mt][3]);
            } else {
              ah[mt][0] = f2tf32(a0); ah[mt][1] = f2tf32(a1); ah[mt][2] = f2tf32(a2); ah[mt][3] = f2tf32(a3);
            }
          }
          const float* wk0 = wseg + (k0 + t) * WSTR + g;
          const float* wk1 = wk0 + 4 * WSTR;
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            if (nt * 8 < ncols) {
              uint32_t bh[2], bl[2];
              const float b0 = wk0[nt * 8], b1 = wk1[nt * 8];
              if (PRECISE) { split_tf32(b0, bh[0], bl[0]); split_tf32(b1, bh[1], bl[1]); }
              else { bh[0] = f2tf32(b0); bh[1] = f2tf32(b1); }
#pragma unroll
              for (int mt = 0; mt < 2; ++mt) {
                if (PRECISE) {
                  mma_tf32(cacc[mt][nt], al[mt], bh);
                  mma_tf32(cacc[mt][nt], ah[mt], bl);
                }
                mma_tf32(cacc[mt][nt], ah[mt], bh);
              }
            }
          }
        }
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[a][b][c] += cacc[a][b][c];
      }
    }

    // ---- epilogue: NHWC store + batch statistics --------------------------------------------------
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int hr = 0; hr < 2; ++hr) {
        const int px = warp * 32 + mt * 16 + hr * 8 + g;
        const int pw = px & (TW - 1), ph = (px >> tg.tw_log) & (TH - 1), ti = px >> (tg.tw_log + tg.th_log);
        const int img = i0 + ti, h = h0 + ph, w = w0 + pw;
        if (img < tg.n && h < tg.oh && w < tg.ow) {
          const long long o = out_task + (((long long)img * tg.oh + h) * tg.ow + w) * p.oc + co0;
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const int col = nt * 8 + 2 * t + j;
              if (col < ncols) {
                const float v = acc[mt][nt][hr * 2 + j];
                p.out[o + col] = v;
                if (p.stat_mode) {
                  const double dv = (double)v;
                  const double other = (p.stat_mode == XM_STAT_SUM_SQ) ? dv : (double)__ldg(p.aux + o + col);
                  st[nt][j][0] += dv;
                  st[nt][j][1] += dv * other;
                }
              }
            }
        }
      }
  }

  if (p.stat_mode) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          double v = st[nt][j][s];
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          if (g == 0) atomicAdd(&sstat[s * 32 + nt * 8 + 2 * t + j], v);
        }
    __syncthreads();
    if (tid < 64) {
      const int s = tid >> 5, col = tid & 31;
      if (col < ncols) atomicAdd(&p.stats[((long long)task * 2 + s) * p.oc + co0 + col], sstat[tid]);
    }
  }
}

int g_precise = 1;
int g_use_tc = 1;
int conv_tc_try(const XmConvArgs* a, cudaStream_t stream);
long long conv_tc_workspace_floats(const XmBlockGeom& g);
int conv_img_try(const XmConvArgs* a, cudaStream_t stream);

}  // namespace xm

using namespace xm;

extern "C" int xm_set_precision(int precise) {
  g_precise = precise < 0 ? 0 : (precise > 2 ? 2 : precise);
  return 0;
}

extern "C" int xm_set_tcgen05(int enable) {
  g_use_tc = enable;
  return 0;
}

extern "C" int64_t xm_conv_workspace_bytes(const XmBlockGeom* g) {
  if (!g || !geom_ok(*g)) return -1;
  return conv_tc_workspace_floats(*g) * (int64_t)sizeof(float);
}

extern "C" int xm_conv(const XmConvArgs* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  XM_REQUIRE(a != nullptr, "xm_conv: null args");
  const XmBlockGeom& g = a->g;
  XM_REQUIRE(geom_ok(g), "xm_conv: inconsistent block geometry");
  XM_REQUIRE(a->mode == XM_CONV_FWD || a->mode == XM_CONV_DGRAD, "xm_conv: bad mode %d", a->mode);
  XM_REQUIRE(a->src1 && a->w1 && a->out, "xm_conv: null src1/w1/out");
  XM_REQUIRE(a->src2 == nullptr || a->w2 != nullptr, "xm_conv: src2 without w2");
  XM_REQUIRE(a->stat_mode >= 0 && a->stat_mode <= 2, "xm_conv: bad stat_mode");
  XM_REQUIRE(a->stat_mode == 0 || a->stats, "xm_conv: stat_mode without stats buffer");
  XM_REQUIRE(a->stat_mode != XM_STAT_SUM_AUX || a->aux, "xm_conv: SUM_AUX without aux");
  XM_REQUIRE(!(a->src_nchw && a->mode != XM_CONV_FWD), "xm_conv: src_nchw is forward-only");
  XM_REQUIRE(!a->src_nchw || (a->row_step > 0 && a->row0 >= 0 &&
             a->row0 + (long long)(g.n - 1) * a->row_step < a->rows_per_task), "xm_conv: bad image row selection");

  if (g_use_tc == 1) {
    // image layer (cin <= 4, stride 1): exact-fp32 CUDA-core kernel, HBM-write bound (conv_img.cu)
    const int rc = conv_img_try(a, stream);
    if (rc != 0) return rc == 1 ? 0 : rc;
  }
  if (g_use_tc && g_precise) {
    // 32-channel stride-1 layers run on the tcgen05 / TMEM kernel (conv_tc.cu)
    const int rc = conv_tc_try(a, stream);
    if (rc != 0) return rc == 1 ? 0 : rc;
  }

  ConvK p{};
  TileGeo& t = p.t;
  p.tasks = g.tasks; t.n = g.n;
  p.wt_cin = g.cin; p.wt_cout = g.cout;
  if (a->mode == XM_CONV_FWD) {
    t.oh = g.hz; t.ow = g.wz; p.oc = g.cout; t.sh = g.hin; t.sw = g.win; t.sc = g.cin;
    t.s_eff = g.stride; t.dilate = 1; p.wmode = 0;
  } else {
    t.oh = g.hin; t.ow = g.win; p.oc = g.cin; t.sh = g.hz; t.sw = g.wz; t.sc = g.cout;
    t.s_eff = 1; t.dilate = g.stride; p.wmode = 1;
  }
  t.src_nchw = a->src_nchw; t.row0 = a->row0; t.row_step = a->row_step; t.rows_per_task = a->rows_per_task;
  finish_tile_geo(t, 4);
  p.nchunks = (t.sc + 31) / 32;
  p.npairs = a->src2 ? 2 : 1;
  const int ccmax = t.sc < 32 ? t.sc : 32;
  p.kseg = (9 * ccmax + 7) & ~7;
  p.stat_mode = a->stat_mode;
  p.src[0] = a->src1; p.w[0] = a->w1; p.wstride[0] = a->w1_task_stride;
  p.src[1] = a->src2; p.w[1] = a->w2; p.wstride[1] = a->w2_task_stride;
  p.out = a->out; p.aux = a->aux; p.stats = a->stats;

  const int nseg = p.npairs * p.nchunks;
  const size_t fixed = (size_t)halo_pixels(t) * t.cstride * 4 + (size_t)p.kseg * 4 + 64 * 8 + 16;
  size_t smem = (size_t)nseg * p.kseg * WSTR * 4 + fixed;
  p.resident = 1;
  if (smem > 200 * 1024) {
    p.resident = 0;
    smem = (size_t)p.kseg * WSTR * 4 + fixed;
  }
  XM_REQUIRE(smem <= 227 * 1024, "xm_conv: %zu bytes of shared memory needed (cin=%d cout=%d too large)",
             smem, g.cin, g.cout);
  if (p.stat_mode)
    XM_CUDA(cudaMemsetAsync(a->stats, 0, (size_t)g.tasks * 2 * p.oc * sizeof(double), stream));

  const int cotiles = (p.oc + 31) / 32;
  int per_task = (num_sms() * 4 + g.tasks * cotiles - 1) / (g.tasks * cotiles);
  if (per_task > t.tiles_per_task) per_task = t.tiles_per_task;
  if (per_task < 1) per_task = 1;
  dim3 grid(per_task, g.tasks, cotiles);
  auto kern = g_precise ? conv_kernel<1> : conv_kernel<0>;
  XM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  kern<<<grid, CONV_THREADS, smem, stream>>>(p);
  return launched("xm_conv");
}
