// wgrad.cu -- xm_wgrad: per-task conv weight gradient with the fused SGD / outer-recursion epilogue.
//
// GEMM view per task: gW[k = (tap, ci)][co] = sum over pixels of x[pixel + tap][ci] * g[pixel][co];
// M = 9*cin rows, N = cout, reduction over the n*hz*wz pixels of the task (44,100 at the Mini-ImageNet
// 42x42 layer).  A CTA owns (task, 32-channel cin chunk, 32-wide cout slice, pixel split): it walks
// over its share of the pixel tiles (same 128-pixel tiles + halo staging as xm_conv), accumulates one
// tile's [<=288][32] block in registers (18 m-tiles x 4 n-tiles spread over 4 warps), folds it into fp32
// master accumulators in shared memory after every tile, and writes one partial block at the end.  A second kernel reduces the pixel splits in double and
// applies the axpy epilogue out = base + scale * gW directly in PyTorch's [co][ci][3][3] layout --
// this is where theta' = theta - lr*g (learn2learn maml_update, core_functions/vision.py:13) is fused.
// Contraction: mma.sync m16n8k8 TF32, 3-term error-compensated split (fp32-level accuracy).
#include "tile.cuh"

namespace xm {

extern int g_precise;
extern int g_use_tc;
int wgrad_tc_try(const XmWgradArgs* a, cudaStream_t stream, int* rc_out, int* ctas_out, int* tiles_out);
long long wgrad_tc_partial_floats(const XmBlockGeom& g);

constexpr int WG_THREADS = 128;
constexpr int GSTR = 40;     // smem row stride of the g tile (32 + 8)
constexpr int MAXMT = 5;     // m-tiles per warp: ceil(18 / 4)

struct WgradK {
  TileGeo t;
  int tasks, cout, cin;
  int nchunks, npairs, splits;
  int krows;                   // 9*min(cin,32) rounded up to 16
  int tab_off;                 // byte offset of the tables / master accumulators behind the staging area
  const float* x[2];
  const float* g[2];
  float* partial;              // [task][split][9*cin][cout]
};

template <int PRECISE>
__global__ void __launch_bounds__(WG_THREADS)
wgrad_kernel(const WgradK p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const TileGeo& tg = p.t;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int task = blockIdx.y, split = blockIdx.x;
  const int cotile = blockIdx.z / p.nchunks, chunk = blockIdx.z - cotile * p.nchunks;
  const int co0 = cotile * 32, ncols = min(32, p.cout - co0);
  const int c0 = chunk * 32, cc = min(32, p.cin - c0);
  const int TW = 1 << tg.tw_log, TH = 1 << tg.th_log;
  const int halo_px = (1 << tg.ti_log) * tg.halo_h * tg.halo_w;
  const int n_mt = (9 * cc + 15) >> 4;
  const bool split_pixels = n_mt < 4;      // few K rows (cin 1..5): warps split the pixels instead

  float* halo = reinterpret_cast<float*>(smem_raw);                 // [halo_px][cstride]
  float* gs = halo + (size_t)halo_px * tg.cstride;                  // [128][GSTR]
  int* offtab = reinterpret_cast<int*>(smem_raw + p.tab_off);       // [krows]
  int* pbtab = offtab + p.krows;                                    // [128]
  // fp32 master accumulators [MAXMT*16 slots][128 threads]: the tensor core adds into its accumulator
  // with truncation, so register accumulators only ever hold ONE tile's contribution (48 chained MMAs)
  // and are folded into these with round-to-nearest adds after every tile.
  float* macc = reinterpret_cast<float*>(pbtab + 128);              // [MAXMT*16][WG_THREADS]

  build_offtab(tg, offtab, p.krows, cc, tid, WG_THREADS);
  for (int i = tid; i < 128; i += WG_THREADS) pbtab[i] = pixel_base(tg, i);
  __syncthreads();

  // m-tiles of this warp and the (tile-invariant) halo offsets of its two A rows per m-tile
  int my_mt[MAXMT], offA[MAXMT][2];
  int n_my = 0;
#pragma unroll
  for (int i = 0; i < MAXMT; ++i) {
    const int mt = split_pixels ? i : warp + 4 * i;
    my_mt[i] = mt;
    if (mt < n_mt) n_my = i + 1;
    offA[i][0] = offtab[min(mt * 16 + g, p.krows - 1)];
    offA[i][1] = offtab[min(mt * 16 + g + 8, p.krows - 1)];
  }

  for (int i = 0; i < MAXMT * 16; ++i) macc[i * WG_THREADS + tid] = 0.f;

  for (int tile = split; tile < tg.tiles_per_task; tile += p.splits) {
    int i0, h0, w0;
    tile_origin(tg, tile, i0, h0, w0);
    float acc[MAXMT][4][4];
#pragma unroll
    for (int a = 0; a < MAXMT; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;
    for (int pair = 0; pair < p.npairs; ++pair) {
      __syncthreads();
      stage_halo(tg, p.x[pair], task, i0, h0, w0, c0, cc, halo, tid, WG_THREADS);
      // g tile: [128 pixels][ncols], zero for pixels outside the task's grid
      const float* G = p.g[pair] + (long long)task * tg.n * tg.oh * tg.ow * p.cout;
      for (int i = tid; i < 128 * 32; i += WG_THREADS) {
        const int px = i >> 5, col = i & 31;
        const int pw = px & (TW - 1), ph = (px >> tg.tw_log) & (TH - 1), ti = px >> (tg.tw_log + tg.th_log);
        const int img = i0 + ti, h = h0 + ph, w = w0 + pw;
        float v = 0.f;
        if (col < ncols && img < tg.n && h < tg.oh && w < tg.ow)
          v = __ldg(G + (((long long)img * tg.oh + h) * tg.ow + w) * p.cout + co0 + col);
        gs[px * GSTR + col] = v;
      }
      __syncthreads();

      const int ks_begin = split_pixels ? warp * 4 : 0, ks_end = split_pixels ? warp * 4 + 4 : 16;
      for (int ks = ks_begin; ks < ks_end; ++ks) {
        const int p0 = ks * 8;
        const int pb0 = pbtab[p0 + t], pb1 = pbtab[p0 + t + 4];
        uint32_t bh[4][2], bl[4][2];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const float b0 = gs[(p0 + t) * GSTR + nt * 8 + g], b1 = gs[(p0 + t + 4) * GSTR + nt * 8 + g];
          if (PRECISE) { split_tf32(b0, bh[nt][0], bl[nt][0]); split_tf32(b1, bh[nt][1], bl[nt][1]); }
          else { bh[nt][0] = f2tf32(b0); bh[nt][1] = f2tf32(b1); }
        }
#pragma unroll
        for (int i = 0; i < MAXMT; ++i) {
          if (i < n_my) {
            uint32_t ah[4], al[4];
            const float a0 = halo[pb0 + offA[i][0]], a1 = halo[pb0 + offA[i][1]];
            const float a2 = halo[pb1 + offA[i][0]], a3 = halo[pb1 + offA[i][1]];
            if (PRECISE) {
              split_tf32(a0, ah[0], al[0]); split_tf32(a1, ah[1], al[1]);
              split_tf32(a2, ah[2], al[2]); split_tf32(a3, ah[3], al[3]);
            } else {
              ah[0] = f2tf32(a0); ah[1] = f2tf32(a1); ah[2] = f2tf32(a2); ah[3] = f2tf32(a3);
            }
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
              if (nt * 8 < ncols) {
                if (PRECISE) {
                  mma_tf32(acc[i][nt], al, bh[nt]);
                  mma_tf32(acc[i][nt], ah, bl[nt]);
                }
                mma_tf32(acc[i][nt], ah, bh[nt]);
              }
            }
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < MAXMT; ++i)
      if (i < n_my)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int r = 0; r < 4; ++r) macc[((i * 4 + nt) * 4 + r) * WG_THREADS + tid] += acc[i][nt][r];
  }

  // ---- write the partial block: rows k = tap*cc + cl -> global row tap*cin + c0 + cl ---------------
  float* P = p.partial + ((long long)task * p.splits + split) * 9 * p.cin * p.cout;
  if (split_pixels) {
    // cross-warp reduction through shared memory (reuse the staging area)
    __syncthreads();
    float* red = reinterpret_cast<float*>(smem_raw);    // [4 warps][n_mt*16][32]
#pragma unroll
    for (int i = 0; i < MAXMT; ++i)
      if (i < n_my)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int row = my_mt[i] * 16 + g + (r >> 1) * 8, col = nt * 8 + 2 * t + (r & 1);
            red[(warp * n_mt * 16 + row) * 32 + col] = macc[((i * 4 + nt) * 4 + r) * WG_THREADS + tid];
          }
    __syncthreads();
    for (int i = tid; i < 9 * cc * 32; i += WG_THREADS) {
      const int row = i >> 5, col = i & 31;
      if (col < ncols) {
        float v = 0.f;
        for (int w = 0; w < 4; ++w) v += red[(w * n_mt * 16 + row) * 32 + col];
        const int tap = row / cc, cl = row - tap * cc;
        P[(long long)(tap * p.cin + c0 + cl) * p.cout + co0 + col] = v;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < MAXMT; ++i)
      if (i < n_my)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int row = my_mt[i] * 16 + g + (r >> 1) * 8, col = nt * 8 + 2 * t + (r & 1);
            if (row < 9 * cc && col < ncols) {
              const int tap = row / cc, cl = row - tap * cc;
              P[(long long)(tap * p.cin + c0 + cl) * p.cout + co0 + col] = macc[((i * 4 + nt) * 4 + r) * WG_THREADS + tid];
            }
          }
  }
}


// ---- image layer (cin <= 4, stride 1, NCHW source): CUDA-core kernel ---------------------------------------
// K = 9*cin <= 36 is far too thin for the tensor pipe and the layer is bound by streaming g (cout floats per
// pixel) anyway: lane = cout channel, warps split the pixels of a 4-row band whose zero-padded source halo
// (6 x (W+2) pixels x 4 channels) sits in shared memory; every g element is loaded once (coalesced 128 B per
// warp) and meets its 27 source values through broadcast 16 B shared loads.
constexpr int WI_THREADS = 256;
constexpr int WI_ROWS = 12;
constexpr int WI_UNROLL = 12;

struct WgradImgK {
  int n, H, W, cin, cout, splits;
  int row0, row_step, rows_per_task;
  const float* x; const float* g;
  float* partial;              // [task][split][9*cin][cout]
};

__global__ void __launch_bounds__(WI_THREADS)
wgrad_img_kernel(const WgradImgK p) {
  extern __shared__ __align__(16) float4 sh4[];                 // [WI_ROWS+2][W+2] pixels (x, y, z, w = channels)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int task = blockIdx.y, split = blockIdx.x, co = blockIdx.z * 32 + lane;
  const int Wp = p.W + 2;
  const int bands_per_img = (p.H + WI_ROWS - 1) / WI_ROWS;
  const int nbands = p.n * bands_per_img;
  float acc[9][4];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[t][c] = 0.f;
  const float* G = p.g + (long long)task * p.n * p.H * p.W * p.cout;

  // x bands are double-buffered: band k+1 streams in with cp.async (zero-filled outside the image) while the
  // g rows of band k are being reduced
  const int band_px = (WI_ROWS + 2) * Wp + WI_UNROLL;
  const int chw = p.H * p.W;
  auto issue_band = [&](int band, float4* dst) {
    const int img = band / bands_per_img, y0 = (band - img * bands_per_img) * WI_ROWS;
    const float* X = p.x + ((long long)task * p.rows_per_task + p.row0 + (long long)img * p.row_step) * p.cin * chw;
    for (int i = tid; i < (WI_ROWS + 2) * Wp; i += WI_THREADS) {
      const int yy = i / Wp, xx = i - yy * Wp;
      const int y = y0 - 1 + yy, x = xx - 1;
      const bool in = y >= 0 && y < p.H && x >= 0 && x < p.W;
      const float* src = in ? X + y * p.W + x : X;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < p.cin) cp_async4(reinterpret_cast<float*>(dst + i) + c, src + (in ? (long long)c * chw : 0), in ? 4 : 0);
    }
    cp_async_commit();
  };
  for (int i = tid; i < 2 * band_px; i += WI_THREADS) sh4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  if (split < nbands) issue_band(split, sh4);
  int kbuf = 0;
  for (int band = split; band < nbands; band += p.splits, kbuf ^= 1) {
    const int img = band / bands_per_img, y0 = (band - img * bands_per_img) * WI_ROWS;
    const float4* const cur = sh4 + kbuf * band_px;
    if (band + p.splits < nbands) {
      issue_band(band + p.splits, sh4 + (kbuf ^ 1) * band_px);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const int rows = min(WI_ROWS, p.H - y0);
    const float* Gb = G + ((long long)img * p.H + y0) * p.W * p.cout + co;
    // A warp takes WI_UNROLL consecutive pixels of one row per iteration and issues all their loads before the
    // first FMA, so each warp keeps WI_UNROLL x 128 B in flight (the kernel streams g once and is latency-bound
    // otherwise).  Packed FFMA2: the g value is the broadcast scalar against channel pairs of the x pixel.
    const int segs_per_row = (p.W + WI_UNROLL - 1) / WI_UNROLL;
    for (int seg = warp; seg < rows * segs_per_row; seg += WI_THREADS / 32) {
      const int ry = seg / segs_per_row, x0 = (seg - ry * segs_per_row) * WI_UNROLL;
      const float* Gr = Gb + ((long long)ry * p.W + x0) * p.cout;
      float gv[WI_UNROLL];
#pragma unroll
      for (int u = 0; u < WI_UNROLL; ++u)
        gv[u] = (x0 + u < p.W && co < p.cout) ? __ldg(Gr + (long long)u * p.cout) : 0.f;
      const float4* xrow = cur + ry * Wp + x0;
      // (pixels past the row end carry gv = 0 and read the finite pad / next-row pixels: no predicate, so the
      // unrolled loop shares the overlapping x loads of neighbouring pixels)
#pragma unroll
      for (int u = 0; u < WI_UNROLL; ++u) {
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const float4 xv = xrow[kh * Wp + u + kw];
            ffma2(acc[kh * 3 + kw][0], acc[kh * 3 + kw][1], gv[u], xv.x, xv.y);
            ffma2(acc[kh * 3 + kw][2], acc[kh * 3 + kw][3], gv[u], xv.z, xv.w);
          }
      }
    }
    __syncthreads();          // every warp is done with `cur` before the next iteration refills it
  }
  // cross-warp reduction, then one partial block per (task, split)
  __syncthreads();
  float* red = reinterpret_cast<float*>(sh4);                   // [8 warps][36][32]
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int c = 0; c < 4; ++c) red[(warp * 36 + t * 4 + c) * 32 + lane] = acc[t][c];
  __syncthreads();
  float* P = p.partial + ((long long)task * p.splits + split) * 9 * p.cin * p.cout;
  for (int i = tid; i < 36 * 32; i += WI_THREADS) {
    const int slot = i >> 5, l = i & 31, t = slot >> 2, c = slot & 3;
    const int col = blockIdx.z * 32 + l;
    if (c < p.cin && col < p.cout) {
      float v = 0.f;
      for (int w = 0; w < WI_THREADS / 32; ++w) v += red[(w * 36 + slot) * 32 + l];
      P[(long long)(t * p.cin + c) * p.cout + col] = v;
    }
  }
}

static size_t wgrad_img_smem(const XmBlockGeom& g) {
  size_t smem = 2 * ((size_t)(WI_ROWS + 2) * (g.win + 2) + WI_UNROLL) * 16;     // two x bands (+ pad pixels)
  const size_t red = (size_t)8 * 36 * 32 * 4;
  return smem < red ? red : smem;
}

static int wgrad_img_splits(const XmBlockGeom& g) {
  const int cotiles = (g.cout + 31) / 32;
  int s = wave_ctas((const void*)wgrad_img_kernel, WI_THREADS, wgrad_img_smem(g)) / (g.tasks * cotiles);   // one wave
  const int nbands = g.n * ((g.hin + WI_ROWS - 1) / WI_ROWS);
  if (s > nbands) s = nbands;
  if (s > 64) s = 64;
  if (s < 1) s = 1;
  return s;
}

// out_w[task][co][ci][tap] = base_w + scale * sum_split partial[task][split][tap*cin+ci][co]  (double sum)
// out_b[task][co] = base_b (the conv-bias gradient is analytically zero under train-mode BN).
// ctas > 0: the partials come from a persistent grid of `ctas` CTAs over the flattened (task, tile) list
// (wgrad_tc.cu); task t then owns slots 0 .. last(t) - first(t) of its `splits` slots.
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int cin, int cout,
                                    float* out_w, float* out_b, long long out_stride,
                                    const float* base_w, const float* base_b, long long base_stride,
                                    float scale, int ctas = 0, int tiles_per_task = 0,
                                    int cin_total = 0, int ci_off = 0, int co_off = 0) {
  // (cin_total > 0: the partials are one 32 x 32 channel block of a wider layer's [cout][cin_total][3][3] gradient)
  // 64 outputs per block, the slots of an output split over 4 thread quarters (a small shard of the meta-batch leaves
  // up to 148 / tasks slots per task: a serial sum is a chain of that many L2 round trips)
  __shared__ double part[4][64];
  if (cin_total == 0) cin_total = cin;
  const int task = blockIdx.y;
  const int per = 9 * cin * cout;
  const int li = threadIdx.x & 63, kq = threadIdx.x >> 6;
  const int i = blockIdx.x * 64 + li;
  int used = splits;
  if (ctas > 0) {
    const long long G = (long long)gridDim.y * tiles_per_task;
    const int first = (int)((((long long)task * tiles_per_task + 1) * ctas + G - 1) / G) - 1;
    const int last = (int)((((long long)(task + 1) * tiles_per_task) * ctas + G - 1) / G) - 1;
    used = last - first + 1;
  }
  double s = 0.0;
  if (i < per) {
    const float* P = partial + (long long)task * splits * per + i;
    int k = kq;
    for (; k + 12 < used; k += 16) {                              // four independent loads in flight per thread
      const float a = P[(long long)k * per], b = P[(long long)(k + 4) * per];
      const float c = P[(long long)(k + 8) * per], d = P[(long long)(k + 12) * per];
      s += ((double)a + (double)b) + ((double)c + (double)d);
    }
    for (; k < used; k += 4) s += (double)P[(long long)k * per];
  }
  part[kq][li] = s;
  __syncthreads();
  if (kq == 0 && i < per) {
    s = (part[0][li] + part[1][li]) + (part[2][li] + part[3][li]);
    const int row = i / cout, co = i - row * cout;
    const int tap = row / cin, ci = row - tap * cin;
    const long long o = ((long long)(co_off + co) * cin_total + ci_off + ci) * 9 + tap;
    const float b = base_w ? base_w[(long long)task * base_stride + o] : 0.f;
    out_w[(long long)task * out_stride + o] = b + scale * (float)s;
  }
  const int ib = blockIdx.x * 256 + threadIdx.x;
  if (out_b && ib < cout)
    out_b[(long long)task * out_stride + co_off + ib] = base_b ? base_b[(long long)task * base_stride + co_off + ib] : 0.f;
}

static void wgrad_geo(const XmBlockGeom& g, TileGeo& t) {
  t = TileGeo{};
  t.n = g.n; t.oh = g.hz; t.ow = g.wz; t.sh = g.hin; t.sw = g.win; t.sc = g.cin;
  t.s_eff = g.stride; t.dilate = 1;
  finish_tile_geo(t, 8);
}

static int wgrad_splits(const XmBlockGeom& g, const TileGeo& t) {
  const int per_task_ctas = ((g.cout + 31) / 32) * ((g.cin + 31) / 32);
  int s = (num_sms() * 3 + g.tasks * per_task_ctas - 1) / (g.tasks * per_task_ctas);
  if (s > t.tiles_per_task) s = t.tiles_per_task;
  if (s > 64) s = 64;
  if (s < 1) s = 1;
  return s;
}

}  // namespace xm

using namespace xm;

extern "C" int64_t xm_wgrad_scratch_bytes(const XmBlockGeom* g) {
  if (!g || !geom_ok(*g)) return -1;
  TileGeo t;
  wgrad_geo(*g, t);
  int64_t floats = (int64_t)g->tasks * wgrad_splits(*g, t) * 9 * g->cin * g->cout;
  const int64_t tc = wgrad_tc_partial_floats(*g);
  if (tc > floats) floats = tc;
  const int64_t im = (int64_t)g->tasks * wgrad_img_splits(*g) * 9 * g->cin * g->cout;
  if (g->cin <= 4 && g->stride == 1 && im > floats) floats = im;
  return floats * (int64_t)sizeof(float);
}

extern "C" int xm_wgrad(const XmWgradArgs* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  XM_REQUIRE(a != nullptr, "xm_wgrad: null args");
  const XmBlockGeom& g = a->g;
  XM_REQUIRE(geom_ok(g), "xm_wgrad: inconsistent block geometry");
  XM_REQUIRE(a->x1 && a->g1 && a->out_w && a->partial, "xm_wgrad: null x1/g1/out_w/partial");
  XM_REQUIRE((a->x2 == nullptr) == (a->g2 == nullptr), "xm_wgrad: x2/g2 must both be given");
  XM_REQUIRE(!a->src_nchw || (a->row_step > 0 && a->row0 >= 0 &&
             a->row0 + (long long)(g.n - 1) * a->row_step < a->rows_per_task), "xm_wgrad: bad image row selection");
  XM_REQUIRE(!(a->src_nchw && a->x2), "xm_wgrad: image sources carry no tangent (x2 must be NULL)");
  if (g_use_tc && g_precise) {
    // 32-channel stride-1 layers run on the tcgen05 / TMEM kernel (wgrad_tc.cu)
    XM_REQUIRE(a->partial_bytes >= wgrad_tc_partial_floats(g) * 4, "xm_wgrad: partial buffer too small");
    int rc = 0, tc_ctas = 0, tc_tiles = 0;
    const int tsplits = wgrad_tc_try(a, stream, &rc, &tc_ctas, &tc_tiles);
    if (tsplits < 0) return rc;
    if (tsplits > 0) {
      const int per = 9 * 32 * 32, blocks = g.cout / 32;
      dim3 rgrid((per + 63) / 64, g.tasks);
      for (int cb = 0; cb < blocks; ++cb)
        for (int ib = 0; ib < blocks; ++ib) {
          const float* part = a->partial + (long long)(cb * blocks + ib) * g.tasks * tsplits * per;
          wgrad_reduce_kernel<<<rgrid, 256, 0, stream>>>(part, tsplits, 32, 32, a->out_w, a->out_b,
                                                        a->out_task_stride, a->base_w, a->base_b,
                                                        a->base_task_stride, a->scale, tc_ctas, tc_tiles,
                                                        g.cin, 32 * ib, 32 * cb);
          if (int rc2 = launched("xm_wgrad(reduce)")) return rc2;
        }
      return 0;
    }
  }
  if (a->src_nchw && g.cin <= 4 && g.stride == 1 && !a->x2) {
    // image layer: CUDA-core streaming kernel
    WgradImgK k{};
    k.n = g.n; k.H = g.hin; k.W = g.win; k.cin = g.cin; k.cout = g.cout;
    k.splits = wgrad_img_splits(g);
    k.row0 = a->row0; k.row_step = a->row_step; k.rows_per_task = a->rows_per_task;
    k.x = a->x1; k.g = a->g1; k.partial = a->partial;
    XM_REQUIRE(a->partial_bytes >= (int64_t)g.tasks * k.splits * 9 * g.cin * g.cout * 4,
               "xm_wgrad: partial buffer too small");
    const size_t smem = wgrad_img_smem(g);
    XM_REQUIRE(smem <= 200 * 1024, "xm_wgrad: image too wide");
    XM_CUDA(cudaFuncSetAttribute(wgrad_img_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    dim3 grid(k.splits, g.tasks, (g.cout + 31) / 32);
    wgrad_img_kernel<<<grid, WI_THREADS, smem, stream>>>(k);
    if (int rc = launched("xm_wgrad(image)")) return rc;
    const int per = 9 * g.cin * g.cout;
    dim3 rgrid((per + 63) / 64, g.tasks);
    wgrad_reduce_kernel<<<rgrid, 256, 0, stream>>>(a->partial, k.splits, g.cin, g.cout, a->out_w, a->out_b,
                                                  a->out_task_stride, a->base_w, a->base_b,
                                                  a->base_task_stride, a->scale);
    return launched("xm_wgrad(reduce)");
  }
  WgradK p{};
  wgrad_geo(g, p.t);
  p.t.src_nchw = a->src_nchw; p.t.row0 = a->row0; p.t.row_step = a->row_step; p.t.rows_per_task = a->rows_per_task;
  p.tasks = g.tasks; p.cout = g.cout; p.cin = g.cin;
  p.nchunks = (g.cin + 31) / 32;
  p.npairs = a->x2 ? 2 : 1;
  p.splits = wgrad_splits(g, p.t);
  const int ccmax = g.cin < 32 ? g.cin : 32;
  p.krows = (9 * ccmax + 15) & ~15;
  p.x[0] = a->x1; p.g[0] = a->g1; p.x[1] = a->x2; p.g[1] = a->g2;
  p.partial = a->partial;
  const int64_t need = (int64_t)g.tasks * p.splits * 9 * g.cin * g.cout * 4;
  XM_REQUIRE(a->partial_bytes >= need, "xm_wgrad: partial buffer too small (%lld < %lld)",
             (long long)a->partial_bytes, (long long)need);
  size_t stage = (size_t)halo_pixels(p.t) * p.t.cstride * 4 + 128 * GSTR * 4;
  const size_t red = (size_t)4 * p.krows * 32 * 4;       // cross-warp reduction area (small-cin mode)
  if (p.krows < 64 && stage < red) stage = red;
  stage = (stage + 15) & ~(size_t)15;
  p.tab_off = (int)stage;
  size_t smem = stage + (size_t)(p.krows + 128) * 4 + (size_t)MAXMT * 16 * WG_THREADS * 4;
  XM_REQUIRE(smem <= 227 * 1024, "xm_wgrad: %zu bytes of shared memory needed", smem);
  dim3 grid(p.splits, g.tasks, ((g.cout + 31) / 32) * p.nchunks);
  auto kern = g_precise ? wgrad_kernel<1> : wgrad_kernel<0>;
  XM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  kern<<<grid, WG_THREADS, smem, stream>>>(p);
  int rc = launched("xm_wgrad");
  if (rc) return rc;
  const int per = 9 * g.cin * g.cout;
  dim3 rgrid((per + 63) / 64, g.tasks);
  wgrad_reduce_kernel<<<rgrid, 256, 0, stream>>>(a->partial, p.splits, g.cin, g.cout, a->out_w, a->out_b,
                                                a->out_task_stride, a->base_w, a->base_b, a->base_task_stride,
                                                a->scale);
  return launched("xm_wgrad(reduce)");
}
