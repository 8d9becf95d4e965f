// tile.cuh -- pixel-tile geometry and shared-memory halo staging shared by xm_conv and xm_wgrad.
#pragma once
#include "common.cuh"

namespace xm {

// A pixel tile is TI images x TH x TW positions of the "output grid" (= 128 GEMM rows); its source
// halo is TI x halo_h x halo_w positions of the source tensor, one <=32-channel chunk at a time.
struct TileGeo {
  int n;                       // images per task
  int oh, ow;                  // output grid
  int sh, sw, sc;              // source tensor dims (per image), channels
  int s_eff;                   // halo step per output pixel (forward: conv stride; dgrad: 1)
  int dilate;                  // 2: source is read as zero-dilated by 2 (dgrad of a stride-2 conv)
  int src_nchw, row0, row_step, rows_per_task;   // source = user images [task][row][c][h][w]
  int tw_log, th_log, ti_log;
  int tiles_w, tiles_h, tiles_i, tiles_per_task;
  int halo_h, halo_w, cstride;
};

// Picks TW/TH/TI (powers of two, product 128) maximising useful pixels per tile, lightly penalising
// halo volume.
inline void pick_tile(int n, int oh, int ow, int s_eff, int& twl, int& thl, int& til) {
  double best = -1.0;
  twl = 4; thl = 3; til = 0;
  for (int a = 1; a <= 4; ++a)
    for (int b = 0; a + b <= 7; ++b) {
      const int c = 7 - a - b;
      const int TW = 1 << a, TH = 1 << b, TI = 1 << c;
      if (TI > 32) continue;
      if (TW >= 2 * ow && a > 1) continue;
      if (TH >= 2 * oh && b > 0) continue;
      const long long tiles = (long long)((ow + TW - 1) / TW) * ((oh + TH - 1) / TH) * ((n + TI - 1) / TI);
      const double useful = (double)n * oh * ow / (tiles * 128.0);
      const double halo = (double)TI * (s_eff * (TH - 1) + 3) * (s_eff * (TW - 1) + 3) / 128.0;
      const double score = useful / (1.0 + 0.15 * halo);
      if (score > best) { best = score; twl = a; thl = b; til = c; }
    }
}

inline void finish_tile_geo(TileGeo& t, int cstride_pad) {
  pick_tile(t.n, t.oh, t.ow, t.s_eff, t.tw_log, t.th_log, t.ti_log);
  const int TW = 1 << t.tw_log, TH = 1 << t.th_log, TI = 1 << t.ti_log;
  t.tiles_w = (t.ow + TW - 1) / TW; t.tiles_h = (t.oh + TH - 1) / TH; t.tiles_i = (t.n + TI - 1) / TI;
  t.tiles_per_task = t.tiles_w * t.tiles_h * t.tiles_i;
  t.halo_h = t.s_eff * (TH - 1) + 3; t.halo_w = t.s_eff * (TW - 1) + 3;
  t.cstride = (t.sc < 32 ? t.sc : 32) + cstride_pad;
}

inline int halo_pixels(const TileGeo& t) { return (1 << t.ti_log) * t.halo_h * t.halo_w; }

__device__ __forceinline__ void tile_origin(const TileGeo& t, int tile, int& i0, int& h0, int& w0) {
  const int tile_w = tile % t.tiles_w;
  const int tile_h = (tile / t.tiles_w) % t.tiles_h;
  const int tile_i = tile / (t.tiles_w * t.tiles_h);
  h0 = tile_h << t.th_log; w0 = tile_w << t.tw_log; i0 = tile_i << t.ti_log;
}

// halo offset (floats) of tile pixel px (0..127), tap (0,0)
__device__ __forceinline__ int pixel_base(const TileGeo& t, int px) {
  const int pw = px & ((1 << t.tw_log) - 1), ph = (px >> t.tw_log) & ((1 << t.th_log) - 1);
  const int ti = px >> (t.tw_log + t.th_log);
  return ((ti * t.halo_h + t.s_eff * ph) * t.halo_w + t.s_eff * pw) * t.cstride;
}

// K-row -> halo offset table for a chunk of cc channels: row k = tap*cc + cl; padded rows -> 0.
__device__ __forceinline__ void build_offtab(const TileGeo& t, int* offtab, int rows, int cc, int tid, int nthr) {
  for (int k = tid; k < rows; k += nthr) {
    int off = 0;
    if (k < 9 * cc) {
      const int tap = k / cc, cl = k - tap * cc;
      off = ((tap / 3) * t.halo_w + (tap % 3)) * t.cstride + cl;
    }
    offtab[k] = off;
  }
}

// Stages channels [c0, c0+cc) of the source halo of the tile at (i0, h0, w0) into halo[px][cstride];
// out-of-range positions (zero padding, dilation holes, images >= n) become 0.
__device__ __forceinline__ void stage_halo(const TileGeo& t, const float* __restrict__ S, int task,
                                           int i0, int h0, int w0, int c0, int cc, float* halo,
                                           int tid, int nthr) {
  const int TI = 1 << t.ti_log;
  const int vy0 = t.s_eff * h0 - 1, vx0 = t.s_eff * w0 - 1;
  const int hw = t.halo_h * t.halo_w;
  if (t.src_nchw) {
    for (int i = tid; i < TI * cc * hw; i += nthr) {
      const int ti = i / (cc * hw);
      int r = i - ti * cc * hw;
      const int cl = r / hw;
      r -= cl * hw;
      const int yy = r / t.halo_w, xx = r - yy * t.halo_w;
      const int y = vy0 + yy, x = vx0 + xx, img = i0 + ti;
      float v = 0.f;
      if (img < t.n && y >= 0 && y < t.sh && x >= 0 && x < t.sw) {
        const long long row = (long long)task * t.rows_per_task + t.row0 + (long long)img * t.row_step;
        v = __ldg(S + ((row * t.sc + c0 + cl) * t.sh + y) * t.sw + x);
      }
      halo[((ti * t.halo_h + yy) * t.halo_w + xx) * t.cstride + cl] = v;
    }
  } else if ((cc & 3) == 0 && (t.sc & 3) == 0) {
    const int c4n = cc >> 2;
    for (int i = tid; i < TI * hw * c4n; i += nthr) {
      const int px = i / c4n, c4 = i - px * c4n;
      const int xx = px % t.halo_w, yy = (px / t.halo_w) % t.halo_h, ti = px / hw;
      int y = vy0 + yy, x = vx0 + xx;
      const int img = i0 + ti;
      bool ok = img < t.n && y >= 0 && x >= 0;
      if (t.dilate == 2) { ok = ok && !((y | x) & 1); y >>= 1; x >>= 1; }
      ok = ok && y < t.sh && x < t.sw;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok) v = __ldg(reinterpret_cast<const float4*>(
                      S + ((((long long)task * t.n + img) * t.sh + y) * t.sw + x) * t.sc + c0) + c4);
      *reinterpret_cast<float4*>(halo + (size_t)px * t.cstride + c4 * 4) = v;
    }
  } else {
    for (int i = tid; i < TI * hw * cc; i += nthr) {
      const int px = i / cc, cl = i - px * cc;
      const int xx = px % t.halo_w, yy = (px / t.halo_w) % t.halo_h, ti = px / hw;
      int y = vy0 + yy, x = vx0 + xx;
      const int img = i0 + ti;
      bool ok = img < t.n && y >= 0 && x >= 0;
      if (t.dilate == 2) { ok = ok && !((y | x) & 1); y >>= 1; x >>= 1; }
      ok = ok && y < t.sh && x < t.sw;
      float v = 0.f;
      if (ok) v = __ldg(S + ((((long long)task * t.n + img) * t.sh + y) * t.sw + x) * t.sc + c0 + cl);
      halo[(size_t)px * t.cstride + cl] = v;
    }
  }
}

}  // namespace xm
