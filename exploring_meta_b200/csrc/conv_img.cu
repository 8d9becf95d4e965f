// conv_img.cu -- xm_conv for the IMAGE layer (cin <= 4, stride 1, NCHW user images): exact-fp32 direct
// convolution on the CUDA cores.
//
// Why not the tensor cores: K = 9*cin <= 36, so the layer is ~10 GFLOP per launch at config 2 against 722 MB of
// output -- it is bound by the HBM write of z (and by the latency of getting there), not by arithmetic.  FFMA2
// (packed fp32 FMA, scalar x broadcast against channel pairs) gives exact fp32 products (no operand splitting),
// the x band of the tile and the per-task weights sit in shared memory, and every output element is written
// once with 128 B-contiguous stores per pixel.
//
// Mapping: a CTA owns one task and walks over bands of CI_ROWS output rows of one image.  The input band
// (+1 halo row/column each side, zero padded) is staged as [row][col] float4 = the (<= 4) channels of a pixel.
// A warp takes 16 consecutive output pixels per iteration; lane = (pixel quad q = lane/8, channel group
// cg = lane%8 -> output channels 4cg..4cg+3 of the CTA's 32-wide slice); a thread computes 4 consecutive pixels
// (4q..4q+3) x 4 channels, so each weight LDS.128 (a broadcast within the 8 lanes of a quad) feeds 16 FMAs.
// Epilogue: NHWC float4 stores + BatchNorm statistics ({sum, sum sq} or {sum, sum v*aux}) per thread in fp32,
// then warp shuffle -> shared -> one double atomic per channel per CTA.
#include "common.cuh"

namespace xm {

constexpr int CI_THREADS = 256;
constexpr int CI_ROWS = 6;

struct ConvImgK {
  int n, H, W, cout, splits;
  int row0, row_step, rows_per_task;
  int stat_mode;
  const float* x; const float* w; long long wstride;
  float* out; const float* aux; double* stats;
};

template <int CIN>
__global__ void __launch_bounds__(CI_THREADS, 3) conv_img_kernel(const ConvImgK p) {
  extern __shared__ __align__(16) float4 sm4[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int task = blockIdx.y, split = blockIdx.x, co0 = blockIdx.z * 32;
  const int cg = lane & 7, quad = lane >> 3;
  const int Wp = p.W + 2;
  float4* wsm = sm4;                                   // [9*CIN][8] float4: weights [tap][ci][co 32]
  float4* band = sm4 + 9 * CIN * 8;                    // 2 x ([CI_ROWS+2][Wp] + 4 pad)
  __shared__ double sred[2][32];

  {
    const float* Wt = p.w + (long long)task * p.wstride;         // [cout][CIN][3][3]
    float* wf = reinterpret_cast<float*>(wsm);
    for (int i = tid; i < 9 * CIN * 32; i += CI_THREADS) {
      const int co = i & 31, r = i >> 5, ci = r % CIN, tap = r / CIN;
      wf[i] = (co0 + co < p.cout) ? __ldg(Wt + ((long long)(co0 + co) * CIN + ci) * 9 + tap) : 0.f;
    }
    if (tid < 64) sred[tid >> 5][tid & 31] = 0.0;
  }
  float ssum[4] = {0.f, 0.f, 0.f, 0.f}, ssq[4] = {0.f, 0.f, 0.f, 0.f};
  const int bands_per_img = (p.H + CI_ROWS - 1) / CI_ROWS;
  const int nbands = p.n * bands_per_img;
  const int chw = p.H * p.W;

  // x bands are double-buffered: band k+1 streams in with cp.async (4 B per channel element, zero-filled
  // outside the image) while band k is being convolved
  const int band_px = (CI_ROWS + 2) * Wp + 4;
  auto issue_band = [&](int b, float4* dst) {
    const int img = b / bands_per_img, y0 = (b - img * bands_per_img) * CI_ROWS;
    const float* X = p.x + ((long long)task * p.rows_per_task + p.row0 + (long long)img * p.row_step) * CIN * chw;
    for (int i = tid; i < (CI_ROWS + 2) * Wp; i += CI_THREADS) {
      const int yy = i / Wp, xx = i - yy * Wp;
      const int y = y0 - 1 + yy, x = xx - 1;
      const bool in = y >= 0 && y < p.H && x >= 0 && x < p.W;
      const float* src = in ? X + y * p.W + x : X;
#pragma unroll
      for (int c = 0; c < CIN; ++c)
        cp_async4(reinterpret_cast<float*>(dst + i) + c, src + (in ? (long long)c * chw : 0), in ? 4 : 0);
    }
    cp_async_commit();
  };
  for (int i = tid; i < 2 * band_px; i += CI_THREADS) band[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  if (split < nbands) issue_band(split, band);
  int kbuf = 0;
  for (int b = split; b < nbands; b += p.splits, kbuf ^= 1) {
    const int img = b / bands_per_img, y0 = (b - img * bands_per_img) * CI_ROWS;
    float4* const cur = band + kbuf * band_px;
    if (b + p.splits < nbands) {
      issue_band(b + p.splits, band + (kbuf ^ 1) * band_px);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const int rows = min(CI_ROWS, p.H - y0);
    const long long obase = (((long long)task * p.n + img) * p.H + y0) * p.W;
    const int gpr = (p.W + 3) >> 2, ngroups = rows * gpr;      // groups of 4 consecutive pixels of one row
    for (int g0 = warp * 4; g0 < ngroups; g0 += (CI_THREADS / 32) * 4) {
      // this thread's group: 4 consecutive pixels of one row (their 3x3 windows overlap, so the unrolled tap
      // loop below re-uses the x loads)
      const int gi = g0 + quad;
      const int gc = gi < ngroups ? gi : 0;
      const int ry = gc / gpr, x0 = (gc - ry * gpr) * 4;
      const int bbase = ry * Wp + x0;
      const int i_first = ry * p.W + x0;
      int boff[4];
      bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        ok[u] = gi < ngroups && x0 + u < p.W;
        boff[u] = bbase + u;            // masked pixels past the row end read (finite) neighbours or the pad
      }
      float acc[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[u][c] = 0.f;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          float xv[4][4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float4 t = cur[boff[u] + kh * Wp + kw];
            xv[u][0] = t.x; xv[u][1] = t.y; xv[u][2] = t.z; xv[u][3] = t.w;
          }
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci) {
            const float4 wv = wsm[((kh * 3 + kw) * CIN + ci) * 8 + cg];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              ffma2(acc[u][0], acc[u][1], xv[u][ci], wv.x, wv.y);
              ffma2(acc[u][2], acc[u][3], xv[u][ci], wv.z, wv.w);
            }
          }
        }
      if (co0 + 4 * cg < p.cout) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (!ok[u]) continue;
          const long long o = (obase + i_first + u) * p.cout + co0 + 4 * cg;
          *reinterpret_cast<float4*>(p.out + o) = make_float4(acc[u][0], acc[u][1], acc[u][2], acc[u][3]);
          if (p.stat_mode == XM_STAT_SUM_SQ) {
#pragma unroll
            for (int c = 0; c < 4; ++c) { ssum[c] += acc[u][c]; ssq[c] = fmaf(acc[u][c], acc[u][c], ssq[c]); }
          } else if (p.stat_mode == XM_STAT_SUM_AUX) {
            const float4 a4 = __ldg(reinterpret_cast<const float4*>(p.aux + o));
            ssum[0] += acc[u][0]; ssum[1] += acc[u][1]; ssum[2] += acc[u][2]; ssum[3] += acc[u][3];
            ssq[0] = fmaf(acc[u][0], a4.x, ssq[0]); ssq[1] = fmaf(acc[u][1], a4.y, ssq[1]);
            ssq[2] = fmaf(acc[u][2], a4.z, ssq[2]); ssq[3] = fmaf(acc[u][3], a4.w, ssq[3]);
          }
        }
      }
    }
    __syncthreads();          // every warp is done with `cur` before the next iteration refills it
  }
  if (p.stat_mode) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      double a = (double)ssum[c], b = (double)ssq[c];
      a += __shfl_xor_sync(0xffffffffu, a, 8);  b += __shfl_xor_sync(0xffffffffu, b, 8);
      a += __shfl_xor_sync(0xffffffffu, a, 16); b += __shfl_xor_sync(0xffffffffu, b, 16);
      if (quad == 0) { atomicAdd(&sred[0][4 * cg + c], a); atomicAdd(&sred[1][4 * cg + c], b); }
    }
    __syncthreads();
    if (tid < 64) {
      const int which = tid >> 5, c = tid & 31;
      if (co0 + c < p.cout)
        atomicAdd(&p.stats[((long long)task * 2 + which) * p.cout + co0 + c], sred[which][c]);
    }
  }
}

// Returns 1 if handled, 0 if the shape is not covered, otherwise an error code.
int conv_img_try(const XmConvArgs* a, cudaStream_t stream) {
  const XmBlockGeom& g = a->g;
  if (!a->src_nchw || a->mode != XM_CONV_FWD || a->src2 || g.stride != 1 || g.cin > 4 || g.cout % 4 != 0) return 0;
  ConvImgK k{};
  k.n = g.n; k.H = g.hin; k.W = g.win; k.cout = g.cout;
  k.row0 = a->row0; k.row_step = a->row_step; k.rows_per_task = a->rows_per_task;
  k.stat_mode = a->stat_mode;
  k.x = a->src1; k.w = a->w1; k.wstride = a->w1_task_stride;
  k.out = a->out; k.aux = a->aux; k.stats = a->stats;
  const int cotiles = (g.cout + 31) / 32;
  const int nbands = g.n * ((g.hin + CI_ROWS - 1) / CI_ROWS);
  const size_t smem = ((size_t)9 * g.cin * 8 + 2 * ((size_t)(CI_ROWS + 2) * (g.win + 2) + 4)) * 16;   // 2 bands (+ 4 pad pixels)
  if (smem > 64 * 1024) return 0;
  const void* kern = g.cin == 1 ? (const void*)conv_img_kernel<1> : g.cin == 2 ? (const void*)conv_img_kernel<2>
                   : g.cin == 3 ? (const void*)conv_img_kernel<3> : (const void*)conv_img_kernel<4>;
  int splits = wave_ctas(kern, CI_THREADS, smem) / (g.tasks * cotiles);      // one wave
  if (splits > nbands) splits = nbands;
  if (splits < 1) splits = 1;
  k.splits = splits;
  if (a->stat_mode) XM_CUDA(cudaMemsetAsync(a->stats, 0, (size_t)g.tasks * 2 * g.cout * sizeof(double), stream));
  dim3 grid(splits, g.tasks, cotiles);
  {   // per call: the attribute is per device, and a process may drive several
    XM_CUDA(cudaFuncSetAttribute(conv_img_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    XM_CUDA(cudaFuncSetAttribute(conv_img_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    XM_CUDA(cudaFuncSetAttribute(conv_img_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    XM_CUDA(cudaFuncSetAttribute(conv_img_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  }
  switch (g.cin) {
    case 1: conv_img_kernel<1><<<grid, CI_THREADS, smem, stream>>>(k); break;
    case 2: conv_img_kernel<2><<<grid, CI_THREADS, smem, stream>>>(k); break;
    case 3: conv_img_kernel<3><<<grid, CI_THREADS, smem, stream>>>(k); break;
    default: conv_img_kernel<4><<<grid, CI_THREADS, smem, stream>>>(k); break;
  }
  if (int rc = launched("xm_conv(image)")) return rc;
  return 1;
}

}  // namespace xm
