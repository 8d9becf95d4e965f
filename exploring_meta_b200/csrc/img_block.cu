// img_block.cu -- the IMAGE block (first ConvBlock: conv3x3(cin <= 4) -> BN(train) -> ReLU -> MaxPool 2x2) as one
// fused unit that never materialises the pre-BN map z, forward / backward and both tangent ("dual") passes.
//
// Why a dedicated path.  At config 2 the image block's z is 84x84x32 per image = 722 MB per 32-task call, and the
// generic conv -> bn -> wgrad chain moves it through HBM 16 times per inner step (6 primal + 10 dual passes): ~70 GB
// of the ~95 GB a meta-iteration touched.  Two facts remove all of it:
//  (1) the block input is the USER IMAGE x, constant over the inner loop and carrying no gradient.  With the im2col
//      matrix X[px][k] (k = (ci, kh, kw) in the weight's own [cin][3][3] order, K = 9 cin <= 36) z = X w per output
//      channel, so every DENSE reduction over z is closed-form in the per-task Gram matrix G = X^T X and column sum
//      sx = X^T 1 (xm_img_gram, once per meta-iteration, double):
//          sum z = w.sx        sum z^2 = w^T G w        sum zdot*z = wd^T G w
//          sum_px xhat X = r (G w - mean sx)            (the "every element" part of the BN backward under wgrad)
//  (2) everything else in the backward is SPARSE: the cotangent gp only reaches the arg-max element of each pooling
//      window, so the forward stores, per pooled element, the winner's z value (zsel) and its 2-bit window position
//      (sel; 255 = ReLU-dead), and the backward is one pass over pooled-resolution tensors that gathers the 3x3xcin
//      input patch of each winner:  S[co][k] = sum_sel gp * X[pos(sel)][k].
//  With m1 = <g>, m2 = <g xhat> (means over ALL n*hz*wz elements, g = gp scattered to the winners) the weight gradient is
//          dW = gamma r ( S - m1 sx - m2 XH ),   XH = r (G w - mean sx)
//  and its tangent (direction wd, gamma_dot; closed forms of SURVEY App. F pushed through wgrad) is
//          dWdot = coef (S - m1 sx - m2 XH) + gamma r ( Sd - e1 sx - m2 XHD - m2dot XH ),   XHD = r (G wd - d1 sx - d2 XH)
//  with Sd the same sparse gather for gpdot.  Traffic per inner step drops from 16 z-sized + 8 p-sized passes to
//  ~13 p-sized ones (z = 4 p); the dense arithmetic that remains is the forward conv itself (FFMA2, exact fp32).
//
// Reference ops replaced: conv2d / native_batch_norm / relu / max_pool2d_with_indices of the first ConvBlock
// (core_functions/vision_models.py:188-193), their backward and double-backward ops (vision/maml_vision.py:112) and
// the fused SGD step of learn2learn's maml_update for this block's four parameters.
#include "img_flat.cuh"

namespace xm {

constexpr int IB_THREADS = 256;
constexpr int IB_ROWS = 6;          // z rows per band of the forward (3 pooled rows)
constexpr int IB_PROWS = 3;         // pooled rows per band of the backward
constexpr int GR_ROWS = 8;          // image rows per band of the Gram kernel (= workers per pair)


__device__ __forceinline__ const float* image_ptr(const ImgK& p, int task, int img, int cin) {
  return p.x + ((long long)task * p.rows_per_task + p.row0 + (long long)img * p.row_step) * cin * p.H * p.W;
}

// Streams `rows` padded image rows starting at image row y_first (may be -1) into dst[row][Wp] float4 (= the <= 4
// channels of a pixel) with cp.async; out-of-image elements are zero-filled.  Unused float4 lanes keep their zeros.
template <int CIN>
__device__ __forceinline__ void issue_band(const ImgK& p, const float* X, int y_first, int rows, float4* dst) {
  const int Wp = p.W + 2, chw = p.H * p.W;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int yy = warp; yy < rows; yy += nwarps) {          // one band row per warp step: no per-element index arithmetic
    const int y = y_first + yy;
    const bool rowin = y >= 0 && y < p.H;
    const float* src = X + (rowin ? y : 0) * p.W - 1;
    float4* d = dst + yy * Wp;
    for (int xx = lane; xx < Wp; xx += 32) {
      const bool in = rowin && xx >= 1 && xx <= p.W;
#pragma unroll
      for (int c = 0; c < CIN; ++c)
        cp_async4(reinterpret_cast<float*>(d + xx) + c, in ? src + xx + (long long)c * chw : X, in ? 4 : 0);
    }
  }
  cp_async_commit();
}

// -------------------------------------------------------------------------------------------------
// Gram matrix.  Row groups a = (ci, kh) (3 cin of them); a thread owns the 3x3 block (kw, kw') of a pair a <= b and
// slides along an image row: per pixel 2 new shared loads feed 9 DFMAs.  Products of two floats are exact in double.
template <int CIN>
__global__ void __launch_bounds__((3 * CIN) * (3 * CIN + 1) / 2 * GR_ROWS) img_gram_kernel(const ImgK p) {
  constexpr int RG = 3 * CIN, NPAIR = RG * (RG + 1) / 2, K = 9 * CIN;
  extern __shared__ __align__(16) double gsm[];
  const int Wp = p.W + 2;
  double* band = gsm;                                   // [CIN][GR_ROWS + 2][Wp]
  double* red = gsm + CIN * (GR_ROWS + 2) * Wp;         // [NPAIR][9] + [RG][3]
  const int tid = threadIdx.x, task = blockIdx.y, split = blockIdx.x;
  const int pair = tid % NPAIR, worker = tid / NPAIR;
  int a = 0, b = 0;
  {
    int rem = pair;
    for (a = 0; a < RG; ++a) { if (rem < RG - a) { b = a + rem; break; } rem -= RG - a; }
  }
  const int cia = a / 3, kha = a % 3, cib = b / 3, khb = b % 3;
  for (int i = tid; i < NPAIR * 9 + RG * 3; i += blockDim.x) red[i] = 0.0;
  double acc[3][3], sacc[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) acc[i][j] = 0.0;
  const int bands_per_img = (p.H + GR_ROWS - 1) / GR_ROWS, nbands = p.n * bands_per_img, chw = p.H * p.W;
  for (int bd = split; bd < nbands; bd += p.splits) {
    const int img = bd / bands_per_img, y0 = (bd - img * bands_per_img) * GR_ROWS;
    const float* X = image_ptr(p, task, img, CIN);
    __syncthreads();
    for (int i = tid; i < CIN * (GR_ROWS + 2) * Wp; i += blockDim.x) {
      const int c = i / ((GR_ROWS + 2) * Wp), r = i - c * (GR_ROWS + 2) * Wp;
      const int yy = r / Wp, xx = r - yy * Wp, y = y0 - 1 + yy, x = xx - 1;
      band[i] = (y >= 0 && y < p.H && x >= 0 && x < p.W) ? (double)__ldg(X + (long long)c * chw + y * p.W + x) : 0.0;
    }
    __syncthreads();
    if (y0 + worker < p.H) {
      const double* ra = band + (cia * (GR_ROWS + 2) + worker + kha) * Wp;
      const double* rb = band + (cib * (GR_ROWS + 2) + worker + khb) * Wp;
      double a0 = ra[0], a1 = ra[1], b0 = rb[0], b1 = rb[1];
      for (int x = 0; x < p.W; ++x) {
        const double a2 = ra[x + 2], b2 = rb[x + 2];
        acc[0][0] = fma(a0, b0, acc[0][0]); acc[0][1] = fma(a0, b1, acc[0][1]); acc[0][2] = fma(a0, b2, acc[0][2]);
        acc[1][0] = fma(a1, b0, acc[1][0]); acc[1][1] = fma(a1, b1, acc[1][1]); acc[1][2] = fma(a1, b2, acc[1][2]);
        acc[2][0] = fma(a2, b0, acc[2][0]); acc[2][1] = fma(a2, b1, acc[2][1]); acc[2][2] = fma(a2, b2, acc[2][2]);
        if (a == b) { sacc[0] += a0; sacc[1] += a1; sacc[2] += a2; }
        a0 = a1; a1 = a2; b0 = b1; b1 = b2;
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) atomicAdd(&red[pair * 9 + i * 3 + j], acc[i][j]);
  if (a == b)
#pragma unroll
    for (int i = 0; i < 3; ++i) atomicAdd(&red[NPAIR * 9 + a * 3 + i], sacc[i]);
  __syncthreads();
  double* G = p.gram + (long long)task * (K * K + K);
  if (worker == 0) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const double v = red[pair * 9 + i * 3 + j];
        const int k = a * 3 + i, k2 = b * 3 + j;
        atomicAdd(&G[k * K + k2], v);
        if (a != b) atomicAdd(&G[k2 * K + k], v);
      }
    if (a == b)
#pragma unroll
      for (int i = 0; i < 3; ++i) atomicAdd(&G[K * K + a * 3 + i], red[NPAIR * 9 + a * 3 + i]);
  }
}

// -------------------------------------------------------------------------------------------------
// Forward (DUAL = 0): conv -> BN (statistics from the Gram matrix) -> ReLU -> pool; writes p, zsel, sel.
// Tangent forward (DUAL = 1): zdot = conv(x, wd) at the stored winners; writes pdot, zdsel, dual_red.
// Mapping: lane = (quad, cg): a thread computes 2 rows x 4 columns (two pooling windows) x 4 channels.
template <int CIN, int DUAL>
__global__ void __launch_bounds__(IB_THREADS, 2) img_fwd_kernel(const ImgK p) {
  constexpr int K = 9 * CIN;
  extern __shared__ __align__(16) float4 sm4[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int task = blockIdx.y, split = blockIdx.x, co0 = blockIdx.z * 32;
  const int cg = lane & 7, quad = lane >> 3;
  const int Wp = p.W + 2;
  const int band_px = (IB_ROWS + 2) * Wp + 8;
  float4* wsm = sm4;                                   // [K][8] float4: conv weights [tap][ci][co] (w, or wd when DUAL)
  float4* band = sm4 + K * 8;                          // 2 x band_px
  float* wpr = reinterpret_cast<float*>(band + 2 * band_px);                 // [K][32] primal w, gram order (prologue)
  float* wdr = wpr + K * 32;                                                  // [K][32] wd, gram order (DUAL prologue)
  double* Gs = reinterpret_cast<double*>(wdr + K * 32);                       // [K*K + K]
  double* red = Gs + K * K + K;                                               // [8][4][32]

  {
    const float* Wt = p.w + (long long)task * p.wstride;          // [cout][CIN][3][3]
    const float* Wd = DUAL ? p.wd + (long long)task * p.wdstride : nullptr;
    float* wf = reinterpret_cast<float*>(wsm);
    for (int i = tid; i < K * 32; i += IB_THREADS) {
      const int co = i & 31, k = i >> 5;                           // k = ci*9 + tap
      const int ci = k / 9, tap = k - ci * 9;
      const float wv = __ldg(Wt + (long long)(co0 + co) * K + k);
      wpr[i] = wv;
      float cv = wv;
      if (DUAL) { cv = __ldg(Wd + (long long)(co0 + co) * K + k); wdr[i] = cv; }
      wf[(tap * CIN + ci) * 32 + co] = cv;
    }
    const double* G = p.gram + (long long)task * (K * K + K);
    for (int i = tid; i < K * K + K; i += IB_THREADS) Gs[i] = G[i];
    for (int i = tid; i < 2 * band_px; i += IB_THREADS) band[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  const int bands_per_img = (p.H + IB_ROWS - 1) / IB_ROWS, nbands = p.n * bands_per_img;
  if (split < nbands) {
    const int img = split / bands_per_img, y0 = (split - img * bands_per_img) * IB_ROWS;
    issue_band<CIN>(p, image_ptr(p, task, img, CIN), y0 - 1, IB_ROWS + 2, band);
  }
  // ---- per-channel scalars from the Gram matrix (identical in every CTA of the task) ----
  {
    const int co = tid & 31, part = tid >> 5;
    double s_ws = 0.0, s_wgw = 0.0, s_ds = 0.0, s_dgw = 0.0;
    for (int k = part; k < K; k += 8) {
      double gw = 0.0;
      for (int k2 = 0; k2 < K; ++k2) gw = fma(Gs[k * K + k2], (double)wpr[k2 * 32 + co], gw);
      const double wk = (double)wpr[k * 32 + co];
      s_ws = fma(wk, Gs[K * K + k], s_ws);
      s_wgw = fma(wk, gw, s_wgw);
      if (DUAL) {
        const double dk = (double)wdr[k * 32 + co];
        s_ds = fma(dk, Gs[K * K + k], s_ds);
        s_dgw = fma(dk, gw, s_dgw);
      }
    }
    red[(part * 4 + 0) * 32 + co] = s_ws;
    red[(part * 4 + 1) * 32 + co] = s_wgw;
    red[(part * 4 + 2) * 32 + co] = s_ds;
    red[(part * 4 + 3) * 32 + co] = s_dgw;
  }
  __syncthreads();
  // per-channel scalars live in shared memory (read back as float4 in the epilogue): [8][32] =
  // gamma, beta, mean, invstd, gamma_dot, beta_dot, d1, d2
  float* chs = reinterpret_cast<float*>(red + 8 * 4 * 32);
  if (tid < 32) {
    const int c = tid;
    double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
    for (int part = 0; part < 8; ++part) {
      t0 += red[(part * 4 + 0) * 32 + c]; t1 += red[(part * 4 + 1) * 32 + c];
      t2 += red[(part * 4 + 2) * 32 + c]; t3 += red[(part * 4 + 3) * 32 + c];
    }
    const long long mi = ((long long)task * 2) * p.cout + co0 + c;
    const bool writer = split == 0;
    chs[0 * 32 + c] = __ldg(p.gamma + (long long)task * p.gbstride + co0 + c);
    if (!DUAL) {
      chs[1 * 32 + c] = __ldg(p.beta + (long long)task * p.gbstride + co0 + c);
      const double m = t0 / p.cnt;
      double var = t1 / p.cnt - m * m;
      if (var < 0.0) var = 0.0;
      const float mf = (float)m, rf = (float)(1.0 / sqrt(var + (double)p.eps));
      chs[2 * 32 + c] = mf;
      chs[3 * 32 + c] = rf;
      if (writer) {
        p.mean_invstd[mi] = mf;
        p.mean_invstd[mi + p.cout] = rf;
        if (p.call_stats) {
          p.call_stats[mi] = mf;
          p.call_stats[mi + p.cout] = (float)(var * (p.cnt / fmax(p.cnt - 1.0, 1.0)));
        }
      }
    } else {
      const float mf = __ldg(p.mean_invstd + mi), rf = __ldg(p.mean_invstd + mi + p.cout);
      chs[2 * 32 + c] = mf;
      chs[3 * 32 + c] = rf;
      chs[4 * 32 + c] = __ldg(p.gammad + (long long)task * p.gbdstride + co0 + c);
      chs[5 * 32 + c] = __ldg(p.betad + (long long)task * p.gbdstride + co0 + c);
      const double e1 = t2 / p.cnt;
      const double e2 = (double)rf * (t3 / p.cnt - (double)mf * e1);
      chs[6 * 32 + c] = (float)e1;
      chs[7 * 32 + c] = (float)e2;
      if (writer) { p.dual_red[mi] = (float)e1; p.dual_red[mi + p.cout] = (float)e2; }
    }
  }
  __syncthreads();

  int kbuf = 0;
  for (int b = split; b < nbands; b += p.splits, kbuf ^= 1) {
    const int img = b / bands_per_img, y0 = (b - img * bands_per_img) * IB_ROWS;
    float4* const cur = band + kbuf * band_px;
    if (b + p.splits < nbands) {
      const int nb = b + p.splits, nimg = nb / bands_per_img, ny0 = (nb - nimg * bands_per_img) * IB_ROWS;
      issue_band<CIN>(p, image_ptr(p, task, nimg, CIN), ny0 - 1, IB_ROWS + 2, band + (kbuf ^ 1) * band_px);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const int rows = min(IB_ROWS, p.H - y0);                     // even (H is even)
    const int gpr = (p.W + 3) >> 2, ngroups = (rows >> 1) * gpr;  // groups of 2 rows x 4 columns
#pragma unroll 1
    for (int g0 = warp * 4; g0 < ngroups; g0 += (IB_THREADS / 32) * 4) {
      const int gi = g0 + quad;
      const int gc = gi < ngroups ? gi : 0;
      const int rp = gc / gpr, x0 = (gc - rp * gpr) * 4;
      float acc[2][4][4];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[r][u][c] = 0.f;
#pragma unroll 1
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          float xv[6][4];
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            const float4 t = cur[(2 * rp + r + kh) * Wp + x0 + j];       // columns past the row end hit the 8-pixel pad
            xv[j][0] = t.x; xv[j][1] = t.y; xv[j][2] = t.z; xv[j][3] = t.w;
          }
#pragma unroll
          for (int kw = 0; kw < 3; ++kw)
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci) {
              const float4 wv = wsm[((kh * 3 + kw) * CIN + ci) * 8 + cg];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                ffma2(acc[r][u][0], acc[r][u][1], xv[u + kw][ci], wv.x, wv.y);
                ffma2(acc[r][u][2], acc[r][u][3], xv[u + kw][ci], wv.z, wv.w);
              }
            }
        }
      const int py = (y0 >> 1) + rp;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int px = (x0 >> 1) + u;
        if (gi >= ngroups || px >= p.wp) continue;
        const long long o = ((((long long)task * p.n + img) * p.hp + py) * p.wp + px) * p.cout + co0 + 4 * cg;
        if (!DUAL) {
          float pv[4], zs[4];
          unsigned int sb = 0;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float gam_c = chs[0 * 32 + 4 * cg + c], bet_c = chs[1 * 32 + 4 * cg + c];
            const float mean_c = chs[2 * 32 + 4 * cg + c], rinv_c = chs[3 * 32 + 4 * cg + c];
            int best = 0;
            float zb = acc[0][2 * u][c];
            float yb = fmaf(gam_c, (zb - mean_c) * rinv_c, bet_c);
#pragma unroll
            for (int d = 1; d < 4; ++d) {
              const float z = acc[d >> 1][2 * u + (d & 1)][c];
              const float y = fmaf(gam_c, (z - mean_c) * rinv_c, bet_c);
              if (y > yb) { yb = y; zb = z; best = d; }
            }
            const bool on = yb > 0.f;
            pv[c] = on ? yb : 0.f;
            zs[c] = on ? zb : 0.f;
            sb |= (on ? (unsigned)best : 255u) << (8 * c);
          }
          *reinterpret_cast<float4*>(p.p + o) = make_float4(pv[0], pv[1], pv[2], pv[3]);
          *reinterpret_cast<float4*>(p.zsel + o) = make_float4(zs[0], zs[1], zs[2], zs[3]);
          *reinterpret_cast<unsigned int*>(p.sel + o) = sb;
        } else {
          const float4 z4 = __ldg(reinterpret_cast<const float4*>(p.zsel + o));
          const unsigned int sb = __ldg(reinterpret_cast<const unsigned int*>(p.sel + o));
          const float zs[4] = {z4.x, z4.y, z4.z, z4.w};
          float pd[4], zd[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float gam_c = chs[0 * 32 + 4 * cg + c], mean_c = chs[2 * 32 + 4 * cg + c], rinv_c = chs[3 * 32 + 4 * cg + c];
            const float gd_c = chs[4 * 32 + 4 * cg + c], bd_c = chs[5 * 32 + 4 * cg + c];
            const float d1_c = chs[6 * 32 + 4 * cg + c], d2_c = chs[7 * 32 + 4 * cg + c];
            const unsigned s = (sb >> (8 * c)) & 255u;
            float zds = acc[0][2 * u][c];
#pragma unroll
            for (int d = 1; d < 4; ++d) if (s == (unsigned)d) zds = acc[d >> 1][2 * u + (d & 1)][c];
            const float xhat = (zs[c] - mean_c) * rinv_c;
            const float xhd = rinv_c * (zds - d1_c - xhat * d2_c);
            const bool on = s != 255u;
            pd[c] = on ? gd_c * xhat + gam_c * xhd + bd_c : 0.f;
            zd[c] = on ? zds : 0.f;
          }
          *reinterpret_cast<float4*>(p.pdot + o) = make_float4(pd[0], pd[1], pd[2], pd[3]);
          *reinterpret_cast<float4*>(p.zdsel + o) = make_float4(zd[0], zd[1], zd[2], zd[3]);
        }
      }
    }
    __syncthreads();
  }
}

// -------------------------------------------------------------------------------------------------
// Backward gather pass.  lane = output channel; a warp walks pooled pixels of the band.  Per (pooled element,
// channel): one coalesced read of gp / zsel / sel (+ gpdot / zdsel when DUAL) and a 3x3xcin patch gather from the
// staged input band at the winner's position.  Accumulates, per (task, channel), into scratch[task][cout][K + 3]:
//   DUAL = 0: S[k] = sum gp X[sel][k],    s1 = sum gp,    s2 = sum gp xhat
//   DUAL = 1: Sd[k] = sum gpd X[sel][k],  E1 = sum gpd,   E2 = sum gpd xhat,   E3 = sum gp zdsel
template <int CIN, int DUAL>
__global__ void __launch_bounds__(IB_THREADS, 2) img_bwd_kernel(const ImgK p) {
  constexpr int K = 9 * CIN, NA = K + 3;
  extern __shared__ __align__(16) float bsm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int task = blockIdx.y, split = blockIdx.x, co0 = blockIdx.z * 32;
  const int Wp = p.W + 2, rows = 2 * IB_PROWS + 2;
  const int plane = rows * Wp, band_fl = CIN * plane + 8;
  // The image band is staged PLANAR ([cin][row][col] floats): the 32 lanes of a patch load then touch at most four
  // addresses (the four window positions) in four different banks -- one shared-memory wavefront per LDS.32, 27 per
  // pooled pixel, where 16-byte pixel vectors cost 36 (a 128-bit load is served a quarter warp at a time).
  float* band = bsm;                                                       // 2 x band_fl
  double* red = reinterpret_cast<double*>(bsm + ((2 * band_fl + 1) & ~1)); // [warps][NA][32]: a thread owns its slots
  for (int i = tid; i < 2 * band_fl; i += IB_THREADS) band[i] = 0.f;
  for (int i = tid; i < (IB_THREADS / 32) * NA * 32; i += IB_THREADS) red[i] = 0.0;
  double* const mine = red + (warp * NA) * 32 + lane;
  __syncthreads();
  const int bands_per_img = (p.hp + IB_PROWS - 1) / IB_PROWS, nbands = p.n * bands_per_img;
  const int chw = p.H * p.W;
  auto issue = [&](int b, float* dst) {
    const int img = b / bands_per_img, py0 = (b - img * bands_per_img) * IB_PROWS;
    const float* X = image_ptr(p, task, img, CIN);
    // one (channel, band row) per warp step, lanes over the columns: no per-element index arithmetic
    for (int cr = warp; cr < CIN * rows; cr += IB_THREADS / 32) {
      const int c = cr / rows, yy = cr - c * rows, y = 2 * py0 - 1 + yy;
      const bool rowin = y >= 0 && y < p.H;
      const float* src = X + (long long)c * chw + (rowin ? y : 0) * p.W - 1;
      float* d = dst + cr * Wp;
      for (int xx = lane; xx < Wp; xx += 32) {
        const bool in = rowin && xx >= 1 && xx <= p.W;
        cp_async4(d + xx, in ? src + xx : X, in ? 4 : 0);
      }
    }
    cp_async_commit();
  };
  if (split < nbands) issue(split, band);
  const long long mi = ((long long)task * 2) * p.cout + co0 + lane;
  const float mean = __ldg(p.mean_invstd + mi), rinv = __ldg(p.mean_invstd + mi + p.cout);
  float facc[NA];                 // fp32 over four bands (~64 winners per thread), then into the thread's double slots
#pragma unroll
  for (int i = 0; i < NA; ++i) facc[i] = 0.f;

  int kbuf = 0, nb_done = 0;
  for (int b = split; b < nbands; b += p.splits, kbuf ^= 1) {
    const int img = b / bands_per_img, py0 = (b - img * bands_per_img) * IB_PROWS;
    const float* const cur = band + kbuf * band_fl;
    if (b + p.splits < nbands) {
      issue(b + p.splits, band + (kbuf ^ 1) * band_fl);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const int prows = min(IB_PROWS, p.hp - py0), npx = prows * p.wp;
    const long long obase = ((((long long)task * p.n + img) * p.hp + py0) * p.wp) * p.cout + co0 + lane;
    // pooled pixel j <-> (row jy[e], column jx[e]) of the band, advanced incrementally (16 pixels per step)
    int jy[2], jx[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int j = warp + e * (IB_THREADS / 32);
      jy[e] = j / p.wp;
      jx[e] = j - jy[e] * p.wp;
    }
    for (int j0 = warp; j0 < npx; j0 += 2 * (IB_THREADS / 32)) {
      float g[2], gd[2], zs[2], zd[2];
      unsigned s[2];
      bool live[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = j0 + e * (IB_THREADS / 32);
        live[e] = j < npx;
        const float* q = p.gp + obase + (long long)(live[e] ? j : j0) * p.cout;
        const long long o = q - p.gp;
        g[e] = __ldg(q);
        zs[e] = __ldg(p.zsel + o);
        s[e] = __ldg(p.sel + o);
        if (DUAL) { gd[e] = p.gpd ? __ldg(p.gpd + o) : 0.f; zd[e] = __ldg(p.zdsel + o); }
        else { gd[e] = 0.f; zd[e] = 0.f; }
      }
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const bool on = live[e] && s[e] != 255u;
        const int pyl = live[e] ? jy[e] : 0, px = live[e] ? jx[e] : 0;
        const int dy = on ? (int)(s[e] >> 1) : 0, dx = on ? (int)(s[e] & 1u) : 0;
        const float gg = on ? g[e] : 0.f, ggd = on ? gd[e] : 0.f;
        const float xhat = (zs[e] - mean) * rinv;
        const float c = DUAL ? ggd : gg;
        facc[K] += c;
        facc[K + 1] = fmaf(c, xhat, facc[K + 1]);
        if (DUAL) facc[K + 2] = fmaf(gg, zd[e], facc[K + 2]);
        const float* src = cur + (2 * pyl + dy) * Wp + 2 * px + dx;
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
          for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw)
              facc[ci * 9 + kh * 3 + kw] = fmaf(c, src[ci * plane + kh * Wp + kw], facc[ci * 9 + kh * 3 + kw]);
        jx[e] += 2 * (IB_THREADS / 32);
        while (jx[e] >= p.wp) { jx[e] -= p.wp; ++jy[e]; }
      }
    }
    if ((++nb_done & 3) == 0) {
#pragma unroll
      for (int i = 0; i < NA; ++i) { mine[i * 32] += (double)facc[i]; facc[i] = 0.f; }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < NA; ++i) mine[i * 32] += (double)facc[i];
  __syncthreads();
  double* out = p.scratch + ((long long)task * p.cout + co0) * NA;
  for (int i = tid; i < NA * 32; i += IB_THREADS) {
    const int co = i / NA, q = i - co * NA;
    double v = 0.0;
#pragma unroll
    for (int wv = 0; wv < IB_THREADS / 32; ++wv) v += red[(wv * NA + q) * 32 + co];
    atomicAdd(&out[co * NA + q], v);
  }
}

// -------------------------------------------------------------------------------------------------
// Closed-form tails, one CTA per task, all in double.  G, sx and the weights are staged in shared memory once; the
// matrix-vector products G w (and G wd) are spread over (channel, k) pairs, then one thread per channel finishes.
template <int CIN, int DUAL>
__global__ void __launch_bounds__(256) img_finalize_kernel(const ImgK p) {
  constexpr int K = 9 * CIN, NA = K + 3;
  extern __shared__ __align__(16) double fsm[];
  const int task = blockIdx.x, tid = threadIdx.x, C = p.cout;
  double* Gs = fsm;                         // [K*K + K]
  double* ws = Gs + K * K + K;              // [C][K]
  double* wds = ws + C * K;                 // [C][K]   (DUAL)
  double* gw = wds + (DUAL ? C * K : 0);    // [C][K]
  double* gwd = gw + C * K;                 // [C][K]   (DUAL)
  const double* G = p.gram + (long long)task * (K * K + K);
  for (int i = tid; i < K * K + K; i += blockDim.x) Gs[i] = G[i];
  const float* W = p.w + (long long)task * p.wstride;
  const float* Wd = DUAL ? p.wd + (long long)task * p.wdstride : nullptr;
  for (int i = tid; i < C * K; i += blockDim.x) {
    ws[i] = (double)W[i];
    if (DUAL) wds[i] = (double)Wd[i];
  }
  __syncthreads();
  for (int i = tid; i < C * K; i += blockDim.x) {
    const int co = i / K, k = i - co * K;
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int k2 = 0; k2 < K; ++k2) {
      const double gv = Gs[k * K + k2];
      a = fma(gv, ws[co * K + k2], a);
      if (DUAL) b = fma(gv, wds[co * K + k2], b);
    }
    gw[i] = a;
    if (DUAL) gwd[i] = b;
  }
  __syncthreads();
  const double* sx = Gs + K * K;
  for (int i = tid; i < C * K; i += blockDim.x) {
    const int co = i / K, k = i - co * K;
    const double* A = p.scratch + ((long long)task * C + co) * NA;
    const long long mi = ((long long)task * 2) * C + co;
    const double mean = (double)p.mean_invstd[mi], r = (double)p.mean_invstd[mi + C];
    const double gamma = (double)p.gamma[(long long)task * p.gbstride + co];
    const double xh = r * (gw[i] - mean * sx[k]);
    double grad;
    if (!DUAL) {
      const double m1 = A[K] / p.cnt, m2 = A[K + 1] / p.cnt;
      grad = gamma * r * (A[k] - m1 * sx[k] - m2 * xh);
    } else {
      const double* S = p.ssum + ((long long)task * C + co) * NA;        // sparse sums of the primal backward
      const double gdot = (double)p.gammad[(long long)task * p.gbdstride + co];
      const double m1 = (double)p.bwd_red[mi], m2 = (double)p.bwd_red[mi + C];
      const double d1 = (double)p.dual_red[mi], d2 = (double)p.dual_red[mi + C];
      const double e1 = A[K] / p.cnt, e2 = A[K + 1] / p.cnt, e3 = A[K + 2] / p.cnt;
      const double q = r * (e3 - d1 * m1 - d2 * m2);          // <g * xhat_dot>
      const double m2dot = e2 + q;
      const double coef = gdot * r - gamma * r * r * d2, gr = gamma * r;
      const double xhd = r * (gwd[i] - d1 * sx[k] - d2 * xh);
      grad = coef * (S[k] - m1 * sx[k] - m2 * xh) + gr * (A[k] - e1 * sx[k] - m2 * xhd - m2dot * xh);
    }
    if (p.out_w) {
      const float base = p.base_w ? p.base_w[(long long)task * p.bstride + i] : 0.f;
      p.out_w[(long long)task * p.ostride + i] = base + p.scale * (float)grad;
    }
  }
  for (int co = tid; co < C; co += blockDim.x) {
    const double* A = p.scratch + ((long long)task * C + co) * NA;
    const long long mi = ((long long)task * 2) * C + co;
    double ggamma, gbeta = A[K];
    if (!DUAL) {
      ggamma = A[K + 1];
      if (p.bwd_red) { p.bwd_red[mi] = (float)(A[K] / p.cnt); p.bwd_red[mi + C] = (float)(A[K + 1] / p.cnt); }
      if (p.ssum) {
        double* dst = p.ssum + ((long long)task * C + co) * NA;
        for (int i = 0; i < NA; ++i) dst[i] = A[i];
      }
    } else {
      const double r = (double)p.mean_invstd[mi + C];
      const double m1 = (double)p.bwd_red[mi], m2 = (double)p.bwd_red[mi + C];
      const double d1 = (double)p.dual_red[mi], d2 = (double)p.dual_red[mi + C];
      const double q = r * (A[K + 2] / p.cnt - d1 * m1 - d2 * m2);
      ggamma = A[K + 1] + q * p.cnt;
    }
    if (p.out_gamma) {
      const float bg = p.base_gamma ? p.base_gamma[(long long)task * p.bstride + co] : 0.f;
      const float bb = p.base_beta ? p.base_beta[(long long)task * p.bstride + co] : 0.f;
      p.out_gamma[(long long)task * p.ostride + co] = bg + p.scale * (float)ggamma;
      p.out_beta[(long long)task * p.ostride + co] = bb + p.scale * (float)gbeta;
    }
    if (p.out_b) p.out_b[(long long)task * p.ostride + co] = p.base_b ? p.base_b[(long long)task * p.bstride + co] : 0.f;
  }
}

// ---- host side -----------------------------------------------------------------------------------
static int pooled_ok(const XmBlockGeom& g) {
  return geom_ok(g) && g.cin >= 1 && g.cin <= 4 && g.stride == 1 && g.pool == 1 && g.hz % 2 == 0 && g.wz % 2 == 0 &&
         g.cout % 32 == 0 && g.cout <= 1024;
}
static int img_ok(const XmBlockGeom& g) { return pooled_ok(g) || flat_ok(g); }

static void fill(const XmImgArgs* a, ImgK& k) {
  const XmBlockGeom& g = a->g;
  k = ImgK{};
  k.n = g.n; k.H = g.hin; k.W = g.win; k.cout = g.cout; k.hp = g.hp; k.wp = g.wp;
  k.row0 = a->row0; k.row_step = a->row_step; k.rows_per_task = a->rows_per_task;
  k.cnt = (double)g.n * g.hz * g.wz; k.eps = a->eps; k.scale = a->scale;
  k.x = a->x; k.gram = a->gram;
  k.w = a->w; k.wstride = a->w_task_stride; k.wd = a->w_dot; k.wdstride = a->wdot_task_stride;
  k.gamma = a->gamma; k.beta = a->beta; k.gbstride = a->gb_task_stride;
  k.gammad = a->gamma_dot; k.betad = a->beta_dot; k.gbdstride = a->gbdot_task_stride;
  k.mean_invstd = a->mean_invstd; k.call_stats = a->call_stats; k.bwd_red = a->bwd_red; k.dual_red = a->dual_red;
  k.p = a->p; k.zsel = a->zsel; k.sel = a->sel; k.pdot = a->pdot; k.zdsel = a->zdsel;
  k.gp = a->gp; k.gpd = a->gpdot; k.ssum = a->ssum; k.scratch = a->scratch;
  k.out_w = a->out_w; k.out_b = a->out_b; k.out_gamma = a->out_gamma; k.out_beta = a->out_beta; k.ostride = a->out_task_stride;
  k.base_w = a->base_w; k.base_b = a->base_b; k.base_gamma = a->base_gamma; k.base_beta = a->base_beta;
  k.bstride = a->base_task_stride;
}

static int common_checks(const XmImgArgs* a, const char* who) {
  XM_REQUIRE(a != nullptr, "%s: null args", who);
  XM_REQUIRE(img_ok(a->g), "%s: geometry not covered by the image-block path (need cin <= 4, stride 1, 2x2 pool, even "
             "hz/wz, cout %% 32 == 0 -- or cin 1, stride 2, no pool, cout 32 / 64); use xm_conv / xm_bn_* / xm_wgrad", who);
  XM_REQUIRE(a->x && a->gram, "%s: null x/gram", who);
  XM_REQUIRE(a->row_step >= 1 && a->row0 >= 0 && a->row0 + (int64_t)(a->g.n - 1) * a->row_step < a->rows_per_task,
             "%s: image row selection out of range", who);
  return 0;
}

#define IMG_DISPATCH(CINV, CALL)                                   \
  switch (CINV) {                                                  \
    case 1: { constexpr int CIN = 1; CALL; } break;                \
    case 2: { constexpr int CIN = 2; CALL; } break;                \
    case 3: { constexpr int CIN = 3; CALL; } break;                \
    default: { constexpr int CIN = 4; CALL; } break;               \
  }

template <int CIN>
static int launch_gram(const XmImgArgs* a, ImgK& k, cudaStream_t stream) {
  constexpr int RG = 3 * CIN, NPAIR = RG * (RG + 1) / 2, K = 9 * CIN, threads = NPAIR * GR_ROWS;
  const XmBlockGeom& g = a->g;
  const size_t smem = ((size_t)CIN * (GR_ROWS + 2) * (g.win + 2) + NPAIR * 9 + RG * 3) * sizeof(double);
  XM_REQUIRE(smem <= 200 * 1024, "xm_img_gram: image too wide");
  XM_CUDA(cudaFuncSetAttribute(img_gram_kernel<CIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int nbands = g.n * ((g.hin + GR_ROWS - 1) / GR_ROWS);
  int splits = wave_ctas((const void*)img_gram_kernel<CIN>, threads, smem) / g.tasks;
  if (splits > nbands) splits = nbands;
  if (splits < 1) splits = 1;
  k.splits = splits;
  XM_CUDA(cudaMemsetAsync(a->gram, 0, (size_t)g.tasks * (K * K + K) * sizeof(double), stream));
  img_gram_kernel<CIN><<<dim3(splits, g.tasks), threads, smem, stream>>>(k);
  return launched("xm_img_gram");
}

template <int CIN, int DUAL>
static int launch_fwd(const XmImgArgs* a, ImgK& k, cudaStream_t stream) {
  constexpr int K = 9 * CIN;
  const XmBlockGeom& g = a->g;
  const size_t smem = ((size_t)K * 8 + 2 * ((size_t)(IB_ROWS + 2) * (g.win + 2) + 8)) * 16 + (size_t)2 * K * 32 * 4 +
                      ((size_t)K * K + K + 8 * 4 * 32) * 8 + 8 * 32 * 4;
  XM_REQUIRE(smem <= 200 * 1024, "xm_img_fwd: image too wide");
  XM_CUDA(cudaFuncSetAttribute(img_fwd_kernel<CIN, DUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int cotiles = g.cout / 32, nbands = g.n * ((g.hin + IB_ROWS - 1) / IB_ROWS);
  int splits = wave_ctas((const void*)img_fwd_kernel<CIN, DUAL>, IB_THREADS, smem) / (g.tasks * cotiles);
  if (splits > nbands) splits = nbands;
  if (splits < 1) splits = 1;
  k.splits = splits;
  img_fwd_kernel<CIN, DUAL><<<dim3(splits, g.tasks, cotiles), IB_THREADS, smem, stream>>>(k);
  return launched(DUAL ? "xm_img_dual_fwd" : "xm_img_fwd");
}

template <int CIN, int DUAL>
static int launch_finalize(const XmImgArgs* a, ImgK& k, cudaStream_t stream);

template <int CIN, int DUAL>
static int launch_bwd(const XmImgArgs* a, ImgK& k, cudaStream_t stream) {
  constexpr int K = 9 * CIN, NA = K + 3;
  const XmBlockGeom& g = a->g;
  const size_t band_fl = (size_t)CIN * (2 * IB_PROWS + 2) * (g.win + 2) + 8;
  const size_t smem = ((2 * band_fl + 1) & ~(size_t)1) * 4 + (size_t)(IB_THREADS / 32) * NA * 32 * 8;
  XM_REQUIRE(smem <= 200 * 1024, "xm_img_bwd: image too wide");
  XM_CUDA(cudaFuncSetAttribute(img_bwd_kernel<CIN, DUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int cotiles = g.cout / 32, nbands = g.n * ((g.hp + IB_PROWS - 1) / IB_PROWS);
  int splits = wave_ctas((const void*)img_bwd_kernel<CIN, DUAL>, IB_THREADS, smem) / (g.tasks * cotiles);
  if (splits > nbands) splits = nbands;
  if (splits < 1) splits = 1;
  k.splits = splits;
  XM_CUDA(cudaMemsetAsync(a->scratch, 0, (size_t)g.tasks * g.cout * NA * sizeof(double), stream));
  img_bwd_kernel<CIN, DUAL><<<dim3(splits, g.tasks, cotiles), IB_THREADS, smem, stream>>>(k);
  if (int rc = launched(DUAL ? "xm_img_dual_bwd(gather)" : "xm_img_bwd(gather)")) return rc;
  return launch_finalize<CIN, DUAL>(a, k, stream);
}

template <int CIN, int DUAL>
static int launch_finalize(const XmImgArgs* a, ImgK& k, cudaStream_t stream) {
  constexpr int K = 9 * CIN;
  const XmBlockGeom& g = a->g;
  const size_t fsmem = ((size_t)K * K + K + (size_t)(DUAL ? 4 : 2) * g.cout * K) * sizeof(double);
  XM_REQUIRE(fsmem <= 200 * 1024, "xm_img_bwd: too many channels for the finalize kernel");
  XM_CUDA(cudaFuncSetAttribute(img_finalize_kernel<CIN, DUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  img_finalize_kernel<CIN, DUAL><<<g.tasks, 256, fsmem, stream>>>(k);
  return launched(DUAL ? "xm_img_dual_bwd(finalize)" : "xm_img_bwd(finalize)");
}

}  // namespace xm

using namespace xm;

extern "C" int xm_img_supported(const XmBlockGeom* g) { return g && img_ok(*g) ? 1 : 0; }

extern "C" int64_t xm_img_gram_bytes(const XmBlockGeom* g) {
  if (!g || !img_ok(*g)) return -1;
  const int64_t K = 9 * g->cin;
  return (int64_t)g->tasks * (K * K + K) * (int64_t)sizeof(double);
}

extern "C" int64_t xm_img_scratch_bytes(const XmBlockGeom* g) {
  if (!g || !img_ok(*g)) return -1;
  return (int64_t)g->tasks * g->cout * (9 * g->cin + 3) * (int64_t)sizeof(double);
}

extern "C" int xm_img_gram(const XmImgArgs* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = common_checks(a, "xm_img_gram")) return rc;
  ImgK k; fill(a, k);
  if (flat_ok(a->g)) return flat_launch_gram(a, k, stream);
  int rc = 0;
  IMG_DISPATCH(a->g.cin, rc = launch_gram<CIN>(a, k, stream));
  return rc;
}

extern "C" int xm_img_fwd(const XmImgArgs* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = common_checks(a, "xm_img_fwd")) return rc;
  const bool flat = flat_ok(a->g);
  XM_REQUIRE(a->w && a->gamma && a->beta && a->mean_invstd && a->p && (flat || (a->zsel && a->sel)),
             "xm_img_fwd: null w/gamma/beta/mean_invstd/p/zsel/sel");
  ImgK k; fill(a, k);
  if (flat) return flat_launch(0, a, k, stream);
  int rc = 0;
  IMG_DISPATCH(a->g.cin, (rc = launch_fwd<CIN, 0>(a, k, stream)));
  return rc;
}

extern "C" int xm_img_dual_fwd(const XmImgArgs* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = common_checks(a, "xm_img_dual_fwd")) return rc;
  const bool flat = flat_ok(a->g);
  XM_REQUIRE(a->w && a->w_dot && a->gamma && a->gamma_dot && a->beta_dot && a->mean_invstd && a->dual_red && a->pdot &&
             (flat ? a->beta != nullptr : (a->zsel && a->sel && a->zdsel)),
             "xm_img_dual_fwd: null w/w_dot/gamma/gamma_dot/beta_dot/mean_invstd/dual_red/zsel/sel/pdot/zdsel (beta for stride 2)");
  ImgK k; fill(a, k);
  if (flat) return flat_launch(2, a, k, stream);
  int rc = 0;
  IMG_DISPATCH(a->g.cin, (rc = launch_fwd<CIN, 1>(a, k, stream)));
  return rc;
}

extern "C" int xm_img_bwd(const XmImgArgs* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = common_checks(a, "xm_img_bwd")) return rc;
  const bool flat = flat_ok(a->g);
  XM_REQUIRE(a->w && a->gamma && a->mean_invstd && a->gp && a->scratch && (flat ? a->beta != nullptr : (a->zsel && a->sel)),
             "xm_img_bwd: null w/gamma/mean_invstd/gp/zsel/sel/scratch (beta for stride 2)");
  XM_REQUIRE((a->out_gamma == nullptr) == (a->out_beta == nullptr), "xm_img_bwd: out_gamma/out_beta must both be given");
  ImgK k; fill(a, k);
  if (flat) {
    if (int rc = flat_launch(1, a, k, stream)) return rc;
    return launch_finalize<1, 0>(a, k, stream);
  }
  int rc = 0;
  IMG_DISPATCH(a->g.cin, (rc = launch_bwd<CIN, 0>(a, k, stream)));
  return rc;
}

extern "C" int xm_img_dual_bwd(const XmImgArgs* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = common_checks(a, "xm_img_dual_bwd")) return rc;
  const bool flat = flat_ok(a->g);
  XM_REQUIRE(a->w && a->w_dot && a->gamma && a->gamma_dot && a->mean_invstd && a->bwd_red && a->dual_red && a->gp &&
             a->ssum && a->scratch && (flat ? a->beta != nullptr : (a->zsel && a->zdsel && a->sel)),
             "xm_img_dual_bwd: null w/w_dot/gamma/gamma_dot/mean_invstd/bwd_red/dual_red/gp/zsel/zdsel/sel/ssum/scratch");
  XM_REQUIRE((a->out_gamma == nullptr) == (a->out_beta == nullptr), "xm_img_dual_bwd: out_gamma/out_beta must both be given");
  ImgK k; fill(a, k);
  if (flat) {
    if (int rc = flat_launch(3, a, k, stream)) return rc;
    return launch_finalize<1, 1>(a, k, stream);
  }
  int rc = 0;
  IMG_DISPATCH(a->g.cin, (rc = launch_bwd<CIN, 1>(a, k, stream)));
  return rc;
}
