// img_flat.cu -- the image block of the Omniglot networks (first ConvBlock with max_pool=False: conv3x3 STRIDE 2 on a
// one-channel image -> BN(train) -> ReLU, core_functions/vision_models.py:158,167,182) in the closed form of
// img_block.cu, without any saved activation: K = 9, so the pre-BN value z = X w of an output element is nine FMAs and
// is simply recomputed from the staged image wherever it is needed (x-hat, the ReLU mask), the dense BatchNorm /
// weight-gradient terms come from the per-task 9x9 Gram matrix of the stride-2 im2col matrix, and the backward is one
// streaming pass over the block-output cotangent:
//   S[co][k] = sum_alive g X[px][k],  s1 = sum_alive g,  s2 = sum_alive g xhat      (alive: y > 0)
// (tangent pass: Sd, e1, e2 with gdot and e3 = sum_alive g zdot).  The closed-form tails are img_block.cu's
// img_finalize_kernel.  Entry points: the xm_img_* functions dispatch here when the geometry has stride 2 / no pool.
#include "img_flat.cuh"

namespace xm {

constexpr int FL_THREADS = 256;

__device__ __forceinline__ const float* flat_image(const ImgK& p, int task, int img) {
  return p.x + ((long long)task * p.rows_per_task + p.row0 + (long long)img * p.row_step) * p.H * p.W;
}

// Gram matrix of the stride-2 im2col matrix (cin = 1): G[k][k'] = sum over output pixels of xp[2y+kh][2x+kw] *
// xp[2y+kh'][2x+kw'], sx[k] likewise; 45 + 9 outputs per task, brute force in double (products of floats are exact).
__global__ void __launch_bounds__(FL_THREADS) flat_gram_kernel(const ImgK p) {
  extern __shared__ __align__(16) double fg_sm[];
  const int Hp = p.H + 2, Wp = p.W + 2;
  double* im = fg_sm;                                  // [Hp][Wp]
  double* red = fg_sm + Hp * Wp;                       // [54]
  const int tid = threadIdx.x, task = blockIdx.y;
  constexpr int NOUT = 54, PARTS = FL_THREADS / 64;    // 4 pixel partitions x 64 output slots (54 used)
  const int o = tid & 63, part = tid >> 6;
  int k1 = 0, k2 = 0;
  if (o < 45) { int rem = o; for (k1 = 0; k1 < 9; ++k1) { if (rem < 9 - k1) { k2 = k1 + rem; break; } rem -= 9 - k1; } }
  else if (o < NOUT) { k1 = o - 45; k2 = -1; }
  const int kh1 = k1 / 3, kw1 = k1 % 3, kh2 = k2 >= 0 ? k2 / 3 : 0, kw2 = k2 >= 0 ? k2 % 3 : 0;
  for (int i = tid; i < NOUT; i += FL_THREADS) red[i] = 0.0;
  double acc = 0.0;
  const int npx = p.hp * p.wp;                         // output pixels per image (hp = hz for this block)
  for (int img = blockIdx.x; img < p.n; img += gridDim.x) {
    const float* X = flat_image(p, task, img);
    __syncthreads();
    for (int i = tid; i < Hp * Wp; i += FL_THREADS) {
      const int yy = i / Wp, xx = i - yy * Wp, y = yy - 1, x = xx - 1;
      im[i] = (y >= 0 && y < p.H && x >= 0 && x < p.W) ? (double)__ldg(X + y * p.W + x) : 0.0;
    }
    __syncthreads();
    if (o < NOUT)
      for (int px = part; px < npx; px += PARTS) {
        const int oy = px / p.wp, ox = px - oy * p.wp;
        const double a = im[(2 * oy + kh1) * Wp + 2 * ox + kw1];
        acc = k2 >= 0 ? fma(a, im[(2 * oy + kh2) * Wp + 2 * ox + kw2], acc) : acc + a;
      }
  }
  __syncthreads();
  if (o < NOUT) atomicAdd(&red[o], acc);
  __syncthreads();
  double* G = p.gram + (long long)task * 90;
  if (tid < 45) {
    atomicAdd(&G[k1 * 9 + k2], red[tid]);
    if (k1 != k2) atomicAdd(&G[k2 * 9 + k1], red[tid]);
  } else if (tid < NOUT) {
    atomicAdd(&G[81 + k1], red[tid]);
  }
}

// MODE 0: forward (writes p, mean_invstd, call_stats)      MODE 1: backward sums (S, s1, s2)
// MODE 2: tangent forward (writes pdot, dual_red)          MODE 3: tangent backward sums (Sd, e1, e2, e3)
// Thread = (pixel slot, 4 channels); its 4 x 9 weights (and tangent weights) live in registers.
template <int MODE>
__global__ void __launch_bounds__(FL_THREADS) flat_kernel(const ImgK p) {
  constexpr int K = 9, NA = K + 3;
  extern __shared__ __align__(16) float fl_sm[];
  const int C = p.cout, c4n = C / 4, slots = FL_THREADS / c4n;
  const int tid = threadIdx.x, task = blockIdx.y, c4 = tid % c4n, slot = tid / c4n, c0 = 4 * c4;
  const int Wp = p.W + 2, Hp = p.H + 2, npx = p.hp * p.wp;
  float* im = fl_sm;                                                    // [Hp][Wp]
  float* chs = im + ((Hp * Wp + 3) & ~3);                               // [4][C]: mean, invstd, d1, d2
  float* xchg = chs + 4 * C;                                            // [warps][c4n][48] flush staging (MODE 1, 3)
  double* dacc = reinterpret_cast<double*>(xchg + (FL_THREADS / 32) * c4n * 48);   // [c4n * 48] (MODE 1, 3)
  double* Gs = dacc + c4n * 48;                                         // [90] (MODE 0, 2 prologue)

  float w[4][K], wd[4][K];
  {
    const float* W = p.w + (long long)task * p.wstride;
    const float* Wd = (MODE >= 2) ? p.wd + (long long)task * p.wdstride : nullptr;
#pragma unroll
    for (int v = 0; v < 4; ++v)
#pragma unroll
      for (int k = 0; k < K; ++k) {
        w[v][k] = __ldg(W + (long long)(c0 + v) * K + k);
        wd[v][k] = (MODE >= 2) ? __ldg(Wd + (long long)(c0 + v) * K + k) : 0.f;
      }
  }
  // ---- per-channel scalars ------------------------------------------------------------------------------------
  if (MODE == 0 || MODE == 2) {
    const double* G = p.gram + (long long)task * 90;
    for (int i = tid; i < 90; i += FL_THREADS) Gs[i] = G[i];
    __syncthreads();
    if (slot == 0) {
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        double ws = 0.0, wgw = 0.0, ds = 0.0, dgw = 0.0;
        for (int k = 0; k < K; ++k) {
          double gw = 0.0;
          for (int k2 = 0; k2 < K; ++k2) gw = fma(Gs[k * K + k2], (double)w[v][k2], gw);
          ws = fma((double)w[v][k], Gs[81 + k], ws);
          wgw = fma((double)w[v][k], gw, wgw);
          if (MODE == 2) { ds = fma((double)wd[v][k], Gs[81 + k], ds); dgw = fma((double)wd[v][k], gw, dgw); }
        }
        const long long mi = ((long long)task * 2) * C + c0 + v;
        if (MODE == 0) {
          const double m = ws / p.cnt;
          double var = wgw / p.cnt - m * m;
          if (var < 0.0) var = 0.0;
          const float mf = (float)m, rf = (float)(1.0 / sqrt(var + (double)p.eps));
          chs[c0 + v] = mf; chs[C + c0 + v] = rf;
          if (blockIdx.x == 0) {
            p.mean_invstd[mi] = mf; p.mean_invstd[mi + C] = rf;
            if (p.call_stats) { p.call_stats[mi] = mf; p.call_stats[mi + C] = (float)(var * (p.cnt / fmax(p.cnt - 1.0, 1.0))); }
          }
        } else {
          const float mf = __ldg(p.mean_invstd + mi), rf = __ldg(p.mean_invstd + mi + C);
          const double e1 = ds / p.cnt, e2 = (double)rf * (dgw / p.cnt - (double)mf * e1);
          chs[c0 + v] = mf; chs[C + c0 + v] = rf; chs[2 * C + c0 + v] = (float)e1; chs[3 * C + c0 + v] = (float)e2;
          if (blockIdx.x == 0) { p.dual_red[mi] = (float)e1; p.dual_red[mi + C] = (float)e2; }
        }
      }
    }
  } else if (slot == 0) {
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const long long mi = ((long long)task * 2) * C + c0 + v;
      chs[c0 + v] = __ldg(p.mean_invstd + mi); chs[C + c0 + v] = __ldg(p.mean_invstd + mi + C);
      if (MODE == 3) { chs[2 * C + c0 + v] = __ldg(p.dual_red + mi); chs[3 * C + c0 + v] = __ldg(p.dual_red + mi + C); }
    }
  }
  if (MODE == 1 || MODE == 3)
    for (int i = tid; i < c4n * 48; i += FL_THREADS) dacc[i] = 0.0;
  __syncthreads();
  float gam[4], bet[4], mean[4], rinv[4], gd[4], bd[4], d1[4], d2[4];
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    gam[v] = __ldg(p.gamma + (long long)task * p.gbstride + c0 + v);
    bet[v] = __ldg(p.beta + (long long)task * p.gbstride + c0 + v);
    mean[v] = chs[c0 + v]; rinv[v] = chs[C + c0 + v];
    d1[v] = (MODE >= 2) ? chs[2 * C + c0 + v] : 0.f; d2[v] = (MODE >= 2) ? chs[3 * C + c0 + v] : 0.f;
    gd[v] = (MODE == 2) ? __ldg(p.gammad + (long long)task * p.gbdstride + c0 + v) : 0.f;
    bd[v] = (MODE == 2) ? __ldg(p.betad + (long long)task * p.gbdstride + c0 + v) : 0.f;
  }
  float S[4][K], st[4][3];
#pragma unroll
  for (int v = 0; v < 4; ++v) {
#pragma unroll
    for (int k = 0; k < K; ++k) S[v][k] = 0.f;
    st[v][0] = st[v][1] = st[v][2] = 0.f;
  }
  // fold the per-thread fp32 sums into the CTA's double accumulators: two pixel slots per warp by shuffle, then the
  // eight warps through shared memory; entry (c4, e) is owned by one thread, so no atomics
  auto flush = [&]() {
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
#pragma unroll
      for (int e = 0; e < NA; ++e) {
        float val = e < K ? S[v][e] : st[v][e - K];
        if (c4n == 16) val += __shfl_xor_sync(0xffffffffu, val, 16);
        else if (c4n == 8) { val += __shfl_xor_sync(0xffffffffu, val, 8); val += __shfl_xor_sync(0xffffffffu, val, 16); }
        if (lane < c4n) xchg[(warp * c4n + lane) * 48 + v * NA + e] = val;
        if (e < K) S[v][e] = 0.f; else st[v][e - K] = 0.f;
      }
    }
    __syncthreads();
    for (int i = tid; i < c4n * 48; i += FL_THREADS) {
      const int cc = i / 48, e = i - cc * 48;
      float s = 0.f;
#pragma unroll
      for (int wv = 0; wv < FL_THREADS / 32; ++wv) s += xchg[(wv * c4n + cc) * 48 + e];
      dacc[i] += (double)s;
    }
    __syncthreads();
  };

  int done = 0;
  for (int img = blockIdx.x; img < p.n; img += gridDim.x) {
    const float* X = flat_image(p, task, img);
    __syncthreads();
    for (int i = tid; i < Hp * Wp; i += FL_THREADS) {
      const int yy = i / Wp, xx = i - yy * Wp, y = yy - 1, x = xx - 1;
      im[i] = (y >= 0 && y < p.H && x >= 0 && x < p.W) ? __ldg(X + y * p.W + x) : 0.f;
    }
    __syncthreads();
    const long long obase = ((long long)task * p.n + img) * npx * C + c0;
    for (int px = slot; px < npx; px += slots) {
      const int oy = px / p.wp, ox = px - oy * p.wp;
      float xv[K];
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) xv[kh * 3 + kw] = im[(2 * oy + kh) * Wp + 2 * ox + kw];
      float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f), gd4 = g4;
      if (MODE == 1 || MODE == 3) g4 = __ldg(reinterpret_cast<const float4*>(p.gp + obase + (long long)px * C));
      if (MODE == 3 && p.gpd) gd4 = __ldg(reinterpret_cast<const float4*>(p.gpd + obase + (long long)px * C));
      const float gv[4] = {g4.x, g4.y, g4.z, g4.w}, gdv[4] = {gd4.x, gd4.y, gd4.z, gd4.w};
      float outv[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        float z = 0.f, zd = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          z = fmaf(w[v][k], xv[k], z);
          if (MODE >= 2) zd = fmaf(wd[v][k], xv[k], zd);
        }
        const float xhat = (z - mean[v]) * rinv[v];
        const float y = fmaf(gam[v], xhat, bet[v]);
        const bool alive = y > 0.f;
        if (MODE == 0) {
          outv[v] = alive ? y : 0.f;
        } else if (MODE == 2) {
          const float xhd = rinv[v] * (zd - d1[v] - xhat * d2[v]);
          outv[v] = alive ? gd[v] * xhat + gam[v] * xhd + bd[v] : 0.f;
        } else {
          const float g = alive ? gv[v] : 0.f, gdot = alive ? gdv[v] : 0.f;
          const float c = MODE == 1 ? g : gdot;
          st[v][0] += c;
          st[v][1] = fmaf(c, xhat, st[v][1]);
          if (MODE == 3) st[v][2] = fmaf(g, zd, st[v][2]);
#pragma unroll
          for (int k = 0; k < K; ++k) S[v][k] = fmaf(c, xv[k], S[v][k]);
        }
      }
      if (MODE == 0) *reinterpret_cast<float4*>(p.p + obase + (long long)px * C) = make_float4(outv[0], outv[1], outv[2], outv[3]);
      if (MODE == 2) *reinterpret_cast<float4*>(p.pdot + obase + (long long)px * C) = make_float4(outv[0], outv[1], outv[2], outv[3]);
    }
    if ((MODE == 1 || MODE == 3) && (++done & 7) == 0) flush();
  }
  if (MODE == 1 || MODE == 3) {
    flush();
    double* out = p.scratch + (long long)task * C * NA;
    for (int i = tid; i < c4n * 48; i += FL_THREADS) {
      const int cc = i / 48, e = i - cc * 48, v = e / NA, q = e - v * NA;
      atomicAdd(&out[(4 * cc + v) * NA + q], dacc[i]);
    }
  }
}

static size_t flat_smem(const XmBlockGeom& g) {
  const size_t im = (((size_t)(g.hin + 2) * (g.win + 2) + 3) & ~(size_t)3) * 4;
  const int c4n = g.cout / 4;
  return im + (size_t)4 * g.cout * 4 + (size_t)(FL_THREADS / 32) * c4n * 48 * 4 + (size_t)c4n * 48 * 8 + 90 * 8 + 16;
}

int flat_ok(const XmBlockGeom& g) {
  return geom_ok(g) && g.cin == 1 && g.stride == 2 && g.pool == 0 && (g.cout == 32 || g.cout == 64) &&
         flat_smem(g) <= 160 * 1024;
}

int flat_launch_gram(const XmImgArgs* a, ImgK& k, cudaStream_t stream) {
  const XmBlockGeom& g = a->g;
  const size_t smem = ((size_t)(g.hin + 2) * (g.win + 2) + 54) * sizeof(double);
  XM_REQUIRE(smem <= 48 * 1024, "xm_img_gram: image too large");
  XM_CUDA(cudaMemsetAsync(a->gram, 0, (size_t)g.tasks * 90 * sizeof(double), stream));
  int per_task = (2 * num_sms() + g.tasks - 1) / g.tasks;
  if (per_task > g.n) per_task = g.n;
  if (per_task < 1) per_task = 1;
  flat_gram_kernel<<<dim3(per_task, g.tasks), FL_THREADS, smem, stream>>>(k);
  return launched("xm_img_gram(stride 2)");
}

template <int MODE>
static int flat_launch_mode(const XmImgArgs* a, ImgK& k, cudaStream_t stream, const char* what) {
  const XmBlockGeom& g = a->g;
  const size_t smem = flat_smem(g);
  XM_CUDA(cudaFuncSetAttribute(flat_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  int per_task = wave_ctas((const void*)flat_kernel<MODE>, FL_THREADS, smem) / g.tasks;
  if (per_task > g.n) per_task = g.n;
  if (per_task < 1) per_task = 1;
  if (MODE == 1 || MODE == 3)
    XM_CUDA(cudaMemsetAsync(a->scratch, 0, (size_t)g.tasks * g.cout * 12 * sizeof(double), stream));
  flat_kernel<MODE><<<dim3(per_task, g.tasks), FL_THREADS, smem, stream>>>(k);
  return launched(what);
}

int flat_launch(int mode, const XmImgArgs* a, ImgK& k, cudaStream_t stream) {
  switch (mode) {
    case 0: return flat_launch_mode<0>(a, k, stream, "xm_img_fwd(stride 2)");
    case 1: return flat_launch_mode<1>(a, k, stream, "xm_img_bwd(stride 2)");
    case 2: return flat_launch_mode<2>(a, k, stream, "xm_img_dual_fwd(stride 2)");
    default: return flat_launch_mode<3>(a, k, stream, "xm_img_dual_bwd(stride 2)");
  }
}

}  // namespace xm
