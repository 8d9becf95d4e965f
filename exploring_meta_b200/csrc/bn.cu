// bn.cu -- train-mode BatchNorm + ReLU + 2x2 max-pool, fused: forward, backward, and the tangent
// ("dual") versions of both that the forward-over-reverse second-order pass needs.
//
// All four entry points are HBM-streaming kernels over the pre-BN conv output z (NHWC, fp32):
//   fwd       : 1 pass   read z                  -> write pooled p                     (+ mean/invstd)
//   bwd       : 2 passes read z, gp (reduce)     ;  read z, gp -> write gz            (+ dgamma/dbeta)
//   dual_fwd  : 1 pass   read z, zdot            -> write pdot
//   dual_bwd  : 2 passes read z, zdot, gp, gpdot ;  same -> write gz, gzdot           (+ tangents of dgamma/dbeta)
// The pool arg-max / ReLU mask is never stored: it is recomputed from z (4 FMAs per element), which
// is cheaper than a mask round trip through HBM.  Per-(task, channel) reductions are accumulated in
// double (thread -> shared -> one global atomic per channel per CTA).  Closed forms: SURVEY App. F.
// A work item is one pooling window (or one pixel when the block has no pool) x VEC channels;
// a thread keeps a fixed channel group, so gamma/beta/mean/invstd live in registers.
// Reference ops replaced: native_batch_norm(train) / relu / max_pool2d_with_indices
// (core_functions/vision_models.py:190-192), their backward ops, batchnorm_double_backward and the
// mask/gather double-backward ops (vision/maml_vision.py:112).
#include "common.cuh"

namespace xm {

struct BnK {
  int n, hz, wz, hp, wp, pool, C;
  int wh, ww;                        // window grid: pool ? ceil(hz/2) x ceil(wz/2) : hz x wz
  double cnt;                        // n*hz*wz
  float eps, scale;
  const float* z; const float* zdot;
  const double* sums; const double* dsums;
  const float* gamma; const float* beta; long long gb_stride;
  const float* gamma_dot; const float* beta_dot; long long gbdot_stride;
  float* mean_invstd; float* call_stats; float* bwd_red; float* dual_red;
  float* p; float* pdot;
  const float* gp; const float* gpdot;
  float* gz; float* gzdot;
  float* out_gamma; float* out_beta; long long out_stride;
  const float* base_gamma; const float* base_beta; long long base_stride;
  double* scratch;                   // [task][4][C]
};

template <int VEC> struct Vec;
template <> struct Vec<4> {
  static __device__ __forceinline__ void ld(const float* p, float (&v)[4]) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void st(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Vec<1> {
  static __device__ __forceinline__ void ld(const float* p, float (&v)[1]) { v[0] = __ldg(p); }
  static __device__ __forceinline__ void st(float* p, const float (&v)[1]) { *p = v[0]; }
};

// One pooling window of one image: offsets of its (up to 4) z elements and of its pooled element.
struct Window {
  long long zoff[4];
  long long poff;
  unsigned valid;    // bit d set: element d = (dy, dx) row-major of the window is inside the map
  bool pooled;       // complete window with a pooled output (every element valid)
};

__device__ __forceinline__ Window make_window(const BnK& k, int task, int item_px, int c0) {
  // item_px indexes (img, wy, wx); 32-bit index arithmetic (the host checks n*wh*ww < 2^31)
  Window w;
  const int row = item_px / k.ww;
  const int wx = item_px - row * k.ww;
  const int img = row / k.wh;
  const int wy = row - img * k.wh;
  const long long zimg = ((long long)task * k.n + img) * k.hz;
  if (!k.pool) {
    w.zoff[0] = ((zimg + wy) * k.wz + wx) * k.C + c0;
    w.valid = 1u;
    w.pooled = true;
    w.poff = w.zoff[0];
    return w;
  }
  w.pooled = wy < k.hp && wx < k.wp;
  w.valid = 0u;
#pragma unroll
  for (int d = 0; d < 4; ++d) {
    const int y = 2 * wy + (d >> 1), x = 2 * wx + (d & 1);
    w.zoff[d] = ((zimg + y) * k.wz + x) * k.C + c0;
    if (y < k.hz && x < k.wz) w.valid |= 1u << d;
  }
  w.poff = w.pooled ? ((((long long)task * k.n + img) * k.hp + wy) * k.wp + wx) * k.C + c0 : 0;
  return w;
}

// Per-thread channel constants.
template <int VEC> struct Chan {
  float gamma[VEC], beta[VEC], mean[VEC], r[VEC];
};

template <int VEC>
__device__ __forceinline__ void load_gb(const BnK& k, int task, int c0, Chan<VEC>& ch) {
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    ch.gamma[v] = __ldg(k.gamma + (long long)task * k.gb_stride + c0 + v);
    ch.beta[v] = __ldg(k.beta + (long long)task * k.gb_stride + c0 + v);
  }
}
template <int VEC>
__device__ __forceinline__ void load_mi(const BnK& k, int task, int c0, Chan<VEC>& ch) {
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    ch.mean[v] = __ldg(k.mean_invstd + ((long long)task * 2) * k.C + c0 + v);
    ch.r[v] = __ldg(k.mean_invstd + ((long long)task * 2 + 1) * k.C + c0 + v);
  }
}

// arg-max element of the window (first maximum in row-major order) and whether it passes the ReLU.
// For windows that are not pooled (odd edge) nothing is selected.
template <int VEC>
__device__ __forceinline__ void select(const Window& w, const Chan<VEC>& ch, const float (&z)[4][VEC],
                                       int (&sel)[VEC]) {
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    int best = 0;
    float ybest = fmaf(ch.gamma[v], (z[0][v] - ch.mean[v]) * ch.r[v], ch.beta[v]);
#pragma unroll
    for (int d = 1; d < 4; ++d)
      if (w.valid == 0xfu) {
        const float y = fmaf(ch.gamma[v], (z[d][v] - ch.mean[v]) * ch.r[v], ch.beta[v]);
        if (y > ybest) { ybest = y; best = d; }
      }
    sel[v] = (w.pooled && ybest > 0.f) ? best : -1;
  }
}

// shared double accumulators [NSTAT][C] -> global scratch [task][4][C]
template <int VEC, int NSTAT>
__device__ __forceinline__ void flush_stats(const BnK& k, int task, int c0, double (&acc)[NSTAT][VEC],
                                            double* sh) {
  for (int i = threadIdx.x; i < NSTAT * k.C; i += blockDim.x) sh[i] = 0.0;
  __syncthreads();
#pragma unroll
  for (int s = 0; s < NSTAT; ++s)
#pragma unroll
    for (int v = 0; v < VEC; ++v) atomicAdd(&sh[s * k.C + c0 + v], acc[s][v]);
  __syncthreads();
  for (int i = threadIdx.x; i < NSTAT * k.C; i += blockDim.x)
    atomicAdd(&k.scratch[((long long)task * 4) * k.C + i], sh[i]);
}

// -------------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(256) bn_fwd_kernel(const BnK k) {
  const int task = blockIdx.y;
  const int cq = k.C / VEC;
  const int c0 = (threadIdx.x % cq) * VEC;
  Chan<VEC> ch;
  load_gb<VEC>(k, task, c0, ch);
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    const double s0 = k.sums[((long long)task * 2) * k.C + c0 + v];
    const double s1 = k.sums[((long long)task * 2 + 1) * k.C + c0 + v];
    const double mean = s0 / k.cnt;
    double var = s1 / k.cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    ch.mean[v] = (float)mean;
    ch.r[v] = (float)(1.0 / sqrt(var + (double)k.eps));
    if (blockIdx.x == 0 && threadIdx.x < cq) {
      k.mean_invstd[((long long)task * 2) * k.C + c0 + v] = ch.mean[v];
      k.mean_invstd[((long long)task * 2 + 1) * k.C + c0 + v] = ch.r[v];
      if (k.call_stats) {
        k.call_stats[((long long)task * 2) * k.C + c0 + v] = ch.mean[v];
        k.call_stats[((long long)task * 2 + 1) * k.C + c0 + v] = (float)(var * (k.cnt / fmax(k.cnt - 1.0, 1.0)));
      }
    }
  }
  // pooled windows only: wh/ww are set to hp/wp by the host for this kernel
  const int npx = k.n * k.wh * k.ww, pslots = blockDim.x / cq;
  for (int px = blockIdx.x * pslots + threadIdx.x / cq; px < npx; px += gridDim.x * pslots) {
    const Window w = make_window(k, task, px, c0);
    float out[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) out[v] = 0.f;       // ReLU floor
#pragma unroll
    for (int d = 0; d < 4; ++d)
      if (w.valid >> d & 1u) {
        float z[VEC];
        Vec<VEC>::ld(k.z + w.zoff[d], z);
#pragma unroll
        for (int v = 0; v < VEC; ++v)
          out[v] = fmaxf(out[v], fmaf(ch.gamma[v], (z[v] - ch.mean[v]) * ch.r[v], ch.beta[v]));
      }
    Vec<VEC>::st(k.p + w.poff, out);
  }
}

// bwd pass 1: s1 = sum gbn, s2 = sum gbn*xhat over the selected elements
template <int VEC>
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const BnK k) {
  extern __shared__ double sh[];
  const int task = blockIdx.y;
  const int cq = k.C / VEC;
  const int c0 = (threadIdx.x % cq) * VEC;
  Chan<VEC> ch;
  load_gb<VEC>(k, task, c0, ch);
  load_mi<VEC>(k, task, c0, ch);
  // fp32 partials over short runs (<= 16 windows), folded into double: the double pipe (and its conversions)
  // would otherwise bound this pass instead of HBM
  double acc[2][VEC];
  float fa[2][VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) { acc[0][v] = acc[1][v] = 0.0; fa[0][v] = fa[1][v] = 0.f; }
  int run = 0;
  const int npx = k.n * k.wh * k.ww, pslots = blockDim.x / cq;     // host sets wh/ww = hp/wp here
  for (int px = blockIdx.x * pslots + threadIdx.x / cq; px < npx; px += gridDim.x * pslots) {
    const Window w = make_window(k, task, px, c0);
    float z[4][VEC];
#pragma unroll
    for (int d = 0; d < 4; ++d)
      if (w.valid >> d & 1u) Vec<VEC>::ld(k.z + w.zoff[d], z[d]);
    int sel[VEC];
    select<VEC>(w, ch, z, sel);
    float gp[VEC];
    Vec<VEC>::ld(k.gp + w.poff, gp);
#pragma unroll
    for (int v = 0; v < VEC; ++v)
      if (sel[v] >= 0) {
        float zs = z[0][v];
#pragma unroll
        for (int d = 1; d < 4; ++d) if (sel[v] == d) zs = z[d][v];
        const float xhat = (zs - ch.mean[v]) * ch.r[v];
        fa[0][v] += gp[v];
        fa[1][v] = fmaf(gp[v], xhat, fa[1][v]);
      }
    if ((++run & 15) == 0) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        acc[0][v] += (double)fa[0][v]; acc[1][v] += (double)fa[1][v];
        fa[0][v] = fa[1][v] = 0.f;
      }
    }
  }
#pragma unroll
  for (int v = 0; v < VEC; ++v) { acc[0][v] += (double)fa[0][v]; acc[1][v] += (double)fa[1][v]; }
  flush_stats<VEC, 2>(k, task, c0, acc, sh);
}

// bwd pass 2: gz = gamma*r*(gbn - m1 - xhat*m2) for every element of z
template <int VEC>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const BnK k) {
  const int task = blockIdx.y;
  const int cq = k.C / VEC;
  const int c0 = (threadIdx.x % cq) * VEC;
  Chan<VEC> ch;
  load_gb<VEC>(k, task, c0, ch);
  load_mi<VEC>(k, task, c0, ch);
  float m1[VEC], m2[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    const double s1 = k.scratch[((long long)task * 4) * k.C + c0 + v];
    const double s2 = k.scratch[((long long)task * 4 + 1) * k.C + c0 + v];
    m1[v] = (float)(s1 / k.cnt);
    m2[v] = (float)(s2 / k.cnt);
    if (blockIdx.x == 0 && threadIdx.x < cq) {
      if (k.bwd_red) {
        k.bwd_red[((long long)task * 2) * k.C + c0 + v] = m1[v];
        k.bwd_red[((long long)task * 2 + 1) * k.C + c0 + v] = m2[v];
      }
      if (k.out_gamma) {
        const float bg = k.base_gamma ? k.base_gamma[(long long)task * k.base_stride + c0 + v] : 0.f;
        const float bb = k.base_beta ? k.base_beta[(long long)task * k.base_stride + c0 + v] : 0.f;
        k.out_gamma[(long long)task * k.out_stride + c0 + v] = bg + k.scale * (float)s2;
        k.out_beta[(long long)task * k.out_stride + c0 + v] = bb + k.scale * (float)s1;
      }
    }
  }
  const int npx = k.n * k.wh * k.ww, pslots = blockDim.x / cq;     // all windows, incl. odd edges
  for (int px = blockIdx.x * pslots + threadIdx.x / cq; px < npx; px += gridDim.x * pslots) {
    const Window w = make_window(k, task, px, c0);
    float z[4][VEC];
#pragma unroll
    for (int d = 0; d < 4; ++d)
      if (w.valid >> d & 1u) Vec<VEC>::ld(k.z + w.zoff[d], z[d]);
    int sel[VEC];
    select<VEC>(w, ch, z, sel);
    float gp[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) gp[v] = 0.f;
    if (w.pooled) Vec<VEC>::ld(k.gp + w.poff, gp);
#pragma unroll
    for (int d = 0; d < 4; ++d)
      if (w.valid >> d & 1u) {
        float o[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          const float xhat = (z[d][v] - ch.mean[v]) * ch.r[v];
          const float gbn = (sel[v] == d) ? gp[v] : 0.f;
          o[v] = ch.gamma[v] * ch.r[v] * (gbn - m1[v] - xhat * m2[v]);
        }
        Vec<VEC>::st(k.gz + w.zoff[d], o);
      }
  }
}

// tangent of the forward: pdot at the selected element
template <int VEC>
__global__ void __launch_bounds__(256) bn_dual_fwd_kernel(const BnK k) {
  const int task = blockIdx.y;
  const int cq = k.C / VEC;
  const int c0 = (threadIdx.x % cq) * VEC;
  Chan<VEC> ch;
  load_gb<VEC>(k, task, c0, ch);
  load_mi<VEC>(k, task, c0, ch);
  float gd[VEC], bd[VEC], d1[VEC], d2[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    gd[v] = __ldg(k.gamma_dot + (long long)task * k.gbdot_stride + c0 + v);
    bd[v] = __ldg(k.beta_dot + (long long)task * k.gbdot_stride + c0 + v);
    const double s0 = k.dsums[((long long)task * 2) * k.C + c0 + v];
    const double s1 = k.dsums[((long long)task * 2 + 1) * k.C + c0 + v];
    const double e1 = s0 / k.cnt;
    const double e2 = (double)ch.r[v] * (s1 / k.cnt - (double)ch.mean[v] * e1);
    d1[v] = (float)e1;
    d2[v] = (float)e2;
    if (blockIdx.x == 0 && threadIdx.x < cq) {
      k.dual_red[((long long)task * 2) * k.C + c0 + v] = d1[v];
      k.dual_red[((long long)task * 2 + 1) * k.C + c0 + v] = d2[v];
    }
  }
  const int npx = k.n * k.wh * k.ww, pslots = blockDim.x / cq;     // pooled windows (host sets hp/wp)
  for (int px = blockIdx.x * pslots + threadIdx.x / cq; px < npx; px += gridDim.x * pslots) {
    const Window w = make_window(k, task, px, c0);
    float z[4][VEC], zd[4][VEC];
#pragma unroll
    for (int d = 0; d < 4; ++d)
      if (w.valid >> d & 1u) { Vec<VEC>::ld(k.z + w.zoff[d], z[d]); Vec<VEC>::ld(k.zdot + w.zoff[d], zd[d]); }
    int sel[VEC];
    select<VEC>(w, ch, z, sel);
    float o[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      o[v] = 0.f;
      if (sel[v] >= 0) {
        float zs = z[0][v], zds = zd[0][v];
#pragma unroll
        for (int d = 1; d < 4; ++d) if (sel[v] == d) { zs = z[d][v]; zds = zd[d][v]; }
        const float xhat = (zs - ch.mean[v]) * ch.r[v];
        const float xhd = ch.r[v] * (zds - d1[v] - xhat * d2[v]);
        o[v] = gd[v] * xhat + ch.gamma[v] * xhd + bd[v];
      }
    }
    Vec<VEC>::st(k.pdot + w.poff, o);
  }
}

// dual bwd pass 1: e1 = sum gbnd, e2 = sum gbnd*xhat, e3 = sum gbn*zdot over the selected elements
template <int VEC>
__global__ void __launch_bounds__(256) bn_dual_bwd_reduce_kernel(const BnK k) {
  extern __shared__ double sh[];
  const int task = blockIdx.y;
  const int cq = k.C / VEC;
  const int c0 = (threadIdx.x % cq) * VEC;
  Chan<VEC> ch;
  load_gb<VEC>(k, task, c0, ch);
  load_mi<VEC>(k, task, c0, ch);
  double acc[3][VEC];
  float fa[3][VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) { acc[0][v] = acc[1][v] = acc[2][v] = 0.0; fa[0][v] = fa[1][v] = fa[2][v] = 0.f; }
  int run = 0;
  const int npx = k.n * k.wh * k.ww, pslots = blockDim.x / cq;     // pooled windows
  for (int px = blockIdx.x * pslots + threadIdx.x / cq; px < npx; px += gridDim.x * pslots) {
    const Window w = make_window(k, task, px, c0);
    float z[4][VEC], zd[4][VEC];
#pragma unroll
    for (int d = 0; d < 4; ++d)
      if (w.valid >> d & 1u) { Vec<VEC>::ld(k.z + w.zoff[d], z[d]); Vec<VEC>::ld(k.zdot + w.zoff[d], zd[d]); }
    int sel[VEC];
    select<VEC>(w, ch, z, sel);
    float gp[VEC], gpd[VEC];
    Vec<VEC>::ld(k.gp + w.poff, gp);
#pragma unroll
    for (int v = 0; v < VEC; ++v) gpd[v] = 0.f;
    if (k.gpdot) Vec<VEC>::ld(k.gpdot + w.poff, gpd);
#pragma unroll
    for (int v = 0; v < VEC; ++v)
      if (sel[v] >= 0) {
        float zs = z[0][v], zds = zd[0][v];
#pragma unroll
        for (int d = 1; d < 4; ++d) if (sel[v] == d) { zs = z[d][v]; zds = zd[d][v]; }
        const float xhat = (zs - ch.mean[v]) * ch.r[v];
        fa[0][v] += gpd[v];
        fa[1][v] = fmaf(gpd[v], xhat, fa[1][v]);
        fa[2][v] = fmaf(gp[v], zds, fa[2][v]);
      }
    if ((++run & 15) == 0) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        acc[0][v] += (double)fa[0][v]; acc[1][v] += (double)fa[1][v]; acc[2][v] += (double)fa[2][v];
        fa[0][v] = fa[1][v] = fa[2][v] = 0.f;
      }
    }
  }
#pragma unroll
  for (int v = 0; v < VEC; ++v) { acc[0][v] += (double)fa[0][v]; acc[1][v] += (double)fa[1][v]; acc[2][v] += (double)fa[2][v]; }
  flush_stats<VEC, 3>(k, task, c0, acc, sh);
}

// dual bwd pass 2: gz and its tangent gzdot for every element of z
template <int VEC>
__global__ void __launch_bounds__(256) bn_dual_bwd_apply_kernel(const BnK k) {
  const int task = blockIdx.y;
  const int cq = k.C / VEC;
  const int c0 = (threadIdx.x % cq) * VEC;
  Chan<VEC> ch;
  load_gb<VEC>(k, task, c0, ch);
  load_mi<VEC>(k, task, c0, ch);
  float m1[VEC], m2[VEC], d1[VEC], d2[VEC], e1[VEC], m2dot[VEC], coef[VEC], gr[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    m1[v] = __ldg(k.bwd_red + ((long long)task * 2) * k.C + c0 + v);
    m2[v] = __ldg(k.bwd_red + ((long long)task * 2 + 1) * k.C + c0 + v);
    d1[v] = __ldg(k.dual_red + ((long long)task * 2) * k.C + c0 + v);
    d2[v] = __ldg(k.dual_red + ((long long)task * 2 + 1) * k.C + c0 + v);
    const float gd = __ldg(k.gamma_dot + (long long)task * k.gbdot_stride + c0 + v);
    const double s1 = k.scratch[((long long)task * 4) * k.C + c0 + v] / k.cnt;       // <gbnd>
    const double s2 = k.scratch[((long long)task * 4 + 1) * k.C + c0 + v] / k.cnt;   // <gbnd*xhat>
    const double s3 = k.scratch[((long long)task * 4 + 2) * k.C + c0 + v] / k.cnt;   // <gbn*zdot>
    const double r = (double)ch.r[v];
    const double q = r * (s3 - (double)d1[v] * (double)m1[v] - (double)d2[v] * (double)m2[v]);   // <gbn*xhat_dot>
    e1[v] = (float)s1;
    m2dot[v] = (float)(s2 + q);
    const float rdot = -ch.r[v] * ch.r[v] * d2[v];
    coef[v] = gd * ch.r[v] + ch.gamma[v] * rdot;
    gr[v] = ch.gamma[v] * ch.r[v];
    if (blockIdx.x == 0 && threadIdx.x < cq && k.out_gamma) {
      const float bg = k.base_gamma ? k.base_gamma[(long long)task * k.base_stride + c0 + v] : 0.f;
      const float bb = k.base_beta ? k.base_beta[(long long)task * k.base_stride + c0 + v] : 0.f;
      k.out_gamma[(long long)task * k.out_stride + c0 + v] = bg + k.scale * (float)((s2 + q) * k.cnt);
      k.out_beta[(long long)task * k.out_stride + c0 + v] = bb + k.scale * (float)(s1 * k.cnt);
    }
  }
  const int npx = k.n * k.wh * k.ww, pslots = blockDim.x / cq;     // all windows
  for (int px = blockIdx.x * pslots + threadIdx.x / cq; px < npx; px += gridDim.x * pslots) {
    const Window w = make_window(k, task, px, c0);
    float z[4][VEC], zd[4][VEC];
#pragma unroll
    for (int d = 0; d < 4; ++d)
      if (w.valid >> d & 1u) { Vec<VEC>::ld(k.z + w.zoff[d], z[d]); Vec<VEC>::ld(k.zdot + w.zoff[d], zd[d]); }
    int sel[VEC];
    select<VEC>(w, ch, z, sel);
    float gp[VEC], gpd[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) gp[v] = gpd[v] = 0.f;
    if (w.pooled) {
      Vec<VEC>::ld(k.gp + w.poff, gp);
      if (k.gpdot) Vec<VEC>::ld(k.gpdot + w.poff, gpd);
    }
#pragma unroll
    for (int d = 0; d < 4; ++d)
      if (w.valid >> d & 1u) {
        float o[VEC], od[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          const float xhat = (z[d][v] - ch.mean[v]) * ch.r[v];
          const float xhd = ch.r[v] * (zd[d][v] - d1[v] - xhat * d2[v]);
          const bool s = sel[v] == d;
          const float gbn = s ? gp[v] : 0.f, gbnd = s ? gpd[v] : 0.f;
          const float proj = gbn - m1[v] - xhat * m2[v];
          o[v] = gr[v] * proj;
          od[v] = coef[v] * proj + gr[v] * (gbnd - e1[v] - xhd * m2[v] - xhat * m2dot[v]);
        }
        if (k.gz) Vec<VEC>::st(k.gz + w.zoff[d], o);
        Vec<VEC>::st(k.gzdot + w.zoff[d], od);
      }
  }
}

// ---- host side -----------------------------------------------------------------------------------
static int fill(const XmBnArgs* a, BnK& k, bool pooled_grid, int& vec, int& threads, int& blocks) {
  const XmBlockGeom& g = a->g;
  k = BnK{};
  k.n = g.n; k.hz = g.hz; k.wz = g.wz; k.hp = g.hp; k.wp = g.wp; k.pool = g.pool; k.C = g.cout;
  if (g.pool) {
    k.wh = pooled_grid ? g.hp : (g.hz + 1) / 2;
    k.ww = pooled_grid ? g.wp : (g.wz + 1) / 2;
  } else { k.wh = g.hz; k.ww = g.wz; }
  k.cnt = (double)g.n * g.hz * g.wz;
  k.eps = a->eps; k.scale = a->scale;
  k.z = a->z; k.zdot = a->zdot; k.sums = a->sums; k.dsums = a->dsums;
  k.gamma = a->gamma; k.beta = a->beta; k.gb_stride = a->gb_task_stride;
  k.gamma_dot = a->gamma_dot; k.beta_dot = a->beta_dot; k.gbdot_stride = a->gbdot_task_stride;
  k.mean_invstd = a->mean_invstd; k.call_stats = a->call_stats; k.bwd_red = a->bwd_red; k.dual_red = a->dual_red;
  k.p = a->p; k.pdot = a->pdot; k.gp = a->gp; k.gpdot = a->gpdot; k.gz = a->gz; k.gzdot = a->gzdot;
  k.out_gamma = a->out_gamma; k.out_beta = a->out_beta; k.out_stride = a->out_task_stride;
  k.base_gamma = a->base_gamma; k.base_beta = a->base_beta; k.base_stride = a->base_task_stride;
  k.scratch = a->scratch;
  vec = (g.cout % 4 == 0) ? 4 : 1;
  const int cq = g.cout / vec;
  if (cq > 256) return 0;
  threads = (256 / cq) * cq;
  const long long items = (long long)g.n * k.wh * k.ww * cq;
  if ((long long)g.n * k.wh * k.ww >= (1LL << 31)) return 0;
  long long want = (items + threads - 1) / threads;
  blocks = (int)(want < 65535 ? want : 65535);      // BN_LAUNCH clamps to one resident wave of the kernel
  if (blocks < 1) blocks = 1;
  return 1;
}

#define BN_LAUNCH(kernel, smem)                                                            \
  do {                                                                                     \
    const void* fn_ = vec == 4 ? (const void*)kernel<4> : (const void*)kernel<1>;          \
    int cap_ = wave_ctas(fn_, threads, smem) / a->g.tasks;                                 \
    if (cap_ < 1) cap_ = 1;                                                                \
    dim3 grid(blocks < cap_ ? blocks : cap_, a->g.tasks);                                  \
    if (vec == 4) kernel<4><<<grid, threads, smem, stream>>>(k);                           \
    else kernel<1><<<grid, threads, smem, stream>>>(k);                                    \
  } while (0)

}  // namespace xm

using namespace xm;

extern "C" int64_t xm_bn_scratch_bytes(const XmBlockGeom* g) {
  if (!g || g->tasks <= 0 || g->cout <= 0) return -1;
  return (int64_t)g->tasks * 4 * g->cout * (int64_t)sizeof(double);
}

static int bn_common_checks(const XmBnArgs* a, const char* who) {
  XM_REQUIRE(a != nullptr, "%s: null args", who);
  XM_REQUIRE(geom_ok(a->g), "%s: inconsistent block geometry", who);
  XM_REQUIRE(a->z && a->gamma && a->beta, "%s: null z/gamma/beta", who);
  XM_REQUIRE(a->g.cout % 4 != 0 || a->g.cout / 4 <= 256, "%s: too many channels", who);
  XM_REQUIRE(a->g.cout % 4 == 0 || a->g.cout <= 256, "%s: too many channels", who);
  return 0;
}

extern "C" int xm_bn_fwd(const XmBnArgs* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = bn_common_checks(a, "xm_bn_fwd")) return rc;
  XM_REQUIRE(a->sums && a->mean_invstd && a->p, "xm_bn_fwd: null sums/mean_invstd/p");
  BnK k; int vec, threads, blocks;
  XM_REQUIRE(fill(a, k, true, vec, threads, blocks), "xm_bn_fwd: unsupported channel count");
  BN_LAUNCH(bn_fwd_kernel, 0);
  return launched("xm_bn_fwd");
}

extern "C" int xm_bn_bwd(const XmBnArgs* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = bn_common_checks(a, "xm_bn_bwd")) return rc;
  XM_REQUIRE(a->gp && a->mean_invstd && a->gz && a->scratch, "xm_bn_bwd: null gp/mean_invstd/gz/scratch");
  XM_REQUIRE((a->out_gamma == nullptr) == (a->out_beta == nullptr), "xm_bn_bwd: out_gamma/out_beta must both be given");
  BnK k; int vec, threads, blocks;
  XM_REQUIRE(fill(a, k, true, vec, threads, blocks), "xm_bn_bwd: unsupported channel count");
  XM_CUDA(cudaMemsetAsync(a->scratch, 0, (size_t)a->g.tasks * 4 * a->g.cout * sizeof(double), stream));
  BN_LAUNCH(bn_bwd_reduce_kernel, 2 * a->g.cout * sizeof(double));
  if (int rc = launched("xm_bn_bwd(reduce)")) return rc;
  XM_REQUIRE(fill(a, k, false, vec, threads, blocks), "xm_bn_bwd: unsupported channel count");
  BN_LAUNCH(bn_bwd_apply_kernel, 0);
  return launched("xm_bn_bwd(apply)");
}

extern "C" int xm_bn_dual_fwd(const XmBnArgs* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = bn_common_checks(a, "xm_bn_dual_fwd")) return rc;
  XM_REQUIRE(a->zdot && a->dsums && a->mean_invstd && a->gamma_dot && a->beta_dot && a->pdot && a->dual_red,
             "xm_bn_dual_fwd: null zdot/dsums/mean_invstd/gamma_dot/beta_dot/pdot/dual_red");
  BnK k; int vec, threads, blocks;
  XM_REQUIRE(fill(a, k, true, vec, threads, blocks), "xm_bn_dual_fwd: unsupported channel count");
  BN_LAUNCH(bn_dual_fwd_kernel, 0);
  return launched("xm_bn_dual_fwd");
}

extern "C" int xm_bn_dual_bwd(const XmBnArgs* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = bn_common_checks(a, "xm_bn_dual_bwd")) return rc;
  XM_REQUIRE(a->zdot && a->gp && a->mean_invstd && a->bwd_red && a->dual_red && a->gamma_dot &&
             a->gzdot && a->scratch, "xm_bn_dual_bwd: null zdot/gp/mean_invstd/bwd_red/dual_red/gamma_dot/gz/gzdot/scratch");
  XM_REQUIRE((a->out_gamma == nullptr) == (a->out_beta == nullptr), "xm_bn_dual_bwd: out_gamma/out_beta must both be given");
  BnK k; int vec, threads, blocks;
  XM_REQUIRE(fill(a, k, true, vec, threads, blocks), "xm_bn_dual_bwd: unsupported channel count");
  XM_CUDA(cudaMemsetAsync(a->scratch, 0, (size_t)a->g.tasks * 4 * a->g.cout * sizeof(double), stream));
  BN_LAUNCH(bn_dual_bwd_reduce_kernel, 3 * a->g.cout * sizeof(double));
  if (int rc = launched("xm_bn_dual_bwd(reduce)")) return rc;
  XM_REQUIRE(fill(a, k, false, vec, threads, blocks), "xm_bn_dual_bwd: unsupported channel count");
  BN_LAUNCH(bn_dual_bwd_apply_kernel, 0);
  return launched("xm_bn_dual_bwd(apply)");
}
