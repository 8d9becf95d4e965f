// head.cu -- classifier head kernels: one CTA per task, everything between the block-4 features and the
// loss fused into one launch.
//   xm_head      : linear (+ spatial mean for Omniglot) + softmax cross-entropy (mean) forward, loss,
//                  arg-max correct count, and backward (feature grad + parameter grads through the axpy
//                  epilogue = fused inner SGD step); `dual` additionally propagates tangents (softmax-CE
//                  Hessian term included) for the forward-over-reverse second-order pass.
//   xm_anil_head : ANIL's head-only adaptation -- `steps` inner steps on the support feature rows, query
//                  loss / accuracy, and the full second-order outer gradient w.r.t. the head
//                  initialisation and ALL feature rows, in one kernel per task.
// Reference ops replaced: aten::linear + cross_entropy fwd/bwd/double-bwd (core_functions/vision_models.py:107-110,
// :51-55; vision/maml_vision.py:86), accuracy (core_functions/vision.py:21-23), and for ANIL the whole
// fast_adapt loop on features (core_functions/vision.py:9-17 with features != None).
#include <cooperative_groups.h>
#include "common.cuh"

namespace xm {

constexpr int HEAD_THREADS = 256;

// Feature accessor: X[i][d] over feat[rows][hw][c].  mode 0: d = c*hw + s (NCHW flatten);
// mode 1: d = c, X = mean over s.
struct Feat {
  const float* f;    // base of this task's rows
  int hw, c, mode, row0, row_step;
  __device__ __forceinline__ float at(int i, int d) const {
    const long long r = (long long)(row0 + i * row_step) * hw;
    if (mode == 0) {
      const int ch = d / hw, s = d - ch * hw;
      return __ldg(f + (r + s) * c + ch);
    }
    float acc = 0.f;
    for (int s = 0; s < hw; ++s) acc += __ldg(f + (r + s) * c + d);
    return acc / (float)hw;
  }
};

// Scatter a gradient w.r.t. X[i][d] back to the feature layout (+= when accumulate).
struct FeatGrad {
  float* f;
  int hw, c, mode, row0, row_step;
  __device__ __forceinline__ void put(int i, int d, float v, bool accumulate) const {
    const long long r = (long long)(row0 + i * row_step) * hw;
    if (mode == 0) {
      const int ch = d / hw, s = d - ch * hw;
      float* q = f + (r + s) * c + ch;
      *q = accumulate ? *q + v : v;
    } else {
      const float u = v / (float)hw;
      for (int s = 0; s < hw; ++s) {
        float* q = f + (r + s) * c + d;
        *q = accumulate ? *q + u : u;
      }
    }
  }
};

// out[i][w] = bias[w] + sum_d X(i,d)*W[w][d] (+ sum_d X2(i,d)*W2[w][d]); one warp per (i, w).
__device__ void logits_pass(const Feat& X, const float* W, const float* bias, const Feat* X2, const float* W2,
                            int n, int ways, int D, float* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int pr = warp; pr < n * ways; pr += nwarps) {
    const int i = pr / ways, w = pr - i * ways;
    float acc = 0.f;
    for (int d = lane; d < D; d += 32) {
      acc = fmaf(X.at(i, d), W[(long long)w * D + d], acc);
      if (X2) acc = fmaf(X2->at(i, d), W2[(long long)w * D + d], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) out[pr] = acc + (bias ? bias[w] : 0.f);
  }
}

// Row-wise softmax statistics.  prob <- softmax(logits); returns per-row loss and correctness through arrays.
__device__ void softmax_rows(const float* logits, const int64_t* labels, int lab0, int lab_step, int n, int ways,
                             float* prob, float* row_loss, int* row_ok) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float* z = logits + i * ways;
    float mx = z[0];
    int arg = 0;
    for (int w = 1; w < ways; ++w)
      if (z[w] > mx) { mx = z[w]; arg = w; }
    float se = 0.f;
    for (int w = 0; w < ways; ++w) se += expf(z[w] - mx);
    const float lse = mx + logf(se);
    const int y = (int)labels[lab0 + (long long)i * lab_step];
    for (int w = 0; w < ways; ++w) prob[i * ways + w] = expf(z[w] - lse);
    if (row_loss) row_loss[i] = lse - z[y];
    if (row_ok) row_ok[i] = (arg == y) ? 1 : 0;
  }
}

struct HeadK {
  int n, ways, c, hw, mode, dual, D;
  const float* feat; const float* feat_dot;
  const int64_t* labels; int label_row0, label_row_step, labels_per_task;
  const float* w; const float* b; long long wb_stride;
  const float* w_dot; const float* b_dot; long long wbdot_stride;
  float* loss; int* correct; float* logits;
  float* g_feat; float* g_feat_dot;
  float* out_w; float* out_b; long long out_stride;
  const float* base_w; const float* base_b; long long base_stride;
  float scale;
};

// head_kernel: one thread-block CLUSTER per task (G <= 8 CTAs).  CTA r owns a slice of the feature index range in the
// features' NATIVE order (e = s*c + ch for the NHWC flatten, e = ch for the spatial mean) and keeps ONLY that slice
// of the rows, the weights and their tangents in shared memory (one coalesced load, all requests in flight at once):
//   1. partial logits of the slice (and their tangent) -> own shared memory
//   2. cluster barrier; every CTA sums the G partials in rank order through distributed shared memory
//   3. softmax / cross-entropy / dL/dlogits (+ tangent through the softmax Hessian) redundantly per CTA (tiny)
//   4. parameter-gradient (axpy epilogue = fused SGD step) and feature-gradient of the slice from shared memory
// The kernel's latency is a few global round trips whatever the task count (it was ~100 serialized L2 round trips:
// 76-100 us per call, 1 ms of the config-2 step and the largest single item at 4 tasks per GPU).
namespace cg = cooperative_groups;

__global__ void __launch_bounds__(HEAD_THREADS) head_kernel(const HeadK k, const int G) {
  extern __shared__ float sm[];
  cg::cluster_group cluster = cg::this_cluster();
  const int task = blockIdx.x / G, part = blockIdx.x - task * G, tid = threadIdx.x;
  const int n = k.n, ways = k.ways, D = k.D, nw = n * ways;
  const int e_lo = (int)((long long)D * part / G), e_hi = (int)((long long)D * (part + 1) / G);
  const int span = e_hi - e_lo;
  const int span_max = (D + G - 1) / G + 1;     // same carve-up in every CTA of the cluster (DSMEM offsets must agree)
  const int ld = span_max | 1;                   // odd row stride: conflict-free column walks
  const bool dual = k.dual != 0, have_xd = dual && k.feat_dot;
  float* part_l = sm;                        // [nw]   partial logits of this slice          (read by the peers)
  float* part_d = part_l + nw;               // [nw]   partial tangent logits                (read by the peers)
  float* logit = part_d + nw;                // [nw]
  float* ldot = logit + nw;                  // [nw]
  float* prob = ldot + nw;                   // [nw]
  float* gl = prob + nw;                     // [nw]   dL/dlogits
  float* gld = gl + nw;                      // [nw]   its tangent
  float* row_loss = gld + nw;                // [n]
  int* row_ok = reinterpret_cast<int*>(row_loss + n);   // [n]
  float* Xs = row_loss + 2 * n;              // [n][ld]
  float* Xd = Xs + (size_t)n * ld;           // [n][ld]      (have_xd)
  float* Wn = Xd + (have_xd ? (size_t)n * ld : 0);       // [ways][ld]
  float* Wdn = Wn + (size_t)ways * ld;       // [ways][ld]   (dual)

  const long long fbase = (long long)task * n * k.hw * k.c;
  const float* F = k.feat + fbase;
  const float* Fd = have_xd ? k.feat_dot + fbase : nullptr;
  const float* W = k.w + (long long)task * k.wb_stride;
  const float* B = k.b + (long long)task * k.wb_stride;
  const float* Wd = dual ? k.w_dot + (long long)task * k.wbdot_stride : nullptr;
  const float* Bd = dual ? k.b_dot + (long long)task * k.wbdot_stride : nullptr;
  const int64_t* lab = k.labels + (long long)task * k.labels_per_task;
  auto perm = [&](int e) { return k.mode == 0 ? (e % k.c) * k.hw + e / k.c : e; };   // native -> PyTorch feature index

  // ---- stage the slice ------------------------------------------------------------------------------------------
  for (int idx = tid; idx < n * span; idx += HEAD_THREADS) {
    const int i = idx / span, j = idx - i * span, e = e_lo + j;
    float x, xd = 0.f;
    if (k.mode == 0) {
      x = __ldg(F + (long long)i * D + e);
      if (have_xd) xd = __ldg(Fd + (long long)i * D + e);
    } else {
      x = 0.f;
      for (int s = 0; s < k.hw; ++s) {
        x += __ldg(F + ((long long)i * k.hw + s) * k.c + e);
        if (have_xd) xd += __ldg(Fd + ((long long)i * k.hw + s) * k.c + e);
      }
      x /= (float)k.hw;
      xd /= (float)k.hw;
    }
    Xs[i * ld + j] = x;
    if (have_xd) Xd[i * ld + j] = xd;
  }
  for (int idx = tid; idx < ways * span; idx += HEAD_THREADS) {
    const int w = idx / span, j = idx - w * span;
    const long long src = (long long)w * D + perm(e_lo + j);
    Wn[w * ld + j] = __ldg(W + src);
    if (dual) Wdn[w * ld + j] = __ldg(Wd + src);
  }
  __syncthreads();

  // ---- partial logits of the slice: thread = (row, class) ---------------------------------------------------------
  for (int o = tid; o < nw; o += HEAD_THREADS) {
    const int i = o / ways, w = o - i * ways;
    const float* xr = Xs + i * ld;
    const float* wr = Wn + w * ld;
    float acc = 0.f, accd = 0.f;
    if (!dual) {
      for (int j = 0; j < span; ++j) acc = fmaf(xr[j], wr[j], acc);
    } else {
      const float* wdr = Wdn + w * ld;
      const float* xdr = Xd + i * ld;
      for (int j = 0; j < span; ++j) {
        acc = fmaf(xr[j], wr[j], acc);
        accd = fmaf(xr[j], wdr[j], accd);
        if (have_xd) accd = fmaf(xdr[j], wr[j], accd);
      }
    }
    part_l[o] = acc;
    part_d[o] = accd;
  }
  cluster.sync();
  for (int o = tid; o < nw; o += HEAD_THREADS) {
    const int w = o % ways;
    float acc = 0.f, accd = 0.f;
    for (int r = 0; r < G; ++r) {                       // rank order: every CTA of the cluster gets the same bits
      const float* peer = cluster.map_shared_rank(part_l, r);
      acc += peer[o];
      if (dual) accd += peer[nw + o];
    }
    logit[o] = acc + B[w];
    if (dual) ldot[o] = accd + Bd[w];
  }
  cluster.sync();                                       // peers have read this CTA's partials; logits visible below

  // ---- softmax, loss, accuracy, dL/dlogits (+ tangent) ----------------------------------------------------------
  for (int i = tid; i < n; i += HEAD_THREADS) {
    const float* z = logit + i * ways;
    float mx = z[0];
    int arg = 0;
    for (int w = 1; w < ways; ++w)
      if (z[w] > mx) { mx = z[w]; arg = w; }
    float se = 0.f;
    for (int w = 0; w < ways; ++w) se += expf(z[w] - mx);
    const float lse = mx + logf(se);
    int y = (int)lab[k.label_row0 + (long long)i * k.label_row_step];
    y = y < 0 ? 0 : (y >= ways ? ways - 1 : y);         // out-of-range labels are clamped (never index out of bounds)
    float pd = 0.f;
    for (int w = 0; w < ways; ++w) {
      const float p = expf(z[w] - lse);
      prob[i * ways + w] = p;
      if (dual) pd += p * ldot[i * ways + w];
    }
    for (int w = 0; w < ways; ++w) {
      const float p = prob[i * ways + w];
      gl[i * ways + w] = (p - (w == y ? 1.f : 0.f)) / (float)n;
      if (dual) gld[i * ways + w] = p * (ldot[i * ways + w] - pd) / (float)n;
    }
    row_loss[i] = lse - z[y];
    row_ok[i] = (arg == y) ? 1 : 0;
  }
  __syncthreads();
  if (part == 0) {
    if (tid == 0) {
      float s = 0.f;
      int ok = 0;
      for (int i = 0; i < n; ++i) { s += row_loss[i]; ok += row_ok[i]; }
      if (k.loss) k.loss[task] = s / (float)n;
      if (k.correct) k.correct[task] = ok;
    }
    if (k.logits)
      for (int i = tid; i < nw; i += HEAD_THREADS) k.logits[(long long)task * nw + i] = logit[i];
  }

  const float* Gq = dual ? gld : gl;
  // ---- parameter gradients of the slice through the axpy epilogue: thread = (class, feature) ----------------------
  if (k.out_w) {
    float* OW = k.out_w + (long long)task * k.out_stride;
    float* OB = k.out_b + (long long)task * k.out_stride;
    const float* BW = k.base_w ? k.base_w + (long long)task * k.base_stride : nullptr;
    const float* BB = k.base_b ? k.base_b + (long long)task * k.base_stride : nullptr;
    for (int idx = tid; idx < ways * span; idx += HEAD_THREADS) {
      const int w = idx / span, j = idx - w * span;
      float acc = 0.f;
      for (int i = 0; i < n; ++i) {
        acc = fmaf(Gq[i * ways + w], Xs[i * ld + j], acc);
        if (have_xd) acc = fmaf(gl[i * ways + w], Xd[i * ld + j], acc);
      }
      const long long dst = (long long)w * D + perm(e_lo + j);
      OW[dst] = (BW ? BW[dst] : 0.f) + k.scale * acc;
    }
    if (part == 0)
      for (int w = tid; w < ways; w += HEAD_THREADS) {
        float acc = 0.f;
        for (int i = 0; i < n; ++i) acc += Gq[i * ways + w];
        OB[w] = (BB ? BB[w] : 0.f) + k.scale * acc;
      }
  }
  // ---- feature gradient of the slice (native order: coalesced stores) -------------------------------------------
  float* GF = dual ? k.g_feat_dot : k.g_feat;
  if (GF) {
    float* out = GF + fbase;
    for (int idx = tid; idx < n * span; idx += HEAD_THREADS) {
      const int i = idx / span, j = idx - i * span, e = e_lo + j;
      float acc = 0.f;
      for (int w = 0; w < ways; ++w) {
        if (dual) {
          acc = fmaf(gld[i * ways + w], Wn[w * ld + j], acc);
          acc = fmaf(gl[i * ways + w], Wdn[w * ld + j], acc);
        } else {
          acc = fmaf(gl[i * ways + w], Wn[w * ld + j], acc);
        }
      }
      if (k.mode == 0) {
        out[(long long)i * D + e] = acc;
      } else {
        const float u = acc / (float)k.hw;
        for (int s = 0; s < k.hw; ++s) out[((long long)i * k.hw + s) * k.c + e] = u;
      }
    }
  }
}

// ---- ANIL ------------------------------------------------------------------------------------------
struct AnilK {
  int rows, ways, c, hw, mode, steps, first_order, D, S;
  float lr;
  const float* feat; const int64_t* labels;
  const float* w; const float* b;
  float* loss; int* correct; float* g_feat;
  float* g_w; float* g_b; long long g_stride;
  float* scratch; long long scratch_per_task;
};

__host__ __device__ inline long long anil_scratch_floats(int steps, int ways, int D, int S) {
  const long long ph = (long long)ways * D + ways;
  return (steps + 1) * ph          // fast weights W_0..W_T, b_0..b_T
         + ph                      // Wb, bb (outer cotangent of the head)
         + 2LL * steps * S * ways  // prob and dL/dlogits of every inner step
         + 4LL * S * ways + 2LL * S;   // work arrays
}

__global__ void __launch_bounds__(HEAD_THREADS) anil_head_kernel(const AnilK k) {
  const int task = blockIdx.x, tid = threadIdx.x;
  const int ways = k.ways, D = k.D, S = k.S, T = k.steps;
  const long long ph = (long long)ways * D + ways;
  float* base = k.scratch + (long long)task * k.scratch_per_task;
  float* Wt = base;                          // [(T+1)][ph]: W then b
  float* Wb = Wt + (T + 1) * ph;             // [ph]
  float* ps = Wb + ph;                       // [T][S][ways]
  float* gls = ps + (long long)T * S * ways; // [T][S][ways]
  float* wk0 = gls + (long long)T * S * ways;   // logits / cot   [S][ways]
  float* wk1 = wk0 + S * ways;               // prob              [S][ways]
  float* wk2 = wk1 + S * ways;               // glq / dl          [S][ways]
  float* wk3 = wk2 + S * ways;               // spare             [S][ways]
  float* row_loss = wk3 + S * ways;          // [S]
  int* row_ok = reinterpret_cast<int*>(row_loss + S);

  const long long fbase = (long long)task * k.rows * k.hw * k.c;
  const Feat Fs{k.feat + fbase, k.hw, k.c, k.mode, 0, 2};
  const Feat Fq{k.feat + fbase, k.hw, k.c, k.mode, 1, 2};
  const FeatGrad Gs{k.g_feat + fbase, k.hw, k.c, k.mode, 0, 2};
  const FeatGrad Gq{k.g_feat + fbase, k.hw, k.c, k.mode, 1, 2};
  const int64_t* lab = k.labels + (long long)task * k.rows;

  for (int i = tid; i < ways * D; i += blockDim.x) Wt[i] = k.w[i];
  for (int i = tid; i < ways; i += blockDim.x) Wt[ways * D + i] = k.b[i];
  __syncthreads();

  // ---- inner loop on the support rows ---------------------------------------------------------------
  for (int t = 0; t < T; ++t) {
    const float* W = Wt + t * ph;
    const float* B = W + ways * D;
    float* Wn = Wt + (t + 1) * ph;
    float* P = ps + (long long)t * S * ways;
    float* G = gls + (long long)t * S * ways;
    logits_pass(Fs, W, B, nullptr, nullptr, S, ways, D, wk0);
    __syncthreads();
    softmax_rows(wk0, lab, 0, 2, S, ways, P, nullptr, nullptr);
    __syncthreads();
    for (int idx = tid; idx < S * ways; idx += blockDim.x) {
      const int i = idx / ways, w = idx - i * ways;
      const int y = (int)lab[2 * i];
      G[idx] = (P[idx] - (w == y ? 1.f : 0.f)) / (float)S;
    }
    __syncthreads();
    for (int idx = tid; idx < ways * D; idx += blockDim.x) {
      const int w = idx / D, d = idx - w * D;
      float acc = 0.f;
      for (int i = 0; i < S; ++i) acc = fmaf(G[i * ways + w], Fs.at(i, d), acc);
      Wn[idx] = W[idx] + (-k.lr * acc);
    }
    for (int w = tid; w < ways; w += blockDim.x) {
      float acc = 0.f;
      for (int i = 0; i < S; ++i) acc += G[i * ways + w];
      Wn[ways * D + w] = B[w] + (-k.lr * acc);
    }
    __syncthreads();
  }

  // ---- query loss / accuracy at the adapted head ----------------------------------------------------
  const float* WT = Wt + T * ph;
  const float* BT = WT + ways * D;
  logits_pass(Fq, WT, BT, nullptr, nullptr, S, ways, D, wk0);
  __syncthreads();
  softmax_rows(wk0, lab, 1, 2, S, ways, wk1, row_loss, row_ok);
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    int ok = 0;
    for (int i = 0; i < S; ++i) { s += row_loss[i]; ok += row_ok[i]; }
    k.loss[task] = s / (float)S;
    k.correct[task] = ok;
  }
  for (int idx = tid; idx < S * ways; idx += blockDim.x) {
    const int i = idx / ways, w = idx - i * ways;
    const int y = (int)lab[2 * i + 1];
    wk2[idx] = (wk1[idx] - (w == y ? 1.f : 0.f)) / (float)S;
  }
  __syncthreads();
  // Wb = glq^T Fq, bb = sum glq ; gFq = glq W_T ; gFs = 0
  for (int idx = tid; idx < ways * D; idx += blockDim.x) {
    const int w = idx / D, d = idx - w * D;
    float acc = 0.f;
    for (int i = 0; i < S; ++i) acc = fmaf(wk2[i * ways + w], Fq.at(i, d), acc);
    Wb[idx] = acc;
  }
  for (int w = tid; w < ways; w += blockDim.x) {
    float acc = 0.f;
    for (int i = 0; i < S; ++i) acc += wk2[i * ways + w];
    Wb[ways * D + w] = acc;
  }
  for (int idx = tid; idx < S * D; idx += blockDim.x) {
    const int i = idx / D, d = idx - i * D;
    float acc = 0.f;
    for (int w = 0; w < ways; ++w) acc = fmaf(wk2[i * ways + w], WT[(long long)w * D + d], acc);
    Gq.put(i, d, acc, false);
    Gs.put(i, d, 0.f, false);
  }
  __syncthreads();

  // ---- second-order reverse sweep through the inner steps --------------------------------------------
  if (!k.first_order) {
    for (int t = T - 1; t >= 0; --t) {
      const float* W = Wt + t * ph;
      const float* P = ps + (long long)t * S * ways;
      const float* G = gls + (long long)t * S * ways;
      // cot[i][w] = Fs(i,:) . uW[w,:] + ub[w],  uW = -lr*Wb, ub = -lr*bb
      logits_pass(Fs, Wb, Wb + ways * D, nullptr, nullptr, S, ways, D, wk0);
      __syncthreads();
      for (int i = tid; i < S; i += blockDim.x) {
        float pc = 0.f;
        for (int w = 0; w < ways; ++w) pc += P[i * ways + w] * (-k.lr * wk0[i * ways + w]);
        for (int w = 0; w < ways; ++w)
          wk2[i * ways + w] = P[i * ways + w] * (-k.lr * wk0[i * ways + w] - pc) / (float)S;     // dl
      }
      __syncthreads();
      // gFs += gl_t uW + dl W_t       (uses Wb before its update)
      for (int idx = tid; idx < S * D; idx += blockDim.x) {
        const int i = idx / D, d = idx - i * D;
        float acc = 0.f;
        for (int w = 0; w < ways; ++w) {
          acc = fmaf(G[i * ways + w], -k.lr * Wb[(long long)w * D + d], acc);
          acc = fmaf(wk2[i * ways + w], W[(long long)w * D + d], acc);
        }
        Gs.put(i, d, acc, true);
      }
      __syncthreads();
      // Wb += dl^T Fs ; bb += sum dl
      for (int idx = tid; idx < ways * D; idx += blockDim.x) {
        const int w = idx / D, d = idx - w * D;
        float acc = 0.f;
        for (int i = 0; i < S; ++i) acc = fmaf(wk2[i * ways + w], Fs.at(i, d), acc);
        Wb[idx] += acc;
      }
      for (int w = tid; w < ways; w += blockDim.x) {
        float acc = 0.f;
        for (int i = 0; i < S; ++i) acc += wk2[i * ways + w];
        Wb[ways * D + w] += acc;
      }
      __syncthreads();
    }
  }
  for (int idx = tid; idx < ways * D; idx += blockDim.x) k.g_w[(long long)task * k.g_stride + idx] = Wb[idx];
  for (int w = tid; w < ways; w += blockDim.x) k.g_b[(long long)task * k.g_stride + w] = Wb[ways * D + w];
}

}  // namespace xm

using namespace xm;

extern "C" int xm_head(const XmHeadArgs* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  XM_REQUIRE(a != nullptr, "xm_head: null args");
  XM_REQUIRE(a->tasks > 0 && a->n > 0 && a->ways > 0 && a->c > 0 && a->hw > 0, "xm_head: bad sizes");
  XM_REQUIRE(a->mode == 0 || a->mode == 1, "xm_head: bad mode");
  XM_REQUIRE(a->feat && a->labels && a->w && a->b, "xm_head: null feat/labels/w/b");
  XM_REQUIRE(a->label_row_step > 0 && a->label_row0 >= 0 &&
             a->label_row0 + (long long)(a->n - 1) * a->label_row_step < a->labels_per_task,
             "xm_head: bad label row selection");
  XM_REQUIRE(!a->dual || (a->w_dot && a->b_dot), "xm_head: dual without w_dot/b_dot");
  XM_REQUIRE((a->out_w == nullptr) == (a->out_b == nullptr), "xm_head: out_w/out_b must both be given");
  HeadK k{};
  k.n = a->n; k.ways = a->ways; k.c = a->c; k.hw = a->hw; k.mode = a->mode; k.dual = a->dual;
  k.D = a->mode == 0 ? a->c * a->hw : a->c;
  k.feat = a->feat; k.feat_dot = a->feat_dot;
  k.labels = a->labels; k.label_row0 = a->label_row0; k.label_row_step = a->label_row_step;
  k.labels_per_task = a->labels_per_task;
  k.w = a->w; k.b = a->b; k.wb_stride = a->wb_task_stride;
  k.w_dot = a->w_dot; k.b_dot = a->b_dot; k.wbdot_stride = a->wbdot_task_stride;
  k.loss = a->loss; k.correct = a->correct; k.logits = a->logits;
  k.g_feat = a->g_feat; k.g_feat_dot = a->g_feat_dot;
  k.out_w = a->out_w; k.out_b = a->out_b; k.out_stride = a->out_task_stride;
  k.base_w = a->base_w; k.base_b = a->base_b; k.base_stride = a->base_task_stride;
  k.scale = a->scale;
  // one cluster per task: G CTAs, each owning ~D/G features (at least 16 each, at most 8 CTAs)
  int G = k.D / 16;
  if (G > 8) G = 8;
  if (G < 1) G = 1;
  const int span_max = (k.D + G - 1) / G + 1, ld = span_max | 1;
  const bool have_xd = a->dual && a->feat_dot;
  const size_t smem = ((size_t)7 * a->n * a->ways + 2 * a->n + (size_t)(have_xd ? 2 : 1) * a->n * ld +
                       (size_t)(a->dual ? 2 : 1) * a->ways * ld) * 4;
  XM_REQUIRE(smem <= 200 * 1024, "xm_head: n * (ways + D/8) too large for shared memory");
  XM_CUDA(cudaFuncSetAttribute(head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(a->tasks * G));
  cfg.blockDim = dim3(HEAD_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = G; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  XM_CUDA(cudaLaunchKernelEx(&cfg, head_kernel, k, G));
  return launched("xm_head");
}

extern "C" int64_t xm_anil_head_scratch_bytes(const XmAnilHeadArgs* a) {
  if (!a || a->tasks <= 0 || a->rows <= 0 || a->ways <= 0 || a->steps < 0) return -1;
  const int D = a->mode == 0 ? a->c * a->hw : a->c;
  return (int64_t)a->tasks * anil_scratch_floats(a->steps, a->ways, D, a->rows / 2) * 4;
}

extern "C" int xm_anil_head(const XmAnilHeadArgs* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  XM_REQUIRE(a != nullptr, "xm_anil_head: null args");
  XM_REQUIRE(a->tasks > 0 && a->rows > 0 && a->rows % 2 == 0 && a->ways > 0 && a->c > 0 && a->hw > 0 && a->steps >= 0,
             "xm_anil_head: bad sizes");
  XM_REQUIRE(a->mode == 0 || a->mode == 1, "xm_anil_head: bad mode");
  XM_REQUIRE(a->feat && a->labels && a->w && a->b && a->loss && a->correct && a->g_feat && a->g_w && a->g_b && a->scratch,
             "xm_anil_head: null pointer argument");
  AnilK k{};
  k.rows = a->rows; k.ways = a->ways; k.c = a->c; k.hw = a->hw; k.mode = a->mode; k.steps = a->steps;
  k.first_order = a->first_order; k.lr = a->lr;
  k.D = a->mode == 0 ? a->c * a->hw : a->c;
  k.S = a->rows / 2;
  k.feat = a->feat; k.labels = a->labels; k.w = a->w; k.b = a->b;
  k.loss = a->loss; k.correct = a->correct; k.g_feat = a->g_feat;
  k.g_w = a->g_w; k.g_b = a->g_b; k.g_stride = a->g_task_stride;
  k.scratch = a->scratch;
  k.scratch_per_task = anil_scratch_floats(k.steps, k.ways, k.D, k.S);
  XM_REQUIRE(a->scratch_bytes >= (int64_t)a->tasks * k.scratch_per_task * 4, "xm_anil_head: scratch too small");
  anil_head_kernel<<<a->tasks, HEAD_THREADS, 0, stream>>>(k);
  return launched("xm_anil_head");
}
