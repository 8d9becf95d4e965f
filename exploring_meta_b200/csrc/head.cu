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
};

// anil_head_kernel: ONE THREAD-BLOCK CLUSTER PER TASK (G <= 8 CTAs), everything resident in shared memory.  CTA r owns
// a slice of the feature index range (native NHWC order) and keeps only that slice of the support / query rows, of
// every fast-weight version W_0..W_T, of the outer cotangent Wb and of the support-feature gradient; the [S][ways]
// logit-type reductions over the full feature range go through distributed shared memory (partials summed in rank
// order by every CTA), the softmax-sized work is replicated.  Phases: T inner steps (fused SGD update of the slice),
// query loss / accuracy / gradient, the second-order reverse sweep (softmax Hessian term included), one write of the
// gradients.  Round 1 ran one 256-thread CTA per task out of global scratch: 1.27 ms per 32-task call (10.5 % of the
// config-3 step).
__global__ void __launch_bounds__(HEAD_THREADS) anil_head_kernel(const AnilK k, const int G) {
  extern __shared__ float sm[];
  cg::cluster_group cluster = cg::this_cluster();
  const int task = blockIdx.x / G, part = blockIdx.x - task * G, tid = threadIdx.x;
  const int ways = k.ways, D = k.D, S = k.S, T = k.steps, sw = S * ways;
  const int e_lo = (int)((long long)D * part / G), e_hi = (int)((long long)D * (part + 1) / G), span = e_hi - e_lo;
  const int ld = ((D + G - 1) / G + 1) | 1;
  // ---- shared-memory carve-up (identical in every CTA of the cluster: DSMEM offsets must agree) ----------------
  float* partl = sm;                                   // [S][ways] partial logits of this slice (read by the peers)
  float* wk0 = partl + sw;                             // [S][ways] logits / cotangent
  float* wk1 = wk0 + sw;                               // [S][ways] query probabilities
  float* wk2 = wk1 + sw;                               // [S][ways] glq / dl
  float* ps = wk2 + sw;                                // [T][S][ways] softmax of every inner step
  float* gls = ps + (size_t)T * sw;                    // [T][S][ways] dL/dlogits of every inner step
  float* bt = gls + (size_t)T * sw;                    // [T+1][ways] biases of every version (replicated)
  float* bb = bt + (T + 1) * ways;                     // [ways] outer cotangent of the bias
  float* row_loss = bb + ways;                         // [S]
  int* row_ok = reinterpret_cast<int*>(row_loss + S);  // [S]
  float* Xs = row_loss + 2 * S;                        // [S][ld] support rows, this slice
  float* Xq = Xs + (size_t)S * ld;                     // [S][ld] query rows
  float* GFs = Xq + (size_t)S * ld;                    // [S][ld] gradient w.r.t. the support rows
  float* Wt = GFs + (size_t)S * ld;                    // [T+1][ways][ld] fast weights W_0 .. W_T
  float* Wb = Wt + (size_t)(T + 1) * ways * ld;        // [ways][ld] outer cotangent of the weights

  const long long fbase = (long long)task * k.rows * k.hw * k.c;
  const float* F = k.feat + fbase;
  const int64_t* lab = k.labels + (long long)task * k.rows;
  auto perm = [&](int e) { return k.mode == 0 ? (e % k.c) * k.hw + e / k.c : e; };   // native -> PyTorch feature index
  auto label = [&](int row) { const int y = (int)lab[row]; return y < 0 ? 0 : (y >= ways ? ways - 1 : y); };

  // ---- stage: rows 2i = support, 2i+1 = query (prepare_batch's split), weights gathered into native order ---------
  for (int idx = tid; idx < 2 * S * span; idx += HEAD_THREADS) {
    const int r = idx / span, j = idx - r * span, e = e_lo + j;
    float x;
    if (k.mode == 0) {
      x = __ldg(F + (long long)r * D + e);
    } else {
      x = 0.f;
      for (int s = 0; s < k.hw; ++s) x += __ldg(F + ((long long)r * k.hw + s) * k.c + e);
      x /= (float)k.hw;
    }
    ((r & 1) ? Xq : Xs)[(r >> 1) * ld + j] = x;
  }
  for (int idx = tid; idx < ways * span; idx += HEAD_THREADS) {
    const int w = idx / span, j = idx - w * span;
    Wt[w * ld + j] = __ldg(k.w + (long long)w * D + perm(e_lo + j));
  }
  for (int idx = tid; idx < S * span; idx += HEAD_THREADS) GFs[(idx / span) * ld + idx % span] = 0.f;
  for (int w = tid; w < ways; w += HEAD_THREADS) bt[w] = __ldg(k.b + w);
  __syncthreads();

  // full[i][w] = bias[w] + sum over the whole feature range of X[i][.] * W[w][.]: slice partials through DSMEM
  auto logits_all = [&](const float* X, const float* W, const float* bias, float* out) {
    for (int o = tid; o < sw; o += HEAD_THREADS) {
      const int i = o / ways, w = o - i * ways;
      const float* xr = X + i * ld;
      const float* wr = W + w * ld;
      float acc = 0.f;
      for (int j = 0; j < span; ++j) acc = fmaf(xr[j], wr[j], acc);
      partl[o] = acc;
    }
    cluster.sync();
    for (int o = tid; o < sw; o += HEAD_THREADS) {
      float acc = 0.f;
      for (int r = 0; r < G; ++r) acc += cluster.map_shared_rank(partl, r)[o];       // rank order: same bits everywhere
      out[o] = acc + bias[o % ways];
    }
    cluster.sync();                                    // every peer has read this CTA's partials
  };
  // prob <- softmax(logits) per row; optional per-row loss / correctness against labels of rows lab0 + 2 i
  auto softmax_all = [&](const float* logits, int lab0, float* prob, bool want_loss) {
    for (int i = tid; i < S; i += HEAD_THREADS) {
      const float* z = logits + i * ways;
      float mx = z[0];
      int arg = 0;
      for (int w = 1; w < ways; ++w)
        if (z[w] > mx) { mx = z[w]; arg = w; }
      float se = 0.f;
      for (int w = 0; w < ways; ++w) se += expf(z[w] - mx);
      const float lse = mx + logf(se);
      for (int w = 0; w < ways; ++w) prob[i * ways + w] = expf(z[w] - lse);
      if (want_loss) {
        const int y = label(lab0 + 2 * i);
        row_loss[i] = lse - z[y];
        row_ok[i] = (arg == y) ? 1 : 0;
      }
    }
    __syncthreads();
  };

  // ---- inner loop on the support rows (core_functions/vision.py:9-13 with features != None) -------------------------
  for (int t = 0; t < T; ++t) {
    const float* W = Wt + (size_t)t * ways * ld;
    float* Wn = Wt + (size_t)(t + 1) * ways * ld;
    float* P = ps + (size_t)t * sw;
    float* Gt = gls + (size_t)t * sw;
    logits_all(Xs, W, bt + t * ways, wk0);
    softmax_all(wk0, 0, P, false);
    for (int idx = tid; idx < sw; idx += HEAD_THREADS) {
      const int i = idx / ways, w = idx - i * ways;
      Gt[idx] = (P[idx] - (w == label(2 * i) ? 1.f : 0.f)) / (float)S;
    }
    __syncthreads();
    for (int idx = tid; idx < ways * span; idx += HEAD_THREADS) {
      const int w = idx / span, j = idx - w * span;
      float acc = 0.f;
      for (int i = 0; i < S; ++i) acc = fmaf(Gt[i * ways + w], Xs[i * ld + j], acc);
      Wn[w * ld + j] = W[w * ld + j] + (-k.lr * acc);
    }
    for (int w = tid; w < ways; w += HEAD_THREADS) {
      float acc = 0.f;
      for (int i = 0; i < S; ++i) acc += Gt[i * ways + w];
      bt[(t + 1) * ways + w] = bt[t * ways + w] + (-k.lr * acc);
    }
    __syncthreads();
  }

  // ---- query loss / accuracy at the adapted head (vision.py:15-17) ------------------------------------------------------
  const float* WT = Wt + (size_t)T * ways * ld;
  logits_all(Xq, WT, bt + T * ways, wk0);
  softmax_all(wk0, 1, wk1, true);
  if (part == 0 && tid == 0) {
    float s = 0.f;
    int ok = 0;
    for (int i = 0; i < S; ++i) { s += row_loss[i]; ok += row_ok[i]; }
    k.loss[task] = s / (float)S;
    k.correct[task] = ok;
  }
  for (int idx = tid; idx < sw; idx += HEAD_THREADS) {
    const int i = idx / ways, w = idx - i * ways;
    wk2[idx] = (wk1[idx] - (w == label(2 * i + 1) ? 1.f : 0.f)) / (float)S;
  }
  __syncthreads();
  // Wb = glq^T Xq, bb = sum glq, gXq = glq W_T
  for (int idx = tid; idx < ways * span; idx += HEAD_THREADS) {
    const int w = idx / span, j = idx - w * span;
    float acc = 0.f;
    for (int i = 0; i < S; ++i) acc = fmaf(wk2[i * ways + w], Xq[i * ld + j], acc);
    Wb[w * ld + j] = acc;
  }
  for (int w = tid; w < ways; w += HEAD_THREADS) {
    float acc = 0.f;
    for (int i = 0; i < S; ++i) acc += wk2[i * ways + w];
    bb[w] = acc;
  }
  float* GF = k.g_feat + fbase;
  auto put_row = [&](int row, int j, float v) {         // gradient w.r.t. feature e_lo + j of row `row`
    const int e = e_lo + j;
    if (k.mode == 0) {
      GF[(long long)row * D + e] = v;
    } else {
      const float u = v / (float)k.hw;
      for (int s = 0; s < k.hw; ++s) GF[((long long)row * k.hw + s) * k.c + e] = u;
    }
  };
  for (int idx = tid; idx < S * span; idx += HEAD_THREADS) {
    const int i = idx / span, j = idx - i * span;
    float acc = 0.f;
    for (int w = 0; w < ways; ++w) acc = fmaf(wk2[i * ways + w], WT[w * ld + j], acc);
    put_row(2 * i + 1, j, acc);
  }
  __syncthreads();

  // ---- second-order reverse sweep through the inner steps ------------------------------------------------------------------
  if (!k.first_order) {
    for (int t = T - 1; t >= 0; --t) {
      const float* W = Wt + (size_t)t * ways * ld;
      const float* P = ps + (size_t)t * sw;
      const float* Gt = gls + (size_t)t * sw;
      logits_all(Xs, Wb, bb, wk0);                     // cot[i][w] = Xs(i,:) . Wb[w,:] + bb[w]
      for (int i = tid; i < S; i += HEAD_THREADS) {
        float pc = 0.f;
        for (int w = 0; w < ways; ++w) pc += P[i * ways + w] * (-k.lr * wk0[i * ways + w]);
        for (int w = 0; w < ways; ++w)
          wk2[i * ways + w] = P[i * ways + w] * (-k.lr * wk0[i * ways + w] - pc) / (float)S;     // dl
      }
      __syncthreads();
      for (int idx = tid; idx < S * span; idx += HEAD_THREADS) {       // gXs += gl_t uW + dl W_t  (Wb before its update)
        const int i = idx / span, j = idx - i * span;
        float acc = 0.f;
        for (int w = 0; w < ways; ++w) {
          acc = fmaf(Gt[i * ways + w], -k.lr * Wb[w * ld + j], acc);
          acc = fmaf(wk2[i * ways + w], W[w * ld + j], acc);
        }
        GFs[i * ld + j] += acc;
      }
      __syncthreads();
      for (int idx = tid; idx < ways * span; idx += HEAD_THREADS) {    // Wb += dl^T Xs ; bb += sum dl
        const int w = idx / span, j = idx - w * span;
        float acc = 0.f;
        for (int i = 0; i < S; ++i) acc = fmaf(wk2[i * ways + w], Xs[i * ld + j], acc);
        Wb[w * ld + j] += acc;
      }
      for (int w = tid; w < ways; w += HEAD_THREADS) {
        float acc = 0.f;
        for (int i = 0; i < S; ++i) acc += wk2[i * ways + w];
        bb[w] += acc;
      }
      __syncthreads();
    }
  }
  for (int idx = tid; idx < S * span; idx += HEAD_THREADS) put_row(2 * (idx / span), idx % span, GFs[(idx / span) * ld + idx % span]);
  for (int idx = tid; idx < ways * span; idx += HEAD_THREADS) {
    const int w = idx / span, j = idx - w * span;
    k.g_w[(long long)task * k.g_stride + (long long)w * D + perm(e_lo + j)] = Wb[w * ld + j];
  }
  if (part == 0)
    for (int w = tid; w < ways; w += HEAD_THREADS) k.g_b[(long long)task * k.g_stride + w] = bb[w];
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
  return 0;          // the kernel keeps every intermediate in shared memory; `scratch` may be NULL
}

extern "C" int xm_anil_head(const XmAnilHeadArgs* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  XM_REQUIRE(a != nullptr, "xm_anil_head: null args");
  XM_REQUIRE(a->tasks > 0 && a->rows > 0 && a->rows % 2 == 0 && a->ways > 0 && a->c > 0 && a->hw > 0 && a->steps >= 0,
             "xm_anil_head: bad sizes");
  XM_REQUIRE(a->mode == 0 || a->mode == 1, "xm_anil_head: bad mode");
  XM_REQUIRE(a->feat && a->labels && a->w && a->b && a->loss && a->correct && a->g_feat && a->g_w && a->g_b,
             "xm_anil_head: null pointer argument");
  AnilK k{};
  k.rows = a->rows; k.ways = a->ways; k.c = a->c; k.hw = a->hw; k.mode = a->mode; k.steps = a->steps;
  k.first_order = a->first_order; k.lr = a->lr;
  k.D = a->mode == 0 ? a->c * a->hw : a->c;
  k.S = a->rows / 2;
  k.feat = a->feat; k.labels = a->labels; k.w = a->w; k.b = a->b;
  k.loss = a->loss; k.correct = a->correct; k.g_feat = a->g_feat;
  k.g_w = a->g_w; k.g_b = a->g_b; k.g_stride = a->g_task_stride;
  int G = k.D / 32;
  if (G > 8) G = 8;
  if (G < 1) G = 1;
  const int ld = ((k.D + G - 1) / G + 1) | 1, sw = k.S * k.ways;
  const size_t smem = ((size_t)4 * sw + 2 * (size_t)k.steps * sw + (size_t)(k.steps + 2) * k.ways + 2 * k.S +
                       3 * (size_t)k.S * ld + (size_t)(k.steps + 2) * k.ways * ld) * 4;
  XM_REQUIRE(smem <= 200 * 1024, "xm_anil_head: rows * (ways + D/8) * steps too large for shared memory");
  XM_CUDA(cudaFuncSetAttribute(anil_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
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
  XM_CUDA(cudaLaunchKernelEx(&cfg, anil_head_kernel, k, G));

  return launched("xm_anil_head");
}
