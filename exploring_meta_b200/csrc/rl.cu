// rl.cu -- config 5: the MAML-TRPO policy-MLP path as task-batched kernels with per-task weights.
//
//   xm_rl_advantages : everything core_functions/rl.py:95-110 (compute_advantages) + cherry derive from a replay alone:
//                      discounted returns, LinearValue ridge fit / values, bootstraps, GAE, normalisation.
//   xm_rl_sweep      : one pass of every task's DiagNormalPolicy (core_functions/policies.py:30-56) over its replay:
//                      forward, tangent forward, backward, tangent backward of the 2-hidden-layer MLP fused with the
//                      Gaussian log-prob / surrogate / KL terms -- the kernels behind trpo_update (rl.py:361-374),
//                      meta_surrogate_loss (:441-473) and the gradient / Fisher-vector products of meta_optimize_trpo
//                      (:409-438).  Second order is forward-over-reverse, as in the vision path: the cotangent through
//                      theta' = theta - lr * g(theta) is v - lr * H v and H v is the TANGENT of the gradient sweep.
//
// One CTA processes tiles of 32 transitions of ONE task with that task's weights resident in shared memory (42 KB,
// plus the tangent weights); the three h x h contractions per tile (forward, data gradient, weight gradient) run on
// the fp32 CUDA cores -- exact fp32 products, 10 k parameters per task: the path is launch / latency bound, not a
// tensor-core workload.  Parameter gradients accumulate in registers across the CTA's tiles and are reduced over the
// task's CTAs in a fixed order (deterministic), through the axpy epilogue.
#include "common.cuh"

namespace xm {

// ====================================================================================================================
// advantages
// ====================================================================================================================
constexpr int ADV_THREADS = 256;
constexpr int ADV_MAX_F = 12;          // features 2*state_dim + 4, state_dim <= 4

__device__ __forceinline__ double block_sum(double v, double* red /* [ADV_THREADS/32] */) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
  return s;
}

// LinearValue features (cherry.models.robotics.LinearValue._features): [s, s^2, t, t^2, t^3, 1], t = index / 100
__device__ __forceinline__ double lv_feature(const float* s, int sd, int k, int i) {
  if (i < sd) return (double)s[k * sd + i];
  if (i < 2 * sd) { const double v = (double)s[k * sd + i - sd]; return v * v; }
  const double t = (double)k / 100.0;
  const int j = i - 2 * sd;
  return j == 0 ? t : (j == 1 ? t * t : (j == 2 ? t * t * t : 1.0));
}

__global__ void __launch_bounds__(ADV_THREADS) rl_adv_kernel(const XmRlAdvArgs a) {
  extern __shared__ double smd[];
  const int rep = blockIdx.x, n = a.n, sd = a.state_dim, tid = threadIdx.x;
  const int F = 2 * sd + 4, E = F * (F + 1) / 2 + F;
  double* r = smd;                              // [n] rewards, later the TD residuals
  double* ret = r + n;                          // [n] returns, later the advantages
  double* boot = ret + n;                       // [n] bootstrapped values
  int* ends = reinterpret_cast<int*>(boot + n); // [n] last index of every episode segment
  float* s = reinterpret_cast<float*>(ends + n);            // [n][sd]
  float* ns = s + (size_t)n * sd;                            // [n][sd]
  float* dn = ns + (size_t)n * sd;                           // [n] done flags
  __shared__ int s_cnt[ADV_THREADS];
  __shared__ int s_nseg;
  __shared__ double s_part[ADV_THREADS];
  __shared__ double s_gram[ADV_MAX_F * (ADV_MAX_F + 1) / 2 + ADV_MAX_F];
  __shared__ double s_coef[ADV_MAX_F];
  __shared__ double s_red[ADV_THREADS / 32];

  const long long base = (long long)rep * n;
  for (int k = tid; k < n; k += ADV_THREADS) {
    r[k] = (double)a.rewards[base + k];
    dn[k] = a.dones[base + k];
  }
  for (int k = tid; k < n * sd; k += ADV_THREADS) {
    s[k] = a.states[base * sd + k];
    ns[k] = a.next_states[base * sd + k];
  }
  __syncthreads();

  // ---- episode segments: a reverse scan restarts after every done (R * (1 - done)) ------------------------------
  const int chunk = (n + ADV_THREADS - 1) / ADV_THREADS;
  const int k0 = min(n, tid * chunk), k1 = min(n, k0 + chunk);
  int cnt = 0;
  for (int k = k0; k < k1; ++k) cnt += (dn[k] != 0.f || k == n - 1) ? 1 : 0;
  s_cnt[tid] = cnt;
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int t = 0; t < ADV_THREADS; ++t) { const int c = s_cnt[t]; s_cnt[t] = run; run += c; }
    s_nseg = run;
  }
  __syncthreads();
  {
    int o = s_cnt[tid];
    for (int k = k0; k < k1; ++k)
      if (dn[k] != 0.f || k == n - 1) ends[o++] = k;
  }
  __syncthreads();
  const int nseg = s_nseg;

  // ---- ch.td.discount(gamma, rewards, dones): sequential inside a segment, one thread per segment ---------------
  for (int j = tid; j < nseg; j += ADV_THREADS) {
    const int start = j ? ends[j - 1] + 1 : 0, end = ends[j];
    double R = 0.0;
    for (int k = end; k >= start; --k) { R = r[k] + a.gamma * R; ret[k] = R; }
  }
  __syncthreads();
  if (a.returns)
    for (int k = tid; k < n; k += ADV_THREADS) a.returns[base + k] = (float)ret[k];

  // ---- LinearValue.fit: A = F^T F + reg I, b = F^T returns; coefficients = A^-1 b ----------------------------------
  {
    const int C = ADV_THREADS / E;                     // sample chunks per Gram entry
    double acc = 0.0;
    const int e = tid % E, c = tid / E;
    if (c < C) {
      int fi, fj;
      if (e < F * (F + 1) / 2) {                       // upper-triangular entry (fi <= fj)
        int idx = e;
        fi = 0;
        while (idx >= F - fi) { idx -= F - fi; ++fi; }
        fj = fi + idx;
      } else {
        fi = e - F * (F + 1) / 2;
        fj = -1;                                       // right-hand side entry
      }
      const int len = (n + C - 1) / C;
      const int a0 = min(n, c * len), a1 = min(n, a0 + len);
      for (int k = a0; k < a1; ++k)
        acc += lv_feature(s, sd, k, fi) * (fj >= 0 ? lv_feature(s, sd, k, fj) : ret[k]);
    }
    s_part[tid] = acc;
    __syncthreads();
    if (tid < E) {
      double t = 0.0;
      for (int cc = 0; cc < C; ++cc) t += s_part[cc * E + tid];
      s_gram[tid] = t;
    }
    __syncthreads();
    if (tid == 0) {
      double A[ADV_MAX_F][ADV_MAX_F + 1];
      int idx = 0;
      for (int i = 0; i < F; ++i)
        for (int j = i; j < F; ++j) { A[i][j] = A[j][i] = s_gram[idx++]; }
      for (int i = 0; i < F; ++i) { A[i][i] += a.reg; A[i][F] = s_gram[idx++]; }
      for (int col = 0; col < F; ++col) {              // Gaussian elimination, partial pivoting
        int piv = col;
        for (int i = col + 1; i < F; ++i)
          if (fabs(A[i][col]) > fabs(A[piv][col])) piv = i;
        if (piv != col)
          for (int j = 0; j <= F; ++j) { const double t = A[col][j]; A[col][j] = A[piv][j]; A[piv][j] = t; }
        const double d = A[col][col];
        for (int i = col + 1; i < F; ++i) {
          const double f = A[i][col] / d;
          for (int j = col; j <= F; ++j) A[i][j] -= f * A[col][j];
        }
      }
      for (int i = F - 1; i >= 0; --i) {
        double t = A[i][F];
        for (int j = i + 1; j < F; ++j) t -= A[i][j] * s_coef[j];
        s_coef[i] = t / A[i][i];
      }
    }
    __syncthreads();
  }

  // ---- values / next values -> bootstraps (rl.py:101-103) ------------------------------------------------------------
  for (int k = tid; k < n; k += ADV_THREADS) {
    double v = 0.0, nv = 0.0;
    for (int i = 0; i < F; ++i) {
      v += lv_feature(s, sd, k, i) * s_coef[i];
      nv += lv_feature(ns, sd, k, i) * s_coef[i];
    }
    const double d = (double)dn[k];
    boot[k] = v * (1.0 - d) + nv * d;
  }
  __syncthreads();
  // ---- ch.pg.generalized_advantage: td = r + gamma (1 - d) next - value, then discount(tau * gamma, td, dones) ----
  for (int k = tid; k < n; k += ADV_THREADS) {
    const double next = (k + 1 < n) ? boot[k + 1] : 0.0;
    r[k] = r[k] + a.gamma * (1.0 - (double)dn[k]) * next - boot[k];
  }
  __syncthreads();
  const double tg = a.tau * a.gamma;
  for (int j = tid; j < nseg; j += ADV_THREADS) {
    const int start = j ? ends[j - 1] + 1 : 0, end = ends[j];
    double R = 0.0;
    for (int k = end; k >= start; --k) { R = r[k] + tg * R; ret[k] = R; }
  }
  __syncthreads();
  if (a.advantages)
    for (int k = tid; k < n; k += ADV_THREADS) a.advantages[base + k] = (float)ret[k];
  // ---- ch.normalize: (x - mean) / (unbiased std + 1e-8) ----------------------------------------------------------------
  double part = 0.0;
  for (int k = tid; k < n; k += ADV_THREADS) part += ret[k];
  const double mean = block_sum(part, s_red) / (double)n;
  part = 0.0;
  for (int k = tid; k < n; k += ADV_THREADS) { const double d = ret[k] - mean; part += d * d; }
  const double var = block_sum(part, s_red) / (double)(n > 1 ? n - 1 : 1);
  const double inv = 1.0 / (sqrt(var) + 1e-8);
  for (int k = tid; k < n; k += ADV_THREADS)
    a.coef[base + k] = (float)(a.coef_scale * (n > 1 ? (ret[k] - mean) * inv : ret[k]));
}

// ====================================================================================================================
// policy sweep
// ====================================================================================================================
constexpr int RL_THREADS = 256;
constexpr int RL_TS = 32;              // transitions per tile
constexpr int RL_MAXH = 128;           // hidden width limit
constexpr int RL_MAXIO = 8;            // state / action dimension limit
constexpr float LOG_EPS = -13.815510557964274f;     // log(1e-6), policies.py:14,51
constexpr float HALF_LOG_2PI = 0.9189385332046727f;

struct SweepK {
  int n, in, out, h1, h2, act, loss, what, G;
  const float* states; const float* actions; const float* coef; const float* mu_old; const float* logstd_old;
  float kl_scale, clip;
  int head_only;
  const float* theta; long long theta_stride;
  const float* theta_dot; long long dot_stride;
  float* mu_out;
  float* partial; double* partial_sc;
  int P;
};

__host__ __device__ inline int rl_num_params(int in, int out, int h1, int h2) {
  return out + h1 * in + h1 + h2 * h1 + h2 + out * h2 + out;
}

__device__ __forceinline__ float act_fn(float z, int act) { return act == XM_ACT_TANH ? tanhf(z) : fmaxf(z, 0.f); }
__device__ __forceinline__ float act_d1(float h, int act) { return act == XM_ACT_TANH ? 1.f - h * h : (h > 0.f ? 1.f : 0.f); }
// d/d eps of act'(z) expressed through h and hdot: tanh: -2 h hdot; relu: 0
__device__ __forceinline__ float act_d1_dot(float h, float hdot, int act) { return act == XM_ACT_TANH ? -2.f * h * hdot : 0.f; }

__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float comp(const float4& v, int r) { return r == 0 ? v.x : (r == 1 ? v.y : (r == 2 ? v.z : v.w)); }
// row stride of W2 in shared memory: a multiple of 4 floats (128-bit loads) whose quarter is odd, so that the eight
// lanes of a load phase, which read rows i, i+1, ..., hit eight different 16-byte bank groups
__host__ __device__ inline int rl_ldw(int h1) { int l = (h1 + 3) & ~3; if (((l >> 2) & 1) == 0) l += 4; return l; }

// DF: tangent forward (theta_dot given); DB: tangent backward too (Hessian-vector product)
template <bool DF, bool DB>
__global__ void __launch_bounds__(RL_THREADS) rl_sweep_kernel(const SweepK k) {
  extern __shared__ __align__(16) float sm[];
  const int task = blockIdx.y, g = blockIdx.x, tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int n = k.n, IN = k.in, OUT = k.out, H1 = k.h1, H2 = k.h2, act = k.act;
  const int ldw = rl_ldw(H1);
  // ---- shared-memory carve-up -------------------------------------------------------------------------------------
  float* p = sm;
  auto take = [&](int count) { float* q = p; p += (count + 3) & ~3; return q; };
  float* sig = take(OUT);  float* W1 = take(H1 * IN); float* b1 = take(H1);
  float* W2 = take(H2 * ldw); float* b2 = take(H2); float* W3 = take(OUT * H2); float* b3 = take(OUT);
  float* sigd = DF ? take(OUT) : nullptr; float* W1d = DF ? take(H1 * IN) : nullptr; float* b1d = DF ? take(H1) : nullptr;
  float* W2d = DF ? take(H2 * ldw) : nullptr; float* b2d = DF ? take(H2) : nullptr;
  float* W3d = DF ? take(OUT * H2) : nullptr; float* b3d = DF ? take(OUT) : nullptr;
  const int HM = H1 > H2 ? H1 : H2;
  float* A1 = take(RL_TS * H1); float* A2 = take(RL_TS * HM);              // h1, h2 (A2 later holds gz1)
  float* D1 = DF ? take(RL_TS * H1) : nullptr; float* D2 = DF ? take(RL_TS * HM) : nullptr;   // tangents (D2 later: gz1 dot)
  float* G2 = take(RL_TS * H2); float* G2d = DB ? take(RL_TS * H2) : nullptr;              // gz2 and its tangent
  float* X = take(RL_TS * IN); float* ACT = take(RL_TS * OUT); float* CF = take(RL_TS);
  float* MU = take(RL_TS * OUT); float* MUD = take(RL_TS * OUT); float* MUO = take(RL_TS * OUT);
  float* GMU = take(RL_TS * OUT); float* GMUD = take(RL_TS * OUT);
  float* lam = take(OUT); float* lamd = take(OUT); float* lamo = take(OUT);
  __shared__ double s_red[RL_THREADS / 32];

  // ---- this task's parameters (and tangent direction) -> shared memory ----------------------------------------------
  const float* th = k.theta + (long long)task * k.theta_stride;
  const float* thd = DF ? k.theta_dot + (long long)task * k.dot_stride : nullptr;
  const int oW1 = OUT, ob1 = oW1 + H1 * IN, oW2 = ob1 + H1, ob2 = oW2 + H2 * H1, oW3 = ob2 + H2, ob3 = oW3 + OUT * H2;
  for (int i = tid; i < OUT; i += RL_THREADS) { sig[i] = th[i]; b3[i] = th[ob3 + i]; if (DF) { sigd[i] = thd[i]; b3d[i] = thd[ob3 + i]; } }
  const bool body_dot = !(k.head_only & 2);             // ANIL cotangent: the body entries of the direction are masked out
  for (int i = tid; i < H1 * IN; i += RL_THREADS) { W1[i] = th[oW1 + i]; if (DF) W1d[i] = body_dot ? thd[oW1 + i] : 0.f; }
  for (int i = tid; i < H1; i += RL_THREADS) { b1[i] = th[ob1 + i]; if (DF) b1d[i] = body_dot ? thd[ob1 + i] : 0.f; }
  for (int i = tid; i < H2; i += RL_THREADS) { b2[i] = th[ob2 + i]; if (DF) b2d[i] = body_dot ? thd[ob2 + i] : 0.f; }
  for (int i = tid; i < OUT * H2; i += RL_THREADS) { W3[i] = th[oW3 + i]; if (DF) W3d[i] = thd[oW3 + i]; }
  for (int i = tid; i < H2 * H1; i += RL_THREADS) {
    const int r = i / H1, c = i - r * H1;
    W2[r * ldw + c] = th[oW2 + i];
    if (DF) W2d[r * ldw + c] = body_dot ? thd[oW2 + i] : 0.f;
  }
  __syncthreads();
  if (tid < OUT) {
    const bool on = sig[tid] >= LOG_EPS;                      // torch.clamp(min): gradient passes where sigma >= min
    lam[tid] = on ? sig[tid] : LOG_EPS;
    lamd[tid] = (DF && on) ? sigd[tid] : 0.f;
    lamo[tid] = k.logstd_old ? k.logstd_old[(long long)task * OUT + tid] : 0.f;
  }
  __syncthreads();

  const bool need_bwd = k.what != XM_RL_FORWARD;
  // ---- register accumulators of the parameter gradient (kept across this CTA's tiles) ----------------------------------
  float accW2[16][4];
#pragma unroll
  for (int q = 0; q < 16; ++q)
#pragma unroll
    for (int r = 0; r < 4; ++r) accW2[q][r] = 0.f;
  float accW3[RL_MAXIO], accW1[RL_MAXIO], accb2 = 0.f, accb1 = 0.f, accb3 = 0.f, accls[RL_MAXIO];
#pragma unroll
  for (int o = 0; o < RL_MAXIO; ++o) accW3[o] = accW1[o] = accls[o] = 0.f;
  double loss_acc = 0.0, kl_acc = 0.0;

  const int tiles = (n + RL_TS - 1) / RL_TS;
  const float invA = 1.f / (float)OUT;
  for (int tile = g; tile < tiles; tile += k.G) {
    const int s0 = tile * RL_TS;
    const long long rowbase = (long long)task * n + s0;
    // ---- stage the tile ---------------------------------------------------------------------------------------------
    for (int i = tid; i < RL_TS * IN; i += RL_THREADS) X[i] = (s0 + i / IN < n) ? k.states[rowbase * IN + i] : 0.f;
    for (int i = tid; i < RL_TS * OUT; i += RL_THREADS) {
      const bool ok = s0 + i / OUT < n;
      ACT[i] = (ok && k.actions) ? k.actions[rowbase * OUT + i] : 0.f;
      MUO[i] = (ok && k.mu_old) ? k.mu_old[rowbase * OUT + i] : 0.f;
    }
    for (int i = tid; i < RL_TS; i += RL_THREADS) CF[i] = (s0 + i < n && k.coef) ? k.coef[rowbase + i] : 0.f;
    __syncthreads();
    // ---- layer 1 ----------------------------------------------------------------------------------------------------
    for (int idx = tid; idx < RL_TS * H1; idx += RL_THREADS) {
      const int s = idx / H1, i = idx - s * H1;
      float z = b1[i], zd = DF ? b1d[i] : 0.f;
      for (int c = 0; c < IN; ++c) {
        z = fmaf(W1[i * IN + c], X[s * IN + c], z);
        if (DF) zd = fmaf(W1d[i * IN + c], X[s * IN + c], zd);
      }
      const float h = act_fn(z, act);
      A1[idx] = h;
      if (DF) D1[idx] = act_d1(h, act) * zd;
    }
    __syncthreads();
    // ---- layer 2: z2[s][i] = b2[i] + sum_j W2[i][j] h1[s][j]; thread = 4 samples x 4 units ------------------------
    {
      float acc[4][4], accd[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int q = 0; q < 4; ++q) { acc[a][q] = 0.f; accd[a][q] = 0.f; }
      const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int j = 0; j < H1; j += 4) {                // 128-bit shared loads: 8 (16 with tangents) per 64 (192) FMAs
        float4 av[4], dv[4], wv[4], wdv[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) { av[a] = lds4(A1 + (ty * 4 + a) * H1 + j); if (DF) dv[a] = lds4(D1 + (ty * 4 + a) * H1 + j); }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int i = tx + 32 * q;
          wv[q] = i < H2 ? lds4(W2 + i * ldw + j) : zero4;
          if (DF) wdv[q] = i < H2 ? lds4(W2d + i * ldw + j) : zero4;
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            acc[a][q] = fmaf(wv[q].x, av[a].x, fmaf(wv[q].y, av[a].y, fmaf(wv[q].z, av[a].z, fmaf(wv[q].w, av[a].w, acc[a][q]))));
            if (DF) {
              accd[a][q] = fmaf(wdv[q].x, av[a].x, fmaf(wdv[q].y, av[a].y, fmaf(wdv[q].z, av[a].z, fmaf(wdv[q].w, av[a].w, accd[a][q]))));
              accd[a][q] = fmaf(wv[q].x, dv[a].x, fmaf(wv[q].y, dv[a].y, fmaf(wv[q].z, dv[a].z, fmaf(wv[q].w, dv[a].w, accd[a][q]))));
            }
          }
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int i = tx + 32 * q, s = ty * 4 + a;
          if (i < H2) {
            const float h = act_fn(acc[a][q] + b2[i], act);
            A2[s * H2 + i] = h;
            if (DF) D2[s * H2 + i] = act_d1(h, act) * (accd[a][q] + b2d[i]);
          }
        }
    }
    __syncthreads();
    // ---- output layer: warp ty handles samples 4 ty .. 4 ty + 3 ---------------------------------------------------------
    for (int a = 0; a < 4; ++a) {
      const int s = ty * 4 + a;
      for (int o = 0; o < OUT; ++o) {
        float m = 0.f, md = 0.f;
        for (int j = tx; j < H2; j += 32) {
          m = fmaf(W3[o * H2 + j], A2[s * H2 + j], m);
          if (DF) md = fmaf(W3d[o * H2 + j], A2[s * H2 + j], fmaf(W3[o * H2 + j], D2[s * H2 + j], md));
        }
        m = warp_sum(m);
        if (DF) md = warp_sum(md);
        if (tx == 0) { MU[s * OUT + o] = m + b3[o]; MUD[s * OUT + o] = DF ? md + b3d[o] : 0.f; }
      }
    }
    __syncthreads();
    // ---- per-sample loss terms: thread = sample ---------------------------------------------------------------------------
    if (tid < RL_TS) {
      const int s = tid;
      const bool ok = s0 + s < n;
      const float c = CF[s];
      if (k.mu_out && ok)
        for (int d = 0; d < OUT; ++d) k.mu_out[(rowbase + s) * OUT + d] = MU[s * OUT + d];
      float lp = 0.f, lpo = 0.f;
      double klv = 0.0;
      for (int d = 0; d < OUT; ++d) {
        const float e = ACT[s * OUT + d] - MU[s * OUT + d], is2 = expf(-2.f * lam[d]);
        lp += -0.5f * e * e * is2 - lam[d] - HALF_LOG_2PI;
        if (k.loss == XM_RL_SURROGATE) {
          const float eo = ACT[s * OUT + d] - MUO[s * OUT + d], iso2 = expf(-2.f * lamo[d]);
          lpo += -0.5f * eo * eo * iso2 - lamo[d] - HALF_LOG_2PI;
          // KL(N(mu, s) || N(mu_o, s_o)) = -dl + (e^{2 dl} - 1) / 2 + dm^2 / (2 s_o^2), dl = log s - log s_o: in double
          // (the fp32 form cancels to ~1e-7 absolute against a KL of 1e-3 .. 1e-2)
          const double dm = (double)MU[s * OUT + d] - (double)MUO[s * OUT + d];
          const double dl = (double)lam[d] - (double)lamo[d];
          klv += 0.5 * expm1(2.0 * dl) - dl + 0.5 * dm * dm * exp(-2.0 * (double)lamo[d]);
        }
      }
      lp *= invA;
      lpo *= invA;
      float wgt = c, wgtd = 0.f;                       // d l / d lp and its tangent
      if (k.loss == XM_RL_A2C) {
        if (ok) loss_acc += (double)(c * lp);
      } else if (k.loss == XM_RL_SURROGATE) {
        const float ratio = expf(lp - lpo);
        wgt = c * ratio;
        float val = wgt;
        if (k.clip > 0.f) {                            // PPO: max(c r, c clamp(r)); the clamped branch has no gradient
          const float rc = fminf(fmaxf(ratio, 1.f - k.clip), 1.f + k.clip);
          const bool inside = ratio >= 1.f - k.clip && ratio <= 1.f + k.clip;
          if (!inside && c * ratio <= c * rc) { wgt = 0.f; val = c * rc; }
        }
        if (ok) { loss_acc += (double)val; kl_acc += (double)k.kl_scale * klv; }
        if (DB) {                                      // tangent of the weight: d(c r)/d eps = c r * lp_dot
          float lpd = 0.f;
          for (int d = 0; d < OUT; ++d) {
            const float e = ACT[s * OUT + d] - MU[s * OUT + d], is2 = expf(-2.f * lam[d]);
            lpd += invA * (e * is2 * MUD[s * OUT + d] + (e * e * is2 - 1.f) * lamd[d]);
          }
          wgtd = wgt * lpd;
        }
      }
#pragma unroll
      for (int d = 0; d < RL_MAXIO; ++d) {
        if (d >= OUT) break;
        const float e = ACT[s * OUT + d] - MU[s * OUT + d], is2 = expf(-2.f * lam[d]);
        float gm, gmd = 0.f;
        if (k.loss == XM_RL_FISHER) {
          gm = ok ? k.kl_scale * MUD[s * OUT + d] * expf(-2.f * lamo[d]) : 0.f;
          if (ok) accls[d] += k.kl_scale * 2.f * lamd[d];
        } else {
          gm = wgt * invA * e * is2;
          if (!DB) {
            accls[d] += wgt * invA * (e * e * is2 - 1.f);
          } else {                                     // tangent of the gradient terms (Hessian-vector product)
            const float mud = MUD[s * OUT + d];
            gmd = wgtd * invA * e * is2 + wgt * invA * is2 * (-mud - 2.f * e * lamd[d]);
            accls[d] += wgtd * invA * (e * e * is2 - 1.f) + wgt * invA * is2 * (-2.f * e * mud - 2.f * e * e * lamd[d]);
          }
        }
        GMU[s * OUT + d] = gm;
        GMUD[s * OUT + d] = gmd;
      }
    }
    __syncthreads();
    if (need_bwd) {
      // ---- gz2 = (W3^T g_mu) * act'(h2)  (+ tangent) ------------------------------------------------------------------
      for (int idx = tid; idx < RL_TS * H2; idx += RL_THREADS) {
        const int s = idx / H2, j = idx - s * H2;
        float gh = 0.f, ghd = 0.f;
        for (int o = 0; o < OUT; ++o) {
          gh = fmaf(W3[o * H2 + j], GMU[s * OUT + o], gh);
          if (DB) ghd = fmaf(W3d[o * H2 + j], GMU[s * OUT + o], fmaf(W3[o * H2 + j], GMUD[s * OUT + o], ghd));
        }
        const float h = A2[idx], d1 = act_d1(h, act);
        G2[idx] = gh * d1;
        if (DB) G2d[idx] = ghd * d1 + gh * act_d1_dot(h, D2[idx], act);
      }
      __syncthreads();
      // ---- output-layer and layer-2 bias gradients: thread = hidden unit (before A2 / D2 are overwritten) ------------
      if (tid < H2) {
        const int j = tid;
        for (int s = 0; s < RL_TS; ++s) {
          accb2 += DB ? G2d[s * H2 + j] : G2[s * H2 + j];
#pragma unroll
          for (int o = 0; o < RL_MAXIO; ++o) {
            if (o >= OUT) break;
            if (DB) accW3[o] = fmaf(GMUD[s * OUT + o], A2[s * H2 + j], fmaf(GMU[s * OUT + o], D2[s * H2 + j], accW3[o]));
            else accW3[o] = fmaf(GMU[s * OUT + o], A2[s * H2 + j], accW3[o]);
          }
        }
      }
      if (tid < OUT)
        for (int s = 0; s < RL_TS; ++s) accb3 += DB ? GMUD[s * OUT + tid] : GMU[s * OUT + tid];
      __syncthreads();
      // ---- gz1 = (W2^T gz2) * act'(h1) (+ tangent): thread = 4 samples x 4 CONSECUTIVE inputs j0 .. j0+3 (one 128-bit load
      // of a W2 row per unit; lanes beyond H1 / 4 idle); written over A2 / D2 ------------------------------------------------
      {
        const int j0 = 4 * tx;
        const bool jact = j0 < H1;
        float acc[4][4], accd[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int c = 0; c < 4; ++c) { acc[a][c] = 0.f; accd[a][c] = 0.f; }
        if (jact)
          for (int i = 0; i < H2; i += 4) {
            float4 gv[4], gdv[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) { gv[a] = lds4(G2 + (ty * 4 + a) * H2 + i); if (DB) gdv[a] = lds4(G2d + (ty * 4 + a) * H2 + i); }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const float4 w = lds4(W2 + (i + r) * ldw + j0);
              float4 wd;
              if (DB) wd = lds4(W2d + (i + r) * ldw + j0);
#pragma unroll
              for (int a = 0; a < 4; ++a) {
                const float gz = comp(gv[a], r);
                acc[a][0] = fmaf(w.x, gz, acc[a][0]); acc[a][1] = fmaf(w.y, gz, acc[a][1]);
                acc[a][2] = fmaf(w.z, gz, acc[a][2]); acc[a][3] = fmaf(w.w, gz, acc[a][3]);
                if (DB) {
                  const float gzd = comp(gdv[a], r);
                  accd[a][0] = fmaf(wd.x, gz, fmaf(w.x, gzd, accd[a][0])); accd[a][1] = fmaf(wd.y, gz, fmaf(w.y, gzd, accd[a][1]));
                  accd[a][2] = fmaf(wd.z, gz, fmaf(w.z, gzd, accd[a][2])); accd[a][3] = fmaf(wd.w, gz, fmaf(w.w, gzd, accd[a][3]));
                }
              }
            }
          }
        __syncthreads();                               // every thread is done reading A2 / D2 of this tile
        // G1 lives in the A2 buffer, G1d in D2 (both hold RL_TS * max(H1, H2) floats, see the host-side size computation)
        if (jact)
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            const int s = ty * 4 + a;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int j = j0 + c;
              const float h = A1[s * H1 + j], d1 = act_d1(h, act);
              A2[s * H1 + j] = acc[a][c] * d1;
              if (DB) D2[s * H1 + j] = accd[a][c] * d1 + acc[a][c] * act_d1_dot(h, D1[s * H1 + j], act);
            }
          }
      }
      __syncthreads();
      // ---- weight gradient of layer 2: thread owns unit blocks 4 (ty + 8 q) .. +3 and inputs 4 tx .. 4 tx + 3 -------------
      {
        const int j0 = 4 * tx;
        if (j0 < H1)
          for (int s = 0; s < RL_TS; ++s) {
            const float4 hv = lds4(A1 + s * H1 + j0);
            float4 hdv;
            if (DB) hdv = lds4(D1 + s * H1 + j0);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int i0 = 4 * (ty + 8 * q);
              if (i0 < H2) {
                const float4 g4 = lds4(G2 + s * H2 + i0);
                float4 gd4;
                if (DB) gd4 = lds4(G2d + s * H2 + i0);
#pragma unroll
                for (int ii = 0; ii < 4; ++ii) {
                  const float gz = comp(g4, ii);
                  float* acc = accW2[4 * q + ii];
                  if (DB) {
                    const float gzd = comp(gd4, ii);
                    acc[0] = fmaf(gzd, hv.x, fmaf(gz, hdv.x, acc[0])); acc[1] = fmaf(gzd, hv.y, fmaf(gz, hdv.y, acc[1]));
                    acc[2] = fmaf(gzd, hv.z, fmaf(gz, hdv.z, acc[2])); acc[3] = fmaf(gzd, hv.w, fmaf(gz, hdv.w, acc[3]));
                  } else {
                    acc[0] = fmaf(gz, hv.x, acc[0]); acc[1] = fmaf(gz, hv.y, acc[1]);
                    acc[2] = fmaf(gz, hv.z, acc[2]); acc[3] = fmaf(gz, hv.w, acc[3]);
                  }
                }
              }
            }
          }
      }
      // ---- layer-1 gradients: thread = hidden unit ----------------------------------------------------------------------------
      if (tid < H1) {
        const int i = tid;
        for (int s = 0; s < RL_TS; ++s) {
          const float gz = DB ? D2[s * H1 + i] : A2[s * H1 + i];
          accb1 += gz;
#pragma unroll
          for (int c = 0; c < RL_MAXIO; ++c)
            if (c < IN) accW1[c] = fmaf(gz, X[s * IN + c], accW1[c]);
        }
      }
    }
    __syncthreads();
  }

  // ---- this CTA's partial sums -> global ------------------------------------------------------------------------------------
  const long long slot = (long long)task * k.G + g;
  {
    const double l = block_sum(loss_acc, s_red);
    const double kl = block_sum(kl_acc, s_red);
    if (tid == 0) { k.partial_sc[slot * 2] = l; k.partial_sc[slot * 2 + 1] = kl; }
  }
  if (need_bwd) {
    float* out = k.partial + slot * k.P;
    if (tid < 32) {                                     // log-std gradients: held by the 32 sample threads
#pragma unroll
      for (int d = 0; d < RL_MAXIO; ++d) {
        if (d >= OUT) break;
        const float v = warp_sum(accls[d]);
        if (tid == 0) out[d] = (sig[d] >= LOG_EPS) ? v : 0.f;
      }
    }
    if (tid < H1) {
#pragma unroll
      for (int c = 0; c < RL_MAXIO; ++c)
        if (c < IN) out[oW1 + tid * IN + c] = accW1[c];
      out[ob1 + tid] = accb1;
    }
    if (tid < H2) {
      out[ob2 + tid] = accb2;
#pragma unroll
      for (int o = 0; o < RL_MAXIO; ++o)
        if (o < OUT) out[oW3 + o * H2 + tid] = accW3[o];
    }
    if (tid < OUT) out[ob3 + tid] = accb3;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int i = 4 * (ty + 8 * (q >> 2)) + (q & 3);
      if (i < H2 && 4 * tx < H1) {
#pragma unroll
        for (int r = 0; r < 4; ++r) out[oW2 + i * H1 + 4 * tx + r] = accW2[q][r];
      }
    }
  }
}

// out[t][p] = (base ? base[t][p] : 0) + scale * sum_g partial[t][g][p]   (fixed order); block (0, t) also reduces the
// scalars.
__global__ void rl_reduce_kernel(const float* __restrict__ partial, const double* __restrict__ partial_sc, int G, int P,
                                 int reduce_vec, int body_lo, int body_hi, float* __restrict__ out, long long out_stride,
                                 const float* __restrict__ base, long long base_stride, float scale,
                                 float* __restrict__ task_loss, float* __restrict__ task_kl) {
  const int task = blockIdx.y;
  const int pidx = blockIdx.x * blockDim.x + threadIdx.x;
  if (reduce_vec && pidx < P) {
    float acc = 0.f;
    for (int g = 0; g < G; ++g) acc += partial[((long long)task * G + g) * P + pidx];
    if (pidx >= body_lo && pidx < body_hi) acc = 0.f;          // ANIL: the body does not adapt in the inner loop
    const float b = base ? base[(long long)task * base_stride + pidx] : 0.f;
    out[(long long)task * out_stride + pidx] = b + scale * acc;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double l = 0.0, kl = 0.0;
    for (int g = 0; g < G; ++g) { l += partial_sc[((long long)task * G + g) * 2]; kl += partial_sc[((long long)task * G + g) * 2 + 1]; }
    if (task_loss) task_loss[task] = (float)l;
    if (task_kl) task_kl[task] = (float)kl;
  }
}

static int sweep_ctas_per_task(int tasks, int n) {
  const int tiles = (n + RL_TS - 1) / RL_TS;
  int G = (2 * num_sms() + tasks - 1) / tasks;
  if (G > tiles) G = tiles;
  if (G < 1) G = 1;
  return G;
}
}  // namespace xm

using namespace xm;

extern "C" int xm_rl_advantages(const XmRlAdvArgs* a, void* stream) {
  XM_REQUIRE(a != nullptr, "xm_rl_advantages: null args");
  XM_REQUIRE(a->replays > 0 && a->n > 0 && a->state_dim >= 1 && a->state_dim <= 4, "xm_rl_advantages: bad sizes");
  XM_REQUIRE(a->states && a->next_states && a->rewards && a->dones && a->coef, "xm_rl_advantages: null pointer");
  const size_t smem = (size_t)a->n * (3 * 8 + 4 + 4 + 2 * 4 * a->state_dim);
  XM_REQUIRE(smem <= 200 * 1024, "xm_rl_advantages: replay too long for one CTA's shared memory");
  XM_CUDA(cudaFuncSetAttribute(rl_adv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  rl_adv_kernel<<<a->replays, ADV_THREADS, smem, (cudaStream_t)stream>>>(*a);
  return launched("xm_rl_advantages");
}

static int sweep_check(const XmRlSweepArgs* a) {
  XM_REQUIRE(a != nullptr, "xm_rl_sweep: null args");
  XM_REQUIRE(a->tasks > 0 && a->n > 0 && a->in_dim >= 1 && a->in_dim <= RL_MAXIO && a->out_dim >= 1 &&
             a->out_dim <= RL_MAXIO && a->h1 >= 1 && a->h1 <= RL_MAXH && a->h2 >= 1 && a->h2 <= RL_MAXH,
             "xm_rl_sweep: bad sizes (dims <= 8, hidden <= 128)");
  XM_REQUIRE(a->h1 % 4 == 0 && a->h2 % 4 == 0, "xm_rl_sweep: hidden widths must be multiples of 4 (128-bit shared loads)");
  XM_REQUIRE(a->activation == XM_ACT_RELU || a->activation == XM_ACT_TANH, "xm_rl_sweep: bad activation");
  XM_REQUIRE(a->loss >= XM_RL_A2C && a->loss <= XM_RL_FISHER && a->what >= XM_RL_FORWARD && a->what <= XM_RL_HVP,
             "xm_rl_sweep: bad loss / what");
  return 0;
}

extern "C" int64_t xm_rl_sweep_scratch_bytes(const XmRlSweepArgs* a) {
  if (sweep_check(a) != 0) return -1;
  const int G = sweep_ctas_per_task(a->tasks, a->n);
  const int P = rl_num_params(a->in_dim, a->out_dim, a->h1, a->h2);
  const int64_t vec = (((int64_t)a->tasks * G * P + 1) / 2) * 2 * 4;
  return vec + (int64_t)a->tasks * G * 2 * 8;
}

extern "C" int xm_rl_sweep(const XmRlSweepArgs* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = sweep_check(a)) return rc;
  XM_REQUIRE(a->states && a->theta && a->partial, "xm_rl_sweep: null states / theta / partial");
  const bool hvp = a->what == XM_RL_HVP, fisher = a->loss == XM_RL_FISHER;
  XM_REQUIRE(!hvp || a->loss != XM_RL_FISHER, "xm_rl_sweep: XM_RL_HVP is defined for XM_RL_A2C / XM_RL_SURROGATE");
  XM_REQUIRE(a->clip >= 0.f && a->clip < 1.f, "xm_rl_sweep: bad clip");
  XM_REQUIRE(!(hvp || fisher) || a->theta_dot, "xm_rl_sweep: theta_dot required");
  XM_REQUIRE(!fisher || a->what == XM_RL_GRAD, "xm_rl_sweep: XM_RL_FISHER produces a gradient");
  XM_REQUIRE(a->loss == XM_RL_FISHER || a->what == XM_RL_FORWARD || (a->actions && a->coef),
             "xm_rl_sweep: actions / coef required");
  XM_REQUIRE(a->loss == XM_RL_A2C || (a->logstd_old && (a->loss == XM_RL_FISHER || (a->mu_old && a->actions && a->coef))),
             "xm_rl_sweep: old-policy outputs required");
  XM_REQUIRE(a->what == XM_RL_FORWARD || a->out, "xm_rl_sweep: out required");
  XM_REQUIRE(a->partial_bytes >= xm_rl_sweep_scratch_bytes(a), "xm_rl_sweep: partial buffer too small");
  const bool DF = hvp || fisher, DB = hvp;
  SweepK k{};
  k.n = a->n; k.in = a->in_dim; k.out = a->out_dim; k.h1 = a->h1; k.h2 = a->h2; k.act = a->activation;
  k.loss = a->loss; k.what = a->what;
  k.G = sweep_ctas_per_task(a->tasks, a->n);
  k.states = a->states; k.actions = a->actions; k.coef = a->coef; k.mu_old = a->mu_old; k.logstd_old = a->logstd_old;
  k.kl_scale = a->kl_scale; k.clip = a->clip; k.head_only = a->head_only;
  k.theta = a->theta; k.theta_stride = a->theta_task_stride;
  k.theta_dot = a->theta_dot; k.dot_stride = a->theta_dot_task_stride;
  k.mu_out = a->mu_out;
  k.P = rl_num_params(a->in_dim, a->out_dim, a->h1, a->h2);
  k.partial = a->partial;
  k.partial_sc = reinterpret_cast<double*>(reinterpret_cast<char*>(a->partial) +
                                           (((int64_t)a->tasks * k.G * k.P + 1) / 2) * 2 * 4);
  auto pad = [](int c) { return (c + 3) & ~3; };
  const int IN = a->in_dim, OUT = a->out_dim, H1 = a->h1, H2 = a->h2, ldw = rl_ldw(H1), HM = H1 > H2 ? H1 : H2;
  const int wset = pad(OUT) + pad(H1 * IN) + pad(H1) + pad(H2 * ldw) + pad(H2) + pad(OUT * H2) + pad(OUT);
  int fl = wset * (DF ? 2 : 1);
  fl += pad(RL_TS * H1) + pad(RL_TS * HM);                          // A1, A2 (A2 also holds gz1: max width)
  if (DF) fl += pad(RL_TS * H1) + pad(RL_TS * HM);                  // D1, D2
  fl += pad(RL_TS * H2) * (DB ? 2 : 1);                             // G2 (+ G2d)
  fl += pad(RL_TS * IN) + pad(RL_TS * OUT) * 6 + pad(RL_TS) + 3 * pad(OUT);
  const size_t smem = (size_t)fl * 4;
  XM_REQUIRE(smem <= 220 * 1024, "xm_rl_sweep: network too large for shared memory");
  dim3 grid(k.G, a->tasks);
  if (DB) {
    XM_CUDA(cudaFuncSetAttribute(rl_sweep_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    rl_sweep_kernel<true, true><<<grid, RL_THREADS, smem, stream>>>(k);
  } else if (DF) {
    XM_CUDA(cudaFuncSetAttribute(rl_sweep_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    rl_sweep_kernel<true, false><<<grid, RL_THREADS, smem, stream>>>(k);
  } else {
    XM_CUDA(cudaFuncSetAttribute(rl_sweep_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    rl_sweep_kernel<false, false><<<grid, RL_THREADS, smem, stream>>>(k);
  }
  if (int rc = launched("xm_rl_sweep")) return rc;
  const int reduce_vec = a->what != XM_RL_FORWARD;
  dim3 rgrid(reduce_vec ? (k.P + 255) / 256 : 1, a->tasks);
  const int body_lo = (a->head_only & 1) ? a->out_dim : 0;
  const int body_hi = (a->head_only & 1) ? a->out_dim + a->h1 * a->in_dim + a->h1 + a->h2 * a->h1 + a->h2 : 0;
  rl_reduce_kernel<<<rgrid, 256, 0, stream>>>(k.partial, k.partial_sc, k.G, k.P, reduce_vec, body_lo, body_hi, a->out, a->out_task_stride,
                                              a->base, a->base_task_stride, a->scale, a->task_loss, a->task_kl);
  return launched("xm_rl_sweep(reduce)");
}
