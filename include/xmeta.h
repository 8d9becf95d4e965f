/* xmeta.h -- C ABI of libxmeta.so: sm_100a kernels for the MAML / ANIL hot path of
 * Kostis-S-Z/exploring_meta (task-batched inner-loop adaptation + second-order outer gradient).
 *
 * The reference has no FFI layer; its boundary is the Python API of core_functions/maml.py,
 * core_functions/vision.py and core_functions/vision_models.py, and every FLOP below that API is an
 * ATen library call.  Each entry point here replaces the ATen calls named in its comment
 * (reference file:line = the call site that bottoms out in them).  INTEGRATION.md shows the
 * ctypes binding a reference maintainer would add.
 *
 * Conventions
 *  - Plain pointers and sizes only.  All pointers are DEVICE pointers owned by the caller (PyTorch
 *    allocations); the library never allocates or frees caller-visible memory.  Scratch comes from
 *    caller-provided buffers whose sizes the *_scratch_bytes queries return.
 *  - Every function returns 0 on success, <0 for an invalid argument (message via xm_last_error()),
 *    >0 = cudaError_t of a failed launch.  No exceptions, no exit/abort.
 *  - Every function is asynchronous on the given stream (cudaStream_t passed as void*), does no
 *    host synchronisation and is CUDA-graph capturable.  Process-wide mutable state: the thread-local last-error
 *    string, the launch counter, the two A/B switches xm_set_precision / xm_set_tcgen05 (read at launch time),
 *    and per-device caches of the SM count / kernel attributes.  XmComm (below) is the only object that owns
 *    device memory.
 *  - "task" = one few-shot task of the meta-batch.  Every tensor carries a leading task dimension and
 *    every parameter pointer a per-task stride in floats (stride 0 = all tasks share the master
 *    weights, as at inner step 0 and for the ANIL body).
 *  - Activations are NHWC fp32: [tasks][n][H][W][C].  User images stay in the reference's NCHW
 *    layout [tasks][rows][C][H][W]; the first conv gathers rows row0, row0+row_step, ... directly
 *    (this is prepare_batch's even/odd split, utils/data_pre.py:121-127, fused into the load).
 *  - Parameters and their gradients use PyTorch's layouts ([cout][cin][3][3], [ways][D]) inside flat
 *    per-task vectors in module.parameters() order (core_functions/vision_models.py:168-185: BN
 *    weight, BN bias, conv weight, conv bias per block; then linear weight, bias).
 *  - Parameter-gradient epilogue ("axpy epilogue"): kernels that produce a parameter gradient g write
 *        out = (base ? base : 0) + scale * g
 *    so that the inner SGD step theta' = theta - lr*g (learn2learn maml_update, called from
 *    core_functions/vision.py:13) and the outer recursion are fused into the producing kernel.
 */
#ifndef XMETA_H_
#define XMETA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XM_VERSION 100

/* Geometry of one ConvBlock call (core_functions/vision_models.py:149-193). */
typedef struct XmBlockGeom {
  int32_t tasks;      /* tasks in this launch                                          */
  int32_t n;          /* images per task in this call (one BN batch)                   */
  int32_t cin, cout;  /* conv channels                                                 */
  int32_t hin, win;   /* block input spatial size                                      */
  int32_t hz, wz;     /* conv output spatial size (pre-pool): stride 1 -> hin, 2 -> (hin+2-3)/2+1 */
  int32_t hp, wp;     /* block output size: pool -> hz/2 (floor), else hz              */
  int32_t stride;     /* 1 (max_pool=True) or 2 (max_pool=False), vision_models.py:158-167 */
  int32_t pool;       /* 1: MaxPool2d(2,2,ceil_mode=False) after the ReLU              */
} XmBlockGeom;

enum { XM_CONV_FWD = 0, XM_CONV_DGRAD = 1 };
enum { XM_STAT_NONE = 0, XM_STAT_SUM_SQ = 1, XM_STAT_SUM_AUX = 2 };

/* xm_conv: 3x3 pad-1 cross-correlation as an implicit GEMM, per-task weights.
 *   FWD  : out[t][n][hz][wz][cout] = conv(src1, w1) (+ conv(src2, w2))            src*: block inputs
 *   DGRAD: out[t][n][hin][win][cin] = conv_transpose(src1, w1) (+ ...(src2, w2))  src*: [t][n][hz][wz][cout]
 * The optional second (src, w) pair is the tangent term of the forward-over-reverse pass
 * (zdot = conv(xdot, W) + conv(x, Wdot)).  Weights are PyTorch-layout [cout][cin][3][3]; the conv bias
 * is never added (it is cancelled by train-mode BN: SURVEY fact 8).
 * Epilogue statistics, accumulated in double per (task, channel) into stats[t][2][cout] (zeroed by
 * the call): SUM_SQ -> {sum v, sum v^2} (BN batch statistics); SUM_AUX -> {sum v, sum v*aux}.
 * Replaces aten::conv2d (vision_models.py:189), convolution_backward's dgrad and the conv terms of
 * _convolution_double_backward (vision/maml_vision.py:112). */
typedef struct XmConvArgs {
  XmBlockGeom g;
  int32_t mode;
  int32_t src_nchw;                       /* 1: src1/src2 are user images (FWD only)   */
  int32_t row0, row_step, rows_per_task;  /* image row selection when src_nchw         */
  int32_t stat_mode;
  const float* src1; const float* w1; int64_t w1_task_stride;
  const float* src2; const float* w2; int64_t w2_task_stride;   /* src2 == NULL: one pair */
  float* out;
  const float* aux;                       /* SUM_AUX: tensor shaped like out           */
  double* stats;
  float* workspace; int64_t workspace_bytes;   /* optional, >= xm_conv_workspace_bytes(&g): lets stride-2 layers run on
                                             the tcgen05 kernels (full-resolution pass + sub-sampling / zero insertion);
                                             NULL = the generic kernels                */
} XmConvArgs;
int64_t xm_conv_workspace_bytes(const XmBlockGeom* g);
int xm_conv(const XmConvArgs* a, void* stream);

/* xm_wgrad: weight gradient gW[co][ci][kh][kw] = sum_{n,h,w} x[n, s*h+kh-1, s*w+kw-1, ci] * g[n,h,w,co]
 * (+ the same for an optional second (x, g) pair: Wdot-gradient = wgrad(x, gzdot) + wgrad(xdot, gz)),
 * split over pixel ranges into `partial` and reduced in double, with the axpy epilogue writing
 * out_w = base_w + scale*gW and out_b = base_b (the conv-bias gradient is analytically zero).
 * Replaces convolution_backward's wgrad + learn2learn update_module's `p + (-lr*g)`. */
typedef struct XmWgradArgs {
  XmBlockGeom g;
  int32_t src_nchw, row0, row_step, rows_per_task;
  const float* x1; const float* g1;
  const float* x2; const float* g2;       /* x2 == NULL: one pair                       */
  float* out_w; float* out_b; int64_t out_task_stride;
  const float* base_w; const float* base_b; int64_t base_task_stride;   /* NULL: zero base */
  float scale;
  float* partial; int64_t partial_bytes;  /* >= xm_wgrad_scratch_bytes(&g)              */
} XmWgradArgs;
int64_t xm_wgrad_scratch_bytes(const XmBlockGeom* g);
int xm_wgrad(const XmWgradArgs* a, void* stream);

/* BatchNorm(train mode, per-call batch statistics) + ReLU + MaxPool, forward / backward and the
 * tangent ("dual") versions of both.  One argument block for the four entry points; each reads the
 * fields its comment names.  Per-(task, channel) scalars live in small fp32 side buffers
 * [tasks][2][cout] that travel between the calls of one block:
 *   mean_invstd : {mean, 1/sqrt(var+eps)}            written by xm_bn_fwd
 *   call_stats  : {mean, unbiased var}               written by xm_bn_fwd (running-stat EMA input)
 *   bwd_red     : {<g>, <g*xhat>}                    written by xm_bn_bwd
 *   dual_red    : {<zdot>, <zdot*xhat>}              written by xm_bn_dual_fwd
 * Replaces native_batch_norm / relu / max_pool2d_with_indices (vision_models.py:190-192), their
 * backward ops, and (dual pair) batchnorm_double_backward + the mask/gather double-backward ops. */
typedef struct XmBnArgs {
  XmBlockGeom g;
  float eps;
  const float* z; const float* zdot;            /* [t][n][hz][wz][cout]                        */
  const double* sums;                           /* {sum z, sum z^2}      from xm_conv SUM_SQ     */
  const double* dsums;                          /* {sum zdot, sum zdot*z} from xm_conv SUM_AUX   */
  const float* gamma; const float* beta; int64_t gb_task_stride;
  const float* gamma_dot; const float* beta_dot; int64_t gbdot_task_stride;
  float* mean_invstd; float* call_stats; float* bwd_red; float* dual_red;
  float* p; float* pdot;                        /* block outputs [t][n][hp][wp][cout]          */
  const float* gp; const float* gpdot;          /* cotangents of p / tangent of gp (NULL = 0)  */
  float* gz; float* gzdot;                      /* gradients w.r.t. z                          */
  float* out_gamma; float* out_beta; int64_t out_task_stride;          /* axpy epilogue      */
  const float* base_gamma; const float* base_beta; int64_t base_task_stride;
  float scale;
  double* scratch;                              /* >= xm_bn_scratch_bytes(&g)                  */
} XmBnArgs;
int64_t xm_bn_scratch_bytes(const XmBlockGeom* g);
/* reads z, sums, gamma, beta                      writes p, mean_invstd, call_stats(opt) */
int xm_bn_fwd(const XmBnArgs* a, void* stream);
/* reads z, gp, mean_invstd, gamma, beta           writes gz, bwd_red, out_gamma/out_beta  */
int xm_bn_bwd(const XmBnArgs* a, void* stream);
/* reads z, zdot, dsums, mean_invstd, gamma(+dot), beta(+dot)   writes pdot, dual_red     */
int xm_bn_dual_fwd(const XmBnArgs* a, void* stream);
/* reads z, zdot, gp, gpdot, mean_invstd, bwd_red, dual_red, gamma(+dot), beta
 * writes gz (recomputed; NULL = not needed), gzdot, out_gamma/out_beta = base + scale * tangent of (g_gamma, g_beta) */
int xm_bn_dual_bwd(const XmBnArgs* a, void* stream);

/* Image block: the FIRST ConvBlock of the network (core_functions/vision_models.py:135-139 -- conv3x3 on the user
 * images, cin <= 4, then BN(train) + ReLU + MaxPool 2x2) as one fused unit that never writes the pre-BN map z.
 * The block input is constant over the inner loop and carries no gradient, so every dense reduction over z is a
 * closed form in the per-task Gram matrix of the im2col'd images (xm_img_gram, once per meta-iteration), and the
 * rest of the backward only visits the pooling winners (derivation: exploring_meta_b200/csrc/img_block.cu).
 * Covered geometry: xm_img_supported(); everything else goes through xm_conv / xm_bn_* / xm_wgrad.
 * K = 9*cin, weight index k = ci*9 + kh*3 + kw (PyTorch's [cout][cin][3][3]).  Side buffers:
 *   gram  [tasks][K*K + K] double : G = X^T X, then sx = X^T 1           (xm_img_gram_bytes)
 *   zsel  [t][n][hp][wp][cout]    : z at the pooling winner (0 where the ReLU is dead)
 *   sel   [t][n][hp][wp][cout] u8 : winner position dy*2+dx inside the window, 255 = ReLU-dead
 *   zdsel like zsel               : tangent of z at the winner
 *   ssum  [tasks][cout][K+3] double: sparse sums {S[K], sum g, sum g*xhat, -} of xm_img_bwd, read by xm_img_dual_bwd
 *   scratch >= xm_img_scratch_bytes
 * mean_invstd / call_stats / bwd_red / dual_red, the axpy epilogue and the row selection are as in XmBnArgs /
 * XmConvArgs.  Replaces, for this block: conv2d + native_batch_norm + relu + max_pool2d_with_indices, their backward
 * and double-backward ops (vision/maml_vision.py:112) and maml_update's `p + (-lr*g)`. */
typedef struct XmImgArgs {
  XmBlockGeom g;
  int32_t row0, row_step, rows_per_task;
  float eps, scale;
  const float* x;                               /* user images [tasks][rows_per_task][cin][hin][win]      */
  double* gram;
  const float* w; int64_t w_task_stride;
  const float* w_dot; int64_t wdot_task_stride;
  const float* gamma; const float* beta; int64_t gb_task_stride;
  const float* gamma_dot; const float* beta_dot; int64_t gbdot_task_stride;
  float* mean_invstd; float* call_stats; float* bwd_red; float* dual_red;
  float* p; float* zsel; uint8_t* sel;
  float* pdot; float* zdsel;
  const float* gp; const float* gpdot;          /* cotangent of p and its tangent (gpdot NULL = 0)         */
  double* ssum;
  double* scratch;
  float* out_w; float* out_b; float* out_gamma; float* out_beta; int64_t out_task_stride;
  const float* base_w; const float* base_b; const float* base_gamma; const float* base_beta; int64_t base_task_stride;
} XmImgArgs;
int xm_img_supported(const XmBlockGeom* g);
int64_t xm_img_gram_bytes(const XmBlockGeom* g);
int64_t xm_img_scratch_bytes(const XmBlockGeom* g);
/* reads x                                              writes gram                                          */
int xm_img_gram(const XmImgArgs* a, void* stream);
/* reads x, gram, w, gamma, beta                        writes p, zsel, sel, mean_invstd, call_stats(opt)    */
int xm_img_fwd(const XmImgArgs* a, void* stream);
/* reads x, gram, w, gamma, mean_invstd, gp, zsel, sel  writes bwd_red, ssum(opt), out_* = base + scale*grad  */
int xm_img_bwd(const XmImgArgs* a, void* stream);
/* reads x, gram, w, w_dot, gamma(+dot), beta_dot, mean_invstd, zsel, sel     writes pdot, zdsel, dual_red     */
int xm_img_dual_fwd(const XmImgArgs* a, void* stream);
/* reads x, gram, w, w_dot, gamma(+dot), mean_invstd, bwd_red, dual_red, ssum, gp, gpdot, zsel, zdsel, sel
 * writes out_* = base + scale * tangent of (gW, gb, g_gamma, g_beta)                                         */
int xm_img_dual_bwd(const XmImgArgs* a, void* stream);

/* xm_head: classifier head on block-4 features, one CTA per task, forward + loss + backward fused.
 *   mode 0 (MiniImagenetCNN.forward, vision_models.py:107-110): X = flatten in NCHW order, D = c*hw
 *   mode 1 (OmniglotCNN.forward :51-55):                        X = mean over hw,           D = c
 *   logits = X W^T + b ; loss = CrossEntropyLoss(mean) (vision/maml_vision.py:86) ; correct = #(argmax == y)
 *   (core_functions/vision.py:21-23, lowest index wins ties).
 * dual == 0: writes loss/correct/logits (each optional), g_feat = dL/dfeat and out_{w,b} = base + scale*dL/d{w,b}.
 * dual == 1: additionally takes tangents (feat_dot, w_dot, b_dot), writes g_feat_dot = tangent of dL/dfeat and
 *            out_{w,b} = base + scale * tangent of dL/d{w,b}   (softmax-CE Hessian term included).
 * Replaces aten::linear + cross_entropy forward/backward/double-backward. */
typedef struct XmHeadArgs {
  int32_t tasks, n, ways, c, hw, mode, dual;
  const float* feat; const float* feat_dot;      /* [t][n][hw][c] (feat_dot NULL = 0)           */
  const int64_t* labels; int32_t label_row0, label_row_step, labels_per_task;
  const float* w; const float* b; int64_t wb_task_stride;
  const float* w_dot; const float* b_dot; int64_t wbdot_task_stride;
  float* loss; int32_t* correct; float* logits;  /* [t], [t], [t][n][ways]; each may be NULL    */
  float* g_feat; float* g_feat_dot;
  float* out_w; float* out_b; int64_t out_task_stride;
  const float* base_w; const float* base_b; int64_t base_task_stride;
  float scale;
} XmHeadArgs;
int xm_head(const XmHeadArgs* a, void* stream);

/* xm_anil_head: ANIL's head-only adaptation (vision/anil_vision.py:116-122 + core_functions/vision.py:9-17
 * with features != None): per task, `steps` second-order (or first-order) inner steps of a Linear(D, ways)
 * on the support feature rows (even rows), query loss / correct count on the odd rows, and the outer
 * gradient w.r.t. the head initialisation (g_w, g_b: per task) and w.r.t. ALL feature rows (g_feat).
 * feat is the body output for all rows of the task: [t][rows][hw][c], flattened like xm_head mode 0/1.
 * One thread-block cluster per task, every intermediate in (distributed) shared memory: `scratch` is unused (may be NULL;
 * xm_anil_head_scratch_bytes returns 0 and is kept for callers of the earlier ABI). */
typedef struct XmAnilHeadArgs {
  int32_t tasks, rows, ways, c, hw, mode, steps, first_order;
  float lr;
  const float* feat; const int64_t* labels;     /* [t][rows][hw][c], [t][rows]                 */
  const float* w; const float* b;               /* shared head initialisation [ways][D], [ways] */
  float* loss; int32_t* correct;                /* [t]                                         */
  float* g_feat;                                /* [t][rows][hw][c]                            */
  float* g_w; float* g_b; int64_t g_task_stride;
  float* scratch; int64_t scratch_bytes;
} XmAnilHeadArgs;
int64_t xm_anil_head_scratch_bytes(const XmAnilHeadArgs* a);
int xm_anil_head(const XmAnilHeadArgs* a, void* stream);

/* xm_sample_tasks: on-device few-shot task sampler over a dataset resident in HBM as uint8: what
 * train_tasks.sample() yields in the reference -- learn2learn's NWays -> KShots(2k) -> LoadData -> RemapLabels ->
 * ConsecutiveLabels (-> RandomClassRotation) chain of utils/data_pre.py:28-37,79-85 -- for `tasks` tasks at once:
 *   x[t][ways*shots2][C][H][W] = scale * u8 + offset, samples grouped by class, one quarter-turn rotation per class
 *   of the task when `rotate`;  y[t][ways*shots2] = 0..ways-1, each shots2 times.
 * Items of class c are data[class_start[c] .. class_start[c+1]) (every class needs >= shots2 items).  Draws are
 * counter based (splitmix64 of seed, first_task + t, draw counter): a batch is a pure function of (seed, first_task);
 * oracle/task_sampler_oracle.py restates the algorithm bit for bit.  items / classes (optional) return the chosen
 * item indices [t][ways*shots2] and class ids [t][ways]. */
typedef struct XmSampleArgs {
  int32_t tasks, ways, shots2, channels, height, width, num_classes, rotate;
  uint64_t seed; int64_t first_task;
  const uint8_t* data; const int32_t* class_start;
  float scale, offset;
  float* x; int64_t* y;
  int32_t* items; int32_t* classes;
} XmSampleArgs;
int xm_sample_tasks(const XmSampleArgs* a, void* stream);

/* dst[i] = (accumulate ? dst[i] : 0) + sum_t src[t*task_stride + i], tasks added in order in fp32 --
 * the order in which eval_loss.backward() accumulates into the master .grad (vision/maml_vision.py:112). */
int xm_accumulate_tasks(const float* src, int64_t task_stride, int32_t tasks, int64_t count,
                        float* dst, int32_t accumulate, void* stream);

/* Outer step (vision/maml_vision.py:139-141): g = grad*grad_scale (the 1/meta_batch_size), then
 * torch.optim.Adam's update with bias correction for step number `step` (1-based), no weight decay. */
int xm_adam_step(float* theta, const float* grad, float* m, float* v, int64_t count, float grad_scale,
                 float lr, float beta1, float beta2, float eps, int32_t step, void* stream);

/* ---- sharded outer step ------------------------------------------------------------
 * The reference sums the tasks' gradients into the master .grad inside one process (vision/maml_vision.py:112) and
 * steps Adam (:139-141).  With the meta-batch sharded over GPUs (one process per GPU) the shard sums are combined
 * over NVLink peer memory INSIDE the Adam kernel: no NCCL call, one launch, graph-capturable.
 *
 * XmComm is the one object of this library that owns device memory: a staging / flag block per rank, exported with
 * cudaIpcGetMemHandle.  xm_comm_create fills handle_out[XM_IPC_HANDLE_BYTES]; the caller exchanges the handles of
 * all ranks out of band (rank order) and passes the world * XM_IPC_HANDLE_BYTES bytes to xm_comm_connect.
 * n_floats = length of the flat buffer every later xm_allreduce_adam call reduces.  xm_comm_error returns 1 after a
 * peer failed to arrive within ~4 s (the kernel gives up instead of hanging the GPU), -1 on a CUDA error. */
#define XM_COMM_MAX_WORLD 8
#define XM_IPC_HANDLE_BYTES 64
typedef struct XmComm XmComm;
int xm_comm_create(int32_t world, int32_t rank, int64_t n_floats, XmComm** out, unsigned char* handle_out);
int xm_comm_connect(XmComm* comm, const unsigned char* handles);
int xm_comm_error(XmComm* comm);
int xm_comm_destroy(XmComm* comm);

/* xm_allreduce_adam: reduced[i] = sum over ranks (rank order) of local[i], i < n_total; then for i < n_params
 * torch.optim.Adam's update of theta / m / v with g = reduced[i] * grad_scale and the bias correction of step
 * number *step (1-based, device memory: advanced by xm_finish_shard, so a captured graph replays correctly).
 * comm == NULL (or world 1): reduced = local, single-GPU outer step.  Every rank must issue the same sequence of
 * calls on its communicator. */
typedef struct XmAdamArgs {
  float* theta; float* m; float* v; int64_t n_params;
  const float* local; float* reduced; int64_t n_total;
  float grad_scale, lr, beta1, beta2, eps;
  const int32_t* step;
} XmAdamArgs;
int xm_allreduce_adam(XmComm* comm, const XmAdamArgs* a, void* stream);

/* out2[0] = sum_t loss[t], out2[1] = sum_t correct[t] (the driver's running sums, vision/maml_vision.py:113-115),
 * *step += 1 (step may be NULL). */
int xm_finish_shard(const float* loss, const int32_t* correct, int32_t tasks, float* out2, int32_t* step,
                    void* stream);

/* ---- config 5: MAML-TRPO policy-MLP path (ABI item 8) ---------------------------------------
 * Policy = core_functions/policies.py:30-56 DiagNormalPolicy with two hidden layers: mean(s) = W3 act(W2 act(W1 s +
 * b1) + b2) + b3, log-std = clamp(sigma, min=log 1e-6), log_prob = MEAN over action dims of the Normal log-density.
 * Flat parameter vector in DiagNormalPolicy.parameters() order: sigma[out], W1[h1][in], b1[h1], W2[h2][h1], b2[h2],
 * W3[out][h2], b3[out] (10,604 floats for 2 -> 100 -> 100 -> 2).  Replays are [tasks][n][dim] fp32, n transitions.
 *
 * xm_rl_advantages: per replay, everything core_functions/rl.py:95-110 + cherry compute from the replay alone --
 * discounted returns (ch.td.discount), LinearValue ridge fit on [s, s^2, t, t^2, t^3, 1] and its values, bootstraps,
 * GAE (ch.pg.generalized_advantage), ch.normalize -- and writes coef[n] = coef_scale * normalised advantage (the
 * per-sample weight of the policy losses: -1/n for a2c.policy_loss :358, -1/(n*tasks) for the surrogate :469).
 * Independent of the policy: computed once per meta-optimisation instead of once per loss evaluation.  Reverse scans
 * run sequentially inside an episode (the reference's order) with one thread per episode; sums in double. */
typedef struct XmRlAdvArgs {
  int32_t replays, n, state_dim;
  double gamma, tau, reg, coef_scale;
  const float* states; const float* next_states;     /* [replays][n][state_dim] */
  const float* rewards; const float* dones;          /* [replays][n]            */
  float* coef;                                       /* [replays][n]            */
  float* returns;                                    /* optional [replays][n] discounted returns */
  float* advantages;                                 /* optional [replays][n] un-normalised GAE advantages */
} XmRlAdvArgs;
int xm_rl_advantages(const XmRlAdvArgs* a, void* stream);

/* xm_rl_sweep: one pass of every task's policy over its n transitions, per-task parameters (theta + t*stride).
 *   loss XM_RL_A2C       l = sum_n coef[n] * log_prob(a_n | s_n)                     (trpo_a2c_loss, rl.py:346-358)
 *        XM_RL_SURROGATE l = sum_n coef[n] * exp(log_prob_new - log_prob_old),  kl = mean KL(new || old)/tasks_total
 *                        with the old policy given by its means mu_old[n][out] and log-std (rl.py:459-469).
 *                        clip > 0: the PPO objective of fast_adapt_ppo (rl.py:264-316, ch.algorithms.ppo.policy_loss):
 *                        l = sum_n max(coef[n] * r, coef[n] * clamp(r, 1 - clip, 1 + clip)), r the probability ratio
 *        XM_RL_FISHER    the Gauss-Newton factor of that KL at new == old: the output cotangent is
 *                        F * (tangent of the outputs in direction theta_dot), F = diag(1/sigma_old^2, 2) * kl_scale
 *   what XM_RL_FORWARD   values only (task_loss / task_kl, optional mu_out): the line search of rl.py:429-438
 *        XM_RL_GRAD      d l / d theta                          (theta_dot: only for XM_RL_FISHER, where it is required)
 *        XM_RL_HVP       d/d eps [ d l / d theta ](theta + eps * theta_dot)   (XM_RL_A2C, XM_RL_SURROGATE): the Hessian-
 *                        vector product of the inner loss = what create_graph=True back-propagates through
 *                        (rl.py:368-374; learner.adapt(loss) at :291 for the PPO inner loop)
 * head_only (the ANIL policy, policies.py:70-126: the body is frozen in the inner loop, `turn_off_body_grads`):
 *   bit 0: the body entries (W1, b1, W2, b2) of the RESULT are zeroed before the epilogue (theta' = theta - lr * M g);
 *   bit 1: the body entries of theta_dot are zeroed on load (cotangent v - lr * H (M v)).
 * Result r [P] per task through the axpy epilogue out = (base ? base : 0) + scale * r  (maml_update fused:
 * theta' = theta - inner_lr * grad; cotangent recursion v - inner_lr * H v).  Deterministic: per-CTA partial sums in
 * `partial` (xm_rl_sweep_scratch_bytes), reduced in a fixed order. */
enum { XM_RL_A2C = 0, XM_RL_SURROGATE = 1, XM_RL_FISHER = 2 };
enum { XM_RL_FORWARD = 0, XM_RL_GRAD = 1, XM_RL_HVP = 2 };
enum { XM_ACT_RELU = 0, XM_ACT_TANH = 1 };
typedef struct XmRlSweepArgs {
  int32_t tasks, n, in_dim, out_dim, h1, h2, activation, loss, what;
  const float* states; const float* actions;                  /* [tasks][n][in_dim], [tasks][n][out_dim] */
  const float* coef;                                          /* [tasks][n] (unused by XM_RL_FISHER)     */
  const float* mu_old; const float* logstd_old;               /* [tasks][n][out_dim], [tasks][out_dim]   */
  float kl_scale;                                             /* 1 / (n * out_dim * tasks_total)         */
  const float* theta; int64_t theta_task_stride;
  const float* theta_dot; int64_t theta_dot_task_stride;
  float* out; int64_t out_task_stride;
  const float* base; int64_t base_task_stride; float scale;
  float* task_loss; float* task_kl;                           /* optional [tasks]                        */
  float* mu_out;                                              /* optional [tasks][n][out_dim]            */
  float* partial; int64_t partial_bytes;
  float clip;                                                 /* PPO clip ratio (0: unclipped surrogate)  */
  int32_t head_only;                                          /* ANIL masks, see above                   */
} XmRlSweepArgs;
int64_t xm_rl_sweep_scratch_bytes(const XmRlSweepArgs* a);
int xm_rl_sweep(const XmRlSweepArgs* a, void* stream);

/* BatchNorm running-statistics side effect, composed sequentially like the reference's shared
 * buffers see it: for o in [0,n_outer) for i in [0,n_inner): r <- (1-m) r + m s(o,i), where
 * s(o,i) = {mean[C], unbiased var[C]} at call_stats + o*outer_stride + i*inner_stride (floats). */
int xm_bn_ema(float* running_mean, float* running_var, const float* call_stats, int32_t n_outer,
              int64_t outer_stride, int32_t n_inner, int64_t inner_stride, int32_t channels,
              float momentum, void* stream);

/* Contraction precision of xm_conv / xm_wgrad: 1 (default) = error-compensated 3xTF32 (fp32-level accuracy, what the
 * parity contract is stated in); 0 = single-pass TF32 (faster, ~1e-3 relative error per contraction); 2 = the same
 * 3-term expansion on fp16 hi / lo pairs with per-tile power-of-two scaling on the tcgen05 kernels (kind::f16: equal
 * or better accuracy than 1, measured NOT faster -- DESIGN section 8 -- kept for A/B comparison; XM_PRECISION=2 in the
 * environment selects it at load time).  Process-wide; set before building / capturing a launch program. */
int xm_set_precision(int precise);
/* Kernel selection for A/B comparison.  1 (default): 32-channel stride-1 contractions run on the tcgen05 / TMEM
 * kernels and the image layer (cin <= 4, stride 1) on the exact-fp32 CUDA-core kernel; 2: the image-layer forward
 * also runs on tcgen05; 0: every shape uses the generic mma.sync kernels. */
int xm_set_tcgen05(int enable);

int xm_version(void);
const char* xm_last_error(void);
/* Number of kernel launches this library has issued from the calling process (bench.py's gpu_launches). */
int64_t xm_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif  /* XMETA_H_ */
