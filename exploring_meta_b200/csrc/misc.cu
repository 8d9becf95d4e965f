// misc.cu -- error slot, launch counter, and the small outer-step kernels of libxmeta.
#include "common.cuh"

namespace xm {
thread_local char g_err[512] = {0};
std::atomic<long long> g_launches{0};

int num_sms() {
  static int cache[64] = {0};                  // per device ordinal: a process may drive several GPUs
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cache[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev] = n;
  }
  return cache[dev];
}

int wave_ctas(const void* kernel, int threads, size_t smem) {
  struct Entry { const void* k; int threads; size_t smem; int value; };
  static thread_local Entry cache[32];
  static thread_local int used = 0;
  for (int i = 0; i < used; ++i)
    if (cache[i].k == kernel && cache[i].threads == threads && cache[i].smem == smem) return cache[i].value;
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    per_sm = 1;
  }
  const int value = per_sm * num_sms();
  if (used < 32) cache[used++] = Entry{kernel, threads, smem, value};
  return value;
}

// dst[i] = (accumulate ? dst[i] : 0) + sum_t src[t*stride + i], tasks added in index order in fp32
// (= the order eval_loss.backward() accumulates into the master .grad, vision/maml_vision.py:112).
__global__ void accumulate_tasks_kernel(const float* __restrict__ src, long long stride, int tasks,
                                        long long count, float* __restrict__ dst, int accumulate) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float acc = accumulate ? dst[i] : 0.f;
  for (int t = 0; t < tasks; ++t) acc += src[(long long)t * stride + i];
  dst[i] = acc;
}

// torch.optim.Adam (no weight decay / amsgrad) on g = grad*grad_scale  (vision/maml_vision.py:139-141)
__global__ void adam_kernel(float* __restrict__ theta, const float* __restrict__ grad, float* __restrict__ m,
                            float* __restrict__ v, long long count, float grad_scale, float step_size,
                            float beta1, float beta2, float eps, float bias2_sqrt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const float g = grad[i] * grad_scale;
  const float mi = beta1 * m[i] + (1.f - beta1) * g;
  const float vi = beta2 * v[i] + (1.f - beta2) * g * g;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bias2_sqrt + eps;
  theta[i] = theta[i] - step_size * (mi / denom);
}

// Sequential EMA over n_outer x n_inner BatchNorm calls (learn2learn clones share the master buffers).
__global__ void bn_ema_kernel(float* __restrict__ rm, float* __restrict__ rv, const float* __restrict__ stats,
                              int n_outer, long long outer_stride, int n_inner, long long inner_stride,
                              int C, float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean = rm[c], var = rv[c];
  for (int o = 0; o < n_outer; ++o)
    for (int i = 0; i < n_inner; ++i) {
      const float* s = stats + o * outer_stride + i * inner_stride;
      mean = (1.f - momentum) * mean + momentum * s[c];
      var = (1.f - momentum) * var + momentum * s[C + c];
    }
  rm[c] = mean;
  rv[c] = var;
}
}  // namespace xm

using namespace xm;

extern "C" int xm_version(void) { return XM_VERSION; }
extern "C" const char* xm_last_error(void) { return g_err; }
extern "C" int64_t xm_launch_count(void) { return (int64_t)g_launches.load(); }

extern "C" int xm_accumulate_tasks(const float* src, int64_t task_stride, int32_t tasks, int64_t count,
                                   float* dst, int32_t accumulate, void* stream) {
  XM_REQUIRE(src && dst && tasks > 0 && count > 0, "xm_accumulate_tasks: bad arguments");
  const int threads = 256;
  accumulate_tasks_kernel<<<(unsigned)((count + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
      src, task_stride, tasks, count, dst, accumulate);
  return launched("xm_accumulate_tasks");
}

extern "C" int xm_adam_step(float* theta, const float* grad, float* m, float* v, int64_t count,
                            float grad_scale, float lr, float beta1, float beta2, float eps, int32_t step,
                            void* stream) {
  XM_REQUIRE(theta && grad && m && v && count > 0 && step >= 1, "xm_adam_step: bad arguments");
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const int threads = 256;
  adam_kernel<<<(unsigned)((count + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
      theta, grad, m, v, count, grad_scale, (float)((double)lr / bc1), beta1, beta2, eps, (float)sqrt(bc2));
  return launched("xm_adam_step");
}

extern "C" int xm_bn_ema(float* running_mean, float* running_var, const float* call_stats, int32_t n_outer,
                         int64_t outer_stride, int32_t n_inner, int64_t inner_stride, int32_t channels,
                         float momentum, void* stream) {
  XM_REQUIRE(running_mean && running_var && call_stats && n_outer > 0 && n_inner > 0 && channels > 0,
             "xm_bn_ema: bad arguments");
  bn_ema_kernel<<<(channels + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      running_mean, running_var, call_stats, n_outer, outer_stride, n_inner, inner_stride, channels, momentum);
  return launched("xm_bn_ema");
}
