// comm.cu -- the outer step of a sharded meta-iteration as ONE kernel: sum the ranks' flat [meta-gradient ; loss sum ;
// correct count ; BN running-statistics partials] buffers over NVLink peer memory and apply the replicated Adam update
// (vision/maml_vision.py:112,139-141 with the task loop sharded over GPUs).
//
//   xm_comm_create / xm_comm_connect / xm_comm_destroy : per-process communicator.  Each rank owns one cudaMalloc'ed
//       block {2 staging buffers ; arrival flags}, exported with cudaIpcGetMemHandle; the ranks exchange the 64-byte
//       handles out of band (torch.distributed all_gather_object in comm.py) and map each other's block.
//   xm_allreduce_adam : every CTA owns a slice of the flat buffer: it publishes its slice to its own staging buffer,
//       raises its arrival flag in every peer's block (st.release.sys over NVLink), waits for the peers' flags of the
//       SAME slice, sums the peers' slices in rank order (identical order on every rank => bit-identical replicas)
//       and runs torch.optim.Adam's update on the parameter part of the slice.  One launch, graph-capturable, no NCCL
//       call, no host synchronisation; the 132 KB exchange costs one NVLink round trip instead of an allreduce launch
//       plus three element-wise kernels.  With comm == NULL it is the single-GPU outer step (no exchange).
//   xm_finish_shard : loss sum / correct count of the shard in task order + advance of the device-side Adam step.
//
// Synchronisation: the arrival flag of (source rank, CTA) carries the call number ("epoch", a per-CTA counter in device
// memory that only that CTA advances).  Staging is double-buffered by epoch parity: a rank can run at most one call
// ahead of a peer (its next call waits for that peer's next flag), so the buffer a slow peer is still reading is never
// the one being rewritten.  A wait that exceeds ~4 s sets the communicator's error word instead of hanging the GPU.
#include <string.h>
#include "common.cuh"

struct XmComm {
  int world, rank, blocks;
  long long n_pad;                 // floats per staging buffer (multiple of 4)
  size_t bytes;
  void* local;                     // this rank's block
  void* peer[XM_COMM_MAX_WORLD];   // mapped blocks (peer[rank] == local)
  // device-side tables
  float** d_stage;                 // [world] staging base of every rank
  unsigned** d_flags;              // [world] flag base of every rank
  unsigned* d_epoch;               // [blocks] per-CTA call counter (local)
  int* d_error;                    // [1] set when a wait timed out
  int device;
};

namespace xm {

constexpr int CA_THREADS = 256;
constexpr int CA_PER_THREAD = 4;
constexpr int CA_CHUNK = CA_THREADS * CA_PER_THREAD;

struct AdamK {
  float* theta; float* m; float* v; long long n_params;
  const float* local; float* reduced; long long n_total;
  float grad_scale, lr, beta1, beta2, eps;
  const int* step;
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float ld_relaxed_sys(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(CA_THREADS)
allreduce_adam_kernel(int world, int rank, int blocks, long long n_pad, float* const* __restrict__ stage,
                      unsigned* const* __restrict__ flags, unsigned* __restrict__ epoch, int* __restrict__ error,
                      const AdamK k) {
  const int b = blockIdx.x, tid = threadIdx.x;
  const long long base = (long long)b * CA_CHUNK;
  float val[CA_PER_THREAD];
#pragma unroll
  for (int j = 0; j < CA_PER_THREAD; ++j) {
    const long long i = base + tid + j * CA_THREADS;
    val[j] = i < k.n_total ? k.local[i] : 0.f;
  }
  if (world > 1) {
    const unsigned e = epoch[b] + 1u;
    float* mine = stage[rank] + (long long)(e & 1u) * n_pad;
#pragma unroll
    for (int j = 0; j < CA_PER_THREAD; ++j) {
      const long long i = base + tid + j * CA_THREADS;
      if (i < k.n_total) mine[i] = val[j];
    }
    __syncthreads();
    if (tid < world && tid != rank) {
      __threadfence_system();
      st_release_sys(flags[tid] + (long long)rank * blocks + b, e);          // "my slice b of call e is published"
    }
    if (tid < world && tid != rank) {
      const unsigned* f = flags[rank] + (long long)tid * blocks + b;
      const long long t0 = clock64();
      while ((int)(ld_acquire_sys(f) - e) < 0) {
        if (clock64() - t0 > 8000000000LL) { atomicExch(error, 1); break; }   // ~4 s at 1.9 GHz: a peer never arrived
        __nanosleep(64);
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < CA_PER_THREAD; ++j) val[j] = 0.f;
    for (int q = 0; q < world; ++q) {                                         // rank order: same sum on every rank
      const float* src = stage[q] + (long long)(e & 1u) * n_pad;
#pragma unroll
      for (int j = 0; j < CA_PER_THREAD; ++j) {
        const long long i = base + tid + j * CA_THREADS;
        if (i < k.n_total) val[j] += ld_relaxed_sys(src + i);
      }
    }
    if (tid == 0) epoch[b] = e;
  }
  // bias corrections in double, exactly as xm_adam_step computes them on the host
  __shared__ float s_adam[2];
  if (tid == 0) {
    const int step = *k.step;
    const double bc1 = 1.0 - pow((double)k.beta1, (double)step);
    const double bc2 = 1.0 - pow((double)k.beta2, (double)step);
    s_adam[0] = (float)((double)k.lr / bc1);
    s_adam[1] = (float)sqrt(bc2);
  }
  __syncthreads();
  const float step_size = s_adam[0], bias2_sqrt = s_adam[1];
#pragma unroll
  for (int j = 0; j < CA_PER_THREAD; ++j) {
    const long long i = base + tid + j * CA_THREADS;
    if (i >= k.n_total) continue;
    k.reduced[i] = val[j];
    if (i < k.n_params) {
      const float g = val[j] * k.grad_scale;
      const float mi = k.beta1 * k.m[i] + (1.f - k.beta1) * g;
      const float vi = k.beta2 * k.v[i] + (1.f - k.beta2) * g * g;
      k.m[i] = mi;
      k.v[i] = vi;
      const float denom = sqrtf(vi) / bias2_sqrt + k.eps;
      k.theta[i] = k.theta[i] - step_size * (mi / denom);
    }
  }
}

// out2[0] = sum_t loss[t], out2[1] = sum_t correct[t] (task order, fp32 like the driver's running sums,
// vision/maml_vision.py:113-115); *step += 1.
__global__ void finish_shard_kernel(const float* __restrict__ loss, const int* __restrict__ correct, int tasks,
                                    float* __restrict__ out2, int* __restrict__ step) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float s = 0.f, c = 0.f;
    for (int t = 0; t < tasks; ++t) { s += loss[t]; c += (float)correct[t]; }
    out2[0] = s;
    out2[1] = c;
    if (step) *step += 1;
  }
}
}  // namespace xm

using namespace xm;

static int comm_blocks(long long n) { return (int)((n + CA_CHUNK - 1) / CA_CHUNK); }

extern "C" int xm_comm_create(int32_t world, int32_t rank, int64_t n_floats, XmComm** out, unsigned char* handle_out) {
  XM_REQUIRE(out && handle_out, "xm_comm_create: null out / handle_out");
  XM_REQUIRE(world >= 1 && world <= XM_COMM_MAX_WORLD && rank >= 0 && rank < world && n_floats > 0,
             "xm_comm_create: bad world / rank / size");
  XmComm* c = new XmComm();
  memset(c, 0, sizeof(*c));
  c->world = world; c->rank = rank;
  c->blocks = comm_blocks(n_floats);
  c->n_pad = (long long)c->blocks * CA_CHUNK;
  const size_t stage_bytes = 2 * (size_t)c->n_pad * 4;
  const size_t flag_bytes = (size_t)world * c->blocks * 4;
  c->bytes = stage_bytes + flag_bytes;
  XM_CUDA(cudaGetDevice(&c->device));
  XM_CUDA(cudaMalloc(&c->local, c->bytes));
  XM_CUDA(cudaMemset(c->local, 0, c->bytes));
  XM_CUDA(cudaMalloc(&c->d_stage, sizeof(float*) * XM_COMM_MAX_WORLD));
  XM_CUDA(cudaMalloc(&c->d_flags, sizeof(unsigned*) * XM_COMM_MAX_WORLD));
  XM_CUDA(cudaMalloc(&c->d_epoch, sizeof(unsigned) * c->blocks));
  XM_CUDA(cudaMemset(c->d_epoch, 0, sizeof(unsigned) * c->blocks));
  XM_CUDA(cudaMalloc(&c->d_error, sizeof(int)));
  XM_CUDA(cudaMemset(c->d_error, 0, sizeof(int)));
  cudaIpcMemHandle_t h;
  XM_CUDA(cudaIpcGetMemHandle(&h, c->local));
  static_assert(sizeof(h) == XM_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
  memcpy(handle_out, &h, sizeof(h));
  XM_CUDA(cudaDeviceSynchronize());
  *out = c;
  return 0;
}

extern "C" int xm_comm_connect(XmComm* c, const unsigned char* handles) {
  XM_REQUIRE(c && handles, "xm_comm_connect: null argument");
  float* stage[XM_COMM_MAX_WORLD] = {nullptr};
  unsigned* flags[XM_COMM_MAX_WORLD] = {nullptr};
  const size_t stage_bytes = 2 * (size_t)c->n_pad * 4;
  for (int q = 0; q < c->world; ++q) {
    if (q == c->rank) {
      c->peer[q] = c->local;
    } else {
      cudaIpcMemHandle_t h;
      memcpy(&h, handles + (size_t)q * XM_IPC_HANDLE_BYTES, sizeof(h));
      XM_CUDA(cudaIpcOpenMemHandle(&c->peer[q], h, cudaIpcMemLazyEnablePeerAccess));
    }
    stage[q] = reinterpret_cast<float*>(c->peer[q]);
    flags[q] = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(c->peer[q]) + stage_bytes);
  }
  XM_CUDA(cudaMemcpy(c->d_stage, stage, sizeof(stage), cudaMemcpyHostToDevice));
  XM_CUDA(cudaMemcpy(c->d_flags, flags, sizeof(flags), cudaMemcpyHostToDevice));
  return 0;
}

extern "C" int xm_comm_error(XmComm* c) {
  if (!c) return 0;
  int e = 0;
  if (cudaMemcpy(&e, c->d_error, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return e;
}

extern "C" int xm_comm_destroy(XmComm* c) {
  if (!c) return 0;
  for (int q = 0; q < c->world; ++q)
    if (q != c->rank && c->peer[q]) cudaIpcCloseMemHandle(c->peer[q]);
  cudaFree(c->local); cudaFree(c->d_stage); cudaFree(c->d_flags); cudaFree(c->d_epoch); cudaFree(c->d_error);
  delete c;
  return 0;
}

extern "C" int xm_allreduce_adam(XmComm* c, const XmAdamArgs* a, void* stream) {
  XM_REQUIRE(a && a->theta && a->m && a->v && a->local && a->reduced && a->step, "xm_allreduce_adam: null argument");
  XM_REQUIRE(a->n_params > 0 && a->n_total >= a->n_params, "xm_allreduce_adam: bad sizes");
  const int blocks = comm_blocks(a->n_total);
  XM_REQUIRE(!c || (c->blocks == blocks), "xm_allreduce_adam: communicator was created for a different buffer size");
  AdamK k{a->theta, a->m, a->v, a->n_params, a->local, a->reduced, a->n_total,
          a->grad_scale, a->lr, a->beta1, a->beta2, a->eps, a->step};
  const bool multi = c && c->world > 1;
  allreduce_adam_kernel<<<blocks, CA_THREADS, 0, (cudaStream_t)stream>>>(
      multi ? c->world : 1, multi ? c->rank : 0, blocks, multi ? c->n_pad : 0, multi ? c->d_stage : nullptr,
      multi ? c->d_flags : nullptr, multi ? c->d_epoch : nullptr, multi ? c->d_error : nullptr, k);
  return launched("xm_allreduce_adam");
}

extern "C" int xm_finish_shard(const float* loss, const int32_t* correct, int32_t tasks, float* out2, int32_t* step,
                               void* stream) {
  XM_REQUIRE(loss && correct && out2 && tasks > 0, "xm_finish_shard: bad arguments");
  finish_shard_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(loss, correct, tasks, out2, step);
  return launched("xm_finish_shard");
}
