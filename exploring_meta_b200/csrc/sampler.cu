// sampler.cu -- on-device few-shot task sampler: replaces train_tasks.sample() +
// prepare_batch's host work (utils/data_pre.py:16-129) for a dataset that is resident in HBM as uint8.
//
// A task is what learn2learn's transform chain NWays -> KShots(2k) -> LoadData -> RemapLabels -> ConsecutiveLabels
// (-> RandomClassRotation for Omniglot) yields (utils/data_pre.py:28-37, 79-85): `ways` distinct classes, 2k distinct
// items of each, samples grouped by class, labels remapped to 0..ways-1, and (Omniglot) one rotation out of
// {0, 90, 180, 270} degrees per class of the task.  The pixel transform is out = scale * u8 + offset
// (Omniglot: ToTensor then 1 - x -> scale -1/255, offset 1, data_pre.py:18-22; Mini-ImageNet: raw 0..255 floats).
//
// Randomness is counter based -- splitmix64 of (seed, global task number, draw counter), reduced to a range by
// multiply-shift -- so a batch is a pure function of (seed, first_task): every rank can draw its own shard without
// communication, and oracle/task_sampler_oracle.py reproduces every index bit for bit.
// Distinctness is by sequential rejection (ways <= 64, 2k <= 64: a few hundred draws by one thread per task).
#include "common.cuh"

namespace xm {

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}
// draw number `ctr` of task `task`: uniform integer in [0, n)
__host__ __device__ __forceinline__ uint32_t draw(uint64_t seed, uint64_t task, uint32_t ctr, uint32_t n) {
  const uint64_t h = splitmix64(seed ^ (task * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)ctr * 0xD1B54A32D192ED03ull));
  return (uint32_t)(((h >> 32) * (uint64_t)n) >> 32);
}

constexpr int SMP_MAX = 64;

__global__ void __launch_bounds__(256) sample_tasks_kernel(const XmSampleArgs a) {
  __shared__ int s_class[SMP_MAX], s_rot[SMP_MAX];
  __shared__ int s_item[SMP_MAX * SMP_MAX];
  const int t = blockIdx.x, tid = threadIdx.x, part = blockIdx.y, nparts = gridDim.y;
  const uint64_t gtask = (uint64_t)a.first_task + (uint64_t)t;
  const int per = a.ways * a.shots2;
  if (tid == 0) {
    uint32_t ctr = 0;
    for (int i = 0; i < a.ways; ++i) {                 // NWays: distinct classes, in draw order = label order
      int c;
      bool dup;
      do {
        c = (int)draw(a.seed, gtask, ctr++, (uint32_t)a.num_classes);
        dup = false;
        for (int j = 0; j < i; ++j) dup |= s_class[j] == c;
      } while (dup);
      s_class[i] = c;
    }
    for (int i = 0; i < a.ways; ++i) {                 // KShots(2k): distinct items of the class
      const int lo = a.class_start[s_class[i]], cnt = a.class_start[s_class[i] + 1] - lo;
      for (int k = 0; k < a.shots2; ++k) {
        int it;
        bool dup;
        do {
          it = lo + (int)draw(a.seed, gtask, ctr++, (uint32_t)cnt);
          dup = false;
          for (int j = 0; j < k; ++j) dup |= s_item[i * a.shots2 + j] == it;
        } while (dup);
        s_item[i * a.shots2 + k] = it;
      }
    }
    for (int i = 0; i < a.ways; ++i)                   // RandomClassRotation: quarter turns, one per class
      s_rot[i] = a.rotate ? (int)draw(a.seed, gtask, ctr++, 4u) : 0;
  }
  __syncthreads();
  if (part == 0) {
    for (int i = tid; i < per; i += blockDim.x) {
      a.y[(long long)t * per + i] = i / a.shots2;      // RemapLabels + ConsecutiveLabels
      if (a.items) a.items[(long long)t * per + i] = s_item[i];
    }
    if (a.classes)
      for (int i = tid; i < a.ways; i += blockDim.x) a.classes[(long long)t * a.ways + i] = s_class[i];
  }
  // the CTAs of a task (every one repeats the few hundred index draws above) split its samples
  const int H = a.height, W = a.width, hw = H * W, chw = a.channels * hw;
  float* X = a.x + (long long)t * per * chw;
  for (int s = part; s < per; s += nparts) {
    const unsigned char* src = a.data + (long long)s_item[s] * chw;
    float* dst = X + (long long)s * chw;
    const int q = s_rot[s / a.shots2];
    if (q == 0 && (chw & 3) == 0) {                    // no rotation: 4 bytes in, one 16-byte store out
      for (int e = tid; e < chw / 4; e += blockDim.x) {
        const uchar4 v = __ldg(reinterpret_cast<const uchar4*>(src) + e);
        reinterpret_cast<float4*>(dst)[e] = make_float4(fmaf(a.scale, (float)v.x, a.offset), fmaf(a.scale, (float)v.y, a.offset),
                                                        fmaf(a.scale, (float)v.z, a.offset), fmaf(a.scale, (float)v.w, a.offset));
      }
    } else {
      for (int r = tid; r < chw; r += blockDim.x) {
        const int c = r / hw, p = r - c * hw, y = p / W, x = p - y * W;
        // counter-clockwise quarter turns (numpy rot90): out[y][x] = in[sy][sx]
        int sy = y, sx = x;
        if (q == 1) { sy = x; sx = W - 1 - y; }
        else if (q == 2) { sy = H - 1 - y; sx = W - 1 - x; }
        else if (q == 3) { sy = H - 1 - x; sx = y; }
        dst[r] = fmaf(a.scale, (float)__ldg(src + c * hw + sy * W + sx), a.offset);
      }
    }
  }
}

}  // namespace xm

using namespace xm;

extern "C" int xm_sample_tasks(const XmSampleArgs* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  XM_REQUIRE(a != nullptr, "xm_sample_tasks: null args");
  XM_REQUIRE(a->tasks > 0 && a->ways > 0 && a->shots2 > 0 && a->channels > 0 && a->height > 0 && a->width > 0,
             "xm_sample_tasks: bad sizes");
  XM_REQUIRE(a->ways <= SMP_MAX && a->shots2 <= SMP_MAX, "xm_sample_tasks: ways and 2*shots must be <= %d", SMP_MAX);
  XM_REQUIRE(a->num_classes >= a->ways, "xm_sample_tasks: fewer classes (%d) than ways (%d)", a->num_classes, a->ways);
  XM_REQUIRE(!a->rotate || a->height == a->width, "xm_sample_tasks: rotation needs square images");
  XM_REQUIRE(a->data && a->class_start && a->x && a->y, "xm_sample_tasks: null data/class_start/x/y");
  // items must be 4-byte aligned for the vector path: chw % 4 == 0 makes every item start aligned when data is
  int parts = (4 * num_sms() + a->tasks - 1) / a->tasks;
  const int per = a->ways * a->shots2;
  if (parts > per) parts = per;
  if (parts < 1) parts = 1;
  sample_tasks_kernel<<<dim3(a->tasks, parts), 256, 0, stream>>>(*a);
  return launched("xm_sample_tasks");
}
