// probe_tmem_shift.cu -- what does tcgen05.shift.down do on sm_100a?  128 threads fill 64 TMEM columns with
// value(row, col) = row * 100 + col through tcgen05.st, one thread issues the shifts of a case + tcgen05.commit, everybody reads
// the columns back.  Prints, per case, which (row, col) each element came from and the cycles from first shift to barrier.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I exploring_meta_b200/csrc scripts/probe_tmem_shift.cu -o scripts/probe_tmem_shift.bin
#include <cstdio>
#include <vector>
#include "tc.cuh"
using namespace xm;

struct Case { int nshift; int col[16]; int lane[16]; };

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}

__global__ void probe(Case c, float* out, long long* cyc) {
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  const uint32_t mine = tm + ((uint32_t)(warp * 32) << 16);
  for (int half = 0; half < 2; ++half) {
    uint32_t r[32];
    for (int k = 0; k < 32; ++k) r[k] = __float_as_uint((float)(tid * 100 + half * 32 + k));
    tmem_st32(mine + half * 32, r);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  long long t0 = 0;
  if (tid == 0) {
    t0 = clock64();
    for (int i = 0; i < c.nshift; ++i) {
      const uint32_t ta = tm + ((uint32_t)c.lane[i] << 16) + (uint32_t)c.col[i];
      asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(ta) : "memory");
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  if (tid == 0) *cyc = clock64() - t0;
  tc_fence_after();
  for (int half = 0; half < 2; ++half) {
    float v[32];
    tmem_ld32(mine + half * 32, v);
    for (int k = 0; k < 32; ++k) out[tid * 64 + half * 32 + k] = v[k];
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(64) : "memory");
}

int main() {
  std::vector<Case> cases;
  { Case c{}; c.nshift = 1; c.col[0] = 0; c.lane[0] = 0; cases.push_back(c); }
  { Case c{}; c.nshift = 2; c.col[0] = c.col[1] = 0; cases.push_back(c); }
  { Case c{}; c.nshift = 1; c.col[0] = 8; cases.push_back(c); }
  { Case c{}; c.nshift = 1; c.col[0] = 4; cases.push_back(c); }
  { Case c{}; c.nshift = 1; c.col[0] = 16; c.lane[0] = 32; cases.push_back(c); }
  { Case c{}; c.nshift = 12; for (int i = 0; i < 12; ++i) { c.col[i] = (i % 8) * 8; } cases.push_back(c); }
  float* d; long long* dc;
  cudaMalloc(&d, 128 * 64 * 4); cudaMalloc(&dc, 8);
  std::vector<float> h(128 * 64);
  for (size_t ci = 0; ci < cases.size(); ++ci) {
    const Case& c = cases[ci];
    cudaMemset(d, 0, 128 * 64 * 4);
    probe<<<1, 128>>>(c, d, dc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("case %zu: %s\n", ci, cudaGetErrorString(e)); return 1; }
    long long cyc; cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(h.data(), d, 128 * 64 * 4, cudaMemcpyDeviceToHost);
    printf("case %zu: %d shifts at", ci, c.nshift);
    for (int i = 0; i < c.nshift; ++i) printf(" (lane %d, col %d)", c.lane[i], c.col[i]);
    printf("; %lld cycles from first shift to barrier\n", cyc);
    { const int col = c.col[0];
      printf("  rows 28..34, 60..66, 92..98, 124..127 of col %d:", col);
      for (int row : {28,29,30,31,32,33,34,60,61,62,63,64,65,66,92,93,94,95,96,97,98,124,125,126,127}) printf(" %d", (int)h[row * 64 + col]);
      printf("\n"); }
    // summarise: for each column, the row offset (source row - row) over rows, and which rows are unchanged
    for (int col = 0; col < 64; ++col) {
      int moved = 0, first = -1, last = -1, delta = 0, colchg = 0;
      for (int row = 0; row < 128; ++row) {
        const int v = (int)h[row * 64 + col], srow = v / 100, scol = v % 100;
        if (srow != row || scol != col) {
          if (!moved) { first = row; delta = srow - row; }
          ++moved; last = row;
          if (scol != col) colchg = 1;
        }
      }
      if (moved) printf("  col %2d: %3d rows changed (rows %d..%d), source row - row = %d%s; row0 = %d row1 = %d row127 = %d\n", col, moved, first, last,
                        delta, colchg ? " COLUMN CHANGED" : "", (int)h[0 * 64 + col], (int)h[1 * 64 + col], (int)h[127 * 64 + col]);
    }
  }
  printf("done\n");
  return 0;
}
