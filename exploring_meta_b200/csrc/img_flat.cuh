// img_flat.cuh -- kernel argument block shared by the two closed-form image-block variants (img_block.cu: stride 1 +
// 2x2 pool, Mini-ImageNet; img_flat.cu: stride 2 without pool, Omniglot) and the entry points of the latter.
#pragma once
#include "common.cuh"

namespace xm {

struct ImgK {
  int n, H, W, cout, splits, hp, wp;
  int row0, row_step, rows_per_task;
  double cnt;
  float eps, scale;
  const float* x; double* gram;
  const float* w; long long wstride;
  const float* wd; long long wdstride;
  const float* gamma; const float* beta; long long gbstride;
  const float* gammad; const float* betad; long long gbdstride;
  float* mean_invstd; float* call_stats; float* bwd_red; float* dual_red;
  float* p; float* zsel; unsigned char* sel; float* pdot; float* zdsel;
  const float* gp; const float* gpd;
  double* ssum; double* scratch;
  float* out_w; float* out_b; float* out_gamma; float* out_beta; long long ostride;
  const float* base_w; const float* base_b; const float* base_gamma; const float* base_beta; long long bstride;
};

// img_flat.cu
int flat_ok(const XmBlockGeom& g);
int flat_launch_gram(const XmImgArgs* a, ImgK& k, cudaStream_t stream);
int flat_launch(int mode, const XmImgArgs* a, ImgK& k, cudaStream_t stream);   // 0 fwd, 1 bwd sums, 2 dual fwd, 3 dual bwd sums

}  // namespace xm
