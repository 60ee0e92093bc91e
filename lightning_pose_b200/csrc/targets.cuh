// Device helpers shared by target generation and the fused target+loss kernel.
// Rules restated from lightning_pose/data/heatmaps.py:37-87.
#pragma once
#include "lpb_common.cuh"

namespace lpb {

enum : int { TARGET_ZERO = 0, TARGET_UNIFORM = 1, TARGET_GAUSS = 2 };

struct TargetPlane {
  int mode;
  float x, y;  // clamped heatmap-grid coordinates (NaN propagates, as torch.clamp does)
};

// vis < 0 means "visibility is None".
__device__ __forceinline__ TargetPlane classify_target(float xr, float yr, int vis, float sx, float sy, int oh, int ow) {
  TargetPlane t;
  const float x = xr * sx, y = yr * sy;  // :41-42
  // :43-49 - only x is tested for NaN; NaN comparisons are false
  const bool bad = isnan(x) || (x < -1.f) || (x > (float)(ow + 1)) || (y < -1.f) || (y > (float)(oh + 1));
  t.x = isnan(x) ? x : fminf(fmaxf(x, -1.f), (float)(ow + 1));  // :52-53
  t.y = isnan(y) ? y : fminf(fmaxf(y, -1.f), (float)(oh + 1));
  if (vis < 0) {
    t.mode = bad ? TARGET_ZERO : TARGET_GAUSS;  // :78-79
  } else if (vis == 0) {
    t.mode = TARGET_ZERO;  // :81
  } else if (vis == 1) {
    t.mode = TARGET_UNIFORM;  // :82
  } else {
    t.mode = (vis == 2 && bad) ? TARGET_ZERO : TARGET_GAUSS;  // :83
  }
  return t;
}

// Fills ex[0..ow), ey[0..oh) with the separable Gaussian factors and returns 1/(sum ex * sum ey).
// `red` needs 2 * (nthreads/32) floats.  All threads of the block must call; ends synchronised.
__device__ __forceinline__ float target_axis_factors(const TargetPlane& tp, int oh, int ow, float two_s2, float* ex,
                                                     float* ey, float* red, int nthreads) {
  float px = 0.f, py = 0.f;
  for (int c = threadIdx.x; c < ow; c += nthreads) {
    const float d = (float)c - tp.x;
    const float e = expf(-(d * d) / two_s2);
    ex[c] = e;
    px += e;
  }
  for (int r = threadIdx.x; r < oh; r += nthreads) {
    const float d = (float)r - tp.y;
    const float e = expf(-(d * d) / two_s2);
    ey[r] = e;
    py += e;
  }
  px = warp_sum(px);
  py = warp_sum(py);
  const int warp = threadIdx.x >> 5, nw = nthreads >> 5;
  if ((threadIdx.x & 31) == 0) {
    red[warp] = px;
    red[nw + warp] = py;
  }
  __syncthreads();
  float sxs = 0.f, sys = 0.f;
  for (int k = 0; k < nw; ++k) {
    sxs += red[k];
    sys += red[nw + k];
  }
  return 1.0f / (sxs * sys);
}

}  // namespace lpb
