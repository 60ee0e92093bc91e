// Adam / AdamW step over a short list of tensors in ONE launch.
// Reference: torch.optim.Adam / AdamW as configured by BaseFeatureExtractor.configure_optimizers
// (lightning_pose/models/base.py:458-477; defaults lr = 1e-3, betas (0.9, 0.999), eps 1e-8, no amsgrad).
// Why it exists: the head has four parameter tensors (80,971 scalars); torch's multi-tensor kernel handles 65,536
// elements per block, so its launch is 5 blocks and 38 us long at the very end of the step, where nothing can
// overlap it.  Here every 256 elements are a block (~3 us).  The step counter lives on the device (a captured CUDA
// graph replays the same launch): every block reads it, the last block to finish advances it.
#include <cstdint>

#include "../../include/lpb200.h"
#include "lpb_common.cuh"

namespace lpb {

constexpr int ADAM_MAX_TENSORS = 16;

struct AdamJobs {
  float* p[ADAM_MAX_TENSORS];
  const float* g[ADAM_MAX_TENSORS];
  float* m[ADAM_MAX_TENSORS];
  float* v[ADAM_MAX_TENSORS];
  long long end[ADAM_MAX_TENSORS];  // exclusive prefix ends, in elements
  int n;
  float* step;        // device scalar: number of steps taken so far
  unsigned* counter;  // device scalar: blocks finished (self-resetting)
  const float* lr_dev;
  double beta1, beta2;  // the bias corrections 1 - beta^t lose five digits in fp32 (beta2 = 0.999): computed in fp64, once per block
  float lr, eps, weight_decay;
  int decoupled;      // 1: AdamW (p *= 1 - lr * wd), 0: Adam (g += wd * p)
};

__global__ void __launch_bounds__(256) adam_step_kernel(const __grid_constant__ AdamJobs J) {
  const float t = *J.step + 1.0f;
  const float lr = J.lr_dev ? *J.lr_dev : J.lr;
  __shared__ float s_bc[2];
  if (threadIdx.x == 0) {
    s_bc[0] = (float)(1.0 - pow(J.beta1, (double)t));        // bias_correction1
    s_bc[1] = (float)sqrt(1.0 - pow(J.beta2, (double)t));  // sqrt(bias_correction2)
  }
  __syncthreads();
  const float b1 = (float)J.beta1, b2 = (float)J.beta2, omb1 = (float)(1.0 - J.beta1), omb2 = (float)(1.0 - J.beta2);
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i < J.end[J.n - 1]) {
    int k = 0;
    while (i >= J.end[k]) ++k;
    const long long j = i - (k ? J.end[k - 1] : 0);
    float p = J.p[k][j], g = J.g[k][j], m = J.m[k][j], v = J.v[k][j];
    if (J.weight_decay != 0.f) {
      if (J.decoupled) p *= 1.0f - lr * J.weight_decay;
      else g = fmaf(J.weight_decay, p, g);
    }
    m = fmaf(b1, m, omb1 * g);       // exp_avg.lerp_(grad, 1 - beta1)
    v = fmaf(b2, v, omb2 * g * g);   // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(v) / s_bc[1] + J.eps;
    p -= (lr / s_bc[0]) * (m / denom);
    J.p[k][j] = p;
    J.m[k][j] = m;
    J.v[k][j] = v;
  }
  // every block has read *step before it arrives here; the last one to arrive publishes step + 1
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(J.counter, 1u) == gridDim.x - 1) {
      *J.step = t;
      *J.counter = 0u;
    }
  }
}

}  // namespace lpb

extern "C" int lpb_adam_step(int n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                             float* const* exp_avg_sq, const int64_t* numel, float* step, uint32_t* block_counter,
                             float lr, const float* lr_dev, double beta1, double beta2, float eps, float weight_decay,
                             int decoupled, void* stream) {
  using namespace lpb;
  LPB_REQUIRE(n_tensors >= 1 && n_tensors <= ADAM_MAX_TENSORS, "adam_step: 1..%d tensors per call (got %d)", ADAM_MAX_TENSORS, n_tensors);
  LPB_REQUIRE(params && grads && exp_avg && exp_avg_sq && numel && step && block_counter, "adam_step: null pointer");
  LPB_REQUIRE(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 && eps >= 0.f, "adam_step: bad hyper-parameters");
  AdamJobs J{};
  long long total = 0;
  for (int k = 0; k < n_tensors; ++k) {
    LPB_REQUIRE(params[k] && grads[k] && exp_avg[k] && exp_avg_sq[k] && numel[k] >= 0, "adam_step: tensor %d: null pointer or bad size", k);
    J.p[k] = params[k];
    J.g[k] = grads[k];
    J.m[k] = exp_avg[k];
    J.v[k] = exp_avg_sq[k];
    total += numel[k];
    J.end[k] = total;
  }
  J.n = n_tensors;
  J.step = step;
  J.counter = block_counter;
  J.lr_dev = lr_dev;
  J.lr = lr;
  J.beta1 = beta1;
  J.beta2 = beta2;
  J.eps = eps;
  J.weight_decay = weight_decay;
  J.decoupled = decoupled;
  const long long blocks = total > 0 ? (total + 255) / 256 : 1;
  LPB_REQUIRE(blocks < (1ll << 31), "adam_step: too many elements");
  adam_step_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(J);
  LPB_CUDA(cudaGetLastError());
  return LPB_OK;
}
