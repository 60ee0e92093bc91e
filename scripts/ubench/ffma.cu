// micro-benchmark: issue rate of FFMA (3-register), FFMA2 (fma.rn.f32x2) and MUFU.EX2 on one B200
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma ffma.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  uint64_t ra = *reinterpret_cast<uint64_t*>(&a), rb = *reinterpret_cast<uint64_t*>(&b), rc = *reinterpret_cast<uint64_t*>(&c), rd;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}

template <int MODE>
__global__ void __launch_bounds__(256) k(const float* in, float* out, int iters) {
  float x = in[threadIdx.x & 31], y = in[32 + (threadIdx.x & 31)];
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = x * (float)(i + 1);
  float2 b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) b[i] = make_float2(a[i], a[i] + 1.f);
  const float2 xy = make_float2(x, y), yy = make_float2(y, x);
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], x, y);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) b[i] = ffma2(b[i], xy, yy);
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
    } else {  // 4 FFMA : 1 MUFU mix
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], x, y);
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[0]));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[4]));
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + b[i].x + b[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int per_iter_lane_ops) {
  float *in, *out;
  cudaMalloc(&in, 256);
  cudaMemset(in, 0, 256);
  const int sms = 148, blocks = sms * 8, iters = 20000;
  cudaMalloc(&out, blocks * 256 * 4);
  k<MODE><<<blocks, 256>>>(in, out, 100);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<blocks, 256>>>(in, out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double ops = (double)blocks * 256 * iters * per_iter_lane_ops;
  printf("%-28s %8.3f ms  %7.1f lane-ops/clk/SM (at 1.965 GHz)\n", name, ms, ops / (ms * 1e-3) / 1.965e9 / sms);
}

int main() {
  run<0>("FFMA (3-reg)", 8);
  run<1>("FFMA2 (f32x2), FMAs counted", 16);
  run<2>("MUFU.EX2", 8);
  run<3>("8 FFMA + 2 MUFU, all ops", 10);
  return 0;
}
