// One launch for everything a head call has to prepare: operand packing of the weights (fp32 master -> bf16 UMMA images),
// clearing the pad rows of freshly allocated row-layout buffers, zeroing gradient accumulators.  Eight tiny launches per
// forward + backward pair became two (each used to cost a launch latency and a tail of its own).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "row_layout.cuh"

namespace lpb {

struct PrepJobs {
  // forward operand packs: W[Cin][Cout][3][3] -> B[stage][shift][kchunk][80][8]   (head_bf16.cu)
  struct { const float* w; const float* bias; int Cin, Cout, nstages; __nv_bfloat16* out; } fpack[2];
  // data-gradient operand packs: -> [tile][shift][kchunk][rows_per_tile][8]            (head_bwd_bf16.cu)
  struct { const float* w; int Cin, Cout, ntiles, rows_per_tile; __nv_bfloat16* out; } dpack[2];
  struct { __nv_bfloat16* buf; RowLayout L; long long nslabs; } pads[2];
  struct { float* p; long long n; } zero[4];
};

int launch_head_prep(const PrepJobs& jobs, cudaStream_t s);

}  // namespace lpb
