// Shape-generic building blocks of the bf16 head (head_rows_bf16.cu), shared with the C-ABI entry in head_bf16.cu.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "row_layout.cuh"

namespace lpb {

enum { CONVT_ROWS_MID = 0, CONVT_ROWS_PLANES = 1, CONVT_ROWS_SOFTMAX = 2 };

struct ConvtRowsParams {
  const __nv_bfloat16* X;    // [B][4*nst][L.rows][8] padded row layout of the conv input
  RowLayout L;
  const __nv_bfloat16* wpk;  // [nst][4 shifts][4 kchunks][80][8] (head_prep_kernel)
  const float* bias;         // [cout] added in the epilogue, or null (bias folded into the GEMM / irrelevant)
  int nst;                   // 32-channel K stages
  int B, cout, mode;
  __nv_bfloat16* mid;        // CONVT_ROWS_MID: [B][4][Lout.rows][8], channel `cout` = 1
  RowLayout Lout;
  float* out;                // planes [B][cout][2Hi][2Wi] (raw or softmaxed)
  int R, rows_alloc, backoff;  // filled by launch_convt_rows
};

int launch_rows_shuffle(const __nv_bfloat16* feat, int B, int C, int H, int W, __nv_bfloat16* xs, cudaStream_t s);
int launch_convt_rows(ConvtRowsParams p, int sms, cudaStream_t s);

}  // namespace lpb
