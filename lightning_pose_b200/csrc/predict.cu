// Batched-inference plumbing: decoded keypoints of one chunk -> rows of the preallocated (N, 3K)
// prediction table, on the device, at a device-resident row cursor, so the whole chunk (head -> decode -> remap ->
// table rows -> cursor += T) is one CUDA-graph replay with no host-side offset.
// Reference: PredictionHandler.unpack_preds / make_pred_arr_undo_resize  lightning_pose/utils/predictions.py:97-144,180-206
// (torch.vstack of per-batch tuples on the host, then numpy interleaving into bp_x, bp_y, bp_likelihood columns).
#include <cstdint>

#include "../../include/lpb200.h"
#include "lpb_common.cuh"

namespace lpb {

__global__ void pack_predictions_kernel(const float* __restrict__ kp, const float* __restrict__ conf, int n, int K,
                                        float* __restrict__ table, int64_t n_rows, const int64_t* __restrict__ cursor,
                                        int64_t row0) {
  const int64_t base = cursor ? *cursor : row0;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * K) return;
  const int f = i / K, k = i - f * K;
  const int64_t row = base + f;
  if (row < 0 || row >= n_rows) return;  // rows past the end of the video (last, partly filled chunk) are dropped
  float* dst = table + (row * K + k) * 3;
  dst[0] = kp[(size_t)f * 2 * K + 2 * k];
  dst[1] = kp[(size_t)f * 2 * K + 2 * k + 1];
  dst[2] = conf[(size_t)f * K + k];
}

__global__ void advance_cursor_kernel(int64_t* cursor, int64_t n) { *cursor += n; }

}  // namespace lpb

extern "C" int lpb_pack_predictions(const float* keypoints, const float* confidences, int n_frames, int K, float* table,
                                    int64_t n_rows, int64_t* cursor, int64_t row0, void* stream) {
  using namespace lpb;
  LPB_REQUIRE(keypoints && confidences && table && n_frames >= 0 && K >= 1 && n_rows >= 0, "pack_predictions: bad arguments");
  if (n_frames == 0) return LPB_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int total = n_frames * K;
  pack_predictions_kernel<<<(total + 255) / 256, 256, 0, s>>>(keypoints, confidences, n_frames, K, table, n_rows, cursor, row0);
  if (cursor) advance_cursor_kernel<<<1, 1, 0, s>>>(cursor, n_frames);
  LPB_CUDA(cudaGetLastError());
  return LPB_OK;
}
