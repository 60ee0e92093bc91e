// Host-side construction of the decode's 1-D upsampling operator and its device tables.
//
// `upsample^ds(h) == U_H h U_W^T` with a banded U (n*2^ds x n): one stage = 2x bicubic
// (A=-0.75, align_corners=False, edge-clamped taps) followed by the zero-padded [1,4,6,4,1]/16
// blur.  Reference: lightning_pose/models/heads/heatmap.py:86-100.
#pragma once
#include <vector>

namespace lpb {

template <int DS>
struct UpsampleGeom {
  static constexpr int F = 1 << DS;   // fine samples per coarse sample
  static constexpr int R = DS + 2;    // band radius in coarse samples (verified on build)
  static constexpr int W = 2 * R + 1; // taps per fine row in window form
};

// Window form: row i (fine), tap t  ->  coarse index i/F - R + t (weight 0 outside [0,n)).
struct HostTable {
  int n = 0, ds = 0;
  std::vector<float> win;    // [n*F][W]
  std::vector<float> phase;  // [F][W] interior (border-free) rows, identical for every n >= 2R+1
  float lip = 0.f;           // max_i sum_c |U[i,c]|
};

// Returns false (and sets the error string) if the band/periodicity assumptions do not hold.
bool build_host_table(int n, int ds, HostTable* out);

struct DeviceTable {
  const float* win = nullptr;  // device copy of HostTable::win
  HostTable host;
};

// Cached per (device, n, ds). Allocates + uploads on first use (synchronous cudaMemcpy).
const DeviceTable* get_device_table(int n, int ds);

}  // namespace lpb
