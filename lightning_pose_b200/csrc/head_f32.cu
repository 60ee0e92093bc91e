// fp32 heatmap head: PixelShuffle(2) -> ConvTranspose2d(k3,s2,p1,op1) [x2] -> spatial softmax (T=1).
// Reference: lightning_pose/models/heads/heatmap.py:20-71 (layer stack), :203-212 (forward).
//
// This is the full-precision path (reference training is fp32, lightning_pose/train.py:411-428):
// CUDA-core FMA with fp32 accumulation, gather-form transposed convolution (each thread owns one
// input pixel and produces its 2x2 output quad for a group of output channels, so there is no
// scatter / atomics), PixelShuffle folded into the shared-memory staging of the first layer.
// The bf16 tensor-core (tcgen05) path lives in head_bf16.cu.
#include <cstdint>

#include "../../include/lpb200.h"
#include "lpb_common.cuh"

namespace lpb {

constexpr int CT_CC = 8;        // input channels staged per iteration
constexpr int CT_OG = 5;        // groups of 4 output channels per pass (20 >= 17 keypoints)
constexpr int CT_KO = 4 * CT_OG;
constexpr int CT_MAX_THREADS = 256;

// out[o, 2m+py, 2n+px] = bias[o] + sum_c sum_taps in[c, m+dm, n+dn] * Wt[c, o, ky, kx]
//   py=0: (dm,ky)=(0,1);  py=1: (0,2),(1,0)     (y = 2*iy - 1 + ky), same along x.
// SHUFFLE: in[c, m, n] = feat[4c + 2(m&1) + (n&1), m>>1, n>>1]   (PixelShuffle(2))
template <bool SHUFFLE>
__global__ void __launch_bounds__(CT_MAX_THREADS) convt3x3s2_kernel(const float* __restrict__ in, int Cin, int Hi, int Wi,
                                                                    const float* __restrict__ wt,
                                                                    const float* __restrict__ bias, int Cout, int TR,
                                                                    float* __restrict__ out) {
  // Hi, Wi: spatial size of the (shuffled) conv input; output is (Cout, 2Hi, 2Wi)
  extern __shared__ __align__(16) float sm[];
  const int xs_row = Wi + 1;
  const int xs_plane = (TR + 1) * xs_row;
  float* xs = sm;                                   // [CT_CC][TR+1][Wi+1]
  float* ws = sm + ((CT_CC * xs_plane + 3) & ~3);   // [CT_CC][CT_KO*9], 16-B aligned rows (180 floats)
  const int tiles_per_img = (Hi + TR - 1) / TR;
  const int b = blockIdx.x / tiles_per_img;
  const int m0 = (blockIdx.x - b * tiles_per_img) * TR;
  const int o0 = blockIdx.y * CT_KO;
  const int nquad = TR * Wi;
  const int q = threadIdx.x;
  const int qm = q / Wi, qn = q - qm * Wi;
  const bool active = (q < nquad) && (m0 + qm < Hi);

  float acc[CT_KO][4];
#pragma unroll
  for (int o = 0; o < CT_KO; ++o) acc[o][0] = acc[o][1] = acc[o][2] = acc[o][3] = 0.f;

  const size_t in_img = SHUFFLE ? (size_t)b * (4 * Cin) * (Hi / 2) * (Wi / 2) : (size_t)b * Cin * Hi * Wi;
  for (int c0 = 0; c0 < Cin; c0 += CT_CC) {
    __syncthreads();
    // ---- stage inputs (rows m0 .. m0+TR, cols 0 .. Wi; zero halo) ----
    for (int i = threadIdx.x; i < CT_CC * xs_plane; i += blockDim.x) {
      const int cc = i / xs_plane, r = (i - cc * xs_plane) / xs_row, n = i - cc * xs_plane - r * xs_row;
      const int c = c0 + cc, m = m0 + r;
      float v = 0.f;
      if (c < Cin && m < Hi && n < Wi) {
        if (SHUFFLE) {
          const int ch = 4 * c + 2 * (m & 1) + (n & 1);
          v = __ldg(in + in_img + ((size_t)ch * (Hi / 2) + (m >> 1)) * (Wi / 2) + (n >> 1));
        } else {
          v = __ldg(in + in_img + ((size_t)c * Hi + m) * Wi + n);
        }
      }
      xs[i] = v;
    }
    // ---- stage weights for output channels o0 .. o0+CT_KO ----
    for (int i = threadIdx.x; i < CT_CC * CT_KO * 9; i += blockDim.x) {
      const int cc = i / (CT_KO * 9), r = i - cc * (CT_KO * 9);
      const int o = o0 + r / 9, c = c0 + cc;
      ws[i] = (c < Cin && o < Cout) ? __ldg(wt + ((size_t)c * Cout + o) * 9 + (r % 9)) : 0.f;
    }
    __syncthreads();
    if (active) {
#pragma unroll 2
      for (int cc = 0; cc < CT_CC; ++cc) {
        const float* xp = xs + cc * xs_plane + qm * xs_row + qn;
        const float x00 = xp[0], x01 = xp[1], x10 = xp[xs_row], x11 = xp[xs_row + 1];
        const float4* w4 = reinterpret_cast<const float4*>(ws + cc * (CT_KO * 9));
#pragma unroll
        for (int g = 0; g < CT_OG; ++g) {
          float w[36];
#pragma unroll
          for (int k = 0; k < 9; ++k) {
            const float4 t = w4[g * 9 + k];
            w[4 * k] = t.x;
            w[4 * k + 1] = t.y;
            w[4 * k + 2] = t.z;
            w[4 * k + 3] = t.w;
          }
#pragma unroll
          for (int oo = 0; oo < 4; ++oo) {
            const float* k9 = w + oo * 9;  // [ky][kx]
            float* a = acc[g * 4 + oo];
            a[0] = fmaf(x00, k9[4], a[0]);                                                      // (even, even)
            a[1] = fmaf(x00, k9[5], fmaf(x01, k9[3], a[1]));                                    // (even, odd)
            a[2] = fmaf(x00, k9[7], fmaf(x10, k9[1], a[2]));                                    // (odd, even)
            a[3] = fmaf(x00, k9[8], fmaf(x01, k9[6], fmaf(x10, k9[2], fmaf(x11, k9[0], a[3]))));  // (odd, odd)
          }
        }
      }
    }
  }
  if (!active) return;
  const int Ho = 2 * Hi, Wo = 2 * Wi;
  const int y = 2 * (m0 + qm), x = 2 * qn;
#pragma unroll
  for (int o = 0; o < CT_KO; ++o) {
    if (o0 + o >= Cout) break;
    const float bv = bias ? __ldg(bias + o0 + o) : 0.f;
    float* dst = out + (((size_t)b * Cout + o0 + o) * Ho + y) * Wo + x;
    *reinterpret_cast<float2*>(dst) = make_float2(acc[o][0] + bv, acc[o][1] + bv);
    *reinterpret_cast<float2*>(dst + Wo) = make_float2(acc[o][2] + bv, acc[o][3] + bv);
  }
}

// in-place softmax over each (b, k) plane: one HBM read + one write (plane staged in smem)
__global__ void __launch_bounds__(256) plane_softmax_kernel(float* __restrict__ x, int hw) {
  extern __shared__ float pl[];
  __shared__ float red[8];
  float* p = x + (size_t)blockIdx.x * hw;
  float mx = -3.0e38f;
  for (int i = threadIdx.x; i < hw; i += 256) {
    const float v = p[i];
    pl[i] = v;
    mx = fmaxf(mx, v);
  }
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int k = 1; k < 8; ++k) mx = fmaxf(mx, red[k]);
  float s = 0.f;
  for (int i = threadIdx.x; i < hw; i += 256) {
    const float e = expf(pl[i] - mx);
    pl[i] = e;
    s += e;
  }
  s = warp_sum(s);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += red[k];
  const float inv = 1.0f / s;
  for (int i = threadIdx.x; i < hw; i += 256) p[i] = pl[i] * inv;
}

static int launch_convt(bool shuffle, const float* in, int B, int Cin, int Hi, int Wi, const float* wt, const float* bias,
                        int Cout, float* out, cudaStream_t s) {
  int TR = CT_MAX_THREADS / Wi;
  LPB_REQUIRE(TR >= 1, "head: conv input width %d > %d unsupported", Wi, CT_MAX_THREADS);
  if (TR > Hi) TR = Hi;
  const int threads = ((TR * Wi + 31) / 32) * 32;
  const int tiles = (Hi + TR - 1) / TR;
  const size_t smem = (size_t)(((CT_CC * (TR + 1) * (Wi + 1) + 3) & ~3) + CT_CC * CT_KO * 9) * sizeof(float);
  dim3 grid((unsigned)(B * tiles), (unsigned)((Cout + CT_KO - 1) / CT_KO));
  if (shuffle)
    convt3x3s2_kernel<true><<<grid, threads, smem, s>>>(in, Cin, Hi, Wi, wt, bias, Cout, TR, out);
  else
    convt3x3s2_kernel<false><<<grid, threads, smem, s>>>(in, Cin, Hi, Wi, wt, bias, Cout, TR, out);
  LPB_CUDA(cudaGetLastError());
  return LPB_OK;
}

}  // namespace lpb

extern "C" int lpb_head_workspace_bytes(int B, int C, int H, int W, int c1, int c2, size_t* bytes) {
  using namespace lpb;
  LPB_REQUIRE(bytes, "head_workspace_bytes: null pointer");
  LPB_REQUIRE(B >= 0 && C >= 4 && C % 4 == 0 && H >= 1 && W >= 1 && c1 >= 1 && c2 >= 0, "head_workspace_bytes: bad shape");
  *bytes = c2 > 0 ? (size_t)B * c1 * (4 * H) * (4 * W) * sizeof(float) : 0;
  return LPB_OK;
}

extern "C" int lpb_head_fwd_f32(const float* features, int B, int C, int H, int W, const float* w1, const float* b1,
                                int c1, const float* w2, const float* b2, int c2, int final_softmax, float* out,
                                void* workspace, void* stream) {
  using namespace lpb;
  LPB_REQUIRE(features && w1 && out, "head_fwd_f32: null pointer");
  LPB_REQUIRE(B >= 0 && C >= 4 && C % 4 == 0 && H >= 1 && W >= 1 && c1 >= 1 && c2 >= 0, "head_fwd_f32: bad shape");
  LPB_REQUIRE(c2 == 0 || (w2 && workspace), "head_fwd_f32: two-layer head needs w2 and workspace");
  if (B == 0) return LPB_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int Hs = 2 * H, Ws = 2 * W;  // after PixelShuffle(2)
  int K, Ho, Wo;
  if (c2 == 0) {
    int rc = launch_convt(true, features, B, C / 4, Hs, Ws, w1, b1, c1, out, s);
    if (rc) return rc;
    K = c1;
    Ho = 2 * Hs;
    Wo = 2 * Ws;
  } else {
    float* mid = static_cast<float*>(workspace);
    int rc = launch_convt(true, features, B, C / 4, Hs, Ws, w1, b1, c1, mid, s);
    if (rc) return rc;
    rc = launch_convt(false, mid, B, c1, 2 * Hs, 2 * Ws, w2, b2, c2, out, s);
    if (rc) return rc;
    K = c2;
    Ho = 4 * Hs;
    Wo = 4 * Ws;
  }
  if (final_softmax) {
    const int hw = Ho * Wo;
    const size_t smem = (size_t)hw * sizeof(float);
    LPB_REQUIRE(smem <= 200 * 1024, "head_fwd_f32: plane %dx%d too large for the softmax stage", Ho, Wo);
    if (smem > 48 * 1024)
      LPB_CUDA(cudaFuncSetAttribute(plane_softmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    plane_softmax_kernel<<<(unsigned)((size_t)B * K), 256, smem, s>>>(out, hw);
    LPB_CUDA(cudaGetLastError());
  }
  return LPB_OK;
}
