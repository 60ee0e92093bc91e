// bf16 heatmap head on the 5th-gen tensor cores (tcgen05 + TMEM), two-deconv (ResNet-family) heads.
// Reference: lightning_pose/models/heads/heatmap.py:20-71 (layer stack), :203-212 (forward).
//
// Formulation ("4-shift GEMM").  A stride-2 3x3 transposed convolution in gather form:
//   out[o, 2m+py, 2n+px] = b[o] + sum_c sum_{(dm,dn) valid for (py,px)} in[c, m+dm, n+dn] * W[c, o, ky, kx]
//   with ky = (py==0 ? 1 : (dm ? 0 : 2)), kx likewise.
// Lay the input out pixel-major, A[row = m*(Wi+1) + n][c] with a zero column n = Wi and a zero row
// m = Hi; then the four (dm,dn) shifts are the SAME matrix read from a start address moved by
// (dm*(Wi+1) + dn) rows, and
//   D[row, cls*20 + o] = sum_shift A_shift[row, :] . B_shift[cls*20 + o, :]
// is an ordinary GEMM accumulated in TMEM over 4 shifts x K.  D needs no col2im: row (m,n), class
// (py,px) IS output pixel (2m+py, 2n+px).  Operands use the K-major no-swizzle UMMA layout with
// SBO = 128 B, where rows are linear in memory, so a row shift is just a different descriptor start
// address (tcgen05.cuh).
//
//   k1a: features (NCHW bf16) --[PixelShuffle folded into a transposing producer]--> smem A stages
//        --tcgen05.mma--> TMEM --epilogue(+bias, ->bf16)--> mid activations in the A layout (global)
//   k1b: mid (A layout, TMA bulk rows) --tcgen05.mma--> TMEM --epilogue--> two-pass plane softmax
//        (pass 0: online max/sum, pass 1: recompute the cheap K=32 GEMM and write normalised fp32)
#include <cuda_bf16.h>

#include <cstdint>

#include "../../include/lpb200.h"
#include "lpb_common.cuh"
#include "tcgen05.cuh"

namespace lpb {

constexpr int HB_THREADS = 288;      // warps 0-3 producer, warp 4 MMA issuer, warps 5-8 epilogue
constexpr int HB_NCOLS = 80;         // 4 classes x 20 (>= 17 keypoints), multiple of 16
constexpr int HB_CLS = 20;
constexpr int HB_KSTAGE = 32;        // channels per pipeline stage (4 K-chunks of 8)
constexpr int HB_BSTAGE_BYTES = 4 * 4 * HB_NCOLS * 16;  // [shift][kchunk][80 rows][16 B]
constexpr int HB_STAGES = 3;

struct HeadGeom {
  int Hi, Wi;       // conv input spatial size (after PixelShuffle for layer 1)
  int P;            // row pitch = Wi + 1 (zero column)
  int rows;         // Hi * P
  int tiles;        // ceil(rows / 128)
  int rows_alloc;   // smem rows per K-chunk (multiple of 8), covers tiles*128 + P + 1
};

__host__ inline HeadGeom make_geom(int Hi, int Wi) {
  HeadGeom g;
  g.Hi = Hi;
  g.Wi = Wi;
  g.P = Wi + 1;
  g.rows = Hi * g.P;
  g.tiles = (g.rows + 127) / 128;
  g.rows_alloc = (g.tiles * 128 + g.P + 1 + 7) & ~7;
  return g;
}

// ---- weight packing: W[Cin][Cout][3][3] (fp32) -> B[stage][shift][kchunk][80][8] bf16 -----------------
__global__ void pack_convt_weights_kernel(const float* __restrict__ w, int Cin, int Cout, int nstages,
                                          __nv_bfloat16* __restrict__ out) {
  const int total = nstages * 4 * 4 * HB_NCOLS * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int e = i & 7;
    int r = i >> 3;
    const int nrow = r % HB_NCOLS;
    r /= HB_NCOLS;
    const int kc = r & 3;
    r >>= 2;
    const int sh = r & 3;
    const int st = r >> 2;
    const int c = st * HB_KSTAGE + kc * 8 + e;
    const int cls = nrow / HB_CLS, o = nrow % HB_CLS;
    const int py = cls >> 1, px = cls & 1, dm = sh >> 1, dn = sh & 1;
    float v = 0.f;
    if (c < Cin && o < Cout && !(py == 0 && dm == 1) && !(px == 0 && dn == 1)) {
      const int ky = py == 0 ? 1 : (dm ? 0 : 2);
      const int kx = px == 0 ? 1 : (dn ? 0 : 2);
      v = w[((size_t)c * Cout + o) * 9 + ky * 3 + kx];
    }
    out[i] = __float2bfloat16_rn(v);
  }
}

// =====================================================================================================
// k1a: PixelShuffle + first transposed convolution
// =====================================================================================================
struct K1aParams {
  const __nv_bfloat16* feat;  // [B][C][H*W]
  const __nv_bfloat16* wpk;   // packed weights [nstages][HB_BSTAGE_BYTES]
  const float* bias;          // [c1]
  __nv_bfloat16* mid;         // [B][4][4*Hi*Wi][8]  (A layout of the next layer, no halo)
  int B, C, HW, W;            // feature geometry (C = 4 * Cin)
  int c1, nstages;
  HeadGeom g;
};

__global__ void __launch_bounds__(HB_THREADS, 1) k1a_shuffle_convt_kernel(const __grid_constant__ K1aParams P) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const HeadGeom g = P.g;
  const int a_stage_bytes = 4 * g.rows_alloc * 16;
  const int stage_bytes = a_stage_bytes + HB_BSTAGE_BYTES;
  unsigned char* stage_base = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + HB_STAGES * stage_bytes);
  uint64_t* full = bars;                    // [HB_STAGES]
  uint64_t* empty = bars + HB_STAGES;       // [HB_STAGES]
  uint64_t* tmem_full = bars + 2 * HB_STAGES;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // zero the A stages once: halo rows/columns are never written again
  for (int i = tid; i < HB_STAGES * stage_bytes / 16; i += HB_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int s = 0; s < HB_STAGES; ++s) {
      mbar_init(&full[s], 129);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 128);
    fence_mbar_init();
  }
  if (warp == 4) tc::tmem_alloc(tmem_ptr, 512);
  fence_proxy_async();  // generic-proxy zero fill -> visible to the async proxy (UMMA operand reads)
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  const int Wo = 2 * g.Wi, rows_mid = 4 * g.Hi * g.Wi;

  if (warp < 4) {
    // ================= producer: NCHW bf16 -> K-major rows, PixelShuffle folded in =================
    const int nchunk = P.HW / 8;  // 16-byte chunks of 8 consecutive spatial positions per channel
    const int ntasks = 4 * 4 * nchunk;
    int it = 0;
    for (int b = blockIdx.x; b < P.B; b += gridDim.x) {
      const __nv_bfloat16* fb = P.feat + (size_t)b * P.C * P.HW;
      for (int st = 0; st < P.nstages; ++st, ++it) {
        const int s = it % HB_STAGES;
        const uint32_t ph = (it / HB_STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        unsigned char* As = stage_base + s * stage_bytes;
        if (tid == 0) {
          mbar_expect_tx(&full[s], HB_BSTAGE_BYTES);
          bulk_g2s(As + a_stage_bytes, reinterpret_cast<const unsigned char*>(P.wpk) + (size_t)st * HB_BSTAGE_BYTES,
                   HB_BSTAGE_BYTES, &full[s]);
        }
        for (int task = tid; task < ntasks; task += 128) {
          const int sc = task % nchunk;
          const int q = (task / nchunk) & 3;
          const int kc = task / (4 * nchunk);
          uint4 v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int ch = 4 * (st * HB_KSTAGE + kc * 8 + e) + q;
            v[e] = __ldg(reinterpret_cast<const uint4*>(fb + (size_t)ch * P.HW + sc * 8));
          }
          const int di = q >> 1, dj = q & 1;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t w0[8] = {v[0].x, v[1].x, v[2].x, v[3].x, v[4].x, v[5].x, v[6].x, v[7].x};
            const uint32_t w1[8] = {v[0].y, v[1].y, v[2].y, v[3].y, v[4].y, v[5].y, v[6].y, v[7].y};
            const uint32_t w2[8] = {v[0].z, v[1].z, v[2].z, v[3].z, v[4].z, v[5].z, v[6].z, v[7].z};
            const uint32_t w3[8] = {v[0].w, v[1].w, v[2].w, v[3].w, v[4].w, v[5].w, v[6].w, v[7].w};
            const uint32_t* wj = j == 0 ? w0 : (j == 1 ? w1 : (j == 2 ? w2 : w3));
#pragma unroll
            for (int hl = 0; hl < 2; ++hl) {
              const uint32_t sel = hl ? 0x7632u : 0x5410u;
              uint4 o;
              o.x = __byte_perm(wj[0], wj[1], sel);
              o.y = __byte_perm(wj[2], wj[3], sel);
              o.z = __byte_perm(wj[4], wj[5], sel);
              o.w = __byte_perm(wj[6], wj[7], sel);
              const int sp = sc * 8 + 2 * j + hl;  // spatial index i*W + jcol
              const int i = sp / P.W, jc = sp - i * P.W;
              const int row = (2 * i + di) * g.P + (2 * jc + dj);
              *reinterpret_cast<uint4*>(As + ((size_t)kc * g.rows_alloc + row) * 16) = o;
            }
          }
        }
        fence_proxy_async();
        tc::mbar_arrive(&full[s]);
      }
    }
  } else if (warp == 4) {
    // ================= MMA issuer =================
    const uint32_t idesc = tc::make_idesc_bf16_f32(128, HB_NCOLS);
    const uint32_t lbo_a = g.rows_alloc * 16, lbo_b = HB_NCOLS * 16;
    int it = 0;
    uint32_t fph = 0;
    for (int b = blockIdx.x; b < P.B; b += gridDim.x) {
      mbar_wait(tmem_empty, fph ^ 1);
      tc::fence_after_sync();
      for (int st = 0; st < P.nstages; ++st, ++it) {
        const int s = it % HB_STAGES;
        const uint32_t ph = (it / HB_STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc::fence_after_sync();
        if (lane == 0) {
          const uint32_t a0 = smem_u32(stage_base + s * stage_bytes);
          const uint32_t b0 = a0 + a_stage_bytes;
          for (int t = 0; t < g.tiles; ++t) {
#pragma unroll
            for (int sh = 0; sh < 4; ++sh) {
              const int shift_rows = (sh >> 1) * g.P + (sh & 1);
#pragma unroll
              for (int k16 = 0; k16 < 2; ++k16) {
                const uint32_t aa = a0 + (2 * k16) * lbo_a + (t * 128 + shift_rows) * 16;
                const uint32_t bb = b0 + (sh * 4 + 2 * k16) * lbo_b;
                const uint64_t ad = tc::make_smem_desc(aa, lbo_a, 128);
                const uint64_t bd = tc::make_smem_desc(bb, lbo_b, 128);
                tc::umma_bf16(tmem_base + t * HB_NCOLS, ad, bd, idesc, (st | sh | k16) != 0 ? 1u : 0u);
              }
            }
          }
          tc::umma_commit(&empty[s]);
        }
        __syncwarp();
      }
      if (lane == 0) tc::umma_commit(tmem_full);
      __syncwarp();
      fph ^= 1;
    }
  } else {
    // ================= epilogue: TMEM -> (+bias) -> bf16 -> mid activations (A layout) =================
    const int q = warp & 3;
    uint32_t fph = 0;
    for (int b = blockIdx.x; b < P.B; b += gridDim.x) {
      mbar_wait(tmem_full, fph);
      tc::fence_after_sync();
      for (int t = 0; t < g.tiles; ++t) {
        float d[HB_NCOLS];
#pragma unroll
        for (int cc = 0; cc < HB_NCOLS / 16; ++cc) {
          float v[16];
          tc::tmem_ld16(tmem_base + ((uint32_t)(32 * q) << 16) + t * HB_NCOLS + cc * 16, v);
#pragma unroll
          for (int i = 0; i < 16; ++i) d[cc * 16 + i] = v[i];
        }
        const int row = t * 128 + 32 * q + lane;
        const int m = row / g.P, n = row - m * g.P;
        if (m < g.Hi && n < g.Wi) {
#pragma unroll
          for (int cls = 0; cls < 4; ++cls) {
            const int y = 2 * m + (cls >> 1), x = 2 * n + (cls & 1);
            const size_t row2 = (size_t)y * Wo + x;
#pragma unroll
            for (int kc = 0; kc < 4; ++kc) {
              uint32_t pk[4];
#pragma unroll
              for (int e2 = 0; e2 < 4; ++e2) {
                const int c0 = kc * 8 + 2 * e2, c1i = c0 + 1;
                const float f0 = (c0 < P.c1 && c0 < HB_CLS) ? d[cls * HB_CLS + (c0 < HB_CLS ? c0 : 0)] + __ldg(P.bias + c0) : 0.f;
                const float f1 = (c1i < P.c1 && c1i < HB_CLS) ? d[cls * HB_CLS + (c1i < HB_CLS ? c1i : 0)] + __ldg(P.bias + c1i) : 0.f;
                __nv_bfloat162 h2 = __floats2bfloat162_rn(f0, f1);
                pk[e2] = *reinterpret_cast<uint32_t*>(&h2);
              }
              *reinterpret_cast<uint4*>(P.mid + ((((size_t)b * 4 + kc) * rows_mid + row2) * 8)) =
                  make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
          }
        }
      }
      tc::fence_before_sync();
      tc::mbar_arrive(tmem_empty);
      fph ^= 1;
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc(tmem_base, 512);
}

// =====================================================================================================
// k1b: second transposed convolution + plane softmax
// =====================================================================================================
struct K1bParams {
  const __nv_bfloat16* mid;  // [B][4][Hi*Wi][8]
  const __nv_bfloat16* wpk;  // packed weights, one stage (K = 32)
  const float* bias;         // [c2]
  float* out;                // [B][c2][2Hi][2Wi]
  int B, c2, final_softmax;
  HeadGeom g;
};

constexpr int K1B_TPB = 3;  // M-tiles per TMEM buffer (3 * 80 = 240 columns; two buffers at 0 and 256)

__global__ void __launch_bounds__(HB_THREADS, 1) k1b_convt_softmax_kernel(const __grid_constant__ K1bParams P) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const HeadGeom g = P.g;
  const int a_bytes = 4 * g.rows_alloc * 16;
  unsigned char* As = smem;
  unsigned char* Bs = smem + a_bytes;
  float* stat = reinterpret_cast<float*>(Bs + HB_BSTAGE_BYTES);  // [2][HB_CLS][128] then fin[2][HB_CLS]
  float* fin = stat + 2 * HB_CLS * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(fin + 2 * HB_CLS + 8);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + 1;
  uint64_t* b_full = bars + 2;
  uint64_t* t_full = bars + 3;   // [2]
  uint64_t* t_empty = bars + 5;  // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 7);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < a_bytes / 16; i += HB_THREADS) reinterpret_cast<uint4*>(As)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    mbar_init(b_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&t_full[i], 1);
      mbar_init(&t_empty[i], 128);
    }
    fence_mbar_init();
  }
  if (warp == 4) tc::tmem_alloc(tmem_ptr, 512);
  fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  const int Ho = 2 * g.Hi, Wo = 2 * g.Wi;
  const int npix = g.Hi * g.Wi;
  const int nbatch = (g.tiles + K1B_TPB - 1) / K1B_TPB;
  const int npass = P.final_softmax ? 2 : 1;

  if (warp == 0) {
    // ================= loader: weights once, then one frame of mid activations per iteration =======
    if (lane == 0) {
      mbar_expect_tx(b_full, HB_BSTAGE_BYTES);
      bulk_g2s(Bs, P.wpk, HB_BSTAGE_BYTES, b_full);
    }
    uint32_t ph = 0;
    for (int b = blockIdx.x; b < P.B; b += gridDim.x) {
      mbar_wait(a_empty, ph ^ 1);
      if (lane == 0) mbar_expect_tx(a_full, (uint32_t)(4 * npix * 16));
      __syncwarp();
      for (int i = lane; i < 4 * g.Hi; i += 32) {
        const int kc = i / g.Hi, y = i - kc * g.Hi;
        bulk_g2s(As + ((size_t)kc * g.rows_alloc + (size_t)y * g.P) * 16,
                 P.mid + (((size_t)b * 4 + kc) * npix + (size_t)y * g.Wi) * 8, (uint32_t)(g.Wi * 16), a_full);
      }
      ph ^= 1;
    }
  } else if (warp == 4) {
    // ================= MMA issuer =================
    const uint32_t idesc = tc::make_idesc_bf16_f32(128, HB_NCOLS);
    const uint32_t lbo_a = g.rows_alloc * 16, lbo_b = HB_NCOLS * 16;
    const uint32_t a0 = smem_u32(As), b0 = smem_u32(Bs);
    mbar_wait(b_full, 0);
    uint32_t aph = 0;
    int nb = 0;  // running batch counter -> buffer + phase
    for (int b = blockIdx.x; b < P.B; b += gridDim.x) {
      mbar_wait(a_full, aph);
      tc::fence_after_sync();
      for (int pass = 0; pass < npass; ++pass) {
        for (int bt = 0; bt < nbatch; ++bt, ++nb) {
          const int buf = nb & 1;
          mbar_wait(&t_empty[buf], ((nb >> 1) & 1) ^ 1);
          tc::fence_after_sync();
          if (lane == 0) {
            for (int tt = 0; tt < K1B_TPB; ++tt) {
              const int t = bt * K1B_TPB + tt;
              if (t >= g.tiles) break;
#pragma unroll
              for (int sh = 0; sh < 4; ++sh) {
                const int shift_rows = (sh >> 1) * g.P + (sh & 1);
#pragma unroll
                for (int k16 = 0; k16 < 2; ++k16) {
                  const uint32_t aa = a0 + (2 * k16) * lbo_a + (t * 128 + shift_rows) * 16;
                  const uint32_t bb = b0 + (sh * 4 + 2 * k16) * lbo_b;
                  const uint64_t ad = tc::make_smem_desc(aa, lbo_a, 128);
                  const uint64_t bd = tc::make_smem_desc(bb, lbo_b, 128);
                  tc::umma_bf16(tmem_base + buf * 256 + tt * HB_NCOLS, ad, bd, idesc, (sh | k16) != 0 ? 1u : 0u);
                }
              }
            }
            tc::umma_commit(&t_full[buf]);
          }
          __syncwarp();
        }
      }
      if (lane == 0) tc::umma_commit(a_empty);  // all reads of this frame's activations have completed
      __syncwarp();
      aph ^= 1;
    }
  } else if (warp >= 5) {
    // ================= epilogue =================
    const int q = warp & 3;
    const int et = (warp - 5) * 32 + lane;  // 0..127
    const float L2E = 1.4426950408889634f;
    int nb = 0;
    for (int b = blockIdx.x; b < P.B; b += gridDim.x) {
      float mx[HB_CLS], sm[HB_CLS];
#pragma unroll
      for (int o = 0; o < HB_CLS; ++o) {
        mx[o] = -3.0e38f;
        sm[o] = 0.f;
      }
      for (int pass = 0; pass < npass; ++pass) {
        const bool write = (pass == npass - 1);
        for (int bt = 0; bt < nbatch; ++bt, ++nb) {
          const int buf = nb & 1;
          mbar_wait(&t_full[buf], (nb >> 1) & 1);
          tc::fence_after_sync();
          for (int tt = 0; tt < K1B_TPB; ++tt) {
            const int t = bt * K1B_TPB + tt;
            if (t >= g.tiles) break;
            float d[HB_NCOLS];
#pragma unroll
            for (int cc = 0; cc < HB_NCOLS / 16; ++cc) {
              float v[16];
              tc::tmem_ld16(tmem_base + ((uint32_t)(32 * q) << 16) + buf * 256 + tt * HB_NCOLS + cc * 16, v);
#pragma unroll
              for (int i = 0; i < 16; ++i) d[cc * 16 + i] = v[i];
            }
            const int row = t * 128 + 32 * q + lane;
            const int m = row / g.P, n = row - m * g.P;
            if (m < g.Hi && n < g.Wi) {
#pragma unroll
              for (int o = 0; o < HB_CLS; ++o) {
                if (o >= P.c2) break;
                const float bo = __ldg(P.bias + o);
                const float l0 = d[o] + bo, l1 = d[HB_CLS + o] + bo, l2 = d[2 * HB_CLS + o] + bo, l3 = d[3 * HB_CLS + o] + bo;
                if (!write) {
                  const float mm = fmaxf(fmaxf(l0, l1), fmaxf(l2, l3));
                  if (mm > mx[o]) {
                    sm[o] *= fast_exp2((mx[o] - mm) * L2E);
                    mx[o] = mm;
                  }
                  sm[o] += fast_exp2((l0 - mx[o]) * L2E) + fast_exp2((l1 - mx[o]) * L2E) + fast_exp2((l2 - mx[o]) * L2E) +
                           fast_exp2((l3 - mx[o]) * L2E);
                } else {
                  float p0 = l0, p1 = l1, p2 = l2, p3 = l3;
                  if (P.final_softmax) {
                    const float M = fin[o], inv = fin[HB_CLS + o];
                    p0 = fast_exp2((l0 - M) * L2E) * inv;
                    p1 = fast_exp2((l1 - M) * L2E) * inv;
                    p2 = fast_exp2((l2 - M) * L2E) * inv;
                    p3 = fast_exp2((l3 - M) * L2E) * inv;
                  }
                  float* dst = P.out + (((size_t)b * P.c2 + o) * Ho + 2 * m) * Wo + 2 * n;
                  *reinterpret_cast<float2*>(dst) = make_float2(p0, p1);        // (even row: px = 0, 1)
                  *reinterpret_cast<float2*>(dst + Wo) = make_float2(p2, p3);   // (odd row)
                }
              }
            }
          }
          tc::fence_before_sync();
          tc::mbar_arrive(&t_empty[buf]);
        }
        if (!write) {
          // reduce the per-thread online-softmax states of the 128 epilogue threads, per plane
#pragma unroll
          for (int o = 0; o < HB_CLS; ++o) {
            stat[o * 128 + et] = mx[o];
            stat[(HB_CLS + o) * 128 + et] = sm[o];
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (et < P.c2) {
            float M = -3.0e38f;
            for (int i = 0; i < 128; ++i) M = fmaxf(M, stat[et * 128 + i]);
            float S = 0.f;
            for (int i = 0; i < 128; ++i) S += stat[(HB_CLS + et) * 128 + i] * fast_exp2((stat[et * 128 + i] - M) * L2E);
            fin[et] = M;
            fin[HB_CLS + et] = 1.0f / S;
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc(tmem_base, 512);
}

static size_t k1a_smem_bytes(const HeadGeom& g) { return (size_t)HB_STAGES * (4 * g.rows_alloc * 16 + HB_BSTAGE_BYTES) + 128; }
static size_t k1b_smem_bytes(const HeadGeom& g) {
  return (size_t)4 * g.rows_alloc * 16 + HB_BSTAGE_BYTES + (2 * HB_CLS * 128 + 2 * HB_CLS + 8) * sizeof(float) + 128;
}

}  // namespace lpb

// workspace layout: [packed w1][packed w2][mid activations]
extern "C" int lpb_head_bf16_workspace_bytes(int B, int C, int H, int W, int c1, int c2, size_t* bytes) {
  using namespace lpb;
  LPB_REQUIRE(bytes, "head_bf16_workspace_bytes: null pointer");
  LPB_REQUIRE(B >= 0 && C >= 128 && C % 128 == 0 && H >= 1 && W >= 1 && c1 >= 1 && c2 >= 1, "head_bf16_workspace_bytes: bad shape");
  const size_t w1 = (size_t)(C / 4 / HB_KSTAGE) * HB_BSTAGE_BYTES, w2 = HB_BSTAGE_BYTES;
  const size_t mid = (size_t)B * 4 * (16 * H * W) * 16;
  *bytes = w1 + w2 + mid;
  return LPB_OK;
}

extern "C" int lpb_head_fwd_bf16(const void* features, int B, int C, int H, int W, const float* w1, const float* b1, int c1,
                                 const float* w2, const float* b2, int c2, int final_softmax, float* out, void* workspace,
                                 void* stream) {
  using namespace lpb;
  LPB_REQUIRE(features && w1 && b1 && w2 && b2 && out && workspace, "head_fwd_bf16: null pointer");
  LPB_REQUIRE(B >= 0 && C >= 128 && C % 128 == 0 && H >= 1 && W >= 1, "head_fwd_bf16: bad feature shape C=%d H=%d W=%d", C, H, W);
  LPB_REQUIRE((H * W) % 8 == 0, "head_fwd_bf16: H*W must be a multiple of 8 (got %d)", H * W);
  LPB_REQUIRE(c1 >= 1 && c1 <= HB_CLS && c2 >= 1 && c2 <= HB_CLS, "head_fwd_bf16: channel counts %d/%d exceed %d", c1, c2, HB_CLS);
  if (B == 0) return LPB_OK;
  const HeadGeom g1 = make_geom(2 * H, 2 * W), g2 = make_geom(4 * H, 4 * W);
  if (g1.tiles * HB_NCOLS > 512) {
    set_error("head_fwd_bf16: %d M-tiles of layer 1 exceed TMEM (feature map %dx%d too large for this build)", g1.tiles, H, W);
    return LPB_ERR_UNSUPPORTED;
  }
  int dev = 0, max_smem = 0, sms = 0;
  LPB_CUDA(cudaGetDevice(&dev));
  LPB_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  LPB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t s1 = k1a_smem_bytes(g1), s2 = k1b_smem_bytes(g2);
  if ((int64_t)s1 > max_smem || (int64_t)s2 > max_smem) {
    set_error("head_fwd_bf16: needs %zu / %zu B shared memory (> %d)", s1, s2, max_smem);
    return LPB_ERR_UNSUPPORTED;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int nst = C / 4 / HB_KSTAGE;
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  __nv_bfloat16* wp1 = reinterpret_cast<__nv_bfloat16*>(ws);
  __nv_bfloat16* wp2 = reinterpret_cast<__nv_bfloat16*>(ws + (size_t)nst * HB_BSTAGE_BYTES);
  __nv_bfloat16* mid = reinterpret_cast<__nv_bfloat16*>(ws + (size_t)(nst + 1) * HB_BSTAGE_BYTES);
  pack_convt_weights_kernel<<<64, 256, 0, s>>>(w1, C / 4, c1, nst, wp1);
  pack_convt_weights_kernel<<<8, 256, 0, s>>>(w2, c1, c2, 1, wp2);
  K1aParams pa;
  pa.feat = static_cast<const __nv_bfloat16*>(features);
  pa.wpk = wp1;
  pa.bias = b1;
  pa.mid = mid;
  pa.B = B;
  pa.C = C;
  pa.HW = H * W;
  pa.W = W;
  pa.c1 = c1;
  pa.nstages = nst;
  pa.g = g1;
  LPB_CUDA(cudaFuncSetAttribute(k1a_shuffle_convt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s1));
  k1a_shuffle_convt_kernel<<<B < sms ? B : sms, HB_THREADS, s1, s>>>(pa);
  K1bParams pb;
  pb.mid = mid;
  pb.wpk = wp2;
  pb.bias = b2;
  pb.out = out;
  pb.B = B;
  pb.c2 = c2;
  pb.final_softmax = final_softmax;
  pb.g = g2;
  LPB_CUDA(cudaFuncSetAttribute(k1b_convt_softmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s2));
  k1b_convt_softmax_kernel<<<B < sms ? B : sms, HB_THREADS, s2, s>>>(pb);
  LPB_CUDA(cudaGetLastError());
  return LPB_OK;
}
