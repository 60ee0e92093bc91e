#include "upsample_tables.cuh"

#include <cmath>
#include <cstdarg>
#include <map>
#include <mutex>
#include <string>
#include <tuple>

#include "lpb_common.cuh"

namespace lpb {

// ---- error plumbing shared by the whole library -----------------------------------------------
static thread_local std::string g_last_error;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}
int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %s (%s) at %s:%d in %s", cudaGetErrorName(e), cudaGetErrorString(e), file, line, what);
  return LPB_ERR_CUDA;
}
const char* last_error_cstr() { return g_last_error.c_str(); }

// ---- bicubic (A = -0.75) + binomial blur, composed in double ------------------------------------
static void cubic_coeffs(double t, double w[4]) {
  const double A = -0.75;
  auto c1 = [&](double x) { return ((A + 2.0) * x - (A + 3.0)) * x * x + 1.0; };          // |x| <= 1
  auto c2 = [&](double x) { return ((A * x - 5.0 * A) * x + 8.0 * A) * x - 4.0 * A; };    // 1 < |x| < 2
  w[0] = c2(t + 1.0);
  w[1] = c1(t);
  w[2] = c1(1.0 - t);
  w[3] = c2(2.0 - t);
}

// dense (n*2^ds) x n operator, row-major
static std::vector<double> dense_operator(int n, int ds) {
  std::vector<double> total((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) total[(size_t)i * n + i] = 1.0;
  int cur = n;
  for (int s = 0; s < ds; ++s) {
    const int nf = 2 * cur;
    // stage = blur(nf x nf) * bicubic(nf x cur); apply to `total` (cur x n)
    std::vector<double> bic((size_t)nf * n, 0.0);
    for (int i = 0; i < nf; ++i) {
      const double src = (i + 0.5) / 2.0 - 0.5;
      const int i0 = (int)std::floor(src);
      double w[4];
      cubic_coeffs(src - i0, w);
      for (int k = 0; k < 4; ++k) {
        int c = i0 - 1 + k;
        c = c < 0 ? 0 : (c > cur - 1 ? cur - 1 : c);
        for (int j = 0; j < n; ++j) bic[(size_t)i * n + j] += w[k] * total[(size_t)c * n + j];
      }
    }
    std::vector<double> next((size_t)nf * n, 0.0);
    const double bw[5] = {1.0 / 16, 4.0 / 16, 6.0 / 16, 4.0 / 16, 1.0 / 16};
    for (int i = 0; i < nf; ++i)
      for (int k = 0; k < 5; ++k) {
        const int r = i - 2 + k;
        if (r < 0 || r >= nf) continue;  // zero padding ("constant" border)
        for (int j = 0; j < n; ++j) next[(size_t)i * n + j] += bw[k] * bic[(size_t)r * n + j];
      }
    total.swap(next);
    cur = nf;
  }
  return total;
}

bool build_host_table(int n, int ds, HostTable* out) {
  if (n < 1 || ds < 1 || ds > 3) {
    set_error("upsample table: unsupported n=%d ds=%d", n, ds);
    return false;
  }
  const int F = 1 << ds, R = ds + 2, W = 2 * R + 1, N = n * F;
  std::vector<double> U = dense_operator(n, ds);
  out->n = n;
  out->ds = ds;
  out->win.assign((size_t)N * W, 0.f);
  double lip = 0.0;
  for (int i = 0; i < N; ++i) {
    const int a = i / F;
    double sabs = 0.0;
    for (int c = 0; c < n; ++c) {
      const double u = U[(size_t)i * n + c];
      sabs += std::fabs(u);
      if (u != 0.0) {
        const int t = c - (a - R);
        if (t < 0 || t >= W) {
          set_error("upsample table: tap outside band (n=%d ds=%d row=%d col=%d)", n, ds, i, c);
          return false;
        }
        out->win[(size_t)i * W + t] = (float)u;
      }
    }
    lip = sabs > lip ? sabs : lip;
  }
  out->lip = (float)(lip * (1.0 + 1e-6));
  // interior phase weights from a large-enough virtual axis (independent of n)
  {
    const int nv = 4 * R + 4;
    std::vector<double> V = dense_operator(nv, ds);
    out->phase.assign((size_t)F * W, 0.f);
    const int amid = nv / 2;
    for (int p = 0; p < F; ++p)
      for (int t = 0; t < W; ++t) out->phase[(size_t)p * W + t] = (float)V[(size_t)(amid * F + p) * nv + (amid - R + t)];
  }
  // verify: rows with R <= a <= n-1-R are phase-periodic
  for (int i = 0; i < N; ++i) {
    const int a = i / F;
    if (a < R || a > n - 1 - R) continue;
    for (int t = 0; t < W; ++t)
      if (std::fabs(out->win[(size_t)i * W + t] - out->phase[(size_t)(i % F) * W + t]) > 1e-7f) {
        set_error("upsample table: interior row %d not phase-periodic (n=%d ds=%d)", i, n, ds);
        return false;
      }
  }
  return true;
}

const DeviceTable* get_device_table(int n, int ds) {
  static std::mutex mu;
  static std::map<std::tuple<int, int, int>, DeviceTable*> cache;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    set_error("cudaGetDevice failed");
    return nullptr;
  }
  std::lock_guard<std::mutex> lock(mu);
  auto key = std::make_tuple(dev, n, ds);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  DeviceTable* t = new DeviceTable();
  if (!build_host_table(n, ds, &t->host)) {
    delete t;
    return nullptr;
  }
  float* d = nullptr;
  const size_t bytes = t->host.win.size() * sizeof(float);
  cudaError_t e = cudaMalloc(&d, bytes);
  if (e == cudaSuccess) e = cudaMemcpy(d, t->host.win.data(), bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cuda_fail(e, "upload upsample table", __FILE__, __LINE__);
    delete t;
    return nullptr;
  }
  t->win = d;
  cache[key] = t;
  return t;
}

}  // namespace lpb
