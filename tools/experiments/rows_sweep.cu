// compile-only experiment: register / code footprint of the rank-1 update swept by rows vs by columns (d = 20)
#include <cuda_runtime.h>
constexpr int kLs = 128;
__device__ __forceinline__ double smax(double a, double b) { return (a < b) ? b : a; }
__device__ __forceinline__ double div_by(double n, double c, double rc) {
  const double q0 = n * rc;
  const double r = fma(-q0, c, n);
  return fma(r, rc, q0);
}
// rows j = J0..J0+1 per trip (rolled over row pairs), k unrolled with static guards
template <int D>
__device__ __forceinline__ double shaper_rows(double* Lc, const double* jump, int jst, double count, double prop_norm) {
  const double eps = 1e-18;
  const double sq = sqrt(count), rsq = 1.0 / sq;
  const double scale = sqrt((count - 1.0) / count);
  double c[D], s[D], rc[D], f[D];
#pragma unroll
  for (int j = 0; j < D; ++j) {
    double xj = div_by(jump[j * jst], sq, rsq);
    double row[D];
#pragma unroll
    for (int k = 0; k <= j; ++k) row[k] = Lc[(j * (j + 1) / 2 + k) * kLs];
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < j; ++k) {
      double Ljk = div_by(row[k] * scale + s[k] * xj, c[k], rc[k]);
      Lc[(j * (j + 1) / 2 + k) * kLs] = Ljk;
      xj = c[k] * xj - s[k] * Ljk;
      acc = fma(Ljk, Ljk, acc);
    }
    const double V = smax(eps, row[j] * scale);
    const double rV = 1.0 / V;
    const double r = sqrt(V * V + xj * xj);
    c[j] = div_by(r, V, rV);
    s[j] = div_by(xj, V, rV);
    rc[j] = 1.0 / c[j];
    Lc[(j * (j + 1) / 2 + j) * kLs] = r;
    f[j] = fma(r, r, acc);
  }
  int nlive = D;
#pragma unroll
  for (int mm = 16; mm > 0; mm >>= 1) {
#pragma unroll
    for (int l = 0; l < mm; ++l)
      if (l + mm < D && l + mm < nlive) f[l] += f[l + mm];
    if (nlive > mm) nlive = mm;
  }
  return prop_norm / smax(sqrt(f[0]), eps);
}
__global__ void k_rows20(double* L, const double* jump, double* out, double count) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  out[t] = shaper_rows<20>(L + (size_t)(t / 128) * 210 * 128 + t % 128, jump + t, 4096, count, 2.5);
}
