/*
 * builtin_targets.cu -- the synthetic likelihoods of BASELINE.json's configs as a plugin .so
 * (libslgpu_targets.so).  Each functor is the device form of a stateline worker's
 *     double f(uint jobType, const std::vector<double>& x)            (src/app/workerwrapper.hpp:21)
 * and returns a negative log-likelihood.  The first one is the reference's own demo worker
 * (src/bin/demo-worker.cpp:37-45); the others have no code in the reference and follow SURVEY.md 8d.
 * A user plugin looks exactly like one of these blocks.
 */
#include "slgpu_device.cuh"

namespace {

using slgpu::u32;

/* config #1 -- gaussianNLL, src/bin/demo-worker.cpp:37-45: 0.5*|x|^2 for every job type */
struct DemoGauss {
  __device__ explicit DemoGauss(const void*) {}
  __device__ double operator()(u32, const double* x, u32 d) const {
    double sq = 0.0;
#pragma unroll
    for (u32 i = 0; i < d; ++i) sq += x[i] * x[i];
    return 0.5 * sq;
  }
};

/* config #2 -- 2-D double well: f(0)=8(x0^2-1)^2, f(1)=0.5*x1^2 */
struct DoubleWell {
  __device__ explicit DoubleWell(const void*) {}
  __device__ double operator()(u32 job, const double* x, u32) const {
    if (job == 0) {
      const double t = x[0] * x[0] - 1.0;
      return 8.0 * (t * t);
    }
    return 0.5 * (x[1] * x[1]);
  }
};

/* config #3 -- factorised Gaussian: job j owns dims 2j, 2j+1; payload = mu[d], inv_s[d] */
struct GaussFact {
  const double* prm;
  __device__ explicit GaussFact(const void* payload) : prm((const double*)payload) {}
  template <int D>
  static constexpr int static_job_types() { return (D + 1) / 2; }
  __device__ double operator()(u32 job, const double* x, u32 d) const {
    const double* mu = prm;
    const double* is = prm + d;
    const u32 a = 2 * job, b = 2 * job + 1;
    const double ta = (x[a] - __ldg(mu + a)) * __ldg(is + a);
    double acc = ta * ta;
    if (b < d) {
      const double tb = (x[b] - __ldg(mu + b)) * __ldg(is + b);
      acc = fma(tb, tb, acc);
    }
    return 0.5 * acc;
  }
};

/* config #4 -- correlated Gaussian 0.5*|U x|^2 split in J row blocks; payload = {J as double, U[d*d] row-major upper, transpose(U)[d*d]} */
struct GaussCorr {
  const double* prm;
  __device__ explicit GaussCorr(const void* payload) : prm((const double*)payload) {}
  __device__ double operator()(u32 job, const double* x, u32 d) const {
    const u32 J = (u32)prm[0];
    const double* U = prm + 1;
    const u32 per = (d + J - 1) / J;
    const u32 r0 = job * per;
    const u32 r1 = (r0 + per < d) ? r0 + per : d;
    double acc = 0.0;
    for (u32 r = r0; r < r1; ++r) {
      double dot = 0.0;
      for (u32 c = r; c < d; ++c) dot = fma(__ldg(U + (size_t)r * d + c), x[c], dot);
      acc = fma(dot, dot, acc);
    }
    return 0.5 * acc;
  }
  /* columns c of row group G: only the row groups m <= G reach the diagonal, and only group G needs the row test.  The
   * columns go in batches whose loads are all issued before the first fma (at most eight per lane). */
  template <int G>
  __device__ __forceinline__ void dot_group(const double* x, u32 d, int lane, const double*& col, double* dot) const {
    constexpr int NG = G + 1;
    constexpr int B = NG >= 3 ? 2 : (NG == 2 ? 4 : 8);
    const u32 c1 = (32u * G + 32u < d) ? 32u * G + 32u : d;
    u32 c = 32u * G;
    for (; c + B <= c1; c += B) {
      double v[B][NG];
#pragma unroll
      for (int b = 0; b < B; ++b)
#pragma unroll
        for (int m = 0; m <= G; ++m) v[b][m] = (m < G || (u32)lane + 32u * (u32)m <= c + b) ? __ldg(col + (size_t)b * d + 32 * m) : 0.0;
#pragma unroll
      for (int b = 0; b < B; ++b) {
        const double xc = x[c + b];
#pragma unroll
        for (int m = 0; m <= G; ++m)
          if (m < G || (u32)lane + 32u * (u32)m <= c + b) dot[m] = fma(v[b][m], xc, dot[m]);
      }
      col += (size_t)B * d;
    }
    for (; c < c1; ++c) {
      const double xc = x[c];
#pragma unroll
      for (int m = 0; m <= G; ++m)
        if (m < G || (u32)lane + 32u * (u32)m <= c) dot[m] = fma(__ldg(col + 32 * m), xc, dot[m]);
      col += d;
    }
    if constexpr (G < 3) {
      if (c1 < d) dot_group<G + 1>(x, d, lane, col, dot);
    }
  }
  /* All job types with one warp (d <= 128).  Lane l owns rows l, l+32, ...: every row's dot product is still its own
   * fma chain over ascending c, the rows of a lane run interleaved, and U is read through its transpose (the second
   * matrix of the payload) so that the lanes of a row group read consecutive addresses.  Then lane j folds the rows
   * of job j in row order. */
  __device__ void warp_jobs(const double* x, u32 d, u32 n_jobs, int lane, double* values, double* dots) const {
    const double* col = prm + 1 + (size_t)d * d + lane; /* Ut[c * d + r] = U[r * d + c], this lane's rows */
    double dot[4] = {0.0, 0.0, 0.0, 0.0};
    dot_group<0>(x, d, lane, col, dot);
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const u32 r = (u32)lane + 32u * (u32)m;
      if (r < d) dots[r] = dot[m];
    }
    __syncwarp();
    const u32 per = (d + n_jobs - 1) / n_jobs;
    for (u32 job = (u32)lane; job < n_jobs; job += 32) {
      const u32 r0 = job * per;
      const u32 r1 = (r0 + per < d) ? r0 + per : d;
      double acc = 0.0;
      for (u32 r = r0; r < r1; ++r) acc = fma(dots[r], dots[r], acc);
      values[job] = 0.5 * acc;
    }
    __syncwarp();
  }
};

/* config #5 -- Rosenbrock/20: d-1 terms split evenly over J jobs; payload = {J as double} */
struct Rosenbrock {
  const double* prm;
  __device__ explicit Rosenbrock(const void* payload) : prm((const double*)payload) {}
  __device__ double operator()(u32 job, const double* x, u32 d) const {
    const u32 J = (u32)prm[0];
    const u32 nterm = d - 1;
    const u32 per = (nterm + J - 1) / J;
    const u32 i0 = job * per;
    const u32 i1 = (i0 + per < nterm) ? i0 + per : nterm;
    double acc = 0.0;
    for (u32 i = i0; i < i1; ++i) {
      const double a = x[i + 1] - x[i] * x[i];
      const double b = 1.0 - x[i];
      acc += (100.0 * (a * a) + b * b) / 20.0;
    }
    return acc;
  }
  /* all job types with one warp: the terms are dealt to the lanes, then lane j adds the terms of job j in order */
  __device__ void warp_jobs(const double* x, u32 d, u32 n_jobs, int lane, double* values, double* terms) const {
    const u32 nterm = d - 1;
    for (u32 i = (u32)lane; i < nterm; i += 32) {
      const double a = x[i + 1] - x[i] * x[i];
      const double b = 1.0 - x[i];
      terms[i] = (100.0 * (a * a) + b * b) / 20.0;
    }
    __syncwarp();
    const u32 per = (nterm + n_jobs - 1) / n_jobs;
    for (u32 job = (u32)lane; job < n_jobs; job += 32) {
      const u32 i0 = job * per;
      const u32 i1 = (i0 + per < nterm) ? i0 + per : nterm;
      double acc = 0.0;
      for (u32 i = i0; i < i1; ++i) acc += terms[i];
      values[job] = acc;
    }
    __syncwarp();
  }
};

}  // namespace

SLGPU_DEFINE_PLUGIN(slgpu_plugin_entry, DemoGauss, "demo_gauss", 0, 0, 0, 4)
SLGPU_DEFINE_PLUGIN(slgpu_plugin_demo_gauss, DemoGauss, "demo_gauss", 0, 0, 0, 4)
SLGPU_DEFINE_PLUGIN(slgpu_plugin_double_well, DoubleWell, "double_well", 2, 2, 0, 2)
SLGPU_DEFINE_PLUGIN(slgpu_plugin_gauss_fact, GaussFact, "gauss_fact", 0, 0, 0, 20)
SLGPU_DEFINE_PLUGIN(slgpu_plugin_gauss_corr, GaussCorr, "gauss_corr", 0, 0, 0)
SLGPU_DEFINE_PLUGIN(slgpu_plugin_rosenbrock, Rosenbrock, "rosenbrock", 0, 0, 0)
