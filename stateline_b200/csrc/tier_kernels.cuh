// tier_kernels.cuh -- the per-temperature-tier RegressionAdapter models on the device.
//
// Reference: RegressionAdapter (src/infer/adaptive.cpp:27-139).  One 3-feature linear model per tier,
// shared by every stack, for sigma (caps [-8,4]) and for beta (caps [0,4]) (src/app/serverwrapper.cpp:48-52).
// The step kernel gathers the sufficient statistics of adaptive.cpp:71-75 per chain; at an adapt point
// k_tier_adapt sums them over the stacks in a fixed-shape tree (same shape as the oracle's
// reduce_tree256), folds them into the running means, re-solves the 3x3 system with a column-pivoted
// Householder QR (adaptive.cpp:78) and rebuilds the beta ladder (computeBetaStack, adaptive.cpp:118-132).
// Compile with -fmad=false (operation order mirrors oracle/sl_oracle.cpp).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "slgpu_math.h"

namespace slgpu_tier {

constexpr int kModelStride = 16;  // mu_xx[9], mu_xy[3], w[3], count
constexpr double kInitialCount = 10.0;  // adaptive.cpp:27
constexpr double kTempVariance = 10.0;  // adaptive.cpp:28
constexpr double kSigMinCap = -8.0, kSigMaxCap = 4.0, kBetaMinCap = 0.0, kBetaMaxCap = 4.0;  // serverwrapper.cpp:48-49

__device__ inline double smax(double a, double b) { return (a < b) ? b : a; }
__device__ inline double smin(double a, double b) { return (b < a) ? b : a; }

// colPivHouseholderQr().solve() for a 3x3 system, A row-major (restated from the oracle's qr_solve3)
__device__ inline void qr_solve3(const double* A, const double* b, double* w) {
  double M[3][3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) M[r][c] = A[3 * r + c];
  int perm[3] = {0, 1, 2};
  double cn[3];
  for (int c = 0; c < 3; ++c) cn[c] = M[0][c] * M[0][c] + M[1][c] * M[1][c] + M[2][c] * M[2][c];
  double tau[3] = {0.0, 0.0, 0.0};
  double cvec[3] = {b[0], b[1], b[2]};
  double maxpivot = 0.0;
  for (int k = 0; k < 3; ++k) {
    int p = k;
    for (int c = k + 1; c < 3; ++c)
      if (cn[c] > cn[p]) p = c;
    if (p != k) {
      for (int r = 0; r < 3; ++r) { const double tmp = M[r][k]; M[r][k] = M[r][p]; M[r][p] = tmp; }
      { const double tmp = cn[k]; cn[k] = cn[p]; cn[p] = tmp; }
      { const int tmp = perm[k]; perm[k] = perm[p]; perm[p] = tmp; }
    }
    const double c0 = M[k][k];
    double tail2 = 0.0;
    for (int r = k + 1; r < 3; ++r) tail2 += M[r][k] * M[r][k];
    double beta;
    if (tail2 == 0.0) {
      tau[k] = 0.0;
      beta = c0;
      for (int r = k + 1; r < 3; ++r) M[r][k] = 0.0;
    } else {
      beta = sqrt(c0 * c0 + tail2);
      if (c0 >= 0.0) beta = -beta;
      const double den = c0 - beta;
      for (int r = k + 1; r < 3; ++r) M[r][k] = M[r][k] / den;
      tau[k] = (beta - c0) / beta;
    }
    M[k][k] = beta;
    if (fabs(beta) > maxpivot) maxpivot = fabs(beta);
    for (int c = k + 1; c < 3; ++c) {
      double s = M[k][c];
      for (int r = k + 1; r < 3; ++r) s += M[r][k] * M[r][c];
      s *= tau[k];
      M[k][c] -= s;
      for (int r = k + 1; r < 3; ++r) M[r][c] -= s * M[r][k];
      cn[c] -= M[k][c] * M[k][c];
    }
    {
      double s = cvec[k];
      for (int r = k + 1; r < 3; ++r) s += M[r][k] * cvec[r];
      s *= tau[k];
      cvec[k] -= s;
      for (int r = k + 1; r < 3; ++r) cvec[r] -= s * M[r][k];
    }
  }
  const double thresh = 2.220446049250313080847263336181640625e-16 * 3.0 * maxpivot;
  int rank = 0;
  for (int k = 0; k < 3; ++k) {
    if (fabs(M[k][k]) > thresh) ++rank;
    else break;
  }
  double y[3] = {0.0, 0.0, 0.0};
  for (int i = rank - 1; i >= 0; --i) {
    double s = cvec[i];
    for (int c = i + 1; c < rank; ++c) s -= M[i][c] * y[c];
    y[i] = s / M[i][i];
  }
  for (int i = 0; i < 3; ++i) w[perm[i]] = y[i];
}

// RegressionAdapter ctor, adaptive.cpp:31-61
__device__ inline void model_init(double* m, double min_cap, double max_cap) {
  const double rej[3] = {min_cap, 0.0, 1.0};
  const double acc[3] = {max_cap, 0.0, 1.0};
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) m[3 * r + c] = 0.5 * rej[r] * rej[c] + 0.5 * acc[r] * acc[c];
  m[4] = kTempVariance;
  for (int r = 0; r < 3; ++r) {
    m[9 + r] = 0.5 * acc[r];
    m[12 + r] = m[9 + r];  // weight_ = mu_xy, not solved (adaptive.cpp:57)
  }
  m[15] = kInitialCount;
}

// adaptive.cpp:94-106
__device__ inline double model_predict(const double* w, double opt, double min_cap, double max_cap, double side) {
  const double eps = 1e-3;
  const double denom = smax(eps, w[0]);
  double numer = -(opt - w[1] * side - w[2]);
  numer = smax(smin(numer, denom * max_cap), denom * min_cap);
  return numer / denom;
}

// running mean of adaptive.cpp:71-75 with alpha = 1/count, applied to N summed samples at once
__device__ inline void model_update_batched(double* m, const double* P) {
  const double n = P[5];
  if (n == 0.0) return;
  const double c0 = m[15];
  const double c1 = c0 + n;
  const double sxx[9] = {P[0], P[1], P[2], P[1], P[3], P[4], P[2], P[4], P[5]};
  const double sxy[3] = {P[6], P[7], P[8]};
  for (int i = 0; i < 9; ++i) m[i] = (m[i] * c0 + sxx[i]) / c1;
  for (int i = 0; i < 3; ++i) m[9 + i] = (m[9 + i] * c0 + sxy[i]) / c1;
  m[15] = c1;
  qr_solve3(m, m + 9, m + 12);
}

// computeBetaStack with the shared models, adaptive.cpp:118-132
__device__ inline void compute_ladder(const double* beta_models, uint32_t T, double opt_swap, double* ladder) {
  double logbeta = 0.0;
  ladder[0] = 1.0;
  for (uint32_t i = 1; i < T; ++i) {
    logbeta -= model_predict(beta_models + (size_t)(i - 1) * kModelStride + 12, opt_swap, kBetaMinCap, kBetaMaxCap, logbeta);
    ladder[i] = slm_exp(logbeta);
  }
}

// models: [2][T][kModelStride] (0 sigma, 1 beta); sig_w: [T][3]; ladder: [T]
__global__ void k_tier_init(double* models, double* sig_w, double* ladder, uint32_t T, double opt_swap) {
  const uint32_t t = threadIdx.x;
  if (t < T) {
    model_init(models + (size_t)t * kModelStride, kSigMinCap, kSigMaxCap);
    model_init(models + (size_t)(T + t) * kModelStride, kBetaMinCap, kBetaMaxCap);
    for (int i = 0; i < 3; ++i) sig_w[3 * t + i] = models[(size_t)t * kModelStride + 12 + i];
  }
  __syncthreads();
  if (t == 0) compute_ladder(models + (size_t)T * kModelStride, T, opt_swap, ladder);
}

// Statistics layout: P[(k * T + t) * S + s], k < 9 (coalesced over the stacks of a tier).
// grid (T, 2, 9), 256 threads: block (t, which, k) sums statistic k of tier t over the stacks in the
// fixed shape of the oracle's reduce_tree256 (slot i adds stacks i, i+256, ... in order, then a halving
// tree), clears it, and writes sums[(which * T + t) * 9 + k].
__global__ void __launch_bounds__(256) k_tier_adapt(double* p_sig, double* p_beta, uint32_t S, uint32_t T, double* sums,
                                                    unsigned int* ticket, double* models, double* sig_w, double opt_swap,
                                                    double* ladder) {
  const uint32_t t = blockIdx.x, which = blockIdx.y, k = blockIdx.z, tid = threadIdx.x;
  // the step kernel behind this launch may start now (programmatic dependent launch): it runs its first round up to the
  // MH test on the chain state alone and waits for this grid before it reads the weights, the ladder or the statistics
  cudaTriggerProgrammaticLaunchCompletion();
  double* P = (which ? p_beta : p_sig) + ((size_t)k * T + t) * S;
  __shared__ double a[256];
  __shared__ bool last;
  double acc = 0.0;
  for (uint32_t m = tid; m < S; m += 256) {
    acc += P[m];
    P[m] = 0.0;
  }
  a[tid] = acc;
  __syncthreads();
  for (uint32_t s = 128; s > 0; s >>= 1) {
    if (tid < s) a[tid] += a[tid + s];
    __syncthreads();
  }
  // The block that finishes last folds the 2T x 9 sums into the models (one thread per model) and rebuilds the
  // ladder: one launch per adapt point instead of a reduce and a one-block solve behind it.
  if (tid == 0) {
    sums[(size_t)(which * T + t) * 9 + k] = a[0];
    __threadfence();
    const unsigned int n = gridDim.x * gridDim.y * gridDim.z;
    last = atomicInc(ticket, n - 1) == n - 1;  // wraps to 0 for the next adapt point
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (tid < 2 * T) {
    double mine[9];
    for (int j = 0; j < 9; ++j) mine[j] = __ldcg(sums + (size_t)tid * 9 + j);
    double* m = models + (size_t)tid * kModelStride;
    model_update_batched(m, mine);
    if (tid < T)
      for (int j = 0; j < 3; ++j) sig_w[3 * tid + j] = m[12 + j];
  }
  __syncthreads();
  if (tid == 0) compute_ladder(models + (size_t)T * kModelStride, T, opt_swap, ladder);
}

// ---- the reference's own schedule ("gpu": {"schedule": "sequential"}): a validation mode ----
// RegressionAdapter::update for ONE sample, adaptive.cpp:63-78: x = (x0, x1, 1) with x0 = -clamp(logval), x1 = side.
__device__ inline void model_update_one(double* m, double x0, double x1, double y) {
  const double x[3] = {x0, x1, 1.0};
  m[15] += 1.0;
  const double alpha = 1.0 / m[15];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) m[3 * r + c] = m[3 * r + c] * (1.0 - alpha) + x[r] * x[c] * alpha;
  for (int r = 0; r < 3; ++r) m[9 + r] = m[9 + r] * (1.0 - alpha) + x[r] * y * alpha;
  qr_solve3(m, m + 9, m + 12);
}

// One thread replays what Sampler::step() does to the shared tier models after a round, chain by chain in the order
// the reference's FIFO delivers results.  That order is data-independent: the constructor proposes chains C-1..0
// (sampler.cpp:128-132), a chain re-enters the queue when it is processed (or, in a swap round, when its colder
// neighbour unlocks it, sampler.cpp:237-257), so every round is processed in descending chain id.  The launch that
// precedes this kernel advanced every chain by exactly ONE round, so p_sig / p_beta hold each chain's single
// regression sample (x0 = P[2], x1 = P[4], y = P[8], P[5] = 1 if there was one).  Per chain, as sampler.cpp:158-186:
//   sigmaAdapter.update, sigma[id] = computeSigma          (adaptive.cpp:63-78,134-139)
//   if the chain was locked (swap round, t < T-1): betaAdapter.betaUpdate; at t == 0 computeBetaStack rewrites
//   beta_values[stack][1..T-1]; then setBeta(id+1, beta_values[id+1]) -- so after a cascade beta[1] is new and
//   beta[2..T-1] are the PREVIOUS ladder (values_ and ChainArray's beta differ; the reference's table prints values_).
// Serial over S*T chains: for validation against the reference's schedule on small configurations only.
__global__ void k_tier_scan_sequential(double* p_sig, double* p_beta, uint32_t S, uint32_t T, double* models, double* sig_w,
                                       double* ladder, double* sigma, double* beta, double* beta_values, double opt_accept,
                                       double opt_swap) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  const uint32_t C = S * T;
  for (uint32_t i = 0; i < C; ++i) {
    const uint32_t id = C - 1 - i, s = id / T, t = id % T;
    double* Ps = p_sig + (size_t)t * S + s;   // statistic k at Ps[k * T * S]
    double* Pb = p_beta + (size_t)t * S + s;
    const size_t ks = (size_t)T * S;
    if (Ps[5 * ks] != 0.0) {
      double* m = models + (size_t)t * kModelStride;
      const double x1 = Ps[4 * ks];
      model_update_one(m, Ps[2 * ks], x1, Ps[8 * ks]);
      sigma[id] = slm_exp(model_predict(m + 12, opt_accept, kSigMinCap, kSigMaxCap, x1));
      for (int k = 0; k < 9; ++k) Ps[k * ks] = 0.0;
    }
    if (Pb[5 * ks] != 0.0) {
      double* m = models + (size_t)(T + t) * kModelStride;
      model_update_one(m, Pb[2 * ks], Pb[4 * ks], Pb[8 * ks]);
      if (t == 0) {
        double logbeta = 0.0;
        for (uint32_t j = 1; j < T; ++j) {
          logbeta -= model_predict(models + (size_t)(T + j - 1) * kModelStride + 12, opt_swap, kBetaMinCap, kBetaMaxCap, logbeta);
          beta_values[id + j] = slm_exp(logbeta);
        }
      }
      beta[id + 1] = beta_values[id + 1];
      for (int k = 0; k < 9; ++k) Pb[k * ks] = 0.0;
    }
  }
  // what the batched kernels read as frozen weights: kept coherent with the models (their sigma / ladder results are
  // overwritten by this kernel after every round)
  for (uint32_t t = 0; t < T; ++t)
    for (int j = 0; j < 3; ++j) sig_w[3 * t + j] = models[(size_t)t * kModelStride + 12 + j];
  compute_ladder(models + (size_t)T * kModelStride, T, opt_swap, ladder);
}

// The acceptance windows of RegressionAdapter::update (adaptive.cpp:80-90), logging only: every chain keeps the last
// accept flags of its sigma adapter (one per round) and of its beta adapter (one per swap attempt of the pair
// (chain, chain+1)).  The reference's deque holds 999 flags once it has seen 1000 (it pushes, then pops when the size
// reaches n_window_ = 1000) and divides by the size before the pop: rate = sum(last 999) / 1000, and sum(all) / k
// while k < 1000.  Here: a 1024-bit ring per chain and adapter, ring[word][chain], updated after every launch from
// the per-round flags the step kernel logged.  win[0..1] = rings (sigma, beta), cnt / sum = [2][C].
constexpr uint32_t kWindow = 1000;  // adaptive.cpp:29
constexpr uint32_t kRingWords = 32;

__device__ inline void window_push(uint32_t* ring, size_t C, size_t c, uint32_t& k, uint32_t& sum, uint32_t bit) {
  k += 1;
  const uint32_t pos = k & 1023u;
  uint32_t* w = ring + (size_t)(pos >> 5) * C + c;
  *w = (*w & ~(1u << (pos & 31u))) | (bit << (pos & 31u));
  sum += bit;
  if (k >= kWindow) {
    const uint32_t old = (k - (kWindow - 1)) & 1023u;
    sum -= (ring[(size_t)(old >> 5) * C + c] >> (old & 31u)) & 1u;
  }
}

__global__ void k_window_push(const unsigned char* flags, uint32_t n_rounds, uint32_t C, uint32_t T, uint32_t* ring_sig,
                              uint32_t* ring_beta, uint32_t* cnt, uint32_t* sum) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  uint32_t ks = cnt[c], kb = cnt[C + c], ss = sum[c], sb = sum[C + c];
  for (uint32_t r = 0; r < n_rounds; ++r) {
    const uint32_t f = flags[(size_t)r * C + c];
    window_push(ring_sig, C, c, ks, ss, f & 1u);
    const uint32_t st = f >> 1;  // swapType: 0 none, 1 accepted, 2 rejected
    if (st != 0 && c % T != T - 1) window_push(ring_beta, C, c, kb, sb, st == 1u ? 1u : 0u);
  }
  cnt[c] = ks; cnt[C + c] = kb; sum[c] = ss; sum[C + c] = sb;
}

}  // namespace slgpu_tier
