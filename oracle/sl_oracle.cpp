/*
 * sl_oracle.cpp -- CPU oracle for the parallel-tempering hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see sl_oracle.h).  A from-scratch restatement, in plain C++17 with no
 * dependencies, of what NICTA/stateline's src/infer does.  Every function cites the reference
 * file:line it follows (paths relative to the reference root).  Nothing here is shipped or called by
 * the product path.
 *
 * Build: g++ -O2 -DNDEBUG -ffp-contract=off -fPIC -shared (oracle/Makefile).  -ffp-contract=off makes
 * every a*b+c two roundings unless std::fma is written out; the CUDA kernels are compiled with
 * -fmad=false and mirror the same explicit fma() calls, so +,-,*,/,sqrt,fma agree bit for bit; exp, log
 * and sin/cos come from include/slgpu_math.h on both sides, so whole runs agree bit for bit.
 */
#include "sl_oracle.h"

/* exp / log / sincos: the deterministic fp64 implementations shared with the device code (include/slgpu_math.h,
 * within 1 ulp of glibc, checked by tests/test_math.py) stand in for the reference's std::exp / std::log
 * (chainarray.cpp:41,64; sampler.cpp:160; adaptive.cpp:100-139) so that CPU and GPU results are bit-identical. */
#include "slgpu_math.h"

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <limits>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#ifdef SLO_LIBM
/* libsl_oracle_libm.so (oracle/Makefile): exp and log from glibc, exactly as the reference calls them
 * (chainarray.cpp:41,64; sampler.cpp:160; adaptive.cpp:116,130,136), instead of the slgpu_math.h kernels the
 * device shares.  Only tests/test_ref_pin.py uses this build: it must agree BIT FOR BIT with the reference's own
 * sources (oracle/_ref); the default build is then tied to it within 1 ulp per call (tests/test_math.py) and to the
 * GPU bit for bit. */
#define slm_exp(x) std::exp(x)
#define slm_log(x) std::log(x)
#endif

namespace {

using u32 = uint32_t;
using u64 = uint64_t;

/* ------------------------------------------------------------------------------------------------
 * Native RNG scheme (amended in DESIGN.md): Philox4x32-10,
 *   key = (seed_lo, seed_hi), ctr = (round_lo, global chain id, purpose | round_hi24<<8, block)
 *   purpose 0: proposal normals, block k -> Box-Muller pair for dims 2k, 2k+1
 *   purpose 1: MH uniform, purpose 2: swap uniform (chain id of the lower chain), purpose 3: initial point
 *   u = (k52 + 0.5) * 2^-52 with k52 the top 52 bits of a 64-bit word: strictly inside (0,1), exact in fp64.
 * The reference draws from mt19937 seeded by random_device (sampler.cpp:60, chainarray.cpp:32-34,57-59)
 * and is irreproducible by construction, so the scheme is ours; parity under it is statistical.
 * ---------------------------------------------------------------------------------------------- */
inline void philox4x32_10(const u32 ctr[4], const u32 key[2], u32 out[4]) {
  u32 c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  u32 k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    const u64 p0 = (u64)0xD2511F53u * c0;
    const u64 p1 = (u64)0xCD9E8D57u * c2;
    const u32 n0 = (u32)(p1 >> 32) ^ c1 ^ k0;
    const u32 n1 = (u32)p1;
    const u32 n2 = (u32)(p0 >> 32) ^ c3 ^ k1;
    const u32 n3 = (u32)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

inline void philox_draw(u64 seed, u32 chain, u64 round, u32 purpose, u32 block, double& ua, double& ub) {
  const u32 ctr[4] = {(u32)round, chain, (purpose & 0xffu) | ((u32)(round >> 32) << 8), block};
  const u32 key[2] = {(u32)seed, (u32)(seed >> 32)};
  u32 o[4];
  philox4x32_10(ctr, key, o);
  const u64 a = (((u64)o[1] << 32) | o[0]) >> 12;
  const u64 b = (((u64)o[3] << 32) | o[2]) >> 12;
  ua = ((double)a + 0.5) * 0x1p-52;
  ub = ((double)b + 0.5) * 0x1p-52;
}

void draw_normals(u64 seed, u32 chain, u64 round, u32 d, double* z) {
  for (u32 k = 0; 2 * k < d; ++k) {
    double ua, ub;
    philox_draw(seed, chain, round, 0, k, ua, ub);
    const double R = std::sqrt(-2.0 * slm_log(ua));
    double sn, cs;
    slm_sincos2pi(ub, &sn, &cs);
    z[2 * k] = R * cs;
    if (2 * k + 1 < d) z[2 * k + 1] = R * sn;
  }
}

double draw_uniform(u64 seed, u32 chain, u64 round, u32 purpose) {
  double ua, ub;
  philox_draw(seed, chain, round, purpose, 0, ua, ub);
  return ua;
}

/* VectorXd::Random(d) is uniform in [-1,1] (serverwrapper.cpp:34) */
void draw_init(u64 seed, u32 chain, u32 d, double* x) {
  for (u32 k = 0; 2 * k < d; ++k) {
    double ua, ub;
    philox_draw(seed, chain, 0, 3, k, ua, ub);
    x[2 * k] = 2.0 * ua - 1.0;
    if (2 * k + 1 < d) x[2 * k + 1] = 2.0 * ub - 1.0;
  }
}

/* ------------------------------------------------------------------------------------------------
 * bouncyBounds -- src/infer/sampler.cpp:23-56
 * ---------------------------------------------------------------------------------------------- */
void bouncy_bounds(const double* val, const double* mn, const double* mx, u32 d, double* out) {
  for (u32 i = 0; i < d; ++i) {
    const double delta = mx[i] - mn[i];
    double result = val[i];
    if (val[i] > mx[i]) {
      const double overstep = val[i] - mx[i];
      const int nSteps = (int)(overstep / delta);
      const double stillToGo = overstep - nSteps * delta;
      result = (nSteps % 2 == 0) ? mx[i] - stillToGo : mn[i] + stillToGo;
    }
    if (val[i] < mn[i]) {
      const double understep = mn[i] - val[i];
      const int nSteps = (int)(understep / delta);
      const double stillToGo = understep - nSteps * delta;
      result = (nSteps % 2 == 0) ? mn[i] + stillToGo : mx[i] - stillToGo;
    }
    out[i] = result;
  }
}

/* ------------------------------------------------------------------------------------------------
 * 3x3 column-pivoted Householder QR solve, standing in for Eigen's
 * mu_xx.colPivHouseholderQr().solve(mu_xy) (src/infer/adaptive.cpp:78).  Eigen (pinned 3.2.0 by
 * tools/fetch-deps:42-48) is not in the image; this restates the published algorithm: at step k the
 * remaining column of largest squared norm is swapped in, a Householder reflector zeroes the column
 * below the diagonal, norms are down-dated, pivots below eps*3*|max pivot| are treated as zero, then
 * Q^T b, back-substitution on the non-zero block and un-permutation.  Unpinned by Eigen itself (no reference
 * test covers it; the Eigen stand-in of oracle/_ref uses this same restatement); any backward-stable solver agrees to cond*eps.
 * A is row-major.
 * ---------------------------------------------------------------------------------------------- */
void qr_solve3(const double A[9], const double b[3], double w[3]) {
  double M[3][3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) M[r][c] = A[3 * r + c];
  int perm[3] = {0, 1, 2};
  double cn[3];
  for (int c = 0; c < 3; ++c) cn[c] = M[0][c] * M[0][c] + M[1][c] * M[1][c] + M[2][c] * M[2][c];
  double tau[3] = {0, 0, 0};
  double cvec[3] = {b[0], b[1], b[2]};
  double maxpivot = 0.0;
  for (int k = 0; k < 3; ++k) {
    int p = k;
    for (int c = k + 1; c < 3; ++c)
      if (cn[c] > cn[p]) p = c;
    if (p != k) {
      for (int r = 0; r < 3; ++r) std::swap(M[r][k], M[r][p]);
      std::swap(cn[k], cn[p]);
      std::swap(perm[k], perm[p]);
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
      beta = std::sqrt(c0 * c0 + tail2);
      if (c0 >= 0.0) beta = -beta;
      const double den = c0 - beta;
      for (int r = k + 1; r < 3; ++r) M[r][k] = M[r][k] / den;
      tau[k] = (beta - c0) / beta;
    }
    M[k][k] = beta;
    if (std::fabs(beta) > maxpivot) maxpivot = std::fabs(beta);
    /* apply H_k = I - tau v v^T (v = [1, essential]) to the trailing columns and to the rhs */
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
  const double thresh = std::numeric_limits<double>::epsilon() * 3.0 * maxpivot;
  int rank = 0;
  for (int k = 0; k < 3; ++k)
    if (std::fabs(M[k][k]) > thresh) ++rank;
    else break;
  double y[3] = {0, 0, 0};
  for (int i = rank - 1; i >= 0; --i) {
    double s = cvec[i];
    for (int c = i + 1; c < rank; ++c) s -= M[i][c] * y[c];
    y[i] = s / M[i][i];
  }
  for (int i = 0; i < 3; ++i) w[perm[i]] = y[i];
}

/* ------------------------------------------------------------------------------------------------
 * RegressionAdapter -- src/infer/adaptive.cpp:27-139
 * ---------------------------------------------------------------------------------------------- */
constexpr double kInitialCount = 10.0;  /* adaptive.cpp:27 */
constexpr double kTempVariance = 10.0;  /* adaptive.cpp:28 */
constexpr u32 kWindow = 1000;           /* adaptive.cpp:29 */

struct Model {
  double mu_xx[9];
  double mu_xy[3];
  double w[3];
  double count;
};

/* adaptive.cpp:31-61 */
void model_init(Model& m, double min_cap, double max_cap) {
  const double rej[3] = {min_cap, 0.0, 1.0};
  const double acc[3] = {max_cap, 0.0, 1.0};
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) m.mu_xx[3 * r + c] = 0.5 * rej[r] * rej[c] + 0.5 * acc[r] * acc[c];
  m.mu_xx[4] = kTempVariance;
  for (int r = 0; r < 3; ++r) {
    m.mu_xy[r] = 0.5 * acc[r];
    m.w[r] = m.mu_xy[r]; /* weight_ = mu_xy, not solved (adaptive.cpp:57) */
  }
  m.count = kInitialCount;
}

/* adaptive.cpp:63-78 (model part of update) */
void model_update(Model& m, double min_cap, double max_cap, double logval, double side, bool acc) {
  logval = std::min(std::max(logval, min_cap), max_cap);
  const double x[3] = {-logval, side, 1.0};
  const double y = acc ? 1.0 : 0.0;
  m.count += 1.0;
  const double alpha = 1.0 / m.count;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) m.mu_xx[3 * r + c] = m.mu_xx[3 * r + c] * (1.0 - alpha) + x[r] * x[c] * alpha;
  for (int r = 0; r < 3; ++r) m.mu_xy[r] = m.mu_xy[r] * (1.0 - alpha) + x[r] * y * alpha;
  qr_solve3(m.mu_xx, m.mu_xy, m.w);
}

/* adaptive.cpp:94-106 */
double model_predict(const double w[3], double opt, double min_cap, double max_cap, double side) {
  const double eps = 1e-3;
  const double denom = std::max(eps, w[0]);
  double numer = -(opt - w[1] * side - w[2]);
  numer = std::max(std::min(numer, denom * max_cap), denom * min_cap);
  return numer / denom;
}

/* Batched form (SLO_SCHED_BATCHED): the running mean of adaptive.cpp:71-75 with alpha=1/count is
 * mu_n = (n0*mu_0 + sum x x^T)/(n0+N); P holds the N-sample sums
 * P = [Sx0x0, Sx0x1, Sx0, Sx1x1, Sx1, N, Sx0y, Sx1y, Sy]. */
void model_update_batched(Model& m, const double P[9]) {
  const double n = P[5];
  if (n == 0.0) return;
  const double c0 = m.count;
  const double c1 = c0 + n;
  const double sxx[9] = {P[0], P[1], P[2], P[1], P[3], P[4], P[2], P[4], P[5]};
  const double sxy[3] = {P[6], P[7], P[8]};
  for (int i = 0; i < 9; ++i) m.mu_xx[i] = (m.mu_xx[i] * c0 + sxx[i]) / c1;
  for (int i = 0; i < 3; ++i) m.mu_xy[i] = (m.mu_xy[i] * c0 + sxy[i]) / c1;
  m.count = c1;
  qr_solve3(m.mu_xx, m.mu_xy, m.w);
}

inline void stats_accumulate(double P[9], double min_cap, double max_cap, double logval, double side, bool acc) {
  logval = std::min(std::max(logval, min_cap), max_cap);
  const double x0 = -logval, x1 = side, y = acc ? 1.0 : 0.0;
  P[0] += x0 * x0;
  P[1] += x0 * x1;
  P[2] += x0;
  P[3] += x1 * x1;
  P[4] += x1;
  P[5] += 1.0;
  P[6] += x0 * y;
  P[7] += x1 * y;
  P[8] += y;
}

/* Fixed-shape reduction used for the cross-stack sums of the batched schedule: slot i<256 sums
 * v[i], v[i+256], ... in order, then a halving tree 128,64,..,1.  The device reduce kernel has the same shape. */
double reduce_tree256(const double* v, u32 n, u32 stride) {
  double a[256];
  for (u32 i = 0; i < 256; ++i) {
    double acc = 0.0;
    for (u32 m = i; m < n; m += 256) acc += v[(size_t)m * stride];
    a[i] = acc;
  }
  for (u32 s = 128; s > 0; s >>= 1)
    for (u32 i = 0; i < s; ++i) a[i] += a[i + s];
  return a[0];
}

struct Regression {
  u32 nTemps = 0, nModels = 0;
  bool perChain = false; /* SLO_SCHED_PER_STACK: model index = chain id */
  double min_cap = 0, max_cap = 0, opt = 0;
  std::vector<Model> models;
  std::vector<double> values; /* values_, per chain, init 1 (adaptive.cpp:37) */
  std::vector<double> rates;  /* rates_, per chain, logging only */
  std::vector<std::deque<int>> window;
  std::vector<long> window_sum;

  void init(u32 nStacks, u32 nT, double optRate, double mn, double mx, bool per_chain) {
    nTemps = nT; perChain = per_chain; min_cap = mn; max_cap = mx; opt = optRate;
    nModels = per_chain ? nStacks * nT : nT;
    models.resize(nModels);
    for (auto& m : models) model_init(m, mn, mx);
    values.assign((size_t)nStacks * nT, 1.0);
    rates.assign((size_t)nStacks * nT, 0.0);
    window.assign((size_t)nStacks * nT, {});
    window_sum.assign((size_t)nStacks * nT, 0);
  }
  u32 midx(u32 chain) const { return perChain ? chain : chain % nTemps; }
  /* adaptive.cpp:63-91 */
  void update(u32 chain, double logval, double side, bool acc) {
    model_update(models[midx(chain)], min_cap, max_cap, logval, side, acc);
    log_window(chain, acc);
  }
  void log_window(u32 chain, bool acc) { /* adaptive.cpp:80-90 */
    const int ia = acc;
    window[chain].push_back(ia);
    window_sum[chain] += ia;
    const u32 n = (u32)window[chain].size();
    if (n >= kWindow) {
      window_sum[chain] -= window[chain][0];
      window[chain].pop_front();
    }
    rates[chain] = (double)window_sum[chain] / (double)n;
  }
  double predict(u32 chain, double side) const {
    return model_predict(models[midx(chain)].w, opt, min_cap, max_cap, side);
  }
  /* adaptive.cpp:108-116 */
  void betaUpdate(u32 chain, double bl, double bh, bool acc) { update(chain, slm_log(bl) - slm_log(bh), slm_log(bl), acc); }
  /* adaptive.cpp:118-132; chain = coldest chain of its stack */
  void computeBetaStack(u32 chain) {
    double logbeta = 0.0;
    for (u32 i = 1; i < nTemps; ++i) {
      logbeta -= predict(chain + i - 1, logbeta);
      values[chain + i] = slm_exp(logbeta);
    }
  }
  /* adaptive.cpp:134-139 */
  double computeSigma(u32 chain, double t) {
    const double sigma = slm_exp(predict(chain, t));
    values[chain] = sigma;
    return sigma;
  }
};

/* ------------------------------------------------------------------------------------------------
 * ProposalShaper -- src/infer/adaptive.cpp:155-209.  Dense row-major d x d, lower triangle used.
 * Frobenius norm: R.norm() (adaptive.cpp:200).  Eigen's summation order is not pinned by any
 * reference test; the canonical order here (and in the kernels) is: every row r sums the squares of
 * its lower-triangle entries left to right with fma (starting from 0); slot l<32 adds the row sums of
 * rows l, l+32, ... in that order; then a halving tree 16,8,4,2,1 over the 32 slots.
 * ---------------------------------------------------------------------------------------------- */
double frobenius_lower(const double* L, u32 d) {
  double slot[32];
  for (u32 l = 0; l < 32; ++l) slot[l] = 0.0;
  for (u32 r = 0; r < d; ++r) {
    double acc = 0.0;
    for (u32 k = 0; k <= r; ++k) acc = std::fma(L[(size_t)r * d + k], L[(size_t)r * d + k], acc);
    slot[r & 31u] += acc;
  }
  for (u32 m = 16; m > 0; m >>= 1)
    for (u32 l = 0; l < m; ++l) slot[l] += slot[l + m];
  return std::sqrt(slot[0]);
}

void shaper_update(double* L, double* N, double& count, const double* step, u32 d, double prop_norm) {
  const double eps = 1e-18;
  count += 1.0; /* count_ is uint in the reference; exact in double far beyond any run length */
  std::vector<double> x(d);
  const double sq = std::sqrt(count);
  for (u32 i = 0; i < d; ++i) x[i] = step[i] / sq;
  const double scale = std::sqrt((count - 1.0) / count);
  for (u32 r = 0; r < d; ++r)
    for (u32 k = 0; k <= r; ++k) L[(size_t)r * d + k] *= scale;
  for (u32 k = 0; k < d; ++k) {
    const double V = std::max(eps, L[(size_t)k * d + k]);
    const double r = std::sqrt(V * V + x[k] * x[k]);
    const double c = r / V;
    const double s = x[k] / V;
    L[(size_t)k * d + k] = r;
    for (u32 j = k + 1; j < d; ++j) {
      double& Ljk = L[(size_t)j * d + k];
      Ljk = (Ljk + s * x[j]) / c;
      x[j] = c * x[j] - s * Ljk;
    }
  }
  const double nscale = prop_norm / std::max(frobenius_lower(L, d), eps);
  for (u32 r = 0; r < d; ++r) {
    for (u32 k = 0; k < d; ++k) N[(size_t)r * d + k] = (k <= r) ? L[(size_t)r * d + k] * nscale : 0.0;
    N[(size_t)r * d + r] += 1e-4; /* adaptive.cpp:203-204 */
  }
}

/* ------------------------------------------------------------------------------------------------
 * Synthetic targets.  The likelihood contract is workerwrapper.hpp:21:
 * double f(uint jobType, const std::vector<double>& x), a negative log-likelihood.
 * The CUDA plugins mirror the operation order of each of these.
 * ---------------------------------------------------------------------------------------------- */
double target_eval(int target, const double* prm, u32 d, u32 J, u32 job, const double* x) {
  switch (target) {
    case SLO_TARGET_DEMO_GAUSS: { /* demo-worker.cpp:37-45 */
      double sq = 0.0;
      for (u32 i = 0; i < d; ++i) sq += x[i] * x[i];
      return 0.5 * sq;
    }
    case SLO_TARGET_DOUBLE_WELL: {
      if (job == 0) {
        const double t = x[0] * x[0] - 1.0;
        return 8.0 * (t * t);
      }
      return 0.5 * (x[1] * x[1]);
    }
    case SLO_TARGET_GAUSS_FACT: {
      const double* mu = prm;
      const double* is = prm + d;
      const u32 a = 2 * job, b = 2 * job + 1;
      const double ta = (x[a] - mu[a]) * is[a];
      double acc = ta * ta;
      if (b < d) {
        const double tb = (x[b] - mu[b]) * is[b];
        acc = std::fma(tb, tb, acc);
      }
      return 0.5 * acc;
    }
    case SLO_TARGET_GAUSS_CORR: {
      const u32 per = (d + J - 1) / J;
      const u32 r0 = job * per, r1 = std::min(d, r0 + per);
      double acc = 0.0;
      for (u32 r = r0; r < r1; ++r) {
        double dot = 0.0;
        for (u32 c = r; c < d; ++c) dot = std::fma(prm[(size_t)r * d + c], x[c], dot);
        acc = std::fma(dot, dot, acc);
      }
      return 0.5 * acc;
    }
    case SLO_TARGET_ROSENBROCK: {
      const u32 nterm = d - 1;
      const u32 per = (nterm + J - 1) / J;
      const u32 i0 = job * per, i1 = std::min(nterm, i0 + per);
      double acc = 0.0;
      for (u32 i = i0; i < i1; ++i) {
        const double a = x[i + 1] - x[i] * x[i];
        const double b = 1.0 - x[i];
        acc += (100.0 * (a * a) + b * b) / 20.0;
      }
      return acc;
    }
  }
  return std::numeric_limits<double>::quiet_NaN();
}

/* std::to_string(double) is "%f" (requester.cpp:37, minion.cpp:53); std::stod back */
inline double wire(double v) {
  char buf[400];
  std::snprintf(buf, sizeof buf, "%f", v);
  return std::strtod(buf, nullptr);
}

/* ------------------------------------------------------------------------------------------------
 * EPSRDiagnostic -- src/infer/diagnostics.cpp:27-76
 * ---------------------------------------------------------------------------------------------- */
struct Epsr {
  u32 nStacks = 0, d = 0;
  std::vector<double> M, S; /* [stack][d] */
  std::vector<int> n;
  void init(u32 s, u32 dims) { nStacks = s; d = dims; M.assign((size_t)s * dims, 0.0); S.assign((size_t)s * dims, 0.0); n.assign(s, 0); }
  void update(u32 stack, const double* x) { /* diagnostics.cpp:27-47 */
    const int nn = n[stack] + 1;
    for (u32 i = 0; i < d; ++i) {
      double& m = M[(size_t)stack * d + i];
      double& s = S[(size_t)stack * d + i];
      const double newM = m + (x[i] - m) / nn;
      s = s + (x[i] - m) * (x[i] - newM);
      m = newM;
    }
    n[stack] = nn;
  }
  void rhat(double* out) const { /* diagnostics.cpp:49-71 */
    const int nmin = *std::min_element(n.begin(), n.end());
    const int m = (int)nStacks;
    for (u32 i = 0; i < d; ++i) {
      double mean = 0.0;
      for (u32 s = 0; s < nStacks; ++s) mean += M[(size_t)s * d + i];
      mean /= m;
      double b = 0.0, w = 0.0;
      for (u32 s = 0; s < nStacks; ++s) {
        const double dm = M[(size_t)s * d + i] - mean;
        b += dm * dm;
        w += (1.0 / m) * S[(size_t)s * d + i] / (nmin - 1.0);
      }
      b *= nmin / (m - 1.0);
      const double vHat = ((nmin - 1.0) / nmin) * w + (1.0 / nmin) * b;
      out[i] = std::sqrt(vHat / (w + 1e-30));
    }
  }
};

/* State row: datatypes.hpp:40-62 */
struct Row {
  std::vector<double> x;
  double energy = 0, sigma = 0, beta = 0;
  int accepted = 0;
  int swapType = 0; /* 0 NoAttempt, 1 Accept, 2 Reject (datatypes.hpp:29-38) */
};

}  // namespace

/* ------------------------------------------------------------------------------------------------
 * The sampler
 * ---------------------------------------------------------------------------------------------- */
struct slo_sampler {
  slo_config cfg;
  std::vector<double> mn, mx, initial, tparams;
  u32 S, T, d, J, C;
  u32 swapInterval, adaptInterval;
  /* replay */
  const double *rz = nullptr, *rumh = nullptr, *ruswap = nullptr, *rxinit = nullptr;
  u64 replayRounds = 0;
  /* ChainArray (chainarray.hpp:39-153): last row, lengths, sigma_/beta_ */
  std::vector<Row> last;
  std::vector<u64> length;
  std::vector<double> sigma_, beta_;
  Regression sigAd, betaAd;
  /* ProposalShaper state */
  std::vector<std::vector<double>> L, N;
  std::vector<double> shCount;
  double propNorm = 0;
  /* Sampler (sampler.hpp:106-146) */
  std::vector<std::vector<double>> prop;
  std::vector<char> locked;
  std::deque<u32> fifo;
  u64 outstanding = 0;
  u64 rounds = 0;
  /* batched schedule: per (stack,tier) partial sums, frozen ladder */
  std::vector<double> Psig, Pbeta; /* [S][T][9] */
  std::vector<double> ladder;      /* [T] */
  /* TableLogger counters + EPSR (logging.cpp:29-48) */
  std::vector<u64> lgLength, lgAccepts, lgSwaps, lgSwapAttempts;
  std::vector<double> lgMinE, lgCurE; /* minEnergies_, energies_ (logging.cpp:22-23) */
  Epsr epsr;
  /* cold rows as CSVChainArrayWriter would receive them (chainarray.cpp:131-155) */
  std::vector<std::vector<double>> cold;

  u32 gchain(u32 id) const { return cfg.stack_offset * T + id; }

  double energy_of(const double* x) const {
    /* Requester round trip: E = sum_j f(j, x) in job-type order (sampler.cpp:148-150, delegator.cpp:163-168) */
    std::vector<double> xq;
    const double* xe = x;
    if (cfg.wire_quantize) {
      xq.resize(d);
      for (u32 i = 0; i < d; ++i) xq[i] = wire(x[i]);
      xe = xq.data();
    }
    double e = 0.0;
    for (u32 j = 0; j < J; ++j) {
      double f = target_eval(cfg.target, tparams.data(), d, J, j, xe);
      if (cfg.wire_quantize) f = wire(f);
      e += f;
    }
    return e;
  }

  void get_z(u64 r, u32 id, double* z) const {
    if (cfg.rng_mode == SLO_RNG_REPLAY) std::memcpy(z, rz + ((size_t)r * C + id) * d, sizeof(double) * d);
    else draw_normals(cfg.seed, gchain(id), r, d, z);
  }
  double get_umh(u64 r, u32 id) const {
    return cfg.rng_mode == SLO_RNG_REPLAY ? rumh[(size_t)r * C + id] : draw_uniform(cfg.seed, gchain(id), r, 1);
  }
  double get_uswap(u64 r, u32 id) const {
    return cfg.rng_mode == SLO_RNG_REPLAY ? ruswap[(size_t)r * C + id] : draw_uniform(cfg.seed, gchain(id), r, 2);
  }

  /* GaussianProposal::propose + boundedPropose, sampler.cpp:79-93: x' = bounce(x + N z sigma).
   * Canonical accumulation order of the dense GEMV: k ascending with fma (Eigen's order is unpinned). */
  void propose_from(u32 id, const double* x, double sigma, u64 r, double* out) const {
    std::vector<double> z(d), raw(d);
    get_z(r, id, z.data());
    const double* Nm = N[id].data();
    for (u32 i = 0; i < d; ++i) {
      double acc = 0.0;
      for (u32 k = 0; k <= i; ++k) acc = std::fma(Nm[(size_t)i * d + k], z[k], acc);
      raw[i] = x[i] + acc * sigma;
    }
    bouncy_bounds(raw.data(), mn.data(), mx.data(), d, out);
  }

  /* acceptProposal, chainarray.cpp:30-45 */
  bool accept_proposal(double eNew, double eOld, double beta, double u) const {
    if (std::isinf(eNew)) return false;
    const double deltaEnergy = eNew - eOld;
    const double p = slm_exp(-1.0 * beta * deltaEnergy);
    return u < p;
  }
  /* acceptSwap as called from ChainArray::swap, chainarray.cpp:55-67,171: the call passes
   * (stateh, statel, beta_h, beta_l) into (stateLow, stateHigh, betaLow, betaHigh). */
  bool accept_swap(double eL, double eH, double bL, double bH, double u) const {
    const double deltaEnergy = eL - eH;
    const double deltaBeta = bL - bH;
    const double p = slm_exp(deltaEnergy * deltaBeta);
    return u < p;
  }

  /* ChainArray::append, chainarray.cpp:94-121 */
  bool append(u32 id, const std::vector<double>& sample, double energy, u64 r) {
    Row& lastRow = last[id];
    const bool accepted = accept_proposal(energy, lastRow.energy, beta_[id], get_umh(r, id));
    if (accepted) {
      lastRow.x = sample;
      lastRow.energy = energy;
      lastRow.sigma = sigma_[id];
      lastRow.beta = beta_[id];
    }
    lastRow.accepted = accepted;
    lastRow.swapType = 0;
    length[id] += 1;
    return accepted;
  }

  /* ChainArray::swap, chainarray.cpp:162-189 */
  bool swap(u32 lId, u32 hId, u64 r) {
    Row& rl = last[lId];
    Row& rh = last[hId];
    const bool swapped = accept_swap(rl.energy, rh.energy, beta_[lId], beta_[hId], get_uswap(r, lId));
    if (swapped) {
      std::swap(rh.x, rl.x);
      std::swap(rh.energy, rl.energy);
      std::swap(rh.accepted, rl.accepted);
      rl.swapType = 1;
    } else {
      rl.swapType = 2;
    }
    return swapped;
  }

  void shaper_step(u32 id, const double* step) { shaper_update(L[id].data(), N[id].data(), shCount[id], step, d, propNorm); }

  /* TableLogger::update, logging.cpp:35-48, fed with the state step() returns */
  void log_update(u32 id, const Row& s) {
    lgLength[id] += 1;
    lgMinE[id] = std::min(lgMinE[id], s.energy);
    lgCurE[id] = s.energy; /* logging.cpp:43 */
    lgAccepts[id] += s.accepted;
    lgSwaps[id] += (s.swapType == 1);
    lgSwapAttempts[id] += (s.swapType != 0);
    if (id % T == 0) { /* diagnostics.cpp:30 */
      epsr.update(id / T, s.x.data());
      emit_cold(id / T, s);
    }
  }
  void emit_cold(u32 stack, const Row& s) {
    auto& v = cold[stack];
    v.insert(v.end(), s.x.begin(), s.x.end());
    v.push_back(s.energy); v.push_back(s.sigma); v.push_back(s.beta);
    v.push_back((double)s.accepted); v.push_back((double)s.swapType);
  }

  /* ---------------- sequential schedules: Sampler::step state machine ---------------- */
  /* Sampler::propose, sampler.cpp:206-213 */
  void propose(u32 id) {
    const double sigma = sigAd.values[id];
    const u64 r = length[id] - 1; /* the round that will evaluate this proposal */
    propose_from(id, last[id].x.data(), sigma, r, prop[id].data());
    fifo.push_back(id); /* requester_.submit: FIFO job list (delegator.cpp:166) */
    outstanding++;
  }
  /* Sampler::unlock, sampler.cpp:237-257 */
  void unlock(u32 id) {
    locked[id] = 0;
    propose(id + 1);
    if (id % T != 0) locked[id - 1] = 1;
    else propose(id);
  }
  /* Sampler::step, sampler.cpp:142-203, given the retrieved (id, energy) */
  void advance(u32 id, double energy) {
    outstanding--;
    const u64 r = length[id] - 1;
    const std::vector<double> prevx = last[id].x; /* previous_state (sampler.cpp:154) */
    append(id, prop[id], energy, r);
    const Row state = last[id]; /* copy, as the reference takes lastState() by value (sampler.cpp:156) */
    /* sigma adaptation, sampler.cpp:160-162 */
    const double logtemper = -slm_log(state.beta);
    sigAd.update(id, slm_log(state.sigma), logtemper, state.accepted);
    sigma_[id] = sigAd.computeSigma(id, logtemper);
    /* proposal shape, sampler.cpp:165-166 */
    if (state.accepted) {
      std::vector<double> stepv(d);
      for (u32 i = 0; i < d; ++i) stepv[i] = state.x[i] - prevx[i];
      shaper_step(id, stepv.data());
    }
    /* swapping logic, sampler.cpp:169-199 */
    if (locked[id]) {
      const bool swapped = swap(id, id + 1, r);
      unlock(id);
      betaAd.betaUpdate(id, beta_[id], beta_[id + 1], swapped);
      if (id % T == 0) betaAd.computeBetaStack(id);
      beta_[id + 1] = betaAd.values[id + 1];
    } else if (id % T == T - 1 && length[id] % swapInterval == 0 && T > 1) {
      locked[id - 1] = 1;
    } else {
      propose(id);
    }
    log_update(id, last[id]); /* serverwrapper.cpp:124-126 with step()'s return value */
  }
  void step_sequential_inline() {
    const u32 id = fifo.front();
    fifo.pop_front();
    advance(id, energy_of(prop[id].data()));
  }

  /* ---------------- batched schedule (DESIGN.md "batched") ---------------- */
  void adapt_point() {
    for (u32 t = 0; t < T; ++t) {
      double P[9];
      for (int k = 0; k < 9; ++k) P[k] = reduce_tree256(Psig.data() + (size_t)t * 9 + k, S, T * 9);
      model_update_batched(sigAd.models[t], P);
      for (int k = 0; k < 9; ++k) P[k] = reduce_tree256(Pbeta.data() + (size_t)t * 9 + k, S, T * 9);
      model_update_batched(betaAd.models[t], P);
    }
    std::fill(Psig.begin(), Psig.end(), 0.0);
    std::fill(Pbeta.begin(), Pbeta.end(), 0.0);
    compute_ladder();
  }
  void compute_ladder() { /* computeBetaStack with the shared models, adaptive.cpp:118-132 */
    double logbeta = 0.0;
    ladder[0] = 1.0;
    for (u32 i = 1; i < T; ++i) {
      logbeta -= model_predict(betaAd.models[i - 1].w, betaAd.opt, betaAd.min_cap, betaAd.max_cap, logbeta);
      ladder[i] = slm_exp(logbeta);
    }
  }
  void round_batched() {
    const u64 r = rounds; /* 0-based */
    std::vector<double> xp(d), stepv(d);
    for (u32 s = 0; s < S; ++s) {
      for (u32 t = 0; t < T; ++t) {
        const u32 id = s * T + t;
        propose_from(id, last[id].x.data(), sigAd.values[id], r, xp.data());
        const double e = energy_of(xp.data());
        const std::vector<double> prevx = last[id].x;
        const std::vector<double> sample(xp.begin(), xp.end());
        const bool acc = append(id, sample, e, r);
        const Row& row = last[id];
        const double logtemper = -slm_log(row.beta);
        stats_accumulate(&Psig[((size_t)s * T + t) * 9], sigAd.min_cap, sigAd.max_cap, slm_log(row.sigma), logtemper, acc);
        sigAd.log_window(id, acc);
        const double sigma = slm_exp(model_predict(sigAd.models[t].w, sigAd.opt, sigAd.min_cap, sigAd.max_cap, logtemper));
        sigAd.values[id] = sigma;
        sigma_[id] = sigma;
        if (acc) {
          for (u32 i = 0; i < d; ++i) stepv[i] = row.x[i] - prevx[i];
          shaper_step(id, stepv.data());
        }
      }
      const bool swapRound = (T > 1) && (length[s * T + T - 1] % swapInterval == 0);
      /* the state step() returns for chain t is taken after swap(t,t+1) and before swap(t-1,t) */
      if (swapRound) {
        log_update(s * T + T - 1, last[s * T + T - 1]);
        for (int t = (int)T - 2; t >= 0; --t) {
          const u32 id = s * T + t;
          const bool swapped = swap(id, id + 1, r);
          const double bl = beta_[id], bh = beta_[id + 1];
          stats_accumulate(&Pbeta[((size_t)s * T + t) * 9], betaAd.min_cap, betaAd.max_cap,
                           slm_log(bl) - slm_log(bh), slm_log(bl), swapped);
          betaAd.log_window(id, swapped);
          log_update(id, last[id]);
        }
        for (u32 t = 1; t < T; ++t) {
          beta_[s * T + t] = ladder[t];
          betaAd.values[s * T + t] = ladder[t];
        }
      } else {
        for (int t = (int)T - 1; t >= 0; --t) log_update(s * T + t, last[s * T + t]);
      }
    }
    rounds++;
    if ((rounds + 1) % adaptInterval == 0) adapt_point(); /* chain length == rounds+1 */
  }
};

extern "C" {

slo_sampler* slo_create(const slo_config* cfg) {
  auto* o = new slo_sampler();
  o->cfg = *cfg;
  o->S = cfg->n_stacks; o->T = cfg->n_temps; o->d = cfg->n_dims; o->J = cfg->n_job_types;
  o->C = o->S * o->T;
  o->swapInterval = cfg->swap_interval;
  o->adaptInterval = cfg->adapt_interval ? cfg->adapt_interval : cfg->swap_interval;
  o->mn.assign(cfg->min, cfg->min + o->d);
  o->mx.assign(cfg->max, cfg->max + o->d);
  if (cfg->use_initial) o->initial.assign(cfg->initial, cfg->initial + o->d);
  size_t np = 0;
  if (cfg->target == SLO_TARGET_GAUSS_FACT) np = 2 * (size_t)o->d;
  if (cfg->target == SLO_TARGET_GAUSS_CORR) np = (size_t)o->d * o->d;
  if (np) o->tparams.assign(cfg->target_params, cfg->target_params + np);
  return o;
}

void slo_destroy(slo_sampler* o) { delete o; }

void slo_set_replay(slo_sampler* o, const double* z, const double* u_mh, const double* u_swap,
                    const double* x_init, uint64_t n_rounds) {
  o->rz = z; o->rumh = u_mh; o->ruswap = u_swap; o->rxinit = x_init; o->replayRounds = n_rounds;
}

/* runSampler up to the Sampler ctor, serverwrapper.cpp:48-87 */
void slo_init_chains(slo_sampler* o) {
  const u32 S = o->S, T = o->T, d = o->d, C = o->C;
  const bool perStack = o->cfg.schedule == SLO_SCHED_PER_STACK;
  o->sigAd.init(S, T, o->cfg.opt_accept, -8.0, 4.0, perStack); /* serverwrapper.cpp:48-51 */
  o->betaAd.init(S, T, o->cfg.opt_swap, 0.0, 4.0, perStack);   /* serverwrapper.cpp:52 */
  /* ProposalShaper ctor, adaptive.cpp:155-171 */
  std::vector<double> range(d);
  double n2 = 0.0;
  for (u32 i = 0; i < d; ++i) {
    range[i] = (o->mx[i] - o->mn[i]) / 4.0;
    n2 += range[i] * range[i];
  }
  o->propNorm = std::sqrt(n2);
  o->L.assign(C, std::vector<double>((size_t)d * d, 0.0));
  o->N.assign(C, std::vector<double>((size_t)d * d, 0.0));
  o->shCount.assign(C, 1000.0); /* serverwrapper.cpp:50 */
  for (u32 c = 0; c < C; ++c)
    for (u32 i = 0; i < d; ++i) {
      o->L[c][(size_t)i * d + i] = range[i];
      o->N[c][(size_t)i * d + i] = range[i] / o->propNorm;
    }
  o->last.assign(C, Row());
  o->length.assign(C, 0);
  o->sigma_.assign(C, 0.0);
  o->beta_.assign(C, 0.0);
  o->prop.assign(C, std::vector<double>(d, 0.0));
  o->locked.assign(C, 0);
  o->lgLength.assign(C, 1);
  o->lgAccepts.assign(C, 1);
  o->lgSwaps.assign(C, 0);
  o->lgSwapAttempts.assign(C, 1);
  o->lgMinE.assign(C, std::numeric_limits<double>::infinity());
  o->lgCurE.assign(C, 0.0);
  o->epsr.init(S, d);
  o->cold.assign(S, {});
  o->Psig.assign((size_t)C * 9, 0.0);
  o->Pbeta.assign((size_t)C * 9, 0.0);
  o->ladder.assign(T, 1.0);
  /* init loop, serverwrapper.cpp:60-81 */
  std::vector<double> raw(d), x0(d);
  for (u32 i = 0; i < C; ++i) {
    if (o->cfg.use_initial) raw = o->initial;
    else if (o->cfg.rng_mode == SLO_RNG_REPLAY) std::memcpy(raw.data(), o->rxinit + (size_t)i * d, sizeof(double) * d);
    else draw_init(o->cfg.seed, o->gchain(i), d, raw.data());
    bouncy_bounds(raw.data(), o->mn.data(), o->mx.data(), d, x0.data());
    const double e = o->energy_of(x0.data());
    if (i % T == 0) o->betaAd.computeBetaStack(i);
    Row& row = o->last[i];
    row.x = x0; row.energy = e;
    row.sigma = o->sigAd.values[i]; row.beta = o->betaAd.values[i];
    row.accepted = 1; row.swapType = 0; /* chainarray.cpp:123-129 */
    o->sigma_[i] = row.sigma; o->beta_[i] = row.beta;
    o->length[i] = 1;
    if (i % T == 0) o->emit_cold(i / T, row); /* the initial state is row 0 of the CSV */
  }
  if (o->cfg.schedule == SLO_SCHED_BATCHED) {
    o->compute_ladder();
  } else {
    /* Sampler ctor, sampler.cpp:128-132: hottest to coldest */
    for (u32 i = 0; i < C; ++i) o->propose(C - i - 1);
  }
  o->rounds = 0;
}

void slo_run_rounds(slo_sampler* o, uint64_t n) {
  if (o->cfg.schedule == SLO_SCHED_BATCHED) {
    for (u64 i = 0; i < n; ++i) o->round_batched();
  } else {
    for (u64 i = 0; i < n; ++i) {
      for (u32 c = 0; c < o->C; ++c) o->step_sequential_inline();
      o->rounds++;
    }
  }
}

/* Sampler::flush, sampler.cpp:216-235: outstanding proposals are appended without adaptation */
void slo_flush(slo_sampler* o) {
  if (o->cfg.schedule == SLO_SCHED_BATCHED) return;
  while (!o->fifo.empty()) {
    const u32 id = o->fifo.front();
    o->fifo.pop_front();
    o->outstanding--;
    o->append(id, o->prop[id], o->energy_of(o->prop[id].data()), o->length[id] - 1);
  }
}

uint64_t slo_rounds_done(const slo_sampler* o) { return o->rounds; }

void slo_get_last_rows(const slo_sampler* o, double* rows) {
  const u32 w = o->d + 5;
  for (u32 c = 0; c < o->C; ++c) {
    const Row& r = o->last[c];
    double* p = rows + (size_t)c * w;
    std::memcpy(p, r.x.data(), sizeof(double) * o->d);
    p[o->d] = r.energy; p[o->d + 1] = r.sigma; p[o->d + 2] = r.beta;
    p[o->d + 3] = r.accepted; p[o->d + 4] = r.swapType;
  }
}

void slo_get_sigma_beta(const slo_sampler* o, double* sigma, double* beta) {
  std::memcpy(sigma, o->sigma_.data(), sizeof(double) * o->C);
  std::memcpy(beta, o->beta_.data(), sizeof(double) * o->C);
}

void slo_get_shaper(const slo_sampler* o, uint32_t chain, double* L, double* N, double* count) {
  const size_t n = (size_t)o->d * o->d;
  if (L) std::memcpy(L, o->L[chain].data(), sizeof(double) * n);
  if (N) std::memcpy(N, o->N[chain].data(), sizeof(double) * n);
  if (count) *count = o->shCount[chain];
}

uint32_t slo_n_models(const slo_sampler* o) { return o->sigAd.nModels; }

void slo_get_models(const slo_sampler* o, int which, double* mu_xx, double* mu_xy, double* weight, double* count) {
  const Regression& a = which ? o->betaAd : o->sigAd;
  for (u32 m = 0; m < a.nModels; ++m) {
    if (mu_xx) std::memcpy(mu_xx + 9 * m, a.models[m].mu_xx, sizeof(double) * 9);
    if (mu_xy) std::memcpy(mu_xy + 3 * m, a.models[m].mu_xy, sizeof(double) * 3);
    if (weight) std::memcpy(weight + 3 * m, a.models[m].w, sizeof(double) * 3);
    if (count) count[m] = a.models[m].count;
  }
}

void slo_get_counters(const slo_sampler* o, uint64_t* length, uint64_t* accepts, uint64_t* swaps,
                      uint64_t* swap_attempts, double* min_energy) {
  for (u32 c = 0; c < o->C; ++c) {
    if (length) length[c] = o->lgLength[c];
    if (accepts) accepts[c] = o->lgAccepts[c];
    if (swaps) swaps[c] = o->lgSwaps[c];
    if (swap_attempts) swap_attempts[c] = o->lgSwapAttempts[c];
    if (min_energy) min_energy[c] = o->lgMinE[c];
  }
}

void slo_get_logger(const slo_sampler* o, double* cur_energy, double* accept_rate, double* swap_rate) {
  for (u32 c = 0; c < o->C; ++c) {
    if (cur_energy) cur_energy[c] = o->lgCurE[c];
    if (accept_rate) accept_rate[c] = o->sigAd.rates[c]; /* sigmaAdapter.rates(), sampler.cpp:198 */
    if (swap_rate) swap_rate[c] = o->betaAd.rates[c];
  }
}

/* the logging window of RegressionAdapter::update (adaptive.cpp:80-90) alone: rate after each of n flags */
void slo_window_rates(const int* flags, uint32_t n, double* rates) {
  Regression r;
  r.init(1, 1, 0.5, 0.0, 1.0, false);
  for (u32 i = 0; i < n; ++i) {
    r.log_window(0, flags[i] != 0);
    rates[i] = r.rates[0];
  }
}

/* TableLogger::update's print branch, logging.cpp:52-87, on the current counters.  The reference streams into a
 * stringstream with fill ' ', precision(5), showpoint | fixed and setw(4 / 9 / 10): by the num_put rules that is
 * printf's "%4u", "%9lu" and "%#10.5f" ('#' is a no-op with a non-zero precision).  The convergence line goes to
 * std::cout with default flags: "%g".  (printf rather than iostream: this library may be linked against a static
 * libstdc++ whose stream locale is never initialised.) */
size_t slo_format_table(const slo_sampler* o, char* out, size_t cap) {
  std::string s = "\n\n  ID    Length    MinEngy   CurrEngy      Sigma     AcptRt  GlbAcptRt       Beta     SwapRt  GlbSwapRt\n";
  s += "------------------------------------------------------------------------------------------------------\n";
  char cell[64];
  auto num = [&](double v) {
    std::snprintf(cell, sizeof cell, "%#10.5f ", v);
    s += cell;
  };
  for (u32 i = 0; i < o->C; i++) {
    if (i % o->T == 0 && i != 0) s += '\n';
    std::snprintf(cell, sizeof cell, "%4u %9lu ", i, (unsigned long)o->lgLength[i]);
    s += cell;
    num(o->lgMinE[i]);
    num(o->lgCurE[i]);
    /* the Sigma and Beta columns are the ADAPTERS' values() (serverwrapper.cpp:124-126), not ChainArray's: after a
     * cascade computeBetaStack has already rewritten values_[2..T-1] while those chains keep their old beta until
     * the next cascade reaches them (sampler.cpp:179-186), so the table shows the ladder to come.  Found by running
     * the reference's own TableLogger (tests/test_ref_pin.py); in the batched schedule both are always equal. */
    num(o->sigAd.values[i]);
    num(o->sigAd.rates[i]);
    num(o->lgAccepts[i] / (double)o->lgLength[i]);
    num(o->betaAd.values[i]);
    num(o->betaAd.rates[i]);
    num(o->lgSwaps[i] / (double)o->lgSwapAttempts[i]);
    s += '\n';
  }
  s += '\n'; /* std::cout << ss.str() << std::endl */
  std::vector<double> rh(o->d);
  o->epsr.rhat(rh.data());
  double mean = 0.0;
  bool conv = true;
  for (u32 i = 0; i < o->d; ++i) {
    mean += rh[i];
    conv = conv && rh[i] < 1.1; /* diagnostics.hpp:35, diagnostics.cpp:73-76 */
  }
  mean /= (double)o->d;
  std::snprintf(cell, sizeof cell, "Convergence test: %g (", mean);
  s += cell;
  s += conv ? "possibly converged" : "not converged";
  s += ")\n";
  if (out && cap) {
    const size_t n = std::min(cap - 1, s.size());
    std::memcpy(out, s.data(), n);
    out[n] = 0;
  }
  return s.size();
}

uint64_t slo_cold_rows(const slo_sampler* o, uint32_t stack, double* rows, uint64_t cap_rows) {
  const u32 w = o->d + 5;
  const u64 n = o->cold[stack].size() / w;
  if (rows) std::memcpy(rows, o->cold[stack].data(), sizeof(double) * w * std::min(n, cap_rows));
  return n;
}

void slo_rhat(const slo_sampler* o, double* rhat) { o->epsr.rhat(rhat); }

/* ---- stand-alone pieces ---- */
void slo_bouncy_bounds(const double* val, const double* mn, const double* mx, uint32_t d, double* out) {
  bouncy_bounds(val, mn, mx, d, out);
}
void slo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox4x32_10(ctr, key, out); }
void slo_draw_normals(uint64_t seed, uint32_t chain, uint64_t round, uint32_t d, double* z) { draw_normals(seed, chain, round, d, z); }
double slo_draw_uniform(uint64_t seed, uint32_t chain, uint64_t round, uint32_t purpose) { return draw_uniform(seed, chain, round, purpose); }
void slo_draw_init(uint64_t seed, uint32_t chain, uint32_t d, double* x) { draw_init(seed, chain, d, x); }
void slo_qr_solve3(const double A[9], const double b[3], double w[3]) { qr_solve3(A, b, w); }
void slo_shaper_update(double* L, double* N, double* count, const double* step, uint32_t d, double prop_norm) {
  shaper_update(L, N, *count, step, d, prop_norm);
}
void slo_epsr(const double* samples, uint32_t n, uint32_t m, uint32_t d, double* rhat) {
  Epsr e;
  e.init(m, d);
  for (u32 i = 0; i < n; ++i)
    for (u32 c = 0; c < m; ++c) e.update(c, samples + ((size_t)i * m + c) * d);
  e.rhat(rhat);
}
void slo_regression_init(double min_cap, double max_cap, double mu_xx[9], double mu_xy[3], double w[3], double* count) {
  Model m;
  model_init(m, min_cap, max_cap);
  std::memcpy(mu_xx, m.mu_xx, sizeof m.mu_xx);
  std::memcpy(mu_xy, m.mu_xy, sizeof m.mu_xy);
  std::memcpy(w, m.w, sizeof m.w);
  *count = m.count;
}
double slo_regression_predict(const double w[3], double opt, double min_cap, double max_cap, double side) {
  return model_predict(w, opt, min_cap, max_cap, side);
}
void slo_regression_update(double mu_xx[9], double mu_xy[3], double w[3], double* count, double min_cap,
                           double max_cap, double logval, double side, int acc) {
  Model m;
  std::memcpy(m.mu_xx, mu_xx, sizeof m.mu_xx);
  std::memcpy(m.mu_xy, mu_xy, sizeof m.mu_xy);
  std::memcpy(m.w, w, sizeof m.w);
  m.count = *count;
  model_update(m, min_cap, max_cap, logval, side, acc != 0);
  std::memcpy(mu_xx, m.mu_xx, sizeof m.mu_xx);
  std::memcpy(mu_xy, m.mu_xy, sizeof m.mu_xy);
  std::memcpy(w, m.w, sizeof m.w);
  *count = m.count;
}
double slo_reduce_tree256(const double* v, uint32_t n, uint32_t stride) { return reduce_tree256(v, n, stride); }

/* operator<<(ostream&, State), db.cpp:18-25: default ostream formatting = "%g" with 6 significant digits */
size_t slo_csv_row(const double* row, uint32_t d, char* out, size_t cap) {
  std::string s;
  char buf[64];
  for (u32 i = 0; i < d; ++i) {
    std::snprintf(buf, sizeof buf, "%g,", row[i]);
    s += buf;
  }
  std::snprintf(buf, sizeof buf, "%g,%g,%g,%d,%d", row[d], row[d + 1], row[d + 2], (int)row[d + 3], (int)row[d + 4]);
  s += buf;
  if (out && cap) {
    const size_t n = std::min(cap - 1, s.size());
    std::memcpy(out, s.data(), n);
    out[n] = 0;
  }
  return s.size();
}

double slo_target_eval(int32_t target, const double* params, uint32_t d, uint32_t J, uint32_t job, const double* x) {
  return target_eval(target, params, d, J, job, x);
}

/* ------------------------------------------------------------------------------------------------
 * CPU baseline: stateline server + N workers.  The sampler thread runs the
 * reference state machine (advance()); every proposal becomes J jobs in a FIFO (delegator.cpp:147-171),
 * worker threads pull jobs, parse the "%f"-formatted, ':'-joined sample (requester.cpp:27-41,
 * minion.cpp:35-49), evaluate one job type (workerwrapper.cpp:30-34) and send the "%f" result back
 * (minion.cpp:51-54).  No ZeroMQ hops, heartbeats or per-step JSON: optimistic for the reference.
 * ---------------------------------------------------------------------------------------------- */
namespace {
struct Job { u32 chain; u32 job; std::string text; };
struct Done { u32 chain; u32 job; std::string text; };
struct Farm {
  std::mutex mj, md;
  std::condition_variable cj, cd;
  std::deque<Job> jobs;
  std::deque<Done> done;
  bool stop = false;
};
}  // namespace

double slo_time_server_workers(const slo_config* cfg, uint32_t n_workers, uint64_t n_rounds, uint64_t* chain_steps) {
  slo_config c = *cfg;
  c.rng_mode = SLO_RNG_PHILOX;
  if (c.schedule == SLO_SCHED_BATCHED) c.schedule = SLO_SCHED_SEQUENTIAL;
  c.wire_quantize = 0;
  slo_sampler* o = slo_create(&c);
  slo_init_chains(o);
  const u64 total = n_rounds * o->C;
  double secs = 0.0;
  if (n_workers == 0) {
    const auto t0 = std::chrono::steady_clock::now();
    for (u64 i = 0; i < total; ++i) o->step_sequential_inline();
    secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  } else {
    Farm farm;
    const u32 d = o->d, J = o->J;
    const int target = c.target;
    const std::vector<double> prm = o->tparams;
    std::vector<std::thread> workers;
    for (u32 wkr = 0; wkr < n_workers; ++wkr) {
      workers.emplace_back([&farm, d, J, target, &prm]() {
        std::vector<double> x(d);
        for (;;) {
          Job jb;
          {
            std::unique_lock<std::mutex> lk(farm.mj);
            farm.cj.wait(lk, [&] { return farm.stop || !farm.jobs.empty(); });
            if (farm.stop && farm.jobs.empty()) return;
            jb = std::move(farm.jobs.front());
            farm.jobs.pop_front();
          }
          const char* p = jb.text.c_str();
          for (u32 i = 0; i < d; ++i) {
            char* e;
            x[i] = std::strtod(p, &e);
            p = (*e == ':') ? e + 1 : e;
          }
          const double f = target_eval(target, prm.data(), d, J, jb.job, x.data());
          char buf[400];
          std::snprintf(buf, sizeof buf, "%f", f);
          {
            std::lock_guard<std::mutex> lk(farm.md);
            farm.done.push_back({jb.chain, jb.job, buf});
          }
          farm.cd.notify_one();
        }
      });
    }
    std::vector<std::vector<double>> results(o->C, std::vector<double>(J, 0.0));
    std::vector<u32> pending(o->C, 0);
    auto submit_pending = [&]() {
      /* everything propose() pushed on the FIFO becomes J text jobs */
      while (!o->fifo.empty()) {
        const u32 id = o->fifo.front();
        o->fifo.pop_front();
        std::string text;
        char buf[400];
        for (u32 i = 0; i < d; ++i) {
          std::snprintf(buf, sizeof buf, "%f", o->prop[id][i]);
          if (i) text += ':';
          text += buf;
        }
        pending[id] = J;
        {
          std::lock_guard<std::mutex> lk(farm.mj);
          for (u32 j = 0; j < J; ++j) farm.jobs.push_back({id, j, text});
        }
        farm.cj.notify_all();
      }
    };
    const auto t0 = std::chrono::steady_clock::now();
    submit_pending();
    u64 stepsDone = 0;
    while (stepsDone < total) {
      Done dn;
      {
        std::unique_lock<std::mutex> lk(farm.md);
        farm.cd.wait(lk, [&] { return !farm.done.empty(); });
        dn = std::move(farm.done.front());
        farm.done.pop_front();
      }
      results[dn.chain][dn.job] = std::strtod(dn.text.c_str(), nullptr);
      if (--pending[dn.chain] == 0) {
        double e = 0.0;
        for (u32 j = 0; j < J; ++j) e += results[dn.chain][j];
        o->advance(dn.chain, e);
        ++stepsDone;
        submit_pending();
      }
    }
    secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    {
      std::lock_guard<std::mutex> lk(farm.mj);
      farm.stop = true;
      farm.jobs.clear();
    }
    farm.cj.notify_all();
    for (auto& t : workers) t.join();
  }
  if (chain_steps) *chain_steps = total;
  slo_destroy(o);
  return secs;
}

}  // extern "C"
