/*
 * slgpu_device.cuh -- the parallel-tempering step kernels (sm_100a), templated on the user's
 * likelihood functor.  Included by plugin translation units only (see slgpu_plugin.h).
 *
 * What it replaces in NICTA/stateline: the body of Sampler::step() (src/infer/sampler.cpp:142-203)
 * and everything it calls -- GaussianProposal::propose + bouncyBounds (sampler.cpp:23-93),
 * the likelihood round trip (src/comms/requester.cpp:27-57 -> user LikelihoodFn, workerwrapper.hpp:21),
 * ChainArray::append/acceptProposal/swap/acceptSwap (src/infer/chainarray.cpp:30-67,94-189),
 * the per-row half of RegressionAdapter::update/predict/computeSigma (src/infer/adaptive.cpp:63-139),
 * ProposalShaper::update (adaptive.cpp:173-209), the TableLogger counters (src/infer/logging.cpp:35-48),
 * EPSRDiagnostic::update (src/infer/diagnostics.cpp:27-47) and the cold-chain row sink
 * (chainarray.cpp:131-155, src/db/db.cpp:18-47).
 *
 * Likelihood contract (device form of `double f(uint jobType, const std::vector<double>& x)`):
 *
 *     struct MyLik {
 *       __device__ explicit MyLik(const void* payload);   // payload: POD copied to the device by the engine
 *       __device__ double operator()(uint32_t job_type, const double* x, uint32_t d) const;  // negative log-likelihood
 *     };
 *
 * pure, re-entrant, no allocation; +inf means "certain reject" (chainarray.cpp:36-37).  The energy of
 * a sample is the sum over job types in job-type order (sampler.cpp:148-150).
 *
 * Mapping: one thread per chain, the T chains of a stack in T adjacent lanes of one warp (floor(32/T)
 * stacks per warp), so the hot->cold swap sweep of a stack runs on warp shuffles.  All arithmetic is
 * fp64.  Compile with -fmad=false: products and sums round separately except where fma() is written,
 * which is exactly the operation order of the CPU oracle, so + - * / sqrt fma agree bit for bit.
 */
#ifndef SLGPU_DEVICE_CUH
#define SLGPU_DEVICE_CUH

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>
#ifdef SLGPU_TIMING
#include <cstdio>
#endif

#include "slgpu_math.h"
#include "slgpu_plugin.h"

#ifndef SLGPU_THREADS
#define SLGPU_THREADS 128
#endif

#ifndef SLGPU_PHILOX_INLINE
#define SLGPU_PHILOX_INLINE __forceinline__
#endif

namespace slgpu {

typedef unsigned long long u64;
typedef uint32_t u32;

/* ---- Philox4x32-10, counter layout of the oracle (oracle/sl_oracle.cpp philox_draw) ---- */
__device__ SLGPU_PHILOX_INLINE void philox4x32_10(u32 c0, u32 c1, u32 c2, u32 c3, u32 k0, u32 k1, u32 out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const u32 h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    const u32 h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    const u32 n0 = h1 ^ c1 ^ k0;
    const u32 n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* two uniforms strictly inside (0,1): u = (k52 + 0.5) * 2^-52 */
__device__ __forceinline__ void philox_draw(u64 seed, u32 chain, u64 round, u32 purpose, u32 block, double& ua, double& ub) {
  u32 o[4];
  philox4x32_10((u32)round, chain, (purpose & 0xffu) | ((u32)(round >> 32) << 8), block, (u32)seed, (u32)(seed >> 32), o);
  /* (k52 + 0.5) * 2^-52 with k52 the top 52 bits, built exactly without an integer->double conversion:
   * [1,2) double from the mantissa bits, minus 1 (exact), plus 2^-53 (exact: 2 k52 + 1 < 2^53) */
  const u64 a = (((u64)o[1] << 32) | o[0]) >> 12;
  const u64 b = (((u64)o[3] << 32) | o[2]) >> 12;
  ua = (__longlong_as_double((long long)(a | 0x3ff0000000000000ull)) - 1.0) + 1.1102230246251565404236316680908203125e-16;
  ub = (__longlong_as_double((long long)(b | 0x3ff0000000000000ull)) - 1.0) + 1.1102230246251565404236316680908203125e-16;
}

/* std::max / std::min comparison order (NaN handling as on the host) */
__device__ __forceinline__ double smax(double a, double b) { return (a < b) ? b : a; }
__device__ __forceinline__ double smin(double a, double b) { return (b < a) ? b : a; }

/* bouncyBounds, one coordinate -- src/infer/sampler.cpp:23-56 */
/* n / c given rc = RN(1/c): q0 = RN(n rc), r = n - q0 c (exact, fma), q = RN(q0 + r rc).  By Markstein's
 * theorem this is the correctly rounded quotient -- the same bits as the IEEE division the reference
 * and the oracle perform -- at 3 fp64 instructions instead of ~12+ (tests/test_math.py checks 10^6 cases). */
__device__ __forceinline__ double div_by(double n, double c, double rc) {
  const double q0 = n * rc;
  const double r = fma(-q0, c, n);
  return fma(r, rc, q0);
}

/* bouncyBounds, one coordinate -- src/infer/sampler.cpp:23-56.  rdelta = RN(1/(mx-mn)).  Only one of the two
 * reference branches can fire (the second tests the original value), so they share one reflection; hot chains
 * (sigma at its cap) reflect on almost every coordinate, so the reflection is inline and division-free. */
__device__ __forceinline__ double bounce1(double v, double mn, double mx, double rdelta) {
  const bool over = v > mx, under = v < mn;
  if (!(over || under)) return v;
  const double delta = mx - mn;
  const double dist = over ? v - mx : mn - v;      /* overstep / understep */
  const int nSteps = (int)div_by(dist, delta, rdelta);
  const double stillToGo = dist - nSteps * delta;
  const bool towards_min = over == ((nSteps & 1) == 0); /* over & even, or under & odd: measured from max downwards */
  return towards_min ? mx - stillToGo : mn + stillToGo;
}

/* RegressionAdapter::predict -- src/infer/adaptive.cpp:94-106 */
__device__ __forceinline__ double predict(const double* w, double opt, double min_cap, double max_cap, double side) {
  const double eps = 1e-3;
  const double denom = smax(eps, w[0]);
  double numer = -(opt - w[1] * side - w[2]);
  numer = smax(smin(numer, denom * max_cap), denom * min_cap);
  return numer / denom;
}

/* one sample of the regression sufficient statistics (model half of adaptive.cpp:63-75, summed form) */
template <int STRIDE>
__device__ __forceinline__ void stats_accumulate(double* P, double min_cap, double max_cap, double logval, double side, bool acc) {
  logval = smin(smax(logval, min_cap), max_cap);
  const double x0 = -logval, x1 = side, y = acc ? 1.0 : 0.0;
  P[0 * STRIDE] += x0 * x0;
  P[1 * STRIDE] += x0 * x1;
  P[2 * STRIDE] += x0;
  P[3 * STRIDE] += x1 * x1;
  P[4 * STRIDE] += x1;
  P[5 * STRIDE] += 1.0;
  P[6 * STRIDE] += x0 * y;
  P[7 * STRIDE] += x1 * y;
  P[8 * STRIDE] += y;
}

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

/* ProposalShaper::update -- src/infer/adaptive.cpp:173-209 on the packed lower triangle L(r,k) =
 * Lc[(r(r+1)/2+k)*ls]; xs holds the accepted jump on entry.  Returns the new N scale. */
template <int D>
__device__ __forceinline__ double shaper_update(double* __restrict__ Lc, size_t ls, int d_rt, double* xs, double& count, double prop_norm) {
  constexpr int DM = D > 0 ? D : (int)SLGPU_MAX_DIMS;
  const int d = D > 0 ? D : d_rt;
  const double eps = 1e-18;
  count += 1.0;
  const double sq = sqrt(count);
  const double scale = sqrt((count - 1.0) / count);
  double fr[DM]; /* per-row sums of squares, left to right */
#pragma unroll
  for (int i = 0; i < d; ++i) {
    xs[i] = xs[i] / sq;
    fr[i] = 0.0;
  }
#pragma unroll
  for (int k = 0; k < d; ++k) {
    const size_t kk = (size_t)(k * (k + 1) / 2 + k) * ls;
    const double Lkk = Lc[kk] * scale;
    const double V = smax(eps, Lkk);
    const double r = sqrt(V * V + xs[k] * xs[k]);
    const double c = r / V;
    const double s = xs[k] / V;
    Lc[kk] = r;
    fr[k] = fma(r, r, fr[k]);
#pragma unroll
    for (int j = k + 1; j < d; ++j) {
      const size_t jk = (size_t)(j * (j + 1) / 2 + k) * ls;
      double Ljk = Lc[jk] * scale;
      Ljk = (Ljk + s * xs[j]) / c;
      Lc[jk] = Ljk;
      xs[j] = c * xs[j] - s * Ljk;
      fr[j] = fma(Ljk, Ljk, fr[j]);
    }
  }
  /* |L|_F in the oracle's canonical order: slot l sums rows l, l+32, ...; halving tree over 32 slots */
  double slot[32];
#pragma unroll
  for (int l = 0; l < 32; ++l) slot[l] = 0.0;
#pragma unroll
  for (int r = 0; r < d; ++r) slot[r & 31] += fr[r];
  int nlive = d < 32 ? d : 32;
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
#pragma unroll
    for (int l = 0; l < m; ++l)
      if (l + m < nlive) slot[l] += slot[l + m]; /* slots >= nlive are +0: adding them is the identity */
    if (nlive > m) nlive = m;
  }
  const double frob = sqrt(slot[0]);
  return prop_norm / smax(frob, eps);
}

/* Optional compile-time job count: a functor may declare
 *     template <int D> static constexpr int static_job_types();
 * so that the sum over job types unrolls (static indexing keeps x in registers).  0 = run-time count. */
template <class L, int D, class = void>
struct StaticJobs { static constexpr int value = 0; };
template <class L, int D>
struct StaticJobs<L, D, std::void_t<decltype(L::template static_job_types<D>())>> {
  static constexpr int value = L::template static_job_types<D>();
};

/* Optional: a functor may evaluate all job types of one point with a whole warp,
 *     __device__ void warp_jobs(const double* x, u32 d, u32 n_jobs, int lane, double* values, double* scratch) const;
 * It is called by the 32 lanes of a warp together; x, values ([n_jobs]) and scratch ([d]) are per-warp shared memory.
 * When it returns, values[j] must hold exactly what operator()(j, x, d) returns (same operations in the same order)
 * and be visible to every lane (end with __syncwarp()).  The warp-per-chain kernel (d > 32) uses it: a likelihood on
 * one lane per job type leaves the other lanes idle. */
template <class L, class = void>
struct HasWarpJobs : std::false_type {};
template <class L>
struct HasWarpJobs<L, std::void_t<decltype(&L::warp_jobs)>> : std::true_type {};

/* base of chain (stack s, temperature t)'s packed L inside its stack group's tile (slgpu_plugin.h); consecutive
 * entries are SLGPU_L_TILE doubles apart */
__device__ __forceinline__ double* l_base(double* L, u32 n_temps, u32 s, u32 t, u32 nL) {
  const u32 spw = 32u / n_temps;
  return L + (size_t)(s / spw) * nL * SLGPU_L_TILE + (s % spw) * n_temps + t;
}

/* entry (r,k), k <= r, of chain (s,t)'s factor in whichever layout the dimension uses */
__device__ __forceinline__ double& l_elem(double* L, u32 n_temps, u32 s, u32 t, int d, int r, int k) {
  const size_t nL = (size_t)d * (d + 1) / 2;
  if (SLGPU_L_LAYOUT((u32)d) == SLGPU_L_COLMAJOR)
    return L[(size_t)(s * n_temps + t) * nL + (size_t)k * d - (size_t)k * (k - 1) / 2 + (size_t)(r - k)];
  return l_base(L, n_temps, s, t, (u32)nL)[(size_t)(r * (r + 1) / 2 + k) * SLGPU_L_TILE];
}

/* regression statistics: P[(k*T + t)*S + s], k < 9 -- coalesced over the stacks of a tier for the reduce */
__device__ __forceinline__ size_t pidx(const slgpu_step_params& p, int k, int t, u32 s) {
  return ((size_t)k * p.n_temps + (size_t)t) * p.n_stacks + s;
}

struct ThreadMap {
  int T, lane, sl, t, base, spw;
  u32 s, c;
  bool active;
  __device__ ThreadMap(const slgpu_step_params& p, u32 block = blockIdx.x) {
    T = (int)p.n_temps;
    spw = 32 / T;
    lane = threadIdx.x & 31;
    sl = lane / T;
    t = lane - sl * T;
    const u32 gw = block * (blockDim.x >> 5) + (threadIdx.x >> 5);
    s = gw * (u32)spw + (u32)sl;
    active = (sl < spw) && (s < p.n_stacks);
    if (sl >= spw) { /* idle tail lanes of a warp: keep every shuffle source inside the warp */
      base = lane;
      t = 0;
    } else {
      base = sl * T;
    }
    c = s * (u32)T + (u32)t;
  }
};

/* --------------------------------------------------------------------------------------------
 * pt_init: chain initialisation, src/app/serverwrapper.cpp:60-81 + generateInitialSample :21-41,
 * ProposalShaper ctor adaptive.cpp:155-171, ChainArray::initialise chainarray.cpp:123-129,
 * TableLogger ctor logging.cpp:29-33.
 * ------------------------------------------------------------------------------------------ */
template <class Lik, int D>
__global__ void __launch_bounds__(SLGPU_THREADS) pt_init_kernel(const slgpu_step_params p) {
  constexpr int DM = D > 0 ? D : (int)SLGPU_MAX_DIMS;
  const int d = D > 0 ? D : (int)p.n_dims;
  const ThreadMap m(p);
  if (!m.active) return;
  const size_t C = (size_t)p.n_stacks * p.n_temps;
  const u32 c = m.c;
  const u32 gchain = p.stack_offset * p.n_temps + c;
  const Lik lik(p.payload);
  double x[DM];
  if (p.initial) {
#pragma unroll
    for (int i = 0; i < d; ++i)
      x[i] = p.initial[i];
  } else if (p.rng_mode == SLGPU_RNG_REPLAY) {
#pragma unroll
    for (int i = 0; i < d; ++i)
      x[i] = p.rp_xinit[(size_t)c * d + i];
  } else { /* VectorXd::Random(d): uniform in [-1,1] */
#pragma unroll
    for (int k = 0; k < (d + 1) / 2; ++k) {
      double ua, ub;
      philox_draw(p.seed, gchain, 0, 3, (u32)k, ua, ub);
      x[2 * k] = 2.0 * ua - 1.0;
      if (2 * k + 1 < d) x[2 * k + 1] = 2.0 * ub - 1.0;
    }
  }
#pragma unroll
  for (int i = 0; i < d; ++i)
    x[i] = bounce1(x[i], p.mn[i], p.mx[i], 1.0 / (p.mx[i] - p.mn[i]));
  double e = 0.0;
  for (u32 j = 0; j < p.n_job_types; ++j) e += lik(j, x, (u32)d);
#pragma unroll
  for (int i = 0; i < d; ++i)
    p.x[(size_t)i * C + c] = x[i];
  for (int r = 0; r < d; ++r)
    for (int k = 0; k <= r; ++k)
      l_elem(p.L, p.n_temps, m.s, (u32)m.t, d, r, k) = (r == k) ? (p.mx[r] - p.mn[r]) / 4.0 : 0.0;
  if (p.moment) { /* "gpu": {"shaper": "welford"}: the second moment behind the initial factor, L0 L0^T = diag(range^2) */
    for (int r = 0; r < d; ++r)
      for (int k = 0; k <= r; ++k) {
        const double rg = (p.mx[r] - p.mn[r]) / 4.0;
        l_elem(p.moment, p.n_temps, m.s, (u32)m.t, d, r, k) = (r == k) ? rg * rg : 0.0;
      }
    p.chol_dirty[c] = 0;
  }
  const double b = p.ladder[m.t];
  p.energy[c] = e;
  p.row_sigma[c] = 1.0;
  p.row_beta[c] = b;
  p.sigma[c] = 1.0;
  p.beta[c] = b;
  p.nscale[c] = 0.0; /* unused until the first shaper update */
  p.sh_count[c] = p.sh_count0;
  p.accepted[c] = 1;
  p.swap_type[c] = 0;
  p.lg_accepts[c] = 1;
  p.lg_swaps[c] = 0;
  p.lg_swap_attempts[c] = 1;
  p.lg_min_energy[c] = __longlong_as_double(0x7ff0000000000000LL);
  p.lg_cur_energy[c] = 0.0; /* energies_ is a zero-filled vector until the first update (logging.cpp:23) */
  for (int k = 0; k < 9; ++k) {
    p.p_sig[pidx(p, k, m.t, m.s)] = 0.0;
    p.p_beta[pidx(p, k, m.t, m.s)] = 0.0;
  }
  if (m.t == 0) {
    for (int i = 0; i < d; ++i) {
      p.ep_m[(size_t)i * p.n_stacks + m.s] = 0.0;
      p.ep_s[(size_t)i * p.n_stacks + m.s] = 0.0;
    }
    if (p.rows_out) { /* the initial state is row 0 of the stack's output (chainarray.cpp:123-129) */
      double* row = p.rows_out + (size_t)m.s * (d + SLGPU_ROW_EXTRA);
      for (int i = 0; i < d; ++i) row[i] = x[i];
      row[d] = e; row[d + 1] = 1.0; row[d + 2] = b; row[d + 3] = 1.0; row[d + 4] = 0.0;
    }
  }
}

/* --------------------------------------------------------------------------------------------
 * pt_step: n_rounds canonical rounds.  Round r (0-based; chain length becomes r+2):
 *   propose -> energy -> MH accept -> sigma statistics + new sigma -> shaper update,
 *   then, when (r+2) % swapInterval == 0, the hot->cold swap sweep of every stack, the beta
 *   statistics and the ladder refresh, then counters / EPSR / cold-row emission.
 * The tier models (sig_w, ladder) are frozen for the launch ("batched" schedule of the oracle).
 * ------------------------------------------------------------------------------------------ */
template <class Lik, int D>
__global__ void __launch_bounds__(SLGPU_THREADS) pt_step_kernel(const slgpu_step_params p) {
  constexpr int DM = D > 0 ? D : (int)SLGPU_MAX_DIMS;
  const int d = D > 0 ? D : (int)p.n_dims;
  const ThreadMap m(p);
  const int T = m.T;
  const size_t C = (size_t)p.n_stacks * p.n_temps;
  const u32 c = m.active ? m.c : 0u; /* idle lanes shadow chain 0 for loads and never store */
  const u32 gchain = p.stack_offset * p.n_temps + c;
  const bool active = m.active;

  __shared__ double s_mn[DM], s_mx[DM], s_n0[DM], s_rd[DM];
  __shared__ double s_w[SLGPU_MAX_TEMPS * 3], s_lad[SLGPU_MAX_TEMPS];
  __shared__ double s_P[18 * SLGPU_THREADS];
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    s_mn[i] = p.mn[i];
    s_mx[i] = p.mx[i];
    s_n0[i] = p.n0_diag[i];
    s_rd[i] = 1.0 / (p.mx[i] - p.mn[i]);
  }
  for (int i = threadIdx.x; i < T; i += blockDim.x) {
    s_w[3 * i + 0] = p.sig_w[3 * i + 0];
    s_w[3 * i + 1] = p.sig_w[3 * i + 1];
    s_w[3 * i + 2] = p.sig_w[3 * i + 2];
    s_lad[i] = p.ladder[i];
  }
  double* P = s_P + threadIdx.x; /* P[k*SLGPU_THREADS], k<9 sigma statistics, 9..17 beta statistics */
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    P[k * SLGPU_THREADS] = active ? p.p_sig[pidx(p, k, m.t, m.s)] : 0.0;
    P[(9 + k) * SLGPU_THREADS] = active ? p.p_beta[pidx(p, k, m.t, m.s)] : 0.0;
  }
  __syncthreads();

  const Lik lik(p.payload);
  double* __restrict__ Lc = l_base(p.L, p.n_temps, m.active ? m.s : 0u, (u32)m.t, (u32)(d * (d + 1) / 2)); /* idle lanes shadow chain 0 */
  const size_t ls = SLGPU_L_TILE;

  double x[DM];
#pragma unroll
  for (int i = 0; i < d; ++i)
    x[i] = p.x[(size_t)i * C + c];
  double energy = p.energy[c], row_sigma = p.row_sigma[c], row_beta = p.row_beta[c];
  double sigma = p.sigma[c], beta = p.beta[c], nscale = p.nscale[c], sh_count = p.sh_count[c];
  int accepted = p.accepted[c], swap_type = p.swap_type[c];
  u32 lg_accepts = p.lg_accepts[c], lg_swaps = p.lg_swaps[c], lg_att = p.lg_swap_attempts[c];
  double lg_min_e = p.lg_min_energy[c];
  const double w3[3] = {s_w[3 * m.t], s_w[3 * m.t + 1], s_w[3 * m.t + 2]};
  const int W = (int)p.swap_interval;

  for (u32 it = 0; it < p.n_rounds; ++it) {
    bool mh_acc = false; /* this chain's own Metropolis-Hastings decision of the round */
    const u64 r = p.round_begin + it;
    if (active) {
      /* ---- proposal: x' = bounce(x + N z sigma), GaussianProposal::propose sampler.cpp:79-93 ---- */
      double z[DM];
      if (p.rng_mode == SLGPU_RNG_REPLAY) {
        const double* zr = p.rp_z + ((size_t)r * C + c) * d;
#pragma unroll
        for (int i = 0; i < d; ++i)
          z[i] = zr[i];
      } else {
#pragma unroll
        for (int k = 0; k < (d + 1) / 2; ++k) {
          double ua, ub;
          philox_draw(p.seed, gchain, r, 0, (u32)k, ua, ub);
          const double R = sqrt(-2.0 * slm_log(ua));
          double sn, cs;
          slm_sincos2pi(ub, &sn, &cs);
          z[2 * k] = R * cs;
          if (2 * k + 1 < d) z[2 * k + 1] = R * sn;
        }
      }
      const bool fresh = (sh_count == p.sh_count0); /* N still the constructor's diag(range/|range|) */
      double xp[DM];
#pragma unroll
      for (int i = 0; i < d; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < i; ++k) acc = fma(Lc[(size_t)(i * (i + 1) / 2 + k) * ls] * nscale, z[k], acc);
        const double nii = fresh ? s_n0[i] : Lc[(size_t)(i * (i + 1) / 2 + i) * ls] * nscale + 1e-4;
        acc = fma(nii, z[i], acc);
        xp[i] = bounce1(x[i] + acc * sigma, s_mn[i], s_mx[i], s_rd[i]);
      }
      /* ---- energy: sum over job types in order, sampler.cpp:148-150 ---- */
      double e_new = 0.0;
      for (u32 j = 0; j < p.n_job_types; ++j) e_new += lik(j, xp, (u32)d);
      /* ---- acceptProposal chainarray.cpp:30-45, append :94-121 ---- */
      bool acc = false;
      if (!isinf(e_new)) {
        double u;
        if (p.rng_mode == SLGPU_RNG_REPLAY) u = p.rp_umh[(size_t)r * C + c];
        else { double ub; philox_draw(p.seed, gchain, r, 1, 0, u, ub); }
        const double deltaEnergy = e_new - energy;
        const double pr = slm_exp(-1.0 * beta * deltaEnergy);
        acc = u < pr;
      }
      if (acc) {
#pragma unroll
        for (int i = 0; i < d; ++i) {
          const double nx = xp[i];
          xp[i] = nx - x[i]; /* xp now holds the accepted jump (sampler.cpp:166) */
          x[i] = nx;
        }
        energy = e_new;
        row_sigma = sigma;
        row_beta = beta;
      }
      accepted = acc ? 1 : 0;
      mh_acc = acc;
      swap_type = 0;
      /* ---- sigma adaptation, sampler.cpp:160-162 (statistics summed, model frozen) ---- */
      const double logtemper = -slm_log(row_beta);
      stats_accumulate<SLGPU_THREADS>(P, -8.0, 4.0, slm_log(row_sigma), logtemper, acc);
      sigma = slm_exp(predict(w3, p.opt_accept, -8.0, 4.0, logtemper));
      /* ---- proposal shape, sampler.cpp:165-166 ---- */
      if (acc) nscale = shaper_update<D>(Lc, ls, d, xp, sh_count, p.prop_norm);
    }

    /* ---- swap sweep, hottest pair first: sampler.cpp:169-199,237-257, chainarray.cpp:55-67,162-189 ---- */
    double mid_e = energy; /* the state step() returns for this chain: after its swap with the chain */
    int mid_acc = accepted; /* above, before the swap with the chain below */
    const bool swapRound = (T > 1) && ((r + 2) % (u64)W == 0);
    if (swapRound) {
      double u_sw = 0.0;
      if (active && m.t < T - 1) {
        if (p.rng_mode == SLGPU_RNG_REPLAY) u_sw = p.rp_uswap[(size_t)r * C + c];
        else { double ub; philox_draw(p.seed, gchain, r, 2, 0, u_sw, ub); }
      }
      int cur = T - 1, my_src = m.t, mid_src = m.t;
      bool my_swapped = false;
      for (int tt = T - 2; tt >= 0; --tt) {
        const double El = shfl_d(energy, m.base + tt);
        const double Eh = shfl_d(energy, m.base + cur);
        const double bl = shfl_d(beta, m.base + tt);
        const double bh = shfl_d(beta, m.base + tt + 1);
        const double u = shfl_d(u_sw, m.base + tt);
        const double deltaEnergy = El - Eh;
        const double deltaBeta = bl - bh;
        const bool sw = u < slm_exp(deltaEnergy * deltaBeta);
        if (m.t == tt + 1) my_src = sw ? tt : cur;
        if (m.t == tt) { my_swapped = sw; mid_src = sw ? cur : tt; }
        if (!sw) cur = tt;
      }
      if (m.t == 0) my_src = cur;
      /* beta statistics of pair (t, t+1): betaUpdate adaptive.cpp:108-116 */
      const double lb = slm_log(beta);
      const double lbh = shfl_d(lb, m.lane + 1 < 32 ? m.lane + 1 : m.lane);
      mid_e = shfl_d(energy, m.base + mid_src);
      mid_acc = __shfl_sync(0xffffffffu, accepted, m.base + mid_src);
      /* the exchange moves sample, energy and accepted; sigma and beta stay (chainarray.cpp:176-178) */
      const double new_e = shfl_d(energy, m.base + my_src);
      const int new_acc = __shfl_sync(0xffffffffu, accepted, m.base + my_src);
#pragma unroll
      for (int i = 0; i < d; ++i)
        x[i] = shfl_d(x[i], m.base + my_src);
      energy = new_e;
      accepted = new_acc;
      if (m.t < T - 1) {
        swap_type = my_swapped ? 1 : 2;
        stats_accumulate<SLGPU_THREADS>(P + 9 * SLGPU_THREADS, 0.0, 4.0, lb - lbh, lb, my_swapped);
      }
      if (m.t > 0) beta = s_lad[m.t]; /* sampler.cpp:186 with the ladder of the last adapt point */
    }

    if (active) {
      /* ---- TableLogger::update, logging.cpp:35-48 ---- */
      lg_min_e = smin(lg_min_e, mid_e);
      lg_accepts += (u32)mid_acc;
      lg_swaps += (swap_type == 1);
      lg_att += (swap_type != 0);
      if (p.round_flags) p.round_flags[(size_t)it * C + c] = (unsigned char)((mh_acc ? 1 : 0) | (swap_type << 1));
      if (it + 1 == p.n_rounds) p.lg_cur_energy[c] = mid_e; /* TableLogger energies_ (logging.cpp:43) */
      if (m.t == 0) {
        /* ---- EPSRDiagnostic::update diagnostics.cpp:27-47; n_ counts the cold rows seen ---- */
        const double nn = (double)(r + 1);
        for (int i = 0; i < d; ++i) {
          const size_t o = (size_t)i * p.n_stacks + m.s;
          const double mo = p.ep_m[o];
          const double xi = x[i];
          const double newM = mo + (xi - mo) / nn;
          p.ep_s[o] = p.ep_s[o] + (xi - mo) * (xi - newM);
          p.ep_m[o] = newM;
        }
        /* ---- cold-chain row, db.cpp:18-25 layout ---- */
        if (p.rows_out) {
          double* row = p.rows_out + ((size_t)it * p.n_stacks + m.s) * (d + SLGPU_ROW_EXTRA);
#pragma unroll
          for (int i = 0; i < d; ++i)
            row[i] = x[i];
          row[d] = energy;
          row[d + 1] = row_sigma;
          row[d + 2] = row_beta;
          row[d + 3] = (double)accepted;
          row[d + 4] = (double)swap_type;
        }
      }
    }
  }

  if (active) {
#pragma unroll
    for (int i = 0; i < d; ++i)
      p.x[(size_t)i * C + c] = x[i];
    p.energy[c] = energy;
    p.row_sigma[c] = row_sigma;
    p.row_beta[c] = row_beta;
    p.sigma[c] = sigma;
    p.beta[c] = beta;
    p.nscale[c] = nscale;
    p.sh_count[c] = sh_count;
    p.accepted[c] = accepted;
    p.swap_type[c] = swap_type;
    p.lg_accepts[c] = lg_accepts;
    p.lg_swaps[c] = lg_swaps;
    p.lg_swap_attempts[c] = lg_att;
    p.lg_min_energy[c] = lg_min_e;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      p.p_sig[pidx(p, k, m.t, m.s)] = P[k * SLGPU_THREADS];
      p.p_beta[pidx(p, k, m.t, m.s)] = P[(9 + k) * SLGPU_THREADS];
    }
  }
}

}  // namespace slgpu

/* d <= 32, compile-time dimension: the thread-per-chain kernel (slgpu_step_fast.cuh) is the default; building a plugin
 * with -DSLGPU_TILE_KERNEL selects the block-per-stack-group kernel with the factors resident in shared memory
 * (slgpu_step_tile.cuh): same bits, HBM traffic = algorithmic, 1.9x the time on config #3 (profiles/README.md) */
#ifdef SLGPU_TILE_KERNEL
#include "slgpu_step_tile.cuh"
#else
#include "slgpu_step_fast.cuh"
#endif
#include "slgpu_step_warp.cuh"

namespace slgpu {

/* contract check: nll[point*J + job] = f(job, x[point]) */
template <class Lik>
__global__ void lik_eval_kernel(const slgpu_step_params p, const double* __restrict__ xs, u32 n_points, double* __restrict__ nll) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_points) return;
  const Lik lik(p.payload);
  double x[SLGPU_MAX_DIMS];
  for (u32 k = 0; k < p.n_dims; ++k) x[k] = xs[(size_t)i * p.n_dims + k];
  for (u32 j = 0; j < p.n_job_types; ++j) nll[(size_t)i * p.n_job_types + j] = lik(j, x, p.n_dims);
}

inline unsigned grid_for(const slgpu_step_params* p, unsigned threads = SLGPU_THREADS) {
  const unsigned spw = 32u / p->n_temps;
  const unsigned per_block = spw * (threads / 32);
  return (p->n_stacks + per_block - 1) / per_block;
}

/* host launchers instantiated per (functor, compile-time dimension); D = 0 is the runtime-d kernel */
template <class Lik, int D>
int launch_init_t(const slgpu_step_params* p, void* stream) {
  pt_init_kernel<Lik, D><<<grid_for(p), SLGPU_THREADS, 0, (cudaStream_t)stream>>>(*p);
  return (int)cudaGetLastError();
}
template <class Lik, int D>
int launch_step_t(const slgpu_step_params* p, void* stream) {
  if (SLGPU_L_LAYOUT(p->n_dims) == SLGPU_L_COLMAJOR) { /* d > 32: one warp per chain, whole stacks per block */
    const unsigned spb = warp_stacks_per_block(p->n_temps), warps = spb * p->n_temps;
    const size_t smem = warp_smem_bytes(p, warps, HasWarpJobs<Lik>::value);
    const unsigned grid = (p->n_stacks + spb - 1) / spb, threads = warps * 32;
    const int rm = (int)(p->n_dims + 31) / 32;
#define SLGPU_WARP_LAUNCH(RMV, MAXTV)                                                                                     \
  {                                                                                                                       \
    auto kern = pt_step_warp_kernel<Lik, RMV, MAXTV>;                                                                     \
    if (smem > 48 * 1024) { /* per device and context: set on every launch (a process may drive several devices) */      \
      const cudaError_t attr = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);        \
      if (attr != cudaSuccess) return (int)attr;                                                                          \
    }                                                                                                                     \
    kern<<<grid, threads, smem, (cudaStream_t)stream>>>(*p);                                                              \
  }
    if (threads <= 256) {
      if (rm <= 2) SLGPU_WARP_LAUNCH(2, 256) else if (rm == 3) SLGPU_WARP_LAUNCH(3, 256) else SLGPU_WARP_LAUNCH(4, 256)
    } else if (threads <= 512) {
      if (rm <= 2) SLGPU_WARP_LAUNCH(2, 512) else if (rm == 3) SLGPU_WARP_LAUNCH(3, 512) else SLGPU_WARP_LAUNCH(4, 512)
    } else {
      if (rm <= 2) SLGPU_WARP_LAUNCH(2, 1024) else if (rm == 3) SLGPU_WARP_LAUNCH(3, 1024) else SLGPU_WARP_LAUNCH(4, 1024)
    }
#undef SLGPU_WARP_LAUNCH
    return (int)cudaGetLastError();
  }
  if constexpr (D > 0 && D <= 32) {
#ifdef SLGPU_TILE_KERNEL
    return launch_step_tile<Lik, D>(p, stream);
#else
    constexpr int JS = StaticJobs<Lik, D>::value;
    if (!(JS > 0 && p->n_job_types != (uint32_t)JS)) {
      constexpr int NT = SLGPU_FAST_THREADS;
      auto kern = p->moment ? pt_step_fast_kernel<Lik, D, NT, true> : pt_step_fast_kernel<Lik, D, NT, false>;
      const size_t smem = fast_smem_bytes<D, NT>(p->n_temps);
      /* the opt-in is per device and context: set it on every launch (a process may drive several devices) */
      const cudaError_t attr = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fast_smem_bytes<D, NT>(1));
      if (attr != cudaSuccess) return (int)attr;
      /* programmatic dependent launch: the kernel may start while the k_tier_adapt in front of it still runs and waits
       * for it at cudaGridDependencySynchronize() (slgpu_step_fast.cuh); after any other kernel it starts as usual */
      cudaLaunchConfig_t lc = {};
      lc.gridDim = dim3(grid_for(p, NT));
      lc.blockDim = dim3(NT);
      lc.dynamicSmemBytes = smem;
      lc.stream = (cudaStream_t)stream;
      cudaLaunchAttribute la[1];
      la[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      la[0].val.programmaticStreamSerializationAllowed = 1;
      lc.attrs = la;
      lc.numAttrs = 1;
      return (int)cudaLaunchKernelEx(&lc, kern, *p);
    }
#endif
  }
  pt_step_kernel<Lik, D><<<grid_for(p), SLGPU_THREADS, 0, (cudaStream_t)stream>>>(*p);
  return (int)cudaGetLastError();
}
/* "gpu": {"shaper": "welford"}: re-factorise the dirty chains' moment matrices (rechol_kernel); only the thread-per-chain
 * kernel of a specialised dimension <= 32 keeps moments */
template <int D>
int launch_rechol_t(const slgpu_step_params* p, void* stream) {
#ifndef SLGPU_TILE_KERNEL
  if constexpr (D > 0 && D <= 32) {
    if (p->moment) {
      const size_t smem = ((size_t)D * (D + 1) / 2 + D + kRecholWarps) * 32 * sizeof(double);
      auto kern = rechol_kernel<D>;
      const cudaError_t attr = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (attr != cudaSuccess) return (int)attr;
      const unsigned spw = 32u / p->n_temps;
      kern<<<(p->n_stacks + spw - 1) / spw, kRecholWarps * 32, smem, (cudaStream_t)stream>>>(*p);
      return (int)cudaGetLastError();
    }
  }
#endif
  return (int)cudaErrorNotSupported;
}
template <class Lik>
int launch_eval_t(const slgpu_step_params* p, const double* x, uint32_t n_points, double* nll, void* stream) {
  if (n_points == 0) return 0;
  lik_eval_kernel<Lik><<<(n_points + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*p, x, n_points, nll);
  return (int)cudaGetLastError();
}

/* dispatch on the run-time dimension over a compile-time list of specialised dimensions */
template <class Lik, int... Ds>
struct Dispatch {
  static int init(const slgpu_step_params* p, void* s) {
    int rc = -1;
    const bool hit = ((p->n_dims == (uint32_t)Ds ? (rc = launch_init_t<Lik, Ds>(p, s), true) : false) || ...);
    return hit ? rc : launch_init_t<Lik, 0>(p, s);
  }
  static int step(const slgpu_step_params* p, void* s) {
    int rc = -1;
    const bool hit = ((p->n_dims == (uint32_t)Ds ? (rc = launch_step_t<Lik, Ds>(p, s), true) : false) || ...);
    return hit ? rc : launch_step_t<Lik, 0>(p, s);
  }
  static int rechol(const slgpu_step_params* p, void* s) {
    int rc = (int)cudaErrorNotSupported;
    const bool hit = ((p->n_dims == (uint32_t)Ds ? (rc = launch_rechol_t<Ds>(p, s), true) : false) || ...);
    (void)hit;
    return rc;
  }
};

}  // namespace slgpu

/* Defines a plugin table named `sym` for functor `Lik`; specialised (fully unrolled) kernels are built
 * for the listed dimensions, every other dimension <= SLGPU_MAX_DIMS runs the runtime-d kernel. */
#define SLGPU_DEFINE_PLUGIN(sym, Lik, plugin_name, n_jobs, n_dims_fixed, payload_size, ...)                       \
  static int sym##_init(const slgpu_step_params* p, void* s) { return slgpu::Dispatch<Lik, ##__VA_ARGS__>::init(p, s); } \
  static int sym##_step(const slgpu_step_params* p, void* s) { return slgpu::Dispatch<Lik, ##__VA_ARGS__>::step(p, s); } \
  static int sym##_eval(const slgpu_step_params* p, const double* x, uint32_t n, double* o, void* s) {            \
    return slgpu::launch_eval_t<Lik>(p, x, n, o, s);                                                              \
  }                                                                                                               \
  static int sym##_rechol(const slgpu_step_params* p, void* s) { return slgpu::Dispatch<Lik, ##__VA_ARGS__>::rechol(p, s); } \
  extern "C" const slgpu_plugin_v1* sym(void) {                                                                   \
    static const slgpu_plugin_v1 table = {SLGPU_PLUGIN_ABI, plugin_name, n_jobs, n_dims_fixed, payload_size,      \
                                          sym##_init,       sym##_step,  sym##_eval, sym##_rechol};               \
    return &table;                                                                                                \
  }

#endif
