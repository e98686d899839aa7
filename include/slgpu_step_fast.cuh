/*
 * slgpu_step_fast.cuh -- pt_step for a compile-time dimension D <= 32 (included by slgpu_device.cuh).
 *
 * Same rounds, same arithmetic in the same order as pt_step_kernel and the oracle; what differs is
 * how the work is laid on the SM:
 *   - one thread per chain, x / x' / the row sums live in registers (static indexing only);
 *   - loops over the column index k stay ROLLED (the instruction stream of a fully unrolled d=20
 *     step is ~250 KB and thrashes the instruction caches); the loops over the row index are unrolled
 *     and guarded by warp-uniform branches on k, so the triangular structure costs no wasted lanes;
 *   - the proposal GEMV runs column by column while the normals are generated, so z is never stored;
 *   - the packed L factors stay in global memory, chain-minor ([idx][chain]): every access of a warp is
 *     one coalesced 256-byte line, and a block's factors (128 chains x 1.7 KB) stay L2-resident over
 *     the rounds of a launch, so HBM sees them about once per launch;
 *   - ProposalShaper::update runs only for accepted proposals (~23%).  Done in place it would drag all
 *     32 lanes of a warp through ~7000 instructions for ~7 useful lanes, so the accepted chains of a block
 *     are COMPACTED: phase A stages (jump vector, chain id, count) of each accepted chain in shared
 *     memory, phase B lets threads 0..n_acc-1 each update one staged chain with full lanes, and the new
 *     N scale is handed back to the owning thread.  Two block barriers per round.
 */
#ifndef SLGPU_STEP_FAST_CUH
#define SLGPU_STEP_FAST_CUH

#include "slgpu_bulk.cuh"

namespace slgpu {

#ifndef SLGPU_FAST_MINBLOCKS
#define SLGPU_FAST_MINBLOCKS 2
#endif
#ifndef SLGPU_FAST_THREADS
#define SLGPU_FAST_THREADS 128
#endif
#ifndef SLGPU_FAST_SMEM_PAD
#define SLGPU_FAST_SMEM_PAD 0 /* development knob: extra dynamic shared memory to cap the blocks per SM */
#endif
#ifndef SLGPU_NORMALS_UNROLL
#define SLGPU_NORMALS_UNROLL 1
#endif
#ifndef SLGPU_LDL
#define SLGPU_LDL(p) (*(p)) /* load of an L entry; development knob: __ldcg / __ldcs / __ldca */
#endif
#ifndef SLGPU_COLBLOCK
#define SLGPU_COLBLOCK 2
#endif

constexpr int kLs = (int)SLGPU_L_TILE; /* doubles between consecutive packed entries of a chain's L */
constexpr int kNormalsUnroll = SLGPU_NORMALS_UNROLL;
constexpr int kColBlock = SLGPU_COLBLOCK; /* columns per block of the blocked triangular sweeps */

/* Columns K0..K0+KB-1 of the proposal GEMV: a_j = fma(N_jk, z_k, a_j) for j >= k.  Rows j >= K0 are
 * loaded unconditionally (for K0 <= j < k the address is still inside the packed triangle, the value is
 * ignored), so a column's loads issue back to back; only the first KB rows need the j >= k guard. */
template <int D, int K0, int K1>
__device__ __forceinline__ void gemv_col(int k, const double* col, double zk, double n0k, double nscale, bool fresh, double* a) {
#pragma unroll
  for (int j = K0; j < D; ++j) {
    if (j >= K1 || j >= k) { /* static for j >= K1, warp-uniform otherwise */
      double njk = col[j] * nscale;
      if (j < K1 && j == k) njk = fresh ? n0k : njk + 1e-4;
      a[j] = fma(njk, zk, a[j]);
    }
  }
}

template <int D, int K0, int KB>
__device__ __forceinline__ void gemv_block(const double* Lc, const double* zs, int zstride, double nscale, bool fresh,
                                           const double* s_n0, double* a) {
  constexpr int K1 = (K0 + KB < D) ? K0 + KB : D;
#pragma unroll 1
  for (int k = K0; k < K1; k += 2) { /* two columns per trip: both columns' loads are in flight together */
    const bool two = (k + 1 < K1);
    const double* Lk = Lc + k * kLs;
    double col0[D], col1[D];
#pragma unroll
    for (int j = K0; j < D; ++j) col0[j] = SLGPU_LDL(Lk + (j * (j + 1) / 2) * kLs);
    if (two) {
#pragma unroll
      for (int j = K0; j < D; ++j) col1[j] = SLGPU_LDL(Lk + (j * (j + 1) / 2 + 1) * kLs);
    }
    gemv_col<D, K0, K1>(k, col0, zs[k * zstride], s_n0[k], nscale, fresh, a);
    if (two) gemv_col<D, K0, K1>(k + 1, col1, zs[(k + 1) * zstride], s_n0[k + 1], nscale, fresh, a);
  }
  if constexpr (K1 < D) gemv_block<D, K1, KB>(Lc, zs, zstride, nscale, fresh, s_n0, a);
}

/* Columns K0..K0+KB-1 of ProposalShaper::update's rank-1 sweep (src/infer/adaptive.cpp:184-197).
 * xs (the scaled jump, then the running rank-1 vector) and fr (row sums of squares) are REGISTER arrays: rows are
 * unrolled, and inside a column block the column index k is one of the KB static candidates K0..K1-1, so x_k and
 * row k's sum are picked by warp-uniform selects -- no shared-memory round trip per element (it was two loads and
 * two stores of the 19 instructions per element, on the MIO pipe the kernel is short of). */
template <int D, int K0, int K1>
__device__ __forceinline__ void shaper_col(int k, double* Lk, const double* col, double scale, double* xs, double* fr) {
  const double eps = 1e-18;
  double xk = xs[K0];
  double Lkk = col[K0];
#pragma unroll
  for (int j = K0 + 1; j < K1; ++j) {
    xk = (j == k) ? xs[j] : xk;
    Lkk = (j == k) ? col[j] : Lkk;
  }
  Lkk = Lkk * scale;
  const double V = smax(eps, Lkk);
  const double rV = 1.0 / V;
  const double r = sqrt(V * V + xk * xk);
  const double c = div_by(r, V, rV);
  const double s = div_by(xk, V, rV);
  const double rc = 1.0 / c;
  Lk[(k * (k + 1) / 2) * kLs] = r;
#pragma unroll
  for (int j = K0; j < K1; ++j)
    if (j == k) fr[j] = fma(r, r, fr[j]); /* the diagonal closes row k's sum */
#pragma unroll
  for (int j = K0 + 1; j < D; ++j) {
    if (j >= K1 || j > k) { /* static for j >= K1, warp-uniform otherwise */
      const double xj = xs[j];
      double Ljk = col[j] * scale;
      Ljk = div_by(Ljk + s * xj, c, rc);
      Lk[(j * (j + 1) / 2) * kLs] = Ljk;
      xs[j] = c * xj - s * Ljk;
      fr[j] = fma(Ljk, Ljk, fr[j]);
    }
  }
}

template <int D, int K0, int KB>
__device__ __forceinline__ void shaper_block(double* Lc, double scale, double* xs, double* fr) {
  constexpr int K1 = (K0 + KB < D) ? K0 + KB : D;
#pragma unroll 1
  for (int k = K0; k < K1; k += 2) { /* two columns per trip; column k's stores never touch column k+1 */
    const bool two = (k + 1 < K1);
    double* Lk = Lc + k * kLs; /* column k: L(j,k) = Lk[(j(j+1)/2)*ls] */
    double col0[D], col1[D];
#pragma unroll
    for (int j = K0; j < D; ++j) col0[j] = SLGPU_LDL(Lk + (j * (j + 1) / 2) * kLs);
    if (two) {
#pragma unroll
      for (int j = K0; j < D; ++j) col1[j] = SLGPU_LDL(Lk + (j * (j + 1) / 2 + 1) * kLs);
    }
    shaper_col<D, K0, K1>(k, Lk, col0, scale, xs, fr);
    if (two) shaper_col<D, K0, K1>(k + 1, Lk + kLs, col1, scale, xs, fr);
  }
  if constexpr (K1 < D) shaper_block<D, K1, KB>(Lc, scale, xs, fr);
}

/* ProposalShaper::update (src/infer/adaptive.cpp:173-209), blocked column sweep; Lc[(r(r+1)/2+k)*ls].
 * jump[j*jst] holds the accepted jump. */
template <int D>
__device__ __forceinline__ double shaper_update_rolled(double* Lc, const double* jump, int jst, double count, double prop_norm) {
  const double eps = 1e-18;
  const double sq = sqrt(count);
  const double rsq = 1.0 / sq;
  const double scale = sqrt((count - 1.0) / count);
  double xs[D], f[D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
    xs[i] = div_by(jump[i * jst], sq, rsq);
    f[i] = 0.0;
  }
  shaper_block<D, 0, kColBlock>(Lc, scale, xs, f);
  /* |L|_F: slot l = row sum l (D <= 32), halving tree over 32 slots; slots >= D are +0 and are skipped */
  int nlive = D;
#pragma unroll
  for (int mm = 16; mm > 0; mm >>= 1) {
#pragma unroll
    for (int l = 0; l < mm; ++l)
      if (l + mm < D && l + mm < nlive) f[l] += f[l + mm];
    if (nlive > mm) nlive = mm;
  }
  const double frob = sqrt(f[0]);
  return prop_norm / smax(frob, eps);
}

/* ---- phase B with TWO lanes per staged chain: the update of one chain is a serial sequence of ~5600
 * instructions, and a block waits for it every round, so its rows are split over a lane pair (lane h takes
 * rows j = K0 + 2m + h of the column pair K0, K0+1); the column heads are computed by both lanes. ---- */
template <int D, int K>
__device__ __forceinline__ void shaper2_col(double* Lc, const double* col, double scale, int h, unsigned pm, double* xs, int xst,
                                            double* fr, int fst) {
  constexpr int K0 = K & ~1;                 /* first column of the pair this column belongs to */
  constexpr int NM = (D - K0 + 1) / 2;       /* row slots per lane */
  constexpr int TK = K * (K + 1) / 2;
  const double eps = 1e-18;
  const double xk = xs[K * xst];
  const double Lkk = Lc[(TK + K) * kLs] * scale;
  const double V = smax(eps, Lkk);
  const double rV = 1.0 / V;
  const double r = sqrt(V * V + xk * xk);
  const double c = div_by(r, V, rV);
  const double s = div_by(xk, V, rV);
  const double rc = 1.0 / c;
  __syncwarp(pm); /* both lanes have read the old diagonal and x_K */
  if (h == (K & 1)) {
    Lc[(TK + K) * kLs] = r;
    fr[K * fst] = fma(r, r, fr[K * fst]); /* the diagonal closes row K's sum */
  }
#pragma unroll
  for (int m = 0; m < NM; ++m) {
    const int j = K0 + 2 * m + h;
    if (j > K && j < D) {
      const int o = ((K0 + 2 * m) * (K0 + 2 * m + 1) / 2 + h * (K0 + 2 * m + 1) + K) * kLs;
      const double xj = xs[j * xst];
      double Ljk = col[m] * scale;
      Ljk = div_by(Ljk + s * xj, c, rc);
      Lc[o] = Ljk;
      xs[j * xst] = c * xj - s * Ljk;
      fr[j * fst] = fma(Ljk, Ljk, fr[j * fst]);
    }
  }
  __syncwarp(pm); /* x_{K+1} is visible to the partner before the next head */
}

template <int D, int K0>
__device__ __forceinline__ void shaper2_block(double* Lc, double scale, int h, unsigned pm, double* xs, int xst, double* fr, int fst) {
  constexpr int NM = (D - K0 + 1) / 2;
  constexpr bool two = (K0 + 1 < D);
  double col0[NM], col1[NM];
#pragma unroll
  for (int m = 0; m < NM; ++m) { /* both columns' rows of this lane first (in-triangle even where unused) */
    const int j = K0 + 2 * m + h;
    if (j < D) {
      const int o = ((K0 + 2 * m) * (K0 + 2 * m + 1) / 2 + h * (K0 + 2 * m + 1) + K0) * kLs;
      col0[m] = SLGPU_LDL(Lc + o);
      if (two) col1[m] = SLGPU_LDL(Lc + o + kLs);
    }
  }
  shaper2_col<D, K0>(Lc, col0, scale, h, pm, xs, xst, fr, fst);
  if constexpr (two) shaper2_col<D, K0 + 1>(Lc, col1, scale, h, pm, xs, xst, fr, fst);
  if constexpr (K0 + 2 < D) shaper2_block<D, K0 + 2>(Lc, scale, h, pm, xs, xst, fr, fst);
}

/* ProposalShaper::update by a lane pair; xs[j*xst] holds the accepted jump on entry, fr[j*fst] is scratch. */
template <int D>
__device__ __forceinline__ double shaper_update_pair2(double* Lc, double* xs, int xst, double* fr, int fst, double count,
                                                      double prop_norm, int h, unsigned pm) {
  const double eps = 1e-18;
  const double sq = sqrt(count);
  const double rsq = 1.0 / sq;
  const double scale = sqrt((count - 1.0) / count);
#pragma unroll
  for (int i = h; i < D; i += 2) {
    xs[i * xst] = div_by(xs[i * xst], sq, rsq);
    fr[i * fst] = 0.0;
  }
  __syncwarp(pm);
  shaper2_block<D, 0>(Lc, scale, h, pm, xs, xst, fr, fst);
  /* |L|_F: slot l = row sum l (D <= 32), halving tree over 32 slots (both lanes, same result) */
  double f[D];
#pragma unroll
  for (int j = 0; j < D; ++j) f[j] = fr[j * fst];
  int nlive = D;
#pragma unroll
  for (int mm = 16; mm > 0; mm >>= 1) {
#pragma unroll
    for (int l = 0; l < mm; ++l)
      if (l + mm < D && l + mm < nlive) f[l] += f[l + mm];
    if (nlive > mm) nlive = mm;
  }
  const double frob = sqrt(f[0]);
  return prop_norm / smax(frob, eps);
}

/* factor of the chain of thread slot ct of this block under ThreadMap (only called for active slots) */
__device__ __forceinline__ double* thread_l_base(const slgpu_step_params& p, u32 blk, int ct, u32 nL) {
  return p.L + (size_t)(blk * (blockDim.x >> 5) + (u32)(ct >> 5)) * nL * SLGPU_L_TILE + (ct & 31);
}

#ifdef SLGPU_STAMPS /* development: block-level time stamps (ns) into the round_flags buffer, 32 u64 per block */
#define SLGPU_STAMP(k)                                                                        \
  do {                                                                                        \
    if (threadIdx.x == 0 && p.round_flags) {                                                  \
      unsigned long long t_;                                                                  \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                   \
      ((unsigned long long*)p.round_flags)[(size_t)blockIdx.x * 32 + (k)] = t_;               \
    }                                                                                         \
  } while (0)
#else
#define SLGPU_STAMP(k) do {} while (0)
#endif

/* dynamic shared memory of pt_step_fast_kernel */
template <int D, int NT>
inline size_t fast_smem_bytes(unsigned n_temps) {
  const size_t nsb = (size_t)(NT / 32) * (32u / n_temps);
  return ((size_t)2 * D * NT + (size_t)D * (NT + 1) + (size_t)2 * D * nsb) * sizeof(double) + SLGPU_FAST_SMEM_PAD;
}

/* WELFORD = false: the reference's ProposalShaper (rank-1 Cholesky update on every accept), bit for bit.
 * WELFORD = true ("gpu": {"shaper": "welford"}, NOT the reference's arithmetic): on an accept the thread updates the second-moment
 * matrix of its chain element by element -- no serial column sweep, no staging, no block barrier in the round -- and the
 * factor is re-computed between launches (rechol_kernel).  north_star's "Welford rank-1 covariance update with periodic
 * re-Cholesky", validated statistically only (tests/test_welford_gpu.py). */
template <class Lik, int D, int NT, bool WELFORD = false>
__global__ void __launch_bounds__(NT, SLGPU_FAST_MINBLOCKS) pt_step_fast_kernel(const slgpu_step_params p) {
  const u32 blk = blockIdx.x; /* stack group of this block */
  const ThreadMap m(p, blk);
  const int T = m.T;
  const size_t C = (size_t)p.n_stacks * p.n_temps;
  const bool active = m.active;
  const u32 c = active ? m.c : 0u;
  const u32 gchain = p.stack_offset * p.n_temps + c;
  const bool replay = (p.rng_mode == SLGPU_RNG_REPLAY);
  const int tid = threadIdx.x;

  __shared__ double s_mn[D], s_mx[D], s_n0[D], s_rd[D];
  __shared__ double s_lad[SLGPU_MAX_TEMPS];
  /* dynamic shared memory (see fast_smem_bytes): three [D][NT] panels + the EPSR state of the block's stacks */
  extern __shared__ double s_dyn[];
  double* const s_jump = s_dyn;              /* staged accepted jumps, [i][slot] */
  double* const s_z = s_dyn + D * NT;        /* this round's normals, [k][thread] */
  constexpr int NX = NT + 1;                 /* row stride of s_x: odd, so that a column (one chain, all i) is conflict-free too */
  double* const s_x = s_dyn + 2 * D * NT;    /* current sample of every chain of the block, [i][thread], rows NX apart */
  const int nsb = (NT / 32) * m.spw;         /* stacks per block */
  double* const s_epm = s_dyn + 2 * D * NT + D * NX;  /* EPSRDiagnostic M, [stack of block][i] */
  double* const s_eps = s_epm + D * nsb;     /* EPSRDiagnostic S */
  const int sb = (tid >> 5) * m.spw + m.sl;  /* this thread's stack within the block */
  __shared__ double s_cnt[NT];      /* shaper count (already incremented) of the staged chain */
  __shared__ double s_nsc[NT];      /* result: new N scale, indexed by owning thread */
  __shared__ unsigned short s_owner[NT]; /* owning thread of the staged chain (its chain id follows from the thread map) */
  __shared__ int s_nacc[2];
  SLGPU_STAMP(0);
  for (int i = tid; i < D; i += NT) {
    s_mn[i] = p.mn[i];
    s_mx[i] = p.mx[i];
    s_n0[i] = p.n0_diag[i];
    s_rd[i] = 1.0 / (p.mx[i] - p.mn[i]);
  }
  /* s_w / s_lad (the tier models' outputs) and the pending regression statistics are loaded in round 0, behind
   * cudaGridDependencySynchronize(): this kernel is launched with programmatic stream serialization, so everything up
   * to there -- state loads, the first round's draws, proposal, likelihood and MH test -- may overlap the k_tier_adapt
   * launch in front of it (which only produces those values) */
  if (tid < 2) s_nacc[tid] = 0;
  __syncthreads();

  const Lik lik(p.payload);
  const double* Lc = l_base(p.L, p.n_temps, active ? m.s : 0u, (u32)m.t, D * (D + 1) / 2); /* rewritten by other threads in phase B: no __restrict__ */
#pragma unroll
  for (int i = 0; i < D; ++i) s_x[i * NX + tid] = p.x[(size_t)i * C + c];
  if (active && m.t == 0) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
      s_epm[sb * D + i] = p.ep_m[(size_t)i * p.n_stacks + m.s];
      s_eps[sb * D + i] = p.ep_s[(size_t)i * p.n_stacks + m.s];
    }
  }
  double Ps[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) Ps[k] = 0.0;
  double energy = p.energy[c], row_sigma = p.row_sigma[c], row_beta = p.row_beta[c];
  double sigma = p.sigma[c], beta = p.beta[c], nscale = p.nscale[c], sh_count = p.sh_count[c];
  int accepted = p.accepted[c], swap_type = p.swap_type[c];
  u32 lg_accepts = p.lg_accepts[c], lg_swaps = p.lg_swaps[c], lg_att = p.lg_swap_attempts[c];
  double lg_min_e = p.lg_min_energy[c];
  int n_logged = WELFORD ? (int)p.chol_dirty[c] : 0; /* jumps of this chain waiting for the next re-Cholesky */
  double w3[3] = {0.0, 0.0, 0.0};
  const int W = (int)p.swap_interval;
  SLGPU_STAMP(1);

  for (u32 it = 0; it < p.n_rounds; ++it) {
    const u64 r = p.round_begin + it;
    const int par = (int)(it & 1u);
    bool acc = false;
    /* ================= phase A: propose, evaluate, accept, sigma ================= */
    if (active) {
      const bool fresh = WELFORD ? (nscale == 0.0) : (sh_count == p.sh_count0); /* N still the constructor's diag(range/|range|) */
      const double* zr = replay ? p.rp_z + ((size_t)r * C + c) * D : nullptr;
      double a[D];
#pragma unroll
      for (int i = 0; i < D; ++i) a[i] = 0.0;
      /* the round's normals go to this thread's column of s_z (rolled loop: one copy of the generator) */
#pragma unroll kNormalsUnroll
      for (int kp = 0; kp < (D + 1) / 2; ++kp) {
        double z0, z1 = 0.0;
        if (replay) {
          z0 = zr[2 * kp];
          if (2 * kp + 1 < D) z1 = zr[2 * kp + 1];
        } else {
          double ua, ub, sn, cs;
          philox_draw(p.seed, gchain, r, 0, (u32)kp, ua, ub);
          const double R = sqrt(-2.0 * slm_log(ua));
          slm_sincos2pi(ub, &sn, &cs);
          z0 = R * cs;
          z1 = R * sn;
        }
        s_z[(2 * kp) * NT + tid] = z0;
        if (2 * kp + 1 < D) s_z[(2 * kp + 1) * NT + tid] = z1;
      }
      /* GaussianProposal::propose (sampler.cpp:79-88): a_i = sum_{k<=i} fma(N_ik, z_k, .), k ascending */
      gemv_block<D, 0, kColBlock>(Lc, s_z + tid, NT, nscale, fresh, s_n0, a);
#pragma unroll
      for (int i = 0; i < D; ++i) a[i] = bounce1(s_x[i * NX + tid] + a[i] * sigma, s_mn[i], s_mx[i], s_rd[i]);
      /* energy = sum over job types in order (sampler.cpp:148-150) */
      double e_new = 0.0;
      if constexpr (StaticJobs<Lik, D>::value > 0) {
#pragma unroll
        for (u32 j = 0; j < (u32)StaticJobs<Lik, D>::value; ++j) e_new += lik(j, a, (u32)D);
      } else {
        for (u32 j = 0; j < p.n_job_types; ++j) e_new += lik(j, a, (u32)D);
      }
      /* acceptProposal (chainarray.cpp:30-45) + append (:94-121) */
      if (!isinf(e_new)) {
        double u;
        if (replay) u = p.rp_umh[(size_t)r * C + c];
        else { double ub; philox_draw(p.seed, gchain, r, 1, 0, u, ub); }
        const double deltaEnergy = e_new - energy;
        acc = u < slm_exp(-1.0 * beta * deltaEnergy);
      }
      if (acc) {
        const int slot = WELFORD ? tid : atomicAdd(&s_nacc[par], 1); /* WELFORD: every chain updates its own moment matrix */
        sh_count += 1.0;
#pragma unroll
        for (int i = 0; i < D; ++i) {
          const double nx = a[i];
          s_jump[i * NT + slot] = nx - s_x[i * NX + tid]; /* the accepted jump (sampler.cpp:166) */
          s_x[i * NX + tid] = nx;
        }
        if constexpr (WELFORD) {
          /* the jump goes to the chain's log; rechol_kernel folds it into the second-moment matrix at the next adapt point */
#pragma unroll
          for (int i = 0; i < D; ++i) __stcs(p.jump_log + ((size_t)n_logged * D + i) * C + c, s_jump[i * NT + tid]);
          n_logged += 1;
        } else {
          s_cnt[slot] = sh_count;
          s_owner[slot] = (unsigned short)tid;
        }
        energy = e_new;
        row_sigma = sigma;
        row_beta = beta;
      }
      accepted = acc ? 1 : 0;
      swap_type = 0;
    }
    if (it == 0) { /* from here on the launch needs what k_tier_adapt wrote: wait for it (a no-op without PDL) */
      cudaGridDependencySynchronize();
      if (tid < T) s_lad[tid] = p.ladder[tid]; /* read after the next block barrier */
      if (active) {
        w3[0] = p.sig_w[3 * m.t]; w3[1] = p.sig_w[3 * m.t + 1]; w3[2] = p.sig_w[3 * m.t + 2];
#pragma unroll
        for (int k = 0; k < 9; ++k) Ps[k] = p.p_sig[pidx(p, k, m.t, m.s)];
      }
    }
    if (active) {
      /* sigma adaptation (sampler.cpp:160-162): statistics summed, model frozen for the launch */
      const double logtemper = -slm_log(row_beta);
      stats_accumulate<1>(Ps, -8.0, 4.0, slm_log(row_sigma), logtemper, acc);
      sigma = slm_exp(predict(w3, p.opt_accept, -8.0, 4.0, logtemper));
    }
    if constexpr (!WELFORD) {
    __syncthreads();
    SLGPU_STAMP(2 + 3 * it);
    /* ================= phase B: compacted ProposalShaper updates ================= */
    {
      const int nacc = s_nacc[par];
      /* slot -> thread map: the staged chains are dealt round-robin to the warps (with few warps resident
       * per SM, parallel part-filled warps beat one full warp) */
#if !defined(SLGPU_SLOTMAP_STRIDED) && !defined(SLGPU_SLOTMAP_CONTIGUOUS) /* measured on config #3: dense 74, strided 78, contiguous runs 95 us per round */
#ifndef SLGPU_SLOT_FIXED /* measured on config #3: rotating 56.4 us per round, always warp 0 first 57.1 */
      /* fill whole warps, but start at a different warp every round and block: a warp sits on a fixed scheduler,
       * and always packing into warp 0 would load the same scheduler's FP64 pipe in every block of the SM */
      const int slot = (tid + NT - 32 * (int)((it + blk) % (u32)(NT / 32))) & (NT - 1);
#else
      const int slot = tid; /* fill the first warps completely */
#endif
      const bool mine = slot < nacc;
#elif defined(SLGPU_SLOTMAP_STRIDED)
      const int slot = (tid & 31) * (NT / 32) + (tid >> 5);
      const bool mine = slot < nacc;
#else
      const int per = (nacc + NT / 32 - 1) / (NT / 32);
      const int slot = (tid >> 5) * per + (tid & 31);
      const bool mine = (tid & 31) < per && slot < nacc;
#endif
#ifndef SLGPU_PHASEB_PAIR /* measured on config #3: one lane per staged chain 59.7 us per round, a lane pair 68.8 */
      if (mine) { /* jump in s_jump[.][slot] */
        const int own = s_owner[slot];
        s_nsc[own] = shaper_update_rolled<D>(thread_l_base(p, blk, own, D * (D + 1) / 2), s_jump + slot, NT, s_cnt[slot], p.prop_norm);
      }
#else
      (void)mine;
      /* a lane pair per staged chain (threads 2u, 2u+1 take slot u, u + NT/2, ...): jump in s_jump[.][u], row sums
       * in column u of s_z (every z is dead by now) */
      const unsigned pm = 3u << (m.lane & ~1);
      for (int u = tid >> 1; u < nacc; u += NT / 2) {
        const double ns = shaper_update_pair2<D>(thread_l_base(p, blk, s_owner[u], D * (D + 1) / 2), s_jump + u, NT, s_z + u, NT, s_cnt[u],
                                                 p.prop_norm, tid & 1, pm);
        if ((tid & 1) == 0) s_nsc[s_owner[u]] = ns;
      }
#endif
      if (tid == 0) s_nacc[par ^ 1] = 0;
    }
    __syncthreads();
    SLGPU_STAMP(3 + 3 * it);
    if (acc) nscale = s_nsc[tid];
    } else {
      if (it == 0) __syncthreads(); /* s_lad (loaded behind the grid dependency in round 0) is read in the swap sweep */
    }
    /* the table logger's per-round log: written here, where acc dies; the swap sweep below adds its two bits */
#ifndef SLGPU_STAMPS
    if (p.round_flags && active) p.round_flags[(size_t)it * C + c] = (unsigned char)(acc ? 1 : 0);
#endif

    /* ================= swap sweep, hottest pair first ================= */
    double mid_e = energy;
    int mid_acc = accepted;
    const bool swapRound = (T > 1) && ((r + 2) % (u64)W == 0);
    if (swapRound) {
      double u_sw = 0.0;
      if (active && m.t < T - 1) {
        if (replay) u_sw = p.rp_uswap[(size_t)r * C + c];
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
      const double lb = slm_log(beta);
      const double lbh = shfl_d(lb, m.lane + 1 < 32 ? m.lane + 1 : m.lane);
      mid_e = shfl_d(energy, m.base + mid_src);
      mid_acc = __shfl_sync(0xffffffffu, accepted, m.base + mid_src);
      const double new_e = shfl_d(energy, m.base + my_src);
      const int new_acc = __shfl_sync(0xffffffffu, accepted, m.base + my_src);
      { /* gather the sample from the source chain's column of s_x (same warp) */
        const int src_tid = (tid & ~31) + ((m.base + my_src) & 31);
        double g[D];
#pragma unroll
        for (int i = 0; i < D; ++i) g[i] = s_x[i * NX + src_tid];
        __syncwarp();
#pragma unroll
        for (int i = 0; i < D; ++i) s_x[i * NX + tid] = g[i];
        __syncwarp();
      }
      energy = new_e;
      accepted = new_acc;
      if (active && m.t < T - 1) {
        swap_type = my_swapped ? 1 : 2;
#ifndef SLGPU_STAMPS
        if (p.round_flags) p.round_flags[(size_t)it * C + c] |= (unsigned char)(swap_type << 1);
#endif
        double Pb[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) Pb[k] = p.p_beta[pidx(p, k, m.t, m.s)];
        stats_accumulate<1>(Pb, 0.0, 4.0, lb - lbh, lb, my_swapped);
#pragma unroll
        for (int k = 0; k < 9; ++k) p.p_beta[pidx(p, k, m.t, m.s)] = Pb[k];
      }
      if (m.t > 0) beta = s_lad[m.t];
    }

    if (active) {
      lg_min_e = smin(lg_min_e, mid_e);
      lg_accepts += (u32)mid_acc;
      lg_swaps += (swap_type == 1);
      lg_att += (swap_type != 0);
      if (it + 1 == p.n_rounds) p.lg_cur_energy[c] = mid_e; /* TableLogger energies_ (logging.cpp:43) */
      if (m.t == 0 && p.rows_out) { /* the five scalar fields of the cold row (db.cpp:18-25) */
        double* row = p.rows_out + ((size_t)it * p.n_stacks + m.s) * (D + SLGPU_ROW_EXTRA);
        row[D] = energy;
        row[D + 1] = row_sigma;
        row[D + 2] = row_beta;
        row[D + 3] = (double)accepted;
        row[D + 4] = (double)swap_type;
      }
    }
    /* EPSRDiagnostic::update (diagnostics.cpp:27-47) and the sample fields of the cold rows: the (cold chain,
     * dimension) pairs of this warp are dealt to all 32 lanes instead of 20 dimensions on each of the few cold lanes */
    __syncwarp();
    {
      const double nn = (double)(r + 1);
      const double rnn = 1.0 / nn;
      const u32 s0 = (blk * (NT / 32) + (u32)(tid >> 5)) * (u32)m.spw; /* first stack of this warp */
      for (int task = m.lane; task < m.spw * D; task += 32) {
        const int sl = task / D, i = task - sl * D;
        const u32 sc = s0 + (u32)sl;
        if (sc < p.n_stacks) {
          const int o = ((tid >> 5) * m.spw + sl) * D + i; /* [stack][i]: consecutive lanes, consecutive words */
          const double mo = s_epm[o];
          const double xi = s_x[i * NX + (tid & ~31) + sl * T];
          const double newM = mo + div_by(xi - mo, nn, rnn);
          s_eps[o] = s_eps[o] + (xi - mo) * (xi - newM);
          s_epm[o] = newM;
          if (p.rows_out) p.rows_out[((size_t)it * p.n_stacks + sc) * (D + SLGPU_ROW_EXTRA) + i] = xi;
        }
      }
    }
    SLGPU_STAMP(4 + 3 * it);
  }

  if (active) {
#pragma unroll
    for (int i = 0; i < D; ++i) p.x[(size_t)i * C + c] = s_x[i * NX + tid];
    if (m.t == 0) {
#pragma unroll
      for (int i = 0; i < D; ++i) {
        p.ep_m[(size_t)i * p.n_stacks + m.s] = s_epm[sb * D + i];
        p.ep_s[(size_t)i * p.n_stacks + m.s] = s_eps[sb * D + i];
      }
    }
    p.energy[c] = energy;
    p.row_sigma[c] = row_sigma;
    p.row_beta[c] = row_beta;
    p.sigma[c] = sigma;
    p.beta[c] = beta;
    if constexpr (WELFORD) p.chol_dirty[c] = (unsigned char)n_logged; /* and rechol_kernel owns nscale */
    else p.nscale[c] = nscale;
    p.sh_count[c] = sh_count;
    p.accepted[c] = accepted;
    p.swap_type[c] = swap_type;
    p.lg_accepts[c] = lg_accepts;
    p.lg_swaps[c] = lg_swaps;
    p.lg_swap_attempts[c] = lg_att;
    p.lg_min_energy[c] = lg_min_e;
#pragma unroll
    for (int k = 0; k < 9; ++k) p.p_sig[pidx(p, k, m.t, m.s)] = Ps[k];
  }
}

/* "gpu": {"shaper": "welford"}, at the adapt points: for every chain that logged accepted jumps, fold them into its
 * second-moment matrix in the order they happened -- C <- (n-1)/n C + s s^T / n with n the shaper count after that accept,
 * the recurrence L L^T obeys under ProposalShaper::update (adaptive.cpp:179-197) -- then L = chol(C) and
 * nscale = prop_norm / |L|_F (adaptive.cpp:200).  One block per stack group: its tile of 32 matrices is staged in shared
 * memory; lane = chain, and the kRecholWarps warps split the matrix ROWS (row j belongs to warp j % kRecholWarps), so every
 * shared-memory access is a full 32-lane row of the tile and a thread carries several independent dot products.  Left-looking
 * Cholesky: the owner of row k publishes the pivot, a barrier, every warp scales its rows.  A non-positive pivot (rounding
 * only: C is a positive diagonal plus outer products) is floored. */
#ifndef SLGPU_RECHOL_WARPS
#define SLGPU_RECHOL_WARPS 4
#endif
constexpr int kRecholWarps = SLGPU_RECHOL_WARPS;

/* the work of warp W of a rechol block: its rows j = W, W + NW, ... of the 32 matrices of the tile live in REGISTERS from the
 * fold to the write-back (every index below is a compile-time constant once W is), the tile in shared memory carries what
 * the warps exchange: the original diagonal, and the finished entries of L, which the other warps read as row k */
template <int D, int NW, int W>
__device__ __forceinline__ void rechol_rows(const slgpu_step_params& p, double* s_t, double* s_j, double* s_fr, unsigned long long* s_bar,
                                            int lane, u32 c, size_t C, bool active, int n_logged, int n_max, double* Cg, double* Lg) {
  constexpr int RPT = (D - W + NW - 1) / NW; /* rows of this warp */
  double R[RPT > 0 ? RPT : 1][D];            /* R[r][k], k <= j_r = W + r NW */
  double* A = s_t + lane;                    /* this chain's matrix in the tile: A[idx * 32] */
  const double cnt = active ? p.sh_count[c] : 1.0;
  double nxt[RPT > 0 ? RPT : 1]; /* this warp's rows of the NEXT logged jump, in flight while the current one is folded */
#pragma unroll
  for (int r = 0; r < RPT; ++r) nxt[r] = (0 < n_logged) ? __ldcs(p.jump_log + (size_t)(W + r * NW) * C + c) : 0.0;
  for (int q = 0; q < n_max; ++q) {
    if (q) __syncthreads(); /* the previous jump has been consumed */
    const bool has = q < n_logged;
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      s_j[(W + r * NW) * 32 + lane] = nxt[r];
      nxt[r] = (q + 1 < n_logged) ? __ldcs(p.jump_log + ((size_t)(q + 1) * D + (W + r * NW)) * C + c) : 0.0;
    }
    __syncthreads();
    if (q == 0) {
      mbar_wait(s_bar, 0); /* the tile is here: take the rows */
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        constexpr int dummy = 0;
        (void)dummy;
        const int j = W + r * NW;
#pragma unroll
        for (int k = 0; k < D; ++k)
          if (k <= j) R[r][k] = A[(j * (j + 1) / 2 + k) * 32];
      }
    }
    if (has) {
      const double n = cnt - (double)(n_logged - 1 - q);
      const double fold = (n - 1.0) / n, rn = 1.0 / n;
      double sjn[RPT > 0 ? RPT : 1];
#pragma unroll
      for (int r = 0; r < RPT; ++r) sjn[r] = s_j[(W + r * NW) * 32 + lane] * rn;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        if (k <= W + (RPT - 1) * NW) { /* some row of this warp reaches column k */
          const double sk = s_j[k * 32 + lane];
#pragma unroll
          for (int r = 0; r < RPT; ++r)
            if (k <= W + r * NW) R[r][k] = fma(sjn[r], sk, R[r][k] * fold);
        }
      }
    }
  }
  const bool redo = n_logged > 0;
  /* the moment matrix itself goes back before it is factorised; the diagonal also to the tile (the other warps' pivots) */
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    const int j = W + r * NW;
#pragma unroll
    for (int k = 0; k < D; ++k)
      if (k <= j) {
        if (redo) Cg[(size_t)(j * (j + 1) / 2 + k) * 32 + lane] = R[r][k];
        if (k == j) A[(j * (j + 1) / 2 + k) * 32] = R[r][k];
      }
  }
  __syncthreads();
  double frob = 0.0;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    constexpr int dummy2 = 0;
    (void)dummy2;
    const int rk = k * (k + 1) / 2;
    double acc[RPT > 0 ? RPT : 1];
#pragma unroll
    for (int r = 0; r < RPT; ++r) acc[r] = (W + r * NW > k) ? R[r][k] : 0.0;
    double piv = A[(rk + k) * 32]; /* C[k][k] */
#pragma unroll
    for (int m = 0; m < k; ++m) {
      const double akm = A[(rk + m) * 32]; /* L[k][m], finished in column m */
      piv = fma(-akm, akm, piv);
#pragma unroll
      for (int r = 0; r < RPT; ++r)
        if (W + r * NW > k) acc[r] = fma(-R[r][m], akm, acc[r]);
    }
    const double lkk = sqrt(smax(piv, 1e-300));
    const double rl = 1.0 / lkk;
    if (k % NW == W) {
      R[(k - W) / NW < RPT ? (k - W) / NW : 0][k] = lkk; /* (the guard only keeps the index in range for other W) */
      frob = fma(lkk, lkk, frob);
    }
#pragma unroll
    for (int r = 0; r < RPT; ++r)
      if (W + r * NW > k) {
        const double w = acc[r] * rl;
        R[r][k] = w;
        A[((W + r * NW) * (W + r * NW + 1) / 2 + k) * 32] = w; /* row j is read by everybody when it becomes the pivot row */
        frob = fma(w, w, frob);
      }
    __syncthreads();
  }
  s_fr[W * 32 + lane] = frob;
  if (redo) {
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int j = W + r * NW;
#pragma unroll
      for (int k = 0; k < D; ++k)
        if (k <= j) Lg[(size_t)(j * (j + 1) / 2 + k) * 32 + lane] = R[r][k];
    }
  }
  __syncthreads();
  if (W == 0 && redo) {
    double f = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) f += s_fr[w * 32 + lane];
    p.nscale[c] = p.prop_norm / smax(sqrt(f), 1e-18);
    p.chol_dirty[c] = 0;
  }
}

template <int D>
__global__ void __launch_bounds__(kRecholWarps * 32) rechol_kernel(const slgpu_step_params p) {
  constexpr int NL = D * (D + 1) / 2;
  constexpr int NW = kRecholWarps;
  extern __shared__ double s_t[]; /* [NL][32] the tile, [D][32] the jump being folded, [NW][32] partial norms */
  double* s_j = s_t + NL * 32;
  double* s_fr = s_j + D * 32;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int T = (int)p.n_temps, spw = 32 / T, sl = lane / T;
  const u32 s = blockIdx.x * (u32)spw + (u32)sl;
  const bool active = sl < spw && s < p.n_stacks;
  const u32 c = active ? s * (u32)T + (u32)(lane - sl * T) : 0u;
  const size_t C = (size_t)p.n_stacks * p.n_temps;
  const int n_logged = active ? (int)p.chol_dirty[c] : 0;
  const int n_max = __reduce_max_sync(0xffffffffu, n_logged); /* every warp sees the same 32 chains */
  if (n_max == 0) return;
  double* Cg = p.moment + (size_t)blockIdx.x * NL * 32;
  double* Lg = p.L + (size_t)blockIdx.x * NL * 32;
  __shared__ unsigned long long s_bar;
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    mbar_expect_tx(&s_bar, NL * 32 * (unsigned)sizeof(double));
    bulk_load(s_t, Cg, NL * 32 * (unsigned)sizeof(double), &s_bar); /* the whole tile in one bulk copy */
  }
  static_assert(NW == 4 || NW == 2 || NW == 8, "rechol_rows is instantiated per warp below");
  switch (wid) { /* warp-uniform: every index inside is static */
    case 0: rechol_rows<D, NW, 0>(p, s_t, s_j, s_fr, &s_bar, lane, c, C, active, n_logged, n_max, Cg, Lg); break;
    case 1: rechol_rows<D, NW, 1>(p, s_t, s_j, s_fr, &s_bar, lane, c, C, active, n_logged, n_max, Cg, Lg); break;
    case 2: if constexpr (NW > 2) rechol_rows<D, NW, 2>(p, s_t, s_j, s_fr, &s_bar, lane, c, C, active, n_logged, n_max, Cg, Lg); break;
    case 3: if constexpr (NW > 2) rechol_rows<D, NW, 3>(p, s_t, s_j, s_fr, &s_bar, lane, c, C, active, n_logged, n_max, Cg, Lg); break;
    case 4: if constexpr (NW > 4) rechol_rows<D, NW, 4>(p, s_t, s_j, s_fr, &s_bar, lane, c, C, active, n_logged, n_max, Cg, Lg); break;
    case 5: if constexpr (NW > 4) rechol_rows<D, NW, 5>(p, s_t, s_j, s_fr, &s_bar, lane, c, C, active, n_logged, n_max, Cg, Lg); break;
    case 6: if constexpr (NW > 4) rechol_rows<D, NW, 6>(p, s_t, s_j, s_fr, &s_bar, lane, c, C, active, n_logged, n_max, Cg, Lg); break;
    default: if constexpr (NW > 4) rechol_rows<D, NW, 7>(p, s_t, s_j, s_fr, &s_bar, lane, c, C, active, n_logged, n_max, Cg, Lg); break;
  }
}

}  // namespace slgpu

#endif
