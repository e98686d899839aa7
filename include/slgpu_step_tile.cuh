/*
 * slgpu_step_tile.cuh -- pt_step for a compile-time dimension D <= 32: ONE THREAD BLOCK PER STACK GROUP, the
 * group's proposal factors resident in shared memory for the whole launch (included by slgpu_device.cuh).
 *
 * Same rounds, same arithmetic in the same order as pt_step_kernel and the oracle.  What differs is how the work is
 * laid on the SM:
 *   - a stack group = the floor(32 / T) stacks whose chains fill one warp's lanes (lane = chain, as in every other
 *     kernel, so the hot->cold swap sweep of a stack stays a warp-shuffle exchange);
 *   - the group's packed factors ([d(d+1)/2][32 slots] doubles, 52.5 KB at d = 20) are ONE contiguous tile of the L
 *     store: one cp.async.bulk brings it into shared memory when the block starts (completion on an mbarrier, the
 *     block's other state loads overlap it), every round of the launch reads and updates it there, one
 *     cp.async.bulk writes it back at the end.  HBM sees every factor once per launch in each direction; no round
 *     leans on L2 residency;
 *   - the block has NW warps and all of them map lane -> chain; what differs between the warps is the ROWS they
 *     own: warp g owns rows j = g, g + NW, ... of x, of the proposal and of L.  A row of L is only ever touched by
 *     its owner: the proposal GEMV is a dot product per row (k ascending, the oracle's order), and in
 *     ProposalShaper::update (adaptive.cpp:173-209) the owner of row j applies the rotations of columns 0..j-1 to its
 *     row and then computes column j's head (r, c, s, 1/c) itself -- 32 chains per instruction, nothing replicated;
 *   - the rank-1 update is a serial chain head_k -> L(k+1,k), x_{k+1} -> head_{k+1} of ~30 dependent fp64 operations
 *     per column.  The owner of row k+1 runs that chain first (its divisor 1/V is precomputed as soon as the row's
 *     previous head is done), publishes head k+1 through shared memory, and one block barrier per column hands it to
 *     the other warps, whose own rows of column k fill the time in between;
 *   - per-chain scalar work (energy sum, MH test, sigma regression sample, swap sweep, logger counters) is warp 0's;
 *     normals are dealt to the warps by pair index, likelihood job types by job index.
 * Three blocks of 8 warps are resident per SM at d = 20 (shared memory: 3 x 71 KB).
 */
#ifndef SLGPU_STEP_TILE_CUH
#define SLGPU_STEP_TILE_CUH

#include "slgpu_bulk.cuh"

namespace slgpu {

#ifndef SLGPU_TILE_MINBLOCKS
#define SLGPU_TILE_MINBLOCKS 3
#endif

/* warps per stack group for dimension D: enough rows per warp to keep every warp busy in the column sweep */
template <int D>
struct TileWarps {
#ifdef SLGPU_TILE_WARPS
  static constexpr int value = (D >= SLGPU_TILE_WARPS) ? SLGPU_TILE_WARPS : (D >= 4 ? 4 : (D >= 2 ? 2 : 1));
#else
  static constexpr int value = (D >= 16) ? 8 : (D >= 6 ? 4 : (D >= 2 ? 2 : 1));
#endif
};

/* shared-memory plan of pt_step_tile_kernel, in doubles */
template <int D>
struct TilePlan {
  static constexpr int NL = D * (D + 1) / 2;
  static constexpr int XS = D | 1;          /* odd row stride of the per-chain proposal panel: conflict-free columns */
  static constexpr int oL = 0;              /* [NL][32]  the factors */
  static constexpr int oZ = oL + NL * 32;   /* [D][32]   normals of the round */
  static constexpr int oXp = oZ + D * 32;   /* [32][XS]  the proposal of every chain, contiguous per chain (functor argument); later [D][32] row sums of squares */
  static constexpr int oHead = oXp + 32 * XS; /* [2][3][32] column heads c, s, 1/c, double buffered by column parity */
  static constexpr int oScal = oHead + 192; /* [6][32]   sigma, nscale, sqrt(count), 1/sqrt(count), scale, spare */
  static constexpr int oBnd = oScal + 192;  /* [4][D]    min, max, 1/(max-min), diag of the initial N */
  static constexpr int oPs = oBnd + 4 * D;  /* [9][32]   sigma regression statistics of the launch (warp 0) */
  static constexpr int oF = oPs + 9 * 32;   /* [J][32]   likelihood of every job type */
  /* + [2][D][stacks of the group] EPSR state behind the job panel */
  static size_t bytes(unsigned n_job_types, unsigned n_temps) {
    return ((size_t)oF + (size_t)n_job_types * 32 + (size_t)2 * D * (32u / n_temps)) * sizeof(double);
  }
};

#ifdef SLGPU_STAMPS /* development: per-block clock stamps of round SLGPU_STAMPS into the round_flags buffer, 32 x i64 per block */
#define SLGPU_TSTAMP(k)                                                                             \
  do {                                                                                              \
    if (threadIdx.x == 0 && p.round_flags && it == (SLGPU_STAMPS))                                  \
      ((long long*)p.round_flags)[(size_t)blockIdx.x * 32 + (k)] = clock64();                       \
  } while (0)
#else
#define SLGPU_TSTAMP(k) do {} while (0)
#endif

template <class Lik, int D, int NW>
__global__ void __launch_bounds__(32 * NW, SLGPU_TILE_MINBLOCKS) pt_step_tile_kernel(const slgpu_step_params p) {
  using Plan = TilePlan<D>;
  constexpr int RM = (D + NW - 1) / NW; /* rows per warp */
  constexpr int NL = Plan::NL, XS = Plan::XS, NP = (D + 1) / 2;
  enum { SIG = Plan::oScal, NSC = Plan::oScal + 32, SQ = Plan::oScal + 64, RSQ = Plan::oScal + 96, SCL = Plan::oScal + 128 };
  /* every shared-memory access below is s_dyn[offset] with integer offsets: pointers into shared memory would be
   * generic ones, and their address arithmetic ends up inside the loops */
  extern __shared__ __align__(16) double s_dyn[];
  __shared__ unsigned long long s_bar;
  __shared__ int s_acc[32], s_src[32], s_fresh[32];
  __shared__ double s_w[SLGPU_MAX_TEMPS * 3], s_lad[SLGPU_MAX_TEMPS];

  int lane, tidx;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane)); /* volatile: kept in a register, not re-read at every use */
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tidx));
  const int g = tidx >> 5, tid = tidx;
  const int T = (int)p.n_temps, spw = 32 / T;
  const int sl = lane / T;
  const bool in_stack = sl < spw;
  const int t = in_stack ? lane - sl * T : 0;
  const int base = in_stack ? sl * T : lane; /* idle tail lanes keep every shuffle source inside the warp */
  const u32 s = blockIdx.x * (u32)spw + (u32)sl;
  const bool active = in_stack && s < p.n_stacks;
  const u32 c = active ? s * (u32)T + (u32)t : 0u; /* idle lanes shadow chain 0 for loads and never store */
  const size_t C = (size_t)p.n_stacks * p.n_temps;
  const u32 gchain = p.stack_offset * p.n_temps + c;
  const bool replay = (p.rng_mode == SLGPU_RNG_REPLAY);
  const int J = (int)p.n_job_types;
  const int W = (int)p.swap_interval;
  const int oLl = Plan::oL + lane; /* entry idx of this lane's chain: s_dyn[oLl + idx * 32] */
  const int oZl = Plan::oZ + lane;
  const int oXl = Plan::oXp + lane * XS; /* this chain's proposal, contiguous */
  const int oFr = Plan::oXp + lane;      /* the same panel as [D][32] row sums at the end of the column sweep */
  const int oHl = Plan::oHead + lane;
  const int oEp = Plan::oF + J * 32;     /* EPSRDiagnostic M and S of the group's stacks: [2][D][spw] */

  double* const gL = p.L + (size_t)blockIdx.x * NL * 32;
  if (tid == 0) mbar_init(&s_bar, 1);
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&s_bar, NL * 32 * (unsigned)sizeof(double));
    bulk_load(&s_dyn[Plan::oL], gL, NL * 32 * (unsigned)sizeof(double), &s_bar);
  }
  /* everything else the block needs, while the tile is in flight */
  for (int i = tid; i < D; i += 32 * NW) {
    s_dyn[Plan::oBnd + i] = p.mn[i];
    s_dyn[Plan::oBnd + D + i] = p.mx[i];
    s_dyn[Plan::oBnd + 2 * D + i] = 1.0 / (p.mx[i] - p.mn[i]);
    s_dyn[Plan::oBnd + 3 * D + i] = p.n0_diag[i];
  }
  for (int i = tid; i < T; i += 32 * NW) {
    s_w[3 * i + 0] = p.sig_w[3 * i + 0];
    s_w[3 * i + 1] = p.sig_w[3 * i + 1];
    s_w[3 * i + 2] = p.sig_w[3 * i + 2];
    s_lad[i] = p.ladder[i];
  }
  for (int i = tid; i < D * spw; i += 32 * NW) { /* [j][stack of the group] */
    const int j = i / spw, q = i - j * spw;
    const u32 sq_ = blockIdx.x * (u32)spw + (u32)q;
    const bool ok = sq_ < p.n_stacks;
    s_dyn[oEp + i] = ok ? p.ep_m[(size_t)j * p.n_stacks + sq_] : 0.0;
    s_dyn[oEp + D * spw + i] = ok ? p.ep_s[(size_t)j * p.n_stacks + sq_] : 0.0;
  }
  double x[RM]; /* current sample, rows of this warp */
  int rowo[RM]; /* s_dyn offset of L(j, 0) of this lane's chain, rows of this warp */
#pragma unroll
  for (int m = 0; m < RM; ++m) {
    const int j = g + NW * m;
    x[m] = (j < D) ? p.x[(size_t)j * C + c] : 0.0;
    rowo[m] = oLl + (j * (j + 1) / 2) * 32;
  }
  /* per-chain scalars: warp 0 */
  double energy = 0.0, row_sigma = 1.0, row_beta = 1.0, sigma = 1.0, beta = 1.0, sh_count = 0.0, lg_min_e = 0.0;
  double lt = 0.0, lrs = 0.0, lsig = 0.0; /* -log(row beta), log(row sigma), log(sigma): see the sigma sample below */
  bool sig_fresh = false, rs_new = false, rb_new = false;
  int accepted = 0, swap_type = 0;
  u32 lg_accepts = 0, lg_swaps = 0, lg_att = 0;
  double w3[3] = {0.0, 0.0, 0.0};
  if (g == 0) {
    energy = p.energy[c]; row_sigma = p.row_sigma[c]; row_beta = p.row_beta[c];
    sigma = p.sigma[c]; beta = p.beta[c]; sh_count = p.sh_count[c];
    accepted = p.accepted[c]; swap_type = p.swap_type[c];
    lg_accepts = p.lg_accepts[c]; lg_swaps = p.lg_swaps[c]; lg_att = p.lg_swap_attempts[c];
    lg_min_e = p.lg_min_energy[c];
    lt = -slm_log(row_beta);
    lrs = slm_log(row_sigma);
#pragma unroll
    for (int k = 0; k < 9; ++k) s_dyn[Plan::oPs + k * 32 + lane] = active ? p.p_sig[pidx(p, k, t, s)] : 0.0;
    s_dyn[SIG + lane] = sigma;
    s_dyn[NSC + lane] = p.nscale[c];
    s_dyn[SQ + lane] = 1.0;
    s_dyn[RSQ + lane] = 1.0;
    s_dyn[SCL + lane] = 1.0;
    s_fresh[lane] = (sh_count == p.sh_count0) ? 1 : 0;
    s_acc[lane] = 0;
    s_src[lane] = lane;
  }
  __syncthreads();
  if (g == 0) {
    w3[0] = s_w[3 * t]; w3[1] = s_w[3 * t + 1]; w3[2] = s_w[3 * t + 2];
  }
  const Lik lik(p.payload);
  bool dirty = false;
  mbar_wait(&s_bar, 0); /* the factors are here */

  for (u32 it = 0; it < p.n_rounds; ++it) {
    const u64 r = p.round_begin + it;
    const bool swapRound = (T > 1) && ((r + 2) % (u64)W == 0);
    SLGPU_TSTAMP(0);
    /* ============ the round's draws: normals dealt to the warps by pair index (last warp first: warp 0 also draws
     * the round's uniforms and prepares sqrt(count + 1) and friends for the chains that will accept) ============ */
#pragma unroll 1
    for (int kp = NW - 1 - g; kp < NP; kp += NW) {
      double z0, z1 = 0.0;
      if (replay) {
        const double* zr = p.rp_z + ((size_t)r * C + c) * D;
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
      s_dyn[oZl + (2 * kp) * 32] = z0;
      if (2 * kp + 1 < D) s_dyn[oZl + (2 * kp + 1) * 32] = z1;
    }
    double u_mh = 0.0, u_sw = 0.0;
    if (g == 0) {
      if (replay) {
        u_mh = p.rp_umh[(size_t)r * C + c];
        if (swapRound && t < T - 1) u_sw = p.rp_uswap[(size_t)r * C + c];
      } else {
        double ub;
        philox_draw(p.seed, gchain, r, 1, 0, u_mh, ub);
        if (swapRound && active && t < T - 1) philox_draw(p.seed, gchain, r, 2, 0, u_sw, ub);
      }
      /* ProposalShaper::update's scalars for count + 1 (adaptive.cpp:178-182): they do not depend on the decision, only
       * whether they are used does; nobody reads them between here and the column sweep */
      const double cnt1 = sh_count + 1.0;
      const double sq = sqrt(cnt1);
      s_dyn[SQ + lane] = sq;
      s_dyn[RSQ + lane] = 1.0 / sq;
      s_dyn[SCL + lane] = sqrt((cnt1 - 1.0) / cnt1);
    }
    __syncthreads();
    SLGPU_TSTAMP(1);

    /* ============ GaussianProposal::propose (sampler.cpp:79-93), rows of this warp: a_j = sum_{k<=j} fma(N_jk, z_k, .),
     * k ascending, N = L * nscale + 1e-4 I (adaptive.cpp:200-204).  The columns are walked in RM segments: in segment
     * sgm the rows m >= sgm of this warp are still below their diagonal, so a segment has no per-element guard and its
     * rows' chains run interleaved ============ */
    {
      const double nscale = s_dyn[NSC + lane];
      const double sig = s_dyn[SIG + lane];
      const bool fresh = s_fresh[lane] != 0;
      double a[RM];
#pragma unroll
      for (int m = 0; m < RM; ++m) a[m] = 0.0;
#pragma unroll
      for (int sgm = 0; sgm < RM; ++sgm) {
        const int k0 = sgm == 0 ? 0 : g + NW * (sgm - 1);
        const int k1 = (g + NW * sgm < D) ? g + NW * sgm : k0; /* rows beyond D do not exist */
#pragma unroll 2
        for (int k = k0; k < k1; ++k) {
          const double zk = s_dyn[oZl + k * 32];
          double lv[RM];
#pragma unroll
          for (int m = sgm; m < RM; ++m) lv[m] = (g + NW * m < D) ? s_dyn[rowo[m] + k * 32] : 0.0; /* below the diagonal of every existing row m >= sgm */
#pragma unroll
          for (int m = sgm; m < RM; ++m)
            if (g + NW * m < D) a[m] = fma(lv[m] * nscale, zk, a[m]);
        }
      }
#pragma unroll
      for (int m = 0; m < RM; ++m) {
        const int j = g + NW * m;
        if (j < D) {
          const double njj = fresh ? s_dyn[Plan::oBnd + 3 * D + j] : s_dyn[rowo[m] + j * 32] * nscale + 1e-4;
          const double aj = fma(njj, s_dyn[oZl + j * 32], a[m]);
          s_dyn[oXl + j] = bounce1(x[m] + aj * sig, s_dyn[Plan::oBnd + j], s_dyn[Plan::oBnd + D + j], s_dyn[Plan::oBnd + 2 * D + j]);
        }
      }
    }
    __syncthreads();
    SLGPU_TSTAMP(2);

    /* ============ likelihood: job types dealt to the warps (workerwrapper.hpp:21 contract, one chain per lane) ============ */
    if (active) {
      const double* xp = &s_dyn[oXl];
#pragma unroll 2
      for (int q = g; q < J; q += NW) s_dyn[Plan::oF + q * 32 + lane] = lik((u32)q, xp, (u32)D);
    }
    __syncthreads();
    SLGPU_TSTAMP(3);

    /* ============ warp 0: energy = sum over job types in order (sampler.cpp:148-150), acceptProposal
     * (chainarray.cpp:30-45), append (:94-121) ============ */
    bool acc = false;
    if (g == 0) {
      if (active) {
        double e_new = 0.0;
        for (int q = 0; q < J; ++q) e_new += s_dyn[Plan::oF + q * 32 + lane];
        if (!isinf(e_new)) {
          const double deltaEnergy = e_new - energy;
          acc = u_mh < slm_exp(-1.0 * beta * deltaEnergy);
        }
        if (acc) {
          sh_count += 1.0;
          energy = e_new;
          row_sigma = sigma; /* the row takes the sigma and beta it was proposed with (chainarray.cpp:94-106) */
          rs_new = true;
          if (row_beta != beta) { /* only after the ladder moved: a swap round since the chain's last accept */
            row_beta = beta;
            rb_new = true;
          }
        }
        accepted = acc ? 1 : 0;
        swap_type = 0;
      }
      s_acc[lane] = acc ? 1 : 0;
    }
    const bool any_acc = __syncthreads_or(acc ? 1 : 0) != 0;
    SLGPU_TSTAMP(4);

    /* ============ ProposalShaper::update (adaptive.cpp:173-209) for the accepted chains of the group ============ */
    if (any_acc) {
      dirty = true;
      const double eps = 1e-18;
      const bool accL = s_acc[lane] != 0;
      const double sq = s_dyn[SQ + lane], rsq = s_dyn[RSQ + lane], scale = s_dyn[SCL + lane];
      double xs[RM], fr[RM];
#pragma unroll
      for (int m = 0; m < RM; ++m) {
        const int j = g + NW * m;
        double jump = 0.0;
        if (j < D) {
          const double nx = s_dyn[oXl + j];
          jump = nx - x[m]; /* the accepted jump (sampler.cpp:166) */
          if (accL) x[m] = nx;
        }
        xs[m] = div_by(jump, sq, rsq);
        fr[m] = 0.0;
      }
      /* The head of column j = the row's own diagonal: r = sqrt(V^2 + x_j^2), c = r / V, s = x_j / V with
       * V = max(eps, L_jj * scale).  V, 1/V and V^2 do not depend on the sweep: prepared ahead, off the serial chain. */
      int jn = g, mn_ = 0; /* row and slot of the next head this warp owes */
      double Vn = 1.0, rVn = 1.0, V2n = 1.0;
      auto prep_head = [&]() {
        if (jn < D) {
          const double Lkk = s_dyn[oLl + (jn * (jn + 1) / 2 + jn) * 32] * scale;
          Vn = smax(eps, Lkk);
          rVn = 1.0 / Vn;
          V2n = Vn * Vn;
        }
      };
      auto do_head = [&](double xk) -> double { /* publishes c, s, 1/c of column jn; returns r */
        const double rr = sqrt(V2n + xk * xk);
        const double cc = div_by(rr, Vn, rVn);
        const double ss = div_by(xk, Vn, rVn);
        const double rcc = 1.0 / cc;
        const int o = oHl + (jn & 1) * 96;
        s_dyn[o] = cc;
        s_dyn[o + 32] = ss;
        s_dyn[o + 64] = rcc;
        if (accL) s_dyn[oLl + (jn * (jn + 1) / 2 + jn) * 32] = rr;
        return rr;
      };
      auto sweep_row = [&](int m, int k, double cc, double ss, double rcc) { /* L(j_m, k) and x_j of row slot m */
        const int o = rowo[m] + k * 32;
        const double xj = xs[m];
        double Ljk = s_dyn[o] * scale;
        Ljk = div_by(Ljk + ss * xj, cc, rcc);
        if (accL) s_dyn[o] = Ljk;
        xs[m] = cc * xj - ss * Ljk;
        fr[m] = fma(Ljk, Ljk, fr[m]);
      };
      prep_head();
      if (g == 0) {
        const double rr = do_head(xs[0]);
        fr[0] = fma(rr, rr, fr[0]); /* the diagonal closes the row's sum */
        jn += NW;
        mn_ = 1;
      }
      __syncthreads();
      if (g == 0) prep_head();
      /* column k: its head is in shared memory.  The warp that owns row k+1 runs the serial chain first -- L(k+1,k), x_{k+1},
       * head k+1 -- and hands it over at the barrier; its other rows of column k wait until after the barrier (they need
       * nothing new, and nobody waits for them).  The other warps do their rows of column k while that head is computed. */
      double pc = 0.0, ps = 0.0, prc = 0.0; /* column whose rows this warp still owes, after having been the owner */
      int pk = -1;
#pragma unroll 1
      for (int k = 0; k < D - 1; ++k) {
        if (pk >= 0) { /* the rows I left behind as the owner of the previous head */
#pragma unroll
          for (int m = 0; m < RM; ++m) {
            const int j = g + NW * m;
            if (j > pk + 1 && j < D) sweep_row(m, pk, pc, ps, prc);
          }
          pk = -1;
        }
        const int oh = oHl + (k & 1) * 96;
        const double cc = s_dyn[oh], ss = s_dyn[oh + 32], rcc = s_dyn[oh + 64];
        if (jn == k + 1) { /* I own row k+1: the serial chain runs through me now */
          double xj = xs[0];
          int ro = rowo[0];
#pragma unroll
          for (int m = 1; m < RM; ++m) {
            xj = (m == mn_) ? xs[m] : xj;
            ro = (m == mn_) ? rowo[m] : ro;
          }
          double Ljk = s_dyn[ro + k * 32] * scale;
          Ljk = div_by(Ljk + ss * xj, cc, rcc);
          if (accL) s_dyn[ro + k * 32] = Ljk;
          xj = cc * xj - ss * Ljk;
          const double rr = do_head(xj); /* row k+1 is complete: its head */
#pragma unroll
          for (int m = 0; m < RM; ++m)
            if (m == mn_) fr[m] = fma(rr, rr, fma(Ljk, Ljk, fr[m]));
          pc = cc; ps = ss; prc = rcc; pk = k;
          __syncthreads();
          jn += NW;
          mn_ += 1;
          prep_head(); /* after the hand-over: the divisor of my next head */
        } else {
#pragma unroll
          for (int m = 0; m < RM; ++m) {
            const int j = g + NW * m;
            if (j > k + 1 && j < D) sweep_row(m, k, cc, ss, rcc);
          }
          __syncthreads();
        }
      }
      if (pk >= 0) {
#pragma unroll
        for (int m = 0; m < RM; ++m) {
          const int j = g + NW * m;
          if (j > pk + 1 && j < D) sweep_row(m, pk, pc, ps, prc);
        }
      }
      /* |L|_F: row sums to shared memory (the proposal panel is dead since the jumps were taken, and nobody writes it
       * before the next round's first barrier), the last warp folds them: slot l = row l, halving tree over 32 slots */
#pragma unroll
      for (int m = 0; m < RM; ++m) {
        const int j = g + NW * m;
        if (j < D) s_dyn[oFr + j * 32] = fr[m];
      }
      __syncthreads();
      if (g == NW - 1 && accL) {
        double f[D];
#pragma unroll
        for (int j = 0; j < D; ++j) f[j] = s_dyn[oFr + j * 32];
        int nlive = D;
#pragma unroll
        for (int mm = 16; mm > 0; mm >>= 1) {
#pragma unroll
          for (int l = 0; l < mm; ++l)
            if (l + mm < D && l + mm < nlive) f[l] += f[l + mm];
          if (nlive > mm) nlive = mm;
        }
        const double frob = sqrt(f[0]);
        s_dyn[NSC + lane] = p.prop_norm / smax(frob, eps);
        s_fresh[lane] = 0;
      }
    }
    SLGPU_TSTAMP(5);

    /* ============ swap sweep, hottest pair first: sampler.cpp:169-199,237-257, chainarray.cpp:55-67,162-189 ============ */
    double mid_e = energy;
    int mid_acc = accepted;
    if (swapRound) {
      if (g == 0) {
        int cur = T - 1, my_src = t, mid_src = t;
        bool my_swapped = false;
        for (int tt = T - 2; tt >= 0; --tt) {
          const double El = shfl_d(energy, base + tt);
          const double Eh = shfl_d(energy, base + cur);
          const double bl = shfl_d(beta, base + tt);
          const double bh = shfl_d(beta, base + tt + 1);
          const double u = shfl_d(u_sw, base + tt);
          const double deltaEnergy = El - Eh;
          const double deltaBeta = bl - bh;
          const bool sw = u < slm_exp(deltaEnergy * deltaBeta);
          if (t == tt + 1) my_src = sw ? tt : cur;
          if (t == tt) { my_swapped = sw; mid_src = sw ? cur : tt; }
          if (!sw) cur = tt;
        }
        if (t == 0) my_src = cur;
        const double lb = slm_log(beta);
        const double lbh = shfl_d(lb, lane + 1 < 32 ? lane + 1 : lane);
        mid_e = shfl_d(energy, base + mid_src);
        mid_acc = __shfl_sync(0xffffffffu, accepted, base + mid_src);
        const double new_e = shfl_d(energy, base + my_src);
        const int new_acc = __shfl_sync(0xffffffffu, accepted, base + my_src);
        s_src[lane] = base + my_src;
        energy = new_e;
        accepted = new_acc;
        if (active && t < T - 1) {
          swap_type = my_swapped ? 1 : 2;
          double Pb[9];
#pragma unroll
          for (int k = 0; k < 9; ++k) Pb[k] = p.p_beta[pidx(p, k, t, s)];
          stats_accumulate<1>(Pb, 0.0, 4.0, lb - lbh, lb, my_swapped);
#pragma unroll
          for (int k = 0; k < 9; ++k) p.p_beta[pidx(p, k, t, s)] = Pb[k];
        }
        if (t > 0) beta = s_lad[t]; /* sampler.cpp:186 with the ladder of the last adapt point */
      }
      __syncthreads();
      const int src = s_src[lane]; /* the exchange moves the sample (chainarray.cpp:176-178): every warp its rows */
#pragma unroll
      for (int m = 0; m < RM; ++m) x[m] = shfl_d(x[m], src);
    }

    /* ============ TableLogger counters (logging.cpp:35-48), cold row (db.cpp:18-25), EPSR (diagnostics.cpp:27-47) ============ */
    if (g == 0 && active) {
      /* sigma adaptation sample (sampler.cpp:160-162) of this round's own decision and row.  Sigma and beta of a row do
       * not move in a swap, so it can wait until here, off the column sweep's path (warp 0 owns head 0).  It needs
       * log(row sigma) and -log(row beta), and the next sigma is exp(predict(w, -log(row beta))) (adaptive.cpp:94-106,
       * 134-139): the tier weights w are frozen for the launch and a row's sigma / beta only change on an accept, so the
       * three are kept and recomputed only when their argument changes -- same functions, same arguments, same bits. */
      if (rs_new) {
        lrs = sig_fresh ? lsig : slm_log(row_sigma); /* sigma is still the one the row was proposed with */
        rs_new = false;
      }
      if (rb_new) {
        lt = -slm_log(row_beta);
        sig_fresh = false;
        rb_new = false;
      }
      stats_accumulate<32>(&s_dyn[Plan::oPs + lane], -8.0, 4.0, lrs, lt, acc);
      if (!sig_fresh) { /* first round of the launch (new weights), or -log(row beta) just changed */
        sigma = slm_exp(predict(w3, p.opt_accept, -8.0, 4.0, lt));
        lsig = slm_log(sigma);
        sig_fresh = true;
        s_dyn[SIG + lane] = sigma; /* read after the next round's first barrier */
      }
      lg_min_e = smin(lg_min_e, mid_e);
      lg_accepts += (u32)mid_acc;
      lg_swaps += (swap_type == 1);
      lg_att += (swap_type != 0);
      if (p.round_flags) {
#ifndef SLGPU_STAMPS
        p.round_flags[(size_t)it * C + c] = (unsigned char)((acc ? 1 : 0) | (swap_type << 1));
#endif
      }
      if (it + 1 == p.n_rounds) p.lg_cur_energy[c] = mid_e;
      if (t == 0 && p.rows_out) {
        double* row = p.rows_out + ((size_t)it * p.n_stacks + s) * (D + SLGPU_ROW_EXTRA);
        row[D] = energy;
        row[D + 1] = row_sigma;
        row[D + 2] = row_beta;
        row[D + 3] = (double)accepted;
        row[D + 4] = (double)swap_type;
      }
    }
    if (active && t == 0) {
      const double nn = (double)(r + 1);
      const double rnn = 1.0 / nn;
#pragma unroll
      for (int m = 0; m < RM; ++m) {
        const int j = g + NW * m;
        if (j < D) {
          const int o = oEp + j * spw + sl;
          const double mo = s_dyn[o];
          const double xi = x[m];
          const double newM = mo + div_by(xi - mo, nn, rnn);
          s_dyn[o + D * spw] = s_dyn[o + D * spw] + (xi - mo) * (xi - newM);
          s_dyn[o] = newM;
          if (p.rows_out) p.rows_out[((size_t)it * p.n_stacks + s) * (D + SLGPU_ROW_EXTRA) + j] = xi;
        }
      }
    }
    SLGPU_TSTAMP(6);
  }

  /* ---- state back to global memory; the tile with one bulk copy ---- */
  if (active) {
#pragma unroll
    for (int m = 0; m < RM; ++m) {
      const int j = g + NW * m;
      if (j < D) p.x[(size_t)j * C + c] = x[m];
    }
    if (g == 0) {
      p.energy[c] = energy; p.row_sigma[c] = row_sigma; p.row_beta[c] = row_beta;
      p.sigma[c] = sigma; p.beta[c] = beta; p.sh_count[c] = sh_count;
      p.accepted[c] = accepted; p.swap_type[c] = swap_type;
      p.lg_accepts[c] = lg_accepts; p.lg_swaps[c] = lg_swaps; p.lg_swap_attempts[c] = lg_att;
      p.lg_min_energy[c] = lg_min_e;
#pragma unroll
      for (int k = 0; k < 9; ++k) p.p_sig[pidx(p, k, t, s)] = s_dyn[Plan::oPs + k * 32 + lane];
    }
  }
  fence_async_smem();
  __syncthreads(); /* every warp's last update of the tile, of nscale and of the EPSR state is in shared memory */
  if (g == 0 && active) p.nscale[c] = s_dyn[NSC + lane];
  for (int i = tid; i < D * spw; i += 32 * NW) {
    const int j = i / spw, q = i - j * spw;
    const u32 sq_ = blockIdx.x * (u32)spw + (u32)q;
    if (sq_ < p.n_stacks) {
      p.ep_m[(size_t)j * p.n_stacks + sq_] = s_dyn[oEp + i];
      p.ep_s[(size_t)j * p.n_stacks + sq_] = s_dyn[oEp + D * spw + i];
    }
  }
  if (tid == 0 && dirty) bulk_store_and_wait(gL, &s_dyn[Plan::oL], NL * 32 * (unsigned)sizeof(double));
}

template <class Lik, int D>
inline int launch_step_tile(const slgpu_step_params* p, void* stream) {
  constexpr int NW = TileWarps<D>::value;
  auto kern = pt_step_tile_kernel<Lik, D, NW>;
  const size_t smem = TilePlan<D>::bytes(p->n_job_types, p->n_temps);
  /* the opt-in is per device: set it on every launch (a process may drive several devices) */
  const cudaError_t attr = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (attr != cudaSuccess) return (int)attr;
  const unsigned spw = 32u / p->n_temps;
  kern<<<(p->n_stacks + spw - 1) / spw, 32 * NW, smem, (cudaStream_t)stream>>>(*p);
  return (int)cudaGetLastError();
}

}  // namespace slgpu

#endif
