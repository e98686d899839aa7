/*
 * slgpu_step_pair.cuh -- pt_step for a compile-time dimension D <= 32 with TWO lanes per chain
 * (included by slgpu_device.cuh).
 *
 * Same rounds, same arithmetic in the same order as pt_step_kernel / pt_step_fast_kernel and the oracle.
 * Why two lanes per chain: the step is latency-bound, and the number of chains an SM can keep resident is
 * capped by the L2 footprint of their L factors (more than ~256 chains per SM and the factors start
 * thrashing between L2 and HBM).  Splitting every chain's ROWS over a lane pair doubles the resident warps
 * at the same footprint and halves the per-thread register arrays:
 *   lane g in {0,1} of a pair owns rows j = 2m + g of x, of the proposal accumulator, of the jump and of L;
 *   a row's accumulation order (k ascending, diagonal last) is untouched, so results do not change;
 *   scalars of the chain (energy, sigma, beta, counters ...) are replicated in both lanes -- both run the
 *   same deterministic instruction sequence on the same inputs, so they stay bit-identical;
 *   what one lane needs from the other travels by shuffle (x_k for the rank-1 heads, partial energies,
 *   row sums) or through the [row][chain] panels in shared memory (normals, proposal, samples).
 * All column loops are fully unrolled over column pairs (only 10 rows per lane at d=20), every guard that
 * depends on the lane is a predicate, loads of a column pair are issued together.
 * ProposalShaper updates are compacted over the block as in pt_step_fast_kernel, one lane pair per staged chain.
 * A warp holds 16 chains = floor(16/T) stacks, so nTemperatures <= 16 is required (else the caller falls back).
 */
#ifndef SLGPU_STEP_PAIR_CUH
#define SLGPU_STEP_PAIR_CUH

namespace slgpu {

#ifndef SLGPU_PAIR_THREADS
#define SLGPU_PAIR_THREADS 256
#endif
#ifndef SLGPU_PAIR_MINBLOCKS
#define SLGPU_PAIR_MINBLOCKS 2
#endif

struct PairMap {
  int T, lane, g, q, sl, t, baseq, spw, cb;
  u32 s, c;
  bool active;
  __device__ PairMap(const slgpu_step_params& p) {
    T = (int)p.n_temps;
    spw = 16 / T;
    lane = threadIdx.x & 31;
    g = lane & 1;
    q = lane >> 1; /* chain slot within the warp */
    sl = q / T;
    t = q - sl * T;
    const u32 gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    s = gw * (u32)spw + (u32)sl;
    active = (sl < spw) && (s < p.n_stacks);
    if (sl >= spw) {
      baseq = q;
      t = 0;
    } else {
      baseq = sl * T;
    }
    c = s * (u32)T + (u32)t;
    cb = threadIdx.x >> 1; /* chain slot within the block */
  }
};

/* packed index of row j = 2m + g at column 0: tri(2m+g) = 2m^2 + m + g(2m+1) */
__device__ __forceinline__ int pair_tri(int m, int g) { return 2 * m * m + m + g * (2 * m + 1); }

/* ---- proposal GEMV, column pair (K0, K0+1) with K0 = 2*M0; rows m >= M0 of this lane ---- */
template <int D, int M0>
__device__ __forceinline__ void gemv_pair(const double* Lc, const double* zs, int zst, double nscale, bool fresh, const double* s_n0,
                                          int g, double* a) {
  constexpr int R = (D + 1) / 2;
  constexpr int K0 = 2 * M0;
  constexpr bool two = (K0 + 1 < D);
  double col0[R], col1[R];
#pragma unroll
  for (int m = M0; m < R; ++m) {
    const int o = (pair_tri(m, g) + K0) * kLs;
    col0[m] = Lc[o];
    if (two) col1[m] = Lc[o + kLs];
  }
  const double z0 = zs[K0 * zst];
  const double z1 = two ? zs[(K0 + 1) * zst] : 0.0;
#pragma unroll
  for (int m = M0; m < R; ++m) {
    if (2 * m + g < D) {
      /* column K0: every row m >= M0 has j >= K0; the diagonal is (m == M0, g == 0) */
      double n0 = col0[m] * nscale;
      if (m == M0 && g == 0) n0 = fresh ? s_n0[K0] : n0 + 1e-4;
      a[m] = fma(n0, z0, a[m]);
      /* column K0+1: rows with j >= K0+1, i.e. m > M0 or (m == M0, g == 1) which is the diagonal */
      if (two && (m > M0 || g == 1)) {
        double n1 = col1[m] * nscale;
        if (m == M0) n1 = fresh ? s_n0[K0 + 1] : n1 + 1e-4;
        a[m] = fma(n1, z1, a[m]);
      }
    }
  }
  if constexpr (M0 + 1 < R) gemv_pair<D, M0 + 1>(Lc, zs, zst, nscale, fresh, s_n0, g, a);
}

/* ---- one column k of the rank-1 sweep (src/infer/adaptive.cpp:184-197); k static, rows split over the pair ---- */
template <int D, int K>
__device__ __forceinline__ void shaper_pair_col(double* Lc, const double* col, double scale, int g, int lane, unsigned pm, double* xs, double* fr) {
  constexpr int R = (D + 1) / 2;
  constexpr int MK = K / 2;      /* row K is row MK of lane (K & 1) */
  constexpr int GK = K & 1;
  const double eps = 1e-18;
  const int owner = (lane & ~1) | GK;
  const double xk = __shfl_sync(pm, xs[MK], owner);
  const double Lkk = __shfl_sync(pm, col[MK], owner) * scale;
  const double V = smax(eps, Lkk);
  const double rV = 1.0 / V;
  const double r = sqrt(V * V + xk * xk);
  const double c = div_by(r, V, rV);
  const double s = div_by(xk, V, rV);
  const double rc = 1.0 / c;
  if (g == GK) {
    Lc[(pair_tri(MK, GK) + K) * kLs] = r;
    fr[MK] = fma(r, r, fr[MK]); /* the diagonal closes row K's sum */
  }
#pragma unroll
  for (int m = MK; m < R; ++m) {
    const int j = 2 * m + g;
    if (j > K && j < D) {
      double Ljk = col[m] * scale;
      Ljk = div_by(Ljk + s * xs[m], c, rc);
      Lc[(pair_tri(m, g) + K) * kLs] = Ljk;
      xs[m] = c * xs[m] - s * Ljk;
      fr[m] = fma(Ljk, Ljk, fr[m]);
    }
  }
}

template <int D, int M0>
__device__ __forceinline__ void shaper_pair(double* Lc, double scale, int g, int lane, unsigned pm, double* xs, double* fr) {
  constexpr int R = (D + 1) / 2;
  constexpr int K0 = 2 * M0;
  constexpr bool two = (K0 + 1 < D);
  double col0[R], col1[R];
#pragma unroll
  for (int m = M0; m < R; ++m) { /* both columns' loads first: column K0's stores never touch column K0+1 */
    const int o = (pair_tri(m, g) + K0) * kLs;
    col0[m] = Lc[o];
    if (two) col1[m] = Lc[o + kLs];
  }
  shaper_pair_col<D, K0>(Lc, col0, scale, g, lane, pm, xs, fr);
  if constexpr (two) shaper_pair_col<D, K0 + 1>(Lc, col1, scale, g, lane, pm, xs, fr);
  if constexpr (M0 + 1 < R) shaper_pair<D, M0 + 1>(Lc, scale, g, lane, pm, xs, fr);
}

/* ProposalShaper::update (src/infer/adaptive.cpp:173-209) by a lane pair; jump[j*jst] holds the accepted jump. */
template <int D>
__device__ __forceinline__ double shaper_update_pair(double* Lc, const double* jump, int jst, double count, double prop_norm, int g,
                                                     int lane) {
  constexpr int R = (D + 1) / 2;
  const double eps = 1e-18;
  const double sq = sqrt(count);
  const double rsq = 1.0 / sq;
  const double scale = sqrt((count - 1.0) / count);
  double xs[R], fr[R];
#pragma unroll
  for (int m = 0; m < R; ++m) {
    const int j = 2 * m + g;
    xs[m] = (j < D) ? div_by(jump[j * jst], sq, rsq) : 0.0;
    fr[m] = 0.0;
  }
  const unsigned pm = 3u << (lane & ~1); /* the two lanes of this pair */
  shaper_pair<D, 0>(Lc, scale, g, lane, pm, xs, fr);
  /* |L|_F: both lanes assemble all D row sums and run the canonical 32-slot halving tree */
  double f[2 * R];
#pragma unroll
  for (int m = 0; m < R; ++m) {
    const double other = __shfl_xor_sync(pm, fr[m], 1);
    f[2 * m] = g ? other : fr[m];
    f[2 * m + 1] = g ? fr[m] : other;
  }
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

/* dynamic shared memory of pt_step_pair_kernel: three [D][NCH] panels, the proposal buffer, the EPSR state of the block's stacks */
template <int D, int NT>
inline size_t pair_smem_bytes(unsigned n_temps) {
  const size_t nch = NT / 2;
  const size_t nsb = (size_t)(NT / 32) * (16u / n_temps);
  return ((size_t)3 * D * nch + (size_t)(D + 1) * nch + (size_t)2 * D * nsb) * sizeof(double);
}

template <class Lik, int D, int NT>
__global__ void __launch_bounds__(NT, SLGPU_PAIR_MINBLOCKS) pt_step_pair_kernel(const slgpu_step_params p) {
  constexpr int R = (D + 1) / 2;
  constexpr int NCH = NT / 2;
  constexpr int NLd = D * (D + 1) / 2;
  const PairMap m(p);
  const int T = m.T, g = m.g, cb = m.cb, lane = m.lane;
  const size_t C = (size_t)p.n_stacks * p.n_temps;
  const bool active = m.active;
  const u32 c = active ? m.c : 0u;
  const u32 gchain = p.stack_offset * p.n_temps + c;
  const bool replay = (p.rng_mode == SLGPU_RNG_REPLAY);
  const int tid = threadIdx.x;

  extern __shared__ double s_dyn[];
  double* const s_jump = s_dyn;               /* staged accepted jumps, [j][slot] */
  double* const s_z = s_dyn + D * NCH;        /* normals, then the proposal x', [k][chain] */
  double* const s_x = s_dyn + 2 * D * NCH;    /* current sample, [j][chain] */
  double* const s_xp = s_dyn + 3 * D * NCH;   /* proposal x', contiguous per chain with an odd stride: what the functor reads */
  double* const xp = s_xp + cb * (D + 1);
  const int nsb = (NT / 32) * m.spw;
  double* const s_epm = s_dyn + 3 * D * NCH + (D + 1) * NCH;
  double* const s_eps = s_epm + D * nsb;
  const int sb = (tid >> 5) * m.spw + m.sl;
  __shared__ double s_mn[D], s_mx[D], s_n0[D], s_rd[D];
  __shared__ double s_w[SLGPU_MAX_TEMPS * 3], s_lad[SLGPU_MAX_TEMPS];
  __shared__ double s_cnt[NCH], s_nsc[NCH];
  __shared__ u32 s_chain[NCH], s_owner[NCH];
  __shared__ int s_nacc[2];
  for (int i = tid; i < D; i += NT) {
    s_mn[i] = p.mn[i];
    s_mx[i] = p.mx[i];
    s_n0[i] = p.n0_diag[i];
    s_rd[i] = 1.0 / (p.mx[i] - p.mn[i]);
  }
  for (int i = tid; i < T; i += NT) {
    s_w[3 * i + 0] = p.sig_w[3 * i + 0];
    s_w[3 * i + 1] = p.sig_w[3 * i + 1];
    s_w[3 * i + 2] = p.sig_w[3 * i + 2];
    s_lad[i] = p.ladder[i];
  }
  if (tid < 2) s_nacc[tid] = 0;

  const Lik lik(p.payload);
  const double* Lc = l_base(p.L, c, NLd);
#pragma unroll
  for (int mm = 0; mm < R; ++mm) {
    const int j = 2 * mm + g;
    if (j < D) s_x[j * NCH + cb] = p.x[(size_t)j * C + c];
  }
  if (active && m.t == 0) {
#pragma unroll
    for (int mm = 0; mm < R; ++mm) {
      const int j = 2 * mm + g;
      if (j < D) {
        s_epm[j * nsb + sb] = p.ep_m[(size_t)j * p.n_stacks + m.s];
        s_eps[j * nsb + sb] = p.ep_s[(size_t)j * p.n_stacks + m.s];
      }
    }
  }
  __syncthreads();
  double Ps[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) Ps[k] = active ? p.p_sig[pidx(p, k, m.t, m.s)] : 0.0;
  double energy = p.energy[c], row_sigma = p.row_sigma[c], row_beta = p.row_beta[c];
  double sigma = p.sigma[c], beta = p.beta[c], nscale = p.nscale[c], sh_count = p.sh_count[c];
  int accepted = p.accepted[c], swap_type = p.swap_type[c];
  u32 lg_accepts = p.lg_accepts[c], lg_swaps = p.lg_swaps[c], lg_att = p.lg_swap_attempts[c];
  double lg_min_e = p.lg_min_energy[c];
  const double w3[3] = {s_w[3 * m.t], s_w[3 * m.t + 1], s_w[3 * m.t + 2]};
  const int W = (int)p.swap_interval;

  for (u32 it = 0; it < p.n_rounds; ++it) {
    const u64 r = p.round_begin + it;
    const int par = (int)(it & 1u);
    bool acc = false;
    /* ================= phase A ================= */
    /* normals: lane g makes the Box-Muller pairs kp = g, g+2, ... of its chain */
    if (active) {
      const double* zr = replay ? p.rp_z + ((size_t)r * C + c) * D : nullptr;
#pragma unroll 1
      for (int kp = g; kp < (D + 1) / 2; kp += 2) {
        double z0, z1 = 0.0;
        if (replay) {
          z0 = zr[2 * kp];
          if (2 * kp + 1 < D) z1 = zr[2 * kp + 1];
        } else {
          double ua, ub, sn, cs;
          philox_draw(p.seed, gchain, r, 0, (u32)kp, ua, ub);
          const double Rr = sqrt(-2.0 * slm_log(ua));
          slm_sincos2pi(ub, &sn, &cs);
          z0 = Rr * cs;
          z1 = Rr * sn;
        }
        s_z[(2 * kp) * NCH + cb] = z0;
        if (2 * kp + 1 < D) s_z[(2 * kp + 1) * NCH + cb] = z1;
      }
    }
    __syncwarp();
    double a[R];
    {
      const bool fresh = (sh_count == p.sh_count0);
#pragma unroll
      for (int mm = 0; mm < R; ++mm) a[mm] = 0.0;
      if (active) gemv_pair<D, 0>(Lc, s_z + cb, NCH, nscale, fresh, s_n0, g, a);
    }
    if (active) {
#pragma unroll
      for (int mm = 0; mm < R; ++mm) {
        const int j = 2 * mm + g;
        if (j < D) {
          a[mm] = bounce1(s_x[j * NCH + cb] + a[mm] * sigma, s_mn[j], s_mx[j], s_rd[j]);
          xp[j] = a[mm];
        }
      }
    }
    __syncwarp(); /* the whole proposal of the chain is visible to both lanes */
    /* energy: lane g evaluates the job types g, g+2, ... on the chain's proposal; both lanes then add all
     * J values in job-type order (sampler.cpp:148-150) */
    double e_new = 0.0;
    {
      const u32 J = (StaticJobs<Lik, D>::value > 0) ? (u32)StaticJobs<Lik, D>::value : p.n_job_types;
#pragma unroll 2
      for (u32 jj = 0; jj < J; jj += 2) {
        double fmine = 0.0;
        if (active && jj + (u32)g < J) fmine = lik(jj + (u32)g, xp, (u32)D);
        const double fother = __shfl_xor_sync(0xffffffffu, fmine, 1);
        e_new += g ? fother : fmine;                 /* job jj   */
        if (jj + 1 < J) e_new += g ? fmine : fother; /* job jj+1 */
      }
    }
    if (active) {
      if (!isinf(e_new)) {
        double u;
        if (replay) u = p.rp_umh[(size_t)r * C + c];
        else { double ub; philox_draw(p.seed, gchain, r, 1, 0, u, ub); }
        const double deltaEnergy = e_new - energy;
        acc = u < slm_exp(-1.0 * beta * deltaEnergy);
      }
    }
    /* staging slot: lane 0 of the pair draws it, the partner copies it */
    int slot = 0;
    if (acc && g == 0) slot = atomicAdd(&s_nacc[par], 1);
    slot = __shfl_sync(0xffffffffu, slot, lane & ~1);
    if (acc) {
      sh_count += 1.0;
#pragma unroll
      for (int mm = 0; mm < R; ++mm) {
        const int j = 2 * mm + g;
        if (j < D) {
          s_jump[j * NCH + slot] = a[mm] - s_x[j * NCH + cb];
          s_x[j * NCH + cb] = a[mm];
        }
      }
      if (g == 0) {
        s_cnt[slot] = sh_count;
        s_chain[slot] = c;
        s_owner[slot] = (u32)cb;
      }
      energy = e_new;
      row_sigma = sigma;
      row_beta = beta;
    }
    if (active) {
      accepted = acc ? 1 : 0;
      swap_type = 0;
      const double logtemper = -slm_log(row_beta);
      stats_accumulate<1>(Ps, -8.0, 4.0, slm_log(row_sigma), logtemper, acc);
      sigma = slm_exp(predict(w3, p.opt_accept, -8.0, 4.0, logtemper));
    }
    __syncthreads();
    /* ================= phase B: compacted ProposalShaper updates, one lane pair per staged chain ================= */
    {
      const int nacc = s_nacc[par];
      /* pair -> slot map: staged chains dealt round-robin to the warps */
      const int pw = lane >> 1;                          /* pair within warp, 0..15 */
      const int bslot = pw * (NT / 32) + (tid >> 5);
      const bool mine = bslot < nacc;
      if (mine) { /* pair-uniform; the shuffles inside are masked to the pair */
        const double ns = shaper_update_pair<D>(l_base(p.L, s_chain[bslot], NLd), s_jump + bslot, NCH, s_cnt[bslot], p.prop_norm, g, lane);
        if (g == 0) s_nsc[s_owner[bslot]] = ns;
      }
      if (tid == 0) s_nacc[par ^ 1] = 0;
    }
    __syncthreads();
    if (acc) nscale = s_nsc[cb];

    /* ================= swap sweep ================= */
    double mid_e = energy;
    int mid_acc = accepted;
    const bool swapRound = (T > 1) && ((r + 2) % (u64)W == 0);
    if (swapRound) {
      double u_sw = 0.0;
      if (active && m.t < T - 1) {
        if (replay) u_sw = p.rp_uswap[(size_t)r * C + c];
        else { double ub; philox_draw(p.seed, gchain, r, 2, 0, u_sw, ub); }
      }
      /* chain slot q of this warp lives in lanes 2q, 2q+1 (replicated scalars): read lane 2q+g */
      int cur = T - 1, my_src = m.t, mid_src = m.t;
      bool my_swapped = false;
      for (int tt = T - 2; tt >= 0; --tt) {
        const double El = shfl_d(energy, 2 * (m.baseq + tt) + g);
        const double Eh = shfl_d(energy, 2 * (m.baseq + cur) + g);
        const double bl = shfl_d(beta, 2 * (m.baseq + tt) + g);
        const double bh = shfl_d(beta, 2 * (m.baseq + tt + 1) + g);
        const double u = shfl_d(u_sw, 2 * (m.baseq + tt) + g);
        const double deltaEnergy = El - Eh;
        const double deltaBeta = bl - bh;
        const bool sw = u < slm_exp(deltaEnergy * deltaBeta);
        if (m.t == tt + 1) my_src = sw ? tt : cur;
        if (m.t == tt) { my_swapped = sw; mid_src = sw ? cur : tt; }
        if (!sw) cur = tt;
      }
      if (m.t == 0) my_src = cur;
      const double lb = slm_log(beta);
      const double lbh = shfl_d(lb, lane + 2 < 32 ? lane + 2 : lane);
      mid_e = shfl_d(energy, 2 * (m.baseq + mid_src) + g);
      mid_acc = __shfl_sync(0xffffffffu, accepted, 2 * (m.baseq + mid_src) + g);
      const double new_e = shfl_d(energy, 2 * (m.baseq + my_src) + g);
      const int new_acc = __shfl_sync(0xffffffffu, accepted, 2 * (m.baseq + my_src) + g);
      { /* gather this lane's rows of the sample from the source chain's column of s_x (same warp) */
        const int src_cb = (cb & ~15) + ((m.baseq + my_src) & 15);
        double gx[R];
#pragma unroll
        for (int mm = 0; mm < R; ++mm) {
          const int j = 2 * mm + g;
          gx[mm] = (j < D) ? s_x[j * NCH + src_cb] : 0.0;
        }
        __syncwarp();
#pragma unroll
        for (int mm = 0; mm < R; ++mm) {
          const int j = 2 * mm + g;
          if (j < D) s_x[j * NCH + cb] = gx[mm];
        }
        __syncwarp();
      }
      energy = new_e;
      accepted = new_acc;
      if (active && m.t < T - 1) {
        swap_type = my_swapped ? 1 : 2;
        if (g == 0) {
          double Pb[9];
#pragma unroll
          for (int k = 0; k < 9; ++k) Pb[k] = p.p_beta[pidx(p, k, m.t, m.s)];
          stats_accumulate<1>(Pb, 0.0, 4.0, lb - lbh, lb, my_swapped);
#pragma unroll
          for (int k = 0; k < 9; ++k) p.p_beta[pidx(p, k, m.t, m.s)] = Pb[k];
        }
      }
      if (m.t > 0) beta = s_lad[m.t];
    }

    if (active) {
      lg_min_e = smin(lg_min_e, mid_e);
      lg_accepts += (u32)mid_acc;
      lg_swaps += (swap_type == 1);
      lg_att += (swap_type != 0);
      if (p.round_flags) p.round_flags[(size_t)it * C + c] = (unsigned char)((acc ? 1 : 0) | (swap_type << 1));
      if (it + 1 == p.n_rounds) p.lg_cur_energy[c] = mid_e; /* TableLogger energies_ (logging.cpp:43) */
      if (m.t == 0) {
        const double nn = (double)(r + 1);
        const double rnn = 1.0 / nn;
        double* row = p.rows_out ? p.rows_out + ((size_t)it * p.n_stacks + m.s) * (D + SLGPU_ROW_EXTRA) : nullptr;
#pragma unroll
        for (int mm = 0; mm < R; ++mm) {
          const int j = 2 * mm + g;
          if (j < D) {
            const int o = j * nsb + sb;
            const double mo = s_epm[o];
            const double xi = s_x[j * NCH + cb];
            const double newM = mo + div_by(xi - mo, nn, rnn);
            s_eps[o] = s_eps[o] + (xi - mo) * (xi - newM);
            s_epm[o] = newM;
            if (row) row[j] = xi;
          }
        }
        if (row && g == 0) {
          row[D] = energy;
          row[D + 1] = row_sigma;
          row[D + 2] = row_beta;
          row[D + 3] = (double)accepted;
          row[D + 4] = (double)swap_type;
        }
      }
    }
  }

  if (active) {
#pragma unroll
    for (int mm = 0; mm < R; ++mm) {
      const int j = 2 * mm + g;
      if (j < D) {
        p.x[(size_t)j * C + c] = s_x[j * NCH + cb];
        if (m.t == 0) {
          p.ep_m[(size_t)j * p.n_stacks + m.s] = s_epm[j * nsb + sb];
          p.ep_s[(size_t)j * p.n_stacks + m.s] = s_eps[j * nsb + sb];
        }
      }
    }
    if (g == 0) {
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
      for (int k = 0; k < 9; ++k) p.p_sig[pidx(p, k, m.t, m.s)] = Ps[k];
    }
  }
}

}  // namespace slgpu

#endif
