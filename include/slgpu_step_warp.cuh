/*
 * slgpu_step_warp.cuh -- pt_step for large dimensions (32 < d <= 128): ONE WARP PER CHAIN
 * (included by slgpu_device.cuh).
 *
 * At d = 100 a chain's factor is 40 KB and a stack of 8 is 323 KB: nothing fits on chip, and a config such as
 * BASELINE #4 (1024 stacks x 8) has only 8192 chains -- too few threads for a thread-per-chain kernel.  So
 * the O(d^2) work of one chain is spread over the 32 lanes of a warp and the factors are streamed from HBM:
 *   - lane l owns rows j = l + 32m (m < 4) of x, of the proposal accumulator, of the jump and of L;
 *   - L is stored per chain, packed COLUMN-major (slgpu_l_layout): column k, rows k..d-1, is contiguous, so a
 *     column sweep (proposal GEMV and rank-1 update both walk columns) is one coalesced warp access per row group;
 *   - a row still accumulates k ascending with the diagonal last, the Frobenius norm uses the canonical
 *     32-slot order (slot l = rows l, l+32, ...; halving tree = __shfl_down), so results equal the oracle's bit for bit;
 *   - accept / reject is warp-uniform: the shaper update runs in place without divergence;
 *   - the chain's scalars are replicated over the lanes; normals, proposal and the running jump go through a
 *     per-warp vector in shared memory; the functor evaluates job types lane, lane+32, ... (or all of them with the
 *     whole warp, if it has warp_jobs) and every lane adds the J values in job order;
 *   - a block holds whole stacks (T warps each); the swap sweep of a stack goes through shared memory.
 */
#ifndef SLGPU_STEP_WARP_CUH
#define SLGPU_STEP_WARP_CUH

namespace slgpu {

/* row groups per lane: RM = ceil(d / 32) is a template parameter (2, 3 or 4) so that small d carries no dead rows */
#ifndef SLGPU_WARP_REGS
#define SLGPU_WARP_REGS 64 /* register budget per thread: 65536 / budget threads resident per SM */
#endif
#define SLGPU_WARP_MINBLOCKS(maxt) (65536 / ((maxt) * SLGPU_WARP_REGS))
#ifndef SLGPU_WARP_GEMV_UNROLL
#define SLGPU_WARP_GEMV_UNROLL 4
#endif
constexpr int kWarpGemvUnroll = SLGPU_WARP_GEMV_UNROLL;

/* dynamic shared memory: per warp a vector [d+1] and [J] job values; per stack the swap scratch; per warp a [d]
 * scratch vector when the functor has warp_jobs */
inline size_t warp_smem_bytes(const slgpu_step_params* p, unsigned warps, bool warp_jobs) {
  const size_t d = p->n_dims, J = p->n_job_types, T = p->n_temps, stacks = warps / T;
  return (warps * (d + 1) + warps * J + stacks * (3 * T + T * d) + (warp_jobs ? warps * d : 0)) * sizeof(double) +
         (stacks * T + 1) / 2 * 2 * sizeof(int);
}

/* stacks per block: as many whole stacks as fit 8 warps (at least one) */
inline unsigned warp_stacks_per_block(unsigned n_temps) { return n_temps >= 8 ? 1u : 8u / n_temps; }

/* ---- column sweeps of the warp kernel.  Lp = Lg + lane; off_k = ck - k (ck = k*d - k(k-1)/2), so that L(j,k) with
 * j = lane + 32m is Lp[off_k + 32m] and off_{k+1} = off_k + d - k - 1: a sweep carries the pointer Lp + off_k from column to
 * column and addresses the row groups with immediate offsets.  The sweeps are cut at the row-group
 * boundaries (columns 32G .. 32G+31) so that which row groups take part is known at compile time. ---- */

/* raw column k, rows j >= k of this lane (0 elsewhere); any k.  pc = Lp + off_k */
template <int RM>
__device__ __forceinline__ void warp_load_col(double* dst, const double* pc, int k, int lane, int d) {
#pragma unroll
  for (int m = 0; m < RM; ++m) {
    const int j = lane + 32 * m;
    dst[m] = (32 * m + 31 >= k && j >= k && j < d) ? pc[32 * m] : 0.0;
  }
}

/* proposal GEMV, columns of row group G: a_j = fma(N_jk, z_k, a_j), k ascending.  pc walks the columns (pc = Lp + off_k):
 * the loads of a column are pc[32 m] with immediate offsets.  Columns go in batches of B: all loads of a batch are issued
 * before its first fma (a loop the compiler unrolls keeps one column in flight: its exits sit between the columns), with
 * B sized so that a batch is at most eight loads per lane. */
template <int RM, int G>
__device__ __forceinline__ void warp_gemv_column(const double* pc, int k, const double* vec, const double* s_n0, double nscale, bool fresh,
                                                 int lane, int d, const double* v, double* a) {
  const double zk = vec[k];
  const double n0k = s_n0[k];
#pragma unroll
  for (int m = G; m < RM; ++m) {
    const int j = lane + 32 * m;
    if ((m > G || j >= k) && j < d) {
      double njk = (v ? v[m - G] : pc[32 * m]) * nscale;
      if (m == G && j == k) njk = fresh ? n0k : njk + 1e-4;
      a[m] = fma(njk, zk, a[m]);
    }
  }
}
template <int RM, int G>
__device__ __forceinline__ void warp_gemv_group(const double*& pc, const double* vec, const double* s_n0, double nscale, bool fresh,
                                                int lane, int d, double* a) {
  constexpr int NG = RM - G;
  constexpr int B = NG >= 3 ? 2 : (NG == 2 ? 4 : 8);
  const int k1 = (32 * G + 32 < d) ? 32 * G + 32 : d;
  int k = 32 * G;
  for (; k + B <= k1; k += B) {
    double v[B][NG];
    const double* q = pc;
#pragma unroll
    for (int b = 0; b < B; ++b) {
#pragma unroll
      for (int m = G; m < RM; ++m) {
        const int j = lane + 32 * m;
        v[b][m - G] = ((m > G || j >= k + b) && j < d) ? q[32 * m] : 0.0;
      }
      q += d - (k + b) - 1;
    }
#pragma unroll
    for (int b = 0; b < B; ++b) warp_gemv_column<RM, G>(nullptr, k + b, vec, s_n0, nscale, fresh, lane, d, v[b], a);
    pc = q;
  }
  for (; k < k1; ++k) {
    warp_gemv_column<RM, G>(pc, k, vec, s_n0, nscale, fresh, lane, d, nullptr, a);
    pc += d - k - 1;
  }
  if constexpr (G + 1 < RM) {
    if (k1 < d) warp_gemv_group<RM, G + 1>(pc, vec, s_n0, nscale, fresh, lane, d, a);
  }
}

/* rank-1 sweep of ProposalShaper::update (adaptive.cpp:184-197), columns of row group G.  cur / nx1 hold the raw
 * columns k and k+1 (loaded two columns ahead of their use: a column's loads never depend on earlier columns'
 * stores); the column head (old diagonal, x_k) comes from the owning lane by shuffle, not from memory. */
template <int RM, int G>
__device__ __forceinline__ void warp_shaper_group(double*& p0, double*& p1, double* cur, double* nx1, double scale, int lane, int d,
                                                  double* xs, double* fr) {
  const double eps = 1e-18;
  const int k1 = (32 * G + 32 < d) ? 32 * G + 32 : d;
#pragma unroll 1
  for (int k = 32 * G; k < k1; ++k) {
    double nx2[RM];
    double* const p2 = p1 + (d - k - 2);
    warp_load_col<RM>(nx2, p2, (k + 2 < d) ? k + 2 : 0x7fff, lane, d); /* past the last column: nothing is loaded */
    const double xk = __shfl_sync(0xffffffffu, xs[G], k & 31);
    const double Lkk = __shfl_sync(0xffffffffu, cur[G], k & 31) * scale;
    const double V = smax(eps, Lkk);
    const double rV = 1.0 / V;
    const double rr = sqrt(V * V + xk * xk);
    const double cc = div_by(rr, V, rV);
    const double ss = div_by(xk, V, rV);
    const double rc = 1.0 / cc;
#pragma unroll
    for (int m = G; m < RM; ++m) {
      const int j = lane + 32 * m;
      if (j < d) {
        if (m == G && j == k) {
          p0[32 * m] = rr;
          fr[m] = fma(rr, rr, fr[m]); /* the diagonal closes row k's sum */
        } else if (m > G || j > k) {
          double Ljk = cur[m] * scale;
          Ljk = div_by(Ljk + ss * xs[m], cc, rc);
          p0[32 * m] = Ljk;
          xs[m] = cc * xs[m] - ss * Ljk;
          fr[m] = fma(Ljk, Ljk, fr[m]);
        }
      }
    }
#pragma unroll
    for (int m = 0; m < RM; ++m) {
      cur[m] = nx1[m];
      nx1[m] = nx2[m];
    }
    p0 = p1;
    p1 = p2;
  }
  if constexpr (G + 1 < RM) {
    if (k1 < d) warp_shaper_group<RM, G + 1>(p0, p1, cur, nx1, scale, lane, d, xs, fr);
  }
}

#ifdef SLGPU_WSTAMPS /* development: clock stamps of round SLGPU_WSTAMPS for every 16th chain, 16 x i64 into the round_flags buffer */
#define SLGPU_WSTAMP(k)                                                                     \
  do {                                                                                      \
    if (lane == 0 && active && p.round_flags && it == (SLGPU_WSTAMPS) && (c & 15u) == 0u)   \
      ((long long*)p.round_flags)[(size_t)c + (k)] = clock64();                             \
  } while (0)
#else
#define SLGPU_WSTAMP(k) \
  do {                  \
  } while (0)
#endif

template <class Lik, int RM, int MAXT>
__global__ void __launch_bounds__(MAXT, SLGPU_WARP_MINBLOCKS(MAXT)) pt_step_warp_kernel(const slgpu_step_params p) {
  const int T = (int)p.n_temps, d = (int)p.n_dims, J = (int)p.n_job_types;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwarps = blockDim.x >> 5, spb = nwarps / T;
  const int sl = warp / T, t = warp - sl * T;
  const u32 s = blockIdx.x * (u32)spb + (u32)sl;
  const bool active = s < p.n_stacks; /* warp-uniform */
  const size_t C = (size_t)p.n_stacks * p.n_temps;
  const u32 c = active ? s * (u32)T + (u32)t : 0u;
  const u32 gchain = p.stack_offset * p.n_temps + c;
  const bool replay = (p.rng_mode == SLGPU_RNG_REPLAY);
  const size_t nL = (size_t)d * (d + 1) / 2;

  extern __shared__ double s_dyn[];
  double* const vec = s_dyn + (size_t)warp * (d + 1);                       /* z, then x', then the running jump */
  double* const fv = s_dyn + (size_t)nwarps * (d + 1) + (size_t)warp * J;   /* per-job values */
  double* const sw = s_dyn + (size_t)nwarps * (d + 1) + (size_t)nwarps * J + (size_t)sl * (3 * T + T * d);
  double* const sw_e = sw;           /* [T] energies */
  double* const sw_b = sw + T;       /* [T] betas */
  double* const sw_u = sw + 2 * T;   /* [T] swap uniforms */
  double* const sw_x = sw + 3 * T;   /* [T][d] samples */
  int* const sw_acc = (int*)(s_dyn + (size_t)nwarps * (d + 1) + (size_t)nwarps * J + (size_t)spb * (3 * T + T * d)) + sl * T;
  /* functor scratch (HasWarpJobs): behind the ints, 8-byte aligned */
  double* const wj = s_dyn + (size_t)nwarps * (d + 1) + (size_t)nwarps * J + (size_t)spb * (3 * T + T * d) + (size_t)(spb * T + 1) / 2 +
                     (size_t)warp * d;
  __shared__ double s_mn[SLGPU_MAX_DIMS], s_mx[SLGPU_MAX_DIMS], s_n0[SLGPU_MAX_DIMS], s_rd[SLGPU_MAX_DIMS];
  __shared__ double s_w[SLGPU_MAX_TEMPS * 3], s_lad[SLGPU_MAX_TEMPS];
  for (int i = tid; i < d; i += blockDim.x) {
    s_mn[i] = p.mn[i];
    s_mx[i] = p.mx[i];
    s_n0[i] = p.n0_diag[i];
    s_rd[i] = 1.0 / (p.mx[i] - p.mn[i]);
  }
  for (int i = tid; i < T; i += blockDim.x) {
    s_w[3 * i + 0] = p.sig_w[3 * i + 0];
    s_w[3 * i + 1] = p.sig_w[3 * i + 1];
    s_w[3 * i + 2] = p.sig_w[3 * i + 2];
    s_lad[i] = p.ladder[i];
  }
  __syncthreads();

  const Lik lik(p.payload);
  double* const Lg = p.L + (size_t)c * nL; /* column k starts at ck = k*d - k(k-1)/2; L(j,k) = Lg[ck + j - k] */
  double x[RM];
#pragma unroll
  for (int m = 0; m < RM; ++m) {
    const int j = lane + 32 * m;
    x[m] = (j < d) ? p.x[(size_t)j * C + c] : 0.0;
  }
  double Ps[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) Ps[k] = active ? p.p_sig[pidx(p, k, t, s)] : 0.0;
  double energy = p.energy[c], row_sigma = p.row_sigma[c], row_beta = p.row_beta[c];
  double sigma = p.sigma[c], beta = p.beta[c], nscale = p.nscale[c], sh_count = p.sh_count[c];
  int accepted = p.accepted[c], swap_type = p.swap_type[c];
  u32 lg_accepts = p.lg_accepts[c], lg_swaps = p.lg_swaps[c], lg_att = p.lg_swap_attempts[c];
  double lg_min_e = p.lg_min_energy[c];
  const double w3[3] = {s_w[3 * t], s_w[3 * t + 1], s_w[3 * t + 2]};
  const int W = (int)p.swap_interval;

  for (u32 it = 0; it < p.n_rounds; ++it) {
    const u64 r = p.round_begin + it;
    bool acc = false;
    SLGPU_WSTAMP(0);
    if (active) {
#ifdef SLGPU_WARP_L2_PREFETCH
      for (int e = 16 * lane; e < (int)nL; e += 512) asm volatile("prefetch.global.L2 [%0];" ::"l"(Lg + e));
#endif
      /* ---- normals: lane l makes the Box-Muller pairs l, l+32, ... ---- */
      const double* zr = replay ? p.rp_z + ((size_t)r * C + c) * d : nullptr;
      for (int kp = lane; kp < (d + 1) / 2; kp += 32) {
        double z0, z1 = 0.0;
        if (replay) {
          z0 = zr[2 * kp];
          if (2 * kp + 1 < d) z1 = zr[2 * kp + 1];
        } else {
          double ua, ub, sn, cs;
          philox_draw(p.seed, gchain, r, 0, (u32)kp, ua, ub);
          const double Rr = sqrt(-2.0 * slm_log(ua));
          slm_sincos2pi(ub, &sn, &cs);
          z0 = Rr * cs;
          z1 = Rr * sn;
        }
        vec[2 * kp] = z0;
        if (2 * kp + 1 < d) vec[2 * kp + 1] = z1;
      }
      __syncwarp();
      SLGPU_WSTAMP(1);
      /* ---- proposal: a_j = sum_{k<=j} fma(N_jk, z_k, .), k ascending; one coalesced column segment per row group ---- */
      const bool fresh = (sh_count == p.sh_count0);
      double a[RM];
#pragma unroll
      for (int m = 0; m < RM; ++m) a[m] = 0.0;
      {
        const double* pc = Lg + lane;
        warp_gemv_group<RM, 0>(pc, vec, s_n0, nscale, fresh, lane, d, a);
      }
      __syncwarp(); /* z fully consumed: the vector takes the proposal */
      SLGPU_WSTAMP(2);
#pragma unroll
      for (int m = 0; m < RM; ++m) {
        const int j = lane + 32 * m;
        if (j < d) {
          a[m] = bounce1(x[m] + a[m] * sigma, s_mn[j], s_mx[j], s_rd[j]);
          vec[j] = a[m];
        }
      }
      __syncwarp();
      /* ---- energy: job types lane, lane+32, ...; every lane adds the J values in job order ---- */
      if constexpr (HasWarpJobs<Lik>::value) {
        lik.warp_jobs(vec, (u32)d, (u32)J, lane, fv, wj);
      } else {
        for (int jj = lane; jj < J; jj += 32) fv[jj] = lik((u32)jj, vec, (u32)d);
        __syncwarp();
      }
      double e_new = 0.0;
      for (int jj = 0; jj < J; ++jj) e_new += fv[jj];
      SLGPU_WSTAMP(3);
      /* ---- acceptProposal (chainarray.cpp:30-45), append (:94-121) ---- */
      if (!isinf(e_new)) {
        double u;
        if (replay) u = p.rp_umh[(size_t)r * C + c];
        else { double ub; philox_draw(p.seed, gchain, r, 1, 0, u, ub); }
        const double deltaEnergy = e_new - energy;
        acc = u < slm_exp(-1.0 * beta * deltaEnergy);
      }
      double xs[RM]; /* the accepted jump, then the running rank-1 vector */
      if (acc) {
#pragma unroll
        for (int m = 0; m < RM; ++m) {
          xs[m] = a[m] - x[m];
          if (lane + 32 * m < d) x[m] = a[m];
        }
        energy = e_new;
        row_sigma = sigma;
        row_beta = beta;
      }
      accepted = acc ? 1 : 0;
      swap_type = 0;
      const double logtemper = -slm_log(row_beta);
      stats_accumulate<1>(Ps, -8.0, 4.0, slm_log(row_sigma), logtemper, acc);
      sigma = slm_exp(predict(w3, p.opt_accept, -8.0, 4.0, logtemper));
      /* ---- ProposalShaper::update (adaptive.cpp:173-209), in place: accept is warp-uniform ---- */
      SLGPU_WSTAMP(4);
      if (acc) {
        const double eps = 1e-18;
        sh_count += 1.0;
        const double sq = sqrt(sh_count);
        const double rsq = 1.0 / sq;
        const double scale = sqrt((sh_count - 1.0) / sh_count);
        double fr[RM];
#pragma unroll
        for (int m = 0; m < RM; ++m) {
          xs[m] = div_by(xs[m], sq, rsq);
          fr[m] = 0.0;
        }
        {
          double cur[RM], nx1[RM];
          double* p0 = Lg + lane;    /* column k:     p0[32 m] = L(lane + 32 m, k) */
          double* p1 = p0 + (d - 1); /* column k + 1 */
          warp_load_col<RM>(cur, p0, 0, lane, d);
          warp_load_col<RM>(nx1, p1, 1 < d ? 1 : 0x7fff, lane, d);
          warp_shaper_group<RM, 0>(p0, p1, cur, nx1, scale, lane, d, xs, fr);
        }
        /* |L|_F: slot l = rows l, l+32, ... in order; halving tree over the 32 slots */
        double slot = 0.0;
#pragma unroll
        for (int m = 0; m < RM; ++m)
          if (lane + 32 * m < d) slot += fr[m];
#pragma unroll
        for (int mm = 16; mm > 0; mm >>= 1) {
          const double other = __shfl_down_sync(0xffffffffu, slot, mm);
          if (lane < mm) slot += other;
        }
        const double frob = sqrt(__shfl_sync(0xffffffffu, slot, 0));
        nscale = p.prop_norm / smax(frob, eps);
      }
    }

    SLGPU_WSTAMP(5);
    /* ================= swap sweep of the stack, through shared memory ================= */
    double mid_e = energy;
    int mid_acc = accepted;
    const bool swapRound = (T > 1) && ((r + 2) % (u64)W == 0); /* block-uniform */
    if (swapRound) {
      if (lane == 0) {
        double u_sw = 0.0;
        if (active && t < T - 1) {
          if (replay) u_sw = p.rp_uswap[(size_t)r * C + c];
          else { double ub; philox_draw(p.seed, gchain, r, 2, 0, u_sw, ub); }
        }
        sw_e[t] = energy;
        sw_b[t] = beta;
        sw_u[t] = u_sw;
        sw_acc[t] = accepted;
      }
#pragma unroll
      for (int m = 0; m < RM; ++m) {
        const int j = lane + 32 * m;
        if (j < d) sw_x[(size_t)t * d + j] = x[m];
      }
      __syncthreads();
      int cur = T - 1, my_src = t, mid_src = t;
      bool my_swapped = false;
      for (int tt = T - 2; tt >= 0; --tt) {
        const double deltaEnergy = sw_e[tt] - sw_e[cur];
        const double deltaBeta = sw_b[tt] - sw_b[tt + 1];
        const bool swp = sw_u[tt] < slm_exp(deltaEnergy * deltaBeta);
        if (t == tt + 1) my_src = swp ? tt : cur;
        if (t == tt) { my_swapped = swp; mid_src = swp ? cur : tt; }
        if (!swp) cur = tt;
      }
      if (t == 0) my_src = cur;
      const double lb = slm_log(beta);
      const double lbh = (t + 1 < T) ? slm_log(sw_b[t + 1]) : lb;
      mid_e = sw_e[mid_src];
      mid_acc = sw_acc[mid_src];
      energy = sw_e[my_src];
      accepted = sw_acc[my_src];
#pragma unroll
      for (int m = 0; m < RM; ++m) {
        const int j = lane + 32 * m;
        if (j < d) x[m] = sw_x[(size_t)my_src * d + j];
      }
      if (active && t < T - 1) {
        swap_type = my_swapped ? 1 : 2;
        if (lane == 0) {
          double Pb[9];
#pragma unroll
          for (int k = 0; k < 9; ++k) Pb[k] = p.p_beta[pidx(p, k, t, s)];
          stats_accumulate<1>(Pb, 0.0, 4.0, lb - lbh, lb, my_swapped);
#pragma unroll
          for (int k = 0; k < 9; ++k) p.p_beta[pidx(p, k, t, s)] = Pb[k];
        }
      }
      if (t > 0) beta = s_lad[t];
      __syncthreads(); /* the scratch is reused next swap round */
    }

    SLGPU_WSTAMP(6);
    if (active) {
      lg_min_e = smin(lg_min_e, mid_e);
      lg_accepts += (u32)mid_acc;
      lg_swaps += (swap_type == 1);
      lg_att += (swap_type != 0);
#ifndef SLGPU_WSTAMPS
      if (p.round_flags) p.round_flags[(size_t)it * C + c] = (unsigned char)((acc ? 1 : 0) | (swap_type << 1));
#else
      if (lane == 0 && p.round_flags && it == (SLGPU_WSTAMPS) && (c & 15u) == 0u) ((long long*)p.round_flags)[(size_t)c + 7] = acc ? 1 : 0;
#endif
      if (it + 1 == p.n_rounds) p.lg_cur_energy[c] = mid_e; /* TableLogger energies_ (logging.cpp:43) */
      if (t == 0) {
        const double nn = (double)(r + 1);
        const double rnn = 1.0 / nn;
        double* row = p.rows_out ? p.rows_out + ((size_t)it * p.n_stacks + s) * (d + SLGPU_ROW_EXTRA) : nullptr;
#pragma unroll
        for (int m = 0; m < RM; ++m) {
          const int j = lane + 32 * m;
          if (j < d) {
            const size_t o = (size_t)j * p.n_stacks + s;
            const double mo = p.ep_m[o];
            const double xi = x[m];
            const double newM = mo + div_by(xi - mo, nn, rnn);
            p.ep_s[o] = p.ep_s[o] + (xi - mo) * (xi - newM);
            p.ep_m[o] = newM;
            if (row) row[j] = xi;
          }
        }
        if (row && lane == 0) {
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
    for (int m = 0; m < RM; ++m) {
      const int j = lane + 32 * m;
      if (j < d) p.x[(size_t)j * C + c] = x[m];
    }
    if (lane == 0) {
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
      for (int k = 0; k < 9; ++k) p.p_sig[pidx(p, k, t, s)] = Ps[k];
    }
  }
}

}  // namespace slgpu

#endif
