/*
 * slgpu_step_quad.cuh -- pt_step for a compile-time dimension D <= 32 with the block's L factors
 * RESIDENT IN SHARED MEMORY for the whole launch and FOUR lanes per chain (included by slgpu_device.cuh).
 *
 * Same rounds, same arithmetic in the same order as the other step kernels and the oracle.  Layout of the work:
 *   - a block owns NCH = NT/4 chains; their packed L factors (d(d+1)/2 doubles each) are copied from the tiled
 *     HBM store into shared memory once per launch (coalesced), used for all rounds of the launch, and
 *     written back once: HBM sees each factor twice per launch instead of being streamed every round, and
 *     no global load is left inside the round loop;
 *   - shared memory then caps the resident chains at ~96 per SM, so every chain is spread over a QUAD of
 *     lanes (lane g owns rows j = 4m + g of x, of the proposal accumulator, of the jump and of L), which
 *     keeps 12 warps per SM resident; a row's accumulation order is untouched, so results do not change;
 *     the chain's scalars are replicated in the four lanes (same deterministic instruction sequence on the
 *     same inputs); x_k for the rank-1 heads and the row sums travel by quad-masked shuffles, normals /
 *     proposal / partial energies through per-chain vectors in shared memory;
 *   - every column loop is fully unrolled (5 rows per lane at d = 20), lane-dependent guards are predicates;
 *   - ProposalShaper updates are compacted over the block: accepted chains stage their jump, and after a
 *     block barrier the staged chains are dealt to the quads of all warps (any quad can update any chain:
 *     the factor is in shared memory).
 * A warp holds 8 chains = floor(8/T) stacks, so nTemperatures <= 8 is required (else the caller falls back).
 */
#ifndef SLGPU_STEP_QUAD_CUH
#define SLGPU_STEP_QUAD_CUH

namespace slgpu {

#ifndef SLGPU_QUAD_THREADS
#define SLGPU_QUAD_THREADS 192
#endif
#ifndef SLGPU_QUAD_MINBLOCKS
#define SLGPU_QUAD_MINBLOCKS 2
#endif

struct QuadMap {
  int T, lane, g, q, sl, t, baseq, spw, cb;
  u32 s, c;
  bool active;
  __device__ QuadMap(const slgpu_step_params& p) {
    T = (int)p.n_temps;
    spw = 8 / T;
    lane = threadIdx.x & 31;
    g = lane & 3;
    q = lane >> 2; /* chain slot within the warp */
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
    cb = threadIdx.x >> 2; /* chain slot within the block */
  }
};

/* chain held by block slot cq (slots follow the thread map above); false for idle slots */
__device__ __forceinline__ bool quad_slot_chain(const slgpu_step_params& p, int cq, u32& cc) {
  const int T = (int)p.n_temps, spw = 8 / T;
  const int w = cq >> 3, qq = cq & 7, sl = qq / T, t = qq - sl * T;
  const u32 s = (blockIdx.x * (blockDim.x >> 5) + (u32)w) * (u32)spw + (u32)sl;
  cc = s * (u32)T + (u32)t;
  return sl < spw && s < p.n_stacks;
}

__device__ __forceinline__ int tri_of(int j) { return (j * (j + 1)) >> 1; }

template <int D>
struct QuadDims {
  static constexpr int R = (D + 3) / 4;             /* rows per lane */
  static constexpr int NL = D * (D + 1) / 2;
  static constexpr int LP = (NL + 3) | 1;           /* odd stride between chains: spreads the quads over the banks */
  static constexpr int DP = D + 1;
};

/* dynamic shared memory of pt_step_quad_kernel */
template <int D, int NT>
inline size_t quad_smem_bytes(unsigned n_temps, unsigned n_jobs) {
  const size_t nch = NT / 4;
  const size_t nsb = (size_t)(NT / 32) * (8u / n_temps);
  return (nch * QuadDims<D>::LP + 2 * nch * QuadDims<D>::DP + nch * n_jobs + 2 * nsb * D) * sizeof(double);
}

/* ProposalShaper::update (src/infer/adaptive.cpp:173-209) by a lane quad on a factor in shared memory
 * (Lu[tri(j)+k]); jump[j] holds the accepted jump. */
template <int D>
__device__ __forceinline__ double shaper_update_quad(double* Lu, const double* jump, double count, double prop_norm, int g, int lane) {
  constexpr int R = QuadDims<D>::R;
  const double eps = 1e-18;
  const unsigned qm = 0xfu << (lane & ~3); /* the four lanes of this quad */
  const double sq = sqrt(count);
  const double rsq = 1.0 / sq;
  const double scale = sqrt((count - 1.0) / count);
  double xs[R], fr[R];
#pragma unroll
  for (int m = 0; m < R; ++m) {
    const int j = 4 * m + g;
    xs[m] = (j < D) ? div_by(jump[j], sq, rsq) : 0.0;
    fr[m] = 0.0;
  }
#pragma unroll
  for (int k = 0; k < D; ++k) {
    const int MK = k >> 2, GK = k & 3; /* row k is row MK of lane GK */
    __syncwarp(qm);
    const double xk = __shfl_sync(qm, xs[MK], (lane & ~3) | GK);
    const double Lkk = Lu[tri_of(k) + k] * scale;
    const double V = smax(eps, Lkk);
    const double rV = 1.0 / V;
    const double r = sqrt(V * V + xk * xk);
    const double c = div_by(r, V, rV);
    const double s = div_by(xk, V, rV);
    const double rc = 1.0 / c;
    __syncwarp(qm);
    if (g == GK) {
      Lu[tri_of(k) + k] = r;
      fr[MK] = fma(r, r, fr[MK]); /* the diagonal closes row k's sum */
    }
#pragma unroll
    for (int m = MK; m < R; ++m) {
      const int j = 4 * m + g;
      if (j > k && j < D) {
        const int o = tri_of(j) + k;
        double Ljk = Lu[o] * scale;
        Ljk = div_by(Ljk + s * xs[m], c, rc);
        Lu[o] = Ljk;
        xs[m] = c * xs[m] - s * Ljk;
        fr[m] = fma(Ljk, Ljk, fr[m]);
      }
    }
  }
  /* |L|_F: every lane assembles all D row sums and runs the canonical 32-slot halving tree */
  double f[4 * R];
#pragma unroll
  for (int m = 0; m < R; ++m) {
#pragma unroll
    for (int gg = 0; gg < 4; ++gg) f[4 * m + gg] = __shfl_sync(qm, fr[m], (lane & ~3) | gg);
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

template <class Lik, int D, int NT>
__global__ void __launch_bounds__(NT, SLGPU_QUAD_MINBLOCKS) pt_step_quad_kernel(const slgpu_step_params p) {
  constexpr int R = QuadDims<D>::R, NLd = QuadDims<D>::NL, LP = QuadDims<D>::LP, DP = QuadDims<D>::DP;
  constexpr int NCH = NT / 4;
  const QuadMap m(p);
  const int T = m.T, g = m.g, cb = m.cb, lane = m.lane;
  const size_t C = (size_t)p.n_stacks * p.n_temps;
  const bool active = m.active;
  const u32 c = active ? m.c : 0u;
  const u32 gchain = p.stack_offset * p.n_temps + c;
  const bool replay = (p.rng_mode == SLGPU_RNG_REPLAY);
  const int tid = threadIdx.x;
  const int JP = (int)p.n_job_types;

  extern __shared__ double s_dyn[];
  double* const s_L = s_dyn;                     /* [chain][LP] packed lower triangles, row-major */
  double* const s_v = s_L + NCH * LP;            /* [chain][DP] normals, then the proposal (what the functor reads) */
  double* const s_j = s_v + NCH * DP;            /* [slot][DP] staged accepted jumps */
  double* const s_f = s_j + NCH * DP;            /* [chain][J] per-job negative log-likelihoods */
  const int nsb = (NT / 32) * m.spw;
  double* const s_epm = s_f + NCH * JP;          /* EPSRDiagnostic M, [stack of block][D] */
  double* const s_eps = s_epm + nsb * D;
  const int sb = (tid >> 5) * m.spw + m.sl;
  __shared__ double s_mn[D], s_mx[D], s_n0[D], s_rd[D];
  __shared__ double s_w[SLGPU_MAX_TEMPS * 3], s_lad[SLGPU_MAX_TEMPS];
  __shared__ double s_cnt[NCH], s_nsc[NCH];
  __shared__ u32 s_owner[NCH];
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
  /* the block's L factors: tiled HBM store -> shared memory, consecutive threads read consecutive chains */
  for (int e = tid; e < NCH * NLd; e += NT) {
    const int cq = e % NCH, idx = e / NCH;
    u32 cc;
    if (quad_slot_chain(p, cq, cc)) s_L[cq * LP + idx] = l_base(p.L, cc, NLd)[(size_t)idx * kLs];
  }

  const Lik lik(p.payload);
  double* const Lc = s_L + cb * LP;
  double* const zv = s_v + cb * DP;
  double x[R];
#pragma unroll
  for (int mm = 0; mm < R; ++mm) {
    const int j = 4 * mm + g;
    x[mm] = (j < D) ? p.x[(size_t)j * C + c] : 0.0;
  }
  if (active && m.t == 0) {
#pragma unroll
    for (int mm = 0; mm < R; ++mm) {
      const int j = 4 * mm + g;
      if (j < D) {
        s_epm[sb * D + j] = p.ep_m[(size_t)j * p.n_stacks + m.s];
        s_eps[sb * D + j] = p.ep_s[(size_t)j * p.n_stacks + m.s];
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
    /* normals: lane g makes the Box-Muller pairs kp = g, g+4, ... of its chain */
    if (active) {
      const double* zr = replay ? p.rp_z + ((size_t)r * C + c) * D : nullptr;
#pragma unroll 1
      for (int kp = g; kp < (D + 1) / 2; kp += 4) {
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
        zv[2 * kp] = z0;
        if (2 * kp + 1 < D) zv[2 * kp + 1] = z1;
      }
    }
    __syncwarp();
    /* GaussianProposal::propose (sampler.cpp:79-88): a_j = sum_{k<=j} fma(N_jk, z_k, .), k ascending, rows j = 4m+g */
    double a[R];
#pragma unroll
    for (int mm = 0; mm < R; ++mm) a[mm] = 0.0;
    if (active) {
      const bool fresh = (sh_count == p.sh_count0);
#pragma unroll
      for (int k = 0; k < D; ++k) {
        const double zk = zv[k];
#pragma unroll
        for (int mm = k >> 2; mm < R; ++mm) {
          const int j = 4 * mm + g;
          if (j >= k && j < D) {
            double njk = Lc[tri_of(j) + k] * nscale;
            if (j == k) njk = fresh ? s_n0[k] : njk + 1e-4;
            a[mm] = fma(njk, zk, a[mm]);
          }
        }
      }
    }
    __syncwarp(); /* all four lanes are done reading z: the vector now takes the proposal */
    if (active) {
#pragma unroll
      for (int mm = 0; mm < R; ++mm) {
        const int j = 4 * mm + g;
        if (j < D) {
          a[mm] = bounce1(x[mm] + a[mm] * sigma, s_mn[j], s_mx[j], s_rd[j]);
          zv[j] = a[mm];
        }
      }
    }
    __syncwarp();
    /* energy: lane g evaluates the job types g, g+4, ...; every lane then adds all J values in job-type
     * order (sampler.cpp:148-150) */
    double e_new = 0.0;
    {
      const u32 J = (StaticJobs<Lik, D>::value > 0) ? (u32)StaticJobs<Lik, D>::value : p.n_job_types;
      if (active) {
        for (u32 jj = (u32)g; jj < J; jj += 4) s_f[cb * JP + jj] = lik(jj, zv, (u32)D);
      }
      __syncwarp();
      if (active) {
        for (u32 jj = 0; jj < J; ++jj) e_new += s_f[cb * JP + jj];
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
    /* staging slot: lane 0 of the quad draws it, the others copy it */
    int slot = 0;
    if (acc && g == 0) slot = atomicAdd(&s_nacc[par], 1);
    slot = __shfl_sync(0xffffffffu, slot, lane & ~3);
    if (acc) {
      sh_count += 1.0;
#pragma unroll
      for (int mm = 0; mm < R; ++mm) {
        const int j = 4 * mm + g;
        if (j < D) {
          s_j[slot * DP + j] = a[mm] - x[mm]; /* the accepted jump (sampler.cpp:166) */
          x[mm] = a[mm];
        }
      }
      if (g == 0) {
        s_cnt[slot] = sh_count;
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
    /* ================= phase B: compacted ProposalShaper updates, one quad per staged chain ================= */
    {
      const int nacc = s_nacc[par];
#ifdef SLGPU_QUAD_STRIDED
      const int bslot = (lane >> 2) * (NT / 32) + (tid >> 5); /* staged chains dealt round-robin to the warps */
#else
      const int bslot = tid >> 2; /* staged chains fill the first warps densely: the kernel is issue-bound, not latency-bound */
#endif
      if (bslot < nacc) {                                     /* quad-uniform; shuffles inside are quad-masked */
        const u32 own = s_owner[bslot];
        const double ns = shaper_update_quad<D>(s_L + own * LP, s_j + bslot * DP, s_cnt[bslot], p.prop_norm, g, lane);
        if (g == 0) s_nsc[own] = ns;
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
      /* chain slot q of this warp lives in lanes 4q..4q+3 (replicated scalars): read lane 4q+g */
      int cur = T - 1, my_src = m.t, mid_src = m.t;
      bool my_swapped = false;
      for (int tt = T - 2; tt >= 0; --tt) {
        const double El = shfl_d(energy, 4 * (m.baseq + tt) + g);
        const double Eh = shfl_d(energy, 4 * (m.baseq + cur) + g);
        const double bl = shfl_d(beta, 4 * (m.baseq + tt) + g);
        const double bh = shfl_d(beta, 4 * (m.baseq + tt + 1) + g);
        const double u = shfl_d(u_sw, 4 * (m.baseq + tt) + g);
        const double deltaEnergy = El - Eh;
        const double deltaBeta = bl - bh;
        const bool sw = u < slm_exp(deltaEnergy * deltaBeta);
        if (m.t == tt + 1) my_src = sw ? tt : cur;
        if (m.t == tt) { my_swapped = sw; mid_src = sw ? cur : tt; }
        if (!sw) cur = tt;
      }
      if (m.t == 0) my_src = cur;
      const double lb = slm_log(beta);
      const double lbh = shfl_d(lb, lane + 4 < 32 ? lane + 4 : lane);
      mid_e = shfl_d(energy, 4 * (m.baseq + mid_src) + g);
      mid_acc = __shfl_sync(0xffffffffu, accepted, 4 * (m.baseq + mid_src) + g);
      const double new_e = shfl_d(energy, 4 * (m.baseq + my_src) + g);
      const int new_acc = __shfl_sync(0xffffffffu, accepted, 4 * (m.baseq + my_src) + g);
#pragma unroll
      for (int mm = 0; mm < R; ++mm) x[mm] = shfl_d(x[mm], 4 * (m.baseq + my_src) + g);
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
          const int j = 4 * mm + g;
          if (j < D) {
            const int o = sb * D + j;
            const double mo = s_epm[o];
            const double xi = x[mm];
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

  __syncthreads();
  for (int e = tid; e < NCH * NLd; e += NT) { /* factors back to the tiled HBM store */
    const int cq = e % NCH, idx = e / NCH;
    u32 cc;
    if (quad_slot_chain(p, cq, cc)) l_base(p.L, cc, NLd)[(size_t)idx * kLs] = s_L[cq * LP + idx];
  }
  if (active) {
#pragma unroll
    for (int mm = 0; mm < R; ++mm) {
      const int j = 4 * mm + g;
      if (j < D) {
        p.x[(size_t)j * C + c] = x[mm];
        if (m.t == 0) {
          p.ep_m[(size_t)j * p.n_stacks + m.s] = s_epm[sb * D + j];
          p.ep_s[(size_t)j * p.n_stacks + m.s] = s_eps[sb * D + j];
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
