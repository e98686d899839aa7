/*
 * slgpu_plugin.h -- binary interface between the engine (libslgpu.so) and a likelihood plugin .so.
 *
 * In NICTA/stateline a user likelihood is a host callback
 *     typedef std::function<double(uint, const std::vector<double>&)> LikelihoodFn;   (src/app/workerwrapper.hpp:21)
 * run by worker processes and reached through the Requester -> Delegator -> Worker -> Minion
 * ZeroMQ round trip (src/comms/requester.cpp:27-57, src/app/workerwrapper.cpp:24-36).  On the GPU
 * path the likelihood is a __device__ functor.  A plugin .so instantiates the step kernels of
 * slgpu_device.cuh with its functor (so the functor inlines; no device function pointers, no -rdc)
 * and exports a table of host launchers through
 *     extern "C" const slgpu_plugin_v1* slgpu_plugin_entry(void);
 * The engine dlopen()s the plugin, fills slgpu_step_params with device pointers it owns and calls
 * the launchers on its compute stream.  Plain C, no CUDA or torch types in the signatures.
 */
#ifndef SLGPU_PLUGIN_H
#define SLGPU_PLUGIN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLGPU_PLUGIN_ABI 5u
#define SLGPU_MAX_DIMS 128u  /* largest nDims the generic (runtime-d) kernels accept */
#define SLGPU_MAX_TEMPS 32u  /* a stack's temperature ladder lives in one warp */
#define SLGPU_L_TILE 32u    /* chain slots per tile of the packed-L store (see slgpu_step_params.L) */
/* how the packed L factors are stored (slgpu_step_params.L): small dimensions are tiled by stack group (the
 * floor(32 / n_temps) stacks whose chains share a warp) for the tile kernel, which keeps a group's tile in shared
 * memory for a whole launch; large ones are kept per chain, packed column-major, for the warp-per-chain kernel */
#define SLGPU_L_TILED 0
#define SLGPU_L_COLMAJOR 1
#define SLGPU_L_LAYOUT(n_dims) ((n_dims) > 32u ? SLGPU_L_COLMAJOR : SLGPU_L_TILED)
/* doubles in the L store of n_stacks x n_temps chains of dimension n_dims */
#define SLGPU_L_DOUBLES(n_stacks, n_temps, n_dims)                                                                      \
  (SLGPU_L_LAYOUT(n_dims) == SLGPU_L_COLMAJOR                                                                           \
       ? (((size_t)(n_stacks) * (n_temps) + 127u) / 128u * 128u) * ((size_t)(n_dims) * ((n_dims) + 1u) / 2u)           \
       : (size_t)(((n_stacks) + 32u / (n_temps) - 1u) / (32u / (n_temps))) * ((size_t)(n_dims) * ((n_dims) + 1u) / 2u) * SLGPU_L_TILE)
#define SLGPU_JUMP_LOG_SLOTS 32u /* "gpu": {"shaper": "welford"}: accepted jumps a chain can log between two re-Cholesky passes */
#define SLGPU_ROW_EXTRA 5u   /* energy, sigma, beta, accepted, swapType (src/db/db.cpp:18-25) */

enum { SLGPU_RNG_REPLAY = 0, SLGPU_RNG_PHILOX = 1 };

/* Everything a launch needs.  All pointers are device pointers owned by the engine.
 * Chain id = stack * n_temps + temperature, temperature 0 coldest (src/infer/chainarray.hpp:28-37).
 * C = n_stacks * n_temps.  Per-chain arrays are indexed [chain]; vectors and matrices are stored
 * structure-of-arrays, element-major: x[i * C + chain]; the packed lower triangle of L (row-major, nL =
 * d(d+1)/2 entries) is tiled by stack group: with spw = 32 / n_temps stacks per group, chain (stack, temp) is slot
 * (stack % spw) * n_temps + temp of tile stack / spw, and entry (r,k) of its factor is
 * L[((stack / spw) * nL + r*(r+1)/2 + k) * 32 + slot] -- a group's factors are one contiguous tile of nL * 256 bytes
 * (one bulk copy into shared memory) and a warp reads one 256-byte line per entry;
 * p_sig[(k * T + temp) * S + stack], ep_m[i * n_stacks + stack]. */
typedef struct slgpu_step_params {
  uint32_t n_stacks, n_temps, n_dims, n_job_types;
  uint32_t swap_interval;  /* swap sweep in the round that makes the chain length a multiple of it */
  uint32_t stack_offset;   /* global index of local stack 0: Philox streams of a sharded run */
  uint32_t n_rounds;       /* rounds this launch advances every chain by */
  int32_t rng_mode;        /* SLGPU_RNG_* */
  uint64_t seed;
  uint64_t round_begin;    /* 0-based index of the first round of this launch */
  uint64_t replay_rounds;  /* rounds held by the replay buffers */
  double opt_accept;       /* optimalAcceptRate */
  double sh_count0;        /* ProposalShaper initial_count (src/app/serverwrapper.cpp:50) */
  double prop_norm;        /* |(max-min)/4| (src/infer/adaptive.cpp:163) */
  const double* mn;        /* [d] */
  const double* mx;        /* [d] */
  const double* n0_diag;   /* [d] range/|range|: diagonal of the initial N (adaptive.cpp:164-169) */
  const double* initial;   /* [d] or NULL (useInitial) */
  /* chain state */
  double* x;               /* [d][C]  sample of the last row */
  double* L;               /* ProposalShaper L_ (adaptive.cpp:158), lower triangle.  d <= 32 (SLGPU_L_TILED):
                            * [ceil(S/spw)][d(d+1)/2][32], row-major packed.  d > 32 (SLGPU_L_COLMAJOR):
                            * [C][d(d+1)/2], column-major packed: L(j,k) at k*d - k(k-1)/2 + (j-k) */
  double* energy;          /* [C] */
  double* row_sigma;       /* [C] sigma stored in the last row (stale after a reject, chainarray.cpp:102-106) */
  double* row_beta;        /* [C] */
  double* sigma;           /* [C] ChainArray::sigma_ == RegressionAdapter::values_ */
  double* beta;            /* [C] ChainArray::beta_ */
  double* nscale;          /* [C] prop_norm / max(|L|_F, eps): N = L * nscale + 1e-4 I (adaptive.cpp:200-204) */
  double* sh_count;        /* [C] ProposalShaper count_ */
  int32_t* accepted;       /* [C] last row's accepted flag */
  int32_t* swap_type;      /* [C] last row's swapType */
  /* TableLogger counters (src/infer/logging.cpp:29-48) */
  uint32_t* lg_accepts;
  uint32_t* lg_swaps;
  uint32_t* lg_swap_attempts;
  double* lg_min_energy;
  double* lg_cur_energy;   /* energy of the last state the logger saw (mid-sweep on swap rounds, logging.cpp:43) */
  /* optional per-round log for the windowed rates of the table logger (RegressionAdapter::rates_, adaptive.cpp:80-90):
   * round_flags[(round - round_begin) * C + chain] = bit 0: this chain's proposal was accepted that round,
   * bits 1-2: swapType of the pair (chain, chain+1) that round.  NULL = not wanted. */
  unsigned char* round_flags;
  /* tier models, frozen while a launch runs */
  const double* sig_w;     /* [T][3] weights of the sigma regression */
  const double* ladder;    /* [T] beta ladder from the last adapt point */
  /* regression sufficient statistics gathered since the last adapt point */
  double* p_sig;           /* [9][T][S] */
  double* p_beta;          /* [9][T][S] */
  /* EPSRDiagnostic Welford state of the cold chains (src/infer/diagnostics.cpp:27-47) */
  double* ep_m;            /* [d][S] */
  double* ep_s;            /* [d][S] */
  /* cold-chain rows of this launch: rows_out[(round - round_begin) * S + stack][d + 5], or NULL */
  double* rows_out;
  /* recorded draws (SLGPU_RNG_REPLAY): z[round][C][d], u_mh[round][C], u_swap[round][C], x_init[C][d] */
  const double* rp_z;
  const double* rp_umh;
  const double* rp_uswap;
  const double* rp_xinit;
  const void* payload;     /* the plugin's POD payload, copied to the device by the engine */
  /* "gpu": {"shaper": "welford"} -- NOT the reference's arithmetic, off by default (moment == NULL): the second moment of the
   * accepted jumps is kept as a matrix, moment[] in the layout of L.  An accept only appends its jump vector to the chain's
   * log, jump_log[(slot * d + i) * C + chain] with slot = chol_dirty[chain]++ (at most SLGPU_JUMP_LOG_SLOTS between two calls of
   * launch_rechol); launch_rechol, at every recholEvery-th adapt point, folds the logged jumps into the matrix (C <- (n-1)/n C + s s^T / n, the
   * recurrence L L^T follows in adaptive.cpp:173-197) and re-factorises L from it. */
  double* moment;
  double* jump_log;
  unsigned char* chol_dirty; /* [C] jumps logged since the last re-Cholesky */
} slgpu_step_params;

/* Launchers return a cudaError_t value (0 = success); stream is a cudaStream_t. */
typedef struct slgpu_plugin_v1 {
  uint32_t abi_version;  /* SLGPU_PLUGIN_ABI */
  const char* name;
  uint32_t n_job_types;  /* job types the likelihood is defined for; 0 = any */
  uint32_t n_dims;       /* dimension the likelihood is defined for; 0 = any */
  size_t payload_bytes;  /* expected payload size; 0 = none / variable */
  /* chain initialisation (src/app/serverwrapper.cpp:60-81); writes the initial cold rows to rows_out[S][d+5] */
  int (*launch_init)(const slgpu_step_params* p, void* stream);
  /* n_rounds canonical rounds of Sampler::step for every chain (src/infer/sampler.cpp:142-203) */
  int (*launch_step)(const slgpu_step_params* p, void* stream);
  /* nll[point * J + job] = f(job, x[point * d ..]) for n_points points (point-major x): contract check */
  int (*launch_eval)(const slgpu_step_params* p, const double* x, uint32_t n_points, double* nll, void* stream);
  /* "gpu": {"shaper": "welford"} only: L = chol(moment), nscale = prop_norm / |L|_F for the chains marked dirty.  Returns
   * cudaErrorNotSupported when the plugin has no Welford kernels for this dimension. */
  int (*launch_rechol)(const slgpu_step_params* p, void* stream);
} slgpu_plugin_v1;

typedef const slgpu_plugin_v1* (*slgpu_plugin_entry_fn)(void);

#ifdef __cplusplus
}
#endif
#endif
