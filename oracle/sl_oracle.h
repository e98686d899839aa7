/*
 * sl_oracle.h -- C API of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a from-scratch CPU restatement of the
 * parallel-tempering hot path of NICTA/stateline (src/infer, the init half of
 * src/app/serverwrapper.cpp and the row layout of src/db/db.cpp).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product path (stateline_b200/, include/) never includes, links or
 * executes anything under oracle/.
 *
 * Parity pin status: PINNED AGAINST THE REFERENCE'S OWN CODE.  The reference's build (cmake + Eigen + ZeroMQ) does
 * not run in this image, but its src/infer/{sampler,chainarray,adaptive}.cpp and src/db/db.cpp compile UNMODIFIED
 * against the stand-in Eigen / zmq headers of oracle/ref_shim into oracle/_ref/libsl_ref.so (oracle/Makefile.ref,
 * oracle/ref_harness.cpp).  Fed the variates the reference itself drew, this oracle (sequential schedule, glibc
 * exp/log build libsl_oracle_libm.so) reproduces the reference bit for bit: chain rows, sigma/beta, tier models,
 * proposal shapes, cold rows, CSV text, flush -- tests/test_ref_pin.py, live and from the committed fixtures
 * tests/golden/ref_*.json.  What stays unpinned by a third party: the three places where the reference delegates
 * to Eigen (GEMV summation order, Frobenius norm order, the 3x3 colPivHouseholderQr), where shim and oracle share
 * one canonical order.  Further pins:
 *   - the reference's own EPSR known answers (src/test/diagnostics.cpp:34,53,78),
 *   - the swap-formula case of src/test/chainarray.cpp:104-125,
 *   - hand-derived KATs in tests/test_oracle.py,
 *   - the README's soft statistics (README.md:421-435).
 */
#ifndef SL_ORACLE_H
#define SL_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* built-in synthetic targets (negative log-likelihood per job type) */
enum {
  SLO_TARGET_DEMO_GAUSS = 0,   /* src/bin/demo-worker.cpp:37-45: 0.5*|x|^2 for every job type   */
  SLO_TARGET_DOUBLE_WELL = 1,  /* f(0)=8(x0^2-1)^2, f(1)=0.5*x1^2                               */
  SLO_TARGET_GAUSS_FACT = 2,   /* job j owns dims 2j,2j+1: 0.5*sum(((x-mu)*inv_s)^2); params = mu[d], inv_s[d] */
  SLO_TARGET_GAUSS_CORR = 3,   /* 0.5*|U x|^2 split in J row blocks; params = U[d*d] row-major upper-tri      */
  SLO_TARGET_ROSENBROCK = 4    /* d-1 terms ((100(x_{i+1}-x_i^2)^2+(1-x_i)^2)/20) split evenly over J jobs     */
};

/* schedules */
enum {
  SLO_SCHED_SEQUENTIAL = 0, /* the reference's own order: Sampler::step() state machine fed by a FIFO   */
  SLO_SCHED_BATCHED = 1,    /* tier models frozen between adapt points, updated from summed statistics  */
  SLO_SCHED_PER_STACK = 2   /* sequential, but every stack owns its tier models (= S servers, nStacks=1)*/
};

enum { SLO_RNG_REPLAY = 0, SLO_RNG_PHILOX = 1 };

typedef struct slo_config {
  uint32_t n_stacks, n_temps, n_dims, n_job_types, swap_interval;
  uint32_t adapt_interval;      /* batched schedule only; 0 => swap_interval */
  double opt_accept, opt_swap;
  const double* min;            /* [d] */
  const double* max;            /* [d] */
  int32_t use_initial;
  const double* initial;        /* [d] when use_initial */
  int32_t target;
  const double* target_params;  /* see enum */
  int32_t schedule;
  int32_t rng_mode;
  uint64_t seed;
  uint32_t stack_offset;        /* global index of stack 0 (multi-GPU sharding of Philox streams) */
  int32_t wire_quantize;        /* 1 => x and nll go through the "%f" text round trip (requester.cpp:37, minion.cpp:53) */
} slo_config;

typedef struct slo_sampler slo_sampler;

slo_sampler* slo_create(const slo_config* cfg);
void slo_destroy(slo_sampler*);

/* replay streams, all indexed by 0-based round r and chain id (id = s*T+t):
 *   z      [n_rounds][S*T][d]  standard normals of the proposal evaluated in round r
 *   u_mh   [n_rounds][S*T]     MH uniforms
 *   u_swap [n_rounds][S*T]     swap uniforms, indexed by the LOWER chain id of the pair
 *   x_init [S*T][d]            pre-bounce initial points (ignored with use_initial)
 * The pointers must stay valid while the sampler runs. */
void slo_set_replay(slo_sampler*, const double* z, const double* u_mh, const double* u_swap,
                    const double* x_init, uint64_t n_rounds);

void slo_init_chains(slo_sampler*);             /* serverwrapper.cpp:60-81 */
void slo_run_rounds(slo_sampler*, uint64_t n);  /* n canonical rounds (every chain appends once per round) */
void slo_flush(slo_sampler*);                   /* Sampler::flush sampler.cpp:216-235 (sequential schedules) */
uint64_t slo_rounds_done(const slo_sampler*);

/* ---- state inspection (all chains) ---- */
/* last row of every chain: rows[S*T][d+5] = x.., energy, sigma, beta, accepted, swapType */
void slo_get_last_rows(const slo_sampler*, double* rows);
void slo_get_sigma_beta(const slo_sampler*, double* sigma, double* beta); /* ChainArray sigma_/beta_ */
void slo_get_shaper(const slo_sampler*, uint32_t chain, double* L /*[d*d] row-major*/, double* N, double* count);
/* tier models: arrays of n_models entries (T for shared, S*T for per-stack) */
uint32_t slo_n_models(const slo_sampler*);
void slo_get_models(const slo_sampler*, int which /*0 sigma,1 beta*/, double* mu_xx /*[n][9]*/,
                    double* mu_xy /*[n][3]*/, double* weight /*[n][3]*/, double* count /*[n]*/);
/* TableLogger counters (infer/logging.cpp:29-48), per chain */
void slo_get_counters(const slo_sampler*, uint64_t* length, uint64_t* accepts, uint64_t* swaps,
                      uint64_t* swap_attempts, double* min_energy);
/* TableLogger energies_ and the adapters' windowed rates_ (logging.cpp:43, adaptive.cpp:80-90), per chain */
void slo_get_logger(const slo_sampler*, double* cur_energy, double* accept_rate, double* swap_rate);
void slo_window_rates(const int* flags, uint32_t n, double* rates);
/* the table + convergence line TableLogger::update prints (logging.cpp:52-87); returns the length */
size_t slo_format_table(const slo_sampler*, char* out, size_t cap);
/* cold-chain output rows in emission order: returns number of rows of that stack; rows[n][d+5].
 * Row 0 is the initial state (accepted=1). */
uint64_t slo_cold_rows(const slo_sampler*, uint32_t stack, double* rows, uint64_t cap_rows);
void slo_rhat(const slo_sampler*, double* rhat /*[d]*/);

/* ---- stand-alone pieces (unit pins) ---- */
void slo_bouncy_bounds(const double* val, const double* mn, const double* mx, uint32_t d, double* out);
void slo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* the draw functions of the native RNG scheme */
void slo_draw_normals(uint64_t seed, uint32_t chain, uint64_t round, uint32_t d, double* z);
double slo_draw_uniform(uint64_t seed, uint32_t chain, uint64_t round, uint32_t purpose);
void slo_draw_init(uint64_t seed, uint32_t chain, uint32_t d, double* x);
void slo_qr_solve3(const double A[9] /*row-major*/, const double b[3], double w[3]);
/* ProposalShaper::update on a dense row-major L (d*d); returns N and the new count */
void slo_shaper_update(double* L, double* N, double* count, const double* step, uint32_t d, double prop_norm);
/* EPSR: feed samples[n][m_chains][d] round-robin, get rhat[d] */
void slo_epsr(const double* samples, uint32_t n, uint32_t m, uint32_t d, double* rhat);
/* RegressionAdapter pieces */
void slo_regression_init(double min_cap, double max_cap, double mu_xx[9], double mu_xy[3], double w[3], double* count);
double slo_regression_predict(const double w[3], double opt, double min_cap, double max_cap, double side);
void slo_regression_update(double mu_xx[9], double mu_xy[3], double w[3], double* count, double min_cap,
                           double max_cap, double logval, double side, int acc);
double slo_reduce_tree256(const double* v, uint32_t n, uint32_t stride);
/* one text row as CSVChainArrayWriter would print it (db.cpp:18-25), without the newline */
size_t slo_csv_row(const double* row, uint32_t d, char* out, size_t cap);
double slo_target_eval(int32_t target, const double* params, uint32_t d, uint32_t n_job_types, uint32_t job,
                       const double* x);

/* ---- CPU baseline: "server + N workers"  ----
 * Runs the sequential schedule with N worker threads pulling (chain, jobType) jobs from a FIFO queue
 * (<=10 in flight per worker, delegator.cpp:228), x and nll passed as "%f" text.  n_workers==0 runs inline
 * (single thread, no queue, no text).  Native Philox draws.  Returns wall seconds for n_rounds rounds
 * after chain initialisation; *chain_steps receives the number of chain steps done. */
double slo_time_server_workers(const slo_config* cfg, uint32_t n_workers, uint64_t n_rounds,
                               uint64_t* chain_steps);

#ifdef __cplusplus
}
#endif
#endif
