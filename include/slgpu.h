/*
 * slgpu.h -- C-ABI of the B200 parallel-tempering engine (libslgpu.so).
 *
 * Drop-in boundary for the src/infer hot path of NICTA/stateline.  What each entry point stands in
 * for on the reference side (paths relative to the reference root):
 *
 *   slgpu_create        StatelineSettings::fromJSON (src/app/serverwrapper.hpp:64-91) +
 *                       ProposalBoundsFromJSON (src/infer/sampler.hpp:34-60) + the construction half of
 *                       runSampler (src/app/serverwrapper.cpp:48-55): adapters, proposal, chain array.
 *                       The likelihood comes from a plugin .so (slgpu_plugin.h) instead of
 *                       WorkerWrapper(LikelihoodFn, ...) (src/app/workerwrapper.hpp:21-31).
 *   slgpu_init_chains   the initialisation loop, src/app/serverwrapper.cpp:60-81.
 *   slgpu_step_rounds   n canonical rounds of Sampler::step() (src/infer/sampler.cpp:142-203): every
 *                       chain of every stack appends once per round.
 *   slgpu_run           the main loop and stop rule of runSampler (src/app/serverwrapper.cpp:100-134):
 *                       initialise, step until #cold-chain samples >= nSamplesTotal, flush the sink.
 *   slgpu_drain         ChainArray::flushToDisk / CSVChainArrayWriter::append (src/infer/chainarray.cpp:131-155,
 *                       src/db/db.cpp:37-47): hands the caller the cold-chain rows
 *                       x_0..x_{d-1},energy,sigma,beta,accepted,swapType in emission order.
 *   slgpu_rhat          EPSRDiagnostic::rHat (src/infer/diagnostics.cpp:49-71).
 *   slgpu_get_*         what TableLogger prints (src/infer/logging.cpp:35-87): per-chain counters,
 *                       sigma, beta; plus adapter / shaper state for parity tests.
 *
 * Error convention: the reference exits the process on a missing config key
 * (src/app/jsonsettings.hpp:22-35) and LOG(FATAL)s on bound-size mismatch (sampler.hpp:42-46).  Nothing
 * here exits or throws across the ABI: every call returns slgpu_status and slgpu_last_error() holds a
 * message for the calling thread.  There is no CPU fallback: without a CUDA device slgpu_create fails
 * with SLGPU_E_CUDA.
 *
 * Ownership: the engine owns all device memory, its streams and the host ring it allocates; the caller
 * owns the config string, the payload (copied at create) and any buffer it passes in.  An engine
 * handle is not thread-safe.
 */
#ifndef SLGPU_H
#define SLGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct slgpu_engine slgpu_engine;

typedef enum slgpu_status {
  SLGPU_OK = 0,
  SLGPU_E_CONFIG = 1, /* missing / malformed key, bound-size mismatch, unsupported sizes */
  SLGPU_E_PLUGIN = 2, /* plugin .so missing, wrong ABI, wrong job-type / dimension count */
  SLGPU_E_CUDA = 3,   /* no device, allocation or launch failure */
  SLGPU_E_ARG = 4,    /* NULL or out-of-range argument */
  SLGPU_E_STATE = 5,  /* call not valid in the engine's current state */
  SLGPU_E_IO = 6,     /* output directory / file failure */
  SLGPU_E_FULL = 7    /* host ring cannot take the rows of the requested rounds: drain first */
} slgpu_status;

/* message of the last failing call on this thread ("" if none) */
const char* slgpu_last_error(void);
/* "slgpu <version> sm_100a" */
const char* slgpu_version(void);

/*
 * config_json: the stateline server config, same keys as src/app/serverwrapper.hpp:64-91, all
 * mandatory except "initial" (needed when useInitial is true):
 *   nStacks nTemperatures nSamplesTotal swapInterval loggingRateSec nJobTypes outputPath
 *   optimalAcceptRate optimalSwapRate useInitial [initial] min max
 * plus an optional "gpu" object:
 *   device (int, default 0)           seed (default 1234)        stackOffset (default 0)
 *   rng "philox" | "replay"           sink "device" | "host" | "csv" | "none" (default "device")
 *   adaptInterval (default swapInterval)   maxRoundsPerLaunch (default 64)
 *   schedule "batched" (default: tier models frozen between adapt points) | "sequential" (validation mode: the
 *     reference's own order, Sampler::step() src/infer/sampler.cpp:142-203 -- one round per launch, then the shared
 *     tier models are replayed chain by chain; bit-identical to the reference's schedule, serial over the chains)
 *   hostRingRounds (default 8 * rounds per launch)   overwrite (bool, default false)
 *   wire "delta" (default) | "full": what the copy engine carries to the host for sink host / csv / bin.  "full" = every
 *     cold row, d + 5 doubles.  "delta" = one flag byte per row plus the d + 3 doubles of a row only when the cold
 *     state changed (a rejected proposal re-pushes a copy of the last state, src/infer/chainarray.cpp:102-106, so about
 *     three rows in four repeat); slgpu_drain / slgpu_peek / the csv files rebuild the exact rows on the host and are
 *     byte-identical for both settings.  With overwrite every segment starts with a full set of rows.
 *   shaper "reference" (default) | "welford": NOT the reference's arithmetic.  "reference" is ProposalShaper::update
 *     (src/infer/adaptive.cpp:173-209: a rank-1 Cholesky update of the second moment of the accepted jumps on every accept)
 *     bit for bit.  "welford" keeps that second moment as a matrix: an accept only logs its jump vector, and at the adapt
 *     points (every adaptInterval rounds) a second kernel folds the logged jumps into the matrix and re-factorises it: the
 *     variant BASELINE's north star names, without the serial column sweep in the round; same adaptation in exact
 *     arithmetic, delayed by at most adaptInterval rounds.  Validated statistically only; nDims <= 32, dimensions the
 *     plugin lists in SLGPU_DEFINE_PLUGIN, batched schedule.  recholEvery (default 1): re-factorise at every n-th adapt
 *     point only; adaptInterval * recholEvery <= 32 (the slots of the jump log per chain).
 *     Measured on config #3: step kernel 30 us per round instead of 47, the re-Cholesky 17 us per round on top: 4 %
 *     faster than the default in steady state, 9-10 % while the hot tiers still accept everything.
 *   sink "bin": the delta stream appended to <outputPath>/chains-<stackOffset>.slgpu as it lands (stateline_b200/binlog.py
 *     reads it back into rows); the sink that keeps up with the GPU, csv is for small runs
 * plugin_path: likelihood plugin .so; plugin_entry: its table symbol (NULL = "slgpu_plugin_entry").
 * payload: POD handed to the functor's constructor, copied to the device (may be NULL).
 */
slgpu_status slgpu_create(const char* config_json, const char* plugin_path, const char* plugin_entry,
                          const void* payload, size_t payload_bytes, slgpu_engine** out);
void slgpu_destroy(slgpu_engine* e);

/* recorded draws for rng "replay" (host pointers, copied to the device):
 * z[n_rounds][C][d], u_mh[n_rounds][C], u_swap[n_rounds][C] (indexed by the lower chain of the pair),
 * x_init[C][d] pre-bounce initial points.  C = nStacks * nTemperatures, chain id = stack*T + temp. */
slgpu_status slgpu_set_replay(slgpu_engine* e, const double* z, const double* u_mh, const double* u_swap,
                              const double* x_init, uint64_t n_rounds);

slgpu_status slgpu_init_chains(slgpu_engine* e);
/* enqueue n rounds (asynchronous; returns once the launches and sink copies are queued) */
slgpu_status slgpu_step_rounds(slgpu_engine* e, uint64_t n_rounds);
/* wait for everything queued so far; reports asynchronous CUDA errors */
slgpu_status slgpu_sync(slgpu_engine* e);
/* wait for everything queued so far WITHOUT finalising the sink (the csv sink keeps its text buffers for the next 64 MiB /
 * end-of-run flush instead of re-opening every <stack>.csv): what a host loop calls between chunks */
slgpu_status slgpu_wait(slgpu_engine* e);
/* the reference's whole run: init (if needed), ceil(nSamplesTotal / nStacks) rounds, flush the sink */
slgpu_status slgpu_run(slgpu_engine* e);
uint64_t slgpu_rounds_done(const slgpu_engine* e);

/* ---- cold-chain rows (sink "host"): emission order is round-major, stack-minor; the first nStacks
 * rows are the initial states.  Row = d + 5 doubles (src/db/db.cpp:18-25). ---- */
slgpu_status slgpu_drain(slgpu_engine* e, double* rows, size_t cap_rows, size_t* n_rows);
/* zero-copy view of the pinned host ring: landed-but-unconsumed rows as up to two spans */
slgpu_status slgpu_peek(slgpu_engine* e, const double** span0, size_t* n0, const double** span1, size_t* n1);
slgpu_status slgpu_advance(slgpu_engine* e, size_t n_rows);
uint64_t slgpu_rows_emitted(const slgpu_engine* e); /* rows produced on the device so far */

/* ---- the delta stream itself (sink "host", wire "delta"): zero-copy view of the oldest landed segment in the pinned
 * ring, for consumers that do not want the rows rebuilt.  A segment covers the cold rows of one launch, rows
 * row_begin .. row_begin + n_rounds * n_stacks - 1 in emission order:
 *   flags[round * n_stacks + stack]   bit 0 accepted, bits 1-2 swapType, bit 3 "changed"
 *   payload[slot * payload_doubles]   x_0..x_{d-1}, energy, sigma, beta of a changed row
 *   base[round * n_chunks + chunk]    slot of the first changed row among stacks 256 * chunk .. 256 * chunk + 255 of that
 *                                     round; the next changed row of the chunk has the next slot
 * An unchanged row repeats x, energy, sigma, beta of the same stack's previous row (src/infer/chainarray.cpp:102-106);
 * a key segment (key != 0: the first one, the first after slgpu_load_state, every one with gpu.overwrite) marks all rows
 * of its first round changed.  slgpu_next_segment leaves n_rounds = 0 when nothing has landed (block = 0) or nothing is
 * queued; slgpu_release_segment frees the ring space.  Mixing with slgpu_drain is allowed at segment boundaries. */
typedef struct slgpu_segment {
  uint32_t n_rounds, n_stacks, n_chunks, payload_doubles, count, key;
  uint64_t row_begin;
  const uint32_t* base;
  const unsigned char* flags;
  const double* payload;
} slgpu_segment;
slgpu_status slgpu_next_segment(slgpu_engine* e, int block, slgpu_segment* out);
slgpu_status slgpu_release_segment(slgpu_engine* e);
/* release without keeping the host's row-rebuilding state in step (slgpu_release_segment copies the changed rows for that:
 * ~4 MB per segment at config #3).  For consumers that only ever take segments.  Afterwards slgpu_drain / slgpu_peek can
 * rebuild rows again only from a key segment (SLGPU_E_STATE before that): slgpu_request_key_segment makes the next launch
 * start one. */
slgpu_status slgpu_discard_segment(slgpu_engine* e);
slgpu_status slgpu_request_key_segment(slgpu_engine* e);
/* bytes handed to the copy engine for the row sink so far (what `e2e.d2h_bytes_per_step` of bench.py counts) */
uint64_t slgpu_d2h_bytes(const slgpu_engine* e);

/* ---- device-side timing of what was queued between the two calls (CUDA events on the compute stream) ---- */
slgpu_status slgpu_timer_start(slgpu_engine* e);
/* the same without the event pair around every step-kernel launch (slgpu_timer_step_kernel reports 0 launches then): an
 * event between two kernels keeps the second from starting early under programmatic dependent launch, so this is the
 * timer that sees the step loop as it runs outside a measurement */
slgpu_status slgpu_timer_start_coarse(slgpu_engine* e);
slgpu_status slgpu_timer_stop(slgpu_engine* e, double* elapsed_ms); /* waits for the stop event */
uint64_t slgpu_launch_count(const slgpu_engine* e);                 /* kernels launched so far */
/* of the last timed region: summed device time of the step-kernel launches alone (their own event pairs) */
slgpu_status slgpu_timer_step_kernel(slgpu_engine* e, double* total_ms, uint64_t* n_launches);

/* ---- state inspection (synchronises) ---- */
uint32_t slgpu_n_dims(const slgpu_engine* e);
uint32_t slgpu_n_chains(const slgpu_engine* e);
/* rows[C][d+5]: last row of every chain */
slgpu_status slgpu_get_last_rows(slgpu_engine* e, double* rows);
slgpu_status slgpu_get_sigma_beta(slgpu_engine* e, double* sigma, double* beta);
/* dense row-major d*d L and N of one chain (either may be NULL) and the shaper count */
slgpu_status slgpu_get_shaper(slgpu_engine* e, uint32_t chain, double* L, double* N, double* count);
/* tier models: which = 0 sigma, 1 beta; mu_xx[T][9], mu_xy[T][3], weight[T][3], count[T] (any may be NULL) */
slgpu_status slgpu_get_models(slgpu_engine* e, int which, double* mu_xx, double* mu_xy, double* weight,
                              double* count);
slgpu_status slgpu_get_ladder(slgpu_engine* e, double* ladder);
/* TableLogger counters per chain (src/infer/logging.cpp:29-48); any pointer may be NULL */
slgpu_status slgpu_get_counters(slgpu_engine* e, uint64_t* length, uint64_t* accepts, uint64_t* swaps,
                                uint64_t* swap_attempts, double* min_energy);
/* EPSR per dimension over this engine's stacks (diagnostics.cpp:49-71) */
slgpu_status slgpu_rhat(slgpu_engine* e, double* rhat);
/* sharded EPSR: stage 1 returns sum_s M_s per dim and the stack / sample counts; after the caller
 * has all-reduced them, stage 2 returns sum_s (M_s - mean)^2 and sum_s S_s per dim for a second
 * all-reduce; slgpu_rhat_finish applies diagnostics.cpp:58-70 to the reduced sums. */
slgpu_status slgpu_rhat_stage1(slgpu_engine* e, double* sum_m, double* n_stacks, double* n_samples);
slgpu_status slgpu_rhat_stage2(slgpu_engine* e, const double* mean, double* sum_dm2, double* sum_s);
void slgpu_rhat_finish(const double* sum_dm2, const double* sum_s, double n_stacks_total, double n_samples,
                       uint32_t d, double* rhat);
/* end-of-run summary as JSON: per-tier accept / swap rates, sigma, beta, R-hat, rounds, rows */
slgpu_status slgpu_summary(slgpu_engine* e, char* json, size_t cap, size_t* needed);

/* TableLogger (src/infer/logging.cpp:35-87).  slgpu_get_logger returns, per chain, energies_ (the energy of the last
 * state the logger saw) and the adapters' windowed accept / swap rates (RegressionAdapter::rates(),
 * adaptive.cpp:80-90: the AcptRt and SwapRt columns); the rates need "gpu": {"windowRates": true} (SLGPU_E_STATE
 * otherwise; any of the three pointers may be NULL).  slgpu_format_table renders the table the reference prints
 * every loggingRateSec -- ID, Length, MinEngy, CurrEngy, Sigma, AcptRt, GlbAcptRt, Beta, SwapRt, GlbSwapRt, a blank
 * line between stacks -- followed by its "Convergence test: <mean R-hat> (not converged|possibly converged)" line,
 * from a snapshot of the device counters.  *needed = bytes including the terminator. */
slgpu_status slgpu_get_logger(slgpu_engine* e, double* cur_energy, double* accept_rate, double* swap_rate);
slgpu_status slgpu_format_table(slgpu_engine* e, char* out, size_t cap, size_t* needed);

/* Checkpoint / resume.  The reference declares ChainArray::recoverFromDisk (src/infer/chainarray.hpp:158-162) and
 * never defines it, so a stateline run cannot be resumed; this is the device-side equivalent.  slgpu_save_state
 * waits for the rounds issued so far and writes every chain's state (sample, energy, sigma, beta, shaper factor,
 * counters), the tier models, the ladder, the pending regression statistics, the R-hat accumulators and the
 * round counter to `path` (written to path.tmp, then renamed).  slgpu_load_state restores them into an engine
 * created from the same configuration and plugin (SLGPU_E_CONFIG otherwise; SLGPU_E_IO for a missing, foreign or
 * truncated file, which leaves the engine untouched); it replaces slgpu_init_chains.  The random streams are
 * addressed by round index, so a resumed run continues bit for bit.  Rows emitted before the checkpoint are not
 * emitted again; with sink "csv" create the resuming engine with "gpu": {"append": true}.  The logging windows of
 * "gpu": {"windowRates": true} are optional on both sides: a checkpoint without them resumes with empty windows. */
slgpu_status slgpu_save_state(slgpu_engine* e, const char* path);
slgpu_status slgpu_load_state(slgpu_engine* e, const char* path);

/* ---- several GPUs, one process.  The reference is ONE `stateline -c cfg` process that owns the whole run
 * (src/bin/stateline.cpp:52-90, src/app/serverwrapper.cpp:44-138); a group does that for several devices.  The stacks of
 * the config are sharded contiguously over the devices (devices = NULL: gpu.device, gpu.device + 1, ... n_devices of them,
 * 0 = all visible; or an explicit list, which may name a device more than once), every shard is an engine with its own
 * tier models and gpu.stackOffset (csv files and Philox streams keep the global stack index), one host thread per device
 * serves each call below, and there is no traffic between the devices in the step loop.  slgpu_group_run applies the
 * reference's stop rule to the total (serverwrapper.cpp:100-107); slgpu_group_rhat is EPSRDiagnostic::rHat over all
 * stacks of all devices (the sums of slgpu_rhat_stage1 / 2 added up in this process).  Everything per engine (drain,
 * counters, table, ...) goes through slgpu_group_engine(g, i). ---- */
typedef struct slgpu_group slgpu_group;
slgpu_status slgpu_group_create(const char* config_json, const char* plugin_path, const char* plugin_entry, const void* payload,
                                size_t payload_bytes, const int* devices, int n_devices, slgpu_group** out);
void slgpu_group_destroy(slgpu_group* g);
int slgpu_group_size(const slgpu_group* g);
slgpu_engine* slgpu_group_engine(slgpu_group* g, int i);
slgpu_status slgpu_group_init_chains(slgpu_group* g);
slgpu_status slgpu_group_step_rounds(slgpu_group* g, uint64_t n_rounds);
slgpu_status slgpu_group_sync(slgpu_group* g);
slgpu_status slgpu_group_wait(slgpu_group* g);
slgpu_status slgpu_group_run(slgpu_group* g);
uint64_t slgpu_group_rounds_done(const slgpu_group* g);
slgpu_status slgpu_group_rhat(slgpu_group* g, double* rhat);
slgpu_status slgpu_group_summary(slgpu_group* g, char* json, size_t cap, size_t* needed);
/* one checkpoint file per shard: <path>.stack<first global stack of the shard> */
slgpu_status slgpu_group_save_state(slgpu_group* g, const char* path);
slgpu_status slgpu_group_load_state(slgpu_group* g, const char* path);

/* contract check: nll[point*J + job] = f(job, x[point*d ..]) through the plugin's functor (host buffers) */
slgpu_status slgpu_eval_likelihood(slgpu_engine* e, const double* x, uint32_t n_points, double* nll);
/* one State row as CSVChainArrayWriter prints it (db.cpp:18-25), no newline; returns the length */
size_t slgpu_format_row(const double* row, uint32_t d, char* out, size_t cap);

/* measurement aid: sustained FP64 FMA throughput of the device (TFLOP/s, 2 flops per DFMA) from a
 * register-resident DFMA loop; the denominator of the fp64 roofline fraction bench.py reports */
slgpu_status slgpu_measure_fp64_peak(int device, double* tflops);

#ifdef __cplusplus
}
#endif
#endif
