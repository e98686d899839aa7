/*
 * oracle/ref_harness.cpp -- drives the REFERENCE'S OWN src/infer code, compiled unmodified from
 * /root/reference (oracle/Makefile.ref), in one process, so that the oracle's restatement can be pinned
 * against it.  TEST INFRASTRUCTURE ONLY: only tests/ (and the golden-fixture generator) load oracle/_ref.
 *
 * What is the reference's code and what is not, in oracle/_ref/libsl_ref.so:
 *   reference, unmodified : src/infer/sampler.cpp (bouncyBounds, GaussianProposal, Sampler::Sampler/step/
 *                           propose/unlock/flush), src/infer/chainarray.cpp (acceptProposal, acceptSwap,
 *                           ChainArray::append/swap/flushToDisk/...), src/infer/adaptive.cpp
 *                           (RegressionAdapter, ProposalShaper), src/db/db.cpp (CSV writer, row format),
 *                           src/infer/logging.cpp (TableLogger: counters and the printed table),
 *                           src/infer/diagnostics.cpp (EPSRDiagnostic),
 *                           and every header they include except Eigen and zmq.hpp.
 *   stand-ins (this repo) : Eigen  -> oracle/ref_shim/Eigen/Core (GEMV, norm and the 3x3 QR solve follow the
 *                           oracle's canonical orders; see its header);
 *                           zmq.hpp -> oracle/ref_shim/zmq.hpp (types only);
 *                           comms::Requester (src/comms/requester.cpp:27-57 needs ZeroMQ and the delegator /
 *                           worker processes) -> the in-process FIFO below: submit() queues (id, x),
 *                           retrieve() pops the oldest, evaluates the caller's likelihood for every job type
 *                           in order and returns (id, results); with `wire` set, x and every result take the
 *                           same "%f" text round trip as requester.cpp:37 / minion.cpp:53;
 *                           the init loop of runSampler (src/app/serverwrapper.cpp:48-87; that file drags in
 *                           the delegator, the HTTP API and the logger) is restated in slref_create(), with
 *                           the caller's pre-bounce start points instead of VectorXd::Random;
 *                           ApiResources::set (src/app/api.cpp:13-18, the HTTP side of that file is not built):
 *                           only so that logging.cpp links; TableLogger::updateApi is never called;
 *                           the three std::mt19937 are seeded from the caller's seed instead of
 *                           std::random_device, and every variate is recorded (oracle/ref_shim/ref_prelude.h).
 */
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <limits>
#include <memory>
#include <numeric>
#include <string>
#include <tuple>
#include <vector>
#include <fstream>
#include <sstream>
#include <iostream>
#include <functional>
#include <chrono>
#include <map>
#include <list>
#include <set>
#include <unordered_map>
#include <mutex>
#include <thread>
#include <algorithm>

#include <easylogging/easylogging++.h>
#include <json.hpp>
#include <Eigen/Dense>
#include <zmq.hpp>

/* the harness reads the private state of the reference's classes to compare it with the oracle's */
#define private public
#include "infer/sampler.hpp"
#include "infer/adaptive.hpp"
#include "infer/chainarray.hpp"
#include "infer/logging.hpp"
#undef private

INITIALIZE_EASYLOGGINGPP

extern "C" {
typedef double (*slref_lik_fn)(void* ctx, uint32_t job_type, const double* x, uint32_t d);

typedef struct slref_config {
  uint32_t n_stacks, n_temps, n_dims, n_job_types, swap_interval;
  double opt_accept, opt_swap;
  const double* min; /* [d] */
  const double* max; /* [d] */
  uint32_t seed;     /* what the three random_device calls return */
  int32_t wire;      /* 1: "%f" text round trip of x and of every result */
  const char* output_path;
  slref_lik_fn lik;   /* the likelihood, or NULL to use target_fn below */
  void* lik_ctx;
  /* alternative: a function of slo_target_eval's signature (oracle/sl_oracle.h) with its arguments */
  double (*target_fn)(int32_t target, const double* params, uint32_t d, uint32_t n_job_types, uint32_t job,
                      const double* x);
  int32_t target;
  const double* target_params;
  uint64_t max_rounds; /* capacity of the draw tables */
} slref_config;
}

namespace slref {

unsigned g_seed = 0;

struct Harness {
  uint32_t S = 0, T = 0, d = 0, J = 0, C = 0;
  bool wire = false;
  slref_lik_fn lik = nullptr;
  void* ctx = nullptr;
  double (*target_fn)(int32_t, const double*, uint32_t, uint32_t, uint32_t, const double*) = nullptr;
  int32_t target = 0;
  const double* target_params = nullptr;
  double eval(uint32_t j, const double* x) const {
    return lik ? lik(ctx, j, x, d) : target_fn(target, target_params, d, J, j, x);
  }
  bool running = false; /* false while the init loop evaluates start points */
  std::deque<std::pair<uint32_t, std::vector<double>>> fifo;
  /* attribution of the recorded variates */
  std::vector<double> pendingNormals;
  std::vector<uint64_t> nProposed, nAppended;
  uint32_t curId = 0;
  uint64_t curRound = 0;
  uint64_t maxRounds = 0;
  std::vector<double> z, umh, uswap; /* [round][C][d], [round][C], [round][C] */
  uint64_t unattributed = 0;
};
Harness* H = nullptr;
/* the reference's uniform distributions are function-local statics: they outlive a run */
const void* g_mhDist = nullptr;
const void* g_swapDist = nullptr;

void log_draw(int kind, const void* dist, double v) {
  if (!H) return;
  if (kind == 0) {
    H->pendingNormals.push_back(v);
    return;
  }
  /* acceptProposal's distribution is the first one ever used (the first append of a run draws it, all
   * targets of the tests being finite inside the bounds); acceptSwap's is the other one */
  if (!g_mhDist) g_mhDist = dist;
  else if (dist != g_mhDist && !g_swapDist) g_swapDist = dist;
  if (H->curRound >= H->maxRounds) { H->unattributed++; return; }
  const size_t at = (size_t)H->curRound * H->C + H->curId;
  if (dist == g_mhDist) H->umh[at] = v;
  else if (dist == g_swapDist) H->uswap[at] = v;
  else H->unattributed++;
}

double wire_round_trip(double v) {
  /* requester.cpp:37 std::to_string(double) == "%f"; the receiver parses with std::stod */
  return std::stod(std::to_string(v));
}

}  // namespace slref

/* ---- comms::Requester as an in-process FIFO (see the header comment) ---- */
namespace stateline {
namespace comms {

Socket::Socket(zmq::context_t& context, int type, const std::string& name, int)
    : socket_(context, type), name_(name) {}

Requester::Requester(zmq::context_t& context) : socket_(context, ZMQ_DEALER, "toDelegator") {}

void Requester::submit(uint id, const std::vector<uint>&, const Eigen::VectorXd& data) {
  slref::Harness& h = *slref::H;
  std::vector<double> x((size_t)data.size());
  for (uint i = 0; i < data.size(); i++) x[i] = data(i);
  if (h.running) {
    /* the d normals GaussianProposal::propose just drew belong to this chain's next evaluated proposal */
    const uint64_t r = h.nProposed[id]++;
    if (h.pendingNormals.size() == h.d && r < h.maxRounds)
      std::memcpy(&h.z[((size_t)r * h.C + id) * h.d], h.pendingNormals.data(), sizeof(double) * h.d);
    else
      h.unattributed++;
  }
  h.pendingNormals.clear();
  h.fifo.emplace_back(id, std::move(x));
}

std::pair<uint, std::vector<double>> Requester::retrieve() {
  slref::Harness& h = *slref::H;
  auto job = std::move(h.fifo.front());
  h.fifo.pop_front();
  const uint id = job.first;
  std::vector<double>& x = job.second;
  if (h.wire)
    for (auto& v : x) v = slref::wire_round_trip(v);
  std::vector<double> results(h.J);
  for (uint32_t j = 0; j < h.J; ++j) {
    double f = h.eval(j, x.data());
    if (h.wire) f = slref::wire_round_trip(f);
    results[j] = f;
  }
  if (h.running) {
    h.curId = id;
    h.curRound = h.nAppended[id]++;
  }
  return std::make_pair(id, results);
}

}  // namespace comms
}  // namespace stateline

namespace stateline {
void ApiResources::set(const std::string& name, json data) { resources_[name] = data.dump(); }
}  // namespace stateline

/* ---- the run object ---- */
struct slref_run {
  slref::Harness h;
  stateline::mcmc::ProposalBounds bounds;
  zmq::context_t context;
  std::unique_ptr<stateline::mcmc::RegressionAdapter> sigmaAdapter, betaAdapter;
  std::unique_ptr<stateline::mcmc::GaussianProposal> proposal;
  std::unique_ptr<stateline::mcmc::ChainArray> chains;
  std::unique_ptr<stateline::comms::Requester> requester;
  std::unique_ptr<stateline::mcmc::Sampler> sampler;
  std::unique_ptr<stateline::mcmc::TableLogger> logger;
  std::string lastTable;
  /* what runSampler's loop sees: the (id, state) pairs step() returns for cold chains, in order */
  std::vector<std::vector<double>> cold; /* [stack] -> rows of d+5 */
  uint64_t steps = 0;
};

static void push_row(std::vector<double>& v, const stateline::mcmc::State& s) {
  for (uint i = 0; i < s.sample.size(); i++) v.push_back(s.sample(i));
  v.push_back(s.energy);
  v.push_back(s.sigma);
  v.push_back(s.beta);
  v.push_back((double)s.accepted);
  v.push_back((double)(int)s.swapType);
}

extern "C" {

/* runSampler up to and including the Sampler constructor (serverwrapper.cpp:48-93).
 * x_init: [S*T][d] pre-bounce start points (what VectorXd::Random would have produced). */
slref_run* slref_create(const slref_config* cfg, const double* x_init) {
  using namespace stateline;
  static bool logging_quiet = false;
  if (!logging_quiet) {
    el::Configurations conf;
    conf.setToDefault();
    conf.setGlobally(el::ConfigurationType::Enabled, "false");
    conf.setGlobally(el::ConfigurationType::ToFile, "false");
    conf.setGlobally(el::ConfigurationType::ToStandardOutput, "false");
    el::Loggers::reconfigureAllLoggers(conf);
    logging_quiet = true;
  }
  if (slref::H) return nullptr; /* one run at a time: the Requester stand-in talks to a global */
  auto* r = new slref_run();
  slref::Harness& h = r->h;
  h.S = cfg->n_stacks; h.T = cfg->n_temps; h.d = cfg->n_dims; h.J = cfg->n_job_types;
  h.C = h.S * h.T;
  h.wire = cfg->wire != 0;
  h.lik = cfg->lik; h.ctx = cfg->lik_ctx;
  h.target_fn = cfg->target_fn; h.target = cfg->target; h.target_params = cfg->target_params;
  h.maxRounds = cfg->max_rounds;
  const double nan = std::numeric_limits<double>::quiet_NaN();
  h.z.assign((size_t)h.maxRounds * h.C * h.d, nan);
  h.umh.assign((size_t)h.maxRounds * h.C, nan);
  h.uswap.assign((size_t)h.maxRounds * h.C, nan);
  h.nProposed.assign(h.C, 0);
  h.nAppended.assign(h.C, 0);
  slref::g_seed = cfg->seed;
  slref::H = &h;

  r->bounds.min.resize(h.d);
  r->bounds.max.resize(h.d);
  for (uint32_t i = 0; i < h.d; ++i) { r->bounds.min[i] = cfg->min[i]; r->bounds.max[i] = cfg->max[i]; }

  /* serverwrapper.cpp:48-56 */
  const double max_log_ratio = 4.;
  const double min_log_ratio = -8.;
  const uint initial_count = 1000;
  r->sigmaAdapter.reset(new mcmc::RegressionAdapter(h.S, h.T, cfg->opt_accept, min_log_ratio, max_log_ratio));
  r->betaAdapter.reset(new mcmc::RegressionAdapter(h.S, h.T, cfg->opt_swap, 0., max_log_ratio));
  r->proposal.reset(new mcmc::GaussianProposal(h.S, h.T, h.d, r->bounds, initial_count));
  r->chains.reset(new mcmc::ChainArray(h.S, h.T, cfg->output_path));
  r->requester.reset(new comms::Requester(r->context));

  std::vector<uint> jobTypes(h.J);
  std::iota(jobTypes.begin(), jobTypes.end(), 0);

  /* serverwrapper.cpp:60-81 with generateInitialSample (:21-41) */
  for (uint i = 0; i < h.C; i++) {
    Eigen::VectorXd raw(h.d);
    for (uint32_t k = 0; k < h.d; ++k) raw[k] = x_init[(size_t)i * h.d + k];
    Eigen::VectorXd sample = mcmc::bouncyBounds(raw, r->bounds.min, r->bounds.max);
    r->requester->submit(0, jobTypes, sample);
    auto result = r->requester->retrieve();
    double energy = std::accumulate(std::begin(result.second), std::end(result.second), 0.0);
    if (i % h.T == 0) r->betaAdapter->computeBetaStack(i);
    r->chains->initialise(i, sample, energy, r->sigmaAdapter->values()[i], r->betaAdapter->values()[i]);
  }
  r->cold.assign(h.S, {});
  for (uint32_t s = 0; s < h.S; ++s) push_row(r->cold[s], r->chains->lastState(s * h.T));

  h.running = true;
  r->sampler.reset(new mcmc::Sampler(*r->requester, jobTypes, *r->chains, *r->proposal, *r->sigmaAdapter,
                                     *r->betaAdapter, cfg->swap_interval));
  /* serverwrapper.cpp:95; the refresh period is "never": slref_step_print forces the print */
  r->logger.reset(new mcmc::TableLogger(h.S, h.T, h.d, 0x7fffffff));
  return r;
}

/* the main loop of runSampler (serverwrapper.cpp:100-107): n_steps calls of Sampler::step() */
void slref_step(slref_run* r, uint64_t n_steps) {
  uint id;
  stateline::mcmc::State state;
  for (uint64_t i = 0; i < n_steps; ++i) {
    std::tie(id, state) = r->sampler->step();
    if (id % r->h.T == 0) push_row(r->cold[id / r->h.T], state);
    /* serverwrapper.cpp:124-126 */
    r->logger->update(id, state, r->sigmaAdapter->values(), r->sigmaAdapter->rates(), r->betaAdapter->values(),
                      r->betaAdapter->rates());
    r->steps++;
  }
}

/* One more step, on which TableLogger::update finds its refresh period expired and prints the table and the
 * convergence line (logging.cpp:50-87) to std::cout, captured here.  Returns the text length. */
size_t slref_step_print(slref_run* r, char* out, size_t cap) {
  std::ostringstream capture;
  std::streambuf* old = std::cout.rdbuf(capture.rdbuf());
  r->logger->msRefresh_ = 0;
  r->logger->lastPrintTime_ = std::chrono::steady_clock::time_point();
  slref_step(r, 1);
  r->logger->msRefresh_ = 0x7fffffff;
  std::cout.rdbuf(old);
  r->lastTable = capture.str();
  if (out && cap) {
    const size_t n = std::min(cap - 1, r->lastTable.size());
    std::memcpy(out, r->lastTable.data(), n);
    out[n] = 0;
  }
  return r->lastTable.size();
}

/* TableLogger's counters and EPSR (logging.cpp:29-48, diagnostics.cpp:49-71) */
void slref_get_logger(const slref_run* r, double* length, double* min_energy, double* cur_energy, double* accepts,
                      double* swaps, double* swap_attempts, double* rhat) {
  const auto& lg = *r->logger;
  for (uint32_t c = 0; c < r->h.C; ++c) {
    length[c] = lg.lengths_[c]; min_energy[c] = lg.minEnergies_[c]; cur_energy[c] = lg.energies_[c];
    accepts[c] = lg.nAcceptsGlobal_[c]; swaps[c] = lg.nSwapsGlobal_[c]; swap_attempts[c] = lg.nSwapAttemptsGlobal_[c];
  }
  Eigen::ArrayXd rh = lg.diagnostic_.rHat();
  for (uint32_t i = 0; i < r->h.d; ++i) rhat[i] = rh(i);
}

void slref_flush(slref_run* r) { r->sampler->flush(); }

/* destroys the sampler and the chain array (whose destructor writes the CSV files) */
void slref_destroy(slref_run* r) {
  if (!r) return;
  r->sampler.reset();
  r->chains.reset();
  slref::H = nullptr;
  delete r;
}

uint64_t slref_unattributed(const slref_run* r) { return r->h.unattributed; }

/* recorded variates, laid out as the oracle's replay tables (sl_oracle.h slo_set_replay) */
void slref_get_draws(const slref_run* r, uint64_t n_rounds, double* z, double* u_mh, double* u_swap) {
  const slref::Harness& h = r->h;
  std::memcpy(z, h.z.data(), sizeof(double) * (size_t)n_rounds * h.C * h.d);
  std::memcpy(u_mh, h.umh.data(), sizeof(double) * (size_t)n_rounds * h.C);
  std::memcpy(u_swap, h.uswap.data(), sizeof(double) * (size_t)n_rounds * h.C);
}

/* rows[S*T][d+5]: ChainArray::lastState of every chain */
void slref_get_last_rows(const slref_run* r, double* rows) {
  const slref::Harness& h = r->h;
  std::vector<double> v;
  for (uint32_t c = 0; c < h.C; ++c) push_row(v, r->chains->lastState(c));
  std::memcpy(rows, v.data(), sizeof(double) * v.size());
}

void slref_get_sigma_beta(const slref_run* r, double* sigma, double* beta) {
  for (uint32_t c = 0; c < r->h.C; ++c) { sigma[c] = r->chains->sigma(c); beta[c] = r->chains->beta(c); }
}

/* ProposalShaper state of one chain, dense row-major */
void slref_get_shaper(const slref_run* r, uint32_t chain, double* L, double* N, double* count) {
  const auto& sh = r->proposal->proposalShape_;
  const uint32_t d = r->h.d;
  for (uint32_t i = 0; i < d; ++i)
    for (uint32_t k = 0; k < d; ++k) {
      L[(size_t)i * d + k] = sh.L_[chain](i, k);
      N[(size_t)i * d + k] = sh.N_[chain](i, k);
    }
  *count = (double)sh.count_[chain];
}

/* tier models of an adapter (which: 0 sigma, 1 beta): mu_xx [T][9] row-major, mu_xy [T][3], weight [T][3],
 * count [T]; values / rates per chain */
void slref_get_models(const slref_run* r, int which, double* mu_xx, double* mu_xy, double* weight, double* count,
                      double* values, double* rates) {
  const auto& a = which ? *r->betaAdapter : *r->sigmaAdapter;
  for (uint32_t t = 0; t < r->h.T; ++t) {
    for (int i = 0; i < 3; ++i) {
      for (int k = 0; k < 3; ++k) mu_xx[t * 9 + i * 3 + k] = a.mu_xx_[t](i, k);
      mu_xy[t * 3 + i] = a.mu_xy_[t](i);
      weight[t * 3 + i] = a.weight_[t](i);
    }
    count[t] = a.count_[t];
  }
  for (uint32_t c = 0; c < r->h.C; ++c) { values[c] = a.values_[c]; rates[c] = a.rates_[c]; }
}

uint64_t slref_cold_rows(const slref_run* r, uint32_t stack, double* rows, uint64_t cap_rows) {
  const auto& v = r->cold[stack];
  const uint64_t n = v.size() / (r->h.d + 5);
  if (rows) std::memcpy(rows, v.data(), sizeof(double) * (size_t)std::min(n, cap_rows) * (r->h.d + 5));
  return n;
}

uint32_t slref_length(const slref_run* r, uint32_t chain) { return r->chains->length(chain); }

/* stand-alone pieces */
/* EPSRDiagnostic (src/infer/diagnostics.cpp) as the reference's own tests drive it (src/test/diagnostics.cpp):
 * samples[n][m][d] fed round-robin to m cold chains (nChains = 1), rhat[d] out */
void slref_epsr(const double* samples, uint32_t n, uint32_t m, uint32_t d, double* rhat, int32_t* converged) {
  stateline::mcmc::EPSRDiagnostic epsr(m, 1, d);
  for (uint32_t i = 0; i < n; ++i)
    for (uint32_t j = 0; j < m; ++j) {
      stateline::mcmc::State state;
      state.sample.resize(d);
      for (uint32_t k = 0; k < d; ++k) state.sample[k] = samples[((size_t)i * m + j) * d + k];
      epsr.update(j, state);
    }
  Eigen::ArrayXd rh = epsr.rHat();
  for (uint32_t k = 0; k < d; ++k) rhat[k] = rh(k);
  if (converged) *converged = epsr.hasConverged() ? 1 : 0;
}

void slref_bouncy_bounds(const double* val, const double* mn, const double* mx, uint32_t d, double* out) {
  Eigen::VectorXd v(d), a(d), b(d);
  for (uint32_t i = 0; i < d; ++i) { v[i] = val[i]; a[i] = mn[i]; b[i] = mx[i]; }
  Eigen::VectorXd o = stateline::mcmc::bouncyBounds(v, a, b);
  for (uint32_t i = 0; i < d; ++i) out[i] = o[i];
}

}  // extern "C"
