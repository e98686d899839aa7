/*
 * oracle/ref_full.cpp -- the reference's COMPLETE CPU path, as its own binaries run it: stateline::ServerWrapper
 * (delegator + sampler + heartbeats over real ZeroMQ sockets, src/app/serverwrapper.cpp, src/comms/*) and N
 * stateline::WorkerWrapper instances (src/app/workerwrapper.cpp: worker + minion threads each), every reference source
 * compiled unmodified from /root/reference (oracle/Makefile.full) and linked against the libzmq shared library that
 * ships with pyzmq in this image.  BENCH INFRASTRUCTURE ONLY (bench.py: cpu_baseline.full_server_workers).
 *
 * Stand-ins: Eigen (oracle/ref_shim/Eigen/Core), libzmq's C header (oracle/ref_shim/zmq_real/zmq.h, declarations only),
 * and this main(), which does what src/bin/stateline.cpp and src/bin/demo-worker.cpp do in one process: build the
 * settings from the config JSON, start the server, start the workers with a likelihood, wait until the server has
 * collected nSamplesTotal cold-chain samples.
 *
 *   ref_full <config.json> <port> <n_workers> <target> <n_dims> [params.bin]
 * prints one line: "<seconds> <cold samples> <chains>"  (seconds from ServerWrapper::start() to the end of the run)
 */
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include <easylogging/easylogging++.h>
#include <json.hpp>

#include "app/serverwrapper.hpp"
#include "app/workerwrapper.hpp"

INITIALIZE_EASYLOGGINGPP

namespace slref {
unsigned g_seed = 0;
void log_draw(int, const void*, double) {}
}  // namespace slref

extern "C" double slo_target_eval(int32_t target, const double* params, uint32_t d, uint32_t n_job_types, uint32_t job,
                                  const double* x);

int main(int argc, char** argv) {
  if (argc < 6) {
    std::fprintf(stderr, "usage: ref_full config.json port n_workers target n_dims [params.bin]\n");
    return 2;
  }
  if (!std::getenv("SLREF_LOG")) {
    el::Configurations conf;
    conf.setToDefault();
    conf.setGlobally(el::ConfigurationType::Enabled, "false");
    conf.setGlobally(el::ConfigurationType::ToFile, "false");
    conf.setGlobally(el::ConfigurationType::ToStandardOutput, "false");
    el::Loggers::reconfigureAllLoggers(conf);
  }
  std::ifstream f(argv[1]);
  nlohmann::json config;
  f >> config;
  const int port = std::atoi(argv[2]);
  const int n_workers = std::atoi(argv[3]);
  const int target = std::atoi(argv[4]);
  const uint32_t d = (uint32_t)std::atoi(argv[5]);
  std::vector<double> params;
  if (argc > 6) {
    std::ifstream pf(argv[6], std::ios::binary);
    double v;
    while (pf.read(reinterpret_cast<char*>(&v), sizeof v)) params.push_back(v);
  }
  slref::g_seed = 12345u;
  stateline::StatelineSettings settings = stateline::StatelineSettings::fromJSON(config);
  const uint32_t J = settings.nJobTypes;
  const double* prm = params.empty() ? nullptr : params.data();
  stateline::LikelihoodFn lik = [=](uint job, const std::vector<double>& x) {
    return slo_target_eval(target, prm, d, J, job, x.data());
  };

  const auto t0 = std::chrono::steady_clock::now();
  stateline::ServerWrapper server(port, settings);
  server.start();
  /* workers are separate processes started after the server in a real deployment: give the delegator time to bind */
  std::this_thread::sleep_for(std::chrono::milliseconds(300));
  std::vector<std::unique_ptr<stateline::WorkerWrapper>> workers;
  const std::string address = "localhost:" + std::to_string(port);
  for (int i = 0; i < n_workers; ++i) {
    workers.emplace_back(new stateline::WorkerWrapper(lik, {0u, J}, address));
    workers.back()->start();
    std::this_thread::sleep_for(std::chrono::milliseconds(50));
  }
  while (server.isRunning()) std::this_thread::sleep_for(std::chrono::milliseconds(2));
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::printf("%.6f %u %u\n", secs, settings.nsamples, settings.nstacks * settings.ntemps);
  std::fflush(stdout);
  /* the reference's shutdown (context teardown with sockets still open in worker threads) can block: the measurement
   * is done, leave without running destructors */
  std::_Exit(0);
}
