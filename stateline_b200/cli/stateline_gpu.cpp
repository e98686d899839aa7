// stateline-gpu -- the `stateline` server binary (src/bin/stateline.cpp:26-90) for the GPU path.
//
// Same command line (-h, -l/--log-level, -p/--port, -c/--config with the same defaults) and the same config
// file; instead of serving likelihood jobs to workers over ZeroMQ it loads a likelihood plugin (--plugin /
// --entry, or "gpu": {"plugin": ..., "entry": ...} in the config) and runs every chain on the device.  Like the
// reference it writes <outputPath>/<stack>.csv (sink "csv" unless the config says otherwise), prints the
// TableLogger table every loggingRateSec (log level >= 0), stops at nSamplesTotal cold samples or on Ctrl+C, and
// flushes before it exits.  Beyond the reference: --checkpoint FILE (written at exit) and --resume FILE.
//
// This file only talks to the C-ABI of include/slgpu.h; it is also the example of a C++ host.
#include <chrono>
#include <csignal>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "json_min.hpp"
#include "slgpu.h"

namespace {

volatile std::sig_atomic_t g_interrupted = 0;  // init::initialiseSignalHandler, src/app/signal.cpp
void on_signal(int) { g_interrupted = 1; }

struct Options {
  std::string config = "config.json";  // stateline.cpp:34
  int log_level = 0;                   // stateline.cpp:32
  int port = 5555;                     // stateline.cpp:33; accepted for compatibility, nothing listens
  std::string plugin, entry, payload, checkpoint, resume;
  int gpus = 1;  // devices this one server process drives (0 = every visible device)
  bool help = false;
};

void usage(const char* argv0) {
  std::printf(
      "Usage: %s [-c config.json] [--plugin likelihood.so] [--entry symbol] [options]\n\n"
      "  -h, --help            Print help message\n"
      "  -l, --log-level N     Logging level (0: table every loggingRateSec; negative: quiet)\n"
      "  -p, --port N          Accepted for compatibility with stateline; no workers connect to the GPU path\n"
      "  -c, --config FILE     Path to configuration file (default config.json)\n"
      "      --plugin FILE     Likelihood plugin .so (default: config \"gpu\".\"plugin\")\n"
      "      --entry SYMBOL    Plugin entry symbol (default: config \"gpu\".\"entry\", else slgpu_plugin_entry)\n"
      "      --payload FILE    Raw bytes handed to the plugin's functor (e.g. an array of doubles)\n"
      "      --gpus N          Shard the stacks over N devices, gpu.device upwards (0 = all visible; default 1)\n"
      "      --checkpoint FILE Save the sampler state there on exit (also after Ctrl+C)\n"
      "      --resume FILE     Continue a run from a checkpoint instead of initialising the chains\n",
      argv0);
}

bool parse_args(int argc, char** argv, Options& o) {
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto value = [&](std::string& dst) {
      if (i + 1 >= argc) return false;
      dst = argv[++i];
      return true;
    };
    std::string v;
    if (a == "-h" || a == "--help") o.help = true;
    else if (a == "-l" || a == "--log-level") { if (!value(v)) return false; o.log_level = std::atoi(v.c_str()); }
    else if (a == "-p" || a == "--port") { if (!value(v)) return false; o.port = std::atoi(v.c_str()); }
    else if (a == "-c" || a == "--config") { if (!value(o.config)) return false; }
    else if (a == "--plugin") { if (!value(o.plugin)) return false; }
    else if (a == "--entry") { if (!value(o.entry)) return false; }
    else if (a == "--payload") { if (!value(o.payload)) return false; }
    else if (a == "--gpus") { if (!value(v)) return false; o.gpus = std::atoi(v.c_str()); }
    else if (a == "--checkpoint") { if (!value(o.checkpoint)) return false; }
    else if (a == "--resume") { if (!value(o.resume)) return false; }
    else return false;
  }
  return true;
}

bool read_file(const std::string& path, std::string& out) {
  std::ifstream f(path, std::ios::binary);
  if (!f.good()) return false;
  std::stringstream ss;
  ss << f.rdbuf();
  out = ss.str();
  return true;
}

// Add "key": value to the config's "gpu" object (creating it) unless the key is already there.
std::string with_gpu_default(const std::string& text, const slgpu_json::Value& cfg, const char* key, const char* json_value) {
  const slgpu_json::Value* gpu = cfg.find("gpu");
  if (gpu && gpu->find(key)) return text;
  const std::string item = std::string("\"") + key + "\": " + json_value;
  if (!gpu) {
    const size_t close = text.rfind('}');
    return text.substr(0, close) + ", \"gpu\": {" + item + "}" + text.substr(close);
  }
  // the "gpu" key itself: the first "gpu" that is followed by ':' and '{'
  size_t at = std::string::npos;
  for (size_t k = text.find("\"gpu\""); k != std::string::npos; k = text.find("\"gpu\"", k + 1)) {
    size_t c = text.find_first_not_of(" \t\r\n", k + 5);
    if (c == std::string::npos || text[c] != ':') continue;
    c = text.find_first_not_of(" \t\r\n", c + 1);
    if (c != std::string::npos && text[c] == '{') {
      at = c;
      break;
    }
  }
  if (at == std::string::npos) return text;
  return text.substr(0, at + 1) + item + (gpu->obj.empty() ? "" : ", ") + text.substr(at + 1);
}

int die(slgpu_group* g, const char* what) {
  std::fprintf(stderr, "%s: %s\n", what, slgpu_last_error());  // LOG(FATAL) in the reference
  if (g) slgpu_group_destroy(g);
  return 1;
}

// the reference prints one table for all chains (logging.cpp:52-87); a group prints one per device, chain ids local to it
void print_table(slgpu_group* g) {
  for (int i = 0; i < slgpu_group_size(g); ++i) {
    slgpu_engine* e = slgpu_group_engine(g, i);
    size_t need = 0;
    if (slgpu_format_table(e, nullptr, 0, &need) != SLGPU_OK) return;
    std::vector<char> buf(need);
    if (slgpu_group_size(g) > 1) std::printf("\n[device shard %d]", i);
    if (slgpu_format_table(e, buf.data(), buf.size(), nullptr) == SLGPU_OK) std::fputs(buf.data(), stdout);
  }
  std::fflush(stdout);
}

}  // namespace

int main(int argc, char** argv) {
  Options opt;
  if (!parse_args(argc, argv, opt) || opt.help) {
    usage(argv[0]);
    return opt.help ? 0 : 2;
  }
  std::string text;
  if (!read_file(opt.config, text)) {
    std::fprintf(stderr, "Could not find config file\n");  // stateline.cpp:44
    return 1;
  }
  slgpu_json::Value cfg;
  try {
    cfg = slgpu_json::parse(text);
  } catch (const std::exception& ex) {
    std::fprintf(stderr, "%s: %s\n", opt.config.c_str(), ex.what());
    return 1;
  }
  if (cfg.kind != slgpu_json::Value::Object) {
    std::fprintf(stderr, "%s: config must be a JSON object\n", opt.config.c_str());
    return 1;
  }
  const slgpu_json::Value* gpu = cfg.find("gpu");
  auto gpu_str = [&](const char* key) {
    const slgpu_json::Value* v = gpu ? gpu->find(key) : nullptr;
    return v && v->kind == slgpu_json::Value::String ? v->str : std::string();
  };
  if (opt.plugin.empty()) opt.plugin = gpu_str("plugin");
  if (opt.entry.empty()) opt.entry = gpu_str("entry");
  if (opt.plugin.empty()) {
    std::fprintf(stderr, "no likelihood plugin: pass --plugin FILE or set \"gpu\": {\"plugin\": ...} (there is no CPU worker path)\n");
    return 2;
  }
  const bool table = opt.log_level >= 0;
  // the reference always writes csv files; the table's windowed-rate columns need the per-round log
  text = with_gpu_default(text, cfg, "sink", "\"csv\"");
  cfg = slgpu_json::parse(text);
  if (table) {
    text = with_gpu_default(text, cfg, "windowRates", "true");
    cfg = slgpu_json::parse(text);
  }
  if (!opt.resume.empty()) {
    text = with_gpu_default(text, cfg, "append", "true");
    cfg = slgpu_json::parse(text);
  }
  std::string payload;
  if (!opt.payload.empty() && !read_file(opt.payload, payload)) {
    std::fprintf(stderr, "cannot read payload %s\n", opt.payload.c_str());
    return 1;
  }

  std::signal(SIGINT, on_signal);
  std::signal(SIGTERM, on_signal);

  // one process owns the whole run, as `stateline -c cfg` does (stateline.cpp:52-90): a group of --gpus devices
  slgpu_group* g = nullptr;
  if (slgpu_group_create(text.c_str(), opt.plugin.c_str(), opt.entry.empty() ? nullptr : opt.entry.c_str(),
                         payload.empty() ? nullptr : payload.data(), payload.size(), nullptr, opt.gpus, &g) != SLGPU_OK)
    return die(nullptr, "slgpu_group_create");
  if (opt.resume.empty() ? slgpu_group_init_chains(g) != SLGPU_OK : slgpu_group_load_state(g, opt.resume.c_str()) != SLGPU_OK)
    return die(g, opt.resume.empty() ? "slgpu_group_init_chains" : "slgpu_group_load_state");

  // stop rule, src/app/serverwrapper.cpp:100-107: cold-chain samples over all stacks >= nSamplesTotal
  const double n_stacks = cfg.find("nStacks")->num, n_total = cfg.find("nSamplesTotal")->num;
  const double log_sec = cfg.find("loggingRateSec") ? cfg.find("loggingRateSec")->num : 1.0;
  const uint64_t total_rounds = (uint64_t)((n_total + n_stacks - 1) / n_stacks);
  using clock = std::chrono::steady_clock;
  auto last_print = clock::now();
  uint64_t chunk = 16;
  while (!g_interrupted && slgpu_group_rounds_done(g) < total_rounds) {
    const uint64_t n = std::min<uint64_t>(chunk, total_rounds - slgpu_group_rounds_done(g));
    const auto t0 = clock::now();
    if (slgpu_group_step_rounds(g, n) != SLGPU_OK) return die(g, "slgpu_group_step_rounds");
    if (slgpu_group_wait(g) != SLGPU_OK) return die(g, "slgpu_group_wait");  // no sink flush per chunk: that is for the end
    const double ms = std::chrono::duration<double, std::milli>(clock::now() - t0).count();
    if (ms < 100.0 && chunk < (1u << 20)) chunk *= 2;  // keep Ctrl+C and the table responsive: ~0.1-0.2 s per chunk
    if (table && std::chrono::duration<double>(clock::now() - last_print).count() > log_sec) {
      last_print = clock::now();
      print_table(g);
    }
  }
  if (slgpu_group_sync(g) != SLGPU_OK) return die(g, "slgpu_group_sync");  // Sampler::flush, sampler.cpp:216-235
  if (table) print_table(g);
  if (!opt.checkpoint.empty() && slgpu_group_save_state(g, opt.checkpoint.c_str()) != SLGPU_OK) return die(g, "slgpu_group_save_state");
  size_t need = 0;
  std::vector<char> buf;
  if (slgpu_group_size(g) == 1) {  // the single-device summary as before
    slgpu_engine* e = slgpu_group_engine(g, 0);
    slgpu_summary(e, nullptr, 0, &need);
    buf.resize(need);
    if (slgpu_summary(e, buf.data(), buf.size(), nullptr) == SLGPU_OK) std::printf("%s\n", buf.data());
  } else {
    slgpu_group_summary(g, nullptr, 0, &need);
    buf.resize(need);
    if (slgpu_group_summary(g, buf.data(), buf.size(), nullptr) == SLGPU_OK) std::printf("%s\n", buf.data());
  }
  slgpu_group_destroy(g);
  return g_interrupted ? 130 : 0;
}
