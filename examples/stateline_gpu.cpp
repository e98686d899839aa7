// examples/stateline_gpu.cpp -- what `stateline -c config.json` (src/bin/stateline.cpp:52-90) becomes on
// the GPU path: read the same config, load a likelihood plugin, run to nSamplesTotal, print a summary.
//
//   g++ -std=c++17 -O2 -I<repo>/include stateline_gpu.cpp -L<repo>/stateline_b200 -lslgpu -Wl,-rpath,<repo>/stateline_b200 -o stateline-gpu
//   ./stateline-gpu -c demo-config.json -p <repo>/stateline_b200/libslgpu_targets.so [-e slgpu_plugin_demo_gauss]
//
// With "gpu": {"sink": "csv"} in the config the output is <outputPath>/<stack>.csv exactly as the reference writes it.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "slgpu.h"

int main(int argc, char** argv) {
  std::string config_path, plugin, entry;
  for (int i = 1; i + 1 < argc; i += 2) {
    if (!strcmp(argv[i], "-c")) config_path = argv[i + 1];
    else if (!strcmp(argv[i], "-p")) plugin = argv[i + 1];
    else if (!strcmp(argv[i], "-e")) entry = argv[i + 1];
  }
  if (config_path.empty() || plugin.empty()) {
    std::fprintf(stderr, "usage: %s -c config.json -p plugin.so [-e entry_symbol]\n", argv[0]);
    return 2;
  }
  std::ifstream f(config_path);
  if (!f.good()) {
    std::fprintf(stderr, "cannot read %s\n", config_path.c_str());
    return 2;
  }
  std::stringstream ss;
  ss << f.rdbuf();
  slgpu_engine* e = nullptr;
  if (slgpu_create(ss.str().c_str(), plugin.c_str(), entry.empty() ? nullptr : entry.c_str(), nullptr, 0, &e) != SLGPU_OK) {
    std::fprintf(stderr, "slgpu_create: %s\n", slgpu_last_error());  // the reference would exit(EXIT_FAILURE) here
    return 1;
  }
  if (slgpu_run(e) != SLGPU_OK) {
    std::fprintf(stderr, "slgpu_run: %s\n", slgpu_last_error());
    slgpu_destroy(e);
    return 1;
  }
  size_t need = 0;
  slgpu_summary(e, nullptr, 0, &need);
  std::vector<char> buf(need);
  slgpu_summary(e, buf.data(), buf.size(), nullptr);
  std::printf("%s\n", buf.data());
  slgpu_destroy(e);
  return 0;
}
