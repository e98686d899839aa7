// engine.cu -- host side of the B200 parallel-tempering engine behind include/slgpu.h.
//
// Replaces, for the GPU path, the construction / initialisation / main-loop half of
// runSampler (src/app/serverwrapper.cpp:44-138) and the ChainArray -> CSVChainArrayWriter sink
// (src/infer/chainarray.cpp:109-155, src/db/db.cpp:27-47).  The step loop itself is the plugin-instantiated
// kernels of include/slgpu_device.cuh; the tier models are csrc/tier_kernels.cuh.
//
// Launch schedule ("batched" schedule): the tier models are frozen between adapt points.
// Round r (0-based) makes the chain length r+2; a swap sweep happens when (r+2) % swapInterval == 0 and
// an adapt point follows the round with (r+2) % adaptInterval == 0.  step_rounds() cuts the requested
// rounds into launches that end at adapt points (or after maxRoundsPerLaunch rounds), each followed,
// at an adapt point, by k_tier_adapt.  Cold-chain rows of a launch land in one of kSegments
// device segment buffers; with sink "host"/"csv" a second stream copies each finished segment into a
// pinned host ring while the next launch runs.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>  // header-only; ranges cost nothing unless a profiler is attached
#include <dlfcn.h>
#include <sys/stat.h>

#include <algorithm>
#include <charconv>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <deque>
#include <fstream>
#include <limits>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#include "json_min.hpp"
#include "slgpu.h"
#include "slgpu_plugin.h"
#include "row_compact.cuh"
#include "tier_kernels.cuh"

// NVTX ranges on the host-side phases of the step loop (nsys / ncu --nvtx timelines)
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

namespace {

thread_local std::string g_err;

slgpu_status fail(slgpu_status st, const std::string& msg) {
  g_err = msg;
  return st;
}

#define CU_TRY(expr)                                                                         \
  do {                                                                                       \
    cudaError_t cu_try_err = (expr);                                                         \
    if (cu_try_err != cudaSuccess)                                                           \
      return fail(SLGPU_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(cu_try_err)); \
  } while (0)

constexpr int kSegments = 3;
constexpr int kCompactBufs = kSegments + 1;  // see enqueue_compact: a buffer is reused only after the host has seen its segment land
enum Sink { SINK_NONE, SINK_DEVICE, SINK_HOST, SINK_CSV, SINK_BIN };
// what crosses PCIe for a host-side sink: every cold row in full, or (default) the delta stream -- one flag byte per row
// and the d + 3 doubles of a row only when the cold state changed (k_compact_rows)
enum Wire { WIRE_DELTA, WIRE_FULL };

struct Settings {
  // src/app/serverwrapper.hpp:64-91
  uint32_t nStacks = 0, nTemps = 0, swapInterval = 0, nJobTypes = 0, nDims = 0;
  uint64_t nSamplesTotal = 0;
  double loggingRateSec = 0, optimalAcceptRate = 0, optimalSwapRate = 0;
  std::string outputPath;
  bool useInitial = false;
  std::vector<double> initial, mn, mx;
  // "gpu" block
  int device = 0;
  uint64_t seed = 1234;
  uint32_t stackOffset = 0;
  int rng = SLGPU_RNG_PHILOX;
  Sink sink = SINK_DEVICE;
  Wire wire = WIRE_DELTA;
  uint32_t adaptInterval = 0, maxRoundsPerLaunch = 64;
  uint64_t hostRingRounds = 0;
  bool overwrite = false;
  bool sequential = false;   // "gpu": {"schedule": "sequential"}: the reference's own update order (validation mode)
  bool windowRates = false;  // keep the adapters' windowed accept / swap rates (the AcptRt / SwapRt columns of the table logger)
  uint32_t csvThreads = 0;  // csv sink: formatter threads (0 = one per host core, at most 16)
  bool append = false;  // csv sink: keep existing <stack>.csv files (a resumed run) instead of truncating them
  uint32_t recholEvery = 1;  // "shaper": "welford": adapt points between two re-Cholesky passes
  bool welford = false;  // "shaper": "welford": second-moment matrix + periodic re-Cholesky (NOT the reference's update; opt-in)
};

struct Segment {
  double* dev = nullptr;
  cudaEvent_t written = nullptr, copied = nullptr;
  bool copy_pending = false;
};

struct PendingCopy {
  cudaEvent_t ev;
  uint64_t rows_end;
};

// ---- the delta stream (wire "delta") -------------------------------------------------------------------------------
// On a reject ChainArray::append re-pushes a copy of the last state (src/infer/chainarray.cpp:102-106): about three
// cold rows in four repeat the row before them in everything but the two flags.  After every launch k_compact_rows
// turns the launch's full rows (device segment) into a compact segment, which is what the copy engine moves:
//   header  { u32 n_rounds, n_stacks, count, key }            count = payload rows
//   base    u32 [n_rounds][n_chunks]                          payload slot of the first changed row of a chunk of 256 stacks
//   flags   u8  [n_rounds][n_stacks]                          bit 0 accepted, bits 1-2 swapType, bit 3 changed
//   payload f64 [count][d + 3]                                x, energy, sigma, beta of the changed rows, in (round, chunk, stack) order within a chunk
// "changed" = any of the d + 3 doubles differs bitwise from the same stack's previous row (the previous launch's last
// round for round 0; a key segment marks every row of its first round changed).  The host rebuilds the exact d + 5
// doubles of every row (materialise), or hands the segment out as it is (slgpu_next_segment).
struct SegLayout {
  uint32_t n_chunks = 0;
  size_t off_base = 16, off_flags = 0, off_payload = 0;
  size_t bytes(uint32_t payload_rows, uint32_t d) const { return off_payload + (size_t)payload_rows * (d + 3) * sizeof(double); }
};
inline SegLayout seg_layout(uint32_t n_rounds, uint32_t S) {
  SegLayout L;
  L.n_chunks = (S + 255u) / 256u;
  L.off_flags = (16 + (size_t)n_rounds * L.n_chunks * 4 + 15) / 16 * 16;
  L.off_payload = (L.off_flags + (size_t)n_rounds * S + 15) / 16 * 16;
  return L;
}
struct CompactBuf {
  unsigned char* dev = nullptr;
  cudaEvent_t written = nullptr, copied = nullptr;
  bool copy_pending = false;
};
struct HostSeg {  // one compact segment in the pinned byte ring, oldest first
  size_t off = 0, reserved = 0;   // region of the byte ring
  uint32_t n_rounds = 0, cap_copied = 0;  // payload rows covered by the first copy
  uint64_t row_begin = 0;         // global index of its first row
  const unsigned char* dev = nullptr;  // where the rest of the payload would come from
  cudaEvent_t ev = nullptr;       // the copy (or the remainder copy) has landed
  bool landed = false, remainder = false;
};

// one device array of the sampler state, as written to / read from a checkpoint
struct StateBuf {
  const char* name;
  void* dev;
  size_t bytes;
};


// Formatter pool of the csv sink.  Rows arrive round-major (row r belongs to stack r % S); worker w owns the stacks
// s with s % n == w -- their text buffers and their files -- so a batch needs no locking beyond the fork / join.
struct CsvPool {
  std::vector<std::thread> threads;
  std::mutex m;
  std::condition_variable go, done;
  uint64_t generation = 0;
  uint32_t running = 0;
  bool stop = false, flush = false, io_error = false;
  uint64_t r0 = 0, r1 = 0;
  std::vector<size_t> buffered;  // bytes held per worker
};

}  // namespace

struct slgpu_engine {
  Settings cfg;
  void* dl = nullptr;
  const slgpu_plugin_v1* plug = nullptr;
  cudaStream_t compute = nullptr, drain = nullptr;
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  bool timing = false;                 // between slgpu_timer_start and slgpu_timer_stop
  std::vector<cudaEvent_t> step_ev;    // (start, stop) pairs around every step-kernel launch of the timed region
  std::vector<cudaEvent_t> ev_time_pool;  // timing events kept for reuse
  double step_kernel_ms = 0.0;
  uint64_t step_kernel_launches = 0;
  std::vector<void*> allocs;
  std::vector<StateBuf> state;         // everything a resumed run needs (slgpu_save_state / slgpu_load_state)
  slgpu_step_params P{};
  double *d_models = nullptr, *d_sig_w = nullptr, *d_ladder = nullptr, *d_sums = nullptr;
  double* d_beta_values = nullptr;  // sequential schedule: RegressionAdapter::values_ of the beta adapter, [C]
  void* d_payload = nullptr;
  unsigned int* d_ticket = nullptr;    // last-block ticket of k_tier_adapt
  unsigned char* d_round_flags = nullptr;  // windowRates: [seg_rounds][C] per-round accept / swap flags of a launch
  uint32_t *d_win_sig = nullptr, *d_win_beta = nullptr, *d_win_cnt = nullptr, *d_win_sum = nullptr;
  double *d_rp_z = nullptr, *d_rp_umh = nullptr, *d_rp_uswap = nullptr, *d_rp_xinit = nullptr;
  bool initialised = false, failed = false;
  uint64_t rounds_done = 0, launches = 0, rows_emitted = 0;
  uint32_t seg_rounds = 0, row_w = 0;
  size_t C = 0;
  Segment seg[kSegments];
  uint64_t seg_next = 0;
  // host ring (sink host / csv): full rows.  Wire "full": pinned, the copy engine writes it.  Wire "delta": pageable,
  // allocated on first use and filled by materialise() from the compact segments of the pinned byte ring.
  double* ring = nullptr;
  size_t ring_rows = 0;
  uint64_t rows_enq = 0, rows_landed = 0, rows_consumed = 0;
  std::deque<PendingCopy> pending;
  std::vector<cudaEvent_t> ev_pool;
  // wire "delta"
  CompactBuf cbuf[kCompactBufs];
  unsigned char* cring = nullptr;      // pinned byte ring of compact segments
  size_t cring_bytes = 0, cring_head = 0;
  std::deque<HostSeg> hsegs;           // enqueued, not yet released
  uint64_t rows_expanded = 0;          // rows materialised into `ring` so far (>= rows_consumed)
  std::vector<double> prev_row;        // [S][d+3] state of the expansion cursor
  double* d_prev[2] = {nullptr, nullptr}; // the last round of full rows of the previous / of this launch (device, ping-pong)
  int prev_cur = 0;                    // d_prev[prev_cur] holds the previous launch's last round
  bool have_prev = false;
  bool prev_valid = true;              // prev_row follows the stream (false after slgpu_discard_segment, until a key segment)
  bool force_key = true;               // the next segment must not lean on earlier rows
  double est_changed_per_round = -1.0; // payload rows per round of recent segments (sizes the copies)
  uint64_t d2h_bytes = 0;              // bytes handed to the copy engine so far
  FILE* bin_file = nullptr;            // sink "bin"
  // csv sink
  std::vector<std::string> csv_buf;
  CsvPool csv;
};

namespace {

// ---------------------------------------------------------------- config
// readSettings<T> semantics (src/app/jsonsettings.hpp:22-35): a missing or mistyped key is fatal.
struct ConfigError {
  std::string msg;
};

const slgpu_json::Value& need(const slgpu_json::Value& j, const char* key, slgpu_json::Value::Kind kind) {
  const slgpu_json::Value* v = j.find(key);
  if (!v || v->kind != kind) throw ConfigError{std::string(key) + " not found in config file and no default value exists."};
  return *v;
}
double need_num(const slgpu_json::Value& j, const char* key) { return need(j, key, slgpu_json::Value::Number).num; }
uint64_t need_uint(const slgpu_json::Value& j, const char* key) {
  const double v = need_num(j, key);
  if (!(v >= 0.0) || v != std::floor(v) || v > 1.8e19) throw ConfigError{std::string(key) + " must be a non-negative integer."};
  return (uint64_t)v;
}
std::vector<double> need_vec(const slgpu_json::Value& j, const char* key) {
  const slgpu_json::Value& a = need(j, key, slgpu_json::Value::Array);
  std::vector<double> out;
  for (const auto& e : a.arr) {
    if (e.kind != slgpu_json::Value::Number) throw ConfigError{std::string(key) + " must be an array of numbers."};
    out.push_back(e.num);
  }
  return out;
}

Settings parse_settings(const char* text) {
  slgpu_json::Value j;
  try {
    j = slgpu_json::parse(text);
  } catch (const std::exception& ex) {
    throw ConfigError{ex.what()};
  }
  if (j.kind != slgpu_json::Value::Object) throw ConfigError{"config must be a JSON object"};
  Settings s;
  s.nStacks = (uint32_t)need_uint(j, "nStacks");
  s.nTemps = (uint32_t)need_uint(j, "nTemperatures");
  s.nSamplesTotal = need_uint(j, "nSamplesTotal");
  s.swapInterval = (uint32_t)need_uint(j, "swapInterval");
  s.loggingRateSec = need_num(j, "loggingRateSec");
  s.nJobTypes = (uint32_t)need_uint(j, "nJobTypes");
  s.outputPath = need(j, "outputPath", slgpu_json::Value::String).str;
  s.optimalAcceptRate = need_num(j, "optimalAcceptRate");
  s.optimalSwapRate = need_num(j, "optimalSwapRate");
  s.useInitial = need(j, "useInitial", slgpu_json::Value::Bool).b;
  if (s.useInitial) s.initial = need_vec(j, "initial");
  s.mn = need_vec(j, "min");
  s.mx = need_vec(j, "max");
  // ProposalBoundsFromJSON, src/infer/sampler.hpp:34-60
  if (s.mn.size() != s.mx.size())
    throw ConfigError{"Proposal bounds dimension mismatch: nMin=" + std::to_string(s.mn.size()) + ", nMax=" + std::to_string(s.mx.size())};
  s.nDims = (uint32_t)s.mn.size();
  if (s.useInitial && s.initial.size() != s.nDims) throw ConfigError{"initial has the wrong number of dimensions"};
  if (s.nDims == 0 || s.nDims > SLGPU_MAX_DIMS) throw ConfigError{"nDims must be in 1.." + std::to_string(SLGPU_MAX_DIMS)};
  if (s.nStacks == 0) throw ConfigError{"nStacks must be positive"};
  if (s.nTemps == 0 || s.nTemps > SLGPU_MAX_TEMPS) throw ConfigError{"nTemperatures must be in 1.." + std::to_string(SLGPU_MAX_TEMPS)};
  if (s.swapInterval == 0) throw ConfigError{"swapInterval must be positive"};
  if (s.nJobTypes == 0) throw ConfigError{"nJobTypes must be positive"};
  for (uint32_t i = 0; i < s.nDims; ++i)
    if (!(s.mx[i] > s.mn[i])) throw ConfigError{"max must exceed min in every dimension"};
  s.adaptInterval = s.swapInterval;
  if (const slgpu_json::Value* g = j.find("gpu")) {
    if (g->kind != slgpu_json::Value::Object) throw ConfigError{"gpu must be an object"};
    if (g->find("device")) s.device = (int)need_uint(*g, "device");
    if (g->find("seed")) s.seed = need_uint(*g, "seed");
    if (g->find("stackOffset")) s.stackOffset = (uint32_t)need_uint(*g, "stackOffset");
    if (g->find("adaptInterval")) s.adaptInterval = (uint32_t)need_uint(*g, "adaptInterval");
    if (g->find("maxRoundsPerLaunch")) s.maxRoundsPerLaunch = (uint32_t)need_uint(*g, "maxRoundsPerLaunch");
    if (g->find("hostRingRounds")) s.hostRingRounds = need_uint(*g, "hostRingRounds");
    if (g->find("overwrite")) s.overwrite = need(*g, "overwrite", slgpu_json::Value::Bool).b;
    if (g->find("schedule")) {
      const std::string& k = need(*g, "schedule", slgpu_json::Value::String).str;
      if (k == "batched") s.sequential = false;
      else if (k == "sequential") s.sequential = true;
      else throw ConfigError{"gpu.schedule must be \"batched\" or \"sequential\""};
    }
    if (g->find("windowRates")) s.windowRates = need(*g, "windowRates", slgpu_json::Value::Bool).b;
    if (g->find("csvThreads")) s.csvThreads = (uint32_t)need_uint(*g, "csvThreads");
    if (g->find("append")) s.append = need(*g, "append", slgpu_json::Value::Bool).b;
    if (g->find("shaper")) {
      const std::string& k = need(*g, "shaper", slgpu_json::Value::String).str;
      if (k == "reference") s.welford = false;
      else if (k == "welford") s.welford = true;
      else throw ConfigError{"gpu.shaper must be \"reference\" or \"welford\""};
    }
    if (g->find("recholEvery")) s.recholEvery = (uint32_t)need_uint(*g, "recholEvery");
    if (g->find("wire")) {
      const std::string& k = need(*g, "wire", slgpu_json::Value::String).str;
      if (k == "delta") s.wire = WIRE_DELTA;
      else if (k == "full") s.wire = WIRE_FULL;
      else throw ConfigError{"gpu.wire must be \"delta\" or \"full\""};
    }
    if (g->find("rng")) {
      const std::string& r = need(*g, "rng", slgpu_json::Value::String).str;
      if (r == "philox") s.rng = SLGPU_RNG_PHILOX;
      else if (r == "replay") s.rng = SLGPU_RNG_REPLAY;
      else throw ConfigError{"gpu.rng must be \"philox\" or \"replay\""};
    }
    if (g->find("sink")) {
      const std::string& k = need(*g, "sink", slgpu_json::Value::String).str;
      if (k == "none") s.sink = SINK_NONE;
      else if (k == "device") s.sink = SINK_DEVICE;
      else if (k == "host") s.sink = SINK_HOST;
      else if (k == "csv") s.sink = SINK_CSV;
      else if (k == "bin") s.sink = SINK_BIN;
      else throw ConfigError{"gpu.sink must be one of none, device, host, csv, bin"};
    }
  }
  if (s.adaptInterval == 0 || s.maxRoundsPerLaunch == 0) throw ConfigError{"gpu.adaptInterval and gpu.maxRoundsPerLaunch must be positive"};
  return s;
}

template <class T>
slgpu_status dev_alloc(slgpu_engine* e, T** out, size_t n) {
  void* p = nullptr;
  CU_TRY(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
  e->allocs.push_back(p);
  *out = (T*)p;
  return SLGPU_OK;
}

slgpu_status upload(double* dst, const std::vector<double>& v) {
  CU_TRY(cudaMemcpy(dst, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice));
  return SLGPU_OK;
}

slgpu_status check_ready(slgpu_engine* e, bool need_init) {
  if (!e) return fail(SLGPU_E_ARG, "engine is NULL");
  if (e->failed) return fail(SLGPU_E_STATE, "engine is in a failed state: " + g_err);
  if (need_init && !e->initialised) return fail(SLGPU_E_STATE, "chains are not initialised (call slgpu_init_chains)");
  CU_TRY(cudaSetDevice(e->cfg.device));
  return SLGPU_OK;
}

cudaEvent_t get_event(slgpu_engine* e) {
  if (!e->ev_pool.empty()) {
    cudaEvent_t ev = e->ev_pool.back();
    e->ev_pool.pop_back();
    return ev;
  }
  cudaEvent_t ev = nullptr;
  cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
  return ev;
}

inline bool delta_wire(const slgpu_engine* e) {
  return e->cfg.wire == WIRE_DELTA && (e->cfg.sink == SINK_HOST || e->cfg.sink == SINK_CSV || e->cfg.sink == SINK_BIN);
}

// wire "delta": has the oldest not-yet-landed segment landed?  Checks the payload count against what the first copy
// covered and fetches the rest if the estimate was short (rare: the accept rates move slowly).
slgpu_status land_segment(slgpu_engine* e, HostSeg& hs, bool block, bool* landed) {
  *landed = false;
  if (block) CU_TRY(cudaEventSynchronize(hs.ev));
  else if (cudaEventQuery(hs.ev) != cudaSuccess) return SLGPU_OK;
  const uint32_t count = ((const uint32_t*)(e->cring + hs.off))[2];
  if (count > hs.cap_copied && !hs.remainder) {
    const SegLayout L = seg_layout(hs.n_rounds, e->cfg.nStacks);
    const size_t from = L.bytes(hs.cap_copied, e->cfg.nDims), to = L.bytes(count, e->cfg.nDims);
    CU_TRY(cudaMemcpyAsync(e->cring + hs.off + from, hs.dev + from, to - from, cudaMemcpyDeviceToHost, e->drain));
    CU_TRY(cudaEventRecord(hs.ev, e->drain));
    e->d2h_bytes += to - from;
    hs.remainder = true;
    hs.cap_copied = count;
    if (!block) return SLGPU_OK;
    CU_TRY(cudaEventSynchronize(hs.ev));
  }
  hs.landed = true;
  *landed = true;
  const bool key = ((const uint32_t*)(e->cring + hs.off))[3] != 0;
  const uint32_t plain_rounds = hs.n_rounds - (key ? 1u : 0u);
  if (plain_rounds) e->est_changed_per_round = (double)(count - (key ? e->cfg.nStacks : 0u)) / plain_rounds;
  return SLGPU_OK;
}

// rows whose D2H copy has completed
void poll_landed(slgpu_engine* e, bool block) {
  if (delta_wire(e)) {
    for (HostSeg& hs : e->hsegs) {
      if (!hs.landed) {
        bool ok = false;
        if (land_segment(e, hs, block, &ok) != SLGPU_OK || !ok) break;
      }
      e->rows_landed = hs.row_begin + (uint64_t)hs.n_rounds * e->cfg.nStacks;
    }
    return;
  }
  while (!e->pending.empty()) {
    PendingCopy& pc = e->pending.front();
    if (block) cudaEventSynchronize(pc.ev);
    else if (cudaEventQuery(pc.ev) != cudaSuccess) break;
    e->rows_landed = pc.rows_end;
    e->ev_pool.push_back(pc.ev);
    e->pending.pop_front();
  }
  if (e->cfg.overwrite && e->rows_landed > e->ring_rows && e->rows_consumed + e->ring_rows < e->rows_landed)
    e->rows_consumed = e->rows_landed - e->ring_rows;
}

// one compact segment applied to the expansion state; rows != NULL also writes the full (d + 5)-double rows
void apply_segment(slgpu_engine* e, const HostSeg& hs, double* ring) {
  const uint32_t S = e->cfg.nStacks, w = e->row_w, wp = w - 2;
  const SegLayout L = seg_layout(hs.n_rounds, S);
  const unsigned char* seg = e->cring + hs.off;
  const uint32_t* base = (const uint32_t*)(seg + L.off_base);
  const unsigned char* flags = seg + L.off_flags;
  const double* pay = (const double*)(seg + L.off_payload);
  if (e->prev_row.size() != (size_t)S * wp) e->prev_row.assign((size_t)S * wp, 0.0);
  for (uint32_t r = 0; r < hs.n_rounds; ++r) {
    for (uint32_t c = 0; c < L.n_chunks; ++c) {
      const double* src = pay + (size_t)base[(size_t)r * L.n_chunks + c] * wp;
      const uint32_t s1 = std::min(S, (c + 1) * (uint32_t)slgpu_rows::kChunk);
      for (uint32_t st = c * slgpu_rows::kChunk; st < s1; ++st) {
        const unsigned f = flags[(size_t)r * S + st];
        double* pv = e->prev_row.data() + (size_t)st * wp;
        if (f & 8u) {
          std::memcpy(pv, src, wp * sizeof(double));
          src += wp;
        }
        if (ring) {
          double* out = ring + (size_t)((hs.row_begin + (uint64_t)r * S + st) % e->ring_rows) * w;
          std::memcpy(out, pv, wp * sizeof(double));
          out[wp] = (double)(f & 1u);
          out[wp + 1] = (double)((f >> 1) & 3u);
        }
      }
    }
  }
}

void release_front_segment(slgpu_engine* e) {
  HostSeg& hs = e->hsegs.front();
  e->ev_pool.push_back(hs.ev);
  e->hsegs.pop_front();
}

// wire "delta": rebuild the full rows of every landed segment into the (pageable) row ring
slgpu_status materialise(slgpu_engine* e) {
  if (!delta_wire(e)) return SLGPU_OK;
  while (!e->hsegs.empty() && e->hsegs.front().landed) {
    if (!e->ring) {
      e->ring = (double*)std::malloc(e->ring_rows * e->row_w * sizeof(double));
      if (!e->ring) return fail(SLGPU_E_STATE, "out of host memory for the row ring");
    }
    const HostSeg& hs = e->hsegs.front();
    const uint64_t end = hs.row_begin + (uint64_t)hs.n_rounds * e->cfg.nStacks;
    if (!e->prev_valid) {
      if (((const uint32_t*)(e->cring + hs.off))[3] == 0)
        return fail(SLGPU_E_STATE, "rows cannot be rebuilt from here: segments were discarded (slgpu_discard_segment) and this one is not a key segment");
      e->prev_valid = true;
    }
    if (e->cfg.overwrite && end > e->rows_consumed + e->ring_rows) e->rows_consumed = end - e->ring_rows;
    apply_segment(e, hs, e->ring);
    e->rows_expanded = end;
    release_front_segment(e);
  }
  return SLGPU_OK;
}

// One value as operator<<(ostream&, double) prints it with default flags (precision 6, general): printf's "%g".
// std::to_chars(general, 6) is specified to produce exactly that text and is several times faster than snprintf.
inline char* put_g(char* p, double v) {
  if (!std::isfinite(v)) return p + std::snprintf(p, 16, "%g", v);
  return std::to_chars(p, p + 32, v, std::chars_format::general, 6).ptr;
}

// operator<<(ostream&, State), src/db/db.cpp:18-25; buf must hold 32 * (d + 5) bytes.  Returns the length.
inline size_t format_row_into(const double* row, uint32_t d, char* buf) {
  char* p = buf;
  for (uint32_t i = 0; i < d + 3; ++i) {
    p = put_g(p, row[i]);
    *p++ = ',';
  }
  p = std::to_chars(p, p + 16, (int)row[d + 3]).ptr;
  *p++ = ',';
  p = std::to_chars(p, p + 16, (int)row[d + 4]).ptr;
  return (size_t)(p - buf);
}

size_t format_row(const double* row, uint32_t d, char* out, size_t cap) {
  std::vector<char> buf((size_t)32 * (d + SLGPU_ROW_EXTRA));
  const size_t len = format_row_into(row, d, buf.data());
  if (out && cap) {
    const size_t n = std::min(cap - 1, len);
    std::memcpy(out, buf.data(), n);
    out[n] = 0;
  }
  return len;
}

std::string csv_path(const slgpu_engine* e, uint32_t local_stack) {
  return e->cfg.outputPath + "/" + std::to_string(e->cfg.stackOffset + local_stack) + ".csv";
}

// worker w of n: format rows [r0, r1) of its stacks, then (flush) append its buffers to <outputPath>/<stack>.csv
void csv_work(slgpu_engine* e, uint32_t w, uint32_t n, uint64_t r0, uint64_t r1, bool flush) {
  const uint32_t S = e->cfg.nStacks, d = e->cfg.nDims;
  char line[32 * (SLGPU_MAX_DIMS + SLGPU_ROW_EXTRA) + 1];
  size_t held = e->csv.buffered[w];
  for (uint32_t s = w; s < S; s += n) {
    std::string& b = e->csv_buf[s];
    uint64_t r = r0 + (s + S - (uint32_t)(r0 % S)) % S;  // first row >= r0 of stack s
    for (; r < r1; r += S) {
      const double* row = e->ring + (size_t)(r % e->ring_rows) * e->row_w;
      size_t len = format_row_into(row, d, line);
      line[len++] = '\n';
      b.append(line, len);
      held += len;
    }
    if (flush && !b.empty()) {
      FILE* f = std::fopen(csv_path(e, s).c_str(), "ab");
      const bool ok = f && std::fwrite(b.data(), 1, b.size(), f) == b.size();
      if (f) std::fclose(f);
      if (!ok) {
        std::lock_guard<std::mutex> lk(e->csv.m);
        e->csv.io_error = true;
      }
      b.clear();
    }
  }
  e->csv.buffered[w] = flush ? 0 : held;
}

void csv_worker_main(slgpu_engine* e, uint32_t w, uint32_t n) {
  CsvPool& cp = e->csv;
  uint64_t seen = 0;
  for (;;) {
    uint64_t r0, r1;
    bool flush;
    {
      std::unique_lock<std::mutex> lk(cp.m);
      cp.go.wait(lk, [&] { return cp.stop || cp.generation != seen; });
      if (cp.stop) return;
      seen = cp.generation;
      r0 = cp.r0; r1 = cp.r1; flush = cp.flush;
    }
    csv_work(e, w, n, r0, r1, flush);
    {
      std::lock_guard<std::mutex> lk(cp.m);
      if (--cp.running == 0) cp.done.notify_one();
    }
  }
}

void csv_pool_start(slgpu_engine* e) {
  uint32_t n = e->cfg.csvThreads ? e->cfg.csvThreads : std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
  n = std::max(1u, std::min(n, e->cfg.nStacks));
  e->csv.buffered.assign(n, 0);
  if (n == 1) return;  // the caller's thread does the work
  for (uint32_t w = 0; w < n; ++w) e->csv.threads.emplace_back(csv_worker_main, e, w, n);
}

void csv_pool_stop(slgpu_engine* e) {
  {
    std::lock_guard<std::mutex> lk(e->csv.m);
    e->csv.stop = true;
  }
  e->csv.go.notify_all();
  for (auto& t : e->csv.threads) t.join();
  e->csv.threads.clear();
}

// format every landed row into the per-stack buffers (one <outputPath>/<stack>.csv per stack, db.cpp:27-35);
// buffers go to the files when they hold 64 MiB together, and at the end (finalise)
slgpu_status csv_pump(slgpu_engine* e, bool block, bool finalise) {
  NvtxRange nvtx_("slgpu:csv_pump");
  poll_landed(e, block);
  {
    slgpu_status st = materialise(e);
    if (st != SLGPU_OK) return st;
  }
  CsvPool& cp = e->csv;
  size_t held = 0;
  for (size_t b : cp.buffered) held += b;
  const uint64_t r0 = e->rows_consumed, r1 = delta_wire(e) ? e->rows_expanded : e->rows_landed;
  const bool flush = finalise || held + (r1 - r0) * 12 * e->row_w > (64u << 20);
  if (r0 == r1 && !(flush && held)) return SLGPU_OK;
  if (cp.threads.empty()) {
    csv_work(e, 0, 1, r0, r1, flush);
  } else {
    std::unique_lock<std::mutex> lk(cp.m);
    cp.r0 = r0; cp.r1 = r1; cp.flush = flush;
    cp.running = (uint32_t)cp.threads.size();
    cp.generation++;
    cp.go.notify_all();
    cp.done.wait(lk, [&] { return cp.running == 0; });
  }
  e->rows_consumed = r1;
  if (cp.io_error) return fail(SLGPU_E_IO, "cannot append to the csv files under " + e->cfg.outputPath);
  return SLGPU_OK;
}

// sink "bin": every landed compact segment goes to <outputPath>/chains-<stackOffset>.slgpu as it is -- a 32-byte
// record header {magic "SLGS", n_rounds, row_begin (u64), bytes (u64), reserved} followed by the segment (header, base,
// flags, payload[count]).  stateline_b200/binlog.py rebuilds the rows.  This is the sink that keeps up with the GPU:
// ~50 bytes per cold row instead of ~150 characters of text.
slgpu_status bin_pump(slgpu_engine* e, bool block, bool finalise) {
  NvtxRange nvtx_("slgpu:bin_pump");
  poll_landed(e, block);
  while (!e->hsegs.empty() && e->hsegs.front().landed) {
    const HostSeg& hs = e->hsegs.front();
    const unsigned char* seg = e->cring + hs.off;
    const uint32_t count = ((const uint32_t*)seg)[2];
    const uint64_t bytes = seg_layout(hs.n_rounds, e->cfg.nStacks).bytes(count, e->cfg.nDims);
    struct { char magic[4]; uint32_t n_rounds; uint64_t row_begin, bytes, reserved; } rec = {{'S', 'L', 'G', 'S'}, hs.n_rounds, hs.row_begin, bytes, 0};
    if (std::fwrite(&rec, sizeof rec, 1, e->bin_file) != 1 || std::fwrite(seg, 1, bytes, e->bin_file) != bytes)
      return fail(SLGPU_E_IO, "short write to the binary chain file under " + e->cfg.outputPath);
    const uint64_t end = hs.row_begin + (uint64_t)hs.n_rounds * e->cfg.nStacks;
    e->rows_consumed = e->rows_expanded = end;
    release_front_segment(e);
  }
  if (finalise && std::fflush(e->bin_file) != 0) return fail(SLGPU_E_IO, "cannot flush the binary chain file under " + e->cfg.outputPath);
  return SLGPU_OK;
}

// queue the D2H copy of a finished segment into the host ring
slgpu_status enqueue_copy(slgpu_engine* e, Segment& sg, uint64_t n_rows) {
  CU_TRY(cudaStreamWaitEvent(e->drain, sg.written, 0));
  uint64_t done = 0;
  while (done < n_rows) {
    const size_t pos = (size_t)((e->rows_enq + done) % e->ring_rows);
    const uint64_t n = std::min<uint64_t>(n_rows - done, e->ring_rows - pos);
    CU_TRY(cudaMemcpyAsync(e->ring + pos * e->row_w, sg.dev + (size_t)done * e->row_w, (size_t)n * e->row_w * sizeof(double),
                           cudaMemcpyDeviceToHost, e->drain));
    e->d2h_bytes += (size_t)n * e->row_w * sizeof(double);
    done += n;
  }
  CU_TRY(cudaEventRecord(sg.copied, e->drain));
  sg.copy_pending = true;
  cudaEvent_t ev = get_event(e);
  CU_TRY(cudaEventRecord(ev, e->drain));
  e->rows_enq += n_rows;
  e->pending.push_back({ev, e->rows_enq});
  return SLGPU_OK;
}

bool ring_has_room(slgpu_engine* e, uint64_t n_rows) {
  if (e->cfg.overwrite) return true;
  return e->rows_enq + n_rows - e->rows_consumed <= e->ring_rows;
}

// wire "delta": `bytes` contiguous bytes of the pinned byte ring (segments are released oldest first)
bool cring_reserve(slgpu_engine* e, size_t bytes, size_t* off) {
  bytes = (bytes + 255) / 256 * 256;
  if (bytes > e->cring_bytes) return false;
  if (e->hsegs.empty()) {
    e->cring_head = 0;
  }
  const size_t tail = e->hsegs.empty() ? 0 : e->hsegs.front().off;
  const bool wrapped = !e->hsegs.empty() && e->cring_head <= tail;  // head has wrapped behind the oldest segment (or the ring is full)
  if (!wrapped) {
    if (e->cring_head + bytes <= e->cring_bytes) {
      *off = e->cring_head;
    } else if (!e->hsegs.empty() && bytes <= tail) {
      *off = 0;
    } else if (e->hsegs.empty()) {
      *off = 0;
    } else {
      return false;
    }
  } else {
    if (e->cring_head + bytes > tail) return false;
    *off = e->cring_head;
  }
  e->cring_head = *off + bytes;
  return true;
}

// wire "delta": compact the launch's rows on the compute stream and queue ONE copy of the compact segment.
// The copy covers the fixed part and as many payload rows as recent segments needed (+25 %); land_segment fetches the
// rest in the rare case that was short.  A compact device buffer is reused only after the host has seen its segment
// land, so that such a late fetch always finds its data.
slgpu_status enqueue_compact(slgpu_engine* e, const double* rows_dev, uint32_t n_rounds) {
  NvtxRange nvtx_("slgpu:compact_rows");
  const uint32_t S = e->cfg.nStacks, d = e->cfg.nDims;
  CompactBuf& cb = e->cbuf[e->seg_next % kCompactBufs];
  for (HostSeg& hs : e->hsegs) {  // the segment that used this buffer last, if it is still in flight
    if (hs.dev == cb.dev && !hs.landed) {
      for (HostSeg& h2 : e->hsegs) {  // segments land in order
        bool ok = false;
        if (!h2.landed) {
          slgpu_status st = land_segment(e, h2, true, &ok);
          if (st != SLGPU_OK) return st;
        }
        if (&h2 == &hs) break;
      }
      break;
    }
  }
  const SegLayout L = seg_layout(n_rounds, S);
  const bool key = e->force_key || e->cfg.overwrite || !e->have_prev;
  CU_TRY(cudaMemsetAsync(cb.dev, 0, 16, e->compute));
  slgpu_rows::k_compact_rows<<<dim3(L.n_chunks, n_rounds), slgpu_rows::kChunk, 0, e->compute>>>(
      rows_dev, e->d_prev[e->prev_cur], e->d_prev[e->prev_cur ^ 1], S, e->row_w, n_rounds, key ? 1 : 0, cb.dev, L.n_chunks,
      (uint32_t)L.off_flags, (uint32_t)L.off_payload);
  e->prev_cur ^= 1;
  e->have_prev = true;
  CU_TRY(cudaGetLastError());
  e->launches++;
  CU_TRY(cudaEventRecord(cb.written, e->compute));
  const uint32_t capacity = n_rounds * S;
  uint32_t cap = capacity;
  if (e->est_changed_per_round >= 0.0) {
    const double want = e->est_changed_per_round * (n_rounds - (key ? 1 : 0)) * 1.25 + 128.0 + (key ? S : 0);
    if (want < (double)capacity) cap = (uint32_t)want;
  }
  size_t off = 0;
  while (!cring_reserve(e, L.bytes(capacity, d), &off)) {
    if (!e->cfg.overwrite || e->hsegs.empty()) return fail(SLGPU_E_FULL, "host ring cannot take the segment: drain first or raise gpu.hostRingRounds");
    HostSeg& old = e->hsegs.front();  // overwrite: the oldest unconsumed segment makes room (every segment is a key segment)
    if (!old.landed) CU_TRY(cudaEventSynchronize(old.ev));
    const uint64_t end = old.row_begin + (uint64_t)old.n_rounds * S;
    e->rows_landed = std::max(e->rows_landed, end);
    e->rows_consumed = std::max(e->rows_consumed, end);
    e->rows_expanded = std::max(e->rows_expanded, end);
    release_front_segment(e);
  }
  CU_TRY(cudaStreamWaitEvent(e->drain, cb.written, 0));
  CU_TRY(cudaMemcpyAsync(e->cring + off, cb.dev, L.bytes(cap, d), cudaMemcpyDeviceToHost, e->drain));
  e->d2h_bytes += L.bytes(cap, d);
  HostSeg hs;
  hs.off = off; hs.reserved = L.bytes(capacity, d); hs.n_rounds = n_rounds; hs.cap_copied = cap;
  hs.row_begin = e->rows_enq; hs.dev = cb.dev; hs.ev = get_event(e);
  CU_TRY(cudaEventRecord(hs.ev, e->drain));
  e->hsegs.push_back(hs);
  e->rows_enq += (uint64_t)n_rounds * S;
  e->force_key = false;
  return SLGPU_OK;
}

// one launch of the plugin's step kernel over [rounds_done, rounds_done + n)
slgpu_status launch_rounds(slgpu_engine* e, uint32_t n) {
  NvtxRange nvtx_("slgpu:launch_rounds");
  const bool emit = e->cfg.sink != SINK_NONE;
  const bool to_host = e->cfg.sink == SINK_HOST || e->cfg.sink == SINK_CSV || e->cfg.sink == SINK_BIN;
  const uint64_t n_rows = (uint64_t)n * e->cfg.nStacks;
  Segment& sg = e->seg[e->seg_next % kSegments];
  if (emit) {
    if (to_host) {
      if (e->cfg.sink == SINK_CSV) {
        slgpu_status st = csv_pump(e, false, false);
        if (st != SLGPU_OK) return st;
        while (!ring_has_room(e, n_rows)) {
          st = csv_pump(e, true, false);
          if (st != SLGPU_OK) return st;
        }
      }
      if (e->cfg.sink == SINK_BIN) {
        slgpu_status st = bin_pump(e, false, false);
        if (st != SLGPU_OK) return st;
        while (!ring_has_room(e, n_rows)) {
          st = bin_pump(e, true, false);
          if (st != SLGPU_OK) return st;
        }
      }
      if (sg.copy_pending) CU_TRY(cudaStreamWaitEvent(e->compute, sg.copied, 0));
    }
    e->P.rows_out = sg.dev;
  } else {
    e->P.rows_out = nullptr;
  }
  e->P.round_begin = e->rounds_done;
  e->P.n_rounds = n;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (e->timing) {  // pooled: creating events is host work inside the timed region
    for (cudaEvent_t* ev : {&ev0, &ev1}) {
      if (!e->ev_time_pool.empty()) {
        *ev = e->ev_time_pool.back();
        e->ev_time_pool.pop_back();
      } else {
        cudaEventCreate(ev);
      }
    }
    cudaEventRecord(ev0, e->compute);
  }
  const int rc = e->plug->launch_step(&e->P, (void*)e->compute);
  if (e->timing) {
    cudaEventRecord(ev1, e->compute);
    e->step_ev.push_back(ev0);
    e->step_ev.push_back(ev1);
  }
  if (rc != 0) {
    e->failed = true;
    return fail(SLGPU_E_CUDA, std::string("launch_step: ") + cudaGetErrorString((cudaError_t)rc));
  }
  e->launches++;
  if (e->d_round_flags) {
    const uint32_t C = (uint32_t)e->C;
    slgpu_tier::k_window_push<<<(C + 255) / 256, 256, 0, e->compute>>>(e->d_round_flags, n, C, e->cfg.nTemps, e->d_win_sig, e->d_win_beta,
                                                                       e->d_win_cnt, e->d_win_sum);
    CU_TRY(cudaGetLastError());
    e->launches++;
  }
  if (emit) {
    e->rows_emitted += n_rows;
    if (to_host) {
      slgpu_status st;
      if (delta_wire(e)) {
        st = enqueue_compact(e, sg.dev, n);
      } else {
        CU_TRY(cudaEventRecord(sg.written, e->compute));
        st = enqueue_copy(e, sg, n_rows);
      }
      if (st != SLGPU_OK) return st;
    }
    e->seg_next++;
  }
  e->rounds_done += n;
  return SLGPU_OK;
}

// sequential schedule: Sampler::step()'s model updates of the round just launched, in the reference's order
slgpu_status sequential_scan(slgpu_engine* e) {
  slgpu_tier::k_tier_scan_sequential<<<1, 32, 0, e->compute>>>(e->P.p_sig, e->P.p_beta, e->cfg.nStacks, e->cfg.nTemps, e->d_models, e->d_sig_w,
                                                               e->d_ladder, e->P.sigma, e->P.beta, e->d_beta_values,
                                                               e->cfg.optimalAcceptRate, e->cfg.optimalSwapRate);
  CU_TRY(cudaGetLastError());
  e->launches += 1;
  return SLGPU_OK;
}

slgpu_status adapt_point(slgpu_engine* e) {
  NvtxRange nvtx_("slgpu:adapt_point");
  const uint32_t T = e->cfg.nTemps;
  slgpu_tier::k_tier_adapt<<<dim3(T, 2, 9), 256, 0, e->compute>>>(e->P.p_sig, e->P.p_beta, e->cfg.nStacks, T, e->d_sums, e->d_ticket, e->d_models,
                                                                  e->d_sig_w, e->cfg.optimalSwapRate, e->d_ladder);
  CU_TRY(cudaGetLastError());
  e->launches += 1;
  // the periodic re-Cholesky: at every recholEvery-th adapt point (a schedule in rounds, not in launches), for the chains that logged a jump
  if (e->cfg.welford && ((e->rounds_done + 1) / e->cfg.adaptInterval) % e->cfg.recholEvery == 0) {
    const int rc2 = e->plug->launch_rechol ? e->plug->launch_rechol(&e->P, (void*)e->compute) : (int)cudaErrorNotSupported;
    if (rc2 != 0) {
      e->failed = true;
      return fail(rc2 == (int)cudaErrorNotSupported ? SLGPU_E_PLUGIN : SLGPU_E_CUDA,
                  std::string("launch_rechol: ") + (rc2 == (int)cudaErrorNotSupported ? "the plugin has no Welford kernels for this dimension (list it in SLGPU_DEFINE_PLUGIN)" : cudaGetErrorString((cudaError_t)rc2)));
    }
    e->launches++;
  }
  return SLGPU_OK;
}

template <class T>
slgpu_status fetch(std::vector<T>& host, const T* dev, size_t n) {
  host.resize(n);
  CU_TRY(cudaMemcpy(host.data(), dev, n * sizeof(T), cudaMemcpyDeviceToHost));
  return SLGPU_OK;
}

slgpu_status sync_all(slgpu_engine* e) {
  cudaError_t a = cudaStreamSynchronize(e->compute);
  cudaError_t b = cudaStreamSynchronize(e->drain);
  if (a != cudaSuccess || b != cudaSuccess) {
    e->failed = true;
    return fail(SLGPU_E_CUDA, std::string("stream synchronize: ") + cudaGetErrorString(a != cudaSuccess ? a : b));
  }
  poll_landed(e, true);
  return SLGPU_OK;
}

// 16 independent DFMA chains per thread: the FP64 FMA pipe roof that roofline fractions are quoted against
__global__ void k_dfma_peak(double* sink, int iters, double m) {
  double a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], m, 1e-12);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 123.456) *sink = s;
}

}  // namespace

extern "C" {

const char* slgpu_last_error(void) { return g_err.c_str(); }
const char* slgpu_version(void) { return "slgpu 0.1 sm_100a"; }

static slgpu_status create_from_settings(const Settings& cfg, const char* plugin_path, const char* plugin_entry, const void* payload,
                                         size_t payload_bytes, slgpu_engine** out);

slgpu_status slgpu_create(const char* config_json, const char* plugin_path, const char* plugin_entry, const void* payload,
                          size_t payload_bytes, slgpu_engine** out) {
  if (!config_json || !plugin_path || !out) return fail(SLGPU_E_ARG, "config_json, plugin_path and out must not be NULL");
  *out = nullptr;
  Settings cfg;
  try {
    cfg = parse_settings(config_json);
  } catch (const ConfigError& ce) {
    return fail(SLGPU_E_CONFIG, ce.msg);
  }
  return create_from_settings(cfg, plugin_path, plugin_entry, payload, payload_bytes, out);
}

static slgpu_status create_from_settings(const Settings& cfg, const char* plugin_path, const char* plugin_entry, const void* payload,
                                         size_t payload_bytes, slgpu_engine** out) {
  // plugin
  void* dl = dlopen(plugin_path, RTLD_NOW | RTLD_LOCAL);
  if (!dl) return fail(SLGPU_E_PLUGIN, std::string("dlopen: ") + dlerror());
  const char* sym = plugin_entry && plugin_entry[0] ? plugin_entry : "slgpu_plugin_entry";
  slgpu_plugin_entry_fn entry = (slgpu_plugin_entry_fn)dlsym(dl, sym);
  if (!entry) {
    dlclose(dl);
    return fail(SLGPU_E_PLUGIN, std::string("plugin has no symbol ") + sym);
  }
  const slgpu_plugin_v1* plug = entry();
  if (!plug || plug->abi_version != SLGPU_PLUGIN_ABI || !plug->launch_init || !plug->launch_step) {
    dlclose(dl);
    return fail(SLGPU_E_PLUGIN, "plugin ABI mismatch");
  }
  if ((plug->n_job_types && plug->n_job_types != cfg.nJobTypes) || (plug->n_dims && plug->n_dims != cfg.nDims) ||
      (plug->payload_bytes && plug->payload_bytes != payload_bytes)) {
    dlclose(dl);
    return fail(SLGPU_E_PLUGIN, std::string("plugin ") + plug->name + " does not match nJobTypes / dimensions / payload size of the config");
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    dlclose(dl);
    return fail(SLGPU_E_CUDA, "no CUDA device: the engine has no CPU fallback");
  }
  if (cfg.device >= ndev) {
    dlclose(dl);
    return fail(SLGPU_E_CONFIG, "gpu.device out of range");
  }
  slgpu_engine* e = new slgpu_engine();
  e->cfg = cfg;
  e->dl = dl;
  e->plug = plug;
  slgpu_status st = [&]() -> slgpu_status {
    CU_TRY(cudaSetDevice(cfg.device));
    CU_TRY(cudaStreamCreateWithFlags(&e->compute, cudaStreamNonBlocking));
    CU_TRY(cudaStreamCreateWithFlags(&e->drain, cudaStreamNonBlocking));
    CU_TRY(cudaEventCreate(&e->t0));
    CU_TRY(cudaEventCreate(&e->t1));
    const uint32_t S = cfg.nStacks, T = cfg.nTemps, d = cfg.nDims;
    const size_t C = (size_t)S * T;
    e->C = C;
    e->row_w = d + SLGPU_ROW_EXTRA;
    slgpu_step_params& P = e->P;
    P.n_stacks = S; P.n_temps = T; P.n_dims = d; P.n_job_types = cfg.nJobTypes;
    P.swap_interval = cfg.swapInterval;
    P.stack_offset = cfg.stackOffset;
    P.rng_mode = cfg.rng;
    P.seed = cfg.seed;
    P.opt_accept = cfg.optimalAcceptRate;
    P.sh_count0 = 1000.0;  // src/app/serverwrapper.cpp:50
    // ProposalShaper ctor, adaptive.cpp:161-164
    std::vector<double> range(d), n0(d);
    double n2 = 0.0;
    for (uint32_t i = 0; i < d; ++i) {
      range[i] = (cfg.mx[i] - cfg.mn[i]) / 4.0;
      n2 += range[i] * range[i];
    }
    P.prop_norm = std::sqrt(n2);
    for (uint32_t i = 0; i < d; ++i) n0[i] = range[i] / P.prop_norm;
    double *d_mn, *d_mx, *d_n0, *d_init = nullptr;
    slgpu_status s2;
#define TRY_ST(x)                           \
  do {                                      \
    if ((s2 = (x)) != SLGPU_OK) return s2;  \
  } while (0)
    TRY_ST(dev_alloc(e, &d_mn, d)); TRY_ST(upload(d_mn, cfg.mn));
    TRY_ST(dev_alloc(e, &d_mx, d)); TRY_ST(upload(d_mx, cfg.mx));
    TRY_ST(dev_alloc(e, &d_n0, d)); TRY_ST(upload(d_n0, n0));
    if (cfg.useInitial) { TRY_ST(dev_alloc(e, &d_init, d)); TRY_ST(upload(d_init, cfg.initial)); }
    P.mn = d_mn; P.mx = d_mx; P.n0_diag = d_n0; P.initial = d_init;
    TRY_ST(dev_alloc(e, &P.x, d * C));
    const size_t nLall = SLGPU_L_DOUBLES(S, T, d);
    TRY_ST(dev_alloc(e, &P.L, nLall));
    CU_TRY(cudaMemset(P.L, 0, nLall * sizeof(double)));  // the idle slots of a tile travel with it: keep them defined
    if (cfg.welford) {
      if (d > 32 || cfg.sequential || cfg.recholEvery == 0 || (uint64_t)cfg.adaptInterval * cfg.recholEvery > SLGPU_JUMP_LOG_SLOTS)
        return fail(SLGPU_E_CONFIG, "gpu.shaper \"welford\" needs nDims <= 32, the batched schedule and adaptInterval * recholEvery <= " + std::to_string(SLGPU_JUMP_LOG_SLOTS));
      TRY_ST(dev_alloc(e, &P.jump_log, (size_t)cfg.adaptInterval * cfg.recholEvery * d * C));
      TRY_ST(dev_alloc(e, &P.moment, nLall));
      CU_TRY(cudaMemset(P.moment, 0, nLall * sizeof(double)));
      TRY_ST(dev_alloc(e, &P.chol_dirty, C));
      CU_TRY(cudaMemset(P.chol_dirty, 0, C));
    }
    TRY_ST(dev_alloc(e, &P.energy, C)); TRY_ST(dev_alloc(e, &P.row_sigma, C)); TRY_ST(dev_alloc(e, &P.row_beta, C));
    TRY_ST(dev_alloc(e, &P.sigma, C)); TRY_ST(dev_alloc(e, &P.beta, C)); TRY_ST(dev_alloc(e, &P.nscale, C));
    TRY_ST(dev_alloc(e, &P.sh_count, C));
    TRY_ST(dev_alloc(e, &P.accepted, C)); TRY_ST(dev_alloc(e, &P.swap_type, C));
    TRY_ST(dev_alloc(e, &P.lg_accepts, C)); TRY_ST(dev_alloc(e, &P.lg_swaps, C)); TRY_ST(dev_alloc(e, &P.lg_swap_attempts, C));
    TRY_ST(dev_alloc(e, &P.lg_min_energy, C)); TRY_ST(dev_alloc(e, &P.lg_cur_energy, C));
    TRY_ST(dev_alloc(e, &P.p_sig, 9 * C)); TRY_ST(dev_alloc(e, &P.p_beta, 9 * C));
    TRY_ST(dev_alloc(e, &P.ep_m, (size_t)d * S)); TRY_ST(dev_alloc(e, &P.ep_s, (size_t)d * S));
    TRY_ST(dev_alloc(e, &e->d_models, (size_t)2 * T * slgpu_tier::kModelStride));
    TRY_ST(dev_alloc(e, &e->d_sig_w, (size_t)3 * T));
    TRY_ST(dev_alloc(e, &e->d_ladder, T));
    TRY_ST(dev_alloc(e, &e->d_sums, (size_t)2 * T * 9));
    TRY_ST(dev_alloc(e, &e->d_ticket, 1));
    CU_TRY(cudaMemset(e->d_ticket, 0, sizeof(unsigned int)));
    P.sig_w = e->d_sig_w;
    P.ladder = e->d_ladder;
    {
      const size_t f8 = sizeof(double), i4 = sizeof(int32_t);
      e->state = {{"x", P.x, d * C * f8}, {"L", P.L, nLall * f8}, {"energy", P.energy, C * f8}, {"row_sigma", P.row_sigma, C * f8},
                  {"row_beta", P.row_beta, C * f8}, {"sigma", P.sigma, C * f8}, {"beta", P.beta, C * f8}, {"nscale", P.nscale, C * f8},
                  {"sh_count", P.sh_count, C * f8}, {"accepted", P.accepted, C * i4}, {"swap_type", P.swap_type, C * i4},
                  {"lg_accepts", P.lg_accepts, C * i4}, {"lg_swaps", P.lg_swaps, C * i4}, {"lg_swap_attempts", P.lg_swap_attempts, C * i4},
                  {"lg_min_energy", P.lg_min_energy, C * f8}, {"lg_cur_energy", P.lg_cur_energy, C * f8}, {"p_sig", P.p_sig, 9 * C * f8}, {"p_beta", P.p_beta, 9 * C * f8},
                  {"ep_m", P.ep_m, (size_t)d * S * f8}, {"ep_s", P.ep_s, (size_t)d * S * f8},
                  {"models", e->d_models, (size_t)2 * T * slgpu_tier::kModelStride * f8}, {"sig_w", e->d_sig_w, (size_t)3 * T * f8},
                  {"ladder", e->d_ladder, T * f8}};
    }
    if (cfg.welford) {
      e->state.push_back({"moment", P.moment, nLall * sizeof(double)});
      e->state.push_back({"chol_dirty", P.chol_dirty, C});
      e->state.push_back({"jump_log", P.jump_log, (size_t)cfg.adaptInterval * cfg.recholEvery * d * C * sizeof(double)});
    }
    if (payload && payload_bytes) {
      TRY_ST(dev_alloc(e, (char**)&e->d_payload, payload_bytes));
      CU_TRY(cudaMemcpy(e->d_payload, payload, payload_bytes, cudaMemcpyHostToDevice));
    }
    P.payload = e->d_payload;
    // sink
    e->seg_rounds = std::min(cfg.adaptInterval, cfg.maxRoundsPerLaunch);
    if (cfg.sequential) {
      e->seg_rounds = 1;  // the tier models are replayed chain by chain after every round
      TRY_ST(dev_alloc(e, &e->d_beta_values, C));
      e->state.push_back({"beta_values", e->d_beta_values, C * sizeof(double)});
    }
    if (cfg.windowRates) {
      const size_t ringw = (size_t)slgpu_tier::kRingWords * C;
      TRY_ST(dev_alloc(e, &e->d_round_flags, (size_t)e->seg_rounds * C));
      TRY_ST(dev_alloc(e, &e->d_win_sig, ringw)); TRY_ST(dev_alloc(e, &e->d_win_beta, ringw));
      TRY_ST(dev_alloc(e, &e->d_win_cnt, 2 * C)); TRY_ST(dev_alloc(e, &e->d_win_sum, 2 * C));
      for (auto pr : {std::make_pair(e->d_win_sig, ringw), std::make_pair(e->d_win_beta, ringw), std::make_pair(e->d_win_cnt, 2 * C),
                      std::make_pair(e->d_win_sum, 2 * C)})
        CU_TRY(cudaMemset(pr.first, 0, pr.second * sizeof(uint32_t)));
      P.round_flags = e->d_round_flags;
      for (const StateBuf& b : {StateBuf{"win_sig", e->d_win_sig, ringw * 4}, StateBuf{"win_beta", e->d_win_beta, ringw * 4},
                                StateBuf{"win_cnt", e->d_win_cnt, 2 * C * 4}, StateBuf{"win_sum", e->d_win_sum, 2 * C * 4}})
        e->state.push_back(b);
    }
    if (cfg.sink != SINK_NONE) {
      const size_t seg_doubles = (size_t)e->seg_rounds * S * e->row_w;
      // full rows leave the device only on wire "full" (three buffers in flight); on the delta wire k_compact_rows consumes a
      // launch's rows on the compute stream before the next launch writes them: one buffer, re-used, stays in L2
      const int nseg = (cfg.sink == SINK_DEVICE || cfg.wire == WIRE_DELTA) ? 1 : kSegments;
      for (int i = 0; i < kSegments; ++i) {
        if (i < nseg) {
          TRY_ST(dev_alloc(e, &e->seg[i].dev, seg_doubles));
        } else {
          e->seg[i].dev = e->seg[0].dev;
        }
        CU_TRY(cudaEventCreateWithFlags(&e->seg[i].written, cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&e->seg[i].copied, cudaEventDisableTiming));
      }
    }
    if (cfg.sink == SINK_BIN && cfg.wire != WIRE_DELTA) return fail(SLGPU_E_CONFIG, "gpu.sink \"bin\" writes the delta stream: it needs gpu.wire \"delta\"");
    if (cfg.sink == SINK_HOST || cfg.sink == SINK_CSV || cfg.sink == SINK_BIN) {
      uint64_t rr = cfg.hostRingRounds ? cfg.hostRingRounds : (uint64_t)8 * e->seg_rounds;
      rr = std::max<uint64_t>(rr, (uint64_t)e->seg_rounds + 1);
      e->ring_rows = (size_t)rr * S;
      if (delta_wire(e)) {
        // compact segments: a pinned byte ring that can hold what the row ring could (a segment reserves its worst case)
        const size_t maxseg = seg_layout(e->seg_rounds, S).bytes(e->seg_rounds * S, d) + 256;
        e->cring_bytes = e->ring_rows * e->row_w * sizeof(double) + (size_t)(rr + 4) * 512 + 2 * maxseg;
        CU_TRY(cudaHostAlloc((void**)&e->cring, e->cring_bytes, cudaHostAllocDefault));
        TRY_ST(dev_alloc(e, &e->d_prev[0], (size_t)S * e->row_w));
        TRY_ST(dev_alloc(e, &e->d_prev[1], (size_t)S * e->row_w));
        for (int i = 0; i < kCompactBufs; ++i) {
          TRY_ST(dev_alloc(e, &e->cbuf[i].dev, maxseg));
          CU_TRY(cudaEventCreateWithFlags(&e->cbuf[i].written, cudaEventDisableTiming));
          CU_TRY(cudaEventCreateWithFlags(&e->cbuf[i].copied, cudaEventDisableTiming));
        }
      } else {
        CU_TRY(cudaHostAlloc((void**)&e->ring, e->ring_rows * e->row_w * sizeof(double), cudaHostAllocDefault));
      }
    }
    if (cfg.sink == SINK_BIN) {
      mkdir(cfg.outputPath.c_str(), 0777);
      const std::string path = cfg.outputPath + "/chains-" + std::to_string(cfg.stackOffset) + ".slgpu";
      e->bin_file = std::fopen(path.c_str(), cfg.append ? "ab" : "wb");
      if (!e->bin_file) return fail(SLGPU_E_IO, "cannot create " + path);
      if (!cfg.append || std::ftell(e->bin_file) == 0) {
        struct { char magic[8]; uint32_t version, n_stacks, n_temps, n_dims, stack_offset, chunk; } fh = {
            {'S', 'L', 'G', 'P', 'U', 'B', 'I', 'N'}, 1, S, T, d, cfg.stackOffset, (uint32_t)slgpu_rows::kChunk};
        if (std::fwrite(&fh, sizeof fh, 1, e->bin_file) != 1) return fail(SLGPU_E_IO, "cannot write " + path);
      }
    }
    if (cfg.sink == SINK_CSV) {
      // CSVChainArrayWriter ctor, db.cpp:27-35: one file per stack, truncated on open
      mkdir(cfg.outputPath.c_str(), 0777);
      e->csv_buf.assign(S, std::string());
      csv_pool_start(e);
      for (uint32_t s = 0; s < S; ++s) {
        std::ofstream f(csv_path(e, s), cfg.append ? std::ios::app : std::ios::trunc);
        if (!f.good()) return fail(SLGPU_E_IO, "cannot create " + csv_path(e, s));
      }
    }
    slgpu_tier::k_tier_init<<<1, 32, 0, e->compute>>>(e->d_models, e->d_sig_w, e->d_ladder, T, cfg.optimalSwapRate);
    CU_TRY(cudaGetLastError());
    e->launches++;
    CU_TRY(cudaStreamSynchronize(e->compute));
#undef TRY_ST
    return SLGPU_OK;
  }();
  if (st != SLGPU_OK) {
    const std::string keep = g_err;
    slgpu_destroy(e);
    g_err = keep;
    return st;
  }
  *out = e;
  return SLGPU_OK;
}

void slgpu_destroy(slgpu_engine* e) {
  if (!e) return;
  cudaSetDevice(e->cfg.device);
  if (e->compute) cudaStreamSynchronize(e->compute);
  if (e->drain) cudaStreamSynchronize(e->drain);
  if (e->cfg.sink == SINK_CSV) {
    if (e->initialised && !e->failed && (e->ring || e->cring)) csv_pump(e, true, true);
    csv_pool_stop(e);
  }
  for (void* p : e->allocs) cudaFree(p);
  for (auto& p : {e->d_rp_z, e->d_rp_umh, e->d_rp_uswap, e->d_rp_xinit})
    if (p) cudaFree(p);
  if (e->cfg.sink == SINK_BIN && e->bin_file) {
    if (e->initialised && !e->failed) bin_pump(e, true, true);
    std::fclose(e->bin_file);
  }
  if (e->ring) {
    if (delta_wire(e)) std::free(e->ring);
    else cudaFreeHost(e->ring);
  }
  if (e->cring) cudaFreeHost(e->cring);
  for (auto& cb : e->cbuf) {
    if (cb.written) cudaEventDestroy(cb.written);
    if (cb.copied) cudaEventDestroy(cb.copied);
  }
  for (auto& hs : e->hsegs) cudaEventDestroy(hs.ev);
  for (auto& sg : e->seg) {
    if (sg.written) cudaEventDestroy(sg.written);
    if (sg.copied) cudaEventDestroy(sg.copied);
  }
  for (auto& pc : e->pending) cudaEventDestroy(pc.ev);
  for (auto ev : e->ev_pool) cudaEventDestroy(ev);
  for (auto ev : e->ev_time_pool) cudaEventDestroy(ev);
  for (auto ev : e->step_ev) cudaEventDestroy(ev);
  if (e->t0) cudaEventDestroy(e->t0);
  if (e->t1) cudaEventDestroy(e->t1);
  if (e->compute) cudaStreamDestroy(e->compute);
  if (e->drain) cudaStreamDestroy(e->drain);
  if (e->dl) dlclose(e->dl);
  delete e;
}

slgpu_status slgpu_set_replay(slgpu_engine* e, const double* z, const double* u_mh, const double* u_swap, const double* x_init,
                              uint64_t n_rounds) {
  slgpu_status st = check_ready(e, false);
  if (st != SLGPU_OK) return st;
  if (e->cfg.rng != SLGPU_RNG_REPLAY) return fail(SLGPU_E_STATE, "gpu.rng is not \"replay\"");
  if (!z || !u_mh || !u_swap || (!x_init && !e->cfg.useInitial)) return fail(SLGPU_E_ARG, "replay buffers must not be NULL");
  st = sync_all(e);
  if (st != SLGPU_OK) return st;
  for (double** p : {&e->d_rp_z, &e->d_rp_umh, &e->d_rp_uswap, &e->d_rp_xinit}) {
    if (*p) cudaFree(*p);
    *p = nullptr;
  }
  const size_t C = e->C, d = e->cfg.nDims;
  auto up = [&](double** dst, const double* src, size_t n) -> slgpu_status {
    CU_TRY(cudaMalloc((void**)dst, std::max<size_t>(n, 1) * sizeof(double)));
    if (n) CU_TRY(cudaMemcpy(*dst, src, n * sizeof(double), cudaMemcpyHostToDevice));
    return SLGPU_OK;
  };
  if ((st = up(&e->d_rp_z, z, (size_t)n_rounds * C * d)) != SLGPU_OK) return st;
  if ((st = up(&e->d_rp_umh, u_mh, (size_t)n_rounds * C)) != SLGPU_OK) return st;
  if ((st = up(&e->d_rp_uswap, u_swap, (size_t)n_rounds * C)) != SLGPU_OK) return st;
  if (x_init && (st = up(&e->d_rp_xinit, x_init, C * d)) != SLGPU_OK) return st;
  e->P.rp_z = e->d_rp_z; e->P.rp_umh = e->d_rp_umh; e->P.rp_uswap = e->d_rp_uswap; e->P.rp_xinit = e->d_rp_xinit;
  e->P.replay_rounds = n_rounds;
  return SLGPU_OK;
}

slgpu_status slgpu_init_chains(slgpu_engine* e) {
  slgpu_status st = check_ready(e, false);
  if (st != SLGPU_OK) return st;
  if (e->initialised) return fail(SLGPU_E_STATE, "chains are already initialised");
  if (e->cfg.rng == SLGPU_RNG_REPLAY && !e->cfg.useInitial && !e->d_rp_xinit)
    return fail(SLGPU_E_STATE, "rng \"replay\" needs slgpu_set_replay before slgpu_init_chains");
  const bool emit = e->cfg.sink != SINK_NONE;
  const bool to_host = e->cfg.sink == SINK_HOST || e->cfg.sink == SINK_CSV || e->cfg.sink == SINK_BIN;
  Segment& sg = e->seg[e->seg_next % kSegments];
  e->P.rows_out = emit ? sg.dev : nullptr;
  e->P.round_begin = 0;
  e->P.n_rounds = 0;
  const int rc = e->plug->launch_init(&e->P, (void*)e->compute);
  if (rc != 0) {
    e->failed = true;
    return fail(SLGPU_E_CUDA, std::string("launch_init: ") + cudaGetErrorString((cudaError_t)rc));
  }
  e->launches++;
  if (e->d_beta_values)  // computeBetaStack of the init loop (serverwrapper.cpp:71-72): values_ == the initial beta
    CU_TRY(cudaMemcpyAsync(e->d_beta_values, e->P.beta, e->C * sizeof(double), cudaMemcpyDeviceToDevice, e->compute));
  if (emit) {
    e->rows_emitted += e->cfg.nStacks;
    if (to_host) {
      if (delta_wire(e)) {
        e->force_key = true;  // the initial rows are everybody's first row
        if ((st = enqueue_compact(e, sg.dev, 1)) != SLGPU_OK) return st;
      } else {
        CU_TRY(cudaEventRecord(sg.written, e->compute));
        if ((st = enqueue_copy(e, sg, e->cfg.nStacks)) != SLGPU_OK) return st;
      }
    }
    e->seg_next++;
  }
  e->initialised = true;
  e->rounds_done = 0;
  return SLGPU_OK;
}

slgpu_status slgpu_step_rounds(slgpu_engine* e, uint64_t n_rounds) {
  NvtxRange nvtx_("slgpu:step_rounds");
  slgpu_status st = check_ready(e, true);
  if (st != SLGPU_OK) return st;
  if (e->cfg.rng == SLGPU_RNG_REPLAY && e->rounds_done + n_rounds > e->P.replay_rounds)
    return fail(SLGPU_E_ARG, "not enough recorded rounds in the replay buffers");
  if (e->cfg.sink == SINK_HOST) {
    poll_landed(e, false);
    if (!ring_has_room(e, n_rounds * e->cfg.nStacks)) {
      poll_landed(e, true);
      if (!ring_has_room(e, n_rounds * e->cfg.nStacks))
        return fail(SLGPU_E_FULL, "host ring cannot take the rows of the requested rounds: drain first or raise gpu.hostRingRounds");
    }
  }
  const uint64_t A = e->cfg.adaptInterval;
  uint64_t left = n_rounds;
  while (left) {
    const uint64_t r = e->rounds_done;
    const uint64_t to_adapt = (A - (r + 2) % A) % A + 1;  // rounds up to and including the next adapt round
    const uint32_t n = (uint32_t)std::min<uint64_t>({left, to_adapt, (uint64_t)e->seg_rounds});
    if ((st = launch_rounds(e, n)) != SLGPU_OK) return st;
    if (e->cfg.sequential) {
      if ((st = sequential_scan(e)) != SLGPU_OK) return st;
    } else if ((e->rounds_done + 1) % A == 0 && (st = adapt_point(e)) != SLGPU_OK) {
      return st;
    }
    left -= n;
  }
  return SLGPU_OK;
}

slgpu_status slgpu_sync(slgpu_engine* e) {
  slgpu_status st = check_ready(e, false);
  if (st != SLGPU_OK) return st;
  if ((st = sync_all(e)) != SLGPU_OK) return st;
  if (e->cfg.sink == SINK_CSV && e->initialised) return csv_pump(e, true, true);  // every landed row is in its file
  if (e->cfg.sink == SINK_BIN && e->initialised) return bin_pump(e, true, true);
  return SLGPU_OK;
}

/* wait for the queued rounds without finalising the sink: the csv sink formats what has landed but leaves its buffers
 * to the 64 MiB / end-of-run flush (slgpu_sync re-opens and closes every <stack>.csv) */
slgpu_status slgpu_wait(slgpu_engine* e) {
  slgpu_status st = check_ready(e, false);
  if (st != SLGPU_OK) return st;
  if ((st = sync_all(e)) != SLGPU_OK) return st;
  if (e->cfg.sink == SINK_CSV && e->initialised) return csv_pump(e, true, false);
  if (e->cfg.sink == SINK_BIN && e->initialised) return bin_pump(e, true, false);
  return SLGPU_OK;
}

slgpu_status slgpu_run(slgpu_engine* e) {
  slgpu_status st = check_ready(e, false);
  if (st != SLGPU_OK) return st;
  if (!e->initialised && (st = slgpu_init_chains(e)) != SLGPU_OK) return st;
  // stop rule, src/app/serverwrapper.cpp:100-107: count cold-chain steps over all stacks
  const uint64_t S = e->cfg.nStacks;
  const uint64_t total = (e->cfg.nSamplesTotal + S - 1) / S;
  while (e->rounds_done < total) {
    uint64_t n = std::min<uint64_t>(total - e->rounds_done, (uint64_t)e->seg_rounds * 4);
    if (e->cfg.sink == SINK_HOST && !e->cfg.overwrite) {
      poll_landed(e, false);
      const uint64_t room = (e->ring_rows - (e->rows_enq - e->rows_consumed)) / S;
      if (room == 0) return fail(SLGPU_E_FULL, "host ring is full: slgpu_run with sink \"host\" needs gpu.overwrite or a ring that holds the run");
      n = std::min(n, room);
    }
    if ((st = slgpu_step_rounds(e, n)) != SLGPU_OK) return st;
  }
  if ((st = sync_all(e)) != SLGPU_OK) return st;
  if (e->cfg.sink == SINK_CSV) return csv_pump(e, true, true);  // Sampler::flush, sampler.cpp:216-235
  if (e->cfg.sink == SINK_BIN) return bin_pump(e, true, true);
  return SLGPU_OK;
}

uint64_t slgpu_rounds_done(const slgpu_engine* e) { return e ? e->rounds_done : 0; }
uint64_t slgpu_rows_emitted(const slgpu_engine* e) { return e ? e->rows_emitted : 0; }
uint64_t slgpu_launch_count(const slgpu_engine* e) { return e ? e->launches : 0; }

slgpu_status slgpu_timer_step_kernel(slgpu_engine* e, double* total_ms, uint64_t* n_launches) {
  if (!e || !total_ms || !n_launches) return fail(SLGPU_E_ARG, "NULL argument");
  *total_ms = e->step_kernel_ms;
  *n_launches = e->step_kernel_launches;
  return SLGPU_OK;
}
uint32_t slgpu_n_dims(const slgpu_engine* e) { return e ? e->cfg.nDims : 0; }
uint32_t slgpu_n_chains(const slgpu_engine* e) { return e ? (uint32_t)e->C : 0; }

slgpu_status slgpu_drain(slgpu_engine* e, double* rows, size_t cap_rows, size_t* n_rows) {
  NvtxRange nvtx_("slgpu:drain");
  slgpu_status st = check_ready(e, false);
  if (st != SLGPU_OK) return st;
  if (!n_rows || (!rows && cap_rows)) return fail(SLGPU_E_ARG, "rows / n_rows must not be NULL");
  if (e->cfg.sink != SINK_HOST) return fail(SLGPU_E_STATE, "slgpu_drain needs gpu.sink \"host\"");
  poll_landed(e, true);
  if ((st = materialise(e)) != SLGPU_OK) return st;
  size_t n = (size_t)std::min<uint64_t>(cap_rows, (delta_wire(e) ? e->rows_expanded : e->rows_landed) - e->rows_consumed);
  size_t done = 0;
  while (done < n) {
    const size_t pos = (size_t)((e->rows_consumed + done) % e->ring_rows);
    const size_t k = std::min(n - done, e->ring_rows - pos);
    std::memcpy(rows + done * e->row_w, e->ring + pos * e->row_w, k * e->row_w * sizeof(double));
    done += k;
  }
  e->rows_consumed += n;
  *n_rows = n;
  return SLGPU_OK;
}

slgpu_status slgpu_peek(slgpu_engine* e, const double** span0, size_t* n0, const double** span1, size_t* n1) {
  slgpu_status st = check_ready(e, false);
  if (st != SLGPU_OK) return st;
  if (!span0 || !n0 || !span1 || !n1) return fail(SLGPU_E_ARG, "span pointers must not be NULL");
  if (e->cfg.sink != SINK_HOST) return fail(SLGPU_E_STATE, "slgpu_peek needs gpu.sink \"host\"");
  poll_landed(e, false);
  if ((st = materialise(e)) != SLGPU_OK) return st;
  const size_t avail = (size_t)((delta_wire(e) ? e->rows_expanded : e->rows_landed) - e->rows_consumed);
  const size_t pos = (size_t)(e->rows_consumed % e->ring_rows);
  const size_t first = std::min(avail, e->ring_rows - pos);
  *span0 = e->ring + pos * e->row_w;
  *n0 = first;
  *span1 = e->ring;
  *n1 = avail - first;
  return SLGPU_OK;
}

slgpu_status slgpu_advance(slgpu_engine* e, size_t n_rows) {
  if (!e) return fail(SLGPU_E_ARG, "engine is NULL");
  if (n_rows > (delta_wire(e) ? e->rows_expanded : e->rows_landed) - e->rows_consumed) return fail(SLGPU_E_ARG, "cannot advance past the landed rows");
  e->rows_consumed += n_rows;
  return SLGPU_OK;
}

slgpu_status slgpu_next_segment(slgpu_engine* e, int block, slgpu_segment* out) {
  slgpu_status st = check_ready(e, false);
  if (st != SLGPU_OK) return st;
  if (!out) return fail(SLGPU_E_ARG, "out must not be NULL");
  if (!(delta_wire(e) && e->cfg.sink == SINK_HOST)) return fail(SLGPU_E_STATE, "slgpu_next_segment needs gpu.sink \"host\" with gpu.wire \"delta\"");
  if (e->rows_consumed != e->rows_expanded) return fail(SLGPU_E_STATE, "drain the materialised rows before taking segments");
  std::memset(out, 0, sizeof *out);
  if (e->hsegs.empty()) return SLGPU_OK;
  HostSeg& hs = e->hsegs.front();
  if (!hs.landed) {
    bool ok = false;
    if ((st = land_segment(e, hs, block != 0, &ok)) != SLGPU_OK) return st;
    if (!ok) return SLGPU_OK;
  }
  const SegLayout L = seg_layout(hs.n_rounds, e->cfg.nStacks);
  const unsigned char* seg = e->cring + hs.off;
  out->n_rounds = hs.n_rounds;
  out->n_stacks = e->cfg.nStacks;
  out->n_chunks = L.n_chunks;
  out->payload_doubles = e->row_w - 2;
  out->count = ((const uint32_t*)seg)[2];
  out->key = ((const uint32_t*)seg)[3];
  out->row_begin = hs.row_begin;
  out->base = (const uint32_t*)(seg + L.off_base);
  out->flags = seg + L.off_flags;
  out->payload = (const double*)(seg + L.off_payload);
  return SLGPU_OK;
}

slgpu_status slgpu_release_segment(slgpu_engine* e) {
  if (!e) return fail(SLGPU_E_ARG, "engine is NULL");
  if (e->hsegs.empty() || !e->hsegs.front().landed) return fail(SLGPU_E_STATE, "no landed segment to release");
  const HostSeg& hs = e->hsegs.front();
  if (!e->prev_valid && ((const uint32_t*)(e->cring + hs.off))[3] != 0) e->prev_valid = true;
  if (e->prev_valid) apply_segment(e, hs, nullptr);  // a later slgpu_drain starts from the right rows
  e->rows_consumed = e->rows_expanded = hs.row_begin + (uint64_t)hs.n_rounds * e->cfg.nStacks;
  e->rows_landed = std::max(e->rows_landed, e->rows_consumed);
  release_front_segment(e);
  return SLGPU_OK;
}

slgpu_status slgpu_discard_segment(slgpu_engine* e) {
  if (!e) return fail(SLGPU_E_ARG, "engine is NULL");
  if (e->hsegs.empty() || !e->hsegs.front().landed) return fail(SLGPU_E_STATE, "no landed segment to discard");
  const HostSeg& hs = e->hsegs.front();
  e->rows_consumed = e->rows_expanded = hs.row_begin + (uint64_t)hs.n_rounds * e->cfg.nStacks;
  e->rows_landed = std::max(e->rows_landed, e->rows_consumed);
  e->prev_valid = false;  // the expansion state no longer follows the stream: rows can be rebuilt again from a key segment
  release_front_segment(e);
  return SLGPU_OK;
}

slgpu_status slgpu_request_key_segment(slgpu_engine* e) {
  if (!e) return fail(SLGPU_E_ARG, "engine is NULL");
  e->force_key = true;
  return SLGPU_OK;
}

uint64_t slgpu_d2h_bytes(const slgpu_engine* e) { return e ? e->d2h_bytes : 0; }

static slgpu_status timer_start(slgpu_engine* e, bool per_launch) {
  slgpu_status st = check_ready(e, false);
  if (st != SLGPU_OK) return st;
  for (cudaEvent_t ev : e->step_ev) e->ev_time_pool.push_back(ev);
  e->step_ev.clear();
  e->timing = per_launch;
  CU_TRY(cudaEventRecord(e->t0, e->compute));
  return SLGPU_OK;
}
slgpu_status slgpu_timer_start(slgpu_engine* e) { return timer_start(e, true); }
/* the same without an event pair around every step-kernel launch: an event between two kernels also keeps the second
 * from starting early under programmatic dependent launch, so this is the timer that sees the step loop as it runs
 * in production (slgpu_timer_step_kernel then reports zero launches) */
slgpu_status slgpu_timer_start_coarse(slgpu_engine* e) { return timer_start(e, false); }

slgpu_status slgpu_timer_stop(slgpu_engine* e, double* elapsed_ms) {
  slgpu_status st = check_ready(e, false);
  if (st != SLGPU_OK) return st;
  if (!elapsed_ms) return fail(SLGPU_E_ARG, "elapsed_ms must not be NULL");
  CU_TRY(cudaEventRecord(e->t1, e->compute));
  CU_TRY(cudaEventSynchronize(e->t1));
  float ms = 0.f;
  CU_TRY(cudaEventElapsedTime(&ms, e->t0, e->t1));
  *elapsed_ms = ms;
  const bool per_launch = e->timing;
  e->timing = false;
  (void)per_launch;
  e->step_kernel_ms = 0.0;
  e->step_kernel_launches = e->step_ev.size() / 2;
  for (size_t i = 0; i + 1 < e->step_ev.size(); i += 2) {
    float k = 0.f;
    if (cudaEventElapsedTime(&k, e->step_ev[i], e->step_ev[i + 1]) == cudaSuccess) e->step_kernel_ms += k;
  }
  for (cudaEvent_t ev : e->step_ev) e->ev_time_pool.push_back(ev);
  e->step_ev.clear();
  return SLGPU_OK;
}

// ---------------------------------------------------------------- state inspection
slgpu_status slgpu_get_last_rows(slgpu_engine* e, double* rows) {
  slgpu_status st = check_ready(e, true);
  if (st != SLGPU_OK) return st;
  if (!rows) return fail(SLGPU_E_ARG, "rows must not be NULL");
  if ((st = sync_all(e)) != SLGPU_OK) return st;
  const size_t C = e->C, d = e->cfg.nDims, w = e->row_w;
  std::vector<double> x, en, rs, rb;
  std::vector<int32_t> ac, sw;
  if ((st = fetch(x, e->P.x, d * C)) != SLGPU_OK || (st = fetch(en, e->P.energy, C)) != SLGPU_OK ||
      (st = fetch(rs, e->P.row_sigma, C)) != SLGPU_OK || (st = fetch(rb, e->P.row_beta, C)) != SLGPU_OK ||
      (st = fetch(ac, e->P.accepted, C)) != SLGPU_OK || (st = fetch(sw, e->P.swap_type, C)) != SLGPU_OK)
    return st;
  for (size_t c = 0; c < C; ++c) {
    double* r = rows + c * w;
    for (size_t i = 0; i < d; ++i) r[i] = x[i * C + c];
    r[d] = en[c]; r[d + 1] = rs[c]; r[d + 2] = rb[c]; r[d + 3] = ac[c]; r[d + 4] = sw[c];
  }
  return SLGPU_OK;
}

slgpu_status slgpu_get_sigma_beta(slgpu_engine* e, double* sigma, double* beta) {
  slgpu_status st = check_ready(e, true);
  if (st != SLGPU_OK) return st;
  if ((st = sync_all(e)) != SLGPU_OK) return st;
  if (sigma) CU_TRY(cudaMemcpy(sigma, e->P.sigma, e->C * sizeof(double), cudaMemcpyDeviceToHost));
  if (beta) CU_TRY(cudaMemcpy(beta, e->P.beta, e->C * sizeof(double), cudaMemcpyDeviceToHost));
  return SLGPU_OK;
}

slgpu_status slgpu_get_shaper(slgpu_engine* e, uint32_t chain, double* L, double* N, double* count) {
  slgpu_status st = check_ready(e, true);
  if (st != SLGPU_OK) return st;
  if (chain >= e->C) return fail(SLGPU_E_ARG, "chain out of range");
  if ((st = sync_all(e)) != SLGPU_OK) return st;
  const size_t d = e->cfg.nDims, nL = d * (d + 1) / 2;
  std::vector<double> packed(nL);
  if (SLGPU_L_LAYOUT((uint32_t)d) == SLGPU_L_COLMAJOR) {
    std::vector<double> colp(nL);
    CU_TRY(cudaMemcpy(colp.data(), e->P.L + (size_t)chain * nL, nL * sizeof(double), cudaMemcpyDeviceToHost));
    for (size_t k = 0; k < d; ++k)
      for (size_t j = k; j < d; ++j) packed[j * (j + 1) / 2 + k] = colp[k * d - k * (k - 1) / 2 + (j - k)];
  } else {
    const uint32_t T = e->cfg.nTemps, spw = 32u / T, stack = chain / T;  // tile = stack group, slot = lane of the chain
    const double* lsrc = e->P.L + (size_t)(stack / spw) * nL * SLGPU_L_TILE + (stack % spw) * T + chain % T;
    CU_TRY(cudaMemcpy2D(packed.data(), sizeof(double), lsrc, SLGPU_L_TILE * sizeof(double), sizeof(double), nL, cudaMemcpyDeviceToHost));
  }
  double cnt = 0.0, ns = 0.0;
  CU_TRY(cudaMemcpy(&cnt, e->P.sh_count + chain, sizeof(double), cudaMemcpyDeviceToHost));
  CU_TRY(cudaMemcpy(&ns, e->P.nscale + chain, sizeof(double), cudaMemcpyDeviceToHost));
  if (count) *count = cnt;
  const bool fresh = e->cfg.welford ? (ns == 0.0) : (cnt == e->P.sh_count0);
  for (size_t r = 0; r < d; ++r)
    for (size_t k = 0; k < d; ++k) {
      const double l = (k <= r) ? packed[r * (r + 1) / 2 + k] : 0.0;
      if (L) L[r * d + k] = l;
      if (N) {
        double n;
        if (fresh) n = (r == k) ? ((e->cfg.mx[r] - e->cfg.mn[r]) / 4.0) / e->P.prop_norm : 0.0;
        else n = (k <= r) ? l * ns + (r == k ? 1e-4 : 0.0) : 0.0;  // adaptive.cpp:200-204
        N[r * d + k] = n;
      }
    }
  return SLGPU_OK;
}

slgpu_status slgpu_get_models(slgpu_engine* e, int which, double* mu_xx, double* mu_xy, double* weight, double* count) {
  slgpu_status st = check_ready(e, false);
  if (st != SLGPU_OK) return st;
  if (which != 0 && which != 1) return fail(SLGPU_E_ARG, "which must be 0 (sigma) or 1 (beta)");
  if ((st = sync_all(e)) != SLGPU_OK) return st;
  const uint32_t T = e->cfg.nTemps;
  std::vector<double> m;
  if ((st = fetch(m, e->d_models + (size_t)which * T * slgpu_tier::kModelStride, (size_t)T * slgpu_tier::kModelStride)) != SLGPU_OK) return st;
  for (uint32_t t = 0; t < T; ++t) {
    const double* p = m.data() + (size_t)t * slgpu_tier::kModelStride;
    if (mu_xx) std::memcpy(mu_xx + 9 * t, p, 9 * sizeof(double));
    if (mu_xy) std::memcpy(mu_xy + 3 * t, p + 9, 3 * sizeof(double));
    if (weight) std::memcpy(weight + 3 * t, p + 12, 3 * sizeof(double));
    if (count) count[t] = p[15];
  }
  return SLGPU_OK;
}

slgpu_status slgpu_get_ladder(slgpu_engine* e, double* ladder) {
  slgpu_status st = check_ready(e, false);
  if (st != SLGPU_OK) return st;
  if (!ladder) return fail(SLGPU_E_ARG, "ladder must not be NULL");
  if ((st = sync_all(e)) != SLGPU_OK) return st;
  CU_TRY(cudaMemcpy(ladder, e->d_ladder, e->cfg.nTemps * sizeof(double), cudaMemcpyDeviceToHost));
  return SLGPU_OK;
}

slgpu_status slgpu_get_counters(slgpu_engine* e, uint64_t* length, uint64_t* accepts, uint64_t* swaps, uint64_t* swap_attempts,
                                double* min_energy) {
  slgpu_status st = check_ready(e, true);
  if (st != SLGPU_OK) return st;
  if ((st = sync_all(e)) != SLGPU_OK) return st;
  const size_t C = e->C;
  std::vector<uint32_t> a, s, t;
  if ((st = fetch(a, e->P.lg_accepts, C)) != SLGPU_OK || (st = fetch(s, e->P.lg_swaps, C)) != SLGPU_OK ||
      (st = fetch(t, e->P.lg_swap_attempts, C)) != SLGPU_OK)
    return st;
  for (size_t c = 0; c < C; ++c) {
    if (length) length[c] = e->rounds_done + 1;  // TableLogger lengths start at 1 (logging.cpp:29-33)
    if (accepts) accepts[c] = a[c];
    if (swaps) swaps[c] = s[c];
    if (swap_attempts) swap_attempts[c] = t[c];
  }
  if (min_energy) CU_TRY(cudaMemcpy(min_energy, e->P.lg_min_energy, C * sizeof(double), cudaMemcpyDeviceToHost));
  return SLGPU_OK;
}

slgpu_status slgpu_rhat_stage1(slgpu_engine* e, double* sum_m, double* n_stacks, double* n_samples) {
  slgpu_status st = check_ready(e, true);
  if (st != SLGPU_OK) return st;
  if (!sum_m || !n_stacks || !n_samples) return fail(SLGPU_E_ARG, "output pointers must not be NULL");
  if ((st = sync_all(e)) != SLGPU_OK) return st;
  const size_t S = e->cfg.nStacks, d = e->cfg.nDims;
  std::vector<double> M;
  if ((st = fetch(M, e->P.ep_m, d * S)) != SLGPU_OK) return st;
  for (size_t i = 0; i < d; ++i) {
    double acc = 0.0;
    for (size_t s = 0; s < S; ++s) acc += M[i * S + s];
    sum_m[i] = acc;
  }
  *n_stacks = (double)S;
  *n_samples = (double)e->rounds_done;  // every stack has seen the same number of cold rows
  return SLGPU_OK;
}

slgpu_status slgpu_rhat_stage2(slgpu_engine* e, const double* mean, double* sum_dm2, double* sum_s) {
  slgpu_status st = check_ready(e, true);
  if (st != SLGPU_OK) return st;
  if (!mean || !sum_dm2 || !sum_s) return fail(SLGPU_E_ARG, "pointers must not be NULL");
  const size_t S = e->cfg.nStacks, d = e->cfg.nDims;
  std::vector<double> M, V;
  if ((st = fetch(M, e->P.ep_m, d * S)) != SLGPU_OK || (st = fetch(V, e->P.ep_s, d * S)) != SLGPU_OK) return st;
  for (size_t i = 0; i < d; ++i) {
    double b = 0.0, w = 0.0;
    for (size_t s = 0; s < S; ++s) {
      const double dm = M[i * S + s] - mean[i];
      b += dm * dm;
      w += V[i * S + s];
    }
    sum_dm2[i] = b;
    sum_s[i] = w;
  }
  return SLGPU_OK;
}

void slgpu_rhat_finish(const double* sum_dm2, const double* sum_s, double m, double n, uint32_t d, double* rhat) {
  // diagnostics.cpp:58-70 with B = n/(m-1) sum (M_s - mean)^2, W = (1/m) sum S_s/(n-1)
  for (uint32_t i = 0; i < d; ++i) {
    const double b = sum_dm2[i] * (n / (m - 1.0));
    const double w = (1.0 / m) * sum_s[i] / (n - 1.0);
    const double vHat = ((n - 1.0) / n) * w + (1.0 / n) * b;
    rhat[i] = std::sqrt(vHat / (w + 1e-30));
  }
}

slgpu_status slgpu_rhat(slgpu_engine* e, double* rhat) {
  slgpu_status st = check_ready(e, true);
  if (st != SLGPU_OK) return st;
  if (!rhat) return fail(SLGPU_E_ARG, "rhat must not be NULL");
  if ((st = sync_all(e)) != SLGPU_OK) return st;
  // EPSRDiagnostic::rHat, diagnostics.cpp:49-71, in the operation order of the oracle
  const size_t S = e->cfg.nStacks, d = e->cfg.nDims;
  std::vector<double> M, V;
  if ((st = fetch(M, e->P.ep_m, d * S)) != SLGPU_OK || (st = fetch(V, e->P.ep_s, d * S)) != SLGPU_OK) return st;
  const int nmin = (int)e->rounds_done;
  const int m = (int)S;
  for (size_t i = 0; i < d; ++i) {
    double mean = 0.0;
    for (size_t s = 0; s < S; ++s) mean += M[i * S + s];
    mean /= m;
    double b = 0.0, w = 0.0;
    for (size_t s = 0; s < S; ++s) {
      const double dm = M[i * S + s] - mean;
      b += dm * dm;
      w += (1.0 / m) * V[i * S + s] / (nmin - 1.0);
    }
    b *= nmin / (m - 1.0);
    const double vHat = ((nmin - 1.0) / nmin) * w + (1.0 / nmin) * b;
    rhat[i] = std::sqrt(vHat / (w + 1e-30));
  }
  return SLGPU_OK;
}

slgpu_status slgpu_summary(slgpu_engine* e, char* json, size_t cap, size_t* needed) {
  slgpu_status st = check_ready(e, true);
  if (st != SLGPU_OK) return st;
  const size_t C = e->C;
  const uint32_t S = e->cfg.nStacks, T = e->cfg.nTemps, d = e->cfg.nDims;
  std::vector<uint64_t> len(C), acc(C), sw(C), att(C);
  std::vector<double> sigma(C), beta(C), rhat(d), ladder(T);
  if ((st = slgpu_get_counters(e, len.data(), acc.data(), sw.data(), att.data(), nullptr)) != SLGPU_OK) return st;
  if ((st = slgpu_get_sigma_beta(e, sigma.data(), beta.data())) != SLGPU_OK) return st;
  if ((st = slgpu_get_ladder(e, ladder.data())) != SLGPU_OK) return st;
  const bool have_rhat = e->rounds_done >= 2 && S >= 2;
  if (have_rhat && (st = slgpu_rhat(e, rhat.data())) != SLGPU_OK) return st;
  std::ostringstream o;
  o.precision(10);
  auto tier = [&](const char* name, auto fn) {
    o << "\"" << name << "\":[";
    for (uint32_t t = 0; t < T; ++t) {
      double v = 0.0;
      for (uint32_t s = 0; s < S; ++s) v += fn((size_t)s * T + t);
      o << (t ? "," : "") << v / S;
    }
    o << "]";
  };
  o << "{\"nStacks\":" << S << ",\"nTemperatures\":" << T << ",\"nDims\":" << d << ",\"rounds\":" << e->rounds_done
    << ",\"rows_emitted\":" << e->rows_emitted << ",\"launches\":" << e->launches << ",";
  tier("accept_rate", [&](size_t c) { return (double)acc[c] / (double)len[c]; });  // GlbAcptRt, logging.cpp:72
  o << ",";
  tier("swap_rate", [&](size_t c) { return (double)sw[c] / (double)att[c]; });  // GlbSwapRt, logging.cpp:75
  o << ",";
  tier("sigma", [&](size_t c) { return sigma[c]; });
  o << ",";
  tier("beta", [&](size_t c) { return beta[c]; });
  o << ",\"ladder\":[";
  for (uint32_t t = 0; t < T; ++t) o << (t ? "," : "") << ladder[t];
  o << "],\"rhat\":[";
  if (have_rhat)
    for (uint32_t i = 0; i < d; ++i) o << (i ? "," : "") << rhat[i];
  o << "]}";
  const std::string s = o.str();
  if (needed) *needed = s.size() + 1;
  if (json && cap) {
    const size_t n = std::min(cap - 1, s.size());
    std::memcpy(json, s.data(), n);
    json[n] = 0;
  }
  return SLGPU_OK;
}

slgpu_status slgpu_eval_likelihood(slgpu_engine* e, const double* x, uint32_t n_points, double* nll) {
  slgpu_status st = check_ready(e, false);
  if (st != SLGPU_OK) return st;
  if (!x || !nll) return fail(SLGPU_E_ARG, "x and nll must not be NULL");
  if (!e->plug->launch_eval) return fail(SLGPU_E_PLUGIN, "plugin has no launch_eval");
  const size_t d = e->cfg.nDims, J = e->cfg.nJobTypes;
  double *dx = nullptr, *dn = nullptr;
  CU_TRY(cudaMalloc((void**)&dx, std::max<size_t>(1, n_points * d) * sizeof(double)));
  CU_TRY(cudaMalloc((void**)&dn, std::max<size_t>(1, n_points * J) * sizeof(double)));
  cudaError_t err = cudaMemcpy(dx, x, n_points * d * sizeof(double), cudaMemcpyHostToDevice);
  if (err == cudaSuccess) err = (cudaError_t)e->plug->launch_eval(&e->P, dx, n_points, dn, (void*)e->compute);
  if (err == cudaSuccess) err = cudaStreamSynchronize(e->compute);
  if (err == cudaSuccess) err = cudaMemcpy(nll, dn, n_points * J * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(dx);
  cudaFree(dn);
  e->launches++;
  if (err != cudaSuccess) return fail(SLGPU_E_CUDA, std::string("eval: ") + cudaGetErrorString(err));
  return SLGPU_OK;
}


// ---------------------------------------------------------------- table logger
slgpu_status slgpu_get_logger(slgpu_engine* e, double* cur_energy, double* accept_rate, double* swap_rate) {
  slgpu_status st = check_ready(e, true);
  if (st != SLGPU_OK) return st;
  if ((accept_rate || swap_rate) && !e->d_win_cnt)
    return fail(SLGPU_E_STATE, "the windowed rates need \"gpu\": {\"windowRates\": true}");
  if ((st = sync_all(e)) != SLGPU_OK) return st;
  const size_t C = e->C;
  if (cur_energy) CU_TRY(cudaMemcpy(cur_energy, e->P.lg_cur_energy, C * sizeof(double), cudaMemcpyDeviceToHost));
  if (accept_rate || swap_rate) {
    std::vector<uint32_t> cnt, sum;
    if ((st = fetch(cnt, e->d_win_cnt, 2 * C)) != SLGPU_OK || (st = fetch(sum, e->d_win_sum, 2 * C)) != SLGPU_OK) return st;
    // rates_ = window_sum / size-before-pop (adaptive.cpp:83-90); 0 until the first update (adaptive.cpp:38)
    auto rate = [&](size_t i) { return cnt[i] ? (double)sum[i] / (double)std::min<uint32_t>(cnt[i], slgpu_tier::kWindow) : 0.0; };
    for (size_t c = 0; c < C; ++c) {
      if (accept_rate) accept_rate[c] = rate(c);
      if (swap_rate) swap_rate[c] = rate(C + c);
    }
  }
  return SLGPU_OK;
}

slgpu_status slgpu_format_table(slgpu_engine* e, char* out, size_t cap, size_t* needed) {
  slgpu_status st = check_ready(e, true);
  if (st != SLGPU_OK) return st;
  const size_t C = e->C;
  const uint32_t S = e->cfg.nStacks, T = e->cfg.nTemps, d = e->cfg.nDims;
  std::vector<uint64_t> len(C), acc(C), sw(C), att(C);
  std::vector<double> minE(C), curE(C), sigma(C), beta(C), ar(C), sr(C), rhat(d);
  if ((st = slgpu_get_counters(e, len.data(), acc.data(), sw.data(), att.data(), minE.data())) != SLGPU_OK) return st;
  if ((st = slgpu_get_logger(e, curE.data(), ar.data(), sr.data())) != SLGPU_OK) return st;
  if ((st = slgpu_get_sigma_beta(e, sigma.data(), beta.data())) != SLGPU_OK) return st;
  if ((st = slgpu_rhat(e, rhat.data())) != SLGPU_OK) return st;
  // TableLogger::update, logging.cpp:52-87: fixed, precision 5, widths 4 / 9 / 10
  std::string s = "\n\n  ID    Length    MinEngy   CurrEngy      Sigma     AcptRt  GlbAcptRt       Beta     SwapRt  GlbSwapRt\n";
  s += std::string(102, '-') + "\n";
  char line[256];
  for (size_t i = 0; i < C; ++i) {
    if (i % T == 0 && i != 0) s += '\n';
    std::snprintf(line, sizeof line, "%4zu %9llu %10.5f %10.5f %10.5f %10.5f %10.5f %10.5f %10.5f %10.5f \n", i, (unsigned long long)len[i],
                  minE[i], curE[i], sigma[i], ar[i], (double)acc[i] / (double)len[i], beta[i], sr[i], (double)sw[i] / (double)att[i]);
    s += line;
  }
  s += '\n';  // std::endl after the table
  double mean = 0.0;
  bool converged = true;
  for (uint32_t i = 0; i < d; ++i) {
    mean += rhat[i];
    converged = converged && rhat[i] < 1.1;  // EPSRDiagnostic threshold, diagnostics.hpp:35
  }
  mean /= (double)d;
  std::snprintf(line, sizeof line, "Convergence test: %g (%s)\n", mean, converged ? "possibly converged" : "not converged");
  s += line;
  (void)S;
  if (needed) *needed = s.size() + 1;
  if (out && cap) {
    const size_t n = std::min(cap - 1, s.size());
    std::memcpy(out, s.data(), n);
    out[n] = 0;
  }
  return SLGPU_OK;
}

// ---------------------------------------------------------------- checkpoint / resume
// The reference declares ChainArray::recoverFromDisk (src/infer/chainarray.hpp:158-162) but never defines it; a
// stateline run cannot be resumed.  Here the whole sampler state is a handful of device arrays plus the round
// counter (the Philox streams are addressed by round index, so there is no generator state to save), and a
// resumed run continues bit for bit.
namespace {

struct CkHeader {
  char magic[8];
  uint32_t version, n_stacks, n_temps, n_dims, n_job_types, swap_interval, adapt_interval, stack_offset, l_layout, rng;
  uint32_t n_sections, reserved;
  uint64_t seed, rounds_done, bounds_hash;
  double opt_accept, opt_swap;
  char plugin[32];
};
struct CkSection {
  char name[24];
  uint64_t bytes;
};
constexpr char kCkMagic[8] = {'S', 'L', 'G', 'P', 'U', 'C', 'K', '1'};

uint64_t fnv1a(const void* data, size_t n, uint64_t h) {
  const unsigned char* p = (const unsigned char*)data;
  for (size_t i = 0; i < n; ++i) h = (h ^ p[i]) * 0x100000001b3ull;
  return h;
}

CkHeader ck_header(const slgpu_engine* e) {
  CkHeader h;
  std::memset(&h, 0, sizeof h);
  std::memcpy(h.magic, kCkMagic, 8);
  const Settings& c = e->cfg;
  h.version = 2;  // 2: L tiled by stack group (SLGPU_L_TILE 32)
  h.n_stacks = c.nStacks; h.n_temps = c.nTemps; h.n_dims = c.nDims; h.n_job_types = c.nJobTypes;
  h.swap_interval = c.swapInterval; h.adapt_interval = c.adaptInterval; h.stack_offset = c.stackOffset;
  h.l_layout = SLGPU_L_LAYOUT(c.nDims); h.rng = (uint32_t)c.rng;
  h.n_sections = (uint32_t)e->state.size();
  h.seed = c.seed;
  h.rounds_done = e->rounds_done;
  uint64_t bh = 0xcbf29ce484222325ull;
  bh = fnv1a(c.mn.data(), c.mn.size() * sizeof(double), bh);
  bh = fnv1a(c.mx.data(), c.mx.size() * sizeof(double), bh);
  h.bounds_hash = bh;
  h.opt_accept = c.optimalAcceptRate; h.opt_swap = c.optimalSwapRate;
  std::strncpy(h.plugin, e->plug->name ? e->plug->name : "", sizeof h.plugin - 1);
  return h;
}

struct FileCloser {
  FILE* f;
  ~FileCloser() { if (f) std::fclose(f); }
};

}  // namespace

slgpu_status slgpu_save_state(slgpu_engine* e, const char* path) {
  NvtxRange nvtx_("slgpu:save_state");
  slgpu_status st = check_ready(e, true);
  if (st != SLGPU_OK) return st;
  if (!path || !path[0]) return fail(SLGPU_E_ARG, "path must not be empty");
  if ((st = sync_all(e)) != SLGPU_OK) return st;
  const std::string tmp = std::string(path) + ".tmp";
  {
    FileCloser fc{std::fopen(tmp.c_str(), "wb")};
    if (!fc.f) return fail(SLGPU_E_IO, "cannot create " + tmp);
    const CkHeader h = ck_header(e);
    bool ok = std::fwrite(&h, sizeof h, 1, fc.f) == 1;
    std::vector<char> host;
    for (const StateBuf& b : e->state) {
      CkSection sec;
      std::memset(&sec, 0, sizeof sec);
      std::strncpy(sec.name, b.name, sizeof sec.name - 1);
      sec.bytes = b.bytes;
      host.resize(b.bytes);
      CU_TRY(cudaMemcpy(host.data(), b.dev, b.bytes, cudaMemcpyDeviceToHost));
      ok = ok && std::fwrite(&sec, sizeof sec, 1, fc.f) == 1 && (b.bytes == 0 || std::fwrite(host.data(), b.bytes, 1, fc.f) == 1);
    }
    ok = ok && std::fflush(fc.f) == 0;
    if (!ok) {
      std::remove(tmp.c_str());
      return fail(SLGPU_E_IO, "short write to " + tmp);
    }
  }
  if (std::rename(tmp.c_str(), path) != 0) return fail(SLGPU_E_IO, std::string("cannot rename ") + tmp + " to " + path);
  return SLGPU_OK;
}

slgpu_status slgpu_load_state(slgpu_engine* e, const char* path) {
  NvtxRange nvtx_("slgpu:load_state");
  slgpu_status st = check_ready(e, false);
  if (st != SLGPU_OK) return st;
  if (!path || !path[0]) return fail(SLGPU_E_ARG, "path must not be empty");
  if ((st = sync_all(e)) != SLGPU_OK) return st;
  FileCloser fc{std::fopen(path, "rb")};
  if (!fc.f) return fail(SLGPU_E_IO, std::string("cannot open ") + path);
  CkHeader h;
  if (std::fread(&h, sizeof h, 1, fc.f) != 1 || std::memcmp(h.magic, kCkMagic, 8) != 0 || h.version != 2)
    return fail(SLGPU_E_IO, std::string(path) + " is not a slgpu checkpoint");
  CkHeader want = ck_header(e);
  want.rounds_done = h.rounds_done;
  want.n_sections = h.n_sections;  // the rate windows are optional on either side (see below)
  if (std::memcmp(&h, &want, sizeof h) != 0)
    return fail(SLGPU_E_CONFIG, std::string(path) + " was written by a run with a different configuration (shape, bounds, seed, rates or plugin)");
  // read everything before touching the device: a truncated file must leave the engine as it was
  std::vector<std::pair<std::string, std::vector<char>>> file_secs(h.n_sections);
  for (auto& fs : file_secs) {
    CkSection sec;
    if (std::fread(&sec, sizeof sec, 1, fc.f) != 1 || sec.bytes > ((uint64_t)1 << 40)) return fail(SLGPU_E_IO, std::string(path) + " is truncated");
    sec.name[sizeof sec.name - 1] = 0;
    fs.first = sec.name;
    fs.second.resize(sec.bytes);
    if (sec.bytes && std::fread(fs.second.data(), sec.bytes, 1, fc.f) != 1) return fail(SLGPU_E_IO, std::string(path) + " is truncated");
  }
  // The logging windows ("win_*", gpu.windowRates) do not feed back into the chains: a checkpoint without them
  // resumes with empty windows, one with them loads into an engine that does not keep any.
  auto optional = [](const std::string& name) { return name.rfind("win_", 0) == 0; };
  std::vector<const std::vector<char>*> src(e->state.size(), nullptr);
  for (size_t i = 0; i < e->state.size(); ++i) {
    for (const auto& fs : file_secs)
      if (fs.first == e->state[i].name) src[i] = &fs.second;
    if (src[i] ? src[i]->size() != e->state[i].bytes : !optional(e->state[i].name))
      return fail(SLGPU_E_IO, std::string(path) + ": section " + e->state[i].name + " is missing or has the wrong size");
  }
  for (size_t i = 0; i < e->state.size(); ++i) {
    if (src[i]) CU_TRY(cudaMemcpy(e->state[i].dev, src[i]->data(), e->state[i].bytes, cudaMemcpyHostToDevice));
    else CU_TRY(cudaMemset(e->state[i].dev, 0, e->state[i].bytes));
  }
  e->rounds_done = h.rounds_done;
  e->initialised = true;
  e->force_key = true;  // the resumed stream starts with a key segment: the host has no earlier row to lean on
  return SLGPU_OK;
}

// development aid (tools/stamps.py): the device buffer of the per-round flags; not part of the documented ABI
void* slgpu_debug_round_flags(slgpu_engine* e) { return e ? (void*)e->d_round_flags : nullptr; }

size_t slgpu_format_row(const double* row, uint32_t d, char* out, size_t cap) { return format_row(row, d, out, cap); }

slgpu_status slgpu_measure_fp64_peak(int device, double* tflops) {
  if (!tflops) return fail(SLGPU_E_ARG, "tflops must not be NULL");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return fail(SLGPU_E_CUDA, "no such CUDA device");
  CU_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, device));
  double* sinkbuf = nullptr;
  CU_TRY(cudaMalloc((void**)&sinkbuf, sizeof(double)));
  cudaEvent_t a, b;
  CU_TRY(cudaEventCreate(&a));
  CU_TRY(cudaEventCreate(&b));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
  double best = 0.0;
  for (int rep = 0; rep < 6; ++rep) {
    CU_TRY(cudaEventRecord(a));
    k_dfma_peak<<<blocks, threads>>>(sinkbuf, iters, 1.0000001);
    CU_TRY(cudaEventRecord(b));
    CU_TRY(cudaEventSynchronize(b));
    float ms = 0.f;
    CU_TRY(cudaEventElapsedTime(&ms, a, b));
    const double flops = 2.0 * 16.0 * (double)iters * (double)blocks * threads;
    if (rep >= 2) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(sinkbuf);
  *tflops = best;
  return SLGPU_OK;
}

}  // extern "C"

// ---------------------------------------------------------------- several devices, one process
// The reference is one `stateline -c cfg` process that owns the whole run (src/bin/stateline.cpp:52-90,
// src/app/serverwrapper.cpp:44-138).  A group does that for several GPUs: the stacks are sharded contiguously, one
// engine per device with its own tier models (the same thing as that many stateline servers), one host
// thread per device for every call that queues work, and the end-of-run R-hat summed in this process.
struct slgpu_group {
  std::vector<slgpu_engine*> eng;
  uint32_t n_stacks = 0, n_dims = 0;
  uint64_t n_samples_total = 0;
};

namespace {
template <class F>
slgpu_status group_parallel(slgpu_group* g, F fn) {
  const size_t n = g->eng.size();
  std::vector<slgpu_status> st(n, SLGPU_OK);
  std::vector<std::string> err(n);
  std::vector<std::thread> th;
  for (size_t i = 0; i < n; ++i)
    th.emplace_back([&, i] {
      st[i] = fn(g->eng[i]);
      if (st[i] != SLGPU_OK) err[i] = g_err;  // the message lives in the worker thread
    });
  for (auto& t : th) t.join();
  for (size_t i = 0; i < n; ++i)
    if (st[i] != SLGPU_OK) return fail(st[i], "device " + std::to_string(g->eng[i]->cfg.device) + ": " + err[i]);
  return SLGPU_OK;
}
}  // namespace

extern "C" {

slgpu_status slgpu_group_create(const char* config_json, const char* plugin_path, const char* plugin_entry, const void* payload,
                                size_t payload_bytes, const int* devices, int n_devices, slgpu_group** out) {
  if (!config_json || !plugin_path || !out) return fail(SLGPU_E_ARG, "config_json, plugin_path and out must not be NULL");
  *out = nullptr;
  Settings cfg;
  try {
    cfg = parse_settings(config_json);
  } catch (const ConfigError& ce) {
    return fail(SLGPU_E_CONFIG, ce.msg);
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(SLGPU_E_CUDA, "no CUDA device: the engine has no CPU fallback");
  std::vector<int> dev;
  if (devices) {
    if (n_devices <= 0) return fail(SLGPU_E_ARG, "n_devices must be positive when a device list is given");
    dev.assign(devices, devices + n_devices);
  } else {
    const int n = n_devices > 0 ? n_devices : ndev;  // 0 = every device of the box
    for (int i = 0; i < n; ++i) dev.push_back(cfg.device + i);
  }
  for (int dv : dev)
    if (dv < 0 || dv >= ndev) return fail(SLGPU_E_CONFIG, "device " + std::to_string(dv) + " out of range (" + std::to_string(ndev) + " visible)");
  const uint32_t G = (uint32_t)dev.size();
  if (cfg.nStacks < G) return fail(SLGPU_E_CONFIG, "fewer stacks than devices");
  slgpu_group* g = new slgpu_group();
  g->n_stacks = cfg.nStacks;
  g->n_dims = cfg.nDims;
  g->n_samples_total = cfg.nSamplesTotal;
  // the stop rule counts cold samples over ALL stacks (serverwrapper.cpp:100-107): every shard runs the same rounds
  const uint64_t rounds = (cfg.nSamplesTotal + cfg.nStacks - 1) / cfg.nStacks;
  const uint32_t base = cfg.nStacks / G, rem = cfg.nStacks % G;
  uint32_t off = 0;
  for (uint32_t i = 0; i < G; ++i) {
    Settings c = cfg;
    c.nStacks = base + (i < rem ? 1u : 0u);
    c.device = dev[i];
    c.stackOffset = cfg.stackOffset + off;
    c.nSamplesTotal = rounds * c.nStacks;
    off += c.nStacks;
    slgpu_engine* e = nullptr;
    const slgpu_status st = create_from_settings(c, plugin_path, plugin_entry, payload, payload_bytes, &e);
    if (st != SLGPU_OK) {
      const std::string keep = g_err;
      slgpu_group_destroy(g);
      return fail(st, keep);
    }
    g->eng.push_back(e);
  }
  *out = g;
  return SLGPU_OK;
}

void slgpu_group_destroy(slgpu_group* g) {
  if (!g) return;
  for (slgpu_engine* e : g->eng) slgpu_destroy(e);
  delete g;
}

int slgpu_group_size(const slgpu_group* g) { return g ? (int)g->eng.size() : 0; }
slgpu_engine* slgpu_group_engine(slgpu_group* g, int i) { return (g && i >= 0 && i < (int)g->eng.size()) ? g->eng[i] : nullptr; }

slgpu_status slgpu_group_init_chains(slgpu_group* g) {
  if (!g) return fail(SLGPU_E_ARG, "group is NULL");
  return group_parallel(g, [](slgpu_engine* e) { return slgpu_init_chains(e); });
}
slgpu_status slgpu_group_step_rounds(slgpu_group* g, uint64_t n_rounds) {
  if (!g) return fail(SLGPU_E_ARG, "group is NULL");
  return group_parallel(g, [n_rounds](slgpu_engine* e) { return slgpu_step_rounds(e, n_rounds); });
}
slgpu_status slgpu_group_sync(slgpu_group* g) {
  if (!g) return fail(SLGPU_E_ARG, "group is NULL");
  return group_parallel(g, [](slgpu_engine* e) { return slgpu_sync(e); });
}
slgpu_status slgpu_group_wait(slgpu_group* g) {
  if (!g) return fail(SLGPU_E_ARG, "group is NULL");
  return group_parallel(g, [](slgpu_engine* e) { return slgpu_wait(e); });
}
slgpu_status slgpu_group_run(slgpu_group* g) {
  if (!g) return fail(SLGPU_E_ARG, "group is NULL");
  return group_parallel(g, [](slgpu_engine* e) { return slgpu_run(e); });
}
slgpu_status slgpu_group_save_state(slgpu_group* g, const char* path) {
  if (!g || !path || !path[0]) return fail(SLGPU_E_ARG, "group and path must not be empty");
  const std::string p = path;
  return group_parallel(g, [&p](slgpu_engine* e) { return slgpu_save_state(e, (p + ".stack" + std::to_string(e->cfg.stackOffset)).c_str()); });
}
slgpu_status slgpu_group_load_state(slgpu_group* g, const char* path) {
  if (!g || !path || !path[0]) return fail(SLGPU_E_ARG, "group and path must not be empty");
  const std::string p = path;
  return group_parallel(g, [&p](slgpu_engine* e) { return slgpu_load_state(e, (p + ".stack" + std::to_string(e->cfg.stackOffset)).c_str()); });
}
uint64_t slgpu_group_rounds_done(const slgpu_group* g) {
  uint64_t r = ~0ull;
  if (g)
    for (const slgpu_engine* e : g->eng) r = std::min(r, e->rounds_done);
  return (g && !g->eng.empty()) ? r : 0;
}

// EPSRDiagnostic::rHat (diagnostics.cpp:49-71) over the stacks of every device: the two-stage sums of slgpu_rhat_stage1/2
slgpu_status slgpu_group_rhat(slgpu_group* g, double* rhat) {
  if (!g || !rhat) return fail(SLGPU_E_ARG, "group and rhat must not be NULL");
  const uint32_t d = g->n_dims;
  std::vector<double> sum_m(d, 0.0), part(d), b(d, 0.0), w(d, 0.0), pb(d), pw(d), mean(d);
  double m_total = 0.0, n_samples = 0.0;
  for (slgpu_engine* e : g->eng) {
    double ns = 0.0, nn = 0.0;
    slgpu_status st = slgpu_rhat_stage1(e, part.data(), &ns, &nn);
    if (st != SLGPU_OK) return st;
    for (uint32_t i = 0; i < d; ++i) sum_m[i] += part[i];
    m_total += ns;
    n_samples = nn;
  }
  for (uint32_t i = 0; i < d; ++i) mean[i] = sum_m[i] / m_total;
  for (slgpu_engine* e : g->eng) {
    slgpu_status st = slgpu_rhat_stage2(e, mean.data(), pb.data(), pw.data());
    if (st != SLGPU_OK) return st;
    for (uint32_t i = 0; i < d; ++i) { b[i] += pb[i]; w[i] += pw[i]; }
  }
  slgpu_rhat_finish(b.data(), w.data(), m_total, n_samples, d, rhat);
  return SLGPU_OK;
}

slgpu_status slgpu_group_summary(slgpu_group* g, char* json, size_t cap, size_t* needed) {
  if (!g) return fail(SLGPU_E_ARG, "group is NULL");
  std::ostringstream o;
  o.precision(10);
  o << "{\"devices\":" << g->eng.size() << ",\"nStacks\":" << g->n_stacks << ",\"rhat\":[";
  uint64_t rounds = slgpu_group_rounds_done(g);
  if (rounds >= 2 && g->n_stacks >= 2) {
    std::vector<double> r(g->n_dims);
    slgpu_status st = slgpu_group_rhat(g, r.data());
    if (st != SLGPU_OK) return st;
    for (uint32_t i = 0; i < g->n_dims; ++i) o << (i ? "," : "") << r[i];
  }
  o << "],\"engines\":[";
  for (size_t k = 0; k < g->eng.size(); ++k) {
    size_t need = 0;
    slgpu_status st = slgpu_summary(g->eng[k], nullptr, 0, &need);
    if (st != SLGPU_OK) return st;
    std::vector<char> buf(need);
    if ((st = slgpu_summary(g->eng[k], buf.data(), buf.size(), nullptr)) != SLGPU_OK) return st;
    o << (k ? "," : "") << "{\"device\":" << g->eng[k]->cfg.device << ",\"stackOffset\":" << g->eng[k]->cfg.stackOffset << ",\"summary\":" << buf.data() << "}";
  }
  o << "]}";
  const std::string s = o.str();
  if (needed) *needed = s.size() + 1;
  if (json && cap) {
    const size_t n = std::min(cap - 1, s.size());
    std::memcpy(json, s.data(), n);
    json[n] = 0;
  }
  return SLGPU_OK;
}

}  // extern "C"
