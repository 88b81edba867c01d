// Error plumbing and launch accounting shared by every entry point of libb200mmor.so.
#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/b200_mmor.h"
#include "common.h"

namespace b200 {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

const char* last_error_cstr() { return g_last_error.c_str(); }

// process-wide tuning switches (common.h): -1 = not read from the environment yet, then 0 / 1
static std::atomic<int> g_pdl{-1}, g_decode_tiles{-1};
static bool env_switch(std::atomic<int>& v, const char* name) {
  int x = v.load(std::memory_order_relaxed);
  if (x < 0) {
    const char* e = getenv(name);
    x = (e != nullptr && e[0] != '\0' && e[0] != '0') ? 1 : 0;
    v.store(x, std::memory_order_relaxed);
  }
  return x != 0;
}
bool pdl_enabled() { return env_switch(g_pdl, "B200_PDL"); }
void set_pdl(bool on) { g_pdl.store(on ? 1 : 0, std::memory_order_relaxed); }
int decode_tiles_mode() {
  int x = g_decode_tiles.load(std::memory_order_relaxed);
  if (x < 0) {
    const char* e = getenv("B200_DECODE_TILES");
    // default 1: timed on B200 at 128 rows x 7B (tools/decode_bench.py, profiles/r2_decode_bench.json): 14.21 ms per
    // decode step against 14.85 with 128-column tiles (mode 0) and 14.36 with two CTAs per SM (mode 2)
    x = (e != nullptr && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1;
    g_decode_tiles.store(x, std::memory_order_relaxed);
  }
  return x;
}
static std::atomic<int> g_fused_rope{-1};
bool fused_rope_enabled() {
  int x = g_fused_rope.load(std::memory_order_relaxed);
  if (x < 0) {
    const char* e = getenv("B200_FUSED_ROPE");
    x = (e != nullptr && e[0] == '0') ? 0 : 1;      // default on
    g_fused_rope.store(x, std::memory_order_relaxed);
  }
  return x != 0;
}
void set_fused_rope(bool on) { g_fused_rope.store(on ? 1 : 0, std::memory_order_relaxed); }
void set_decode_tiles(int mode) { g_decode_tiles.store(mode < 0 ? 0 : (mode > 2 ? 2 : mode), std::memory_order_relaxed); }

// ---------------------------------------------------------------------------------------------
// launch counter + per-family event timing
// ---------------------------------------------------------------------------------------------
static std::atomic<long long> g_launches{0};
static std::atomic<int> g_prof_on{0};

struct ProfSlot {
  cudaEvent_t e0, e1;
  int family;
  bool used;
};
static std::mutex g_prof_mu;
static std::vector<ProfSlot> g_slots;      // event pool; [0, g_slots_used) hold the current measurement
static size_t g_slots_used = 0;
static double g_fam_bytes[kFamCount], g_fam_flops[kFamCount];
static long long g_fam_launches[kFamCount];

LaunchScope::LaunchScope(int family, cudaStream_t stream, double alg_bytes, double alg_flops, int kernels)
    : stream_(stream), slot_(-1) {
  g_launches.fetch_add(kernels, std::memory_order_relaxed);
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (g_slots_used == g_slots.size()) {
    ProfSlot s{};
    if (cudaEventCreate(&s.e0) != cudaSuccess || cudaEventCreate(&s.e1) != cudaSuccess) return;
    g_slots.push_back(s);
  }
  ProfSlot& s = g_slots[g_slots_used];
  s.family = family;
  s.used = true;
  slot_ = static_cast<int>(g_slots_used++);
  g_fam_bytes[family] += alg_bytes;
  g_fam_flops[family] += alg_flops;
  g_fam_launches[family] += kernels;
  cudaEventRecord(s.e0, stream);
}

LaunchScope::~LaunchScope() {
  if (slot_ < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEventRecord(g_slots[slot_].e1, stream_);
}

}  // namespace b200

using namespace b200;

extern "C" {

long long b200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int b200_set_option(const char* name, int value) {
  if (name == nullptr) return fail(-2, "b200_set_option: null name");
  const std::string n(name);
  if (n == "pdl") set_pdl(value != 0);
  else if (n == "decode_tiles") {
    if (value < 0 || value > 2) return fail(-2, "b200_set_option: decode_tiles takes 0, 1 or 2 (got %d)", value);
    set_decode_tiles(value);
  }
  else if (n == "fused_rope") set_fused_rope(value != 0);
  else return fail(-2, "b200_set_option: unknown option '%s' (pdl, decode_tiles, fused_rope)", name);
  return 0;
}
int b200_get_option(const char* name) {
  if (name == nullptr) return fail(-2, "b200_get_option: null name");
  const std::string n(name);
  if (n == "pdl") return pdl_enabled() ? 1 : 0;
  if (n == "decode_tiles") return decode_tiles_mode();
  if (n == "fused_rope") return fused_rope_enabled() ? 1 : 0;
  return fail(-2, "b200_get_option: unknown option '%s' (pdl, decode_tiles, fused_rope)", name);
}

int b200_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_slots_used = 0;
  for (int f = 0; f < kFamCount; ++f) {
    g_fam_bytes[f] = g_fam_flops[f] = 0.0;
    g_fam_launches[f] = 0;
  }
  g_prof_on.store(on ? 1 : 0);
  return 0;
}

int b200_prof_family_count(void) { return kFamCount; }

const char* b200_prof_family_name(int family) {
  static const char* names[kFamCount] = {"gemm",  "gemm_skinny", "flash_attn", "decode_attn", "norm",  "rope_kv",
                                         "embed", "argmax",      "patchify",   "segmask",     "misc",  "train",
                                         "pointcloud"};
  return family >= 0 && family < kFamCount ? names[family] : "";
}

int b200_prof_collect(double* ms, double* alg_bytes, double* alg_flops, long long* launches) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int f = 0; f < kFamCount; ++f) {
    ms[f] = 0.0;
    alg_bytes[f] = g_fam_bytes[f];
    alg_flops[f] = g_fam_flops[f];
    launches[f] = g_fam_launches[f];
  }
  for (size_t i = 0; i < g_slots_used; ++i) {
    ProfSlot& s = g_slots[i];
    B200_CUDA_OK(cudaEventSynchronize(s.e1));
    float t = 0.f;
    B200_CUDA_OK(cudaEventElapsedTime(&t, s.e0, s.e1));
    ms[s.family] += t;
  }
  return 0;
}

}  // extern "C"
