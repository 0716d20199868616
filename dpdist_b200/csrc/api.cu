// Library-level entry points: version, error reporting, device queries.
#include <stdarg.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include <nvtx3/nvToolsExt.h>   // header-only; resolves the tools library lazily, no-ops when no tool is attached
#include <stdlib.h>

namespace dpd {

char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

namespace {
std::atomic<long long> g_launches{0};
std::atomic<int> g_prof_on{0};
std::mutex g_prof_mu;
struct Pending { const char* name; cudaEvent_t e0, e1; };
std::vector<Pending> g_pending;
std::vector<cudaEvent_t> g_free_events;
struct Total { double ms = 0; long long n = 0; };
std::map<std::string, Total> g_totals;

cudaEvent_t get_event() {
  if (!g_free_events.empty()) { cudaEvent_t e = g_free_events.back(); g_free_events.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

// DPD_NVTX=1: every kernel launch of the library is wrapped in an NVTX range carrying its profile name, so that nsys /
// ncu timelines show the three stages of the path (3DmFV, voxel assignment, head layers) by name
static bool nvtx_on() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("DPD_NVTX"); v = (e && atoi(e) != 0) ? 1 : 0; }
  return v != 0;
}

ProfScope::ProfScope(const char* name, cudaStream_t st) : name_(name), st_(st), e0_(nullptr), e1_(nullptr), on_(false) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (nvtx_on()) nvtxRangePushA(name);
  if (g_prof_on.load(std::memory_order_relaxed)) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    e0_ = get_event(); e1_ = get_event();
    on_ = e0_ && e1_;
    if (on_) cudaEventRecord(e0_, st_);
  }
}

ProfScope::~ProfScope() {
  if (nvtx_on()) nvtxRangePop();
  if (on_) {
    cudaEventRecord(e1_, st_);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_pending.push_back({name_, e0_, e1_});
  }
}

}  // namespace dpd

extern "C" long long dpd_launch_count(void) { return dpd::g_launches.load(); }

extern "C" int dpd_profile_enable(int on) {
  dpd::g_prof_on.store(on ? 1 : 0);
  return 0;
}

extern "C" int dpd_profile_read(dpd_profile_entry* h_entries, int max_entries, int reset) {
  using namespace dpd;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& p : g_pending) {
    float ms = 0.f;
    if (cudaEventSynchronize(p.e1) == cudaSuccess && cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess) {
      Total& t = g_totals[p.name];
      t.ms += ms; t.n += 1;
    }
    g_free_events.push_back(p.e0); g_free_events.push_back(p.e1);
  }
  g_pending.clear();
  int n = 0;
  for (auto& kv : g_totals) {
    if (h_entries && n < max_entries) {
      memset(&h_entries[n], 0, sizeof(dpd_profile_entry));
      strncpy(h_entries[n].name, kv.first.c_str(), sizeof(h_entries[n].name) - 1);
      h_entries[n].ms = kv.second.ms; h_entries[n].launches = kv.second.n;
      ++n;
    }
  }
  if (reset) g_totals.clear();
  return n;
}

// CRC32C (Castagnoli, reflected polynomial 0x82F63B78) of a HOST buffer: the checksum TensorFlow's tensor-bundle
// checkpoints store per variable (dpdist_b200/tf_checkpoint.py verifies it when loading a reference model.ckpt).
extern "C" uint32_t dpd_crc32c(const void* h_data, size_t n) {
  static uint32_t table[256];
  static std::once_flag once;
  std::call_once(once, [] {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      table[i] = c;
    }
  });
  const unsigned char* p = (const unsigned char*)h_data;
  uint32_t c = 0xFFFFFFFFu;
  for (size_t i = 0; i < n; ++i) c = table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}

extern "C" int dpd_version(void) { return DPD_ABI_VERSION; }
extern "C" const char* dpd_last_error(void) { return dpd::last_error_buf(); }
