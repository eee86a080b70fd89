// Error reporting and library-level entry points of the C ABI (include/mintime_b200.h).
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <utility>
#include <vector>

#include "common.cuh"

namespace mt {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_status(cudaError_t e, const char* what) {
  set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  return (int)e;
}

namespace {
std::mutex g_once_mu;
std::vector<std::pair<int, const void*>> g_once;
}  // namespace

bool first_use_on_device(const void* key) {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(g_once_mu);
  for (const auto& e : g_once)
    if (e.first == dev && e.second == key) return false;
  g_once.emplace_back(dev, key);
  return true;
}

int current_sms() {
  int dev = 0, n = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n > 0 ? n : 148;
}

// ---------------------------------------------------------------------------------- diagnostics
namespace {
struct ProfRec {
  char name[64];
  double flops, bytes;
  cudaEvent_t e0, e1;
};
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof;
std::atomic<int> g_prof_on{0};
std::atomic<unsigned long long> g_launches{0};
}  // namespace

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

ProfScope::ProfScope(cudaStream_t stream, double flops, double bytes, const char* fmt, ...) : slot(-1), st(stream) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  ProfRec r;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(r.name, sizeof(r.name), fmt, ap);
  va_end(ap);
  r.flops = flops;
  r.bytes = bytes;
  if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
  cudaEventRecord(r.e0, st);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.push_back(r);
  slot = (int)g_prof.size() - 1;
}

ProfScope::~ProfScope() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (slot < (int)g_prof.size()) cudaEventRecord(g_prof[slot].e1, st);
}

}  // namespace mt

extern "C" void mt_prof_enable(int on) { mt::g_prof_on.store(on ? 1 : 0); }

extern "C" unsigned long long mt_prof_launch_count(void) { return mt::g_launches.load(); }

extern "C" void mt_prof_reset(void) {
  std::lock_guard<std::mutex> lk(mt::g_prof_mu);
  for (auto& r : mt::g_prof) {
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  mt::g_prof.clear();
}

// Synchronises on the recorded events, aggregates by name (sorted by total time, descending) and
// returns the number of distinct names (at most `max_entries` are written).
extern "C" int mt_prof_collect(mt_prof_entry_t* out, int max_entries) {
  std::lock_guard<std::mutex> lk(mt::g_prof_mu);
  std::vector<mt_prof_entry_t> agg;
  for (auto& r : mt::g_prof) {
    if (cudaEventSynchronize(r.e1) != cudaSuccess) continue;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) continue;
    mt_prof_entry_t* e = nullptr;
    for (auto& a : agg)
      if (!strcmp(a.name, r.name)) { e = &a; break; }
    if (!e) {
      mt_prof_entry_t n;
      memset(&n, 0, sizeof(n));
      strncpy(n.name, r.name, sizeof(n.name) - 1);
      agg.push_back(n);
      e = &agg.back();
    }
    e->ms_total += ms;
    e->flops_total += r.flops;
    e->bytes_total += r.bytes;
    e->count += 1;
  }
  std::sort(agg.begin(), agg.end(), [](const mt_prof_entry_t& a, const mt_prof_entry_t& b) { return a.ms_total > b.ms_total; });
  for (int i = 0; i < (int)agg.size() && i < max_entries; ++i) out[i] = agg[i];
  return (int)agg.size();
}

extern "C" int mt_abi_version(void) { return MT_ABI_VERSION; }

extern "C" const char* mt_last_error(void) { return mt::g_err; }

extern "C" int mt_device_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return mt::cuda_status(e, "cudaGetDevice");
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return mt::cuda_status(e, "cudaDeviceGetAttribute");
  if (major != 10) {
    mt::set_error("libmintime_b200 is built for sm_100a only; device %d has compute capability %d.x", dev, major);
    return MT_ERR_UNSUPPORTED;
  }
  return MT_OK;
}
