// Launch accounting and optional per-launch-site CUDA-event timing (bench.py's roofline numbers
// come from these events, recorded on the stream the kernels are launched on).
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include <string.h>
#include "common.cuh"
#include "profile.cuh"

namespace drl {

static std::mutex g_mu;
static bool g_enabled = false;
static long long g_launches = 0;
struct Rec { const char* name; cudaEvent_t a, b; };
static std::vector<Rec> g_recs;
static std::vector<cudaEvent_t> g_pool;

void count_launch() { ++g_launches; }

static cudaEvent_t get_event() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

ProfileScope::ProfileScope(const char* name, cudaStream_t st) : name_(name), st_(st), on_(false) {
  if (!g_enabled) return;
  std::lock_guard<std::mutex> lk(g_mu);
  a_ = get_event();
  b_ = get_event();
  on_ = a_ && b_;
  if (on_) cudaEventRecord(a_, st_);
}
ProfileScope::~ProfileScope() {
  if (!on_) return;
  cudaEventRecord(b_, st_);
  std::lock_guard<std::mutex> lk(g_mu);
  g_recs.push_back(Rec{name_, a_, b_});
}

}  // namespace drl

using namespace drl;

extern "C" long long drl_launch_count(void) { return g_launches; }

extern "C" int drl_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_enabled = on != 0;
  return DRL_OK;
}

// Synchronises, sums the elapsed time per launch site since the last read, clears the records.
// names_out: max_sites x 48 chars; returns the number of sites through *n_sites.
extern "C" int drl_profile_read(int max_sites, char* names_out, float* total_ms, int* counts, int* n_sites) {
  if (!names_out || !total_ms || !counts || !n_sites) return DRL_ERR_BAD_ARG;
  DRL_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(g_mu);
  std::map<std::string, std::pair<double, int>> acc;
  std::vector<std::string> order;
  for (const Rec& r : g_recs) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    auto it = acc.find(r.name);
    if (it == acc.end()) { acc[r.name] = {ms, 1}; order.push_back(r.name); }
    else { it->second.first += ms; it->second.second += 1; }
    g_pool.push_back(r.a);
    g_pool.push_back(r.b);
  }
  g_recs.clear();
  int n = 0;
  for (const std::string& k : order) {
    if (n >= max_sites) break;
    strncpy(names_out + 48 * n, k.c_str(), 47);
    names_out[48 * n + 47] = 0;
    total_ms[n] = (float)acc[k].first;
    counts[n] = acc[k].second;
    ++n;
  }
  *n_sites = n;
  return DRL_OK;
}
