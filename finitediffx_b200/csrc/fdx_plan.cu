// Plan lifetime and library introspection (include/fdx_b200.h).
//
// A plan is the device-resident form of one axis operator.  It plays the role of the reference's
// only cache -- the jax.jit cache around the coefficient solve, finitediffx/_src/utils.py:100 --
// except that the coefficients arrive already solved (exactly, in float64) from the host.
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <utility>
#include <vector>

extern char** environ;

#include "fdx_common.cuh"

namespace fdx {
thread_local const char* g_last_kernel = "";
std::atomic<uint64_t> g_epoch{1};
static std::atomic<int64_t> g_launches{0};
void count_launch(const char* family, int launches) {
  g_last_kernel = family;
  g_launches.fetch_add(launches, std::memory_order_relaxed);
}

int num_sms() {
  static std::atomic<int> cache[64];  // per device ordinal; 0 = not queried yet
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

// ---- snapshot of the FDX_* environment switches
using EnvTable = std::vector<std::pair<std::string, std::string>>;
static std::mutex g_env_mu;
static const EnvTable* snapshot_env() {
  auto* t = new EnvTable();
  for (char** e = environ; e && *e; ++e) {
    if (std::strncmp(*e, "FDX_", 4) != 0) continue;
    const char* eq = std::strchr(*e, '=');
    if (eq) t->emplace_back(std::string(*e, eq - *e), std::string(eq + 1));
  }
  return t;
}
static std::atomic<const EnvTable*> g_env_cur{nullptr};
static const EnvTable& env_table() {
  const EnvTable* t = g_env_cur.load(std::memory_order_acquire);
  if (!t) {
    std::lock_guard<std::mutex> lk(g_env_mu);
    t = g_env_cur.load(std::memory_order_acquire);
    if (!t) {
      t = snapshot_env();
      g_env_cur.store(t, std::memory_order_release);
    }
  }
  return *t;
}
const char* env_str(const char* name) {
  for (const auto& kv : env_table())
    if (kv.first == name) return kv.second.c_str();
  return nullptr;
}
int env_int(const char* name, int dflt) {
  const char* s = env_str(name);
  return s ? std::atoi(s) : dflt;
}
}  // namespace fdx

// Re-read the FDX_* switches (they are snapshotted at first use).  Not for concurrent use with launches that are
// reading an older snapshot's strings: tests and timing tools call it between launches.
extern "C" void fdx_reload_env(void) {
  std::lock_guard<std::mutex> lk(fdx::g_env_mu);
  fdx::g_env_cur.store(fdx::snapshot_env(), std::memory_order_release);  // the old table is leaked (a few bytes per call)
  fdx::g_epoch.fetch_add(1, std::memory_order_acq_rel);
}

extern "C" int fdx_plan_create(fdx_plan_t* plan, const fdx_axis_desc* d) {
  if (!plan || !d) return FDX_EINVAL;
  *plan = nullptr;
  const int taps = d->hi - d->lo + 1;
  if (d->n < 0 || taps < 1 || taps > FDX_MAX_TAPS) return FDX_EINVAL;
  if (d->nl < 0 || d->nr < 0 || d->wl < 0 || d->wr < 0) return FDX_EINVAL;
  if (d->nl + d->nr > d->n || d->wl > d->n || d->wr > d->n) return FDX_EINVAL;
  if ((d->nl > 0 && (!d->band_l || d->wl == 0)) || (d->nr > 0 && (!d->band_r || d->wr == 0))) return FDX_EINVAL;
  // main rows must be able to reach all their taps unless the caller supplies halos (slab plans
  // with nl == 0 / nr == 0 read the halo planes instead; the fused entry points check that).
  if (d->nl > 0 && d->nl < -d->lo) return FDX_EINVAL;
  if (d->nr > 0 && d->nr < d->hi) return FDX_EINVAL;

  fdx_plan_s* p = new (std::nothrow) fdx_plan_s();
  if (!p) return (int)cudaErrorMemoryAllocation;
  p->n = d->n, p->lo = d->lo, p->hi = d->hi;
  p->nl = d->nl, p->wl = d->wl, p->nr = d->nr, p->wr = d->wr;
  std::memcpy(p->main_taps, d->main_taps, sizeof(p->main_taps));
  p->band_l = p->band_r = p->d_main = nullptr;
  p->h_band_l = p->h_band_r = nullptr;
  cudaError_t e = cudaGetDevice(&p->device);
  if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_main, sizeof(p->main_taps));
  if (e == cudaSuccess) e = cudaMemcpy(p->d_main, p->main_taps, sizeof(p->main_taps), cudaMemcpyHostToDevice);
  const size_t bl = sizeof(double) * (size_t)d->nl * d->wl, br = sizeof(double) * (size_t)d->nr * d->wr;
  if (e == cudaSuccess && bl) e = cudaMalloc((void**)&p->band_l, bl);
  if (e == cudaSuccess && bl) e = cudaMemcpy(p->band_l, d->band_l, bl, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && br) e = cudaMalloc((void**)&p->band_r, br);
  if (e == cudaSuccess && br) e = cudaMemcpy(p->band_r, d->band_r, br, cudaMemcpyHostToDevice);
  auto host_copy = [&](double** dst, const double* src, size_t count) {
    if (e != cudaSuccess || !count) return;
    *dst = new (std::nothrow) double[count];
    if (!*dst) {
      e = cudaErrorMemoryAllocation;
      return;
    }
    std::memcpy(*dst, src, sizeof(double) * count);
  };
  host_copy(&p->h_band_l, d->band_l, (size_t)d->nl * d->wl);
  host_copy(&p->h_band_r, d->band_r, (size_t)d->nr * d->wr);
  if (e != cudaSuccess) {
    cudaFree(p->band_l);
    cudaFree(p->band_r);
    cudaFree(p->d_main);
    delete[] p->h_band_l;
    delete[] p->h_band_r;
    delete p;
    return (int)e;
  }
  *plan = p;
  return FDX_OK;
}

extern "C" int fdx_plan_destroy(fdx_plan_t plan) {
  if (!plan) return FDX_OK;
  fdx::g_epoch.fetch_add(1, std::memory_order_acq_rel);
  cudaFree(plan->band_l);
  cudaFree(plan->band_r);
  cudaFree(plan->d_main);
  delete[] plan->h_band_l;
  delete[] plan->h_band_r;
  delete plan;
  return FDX_OK;
}

extern "C" int fdx_version(void) { return 1000 * 0 + 1; }

extern "C" const char* fdx_error_string(int code) {
  switch (code) {
    case FDX_OK:
      return "ok";
    case FDX_EINVAL:
      return "fdx: invalid argument";
    case FDX_EUNSUPPORTED:
      return "fdx: unsupported shape/plan combination for this entry point";
    case FDX_EALIGN:
      return "fdx: pointer not aligned to its element size";
    default:
      return code > 0 ? cudaGetErrorString((cudaError_t)code) : "fdx: unknown error";
  }
}

extern "C" int64_t fdx_launch_count(void) { return fdx::g_launches.load(std::memory_order_relaxed); }
extern "C" const char* fdx_last_kernel(void) { return fdx::g_last_kernel; }
