// Shared device/host helpers for the fdx_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/fdx_b200.h"

// Device-resident plan (see include/fdx_b200.h :: fdx_axis_desc and finitediffx_b200/plan.py).
struct fdx_plan_s {
  int32_t n, lo, hi;
  int32_t nl, wl, nr, wr;
  double main_taps[FDX_MAX_TAPS];
  double* band_l;  // device, nl*wl
  double* band_r;  // device, nr*wr
  double* d_main;  // device copy of main_taps (kernels that index the taps at run time: an indexed constant-bank load is slow)
  double* h_band_l;  // host copies (the fused kernels carry the bands in their parameter block)
  double* h_band_r;
  int device;
};

namespace fdx {

// Number of SMs of the current device (148 on B200), queried once per device: the chunk / grid sizing of every kernel
// works in multiples of it.  (A build-time constant in round 1.)
int num_sms();

// Main taps travel by value in the kernel parameter block, pre-multiplied by alpha on the host in
// float64 and rounded once to T, so the inner loops are FFMA/DFMA with constant-bank operands.
template <typename T>
struct Taps {
  T c[FDX_MAX_TAPS];
};

template <typename T>
inline Taps<T> scaled_taps(const fdx_plan_s* p, double alpha) {
  Taps<T> t;
  for (int k = 0; k < FDX_MAX_TAPS; ++k) t.c[k] = (k <= p->hi - p->lo) ? (T)(p->main_taps[k] * alpha) : (T)0;
  return t;
}

// Boundary bands of one axis, as the kernels see them.
struct Bands {
  const double* l;
  const double* r;
  int32_t nl, wl, nr, wr;
};
inline Bands bands_of(const fdx_plan_s* p) { return Bands{p->band_l, p->band_r, p->nl, p->wl, p->nr, p->wr}; }

// ---- 16-byte vectors -------------------------------------------------------------------------
template <typename T, int V>
struct alignas(sizeof(T) * V) Vec {
  T v[V];
};

template <typename T, int V>
__device__ __forceinline__ Vec<T, V> ld_stream(const T* p) {
  // read-only, streamed once: non-coherent path, do not allocate in L1
  Vec<T, V> r;
  if constexpr (sizeof(T) * V == 16) {
    uint32_t a, b, c, d;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
    uint32_t w[4] = {a, b, c, d};
    memcpy(&r, w, 16);
  } else if constexpr (sizeof(T) * V == 8) {
    uint32_t a, b;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "l"(p));
    uint32_t w[2] = {a, b};
    memcpy(&r, w, 8);
  } else {
    static_assert(sizeof(T) * V == 4, "unsupported vector width");
    uint32_t a;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(a) : "l"(p));
    memcpy(&r, &a, 4);
  }
  return r;
}

template <typename T, int V>
__device__ __forceinline__ Vec<T, V> ld_cached(const T* p) {
  // read-only with reuse across neighbouring threads: keep in L1
  return *reinterpret_cast<const Vec<T, V>*>(__builtin_assume_aligned(p, sizeof(T) * V));
}

template <typename T, int V>
__device__ __forceinline__ void st_vec(T* p, const Vec<T, V>& r) {
  *reinterpret_cast<Vec<T, V>*>(__builtin_assume_aligned(p, sizeof(T) * V)) = r;
}

// The FDX_* environment switches (experiments, bit-equality tests, timing tools) are read ONCE, when the library is
// first used, into a snapshot; launches look them up there (usually an empty table) instead of calling getenv ~15
// times per launch.  fdx_reload_env() takes a new snapshot (tests and tools that flip switches at run time).
int env_int(const char* name, int dflt);
const char* env_str(const char* name);  // nullptr when unset

// Library epoch: bumped whenever a plan is destroyed (its address may be reused) or the switches are re-read; caches of
// launch parameters keyed by plan pointers are valid for one epoch only.
extern std::atomic<uint64_t> g_epoch;

extern thread_local const char* g_last_kernel;
void count_launch(const char* family, int launches = 1);

}  // namespace fdx

#define FDX_CUDA_TRY(expr)                  \
  do {                                      \
    cudaError_t _e = (expr);                \
    if (_e != cudaSuccess) return (int)_e;  \
  } while (0)

#define FDX_LAUNCH_CHECK()                  \
  do {                                      \
    cudaError_t _e = cudaPeekAtLastError(); \
    if (_e != cudaSuccess) return (int)_e;  \
  } while (0)
