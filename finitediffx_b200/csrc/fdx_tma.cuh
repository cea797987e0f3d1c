// Shared pieces of the persistent TMA-ring axis kernels (fdx_rowtma.cu, fdx_coltma.cu): mbarrier / TMA PTX wrappers,
// the tensor-map encoder entry point and the per-device pool of claim counters.
#pragma once
#include <cuda.h>

#include <atomic>
#include <mutex>

#include "fdx_common.cuh"

namespace fdx {
namespace tma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
template <typename T, int V>
__device__ __forceinline__ Vec<T, V> lds_vec(uint32_t addr) {
  Vec<T, V> r;
  if constexpr (sizeof(T) == 4)
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]) : "r"(addr));
  else
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(r.v[0]), "=d"(r.v[1]) : "r"(addr));
  return r;
}
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(unsigned long long p, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p));
}

__host__ __device__ constexpr int cdiv(int a, int b) { return (a + b - 1) / b; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

// Counter pairs for the item claims: a pool per device.  Eager launches: ONE pair per stream (handed out at the
// stream's first launch) -- launches on a stream are ordered, so a pair is never used by two grids at once, and the grid
// that used it rewinds it to zero (the last producer).  Launches inside a stream capture: a pair of their own from a
// second region of the pool (a captured launch keeps its pair for every replay of its graph, so it must not share it
// with anything; a memset node in front of the kernel zeroes it at every replay).  nullptr -- the caller takes the
// non-persistent kernel instead -- beyond kSlots streams or kCapture captured launches on a device, or when the very
// first use on a device happens inside a capture (the pool is allocated and zeroed outside).
unsigned int* claim_slot(cudaStream_t st);  // fdx_rowtma.cu
#ifdef FDX_TMA_DEFINE_CLAIM_SLOT
unsigned int* claim_slot(cudaStream_t st) {
  constexpr int kSlots = 256, kCapture = 8192, kDevs = 64;
  struct Pool {
    unsigned int* base = nullptr;
    int used = 0, captured = 0;
    cudaStream_t owner[kSlots];
  };
  static Pool pool[kDevs];
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kDevs) return nullptr;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) return nullptr;
  const bool capturing = cs != cudaStreamCaptureStatusNone;
  std::lock_guard<std::mutex> lk(mu);
  Pool& p = pool[dev];
  if (!p.base) {
    if (capturing) return nullptr;
    unsigned int* b = nullptr;
    const size_t bytes = (size_t)(kSlots + kCapture) * 2 * sizeof(unsigned int);
    if (cudaMalloc(&b, bytes) != cudaSuccess) return nullptr;
    if (cudaMemset(b, 0, bytes) != cudaSuccess) {
      cudaFree(b);
      return nullptr;
    }
    p.base = b;
  }
  if (capturing) {
    if (p.captured >= kCapture) return nullptr;
    unsigned int* pair = p.base + 2 * (kSlots + p.captured);
    if (cudaMemsetAsync(pair, 0, 2 * sizeof(unsigned int), st) != cudaSuccess) return nullptr;  // (a node of the graph)
    ++p.captured;
    return pair;
  }
  for (int i = 0; i < p.used; ++i)
    if (p.owner[i] == st) return p.base + 2 * i;
  if (p.used >= kSlots) return nullptr;
  p.owner[p.used] = st;
  return p.base + 2 * p.used++;
}
#endif

}  // namespace tma
}  // namespace fdx
