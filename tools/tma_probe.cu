// Micro-benchmark: how fast can halo boxes be streamed through shared memory on B200?
// Copies a (n0,n1,n2) fp32 grid plane by plane: each CTA owns a (TY x TX) tile and a z chunk and loads
// the haloed box (TY+2P) x (TX+2HX) of every plane either with TMA or with cp.async, then writes the
// tile interior back with 128-bit stores.  No stencil arithmetic: this is the ceiling of the load path.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tma_probe tools/tma_probe.cu
//   ./tools/tma_probe [n]
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(
          smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

struct Params {
  CUtensorMap tmap;
  const float* in;
  float* out;
  int n0, n1, n2, chunk, tiles_x, tiles_y;
};

// MODE 0: TMA box per plane.  MODE 1: cp.async 16-byte cooperative loads (zero fill out of range).
template <int TX, int TY, int P, int HX, int STAGES, int NT, int MODE>
__global__ void __launch_bounds__(NT) probe(const __grid_constant__ Params prm) {
  constexpr int BX = TX + 2 * HX, BY = TY + 2 * P;
  constexpr int STAGE_BYTES = ((BX * BY * 4 + 127) / 128) * 128;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_bar = full_bar + STAGES;
  const uint32_t raw = smem_u32(smem);
  const uint32_t ring = ((raw + 256 + 127) & ~127u) - raw;
  const int tid = threadIdx.x;
  int b = blockIdx.x;
  const int tile_x = b % prm.tiles_x;
  b /= prm.tiles_x;
  const int tile_y = b % prm.tiles_y;
  const int chunk = b / prm.tiles_y;
  const int x0 = tile_x * TX, y0 = tile_y * TY;
  const int zs = chunk * prm.chunk, ze = min(prm.n0, zs + prm.chunk);
  const int n_planes = ze - zs;
  const int64_t plane = (int64_t)prm.n1 * prm.n2;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], MODE == 0 ? 1 : NT);
      mbar_init(&empty_bar[s], NT / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int it) {
    const int s = it % STAGES;
    unsigned char* dst = smem + ring + (size_t)s * STAGE_BYTES;
    if constexpr (MODE == 0) {
      if (tid == 0) {
        if (it >= STAGES) mbar_wait(&empty_bar[s], ((it / STAGES) - 1) & 1);
        mbar_expect_tx(&full_bar[s], BX * BY * 4);
        tma_load_3d(dst, &prm.tmap, &full_bar[s], x0 - HX, y0 - P, zs + it);
      }
    } else {
      if (it >= STAGES) mbar_wait(&empty_bar[s], ((it / STAGES) - 1) & 1);
      const float* src_plane = prm.in + (int64_t)(zs + it) * plane;
      constexpr int VPR = BX / 4;  // 16-byte vectors per box row
      for (int i = tid; i < VPR * BY; i += NT) {
        const int r = i / VPR, c = i - r * VPR;
        const int gy = y0 - P + r, gx = x0 - HX + c * 4;
        const bool ok = gy >= 0 && gy < prm.n1 && gx >= 0 && gx < prm.n2;
        const float* src = ok ? src_plane + (int64_t)gy * prm.n2 + gx : prm.in;
        const uint32_t d = smem_u32(dst + (r * BX + c * 4) * 4);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(ok ? 16 : 0) : "memory");
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full_bar[s])) : "memory");
    }
  };
  for (int it = 0; it < STAGES - 1 && it < n_planes; ++it) issue(it);
  constexpr int TXV = TX / 4;
  const int tx = tid % TXV, ty = tid / TXV;
  constexpr int ROWS_PER_PASS = NT / TXV;
  for (int it = 0; it < n_planes; ++it) {
    const int s = it % STAGES;
    if (it + STAGES - 1 < n_planes) issue(it + STAGES - 1);
    mbar_wait(&full_bar[s], (it / STAGES) & 1);
    const float* t = reinterpret_cast<const float*>(smem + ring + (size_t)s * STAGE_BYTES);
    float4 v[(TY + ROWS_PER_PASS - 1) / ROWS_PER_PASS];
#pragma unroll
    for (int k = 0; k < (TY + ROWS_PER_PASS - 1) / ROWS_PER_PASS; ++k) {
      const int r = ty + k * ROWS_PER_PASS;
      if (r < TY) v[k] = *reinterpret_cast<const float4*>(t + (r + P) * BX + HX + tx * 4);
    }
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&empty_bar[s]);
    float* o = prm.out + (int64_t)(zs + it) * plane;
#pragma unroll
    for (int k = 0; k < (TY + ROWS_PER_PASS - 1) / ROWS_PER_PASS; ++k) {
      const int r = ty + k * ROWS_PER_PASS;
      const int gy = y0 + r, gx = x0 + tx * 4;
      if (r < TY && gy < prm.n1 && gx < prm.n2) *reinterpret_cast<float4*>(o + (int64_t)gy * prm.n2 + gx) = v[k];
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int TX, int TY, int P, int HX, int STAGES, int NT, int MODE>
void run(const char* name, const float* in, float* out, int n, int chunk, int promo) {
  constexpr int BX = TX + 2 * HX, BY = TY + 2 * P;
  constexpr int STAGE_BYTES = ((BX * BY * 4 + 127) / 128) * 128;
  const int smem_bytes = 512 + STAGES * STAGE_BYTES;
  Params prm;
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  cuuint64_t dims[3] = {(cuuint64_t)n, (cuuint64_t)n, (cuuint64_t)n};
  cuuint64_t strides[2] = {(cuuint64_t)n * 4, (cuuint64_t)n * n * 4};
  cuuint32_t box[3] = {BX, BY, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = ((EncodeTiledFn)fp)(&prm.tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)in, dims, strides, box, es,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)promo,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("%s: tensor map encode failed (%d)\n", name, (int)r);
    return;
  }
  prm.in = in, prm.out = out, prm.n0 = prm.n1 = prm.n2 = n, prm.chunk = chunk;
  prm.tiles_x = (n + TX - 1) / TX, prm.tiles_y = (n + TY - 1) / TY;
  const int blocks = prm.tiles_x * prm.tiles_y * ((n + chunk - 1) / chunk);
  auto k = probe<TX, TY, P, HX, STAGES, NT, MODE>;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, NT, smem_bytes));
  for (int i = 0; i < 2; ++i) k<<<blocks, NT, smem_bytes>>>(prm);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int reps = 10;
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) k<<<blocks, NT, smem_bytes>>>(prm);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= reps;
  const double gb = 2.0 * n * (double)n * n * 4 / 1e9;
  printf("%-44s box %3dx%3d chunk %3d ctas %6d occ %d  %.3f ms  %.0f GB/s (algorithmic), box rows/us/SM %.1f\n", name, BX, BY,
         chunk, blocks, occ, ms, gb / ms * 1e3, (double)blocks / 148 * (chunk)*BY / (ms * 1e3));
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 512;
  float *in, *out;
  CK(cudaMalloc(&in, (size_t)n * n * n * 4));
  CK(cudaMalloc(&out, (size_t)n * n * n * 4));
  CK(cudaMemset(in, 1, (size_t)n * n * n * 4));
  // reference: plain device-to-device copy
  {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    CK(cudaMemcpy(out, in, (size_t)n * n * n * 4, cudaMemcpyDeviceToDevice));
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) CK(cudaMemcpyAsync(out, in, (size_t)n * n * n * 4, cudaMemcpyDeviceToDevice));
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("cudaMemcpy D2D: %.3f ms  %.0f GB/s\n", ms / 10, 2.0 * n * (double)n * n * 4 / 1e9 / (ms / 10) * 1e3);
  }
  const int P256 = (int)CU_TENSOR_MAP_L2_PROMOTION_L2_256B, P128 = (int)CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            P0 = (int)CU_TENSOR_MAP_L2_PROMOTION_NONE;
  //            TX   TY  P HX S  NT  MODE
  run<64, 32, 3, 4, 4, 256, 0>("tma  64x32 halo3 S4 nt256 promo256", in, out, n, 32, P256);
  run<64, 32, 3, 4, 4, 256, 0>("tma  64x32 halo3 S4 nt256 promo128", in, out, n, 32, P128);
  run<64, 32, 3, 4, 4, 256, 0>("tma  64x32 halo3 S4 nt256 promo0", in, out, n, 32, P0);
  run<64, 32, 0, 0, 4, 256, 0>("tma  64x32 nohalo S4 nt256", in, out, n, 32, P256);
  run<64, 32, 3, 0, 4, 256, 0>("tma  64x32 yhalo only S4 nt256", in, out, n, 32, P256);
  run<64, 32, 0, 4, 4, 256, 0>("tma  64x32 xhalo only S4 nt256", in, out, n, 32, P256);
  run<64, 32, 3, 4, 8, 256, 0>("tma  64x32 halo3 S8 nt256", in, out, n, 32, P256);
  run<64, 32, 3, 4, 4, 256, 0>("tma  64x32 halo3 S4 nt256 chunk128", in, out, n, 128, P256);
  run<64, 64, 3, 4, 4, 256, 0>("tma  64x64 halo3 S4 nt256", in, out, n, 32, P256);
  run<128, 32, 3, 4, 4, 256, 0>("tma 128x32 halo3 S4 nt256", in, out, n, 32, P256);
  run<128, 16, 3, 4, 4, 256, 0>("tma 128x16 halo3 S4 nt256", in, out, n, 32, P256);
  run<248, 16, 3, 4, 4, 256, 0>("tma 248x16 halo3 S4 nt256 (box 256 wide)", in, out, n, 32, P256);
  run<64, 16, 3, 4, 6, 128, 0>("tma  64x16 halo3 S6 nt128", in, out, n, 32, P256);
  run<64, 32, 3, 4, 4, 256, 1>("cpasync  64x32 halo3 S4 nt256", in, out, n, 32, P256);
  run<64, 32, 3, 4, 8, 256, 1>("cpasync  64x32 halo3 S8 nt256", in, out, n, 32, P256);
  run<64, 64, 3, 4, 4, 256, 1>("cpasync  64x64 halo3 S4 nt256", in, out, n, 32, P256);
  run<128, 32, 3, 4, 4, 256, 1>("cpasync 128x32 halo3 S4 nt256", in, out, n, 32, P256);
  run<64, 32, 3, 4, 4, 256, 1>("cpasync  64x32 halo3 S4 nt256 chunk128", in, out, n, 128, P256);
  run<64, 16, 3, 4, 6, 128, 1>("cpasync  64x16 halo3 S6 nt128", in, out, n, 32, P256);
  return 0;
}
