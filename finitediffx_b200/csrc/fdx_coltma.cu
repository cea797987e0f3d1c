// `difference` along a STRIDED axis of a large array (outer x n x inner, inner contiguous): persistent TMA-staged kernel.
//
// The sibling of fdx_rowtma.cu behind the same entry point (fdx_axis_apply_*; reference:
// finitediffx/_src/finite_diff.py:168-297 with axis != last), for big arrays in place of axis_stream (fdx_axis.cu).
// Work items are (a strip of TX contiguous elements) x (a block of RB output rows along the axis); the resident CTAs
// claim them from a global counter, an item is one stage of a shared-memory ring:
//   * the producer warp issues `cp.async.bulk.tensor.3d` for a box of RB + W - 1 rows (the block and the rows its stencils
//     reach; W = taps) x TX elements.  The W - 1 halo rows are read again by the item below: an L2 hit, since
//     neighbouring items are claimed at about the same time;
//   * a consumer warp owns RW = RB / 4 consecutive output rows of the stage, a lane one 16-byte vector column: it walks
//     down its RW + W - 1 rows with a register window (one 128-bit shared load per row, every tap a packed FFMA2 --
//     along a strided axis the operand pairs are always aligned) and stores 512 bytes per warp and row;
//   * band rows (one-sided stencils at the two ends of the axis) by a warp of their own, arithmetic of band_element.
// Per output the taps are summed in ascending order with the fma chain of axis_stream: bit-identical results.
#include <algorithm>
#include <type_traits>

#include "fdx_tma.cuh"

namespace fdx {

namespace coltma {
using namespace tma;

template <typename T, int W_>
struct Geo {
  static constexpr int V = 16 / (int)sizeof(T), W = W_;
  static constexpr int TXV = 32, TX = TXV * V;  // a warp row: 512 bytes
  static constexpr int NW = 4, RW = 8, RB = NW * RW, NT = (NW + 2) * 32;
  static constexpr int BOX_Y = RB + W - 1;
  static constexpr int ROW_BYTES = TX * (int)sizeof(T);  // 512
  static constexpr int STAGE_BYTES = BOX_Y * ROW_BYTES;  // a multiple of 128
  static constexpr int STAGES = 4, CLAIM = 4;
  static constexpr int SMEM_BYTES = 256 + 128 + STAGES * STAGE_BYTES;
};

template <typename T>
struct Params {
  CUtensorMap tmap;  // (inner, n, outer)
  const T* x;
  T* out;
  T *out2, *out3;  // optional further destinations of the same values (nullptr: none)
  int64_t inner;
  int n, inner_v, outer, r0, r1, lo;
  int strips, rblocks;  // ceil(inner / TX), ceil((r1 - r0) / RB)
  unsigned int items;   // strips * rblocks * outer
  int64_t band_total;   // outer * (nl + nr) * inner
  unsigned int* ctr;
  Bands b;
  T alpha;
  Taps<T> taps;
};

template <int I, int N, typename F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

template <typename T, int W>
__global__ void __launch_bounds__(192, 2) axis_coltma_kernel(const __grid_constant__ Params<T> prm) {
  using G = Geo<T, W>;
  constexpr int V = G::V, STAGES = G::STAGES, RW = G::RW;
  using VT = Vec<T, V>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint32_t smem0 = smem_u32(smem_raw);
  const uint32_t ring = (smem0 + 256u + 127u) & ~127u;
  const uint32_t bar_full = smem0, bar_empty = smem0 + 8u * STAGES, item_xyz = smem0 + 128u;  // (strip, row block, outer) per stage
  const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  const int b = blockIdx.x, gsz = gridDim.x;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8u * s, 1);
      mbar_init(bar_empty + 8u * s, G::NW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&prm.tmap) : "memory");
  }
  __syncthreads();

  // ---- producer warp: claims items and fills the ring; the stage after the last item carries strip = -1
  if (wrp == G::NW) {
    if (lane != 0) return;
    unsigned int base = 0;
    int left = 0;
    const unsigned per_o = (unsigned)prm.strips * (unsigned)prm.rblocks;
    for (int k = 0;; ++k) {
      if (left == 0) base = atomicAdd(prm.ctr, (unsigned)G::CLAIM), left = G::CLAIM;
      const unsigned int t = base++;
      --left;
      const int s = k % STAGES;
      if (k >= STAGES) mbar_wait(bar_empty + 8u * s, (uint32_t)((k / STAGES) - 1) & 1u);
      if (t < prm.items) {
        const unsigned o = t / per_o, rem = t - o * per_o;
        const int rb = (int)(rem / (unsigned)prm.strips), strip = (int)(rem - (unsigned)rb * (unsigned)prm.strips);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(item_xyz + 16u * s), "r"(strip), "r"(rb), "r"((int)o), "r"(0) : "memory");
        mbar_expect_tx(bar_full + 8u * s, (uint32_t)G::STAGE_BYTES);
        tma_load_3d(ring + (uint32_t)s * G::STAGE_BYTES, &prm.tmap, bar_full + 8u * s, strip * G::TX, prm.r0 + rb * G::RB + prm.lo, (int)o);
      } else {
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(item_xyz + 16u * s), "r"(-1), "r"(-1), "r"(-1), "r"(0) : "memory");
        mbar_arrive(bar_full + 8u * s);
        break;
      }
    }
    __threadfence();
    if (atomicAdd(prm.ctr + 1, 1u) == (unsigned)gsz - 1u) {  // the last producer of the grid rewinds the counters
      prm.ctr[0] = 0u, prm.ctr[1] = 0u;
      __threadfence();
    }
    return;
  }

  // ---- band warp: the band rows, one element per lane (contiguous along `inner`), loads in batches of four
  if (wrp == G::NW + 1) {
    const int nb = prm.b.nl + prm.b.nr;
    for (int64_t g = (int64_t)b * 32 + lane; g < prm.band_total; g += (int64_t)gsz * 32) {
      const int64_t j = g % prm.inner;
      const int64_t rest = g / prm.inner;
      const int bi = (int)(rest % nb);
      const int64_t o = rest / nb;
      int64_t row, col0;
      int w;
      const double* coef;
      if (bi < prm.b.nl) {
        row = bi, col0 = 0, w = prm.b.wl, coef = prm.b.l + (int64_t)bi * prm.b.wl;
      } else {
        const int r = bi - prm.b.nl;
        row = prm.n - prm.b.nr + r, col0 = prm.n - prm.b.wr, w = prm.b.wr, coef = prm.b.r + (int64_t)r * prm.b.wr;
      }
      const T* xp = prm.x + (o * prm.n + col0) * prm.inner + j;
      T acc = 0;
      for (int c = 0; c < w; c += 4) {
        T cf[4], xv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int cu = min(c + u, w - 1);
          cf[u] = c + u < w ? (T)__ldg(coef + cu) : (T)0;
          xv[u] = __ldg(xp + (int64_t)cu * prm.inner);
        }
        // (a tap whose coefficient is exactly zero is skipped, as in band_element: 0 * inf must not become NaN)
#pragma unroll
        for (int u = 0; u < 4; ++u) acc = cf[u] != (T)0 ? fma(cf[u], xv[u], acc) : acc;
      }
      const T val = acc * prm.alpha;
      const int64_t at = (o * prm.n + row) * prm.inner + j;
      prm.out[at] = val;
      if (prm.out2) prm.out2[at] = val;
      if (prm.out3) prm.out3[at] = val;
    }
    return;
  }

  // ---- consumer warps
  int stage = 0;
  uint32_t parity = 0;
  for (;;) {
    mbar_wait(bar_full + 8u * stage, parity);
    int strip, rb, o, pad;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(strip), "=r"(rb), "=r"(o), "=r"(pad) : "r"(item_xyz + 16u * stage));
    if (strip < 0) break;
    const int jv = strip * G::TXV + lane;         // this lane's vector column
    const int i0 = prm.r0 + rb * G::RB + wrp * RW;  // first output row of this warp
    const uint32_t base = ring + (uint32_t)stage * G::STAGE_BYTES + (uint32_t)(wrp * RW * G::ROW_BYTES + lane * 16);
    const int64_t at0 = ((int64_t)o * prm.n + i0) * prm.inner + (int64_t)jv * V;
    const bool col_ok = jv < prm.inner_v;
    VT q[W];
    static_for<0, W - 1>([&](auto s) { q[s.value] = lds_vec<T, V>(base + (uint32_t)(s.value * G::ROW_BYTES)); });
    static_for<0, RW>([&](auto r_) {
      constexpr int r = r_.value;
      q[(W - 1 + r) % W] = lds_vec<T, V>(base + (uint32_t)((W - 1 + r) * G::ROW_BYTES));
      VT res;
      if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          unsigned long long a;
          asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(a) : "l"(pack2(prm.taps.c[0], prm.taps.c[0])), "l"(pack2(q[r % W].v[2 * h], q[r % W].v[2 * h + 1])));
          static_for<1, W>([&](auto k_) {
            constexpr int k = k_.value;
            asm("fma.rn.f32x2 %0, %1, %2, %0;"
                : "+l"(a)
                : "l"(pack2(prm.taps.c[k], prm.taps.c[k])), "l"(pack2(q[(r + k) % W].v[2 * h], q[(r + k) % W].v[2 * h + 1])));
          });
          unpack2(a, res.v[2 * h], res.v[2 * h + 1]);
        }
      } else {
#pragma unroll
        for (int v = 0; v < V; ++v) {
          T acc = prm.taps.c[0] * q[r % W].v[v];
          static_for<1, W>([&](auto k_) { acc = fma(prm.taps.c[k_.value], q[(r + k_.value) % W].v[v], acc); });
          res.v[v] = acc;
        }
      }
      if (col_ok && i0 + r < prm.r1) {
        const int64_t at = at0 + (int64_t)r * prm.inner;
        st_vec<T, V>(prm.out + at, res);
        if (prm.out2) st_vec<T, V>(prm.out2 + at, res);
        if (prm.out3) st_vec<T, V>(prm.out3 + at, res);
      }
    });
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty + 8u * stage);
    if (++stage == STAGES) stage = 0, parity ^= 1u;
  }
}

template <typename T, int W>
static int launch(const fdx_plan_s* p, const T* x, T* out, int64_t outer, int64_t inner, double alpha, cudaStream_t st, bool* band_done, T* out2, T* out3) {
  using G = Geo<T, W>;
  EncodeTiledFn enc = encode_fn();
  if (!enc) return FDX_EUNSUPPORTED;
  Params<T> prm;
  cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)p->n, (cuuint64_t)outer};
  cuuint64_t strides[2] = {(cuuint64_t)inner * sizeof(T), (cuuint64_t)inner * p->n * sizeof(T)};
  cuuint32_t box[3] = {(cuuint32_t)G::TX, (cuuint32_t)G::BOX_Y, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  if (enc(&prm.tmap, sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)x, dims, strides, box, estr,
          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return FDX_EUNSUPPORTED;
  const int r0 = p->nl, r1 = p->n - p->nr;
  const int64_t strips = (inner + G::TX - 1) / G::TX, rblocks = (r1 - r0 + G::RB - 1) / G::RB;
  const int64_t items = strips * rblocks * outer;
  if (items > 0x3fffffffLL || strips > 0x7fffffffLL / G::TX) return FDX_EUNSUPPORTED;
  prm.x = x, prm.out = out, prm.out2 = out2, prm.out3 = out3, prm.inner = inner, prm.n = p->n, prm.inner_v = (int)(inner / G::V), prm.outer = (int)outer;
  prm.r0 = r0, prm.r1 = r1, prm.lo = p->lo, prm.strips = (int)strips, prm.rblocks = (int)rblocks, prm.items = (unsigned)items;
  prm.b = bands_of(p), prm.alpha = (T)alpha, prm.taps = scaled_taps<T>(p, alpha);
  prm.band_total = outer * (int64_t)(p->nl + p->nr) * inner;
  prm.ctr = claim_slot(st);
  if (!prm.ctr) return FDX_EUNSUPPORTED;
  auto kern = axis_coltma_kernel<T, W>;
  {
    static std::atomic<unsigned long long> attr_done{0};
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = (dev >= 0 && dev < 64) ? (1ull << dev) : 0ull;
    if (!bit || !(attr_done.load(std::memory_order_acquire) & bit)) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES);
      if (e != cudaSuccess) return (int)e;
      attr_done.fetch_or(bit, std::memory_order_release);
    }
  }
  const int per_sm = std::max(1, std::min(4, env_int("FDX_COLTMA_CPS", (int)(224 * 1024 / (G::SMEM_BYTES + 1024)))));
  const int64_t grid = std::min<int64_t>(items, (int64_t)num_sms() * per_sm);
  kern<<<(unsigned)grid, G::NT, G::SMEM_BYTES, st>>>(prm);
  count_launch("axis_coltma");
  *band_done = true;
  FDX_LAUNCH_CHECK();
  return FDX_OK;
}

template <typename T>
static int dispatch(const fdx_plan_s* p, const T* x, T* out, int64_t outer, int64_t inner, double alpha, cudaStream_t st, bool* band_done, T* out2, T* out3) {
  switch (p->hi - p->lo + 1) {
#define FDX_CASE(W_) \
  case W_:           \
    return launch<T, W_>(p, x, out, outer, inner, alpha, st, band_done, out2, out3);
    FDX_CASE(2) FDX_CASE(3) FDX_CASE(4) FDX_CASE(5) FDX_CASE(6) FDX_CASE(7) FDX_CASE(8) FDX_CASE(9) FDX_CASE(10)
    FDX_CASE(11) FDX_CASE(12) FDX_CASE(13) FDX_CASE(14) FDX_CASE(15) FDX_CASE(16)
#undef FDX_CASE
    default:
      return FDX_EUNSUPPORTED;
  }
}

}  // namespace coltma

// Entry for fdx_axis.cu: FDX_EUNSUPPORTED = not this kernel's case (the caller goes on to axis_stream).
template <typename T>
int col_tma_apply(const fdx_plan_s* p, const T* x, T* out, int64_t outer, int64_t inner, double alpha, cudaStream_t st, bool* band_done, T* out2, T* out3) {
  constexpr int V = 16 / (int)sizeof(T);
  if (env_int("FDX_COLTMA_DISABLE", 0)) return FDX_EUNSUPPORTED;
  if (inner % V != 0 || (((uintptr_t)x | (uintptr_t)out | (uintptr_t)out2 | (uintptr_t)out3) & 15) != 0) return FDX_EUNSUPPORTED;
  if (p->n - p->nl - p->nr <= 0 || outer <= 0 || outer > 0x7fffffffLL || inner < 32 * V) return FDX_EUNSUPPORTED;
  const int64_t bytes = outer * (int64_t)p->n * inner * (int64_t)sizeof(T);
  if (bytes < (int64_t)env_int("FDX_COLTMA_MIN_MB", 8) * (1 << 20)) return FDX_EUNSUPPORTED;
  return coltma::dispatch<T>(p, x, out, outer, inner, alpha, st, band_done, out2, out3);
}
template int col_tma_apply<float>(const fdx_plan_s*, const float*, float*, int64_t, int64_t, double, cudaStream_t, bool*, float*, float*);
template int col_tma_apply<double>(const fdx_plan_s*, const double*, double*, int64_t, int64_t, double, cudaStream_t, bool*, double*, double*);

}  // namespace fdx
