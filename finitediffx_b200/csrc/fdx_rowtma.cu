// `difference` along the CONTIGUOUS axis of a large array: persistent TMA-staged row kernel.
//
// Replaces, for big arrays, the L1-gather form of axis_row (fdx_axis.cu) behind the same entry point
// (fdx_axis_apply_*; reference: finitediffx/_src/finite_diff.py:168-297 with axis = last).  The rows of the array
// are independent, so the work is a flat list of items -- (a strip of TX columns) x (a block of TYR rows) -- that the
// resident CTAs CLAIM from a global counter (a persistent grid under a saturated DRAM is not served evenly: with a static
// round-robin split the slowest SM finished 40 % behind the fastest); an item is one stage of a shared-memory ring:
//   * a producer warp (one lane) claims items four at a time, writes each item's coordinates next to its stage and
//     issues `cp.async.bulk.tensor.2d` (box = TYR rows x (TX + halo vectors) columns; columns outside the row are
//     zero-filled by the TMA unit), mbarrier full / empty pairs; it runs ahead as far as the ring allows;
//   * four consumer warps: a warp owns RW rows of the stage, a lane one 16-byte vector per row -- its window comes from
//     shared memory with aligned 128-bit loads (no tag lookups, no per-thread address arithmetic), the taps are FFMA2
//     where the operand pair is an aligned register pair, the row segment of a warp leaves as one coalesced 512-byte store;
//   * the ring never drains: there are no halos between items, so the CTA's prologue latency is paid once.
// The band columns (one-sided stencils next to the two ends of the axis) are computed by a sixth warp of the same CTAs,
// one element per lane, with the arithmetic of the other axis kernels (band_element).  Per output the taps are summed in
// ascending order with the same fma chain as axis_row: the two kernels agree bit for bit.
#include <algorithm>

#define FDX_TMA_DEFINE_CLAIM_SLOT
#include "fdx_tma.cuh"

namespace fdx {

namespace rowtma {
using namespace tma;

template <typename T, int LO, int HI>
struct Geo {
  static constexpr int V = 16 / (int)sizeof(T);
  static constexpr int HLV = LO < 0 ? cdiv(-LO, V) : 0, HRV = HI > 0 ? cdiv(HI, V) : 0;  // halo vectors left / right
  static constexpr int NV = HLV + 1 + HRV;                                                // window vectors of a lane
  static constexpr int TXV = 32, TX = TXV * V;                                            // a warp row: 512 bytes of outputs
  static constexpr int BOX_X = TX + (HLV + HRV) * V;
  static constexpr int NW = 4, RW = 2, TYR = NW * RW, NT = (NW + 2) * 32;  // four consumer warps + producer + band warp
  static constexpr int STAGE_BYTES = ((BOX_X * TYR * (int)sizeof(T) + 127) / 128) * 128;
  static constexpr int STAGES = 6, CLAIM = 4;  // items per claim of the global counter
  static constexpr int SMEM_BYTES = 256 + 128 + STAGES * STAGE_BYTES;  // barriers, item coordinates, slack, ring
  static constexpr int W = HI - LO + 1;
};

template <typename T>
struct Params {
  CUtensorMap tmap;
  const T* x;
  T* out;
  T *out2, *out3;  // optional further destinations of the same values (nullptr: none)
  int rows;  // outer
  int n, n_v, r0, r1;
  int strips;       // ceil(n / TX)
  int items;        // strips * ceil(rows / TYR)
  int64_t band_total;  // band elements of the whole array (0: none ride along)
  unsigned int* ctr;   // [0] next unclaimed item, [1] producers that have finished; both back to 0 when the grid is done
  Bands b;
  T alpha;
  Taps<T> taps;
};

template <typename T, int LO, int HI>
__global__ void __launch_bounds__(192, 5) axis_rowtma_kernel(const __grid_constant__ Params<T> prm) {
  using G = Geo<T, LO, HI>;
  constexpr int V = G::V, W = G::W, STAGES = G::STAGES;
  using VT = Vec<T, V>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint32_t smem0 = smem_u32(smem_raw);
  const uint32_t ring = (smem0 + 256u + 127u) & ~127u;
  const uint32_t bar_full = smem0, bar_empty = smem0 + 8u * STAGES, item_xy = smem0 + 128u;  // (strip, row block) per stage
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
    for (int k = 0;; ++k) {
      if (left == 0) base = atomicAdd(prm.ctr, (unsigned)G::CLAIM), left = G::CLAIM;
      const unsigned int t = base++;
      --left;
      const int s = k % STAGES;
      if (k >= STAGES) mbar_wait(bar_empty + 8u * s, (uint32_t)((k / STAGES) - 1) & 1u);
      if (t < (unsigned)prm.items) {
        const int rb = (int)(t / (unsigned)prm.strips), strip = (int)(t - (unsigned)rb * (unsigned)prm.strips);
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(item_xy + 8u * s), "r"(strip), "r"(rb) : "memory");
        mbar_expect_tx(bar_full + 8u * s, (uint32_t)(G::BOX_X * G::TYR * sizeof(T)));
        tma_load_2d(ring + (uint32_t)s * G::STAGE_BYTES, &prm.tmap, bar_full + 8u * s, strip * G::TX - G::HLV * V, rb * G::TYR);
      } else {
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(item_xy + 8u * s), "r"(-1), "r"(-1) : "memory");
        mbar_arrive(bar_full + 8u * s);
        break;
      }
    }
    // every producer has made its last claim once all of them are here: the last one rewinds the counters for the
    // next launch that uses this slot
    __threadfence();
    if (atomicAdd(prm.ctr + 1, 1u) == (unsigned)gsz - 1u) {
      prm.ctr[0] = 0u, prm.ctr[1] = 0u;
      __threadfence();
    }
    return;
  }

  // The band columns (a handful of elements per row, each a chain of dependent loads: ~10 us a piece from cold DRAM) are
  // the job of a warp of their own: the ring never waits for them.  Loads in batches of four.
  if (wrp == G::NW + 1) {
    const int nb = prm.b.nl + prm.b.nr;
    for (int64_t g = (int64_t)b * 32 + lane; g < prm.band_total; g += (int64_t)gsz * 32) {
      const int bi = (int)(g % nb);
      const int64_t o = g / nb;
      int64_t row, col0;
      int w;
      const double* coef;
      if (bi < prm.b.nl) {
        row = bi, col0 = 0, w = prm.b.wl, coef = prm.b.l + (int64_t)bi * prm.b.wl;
      } else {
        const int r = bi - prm.b.nl;
        row = prm.n - prm.b.nr + r, col0 = prm.n - prm.b.wr, w = prm.b.wr, coef = prm.b.r + (int64_t)r * prm.b.wr;
      }
      const T* xp = prm.x + o * prm.n + col0;
      T acc = 0;
      for (int c = 0; c < w; c += 4) {
        T cf[4], xv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int cu = min(c + u, w - 1);
          cf[u] = c + u < w ? (T)__ldg(coef + cu) : (T)0;
          xv[u] = __ldg(xp + cu);
        }
        // (a tap whose coefficient is exactly zero is skipped, as in band_element: 0 * inf must not become NaN)
#pragma unroll
        for (int u = 0; u < 4; ++u) acc = cf[u] != (T)0 ? fma(cf[u], xv[u], acc) : acc;
      }
      const T val = acc * prm.alpha;
      prm.out[o * prm.n + row] = val;
      if (prm.out2) prm.out2[o * prm.n + row] = val;
      if (prm.out3) prm.out3[o * prm.n + row] = val;
    }
    return;
  }

  // ---- consumer warps
  int stage = 0;
  uint32_t parity = 0;
  for (;;) {
    mbar_wait(bar_full + 8u * stage, parity);
    int strip, rb;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(strip), "=r"(rb) : "r"(item_xy + 8u * stage));
    if (strip < 0) break;
    const int tv = strip * G::TXV + lane;  // this lane's vector along the row
    const int i_first = tv * V;
    const uint32_t base = ring + (uint32_t)stage * G::STAGE_BYTES + (uint32_t)(lane * 16);
    const bool has_main = tv < prm.n_v && i_first + V > prm.r0 && i_first < prm.r1;
    const bool whole = i_first >= prm.r0 && i_first + V <= prm.r1;
#pragma unroll
    for (int rr = 0; rr < G::RW; ++rr) {
      const int lr = wrp * G::RW + rr;  // row inside the stage
      const int row = rb * G::TYR + lr;
      T w[G::NV * V];
#pragma unroll
      for (int u = 0; u < G::NV; ++u) {
        const VT r = lds_vec<T, V>(base + (uint32_t)(lr * G::BOX_X * (int)sizeof(T) + u * 16));
#pragma unroll
        for (int v = 0; v < V; ++v) w[u * V + v] = r.v[v];
      }
      // output v reads w[J0 + v + k], taps ascending
      constexpr int J0 = G::HLV * V + LO;
      VT res;
      if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float s0 = 0.f, s1 = 0.f;
#pragma unroll
          for (int q = 0; q < W; ++q) {
            const int j = J0 + 2 * h + q;
            if (q == 0) {
              s0 = prm.taps.c[0] * w[j], s1 = prm.taps.c[0] * w[j + 1];
            } else if (j % 2 == 0) {
              unsigned long long a = pack2(s0, s1);
              asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a) : "l"(pack2(prm.taps.c[q], prm.taps.c[q])), "l"(pack2(w[j], w[j + 1])));
              unpack2(a, s0, s1);
            } else {
              s0 = fma(prm.taps.c[q], w[j], s0), s1 = fma(prm.taps.c[q], w[j + 1], s1);
            }
          }
          res.v[2 * h] = s0, res.v[2 * h + 1] = s1;
        }
      } else {
#pragma unroll
        for (int v = 0; v < V; ++v) {
          T s = prm.taps.c[0] * w[J0 + v];
#pragma unroll
          for (int q = 1; q < W; ++q) s = fma(prm.taps.c[q], w[J0 + v + q], s);
          res.v[v] = s;
        }
      }
      if (has_main && row < prm.rows) {
        const int64_t at = (int64_t)row * prm.n + i_first;
        auto put = [&](T* dst) {
          T* orow = dst + at;
          if (whole) {
            st_vec<T, V>(orow, res);
          } else {
#pragma unroll
            for (int v = 0; v < V; ++v)
              if (i_first + v >= prm.r0 && i_first + v < prm.r1) orow[v] = res.v[v];
          }
        };
        put(prm.out);
        if (prm.out2) put(prm.out2);
        if (prm.out3) put(prm.out3);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty + 8u * stage);
    if (++stage == STAGES) stage = 0, parity ^= 1u;
  }
}

template <typename T, int LO, int HI>
static int launch(const fdx_plan_s* p, const T* x, T* out, int64_t outer, double alpha, cudaStream_t st, bool* band_done, T* out2, T* out3) {
  using G = Geo<T, LO, HI>;
  EncodeTiledFn enc = encode_fn();
  if (!enc) return FDX_EUNSUPPORTED;
  Params<T> prm;
  cuuint64_t dims[2] = {(cuuint64_t)p->n, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)p->n * sizeof(T)};
  cuuint32_t box[2] = {(cuuint32_t)G::BOX_X, (cuuint32_t)G::TYR};
  cuuint32_t estr[2] = {1, 1};
  if (enc(&prm.tmap, sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)x, dims, strides, box, estr,
          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return FDX_EUNSUPPORTED;
  const int64_t items = (int64_t)((p->n + G::TX - 1) / G::TX) * ((outer + G::TYR - 1) / G::TYR);
  if (items > 0x7fffffffLL) return FDX_EUNSUPPORTED;
  prm.x = x, prm.out = out, prm.out2 = out2, prm.out3 = out3, prm.rows = (int)outer, prm.n = p->n, prm.n_v = p->n / G::V, prm.r0 = p->nl, prm.r1 = p->n - p->nr;
  prm.strips = (p->n + G::TX - 1) / G::TX;
  prm.items = (int)items;
  prm.b = bands_of(p), prm.alpha = (T)alpha, prm.taps = scaled_taps<T>(p, alpha);
  prm.band_total = outer * (int64_t)(p->nl + p->nr);
  prm.ctr = claim_slot(st);
  if (!prm.ctr) return FDX_EUNSUPPORTED;
  auto kern = axis_rowtma_kernel<T, LO, HI>;
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
  const int per_sm = std::max(1, std::min(5, env_int("FDX_ROWTMA_CPS", (int)(220 * 1024 / (G::SMEM_BYTES + 1024)))));
  const int64_t grid = std::min<int64_t>(items, (int64_t)num_sms() * per_sm);
  kern<<<(unsigned)grid, G::NT, G::SMEM_BYTES, st>>>(prm);
  count_launch("axis_rowtma");
  *band_done = true;
  FDX_LAUNCH_CHECK();
  return FDX_OK;
}

template <typename T>
static int dispatch(const fdx_plan_s* p, const T* x, T* out, int64_t outer, double alpha, cudaStream_t st, bool* band_done, T* out2, T* out3) {
  const int lo = p->lo, hi = p->hi;
#define FDX_C(P_) \
  if (lo == -P_ && hi == P_) return launch<T, -P_, P_>(p, x, out, outer, alpha, st, band_done, out2, out3);
#define FDX_F(L_)                                                                        \
  if (lo == 0 && hi == L_) return launch<T, 0, L_>(p, x, out, outer, alpha, st, band_done, out2, out3); \
  if (lo == -L_ && hi == 0) return launch<T, -L_, 0>(p, x, out, outer, alpha, st, band_done, out2, out3);
  FDX_C(1) FDX_C(2) FDX_C(3) FDX_C(4) FDX_C(5) FDX_C(6)
  FDX_F(1) FDX_F(2) FDX_F(3) FDX_F(4) FDX_F(5) FDX_F(6) FDX_F(7) FDX_F(8) FDX_F(9) FDX_F(10) FDX_F(11)
#undef FDX_C
#undef FDX_F
  return FDX_EUNSUPPORTED;
}

}  // namespace rowtma

// Entry for fdx_axis.cu: FDX_EUNSUPPORTED = not this kernel's case (the caller goes on to axis_row).  Big aligned arrays
// only: below ~8 MB the L1-gather kernel's short CTAs fill the machine faster than a persistent ring starts up.
template <typename T>
int row_tma_apply(const fdx_plan_s* p, const T* x, T* out, int64_t outer, double alpha, cudaStream_t st, bool* band_done, T* out2, T* out3) {
  constexpr int V = 16 / (int)sizeof(T);
  if (env_int("FDX_ROWTMA_DISABLE", 0)) return FDX_EUNSUPPORTED;
  if (p->n % V != 0 || (((uintptr_t)x | (uintptr_t)out | (uintptr_t)out2 | (uintptr_t)out3) & 15) != 0) return FDX_EUNSUPPORTED;
  if (p->n - p->nl - p->nr <= 0 || outer <= 0 || outer > 0x7fffffffLL) return FDX_EUNSUPPORTED;
  const int64_t bytes = outer * (int64_t)p->n * (int64_t)sizeof(T);
  if (bytes < (int64_t)env_int("FDX_ROWTMA_MIN_MB", 8) * (1 << 20) || p->n < 32 * V) return FDX_EUNSUPPORTED;  // (at least one full strip)
  return rowtma::dispatch<T>(p, x, out, outer, alpha, st, band_done, out2, out3);
}
template int row_tma_apply<float>(const fdx_plan_s*, const float*, float*, int64_t, double, cudaStream_t, bool*, float*, float*);
template int row_tma_apply<double>(const fdx_plan_s*, const double*, double*, int64_t, double, cudaStream_t, bool*, double*, double*);

}  // namespace fdx
