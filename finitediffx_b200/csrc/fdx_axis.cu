// 1-D axis operator: out = alpha * D x (+ out), D described by an fdx_plan (main taps + bands).
//
// Replaces the reference's `difference` (finitediffx/_src/finite_diff.py:168-297) and its four
// row-region helpers (:45-165).  The reference builds each row region as a sum of shifted slices
// and concatenates them (>= 2 full read+write passes); here every element is read from HBM once
// and written once:
//
//   axis_stream  (axis is NOT the contiguous one; array seen as outer x n x inner)
//       one thread owns a 16-byte vector of `inner` and marches along the axis with a register
//       queue of W = hi-lo+1 vectors: 1 coalesced 128-bit load + 1 store per output vector.
//   axis_row     (axis IS the contiguous one)
//       one thread owns 16 bytes of outputs and gathers its window with aligned 128-bit loads;
//       the overlap between neighbouring threads is served by L1.
//   band rows    the few boundary rows with their own dense taps (one-sided stencils,
//       finite_diff.py:75-78,85-88 etc., or the wider bands of a transposed plan): they ride along as
//       the leading CTAs of the two kernels above (one launch per call); axis_band is the same code as a
//       kernel of its own behind the fallbacks.
//   *_scalar     shape/alignment fallbacks of the above (still CUDA, never the host).
#include <algorithm>
#include <type_traits>

#include "fdx_common.cuh"

namespace fdx {

// fdx_rowtma.cu / fdx_coltma.cu: the contiguous axis / a strided axis of big aligned arrays (FDX_EUNSUPPORTED: not their case)
template <typename T>
int col_tma_apply(const fdx_plan_s* p, const T* x, T* out, int64_t outer, int64_t inner, double alpha, cudaStream_t st, bool* band_done, T* out2 = nullptr,
                  T* out3 = nullptr);
template <typename T>
int row_tma_apply(const fdx_plan_s* p, const T* x, T* out, int64_t outer, double alpha, cudaStream_t st, bool* band_done, T* out2 = nullptr, T* out3 = nullptr);

// ---------------------------------------------------------------------------------- band rows
// The few boundary rows with their own dense taps.  One element per thread; `g` numbers the band elements of the
// whole array.  Runs as the leading CTAs of the main kernels (FusedBand: one launch per call, the band rows'
// latency hides behind the main rows) or as a kernel of its own behind the shape/alignment fallbacks.
template <typename T>
struct FusedBand {
  Bands b;
  T alpha;
  int64_t n, inner, outer;
  int blocks;  // leading CTAs that work on band elements (0: the caller launches axis_band_kernel)
  // runtime scale read ON THE DEVICE (the *_dev_alpha entry points: alpha is a traced value in the caller's graph and
  // must not be fetched by the host); nullptr: the scale is already folded into the taps / `alpha` on the host
  const double* alpha_dev;
};

template <typename T>
__device__ __forceinline__ T dev_scale(const double* alpha_dev) {
  return alpha_dev ? (T)__ldg(alpha_dev) : (T)1;
}

template <typename T, bool ACC>
__device__ __forceinline__ void band_element(const T* __restrict__ x, T* __restrict__ out, int64_t n, int64_t inner,
                                             int64_t outer, const Bands& b, T alpha, int64_t g, const double* alpha_dev = nullptr) {
  const int nb = b.nl + b.nr;
  if (g >= outer * nb * inner) return;
  const int64_t j = g % inner;
  const int64_t rest = g / inner;
  const int bi = (int)(rest % nb);
  const int64_t o = rest / nb;
  int64_t row, col0;
  int w;
  const double* coef;
  if (bi < b.nl) {
    row = bi, col0 = 0, w = b.wl, coef = b.l + (int64_t)bi * b.wl;
  } else {
    const int r = bi - b.nl;
    row = n - b.nr + r, col0 = n - b.wr, w = b.wr, coef = b.r + (int64_t)r * b.wr;
  }
  const T* xp = x + (o * n + col0) * inner + j;
  T acc = 0;
  for (int c = 0; c < w; ++c) {
    const T cf = (T)__ldg(coef + c);
    if (cf != (T)0) acc = fma(cf, __ldg(xp + (int64_t)c * inner), acc);
  }
  acc *= alpha * dev_scale<T>(alpha_dev);
  T* o_ptr = out + (o * n + row) * inner + j;
  *o_ptr = ACC ? *o_ptr + acc : acc;
}

// ------------------------------------------------------------------------------- axis_stream
template <typename T, int V, int W, bool ACC, bool DEVA>
__global__ void __launch_bounds__(128) axis_stream_kernel(const T* __restrict__ x, T* __restrict__ out, int64_t n,
                                                          int64_t inner_v, int lo, int r0, int r1, int rows_per_chunk,
                                                          int n_chunks, int64_t n_tiles, const __grid_constant__ Taps<T> taps,
                                                          const __grid_constant__ FusedBand<T> fb) {
  if ((int)blockIdx.x < fb.blocks) {
    band_element<T, ACC>(x, out, fb.n, fb.inner, fb.outer, fb.b, fb.alpha, (int64_t)blockIdx.x * blockDim.x + threadIdx.x, fb.alpha_dev);
    return;
  }
  int64_t b = (int64_t)blockIdx.x - fb.blocks;
  const int64_t tile = b % n_tiles;
  b /= n_tiles;
  const int c = (int)(b % n_chunks);
  const int64_t o = b / n_chunks;
  const int64_t jv = tile * blockDim.x + threadIdx.x;
  if (jv >= inner_v) return;
  const int i0 = r0 + c * rows_per_chunk;
  const int i1 = min(r1, i0 + rows_per_chunk);
  const int64_t stride = inner_v * V;
  const T* xp = x + (o * n + (int64_t)(i0 + lo)) * stride + jv * V;
  T* op = out + (o * n + (int64_t)i0) * stride + jv * V;

  using VT = Vec<T, V>;
  T scale = 1;
  if constexpr (DEVA) scale = dev_scale<T>(fb.alpha_dev);
  VT q[W];
#pragma unroll
  for (int k = 0; k < W - 1; ++k) {
    q[k] = ld_stream<T, V>(xp);
    xp += stride;
  }
  auto emit = [&](int s) {
    VT r;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      T acc = taps.c[0] * q[s % W].v[v];
#pragma unroll
      for (int k = 1; k < W; ++k) acc = fma(taps.c[k], q[(s + k) % W].v[v], acc);
      r.v[v] = DEVA ? acc * scale : acc;
    }
    if constexpr (ACC) {
      VT old = ld_cached<T, V>(op);
#pragma unroll
      for (int v = 0; v < V; ++v) r.v[v] += old.v[v];
    }
    st_vec<T, V>(op, r);
    op += stride;
  };
  int i = i0;
  for (; i + W <= i1; i += W) {
    // issue the W loads of this group up front: W independent 128-bit requests in flight
    VT nxt[W];
#pragma unroll
    for (int s = 0; s < W; ++s) nxt[s] = ld_stream<T, V>(xp + (int64_t)s * stride);
    xp += (int64_t)W * stride;
#pragma unroll
    for (int s = 0; s < W; ++s) {
      q[(W - 1 + s) % W] = nxt[s];
      emit(s);
    }
  }
#pragma unroll
  for (int s = 0; s < W - 1; ++s) {
    if (i + s < i1) {
      q[(W - 1 + s) % W] = ld_stream<T, V>(xp);
      xp += stride;
      emit(s);
    }
  }
}

// ---------------------------------------------------------------------------------- axis_row
__host__ __device__ constexpr int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

template <typename T, int V, int LO, int HI, bool ACC, bool DEVA, bool LOOP>
__global__ void __launch_bounds__(256) axis_row_kernel(const T* __restrict__ x, T* __restrict__ out, int64_t total_vecs,
                                                       int n_v, int r0, int r1, const __grid_constant__ Taps<T> taps,
                                                       const __grid_constant__ FusedBand<T> fb) {
  if ((int)blockIdx.x < fb.blocks) {
    band_element<T, ACC>(x, out, fb.n, fb.inner, fb.outer, fb.b, fb.alpha, (int64_t)blockIdx.x * blockDim.x + threadIdx.x, fb.alpha_dev);
    return;
  }
  // LOOP: a thread serves several output vectors, a whole grid of threads apart (long stencils: the index arithmetic
  // is paid once and several windows are in flight: the 11-13 tap cases of BASELINE config 2 gain 5 %); short stencils
  // keep one vector per thread and straight-line code (the loop form measured 9-17 % slower at the median of the 84
  // cases of an axis, with the same instructions inside: the early exits matter)
  T scale = 1;
  if constexpr (DEVA) scale = dev_scale<T>(fb.alpha_dev);
  auto one = [&](int64_t g) {
    const int64_t row = g / n_v;
    const int t = (int)(g - row * n_v);
    const int i_first = t * V;
    if (i_first + V <= r0 || i_first >= r1) return;  // band-only vector
    const T* xr = x + row * ((int64_t)n_v * V);
    constexpr int A = floor_div(LO, V), B = floor_div(V - 1 + HI, V), NV = B - A + 1, W = HI - LO + 1;
    T w[NV * V];
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      const int vi = min(max(t + A + u, 0), n_v - 1);
      Vec<T, V> r = ld_cached<T, V>(xr + (int64_t)vi * V);
#pragma unroll
      for (int v = 0; v < V; ++v) w[u * V + v] = r.v[v];
    }
    Vec<T, V> res;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      T acc = taps.c[0] * w[v + LO - A * V];
#pragma unroll
      for (int k = 1; k < W; ++k) acc = fma(taps.c[k], w[v + LO + k - A * V], acc);
      res.v[v] = DEVA ? acc * scale : acc;
    }
    T* orow = out + row * ((int64_t)n_v * V) + i_first;
    if (i_first >= r0 && i_first + V <= r1) {
      if constexpr (ACC) {
        Vec<T, V> old = ld_cached<T, V>(orow);
#pragma unroll
        for (int v = 0; v < V; ++v) res.v[v] += old.v[v];
      }
      st_vec<T, V>(orow, res);
    } else {
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const int i = i_first + v;
        if (i >= r0 && i < r1) orow[v] = ACC ? orow[v] + res.v[v] : res.v[v];
      }
    }
  };
  const int64_t g0 = ((int64_t)blockIdx.x - fb.blocks) * blockDim.x + threadIdx.x;
  if constexpr (LOOP) {
    const int64_t nthreads = (int64_t)(gridDim.x - fb.blocks) * blockDim.x;
#pragma unroll 1
    for (int64_t g = g0; g < total_vecs; g += nthreads) one(g);
  } else {
    if (g0 < total_vecs) one(g0);
  }
}

// contiguous axis, any length/alignment: one thread per main-row output, taps through L1
template <typename T, bool ACC>
__global__ void __launch_bounds__(256) axis_row_scalar_kernel(const T* __restrict__ x, T* __restrict__ out, int64_t outer,
                                                              int n, int lo, int taps_n, int r0, int r1, const __grid_constant__ Taps<T> taps,
                                                              const double* alpha_dev) {
  const int m = r1 - r0;
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= outer * m) return;
  const int64_t row = g / m;
  const int i = r0 + (int)(g - row * m);
  const T* xr = x + row * n + i + lo;
  T acc = taps.c[0] * __ldg(xr);
  for (int k = 1; k < taps_n; ++k) acc = fma(taps.c[k], __ldg(xr + k), acc);
  acc *= dev_scale<T>(alpha_dev);
  T* o = out + row * n + i;
  *o = ACC ? *o + acc : acc;
}

// ---------------------------------------------------------------------------------- axis_band
template <typename T, bool ACC>
__global__ void __launch_bounds__(256) axis_band_kernel(const T* __restrict__ x, T* __restrict__ out, int64_t n,
                                                        int64_t inner, int64_t outer, Bands b, T alpha, const double* alpha_dev) {
  band_element<T, ACC>(x, out, n, inner, outer, b, alpha, (int64_t)blockIdx.x * blockDim.x + threadIdx.x, alpha_dev);
}

// The band rows as leading CTAs of a main kernel with `threads` threads per CTA (blocks = 0 if there are none or
// too many; the caller then launches axis_band_kernel).
template <typename T>
static FusedBand<T> fused_band(const fdx_plan_s* p, int64_t outer, int64_t inner, double alpha, int threads, const double* alpha_dev) {
  FusedBand<T> fb;
  fb.b = bands_of(p), fb.alpha = (T)alpha, fb.n = p->n, fb.inner = inner, fb.outer = outer, fb.blocks = 0, fb.alpha_dev = alpha_dev;
  const int64_t total = outer * (int64_t)(p->nl + p->nr) * inner;
  const int64_t blocks = (total + threads - 1) / threads;
  if (blocks > 0 && blocks <= (1 << 20)) fb.blocks = (int)blocks;
  return fb;
}

// ----------------------------------------------------------------------------------- dispatch
template <typename T, int V, int W, bool ACC>
static int launch_stream(const fdx_plan_s* p, const T* x, T* out, int64_t outer, int64_t inner, double alpha,
                         cudaStream_t st, bool* band_done, const double* alpha_dev) {
  const int r0 = p->nl, r1 = p->n - p->nr;
  const int m = r1 - r0;
  if (m <= 0) return FDX_OK;
  const int64_t inner_v = inner / V;
  const int64_t n_tiles = (inner_v + 127) / 128;
  // Rows per chunk: a chunk re-reads W - 1 halo rows, so long stencils want long chunks (B200 sweep of BASELINE
  // config 2, tools/diff_probe.py: 7+ taps gain 4-9 % at ~220 rows, 3-5 taps are best at 24-60), but never fewer than
  // ~4 CTAs per SM and never shorter than 8 groups.  FDX_AXIS_CPS = CTAs per SM instead (sweeps).
  const int min_rows = 8 * W;
  int rows;
  if (const int cps = env_int("FDX_AXIS_CPS", 0)) {
    const int64_t want = (int64_t)num_sms() * std::max(1, cps);
    const int64_t chunks = (want + n_tiles * outer - 1) / (n_tiles * outer);
    rows = (int)((m + chunks - 1) / chunks);
  } else {
    const int64_t chunks_min = (4LL * num_sms() + n_tiles * outer - 1) / (n_tiles * outer);
    const int rows_max = (int)((m + chunks_min - 1) / chunks_min);
    rows = std::min((W <= 5 ? 8 : 32) * W, rows_max);
  }
  if (rows < min_rows) rows = min_rows;
  rows = ((rows + W - 1) / W) * W;
  if (rows > m) rows = m;
  const int n_chunks = (m + rows - 1) / rows;
  const FusedBand<T> fb = fused_band<T>(p, outer, inner, alpha, 128, alpha_dev);
  const int64_t blocks = n_tiles * n_chunks * outer + fb.blocks;
  if (blocks > 0x7fffffffLL) return FDX_EUNSUPPORTED;
  if (alpha_dev)
    axis_stream_kernel<T, V, W, ACC, true><<<(unsigned)blocks, 128, 0, st>>>(x, out, p->n, inner_v, p->lo, r0, r1, rows, n_chunks,
                                                                             n_tiles, scaled_taps<T>(p, alpha), fb);
  else
    axis_stream_kernel<T, V, W, ACC, false><<<(unsigned)blocks, 128, 0, st>>>(x, out, p->n, inner_v, p->lo, r0, r1, rows, n_chunks,
                                                                              n_tiles, scaled_taps<T>(p, alpha), fb);
  count_launch("axis_stream");
  *band_done = fb.blocks > 0;
  return FDX_OK;
}

template <typename T, int V, bool ACC>
static int dispatch_stream(const fdx_plan_s* p, const T* x, T* out, int64_t outer, int64_t inner, double alpha,
                           cudaStream_t st, bool* band_done, const double* alpha_dev) {
  switch (p->hi - p->lo + 1) {
#define FDX_CASE(W_) \
  case W_:           \
    return launch_stream<T, V, W_, ACC>(p, x, out, outer, inner, alpha, st, band_done, alpha_dev);
    FDX_CASE(1) FDX_CASE(2) FDX_CASE(3) FDX_CASE(4) FDX_CASE(5) FDX_CASE(6) FDX_CASE(7) FDX_CASE(8) FDX_CASE(9) FDX_CASE(10)
    FDX_CASE(11) FDX_CASE(12) FDX_CASE(13) FDX_CASE(14) FDX_CASE(15) FDX_CASE(16)  // up to FDX_MAX_TAPS
#undef FDX_CASE
    default:
      return FDX_EUNSUPPORTED;
  }
}

template <typename T, int V, int LO, int HI, bool ACC>
static int launch_row(const fdx_plan_s* p, const T* x, T* out, int64_t outer, double alpha, cudaStream_t st, bool* band_done,
                      const double* alpha_dev) {
  const int n_v = p->n / V;
  const int64_t total = outer * n_v;
  const FusedBand<T> fb = fused_band<T>(p, outer, 1, alpha, 256, alpha_dev);
  // vectors per thread (FDX_ROW_VPT; default 4 for stencils of 9+ taps, 1 below)
  constexpr bool kLong = HI - LO + 1 >= 9;
  int vpt = env_int("FDX_ROW_VPT", kLong ? 4 : 1);
  if (vpt < 1) vpt = 1;
  while (vpt > 1 && (total + 256LL * vpt - 1) / (256LL * vpt) < 4LL * num_sms()) vpt /= 2;  // small arrays: keep the machine full
  const int64_t blocks = (total + 256LL * vpt - 1) / (256LL * vpt) + fb.blocks;
  if (blocks > 0x7fffffffLL) return FDX_EUNSUPPORTED;
  auto go = [&](auto deva, auto loop) {
    axis_row_kernel<T, V, LO, HI, ACC, decltype(deva)::value, decltype(loop)::value><<<(unsigned)blocks, 256, 0, st>>>(
        x, out, total, n_v, p->nl, p->n - p->nr, scaled_taps<T>(p, alpha), fb);
  };
  using Tr = std::true_type;
  using Fa = std::false_type;
  if (alpha_dev)
    go(Tr{}, Tr{});  // (the traced-scale path always takes the loop form: one instantiation)
  else if (vpt > 1)
    go(Fa{}, Tr{});
  else
    go(Fa{}, Fa{});
  count_launch("axis_row");
  *band_done = fb.blocks > 0;
  return FDX_OK;
}

template <typename T, int V, bool ACC>
static int dispatch_row(const fdx_plan_s* p, const T* x, T* out, int64_t outer, double alpha, cudaStream_t st, bool* band_done,
                        const double* alpha_dev) {
  const int lo = p->lo, hi = p->hi;
#define FDX_C(P_) \
  if (lo == -P_ && hi == P_) return launch_row<T, V, -P_, P_, ACC>(p, x, out, outer, alpha, st, band_done, alpha_dev);
#define FDX_F(L_)                                                                              \
  if (lo == 0 && hi == L_) return launch_row<T, V, 0, L_, ACC>(p, x, out, outer, alpha, st, band_done, alpha_dev); \
  if (lo == -L_ && hi == 0) return launch_row<T, V, -L_, 0, ACC>(p, x, out, outer, alpha, st, band_done, alpha_dev);
  FDX_C(1) FDX_C(2) FDX_C(3) FDX_C(4) FDX_C(5) FDX_C(6)
  FDX_F(1) FDX_F(2) FDX_F(3) FDX_F(4) FDX_F(5) FDX_F(6) FDX_F(7) FDX_F(8) FDX_F(9) FDX_F(10) FDX_F(11)
#undef FDX_C
#undef FDX_F
  return FDX_EUNSUPPORTED;
}

template <typename T, bool ACC>
static int apply_impl(const fdx_plan_s* p, const T* x, T* out, int64_t outer, int64_t inner, double alpha,
                      cudaStream_t st, const double* alpha_dev) {
  constexpr int V = 16 / sizeof(T);
  const int m = p->n - p->nl - p->nr;
  const bool aligned16 = (((uintptr_t)x | (uintptr_t)out) & 15) == 0;
  int rc = FDX_OK;
  bool band_done = false;  // the band rows rode along with the main kernel
  if (m > 0) {
    if (inner > 1) {
      rc = FDX_EUNSUPPORTED;
      // big arrays: the persistent TMA-staged kernel for strided axes (fdx_coltma.cu); it writes, so not for the accumulating form
      if constexpr (!ACC) {
        if (!alpha_dev) rc = col_tma_apply<T>(p, x, out, outer, inner, alpha, st, &band_done);
      }
      if (rc != FDX_EUNSUPPORTED) {
      } else if (aligned16 && inner % V == 0)
        rc = dispatch_stream<T, V, ACC>(p, x, out, outer, inner, alpha, st, &band_done, alpha_dev);
      else
        rc = dispatch_stream<T, 1, ACC>(p, x, out, outer, inner, alpha, st, &band_done, alpha_dev);
    } else {
      rc = FDX_EUNSUPPORTED;
      // big arrays: the persistent TMA-staged row kernel (fdx_rowtma.cu); it writes, so not for the accumulating form
      if constexpr (!ACC) {
        if (!alpha_dev) rc = row_tma_apply<T>(p, x, out, outer, alpha, st, &band_done);
      }
      if (rc == FDX_EUNSUPPORTED && aligned16 && p->n % V == 0)
        rc = dispatch_row<T, V, ACC>(p, x, out, outer, alpha, st, &band_done, alpha_dev);
      if (rc == FDX_EUNSUPPORTED) {
        const int64_t total = outer * m;
        const int64_t blocks = (total + 255) / 256;
        if (blocks > 0x7fffffffLL) return FDX_EUNSUPPORTED;
        axis_row_scalar_kernel<T, ACC><<<(unsigned)blocks, 256, 0, st>>>(x, out, outer, p->n, p->lo, p->hi - p->lo + 1, p->nl,
                                                                         p->n - p->nr, scaled_taps<T>(p, alpha), alpha_dev);
        count_launch("axis_row_scalar");
        rc = FDX_OK;
      }
    }
    if (rc != FDX_OK) return rc;
    FDX_LAUNCH_CHECK();
  }
  const int nb = p->nl + p->nr;
  if (nb > 0 && !band_done) {
    const int64_t total = outer * nb * inner;
    const int64_t blocks = (total + 255) / 256;
    if (blocks > 0x7fffffffLL) return FDX_EUNSUPPORTED;
    axis_band_kernel<T, ACC><<<(unsigned)blocks, 256, 0, st>>>(x, out, p->n, inner, outer, bands_of(p), (T)alpha, alpha_dev);
    count_launch("axis_band");
    FDX_LAUNCH_CHECK();
  }
  return FDX_OK;
}

template <typename T>
static int apply_checked(fdx_plan_t plan, const T* x, T* out, int64_t outer, int64_t inner, double alpha, int accumulate,
                         void* stream, const double* alpha_dev = nullptr) {
  if (!plan || !x || !out || outer < 0 || inner < 0) return FDX_EINVAL;
  if ((((uintptr_t)x | (uintptr_t)out) & (sizeof(T) - 1)) != 0) return FDX_EALIGN;
  if (outer == 0 || inner == 0 || plan->n == 0) return FDX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  return accumulate ? apply_impl<T, true>(plan, x, out, outer, inner, alpha, st, alpha_dev)
                    : apply_impl<T, false>(plan, x, out, outer, inner, alpha, st, alpha_dev);
}

// out = out2 = out3 = alpha * D x in one pass where the TMA kernels take the call (their stores are simply repeated);
// otherwise the ordinary pass into `out` and device-to-device copies behind it on the same stream.
template <typename T>
static int apply_dup(fdx_plan_t plan, const T* x, T* out, T* out2, T* out3, int64_t outer, int64_t inner, double alpha, void* stream) {
  if (!out2 && out3) out2 = out3, out3 = nullptr;
  if (!out2) return apply_checked<T>(plan, x, out, outer, inner, alpha, 0, stream);
  if (!plan || !x || !out || outer < 0 || inner < 0) return FDX_EINVAL;
  if ((((uintptr_t)x | (uintptr_t)out | (uintptr_t)out2 | (uintptr_t)out3) & (sizeof(T) - 1)) != 0) return FDX_EALIGN;
  if (outer == 0 || inner == 0 || plan->n == 0) return FDX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  bool band_done = false;
  int rc = FDX_EUNSUPPORTED;
  if (plan->n - plan->nl - plan->nr > 0)
    rc = inner > 1 ? col_tma_apply<T>(plan, x, out, outer, inner, alpha, st, &band_done, out2, out3)
                   : row_tma_apply<T>(plan, x, out, outer, alpha, st, &band_done, out2, out3);
  if (rc != FDX_EUNSUPPORTED) return rc;
  rc = apply_checked<T>(plan, x, out, outer, inner, alpha, 0, stream);
  if (rc != FDX_OK) return rc;
  const size_t bytes = (size_t)outer * (size_t)plan->n * (size_t)inner * sizeof(T);
  FDX_CUDA_TRY(cudaMemcpyAsync(out2, out, bytes, cudaMemcpyDeviceToDevice, st));
  if (out3) FDX_CUDA_TRY(cudaMemcpyAsync(out3, out, bytes, cudaMemcpyDeviceToDevice, st));
  return FDX_OK;
}

}  // namespace fdx

extern "C" int fdx_axis_apply_dup_f32(fdx_plan_t plan, const float* x, float* out, float* out2, float* out3, int64_t outer, int64_t inner,
                                      double alpha, void* stream) {
  return fdx::apply_dup<float>(plan, x, out, out2, out3, outer, inner, alpha, stream);
}

extern "C" int fdx_axis_apply_dup_f64(fdx_plan_t plan, const double* x, double* out, double* out2, double* out3, int64_t outer, int64_t inner,
                                      double alpha, void* stream) {
  return fdx::apply_dup<double>(plan, x, out, out2, out3, outer, inner, alpha, stream);
}

extern "C" int fdx_axis_apply_f32(fdx_plan_t plan, const float* x, float* out, int64_t outer, int64_t inner, double alpha,
                                  int accumulate, void* stream) {
  return fdx::apply_checked<float>(plan, x, out, outer, inner, alpha, accumulate, stream);
}

extern "C" int fdx_axis_apply_f64(fdx_plan_t plan, const double* x, double* out, int64_t outer, int64_t inner,
                                  double alpha, int accumulate, void* stream) {
  return fdx::apply_checked<double>(plan, x, out, outer, inner, alpha, accumulate, stream);
}

// ---- alpha read on the device (XLA custom calls: `step_size` is a traced value, finite_diff.py:168): the host never
// touches it, so the call needs no synchronisation and is safe under stream capture.  out = sign * (*alpha_dev) * D x.
extern "C" int fdx_axis_apply_dev_alpha_f32(fdx_plan_t plan, const float* x, float* out, int64_t outer, int64_t inner,
                                            const double* alpha_dev, double sign, int accumulate, void* stream) {
  if (!alpha_dev) return FDX_EINVAL;
  return fdx::apply_checked<float>(plan, x, out, outer, inner, sign, accumulate, stream, alpha_dev);
}

extern "C" int fdx_axis_apply_dev_alpha_f64(fdx_plan_t plan, const double* x, double* out, int64_t outer, int64_t inner,
                                            const double* alpha_dev, double sign, int accumulate, void* stream) {
  if (!alpha_dev) return FDX_EINVAL;
  return fdx::apply_checked<double>(plan, x, out, outer, inner, sign, accumulate, stream, alpha_dev);
}
