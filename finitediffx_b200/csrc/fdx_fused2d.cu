// Fused single-pass 2-D operators: laplacian, divergence (and, with swapped inputs and a negated scale, the 2-D
// curl), gradient  (include/fdx_b200.h, fdx_*2d_*).
//
// The reference computes them as two `difference` passes plus an add / stack
// (finitediffx/_src/finite_diff.py:338-351, 443-453, 565-575, 586-606); most of its own tests run on 2-D grids
// (tests/test_finite_diff.py:161-307).  Here one kernel reads every input once and writes every output once.
//
// Design -- the two axis kernels of fdx_axis.cu welded together.  The grid is C-contiguous (n0, n1):
//   * a thread owns one 16-byte vector of columns and marches along axis 0 over a chunk of rows with a register
//     window of 2P + 1 row vectors (the axis-0 stencil in gather form: one coalesced 128-bit load per new row);
//   * the axis-1 stencil of a row needs the neighbouring vectors of the same row: they are the vectors the
//     neighbouring lanes load as their own, so the extra loads hit L1 (cached loads throughout);
//   * the main rows and columns -- where every tap of both stencils lies inside the grid -- take this path; the
//     frame (boundary rows of axis 0, boundary vectors of axis 1) is evaluated point by point by the leading CTAs of
//     the same launch with the plans' dense band rows (one-sided stencils, finite_diff.py:75-88, or the wide bands of
//     a transposed plan), each term from its band or its main stencil as the point requires.
//   Both stencils run at the larger radius P; an axis with a shorter or one-sided stencil has zero taps (the main
//   region is shrunk so that zero taps still read inside the array).
// HBM bound: 2s / 3s / 3s bytes per point (laplacian / divergence / gradient), like the 1-D kernels.
#include <algorithm>

#include "fdx_common.cuh"

namespace fdx {
namespace {

enum Op2 { LAP2 = 0, DIV2 = 1, GRAD2 = 2 };

template <typename T>
struct Axis2 {
  Bands b;        // dense band rows (device, float64)
  int n, lo, hi;  // true extent of the main stencil
  T alpha;
  T c[FDX_MAX_TAPS];  // kernel taps at offsets -P .. P, scaled by alpha (zero outside lo .. hi)
  T m[FDX_MAX_TAPS];  // the plan's own main taps (offsets lo .. hi), unscaled: the frame path
};

template <typename T>
struct Params2 {
  const T* in[2];
  T* out[2];
  Axis2<T> ax[2];
  int r0, r1;    // main rows [r0, r1)
  int v0, v1;    // main vectors [v0, v1) of a row
  int n1v;       // vectors per row
  int rows_per_chunk, n_chunks, n_tiles;
  int frame_blocks;      // leading CTAs that evaluate frame points
  int64_t frame_points;  // (r0 + n0 - r1) * n1 + (r1 - r0) * (v0 + n1v - v1) * V
};

// one term at one point, from the band or the main stencil of its axis (the frame path)
template <typename T>
__device__ __forceinline__ T axis_dot(const Axis2<T>& a, const T* x, int64_t base, int64_t stride, int i) {
  T acc = 0;
  if (i < a.b.nl) {
    const double* cf = a.b.l + (int64_t)i * a.b.wl;
    for (int c = 0; c < a.b.wl; ++c) {
      const T k = (T)__ldg(cf + c);
      if (k != (T)0) acc = fma(k, __ldg(x + base + (int64_t)c * stride), acc);
    }
  } else if (i >= a.n - a.b.nr) {
    const double* cf = a.b.r + (int64_t)(i - (a.n - a.b.nr)) * a.b.wr;
    const int64_t c0 = a.n - a.b.wr;
    for (int c = 0; c < a.b.wr; ++c) {
      const T k = (T)__ldg(cf + c);
      if (k != (T)0) acc = fma(k, __ldg(x + base + (c0 + c) * stride), acc);
    }
  } else {
    for (int k = 0; k <= a.hi - a.lo; ++k) acc = fma(a.m[k], __ldg(x + base + (int64_t)(i + a.lo + k) * stride), acc);
  }
  return acc * a.alpha;
}

template <Op2 OP, typename T, int V, int P>
__global__ void __launch_bounds__(128) fused2d_kernel(const __grid_constant__ Params2<T> prm) {
  constexpr int W = 2 * P + 1;
  constexpr int HXV = (P + V - 1) / V;
  using VT = Vec<T, V>;
  const int n1 = prm.n1v * V;
  if ((int)blockIdx.x < prm.frame_blocks) {
    // ---- frame: boundary rows (all columns), then the boundary vectors of the main rows
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= prm.frame_points) return;
    const int top = prm.r0, bot = prm.ax[0].n - prm.r1;
    int i0, i1;
    if (g < (int64_t)(top + bot) * n1) {
      const int r = (int)(g / n1);
      i1 = (int)(g - (int64_t)r * n1);
      i0 = r < top ? r : prm.r1 + (r - top);
    } else {
      const int64_t h = g - (int64_t)(top + bot) * n1;
      const int wl = prm.v0 * V, wr = (prm.n1v - prm.v1) * V;
      const int r = (int)(h / (wl + wr));
      const int c = (int)(h - (int64_t)r * (wl + wr));
      i0 = prm.r0 + r;
      i1 = c < wl ? c : prm.v1 * V + (c - wl);
    }
    const int64_t at = (int64_t)i0 * n1 + i1;
    const int ci0 = OP == DIV2 ? 0 : 0, ci1 = OP == DIV2 ? 1 : 0;
    const T d0 = axis_dot<T>(prm.ax[0], prm.in[ci0], i1, n1, i0);               // along axis 0: column i1
    const T d1 = axis_dot<T>(prm.ax[1], prm.in[ci1], (int64_t)i0 * n1, 1, i1);  // along axis 1: row i0
    if constexpr (OP == GRAD2) {
      prm.out[0][at] = d0;
      prm.out[1][at] = d1;
    } else {
      prm.out[0][at] = d0 + d1;
    }
    return;
  }
  // ---- main region
  int b = (int)blockIdx.x - prm.frame_blocks;
  const int tile = b % prm.n_tiles;
  const int chunk = b / prm.n_tiles;
  const int jv = prm.v0 + tile * (int)blockDim.x + (int)threadIdx.x;
  if (jv >= prm.v1) return;
  const int i_begin = prm.r0 + chunk * prm.rows_per_chunk;
  const int i_end = min(prm.r1, i_begin + prm.rows_per_chunk);
  const int64_t stride = n1;
  const T* xa = prm.in[0];                       // field under the axis-0 stencil
  const T* xb = prm.in[OP == DIV2 ? 1 : 0];      // field under the axis-1 stencil
  const Axis2<T>& A0 = prm.ax[0];
  const Axis2<T>& A1 = prm.ax[1];
  const int64_t col = (int64_t)jv * V;
  // register window of the axis-0 stencil: rows i - P .. i + P of this thread's vector
  VT q[W];
#pragma unroll
  for (int k = 0; k < W - 1; ++k) q[k] = ld_cached<T, V>(xa + (int64_t)(i_begin - P + k) * stride + col);
#pragma unroll 1
  for (int i = i_begin; i < i_end; i += W) {
    // the W new rows of this group are requested up front: W independent 128-bit loads in flight per thread (a row
    // loaded where it is first used costs one DRAM round trip per row: measured 2x slower at 4 CTAs per SM)
    VT nxt[W];
#pragma unroll
    for (int s = 0; s < W; ++s)
      if (i + s < i_end) nxt[s] = ld_cached<T, V>(xa + (int64_t)(i + s + P) * stride + col);
#pragma unroll
    for (int s = 0; s < W; ++s) {
      if (i + s < i_end) {
        const int row = i + s;
        q[(W - 1 + s) % W] = nxt[s];
        // axis 0: taps over the window (slot (s + k) % W holds row - P + k)
        VT d0;
#pragma unroll
        for (int v = 0; v < V; ++v) {
          T acc = A0.c[0] * q[s % W].v[v];
#pragma unroll
          for (int k = 1; k < W; ++k) acc = fma(A0.c[k], q[(s + k) % W].v[v], acc);
          d0.v[v] = acc;
        }
        // axis 1: this row's vector and its neighbours (L1 hits: the neighbouring lanes' own vectors)
        T w[(2 * HXV + 1) * V];
        const T* rowp = xb + (int64_t)row * stride + col;
#pragma unroll
        for (int u = 0; u < 2 * HXV + 1; ++u) {
          VT r;
          if (OP != DIV2 && u == HXV)
            r = q[(s + P) % W];  // the centre vector is already in the window
          else
            r = ld_cached<T, V>(rowp + (u - HXV) * V);
#pragma unroll
          for (int v = 0; v < V; ++v) w[u * V + v] = r.v[v];
        }
        VT d1;
#pragma unroll
        for (int v = 0; v < V; ++v) {
          T acc = A1.c[0] * w[HXV * V + v - P];
#pragma unroll
          for (int k = 1; k < W; ++k) acc = fma(A1.c[k], w[HXV * V + v - P + k], acc);
          d1.v[v] = acc;
        }
        const int64_t at = (int64_t)row * stride + col;
        if constexpr (OP == GRAD2) {
          st_vec<T, V>(prm.out[0] + at, d0);
          st_vec<T, V>(prm.out[1] + at, d1);
        } else {
#pragma unroll
          for (int v = 0; v < V; ++v) d0.v[v] += d1.v[v];
          st_vec<T, V>(prm.out[0] + at, d0);
        }
      }
    }
  }
}

template <Op2 OP, typename T, int P>
int launch2(const fdx_plan_s* const plans[2], const T* const in[2], T* const out[2], const double alpha[2], cudaStream_t st) {
  constexpr int V = 16 / (int)sizeof(T);
  constexpr int W = 2 * P + 1;
  constexpr int HXV = (P + V - 1) / V;
  Params2<T> prm;
  const int n0 = plans[0]->n, n1 = plans[1]->n;
  for (int f = 0; f < 2; ++f) prm.in[f] = in[f], prm.out[f] = out[f];
  for (int k = 0; k < 2; ++k) {
    const fdx_plan_s* p = plans[k];
    Axis2<T>& a = prm.ax[k];
    a.b = bands_of(p), a.n = p->n, a.lo = p->lo, a.hi = p->hi, a.alpha = (T)alpha[k];
    for (int t = 0; t < FDX_MAX_TAPS; ++t) {
      const int o = t - P;
      a.c[t] = (t < W && o >= p->lo && o <= p->hi) ? (T)(p->main_taps[o - p->lo] * alpha[k]) : (T)0;
      a.m[t] = t <= p->hi - p->lo ? (T)p->main_taps[t] : (T)0;
    }
  }
  // main region: every tap of the radius-P stencils (zero taps included) reads inside the array, and no band row
  prm.r0 = std::max(plans[0]->nl, P), prm.r1 = n0 - std::max(plans[0]->nr, P);
  prm.n1v = n1 / V;
  prm.v0 = std::max((plans[1]->nl + V - 1) / V, HXV), prm.v1 = prm.n1v - std::max((plans[1]->nr + V - 1) / V, HXV);
  if (prm.r1 < prm.r0) prm.r1 = prm.r0;
  if (prm.v1 < prm.v0) prm.v1 = prm.v0;
  const int rows = prm.r1 - prm.r0, vecs = prm.v1 - prm.v0;
  prm.frame_points = (int64_t)(n0 - rows) * n1 + (int64_t)rows * (prm.n1v - vecs) * V;
  const int64_t fb = (prm.frame_points + 127) / 128;
  prm.n_tiles = std::max(1, (vecs + 127) / 128);
  // rows per chunk as in axis_stream: long chunks for long stencils (W - 1 halo rows are re-read per chunk), but
  // at least ~4 CTAs per SM
  const int64_t chunks_min = std::max<int64_t>(1, (4LL * num_sms() + prm.n_tiles - 1) / prm.n_tiles);
  int rpc = std::min((W <= 5 ? 8 : 32) * W, (int)((rows + chunks_min - 1) / chunks_min));
  rpc = std::max(rpc, std::min(rows, 8 * W));
  rpc = std::max(1, ((rpc + W - 1) / W) * W);
  prm.rows_per_chunk = rpc;
  prm.n_chunks = rows > 0 && vecs > 0 ? (rows + rpc - 1) / rpc : 0;
  const int64_t blocks = fb + (int64_t)prm.n_tiles * prm.n_chunks;
  if (fb > 0x3fffffffLL || blocks > 0x7fffffffLL) return FDX_EUNSUPPORTED;
  prm.frame_blocks = (int)fb;
  if (blocks == 0) return FDX_OK;
  fused2d_kernel<OP, T, V, P><<<(unsigned)blocks, 128, 0, st>>>(prm);
  count_launch(OP == LAP2 ? "fused_laplacian2d" : OP == DIV2 ? "fused_divergence2d" : "fused_gradient2d");
  FDX_LAUNCH_CHECK();
  return FDX_OK;
}

template <Op2 OP, typename T>
int run2(const fdx_plan_t plans_in[2], const T* const in[2], T* const out[2], const double alpha[2], void* stream) {
  constexpr int V = 16 / (int)sizeof(T);
  if (!plans_in || !plans_in[0] || !plans_in[1] || !alpha) return FDX_EINVAL;
  if (env_int("FDX_FUSED2D_DISABLE", 0)) return FDX_EUNSUPPORTED;  // (tests: the composition of axis kernels instead)
  const int nin = OP == DIV2 ? 2 : 1, nout = OP == GRAD2 ? 2 : 1;
  for (int f = 0; f < nin; ++f)
    if (!in[f]) return FDX_EINVAL;
  for (int f = 0; f < nout; ++f)
    if (!out[f]) return FDX_EINVAL;
  const fdx_plan_s* plans[2] = {plans_in[0], plans_in[1]};
  if (plans[0]->n == 0 || plans[1]->n == 0) return FDX_OK;
  int P = 0;
  for (int k = 0; k < 2; ++k) {
    const fdx_plan_s* p = plans[k];
    if (p->lo > 0 || p->hi < 0) return FDX_EUNSUPPORTED;
    if (p->nl + p->nr > p->n) return FDX_EUNSUPPORTED;  // dense (tiny-n) plans: composition
    if (p->nl < -p->lo || p->nr < p->hi) return FDX_EUNSUPPORTED;
    P = std::max(P, std::max(-p->lo, p->hi));
  }
  if (P < 1 || P > 6) return FDX_EUNSUPPORTED;
  if (plans[1]->n % V != 0) return FDX_EUNSUPPORTED;
  if ((int64_t)plans[0]->n * plans[1]->n > 0x7fffffffLL * 64) return FDX_EUNSUPPORTED;
  for (int f = 0; f < nin; ++f)
    if (((uintptr_t)in[f] & 15) != 0) return FDX_EUNSUPPORTED;
  for (int f = 0; f < nout; ++f)
    if (((uintptr_t)out[f] & 15) != 0) return FDX_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const T* in2[2] = {in[0], nin > 1 ? in[1] : nullptr};
  T* out2[2] = {out[0], nout > 1 ? out[1] : nullptr};
  switch (P) {
    case 1: return launch2<OP, T, 1>(plans, in2, out2, alpha, st);
    case 2: return launch2<OP, T, 2>(plans, in2, out2, alpha, st);
    case 3: return launch2<OP, T, 3>(plans, in2, out2, alpha, st);
    case 4: return launch2<OP, T, 4>(plans, in2, out2, alpha, st);
    case 5: return launch2<OP, T, 5>(plans, in2, out2, alpha, st);
    default: return launch2<OP, T, 6>(plans, in2, out2, alpha, st);
  }
}

}  // namespace
}  // namespace fdx

#define FDX_DEF2(SUFFIX, T)                                                                                                  \
  extern "C" int fdx_laplacian2d_##SUFFIX(const fdx_plan_t plans[2], const T* x, T* out, const double alpha[2], void* stream) { \
    const T* in[2] = {x, nullptr};                                                                                           \
    T* o[2] = {out, nullptr};                                                                                                \
    return fdx::run2<fdx::LAP2, T>(plans, in, o, alpha, stream);                                                             \
  }                                                                                                                          \
  extern "C" int fdx_divergence2d_##SUFFIX(const fdx_plan_t plans[2], const T* const x[2], T* out, const double alpha[2],    \
                                           void* stream) {                                                                   \
    if (!x) return FDX_EINVAL;                                                                                               \
    const T* in[2] = {x[0], x[1]};                                                                                           \
    T* o[2] = {out, nullptr};                                                                                                \
    return fdx::run2<fdx::DIV2, T>(plans, in, o, alpha, stream);                                                             \
  }                                                                                                                          \
  extern "C" int fdx_gradient2d_##SUFFIX(const fdx_plan_t plans[2], const T* x, T* const out[2], const double alpha[2],      \
                                         void* stream) {                                                                     \
    if (!out) return FDX_EINVAL;                                                                                             \
    const T* in[2] = {x, nullptr};                                                                                           \
    T* o[2] = {out[0], out[1]};                                                                                              \
    return fdx::run2<fdx::GRAD2, T>(plans, in, o, alpha, stream);                                                            \
  }

FDX_DEF2(f32, float)
FDX_DEF2(f64, double)
