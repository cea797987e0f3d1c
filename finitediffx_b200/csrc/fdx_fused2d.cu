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
//
// Big grids (>= 8 MB per field) take the persistent TMA-ring form of the same arithmetic (fused2d_tma_kernel below, the
// skeleton of fdx_coltma.cu): items = (512-byte strip) x (32 rows) claimed from a global counter, a producer warp fills
// the ring with halo boxes, four consumer warps walk 8 rows each with a register window along axis 0 and take the
// axis-1 neighbours from shared memory; the frame by a warp of its own.  Same fma chains: bit-identical results.
#include <algorithm>
#include <type_traits>

#include "fdx_tma.cuh"

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

// (explicitly rounded product / sum: the frame code is inlined into two kernels and must give the same bits in both --
// left to the compiler, `acc * alpha + other` may or may not be contracted into an fma)
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }

// one term at one point, from the band or the main stencil of its axis (the frame path)
template <typename T>
__device__ __forceinline__ T axis_dot(const Axis2<T>& a, const T* x, int64_t base, int64_t stride, int i) {
  // taps in ascending order; the loads of four taps are issued together (a frame point is a chain of cold loads
  // otherwise: ~1 us per tap).  A band tap whose coefficient is exactly zero is skipped (0 * inf must not become NaN).
  T acc = 0;
  const bool left = i < a.b.nl, right = i >= a.n - a.b.nr;
  if (left || right) {
    const int w = left ? a.b.wl : a.b.wr;
    const double* cf = left ? a.b.l + (int64_t)i * a.b.wl : a.b.r + (int64_t)(i - (a.n - a.b.nr)) * a.b.wr;
    const T* xp = x + base + (left ? (int64_t)0 : (int64_t)(a.n - a.b.wr) * stride);
    for (int c = 0; c < w; c += 4) {
      T k[4], v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int cu = min(c + u, w - 1);
        k[u] = c + u < w ? (T)__ldg(cf + cu) : (T)0;
        v[u] = __ldg(xp + (int64_t)cu * stride);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) acc = k[u] != (T)0 ? fma(k[u], v[u], acc) : acc;
    }
  } else {
    const int w = a.hi - a.lo + 1;
    const T* xp = x + base + (int64_t)(i + a.lo) * stride;
    for (int c = 0; c < w; c += 4) {
      T v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg(xp + (int64_t)min(c + u, w - 1) * stride);
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (c + u < w) acc = fma(a.m[c + u], v[u], acc);
    }
  }
  return mul_rn(acc, a.alpha);
}

// one frame point g (boundary rows with all their columns first, then the boundary vectors of the main rows): each term
// from its band or its main stencil as the point requires
template <Op2 OP, typename T, int V>
__device__ __forceinline__ void frame_point(const Params2<T>& prm, int64_t g) {
  const int n1 = prm.n1v * V;
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
  const int ci0 = 0, ci1 = OP == DIV2 ? 1 : 0;
  const T d0 = axis_dot<T>(prm.ax[0], prm.in[ci0], i1, n1, i0);               // along axis 0: column i1
  const T d1 = axis_dot<T>(prm.ax[1], prm.in[ci1], (int64_t)i0 * n1, 1, i1);  // along axis 1: row i0
  if constexpr (OP == GRAD2) {
    prm.out[0][at] = d0;
    prm.out[1][at] = d1;
  } else {
    prm.out[0][at] = add_rn(d0, d1);
  }
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
    if (g < prm.frame_points) frame_point<OP, T, V>(prm, g);
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

// ------------------------------------------------------------------------------ persistent TMA-ring form (big grids)
template <Op2 OP, typename T, int P>
struct Geo2 {
  static constexpr int V = 16 / (int)sizeof(T), W = 2 * P + 1, HXV = (P + V - 1) / V;
  static constexpr int TXV = 32, TX = TXV * V;
  static constexpr int NW = 4, RW = 8, RB = NW * RW, NFW = 2, NT = (NW + 1 + NFW) * 32;  // consumers, producer, frame warps
  // field a (axis-0 stencil): RB + 2P rows; with the x halos when it is also the field of the axis-1 stencil
  static constexpr int AX = OP == DIV2 ? TX : TX + 2 * HXV * V, AY = RB + 2 * P;
  static constexpr int A_BYTES = ((AX * AY * (int)sizeof(T) + 127) / 128) * 128;
  // field b (DIV2 only: the axis-1 stencil's field): RB rows with x halos
  static constexpr int BX = TX + 2 * HXV * V, BY = RB;
  static constexpr int B_BYTES = OP == DIV2 ? ((BX * BY * (int)sizeof(T) + 127) / 128) * 128 : 0;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = OP == DIV2 ? 3 : 4, CLAIM = 4;
  static constexpr int SMEM_BYTES = 256 + 128 + STAGES * STAGE_BYTES;
};

template <typename T>
struct TmaExtra {
  CUtensorMap tmap[2];
  int strips, rblocks;
  unsigned int items;
  unsigned int* ctr;
};

template <int I, int N, typename F>
__device__ __forceinline__ void static_for2(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for2<I + 1, N>(f);
  }
}

template <Op2 OP, typename T, int P>
__global__ void __launch_bounds__(224, 2) fused2d_tma_kernel(const __grid_constant__ Params2<T> prm, const __grid_constant__ TmaExtra<T> ex) {
  using namespace tma;
  using G = Geo2<OP, T, P>;
  constexpr int V = G::V, W = G::W, HXV = G::HXV, STAGES = G::STAGES, RW = G::RW;
  using VT = Vec<T, V>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint32_t smem0 = smem_u32(smem_raw);
  const uint32_t ring = (smem0 + 256u + 127u) & ~127u;
  const uint32_t bar_full = smem0, bar_empty = smem0 + 8u * STAGES, item_xy = smem0 + 128u;
  const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  const int b = blockIdx.x, gsz = gridDim.x;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8u * s, 1);
      mbar_init(bar_empty + 8u * s, G::NW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&ex.tmap[0]) : "memory");
    if constexpr (OP == DIV2) asm volatile("prefetch.tensormap [%0];" ::"l"(&ex.tmap[1]) : "memory");
  }
  __syncthreads();

  // ---- producer warp
  if (wrp == G::NW) {
    if (lane != 0) return;
    unsigned int base = 0;
    int left = 0;
    for (int k = 0;; ++k) {
      if (left == 0) base = atomicAdd(ex.ctr, (unsigned)G::CLAIM), left = G::CLAIM;
      const unsigned int t = base++;
      --left;
      const int s = k % STAGES;
      if (k >= STAGES) mbar_wait(bar_empty + 8u * s, (uint32_t)((k / STAGES) - 1) & 1u);
      if (t < ex.items) {
        const int rb = (int)(t / (unsigned)ex.strips), strip = (int)(t - (unsigned)rb * (unsigned)ex.strips);
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(item_xy + 8u * s), "r"(strip), "r"(rb) : "memory");
        const int x0 = (prm.v0 + strip * G::TXV) * V, y0 = prm.r0 + rb * G::RB;
        const uint32_t dst = ring + (uint32_t)s * G::STAGE_BYTES;
        if constexpr (OP == DIV2) {
          mbar_expect_tx(bar_full + 8u * s, (uint32_t)((G::AX * G::AY + G::BX * G::BY) * sizeof(T)));
          tma_load_2d(dst, &ex.tmap[0], bar_full + 8u * s, x0, y0 - P);
          tma_load_2d(dst + G::A_BYTES, &ex.tmap[1], bar_full + 8u * s, x0 - HXV * V, y0);
        } else {
          mbar_expect_tx(bar_full + 8u * s, (uint32_t)(G::AX * G::AY * sizeof(T)));
          tma_load_2d(dst, &ex.tmap[0], bar_full + 8u * s, x0 - HXV * V, y0 - P);
        }
      } else {
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(item_xy + 8u * s), "r"(-1), "r"(-1) : "memory");
        mbar_arrive(bar_full + 8u * s);
        break;
      }
    }
    __threadfence();
    if (atomicAdd(ex.ctr + 1, 1u) == (unsigned)gsz - 1u) {  // the last producer of the grid rewinds the counters
      ex.ctr[0] = 0u, ex.ctr[1] = 0u;
      __threadfence();
    }
    return;
  }
  // ---- frame warps
  if (wrp > G::NW) {
    const int fl = (wrp - G::NW - 1) * 32 + lane;
    for (int64_t g = (int64_t)b * (G::NFW * 32) + fl; g < prm.frame_points; g += (int64_t)gsz * (G::NFW * 32)) frame_point<OP, T, V>(prm, g);
    return;
  }
  // ---- consumer warps
  const Axis2<T>& A0 = prm.ax[0];
  const Axis2<T>& A1 = prm.ax[1];
  const int n1 = prm.n1v * V;
  constexpr int APITCH = G::AX * (int)sizeof(T), BPITCH = G::BX * (int)sizeof(T);
  constexpr int AXOFF = OP == DIV2 ? 0 : HXV * 16;  // byte offset of this lane's own vector inside a row of field a
  int stage = 0;
  uint32_t parity = 0;
  for (;;) {
    mbar_wait(bar_full + 8u * stage, parity);
    int strip, rb;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(strip), "=r"(rb) : "r"(item_xy + 8u * stage));
    if (strip < 0) break;
    const int jv = prm.v0 + strip * G::TXV + lane;
    const int i0 = prm.r0 + rb * G::RB + wrp * RW;
    const uint32_t sa = ring + (uint32_t)stage * G::STAGE_BYTES + (uint32_t)(wrp * RW * APITCH + AXOFF + lane * 16);  // box row of input row i0 - P
    const uint32_t sb = OP == DIV2 ? ring + (uint32_t)stage * G::STAGE_BYTES + (uint32_t)(G::A_BYTES + wrp * RW * BPITCH + lane * 16)
                                   : sa + (uint32_t)(P * APITCH) - (uint32_t)(HXV * 16);  // first window vector of output row i0
    const bool col_ok = jv < prm.v1;
    const int64_t at0 = (int64_t)i0 * n1 + (int64_t)jv * V;
    VT q[W];
    static_for2<0, W - 1>([&](auto s_) { q[s_.value] = lds_vec<T, V>(sa + (uint32_t)(s_.value * APITCH)); });
    static_for2<0, RW>([&](auto r_) {
      constexpr int r = r_.value;
      q[(W - 1 + r) % W] = lds_vec<T, V>(sa + (uint32_t)((W - 1 + r) * APITCH));
      // axis 0: taps over the window (slot (r + k) % W holds input row i0 + r - P + k)
      VT d0;
      if constexpr (sizeof(T) == 4) {  // packed pairs (FFMA2: two independent fmas, the bits of the scalar chain)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          unsigned long long a;
          asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(a) : "l"(pack2(A0.c[0], A0.c[0])), "l"(pack2(q[r % W].v[2 * h], q[r % W].v[2 * h + 1])));
          static_for2<1, W>([&](auto k_) {
            constexpr int k = k_.value;
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a) : "l"(pack2(A0.c[k], A0.c[k])), "l"(pack2(q[(r + k) % W].v[2 * h], q[(r + k) % W].v[2 * h + 1])));
          });
          unpack2(a, d0.v[2 * h], d0.v[2 * h + 1]);
        }
      } else {
#pragma unroll
        for (int v = 0; v < V; ++v) {
          T acc = A0.c[0] * q[r % W].v[v];
          static_for2<1, W>([&](auto k_) { acc = fma(A0.c[k_.value], q[(r + k_.value) % W].v[v], acc); });
          d0.v[v] = acc;
        }
      }
      // axis 1: this row's vector and its neighbours from shared memory
      T w[(2 * HXV + 1) * V];
#pragma unroll
      for (int u = 0; u < 2 * HXV + 1; ++u) {
        VT t;
        if (OP != DIV2 && u == HXV)
          t = q[(r + P) % W];  // the centre vector is already in the window
        else
          t = lds_vec<T, V>(sb + (uint32_t)(r * (OP == DIV2 ? BPITCH : APITCH) + u * 16));
#pragma unroll
        for (int v = 0; v < V; ++v) w[u * V + v] = t.v[v];
      }
      VT d1;
      if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float s0 = A1.c[0] * w[HXV * V + 2 * h - P], s1 = A1.c[0] * w[HXV * V + 2 * h + 1 - P];
#pragma unroll
          for (int k = 1; k < W; ++k) {
            const int j = HXV * V + 2 * h - P + k;
            if (j % 2 == 0) {
              unsigned long long a = pack2(s0, s1);
              asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a) : "l"(pack2(A1.c[k], A1.c[k])), "l"(pack2(w[j], w[j + 1])));
              unpack2(a, s0, s1);
            } else {
              s0 = fma(A1.c[k], w[j], s0), s1 = fma(A1.c[k], w[j + 1], s1);
            }
          }
          d1.v[2 * h] = s0, d1.v[2 * h + 1] = s1;
        }
      } else {
#pragma unroll
        for (int v = 0; v < V; ++v) {
          T acc = A1.c[0] * w[HXV * V + v - P];
#pragma unroll
          for (int k = 1; k < W; ++k) acc = fma(A1.c[k], w[HXV * V + v - P + k], acc);
          d1.v[v] = acc;
        }
      }
      if (col_ok && i0 + r < prm.r1) {
        const int64_t at = at0 + (int64_t)r * n1;
        if constexpr (OP == GRAD2) {
          st_vec<T, V>(prm.out[0] + at, d0);
          st_vec<T, V>(prm.out[1] + at, d1);
        } else {
#pragma unroll
          for (int v = 0; v < V; ++v) d0.v[v] += d1.v[v];
          st_vec<T, V>(prm.out[0] + at, d0);
        }
      }
    });
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty + 8u * stage);
    if (++stage == STAGES) stage = 0, parity ^= 1u;
  }
}

template <Op2 OP, typename T, int P>
static int launch2_tma(Params2<T>& prm, cudaStream_t st) {
  using namespace tma;
  using G = Geo2<OP, T, P>;
  const int rows = prm.r1 - prm.r0, vecs = prm.v1 - prm.v0;
  const int n0 = prm.ax[0].n, n1 = prm.n1v * G::V;
  if (rows <= 0 || vecs <= 0) return FDX_EUNSUPPORTED;
  if ((int64_t)n0 * n1 * (int64_t)sizeof(T) < (int64_t)env_int("FDX_FUSED2D_TMA_MIN_MB", 8) * (1 << 20) || env_int("FDX_FUSED2D_TMA_DISABLE", 0)) return FDX_EUNSUPPORTED;
  EncodeTiledFn enc = encode_fn();
  if (!enc) return FDX_EUNSUPPORTED;
  TmaExtra<T> ex;
  auto make = [&](CUtensorMap* m, const T* base, int bx, int by) {
    cuuint64_t dims[2] = {(cuuint64_t)n1, (cuuint64_t)n0};
    cuuint64_t strides[1] = {(cuuint64_t)n1 * sizeof(T)};
    cuuint32_t box[2] = {(cuuint32_t)bx, (cuuint32_t)by};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  if (!make(&ex.tmap[0], prm.in[0], G::AX, G::AY)) return FDX_EUNSUPPORTED;
  if (OP == DIV2 && !make(&ex.tmap[1], prm.in[1], G::BX, G::BY)) return FDX_EUNSUPPORTED;
  const int64_t strips = (vecs + G::TXV - 1) / G::TXV, rblocks = (rows + G::RB - 1) / G::RB;
  if (strips * rblocks > 0x3fffffffLL) return FDX_EUNSUPPORTED;
  ex.strips = (int)strips, ex.rblocks = (int)rblocks, ex.items = (unsigned)(strips * rblocks);
  ex.ctr = claim_slot(st);
  if (!ex.ctr) return FDX_EUNSUPPORTED;
  auto kern = fused2d_tma_kernel<OP, T, P>;
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
  const int per_sm = std::max(1, std::min(4, (int)(224 * 1024 / (G::SMEM_BYTES + 1024))));
  const int64_t grid = std::min<int64_t>(ex.items, (int64_t)num_sms() * per_sm);
  kern<<<(unsigned)grid, G::NT, G::SMEM_BYTES, st>>>(prm, ex);
  count_launch(OP == LAP2 ? "fused_laplacian2d_tma" : OP == DIV2 ? "fused_divergence2d_tma" : "fused_gradient2d_tma");
  FDX_LAUNCH_CHECK();
  return FDX_OK;
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
  prm.frame_blocks = 0;
  {
    const int rc = launch2_tma<OP, T, P>(prm, st);  // big grids: the persistent TMA-ring form
    if (rc != FDX_EUNSUPPORTED) return rc;
  }
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
