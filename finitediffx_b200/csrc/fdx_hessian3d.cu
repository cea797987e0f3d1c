// Fused 3-D hessian: replaces the reference's `hessian` on a 3-D scalar field (finitediffx/_src/finite_diff.py:460-529:
// ndim second differences + ndim (ndim - 1) / 2 nested first differences, each a `concatenate`, then ndim + 1 `stack`s)
// by ONE pass over the input: 1 read + 9 writes per point (10 s bytes; the composition of axis passes moves 18 s).
//
// The reference's result table is F[ax1 + ax2] (:499, :510), so for ndim = 3 the five surviving entries are
//   F[0] = D2_0 x,  F[1] = D_1 D_0 x,  F[2] = D_2 D_0 x (it replaces D2_1 x: SURVEY F4),  F[3] = D_2 D_1 x,  F[4] = D2_2 x
// and H[i][j] = F[i + j] (:524-529).  This kernel produces exactly those five fields and stores each of them in every
// slot (i, j) with i + j = key of the (3, 3, n0, n1, n2) result.
//
// Two launches (disjoint output points):
//   the march          the points whose rows are main rows on all three axes.  There the nested differences are pure
//                      tensor products of the main stencils (operators on different axes commute), so a CTA marches a
//                      (8 x 32 vectors) tile along axis 0: plane z is staged in shared memory (cp.async, zero fill outside
//                      the grid, a ring of four planes), a first stage writes gx = D_2 x of the tile and its y halo rows to
//                      shared memory, the second stage gives every thread D2_2 x, gy = D_1 x and D_1 gx of its vector, and
//                      three shift-register queues along z (x -> D2_0 x, gy -> D_0 D_1 x, gx -> D_0 D_2 x) complete
//                      planes z - P2 / z - P1.  All taps sit in the parameter block pre-multiplied by the scales.
//   the shell          the boundary shell (points in a band row of any axis: the reference's one-sided rows,
//                      finite_diff.py:75-88; along the last axis rounded to whole vectors): nested dense gathers with
//                      the plans' band tables, accumulated in float64 (a kernel of its own: 64 registers, full
//                      occupancy for its latency-bound gathers; as leading CTAs of the march it ran 5x slower).
#include <algorithm>

#include "fdx_common.cuh"

namespace fdx {
namespace {

// P1 = reach of the first-derivative stencils (all three axes), P2 = P1 + 1 = reach of the second-derivative stencils
// (axes 0 and 2): at equal accuracy the reference's second-derivative stencil is one point longer on each side
// (utils.py:59-79 through finite_diff.py:247-258); shorter plans are padded with zero taps.
template <typename T, int P1>
struct HessGeom {
  static constexpr int V = 16 / (int)sizeof(T);
  static constexpr int P2 = P1 + 1;
  static constexpr int W1 = 2 * P1 + 1, W2 = 2 * P2 + 1;
  static constexpr int HXV = (P2 + V - 1) / V;  // halo vectors per side
  static constexpr int HX = HXV * V;
  static constexpr int TXV = 32, TX = TXV * V, TY = 8;
  static constexpr int ROWS = TY + 2 * P1;
  static constexpr int XSV = TXV + 2 * HXV;  // vectors per staged row
  static constexpr int WV = 1 + 2 * HXV;     // vectors of a thread's x window
  static constexpr int XS_ELEMS = ROWS * XSV * V;
  static constexpr int GS_ELEMS = ROWS * TX;
  static constexpr int NT = 256;
  static constexpr int NS = 4;  // planes in flight (cp.async ring)
  static constexpr int NLD = (ROWS * XSV + NT - 1) / NT;
  static constexpr int NGX = (ROWS * TXV + NT - 1) / NT;
};

struct XBand {
  const double* l;
  const double* r;
  int32_t nl, wl, nr, wr;
  double alpha;
};
constexpr int kXBandMax = 4;  // band columns per side the march handles (reach of the second derivative <= 4)
constexpr int kXBandW = 12;   // taps per band row (first-derivative reach <= 3: wl <= 4 + 7)
constexpr int kYBandMax = 3;  // band rows of axis 1 per side (first-derivative reach <= 3)

// One band element: kXBandW taps from the CTA's table in shared memory (scaled, zero-padded) over a staged row; `row`
// points at the staged element under the band's first column, `w` is the band's width (lanes past it re-read its last
// column under a zero tap).  Zero taps never touch their data (0 * inf).
template <typename T>
__device__ __forceinline__ T xband_dot(const T* tab, int w, const T* row) {
  T acc = (T)0;
#pragma unroll
  for (int c = 0; c < kXBandW; ++c) {
    const T k = tab[c];
    const T v = row[min(c, w - 1)];
    acc = fma(k, k != (T)0 ? v : (T)0, acc);
  }
  return acc;
}

template <typename T, int P1>
struct HessParams {
  const T* x;
  T* h;  // (3, 3, n0, n1, n2)
  int32_t n0, n1, n2;
  int32_t z_lo, z_hi;  // output planes, rows and (vector-aligned) columns of the march: main rows on all three axes
  int32_t y_lo, y_hi, x_lo, x_hi;
  int32_t zc;          // planes per chunk
  int32_t tiles_x, tiles_y;
  // radius-P2 / radius-P1 taps (zero where a plan is shorter), scales folded in
  T c2z[2 * P1 + 3], c2x[2 * P1 + 3], c1z[2 * P1 + 1], c1y[2 * P1 + 1], c1x[2 * P1 + 1];
  int32_t plain_stores;  // FDX_HESS_PLAIN_ST=1: write-back stores instead of streaming (evict-first) ones
  // boundary rows of the last axis (dense bands of the first- / second-derivative plan, unscaled float64 in device memory):
  // evaluated by the march itself from its staged rows, a lane per band column
  XBand xb1, xb2;
  XBand yb;  // boundary rows of axis 1 (first derivative): a warp whose row is a band row runs the dense row instead
};

// ---- boundary shell -------------------------------------------------------------------------------------------------
struct AxisTab {
  const double* bl;
  const double* br;
  const double* main;  // device copy of the main taps (an indexed constant-bank load per tap made the shell LDC-bound)
  int32_t n, lo, hi, nl, wl, nr, wr;
};
struct RowOp {
  const double* c;  // the row's taps (device memory)
  int32_t s, cnt;
};
__device__ __forceinline__ RowOp row_of(const AxisTab& a, int i) {
  if (i < a.nl) return RowOp{a.bl + (int64_t)i * a.wl, 0, a.wl};
  if (i >= a.n - a.nr) return RowOp{a.br + (int64_t)(i - (a.n - a.nr)) * a.wr, a.n - a.wr, a.wr};
  return RowOp{a.main, i + a.lo, a.hi - a.lo + 1};
}
__device__ __forceinline__ double tap_of(const AxisTab&, const RowOp& r, int k) { return __ldg(r.c + k); }

// One row of an axis operator applied to a strided line: taps in batches of eight, all loads of a batch issued before
// the first use (the shell is latency-bound: one load in flight per thread cost 275 us at 384^3, five times the march's
// share of the work).  Lanes past the row's length re-read its last element with a zero coefficient.
template <typename T>
__device__ __forceinline__ double row_dot(const AxisTab& a, const RowOp& r, const T* base, int64_t stride) {
  double acc = 0;
  const T* p0 = base + (int64_t)r.s * stride;
  for (int k0 = 0; k0 < r.cnt; k0 += 8) {
    double c[8];
    T v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int k = min(k0 + u, r.cnt - 1);
      c[u] = k0 + u < r.cnt ? tap_of(a, r, k) : 0.0;
      v[u] = p0[(int64_t)k * stride];
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) acc = fma(c[u], c[u] != 0.0 ? (double)v[u] : 0.0, acc);
  }
  return acc;
}

template <typename T>
struct ShellParams {
  const T* x;
  T* h;
  int32_t n0, n1, n2;
  int32_t bz0, bz1, by0, by1, bx0, bx1;  // main (band-free) index ranges per axis
  int64_t cnt_a, cnt_b, cnt_c;
  double s00, s22, s01, s02, s12;  // scales of the five entries
  AxisTab d1z, d1y, d1x, d2z, d2x;
};

template <typename T>
__device__ __forceinline__ void shell_point(const ShellParams<T>& prm, int64_t idx) {
  int z, y, x;
  const int nzb = prm.n0 - (prm.bz1 - prm.bz0), nyb = prm.n1 - (prm.by1 - prm.by0), nxb = prm.n2 - (prm.bx1 - prm.bx0);
  if (idx < prm.cnt_a) {  // band planes of axis 0, whole planes
    const int64_t pl = (int64_t)prm.n1 * prm.n2;
    const int zi = (int)(idx / pl);
    const int64_t rem = idx - (int64_t)zi * pl;
    z = zi < prm.bz0 ? zi : prm.bz1 + (zi - prm.bz0);
    y = (int)(rem / prm.n2), x = (int)(rem - (int64_t)y * prm.n2);
    (void)nzb;
  } else if (idx < prm.cnt_a + prm.cnt_b) {  // main planes, band rows of axis 1
    idx -= prm.cnt_a;
    const int64_t pl = (int64_t)nyb * prm.n2;
    const int zi = (int)(idx / pl);
    const int64_t rem = idx - (int64_t)zi * pl;
    const int yi = (int)(rem / prm.n2);
    z = prm.bz0 + zi;
    y = yi < prm.by0 ? yi : prm.by1 + (yi - prm.by0);
    x = (int)(rem - (int64_t)yi * prm.n2);
  } else if (idx < prm.cnt_a + prm.cnt_b + prm.cnt_c) {  // main planes and rows, band columns of axis 2
    idx -= prm.cnt_a + prm.cnt_b;
    const int64_t pl = (int64_t)(prm.by1 - prm.by0) * nxb;
    const int zi = (int)(idx / pl);
    const int64_t rem = idx - (int64_t)zi * pl;
    const int yi = (int)(rem / nxb), xi = (int)(rem - (int64_t)yi * nxb);
    z = prm.bz0 + zi, y = prm.by0 + yi;
    x = xi < prm.bx0 ? xi : prm.bx1 + (xi - prm.bx0);
  } else {
    return;
  }
  const int64_t s1 = prm.n2, s0 = (int64_t)prm.n1 * prm.n2, N = s0 * prm.n0;
  const T* X = prm.x;
  const int64_t at = (int64_t)z * s0 + (int64_t)y * s1 + x;

  const RowOp rz = row_of(prm.d1z, z), ry = row_of(prm.d1y, y), rx = row_of(prm.d1x, x);
  const RowOp rz2 = row_of(prm.d2z, z), rx2 = row_of(prm.d2x, x);
  const double d00 = row_dot<T>(prm.d2z, rz2, X + (int64_t)y * s1 + x, s0);
  const double d22 = row_dot<T>(prm.d2x, rx2, X + (int64_t)z * s0 + (int64_t)y * s1, 1);
  double d01 = 0, d02 = 0, d12 = 0;
  for (int k = 0; k < rz.cnt; ++k) {
    const double cz = tap_of(prm.d1z, rz, k);
    if (cz == 0.0) continue;  // (zero taps never touch their data: 0 * inf stays out, as in the axis kernels' band rows)
    const T* pz = X + (int64_t)(rz.s + k) * s0;
    d01 = fma(cz, row_dot<T>(prm.d1y, ry, pz + x, s1), d01);
    d02 = fma(cz, row_dot<T>(prm.d1x, rx, pz + (int64_t)y * s1, 1), d02);
  }
  for (int j = 0; j < ry.cnt; ++j) {
    const double cy = tap_of(prm.d1y, ry, j);
    if (cy == 0.0) continue;
    d12 = fma(cy, row_dot<T>(prm.d1x, rx, X + (int64_t)z * s0 + (int64_t)(ry.s + j) * s1, 1), d12);
  }
  T* o = prm.h + at;
  const T v0 = (T)(d00 * prm.s00), v1 = (T)(d01 * prm.s01), v2 = (T)(d02 * prm.s02), v3 = (T)(d12 * prm.s12), v4 = (T)(d22 * prm.s22);
  o[0] = v0;
  o[1 * N] = v1, o[3 * N] = v1;
  o[2 * N] = v2, o[4 * N] = v2, o[6 * N] = v2;
  o[5 * N] = v3, o[7 * N] = v3;
  o[8 * N] = v4;
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  const int sz = valid ? 16 : 0;  // src-size 0: the 16 bytes are zero-filled, nothing is read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename T, int V>
__device__ __forceinline__ void st_cs(T* p, const Vec<T, V>& r, bool plain = false) {
  if (plain) {
    st_vec<T, V>(p, r);
    return;
  }
  if constexpr (sizeof(T) == 4) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(r.v[0], r.v[1], r.v[2], r.v[3]));
  } else {
    __stcs(reinterpret_cast<double2*>(p), make_double2(r.v[0], r.v[1]));
  }
}

template <typename T>
__global__ void __launch_bounds__(128) hessian3d_shell(const __grid_constant__ ShellParams<T> prm) {
  shell_point<T>(prm, (int64_t)blockIdx.x * 128 + threadIdx.x);
}

template <typename T, int P1, int MINB>
__global__ void __launch_bounds__(256, MINB) hessian3d_kernel(const __grid_constant__ HessParams<T, P1> prm) {
  using G = HessGeom<T, P1>;
  constexpr int V = G::V, P2 = G::P2, W1 = G::W1, W2 = G::W2, HX = G::HX, HXV = G::HXV, XSV = G::XSV, TX = G::TX;
  using VecT = Vec<T, V>;
  constexpr int NS = G::NS;
  __shared__ __align__(16) T xs[NS][G::XS_ELEMS];
  __shared__ __align__(16) T gs[G::GS_ELEMS];
  __shared__ T bt[4][kXBandMax][kXBandW];     // band rows of the last axis: {D_2 left, D_2 right, D2_2 left, D2_2 right}
  __shared__ T d22b[G::TY][2 * kXBandMax];    // D2_2 x on the band columns of the tile rows (left | right)
  __shared__ T ybt[2][kYBandMax][kXBandW];    // band rows of axis 1 (top | bottom), scaled

  const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  int b = blockIdx.x;
  const int tx = b % prm.tiles_x;
  b /= prm.tiles_x;
  const int ty = b % prm.tiles_y;
  b /= prm.tiles_y;
  const int zc0 = prm.z_lo + b * prm.zc, zc1 = min(zc0 + prm.zc, prm.z_hi);
  // (the last tile of a row is anchored at the end of the row: it overlaps its neighbour -- the same values are stored
  // twice -- and always holds the columns under the right band rows)
  const int x0 = min(tx * TX, max(0, prm.n2 - TX)), y0 = min(ty * G::TY, max(0, prm.n1 - G::TY));
  const bool edge_l = x0 == 0 && (prm.xb1.nl > 0 || prm.xb2.nl > 0), edge_r = x0 + TX >= prm.n2 && (prm.xb1.nr > 0 || prm.xb2.nr > 0);
  if ((edge_l | edge_r) && tid < 4 * kXBandMax * kXBandW) {
    const int t = tid / (kXBandMax * kXBandW), j = (tid / kXBandW) % kXBandMax, c = tid % kXBandW;
    const XBand& xb = t < 2 ? prm.xb1 : prm.xb2;
    const bool left = (t & 1) == 0;
    const int rows = left ? xb.nl : xb.nr, w = left ? xb.wl : xb.wr;
    const double* tab = left ? xb.l : xb.r;
    bt[t][j][c] = (j < rows && c < w) ? (T)(tab[j * w + c] * xb.alpha) : (T)0;  // (visible after the loop's first barrier)
  }
  const bool edge_t = y0 < prm.yb.nl, edge_b = y0 + G::TY > prm.n1 - prm.yb.nr;
  if ((edge_t | edge_b) && tid < 2 * kYBandMax * kXBandW) {
    const int u = tid, t = u / (kYBandMax * kXBandW), j = (u / kXBandW) % kYBandMax, c = u % kXBandW;
    const int rows = t == 0 ? prm.yb.nl : prm.yb.nr, w = t == 0 ? prm.yb.wl : prm.yb.wr;
    const double* tab = t == 0 ? prm.yb.l : prm.yb.r;
    ybt[t][j][c] = (j < rows && c < w) ? (T)(tab[j * w + c] * prm.yb.alpha) : (T)0;
  }
  // helper work of an edge tile, done by warps 4..7 in stage B (warps 0..3 have a second row there): D_2 x on the band
  // columns of every staged row -> gs, D2_2 x on the band columns of the tile rows -> d22b
  const int hb0 = edge_l ? G::ROWS * prm.xb1.nl : 0, hb1 = hb0 + (edge_r ? G::ROWS * prm.xb1.nr : 0);
  const int hb2 = hb1 + (edge_l ? G::TY * prm.xb2.nl : 0), hb3 = hb2 + (edge_r ? G::TY * prm.xb2.nr : 0);
  // band elements of this lane's vector (bit p: element p is a band column of D_2 / D2_2)
  unsigned bm1 = 0, bm2 = 0;
#pragma unroll
  for (int p = 0; p < V; ++p) {
    const int col = x0 + lane * V + p;
    if ((edge_l && col < prm.xb1.nl) || (edge_r && col >= prm.n2 - prm.xb1.nr && col < prm.n2)) bm1 |= 1u << p;
    if ((edge_l && col < prm.xb2.nl) || (edge_r && col >= prm.n2 - prm.xb2.nr && col < prm.n2)) bm2 |= 1u << p;
  }
  const int64_t plane = (int64_t)prm.n1 * prm.n2;
  const int64_t N = plane * prm.n0;
  const bool plain = prm.plain_stores != 0;

  // what this thread stages of every plane (the same slots for all planes)
  int goff[G::NLD], soff[G::NLD];
  bool ok[G::NLD];
#pragma unroll
  for (int s = 0; s < G::NLD; ++s) {
    const int idx = tid + s * G::NT;
    soff[s] = -1, goff[s] = 0, ok[s] = false;
    if (idx < G::ROWS * XSV) {
      const int r = idx / XSV, v = idx - r * XSV;
      const int yy = y0 - P1 + r, xx = x0 - HX + v * V;
      soff[s] = idx * V;
      ok[s] = yy >= 0 && yy < prm.n1 && xx >= 0 && xx < prm.n2;
      goff[s] = ok[s] ? yy * prm.n2 + xx : 0;
    }
  }
  auto stage = [&](int z, int buf) {
    const bool zin = z >= 0 && z < prm.n0;
    const T* src = prm.x + (zin ? (int64_t)z * plane : 0);
#pragma unroll
    for (int s = 0; s < G::NLD; ++s)
      if (soff[s] >= 0) cp_async16(&xs[buf][soff[s]], src + goff[s], zin && ok[s]);
  };

  // own vector
  const int yo = y0 + wrp, xo = x0 + lane * V;
  const bool own = yo >= prm.y_lo && yo < prm.y_hi && xo >= prm.x_lo && xo < prm.x_hi;
  // a warp whose row is a boundary row of axis 1: band table (0 top, 1 bottom), row in it, staged row under its first column
  const int ybk = yo < prm.yb.nl ? 0 : (yo >= prm.n1 - prm.yb.nr && yo < prm.n1 ? 1 : -1);
  const int ybj = ybk == 0 ? yo : yo - (prm.n1 - prm.yb.nr);
  const int ybw = ybk == 0 ? prm.yb.wl : prm.yb.wr;
  const int ybr = (ybk == 0 ? 0 : prm.n1 - prm.yb.wr) - y0 + P1;
  const int64_t ooff = (int64_t)yo * prm.n2 + xo;

  T q0[2 * P2][V], q1[2 * P1][V], q2[2 * P1][V];
#pragma unroll
  for (int j = 0; j < 2 * P2; ++j)
#pragma unroll
    for (int p = 0; p < V; ++p) q0[j][p] = (T)0;
#pragma unroll
  for (int j = 0; j < 2 * P1; ++j)
#pragma unroll
    for (int p = 0; p < V; ++p) q1[j][p] = q2[j][p] = (T)0;

  const int zfirst = zc0 - P2, zlast = zc1 - 1 + P2;
#pragma unroll
  for (int s = 0; s < NS - 1; ++s) {  // (one group per plane, empty groups past the end keep the count uniform)
    if (zfirst + s <= zlast) stage(zfirst + s, s);
    cp_async_commit();
  }
  int cur = 0;
  for (int z = zfirst; z <= zlast; ++z) {
    cp_async_wait<NS - 2>();  // plane z has landed
    __syncthreads();          // ... for every thread; and the buffer of plane z - 1 is free
    {
      const int nb = cur == 0 ? NS - 1 : cur - 1;
      if (z + NS - 1 <= zlast) stage(z + NS - 1, nb);
      cp_async_commit();
    }
    const T* X = xs[cur];
    cur = cur + 1 == NS ? 0 : cur + 1;

    // ---- stage B: gx = D_2 x on the tile rows and their y halo
#pragma unroll
    for (int s = 0; s < G::NGX; ++s) {
      const int r = wrp + s * 8;
      if (r < G::ROWS) {
        T w[G::WV * V];
#pragma unroll
        for (int u = 0; u < G::WV; ++u) {
          const VecT t = *reinterpret_cast<const VecT*>(&X[(r * XSV + lane + u) * V]);
#pragma unroll
          for (int p = 0; p < V; ++p) w[u * V + p] = t.v[p];
        }
        VecT g;
#pragma unroll
        for (int p = 0; p < V; ++p) {
          T a = (T)0;
#pragma unroll
          for (int k = 0; k < W1; ++k) a = fma(prm.c1x[k], w[HX - P1 + p + k], a);
          g.v[p] = a;
        }
        if (bm1 == 0) {
          *reinterpret_cast<VecT*>(&gs[r * TX + lane * V]) = g;
        } else {  // (the band elements of this vector come from the helper lanes below)
#pragma unroll
          for (int p = 0; p < V; ++p)
            if (!((bm1 >> p) & 1)) gs[r * TX + lane * V + p] = g.v[p];
        }
      }
    }
    if (tid >= 128 && hb3 > 0) {
      for (int q = tid - 128; q < hb3; q += 128) {
        const bool second = q >= hb1;                              // D2_2 (tile rows) instead of D_2 (staged rows)
        const bool right = second ? q >= hb2 : q >= hb0;
        const XBand& xb = second ? prm.xb2 : prm.xb1;
        const int nb = right ? xb.nr : xb.nl, w = right ? xb.wr : xb.wl;
        const int e = q - (second ? (right ? hb2 : hb1) : (right ? hb0 : 0));
        const int row = e / nb, j = e - row * nb;
        const int r = second ? row + P1 : row;
        const int c0 = right ? prm.n2 - w : 0;                      // first grid column under the band
        const T v = xband_dot<T>(bt[(second ? 2 : 0) + (right ? 1 : 0)][j], w, &X[r * XSV * V + HX + c0 - x0]);
        if (second)
          d22b[row][(right ? kXBandMax : 0) + j] = v;
        else
          gs[r * TX + (right ? prm.n2 - nb + j : j) - x0] = v;
      }
    }
    __syncthreads();

    // ---- stage C: the thread's own vector of plane z
    const bool inplane = z >= zc0 && z < zc1;
    T xc[V], gyv[V], gxc[V];
    {
      T w[G::WV * V];
      const int r = wrp + P1;
#pragma unroll
      for (int u = 0; u < G::WV; ++u) {
        const VecT t = *reinterpret_cast<const VecT*>(&X[(r * XSV + lane + u) * V]);
#pragma unroll
        for (int p = 0; p < V; ++p) w[u * V + p] = t.v[p];
      }
#pragma unroll
      for (int p = 0; p < V; ++p) xc[p] = w[HX + p];
      VecT d22, d12;
#pragma unroll
      for (int p = 0; p < V; ++p) {
        T a = (T)0;
#pragma unroll
        for (int k = 0; k < W2; ++k) a = fma(prm.c2x[k], w[HX - P2 + p + k], a);
        d22.v[p] = a;
        gyv[p] = (T)0, d12.v[p] = (T)0;
      }
      if (bm2 != 0) {  // band columns of D2_2 x: from the helper lanes
#pragma unroll
        for (int p = 0; p < V; ++p) {
          const int col = x0 + lane * V + p;
          if ((bm2 >> p) & 1) d22.v[p] = col < prm.xb2.nl && edge_l ? d22b[wrp][col] : d22b[wrp][kXBandMax + col - (prm.n2 - prm.xb2.nr)];
        }
      }
      if (ybk >= 0) {  // (warp-uniform) dense boundary row of axis 1 over the staged rows; zero taps skipped
        const VecT gc = *reinterpret_cast<const VecT*>(&gs[(wrp + P1) * TX + lane * V]);
#pragma unroll
        for (int p = 0; p < V; ++p) gxc[p] = gc.v[p];
        for (int c = 0; c < ybw; ++c) {
          const T k = ybt[ybk][ybj][c];
          if (k == (T)0) continue;
          const VecT gj = *reinterpret_cast<const VecT*>(&gs[(ybr + c) * TX + lane * V]);
          const VecT xj = *reinterpret_cast<const VecT*>(&X[((ybr + c) * XSV + lane + HXV) * V]);
#pragma unroll
          for (int p = 0; p < V; ++p) {
            gyv[p] = fma(k, xj.v[p], gyv[p]);
            d12.v[p] = fma(k, gj.v[p], d12.v[p]);
          }
        }
      } else
#pragma unroll
      for (int j = 0; j < W1; ++j) {
        const VecT gj = *reinterpret_cast<const VecT*>(&gs[(wrp + j) * TX + lane * V]);
        VecT xj;
        if (j == P1) {
#pragma unroll
          for (int p = 0; p < V; ++p) xj.v[p] = xc[p], gxc[p] = gj.v[p];
        } else {
          xj = *reinterpret_cast<const VecT*>(&X[((wrp + j) * XSV + lane + HXV) * V]);
        }
#pragma unroll
        for (int p = 0; p < V; ++p) {
          gyv[p] = fma(prm.c1y[j], xj.v[p], gyv[p]);
          d12.v[p] = fma(prm.c1y[j], gj.v[p], d12.v[p]);
        }
      }
      if (inplane && own) {
        T* o = prm.h + (int64_t)z * plane + ooff;
        st_cs<T, V>(o + 8 * N, d22, plain);  // key 4: (2,2)
        st_cs<T, V>(o + 5 * N, d12, plain);  // key 3: (1,2), (2,1)
        st_cs<T, V>(o + 7 * N, d12, plain);
      }
    }
    // ---- z queues: plane z completes D2_0 x of plane z - P2 and the mixed entries of plane z - P1
    VecT d00, d01, d02;
#pragma unroll
    for (int p = 0; p < V; ++p) {
      d00.v[p] = fma(prm.c2z[2 * P2], xc[p], q0[0][p]);
      d01.v[p] = fma(prm.c1z[2 * P1], gyv[p], q1[0][p]);
      d02.v[p] = fma(prm.c1z[2 * P1], gxc[p], q2[0][p]);
    }
#pragma unroll
    for (int j = 1; j < 2 * P2; ++j)
#pragma unroll
      for (int p = 0; p < V; ++p) q0[j - 1][p] = fma(prm.c2z[2 * P2 - j], xc[p], q0[j][p]);
#pragma unroll
    for (int j = 1; j < 2 * P1; ++j)
#pragma unroll
      for (int p = 0; p < V; ++p) {
        q1[j - 1][p] = fma(prm.c1z[2 * P1 - j], gyv[p], q1[j][p]);
        q2[j - 1][p] = fma(prm.c1z[2 * P1 - j], gxc[p], q2[j][p]);
      }
#pragma unroll
    for (int p = 0; p < V; ++p) {
      q0[2 * P2 - 1][p] = prm.c2z[0] * xc[p];
      q1[2 * P1 - 1][p] = prm.c1z[0] * gyv[p];
      q2[2 * P1 - 1][p] = prm.c1z[0] * gxc[p];
    }
    const int zo2 = z - P2, zo = z - P1;
    if (zo2 >= zc0 && zo2 < zc1 && own) st_cs<T, V>(prm.h + (int64_t)zo2 * plane + ooff, d00, plain);  // key 0: (0,0)
    if (zo >= zc0 && zo < zc1 && own) {
      T* o = prm.h + (int64_t)zo * plane + ooff;
      st_cs<T, V>(o + 1 * N, d01, plain);  // key 1: (0,1), (1,0)
      st_cs<T, V>(o + 3 * N, d01, plain);
      st_cs<T, V>(o + 2 * N, d02, plain);  // key 2: (0,2), (1,1), (2,0)
      st_cs<T, V>(o + 4 * N, d02, plain);
      st_cs<T, V>(o + 6 * N, d02, plain);
    }
  }
}

inline AxisTab tab_of(const fdx_plan_s* p) {
  AxisTab a;
  a.bl = p->band_l, a.br = p->band_r, a.main = p->d_main;
  a.n = p->n, a.lo = p->lo, a.hi = p->hi, a.nl = p->nl, a.wl = p->wl, a.nr = p->nr, a.wr = p->wr;
  return a;
}

template <typename T, int P>
void fill_taps(T (&c)[2 * P + 1], const fdx_plan_s* p, double scale) {  // (P: the radius the kernel runs this stencil at)
  for (int t = 0; t < 2 * P + 1; ++t) {
    const int o = t - P;
    c[t] = (o >= p->lo && o <= p->hi) ? (T)(p->main_taps[o - p->lo] * scale) : (T)0;
  }
}

template <typename T, int P1>
int launch_hess(const fdx_plan_s* const d1[3], const fdx_plan_s* d2z, const fdx_plan_s* d2x, const T* x, T* h, const double a1[3],
                double a2z, double a2x, const ShellParams<T>& sh, cudaStream_t st) {
  using G = HessGeom<T, P1>;
  HessParams<T, P1> prm;
  prm.x = x, prm.h = h;
  prm.n0 = sh.n0, prm.n1 = sh.n1, prm.n2 = sh.n2;
  prm.z_lo = sh.bz0, prm.z_hi = sh.bz1;
  prm.plain_stores = env_int("FDX_HESS_PLAIN_ST", 0);
  prm.xb1 = XBand{d1[2]->band_l, d1[2]->band_r, d1[2]->nl, d1[2]->wl, d1[2]->nr, d1[2]->wr, a1[2]};
  prm.xb2 = XBand{d2x->band_l, d2x->band_r, d2x->nl, d2x->wl, d2x->nr, d2x->wr, a2x};
  prm.yb = XBand{d1[1]->band_l, d1[1]->band_r, d1[1]->nl, d1[1]->wl, d1[1]->nr, d1[1]->wr, a1[1]};
  fill_taps<T, P1 + 1>(prm.c2z, d2z, a2z);
  fill_taps<T, P1 + 1>(prm.c2x, d2x, a2x);
  fill_taps<T, P1>(prm.c1z, d1[0], a1[0]);
  fill_taps<T, P1>(prm.c1y, d1[1], a1[1]);
  fill_taps<T, P1>(prm.c1x, d1[2], a1[2]);
  prm.y_lo = sh.by0, prm.y_hi = sh.by1, prm.x_lo = sh.bx0, prm.x_hi = sh.bx1;
  prm.tiles_x = (sh.n2 + G::TX - 1) / G::TX, prm.tiles_y = (sh.n1 + G::TY - 1) / G::TY;
  prm.zc = 1;
  const int nz = sh.bz1 - sh.bz0;
  const int part = env_int("FDX_HESS_PART", 0);  // timing tools: 1 = the march alone, 2 = the shell alone
  int64_t main_blocks = 0;
  if (sh.by1 > sh.by0 && sh.bx1 > sh.bx0 && nz > 0 && part != 2) {
    const int64_t tiles = (int64_t)prm.tiles_x * prm.tiles_y;
    // z chunks of ~64 planes (a chunk re-reads 2 P2 planes), shorter when that leaves fewer than ~3 CTAs per SM
    int zc = env_int("FDX_HESS_ZC", 0);
    if (zc <= 0) {
      const int64_t want = (int64_t)env_int("FDX_HESS_CPS", 3) * num_sms();
      const int64_t chunks = std::max<int64_t>((nz + 63) / 64, (want + tiles - 1) / tiles);
      zc = std::max((int)((nz + chunks - 1) / chunks), 16);
    }
    zc = std::min(zc, nz);
    prm.zc = zc;
    main_blocks = tiles * ((nz + zc - 1) / zc);
  }
  const int64_t pts = part == 1 ? 0 : sh.cnt_a + sh.cnt_b + sh.cnt_c;
  const int64_t shell_blocks = (pts + 127) / 128;
  if (main_blocks + shell_blocks > 0x7fffffffLL) return FDX_EUNSUPPORTED;
  if (main_blocks + shell_blocks == 0) return FDX_OK;
  if (shell_blocks) {
    hessian3d_shell<T><<<(unsigned)shell_blocks, 128, 0, st>>>(sh);
    count_launch("fused_hessian3d");
    FDX_LAUNCH_CHECK();
  }
  if (main_blocks) {
    bool three = false;
    if constexpr (P1 == 2) three = env_int("FDX_HESS_MINB", 2) == 3;
    if constexpr (P1 == 2) {
      if (three) hessian3d_kernel<T, P1, 3><<<(unsigned)main_blocks, G::NT, 0, st>>>(prm);
    }
    if (!three) hessian3d_kernel<T, P1, (P1 <= 1 ? 3 : 2)><<<(unsigned)main_blocks, G::NT, 0, st>>>(prm);
    count_launch("fused_hessian3d");
    FDX_LAUNCH_CHECK();
  }
  return FDX_OK;
}

template <typename T>
int run_hess(const fdx_plan_t d1_in[3], const fdx_plan_t d2_in[3], const T* x, T* h, const double a1[3], const double a2[3], void* stream) {
  constexpr int V = 16 / (int)sizeof(T);
  if (!d1_in || !d2_in || !a1 || !a2 || !x || !h) return FDX_EINVAL;
  if (!d1_in[0] || !d1_in[1] || !d1_in[2] || !d2_in[0] || !d2_in[2]) return FDX_EINVAL;
  if (env_int("FDX_HESS_DISABLE", 0)) return FDX_EUNSUPPORTED;  // (tests: the composition of axis passes instead)
  const fdx_plan_s* d1[3] = {d1_in[0], d1_in[1], d1_in[2]};
  const fdx_plan_s *d2z = d2_in[0], *d2x = d2_in[2];
  if (d2z->n != d1[0]->n || d2x->n != d1[2]->n) return FDX_EINVAL;
  const int n0 = d1[0]->n, n1 = d1[1]->n, n2 = d1[2]->n;
  if (n0 == 0 || n1 == 0 || n2 == 0) return FDX_OK;
  const fdx_plan_s* all[5] = {d1[0], d1[1], d1[2], d2z, d2x};
  int P1 = 1;
  for (int k = 0; k < 5; ++k) {
    const fdx_plan_s* p = all[k];
    if (p->nl + p->nr > p->n) return FDX_EUNSUPPORTED;  // dense (tiny-n) plans: composition
    if (p->nl < -p->lo || p->nr < p->hi) return FDX_EUNSUPPORTED;
    if (p->wl > p->n || p->wr > p->n) return FDX_EUNSUPPORTED;
    const int reach = std::max(std::abs(p->lo), std::abs(p->hi));
    P1 = std::max(P1, k < 3 ? reach : reach - 1);  // the kernel runs the second derivatives at radius P1 + 1
  }
  if (P1 > 3) return FDX_EUNSUPPORTED;
  if (n2 % V != 0 || ((uintptr_t)x & 15) != 0 || ((uintptr_t)h & 15) != 0) return FDX_EUNSUPPORTED;
  if ((int64_t)n1 * n2 >= 0x7fffffffLL) return FDX_EUNSUPPORTED;
  if (((int64_t)n0 * n1 * n2) % V != 0) return FDX_EUNSUPPORTED;

  ShellParams<T> sh;
  sh.x = x, sh.h = h, sh.n0 = n0, sh.n1 = n1, sh.n2 = n2;
  sh.bz0 = std::max(d1[0]->nl, d2z->nl), sh.bz1 = n0 - std::max(d1[0]->nr, d2z->nr);
  // (rows: the march runs the band rows of axis 1 itself, a warp per row; its last tile is anchored at the end of the axis
  // so that the staged rows hold every row under the bottom band)
  sh.by0 = 0, sh.by1 = n1;
  if (d1[1]->nl > kYBandMax || d1[1]->nr > kYBandMax || d1[1]->wl > kXBandW || d1[1]->wr > kXBandW) return FDX_EUNSUPPORTED;
  // (columns: the march evaluates the band rows of the last axis itself -- a sector written half by the march and half by
  // the shell is two partial writes (+50 % on the fp32 march), a shell of whole sectors is 16 columns per row)
  sh.bx0 = 0, sh.bx1 = n2;
  if (std::max(d1[2]->nl, d2x->nl) > kXBandMax || std::max(d1[2]->nr, d2x->nr) > kXBandMax) return FDX_EUNSUPPORTED;
  if (sh.bz1 < sh.bz0) sh.bz1 = sh.bz0;
  if (sh.by1 < sh.by0) sh.by1 = sh.by0;
  if (sh.bx1 < sh.bx0) sh.bx1 = sh.bx0;
  const int64_t mz = sh.bz1 - sh.bz0, my = sh.by1 - sh.by0, mx = sh.bx1 - sh.bx0;
  sh.cnt_a = (int64_t)(n0 - mz) * n1 * n2;
  sh.cnt_b = mz * (n1 - my) * n2;
  sh.cnt_c = mz * my * (n2 - mx);
  sh.s00 = a2[0], sh.s22 = a2[2], sh.s01 = a1[0] * a1[1], sh.s02 = a1[0] * a1[2], sh.s12 = a1[1] * a1[2];
  sh.d1z = tab_of(d1[0]), sh.d1y = tab_of(d1[1]), sh.d1x = tab_of(d1[2]), sh.d2z = tab_of(d2z), sh.d2x = tab_of(d2x);
  cudaStream_t st = (cudaStream_t)stream;
  switch (P1) {
    case 1: return launch_hess<T, 1>(d1, d2z, d2x, x, h, a1, a2[0], a2[2], sh, st);
    case 2: return launch_hess<T, 2>(d1, d2z, d2x, x, h, a1, a2[0], a2[2], sh, st);
    default: return launch_hess<T, 3>(d1, d2z, d2x, x, h, a1, a2[0], a2[2], sh, st);
  }
}

}  // namespace
}  // namespace fdx

extern "C" int fdx_hessian3d_f32(const fdx_plan_t d1[3], const fdx_plan_t d2[3], const float* x, float* h, const double a1[3],
                                 const double a2[3], void* stream) {
  return fdx::run_hess<float>(d1, d2, x, h, a1, a2, stream);
}
extern "C" int fdx_hessian3d_f64(const fdx_plan_t d1[3], const fdx_plan_t d2[3], const double* x, double* h, const double a1[3],
                                 const double a2[3], void* stream) {
  return fdx::run_hess<double>(d1, d2, x, h, a1, a2, stream);
}
