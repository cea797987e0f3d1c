// Fused 3-D hessian: replaces the reference's `hessian` on a 3-D scalar field (finitediffx/_src/finite_diff.py:460-529:
// ndim second differences + ndim (ndim - 1) / 2 nested first differences, each a `concatenate`, then ndim + 1 `stack`s)
// by ONE pass over the input: 1 read + 9 writes per point (10 s bytes; the composition of axis passes moves 18 s).
//
// The reference's result table is F[ax1 + ax2] (:499, :510), so for ndim = 3 the five surviving entries are
//   F[0] = D2_0 x,  F[1] = D_1 D_0 x,  F[2] = D_2 D_0 x (it replaces D2_1 x: SURVEY F4),  F[3] = D_2 D_1 x,  F[4] = D2_2 x
// and H[i][j] = F[i + j] (:524-529).  This kernel produces exactly those five fields and stores each of them in every
// slot (i, j) with i + j = key of the (3, 3, n0, n1, n2) result.
//
// ONE launch.  Operators on different axes commute, so the nested differences are tensor products of the 1-D rows
// (main stencil or dense boundary row, finite_diff.py:75-88) -- which is what lets a single march produce them:
//   chunks of the march  a CTA marches a (8 rows x 32 vectors) tile along axis 0 over the planes whose rows of axis 0 are
//                        main rows: plane z is staged in shared memory (cp.async, zero fill outside the grid, a ring of
//                        four planes); stage B writes gx = D_2 x of the tile rows and their y halo to shared memory;
//                        stage C gives every thread D2_2 x, gy = D_1 x and D_1 gx of its vector; three shift-register
//                        queues along z (x -> D2_0 x, gy -> D_0 D_1 x, gx -> D_0 D_2 x) complete planes z - P2 / z - P1.
//                        Main taps sit in the parameter block pre-multiplied by the scales.
//   boundary rows        axis 2: the band columns of gx (all staged rows) and of D2_2 x (tile rows) are evaluated by helper
//                        lanes of warps 4..7 in stage B (warps 0..3 have a second staged row there) from a scaled
//                        coefficient table in shared memory, and patched into the owners' vectors; the last tile of a row
//                        is anchored at the end of the row so that it holds the columns under the right band.
//                        axis 1: a warp whose row is a band row runs the dense row over the staged rows (warp-uniform).
//                        axis 0: the band planes get short CTAs of their own behind the chunks: with Gz = (row zb of D_0) x
//                        -- ONE plane, built in shared memory from the planes under the band -- the mixed entries are
//                        D_1 Gz and D_2 Gz: stages B and C run once on the Gz tile and once on plane zb itself.
// (History, measured at 384^3 fp32 accuracy 4: a separate gather kernel over the boundary shell cost 0.28-0.44 ms next to a
// 0.47 ms march -- nested gathers at ~2000 instructions per point -- and a 32-byte sector written half by the march and half
// by the shell is two partial writes: +50 % on the march.  All rows in the march: 0.49 ms for everything.)
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

struct XBand {
  const double* l;
  const double* r;
  int32_t nl, wl, nr, wr;
  double alpha;
};
constexpr int kXBandMax = 4;  // band columns per side the march handles (reach of the second derivative <= 4)
constexpr int kXBandW = 12;   // taps per band row (first-derivative reach <= 3: wl <= 4 + 7)
constexpr int kYBandMax = 3;  // band rows of axis 1 per side (first-derivative reach <= 3)
constexpr int kXBandW2 = 16;  // 2-D kernel: taps per band row incl. the shift of a right band to a vector boundary

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
  // boundary planes of axis 0: CTAs of their own behind the chunks of the march (block index / tiles >= n_chunks)
  AxisTab zt1, zt2;
  double a1z, a2z;
  int32_t n_chunks;
};

// ---- boundary shell -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  const int sz = valid ? 16 : 0;  // src-size 0: the 16 bytes are zero-filled, nothing is read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// acc += c * x, dst = c * x + src, dst = c * x on 16-byte vectors.  fp32: the packed FFMA2 / FMUL2 (fma.rn.f32x2, sm_100+; two
// independent IEEE fmas, the bits of the scalar form) -- the fp32 march is issue-limited (ncu: 454 instructions per warp and
// plane at 2 CTAs per SM), the y taps and the three z queues are whole-vector operations.
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(unsigned long long p, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p));
}
template <typename T, int V, bool PACK = true>
__device__ __forceinline__ void vfma_to(Vec<T, V>& dst, T c, const Vec<T, V>& x, const Vec<T, V>& src) {
  if constexpr (sizeof(T) == 4 && PACK) {
    const unsigned long long cc = pack2(c, c);
#pragma unroll
    for (int h = 0; h < V / 2; ++h) {
      unsigned long long d;
      asm("fma.rn.f32x2 %0, %1, %2, %3;"
          : "=l"(d)
          : "l"(cc), "l"(pack2(x.v[2 * h], x.v[2 * h + 1])), "l"(pack2(src.v[2 * h], src.v[2 * h + 1])));
      unpack2(d, dst.v[2 * h], dst.v[2 * h + 1]);
    }
  } else {
#pragma unroll
    for (int p = 0; p < V; ++p) dst.v[p] = fma(c, x.v[p], src.v[p]);
  }
}
template <typename T, int V, bool PACK = true>
__device__ __forceinline__ void vfma(Vec<T, V>& acc, T c, const Vec<T, V>& x) {
  vfma_to<T, V, PACK>(acc, c, x, acc);
}
template <typename T, int V, bool PACK = true>
__device__ __forceinline__ void vmul(Vec<T, V>& dst, T c, const Vec<T, V>& x) {
  if constexpr (sizeof(T) == 4 && PACK) {
    const unsigned long long cc = pack2(c, c);
#pragma unroll
    for (int h = 0; h < V / 2; ++h) {
      unsigned long long d;
      asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(cc), "l"(pack2(x.v[2 * h], x.v[2 * h + 1])));
      unpack2(d, dst.v[2 * h], dst.v[2 * h + 1]);
    }
  } else {
#pragma unroll
    for (int p = 0; p < V; ++p) dst.v[p] = c * x.v[p];
  }
}

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

template <typename T, int P1, int MINB>
__global__ void __launch_bounds__(256, MINB) hessian3d_kernel(const __grid_constant__ HessParams<T, P1> prm) {
  using G = HessGeom<T, P1>;
  constexpr int V = G::V, P2 = G::P2, W1 = G::W1, W2 = G::W2, HX = G::HX, HXV = G::HXV, XSV = G::XSV, TX = G::TX;
  using VecT = Vec<T, V>;
  // packed fp32 arithmetic for the y taps and the z queues -- except at reach 2, which sits at 128 registers: packed (aligned
  // register pairs for its 56 queue registers) it spills: 384^3 accuracy 4 0.48 -> 0.53 ms (y taps alone packed: 0.61 ms);
  // reach 3 gains 11 % (0.69 -> 0.61 ms), reach 1 is unchanged
  constexpr bool PQ = P1 != 2;
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
  // (the last tile of a row is anchored at the end of the row: it overlaps its neighbour, which gives up the shared
  // columns, and always holds the columns under the right band rows)
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
  // columns / rows this tile stores (no point is stored twice).  Columns: natural tiles, except that the anchored last tile's
  // share starts at or below the first right band column -- the band columns must all belong to the tile that evaluates
  // them (fp64: with n2 % 64 == 2 and 3+ band columns the tile before the last one overlaps them and would store
  // main-stencil values there); the left band stays with tile 0.  Rows: every tile runs the band rows it overlaps.
  const int xsplit = min((prm.tiles_x - 1) * TX, (prm.n2 - max(prm.xb1.nr, prm.xb2.nr)) / V * V);
  const bool xlast = tx + 1 == prm.tiles_x;
  const int xown_lo = xlast && prm.tiles_x > 1 ? xsplit : x0, xown_hi = xlast ? prm.n2 : min((tx + 1) * TX, xsplit);
  const int yown_hi = ty + 1 == prm.tiles_y ? prm.n1 : min((ty + 1) * G::TY, prm.n1 - G::TY);
  const bool own = yo >= prm.y_lo && yo < min(prm.y_hi, yown_hi) && xo >= max(prm.x_lo, xown_lo) && xo < min(prm.x_hi, xown_hi);
  // a warp whose row is a boundary row of axis 1: band table (0 top, 1 bottom), row in it, staged row under its first column
  const int ybk = yo < prm.yb.nl ? 0 : (yo >= prm.n1 - prm.yb.nr && yo < prm.n1 ? 1 : -1);
  const int ybj = ybk == 0 ? yo : yo - (prm.n1 - prm.yb.nr);
  const int ybw = ybk == 0 ? prm.yb.wl : prm.yb.wr;
  const int ybr = (ybk == 0 ? 0 : prm.n1 - prm.yb.wr) - y0 + P1;
  const int64_t ooff = (int64_t)yo * prm.n2 + xo;

  // ---- stage B: gx = D_2 x on the staged rows (tile rows and their y halo) -> gs; edge tiles: the band columns of it and
  // of D2_2 x by the helper lanes
  auto stage_b = [&](const T* X) {
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
  };
  // ---- stage C: the thread's own vector of the staged plane: its value, D2_2 x, gy = D_1 x, gx = D_2 x and D_1 gx
  auto stage_c = [&](const T* X, VecT& xc, VecT& gyv, VecT& gxc, VecT& d22, VecT& d12) {
      T w[G::WV * V];
      const int r = wrp + P1;
#pragma unroll
      for (int u = 0; u < G::WV; ++u) {
        const VecT t = *reinterpret_cast<const VecT*>(&X[(r * XSV + lane + u) * V]);
#pragma unroll
        for (int p = 0; p < V; ++p) w[u * V + p] = t.v[p];
      }
#pragma unroll
      for (int p = 0; p < V; ++p) xc.v[p] = w[HX + p];
#pragma unroll
      for (int p = 0; p < V; ++p) {
        T a = (T)0;
#pragma unroll
        for (int k = 0; k < W2; ++k) a = fma(prm.c2x[k], w[HX - P2 + p + k], a);
        d22.v[p] = a;
        gyv.v[p] = (T)0, d12.v[p] = (T)0;
      }
      if (bm2 != 0) {  // band columns of D2_2 x: from the helper lanes
#pragma unroll
        for (int p = 0; p < V; ++p) {
          const int col = x0 + lane * V + p;
          if ((bm2 >> p) & 1) d22.v[p] = col < prm.xb2.nl && edge_l ? d22b[wrp][col] : d22b[wrp][kXBandMax + col - (prm.n2 - prm.xb2.nr)];
        }
      }
      if (ybk >= 0) {  // (warp-uniform) dense boundary row of axis 1 over the staged rows; zero taps skipped
        gxc = *reinterpret_cast<const VecT*>(&gs[(wrp + P1) * TX + lane * V]);
        for (int c = 0; c < ybw; ++c) {
          const T k = ybt[ybk][ybj][c];
          if (k == (T)0) continue;
          const VecT gj = *reinterpret_cast<const VecT*>(&gs[(ybr + c) * TX + lane * V]);
          const VecT xj = *reinterpret_cast<const VecT*>(&X[((ybr + c) * XSV + lane + HXV) * V]);
          vfma<T, V, PQ>(gyv, k, xj);
          vfma<T, V, PQ>(d12, k, gj);
        }
      } else
#pragma unroll
      for (int j = 0; j < W1; ++j) {
        const VecT gj = *reinterpret_cast<const VecT*>(&gs[(wrp + j) * TX + lane * V]);
        VecT xj;
        if (j == P1) {
          xj = xc, gxc = gj;
        } else {
          xj = *reinterpret_cast<const VecT*>(&X[((wrp + j) * XSV + lane + HXV) * V]);
        }
        vfma<T, V, PQ>(gyv, prm.c1y[j], xj);
        vfma<T, V, PQ>(d12, prm.c1y[j], gj);
      }
  };
  if (b >= prm.n_chunks) {
    // ---- a boundary plane of axis 0 (the dense rows of both plans there): no queue.  With Gz = (row zb of D_0) x, a single
    // plane, the mixed entries are D_1 Gz and D_2 Gz -- the in-plane stages run on the staged Gz tile and on plane zb itself
    const int zi = b - prm.n_chunks;
    const int zb = zi < prm.z_lo ? zi : prm.z_hi + (zi - prm.z_lo);
    stage(zb, 1);
    cp_async_commit();
    const RowOp r1 = row_of(prm.zt1, zb), r2 = row_of(prm.zt2, zb);
#pragma unroll
    for (int s = 0; s < G::NLD; ++s)
      if (soff[s] >= 0) {
        VecT acc;
#pragma unroll
        for (int p = 0; p < V; ++p) acc.v[p] = (T)0;
        if (ok[s])
          for (int c = 0; c < r1.cnt; ++c) {
            const T k = (T)(__ldg(r1.c + c) * prm.a1z);
            if (k == (T)0) continue;
            const VecT v = ld_cached<T, V>(prm.x + (int64_t)(r1.s + c) * plane + goff[s]);
#pragma unroll
            for (int p = 0; p < V; ++p) acc.v[p] = fma(k, v.v[p], acc.v[p]);
          }
        *reinterpret_cast<VecT*>(&xs[0][soff[s]]) = acc;
      }
    VecT d00;
#pragma unroll
    for (int p = 0; p < V; ++p) d00.v[p] = (T)0;
    if (own)
      for (int c = 0; c < r2.cnt; ++c) {
        const T k = (T)(__ldg(r2.c + c) * prm.a2z);
        if (k == (T)0) continue;
        const VecT v = ld_cached<T, V>(prm.x + (int64_t)(r2.s + c) * plane + ooff);
#pragma unroll
        for (int p = 0; p < V; ++p) d00.v[p] = fma(k, v.v[p], d00.v[p]);
      }
    cp_async_wait<0>();
    __syncthreads();
    VecT xc, gyv, gxc;
    VecT d22, d12, d01, d02;
    stage_b(xs[0]);
    __syncthreads();
    stage_c(xs[0], xc, gyv, gxc, d22, d12);
    d01 = gyv, d02 = gxc;
    __syncthreads();
    stage_b(xs[1]);
    __syncthreads();
    stage_c(xs[1], xc, gyv, gxc, d22, d12);
    if (own) {
      T* o = prm.h + (int64_t)zb * plane + ooff;
      st_cs<T, V>(o, d00, plain);
      st_cs<T, V>(o + 1 * N, d01, plain);
      st_cs<T, V>(o + 3 * N, d01, plain);
      st_cs<T, V>(o + 2 * N, d02, plain);
      st_cs<T, V>(o + 4 * N, d02, plain);
      st_cs<T, V>(o + 6 * N, d02, plain);
      st_cs<T, V>(o + 5 * N, d12, plain);
      st_cs<T, V>(o + 7 * N, d12, plain);
      st_cs<T, V>(o + 8 * N, d22, plain);
    }
    return;
  }

  VecT q0[2 * P2], q1[2 * P1], q2[2 * P1];
#pragma unroll
  for (int j = 0; j < 2 * P2; ++j)
#pragma unroll
    for (int p = 0; p < V; ++p) q0[j].v[p] = (T)0;
#pragma unroll
  for (int j = 0; j < 2 * P1; ++j)
#pragma unroll
    for (int p = 0; p < V; ++p) q1[j].v[p] = q2[j].v[p] = (T)0;

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

    stage_b(X);
    __syncthreads();

    // the thread's own vector of plane z
    const bool inplane = z >= zc0 && z < zc1;
    VecT xc, gyv, gxc;
    {
      VecT d22, d12;
      stage_c(X, xc, gyv, gxc, d22, d12);
      if (inplane && own) {
        T* o = prm.h + (int64_t)z * plane + ooff;
        st_cs<T, V>(o + 8 * N, d22, plain);  // key 4: (2,2)
        st_cs<T, V>(o + 5 * N, d12, plain);  // key 3: (1,2), (2,1)
        st_cs<T, V>(o + 7 * N, d12, plain);
      }
    }
    // ---- z queues: plane z completes D2_0 x of plane z - P2 and the mixed entries of plane z - P1
    VecT d00, d01, d02;
    vfma_to<T, V, PQ>(d00, prm.c2z[2 * P2], xc, q0[0]);
    vfma_to<T, V, PQ>(d01, prm.c1z[2 * P1], gyv, q1[0]);
    vfma_to<T, V, PQ>(d02, prm.c1z[2 * P1], gxc, q2[0]);
#pragma unroll
    for (int j = 1; j < 2 * P2; ++j) vfma_to<T, V, PQ>(q0[j - 1], prm.c2z[2 * P2 - j], xc, q0[j]);
#pragma unroll
    for (int j = 1; j < 2 * P1; ++j) {
      vfma_to<T, V, PQ>(q1[j - 1], prm.c1z[2 * P1 - j], gyv, q1[j]);
      vfma_to<T, V, PQ>(q2[j - 1], prm.c1z[2 * P1 - j], gxc, q2[j]);
    }
    vmul<T, V, PQ>(q0[2 * P2 - 1], prm.c2z[0], xc);
    vmul<T, V, PQ>(q1[2 * P1 - 1], prm.c1z[0], gyv);
    vmul<T, V, PQ>(q2[2 * P1 - 1], prm.c1z[0], gxc);
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
                double a2z, double a2x, int bz0, int bz1, cudaStream_t st) {
  using G = HessGeom<T, P1>;
  HessParams<T, P1> prm;
  prm.x = x, prm.h = h;
  prm.n0 = d1[0]->n, prm.n1 = d1[1]->n, prm.n2 = d1[2]->n;
  prm.z_lo = bz0, prm.z_hi = bz1;
  prm.plain_stores = env_int("FDX_HESS_PLAIN_ST", 0);
  prm.xb1 = XBand{d1[2]->band_l, d1[2]->band_r, d1[2]->nl, d1[2]->wl, d1[2]->nr, d1[2]->wr, a1[2]};
  prm.xb2 = XBand{d2x->band_l, d2x->band_r, d2x->nl, d2x->wl, d2x->nr, d2x->wr, a2x};
  prm.yb = XBand{d1[1]->band_l, d1[1]->band_r, d1[1]->nl, d1[1]->wl, d1[1]->nr, d1[1]->wr, a1[1]};
  prm.zt1 = tab_of(d1[0]), prm.zt2 = tab_of(d2z), prm.a1z = a1[0], prm.a2z = a2z;
  fill_taps<T, P1 + 1>(prm.c2z, d2z, a2z);
  fill_taps<T, P1 + 1>(prm.c2x, d2x, a2x);
  fill_taps<T, P1>(prm.c1z, d1[0], a1[0]);
  fill_taps<T, P1>(prm.c1y, d1[1], a1[1]);
  fill_taps<T, P1>(prm.c1x, d1[2], a1[2]);
  prm.y_lo = 0, prm.y_hi = prm.n1, prm.x_lo = 0, prm.x_hi = prm.n2;
  prm.tiles_x = (prm.n2 + G::TX - 1) / G::TX, prm.tiles_y = (prm.n1 + G::TY - 1) / G::TY;
  prm.zc = 1, prm.n_chunks = 0;
  const int nz = bz1 - bz0;
  const int part = env_int("FDX_HESS_PART", 0);  // timing tools: 1 = the chunks of the march alone, 2 = the z-band planes alone
  const int64_t tiles = (int64_t)prm.tiles_x * prm.tiles_y;
  if (nz > 0 && part != 2) {
    // z chunks of ~32 planes (a chunk re-reads 2 P2 planes; 384^3: 32 beats 48 / 64 / 96 by 1-6 %), shorter when that
    // leaves fewer than ~3 CTAs per SM
    int zc = env_int("FDX_HESS_ZC", 0);
    if (zc <= 0) {
      const int64_t want = (int64_t)env_int("FDX_HESS_CPS", 3) * num_sms();
      const int64_t chunks = std::max<int64_t>((nz + 31) / 32, (want + tiles - 1) / tiles);
      zc = std::max((int)((nz + chunks - 1) / chunks), 16);
    }
    zc = std::min(zc, nz);
    prm.zc = zc;
    prm.n_chunks = (nz + zc - 1) / zc;
  }
  // the boundary planes of axis 0: one short CTA per tile and plane, behind the chunks (they fill the tail of the launch)
  const int64_t zband_planes = part == 1 ? 0 : prm.n0 - nz;
  const int64_t blocks = tiles * (prm.n_chunks + zband_planes);
  if (blocks > 0x7fffffffLL) return FDX_EUNSUPPORTED;
  if (blocks == 0) return FDX_OK;
  bool three = false;
  if constexpr (P1 <= 2) three = env_int("FDX_HESS_MINB", 2) == 3;
  if constexpr (P1 <= 2) {
    if (three) hessian3d_kernel<T, P1, 3><<<(unsigned)blocks, G::NT, 0, st>>>(prm);
  }
  if (!three) hessian3d_kernel<T, P1, 2><<<(unsigned)blocks, G::NT, 0, st>>>(prm);
  count_launch("fused_hessian3d");
  FDX_LAUNCH_CHECK();
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

  // main rows of axis 0: the chunks of the march; the band planes on either side get CTAs of their own.  The band rows of
  // axes 1 and 2 are evaluated inside the march (a warp per band row, helper lanes per band column): a separate pass over
  // the boundary shell cost more than the march itself (nested gathers), and a 32-byte sector written half by each pass is
  // two partial writes
  int bz0 = std::max(d1[0]->nl, d2z->nl), bz1 = n0 - std::max(d1[0]->nr, d2z->nr);
  if (bz1 < bz0) bz1 = bz0;
  if (d1[1]->nl > kYBandMax || d1[1]->nr > kYBandMax || d1[1]->wl > kXBandW || d1[1]->wr > kXBandW) return FDX_EUNSUPPORTED;
  if (std::max(d1[2]->nl, d2x->nl) > kXBandMax || std::max(d1[2]->nr, d2x->nr) > kXBandMax) return FDX_EUNSUPPORTED;
  if (std::max(std::max(d1[2]->wl, d2x->wl), std::max(d1[2]->wr, d2x->wr)) > kXBandW) return FDX_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  switch (P1) {
    case 1: return launch_hess<T, 1>(d1, d2z, d2x, x, h, a1, a2[0], a2[2], bz0, bz1, st);
    case 2: return launch_hess<T, 2>(d1, d2z, d2x, x, h, a1, a2[0], a2[2], bz0, bz1, st);
    default: return launch_hess<T, 3>(d1, d2z, d2x, x, h, a1, a2[0], a2[2], bz0, bz1, st);
  }
}


// ---- 2-D hessian ------------------------------------------------------------------------------------------------------
// hessian of a 2-D scalar field (finite_diff.py:460-529 with ndim = 2: F[0] = D2_0 x, F[1] = D_1 D_0 x, F[2] = D2_1 x, H =
// [[F0, F1], [F1, F2]]) in one pass: 1 read + 4 writes per point (the composition of axis passes: 4 reads + 5 writes).
// A WARP marches a strip of 32 vectors down axis 0; no shared data between warps, no block barrier:
//   row y: the lane's window of the row (its vector and the halo vectors, L1-shared with the neighbour lanes) gives
//          gx = D_1 x and D2_1 x of its vector; two shift-register queues along axis 0 (x -> D2_0 x, reach P2; gx -> D_0 gx,
//          reach P1) complete rows y - P2 / y - P1;
//   boundary columns of axis 1 (edge strips; the last strip is anchored at the end of the row): the columns under the band
//          are broadcast from their owner lanes with shuffles, lane j evaluates band row j from a scaled table in shared
//          memory and hands it to the owner of column j;
//   boundary rows of axis 0: work items of their own -- the band row collapses the rows under it into ONE virtual row
//          G = (row yb of D_0) x, and D_1 D_0 x = D_1 G: the in-row code runs on G's window and on row yb's.
template <typename T, int P1>
struct Hess2Params {
  const T* x;
  T* h;  // (2, 2, n0, n1)
  int32_t n0, n1;
  int32_t y_lo, y_hi;  // main rows of axis 0
  int32_t rc, n_chunks, strips;
  int32_t pf;  // rows ahead of the march whose lines are pulled into L2 (prefetch.global.L2: no registers held)
  int64_t march_items, items;
  T c2y[2 * P1 + 3], c2x[2 * P1 + 3], c1y[2 * P1 + 1], c1x[2 * P1 + 1];
  XBand xb1, xb2;
  AxisTab yt1, yt2;
  double a1y, a2y;
};

template <typename T, int P1>
struct Hess2Geom {
  static constexpr int V = 16 / (int)sizeof(T), P2 = P1 + 1, W1 = 2 * P1 + 1, W2 = 2 * P2 + 1;
  static constexpr int HXV = (P2 + V - 1) / V, HX = HXV * V, WV = 1 + 2 * HXV, TX = 32 * V;
};
// window of a row: the lane's vector and HXV vectors on either side (zero outside the row / the grid)
template <typename T, int P1>
__device__ __forceinline__ void h2_load_win(const Hess2Params<T, P1>& prm, int xo, int y, Vec<T, 16 / (int)sizeof(T)> (&w)[Hess2Geom<T, P1>::WV]) {
  using G = Hess2Geom<T, P1>;
  constexpr int V = G::V, HXV = G::HXV, WV = G::WV;

    const bool yin = y >= 0 && y < prm.n0;
    const T* row = prm.x + (yin ? (int64_t)y * prm.n1 : 0);
#pragma unroll
    for (int u = 0; u < WV; ++u) {
      const int col = xo + (u - HXV) * V;
      if (yin && col >= 0 && col < prm.n1) {
        w[u] = ld_cached<T, V>(row + col);
      } else {
#pragma unroll
        for (int p = 0; p < V; ++p) w[u].v[p] = (T)0;
      }
    }
}
// in-row work on a window: gx = D_1 (row), xx = D2_1 (row) of the lane's vector, boundary columns included
template <typename T, int P1>
__device__ __forceinline__ void h2_row_ops(const Hess2Params<T, P1>& prm, const T (*bt)[kXBandMax][kXBandW2], int x0, int lane, bool edge_l,
                                           bool edge_r, const Vec<T, 16 / (int)sizeof(T)> (&w)[Hess2Geom<T, P1>::WV],
                                           Vec<T, 16 / (int)sizeof(T)>& gx, Vec<T, 16 / (int)sizeof(T)>& xx) {
  using G = Hess2Geom<T, P1>;
  constexpr int V = G::V, P2 = G::P2, W1 = G::W1, W2 = G::W2, HXV = G::HXV, HX = G::HX, WV = G::WV, TX = G::TX;
  using VecT = Vec<T, V>;

    T f[WV * V];
#pragma unroll
    for (int u = 0; u < WV; ++u)
#pragma unroll
      for (int p = 0; p < V; ++p) f[u * V + p] = w[u].v[p];
#pragma unroll
    for (int p = 0; p < V; ++p) {
      T a = (T)0, b = (T)0;
#pragma unroll
      for (int k = 0; k < W1; ++k) a = fma(prm.c1x[k], f[HX - P1 + p + k], a);
#pragma unroll
      for (int k = 0; k < W2; ++k) b = fma(prm.c2x[k], f[HX - P2 + p + k], b);
      gx.v[p] = a, xx.v[p] = b;
    }
    if (edge_l | edge_r) {  // (warp-uniform)
      const VecT& c = w[HXV];
#pragma unroll
      for (int side = 0; side < 2; ++side) {
        if (side == 0 ? !edge_l : !edge_r) continue;
        const int nb1 = side == 0 ? prm.xb1.nl : prm.xb1.nr, nb2 = side == 0 ? prm.xb2.nl : prm.xb2.nr;
        const int w1 = side == 0 ? prm.xb1.wl : prm.xb1.wr, w2 = side == 0 ? prm.xb2.wl : prm.xb2.wr;
        // the table of a right band starts at the vector boundary at or below its first column (leading zero taps), so
        // that tap cc always sits in element cc % V of its owner lane: every register index below is static
        const int f1 = side == 0 ? 0 : (prm.n1 - w1) / V * V, f2 = side == 0 ? 0 : (prm.n1 - w2) / V * V;
        const int e1 = side == 0 ? w1 : prm.n1 - f1, e2 = side == 0 ? w2 : prm.n1 - f2;  // taps of the shifted tables
        const int emax = max(e1, e2);
        const int s1 = (f1 - x0) / V, s2 = (f2 - x0) / V;  // lanes of the first vectors under the bands
        // lane j evaluates band row j of both plans; the columns under the bands come from their owner lanes
        T a1 = (T)0, a2 = (T)0;
        const int jl = min(lane, kXBandMax - 1);
#pragma unroll
        for (int cc = 0; cc < kXBandW2; ++cc) {
          if (cc >= emax) break;
          const T v1 = __shfl_sync(0xffffffffu, c.v[cc % V], min(s1 + cc / V, 31));
          const T v2 = __shfl_sync(0xffffffffu, c.v[cc % V], min(s2 + cc / V, 31));
          const T k1 = bt[side][jl][cc], k2 = bt[2 + side][jl][cc];
          a1 = fma(k1, k1 != (T)0 ? v1 : (T)0, a1);
          a2 = fma(k2, k2 != (T)0 ? v2 : (T)0, a2);
        }
        // the owner of a band column fetches its value from the lane that evaluated its band row
        const int b1 = side == 0 ? 0 : prm.n1 - nb1, b2 = side == 0 ? 0 : prm.n1 - nb2;  // first band columns
#pragma unroll
        for (int p = 0; p < V; ++p) {
          const int col = x0 + lane * V + p;
          const int j1 = col - b1, j2 = col - b2;
          const T r1 = __shfl_sync(0xffffffffu, a1, max(0, min(j1, 31))), r2 = __shfl_sync(0xffffffffu, a2, max(0, min(j2, 31)));
          if (j1 >= 0 && j1 < nb1 && col < prm.n1) gx.v[p] = r1;
          if (j2 >= 0 && j2 < nb2 && col < prm.n1) xx.v[p] = r2;
        }
      }
    }
}

template <typename T, int P1, int MINB>
__global__ void __launch_bounds__(128, MINB) hessian2d_kernel(const __grid_constant__ Hess2Params<T, P1> prm) {
  constexpr int V = 16 / (int)sizeof(T), P2 = P1 + 1, W1 = 2 * P1 + 1, W2 = 2 * P2 + 1;
  constexpr int HXV = (P2 + V - 1) / V, HX = HXV * V, WV = 1 + 2 * HXV, TX = 32 * V;
  using VecT = Vec<T, V>;
  __shared__ T bt[4][kXBandMax][kXBandW2];  // band rows of axis 1: {D_1 left, D_1 right, D2_1 left, D2_1 right}, scaled
  const int tid = threadIdx.x, lane = tid & 31;
  for (int e = tid; e < 4 * kXBandMax * kXBandW2; e += 128) {
    const int t = e / (kXBandMax * kXBandW2), j = (e / kXBandW2) % kXBandMax, cc = e % kXBandW2;
    const XBand& xb = t < 2 ? prm.xb1 : prm.xb2;
    const bool left = (t & 1) == 0;
    const int rows = left ? xb.nl : xb.nr, w = left ? xb.wl : xb.wr;
    const int off = left ? 0 : (prm.n1 - w) % V;  // (right bands: shifted to the vector boundary below their first column)
    const double* tab = left ? xb.l : xb.r;
    const int c = cc - off;
    bt[t][j][cc] = (j < rows && c >= 0 && c < w) ? (T)(tab[j * w + c] * xb.alpha) : (T)0;
  }
  __syncthreads();
  const int64_t item = (int64_t)blockIdx.x * 4 + (tid >> 5);
  if (item >= prm.items) return;
  const bool band_item = item >= prm.march_items;
  const int64_t it = band_item ? item - prm.march_items : item;
  const int strip = (int)(it % prm.strips), rest = (int)(it / prm.strips);
  const int x0 = min(strip * TX, max(0, prm.n1 - TX));
  const int xo = x0 + lane * V;
  // (natural strips, except that the anchored last strip's share starts at or below the first right band column: the band
  // columns belong to the strip that evaluates them; the left band stays with strip 0)
  const int xsplit = min((prm.strips - 1) * TX, (prm.n1 - max(prm.xb1.nr, prm.xb2.nr)) / V * V);
  const bool xlast = strip + 1 == prm.strips;
  const bool own = xo >= (xlast && prm.strips > 1 ? xsplit : x0) && xo < (xlast ? prm.n1 : min((strip + 1) * TX, xsplit));
  const bool edge_l = x0 == 0 && (prm.xb1.nl > 0 || prm.xb2.nl > 0), edge_r = x0 + TX >= prm.n1 && (prm.xb1.nr > 0 || prm.xb2.nr > 0);
  const int64_t N = (int64_t)prm.n0 * prm.n1;

  const int64_t ooff = xo;

  if (band_item) {
    // ---- a boundary row of axis 0
    const int yi = rest;
    const int yb = yi < prm.y_lo ? yi : prm.y_hi + (yi - prm.y_lo);
    const RowOp r1 = row_of(prm.yt1, yb), r2 = row_of(prm.yt2, yb);
    VecT g[WV], w[WV], yy;
#pragma unroll
    for (int u = 0; u < WV; ++u)
#pragma unroll
      for (int p = 0; p < V; ++p) g[u].v[p] = (T)0;
#pragma unroll
    for (int p = 0; p < V; ++p) yy.v[p] = (T)0;
    for (int c = 0; c < r1.cnt; ++c) {
      const T k = (T)(__ldg(r1.c + c) * prm.a1y);
      if (k == (T)0) continue;
      h2_load_win<T, P1>(prm, xo, r1.s + c, w);
#pragma unroll
      for (int u = 0; u < WV; ++u)
#pragma unroll
        for (int p = 0; p < V; ++p) g[u].v[p] = fma(k, w[u].v[p], g[u].v[p]);
    }
    for (int c = 0; c < r2.cnt; ++c) {
      const T k = (T)(__ldg(r2.c + c) * prm.a2y);
      if (k == (T)0 || !own) continue;
      const VecT v = ld_cached<T, V>(prm.x + (int64_t)(r2.s + c) * prm.n1 + xo);
#pragma unroll
      for (int p = 0; p < V; ++p) yy.v[p] = fma(k, v.v[p], yy.v[p]);
    }
    VecT mixed, xx, dummy;
    h2_row_ops<T, P1>(prm, bt, x0, lane, edge_l, edge_r, g, mixed, dummy);
    h2_load_win<T, P1>(prm, xo, yb, w);
    h2_row_ops<T, P1>(prm, bt, x0, lane, edge_l, edge_r, w, dummy, xx);
    if (own) {
      T* o = prm.h + (int64_t)yb * prm.n1 + ooff;
      st_cs<T, V>(o, yy);
      st_cs<T, V>(o + N, mixed);
      st_cs<T, V>(o + 2 * N, mixed);
      st_cs<T, V>(o + 3 * N, xx);
    }
    return;
  }

  // ---- a chunk of main rows
  const int yc0 = prm.y_lo + rest * prm.rc, yc1 = min(yc0 + prm.rc, prm.y_hi);
  VecT q0[2 * P2], q1[2 * P1];
#pragma unroll
  for (int j = 0; j < 2 * P2; ++j)
#pragma unroll
    for (int p = 0; p < V; ++p) q0[j].v[p] = (T)0;
#pragma unroll
  for (int j = 0; j < 2 * P1; ++j)
#pragma unroll
    for (int p = 0; p < V; ++p) q1[j].v[p] = (T)0;
  VecT w[WV], wn[WV];
  h2_load_win<T, P1>(prm, xo, yc0 - P2, w);
  for (int y = yc0 - P2; y <= yc1 - 1 + P2; ++y) {
    if (y < yc1 - 1 + P2) h2_load_win<T, P1>(prm, xo, y + 1, wn);  // (next row in flight behind this row's arithmetic)
    if (prm.pf > 0 && y + prm.pf < prm.n0 && y + prm.pf >= 0 && own)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(prm.x + (int64_t)(y + prm.pf) * prm.n1 + xo));
    VecT gx, xx;
    h2_row_ops<T, P1>(prm, bt, x0, lane, edge_l, edge_r, w, gx, xx);
    const VecT& xc = w[HXV];
    if (y >= yc0 && y < yc1 && own) st_cs<T, V>(prm.h + 3 * N + (int64_t)y * prm.n1 + ooff, xx);
    VecT yy, mixed;
    vfma_to<T, V>(yy, prm.c2y[2 * P2], xc, q0[0]);
    vfma_to<T, V>(mixed, prm.c1y[2 * P1], gx, q1[0]);
#pragma unroll
    for (int j = 1; j < 2 * P2; ++j) vfma_to<T, V>(q0[j - 1], prm.c2y[2 * P2 - j], xc, q0[j]);
#pragma unroll
    for (int j = 1; j < 2 * P1; ++j) vfma_to<T, V>(q1[j - 1], prm.c1y[2 * P1 - j], gx, q1[j]);
    vmul<T, V>(q0[2 * P2 - 1], prm.c2y[0], xc);
    vmul<T, V>(q1[2 * P1 - 1], prm.c1y[0], gx);
    const int yo2 = y - P2, yo1 = y - P1;
    if (yo2 >= yc0 && yo2 < yc1 && own) st_cs<T, V>(prm.h + (int64_t)yo2 * prm.n1 + ooff, yy);
    if (yo1 >= yc0 && yo1 < yc1 && own) {
      T* o = prm.h + (int64_t)yo1 * prm.n1 + ooff;
      st_cs<T, V>(o + N, mixed);
      st_cs<T, V>(o + 2 * N, mixed);
    }
#pragma unroll
    for (int u = 0; u < WV; ++u) w[u] = wn[u];
  }
}

template <typename T, int P1>
int launch_hess2(const fdx_plan_s* const d1[2], const fdx_plan_s* const d2[2], const T* x, T* h, const double a1[2], const double a2[2],
                 int by0, int by1, cudaStream_t st) {
  constexpr int V = 16 / (int)sizeof(T), TX = 32 * V;
  Hess2Params<T, P1> prm;
  prm.x = x, prm.h = h, prm.n0 = d1[0]->n, prm.n1 = d1[1]->n;
  prm.y_lo = by0, prm.y_hi = by1;
  fill_taps<T, P1 + 1>(prm.c2y, d2[0], a2[0]);
  fill_taps<T, P1 + 1>(prm.c2x, d2[1], a2[1]);
  fill_taps<T, P1>(prm.c1y, d1[0], a1[0]);
  fill_taps<T, P1>(prm.c1x, d1[1], a1[1]);
  prm.xb1 = XBand{d1[1]->band_l, d1[1]->band_r, d1[1]->nl, d1[1]->wl, d1[1]->nr, d1[1]->wr, a1[1]};
  prm.xb2 = XBand{d2[1]->band_l, d2[1]->band_r, d2[1]->nl, d2[1]->wl, d2[1]->nr, d2[1]->wr, a2[1]};
  prm.yt1 = tab_of(d1[0]), prm.yt2 = tab_of(d2[0]), prm.a1y = a1[0], prm.a2y = a2[0];
  prm.strips = (prm.n1 + TX - 1) / TX;
  const int ny = by1 - by0;
  prm.rc = 1, prm.n_chunks = 0;
  // (8192^2 accuracy 4, rows ahead 0 / 2 / 4 / 8 / 16: fp32 0.416 / 0.332 / 0.318 / 0.340 / 0.382 ms, fp64 0.645 / 0.555 / 0.598 / 0.623 / 0.634)
  // fp64, 2 against 4 rows ahead: accuracy 2 0.597 / 0.565, accuracy 4 0.555 / 0.598, accuracy 6 0.699 / 0.658 ms: 4 for both types
  prm.pf = env_int("FDX_HESS2D_PF", 4);
  if (ny > 0) {
    // rows per chunk: a chunk re-reads 2 P2 rows; ~32 warps per SM over the launch, at least 32 rows per chunk
    const int64_t want = (int64_t)env_int("FDX_HESS2D_WPS", 32) * num_sms();
    const int64_t chunks = std::max<int64_t>(1, std::min<int64_t>((ny + 31) / 32, (want + prm.strips - 1) / prm.strips));
    prm.rc = (int)((ny + chunks - 1) / chunks);
    prm.n_chunks = (ny + prm.rc - 1) / prm.rc;
  }
  prm.march_items = (int64_t)prm.strips * prm.n_chunks;
  prm.items = prm.march_items + (int64_t)prm.strips * (prm.n0 - ny);
  const int64_t blocks = (prm.items + 3) / 4;
  if (blocks > 0x7fffffffLL) return FDX_EUNSUPPORTED;
  if (blocks == 0) return FDX_OK;
  // 4 CTAs (16 warps) per SM at 128 registers: 8192^2 accuracy 4 fp32 0.347 ms against 0.406 (3 CTAs, 150 registers) and
  // 0.506 (6 CTAs, 80 registers, spills); fp64 0.616 / 0.727 / 1.78
  // (fp64 at reach 3 spills at 128 registers and then loses to the composition: 3 CTAs per SM there)
  constexpr int MB = (sizeof(T) == 8 && P1 == 3) ? 3 : 4;
  hessian2d_kernel<T, P1, MB><<<(unsigned)blocks, 128, 0, st>>>(prm);
  count_launch("fused_hessian2d");
  FDX_LAUNCH_CHECK();
  return FDX_OK;
}

template <typename T>
int run_hess2(const fdx_plan_t d1_in[2], const fdx_plan_t d2_in[2], const T* x, T* h, const double a1[2], const double a2[2], void* stream) {
  constexpr int V = 16 / (int)sizeof(T);
  if (!d1_in || !d2_in || !a1 || !a2 || !x || !h) return FDX_EINVAL;
  if (!d1_in[0] || !d1_in[1] || !d2_in[0] || !d2_in[1]) return FDX_EINVAL;
  if (env_int("FDX_HESS_DISABLE", 0)) return FDX_EUNSUPPORTED;
  const fdx_plan_s* d1[2] = {d1_in[0], d1_in[1]};
  const fdx_plan_s* d2[2] = {d2_in[0], d2_in[1]};
  if (d2[0]->n != d1[0]->n || d2[1]->n != d1[1]->n) return FDX_EINVAL;
  const int n0 = d1[0]->n, n1 = d1[1]->n;
  if (n0 == 0 || n1 == 0) return FDX_OK;
  const fdx_plan_s* all[4] = {d1[0], d1[1], d2[0], d2[1]};
  int P1 = 1;
  for (int k = 0; k < 4; ++k) {
    const fdx_plan_s* p = all[k];
    if (p->nl + p->nr > p->n) return FDX_EUNSUPPORTED;
    if (p->nl < -p->lo || p->nr < p->hi) return FDX_EUNSUPPORTED;
    if (p->wl > p->n || p->wr > p->n) return FDX_EUNSUPPORTED;
    const int reach = std::max(std::abs(p->lo), std::abs(p->hi));
    P1 = std::max(P1, k < 2 ? reach : reach - 1);
  }
  if (P1 > 3) return FDX_EUNSUPPORTED;
  if (n1 % V != 0 || ((uintptr_t)x & 15) != 0 || ((uintptr_t)h & 15) != 0) return FDX_EUNSUPPORTED;
  if (((int64_t)n0 * n1) % V != 0) return FDX_EUNSUPPORTED;
  if (std::max(d1[1]->nl, d2[1]->nl) > kXBandMax || std::max(d1[1]->nr, d2[1]->nr) > kXBandMax) return FDX_EUNSUPPORTED;
  if (std::max(std::max(d1[1]->wl, d2[1]->wl), std::max(d1[1]->wr, d2[1]->wr)) > kXBandW) return FDX_EUNSUPPORTED;
  if (std::max(std::max(d1[1]->wl, d2[1]->wl), std::max(d1[1]->wr, d2[1]->wr)) > 32 * V) return FDX_EUNSUPPORTED;
  int by0 = std::max(d1[0]->nl, d2[0]->nl), by1 = n0 - std::max(d1[0]->nr, d2[0]->nr);
  if (by1 < by0) by1 = by0;
  cudaStream_t st = (cudaStream_t)stream;
  switch (P1) {
    case 1: return launch_hess2<T, 1>(d1, d2, x, h, a1, a2, by0, by1, st);
    case 2: return launch_hess2<T, 2>(d1, d2, x, h, a1, a2, by0, by1, st);
    default: return launch_hess2<T, 3>(d1, d2, x, h, a1, a2, by0, by1, st);
  }
}

}  // namespace
}  // namespace fdx

extern "C" int fdx_hessian2d_f32(const fdx_plan_t d1[2], const fdx_plan_t d2[2], const float* x, float* h, const double a1[2],
                                 const double a2[2], void* stream) {
  return fdx::run_hess2<float>(d1, d2, x, h, a1, a2, stream);
}
extern "C" int fdx_hessian2d_f64(const fdx_plan_t d1[2], const fdx_plan_t d2[2], const double* x, double* h, const double a1[2],
                                 const double a2[2], void* stream) {
  return fdx::run_hess2<double>(d1, d2, x, h, a1, a2, stream);
}
extern "C" int fdx_hessian3d_f32(const fdx_plan_t d1[3], const fdx_plan_t d2[3], const float* x, float* h, const double a1[3],
                                 const double a2[3], void* stream) {
  return fdx::run_hess<float>(d1, d2, x, h, a1, a2, stream);
}
extern "C" int fdx_hessian3d_f64(const fdx_plan_t d1[3], const fdx_plan_t d2[3], const double* x, double* h, const double a1[3],
                                 const double a2[3], void* stream) {
  return fdx::run_hess<double>(d1, d2, x, h, a1, a2, stream);
}
