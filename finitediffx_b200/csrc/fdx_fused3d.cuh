// Fused single-pass 3-D operators: laplacian, divergence, gradient, curl  (include/fdx_b200.h).
//
// The reference computes each of these as 3-6 separate `difference` passes plus adds/stack
// (finitediffx/_src/finite_diff.py:338-351, 443-453, 565-575, 616-668).  Here one kernel reads
// every input plane once and writes every output once.
//
// Design (sm_100a):
//   * grid of (y,x) tiles x z-chunks; a CTA marches its tile along z (axis 0, the slowest axis);
//   * a producer warp streams one (TY+2P) x (TX+2HX) halo box per input field and plane into a
//     multi-stage shared-memory ring with TMA (cp.async.bulk.tensor, mbarrier full/empty pairs);
//     out-of-range halo elements are zero-filled by the TMA unit, so the main path has no bounds
//     checks;
//   * consumer threads own RY rows x 16 bytes (4 fp32 / 2 fp64) of the tile.  In-plane taps come
//     from shared memory with 128-bit loads; along z each incoming plane is scattered into a
//     register queue of 2P+1 partial sums per point, so a plane lives in shared memory for one
//     step only and every output is written exactly once with a 128-bit store;
//   * rows/columns/planes next to a physical boundary use the plan's dense band taps (the
//     reference's one-sided stencils or their transposes): those few points re-gather their taps
//     from global memory (L2-resident, the planes were just streamed).
//   * slabs (multi-GPU halos, host streaming): the buffers hold a window of planes of the global grid;
//     plane indices stay global, so the boundary logic is the single-GPU one.
//
// Anything this kernel does not cover (non-central plans, mixed radii, unaligned shapes) falls
// back -- inside the library, still on the GPU -- to the composition of axis kernels.
#include <cuda.h>

#include <algorithm>
#include <cmath>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>
#include <vector>

#pragma once
#include "fdx_common.cuh"

extern "C" int fdx_axis_apply_f32(fdx_plan_t, const float*, float*, int64_t, int64_t, double, int, void*);
extern "C" int fdx_axis_apply_f64(fdx_plan_t, const double*, double*, int64_t, int64_t, double, int, void*);

namespace fdx {

enum Op { LAP = 0, DIV = 1, GRAD = 2, CURL = 3 };

// ------------------------------------------------------------------------------ PTX helpers
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
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}" ::"r"(smem_u32(bar)),
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

// ------------------------------------------------------------------------------ parameters
// Band tables, sized by the stencil radius so that TRANSPOSED plans stay fused at every accuracy the forward kernels
// cover (the adjoint's bands are P + L rows of up to 2L + 1 taps: 6 sides x 18 x 24 at P = 6).  Radius <= 3 keeps the
// small block (the latency-bound small grids of BASELINE config 1 pay for the size of the parameter block).
__host__ __device__ constexpr int tab_max(int P) { return P <= 3 ? 640 : (P == 4 ? 1280 : (P == 5 ? 1920 : 2688)); }  // band coefficients of all 6 sides
__host__ __device__ constexpr int xtab_max(int P) { return P <= 3 ? 320 : (P == 4 ? 384 : (P == 5 ? 640 : 896)); }    // x-band coefficients of both sides (staged in shared memory by the x-edge CTAs)
constexpr int kRowsMax = 128;  // band rows of all 6 sides
constexpr int kMaxPlanes = 383;  // planes one CTA may march through (per-plane flag bytes in shared memory)
constexpr int kMaxCuts = 128;    // z chunks per launch

template <typename T>
struct AxisDev {
  int n, nl, wl, nr, wr;
  int tab_l, tab_r;   // offsets of the dense band taps inside FusedParams::tab
  int rng_l, rng_r;   // offsets of the per-row nonzero ranges inside FusedParams::rng
  // head form of the bands (transposed plans, see corr_form()): a band row is the main stencil over the columns beyond
  // the head plus a few dense taps on the first kcl / last kcr elements of the axis.  The tables at tab_l / tab_r then
  // hold head[c][row] (alpha folded in): left rows 0 .. nlp-1, right rows n-nrp .. n-1 (nlp, nrp: nl, nr rounded up to
  // a vector; rows that are not band rows carry zeros), right columns n-kcr .. n-1.
  int corr, kcl, kcr, nlp, nrp;
  T alpha;            // scale applied on the band path
  T c[FDX_MAX_TAPS];  // main taps, pre-multiplied by alpha
};

template <typename T, int P_>
struct FusedParams {
  CUtensorMap tmap[3];  // one per input field
  const T* in[3];       // logical plane 0 of each input field (band gathers)
  T* out[3];
  AxisDev<T> ax[3];
  int z_begin, z_end;  // output planes
  int z_lo, z_hi;      // addressable input planes [z_lo, z_hi]
  int tz_off;          // tensor-map plane index = global plane index + tz_off (= -in_first)
  int n_chunks, tiles_x, tiles_y;
  // z chunks in launch order (graded: long chunks first, short ones last so that the tail of the
  // launch is made of short CTAs): first output plane and number of output planes of every chunk
  int cut_z[kMaxCuts];
  short cut_n[kMaxCuts];
  // fast x bands (band rows fit one 16-byte vector, taps fit three): per-tap coefficient vectors,
  // xcl[t][v] = tap t of left band row v (column v + t); xcr[t][v] = tap of right band row
  // v - (V - nr) that sits t columns before the row's own column
  static constexpr int kXV = 16 / (int)sizeof(T), kXT = 2 * kXV + 1;
  int xfast, xnt_l, xnt_r;
  int plain_ok;  // interior CTAs may take the specialised (flag-free) march
  int l2_ahead;   // planes the DRAM -> L2 prefetch runs ahead of the TMA loads (FDX_L2_AHEAD; 0 = none)
  int issue_rot;  // the TMA issue duty rotates over the warps of the CTA (FDX_ISSUE_ROT=0 switches it off)
  int rim_first;  // launch order inside a chunk layer: rim tiles first (FDX_FUSED_RIMFIRST=0 switches it off)
  int dbg;       // experiments only (FDX_FUSED_DBG): 2 no x-band work, 16 every tile takes the interior copy (both: wrong results at the edges)
  int ydirect;   // ... and the tiles with y-band rows (the first / last tile row holds every band row and all of its taps)
  int xdirect;   // ... and so may the tiles with x-band columns (the owner's window holds all band taps)
  T xcl[kXT][kXV], xcr[kXT][kXV];
  short2 rng[kRowsMax];  // [first, last+1) nonzero band column of every band row
  T tab[tab_max(P_)];
};

// TXV x TYT threads own RY rows x one 16-byte vector each; a warp covers LX x (32 / LX) of them.  A narrow
// warp (LX < TXV) keeps the x-band work of the first / last tile of a row inside fewer warps.
template <int TXV_, int TYT_, int RY_, int STAGES_, int MINB_ = 1, int LX_ = 0, int PW_ = 0>
struct Cfg {
  static constexpr int TXV = TXV_, TYT = TYT_, RY = RY_, STAGES = STAGES_, MINB = MINB_;
  static constexpr int PW = PW_;  // (sweeps) a producer warp of its own: the consumers never issue TMA
  static constexpr int LX = LX_ > 0 ? LX_ : (TXV_ < 32 ? TXV_ : 32);
  static_assert(32 % LX == 0 && TXV_ % LX == 0 && TYT_ % (32 / LX) == 0 && LX >= RY_, "warp shape");
};

template <Op OP>
struct OpTraits;
template <>
struct OpTraits<LAP> {
  static constexpr int NIN = 1, NOUT = 1, NACC = 1;
  __host__ __device__ static constexpr bool hx(int f) { return f == 0; }
  __host__ __device__ static constexpr bool hy(int f) { return f == 0; }
};
template <>
struct OpTraits<DIV> {
  static constexpr int NIN = 3, NOUT = 1, NACC = 1;
  __host__ __device__ static constexpr bool hx(int f) { return f == 2; }
  __host__ __device__ static constexpr bool hy(int f) { return f == 1; }
};
template <>
struct OpTraits<GRAD> {
  static constexpr int NIN = 1, NOUT = 3, NACC = 1;
  __host__ __device__ static constexpr bool hx(int f) { return f == 0; }
  __host__ __device__ static constexpr bool hy(int f) { return f == 0; }
};
template <>
struct OpTraits<CURL> {
  static constexpr int NIN = 3, NOUT = 3, NACC = 2;
  __host__ __device__ static constexpr bool hx(int f) { return f == 0 || f == 1; }
  __host__ __device__ static constexpr bool hy(int f) { return f == 0 || f == 2; }
};

template <Op OP, typename T, int P, typename CFG>
struct Geometry {
  using Tr = OpTraits<OP>;
  static constexpr int V = 16 / sizeof(T);
  static constexpr int HXV = (P + V - 1) / V;  // halo vectors each side in x
  static constexpr int HX = HXV * V;
  static constexpr int TX = CFG::TXV * V, TY = CFG::TYT * CFG::RY;
  static constexpr int NC = CFG::TXV * CFG::TYT;        // consumer threads
  static constexpr int NT = NC + (CFG::PW ? 32 : 0);   // threads (without a producer warp: elected consumer lanes issue TMA)
  static constexpr int WZ = 2 * P + 1;
  __host__ __device__ static constexpr int xoff(int f) { return Tr::hx(f) ? HX : 0; }
  __host__ __device__ static constexpr int yoff(int f) { return Tr::hy(f) ? P : 0; }
  __host__ __device__ static constexpr int box_x(int f) { return TX + 2 * xoff(f); }
  __host__ __device__ static constexpr int box_y(int f) { return TY + 2 * yoff(f); }
  __host__ __device__ static constexpr int field_bytes(int f) { return ((box_x(f) * box_y(f) * (int)sizeof(T) + 127) / 128) * 128; }
  __host__ __device__ static constexpr int field_off(int f) { return f == 0 ? 0 : field_off(f - 1) + field_bytes(f - 1); }
  static constexpr int STAGE_BYTES = field_off(Tr::NIN);
  __host__ __device__ static constexpr int tx_bytes() {
    int b = 0;
    for (int f = 0; f < Tr::NIN; ++f) b += box_x(f) * box_y(f) * (int)sizeof(T);
    return b;
  }
  static constexpr int XTAB_BYTES = ((xtab_max(P) * (int)sizeof(T) + 127) / 128) * 128;
  // barriers (128 B) + plane flags (384 B) + x-band table + alignment slack + stage ring
  static constexpr int SMEM_BYTES = 512 + XTAB_BYTES + 128 + CFG::STAGES * STAGE_BYTES;
};

template <int N>
using IC = std::integral_constant<int, N>;

// ------------------------------------------------------------------------------ vector FMA
// acc += c * x on 16-byte vectors.  fp32 uses the packed FFMA2 (fma.rn.f32x2, sm_100+): the
// coefficient is a uniform-register scalar broadcast to both halves, so one instruction issues two
// FMAs -- this kernel is bound by issue slots, not by the FP32 pipe.
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(unsigned long long p, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p));
}
template <bool INIT>
__device__ __forceinline__ void fma_vec(Vec<float, 4>& acc, float c, const Vec<float, 4>& x) {
  const unsigned long long cc = pack2(c, c);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    unsigned long long a = INIT ? 0ull : pack2(acc.v[2 * h], acc.v[2 * h + 1]);
    const unsigned long long xx = pack2(x.v[2 * h], x.v[2 * h + 1]);
    if constexpr (INIT)
      asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(a) : "l"(cc), "l"(xx));
    else
      asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a) : "l"(cc), "l"(xx));
    unpack2(a, acc.v[2 * h], acc.v[2 * h + 1]);
  }
}
template <bool INIT>
__device__ __forceinline__ void fma_vec(Vec<double, 2>& acc, double c, const Vec<double, 2>& x) {
#pragma unroll
  for (int v = 0; v < 2; ++v) acc.v[v] = INIT ? c * x.v[v] : fma(c, x.v[v], acc.v[v]);
}
__device__ __forceinline__ void add_vec(Vec<float, 4>& acc, const Vec<float, 4>& x) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    unsigned long long a = pack2(acc.v[2 * h], acc.v[2 * h + 1]);
    asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(pack2(x.v[2 * h], x.v[2 * h + 1])));
    unpack2(a, acc.v[2 * h], acc.v[2 * h + 1]);
  }
}
__device__ __forceinline__ void add_vec(Vec<double, 2>& acc, const Vec<double, 2>& x) {
#pragma unroll
  for (int v = 0; v < 2; ++v) acc.v[v] += x.v[v];
}
// dst = c * x + src (three-operand form: the queue shift of the z march)
__device__ __forceinline__ void fma_to(Vec<float, 4>& dst, float c, const Vec<float, 4>& x, const Vec<float, 4>& src) {
  const unsigned long long cc = pack2(c, c);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(d)
        : "l"(cc), "l"(pack2(x.v[2 * h], x.v[2 * h + 1])), "l"(pack2(src.v[2 * h], src.v[2 * h + 1])));
    unpack2(d, dst.v[2 * h], dst.v[2 * h + 1]);
  }
}
__device__ __forceinline__ void fma_to(Vec<double, 2>& dst, double c, const Vec<double, 2>& x, const Vec<double, 2>& src) {
#pragma unroll
  for (int v = 0; v < 2; ++v) dst.v[v] = fma(c, x.v[v], src.v[v]);
}

// ------------------------------------------------------------------------------ the kernel
// Shared-memory accesses of the hot path use 32-bit shared-window addresses (no generic->shared
// conversion inside the loop); `volatile` keeps them behind the mbarrier wait.
template <typename T, int V>
__device__ __forceinline__ Vec<T, V> lds_vec(uint32_t addr) {
  Vec<T, V> r;
  if constexpr (sizeof(T) == 4) {
    static_assert(V == 4, "fp32 vectors are 4 wide");
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]) : "r"(addr));
  } else {
    static_assert(V == 2, "fp64 vectors are 2 wide");
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(r.v[0]), "=d"(r.v[1]) : "r"(addr));
  }
  return r;
}
// (streaming stores, evict-first: an output is written once and never read back by the kernel, so it should not push
// the input lines that neighbouring tiles and chunks are about to re-read out of L2.  Same-box A/B against plain
// stores: laplacian / divergence 512^3 fp32 -0.4 %, gradient -0.9 %, fp64 neutral; in the TMA-ring axis and 2-D
// kernels it made no difference or cost 2 % (2-D gradient), they keep plain stores.)
template <typename T, int V>
__device__ __forceinline__ void stg_vec(uint64_t addr, const Vec<T, V>& r) {
  if constexpr (sizeof(T) == 4)
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(addr), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]) : "memory");
  else
    asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" ::"l"(addr), "d"(r.v[0]), "d"(r.v[1]) : "memory");
}
__device__ __forceinline__ void mbar_init_a(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void tma_load_3d_a(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// Which field feeds the q-th y-derivative / x-derivative of the in-plane part
//   LAP : ya0 = D1 x       xa0 = D2 x
//   DIV : ya0 = D1 x1      xa0 = D2 x2
//   GRAD: ya0 = D1 x       xa0 = D2 x
//   CURL: ya0 = D1 x2, ya1 = D1 x0      xa0 = D2 x1, xa1 = D2 x0
template <Op OP>
struct InPlane {
  static constexpr int NY = OP == CURL ? 2 : 1, NX = OP == CURL ? 2 : 1;
  __host__ __device__ static constexpr int yfield(int q) { return OP == DIV ? 1 : (OP == CURL ? (q == 0 ? 2 : 0) : 0); }
  __host__ __device__ static constexpr int xfield(int q) { return OP == DIV ? 2 : (OP == CURL ? (q == 0 ? 1 : 0) : 0); }
};

// output o of OP is written at plane zc - lag: P for the outputs that leave the z queue, 0 for in-plane ones
template <Op OP, int P>
__host__ __device__ constexpr int out_lag(int o) {
  return (OP == LAP || OP == DIV) ? P : (OP == GRAD ? (o == 0 ? P : 0) : (o == 0 ? 0 : P));
}

// CORR: the instantiation that knows the head form of the bands (AxisDev::corr; launched when any axis has it:
// the adjoints).  The plain instantiation carries none of that code, so the forward operators keep their registers.
template <Op OP, typename T, int P, typename CFG, bool CORR>
__global__ void __launch_bounds__(Geometry<OP, T, P, CFG>::NT, CFG::MINB)
    fused3d_kernel(const __grid_constant__ FusedParams<T, P> prm) {
  using G = Geometry<OP, T, P, CFG>;
  using Tr = OpTraits<OP>;
  using IP = InPlane<OP>;
  constexpr int V = G::V, RY = CFG::RY, WZ = G::WZ, STAGES = CFG::STAGES, HX = G::HX, HXV = G::HXV;
  constexpr int NY = IP::NY, NX = IP::NX;
  constexpr int LX = CFG::LX, TRW = 32 / LX;  // a warp covers LX x TRW threads
  constexpr int WR = TRW * RY;               // grid rows per warp
  // planes in flight: the stage refilled at step `it` was released at step it - (STAGES - LOOK), so
  // with LOOK = STAGES - 2 the issuing thread almost never waits for the slowest warp
  constexpr int LOOK = STAGES >= 4 ? STAGES - 2 : STAGES - 1;
  static_assert(STAGES * 16 <= 128, "barrier area");
  using VT = Vec<T, V>;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  // barriers first (full[STAGES], empty[STAGES]), one flag byte per plane of the march at +128, then
  // the stage ring rounded up to 128 B for TMA
  uint32_t smem0 = smem_u32(smem_raw);
  asm volatile("" : "+r"(smem0));  // opaque: computed once, kept in a register
  const uint32_t ring = (smem0 + 512u + (uint32_t)G::XTAB_BYTES + 127u) & ~127u;
  const uint32_t bar_full = smem0, bar_empty = smem0 + 8u * STAGES, flag_base = smem0 + 128u, xtab = smem0 + 512u;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int n0 = prm.ax[0].n, n1 = prm.ax[1].n, n2 = prm.ax[2].n;

  // ---- work item: (tile, chunk); tile fastest so that concurrently resident CTAs share halos in L2
  // Within a chunk layer the tiles on the rim of the (y,x) plane come first: they carry the band work and are
  // the slowest CTAs, so they must not be the last ones of the launch.
  int tile_x, tile_y, cslot;
  {
    const int txn = prm.tiles_x, tyn = prm.tiles_y, tiles = txn * tyn;
    const int b = blockIdx.x;
    cslot = b / tiles;  // chunks are listed in launch order (see launch_cfg)
    int r = b - cslot * tiles;
    if (txn < 3 || tyn < 3 || prm.rim_first == 0) {
      tile_x = r % txn, tile_y = r / txn;
    } else if (r < 2 * txn) {
      tile_y = r < txn ? 0 : tyn - 1, tile_x = r < txn ? r : r - txn;
    } else if (r < 2 * txn + 2 * (tyn - 2)) {
      r -= 2 * txn;
      tile_y = 1 + (r >> 1), tile_x = (r & 1) ? txn - 1 : 0;
    } else {
      r -= 2 * txn + 2 * (tyn - 2);
      tile_y = 1 + r / (txn - 2), tile_x = 1 + r % (txn - 2);
    }
  }
  // the last tile of a row is shifted left so that it is a full tile ending at column n2 (it then holds
  // the whole right band and all of its taps however n2 relates to TX); a tile writes only the columns
  // before the next tile's origin, so no output is written twice
  const int x_shift = max(0, n2 - G::TX);
  const int x0 = min(tile_x * G::TX, x_shift), y0 = tile_y * G::TY;
  // ... except that with two tiles the first one keeps every left band column (wide adjoint bands may reach past
  // the origin of the shifted tile): the cut between the two is max(x_shift, nl rounded up to a vector)
  const int x_cut2 = min(G::TX, max(x_shift, ((prm.ax[2].nl + V - 1) / V) * V));
  const int x_hi = tile_x + 1 < prm.tiles_x ? (prm.tiles_x == 2 ? x_cut2 : min((tile_x + 1) * G::TX, x_shift)) : n2;
  const int x_lo = (prm.tiles_x == 2 && tile_x == 1) ? x_cut2 : x0;  // first column this tile writes
  const int zs = prm.cut_z[cslot];
  const int ze = zs + prm.cut_n[cslot];
  const int zin0 = max(zs - P, prm.z_lo);
  const int zin1 = min(ze - 1 + P, prm.z_hi);
  const int n_planes = zin1 - zin0 + 1;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init_a(bar_full + 8u * s, 1);
      mbar_init_a(bar_empty + 8u * s, G::NC / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // ---- CTA classes (uniform): only `edge` CTAs and z-boundary planes ever enter the cold section
  const bool xedge_l = (tile_x == 0) && prm.ax[2].nl > 0;  // first / last tile along x hold the whole
  const bool xedge_r = (tile_x == prm.tiles_x - 1) && prm.ax[2].nr > 0;  // left / right band (host-checked)
  const bool yedge = (y0 < prm.ax[1].nl) || (y0 + G::TY > n1 - prm.ax[1].nr);
  const bool edge = xedge_l || xedge_r || yedge;
  // per-plane flags of this CTA's march, one byte each: the loop reads them instead of re-deriving them
  // (z axis in head form: the band planes leave the queue like every other plane -- planes outside the grid count
  // as zero, head planes are kept out of the queue -- and receive their head taps when they are emitted: F_EZB;
  // no direct gathers, no F_ZB)
  enum : uint32_t { F_INR = 1, F_EMIT = 2, F_ZB = 4, F_SLOW = 8, F_EZB = 16 };
  const bool zcorr = CORR && prm.ax[0].corr != 0;
  for (int i = tid; i < n_planes; i += G::NT) {
    const int z = zin0 + i, zo = z - P;
    const bool inr = z >= zs && z < ze;                                             // an output plane of ours
    const bool zb = !zcorr && inr && (z < prm.ax[0].nl || z >= n0 - prm.ax[0].nr);  // next to a physical z boundary
    const bool zbo = zo < prm.ax[0].nl || zo >= n0 - prm.ax[0].nr;
    const bool emit = zo >= zs && (zcorr || !zbo);                                  // plane z - P leaves the queue
    const uint32_t f = (inr ? F_INR : 0u) | (emit ? F_EMIT : 0u) | (zb ? F_ZB : 0u) | ((inr && (edge || zb)) ? F_SLOW : 0u) |
                       ((CORR && emit && zbo) ? F_EZB : 0u);
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(flag_base + (uint32_t)i), "r"(f) : "memory");
  }
  if (xedge_l || xedge_r) {  // the dense x-band rows (left, then right) or the x head tables go to shared memory
    const int cnt = (CORR && prm.ax[2].corr) ? prm.ax[2].kcl * prm.ax[2].nlp + prm.ax[2].kcr * prm.ax[2].nrp
                                   : prm.ax[2].nl * prm.ax[2].wl + prm.ax[2].nr * prm.ax[2].wr;  // <= xtab_max(P), host-checked
    for (int i = tid; i < cnt; i += G::NT) {
      const T c = prm.tab[prm.ax[2].tab_l + i];
      if constexpr (sizeof(T) == 4)
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(xtab + 4u * (uint32_t)i), "f"(c) : "memory");
      else
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(xtab + 8u * (uint32_t)i), "d"(c) : "memory");
    }
  }
  __syncthreads();

  // TMA issue (thread 0 only): plane `it` of this CTA's march goes to stage it % STAGES.
  // (Round 2, measured and dropped: a non-blocking issue -- mbarrier.test_wait on the empty barrier, planes that cannot go
  // out yet go out a step later, up to STAGES planes in flight.  ncu had shown thread 0's blocking wait on the empty
  // barrier as the most-sampled stall, but that wait is benign -- warp 0 runs ahead of the slowest warp and the ring
  // throttles it -- while the extra test / counter work sits on warp 0's critical path: laplacian fp32 -7 %, fp64 -8 %.)
  auto issue = [&](int it) {
    const int s = it % STAGES;
    if (it >= STAGES) mbar_wait_a(bar_empty + 8u * s, ((it / STAGES) - 1) & 1);  // all warps released it
    mbar_expect_tx_a(bar_full + 8u * s, G::tx_bytes());
    const uint32_t dst = ring + (uint32_t)s * G::STAGE_BYTES;
    const int zt = zin0 + it + prm.tz_off;
#pragma unroll
    for (int f = 0; f < Tr::NIN; ++f)
      tma_load_3d_a(dst + G::field_off(f), &prm.tmap[f], bar_full + 8u * s, x0 - G::xoff(f), y0 - G::yoff(f), zt);
    // DRAM -> L2 prefetch of a plane further ahead than the shared-memory ring reaches (prm.l2_ahead planes; 0 = off)
    if (prm.l2_ahead > 0 && it + prm.l2_ahead < n_planes) {
#pragma unroll
      for (int f = 0; f < Tr::NIN; ++f)
        tma_prefetch_3d(&prm.tmap[f], x0 - G::xoff(f), y0 - G::yoff(f), zt + prm.l2_ahead);
    }
  };
  if constexpr (CFG::PW) {
    // (sweep configurations) the producer warp: one lane issues every plane of the march, as far ahead as the ring allows
    if (tid >= G::NC) {
      if (tid == G::NC) {
        for (int f = 0; f < Tr::NIN; ++f) asm volatile("prefetch.tensormap [%0];" ::"l"(&prm.tmap[f]) : "memory");
        for (int it = 0; it < n_planes; ++it) issue(it);
      }
      return;
    }
  } else if (tid == 0) {
    for (int f = 0; f < Tr::NIN; ++f) asm volatile("prefetch.tensormap [%0];" ::"l"(&prm.tmap[f]) : "memory");
    for (int it = 0; it < LOOK && it < n_planes; ++it) issue(it);
  }
  // The issue duty rotates over the warps (lane 0 of warp it % NW issues at step it): the ~35 instructions of an issue
  // are then paid by every warp once in NW steps instead of by warp 0 at every step -- warp 0 was the slowest warp
  // of an interior CTA and the step time of a CTA is that of its slowest warp.  (FDX_ISSUE_ROT=0: thread 0 always.)
  constexpr int NW = G::NC / 32;
  const bool rot = prm.issue_rot != 0;
  auto issuer = [&](int it) { return CFG::PW ? -1 : (rot ? (it % NW) * 32 : 0); };

  const int wrp = tid / 32;
  const int tx = (wrp % (CFG::TXV / LX)) * LX + lane % LX, ty = (wrp / (CFG::TXV / LX)) * TRW + lane / LX;
  const int xg = x0 + tx * V;   // first global x of this thread's vector
  const int yg = y0 + ty * RY;  // first global y of this thread's rows
  const AxisDev<T>& A0 = prm.ax[0];
  const AxisDev<T>& A1 = prm.ax[1];
  const AxisDev<T>& A2 = prm.ax[2];
  const bool xcorr = CORR && A2.corr != 0, ycorr = CORR && A1.corr != 0;  // (constants in the plain instantiation)
  const int plane = n1 * n2;  // < 2^31, host-checked
  // n2 % V == 0, so a vector is entirely in or out of the grid
  // row j of this thread lies inside the grid: bit j (opaque, so that it stays one register)
  uint32_t rmask = 0;
#pragma unroll
  for (int j = 0; j < RY; ++j) rmask |= ((xg >= x_lo) && (xg < x_hi) && (yg + j < n1)) ? (1u << j) : 0u;
  asm volatile("" : "+r"(rmask));
  auto rok = [&](int j) { return (rmask >> j) & 1u; };

  const bool yband = yedge && rok(0) && ((yg < A1.nl) || (yg + RY > n1 - A1.nr));

  const int64_t thread_ofs = (int64_t)yg * n2 + xg;  // this thread's first element inside a plane
  // shared address (stage 0) of this thread's vector in tile row ty*RY of each field
  uint32_t tb[Tr::NIN];
#pragma unroll
  for (int f = 0; f < Tr::NIN; ++f)
  {
    tb[f] = ring + G::field_off(f) + (uint32_t)sizeof(T) * ((ty * RY) * G::box_x(f) + G::xoff(f) + tx * V);
    asm volatile("" : "+r"(tb[f]));  // opaque: one register instead of a re-derivation per plane
  }
  // output row bases: address of element (plane 0, row yg + j, column xg) of every output, held in
  // registers (opaque) so that a store costs one IMAD.WIDE for its address
  uint64_t ob[Tr::NOUT][RY];
#pragma unroll
  for (int o = 0; o < Tr::NOUT; ++o)
#pragma unroll
    for (int j = 0; j < RY; ++j) {
      ob[o][j] = (uint64_t)(uintptr_t)(prm.out[o] + thread_ofs + (int64_t)j * n2);
      asm volatile("" : "+l"(ob[o][j]));
    }
  const int plane_b = plane * (int)sizeof(T);  // < 2^31, host-checked

  // Bands in head form (CORR instantiation, AxisDev::corr): the main stencil of a band row must not see the head
  // elements of its axis (the first kcl / last kcr: they enter through the head table instead).  Every output that a
  // head element reaches through the main stencil is a band row (host-checked: kc + P <= nl), so the elements are
  // simply hidden from the main path -- bit j of xmask: entry j of this thread's x window is a head column; bit i of
  // ymask: centre row i of its y march is a head row (thread constants); along z the head planes skip the queue.
  constexpr int XW = (2 * HXV + 1) * V;
  uint32_t xmask = 0, ymask = 0;
  if constexpr (CORR) {
    if (xcorr) {
#pragma unroll
      for (int j = 0; j < XW; ++j) {
        const int col = xg - HX + j;
        if ((col >= 0 && col < A2.kcl) || (col >= n2 - A2.kcr && col < n2)) xmask |= 1u << j;
      }
    }
    if (ycorr) {
#pragma unroll
      for (int i = 0; i < RY + 2 * P; ++i) {
        const int row = yg - P + i;
        if ((row >= 0 && row < A1.kcl) || (row >= n1 - A1.kcr && row < n1)) ymask |= 1u << i;
      }
    }
    asm volatile("" : "+r"(xmask), "+r"(ymask));
  }

  // ------------------------------------------------------------------ hot in-plane helpers
  // D_y, streaming: centre row i of the RY+2P rows a thread sees feeds output rows j = i-2P .. i
  // with tap i-j, so only one centre row is live at a time (keeps the register footprint small)
  // (mk: a constant at every call site -- the copies of the march that serve tiles with band rows apply the masks)
  auto dy_feed = [&](VT (&ys)[RY], const VT& ci_in, int i, bool mk) {  // i is a constant after unrolling
    VT ci = ci_in;
    if constexpr (CORR) {
      if (mk && ((ymask >> i) & 1u)) {
#pragma unroll
        for (int v = 0; v < V; ++v) ci.v[v] = 0;
      }
    }
#pragma unroll
    for (int j = 0; j < RY; ++j) {
      const int k = i - j;
      if (k == 0)
        fma_vec<true>(ys[j], A1.c[0], ci);
      else if (k > 0 && k < WZ)
        fma_vec<false>(ys[j], A1.c[k], ci);
    }
  };
  // D_x for one row: `row` is the shared address of this thread's vector inside a tile row with x halos
  // With `xb` (a constant at every call site; specialised march of a tile with x-band columns) the thread
  // that owns the band columns of a side replaces them: band row v is the dot product of its taps
  // xcl[t][v] / xcr[t][v] (prm.xfast layout) with the window it already holds, taps in ascending column
  // order like the reference's sum.  The host guarantees that the window reaches far enough (prm.xdirect).
  auto lds_s = [&](uint32_t a) {
    T r;
    if constexpr (sizeof(T) == 4)
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(a));
    else
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r) : "r"(a));
    return r;
  };
  // x bands in head form (A2.corr): the band rows inside this thread's vector receive their head taps -- the first /
  // last columns of the grid row -- on top of the main stencil, which ran with those columns masked (xmask).  What
  // depends on the thread only is computed once: xc_kc columns (0: no band row of either side in this vector -- a
  // vector never holds band rows of both, corr_form()), xc_xoff = bytes from the thread's vector to the first of
  // those columns inside a tile row, xc_tab = shared address of its first coefficient vector (tables at `xtab`: left
  // block, then right block).  Columns in ascending order; the rim copy and the generic march share the code (same bits).
  int xc_kc = 0, xc_xoff = 0;
  uint32_t xc_tab = 0;
  if (xcorr) {
    if (xedge_l && xg < A2.nl) {
      xc_kc = A2.kcl, xc_xoff = -xg * (int)sizeof(T), xc_tab = xtab + (uint32_t)(xg * (int)sizeof(T));
    } else if (xedge_r && xg + V > n2 - A2.nr && xg < n2) {
      xc_kc = A2.kcr, xc_xoff = (n2 - A2.kcr - xg) * (int)sizeof(T);
      xc_tab = xtab + (uint32_t)((A2.kcl * A2.nlp + (xg - (n2 - A2.nrp))) * (int)sizeof(T));
    }
  }
  // xs[j] += sum_c head[c][.] * x[row j][c]; row0 = shared address of this thread's vector in the row of yg
  auto xcorr_apply = [&](uint32_t row0, int pitch_b, VT (&xs)[RY]) {
    if (xc_kc == 0) return;
    const uint32_t tpitch = (uint32_t)((xg < A2.nl ? A2.nlp : A2.nrp) * (int)sizeof(T));
    uint32_t xa_ = row0 + (uint32_t)xc_xoff, ca = xc_tab;
#pragma unroll 1
    for (int c = xc_kc; c > 0; --c) {
      const VT cf = lds_vec<T, V>(ca);
#pragma unroll
      for (int j = 0; j < RY; ++j) {
        const T xv = lds_s(xa_ + (uint32_t)(j * pitch_b));
#pragma unroll
        for (int v = 0; v < V; ++v) xs[j].v[v] = fma(cf.v[v], xv, xs[j].v[v]);
      }
      xa_ += (uint32_t)sizeof(T), ca += tpitch;
    }
  };
  const bool xdir = prm.xdirect && !(prm.dbg & 2);
  const bool own_l = xedge_l && tx == 0 && xdir, own_r = xedge_r && tx == (n2 - V - x0) / V && xdir;
  auto dx_main = [&](uint32_t row, const VT& centre, bool xb, bool mk) {
    constexpr int W = (2 * HXV + 1) * V;
    T w[W];
#pragma unroll
    for (int h = 0; h < HXV; ++h) {
      const VT l = lds_vec<T, V>(row - (uint32_t)((HXV - h) * V * sizeof(T)));
      const VT rr = lds_vec<T, V>(row + (uint32_t)((h + 1) * V * sizeof(T)));
#pragma unroll
      for (int v = 0; v < V; ++v) {
        w[h * V + v] = l.v[v];
        w[(HXV + 1 + h) * V + v] = rr.v[v];
      }
    }
#pragma unroll
    for (int v = 0; v < V; ++v) w[HXV * V + v] = centre.v[v];
    if constexpr (CORR) {
      if (mk && xmask) {
#pragma unroll
        for (int j = 0; j < W; ++j)
          if ((xmask >> j) & 1u) w[j] = 0;
      }
    }
    VT r;
    if constexpr (sizeof(T) == 4) {
      // outputs (0,1) and (2,3): tap k of output v reads w[HX + v - P + k].  When that index is even for the first
      // output of a pair, the two operands are an aligned register pair of the window and the tap is one packed
      // FFMA2; the odd taps stay scalar (building their pairs would cost two register moves each: measured, no
      // gain).  Per output the taps are summed in ascending order either way: the bits of the scalar form.
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int k = 0; k < WZ; ++k) {
          const int j = HX + 2 * h - P + k;
          if (j % 2 == 0 && k > 0) {
            unsigned long long a = pack2(s0, s1);
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a) : "l"(pack2(A2.c[k], A2.c[k])), "l"(pack2(w[j], w[j + 1])));
            unpack2(a, s0, s1);
          } else if (k == 0) {
            s0 = A2.c[0] * w[j], s1 = A2.c[0] * w[j + 1];
          } else {
            s0 = fma(A2.c[k], w[j], s0), s1 = fma(A2.c[k], w[j + 1], s1);
          }
        }
        r.v[2 * h] = s0, r.v[2 * h + 1] = s1;
      }
    } else {
#pragma unroll
      for (int v = 0; v < V; ++v) {
        T s = A2.c[0] * w[HX + v - P];
#pragma unroll
        for (int k = 1; k < WZ; ++k) s = fma(A2.c[k], w[HX + v - P + k], s);
        r.v[v] = s;
      }
    }
    if (xb && !xcorr) {
      constexpr int XT = FusedParams<T, P>::kXT;
      if (own_l) {
        T val[V];
#pragma unroll
        for (int v = 0; v < V; ++v) val[v] = 0;
#pragma unroll
        for (int t = 0; t < XT; ++t) {
          if (t < prm.xnt_l) {
#pragma unroll
            for (int v = 0; v < V; ++v)
              if (HX + v + t < W) val[v] = fma(prm.xcl[t][v], w[HX + v + t], val[v]);
          }
        }
#pragma unroll
        for (int v = 0; v < V; ++v)
          if (v < A2.nl) r.v[v] = val[v] * A2.alpha;
      }
      if (own_r) {
        T val[V];
#pragma unroll
        for (int v = 0; v < V; ++v) val[v] = 0;
#pragma unroll
        for (int t = XT - 1; t >= 0; --t) {
          if (t < prm.xnt_r) {
#pragma unroll
            for (int v = 0; v < V; ++v)
              if (HX + v - t >= 0) val[v] = fma(prm.xcr[t][v], w[HX + v - t], val[v]);
          }
        }
#pragma unroll
        for (int v = 0; v < V; ++v)
          if (v >= V - A2.nr) r.v[v] = val[v] * A2.alpha;
      }
    }
    return r;
  };

  // ------------------------------------------------------------------ cold band helpers
  // band row/column/plane i of axis A: coefficient offset, nonzero column range, first band column
  auto band_rng = [&](const AxisDev<T>& A, int i, int& tab, int& c0, int& c1, int& colbase) {
    const bool left = i < A.nl;
    const int r = left ? i : i - (A.n - A.nr);
    const int w = left ? A.wl : A.wr;
    tab = (left ? A.tab_l : A.tab_r) + r * w;
    const short2 rg = prm.rng[(left ? A.rng_l : A.rng_r) + r];
    c0 = rg.x, c1 = rg.y;
    colbase = left ? 0 : A.n - A.wr;
  };
  // taps c in [c0, c1) in batches of four: all loads of a batch are issued before the first FMA, so
  // a cold path costs about one load latency per batch instead of one per tap
  auto band_dot_vec = [&](int tab, int c0, int c1, auto load /* (int c) -> VT */) {
    VT a;
#pragma unroll
    for (int v = 0; v < V; ++v) a.v[v] = 0;
    for (int c = c0; c < c1; c += 4) {
      T cf[4];
      VT x[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int cc = min(c + q, c1 - 1);
        cf[q] = prm.tab[tab + cc];
        if (c + q >= c1) cf[q] = 0;
        x[q] = load(cc);
      }
      // a tap whose coefficient is exactly zero (the padding of the last batch, zeros inside a transposed band) is
      // skipped, as the axis kernels' band_element does: 0 * inf must not turn a band row into NaN
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int v = 0; v < V; ++v) a.v[v] = cf[q] != (T)0 ? fma(cf[q], x[q].v[v], a.v[v]) : a.v[v];
    }
    return a;
  };
  // y band of global row y: taps from the tile when it holds all of them, else from global memory.
  // tcol: generic address of (tile row 0, this thread's x); gcol: global address of (z, row 0, xg)
  auto band_y = [&](const T* tcol, int pitch, int boxy, const T* gcol, int y, bool tile_only) {
    int tab, c0, c1, cb;
    band_rng(A1, y, tab, c0, c1, cb);
    VT a;
    const int r0 = cb - (y0 - P);  // tile row of band column 0
    if (tile_only || (r0 + c0 >= 0 && r0 + c1 <= boxy)) {  // (tile_only: host-checked, prm.ydirect)
      const T* p = tcol + r0 * pitch;
      a = band_dot_vec(tab, c0, c1, [&](int c) { return *reinterpret_cast<const VT*>(p + c * pitch); });
    } else {
      const T* p = gcol + (int64_t)cb * n2;
      a = band_dot_vec(tab, c0, c1, [&](int c) { return ld_cached<T, V>(p + (int64_t)c * n2); });
    }
#pragma unroll
    for (int v = 0; v < V; ++v) a.v[v] *= A1.alpha;
    return a;
  };
  // y band of global row y from the tile alone (specialised march; host-checked, prm.ydirect): one 128-bit
  // shared load and one coefficient per tap, taps in ascending order like band_dot_vec (same bits).
  // tcol: shared address of (tile row 0, this thread's x) of the field
  auto band_y_tile = [&](uint32_t tcol, int pitch_b, int y) {
    int tab, c0, c1, cb;
    band_rng(A1, y, tab, c0, c1, cb);
    uint32_t addr = tcol + (uint32_t)((cb - (y0 - P) + c0) * pitch_b);
    VT a;
#pragma unroll
    for (int v = 0; v < V; ++v) a.v[v] = 0;
#pragma unroll 2
    for (int c = c0; c < c1; ++c) {
      const T cf = prm.tab[tab + c];
      const VT x = lds_vec<T, V>(addr);
      fma_vec<false>(a, cf, x);  // (packed for fp32: two independent fmas, the bits of the scalar form)
      addr += (uint32_t)pitch_b;
    }
#pragma unroll
    for (int v = 0; v < V; ++v) a.v[v] *= A1.alpha;
    return a;
  };
  // x bands, warp-cooperative: the band outputs of the WR grid rows of this warp are spread over
  // the lanes, one element each -- lane = r + WR * bl evaluates band column b0 + bl of warp row r
  // from the tile in shared memory (scalar loads, coefficients from the table staged at `xtab`, taps
  // in ascending column order like the reference's sum); the results travel to the lanes that own
  // those elements with shuffles.  The host guarantees that the first (last) tile holds every left
  // (right) band output and all of its taps.  `row0` = shared address of tile column 0 of the tile
  // row that holds grid row y0 of this field.
  constexpr int NB = 32 / WR;  // band columns per pass
  auto xband_coop = [&](uint32_t row0, int pitch, VT (&xs)[RY]) {
    const int r = lane % WR, bl = lane / WR;
    const uint32_t rowa = row0 + (uint32_t)(((ty / TRW) * WR + r) * pitch * (int)sizeof(T));
    const int myrow = (ty % TRW) * RY;
#pragma unroll 1
    for (int side = 0; side < 2; ++side) {
      if (!(side ? xedge_r : xedge_l)) continue;
      {  // warp-uniform: a warp whose columns hold no band element of this side has nothing to do
        const int wx0 = x0 + (wrp % (CFG::TXV / LX)) * LX * V;
        if (side ? (wx0 + LX * V <= n2 - A2.nr) : (wx0 >= A2.nl)) continue;
      }
      const int nb = side ? A2.nr : A2.nl, w = side ? A2.wr : A2.wl;
      const int tab0 = side ? A2.nl * A2.wl : 0;          // first coefficient of this side in the table
      const int col0 = (side ? n2 - A2.wr : 0) - (x0 - HX);  // tile column of band column 0
      const int xb0 = side ? n2 - A2.nr : 0;              // global x of band row 0
#pragma unroll 1
      for (int b0 = 0; b0 < nb; b0 += NB) {
        const int b = b0 + bl;
        T val = 0;
        if (b < nb) {
          const short2 rg = prm.rng[(side ? A2.rng_r : A2.rng_l) + b];
          uint32_t ca = xtab + (uint32_t)((tab0 + b * w + rg.x) * (int)sizeof(T));
          uint32_t xa_ = rowa + (uint32_t)((col0 + rg.x) * (int)sizeof(T));
          T a = 0;
#pragma unroll 2
          for (int c = rg.x; c < rg.y; ++c) {
            a = fma(lds_s(ca), lds_s(xa_), a);
            ca += (uint32_t)sizeof(T), xa_ += (uint32_t)sizeof(T);
          }
          val = a * A2.alpha;
        }
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const int bb = xg + v - xb0 - b0;  // band column of this element relative to the pass
          const bool mine = bb >= 0 && bb < NB && bb + b0 < nb && xg < n2;
          const int col = WR * (mine ? bb : 0);
#pragma unroll
          for (int j = 0; j < RY; ++j) {
            const T got = __shfl_sync(0xffffffffu, val, (myrow + j + col) & 31);
            if (mine) xs[j].v[v] = got;
          }
        }
      }
    }
  };
  // x bands, fast form (prm.xfast): every band row of a side lies in ONE thread's vector (the owner,
  // tx = 0 on the left, the last valid vector on the right) and its taps in three vectors, so they are
  // evaluated from registers with static indices, taps in ascending column order like the reference's
  // sum.  The owner's RY rows are spread over RY neighbouring lanes (helper h takes row h) and come
  // back with shuffles, so a warp pays for one row, not RY.
  const bool xfast = prm.xfast != 0;
  auto xband_fast = [&](uint32_t row /* shared address of this thread's vector in the row of yg */, int pitch, VT (&xs)[RY]) {
    constexpr int XT = FusedParams<T, P>::kXT;
#pragma unroll
    for (int side = 0; side < 2; ++side) {
      if (!(side ? xedge_r : xedge_l)) continue;  // CTA-uniform
      const int txo = side ? (n2 - V - x0) / V : 0;  // the owner's tx
      const int h = side ? txo - tx : tx;             // this lane helps with row h of the owner
      T val[V];
#pragma unroll
      for (int v = 0; v < V; ++v) val[v] = 0;
      if (h >= 0 && h < RY) {
        // the owner's vector in row h, and the two vectors towards the interior
        const uint32_t base = row + (uint32_t)(((txo - tx) * V + h * pitch) * (int)sizeof(T));
        T w[3 * V];
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          const VT t = lds_vec<T, V>(side ? base - (uint32_t)((2 - g) * V * sizeof(T)) : base + (uint32_t)(g * V * sizeof(T)));
#pragma unroll
          for (int v = 0; v < V; ++v) w[g * V + v] = t.v[v];
        }
        if (side == 0) {
#pragma unroll
          for (int t = 0; t < XT; ++t) {
            if (t < prm.xnt_l) {
#pragma unroll
              for (int v = 0; v < V; ++v) val[v] = fma(prm.xcl[t][v], w[v + t], val[v]);
            }
          }
        } else {
#pragma unroll
          for (int t = XT - 1; t >= 0; --t) {
            if (t < prm.xnt_r) {
#pragma unroll
              for (int v = 0; v < V; ++v) val[v] = fma(prm.xcr[t][v], w[2 * V + v - t], val[v]);
            }
          }
        }
#pragma unroll
        for (int v = 0; v < V; ++v) val[v] *= A2.alpha;
      }
      const bool owner = tx == txo;
#pragma unroll
      for (int j = 0; j < RY; ++j) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const T got = j == 0 ? val[v] : __shfl_sync(0xffffffffu, val[v], (lane + (side ? -j : j)) & 31);
          if (owner && (side ? v >= V - A2.nr : v < A2.nl)) xs[j].v[v] = got;
        }
      }
    }
  };
  // z band of plane z (always from global memory: the other planes are not in the ring)
  auto band_z = [&](const T* gcol0 /* global address of (plane 0, this row, xg) */, int z) {
    int tab, c0, c1, cb;
    band_rng(A0, z, tab, c0, c1, cb);
    const T* p = gcol0 + (int64_t)cb * plane;
    VT a = band_dot_vec(tab, c0, c1, [&](int c) { return ld_cached<T, V>(p + (int64_t)c * plane); });
#pragma unroll
    for (int v = 0; v < V; ++v) a.v[v] *= A0.alpha;
    return a;
  };
  auto is_yband = [&](int j) { return rok(j) && (yg + j < A1.nl || yg + j >= n1 - A1.nr); };
  // y / z bands in head form: band row y (plane z) = the main stencil over the rows beyond the head (already in `acc`)
  // + the head taps on the first / last rows (planes) of the axis, ascending.  `load(c)` returns this thread's vector of row (plane) c0 + c.
  auto corr_rows = [&](const AxisDev<T>& A, int i, T sgn, VT& acc, auto load /* (int global row or plane) -> VT */, auto batched) {
    const bool left = i < A.nl;
    const int kc = left ? A.kcl : A.kcr, pitch = left ? A.nlp : A.nrp;
    const int c0 = left ? 0 : A.n - A.kcr;
    int ti = left ? A.tab_l + i : A.tab_r + (i - (A.n - A.nrp));
    if constexpr (decltype(batched)::value) {
      // four rows per batch, their loads issued together (the z form gathers from L2: one round trip per batch); a row
      // past the last one repeats it with a zero factor, so that every element of the batch is defined
#pragma unroll 1
      for (int c = 0; c < kc; c += 4) {
        T cf[4];
        VT x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int cu = min(u, kc - 1 - c);
          cf[u] = (c + u < kc) ? sgn * prm.tab[ti + cu * pitch] : (T)0;
          x[u] = load(c0 + c + cu);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < V; ++v) acc.v[v] = fma(cf[u], x[u].v[v], acc.v[v]);
        ti += 4 * pitch;
      }
    } else {
#pragma unroll 1
      for (int c = 0; c < kc; ++c) {
        const T cf = sgn * prm.tab[ti];
        const VT x = load(c0 + c);
#pragma unroll
        for (int v = 0; v < V; ++v) acc.v[v] = fma(cf, x.v[v], acc.v[v]);
        ti += pitch;
      }
    }
  };
  // ... y, rows from the tile (tcol: shared address of (tile row 0, this thread's x) of the field)
  auto ycorr_tile = [&](uint32_t tcol, int pitch_b, int y, VT& acc) {
    corr_rows(A1, y, (T)1, acc, [&](int row) { return lds_vec<T, V>(tcol + (uint32_t)((row - (y0 - P)) * pitch_b)); }, std::false_type{});
  };
  // does the tile hold the head rows of its band rows?  (first / last tile row; uniform per CTA)
  const bool ycorr_ok = ycorr && (!(y0 < A1.nl) || (y0 <= P && A1.kcl <= y0 + G::TY + P)) &&
                        (!(y0 + G::TY > n1 - A1.nr) || (n1 - A1.kcr >= y0 - P && n1 <= y0 + G::TY + P));
  auto store = [&](int o, int z, int j, const VT& val) {
    if (rok(j)) stg_vec<T, V>(ob[o][j] + (uint64_t)((int64_t)z * (int64_t)plane_b), val);
  };

  // ---- march along z ---------------------------------------------------------------------
  // The register queue is a shift register: before plane zc arrives, q[a][i][j] holds the partial sum
  // of output plane zc + P - 1 - i, which has received its taps 0 .. i.  Plane zc completes the oldest
  // entry (tap 2P, emitted) and moves every other entry one place with a three-operand FMA,
  //     q[i] = q[i-1] + c_i * x      (i = 2P-1 .. 1),      q[0] = c_0 * x,
  // so the rotation costs nothing, all register indices are static and there is one copy of the code.
  VT q[Tr::NACC][2 * P][RY];
#pragma unroll
  for (int a = 0; a < Tr::NACC; ++a)
#pragma unroll
    for (int i = 0; i < 2 * P; ++i)
#pragma unroll
      for (int j = 0; j < RY; ++j)
#pragma unroll
        for (int v = 0; v < V; ++v) q[a][i][j].v[v] = 0;

  int stage = 0;           // it % STAGES
  uint32_t soff = 0;       // stage * STAGE_BYTES
  uint32_t full_parity = 0;

  // ---------------- pieces of one march step (shared by the generic and the specialised loop)
  // hot: loads of plane zc and its in-plane taps
  auto hot_inplane = [&](VT (&zv)[Tr::NACC][RY], VT (&ya)[NY][RY], VT (&xa)[NX][RY], bool xb, bool mk) {
    if constexpr (OP == LAP || OP == GRAD) {
#pragma unroll
      for (int i = 0; i < RY + 2 * P; ++i) {
        const uint32_t row = tb[0] + soff + (uint32_t)(i * G::box_x(0) * sizeof(T));
        const VT ci = lds_vec<T, V>(row);
        dy_feed(ya[0], ci, i, mk);
        if (i >= P && i < P + RY) {
          zv[0][i - P] = ci;
          xa[0][i - P] = dx_main(row, ci, xb, mk);
        }
      }
    } else if constexpr (OP == DIV) {
#pragma unroll
      for (int j = 0; j < RY; ++j) zv[0][j] = lds_vec<T, V>(tb[0] + soff + (uint32_t)(j * G::box_x(0) * sizeof(T)));
#pragma unroll
      for (int i = 0; i < RY + 2 * P; ++i)
        dy_feed(ya[0], lds_vec<T, V>(tb[1] + soff + (uint32_t)(i * G::box_x(1) * sizeof(T))), i, mk);
#pragma unroll
      for (int j = 0; j < RY; ++j) {
        const uint32_t row = tb[2] + soff + (uint32_t)(j * G::box_x(2) * sizeof(T));
        xa[0][j] = dx_main(row, lds_vec<T, V>(row), xb, mk);
      }
    } else {  // CURL
#pragma unroll
      for (int i = 0; i < RY + 2 * P; ++i) {
        const VT ci = lds_vec<T, V>(tb[2] + soff + (uint32_t)(i * G::box_x(2) * sizeof(T)));
        dy_feed(ya[0], ci, i, mk);                    // D1 x2
        if (i >= P && i < P + RY) zv[0][i - P] = ci;  // x2 -> -D0 x2 into out_1
      }
#pragma unroll
      for (int j = 0; j < RY; ++j) {
        const uint32_t row = tb[1] + soff + (uint32_t)(j * G::box_x(1) * sizeof(T));
        zv[1][j] = lds_vec<T, V>(row);      // x1 -> +D0 x1 into out_2
        xa[0][j] = dx_main(row, zv[1][j], xb, mk);  // D2 x1
      }
#pragma unroll
      for (int i = 0; i < RY + 2 * P; ++i) {
        const uint32_t row = tb[0] + soff + (uint32_t)(i * G::box_x(0) * sizeof(T));
        const VT ci = lds_vec<T, V>(row);
        dy_feed(ya[1], ci, i, mk);                                   // D1 x0
        if (i >= P && i < P + RY) xa[1][i - P] = dx_main(row, ci, xb, mk);  // D2 x0
      }
    }
  };
  // warm-up / cool-down plane: only the values that enter the z taps
  auto zonly_loads = [&](VT (&zv)[Tr::NACC][RY]) {
#pragma unroll
    for (int j = 0; j < RY; ++j) {
      if constexpr (OP == LAP || OP == GRAD) {
        zv[0][j] = lds_vec<T, V>(tb[0] + soff + (uint32_t)((P + j) * G::box_x(0) * sizeof(T)));
      } else if constexpr (OP == DIV) {
        zv[0][j] = lds_vec<T, V>(tb[0] + soff + (uint32_t)(j * G::box_x(0) * sizeof(T)));
      } else {
        zv[0][j] = lds_vec<T, V>(tb[2] + soff + (uint32_t)((P + j) * G::box_x(2) * sizeof(T)));
        zv[1][j] = lds_vec<T, V>(tb[1] + soff + (uint32_t)(j * G::box_x(1) * sizeof(T)));
      }
    }
  };
  // in-plane results: straight out (ip: GRAD out_1, out_2; CURL out_0 -- stored by the caller at plane zc), or
  // towards output plane zc of the queue (tin)
  constexpr int NIP = OP == GRAD ? 2 : (OP == CURL ? 1 : 0);
  auto combine = [&](VT (&ya)[NY][RY], VT (&xa)[NX][RY], VT (&tin)[Tr::NACC][RY], VT (&ip)[NIP + 1][RY]) {
#pragma unroll
    for (int j = 0; j < RY; ++j) {
      if constexpr (OP == LAP || OP == DIV) {
        tin[0][j] = ya[0][j];
        add_vec(tin[0][j], xa[0][j]);
      } else if constexpr (OP == GRAD) {
        ip[0][j] = ya[0][j];
        ip[1][j] = xa[0][j];
      } else {
#pragma unroll
        for (int v = 0; v < V; ++v) {
          ip[0][j].v[v] = ya[0][j].v[v] - xa[0][j].v[v];
          tin[0][j].v[v] = xa[1][j].v[v];
          tin[1][j].v[v] = -ya[1][j].v[v];
        }
      }
    }
  };
  // this warp no longer reads the stage: release it (one arrive per warp), move to the next stage
  auto release = [&]() {
    __syncwarp();
    if (lane == 0) mbar_arrive_a(bar_empty + 8u * stage);
    if (++stage == STAGES) {
      stage = 0, soff = 0, full_parity ^= 1u;
    } else {
      soff += G::STAGE_BYTES;
    }
  };
  // z taps: complete the oldest queue entry (eo: output plane zc - P, stored by the caller), shift the others
  auto ztaps = [&](VT (&zv)[Tr::NACC][RY], bool emit, VT (&eo)[Tr::NACC][RY]) {
#pragma unroll
    for (int a = 0; a < Tr::NACC; ++a) {
      const T sgn = (OP == CURL && a == 0) ? (T)-1 : (T)1;
      if (emit) {
#pragma unroll
        for (int j = 0; j < RY; ++j) fma_to(eo[a][j], sgn * A0.c[2 * P], zv[a][j], q[a][2 * P - 1][j]);
      }
#pragma unroll
      for (int i = 2 * P - 1; i >= 1; --i)
#pragma unroll
        for (int j = 0; j < RY; ++j) fma_to(q[a][i][j], sgn * A0.c[i], zv[a][j], q[a][i - 1][j]);
#pragma unroll
      for (int j = 0; j < RY; ++j) fma_vec<true>(q[a][0][j], sgn * A0.c[0], zv[a][j]);
    }
  };
  // in-plane part of plane zc joins its output (which has just received tap P)
  auto join = [&](VT (&tin)[Tr::NACC][RY]) {
    if constexpr (OP != GRAD) {
#pragma unroll
      for (int a = 0; a < Tr::NACC; ++a)
#pragma unroll
        for (int j = 0; j < RY; ++j) add_vec(q[a][P][j], tin[a][j]);
    }
  };

  // ---- specialised march: a CTA with a full tile whose chunk stays clear of the z bands knows its whole
  // schedule -- P warm-up planes, P planes that do not emit yet, the steady state, P cool-down planes --
  // so the loop carries no per-plane flag bytes, no row masks and running output pointers.  Three
  // copies: interior tiles and tiles with band rows / columns (partial tiles and the chunks that hold
  // z-band planes take the generic march below).
  const int nout = ze - zs;
  // the rim copy takes its y bands from the tile alone: every band row of this tile must find all of its taps in
  // the tile's rows [y0 - P, y0 + TY + P)
  const bool ytile_ok = prm.ydirect && (!(y0 < A1.nl) || (y0 <= P && A1.wl <= y0 + G::TY + P)) &&
                        (!(y0 + G::TY > n1 - A1.nr) || (n1 - A1.wr >= y0 - P && n1 <= y0 + G::TY + P));
  const bool fast_ok = prm.plain_ok && (x0 + G::TX <= x_hi) && (x_lo == x0) && (y0 + G::TY <= n1) && zs >= A0.nl && ze <= n0 - A0.nr &&
                       n_planes == nout + 2 * P && (!(xedge_l || xedge_r) || prm.xdirect || xcorr || (prm.dbg & 48)) &&
                       (!yedge || (ycorr ? ycorr_ok : ytile_ok) || (prm.dbg & 16));
  if (fast_ok) {
    uint64_t oz[Tr::NOUT][RY];
#pragma unroll
    for (int o = 0; o < Tr::NOUT; ++o)
#pragma unroll
      for (int j = 0; j < RY; ++j) oz[o][j] = ob[o][j] + (uint64_t)((int64_t)(zin0 - out_lag<OP, P>(o)) * (int64_t)plane_b);
    auto st = [&](int o, int j, const VT& val) { stg_vec<T, V>(oz[o][j], val); };
    int it = 0;
    // be: the copy for tiles with band rows / columns (constant at every call site).  phase: 0 = inr / emit
    // from the plane counter (one loop); 1 .. 4 = constants (warm-up, no emit yet, steady state, cool-down:
    // four loops, no flags at all)
    auto step = [&](bool be, int phase) {
      const bool xe = be, ye = be;
      const bool inr = phase == 0 ? (it >= P && it < P + nout) : (phase == 2 || phase == 3);  // also an output plane of ours
      const bool emit = phase == 0 ? (it >= 2 * P) : (phase >= 3);  // output plane zc - P is complete after this step
      if (tid == issuer(it) && it + LOOK < n_planes) issue(it + LOOK);
      mbar_wait_a(bar_full + 8u * stage, full_parity);
      VT zv[Tr::NACC][RY], tin[Tr::NACC][RY], eo[Tr::NACC][RY];
      if (inr) {
        VT ya[NY][RY], xa[NX][RY], ip[NIP + 1][RY];
        // x bands of the rim copy, single-field operators: the helper-lane form (xband_fast: the owner's RY rows go to RY
        // neighbouring lanes, ONE pass per warp) instead of RY passes over the owner's register windows -- the tile
        // runs at the pace of its slowest warp: laplacian 512^3 fp32 -3 % (divergence / curl: no difference, they keep the
        // register form).  FDX_FUSED_DBG=256 (experiments): the register form everywhere.
        constexpr bool kRimFast = OP == LAP || OP == GRAD;
        const bool rimfast = kRimFast && xe && !xcorr && xfast && !(prm.dbg & 256) && (xedge_l || xedge_r);
        hot_inplane(zv, ya, xa, xe && !rimfast, be);
        if (rimfast) {
#pragma unroll
          for (int qq = 0; qq < NX; ++qq) {
            const int f = IP::xfield(qq);
            xband_fast(tb[f] + soff + (uint32_t)(G::yoff(f) * G::box_x(f) * sizeof(T)), G::box_x(f), xa[qq]);
          }
        }
        if (xe && xcorr && !(prm.dbg & 64)) {
#pragma unroll
          for (int qq = 0; qq < NX; ++qq) {
            const int f = IP::xfield(qq);
            xcorr_apply(tb[f] + soff + (uint32_t)(G::yoff(f) * G::box_x(f) * sizeof(T)), G::box_x(f) * (int)sizeof(T), xa[qq]);
          }
        }
        if (ye) {
          // y bands of the rim copy.  Dense form with two rows per thread: a thread whose second row is a band row hands
          // it to the thread two thread-rows further down in its warp (lane ^ 16: same column) when that one has no
          // band row of its own, and gets the result back with a shuffle -- the three band rows of a warp take ONE pass
          // instead of two.  Same taps, same order: the bits of band_y_tile.  (FDX_FUSED_DBG=512: one pass per row.)
          constexpr bool kYShare = RY == 2 && TRW >= 2;
          if (kYShare && yedge && !ycorr && !(prm.dbg & (4 | 512))) {  // (CTA-uniform: the shuffles see whole warps)
            const int ygp = y0 + (ty ^ (TRW / 2)) * RY;  // first row of the partner thread
            auto bandrow = [&](int y) { return y < n1 && (y < A1.nl || y >= n1 - A1.nr); };
            const bool colok = xg >= x_lo && xg < x_hi;
            const bool own0 = colok && bandrow(yg), own1 = colok && bandrow(yg + 1);
            const bool p0 = colok && bandrow(ygp), p1 = colok && bandrow(ygp + 1);
            const int job = own0 ? yg : (p1 ? ygp + 1 : -1);  // this pass: my first row, else my partner's second row
#pragma unroll
            for (int qq = 0; qq < NY; ++qq) {
              const int f = IP::yfield(qq);
              const uint32_t tcol = tb[f] + soff - (uint32_t)(ty * RY * G::box_x(f) * (int)sizeof(T));
              VT val, got;
#pragma unroll
              for (int v = 0; v < V; ++v) val.v[v] = 0;
              if (job >= 0) val = band_y_tile(tcol, G::box_x(f) * (int)sizeof(T), job);
#pragma unroll
              for (int v = 0; v < V; ++v) got.v[v] = __shfl_xor_sync(0xffffffffu, val.v[v], 16);
              if (own0) ya[qq][0] = val;
              if (own1) ya[qq][1] = p0 ? band_y_tile(tcol, G::box_x(f) * (int)sizeof(T), yg + 1) : got;  // (partner busy: a pass of my own)
            }
          } else if (yband && !(prm.dbg & 4)) {
#pragma unroll
            for (int qq = 0; qq < NY; ++qq) {
              const int f = IP::yfield(qq);
              // (tile row 0, this thread's x) = this thread's vector in tile row ty * RY, moved up
              const uint32_t tcol = tb[f] + soff - (uint32_t)(ty * RY * G::box_x(f) * (int)sizeof(T));
#pragma unroll
              for (int j = 0; j < RY; ++j)
                if (is_yband(j)) {
                  if (ycorr) {
                    if (!(prm.dbg & 128)) ycorr_tile(tcol, G::box_x(f) * (int)sizeof(T), yg + j, ya[qq][j]);
                  }
                  else
                    ya[qq][j] = band_y_tile(tcol, G::box_x(f) * (int)sizeof(T), yg + j);
                }
            }
          }
        }
        combine(ya, xa, tin, ip);
#pragma unroll
        for (int j = 0; j < RY; ++j) {
          if constexpr (OP == GRAD) st(1, j, ip[0][j]), st(2, j, ip[1][j]);
          if constexpr (OP == CURL) st(0, j, ip[0][j]);
        }
      } else {
        zonly_loads(zv);
      }
      release();
      ztaps(zv, emit, eo);
      if (emit) {
#pragma unroll
        for (int a = 0; a < Tr::NACC; ++a)
#pragma unroll
          for (int j = 0; j < RY; ++j) st(OP == CURL ? a + 1 : 0, j, eo[a][j]);
      }
      if (inr) join(tin);
      ++it;
#pragma unroll
      for (int o = 0; o < Tr::NOUT; ++o)
#pragma unroll
        for (int j = 0; j < RY; ++j) oz[o][j] += (uint64_t)plane_b;
    };
    if ((edge && !(prm.dbg & 16)) || nout < P) {
#pragma unroll 1
      for (int k = n_planes; k > 0; --k) step(true, 0);
    } else {
#pragma unroll 1
      for (int k = 0; k < P; ++k) step(false, 1);
#pragma unroll 1
      for (int k = 0; k < P; ++k) step(false, 2);
#pragma unroll 1
      for (int k = nout - P; k > 0; --k) step(false, 3);
#pragma unroll 1
      for (int k = 0; k < P; ++k) step(false, 4);
    }
    return;
  }

  // ---- generic march: per-plane flags, band sections
  int zc = zin0;
  auto lds_flag = [&](int i) {
    uint32_t f;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(f) : "r"(flag_base + (uint32_t)i));
    return f;
  };
  // z bands in head form: output plane zo (a band plane, just completed by the queue) receives its head taps on
  // the first / last planes of the grid, gathered from global memory (L2: those planes were just streamed)
  auto zcorr_emit = [&](int zo, VT (&eo)[Tr::NACC][RY]) {
#pragma unroll
    for (int a = 0; a < Tr::NACC; ++a) {
      const int f = OP == CURL ? (a == 0 ? 2 : 1) : 0;                 // LAP / DIV / GRAD: D0 of field 0; CURL: -D0 x2, +D0 x1
      const T sgn = (OP == CURL && a == 0) ? (T)-1 : (T)1;
#pragma unroll
      for (int j = 0; j < RY; ++j) {
        if (!rok(j)) continue;
        const T* gcol0 = prm.in[f] + thread_ofs + (int64_t)j * n2;
        corr_rows(A0, zo, sgn, eo[a][j], [&](int pl) { return ld_cached<T, V>(gcol0 + (int64_t)pl * plane); }, std::true_type{});
      }
    }
  };
  uint32_t fl_next = lds_flag(0);
  for (int it = 0; it < n_planes; ++it, ++zc) {
    const uint32_t fl = fl_next;
    const bool inr = fl & F_INR;    // this plane is also an output plane of ours
    const bool zb = fl & F_ZB;      // ... next to a physical z boundary
    const bool slow = fl & F_SLOW;  // the in-plane result needs the cold section
    const bool emit = fl & F_EMIT;  // output plane zc - P is complete after this step
    if (tid == issuer(it) && it + LOOK < n_planes) issue(it + LOOK);
    mbar_wait_a(bar_full + 8u * stage, full_parity);
    fl_next = lds_flag(it + 1);  // (one byte past the last plane is still inside the flag area)

    VT zv[Tr::NACC][RY];   // values that enter the z taps
    VT tin[Tr::NACC][RY];  // in-plane contribution to output plane zc of each accumulator set
    if (inr) {
      VT ya[NY][RY], xa[NX][RY];  // raw in-plane derivatives of plane zc
      hot_inplane(zv, ya, xa, false, true);
      // ---------------- cold: rows / columns next to a physical boundary of the grid
      if (slow && edge) {
        const int64_t zofs = (int64_t)zc * plane;
        auto tile = [&](int f) { return reinterpret_cast<const T*>(smem_raw + (ring - smem0) + soff + G::field_off(f)); };
        if (yband) {
#pragma unroll
          for (int qq = 0; qq < NY; ++qq) {
            const int f = IP::yfield(qq);
#pragma unroll
            for (int j = 0; j < RY; ++j)
              if (is_yband(j)) {
                if (ycorr) {
                  if (ycorr_ok) {  // (the same loads as the rim copy)
                    ycorr_tile(tb[f] + soff - (uint32_t)(ty * RY * G::box_x(f) * (int)sizeof(T)), G::box_x(f) * (int)sizeof(T), yg + j, ya[qq][j]);
                  } else {
                    const T* gcol = prm.in[f] + zofs + xg;
                    corr_rows(A1, yg + j, (T)1, ya[qq][j], [&](int row) { return ld_cached<T, V>(gcol + (int64_t)row * n2); }, std::true_type{});
                  }
                } else {
                  ya[qq][j] = band_y(tile(f) + G::xoff(f) + tx * V, G::box_x(f), G::box_y(f), prm.in[f] + zofs + xg, yg + j, false);
                }
              }
          }
        }
        if (xedge_l || xedge_r) {
#pragma unroll
          for (int qq = 0; qq < NX; ++qq) {
            const int f = IP::xfield(qq);
            if (xcorr) {
              xcorr_apply(tb[f] + soff + (uint32_t)(G::yoff(f) * G::box_x(f) * sizeof(T)), G::box_x(f) * (int)sizeof(T), xa[qq]);
            } else if (xfast) {
              xband_fast(tb[f] + soff + (uint32_t)(G::yoff(f) * G::box_x(f) * sizeof(T)), G::box_x(f), xa[qq]);
            } else {
              xband_coop(ring + soff + G::field_off(f) + (uint32_t)(G::yoff(f) * G::box_x(f) * sizeof(T)), G::box_x(f), xa[qq]);
            }
          }
        }
      }
      VT ip[NIP + 1][RY];
      combine(ya, xa, tin, ip);
#pragma unroll
      for (int j = 0; j < RY; ++j) {
        if constexpr (OP == GRAD) store(1, zc, j, ip[0][j]), store(2, zc, j, ip[1][j]);
        if constexpr (OP == CURL) store(0, zc, j, ip[0][j]);
      }
      if (zb) {  // cold: plane next to a physical z boundary, one-sided taps gathered directly
#pragma unroll
        for (int j = 0; j < RY; ++j) {
          if (!rok(j)) continue;
          const int64_t cofs = thread_ofs + (int64_t)j * n2;
          if constexpr (OP == LAP || OP == DIV || OP == GRAD) {
            VT zz = band_z(prm.in[0] + cofs, zc);
            if constexpr (OP != GRAD) add_vec(zz, tin[0][j]);
            store(0, zc, j, zz);
          } else {
            const VT z2 = band_z(prm.in[2] + cofs, zc), z1 = band_z(prm.in[1] + cofs, zc);
            VT o1, o2;
#pragma unroll
            for (int v = 0; v < V; ++v) {
              o1.v[v] = tin[0][j].v[v] - z2.v[v];
              o2.v[v] = tin[1][j].v[v] + z1.v[v];
            }
            store(1, zc, j, o1);
            store(2, zc, j, o2);
          }
        }
      }
    } else {
      zonly_loads(zv);
    }
    release();
    if constexpr (CORR) {
      if (zcorr && (zc < A0.kcl || zc >= n0 - A0.kcr)) {  // a head plane reaches its outputs through the head table only
#pragma unroll
        for (int a = 0; a < Tr::NACC; ++a)
#pragma unroll
          for (int j = 0; j < RY; ++j)
#pragma unroll
            for (int v = 0; v < V; ++v) zv[a][j].v[v] = 0;
      }
    }
    VT eo[Tr::NACC][RY];
    ztaps(zv, emit, eo);
    if (emit) {
      if constexpr (CORR) {
        if (fl & F_EZB) zcorr_emit(zc - P, eo);
      }
#pragma unroll
      for (int a = 0; a < Tr::NACC; ++a)
#pragma unroll
        for (int j = 0; j < RY; ++j) store(OP == CURL ? a + 1 : 0, zc - P, j, eo[a][j]);
    }
    if (inr && !zb) join(tin);
  }
  if (zcorr) {
    // the last P planes of the grid leave the queue behind the last input plane: the planes beyond the grid are zero
    for (int zo = max(zs, zin1 - P + 1); zo < ze; ++zo) {
      VT eo[Tr::NACC][RY];
#pragma unroll
      for (int a = 0; a < Tr::NACC; ++a) {
#pragma unroll
        for (int j = 0; j < RY; ++j) eo[a][j] = q[a][2 * P - 1][j];
#pragma unroll
        for (int i = 2 * P - 1; i >= 1; --i)
#pragma unroll
          for (int j = 0; j < RY; ++j) q[a][i][j] = q[a][i - 1][j];
#pragma unroll
        for (int j = 0; j < RY; ++j)
#pragma unroll
          for (int v = 0; v < V; ++v) q[a][0][j].v[v] = 0;
      }
      zcorr_emit(zo, eo);
#pragma unroll
      for (int a = 0; a < Tr::NACC; ++a)
#pragma unroll
        for (int j = 0; j < RY; ++j) store(OP == CURL ? a + 1 : 0, zo, j, eo[a][j]);
    }
  }
}

// ------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

template <typename T>
static int make_tmap(CUtensorMap* m, const T* base, int64_t n2, int64_t n1, int64_t nz, int bx, int by) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return FDX_EUNSUPPORTED;
  cuuint64_t dims[3] = {(cuuint64_t)n2, (cuuint64_t)n1, (cuuint64_t)nz};
  cuuint64_t strides[2] = {(cuuint64_t)n2 * sizeof(T), (cuuint64_t)n2 * n1 * sizeof(T)};
  cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)base,
                   dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? FDX_OK : FDX_EUNSUPPORTED;
}

// (env_int / env_str: the library's snapshot of the FDX_* switches, fdx_common.cuh -- no getenv on the launch path)

struct SlabArgs {
  int z_begin, z_end, in_first, in_planes, out_first;
  int z_begin2, z_end2;  // optional second range of output planes (empty: z_begin2 == z_end2)
};

static SlabArgs slab_args(const fdx_plan_t plans[3], const fdx_slab* s) {
  if (!s) return SlabArgs{0, plans[0]->n, 0, plans[0]->n, 0, 0, 0};
  return SlabArgs{s->z_begin, s->z_end, s->in_first, s->in_planes, s->out_first, s->z_begin2, s->z_end2 > s->z_begin2 ? s->z_end2 : s->z_begin2};
}

// A plan as the radius-P kernel wants it: at least P band rows on either side.  The kernel assumes that every MAIN row
// finds all of its 2P + 1 taps inside the grid (zero taps included: the z queue completes an output plane P planes
// later, the in-plane halos are P wide).  A one-sided or shorter stencil has fewer band rows than P (method="backward":
// none on the right); the rows in between become band rows that carry the main stencil as dense taps -- the same sums.
struct PaddedPlan {
  fdx_plan_s p;  // h_band_l / h_band_r point into the vectors below; band_l / band_r (device) are unused here
  std::vector<double> bl, br;
};
static bool pad_plan(PaddedPlan& out, const fdx_plan_s* src, int P) {
  out.p = *src;
  const int n = src->n, lo = src->lo, hi = src->hi, taps = hi - lo + 1;
  const int nl = std::max(src->nl, std::min(P, n)), nr = std::max(src->nr, std::min(P, n));
  const int wl = std::max(src->wl, nl + hi), wr = std::max(src->wr, nr - lo);
  if (nl + nr > n || wl > n || wr > n) return false;
  out.bl.assign((size_t)nl * wl, 0.0), out.br.assign((size_t)nr * wr, 0.0);
  for (int r = 0; r < nl; ++r)
    for (int c = 0; c < wl; ++c) {
      double v = 0.0;
      if (r < src->nl)
        v = c < src->wl ? src->h_band_l[r * src->wl + c] : 0.0;
      else if (c - r >= lo && c - r <= hi)
        v = src->main_taps[c - r - lo];
      out.bl[(size_t)r * wl + c] = v;
    }
  for (int r = 0; r < nr; ++r) {  // row n - nr + r, columns n - wr + c
    const int sr = r - (nr - src->nr);  // row of the source's right band (negative: a main row)
    for (int c = 0; c < wr; ++c) {
      double v = 0.0;
      if (sr >= 0) {
        const int sc = c - (wr - src->wr);
        v = sc >= 0 ? src->h_band_r[sr * src->wr + sc] : 0.0;
      } else {
        const int o = (wr - nr) + r;  // column of the row's own element
        if (c - o >= lo && c - o <= hi) v = src->main_taps[c - o - lo];
      }
      out.br[(size_t)r * wr + c] = v;
    }
  }
  (void)taps;
  out.p.nl = nl, out.p.wl = wl, out.p.nr = nr, out.p.wr = wr;
  out.p.h_band_l = out.bl.data(), out.p.h_band_r = out.br.data();
  return true;
}

// Head form of a plan's bands.  A TRANSPOSED plan (the adjoints: P + L band rows of up to 2L + 1 taps) differs from the
// radius-P main stencil only on its first / last few columns -- the columns that belong to the forward plan's one-sided
// rows: band row r = [main stencil over the columns beyond the head] + sum over the kc head columns of band[r][c] x[c].
// The kernel evaluates it that way: the main path runs on the band rows too, with the head elements hidden from it
// (masks; elements outside the grid are zero fill anyway), and a kc-tap head table adds the rest -- the same products
// as the dense row, tail first.  x-edge tiles stay on the specialised march, z-band planes gather kc planes instead of
// up to 2L + 1.  The reference's forward plans (one-sided rows: as wide as the rows) keep the dense form.
// Conditions: every row that reaches a head column through the main stencil is a band row (kc + P <= nl, so that
// hiding the head from the main path changes no main row), and the dense rows do not extend past the main stencil's
// reach.  main / band are compared as T values scaled by alpha, so that rows that merely repeat the main stencil
// (pad_plan) count as equal.  kcl / kcr: columns [0, kcl) and [n - kcr, n) form the heads.
template <typename T>
static bool corr_form(const fdx_plan_s* p, int P, double alpha, int& kcl, int& kcr) {
  kcl = kcr = 0;
  if (env_int("FDX_FUSED_NOCORR", 0)) return false;
  auto main_t = [&](int o) { return (o >= p->lo && o <= p->hi && o >= -P && o <= P) ? (T)(p->main_taps[o - p->lo] * alpha) : (T)0; };
  for (int r = 0; r < p->nl; ++r)
    for (int c = 0; c < p->wl; ++c)
      if ((T)(p->h_band_l[r * p->wl + c] * alpha) != main_t(c - r)) kcl = std::max(kcl, c + 1);
  for (int r = 0; r < p->nr; ++r)
    for (int c = 0; c < p->wr; ++c)
      if ((T)(p->h_band_r[r * p->wr + c] * alpha) != main_t((c - p->wr) - (r - p->nr))) kcr = std::max(kcr, p->wr - c);
  // the main stencil of row r reaches columns r - P .. r + P, the dense row covers [0, wl): beyond it the main stencil
  // must have no taps
  for (int r = 0; r < p->nl; ++r)
    for (int o = -P; o <= P; ++o)
      if (r + o >= p->wl && main_t(o) != (T)0) return false;
  for (int r = 0; r < p->nr; ++r)  // row n - nr + r, dense columns [n - wr, n)
    for (int o = -P; o <= P; ++o)
      if ((p->n - p->nr + r) + o < p->n - p->wr && main_t(o) != (T)0) return false;
  const int kmax = std::max(kcl, kcr), wmax = std::max(p->wl, p->wr);
  if (kmax == 0) return false;
  constexpr int V = 16 / (int)sizeof(T);
  const int nlp = ((p->nl + V - 1) / V) * V, nrp = ((p->nr + V - 1) / V) * V;
  if (kcl + P > p->nl || kcr + P > p->nr) return false;
  return kmax <= 8 && 2 * kmax <= wmax && nlp + nrp <= p->n;
}

// Everything of a launch that does not depend on the data pointers: band tables, x-band layouts, tile counts, the
// list of z chunks.  Cached per (plans, alpha, slab) by launch_cfg.  prm.n_chunks == 0 on return: nothing to do.
template <Op OP, typename T, int P, typename CFG>
static int build_params(FusedParams<T, P>& prm, const fdx_plan_t plans_in[3], const double alpha[3], const SlabArgs& sl, bool allow_corr) {
  using G = Geometry<OP, T, P, CFG>;
  PaddedPlan padded[3];
  const fdx_plan_s* plans[3];
  for (int k = 0; k < 3; ++k) {
    if (!pad_plan(padded[k], plans_in[k], P)) return FDX_EUNSUPPORTED;
    plans[k] = &padded[k].p;
  }
  const int n0 = plans[0]->n, n1 = plans[1]->n, n2 = plans[2]->n;
  (void)n1;
  prm.n_chunks = 0;
  int tab_used = 0, rows_used = 0;
  for (int k = 0; k < 3; ++k) {
    const fdx_plan_s* p = plans[k];
    AxisDev<T>& a = prm.ax[k];
    a.n = p->n, a.nl = p->nl, a.wl = p->wl, a.nr = p->nr, a.wr = p->wr;
    a.alpha = (T)alpha[k];
    for (int t = 0; t < FDX_MAX_TAPS; ++t) {  // tap t of the kernel sits at offset t - P; the plan's taps cover lo .. hi
      const int o = t - P;
      a.c[t] = (t < 2 * P + 1 && o >= p->lo && o <= p->hi) ? (T)(p->main_taps[o - p->lo] * alpha[k]) : (T)0;
    }
    a.corr = (allow_corr && corr_form<T>(p, P, alpha[k], a.kcl, a.kcr)) ? 1 : 0;
    a.nlp = ((p->nl + G::V - 1) / G::V) * G::V, a.nrp = ((p->nr + G::V - 1) / G::V) * G::V;
    if (a.corr) {
      // head[c][row] tables, alpha folded in: left rows 0 .. nlp - 1 and columns 0 .. kcl - 1, right rows n - nrp .. n - 1
      // and columns n - kcr .. n - 1 (rows that are not band rows: zeros)
      if (tab_used + a.kcl * a.nlp + a.kcr * a.nrp > tab_max(P)) return FDX_EUNSUPPORTED;
      a.tab_l = tab_used, a.rng_l = a.rng_r = 0;
      for (int c = 0; c < a.kcl; ++c)
        for (int r = 0; r < a.nlp; ++r) prm.tab[tab_used++] = (r < p->nl && c < p->wl) ? (T)(p->h_band_l[r * p->wl + c] * alpha[k]) : (T)0;
      a.tab_r = tab_used;
      for (int c = 0; c < a.kcr; ++c)      // global column n - kcr + c
        for (int r = 0; r < a.nrp; ++r) {  // global row n - nrp + r
          const int br = r - (a.nrp - p->nr), bc = p->wr - a.kcr + c;  // row / column inside the dense right band
          prm.tab[tab_used++] = (br >= 0 && bc >= 0) ? (T)(p->h_band_r[br * p->wr + bc] * alpha[k]) : (T)0;
        }
      continue;
    }
    for (int side = 0; side < 2; ++side) {
      const int rows = side ? p->nr : p->nl, w = side ? p->wr : p->wl;
      const double* h = side ? p->h_band_r : p->h_band_l;
      if (tab_used + rows * w > tab_max(P) || rows_used + rows > kRowsMax) return FDX_EUNSUPPORTED;
      (side ? a.tab_r : a.tab_l) = tab_used;
      (side ? a.rng_r : a.rng_l) = rows_used;
      for (int r = 0; r < rows; ++r) {
        int c0 = w, c1 = 0;
        for (int c = 0; c < w; ++c) {
          prm.tab[tab_used + r * w + c] = (T)h[r * w + c];
          if (h[r * w + c] != 0.0) {
            if (c < c0) c0 = c;
            c1 = c + 1;
          }
        }
        if (c1 <= c0) c0 = c1 = 0;
        prm.rng[rows_used + r] = make_short2((short)c0, (short)c1);
      }
      tab_used += rows * w;
      rows_used += rows;
    }
  }
  {
    // fast x bands: rows in one vector, row r's taps inside columns [r, r + nt)
    constexpr int XV = FusedParams<T, P>::kXV, XT = FusedParams<T, P>::kXT;
    const fdx_plan_s* px = plans[2];
    const int ntl = px->wl - px->nl + 1, ntr = px->wr - px->nr + 1;
    bool ok = px->nl <= XV && px->nr <= XV && ntl >= 1 && ntr >= 1 && ntl <= XT && ntr <= XT && n2 >= 3 * XV;
    for (int t = 0; t < XT; ++t)
      for (int v = 0; v < XV; ++v) prm.xcl[t][v] = prm.xcr[t][v] = (T)0;
    for (int r = 0; ok && r < px->nl; ++r)
      for (int c = 0; c < px->wl; ++c) {
        const double h = px->h_band_l[r * px->wl + c];
        if (c >= r && c < r + ntl)
          prm.xcl[c - r][r] = (T)h;
        else if (h != 0.0)
          ok = false;
      }
    for (int r = 0; ok && r < px->nr; ++r)
      for (int c = 0; c < px->wr; ++c) {
        const double h = px->h_band_r[r * px->wr + c];
        if (c >= r && c < r + ntr)
          prm.xcr[r + ntr - 1 - c][XV - px->nr + r] = (T)h;
        else if (h != 0.0)
          ok = false;
      }
    // the last tile must hold the three vectors that end at column n2 - 1
    const int xl_last = n2 > G::TX ? n2 - G::TX : 0;  // origin of the (shifted) last tile
    if (n2 - xl_last < 3 * XV - G::HX) ok = false;
    if (((n2 - xl_last) / XV - 1) % CFG::LX < CFG::RY - 1) ok = false;  // the helper lanes sit in the owner's warp
    if (env_int("FDX_XBAND_SLOW", 0)) ok = false;
    prm.xfast = ok ? 1 : 0, prm.xnt_l = ntl, prm.xnt_r = ntr;
    prm.plain_ok = env_int("FDX_FUSED_NOPLAIN", 0) ? 0 : 1;
    prm.issue_rot = env_int("FDX_ISSUE_ROT", 1);
    prm.l2_ahead = env_int("FDX_L2_AHEAD", 0);
    {
      const fdx_plan_s* py = plans[1];
      prm.ydirect = (py->nl <= G::TY && py->nr <= G::TY && py->wl <= G::TY + P && py->wr <= G::TY + P) ? 1 : 0;
    }
    prm.dbg = env_int("FDX_FUSED_EXPERIMENTS", 0) ? env_int("FDX_FUSED_DBG", 0) : 0;  // (wrong results at the rim: timing experiments only)
    prm.rim_first = -1;  // decided below, once the tile counts are known
    // direct x bands: every tap of every band row lies inside the owner's window of 2 HXV + 1 vectors
    prm.xdirect = (ok && (px->nl - 1) + (ntl - 1) <= G::HX + XV - 1 && (px->nr - 1) + (ntr - 1) <= G::HX + XV - 1 &&
                   !env_int("FDX_XBAND_NODIRECT", 0)) ? 1 : 0;
  }
  prm.z_begin = sl.z_begin, prm.z_end = sl.z_end;
  prm.tz_off = -sl.in_first;
  prm.z_lo = 0, prm.z_hi = n0 - 1;
  prm.tiles_x = (n2 + G::TX - 1) / G::TX;
  prm.tiles_y = (n1 + G::TY - 1) / G::TY;
  {
    // the x bands are evaluated from the first / last tile in shared memory: those tiles must hold
    // every band output of their side and all of its taps (otherwise: composition fallback)
    const fdx_plan_s* px = plans[2];
    const int xl = n2 > G::TX ? n2 - G::TX : 0;  // origin of the (shifted) last tile
    if (px->nl > G::TX || px->wl > G::TX + G::HX || n2 - xl < px->nr || xl > n2 - px->wr + G::HX) return FDX_EUNSUPPORTED;
    if ((prm.ax[2].corr ? prm.ax[2].kcl * prm.ax[2].nlp + prm.ax[2].kcr * prm.ax[2].nrp : px->nl * px->wl + px->nr * px->wr) > xtab_max(P))
      return FDX_EUNSUPPORTED;
    if (prm.tiles_x == 2) {  // the cut between the first tile and the shifted second one (see the kernel): the right band
      const int cut = std::min(G::TX, std::max(xl, ((px->nl + G::V - 1) / G::V) * G::V));  // must lie behind it
      if (n2 - px->nr < cut) return FDX_EUNSUPPORTED;
    }
  }
  const int64_t tiles = (int64_t)prm.tiles_x * prm.tiles_y;
  // rim tiles first only while a layer of tiles fits the machine (measured: +3 % at 512^3 = 256 tiles per layer,
  // -3 % at 1024^3 = 1024 tiles, where launching the rim apart from its neighbours costs L2 halo sharing)
  prm.rim_first = env_int("FDX_FUSED_RIMFIRST", tiles <= (int64_t)num_sms() * CFG::MINB ? 1 : 0);
  // z chunking.  The planes next to a physical z boundary (one-sided taps, gathered directly) form
  // two chunks of their own, launched first, so that every other chunk is free of them and can take
  // the specialised march.  Every chunk pays 2P warm-up planes, so long chunks are cheaper, but the
  // launch ends when its last CTA does: the chunks are graded -- long ones first, then medium, and
  // the CTAs of the last ~1.5 waves are short -- which keeps both the redundancy and the tail small.
  {
    const bool zsplit = !env_int("FDX_FUSED_ZB_MERGE", 0);
    const int zlo_end = zsplit ? std::min(sl.z_end, std::max(sl.z_begin, plans[0]->nl)) : sl.z_begin;     // [z_begin, zlo_end): left band planes
    const int zhi_beg = zsplit ? std::max(zlo_end, std::min(sl.z_end, n0 - plans[0]->nr)) : sl.z_end;  // [zhi_beg, z_end): right band planes
    const int nz = zhi_beg - zlo_end;                                                     // planes of the graded middle part
    const int64_t slots = (int64_t)num_sms() * CFG::MINB;
    int L = env_int("FDX_FUSED_CHUNK", 0), S = env_int("FDX_FUSED_CHUNK_SHORT", 0);
    const bool uniform = L > 0 && S <= 0;  // an explicit chunk alone means uniform chunks (sweeps)
    if (L <= 0) {
      // long chunks: 2P warm-up planes per chunk against the imbalance of long CTAs near the end of the launch.
      // L* = sqrt(2P nz tiles / (4 x 148)) fits the B200 sweeps (512^3: 40 at P = 3, 32 at P = 2; 1024^3: the cap)
      const int cap = 64 > 8 * P ? 64 : 8 * P;
      const double lstar = std::sqrt(2.0 * P * (double)(nz > 0 ? nz : 1) * (double)tiles / (4.0 * num_sms()));
      L = ((int)(lstar + 2.0) / 4) * 4;
      if (L > cap) L = cap;
      if (L < 16) L = 16;
      if (L < 4 * P) L = 4 * P;
    }
    if (S <= 0) S = 16 > 4 * P ? 16 : 4 * P;
    if (uniform) S = L;
    if (L > kMaxPlanes - 2 * P) L = kMaxPlanes - 2 * P;
    if (S > L) S = L;
    if ((int64_t)nz > (int64_t)L * (kMaxCuts - 30)) L = (int)((nz + kMaxCuts - 31) / (kMaxCuts - 30));
    if (L > kMaxPlanes - 2 * P) return FDX_EUNSUPPORTED;
    // small grids: keep at least ~2 waves of CTAs
    while (!uniform && L > S && tiles * ((nz + L - 1) / L) < 2 * slots) L = L / 2 > S ? L / 2 : S;
    // ... and tiny grids are latency-bound (a march is a serial chain of ~1 us per plane, measured
    // 35 -> 16 us for the README's 100^3 divergence): many very short chunks until two waves are filled
    if (!uniform && L == S)
      while (S > 4 && tiles * ((nz + S - 1) / S) < 2 * slots) S = L = (S / 2 > 4 ? S / 2 : 4);
    // ... but never more chunks than the cut table holds (few tiles and a long z axis, e.g. 512 x 80 x 192: the
    // launch used to be refused and the call fell back to the three-pass composition)
    if ((int64_t)nz > (int64_t)L * (kMaxCuts - 30)) L = (int)((nz + kMaxCuts - 31) / (kMaxCuts - 30));
    if (S > L) S = L;
    // sizes from the end of the middle range backwards: short layers, as many medium ones, then long ones
    int sizes[kMaxCuts];
    int n = 0, rem = nz;
    int n_short = (int)((3 * slots + 2 * tiles - 1) / (2 * tiles));  // ~1.5 waves of CTAs
    if (n_short > 12) n_short = 12;
    if (uniform || L == S) n_short = 0;
    for (int i = 0; i < n_short && rem > nz / 2 && rem > S; ++i) sizes[n++] = S, rem -= S;
    const int M = 2 * S < L ? 2 * S : L;
    for (int i = 0; i < n_short && M < L && rem > nz / 2 && rem > M; ++i) sizes[n++] = M, rem -= M;
    if (rem > 0) {
      const int nl = (rem + L - 1) / L;
      for (int i = 0; i < nl; ++i) {
        const int sz = (rem + (nl - i) - 1) / (nl - i);
        sizes[n++] = sz, rem -= sz;
      }
    }
    if (const char* cuts = env_str("FDX_FUSED_CUTS")) {  // experiments: explicit sizes, last z chunk first
      n = 0, rem = nz;
      for (const char* c = cuts; *c && rem > 0 && n < kMaxCuts - 3;) {
        int v = std::atoi(c);
        if (v <= 0) break;
        if (v > kMaxPlanes - 2 * P) v = kMaxPlanes - 2 * P;
        if (v > rem) v = rem;
        sizes[n++] = v, rem -= v;
        while (*c && *c != ',') ++c;
        if (*c == ',') ++c;
      }
      while (rem > 0 && n < kMaxCuts - 2) {
        const int v = rem < L ? rem : L;
        sizes[n++] = v, rem -= v;
      }
      if (rem > 0) return FDX_EUNSUPPORTED;
    }
    if (n > kMaxCuts - 2) return FDX_EUNSUPPORTED;
    // launch order: the middle in z order, long to short, then the two band chunks (zb_last)
    int k = 0;
    // the two chunks of z-band planes (few planes, direct gathers: short but slow CTAs) go to the END of the launch, behind
    // the graded middle chunks -- they fill the tail (B200 A/B: laplacian 512^3 fp32 -0.9 %, fp64 accuracy 8 -3.7 %,
    // 256^3 -1.8 %, the other operators -0.4 % or neutral); FDX_FUSED_ZB_LAST=0: in front, as in round 1
    const bool zb_last = env_int("FDX_FUSED_ZB_LAST", 1) != 0;
    if (!zb_last) {
      if (zlo_end > sl.z_begin) prm.cut_z[k] = sl.z_begin, prm.cut_n[k] = (short)(zlo_end - sl.z_begin), ++k;
      if (sl.z_end > zhi_beg) prm.cut_z[k] = zhi_beg, prm.cut_n[k] = (short)(sl.z_end - zhi_beg), ++k;
    }
    int z = zlo_end;
    if (!zsplit && n > 1) {  // (experiment) the old order: first and last chunk first
      prm.cut_z[k] = zlo_end, prm.cut_n[k] = (short)sizes[n - 1], ++k;
      prm.cut_z[k] = zhi_beg - sizes[0], prm.cut_n[k] = (short)sizes[0], ++k;
      z += sizes[n - 1];
      for (int i = n - 2; i >= 1; --i) prm.cut_z[k] = z, prm.cut_n[k] = (short)sizes[i], ++k, z += sizes[i];
    } else
    for (int i = n - 1; i >= 0; --i) prm.cut_z[k] = z, prm.cut_n[k] = (short)sizes[i], ++k, z += sizes[i];
    if (zb_last) {
      if (zlo_end > sl.z_begin) prm.cut_z[k] = sl.z_begin, prm.cut_n[k] = (short)(zlo_end - sl.z_begin), ++k;
      if (sl.z_end > zhi_beg) prm.cut_z[k] = zhi_beg, prm.cut_n[k] = (short)(sl.z_end - zhi_beg), ++k;
    }
    // the optional second range: pieces of at most L planes behind the first range's chunks (any chunking gives the
    // same bits; a piece that holds z-band planes takes the generic march)
    for (int z2 = sl.z_begin2; z2 < sl.z_end2; z2 += L) {
      if (k >= kMaxCuts) return FDX_EUNSUPPORTED;
      prm.cut_z[k] = z2, prm.cut_n[k] = (short)std::min(L, sl.z_end2 - z2), ++k;
    }
    if (k == 0) return FDX_OK;  // nothing to do
    prm.n_chunks = k;
  }
  if (tiles * prm.n_chunks > 0x7fffffffLL) return FDX_EUNSUPPORTED;
  return FDX_OK;
}

// ALLOW_CORR: this tile configuration also exists as the CORR instantiation of the kernel (the default configurations
// only: the alternates of the sweeps and the wide-band shape keep the dense bands, and the build its size)
template <Op OP, typename T, int P, typename CFG, bool ALLOW_CORR = false>
static int launch_cfg(const fdx_plan_t plans[3], const T* const in[3], T* const out[3], const double alpha[3],
                      const SlabArgs& sl, cudaStream_t st) {
  using G = Geometry<OP, T, P, CFG>;
  using Tr = OpTraits<OP>;
  const int n1 = plans[1]->n, n2 = plans[2]->n;
  FusedParams<T, P> prm;
  int rc;
  {
    // the pointer-independent part of the parameter block, cached (small grids are bound by the host launch path);
    // an entry dies with the library epoch: any plan destroyed (its address may be reused) or the switches re-read
    struct Entry {
      uint64_t epoch;
      const fdx_plan_s* p[3];
      double a[3];
      SlabArgs sl;
      int rc;
      FusedParams<T, P> prm;
    };
    constexpr int kEntries = 4;
    static Entry cache[kEntries];
    static int used = 0, next = 0;
    static std::mutex mu;
    const uint64_t epoch = g_epoch.load(std::memory_order_acquire);
    std::lock_guard<std::mutex> lk(mu);
    const Entry* hit = nullptr;
    for (int i = 0; i < used && !hit; ++i) {
      const Entry& e = cache[i];
      if (e.epoch == epoch && e.p[0] == plans[0] && e.p[1] == plans[1] && e.p[2] == plans[2] && e.a[0] == alpha[0] && e.a[1] == alpha[1] &&
          e.a[2] == alpha[2] && std::memcmp(&e.sl, &sl, sizeof(SlabArgs)) == 0)
        hit = &e;
    }
    if (!hit) {
      Entry& e = cache[next];
      next = (next + 1) % kEntries;
      if (used < kEntries) ++used;
      e.epoch = epoch;
      for (int k = 0; k < 3; ++k) e.p[k] = plans[k], e.a[k] = alpha[k];
      e.sl = sl;
      e.rc = build_params<OP, T, P, CFG>(e.prm, plans, alpha, sl, ALLOW_CORR);
      hit = &e;
    }
    rc = hit->rc;
    if (rc == FDX_OK) prm = hit->prm;
  }
  if (rc != FDX_OK) return rc;
  if (prm.n_chunks == 0) return FDX_OK;  // nothing to do
  // plain loads/stores address planes by their global index through pointers to (virtual) plane 0;
  // only addresses inside the caller's window are ever dereferenced.  TMA gets the real base.
  const int64_t plane_elems = (int64_t)n1 * n2;
  for (int f = 0; f < 3; ++f) {
    prm.in[f] = f < Tr::NIN ? in[f] - (int64_t)sl.in_first * plane_elems : nullptr;
    prm.out[f] = f < Tr::NOUT ? out[f] - (int64_t)sl.out_first * plane_elems : nullptr;
  }
  for (int f = 0; f < Tr::NIN; ++f) {
    int rc2 = make_tmap<T>(&prm.tmap[f], in[f], n2, n1, (int64_t)sl.in_planes, G::box_x(f), G::box_y(f));
    if (rc2) return rc2;
  }
  const int64_t blocks = (int64_t)prm.tiles_x * prm.tiles_y * prm.n_chunks;
  const bool corr = ALLOW_CORR && (prm.ax[0].corr | prm.ax[1].corr | prm.ax[2].corr) != 0;
  auto go = [&](auto kern, std::atomic<unsigned long long>& attr_done) -> int {
    // once per kernel instantiation and device (the attribute is per device; ordinals beyond 63 set it at every launch)
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = (dev >= 0 && dev < 64) ? (1ull << dev) : 0ull;
    if (!bit || !(attr_done.load(std::memory_order_acquire) & bit)) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES);
      if (e != cudaSuccess) return (int)e;
      attr_done.fetch_or(bit, std::memory_order_release);
    }
    kern<<<(unsigned)blocks, G::NT, G::SMEM_BYTES, st>>>(prm);
    return FDX_OK;
  };
  static std::atomic<unsigned long long> attr_plain{0}, attr_corr{0};
  int rc_l;
  if constexpr (ALLOW_CORR)
    rc_l = corr ? go(fused3d_kernel<OP, T, P, CFG, true>, attr_corr) : go(fused3d_kernel<OP, T, P, CFG, false>, attr_plain);
  else
    rc_l = go(fused3d_kernel<OP, T, P, CFG, false>, attr_plain);
  if (rc_l != FDX_OK) return rc_l;
  count_launch(OP == LAP ? "fused_laplacian3d" : OP == DIV ? "fused_divergence3d" : OP == GRAD ? "fused_gradient3d" : "fused_curl3d");
  FDX_LAUNCH_CHECK();
  return FDX_OK;
}

template <Op OP, typename T, int P>
static int launch_p(const fdx_plan_t plans[3], const T* const in[3], T* const out[3], const double alpha[3],
                    const SlabArgs& sl, cudaStream_t st) {
  // tile configurations: Cfg<lanes in x, thread rows, rows per thread, stages, min CTAs per SM>
  const int cfg = env_int("FDX_FUSED_CFG", 0);
#define FDX_CFG(ID, ...) \
  case ID:               \
    return launch_cfg<OP, T, P, Cfg<__VA_ARGS__>>(plans, in, out, alpha, sl, st);
#ifdef FDX_SWEEP
#ifndef FDX_SWEEP_TSIZE
#define FDX_SWEEP_TSIZE 4
#endif
#ifndef FDX_SWEEP_P
#define FDX_SWEEP_P 3
#endif
  if constexpr (sizeof(T) == FDX_SWEEP_TSIZE && P == FDX_SWEEP_P) {
    switch (cfg) {
      // round-2 sweep list (tools/sweep_fused.py): Cfg<lanes x, thread rows, rows/thread, stages, min CTAs/SM, lanes x per warp>
      FDX_CFG(1, 16, 8, 2, 8, 4, 8)    // the fp32 laplacian default: 64x16 tile, 128 threads
      FDX_CFG(2, 16, 8, 2, 8, 4, 4)    // warps 4 lanes wide x 8 thread rows: the x-band work of an x-edge tile in ONE warp
      FDX_CFG(3, 16, 8, 2, 8, 4, 16)   // full-width warps
      FDX_CFG(4, 32, 4, 2, 8, 4, 8)    // 128x8 tile: half the TMA box rows per point
      FDX_CFG(5, 32, 8, 2, 8, 2, 8)    // 128x16 tile, 256 threads
      FDX_CFG(6, 16, 16, 2, 8, 2, 8)   // 64x32 tile, 256 threads
      FDX_CFG(7, 16, 8, 2, 6, 4, 8)    // 6 stages
      FDX_CFG(8, 8, 16, 2, 8, 4, 8)    // 32x32 tile
      FDX_CFG(9, 16, 8, 2, 4, 3)       // the generic default
      FDX_CFG(10, 16, 16, 1, 8, 4, 8)  // one row per thread, 256 threads, 64 registers
      FDX_CFG(11, 32, 4, 2, 8, 4, 16)
      FDX_CFG(12, 16, 4, 2, 8, 8, 8)   // 64x8 tile, 64 threads, 8 CTAs per SM
      // fp64 wide stencils (FDX_SWEEP_TSIZE=8 FDX_SWEEP_P=5): deeper rings, fewer / more CTAs per SM, narrow warps
      FDX_CFG(13, 16, 8, 2, 6, 3)
      FDX_CFG(14, 16, 8, 2, 7, 3)
      FDX_CFG(15, 16, 8, 2, 5, 3)
      FDX_CFG(16, 16, 8, 2, 6, 2)
      FDX_CFG(17, 16, 8, 2, 4, 3, 8)
      FDX_CFG(18, 32, 4, 2, 4, 3)
      FDX_CFG(19, 16, 8, 2, 4, 2)
      FDX_CFG(20, 16, 4, 2, 6, 5)
      FDX_CFG(21, 16, 4, 2, 6, 6)
      FDX_CFG(22, 8, 16, 2, 4, 3)
      FDX_CFG(23, 16, 8, 2, 8, 2)
      FDX_CFG(24, 16, 8, 2, 6, 5, 8)   // fp32: 5 CTAs per SM (96 registers)
      FDX_CFG(25, 16, 8, 2, 5, 6, 8)   // fp32: 6 CTAs per SM (80 registers)
      // fp64 wide stencils, four rows per thread (14 row loads per 4 vectors instead of 12 per 2): 160 queue registers
      FDX_CFG(26, 16, 8, 4, 4, 2)      // 64x32 tile, 128 threads, 2 CTAs per SM
      FDX_CFG(27, 16, 4, 4, 3, 4)      // 64x16 tile, 64 threads, 4 CTAs per SM
      FDX_CFG(28, 16, 8, 4, 3, 2)
      FDX_CFG(29, 16, 4, 4, 4, 3)
      // fp32: three CTAs per SM with deep rings (how much do 12 instead of 16 warps cost?)
      FDX_CFG(30, 16, 8, 2, 8, 3, 8)
      FDX_CFG(32, 16, 8, 2, 8, 3, 8, 1)  // ... with a producer warp
      default:
        break;
    }
  }
#endif
  // defaults from the sweeps on B200 (tools/sweep_fused.py, profiles/): 64x16-point tiles of 128
  // threads everywhere; the fp32 laplacian (most arithmetic per byte) runs 4 CTAs per SM with 8 TMA
  // stages and narrow warps, curl (two queues, three inputs) one row per thread
  if constexpr (OP == CURL) {
    switch (cfg) {
      FDX_CFG(1, 16, 8, 2, 4, 3)
      default:
        // two register queues: from radius 4 on they need more than 128 registers
        return launch_cfg<OP, T, P, Cfg<16, 16, 1, 4, (P <= 3 ? 2 : 1)>, true>(plans, in, out, alpha, sl, st);
    }
  } else if constexpr (OP == LAP && sizeof(T) == 4 && P <= 3) {
    // wide x bands (transposed plans: P + L band columns) go through the warp-cooperative path, whose cost per warp
    // grows with the rows a warp covers: full-width warps (LX = 16, 4 rows) halve it
    // (bands in head form cost the same at any width: they keep the default shape)
    bool wide = plans[2]->nl > 16 / (int)sizeof(T) || plans[2]->nr > 16 / (int)sizeof(T);
    if (wide) {
      PaddedPlan pp;
      int kl, kr;
      if (pad_plan(pp, plans[2], P) && corr_form<T>(&pp.p, P, alpha[2], kl, kr)) wide = false;
    }
    switch (cfg == 0 && wide && !env_int("FDX_FUSED_NOWIDECFG", 0) ? 2 : cfg) {
      FDX_CFG(1, 16, 8, 2, 4, 3)
      FDX_CFG(2, 16, 8, 2, 8, 4, 16)
      default:
        return launch_cfg<OP, T, P, Cfg<16, 8, 2, 8, 4, 8>, true>(plans, in, out, alpha, sl, st);
    }
  } else {
    switch (cfg) {
      FDX_CFG(1, 16, 16, 1, 4, 2)
      default:
        return launch_cfg<OP, T, P, Cfg<16, 8, 2, 4, 3>, true>(plans, in, out, alpha, sl, st);
    }
  }
#undef FDX_CFG
}

// Can the TMA kernel take this call?  (otherwise: composition of axis kernels)
template <Op OP, typename T>
static int fused_radius(const fdx_plan_t plans[3], const T* const in[3], T* const out[3], const SlabArgs& sl) {
  using Tr = OpTraits<OP>;
  constexpr int V = 16 / sizeof(T);
  if (env_int("FDX_FUSED_DISABLE", 0)) return 0;
  // The kernel radius is the largest reach of any axis' main stencil.  An axis with a shorter or one-sided stencil
  // (per-axis accuracy, method = "forward" / "backward": offsets 0..L or -L..0, finite_diff.py:118-165) runs as a
  // radius-P stencil whose missing taps are zero: the same sums, the taps outside the grid multiply TMA's zero fill.
  int P = 0;
  for (int k = 0; k < 3; ++k) {
    const fdx_plan_s* p = plans[k];
    if (p->lo > 0 || p->hi < 0) return 0;
    P = std::max(P, std::max(-p->lo, p->hi));
    if (p->nl < -p->lo || p->nr < p->hi) return 0;  // every main row must reach its (non-zero) taps inside the grid
    if (p->nl + p->nr > p->n) return 0;              // dense (tiny-n) plans
    if (p->n < 1) return 0;
  }
  if (P < 1 || P > 6) return 0;
  if (plans[2]->n % V != 0) return 0;
  if ((int64_t)plans[1]->n * plans[2]->n * (int64_t)sizeof(T) > 0x7fffffffLL) return 0;  // 32-bit plane stride (bytes) in the kernel
  for (int f = 0; f < Tr::NIN; ++f)
    if (((uintptr_t)in[f] & 15) != 0) return 0;
  for (int f = 0; f < Tr::NOUT; ++f)
    if (((uintptr_t)out[f] & 15) != 0) return 0;
  if (sl.z_begin < 0 || sl.z_end > plans[0]->n || sl.z_begin >= sl.z_end) return 0;
  if (sl.z_end2 > sl.z_begin2 && (sl.z_begin2 < 0 || sl.z_end2 > plans[0]->n || (sl.z_begin2 < sl.z_end && sl.z_begin < sl.z_end2))) return -1;  // overlap
  for (int r = 0; r < 2; ++r) {
    // the planes this call reads must lie inside the caller's window
    const int zb = r ? sl.z_begin2 : sl.z_begin, ze = r ? sl.z_end2 : sl.z_end;
    if (ze <= zb) continue;
    const fdx_plan_s* p0 = plans[0];
    const int nl = std::max(p0->nl, P), nr = std::max(p0->nr, P);  // (the kernel's band rows: see pad_plan)
    const int wl = std::max(p0->wl, nl + p0->hi), wr = std::max(p0->wr, nr - p0->lo);
    int lo = zb - P < 0 ? 0 : zb - P;
    int hi = ze + P > p0->n ? p0->n : ze + P;
    if (zb < nl) lo = 0, hi = hi > wl ? hi : wl;
    if (ze > p0->n - nr) lo = lo < p0->n - wr ? lo : p0->n - wr, hi = p0->n;
    if (lo < sl.in_first || hi > sl.in_first + sl.in_planes) return -1;
    if (zb < sl.out_first) return -1;
  }
  if (!encode_fn()) return 0;
  return P;
}

template <typename T>
static int axis_apply(fdx_plan_t p, const T* x, T* out, int64_t outer, int64_t inner, double alpha, int acc, void* st) {
  if constexpr (sizeof(T) == 4)
    return fdx_axis_apply_f32(p, x, out, outer, inner, alpha, acc, st);
  else
    return fdx_axis_apply_f64(p, x, out, outer, inner, alpha, acc, st);
}

// D_k on the (n0, n1, n2) grid through the generic axis kernels
template <typename T>
static int apply_axis3(const fdx_plan_t plans[3], int k, const T* x, T* out, double alpha, int acc, void* st) {
  const int64_t n[3] = {plans[0]->n, plans[1]->n, plans[2]->n};
  const int64_t outer = (k == 0) ? 1 : (k == 1 ? n[0] : n[0] * n[1]);
  const int64_t inner = (k == 0) ? n[1] * n[2] : (k == 1 ? n[2] : 1);
  return axis_apply<T>(plans[k], x, out, outer, inner, alpha, acc, st);
}

template <Op OP, typename T>
static int composed(const fdx_plan_t plans[3], const T* const in[3], T* const out[3], const double alpha[3], void* st) {
  int rc = FDX_OK;
  if constexpr (OP == LAP) {
    for (int k = 0; k < 3 && !rc; ++k) rc = apply_axis3<T>(plans, k, in[0], out[0], alpha[k], k > 0, st);
  } else if constexpr (OP == DIV) {
    for (int k = 0; k < 3 && !rc; ++k) rc = apply_axis3<T>(plans, k, in[k], out[0], alpha[k], k > 0, st);
  } else if constexpr (OP == GRAD) {
    for (int k = 0; k < 3 && !rc; ++k) rc = apply_axis3<T>(plans, k, in[0], out[k], alpha[k], 0, st);
  } else {
    const int plus_axis[3] = {1, 2, 0}, plus_comp[3] = {2, 0, 1};
    const int minus_axis[3] = {2, 0, 1}, minus_comp[3] = {1, 2, 0};
    for (int c = 0; c < 3 && !rc; ++c) {
      rc = apply_axis3<T>(plans, plus_axis[c], in[plus_comp[c]], out[c], alpha[plus_axis[c]], 0, st);
      if (!rc) rc = apply_axis3<T>(plans, minus_axis[c], in[minus_comp[c]], out[c], -alpha[minus_axis[c]], 1, st);
    }
  }
  return rc;
}

template <Op OP, typename T>
static int run(const fdx_plan_t plans[3], const T* const in[3], T* const out[3], const double alpha[3],
               const fdx_slab* slab, void* stream) {
  using Tr = OpTraits<OP>;
  if (!plans || !plans[0] || !plans[1] || !plans[2] || !alpha || !in || !out) return FDX_EINVAL;
  for (int f = 0; f < Tr::NIN; ++f)
    if (!in[f]) return FDX_EINVAL;
  for (int f = 0; f < Tr::NOUT; ++f)
    if (!out[f]) return FDX_EINVAL;
  const SlabArgs sl = slab_args(plans, slab);
  if (plans[0]->n == 0 || plans[1]->n == 0 || plans[2]->n == 0) return FDX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int P = fused_radius<OP, T>(plans, in, out, sl);
  if (P < 0) return FDX_EINVAL;
  int rc = FDX_EUNSUPPORTED;
  switch (P) {
    case 1:
      rc = launch_p<OP, T, 1>(plans, in, out, alpha, sl, st);
      break;
    case 2:
      rc = launch_p<OP, T, 2>(plans, in, out, alpha, sl, st);
      break;
    case 3:
      rc = launch_p<OP, T, 3>(plans, in, out, alpha, sl, st);
      break;
    case 4:
      rc = launch_p<OP, T, 4>(plans, in, out, alpha, sl, st);
      break;
    case 5:
      rc = launch_p<OP, T, 5>(plans, in, out, alpha, sl, st);
      break;
    case 6:
      rc = launch_p<OP, T, 6>(plans, in, out, alpha, sl, st);
      break;
    default:
      break;
  }
  if (rc != FDX_EUNSUPPORTED) return rc;
  // fallback: only the whole-grid, no-halo form can be composed from the axis kernels
  if (sl.z_begin != 0 || sl.z_end != plans[0]->n || sl.in_first != 0 || sl.out_first != 0 || sl.z_end2 > sl.z_begin2) return FDX_EUNSUPPORTED;
  return composed<OP, T>(plans, in, out, alpha, stream);
}

}  // namespace fdx

