// Fused 3-D operators (include/fdx_b200.h).  First cut: composition of the axis operator with
// accumulation, inside the library, so that every entry point is live; the single-pass TMA
// kernels replace these paths shape by shape.
#include "fdx_common.cuh"

extern "C" int fdx_axis_apply_f32(fdx_plan_t, const float*, float*, int64_t, int64_t, double, int, void*);
extern "C" int fdx_axis_apply_f64(fdx_plan_t, const double*, double*, int64_t, int64_t, double, int, void*);

namespace fdx {

template <typename T>
static int axis_apply(fdx_plan_t p, const T* x, T* out, int64_t outer, int64_t inner, double alpha, int acc, void* st) {
  if constexpr (sizeof(T) == 4)
    return fdx_axis_apply_f32(p, x, out, outer, inner, alpha, acc, st);
  else
    return fdx_axis_apply_f64(p, x, out, outer, inner, alpha, acc, st);
}

static bool default_slab(const fdx_plan_t plans[3], const fdx_slab* s) {
  return !s || (s->z_begin == 0 && s->z_end == plans[0]->n && s->halo_lo == 0 && s->halo_hi == 0);
}

static bool plans_ok(const fdx_plan_t plans[3]) { return plans && plans[0] && plans[1] && plans[2]; }

// D_k applied on the (n0, n1, n2) grid: axis k seen as (outer, n_k, inner)
template <typename T>
static int apply_axis3(const fdx_plan_t plans[3], int k, const T* x, T* out, double alpha, int acc, void* st) {
  const int64_t n[3] = {plans[0]->n, plans[1]->n, plans[2]->n};
  const int64_t outer = (k == 0) ? 1 : (k == 1 ? n[0] : n[0] * n[1]);
  const int64_t inner = (k == 0) ? n[1] * n[2] : (k == 1 ? n[2] : 1);
  return axis_apply<T>(plans[k], x, out, outer, inner, alpha, acc, st);
}

template <typename T>
static int laplacian_composed(const fdx_plan_t plans[3], const T* x, T* out, const double alpha[3], void* st) {
  for (int k = 0; k < 3; ++k) {
    int rc = apply_axis3<T>(plans, k, x, out, alpha[k], k > 0, st);
    if (rc) return rc;
  }
  return FDX_OK;
}

template <typename T>
static int divergence_composed(const fdx_plan_t plans[3], const T* const x[3], T* out, const double alpha[3], void* st) {
  for (int k = 0; k < 3; ++k) {
    int rc = apply_axis3<T>(plans, k, x[k], out, alpha[k], k > 0, st);
    if (rc) return rc;
  }
  return FDX_OK;
}

template <typename T>
static int gradient_composed(const fdx_plan_t plans[3], const T* x, T* const out[3], const double alpha[3], void* st) {
  for (int k = 0; k < 3; ++k) {
    int rc = apply_axis3<T>(plans, k, x, out[k], alpha[k], 0, st);
    if (rc) return rc;
  }
  return FDX_OK;
}

template <typename T>
static int curl_composed(const fdx_plan_t plans[3], const T* const x[3], T* const out[3], const double alpha[3], void* st) {
  // out_0 = a1 D1 x2 - a2 D2 x1 ; out_1 = a2 D2 x0 - a0 D0 x2 ; out_2 = a0 D0 x1 - a1 D1 x0
  const int plus_axis[3] = {1, 2, 0}, plus_comp[3] = {2, 0, 1};
  const int minus_axis[3] = {2, 0, 1}, minus_comp[3] = {1, 2, 0};
  for (int c = 0; c < 3; ++c) {
    int rc = apply_axis3<T>(plans, plus_axis[c], x[plus_comp[c]], out[c], alpha[plus_axis[c]], 0, st);
    if (rc) return rc;
    rc = apply_axis3<T>(plans, minus_axis[c], x[minus_comp[c]], out[c], -alpha[minus_axis[c]], 1, st);
    if (rc) return rc;
  }
  return FDX_OK;
}

}  // namespace fdx

#define FDX_DEFINE_FUSED(SUFFIX, T)                                                                                     \
  extern "C" int fdx_laplacian3d_##SUFFIX(const fdx_plan_t plans[3], const T* x, T* out, const double alpha[3],          \
                                          const fdx_slab* slab, void* stream) {                                         \
    if (!fdx::plans_ok(plans) || !x || !out || !alpha) return FDX_EINVAL;                                               \
    if (!fdx::default_slab(plans, slab)) return FDX_EUNSUPPORTED;                                                       \
    return fdx::laplacian_composed<T>(plans, x, out, alpha, stream);                                                    \
  }                                                                                                                     \
  extern "C" int fdx_divergence3d_##SUFFIX(const fdx_plan_t plans[3], const T* const x[3], T* out, const double alpha[3], \
                                           const fdx_slab* slab, void* stream) {                                        \
    if (!fdx::plans_ok(plans) || !x || !x[0] || !x[1] || !x[2] || !out || !alpha) return FDX_EINVAL;                    \
    if (!fdx::default_slab(plans, slab)) return FDX_EUNSUPPORTED;                                                       \
    return fdx::divergence_composed<T>(plans, x, out, alpha, stream);                                                   \
  }                                                                                                                     \
  extern "C" int fdx_gradient3d_##SUFFIX(const fdx_plan_t plans[3], const T* x, T* const out[3], const double alpha[3],  \
                                         const fdx_slab* slab, void* stream) {                                          \
    if (!fdx::plans_ok(plans) || !x || !out || !out[0] || !out[1] || !out[2] || !alpha) return FDX_EINVAL;              \
    if (!fdx::default_slab(plans, slab)) return FDX_EUNSUPPORTED;                                                       \
    return fdx::gradient_composed<T>(plans, x, out, alpha, stream);                                                     \
  }                                                                                                                     \
  extern "C" int fdx_curl3d_##SUFFIX(const fdx_plan_t plans[3], const T* const x[3], T* const out[3],                    \
                                     const double alpha[3], const fdx_slab* slab, void* stream) {                       \
    if (!fdx::plans_ok(plans) || !x || !x[0] || !x[1] || !x[2] || !out || !out[0] || !out[1] || !out[2] || !alpha)      \
      return FDX_EINVAL;                                                                                                \
    if (!fdx::default_slab(plans, slab)) return FDX_EUNSUPPORTED;                                                       \
    return fdx::curl_composed<T>(plans, x, out, alpha, stream);                                                         \
  }

FDX_DEFINE_FUSED(f32, float)
FDX_DEFINE_FUSED(f64, double)
