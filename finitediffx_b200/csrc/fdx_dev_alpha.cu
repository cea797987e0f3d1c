// Fused-operator entry points whose scale is read ON THE DEVICE (include/fdx_b200.h, *_dev_alpha).
//
// The reference traces `step_size` (finitediffx/_src/finite_diff.py:168: not in static_argnames), so inside a jitted
// caller the scale 1 / step_size**derivative is a device value.  An XLA custom call must not fetch it (a host read
// stalls the stream and cannot be captured), so these entry points take `alpha_dev`, a device pointer to three
// float64 scales, one per axis.  They only enqueue kernels: no allocation, no synchronisation, capture-safe.
//
// The single-pass fused kernels fold the scale into their constant-bank taps on the host, which a device scale cannot
// do; these entry points therefore run the operator as a composition of the axis kernels (three passes for laplacian /
// divergence / gradient, six for curl), each of which multiplies its result by the device scale.  A caller with
// concrete step sizes -- the common case, and the one every BASELINE config is in -- uses the fused entry points.
#include "fdx_common.cuh"

extern "C" int fdx_axis_apply_dev_alpha_f32(fdx_plan_t, const float*, float*, int64_t, int64_t, const double*, double, int, void*);
extern "C" int fdx_axis_apply_dev_alpha_f64(fdx_plan_t, const double*, double*, int64_t, int64_t, const double*, double, int, void*);

namespace fdx {
namespace {

template <typename T>
int axis3(const fdx_plan_t plans[3], int k, const T* x, T* out, const double* alpha_dev, double sign, int acc, void* st) {
  const int64_t n[3] = {plans[0]->n, plans[1]->n, plans[2]->n};
  const int64_t outer = (k == 0) ? 1 : (k == 1 ? n[0] : n[0] * n[1]);
  const int64_t inner = (k == 0) ? n[1] * n[2] : (k == 1 ? n[2] : 1);
  if constexpr (sizeof(T) == 4)
    return fdx_axis_apply_dev_alpha_f32(plans[k], x, out, outer, inner, alpha_dev + k, sign, acc, st);
  else
    return fdx_axis_apply_dev_alpha_f64(plans[k], x, out, outer, inner, alpha_dev + k, sign, acc, st);
}

bool ok(const fdx_plan_t plans[3], const double* alpha_dev) { return plans && plans[0] && plans[1] && plans[2] && alpha_dev; }

template <typename T>
int laplacian(const fdx_plan_t plans[3], const T* x, T* out, const double* a, void* st) {
  if (!ok(plans, a) || !x || !out) return FDX_EINVAL;
  int rc = FDX_OK;
  for (int k = 0; k < 3 && !rc; ++k) rc = axis3<T>(plans, k, x, out, a, 1.0, k > 0, st);
  return rc;
}
template <typename T>
int divergence(const fdx_plan_t plans[3], const T* const x[3], T* out, const double* a, void* st) {
  if (!ok(plans, a) || !x || !x[0] || !x[1] || !x[2] || !out) return FDX_EINVAL;
  int rc = FDX_OK;
  for (int k = 0; k < 3 && !rc; ++k) rc = axis3<T>(plans, k, x[k], out, a, 1.0, k > 0, st);
  return rc;
}
template <typename T>
int gradient(const fdx_plan_t plans[3], const T* x, T* const out[3], const double* a, void* st) {
  if (!ok(plans, a) || !x || !out || !out[0] || !out[1] || !out[2]) return FDX_EINVAL;
  int rc = FDX_OK;
  for (int k = 0; k < 3 && !rc; ++k) rc = axis3<T>(plans, k, x, out[k], a, 1.0, 0, st);
  return rc;
}
template <typename T>
int curl(const fdx_plan_t plans[3], const T* const x[3], T* const out[3], const double* a, void* st) {
  if (!ok(plans, a) || !x || !x[0] || !x[1] || !x[2] || !out || !out[0] || !out[1] || !out[2]) return FDX_EINVAL;
  // out_c = D_{c+1} x_{c+2} - D_{c+2} x_{c+1}   (indices mod 3; finite_diff.py:616-668)
  const int plus_axis[3] = {1, 2, 0}, plus_comp[3] = {2, 0, 1}, minus_axis[3] = {2, 0, 1}, minus_comp[3] = {1, 2, 0};
  int rc = FDX_OK;
  for (int c = 0; c < 3 && !rc; ++c) {
    rc = axis3<T>(plans, plus_axis[c], x[plus_comp[c]], out[c], a, 1.0, 0, st);
    if (!rc) rc = axis3<T>(plans, minus_axis[c], x[minus_comp[c]], out[c], a, -1.0, 1, st);
  }
  return rc;
}

}  // namespace
}  // namespace fdx

#define FDX_DEV_ALPHA(SUFFIX, T)                                                                                              \
  extern "C" int fdx_laplacian3d_dev_alpha_##SUFFIX(const fdx_plan_t plans[3], const T* x, T* out, const double* alpha_dev,   \
                                                    void* stream) {                                                           \
    return fdx::laplacian<T>(plans, x, out, alpha_dev, stream);                                                               \
  }                                                                                                                           \
  extern "C" int fdx_divergence3d_dev_alpha_##SUFFIX(const fdx_plan_t plans[3], const T* const x[3], T* out,                  \
                                                     const double* alpha_dev, void* stream) {                                 \
    return fdx::divergence<T>(plans, x, out, alpha_dev, stream);                                                              \
  }                                                                                                                           \
  extern "C" int fdx_gradient3d_dev_alpha_##SUFFIX(const fdx_plan_t plans[3], const T* x, T* const out[3],                    \
                                                   const double* alpha_dev, void* stream) {                                   \
    return fdx::gradient<T>(plans, x, out, alpha_dev, stream);                                                                \
  }                                                                                                                           \
  extern "C" int fdx_curl3d_dev_alpha_##SUFFIX(const fdx_plan_t plans[3], const T* const x[3], T* const out[3],               \
                                                const double* alpha_dev, void* stream) {                                      \
    return fdx::curl<T>(plans, x, out, alpha_dev, stream);                                                                    \
  }

FDX_DEV_ALPHA(f32, float)
FDX_DEV_ALPHA(f64, double)
