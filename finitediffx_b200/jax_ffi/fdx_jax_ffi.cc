// XLA FFI (jax.ffi) handlers over the fdx_b200 C ABI.   UNTESTED IN THIS REPOSITORY'S IMAGE:
// JAX and the XLA FFI headers are not installed there, so this file has never been
// compiled.  It is built only when `jax.ffi.include_dir()` resolves (finitediffx_b200/jax_ffi/__init__.py);
// the CUDA kernels and the C entry points it calls are the ones tests/ exercise through ctypes
// (tests/test_gpu_parity.py::test_step_size_as_a_device_value... covers the *_dev_alpha entries,
// under CUDA-graph capture).
//
//   g++ -O2 -fPIC -shared -std=c++17 -I$(python -c "import jax; print(jax.ffi.include_dir())") \
//       -I/usr/local/cuda/include -I../../include fdx_jax_ffi.cc -L../_lib -lfdx_b200 -lcudart \
//       -o ../_lib/libfdx_jax_ffi.so
//
// Each handler replaces one reference operator on the `jax.jit` path:
//   fdx_axis_apply   <- finitediffx/_src/finite_diff.py:168-297  (difference)
//   fdx_laplacian3d  <- :532-575   fdx_divergence3d <- :404-457
//   fdx_gradient3d   <- :300-351   fdx_curl3d       <- :609-668
//
// Contract of every handler
//   * operands: the array(s) and `alpha`, a float64 DEVICE vector of scales 1 / step_size**derivative (one element for
//     the axis operator, three for the 3-D operators).  `step_size` is traced in the reference (finite_diff.py:168), so
//     alpha stays on the device: the handlers pass its pointer to the *_dev_alpha entry points and never read it
//     (no cudaStreamSynchronize, safe under XLA's command-buffer capture);
//   * attributes: the plans, fully described by the Python side (finitediffx_b200.plan.AxisPlan: main taps and the
//     two band tables), so this shim holds no coefficient mathematics;
//   * plans are device objects (cudaMalloc + cudaMemcpy at creation): they are created in the INSTANTIATE stage of
//     the handler and looked up in the execute stage.  The cache key holds the device ordinal -- a plan is bound to
//     the device that was current when it was created.
#include <cuda_runtime_api.h>

#include <array>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "fdx_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

struct PlanKey {
  int device;
  int32_t n, lo, hi, nl, wl, nr, wr;
  std::vector<double> main, band_l, band_r;
  bool operator<(const PlanKey& o) const {
    return std::tie(device, n, lo, hi, nl, wl, nr, wr, main, band_l, band_r) <
           std::tie(o.device, o.n, o.lo, o.hi, o.nl, o.wl, o.nr, o.wr, o.main, o.band_l, o.band_r);
  }
};

std::mutex g_mu;
std::map<PlanKey, fdx_plan_t> g_plans;

// create == true: instantiate stage (allocation allowed).  create == false: execute stage, lookup only.
fdx_plan_t GetPlan(const PlanKey& k, bool create) {
  std::lock_guard<std::mutex> lock(g_mu);
  auto it = g_plans.find(k);
  if (it != g_plans.end()) return it->second;
  if (!create) return nullptr;
  fdx_axis_desc d{};
  d.n = k.n, d.lo = k.lo, d.hi = k.hi;
  for (size_t i = 0; i < k.main.size() && i < FDX_MAX_TAPS; ++i) d.main_taps[i] = k.main[i];
  d.nl = k.nl, d.wl = k.wl, d.nr = k.nr, d.wr = k.wr;
  d.band_l = k.band_l.empty() ? nullptr : k.band_l.data();
  d.band_r = k.band_r.empty() ? nullptr : k.band_r.data();
  fdx_plan_t p = nullptr;
  if (fdx_plan_create(&p, &d) != FDX_OK) return nullptr;
  g_plans.emplace(k, p);
  return p;
}

// Plan attributes travel flattened: `plan_shape` holds 7 int32 per plan (n, lo, hi, nl, wl, nr, wr); `main_taps`,
// `band_l`, `band_r` hold the plans' arrays back to back (lengths follow from plan_shape).
bool KeysFrom(int n_plans, ffi::Span<const int32_t> shape, ffi::Span<const double> main, ffi::Span<const double> bl,
              ffi::Span<const double> br, std::vector<PlanKey>* keys) {
  int device = 0;
  if (cudaGetDevice(&device) != cudaSuccess || (int)shape.size() != 7 * n_plans) return false;
  size_t om = 0, ol = 0, orr = 0;
  for (int i = 0; i < n_plans; ++i) {
    const int32_t* s = shape.begin() + 7 * i;
    PlanKey k{device, s[0], s[1], s[2], s[3], s[4], s[5], s[6], {}, {}, {}};
    const size_t nm = (size_t)(s[2] - s[1] + 1), nl = (size_t)s[3] * s[4], nr = (size_t)s[5] * s[6];
    if (om + nm > main.size() || ol + nl > bl.size() || orr + nr > br.size()) return false;
    k.main.assign(main.begin() + om, main.begin() + om + nm);
    k.band_l.assign(bl.begin() + ol, bl.begin() + ol + nl);
    k.band_r.assign(br.begin() + orr, br.begin() + orr + nr);
    om += nm, ol += nl, orr += nr;
    keys->push_back(std::move(k));
  }
  return true;
}

ffi::Error Status(int rc, const char* what) {
  if (rc == FDX_OK) return ffi::Error::Success();
  return ffi::Error(rc == FDX_EUNSUPPORTED ? ffi::ErrorCode::kUnimplemented : ffi::ErrorCode::kInternal,
                    std::string(what) + ": " + fdx_error_string(rc));
}

// ---- instantiate stage, shared by every target: create the plans the execute stage will look up
ffi::Error InstantiatePlans(ffi::Span<const int32_t> plan_shape, ffi::Span<const double> main, ffi::Span<const double> bl,
                            ffi::Span<const double> br) {
  std::vector<PlanKey> keys;
  if (!KeysFrom((int)plan_shape.size() / 7, plan_shape, main, bl, br, &keys)) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "fdx: malformed plan attributes");
  for (const auto& k : keys)
    if (!GetPlan(k, /*create=*/true)) return ffi::Error(ffi::ErrorCode::kInternal, "fdx_plan_create failed");
  return ffi::Error::Success();
}

bool LookupPlans(int n_plans, ffi::Span<const int32_t> shape, ffi::Span<const double> main, ffi::Span<const double> bl,
                 ffi::Span<const double> br, fdx_plan_t* out) {
  std::vector<PlanKey> keys;
  if (!KeysFrom(n_plans, shape, main, bl, br, &keys)) return false;
  for (int i = 0; i < n_plans; ++i) {
    // (a runtime that skips the instantiate stage still works: the plan is created here, outside any capture)
    out[i] = GetPlan(keys[i], /*create=*/true);
    if (!out[i]) return false;
  }
  return true;
}

// ---- difference along one axis: out = alpha[0] * D x ---------------------------------------------
template <typename T, ffi::DataType DT>
ffi::Error AxisApplyImpl(cudaStream_t stream, ffi::Buffer<DT> x, ffi::Buffer<ffi::F64> alpha, ffi::ResultBuffer<DT> out,
                         int64_t axis, ffi::Span<const int32_t> plan_shape, ffi::Span<const double> main,
                         ffi::Span<const double> band_l, ffi::Span<const double> band_r) {
  fdx_plan_t plan = nullptr;
  if (!LookupPlans(1, plan_shape, main, band_l, band_r, &plan)) return ffi::Error(ffi::ErrorCode::kInternal, "fdx: no plan");
  auto dims = x.dimensions();
  int64_t outer = 1, inner = 1;
  for (int64_t i = 0; i < axis; ++i) outer *= dims[i];
  for (size_t i = axis + 1; i < dims.size(); ++i) inner *= dims[i];
  int rc;
  if constexpr (sizeof(T) == 4)
    rc = fdx_axis_apply_dev_alpha_f32(plan, x.typed_data(), out->typed_data(), outer, inner, alpha.typed_data(), 1.0, 0, stream);
  else
    rc = fdx_axis_apply_dev_alpha_f64(plan, x.typed_data(), out->typed_data(), outer, inner, alpha.typed_data(), 1.0, 0, stream);
  return Status(rc, "fdx_axis_apply_dev_alpha");
}

// ---- the 3-D operators.  Vector fields are one (3, n0, n1, n2) buffer: component k starts at k * n0 * n1 * n2.
enum Kind { kLaplacian, kDivergence, kGradient, kCurl };

template <typename T, ffi::DataType DT, Kind K>
ffi::Error Fused3dImpl(cudaStream_t stream, ffi::Buffer<DT> x, ffi::Buffer<ffi::F64> alpha, ffi::ResultBuffer<DT> out,
                       ffi::Span<const int32_t> plan_shape, ffi::Span<const double> main, ffi::Span<const double> band_l,
                       ffi::Span<const double> band_r) {
  fdx_plan_t plans[3];
  if (!LookupPlans(3, plan_shape, main, band_l, band_r, plans)) return ffi::Error(ffi::ErrorCode::kInternal, "fdx: no plans");
  const int64_t field = (int64_t)plan_shape[0] * plan_shape[7] * plan_shape[14];
  const T* xp = x.typed_data();
  T* op = out->typed_data();
  const T* xs[3] = {xp, xp + field, xp + 2 * field};
  T* os[3] = {op, op + field, op + 2 * field};
  const double* a = alpha.typed_data();
  int rc;
  if constexpr (sizeof(T) == 4) {
    if constexpr (K == kLaplacian) rc = fdx_laplacian3d_dev_alpha_f32(plans, xp, op, a, stream);
    if constexpr (K == kDivergence) rc = fdx_divergence3d_dev_alpha_f32(plans, xs, op, a, stream);
    if constexpr (K == kGradient) rc = fdx_gradient3d_dev_alpha_f32(plans, xp, os, a, stream);
    if constexpr (K == kCurl) rc = fdx_curl3d_dev_alpha_f32(plans, xs, os, a, stream);
  } else {
    if constexpr (K == kLaplacian) rc = fdx_laplacian3d_dev_alpha_f64(plans, xp, op, a, stream);
    if constexpr (K == kDivergence) rc = fdx_divergence3d_dev_alpha_f64(plans, xs, op, a, stream);
    if constexpr (K == kGradient) rc = fdx_gradient3d_dev_alpha_f64(plans, xp, os, a, stream);
    if constexpr (K == kCurl) rc = fdx_curl3d_dev_alpha_f64(plans, xs, os, a, stream);
  }
  return Status(rc, "fdx_*3d_dev_alpha");
}

// A static step size (the common case: a Python float) is known when the call is built; `alpha_host` then carries the
// three scales as an attribute and the single-pass fused kernels run (alpha folded into their taps on the host).
template <typename T, ffi::DataType DT, Kind K>
ffi::Error Fused3dStaticImpl(cudaStream_t stream, ffi::Buffer<DT> x, ffi::ResultBuffer<DT> out, ffi::Span<const double> alpha_host,
                             ffi::Span<const int32_t> plan_shape, ffi::Span<const double> main, ffi::Span<const double> band_l,
                             ffi::Span<const double> band_r) {
  fdx_plan_t plans[3];
  if (alpha_host.size() != 3 || !LookupPlans(3, plan_shape, main, band_l, band_r, plans)) return ffi::Error(ffi::ErrorCode::kInternal, "fdx: no plans");
  const int64_t field = (int64_t)plan_shape[0] * plan_shape[7] * plan_shape[14];
  const T* xp = x.typed_data();
  T* op = out->typed_data();
  const T* xs[3] = {xp, xp + field, xp + 2 * field};
  T* os[3] = {op, op + field, op + 2 * field};
  const double a[3] = {alpha_host[0], alpha_host[1], alpha_host[2]};
  int rc;
  if constexpr (sizeof(T) == 4) {
    if constexpr (K == kLaplacian) rc = fdx_laplacian3d_f32(plans, xp, op, a, nullptr, stream);
    if constexpr (K == kDivergence) rc = fdx_divergence3d_f32(plans, xs, op, a, nullptr, stream);
    if constexpr (K == kGradient) rc = fdx_gradient3d_f32(plans, xp, os, a, nullptr, stream);
    if constexpr (K == kCurl) rc = fdx_curl3d_f32(plans, xs, os, a, nullptr, stream);
  } else {
    if constexpr (K == kLaplacian) rc = fdx_laplacian3d_f64(plans, xp, op, a, nullptr, stream);
    if constexpr (K == kDivergence) rc = fdx_divergence3d_f64(plans, xs, op, a, nullptr, stream);
    if constexpr (K == kGradient) rc = fdx_gradient3d_f64(plans, xp, os, a, nullptr, stream);
    if constexpr (K == kCurl) rc = fdx_curl3d_f64(plans, xs, os, a, nullptr, stream);
  }
  return Status(rc, "fdx_*3d");
}

}  // namespace

#define FDX_PLAN_ATTRS()                            \
  .Attr<ffi::Span<const int32_t>>("plan_shape")     \
      .Attr<ffi::Span<const double>>("main_taps")   \
      .Attr<ffi::Span<const double>>("band_l")      \
      .Attr<ffi::Span<const double>>("band_r")

// instantiate stage (one symbol serves every target): plans are created here, outside the execute stage
XLA_FFI_DEFINE_HANDLER_SYMBOL(fdx_instantiate_ffi, InstantiatePlans,
                              ffi::Ffi::Bind<ffi::ExecutionStage::kInstantiate>() FDX_PLAN_ATTRS());

#define FDX_AXIS(NAME, T, DT)                                                                                             \
  XLA_FFI_DEFINE_HANDLER_SYMBOL(NAME, (AxisApplyImpl<T, DT>),                                                             \
                                ffi::Ffi::Bind()                                                                          \
                                    .Ctx<ffi::PlatformStream<cudaStream_t>>()                                             \
                                    .Arg<ffi::Buffer<DT>>()                                                               \
                                    .Arg<ffi::Buffer<ffi::F64>>()                                                         \
                                    .Ret<ffi::Buffer<DT>>()                                                               \
                                    .Attr<int64_t>("axis") FDX_PLAN_ATTRS())
FDX_AXIS(fdx_axis_apply_f32_ffi, float, ffi::F32);
FDX_AXIS(fdx_axis_apply_f64_ffi, double, ffi::F64);

#define FDX_FUSED(NAME, T, DT, K)                                                                                         \
  XLA_FFI_DEFINE_HANDLER_SYMBOL(NAME##_ffi, (Fused3dImpl<T, DT, K>),                                                      \
                                ffi::Ffi::Bind()                                                                          \
                                    .Ctx<ffi::PlatformStream<cudaStream_t>>()                                             \
                                    .Arg<ffi::Buffer<DT>>()                                                               \
                                    .Arg<ffi::Buffer<ffi::F64>>()                                                         \
                                    .Ret<ffi::Buffer<DT>>() FDX_PLAN_ATTRS());                                            \
  XLA_FFI_DEFINE_HANDLER_SYMBOL(NAME##_static_ffi, (Fused3dStaticImpl<T, DT, K>),                                         \
                                ffi::Ffi::Bind()                                                                          \
                                    .Ctx<ffi::PlatformStream<cudaStream_t>>()                                             \
                                    .Arg<ffi::Buffer<DT>>()                                                               \
                                    .Ret<ffi::Buffer<DT>>()                                                               \
                                    .Attr<ffi::Span<const double>>("alpha") FDX_PLAN_ATTRS())
FDX_FUSED(fdx_laplacian3d_f32, float, ffi::F32, kLaplacian);
FDX_FUSED(fdx_laplacian3d_f64, double, ffi::F64, kLaplacian);
FDX_FUSED(fdx_divergence3d_f32, float, ffi::F32, kDivergence);
FDX_FUSED(fdx_divergence3d_f64, double, ffi::F64, kDivergence);
FDX_FUSED(fdx_gradient3d_f32, float, ffi::F32, kGradient);
FDX_FUSED(fdx_gradient3d_f64, double, ffi::F64, kGradient);
FDX_FUSED(fdx_curl3d_f32, float, ffi::F32, kCurl);
FDX_FUSED(fdx_curl3d_f64, double, ffi::F64, kCurl);
