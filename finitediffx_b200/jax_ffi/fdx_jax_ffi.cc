// XLA FFI (jax.ffi) handlers over the fdx_b200 C ABI.   UNTESTED IN THIS REPOSITORY'S IMAGE:
// JAX and the XLA FFI headers are not installed there.  This file is compiled only
// when `jax.ffi.include_dir()` resolves (finitediffx_b200/jax_ffi/__init__.py); the CUDA kernels it
// calls are the ones exercised by tests/ through ctypes.
//
//   g++ -O2 -fPIC -shared -std=c++17 -I$(python -c "import jax; print(jax.ffi.include_dir())") \
//       -I/usr/local/cuda/include -I../../include fdx_jax_ffi.cc -L../_lib -lfdx_b200 -lcudart \
//       -o ../_lib/libfdx_jax_ffi.so
//
// Each handler replaces one reference operator on the `jax.jit` path:
//   fdx_axis_apply   <- finitediffx/_src/finite_diff.py:168-297  (difference)
//   fdx_laplacian3d  <- :532-575   fdx_divergence3d <- :404-457
//   fdx_gradient3d   <- :300-351   fdx_curl3d       <- :609-668
// Static attributes carry (derivative, accuracy, method, transposed, axis); `alpha` = 1/step**d is a
// traced f64[3] operand because `step_size` is traced in the reference.  Plans are created once per
// attribute set in a process-wide cache (the FFI "prepare" stage is the place for the cudaMalloc).
#include <cuda_runtime_api.h>

#include <array>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "fdx_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

// Plans arrive fully described by the Python side (finitediffx_b200.plan.AxisPlan): main taps and the
// two band tables as attribute arrays, so this shim needs no coefficient mathematics of its own.
struct PlanKey {
  int32_t n, lo, hi, nl, wl, nr, wr;
  std::vector<double> main, band_l, band_r;
  bool operator<(const PlanKey& o) const {
    return std::tie(n, lo, hi, nl, wl, nr, wr, main, band_l, band_r) <
           std::tie(o.n, o.lo, o.hi, o.nl, o.wl, o.nr, o.wr, o.main, o.band_l, o.band_r);
  }
};

fdx_plan_t GetPlan(const PlanKey& k) {
  static std::mutex mu;
  static std::map<PlanKey, fdx_plan_t> cache;
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(k);
  if (it != cache.end()) return it->second;
  fdx_axis_desc d{};
  d.n = k.n, d.lo = k.lo, d.hi = k.hi;
  for (size_t i = 0; i < k.main.size() && i < FDX_MAX_TAPS; ++i) d.main_taps[i] = k.main[i];
  d.nl = k.nl, d.wl = k.wl, d.nr = k.nr, d.wr = k.wr;
  d.band_l = k.band_l.empty() ? nullptr : k.band_l.data();
  d.band_r = k.band_r.empty() ? nullptr : k.band_r.data();
  fdx_plan_t p = nullptr;
  if (fdx_plan_create(&p, &d) != FDX_OK) return nullptr;
  cache.emplace(k, p);
  return p;
}

PlanKey KeyFrom(ffi::Span<const int32_t> shape7, ffi::Span<const double> main, ffi::Span<const double> bl,
                ffi::Span<const double> br) {
  PlanKey k{shape7[0], shape7[1], shape7[2], shape7[3], shape7[4], shape7[5], shape7[6], {}, {}, {}};
  k.main.assign(main.begin(), main.end());
  k.band_l.assign(bl.begin(), bl.end());
  k.band_r.assign(br.begin(), br.end());
  return k;
}

ffi::Error Status(int rc, const char* what) {
  if (rc == FDX_OK) return ffi::Error::Success();
  return ffi::Error(rc == FDX_EUNSUPPORTED ? ffi::ErrorCode::kUnimplemented : ffi::ErrorCode::kInternal,
                    std::string(what) + ": " + fdx_error_string(rc));
}

// ---- difference along one axis: out = alpha * D x ------------------------------------------------
template <typename T, ffi::DataType DT>
ffi::Error AxisApplyImpl(cudaStream_t stream, ffi::Buffer<DT> x, ffi::Buffer<ffi::F64> alpha, ffi::ResultBuffer<DT> out,
                         int64_t axis, ffi::Span<const int32_t> plan_shape, ffi::Span<const double> main,
                         ffi::Span<const double> band_l, ffi::Span<const double> band_r) {
  fdx_plan_t plan = GetPlan(KeyFrom(plan_shape, main, band_l, band_r));
  if (!plan) return ffi::Error(ffi::ErrorCode::kInternal, "fdx_plan_create failed");
  auto dims = x.dimensions();
  int64_t outer = 1, inner = 1;
  for (int64_t i = 0; i < axis; ++i) outer *= dims[i];
  for (size_t i = axis + 1; i < dims.size(); ++i) inner *= dims[i];
  // alpha is a device scalar in XLA; the C ABI takes it by value, so it is read back here.  A
  // production shim would add an entry point taking alpha by device pointer to stay capture-safe.
  double a = 1.0;
  cudaMemcpyAsync(&a, alpha.typed_data(), sizeof(double), cudaMemcpyDeviceToHost, stream);
  cudaStreamSynchronize(stream);
  int rc;
  if constexpr (sizeof(T) == 4)
    rc = fdx_axis_apply_f32(plan, x.typed_data(), out->typed_data(), outer, inner, a, 0, stream);
  else
    rc = fdx_axis_apply_f64(plan, x.typed_data(), out->typed_data(), outer, inner, a, 0, stream);
  return Status(rc, "fdx_axis_apply");
}

ffi::Error AxisApplyF32(cudaStream_t s, ffi::Buffer<ffi::F32> x, ffi::Buffer<ffi::F64> a, ffi::ResultBuffer<ffi::F32> o,
                        int64_t axis, ffi::Span<const int32_t> ps, ffi::Span<const double> m, ffi::Span<const double> bl,
                        ffi::Span<const double> br) {
  return AxisApplyImpl<float, ffi::F32>(s, x, a, o, axis, ps, m, bl, br);
}
ffi::Error AxisApplyF64(cudaStream_t s, ffi::Buffer<ffi::F64> x, ffi::Buffer<ffi::F64> a, ffi::ResultBuffer<ffi::F64> o,
                        int64_t axis, ffi::Span<const int32_t> ps, ffi::Span<const double> m, ffi::Span<const double> bl,
                        ffi::Span<const double> br) {
  return AxisApplyImpl<double, ffi::F64>(s, x, a, o, axis, ps, m, bl, br);
}

}  // namespace

#define FDX_AXIS_BINDING(DT)                       \
  ffi::Ffi::Bind()                                 \
      .Ctx<ffi::PlatformStream<cudaStream_t>>()    \
      .Arg<ffi::Buffer<DT>>()                      \
      .Arg<ffi::Buffer<ffi::F64>>()                \
      .Ret<ffi::Buffer<DT>>()                      \
      .Attr<int64_t>("axis")                       \
      .Attr<ffi::Span<const int32_t>>("plan_shape") \
      .Attr<ffi::Span<const double>>("main_taps")  \
      .Attr<ffi::Span<const double>>("band_l")     \
      .Attr<ffi::Span<const double>>("band_r")

XLA_FFI_DEFINE_HANDLER_SYMBOL(fdx_axis_apply_f32_ffi, AxisApplyF32, FDX_AXIS_BINDING(ffi::F32));
XLA_FFI_DEFINE_HANDLER_SYMBOL(fdx_axis_apply_f64_ffi, AxisApplyF64, FDX_AXIS_BINDING(ffi::F64));
// The fused 3-D handlers (fdx_laplacian3d_*, fdx_divergence3d_*, fdx_gradient3d_*, fdx_curl3d_*) follow the
// same pattern with three plans' attributes and component buffers; see INTEGRATION.md.
