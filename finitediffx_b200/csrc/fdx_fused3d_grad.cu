// One translation unit per fused operator so that nvcc compiles them in parallel.
#include "fdx_fused3d.cuh"

#define FDX_DEF(SUFFIX, T)                                                                                             \
  extern "C" int fdx_gradient3d_##SUFFIX(const fdx_plan_t plans[3], const T* x, T* const out[3], const double alpha[3], \
                                         const fdx_slab* slab, void* stream) {                                         \
    if (!out) return FDX_EINVAL;                                                                                       \
    const T* in[3] = {x, nullptr, nullptr};                                                                            \
    return fdx::run<fdx::GRAD, T>(plans, in, out, alpha, slab, stream);                                                \
  }

FDX_DEF(f32, float)
FDX_DEF(f64, double)
