// One translation unit per fused operator so that nvcc compiles them in parallel.
#include "fdx_fused3d.cuh"

#define FDX_DEF(SUFFIX, T)                                                                           \
  extern "C" int fdx_curl3d_##SUFFIX(const fdx_plan_t plans[3], const T* const x[3], T* const out[3], \
                                     const double alpha[3], const fdx_slab* slab, void* stream) {    \
    if (!x || !out) return FDX_EINVAL;                                                               \
    return fdx::run<fdx::CURL, T>(plans, x, out, alpha, slab, stream);                               \
  }

FDX_DEF(f32, float)
FDX_DEF(f64, double)
