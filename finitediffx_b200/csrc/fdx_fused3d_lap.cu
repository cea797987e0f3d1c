// One translation unit per fused operator so that nvcc compiles them in parallel.
#include "fdx_fused3d.cuh"

#define FDX_DEF(SUFFIX, T)                                                                                      \
  extern "C" int fdx_laplacian3d_##SUFFIX(const fdx_plan_t plans[3], const T* x, T* out, const double alpha[3],  \
                                          const fdx_slab* slab, void* stream) {                                 \
    const T* in[3] = {x, nullptr, nullptr};                                                                     \
    T* o[3] = {out, nullptr, nullptr};                                                                          \
    return fdx::run<fdx::LAP, T>(plans, in, o, alpha, slab, stream);                                            \
  }

FDX_DEF(f32, float)
FDX_DEF(f64, double)
