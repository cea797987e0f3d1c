/*
 * fdx_b200 -- C ABI of the B200-native finite-difference stencil library.
 *
 * The reference (ASEM000/finitediffX) has no native seam: its array path is pure Python on
 * jax.numpy.  Every public operator funnels through `difference`
 * (finitediffx/_src/finite_diff.py:168-297), so that is where this ABI cuts: one entry point
 * for the 1-D axis operator, and one per fused vector-calculus operator.  A binding
 * (ctypes here, jax.ffi / XLA custom call in INTEGRATION.md) only has to marshal raw device
 * pointers, sizes and a stream.
 *
 * Conventions
 *   - all array pointers are DEVICE pointers, C-contiguous, caller-owned, outputs pre-allocated,
 *     inputs and outputs must not alias;
 *   - `stream` is a cudaStream_t passed as void*; compute entry points only enqueue work:
 *     no allocation, no synchronisation, re-entrant from any host thread;
 *   - every function returns 0 on success, a negative FDX_E* code on misuse, or a positive
 *     cudaError_t value; no exception crosses the ABI;
 *   - `alpha` is the runtime scale (1 / step_size**derivative in the reference,
 *     finite_diff.py:257,270,280,290); it is NOT part of a plan because `step_size` is a traced
 *     value in the reference.
 *
 * A plan (fdx_plan_t) is the device-resident form of one 1-D operator: uniform "main" taps plus
 * dense boundary bands (the reference's one-sided boundary stencils, finite_diff.py:75-88,
 * 105-113, 130-138, 155-163, or their transposes for the adjoint).  It replaces the reference's
 * only cache, the jax.jit cache around the coefficient solve (finitediffx/_src/utils.py:100).
 */
#ifndef FDX_B200_H_
#define FDX_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FDX_MAX_TAPS 16
#define FDX_MAX_BAND 64 /* max band rows / band columns per side */

#define FDX_OK 0
#define FDX_EINVAL (-1)      /* bad argument (null pointer, negative size, taps > FDX_MAX_TAPS ...) */
#define FDX_EUNSUPPORTED (-2) /* shape/plan combination this entry point does not cover; use fdx_axis_apply */
#define FDX_EALIGN (-3)       /* pointer not aligned to its element size */

typedef struct fdx_plan_s* fdx_plan_t;

/* Host-side description of one axis operator (see finitediffx_b200/plan.py):
 *   out[i] = sum_k main[k] * x[i + lo + k]                    nl <= i < n - nr
 *   out[i] = sum_c band_l[i*wl + c] * x[c]                    i < nl
 *   out[i] = sum_c band_r[(i-(n-nr))*wr + c] * x[n - wr + c]  i >= n - nr
 * Coefficients are unscaled float64 (solved exactly on the host; replaces utils.py:100-111). */
typedef struct {
  int32_t n;
  int32_t lo, hi;
  double main_taps[FDX_MAX_TAPS];
  int32_t nl, wl, nr, wr;
  const double* band_l; /* HOST pointer, nl*wl, row major (may be NULL when nl == 0) */
  const double* band_r; /* HOST pointer, nr*wr, row major (may be NULL when nr == 0) */
} fdx_axis_desc;

/* Plan lifetime.  create/destroy allocate and copy (cudaMalloc + cudaMemcpy): call them outside
 * the hot path and outside stream capture.  A plan is bound to the device current at creation. */
int fdx_plan_create(fdx_plan_t* plan, const fdx_axis_desc* desc);
int fdx_plan_destroy(fdx_plan_t plan);

/* ---- 1-D axis operator: replaces `difference` (finite_diff.py:168-297) and, with a transposed
 * plan, its adjoint.  The array is viewed as (outer, n, inner), n == plan n, inner contiguous.
 *   out = alpha * D x            accumulate == 0
 *   out = out + alpha * D x      accumulate != 0                                             */
int fdx_axis_apply_f32(fdx_plan_t plan, const float* x, float* out, int64_t outer, int64_t inner,
                       double alpha, int accumulate, void* stream);
int fdx_axis_apply_f64(fdx_plan_t plan, const double* x, double* out, int64_t outer, int64_t inner,
                       double alpha, int accumulate, void* stream);
/* The same pass storing its result in up to three arrays of the same shape (out2 / out3 may be NULL): the repeated
 * entries H[i][j] = F[i + j] of the reference's hessian (finite_diff.py:524-529) without copy passes. */
int fdx_axis_apply_dup_f32(fdx_plan_t plan, const float* x, float* out, float* out2, float* out3, int64_t outer,
                           int64_t inner, double alpha, void* stream);
int fdx_axis_apply_dup_f64(fdx_plan_t plan, const double* x, double* out, double* out2, double* out3, int64_t outer,
                           int64_t inner, double alpha, void* stream);

/* ---- fused 3-D operators on a C-contiguous (n0, n1, n2) grid; plans[k] acts along axis k and
 * plans[k]->n == n_k.  One pass: every input plane is read once, every output written once.
 *
 * Slab arguments (multi-GPU halo exchange or host streaming along axis 0; NULL = the whole grid):
 * the plans describe the GLOBAL grid (plans[0]->n global planes); the caller's buffers hold only a
 * window of planes:
 *   z_begin,z_end      : output planes [z_begin, z_end) (global indices) are computed and written;
 *   in_first,in_planes : every INPUT pointer designates global plane in_first and in_planes planes
 *                        are addressable behind it.  The call reads planes [z_begin-P, z_end+P)
 *                        clipped to the grid (P = stencil radius) and, when the range touches a
 *                        physical boundary, the planes under the one-sided boundary stencils;
 *                        all of them must lie inside the window (FDX_EINVAL otherwise);
 *   out_first          : every OUTPUT pointer designates global plane out_first;
 *   z_begin2,z_end2    : an optional second, disjoint range of output planes served by the same launch.        */
typedef struct {
  int32_t z_begin, z_end;
  int32_t in_first, in_planes;
  int32_t out_first;
  /* optional SECOND range of output planes in the same launch (z_begin2 == z_end2: none), disjoint from the
   * first: the two P-plane shells of a slab next to its two cuts go out as ONE launch behind the halo exchange */
  int32_t z_begin2, z_end2;
} fdx_slab;

/* laplacian (finite_diff.py:532-575):  out = sum_k alpha[k] * D_k x                         */
int fdx_laplacian3d_f32(const fdx_plan_t plans[3], const float* x, float* out, const double alpha[3],
                        const fdx_slab* slab, void* stream);
int fdx_laplacian3d_f64(const fdx_plan_t plans[3], const double* x, double* out, const double alpha[3],
                        const fdx_slab* slab, void* stream);

/* divergence (finite_diff.py:404-457): out = sum_k alpha[k] * D_k x_k                       */
int fdx_divergence3d_f32(const fdx_plan_t plans[3], const float* const x[3], float* out,
                         const double alpha[3], const fdx_slab* slab, void* stream);
int fdx_divergence3d_f64(const fdx_plan_t plans[3], const double* const x[3], double* out,
                         const double alpha[3], const fdx_slab* slab, void* stream);

/* gradient (finite_diff.py:300-351):   out_k = alpha[k] * D_k x
 * (with transposed plans this is also the adjoint of divergence)                            */
int fdx_gradient3d_f32(const fdx_plan_t plans[3], const float* x, float* const out[3],
                       const double alpha[3], const fdx_slab* slab, void* stream);
int fdx_gradient3d_f64(const fdx_plan_t plans[3], const double* x, double* const out[3],
                       const double alpha[3], const fdx_slab* slab, void* stream);

/* curl (finite_diff.py:609-668):
 *   out_0 = alpha[1] D_1 x_2 - alpha[2] D_2 x_1
 *   out_1 = alpha[2] D_2 x_0 - alpha[0] D_0 x_2
 *   out_2 = alpha[0] D_0 x_1 - alpha[1] D_1 x_0
 * (with transposed plans and negated alpha this is also its own adjoint)                    */
int fdx_curl3d_f32(const fdx_plan_t plans[3], const float* const x[3], float* const out[3],
                   const double alpha[3], const fdx_slab* slab, void* stream);
int fdx_curl3d_f64(const fdx_plan_t plans[3], const double* const x[3], double* const out[3],
                   const double alpha[3], const fdx_slab* slab, void* stream);

/* ---- fused 2-D operators on a C-contiguous (n0, n1) grid; plans[k] acts along axis k.  One pass: every input is read
 * once, every output written once (the reference: two `difference` passes plus an add / stack, finite_diff.py:338-351,
 * 443-453, 565-575, 586-606).  FDX_EUNSUPPORTED (use fdx_axis_apply) for stencil reach > 6, n1 not a multiple of the
 * 16-byte vector, unaligned pointers and dense tiny-n plans.
 *   laplacian : out = alpha[0] D_0 x + alpha[1] D_1 x
 *   divergence: out = alpha[0] D_0 x[0] + alpha[1] D_1 x[1]      (2-D curl: x = {F_1, F_0}, alpha[1] negated)
 *   gradient  : out[0] = alpha[0] D_0 x, out[1] = alpha[1] D_1 x  (with transposed plans: the adjoint of divergence) */
int fdx_laplacian2d_f32(const fdx_plan_t plans[2], const float* x, float* out, const double alpha[2], void* stream);
int fdx_laplacian2d_f64(const fdx_plan_t plans[2], const double* x, double* out, const double alpha[2], void* stream);
int fdx_divergence2d_f32(const fdx_plan_t plans[2], const float* const x[2], float* out, const double alpha[2], void* stream);
int fdx_divergence2d_f64(const fdx_plan_t plans[2], const double* const x[2], double* out, const double alpha[2], void* stream);
int fdx_gradient2d_f32(const fdx_plan_t plans[2], const float* x, float* const out[2], const double alpha[2], void* stream);
int fdx_gradient2d_f64(const fdx_plan_t plans[2], const double* x, double* const out[2], const double alpha[2], void* stream);

/* ---- fused 3-D hessian of a scalar field (finite_diff.py:460-529), one pass: 1 read + 9 writes per point.
 * d1[k] / d2[k]: the first- / second-derivative plan along axis k (d2[1] is not used and may be NULL: the reference's
 * table F[ax1 + ax2] (:499, :510) lets the mixed entry (0, 2) replace the diagonal entry of axis 1).
 * h is the (3, 3, n0, n1, n2) result; every slot H[i][j] receives F[i + j] (:524-529) with
 *   F[0] = a2[0] D2_0 x, F[1] = a1[0] a1[1] D_1 D_0 x, F[2] = a1[0] a1[2] D_2 D_0 x, F[3] = a1[1] a1[2] D_2 D_1 x,
 *   F[4] = a2[2] D2_2 x    (nested differences keep the boundary rows of both plans, as `difference(difference(.))` does).
 * FDX_EUNSUPPORTED (compose fdx_axis_apply_dup instead) for first-derivative reach > 3 (second-derivative reach > 4), n2 not a multiple of the 16-byte vector,
 * unaligned pointers and dense tiny-n plans. */
int fdx_hessian3d_f32(const fdx_plan_t d1[3], const fdx_plan_t d2[3], const float* x, float* h, const double a1[3],
                      const double a2[3], void* stream);
int fdx_hessian3d_f64(const fdx_plan_t d1[3], const fdx_plan_t d2[3], const double* x, double* h, const double a1[3],
                      const double a2[3], void* stream);

/* the 2-D form (ndim = 2: F[0] = a2[0] D2_0 x, F[1] = a1[0] a1[1] D_1 D_0 x, F[2] = a2[1] D2_1 x; h is (2, 2, n0, n1) with
 * H = [[F0, F1], [F1, F2]]): one pass, 1 read + 4 writes per point.  Same coverage rules. */
int fdx_hessian2d_f32(const fdx_plan_t d1[2], const fdx_plan_t d2[2], const float* x, float* h, const double a1[2],
                      const double a2[2], void* stream);
int fdx_hessian2d_f64(const fdx_plan_t d1[2], const fdx_plan_t d2[2], const double* x, double* h, const double a1[2],
                      const double a2[2], void* stream);

/* ---- scale read on the device.  The reference traces `step_size` (finite_diff.py:168), so inside a jitted caller
 * alpha = 1 / step_size**derivative is a DEVICE value; an XLA custom call must not fetch it.  `alpha_dev` is a device
 * pointer to float64: one value for the axis operator (out = sign * alpha_dev[0] * D x), three (one per axis) for the
 * 3-D operators.  These calls only enqueue kernels (no synchronisation, capture-safe).  The 3-D forms run as a
 * composition of axis passes (the single-pass fused kernels fold alpha into their taps on the host); callers with
 * concrete step sizes use the entry points above. */
int fdx_axis_apply_dev_alpha_f32(fdx_plan_t plan, const float* x, float* out, int64_t outer, int64_t inner,
                                 const double* alpha_dev, double sign, int accumulate, void* stream);
int fdx_axis_apply_dev_alpha_f64(fdx_plan_t plan, const double* x, double* out, int64_t outer, int64_t inner,
                                 const double* alpha_dev, double sign, int accumulate, void* stream);
int fdx_laplacian3d_dev_alpha_f32(const fdx_plan_t plans[3], const float* x, float* out, const double* alpha_dev, void* stream);
int fdx_laplacian3d_dev_alpha_f64(const fdx_plan_t plans[3], const double* x, double* out, const double* alpha_dev, void* stream);
int fdx_divergence3d_dev_alpha_f32(const fdx_plan_t plans[3], const float* const x[3], float* out, const double* alpha_dev, void* stream);
int fdx_divergence3d_dev_alpha_f64(const fdx_plan_t plans[3], const double* const x[3], double* out, const double* alpha_dev, void* stream);
int fdx_gradient3d_dev_alpha_f32(const fdx_plan_t plans[3], const float* x, float* const out[3], const double* alpha_dev, void* stream);
int fdx_gradient3d_dev_alpha_f64(const fdx_plan_t plans[3], const double* x, double* const out[3], const double* alpha_dev, void* stream);
int fdx_curl3d_dev_alpha_f32(const fdx_plan_t plans[3], const float* const x[3], float* const out[3], const double* alpha_dev, void* stream);
int fdx_curl3d_dev_alpha_f64(const fdx_plan_t plans[3], const double* const x[3], double* const out[3], const double* alpha_dev, void* stream);

/* ---- introspection */
int fdx_version(void);                 /* 1000*major + minor */
const char* fdx_error_string(int code); /* static string for FDX_E* and cudaError_t values */
/* number of kernel launches issued through this library since load (bench.py's gpu_launches) */
int64_t fdx_launch_count(void);
/* name of the kernel family the last call on this thread dispatched to ("axis_stream", ...) */
const char* fdx_last_kernel(void);

/* The FDX_* environment switches of the fused kernels (experiments, bit-equality tests) are snapshotted when the
 * library is first used; this takes a new snapshot.  Test / tooling hook, not part of the hot path. */
void fdx_reload_env(void);

#ifdef __cplusplus
}
#endif
#endif /* FDX_B200_H_ */
