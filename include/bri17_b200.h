/*
 * bri17_b200.h -- C ABI of libbri17_b200.so, the B200 (sm_100a) implementation
 * of the bri17 modal operator path.
 *
 * This is the drop-in boundary: plain pointers and sizes, `int` status codes,
 * no C++/torch types.  The C++ header include/bri17/bri17.hpp
 * (bri17::ModalOperator) and the Python mirror (bri17_b200/) are thin wrappers
 * over exactly these entry points; INTEGRATION.md shows the binding a bri17
 * maintainer would add.
 *
 * Reference interfaces replaced (file:line relative to the bri17 repository):
 *   include/bri17/bri17.hpp:34-58    CartesianGrid<T,DIM>{shape, L}   -> bri17_plan_create(dim, shape, L, ...)
 *   include/bri17/bri17.hpp:176-194  Hooke<T,DIM>{mu, nu, grid}       -> bri17_plan_create(..., mu, nu, ...)
 *   include/bri17/bri17.hpp:247-292  Hooke::modal_stiffness(k, K)     -> bri17_modal_stiffness_mode_f64 (one mode, host)
 *                                                                        bri17_modal_stiffness_field_f64 (every mode, device)
 *   include/bri17/bri17.hpp:212-236  Hooke::modal_strain_displacement -> bri17_modal_strain_displacement_mode_f64 / _field_f64
 *   tests/test_bri17.cpp:58-92       loop nest: gather, K^[k]*u^[k], scatter
 *                                                                     -> bri17_modal_stiffness_apply_f64 (device buffers)
 *                                                                        bri17_modal_stiffness_apply_host_f64 (host buffers)
 *   tests/test_bri17.cpp:62-64,71 / :76-79,88   frequency <-> linear index  -> bri17_freq_index_map
 *   tests/test_bri17.cpp:194-235     strain recovery loop (compute_Bu) -> bri17_strain_displacement_apply_f64
 *   tests/test_bri17.cpp:56-107      real-space apply (FFT, K^, iFFT, |h|/|N|) -> bri17_real_space_apply_f64
 *
 * Data layout (tests/test_bri17.cpp:66-67, :81-83): a field is PLANAR by
 * component; each component is a row-major block of interleaved complex
 * doubles (re, im) = std::complex<double> = 16 bytes.  Element (c, i) of a
 * field lives at base[i + c*comp_stride] (in complex elements).  A "block" is
 * the set of frequencies k = k_begin + [0, local_shape) -- the whole grid for
 * the reference's use (k_begin = 0, local_shape = shape), a slab of it for
 * multi-GPU runs.  Frequencies are raw indices in [0, N_d): no fftshift, no
 * negative wrap (bri17.hpp:260 uses k as is).
 *
 * Ownership: the caller owns and preallocates every buffer (as in
 * bri17.hpp:207-211, :241-246).  A plan owns only its per-axis tables and, for
 * the *_host_* entry point, its staging buffers.  Input and output may alias.
 *
 * Threading: a plan is immutable after creation as far as the device entry
 * points are concerned (the reference's methods are const and reentrant,
 * bri17.hpp:212,247); any number of host threads may launch on one plan
 * concurrently, each on its own stream.  The *_host_* entry point shares one
 * staging set per plan: concurrent calls on the same plan are serialised by a
 * mutex inside the plan.  bri17_plan_set_option must not race with launches.
 *
 * Errors: every function returns BRI17_OK or an error code and never throws;
 * bri17_last_error() returns a thread-local message.  There is no CPU
 * fallback: without a CUDA device plan creation fails with BRI17_ERR_CUDA.
 */
#ifndef BRI17_B200_H
#define BRI17_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define BRI17_API __attribute__((visibility("default")))
#else
#define BRI17_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define BRI17_VERSION 100 /* 0.1.0, tracks metadata/version.txt of the reference */

enum {
  BRI17_OK = 0,
  BRI17_ERR_INVALID_ARG = 1, /* bad dim/shape/pointer/alignment: std::invalid_argument in the C++ wrapper */
  BRI17_ERR_CUDA = 2,        /* CUDA runtime failure: std::runtime_error */
  BRI17_ERR_NCCL = 3,        /* NCCL failure (real-space apply only) */
  BRI17_ERR_UNSUPPORTED = 4,
  BRI17_ERR_BREAKDOWN = 5    /* iterative solve broke down (non-finite residual): std::runtime_error */
};

typedef struct bri17_plan bri17_plan;

/* `device` value for a plan that only serves the per-mode HOST API
 * (bri17_modal_*_mode_f64, bri17_plan_get_tables).  Every whole-grid entry
 * point fails with BRI17_ERR_CUDA on such a plan: there is no CPU fallback. */
#define BRI17_DEVICE_NONE (-1)

/* Thread-local description of the last failure on the calling thread. */
BRI17_API const char *bri17_last_error(void);
BRI17_API int bri17_version(void);
/* Sets that message; used by companion libraries (libbri17_b200_rs.so) so that
 * one bri17_last_error() serves every entry point. */
BRI17_API void bri17_set_last_error(const char *msg);

/*
 * Create the operator for CartesianGrid{shape, L} + Hooke{mu, nu} on CUDA
 * device `device` (bri17.hpp:54, :193).  dim is 2 or 3.  Builds the per-axis
 * tables phi/chi/psi (bri17.hpp:259-263) and c/s (bri17.hpp:218-221) on the
 * host with libm, in the reference's operation order, and uploads them
 * (phi/chi/psi, c/s, alpha and sin(alpha) per grid line: K^ itself is never materialised).
 */
BRI17_API int bri17_plan_create(bri17_plan **out, int dim, const int *shape,
                      const double *L, double mu, double nu, int device);
BRI17_API int bri17_plan_destroy(bri17_plan *plan);

/* Tuning/diagnostic knobs ("apply_variant", "host_chunk_rows", ...). */
BRI17_API int bri17_plan_set_option(bri17_plan *plan, const char *key, int64_t value);
BRI17_API int bri17_plan_get_info(const bri17_plan *plan, const char *key, int64_t *value);

/* Host copy of the tables of one axis; any output pointer may be NULL.
 * Each array has shape[axis] entries. */
BRI17_API int bri17_plan_get_tables(const bri17_plan *plan, int axis, double *phi,
                          double *chi, double *psi, double *c, double *s);

/* ---- one frequency, host side (the reference's per-mode API) ------------- */

/* Hooke::modal_stiffness (bri17.hpp:247-292): K[dim*dim] interleaved complex,
 * row-major K[dim*i+j], imaginary parts zero.  k[d] must lie in [0, shape[d]). */
BRI17_API int bri17_modal_stiffness_mode_f64(const bri17_plan *plan, const int *k, double *K);
/* Hooke::modal_strain_displacement (bri17.hpp:212-236): B[dim] interleaved. */
BRI17_API int bri17_modal_strain_displacement_mode_f64(const bri17_plan *plan, const int *k, double *B);

/* Hooke::modal_eigenstress_to_opposite_strain (bri17.hpp:308-355): tau[nsym] ->
 * eta[nsym], Mandel notation (nsym = 3 in 2-D, 6 in 3-D), zero at k = 0.  The
 * reference solves K^ u = tau.conj(B^) with Eigen's LLT; this build uses its own
 * Cholesky (no reference test pins the result: parity unpinned). */
BRI17_API int bri17_modal_eigenstress_to_opposite_strain_mode_f64(const bri17_plan *plan, const int *k,
                                                                  const double *tau, double *eta);

/* ---- every frequency of a block, device side ----------------------------- */

/*
 * f^[c,k] = out_scale * sum_j K^[k][c,j] u^[j,k]   (tests/test_bri17.cpp:58-92)
 *
 * u_hat_dev, f_hat_dev: device pointers, 16-byte aligned, planar layout above.
 * k_begin, local_shape: dim ints each, or NULL for the whole grid.
 * comp_stride: complex elements between components; 0 means prod(local_shape).
 * out_scale: 1.0 reproduces the reference bit for bit (no multiply is issued);
 *            the real-space apply passes |h|/|N| (tests/test_bri17.cpp:98).
 * stream: a cudaStream_t (NULL = default stream).  Asynchronous.
 */
BRI17_API int bri17_modal_stiffness_apply_f64(bri17_plan *plan, const void *u_hat_dev,
                                    void *f_hat_dev, const int *k_begin,
                                    const int *local_shape, int64_t comp_stride,
                                    double out_scale, void *stream);

/*
 * bri17_modal_stiffness_apply_f64 that ALSO returns, in the device scalar *dot_dev,
 *     sum_k w_k Re( u^[k]^H f^[k] )        (f^ including out_scale)
 * over the block (stream-ordered; deterministic summation order).  By Parseval this is |N| <u, F>
 * for the real-space fields, so a CG iteration gets <p, A p> from the operator application itself
 * (tests/test_bri17.cpp:56-107 has no such product; this is what the CG of BASELINE config 5
 * needs).  hermitian_n = 0: every mode counts once (w_k = 1).  hermitian_n = N > 0: the fastest
 * axis of the block is the half spectrum k <= N/2 of a real field of length N (K^(N-k) = K^(k)):
 * modes with 0 < k < N/2 stand for a conjugate pair, w_k = 2.  scratch_dev: scratch_count >= 1
 * doubles (one partial sum per CTA; the grid is capped at scratch_count, 1184 keeps it full).
 */
BRI17_API int bri17_modal_stiffness_apply_dot_f64(bri17_plan *plan, const void *u_hat_dev, void *f_hat_dev,
                                                  const int *k_begin, const int *local_shape,
                                                  int64_t comp_stride, double out_scale, int hermitian_n,
                                                  double *dot_dev, double *scratch_dev, int scratch_count,
                                                  void *stream);

/*
 * Same operation on HOST buffers: the block is cut along its slowest axis into
 * chunks that are copied in, processed and copied out on rotating streams so
 * that both PCIe directions and the kernel overlap.  Synchronous.  Pinned
 * (page-locked) buffers give full PCIe bandwidth; pageable memory works too.
 */
BRI17_API int bri17_modal_stiffness_apply_host_f64(bri17_plan *plan, const void *u_hat_host,
                                         void *f_hat_host, const int *k_begin,
                                         const int *local_shape, int64_t comp_stride,
                                         double out_scale);

/* K^[k] for every mode of the block, mode-major: K_dev[(i*dim*dim + r*dim + j)]
 * complex, i the row-major linear index in the block (what a loop over
 * Hooke::modal_stiffness would write).  144 B/mode in 3-D: diagnostic only. */
BRI17_API int bri17_modal_stiffness_field_f64(bri17_plan *plan, void *K_dev, const int *k_begin,
                                    const int *local_shape, void *stream);

/* B^[k] for every mode of the block, mode-major: B_dev[i*dim + j] complex. */
BRI17_API int bri17_modal_strain_displacement_field_f64(bri17_plan *plan, void *B_dev,
                                              const int *k_begin, const int *local_shape,
                                              void *stream);

/*
 * eps^ = 1/2 (B^ (x) u^ + u^ (x) B^) in Mandel order (tests/test_bri17.cpp:194-235):
 * 2-D [00, 11, sqrt2*01], 3-D [00, 11, 22, sqrt2*12, sqrt2*20, sqrt2*01];
 * planar output with eps_stride complex elements between components (0 = dense).
 */
BRI17_API int bri17_strain_displacement_apply_f64(bri17_plan *plan, const void *u_hat_dev,
                                        void *eps_hat_dev, const int *k_begin,
                                        const int *local_shape, int64_t u_stride,
                                        int64_t eps_stride, double out_scale, void *stream);

/*
 * Per-mode direct solves, every mode of the block (bri17.hpp:308-355 batched;
 * the reference drives them from a Python double loop, python/demo.py:33-40).
 * Element (s, i) of a field lives at base[i*mode_stride + s*comp_stride]:
 *   planar      mode_stride = 1, comp_stride >= modes   (0, 0 selects this)
 *   mode-major  mode_stride = ncomp, comp_stride = 1     (python/demo.py:21,37-38)
 *
 * bri17_modal_stiffness_solve_f64:           u^ = K^-1 f^, u^(0) = 0 (theory.rst:208-212); dim -> dim
 * bri17_eigenstress_to_displacement_f64:     u^ = K^-1 (tau^ . conj(B^))  (bri17.hpp:340-341); nsym -> dim
 * bri17_eigenstress_to_opposite_strain_f64:  eta^ = sym(B^ (x) u^), Mandel (bri17.hpp:342-353); nsym -> nsym
 */
BRI17_API int bri17_modal_stiffness_solve_f64(bri17_plan *plan, const void *f_hat_dev, void *u_hat_dev,
                                              const int *k_begin, const int *local_shape,
                                              int64_t comp_stride, int64_t mode_stride, void *stream);
BRI17_API int bri17_eigenstress_to_displacement_f64(bri17_plan *plan, const void *tau_hat_dev,
                                                    void *u_hat_dev, const int *k_begin,
                                                    const int *local_shape, int64_t tau_comp_stride,
                                                    int64_t tau_mode_stride, int64_t u_comp_stride,
                                                    int64_t u_mode_stride, void *stream);
BRI17_API int bri17_eigenstress_to_opposite_strain_f64(bri17_plan *plan, const void *tau_hat_dev,
                                                       void *eta_hat_dev, const int *k_begin,
                                                       const int *local_shape, int64_t comp_stride,
                                                       int64_t mode_stride, void *stream);
/* f^ = tau^ . conj(B^) per mode (bri17.hpp:324-332, :340: the right-hand side of the solve above,
 * without the solve); nsym -> dim, zero at k = 0.  b = (|h|/|N|) iDFT_unnormalised(f^) is the nodal
 * force (theory.rst:151-157) of the periodic inclusion problem of python/demo.py:11-23, i.e. the
 * right-hand side of A x = b that bri17_cg_solve_f64 solves matrix-free (BASELINE config 5). */
BRI17_API int bri17_eigenstress_to_force_f64(bri17_plan *plan, const void *tau_hat_dev, void *f_hat_dev,
                                             const int *k_begin, const int *local_shape,
                                             int64_t tau_comp_stride, int64_t tau_mode_stride,
                                             int64_t f_comp_stride, int64_t f_mode_stride, void *stream);

/* Frequency multi-index the kernels derive for every linear element of the
 * block: k_out_dev[i*dim + d] (int32).  Uses the same tile cursor as the apply
 * kernels; the parity tests require it to be bit-identical to the reference's
 * loop nest (tests/test_bri17.cpp:62-64,71 / :76-79,88). */
BRI17_API int bri17_freq_index_map(bri17_plan *plan, int32_t *k_out_dev, const int *k_begin,
                         const int *local_shape, void *stream);

/* Diagnostic, host only (works on a BRI17_DEVICE_NONE plan): replays on the CPU the
 * tile cursor that persistent CTA `cta` of a grid of at most `max_ctas` CTAs follows over
 * the block, for tiles of `tile_modes` modes.  Writes (tile, row, chunk, a, b) per visited
 * tile into out[5*i..] (up to `cap` tiles), the grid size into *grid, and returns the
 * number of tiles visited (-1 on bad arguments). */
BRI17_API int bri17_debug_walk_tiles(const bri17_plan *plan, const int *k_begin, const int *local_shape,
                                     int tile_modes, int max_ctas, int cta, int64_t *out, int cap,
                                     int *grid);

#ifdef __cplusplus
}
#endif
#endif /* BRI17_B200_H */
