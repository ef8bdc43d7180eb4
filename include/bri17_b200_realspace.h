/*
 * bri17_b200_realspace.h -- C ABI of libbri17_b200_rs.so: the end-to-end
 * real-space operator and the CG solve built on it.
 *
 * Reference interface replaced (file:line in the bri17 repository):
 *   tests/test_bri17.cpp:56-107   StiffnessMatrixFactory::compute_Ku
 *        F = (|h|/|N|) * iDFT_unnormalised( K^ . DFT(u) ),  DIM forward and DIM
 *        backward c2c transforms on the planar component blocks (:117-127,
 *        FFTW_FORWARD / FFTW_BACKWARD), correction |h|/|N| (:93-106)
 *   sphinx/theory.rst:60,72,151-157  DFT sign/scale conventions, eqs (1),(3),(11)-(12)
 *   python/demo.py:11-40 + bri17.hpp:336-341   the periodic inclusion problem
 *        K^ u^ = tau^ . conj(B^), u^(0) = 0, solved here matrix-free by CG.
 *
 * FFTW (serial, host) becomes cuFFT (local transforms) plus a slab
 * decomposition over `nranks` GPUs, one process per GPU:
 *
 *   real space   u[c][n0 in my slab][n1][n2]          (slab over axis 0)
 *   Fourier      u^[c][k0][k1 in my slab][k2]         (slab over axis 1)
 *
 * forward = local FFT over axes 1.. -> all-to-all (slab transpose over
 * NVLink) -> FFT along axis 0; the modal operator runs directly on the
 * Fourier-side block (k_begin = {0, k1_begin, 0}), then the inverse path.
 * The all-to-all is either NCCL grouped send/recv with a pack/unpack kernel
 * on the strided side (mode 0), or ONE kernel per direction that reads the
 * local slab and stores straight into the peers' buffers through CUDA-IPC
 * mapped memory, so that the transposition is the transfer (mode 1).
 *
 * Layout, ownership and error conventions are those of bri17_b200.h.  Real
 * fields are carried as complex numbers with zero imaginary part, exactly
 * like the reference (tests/test_bri17.cpp:133-136).
 */
#ifndef BRI17_B200_REALSPACE_H
#define BRI17_B200_REALSPACE_H

#include <stdint.h>

#include "bri17_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bri17_rs_plan bri17_rs_plan;

#define BRI17_NCCL_UNIQUE_ID_BYTES 128

/* Rank 0 calls this and distributes the 128 bytes to the other ranks (any
 * transport: torch.distributed, MPI, a file); not needed when nranks == 1. */
BRI17_API int bri17_rs_unique_id(void *out128);

/*
 * Collective over the `nranks` processes.  shape[0] and shape[1] are split
 * into contiguous balanced slabs (rank g owns [g*N/P, (g+1)*N/P)).
 * exchange_mode: 0 = NCCL send/recv + pack kernels, 1 = fused peer-store kernel.
 */
BRI17_API int bri17_rs_plan_create(bri17_rs_plan **out, int dim, const int *shape, const double *L,
                                   double mu, double nu, int device, int rank, int nranks,
                                   const void *nccl_unique_id, int exchange_mode);
BRI17_API int bri17_rs_plan_destroy(bri17_rs_plan *plan);

/* Tuning knobs: "pipeline" (1 = overlap the exchange with the local transforms on a second
 * stream, sub-slab by sub-slab; default 1, fused exchange only), "exchange_chunks" (sub-slabs of
 * n0 planes per component in that pipeline: 1..4, 0 = by size, at least 128 MiB each; default 0),
 * "fft_chunk_mib" (> 0: run the local 2-D transforms in chunks of planes of this size so that
 * cuFFT's second kernel could hit L2; measured slower at every size, default 0 = whole slab;
 * "fft_chunk_planes": the same in planes),
 * "copy_ctas" (grid cap of the exchange kernel; 0 = default: 4 CTAs per SM when it runs alone, 2 per SM
 * (complex) or 2/3 per SM (half spectrum) in the pipelined apply, where it shares the SMs with cuFFT), "fused_axis0" (1 = run FFT(axis 0) ->
 * K^ -> inverse FFT(axis 0) as one kernel when shape[0] is 16..1024 and a power of two;
 * default 1; 0 = cuFFT + modal kernel + cuFFT, the only path for other lengths),
 * "k1_major" (with the fused pass in 3-D, keep the Fourier-side block as [c][k1][n0][k2] so that
 * the rows a tile gathers are S2e*16 bytes apart instead of n1*S2e*16: 1 always, 0 never, -1
 * (default) whenever the fused exchange writes it, and on one GPU for complex fields only). */
BRI17_API int bri17_rs_plan_set_option(bri17_rs_plan *plan, const char *key, int64_t value);
/* "fused_axis0" (is the fused pass in use), "k1_major" / "k1_major_real" (is the k1-major layout
 * in use for complex / real fields), "fused_launches", "pipeline", "exchange_mode", "barriers"
 * (flag barriers issued so far). */
BRI17_API int bri17_rs_plan_get_info(const bri17_rs_plan *plan, const char *key, int64_t *value);

/* Geometry of this rank: real-space slab [n0_begin, n0_begin+n0_count) of axis
 * 0, Fourier-space slab [k1_begin, k1_begin+k1_count) of axis 1. */
BRI17_API int bri17_rs_plan_local(const bri17_rs_plan *plan, int *n0_begin, int *n0_count,
                                  int *k1_begin, int *k1_count);
/* Complex elements per component of the real-space slab / the Fourier slab. */
BRI17_API int64_t bri17_rs_plan_real_count(const bri17_rs_plan *plan);
BRI17_API int64_t bri17_rs_plan_fourier_count(const bri17_rs_plan *plan);

/*
 * F = (|h|/|N|) iDFT( K^ DFT(u) )   (tests/test_bri17.cpp:56-107)
 * u_dev, F_dev: [dim][n0_count][N1][(N2)] complex128, distinct buffers; F is
 * also used as scratch.  Collective; asynchronous on `stream`.  A rank whose slab
 * is empty (shape[0] < nranks) passes NULL fields and must still make the call.
 */
BRI17_API int bri17_real_space_apply_f64(bri17_rs_plan *plan, const void *u_dev, void *F_dev,
                                         void *stream);

/*
 * Same operator for REAL fields stored as plain doubles, [dim][n0_count][N1][(N2)]
 * (the reference only ever applies compute_Ku to real data carried as complex,
 * tests/test_bri17.cpp:133-136).  r2c over the trailing axes keeps the
 * non-redundant half spectrum of the last axis (K^(N-k) = K^(k)): half the FFT
 * work, half the all-to-all bytes, half the modal traffic.  u_dev is preserved.
 */
BRI17_API int bri17_real_space_apply_real_f64(bri17_rs_plan *plan, const void *u_dev, void *F_dev,
                                              void *stream);

/* The two halves, exposed for callers that work in Fourier space:
 * forward: x_dev (real-space slab, preserved) -> x_hat_dev (Fourier slab
 *          [dim][N0][k1_count][(N2)]), unnormalised, sign -1 (theory.rst:60);
 * inverse: x_hat_dev (destroyed) -> x_dev, multiplied by `scale`
 *          (pass 1/|N| for theory.rst:72). ncomp components (dim or 6, ...). */
BRI17_API int bri17_rs_forward_fft_f64(bri17_rs_plan *plan, const void *x_dev, void *x_hat_dev,
                                       int ncomp, void *stream);
BRI17_API int bri17_rs_inverse_fft_f64(bri17_rs_plan *plan, void *x_hat_dev, void *x_dev,
                                       int ncomp, double scale, void *stream);

/* The modal plan of the Fourier-side block (k_begin/local_shape helpers). */
BRI17_API bri17_plan *bri17_rs_plan_modal(bri17_rs_plan *plan);

/* Milliseconds of the phases of the last real-space apply, after a stream
 * synchronisation: [0] local forward FFT, [1] forward exchange (pack + all-to-
 * all), [2] axis-0 forward FFT, [3] modal apply, [4] axis-0 inverse FFT,
 * [5] backward exchange, [6] local inverse FFT, [7] total.  n <= 8. */
BRI17_API int bri17_rs_plan_last_timings(bri17_rs_plan *plan, double *ms, int n);
/* Bytes this rank sends to other ranks in one exchange (one direction), for the
 * complex (real_layout = 0) or the half-spectrum (real_layout = 1) path. */
BRI17_API int64_t bri17_rs_plan_exchange_bytes(const bri17_rs_plan *plan, int real_layout);

/*
 * F = A u as above AND the global scalar <u, A u> (summed over ranks, written to
 * *dot_host after a stream synchronisation): the Parseval sum  sum_k Re(u^_k^H f^_k)
 * is accumulated by the kernel that applies K^, so no extra pass over u and F is
 * needed.  real_fields: 0 = complex128 fields, 1 = float64 fields (r2c path).
 */
BRI17_API int bri17_real_space_apply_dot_f64(bri17_rs_plan *plan, const void *u_dev, void *F_dev,
                                             int real_fields, double *dot_host, void *stream);

/*
 * Matrix-free conjugate gradients on  A x = b,  A = the real-space operator
 * above (symmetric positive semi-definite; null space = constant fields).  The
 * component means of b are projected out first (K^(0) = 0: the reference skips
 * the null frequency, bri17.hpp:336-339) and the zero-mean solution is returned
 * -- the u^(0) = 0 choice of theory.rst:208-212.
 * b_dev, x_dev: real-space slabs (x is overwritten, start from 0).
 * Stops when |r| <= rtol*|b| or after max_iter iterations; all scalars stay
 * on the device (no host synchronisation inside an iteration except every
 * `check_every` iterations for the stopping test).  Per iteration: one
 * operator application (which also returns <p, A p>), r -= alpha A p with
 * <r, r>, then x += alpha p and p = r + beta p in one pass; two scalar
 * all-reduces.  Returns BRI17_ERR_BREAKDOWN if the residual stops being finite.
 */
BRI17_API int bri17_cg_solve_f64(bri17_rs_plan *plan, const void *b_dev, void *x_dev, double rtol,
                                 int max_iter, int check_every, int *iterations,
                                 double *rel_residual, void *stream);

/* Same on real fields (plain doubles) through bri17_real_space_apply_real_f64. */
BRI17_API int bri17_cg_solve_real_f64(bri17_rs_plan *plan, const void *b_dev, void *x_dev, double rtol,
                                      int max_iter, int check_every, int *iterations,
                                      double *rel_residual, void *stream);

/* Test aid, HOST memory, no device needed: replays the fused axis-0 kernel
 * (FFT(axis 0) -> K^ * out_scale -> inverse FFT(axis 0)) thread by thread on the CPU,
 * in place on X_host[dim][N0][S] (complex128).  Column j is (k1, k2) = (k1_begin +
 * j / S2e, j % S2e) in 3-D, k1 = k1_begin + j in 2-D.  tab_d: phi|chi|psi of axis d
 * ([3][N_d], bri17_plan_get_tables).  k1_major (3-D): X_host is [dim][S/S2e][N0][S2e] instead of
 * [dim][N0][S] (the layout the operator keeps its Fourier-side block in, see
 * bri17_rs_plan_set_option "k1_major").  dot_out (optional): sum_k w_k Re(u^H f). */
BRI17_API int bri17_debug_axis0_fused_host(int dim, int N0, int64_t S, int S2e, int k1_begin, int N1,
                                           int N2, const double *tab0, const double *tab1,
                                           const double *tab2, double mu, double nu, double out_scale,
                                           int hermitian_n, int k1_major, void *X_host, double *dot_out);

/* Test aid, HOST memory, no device needed: replays the fused all-to-all exchange of `nranks` virtual ranks on the
 * CPU with the copy plans the GPU path builds.  direction 0: in[r] = [dim][n0_r][S1][S2e] (local-transform layout
 * of rank r) -> out[q] = Fourier-side block of rank q, [dim][N0][n1_q][S2e] or, k1_major, [dim][n1_q][N0][S2e];
 * direction 1: back.  real_layout: S = N/2+1 along the last axis.  nchunks: sub-slabs per component (1..4). */
BRI17_API int bri17_debug_exchange_host(int dim, const int *shape, int nranks, int real_layout, int k1_major,
                                        int nchunks, int direction, void *const *in, void *const *out);

#ifdef __cplusplus
}
#endif
#endif /* BRI17_B200_REALSPACE_H */
