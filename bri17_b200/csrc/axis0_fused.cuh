// axis0_fused.cuh -- ONE kernel for  FFT(axis 0) -> K^ . (.) * |h|/|N| -> inverse FFT(axis 0)
// on the Fourier-side block X[c][n0][column] of the real-space operator
// (tests/test_bri17.cpp:57 [axis 0 of the forward transforms], :58-92 [modal loop], :93-106
// [scale], :95 [axis 0 of the backward transforms] of the reference).
//
// Round 1 ran these as three passes over HBM (cuFFT strided Z2Z, modal kernel, cuFFT): 7.4 of the
// 15.8 ms of a 512^3 apply.  Here a CTA owns a tile of W consecutive columns x all N0 points x all
// components in shared memory (96 KiB for N0 <= 512, 192 KiB for N0 = 1024), so the block is read
// once and written once:
//
//   global --(stage 0: radix R0 butterflies in registers)--> smem --(stages 1..)--> smem
//          --(last forward stage, K^ per mode, first inverse stage: same owner thread, no barrier)-->
//   smem --(inverse stages)--> (last inverse stage stores straight to global, natural order)
//
// Forward = decimation in frequency (natural order in, digit-reversed out), inverse = the
// transposed flow graph with conjugated twiddles (digit-reversed in, natural out): no reordering
// pass, and K^ is applied in digit-reversed position with k0 recovered from the position's digits.
// Twiddles exp(-2 pi i j/N0) come from a host-built table staged in shared memory; complex
// products use explicit fma (the library is compiled with -fmad=false for the bit-exact kernels).
// The modal arithmetic is the reference's, in its written order (bri17.hpp:266-288).
//
// Shared-memory layout: data[c][n][w] (complex), w fastest.  A 128-bit access is served in
// quarter-warps (8 lanes = 128 B = all banks once).  W >= 8: the 8 lanes are 8 consecutive w of one
// n -> conflict-free.  W = 4: 8 lanes = 2 butterflies x 4 columns; in the stride-1 stage (radix 8)
// the two butterflies are 8 rows = 512 B apart -> same banks; rows are therefore stored at
// n ^ ((n >> 3) & 1), which puts the two rows in different halves of the 128 B line and leaves the
// other stages' quarter-warps contiguous.
//
// Global layout.  A tile reads N0 row segments of W*16 bytes per component.  With the natural layout
// [c][n0][k1][k2] the rows of a column are n1*S2e*16 bytes apart (4 MiB at 512^3): every row is another
// 2 MB page and another DRAM page, and a pure copy with this pattern reaches 2.3 TB/s on B200
// (tests/cpp/seg_bw.cu: 64-byte segments, 4 MiB stride) -- the first version of this kernel ran at that
// limit.  The k1-major layout [c][k1][n0][k2] puts the rows S2e*16 bytes apart (8 KiB): 5.0 TB/s for
// the same copy.  The producers write it for free: the exchange kernel stores every received row at
// (k1, n0) instead of (n0, k1); on one GPU cuFFT's advanced output layout does.
//
// Everything below the kernel is __host__ __device__: tests replay the phases thread by thread on
// the CPU (bri17_debug_axis0_fused_host) against numpy's FFT, so that index arithmetic is checked
// without a GPU.
#pragma once

#include <cuda_runtime.h>

#include <cmath>

#ifdef __CUDACC__
#define A0_HD __host__ __device__ __forceinline__
#else
#define A0_HD inline
#endif

namespace bri17b200 {
namespace axis0 {

struct Params {
  double2 *X;              // dim components of N0 x S complex, transformed in place; element (c, n0,
                           // column j) at c*comp_stride + (j / blk_cols)*blk_stride + j % blk_cols + n0*row_stride
  long long comp_stride;   // complex elements between components (>= N0*S)
  long long row_stride;    // ... between consecutive n0 of one column
  long long blk_stride;    // ... between consecutive blocks of blk_cols columns
  long long blk_cols;      // natural layout [n0][S]: blk_cols = S, row_stride = S;
                           // k1-major layout [k1][n0][S2e]: blk_cols = S2e, row_stride = S2e, blk_stride = N0*S2e
  long long S;             // columns
  long long n_tiles;       // ceil(S / W)
  int N0;
  int S2e;                 // 3-D: extent of the fastest axis in the block (column = b*S2e + k2); 2-D: 1
  int k1_begin;            // global index of the first k1 (axis-1 frequency) of the block
  const double *tab0, *tab1, *tab2;  // per-axis phi|chi|psi, [3][N_d] (tab2 unused in 2-D)
  int N1, N2;              // table lengths of axes 1 and 2
  const double2 *twiddle;  // per-stage twiddle tables (fill_twiddles), N0 complex
  double mu, scaling, out_scale;
  double *dot_partial;     // optional: one partial sum of w_k Re(u^H f) per CTA
  int herm_n;              // > 0: the fastest axis is the half spectrum of a real field of this length
};

// ---- complex helpers -------------------------------------------------------
A0_HD double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
A0_HD double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
A0_HD double fma_(double a, double b, double c) {
#ifdef __CUDA_ARCH__
  return __fma_rn(a, b, c);
#else
  return std::fma(a, b, c);
#endif
}
// a * (c, s) with s the IMAGINARY part of the factor actually applied
A0_HD double2 cmul(double2 a, double cr, double ci) {
  return make_double2(fma_(a.x, cr, -(a.y * ci)), fma_(a.x, ci, a.y * cr));
}
// forward: a * (-i); inverse: a * (+i)
template <bool INV>
A0_HD double2 rot90(double2 a) {
  return INV ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x);
}

// ---- radix-R DFT in registers, natural order in and out: a[m] <- sum_r a[r] w_R^{+-rm} -------
template <int R, bool INV>
struct Dft;

template <bool INV>
struct Dft<2, INV> {
  static A0_HD void run(double2 (&a)[2]) {
    const double2 t = a[0];
    a[0] = cadd(t, a[1]);
    a[1] = csub(t, a[1]);
  }
};

template <bool INV>
struct Dft<4, INV> {
  static A0_HD void run(double2 (&a)[4]) {
    const double2 t0 = cadd(a[0], a[2]), t1 = csub(a[0], a[2]);
    const double2 t2 = cadd(a[1], a[3]), t3 = rot90<INV>(csub(a[1], a[3]));
    a[0] = cadd(t0, t2);
    a[2] = csub(t0, t2);
    a[1] = cadd(t1, t3);
    a[3] = csub(t1, t3);
  }
};

template <bool INV>
struct Dft<8, INV> {
  static A0_HD void run(double2 (&a)[8]) {
    constexpr double h = 0.70710678118654752440;
    double2 e[4] = {a[0], a[2], a[4], a[6]}, o[4] = {a[1], a[3], a[5], a[7]};
    Dft<4, INV>::run(e);
    Dft<4, INV>::run(o);
    // o[m] *= w8^{+-m}
    const double2 o1 = INV ? make_double2(h * (o[1].x - o[1].y), h * (o[1].x + o[1].y))
                           : make_double2(h * (o[1].x + o[1].y), h * (o[1].y - o[1].x));
    const double2 o2 = rot90<INV>(o[2]);
    const double2 o3 = INV ? make_double2(-h * (o[3].x + o[3].y), h * (o[3].x - o[3].y))
                           : make_double2(h * (o[3].y - o[3].x), -h * (o[3].x + o[3].y));
    a[0] = cadd(e[0], o[0]); a[4] = csub(e[0], o[0]);
    a[1] = cadd(e[1], o1);   a[5] = csub(e[1], o1);
    a[2] = cadd(e[2], o2);   a[6] = csub(e[2], o2);
    a[3] = cadd(e[3], o3);   a[7] = csub(e[3], o3);
  }
};

template <bool INV>
struct Dft<16, INV> {
  static A0_HD void run(double2 (&a)[16]) {
    constexpr double h = 0.70710678118654752440;
    constexpr double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173;  // cos, sin(pi/8)
    double2 e[8], o[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { e[i] = a[2 * i]; o[i] = a[2 * i + 1]; }
    Dft<8, INV>::run(e);
    Dft<8, INV>::run(o);
    constexpr double sg = INV ? 1. : -1.;  // sign of the imaginary part of w16^{+-m}
    o[1] = cmul(o[1], c1, sg * s1);
    o[2] = cmul(o[2], h, sg * h);
    o[3] = cmul(o[3], s1, sg * c1);
    o[4] = rot90<INV>(o[4]);
    o[5] = cmul(o[5], -s1, sg * c1);
    o[6] = cmul(o[6], -h, sg * h);
    o[7] = cmul(o[7], -c1, sg * s1);
#pragma unroll
    for (int m = 0; m < 8; m++) { a[m] = cadd(e[m], o[m]); a[m + 8] = csub(e[m], o[m]); }
  }
};

// ---- global / shared accessors ----------------------------------------------
A0_HD double2 ld_stream(const double2 *p) {
#ifdef __CUDA_ARCH__
  return __ldcs(p);
#else
  return *p;
#endif
}
A0_HD void st_stream(double2 *p, double2 v) {
#ifdef __CUDA_ARCH__
  __stcs(p, v);
#else
  *p = v;
#endif
}
A0_HD double ld_tab(const double *p) {
#ifdef __CUDA_ARCH__
  return __ldg(p);
#else
  return *p;
#endif
}

template <int W>
A0_HD int swz(int n) {
  return W == 4 ? (n ^ ((n >> 3) & 1)) : n;
}

// Row index swz(nb + m*STRIDE) written so that, with m a compile-time constant, it is ONE of two
// per-item base values plus an immediate: the swizzle only toggles bit 0 with bit 3, and the
// strides that occur for W = 4 are multiples of 16 (bit 3 of nb decides), 8 (bit 3 alternates
// with m) and 1 with nb a multiple of 8 (bit 0 of m is toggled by bit 3 of nb).
template <int W, int STRIDE>
A0_HD int row_of(int nb, int m) {
  if constexpr (W != 4) {
    return nb + m * STRIDE;
  } else if constexpr (STRIDE % 16 == 0) {
    return (nb ^ ((nb >> 3) & 1)) + m * STRIDE;
  } else if constexpr (STRIDE == 8) {
    return (nb ^ (((nb >> 3) ^ m) & 1)) + m * STRIDE;
  } else {
    static_assert(STRIDE == 1, "unexpected stride for W = 4");
    const int b = (nb >> 3) & 1;  // nb is a multiple of 8
    return (m & 1) ? nb - b + m : nb + b + m;
  }
}

// 16-byte asynchronous copy global -> shared (LDGSTS): no register is held while the data is in
// flight.  Host replay: an ordinary copy.
A0_HD void cp_async16(double2 *smem_dst, const double2 *gsrc) {
#ifdef __CUDA_ARCH__
  const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
#else
  *smem_dst = *gsrc;
#endif
}
A0_HD void cp_async_commit() {
#ifdef __CUDA_ARCH__
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
// waits until at most `pending` of this thread's most recent copy groups are still in flight
A0_HD void cp_async_wait(int pending) {
#ifdef __CUDA_ARCH__
  switch (pending) {
    case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
  }
#else
  (void)pending;
#endif
}

// Offset of column j (row n0 = 0) inside a component.  Column counts fit 32 bits (checked by the
// launcher): one 32-bit division instead of a 64-bit one per thread and tile.
A0_HD long long col_base(const Params &p, long long j) {
  const unsigned ju = unsigned(j), bc = unsigned(p.blk_cols), b = ju / bc;
  return (long long)b * p.blk_stride + (long long)(ju - b * bc);
}

// Configuration: N0 = R0*R1*R2 (R2 = 1: two stages), W columns per tile.
template <int N0_, int W_, int R0_, int R1_, int R2_, int KEEP_ = 1>
struct Cfg {
  static constexpr int N0 = N0_, W = W_, R0 = R0_, R1 = R1_, R2 = R2_;
  // components whose RL values stay in REGISTERS from the last forward stage through K^ to the first
  // inverse stage (the others make three round trips through shared memory in the middle phase)
  static constexpr int KEEP = KEEP_;
  static constexpr int NS = R2_ == 1 ? 2 : 3;
  static constexpr int RL = NS == 3 ? R2_ : R1_;            // radix of the last (stride-1) stage
  static constexpr int NPH = 2 * NS - 1;                    // phases separated by CTA barriers
  static constexpr int THREADS = 256;
  static constexpr int MINB = (N0_ * W_ * 3 * 16 + N0_ * 16) <= 110 * 1024 ? 2 : 1;
  // Three-stage plans with W <= 4: the 32 items of a warp are the 8 butterflies x W columns of ONE
  // block of R1*R2 = 64 rows in the second stage, the middle phase and the second-to-last stage, so
  // those phases only need a warp barrier between them: two CTA barriers per tile instead of four.
  static constexpr bool WARP_LOCAL = R2_ == 8 && R1_ == 8 && W_ == 4;
  // first-stage items per thread; 1: the loads of a tile are waited for component by component
  static constexpr int ITEMS0 = (N0_ / R0_ * W_ + THREADS - 1) / THREADS;
  static constexpr size_t smem_bytes(int dim) { return size_t(N0_) * 16 + size_t(dim) * N0_ * W_ * 16; }
  // twiddle table (device array of N0 complex, staged in shared memory): stage 0 at [0], entry
  // (m-1)*stride0 + j = w_N0^(j m); stage 1 (three-stage plans) behind it, entry (m-1)*stride1 + j =
  // w_{N0/R0}^(j m).  Consecutive lanes (consecutive j) read consecutive words: no bank conflicts.
  static constexpr int TW1 = (R0_ - 1) * (N0_ / R0_);
  static_assert(R0_ * R1_ * R2_ == N0_, "radices must multiply to N0");
  static_assert(RL == 8, "the stride-1 stage must be radix 8 (bank swizzle)");
  static_assert(W_ % 4 == 0 && THREADS % W_ == 0, "W must divide the CTA size");
};

// ---- loads of a tile: every thread copies the inputs of ITS first-stage butterflies, all
// components, asynchronously into the shared-memory slots those butterflies work in.  The same
// thread owns the same slots in the last inverse stage, so the copies for the NEXT tile are issued
// there, slot by slot as soon as the current tile's value has been read (inv_stage<LAST>): the global
// latency hides behind that stage instead of being exposed after a barrier.
template <class C, int DIM>
A0_HD void issue_tile_loads(int tid, double2 *data, const Params &p, long long col0) {
  constexpr int N0 = C::N0, W = C::W, R = C::R0, stride = N0 / R, nbf = N0 / R;
  for (int item = tid; item < nbf * W; item += C::THREADS) {
    const int w = item % W, nb = item / W;  // first stage: one block, j = q
    if (col0 + w >= p.S) continue;
    const double2 *g = p.X + col_base(p, col0 + w) + (long long)nb * p.row_stride;
    double2 *d = data + w;
    for (int c = 0; c < DIM; c++, d += N0 * W, g += p.comp_stride) {
#pragma unroll
      for (int r = 0; r < R; r++) cp_async16(d + row_of<W, stride>(nb, r) * W, g + (long long)r * stride * p.row_stride);
      cp_async_commit();  // one group per component (see fwd_stage<FIRST>)
    }
  }
}

// ---- forward stage (not the last one): radix R on sub-blocks of size BS -----------------
// tws: this stage's twiddle table ([m-1][j], see Cfg).  FIRST: the inputs were copied into this
// thread's own slots by issue_tile_loads / the previous tile's last stage; wait for them (no CTA
// barrier is needed: nobody else touches these slots before the barrier that ends this stage).
template <class C, int DIM, int R, int BS, bool FIRST>
A0_HD void fwd_stage(int tid, double2 *data, const double2 *tws, const Params &p, long long col0) {
  constexpr int N0 = C::N0, W = C::W, stride = BS / R, nbf = N0 / R;
  // FIRST: the copies were committed one group per component, in component order; with one item per
  // thread component c is complete once at most DIM-1-c groups are pending, so the later components'
  // copies stay in flight while the first ones are transformed
  constexpr bool PER_COMP = FIRST && C::ITEMS0 == 1;
  if constexpr (FIRST && !PER_COMP) cp_async_wait(0);
  for (int item = tid; item < nbf * W; item += C::THREADS) {
    const int w = item % W, q = item / W;
    if (col0 + w >= p.S) continue;
    const int blk = q / stride, j = q % stride;
    const int nb = blk * BS + j;
    double2 t[R - 1];
#pragma unroll
    for (int m = 1; m < R; m++) t[m - 1] = tws[(m - 1) * stride + j];
    double2 *d = data + w;
    if constexpr (R <= 8) {
      // software pipeline over the components: the shared-memory loads of component c+1 are issued
      // before the butterflies of component c (short-scoreboard stalls were the top stall reason)
      double2 a[R], nx[R] = {};
      if constexpr (PER_COMP) cp_async_wait(DIM - 1);
#pragma unroll
      for (int r = 0; r < R; r++) a[r] = d[row_of<W, stride>(nb, r) * W];
#pragma unroll
      for (int c = 0; c < DIM; c++) {
        if (c + 1 < DIM) {
          if constexpr (PER_COMP) cp_async_wait(DIM - 2 - c);
#pragma unroll
          for (int r = 0; r < R; r++) nx[r] = d[((c + 1) * N0 + row_of<W, stride>(nb, r)) * W];
        }
        Dft<R, false>::run(a);
#pragma unroll
        for (int m = 1; m < R; m++) a[m] = cmul(a[m], t[m - 1].x, t[m - 1].y);
#pragma unroll
        for (int m = 0; m < R; m++) d[(c * N0 + row_of<W, stride>(nb, m)) * W] = a[m];
        if (c + 1 < DIM) {
#pragma unroll
          for (int r = 0; r < R; r++) a[r] = nx[r];
        }
      }
    } else {
#pragma unroll 1
      for (int c = 0; c < DIM; c++, d += N0 * W) {
        if constexpr (PER_COMP) cp_async_wait(DIM - 1 - c);
        double2 a[R];
#pragma unroll
        for (int r = 0; r < R; r++) a[r] = d[row_of<W, stride>(nb, r) * W];
        Dft<R, false>::run(a);
#pragma unroll
        for (int m = 1; m < R; m++) a[m] = cmul(a[m], t[m - 1].x, t[m - 1].y);
#pragma unroll
        for (int m = 0; m < R; m++) d[row_of<W, stride>(nb, m) * W] = a[m];
      }
    }
  }
}

// ---- inverse stage (not the first one): transposed forward stage, conjugated twiddles --------
// LAST: results go straight to global memory (natural order), and the freed slots are refilled
// with the next tile's inputs (next_col0; < 0 or >= S: nothing to prefetch).
template <class C, int DIM, int R, int BS, bool LAST>
A0_HD void inv_stage(int tid, double2 *data, const double2 *tws, const Params &p, long long col0,
                     long long next_col0 = -1) {
  constexpr int N0 = C::N0, W = C::W, stride = BS / R, nbf = N0 / R;
  for (int item = tid; item < nbf * W; item += C::THREADS) {
    const int w = item % W, q = item / W;
    const bool have = col0 + w < p.S;
    const bool more = LAST && next_col0 >= 0 && next_col0 + w < p.S;
    if (!have && !more) continue;
    const int blk = q / stride, j = q % stride;
    const int nb = blk * BS + j;
    double2 t[R - 1];
#pragma unroll
    for (int m = 1; m < R; m++) t[m - 1] = tws[(m - 1) * stride + j];
    double2 *g = p.X + (have ? col_base(p, col0 + w) : 0) + (long long)nb * p.row_stride;
    const double2 *gn = p.X + (more ? col_base(p, next_col0 + w) : 0) + (long long)nb * p.row_stride;
    double2 *d = data + w;
    if constexpr (R <= 8) {
      // software pipeline over the components (see fwd_stage); LAST: the slots of a component are
      // refilled with the next tile's inputs right after they have been read
      double2 a[R], nx[R] = {};
      auto fetch = [&](double2 (&v)[R], int c) {
#pragma unroll
        for (int m = 0; m < R; m++)
          v[m] = have ? d[(c * N0 + row_of<W, stride>(nb, m)) * W] : make_double2(0., 0.);
        if constexpr (LAST) {
          if (more) {
#pragma unroll
            for (int r = 0; r < R; r++)
              cp_async16(d + (c * N0 + row_of<W, stride>(nb, r)) * W,
                         gn + c * p.comp_stride + (long long)r * stride * p.row_stride);
          }
          cp_async_commit();
        }
      };
      fetch(a, 0);
#pragma unroll
      for (int c = 0; c < DIM; c++) {
        if (c + 1 < DIM) fetch(nx, c + 1);
        if (have) {
#pragma unroll
          for (int m = 1; m < R; m++) a[m] = cmul(a[m], t[m - 1].x, -t[m - 1].y);
          Dft<R, true>::run(a);
#pragma unroll
          for (int r = 0; r < R; r++) {
            if (LAST) st_stream(g + c * p.comp_stride + (long long)r * stride * p.row_stride, a[r]);
            else d[(c * N0 + row_of<W, stride>(nb, r)) * W] = a[r];
          }
        }
        if (c + 1 < DIM) {
#pragma unroll
          for (int r = 0; r < R; r++) a[r] = nx[r];
        }
      }
    } else {
#pragma unroll 1
      for (int c = 0; c < DIM; c++, d += N0 * W, g += p.comp_stride, gn += p.comp_stride) {
        double2 a[R];
#pragma unroll
        for (int m = 0; m < R; m++) a[m] = have ? d[row_of<W, stride>(nb, m) * W] : make_double2(0., 0.);
        if constexpr (LAST) {
          if (more) {
#pragma unroll
            for (int r = 0; r < R; r++)
              cp_async16(d + row_of<W, stride>(nb, r) * W, gn + (long long)r * stride * p.row_stride);
          }
          cp_async_commit();
        }
        if (!have) continue;
#pragma unroll
        for (int m = 1; m < R; m++) a[m] = cmul(a[m], t[m - 1].x, -t[m - 1].y);
        Dft<R, true>::run(a);
#pragma unroll
        for (int r = 0; r < R; r++) {
          if (LAST) st_stream(g + (long long)r * stride * p.row_stride, a[r]);
          else d[row_of<W, stride>(nb, r) * W] = a[r];
        }
      }
    }
  }
}

// K^ u of one mode.  The factors that do not depend on k0 are combined once per column
// (ColumnFactors, with |h|/|N| folded in); per mode that leaves 12 flops for the six distinct entries
// of K^ (bri17.hpp:266-288) and 18 fused multiply-adds for the product (tests/test_bri17.cpp:68 / :84).
// The products are re-associated with respect to the reference's written order: this path is
// compared at 1e-13, not bitwise (an FFT is involved anyway); the bit-exact kernels are in
// modal_kernels.cu.
template <int DIM>
struct ColumnFactors {
  double a, b, c, d, e, g;  // 3-D: chi1 chi2, phi1 chi2, chi1 phi2, s psi1 chi2, s chi1 psi2, s psi1 psi2 (x out_scale)
};

template <int DIM>
A0_HD ColumnFactors<DIM> column_factors(const double *phi, const double *chi, const double *psi, double scaling,
                                        double out_scale) {
  ColumnFactors<DIM> f;
  if constexpr (DIM == 3) {
    const double c2 = chi[2] * out_scale, sp1 = scaling * psi[1];
    f.a = chi[1] * c2;
    f.b = phi[1] * c2;
    f.c = chi[1] * (phi[2] * out_scale);
    f.d = sp1 * c2;
    f.e = (scaling * chi[1]) * (psi[2] * out_scale);
    f.g = sp1 * (psi[2] * out_scale);
  } else {
    f.a = chi[1] * out_scale;
    f.b = phi[1] * out_scale;
    f.d = (scaling * psi[1]) * out_scale;
    f.c = f.e = f.g = 0.;
  }
  return f;
}

template <int DIM>
A0_HD void stiffness_times(double phi0, double chi0, double psi0, const ColumnFactors<DIM> &cf, double mu,
                           double scaling, const double2 *u, double2 *f) {
  if constexpr (DIM == 3) {
    const double H00 = phi0 * cf.a, H11 = chi0 * cf.b, H22 = chi0 * cf.c;
    const double Kd = mu * ((H00 + H11) + H22);
    const double K00 = fma_(scaling, H00, Kd), K11 = fma_(scaling, H11, Kd), K22 = fma_(scaling, H22, Kd);
    const double K01 = psi0 * cf.d, K02 = psi0 * cf.e, K12 = chi0 * cf.g;
    f[0] = make_double2(fma_(K02, u[2].x, fma_(K01, u[1].x, K00 * u[0].x)), fma_(K02, u[2].y, fma_(K01, u[1].y, K00 * u[0].y)));
    f[1] = make_double2(fma_(K12, u[2].x, fma_(K11, u[1].x, K01 * u[0].x)), fma_(K12, u[2].y, fma_(K11, u[1].y, K01 * u[0].y)));
    f[2] = make_double2(fma_(K22, u[2].x, fma_(K12, u[1].x, K02 * u[0].x)), fma_(K22, u[2].y, fma_(K12, u[1].y, K02 * u[0].y)));
  } else {
    const double H00 = phi0 * cf.a, H11 = chi0 * cf.b;
    const double Kd = mu * (H00 + H11);
    const double K00 = fma_(scaling, H00, Kd), K11 = fma_(scaling, H11, Kd);
    const double K01 = psi0 * cf.d;
    f[0] = make_double2(fma_(K01, u[1].x, K00 * u[0].x), fma_(K01, u[1].y, K00 * u[0].y));
    f[1] = make_double2(fma_(K11, u[1].x, K01 * u[0].x), fma_(K11, u[1].y, K01 * u[0].y));
  }
}

// ---- middle phase: last forward stage (stride 1), K^ in digit-reversed position, first inverse
// stage.  The same thread owns the RL positions q*RL .. q*RL+RL-1 of its column in all three
// steps, so no CTA barrier is needed between them. --------------------------------------------
template <class C, int DIM>
A0_HD void mid_phase(int tid, double2 *data, const Params &p, long long col0, double &dot_acc) {
  constexpr int N0 = C::N0, W = C::W, R = C::RL, nbf = N0 / R;
  for (int item = tid; item < nbf * W; item += C::THREADS) {
    const int w = item % W, q = item / W;
    const long long col = col0 + w;
    if (col >= p.S) continue;
    const int nb = q * R;
    double2 *d = data + w;
    // last forward stage (no twiddle after it); the last KEEP components stay in registers
    constexpr int KEEP = C::KEEP < DIM ? C::KEEP : DIM, VIA = DIM - KEEP;
    double2 keep[KEEP][R];
#pragma unroll 1
    for (int c = 0; c < VIA; c++) {
      double2 a[R];
#pragma unroll
      for (int r = 0; r < R; r++) a[r] = d[(c * N0 + row_of<W, 1>(nb, r)) * W];
      Dft<R, false>::run(a);
#pragma unroll
      for (int r = 0; r < R; r++) d[(c * N0 + row_of<W, 1>(nb, r)) * W] = a[r];
    }
#pragma unroll
    for (int kc = 0; kc < KEEP; kc++) {
#pragma unroll
      for (int r = 0; r < R; r++) keep[kc][r] = d[((VIA + kc) * N0 + row_of<W, 1>(nb, r)) * W];
      Dft<R, false>::run(keep[kc]);
    }
    // position nb + r holds frequency k0 = kbase + r * (N0 / R): digits of q, least significant
    // stage first (decimation in frequency leaves the spectrum in digit-reversed order)
    int kbase;
    if constexpr (C::NS == 3) kbase = q / C::R1 + C::R0 * (q % C::R1);
    else kbase = q;
    double phi[3], chi[3], psi[3];
    int k_last;
    const unsigned colu = unsigned(col);  // column counts fit 32 bits (launcher)
    if constexpr (DIM == 3) {
      const unsigned b = colu / unsigned(p.S2e);
      const int k2 = int(colu - b * unsigned(p.S2e)), k1 = p.k1_begin + int(b);
      phi[1] = ld_tab(p.tab1 + k1); chi[1] = ld_tab(p.tab1 + p.N1 + k1); psi[1] = ld_tab(p.tab1 + 2 * p.N1 + k1);
      phi[2] = ld_tab(p.tab2 + k2); chi[2] = ld_tab(p.tab2 + p.N2 + k2); psi[2] = ld_tab(p.tab2 + 2 * p.N2 + k2);
      k_last = k2;
    } else {
      const int k1 = p.k1_begin + int(colu);
      phi[1] = ld_tab(p.tab1 + k1); chi[1] = ld_tab(p.tab1 + p.N1 + k1); psi[1] = ld_tab(p.tab1 + 2 * p.N1 + k1);
      phi[2] = chi[2] = psi[2] = 0.;
      k_last = k1;
    }
    const ColumnFactors<DIM> cf = column_factors<DIM>(phi, chi, psi, p.scaling, p.out_scale);
    const double wgt = (p.herm_n > 0 && k_last != 0 && 2 * k_last != p.herm_n) ? 2. : 1.;
#pragma unroll
    for (int r = 0; r < R; r++) {
      const int k0 = kbase + r * nbf;
      const double phi0 = ld_tab(p.tab0 + k0), chi0 = ld_tab(p.tab0 + N0 + k0), psi0 = ld_tab(p.tab0 + 2 * N0 + k0);
      double2 u[DIM], f[DIM];
#pragma unroll
      for (int c = 0; c < VIA; c++) u[c] = d[(c * N0 + row_of<W, 1>(nb, r)) * W];
#pragma unroll
      for (int kc = 0; kc < KEEP; kc++) u[VIA + kc] = keep[kc][r];
      stiffness_times<DIM>(phi0, chi0, psi0, cf, p.mu, p.scaling, u, f);
#pragma unroll
      for (int c = 0; c < VIA; c++) d[(c * N0 + row_of<W, 1>(nb, r)) * W] = f[c];
#pragma unroll
      for (int kc = 0; kc < KEEP; kc++) keep[kc][r] = f[VIA + kc];
      if (p.dot_partial) {
        double dd = u[0].x * f[0].x;
        dd = fma_(u[0].y, f[0].y, dd);
#pragma unroll
        for (int c = 1; c < DIM; c++) { dd = fma_(u[c].x, f[c].x, dd); dd = fma_(u[c].y, f[c].y, dd); }
        dot_acc = fma_(wgt, dd, dot_acc);
      }
    }
    // first inverse stage (stride 1, no twiddle before it): the register-resident components first
#pragma unroll
    for (int kc = 0; kc < KEEP; kc++) {
      Dft<R, true>::run(keep[kc]);
#pragma unroll
      for (int r = 0; r < R; r++) d[((VIA + kc) * N0 + row_of<W, 1>(nb, r)) * W] = keep[kc][r];
    }
#pragma unroll 1
    for (int c = 0; c < VIA; c++) {
      double2 a[R];
#pragma unroll
      for (int r = 0; r < R; r++) a[r] = d[(c * N0 + row_of<W, 1>(nb, r)) * W];
      Dft<R, true>::run(a);
#pragma unroll
      for (int r = 0; r < R; r++) d[(c * N0 + row_of<W, 1>(nb, r)) * W] = a[r];
    }
  }
}

// Phase PH of a tile for thread `tid`.  A CTA barrier separates consecutive phases of a tile; none is
// needed between the last phase of a tile and the first phase of the next (same slot ownership).
// next_col0: first column of the tile this CTA processes next (its inputs are prefetched by the last
// phase), -1 if none.
// (An L2 prefetch instruction per row was tried instead and removed: it raised the DRAM reads of a
// 512^3 pass from 6.4 to 11.3 GB without shortening the load phase -- profiles/r02_axis0_fused.md.)
template <class C, int DIM, int PH>
A0_HD void phase(int tid, double2 *data, const double2 *tw, const Params &p, long long col0, long long next_col0,
                 double &dot_acc) {
  if constexpr (C::NS == 3) {
    if constexpr (PH == 0) fwd_stage<C, DIM, C::R0, C::N0, true>(tid, data, tw, p, col0);
    else if constexpr (PH == 1) fwd_stage<C, DIM, C::R1, C::N0 / C::R0, false>(tid, data, tw + C::TW1, p, col0);
    else if constexpr (PH == 2) mid_phase<C, DIM>(tid, data, p, col0, dot_acc);
    else if constexpr (PH == 3) inv_stage<C, DIM, C::R1, C::N0 / C::R0, false>(tid, data, tw + C::TW1, p, col0);
    else inv_stage<C, DIM, C::R0, C::N0, true>(tid, data, tw, p, col0, next_col0);
  } else {
    if constexpr (PH == 0) fwd_stage<C, DIM, C::R0, C::N0, true>(tid, data, tw, p, col0);
    else if constexpr (PH == 1) mid_phase<C, DIM>(tid, data, p, col0, dot_acc);
    else inv_stage<C, DIM, C::R0, C::N0, true>(tid, data, tw, p, col0, next_col0);
  }
}

// CPU replay of the kernel, thread by thread (tests only; a barrier = the end of a tid loop).
template <class C, int DIM>
void emulate_host(const Params &p, double *dot_out) {
  double2 *data = new double2[size_t(DIM) * C::N0 * C::W];
  double *acc = new double[C::THREADS]();
  // same schedule as the kernel with a grid of 3 CTAs replayed one after the other, so that the
  // "next tile" prefetch and the missing barrier between tiles are exercised
  const long long G = 3;
  for (long long cta = 0; cta < G; cta++) {
    if (cta < p.n_tiles)
      for (int t = 0; t < C::THREADS; t++) issue_tile_loads<C, DIM>(t, data, p, cta * C::W);
    for (long long tile = cta; tile < p.n_tiles; tile += G) {
      const long long col0 = tile * C::W, nx = tile + G < p.n_tiles ? (tile + G) * C::W : -1;
      // no barrier between the last phase of a tile and phase 0 of the next: replay them fused per
      // thread where the kernel would run them back to back -- here phase 0 follows in program order
      for (int t = 0; t < C::THREADS; t++) phase<C, DIM, 0>(t, data, p.twiddle, p, col0, nx, acc[t]);
      for (int t = 0; t < C::THREADS; t++) phase<C, DIM, 1>(t, data, p.twiddle, p, col0, nx, acc[t]);
      for (int t = 0; t < C::THREADS; t++) phase<C, DIM, 2>(t, data, p.twiddle, p, col0, nx, acc[t]);
      if constexpr (C::NPH > 3) {
        for (int t = 0; t < C::THREADS; t++) phase<C, DIM, 3>(t, data, p.twiddle, p, col0, nx, acc[t]);
        for (int t = 0; t < C::THREADS; t++) phase<C, DIM, 4>(t, data, p.twiddle, p, col0, nx, acc[t]);
      }
    }
  }
  if (dot_out) {
    double s = 0.;
    for (int t = 0; t < C::THREADS; t++) s += acc[t];
    *dot_out = s;
  }
  delete[] data;
  delete[] acc;
}

// Supported axis-0 lengths and their factorizations.
using Cfg16 = Cfg<16, 128, 2, 8, 1>;
using Cfg32 = Cfg<32, 64, 4, 8, 1>;
using Cfg64 = Cfg<64, 32, 8, 8, 1>;
using Cfg128 = Cfg<128, 16, 2, 8, 8>;
using Cfg256 = Cfg<256, 8, 4, 8, 8>;
using Cfg512 = Cfg<512, 4, 8, 8, 8, 2>;
using Cfg1024 = Cfg<1024, 4, 16, 8, 8, 3>;

inline bool supported(int N0) {
  return N0 == 16 || N0 == 32 || N0 == 64 || N0 == 128 || N0 == 256 || N0 == 512 || N0 == 1024;
}

}  // namespace axis0
}  // namespace bri17b200
