// mode_math.h -- per-frequency arithmetic shared by host (plan.cu) and device
// (modal_kernels.cu) code: K^ and B^ from the per-axis factors, the 2x2/3x3
// Cholesky solve and the eigenstress -> displacement / strain maps of
// Hooke::modal_eigenstress_to_opposite_strain (bri17.hpp:308-355).
//
// Every operation is written out one rounding at a time, in the order of the
// reference's expressions; the translation units that include this header are
// built with -fmad=false / -ffp-contract=off, so host and device agree bit for
// bit on everything except the transcendental prefactor of B^.
#pragma once

#ifdef __CUDACC__
#define BRI17_HD __host__ __device__ __forceinline__
#else
#define BRI17_HD inline
#endif

namespace bri17b200 {

struct Cplx {
  double re, im;
};
BRI17_HD Cplx cadd(Cplx a, Cplx b) { return {a.re + b.re, a.im + b.im}; }
BRI17_HD Cplx csub(Cplx a, Cplx b) { return {a.re - b.re, a.im - b.im}; }
BRI17_HD Cplx cmulc(Cplx a, Cplx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
BRI17_HD Cplx cscale(double s, Cplx a) { return {s * a.re, s * a.im}; }
BRI17_HD Cplx cdivr(Cplx a, double d) { return {a.re / d, a.im / d}; }
BRI17_HD Cplx cconj(Cplx a) { return {a.re, -a.im}; }

// Distinct entries of the real symmetric K^ (bri17.hpp:266-288).  phi/chi/psi: per-axis factors.
template <int DIM>
BRI17_HD void stiffness_entries(const double *phi, const double *chi, const double *psi, double mu,
                                double scaling, double (&K)[DIM][DIM]) {
  if constexpr (DIM == 2) {
    const double H00 = phi[0] * chi[1];
    const double H11 = chi[0] * phi[1];
    const double Kd = mu * (H00 + H11);
    K[0][0] = scaling * H00 + Kd;
    K[0][1] = K[1][0] = scaling * psi[0] * psi[1];
    K[1][1] = scaling * H11 + Kd;
  } else {
    const double H00 = phi[0] * chi[1] * chi[2];
    const double H11 = chi[0] * phi[1] * chi[2];
    const double H22 = chi[0] * chi[1] * phi[2];
    const double Kd = mu * (H00 + H11 + H22);
    K[0][0] = scaling * H00 + Kd;
    K[1][1] = scaling * H11 + Kd;
    K[2][2] = scaling * H22 + Kd;
    K[0][1] = K[1][0] = scaling * psi[0] * psi[1] * chi[2];
    K[0][2] = K[2][0] = scaling * psi[0] * chi[1] * psi[2];
    K[1][2] = K[2][1] = scaling * chi[0] * psi[1] * psi[2];
  }
}

// B^ from the half-angle factors and the prefactor (-2 sin S, 2 cos S) (bri17.hpp:224-232).
template <int DIM>
BRI17_HD void strain_displacement_entries(const double *c, const double *s, Cplx pre, Cplx (&B)[DIM]) {
  for (int i = 0; i < DIM; i++) {
    Cplx b = pre;
    for (int d = 0; d < DIM; d++) b = cscale(d == i ? s[d] : c[d], b);
    B[i] = b;
  }
}

// x <- K^-1 x for a real SPD K (destroyed), complex x: Cholesky K = L L^T,
// forward then backward substitution.  Stands in for Eigen's K.llt().solve(rhs)
// (bri17.hpp:341) in the operation order Eigen 3.3/3.4 uses for fixed sizes:
// llt_inplace<Lower>::unblocked (x = A_kk - A10.squaredNorm(); A21 -= A20 * A10^H;
// A21 /= x) and the unrolled triangular solves (rhs_i -= (row_i . rhs).sum();
// rhs_i /= L_ii): sums of products are formed first and subtracted once.  No
// reference test pins the result (parity unpinned).
template <int DIM>
BRI17_HD void cholesky_solve(double (&A)[DIM][DIM], Cplx (&x)[DIM]) {
#ifdef __CUDA_ARCH__
  // Device: fp64 divide and sqrt are ~30-instruction software sequences and would make
  // the batched solve FP64-bound.  Same factorisation and grouping with one rsqrt per
  // pivot and multiplications by the reciprocal pivots (differs from the host path by a
  // few ulp; tested at 1e-12).
  double r[DIM];  // 1 / L_jj
  for (int j = 0; j < DIM; j++) {
    double d = A[j][j];
    if (j > 0) {
      double sq = A[j][0] * A[j][0];
      for (int p = 1; p < j; p++) sq = sq + A[j][p] * A[j][p];
      d = d - sq;
    }
    r[j] = rsqrt(d);
    for (int i = j + 1; i < DIM; i++) {
      double t = A[i][j];
      if (j > 0) {
        double dot = A[i][0] * A[j][0];
        for (int p = 1; p < j; p++) dot = dot + A[i][p] * A[j][p];
        t = t - dot;
      }
      A[i][j] = t * r[j];
    }
  }
  for (int i = 0; i < DIM; i++) {
    Cplx t = x[i];
    if (i > 0) {
      Cplx acc = cscale(A[i][0], x[0]);
      for (int p = 1; p < i; p++) acc = cadd(acc, cscale(A[i][p], x[p]));
      t = csub(t, acc);
    }
    x[i] = cscale(r[i], t);
  }
  for (int i = DIM - 1; i >= 0; i--) {
    Cplx t = x[i];
    if (i < DIM - 1) {
      Cplx acc = cscale(A[i + 1][i], x[i + 1]);
      for (int p = i + 2; p < DIM; p++) acc = cadd(acc, cscale(A[p][i], x[p]));
      t = csub(t, acc);
    }
    x[i] = cscale(r[i], t);
  }
#else
  for (int j = 0; j < DIM; j++) {
    double d = A[j][j];
    if (j > 0) {
      double sq = A[j][0] * A[j][0];
      for (int p = 1; p < j; p++) sq = sq + A[j][p] * A[j][p];
      d = d - sq;
    }
    d = sqrt(d);
    A[j][j] = d;
    for (int i = j + 1; i < DIM; i++) {
      double t = A[i][j];
      if (j > 0) {
        double dot = A[i][0] * A[j][0];
        for (int p = 1; p < j; p++) dot = dot + A[i][p] * A[j][p];
        t = t - dot;
      }
      A[i][j] = t / d;
    }
  }
  for (int i = 0; i < DIM; i++) {
    Cplx t = x[i];
    if (i > 0) {
      Cplx acc = cscale(A[i][0], x[0]);
      for (int p = 1; p < i; p++) acc = cadd(acc, cscale(A[i][p], x[p]));
      t = csub(t, acc);
    }
    x[i] = cdivr(t, A[i][i]);
  }
  for (int i = DIM - 1; i >= 0; i--) {
    Cplx t = x[i];
    if (i < DIM - 1) {
      Cplx acc = cscale(A[i + 1][i], x[i + 1]);
      for (int p = i + 2; p < DIM; p++) acc = cadd(acc, cscale(A[p][i], x[p]));
      t = csub(t, acc);
    }
    x[i] = cdivr(t, A[i][i]);
  }
#endif
}

template <int DIM>
struct Mandel {
  static constexpr int nsym = DIM * (DIM + 1) / 2;
};
// (row, column) of Mandel component s: 2-D [00, 11, 01]; 3-D [00, 11, 22, 12, 20, 01]
// (bri17.hpp:324-332, :344-353; tests/test_bri17.cpp:207-209, :224-230)
template <int DIM>
BRI17_HD void mandel_pair(int s, int &p, int &q) {
  if constexpr (DIM == 2) {
    p = s == 1 ? 1 : 0;
    q = s == 0 ? 0 : 1;
  } else {
    const int P[6] = {0, 1, 2, 1, 2, 0}, Q[6] = {0, 1, 2, 2, 0, 1};
    p = P[s];
    q = Q[s];
  }
}

// f = tau . conj(B)  (bri17.hpp:324-332, :340): the right-hand side of the per-mode solve,
// i.e. the modal force equivalent to the eigenstress tau (Mandel notation).
template <int DIM>
BRI17_HD void eigenstress_to_force(const Cplx *tau, const Cplx (&B)[DIM], Cplx (&f)[DIM]) {
  Cplx t[DIM][DIM];
  for (int s = 0; s < Mandel<DIM>::nsym; s++) {
    int p, q;
    mandel_pair<DIM>(s, p, q);
    if (p == q) t[p][p] = tau[s];
    else {
#ifdef __CUDA_ARCH__
      t[p][q] = t[q][p] = cscale(0.7071067811865476, tau[s]);  // 1/sqrt2: avoids two fp64 divides per entry
#else
      t[p][q] = t[q][p] = cdivr(tau[s], 1.4142135623730951);  // std::numbers::sqrt2_v<double>
#endif
    }
  }
  for (int i = 0; i < DIM; i++) {
    Cplx acc = cmulc(t[i][0], cconj(B[0]));
    for (int j = 1; j < DIM; j++) acc = cadd(acc, cmulc(t[i][j], cconj(B[j])));
    f[i] = acc;
  }
}

// u = K^-1 (tau . conj(B))  (bri17.hpp:324-341).  tau in Mandel notation.  K is destroyed.
template <int DIM>
BRI17_HD void eigenstress_to_displacement(const Cplx *tau, const Cplx (&B)[DIM], double (&K)[DIM][DIM],
                                          Cplx (&u)[DIM]) {
  eigenstress_to_force<DIM>(tau, B, u);
  cholesky_solve<DIM>(K, u);
}

// eta = sym(B (x) u) in Mandel notation (bri17.hpp:342-353).
template <int DIM>
BRI17_HD void displacement_to_strain(const Cplx (&B)[DIM], const Cplx (&u)[DIM], Cplx *eta) {
  const double sqrt2 = 1.4142135623730951;
  for (int s = 0; s < Mandel<DIM>::nsym; s++) {
    int p, q;
    mandel_pair<DIM>(s, p, q);
    Cplx e = cscale(0.5, cadd(cmulc(B[p], u[q]), cmulc(u[p], B[q])));
    if (p != q) e = cscale(sqrt2, e);
    eta[s] = e;
  }
}

}  // namespace bri17b200
