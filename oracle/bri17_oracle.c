/*
 * bri17_oracle.c -- CPU restatement of the bri17 modal operator path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA
 * path in bri17_b200/csrc.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product
 * (libbri17_b200.so, include/bri17/bri17.hpp) never links or calls it.
 *
 * Every function restates, in plain C and with the SAME floating-point
 * operation order, a piece of the reference, cited as file:line relative to
 * /root/reference.  Build with -O2 -ffp-contract=off (oracle/Makefile) so
 * that no multiply-add is contracted: the results are then bit-identical to
 * the reference header compiled for baseline x86-64 (no FMA), which is what
 * tests/test_oracle_vs_ref.py checks against oracle/_ref/libbri17_ref.so.
 *
 * Parity status: PINNED for the hot path (UNPINNED only for
 * oracle_modal_eigenstress_to_opposite_strain, see its comment).  (1) bit-for-bit against the unmodified reference
 * header compiled in oracle/_ref (modal_stiffness, modal_strain_displacement,
 * whole-grid apply); (2) against the reference's own known-answer tests --
 * the Maxima-derived element matrices of tests/test_bri17.cpp:344-356,
 * :372-531, :547-552, :570-598 -- through the FFT sandwich restated with
 * numpy.fft in oracle/kat.py, at the reference tolerance 1e-15*|e|+1e-14.
 *
 * Third-party arithmetic that the reference delegates and that is NOT under
 * /root/reference:
 *   - libm sin/cos (glibc; bri17.hpp:220-221,224,261-263): called here
 *     through the same libm.
 *   - Eigen (>=3.3, unpinned, tests/CMakeLists.txt:12) fixed-size product
 *     K_k * u_k (tests/test_bri17.cpp:68,84): restated as a real-matrix
 *     times complex-vector product accumulated left to right.  Im(K)=0
 *     (bri17.hpp:271-288 assign real values to complex<T>), so the complex
 *     products reduce to real*real; Eigen's exact summation order for the
 *     three terms is not pinned by any reference test (difference <= 1 ulp,
 *     inside the 1e-12 per-mode gate).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* std::numbers::pi_v<double> (bri17.hpp:218,260) */
#define ORACLE_PI 3.141592653589793238462643383279502884

/*
 * Hooke<double,DIM>::modal_stiffness -- bri17.hpp:247-292.
 * K is DIM*DIM complex numbers, interleaved (re,im), row-major K[DIM*i+j].
 */
void oracle_modal_stiffness(int dim, const int *shape, const double *L,
                            double mu, double nu, const int *k, double *K) {
  double phi[3], psi[3], chi[3];
  for (int i = 0; i < dim; i++) {
    double h = L[i] / shape[i];                      /* :259 */
    double beta = 2 * ORACLE_PI * k[i] / shape[i];   /* :260 ((2*pi)*k)/N */
    phi[i] = 2 * (1 - cos(beta)) / h / h;            /* :261 */
    chi[i] = (2 + cos(beta)) / 3;                    /* :262 */
    psi[i] = sin(beta) / h;                          /* :263 */
  }
  const double scaling = mu / (1. - 2. * nu);        /* :266 */
  if (dim == 2) {
    double H_00 = phi[0] * chi[1];                   /* :268 */
    double H_11 = chi[0] * phi[1];                   /* :269 */
    double K_diag = mu * (H_00 + H_11);              /* :270 */
    double K00 = scaling * H_00 + K_diag;            /* :271 */
    double K01 = scaling * psi[0] * psi[1];          /* :272 */
    double K11 = scaling * H_11 + K_diag;            /* :274 */
    K[0] = K00; K[1] = 0.;
    K[2] = K01; K[3] = 0.;
    K[4] = K01; K[5] = 0.;                           /* :273 */
    K[6] = K11; K[7] = 0.;
  } else {
    double H_00 = phi[0] * chi[1] * chi[2];          /* :276 */
    double H_11 = chi[0] * phi[1] * chi[2];          /* :277 */
    double H_22 = chi[0] * chi[1] * phi[2];          /* :278 */
    double K_diag = mu * (H_00 + H_11 + H_22);       /* :279 */
    double K00 = scaling * H_00 + K_diag;            /* :280 */
    double K01 = scaling * psi[0] * psi[1] * chi[2]; /* :281 */
    double K02 = scaling * psi[0] * chi[1] * psi[2]; /* :282 */
    double K11 = scaling * H_11 + K_diag;            /* :284 */
    double K12 = scaling * chi[0] * psi[1] * psi[2]; /* :285 */
    double K22 = scaling * H_22 + K_diag;            /* :288 */
    double Kr[9] = {K00, K01, K02, K01, K11, K12, K02, K12, K22};
    for (int i = 0; i < 9; i++) { K[2 * i] = Kr[i]; K[2 * i + 1] = 0.; }
  }
}

/*
 * Hooke<double,DIM>::modal_strain_displacement -- bri17.hpp:212-236.
 * B is DIM complex numbers, interleaved.  std::complex<T> * T multiplies both
 * parts by the scalar, left to right as written at :227-232.
 */
void oracle_modal_strain_displacement(int dim, const int *shape,
                                      const double *L, const int *k,
                                      double *B) {
  double c[3], s[3];
  double sum_alpha = 0.;                             /* :215 */
  for (int i = 0; i < dim; i++) {
    double alpha = ORACLE_PI * k[i] / shape[i];      /* :218 (pi*k)/N */
    sum_alpha += alpha;                              /* :219 */
    c[i] = cos(alpha);                               /* :220 */
    s[i] = sin(alpha) * shape[i] / L[i];             /* :221 (sin*N)/L */
  }
  double pre_re = -2 * sin(sum_alpha);               /* :224 */
  double pre_im = 2 * cos(sum_alpha);
  if (dim == 2) {
    B[0] = pre_re * s[0] * c[1]; B[1] = pre_im * s[0] * c[1]; /* :227 */
    B[2] = pre_re * c[0] * s[1]; B[3] = pre_im * c[0] * s[1]; /* :228 */
  } else {
    B[0] = pre_re * s[0] * c[1] * c[2]; B[1] = pre_im * s[0] * c[1] * c[2];
    B[2] = pre_re * c[0] * s[1] * c[2]; B[3] = pre_im * c[0] * s[1] * c[2];
    B[4] = pre_re * c[0] * c[1] * s[2]; B[5] = pre_im * c[0] * c[1] * s[2];
  }
}

/* CartesianGrid::get_node_at -- bri17.hpp:78-81, :90-93 (row-major). */
int oracle_get_node_at(int dim, const int *shape, const int *ijk) {
  if (dim == 2) return ijk[0] * shape[1] + ijk[1];
  return (ijk[0] * shape[1] + ijk[1]) * shape[2] + ijk[2];
}

/* CartesianGrid::get_cell_nodes -- bri17.hpp:127-158 (periodic wrap, last
 * axis fastest within the cell). */
void oracle_get_cell_nodes(int dim, const int *shape, int cell, int *nodes) {
  if (dim == 2) {
    int i1 = cell / shape[1], j1 = cell % shape[1];
    int i2 = i1 == shape[0] - 1 ? 0 : i1 + 1;
    int j2 = j1 == shape[1] - 1 ? 0 : j1 + 1;
    int a[4][2] = {{i1, j1}, {i1, j2}, {i2, j1}, {i2, j2}};
    for (int n = 0; n < 4; n++) nodes[n] = oracle_get_node_at(2, shape, a[n]);
  } else {
    int k1 = cell % shape[2], ij1 = cell / shape[2];
    int j1 = ij1 % shape[1], i1 = ij1 / shape[1];
    int i2 = i1 == shape[0] - 1 ? 0 : i1 + 1;
    int j2 = j1 == shape[1] - 1 ? 0 : j1 + 1;
    int k2 = k1 == shape[2] - 1 ? 0 : k1 + 1;
    int a[8][3] = {{i1, j1, k1}, {i1, j1, k2}, {i1, j2, k1}, {i1, j2, k2},
                   {i2, j1, k1}, {i2, j1, k2}, {i2, j2, k1}, {i2, j2, k2}};
    for (int n = 0; n < 8; n++) nodes[n] = oracle_get_node_at(3, shape, a[n]);
  }
}

/*
 * Frequency index map -- the loop nest + running counter of
 * tests/test_bri17.cpp:58,62-64,71 (2-D) and :76-79,88 (3-D): linear element
 * i of a row-major block corresponds to k = k_begin + unravel(i, local).
 * Writes dim ints per element.  No fftshift, no negative wrap.
 */
void oracle_freq_index_map(int dim, const int *k_begin, const int *local_shape,
                           int32_t *k_out) {
  int64_t i = 0;
  int n2 = dim == 3 ? local_shape[2] : 1;
  for (int a = 0; a < local_shape[0]; a++)
    for (int b = 0; b < local_shape[1]; b++)
      for (int c = 0; c < n2; c++) {
        k_out[dim * i + 0] = k_begin[0] + a;
        k_out[dim * i + 1] = k_begin[1] + b;
        if (dim == 3) k_out[dim * i + 2] = k_begin[2] + c;
        i++;
      }
}

/*
 * Block-diagonal apply f^[c,k] = sum_j K^[k][c,j] u^[j,k] over a row-major
 * block of frequencies -- tests/test_bri17.cpp:58-92.  Planar layout: element
 * (c, i) at buf[i + c*comp_stride] (:66-67, :81-83), interleaved complex.
 * The block covers k = k_begin + [0, local_shape) (the reference always runs
 * the full grid: k_begin = 0, local_shape = shape, comp_stride = grid.size).
 * One libm-backed modal_stiffness call per mode, exactly like the reference.
 * OpenMP over the slowest index only when compiled with -fopenmp and
 * nthreads > 1 (the reference itself is single-threaded).
 */
void oracle_apply_modal_stiffness(int dim, const int *shape, const double *L,
                                  double mu, double nu, const int *k_begin,
                                  const int *local_shape, int64_t comp_stride,
                                  const double *u_hat, double *f_hat,
                                  int nthreads) {
  const int n0 = local_shape[0], n1 = local_shape[1];
  const int n2 = dim == 3 ? local_shape[2] : 1;
  (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
#endif
  for (int a = 0; a < n0; a++) {
    int k[3];
    double K[18];
    k[0] = k_begin[0] + a;
    for (int b = 0; b < n1; b++) {
      k[1] = k_begin[1] + b;
      for (int c = 0; c < n2; c++) {
        if (dim == 3) k[2] = k_begin[2] + c;
        int64_t i = ((int64_t)a * n1 + b) * n2 + c;      /* :71, :88 */
        oracle_modal_stiffness(dim, shape, L, mu, nu, k, K); /* :65, :80 */
        double ur[3], ui[3];
        for (int j = 0; j < dim; j++) {                  /* gather :66-67 */
          ur[j] = u_hat[2 * (i + j * comp_stride)];
          ui[j] = u_hat[2 * (i + j * comp_stride) + 1];
        }
        for (int r = 0; r < dim; r++) {                  /* matvec :68, :84 */
          double fr = K[2 * (dim * r)] * ur[0];
          double fi = K[2 * (dim * r)] * ui[0];
          for (int j = 1; j < dim; j++) {
            fr = fr + K[2 * (dim * r + j)] * ur[j];
            fi = fi + K[2 * (dim * r + j)] * ui[j];
          }
          f_hat[2 * (i + r * comp_stride)] = fr;         /* scatter :69-70 */
          f_hat[2 * (i + r * comp_stride) + 1] = fi;
        }
      }
    }
  }
}

/* complex product (a+ib)(c+id), one rounding per product, one per sum */
static inline void cmul(double ar, double ai, double br, double bi,
                        double *re, double *im) {
  *re = ar * br - ai * bi;
  *im = ar * bi + ai * br;
}

/*
 * Batched strain recovery eps^ = 1/2 (B^ (x) u^ + u^ (x) B^) in Mandel order
 * -- tests/test_bri17.cpp:194-235 (compute_Bu middle).  Input planar u^
 * (dim components, stride u_stride), output planar eps^ (3 or 6 components,
 * stride e_stride): 2-D [00, 11, sqrt2*01] (:207-209); 3-D [00, 11, 22,
 * sqrt2*12, sqrt2*20, sqrt2*01] (:224-230).
 */
void oracle_apply_strain_displacement(int dim, const int *shape,
                                      const double *L, const int *k_begin,
                                      const int *local_shape, int64_t u_stride,
                                      int64_t e_stride, const double *u_hat,
                                      double *eps_hat) {
  const int n0 = local_shape[0], n1 = local_shape[1];
  const int n2 = dim == 3 ? local_shape[2] : 1;
  const double sqrt2 = sqrt(2);                          /* :209, :228 */
  /* Mandel pairs, in output order */
  const int p2[3][2] = {{0, 0}, {1, 1}, {0, 1}};
  const int p3[6][2] = {{0, 0}, {1, 1}, {2, 2}, {1, 2}, {2, 0}, {0, 1}};
  const int nsym = dim == 2 ? 3 : 6;
  for (int a = 0; a < n0; a++)
    for (int b = 0; b < n1; b++)
      for (int c = 0; c < n2; c++) {
        int k[3] = {k_begin[0] + a, k_begin[1] + b, dim == 3 ? k_begin[2] + c : 0};
        int64_t i = ((int64_t)a * n1 + b) * n2 + c;
        double B[6], u[6];
        oracle_modal_strain_displacement(dim, shape, L, k, B); /* :203,:219 */
        for (int j = 0; j < dim; j++) {
          u[2 * j] = u_hat[2 * (i + j * u_stride)];
          u[2 * j + 1] = u_hat[2 * (i + j * u_stride) + 1];
        }
        for (int s = 0; s < nsym; s++) {
          int p = dim == 2 ? p2[s][0] : p3[s][0];
          int q = dim == 2 ? p2[s][1] : p3[s][1];
          double t1r, t1i, t2r, t2i;
          /* eps(p,q) = 0.5 * (B_p u_q + u_p B_q)  (:206, :223) */
          cmul(B[2 * p], B[2 * p + 1], u[2 * q], u[2 * q + 1], &t1r, &t1i);
          cmul(u[2 * p], u[2 * p + 1], B[2 * q], B[2 * q + 1], &t2r, &t2i);
          double er = 0.5 * (t1r + t2r), ei = 0.5 * (t1i + t2i);
          if (p != q) { er = sqrt2 * er; ei = sqrt2 * ei; }
          eps_hat[2 * (i + s * e_stride)] = er;
          eps_hat[2 * (i + s * e_stride) + 1] = ei;
        }
      }
}

/*
 * Hooke<double,DIM>::modal_eigenstress_to_opposite_strain -- bri17.hpp:308-355.
 * tau, eta: nsym complex numbers (Mandel notation), interleaved.
 *
 * PARITY UNPINNED: the reference solves K u = rhs with Eigen's K.llt().solve()
 * (:341); Eigen (>= 3.3, no version pin) is neither vendored nor installed and no
 * reference test calls this method (only python/demo.py does).  The solve is
 * restated in the operation order Eigen 3.3/3.4 publishes for fixed sizes < 32:
 * Cholesky/LLT.h llt_inplace<Scalar,Lower>::unblocked (x = A_kk - A10.squaredNorm();
 * A21 -= A20 * A10^H; A21 /= x) followed by matrixL().solveInPlace and
 * matrixU().solveInPlace through SolveTriangular.h triangular_solver_unroller
 * (rhs_i -= (row_i . rhs).sum(); rhs_i /= L_ii).  K^ is real (Im K = 0), so the
 * factor is real; a complex number divided by (L_ii + 0i) under GCC's __divdc3
 * equals the component-wise real division used here.  What stays unpinned: the
 * Eigen version and the user's compiler flags (FMA contraction, vectorised pmadd).
 * out_u (may be NULL) receives the intermediate displacement u (DIM complex).
 */
void oracle_modal_eigenstress_to_opposite_strain(int dim, const int *shape, const double *L,
                                                 double mu, double nu, const int *k,
                                                 const double *tau, double *eta, double *out_u) {
  const int nsym = dim == 2 ? 3 : 6;                       /* :315 */
  const double sqrt2 = 1.4142135623730951;                 /* :313 */
  double Bf[6], Kf[18];
  oracle_modal_strain_displacement(dim, shape, L, k, Bf);  /* :317 */
  oracle_modal_stiffness(dim, shape, L, mu, nu, k, Kf);    /* :319 */
  int null_frequency = 1;                                  /* :327, :334 */
  for (int d = 0; d < dim; d++) null_frequency = null_frequency && (k[d] == 0);
  if (null_frequency) {                                    /* :336-339 */
    for (int i = 0; i < 2 * nsym; i++) eta[i] = 0.;
    if (out_u) for (int i = 0; i < 2 * dim; i++) out_u[i] = 0.;
    return;
  }
  /* tau_mat, :324-325 (2-D), :330-332 (3-D) */
  double tr[3][3], ti[3][3];
  const int p2[3][2] = {{0, 0}, {1, 1}, {0, 1}};
  const int p3[6][2] = {{0, 0}, {1, 1}, {2, 2}, {1, 2}, {2, 0}, {0, 1}};
  for (int s = 0; s < nsym; s++) {
    int p = dim == 2 ? p2[s][0] : p3[s][0], q = dim == 2 ? p2[s][1] : p3[s][1];
    if (p == q) { tr[p][p] = tau[2 * s]; ti[p][p] = tau[2 * s + 1]; }
    else {
      tr[p][q] = tr[q][p] = tau[2 * s] / sqrt2;
      ti[p][q] = ti[q][p] = tau[2 * s + 1] / sqrt2;
    }
  }
  /* rhs = tau_mat * conj(B), :340 */
  double ur[3], ui[3];
  for (int i = 0; i < dim; i++) {
    double accr = 0., acci = 0.;
    for (int j = 0; j < dim; j++) {
      double br = Bf[2 * j], bi = -Bf[2 * j + 1];
      double pr = tr[i][j] * br - ti[i][j] * bi;
      double pi = tr[i][j] * bi + ti[i][j] * br;
      if (j == 0) { accr = pr; acci = pi; } else { accr = accr + pr; acci = acci + pi; }
    }
    ur[i] = accr; ui[i] = acci;
  }
  /* u = K.llt().solve(rhs), :341 */
  double A[3][3];
  for (int i = 0; i < dim; i++)
    for (int j = 0; j < dim; j++) A[i][j] = Kf[2 * (dim * i + j)];
  /* llt_inplace<Lower>::unblocked: x = A_kk - A10.squaredNorm(); A21 -= A20 * A10^H; A21 /= x */
  for (int j = 0; j < dim; j++) {
    double d = A[j][j];
    if (j > 0) {
      double sq = A[j][0] * A[j][0];
      for (int p = 1; p < j; p++) sq = sq + A[j][p] * A[j][p];
      d = d - sq;
    }
    d = sqrt(d);
    A[j][j] = d;
    for (int i = j + 1; i < dim; i++) {
      double t = A[i][j];
      if (j > 0) {
        double dot = A[i][0] * A[j][0];
        for (int p = 1; p < j; p++) dot = dot + A[i][p] * A[j][p];
        t = t - dot;
      }
      A[i][j] = t / d;
    }
  }
  /* matrixL().solveInPlace: x_i = (x_i - sum_p L_ip x_p) / L_ii, the sum formed first */
  for (int i = 0; i < dim; i++) {
    double sr = ur[i], si = ui[i];
    if (i > 0) {
      double ar = A[i][0] * ur[0], ai = A[i][0] * ui[0];
      for (int p = 1; p < i; p++) { ar = ar + A[i][p] * ur[p]; ai = ai + A[i][p] * ui[p]; }
      sr = sr - ar; si = si - ai;
    }
    ur[i] = sr / A[i][i]; ui[i] = si / A[i][i];
  }
  /* matrixU().solveInPlace (U = L^H), last row first, sums over increasing column index */
  for (int i = dim - 1; i >= 0; i--) {
    double sr = ur[i], si = ui[i];
    if (i < dim - 1) {
      double ar = A[i + 1][i] * ur[i + 1], ai = A[i + 1][i] * ui[i + 1];
      for (int p = i + 2; p < dim; p++) { ar = ar + A[p][i] * ur[p]; ai = ai + A[p][i] * ui[p]; }
      sr = sr - ar; si = si - ai;
    }
    ur[i] = sr / A[i][i]; ui[i] = si / A[i][i];
  }
  if (out_u) for (int i = 0; i < dim; i++) { out_u[2 * i] = ur[i]; out_u[2 * i + 1] = ui[i]; }
  /* eta_mat = 0.5 (B u^T + u B^T), :342; Mandel with sqrt2 on shear, :344-353 */
  for (int s = 0; s < nsym; s++) {
    int p = dim == 2 ? p2[s][0] : p3[s][0], q = dim == 2 ? p2[s][1] : p3[s][1];
    double t1r, t1i, t2r, t2i;
    cmul(Bf[2 * p], Bf[2 * p + 1], ur[q], ui[q], &t1r, &t1i);
    cmul(ur[p], ui[p], Bf[2 * q], Bf[2 * q + 1], &t2r, &t2i);
    double er = 0.5 * (t1r + t2r), ei = 0.5 * (t1i + t2i);
    if (p != q) { er = sqrt2 * er; ei = sqrt2 * ei; }
    eta[2 * s] = er; eta[2 * s + 1] = ei;
  }
}

/*
 * The per-mode loop of python/demo.py:33-40 over a row-major block, planar
 * fields: eta^ (nsym comps) and u^ (dim comps) from tau^ (nsym comps).  Either
 * output may be NULL.
 */
void oracle_apply_eigenstress(int dim, const int *shape, const double *L, double mu, double nu,
                              const int *k_begin, const int *local_shape, const double *tau_hat,
                              double *eta_hat, double *u_hat) {
  const int n0 = local_shape[0], n1 = local_shape[1], n2 = dim == 3 ? local_shape[2] : 1;
  const int nsym = dim == 2 ? 3 : 6;
  const int64_t M = (int64_t)n0 * n1 * n2;
  for (int a = 0; a < n0; a++)
    for (int b = 0; b < n1; b++)
      for (int c = 0; c < n2; c++) {
        int k[3] = {k_begin[0] + a, k_begin[1] + b, dim == 3 ? k_begin[2] + c : 0};
        int64_t i = ((int64_t)a * n1 + b) * n2 + c;
        double tau[12], eta[12], u[6];
        for (int s = 0; s < nsym; s++) { tau[2 * s] = tau_hat[2 * (i + s * M)]; tau[2 * s + 1] = tau_hat[2 * (i + s * M) + 1]; }
        oracle_modal_eigenstress_to_opposite_strain(dim, shape, L, mu, nu, k, tau, eta, u);
        if (eta_hat) for (int s = 0; s < nsym; s++) { eta_hat[2 * (i + s * M)] = eta[2 * s]; eta_hat[2 * (i + s * M) + 1] = eta[2 * s + 1]; }
        if (u_hat) for (int s = 0; s < dim; s++) { u_hat[2 * (i + s * M)] = u[2 * s]; u_hat[2 * (i + s * M) + 1] = u[2 * s + 1]; }
      }
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
