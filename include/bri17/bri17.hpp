// bri17.hpp -- drop-in header of the B200-native bri17 build.
//
// Same public surface as the reference header (include/bri17/bri17.hpp of
// sbrisard/bri17; citations below are file:line in that repository):
//
//   bri17::CartesianGrid<T, DIM>   shape, L, size, num_nodes_per_cell, repr(),
//                                  get_node_at(), get_cell_nodes()      (:34-159)
//   bri17::Hooke<T, DIM>           mu, nu, grid, repr(),
//                                  modal_strain_displacement(k, B)      (:212-236)
//                                  modal_stiffness(k, K)                (:247-292)
//                                  modal_eigenstress_to_opposite_strain (:308-355)
//   operator<< for both                                                (:162-165, :359-362)
//
// The per-frequency methods stay host-side C++ and evaluate exactly the
// reference's floating-point expressions in the reference's order (they are
// what the parity tests compare against).  They need no Eigen.
//
// NEW: bri17::ModalOperator<DIM> -- the whole-grid operator that the reference
// only has as a loop nest in its test harness (tests/test_bri17.cpp:56-107).
// It forwards to the C ABI of libbri17_b200.so (include/bri17_b200.h), i.e. to
// the hand-written sm_100a kernels.  There is no CPU fallback: the constructor
// throws std::runtime_error when no CUDA device is available.
//
// Define BRI17_NO_DEVICE before including this header to get the per-mode API
// only (no dependency on libbri17_b200.so).
#pragma once

#include <array>
#include <cmath>
#include <complex>
#include <concepts>
#include <cstddef>
#include <cstdint>
#include <numbers>
#include <ostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <typeinfo>

#ifndef BRI17_NO_DEVICE
#include "bri17_b200.h"
#ifdef BRI17_WITH_REALSPACE
#include "bri17_b200_realspace.h"
#endif
#endif

namespace bri17 {

template <typename T, int DIM>
concept ValidGridSpec = std::floating_point<T> && (DIM == 2 || DIM == 3);

// ---------------------------------------------------------------------------
// CartesianGrid (reference :34-159)
// ---------------------------------------------------------------------------
template <typename T, int DIM>
  requires ValidGridSpec<T, DIM>
class CartesianGrid {
 public:
  static constexpr int num_nodes_per_cell = 1 << DIM;  // :39

  std::array<int, DIM> const shape;  // cells per direction (:42)
  std::array<T, DIM> const L;        // edge lengths (:45)
  int const size;                    // product of shape, kept as int like the reference (:48)

  CartesianGrid(std::array<int, DIM> shape_, std::array<T, DIM> L_)
      : shape{shape_}, L{L_}, size{product(shape_)} {}

  // Exact number of cells.  `size` is an int for API fidelity and overflows
  // beyond 2^31-1 cells; every batched path of this build uses this instead.
  std::int64_t size64() const {
    std::int64_t n = 1;
    for (int extent : shape) n *= extent;
    return n;
  }

  std::string repr() const {  // same text as :61-69
    std::ostringstream out;
    out << "CartesianGrid<" << typeid(T).name() << "," << DIM << ">{shape={";
    for (int extent : shape) out << extent << ",";
    out << "},L={";
    for (T length : L) out << length << ",";
    out << "}}";
    return out.str();
  }

  // Row-major node index (:78-93).
  int get_node_at(int i, int j) const {
    static_assert(DIM == 2, "this method expects a 2D grid");
    return i * shape[1] + j;
  }
  int get_node_at(int i, int j, int k) const {
    static_assert(DIM == 3, "this method expects a 3D grid");
    return (i * shape[1] + j) * shape[2] + k;
  }

  // Vertices of a cell, periodic, last axis fastest inside the cell (:127-158).
  std::array<int, num_nodes_per_cell> get_cell_nodes(int cell) const {
    int origin[DIM];
    for (int d = DIM - 1; d >= 0; d--) {
      origin[d] = cell % shape[d];
      cell /= shape[d];
    }
    std::array<int, num_nodes_per_cell> nodes;
    for (int local = 0; local < num_nodes_per_cell; local++) {
      int node = 0;
      for (int d = 0; d < DIM; d++) {
        const int step = (local >> (DIM - 1 - d)) & 1;
        const int coord = (step && origin[d] == shape[d] - 1) ? 0 : origin[d] + step;
        node = node * shape[d] + coord;
      }
      nodes[local] = node;
    }
    return nodes;
  }

 private:
  static int product(const std::array<int, DIM> &s) {
    int n = 1;
    for (int extent : s) n *= extent;
    return n;
  }
};

template <typename T, int DIM>
std::ostream &operator<<(std::ostream &os, const CartesianGrid<T, DIM> &grid) {
  return os << grid.repr();
}

namespace detail {

// Per-axis factors of the modal stiffness for one frequency index (:258-264).
template <typename T>
struct AxisStiffnessFactors {
  T phi, chi, psi;
  AxisStiffnessFactors(int k, int n, T length) {
    const T h = length / n;
    const T beta = 2 * std::numbers::pi_v<T> * k / n;
    phi = 2 * (1 - std::cos(beta)) / h / h;
    chi = (2 + std::cos(beta)) / 3;
    psi = std::sin(beta) / h;
  }
};

// In-place Cholesky factorisation of a real SPD DIM x DIM matrix (lower
// triangle, row-major) and solve for a complex right-hand side.  Replaces
// Eigen's K.llt().solve(rhs) of the reference (:341); K^ has zero imaginary
// part, so the factor is real.  The operation order is the one Eigen 3.3/3.4
// uses for fixed sizes: llt_inplace<Lower>::unblocked (x = A_kk -
// A10.squaredNorm(); A21 -= A20 * A10^H; A21 /= x), then the unrolled
// triangular solves (rhs_i -= (row_i . rhs).sum(); rhs_i /= L_ii).
template <typename T, int DIM>
void cholesky_solve(T (&A)[DIM][DIM], std::complex<T> (&x)[DIM]) {
  for (int j = 0; j < DIM; j++) {
    T d = A[j][j];
    if (j > 0) {
      T sq = A[j][0] * A[j][0];
      for (int p = 1; p < j; p++) sq = sq + A[j][p] * A[j][p];
      d = d - sq;
    }
    d = std::sqrt(d);
    A[j][j] = d;
    for (int i = j + 1; i < DIM; i++) {
      T s = A[i][j];
      if (j > 0) {
        T dot = A[i][0] * A[j][0];
        for (int p = 1; p < j; p++) dot = dot + A[i][p] * A[j][p];
        s = s - dot;
      }
      A[i][j] = s / d;
    }
  }
  for (int i = 0; i < DIM; i++) {  // L y = b
    std::complex<T> s = x[i];
    if (i > 0) {
      std::complex<T> acc = A[i][0] * x[0];
      for (int p = 1; p < i; p++) acc = acc + A[i][p] * x[p];
      s = s - acc;
    }
    x[i] = s / A[i][i];
  }
  for (int i = DIM - 1; i >= 0; i--) {  // L^T x = y
    std::complex<T> s = x[i];
    if (i < DIM - 1) {
      std::complex<T> acc = A[i + 1][i] * x[i + 1];
      for (int p = i + 2; p < DIM; p++) acc = acc + A[p][i] * x[p];
      s = s - acc;
    }
    x[i] = s / A[i][i];
  }
}

}  // namespace detail

// ---------------------------------------------------------------------------
// Hooke (reference :176-356)
// ---------------------------------------------------------------------------
template <typename T, int DIM>
  requires ValidGridSpec<T, DIM>
class Hooke {
 public:
  T const mu;                        // shear modulus (:180)
  T const nu;                        // Poisson ratio (:183)
  CartesianGrid<T, DIM> const grid;  // by-value copy (:186)

  // Non-const lvalue reference like the reference (:193), so that CTAD
  // `bri17::Hooke hooke{mu, nu, grid}` keeps working (tests/test_bri17.cpp:338).
  Hooke(T mu_, T nu_, CartesianGrid<T, DIM> &grid_) : mu{mu_}, nu{nu_}, grid{grid_} {}

  std::string repr() const {  // same text as :197-202 (ends with a newline, no closing brace)
    std::ostringstream out;
    out << "Hooke<" << typeid(T).name() << "," << DIM << ">{mu=" << mu << ",nu=" << nu
        << ",grid=" << grid << std::endl;
    return out.str();
  }

  // B^[k, :] -- DIM complex numbers (:212-236).  Half angles: NOT N-periodic
  // in k, so k must be given in [0, N).
  void modal_strain_displacement(int const *k, std::complex<T> *B) const {
    T c[DIM], s[DIM];
    T sum_alpha{};
    for (int d = 0; d < DIM; d++) {
      const T alpha = std::numbers::pi_v<T> * k[d] / grid.shape[d];
      sum_alpha += alpha;
      c[d] = std::cos(alpha);
      s[d] = std::sin(alpha) * grid.shape[d] / grid.L[d];
    }
    const std::complex<T> prefactor{-2 * std::sin(sum_alpha), 2 * std::cos(sum_alpha)};
    for (int i = 0; i < DIM; i++) {
      // prefactor * f_0 * f_1 (* f_2), f_d = s[d] on the diagonal, c[d] elsewhere,
      // multiplied left to right as written at :227-232
      std::complex<T> b = prefactor;
      for (int d = 0; d < DIM; d++) b = b * (d == i ? s[d] : c[d]);
      B[i] = b;
    }
  }

  // K^[k, :, :] -- DIM*DIM complex numbers, row-major, zero imaginary part
  // (:247-292).  k may be any integer (cos/sin are N-periodic).
  void modal_stiffness(int const *k, std::complex<T> *K) const {
    T phi[DIM], chi[DIM], psi[DIM];
    for (int d = 0; d < DIM; d++) {
      const detail::AxisStiffnessFactors<T> f(k[d], grid.shape[d], grid.L[d]);
      phi[d] = f.phi;
      chi[d] = f.chi;
      psi[d] = f.psi;
    }
    const double scaling = mu / (1. - 2. * nu);  // double whatever T is (:266)
    if constexpr (DIM == 2) {
      const auto H00 = phi[0] * chi[1];
      const auto H11 = chi[0] * phi[1];
      const auto Kd = mu * (H00 + H11);
      K[0] = scaling * H00 + Kd;
      K[1] = scaling * psi[0] * psi[1];
      K[2] = K[1];
      K[3] = scaling * H11 + Kd;
    } else {
      const auto H00 = phi[0] * chi[1] * chi[2];
      const auto H11 = chi[0] * phi[1] * chi[2];
      const auto H22 = chi[0] * chi[1] * phi[2];
      const auto Kd = mu * (H00 + H11 + H22);
      K[0] = scaling * H00 + Kd;
      K[4] = scaling * H11 + Kd;
      K[8] = scaling * H22 + Kd;
      K[1] = K[3] = scaling * psi[0] * psi[1] * chi[2];
      K[2] = K[6] = scaling * psi[0] * chi[1] * psi[2];
      K[5] = K[7] = scaling * chi[0] * psi[1] * psi[2];
    }
  }

  // eta^[k] = -eps^[k] induced by the eigenstress tau^[k], Mandel notation
  // (:308-355): solve K^ u = tau . conj(B^), eta = sym(B^ (x) u); zero at k = 0.
  // The reference delegates the solve to Eigen's LLT (unpinned by any
  // reference test); this build uses detail::cholesky_solve.
  void modal_eigenstress_to_opposite_strain(int const *k, std::complex<T> const *tau,
                                            std::complex<T> *eta) const {
    constexpr int sym = (DIM * (DIM + 1)) / 2;
    constexpr T sqrt2 = std::numbers::sqrt2_v<T>;
    bool null_frequency = true;
    for (int d = 0; d < DIM; d++) null_frequency = null_frequency && (k[d] == 0);
    if (null_frequency) {
      for (int i = 0; i < sym; i++) eta[i] = std::complex<T>{};
      return;
    }
    std::complex<T> B[DIM], Kc[DIM * DIM];
    modal_strain_displacement(k, B);
    modal_stiffness(k, Kc);
    // Mandel -> tensor: shear entries carry 1/sqrt2 (:324-325, :330-332)
    std::complex<T> t[DIM][DIM];
    for (int i = 0; i < DIM; i++) t[i][i] = tau[i];
    if constexpr (DIM == 2) {
      t[0][1] = t[1][0] = tau[2] / sqrt2;
    } else {
      t[1][2] = t[2][1] = tau[3] / sqrt2;
      t[2][0] = t[0][2] = tau[4] / sqrt2;
      t[0][1] = t[1][0] = tau[5] / sqrt2;
    }
    std::complex<T> u[DIM];
    T A[DIM][DIM];
    for (int i = 0; i < DIM; i++) {
      u[i] = std::complex<T>{};
      for (int j = 0; j < DIM; j++) {
        u[i] += t[i][j] * std::conj(B[j]);  // rhs = tau . conj(B) (:340)
        A[i][j] = Kc[DIM * i + j].real();
      }
    }
    detail::cholesky_solve<T, DIM>(A, u);  // :341
    auto e = [&](int i, int j) { return T(0.5) * (B[i] * u[j] + u[i] * B[j]); };  // :342
    for (int i = 0; i < DIM; i++) eta[i] = e(i, i);
    if constexpr (DIM == 2) {
      eta[2] = sqrt2 * e(0, 1);
    } else {
      eta[3] = sqrt2 * e(1, 2);
      eta[4] = sqrt2 * e(2, 0);
      eta[5] = sqrt2 * e(0, 1);
    }
  }
};

template <typename T, int DIM>
std::ostream &operator<<(std::ostream &os, const Hooke<T, DIM> &hooke) {
  return os << hooke.repr();
}

// Aliases in the spelling of the reference's older documentation (docs/cpp_api.html).
template <int DIM>
using CartesianGridD = CartesianGrid<double, DIM>;
template <int DIM>
using HookeD = Hooke<double, DIM>;

#ifndef BRI17_NO_DEVICE
// ---------------------------------------------------------------------------
// ModalOperator: every frequency at once, on the GPU.
//
// Fields are planar by component, interleaved complex<double>, element (c, i)
// at buf[i + c*comp_stride] -- the layout of the reference harness
// (tests/test_bri17.cpp:66-67, :81-83).  All buffers are caller-owned DEVICE
// memory except in apply_modal_stiffness_host.  `stream` is a cudaStream_t
// passed as void* so that this header needs no CUDA include.
// ---------------------------------------------------------------------------
template <int DIM>
  requires(DIM == 2 || DIM == 3)
class ModalOperator {
 public:
  using complex_t = std::complex<double>;

  explicit ModalOperator(const Hooke<double, DIM> &hooke, int device = 0) : hooke_{hooke} {
    check(bri17_plan_create(&plan_, DIM, hooke.grid.shape.data(), hooke.grid.L.data(), hooke.mu,
                            hooke.nu, device));
  }
  ~ModalOperator() { bri17_plan_destroy(plan_); }
  ModalOperator(const ModalOperator &) = delete;
  ModalOperator &operator=(const ModalOperator &) = delete;

  const Hooke<double, DIM> &hooke() const { return hooke_; }

  // f^[c,k] = sum_j K^[k][c,j] u^[j,k] over the whole grid (tests/test_bri17.cpp:58-92).
  void apply_modal_stiffness(const complex_t *u_hat_dev, complex_t *f_hat_dev,
                             void *stream = nullptr, double out_scale = 1.0) const {
    check(bri17_modal_stiffness_apply_f64(plan_, u_hat_dev, f_hat_dev, nullptr, nullptr, 0,
                                          out_scale, stream));
  }
  // ... over the block k_begin + [0, local_shape) (a k0 slab per GPU).
  void apply_modal_stiffness(const complex_t *u_hat_dev, complex_t *f_hat_dev,
                             const std::array<int, DIM> &k_begin,
                             const std::array<int, DIM> &local_shape, std::int64_t comp_stride = 0,
                             void *stream = nullptr, double out_scale = 1.0) const {
    check(bri17_modal_stiffness_apply_f64(plan_, u_hat_dev, f_hat_dev, k_begin.data(),
                                          local_shape.data(), comp_stride, out_scale, stream));
  }
  // Host buffers (pinned memory recommended); synchronous.
  void apply_modal_stiffness_host(const complex_t *u_hat, complex_t *f_hat) const {
    check(bri17_modal_stiffness_apply_host_f64(plan_, u_hat, f_hat, nullptr, nullptr, 0, 1.0));
  }
  // eps^ = sym(B^ (x) u^) in Mandel order (tests/test_bri17.cpp:194-235).
  void apply_strain_displacement(const complex_t *u_hat_dev, complex_t *eps_hat_dev,
                                 void *stream = nullptr) const {
    check(bri17_strain_displacement_apply_f64(plan_, u_hat_dev, eps_hat_dev, nullptr, nullptr, 0, 0,
                                              1.0, stream));
  }
  // K^[k] / B^[k] for every k, mode-major (a loop over Hooke::modal_* would write the same).
  void modal_stiffness_field(complex_t *K_dev, void *stream = nullptr) const {
    check(bri17_modal_stiffness_field_f64(plan_, K_dev, nullptr, nullptr, stream));
  }
  void modal_strain_displacement_field(complex_t *B_dev, void *stream = nullptr) const {
    check(bri17_modal_strain_displacement_field_f64(plan_, B_dev, nullptr, nullptr, stream));
  }
  void freq_index_map(std::int32_t *k_out_dev, void *stream = nullptr) const {
    check(bri17_freq_index_map(plan_, k_out_dev, nullptr, nullptr, stream));
  }
  // Per-mode direct solves over the whole grid, planar fields (Hooke::modal_eigenstress_to_opposite_strain
  // batched, reference :308-355): u^ = K^-1 f^; f^ = tau^ . conj(B^) (:340); u^ = K^-1 f^ of that (:341);
  // eta^ = sym(B^ (x) u^) (:342-353).
  void solve_modal_stiffness(const complex_t *f_hat_dev, complex_t *u_hat_dev, void *stream = nullptr) const {
    check(bri17_modal_stiffness_solve_f64(plan_, f_hat_dev, u_hat_dev, nullptr, nullptr, 0, 0, stream));
  }
  void eigenstress_to_force(const complex_t *tau_hat_dev, complex_t *f_hat_dev, void *stream = nullptr) const {
    check(bri17_eigenstress_to_force_f64(plan_, tau_hat_dev, f_hat_dev, nullptr, nullptr, 0, 0, 0, 0, stream));
  }
  void eigenstress_to_displacement(const complex_t *tau_hat_dev, complex_t *u_hat_dev,
                                   void *stream = nullptr) const {
    check(bri17_eigenstress_to_displacement_f64(plan_, tau_hat_dev, u_hat_dev, nullptr, nullptr, 0, 0, 0, 0,
                                                stream));
  }
  void eigenstress_to_opposite_strain(const complex_t *tau_hat_dev, complex_t *eta_hat_dev,
                                      void *stream = nullptr) const {
    check(bri17_eigenstress_to_opposite_strain_f64(plan_, tau_hat_dev, eta_hat_dev, nullptr, nullptr, 0, 0,
                                                   stream));
  }

  bri17_plan *c_plan() const { return plan_; }

 private:
  static void check(int rc) {
    if (rc == BRI17_OK) return;
    const std::string msg = bri17_last_error();
    if (rc == BRI17_ERR_INVALID_ARG) throw std::invalid_argument(msg);
    throw std::runtime_error(msg);
  }
  Hooke<double, DIM> hooke_;
  bri17_plan *plan_ = nullptr;
};
#ifdef BRI17_WITH_REALSPACE
// ---------------------------------------------------------------------------
// RealSpaceOperator: F = (|h|/|N|) iDFT(K^ DFT(u)), i.e. the whole
// StiffnessMatrixFactory::compute_Ku of the reference harness
// (tests/test_bri17.cpp:56-107) on the GPU (cuFFT instead of FFTW), optionally
// slab-distributed over `nranks` processes, plus matrix-free CG.  Define
// BRI17_WITH_REALSPACE and link libbri17_b200_rs.so.
// Fields: [DIM][n0_count][N1][(N2)] of this rank's slab, complex<double>
// (real data carried as complex, like the reference) or plain double (r2c path).
// ---------------------------------------------------------------------------
template <int DIM>
  requires(DIM == 2 || DIM == 3)
class RealSpaceOperator {
 public:
  using complex_t = std::complex<double>;

  explicit RealSpaceOperator(const Hooke<double, DIM> &hooke, int device = 0, int rank = 0, int nranks = 1,
                             const void *nccl_unique_id = nullptr, int exchange_mode = 1) {
    check(bri17_rs_plan_create(&plan_, DIM, hooke.grid.shape.data(), hooke.grid.L.data(), hooke.mu, hooke.nu,
                               device, rank, nranks, nccl_unique_id, exchange_mode));
    bri17_rs_plan_local(plan_, &n0_begin_, &n0_count_, nullptr, nullptr);
  }
  ~RealSpaceOperator() { bri17_rs_plan_destroy(plan_); }
  RealSpaceOperator(const RealSpaceOperator &) = delete;
  RealSpaceOperator &operator=(const RealSpaceOperator &) = delete;

  int n0_begin() const { return n0_begin_; }
  int n0_count() const { return n0_count_; }
  std::int64_t slab_count() const { return bri17_rs_plan_real_count(plan_); }  // elements per component

  void apply(const complex_t *u_dev, complex_t *F_dev, void *stream = nullptr) const {
    check(bri17_real_space_apply_f64(plan_, u_dev, F_dev, stream));
  }
  void apply(const double *u_dev, double *F_dev, void *stream = nullptr) const {
    check(bri17_real_space_apply_real_f64(plan_, u_dev, F_dev, stream));
  }
  // F = A u and the global scalar <u, A u> (Parseval sum taken inside the K^ kernel).
  double apply_with_dot(const complex_t *u_dev, complex_t *F_dev, void *stream = nullptr) const {
    double dot = 0.;
    check(bri17_real_space_apply_dot_f64(plan_, u_dev, F_dev, 0, &dot, stream));
    return dot;
  }
  double apply_with_dot(const double *u_dev, double *F_dev, void *stream = nullptr) const {
    double dot = 0.;
    check(bri17_real_space_apply_dot_f64(plan_, u_dev, F_dev, 1, &dot, stream));
    return dot;
  }
  // Conjugate gradients on A x = b (the component means of b are projected out: K^(0) = 0,
  // reference :336-339); returns the iteration count.
  int solve(const complex_t *b_dev, complex_t *x_dev, double rtol = 1e-8, int max_iter = 1000,
            double *rel_residual = nullptr, void *stream = nullptr) const {
    int it = 0;
    check(bri17_cg_solve_f64(plan_, b_dev, x_dev, rtol, max_iter, 10, &it, rel_residual, stream));
    return it;
  }
  int solve(const double *b_dev, double *x_dev, double rtol = 1e-8, int max_iter = 1000,
            double *rel_residual = nullptr, void *stream = nullptr) const {
    int it = 0;
    check(bri17_cg_solve_real_f64(plan_, b_dev, x_dev, rtol, max_iter, 10, &it, rel_residual, stream));
    return it;
  }
  bri17_rs_plan *c_plan() const { return plan_; }

 private:
  static void check(int rc) {
    if (rc == BRI17_OK) return;
    const std::string msg = bri17_last_error();
    if (rc == BRI17_ERR_INVALID_ARG) throw std::invalid_argument(msg);
    throw std::runtime_error(msg);
  }
  bri17_rs_plan *plan_ = nullptr;
  int n0_begin_ = 0, n0_count_ = 0;
};
#endif  // BRI17_WITH_REALSPACE
#endif  // BRI17_NO_DEVICE

}  // namespace bri17
