// ref_driver.cpp -- C entry points over the UNMODIFIED reference header.
//
// TEST INFRASTRUCTURE ONLY (see oracle/bri17_oracle.c for the rules).  This
// translation unit is compiled with -I/root/reference/include, i.e. the
// arithmetic of modal_stiffness / modal_strain_displacement / get_cell_nodes
// executed here IS the reference's (include/bri17/bri17.hpp:34-292); no
// reference source is copied into this repository.  What is restated here is
// only the harness loop that the reference keeps in its Catch2 test TU
// (tests/test_bri17.cpp:58-92 and :194-235), because that TU needs Eigen,
// FFTW and Catch2, none of which is installed; the Eigen fixed-size product
// K_k * u_k (:68, :84) is written out as a left-to-right real*complex matvec.
//
// Outputs go to oracle/_ref/ (git-ignored, but shipped to the GPU box).
#include <sstream>   // the reference header relies on it (tests/test_bri17.cpp:1)
#include <typeinfo>
#include <cstdint>
#include <cstring>
#include <complex>

#include "bri17/bri17.hpp"

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

template <int DIM>
bri17::Hooke<double, DIM> make_hooke(const int *shape, const double *L,
                                     double mu, double nu) {
  std::array<int, DIM> s;
  std::array<double, DIM> l;
  for (int i = 0; i < DIM; i++) { s[i] = shape[i]; l[i] = L[i]; }
  bri17::CartesianGrid<double, DIM> grid{s, l};
  return bri17::Hooke<double, DIM>{mu, nu, grid};
}

template <int DIM>
void apply_K(const int *shape, const double *L, double mu, double nu,
             const int *k_begin, const int *local_shape, int64_t comp_stride,
             const std::complex<double> *u_hat, std::complex<double> *f_hat,
             int nthreads) {
  const auto hooke = make_hooke<DIM>(shape, L, mu, nu);
  const int n0 = local_shape[0], n1 = local_shape[1];
  const int n2 = DIM == 3 ? local_shape[2] : 1;
  (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
#endif
  for (int a = 0; a < n0; a++) {
    int k[3] = {k_begin[0] + a, 0, 0};
    std::complex<double> K[DIM * DIM];
    for (int b = 0; b < n1; b++) {
      k[1] = k_begin[1] + b;
      for (int c = 0; c < n2; c++) {
        if (DIM == 3) k[2] = k_begin[2] + c;
        const int64_t i = (int64_t(a) * n1 + b) * n2 + c;
        hooke.modal_stiffness(k, K);
        double ur[DIM], ui[DIM];
        for (int j = 0; j < DIM; j++) {
          ur[j] = u_hat[i + j * comp_stride].real();
          ui[j] = u_hat[i + j * comp_stride].imag();
        }
        for (int r = 0; r < DIM; r++) {
          double fr = K[DIM * r].real() * ur[0];
          double fi = K[DIM * r].real() * ui[0];
          for (int j = 1; j < DIM; j++) {
            fr = fr + K[DIM * r + j].real() * ur[j];
            fi = fi + K[DIM * r + j].real() * ui[j];
          }
          f_hat[i + r * comp_stride] = {fr, fi};
        }
      }
    }
  }
}

template <int DIM>
void apply_B(const int *shape, const double *L, const int *k_begin,
             const int *local_shape, int64_t u_stride, int64_t e_stride,
             const std::complex<double> *u_hat, std::complex<double> *eps_hat) {
  const auto hooke = make_hooke<DIM>(shape, L, 1.0, 0.25);
  const int n0 = local_shape[0], n1 = local_shape[1];
  const int n2 = DIM == 3 ? local_shape[2] : 1;
  constexpr int nsym = DIM == 2 ? 3 : 6;
  const int p2[3][2] = {{0, 0}, {1, 1}, {0, 1}};
  const int p3[6][2] = {{0, 0}, {1, 1}, {2, 2}, {1, 2}, {2, 0}, {0, 1}};
  for (int a = 0; a < n0; a++)
    for (int b = 0; b < n1; b++)
      for (int c = 0; c < n2; c++) {
        int k[3] = {k_begin[0] + a, k_begin[1] + b, DIM == 3 ? k_begin[2] + c : 0};
        const int64_t i = (int64_t(a) * n1 + b) * n2 + c;
        std::complex<double> B[DIM], u[DIM];
        hooke.modal_strain_displacement(k, B);
        for (int j = 0; j < DIM; j++) u[j] = u_hat[i + j * u_stride];
        for (int s = 0; s < nsym; s++) {
          const int p = DIM == 2 ? p2[s][0] : p3[s][0];
          const int q = DIM == 2 ? p2[s][1] : p3[s][1];
          std::complex<double> e = 0.5 * (B[p] * u[q] + u[p] * B[q]);
          if (p != q) e = sqrt(2) * e;
          eps_hat[i + s * e_stride] = e;
        }
      }
}

}  // namespace

extern "C" {

void ref_modal_stiffness(int dim, const int *shape, const double *L, double mu,
                         double nu, const int *k, double *K) {
  auto *Kc = reinterpret_cast<std::complex<double> *>(K);
  if (dim == 2) make_hooke<2>(shape, L, mu, nu).modal_stiffness(k, Kc);
  else make_hooke<3>(shape, L, mu, nu).modal_stiffness(k, Kc);
}

void ref_modal_strain_displacement(int dim, const int *shape, const double *L,
                                   const int *k, double *B) {
  auto *Bc = reinterpret_cast<std::complex<double> *>(B);
  if (dim == 2) make_hooke<2>(shape, L, 1.0, 0.25).modal_strain_displacement(k, Bc);
  else make_hooke<3>(shape, L, 1.0, 0.25).modal_strain_displacement(k, Bc);
}

void ref_get_cell_nodes(int dim, const int *shape, int cell, int *nodes) {
  const double L[3] = {1., 1., 1.};
  if (dim == 2) {
    auto n = make_hooke<2>(shape, L, 1., .25).grid.get_cell_nodes(cell);
    for (int i = 0; i < 4; i++) nodes[i] = n[i];
  } else {
    auto n = make_hooke<3>(shape, L, 1., .25).grid.get_cell_nodes(cell);
    for (int i = 0; i < 8; i++) nodes[i] = n[i];
  }
}

int ref_grid_size(int dim, const int *shape) {
  const double L[3] = {1., 1., 1.};
  return dim == 2 ? make_hooke<2>(shape, L, 1., .25).grid.size
                  : make_hooke<3>(shape, L, 1., .25).grid.size;
}

int ref_repr(int dim, const int *shape, const double *L, double mu, double nu,
             int which, char *out, int cap) {
  std::string s;
  if (dim == 2) {
    auto h = make_hooke<2>(shape, L, mu, nu);
    s = which == 0 ? h.grid.repr() : h.repr();
  } else {
    auto h = make_hooke<3>(shape, L, mu, nu);
    s = which == 0 ? h.grid.repr() : h.repr();
  }
  std::strncpy(out, s.c_str(), cap - 1);
  out[cap - 1] = 0;
  return int(s.size());
}

void ref_apply_modal_stiffness(int dim, const int *shape, const double *L,
                               double mu, double nu, const int *k_begin,
                               const int *local_shape, int64_t comp_stride,
                               const double *u_hat, double *f_hat,
                               int nthreads) {
  auto *u = reinterpret_cast<const std::complex<double> *>(u_hat);
  auto *f = reinterpret_cast<std::complex<double> *>(f_hat);
  if (dim == 2) apply_K<2>(shape, L, mu, nu, k_begin, local_shape, comp_stride, u, f, nthreads);
  else apply_K<3>(shape, L, mu, nu, k_begin, local_shape, comp_stride, u, f, nthreads);
}

void ref_apply_strain_displacement(int dim, const int *shape, const double *L,
                                   const int *k_begin, const int *local_shape,
                                   int64_t u_stride, int64_t e_stride,
                                   const double *u_hat, double *eps_hat) {
  auto *u = reinterpret_cast<const std::complex<double> *>(u_hat);
  auto *e = reinterpret_cast<std::complex<double> *>(eps_hat);
  if (dim == 2) apply_B<2>(shape, L, k_begin, local_shape, u_stride, e_stride, u, e);
  else apply_B<3>(shape, L, k_begin, local_shape, u_stride, e_stride, u, e);
}

int ref_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
