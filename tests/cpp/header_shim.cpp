// C entry points over the per-mode API of include/bri17/bri17.hpp (THIS
// repository's drop-in header), with the same signatures as oracle/ref_driver.cpp
// so that the tests can compare the two implementations bit for bit.
#define BRI17_NO_DEVICE
#include "bri17/bri17.hpp"

#include <cstring>

namespace {
template <int DIM>
bri17::Hooke<double, DIM> make(const int *shape, const double *L, double mu, double nu) {
  std::array<int, DIM> s;
  std::array<double, DIM> l;
  for (int i = 0; i < DIM; i++) { s[i] = shape[i]; l[i] = L[i]; }
  bri17::CartesianGrid<double, DIM> grid{s, l};
  bri17::Hooke hooke{mu, nu, grid};  // CTAD, as in tests/test_bri17.cpp:338
  return hooke;
}
}  // namespace

extern "C" {
void hdr_modal_stiffness(int dim, const int *shape, const double *L, double mu, double nu,
                         const int *k, double *K) {
  auto *Kc = reinterpret_cast<std::complex<double> *>(K);
  if (dim == 2) make<2>(shape, L, mu, nu).modal_stiffness(k, Kc);
  else make<3>(shape, L, mu, nu).modal_stiffness(k, Kc);
}
void hdr_modal_strain_displacement(int dim, const int *shape, const double *L, const int *k, double *B) {
  auto *Bc = reinterpret_cast<std::complex<double> *>(B);
  if (dim == 2) make<2>(shape, L, 1.0, 0.25).modal_strain_displacement(k, Bc);
  else make<3>(shape, L, 1.0, 0.25).modal_strain_displacement(k, Bc);
}
void hdr_eigenstress_to_opposite_strain(int dim, const int *shape, const double *L, double mu, double nu,
                                        const int *k, const double *tau, double *eta) {
  auto *t = reinterpret_cast<const std::complex<double> *>(tau);
  auto *e = reinterpret_cast<std::complex<double> *>(eta);
  if (dim == 2) make<2>(shape, L, mu, nu).modal_eigenstress_to_opposite_strain(k, t, e);
  else make<3>(shape, L, mu, nu).modal_eigenstress_to_opposite_strain(k, t, e);
}
void hdr_get_cell_nodes(int dim, const int *shape, int cell, int *nodes) {
  const double L[3] = {1., 1., 1.};
  if (dim == 2) { auto n = make<2>(shape, L, 1., .25).grid.get_cell_nodes(cell); std::memcpy(nodes, n.data(), sizeof(int) * 4); }
  else { auto n = make<3>(shape, L, 1., .25).grid.get_cell_nodes(cell); std::memcpy(nodes, n.data(), sizeof(int) * 8); }
}
int hdr_repr(int dim, const int *shape, const double *L, double mu, double nu, int which, char *out, int cap) {
  std::ostringstream os;
  if (dim == 2) { auto h = make<2>(shape, L, mu, nu); if (which == 0) os << h.grid; else os << h; }
  else { auto h = make<3>(shape, L, mu, nu); if (which == 0) os << h.grid; else os << h; }
  std::strncpy(out, os.str().c_str(), cap - 1);
  out[cap - 1] = 0;
  return int(os.str().size());
}
int hdr_float_instantiates(void) {
  // the reference is a template over any floating-point T (bri17.hpp:34-36)
  bri17::CartesianGrid<float, 2> g{{4, 4}, {1.f, 1.f}};
  bri17::Hooke<float, 2> h{1.f, 0.3f, g};
  int k[2] = {1, 2};
  std::complex<float> K[4];
  h.modal_stiffness(k, K);
  return K[0].real() > 0.f && K[1] == K[2];
}
}
