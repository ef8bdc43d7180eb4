// C++ host program over the drop-in header: builds the operator the way a
// bri17 user would (CartesianGrid + Hooke), runs bri17::ModalOperator on the
// GPU through the C ABI, and compares with the reference-style per-mode loop
// (tests/test_bri17.cpp:76-91) evaluated with the header's own host-side
// Hooke::modal_stiffness.  Prints "max_abs_diff <x>"; exit code 0 iff x == 0.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "bri17/bri17.hpp"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { std::printf("CUDA error %s\n", cudaGetErrorString(e)); return 2; } } while (0)

template <int DIM>
int run(std::array<int, DIM> shape) {
  using C = std::complex<double>;
  std::array<double, DIM> L;
  const double spacing[3] = {1.1, 1.2, 1.3};
  for (int d = 0; d < DIM; d++) L[d] = shape[d] * spacing[d];
  bri17::CartesianGrid<double, DIM> grid{shape, L};
  bri17::Hooke hooke{5.6, 0.3, grid};
  const std::int64_t M = grid.size64();

  std::mt19937_64 rng(42);
  std::normal_distribution<double> gauss;
  std::vector<C> u(DIM * M), f(DIM * M), expected(DIM * M);
  for (auto &x : u) x = C{gauss(rng), gauss(rng)};

  // reference-style loop nest, planar layout
  int k[3] = {0, 0, 0};
  C K[DIM * DIM];
  const int n2 = DIM == 3 ? shape[DIM - 1] : 1;
  std::int64_t i = 0;
  for (k[0] = 0; k[0] < shape[0]; k[0]++)
    for (k[1] = 0; k[1] < shape[1]; k[1]++)
      for (int c = 0; c < n2; c++, i++) {
        if (DIM == 3) k[2] = c;
        hooke.modal_stiffness(k, K);
        for (int r = 0; r < DIM; r++) {
          double re = K[DIM * r].real() * u[i].real(), im = K[DIM * r].real() * u[i].imag();
          for (int j = 1; j < DIM; j++) {
            re = re + K[DIM * r + j].real() * u[i + j * M].real();
            im = im + K[DIM * r + j].real() * u[i + j * M].imag();
          }
          expected[i + r * M] = C{re, im};
        }
      }

  C *du = nullptr, *df = nullptr;
  CK(cudaMalloc(&du, sizeof(C) * DIM * M));
  CK(cudaMalloc(&df, sizeof(C) * DIM * M));
  CK(cudaMemcpy(du, u.data(), sizeof(C) * DIM * M, cudaMemcpyHostToDevice));
  bri17::ModalOperator<DIM> op{hooke};
  op.apply_modal_stiffness(du, df);
  CK(cudaMemcpy(f.data(), df, sizeof(C) * DIM * M, cudaMemcpyDeviceToHost));
  double worst = 0;
  for (std::int64_t j = 0; j < DIM * M; j++) worst = std::max(worst, std::abs(f[j] - expected[j]));
  // host-buffer entry point must agree as well
  std::vector<C> f2(DIM * M);
  op.apply_modal_stiffness_host(u.data(), f2.data());
  for (std::int64_t j = 0; j < DIM * M; j++) worst = std::max(worst, std::abs(f2[j] - expected[j]));
  // error translation: a block outside the grid -> std::invalid_argument
  bool threw = false;
  try {
    std::array<int, DIM> kb{}, ls = shape;
    kb[0] = 1;
    op.apply_modal_stiffness(du, df, kb, ls);
  } catch (const std::invalid_argument &) { threw = true; }
  cudaFree(du);
  cudaFree(df);
  std::printf("dim %d modes %lld max_abs_diff %.3e invalid_argument_thrown %d\n", DIM, (long long)M, worst, int(threw));
  return (worst == 0.0 && threw) ? 0 : 1;
}

int main() {
  int rc = run<2>({48, 80});
  rc |= run<3>({12, 10, 70});
  return rc;
}
