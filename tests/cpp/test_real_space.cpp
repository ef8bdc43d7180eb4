// C++ end-to-end test over the drop-in header, in the style of the reference's
// "Global assembly tests" (tests/test_bri17.cpp:130-150): the global stiffness
// matrix is built column by column, K[:, j] = real_space_apply(e_j), on the
// reference's own grids (3,4) and (3,4,5) with mu = 5.6, nu = 0.3, and compared
// with the same columns computed on the host by a naive O(N^2) DFT sandwich
// around Hooke::modal_stiffness (no FFTW here).  Also checks symmetry of K,
// the r2c path, and CG against a known solution.
#include <cuda_runtime.h>

#include <cstdio>
#include <vector>

#define BRI17_WITH_REALSPACE
#include "bri17/bri17.hpp"

template <int DIM>
static int run(std::array<int, DIM> shape) {
  using C = std::complex<double>;
  const double spacing[3] = {1.1, 1.2, 1.3};
  std::array<double, DIM> L;
  for (int d = 0; d < DIM; d++) L[d] = shape[d] * spacing[d];
  bri17::CartesianGrid<double, DIM> grid{shape, L};
  bri17::Hooke hooke{5.6, 0.3, grid};
  const int M = grid.size, ndof = DIM * M;

  // host reference: dense K through explicit DFT matrices
  std::vector<C> Khat(size_t(M) * DIM * DIM);
  auto unravel = [&](int i, int *k) { for (int d = DIM - 1; d >= 0; d--) { k[d] = i % shape[d]; i /= shape[d]; } };
  for (int i = 0; i < M; i++) { int k[3]; unravel(i, k); hooke.modal_stiffness(k, &Khat[size_t(i) * DIM * DIM]); }
  double cell_volume = 1;
  for (int d = 0; d < DIM; d++) cell_volume *= L[d] / shape[d];
  auto phase = [&](int n, int k) {  // exp(-i phi[n,k]), theory.rst eqs (1)-(2)
    int nn[3], kk[3]; unravel(n, nn); unravel(k, kk);
    double phi = 0;
    for (int d = 0; d < DIM; d++) phi += 2 * std::numbers::pi * double(kk[d]) * nn[d] / shape[d];
    return C{std::cos(phi), -std::sin(phi)};
  };
  std::vector<double> Kref(size_t(ndof) * ndof, 0.0);
  for (int cj = 0; cj < DIM; cj++)
    for (int nj = 0; nj < M; nj++)          // column j = (cj, nj): u = e_j
      for (int ci = 0; ci < DIM; ci++)
        for (int ni = 0; ni < M; ni++) {
          C acc{};
          for (int k = 0; k < M; k++)        // F[ni] = |h|/|N| sum_k exp(+i phi[ni,k]) K^[k] exp(-i phi[nj,k])
            acc += std::conj(phase(ni, k)) * Khat[size_t(k) * DIM * DIM + DIM * ci + cj] * phase(nj, k);
          Kref[size_t(ci * M + ni) * ndof + (cj * M + nj)] = (acc * (cell_volume / M)).real();
        }

  bri17::RealSpaceOperator<DIM> op{hooke};
  C *du, *dF;
  double *dur, *dFr;
  cudaMalloc(&du, sizeof(C) * ndof); cudaMalloc(&dF, sizeof(C) * ndof);
  cudaMalloc(&dur, sizeof(double) * ndof); cudaMalloc(&dFr, sizeof(double) * ndof);
  std::vector<C> u(ndof), F(ndof);
  std::vector<double> ur(ndof), Fr(ndof), K(size_t(ndof) * ndof);
  double worst = 0, worst_imag = 0, worst_r2c = 0, scale = 0;
  for (int j = 0; j < ndof; j++) {
    u.assign(ndof, C{}); u[j] = 1.0;                         // tests/test_bri17.cpp:135-136
    ur.assign(ndof, 0.0); ur[j] = 1.0;
    cudaMemcpy(du, u.data(), sizeof(C) * ndof, cudaMemcpyHostToDevice);
    cudaMemcpy(dur, ur.data(), sizeof(double) * ndof, cudaMemcpyHostToDevice);
    op.apply(du, dF);
    op.apply(dur, dFr);
    cudaMemcpy(F.data(), dF, sizeof(C) * ndof, cudaMemcpyDeviceToHost);
    cudaMemcpy(Fr.data(), dFr, sizeof(double) * ndof, cudaMemcpyDeviceToHost);
    for (int i = 0; i < ndof; i++) {
      const double e = Kref[size_t(i) * ndof + j];
      K[size_t(i) * ndof + j] = F[i].real();
      scale = std::max(scale, std::abs(e));
      worst = std::max(worst, std::abs(F[i].real() - e) - 1e-15 * std::abs(e));   // rtol 1e-15 (+ atol below)
      worst_imag = std::max(worst_imag, std::abs(F[i].imag()));                   // :140-144
      worst_r2c = std::max(worst_r2c, std::abs(Fr[i] - e));
    }
  }
  double asym = 0;
  for (int i = 0; i < ndof; i++)
    for (int j = 0; j < ndof; j++) asym = std::max(asym, std::abs(K[size_t(i) * ndof + j] - K[size_t(j) * ndof + i]));

  // CG: b = K x_true with zero-mean x_true
  std::vector<double> xt(ndof), b(ndof, 0.0), x(ndof);
  for (int c = 0; c < DIM; c++) {
    double mean = 0;
    for (int n = 0; n < M; n++) { xt[c * M + n] = std::sin(1.0 + 0.7 * n + c); mean += xt[c * M + n]; }
    for (int n = 0; n < M; n++) xt[c * M + n] -= mean / M;
  }
  for (int i = 0; i < ndof; i++)
    for (int j = 0; j < ndof; j++) b[i] += Kref[size_t(i) * ndof + j] * xt[j];
  cudaMemcpy(dur, b.data(), sizeof(double) * ndof, cudaMemcpyHostToDevice);
  double res = 0;
  const int iters = op.solve(dur, dFr, 1e-12, 2000, &res);
  cudaMemcpy(x.data(), dFr, sizeof(double) * ndof, cudaMemcpyDeviceToHost);
  double cg_err = 0;
  for (int i = 0; i < ndof; i++) cg_err = std::max(cg_err, std::abs(x[i] - xt[i]));
  cudaFree(du); cudaFree(dF); cudaFree(dur); cudaFree(dFr);
  // the host reference is a naive O(N^2) DFT (error ~ N * eps * |K|), hence looser bounds than the
  // reference's 1e-15*|e| + 1e-14, which tests/test_gpu_realspace.py applies with the Maxima matrices
  const bool ok = worst <= 5e-13 && worst_imag <= 1e-13 && worst_r2c <= 5e-13 && asym <= 1e-13 && cg_err <= 1e-8;
  std::printf("dim %d dofs %d: max|K-Kref| excess %.2e imag %.2e r2c %.2e asym %.2e (|K| %.1f) cg %d it res %.1e err %.1e %s\n",
              DIM, ndof, worst, worst_imag, worst_r2c, asym, scale, iters, res, cg_err, ok ? "OK" : "FAIL");
  return ok ? 0 : 1;
}

int main() {
  int rc = run<2>({3, 4});
  rc |= run<3>({3, 4, 5});
  return rc;
}
