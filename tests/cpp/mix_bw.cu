// mix_bw.cu -- what the HBM of a B200 sustains for a given READ : WRITE mix of planar streams.
// The secondary kernels of bri17_b200 are not 1 : 1 copies: strain recovery reads 3 and writes 6
// complex planes per mode, eigenstress -> displacement reads 6 and writes 3, the K^ / B^ field
// writers only write.  This kernel moves NR input planes and NW output planes with exactly the
// access pattern of those kernels (persistent grid of 296 CTAs x 256 threads, 2 x 128-bit words per
// plane and thread, streaming ld/st) and no arithmetic worth mentioning, so its GB/s is the ceiling
// of each mix: the denominator to judge those kernels by (profiles/r02_measurements.md).
#include <cuda_runtime.h>

#include <cstdio>

template <int NR, int NW>
__global__ void __launch_bounds__(256, 2) mix_kernel(const double2 *in, double2 *out, long long n) {
  constexpr int VEC = 2;
  for (long long base = (long long)blockIdx.x * 256 * VEC; base < n; base += (long long)gridDim.x * 256 * VEC) {
    double2 v[VEC][NR > 0 ? NR : 1];
    double2 acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; j++) {
      acc[j] = make_double2(double(base), 1.0);
#pragma unroll
      for (int c = 0; c < NR; c++) v[j][c] = __ldcs(in + c * n + base + j * 256 + threadIdx.x);
    }
#pragma unroll
    for (int j = 0; j < VEC; j++)
#pragma unroll
      for (int c = 0; c < NR; c++) { acc[j].x += v[j][c].x; acc[j].y += v[j][c].y; }
#pragma unroll
    for (int j = 0; j < VEC; j++) {
#pragma unroll
      for (int c = 0; c < NW; c++) __stcs(out + c * n + base + j * 256 + threadIdx.x, make_double2(acc[j].x + c, acc[j].y));
      if (NW == 0 && acc[j].x == 1.2345e300) out[0] = acc[j];  // keep the loads alive
    }
  }
}

template <int NR, int NW>
static void run(const double2 *in, double2 *out, long long n, bool &first) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int it = 0; it < 6; it++) {
    if (it == 1) cudaEventRecord(e0);
    mix_kernel<NR, NW><<<296, 256>>>(in, out, n);
  }
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= 5;
  printf("%s{\"read_planes\": %d, \"write_planes\": %d, \"ms\": %.4f, \"gbs\": %.0f}", first ? "" : ", ", NR, NW, ms,
         16.0 * n * (NR + NW) / 1e9 / (ms * 1e-3));
  first = false;
}

int main() {
  const long long n = 1LL << 27;  // 512^3 modes, 2 GiB per plane
  double2 *in, *out;
  if (cudaMalloc(&in, size_t(n) * 16 * 6) != cudaSuccess || cudaMalloc(&out, size_t(n) * 16 * 9) != cudaSuccess) return 1;
  cudaMemset(in, 0, size_t(n) * 16 * 6);
  bool first = true;
  printf("{\"modes\": %lld, \"results\": [", n);
  run<3, 3>(in, out, n, first);  // modal stiffness apply, K^-1
  run<3, 6>(in, out, n, first);  // strain recovery
  run<6, 3>(in, out, n, first);  // eigenstress -> displacement
  run<6, 6>(in, out, n, first);  // eigenstress -> opposite strain
  run<0, 3>(in, out, n, first);  // B^ field (write only)
  run<0, 9>(in, out, n, first);  // K^ field (write only)
  run<3, 0>(in, out, n, first);  // read only
  printf("]}\n");
  return 0;
}
