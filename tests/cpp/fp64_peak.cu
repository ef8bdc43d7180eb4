// DFMA micro-benchmark: measures the vector FP64 peak of the device, the
// compute-side denominator of the roofline (SURVEY.md section 8d asks for a
// measured figure: MEASURED_PEAKS.json has none).  Prints TFLOP/s (2 flop per DFMA).
#include <cuda_runtime.h>

#include <cstdio>

__global__ void __launch_bounds__(256) dfma_kernel(double *out, double a, double b, int iters) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = sms * 8, block = 256, iters = 1 << 14;
  double *out;
  cudaMalloc(&out, sizeof(double) * grid * block);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    dfma_kernel<<<grid, block>>>(out, 0.999999, 1e-9, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double tflops = 2.0 * 8 * double(iters) * grid * block / (ms * 1e-3) / 1e12;
    if (tflops > best) best = tflops;
  }
  std::printf("{\"fp64_dfma_tflops\": %.2f, \"sms\": %d}\n", best, sms);
  return cudaGetLastError() != cudaSuccess;
}
