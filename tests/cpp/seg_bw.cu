// seg_bw.cu -- microbenchmark behind the layout of the fused axis-0 pass:
// how fast can a B200 gather/scatter short row segments at a given row stride?
// Each CTA "tile" reads ROWS segments of SEG bytes, ROW_STRIDE bytes apart (one 16-byte word per
// lane), and writes them back, 8 independent 128-bit loads in flight per thread -- the access
// pattern of stage 0 / the last stage of axis0_fused_kernel.  Prints GB/s (read + write).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>

__global__ void __launch_bounds__(256) seg_copy(const double2 *in, double2 *out, long long row_stride, int seg_words,
                                                int rows, long long tiles, long long tiles_per_block) {
  // thread -> (row group q, word w): w fastest; 8 rows per thread, q + (rows/8)*r
  const int w = threadIdx.x % seg_words, q = threadIdx.x / seg_words;
  const int qn = 256 / seg_words, per = rows / 8;
  for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
    for (int q0 = q; q0 < per; q0 += qn) {
      // tile t: segment t % tiles_per_block of block t / tiles_per_block (a block = `rows` full rows)
      const long long base = (t / tiles_per_block) * rows * row_stride + (t % tiles_per_block) * seg_words +
                             (long long)q0 * row_stride + w;
      double2 v[8];
#pragma unroll
      for (int r = 0; r < 8; r++) v[r] = __ldcs(in + base + (long long)r * per * row_stride);
#pragma unroll
      for (int r = 0; r < 8; r++) __stcs(out + base + (long long)r * per * row_stride, v[r]);
    }
  }
}

int main() {
  const size_t words = size_t(3) << 28;  // 12 GiB per buffer
  double2 *in, *out;
  if (cudaMalloc(&in, words * 16) != cudaSuccess || cudaMalloc(&out, words * 16) != cudaSuccess) return 1;
  cudaMemset(in, 1, words * 16);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  printf("{\"rows\": 512, \"results\": [");
  bool first = true;
  const int rows = 512;
  for (int seg_bytes : {64, 128, 256}) {
    for (long long stride_bytes : {512LL, 8192LL, 1LL << 20, 4LL << 20}) {
      const int seg_words = seg_bytes / 16;
      const long long row_stride = stride_bytes / 16;
      // tiles: adjacent segments along the row; a "super row" holds row_stride/seg_words tiles, then
      // the next block of `rows` rows starts
      const long long tiles_per_block = row_stride / seg_words;
      const long long block_words = (long long)rows * row_stride;
      const long long nblocks = (long long)(words / block_words);
      if (nblocks < 1) continue;
      const long long use_tiles = std::min<long long>(tiles_per_block * nblocks, (1LL << 33) / (rows * seg_bytes));
      for (int grid : {296, 592}) {
        for (int it = 0; it < 3; it++) {
          if (it == 1) cudaEventRecord(e0);
          seg_copy<<<grid, 256>>>(in, out, row_stride, seg_words, rows, use_tiles, tiles_per_block);
        }
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        ms /= 2;
        const double gb = 2.0 * use_tiles * rows * seg_bytes / 1e9;
        printf("%s{\"seg_bytes\": %d, \"row_stride_bytes\": %lld, \"grid\": %d, \"tiles\": %lld, \"ms\": %.4f, \"gbs\": %.0f}",
               first ? "" : ", ", seg_bytes, stride_bytes, grid, use_tiles, ms, gb / (ms * 1e-3));
        first = false;
      }
    }
  }
  printf("]}\n");
  return 0;
}
