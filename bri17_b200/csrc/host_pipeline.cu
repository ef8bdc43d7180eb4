// host_pipeline.cu -- modal stiffness apply on HOST buffers.
//
// The reference's harness owns host arrays (fftw_malloc'ed, tests/test_bri17.cpp:30-41)
// and walks them in place; a drop-in call therefore starts and ends in host
// memory.  The block is cut along its slowest axis into chunks; each chunk is
// copied host->device, processed by the apply kernel and copied back, on
// `host_streams` rotating streams, so the H2D copy engine, the SMs and the D2H
// copy engine all stay busy.  The path is PCIe-bound (96 B/mode cross the
// link in 3-D); the kernel itself is ~100x faster than the link.
#include <algorithm>

#include "internal.h"

namespace bri17b200 {

void free_host_stages(bri17_plan *p) {
  for (auto &s : p->stages) {
    if (s.in) cudaFree(s.in);
    if (s.out) cudaFree(s.out);
    if (s.stream) cudaStreamDestroy(s.stream);
  }
  p->stages.clear();
  p->stage_bytes = 0;
}

static int ensure_stages(bri17_plan *p, int64_t bytes) {
  if (int(p->stages.size()) == p->host_streams && p->stage_bytes >= bytes && p->stages[0].in) return BRI17_OK;
  free_host_stages(p);
  p->stages.resize(p->host_streams);
  for (auto &s : p->stages) {
    BRI17_CUDA_TRY(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    BRI17_CUDA_TRY(cudaMalloc(&s.in, bytes));
    BRI17_CUDA_TRY(cudaMalloc(&s.out, bytes));
  }
  p->stage_bytes = bytes;
  return BRI17_OK;
}

// Zero-copy variant (option "host_zero_copy"): when both buffers are page-locked and mapped,
// the apply kernel reads u^ and writes f^ directly in host memory over PCIe, one launch, no
// staging buffers.  Kept as an option: measured against the chunked pipeline in
// profiles/r01_measurements.md.
static int apply_host_zero_copy(bri17_plan *p, const Block &b, const void *u_host, void *f_host,
                                int64_t comp_stride, double out_scale, bool *done) {
  *done = false;
  cudaPointerAttributes au{}, af{};
  if (cudaPointerGetAttributes(&au, u_host) != cudaSuccess ||
      cudaPointerGetAttributes(&af, f_host) != cudaSuccess) {
    cudaGetLastError();
    return BRI17_OK;
  }
  if (au.type != cudaMemoryTypeHost || af.type != cudaMemoryTypeHost || !au.devicePointer || !af.devicePointer)
    return BRI17_OK;  // pageable memory: use the staged pipeline
  if (p->stages.empty()) {
    p->stages.resize(1);
    BRI17_CUDA_TRY(cudaStreamCreateWithFlags(&p->stages[0].stream, cudaStreamNonBlocking));
  }
  cudaStream_t st = p->stages[0].stream;
  int rc = launch_apply(p, b, au.devicePointer, af.devicePointer, comp_stride, comp_stride, out_scale, st);
  if (rc) return rc;
  BRI17_CUDA_TRY(cudaStreamSynchronize(st));
  *done = true;
  return BRI17_OK;
}

int apply_host(bri17_plan *p, const Block &b, const void *u_host, void *f_host,
               int64_t comp_stride, double out_scale) {
  const int dim = b.dim;
  if (p->host_zero_copy) {
    bool done = false;
    int zrc = apply_host_zero_copy(p, b, u_host, f_host, comp_stride, out_scale, &done);
    if (zrc || done) return zrc;
  }
  const int64_t plane = b.modes / b.n[0];  // modes per index of the slowest axis
  int64_t rows = p->host_chunk_rows;
  if (rows <= 0) {  // default: ~32 MiB per component per chunk
    rows = std::max<int64_t>(1, (int64_t(32) << 20) / (plane * 16));
  }
  rows = std::min<int64_t>(rows, b.n[0]);
  const int64_t chunk_modes = rows * plane;
  int rc = ensure_stages(p, chunk_modes * 16 * dim);
  if (rc) return rc;

  const char *src = static_cast<const char *>(u_host);
  char *dst = static_cast<char *>(f_host);
  int si = 0;
  for (int64_t a0 = 0; a0 < b.n[0]; a0 += rows, si = (si + 1) % p->host_streams) {
    auto &s = p->stages[si];
    const int64_t r = std::min<int64_t>(rows, b.n[0] - a0);
    const int64_t modes = r * plane;
    Block cb = b;
    cb.kb[0] = b.kb[0] + int(a0);
    cb.n[0] = int(r);
    cb.modes = modes;
    for (int c = 0; c < dim; c++)
      BRI17_CUDA_TRY(cudaMemcpyAsync(static_cast<char *>(s.in) + c * modes * 16,
                                     src + (c * comp_stride + a0 * plane) * 16, modes * 16,
                                     cudaMemcpyHostToDevice, s.stream));
    rc = launch_apply(p, cb, s.in, s.out, modes, modes, out_scale, s.stream);
    if (rc) return rc;
    for (int c = 0; c < dim; c++)
      BRI17_CUDA_TRY(cudaMemcpyAsync(dst + (c * comp_stride + a0 * plane) * 16,
                                     static_cast<char *>(s.out) + c * modes * 16, modes * 16,
                                     cudaMemcpyDeviceToHost, s.stream));
  }
  for (auto &s : p->stages) BRI17_CUDA_TRY(cudaStreamSynchronize(s.stream));
  return BRI17_OK;
}

}  // namespace bri17b200
