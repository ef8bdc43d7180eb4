// realspace.cu -- libbri17_b200_rs.so: real-space operator F = (|h|/|N|) iDFT(K^ DFT(u))
// (tests/test_bri17.cpp:56-107 of the reference) on 1..16 GPUs, and CG on top of it.
//
// Local transforms: cuFFT Z2Z (the reference uses FFTW c2c on the planar
// component blocks, tests/test_bri17.cpp:117-127).  Distribution: slab over
// axis 0 in real space, slab over axis 1 in Fourier space; one exchange per
// direction.  The modal operator (libbri17_b200.so) runs on the Fourier-side
// block [dim][N0][k1 slab][N2] with k_begin = {0, k1_begin, 0}: no transpose
// back is needed before applying K^.
//
// Exchange = "segment copy" kernel (slab_copy_kernel) + transport:
//   mode 0: pack into per-peer contiguous pieces, NCCL grouped send/recv,
//           unpack on the way back (2 extra HBM passes per apply);
//   mode 1: the same kernel stores straight into the peers' buffers through
//           CUDA-IPC mappings (NVLink peer memory): the transposition IS the
//           transfer, no pack/unpack pass, NCCL only as a stream-ordered barrier.
#include <cuda_runtime.h>
#include <cufft.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "bri17_b200_realspace.h"
#include "internal.h"

// libbri17_b200.so keeps its internals hidden: this library has its own copy of the
// error helpers, which forward to the exported thread-local message of the core library.
namespace bri17b200 {
void set_error(const std::string &msg) { bri17_set_last_error(msg.c_str()); }
int fail(int code, const std::string &msg) {
  bri17_set_last_error(msg.c_str());
  return code;
}
}  // namespace bri17b200
using bri17b200::fail;

#define RS_CUFFT_TRY(expr)                                                                  \
  do {                                                                                      \
    cufftResult _r = (expr);                                                                \
    if (_r != CUFFT_SUCCESS)                                                                \
      return fail(BRI17_ERR_CUDA, std::string(#expr) + ": cuFFT error " + std::to_string(int(_r))); \
  } while (0)
#define RS_NCCL_TRY(expr)                                                                   \
  do {                                                                                      \
    ncclResult_t _r = (expr);                                                               \
    if (_r != ncclSuccess)                                                                  \
      return fail(BRI17_ERR_NCCL, std::string(#expr) + ": " + ncclGetErrorString(_r));      \
  } while (0)
#define RS_TRY(expr)            \
  do {                          \
    int _rc = (expr);           \
    if (_rc != BRI17_OK) return _rc; \
  } while (0)

namespace {

constexpr int MAX_RANKS = 16;
constexpr int NUM_EVENTS = 8;

// One family of equal-length row pieces to move: for c < ncomp, a < rows:
//   dst[c*dst_cs + a*dst_rs + (0..len)] = scale * src[c*src_cs + a*src_rs + (0..len)]
struct CopySeg {
  const double2 *src;
  double2 *dst;
  long long src_cs, src_rs, dst_cs, dst_rs;
  int rows;
  int len;
};
struct CopyPlan {
  CopySeg seg[MAX_RANKS];
  int nseg;
  int ncomp;
  int parts;  // CTAs per row piece
  double scale;
  long long ctas_per_seg;  // grid = nseg * ctas_per_seg, segment index fastest
};

// Streams contiguous row pieces with 128-bit accesses; the destination may be
// local HBM or a peer GPU's memory (NVLink stores).  Consecutive CTAs belong to
// DIFFERENT segments (= destination GPUs) and every rank starts its segment
// list at its right-hand neighbour, so at any instant each GPU spreads its
// stores over all peers and receives from all peers: an all-to-all that walks
// the peers in lock step would serialise on one receiver's NVLink ingress.
__global__ void __launch_bounds__(256) slab_copy_kernel(const CopyPlan cp, int fence_system) {
  const long long cta = blockIdx.x;
  const int s = int(cta % cp.nseg);
  const CopySeg &g = cp.seg[s];
  long long local = cta / cp.nseg;
  if (local >= (long long)cp.ncomp * g.rows * cp.parts) return;
  const int part = int(local % cp.parts);
  local /= cp.parts;
  const int a = int(local % g.rows);
  const int c = int(local / g.rows);
  const int per = ((g.len + cp.parts - 1) / cp.parts + 31) & ~31;
  const int begin = part * per;
  const int end = min(g.len, begin + per);
  const double2 *src = g.src + c * g.src_cs + a * g.src_rs;
  double2 *dst = g.dst + c * g.dst_cs + a * g.dst_rs;
  const bool scaled = cp.scale != 1.0;
  int i = begin + threadIdx.x;
  for (; i + 3 * 256 < end; i += 4 * 256) {
    double2 v0 = __ldcs(src + i), v1 = __ldcs(src + i + 256), v2 = __ldcs(src + i + 512),
            v3 = __ldcs(src + i + 768);
    if (scaled) {
      v0.x *= cp.scale; v0.y *= cp.scale; v1.x *= cp.scale; v1.y *= cp.scale;
      v2.x *= cp.scale; v2.y *= cp.scale; v3.x *= cp.scale; v3.y *= cp.scale;
    }
    dst[i] = v0; dst[i + 256] = v1; dst[i + 512] = v2; dst[i + 768] = v3;
  }
  for (; i < end; i += 256) {
    double2 v = __ldcs(src + i);
    if (scaled) { v.x *= cp.scale; v.y *= cp.scale; }
    dst[i] = v;
  }
  if (fence_system) __threadfence_system();
}

__global__ void scale_kernel(double2 *x, long long n, double scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    double2 v = x[i];
    v.x *= scale; v.y *= scale;
    x[i] = v;
  }
}

// ---- CG vector kernels (K5): deterministic two-stage reductions, device scalars ----
constexpr int RED_CTAS = 1184;  // 148 SMs x 8
constexpr int RED_THREADS = 256;

__device__ __forceinline__ double block_sum(double v) {
  __shared__ double sh[RED_THREADS / 32];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.;
  if (threadIdx.x < RED_THREADS / 32) t = sh[threadIdx.x];
  if (threadIdx.x < 32)
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  __syncthreads();
  return t;  // valid in thread 0
}

// partial[b] = sum over this CTA's elements of Re(x conj(y))
__global__ void __launch_bounds__(RED_THREADS) cg_dot_kernel(const double2 *x, const double2 *y,
                                                              long long n, double *partial) {
  double acc = 0.;
  for (long long i = blockIdx.x * (long long)RED_THREADS + threadIdx.x; i < n;
       i += (long long)gridDim.x * RED_THREADS) {
    const double2 a = x[i], b = y[i];
    acc += a.x * b.x + a.y * b.y;
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

__global__ void __launch_bounds__(RED_THREADS) cg_finish_kernel(const double *partial, int n,
                                                                 double *out) {
  double acc = 0.;
  for (int i = threadIdx.x; i < n; i += RED_THREADS) acc += partial[i];
  acc = block_sum(acc);
  if (threadIdx.x == 0) *out = acc;
}

// alpha = rr/pAp;  x += alpha p;  r -= alpha Ap;  partial <r,r>   (fused axpy + axpy + dot)
__global__ void __launch_bounds__(RED_THREADS) cg_update_kernel(double2 *x, double2 *r, const double2 *p,
                                                                 const double2 *Ap, long long n,
                                                                 const double *rr, const double *pAp,
                                                                 double *partial) {
  const double alpha = *rr / *pAp;
  double acc = 0.;
  for (long long i = blockIdx.x * (long long)RED_THREADS + threadIdx.x; i < n;
       i += (long long)gridDim.x * RED_THREADS) {
    const double2 pi = p[i], ai = Ap[i];
    double2 xi = x[i], ri = r[i];
    xi.x += alpha * pi.x; xi.y += alpha * pi.y;
    ri.x -= alpha * ai.x; ri.y -= alpha * ai.y;
    x[i] = xi; r[i] = ri;
    acc += ri.x * ri.x + ri.y * ri.y;
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

// beta = rr_new/rr;  p = r + beta p
__global__ void __launch_bounds__(RED_THREADS) cg_direction_kernel(double2 *p, const double2 *r, long long n,
                                                                    const double *rr_new, const double *rr) {
  const double beta = *rr_new / *rr;
  for (long long i = blockIdx.x * (long long)RED_THREADS + threadIdx.x; i < n;
       i += (long long)gridDim.x * RED_THREADS) {
    const double2 ri = r[i];
    double2 pi = p[i];
    pi.x = ri.x + beta * pi.x; pi.y = ri.y + beta * pi.y;
    p[i] = pi;
  }
}

}  // namespace

struct bri17_rs_plan {
  int dim = 0, shape[3] = {1, 1, 1};
  double L[3] = {1, 1, 1};
  int device = 0, rank = 0, nranks = 1, mode = 0;
  int N2e = 1;                 // trailing extent (N2 in 3-D, 1 in 2-D)
  int n0_beg[MAX_RANKS + 1], k1_beg[MAX_RANKS + 1];
  int n0_loc = 0, n1_loc = 0;
  int64_t real_count = 0, fourier_count = 0;
  double correction = 1.0;     // |h|/|N|, tests/test_bri17.cpp:93-98
  bri17_plan *modal = nullptr;
  cufftHandle fft_local = 0, fft_axis0 = 0;
  bool have_local = false, have_axis0 = false;
  ncclComm_t comm = nullptr;
  double2 *W = nullptr, *W2 = nullptr;  // exchange buffers, dim components each
  double2 *peerW[MAX_RANKS] = {}, *peerW2[MAX_RANKS] = {};
  double *barrier_word = nullptr;
  cudaEvent_t ev[NUM_EVENTS + 1] = {};
  bool timings_valid = false;
  // CG work space
  double2 *cg_r = nullptr, *cg_p = nullptr, *cg_Ap = nullptr;
  double *cg_partial = nullptr, *cg_scalars = nullptr;
};

namespace {

int launch_copy(CopyPlan &cp, int fence, cudaStream_t st) {
  long long per_seg = 0;
  for (int s = 0; s < cp.nseg; s++)
    per_seg = std::max(per_seg, (long long)cp.ncomp * cp.seg[s].rows * cp.parts);
  cp.ctas_per_seg = per_seg;
  const long long total = per_seg * cp.nseg;
  if (total == 0) return BRI17_OK;
  if (total > 0x7fffffffLL) return fail(BRI17_ERR_UNSUPPORTED, "exchange grid too large");
  slab_copy_kernel<<<(unsigned)total, 256, 0, st>>>(cp, fence);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(BRI17_ERR_CUDA, std::string("slab_copy launch: ") + cudaGetErrorString(e));
  return BRI17_OK;
}

int choose_parts(int len, long long rows_total) {
  // enough CTAs to fill the chip, at least 2048 elements (32 KiB) per CTA
  int parts = 1;
  while (parts < 64 && rows_total * parts < 148 * 8 && len / (parts * 2) >= 2048) parts *= 2;
  return parts;
}

// cross-GPU, stream-ordered barrier (and memory fence) = 1-element all-reduce
int stream_barrier(bri17_rs_plan *p, cudaStream_t st) {
  if (p->nranks == 1) return BRI17_OK;
  RS_NCCL_TRY(ncclAllReduce(p->barrier_word, p->barrier_word, 1, ncclDouble, ncclSum, p->comm, st));
  return BRI17_OK;
}

// Forward exchange: real-space layout T[c][a][b][k2] -> Fourier-side layout X[c][n0][b_loc][k2].
// `S` is the packed send buffer (mode 0).  ncomp <= dim.
int exchange_forward(bri17_rs_plan *p, const double2 *T, double2 *X, double2 *S, int ncomp,
                     cudaStream_t st) {
  const int P = p->nranks, r = p->rank, N0 = p->shape[0], N1 = p->shape[1], N2e = p->N2e;
  CopyPlan cp{};
  cp.ncomp = ncomp;
  cp.scale = 1.0;
  std::vector<long long> off(P + 1, 0);  // packed offsets (elements)
  for (int q = 0; q < P; q++)
    off[q + 1] = off[q] + (long long)ncomp * p->n0_loc * (p->k1_beg[q + 1] - p->k1_beg[q]) * N2e;
  int maxlen = 0;
  for (int i = 1; i <= P; i++) {
    const int q = (r + i) % P;
    const int n1q = p->k1_beg[q + 1] - p->k1_beg[q];
    if (n1q == 0 || p->n0_loc == 0) continue;
    CopySeg &g = cp.seg[cp.nseg++];
    g.src = T + (long long)p->k1_beg[q] * N2e;
    g.src_cs = (long long)p->n0_loc * N1 * N2e;
    g.src_rs = (long long)N1 * N2e;
    g.rows = p->n0_loc;
    g.len = n1q * N2e;
    maxlen = std::max(maxlen, g.len);
    const bool direct = (q == r) || p->mode == 1;  // store at the final position
    if (direct) {
      double2 *base = (q == r) ? X : p->peerW[q];
      g.dst = base + (long long)p->n0_beg[r] * n1q * N2e;
      g.dst_cs = (long long)N0 * n1q * N2e;
      g.dst_rs = (long long)n1q * N2e;
    } else {
      g.dst = S + off[q];
      g.dst_cs = (long long)p->n0_loc * n1q * N2e;
      g.dst_rs = (long long)n1q * N2e;
    }
  }
  cp.parts = choose_parts(maxlen, (long long)ncomp * p->n0_loc * P);
  if (p->mode == 1) RS_TRY(stream_barrier(p, st));  // peers' buffers are free to overwrite
  RS_TRY(launch_copy(cp, p->mode == 1, st));
  if (P == 1) return BRI17_OK;
  if (p->mode == 1) return stream_barrier(p, st);   // everybody's stores have landed
  RS_NCCL_TRY(ncclGroupStart());
  for (int q = 0; q < P; q++) {
    if (q == r) continue;
    const long long n1q = p->k1_beg[q + 1] - p->k1_beg[q], n0q = p->n0_beg[q + 1] - p->n0_beg[q];
    for (int c = 0; c < ncomp; c++) {
      const long long scount = (long long)p->n0_loc * n1q * N2e, rcount = n0q * p->n1_loc * N2e;
      if (scount) RS_NCCL_TRY(ncclSend(S + off[q] + c * scount, size_t(2 * scount), ncclDouble, q, p->comm, st));
      if (rcount)
        RS_NCCL_TRY(ncclRecv(X + ((long long)c * N0 + p->n0_beg[q]) * p->n1_loc * N2e, size_t(2 * rcount),
                             ncclDouble, q, p->comm, st));
    }
  }
  RS_NCCL_TRY(ncclGroupEnd());
  return BRI17_OK;
}

// Backward exchange: Fourier-side X[c][n0][b_loc][k2] -> real-space layout D[c][a][b][k2] (times scale).
// mode 0: R = packed receive buffer, D written by the unpack; mode 1: peers store into our W2 (= D).
int exchange_backward(bri17_rs_plan *p, const double2 *X, double2 *D, double2 *R, int ncomp, double scale,
                      cudaStream_t st) {
  const int P = p->nranks, r = p->rank, N0 = p->shape[0], N1 = p->shape[1], N2e = p->N2e;
  if (p->mode == 1 || P == 1) {
    // every row (c, n0) goes, whole, to the owner of n0, at its final position
    CopyPlan cp{};
    cp.ncomp = ncomp;
    cp.scale = scale;
    for (int i = 1; i <= P; i++) {
      const int q = (r + i) % P;
      const int n0q = p->n0_beg[q + 1] - p->n0_beg[q];
      if (n0q == 0 || p->n1_loc == 0) continue;
      CopySeg &g = cp.seg[cp.nseg++];
      g.src = X + (long long)p->n0_beg[q] * p->n1_loc * N2e;
      g.src_cs = (long long)N0 * p->n1_loc * N2e;
      g.src_rs = (long long)p->n1_loc * N2e;
      g.rows = n0q;
      g.len = p->n1_loc * N2e;
      double2 *base = (q == r) ? D : p->peerW2[q];
      g.dst = base + (long long)p->k1_beg[r] * N2e;
      g.dst_cs = (long long)n0q * N1 * N2e;
      g.dst_rs = (long long)N1 * N2e;
    }
    cp.parts = choose_parts(p->n1_loc * N2e, (long long)ncomp * N0);
    RS_TRY(stream_barrier(p, st));
    RS_TRY(launch_copy(cp, P > 1, st));
    return stream_barrier(p, st);
  }
  std::vector<long long> off(P + 1, 0);
  for (int q = 0; q < P; q++)
    off[q + 1] = off[q] + (long long)ncomp * p->n0_loc * (p->k1_beg[q + 1] - p->k1_beg[q]) * N2e;
  RS_NCCL_TRY(ncclGroupStart());
  for (int q = 0; q < P; q++) {
    if (q == r) continue;
    const long long n1q = p->k1_beg[q + 1] - p->k1_beg[q], n0q = p->n0_beg[q + 1] - p->n0_beg[q];
    for (int c = 0; c < ncomp; c++) {
      const long long scount = n0q * p->n1_loc * N2e, rcount = (long long)p->n0_loc * n1q * N2e;
      if (scount)
        RS_NCCL_TRY(ncclSend(X + ((long long)c * N0 + p->n0_beg[q]) * p->n1_loc * N2e, size_t(2 * scount),
                             ncclDouble, q, p->comm, st));
      if (rcount) RS_NCCL_TRY(ncclRecv(R + off[q] + c * rcount, size_t(2 * rcount), ncclDouble, q, p->comm, st));
    }
  }
  RS_NCCL_TRY(ncclGroupEnd());
  CopyPlan cp{};
  cp.ncomp = ncomp;
  cp.scale = scale;
  int maxlen = 0;
  for (int q = 0; q < P; q++) {
    const int n1q = p->k1_beg[q + 1] - p->k1_beg[q];
    if (n1q == 0 || p->n0_loc == 0) continue;
    CopySeg &g = cp.seg[cp.nseg++];
    g.rows = p->n0_loc;
    g.len = n1q * N2e;
    maxlen = std::max(maxlen, g.len);
    if (q == r) {  // own block straight from X
      g.src = X + (long long)p->n0_beg[r] * n1q * N2e;
      g.src_cs = (long long)N0 * n1q * N2e;
      g.src_rs = (long long)n1q * N2e;
    } else {
      g.src = R + off[q];
      g.src_cs = (long long)p->n0_loc * n1q * N2e;
      g.src_rs = (long long)n1q * N2e;
    }
    g.dst = D + (long long)p->k1_beg[q] * N2e;
    g.dst_cs = (long long)p->n0_loc * N1 * N2e;
    g.dst_rs = (long long)N1 * N2e;
  }
  cp.parts = choose_parts(maxlen, (long long)ncomp * p->n0_loc * P);
  return launch_copy(cp, 0, st);
}

int fft_local(bri17_rs_plan *p, const double2 *in, double2 *out, int ncomp, int dir, cudaStream_t st) {
  if (!p->have_local || p->n0_loc == 0) {
    if (in != out && p->real_count)
      BRI17_CUDA_TRY(cudaMemcpyAsync(out, in, sizeof(double2) * ncomp * p->real_count,
                                     cudaMemcpyDeviceToDevice, st));
    return BRI17_OK;
  }
  RS_CUFFT_TRY(cufftSetStream(p->fft_local, st));
  for (int c = 0; c < ncomp; c++)
    RS_CUFFT_TRY(cufftExecZ2Z(p->fft_local, (cufftDoubleComplex *)(in + c * p->real_count),
                              (cufftDoubleComplex *)(out + c * p->real_count), dir));
  return BRI17_OK;
}

int fft_axis0(bri17_rs_plan *p, double2 *x, int ncomp, int dir, cudaStream_t st) {
  if (!p->have_axis0 || p->n1_loc == 0) return BRI17_OK;
  RS_CUFFT_TRY(cufftSetStream(p->fft_axis0, st));
  for (int c = 0; c < ncomp; c++)
    RS_CUFFT_TRY(cufftExecZ2Z(p->fft_axis0, (cufftDoubleComplex *)(x + c * p->fourier_count),
                              (cufftDoubleComplex *)(x + c * p->fourier_count), dir));
  return BRI17_OK;
}

struct DeviceGuard {
  int prev = -1, dev;
  explicit DeviceGuard(int d) : dev(d) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    if (prev != dev && prev >= 0) cudaSetDevice(prev);
  }
};

void mark(bri17_rs_plan *p, int i, cudaStream_t st) { cudaEventRecord(p->ev[i], st); }

}  // namespace

extern "C" {

int bri17_rs_unique_id(void *out128) {
  if (!out128) return fail(BRI17_ERR_INVALID_ARG, "out128 is NULL");
  static_assert(sizeof(ncclUniqueId) == BRI17_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  RS_NCCL_TRY(ncclGetUniqueId(&id));
  std::memcpy(out128, &id, sizeof(id));
  return BRI17_OK;
}

int bri17_rs_plan_destroy(bri17_rs_plan *p) {
  if (!p) return BRI17_OK;
  DeviceGuard guard(p->device);
  cudaDeviceSynchronize();
  for (int q = 0; q < p->nranks; q++) {
    if (q == p->rank) continue;
    if (p->peerW[q]) cudaIpcCloseMemHandle(p->peerW[q]);
    if (p->peerW2[q]) cudaIpcCloseMemHandle(p->peerW2[q]);
  }
  if (p->comm) ncclCommDestroy(p->comm);
  if (p->have_local) cufftDestroy(p->fft_local);
  if (p->have_axis0) cufftDestroy(p->fft_axis0);
  for (void *ptr : {(void *)p->W, (void *)p->W2, (void *)p->barrier_word, (void *)p->cg_r, (void *)p->cg_p,
                    (void *)p->cg_Ap, (void *)p->cg_partial, (void *)p->cg_scalars})
    if (ptr) cudaFree(ptr);
  for (auto &e : p->ev)
    if (e) cudaEventDestroy(e);
  if (p->modal) bri17_plan_destroy(p->modal);
  delete p;
  return BRI17_OK;
}

int bri17_rs_plan_create(bri17_rs_plan **out, int dim, const int *shape, const double *L, double mu,
                         double nu, int device, int rank, int nranks, const void *nccl_unique_id,
                         int exchange_mode) {
  if (!out) return fail(BRI17_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  if (dim != 2 && dim != 3) return fail(BRI17_ERR_INVALID_ARG, "dim must be 2 or 3");
  if (!shape || !L) return fail(BRI17_ERR_INVALID_ARG, "shape/L is NULL");
  if (nranks < 1 || nranks > MAX_RANKS || rank < 0 || rank >= nranks)
    return fail(BRI17_ERR_INVALID_ARG, "bad rank/nranks (at most 16 ranks)");
  if (nranks > 1 && !nccl_unique_id) return fail(BRI17_ERR_INVALID_ARG, "nccl_unique_id is NULL");
  if (exchange_mode != 0 && exchange_mode != 1) return fail(BRI17_ERR_INVALID_ARG, "exchange_mode must be 0 or 1");

  auto *p = new bri17_rs_plan;
  p->dim = dim;
  p->device = device;
  p->rank = rank;
  p->nranks = nranks;
  p->mode = nranks > 1 ? exchange_mode : 0;
  double cell_volume = 1.0;
  int64_t size = 1;
  for (int d = 0; d < dim; d++) {
    if (shape[d] < 1) { delete p; return fail(BRI17_ERR_INVALID_ARG, "shape entries must be >= 1"); }
    p->shape[d] = shape[d];
    p->L[d] = L[d];
    cell_volume *= L[d] / shape[d];  // tests/test_bri17.cpp:96
    size *= shape[d];
  }
  p->correction = cell_volume / double(size);  // :98
  p->N2e = dim == 3 ? shape[2] : 1;
  for (int q = 0; q <= nranks; q++) {
    p->n0_beg[q] = int((int64_t)q * shape[0] / nranks);
    p->k1_beg[q] = int((int64_t)q * shape[1] / nranks);
  }
  p->n0_loc = p->n0_beg[rank + 1] - p->n0_beg[rank];
  p->n1_loc = p->k1_beg[rank + 1] - p->k1_beg[rank];
  p->real_count = (int64_t)p->n0_loc * shape[1] * p->N2e;
  p->fourier_count = (int64_t)shape[0] * p->n1_loc * p->N2e;

  int rc = bri17_plan_create(&p->modal, dim, shape, L, mu, nu, device);
  if (rc) { delete p; return rc; }
  DeviceGuard guard(device);
  auto bail = [&](int code) { bri17_rs_plan_destroy(p); return code; };

  for (auto &e : p->ev)
    if (cudaEventCreate(&e) != cudaSuccess) return bail(fail(BRI17_ERR_CUDA, "cudaEventCreate failed"));

  // local transform over the trailing axes, one plane of the slab per batch entry
  if (p->n0_loc > 0) {
    int n[2] = {shape[1], p->N2e};
    const int frank = dim - 1;
    const long long plane = (long long)shape[1] * p->N2e;
    cufftResult cr = cufftCreate(&p->fft_local);
    size_t ws = 0;
    long long n64[2] = {n[0], n[1]};
    if (cr == CUFFT_SUCCESS)
      cr = cufftMakePlanMany64(p->fft_local, frank, n64, nullptr, 1, plane, nullptr, 1, plane, CUFFT_Z2Z,
                               p->n0_loc, &ws);
    if (cr != CUFFT_SUCCESS) return bail(fail(BRI17_ERR_CUDA, "cuFFT local plan failed: " + std::to_string(int(cr))));
    p->have_local = true;
  }
  // transform along axis 0 on the Fourier-side block: stride = n1_loc*N2e, batch = n1_loc*N2e
  if (p->n1_loc > 0) {
    const long long S = (long long)p->n1_loc * p->N2e;
    long long n64[1] = {shape[0]};
    long long embed[1] = {shape[0]};
    size_t ws = 0;
    cufftResult cr = cufftCreate(&p->fft_axis0);
    if (cr == CUFFT_SUCCESS)
      cr = cufftMakePlanMany64(p->fft_axis0, 1, n64, embed, S, 1, embed, S, 1, CUFFT_Z2Z, S, &ws);
    if (cr != CUFFT_SUCCESS) return bail(fail(BRI17_ERR_CUDA, "cuFFT axis-0 plan failed: " + std::to_string(int(cr))));
    p->have_axis0 = true;
  }

  if (nranks > 1) {
    ncclUniqueId id;
    std::memcpy(&id, nccl_unique_id, sizeof(id));
    ncclResult_t nr = ncclCommInitRank(&p->comm, nranks, id, rank);
    if (nr != ncclSuccess) return bail(fail(BRI17_ERR_NCCL, std::string("ncclCommInitRank: ") + ncclGetErrorString(nr)));
    const size_t cap = sizeof(double2) * dim * size_t(std::max(p->real_count, p->fourier_count));
    if (cudaMalloc(&p->W, std::max<size_t>(cap, 16)) != cudaSuccess ||
        cudaMalloc(&p->W2, std::max<size_t>(cap, 16)) != cudaSuccess ||
        cudaMalloc(&p->barrier_word, 256) != cudaSuccess)
      return bail(fail(BRI17_ERR_CUDA, "exchange buffer allocation failed"));
    cudaMemset(p->barrier_word, 0, 256);
    if (p->mode == 1) {
      // exchange CUDA-IPC handles of W and W2 through NCCL itself
      struct Handles { cudaIpcMemHandle_t w, w2; };
      Handles mine, *dev_all = nullptr;
      std::vector<Handles> all(nranks);
      if (cudaIpcGetMemHandle(&mine.w, p->W) != cudaSuccess || cudaIpcGetMemHandle(&mine.w2, p->W2) != cudaSuccess)
        return bail(fail(BRI17_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(cudaGetLastError())));
      if (cudaMalloc(&dev_all, sizeof(Handles) * nranks) != cudaSuccess)
        return bail(fail(BRI17_ERR_CUDA, "handle buffer allocation failed"));
      cudaMemcpy(dev_all + rank, &mine, sizeof(Handles), cudaMemcpyHostToDevice);
      nr = ncclAllGather(dev_all + rank, dev_all, sizeof(Handles), ncclChar, p->comm, nullptr);
      cudaError_t ce = cudaDeviceSynchronize();
      if (nr == ncclSuccess && ce == cudaSuccess)
        ce = cudaMemcpy(all.data(), dev_all, sizeof(Handles) * nranks, cudaMemcpyDeviceToHost);
      cudaFree(dev_all);
      if (nr != ncclSuccess || ce != cudaSuccess) return bail(fail(BRI17_ERR_NCCL, "IPC handle all-gather failed"));
      for (int q = 0; q < nranks; q++) {
        if (q == rank) { p->peerW[q] = p->W; p->peerW2[q] = p->W2; continue; }
        void *a = nullptr, *b = nullptr;
        if (cudaIpcOpenMemHandle(&a, all[q].w, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
            cudaIpcOpenMemHandle(&b, all[q].w2, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
          return bail(fail(BRI17_ERR_CUDA, std::string("cudaIpcOpenMemHandle (peer ") + std::to_string(q) +
                                               "): " + cudaGetErrorString(cudaGetLastError())));
        p->peerW[q] = static_cast<double2 *>(a);
        p->peerW2[q] = static_cast<double2 *>(b);
      }
    }
  }
  *out = p;
  return BRI17_OK;
}

int bri17_rs_plan_local(const bri17_rs_plan *p, int *n0_begin, int *n0_count, int *k1_begin, int *k1_count) {
  if (!p) return fail(BRI17_ERR_INVALID_ARG, "plan is NULL");
  if (n0_begin) *n0_begin = p->n0_beg[p->rank];
  if (n0_count) *n0_count = p->n0_loc;
  if (k1_begin) *k1_begin = p->k1_beg[p->rank];
  if (k1_count) *k1_count = p->n1_loc;
  return BRI17_OK;
}
int64_t bri17_rs_plan_real_count(const bri17_rs_plan *p) { return p ? p->real_count : -1; }
int64_t bri17_rs_plan_fourier_count(const bri17_rs_plan *p) { return p ? p->fourier_count : -1; }
bri17_plan *bri17_rs_plan_modal(bri17_rs_plan *p) { return p ? p->modal : nullptr; }

int64_t bri17_rs_plan_exchange_bytes(const bri17_rs_plan *p) {
  if (!p) return -1;
  const int64_t others = p->shape[1] - p->n1_loc;
  return int64_t(16) * p->dim * p->n0_loc * others * p->N2e;
}

int bri17_rs_forward_fft_f64(bri17_rs_plan *p, const void *x_dev, void *x_hat_dev, int ncomp, void *stream) {
  if (!p || !x_dev || !x_hat_dev) return fail(BRI17_ERR_INVALID_ARG, "NULL argument");
  if (ncomp < 1) return fail(BRI17_ERR_INVALID_ARG, "ncomp < 1");
  DeviceGuard guard(p->device);
  cudaStream_t st = cudaStream_t(stream);
  const double2 *x = static_cast<const double2 *>(x_dev);
  double2 *xh = static_cast<double2 *>(x_hat_dev);
  if (p->nranks == 1) {
    RS_TRY(fft_local(p, x, xh, ncomp, CUFFT_FORWARD, st));
    return fft_axis0(p, xh, ncomp, CUFFT_FORWARD, st);
  }
  for (int c0 = 0; c0 < ncomp; c0 += p->dim) {  // the exchange buffers hold dim components
    const int nc = std::min(p->dim, ncomp - c0);
    RS_TRY(fft_local(p, x + c0 * p->real_count, p->W2, nc, CUFFT_FORWARD, st));
    double2 *X = xh + c0 * p->fourier_count;
    if (p->mode == 1) {  // peers store into our W: stage through it
      RS_TRY(exchange_forward(p, p->W2, p->W, nullptr, nc, st));
      BRI17_CUDA_TRY(cudaMemcpyAsync(X, p->W, sizeof(double2) * nc * p->fourier_count, cudaMemcpyDeviceToDevice, st));
    } else {
      RS_TRY(exchange_forward(p, p->W2, X, p->W, nc, st));
    }
    RS_TRY(fft_axis0(p, X, nc, CUFFT_FORWARD, st));
  }
  return BRI17_OK;
}

int bri17_rs_inverse_fft_f64(bri17_rs_plan *p, void *x_hat_dev, void *x_dev, int ncomp, double scale, void *stream) {
  if (!p || !x_dev || !x_hat_dev) return fail(BRI17_ERR_INVALID_ARG, "NULL argument");
  if (ncomp < 1) return fail(BRI17_ERR_INVALID_ARG, "ncomp < 1");
  DeviceGuard guard(p->device);
  cudaStream_t st = cudaStream_t(stream);
  double2 *x = static_cast<double2 *>(x_dev);
  double2 *xh = static_cast<double2 *>(x_hat_dev);
  if (p->nranks == 1) {
    RS_TRY(fft_axis0(p, xh, ncomp, CUFFT_INVERSE, st));
    RS_TRY(fft_local(p, xh, x, ncomp, CUFFT_INVERSE, st));
    if (scale != 1.0 && p->real_count)
      scale_kernel<<<1184, 256, 0, st>>>(x, (long long)ncomp * p->real_count, scale);
    return BRI17_OK;
  }
  for (int c0 = 0; c0 < ncomp; c0 += p->dim) {
    const int nc = std::min(p->dim, ncomp - c0);
    double2 *X = xh + c0 * p->fourier_count, *D = x + c0 * p->real_count;
    RS_TRY(fft_axis0(p, X, nc, CUFFT_INVERSE, st));
    if (p->mode == 1) {
      RS_TRY(exchange_backward(p, X, p->W2, nullptr, nc, scale, st));
      RS_TRY(fft_local(p, p->W2, D, nc, CUFFT_INVERSE, st));
    } else {
      RS_TRY(exchange_backward(p, X, D, p->W2, nc, scale, st));
      RS_TRY(fft_local(p, D, D, nc, CUFFT_INVERSE, st));
    }
  }
  return BRI17_OK;
}

int bri17_real_space_apply_f64(bri17_rs_plan *p, const void *u_dev, void *F_dev, void *stream) {
  if (!p || !u_dev || !F_dev) return fail(BRI17_ERR_INVALID_ARG, "NULL argument");
  if (u_dev == F_dev) return fail(BRI17_ERR_INVALID_ARG, "u_dev and F_dev must be distinct (F is scratch)");
  DeviceGuard guard(p->device);
  cudaStream_t st = cudaStream_t(stream);
  const double2 *u = static_cast<const double2 *>(u_dev);
  double2 *F = static_cast<double2 *>(F_dev);
  const int dim = p->dim;
  int kb[3] = {0, p->k1_beg[p->rank], 0};
  int ls[3] = {p->shape[0], p->n1_loc, p->shape[2]};
  p->timings_valid = false;
  mark(p, 0, st);
  RS_TRY(fft_local(p, u, F, dim, CUFFT_FORWARD, st));                       // :57 (axes 1..)
  mark(p, 1, st);
  double2 *X = F;  // Fourier-side block
  if (p->nranks > 1) {
    X = p->W;
    RS_TRY(exchange_forward(p, F, p->W, p->W2, dim, st));
  }
  mark(p, 2, st);
  RS_TRY(fft_axis0(p, X, dim, CUFFT_FORWARD, st));                          // :57 (axis 0)
  mark(p, 3, st);
  if (p->fourier_count)                                                     // :58-92, scale :93-106
    RS_TRY(bri17_modal_stiffness_apply_f64(p->modal, X, X, kb, ls, 0, p->correction, st));
  mark(p, 4, st);
  RS_TRY(fft_axis0(p, X, dim, CUFFT_INVERSE, st));                          // :95
  mark(p, 5, st);
  if (p->nranks > 1) {
    if (p->mode == 1) {
      RS_TRY(exchange_backward(p, X, p->W2, nullptr, dim, 1.0, st));
      mark(p, 6, st);
      RS_TRY(fft_local(p, p->W2, F, dim, CUFFT_INVERSE, st));
    } else {
      RS_TRY(exchange_backward(p, X, F, p->W2, dim, 1.0, st));
      mark(p, 6, st);
      RS_TRY(fft_local(p, F, F, dim, CUFFT_INVERSE, st));
    }
  } else {
    mark(p, 6, st);
    RS_TRY(fft_local(p, F, F, dim, CUFFT_INVERSE, st));
  }
  mark(p, 7, st);
  p->timings_valid = true;
  return BRI17_OK;
}

int bri17_rs_plan_last_timings(bri17_rs_plan *p, double *ms, int n) {
  if (!p || !ms) return fail(BRI17_ERR_INVALID_ARG, "NULL argument");
  if (!p->timings_valid) return fail(BRI17_ERR_INVALID_ARG, "no real-space apply has been timed yet");
  DeviceGuard guard(p->device);
  BRI17_CUDA_TRY(cudaEventSynchronize(p->ev[7]));
  for (int i = 0; i < n && i < 8; i++) {
    float t = 0.f;
    BRI17_CUDA_TRY(cudaEventElapsedTime(&t, p->ev[i < 7 ? i : 0], p->ev[i < 7 ? i + 1 : 7]));
    ms[i] = t;
  }
  return BRI17_OK;
}

int bri17_cg_solve_f64(bri17_rs_plan *p, const void *b_dev, void *x_dev, double rtol, int max_iter,
                       int check_every, int *iterations, double *rel_residual, void *stream) {
  if (!p || !b_dev || !x_dev) return fail(BRI17_ERR_INVALID_ARG, "NULL argument");
  if (max_iter < 0) return fail(BRI17_ERR_INVALID_ARG, "max_iter < 0");
  DeviceGuard guard(p->device);
  cudaStream_t st = cudaStream_t(stream);
  const long long n = (long long)p->dim * p->real_count;
  const size_t bytes = sizeof(double2) * std::max<long long>(n, 1);
  if (!p->cg_r) {
    BRI17_CUDA_TRY(cudaMalloc(&p->cg_r, bytes));
    BRI17_CUDA_TRY(cudaMalloc(&p->cg_p, bytes));
    BRI17_CUDA_TRY(cudaMalloc(&p->cg_Ap, bytes));
    BRI17_CUDA_TRY(cudaMalloc(&p->cg_partial, sizeof(double) * RED_CTAS));
    BRI17_CUDA_TRY(cudaMalloc(&p->cg_scalars, sizeof(double) * 8));
  }
  const double2 *b = static_cast<const double2 *>(b_dev);
  double2 *x = static_cast<double2 *>(x_dev), *r = p->cg_r, *d = p->cg_p, *Ad = p->cg_Ap;
  double *sc = p->cg_scalars;  // [0] rr (even iter) [1] rr (odd iter) [2] pAp [3] bb
  auto reduce_to = [&](double *slot) -> int {
    cg_finish_kernel<<<1, RED_THREADS, 0, st>>>(p->cg_partial, RED_CTAS, slot);
    if (p->nranks > 1) RS_NCCL_TRY(ncclAllReduce(slot, slot, 1, ncclDouble, ncclSum, p->comm, st));
    return BRI17_OK;
  };
  // x = 0, r = b, d = r, rr = <r,r>
  BRI17_CUDA_TRY(cudaMemsetAsync(x, 0, bytes, st));
  BRI17_CUDA_TRY(cudaMemcpyAsync(r, b, bytes, cudaMemcpyDeviceToDevice, st));
  BRI17_CUDA_TRY(cudaMemcpyAsync(d, b, bytes, cudaMemcpyDeviceToDevice, st));
  cg_dot_kernel<<<RED_CTAS, RED_THREADS, 0, st>>>(r, r, n, p->cg_partial);
  RS_TRY(reduce_to(sc + 0));
  BRI17_CUDA_TRY(cudaMemcpyAsync(sc + 3, sc + 0, sizeof(double), cudaMemcpyDeviceToDevice, st));
  double h[4] = {0, 0, 0, 0};
  BRI17_CUDA_TRY(cudaMemcpyAsync(h, sc, sizeof(double) * 4, cudaMemcpyDeviceToHost, st));
  BRI17_CUDA_TRY(cudaStreamSynchronize(st));
  const double bb = h[3];
  int it = 0;
  double rr_host = bb;
  if (bb > 0.) {
    for (; it < max_iter;) {
      double *rr = sc + (it & 1), *rr_new = sc + ((it + 1) & 1);
      RS_TRY(bri17_real_space_apply_f64(p, d, Ad, st));
      cg_dot_kernel<<<RED_CTAS, RED_THREADS, 0, st>>>(d, Ad, n, p->cg_partial);
      RS_TRY(reduce_to(sc + 2));
      cg_update_kernel<<<RED_CTAS, RED_THREADS, 0, st>>>(x, r, d, Ad, n, rr, sc + 2, p->cg_partial);
      RS_TRY(reduce_to(rr_new));
      cg_direction_kernel<<<RED_CTAS, RED_THREADS, 0, st>>>(d, r, n, rr_new, rr);
      it++;
      if (check_every > 0 && (it % check_every == 0 || it == max_iter)) {
        BRI17_CUDA_TRY(cudaMemcpyAsync(&rr_host, rr_new, sizeof(double), cudaMemcpyDeviceToHost, st));
        BRI17_CUDA_TRY(cudaStreamSynchronize(st));
        if (rr_host <= rtol * rtol * bb) break;
      }
    }
    if (check_every <= 0) {
      BRI17_CUDA_TRY(cudaMemcpyAsync(&rr_host, sc + (it & 1), sizeof(double), cudaMemcpyDeviceToHost, st));
      BRI17_CUDA_TRY(cudaStreamSynchronize(st));
    }
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(BRI17_ERR_CUDA, std::string("CG: ") + cudaGetErrorString(e));
  if (iterations) *iterations = it;
  if (rel_residual) *rel_residual = bb > 0. ? std::sqrt(rr_host / bb) : 0.;
  return BRI17_OK;
}

}  // extern "C"
