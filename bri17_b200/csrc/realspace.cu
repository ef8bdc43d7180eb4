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
//           transfer, no pack/unpack pass.  Cross-GPU synchronisation = release/
//           acquire flags in the same peer-mapped memory (flag_barrier_kernel, a few
//           microseconds; round 1 used a 1-element ncclAllReduce, ~20 us each), and
//           the CG scalars are summed the same way (peer_allreduce_kernel).  NCCL is
//           only used at plan creation (IPC handle all-gather) in this mode.
//
// Axis 0: when N0 is a supported power of two the three passes  FFT(axis 0) -> K^ ->
// inverse FFT(axis 0)  are ONE kernel (axis0_fused.cuh); otherwise cuFFT + the modal kernel.
#include <cuda_runtime.h>
#include <cufft.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "axis0_fused.cuh"
#include "bri17_b200_realspace.h"
#include "internal.h"

namespace bri17b200 {
namespace axis0 {
int launch(Params p, int dim, int sm_count, int max_grid, cudaStream_t st, int *grid_out);
void fill_twiddles(int N0, double2 *tw);
int emulate(const Params &p, int dim, double *dot_out);
}  // namespace axis0
}  // namespace bri17b200

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
constexpr int MAX_XCHUNKS = 4;  // sub-slabs per component in the pipelined apply

// One family of equal-length row pieces to move: for c < ncomp, a < rows, e < len:
//   dst[c*dst_cs + a*dst_rs + off(e, dst_bs)] = scale * src[c*src_cs + a*src_rs + off(e, src_bs)]
// off(e, bs) = (e / inner)*bs + e % inner: a piece is len/inner runs of `inner` contiguous elements,
// `bs` apart (bs = inner: the piece is contiguous).  The k1-major Fourier-side layout
// [c][k1][n0][k2] makes the (c, n0) piece of a rank a set of k1 runs N0*S2e apart.
struct CopySeg {
  const double2 *src;
  double2 *dst;
  long long src_cs, src_rs, dst_cs, dst_rs;
  long long src_bs, dst_bs;
  int rows;
  int len;
  int inner;
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
// The grid is capped (a few CTAs per SM walk the virtual CTA list) so that transform kernels
// of the next/previous component can run beside it when the apply is pipelined.
__global__ void __launch_bounds__(256) slab_copy_kernel(const CopyPlan cp, int fence_system) {
 for (long long cta = blockIdx.x; cta < cp.ctas_per_seg * cp.nseg; cta += gridDim.x) {
  const int s = int(cta % cp.nseg);
  const CopySeg &g = cp.seg[s];
  long long local = cta / cp.nseg;
  if (local >= (long long)cp.ncomp * g.rows * cp.parts) continue;
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
  if (g.src_bs == g.inner && g.dst_bs == g.inner) {  // contiguous on both sides
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
  } else {  // runs of `inner` elements, src_bs / dst_bs apart
    const unsigned inner = unsigned(g.inner);
    for (; i + 3 * 256 < end; i += 4 * 256) {
      long long so[4], dd[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const unsigned e = unsigned(i + j * 256), b = e / inner, k = e - b * inner;
        so[j] = b * g.src_bs + k;
        dd[j] = b * g.dst_bs + k;
      }
      double2 v[4];
#pragma unroll
      for (int j = 0; j < 4; j++) v[j] = __ldcs(src + so[j]);
#pragma unroll
      for (int j = 0; j < 4; j++) {
        if (scaled) { v[j].x *= cp.scale; v[j].y *= cp.scale; }
        dst[dd[j]] = v[j];
      }
    }
    for (; i < end; i += 256) {
      const unsigned e = unsigned(i), b = e / inner, k = e - b * inner;
      double2 v = __ldcs(src + b * g.src_bs + k);
      if (scaled) { v.x *= cp.scale; v.y *= cp.scale; }
      dst[b * g.dst_bs + k] = v;
    }
  }
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

// ---- cross-GPU flags in peer memory (exchange mode 1) -------------------------------------
// Every rank owns one page of 64-bit words at the tail of its exchange buffer W, mapped by all
// peers through the same CUDA-IPC handle as W itself:
//   [FLAG_BAR0 + q]  barrier counter written by rank q, caller's stream
//   [FLAG_BAR1 + q]  same for the exchange stream of the pipelined apply
//   [FLAG_RED + q]   all-reduce counter written by rank q
//   [FLAG_VAL + (parity*MAX_RANKS + q)*RED_MAX + i]   all-reduce values of rank q (doubles)
// Counters only grow; every rank executes the same sequence of barriers / reductions per stream,
// so "counter >= my epoch" is the arrival test.
constexpr int FLAG_BAR0 = 0, FLAG_BAR1 = 16, FLAG_RED = 32, FLAG_VAL = 64, RED_MAX = 8;
constexpr int FLAG_WORDS = FLAG_VAL + 2 * MAX_RANKS * RED_MAX;  // 320 words
constexpr size_t FLAG_BYTES = 4096;
static_assert(FLAG_WORDS * 8 <= FLAG_BYTES, "flag page too small");
constexpr unsigned long long SPIN_TIMEOUT_NS = 60ull * 1000000000ull;

struct PeerFlags {
  unsigned long long *page[MAX_RANKS];  // flag page of every rank (own page at [rank])
  int rank, nranks;
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// Spins until *p >= epoch.  A peer that never arrives (crashed rank) must not hang the GPU:
// after SPIN_TIMEOUT_NS the kernel traps, which surfaces as a CUDA error on the host.
__device__ __forceinline__ void wait_flag(const unsigned long long *p, unsigned long long epoch) {
  const unsigned long long t0 = global_timer_ns();
  while (ld_acquire_sys(p) < epoch) {
    __nanosleep(64);
    if (global_timer_ns() - t0 > SPIN_TIMEOUT_NS) __trap();
  }
}

// Stream-ordered barrier across the ranks: lane q tells rank q "I am at `epoch`" and waits for
// rank q's message.  Everything this GPU wrote before (earlier kernels of the stream, including
// stores into peer memory) is released to the system first; everything peers wrote before their
// arrival is acquired.
__global__ void flag_barrier_kernel(PeerFlags f, int base, unsigned long long epoch) {
  const int q = threadIdx.x;
  if (q >= f.nranks || q == f.rank) return;
  __threadfence_system();
  st_release_sys(f.page[q] + base + f.rank, epoch);
  wait_flag(f.page[f.rank] + base + q, epoch);
  __threadfence_system();
}

// v[0..n) <- sum over ranks of v[0..n), n <= RED_MAX, in rank order on every rank (deterministic
// and identical everywhere).  One warp: lane q pushes this rank's values into rank q's page.
__global__ void peer_allreduce_kernel(PeerFlags f, double *v, int n, unsigned long long epoch) {
  const int q = threadIdx.x;
  const int par = int(epoch & 1);
  if (q < f.nranks) {
    double *dst = reinterpret_cast<double *>(f.page[q] + FLAG_VAL) + (par * MAX_RANKS + f.rank) * RED_MAX;
    for (int i = 0; i < n; i++) dst[i] = v[i];
    if (q != f.rank) {
      st_release_sys(f.page[q] + FLAG_RED + f.rank, epoch);
      wait_flag(f.page[f.rank] + FLAG_RED + q, epoch);
    }
  }
  __syncwarp();
  if (q < n) {
    const double *src = reinterpret_cast<const double *>(f.page[f.rank] + FLAG_VAL) + par * MAX_RANKS * RED_MAX;
    double acc = 0.;
    for (int r = 0; r < f.nranks; r++) acc += reinterpret_cast<const volatile double *>(src)[r * RED_MAX + q];
    v[q] = acc;
  }
}

// ---- CG vector kernels (K5): deterministic two-stage reductions, device scalars ----
// Vectors are plain double arrays (a complex field is its interleaved doubles:
// sum x.y over doubles = sum Re(x conj y)), processed as double2 when possible.
// Per iteration (cg_core): the operator application returns <p, A p> itself (Parseval sum in
// the modal kernel), then
//   cg_residual_kernel:  r -= alpha A p, partial <r, r>          (2 reads + 1 write)
//   cg_direction_kernel: x += alpha p,  p = r + beta p           (3 reads + 2 writes)
// = 8 vector passes; round 1 spent 11 (separate <p, A p> pass, x updated with r).
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

__global__ void __launch_bounds__(RED_THREADS) cg_finish_kernel(const double *partial, int n, double *out) {
  double acc = 0.;
  for (int i = threadIdx.x; i < n; i += RED_THREADS) acc += partial[i];
  acc = block_sum(acc);
  if (threadIdx.x == 0) *out = acc;
}

// Sums of the `nslot` interleaved sub-fields of b: slot = c*inter + t for value (c*count + i)*inter + t
// (complex fields: inter = 2, real and imaginary parts are separate slots).  grid = (CTAs, nslot).
constexpr int MEAN_CTAS = 148;
__global__ void __launch_bounds__(RED_THREADS) cg_slot_sum_kernel(const double *b, long long count, int inter,
                                                                   double *partial) {
  const int slot = blockIdx.y, c = slot / inter, t = slot % inter;
  const double *src = b + (long long)c * count * inter + t;
  double acc = 0.;
  for (long long i = blockIdx.x * (long long)RED_THREADS + threadIdx.x; i < count;
       i += (long long)gridDim.x * RED_THREADS)
    acc += src[i * inter];
  acc = block_sum(acc);
  if (threadIdx.x == 0) partial[slot * MEAN_CTAS + blockIdx.x] = acc;
}
__global__ void __launch_bounds__(RED_THREADS) cg_slot_finish_kernel(const double *partial, double *sums) {
  double acc = 0.;
  for (int i = threadIdx.x; i < MEAN_CTAS; i += RED_THREADS) acc += partial[blockIdx.x * MEAN_CTAS + i];
  acc = block_sum(acc);
  if (threadIdx.x == 0) sums[blockIdx.x] = acc;
}

// r = p = b - mean(slot), x = 0, partial <r, r>: the zero-frequency component of the right-hand
// side is projected out (K^(0) = 0: bri17.hpp:336-339 skips the null frequency, theory.rst:208-212).
__global__ void __launch_bounds__(RED_THREADS) cg_init_kernel(const double *b, double *x, double *r, double *p,
                                                               long long count, int inter, int nslot,
                                                               const double *sums, double inv_total,
                                                               double *partial) {
  const long long n = count * inter * (nslot / inter);
  double acc = 0.;
  for (long long i = blockIdx.x * (long long)RED_THREADS + threadIdx.x; i < n;
       i += (long long)gridDim.x * RED_THREADS) {
    const int slot = int(i / (count * inter)) * inter + int(i % inter);
    const double v = b[i] - sums[slot] * inv_total;
    x[i] = 0.;
    r[i] = v;
    p[i] = v;
    acc += v * v;
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

// Breakdown guards: a zero denominator (r == 0 exactly, or a null direction) freezes the
// iteration instead of producing NaN; the host stops at its next residual check.
__device__ __forceinline__ double safe_ratio(double num, double den) { return den > 0. ? num / den : 0.; }

// alpha = rr/pAp;  r -= alpha Ap;  partial <r,r>
__global__ void __launch_bounds__(RED_THREADS) cg_residual_kernel(double *r, const double *Ap, long long n,
                                                                   const double *rr, const double *pAp,
                                                                   double *partial) {
  const double alpha = safe_ratio(*rr, *pAp);
  double2 *r2 = reinterpret_cast<double2 *>(r);
  const double2 *A2 = reinterpret_cast<const double2 *>(Ap);
  const long long n2 = n >> 1;
  double acc = 0.;
  for (long long i = blockIdx.x * (long long)RED_THREADS + threadIdx.x; i < n2;
       i += (long long)gridDim.x * RED_THREADS) {
    const double2 ai = A2[i];
    double2 ri = r2[i];
    ri.x -= alpha * ai.x; ri.y -= alpha * ai.y;
    r2[i] = ri;
    acc += ri.x * ri.x + ri.y * ri.y;
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const double t = r[n - 1] - alpha * Ap[n - 1];
    r[n - 1] = t;
    acc += t * t;
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

// alpha = rr/pAp, beta = rr_new/rr;  x += alpha p;  p = r + beta p
__global__ void __launch_bounds__(RED_THREADS) cg_direction_kernel(double *x, double *p, const double *r,
                                                                    long long n, const double *rr,
                                                                    const double *pAp, const double *rr_new) {
  const double alpha = safe_ratio(*rr, *pAp), beta = safe_ratio(*rr_new, *rr);
  double2 *x2 = reinterpret_cast<double2 *>(x), *p2 = reinterpret_cast<double2 *>(p);
  const double2 *r2 = reinterpret_cast<const double2 *>(r);
  const long long n2 = n >> 1;
  for (long long i = blockIdx.x * (long long)RED_THREADS + threadIdx.x; i < n2;
       i += (long long)gridDim.x * RED_THREADS) {
    const double2 ri = r2[i];
    double2 pi = p2[i], xi = x2[i];
    xi.x += alpha * pi.x; xi.y += alpha * pi.y;
    pi.x = ri.x + beta * pi.x; pi.y = ri.y + beta * pi.y;
    x2[i] = xi;
    p2[i] = pi;
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    x[n - 1] += alpha * p[n - 1];
    p[n - 1] = r[n - 1] + beta * p[n - 1];
  }
}

}  // namespace

// Batched transform over the trailing axes of a slab, one plane per batch entry, optionally executed
// in CHUNKS of consecutive planes (option "fft_chunk_mib").  cuFFT runs a 2-D transform as two kernels
// (one per axis) with the whole batch between them: over a 512^3 component that is two full passes over
// HBM, each at the copy roofline (0.62-0.78 ms for 4.3 GB).  The idea of the chunks: with a chunk that
// fits the 126 MB L2 the second kernel could find the first one's output there.  MEASURED (512^3, one
// B200, profiles/r02_measurements.md): it does not pay -- 4.8 ms per direction for the whole slab, 7.0 ms
// with 32 MiB chunks, 5.8 ms with 96 MiB chunks (hundreds of 20 us kernels with dependent launches lose
// more than the L2 hits win).  Default: one chunk = the whole slab.
struct BatchFft {
  int rank = 0;
  long long n[2] = {1, 1}, inembed[2] = {1, 1}, onembed[2] = {1, 1};
  bool embed = false;               // advanced layout (inembed/onembed) or the packed default
  long long istride = 1, ostride = 1;
  long long idist = 0, odist = 0;   // elements of the input / output type between consecutive planes
  cufftType type = CUFFT_Z2Z;
  int planes = 0, chunk = 0;        // planes of the slab; planes per cuFFT call in batch_exec
  struct { int batch; cufftHandle h; } cache[8] = {};  // one cuFFT plan per batch size in use
  int ncache = 0;
  bool made = false;
};

// Geometry + cuFFT plans of one spectral layout.  "complex": the reference's c2c
// transforms (tests/test_bri17.cpp:117-127).  "real": r2c/c2r over the trailing
// axes, only the non-redundant half spectrum of the last axis is kept
// (K^(N-k) = K^(k)).
struct Layout {
  bool real = false, ready = false;
  int S1 = 1, S2e = 1;        // spectral extents of axis 1 and of the trailing axis (1 in 2-D)
  int k1_beg[16 + 1] = {};
  int n1_loc = 0;
  int64_t t_count = 0;        // complex elements / component after the local transform [n0_loc][S1][S2e]
  int64_t fourier_count = 0;  // ... of the Fourier-side block [N0][n1_loc][S2e]
  BatchFft fwd_local, inv_local;   // c2c: one plan serves both directions (inv_local unused)
  cufftHandle axis0 = 0;
  bool have_local = false, have_axis0 = false;
  // single GPU, 3-D, fused axis-0 pass: the same local transforms writing / reading the k1-major
  // layout [k1][n0][S2e] through cuFFT's advanced data layout (created on first use)
  BatchFft fwd_local_t, inv_local_t;
  bool have_local_t = false;
};

struct bri17_rs_plan {
  int dim = 0, shape[3] = {1, 1, 1};
  double L[3] = {1, 1, 1};
  double mu = 0, nu = 0;
  int device = 0, rank = 0, nranks = 1, mode = 0, sm_count = 148;
  int N2e = 1;                 // trailing extent in real space (N2 in 3-D, 1 in 2-D)
  int n0_beg[16 + 1];
  int n0_loc = 0;
  int64_t real_count = 0;      // elements / component of the real-space slab [n0_loc][N1][N2e]
  double correction = 1.0;     // |h|/|N|, tests/test_bri17.cpp:93-98
  Layout lc, lr;               // complex (c2c) and real (r2c) layouts
  bri17_plan *modal = nullptr;
  ncclComm_t comm = nullptr;
  cudaStream_t sx = nullptr;                      // exchange stream of the pipelined apply
  cudaEvent_t ev_a[3 * MAX_XCHUNKS] = {}, ev_b[3 * MAX_XCHUNKS] = {};  // per-sub-slab hand-offs st <-> sx
  int xchunks = 0;                                // option "exchange_chunks": sub-slabs per component (0 = by size)
  int pipeline = 1;                               // overlap the exchange of component c with the FFTs of c+-1
  int fft_chunk_planes = 0;                       // same in planes (takes precedence; tests)
  int fft_chunk_mib = 0;                          // > 0: local 2-D transforms run in chunks of planes of this size (BatchFft); measured slower, off
  int copy_ctas = 0;                              // grid cap of slab_copy_kernel; 0 = copy_grid_cap() decides
  double2 *W = nullptr, *W2 = nullptr;  // exchange buffers, dim components of the c2c layout each
  size_t buf_bytes = 0;
  int64_t real_upper = 0;      // element offset of the upper region of W2 used by the real path
  double2 *peerW[16] = {}, *peerW2[16] = {};
  double *barrier_word = nullptr;
  // release/acquire flags in peer memory (mode 1): one page at the tail of every rank's W
  PeerFlags flags{};
  unsigned long long epoch_bar[2] = {0, 0}, epoch_red = 0;
  // fused axis-0 pass (axis0_fused.cuh): own device copies of phi|chi|psi per axis + twiddles
  int fused = 1;               // option "fused_axis0"; used when axis0::supported(shape[0])
  int xt = -1;                 // option "k1_major": k1-major Fourier-side layout with the fused pass (3-D); -1 auto
  double *tabs = nullptr;      // [axis][3][N_axis]
  int64_t tab_off[3] = {0, 0, 0};
  double2 *twiddle = nullptr;
  double *dot_scratch = nullptr;  // one partial per CTA of the pass that also returns <u^, f^>
  int64_t fused_launches = 0;
  cudaEvent_t ev[8 + 1] = {};
  bool timings_valid = false;
  // CG work space (doubles)
  double *cg_r = nullptr, *cg_p = nullptr, *cg_Ap = nullptr;
  size_t cg_len = 0;
  double *cg_partial = nullptr, *cg_scalars = nullptr;
  double *rbuf = nullptr;      // staging for real components that are not 16-byte aligned (odd slab sizes)
};

namespace {

// Grid cap of the exchange kernel.  Alone on the GPU it takes 4 CTAs per SM.  In the pipelined apply it
// shares the SMs with cuFFT's kernels, and peer stores that wait for NVLink back up the memory pipeline of
// every SM they are issued from: with 4 CTAs on every SM the "overlapped" time was the SUM of transform and
// exchange.  Measured on 8 GPUs at 1024^3 (profiles/r02_measurements.md, ms per apply complex / real):
// 592 CTAs 27.5 / 16.5, 296 -> 23.9 / 15.7, 148 -> 25.9 / 15.5, 96 -> 27.0 / 14.6, 64 -> 28.2 / 15.7.
// The half spectrum moves half the bytes and needs fewer CTAs to keep NVLink busy.
int copy_grid_cap(const bri17_rs_plan *p, const Layout &l, bool pipelined) {
  if (p->copy_ctas > 0) return p->copy_ctas;
  if (!pipelined) return p->sm_count * 4;
  return l.real ? std::max(64, (p->sm_count * 2) / 3) : p->sm_count * 2;
}

int launch_copy(CopyPlan &cp, int fence, cudaStream_t st, int max_ctas) {
  long long per_seg = 0;
  for (int s = 0; s < cp.nseg; s++)
    per_seg = std::max(per_seg, (long long)cp.ncomp * cp.seg[s].rows * cp.parts);
  cp.ctas_per_seg = per_seg;
  const long long total = per_seg * cp.nseg;
  if (total == 0) return BRI17_OK;
  const unsigned grid = unsigned(std::min<long long>(total, max_ctas));
  slab_copy_kernel<<<grid, 256, 0, st>>>(cp, fence);
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

// cross-GPU, stream-ordered barrier (and memory fence).  Mode 1: flags in peer memory
// (which = 0 caller's stream, 1 exchange stream: independent counters); mode 0: 1-element all-reduce.
int stream_barrier(bri17_rs_plan *p, cudaStream_t st, int which = 0) {
  if (p->nranks == 1) return BRI17_OK;
  if (p->mode == 1) {
    flag_barrier_kernel<<<1, 32, 0, st>>>(p->flags, which ? FLAG_BAR1 : FLAG_BAR0, ++p->epoch_bar[which]);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(BRI17_ERR_CUDA, std::string("flag barrier launch: ") + cudaGetErrorString(e));
    return BRI17_OK;
  }
  RS_NCCL_TRY(ncclAllReduce(p->barrier_word, p->barrier_word, 1, ncclDouble, ncclSum, p->comm, st));
  return BRI17_OK;
}
// same on the exchange stream of the pipelined apply (mode 1 only)
int exchange_barrier(bri17_rs_plan *p) { return stream_barrier(p, p->sx, 1); }

// v[0..n) <- sum over ranks, on the stream, result identical on every rank (n <= RED_MAX)
int scalar_allreduce(bri17_rs_plan *p, double *v, int n, cudaStream_t st) {
  if (p->nranks == 1) return BRI17_OK;
  if (p->mode == 1) {
    peer_allreduce_kernel<<<1, 32, 0, st>>>(p->flags, v, n, ++p->epoch_red);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(BRI17_ERR_CUDA, std::string("peer all-reduce launch: ") + cudaGetErrorString(e));
    return BRI17_OK;
  }
  RS_NCCL_TRY(ncclAllReduce(v, v, size_t(n), ncclDouble, ncclSum, p->comm, st));
  return BRI17_OK;
}

// Forward exchange: local-transform layout T[c][a][b][k2] (a in my n0 slab, b over S1)
// -> Fourier-side layout X[c][n0][b_loc][k2].  `S` is the packed send buffer (mode 0).
// c0/ncomp select components of the multi-component arrays T and X; `barriers` = false leaves
// the cross-GPU synchronisation to the caller (pipelined apply).
// a0/na (mode 1 only): restrict the exchange to planes [a0, a0 + na) of this rank's n0 slab (na < 0: all).
int exchange_forward(bri17_rs_plan *p, const Layout &l, const double2 *T, double2 *X, double2 *S, int ncomp,
                     cudaStream_t st, int c0 = 0, bool barriers = true, bool xt = false, int a0 = 0, int na = -1,
                     CopyPlan *plan_only = nullptr) {
  const int P = p->nranks, r = p->rank, N0 = p->shape[0], S1 = l.S1, S2e = l.S2e;
  CopyPlan cp{};
  cp.ncomp = ncomp;
  cp.scale = 1.0;
  std::vector<long long> off(P + 1, 0);  // packed offsets (elements)
  for (int q = 0; q < P; q++)
    off[q + 1] = off[q] + (long long)ncomp * p->n0_loc * (l.k1_beg[q + 1] - l.k1_beg[q]) * S2e;
  int maxlen = 0;
  for (int i = 1; i <= P; i++) {
    const int q = (r + i) % P;
    const int n1q = l.k1_beg[q + 1] - l.k1_beg[q];
    if (n1q == 0 || p->n0_loc == 0) continue;
    CopySeg &g = cp.seg[cp.nseg++];
    g.src_cs = (long long)p->n0_loc * S1 * S2e;
    g.src_rs = (long long)S1 * S2e;
    g.src = T + (long long)l.k1_beg[q] * S2e + c0 * g.src_cs;
    g.rows = p->n0_loc;
    g.len = n1q * S2e;
    g.inner = S2e;
    g.src_bs = g.dst_bs = S2e;
    maxlen = std::max(maxlen, g.len);
    const bool direct = (q == r) || p->mode == 1;  // store at the final position
    if (direct && xt) {  // k1-major Fourier-side layout X[c][b][n0][k2]
      double2 *base = (q == r) ? X : p->peerW[q];
      g.dst_cs = (long long)N0 * n1q * S2e;
      g.dst_rs = S2e;
      g.dst_bs = (long long)N0 * S2e;
      g.dst = base + (long long)p->n0_beg[r] * S2e + c0 * g.dst_cs;
    } else if (direct) {
      double2 *base = (q == r) ? X : p->peerW[q];
      g.dst_cs = (long long)N0 * n1q * S2e;
      g.dst_rs = (long long)n1q * S2e;
      g.dst = base + (long long)p->n0_beg[r] * n1q * S2e + c0 * g.dst_cs;
    } else {
      g.dst = S + off[q];
      g.dst_cs = (long long)p->n0_loc * n1q * S2e;
      g.dst_rs = (long long)n1q * S2e;
    }
    if (na >= 0) {  // sub-slab of planes
      if (na == 0) { cp.nseg--; continue; }
      g.rows = na;
      g.src += (long long)a0 * g.src_rs;
      g.dst += (long long)a0 * g.dst_rs;
    }
  }
  cp.parts = choose_parts(maxlen, (long long)ncomp * (na >= 0 ? na : p->n0_loc) * P);
  if (plan_only) { *plan_only = cp; return BRI17_OK; }  // CPU replay of the index arithmetic (tests)
  if (p->mode == 1 && barriers) RS_TRY(stream_barrier(p, st));  // peers' buffers are free to overwrite
  RS_TRY(launch_copy(cp, p->mode == 1, st, copy_grid_cap(p, l, !barriers)));
  if (P == 1) return BRI17_OK;
  if (p->mode == 1) return barriers ? stream_barrier(p, st) : BRI17_OK;   // everybody's stores have landed
  RS_NCCL_TRY(ncclGroupStart());
  for (int q = 0; q < P; q++) {
    if (q == r) continue;
    const long long n1q = l.k1_beg[q + 1] - l.k1_beg[q], n0q = p->n0_beg[q + 1] - p->n0_beg[q];
    for (int c = 0; c < ncomp; c++) {
      const long long scount = (long long)p->n0_loc * n1q * S2e, rcount = n0q * l.n1_loc * S2e;
      if (scount) RS_NCCL_TRY(ncclSend(S + off[q] + c * scount, size_t(2 * scount), ncclDouble, q, p->comm, st));
      if (rcount)
        RS_NCCL_TRY(ncclRecv(X + ((long long)c * N0 + p->n0_beg[q]) * l.n1_loc * S2e, size_t(2 * rcount),
                             ncclDouble, q, p->comm, st));
    }
  }
  RS_NCCL_TRY(ncclGroupEnd());
  return BRI17_OK;
}

// Backward exchange: Fourier-side X[c][n0][b_loc][k2] -> local-transform layout D[c][a][b][k2] (times scale).
// mode 0: R = packed receive buffer, D written by the unpack; mode 1: peers store into our W2 (= D).
// chunk/nchunks (mode 1 only): restrict the exchange to the chunk-th of nchunks parts of EVERY
// destination's n0 slab (planes [n0q*chunk/nchunks, n0q*(chunk+1)/nchunks) of rank q).
int exchange_backward(bri17_rs_plan *p, const Layout &l, const double2 *X, double2 *D, double2 *R, int ncomp,
                      double scale, cudaStream_t st, int c0 = 0, bool barriers = true, bool xt = false,
                      int chunk = 0, int nchunks = 1, CopyPlan *plan_only = nullptr) {
  const int P = p->nranks, r = p->rank, N0 = p->shape[0], S1 = l.S1, S2e = l.S2e;
  if (p->mode == 1 || P == 1) {
    // every row (c, n0) goes, whole, to the owner of n0, at its final position
    CopyPlan cp{};
    cp.ncomp = ncomp;
    cp.scale = scale;
    for (int i = 1; i <= P; i++) {
      const int q = (r + i) % P;
      const int n0q = p->n0_beg[q + 1] - p->n0_beg[q];
      if (n0q == 0 || l.n1_loc == 0) continue;
      CopySeg &g = cp.seg[cp.nseg++];
      g.inner = S2e;
      g.src_bs = g.dst_bs = S2e;
      g.src_cs = (long long)N0 * l.n1_loc * S2e;
      if (xt) {  // k1-major source X[c][b][n0][k2]
        g.src_rs = S2e;
        g.src_bs = (long long)N0 * S2e;
        g.src = X + (long long)p->n0_beg[q] * S2e + c0 * g.src_cs;
      } else {
        g.src_rs = (long long)l.n1_loc * S2e;
        g.src = X + (long long)p->n0_beg[q] * l.n1_loc * S2e + c0 * g.src_cs;
      }
      g.rows = n0q;
      g.len = l.n1_loc * S2e;
      double2 *base = (q == r) ? D : p->peerW2[q];
      g.dst_cs = (long long)n0q * S1 * S2e;
      g.dst_rs = (long long)S1 * S2e;
      g.dst = base + (long long)l.k1_beg[r] * S2e + c0 * g.dst_cs;
      if (nchunks > 1) {
        const int a0 = int((long long)n0q * chunk / nchunks), a1 = int((long long)n0q * (chunk + 1) / nchunks);
        if (a1 == a0) { cp.nseg--; continue; }
        g.rows = a1 - a0;
        g.src += (long long)a0 * g.src_rs;
        g.dst += (long long)a0 * g.dst_rs;
      }
    }
    cp.parts = choose_parts(l.n1_loc * S2e, (long long)ncomp * N0 / nchunks);
    if (plan_only) { *plan_only = cp; return BRI17_OK; }  // CPU replay (tests)
    if (barriers) RS_TRY(stream_barrier(p, st));
    RS_TRY(launch_copy(cp, P > 1, st, copy_grid_cap(p, l, !barriers)));
    return barriers ? stream_barrier(p, st) : BRI17_OK;
  }
  std::vector<long long> off(P + 1, 0);
  for (int q = 0; q < P; q++)
    off[q + 1] = off[q] + (long long)ncomp * p->n0_loc * (l.k1_beg[q + 1] - l.k1_beg[q]) * S2e;
  RS_NCCL_TRY(ncclGroupStart());
  for (int q = 0; q < P; q++) {
    if (q == r) continue;
    const long long n1q = l.k1_beg[q + 1] - l.k1_beg[q], n0q = p->n0_beg[q + 1] - p->n0_beg[q];
    for (int c = 0; c < ncomp; c++) {
      const long long scount = n0q * l.n1_loc * S2e, rcount = (long long)p->n0_loc * n1q * S2e;
      if (scount)
        RS_NCCL_TRY(ncclSend(X + ((long long)c * N0 + p->n0_beg[q]) * l.n1_loc * S2e, size_t(2 * scount),
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
    const int n1q = l.k1_beg[q + 1] - l.k1_beg[q];
    if (n1q == 0 || p->n0_loc == 0) continue;
    CopySeg &g = cp.seg[cp.nseg++];
    g.rows = p->n0_loc;
    g.len = n1q * S2e;
    g.inner = S2e;
    g.src_bs = g.dst_bs = S2e;
    maxlen = std::max(maxlen, g.len);
    if (q == r) {  // own block straight from X
      g.src = X + (long long)p->n0_beg[r] * n1q * S2e;
      g.src_cs = (long long)N0 * n1q * S2e;
      g.src_rs = (long long)n1q * S2e;
    } else {
      g.src = R + off[q];
      g.src_cs = (long long)p->n0_loc * n1q * S2e;
      g.src_rs = (long long)n1q * S2e;
    }
    g.dst = D + (long long)l.k1_beg[q] * S2e;
    g.dst_cs = (long long)p->n0_loc * S1 * S2e;
    g.dst_rs = (long long)S1 * S2e;
  }
  cp.parts = choose_parts(maxlen, (long long)ncomp * p->n0_loc * P);
  return launch_copy(cp, 0, st, copy_grid_cap(p, l, false));
}

// ---- chunked batched transforms (BatchFft) ---------------------------------------------------
// Planes per chunk for a slab whose spectral planes hold `plane_elems` complex values.
int chunk_planes(const bri17_rs_plan *p, long long plane_elems, long long real_plane, int planes) {
  if ((p->fft_chunk_mib <= 0 && p->fft_chunk_planes <= 0) || planes <= 1) return std::max(planes, 1);
  long long c = p->fft_chunk_planes > 0
                    ? p->fft_chunk_planes
                    : (static_cast<long long>(p->fft_chunk_mib) << 20) / std::max<long long>(plane_elems * 16, 1);
  c = std::max<long long>(1, std::min<long long>(c, planes));
  // real planes with an odd number of doubles: keep every chunk start 16-byte aligned
  if ((real_plane & 1) && (c & 1) && c < planes) c = c > 1 ? c - 1 : 2;
  return int(std::min<long long>(c, planes));
}

int batch_handle(BatchFft &f, int batch, cufftHandle *out) {
  for (int i = 0; i < f.ncache; i++)
    if (f.cache[i].batch == batch) { *out = f.cache[i].h; return BRI17_OK; }
  if (f.ncache == 8) return fail(BRI17_ERR_UNSUPPORTED, "too many distinct batch sizes for one transform");
  size_t ws = 0;
  cufftHandle h = 0;
  RS_CUFFT_TRY(cufftCreate(&h));
  cufftResult r = cufftMakePlanMany64(h, f.rank, f.n, f.embed ? f.inembed : nullptr, f.istride, f.idist,
                                      f.embed ? f.onembed : nullptr, f.ostride, f.odist, f.type, batch, &ws);
  if (r != CUFFT_SUCCESS) {
    cufftDestroy(h);
    return fail(BRI17_ERR_CUDA, "cufftMakePlanMany64: cuFFT error " + std::to_string(int(r)));
  }
  f.cache[f.ncache].batch = batch;
  f.cache[f.ncache].h = h;
  f.ncache++;
  *out = h;
  return BRI17_OK;
}

int batch_make(BatchFft &f, int rank, const long long *n, const long long *inembed, long long istride, long long idist,
               const long long *onembed, long long ostride, long long odist, cufftType type, int planes, int chunk) {
  f = BatchFft{};
  f.rank = rank;
  for (int i = 0; i < rank; i++) {
    f.n[i] = n[i];
    if (inembed) f.inembed[i] = inembed[i];
    if (onembed) f.onembed[i] = onembed[i];
  }
  f.embed = inembed != nullptr;
  f.istride = istride; f.ostride = ostride;
  f.idist = idist; f.odist = odist;
  f.type = type;
  f.planes = planes;
  f.chunk = std::max(1, std::min(chunk, planes));
  f.made = true;
  cufftHandle h;
  return batch_handle(f, f.chunk, &h);  // the plan every full chunk uses, created up front
}

void batch_destroy(BatchFft &f) {
  for (int i = 0; i < f.ncache; i++) cufftDestroy(f.cache[i].h);
  f = BatchFft{};
}

// `na` planes, in / out pointing AT the first of them; dir is used by Z2Z only.
int batch_exec_at(BatchFft &f, const void *in, void *out, int na, int dir, cudaStream_t st) {
  if (!f.made || na <= 0) return BRI17_OK;
  cufftHandle h;
  RS_TRY(batch_handle(f, na, &h));
  RS_CUFFT_TRY(cufftSetStream(h, st));
  void *a = const_cast<void *>(in), *b = out;
  if (f.type == CUFFT_Z2Z) RS_CUFFT_TRY(cufftExecZ2Z(h, (cufftDoubleComplex *)a, (cufftDoubleComplex *)b, dir));
  else if (f.type == CUFFT_D2Z) RS_CUFFT_TRY(cufftExecD2Z(h, (cufftDoubleReal *)a, (cufftDoubleComplex *)b));
  else RS_CUFFT_TRY(cufftExecZ2D(h, (cufftDoubleComplex *)a, (cufftDoubleReal *)b));
  return BRI17_OK;
}

// Planes [a0, a0 + na) of the component whose first plane is at in / out.
int batch_exec_range(BatchFft &f, const void *in, void *out, int a0, int na, int dir, cudaStream_t st) {
  if (!f.made || na <= 0) return BRI17_OK;
  const size_t isz = f.type == CUFFT_D2Z ? 8 : 16, osz = f.type == CUFFT_Z2D ? 8 : 16;
  return batch_exec_at(f, static_cast<const char *>(in) + size_t(a0) * f.idist * isz,
                       static_cast<char *>(out) + size_t(a0) * f.odist * osz, na, dir, st);
}

// The whole slab of one component, `chunk` planes per call.
int batch_exec(BatchFft &f, const void *in, void *out, int dir, cudaStream_t st) {
  if (!f.made) return BRI17_OK;
  for (int a0 = 0; a0 < f.planes; a0 += f.chunk)
    RS_TRY(batch_exec_range(f, in, out, a0, std::min(f.chunk, f.planes - a0), dir, st));
  return BRI17_OK;
}

// c2c local transform over the trailing axes (complex layout only)
int fft_local_c2c(bri17_rs_plan *p, const double2 *in, double2 *out, int ncomp, int dir, cudaStream_t st) {
  Layout &l = p->lc;
  if (!l.have_local) {
    if (in != out && l.t_count)
      BRI17_CUDA_TRY(cudaMemcpyAsync(out, in, sizeof(double2) * ncomp * l.t_count, cudaMemcpyDeviceToDevice, st));
    return BRI17_OK;
  }
  for (int c = 0; c < ncomp; c++) RS_TRY(batch_exec(l.fwd_local, in + c * l.t_count, out + c * l.t_count, dir, st));
  return BRI17_OK;
}

int fft_axis0(bri17_rs_plan *p, const Layout &l, double2 *x, int ncomp, int dir, cudaStream_t st) {
  (void)p;
  if (!l.have_axis0) return BRI17_OK;
  RS_CUFFT_TRY(cufftSetStream(l.axis0, st));
  for (int c = 0; c < ncomp; c++)
    RS_CUFFT_TRY(cufftExecZ2Z(l.axis0, (cufftDoubleComplex *)(x + c * l.fourier_count),
                              (cufftDoubleComplex *)(x + c * l.fourier_count), dir));
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

// Geometry of a layout (no cuFFT plan yet).
void layout_geometry(const bri17_rs_plan *p, Layout &l, bool real) {
  const int dim = p->dim, P = p->nranks;
  l.real = real;
  const int last = p->shape[dim - 1];
  if (dim == 3) { l.S1 = p->shape[1]; l.S2e = real ? last / 2 + 1 : last; }
  else { l.S1 = real ? last / 2 + 1 : last; l.S2e = 1; }
  for (int q = 0; q <= P; q++) l.k1_beg[q] = int((int64_t)q * l.S1 / P);
  l.n1_loc = l.k1_beg[p->rank + 1] - l.k1_beg[p->rank];
  l.t_count = (int64_t)p->n0_loc * l.S1 * l.S2e;
  l.fourier_count = (int64_t)p->shape[0] * l.n1_loc * l.S2e;
}

// cuFFT plans of a layout whose geometry is set.
int setup_layout(bri17_rs_plan *p, Layout &l) {
  const int dim = p->dim;
  const bool real = l.real;
  size_t ws = 0;
  if (p->n0_loc > 0) {  // local transform over the trailing axes, one slab plane per batch entry
    long long n64[2] = {p->shape[1], p->N2e};
    const int frank = dim - 1;
    const long long rplane = (long long)p->shape[1] * p->N2e, splane = (long long)l.S1 * l.S2e;
    // a 1-D transform (2-D grids) is a single kernel: nothing to gain from chunks
    const int chunk = frank == 2 ? chunk_planes(p, splane, real ? rplane : 0, p->n0_loc) : p->n0_loc;
    if (!real) {
      RS_TRY(batch_make(l.fwd_local, frank, n64, nullptr, 1, rplane, nullptr, 1, rplane, CUFFT_Z2Z, p->n0_loc, chunk));
    } else {
      RS_TRY(batch_make(l.fwd_local, frank, n64, nullptr, 1, rplane, nullptr, 1, splane, CUFFT_D2Z, p->n0_loc, chunk));
      RS_TRY(batch_make(l.inv_local, frank, n64, nullptr, 1, splane, nullptr, 1, rplane, CUFFT_Z2D, p->n0_loc, chunk));
    }
    l.have_local = true;
  }
  if (l.n1_loc > 0) {  // axis 0 on the Fourier-side block: stride = batch = n1_loc*S2e
    const long long S = (long long)l.n1_loc * l.S2e;
    long long n64[1] = {p->shape[0]}, embed[1] = {p->shape[0]};
    RS_CUFFT_TRY(cufftCreate(&l.axis0));
    RS_CUFFT_TRY(cufftMakePlanMany64(l.axis0, 1, n64, embed, S, 1, embed, S, 1, CUFFT_Z2Z, S, &ws));
    l.have_axis0 = true;
  }
  l.ready = true;
  return BRI17_OK;
}

// Drops the cuFFT plans, keeps the geometry.
void destroy_layout(Layout &l) {
  batch_destroy(l.fwd_local);
  batch_destroy(l.inv_local);
  if (l.have_axis0) cufftDestroy(l.axis0);
  batch_destroy(l.fwd_local_t);
  batch_destroy(l.inv_local_t);
  l.axis0 = 0;
  l.have_local = l.have_axis0 = l.have_local_t = l.ready = false;
}

// Exchange buffers: always present for P > 1 (allocated at creation so that their IPC
// handles can be exchanged); for P == 1 only the real path needs one (lazily).
int ensure_buffers(bri17_rs_plan *p, size_t bytes) {
  if (p->buf_bytes >= bytes) return BRI17_OK;
  if (p->nranks > 1) return fail(BRI17_ERR_UNSUPPORTED, "exchange buffers cannot grow after creation");
  if (p->W2) cudaFree(p->W2);
  p->W2 = nullptr;
  BRI17_CUDA_TRY(cudaMalloc(&p->W2, bytes));
  p->buf_bytes = bytes;
  return BRI17_OK;
}

// Is the fused axis-0 pass used for this plan?
bool use_fused(const bri17_rs_plan *p) { return p->fused && p->tabs && bri17b200::axis0::supported(p->shape[0]); }

// Is the Fourier-side block kept k1-major, [c][k1][n0][k2] (axis0_fused.cuh, "Global layout")?
// 3-D only, with the fused pass, when the producer can write it: the fused exchange kernel (mode 1)
// or, on one GPU, cuFFT's advanced output layout.
// Option "k1_major": 1 always, 0 never, -1 (default) where it pays: always with the fused exchange
// (free); on one GPU only for complex fields -- cuFFT's transposed output costs 0.8 ms per 512^3
// apply, the fused pass gains 1.8 ms on complex fields (12.96 vs 14.06 ms per apply) but only 0.56 ms
// on the half spectrum (8.53 vs 8.24 ms), profiles/r02_measurements.md.
bool use_xt(const bri17_rs_plan *p, const Layout &l) {
  if (p->xt == 0 || p->dim != 3 || !use_fused(p) || !(p->nranks == 1 || p->mode == 1)) return false;
  if (p->xt > 0 || p->nranks > 1) return true;
  return !l.real;
}

// Single GPU: local transforms over axes (1, 2) whose spectral side is k1-major:
// element (k1, k2) of plane n0 at k1*(N0*S2e) + n0*S2e + k2.
int setup_layout_t(bri17_rs_plan *p, Layout &l) {
  if (l.have_local_t || p->n0_loc == 0) return BRI17_OK;
  long long n64[2] = {p->shape[1], p->N2e};
  const long long rplane = (long long)p->shape[1] * p->N2e;
  long long nat[2] = {p->shape[1], p->N2e};                               // natural real-space planes
  long long spec[2] = {l.S1, (long long)p->shape[0] * l.S2e};            // k1 stride = N0*S2e
  const int chunk = chunk_planes(p, (long long)l.S1 * l.S2e, l.real ? rplane : 0, p->n0_loc);
  if (!l.real) {
    RS_TRY(batch_make(l.fwd_local_t, 2, n64, nat, 1, rplane, spec, 1, l.S2e, CUFFT_Z2Z, p->n0_loc, chunk));
    RS_TRY(batch_make(l.inv_local_t, 2, n64, spec, 1, l.S2e, nat, 1, rplane, CUFFT_Z2Z, p->n0_loc, chunk));
  } else {
    RS_TRY(batch_make(l.fwd_local_t, 2, n64, nat, 1, rplane, spec, 1, l.S2e, CUFFT_D2Z, p->n0_loc, chunk));
    RS_TRY(batch_make(l.inv_local_t, 2, n64, spec, 1, l.S2e, nat, 1, rplane, CUFFT_Z2D, p->n0_loc, chunk));
  }
  l.have_local_t = true;
  return BRI17_OK;
}

// Axis-0 section of the apply on the Fourier-side block X of a layout, in place:
//   FFT(axis 0) -> K^ . (.) * |h|/|N| -> unnormalised inverse FFT(axis 0)
// (tests/test_bri17.cpp:57 [axis 0], :58-92, :93-106, :95 [axis 0]).  One kernel when N0 is a
// supported power of two, cuFFT + modal kernel + cuFFT otherwise.  dot_dev (optional, device
// scalar) receives sum_k w_k Re(u^_k^H f^_k) over THIS rank's block (= its share of <u, A u>).
// Events 3 and 4 bracket the modal kernel (or the whole fused pass).
int modal_section(bri17_rs_plan *p, const Layout &l, double2 *X, cudaStream_t st, double *dot_dev, bool xt) {
  const int dim = p->dim;
  const int herm_n = l.real ? p->shape[dim - 1] : 0;
  if (dot_dev && !p->dot_scratch) BRI17_CUDA_TRY(cudaMalloc(&p->dot_scratch, sizeof(double) * RED_CTAS));
  if (l.fourier_count == 0) {
    mark(p, 3, st);
    mark(p, 4, st);
    mark(p, 5, st);
    if (dot_dev) BRI17_CUDA_TRY(cudaMemsetAsync(dot_dev, 0, sizeof(double), st));
    return BRI17_OK;
  }
  if (use_fused(p)) {
    mark(p, 3, st);
    bri17b200::axis0::Params a{};
    a.X = X;
    a.comp_stride = l.fourier_count;
    a.S = (long long)l.n1_loc * l.S2e;
    if (xt) {  // [k1][n0][k2]
      a.row_stride = l.S2e;
      a.blk_cols = l.S2e;
      a.blk_stride = (long long)p->shape[0] * l.S2e;
    } else {   // [n0][k1][k2]
      a.row_stride = a.S;
      a.blk_cols = std::max<long long>(a.S, 1);
      a.blk_stride = 0;
    }
    a.N0 = p->shape[0];
    a.S2e = dim == 3 ? l.S2e : 1;
    a.k1_begin = l.k1_beg[p->rank];
    a.tab0 = p->tabs + p->tab_off[0];
    a.tab1 = p->tabs + p->tab_off[1];
    a.tab2 = dim == 3 ? p->tabs + p->tab_off[2] : nullptr;
    a.N1 = p->shape[1];
    a.N2 = dim == 3 ? p->shape[2] : 1;
    a.twiddle = p->twiddle;
    a.mu = p->mu;
    a.scaling = p->mu / (1. - 2. * p->nu);  // bri17.hpp:266
    a.out_scale = p->correction;
    a.dot_partial = dot_dev ? p->dot_scratch : nullptr;
    a.herm_n = herm_n;
    int grid = 0;
    RS_TRY(bri17b200::axis0::launch(a, dim, p->sm_count, dot_dev ? RED_CTAS : 0, st, &grid));
    p->fused_launches++;
    if (dot_dev) cg_finish_kernel<<<1, RED_THREADS, 0, st>>>(p->dot_scratch, grid, dot_dev);
    mark(p, 4, st);
    mark(p, 5, st);
    return BRI17_OK;
  }
  if (xt) return fail(BRI17_ERR_UNSUPPORTED, "k1-major layout without the fused axis-0 pass");
  RS_TRY(fft_axis0(p, l, X, dim, CUFFT_FORWARD, st));                       // :57 (axis 0)
  mark(p, 3, st);
  int kb[3] = {0, l.k1_beg[p->rank], 0};
  int ls[3] = {p->shape[0], l.n1_loc, l.S2e};
  if (dot_dev)
    RS_TRY(bri17_modal_stiffness_apply_dot_f64(p->modal, X, X, kb, ls, 0, p->correction, herm_n, dot_dev,
                                               p->dot_scratch, RED_CTAS, st));
  else
    RS_TRY(bri17_modal_stiffness_apply_f64(p->modal, X, X, kb, ls, 0, p->correction, st));  // :58-92, scale :93-106
  mark(p, 4, st);
  RS_TRY(fft_axis0(p, l, X, dim, CUFFT_INVERSE, st));                       // :95
  mark(p, 5, st);
  return BRI17_OK;
}

// Shared CG driver over double arrays; `apply(d, Ad, dot)` is the operator, which also leaves this
// rank's share of <d, A d> in the device scalar `dot`.  The fields hold `nslot/inter` components
// of `count*inter` doubles (inter = 2: interleaved complex).
template <typename Apply>
int cg_core(bri17_rs_plan *p, Apply apply, const double *b, double *x, long long count, int inter, double rtol,
            int max_iter, int check_every, int *iterations, double *rel_residual, cudaStream_t st) {
  const int nslot = p->dim * inter;
  const long long n = count * nslot;
  const size_t bytes = sizeof(double) * std::max<long long>(n, 2);
  if (p->cg_len < bytes) {
    for (double *ptr : {p->cg_r, p->cg_p, p->cg_Ap})
      if (ptr) cudaFree(ptr);
    p->cg_r = p->cg_p = p->cg_Ap = nullptr;
    BRI17_CUDA_TRY(cudaMalloc(&p->cg_r, bytes));
    BRI17_CUDA_TRY(cudaMalloc(&p->cg_p, bytes));
    BRI17_CUDA_TRY(cudaMalloc(&p->cg_Ap, bytes));
    p->cg_len = bytes;
  }
  if (!p->cg_partial) {
    BRI17_CUDA_TRY(cudaMalloc(&p->cg_partial, sizeof(double) * std::max(RED_CTAS, 8 * MEAN_CTAS)));
    BRI17_CUDA_TRY(cudaMalloc(&p->cg_scalars, sizeof(double) * 16));
  }
  double *r = p->cg_r, *d = p->cg_p, *Ad = p->cg_Ap;
  double *sc = p->cg_scalars;  // [0] rr (even iter) [1] rr (odd iter) [2] pAp [8..8+nslot) component sums
  auto reduce_to = [&](double *slot) -> int {
    cg_finish_kernel<<<1, RED_THREADS, 0, st>>>(p->cg_partial, RED_CTAS, slot);
    return scalar_allreduce(p, slot, 1, st);
  };
  // zero-frequency projection of the right-hand side, then r = p = b - mean, x = 0, <r, r>
  double total = 1.;
  for (int dd = 0; dd < p->dim; dd++) total *= p->shape[dd];
  cg_slot_sum_kernel<<<dim3(MEAN_CTAS, nslot), RED_THREADS, 0, st>>>(b, count, inter, p->cg_partial);
  cg_slot_finish_kernel<<<nslot, RED_THREADS, 0, st>>>(p->cg_partial, sc + 8);
  RS_TRY(scalar_allreduce(p, sc + 8, nslot, st));
  cg_init_kernel<<<RED_CTAS, RED_THREADS, 0, st>>>(b, x, r, d, count, inter, nslot, sc + 8, 1. / total,
                                                    p->cg_partial);
  RS_TRY(reduce_to(sc + 0));
  double bb = 0.;
  BRI17_CUDA_TRY(cudaMemcpyAsync(&bb, sc + 0, sizeof(double), cudaMemcpyDeviceToHost, st));
  BRI17_CUDA_TRY(cudaStreamSynchronize(st));
  if (!std::isfinite(bb)) return fail(BRI17_ERR_BREAKDOWN, "CG: the right-hand side is not finite");
  int it = 0;
  double rr_host = bb;
  if (bb > 0.) {
    for (; it < max_iter;) {
      double *rr = sc + (it & 1), *rr_new = sc + ((it + 1) & 1);
      RS_TRY(apply(d, Ad, sc + 2));
      RS_TRY(scalar_allreduce(p, sc + 2, 1, st));
      cg_residual_kernel<<<RED_CTAS, RED_THREADS, 0, st>>>(r, Ad, n, rr, sc + 2, p->cg_partial);
      RS_TRY(reduce_to(rr_new));
      cg_direction_kernel<<<RED_CTAS, RED_THREADS, 0, st>>>(x, d, r, n, rr, sc + 2, rr_new);
      it++;
      if (check_every > 0 && (it % check_every == 0 || it == max_iter)) {
        BRI17_CUDA_TRY(cudaMemcpyAsync(&rr_host, rr_new, sizeof(double), cudaMemcpyDeviceToHost, st));
        BRI17_CUDA_TRY(cudaStreamSynchronize(st));
        if (!std::isfinite(rr_host)) break;
        if (rr_host <= rtol * rtol * bb) break;
      }
    }
    if (check_every <= 0) {
      BRI17_CUDA_TRY(cudaMemcpyAsync(&rr_host, sc + (it & 1), sizeof(double), cudaMemcpyDeviceToHost, st));
      BRI17_CUDA_TRY(cudaStreamSynchronize(st));
    }
  } else {
    BRI17_CUDA_TRY(cudaStreamSynchronize(st));
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(BRI17_ERR_CUDA, std::string("CG: ") + cudaGetErrorString(e));
  if (iterations) *iterations = it;
  if (rel_residual) *rel_residual = bb > 0. ? std::sqrt(rr_host / bb) : 0.;
  if (!std::isfinite(rr_host))
    return fail(BRI17_ERR_BREAKDOWN, "CG broke down: the residual norm is not finite after " +
                                         std::to_string(it) + " iterations");
  return BRI17_OK;
}

// Pipelined apply for nranks > 1 with the fused peer-store exchange: the exchange runs on its own
// high-priority stream while the transforms run on the caller's stream, so NVLink-bound and
// HBM-bound work overlap.  The unit of the pipeline is a SUB-SLAB: J chunks of n0 planes per
// component (J = exchange_chunks(): 1..4, larger for larger slabs).
//   st: L00 L01 .. L2J |wait| fused axis-0 pass |      wait Q00: L'00, wait Q01: L'01 ...
//   sx:     P00 P01 ..  P2J, barrier            | Q00 b Q01 b ... Q2J b
// L = local transform over the trailing axes of a sub-slab, P/Q = slab_copy_kernel storing into the
// peers' W / W2 (b = flag barrier).  Only the last P and the first Q are exposed: 1/(3J) of each
// exchange (round 1 pipelined whole components: 1/3).  The forward direction needs ONE barrier (its
// consumer, the axis-0 pass, needs every piece anyway); backward, the consumer of sub-slab (c, j) is
// the local inverse transform of exactly those planes, so every Q is followed by a barrier.
// No "buffer free" barrier is needed: every exchange ends with a barrier after its last read, and
// the next writer is ordered behind it.
// Without the fused axis-0 pass (cuFFT per component) the schedule is per component, as in round 1:
//   st: L0 L1 L2 |wait 0| A0 |wait 1| A1 A2  modal  A'0 A'1 A'2      |wait| L'0 L'1 L'2
//   sx:    P0b P1b P2b                               Q0b Q1b Q2b
int exchange_chunks(const bri17_rs_plan *p, const Layout &l) {
  if (p->xchunks > 0) return std::min(p->xchunks, MAX_XCHUNKS);
  // from the LARGEST slab of the decomposition, not this rank's own: every rank must come to the same
  // count (one barrier per backward sub-slab), also when the slabs are uneven or some are empty
  const long long planes = (p->shape[0] + p->nranks - 1) / p->nranks;
  const long long bytes = 16ll * planes * l.S1 * l.S2e;  // one component of that slab
  return int(std::max<long long>(1, std::min<long long>(MAX_XCHUNKS, bytes >> 27)));  // >= 128 MiB per sub-slab
}

template <typename LocalFwd, typename LocalInv>
int apply_pipelined(bri17_rs_plan *p, const Layout &l, const double2 *T, LocalFwd local_fwd, LocalInv local_inv,
                    cudaStream_t st, double *dot_dev) {
  const int dim = p->dim;
  double2 *X = p->W;
  const bool xt = use_xt(p, l);
  const bool fused = use_fused(p);
  p->timings_valid = false;
  mark(p, 0, st);
  if (fused || dot_dev) {
    const int J = exchange_chunks(p, l);
    auto lo = [&](int j) { return int((long long)p->n0_loc * j / J); };
    for (int c = 0; c < dim; c++)
      for (int j = 0; j < J; j++) {
        const int e = c * J + j, a0 = lo(j), na = lo(j + 1) - a0;
        RS_TRY(local_fwd(c, a0, na));
        BRI17_CUDA_TRY(cudaEventRecord(p->ev_a[e], st));
        BRI17_CUDA_TRY(cudaStreamWaitEvent(p->sx, p->ev_a[e], 0));
        RS_TRY(exchange_forward(p, l, T, X, nullptr, 1, p->sx, c, false, xt, a0, na));
      }
    RS_TRY(exchange_barrier(p));
    BRI17_CUDA_TRY(cudaEventRecord(p->ev_b[0], p->sx));
    mark(p, 1, st);
    mark(p, 2, st);
    BRI17_CUDA_TRY(cudaStreamWaitEvent(st, p->ev_b[0], 0));
    RS_TRY(modal_section(p, l, X, st, dot_dev, xt));
    BRI17_CUDA_TRY(cudaEventRecord(p->ev_a[0], st));
    BRI17_CUDA_TRY(cudaStreamWaitEvent(p->sx, p->ev_a[0], 0));
    for (int c = 0; c < dim; c++)
      for (int j = 0; j < J; j++) {
        RS_TRY(exchange_backward(p, l, X, p->W2, nullptr, 1, 1.0, p->sx, c, false, xt, j, J));
        RS_TRY(exchange_barrier(p));
        BRI17_CUDA_TRY(cudaEventRecord(p->ev_b[c * J + j], p->sx));
      }
    mark(p, 6, st);
    for (int c = 0; c < dim; c++)
      for (int j = 0; j < J; j++) {
        BRI17_CUDA_TRY(cudaStreamWaitEvent(st, p->ev_b[c * J + j], 0));
        RS_TRY(local_inv(c, lo(j), lo(j + 1) - lo(j)));
      }
    mark(p, 7, st);
    p->timings_valid = true;
    return BRI17_OK;
  }
  for (int c = 0; c < dim; c++) {
    RS_TRY(local_fwd(c, 0, p->n0_loc));
    BRI17_CUDA_TRY(cudaEventRecord(p->ev_a[c], st));
    BRI17_CUDA_TRY(cudaStreamWaitEvent(p->sx, p->ev_a[c], 0));
    RS_TRY(exchange_forward(p, l, T, X, nullptr, 1, p->sx, c, false, xt));
    RS_TRY(exchange_barrier(p));
    BRI17_CUDA_TRY(cudaEventRecord(p->ev_b[c], p->sx));
  }
  mark(p, 1, st);
  mark(p, 2, st);
  for (int c = 0; c < dim; c++) {
    BRI17_CUDA_TRY(cudaStreamWaitEvent(st, p->ev_b[c], 0));
    RS_TRY(fft_axis0(p, l, X + c * l.fourier_count, 1, CUFFT_FORWARD, st));
  }
  mark(p, 3, st);
  int kb[3] = {0, l.k1_beg[p->rank], 0};
  int ls[3] = {p->shape[0], l.n1_loc, l.S2e};
  if (l.fourier_count)
    RS_TRY(bri17_modal_stiffness_apply_f64(p->modal, X, X, kb, ls, 0, p->correction, st));
  mark(p, 4, st);
  for (int c = 0; c < dim; c++) {
    RS_TRY(fft_axis0(p, l, X + c * l.fourier_count, 1, CUFFT_INVERSE, st));
    BRI17_CUDA_TRY(cudaEventRecord(p->ev_a[c], st));
    BRI17_CUDA_TRY(cudaStreamWaitEvent(p->sx, p->ev_a[c], 0));
    RS_TRY(exchange_backward(p, l, X, p->W2, nullptr, 1, 1.0, p->sx, c, false));
    RS_TRY(exchange_barrier(p));
    BRI17_CUDA_TRY(cudaEventRecord(p->ev_b[c], p->sx));
  }
  mark(p, 5, st);
  mark(p, 6, st);
  for (int c = 0; c < dim; c++) {
    BRI17_CUDA_TRY(cudaStreamWaitEvent(st, p->ev_b[c], 0));
    RS_TRY(local_inv(c, 0, p->n0_loc));
  }
  mark(p, 7, st);
  p->timings_valid = true;
  return BRI17_OK;
}

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
  if (p->sx) cudaStreamDestroy(p->sx);
  for (auto &e : p->ev_a) if (e) cudaEventDestroy(e);
  for (auto &e : p->ev_b) if (e) cudaEventDestroy(e);
  destroy_layout(p->lc);
  destroy_layout(p->lr);
  for (void *ptr : {(void *)p->W, (void *)p->W2, (void *)p->barrier_word, (void *)p->cg_r, (void *)p->cg_p,
                    (void *)p->cg_Ap, (void *)p->cg_partial, (void *)p->cg_scalars, (void *)p->rbuf,
                    (void *)p->tabs, (void *)p->twiddle, (void *)p->dot_scratch})
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
  p->mu = mu;
  p->nu = nu;
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
  for (int q = 0; q <= nranks; q++) p->n0_beg[q] = int((int64_t)q * shape[0] / nranks);
  p->n0_loc = p->n0_beg[rank + 1] - p->n0_beg[rank];
  p->real_count = (int64_t)p->n0_loc * shape[1] * p->N2e;

  int rc = bri17_plan_create(&p->modal, dim, shape, L, mu, nu, device);
  if (rc) { delete p; return rc; }
  DeviceGuard guard(device);
  auto bail = [&](int code) { bri17_rs_plan_destroy(p); return code; };

  for (auto &e : p->ev)
    if (cudaEventCreate(&e) != cudaSuccess) return bail(fail(BRI17_ERR_CUDA, "cudaEventCreate failed"));
  layout_geometry(p, p->lc, false);
  layout_geometry(p, p->lr, true);
  // the real path keeps its local-transform data in the lower part of W2 and its packed
  // send/receive pieces (NCCL mode) above it
  p->real_upper = int64_t(dim) * std::max(p->lr.t_count, p->lr.fourier_count);
  if ((rc = setup_layout(p, p->lc))) return bail(rc);
  cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, device);

  // fused axis-0 pass: device copies of the host-built phi|chi|psi tables (bri17.hpp:259-263,
  // same libm values as the modal plan) and the twiddles of the axis-0 transform
  if (bri17b200::axis0::supported(shape[0])) {
    int64_t total = 0;
    for (int d = 0; d < dim; d++) { p->tab_off[d] = total; total += 3 * int64_t(shape[d]); }
    std::vector<double> host(total);
    for (int d = 0; d < dim; d++) {
      double *t = host.data() + p->tab_off[d];
      if ((rc = bri17_plan_get_tables(p->modal, d, t, t + shape[d], t + 2 * shape[d], nullptr, nullptr)))
        return bail(rc);
    }
    std::vector<double2> tw(shape[0]);
    bri17b200::axis0::fill_twiddles(shape[0], tw.data());
    if (cudaMalloc(&p->tabs, sizeof(double) * total) != cudaSuccess ||
        cudaMalloc(&p->twiddle, sizeof(double2) * shape[0]) != cudaSuccess ||
        cudaMemcpy(p->tabs, host.data(), sizeof(double) * total, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(p->twiddle, tw.data(), sizeof(double2) * shape[0], cudaMemcpyHostToDevice) != cudaSuccess)
      return bail(fail(BRI17_ERR_CUDA, "fused axis-0 tables: allocation/upload failed"));
  }

  if (nranks > 1) {
    ncclUniqueId id;
    std::memcpy(&id, nccl_unique_id, sizeof(id));
    ncclResult_t nr = ncclCommInitRank(&p->comm, nranks, id, rank);
    if (nr != ncclSuccess) return bail(fail(BRI17_ERR_NCCL, std::string("ncclCommInitRank: ") + ncclGetErrorString(nr)));
    // Same size on EVERY rank (the largest any rank needs): the flag page sits right behind the
    // data, and peers address it through their own copy of this number.
    size_t cap_elems = 1;
    for (int q = 0; q < nranks; q++) {
      const size_t n0q = size_t(p->n0_beg[q + 1] - p->n0_beg[q]);
      for (const Layout *l : {&p->lc, &p->lr}) {
        const size_t n1q = size_t(l->k1_beg[q + 1] - l->k1_beg[q]);
        const size_t t = n0q * l->S1 * l->S2e, f = size_t(shape[0]) * n1q * l->S2e;
        cap_elems = std::max(cap_elems, size_t(dim) * std::max(t, f) * (l->real ? 2 : 1));
      }
    }
    const size_t cap = sizeof(double2) * cap_elems;
    const size_t cap_pad = (cap + 255) & ~size_t(255);  // flag page behind the data, 256-byte aligned
    if (cudaMalloc(&p->W, cap_pad + FLAG_BYTES) != cudaSuccess || cudaMalloc(&p->W2, cap) != cudaSuccess ||
        cudaMalloc(&p->barrier_word, 256) != cudaSuccess)
      return bail(fail(BRI17_ERR_CUDA, "exchange buffer allocation failed"));
    p->buf_bytes = cap;
    cudaMemset(p->barrier_word, 0, 256);
    cudaMemset(reinterpret_cast<char *>(p->W) + cap_pad, 0, FLAG_BYTES);
    {  // exchange stream (high priority) and its events
      int lo = 0, hi = 0;
      cudaDeviceGetStreamPriorityRange(&lo, &hi);
      if (cudaStreamCreateWithPriority(&p->sx, cudaStreamNonBlocking, hi) != cudaSuccess)
        return bail(fail(BRI17_ERR_CUDA, "exchange stream creation failed"));
      for (int c = 0; c < 3 * MAX_XCHUNKS; c++)
        if (cudaEventCreateWithFlags(&p->ev_a[c], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&p->ev_b[c], cudaEventDisableTiming) != cudaSuccess)
          return bail(fail(BRI17_ERR_CUDA, "cudaEventCreate failed"));
    }
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
      p->flags.rank = rank;
      p->flags.nranks = nranks;
      for (int q = 0; q < nranks; q++)
        p->flags.page[q] = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(p->peerW[q]) + cap_pad);
      // nobody may signal before every rank has zeroed its page and mapped the others
      cudaDeviceSynchronize();
      nr = ncclAllReduce(p->barrier_word, p->barrier_word, 1, ncclDouble, ncclSum, p->comm, nullptr);
      if (nr != ncclSuccess || cudaDeviceSynchronize() != cudaSuccess)
        return bail(fail(BRI17_ERR_NCCL, "plan creation barrier failed"));
    }
  }
  *out = p;
  return BRI17_OK;
}

int bri17_rs_plan_local(const bri17_rs_plan *p, int *n0_begin, int *n0_count, int *k1_begin, int *k1_count) {
  if (!p) return fail(BRI17_ERR_INVALID_ARG, "plan is NULL");
  if (n0_begin) *n0_begin = p->n0_beg[p->rank];
  if (n0_count) *n0_count = p->n0_loc;
  if (k1_begin) *k1_begin = p->lc.k1_beg[p->rank];
  if (k1_count) *k1_count = p->lc.n1_loc;
  return BRI17_OK;
}
int64_t bri17_rs_plan_real_count(const bri17_rs_plan *p) { return p ? p->real_count : -1; }
int64_t bri17_rs_plan_fourier_count(const bri17_rs_plan *p) { return p ? p->lc.fourier_count : -1; }
bri17_plan *bri17_rs_plan_modal(bri17_rs_plan *p) { return p ? p->modal : nullptr; }

int64_t bri17_rs_plan_exchange_bytes(const bri17_rs_plan *p, int real_layout) {
  if (!p) return -1;
  const Layout *l = real_layout ? &p->lr : &p->lc;
  return int64_t(16) * p->dim * p->n0_loc * (l->S1 - l->n1_loc) * l->S2e;
}

int bri17_rs_forward_fft_f64(bri17_rs_plan *p, const void *x_dev, void *x_hat_dev, int ncomp, void *stream) {
  if (!p) return fail(BRI17_ERR_INVALID_ARG, "plan is NULL");
  // a rank without rows (or without k1 columns) passes NULL on that side and still takes part
  if ((p->lc.t_count > 0 && !x_dev) || (p->lc.fourier_count > 0 && !x_hat_dev))
    return fail(BRI17_ERR_INVALID_ARG, "NULL field on a rank that owns a non-empty slab");
  if (ncomp < 1) return fail(BRI17_ERR_INVALID_ARG, "ncomp < 1");
  DeviceGuard guard(p->device);
  cudaStream_t st = cudaStream_t(stream);
  if (!p->lc.ready) RS_TRY(setup_layout(p, p->lc));
  const Layout &l = p->lc;
  const double2 *x = static_cast<const double2 *>(x_dev);
  double2 *xh = static_cast<double2 *>(x_hat_dev);
  if (p->nranks == 1) {
    RS_TRY(fft_local_c2c(p, x, xh, ncomp, CUFFT_FORWARD, st));
    return fft_axis0(p, l, xh, ncomp, CUFFT_FORWARD, st);
  }
  for (int c0 = 0; c0 < ncomp; c0 += p->dim) {  // the exchange buffers hold dim components
    const int nc = std::min(p->dim, ncomp - c0);
    RS_TRY(fft_local_c2c(p, x + c0 * l.t_count, p->W2, nc, CUFFT_FORWARD, st));
    double2 *X = xh + c0 * l.fourier_count;
    if (p->mode == 1) {  // peers store into our W: stage through it
      RS_TRY(exchange_forward(p, l, p->W2, p->W, nullptr, nc, st));
      if (l.fourier_count)
        BRI17_CUDA_TRY(cudaMemcpyAsync(X, p->W, sizeof(double2) * nc * l.fourier_count, cudaMemcpyDeviceToDevice, st));
      // W is free again only once EVERY rank has copied its part out: the next writer of our W
      // (a peer's exchange kernel, e.g. of a pipelined apply that starts without a barrier) must
      // be ordered behind this copy
      RS_TRY(stream_barrier(p, st));
    } else {
      RS_TRY(exchange_forward(p, l, p->W2, X, p->W, nc, st));
    }
    RS_TRY(fft_axis0(p, l, X, nc, CUFFT_FORWARD, st));
  }
  return BRI17_OK;
}

int bri17_rs_inverse_fft_f64(bri17_rs_plan *p, void *x_hat_dev, void *x_dev, int ncomp, double scale, void *stream) {
  if (!p) return fail(BRI17_ERR_INVALID_ARG, "plan is NULL");
  if ((p->lc.t_count > 0 && !x_dev) || (p->lc.fourier_count > 0 && !x_hat_dev))
    return fail(BRI17_ERR_INVALID_ARG, "NULL field on a rank that owns a non-empty slab");
  if (ncomp < 1) return fail(BRI17_ERR_INVALID_ARG, "ncomp < 1");
  DeviceGuard guard(p->device);
  cudaStream_t st = cudaStream_t(stream);
  if (!p->lc.ready) RS_TRY(setup_layout(p, p->lc));
  const Layout &l = p->lc;
  double2 *x = static_cast<double2 *>(x_dev);
  double2 *xh = static_cast<double2 *>(x_hat_dev);
  if (p->nranks == 1) {
    RS_TRY(fft_axis0(p, l, xh, ncomp, CUFFT_INVERSE, st));
    RS_TRY(fft_local_c2c(p, xh, x, ncomp, CUFFT_INVERSE, st));
    if (scale != 1.0 && l.t_count)
      scale_kernel<<<1184, 256, 0, st>>>(x, (long long)ncomp * l.t_count, scale);
    return BRI17_OK;
  }
  for (int c0 = 0; c0 < ncomp; c0 += p->dim) {
    const int nc = std::min(p->dim, ncomp - c0);
    double2 *X = xh + c0 * l.fourier_count, *D = x + c0 * l.t_count;
    RS_TRY(fft_axis0(p, l, X, nc, CUFFT_INVERSE, st));
    if (p->mode == 1) {
      RS_TRY(exchange_backward(p, l, X, p->W2, nullptr, nc, scale, st));
      RS_TRY(fft_local_c2c(p, p->W2, D, nc, CUFFT_INVERSE, st));
    } else {
      RS_TRY(exchange_backward(p, l, X, D, p->W2, nc, scale, st));
      RS_TRY(fft_local_c2c(p, D, D, nc, CUFFT_INVERSE, st));
    }
  }
  return BRI17_OK;
}

}  // extern "C"

namespace {

// A rank whose slab is empty (shape[0] < nranks) passes NULL fields but still takes part in every
// exchange and barrier, otherwise the other ranks would wait for it forever.
int check_fields(const bri17_rs_plan *p, const void *a, const void *b) {
  if (!p) return fail(BRI17_ERR_INVALID_ARG, "plan is NULL");
  if (p->real_count > 0 && (!a || !b)) return fail(BRI17_ERR_INVALID_ARG, "NULL field on a rank that owns a non-empty slab");
  return BRI17_OK;
}

int apply_complex(bri17_rs_plan *p, const void *u_dev, void *F_dev, cudaStream_t st, double *dot_dev) {
  Layout &l = p->lc;
  if (!l.ready) RS_TRY(setup_layout(p, l));
  const double2 *u = static_cast<const double2 *>(u_dev);
  double2 *F = static_cast<double2 *>(F_dev);
  const int dim = p->dim;
  const bool xt = use_xt(p, l);
  if (p->nranks > 1 && p->mode == 1 && p->pipeline) {
    auto fwd = [&](int c, int a0, int na) {
      return batch_exec_range(l.fwd_local, u + c * l.t_count, F + c * l.t_count, a0, na, CUFFT_FORWARD, st);
    };
    auto inv = [&](int c, int a0, int na) {
      return batch_exec_range(l.fwd_local, p->W2 + c * l.t_count, F + c * l.t_count, a0, na, CUFFT_INVERSE, st);
    };
    return apply_pipelined(p, l, F, fwd, inv, st, dot_dev);
  }
  p->timings_valid = false;
  if (p->nranks == 1 && xt) {
    // one GPU, k1-major spectral layout: u -> (cuFFT, transposed output) W2 -> fused pass in place
    // -> (cuFFT, transposed input) F
    RS_TRY(setup_layout_t(p, l));
    RS_TRY(ensure_buffers(p, std::max<size_t>(sizeof(double2) * dim * l.t_count, 16)));
    double2 *X = p->W2;
    mark(p, 0, st);
    for (int c = 0; c < dim; c++)
      RS_TRY(batch_exec(l.fwd_local_t, u + c * l.t_count, X + c * l.t_count, CUFFT_FORWARD, st));
    mark(p, 1, st);
    mark(p, 2, st);
    RS_TRY(modal_section(p, l, X, st, dot_dev, true));
    mark(p, 6, st);
    for (int c = 0; c < dim; c++)
      RS_TRY(batch_exec(l.inv_local_t, X + c * l.t_count, F + c * l.t_count, CUFFT_INVERSE, st));
    mark(p, 7, st);
    p->timings_valid = true;
    return BRI17_OK;
  }
  mark(p, 0, st);
  RS_TRY(fft_local_c2c(p, u, F, dim, CUFFT_FORWARD, st));                   // :57 (axes 1..)
  mark(p, 1, st);
  double2 *X = F;  // Fourier-side block
  if (p->nranks > 1) {
    X = p->W;
    RS_TRY(exchange_forward(p, l, F, p->W, p->W2, dim, st, 0, true, xt));
  }
  mark(p, 2, st);
  RS_TRY(modal_section(p, l, X, st, dot_dev, xt));                          // :57 (axis 0), :58-106, :95 (axis 0)
  if (p->nranks > 1) {
    if (p->mode == 1) {
      RS_TRY(exchange_backward(p, l, X, p->W2, nullptr, dim, 1.0, st, 0, true, xt));
      mark(p, 6, st);
      RS_TRY(fft_local_c2c(p, p->W2, F, dim, CUFFT_INVERSE, st));
    } else {
      RS_TRY(exchange_backward(p, l, X, F, p->W2, dim, 1.0, st));
      mark(p, 6, st);
      RS_TRY(fft_local_c2c(p, F, F, dim, CUFFT_INVERSE, st));
    }
  } else {
    mark(p, 6, st);
    RS_TRY(fft_local_c2c(p, F, F, dim, CUFFT_INVERSE, st));
  }
  mark(p, 7, st);
  p->timings_valid = true;
  return BRI17_OK;
}

// Real fields (plain doubles, [dim][n0_count][N1][(N2)]): r2c over the trailing axes, half
// spectrum everywhere in between, c2r back.
int apply_real(bri17_rs_plan *p, const void *u_dev, void *F_dev, cudaStream_t st, double *dot_dev) {
  Layout &l = p->lr;
  if (!l.ready) RS_TRY(setup_layout(p, l));
  const int dim = p->dim;
  if (p->nranks == 1) RS_TRY(ensure_buffers(p, std::max<size_t>(sizeof(double2) * dim * l.t_count, 16)));
  const double *u = static_cast<const double *>(u_dev);
  double *F = static_cast<double *>(F_dev);
  double2 *T = p->W2;  // local-transform layout [c][n0_loc][S1][S2e] ([c][S1][n0][S2e] when k1-major on one GPU)
  const bool xt = use_xt(p, l);
  const bool local_t = xt && p->nranks == 1;
  if (local_t) RS_TRY(setup_layout_t(p, l));
  // cuFFT wants 16-byte aligned real arrays; component c starts at c*real_count doubles, which
  // is misaligned when the slab holds an odd number of values: stage those through rbuf.
  if ((p->real_count & 1) && !p->rbuf) BRI17_CUDA_TRY(cudaMalloc(&p->rbuf, sizeof(double) * (p->real_count + 2)));
  auto misaligned = [](const void *ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) != 0; };
  const long long rplane = (long long)p->shape[1] * p->N2e;  // doubles per real-space plane
  auto local_fwd = [&](int c, int a0, int na) -> int {  // D2Z planes [a0, a0 + na) of u_c -> T_c
    if (!l.have_local || na <= 0) return BRI17_OK;
    BatchFft &plan = local_t ? l.fwd_local_t : l.fwd_local;
    const double *src = u + c * p->real_count + a0 * rplane;
    if (misaligned(src)) {
      BRI17_CUDA_TRY(cudaMemcpyAsync(p->rbuf, src, sizeof(double) * na * rplane, cudaMemcpyDeviceToDevice, st));
      src = p->rbuf;
    }
    return batch_exec_at(plan, src, T + c * l.t_count + a0 * plan.odist, na, 0, st);
  };
  auto local_inv = [&](int c, int a0, int na) -> int {  // Z2D planes [a0, a0 + na) of (W2)_c -> F_c
    if (!l.have_local || na <= 0) return BRI17_OK;
    BatchFft &plan = local_t ? l.inv_local_t : l.inv_local;
    double *dst = F + c * p->real_count + a0 * rplane;
    double *out = misaligned(dst) ? p->rbuf : dst;
    RS_TRY(batch_exec_at(plan, p->W2 + c * l.t_count + a0 * plan.idist, out, na, 0, st));
    if (out != dst)
      BRI17_CUDA_TRY(cudaMemcpyAsync(dst, out, sizeof(double) * na * rplane, cudaMemcpyDeviceToDevice, st));
    return BRI17_OK;
  };
  // the whole slab of component c, in the plan's chunks (one chunk unless "fft_chunk_mib" is set)
  auto whole = [&](auto &fn, BatchFft &plan, int c) -> int {
    const int step = plan.made ? plan.chunk : std::max(p->n0_loc, 1);
    for (int a0 = 0; a0 < p->n0_loc; a0 += step) RS_TRY(fn(c, a0, std::min(step, p->n0_loc - a0)));
    return BRI17_OK;
  };
  if (p->nranks > 1 && p->mode == 1 && p->pipeline) return apply_pipelined(p, l, T, local_fwd, local_inv, st, dot_dev);

  p->timings_valid = false;
  mark(p, 0, st);
  for (int c = 0; c < dim; c++) RS_TRY(whole(local_fwd, local_t ? l.fwd_local_t : l.fwd_local, c));
  mark(p, 1, st);
  double2 *X = T;
  if (p->nranks > 1) {
    X = p->W;
    double2 *S = p->W2 + p->real_upper;  // packed send pieces (NCCL mode)
    RS_TRY(exchange_forward(p, l, T, p->W, S, dim, st, 0, true, xt));
  }
  mark(p, 2, st);
  RS_TRY(modal_section(p, l, X, st, dot_dev, xt));
  if (p->nranks > 1) {
    if (p->mode == 1) {
      RS_TRY(exchange_backward(p, l, X, p->W2, nullptr, dim, 1.0, st, 0, true, xt));   // peers store into our W2
    } else {
      double2 *R = p->W2 + p->real_upper;                                 // packed receive pieces
      RS_TRY(exchange_backward(p, l, X, p->W2, R, dim, 1.0, st));
    }
  }
  mark(p, 6, st);
  for (int c = 0; c < dim; c++) RS_TRY(whole(local_inv, local_t ? l.inv_local_t : l.inv_local, c));
  mark(p, 7, st);
  p->timings_valid = true;
  return BRI17_OK;
}

}  // namespace

extern "C" {

int bri17_real_space_apply_f64(bri17_rs_plan *p, const void *u_dev, void *F_dev, void *stream) {
  RS_TRY(check_fields(p, u_dev, F_dev));
  if (u_dev && u_dev == F_dev) return fail(BRI17_ERR_INVALID_ARG, "u_dev and F_dev must be distinct (F is scratch)");
  DeviceGuard guard(p->device);
  return apply_complex(p, u_dev, F_dev, cudaStream_t(stream), nullptr);
}

// Same operator as bri17_real_space_apply_f64 restricted to real input (what the reference
// always feeds it, tests/test_bri17.cpp:133-136).
int bri17_real_space_apply_real_f64(bri17_rs_plan *p, const void *u_dev, void *F_dev, void *stream) {
  RS_TRY(check_fields(p, u_dev, F_dev));
  DeviceGuard guard(p->device);
  return apply_real(p, u_dev, F_dev, cudaStream_t(stream), nullptr);
}

int bri17_rs_plan_set_option(bri17_rs_plan *p, const char *key, int64_t value) {
  if (!p || !key) return fail(BRI17_ERR_INVALID_ARG, "plan/key is NULL");
  if (!std::strcmp(key, "pipeline")) p->pipeline = value != 0;
  else if (!std::strcmp(key, "fused_axis0")) p->fused = value != 0;
  else if (!std::strcmp(key, "k1_major")) p->xt = value < 0 ? -1 : (value != 0);
  else if (!std::strcmp(key, "exchange_chunks")) {
    if (value < 0 || value > MAX_XCHUNKS) return fail(BRI17_ERR_INVALID_ARG, "exchange_chunks must be 0 (auto) .. 4");
    p->xchunks = int(value);
  } else if (!std::strcmp(key, "fft_chunk_mib") || !std::strcmp(key, "fft_chunk_planes")) {
    if (value < 0) return fail(BRI17_ERR_INVALID_ARG, std::string(key) + " < 0");
    DeviceGuard guard(p->device);
    cudaDeviceSynchronize();
    (key[10] == 'm' ? p->fft_chunk_mib : p->fft_chunk_planes) = int(std::min<int64_t>(value, 1 << 20));
    destroy_layout(p->lc);   // plans are rebuilt with the new chunk on next use
    destroy_layout(p->lr);
  } else if (!std::strcmp(key, "copy_ctas")) {
    if (value < 0) return fail(BRI17_ERR_INVALID_ARG, "copy_ctas < 0");
    p->copy_ctas = int(value);
  } else return fail(BRI17_ERR_INVALID_ARG, std::string("unknown option ") + key);
  return BRI17_OK;
}

int bri17_rs_plan_get_info(const bri17_rs_plan *p, const char *key, int64_t *value) {
  if (!p || !key || !value) return fail(BRI17_ERR_INVALID_ARG, "NULL argument");
  if (!std::strcmp(key, "fused_axis0")) *value = use_fused(p) ? 1 : 0;
  else if (!std::strcmp(key, "k1_major")) *value = use_xt(p, p->lc) ? 1 : 0;
  else if (!std::strcmp(key, "k1_major_real")) *value = use_xt(p, p->lr) ? 1 : 0;
  else if (!std::strcmp(key, "fused_launches")) *value = p->fused_launches;
  else if (!std::strcmp(key, "pipeline")) *value = (p->nranks > 1 && p->mode == 1 && p->pipeline) ? 1 : 0;
  else if (!std::strcmp(key, "exchange_mode")) *value = p->mode;
  else if (!std::strcmp(key, "fft_chunk_mib")) *value = p->fft_chunk_mib;
  else if (!std::strcmp(key, "exchange_chunks")) *value = exchange_chunks(p, p->lc);
  else if (!std::strcmp(key, "exchange_chunks_real")) *value = exchange_chunks(p, p->lr);
  else if (!std::strcmp(key, "fft_chunk_planes")) *value = p->lc.fwd_local.made ? p->lc.fwd_local.chunk : 0;
  else if (!std::strcmp(key, "barriers")) *value = int64_t(p->epoch_bar[0] + p->epoch_bar[1]);
  else return fail(BRI17_ERR_INVALID_ARG, std::string("unknown info key ") + key);
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
  RS_TRY(check_fields(p, b_dev, x_dev));
  if (max_iter < 0) return fail(BRI17_ERR_INVALID_ARG, "max_iter < 0");
  DeviceGuard guard(p->device);
  cudaStream_t st = cudaStream_t(stream);
  // interleaved complex: dim components of real_count (re, im) pairs
  return cg_core(p, [&](const double *d, double *Ad, double *dot) { return apply_complex(p, d, Ad, st, dot); },
                 static_cast<const double *>(b_dev), static_cast<double *>(x_dev), p->real_count, 2, rtol,
                 max_iter, check_every, iterations, rel_residual, st);
}

int bri17_cg_solve_real_f64(bri17_rs_plan *p, const void *b_dev, void *x_dev, double rtol, int max_iter,
                            int check_every, int *iterations, double *rel_residual, void *stream) {
  RS_TRY(check_fields(p, b_dev, x_dev));
  if (max_iter < 0) return fail(BRI17_ERR_INVALID_ARG, "max_iter < 0");
  DeviceGuard guard(p->device);
  cudaStream_t st = cudaStream_t(stream);
  return cg_core(p, [&](const double *d, double *Ad, double *dot) { return apply_real(p, d, Ad, st, dot); },
                 static_cast<const double *>(b_dev), static_cast<double *>(x_dev), p->real_count, 1, rtol,
                 max_iter, check_every, iterations, rel_residual, st);
}

// <u, A u> together with F = A u (what one CG iteration consumes): *dot_host receives the GLOBAL
// value (summed over ranks).  real_fields = 1: plain double fields through the r2c path.
int bri17_real_space_apply_dot_f64(bri17_rs_plan *p, const void *u_dev, void *F_dev, int real_fields,
                                   double *dot_host, void *stream) {
  RS_TRY(check_fields(p, u_dev, F_dev));
  if (!dot_host) return fail(BRI17_ERR_INVALID_ARG, "dot_host is NULL");
  DeviceGuard guard(p->device);
  cudaStream_t st = cudaStream_t(stream);
  if (!p->cg_scalars) {
    BRI17_CUDA_TRY(cudaMalloc(&p->cg_partial, sizeof(double) * std::max(RED_CTAS, 8 * MEAN_CTAS)));
    BRI17_CUDA_TRY(cudaMalloc(&p->cg_scalars, sizeof(double) * 16));
  }
  double *slot = p->cg_scalars + 2;
  RS_TRY(real_fields ? apply_real(p, u_dev, F_dev, st, slot) : apply_complex(p, u_dev, F_dev, st, slot));
  RS_TRY(scalar_allreduce(p, slot, 1, st));
  BRI17_CUDA_TRY(cudaMemcpyAsync(dot_host, slot, sizeof(double), cudaMemcpyDeviceToHost, st));
  BRI17_CUDA_TRY(cudaStreamSynchronize(st));
  return BRI17_OK;
}

// Test aid, HOST memory, no device needed: replays the fused axis-0 kernel thread by thread on
// the CPU.  X: [dim][N0][S] complex, in place.  tabs: phi|chi|psi of axis d at tabs_d ([3][N_d]).
int bri17_debug_axis0_fused_host(int dim, int N0, int64_t S, int S2e, int k1_begin, int N1, int N2,
                                 const double *tab0, const double *tab1, const double *tab2, double mu,
                                 double nu, double out_scale, int hermitian_n, int k1_major, void *X_host,
                                 double *dot_out) {
  if (!X_host || !tab0 || !tab1 || (dim == 3 && !tab2)) return fail(BRI17_ERR_INVALID_ARG, "NULL argument");
  if (dim != 2 && dim != 3) return fail(BRI17_ERR_INVALID_ARG, "dim must be 2 or 3");
  if (!bri17b200::axis0::supported(N0)) return fail(BRI17_ERR_UNSUPPORTED, "unsupported N0");
  std::vector<double2> tw(N0);
  bri17b200::axis0::fill_twiddles(N0, tw.data());
  bri17b200::axis0::Params a{};
  a.X = static_cast<double2 *>(X_host);
  a.comp_stride = int64_t(N0) * S;
  a.S = S;
  if (k1_major && dim == 3) {
    a.row_stride = S2e;
    a.blk_cols = S2e;
    a.blk_stride = int64_t(N0) * S2e;
  } else {
    a.row_stride = S;
    a.blk_cols = std::max<int64_t>(S, 1);
    a.blk_stride = 0;
  }
  a.N0 = N0;
  a.S2e = dim == 3 ? S2e : 1;
  a.k1_begin = k1_begin;
  a.tab0 = tab0; a.tab1 = tab1; a.tab2 = tab2;
  a.N1 = N1; a.N2 = N2;
  a.twiddle = tw.data();
  a.mu = mu;
  a.scaling = mu / (1. - 2. * nu);
  a.out_scale = out_scale;
  a.herm_n = hermitian_n;
  return bri17b200::axis0::emulate(a, dim, dot_out);
}

// Test aid, HOST memory, no device needed: replays the fused exchange (mode 1) of `nranks` virtual ranks
// on the CPU with the very CopyPlans the GPU path builds (exchange_forward / exchange_backward), executed by
// a plain loop over their documented semantics.  direction 0: in[r] = local-transform layout of rank r,
// [dim][n0_r][S1][S2e] -> out[q] = Fourier-side block of rank q ([dim][N0][n1_q][S2e], or k1-major
// [dim][n1_q][N0][S2e]); direction 1: the way back.  real_layout: half spectrum of the last axis.
// nchunks: sub-slabs per component, as in the pipelined apply.  Buffers hold complex128.
int bri17_debug_exchange_host(int dim, const int *shape, int nranks, int real_layout, int k1_major, int nchunks,
                              int direction, void *const *in, void *const *out) {
  if (!shape || !in || !out) return fail(BRI17_ERR_INVALID_ARG, "NULL argument");
  if ((dim != 2 && dim != 3) || nranks < 1 || nranks > MAX_RANKS || nchunks < 1 || nchunks > MAX_XCHUNKS)
    return fail(BRI17_ERR_INVALID_ARG, "bad dim / nranks / nchunks");
  auto run = [](const CopyPlan &cp) {
    for (int si = 0; si < cp.nseg; si++) {
      const CopySeg &g = cp.seg[si];
      for (int c = 0; c < cp.ncomp; c++)
        for (int a = 0; a < g.rows; a++)
          for (int e = 0; e < g.len; e++) {
            const long long so = (long long)(e / g.inner) * g.src_bs + e % g.inner;
            const long long dd = (long long)(e / g.inner) * g.dst_bs + e % g.inner;
            const double2 v = g.src[c * g.src_cs + a * g.src_rs + so];
            g.dst[c * g.dst_cs + a * g.dst_rs + dd] = make_double2(v.x * cp.scale, v.y * cp.scale);
          }
    }
  };
  for (int r = 0; r < nranks; r++) {
    bri17_rs_plan p;  // geometry only: no CUDA object is created or touched
    p.dim = dim;
    for (int d = 0; d < dim; d++) p.shape[d] = shape[d];
    p.rank = r;
    p.nranks = nranks;
    p.mode = 1;
    p.N2e = dim == 3 ? shape[2] : 1;
    for (int q = 0; q <= nranks; q++) p.n0_beg[q] = int((int64_t)q * shape[0] / nranks);
    p.n0_loc = p.n0_beg[r + 1] - p.n0_beg[r];
    Layout l;
    layout_geometry(&p, l, real_layout != 0);
    const bool xt = k1_major && dim == 3;
    for (int q = 0; q < nranks; q++) {
      p.peerW[q] = static_cast<double2 *>(direction == 0 ? out[q] : in[q]);
      p.peerW2[q] = static_cast<double2 *>(direction == 0 ? in[q] : out[q]);
    }
    for (int c = 0; c < dim; c++)
      for (int j = 0; j < nchunks; j++) {
        CopyPlan cp{};
        if (direction == 0) {
          const int a0 = int((long long)p.n0_loc * j / nchunks), a1 = int((long long)p.n0_loc * (j + 1) / nchunks);
          RS_TRY(exchange_forward(&p, l, p.peerW2[r], p.peerW[r], nullptr, 1, nullptr, c, false, xt,
                                  nchunks > 1 ? a0 : 0, nchunks > 1 ? a1 - a0 : -1, &cp));
        } else {
          RS_TRY(exchange_backward(&p, l, p.peerW[r], p.peerW2[r], nullptr, 1, 1.0, nullptr, c, false, xt, j, nchunks, &cp));
        }
        run(cp);
      }
  }
  return BRI17_OK;
}

}  // extern "C"
