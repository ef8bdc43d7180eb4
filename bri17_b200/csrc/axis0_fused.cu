// axis0_fused.cu -- kernel wrapper, launcher and CPU replay of the fused
// FFT(axis 0) -> K^ -> inverse FFT(axis 0) pass (see axis0_fused.cuh).
#include "axis0_fused.cuh"

#include <algorithm>
#include <string>
#include <vector>

#include "internal.h"

namespace bri17b200 {
namespace axis0 {

template <class C, int DIM>
__global__ void __launch_bounds__(C::THREADS, C::MINB) axis0_fused_kernel(const Params p) {
  extern __shared__ __align__(16) unsigned char a0_raw[];
  double2 *tw = reinterpret_cast<double2 *>(a0_raw);
  double2 *data = tw + C::N0;
  const int tid = threadIdx.x;
  for (int i = tid; i < C::N0; i += C::THREADS) tw[i] = p.twiddle[i];
  __syncthreads();
  double dot = 0.;
  if (blockIdx.x < p.n_tiles) issue_tile_loads<C, DIM>(tid, data, p, (long long)blockIdx.x * C::W);
  for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
    const long long col0 = tile * C::W;
    const long long nx = tile + gridDim.x < p.n_tiles ? (tile + gridDim.x) * C::W : -1;
    phase<C, DIM, 0>(tid, data, tw, p, col0, nx, dot);  // waits for this thread's async copies
    __syncthreads();
    phase<C, DIM, 1>(tid, data, tw, p, col0, nx, dot);
    if constexpr (C::NPH > 3) {
      // phases 1..3 of a warp touch only that warp's 64-row block (Cfg::WARP_LOCAL)
      if constexpr (C::WARP_LOCAL) __syncwarp(); else __syncthreads();
      phase<C, DIM, 2>(tid, data, tw, p, col0, nx, dot);
      if constexpr (C::WARP_LOCAL) __syncwarp(); else __syncthreads();
      phase<C, DIM, 3>(tid, data, tw, p, col0, nx, dot);
      __syncthreads();
      phase<C, DIM, 4>(tid, data, tw, p, col0, nx, dot);
    } else {
      __syncthreads();
      phase<C, DIM, 2>(tid, data, tw, p, col0, nx, dot);
    }
    // no barrier here: the next tile's first stage works on the slots this thread just used
  }
  if (p.dot_partial) {  // deterministic CTA sum -> one partial per CTA
    __shared__ double sh[C::THREADS / 32];
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if ((tid & 31) == 0) sh[tid >> 5] = dot;
    __syncthreads();
    if (tid == 0) {
      double t = 0.;
      for (int i = 0; i < C::THREADS / 32; i++) t += sh[i];
      p.dot_partial[blockIdx.x] = t;
    }
  }
}

template <class C>
static int launch_cfg(const Params &p, int dim, int sm_count, int max_grid, cudaStream_t st, int *grid_out) {
  void (*kern)(const Params) = dim == 3 ? axis0_fused_kernel<C, 3> : axis0_fused_kernel<C, 2>;
  const size_t smem = C::smem_bytes(dim);
  BRI17_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  int occ = 0;
  BRI17_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, C::THREADS, smem));
  if (occ < 1) return fail(BRI17_ERR_CUDA, "fused axis-0 kernel does not fit on an SM");
  long long grid = std::min<long long>(p.n_tiles, (long long)sm_count * occ);
  if (max_grid > 0) grid = std::min<long long>(grid, max_grid);
  if (grid < 1) grid = 1;
  kern<<<unsigned(grid), C::THREADS, smem, st>>>(p);
  if (grid_out) *grid_out = int(grid);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(BRI17_ERR_CUDA, std::string("axis0_fused launch: ") + cudaGetErrorString(e));
  return BRI17_OK;
}

int columns_per_tile(int N0) {
  switch (N0) {
    case 16: return Cfg16::W;
    case 32: return Cfg32::W;
    case 64: return Cfg64::W;
    case 128: return Cfg128::W;
    case 256: return Cfg256::W;
    case 512: return Cfg512::W;
    case 1024: return Cfg1024::W;
  }
  return 0;
}

// Launches the fused pass on `st`.  p.n_tiles is filled here.  max_grid > 0 caps the grid (one dot
// partial per CTA).  *grid_out receives the grid size.
int launch(Params p, int dim, int sm_count, int max_grid, cudaStream_t st, int *grid_out) {
  const int W = columns_per_tile(p.N0);
  if (!W) return fail(BRI17_ERR_UNSUPPORTED, "fused axis-0 pass: unsupported N0");
  p.n_tiles = (p.S + W - 1) / W;
  if (p.n_tiles == 0) { if (grid_out) *grid_out = 0; return BRI17_OK; }
  if (p.S >= (1ll << 31) || p.blk_cols >= (1ll << 31))
    return fail(BRI17_ERR_UNSUPPORTED, "fused axis-0 pass: more than 2^31 columns");
  switch (p.N0) {
    case 16: return launch_cfg<Cfg16>(p, dim, sm_count, max_grid, st, grid_out);
    case 32: return launch_cfg<Cfg32>(p, dim, sm_count, max_grid, st, grid_out);
    case 64: return launch_cfg<Cfg64>(p, dim, sm_count, max_grid, st, grid_out);
    case 128: return launch_cfg<Cfg128>(p, dim, sm_count, max_grid, st, grid_out);
    case 256: return launch_cfg<Cfg256>(p, dim, sm_count, max_grid, st, grid_out);
    case 512: return launch_cfg<Cfg512>(p, dim, sm_count, max_grid, st, grid_out);
    case 1024: return launch_cfg<Cfg1024>(p, dim, sm_count, max_grid, st, grid_out);
  }
  return fail(BRI17_ERR_UNSUPPORTED, "fused axis-0 pass: unsupported N0");
}

// exp(-2 pi i q / N0) for 0 <= q < N0, evaluated in the first octant and mirrored (libm, host).
static double2 unit_root(int N0, int q) {
  const double two_pi = 6.283185307179586476925286766559;
  const int oct = (8 * q) / N0;  // octant 0..7
  double c, s, cc, ss;
  auto cs = [&](int jj, double &co, double &si) { const double a = two_pi * jj / N0; co = std::cos(a); si = std::sin(a); };
  switch (oct) {
    case 0: cs(q, c, s); break;
    case 1: cs(N0 / 4 - q, cc, ss); c = ss; s = cc; break;
    case 2: cs(q - N0 / 4, cc, ss); c = -ss; s = cc; break;
    case 3: cs(N0 / 2 - q, cc, ss); c = -cc; s = ss; break;
    case 4: cs(q - N0 / 2, cc, ss); c = -cc; s = -ss; break;
    case 5: cs(3 * N0 / 4 - q, cc, ss); c = -ss; s = -cc; break;
    case 6: cs(q - 3 * N0 / 4, cc, ss); c = ss; s = -cc; break;
    default: cs(N0 - q, cc, ss); c = cc; s = -ss; break;
  }
  return make_double2(c, -s);
}

template <class C>
static void fill_cfg(double2 *tw) {
  for (int i = 0; i < C::N0; i++) tw[i] = make_double2(1., 0.);
  constexpr int s0 = C::N0 / C::R0;  // stage 0: block N0, stride s0, twiddle w_N0^(j m)
  for (int m = 1; m < C::R0; m++)
    for (int j = 0; j < s0; j++) tw[(m - 1) * s0 + j] = unit_root(C::N0, j * m);
  if constexpr (C::NS == 3) {  // stage 1: block N0/R0, stride s1, twiddle w_{N0/R0}^(j m) = w_N0^(j m R0)
    constexpr int s1 = C::N0 / (C::R0 * C::R1);
    static_assert(C::TW1 + (C::R1 - 1) * s1 <= C::N0, "twiddle table overflows N0 entries");
    for (int m = 1; m < C::R1; m++)
      for (int j = 0; j < s1; j++) tw[C::TW1 + (m - 1) * s1 + j] = unit_root(C::N0, j * m * C::R0);
  }
}

// Per-stage twiddle tables of the plan for N0 (layout: Cfg), N0 complex in all.
void fill_twiddles(int N0, double2 *tw) {
  switch (N0) {
    case 16: fill_cfg<Cfg16>(tw); break;
    case 32: fill_cfg<Cfg32>(tw); break;
    case 64: fill_cfg<Cfg64>(tw); break;
    case 128: fill_cfg<Cfg128>(tw); break;
    case 256: fill_cfg<Cfg256>(tw); break;
    case 512: fill_cfg<Cfg512>(tw); break;
    case 1024: fill_cfg<Cfg1024>(tw); break;
  }
}

// CPU replay (tests only): same phase code, threads run one after the other.
int emulate(const Params &p_in, int dim, double *dot_out) {
  Params p = p_in;
  const int W = columns_per_tile(p.N0);
  if (!W) return fail(BRI17_ERR_UNSUPPORTED, "fused axis-0 pass: unsupported N0");
  p.n_tiles = (p.S + W - 1) / W;
  double dummy = 0.;
  if (dot_out) p.dot_partial = &dummy;  // non-NULL switches the accumulation on
#define BRI17_A0_CASE(N, C)                                        \
  case N:                                                          \
    if (dim == 3) emulate_host<C, 3>(p, dot_out);                  \
    else emulate_host<C, 2>(p, dot_out);                           \
    return BRI17_OK;
  switch (p.N0) {
    BRI17_A0_CASE(16, Cfg16)
    BRI17_A0_CASE(32, Cfg32)
    BRI17_A0_CASE(64, Cfg64)
    BRI17_A0_CASE(128, Cfg128)
    BRI17_A0_CASE(256, Cfg256)
    BRI17_A0_CASE(512, Cfg512)
    BRI17_A0_CASE(1024, Cfg1024)
  }
#undef BRI17_A0_CASE
  return fail(BRI17_ERR_UNSUPPORTED, "fused axis-0 pass: unsupported N0");
}

}  // namespace axis0
}  // namespace bri17b200
