// internal.h -- shared declarations of libbri17_b200.so (not installed).
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <mutex>
#include <string>
#include <vector>

#include "bri17_b200.h"

namespace bri17b200 {

// ---- error plumbing ---------------------------------------------------------
void set_error(const std::string &msg);
int fail(int code, const std::string &msg);

#define BRI17_CUDA_TRY(expr)                                                        \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess)                                                          \
      return ::bri17b200::fail(BRI17_ERR_CUDA, std::string(#expr) + ": " +          \
                                                   cudaGetErrorString(_e));         \
  } while (0)

// ---- per-axis tables ----------------------------------------------------------
// Device layout of one axis: seven arrays of N doubles, back to back:
//   phi | chi | psi | c | s | alpha | sin(alpha)   (bri17.hpp:261-263, :220-221 and :218)
// sin(alpha) (unscaled, unlike s = sin(alpha)*N/L) lets the device build the prefactor
// e^{i sum(alpha)} of B^ (:224) as a product of per-axis (cos, sin) pairs instead of a sincos.
enum { TAB_PHI = 0, TAB_CHI = 1, TAB_PSI = 2, TAB_C = 3, TAB_S = 4, TAB_ALPHA = 5, TAB_SINA = 6, TAB_COUNT = 7 };

struct AxisTables {
  std::vector<double> host;  // TAB_COUNT * n
  double *dev = nullptr;     // same layout
  int n = 0;
  const double *h(int which) const { return host.data() + size_t(which) * n; }
};

// A row-major block of frequencies, normalised to 3 memory dimensions
// (2-D grids get a leading dimension of extent 1 ... see make_block()).
struct Block {
  int dim;
  int n[3];         // local extents, slowest -> fastest (dim entries used)
  int kb[3];        // first frequency along each axis
  int64_t modes;    // prod(n)
};

// Everything a kernel needs to walk a block tile by tile.  Passed by value.
struct TileGeom {
  long long n_rows;    // product of all extents but the fastest
  long long n_tiles;   // n_rows * cpr
  int n_inner;         // fastest extent
  int n_mid;           // 3-D: extent of the middle axis (row = a*n_mid + b); 2-D: 1
  int n_outer;         // extent of the slowest axis
  int cpr;             // tiles ("chunks") per row
  int kb_outer;        // k_begin of the slowest axis
  int kb_mid;          // k_begin of the middle axis (3-D)
  int kb_inner;        // k_begin of the fastest axis
  // per-iteration increments of a persistent CTA (grid stride = gridDim.x tiles)
  long long d_row;     // gridDim.x / cpr
  int d_chunk;         // gridDim.x % cpr
  int d_a, d_b;        // d_row / n_mid, d_row % n_mid
};

struct ApplyParams {
  const double2 *u;
  double2 *f;
  long long u_stride;   // complex elements between input components
  long long f_stride;   // ... between output components
  long long u_mstride;  // ... between consecutive modes of the input (1 = planar); solve kernels only
  long long f_mstride;  // ... of the output
  TileGeom g;
  const double *tab_outer;  // device tables of the slowest axis
  const double *tab_mid;    // middle axis (3-D only)
  const double *tab_inner;  // fastest axis
  int N_outer, N_mid, N_inner;  // table lengths (global shape)
  double mu, scaling, out_scale;
  int stage_outer;      // outer/mid tables of the local range are staged in smem
  // dot variant of the flat kernel only: per-CTA partial sums of w_k Re(u^_k^H f^_k)
  double *dot_partial;
  int herm_n;           // > 0: the fastest axis is the half spectrum of a real field of this
                        // length; modes with 0 < k < N/2 stand for a conjugate pair (w_k = 2)
};

struct Variant {
  const char *name;
  int threads;
  int vec;           // modes per thread per tile
  int min_blocks;    // __launch_bounds__ minBlocksPerSM
  int load_hint;     // 0 default, 1 ld.global.cs, 2 ld.global.nc, 3 nc + L1::no_allocate
  int store_hint;    // 0 default, 1 st.global.cs
};

}  // namespace bri17b200

struct bri17_plan {
  int dim = 0;
  int shape[3] = {1, 1, 1};
  double L[3] = {1, 1, 1};
  double mu = 0, nu = 0;
  double scaling = 0;  // mu / (1 - 2 nu), bri17.hpp:266
  int device = 0;
  int sm_count = 0;
  bri17b200::AxisTables tab[3];
  int apply_variant = -1;  // -1: default
  int solve_variant = 0;   // per-mode solves in 3-D: 1 = one mode per thread, <= 85 registers, 3 CTAs / SM (measured slower); 0 = default
  int mapping = 0;         // 0 auto, 1 always row tiles, 2 always flat tiles
  int64_t host_chunk_rows = 0;
  int host_streams = 3;
  int host_zero_copy = 0;  // 1: kernels access pinned host memory directly (no staging)
  // staging for the host-buffer path (lazily allocated, owned by the plan)
  struct HostStage {
    cudaStream_t stream = nullptr;
    void *in = nullptr, *out = nullptr;
  };
  std::vector<HostStage> stages;
  int64_t stage_bytes = 0;
  std::mutex host_mutex;   // the staging set is shared: host-buffer calls on one plan are serialised
  // diagnostics only (bri17_plan_get_info); atomics so that concurrent launches on one plan do not race
  std::atomic<int64_t> last_grid{0}, last_block{0}, last_smem{0}, launches{0};
  std::atomic<int> last_flat{0};
};

namespace bri17b200 {

int make_block(const bri17_plan *p, const int *k_begin, const int *local_shape, Block *b);
int num_variants();
const Variant &variant(int i);
int default_variant(const bri17_plan *p);

int launch_apply(bri17_plan *p, const Block &b, const void *u, void *f, int64_t u_stride,
                 int64_t f_stride, double out_scale, cudaStream_t stream);
int launch_index_map(bri17_plan *p, const Block &b, int32_t *k_out, cudaStream_t stream);
int launch_stiffness_field(bri17_plan *p, const Block &b, void *K, cudaStream_t stream);
int launch_strain_field(bri17_plan *p, const Block &b, void *B, cudaStream_t stream);
int launch_strain_apply(bri17_plan *p, const Block &b, const void *u, void *eps,
                        int64_t u_stride, int64_t e_stride, double out_scale,
                        cudaStream_t stream);
int launch_apply_dot(bri17_plan *p, const Block &b, const void *u, void *f, int64_t u_stride,
                     int64_t f_stride, double out_scale, int herm_n, double *dot_out, double *scratch,
                     int scratch_count, cudaStream_t stream);
// mode: 0 u^ = K^-1 f^; 1 u^ = K^-1 (tau^ . conj B^); 2 eta^ = sym(B^ (x) u^) of mode 1;
//       3 f^ = tau^ . conj B^ (no solve: the right-hand side of mode 1, bri17.hpp:340)
int launch_modal_solve(bri17_plan *p, const Block &b, int mode, const void *in, void *out,
                       int64_t in_cs, int64_t in_ms, int64_t out_cs, int64_t out_ms, cudaStream_t stream);
int walk_tiles_host(const Block &b, int tile_modes, int max_ctas, int cta, int64_t *out, int cap, int *grid_out);
int apply_host(bri17_plan *p, const Block &b, const void *u_host, void *f_host,
               int64_t comp_stride, double out_scale);
void free_host_stages(bri17_plan *p);

}  // namespace bri17b200
