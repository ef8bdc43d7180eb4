// modal_kernels.cu -- hand-written sm_100a kernels of the modal operator path.
//
// K1/K2  modal_stiffness_apply<DIM>   f^[k] = K^[k] u^[k] for every frequency of a block
//                                     (replaces the loop nest of tests/test_bri17.cpp:58-92;
//                                      K^ per bri17.hpp:247-292, regenerated per mode in registers)
// K3     strain_displacement_*        B^[k] field and eps^ = sym(B^ (x) u^) (bri17.hpp:212-236,
//                                      tests/test_bri17.cpp:194-235)
// K6     freq_index_map               the multi-index every thread derives (parity check)
//        modal_stiffness_field        K^[k] written out mode by mode (diagnostic)
//
// Design:
//  * HBM-bound streaming: 96 B/mode (3-D) and 64 B/mode (2-D) of algorithmic
//    traffic, ~45 FP64 instructions per mode.  No tensor cores, no GEMM shape.
//  * Persistent CTAs (grid = SMs x resident CTAs) walk (row, chunk) tiles with
//    a division-free cursor; a row is one line of the fastest axis, so every
//    128-bit access of a warp is one contiguous 512-byte run per component.
//  * A CTA keeps the same chunk of the fastest axis from one row to the next,
//    so the fastest-axis phi/chi/psi of its columns stay in REGISTERS; the
//    tables of the slower axes are staged once per CTA in shared memory and
//    read as broadcasts, six doubles per row.
//  * All loads of a tile are issued before the first use (VEC x DIM
//    independent 16-byte loads per thread in flight).
//  * Arithmetic uses __dmul_rn/__dadd_rn in the reference's written order, so
//    no multiply-add is contracted and the result is bit-identical to the
//    reference compiled without FMA (the oracle's parity build).
#include <algorithm>
#include <cstdio>

#include "internal.h"
#include "mode_math.h"

namespace bri17b200 {

// ---------------------------------------------------------------------------
// memory access helpers
// ---------------------------------------------------------------------------
template <int HINT>
__device__ __forceinline__ double2 ld_c128(const double2 *p) {
  if constexpr (HINT == 1) {
    return __ldcs(p);  // ld.global.cs: streaming, evict first
  } else if constexpr (HINT == 2) {
    return __ldg(p);  // ld.global.nc
  } else if constexpr (HINT == 3) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];"
                 : "=d"(v.x), "=d"(v.y)
                 : "l"(p));
    return v;
  } else if constexpr (HINT == 4) {
    double2 v;
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];"
                 : "=d"(v.x), "=d"(v.y)
                 : "l"(p));
    return v;
  } else {
    return *p;
  }
}

template <int HINT>
__device__ __forceinline__ void st_c128(double2 *p, double2 v) {
  if constexpr (HINT == 1) {
    __stcs(p, v);  // st.global.cs
  } else if constexpr (HINT == 2) {
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y)
                 : "memory");
  } else {
    *p = v;
  }
}

// ---------------------------------------------------------------------------
// tile cursor: which (row, chunk) a persistent CTA works on, and the
// frequency indices that go with it.  Shared by every kernel in this file, so
// the index map kernel (K6) tests the mapping the apply kernels use.
// ---------------------------------------------------------------------------
struct TileCursor {
  long long tile;
  long long row;  // row-major index over all axes but the fastest
  int chunk;      // which TILE-wide piece of the row
  int a, b;       // row = a * n_mid + b

  __host__ __device__ __forceinline__ void init(const TileGeom &g, unsigned block) {
    tile = block;
    row = tile / g.cpr;
    chunk = int(tile - row * g.cpr);
    a = int(row / g.n_mid);
    b = int(row - (long long)a * g.n_mid);
  }
  __host__ __device__ __forceinline__ bool valid(const TileGeom &g) const { return tile < g.n_tiles; }
  __host__ __device__ __forceinline__ void next(const TileGeom &g, unsigned grid) {
    tile += grid;
    row += g.d_row;
    chunk += g.d_chunk;
    a += g.d_a;
    b += g.d_b;
    if (chunk >= g.cpr) { chunk -= g.cpr; row++; b++; }
    if (b >= g.n_mid) { b -= g.n_mid; a++; }
  }
};

__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }

// Shared-memory staging of the slow-axis tables (local range only).
// Layout: [phi|chi|psi] of the outer axis (n_outer each), then of the mid axis.
struct SlowTables {
  const double *outer;  // phi at [0,n), chi at [n,2n), psi at [2n,3n); n = stride_o
  const double *mid;
  int stride_o, stride_m;
  int off_o, off_m;  // index of local 0 in the arrays
};

template <int DIM>
__device__ __forceinline__ SlowTables stage_slow_tables(const ApplyParams &p, double *smem,
                                                        int n_outer_local) {
  SlowTables t;
  if (p.stage_outer) {
    const int no = n_outer_local, nm = DIM == 3 ? p.g.n_mid : 0;
    for (int i = threadIdx.x; i < 3 * no; i += blockDim.x) {
      const int w = i / no, k = i - w * no;
      smem[i] = p.tab_outer[size_t(w) * p.N_outer + p.g.kb_outer + k];
    }
    if constexpr (DIM == 3) {
      for (int i = threadIdx.x; i < 3 * nm; i += blockDim.x) {
        const int w = i / nm, k = i - w * nm;
        smem[3 * no + i] = p.tab_mid[size_t(w) * p.N_mid + p.g.kb_mid + k];
      }
    }
    __syncthreads();
    t.outer = smem; t.stride_o = no; t.off_o = 0;
    t.mid = smem + 3 * no; t.stride_m = nm; t.off_m = 0;
  } else {  // tables too large for the smem budget: read-only L1/L2 path
    t.outer = p.tab_outer; t.stride_o = p.N_outer; t.off_o = p.g.kb_outer;
    t.mid = p.tab_mid; t.stride_m = p.N_mid; t.off_m = p.g.kb_mid;
  }
  return t;
}

// ---------------------------------------------------------------------------
// K1 / K2: modal stiffness apply
// ---------------------------------------------------------------------------
template <int DIM, int THREADS, int VEC, int MINB, int LDH, int STH>
__global__ void __launch_bounds__(THREADS, MINB) modal_stiffness_apply_kernel(const ApplyParams p) {
  extern __shared__ double smem[];
  constexpr int TILE = THREADS * VEC;
  const TileGeom &g = p.g;
  const SlowTables st = stage_slow_tables<DIM>(p, smem, g.n_outer);

  const double mu = p.mu, scaling = p.scaling;
  const bool scale_out = p.out_scale != 1.0;

  TileCursor cur;
  cur.init(g, blockIdx.x);
  int cached_chunk = -1;
  int col[VEC];
  bool ok[VEC];
  double phiI[VEC], chiI[VEC], psiI[VEC];  // fastest-axis table entries of my columns

  for (; cur.valid(g); cur.next(g, gridDim.x)) {
    if (cur.chunk != cached_chunk) {  // CTA-uniform; taken once when gridDim.x % cpr == 0
      cached_chunk = cur.chunk;
#pragma unroll
      for (int j = 0; j < VEC; j++) {
        col[j] = cur.chunk * TILE + j * THREADS + threadIdx.x;
        ok[j] = col[j] < g.n_inner;
        const int k = g.kb_inner + (ok[j] ? col[j] : 0);
        phiI[j] = __ldg(p.tab_inner + size_t(TAB_PHI) * p.N_inner + k);
        chiI[j] = __ldg(p.tab_inner + size_t(TAB_CHI) * p.N_inner + k);
        psiI[j] = __ldg(p.tab_inner + size_t(TAB_PSI) * p.N_inner + k);
      }
    }

    // ---- issue every load of the tile first ----
    const long long base = cur.row * g.n_inner;
    double2 u[VEC][DIM];
#pragma unroll
    for (int j = 0; j < VEC; j++)
#pragma unroll
      for (int c = 0; c < DIM; c++)
        if (ok[j]) u[j][c] = ld_c128<LDH>(p.u + c * p.u_stride + base + col[j]);

    // ---- row invariants (broadcast reads) ----
    const int ia = st.off_o + cur.a;
    const double p0 = st.outer[ia], c0 = st.outer[st.stride_o + ia], s0 = st.outer[2 * st.stride_o + ia];

    if constexpr (DIM == 3) {
      const int ib = st.off_m + cur.b;
      const double p1 = st.mid[ib], c1 = st.mid[st.stride_m + ib], s1 = st.mid[2 * st.stride_m + ib];
      // bri17.hpp:276-288, factors that do not depend on the fastest axis
      const double h00r = mul(p0, c1);            // phi0*chi1
      const double h11r = mul(c0, p1);            // chi0*phi1
      const double h22r = mul(c0, c1);            // chi0*chi1
      const double k01r = mul(mul(scaling, s0), s1);  // scaling*psi0*psi1
      const double k02r = mul(mul(scaling, s0), c1);  // scaling*psi0*chi1
      const double k12r = mul(mul(scaling, c0), s1);  // scaling*chi0*psi1
#pragma unroll
      for (int j = 0; j < VEC; j++) {
        if (!ok[j]) continue;
        const double H00 = mul(h00r, chiI[j]);                 // :276
        const double H11 = mul(h11r, chiI[j]);                 // :277
        const double H22 = mul(h22r, phiI[j]);                 // :278
        const double Kd = mul(mu, add(add(H00, H11), H22));    // :279
        const double K00 = add(mul(scaling, H00), Kd);         // :280
        const double K11 = add(mul(scaling, H11), Kd);         // :284
        const double K22 = add(mul(scaling, H22), Kd);         // :288
        const double K01 = mul(k01r, chiI[j]);                 // :281
        const double K02 = mul(k02r, psiI[j]);                 // :282
        const double K12 = mul(k12r, psiI[j]);                 // :285
        const double2 u0 = u[j][0], u1 = u[j][1], u2 = u[j][2];
        // tests/test_bri17.cpp:84, Im(K)=0, accumulated left to right
        double2 f0, f1, f2;
        f0.x = add(add(mul(K00, u0.x), mul(K01, u1.x)), mul(K02, u2.x));
        f0.y = add(add(mul(K00, u0.y), mul(K01, u1.y)), mul(K02, u2.y));
        f1.x = add(add(mul(K01, u0.x), mul(K11, u1.x)), mul(K12, u2.x));
        f1.y = add(add(mul(K01, u0.y), mul(K11, u1.y)), mul(K12, u2.y));
        f2.x = add(add(mul(K02, u0.x), mul(K12, u1.x)), mul(K22, u2.x));
        f2.y = add(add(mul(K02, u0.y), mul(K12, u1.y)), mul(K22, u2.y));
        if (scale_out) {
          f0.x = mul(f0.x, p.out_scale); f0.y = mul(f0.y, p.out_scale);
          f1.x = mul(f1.x, p.out_scale); f1.y = mul(f1.y, p.out_scale);
          f2.x = mul(f2.x, p.out_scale); f2.y = mul(f2.y, p.out_scale);
        }
        double2 *o = p.f + base + col[j];
        st_c128<STH>(o, f0);
        st_c128<STH>(o + p.f_stride, f1);
        st_c128<STH>(o + 2 * p.f_stride, f2);
      }
    } else {
      const double k01r = mul(scaling, s0);  // scaling*psi0
#pragma unroll
      for (int j = 0; j < VEC; j++) {
        if (!ok[j]) continue;
        const double H00 = mul(p0, chiI[j]);               // bri17.hpp:268
        const double H11 = mul(c0, phiI[j]);               // :269
        const double Kd = mul(mu, add(H00, H11));          // :270
        const double K00 = add(mul(scaling, H00), Kd);     // :271
        const double K01 = mul(k01r, psiI[j]);             // :272
        const double K11 = add(mul(scaling, H11), Kd);     // :274
        const double2 u0 = u[j][0], u1 = u[j][1];
        double2 f0, f1;  // tests/test_bri17.cpp:68
        f0.x = add(mul(K00, u0.x), mul(K01, u1.x));
        f0.y = add(mul(K00, u0.y), mul(K01, u1.y));
        f1.x = add(mul(K01, u0.x), mul(K11, u1.x));
        f1.y = add(mul(K01, u0.y), mul(K11, u1.y));
        if (scale_out) {
          f0.x = mul(f0.x, p.out_scale); f0.y = mul(f0.y, p.out_scale);
          f1.x = mul(f1.x, p.out_scale); f1.y = mul(f1.y, p.out_scale);
        }
        double2 *o = p.f + base + col[j];
        st_c128<STH>(o, f0);
        st_c128<STH>(o + p.f_stride, f1);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// K1f / K2f: "flat" modal stiffness apply for rows that do not fill whole tiles
// (fastest extent such as 513 = N/2+1 of a half spectrum, 300, 5, ...).  A tile
// is THREADS*VEC consecutive elements of the row-major block, rows concatenated,
// so no lane is idle except in the very last tile; each element derives its
// (a, b, col) by two integer divisions and reads its nine factors from the
// shared-memory tables of all three axes.  Same arithmetic, same results.
// ---------------------------------------------------------------------------
struct FlatIndex {
  long long i;
  int a, b, col;
};

// Weight of frequency k of the fastest axis in a sum over the FULL spectrum when only the half
// spectrum k <= N/2 of a real field is stored: k and N-k are a conjugate pair except k = 0 and
// k = N/2 (even N).  herm_n = 0: the axis is complete, every mode counts once.
__device__ __forceinline__ double pair_weight(int herm_n, int k) {
  return (herm_n > 0 && k != 0 && 2 * k != herm_n) ? 2. : 1.;
}

// Sum over the CTA, valid in thread 0 (fixed order: deterministic).
template <int THREADS>
__device__ __forceinline__ double cta_sum(double v) {
  __shared__ double sh[THREADS / 32];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.;
  if (threadIdx.x < THREADS / 32) t = sh[threadIdx.x];
  if (threadIdx.x < 32)
    for (int o = THREADS / 64; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;
}
__device__ __forceinline__ FlatIndex flat_index(const TileGeom &g, long long i) {
  FlatIndex f;
  f.i = i;
  if (g.n_tiles <= 0x7fffffffLL / 1 && g.n_rows * g.n_inner <= 0x7fffffffLL) {  // 32-bit fast path
    const unsigned iu = unsigned(i), row = iu / unsigned(g.n_inner);
    f.col = int(iu - row * unsigned(g.n_inner));
    f.a = int(row / unsigned(g.n_mid));
    f.b = int(row - unsigned(f.a) * unsigned(g.n_mid));
  } else {
    const long long row = i / g.n_inner;
    f.col = int(i - row * g.n_inner);
    f.a = int(row / g.n_mid);
    f.b = int(row - (long long)f.a * g.n_mid);
  }
  return f;
}

// DOT: also accumulates sum_k w_k Re(u^_k^H f^_k) (f^ including out_scale) into one partial sum per
// CTA -- by Parseval this is |N| <u, F> of the real-space fields, which lets CG take <p, A p> from
// the operator application itself instead of a separate pass over both vectors.
template <int DIM, int THREADS, int VEC, int MINB, bool DOT>
__global__ void __launch_bounds__(THREADS, MINB) modal_stiffness_apply_flat_kernel(const ApplyParams p) {
  extern __shared__ double smem[];
  double dot_acc = 0.;
  constexpr int TILE = THREADS * VEC;
  const TileGeom &g = p.g;
  // tables of the local ranges: outer | mid | inner, each [phi|chi|psi]
  const double *tO, *tM, *tI;
  int sO, sM, sI, oO, oM, oI;
  if (p.stage_outer) {
    const int no = g.n_outer, nm = DIM == 3 ? g.n_mid : 0, ni = g.n_inner;
    for (int x = threadIdx.x; x < 3 * no; x += THREADS) {
      const int w = x / no, kk = x - w * no;
      smem[x] = p.tab_outer[size_t(w) * p.N_outer + g.kb_outer + kk];
    }
    if constexpr (DIM == 3)
      for (int x = threadIdx.x; x < 3 * nm; x += THREADS) {
        const int w = x / nm, kk = x - w * nm;
        smem[3 * no + x] = p.tab_mid[size_t(w) * p.N_mid + g.kb_mid + kk];
      }
    for (int x = threadIdx.x; x < 3 * ni; x += THREADS) {
      const int w = x / ni, kk = x - w * ni;
      smem[3 * (no + nm) + x] = p.tab_inner[size_t(w) * p.N_inner + g.kb_inner + kk];
    }
    __syncthreads();
    tO = smem; sO = no; oO = 0;
    tM = smem + 3 * no; sM = nm; oM = 0;
    tI = smem + 3 * (no + nm); sI = ni; oI = 0;
  } else {
    tO = p.tab_outer; sO = p.N_outer; oO = g.kb_outer;
    tM = p.tab_mid; sM = p.N_mid; oM = g.kb_mid;
    tI = p.tab_inner; sI = p.N_inner; oI = g.kb_inner;
  }
  const double mu = p.mu, scaling = p.scaling;
  const bool scale_out = p.out_scale != 1.0;
  const long long total = g.n_rows * g.n_inner;

  for (long long tile = blockIdx.x; tile * TILE < total; tile += gridDim.x) {
    FlatIndex f[VEC];
    bool ok[VEC];
    double2 u[VEC][DIM];
#pragma unroll
    for (int j = 0; j < VEC; j++) {
      const long long i = tile * TILE + j * THREADS + threadIdx.x;
      ok[j] = i < total;
      f[j] = flat_index(g, ok[j] ? i : 0);
#pragma unroll
      for (int c = 0; c < DIM; c++)
        if (ok[j]) u[j][c] = __ldcs(p.u + c * p.u_stride + i);
    }
#pragma unroll
    for (int j = 0; j < VEC; j++) {
      if (!ok[j]) continue;
      const int ia = oO + f[j].a, ii = oI + f[j].col;
      const double p0 = tO[ia], c0 = tO[sO + ia], s0 = tO[2 * sO + ia];
      const double pI = tI[ii], cI = tI[sI + ii], sIv = tI[2 * sI + ii];
      double2 *o = p.f + f[j].i;
      if constexpr (DIM == 3) {
        const int ib = oM + f[j].b;
        const double p1 = tM[ib], c1 = tM[sM + ib], s1 = tM[2 * sM + ib];
        const double H00 = mul(mul(p0, c1), cI);                 // bri17.hpp:276
        const double H11 = mul(mul(c0, p1), cI);                 // :277
        const double H22 = mul(mul(c0, c1), pI);                 // :278
        const double Kd = mul(mu, add(add(H00, H11), H22));      // :279
        const double K00 = add(mul(scaling, H00), Kd);           // :280
        const double K11 = add(mul(scaling, H11), Kd);           // :284
        const double K22 = add(mul(scaling, H22), Kd);           // :288
        const double K01 = mul(mul(mul(scaling, s0), s1), cI);   // :281
        const double K02 = mul(mul(mul(scaling, s0), c1), sIv);  // :282
        const double K12 = mul(mul(mul(scaling, c0), s1), sIv);  // :285
        const double2 u0 = u[j][0], u1 = u[j][1], u2 = u[j][2];
        double2 f0, f1, f2;
        f0.x = add(add(mul(K00, u0.x), mul(K01, u1.x)), mul(K02, u2.x));
        f0.y = add(add(mul(K00, u0.y), mul(K01, u1.y)), mul(K02, u2.y));
        f1.x = add(add(mul(K01, u0.x), mul(K11, u1.x)), mul(K12, u2.x));
        f1.y = add(add(mul(K01, u0.y), mul(K11, u1.y)), mul(K12, u2.y));
        f2.x = add(add(mul(K02, u0.x), mul(K12, u1.x)), mul(K22, u2.x));
        f2.y = add(add(mul(K02, u0.y), mul(K12, u1.y)), mul(K22, u2.y));
        if (scale_out) {
          f0.x = mul(f0.x, p.out_scale); f0.y = mul(f0.y, p.out_scale);
          f1.x = mul(f1.x, p.out_scale); f1.y = mul(f1.y, p.out_scale);
          f2.x = mul(f2.x, p.out_scale); f2.y = mul(f2.y, p.out_scale);
        }
        __stcs(o, f0);
        __stcs(o + p.f_stride, f1);
        __stcs(o + 2 * p.f_stride, f2);
        if constexpr (DOT) {
          const double w = pair_weight(p.herm_n, g.kb_inner + f[j].col);
          dot_acc += w * (u0.x * f0.x + u0.y * f0.y + u1.x * f1.x + u1.y * f1.y + u2.x * f2.x + u2.y * f2.y);
        }
      } else {
        const double H00 = mul(p0, cI);                          // :268
        const double H11 = mul(c0, pI);                          // :269
        const double Kd = mul(mu, add(H00, H11));                // :270
        const double K00 = add(mul(scaling, H00), Kd);           // :271
        const double K01 = mul(mul(scaling, s0), sIv);           // :272
        const double K11 = add(mul(scaling, H11), Kd);           // :274
        const double2 u0 = u[j][0], u1 = u[j][1];
        double2 f0, f1;
        f0.x = add(mul(K00, u0.x), mul(K01, u1.x));
        f0.y = add(mul(K00, u0.y), mul(K01, u1.y));
        f1.x = add(mul(K01, u0.x), mul(K11, u1.x));
        f1.y = add(mul(K01, u0.y), mul(K11, u1.y));
        if (scale_out) {
          f0.x = mul(f0.x, p.out_scale); f0.y = mul(f0.y, p.out_scale);
          f1.x = mul(f1.x, p.out_scale); f1.y = mul(f1.y, p.out_scale);
        }
        __stcs(o, f0);
        __stcs(o + p.f_stride, f1);
        if constexpr (DOT) {
          const double w = pair_weight(p.herm_n, g.kb_inner + f[j].col);
          dot_acc += w * (u0.x * f0.x + u0.y * f0.y + u1.x * f1.x + u1.y * f1.y);
        }
      }
    }
  }
  if constexpr (DOT) {
    dot_acc = cta_sum<THREADS>(dot_acc);
    if (threadIdx.x == 0) p.dot_partial[blockIdx.x] = dot_acc;
  }
}

// partial[0..n) -> *out, one CTA, fixed summation order (deterministic)
__global__ void __launch_bounds__(256) dot_finish_kernel(const double *partial, int n, double *out) {
  double acc = 0.;
  for (int i = threadIdx.x; i < n; i += 256) acc += partial[i];
  acc = cta_sum<256>(acc);
  if (threadIdx.x == 0) *out = acc;
}

// K6, flat mapping
template <int DIM>
__global__ void __launch_bounds__(256) freq_index_map_flat_kernel(const TileGeom g, int32_t *k_out) {
  constexpr int THREADS = 256, VEC = 2, TILE = THREADS * VEC;
  const long long total = g.n_rows * g.n_inner;
  for (long long tile = blockIdx.x; tile * TILE < total; tile += gridDim.x)
#pragma unroll
    for (int j = 0; j < VEC; j++) {
      const long long i = tile * TILE + j * THREADS + threadIdx.x;
      if (i >= total) continue;
      const FlatIndex f = flat_index(g, i);
      int32_t *o = k_out + i * DIM;
      o[0] = g.kb_outer + f.a;
      if constexpr (DIM == 3) { o[1] = g.kb_mid + f.b; o[2] = g.kb_inner + f.col; }
      else o[1] = g.kb_inner + f.col;
    }
}

// ---------------------------------------------------------------------------
// K6: frequency index map (same cursor, THREADS=256, VEC=2 tiles)
// ---------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(256) freq_index_map_kernel(const TileGeom g, int32_t *k_out) {
  constexpr int THREADS = 256, VEC = 2, TILE = THREADS * VEC;
  TileCursor cur;
  for (cur.init(g, blockIdx.x); cur.valid(g); cur.next(g, gridDim.x)) {
    const long long base = cur.row * g.n_inner;
#pragma unroll
    for (int j = 0; j < VEC; j++) {
      const int col = cur.chunk * TILE + j * THREADS + threadIdx.x;
      if (col >= g.n_inner) continue;
      int32_t *o = k_out + (base + col) * DIM;
      o[0] = g.kb_outer + cur.a;
      if constexpr (DIM == 3) {
        o[1] = g.kb_mid + cur.b;
        o[2] = g.kb_inner + col;
      } else {
        o[1] = g.kb_inner + col;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// K3: strain-displacement.  B^_i = prefactor * s_i * prod_{j != i} c_j with
// prefactor = (-2 sin S, 2 cos S), S = sum of the half angles (bri17.hpp:224).
// c/s come from the host tables.  sin S / cos S are NOT evaluated with a device
// sincos (a ~100-instruction fp64 software sequence that made these kernels
// FP64-bound in round 1): e^{iS} = prod_d (cos a_d + i sin a_d) with both factors
// of every axis taken from the host tables, two complex products per mode.  Each
// table entry is correctly rounded libm output, so cos S / sin S carry an absolute
// error of a few ulp of 1 -- the same as a device sincos -- and B^ / eps^ agree with
// the reference to <= 1e-12 per mode (max-norm), not bitwise.
// ---------------------------------------------------------------------------
// Half-angle factors of one grid line: c = cos(alpha), s = sin(alpha)*N/L, sa = sin(alpha),
// alpha = (pi*k)/N, all three from the host tables (bri17.hpp:218-221).
struct HalfAngle {
  double c, s, sa;
};
__device__ __forceinline__ HalfAngle half_angle(const double *tab, int N, int k) {
  HalfAngle h;
  h.c = __ldg(tab + size_t(TAB_C) * N + k);
  h.s = __ldg(tab + size_t(TAB_S) * N + k);
  h.sa = __ldg(tab + size_t(TAB_SINA) * N + k);
  return h;
}

// (cos, sin) of the sum of two angles from their (cos, sin) pairs.
__device__ __forceinline__ double2 rot_mul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// B^ of one mode from the per-axis factors (bri17.hpp:224-232); e_row = (cos, sin) of the
// sum of the half angles of the slower axes (1, 0 when there is none).
template <int DIM>
__device__ __forceinline__ void modal_B(const HalfAngle &h0, const HalfAngle &h1, const HalfAngle &hI,
                                        double2 e_row, double2 *B) {
  const double2 e = rot_mul(e_row, make_double2(hI.c, hI.sa));
  const double pre_re = mul(-2., e.y), pre_im = mul(2., e.x);  // :224
  if constexpr (DIM == 3) {
    B[0] = make_double2(mul(mul(mul(pre_re, h0.s), h1.c), hI.c), mul(mul(mul(pre_im, h0.s), h1.c), hI.c));
    B[1] = make_double2(mul(mul(mul(pre_re, h0.c), h1.s), hI.c), mul(mul(mul(pre_im, h0.c), h1.s), hI.c));
    B[2] = make_double2(mul(mul(mul(pre_re, h0.c), h1.c), hI.s), mul(mul(mul(pre_im, h0.c), h1.c), hI.s));
  } else {
    B[0] = make_double2(mul(mul(pre_re, h0.s), hI.c), mul(mul(pre_im, h0.s), hI.c));
    B[1] = make_double2(mul(mul(pre_re, h0.c), hI.s), mul(mul(pre_im, h0.c), hI.s));
  }
}

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(add(mul(a.x, b.x), -mul(a.y, b.y)), add(mul(a.x, b.y), mul(a.y, b.x)));
}

// ---------------------------------------------------------------------------
// Mode-major field writers: what a loop over Hooke::modal_strain_displacement (WHAT 0,
// DIM complex per mode) or Hooke::modal_stiffness (WHAT 1, DIM*DIM complex per mode,
// diagnostic) would write.  A tile is 256 consecutive modes of a row, one per thread; its
// output is ONE contiguous run of 256*NC complex numbers, so the values are staged in shared
// memory and written out with fully coalesced 16-byte stores (round 1 stored 48-144 B per
// lane with a stride: 0.19-0.48 of the HBM roofline).
// ---------------------------------------------------------------------------
template <int DIM, int WHAT>
__global__ void __launch_bounds__(256) modal_field_kernel(const ApplyParams p) {
  constexpr int THREADS = 256, TILE = THREADS;
  constexpr int NC = WHAT == 0 ? DIM : DIM * DIM;
  __shared__ double2 stage[TILE * NC];  // 3-D K^: 36 KiB
  const TileGeom &g = p.g;
  const double mu = p.mu, scaling = p.scaling;
  TileCursor cur;
  for (cur.init(g, blockIdx.x); cur.valid(g); cur.next(g, gridDim.x)) {
    const int col0 = cur.chunk * TILE;
    const int valid = min(TILE, g.n_inner - col0);
    if (int(threadIdx.x) < valid) {
      const int ka = g.kb_outer + cur.a, kb = g.kb_mid + cur.b, ki = g.kb_inner + col0 + threadIdx.x;
      double2 *o = stage + threadIdx.x * NC;
      if constexpr (WHAT == 0) {
        const HalfAngle h0 = half_angle(p.tab_outer, p.N_outer, ka);
        HalfAngle h1 = h0;
        double2 e_row = make_double2(h0.c, h0.sa);
        if constexpr (DIM == 3) {
          h1 = half_angle(p.tab_mid, p.N_mid, kb);
          e_row = rot_mul(e_row, make_double2(h1.c, h1.sa));
        }
        double2 B[DIM];
        modal_B<DIM>(h0, h1, half_angle(p.tab_inner, p.N_inner, ki), e_row, B);
#pragma unroll
        for (int c = 0; c < DIM; c++) o[c] = B[c];
      } else {
        const double p0 = __ldg(p.tab_outer + ka), c0 = __ldg(p.tab_outer + p.N_outer + ka),
                     s0 = __ldg(p.tab_outer + 2 * size_t(p.N_outer) + ka);
        const double pI = __ldg(p.tab_inner + ki), cI = __ldg(p.tab_inner + p.N_inner + ki),
                     sI = __ldg(p.tab_inner + 2 * size_t(p.N_inner) + ki);
        if constexpr (DIM == 3) {
          const double p1 = __ldg(p.tab_mid + kb), c1 = __ldg(p.tab_mid + p.N_mid + kb),
                       s1 = __ldg(p.tab_mid + 2 * size_t(p.N_mid) + kb);
          const double H00 = mul(mul(p0, c1), cI);
          const double H11 = mul(mul(c0, p1), cI);
          const double H22 = mul(mul(c0, c1), pI);
          const double Kd = mul(mu, add(add(H00, H11), H22));
          const double K00 = add(mul(scaling, H00), Kd);
          const double K01 = mul(mul(mul(scaling, s0), s1), cI);
          const double K02 = mul(mul(mul(scaling, s0), c1), sI);
          const double K11 = add(mul(scaling, H11), Kd);
          const double K12 = mul(mul(mul(scaling, c0), s1), sI);
          const double K22 = add(mul(scaling, H22), Kd);
          o[0] = make_double2(K00, 0.); o[1] = make_double2(K01, 0.); o[2] = make_double2(K02, 0.);
          o[3] = make_double2(K01, 0.); o[4] = make_double2(K11, 0.); o[5] = make_double2(K12, 0.);
          o[6] = make_double2(K02, 0.); o[7] = make_double2(K12, 0.); o[8] = make_double2(K22, 0.);
        } else {
          const double H00 = mul(p0, cI);
          const double H11 = mul(c0, pI);
          const double Kd = mul(mu, add(H00, H11));
          const double K00 = add(mul(scaling, H00), Kd);
          const double K01 = mul(mul(scaling, s0), sI);
          const double K11 = add(mul(scaling, H11), Kd);
          o[0] = make_double2(K00, 0.); o[1] = make_double2(K01, 0.);
          o[2] = make_double2(K01, 0.); o[3] = make_double2(K11, 0.);
        }
      }
    }
    __syncthreads();
    double2 *out = p.f + (cur.row * g.n_inner + col0) * NC;
    for (int i = threadIdx.x; i < valid * NC; i += THREADS) __stcs(out + i, stage[i]);
    __syncthreads();
  }
}

// eps^ = sym(B^ (x) u^), planar Mandel (tests/test_bri17.cpp:194-235).
// Same persistent-CTA structure as the stiffness apply: the fastest-axis
// factors of a thread's columns stay in registers from row to row.
template <int DIM>
__global__ void __launch_bounds__(256, 2) strain_displacement_kernel(const ApplyParams p) {
  constexpr int THREADS = 256, VEC = 2, TILE = THREADS * VEC;
  constexpr int NSYM = DIM * (DIM + 1) / 2;
  const TileGeom &g = p.g;
  const bool scale_out = p.out_scale != 1.0;
  TileCursor cur;
  int cached_chunk = -1;
  int col[VEC];
  bool ok[VEC];
  HalfAngle hI[VEC];
  for (cur.init(g, blockIdx.x); cur.valid(g); cur.next(g, gridDim.x)) {
    if (cur.chunk != cached_chunk) {
      cached_chunk = cur.chunk;
#pragma unroll
      for (int j = 0; j < VEC; j++) {
        col[j] = cur.chunk * TILE + j * THREADS + threadIdx.x;
        ok[j] = col[j] < g.n_inner;
        hI[j] = half_angle(p.tab_inner, p.N_inner, g.kb_inner + (ok[j] ? col[j] : 0));
      }
    }
    const long long base = cur.row * g.n_inner;
    double2 u[VEC][DIM];
#pragma unroll
    for (int j = 0; j < VEC; j++)
#pragma unroll
      for (int c = 0; c < DIM; c++)
        if (ok[j]) u[j][c] = __ldcs(p.u + c * p.u_stride + base + col[j]);
    const HalfAngle h0 = half_angle(p.tab_outer, p.N_outer, g.kb_outer + cur.a);
    HalfAngle h1 = h0;
    double2 e_row = make_double2(h0.c, h0.sa);
    if constexpr (DIM == 3) {
      h1 = half_angle(p.tab_mid, p.N_mid, g.kb_mid + cur.b);
      e_row = rot_mul(e_row, make_double2(h1.c, h1.sa));
    }
#pragma unroll
    for (int j = 0; j < VEC; j++) {
      if (!ok[j]) continue;
      double2 B[DIM];
      modal_B<DIM>(h0, h1, hI[j], e_row, B);
      // Mandel order: tests/test_bri17.cpp:207-209 (2-D), :224-230 (3-D)
      constexpr int P3[6] = {0, 1, 2, 1, 2, 0}, Q3[6] = {0, 1, 2, 2, 0, 1};
      constexpr int P2[3] = {0, 1, 0}, Q2[3] = {0, 1, 1};
      // sqrt(2)*(0.5*x) == (sqrt(2)/2)*x bit for bit: halving is exact
      const double half_sqrt2 = 0.5 * 1.4142135623730951;
      double2 *o = p.f + base + col[j];
#pragma unroll
      for (int s = 0; s < NSYM; s++) {
        const int pp = DIM == 3 ? P3[s] : P2[s], qq = DIM == 3 ? Q3[s] : Q2[s];
        double2 e;
        if (pp == qq) {
          // 0.5*(B_p u_p + u_p B_p): both products are bit-identical, so the sum is an
          // exact doubling and the half undoes it exactly (:206, :223)
          e = cmul(B[pp], u[j][pp]);
        } else {
          const double2 t1 = cmul(B[pp], u[j][qq]);
          const double2 t2 = cmul(u[j][pp], B[qq]);
          e = make_double2(mul(half_sqrt2, add(t1.x, t2.x)), mul(half_sqrt2, add(t1.y, t2.y)));
        }
        if (scale_out) e = make_double2(mul(e.x, p.out_scale), mul(e.y, p.out_scale));
        __stcs(o + s * p.f_stride, e);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Per-mode direct solves (bri17.hpp:308-355).
//   MODE 0: u^ = K^-1 f^                      (exact inverse of the stiffness apply, u^(0) = 0)
//   MODE 1: u^ = K^-1 (tau^ . conj(B^))        (bri17.hpp:324-341)
//   MODE 2: eta^ = sym(B^ (x) u^), u^ of MODE 1 (bri17.hpp:342-353): -strain induced by an eigenstress
//   MODE 3: f^ = tau^ . conj(B^)              (bri17.hpp:340 without the solve: the modal force that
//                                              is the right-hand side of the inclusion problem)
// K^ and B^ are rebuilt per mode from the tables; the 2x2/3x3 Cholesky runs in registers.
// Input/output element (s, i) at base[i*mstride + s*cstride]: planar (mstride 1) or
// mode-major (cstride 1), the layout of python/demo.py:21,37-38.
// ---------------------------------------------------------------------------
struct AxisFactors {
  double phi, chi, psi, c, s, sa;
};
__device__ __forceinline__ AxisFactors axis_factors(const double *tab, int N, int k) {
  AxisFactors f;
  f.phi = __ldg(tab + size_t(TAB_PHI) * N + k);
  f.chi = __ldg(tab + size_t(TAB_CHI) * N + k);
  f.psi = __ldg(tab + size_t(TAB_PSI) * N + k);
  f.c = __ldg(tab + size_t(TAB_C) * N + k);
  f.s = __ldg(tab + size_t(TAB_S) * N + k);
  f.sa = __ldg(tab + size_t(TAB_SINA) * N + k);
  return f;
}

template <int DIM, int MODE, int VEC, int MINB = 2>
__global__ void __launch_bounds__(256, MINB) modal_solve_kernel(const ApplyParams p) {
  constexpr int THREADS = 256, TILE = THREADS * VEC;
  constexpr int NIN = MODE == 0 ? DIM : DIM * (DIM + 1) / 2;
  constexpr int NOUT = MODE == 2 ? DIM * (DIM + 1) / 2 : DIM;
  const TileGeom &g = p.g;
  TileCursor cur;
  int cached_chunk = -1;
  int col[VEC];
  bool ok[VEC];
  AxisFactors fI[VEC];
  for (cur.init(g, blockIdx.x); cur.valid(g); cur.next(g, gridDim.x)) {
    if (cur.chunk != cached_chunk) {
      cached_chunk = cur.chunk;
#pragma unroll
      for (int j = 0; j < VEC; j++) {
        col[j] = cur.chunk * TILE + j * THREADS + threadIdx.x;
        ok[j] = col[j] < g.n_inner;
        fI[j] = axis_factors(p.tab_inner, p.N_inner, g.kb_inner + (ok[j] ? col[j] : 0));
      }
    }
    // every load of the tile first
    double2 raw[VEC][NIN];
#pragma unroll
    for (int j = 0; j < VEC; j++)
#pragma unroll
      for (int s = 0; s < NIN; s++)
        if (ok[j]) raw[j][s] = __ldcs(p.u + (cur.row * g.n_inner + col[j]) * p.u_mstride + s * p.u_stride);
    const int k0 = g.kb_outer + cur.a, k1 = g.kb_mid + cur.b;
    const AxisFactors f0 = axis_factors(p.tab_outer, p.N_outer, k0);
    AxisFactors f1 = f0;
    if constexpr (DIM == 3) f1 = axis_factors(p.tab_mid, p.N_mid, k1);
#pragma unroll
    for (int j = 0; j < VEC; j++) {
      if (!ok[j]) continue;
      const long long i = cur.row * g.n_inner + col[j];
      const int kI = g.kb_inner + col[j];
      Cplx in[NIN];
#pragma unroll
      for (int s = 0; s < NIN; s++) in[s] = {raw[j][s].x, raw[j][s].y};
      Cplx out[NOUT];
      const bool null_frequency = (k0 == 0) && (DIM == 2 || k1 == 0) && (kI == 0);   // bri17.hpp:327,334
      if (null_frequency) {
#pragma unroll
        for (int s = 0; s < NOUT; s++) out[s] = {0., 0.};                            // :336-339
      } else {
        double phi[DIM], chi[DIM], psi[DIM], c[DIM], sh[DIM];
        phi[0] = f0.phi; chi[0] = f0.chi; psi[0] = f0.psi; c[0] = f0.c; sh[0] = f0.s;
        if constexpr (DIM == 3) { phi[1] = f1.phi; chi[1] = f1.chi; psi[1] = f1.psi; c[1] = f1.c; sh[1] = f1.s; }
        phi[DIM - 1] = fI[j].phi; chi[DIM - 1] = fI[j].chi; psi[DIM - 1] = fI[j].psi;
        c[DIM - 1] = fI[j].c; sh[DIM - 1] = fI[j].s;
        double K[DIM][DIM];
        if constexpr (MODE != 3) stiffness_entries<DIM>(phi, chi, psi, p.mu, p.scaling, K);
        Cplx u[DIM];
        if constexpr (MODE == 0) {
#pragma unroll
          for (int d = 0; d < DIM; d++) u[d] = in[d];
          cholesky_solve<DIM>(K, u);
#pragma unroll
          for (int d = 0; d < DIM; d++) out[d] = u[d];
        } else {
          // e^{i sum(alpha)} from the per-axis (cos, sin) table entries (see modal_B)
          double2 e = make_double2(f0.c, f0.sa);
          if constexpr (DIM == 3) e = rot_mul(e, make_double2(f1.c, f1.sa));
          e = rot_mul(e, make_double2(fI[j].c, fI[j].sa));
          Cplx B[DIM];
          strain_displacement_entries<DIM>(c, sh, Cplx{-2. * e.y, 2. * e.x}, B);
          if constexpr (MODE == 3) {
            eigenstress_to_force<DIM>(in, B, u);
          } else {
            eigenstress_to_displacement<DIM>(in, B, K, u);
          }
          if constexpr (MODE == 1 || MODE == 3) {
#pragma unroll
            for (int d = 0; d < DIM; d++) out[d] = u[d];
          } else {
            displacement_to_strain<DIM>(B, u, out);
          }
        }
      }
#pragma unroll
      for (int s = 0; s < NOUT; s++)
        __stcs(p.f + i * p.f_mstride + s * p.f_stride, make_double2(out[s].re, out[s].im));
    }
  }
}

// ---------------------------------------------------------------------------
// host side: geometry, variants, launchers
// ---------------------------------------------------------------------------
static const Variant kVariants[] = {
    //  name              threads vec minb ld st
    {"t256v2_cs", 256, 2, 2, 1, 1},       // 0
    {"t256v2_plain", 256, 2, 2, 0, 0},    // 1
    {"t256v1_cs", 256, 1, 6, 1, 1},       // 2
    {"t256v4_cs", 256, 4, 1, 1, 1},       // 3
    {"t128v2_cs", 128, 2, 4, 1, 1},       // 4
    {"t128v4_cs", 128, 4, 3, 1, 1},       // 5
    {"t512v1_cs", 512, 1, 3, 1, 1},       // 6
    {"t512v2_cs", 512, 2, 1, 1, 1},       // 7
    {"t256v2_ncna_cs", 256, 2, 2, 3, 1},  // 8
    {"t256v2_ldg_plain", 256, 2, 2, 2, 0},// 9
    {"t256v2_na_na", 256, 2, 2, 4, 2},    // 10
    {"t128v1_cs", 128, 1, 12, 1, 1},      // 11
    {"t256v1_plain", 256, 1, 6, 0, 0},    // 12
    {"t256v1_ncna_cs", 256, 1, 6, 3, 1},  // 13
};

int num_variants() { return int(sizeof(kVariants) / sizeof(kVariants[0])); }
const Variant &variant(int i) { return kVariants[i]; }
// Measured on B200 (profiles/r01_variant_sweep.md): 3-D 512^3 -> t256v2_cs 98.2 % of the measured
// HBM copy peak; 2-D 4096^2 (1 GiB, 0.17 ms) prefers smaller CTAs -> t128v2_cs 98.5 %.
int default_variant(const bri17_plan *p) { return p->dim == 2 ? 4 : 0; }

typedef void (*ApplyKernel)(const ApplyParams);

template <int DIM>
static ApplyKernel apply_kernel_for(int v) {
  switch (v) {
    case 0: return modal_stiffness_apply_kernel<DIM, 256, 2, 2, 1, 1>;
    case 1: return modal_stiffness_apply_kernel<DIM, 256, 2, 2, 0, 0>;
    case 2: return modal_stiffness_apply_kernel<DIM, 256, 1, 6, 1, 1>;
    case 3: return modal_stiffness_apply_kernel<DIM, 256, 4, 1, 1, 1>;
    case 4: return modal_stiffness_apply_kernel<DIM, 128, 2, 4, 1, 1>;
    case 5: return modal_stiffness_apply_kernel<DIM, 128, 4, 3, 1, 1>;
    case 6: return modal_stiffness_apply_kernel<DIM, 512, 1, 3, 1, 1>;
    case 7: return modal_stiffness_apply_kernel<DIM, 512, 2, 1, 1, 1>;
    case 8: return modal_stiffness_apply_kernel<DIM, 256, 2, 2, 3, 1>;
    case 9: return modal_stiffness_apply_kernel<DIM, 256, 2, 2, 2, 0>;
    case 10: return modal_stiffness_apply_kernel<DIM, 256, 2, 2, 4, 2>;
    case 11: return modal_stiffness_apply_kernel<DIM, 128, 1, 12, 1, 1>;
    case 12: return modal_stiffness_apply_kernel<DIM, 256, 1, 6, 0, 0>;
    case 13: return modal_stiffness_apply_kernel<DIM, 256, 1, 6, 3, 1>;
  }
  return nullptr;
}

// Tile geometry of a block for tiles of `tile` modes and a grid of at most
// `max_ctas` persistent CTAs.  Returns the grid size.
static int make_geom(const Block &b, int tile, int max_ctas, TileGeom *g) {
  const int dim = b.dim;
  g->n_inner = b.n[dim - 1];
  g->n_mid = dim == 3 ? b.n[1] : 1;
  g->n_outer = b.n[0];
  g->n_rows = (long long)b.n[0] * g->n_mid;
  g->cpr = (g->n_inner + tile - 1) / tile;
  g->n_tiles = g->n_rows * g->cpr;
  g->kb_outer = b.kb[0];
  g->kb_mid = dim == 3 ? b.kb[1] : 0;
  g->kb_inner = b.kb[dim - 1];
  long long grid = max_ctas;
  if (grid > g->n_tiles) grid = g->n_tiles;
  // keep the chunk of a CTA fixed from one row to the next when possible
  if (grid >= g->cpr) grid -= grid % g->cpr;
  if (grid < 1) grid = 1;
  g->d_row = grid / g->cpr;
  g->d_chunk = int(grid % g->cpr);
  g->d_a = int(g->d_row / g->n_mid);
  g->d_b = int(g->d_row % g->n_mid);
  return int(grid);
}

static void fill_params(const bri17_plan *p, const Block &b, ApplyParams *ap) {
  const int dim = b.dim;
  ap->tab_outer = p->tab[0].dev;
  ap->N_outer = p->shape[0];
  ap->tab_mid = dim == 3 ? p->tab[1].dev : p->tab[0].dev;
  ap->N_mid = dim == 3 ? p->shape[1] : 1;
  ap->tab_inner = p->tab[dim - 1].dev;
  ap->N_inner = p->shape[dim - 1];
  ap->mu = p->mu;
  ap->scaling = p->scaling;
  ap->out_scale = 1.0;
  ap->stage_outer = 0;
  ap->u = nullptr;
  ap->f = nullptr;
  ap->u_stride = ap->f_stride = 0;
  ap->u_mstride = ap->f_mstride = 1;
  ap->dot_partial = nullptr;
  ap->herm_n = 0;
}

static int check_launch(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return fail(BRI17_ERR_CUDA, std::string(what) + " launch: " + cudaGetErrorString(e));
  return BRI17_OK;
}

// Row tiling pads every row to a multiple of the tile; beyond a few per cent of idle
// lanes the flat mapping is used instead.
static bool rows_fill_tiles(const Block &b, int tile) {
  const int n_inner = b.n[b.dim - 1];
  const long long padded = (long long)((n_inner + tile - 1) / tile) * tile;
  return padded * 100 <= (long long)n_inner * 103;
}

static int launch_apply_flat(bri17_plan *p, const Block &b, ApplyParams &ap, cudaStream_t stream,
                             int max_grid = 0) {
  constexpr int THREADS = 256, VEC = 2;
  const bool dot = ap.dot_partial != nullptr;
  void (*kern)(const ApplyParams) =
      dot ? (b.dim == 3 ? modal_stiffness_apply_flat_kernel<3, THREADS, VEC, 2, true>
                        : modal_stiffness_apply_flat_kernel<2, THREADS, VEC, 2, true>)
          : (b.dim == 3 ? modal_stiffness_apply_flat_kernel<3, THREADS, VEC, 2, false>
                        : modal_stiffness_apply_flat_kernel<2, THREADS, VEC, 2, false>);
  const int n_mid = b.dim == 3 ? b.n[1] : 0;
  size_t smem = size_t(3) * (size_t(b.n[0]) + n_mid + b.n[b.dim - 1]) * sizeof(double);
  if (smem <= 48 * 1024) ap.stage_outer = 1; else { smem = 0; ap.stage_outer = 0; }
  int occ = 0;
  BRI17_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem));
  if (occ < 1) return fail(BRI17_ERR_CUDA, "flat apply kernel does not fit on an SM");
  make_geom(b, THREADS * VEC, p->sm_count * occ, &ap.g);
  const long long tiles = (b.modes + THREADS * VEC - 1) / (THREADS * VEC);
  int grid = int(std::min<long long>(tiles, (long long)p->sm_count * occ));
  if (max_grid > 0) grid = std::min(grid, max_grid);
  kern<<<grid, THREADS, smem, stream>>>(ap);
  p->last_grid = grid; p->last_block = THREADS; p->last_smem = int64_t(smem);
  p->last_flat = 1;
  p->launches++;
  return check_launch("modal_stiffness_apply_flat");
}

// Stiffness apply that also returns sum_k w_k Re(u^_k^H f^_k) in *dot_out (device scalar);
// `scratch` receives one partial sum per CTA, so the grid is capped at scratch_count.
int launch_apply_dot(bri17_plan *p, const Block &b, const void *u, void *f, int64_t u_stride,
                     int64_t f_stride, double out_scale, int herm_n, double *dot_out, double *scratch,
                     int scratch_count, cudaStream_t stream) {
  if (b.modes == 0) {
    BRI17_CUDA_TRY(cudaMemsetAsync(dot_out, 0, sizeof(double), stream));
    return BRI17_OK;
  }
  ApplyParams ap;
  fill_params(p, b, &ap);
  ap.u = static_cast<const double2 *>(u);
  ap.f = static_cast<double2 *>(f);
  ap.u_stride = u_stride;
  ap.f_stride = f_stride;
  ap.out_scale = out_scale;
  ap.dot_partial = scratch;
  ap.herm_n = herm_n;
  int rc = launch_apply_flat(p, b, ap, stream, scratch_count);
  if (rc) return rc;
  dot_finish_kernel<<<1, 256, 0, stream>>>(scratch, int(p->last_grid.load()), dot_out);
  p->launches++;
  return check_launch("dot_finish");
}

int launch_apply(bri17_plan *p, const Block &b, const void *u, void *f, int64_t u_stride,
                 int64_t f_stride, double out_scale, cudaStream_t stream) {
  const int vi = p->apply_variant < 0 ? default_variant(p) : p->apply_variant;
  const Variant &v = kVariants[vi];
  const bool flat = p->mapping == 2 || (p->mapping == 0 && !rows_fill_tiles(b, v.threads * v.vec));
  if (flat) {
    ApplyParams fp;
    fill_params(p, b, &fp);
    fp.u = static_cast<const double2 *>(u);
    fp.f = static_cast<double2 *>(f);
    fp.u_stride = u_stride;
    fp.f_stride = f_stride;
    fp.out_scale = out_scale;
    return launch_apply_flat(p, b, fp, stream);
  }
  p->last_flat = 0;
  ApplyKernel kern = b.dim == 3 ? apply_kernel_for<3>(vi) : apply_kernel_for<2>(vi);
  if (!kern) return fail(BRI17_ERR_UNSUPPORTED, "unknown apply variant");

  ApplyParams ap;
  fill_params(p, b, &ap);
  ap.u = static_cast<const double2 *>(u);
  ap.f = static_cast<double2 *>(f);
  ap.u_stride = u_stride;
  ap.f_stride = f_stride;
  ap.out_scale = out_scale;

  // shared memory: slow-axis tables of the local range, if they fit 48 KB
  const int n_outer = b.n[0], n_mid = b.dim == 3 ? b.n[1] : 0;
  size_t smem = size_t(3) * (n_outer + n_mid) * sizeof(double);
  if (smem <= 48 * 1024) ap.stage_outer = 1; else smem = 0;

  int occ = 0;
  BRI17_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, v.threads, smem));
  if (occ < 1) return fail(BRI17_ERR_CUDA, "apply kernel does not fit on an SM");
  const int grid = make_geom(b, v.threads * v.vec, p->sm_count * occ, &ap.g);

  kern<<<grid, v.threads, smem, stream>>>(ap);
  p->last_grid = grid; p->last_block = v.threads; p->last_smem = int64_t(smem);
  p->launches++;
  return check_launch("modal_stiffness_apply");
}

int launch_index_map(bri17_plan *p, const Block &b, int32_t *k_out, cudaStream_t stream) {
  TileGeom g;
  const int grid = make_geom(b, 512, p->sm_count * 8, &g);
  // same choice of mapping as the apply kernels, so that this exposes what they derive
  const bool flat = p->mapping == 2 || (p->mapping == 0 && !rows_fill_tiles(b, 512));
  if (flat) {
    const long long tiles = (b.modes + 511) / 512;
    const int fgrid = int(std::min<long long>(tiles, (long long)p->sm_count * 8));
    if (b.dim == 3) freq_index_map_flat_kernel<3><<<fgrid, 256, 0, stream>>>(g, k_out);
    else freq_index_map_flat_kernel<2><<<fgrid, 256, 0, stream>>>(g, k_out);
  } else if (b.dim == 3) freq_index_map_kernel<3><<<grid, 256, 0, stream>>>(g, k_out);
  else freq_index_map_kernel<2><<<grid, 256, 0, stream>>>(g, k_out);
  p->launches++;
  return check_launch("freq_index_map");
}

int launch_stiffness_field(bri17_plan *p, const Block &b, void *K, cudaStream_t stream) {
  ApplyParams ap;
  fill_params(p, b, &ap);
  ap.f = static_cast<double2 *>(K);
  const int grid = make_geom(b, 256, p->sm_count * 4, &ap.g);
  if (b.dim == 3) modal_field_kernel<3, 1><<<grid, 256, 0, stream>>>(ap);
  else modal_field_kernel<2, 1><<<grid, 256, 0, stream>>>(ap);
  p->launches++;
  return check_launch("modal_stiffness_field");
}

int launch_strain_field(bri17_plan *p, const Block &b, void *B, cudaStream_t stream) {
  ApplyParams ap;
  fill_params(p, b, &ap);
  ap.f = static_cast<double2 *>(B);
  const int grid = make_geom(b, 256, p->sm_count * 8, &ap.g);
  if (b.dim == 3) modal_field_kernel<3, 0><<<grid, 256, 0, stream>>>(ap);
  else modal_field_kernel<2, 0><<<grid, 256, 0, stream>>>(ap);
  p->launches++;
  return check_launch("modal_strain_displacement_field");
}

int launch_strain_apply(bri17_plan *p, const Block &b, const void *u, void *eps,
                        int64_t u_stride, int64_t e_stride, double out_scale,
                        cudaStream_t stream) {
  ApplyParams ap;
  fill_params(p, b, &ap);
  ap.u = static_cast<const double2 *>(u);
  ap.f = static_cast<double2 *>(eps);
  ap.u_stride = u_stride;
  ap.f_stride = e_stride;
  ap.out_scale = out_scale;
  const int grid = make_geom(b, 512, p->sm_count * 2, &ap.g);
  if (b.dim == 3) strain_displacement_kernel<3><<<grid, 256, 0, stream>>>(ap);
  else strain_displacement_kernel<2><<<grid, 256, 0, stream>>>(ap);
  p->launches++;
  return check_launch("strain_displacement_apply");
}

int launch_modal_solve(bri17_plan *p, const Block &b, int mode, const void *in, void *out,
                       int64_t in_cs, int64_t in_ms, int64_t out_cs, int64_t out_ms, cudaStream_t stream) {
  ApplyParams ap;
  fill_params(p, b, &ap);
  ap.u = static_cast<const double2 *>(in);
  ap.f = static_cast<double2 *>(out);
  ap.u_stride = in_cs; ap.u_mstride = in_ms;
  ap.f_stride = out_cs; ap.f_mstride = out_ms;
  // Two modes per thread for the lighter maps (more loads in flight), one for eigenstress -> strain
  // (register budget: no spills at 128).  Option "solve_variant" = 1 (3-D, K^-1 and eigenstress ->
  // displacement): one mode per thread within 85 registers, three CTAs (24 warps) per SM instead of
  // two -- MEASURED SLOWER (5600 vs 5648 and 4956 vs 5572 GB/s, profiles/r02_measurements.md), so
  // these maps are not occupancy-bound either; kept as an option, off.
  const bool dense = b.dim == 3 && p->solve_variant && (mode == 0 || mode == 1);
  const int vec = dense ? 1 : ((mode == 0 || mode == 3 || (mode == 1 && b.dim == 2)) ? 2 : 1);
  const int grid = make_geom(b, 256 * vec, p->sm_count * (dense ? 3 : 2), &ap.g);
  if (b.dim == 3) {
    if (mode == 0 && dense) modal_solve_kernel<3, 0, 1, 3><<<grid, 256, 0, stream>>>(ap);
    else if (mode == 0) modal_solve_kernel<3, 0, 2><<<grid, 256, 0, stream>>>(ap);
    else if (mode == 1 && dense) modal_solve_kernel<3, 1, 1, 3><<<grid, 256, 0, stream>>>(ap);
    else if (mode == 1) modal_solve_kernel<3, 1, 1><<<grid, 256, 0, stream>>>(ap);
    else if (mode == 2) modal_solve_kernel<3, 2, 1><<<grid, 256, 0, stream>>>(ap);
    else modal_solve_kernel<3, 3, 2><<<grid, 256, 0, stream>>>(ap);
  } else {
    if (mode == 0) modal_solve_kernel<2, 0, 2><<<grid, 256, 0, stream>>>(ap);
    else if (mode == 1) modal_solve_kernel<2, 1, 2><<<grid, 256, 0, stream>>>(ap);
    else if (mode == 2) modal_solve_kernel<2, 2, 1><<<grid, 256, 0, stream>>>(ap);
    else modal_solve_kernel<2, 3, 2><<<grid, 256, 0, stream>>>(ap);
  }
  p->launches++;
  return check_launch("modal_solve");
}

// Host replay of the tile cursor of one persistent CTA (diagnostic, no device needed):
// writes (tile, row, chunk, a, b) per visited tile.  Lets the CPU tests check that the
// division-free stepping covers every tile exactly once with consistent indices.
int walk_tiles_host(const Block &b, int tile_modes, int max_ctas, int cta, int64_t *out, int cap, int *grid_out) {
  TileGeom g;
  const int grid = make_geom(b, tile_modes, max_ctas, &g);
  if (grid_out) *grid_out = grid;
  if (cta < 0 || cta >= grid) return 0;
  TileCursor cur;
  int n = 0;
  for (cur.init(g, unsigned(cta)); cur.valid(g); cur.next(g, unsigned(grid))) {
    if (n < cap) {
      out[5 * n + 0] = cur.tile; out[5 * n + 1] = cur.row; out[5 * n + 2] = cur.chunk;
      out[5 * n + 3] = cur.a; out[5 * n + 4] = cur.b;
    }
    n++;
  }
  return n;
}

}  // namespace bri17b200
