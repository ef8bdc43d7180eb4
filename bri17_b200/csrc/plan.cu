// plan.cu -- plan lifetime, per-axis tables, per-mode host API, C-ABI glue.
//
// The per-axis tables are the ONLY place transcendental functions are
// evaluated.  They are computed on the host with libm in exactly the
// operation order of the reference (bri17.hpp:259-263 for phi/chi/psi,
// :218-221 for c/s), because 2(1-cos b)/h^2 is ill-conditioned at low
// frequency: one ulp of cos(b) moves phi by 1.5e-12 (relative) at N=512, so a
// device cos() would break the 1e-12 per-mode parity gate.  This translation
// unit is compiled with -ffp-contract=off for the same reason.
#include <cmath>
#include <cstring>
#include <numbers>

#include "internal.h"
#include "mode_math.h"

namespace bri17b200 {

static thread_local std::string g_last_error;

void set_error(const std::string &msg) { g_last_error = msg; }
int fail(int code, const std::string &msg) {
  g_last_error = msg;
  return code;
}

// bri17.hpp:258-264 and :217-222, evaluated for every grid line of one axis.
static void fill_axis_tables(AxisTables &t, int n, double L) {
  t.n = n;
  t.host.resize(size_t(TAB_COUNT) * n);
  double *phi = t.host.data() + size_t(TAB_PHI) * n;
  double *chi = t.host.data() + size_t(TAB_CHI) * n;
  double *psi = t.host.data() + size_t(TAB_PSI) * n;
  double *c = t.host.data() + size_t(TAB_C) * n;
  double *s = t.host.data() + size_t(TAB_S) * n;
  double *al = t.host.data() + size_t(TAB_ALPHA) * n;
  double *sa = t.host.data() + size_t(TAB_SINA) * n;
  for (int k = 0; k < n; k++) {
    const double h = L / n;                                   // :259
    const double beta = 2 * std::numbers::pi_v<double> * k / n;  // :260
    phi[k] = 2 * (1 - std::cos(beta)) / h / h;                // :261
    chi[k] = (2 + std::cos(beta)) / 3;                        // :262
    psi[k] = std::sin(beta) / h;                              // :263
    const double alpha = std::numbers::pi_v<double> * k / n;  // :218
    c[k] = std::cos(alpha);                                   // :220
    s[k] = std::sin(alpha) * n / L;                           // :221
    al[k] = alpha;                                            // summed per mode by the host path (:219)
    sa[k] = std::sin(alpha);                                  // device: e^{i sum(alpha)} = prod (c + i sa), :224
  }
}

int make_block(const bri17_plan *p, const int *k_begin, const int *local_shape, Block *b) {
  b->dim = p->dim;
  b->modes = 1;
  for (int d = 0; d < 3; d++) { b->n[d] = 1; b->kb[d] = 0; }
  for (int d = 0; d < p->dim; d++) {
    b->kb[d] = k_begin ? k_begin[d] : 0;
    b->n[d] = local_shape ? local_shape[d] : p->shape[d];
    if (b->n[d] < 0 || b->kb[d] < 0 || int64_t(b->kb[d]) + b->n[d] > p->shape[d])
      return fail(BRI17_ERR_INVALID_ARG,
                  "block [k_begin, k_begin+local_shape) exceeds the grid along axis " +
                      std::to_string(d));
    b->modes *= b->n[d];
  }
  return BRI17_OK;
}

}  // namespace bri17b200

using namespace bri17b200;

static int check_dev_ptr(const void *ptr, const char *what) {
  if (!ptr) return fail(BRI17_ERR_INVALID_ARG, std::string(what) + " is NULL");
  if (reinterpret_cast<uintptr_t>(ptr) % 16)
    return fail(BRI17_ERR_INVALID_ARG, std::string(what) + " must be 16-byte aligned");
  return BRI17_OK;
}

// Runs `body` with the plan's device current, restoring the caller's device.
template <typename F>
static int on_device(bri17_plan *p, F body) {
  if (p->device < 0)
    return fail(BRI17_ERR_CUDA, "plan was created with BRI17_DEVICE_NONE: whole-grid operators "
                                "need a CUDA device (there is no CPU fallback)");
  int prev = 0;
  BRI17_CUDA_TRY(cudaGetDevice(&prev));
  if (prev != p->device) BRI17_CUDA_TRY(cudaSetDevice(p->device));
  int rc = body();
  if (prev != p->device) cudaSetDevice(prev);
  return rc;
}

extern "C" {

const char *bri17_last_error(void) { return g_last_error.c_str(); }
int bri17_version(void) { return BRI17_VERSION; }
void bri17_set_last_error(const char *msg) { g_last_error = msg ? msg : ""; }

int bri17_plan_create(bri17_plan **out, int dim, const int *shape, const double *L,
                      double mu, double nu, int device) {
  if (!out) return fail(BRI17_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  if (dim != 2 && dim != 3)  // bri17.hpp:35-36
    return fail(BRI17_ERR_INVALID_ARG, "dim must be 2 or 3");
  if (!shape || !L) return fail(BRI17_ERR_INVALID_ARG, "shape/L is NULL");
  for (int d = 0; d < dim; d++)
    if (shape[d] < 1) return fail(BRI17_ERR_INVALID_ARG, "shape entries must be >= 1");

  if (device == BRI17_DEVICE_NONE) {  // per-mode host API only (bri17.hpp:212-292)
    auto *p = new bri17_plan;
    p->dim = dim; p->mu = mu; p->nu = nu; p->device = device;
    p->scaling = mu / (1. - 2. * nu);
    for (int d = 0; d < dim; d++) {
      p->shape[d] = shape[d]; p->L[d] = L[d];
      fill_axis_tables(p->tab[d], shape[d], L[d]);
    }
    *out = p;
    return BRI17_OK;
  }

  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(BRI17_ERR_CUDA,
                std::string("no CUDA device available (there is no CPU fallback): ") +
                    cudaGetErrorString(e));
  if (device < 0 || device >= count)
    return fail(BRI17_ERR_INVALID_ARG, "device ordinal out of range");

  int prev = 0;
  BRI17_CUDA_TRY(cudaGetDevice(&prev));
  BRI17_CUDA_TRY(cudaSetDevice(device));

  auto *p = new bri17_plan;
  p->dim = dim;
  p->mu = mu;
  p->nu = nu;
  p->scaling = mu / (1. - 2. * nu);  // bri17.hpp:266
  p->device = device;
  for (int d = 0; d < dim; d++) { p->shape[d] = shape[d]; p->L[d] = L[d]; }

  int rc = BRI17_OK;
  cudaError_t ce = cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, device);
  for (int d = 0; d < dim && ce == cudaSuccess; d++) {
    fill_axis_tables(p->tab[d], shape[d], L[d]);
    const size_t bytes = p->tab[d].host.size() * sizeof(double);
    ce = cudaMalloc(&p->tab[d].dev, bytes);
    if (ce == cudaSuccess)
      ce = cudaMemcpy(p->tab[d].dev, p->tab[d].host.data(), bytes, cudaMemcpyHostToDevice);
  }
  if (ce != cudaSuccess) {
    rc = fail(BRI17_ERR_CUDA, std::string("plan creation: ") + cudaGetErrorString(ce));
    bri17_plan_destroy(p);
    p = nullptr;
  }
  cudaSetDevice(prev);
  *out = p;
  return rc;
}

int bri17_plan_destroy(bri17_plan *p) {
  if (!p) return BRI17_OK;
  if (p->device < 0) { delete p; return BRI17_OK; }
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(p->device);
  free_host_stages(p);
  for (int d = 0; d < 3; d++)
    if (p->tab[d].dev) cudaFree(p->tab[d].dev);
  cudaSetDevice(prev);
  delete p;
  return BRI17_OK;
}

int bri17_plan_set_option(bri17_plan *p, const char *key, int64_t value) {
  if (!p || !key) return fail(BRI17_ERR_INVALID_ARG, "plan/key is NULL");
  std::lock_guard<std::mutex> lock(p->host_mutex);  // the host_* options free the staging set
  if (!std::strcmp(key, "apply_variant")) {
    if (value < -1 || value >= num_variants())
      return fail(BRI17_ERR_INVALID_ARG, "apply_variant out of range");
    p->apply_variant = int(value);
  } else if (!std::strcmp(key, "mapping")) {
    if (value < 0 || value > 2) return fail(BRI17_ERR_INVALID_ARG, "mapping must be 0 (auto), 1 (rows) or 2 (flat)");
    p->mapping = int(value);
  } else if (!std::strcmp(key, "host_chunk_rows")) {
    if (value < 0) return fail(BRI17_ERR_INVALID_ARG, "host_chunk_rows < 0");
    p->host_chunk_rows = value;
    free_host_stages(p);
  } else if (!std::strcmp(key, "solve_variant")) {
    p->solve_variant = value != 0;
  } else if (!std::strcmp(key, "host_zero_copy")) {
    p->host_zero_copy = value != 0;
    free_host_stages(p);
  } else if (!std::strcmp(key, "host_streams")) {
    if (value < 1 || value > 8) return fail(BRI17_ERR_INVALID_ARG, "host_streams must be 1..8");
    p->host_streams = int(value);
    free_host_stages(p);
  } else {
    return fail(BRI17_ERR_INVALID_ARG, std::string("unknown option ") + key);
  }
  return BRI17_OK;
}

int bri17_plan_get_info(const bri17_plan *p, const char *key, int64_t *value) {
  if (!p || !key || !value) return fail(BRI17_ERR_INVALID_ARG, "NULL argument");
  if (!std::strcmp(key, "num_variants")) *value = num_variants();
  else if (!std::strcmp(key, "apply_variant")) *value = p->apply_variant < 0 ? default_variant(p) : p->apply_variant;
  else if (!std::strcmp(key, "solve_variant")) *value = p->solve_variant;
  else if (!std::strcmp(key, "sm_count")) *value = p->sm_count;
  else if (!std::strcmp(key, "last_grid")) *value = p->last_grid.load();
  else if (!std::strcmp(key, "last_block")) *value = p->last_block.load();
  else if (!std::strcmp(key, "last_smem")) *value = p->last_smem.load();
  else if (!std::strcmp(key, "launches")) *value = p->launches.load();
  else if (!std::strcmp(key, "last_flat")) *value = p->last_flat.load();
  else if (!std::strcmp(key, "device")) *value = p->device;
  else if (!std::strcmp(key, "table_bytes")) {
    int64_t b = 0;
    for (int d = 0; d < p->dim; d++) b += int64_t(p->tab[d].host.size()) * 8;
    *value = b;
  } else return fail(BRI17_ERR_INVALID_ARG, std::string("unknown info key ") + key);
  return BRI17_OK;
}

int bri17_plan_get_tables(const bri17_plan *p, int axis, double *phi, double *chi,
                          double *psi, double *c, double *s) {
  if (!p || axis < 0 || axis >= p->dim) return fail(BRI17_ERR_INVALID_ARG, "bad plan/axis");
  const AxisTables &t = p->tab[axis];
  double *outs[5] = {phi, chi, psi, c, s};
  for (int w = 0; w < 5; w++)
    if (outs[w]) std::memcpy(outs[w], t.h(w), sizeof(double) * t.n);
  return BRI17_OK;
}

// Hooke::modal_stiffness, bri17.hpp:247-292, from the tables.
int bri17_modal_stiffness_mode_f64(const bri17_plan *p, const int *k, double *K) {
  if (!p || !k || !K) return fail(BRI17_ERR_INVALID_ARG, "NULL argument");
  double phi[3], chi[3], psi[3];
  for (int d = 0; d < p->dim; d++) {
    if (k[d] < 0 || k[d] >= p->shape[d])
      return fail(BRI17_ERR_INVALID_ARG, "frequency index outside [0, shape)");
    phi[d] = p->tab[d].h(TAB_PHI)[k[d]];
    chi[d] = p->tab[d].h(TAB_CHI)[k[d]];
    psi[d] = p->tab[d].h(TAB_PSI)[k[d]];
  }
  const double mu = p->mu, scaling = p->scaling;
  if (p->dim == 2) {
    const double H_00 = phi[0] * chi[1];
    const double H_11 = chi[0] * phi[1];
    const double K_diag = mu * (H_00 + H_11);
    const double Kr[4] = {scaling * H_00 + K_diag, scaling * psi[0] * psi[1], 0,
                          scaling * H_11 + K_diag};
    K[0] = Kr[0]; K[2] = Kr[1]; K[4] = Kr[1]; K[6] = Kr[3];
    K[1] = K[3] = K[5] = K[7] = 0.;
  } else {
    const double H_00 = phi[0] * chi[1] * chi[2];
    const double H_11 = chi[0] * phi[1] * chi[2];
    const double H_22 = chi[0] * chi[1] * phi[2];
    const double K_diag = mu * (H_00 + H_11 + H_22);
    const double K00 = scaling * H_00 + K_diag;
    const double K01 = scaling * psi[0] * psi[1] * chi[2];
    const double K02 = scaling * psi[0] * chi[1] * psi[2];
    const double K11 = scaling * H_11 + K_diag;
    const double K12 = scaling * chi[0] * psi[1] * psi[2];
    const double K22 = scaling * H_22 + K_diag;
    const double Kr[9] = {K00, K01, K02, K01, K11, K12, K02, K12, K22};
    for (int i = 0; i < 9; i++) { K[2 * i] = Kr[i]; K[2 * i + 1] = 0.; }
  }
  return BRI17_OK;
}

// Hooke::modal_strain_displacement, bri17.hpp:212-236.  The prefactor needs
// sin/cos of the SUM of the half angles (:224), which is not a per-axis
// quantity: evaluated here with libm exactly like the reference.
int bri17_modal_strain_displacement_mode_f64(const bri17_plan *p, const int *k, double *B) {
  if (!p || !k || !B) return fail(BRI17_ERR_INVALID_ARG, "NULL argument");
  double c[3] = {1., 1., 1.}, s[3] = {0., 0., 0.};
  double sum_alpha = 0.;
  for (int d = 0; d < p->dim; d++) {
    if (k[d] < 0 || k[d] >= p->shape[d])
      return fail(BRI17_ERR_INVALID_ARG, "frequency index outside [0, shape)");
    const double alpha = std::numbers::pi_v<double> * k[d] / p->shape[d];
    sum_alpha += alpha;
    c[d] = p->tab[d].h(TAB_C)[k[d]];
    s[d] = p->tab[d].h(TAB_S)[k[d]];
  }
  const double pre[2] = {-2 * std::sin(sum_alpha), 2 * std::cos(sum_alpha)};
  for (int ri = 0; ri < 2; ri++) {
    if (p->dim == 2) {
      B[0 + ri] = pre[ri] * s[0] * c[1];
      B[2 + ri] = pre[ri] * c[0] * s[1];
    } else {
      B[0 + ri] = pre[ri] * s[0] * c[1] * c[2];
      B[2 + ri] = pre[ri] * c[0] * s[1] * c[2];
      B[4 + ri] = pre[ri] * c[0] * c[1] * s[2];
    }
  }
  return BRI17_OK;
}

// Hooke::modal_eigenstress_to_opposite_strain, bri17.hpp:308-355, one mode on the host.
int bri17_modal_eigenstress_to_opposite_strain_mode_f64(const bri17_plan *p, const int *k,
                                                        const double *tau, double *eta) {
  if (!p || !k || !tau || !eta) return fail(BRI17_ERR_INVALID_ARG, "NULL argument");
  const int dim = p->dim, nsym = dim * (dim + 1) / 2;
  double phi[3], chi[3], psi[3], c[3], s[3];
  double sum_alpha = 0.;
  bool null_frequency = true;
  for (int d = 0; d < dim; d++) {
    if (k[d] < 0 || k[d] >= p->shape[d])
      return fail(BRI17_ERR_INVALID_ARG, "frequency index outside [0, shape)");
    null_frequency = null_frequency && k[d] == 0;
    phi[d] = p->tab[d].h(TAB_PHI)[k[d]]; chi[d] = p->tab[d].h(TAB_CHI)[k[d]];
    psi[d] = p->tab[d].h(TAB_PSI)[k[d]];
    c[d] = p->tab[d].h(TAB_C)[k[d]]; s[d] = p->tab[d].h(TAB_S)[k[d]];
    sum_alpha += p->tab[d].h(TAB_ALPHA)[k[d]];
  }
  if (null_frequency) {  // :336-339
    for (int i = 0; i < 2 * nsym; i++) eta[i] = 0.;
    return BRI17_OK;
  }
  const Cplx pre{-2 * std::sin(sum_alpha), 2 * std::cos(sum_alpha)};
  Cplx t[6], e[6];
  for (int i = 0; i < nsym; i++) t[i] = {tau[2 * i], tau[2 * i + 1]};
  if (dim == 2) {
    double K[2][2];
    Cplx B[2], u[2];
    stiffness_entries<2>(phi, chi, psi, p->mu, p->scaling, K);
    strain_displacement_entries<2>(c, s, pre, B);
    eigenstress_to_displacement<2>(t, B, K, u);
    displacement_to_strain<2>(B, u, e);
  } else {
    double K[3][3];
    Cplx B[3], u[3];
    stiffness_entries<3>(phi, chi, psi, p->mu, p->scaling, K);
    strain_displacement_entries<3>(c, s, pre, B);
    eigenstress_to_displacement<3>(t, B, K, u);
    displacement_to_strain<3>(B, u, e);
  }
  for (int i = 0; i < nsym; i++) { eta[2 * i] = e[i].re; eta[2 * i + 1] = e[i].im; }
  return BRI17_OK;
}

static int solve_entry(bri17_plan *p, int mode, const void *in, void *out, const int *k_begin,
                       const int *local_shape, int64_t in_cs, int64_t in_ms, int64_t out_cs,
                       int64_t out_ms, void *stream) {
  if (!p) return fail(BRI17_ERR_INVALID_ARG, "plan is NULL");
  Block b;
  int rc = make_block(p, k_begin, local_shape, &b);
  if (rc) return rc;
  if (b.modes == 0) return BRI17_OK;
  if ((rc = check_dev_ptr(in, "input field")) || (rc = check_dev_ptr(out, "output field"))) return rc;
  if (in_ms == 0) in_ms = 1;
  if (out_ms == 0) out_ms = 1;
  if (in_cs == 0) in_cs = in_ms == 1 ? b.modes : 1;
  if (out_cs == 0) out_cs = out_ms == 1 ? b.modes : 1;
  if (in_cs < 1 || in_ms < 1 || out_cs < 1 || out_ms < 1)
    return fail(BRI17_ERR_INVALID_ARG, "strides must be positive");
  return on_device(p, [&] {
    return launch_modal_solve(p, b, mode, in, out, in_cs, in_ms, out_cs, out_ms, cudaStream_t(stream));
  });
}

int bri17_modal_stiffness_solve_f64(bri17_plan *p, const void *f, void *u, const int *k_begin,
                                    const int *local_shape, int64_t comp_stride, int64_t mode_stride,
                                    void *stream) {
  return solve_entry(p, 0, f, u, k_begin, local_shape, comp_stride, mode_stride, comp_stride, mode_stride, stream);
}

int bri17_eigenstress_to_displacement_f64(bri17_plan *p, const void *tau, void *u, const int *k_begin,
                                          const int *local_shape, int64_t tau_cs, int64_t tau_ms,
                                          int64_t u_cs, int64_t u_ms, void *stream) {
  return solve_entry(p, 1, tau, u, k_begin, local_shape, tau_cs, tau_ms, u_cs, u_ms, stream);
}

int bri17_eigenstress_to_opposite_strain_f64(bri17_plan *p, const void *tau, void *eta, const int *k_begin,
                                             const int *local_shape, int64_t comp_stride,
                                             int64_t mode_stride, void *stream) {
  return solve_entry(p, 2, tau, eta, k_begin, local_shape, comp_stride, mode_stride, comp_stride, mode_stride, stream);
}

int bri17_eigenstress_to_force_f64(bri17_plan *p, const void *tau, void *f, const int *k_begin,
                                   const int *local_shape, int64_t tau_cs, int64_t tau_ms,
                                   int64_t f_cs, int64_t f_ms, void *stream) {
  return solve_entry(p, 3, tau, f, k_begin, local_shape, tau_cs, tau_ms, f_cs, f_ms, stream);
}

int bri17_modal_stiffness_apply_dot_f64(bri17_plan *p, const void *u, void *f, const int *k_begin,
                                        const int *local_shape, int64_t comp_stride, double out_scale,
                                        int hermitian_n, double *dot_dev, double *scratch_dev,
                                        int scratch_count, void *stream) {
  if (!p) return fail(BRI17_ERR_INVALID_ARG, "plan is NULL");
  if (!dot_dev || !scratch_dev || scratch_count < 1)
    return fail(BRI17_ERR_INVALID_ARG, "dot_dev/scratch_dev is NULL or scratch_count < 1");
  if (hermitian_n < 0) return fail(BRI17_ERR_INVALID_ARG, "hermitian_n < 0");
  Block b;
  int rc = make_block(p, k_begin, local_shape, &b);
  if (rc) return rc;
  if (b.modes > 0 && ((rc = check_dev_ptr(u, "u_hat_dev")) || (rc = check_dev_ptr(f, "f_hat_dev")))) return rc;
  if (comp_stride == 0) comp_stride = b.modes;
  if (comp_stride < b.modes)
    return fail(BRI17_ERR_INVALID_ARG, "comp_stride smaller than the block");
  return on_device(p, [&] {
    return launch_apply_dot(p, b, u, f, comp_stride, comp_stride, out_scale, hermitian_n, dot_dev,
                            scratch_dev, scratch_count, cudaStream_t(stream));
  });
}

int bri17_modal_stiffness_apply_f64(bri17_plan *p, const void *u, void *f, const int *k_begin,
                                    const int *local_shape, int64_t comp_stride,
                                    double out_scale, void *stream) {
  if (!p) return fail(BRI17_ERR_INVALID_ARG, "plan is NULL");
  Block b;
  int rc = make_block(p, k_begin, local_shape, &b);
  if (rc) return rc;
  if (b.modes == 0) return BRI17_OK;
  if ((rc = check_dev_ptr(u, "u_hat_dev")) || (rc = check_dev_ptr(f, "f_hat_dev"))) return rc;
  if (comp_stride == 0) comp_stride = b.modes;
  if (comp_stride < b.modes)
    return fail(BRI17_ERR_INVALID_ARG, "comp_stride smaller than the block");
  return on_device(p, [&] {
    return launch_apply(p, b, u, f, comp_stride, comp_stride, out_scale, cudaStream_t(stream));
  });
}

int bri17_modal_stiffness_apply_host_f64(bri17_plan *p, const void *u, void *f,
                                         const int *k_begin, const int *local_shape,
                                         int64_t comp_stride, double out_scale) {
  if (!p) return fail(BRI17_ERR_INVALID_ARG, "plan is NULL");
  Block b;
  int rc = make_block(p, k_begin, local_shape, &b);
  if (rc) return rc;
  if (b.modes == 0) return BRI17_OK;
  if (!u || !f) return fail(BRI17_ERR_INVALID_ARG, "host buffer is NULL");
  if (comp_stride == 0) comp_stride = b.modes;
  if (comp_stride < b.modes)
    return fail(BRI17_ERR_INVALID_ARG, "comp_stride smaller than the block");
  // one staging set per plan: concurrent host-buffer calls on the same plan take turns
  std::lock_guard<std::mutex> lock(p->host_mutex);
  return on_device(p, [&] { return apply_host(p, b, u, f, comp_stride, out_scale); });
}

int bri17_modal_stiffness_field_f64(bri17_plan *p, void *K, const int *k_begin,
                                    const int *local_shape, void *stream) {
  if (!p) return fail(BRI17_ERR_INVALID_ARG, "plan is NULL");
  Block b;
  int rc = make_block(p, k_begin, local_shape, &b);
  if (rc) return rc;
  if (b.modes == 0) return BRI17_OK;
  if ((rc = check_dev_ptr(K, "K_dev"))) return rc;
  return on_device(p, [&] { return launch_stiffness_field(p, b, K, cudaStream_t(stream)); });
}

int bri17_modal_strain_displacement_field_f64(bri17_plan *p, void *B, const int *k_begin,
                                              const int *local_shape, void *stream) {
  if (!p) return fail(BRI17_ERR_INVALID_ARG, "plan is NULL");
  Block b;
  int rc = make_block(p, k_begin, local_shape, &b);
  if (rc) return rc;
  if (b.modes == 0) return BRI17_OK;
  if ((rc = check_dev_ptr(B, "B_dev"))) return rc;
  return on_device(p, [&] { return launch_strain_field(p, b, B, cudaStream_t(stream)); });
}

int bri17_strain_displacement_apply_f64(bri17_plan *p, const void *u, void *eps,
                                        const int *k_begin, const int *local_shape,
                                        int64_t u_stride, int64_t eps_stride, double out_scale,
                                        void *stream) {
  if (!p) return fail(BRI17_ERR_INVALID_ARG, "plan is NULL");
  Block b;
  int rc = make_block(p, k_begin, local_shape, &b);
  if (rc) return rc;
  if (b.modes == 0) return BRI17_OK;
  if ((rc = check_dev_ptr(u, "u_hat_dev")) || (rc = check_dev_ptr(eps, "eps_hat_dev"))) return rc;
  if (u_stride == 0) u_stride = b.modes;
  if (eps_stride == 0) eps_stride = b.modes;
  if (u_stride < b.modes || eps_stride < b.modes)
    return fail(BRI17_ERR_INVALID_ARG, "component stride smaller than the block");
  return on_device(p, [&] {
    return launch_strain_apply(p, b, u, eps, u_stride, eps_stride, out_scale, cudaStream_t(stream));
  });
}

int bri17_debug_walk_tiles(const bri17_plan *p, const int *k_begin, const int *local_shape, int tile_modes,
                           int max_ctas, int cta, int64_t *out, int cap, int *grid) {
  if (!p || !out || tile_modes < 1 || max_ctas < 1) return -1;
  Block b;
  if (make_block(p, k_begin, local_shape, &b)) return -1;
  if (b.modes == 0) return 0;
  return walk_tiles_host(b, tile_modes, max_ctas, cta, out, cap, grid);
}

int bri17_freq_index_map(bri17_plan *p, int32_t *k_out, const int *k_begin,
                         const int *local_shape, void *stream) {
  if (!p) return fail(BRI17_ERR_INVALID_ARG, "plan is NULL");
  Block b;
  int rc = make_block(p, k_begin, local_shape, &b);
  if (rc) return rc;
  if (b.modes == 0) return BRI17_OK;
  if (!k_out) return fail(BRI17_ERR_INVALID_ARG, "k_out_dev is NULL");
  return on_device(p, [&] { return launch_index_map(p, b, k_out, cudaStream_t(stream)); });
}

}  // extern "C"
