// api.cu -- host side of libcdpr_b200.so: the C ABI of include/cdpr_b200.h over the sm_100a kernels.
// No CPU fallback anywhere: without a CUDA device every entry point fails.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/cdpr_b200.h"
#include "common.cuh"
#include "launch.h"
#include "misc_kernels.cuh"

using namespace cdpr;

static thread_local std::string g_create_error;
static const std::vector<FastEntry> &fast_table();

struct cdpr_batch {
  cdpr_config cfg;
  int device = 0;
  long long n = 0, np = 0;
  bool general = false;  // controller state in the general layout (time-stamp rings, biquad state): flex and HBM variants
  bool flex = false;     // the on-chip full-semantics kernel (step_flex.cuh): per-instance modes and commands
  bool flex_capable = false;
  int flex_tpb = 0, flex_ps = 0, flex_ds = 0, flex_nf = 0, flex_unroll = 2, flex_lanes = 1;
  std::string detail;    // cdpr_kernel_detail
  bool flexr = false;    // ... in its rebuilt form (step_flexr.cuh): at most two biquad stages per filter, no leg model
  bool flexr_hold = false;
  size_t flex_smem = 0;
  DevLayout L{};
  RobotConsts rc{};
  PidConsts pc[2]{};
  double fir[2][kMaxDbuf]{};
  double dmom[2][3]{};
  bool dmom_ok[2]{};
  bool force_fir = false;  // CDPR_OPT_DTERM_FIR
  int mode = MODE_POSITION;
  bool vel_pending = false, pos_pending = false;
  int sec = 0, nsec = 0, dt_ns = 0;
  long long step_count = 0;
  bool sine_on = false;
  int pub_shape = 0;     // CDPR_OPT_PUBLISHER_SHAPE: wave form of the in-kernel publisher
  int sine_period = 10;
  double sine_time = 0.0, sine_pub_dt = 0.01;
  double *snap_peers[8] = {nullptr};
  int n_snap_peers = 0;
  bool snap_multimem = false;
  long long snap_stride = 0, snap_offset = 0;
  long long snap_every = 0, snap_written = 0, snap_capacity = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // async mode: everything that is not a step kernel (reset, uploads, layout kernels, downloads) runs on a
  // high-priority stream of its own, so it is not starved behind step kernels that fill every SM (another handle's)
  cudaStream_t io_stream = nullptr;
  cudaEvent_t io_done = nullptr, main_done = nullptr;
  bool io_pending = false, main_pending = false;
  bool timed = false;
  bool timing = true;  // CDPR_OPT_KERNEL_TIMING: event records around the launches
  bool async_copies = false;  // host-buffer calls only enqueue; the caller synchronises (pinned buffers)
  bool targets_uniform = false;  // all cables of an instance hold the same velocity target (zeros after Load, or written by the sine publisher)
  long long launches = 0;
  void *stage = nullptr;
  // cdpr_update: pinned, device-mapped host memory the step kernel publishes into and the command scatter reads from
  double *pub_host = nullptr;
  float *cmd_host[2] = {nullptr, nullptr};
  double *pub_ptr[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // set for the duration of one cdpr_update
  size_t stage_bytes = 0;
  double *cost_dev = nullptr;
  float *cmd_dev = nullptr;
  size_t cmd_dev_bytes = 0;
  std::vector<void *> allocs;
  std::string err;
};

#define CK(h, call)                                                                         \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess) {                                                                \
      (h)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                        \
      return CDPR_ERR_CUDA;                                                                 \
    }                                                                                       \
  } while (0)

static int fail(cdpr_handle h, int code, const std::string &msg) {
  if (h) h->err = msg; else g_create_error = msg;
  return code;
}

// ---------------------------------------------------------------------------------------------
// constants
// ---------------------------------------------------------------------------------------------
extern "C" int cdpr_config_default(cdpr_config *cfg, int n_cables) {
  if (!cfg || n_cables < 1 || n_cables > CDPR_MAX_CABLES) return CDPR_ERR_BAD_ARG;
  std::memset(cfg, 0, sizeof(*cfg));
  cfg->n_cables = n_cables;
  // sdf/cube.sdf: frame anchors :383,559,735,911; platform anchors :458,634,810,986 relative to :310.
  // Cables 4..7 (synthetic 8-cable extension): same corners on the lower frame face z = 0.
  const double sx[4] = {-1, -1, 1, 1}, sy[4] = {-1, 1, 1, -1};
  for (int c = 0; c < n_cables; ++c) {
    cfg->frame_anchor[c][0] = 0.3 * sx[c & 3];
    cfg->frame_anchor[c][1] = 0.3 * sy[c & 3];
    cfg->frame_anchor[c][2] = c < 4 ? 0.6 : 0.0;
    cfg->platform_anchor[c][0] = 0.03 * sx[c & 3];
    cfg->platform_anchor[c][1] = 0.03 * sy[c & 3];
    cfg->platform_anchor[c][2] = 0.0;
  }
  cfg->home_pos[2] = 0.3;
  cfg->home_quat[0] = 1.0;
  cfg->mass = 1.0;
  cfg->inertia[0] = cfg->inertia[1] = cfg->inertia[2] = 1.0;
  cfg->gravity[2] = -9.8;
  cfg->cable_damping = 1.0;
  cfg->effort_limit = 100.0;
  cfg->dt = 0.001;
  cdpr_pid_params &v = cfg->vel_pid, &p = cfg->pos_pid;  // launch/cdpr_gazebo.launch:19-39
  v.forward_gain = 0.0; v.p_gain = 200.0; v.i_gain = 20.0; v.d_gain = 1.0;
  v.d_degree = 2; v.d_buffer_length = 11; v.i_limit = 100.0; v.cmd_limit = 100.0;
  v.p_cutoff = 0.1; v.p_quality = 0.707; v.p_cascade = 0;
  v.d_cutoff = 0.1; v.d_quality = 0.707; v.d_cascade = 0;
  p = v;
  p.forward_gain = 0.0; p.p_gain = 200.0; p.i_gain = 70.0; p.d_gain = 80.0;
  p.p_cascade = p.d_cascade = 0;
  cfg->velocity_epsilon = -0.001;
  cfg->sine_publish_hz = 100.0;
  // leg links and passive joints (cube.sdf:344-518); off by default = the reduced model
  cfg->leg_model = 0;
  cfg->leg_link_mass = 0.001; cfg->leg_link_inertia = 0.001;
  cfg->leg_cable_com = 0.51961524;  // l/2, l = |(0.6, 0.6, 0.6)| (gen_cdpr.py:104,124-125)
  cfg->passive_damping = 0.01;
  cfg->slider_lower = -0.51961524; cfg->slider_upper = 0.51961524; cfg->slider_velocity_limit = 10.0;
  for (int c = 0; c < n_cables; ++c) {
    // gen_cdpr.py:113-125,152: the leg frame is the rotation about z x u_fp that takes z onto the frame -> platform
    // direction u_fp; rev_X turns about its first column (cube.sdf:390)
    double ufp[3], n = 0.0;
    for (int k = 0; k < 3; ++k) { ufp[k] = cfg->home_pos[k] + cfg->platform_anchor[c][k] - cfg->frame_anchor[c][k]; n += ufp[k] * ufp[k]; }
    n = std::sqrt(n);
    for (int k = 0; k < 3; ++k) ufp[k] /= n;
    double ax[3] = {-ufp[1], ufp[0], 0.0};
    const double sn = std::sqrt(ax[0] * ax[0] + ax[1] * ax[1]), cs = ufp[2];
    if (sn > 0.0) { ax[0] /= sn; ax[1] /= sn; }
    cfg->leg_axis_frame[c][0] = cs + ax[0] * ax[0] * (1.0 - cs);
    cfg->leg_axis_frame[c][1] = ax[2] * sn + ax[1] * ax[0] * (1.0 - cs);
    cfg->leg_axis_frame[c][2] = -ax[1] * sn + ax[2] * ax[0] * (1.0 - cs);
    cfg->leg_axis_cable[c][2] = 1.0;     // "0 0 1", model frame (SDF 1.4)
    cfg->leg_axis_platform[c][0] = 1.0;  // "1 0 0"
  }
  return CDPR_OK;
}

static void biquad_coeffs(double fc, double q, double out[5]) {  // Filter.h:130-140 with fs = 1 (Pid.cpp:34)
  const double k = std::tan(M_PI * fc / 1.0);
  const double den = k * k + k / q + 1.0;
  out[0] = k * k / den;
  out[1] = 2 * out[0];
  out[2] = out[0];
  out[3] = 2 * (k * k - 1.0) / den;
  out[4] = (k * k - k / q + 1.0) / den;
}

static void make_pid_consts(const cdpr_pid_params &p, PidConsts &o) {  // Pid.cpp:63-77
  std::memset(&o, 0, sizeof(o));
  o.kf = p.forward_gain; o.kp = p.p_gain; o.ki = p.i_gain; o.kd = p.d_gain;
  o.i_max = std::fabs(p.i_limit); o.i_min = -std::fabs(p.i_limit);
  o.i_max_over_ki = o.i_max / p.i_gain; o.i_min_over_ki = o.i_min / p.i_gain;
  o.cmd_max = std::fabs(p.cmd_limit); o.cmd_min = -std::fabs(p.cmd_limit);
  o.degree = p.d_degree; o.len = p.d_buffer_length; o.p_casc = p.p_cascade; o.d_casc = p.d_cascade;
  if (p.p_cascade > 0) biquad_coeffs(p.p_cutoff, p.p_quality, o.pf);
  if (p.d_cascade > 0) biquad_coeffs(p.d_cutoff, p.d_quality, o.df);
}

// FIR form of Pid::derive for uniformly spaced samples: weights w_j with
//   d/dt p(now) = sum_j w_j * y_j,   p = least-squares polynomial of `degree` through the window.
// Row 1 of (V^T V)^-1 V^T on the abscissae x_j = (j - (len-1)) / (len-1), divided by the span.
static void fir_weights(int degree, int len, double dt, double *w) {
  for (int j = 0; j < kMaxDbuf; ++j) w[j] = 0.0;
  if (degree < 1 || len < 2) return;
  const int m = degree + 1;
  long double G[kMaxDegree + 1][kMaxDegree + 2];
  std::vector<long double> x(len);
  for (int j = 0; j < len; ++j) x[j] = (long double)(j - (len - 1)) / (long double)(len - 1);
  for (int r = 0; r < m; ++r) {
    for (int q = 0; q < m; ++q) {
      long double s = 0;
      for (int j = 0; j < len; ++j) s += powl(x[j], r + q);
      G[r][q] = s;
    }
    G[r][m] = (r == 1) ? 1.0L : 0.0L;
  }
  for (int col = 0; col < m; ++col) {
    int piv = col;
    for (int r = col + 1; r < m; ++r) if (fabsl(G[r][col]) > fabsl(G[piv][col])) piv = r;
    for (int q = 0; q <= m; ++q) std::swap(G[col][q], G[piv][q]);
    for (int r = 0; r < m; ++r) {
      if (r == col) continue;
      const long double f = G[r][col] / G[col][col];
      for (int q = col; q <= m; ++q) G[r][q] -= f * G[col][q];
    }
  }
  const long double span = (long double)(len - 1) * (long double)dt;
  for (int j = 0; j < len; ++j) {
    long double s = 0;
    for (int k = 0; k < m; ++k) s += (G[k][m] / G[k][k]) * powl(x[j], k);
    w[j] = (double)(s / span);
  }
}

// The FIR weights of a degree <= 2 fit are a quadratic in the sample position p = j + 1 (oldest 1 .. newest len).
// Returns false when they are not (degree > 2): the kernel then keeps the plain FIR.
static bool fir_as_quadratic(const double *w, int len, double *abc) {
  const long double K = -1.0L;  // position = j - K = j + 1
  // least-squares quadratic through (k_j, w_j) via normal equations in long double, then check the residual
  long double s[5] = {0, 0, 0, 0, 0}, t[3] = {0, 0, 0};
  for (int j = 0; j < len; ++j) {
    const long double k = j - K;
    long double p = 1;
    for (int q = 0; q < 5; ++q) { s[q] += p; if (q < 3) t[q] += p * w[j]; p *= k; }
  }
  long double M[3][4] = {{s[0], s[1], s[2], t[0]}, {s[1], s[2], s[3], t[1]}, {s[2], s[3], s[4], t[2]}};
  for (int col = 0; col < 3; ++col) {
    int piv = col;
    for (int r = col + 1; r < 3; ++r) if (fabsl(M[r][col]) > fabsl(M[piv][col])) piv = r;
    if (M[piv][col] == 0) return false;
    for (int q = 0; q < 4; ++q) std::swap(M[col][q], M[piv][q]);
    for (int r = 0; r < 3; ++r) {
      if (r == col) continue;
      const long double f = M[r][col] / M[col][col];
      for (int q = col; q < 4; ++q) M[r][q] -= f * M[col][q];
    }
  }
  long double c[3] = {M[0][3] / M[0][0], M[1][3] / M[1][1], M[2][3] / M[2][2]};
  long double worst = 0, scale = 0;
  for (int j = 0; j < len; ++j) {
    const long double k = j - K;
    worst = std::max(worst, fabsl(c[0] + c[1] * k + c[2] * k * k - w[j]));
    scale = std::max(scale, fabsl((long double)w[j]));
  }
  for (int q = 0; q < 3; ++q) abc[q] = (double)c[q];
  return worst <= 1e-13L * std::max(scale, (long double)1e-300);
}

static void quat_rot_host(const double q[4], double R[3][3]) {
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  R[0][0] = 1 - 2 * (y * y + z * z); R[0][1] = 2 * (x * y - w * z); R[0][2] = 2 * (x * z + w * y);
  R[1][0] = 2 * (x * y + w * z); R[1][1] = 1 - 2 * (x * x + z * z); R[1][2] = 2 * (y * z - w * x);
  R[2][0] = 2 * (x * z - w * y); R[2][1] = 2 * (y * z + w * x); R[2][2] = 1 - 2 * (x * x + y * y);
}

static int make_robot_consts(const cdpr_config &c, RobotConsts &o, std::string &err) {
  std::memset(&o, 0, sizeof(o));
  double R[3][3];
  quat_rot_host(c.home_quat, R);
  for (int i = 0; i < c.n_cables; ++i) {
    double d[3];
    for (int k = 0; k < 3; ++k) {
      o.a[i][k] = c.frame_anchor[i][k];
      o.b[i][k] = c.platform_anchor[i][k];
    }
    for (int k = 0; k < 3; ++k) {
      const double r = R[k][0] * o.b[i][0] + R[k][1] * o.b[i][1] + R[k][2] * o.b[i][2];
      d[k] = o.a[i][k] - c.home_pos[k] - r;
    }
    o.home_len[i] = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  }
  for (int k = 0; k < 3; ++k) o.mg[k] = c.mass * c.gravity[k];
  o.h = c.dt; o.h_over_m = c.dt / c.mass; o.half_h = 0.5 * c.dt;
  const double *I = c.inertia;
  const double a = I[0], b = I[1], cc = I[2], d = I[3], e = I[4], f = I[5];  // [a d e; d b f; e f c]
  const double det = a * (b * cc - f * f) - d * (d * cc - f * e) + e * (d * f - b * e);
  if (!(det != 0.0) || !(c.mass > 0.0)) { err = "singular inertia or non-positive mass"; return CDPR_ERR_BAD_ARG; }
  for (int k = 0; k < 6; ++k) o.ib[k] = I[k];
  o.ib_inv[0] = (b * cc - f * f) / det; o.ib_inv[1] = (a * cc - e * e) / det; o.ib_inv[2] = (a * b - d * d) / det;
  o.ib_inv[3] = (e * f - d * cc) / det; o.ib_inv[4] = (d * f - e * b) / det; o.ib_inv[5] = (d * e - a * f) / det;
  o.diag_inertia = (d == 0.0 && e == 0.0 && f == 0.0) ? 1 : 0;
  o.spec = 0;
  if (o.diag_inertia) o.spec |= SPEC_DIAG;
  if (o.diag_inertia && I[0] == I[1] && I[1] == I[2]) o.spec |= SPEC_ISO;
  bool bz0 = true;
  for (int i = 0; i < c.n_cables; ++i) bz0 = bz0 && (c.platform_anchor[i][2] == 0.0);
  if (bz0) o.spec |= SPEC_BZ0;
  // cables c and c + NC/2 from the same platform anchor to frame anchors above each other (same x, y)?
  if (c.n_cables >= 2 && c.n_cables % 2 == 0) {
    const int half = c.n_cables / 2;
    bool pair = true;
    for (int i = 0; i < half; ++i) {
      for (int k = 0; k < 3; ++k) pair = pair && (c.platform_anchor[i][k] == c.platform_anchor[i + half][k]);
      pair = pair && (c.frame_anchor[i][0] == c.frame_anchor[i + half][0]) && (c.frame_anchor[i][1] == c.frame_anchor[i + half][1]);
      o.pair_dz[i] = c.frame_anchor[i + half][2] - c.frame_anchor[i][2];
    }
    if (pair) o.spec |= SPEC_PAIR;
  }
  o.cdamp = c.cable_damping; o.effort_limit = c.effort_limit; o.vel_eps = c.velocity_epsilon;
  o.effort_limit_abs = c.effort_limit >= 0.0 ? c.effort_limit : INFINITY;
  o.mass = c.mass;
  for (int k = 0; k < 3; ++k) o.grav[k] = c.gravity[k];
  o.leg_model = c.leg_model;
  if (c.leg_model) {
    if (!(c.leg_link_mass >= 0.0) || !(c.leg_link_inertia >= 0.0) || !(c.passive_damping >= 0.0)) { err = "leg constants must be non-negative"; return CDPR_ERR_BAD_ARG; }
    o.leg_sI = std::sqrt(c.leg_link_inertia); o.leg_s2I = std::sqrt(2.0 * c.leg_link_inertia);
    o.leg_sm = std::sqrt(c.leg_link_mass); o.leg_s2m = std::sqrt(2.0 * c.leg_link_mass);
    o.leg_sc = std::sqrt(c.passive_damping); o.leg_lc = c.leg_cable_com;
    for (int i = 0; i < c.n_cables; ++i) {
      // body triad of the leg at the home pose: e2 = (u x x0)/c, e1 = e2 x u, u; the rev_Zpf axis keeps these components
      double d[3], u[3], x0[3], e1[3], e2[3], nx = 0.0;
      for (int k = 0; k < 3; ++k) { x0[k] = c.leg_axis_frame[i][k]; nx += x0[k] * x0[k]; }
      if (!(nx > 0.0)) { err = "leg_axis_frame must be non-zero"; return CDPR_ERR_BAD_ARG; }
      for (int k = 0; k < 3; ++k) {
        const double r = R[k][0] * o.b[i][0] + R[k][1] * o.b[i][1] + R[k][2] * o.b[i][2];
        d[k] = o.a[i][k] - c.home_pos[k] - r;
      }
      const double L = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
      for (int k = 0; k < 3; ++k) u[k] = d[k] / L;
      const double s = u[0] * x0[0] + u[1] * x0[1] + u[2] * x0[2], cc = std::sqrt(1.0 - s * s);
      if (!(cc > 1e-6)) { err = "leg_axis_frame must not be parallel to the leg"; return CDPR_ERR_BAD_ARG; }
      for (int k = 0; k < 3; ++k) e1[k] = (x0[k] - s * u[k]) / cc;
      e2[0] = u[1] * e1[2] - u[2] * e1[1]; e2[1] = u[2] * e1[0] - u[0] * e1[2]; e2[2] = u[0] * e1[1] - u[1] * e1[0];
      const double *a3 = c.leg_axis_cable[i];
      o.leg_alpha[i][0] = e1[0] * a3[0] + e1[1] * a3[1] + e1[2] * a3[2];
      o.leg_alpha[i][1] = e2[0] * a3[0] + e2[1] * a3[1] + e2[2] * a3[2];
      o.leg_alpha[i][2] = u[0] * a3[0] + u[1] * a3[1] + u[2] * a3[2];
      for (int k = 0; k < 3; ++k) { o.leg_x0[i][k] = x0[k]; o.leg_a1[i][k] = c.leg_axis_platform[i][k]; }
    }
  }
  return CDPR_OK;
}

// ---------------------------------------------------------------------------------------------
// lifecycle
// ---------------------------------------------------------------------------------------------
static int dev_alloc(cdpr_handle h, void **p, size_t bytes) {
  cudaError_t e = cudaMalloc(p, bytes);
  if (e != cudaSuccess) { h->err = std::string("cudaMalloc: ") + cudaGetErrorString(e); return CDPR_ERR_NOMEM; }
  h->allocs.push_back(*p);
  return CDPR_OK;
}

static int ensure_stage(cdpr_handle h, size_t bytes) {
  if (bytes <= h->stage_bytes) return CDPR_OK;
  if (h->stage) cudaFree(h->stage);
  h->stage = nullptr; h->stage_bytes = 0;
  cudaError_t e = cudaMalloc(&h->stage, bytes);
  if (e != cudaSuccess) { h->err = std::string("cudaMalloc(stage): ") + cudaGetErrorString(e); return CDPR_ERR_NOMEM; }
  h->stage_bytes = bytes;
  return CDPR_OK;
}

static inline cudaError_t sync_unless_async(cdpr_handle h) { return h->async_copies ? cudaSuccess : cudaStreamSynchronize(h->stream); }

// Stream for a non-step operation: the handle's stream, or in async mode the io stream ordered after the last step
static cudaStream_t io_begin(cdpr_handle h) {
  if (!h->async_copies) return h->stream;
  if (h->main_pending) { cudaStreamWaitEvent(h->io_stream, h->main_done, 0); h->main_pending = false; }
  return h->io_stream;
}
static void io_end(cdpr_handle h) {
  if (!h->async_copies) return;
  cudaEventRecord(h->io_done, h->io_stream);
  h->io_pending = true;
}
// before / after work on the handle's own stream that touches the state (step kernels, checkpoints, rollouts)
static void main_begin(cdpr_handle h) {
  if (h->io_pending) { cudaStreamWaitEvent(h->stream, h->io_done, 0); h->io_pending = false; }
}
static void main_end(cdpr_handle h) {
  if (!h->async_copies) return;
  cudaEventRecord(h->main_done, h->stream);
  h->main_pending = true;
}

static inline unsigned grid_for(long long n, int tpb) { return (unsigned)((n + tpb - 1) / tpb); }

static int reset_to_load_state(cdpr_handle h, cudaStream_t st) {
  const DevLayout &L = h->L;
  CK(h, cudaMemsetAsync(L.cab, 0, sizeof(double) * L.nc * CAB_F * L.np, st));
  CK(h, cudaMemsetAsync(L.pid, 0, sizeof(double) * L.nc * 2 * PID_F * L.np, st));
  CK(h, cudaMemsetAsync(L.win_y, 0, sizeof(double) * L.nc * 2 * L.len * L.np, st));
  if (L.mom) CK(h, cudaMemsetAsync(L.mom, 0, sizeof(double) * L.nc * 2 * 3 * L.np, st));
  if (L.win_x) CK(h, cudaMemsetAsync(L.win_x, 0, sizeof(double) * L.nc * 2 * L.len * L.np, st));
  if (L.filt) CK(h, cudaMemsetAsync(L.filt, 0, sizeof(double) * L.nc * 2 * 2 * L.casc * 4 * L.np, st));
  const cdpr_config &c = h->cfg;
  k_init_state<<<grid_for(L.np, 256), 256, 0, st>>>(L, h->rc, c.home_pos[0], c.home_pos[1], c.home_pos[2], c.home_quat[0],
                                                           c.home_quat[1], c.home_quat[2], c.home_quat[3],
                                                           (unsigned)c.vel_pid.d_buffer_length, (unsigned)c.pos_pid.d_buffer_length, h->general ? 1 : 0);
  CK(h, cudaGetLastError());
  h->mode = MODE_POSITION;  // CdprGazeboPlugin.cpp:154
  h->vel_pending = h->pos_pending = false;
  h->sec = h->nsec = 0;
  h->step_count = 0;
  h->sine_time = 0.0;
  h->targets_uniform = true;
  return CDPR_OK;
}

static void flex_prepare_any(cdpr_handle h) {
  if (h->flexr) flexr_prepare(h->L.nc, h->flex_nf, h->flexr_hold, h->flex_lanes);
  else flex_prepare(h->L.nc, h->flex_nf, h->flex_unroll, h->flex_lanes);
}

extern "C" int cdpr_create(const cdpr_config *cfg, int64_t n_instances, int device, cdpr_handle *out) {
  if (!cfg || !out) return fail(nullptr, CDPR_ERR_BAD_ARG, "null argument");
  *out = nullptr;
  if (cfg->n_cables < 1 || cfg->n_cables > CDPR_MAX_CABLES)
    return fail(nullptr, CDPR_ERR_BAD_CABLE_COUNT, "invalid joint count");  // CdprGazeboPlugin.cpp:167-169
  if (n_instances < 1 || n_instances > (1LL << 31) - 256) return fail(nullptr, CDPR_ERR_BAD_ARG, "n_instances out of range");
  const cdpr_pid_params *pp[2] = {&cfg->vel_pid, &cfg->pos_pid};
  for (int k = 0; k < 2; ++k) {
    if (pp[k]->d_buffer_length < 2 || pp[k]->d_buffer_length > CDPR_MAX_DBUF || pp[k]->d_degree < 0 ||
        pp[k]->d_degree > CDPR_MAX_DEGREE || pp[k]->d_degree + 1 > pp[k]->d_buffer_length || pp[k]->p_cascade < 0 ||
        pp[k]->p_cascade > CDPR_MAX_CASCADE || pp[k]->d_cascade < 0 || pp[k]->d_cascade > CDPR_MAX_CASCADE)
      return fail(nullptr, CDPR_ERR_BAD_ARG, "pid parameters out of range");
  }
  const double ns = cfg->dt * 1e9;
  if (!(cfg->dt > 0.0) || std::fabs(ns - std::round(ns)) > 1e-6 || ns > 1e9)
    return fail(nullptr, CDPR_ERR_BAD_ARG, "dt must be a whole number of nanoseconds in (0, 1 s]");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev)
    return fail(nullptr, CDPR_ERR_NO_DEVICE, "no CUDA device (this library has no CPU fallback)");
  cdpr_handle h = new (std::nothrow) cdpr_batch();
  if (!h) return fail(nullptr, CDPR_ERR_NOMEM, "out of host memory");
  h->cfg = *cfg;
  h->device = device;
  int rc = make_robot_consts(*cfg, h->rc, h->err);
  if (rc != CDPR_OK) { g_create_error = h->err; delete h; return rc; }
  make_pid_consts(cfg->vel_pid, h->pc[PID_VEL]);
  make_pid_consts(cfg->pos_pid, h->pc[PID_POS]);
  fir_weights(cfg->vel_pid.d_degree, cfg->vel_pid.d_buffer_length, cfg->dt, h->fir[PID_VEL]);
  fir_weights(cfg->pos_pid.d_degree, cfg->pos_pid.d_buffer_length, cfg->dt, h->fir[PID_POS]);
  h->dmom_ok[PID_VEL] = fir_as_quadratic(h->fir[PID_VEL], cfg->vel_pid.d_buffer_length, h->dmom[PID_VEL]);
  h->dmom_ok[PID_POS] = fir_as_quadratic(h->fir[PID_POS], cfg->pos_pid.d_buffer_length, h->dmom[PID_POS]);
  if (!(cfg->sine_publish_hz > 0.0) || !std::isfinite(cfg->sine_publish_hz)) {
    g_create_error = "sine_publish_hz must be positive and finite"; delete h; return CDPR_ERR_BAD_ARG;
  }
  h->dt_ns = (int)std::llround(ns);
  h->sine_pub_dt = 1.0 / cfg->sine_publish_hz;  // sinevelocitytest.cpp:48
  h->sine_period = (int)std::llround(h->sine_pub_dt / cfg->dt);
  if (h->sine_period < 1) h->sine_period = 1;
  // Which kernel variant? The fast one needs: hold impossible, no filters, cmdLimit != 0, window 11 for both Pids.
  const bool fast_ok = cfg->velocity_epsilon < 0.0 && cfg->vel_pid.p_cascade == 0 && cfg->vel_pid.d_cascade == 0 &&
                       cfg->pos_pid.p_cascade == 0 && cfg->pos_pid.d_cascade == 0 && cfg->vel_pid.cmd_limit != 0.0 &&
                       cfg->pos_pid.cmd_limit != 0.0 && cfg->vel_pid.i_gain >= 0.0 && cfg->pos_pid.i_gain >= 0.0 && cfg->vel_pid.d_buffer_length == 11 && cfg->pos_pid.d_buffer_length == 11 &&
                       (cfg->n_cables == 4 || cfg->n_cables == 8) && cfg->leg_model == 0;
  h->general = !fast_ok;
  // The on-chip full-semantics variant (step_flex.cuh): windows of 11 fitted with one degree, cmdLimit != 0, 4 or 8
  // cables, and a block shape whose controller state fits in shared memory.
  h->flex_ps = std::max(cfg->vel_pid.p_cascade, cfg->pos_pid.p_cascade);
  h->flex_ds = std::max(cfg->vel_pid.d_cascade, cfg->pos_pid.d_cascade);
  {
    const bool shape_ok = (cfg->n_cables == 4 || cfg->n_cables == 8) && cfg->vel_pid.d_buffer_length == 11 && cfg->pos_pid.d_buffer_length == 11 &&
                          cfg->vel_pid.d_degree == cfg->pos_pid.d_degree && cfg->vel_pid.cmd_limit != 0.0 && cfg->pos_pid.cmd_limit != 0.0;
    h->flex_nf = flex_stage_slots(h->flex_ps, h->flex_ds);
    h->flex_unroll = cfg->velocity_epsilon < 0.0 ? 4 : 2;
    // tuning override; unroll 4 is compiled without the hold test, so it is only allowed when hold is impossible
    if (const char *env = std::getenv("CDPR_FLEX_UNROLL")) h->flex_unroll = (std::atoi(env) >= 4 && cfg->velocity_epsilon < 0.0) ? 4 : 2;
    h->flex_tpb = flex_tpb();
    {  // CDPR_FLEX_LANES: tuning override of the number of lanes that share one robot (step_flex.cuh)
      const char *env = std::getenv("CDPR_FLEX_LANES");
      // measured (step_flex.cuh): two lanes per robot pay at 8 cables as soon as Pids can change inside a run or filters
      // shrink residency; the steady launch configuration and the 4-cable robot are faster with one thread per robot
      const int lanes_default = (cfg->n_cables == 8 && (cfg->velocity_epsilon >= 0.0 || h->flex_nf > 0)) ? 2 : 1;
      h->flex_lanes = flex_lanes_supported(cfg->n_cables, env ? std::atoi(env) : lanes_default);
    }
    h->flex_smem = shape_ok ? flex_smem_bytes(cfg->n_cables, h->flex_nf, h->flex_lanes) : 0;
    h->flex_capable = shape_ok && h->flex_smem <= 227u * 1024u;
    // The rebuilt kernel (step_flexr.cuh) wherever it is compiled; CDPR_FLEX_CLASSIC=1 keeps k_step_flex (A/B runs).
    const char *classic = std::getenv("CDPR_FLEX_CLASSIC");
    const char *env_lanes = std::getenv("CDPR_FLEX_LANES");
    const int rnf = std::max(h->flex_ps, h->flex_ds);  // k_step_flexr holds exactly the stages there are (0, 1 or 2 per filter)
    const int rl = flexr_lanes(cfg->n_cables, rnf, env_lanes ? std::atoi(env_lanes) : 2);
    // one coefficient set per filter: two Pids that both have the stage must share its coefficients (pc is filled above)
    const bool coefs_ok = (h->pc[0].p_casc == 0 || h->pc[1].p_casc == 0 || std::memcmp(h->pc[0].pf, h->pc[1].pf, sizeof(h->pc[0].pf)) == 0) &&
                          (h->pc[0].d_casc == 0 || h->pc[1].d_casc == 0 || std::memcmp(h->pc[0].df, h->pc[1].df, sizeof(h->pc[0].df)) == 0);
    if (h->flex_capable && rl > 0 && coefs_ok && cfg->leg_model == 0 && !(classic && std::atoi(classic) != 0)) {
      h->flexr = true;
      h->flexr_hold = cfg->velocity_epsilon >= 0.0;
      h->flex_lanes = rl;
      h->flex_nf = rnf;
      h->flex_smem = flexr_smem_bytes(cfg->n_cables, rnf, rl);
    }
  }
  h->flex = h->general && h->flex_capable;
  if (cfg->leg_model && !h->flex) {
    g_create_error = "leg_model = 1 runs in the flex kernel only (4 or 8 cables, windows of 11 with one degree, cmdLimit != 0)";
    delete h;
    return CDPR_ERR_UNSUPPORTED;
  }

  auto bail = [&](int code) { g_create_error = h->err; cdpr_destroy(h); return code; };
  if (cudaSetDevice(device) != cudaSuccess) { h->err = "cudaSetDevice failed"; return bail(CDPR_ERR_CUDA); }
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { h->err = "stream create failed"; return bail(CDPR_ERR_CUDA); }
  h->own_stream = true;
  cudaEventCreate(&h->ev0);
  cudaEventCreate(&h->ev1);
  {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);  // hi = numerically lowest = highest priority
    cudaStreamCreateWithPriority(&h->io_stream, cudaStreamNonBlocking, hi);
    cudaEventCreateWithFlags(&h->io_done, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->main_done, cudaEventDisableTiming);
  }
  h->n = n_instances;
  h->np = (n_instances + kTpb - 1) / kTpb * kTpb;
  DevLayout &L = h->L;
  L.n = (int)h->n; L.np = h->np; L.nc = cfg->n_cables;
  L.len = std::max(cfg->vel_pid.d_buffer_length, cfg->pos_pid.d_buffer_length);
  L.casc = std::max(std::max(cfg->vel_pid.p_cascade, cfg->vel_pid.d_cascade), std::max(cfg->pos_pid.p_cascade, cfg->pos_pid.d_cascade));
  const size_t col = sizeof(double) * (size_t)L.np;
  if ((rc = dev_alloc(h, (void **)&L.plat, col * 13))) return bail(rc);
  if ((rc = dev_alloc(h, (void **)&L.cab, col * L.nc * CAB_F))) return bail(rc);
  if ((rc = dev_alloc(h, (void **)&L.pid, col * L.nc * 2 * PID_F))) return bail(rc);
  if ((rc = dev_alloc(h, (void **)&L.win_y, col * L.nc * 2 * L.len))) return bail(rc);
  if (!h->general && (rc = dev_alloc(h, (void **)&L.mom, col * L.nc * 2 * 3))) return bail(rc);
  if (h->general) {
    if ((rc = dev_alloc(h, (void **)&L.win_x, col * L.nc * 2 * L.len))) return bail(rc);
    if (L.casc > 0 && (rc = dev_alloc(h, (void **)&L.filt, col * L.nc * 2 * 2 * L.casc * 4))) return bail(rc);
  }
  if ((rc = dev_alloc(h, (void **)&L.ctl, sizeof(uint32_t) * (size_t)L.np * L.nc))) return bail(rc);
  if ((rc = dev_alloc(h, (void **)&L.ictl, sizeof(uint32_t) * (size_t)L.np))) return bail(rc);
  if ((rc = dev_alloc(h, (void **)&L.sine, col * 3))) return bail(rc);
  if (cudaMemsetAsync(L.sine, 0, col * 3, h->stream) != cudaSuccess) { h->err = "memset failed"; return bail(CDPR_ERR_CUDA); }
  if ((rc = reset_to_load_state(h, h->stream))) return bail(rc);
  if (!h->general) {
    for (const FastEntry &e : fast_table()) {
      if (e.nc != cfg->n_cables) continue;
      cudaFuncSetAttribute(e.func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e.smem);
      cudaFuncSetAttribute(e.func, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
  }
  if (h->flex) flex_prepare_any(h);
  if (cudaStreamSynchronize(h->stream) != cudaSuccess) { h->err = "initialisation kernels failed"; return bail(CDPR_ERR_CUDA); }
  *out = h;
  return CDPR_OK;
}

extern "C" int cdpr_destroy(cdpr_handle h) {
  if (!h) return CDPR_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (void *p : h->allocs) cudaFree(p);
  if (h->stage) cudaFree(h->stage);
  if (h->pub_host) cudaFreeHost(h->pub_host);
  if (h->cmd_host[0]) cudaFreeHost(h->cmd_host[0]);
  if (h->cost_dev) cudaFree(h->cost_dev);
  if (h->cmd_dev) cudaFree(h->cmd_dev);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->io_stream) { cudaStreamSynchronize(h->io_stream); cudaStreamDestroy(h->io_stream); }
  if (h->io_done) cudaEventDestroy(h->io_done);
  if (h->main_done) cudaEventDestroy(h->main_done);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return CDPR_OK;
}

extern "C" int cdpr_reset(cdpr_handle h) {
  if (!h) return CDPR_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  int rc = reset_to_load_state(h, io_begin(h));
  io_end(h);
  if (rc) return rc;
  h->snap_written = 0;
  h->launches += 1;
  return CDPR_OK;
}

extern "C" const char *cdpr_last_error(cdpr_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

extern "C" int cdpr_set_stream(cdpr_handle h, void *cuda_stream) {
  if (!h) return CDPR_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->io_stream) cudaStreamSynchronize(h->io_stream);
  h->io_pending = h->main_pending = false;
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  h->stream = (cudaStream_t)cuda_stream;
  h->own_stream = false;
  return CDPR_OK;
}

extern "C" int cdpr_set_async(cdpr_handle h, int on) {
  if (!h) return CDPR_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  cudaStreamSynchronize(h->io_stream);
  h->io_pending = h->main_pending = false;
  h->async_copies = on != 0;
  return CDPR_OK;
}

extern "C" int cdpr_synchronize(cdpr_handle h) {
  if (!h) return CDPR_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  CK(h, cudaStreamSynchronize(h->stream));
  CK(h, cudaStreamSynchronize(h->io_stream));
  return CDPR_OK;
}

extern "C" int cdpr_set_option(cdpr_handle h, int option, int64_t value) {
  if (!h) return CDPR_ERR_BAD_ARG;
  switch (option) {
    case CDPR_OPT_KERNEL_TIMING:
      h->timing = value != 0;
      if (!h->timing) h->timed = false;
      return CDPR_OK;
    case CDPR_OPT_INDEPENDENT: {
      if (h->flex) return CDPR_OK;  // the flex variant is always per-instance
      if (value == 0) return CDPR_OK;
      if (!h->flex_capable) return fail(h, CDPR_ERR_UNSUPPORTED, "this configuration has no on-chip full-semantics kernel (needs 4 or 8 cables, windows of 11, cmdLimit != 0)");
      if (h->step_count != 0) return fail(h, CDPR_ERR_BAD_ARG, "switch to independent robots before the first step (or right after cdpr_reset)");
      cudaSetDevice(h->device);
      DevLayout &L = h->L;
      const size_t col = sizeof(double) * (size_t)L.np;
      int rc;
      if (!L.win_x && (rc = dev_alloc(h, (void **)&L.win_x, col * L.nc * 2 * L.len))) return rc;
      if (!L.filt && L.casc > 0 && (rc = dev_alloc(h, (void **)&L.filt, col * L.nc * 2 * 2 * L.casc * 4))) return rc;
      h->general = true; h->flex = true;
      flex_prepare_any(h);
      cudaStream_t st = io_begin(h);
      rc = reset_to_load_state(h, st);
      io_end(h);
      if (rc) return rc;
      CK(h, sync_unless_async(h));
      return CDPR_OK;
    }
    case CDPR_OPT_PUBLISHER_SHAPE:
      if (value != 0 && value != 1) return fail(h, CDPR_ERR_BAD_ARG, "publisher shape: 0 (sinevelocitytest) or 1 (squarevelocitytest)");
      h->pub_shape = (int)value;
      return CDPR_OK;
    case CDPR_OPT_DTERM_FIR:
      if (h->step_count != 0) return fail(h, CDPR_ERR_BAD_ARG, "the D-term form can only change before the first step");
      h->force_fir = value != 0;
      return CDPR_OK;
    default:
      return fail(h, CDPR_ERR_BAD_ARG, "unknown option");
  }
}

// ---------------------------------------------------------------------------------------------
// commands
// ---------------------------------------------------------------------------------------------
template <typename T>
static int scatter_cmd(cdpr_handle h, const T *host, const unsigned char *mask, int64_t n_instances, int n_axes, int field) {
  if (!h || !host) return CDPR_ERR_BAD_ARG;
  // the plugin drops a Joy message whose axes.size() != cWireCount (CdprGazeboPlugin.cpp:68,77)
  if (n_axes != h->L.nc) return fail(h, CDPR_ERR_BAD_LENGTH, "command length != cable count: dropped");
  if (n_instances != h->n) return fail(h, CDPR_ERR_BAD_ARG, "n_instances does not match the handle");
  if (mask && !h->flex)
    return fail(h, CDPR_ERR_UNSUPPORTED, "per-instance commands need the flex variant: cdpr_set_option(h, CDPR_OPT_INDEPENDENT, 1) before the first step");
  cudaSetDevice(h->device);
  const size_t bytes = sizeof(T) * (size_t)h->n * h->L.nc;
  const size_t mask_off = (bytes + 255) / 256 * 256;
  int rc = ensure_stage(h, mask_off + (mask ? (size_t)h->n : 0));
  if (rc) return rc;
  cudaStream_t st = io_begin(h);
  CK(h, cudaMemcpyAsync(h->stage, host, bytes, cudaMemcpyHostToDevice, st));
  const unsigned char *dmask = nullptr;
  if (mask) {
    CK(h, cudaMemcpyAsync((char *)h->stage + mask_off, mask, (size_t)h->n, cudaMemcpyHostToDevice, st));
    dmask = (const unsigned char *)h->stage + mask_off;
  }
  if (h->flex) {
    const unsigned bit = (field == CAB_VEL_TARGET) ? 4u : (field == CAB_POS_TARGET) ? 8u : 0u;
    k_scatter_cab_masked<T><<<grid_for(h->n, 256), 256, 0, st>>>(h->L, field, (const T *)h->stage, dmask, bit, field == CAB_FORCE_CMD ? 1 : 0);
  } else {
    k_scatter_cab<T><<<grid_for(h->n, 256), 256, 0, st>>>(h->L, field, (const T *)h->stage);
  }
  CK(h, cudaGetLastError());
  ++h->launches;
  io_end(h);
  CK(h, sync_unless_async(h));  // the caller may reuse its buffers
  return CDPR_OK;
}

extern "C" int cdpr_set_velocity_cmd_masked(cdpr_handle h, const float *axes, const unsigned char *mask, int64_t n_instances, int n_axes) {
  int rc = scatter_cmd<float>(h, axes, mask, n_instances, n_axes, CAB_VEL_TARGET);
  if (rc == CDPR_OK) { h->vel_pending = true; h->targets_uniform = false; }
  return rc;
}
extern "C" int cdpr_set_position_cmd_masked(cdpr_handle h, const float *axes, const unsigned char *mask, int64_t n_instances, int n_axes) {
  int rc = scatter_cmd<float>(h, axes, mask, n_instances, n_axes, CAB_POS_TARGET);
  if (rc == CDPR_OK) h->pos_pending = true;
  return rc;
}
extern "C" int cdpr_set_effort_cmd_masked(cdpr_handle h, const double *force, const unsigned char *mask, int64_t n_instances, int n_axes) {
  int rc = scatter_cmd<double>(h, force, mask, n_instances, n_axes, CAB_FORCE_CMD);
  if (rc == CDPR_OK) h->mode = MODE_FORCE;  // JointForceCalculator.h:92-95: immediate
  return rc;
}
extern "C" int cdpr_set_velocity_cmd(cdpr_handle h, const float *axes, int64_t n_instances, int n_axes) {
  return cdpr_set_velocity_cmd_masked(h, axes, nullptr, n_instances, n_axes);
}
extern "C" int cdpr_set_position_cmd(cdpr_handle h, const float *axes, int64_t n_instances, int n_axes) {
  return cdpr_set_position_cmd_masked(h, axes, nullptr, n_instances, n_axes);
}
extern "C" int cdpr_set_effort_cmd(cdpr_handle h, const double *force, int64_t n_instances, int n_axes) {
  return cdpr_set_effort_cmd_masked(h, force, nullptr, n_instances, n_axes);
}

extern "C" int cdpr_get_modes(cdpr_handle h, int32_t *modes) {
  if (!h || !modes) return CDPR_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  if (!h->flex) {  // batch-uniform mode; commands still pending are applied by the next step, like in the plugin
    cudaStreamSynchronize(h->stream);
    for (long long i = 0; i < h->n; ++i) modes[i] = h->mode;
    return CDPR_OK;
  }
  int rc = ensure_stage(h, sizeof(int) * (size_t)h->n);
  if (rc) return rc;
  cudaStream_t st = io_begin(h);
  k_modes<<<grid_for(h->n, 256), 256, 0, st>>>(h->L, (int *)h->stage);
  CK(h, cudaGetLastError());
  CK(h, cudaMemcpyAsync(modes, h->stage, sizeof(int) * (size_t)h->n, cudaMemcpyDeviceToHost, st));
  io_end(h);
  CK(h, sync_unless_async(h));
  return CDPR_OK;
}

extern "C" int cdpr_set_sine_cmd(cdpr_handle h, const double *amp, const double *freq, const double *phase, int64_t n_instances) {
  if (!h) return CDPR_ERR_BAD_ARG;
  if (!amp) { h->sine_on = false; return CDPR_OK; }
  if (n_instances != h->n) return fail(h, CDPR_ERR_BAD_ARG, "n_instances does not match the handle");
  cudaSetDevice(h->device);
  const size_t bytes = sizeof(double) * (size_t)h->n;
  std::vector<double> def;
  const double *src[3] = {amp, freq, phase};
  const double defaults[3] = {0.05, 0.1, 0.0};  // sinevelocitytest.cpp:8-9
  for (int k = 0; k < 3; ++k) {
    if (!src[k]) { def.assign((size_t)h->n, defaults[k]); src[k] = def.data(); }
    const bool temporary = (src[k] == def.data());
    cudaStream_t st = io_begin(h);
    CK(h, cudaMemcpyAsync(h->L.sine + (size_t)k * h->L.np, src[k], bytes, cudaMemcpyHostToDevice, st));
    io_end(h);
    if (temporary) CK(h, cudaStreamSynchronize(st)); else CK(h, sync_unless_async(h));
  }
  h->sine_on = true;
  h->sine_time = 0.0;
  return CDPR_OK;
}

// ---------------------------------------------------------------------------------------------
// stepping
// ---------------------------------------------------------------------------------------------
static int reset_pid(cdpr_handle h, int k) {
  const unsigned len = (unsigned)(k == PID_VEL ? h->cfg.vel_pid.d_buffer_length : h->cfg.pos_pid.d_buffer_length);
  k_reset_pid<<<grid_for(h->L.np, 256), 256, 0, h->stream>>>(h->L, k, len, h->general ? 1 : 0);
  CK(h, cudaGetLastError());
  ++h->launches;
  return CDPR_OK;
}

static void fill_args(cdpr_handle h, StepArgs &A, int k_steps, bool sine) {
  std::memset(&A, 0, sizeof(A));
  A.L = h->L; A.rc = h->rc; A.pc[0] = h->pc[0]; A.pc[1] = h->pc[1];
  A.live_idx = (h->mode == MODE_POSITION) ? PID_POS : PID_VEL;
  A.live = h->pc[A.live_idx];
  std::memcpy(A.fir, h->fir[A.live_idx], sizeof(A.fir));
  std::memcpy(A.fir2, h->fir, sizeof(A.fir2));
  std::memcpy(A.dmom, h->dmom[A.live_idx], sizeof(A.dmom));
  {  // D = a S0 + b S1 + c S2 over positions 1..LEN; one slide of the window, substituted into Kd * D
    const double a = A.dmom[0], b = A.dmom[1], c = A.dmom[2], kd = A.live.kd, len = (double)A.live.len;
    A.dk[0] = kd * (a + len * b + len * len * c);
    A.dk[1] = kd * (c - b);
    A.dk[2] = -2.0 * kd * c;
    A.dk[3] = -kd * a;
  }
  A.flex_ps = h->flex_ps; A.flex_ds = h->flex_ds;
  for (int j = 0; j < 21; ++j) { A.firx[j] = A.fir[j % 11]; A.firx0[j] = (j % 11 == 10) ? 0.0 : A.fir[j % 11]; }
  for (int k = 1; k >= 0; --k) {  // the velocity Pid's coefficients when both Pids have the stage (they are equal then)
    if (h->pc[k].p_casc > 0) std::memcpy(A.flex_pf, h->pc[k].pf, sizeof(A.flex_pf));
    if (h->pc[k].d_casc > 0) std::memcpy(A.flex_df, h->pc[k].df, sizeof(A.flex_df));
  }
  A.pub_pos = h->pub_ptr[0]; A.pub_vel = h->pub_ptr[1]; A.pub_eff = h->pub_ptr[2]; A.pub_pose = h->pub_ptr[3]; A.pub_twist = h->pub_ptr[4];
  A.effort_ge_cmd = h->rc.effort_limit_abs >= A.live.cmd_max ? 1 : 0;
  A.sat_thr = fmin(A.live.cmd_max, h->rc.effort_limit_abs);
  A.mode = h->mode; A.k_steps = k_steps; A.n0 = h->step_count;
  A.sec0 = h->sec; A.nsec0 = h->nsec; A.dt_ns = h->dt_ns; A.t0 = time_double(h->sec, h->nsec);
  A.pub_shape = h->pub_shape;
  A.sine_on = sine ? 1 : 0; A.sine_period = h->sine_period; A.sine_time0 = h->sine_time; A.sine_pub_dt = h->sine_pub_dt;
  for (int p = 0; p < 8; ++p) A.snap_peers[p] = h->snap_peers[p];
  A.snap_multimem = (h->snap_multimem && (!h->general || h->flex)) ? 1 : 0;
  A.n_snap_peers = h->n_snap_peers; A.snap_stride = h->snap_stride; A.snap_offset = h->snap_offset;
  A.snap_every = h->n_snap_peers > 0 ? h->snap_every : 0; A.snap_written0 = h->snap_written; A.snap_capacity = h->snap_capacity;
}

// every k_step_fast instance of the library, collected once from the translation units that hold them
static const std::vector<FastEntry> &fast_table() {
  static const std::vector<FastEntry> table = [] {
    std::vector<FastEntry> t;
    fast_entries_nc4_base(t); fast_entries_nc4_diag(t); fast_entries_nc4_spec(t);
    fast_entries_nc8_base(t); fast_entries_nc8_diag(t); fast_entries_nc8_spec(t); fast_entries_nc8_pair(t);
    return t;
  }();
  return table;
}
static const FastEntry *fast_find(int nc, int mode, bool dmom, int spec) {
  for (const FastEntry &e : fast_table())
    if (e.nc == nc && e.mode == mode && e.dmom == dmom && e.spec == spec) return &e;
  return nullptr;
}

static int launch_step(cdpr_handle h, const StepArgs &A) {
  if (h->flex) {
    if (h->flexr) flexr_launch(h->L.nc, h->flex_nf, h->flexr_hold, h->flex_lanes, grid_for(h->np * h->flex_lanes, h->flex_tpb), A, h->stream);
    else flex_launch(h->L.nc, h->flex_nf, h->flex_unroll, h->flex_lanes, grid_for(h->np * h->flex_lanes, h->flex_tpb), A, h->stream);
  } else if (h->general) {
    general_launch(std::max(h->pc[PID_VEL].degree, h->pc[PID_POS].degree), (unsigned)(h->np / kTpb), A, h->stream);
  } else {
    const bool dm = h->dmom_ok[A.live_idx] && !h->force_fir && A.mode != MODE_FORCE;
    // the velocity and position modes with the moment D-term are specialised on the robot constants; every
    // other combination runs the diagonal-inertia or fully general instance
    const int spec_full = h->rc.spec, spec_base = h->rc.spec & SPEC_DIAG;
    // one target for all cables (the sine publisher) and no feed-forward term: targets live in a register
    const bool uniform_noff = A.sine_on && !A.cmd_table && A.live.kf == 0.0 && h->targets_uniform;
    // most specialised instance first: with / without the paired-anchor form, with / without the one-target layout
    const FastEntry *e = nullptr;
    for (int spec_try : {spec_full, spec_full & ~SPEC_PAIR}) {
      if (e || !dm) break;
      if (A.mode == MODE_VELOCITY && uniform_noff) e = fast_find(h->L.nc, A.mode, true, spec_try | SPEC_NOFF | SPEC_UTGT);
      if (!e) e = fast_find(h->L.nc, A.mode, true, spec_try);
    }
    if (!e) e = fast_find(h->L.nc, A.mode, dm, spec_base);
    if (!e) return fail(h, CDPR_ERR_UNSUPPORTED, "no step kernel instance for this configuration");
    e->launch(grid_for(h->np, e->tpb), A, h->stream);
  }
  CK(h, cudaGetLastError());
  ++h->launches;
  return CDPR_OK;
}

static void advance_host_clock(cdpr_handle h, long long k, bool sine) {
  if (sine) {  // one publish before every step whose 0-based index m is a multiple of the period, m in [step_count, step_count + k)
    const long long p = h->sine_period, lo = h->step_count, hi = h->step_count + k;
    const long long publishes = (hi + p - 1) / p - (lo + p - 1) / p;
    // repeated addition, like the device (and the driver's `time += 1.0 / 100.0`, sinevelocitytest.cpp:48): bitwise the same
    for (long long e = 0; e < publishes; ++e) h->sine_time = h->sine_time + h->sine_pub_dt;
  }
  if (h->n_snap_peers > 0 && h->snap_every > 0) h->snap_written += (h->step_count + k) / h->snap_every - h->step_count / h->snap_every;
  long long ns = (long long)h->nsec + (long long)h->dt_ns * k;
  h->sec += (int)(ns / 1000000000LL);
  h->nsec = (int)(ns % 1000000000LL);
  h->step_count += k;
}

extern "C" int cdpr_step(cdpr_handle h, int64_t k_steps) {
  if (!h || k_steps < 0) return CDPR_ERR_BAD_ARG;
  if (k_steps == 0) return CDPR_OK;
  cudaSetDevice(h->device);
  main_begin(h);
  if (h->timing) CK(h, cudaEventRecord(h->ev0, h->stream));
  long long remaining = k_steps;
  while (h->flex && remaining > 0) {
    // every instance latches its own commands and switches its own mode inside the kernel (step_flex.cuh): nothing to
    // decide on the host
    const long long seg = std::min<long long>(remaining, 1 << 30);
    StepArgs A;
    fill_args(h, A, (int)seg, h->sine_on);
    int rc = launch_step(h, A);
    if (rc) return rc;
    advance_host_clock(h, seg, h->sine_on);
    remaining -= seg;
  }
  while (remaining > 0) {
    // CdprGazeboPlugin::update, .cpp:206-221: a pending velocity command is fanned out first, then a pending
    // position command; each setter resets its Pid when the mode changes (JointForceCalculator.cpp:99-119).
    // A publish of the in-kernel sine generator IS a velocity command arriving with this step, so it takes part in
    // the velocity fan-out -- a position command pending at a publish step is applied after it and wins until the
    // next publish.
    const bool publish_now = h->sine_on && (h->step_count % h->sine_period == 0);
    if (h->vel_pending || publish_now) {
      if (h->mode != MODE_VELOCITY) { int rc = reset_pid(h, PID_VEL); if (rc) return rc; }
      h->mode = MODE_VELOCITY; h->vel_pending = false;
    }
    if (h->pos_pending) {
      if (h->mode != MODE_POSITION) { int rc = reset_pid(h, PID_POS); if (rc) return rc; }
      h->mode = MODE_POSITION; h->pos_pending = false;
    }
    long long seg = std::min<long long>(remaining, 1 << 30);
    bool sine = false, silent_publish = false;
    if (h->sine_on) {
      if (h->mode == MODE_VELOCITY) sine = true;
      else {  // Force / Position mode runs until the next publish switches back to Velocity
        const long long r = h->step_count % h->sine_period;
        seg = std::min<long long>(seg, h->sine_period - r);
        silent_publish = (r == 0);  // the publisher ticked, but its target was overridden in the same update
      }
    }
    StepArgs A;
    fill_args(h, A, (int)seg, sine);
    int rc = launch_step(h, A);
    if (rc) return rc;
    if (sine) {  // did the publisher write a command in this segment? then every cable holds that one value
      const long long first = ((h->step_count + h->sine_period - 1) / h->sine_period) * h->sine_period;  // first n0' >= step_count with n0' % period == 0
      if (first < h->step_count + seg) h->targets_uniform = true;
    }
    advance_host_clock(h, seg, sine);
    if (silent_publish) h->sine_time = h->sine_time + h->sine_pub_dt;
    remaining -= seg;
  }
  if (h->timing) { CK(h, cudaEventRecord(h->ev1, h->stream)); h->timed = true; }
  main_end(h);
  return CDPR_OK;
}

// One plugin update in one call: see include/cdpr_b200.h.  The step kernel itself publishes (last-step body) into pinned host
// memory mapped into the device, so an update is [command scatter, only when a message arrived] + ONE kernel + one
// stream synchronisation -- no pack kernels, no memcpy calls, no event records.
extern "C" int cdpr_update(cdpr_handle h, const float *vel_axes, const float *pos_axes, double *position, double *velocity, double *effort,
                           double *pose7, double *twist6) {
  if (!h) return CDPR_ERR_BAD_ARG;
  if (h->async_copies) return fail(h, CDPR_ERR_UNSUPPORTED, "cdpr_update is synchronous: leave async mode first");
  cudaSetDevice(h->device);
  const size_t nj = (size_t)h->n * h->L.nc, per = 3 * nj + 13 * (size_t)h->n;
  if (!h->pub_host) {
    if (cudaHostAlloc((void **)&h->pub_host, sizeof(double) * per, cudaHostAllocMapped) != cudaSuccess) return fail(h, CDPR_ERR_NOMEM, "cudaHostAlloc(publish buffer) failed");
    if (cudaHostAlloc((void **)&h->cmd_host[0], sizeof(float) * 2 * nj, cudaHostAllocMapped) != cudaSuccess) return fail(h, CDPR_ERR_NOMEM, "cudaHostAlloc(command buffer) failed");
    h->cmd_host[1] = h->cmd_host[0] + nj;
  }
  // subscriber callbacks: latch the messages (CdprGazeboPlugin.cpp:67-83); the scatter kernel reads the mapped buffer
  for (int which = 0; which < 2; ++which) {
    const float *axes = which == 0 ? vel_axes : pos_axes;
    if (!axes) continue;
    std::memcpy(h->cmd_host[which], axes, sizeof(float) * nj);
    const int field = which == 0 ? CAB_VEL_TARGET : CAB_POS_TARGET;
    if (h->flex) k_scatter_cab_masked<float><<<grid_for(h->n, 256), 256, 0, h->stream>>>(h->L, field, h->cmd_host[which], nullptr, which == 0 ? 4u : 8u, 0);
    else k_scatter_cab<float><<<grid_for(h->n, 256), 256, 0, h->stream>>>(h->L, field, h->cmd_host[which]);
    CK(h, cudaGetLastError());
    ++h->launches;
    if (which == 0) { h->vel_pending = true; h->targets_uniform = false; } else h->pos_pending = true;
  }
  double *o = h->pub_host;
  h->pub_ptr[0] = position ? o : nullptr; h->pub_ptr[1] = velocity ? o + nj : nullptr; h->pub_ptr[2] = effort ? o + 2 * nj : nullptr;
  h->pub_ptr[3] = pose7 ? o + 3 * nj : nullptr; h->pub_ptr[4] = twist6 ? o + 3 * nj + 7 * (size_t)h->n : nullptr;
  const bool timing = h->timing;
  h->timing = false;
  int rc = cdpr_step(h, 1);
  h->timing = timing;
  for (double *&p : h->pub_ptr) p = nullptr;
  if (rc) return rc;
  CK(h, cudaStreamSynchronize(h->stream));
  if (position) std::memcpy(position, o, sizeof(double) * nj);
  if (velocity) std::memcpy(velocity, o + nj, sizeof(double) * nj);
  if (effort) std::memcpy(effort, o + 2 * nj, sizeof(double) * nj);
  if (pose7) std::memcpy(pose7, o + 3 * nj, sizeof(double) * 7 * (size_t)h->n);
  if (twist6) std::memcpy(twist6, o + 3 * nj + 7 * (size_t)h->n, sizeof(double) * 6 * (size_t)h->n);
  return CDPR_OK;
}

extern "C" int64_t cdpr_step_count(cdpr_handle h) { return h ? h->step_count : -1; }
extern "C" double cdpr_sim_time(cdpr_handle h) { return h ? time_double(h->sec, h->nsec) : -1.0; }

// ---------------------------------------------------------------------------------------------
// outputs
// ---------------------------------------------------------------------------------------------
extern "C" int cdpr_get_platform_state(cdpr_handle h, double *pose7, double *twist6) {
  if (!h) return CDPR_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  const size_t nb = sizeof(double) * (size_t)h->n;
  int rc = ensure_stage(h, nb * 13);
  if (rc) return rc;
  double *dp = (double *)h->stage, *dt = dp + 7 * h->n;
  cudaStream_t st = io_begin(h);
  k_pack_platform<<<grid_for(h->n, 256), 256, 0, st>>>(h->L, pose7 ? dp : nullptr, twist6 ? dt : nullptr);
  CK(h, cudaGetLastError());
  if (pose7) CK(h, cudaMemcpyAsync(pose7, dp, nb * 7, cudaMemcpyDeviceToHost, st));
  if (twist6) CK(h, cudaMemcpyAsync(twist6, dt, nb * 6, cudaMemcpyDeviceToHost, st));
  io_end(h);
  CK(h, sync_unless_async(h));
  return CDPR_OK;
}

extern "C" int cdpr_set_platform_state(cdpr_handle h, const double *pose7, const double *twist6) {
  if (!h) return CDPR_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  const size_t nb = sizeof(double) * (size_t)h->n;
  int rc = ensure_stage(h, nb * 13);
  if (rc) return rc;
  double *dp = (double *)h->stage, *dt = dp + 7 * h->n;
  cudaStream_t st = io_begin(h);
  if (pose7) CK(h, cudaMemcpyAsync(dp, pose7, nb * 7, cudaMemcpyHostToDevice, st));
  if (twist6) CK(h, cudaMemcpyAsync(dt, twist6, nb * 6, cudaMemcpyHostToDevice, st));
  k_unpack_platform<<<grid_for(h->n, 256), 256, 0, st>>>(h->L, pose7 ? dp : nullptr, twist6 ? dt : nullptr, 1);
  CK(h, cudaGetLastError());
  io_end(h);
  CK(h, sync_unless_async(h));
  return CDPR_OK;
}

extern "C" int cdpr_get_joint_states(cdpr_handle h, double *position, double *velocity, double *effort) {
  if (!h) return CDPR_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  const size_t nb = sizeof(double) * (size_t)h->n * h->L.nc;
  int rc = ensure_stage(h, nb * 3);
  if (rc) return rc;
  double *d0 = (double *)h->stage, *d1 = d0 + h->n * h->L.nc, *d2 = d1 + h->n * h->L.nc;
  cudaStream_t st = io_begin(h);
  k_joint_states<<<grid_for(h->n, 256), 256, 0, st>>>(h->L, h->rc, position ? d0 : nullptr, velocity ? d1 : nullptr, effort ? d2 : nullptr);
  CK(h, cudaGetLastError());
  if (position) CK(h, cudaMemcpyAsync(position, d0, nb, cudaMemcpyDeviceToHost, st));
  if (velocity) CK(h, cudaMemcpyAsync(velocity, d1, nb, cudaMemcpyDeviceToHost, st));
  if (effort) CK(h, cudaMemcpyAsync(effort, d2, nb, cudaMemcpyDeviceToHost, st));
  io_end(h);
  CK(h, sync_unless_async(h));
  return CDPR_OK;
}

extern "C" int cdpr_get_pid_state(cdpr_handle h, double *out) {
  if (!h || !out) return CDPR_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  const size_t nb = sizeof(double) * (size_t)h->n * h->L.nc * 6;
  int rc = ensure_stage(h, nb);
  if (rc) return rc;
  cudaStream_t st = io_begin(h);
  k_pid_state<<<grid_for(h->n, 256), 256, 0, st>>>(h->L, h->mode, h->flex ? 1 : 0, (double *)h->stage);
  CK(h, cudaGetLastError());
  CK(h, cudaMemcpyAsync(out, h->stage, nb, cudaMemcpyDeviceToHost, st));
  io_end(h);
  CK(h, sync_unless_async(h));
  return CDPR_OK;
}

extern "C" int cdpr_get_pid_terms(cdpr_handle h, double *out) {
  if (!h || !out) return CDPR_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  const size_t nb = sizeof(double) * (size_t)h->n * h->L.nc * 5;
  int rc = ensure_stage(h, nb);
  if (rc) return rc;
  cudaStream_t st = io_begin(h);
  k_pid_terms<<<grid_for(h->n, 256), 256, 0, st>>>(h->L, (double *)h->stage);
  CK(h, cudaGetLastError());
  CK(h, cudaMemcpyAsync(out, h->stage, nb, cudaMemcpyDeviceToHost, st));
  io_end(h);
  CK(h, sync_unless_async(h));
  return CDPR_OK;
}

// ---------------------------------------------------------------------------------------------
// checkpoint / resume: header + the raw device arrays in the order of common.cuh
// ---------------------------------------------------------------------------------------------
struct BlobHeader {
  uint64_t magic;
  int64_t n, np;
  int32_t nc, len, casc, general, mode, vel_pending, pos_pending, sec, nsec, sine_on;
  int64_t step_count;
  double sine_time;
  uint64_t cfg_hash;  // FNV-1a of the cdpr_config the blob was taken under (gains and dt shape the window state)
  uint8_t pad[152];
};
static uint64_t config_hash(const cdpr_config &c) {
  uint64_t hsh = 1469598103934665603ULL;
  const unsigned char *p = (const unsigned char *)&c;
  for (size_t i = 0; i < sizeof(c); ++i) { hsh ^= p[i]; hsh *= 1099511628211ULL; }
  return hsh;
}
static const uint64_t kMagic = 0x3030324252504443ULL;  // "CDPRB200"

struct Section { void *ptr; size_t bytes; };
static std::vector<Section> sections(cdpr_handle h) {
  const DevLayout &L = h->L;
  const size_t col = sizeof(double) * (size_t)L.np;
  std::vector<Section> s = {{L.plat, col * 13}, {L.cab, col * L.nc * CAB_F}, {L.pid, col * L.nc * 2 * PID_F}, {L.win_y, col * L.nc * 2 * L.len}};
  if (L.mom) s.push_back({L.mom, col * L.nc * 2 * 3});
  if (L.win_x) s.push_back({L.win_x, col * L.nc * 2 * L.len});
  if (L.filt) s.push_back({L.filt, col * L.nc * 2 * 2 * L.casc * 4});
  s.push_back({L.ctl, sizeof(uint32_t) * (size_t)L.np * L.nc});
  s.push_back({L.ictl, sizeof(uint32_t) * (size_t)L.np});
  s.push_back({L.sine, col * 3});
  return s;
}

extern "C" size_t cdpr_state_bytes(cdpr_handle h) {
  if (!h) return 0;
  size_t b = sizeof(BlobHeader);
  for (auto &s : sections(h)) b += s.bytes;
  return b;
}

extern "C" int cdpr_get_state(cdpr_handle h, void *blob, size_t bytes) {
  if (!h || !blob || bytes < cdpr_state_bytes(h)) return CDPR_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  main_begin(h);
  BlobHeader hd;
  std::memset(&hd, 0, sizeof(hd));
  hd.magic = kMagic; hd.n = h->n; hd.np = h->np; hd.nc = h->L.nc; hd.len = h->L.len; hd.casc = h->L.casc; hd.general = (int)h->general + (int)h->flex;
  hd.mode = h->mode; hd.vel_pending = h->vel_pending; hd.pos_pending = h->pos_pending; hd.sec = h->sec; hd.nsec = h->nsec;
  hd.sine_on = h->sine_on ? 1 + h->pub_shape : 0; hd.step_count = h->step_count; hd.sine_time = h->sine_time; hd.cfg_hash = config_hash(h->cfg);
  std::memcpy(blob, &hd, sizeof(hd));
  uint8_t *o = (uint8_t *)blob + sizeof(hd);
  for (auto &s : sections(h)) {
    CK(h, cudaMemcpyAsync(o, s.ptr, s.bytes, cudaMemcpyDeviceToHost, h->stream));
    o += s.bytes;
  }
  CK(h, cudaStreamSynchronize(h->stream));
  return CDPR_OK;
}

extern "C" int cdpr_set_state(cdpr_handle h, const void *blob, size_t bytes) {
  if (!h || !blob || bytes < cdpr_state_bytes(h)) return CDPR_ERR_BAD_ARG;
  BlobHeader hd;
  std::memcpy(&hd, blob, sizeof(hd));
  if (hd.magic != kMagic || hd.n != h->n || hd.np != h->np || hd.nc != h->L.nc || hd.len != h->L.len || hd.casc != h->L.casc ||
      hd.general != (int)h->general + (int)h->flex)
    return fail(h, CDPR_ERR_BAD_ARG, "checkpoint does not match this handle's shape");
  if (hd.cfg_hash != config_hash(h->cfg)) return fail(h, CDPR_ERR_BAD_ARG, "checkpoint was taken under a different cdpr_config");
  if (hd.mode < MODE_FORCE || hd.mode > MODE_VELOCITY || hd.step_count < 0) return fail(h, CDPR_ERR_BAD_ARG, "corrupt checkpoint header");
  cudaSetDevice(h->device);
  main_begin(h);
  const uint8_t *o = (const uint8_t *)blob + sizeof(hd);
  for (auto &s : sections(h)) {
    CK(h, cudaMemcpyAsync(s.ptr, o, s.bytes, cudaMemcpyHostToDevice, h->stream));
    o += s.bytes;
  }
  CK(h, cudaStreamSynchronize(h->stream));
  h->targets_uniform = false;
  h->mode = hd.mode; h->vel_pending = hd.vel_pending; h->pos_pending = hd.pos_pending; h->sec = hd.sec; h->nsec = hd.nsec;
  h->sine_on = hd.sine_on != 0; if (hd.sine_on) h->pub_shape = hd.sine_on - 1; h->step_count = hd.step_count; h->sine_time = hd.sine_time;
  return CDPR_OK;
}

// ---------------------------------------------------------------------------------------------
// snapshots
// ---------------------------------------------------------------------------------------------
extern "C" int cdpr_set_snapshot_peers(cdpr_handle h, int64_t every, void *const *peer_bufs, int n_peers, int64_t instance_offset,
                                       int64_t total_instances, int64_t capacity) {
  if (!h || every < 0 || capacity < 0 || n_peers < 0 || n_peers > 8) return CDPR_ERR_BAD_ARG;
  h->n_snap_peers = 0; h->snap_every = 0; h->snap_written = 0; h->snap_capacity = 0; h->snap_multimem = false;
  if (every == 0 || n_peers == 0 || !peer_bufs) return CDPR_OK;
  if (instance_offset < 0 || instance_offset + h->n > total_instances) return fail(h, CDPR_ERR_BAD_ARG, "instance range outside the gather buffer");
  for (int p = 0; p < n_peers; ++p) {
    if (!peer_bufs[p]) return fail(h, CDPR_ERR_BAD_ARG, "null peer buffer");
    h->snap_peers[p] = (double *)peer_bufs[p];
  }
  h->n_snap_peers = n_peers; h->snap_stride = total_instances; h->snap_offset = instance_offset;
  h->snap_every = every; h->snap_capacity = capacity;
  return CDPR_OK;
}

extern "C" int cdpr_set_snapshot_multicast(cdpr_handle h, int64_t every, void *multicast_buf, int64_t instance_offset,
                                           int64_t total_instances, int64_t capacity) {
  if (!h) return CDPR_ERR_BAD_ARG;
  if (h->general && !h->flex) return fail(h, CDPR_ERR_UNSUPPORTED, "multicast snapshots need the fast or flex kernel variant");
  void *one[1] = {multicast_buf};
  int rc = cdpr_set_snapshot_peers(h, multicast_buf ? every : 0, one, multicast_buf ? 1 : 0, instance_offset, total_instances, capacity);
  if (rc == CDPR_OK && multicast_buf && every > 0) h->snap_multimem = true;
  return rc;
}

extern "C" int cdpr_set_snapshots(cdpr_handle h, int64_t every, void *dev_buf, int64_t capacity) {
  if (!h) return CDPR_ERR_BAD_ARG;
  void *one[1] = {dev_buf};
  return cdpr_set_snapshot_peers(h, dev_buf ? every : 0, one, dev_buf ? 1 : 0, 0, h->n, capacity);
}
extern "C" int64_t cdpr_snapshot_count(cdpr_handle h) { return h ? std::min(h->snap_written, h->snap_capacity) : -1; }

// ---------------------------------------------------------------------------------------------
// kinematics only
// ---------------------------------------------------------------------------------------------
static int launch_ik(cdpr_handle h, const IkArgs &A, bool aos) {
  const unsigned grid = grid_for(A.n, 256);
  // device path, even pose count and 16-byte aligned buffers: one thread per (cable, pose pair), double2 accesses
  if (!aos && A.n <= (1 << 19) && (A.n % 2) == 0 && ((uintptr_t)A.state13 % 16) == 0 && ((uintptr_t)A.out % 16) == 0 && (h->L.nc == 4 || h->L.nc == 8)) {
    k_ik_pair<<<dim3(grid_for(A.n / 2, 256), (unsigned)h->L.nc), 256, 0, h->stream>>>(A);
    CK(h, cudaGetLastError());
    ++h->launches;
    return CDPR_OK;
  }
  if (h->L.nc == 4) { if (aos) k_ik<4, true><<<grid, 256, 0, h->stream>>>(A); else k_ik<4, false><<<grid, 256, 0, h->stream>>>(A); }
  else if (h->L.nc == 8) { if (aos) k_ik<8, true><<<grid, 256, 0, h->stream>>>(A); else k_ik<8, false><<<grid, 256, 0, h->stream>>>(A); }
  else return fail(h, CDPR_ERR_UNSUPPORTED, "cdpr_ik supports 4 or 8 cables");
  CK(h, cudaGetLastError());
  ++h->launches;
  return CDPR_OK;
}

extern "C" int cdpr_ik_device(cdpr_handle h, int64_t n, const void *dev_state13, void *dev_out) {
  if (!h || n < 1 || !dev_state13 || !dev_out) return CDPR_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  main_begin(h);
  IkArgs A;
  std::memset(&A, 0, sizeof(A));
  A.rc = h->rc; A.nc = h->L.nc; A.n = n; A.state13 = (const double *)dev_state13; A.out = (double *)dev_out;
  if (h->timing) CK(h, cudaEventRecord(h->ev0, h->stream));
  int rc = launch_ik(h, A, false);
  if (rc) return rc;
  if (h->timing) { CK(h, cudaEventRecord(h->ev1, h->stream)); h->timed = true; }
  return CDPR_OK;
}

extern "C" int cdpr_ik(cdpr_handle h, int64_t n, const double *pose7, const double *twist6, double *length, double *length_rate, double *wmat) {
  if (!h || n < 1 || !pose7 || !twist6 || !length || !length_rate || !wmat) return CDPR_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  const int nc = h->L.nc;
  const size_t in_b = sizeof(double) * (size_t)n * 13, out_b = sizeof(double) * (size_t)n * nc * 8;
  int rc = ensure_stage(h, in_b + out_b);
  if (rc) return rc;
  main_begin(h);
  double *dp = (double *)h->stage, *dt = dp + 7 * n, *dl = dt + 6 * n, *dr = dl + n * nc, *dw = dr + n * nc;
  CK(h, cudaMemcpyAsync(dp, pose7, sizeof(double) * n * 7, cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaMemcpyAsync(dt, twist6, sizeof(double) * n * 6, cudaMemcpyHostToDevice, h->stream));
  IkArgs A;
  std::memset(&A, 0, sizeof(A));
  A.rc = h->rc; A.nc = nc; A.n = n; A.pose7 = dp; A.twist6 = dt; A.length = dl; A.length_rate = dr; A.wmat = dw;
  if (h->timing) CK(h, cudaEventRecord(h->ev0, h->stream));
  rc = launch_ik(h, A, true);
  if (rc) return rc;
  if (h->timing) { CK(h, cudaEventRecord(h->ev1, h->stream)); h->timed = true; }
  CK(h, cudaMemcpyAsync(length, dl, sizeof(double) * n * nc, cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaMemcpyAsync(length_rate, dr, sizeof(double) * n * nc, cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaMemcpyAsync(wmat, dw, sizeof(double) * n * nc * 6, cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  return CDPR_OK;
}

// ---------------------------------------------------------------------------------------------
// sampled rollouts
// ---------------------------------------------------------------------------------------------
extern "C" int cdpr_rollout(cdpr_handle h, int64_t n_robots, int64_t n_seq, const double *pose7, const double *twist6, const float *cmds,
                            int64_t n_cmd, int64_t steps_per_cmd, const double target_pos[3], double lambda, void *dev_cost_seq,
                            double *host_cost) {
  if (!h || !cmds || !target_pos || n_robots < 1 || n_seq < 1 || n_cmd < 1 || steps_per_cmd < 1) return CDPR_ERR_BAD_ARG;
  if (n_robots * n_seq != h->n) return fail(h, CDPR_ERR_BAD_ARG, "n_robots * n_seq must equal the handle's instance count");
  if (n_cmd * steps_per_cmd > (1 << 30)) return fail(h, CDPR_ERR_BAD_ARG, "rollout too long");
  cudaSetDevice(h->device);
  main_begin(h);
  int rc = reset_to_load_state(h, h->stream);
  if (rc) return rc;
  if (pose7 || twist6) {
    const size_t nb = sizeof(double) * (size_t)n_robots;
    if ((rc = ensure_stage(h, nb * 13))) return rc;
    double *dp = (double *)h->stage, *dt = dp + 7 * n_robots;
    if (pose7) CK(h, cudaMemcpyAsync(dp, pose7, nb * 7, cudaMemcpyHostToDevice, h->stream));
    if (twist6) CK(h, cudaMemcpyAsync(dt, twist6, nb * 6, cudaMemcpyHostToDevice, h->stream));
    k_unpack_platform<<<grid_for(h->n, 256), 256, 0, h->stream>>>(h->L, pose7 ? dp : nullptr, twist6 ? dt : nullptr, n_seq);
    CK(h, cudaGetLastError());
  }
  const size_t cmd_bytes = sizeof(float) * (size_t)n_seq * n_cmd * h->L.nc;
  if (cmd_bytes > h->cmd_dev_bytes) {
    if (h->cmd_dev) cudaFree(h->cmd_dev);
    h->cmd_dev = nullptr; h->cmd_dev_bytes = 0;
    if (cudaMalloc((void **)&h->cmd_dev, cmd_bytes) != cudaSuccess) return fail(h, CDPR_ERR_NOMEM, "cudaMalloc(cmds) failed");
    h->cmd_dev_bytes = cmd_bytes;
  }
  if (!h->cost_dev && cudaMalloc((void **)&h->cost_dev, sizeof(double) * (size_t)h->np) != cudaSuccess)
    return fail(h, CDPR_ERR_NOMEM, "cudaMalloc(cost) failed");
  CK(h, cudaMemcpyAsync(h->cmd_dev, cmds, cmd_bytes, cudaMemcpyHostToDevice, h->stream));
  // the first command arrives with step 1: setVelocityTarget from Position mode resets the velocity Pid
  // (already in its reset state) and switches the mode (JointForceCalculator.cpp:111-119)
  h->mode = MODE_VELOCITY;
  StepArgs A;
  fill_args(h, A, (int)(n_cmd * steps_per_cmd), false);
  A.n_snap_peers = 0; A.snap_every = 0;
  A.cmd_table = h->cmd_dev; A.n_seq = (int)n_seq; A.n_cmd = (int)n_cmd; A.steps_per_cmd = (int)steps_per_cmd;
  A.cost = h->cost_dev; A.target[0] = target_pos[0]; A.target[1] = target_pos[1]; A.target[2] = target_pos[2]; A.lambda = lambda;
  if (h->timing) CK(h, cudaEventRecord(h->ev0, h->stream));
  if ((rc = launch_step(h, A))) return rc;
  if (dev_cost_seq) {
    k_reduce_cost_seq<<<grid_for(n_seq, 128), 128, 0, h->stream>>>(h->cost_dev, n_robots, n_seq, (double *)dev_cost_seq);
    CK(h, cudaGetLastError());
    ++h->launches;
  }
  if (h->timing) { CK(h, cudaEventRecord(h->ev1, h->stream)); h->timed = true; }
  h->targets_uniform = false;
  const int peers_saved = h->n_snap_peers;
  h->n_snap_peers = 0;  // rollouts write no snapshots
  advance_host_clock(h, n_cmd * steps_per_cmd, false);
  h->n_snap_peers = peers_saved;
  main_end(h);
  if (host_cost) {
    CK(h, cudaMemcpyAsync(host_cost, h->cost_dev, sizeof(double) * (size_t)h->n, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
  }
  return CDPR_OK;
}

// ---------------------------------------------------------------------------------------------
// raw device access, measurement helpers
// ---------------------------------------------------------------------------------------------
// Host-side constants of the D-term, computable without a device (used by the CPU tests)
extern "C" int cdpr_dterm_weights(const cdpr_pid_params *pid, double dt, double *fir, double *quadratic, int *is_quadratic) {
  if (!pid || !fir || pid->d_buffer_length < 2 || pid->d_buffer_length > CDPR_MAX_DBUF || pid->d_degree < 0 ||
      pid->d_degree > CDPR_MAX_DEGREE || !(dt > 0.0))
    return CDPR_ERR_BAD_ARG;
  double w[kMaxDbuf], abc[3] = {0, 0, 0};
  fir_weights(pid->d_degree, pid->d_buffer_length, dt, w);
  for (int j = 0; j < pid->d_buffer_length; ++j) fir[j] = w[j];
  const bool ok = fir_as_quadratic(w, pid->d_buffer_length, abc);
  if (quadratic) for (int q = 0; q < 3; ++q) quadratic[q] = abc[q];
  if (is_quadratic) *is_quadratic = ok ? 1 : 0;
  return CDPR_OK;
}

extern "C" int64_t cdpr_padded_instances(cdpr_handle h) { return h ? h->np : -1; }
extern "C" void *cdpr_device_platform_state(cdpr_handle h) { return h ? h->L.plat : nullptr; }
extern "C" int64_t cdpr_launch_count(cdpr_handle h) { return h ? h->launches : -1; }
extern "C" const char *cdpr_kernel_detail(cdpr_handle h) {
  if (!h) return "";
  char buf[96];
  if (h->flex && h->flexr) std::snprintf(buf, sizeof(buf), "flexr:lanes=%d,nf=%d,hold=%d", h->flex_lanes, h->flex_nf, h->flexr_hold ? 1 : 0);
  else if (h->flex) std::snprintf(buf, sizeof(buf), "flex:lanes=%d,nf=%d,unroll=%d", h->flex_lanes, h->flex_nf, h->flex_unroll);
  else std::snprintf(buf, sizeof(buf), "%s", h->general ? "general" : "fast");
  h->detail = buf;
  return h->detail.c_str();
}
extern "C" const char *cdpr_kernel_variant(cdpr_handle h) { return !h ? "" : (h->flex ? "flex" : h->general ? "general" : "fast"); }

extern "C" float cdpr_last_kernel_ms(cdpr_handle h) {
  if (!h || !h->timed) return -1.0f;
  cudaSetDevice(h->device);
  if (cudaEventSynchronize(h->ev1) != cudaSuccess) return -1.0f;
  float ms = -1.0f;
  if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) != cudaSuccess) return -1.0f;
  return ms;
}

extern "C" double cdpr_measure_fp64_tflops(int device, int iters) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return -1.0;
  cudaSetDevice(device);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1.0;
  if (iters < 1) iters = 4096;
  const int blocks = prop.multiProcessorCount * 8, tpb = 256;
  double *out = nullptr;
  if (cudaMalloc((void **)&out, sizeof(double) * blocks * tpb) != cudaSuccess) return -1.0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_dfma_peak<<<blocks, tpb>>>(out, iters / 8 + 1, 1.0);  // warm-up
  double best = -1.0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k_dfma_peak<<<blocks, tpb>>>(out, iters, 1.0 + rep);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { best = -1.0; break; }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 64.0 * (double)iters * (double)blocks * tpb;
    best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(out);
  return best;
}
