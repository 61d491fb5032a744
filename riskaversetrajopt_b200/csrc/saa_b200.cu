// libsaa_b200.so -- C ABI implementation (see include/saa_b200.h).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <limits>

#include "saa_common.cuh"
#include "layout.cuh"
#include "drone_kernels.cuh"
#include "drone32_kernels.cuh"
#include "car_kernels.cuh"
#include "car32_kernels.cuh"
#include "hopper_kernels.cuh"
#include "tail_kernels.cuh"
#include "generic_kernels.cuh"
#include "qp_kernels.cuh"

using namespace saa;

namespace {

constexpr int kVersion = 100;
constexpr int kS = 20;            // horizon the kernels are instantiated for (reference: S = 20)
#ifndef SAA_WARPS
#define SAA_WARPS 6
#endif
constexpr int kWarps = SAA_WARPS;    // warps per block of the assemble kernels
constexpr int kBlocksPerSM = SAA_BPS; // resident blocks per SM (persistent grid = SMs x this)
#ifndef SAA_DRONE_TILE_DEFAULT
#define SAA_DRONE_TILE_DEFAULT 16
#endif
#ifndef SAA_CAR_TILE_DEFAULT
#define SAA_CAR_TILE_DEFAULT 16
#endif
#ifndef SAA_WARPS32
#define SAA_WARPS32 7
#endif
constexpr int kWarps32 = SAA_WARPS32; // warps per block of the 32-sample-tile drone kernel (one block per SM)

thread_local std::string g_create_error;

}  // namespace

struct saa_handle {
  int problem = 0, method = 0, variant = 0, S = 0, precision = 64, device = 0;
  i64 M_local = 0, M_global = 0, sample_offset = 0;
  i64 M_cap = 0;                     // capacity (M_local at creation); saa_set_active may lower M_local
  void *qp = nullptr;                // column / pair tables of the device QP (qp_host.cuh)
  bool qp_generic = false;           // force the run-time-horizon QP kernels (SAA_QP_GENERIC=1; tests)
  double *d_select = nullptr; i64 select_len = 0;   // state / histogram / block counts of the radix select
  i64 M_out = 0, first_out = 0;
  double alpha = 0.1;
  bool params_set = false, samples_set = false;
  saa_drone_params dp{};
  saa_car_params cp{};
  double car_ego0[4] = {0, 0, 0, 0};   // ego initial state (identical for all samples)
  // packed samples (device, library-owned)
  i64 Mpad = 0;
  void *d_a = nullptr, *d_b = nullptr, *d_c = nullptr, *d_d = nullptr;
  // hopper
  int n_feat = 0; double mu_nom = 0.0;
  // scratch
  int n_sms = kSMs;
  int reserve_sms = 0;         // SMs the persistent assemble grids leave free (saa_reserve_sms)
  int car_tile = SAA_CAR_TILE_DEFAULT;       // samples per warp tile of the car assemble kernel: 16 | 32 (env SAA_CAR_TILE)
  int drone_tile = SAA_DRONE_TILE_DEFAULT;   // samples per warp tile of the drone assemble kernel: 16 | 32 (env SAA_DRONE_TILE)
  double *d_partials = nullptr; i64 partials_len = 0;
  double *d_sums = nullptr;
  double *d_means_scratch = nullptr; i64 means_scratch_len = 0;   // saa_linearize_means (may run on a side stream)
  i64 *d_fin_off = nullptr;          // [0..255]: normal pattern, [256..511]: car relaxed pattern
  void *d_relax_scratch = nullptr; i64 relax_scratch_bytes = 0;   // car scp_iter 0: sample 0 alone
  unsigned long long *d_nonfinite = nullptr;                      // samples with a non-finite rollout (saa_check_finite)
  double *d_hopper_geo = nullptr;                                 // hopper Hessian: [6][32] contact geometry
  Layout lay;                 // destination geometry (M_out)
  mutable std::string err;
};

namespace {

int fail(const saa_handle *h, int code, const std::string &msg) {
  if (h) h->err = msg; else g_create_error = msg;
  return code;
}
#define SAA_CUDA(h, call)                                                              \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess)                                                             \
      return fail(h, SAA_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

// precision 32 = FP32 STORAGE of the outputs; the packed samples and the arithmetic stay FP64
constexpr size_t kInSize = sizeof(double);

int relax_threshold(const saa_handle *h) { return h->problem == SAA_DRONE ? 2 : 1; }

// relaxation constants of get_constraints_coeffs (drone_risk.py:413-417, drone_times.py:421-425)
void drone_relax(const saa_handle *h, double *scale, double *bound) {
  if (h->variant == SAA_VARIANT_TIMES) { *scale = 1e-5; *bound = 10.0; }
  else { *scale = 1e-7; *bound = 0.1; }
}
// multiplier / padding of the sample rows (drone_risk.py:310, :324-325, :352; drone_times.py:324-334)
void drone_mult(const saa_handle *h, double *mult, double *pad) {
  if (h->method == SAA_METHOD_BASELINE) {
    if (h->variant == SAA_VARIANT_TIMES) { *mult = 1.0; *pad = 0.0; }
    else { *mult = 0.01; *pad = 1e-3; }
  } else { *mult = 0.01; *pad = 0.0; }
}

int ensure_scratch(saa_handle *h, i64 partial_doubles) {
  if (h->partials_len < partial_doubles) {
    if (h->d_partials) cudaFree(h->d_partials);
    h->d_partials = nullptr; h->partials_len = 0;
    SAA_CUDA(h, cudaMalloc(&h->d_partials, partial_doubles * sizeof(double)));
    h->partials_len = partial_doubles;
  }
  if (!h->d_sums) SAA_CUDA(h, cudaMalloc(&h->d_sums, 1024 * sizeof(double)));
  if (!h->d_nonfinite) {
    SAA_CUDA(h, cudaMalloc(&h->d_nonfinite, sizeof(unsigned long long)));
    SAA_CUDA(h, cudaMemset(h->d_nonfinite, 0, sizeof(unsigned long long)));
  }
  return SAA_OK;
}

// offsets (in the destination Ax) of the sample-mean entries, in the order of
// the kernels' reduction slots
int fin_offsets_for(saa_handle *h, const Layout &L, std::vector<i64> &off) {
  const int S = h->S;
  int rows[4];
  auto pos = [&](int c, int row) -> i64 {
    const int n = L.fin_rows(c, rows);
    for (int r = 0; r < n; ++r) if (rows[r] == row) return L.ucol[c] + r;
    return -1;
  };
  off.clear();
  if (h->problem == SAA_DRONE) {
    for (int a = 0; a < 3; ++a) for (int j = 0; j < S - 1; ++j) off.push_back(pos(j * 3 + a, a));
    for (int a = 0; a < 3; ++a) for (int j = 0; j < S; ++j) off.push_back(pos(j * 3 + a, 3 + a));
  } else {
    // car slots (CarRed): PX (c<2, j<S-1), PY (c<2, j<S-1), V (j<S), PHI (j<S)
    for (int c = 0; c < 2; ++c) for (int j = 0; j < S - 1; ++j) off.push_back(pos(j * 2 + c, 0));
    for (int c = 0; c < 2; ++c) for (int j = 0; j < S - 1; ++j) off.push_back(pos(j * 2 + c, 1));
    for (int j = 0; j < S; ++j) off.push_back(pos(j * 2 + 0, 2));
    for (int j = 0; j < S; ++j) off.push_back(pos(j * 2 + 1, 3));
  }
  for (i64 o : off) if (o < 0) return fail(h, SAA_ERR_STATE, "internal: final-row offset");
  return SAA_OK;
}

int upload_fin_offsets(saa_handle *h) {
  std::vector<i64> off, all(512, 0);
  int rc = fin_offsets_for(h, h->lay, off);
  if (rc) return rc;
  std::copy(off.begin(), off.end(), all.begin());
  if (h->problem == SAA_CAR) {
    Layout Lr; Lr.build(h->problem, h->method, h->S, h->M_out, true);
    rc = fin_offsets_for(h, Lr, off);
    if (rc) return rc;
    std::copy(off.begin(), off.end(), all.begin() + 256);
  }
  if (!h->d_fin_off) SAA_CUDA(h, cudaMalloc(&h->d_fin_off, 512 * sizeof(i64)));
  SAA_CUDA(h, cudaMemcpy(h->d_fin_off, all.data(), all.size() * sizeof(i64), cudaMemcpyHostToDevice));
  return SAA_OK;
}

int set_geometry(saa_handle *h, i64 M_out, i64 first_out) {
  if (M_out < h->M_local || first_out < 0 || first_out + h->M_local > M_out)
    return fail(h, SAA_ERR_ARG, "output geometry does not contain the local samples");
  const bool same = M_out == h->M_out && h->d_fin_off != nullptr;      // the final-row offsets depend on M_out only
  h->M_out = M_out; h->first_out = first_out;
  h->lay.build(h->problem, h->method, h->S, M_out, false);
  if (h->problem != SAA_HOPPER && !same) return upload_fin_offsets(h);
  return SAA_OK;
}

// ---------------------------------------------------------------------------
// constants kernel (generic over drone / car)
// ---------------------------------------------------------------------------
struct ConstArgs {
  i64 M_local, first_out, M_out;
  int R, nu, n_fin;
  int saa;                 // method == SAA
  int write_shared;
  i64 ycol0, ycol_len, slackcol, tcol;
  i64 row_cvar, row_y0, row_s0, row_slack, row_ctrl0;
  double cvar_y, cvar_slack, cvar_t, y_diag, y_slack, y_rows, t_rows, slack_last;
  double l_rows, u_cvar, u_y, u_slack;      // bounds of the risk rows
  int write_u_rows; double u_rows;          // relaxed: sample-row upper bounds are constant
  double u_max;
  i64 ucol_last[128];                       // position of the control-identity entry per u column
};

template <typename T>
__global__ void write_constants_kernel(const __grid_constant__ ConstArgs C, T *Ax, T *l, T *u) {
  const i64 tid = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  const i64 nth = (i64)gridDim.x * blockDim.x;
  const i64 n_samp_rows = C.M_local * C.R;
  // per-sample-row constants
  for (i64 e = tid; e < n_samp_rows; e += nth) {
    const i64 grow = C.first_out * C.R + e;           // sample-row index in the destination
    l[C.row_s0 + grow] = (T)C.l_rows;
    if (C.write_u_rows) u[C.row_s0 + grow] = (T)C.u_rows;
    if (C.saa) {
      Ax[C.tcol + 1 + grow] = (T)C.t_rows;
      const i64 i = e / C.R, r = e - i * C.R;
      Ax[C.ycol0 + (C.first_out + i) * C.ycol_len + 2 + r] = (T)C.y_rows;
    }
  }
  if (C.saa) {
    for (i64 i = tid; i < C.M_local; i += nth) {
      const i64 gi = C.first_out + i;
      Ax[C.ycol0 + gi * C.ycol_len] = (T)C.cvar_y;
      Ax[C.ycol0 + gi * C.ycol_len + 1] = (T)C.y_diag;
      Ax[C.slackcol + 1 + gi] = (T)C.y_slack;
      l[C.row_y0 + gi] = (T)C.l_rows;
      u[C.row_y0 + gi] = (T)C.u_y;
    }
  }
  if (C.write_shared && tid < C.nu) {
    Ax[C.ucol_last[tid]] = (T)1.0;
    l[C.row_ctrl0 + tid] = (T)(-C.u_max);
    u[C.row_ctrl0 + tid] = (T)C.u_max;
  }
  if (C.write_shared && C.saa && tid == 0) {
    Ax[C.slackcol] = (T)C.cvar_slack;
    Ax[C.slackcol + 1 + C.M_out] = (T)C.slack_last;
    Ax[C.tcol] = (T)C.cvar_t;
    l[C.row_cvar] = (T)C.l_rows; u[C.row_cvar] = (T)C.u_cvar;
    l[C.row_slack] = (T)C.l_rows; u[C.row_slack] = (T)C.u_slack;
  }
}

template <typename T>
int launch_constants(saa_handle *h, int scp_iter, int write_shared, void *Ax, void *l, void *u,
                     cudaStream_t st) {
  const Layout &L = h->lay;
  if (L.nu > 128) return fail(h, SAA_ERR_ARG, "n_u*S > 128 unsupported");
  ConstArgs C{};
  C.M_local = h->M_local; C.first_out = h->first_out; C.M_out = h->M_out;
  C.R = L.R; C.nu = L.nu; C.n_fin = L.n_fin;
  C.saa = h->method == SAA_METHOD_SAA;
  C.write_shared = write_shared;
  C.ycol0 = L.ycol0; C.ycol_len = L.ycol_len(); C.slackcol = L.slackcol; C.tcol = L.tcol;
  C.row_cvar = L.row_cvar; C.row_y0 = L.row_y0; C.row_s0 = L.row_s0; C.row_slack = L.row_slack;
  C.row_ctrl0 = L.row_ctrl0;
  const bool relaxed = scp_iter < relax_threshold(h);
  const double inf = std::numeric_limits<double>::infinity();
  double mult = 1.0, pad = 0.0, scale = 1.0, bound = 0.0;
  if (h->problem == SAA_DRONE) { drone_mult(h, &mult, &pad); drone_relax(h, &scale, &bound); }
  if (h->problem == SAA_CAR && relaxed)
    return fail(h, SAA_ERR_ARG, "car scp_iter == 0 uses the relaxed pattern: call saa_car_relaxed");
  const double rs = relaxed ? scale : 1.0;
  C.cvar_y = 1.0 * rs; C.cvar_slack = 1.0 * rs;               // slack gets 1.0: fill slice quirk (:337)
  C.cvar_t = (double)h->M_global * h->alpha * rs;
  C.y_diag = -1.0 * rs; C.y_slack = -1.0 * rs; C.slack_last = -1.0 * rs;
  C.y_rows = -mult * rs; C.t_rows = -mult * rs;
  C.l_rows = relaxed ? -bound : -inf;
  C.u_cvar = C.u_y = C.u_slack = relaxed ? bound : 0.0;
  C.write_u_rows = relaxed; C.u_rows = bound;
  C.u_max = h->problem == SAA_DRONE ? h->dp.u_max : h->cp.u_max;
  for (int c = 0; c < L.nu; ++c) C.ucol_last[c] = L.ucol[c + 1] - 1;
  const i64 work = std::max<i64>(h->M_local * L.R, 64);
  const int threads = 256;
  const int blocks = (int)std::min<i64>((work + threads - 1) / threads, (i64)h->n_sms * 8);
  write_constants_kernel<T><<<blocks, threads, 0, st>>>(C, (T *)Ax, (T *)l, (T *)u);
  SAA_CUDA(h, cudaGetLastError());
  return SAA_OK;
}

// ---------------------------------------------------------------------------
// drone launchers
// ---------------------------------------------------------------------------
template <typename T>
void fill_drone_common(const saa_handle *h, const double *us, T *us_out, T &dt, T &noise_c, T &drag,
                       T &kp, T &kd, T (&x0)[6], T (&oc)[3][2]) {
  for (int i = 0; i < kS * 3; ++i) us_out[i] = (T)us[i];
  dt = (T)h->dp.dt;
  noise_c = (T)(std::sqrt(h->dp.dt) * h->dp.beta);
  drag = (T)h->dp.drag_coefficient; kp = (T)h->dp.gain_p; kd = (T)h->dp.gain_v;
  for (int i = 0; i < 6; ++i) x0[i] = (T)h->dp.x_init[i];
  for (int o = 0; o < 3; ++o) for (int a = 0; a < 2; ++a) oc[o][a] = (T)h->dp.obs_positions[o][a];
}

int grid_for(const saa_handle *h, i64 ntiles, int warps, int blocks_per_sm) {
  const i64 want = (ntiles + warps - 1) / warps;
  return (int)std::max<i64>(1, std::min<i64>(want, (i64)std::max(1, h->n_sms - h->reserve_sms) * blocks_per_sm));
}

// same formula as DroneChain<S,J>::CA/CB (drone_kernels.cuh)
i64 drone_col_start(int j, int a, int S, i64 M) {
  return (9 * j + 3 * a + 2) + M * (i64)(6 * j * (S - 1) - 3 * j * (j - 1) + 3 * a * (S - 1 - j));
}

// mode: DRONE_FULL (Ax), DRONE_FACTOR (fsp/fp record instead of Ax), DRONE_EXPAND (record -> Ax for
// samples [s_begin, s_begin + count) of the output geometry)
template <typename TO, int MODE>
int launch_drone_assemble(saa_handle *h, const double *us, int scp_iter, void *Ax, void *u, void *Z,
                          void *fsp, void *fp, i64 s_begin, i64 count, double *sums, cudaStream_t st) {
  using T = double;
  using Args = DroneArgs<T, TO, kS>;
  using Smem = DroneSmem<TO, kS, kWarps>;
  Args A{};
  A.mass = (const T *)h->d_a; A.dw = (const T *)h->d_b; A.q = (const T *)h->d_c;
  A.M = MODE == DRONE_EXPAND ? count : h->M_local; A.Mpad = h->Mpad;
  if (us) fill_drone_common<T>(h, us, A.us, A.dt, A.noise_c, A.drag, A.kp, A.kd, A.x0, A.oc);
  else for (int o = 0; o < 3; ++o) for (int a = 0; a < 2; ++a) A.oc[o][a] = (T)h->dp.obs_positions[o][a];
  for (int i = 0; i < 6; ++i) A.xf[i] = (T)h->dp.x_final[i];
  double mult, pad, scale, bound;
  drone_mult(h, &mult, &pad); drone_relax(h, &scale, &bound);
  const bool relaxed = scp_iter < 2;
  A.escale = (T)(relaxed ? mult * scale : mult);
  A.ubscale = (T)mult; A.ubpad = (T)pad; A.ztol = (T)0;
  A.Ax = (TO *)Ax; A.fsp = (TO *)fsp; A.fp = (TO *)fp; A.s_begin = s_begin;
  const Layout &L = h->lay;
  A.M_out = h->M_out; A.first_out = h->first_out;
  // the kernel derives the column positions in closed form; they must agree with the layout
  for (int a = 0; a < 2; ++a)
    for (int j = 0; j < kS - 1; ++j)
      if (L.run_start(j * 3 + a) != drone_col_start(j, a, kS, h->M_out))
        return fail(h, SAA_ERR_STATE, "internal: closed-form column offsets disagree with the layout");
  A.ub = relaxed ? nullptr : (TO *)u;         // relaxed: bounds are the constant +-bound
  A.ub_off = L.row_s0 + h->first_out * L.R;
  A.Z = (TO *)Z;
  const bool tile32 = MODE == DRONE_FULL && h->drone_tile == 32;
  const i64 ntiles = tile32 ? (A.M + kTile32 - 1) / kTile32 : (A.M + kTileSamples - 1) / kTileSamples;
  const int grid = tile32 ? grid_for(h, ntiles, kWarps32, 1) : grid_for(h, ntiles, kWarps, kBlocksPerSM);
  // rows [0, grid) of the partial sums come from the assemble kernel, rows [grid, grid + gridz)
  // from the z-axis kernel
  constexpr int kZWarps = 4;
  const int gridz = MODE == DRONE_EXPAND ? 0
      : (int)std::max<i64>(1, std::min<i64>((h->M_local + kZWarps * 32 - 1) / (kZWarps * 32), (i64)h->n_sms * 2));
  int rc = ensure_scratch(h, (i64)(grid + gridz) * DroneRed<kS>::N);
  if (rc) return rc;
  A.partials = h->d_partials;
  if (tile32) {
    using Smem32 = Drone32Smem<TO, kS, kWarps32>;
    static_assert(sizeof(Smem32) <= 232448, "drone32: shared memory over the per-block limit");
    auto kern32 = drone_assemble32_kernel<T, TO, kS, kWarps32>;
    SAA_CUDA(h, cudaFuncSetAttribute(kern32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem32)));
    kern32<<<grid, kWarps32 * 32, sizeof(Smem32), st>>>(A);
  } else {
    auto kern = drone_assemble_kernel<T, TO, kS, kWarps, MODE>;
    SAA_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
    kern<<<grid, kWarps * 32, sizeof(Smem), st>>>(A);
  }
  SAA_CUDA(h, cudaGetLastError());
  if (MODE == DRONE_EXPAND) return SAA_OK;
  drone_axis_mean_kernel<T, TO, kS, kZWarps><<<gridz, kZWarps * 32, 0, st>>>(A, 2, h->d_partials + (i64)grid * DroneRed<kS>::N);
  SAA_CUDA(h, cudaGetLastError());
  const int n = DroneRed<kS>::N;
  reduce_partials_kernel<double><<<(n + 3) / 4, 128, 0, st>>>(h->d_partials, grid + gridz, n, sums);
  SAA_CUDA(h, cudaGetLastError());
  return SAA_OK;
}

template <typename TO>
int launch_drone_rollout(saa_handle *h, const double *us, void *Xs, void *Z, double t_risk,
                         double sat_tol, double ztol, double *out3, cudaStream_t st) {
  constexpr int W = 4;
  using T = double;
  using Args = DroneRollArgs<T, TO, kS>;
  Args A{};
  A.mass = (const T *)h->d_a; A.dw = (const T *)h->d_b; A.q = (const T *)h->d_c;
  A.M = h->M_local; A.Mpad = h->Mpad;
  fill_drone_common<T>(h, us, A.us, A.dt, A.noise_c, A.drag, A.kp, A.kd, A.x0, A.oc);
  A.Xs = (TO *)Xs; A.Z = (TO *)Z;
  A.ztol = (T)ztol; A.t_risk = (T)t_risk; A.sat_tol = (T)sat_tol;
  const i64 ntiles = (h->M_local + 31) / 32;
  const int grid = grid_for(h, ntiles, W, 4);
  int rc = ensure_scratch(h, (i64)grid * 3);
  if (rc) return rc;
  A.partials = out3 ? h->d_partials : nullptr;
  const size_t smem = Xs ? (size_t)W * 32 * (((kS + 1) * 6) | 1) * sizeof(TO) : 0;
  auto kern = drone_rollout_kernel<T, TO, kS, W>;
  SAA_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(W * 32 * (((kS + 1) * 6) | 1) * sizeof(TO))));
  kern<<<grid, W * 32, smem, st>>>(A);
  SAA_CUDA(h, cudaGetLastError());
  if (out3) {
    reduce_cvar_kernel<<<1, 32, 0, st>>>(h->d_partials, grid, out3);
    SAA_CUDA(h, cudaGetLastError());
  }
  return SAA_OK;
}

template <typename T>
int launch_finalize(saa_handle *h, const double *sums, int relaxed_pattern, void *Ax, void *l, void *u,
                    cudaStream_t st) {
  const int n_entries = (int)saa_mean_len(h) - h->lay.n_fin;
  const int n = n_entries + h->lay.n_fin;
  const i64 *fin_off = h->d_fin_off + ((relaxed_pattern && h->problem == SAA_CAR) ? 256 : 0);
  scatter_means_kernel<T><<<(n + 127) / 128, 128, 0, st>>>(sums, 1.0 / (double)h->M_global, n_entries,
                                                         fin_off, h->lay.n_fin, (T *)Ax, (T *)l,
                                                         (T *)u);
  SAA_CUDA(h, cudaGetLastError());
  return SAA_OK;
}

// ---------------------------------------------------------------------------
// Fused all-reduce + finalize over NVLink peer memory (multi-GPU).  Every rank owns an inbox
//   struct { double data[2][W][128]; unsigned long long flag[W]; }
// mapped by all the others (CUDA IPC).  One 128-thread block per rank: store my sums into slot
// [epoch & 1][rank] of EVERY inbox (peer stores), fence, raise flag[rank] = epoch in every inbox, spin
// until my own inbox holds the epoch of every rank, then add the W contributions in rank order and
// scatter the means.  One launch, a few microseconds of NVLink latency, no NCCL kernel -- and because
// every rank adds the same numbers in the same order, the expectation rows are bitwise identical on
// all ranks and independent of arrival order.  The double buffer makes epoch e+1 writes safe while a
// slow rank still reads epoch e (a rank cannot run two epochs ahead: it needs the slow rank's flag).
// ---------------------------------------------------------------------------
constexpr int kPeerSlots = 128, kPeerMaxWorld = 16;
struct PeerInbox { double data[2][kPeerMaxWorld][kPeerSlots]; unsigned long long flag[kPeerMaxWorld]; };
struct PeerArgs { PeerInbox *inbox[kPeerMaxWorld]; int rank, world; unsigned long long epoch; };

template <typename T>
__global__ void __launch_bounds__(kPeerSlots)
peer_allreduce_finalize_kernel(const __grid_constant__ PeerArgs P, const double *__restrict__ sums, double inv_M,
                               int n_entries, const i64 *__restrict__ fin_off, int n_fin, T *Ax, T *l, T *u) {
  const int r = threadIdx.x, n = n_entries + n_fin, par = (int)(P.epoch & 1);
  if (r < n) {
    const double v = sums[r];
    for (int q = 0; q < P.world; ++q) {
      volatile double *dst = &P.inbox[q]->data[par][P.rank][r];
      *dst = v;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (r < P.world) {
    volatile unsigned long long *f = &P.inbox[r]->flag[P.rank];
    *f = P.epoch;
    volatile unsigned long long *mine = &P.inbox[P.rank]->flag[r];
    while (*mine < P.epoch) { }
  }
  __threadfence_system();
  __syncthreads();
  if (r < n) {
    double acc = 0.0;
    for (int p = 0; p < P.world; ++p) acc += ((volatile double *)P.inbox[P.rank]->data[par][p])[r];
    const T v = (T)(acc * inv_M);
    if (r < n_entries) Ax[fin_off[r]] = v;
    else { l[r - n_entries] = v; u[r - n_entries] = v; }
  }
}

// ---------------------------------------------------------------------------
// merge a gathered compact shard into the destination matrix (multi-GPU, NCCL path)
// ---------------------------------------------------------------------------
struct MergeArgs {
  int ncols;
  i64 src_off[64], dst_off[64], count[64];   // u columns that carry sample rows, then the bounds
};

template <typename T>
__global__ void merge_shard_kernel(const __grid_constant__ MergeArgs G, const T *__restrict__ sAx,
                                   const T *__restrict__ su, T *__restrict__ Ax, T *__restrict__ u) {
  // blockIdx.y = run, blockIdx.x strides over the run
  const int r = blockIdx.y;
  const bool bounds = (r == G.ncols - 1);
  const T *src = (bounds ? su : sAx) + G.src_off[r];
  T *dst = (bounds ? u : Ax) + G.dst_off[r];
  const i64 n = G.count[r];
  for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (i64)gridDim.x * blockDim.x)
    dst[e] = src[e];
}

template <typename T>
int launch_merge(saa_handle *h, const void *sAx, const void *su, i64 M_shard, i64 first, void *Ax,
                 void *u, cudaStream_t st) {
  const Layout &L = h->lay;
  Layout Ls; Ls.build(h->problem, h->method, h->S, M_shard, false);
  MergeArgs G{};
  int n = 0;
  for (int c = 0; c < L.nu; ++c) {
    const int len = L.run_len(c);
    if (len == 0) continue;
    if (n >= 63) return fail(h, SAA_ERR_ARG, "too many columns");
    G.src_off[n] = Ls.run_start(c);
    G.dst_off[n] = L.run_start(c) + first * len;
    G.count[n] = M_shard * len;
    ++n;
  }
  G.src_off[n] = Ls.row_s0; G.dst_off[n] = L.row_s0 + first * L.R; G.count[n] = M_shard * L.R;
  G.ncols = ++n;
  const int threads = 256;
  const int bx = (int)std::max<i64>(1, std::min<i64>((M_shard * 57 + threads - 1) / threads, (i64)h->n_sms * 2));
  merge_shard_kernel<T><<<dim3(bx, n), threads, 0, st>>>(G, (const T *)sAx, (const T *)su, (T *)Ax, (T *)u);
  SAA_CUDA(h, cudaGetLastError());
  return SAA_OK;
}

}  // namespace

int64_t saa_mean_len(const saa_handle *h);

#include "car_host.cuh"
#include "hopper_host.cuh"
#include "tail_host.cuh"
#include "generic_host.cuh"
#include "qp_host.cuh"

// =============================================================================
// C ABI
// =============================================================================
extern "C" {

int saa_version(void) { return kVersion; }

const char *saa_last_error(const saa_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int saa_create(saa_handle **out, int problem, int method, int variant, int64_t M_local,
               int64_t M_global, int64_t sample_offset, int S, double alpha, int precision,
               int device) {
  if (!out) return fail(nullptr, SAA_ERR_ARG, "out is NULL");
  *out = nullptr;
  if (problem < SAA_DRONE || problem > SAA_HOPPER) return fail(nullptr, SAA_ERR_ARG, "unknown problem");
  if (method != SAA_METHOD_SAA && method != SAA_METHOD_BASELINE) return fail(nullptr, SAA_ERR_ARG, "unknown method");
  if (precision != 64 && precision != 32) return fail(nullptr, SAA_ERR_ARG, "precision must be 64 or 32");
  if (M_local <= 0 || M_global < M_local || sample_offset < 0 || sample_offset + M_local > M_global)
    return fail(nullptr, SAA_ERR_ARG, "need 0 < M_local <= M_global and the local range inside the global one");
  if (problem != SAA_HOPPER && (S < 3 || S > kGenSMax))
    return fail(nullptr, SAA_ERR_ARG, "need 3 <= S <= 32 (S = 20, the reference horizon, runs the tuned kernels; "
                                      "other horizons the generic ones)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(nullptr, SAA_ERR_NO_DEVICE, "no CUDA device: libsaa_b200 has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(nullptr, SAA_ERR_ARG, "bad device index");
  if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, SAA_ERR_CUDA, "cudaSetDevice failed");
  saa_handle *h = new (std::nothrow) saa_handle();
  if (!h) return fail(nullptr, SAA_ERR_ARG, "out of host memory");
  h->problem = problem; h->method = method; h->variant = variant; h->S = S;
  h->precision = precision; h->device = device;
  h->qp_generic = std::getenv("SAA_QP_GENERIC") != nullptr;
  h->M_local = M_local; h->M_cap = M_local; h->M_global = M_global; h->sample_offset = sample_offset; h->alpha = alpha;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) h->n_sms = prop.multiProcessorCount;
  if (const char *e = std::getenv("SAA_DRONE_TILE")) h->drone_tile = std::atoi(e) == 32 ? 32 : 16;
  if (const char *e = std::getenv("SAA_CAR_TILE")) h->car_tile = std::atoi(e) == 32 ? 32 : 16;
  if (problem != SAA_HOPPER) {
    int rc = set_geometry(h, M_global, sample_offset);
    if (rc) { g_create_error = h->err; delete h; return rc; }
  }
  *out = h;
  return SAA_OK;
}

int saa_destroy(saa_handle *h) {
  if (!h) return SAA_OK;
  cudaSetDevice(h->device);
  cudaFree(h->d_a); cudaFree(h->d_b); cudaFree(h->d_c); cudaFree(h->d_d);
  cudaFree(h->d_partials); cudaFree(h->d_sums); cudaFree(h->d_fin_off); cudaFree(h->d_relax_scratch);
  cudaFree(h->d_nonfinite); cudaFree(h->d_hopper_geo); cudaFree(h->d_means_scratch); cudaFree(h->d_select);
  qp_free(h);
  delete h;
  return SAA_OK;
}

int saa_set_params_drone(saa_handle *h, const saa_drone_params *p) {
  if (!h || !p) return fail(h, SAA_ERR_ARG, "NULL argument");
  if (h->problem != SAA_DRONE) return fail(h, SAA_ERR_ARG, "handle is not a drone problem");
  if (p->n_obs != 3) return fail(h, SAA_ERR_ARG, "kernels are instantiated for n_obs = 3");
  if (!(p->dt > 0)) return fail(h, SAA_ERR_ARG, "dt must be positive");
  h->dp = *p; h->params_set = true;
  return SAA_OK;
}

int saa_set_samples_drone(saa_handle *h, const double *masses, const double *DWs, const double *obs_Qs,
                          void *stream) {
  if (!h || !masses || !DWs || !obs_Qs) return fail(h, SAA_ERR_ARG, "NULL argument");
  if (h->problem != SAA_DRONE) return fail(h, SAA_ERR_ARG, "handle is not a drone problem");
  SAA_CUDA(h, cudaSetDevice(h->device));
  const i64 M = h->M_local;
  h->Mpad = (M + kTileSamples - 1) / kTileSamples * kTileSamples;
  const size_t es = kInSize;
  if (!h->d_a) {
    void *a = nullptr, *b = nullptr, *c = nullptr;       // commit to the handle only if all succeed
    cudaError_t e = cudaMalloc(&a, h->Mpad * es);
    if (e == cudaSuccess) e = cudaMalloc(&b, h->Mpad * es * 3 * h->S);
    if (e == cudaSuccess) e = cudaMalloc(&c, h->Mpad * es * 6);
    if (e != cudaSuccess) {
      cudaFree(a); cudaFree(b); cudaFree(c);
      return fail(h, SAA_ERR_CUDA, std::string("cudaMalloc (packed samples): ") + cudaGetErrorString(e));
    }
    h->d_a = a; h->d_b = b; h->d_c = c;
  }
  const int threads = 128;
  const int blocks = (int)((h->Mpad + threads - 1) / threads);
  cudaStream_t st = (cudaStream_t)stream;
  drone_pack_kernel<double><<<blocks, threads, 0, st>>>(masses, DWs, obs_Qs, M, h->Mpad, h->S,
                                                      (double *)h->d_a, (double *)h->d_b, (double *)h->d_c);
  SAA_CUDA(h, cudaGetLastError());
  h->samples_set = true;
  return SAA_OK;
}

int saa_pattern_sizes(const saa_handle *h, int relaxed_pattern, int64_t *n_rows, int64_t *n_cols,
                      int64_t *nnz) {
  if (!h) return fail(h, SAA_ERR_ARG, "NULL handle");
  if (h->problem == SAA_HOPPER) return fail(h, SAA_ERR_ARG, "hopper has no QP pattern");
  Layout L; L.build(h->problem, h->method, h->S, h->M_out, relaxed_pattern != 0);
  if (n_rows) *n_rows = L.n_rows;
  if (n_cols) *n_cols = L.n_cols;
  if (nnz) *nnz = L.nnz;
  return SAA_OK;
}

int saa_pattern_i32(const saa_handle *h, int relaxed_pattern, int32_t *indptr, int32_t *indices) {
  if (!h || !indptr || !indices) return fail(h, SAA_ERR_ARG, "NULL argument");
  if (h->problem == SAA_HOPPER) return fail(h, SAA_ERR_ARG, "hopper has no QP pattern");
  Layout L; L.build(h->problem, h->method, h->S, h->M_out, relaxed_pattern != 0);
  if (L.nnz > 2147483647LL || L.n_rows > 2147483647LL)
    return fail(h, SAA_ERR_ARG, "pattern needs 64-bit indices: use saa_pattern_i64");
  L.fill<int32_t>(indptr, indices);
  return SAA_OK;
}

int saa_pattern_i64(const saa_handle *h, int relaxed_pattern, int64_t *indptr, int64_t *indices) {
  if (!h || !indptr) return fail(h, SAA_ERR_ARG, "NULL argument");
  if (h->problem == SAA_HOPPER) return fail(h, SAA_ERR_ARG, "hopper has no QP pattern");
  Layout L; L.build(h->problem, h->method, h->S, h->M_out, relaxed_pattern != 0);
  L.fill<int64_t>(indptr, indices);
  return SAA_OK;
}

static int static_args_ok(int problem, int method, int S, int64_t M) {
  if (problem != SAA_DRONE && problem != SAA_CAR) return fail(nullptr, SAA_ERR_ARG, "pattern: drone or car only");
  if (method != SAA_METHOD_SAA && method != SAA_METHOD_BASELINE) return fail(nullptr, SAA_ERR_ARG, "unknown method");
  if (S < 2 || M < 1) return fail(nullptr, SAA_ERR_ARG, "need S >= 2 and M >= 1");
  return SAA_OK;
}

int saa_static_pattern_sizes(int problem, int method, int S, int64_t M, int relaxed_pattern,
                             int64_t *n_rows, int64_t *n_cols, int64_t *nnz) {
  if (int rc = static_args_ok(problem, method, S, M)) return rc;
  Layout L; L.build(problem, method, S, M, relaxed_pattern != 0);
  if (n_rows) *n_rows = L.n_rows;
  if (n_cols) *n_cols = L.n_cols;
  if (nnz) *nnz = L.nnz;
  return SAA_OK;
}

int saa_static_pattern_i32(int problem, int method, int S, int64_t M, int relaxed_pattern,
                           int32_t *indptr, int32_t *indices) {
  if (int rc = static_args_ok(problem, method, S, M)) return rc;
  if (!indptr || !indices) return fail(nullptr, SAA_ERR_ARG, "NULL argument");
  Layout L; L.build(problem, method, S, M, relaxed_pattern != 0);
  if (L.nnz > 2147483647LL || L.n_rows > 2147483647LL)
    return fail(nullptr, SAA_ERR_ARG, "pattern needs 64-bit indices");
  L.fill<int32_t>(indptr, indices);
  return SAA_OK;
}

int saa_static_pattern_i64(int problem, int method, int S, int64_t M, int relaxed_pattern,
                           int64_t *indptr, int64_t *indices) {
  if (int rc = static_args_ok(problem, method, S, M)) return rc;
  if (!indptr) return fail(nullptr, SAA_ERR_ARG, "NULL argument");
  Layout L; L.build(problem, method, S, M, relaxed_pattern != 0);
  L.fill<int64_t>(indptr, indices);
  return SAA_OK;
}

int saa_set_output_geometry(saa_handle *h, int64_t M_out, int64_t first_out) {
  if (!h) return fail(h, SAA_ERR_ARG, "NULL handle");
  if (h->problem == SAA_HOPPER) return fail(h, SAA_ERR_ARG, "hopper has no QP geometry");
  SAA_CUDA(h, cudaSetDevice(h->device));
  return set_geometry(h, M_out, first_out);
}

int saa_set_active(saa_handle *h, int64_t M_active, int64_t M_out, int64_t first_out) {
  if (!h) return fail(h, SAA_ERR_ARG, "NULL handle");
  if (h->problem == SAA_HOPPER) return fail(h, SAA_ERR_ARG, "not available for the hopper");
  if (M_active < 0 || M_active > h->M_cap) return fail(h, SAA_ERR_ARG, "need 0 <= M_active <= the handle's capacity");
  SAA_CUDA(h, cudaSetDevice(h->device));
  h->M_local = M_active;
  return set_geometry(h, M_out, first_out);
}

int saa_write_constants(saa_handle *h, int scp_iter, int write_shared, void *Ax, void *l, void *u,
                        void *stream) {
  if (!h || !Ax || !l || !u) return fail(h, SAA_ERR_ARG, "NULL argument");
  if (h->problem == SAA_HOPPER) return fail(h, SAA_ERR_ARG, "hopper has no QP constants");
  if (!h->params_set) return fail(h, SAA_ERR_STATE, "set params first");
  if (h->M_local == 0 && !write_shared) return SAA_OK;
  SAA_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (h->problem == SAA_CAR && scp_iter < 1)
    return car_write_constants_relaxed(h, write_shared, Ax, l, u, st);
  return h->precision == 64 ? launch_constants<double>(h, scp_iter, write_shared, Ax, l, u, st)
                            : launch_constants<float>(h, scp_iter, write_shared, Ax, l, u, st);
}

int64_t saa_mean_len(const saa_handle *h) {
  if (!h) return 0;
  if (h->problem == SAA_DRONE) return GenDroneRed{h->S}.n();     // == DroneRed<S>::N
  if (h->problem == SAA_CAR) return GenCarRed{h->S}.n();         // == CarRed<S>::N
  return 0;
}

int saa_finalize_means(saa_handle *h, const double *mean_sums, int scp_iter, void *Ax, void *l, void *u,
                       void *stream) {
  if (!h || !mean_sums || !Ax || !l || !u) return fail(h, SAA_ERR_ARG, "NULL argument");
  if (h->problem == SAA_HOPPER) return fail(h, SAA_ERR_ARG, "hopper has no mean rows");
  SAA_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int relaxed_pattern = h->problem == SAA_CAR && scp_iter < 1;
  return h->precision == 64 ? launch_finalize<double>(h, mean_sums, relaxed_pattern, Ax, l, u, st)
                            : launch_finalize<float>(h, mean_sums, relaxed_pattern, Ax, l, u, st);
}

int saa_linearize_assemble(saa_handle *h, const double *us, int scp_iter, void *Ax, void *l, void *u,
                           void *Z, double *mean_sums, int finalize, void *stream) {
  if (!h || !us || !Ax || !u || !l) return fail(h, SAA_ERR_ARG, "NULL argument");
  if (h->problem == SAA_HOPPER) return fail(h, SAA_ERR_ARG, "use saa_hopper_friction for the hopper");
  if (!h->params_set || !h->samples_set) return fail(h, SAA_ERR_STATE, "set params and samples first");
  SAA_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  int rc = ensure_scratch(h, 1);
  if (rc) return rc;
  double *sums = mean_sums ? mean_sums : h->d_sums;
  if (h->M_local == 0) {                                  // saa_set_active(0): nothing to assemble on this rank
    SAA_CUDA(h, cudaMemsetAsync(sums, 0, saa_mean_len(h) * sizeof(double), st));
    return SAA_OK;
  }
  if (h->S != kS) {
    // horizon-generic kernels (generic_kernels.cuh)
    if (h->problem == SAA_DRONE)
      rc = h->precision == 64 ? launch_drone_generic<double>(h, us, scp_iter, Ax, u, Z, sums, st)
                              : launch_drone_generic<float>(h, us, scp_iter, Ax, u, Z, sums, st);
    else
      rc = h->precision == 64 ? launch_car_generic<double>(h, us, scp_iter, Ax, u, Z, sums, st)
                              : launch_car_generic<float>(h, us, scp_iter, Ax, u, Z, sums, st);
  } else if (h->problem == SAA_DRONE)
    rc = h->precision == 64
             ? launch_drone_assemble<double, DRONE_FULL>(h, us, scp_iter, Ax, u, Z, nullptr, nullptr, 0, 0, sums, st)
             : launch_drone_assemble<float, DRONE_FULL>(h, us, scp_iter, Ax, u, Z, nullptr, nullptr, 0, 0, sums, st);
  else
    rc = h->precision == 64 ? launch_car_assemble<double>(h, us, scp_iter, Ax, u, Z, sums, st)
                            : launch_car_assemble<float>(h, us, scp_iter, Ax, u, Z, sums, st);
  if (rc) return rc;
  if (finalize) return saa_finalize_means(h, sums, scp_iter, Ax, l, u, stream);
  return SAA_OK;
}

int64_t saa_peer_inbox_bytes(void) { return (int64_t)sizeof(PeerInbox); }

int saa_peer_allreduce_finalize(saa_handle *h, const double *sums, int scp_iter, void *Ax, void *l, void *u,
                                int rank, int world, void *const *inboxes, uint64_t epoch, void *stream) {
  if (!h || !sums || !Ax || !l || !u || !inboxes) return fail(h, SAA_ERR_ARG, "NULL argument");
  if (h->problem == SAA_HOPPER) return fail(h, SAA_ERR_ARG, "hopper has no mean rows");
  if (world < 1 || world > kPeerMaxWorld || rank < 0 || rank >= world) return fail(h, SAA_ERR_ARG, "bad rank / world (<= 16 ranks)");
  if (epoch == 0) return fail(h, SAA_ERR_ARG, "epochs start at 1");
  SAA_CUDA(h, cudaSetDevice(h->device));
  PeerArgs P{};
  for (int q = 0; q < world; ++q) {
    if (!inboxes[q]) return fail(h, SAA_ERR_ARG, "NULL inbox pointer");
    P.inbox[q] = (PeerInbox *)inboxes[q];
  }
  P.rank = rank; P.world = world; P.epoch = epoch;
  const int n_entries = (int)saa_mean_len(h) - h->lay.n_fin, n = n_entries + h->lay.n_fin;
  if (n > kPeerSlots) return fail(h, SAA_ERR_STATE, "internal: mean block larger than the inbox slots");
  const bool relaxed_pattern = h->problem == SAA_CAR && scp_iter < 1;
  const i64 *fin_off = h->d_fin_off + (relaxed_pattern ? 256 : 0);
  cudaStream_t st = (cudaStream_t)stream;
  const double inv_M = 1.0 / (double)h->M_global;
  if (h->precision == 64)
    peer_allreduce_finalize_kernel<double><<<1, kPeerSlots, 0, st>>>(P, sums, inv_M, n_entries, fin_off, h->lay.n_fin,
                                                                    (double *)Ax, (double *)l, (double *)u);
  else
    peer_allreduce_finalize_kernel<float><<<1, kPeerSlots, 0, st>>>(P, sums, inv_M, n_entries, fin_off, h->lay.n_fin,
                                                                   (float *)Ax, (float *)l, (float *)u);
  SAA_CUDA(h, cudaGetLastError());
  return SAA_OK;
}

int saa_merge_shard(saa_handle *h, const void *shard_Ax, const void *shard_u, int64_t M_shard,
                    int64_t first, void *Ax, void *u, void *stream) {
  if (!h || !shard_Ax || !shard_u || !Ax || !u) return fail(h, SAA_ERR_ARG, "NULL argument");
  if (h->problem == SAA_HOPPER) return fail(h, SAA_ERR_ARG, "hopper has no QP matrix");
  if (M_shard < 1 || first < 0 || first + M_shard > h->M_out)
    return fail(h, SAA_ERR_ARG, "shard does not fit the destination geometry");
  SAA_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  return h->precision == 64 ? launch_merge<double>(h, shard_Ax, shard_u, M_shard, first, Ax, u, st)
                            : launch_merge<float>(h, shard_Ax, shard_u, M_shard, first, Ax, u, st);
}

int saa_shared_alloc(int device, int64_t bytes, void **ptr_out, unsigned char handle_out[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (!ptr_out || !handle_out || bytes <= 0) return fail(nullptr, SAA_ERR_ARG, "bad argument");
  if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, SAA_ERR_CUDA, "cudaSetDevice failed");
  void *p = nullptr;
  cudaError_t e = cudaMalloc(&p, (size_t)bytes);
  if (e != cudaSuccess) return fail(nullptr, SAA_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  if (bytes <= (1 << 20)) cudaMemset(p, 0, (size_t)bytes);     // small control blocks (peer inboxes) start zeroed
  cudaIpcMemHandle_t hd;
  e = cudaIpcGetMemHandle(&hd, p);
  if (e != cudaSuccess) { cudaFree(p); return fail(nullptr, SAA_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e)); }
  std::memcpy(handle_out, &hd, 64);
  *ptr_out = p;
  return SAA_OK;
}

int saa_shared_open(int device, const unsigned char handle[64], void **ptr_out) {
  if (!ptr_out || !handle) return fail(nullptr, SAA_ERR_ARG, "bad argument");
  if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, SAA_ERR_CUDA, "cudaSetDevice failed");
  cudaIpcMemHandle_t hd;
  std::memcpy(&hd, handle, 64);
  void *p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return fail(nullptr, SAA_ERR_CUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
  *ptr_out = p;
  return SAA_OK;
}

int saa_shared_close(int device, void *ptr) {
  if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, SAA_ERR_CUDA, "cudaSetDevice failed");
  return cudaIpcCloseMemHandle(ptr) == cudaSuccess ? SAA_OK : fail(nullptr, SAA_ERR_CUDA, "cudaIpcCloseMemHandle failed");
}

int saa_shared_free(int device, void *ptr) {
  if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, SAA_ERR_CUDA, "cudaSetDevice failed");
  return cudaFree(ptr) == cudaSuccess ? SAA_OK : fail(nullptr, SAA_ERR_CUDA, "cudaFree failed");
}

int saa_factored_sizes(const saa_handle *h, int64_t *n_sp, int64_t *n_p) {
  if (!h) return fail(h, SAA_ERR_ARG, "NULL handle");
  if (h->problem != SAA_DRONE || h->S != kS) return fail(h, SAA_ERR_ARG, "the factored record is implemented for the drone at S = 20");
  if (n_sp) *n_sp = h->M_out * (i64)(kS * (kS - 1));              // 2 axes x S(S-1)/2 sensitivities
  if (n_p) *n_p = h->M_out * (i64)(2 * DroneFac<kS>::ROWS);
  return SAA_OK;
}

int saa_linearize_factored(saa_handle *h, const double *us, int scp_iter, void *fsp, void *fp, void *u,
                           void *Z, double *mean_sums, void *stream) {
  if (!h || !us || !fsp || !fp || !u) return fail(h, SAA_ERR_ARG, "NULL argument");
  if (h->problem != SAA_DRONE || h->S != kS) return fail(h, SAA_ERR_ARG, "the factored record is implemented for the drone at S = 20");
  if (!h->params_set || !h->samples_set) return fail(h, SAA_ERR_STATE, "set params and samples first");
  SAA_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  int rc = ensure_scratch(h, 1);
  if (rc) return rc;
  double *sums = mean_sums ? mean_sums : h->d_sums;
  if (h->M_local == 0) {
    SAA_CUDA(h, cudaMemsetAsync(sums, 0, saa_mean_len(h) * sizeof(double), st));
    return SAA_OK;
  }
  return h->precision == 64
             ? launch_drone_assemble<double, DRONE_FACTOR>(h, us, scp_iter, nullptr, u, Z, fsp, fp, 0, 0, sums, st)
             : launch_drone_assemble<float, DRONE_FACTOR>(h, us, scp_iter, nullptr, u, Z, fsp, fp, 0, 0, sums, st);
}

int saa_expand_factored(saa_handle *h, int scp_iter, const void *fsp, const void *fp, int64_t sample_begin,
                        int64_t sample_count, void *Ax, void *stream) {
  if (!h || !fsp || !fp || !Ax) return fail(h, SAA_ERR_ARG, "NULL argument");
  if (h->problem != SAA_DRONE || h->S != kS) return fail(h, SAA_ERR_ARG, "the factored record is implemented for the drone at S = 20");
  if (!h->params_set) return fail(h, SAA_ERR_STATE, "set params first");
  if (sample_begin < 0 || sample_count < 0 || sample_begin + sample_count > h->M_out)
    return fail(h, SAA_ERR_ARG, "sample range outside the output geometry");
  if (sample_count == 0) return SAA_OK;
  SAA_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  return h->precision == 64
             ? launch_drone_assemble<double, DRONE_EXPAND>(h, nullptr, scp_iter, Ax, nullptr, nullptr, (void *)fsp,
                                                           (void *)fp, sample_begin, sample_count, nullptr, st)
             : launch_drone_assemble<float, DRONE_EXPAND>(h, nullptr, scp_iter, Ax, nullptr, nullptr, (void *)fsp,
                                                          (void *)fp, sample_begin, sample_count, nullptr, st);
}

int saa_rollout(saa_handle *h, const double *us, void *Xs, void *stream) {
  if (!h || !us || !Xs) return fail(h, SAA_ERR_ARG, "NULL argument");
  if (h->problem == SAA_HOPPER) return fail(h, SAA_ERR_ARG, "the hopper trajectory is a decision variable");
  if (!h->params_set || !h->samples_set) return fail(h, SAA_ERR_STATE, "set params and samples first");
  SAA_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (h->S != kS) {
    if (h->problem == SAA_DRONE)
      return h->precision == 64 ? launch_drone_generic_rollout<double>(h, us, Xs, nullptr, 0, 0, 0, nullptr, st)
                                : launch_drone_generic_rollout<float>(h, us, Xs, nullptr, 0, 0, 0, nullptr, st);
    return h->precision == 64 ? launch_car_generic_rollout<double>(h, us, Xs, nullptr, 0, 0, 0, nullptr, st)
                              : launch_car_generic_rollout<float>(h, us, Xs, nullptr, 0, 0, 0, nullptr, st);
  }
  if (h->problem == SAA_DRONE)
    return h->precision == 64 ? launch_drone_rollout<double>(h, us, Xs, nullptr, 0, 0, 0, nullptr, st)
                              : launch_drone_rollout<float>(h, us, Xs, nullptr, 0, 0, 0, nullptr, st);
  return h->precision == 64 ? launch_car_rollout<double>(h, us, Xs, nullptr, 0, 0, 0, nullptr, st)
                            : launch_car_rollout<float>(h, us, Xs, nullptr, 0, 0, 0, nullptr, st);
}

// Measured FP64 FMA throughput of the device: the denominator of the FP64 rooflines (car, hopper).
// 8 independent dependent-FMA chains per thread, 32 warps per SM, best of 5 launches.
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters, double a, double b) {
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 12345.678) out[0] = s;     // never true: keeps the chains alive
}

int saa_measure_fp64_peak(int device, double *fma_per_s) {
  if (!fma_per_s) return fail(nullptr, SAA_ERR_ARG, "NULL argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(nullptr, SAA_ERR_NO_DEVICE, "no CUDA device");
  if (device < 0 || device >= ndev || cudaSetDevice(device) != cudaSuccess) return fail(nullptr, SAA_ERR_ARG, "bad device index");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(nullptr, SAA_ERR_CUDA, "cudaGetDeviceProperties failed");
  double *out = nullptr;
  if (cudaMalloc(&out, 8) != cudaSuccess) return fail(nullptr, SAA_ERR_CUDA, "cudaMalloc failed");
  const int iters = 2048, blocks = prop.multiProcessorCount * 4;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0);
    fp64_peak_kernel<<<blocks, 256>>>(out, iters, 0.9999999, 1e-7);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(out); return fail(nullptr, SAA_ERR_CUDA, "fp64 peak kernel failed"); }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
  *fma_per_s = (double)blocks * 256 * iters * 64 / ((double)best * 1e-3);
  return SAA_OK;
}

int saa_reserve_sms(saa_handle *h, int n_sms) {
  if (!h) return fail(h, SAA_ERR_ARG, "NULL handle");
  if (n_sms < 0 || n_sms >= h->n_sms) return fail(h, SAA_ERR_ARG, "need 0 <= n_sms < number of SMs");
  h->reserve_sms = n_sms;
  return SAA_OK;
}

int saa_check_finite(saa_handle *h, int64_t *count_out, void *stream) {
  if (!h) return fail(h, SAA_ERR_ARG, "NULL handle");
  SAA_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  int rc = ensure_scratch(h, 1);
  if (rc) return rc;
  unsigned long long n = 0;
  SAA_CUDA(h, cudaMemcpyAsync(&n, h->d_nonfinite, sizeof(n), cudaMemcpyDeviceToHost, st));
  SAA_CUDA(h, cudaMemsetAsync(h->d_nonfinite, 0, sizeof(n), st));
  SAA_CUDA(h, cudaStreamSynchronize(st));
  if (count_out) *count_out = (int64_t)n;
  if (n) return fail(h, SAA_ERR_NONFINITE, std::to_string(n) + " sample(s) met a zero / non-finite ego-pedestrian distance "
                                           "(car/driving.py:154 divides by it): their rows are NaN");
  return SAA_OK;
}

int saa_cvar_terms(saa_handle *h, const double *us, double t_risk, double sat_tol, void *Z,
                   double *out3, void *stream) {
  if (!h || !us || !out3) return fail(h, SAA_ERR_ARG, "NULL argument");
  if (h->problem == SAA_HOPPER) return fail(h, SAA_ERR_ARG, "use saa_hopper_cvar_terms for the hopper");
  if (!h->params_set || !h->samples_set) return fail(h, SAA_ERR_STATE, "set params and samples first");
  SAA_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (h->S != kS) {
    const double tol = h->problem == SAA_DRONE ? h->dp.osqp_tol : h->cp.osqp_tol;
    if (h->problem == SAA_DRONE)
      return h->precision == 64 ? launch_drone_generic_rollout<double>(h, us, nullptr, Z, t_risk, sat_tol, tol, out3, st)
                                : launch_drone_generic_rollout<float>(h, us, nullptr, Z, t_risk, sat_tol, tol, out3, st);
    return h->precision == 64 ? launch_car_generic_rollout<double>(h, us, nullptr, Z, t_risk, sat_tol, tol, out3, st)
                              : launch_car_generic_rollout<float>(h, us, nullptr, Z, t_risk, sat_tol, tol, out3, st);
  }
  if (h->problem == SAA_DRONE) {
    const double tol = h->dp.osqp_tol;
    return h->precision == 64 ? launch_drone_rollout<double>(h, us, nullptr, Z, t_risk, sat_tol, tol, out3, st)
                              : launch_drone_rollout<float>(h, us, nullptr, Z, t_risk, sat_tol, tol, out3, st);
  }
  const double tol = h->cp.osqp_tol;
  return h->precision == 64 ? launch_car_rollout<double>(h, us, nullptr, Z, t_risk, sat_tol, tol, out3, st)
                            : launch_car_rollout<float>(h, us, nullptr, Z, t_risk, sat_tol, tol, out3, st);
}

}  // extern "C"
