// Host-side launchers of the car kernels (included by saa_b200.cu).
namespace {

constexpr int kCarWarps = SAA_CAR_WARPS;   // one block per SM; shared memory = warps x CarPass::SIZE

// same formula as CarCol<S,J>::CA/CB (car_kernels.cuh)
i64 car_col_start(int j, int c, int S, i64 M) {
  return (8 * j + 4 * c + 3) + M * (i64)(2 * j * (S - 1) - j * (j - 1) + c * (S - 1 - j));
}

template <typename T, typename Args>
void fill_car_common(const saa_handle *h, const double *us, Args &A) {
  for (int i = 0; i < kS * 2; ++i) A.us[i] = (T)us[i];
  for (int i = 0; i < 4; ++i) A.ego0[i] = (T)h->car_ego0[i];
  A.dt = (T)h->cp.dt;
  A.noise_c = (T)(std::sqrt(h->cp.dt) * h->cp.beta);   // sqrt(dt) * sigma (car/driving.py:183, :200)
  A.v_des = (T)h->cp.speed_ped_des;
  A.d_min = (T)h->cp.min_separation_distance;
  A.x0 = (const T *)h->d_a; A.om = (const T *)h->d_b; A.dw = (const T *)h->d_c;
  A.M = h->M_local; A.Mpad = h->Mpad;
}

template <typename TO>
int launch_car_relaxed_rows(saa_handle *h, const double *us, void *Ax, void *u, cudaStream_t st);

// scp_iter < 1: rows >= n_x are multiplied by exactly 0 (car/driving.py:411-415) and vanish
// from the pattern; only the sample-independent final rows carry values.
template <typename TO>
int launch_car_assemble(saa_handle *h, const double *us, int scp_iter, void *Ax, void *u, void *Z,
                        double *sums, cudaStream_t st) {
  using T = double;
  using Args = CarArgs<T, TO, kS>;
  using Smem = CarSmem<T, TO, kS, kCarWarps>;
  static_assert(sizeof(Smem) <= 232448, "car kernel: shared memory over the 227 KB per-block limit (SAA_CAR_WARPS / SAA_CAR_CAP)");
  const bool relaxed = scp_iter < 1;
  Args A{};
  fill_car_common<T>(h, us, A);
  for (int i = 0; i < 4; ++i) A.goal[i] = (T)h->cp.goal[i];
  A.ztol = (T)0;
  const Layout &L = h->lay;
  A.M_out = h->M_out; A.first_out = h->first_out;
  for (int c = 0; c < 2; ++c)
    for (int j = 0; j < kS - 1; ++j)
      if (L.run_start(j * 2 + c) != car_col_start(j, c, kS, h->M_out))
        return fail(h, SAA_ERR_STATE, "internal: closed-form column offsets disagree with the layout");
  A.Ax = relaxed ? nullptr : (TO *)Ax;
  A.ub = relaxed ? nullptr : (TO *)u;
  A.ub_off = L.row_s0 + h->first_out * L.R;
  A.Z = relaxed ? nullptr : (TO *)Z;
  A.sums = sums;
  A.nonfinite = h->d_nonfinite;
  if (!relaxed && h->car_tile == 32) {
    // 32-sample tiles, lane = sample with both controls (car32_kernels.cuh)
    using Smem32 = CarSmem32<T, TO, kS, SAA_CAR32_WARPS>;
    static_assert(sizeof(Smem32) <= 232448, "car32: shared memory over the 227 KB per-block limit");
    const i64 nt32 = (h->M_local + kCarTile32 - 1) / kCarTile32;
    auto kern32 = car_assemble32_kernel<T, TO, kS, SAA_CAR32_WARPS>;
    SAA_CUDA(h, cudaFuncSetAttribute(kern32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem32)));
    kern32<<<grid_for(h, nt32, SAA_CAR32_WARPS, 1), SAA_CAR32_WARPS * 32, sizeof(Smem32), st>>>(A);
    SAA_CUDA(h, cudaGetLastError());
    return SAA_OK;
  }
  const i64 ntiles = (h->M_local + kTileSamples - 1) / kTileSamples;
  const int grid = relaxed ? 1 : grid_for(h, ntiles, kCarWarps, 1);
  auto kern = car_assemble_kernel<T, TO, kS, kCarWarps>;
  SAA_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
  kern<<<grid, kCarWarps * 32, sizeof(Smem), st>>>(A);
  SAA_CUDA(h, cudaGetLastError());
  if (relaxed && Ax != nullptr) return launch_car_relaxed_rows<TO>(h, us, Ax, u, st);
  return SAA_OK;
}

template <typename TO>
int launch_car_rollout(saa_handle *h, const double *us, void *Xs, void *Z, double t_risk,
                       double sat_tol, double ztol, double *out3, cudaStream_t st) {
  constexpr int W = 4;
  using T = double;
  using Args = CarRollArgs<T, TO, kS>;
  Args A{};
  fill_car_common<T>(h, us, A);
  A.Xs = (TO *)Xs; A.Z = (TO *)Z;
  A.ztol = (T)ztol; A.t_risk = (T)t_risk; A.sat_tol = (T)sat_tol;
  const i64 ntiles = (h->M_local + 31) / 32;
  const int grid = grid_for(h, ntiles, W, 4);
  int rc = ensure_scratch(h, (i64)grid * 3);
  if (rc) return rc;
  A.partials = out3 ? h->d_partials : nullptr;
  const size_t full = (size_t)W * 32 * (((kS + 1) * 8) | 1) * sizeof(TO);
  auto kern = car_rollout_kernel<T, TO, kS, W>;
  SAA_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)full));
  kern<<<grid, W * 32, Xs ? full : 0, st>>>(A);
  SAA_CUDA(h, cudaGetLastError());
  if (out3) {
    reduce_cvar_kernel<<<1, 32, 0, st>>>(h->d_partials, grid, out3);
    SAA_CUDA(h, cudaGetLastError());
  }
  return SAA_OK;
}

// ---- constants of the relaxed (scp_iter == 0) car problem ---------------------------
// Rows >= n_x = 8 of the constraint block are multiplied by exactly 0 (car/driving.py:411-415):
// their entries vanish, l = -inf * 0 = nan, u = 0.  Rows < n_x survive untouched (Layout::keep_y,
// keep_s): the CVaR row, the first "-y_i" rows and possibly the first separation rows of sample 0.
struct CarRelaxArgs {
  i64 M_local, first_out, M_out;
  int nu, keep_y, keep_s, saa, n_x, R;
  i64 ycol0, slackcol, tcol, row_cvar, row_y0, row_s0, row_ctrl0;
  double cvar_t, u_max;
  i64 ucol_last[128];
};

template <typename T>
__global__ void car_relaxed_constants_kernel(const __grid_constant__ CarRelaxArgs C, int write_shared,
                                             T *Ax, T *l, T *u) {
  const i64 tid = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  const i64 nth = (i64)gridDim.x * blockDim.x;
  const T inf = (T)INFINITY;
  const T nan = inf * (T)0;                                   // -inf * 0 (car/driving.py:414)
  if (C.saa) {
    // y columns: [CVaR row] [-y_i row if i < keep_y] [sample 0's surviving rows if i == 0]
    for (i64 i = tid; i < C.M_local; i += nth) {
      const i64 gi = C.first_out + i;
      const i64 pos = C.ycol0 + gi + (gi < C.keep_y ? gi : C.keep_y) + (gi > 0 ? C.keep_s : 0);
      Ax[pos] = (T)1.0;
      i64 q = pos + 1;
      if (gi < C.keep_y) Ax[q++] = (T)-1.0;
      if (gi == 0) for (int e = 0; e < C.keep_s; ++e) Ax[q++] = (T)-1.0;
      const i64 r = C.row_y0 + gi;
      l[r] = (r < C.n_x) ? -inf : nan;
      u[r] = (T)0;
    }
  }
  // this rank's sample rows (the upper bounds of surviving rows are written per iteration)
  const i64 r_lo = C.row_s0 + C.first_out * C.R, r_hi = r_lo + C.M_local * C.R;
  for (i64 r = r_lo + tid; r < r_hi; r += nth) { l[r] = (r < C.n_x) ? -inf : nan; u[r] = (T)0; }
  if (!write_shared) return;
  if (tid < C.nu) {
    Ax[C.ucol_last[tid]] = (T)1.0;
    l[C.row_ctrl0 + tid] = (T)(-C.u_max);
    u[C.row_ctrl0 + tid] = (T)C.u_max;
  }
  if (tid == 0 && C.saa) {
    Ax[C.slackcol] = (T)1.0;                                   // slack in the CVaR row: fill-slice quirk (:341)
    for (int i = 0; i < C.keep_y; ++i) Ax[C.slackcol + 1 + i] = (T)-1.0;
    Ax[C.tcol] = (T)C.cvar_t;
    for (int e = 0; e < C.keep_s; ++e) Ax[C.tcol + 1 + e] = (T)-1.0;
    l[C.row_cvar] = -inf; u[C.row_cvar] = (T)0;
    l[C.row_ctrl0 - 1] = nan; u[C.row_ctrl0 - 1] = (T)0;      // "-slack <= 0" row, zeroed too
  }
}

int car_write_constants_relaxed(saa_handle *h, int write_shared, void *Ax, void *l, void *u,
                                cudaStream_t st) {
  Layout L; L.build(SAA_CAR, h->method, h->S, h->M_out, true);
  CarRelaxArgs C{};
  C.M_local = h->M_local; C.first_out = h->first_out; C.M_out = h->M_out;
  C.nu = L.nu; C.keep_y = L.keep_y; C.keep_s = L.keep_s; C.saa = h->method == SAA_METHOD_SAA;
  C.n_x = L.n_x; C.R = L.R;
  C.ycol0 = L.ycol0; C.slackcol = L.slackcol; C.tcol = L.tcol;
  C.row_cvar = L.row_cvar; C.row_y0 = L.row_y0; C.row_s0 = L.row_s0; C.row_ctrl0 = L.row_ctrl0;
  C.cvar_t = (double)h->M_global * h->alpha;
  C.u_max = h->cp.u_max;
  for (int c = 0; c < L.nu; ++c) C.ucol_last[c] = L.ucol[c + 1] - 1;
  const int threads = 256;
  const int blocks = (int)std::min<i64>((h->M_local * L.R + threads - 1) / threads, (i64)h->n_sms * 8);
  if (h->precision == 64)
    car_relaxed_constants_kernel<double><<<std::max(blocks, 1), threads, 0, st>>>(C, write_shared, (double *)Ax, (double *)l, (double *)u);
  else
    car_relaxed_constants_kernel<float><<<std::max(blocks, 1), threads, 0, st>>>(C, write_shared, (float *)Ax, (float *)l, (float *)u);
  SAA_CUDA(h, cudaGetLastError());
  return SAA_OK;
}

// ---- surviving separation rows of sample 0 at scp_iter == 0 (baseline, or saa with M < 3) ----
// The ordinary kernel linearises sample 0 alone into a scratch matrix with the normal
// one-sample pattern; the (<= 12 + 4) surviving values are then moved to their places in the
// relaxed pattern.
struct CarPickArgs { int n_ax, n_u; i64 src_ax[16], dst_ax[16], src_u[4], dst_u[4]; };

template <typename T>
__global__ void car_relaxed_pick_kernel(const __grid_constant__ CarPickArgs P, const T *__restrict__ sAx,
                                        const T *__restrict__ su, T *Ax, T *u) {
  const int t = threadIdx.x;
  if (t < P.n_ax) Ax[P.dst_ax[t]] = sAx[P.src_ax[t]];
  if (t < P.n_u) u[P.dst_u[t]] = su[P.src_u[t]];
}

template <typename TO>
int launch_car_relaxed_rows(saa_handle *h, const double *us, void *Ax, void *u, cudaStream_t st) {
  using T = double;
  using Args = CarArgs<T, TO, kS>;
  using Smem = CarSmem<T, TO, kS, kCarWarps>;
  Layout Lr; Lr.build(SAA_CAR, h->method, h->S, h->M_out, true);
  if (Lr.keep_s == 0 || h->first_out != 0) return SAA_OK;      // the owner of sample 0 writes them
  Layout L1; L1.build(SAA_CAR, h->method, h->S, 1, false);
  const i64 need = (L1.nnz + L1.n_rows) * (i64)sizeof(TO);
  if (h->relax_scratch_bytes < need) {
    if (h->d_relax_scratch) cudaFree(h->d_relax_scratch);
    h->d_relax_scratch = nullptr; h->relax_scratch_bytes = 0;
    SAA_CUDA(h, cudaMalloc(&h->d_relax_scratch, need));
    h->relax_scratch_bytes = need;
  }
  TO *sAx = (TO *)h->d_relax_scratch, *su = sAx + L1.nnz;
  Args A{};
  fill_car_common<T>(h, us, A);
  for (int i = 0; i < 4; ++i) A.goal[i] = (T)h->cp.goal[i];
  A.M = 1; A.M_out = 1; A.first_out = 0;
  A.Ax = sAx; A.ub = su; A.ub_off = L1.row_s0; A.Z = nullptr; A.sums = nullptr; A.nonfinite = nullptr;
  auto kern = car_assemble_kernel<T, TO, kS, kCarWarps>;
  SAA_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
  kern<<<1, kCarWarps * 32, sizeof(Smem), st>>>(A);
  SAA_CUDA(h, cudaGetLastError());
  CarPickArgs P{};
  int rows[4];
  for (int c = 0; c < Lr.nu; ++c) {
    const int j = c / 2, nf = Lr.fin_rows(c, rows);
    for (int e = 0; e < Lr.relaxed_extra(c); ++e) {           // step k = j + 2 + e of sample 0
      if (P.n_ax >= 16) return fail(h, SAA_ERR_STATE, "internal: relaxed pick table");
      P.src_ax[P.n_ax] = L1.run_start(c) + e;
      P.dst_ax[P.n_ax] = Lr.ucol[c] + nf + e;
      ++P.n_ax;
    }
    (void)j;
  }
  for (int e = 0; e < Lr.keep_s; ++e) { P.src_u[e] = L1.row_s0 + e; P.dst_u[e] = Lr.row_s0 + e; }
  P.n_u = Lr.keep_s;
  car_relaxed_pick_kernel<TO><<<1, 32, 0, st>>>(P, sAx, su, (TO *)Ax, (TO *)u);
  SAA_CUDA(h, cudaGetLastError());
  return SAA_OK;
}

}  // namespace

extern "C" {

int saa_set_params_car(saa_handle *h, const saa_car_params *p) {
  if (!h || !p) return fail(h, SAA_ERR_ARG, "NULL argument");
  if (h->problem != SAA_CAR) return fail(h, SAA_ERR_ARG, "handle is not a car problem");
  if (!(p->dt > 0)) return fail(h, SAA_ERR_ARG, "dt must be positive");
  h->cp = *p; h->params_set = true;
  return SAA_OK;
}

int saa_set_samples_car(saa_handle *h, const double *states_init, const double *w_s, const double *w_r,
                        const double *DWs, void *stream) {
  if (!h || !states_init || !w_s || !w_r || !DWs) return fail(h, SAA_ERR_ARG, "NULL argument");
  if (h->problem != SAA_CAR) return fail(h, SAA_ERR_ARG, "handle is not a car problem");
  SAA_CUDA(h, cudaSetDevice(h->device));
  const i64 M = h->M_local;
  h->Mpad = (M + 31) / 32 * 32;
  const size_t es = kInSize;
  if (!h->d_a) {
    void *a = nullptr, *b = nullptr, *c = nullptr, *d = nullptr;   // commit only if all succeed
    cudaError_t e = cudaMalloc(&a, h->Mpad * es * 4);
    if (e == cudaSuccess) e = cudaMalloc(&b, h->Mpad * es * 2);
    if (e == cudaSuccess) e = cudaMalloc(&c, h->Mpad * es * 2 * h->S);
    if (e == cudaSuccess) e = cudaMalloc(&d, sizeof(int));
    if (e != cudaSuccess) {
      cudaFree(a); cudaFree(b); cudaFree(c); cudaFree(d);
      return fail(h, SAA_ERR_CUDA, std::string("cudaMalloc (packed samples): ") + cudaGetErrorString(e));
    }
    h->d_a = a; h->d_b = b; h->d_c = c; h->d_d = d;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int threads = 128;
  const int blocks = (int)((h->Mpad + threads - 1) / threads);
  SAA_CUDA(h, cudaMemsetAsync(h->d_d, 0, sizeof(int), st));
  car_check_ego_kernel<<<blocks, threads, 0, st>>>(states_init, M, (int *)h->d_d);
  car_pack_kernel<double><<<blocks, threads, 0, st>>>(states_init, w_s, w_r, DWs, M, h->Mpad, h->S,
                                                    (double *)h->d_a, (double *)h->d_b, (double *)h->d_c);
  SAA_CUDA(h, cudaGetLastError());
  int flag = 0;
  SAA_CUDA(h, cudaMemcpyAsync(&flag, h->d_d, sizeof(int), cudaMemcpyDeviceToHost, st));
  SAA_CUDA(h, cudaMemcpyAsync(h->car_ego0, states_init, 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
  SAA_CUDA(h, cudaStreamSynchronize(st));
  if (flag)
    return fail(h, SAA_ERR_ARG, "states_init[:, :4] (ego) must be identical for all samples "
                                "(the reference perturbs only the pedestrian part)");
  h->samples_set = true;
  return SAA_OK;
}

}  // extern "C"
