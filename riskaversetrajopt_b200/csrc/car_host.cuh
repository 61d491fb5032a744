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

// scp_iter < 1: rows >= n_x are multiplied by exactly 0 (car/driving.py:411-415) and vanish
// from the pattern; only the sample-independent final rows carry values.
template <typename T>
int launch_car_assemble(saa_handle *h, const double *us, int scp_iter, void *Ax, void *u, void *Z,
                        double *sums, cudaStream_t st) {
  using Args = CarArgs<T, kS>;
  using Smem = CarSmem<T, kS, kCarWarps>;
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
  A.Ax = relaxed ? nullptr : (T *)Ax;
  A.ub = relaxed ? nullptr : (T *)u;
  A.ub_off = L.row_s0 + h->first_out * L.R;
  A.Z = relaxed ? nullptr : (T *)Z;
  A.sums = sums;
  const i64 ntiles = (h->M_local + kTileSamples - 1) / kTileSamples;
  const int grid = relaxed ? 1 : grid_for(h, ntiles, kCarWarps, 1);
  auto kern = car_assemble_kernel<T, kS, kCarWarps>;
  SAA_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
  kern<<<grid, kCarWarps * 32, sizeof(Smem), st>>>(A);
  SAA_CUDA(h, cudaGetLastError());
  return SAA_OK;
}

template <typename T>
int launch_car_rollout(saa_handle *h, const double *us, void *Xs, void *Z, double t_risk,
                       double sat_tol, double ztol, double *out3, cudaStream_t st) {
  constexpr int W = 4;
  using Args = CarRollArgs<T, kS>;
  Args A{};
  fill_car_common<T>(h, us, A);
  A.Xs = (T *)Xs; A.Z = (T *)Z;
  A.ztol = (T)ztol; A.t_risk = (T)t_risk; A.sat_tol = (T)sat_tol;
  const i64 ntiles = (h->M_local + 31) / 32;
  const int grid = grid_for(h, ntiles, W, 4);
  int rc = ensure_scratch(h, (i64)grid * 3);
  if (rc) return rc;
  A.partials = out3 ? h->d_partials : nullptr;
  const size_t full = (size_t)W * 32 * (((kS + 1) * 8) | 1) * sizeof(T);
  auto kern = car_rollout_kernel<T, kS, W>;
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
struct CarRelaxArgs {
  i64 M_local, first_out, M_out;
  int nu, keep;
  i64 ycol0, slackcol, tcol, n_rows, row_cvar, row_y0, row_ctrl0;
  double cvar_t, u_max;
  i64 ucol_last[64];
};

template <typename T>
__global__ void car_relaxed_constants_kernel(const __grid_constant__ CarRelaxArgs C, int write_shared,
                                             T *Ax, T *l, T *u) {
  const i64 tid = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  const i64 nth = (i64)gridDim.x * blockDim.x;
  const T inf = (T)INFINITY;
  const T nan = inf * (T)0;                                   // -inf * 0 (car/driving.py:414)
  // y columns: y_i for i < keep holds (CVaR row, -y_i row), the others only the CVaR row
  for (i64 i = tid; i < C.M_local; i += nth) {
    const i64 gi = C.first_out + i;
    const i64 pos = C.ycol0 + gi + (gi < C.keep ? gi : C.keep);
    Ax[pos] = (T)1.0;
    if (gi < C.keep) Ax[pos + 1] = (T)-1.0;
  }
  // rows >= n_x of the constraint block: l = nan, u = 0; this rank's share of them
  const i64 R = 20;   // car: rows per sample
  const i64 r_lo = C.row_y0 + C.M_out + C.first_out * R, r_hi = r_lo + C.M_local * R;
  for (i64 r = r_lo + tid; r < r_hi; r += nth) { l[r] = nan; u[r] = (T)0; }
  for (i64 i = tid; i < C.M_local; i += nth) {
    const i64 r = C.row_y0 + C.first_out + i;
    l[r] = (r < 8) ? -inf : nan;
    u[r] = (T)0;
  }
  if (!write_shared) return;
  if (tid < C.nu) {
    Ax[C.ucol_last[tid]] = (T)1.0;
    l[C.row_ctrl0 + tid] = (T)(-C.u_max);
    u[C.row_ctrl0 + tid] = (T)C.u_max;
  }
  if (tid == 0) {
    Ax[C.slackcol] = (T)1.0;
    for (int i = 0; i < C.keep; ++i) Ax[C.slackcol + 1 + i] = (T)-1.0;
    Ax[C.tcol] = (T)C.cvar_t;
    l[C.row_cvar] = -inf; u[C.row_cvar] = (T)0;
    l[C.row_ctrl0 - 1] = nan; u[C.row_ctrl0 - 1] = (T)0;      // "-slack <= 0" row, zeroed too
  }
}

int car_write_constants_relaxed(saa_handle *h, int write_shared, void *Ax, void *l, void *u,
                                cudaStream_t st) {
  if (h->method != SAA_METHOD_SAA)
    return fail(h, SAA_ERR_ARG, "car baseline at scp_iter 0 is not supported (reference multiplies +-inf bounds by 0)");
  if (h->M_out < 3) return fail(h, SAA_ERR_ARG, "car scp_iter 0 needs M >= 3 (rows < n_x must be risk rows)");
  Layout L; L.build(SAA_CAR, h->method, h->S, h->M_out, true);
  CarRelaxArgs C{};
  C.M_local = h->M_local; C.first_out = h->first_out; C.M_out = h->M_out;
  C.nu = L.nu; C.keep = 3;
  C.ycol0 = L.ycol0; C.slackcol = L.slackcol; C.tcol = L.tcol; C.n_rows = L.n_rows;
  C.row_cvar = L.row_cvar; C.row_y0 = L.row_y0; C.row_ctrl0 = L.row_ctrl0;
  C.cvar_t = (double)h->M_global * h->alpha;
  C.u_max = h->cp.u_max;
  for (int c = 0; c < L.nu; ++c) C.ucol_last[c] = L.ucol[c + 1] - 1;
  const int threads = 256;
  const int blocks = (int)std::min<i64>((h->M_local * 20 + threads - 1) / threads, (i64)h->n_sms * 8);
  if (h->precision == 64)
    car_relaxed_constants_kernel<double><<<std::max(blocks, 1), threads, 0, st>>>(C, write_shared, (double *)Ax, (double *)l, (double *)u);
  else
    car_relaxed_constants_kernel<float><<<std::max(blocks, 1), threads, 0, st>>>(C, write_shared, (float *)Ax, (float *)l, (float *)u);
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
  const size_t es = esize(h);
  if (!h->d_a) {
    SAA_CUDA(h, cudaMalloc(&h->d_a, h->Mpad * es * 4));
    SAA_CUDA(h, cudaMalloc(&h->d_b, h->Mpad * es * 2));
    SAA_CUDA(h, cudaMalloc(&h->d_c, h->Mpad * es * 2 * h->S));
    SAA_CUDA(h, cudaMalloc(&h->d_d, sizeof(int)));
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int threads = 128;
  const int blocks = (int)((h->Mpad + threads - 1) / threads);
  SAA_CUDA(h, cudaMemsetAsync(h->d_d, 0, sizeof(int), st));
  car_check_ego_kernel<<<blocks, threads, 0, st>>>(states_init, M, (int *)h->d_d);
  if (h->precision == 64)
    car_pack_kernel<double><<<blocks, threads, 0, st>>>(states_init, w_s, w_r, DWs, M, h->Mpad, h->S,
                                                      (double *)h->d_a, (double *)h->d_b, (double *)h->d_c);
  else
    car_pack_kernel<float><<<blocks, threads, 0, st>>>(states_init, w_s, w_r, DWs, M, h->Mpad, h->S,
                                                     (float *)h->d_a, (float *)h->d_b, (float *)h->d_c);
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
