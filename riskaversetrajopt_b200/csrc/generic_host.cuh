// Host-side launchers of the horizon-generic kernels (included by saa_b200.cu): every S != 20.
namespace {

int gen_grid(const saa_handle *h) {
  const i64 want = (h->M_local + kGenThreads - 1) / kGenThreads;
  return (int)std::max<i64>(1, std::min<i64>(want, (i64)h->n_sms * 8));
}

template <typename TO>
void fill_gen_drone(const saa_handle *h, const double *us, GenDroneArgs<TO> &A) {
  A.mass = (const double *)h->d_a; A.dw = (const double *)h->d_b; A.q = (const double *)h->d_c;
  A.M = h->M_local; A.Mpad = h->Mpad; A.S = h->S;
  for (int i = 0; i < h->S * 3; ++i) A.us[i] = us[i];
  A.dt = h->dp.dt; A.noise_c = std::sqrt(h->dp.dt) * h->dp.beta;
  A.drag = h->dp.drag_coefficient; A.kp = h->dp.gain_p; A.kd = h->dp.gain_v;
  for (int i = 0; i < 6; ++i) { A.x0[i] = h->dp.x_init[i]; A.xf[i] = h->dp.x_final[i]; }
  for (int o = 0; o < 3; ++o) for (int a = 0; a < 2; ++a) A.oc[o][a] = h->dp.obs_positions[o][a];
}

template <typename TO>
int launch_drone_generic(saa_handle *h, const double *us, int scp_iter, void *Ax, void *u, void *Z,
                         double *sums, cudaStream_t st) {
  GenDroneArgs<TO> A{};
  fill_gen_drone(h, us, A);
  double mult, pad, scale, bound;
  drone_mult(h, &mult, &pad); drone_relax(h, &scale, &bound);
  const bool relaxed = scp_iter < 2;
  A.escale = relaxed ? mult * scale : mult; A.ubscale = mult; A.ubpad = pad; A.ztol = 0.0;
  A.Ax = (TO *)Ax; A.M_out = h->M_out; A.first_out = h->first_out;
  A.ub = relaxed ? nullptr : (TO *)u; A.ub_off = h->lay.row_s0 + h->first_out * h->lay.R;
  A.Z = (TO *)Z;
  const int grid = gen_grid(h), n = (int)saa_mean_len(h);
  const i64 nrows = (i64)grid * (kGenThreads / 32);
  int rc = ensure_scratch(h, nrows * n);
  if (rc) return rc;
  A.partials = h->d_partials;
  drone_generic_kernel<TO><<<grid, kGenThreads, 0, st>>>(A);
  SAA_CUDA(h, cudaGetLastError());
  gen_reduce_kernel<<<(n + 127) / 128, 128, 0, st>>>(h->d_partials, nrows, n, 0, sums);
  SAA_CUDA(h, cudaGetLastError());
  return SAA_OK;
}

template <typename TO>
int launch_drone_generic_rollout(saa_handle *h, const double *us, void *Xs, void *Z, double t_risk, double sat_tol,
                                 double ztol, double *out3, cudaStream_t st) {
  GenDroneArgs<TO> A{};
  fill_gen_drone(h, us, A);
  A.Xs = (TO *)Xs; A.Z = (TO *)Z; A.ztol = ztol; A.t_risk = t_risk; A.sat_tol = sat_tol;
  const int grid = gen_grid(h);
  const i64 nrows = (i64)grid * (kGenThreads / 32);
  int rc = ensure_scratch(h, nrows * 3);
  if (rc) return rc;
  A.partials = out3 ? h->d_partials : nullptr;
  drone_generic_kernel<TO><<<grid, kGenThreads, 0, st>>>(A);
  SAA_CUDA(h, cudaGetLastError());
  if (out3) {
    gen_reduce_kernel<<<1, 32, 0, st>>>(h->d_partials, nrows, 3, 1, out3);
    SAA_CUDA(h, cudaGetLastError());
  }
  return SAA_OK;
}

template <typename TO>
void fill_gen_car(const saa_handle *h, const double *us, GenCarArgs<TO> &A) {
  A.x0 = (const double *)h->d_a; A.om = (const double *)h->d_b; A.dw = (const double *)h->d_c;
  A.M = h->M_local; A.Mpad = h->Mpad; A.S = h->S;
  for (int i = 0; i < h->S * 2; ++i) A.us[i] = us[i];
  for (int i = 0; i < 4; ++i) { A.ego0[i] = h->car_ego0[i]; A.goal[i] = h->cp.goal[i]; }
  A.dt = h->cp.dt; A.noise_c = std::sqrt(h->cp.dt) * h->cp.beta;
  A.v_des = h->cp.speed_ped_des; A.d_min = h->cp.min_separation_distance;
}

template <typename TO>
int launch_car_generic(saa_handle *h, const double *us, int scp_iter, void *Ax, void *u, void *Z, double *sums,
                       cudaStream_t st) {
  const bool relaxed = scp_iter < 1;
  GenCarArgs<TO> A{};
  fill_gen_car(h, us, A);
  A.ztol = 0.0;
  A.Ax = relaxed ? nullptr : (TO *)Ax; A.M_out = h->M_out; A.first_out = h->first_out;
  A.ub = relaxed ? nullptr : (TO *)u; A.ub_off = h->lay.row_s0 + h->first_out * h->lay.R;
  A.Z = relaxed ? nullptr : (TO *)Z;
  A.sums = sums; A.nonfinite = relaxed ? nullptr : h->d_nonfinite;
  car_generic_kernel<TO><<<relaxed ? 1 : gen_grid(h), kGenThreads, 0, st>>>(A);
  SAA_CUDA(h, cudaGetLastError());
  if (!relaxed || Ax == nullptr) return SAA_OK;
  // scp_iter 0: surviving separation rows of sample 0 (see launch_car_relaxed_rows)
  Layout Lr; Lr.build(SAA_CAR, h->method, h->S, h->M_out, true);
  if (Lr.keep_s == 0 || h->first_out != 0) return SAA_OK;
  Layout L1; L1.build(SAA_CAR, h->method, h->S, 1, false);
  const i64 need = (L1.nnz + L1.n_rows) * (i64)sizeof(TO);
  if (h->relax_scratch_bytes < need) {
    if (h->d_relax_scratch) cudaFree(h->d_relax_scratch);
    h->d_relax_scratch = nullptr; h->relax_scratch_bytes = 0;
    SAA_CUDA(h, cudaMalloc(&h->d_relax_scratch, need));
    h->relax_scratch_bytes = need;
  }
  TO *sAx = (TO *)h->d_relax_scratch, *su = sAx + L1.nnz;
  GenCarArgs<TO> B{};
  fill_gen_car(h, us, B);
  B.M = 1; B.M_out = 1; B.first_out = 0; B.Ax = sAx; B.ub = su; B.ub_off = L1.row_s0;
  car_generic_kernel<TO><<<1, kGenThreads, 0, st>>>(B);
  SAA_CUDA(h, cudaGetLastError());
  CarPickArgs P{};
  int rows[4];
  for (int c = 0; c < Lr.nu; ++c) {
    const int nf = Lr.fin_rows(c, rows);
    for (int e = 0; e < Lr.relaxed_extra(c); ++e) {
      if (P.n_ax >= 16) return fail(h, SAA_ERR_STATE, "internal: relaxed pick table");
      P.src_ax[P.n_ax] = L1.run_start(c) + e; P.dst_ax[P.n_ax] = Lr.ucol[c] + nf + e; ++P.n_ax;
    }
  }
  for (int e = 0; e < Lr.keep_s; ++e) { P.src_u[e] = L1.row_s0 + e; P.dst_u[e] = Lr.row_s0 + e; }
  P.n_u = Lr.keep_s;
  car_relaxed_pick_kernel<TO><<<1, 32, 0, st>>>(P, sAx, su, (TO *)Ax, (TO *)u);
  SAA_CUDA(h, cudaGetLastError());
  return SAA_OK;
}

template <typename TO>
int launch_car_generic_rollout(saa_handle *h, const double *us, void *Xs, void *Z, double t_risk, double sat_tol,
                               double ztol, double *out3, cudaStream_t st) {
  GenCarArgs<TO> A{};
  fill_gen_car(h, us, A);
  A.Xs = (TO *)Xs; A.Z = (TO *)Z; A.ztol = ztol; A.t_risk = t_risk; A.sat_tol = sat_tol;
  A.nonfinite = h->d_nonfinite;
  const int grid = gen_grid(h);
  const i64 nrows = (i64)grid * (kGenThreads / 32);
  int rc = ensure_scratch(h, nrows * 3);
  if (rc) return rc;
  A.partials = h->d_partials;
  car_generic_kernel<TO><<<grid, kGenThreads, 0, st>>>(A);
  SAA_CUDA(h, cudaGetLastError());
  if (out3) {
    gen_reduce_kernel<<<1, 32, 0, st>>>(h->d_partials, nrows, 3, 1, out3);
    SAA_CUDA(h, cudaGetLastError());
  }
  return SAA_OK;
}

}  // namespace
