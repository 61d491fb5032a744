// Host-side launchers of the means-only pass, the tail selection and the sample gather
// (included by saa_b200.cu).
namespace {

// Expectation rows and Z_i of ALL local samples without touching the matrix: the three per-axis
// mean kernels (drone) or the sample-independent final rows (car), plus the rollout kernel for Z.
template <typename TO>
int launch_drone_means(saa_handle *h, const double *us, void *Z, double *sums, cudaStream_t st) {
  using T = double;
  using Args = DroneArgs<T, TO, kS>;
  Args A{};
  A.mass = (const T *)h->d_a; A.dw = (const T *)h->d_b; A.q = (const T *)h->d_c;
  A.M = h->M_local; A.Mpad = h->Mpad;
  fill_drone_common<T>(h, us, A.us, A.dt, A.noise_c, A.drag, A.kp, A.kd, A.x0, A.oc);
  for (int i = 0; i < 6; ++i) A.xf[i] = (T)h->dp.x_final[i];
  constexpr int kZWarps = 4;
  const int gridz = (int)std::max<i64>(1, std::min<i64>((h->M_local + kZWarps * 32 - 1) / (kZWarps * 32), (i64)h->n_sms * 2));
  const int n = DroneRed<kS>::N;
  // own scratch: this pass may run on a side stream while an assemble launch uses d_partials
  const i64 need = (i64)3 * gridz * n;
  if (h->means_scratch_len < need) {
    if (h->d_means_scratch) cudaFree(h->d_means_scratch);
    h->d_means_scratch = nullptr; h->means_scratch_len = 0;
    SAA_CUDA(h, cudaMalloc(&h->d_means_scratch, need * sizeof(double)));
    h->means_scratch_len = need;
  }
  for (int axis = 0; axis < 3; ++axis) {
    drone_axis_mean_kernel<T, TO, kS, kZWarps><<<gridz, kZWarps * 32, 0, st>>>(A, axis, h->d_means_scratch + (i64)axis * gridz * n);
    SAA_CUDA(h, cudaGetLastError());
  }
  reduce_partials_kernel<double><<<(n + 3) / 4, 128, 0, st>>>(h->d_means_scratch, 3 * gridz, n, sums);
  SAA_CUDA(h, cudaGetLastError());
  if (Z) return launch_drone_rollout<TO>(h, us, nullptr, Z, 0.0, 0.0, 0.0, nullptr, st);
  return SAA_OK;
}

template <typename T>
int launch_select(saa_handle *h, const void *Z, i64 K, i64 *idx_out, cudaStream_t st) {
  const i64 n = h->M_local;
  const i64 nblk = (n + kSelThreads * kSelItems - 1) / (kSelThreads * kSelItems);
  // scratch: [SelectState (4 x 8 B)] [hist 256 x 4 B = 128 doubles] [counts 2 nblk]
  int rc = ensure_scratch(h, 4 + 128 + 2 * nblk);
  if (rc) return rc;
  SelectState *state = (SelectState *)h->d_partials;
  unsigned *hist = (unsigned *)(h->d_partials + 4);
  i64 *counts = (i64 *)(h->d_partials + 4 + 128);
  SelectState init{0ull, (long long)K, 0ll};
  SAA_CUDA(h, cudaMemcpyAsync(state, &init, sizeof(init), cudaMemcpyHostToDevice, st));
  SAA_CUDA(h, cudaMemsetAsync(hist, 0, 256 * sizeof(unsigned), st));
  const int hgrid = (int)std::max<i64>(1, std::min<i64>((n + 255) / 256, (i64)h->n_sms * 8));
  for (int shift = SelectBits<T>::value - 8; shift >= 0; shift -= 8) {
    select_hist_kernel<T><<<hgrid, 256, 0, st>>>((const T *)Z, n, shift, state, hist);
    select_pick_kernel<<<1, 256, 0, st>>>(hist, shift, state);
  }
  SAA_CUDA(h, cudaGetLastError());
  select_count_kernel<T><<<(int)nblk, kSelThreads, 0, st>>>((const T *)Z, n, state, counts);
  select_scan_kernel<<<1, 1024, 0, st>>>(counts, nblk);
  select_write_kernel<T><<<(int)nblk, kSelThreads, 0, st>>>((const T *)Z, n, state, counts, idx_out);
  SAA_CUDA(h, cudaGetLastError());
  return SAA_OK;
}

int launch_gather(saa_handle *dst, const saa_handle *src, const i64 *idx, cudaStream_t st) {
  using T = double;                                              // the packed samples are always double
  const i64 K = dst->M_local;
  const int S = src->S;
  const int rows_a = src->problem == SAA_DRONE ? 1 : 4;        // mass | pedestrian initial state
  const int rows_b = src->problem == SAA_DRONE ? 3 * S : 2;    // dw   | omegas
  const int rows_c = src->problem == SAA_DRONE ? 6 : 2 * S;    // Q    | dw
  dst->Mpad = (K + kTileSamples - 1) / kTileSamples * kTileSamples;
  const size_t es = kInSize;
  if (!dst->d_a) {
    void *a = nullptr, *b = nullptr, *c = nullptr;               // commit only if all succeed
    cudaError_t e = cudaMalloc(&a, dst->Mpad * es * rows_a);
    if (e == cudaSuccess) e = cudaMalloc(&b, dst->Mpad * es * rows_b);
    if (e == cudaSuccess) e = cudaMalloc(&c, dst->Mpad * es * rows_c);
    if (e != cudaSuccess) {
      cudaFree(a); cudaFree(b); cudaFree(c);
      return fail(dst, SAA_ERR_CUDA, std::string("cudaMalloc (gathered samples): ") + cudaGetErrorString(e));
    }
    dst->d_a = a; dst->d_b = b; dst->d_c = c;
  }
  const int blocks = (int)((dst->Mpad + 255) / 256);
  gather_rows_kernel<T><<<blocks, 256, 0, st>>>((const T *)src->d_a, src->Mpad, (T *)dst->d_a, dst->Mpad, rows_a, idx, K);
  gather_rows_kernel<T><<<blocks, 256, 0, st>>>((const T *)src->d_b, src->Mpad, (T *)dst->d_b, dst->Mpad, rows_b, idx, K);
  gather_rows_kernel<T><<<blocks, 256, 0, st>>>((const T *)src->d_c, src->Mpad, (T *)dst->d_c, dst->Mpad, rows_c, idx, K);
  SAA_CUDA(dst, cudaGetLastError());
  if (src->problem == SAA_CAR) std::memcpy(dst->car_ego0, src->car_ego0, sizeof(dst->car_ego0));
  dst->samples_set = true;
  return SAA_OK;
}

}  // namespace

extern "C" {

int saa_linearize_means(saa_handle *h, const double *us, void *Z, double *mean_sums, void *stream) {
  if (!h || !us || !mean_sums) return fail(h, SAA_ERR_ARG, "NULL argument");
  if (h->problem == SAA_HOPPER) return fail(h, SAA_ERR_ARG, "the hopper has no expectation rows");
  if (!h->params_set || !h->samples_set) return fail(h, SAA_ERR_STATE, "set params and samples first");
  if (h->S != kS) return fail(h, SAA_ERR_ARG, "the means-only pass (tail-reduced subproblem) is implemented for S = 20");
  SAA_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (h->problem == SAA_DRONE)
    return h->precision == 64 ? launch_drone_means<double>(h, us, Z, mean_sums, st)
                              : launch_drone_means<float>(h, us, Z, mean_sums, st);
  // car: the final rows do not depend on the samples (relaxed-iteration path of the assemble launch)
  int rc = h->precision == 64 ? launch_car_assemble<double>(h, us, 0, nullptr, nullptr, nullptr, mean_sums, st)
                              : launch_car_assemble<float>(h, us, 0, nullptr, nullptr, nullptr, mean_sums, st);
  if (rc || !Z) return rc;
  return h->precision == 64 ? launch_car_rollout<double>(h, us, nullptr, Z, 0.0, 0.0, 0.0, nullptr, st)
                            : launch_car_rollout<float>(h, us, nullptr, Z, 0.0, 0.0, 0.0, nullptr, st);
}

int saa_select_tail(saa_handle *h, const void *Z, int64_t K, int64_t *idx_out, void *stream) {
  if (!h || !Z || !idx_out) return fail(h, SAA_ERR_ARG, "NULL argument");
  if (K < 1 || K > h->M_local) return fail(h, SAA_ERR_ARG, "need 1 <= K <= M_local");
  SAA_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  return h->precision == 64 ? launch_select<double>(h, Z, K, (i64 *)idx_out, st)
                            : launch_select<float>(h, Z, K, (i64 *)idx_out, st);
}

int saa_gather_samples(saa_handle *dst, const saa_handle *src, const int64_t *idx, void *stream) {
  if (!dst || !src || !idx) return fail(dst, SAA_ERR_ARG, "NULL argument");
  if (dst->problem != src->problem || dst->S != src->S || dst->precision != src->precision ||
      dst->device != src->device)
    return fail(dst, SAA_ERR_ARG, "source and destination handles must agree in problem, S, precision and device");
  if (src->problem == SAA_HOPPER) return fail(dst, SAA_ERR_ARG, "not available for the hopper");
  if (!src->samples_set) return fail(dst, SAA_ERR_STATE, "the source handle has no samples");
  if (dst->M_local > src->M_local) return fail(dst, SAA_ERR_ARG, "destination holds more samples than the source");
  SAA_CUDA(dst, cudaSetDevice(dst->device));
  cudaStream_t st = (cudaStream_t)stream;
  return launch_gather(dst, src, (const i64 *)idx, st);
}

}  // extern "C"
