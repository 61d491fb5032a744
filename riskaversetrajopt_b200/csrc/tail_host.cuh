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

// ---- distributed exact selection: the radix select with its histogram exposed between the passes ----
struct SelectScratch { SelectState *state; unsigned *hist; i64 *counts; i64 nblk; };

int select_scratch(saa_handle *h, SelectScratch &sc) {
  const i64 n = h->M_local;
  sc.nblk = (n + kSelThreads * kSelItems - 1) / (kSelThreads * kSelItems);
  const i64 need = 4 + 128 + 2 * std::max<i64>(sc.nblk, 1) + 2;
  if (h->select_len < need) {
    if (h->d_select) cudaFree(h->d_select);
    h->d_select = nullptr; h->select_len = 0;
    SAA_CUDA(h, cudaMalloc(&h->d_select, need * sizeof(double)));
    h->select_len = need;
  }
  sc.state = (SelectState *)h->d_select;
  sc.hist = (unsigned *)(h->d_select + 4);
  sc.counts = (i64 *)(h->d_select + 4 + 128);
  return SAA_OK;
}

__global__ void select_totals_kernel(const i64 *__restrict__ counts, i64 nblk, i64 *__restrict__ out2) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    i64 gt = 0, eq = 0;
    for (i64 b = 0; b < nblk; ++b) { gt += counts[2 * b]; eq += counts[2 * b + 1]; }
    out2[0] = gt; out2[1] = eq;
  }
}
__global__ void select_set_ties_kernel(SelectState *st, long long take) { if (threadIdx.x == 0) st->k_rem = take; }

int launch_gather(saa_handle *dst, const saa_handle *src, const i64 *idx, cudaStream_t st) {
  using T = double;                                              // the packed samples are always double
  const i64 K = dst->M_local;
  if (K == 0) { dst->samples_set = true; return SAA_OK; }        // saa_set_active(0): nothing to gather
  const int S = src->S;
  const int rows_a = src->problem == SAA_DRONE ? 1 : 4;        // mass | pedestrian initial state
  const int rows_b = src->problem == SAA_DRONE ? 3 * S : 2;    // dw   | omegas
  const int rows_c = src->problem == SAA_DRONE ? 6 : 2 * S;    // Q    | dw
  dst->Mpad = (dst->M_cap + kTileSamples - 1) / kTileSamples * kTileSamples;   // stride of the packed rows: the capacity
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
  // entries [0, roundup(K, 32)) of every packed row are written (the padding repeats the last sample)
  const i64 nw = std::min<i64>(dst->Mpad, (K + 31) / 32 * 32);
  const int blocks = (int)((nw + 255) / 256);
  gather_rows_kernel<T><<<blocks, 256, 0, st>>>((const T *)src->d_a, src->Mpad, (T *)dst->d_a, dst->Mpad, nw, rows_a, idx, K);
  gather_rows_kernel<T><<<blocks, 256, 0, st>>>((const T *)src->d_b, src->Mpad, (T *)dst->d_b, dst->Mpad, nw, rows_b, idx, K);
  gather_rows_kernel<T><<<blocks, 256, 0, st>>>((const T *)src->d_c, src->Mpad, (T *)dst->d_c, dst->Mpad, nw, rows_c, idx, K);
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

int saa_select_passes(const saa_handle *h) { return h ? (h->precision == 64 ? 8 : 4) : 0; }

int saa_select_begin(saa_handle *h, int64_t K_global, void *stream) {
  if (!h) return fail(h, SAA_ERR_ARG, "NULL handle");
  if (K_global < 1) return fail(h, SAA_ERR_ARG, "need K_global >= 1");
  SAA_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  SelectScratch sc;
  if (int rc = select_scratch(h, sc)) return rc;
  SelectState init{0ull, (long long)K_global, 0ll};
  SAA_CUDA(h, cudaMemcpyAsync(sc.state, &init, sizeof(init), cudaMemcpyHostToDevice, st));
  return SAA_OK;
}

int saa_select_pass_hist(saa_handle *h, const void *Z, int pass, uint32_t *hist, void *stream) {
  if (!h || !Z || !hist) return fail(h, SAA_ERR_ARG, "NULL argument");
  const int bits = h->precision == 64 ? 64 : 32, shift = bits - 8 * (pass + 1);
  if (pass < 0 || shift < 0) return fail(h, SAA_ERR_ARG, "bad pass index");
  SAA_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  SelectScratch sc;
  if (int rc = select_scratch(h, sc)) return rc;
  SAA_CUDA(h, cudaMemsetAsync(hist, 0, 256 * sizeof(unsigned), st));
  const i64 n = h->M_local;
  if (n > 0) {
    const int hgrid = (int)std::max<i64>(1, std::min<i64>((n + 255) / 256, (i64)h->n_sms * 8));
    if (h->precision == 64) select_hist_kernel<double><<<hgrid, 256, 0, st>>>((const double *)Z, n, shift, sc.state, hist);
    else select_hist_kernel<float><<<hgrid, 256, 0, st>>>((const float *)Z, n, shift, sc.state, hist);
    SAA_CUDA(h, cudaGetLastError());
  }
  return SAA_OK;
}

int saa_select_pass_pick(saa_handle *h, uint32_t *hist, int pass, void *stream) {
  if (!h || !hist) return fail(h, SAA_ERR_ARG, "NULL argument");
  const int bits = h->precision == 64 ? 64 : 32, shift = bits - 8 * (pass + 1);
  if (pass < 0 || shift < 0) return fail(h, SAA_ERR_ARG, "bad pass index");
  SAA_CUDA(h, cudaSetDevice(h->device));
  SelectScratch sc;
  if (int rc = select_scratch(h, sc)) return rc;
  select_pick_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(hist, shift, sc.state);
  SAA_CUDA(h, cudaGetLastError());
  return SAA_OK;
}

int saa_select_counts(saa_handle *h, const void *Z, int64_t *gt_eq, void *stream) {
  if (!h || !Z || !gt_eq) return fail(h, SAA_ERR_ARG, "NULL argument");
  SAA_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  SelectScratch sc;
  if (int rc = select_scratch(h, sc)) return rc;
  const i64 n = h->M_local;
  if (n == 0) { SAA_CUDA(h, cudaMemsetAsync(gt_eq, 0, 2 * sizeof(int64_t), st)); return SAA_OK; }
  if (h->precision == 64) select_count_kernel<double><<<(int)sc.nblk, kSelThreads, 0, st>>>((const double *)Z, n, sc.state, sc.counts);
  else select_count_kernel<float><<<(int)sc.nblk, kSelThreads, 0, st>>>((const float *)Z, n, sc.state, sc.counts);
  select_totals_kernel<<<1, 32, 0, st>>>(sc.counts, sc.nblk, (i64 *)gt_eq);
  SAA_CUDA(h, cudaGetLastError());
  return SAA_OK;
}

int saa_select_finish(saa_handle *h, const void *Z, int64_t take_ties, int64_t *idx_out, void *stream) {
  if (!h || !Z || !idx_out) return fail(h, SAA_ERR_ARG, "NULL argument");
  if (take_ties < 0) return fail(h, SAA_ERR_ARG, "take_ties must be >= 0");
  SAA_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  SelectScratch sc;
  if (int rc = select_scratch(h, sc)) return rc;
  const i64 n = h->M_local;
  if (n == 0) return SAA_OK;
  select_set_ties_kernel<<<1, 32, 0, st>>>(sc.state, (long long)take_ties);
  select_scan_kernel<<<1, 1024, 0, st>>>(sc.counts, sc.nblk);          // counts were left by saa_select_counts
  if (h->precision == 64) select_write_kernel<double><<<(int)sc.nblk, kSelThreads, 0, st>>>((const double *)Z, n, sc.state, sc.counts, (i64 *)idx_out);
  else select_write_kernel<float><<<(int)sc.nblk, kSelThreads, 0, st>>>((const float *)Z, n, sc.state, sc.counts, (i64 *)idx_out);
  SAA_CUDA(h, cudaGetLastError());
  return SAA_OK;
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
