// Host-side launchers of the hopper kernels (included by saa_b200.cu).
namespace {

#ifndef SAA_HOPPER_SMEM_KB
#define SAA_HOPPER_SMEM_KB 36   // feature staging per block (48 B per feature): 6 blocks (24 warps) per SM
#endif

struct HopperOut {
  void *mu = nullptr, *dmu = nullptr, *g = nullptr, *jac = nullptr, *Z = nullptr;
  const double *y = nullptr, *lambda = nullptr;
  double *hess_sums = nullptr;      // [n_c][2]
  double *out3 = nullptr;           // CVaR terms
  double sat_tol = 0.0;
};

template <typename TO>
int launch_hopper(saa_handle *h, const saa_hopper_point &pt, const HopperOut &o, cudaStream_t st) {
  using T = double;
  const int n_c = pt.n_c;
  HopperArgs<T, TO> A{};
  A.I = (const T *)h->d_a; A.theta = (const T *)h->d_b; A.tau = (const T *)h->d_c;
  A.M = h->M_local; A.F = h->n_feat; A.n_c = n_c; A.mu_nom = (T)h->mu_nom;
  A.saa = h->method == SAA_METHOD_SAA;
  for (int c = 0; c < n_c; ++c) {
    A.px[c] = pt.px[c]; A.fx[c] = pt.fx[c]; A.fz[c] = pt.fz[c];
    A.gp[0][c] = 1.0; A.gp[1][c] = pt.x3[c] * std::cos(pt.x2[c]); A.gp[2][c] = std::sin(pt.x2[c]);
  }
  A.t_risk = pt.t_risk; A.slack = pt.slack; A.y = o.y;
  A.mu = (TO *)o.mu; A.dmu = (TO *)o.dmu; A.g = (TO *)o.g; A.jac = (TO *)o.jac;
  A.lambda = o.lambda; A.Z = (TO *)o.Z; A.sat_tol = o.sat_tol;
  const bool hess = o.lambda != nullptr, cvar = o.out3 != nullptr || o.Z != nullptr;
  // chunk of samples staged per block iteration: features (4 doubles each) within the budget, and a
  // multiple of the block size in (sample, contact) pairs when one exists
  const size_t per_sample = sizeof(HopperFeat<T>) * (size_t)h->n_feat +
                            ((hess ? 2 : 0) + (cvar ? 1 : 0)) * sizeof(double) * (size_t)n_c;
  int cap = (int)std::max<size_t>(1, std::min<size_t>(64, ((size_t)SAA_HOPPER_SMEM_KB * 1024) / per_sample));
  int chunk = cap;
  for (int c = cap; c >= std::max(1, (3 * cap) / 4); --c)
    if ((c * n_c) % kHopperThreads == 0) { chunk = c; break; }
  A.chunk = chunk;
  const size_t smem = per_sample * (size_t)chunk;
  if (smem > 200 * 1024) return fail(h, SAA_ERR_ARG, "n_features too large for the shared-memory staging");
  const i64 nchunks = (h->M_local + chunk - 1) / chunk;
  const int blocks = (int)std::max<i64>(1, std::min<i64>(nchunks, (i64)h->n_sms * 16));
  int rc = ensure_scratch(h, (i64)blocks * (2 * n_c + 3));
  if (rc) return rc;
  A.w = h->d_partials;
  A.cvar = o.out3 ? h->d_partials + (i64)blocks * 2 * n_c : nullptr;
  auto go = [&](auto kern) -> int {
    SAA_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<blocks, kHopperThreads, smem, st>>>(A);
    SAA_CUDA(h, cudaGetLastError());
    return SAA_OK;
  };
  if (hess && cvar) rc = go(hopper_friction_kernel<T, TO, true, true>);
  else if (hess) rc = go(hopper_friction_kernel<T, TO, true, false>);
  else if (cvar) rc = go(hopper_friction_kernel<T, TO, false, true>);
  else rc = go(hopper_friction_kernel<T, TO, false, false>);
  if (rc) return rc;
  if (hess) {
    hopper_reduce_kernel<<<(2 * n_c * 32 + 255) / 256, 256, 0, st>>>(h->d_partials, blocks, n_c, o.hess_sums);
    SAA_CUDA(h, cudaGetLastError());
  }
  if (o.out3) {
    reduce_cvar_kernel<<<1, 32, 0, st>>>(h->d_partials + (i64)blocks * 2 * n_c, blocks, o.out3);
    SAA_CUDA(h, cudaGetLastError());
  }
  return SAA_OK;
}

int hopper_dispatch(saa_handle *h, const saa_hopper_point &pt, const HopperOut &o, cudaStream_t st) {
  return h->precision == 64 ? launch_hopper<double>(h, pt, o, st) : launch_hopper<float>(h, pt, o, st);
}

int hopper_check(saa_handle *h, const saa_hopper_point *pt) {
  if (!h || !pt) return fail(h, SAA_ERR_ARG, "NULL argument");
  if (h->problem != SAA_HOPPER) return fail(h, SAA_ERR_ARG, "handle is not a hopper problem");
  if (!h->samples_set) return fail(h, SAA_ERR_STATE, "set the friction features first");
  if (pt->n_c < 1 || pt->n_c > kHopperMaxContacts) return fail(h, SAA_ERR_ARG, "need 1 <= n_c <= 32 contact instants");
  return SAA_OK;
}

// rows of the block that do not depend on the samples (saa): M alpha t + sum y, -y_i, trailing 0
template <typename TO>
int launch_hopper_head(saa_handle *h, const saa_hopper_point &pt, const double *y, void *g, cudaStream_t st) {
  const i64 M = h->M_local;
  const int blocks = (int)std::max<i64>(1, std::min<i64>((M + 255) / 256, (i64)h->n_sms * 4));
  int rc = ensure_scratch(h, blocks);
  if (rc) return rc;
  TO *gg = (TO *)g;
  hopper_head_rows_kernel<TO><<<blocks, 256, 0, st>>>(y, M, gg + 1, h->d_partials);
  hopper_head_finish_kernel<TO><<<1, 32, 0, st>>>(h->d_partials, blocks, (double)h->M_global * h->alpha * pt.t_risk,
                                                 gg, gg + 1 + M + M * pt.n_c);
  SAA_CUDA(h, cudaGetLastError());
  return SAA_OK;
}

}  // namespace

extern "C" {

int saa_set_samples_hopper(saa_handle *h, int32_t n_features, double mu_nom, const double *I,
                           const double *thetas, const double *taus, void *stream) {
  if (!h || !I || !thetas || !taus) return fail(h, SAA_ERR_ARG, "NULL argument");
  if (h->problem != SAA_HOPPER) return fail(h, SAA_ERR_ARG, "handle is not a hopper problem");
  if (n_features < 1) return fail(h, SAA_ERR_ARG, "n_features must be positive");
  SAA_CUDA(h, cudaSetDevice(h->device));
  const i64 n = h->M_local * n_features;
  const size_t es = kInSize;
  if (h->d_a && h->n_feat != n_features) {
    cudaFree(h->d_a); cudaFree(h->d_b); cudaFree(h->d_c);
    h->d_a = h->d_b = h->d_c = nullptr;
  }
  if (!h->d_a) {
    void *a = nullptr, *b = nullptr, *c = nullptr;       // commit to the handle only if all succeed
    cudaError_t e = cudaMalloc(&a, n * es);
    if (e == cudaSuccess) e = cudaMalloc(&b, n * es);
    if (e == cudaSuccess) e = cudaMalloc(&c, n * es);
    if (e != cudaSuccess) {
      cudaFree(a); cudaFree(b); cudaFree(c);
      return fail(h, SAA_ERR_CUDA, std::string("cudaMalloc (friction features): ") + cudaGetErrorString(e));
    }
    h->d_a = a; h->d_b = b; h->d_c = c;
  }
  h->n_feat = n_features; h->mu_nom = mu_nom;
  cudaStream_t st = (cudaStream_t)stream;
  const int threads = 256;
  const int blocks = (int)((n + threads - 1) / threads);
  hopper_pack_kernel<double><<<blocks, threads, 0, st>>>(I, thetas, taus, n, (double *)h->d_a, (double *)h->d_b, (double *)h->d_c);
  SAA_CUDA(h, cudaGetLastError());
  h->samples_set = true; h->params_set = true;
  return SAA_OK;
}

int saa_hopper_friction(saa_handle *h, int32_t n_c, const double *px, void *mu, void *dmu,
                        const double *lambda, double *hess_sums, void *stream) {
  if (!h || !px || !mu || !dmu) return fail(h, SAA_ERR_ARG, "NULL argument");
  if (n_c < 1 || n_c > kHopperMaxContacts) return fail(h, SAA_ERR_ARG, "need 1 <= n_c <= 32 contact instants");
  saa_hopper_point pt{};
  pt.n_c = n_c;
  for (int c = 0; c < n_c; ++c) pt.px[c] = px[c];
  if (int rc = hopper_check(h, &pt)) return rc;
  if ((lambda == nullptr) != (hess_sums == nullptr))
    return fail(h, SAA_ERR_ARG, "lambda_dev and hess_sums_dev go together");
  SAA_CUDA(h, cudaSetDevice(h->device));
  HopperOut o; o.mu = mu; o.dmu = dmu; o.lambda = lambda; o.hess_sums = hess_sums;
  return hopper_dispatch(h, pt, o, (cudaStream_t)stream);
}

int saa_hopper_g(saa_handle *h, const saa_hopper_point *pt, const double *y, void *g, void *stream) {
  if (int rc = hopper_check(h, pt)) return rc;
  if (!g) return fail(h, SAA_ERR_ARG, "NULL argument");
  const bool saa = h->method == SAA_METHOD_SAA;
  if (saa && !y) return fail(h, SAA_ERR_ARG, "method saa needs the risk variables y_dev");
  SAA_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t es = h->precision == 64 ? 8 : 4;
  HopperOut o; o.y = saa ? y : nullptr;
  o.g = saa ? (void *)((char *)g + (size_t)(1 + h->M_local) * es) : g;
  if (int rc = hopper_dispatch(h, *pt, o, st)) return rc;
  if (!saa) return SAA_OK;
  return h->precision == 64 ? launch_hopper_head<double>(h, *pt, y, g, st) : launch_hopper_head<float>(h, *pt, y, g, st);
}

int saa_hopper_g_jac(saa_handle *h, const saa_hopper_point *pt, const double *y, void *g, void *jac,
                     void *stream) {
  if (int rc = hopper_check(h, pt)) return rc;
  if (!g || !jac) return fail(h, SAA_ERR_ARG, "NULL argument");
  const bool saa = h->method == SAA_METHOD_SAA;
  if (saa && !y) return fail(h, SAA_ERR_ARG, "method saa needs the risk variables y_dev");
  SAA_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t es = h->precision == 64 ? 8 : 4;
  HopperOut o; o.y = saa ? y : nullptr; o.jac = jac;
  o.g = saa ? (void *)((char *)g + (size_t)(1 + h->M_local) * es) : g;
  if (int rc = hopper_dispatch(h, *pt, o, st)) return rc;
  if (!saa) return SAA_OK;
  return h->precision == 64 ? launch_hopper_head<double>(h, *pt, y, g, st) : launch_hopper_head<float>(h, *pt, y, g, st);
}

int saa_hopper_jac(saa_handle *h, const saa_hopper_point *pt, void *jac, void *stream) {
  if (int rc = hopper_check(h, pt)) return rc;
  if (!jac) return fail(h, SAA_ERR_ARG, "NULL argument");
  SAA_CUDA(h, cudaSetDevice(h->device));
  HopperOut o; o.jac = jac;
  return hopper_dispatch(h, *pt, o, (cudaStream_t)stream);
}

int saa_hopper_hess(saa_handle *h, const saa_hopper_point *pt, const double *lambda, double *hess, void *stream) {
  if (int rc = hopper_check(h, pt)) return rc;
  if (!lambda || !hess) return fail(h, SAA_ERR_ARG, "NULL argument");
  SAA_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  int rc = ensure_scratch(h, 1);
  if (rc) return rc;
  if (!h->d_hopper_geo) SAA_CUDA(h, cudaMalloc(&h->d_hopper_geo, (6 * 32 + 64) * sizeof(double)));
  const int n_c = pt->n_c;
  double geo[6 * 32];
  for (int c = 0; c < n_c; ++c) {
    geo[c] = pt->fz[c];
    geo[n_c + c] = 1.0;
    geo[2 * n_c + c] = pt->x3[c] * std::cos(pt->x2[c]);
    geo[3 * n_c + c] = std::sin(pt->x2[c]);
    geo[4 * n_c + c] = -pt->x3[c] * std::sin(pt->x2[c]);     // d2 p / d x2^2
    geo[5 * n_c + c] = std::cos(pt->x2[c]);                  // d2 p / d x2 d x3
  }
  SAA_CUDA(h, cudaMemcpyAsync(h->d_hopper_geo, geo, 6 * n_c * sizeof(double), cudaMemcpyHostToDevice, st));
  double *sums = h->d_hopper_geo + 6 * 32;                   // [n_c][2]
  HopperOut o; o.lambda = lambda; o.hess_sums = sums;
  rc = hopper_dispatch(h, *pt, o, st);
  if (rc) return rc;
  hopper_hess_blocks_kernel<<<1, 32, 0, st>>>(sums, n_c, h->d_hopper_geo, hess);
  SAA_CUDA(h, cudaGetLastError());
  return SAA_OK;
}

int saa_hopper_cvar_terms(saa_handle *h, const saa_hopper_point *pt, double sat_tol, void *Z, double *out3,
                          void *stream) {
  if (int rc = hopper_check(h, pt)) return rc;
  if (!out3) return fail(h, SAA_ERR_ARG, "NULL argument");
  SAA_CUDA(h, cudaSetDevice(h->device));
  HopperOut o; o.Z = Z; o.out3 = out3; o.sat_tol = sat_tol;
  return hopper_dispatch(h, *pt, o, (cudaStream_t)stream);
}

}  // extern "C"
