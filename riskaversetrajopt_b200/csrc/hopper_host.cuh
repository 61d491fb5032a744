// Host-side launchers of the hopper kernels (included by saa_b200.cu).
namespace {
template <typename T>
int launch_hopper(saa_handle *h, int n_c, const double *px, void *mu, void *dmu, const double *lambda,
                  double *hess_sums, cudaStream_t st) {
  HopperArgs<T> A{};
  A.I = (const T *)h->d_a; A.theta = (const T *)h->d_b; A.tau = (const T *)h->d_c;
  A.M = h->M_local; A.F = h->n_feat; A.n_c = n_c; A.mu_nom = (T)h->mu_nom;
  for (int c = 0; c < n_c; ++c) A.px[c] = (T)px[c];
  A.mu = (T *)mu; A.dmu = (T *)dmu; A.lambda = lambda;
  // chunk of samples staged per block iteration: features (4 arrays) within ~48 KB, and a
  // multiple of the block size in (sample, contact) pairs when one exists
  const size_t es = sizeof(T);
  int cap = (int)std::max<size_t>(1, std::min<size_t>(64, (48 * 1024) / (4 * es * (size_t)h->n_feat)));
  int chunk = cap;
  for (int c = cap; c >= std::max(1, cap / 2); --c)
    if ((c * n_c) % kHopperThreads == 0) { chunk = c; break; }
  A.chunk = chunk;
  const size_t smem = 4 * es * (size_t)chunk * h->n_feat + (lambda ? 2 * sizeof(double) * (size_t)chunk * n_c : 0);
  if (smem > 200 * 1024) return fail(h, SAA_ERR_ARG, "n_features too large for the shared-memory staging");
  const i64 nchunks = (h->M_local + chunk - 1) / chunk;
  const int blocks = (int)std::max<i64>(1, std::min<i64>(nchunks, (i64)h->n_sms * 8));
  if (lambda) {
    int rc = ensure_scratch(h, (i64)blocks * 2 * n_c);
    if (rc) return rc;
    A.w = h->d_partials;
    auto kern = hopper_friction_kernel<T, true>;
    SAA_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<blocks, kHopperThreads, smem, st>>>(A);
  } else {
    auto kern = hopper_friction_kernel<T, false>;
    SAA_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<blocks, kHopperThreads, smem, st>>>(A);
  }
  SAA_CUDA(h, cudaGetLastError());
  if (lambda) {
    hopper_reduce_kernel<<<(2 * n_c * 32 + 255) / 256, 256, 0, st>>>(h->d_partials, blocks, n_c, hess_sums);
    SAA_CUDA(h, cudaGetLastError());
  }
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
  const size_t es = esize(h);
  if (h->d_a && h->n_feat != n_features) {
    cudaFree(h->d_a); cudaFree(h->d_b); cudaFree(h->d_c);
    h->d_a = h->d_b = h->d_c = nullptr;
  }
  if (!h->d_a) {
    SAA_CUDA(h, cudaMalloc(&h->d_a, n * es));
    SAA_CUDA(h, cudaMalloc(&h->d_b, n * es));
    SAA_CUDA(h, cudaMalloc(&h->d_c, n * es));
  }
  h->n_feat = n_features; h->mu_nom = mu_nom;
  cudaStream_t st = (cudaStream_t)stream;
  const int threads = 256;
  const int blocks = (int)((n + threads - 1) / threads);
  if (h->precision == 64)
    hopper_pack_kernel<double><<<blocks, threads, 0, st>>>(I, thetas, taus, n, (double *)h->d_a, (double *)h->d_b, (double *)h->d_c);
  else
    hopper_pack_kernel<float><<<blocks, threads, 0, st>>>(I, thetas, taus, n, (float *)h->d_a, (float *)h->d_b, (float *)h->d_c);
  SAA_CUDA(h, cudaGetLastError());
  h->samples_set = true; h->params_set = true;
  return SAA_OK;
}

int saa_hopper_friction(saa_handle *h, int32_t n_c, const double *px, void *mu, void *dmu,
                        const double *lambda, double *hess_sums, void *stream) {
  if (!h || !px || !mu || !dmu) return fail(h, SAA_ERR_ARG, "NULL argument");
  if (h->problem != SAA_HOPPER) return fail(h, SAA_ERR_ARG, "handle is not a hopper problem");
  if (!h->samples_set) return fail(h, SAA_ERR_STATE, "set the friction features first");
  if (n_c < 1 || n_c > kHopperMaxContacts) return fail(h, SAA_ERR_ARG, "need 1 <= n_c <= 32 contact instants");
  if ((lambda == nullptr) != (hess_sums == nullptr))
    return fail(h, SAA_ERR_ARG, "lambda_dev and hess_sums_dev go together");
  SAA_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  return h->precision == 64 ? launch_hopper<double>(h, n_c, px, mu, dmu, lambda, hess_sums, st)
                            : launch_hopper<float>(h, n_c, px, mu, dmu, lambda, hess_sums, st);
}

}  // extern "C"
