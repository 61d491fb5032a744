// Host side of the device-resident ADMM (qp_kernels.cuh): column tables, launches, C ABI.
// Included by saa_b200.cu inside its anonymous namespace / extern "C" sections (see the markers).
#pragma once

namespace {

struct QpPlan {
  i64 M_out = -1;                   // the layout the tables were built for
  int nact = 0, nnzJ = 0, npairs = 0;
  std::vector<QpCol> cols;
  int2 *d_pairs = nullptr;
  std::vector<int> active;          // QP column of every active column
  std::vector<int2> pairs;
};

int qp_plan(saa_handle *h, QpPlan **out) {
  if (h->problem == SAA_HOPPER) return fail(h, SAA_ERR_ARG, "the hopper has no QP");
  if (h->method != SAA_METHOD_SAA) return fail(h, SAA_ERR_ARG, "the device QP solves the CVaR ('saa') program");
  if (h->precision != 64) return fail(h, SAA_ERR_ARG, "the device QP reads FP64 values");
  if (!h->qp) h->qp = new QpPlan();
  QpPlan *p = static_cast<QpPlan *>(h->qp);
  const Layout &L = h->lay;
  if (L.relaxed_pattern) return fail(h, SAA_ERR_ARG, "relaxed car pattern (scp_iter 0) has no sample rows");
  if (p->M_out != L.M) {
    std::vector<QpCol> cols;
    int off = 0;
    for (int c = 0; c < L.nu; ++c) {
      const int len = L.run_len(c);
      if (len == 0) continue;
      const int j = c / L.n_u;
      cols.push_back(QpCol{L.run_start(c), c, j, len, L.S - 1 - j, off});
      off += len;
    }
    if ((int)cols.size() > kQpMaxCols || L.R > 32 * kQpMaxQ) return fail(h, SAA_ERR_ARG, "horizon too long for the device QP");
    p->nact = (int)cols.size(); p->nnzJ = off;
    p->active.clear();
    for (auto &c : cols) p->active.push_back(c.c);
    p->pairs.clear();
    const int nb = p->nact + 2;
    for (int a = 0; a < nb; ++a) for (int b = a; b < nb; ++b) p->pairs.push_back(make_int2(a, b));
    p->npairs = (int)p->pairs.size();
    // order the pairs by cost (common rows) so that a lane-strided assignment is balanced
    std::stable_sort(p->pairs.begin(), p->pairs.end(), [&](const int2 &x, const int2 &y) {
      auto cost = [&](const int2 &q) { return q.y < p->nact ? L.S - 1 - std::max(cols[q.x].j, cols[q.y].j) : 0; };
      return cost(x) > cost(y);
    });
    if (!p->d_pairs) SAA_CUDA(h, cudaMalloc(&p->d_pairs, sizeof(int2) * (kQpMaxCols + 2) * (kQpMaxCols + 3) / 2));
    p->cols = cols;
    SAA_CUDA(h, cudaMemcpy(p->d_pairs, p->pairs.data(), sizeof(int2) * p->pairs.size(), cudaMemcpyHostToDevice));
    p->M_out = L.M;
  }
  *out = p;
  return SAA_OK;
}

void qp_free(saa_handle *h) {
  if (!h->qp) return;
  QpPlan *p = static_cast<QpPlan *>(h->qp);
  cudaFree(p->d_pairs);
  delete p;
  h->qp = nullptr;
}

enum { QP_SCALE = 0, QP_GRAM = 1, QP_PASS = 2, QP_CHECK = 3 };
constexpr int kQpWarps = 4;

// tuning switches of the ADMM pass (defaults = the measured best): staging buffers per warp, min blocks per SM
int qp_nbuf() { static const int v = std::getenv("SAA_QP_NBUF") ? std::atoi(std::getenv("SAA_QP_NBUF")) : 1; return v == 2 ? 2 : 1; }
int qp_minb() { static const int v = std::getenv("SAA_QP_MINB") ? std::atoi(std::getenv("SAA_QP_MINB")) : 2; return v; }

int qp_plen(const saa_handle *h, const QpPlan *p, int kind) {
  const int nu = h->lay.nu, nb = p->nact + 2;
  switch (kind) {
    case QP_SCALE: return nu + 3;
    case QP_GRAM: return p->npairs + 2 * nb + 1;
    case QP_PASS: return nu + 4;
    default: return nu + 6;
  }
}
size_t qp_smem(const saa_handle *h, const QpPlan *p, int kind) {
  const int plen = qp_plen(h, p, kind);
  const int per_warp = kind == QP_GRAM ? p->nnzJ + 2 * h->lay.R + p->nact + 2
                     : kind == QP_PASS ? qp_nbuf() * (p->nnzJ + 5 * h->lay.R + 8) + h->lay.R : p->nnzJ + h->lay.R;
  return sizeof(QpShared) + sizeof(double) * (size_t)kQpWarps * (plen + per_warp);
}

template <typename K>
int qp_grid(saa_handle *h, K kernel, size_t smem, int *blocks) {
  SAA_CUDA(h, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  SAA_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kQpWarps * 32, smem));
  if (per_sm < 1) return fail(h, SAA_ERR_CUDA, "device QP kernel does not fit an SM");
  const i64 need = (h->M_local + kQpWarps - 1) / kQpWarps;
  *blocks = (int)std::max<i64>(1, std::min<i64>(need, (i64)h->n_sms * per_sm));
  return SAA_OK;
}

// the compile-time-horizon kernels apply when the layout is the reference's (S = 20, two active axes per step)
bool qp_fixed(const saa_handle *h, const QpPlan *p) {
  return h->lay.S == kS && p->nact == 2 * (kS - 1) && (h->lay.blk == 3 || h->lay.blk == 1) && !h->qp_generic;
}

template <typename F>
int qp_pass_dispatch(const saa_handle *h, const QpPlan *p, F &&f) {
  const int mb = qp_minb();
  if (!qp_fixed(h, p)) return f(qp_admm_pass_kernel<0, 0, 2>);
  if (h->lay.blk == 3) return mb >= 4 ? f(qp_admm_pass_kernel<kS, 3, 4>) : mb == 3 ? f(qp_admm_pass_kernel<kS, 3, 3>) : f(qp_admm_pass_kernel<kS, 3, 2>);
  return mb >= 4 ? f(qp_admm_pass_kernel<kS, 1, 4>) : mb == 3 ? f(qp_admm_pass_kernel<kS, 1, 3>) : f(qp_admm_pass_kernel<kS, 1, 2>);
}

int qp_blocks(saa_handle *h, const QpPlan *p, int kind, int *blocks) {
  const size_t smem = qp_smem(h, p, kind);
  switch (kind) {
    case QP_SCALE: return qp_grid(h, qp_scale_pass_kernel, smem, blocks);
    case QP_GRAM: return qp_grid(h, qp_gram_pass_kernel, smem, blocks);
    case QP_PASS:
      return qp_pass_dispatch(h, p, [&](auto kernel) { return qp_grid(h, kernel, smem, blocks); });
    default:
      if (qp_fixed(h, p)) return h->lay.blk == 3 ? qp_grid(h, qp_check_pass_kernel<kS, 3>, smem, blocks) : qp_grid(h, qp_check_pass_kernel<kS, 1>, smem, blocks);
      return qp_grid(h, qp_check_pass_kernel<0, 0>, smem, blocks);
  }
}

QpArgs qp_args(const saa_handle *h, const QpPlan *p, const double *Ax, const double *l, const double *u,
               const double *Dw, const saa_qp_sample_state *st, int kind, double *partials) {
  const Layout &L = h->lay;
  QpArgs A{};
  A.Ax = Ax; A.l = l; A.u = u;
  A.M_local = h->M_local; A.first_out = h->first_out;
  A.row_y0 = L.row_y0; A.row_s0 = L.row_s0; A.ycol0 = L.ycol0; A.slackcol = L.slackcol; A.tcol = L.tcol;
  A.R = L.R; A.S = L.S; A.blk = L.blk; A.nu = L.nu; A.nact = p->nact; A.nnzJ = p->nnzJ;
  for (int a = 0; a < p->nact; ++a) A.cols[a] = p->cols[a];
  A.Dw = Dw;
  A.st = QpSampleState{st->Dy, st->Ey, st->Es, st->xy, st->rloc, st->zy, st->ly, st->zs, st->ls};
  A.partials = partials; A.plen = qp_plen(h, p, kind);
  A.npairs = p->npairs; A.pairs = p->d_pairs;
  return A;
}

QpDense qp_dense_layout(const saa_handle *h) {
  QpDense D{};
  const Layout &L = h->lay;
  D.nu = L.nu; D.nf = L.n_fin; D.nw = L.nu + 2; D.ng = L.n_fin + 2 + L.nu;
  int o = 0;
  auto take = [&](int n) { int r = o; o += n; return r; };
  D.Fs = take(D.nf * D.nu); D.ctS = take(D.nu); D.slS = take(1); D.vcw = take(2);
  D.rho = take(D.ng); D.lo = take(D.ng); D.hi = take(D.ng); D.z = take(D.ng); D.lam = take(D.ng);
  D.Sinv = take(D.nw * D.nw); D.h = take(D.nw); D.pw = take(D.nw); D.scal = take(4); D.qw = take(D.nw);
  D.xw = take(D.nw); D.xt = take(D.nw + 1); D.total = o;
  return D;
}

}  // namespace

extern "C" {

int saa_qp_layout(saa_handle *h, int64_t *out, int64_t cap) {
  if (!h || !out) return fail(h, SAA_ERR_ARG, "NULL argument");
  QpPlan *p;
  if (int rc = qp_plan(h, &p)) return rc;
  const Layout &L = h->lay;
  const QpDense D = qp_dense_layout(h);
  std::vector<int64_t> v = {L.n_fin, L.nu, L.R, L.S, L.blk, L.M, h->first_out, L.row_cvar, L.row_y0, L.row_s0,
                            L.row_slack, L.row_ctrl0, L.ycol0, L.slackcol, L.tcol, L.nnz, L.n_rows,
                            p->nact, p->nnzJ, p->npairs,
                            D.Fs, D.ctS, D.slS, D.vcw, D.rho, D.lo, D.hi, D.z, D.lam, D.Sinv, D.h, D.pw, D.scal, D.qw,
                            D.xw, D.xt, D.total};
  // per u column: element offset, number of final rows, the final rows (4 slots), entries per sample
  for (int c = 0; c < L.nu; ++c) {
    int rows[4] = {-1, -1, -1, -1};
    const int nf = L.fin_rows(c, rows);
    v.push_back(L.ucol[c]); v.push_back(nf);
    for (int r = 0; r < 4; ++r) v.push_back(rows[r]);
    v.push_back(L.run_len(c));
  }
  for (int a = 0; a < p->nact; ++a) v.push_back(p->active[a]);
  for (int q = 0; q < p->npairs; ++q) { v.push_back(p->pairs[q].x); v.push_back(p->pairs[q].y); }
  if ((int64_t)v.size() > cap) return fail(h, SAA_ERR_ARG, "saa_qp_layout: buffer too small (need " + std::to_string(v.size()) + ")");
  std::memcpy(out, v.data(), v.size() * sizeof(int64_t));
  return SAA_OK;
}

int saa_qp_partials(saa_handle *h, int kind, int64_t *nblocks, int64_t *plen) {
  if (!h || !nblocks || !plen) return fail(h, SAA_ERR_ARG, "NULL argument");
  if (kind < 0 || kind > 3) return fail(h, SAA_ERR_ARG, "kind must be 0..3");
  SAA_CUDA(h, cudaSetDevice(h->device));
  QpPlan *p;
  if (int rc = qp_plan(h, &p)) return rc;
  int blocks = 0;
  if (int rc = qp_blocks(h, p, kind, &blocks)) return rc;
  *nblocks = blocks; *plen = qp_plen(h, p, kind);
  return SAA_OK;
}

#define QP_COMMON(kind_)                                                                   \
  if (!h || !Ax || !l || !u || !Dw || !st || !partials) return fail(h, SAA_ERR_ARG, "NULL argument"); \
  SAA_CUDA(h, cudaSetDevice(h->device));                                                   \
  QpPlan *p;                                                                               \
  if (int rc = qp_plan(h, &p)) return rc;                                                  \
  int blocks = 0;                                                                          \
  if (int rc = qp_blocks(h, p, kind_, &blocks)) return rc;                                 \
  QpArgs A = qp_args(h, p, (const double *)Ax, (const double *)l, (const double *)u, Dw, st, kind_, partials); \
  const size_t smem = qp_smem(h, p, kind_);                                                \
  cudaStream_t s_ = (cudaStream_t)stream;

int saa_qp_scale_pass(saa_handle *h, const void *Ax, const void *l, const void *u, const double *Dw, double Ec,
                      const saa_qp_sample_state *st, double *partials, void *stream) {
  QP_COMMON(QP_SCALE)
  A.Ec = Ec;
  qp_scale_pass_kernel<<<blocks, kQpWarps * 32, smem, s_>>>(A);
  SAA_CUDA(h, cudaGetLastError());
  return SAA_OK;
}

int saa_qp_gram_pass(saa_handle *h, const void *Ax, const void *l, const void *u, const double *Dw, double Ec,
                     double rho, double sigma, const saa_qp_sample_state *st, double *partials, void *stream) {
  QP_COMMON(QP_GRAM)
  A.Ec = Ec; A.rho = rho; A.sigma = sigma;
  qp_gram_pass_kernel<<<blocks, kQpWarps * 32, smem, s_>>>(A);
  SAA_CUDA(h, cudaGetLastError());
  return SAA_OK;
}

int saa_qp_admm_pass(saa_handle *h, const void *Ax, const void *l, const void *u, const double *Dw, double Ec,
                     double rho, double sigma, double alpha, const saa_qp_sample_state *st, const double *xt,
                     int first, double *partials, void *stream) {
  QP_COMMON(QP_PASS)
  if (!xt) return fail(h, SAA_ERR_ARG, "NULL argument");
  A.Ec = Ec; A.rho = rho; A.sigma = sigma; A.alpha = alpha; A.xt = xt; A.first = first;
  A.nbuf = qp_nbuf();
  qp_pass_dispatch(h, p, [&](auto kernel) { kernel<<<blocks, kQpWarps * 32, smem, s_>>>(A); return 0; });
  SAA_CUDA(h, cudaGetLastError());
  return SAA_OK;
}

int saa_qp_check_pass(saa_handle *h, const void *Ax, const void *l, const void *u, const double *Dw, double Ec,
                      const saa_qp_sample_state *st, const double *xw_lamc, double *partials, void *stream) {
  QP_COMMON(QP_CHECK)
  if (!xw_lamc) return fail(h, SAA_ERR_ARG, "NULL argument");
  A.Ec = Ec; A.xt = xw_lamc;
  if (!qp_fixed(h, p)) qp_check_pass_kernel<0, 0><<<blocks, kQpWarps * 32, smem, s_>>>(A);
  else if (h->lay.blk == 3) qp_check_pass_kernel<kS, 3><<<blocks, kQpWarps * 32, smem, s_>>>(A);
  else qp_check_pass_kernel<kS, 1><<<blocks, kQpWarps * 32, smem, s_>>>(A);
  SAA_CUDA(h, cudaGetLastError());
  return SAA_OK;
}
#undef QP_COMMON

int saa_qp_reduce(saa_handle *h, const double *partials, int64_t nblocks, int64_t plen, int64_t n_max, double *out,
                  void *stream) {
  if (!h || !partials || !out) return fail(h, SAA_ERR_ARG, "NULL argument");
  if (nblocks < 1 || plen < 1) return fail(h, SAA_ERR_ARG, "empty reduction");
  SAA_CUDA(h, cudaSetDevice(h->device));
  qp_reduce_kernel<<<(int)((plen * 32 + 127) / 128), 128, 0, (cudaStream_t)stream>>>(partials, (int)nblocks, (int)plen, (int)n_max, out);
  SAA_CUDA(h, cudaGetLastError());
  return SAA_OK;
}

int saa_qp_dense_step(saa_handle *h, double *G, const double *red, int first, void *stream) {
  if (!h || !G || !red) return fail(h, SAA_ERR_ARG, "NULL argument");
  SAA_CUDA(h, cudaSetDevice(h->device));
  QpPlan *p;
  if (int rc = qp_plan(h, &p)) return rc;
  qp_dense_step_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(qp_dense_layout(h), G, red, first);
  SAA_CUDA(h, cudaGetLastError());
  return SAA_OK;
}

}  // extern "C"
