// Closed-form CSC layout of the assembled constraint matrix (host code).
//
// The reference obtains the pattern by scanning a dense matrix
// (sp.csr_matrix(dense) then sp.vstack(format='csc'), drone/drone_risk.py:419-420,
// car/driving.py:417-418).  The pattern is in fact static; this file writes it
// down.  Within a column SciPy orders entries by row, which gives, for u column
// (j, c):  [final rows] [sample i = 0..M-1: group o = 0..blk-1: step k = j+2..S]
// [control row]  -- i.e. every sample owns a contiguous sub-run.
#pragma once
#include "saa_common.cuh"

namespace saa {

// Which final rows does u column c = j*n_u + a touch?
//   drone (final rows = x_S - x_final, 6 rows): p_a row (needs j <= S-2), v_a row.
//   car   (final rows = ego (px,py,v,phi), 4 rows): px, py (j <= S-2, both
//   controls), v (a == 0), phi (a == 1).
inline int Layout::fin_rows(int c, int rows[4]) const {
  const int j = c / n_u, a = c % n_u;
  int n = 0;
  if (problem == SAA_DRONE) {
    if (j <= S - 2) rows[n++] = a;
    rows[n++] = 3 + a;
  } else {
    if (j <= S - 2) { rows[n++] = 0; rows[n++] = 1; }
    rows[n++] = 2 + a;
  }
  return n;
}

// Sample rows (i, o, k) depend on u_{j,a} iff j <= k-2 (one step for the control
// to reach the velocity, one more to reach the position); the drone's z axis
// never enters the planar obstacle constraint.
inline int Layout::run_len(int c) const {
  if (relaxed_pattern) return 0;     // only sample 0 may keep rows: relaxed_extra()
  const int j = c / n_u, a = c % n_u;
  if (j > S - 2) return 0;
  if (problem == SAA_DRONE && a == 2) return 0;
  return blk * (S - 1 - j);
}

// Car, scp_iter == 0 (car/driving.py:411-415): rows >= n_x = 8 are multiplied by exactly 0 and
// vanish, rows < n_x survive untouched.  There are only n_fin = 4 final rows, so the survivors
// are the CVaR row and the first keep_y "-y_i" rows (saa), and -- when the risk rows run out
// before row 8 (baseline: at once; saa: M < 3) -- the first keep_s separation rows of sample 0,
// WITH their Jacobian values.  Entries of u column (j, c) among them: steps k = j+2 .. keep_s.
inline int Layout::relaxed_extra(int c) const {
  if (!relaxed_pattern) return 0;
  const int j = c / n_u;
  const int n = keep_s - (j + 1);
  return n > 0 ? n : 0;
}

inline i64 Layout::ycol_len() const {
  if (method != SAA_METHOD_SAA) return 0;
  return 2 + R;   // CVaR row, -y_i row, the sample's R rows
}

inline void Layout::build(int problem_, int method_, int S_, i64 M_, bool relaxed_pattern_) {
  problem = problem_; method = method_; S = S_; M = M_;
  relaxed_pattern = relaxed_pattern_ && problem_ == SAA_CAR;
  if (problem == SAA_DRONE) { n_u = 3; n_x = 6; n_fin = 6; blk = 3; }
  else { n_u = 2; n_x = 8; n_fin = 4; blk = 1; }
  R = blk * S;
  nu = n_u * S;
  if (method == SAA_METHOD_SAA) {
    row_cvar = n_fin; row_y0 = n_fin + 1; row_s0 = n_fin + 1 + M;
    row_slack = row_s0 + M * R; row_ctrl0 = row_slack + 1;
  } else {
    row_cvar = row_y0 = row_slack = -1; row_s0 = n_fin; row_ctrl0 = row_s0 + M * R;
  }
  n_rows = row_ctrl0 + nu;
  n_cols = nu + M + 2;
  keep_y = keep_s = 0;
  if (relaxed_pattern) {
    if (method == SAA_METHOD_SAA) keep_y = (int)(n_x - n_fin - 1 < M ? n_x - n_fin - 1 : M);
    keep_s = (int)(n_x > row_s0 ? n_x - row_s0 : 0);
    if (keep_s > S) keep_s = S;
  }
  ucol.assign(nu + 1, 0);
  int rows[4];
  for (int c = 0; c < nu; ++c)
    ucol[c + 1] = ucol[c] + fin_rows(c, rows) + M * run_len(c) + relaxed_extra(c) + 1;
  ycol0 = ucol[nu];
  if (method != SAA_METHOD_SAA) {
    slackcol = tcol = nnz = ycol0;      // y / slack / t columns are empty
    return;
  }
  if (!relaxed_pattern) {
    slackcol = ycol0 + M * ycol_len();
    tcol = slackcol + (M + 2);          // CVaR row (quirk), M rows "-y_i - slack", last row
    nnz = tcol + 1 + M * R;
  } else {
    // each y: CVaR row; y_i, i < keep_y: also row n_fin+1+i; y_0: sample 0's surviving rows
    slackcol = ycol0 + M + keep_y + keep_s;
    tcol = slackcol + 1 + keep_y;
    nnz = tcol + 1 + keep_s;
  }
}

template <typename I>
void Layout::fill(I *indptr, I *indices) const {
  int rows[4];
  for (int c = 0; c <= nu; ++c) indptr[c] = (I)ucol[c];
  if (indices == nullptr) {            // column pointers only (the row indices of 10^6+ samples are tens of GB)
    if (method != SAA_METHOD_SAA) { for (i64 i = 0; i <= M + 1; ++i) indptr[nu + 1 + i] = (I)nnz; return; }
    if (!relaxed_pattern) {
      const i64 yl = ycol_len();
      for (i64 i = 0; i < M; ++i) indptr[nu + i] = (I)(ycol0 + i * yl);
    } else {
      i64 pos = ycol0;
      for (i64 i = 0; i < M; ++i) { indptr[nu + i] = (I)pos; pos += 1 + (i < keep_y) + (i == 0 ? keep_s : 0); }
    }
    indptr[nu + M] = (I)slackcol; indptr[nu + M + 1] = (I)tcol; indptr[nu + M + 2] = (I)nnz;
    return;
  }
  for (int c = 0; c < nu; ++c) {
    I *out = indices + ucol[c];
    const int nf = fin_rows(c, rows);
    for (int r = 0; r < nf; ++r) *out++ = (I)rows[r];
    const int j = c / n_u, L = S - 1 - j, len = run_len(c);
    if (len > 0) {
#pragma omp parallel for schedule(static)
      for (i64 i = 0; i < M; ++i) {
        I *o_ = out + i * len;
        for (int o = 0; o < blk; ++o)
          for (int kk = 0; kk < L; ++kk)
            *o_++ = (I)(row_s0 + i * R + o * S + (j + 1 + kk));   // row of step k = j+2+kk
      }
      out += M * len;
    }
    for (int e = 0; e < relaxed_extra(c); ++e) *out++ = (I)(row_s0 + (j + 1 + e));   // sample 0, step j+2+e
    *out++ = (I)(row_ctrl0 + c);
  }
  if (method != SAA_METHOD_SAA) {
    for (i64 i = 0; i <= M + 1; ++i) indptr[nu + 1 + i] = (I)nnz;
    return;
  }
  if (!relaxed_pattern) {
    const i64 yl = ycol_len();
#pragma omp parallel for schedule(static)
    for (i64 i = 0; i < M; ++i) {
      indptr[nu + i] = (I)(ycol0 + i * yl);
      I *o_ = indices + ycol0 + i * yl;
      *o_++ = (I)row_cvar;
      *o_++ = (I)(row_y0 + i);
      for (int r = 0; r < R; ++r) *o_++ = (I)(row_s0 + i * R + r);
    }
    indptr[nu + M] = (I)slackcol;
    I *o_ = indices + slackcol;
    *o_++ = (I)row_cvar;
    for (i64 i = 0; i < M; ++i) *o_++ = (I)(row_y0 + i);
    *o_++ = (I)row_slack;
    indptr[nu + M + 1] = (I)tcol;
    o_ = indices + tcol;
    *o_++ = (I)row_cvar;
#pragma omp parallel for schedule(static)
    for (i64 r = 0; r < M * R; ++r) o_[r] = (I)(row_s0 + r);
    indptr[nu + M + 2] = (I)nnz;
  } else {
    i64 pos = ycol0;
    for (i64 i = 0; i < M; ++i) {
      indptr[nu + i] = (I)pos;
      indices[pos++] = (I)row_cvar;
      if (i < keep_y) indices[pos++] = (I)(row_y0 + i);
      if (i == 0) for (int e = 0; e < keep_s; ++e) indices[pos++] = (I)(row_s0 + e);
    }
    indptr[nu + M] = (I)pos;
    indices[pos++] = (I)row_cvar;
    for (i64 i = 0; i < keep_y; ++i) indices[pos++] = (I)(row_y0 + i);
    indptr[nu + M + 1] = (I)pos;
    indices[pos++] = (I)row_cvar;
    for (int e = 0; e < keep_s; ++e) indices[pos++] = (I)(row_s0 + e);
    indptr[nu + M + 2] = (I)pos;
  }
}

}  // namespace saa
