// Device-resident ADMM for the CVaR program of one SCP iteration (second option).
//
// The QP the reference hands to OSQP (drone/drone_risk.py:327-399, 425-452; car/driving.py:331-397) has
// variables (u, y_1..y_M, slack, t) and an arrow-shaped constraint matrix: nu dense u columns, one y_i
// column per sample that touches only that sample's rows and the CVaR row, and the slack / t columns.
// The OSQP iteration solves (P + sigma I + A' R A) x = rhs every step.  Here that system is never
// factorised as a sparse matrix: the CVaR row is a rank-one term (Sherman-Morrison), the y block is then
// diagonal, and the coupling to the nu + 2 dense variables is eliminated sample by sample (Schur
// complement, a (nu+2)^2 matrix).  What remains per ADMM iteration is ONE pass over the samples' Jacobian
// values (read in place from the assembled CSC array) with a (nu + 4)-double reduction, plus a one-block
// kernel for the dense variables and the 8 + nu sample-independent rows.  oracle/arrow_admm.py is the
// NumPy restatement of the same algebra (checked iterate for iterate against qp.OSQPLike).
//
// One warp owns one sample: its Jacobian values (1140 doubles for the drone, 380 for the car) are staged
// in shared memory once and used for the forward product (z~ = A x~), the multiplier / projection update
// and the transposed product (A'(rho z - lambda)) of the next iteration's right-hand side.
#pragma once
#include "saa_common.cuh"

namespace saa {

struct QpCol {          // a u column that carries sample rows
  i64 gbase;            // element offset of the column's sample run in Ax (sample 0 of the matrix)
  int c, j, len, L, off;   // QP column, control step, entries per sample (= blk * L), L = S - 1 - j, offset in the staged sample
};

constexpr int kQpMaxCols = 64;     // S <= 32: 2 * 31 active columns
constexpr int kQpMaxQ = 3;         // rows per lane: R <= 96
constexpr double kQpInf = 1e20;

struct QpSampleState {             // mirrors saa_qp_sample_state (include/saa_b200.h)
  double *Dy, *Ey, *Es, *xy, *rloc, *zy, *ly, *zs, *ls;
};

struct QpArgs {
  const double *Ax, *l, *u;
  i64 M_local, first_out;
  i64 row_y0, row_s0, ycol0, slackcol, tcol;
  int R, S, blk, nu, nact, nnzJ;
  QpCol cols[kQpMaxCols];          // by value: lives in the constant bank, read with warp-uniform loads
  const double *Dw;                // nu + 2: D of (u, slack, t)
  double Ec, rho, sigma, alpha;
  QpSampleState st;
  const double *xt;                // pass: x~_w (nu + 2) and gamma'; check: x_w (nu + 2) and lambda of the CVaR row
  int first;
  int nbuf;                        // pass: staging buffers per warp (2 = the next sample is prefetched while this one is processed)
  double *partials;                // [gridDim.x][plen]
  int plen;
  // gram
  int npairs;                      // pairs over (active columns, slack, t)
  const int2 *pairs;
};

struct QpConsts { double cvar_y, yd, ys, yr, tr; };

__device__ __forceinline__ double qp_guard(double v) { return fmin(v < 1e-4 ? 1.0 : v, 1e4); }

// per-row rho as OSQP assigns it: equality rows 10^3 rho, free rows 10^-6 (qp.py:_rho_vec)
__device__ __forceinline__ double qp_rho(double l_raw, double u_raw, double E, double rho, double &ls, double &us) {
  l_raw = l_raw != l_raw ? -kQpInf : fmax(l_raw, -kQpInf);
  u_raw = u_raw != u_raw ? kQpInf : fmin(u_raw, kQpInf);
  ls = l_raw <= -kQpInf ? -kQpInf : E * l_raw;
  us = u_raw >= kQpInf ? kQpInf : E * u_raw;
  if (fabs(us - ls) < 1e-10) return 1e3 * rho;
  if (ls <= -kQpInf && us >= kQpInf) return 1e-6;
  return rho;
}
// 1 / (that rho) without a division: rho takes three values per solve
__device__ __forceinline__ double qp_inv_rho(double rho_r, double rho, double inv_rho) {
  return rho_r == rho ? inv_rho : (rho_r == 1e-6 ? 1e6 : 1e-3 * inv_rho);
}

struct QpShared {                  // block-shared header of the dynamic shared memory
  QpCol cols[kQpMaxCols];          // copy for lane-indexed access (the uniform loops read the kernel parameters)
  double ud[kQpMaxCols];           // D_c * x_c of the active columns
  double dcol[kQpMaxCols];         // D_c
  QpConsts k;
};

__device__ __forceinline__ void qp_block_setup(const QpArgs &A, QpShared *sh, bool with_x) {
  for (int a = threadIdx.x; a < A.nact; a += blockDim.x) {
    sh->cols[a] = A.cols[a];
    const double d = A.Dw[A.cols[a].c];
    sh->dcol[a] = d;
    sh->ud[a] = with_x ? d * A.xt[A.cols[a].c] : 0.0;
  }
  if (threadIdx.x == 0) {
    const i64 yl = 2 + A.R, g0 = A.first_out;
    sh->k.cvar_y = A.Ax[A.ycol0 + g0 * yl];
    sh->k.yd = A.Ax[A.ycol0 + g0 * yl + 1];
    sh->k.yr = A.Ax[A.ycol0 + g0 * yl + 2];
    sh->k.ys = A.Ax[A.slackcol + 1 + g0];
    sh->k.tr = A.Ax[A.tcol + 1 + g0 * A.R];
  }
  __syncthreads();
}

// stage sample gi's Jacobian values: Jt[off_c + o * L_c + kk] = entry (row o*S + j+1+kk, column c)
__device__ __forceinline__ void qp_load_sample(const QpArgs &A, const QpShared *sh, double *Jt, i64 gi, int lane,
                                               bool wait = true) {
  // asynchronous global -> shared copies (LDGSTS): all of the sample's column pieces are in flight at once
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(Jt);
  for (int a = 0; a < A.nact; ++a) {
    const QpCol &c = A.cols[a];
    const double *src = A.Ax + c.gbase + gi * c.len;
    for (int e = lane; e < c.len; e += 32)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sbase + 8u * (unsigned)(c.off + e)), "l"(src + e) : "memory");
  }
  if (wait) {
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
  }
}

// sum_c J[r, c] * v[c] over the active columns for row r = o * S + kr
__device__ __forceinline__ double qp_row_dot(const QpArgs &A, const QpShared *sh, const double *Jt, const double *v,
                                             int o, int kr) {
  double s = 0.0;
  for (int a = 0; a < A.nact; ++a) {
    const QpCol &c = A.cols[a];
    if (c.j < kr) s = fma(Jt[c.off + o * c.L + (kr - 1 - c.j)], v[a], s);
  }
  return s;
}
__device__ __forceinline__ double qp_row_absmax(const QpArgs &A, const QpShared *sh, const double *Jt, const double *v,
                                                int o, int kr) {
  double s = 0.0;
  for (int a = 0; a < A.nact; ++a) {
    const QpCol &c = A.cols[a];
    if (c.j < kr) s = fmax(s, fabs(Jt[c.off + o * c.L + (kr - 1 - c.j)]) * v[a]);
  }
  return s;
}
// sum_r J[r, c] * w[r] for active column a
__device__ __forceinline__ double qp_col_dot(const QpArgs &A, const QpCol &c, const double *Jt, const double *w) {
  double s = 0.0;
  for (int o = 0; o < A.blk; ++o) {
    const double *Jc = Jt + c.off + o * c.L, *wr = w + o * A.S + c.j + 1;
    for (int kk = 0; kk < c.L; ++kk) s = fma(Jc[kk], wr[kk], s);
  }
  return s;
}
__device__ __forceinline__ double qp_col_absmax(const QpArgs &A, const QpCol &c, const double *Jt, const double *w) {
  double s = 0.0;
  for (int o = 0; o < A.blk; ++o) {
    const double *Jc = Jt + c.off + o * c.L, *wr = w + o * A.S + c.j + 1;
    for (int kk = 0; kk < c.L; ++kk) s = fmax(s, fabs(Jc[kk]) * wr[kk]);
  }
  return s;
}

// ------------------------------------------------------------------------------------------------
// Horizon known at compile time (S = 20, the reference's): the same three primitives with every offset an
// immediate.  Active column index = 2 j + a (a = control axis 0 / 1), L_j = S - 1 - j, packed offset
// off(j, a) = BLK (2 sum_{j' < j} L_j' + a L_j).  S = 0 selects the run-time versions above.
// ------------------------------------------------------------------------------------------------
template <int S> __host__ __device__ constexpr int qp_sumL(int j) { return j * (S - 1) - j * (j - 1) / 2; }
template <int S, int BLK> __host__ __device__ constexpr int qp_off(int j, int a) { return BLK * (2 * qp_sumL<S>(j) + a * (S - 1 - j)); }

template <int S, int BLK>
__device__ __forceinline__ void qp_load(const QpArgs &A, const QpShared *sh, double *Jt, i64 gi, int lane,
                                        bool wait = true) {
  if constexpr (S == 0) {
    qp_load_sample(A, sh, Jt, gi, lane, wait);
  } else {
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(Jt);
    static_for<0, S - 1>([&](auto J) {
      constexpr int j = decltype(J)::value, len = BLK * (S - 1 - j);
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const double *src = A.Ax + A.cols[2 * j + a].gbase + gi * len;
        constexpr int off0 = qp_off<S, BLK>(j, 0);
        const unsigned dst = sbase + 8u * (unsigned)(off0 + a * len);
        if (lane < len)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 8u * lane), "l"(src + lane) : "memory");
        if constexpr (len > 32) {
          if (lane + 32 < len)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 8u * (lane + 32)), "l"(src + lane + 32) : "memory");
        }
      }
    });
    if (wait) {
      asm volatile("cp.async.wait_all;" ::: "memory");
      __syncwarp();
    }
  }
}

template <int S, int BLK>
__device__ __forceinline__ double qp_row(const QpArgs &A, const QpShared *sh, const double *Jt, const double *v, int r) {
  if constexpr (S == 0) {
    return qp_row_dot(A, sh, Jt, v, r / A.S, r % A.S);
  } else {
    const int o = r / S, kr = r - o * S;
    const int t0 = o * (S - 1) + kr - 1;         // entry (j, a) of this row: off(j, a) + t0 - j (o + 1)
    double s = 0.0;
    static_for<0, S - 1>([&](auto J) {
      constexpr int j = decltype(J)::value;
      if (j < kr) {
        constexpr int off0 = qp_off<S, BLK>(j, 0), off1 = qp_off<S, BLK>(j, 1);
        const int tj = t0 - j * (o + 1);
        s = fma(Jt[off0 + tj], v[2 * j], s);
        s = fma(Jt[off1 + tj], v[2 * j + 1], s);
      }
    });
    return s;
  }
}

// transposed product for active column `a` = lane (ROUND 0) or lane + 32 (ROUND 1)
template <int S, int BLK, int ROUND>
__device__ __forceinline__ double qp_col(const QpArgs &A, const QpShared *sh, const double *Jt, const double *w, int lane) {
  const int a = lane + 32 * ROUND;
  if (a >= A.nact) return 0.0;
  if constexpr (S == 0) {
    return qp_col_dot(A, sh->cols[a], Jt, w);
  } else {
    const int j = a >> 1, L = S - 1 - j;
    const int off = BLK * (2 * (j * (S - 1) - j * (j - 1) / 2) + (a & 1) * L);
    constexpr int KR0 = ROUND == 0 ? 1 : 17;      // columns 32.. have j >= 16
    double s = 0.0;
#pragma unroll
    for (int o = 0; o < BLK; ++o) {
      const double *Jc = Jt + off + o * L - 1 - j;
#pragma unroll
      for (int kr = KR0; kr < S; ++kr)
        if (kr > j) s = fma(Jc[kr], w[o * S + kr], s);
    }
    return s;
  }
}

// block-level reduction of the per-warp accumulators scratch[warp][plen] into out[plen]: the first n_max entries
// by max, the others by sum, warps in order
__device__ __forceinline__ void qp_block_reduce(const double *scratch, int plen, int n_max, double *out, int nwarps) {
  __syncthreads();
  for (int e = threadIdx.x; e < plen; e += blockDim.x) {
    double v = scratch[e];
    for (int w = 1; w < nwarps; ++w) v = e < n_max ? fmax(v, scratch[w * plen + e]) : v + scratch[w * plen + e];
    out[e] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// ADMM pass: (if !first) finish the iteration -- y back-substitution, z~ = A x~, relaxation, projection,
// multiplier update -- then the sample part of the next right-hand side.
// partial layout: [0, nu) R_u, nu R_s, nu+1 R_t, nu+2 sigma1, nu+3 cv
// ------------------------------------------------------------------------------------------------
template <int S, int BLK, int MINB>
__global__ void __launch_bounds__(128, MINB) qp_admm_pass_kernel(QpArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  QpShared *sh = reinterpret_cast<QpShared *>(smem_raw);
  const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *scratch = reinterpret_cast<double *>(sh + 1);                  // [nwarps][plen]
  // per warp: two staging buffers -- the next sample's Jacobian values AND its state (scalings, bounds, z,
  // multipliers) are in flight (cp.async) while this one is processed: no global load is waited for -- + wv
  const int bstride = A.nnzJ + 5 * A.R + 8;
  double *Jbuf = scratch + nwarps * A.plen + warp * (A.nbuf * bstride + A.R);
  double *wv = Jbuf + A.nbuf * bstride;
  auto stage = [&](double *dst, i64 s, i64 gi) {
    qp_load<S, BLK>(A, sh, dst, gi, lane, false);
    const unsigned sb = (unsigned)__cvta_generic_to_shared(dst + A.nnzJ);
    auto cp8 = [&](unsigned d, const double *src) {
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
    };
    for (int r = lane; r < A.R; r += 32) {
      const i64 gr = A.row_s0 + gi * A.R + r;
      cp8(sb + 8u * r, A.st.Es + s * A.R + r);
      cp8(sb + 8u * (A.R + r), A.l + gr);
      cp8(sb + 8u * (2 * A.R + r), A.u + gr);
      cp8(sb + 8u * (3 * A.R + r), A.st.zs + s * A.R + r);
      cp8(sb + 8u * (4 * A.R + r), A.st.ls + s * A.R + r);
    }
    if (lane < 8) {
      const double *src = lane == 0 ? A.st.Dy + s : lane == 1 ? A.st.Ey + s : lane == 2 ? A.st.xy + s
                        : lane == 3 ? A.st.rloc + s : lane == 4 ? A.st.zy + s : lane == 5 ? A.st.ly + s
                        : lane == 6 ? A.l + A.row_y0 + gi : A.u + A.row_y0 + gi;
      cp8(sb + 8u * (5 * A.R + lane), src);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  qp_block_setup(A, sh, true);
  const QpConsts k = sh->k;
  const int nu = A.nu;
  const double Ds = A.Dw[nu], Dt = A.Dw[nu + 1];
  const double s_t = A.xt[nu], t_t = A.xt[nu + 1], gam = A.xt[nu + 2];
  const double alpha = A.alpha, sigma = A.sigma, inv_rho = 1.0 / A.rho;
  double acc0 = 0.0, acc1 = 0.0, acc_s = 0.0, acc_t = 0.0, acc_s1 = 0.0, acc_cv = 0.0;
  const i64 s_first = (i64)blockIdx.x * nwarps + warp, s_step = (i64)gridDim.x * nwarps;
  int buf = 0;
  const bool dbl = A.nbuf == 2;
  if (dbl && s_first < A.M_local) stage(Jbuf, s_first, A.first_out + s_first);
  for (i64 s = s_first; s < A.M_local; s += s_step, buf ^= (dbl ? 1 : 0)) {
    const i64 gi = A.first_out + s;
    const double *Jt = Jbuf + buf * bstride;
    const double *sbuf = Jt + A.nnzJ;            // [Es | l | u | zs | ls] (R each), then Dy Ey xy rloc zy ly l_y u_y
    if (!dbl) {
      stage(Jbuf, s, gi);
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else if (s + s_step < A.M_local) {
      stage(Jbuf + (buf ^ 1) * bstride, s + s_step, gi + s_step);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncwarp();
    const double Dy = sbuf[5 * A.R], Ey = sbuf[5 * A.R + 1];
    double Es[kQpMaxQ], rho_r[kQpMaxQ], lo[kQpMaxQ], hi[kQpMaxQ], zu[kQpMaxQ], z[kQpMaxQ], lam[kQpMaxQ];
    double sA = 0.0, sB = 0.0;
#pragma unroll
    for (int q = 0; q < kQpMaxQ; ++q) {
      const int r = lane + 32 * q;
      Es[q] = rho_r[q] = zu[q] = z[q] = lam[q] = 0.0; lo[q] = hi[q] = 0.0;
      if (r < A.R) {
        Es[q] = sbuf[r];
        rho_r[q] = qp_rho(sbuf[A.R + r], sbuf[2 * A.R + r], Es[q], A.rho, lo[q], hi[q]);
        z[q] = sbuf[3 * A.R + r];
        lam[q] = sbuf[4 * A.R + r];
        const double yrS = Es[q] * k.yr * Dy;
        sA += rho_r[q] * yrS * yrS;
        if (!A.first) {
          zu[q] = Es[q] * qp_row<S, BLK>(A, sh, Jt, sh->ud, r);
          sB += rho_r[q] * yrS * (zu[q] + Es[q] * k.tr * Dt * t_t);
        }
      }
    }
    sA = sum32(sA); sB = sum32(sB);
    double ylo, yhi;
    const double ry = qp_rho(sbuf[5 * A.R + 6], sbuf[5 * A.R + 7], Ey, A.rho, ylo, yhi);
    const double ydS = Ey * k.yd * Dy, ysS = Ey * k.ys * Ds;
    const double a_i = sigma + ry * ydS * ydS + sA, inv_a = 1.0 / a_i;
    const double e_i = A.Ec * k.cvar_y * Dy;
    double xy = sbuf[5 * A.R + 2], zy = sbuf[5 * A.R + 4], ly = sbuf[5 * A.R + 5];
    if (!A.first) {
      const double bx = sB + ry * ydS * ysS * s_t;
      const double yt = (sbuf[5 * A.R + 3] + e_i * gam - bx) * inv_a;
      acc_cv += e_i * yt;
      xy = alpha * yt + (1.0 - alpha) * xy;
      {
        const double zt = ydS * yt + ysS * s_t;
        const double zr = alpha * zt + (1.0 - alpha) * zy;
        const double zn = fmin(fmax(fma(ly, qp_inv_rho(ry, A.rho, inv_rho), zr), ylo), yhi);
        ly += ry * (zr - zn);
        zy = zn;
      }
#pragma unroll
      for (int q = 0; q < kQpMaxQ; ++q) {
        const int r = lane + 32 * q;
        if (r < A.R) {
          const double zt = zu[q] + Es[q] * (k.yr * Dy * yt + k.tr * Dt * t_t);
          const double zr = alpha * zt + (1.0 - alpha) * z[q];
          const double zn = fmin(fmax(fma(lam[q], qp_inv_rho(rho_r[q], A.rho, inv_rho), zr), lo[q]), hi[q]);
          lam[q] += rho_r[q] * (zr - zn);
          z[q] = zn;
          A.st.zs[s * A.R + r] = zn;
          A.st.ls[s * A.R + r] = lam[q];
        }
      }
      if (lane == 0) { A.st.xy[s] = xy; A.st.zy[s] = zy; A.st.ly[s] = ly; }
    }
    // sample part of r = sigma x - q + A'(rho z - lambda)
    const double wy = ry * zy - ly;
    double sw = 0.0;
    double w[kQpMaxQ];
#pragma unroll
    for (int q = 0; q < kQpMaxQ; ++q) {
      w[q] = rho_r[q] * z[q] - lam[q];
      sw += Es[q] * k.yr * Dy * w[q];
    }
    sw = sum32(sw);
    const double rloc = sigma * xy + ydS * wy + sw;
    const double f = rloc * inv_a;
    double st_ = 0.0;
#pragma unroll
    for (int q = 0; q < kQpMaxQ; ++q) {
      const int r = lane + 32 * q;
      if (r < A.R) {
        const double v = w[q] - rho_r[q] * Es[q] * k.yr * Dy * f;
        wv[r] = Es[q] * v;
        st_ += Es[q] * k.tr * Dt * v;
      }
    }
    acc_t += sum32(st_);
    acc_s += ysS * (wy - ry * ydS * f);
    acc_s1 += e_i * f;
    if (lane == 0) A.st.rloc[s] = rloc;
    __syncwarp();
    if (lane < A.nact) acc0 += sh->dcol[lane] * qp_col<S, BLK, 0>(A, sh, Jt, wv, lane);
    if (lane + 32 < A.nact) acc1 += sh->dcol[lane + 32] * qp_col<S, BLK, 1>(A, sh, Jt, wv, lane);
    __syncwarp();
  }
  double *mine = scratch + warp * A.plen;
  for (int e = lane; e < A.plen; e += 32) mine[e] = 0.0;
  __syncwarp();
  if (lane < A.nact) mine[sh->cols[lane].c] = acc0;
  if (lane + 32 < A.nact) mine[sh->cols[lane + 32].c] = acc1;
  if (lane == 0) { mine[nu] = acc_s; mine[nu + 1] = acc_t; mine[nu + 2] = acc_s1; mine[nu + 3] = acc_cv; }
  qp_block_reduce(scratch, A.plen, 0, A.partials + (i64)blockIdx.x * A.plen, nwarps);
}

// ------------------------------------------------------------------------------------------------
// Residual pieces of the sample rows (OSQP's termination test, qp.py:_residuals).
// partial layout: [0] max |(Ax - z)/E|, [1] max |Ax/E|, [2] max |z/E|, [3] max |(A'lam)_y / D_y|,
//                 [4, 4+nu) (A'lam)_u, 4+nu (A'lam)_s, 5+nu (A'lam)_t      (sample rows' share)
// A.xt = (x_w (nu+2), lambda of the CVaR row)
// ------------------------------------------------------------------------------------------------
template <int S, int BLK>
__global__ void __launch_bounds__(128) qp_check_pass_kernel(QpArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  QpShared *sh = reinterpret_cast<QpShared *>(smem_raw);
  const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *scratch = reinterpret_cast<double *>(sh + 1);
  double *Jt = scratch + nwarps * A.plen + warp * (A.nnzJ + A.R);
  double *wv = Jt + A.nnzJ;
  qp_block_setup(A, sh, true);
  const QpConsts k = sh->k;
  const int nu = A.nu;
  const double Ds = A.Dw[nu], Dt = A.Dw[nu + 1];
  const double s_x = A.xt[nu], t_x = A.xt[nu + 1], lam_c = A.xt[nu + 2];
  double acc0 = 0.0, acc1 = 0.0, acc_s = 0.0, acc_t = 0.0, m_rp = 0.0, m_ax = 0.0, m_z = 0.0, m_gy = 0.0;
  for (i64 s = (i64)blockIdx.x * nwarps + warp; s < A.M_local; s += (i64)gridDim.x * nwarps) {
    const i64 gi = A.first_out + s;
    qp_load<S, BLK>(A, sh, Jt, gi, lane);
    const double Dy = A.st.Dy[s], Ey = A.st.Ey[s], xy = A.st.xy[s];
    double gy = 0.0, gt = 0.0;
#pragma unroll
    for (int q = 0; q < kQpMaxQ; ++q) {
      const int r = lane + 32 * q;
      if (r < A.R) {
        const double Es = A.st.Es[s * A.R + r], z = A.st.zs[s * A.R + r], lam = A.st.ls[s * A.R + r];
        const double ax = Es * (qp_row<S, BLK>(A, sh, Jt, sh->ud, r) + k.yr * Dy * xy + k.tr * Dt * t_x);
        m_rp = fmax(m_rp, fabs(ax - z) / Es);
        m_ax = fmax(m_ax, fabs(ax) / Es);
        m_z = fmax(m_z, fabs(z) / Es);
        wv[r] = Es * lam;
        gy += Es * k.yr * Dy * lam;
        gt += Es * k.tr * Dt * lam;
      }
    }
    gy = sum32(gy); gt = sum32(gt);
    {
      const double zy = A.st.zy[s], ly = A.st.ly[s];
      const double ydS = Ey * k.yd * Dy, ysS = Ey * k.ys * Ds;
      const double ax = ydS * xy + ysS * s_x;
      m_rp = fmax(m_rp, fabs(ax - zy) / Ey);
      m_ax = fmax(m_ax, fabs(ax) / Ey);
      m_z = fmax(m_z, fabs(zy) / Ey);
      gy += A.Ec * k.cvar_y * Dy * lam_c + ydS * ly;
      m_gy = fmax(m_gy, fabs(gy) / Dy);
      acc_s += ysS * ly;
      acc_t += gt;
    }
    __syncwarp();
    if (lane < A.nact) acc0 += sh->dcol[lane] * qp_col<S, BLK, 0>(A, sh, Jt, wv, lane);
    if (lane + 32 < A.nact) acc1 += sh->dcol[lane + 32] * qp_col<S, BLK, 1>(A, sh, Jt, wv, lane);
    __syncwarp();
  }
  m_rp = max32(m_rp); m_ax = max32(m_ax); m_z = max32(m_z);
  double *mine = scratch + warp * A.plen;
  for (int e = lane; e < A.plen; e += 32) mine[e] = 0.0;
  __syncwarp();
  if (lane < A.nact) mine[4 + sh->cols[lane].c] = acc0;
  if (lane + 32 < A.nact) mine[4 + sh->cols[lane + 32].c] = acc1;
  if (lane == 0) {
    mine[0] = m_rp; mine[1] = m_ax; mine[2] = m_z; mine[3] = m_gy;
    mine[4 + nu] = acc_s; mine[5 + nu] = acc_t;
  }
  qp_block_reduce(scratch, A.plen, 4, A.partials + (i64)blockIdx.x * A.plen, nwarps);
}

// ------------------------------------------------------------------------------------------------
// One Ruiz equilibration sweep over the samples (qp.py:_scale): norms with the OLD scalings, then the
// sample-local scalings (rows of the sample, -y_i row, y_i column) are updated in place.
// partial layout (all max): [0, nu) column norms of the u columns (sample rows' share), nu: slack column,
//                 nu+1: t column, nu+2: max D_y (old; the CVaR row's norm needs it)
// A.xt unused.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) qp_scale_pass_kernel(QpArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  QpShared *sh = reinterpret_cast<QpShared *>(smem_raw);
  const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *scratch = reinterpret_cast<double *>(sh + 1);
  double *Jt = scratch + nwarps * A.plen + warp * (A.nnzJ + A.R);
  double *wv = Jt + A.nnzJ;
  qp_block_setup(A, sh, false);
  const QpConsts k = sh->k;
  const int nu = A.nu;
  const double Ds = A.Dw[nu], Dt = A.Dw[nu + 1];
  double acc0 = 0.0, acc1 = 0.0, m_s = 0.0, m_t = 0.0, m_dy = 0.0;
  for (i64 s = (i64)blockIdx.x * nwarps + warp; s < A.M_local; s += (i64)gridDim.x * nwarps) {
    const i64 gi = A.first_out + s;
    qp_load_sample(A, sh, Jt, gi, lane);
    const double Dy = A.st.Dy[s], Ey = A.st.Ey[s];
    double Es[kQpMaxQ], esmax = 0.0;
#pragma unroll
    for (int q = 0; q < kQpMaxQ; ++q) {
      const int r = lane + 32 * q;
      Es[q] = 0.0;
      if (r < A.R) {
        Es[q] = A.st.Es[s * A.R + r];
        wv[r] = Es[q];
        esmax = fmax(esmax, Es[q]);
      }
    }
    esmax = max32(esmax);
    __syncwarp();
    // column norms with the old scalings
    if (lane < A.nact) acc0 = fmax(acc0, sh->dcol[lane] * qp_col_absmax(A, sh->cols[lane], Jt, wv));
    if (lane + 32 < A.nact) acc1 = fmax(acc1, sh->dcol[lane + 32] * qp_col_absmax(A, sh->cols[lane + 32], Jt, wv));
    m_s = fmax(m_s, fabs(k.ys) * Ey * Ds);
    m_t = fmax(m_t, fabs(k.tr) * esmax * Dt);
    m_dy = fmax(m_dy, Dy);
    const double coly = fmax(fmax(fabs(k.cvar_y) * A.Ec * Dy, fabs(k.yd) * Ey * Dy), fabs(k.yr) * esmax * Dy);
    const double rowy = fmax(fabs(k.yd) * Ey * Dy, fabs(k.ys) * Ey * Ds);
    // row norms, then the local updates
#pragma unroll
    for (int q = 0; q < kQpMaxQ; ++q) {
      const int r = lane + 32 * q;
      if (r < A.R) {
        const double rown = fmax(fmax(Es[q] * qp_row_absmax(A, sh, Jt, sh->dcol, r / A.S, r % A.S),
                                      fabs(k.yr) * Es[q] * Dy), fabs(k.tr) * Es[q] * Dt);
        A.st.Es[s * A.R + r] = Es[q] / sqrt(qp_guard(rown));
      }
    }
    if (lane == 0) {
      A.st.Ey[s] = Ey / sqrt(qp_guard(rowy));
      A.st.Dy[s] = Dy / sqrt(qp_guard(coly));
    }
    __syncwarp();
  }
  double *mine = scratch + warp * A.plen;
  for (int e = lane; e < A.plen; e += 32) mine[e] = 0.0;
  __syncwarp();
  if (lane < A.nact) mine[sh->cols[lane].c] = acc0;
  if (lane + 32 < A.nact) mine[sh->cols[lane + 32].c] = acc1;
  if (lane == 0) { mine[nu] = m_s; mine[nu + 1] = m_t; mine[nu + 2] = m_dy; }
  qp_block_reduce(scratch, A.plen, A.plen, A.partials + (i64)blockIdx.x * A.plen, nwarps);
}

// ------------------------------------------------------------------------------------------------
// Sample share of the Schur complement S = K_ww - sum_i b_i b_i' / a_i and of the Sherman-Morrison
// vectors (oracle/arrow_admm.py:_factor).  Pair p = (alpha <= beta) over ("active u column", slack = nact,
// t = nact + 1), nb = nact + 2.  Partial layout (all sums):
//   [0, npairs)                      sum_i (u,u products of the sample rows) - b_i[alpha] b_i[beta] / a_i
//   [npairs, npairs + nb)            h = sum_i b_i e_i / a_i
//   [npairs + nb]                    eps_c = sum_i e_i^2 / a_i
//   [npairs + nb + 1, npairs + 2 nb + 1)   row products that are not (u,u): (a, t) for a < nact, (s, s) at nact,
//                                    (t, t) at nact + 1
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) qp_gram_pass_kernel(QpArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  QpShared *sh = reinterpret_cast<QpShared *>(smem_raw);
  const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *scratch = reinterpret_cast<double *>(sh + 1);                   // [nwarps][plen]: accumulated in place
  const int nb = A.nact + 2;
  double *Jt = scratch + nwarps * A.plen + warp * (A.nnzJ + 2 * A.R + nb);
  double *wv = Jt + A.nnzJ;               // rho_r Es_r^2
  double *w2 = wv + A.R;                  // rho_r Es_r^2 yr Dy   (for b_u)
  double *bv = w2 + A.R;                  // b_i over (active columns, slack, t)
  qp_block_setup(A, sh, false);
  const QpConsts k = sh->k;
  const int nu = A.nu;
  const double Ds = A.Dw[nu], Dt = A.Dw[nu + 1];
  double *mine = scratch + warp * A.plen;
  for (int e = lane; e < A.plen; e += 32) mine[e] = 0.0;
  __syncwarp();
  for (i64 s = (i64)blockIdx.x * nwarps + warp; s < A.M_local; s += (i64)gridDim.x * nwarps) {
    const i64 gi = A.first_out + s;
    qp_load_sample(A, sh, Jt, gi, lane);
    const double Dy = A.st.Dy[s], Ey = A.st.Ey[s];
    double sA = 0.0, sT = 0.0;
#pragma unroll
    for (int q = 0; q < kQpMaxQ; ++q) {
      const int r = lane + 32 * q;
      if (r < A.R) {
        const i64 gr = A.row_s0 + gi * A.R + r;
        const double Es = A.st.Es[s * A.R + r];
        double lo, hi;
        const double rr = qp_rho(A.l[gr], A.u[gr], Es, A.rho, lo, hi);
        const double wgt = rr * Es * Es;
        wv[r] = wgt;
        w2[r] = wgt * k.yr * Dy;
        sA += wgt * k.yr * Dy * k.yr * Dy;
        sT += wgt;
      }
    }
    sA = sum32(sA); sT = sum32(sT);
    double ylo, yhi;
    const i64 gy = A.row_y0 + gi;
    const double ry = qp_rho(A.l[gy], A.u[gy], Ey, A.rho, ylo, yhi);
    const double ydS = Ey * k.yd * Dy, ysS = Ey * k.ys * Ds;
    const double a_i = A.sigma + ry * ydS * ydS + sA;
    const double e_i = A.Ec * k.cvar_y * Dy;
    const double trD = k.tr * Dt;
    __syncwarp();
    // b_i and the (u, t) products
    double ut0 = 0.0, ut1 = 0.0;
    if (lane < A.nact) { bv[lane] = sh->dcol[lane] * qp_col_dot(A, sh->cols[lane], Jt, w2); ut0 = sh->dcol[lane] * trD * qp_col_dot(A, sh->cols[lane], Jt, wv); }
    if (lane + 32 < A.nact) { bv[lane + 32] = sh->dcol[lane + 32] * qp_col_dot(A, sh->cols[lane + 32], Jt, w2); ut1 = sh->dcol[lane + 32] * trD * qp_col_dot(A, sh->cols[lane + 32], Jt, wv); }
    if (lane == 0) { bv[A.nact] = ry * ydS * ysS; bv[A.nact + 1] = sT * k.yr * Dy * trD; }
    __syncwarp();
    const double inv_a = 1.0 / a_i;
    for (int p = lane; p < A.npairs; p += 32) {
      const int2 pr = A.pairs[p];
      double g = 0.0;
      if (pr.y < A.nact) {                                   // (u, u): common rows kr > max(j1, j2)
        const QpCol &c1 = sh->cols[pr.x], &c2 = sh->cols[pr.y];
        const int j0 = c1.j > c2.j ? c1.j : c2.j;
        for (int o = 0; o < A.blk; ++o) {
          const double *J1 = Jt + c1.off + o * c1.L - 1 - c1.j, *J2 = Jt + c2.off + o * c2.L - 1 - c2.j, *wr = wv + o * A.S;
          for (int kr = j0 + 1; kr < A.S; ++kr) g = fma(J1[kr] * J2[kr], wr[kr], g);
        }
        g *= sh->dcol[pr.x] * sh->dcol[pr.y];
      }
      mine[p] += g - bv[pr.x] * bv[pr.y] * inv_a;
    }
    // (u, t), (s, s), (t, t) products of the rows themselves: a separate section (the host adds them into S)
    __syncwarp();
    if (lane < A.nact) mine[A.npairs + nb + 1 + lane] += ut0;
    if (lane + 32 < A.nact) mine[A.npairs + nb + 1 + lane + 32] += ut1;
    if (lane == 0) {
      mine[A.npairs + nb + 1 + A.nact] += ry * ysS * ysS;            // (s, s)
      mine[A.npairs + nb + 1 + A.nact + 1] += sT * trD * trD;        // (t, t)
      mine[A.npairs + nb] += e_i * e_i * inv_a;                      // eps_c
    }
    for (int a = lane; a < nb; a += 32) mine[A.npairs + a] += bv[a] * e_i * inv_a;    // h
    __syncwarp();
  }
  qp_block_reduce(scratch, A.plen, 0, A.partials + (i64)blockIdx.x * A.plen, nwarps);
}

// out[e] = reduce over blocks of partials[b][e]: the first n_max entries by max, the rest by sum.  One warp per
// entry: lane l folds blocks l, l+32, ... in order, then a fixed shuffle tree -- the order depends only on nblocks
// (bitwise reproducible for a given grid).
__global__ void qp_reduce_kernel(const double *__restrict__ partials, int nblocks, int plen, int n_max,
                                 double *__restrict__ out) {
  const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (e >= plen) return;
  const bool mx = e < n_max;
  double v = mx ? -kQpInf : 0.0;
  for (int b = lane; b < nblocks; b += 32) {
    const double p = partials[(i64)b * plen + e];
    v = mx ? fmax(v, p) : v + p;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double w = __shfl_xor_sync(0xffffffffu, v, o);
    v = mx ? fmax(v, w) : v + w;
  }
  if (lane == 0) out[e] = v;
}

// ------------------------------------------------------------------------------------------------
// The dense variables w = (u, slack, t) and the sample-independent rows (final rows, CVaR row, slack row,
// control rows): one block.  G is the packed global state (offsets: QpDense, qp_host.cuh).
// ------------------------------------------------------------------------------------------------
struct QpDense {
  int nu, nf, nw, ng;      // ng = nf + 2 + nu global rows: [final | cvar | slack | control]
  int Fs, ctS, slS, vcw, rho, lo, hi, z, lam, Sinv, h, pw, scal, qw, xw, xt, total;
};

__global__ void __launch_bounds__(128) qp_dense_step_kernel(QpDense L, double *__restrict__ G, const double *__restrict__ red, int first) {
  __shared__ double xt[kQpMaxCols * 2 + 4], w[kQpMaxCols * 2 + 16], rw[kQpMaxCols * 2 + 4], x0[kQpMaxCols * 2 + 4];
  __shared__ double s_kappa;
  const int t = threadIdx.x, nu = L.nu, nf = L.nf, nw = L.nw, ng = L.ng;
  const double sigma = G[L.scal + 2], alpha = G[L.scal + 3];
  for (int c = t; c <= nw; c += blockDim.x) xt[c] = G[L.xt + c];
  __syncthreads();
  // global row g: 0..nf-1 final, nf cvar, nf+1 slack, nf+2.. control
  for (int g = t; g < ng; g += blockDim.x) {
    double z = G[L.z + g], lam = G[L.lam + g];
    const double rho = G[L.rho + g];
    if (!first) {
      double zt;
      if (g < nf) { zt = 0.0; for (int c = 0; c < nu; ++c) zt = fma(G[L.Fs + g * nu + c], xt[c], zt); }
      else if (g == nf) zt = G[L.vcw] * xt[nu] + G[L.vcw + 1] * xt[nu + 1] + red[nu + 3];
      else if (g == nf + 1) zt = G[L.slS] * xt[nu];
      else zt = G[L.ctS + (g - nf - 2)] * xt[g - nf - 2];
      const double zr = alpha * zt + (1.0 - alpha) * z;
      const double zn = fmin(fmax(zr + lam / rho, G[L.lo + g]), G[L.hi + g]);
      lam += rho * (zr - zn);
      z = zn;
      G[L.z + g] = z; G[L.lam + g] = lam;
    }
    w[g] = rho * z - lam;
  }
  if (!first) for (int c = t; c < nw; c += blockDim.x) G[L.xw + c] = alpha * xt[c] + (1.0 - alpha) * G[L.xw + c];
  __syncthreads();
  const double gam = w[nf];
  for (int c = t; c < nw; c += blockDim.x) {
    double v = sigma * G[L.xw + c] - G[L.qw + c] + red[c] - gam * G[L.h + c];
    if (c < nu) {
      for (int g = 0; g < nf; ++g) v = fma(G[L.Fs + g * nu + c], w[g], v);
      v = fma(G[L.ctS + c], w[nf + 2 + c], v);
    } else if (c == nu) v += G[L.slS] * w[nf + 1] + gam * G[L.vcw];
    else v += gam * G[L.vcw + 1];
    rw[c] = v;
  }
  __syncthreads();
  for (int c = t; c < nw; c += blockDim.x) {
    double v = 0.0;
    for (int d = 0; d < nw; ++d) v = fma(G[L.Sinv + c * nw + d], rw[d], v);
    x0[c] = v;
  }
  __syncthreads();
  if (t == 0) {
    double vx0 = G[L.vcw] * x0[nu] + G[L.vcw + 1] * x0[nu + 1] + red[nu + 2] + gam * G[L.scal];
    for (int c = 0; c < nw; ++c) vx0 -= G[L.h + c] * x0[c];
    const double rc = G[L.rho + nf];
    s_kappa = rc * vx0 / (1.0 + rc * G[L.scal + 1]);
  }
  __syncthreads();
  for (int c = t; c < nw; c += blockDim.x) G[L.xt + c] = x0[c] - s_kappa * G[L.pw + c];
  if (t == 0) G[L.xt + nw] = gam - s_kappa;
}

}  // namespace saa
