// Horizon-generic SAA kernels (drone, car): run-time S in [3, kGenSMax].
//
// The tuned kernels (drone_kernels.cuh, car_kernels.cuh) are instantiated for the reference's
// horizon S = 20 (drone_params.py:9, driving_params.py:11): their register allocation, staging
// buffers and column offsets are compile-time functions of S.  The reference's Model takes S as a
// constructor argument (drone/drone_risk.py:70-83), so any other horizon is served by the kernels
// in this file: one thread per sample, the same closed-form recursions, trajectories in local
// memory, entries stored straight to their CSC positions (uncoalesced: this path trades speed for
// generality, ~5-10x below the tuned kernels).  Same outputs, same reduction slots, same
// determinism (per-warp partial rows, fixed-order final reduction).
#pragma once
#include "saa_common.cuh"
#include "car_kernels.cuh"   // rsqrt_t, sincos_t

namespace saa {

constexpr int kGenSMax = 32;
constexpr int kGenThreads = 128;

template <typename TO> struct GenDroneArgs {
  const double *mass, *dw, *q;   // packed like the tuned path: mass[M]; dw[(k*3+a)*Mpad+s]; q[(o*2+a)*Mpad+s]
  i64 M, Mpad;
  int S;
  double us[kGenSMax * 3];
  double dt, noise_c, drag, kp, kd;
  double x0[6], xf[6], oc[3][2];
  double escale, ubscale, ubpad, ztol;
  TO *Ax; i64 M_out, first_out;
  TO *ub; i64 ub_off;            // nullptr: relaxed iteration (constant bounds)
  TO *Z;
  TO *Xs;                        // rollout-only mode: (M, S+1, 6); Ax == nullptr
  double t_risk, sat_tol;
  double *partials;              // [n_warps][N] mean sums (assemble) | [n_warps][3] CVaR terms (rollout mode)
};

// slots of the mean sums for run-time S (same order as DroneRed<S>)
struct GenDroneRed {
  int S;
  __host__ __device__ int fin_p(int a, int j) const { return a * (S - 1) + j; }
  __host__ __device__ int fin_v(int a, int j) const { return 3 * (S - 1) + a * S + j; }
  __host__ __device__ int val(int r) const { return 3 * (S - 1) + 3 * S + r; }
  __host__ __device__ int n() const { return 3 * (S - 1) + 3 * S + 6; }
};

__device__ __forceinline__ void gen_warp_add(double *row, int slot, double v, int lane) {
  v = sum32(v);
  if (lane == 0) row[slot] += v;
}

template <typename TO>
__global__ void __launch_bounds__(kGenThreads)
drone_generic_kernel(const __grid_constant__ GenDroneArgs<TO> A) {
  const int S = A.S, lane = threadIdx.x & 31;
  const GenDroneRed R{S};
  const i64 warp_id = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const bool assemble = A.Ax != nullptr;
  double *row = A.partials ? A.partials + warp_id * (assemble ? R.n() : 3) : nullptr;
  if (row) {
    for (int r = lane; r < (assemble ? R.n() : 3); r += 32) row[r] = (assemble || r < 2) ? 0.0 : -INFINITY;
    __syncwarp();
  }
  const i64 nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
  const i64 ntiles = (A.M + 31) / 32;
  double acc_excess = 0.0, acc_sat = 0.0, acc_max = -INFINITY;
  for (i64 tile = warp_id; tile < ntiles; tile += nwarps) {
    const i64 s_raw = tile * 32 + lane;
    const bool active = s_raw < A.M;
    const i64 s = active ? s_raw : A.M - 1;
    const double inv_m = 1.0 / A.mass[s];
    const double dt = A.dt, dtm = dt * inv_m, a21 = -A.kp * dtm, nz = A.noise_c * inv_m, c2 = 2.0 * A.drag;
    double P[3][kGenSMax + 1], A22[3][kGenSMax];
    double qq[3][2];
    for (int o = 0; o < 3; ++o) for (int a = 0; a < 2; ++a) qq[o][a] = A.q[(o * 2 + a) * A.Mpad + s];
    // ---- rollout of the three axes with the tangent along u (drone_risk.py:139-155) ----
    double p[3], v[3], tp[3] = {0, 0, 0}, tv[3] = {0, 0, 0};
    for (int a = 0; a < 3; ++a) { p[a] = A.x0[a]; v[a] = A.x0[3 + a]; P[a][0] = p[a]; }
    double zmax = -INFINITY;
    TO *xs = A.Xs ? A.Xs + s * (i64)(S + 1) * 6 : nullptr;
    if (xs && active) for (int a = 0; a < 3; ++a) { xs[a] = (TO)p[a]; xs[3 + a] = (TO)v[a]; }
    for (int k = 0; k < S; ++k) {
      for (int a = 0; a < 3; ++a) {
        const double absv = fabs(v[a]);
        const double a22 = 1.0 - dtm * (A.kd + c2 * absv);
        A22[a][k] = a22;
        const double u = A.us[k * 3 + a];
        const double acc = (u - A.kp * p[a] - A.kd * v[a] - A.drag * absv * v[a]) * inv_m;
        const double ntp = fma(dt, tv[a], tp[a]);
        const double ntv = fma(a22, tv[a], fma(a21, tp[a], dtm * u));
        const double np_ = fma(dt, v[a], p[a]);
        v[a] = v[a] + dt * acc + nz * A.dw[(i64)(k * 3 + a) * A.Mpad + s];
        p[a] = np_; tp[a] = ntp; tv[a] = ntv;
        P[a][k + 1] = p[a];
      }
      if (xs && active) for (int a = 0; a < 3; ++a) { xs[(k + 1) * 6 + a] = (TO)p[a]; xs[(k + 1) * 6 + 3 + a] = (TO)v[a]; }
      for (int o = 0; o < 3; ++o) {
        double wsum = 0.0, esum = 0.0;
        for (int a = 0; a < 2; ++a) {
          const double d = p[a] - A.oc[o][a];
          const double w = qq[o][a] * d * d;
          wsum += w;
          esum += fma(-2.0 * qq[o][a] * d, tp[a], w);
        }
        zmax = fmax(zmax, 1.0 - wsum);
        if (assemble && A.ub && active)
          A.ub[A.ub_off + s * (i64)(3 * S) + o * S + k] = (TO)fma(esum - 1.0, A.ubscale, -A.ubpad);
      }
    }
    const double Zi = zmax - A.ztol;
    if (A.Z && active) A.Z[s] = (TO)Zi;
    if (!assemble) {
      if (active) {
        acc_excess += fmax(Zi - A.t_risk, 0.0);
        acc_sat += (Zi <= A.sat_tol) ? 1.0 : 0.0;
        acc_max = fmax(acc_max, Zi);
      }
      continue;
    }
    // linearisation offsets of the final rows (:271)
    for (int a = 0; a < 3; ++a) {
      gen_warp_add(row, R.val(a), active ? -(p[a] - A.xf[a]) + tp[a] : 0.0, lane);
      gen_warp_add(row, R.val(3 + a), active ? -(v[a] - A.xf[3 + a]) + tv[a] : 0.0, lane);
    }
    // ---- sensitivity chains: column (j, a) holds rows k = j+2..S of the three obstacles ----
    for (int a = 0; a < 3; ++a) {
      for (int j = 0; j < S; ++j) {
        double sp = 0.0, sv = dtm;
        const int L = S - 1 - j;
        const i64 base = (9 * j + 3 * a + 2) + A.M_out * (i64)(6 * j * (S - 1) - 3 * j * (j - 1) + 3 * a * L) +
                         (A.first_out + s) * (i64)(3 * L);
        for (int k = j + 1; k < S; ++k) {
          const double nsp = fma(dt, sv, sp);
          const double nsv = fma(A22[a][k], sv, a21 * sp);
          sp = nsp; sv = nsv;                               // d(p,v)_{k+1} / du_j
          if (a < 2 && active) {
            for (int o = 0; o < 3; ++o) {
              const double q2 = -2.0 * A.escale * qq[o][a];
              const double coef = fma(q2, P[a][k + 1], -q2 * A.oc[o][a]);
              A.Ax[base + o * L + (k - j - 1)] = (TO)(coef * sp);
            }
          }
        }
        if (j < S - 1) gen_warp_add(row, R.fin_p(a, j), active ? sp : 0.0, lane);
        gen_warp_add(row, R.fin_v(a, j), active ? sv : 0.0, lane);
      }
    }
  }
  if (!assemble && row) {
    acc_excess = sum32(acc_excess); acc_sat = sum32(acc_sat); acc_max = max32(acc_max);
    if (lane == 0) { row[0] = acc_excess; row[1] = acc_sat; row[2] = acc_max; }
  }
}

// sums_out[r] = sum over warp rows (fixed order); CVaR mode: [sum, sum, max]
__global__ void gen_reduce_kernel(const double *__restrict__ partials, i64 nrows, int n, int cvar,
                                  double *__restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  double acc = (cvar && r == 2) ? -INFINITY : 0.0;
  for (i64 b = 0; b < nrows; ++b) {
    const double v = partials[b * n + r];
    acc = (cvar && r == 2) ? fmax(acc, v) : acc + v;
  }
  out[r] = acc;
}

// ------------------------------------------------------------------------------------------------
// car
// ------------------------------------------------------------------------------------------------
template <typename TO> struct GenCarArgs {
  const double *x0, *om, *dw;    // packed like the tuned path
  i64 M, Mpad;
  int S;
  double us[kGenSMax * 2];
  double ego0[4], goal[4];
  double dt, noise_c, v_des, d_min, ztol;
  TO *Ax; i64 M_out, first_out;
  TO *ub; i64 ub_off;
  TO *Z;
  TO *Xs;                        // rollout-only mode: (M, S+1, 8); Ax == nullptr
  double t_risk, sat_tol;
  double *sums;                  // assemble: M * (sample-independent final-row values), slots of GenCarRed
  double *partials;              // rollout mode: [n_warps][3]
  unsigned long long *nonfinite;
};

struct GenCarRed {
  int S;
  __host__ __device__ int px(int c, int j) const { return c * (S - 1) + j; }
  __host__ __device__ int py(int c, int j) const { return 2 * (S - 1) + c * (S - 1) + j; }
  __host__ __device__ int v(int j) const { return 4 * (S - 1) + j; }
  __host__ __device__ int phi(int j) const { return 4 * (S - 1) + S + j; }
  __host__ __device__ int val(int r) const { return 4 * (S - 1) + 2 * S + r; }
  __host__ __device__ int n() const { return 4 * (S - 1) + 2 * S + 4; }
};

struct GenCarEgo {
  double p[kGenSMax + 1][2], st[kGenSMax + 1][2];   // position, (v, phi)
  double tt[2][kGenSMax][2];
  double ucum[2][kGenSMax + 1];
  double fin[4];
};

template <typename TO>
__global__ void __launch_bounds__(kGenThreads)
car_generic_kernel(const __grid_constant__ GenCarArgs<TO> A) {
  __shared__ GenCarEgo E;
  const int S = A.S, lane = threadIdx.x & 31;
  const GenCarRed R{S};
  if (threadIdx.x == 0) {                              // ego rollout (car/driving.py:167-172): sample independent
    double px = A.ego0[0], py = A.ego0[1], v = A.ego0[2], phi = A.ego0[3], u0 = 0.0, u1 = 0.0;
    const double dt2 = A.dt * A.dt;
    for (int k = 0; k < S; ++k) {
      E.p[k][0] = px; E.p[k][1] = py; E.st[k][0] = v; E.st[k][1] = phi;
      E.ucum[0][k] = u0; E.ucum[1][k] = u1;
      double sn, cs;
      sincos(phi, &sn, &cs);
      E.tt[0][k][0] = dt2 * cs;      E.tt[0][k][1] = dt2 * sn;
      E.tt[1][k][0] = -dt2 * v * sn; E.tt[1][k][1] = dt2 * v * cs;
      px = px + A.dt * (v * cs); py = py + A.dt * (v * sn);
      v = v + A.dt * A.us[2 * k]; phi = phi + A.dt * A.us[2 * k + 1];
      u0 += A.us[2 * k]; u1 += A.us[2 * k + 1];
    }
    E.p[S][0] = px; E.p[S][1] = py; E.st[S][0] = v; E.st[S][1] = phi;
    E.ucum[0][S] = u0; E.ucum[1][S] = u1;
    E.fin[0] = px; E.fin[1] = py; E.fin[2] = v; E.fin[3] = phi;
  }
  __syncthreads();
  const bool assemble = A.Ax != nullptr || A.sums != nullptr;
  if (blockIdx.x == 0 && A.sums != nullptr) {          // sample-independent final rows (:217-221, :271, :311-313)
    const double Md = (double)A.M;
    for (int r = threadIdx.x; r < R.n(); r += blockDim.x) {
      double val = 0.0;
      if (r < R.v(0)) {
        const int comp = r >= R.py(0, 0), q = r - (comp ? R.py(0, 0) : 0), c = q / (S - 1), j = q % (S - 1);
        for (int m = j + 1; m < S; ++m) val += E.tt[c][m][comp];
      } else if (r < R.val(0)) {
        val = A.dt;
      } else {
        const int rw = r - R.val(0);
        double ju = 0.0;
        if (rw < 2) { for (int c = 0; c < 2; ++c) for (int m = 0; m < S; ++m) ju += E.tt[c][m][rw] * E.ucum[c][m]; }
        else ju = A.dt * E.ucum[rw - 2][S];
        val = -(E.fin[rw] - A.goal[rw]) + ju;
      }
      A.sums[r] = Md * val;
    }
  }
  if (A.Ax == nullptr && A.Xs == nullptr && A.Z == nullptr && A.partials == nullptr) return;
  const i64 warp_id = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const i64 nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
  const i64 ntiles = (A.M + 31) / 32;
  double acc_excess = 0.0, acc_sat = 0.0, acc_max = -INFINITY;
  bool bad = false;
  for (i64 tile = warp_id; tile < ntiles; tile += nwarps) {
    const i64 s_raw = tile * 32 + lane;
    const bool active = s_raw < A.M;
    const i64 s = active ? s_raw : A.M - 1;
    double qx = A.x0[s], qy = A.x0[A.Mpad + s], wx = A.x0[2 * A.Mpad + s], wy = A.x0[3 * A.Mpad + s];
    const double w_s = A.om[s], w_r = A.om[A.Mpad + s];
    const double dt = A.dt, wsdt = w_s * dt, dtwr = dt * w_r;
    double NX[kGenSMax + 1], NY[kGenSMax + 1], OM[kGenSMax + 1];
    double cu[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};    // tangent along u of each control: rx, ry, wx, wy
    double zmax = -INFINITY;
    TO *xs = A.Xs ? A.Xs + s * (i64)(S + 1) * 8 : nullptr;
    for (int k = 0; k <= S; ++k) {
      if (xs && active) {
        xs[k * 8] = (TO)E.p[k][0]; xs[k * 8 + 1] = (TO)E.p[k][1]; xs[k * 8 + 2] = (TO)E.st[k][0]; xs[k * 8 + 3] = (TO)E.st[k][1];
        xs[k * 8 + 4] = (TO)qx; xs[k * 8 + 5] = (TO)qy; xs[k * 8 + 6] = (TO)wx; xs[k * 8 + 7] = (TO)wy;
      }
      const double dx = E.p[k][0] - qx, dy = E.p[k][1] - qy;
      const double n2 = fma(dx, dx, dy * dy);
      const double inv_n = rsqrt(n2);
      bad |= active && (!(n2 > 0.0) || !(n2 < INFINITY));
      const double nhx = dx * inv_n, nhy = dy * inv_n, om_n = dtwr * inv_n;
      NX[k] = nhx; NY[k] = nhy; OM[k] = om_n;
      if (k >= 1) {
        const double gu = -(fma(nhx, cu[0][0], nhy * cu[0][1]) + fma(nhx, cu[1][0], nhy * cu[1][1]));
        const double g = A.d_min - n2 * inv_n;
        zmax = fmax(zmax, g);
        if (assemble && A.ub && active) A.ub[A.ub_off + s * (i64)S + (k - 1)] = (TO)(gu - g);
      }
      if (k < S) {
        for (int c = 0; c < 2; ++c) {
          const double ttx = E.tt[c][k][0], tty = E.tt[c][k][1], uc = E.ucum[c][k];
          const double d = fma(nhx, cu[c][0], nhy * cu[c][1]);
          const double nwx = fma(om_n * nhx, d, fma(-om_n, cu[c][0], fma(-wsdt, cu[c][3], cu[c][2])));
          const double nwy = fma(om_n * nhy, d, fma(-om_n, cu[c][1], fma(-wsdt, cu[c][3], cu[c][3])));
          cu[c][0] = fma(-dt, cu[c][2], fma(uc, ttx, cu[c][0]));
          cu[c][1] = fma(-dt, cu[c][3], fma(uc, tty, cu[c][1]));
          cu[c][2] = nwx; cu[c][3] = nwy;
        }
        const double sp = w_s * (A.v_des - wy);
        const double fx = fma(-w_r, nhx, sp), fy = fma(-w_r, nhy, sp);
        const double nqx = fma(dt, wx, qx), nqy = fma(dt, wy, qy);
        wx = wx + dt * fx + A.noise_c * A.dw[(i64)(2 * k) * A.Mpad + s];
        wy = wy + dt * fy + A.noise_c * A.dw[(i64)(2 * k + 1) * A.Mpad + s];
        qx = nqx; qy = nqy;
      }
    }
    const double Zi = zmax - A.ztol;
    if (A.Z && active) A.Z[s] = (TO)Zi;
    if (active) {
      acc_excess += fmax(Zi - A.t_risk, 0.0);
      acc_sat += (Zi <= A.sat_tol) ? 1.0 : 0.0;
      acc_max = fmax(acc_max, Zi);
    }
    if (A.Ax == nullptr || !active) continue;
    // ---- chains: column (j, c) holds rows k = j+2..S (car/driving.py:261-298) ----
    for (int c = 0; c < 2; ++c) {
      for (int j = 0; j < S - 1; ++j) {
        const int L = S - 1 - j;
        const i64 base = (8 * j + 4 * c + 3) + A.M_out * (i64)(2 * j * (S - 1) - j * (j - 1) + c * L) +
                         (A.first_out + s) * (i64)L;
        double rx = E.tt[c][j + 1][0], ry = E.tt[c][j + 1][1], cwx = 0.0, cwy = 0.0;   // state at k = j + 2
        for (int k = j + 2; k <= S; ++k) {
          const double d = fma(NX[k], rx, NY[k] * ry);
          A.Ax[base + (k - j - 2)] = (TO)(-d);
          if (k < S) {
            const double nwx = fma(OM[k] * NX[k], d, fma(-OM[k], rx, fma(-wsdt, cwy, cwx)));
            const double nwy = fma(OM[k] * NY[k], d, fma(-OM[k], ry, fma(-wsdt, cwy, cwy)));
            rx = fma(-dt, cwx, rx + E.tt[c][k][0]);
            ry = fma(-dt, cwy, ry + E.tt[c][k][1]);
            cwx = nwx; cwy = nwy;
          }
        }
      }
    }
  }
  if (A.partials != nullptr) {
    double *row = A.partials + warp_id * 3;
    acc_excess = sum32(acc_excess); acc_sat = sum32(acc_sat); acc_max = max32(acc_max);
    if (lane == 0) { row[0] = acc_excess; row[1] = acc_sat; row[2] = acc_max; }
  }
  if (A.nonfinite != nullptr) {
    const unsigned m = __ballot_sync(0xffffffffu, bad);
    if (lane == 0 && m) atomicAdd(A.nonfinite, (unsigned long long)__popc(m));
  }
}

}  // namespace saa
