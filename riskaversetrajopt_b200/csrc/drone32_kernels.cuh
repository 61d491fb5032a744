// Quadrotor linearize + assemble with 32-sample tiles (K1, variant selected by SAA_DRONE_TILE=32).
//
// Same arithmetic per sample as drone_assemble_kernel (drone_kernels.cuh) -- the entries are
// bitwise identical -- but a different work split: a warp owns 32 samples, LANE = SAMPLE, and each
// lane runs BOTH planar axes.  What changes:
//   * the column sub-run a warp writes is 32 samples long (768 (S-1-j) bytes instead of 384): the
//     store-pattern probe (tools/wbw3.cu, profiles/README.md) gives 5.85 TB/s for 32-sample runs
//     against 5.34 TB/s for 16-sample runs -- half as many run boundaries = half as many 128-byte
//     lines written in two pieces;
//   * no cross-lane exchange in the rollout (the obstacle rows need w_x + w_y of the same sample:
//     240 shuffles per tile in the 16-sample kernel, none here);
//   * two independent sensitivity chains (x and y) per lane instead of one.
// Price: 164 registers of trajectory state per lane (P, A22 of both axes), 29 KB of staging per
// warp -> 7 warps per SM (224 samples in flight against 192).
#pragma once
#include "drone_kernels.cuh"

namespace saa {

constexpr int kTile32 = 32;

template <typename TO, int S> struct Drone32Geom {
  __host__ __device__ static constexpr int stage_size() {       // elements: the largest column pair / the bounds
    int m = Stager<TO, 3 * S, kTile32>::SIZE / 2 + 8;            // upper bounds use column slot 0 only
    for (int j = 0; j < S - 1; ++j) {
      const int len = 3 * (S - 1 - j);
      const int vec = 16 / (int)sizeof(TO);
      const bool roww = (len % (2 * vec)) == 0;
      const int ybase = ((kTile32 * len + vec - 1) / vec + 1) * vec;
      const int sz = roww ? 2 * kTile32 * (len + vec) : 2 * ybase;
      m = sz > m ? sz : m;
    }
    return (m + 7) / 8 * 8;
  }
};

template <typename TO, int S, int WARPS> struct Drone32Smem {
  static constexpr int STAGE = Drone32Geom<TO, S>::stage_size();
  alignas(16) TO stage[WARPS][STAGE];
  double wacc[WARPS][DroneRed<S>::N];
};

template <typename T, typename TO, int S, int J>
__device__ __forceinline__ void drone32_chains(const DroneArgs<T, TO, S> &A, const DroneOut<TO> &O,
                                               const T (&Px)[S + 1], const T (&Py)[S + 1], const T (&A22x)[S],
                                               const T (&A22y)[S], const T (&q2x)[3], const T (&q2y)[3], T a21,
                                               T dtm, TO *stage, double *wacc, int lane, i64 s0, int ns,
                                               bool active) {
  using Rd = DroneRed<S>;
  if constexpr (J >= S - 1) {
    const double rv = sum32((double)(active ? dtm : T(0)));
    if (lane == 0) { wacc[Rd::FIN_V + (S - 1)] += rv; wacc[Rd::FIN_V + S + (S - 1)] += rv; }
  } else {
    using C = DroneChain<S, J>;
    using St = Stager<TO, C::LEN, kTile32>;
    static_assert(St::SIZE <= Drone32Geom<TO, S>::stage_size(), "staging buffer too small");
    // same optimisation barriers as drone_chains (keep the per-chain coefficients out of the
    // unrolled common-subexpression pool)
    T qx[3] = {q2x[0], q2x[1], q2x[2]}, qy[3] = {q2y[0], q2y[1], q2y[2]};
    T cx[3] = {-q2x[0] * A.oc[0][0], -q2x[1] * A.oc[1][0], -q2x[2] * A.oc[2][0]};
    T cy[3] = {-q2y[0] * A.oc[0][1], -q2y[1] * A.oc[1][1], -q2y[2] * A.oc[2][1]};
#pragma unroll
    for (int o = 0; o < 3; ++o) { opaque(qx[o]); opaque(qy[o]); opaque(cx[o]); opaque(cy[o]); }
    i64 sbase = s0 + O.first, mout = O.mout;
    opaque(sbase); opaque(mout);
    const i64 g0x = C::CA0 + mout * C::CB0 + sbase * C::LEN, g0y = C::CA1 + mout * C::CB1 + sbase * C::LEN;
    TO *mx = St::mine(stage, 0, lane, g0x), *my = St::mine(stage, 1, lane, g0y);
    T spx = T(0), svx = dtm, spy = T(0), svy = dtm;        // d(p,v)_{J+1}/du_J = (0, dt/m), both axes
#pragma unroll
    for (int k = J + 1; k < S; ++k) {
      const int kk = k - J - 1;
      const T nspx = fma(A.dt, svx, spx), nsvx = fma(A22x[k], svx, a21 * spx);
      const T nspy = fma(A.dt, svy, spy), nsvy = fma(A22y[k], svy, a21 * spy);
      spx = nspx; svx = nsvx; spy = nspy; svy = nsvy;
#pragma unroll
      for (int o = 0; o < 3; ++o) {
        mx[o * C::L + kk] = (TO)(fma(qx[o], Px[k + 1], cx[o]) * spx);
        my[o * C::L + kk] = (TO)(fma(qy[o], Py[k + 1], cy[o]) * spy);
      }
    }
    {
      const double rpx = sum32((double)(active ? spx : T(0))), rvx = sum32((double)(active ? svx : T(0)));
      const double rpy = sum32((double)(active ? spy : T(0))), rvy = sum32((double)(active ? svy : T(0)));
      if (lane == 0) {
        wacc[Rd::FIN_P + J] += rpx; wacc[Rd::FIN_V + J] += rvx;
        wacc[Rd::FIN_P + (S - 1) + J] += rpy; wacc[Rd::FIN_V + S + J] += rvy;
      }
    }
    __syncwarp();
    St::copy_vec(O.Ax, stage, 0, g0x, ns, lane);
    St::copy_vec(O.Ax, stage, 1, g0y, ns, lane);
    __syncwarp();
    drone32_chains<T, TO, S, J + 1>(A, O, Px, Py, A22x, A22y, q2x, q2y, a21, dtm, stage, wacc, lane, s0, ns, active);
  }
}

template <typename T, typename TO, int S, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
drone_assemble32_kernel(const __grid_constant__ DroneArgs<T, TO, S> A) {
  using Rd = DroneRed<S>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  auto &sm = *reinterpret_cast<Drone32Smem<TO, S, WARPS> *>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  TO *stage = sm.stage[warp];
  double *wacc = sm.wacc[warp];
  for (int r = lane; r < Rd::N; r += 32) wacc[r] = 0.0;
  __syncwarp();
  DroneOut<TO> O{A.Ax, A.fsp, A.M_out, A.first_out};
  {
    i64 ax = (i64)O.Ax;
    opaque(ax); opaque(O.mout); opaque(O.first);
    O.Ax = (TO *)ax;
  }
  i64 ub_base = (i64)A.ub, ub_off = A.ub_off;
  opaque(ub_base); opaque(ub_off);
  TO *const ub_ptr = (TO *)ub_base;
  const i64 ntiles = (A.M + kTile32 - 1) / kTile32;
#pragma unroll 1
  for (i64 tile = (i64)blockIdx.x * WARPS + warp; tile < ntiles; tile += (i64)gridDim.x * WARPS) {
    const i64 s0 = tile * kTile32;
    const int ns = (int)min((i64)kTile32, A.M - s0);
    const bool active = lane < ns;
    const i64 s = s0 + (active ? lane : 0);
    // ---- all inputs of the sample in flight at once ------------------------------------------
    T dwx[S], dwy[S];
#pragma unroll
    for (int k = 0; k < S; ++k) {
      dwx[k] = __ldcs(A.dw + (i64)(k * 3) * A.Mpad + s);
      dwy[k] = __ldcs(A.dw + (i64)(k * 3 + 1) * A.Mpad + s);
    }
    T q[3][2];
#pragma unroll
    for (int o = 0; o < 3; ++o) { q[o][0] = __ldcs(A.q + (o * 2) * A.Mpad + s); q[o][1] = __ldcs(A.q + (o * 2 + 1) * A.Mpad + s); }
    const T inv_m = T(1) / __ldcs(A.mass + s);
    const T dt = A.dt, dtm = dt * inv_m, a21 = -A.kp * dtm, nz = A.noise_c * inv_m, c2 = T(2) * A.drag;
    T q2x[3], q2y[3];
#pragma unroll
    for (int o = 0; o < 3; ++o) { q2x[o] = T(-2) * A.escale * q[o][0]; q2y[o] = T(-2) * A.escale * q[o][1]; }
    // ---- rollout of both planar axes + constraint values + linearisation offsets ---------------
    T Px[S + 1], Py[S + 1], A22x[S], A22y[S];
    T px = A.x0[0], vx = A.x0[3], py = A.x0[1], vy = A.x0[4], tpx = T(0), tvx = T(0), tpy = T(0), tvy = T(0);
    T zmax = -INFINITY;
    Px[0] = px; Py[0] = py;
    using StU = Stager<TO, 3 * S, kTile32>;
    const i64 gu = ub_off + s0 * (3 * S);
    TO *ubrow = StU::mine(stage, 0, lane, gu);
#pragma unroll
    for (int k = 0; k < S; ++k) {
      {
        const T absv = fabs(vx);
        const T a22 = T(1) - dtm * (A.kd + c2 * absv);
        A22x[k] = a22;
        const T u = A.us[k * 3];
        const T acc = (u - A.kp * px - A.kd * vx - A.drag * absv * vx) * inv_m;
        const T ntp = fma(dt, tvx, tpx), ntv = fma(a22, tvx, fma(a21, tpx, dtm * u));
        const T np_ = fma(dt, vx, px);
        vx = vx + dt * acc + nz * dwx[k];
        px = np_; tpx = ntp; tvx = ntv;
        Px[k + 1] = px;
      }
      {
        const T absv = fabs(vy);
        const T a22 = T(1) - dtm * (A.kd + c2 * absv);
        A22y[k] = a22;
        const T u = A.us[k * 3 + 1];
        const T acc = (u - A.kp * py - A.kd * vy - A.drag * absv * vy) * inv_m;
        const T ntp = fma(dt, tvy, tpy), ntv = fma(a22, tvy, fma(a21, tpy, dtm * u));
        const T np_ = fma(dt, vy, py);
        vy = vy + dt * acc + nz * dwy[k];
        py = np_; tpy = ntp; tvy = ntv;
        Py[k + 1] = py;
      }
#pragma unroll
      for (int o = 0; o < 3; ++o) {
        const T dx = px - A.oc[o][0], dy = py - A.oc[o][1];
        const T wx = q[o][0] * dx * dx, wy = q[o][1] * dy * dy;
        const T ex = fma(T(-2) * q[o][0] * dx, tpx, wx), ey = fma(T(-2) * q[o][1] * dy, tpy, wy);
        // same association as the 16-sample kernel: (own axis) + (other axis), x first
        zmax = fmax(zmax, T(1) - (wx + wy));
        ubrow[o * S + k] = (TO)fma((ex + ey) - T(1), A.ubscale, -A.ubpad);
      }
    }
    {
      const double rpx = sum32((double)(active ? -(px - A.xf[0]) + tpx : T(0)));
      const double rvx = sum32((double)(active ? -(vx - A.xf[3]) + tvx : T(0)));
      const double rpy = sum32((double)(active ? -(py - A.xf[1]) + tpy : T(0)));
      const double rvy = sum32((double)(active ? -(vy - A.xf[4]) + tvy : T(0)));
      if (lane == 0) { wacc[Rd::VAL] += rpx; wacc[Rd::VAL + 3] += rvx; wacc[Rd::VAL + 1] += rpy; wacc[Rd::VAL + 4] += rvy; }
    }
    if (A.Z != nullptr && active) A.Z[s] = (TO)(zmax - A.ztol);
    __syncwarp();
    if (ub_ptr != nullptr) StU::copy_vec(ub_ptr, stage, 0, gu, ns, lane);
    __syncwarp();
    drone32_chains<T, TO, S, 0>(A, O, Px, Py, A22x, A22y, q2x, q2y, a21, dtm, stage, wacc, lane, s0, ns, active);
  }
  __syncthreads();
  for (int r = threadIdx.x; r < Rd::N; r += WARPS * 32) {
    double acc = 0.0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) acc += sm.wacc[w][r];
    A.partials[(i64)blockIdx.x * Rd::N + r] = acc;
  }
}

}  // namespace saa
