// Ego car + pedestrian SAA kernels (K2 assemble, K4 rollout, K5 CVaR terms), sm_100a.
//
// What is computed (reference car/driving.py):
//   force_on_pedestrian :146-158  F = -w_r d/|d| + w_s (1.3 - x[7]) (scalar broadcast on both
//                                  components; x[7] is the pedestrian's v_y -- reference quirk, kept)
//   b, sigma            :161-184  ego unicycle (px,py,v,phi), pedestrian double integrator + F
//   rollout             :187-204  Euler-Maruyama, noise on states 6,7 only
//   constraints         :217-236  ego_S - goal ;  g_k = -(|p_ego,k - p_ped,k| - d_min), k = 1..S
//   per-sample Jacobian :261-298  jacfwd wrt us (here: closed-form forward sensitivities)
//   packing             :302-373
//
// Structure used here.  The ego states carry no noise and no pedestrian coupling,
// so the ego trajectory and d ego/du are sample independent: with
//   T_0(k) = dt^2 (cos phi_k, sin phi_k),  T_1(k) = dt^2 (-v_k sin phi_k, v_k cos phi_k)
// d p_ego,k / d u_{j,c} = sum_{m=j+1}^{k-1} T_c(m).  For the pedestrian, with
// rho = d p_ego/du - d p_ped/du and w = d v_ped/du (per control (j,c)):
//   rho_{k+1} = rho_k + T_c(k) - dt w_k
//   w_{k+1}   = w_k + dt (G_k rho_k - w_s w_k.y (1,1)^T),   G_k = -w_r (I - n n^T)/|d_k|, n = d_k/|d_k|
//   d g_k / d u_{j,c} = -n_k . rho_k                         (non-zero for j <= k-2)
// All 19 chains of a control advance together in one pass over k (forward mode),
// so G_k and n_k are computed once per step.  Lane = (sample, control); a warp owns
// 16 samples and stages the tile's 16 x 380 entries in shared memory in CSC order
// [control][j][sample][k], then streams the 38 column sub-runs out coalesced.
#pragma once
#include "saa_common.cuh"

namespace saa {

template <int S> struct CarRed {
  static constexpr int PX = 0;                    // + c*(S-1) + j    (c<2, j<S-1)
  static constexpr int PY = 2 * (S - 1);          // + c*(S-1) + j
  static constexpr int V = 4 * (S - 1);           // + j              (control 0)
  static constexpr int PHI = 4 * (S - 1) + S;     // + j              (control 1)
  static constexpr int VAL = 4 * (S - 1) + 2 * S; // + r  (r<4)
  static constexpr int N = VAL + 4;
};

template <typename T, int S> struct CarArgs {
  const T *x0;       // packed [f*Mpad + s], f = 0..3: pedestrian (qx, qy, wx, wy) initial
  const T *om;       // packed [f*Mpad + s], f = 0: omega_speed, 1: omega_repulsive
  const T *dw;       // packed [(k*2+f)*Mpad + s], f = 0,1: noise on states 6,7
  i64 M, Mpad;
  T us[S * 2];
  T ego0[4];         // ego initial (px, py, v, phi) (same for every sample)
  T dt, noise_c, v_des, d_min;
  T escale;          // scale applied to the Jacobian entries (1, or 0 -> not used)
  T ztol;
  T *Ax;
  i64 col_off[2 * (S - 1)];   // [c*(S-1)+j]
  T *ub; i64 ub_off;
  T *Z;
};

// Ego trajectory, shared by the whole block: E[k] = (px,py,v,phi)_k, TT[c][k] = T_c(k)
template <typename T, int S> struct CarEgo {
  T e[S + 1][4];
  T tt[2][S][2];
  T ucum[2][S + 1];   // ucum[c][k] = sum_{j<k} u_{j,c}
};

template <typename T, int S>
__device__ __forceinline__ void car_ego_rollout(const T *us, const T *ego0, T dt, CarEgo<T, S> &E) {
  // executed by one thread; O(S) work
  T px = ego0[0], py = ego0[1], v = ego0[2], phi = ego0[3];
  T u0 = T(0), u1 = T(0);
  E.ucum[0][0] = T(0); E.ucum[1][0] = T(0);
  for (int k = 0; k < S; ++k) {
    E.e[k][0] = px; E.e[k][1] = py; E.e[k][2] = v; E.e[k][3] = phi;
    T s, c;
    sincos((double)phi, (double *)nullptr, (double *)nullptr);  // placeholder removed below
    (void)s; (void)c;
  }
}

}  // namespace saa
