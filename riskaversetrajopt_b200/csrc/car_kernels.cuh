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
// All S-1 chains of a control advance together in one pass over k (forward mode),
// so G_k and n_k are computed once per step.  Lane = (sample, control); a warp owns
// 16 samples and stages the tile's 16 x 380 entries in shared memory in CSC order
// [control][j][sample][k], then streams the 2(S-1) column sub-runs out coalesced.
#pragma once
#include "saa_common.cuh"

namespace saa {

__device__ __forceinline__ double rsqrt_t(double x) { return rsqrt(x); }
__device__ __forceinline__ float rsqrt_t(float x) { return rsqrtf(x); }
__device__ __forceinline__ void sincos_t(double x, double *s, double *c) { sincos(x, s, c); }
__device__ __forceinline__ void sincos_t(float x, float *s, float *c) { sincosf(x, s, c); }

// FP64 sincos for the FP64-pipe bound friction kernel: 22 FP64-pipe instructions, no D2I/I2D
// conversions, no FP64 compares or selects; double-precision accuracy on |x| < 1e5 (measured
// against sincos(): tools/sincos_check.cu).  Quadrant n = rint(x 2/pi) by the 1.5 2^52 shift,
// three-step Cody-Waite reduction r = x - n pi/2 with FMAs (pi/2 split in three doubles; the FMA
// forms n * hi exactly), fdlibm minimax kernels on |r| <= pi/4, quadrant fix-up on the integer
// pipe.  Larger arguments take the library path.
// Coefficients live in constant memory: FP64 instructions take a constant-bank operand directly,
// whereas 64-bit immediates would be re-materialised with two moves per use inside the loop.
__constant__ double kSinCosC[17] = {
    6.36619772367581382433e-01,                        // 0: 2/pi
    -1.57079632679489655800e+00,                       // 1: -pi/2 (hi)
    -6.12323399573676603587e-17,                       // 2: -pi/2 (mid)
    1.49738490485916983e-33,                           // 3: -pi/2 (lo; this part of pi/2 is negative)
    1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06,    // 4..9: sin
    -1.98412698298579493134e-04, 8.33333333332248946124e-03, -1.66666666666666324348e-01,
    -1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07,   // 10..15: cos
    2.48015872894767294178e-05, -1.38888888888741095749e-03, 4.16666666666666019037e-02,
    -0.5};

// branch-free core, valid for |x| < 1e5
__device__ __forceinline__ void sincos_core(double x, double *s, double *c) {
  const double *K = kSinCosC;
  const double SHIFT = 6755399441055744.0;                    // 1.5 * 2^52
  const double t = fma(x, K[0], SHIFT);                       // x * 2/pi, integer part in the low bits
  const int n = __double2loint(t);
  const double nf = t - SHIFT;
  double r = fma(nf, K[1], x);
  r = fma(nf, K[2], r);
  r = fma(nf, K[3], r);            // keeps r relatively accurate next to multiples of pi/2
  const double z = r * r;
  double ps = fma(z, K[4], K[5]);
  ps = fma(z, ps, K[6]);
  ps = fma(z, ps, K[7]);
  ps = fma(z, ps, K[8]);
  ps = fma(z, ps, K[9]);
  const double sn = fma(r * z, ps, r);
  double pc = fma(z, K[10], K[11]);
  pc = fma(z, pc, K[12]);
  pc = fma(z, pc, K[13]);
  pc = fma(z, pc, K[14]);
  pc = fma(z, pc, K[15]);
  const double cs = fma(z * z, pc, fma(z, K[16], 1.0));
  // n mod 4: 0 (s, c)  1 (c, -s)  2 (-s, -c)  3 (-c, s)
  const bool swap = n & 1;
  double so = swap ? cs : sn, co = swap ? sn : cs;
  const int sflip = (n & 2) << 30, cflip = ((n + 1) & 2) << 30;
  so = __hiloint2double(__double2hiint(so) ^ sflip, __double2loint(so));
  co = __hiloint2double(__double2hiint(co) ^ cflip, __double2loint(co));
  *s = so; *c = co;
}
// FP32 counterpart (|x| < 1e4): same scheme with the 1.5 2^23 shift, pi/2 split in three floats,
// cephes sinf / cosf kernels (about 1 ulp in float)
__device__ __forceinline__ void sincos_core(float x, float *s, float *c) {
  const float SHIFT = 12582912.0f;                            // 1.5 * 2^23
  const float t = fmaf(x, 0.636619747f, SHIFT);
  const int n = __float_as_int(t);                            // low mantissa bits = n (mod 4 is all we need)
  const float nf = t - SHIFT;
  float r = fmaf(nf, -1.57079637e+00f, x);
  r = fmaf(nf, 4.37113883e-08f, r);
  r = fmaf(nf, 1.71512451e-15f, r);
  const float z = r * r;
  float ps = fmaf(z, -1.9515295891e-4f, 8.3321608736e-3f);
  ps = fmaf(z, ps, -1.6666654611e-1f);
  const float sn = fmaf(r * z, ps, r);
  float pc = fmaf(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
  pc = fmaf(z, pc, 4.166664568298827e-2f);
  const float cs = fmaf(z * z, pc, fmaf(z, -0.5f, 1.0f));
  const bool swap = n & 1;
  float so = swap ? cs : sn, co = swap ? sn : cs;
  so = __int_as_float(__float_as_int(so) ^ ((n & 2) << 30));
  co = __int_as_float(__float_as_int(co) ^ (((n + 1) & 2) << 30));
  *s = so; *c = co;
}
// N independent evaluations written stage by stage (source order alternates between them; ptxas
// is free to re-serialise, and does so in FP64)
template <int N>
__device__ __forceinline__ void sincos_core_n(const double (&x)[N], double (&s)[N], double (&c)[N]) {
  const double *K = kSinCosC;
  const double SHIFT = 6755399441055744.0;
  double t[N], nf[N], r[N], z[N], ps[N], pc[N], sn[N], cs[N];
  int n[N];
#pragma unroll
  for (int q = 0; q < N; ++q) t[q] = fma(x[q], K[0], SHIFT);
#pragma unroll
  for (int q = 0; q < N; ++q) { n[q] = __double2loint(t[q]); nf[q] = t[q] - SHIFT; }
#pragma unroll
  for (int q = 0; q < N; ++q) r[q] = fma(nf[q], K[1], x[q]);
#pragma unroll
  for (int q = 0; q < N; ++q) r[q] = fma(nf[q], K[2], r[q]);
#pragma unroll
  for (int q = 0; q < N; ++q) r[q] = fma(nf[q], K[3], r[q]);
#pragma unroll
  for (int q = 0; q < N; ++q) z[q] = r[q] * r[q];
#pragma unroll
  for (int q = 0; q < N; ++q) { ps[q] = fma(z[q], K[4], K[5]); pc[q] = fma(z[q], K[10], K[11]); }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int q = 0; q < N; ++q) { ps[q] = fma(z[q], ps[q], K[6 + i]); pc[q] = fma(z[q], pc[q], K[12 + i]); }
  }
#pragma unroll
  for (int q = 0; q < N; ++q) {
    sn[q] = fma(r[q] * z[q], ps[q], r[q]);
    cs[q] = fma(z[q] * z[q], pc[q], fma(z[q], K[16], 1.0));
  }
#pragma unroll
  for (int q = 0; q < N; ++q) {
    const bool swap = n[q] & 1;
    double so = swap ? cs[q] : sn[q], co = swap ? sn[q] : cs[q];
    so = __hiloint2double(__double2hiint(so) ^ ((n[q] & 2) << 30), __double2loint(so));
    co = __hiloint2double(__double2hiint(co) ^ (((n[q] + 1) & 2) << 30), __double2loint(co));
    s[q] = so; c[q] = co;
  }
}
template <int N>
__device__ __forceinline__ void sincos_core_n(const float (&x)[N], float (&s)[N], float (&c)[N]) {
#pragma unroll
  for (int q = 0; q < N; ++q) sincos_core(x[q], &s[q], &c[q]);
}

// true if the core's reduction does not cover x (|x| >= 1e5 in FP64, 1e4 in FP32, NaN, Inf)
__device__ __forceinline__ bool sincos_big(double x) { return (__double2hiint(x) & 0x7fffffff) >= 0x40f86a00; }
__device__ __forceinline__ bool sincos_big(float x) { return (__float_as_int(x) & 0x7fffffff) >= 0x461c4000; }

__device__ __forceinline__ void sincos_fast(double x, double *s, double *c) {
  if (sincos_big(x)) sincos(x, s, c);      // library slow path
  else sincos_core(x, s, c);
}
__device__ __forceinline__ void sincos_fast(float x, float *s, float *c) {
  if (sincos_big(x)) sincosf(x, s, c);
  else sincos_core(x, s, c);
}

template <int S> struct CarRed {
  static constexpr int PX = 0;                    // + c*(S-1) + j    (c<2, j<S-1)
  static constexpr int PY = 2 * (S - 1);          // + c*(S-1) + j
  static constexpr int V = 4 * (S - 1);           // + j              (control 0)
  static constexpr int PHI = 4 * (S - 1) + S;     // + j              (control 1)
  static constexpr int VAL = 4 * (S - 1) + 2 * S; // + r  (r<4)
  static constexpr int N = VAL + 4;
};

// CSC position of sample 0's run of u column 2J + c in a matrix with M samples:
// CAc + M * CBc.  Every earlier control step holds 2 x (3 final-row + 1 control-row)
// entries and 2 columns of (S-1-j') values per sample (layout.cuh, Layout::build).
template <int S, int J> struct CarCol {
  static constexpr int L = S - 1 - J;
  static constexpr int STRIDE = L | 1;
  static constexpr i64 CA0 = 8 * J + 3, CA1 = CA0 + 4;
  static constexpr i64 CB0 = 2 * J * (S - 1) - J * (J - 1), CB1 = CB0 + L;
};

template <typename T, typename TO, int S> struct CarArgs {
  const T *x0;       // packed [f*Mpad + s], f = 0..3: pedestrian (qx, qy, wx, wy) initial
  const T *om;       // packed [f*Mpad + s], f = 0: omega_speed, 1: omega_repulsive
  const T *dw;       // packed [(k*2+f)*Mpad + s], f = 0,1: noise on states 6,7
  i64 M, Mpad;
  T us[S * 2];
  T ego0[4];         // ego initial (px, py, v, phi) (same for every sample)
  T goal[4];
  T dt, noise_c, v_des, d_min;
  T ztol;
  TO *Ax; i64 M_out, first_out;   // TO = storage type of the outputs (double | float), T = arithmetic type
  TO *ub; i64 ub_off;
  TO *Z;
  double *sums;      // CarRed<S>::N doubles: M * (sample-independent final-row values)
  unsigned long long *nonfinite;   // += number of samples whose rollout met |d| = 0 / a non-finite value
};

// Ego trajectory, shared by the whole block
template <typename T, int S> struct CarEgo {
  T p[S + 1][2];      // ego position at step k
  T tt[2][S][2];      // T_c(k)
  T ucum[2][S + 1];   // sum_{j<k} u_{j,c}
  T fin[4];           // ego state at step S
};

// executed by one thread; O(S) work, identical for all samples (car/driving.py:167-172)
template <typename T, int S>
__device__ void car_ego_rollout(const T *us, const T *ego0, T dt, CarEgo<T, S> &E) {
  T px = ego0[0], py = ego0[1], v = ego0[2], phi = ego0[3];
  T u0 = T(0), u1 = T(0);
  for (int k = 0; k < S; ++k) {
    E.p[k][0] = px; E.p[k][1] = py;
    E.ucum[0][k] = u0; E.ucum[1][k] = u1;
    T sn, cs;
    sincos_t(phi, &sn, &cs);
    const T dt2 = dt * dt;
    E.tt[0][k][0] = dt2 * cs;      E.tt[0][k][1] = dt2 * sn;
    E.tt[1][k][0] = -dt2 * v * sn; E.tt[1][k][1] = dt2 * v * cs;
    px = px + dt * (v * cs);
    py = py + dt * (v * sn);
    v = v + dt * us[2 * k];
    phi = phi + dt * us[2 * k + 1];
    u0 += us[2 * k]; u1 += us[2 * k + 1];
  }
  E.p[S][0] = px; E.p[S][1] = py;
  E.ucum[0][S] = u0; E.ucum[1][S] = u1;
  E.fin[0] = px; E.fin[1] = py; E.fin[2] = v; E.fin[3] = phi;
}

// sample-independent final rows (car/driving.py:217-221, :271, :311-313):
// sums[r] = M * value so that the generic "sum / M_global" finalize applies across ranks
template <typename T, typename TO, int S>
__device__ void car_final_rows(const CarArgs<T, TO, S> &A, const CarEgo<T, S> &E, int tid, int nthreads) {
  using Rd = CarRed<S>;
  const double Md = (double)A.M;
  for (int r = tid; r < Rd::N; r += nthreads) {
    double val = 0.0;
    if (r < Rd::V) {                       // d p_S / d u_{j,c} = sum_{m=j+1}^{S-1} T_c(m)
      const int comp = r >= Rd::PY, q = r - (comp ? Rd::PY : Rd::PX), c = q / (S - 1), j = q % (S - 1);
      for (int m = j + 1; m < S; ++m) val += (double)E.tt[c][m][comp];
    } else if (r < Rd::VAL) {              // d v_S / d u_{j,0} = d phi_S / d u_{j,1} = dt
      val = (double)A.dt;
    } else {                               // linearisation offset -(ego_S - goal) + J u
      const int row = r - Rd::VAL;
      double ju = 0.0;
      if (row < 2) {
        for (int c = 0; c < 2; ++c)
          for (int m = 0; m < S; ++m) ju += (double)E.tt[c][m][row] * (double)E.ucum[c][m];
      } else {
        ju = (double)A.dt * (double)E.ucum[row - 2][S];
      }
      val = -((double)E.fin[row] - (double)A.goal[row]) + ju;
    }
    A.sums[r] = Md * val;
  }
}

// per-chain sensitivity state
template <typename T> struct CarChain { T rx, ry, wx, wy; };

#ifndef SAA_CAR_WARPS
#define SAA_CAR_WARPS 16   // warps per block (one block per SM; shared memory = WARPS x (geometry + staging))
#endif
#ifndef SAA_CAR_COPY_RT
#define SAA_CAR_COPY_RT 0  // 1: one run-time-parametrised copy loop for all columns; 0: unrolled per column
#endif
#ifndef SAA_CAR_GFORM
#define SAA_CAR_GFORM 1    // 1: G rho = -om rho + (om n)(n . rho) reusing the entry's dot product; 0: dense 2x2 G
#endif
#ifndef SAA_CAR_CAP
#define SAA_CAR_CAP 23     // staged values per sample and control in one chain pass (>= S-1); see CarPass
#endif

// Kernel structure (round 2).  A warp owns a tile of 16 samples, lane = (sample, control c).
//   pass A   rollout of the pedestrian ONCE per tile: per state k the unit vector n_k = d_k/|d_k|
//            and om_k = dt w_r / |d_k| go to shared memory (3 doubles per sample and step); the
//            tangent along u itself gives grad g . u for the upper bounds; Z_i.  The noise
//            increments are parked in the (still idle) staging buffer, not in 80 registers.
//   pass p   the forward-sensitivity chains j in [J0, J1) of this lane's control, all advancing
//            together over k = J0+1..S from the stored geometry (no rollout, no rsqrt, G_k rebuilt
//            with 5 flops); their entries are staged in CSC order [c][j][sample][k] and streamed
//            out as 16-byte stores.  Consecutive chains are packed into a pass while their padded
//            lengths fit SAA_CAR_CAP values per sample.
// Round 1 recomputed rollout + rsqrt in every pass and kept the noise in registers over all of
// them (168 registers, 160 B of spills, 12 warps/SM): 1.36 ms at M = 1e6.
template <int S, int CAP> struct CarPass {
  static_assert(CAP >= ((S - 1) | 1), "SAA_CAR_CAP must hold the longest chain");
  __host__ __device__ static constexpr int bound(int p) {           // first chain of pass p
    int j = 0;
    for (int q = 0; q < p && j < S - 1; ++q) {
      int acc = 0;
      while (j < S - 1 && acc + ((S - 1 - j) | 1) <= CAP) { acc += (S - 1 - j) | 1; ++j; }
    }
    return j;
  }
  __host__ __device__ static constexpr int npass() {
    int p = 0;
    while (bound(p) < S - 1) ++p;
    return p;
  }
  static constexpr int NPASS = npass();
  __host__ __device__ static constexpr int pre(int j0, int j) {     // staging offset of column j inside its pass
    int s = 0;
    for (int jj = j0; jj < j; ++jj) s += kTileSamples * ((S - 1 - jj) | 1);
    return s;
  }
  static constexpr int PER_C = kTileSamples * CAP;                  // staged block of one control
  static constexpr int UBROW = S | 1;
  // pass A stages the upper-bound rows 16 x UBROW here, the chain passes their columns
  __host__ __device__ static constexpr int stage_bytes(int szTO) {
    const int a = kTileSamples * UBROW * szTO;
    const int b = 2 * PER_C * szTO;
    return ((a > b ? a : b) + 15) / 16 * 16;
  }
  // Geometry rows of a warp: row 3k + q holds (n_x, n_y, om)[q] at state k for the 16 samples.
  // Pass A parks the noise increment (k, c) in row DWBASE + 2k + c, i.e. in the rows of LATER
  // states: step k writes rows 3k..3k+2 < DWBASE + 2k, the first noise row still unread.
  static constexpr int GEO_ROWS = 3 * (S + 1);
  static constexpr int DWBASE = GEO_ROWS - 2 * S;
};

template <typename T, typename TO, int S, int WARPS> struct CarSmem {
  using Ps = CarPass<S, SAA_CAR_CAP>;
  static constexpr int STAGE_BYTES = Ps::stage_bytes((int)sizeof(TO));
  CarEgo<T, S> ego;
  T geo[WARPS][Ps::GEO_ROWS][kTileSamples];              // n_x, n_y, om per sample and state (CarPass)
  alignas(16) unsigned char stage[WARPS][STAGE_BYTES];
};

// Copy a column run (ns samples x L values, rows of stride L|1 in shared memory) to its
// contiguous place in global memory with 16-byte streaming stores; the run starts 8-byte (4-byte)
// aligned only, so up to VEC-1 head / tail elements go out as scalars.  L is a run-time value
// (one instance of this code serves all columns: the fully unrolled per-column version was 200 KB
// of SASS and stalled on instruction fetch); rows are found with a multiply-high by ceil(2^32/L).
template <typename TO>
__device__ __forceinline__ void car_copy_run(TO *__restrict__ dst, const TO *__restrict__ src, int L, int n,
                                             int lane) {
  constexpr int VEC = 16 / (int)sizeof(TO);
  const int pad = (L & 1) ^ 1;                                   // row stride L|1
  const unsigned magic = (unsigned)(0x100000000ull / (unsigned)L) + 1u;
  auto at = [&](int e) -> TO { return src[e + (int)__umulhi((unsigned)e, magic) * pad]; };
  const int head = min(n, (int)((VEC - (((uintptr_t)dst / sizeof(TO)) & (VEC - 1))) & (VEC - 1)));
  const int nvec = (n - head) / VEC, tail = n - head - nvec * VEC;
  if (lane < head) st_stream(dst + lane, at(lane));
#pragma unroll 2
  for (int v = lane; v < nvec; v += 32) {
    const int e = head + v * VEC;
    if constexpr (VEC == 2) {
      double2 val;
      val.x = (double)at(e); val.y = (double)at(e + 1);
      __stcs(reinterpret_cast<double2 *>(dst + e), val);
    } else {
      float4 val;
      val.x = (float)at(e); val.y = (float)at(e + 1); val.z = (float)at(e + 2); val.w = (float)at(e + 3);
      __stcs(reinterpret_cast<float4 *>(dst + e), val);
    }
  }
  if (lane < tail) st_stream(dst + head + nvec * VEC + lane, at(head + nvec * VEC + lane));
}

// columns j in [j0, j1) of both controls; position of the run in the CSC array: CarCol
template <typename T, typename TO, int S>
__device__ __forceinline__ void car_copy_cols(const CarArgs<T, TO, S> &A, const TO *stage, int j0, int j1,
                                              i64 sbase, int ns, int lane) {
  using Ps = CarPass<S, SAA_CAR_CAP>;
  int off = 0;
#pragma unroll 1
  for (int j = j0; j < j1; ++j) {
    const int L = S - 1 - j;
    const i64 ca0 = 8 * j + 3, cb0 = (i64)2 * j * (S - 1) - (i64)j * (j - 1);
    TO *d0 = A.Ax + (ca0 + A.M_out * cb0 + sbase * L);
    TO *d1 = A.Ax + (ca0 + 4 + A.M_out * (cb0 + L) + sbase * L);
    car_copy_run<TO>(d0, stage + off, L, ns * L, lane);
    car_copy_run<TO>(d1, stage + Ps::PER_C + off, L, ns * L, lane);
    off += kTileSamples * (L | 1);
  }
}

// ---- compile-time unrolled variant of the copy-out (one instance per column; SAA_CAR_COPY_RT=0) ----
// Copy a column run (ns samples x L values, rows of stride STRIDE in shared memory) to its
// contiguous place in global memory with 16-byte streaming stores; the run starts 8-byte (4-byte)
// aligned only, so up to VEC-1 head / tail elements go out as scalars.
template <typename TO, int L, int STRIDE>
__device__ __forceinline__ void car_copy_run_t(TO *__restrict__ dst, const TO *__restrict__ src, int n, int lane) {
  constexpr int VEC = 16 / (int)sizeof(TO);
  auto at = [&](int e) -> TO { const int row = e / L; return src[e + row * (STRIDE - L)]; };
  const int head = min(n, (int)((VEC - (((uintptr_t)dst / sizeof(TO)) & (VEC - 1))) & (VEC - 1)));
  const int nvec = (n - head) / VEC, tail = n - head - nvec * VEC;
  if (lane < head) st_stream(dst + lane, at(lane));
#pragma unroll 2
  for (int v = lane; v < nvec; v += 32) {
    const int e = head + v * VEC;
    if constexpr (VEC == 2) {
      double2 val;
      val.x = (double)at(e); val.y = (double)at(e + 1);
      __stcs(reinterpret_cast<double2 *>(dst + e), val);
    } else {
      float4 val;
      val.x = (float)at(e); val.y = (float)at(e + 1); val.z = (float)at(e + 2); val.w = (float)at(e + 3);
      __stcs(reinterpret_cast<float4 *>(dst + e), val);
    }
  }
  if (lane < tail) st_stream(dst + head + nvec * VEC + lane, at(head + nvec * VEC + lane));
}

template <typename T, typename TO, int S, int J0, int J, int J1>
__device__ __forceinline__ void car_copy_cols_t(const CarArgs<T, TO, S> &A, const TO *stage, i64 sbase, int ns,
                                              int lane) {
  if constexpr (J < J1) {
    using C = CarCol<S, J>;
    using Ps = CarPass<S, SAA_CAR_CAP>;
    i64 sb = sbase, mout = A.M_out;
    opaque(sb); opaque(mout);   // 2 IMADs per column instead of 2(S-1) live 64-bit bases
    car_copy_run_t<TO, C::L, C::STRIDE>(A.Ax + (C::CA0 + mout * C::CB0 + sb * C::L), stage + Ps::pre(J0, J),
                                      ns * C::L, lane);
    car_copy_run_t<TO, C::L, C::STRIDE>(A.Ax + (C::CA1 + mout * C::CB1 + sb * C::L),
                                      stage + Ps::PER_C + Ps::pre(J0, J), ns * C::L, lane);
    car_copy_cols_t<T, TO, S, J0, J + 1, J1>(A, stage, sbase, ns, lane);
  }
}

// pass A: rollout, geometry -> shared memory, upper bounds, Z_i
template <typename T, typename TO, int S>
__device__ __forceinline__ void car_rollout_pass(const CarArgs<T, TO, S> &A, const CarEgo<T, S> &E,
                                                 T (*geo)[kTileSamples], unsigned char *stage_raw,
                                                 int c, int si, i64 s, T w_s, T w_r, T &zmax_out, bool &bad) {
  using Ps = CarPass<S, SAA_CAR_CAP>;
  TO *ubrow = reinterpret_cast<TO *>(stage_raw) + si * Ps::UBROW;
  // this lane's half of the noise (c = 0: state 6, c = 1: state 7), all loads in flight at once,
  // parked in the geometry rows of later states (CarPass::DWBASE)
  {
    T tmp[S];
#pragma unroll
    for (int k = 0; k < S; ++k) tmp[k] = __ldcs(A.dw + (i64)(2 * k + c) * A.Mpad + s);
#pragma unroll
    for (int k = 0; k < S; ++k) geo[Ps::DWBASE + 2 * k + c][si] = tmp[k];
  }
  T qx = __ldcs(A.x0 + s), qy = __ldcs(A.x0 + A.Mpad + s);
  T wx = __ldcs(A.x0 + 2 * A.Mpad + s), wy = __ldcs(A.x0 + 3 * A.Mpad + s);
  __syncwarp();
  const T dt = A.dt, wsdt = w_s * dt, dtwr = dt * w_r;
  CarChain<T> cu{T(0), T(0), T(0), T(0)};     // tangent along u itself (for grad g . u)
  T zmax = -INFINITY;
  static_for<0, S + 1>([&](auto kc) {
    constexpr int k = decltype(kc)::value;
    // geometry at state k (car/driving.py:150-154, :227-231)
    const T dx = E.p[k][0] - qx, dy = E.p[k][1] - qy;
    const T n2 = fma(dx, dx, dy * dy);
    const T inv_n = rsqrt_t(n2);
    bad |= !(n2 > T(0)) || !(n2 < T(INFINITY));
    const T nhx = dx * inv_n, nhy = dy * inv_n;
    const T om_n = dtwr * inv_n;
    if constexpr (k >= 1) {
      // upper bound -g_k + grad g_k . u (:278), grad g . u summed over both controls
      T gu = -fma(nhx, cu.rx, nhy * cu.ry);
      gu += __shfl_xor_sync(0xffffffffu, gu, 16);        // also: every lane has read the noise of step k-1
      const T g = A.d_min - n2 * inv_n;
      zmax = fmax(zmax, g);
      if (c == (k & 1)) ubrow[k - 1] = (TO)(gu - g);
    }
    // rows 3k..3k+2 held noise increments of steps <= (3k + 2 - DWBASE) / 2 < k: every lane has read
    // them (program order + this barrier; racecheck does not count the shuffle above as one)
    __syncwarp();
    if (c == 0) { geo[3 * k][si] = nhx; geo[3 * k + 2][si] = om_n; }
    else geo[3 * k + 1][si] = nhy;
    if constexpr (k < S) {
      // dt dF/dp_ego = G_k = -om (I - n n^T)
      const T ttx = E.tt[c][k][0], tty = E.tt[c][k][1];
      const T uc = E.ucum[c][k];
      const T d = fma(nhx, cu.rx, nhy * cu.ry);
      const T nwx = fma(om_n * nhx, d, fma(-om_n, cu.rx, fma(-wsdt, cu.wy, cu.wx)));
      const T nwy = fma(om_n * nhy, d, fma(-om_n, cu.ry, fma(-wsdt, cu.wy, cu.wy)));
      cu.rx = fma(-dt, cu.wx, fma(uc, ttx, cu.rx));
      cu.ry = fma(-dt, cu.wy, fma(uc, tty, cu.ry));
      cu.wx = nwx; cu.wy = nwy;
      // one Euler-Maruyama step of the pedestrian (:187-204)
      const T sp = w_s * (A.v_des - wy);
      const T fx = fma(-w_r, nhx, sp), fy = fma(-w_r, nhy, sp);
      const T nqx = fma(dt, wx, qx), nqy = fma(dt, wy, qy);
      wx = wx + dt * fx + A.noise_c * geo[Ps::DWBASE + 2 * k][si];
      wy = wy + dt * fy + A.noise_c * geo[Ps::DWBASE + 2 * k + 1][si];
      qx = nqx; qy = nqy;
    }
  });
  zmax_out = zmax;
}

// One pass over the horizon for the chains j in [J0, J1) of this lane's control c, driven by the
// geometry pass A left in shared memory.
template <typename T, typename TO, int S, int J0, int J1>
__device__ __forceinline__ void car_chain_pass(const CarArgs<T, TO, S> &A, const CarEgo<T, S> &E,
                                               const T (*geo)[kTileSamples], TO *stage, int c, int si,
                                               T wsdt) {
  using Ps = CarPass<S, SAA_CAR_CAP>;
  constexpr int NJ = J1 - J0;
  const T dt = A.dt;
  TO *cstage = stage + c * Ps::PER_C;
#pragma nv_diag_suppress 549                  // every chain is initialised at its birth step (k = j + 1) before use
  CarChain<T> ch[NJ > 0 ? NJ : 1];
  static_for<J0 + 1, S + 1>([&](auto kc) {
    constexpr int k = decltype(kc)::value;
    // chains of this pass that are alive at state k: J0 <= j <= min(k-2, J1-1)
    constexpr int JE = (k - 1 < J1) ? (k - 1) : J1;          // exclusive end
    constexpr bool ALIVE = JE > J0;
    T nhx = T(0), nhy = T(0);
    [[maybe_unused]] T dot[NJ > 0 ? NJ : 1];   // n_k . rho_k per chain: the (negated) entry AND what G_k rho_k needs
    if constexpr (ALIVE) {
      nhx = geo[3 * k][si]; nhy = geo[3 * k + 1][si];
      // row k of the sample: d g_k / d u_{j,c} = -n_k . rho_k^{(j,c)}
      static_for<J0, JE>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        dot[j - J0] = fma(nhx, ch[j - J0].rx, nhy * ch[j - J0].ry);
        cstage[Ps::pre(J0, j) + si * CarCol<S, j>::STRIDE + (k - j - 2)] = (TO)(-dot[j - J0]);
      });
    }
    if constexpr (k < S) {
      const T ttx = E.tt[c][k][0], tty = E.tt[c][k][1];
      if constexpr (ALIVE) {
#if SAA_CAR_GFORM
        // dt dF/dp_ego = G_k = -om (I - n n^T):  G rho = -om rho + (om n) (n . rho)
        const T om_n = geo[3 * k + 2][si];
        const T onx = om_n * nhx, ony = om_n * nhy;
        static_for<J0, JE>([&](auto jc) {
          constexpr int j = decltype(jc)::value;
          CarChain<T> &h = ch[j - J0];
          const T d = dot[j - J0];
          const T nwx = fma(onx, d, fma(-om_n, h.rx, fma(-wsdt, h.wy, h.wx)));
          const T nwy = fma(ony, d, fma(-om_n, h.ry, fma(-wsdt, h.wy, h.wy)));
          h.rx = fma(-dt, h.wx, h.rx + ttx);
          h.ry = fma(-dt, h.wy, h.ry + tty);
          h.wx = nwx; h.wy = nwy;
        });
      }
#else
        const T om_n = geo[3 * k + 2][si];
        const T g11 = -om_n * fma(-nhx, nhx, T(1)), g12 = om_n * nhx * nhy,
                g22 = -om_n * fma(-nhy, nhy, T(1));           // dt * dF/dp_ego
        static_for<J0, JE>([&](auto jc) {
          constexpr int j = decltype(jc)::value;
          CarChain<T> &h = ch[j - J0];
          const T sy = wsdt * h.wy;
          const T nwx = fma(g11, h.rx, fma(g12, h.ry, h.wx - sy));
          const T nwy = fma(g12, h.rx, fma(g22, h.ry, h.wy - sy));
          h.rx = fma(-dt, h.wx, h.rx + ttx);
          h.ry = fma(-dt, h.wy, h.ry + tty);
          h.wx = nwx; h.wy = nwy;
        });
      }
#endif
      // chain j = k-1 is born at this step: rho_{k+1} = T_c(k), w_{k+1} = 0
      if constexpr (k - 1 >= J0 && k - 1 < J1) {
        ch[k - 1 - J0].rx = ttx; ch[k - 1 - J0].ry = tty; ch[k - 1 - J0].wx = T(0); ch[k - 1 - J0].wy = T(0);
      }
    }
  });
}

template <typename T, typename TO, int S, int P>
__device__ __forceinline__ void car_passes(const CarArgs<T, TO, S> &A, const CarEgo<T, S> &E,
                                           const T (*geo)[kTileSamples], TO *stage, int c, int si,
                                           int lane, i64 s0, int ns, T wsdt) {
  using Ps = CarPass<S, SAA_CAR_CAP>;
  if constexpr (P < Ps::NPASS) {
    constexpr int J0 = Ps::bound(P), J1 = Ps::bound(P + 1);
    car_chain_pass<T, TO, S, J0, J1>(A, E, geo, stage, c, si, wsdt);
    __syncwarp();
#if SAA_CAR_COPY_RT
    car_copy_cols<T, TO, S>(A, stage, J0, J1, s0 + A.first_out, ns, lane);
#else
    car_copy_cols_t<T, TO, S, J0, J0, J1>(A, stage, s0 + A.first_out, ns, lane);
#endif
    __syncwarp();
    car_passes<T, TO, S, P + 1>(A, E, geo, stage, c, si, lane, s0, ns, wsdt);
  }
}

// ---- K2: linearize + assemble ---------------------------------------------------
template <typename T, typename TO, int S, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
car_assemble_kernel(const __grid_constant__ CarArgs<T, TO, S> A) {
  using Ps = CarPass<S, SAA_CAR_CAP>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  auto &sm = *reinterpret_cast<CarSmem<T, TO, S, WARPS> *>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = lane >> 4, si = lane & 15;
  if (threadIdx.x == 0) car_ego_rollout<T, S>(A.us, A.ego0, A.dt, sm.ego);
  __syncthreads();
  if (blockIdx.x == 0 && A.sums != nullptr) car_final_rows<T, TO, S>(A, sm.ego, threadIdx.x, WARPS * 32);
  if (A.Ax == nullptr) return;              // relaxed iteration: only the final rows are needed
  const CarEgo<T, S> &E = sm.ego;
  T (*geo)[kTileSamples] = sm.geo[warp];
  unsigned char *stage_raw = sm.stage[warp];
  bool bad = false;

  const i64 ntiles = (A.M + kTileSamples - 1) / kTileSamples;
#pragma unroll 1
  for (i64 tile = (i64)blockIdx.x * WARPS + warp; tile < ntiles; tile += (i64)gridDim.x * WARPS) {
    const i64 s0 = tile * kTileSamples;
    const int ns = (int)min((i64)kTileSamples, A.M - s0);
    const bool active = si < ns;
    const i64 s = s0 + (active ? si : 0);
    const T w_s = __ldcs(A.om + s), w_r = __ldcs(A.om + A.Mpad + s);
    T zmax;
    bool tile_bad = false;
    car_rollout_pass<T, TO, S>(A, E, geo, stage_raw, c, si, s, w_s, w_r, zmax, tile_bad);
    bad |= tile_bad && active && c == 0;
    if (A.Z != nullptr && c == 0 && active) A.Z[s] = (TO)(zmax - A.ztol);
    __syncwarp();
    if (A.ub != nullptr)
      car_copy_run<TO>(A.ub + A.ub_off + s0 * S, reinterpret_cast<const TO *>(stage_raw), S, ns * S, lane);
    __syncwarp();
    car_passes<T, TO, S, 0>(A, E, geo, reinterpret_cast<TO *>(stage_raw), c, si, lane, s0, ns, w_s * A.dt);
  }
  if (A.nonfinite != nullptr) {
    const unsigned m = __ballot_sync(0xffffffffu, bad);
    if (lane == 0 && m) atomicAdd(A.nonfinite, (unsigned long long)__popc(m));
  }
}

// ---- K4 / K5: rollout only ----------------------------------------------------------
template <typename T, typename TO, int S> struct CarRollArgs {
  const T *x0, *om, *dw;
  i64 M, Mpad;
  T us[S * 2];
  T ego0[4];
  T dt, noise_c, v_des, d_min;
  TO *Xs;           // (M, S+1, 8) or nullptr
  TO *Z;            // (M) or nullptr
  T ztol, t_risk, sat_tol;
  double *partials; // [gridDim.x][3] or nullptr
};

template <typename T, typename TO, int S, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
car_rollout_kernel(const __grid_constant__ CarRollArgs<T, TO, S> A) {
  constexpr int ROW = (S + 1) * 8, STRIDE = ROW | 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ T ego[S + 1][4];
  __shared__ double red[WARPS][3];
  TO *stage = reinterpret_cast<TO *>(smem_raw) + (threadIdx.x >> 5) * 32 * STRIDE;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    T px = A.ego0[0], py = A.ego0[1], v = A.ego0[2], phi = A.ego0[3];
    for (int k = 0; k <= S; ++k) {
      ego[k][0] = px; ego[k][1] = py; ego[k][2] = v; ego[k][3] = phi;
      if (k == S) break;
      T sn, cs;
      sincos_t(phi, &sn, &cs);
      px = px + A.dt * (v * cs); py = py + A.dt * (v * sn);
      v = v + A.dt * A.us[2 * k]; phi = phi + A.dt * A.us[2 * k + 1];
    }
  }
  __syncthreads();
  double acc_excess = 0.0, acc_sat = 0.0, acc_max = -INFINITY;
  const i64 ntiles = (A.M + 31) / 32;
  for (i64 tile = (i64)blockIdx.x * WARPS + warp; tile < ntiles; tile += (i64)gridDim.x * WARPS) {
    const i64 s0 = tile * 32;
    const int ns = (int)min((i64)32, A.M - s0);
    const bool active = lane < ns;
    const i64 s = s0 + (active ? lane : 0);
    T qx = A.x0[s], qy = A.x0[A.Mpad + s], wx = A.x0[2 * A.Mpad + s], wy = A.x0[3 * A.Mpad + s];
    const T w_s = A.om[s], w_r = A.om[A.Mpad + s];
    T zmax = -INFINITY;
    TO *mine = stage + lane * STRIDE;
    T dwx[S], dwy[S];
#pragma unroll
    for (int k = 0; k < S; ++k) {
      dwx[k] = __ldcs(A.dw + (i64)(2 * k) * A.Mpad + s);
      dwy[k] = __ldcs(A.dw + (i64)(2 * k + 1) * A.Mpad + s);
    }
#pragma unroll
    for (int k = 0; k <= S; ++k) {
      if (A.Xs != nullptr) {
#pragma unroll
        for (int f = 0; f < 4; ++f) mine[k * 8 + f] = (TO)ego[k][f];
        mine[k * 8 + 4] = (TO)qx; mine[k * 8 + 5] = (TO)qy; mine[k * 8 + 6] = (TO)wx; mine[k * 8 + 7] = (TO)wy;
      }
      const T dx = ego[k][0] - qx, dy = ego[k][1] - qy;
      const T n2 = fma(dx, dx, dy * dy);
      const T inv_n = rsqrt_t(n2);
      if (k >= 1) zmax = fmax(zmax, A.d_min - n2 * inv_n);
      if (k < S) {
        const T sp = w_s * (A.v_des - wy);
        const T fx = fma(-w_r, dx * inv_n, sp), fy = fma(-w_r, dy * inv_n, sp);
        const T nqx = fma(A.dt, wx, qx), nqy = fma(A.dt, wy, qy);
        wx = wx + A.dt * fx + A.noise_c * dwx[k];
        wy = wy + A.dt * fy + A.noise_c * dwy[k];
        qx = nqx; qy = nqy;
      }
    }
    const T Zi = zmax - A.ztol;
    if (A.Z != nullptr && active) A.Z[s] = (TO)Zi;
    if (active) {
      acc_excess += (double)fmax(Zi - A.t_risk, T(0));
      acc_sat += (Zi <= A.sat_tol) ? 1.0 : 0.0;
      acc_max = fmax(acc_max, (double)Zi);
    }
    if (A.Xs != nullptr) {
      __syncwarp();
      copy_run<TO, ROW, STRIDE>(A.Xs + s0 * ROW, stage, ns * ROW, lane);
      __syncwarp();
    }
  }
  if (A.partials != nullptr) {
    acc_excess = sum32(acc_excess); acc_sat = sum32(acc_sat); acc_max = max32(acc_max);
    if (lane == 0) { red[warp][0] = acc_excess; red[warp][1] = acc_sat; red[warp][2] = acc_max; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double e = 0.0, cc = 0.0, m = -INFINITY;
      for (int w = 0; w < WARPS; ++w) { e += red[w][0]; cc += red[w][1]; m = fmax(m, red[w][2]); }
      A.partials[(i64)blockIdx.x * 3 + 0] = e;
      A.partials[(i64)blockIdx.x * 3 + 1] = cc;
      A.partials[(i64)blockIdx.x * 3 + 2] = m;
    }
  }
}

// ---- repack the reference-layout sample set (car/driving.py:95-120) ----------------
template <typename T>
__global__ void car_pack_kernel(const double *__restrict__ states_init, const double *__restrict__ w_s,
                                const double *__restrict__ w_r, const double *__restrict__ DWs, i64 M,
                                i64 Mpad, int S, T *x0, T *om, T *dw) {
  const i64 s = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= Mpad) return;
  const i64 src = s < M ? s : M - 1;
  for (int f = 0; f < 4; ++f) x0[f * Mpad + s] = (T)states_init[src * 8 + 4 + f];
  om[s] = (T)w_s[src];
  om[Mpad + s] = (T)w_r[src];
  for (int k = 0; k < S; ++k)
    for (int f = 0; f < 2; ++f) dw[(i64)(2 * k + f) * Mpad + s] = (T)DWs[(src * S + k) * 8 + 6 + f];
}

// Ego initial state must be the same for every sample (the reference only perturbs the
// pedestrian part, car/driving.py:104-110); checked on the device at set_samples time.
__global__ void car_check_ego_kernel(const double *__restrict__ states_init, i64 M, int *flag) {
  const i64 s = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= M) return;
  for (int f = 0; f < 4; ++f)
    if (states_init[s * 8 + f] != states_init[f]) *flag = 1;
}

}  // namespace saa
