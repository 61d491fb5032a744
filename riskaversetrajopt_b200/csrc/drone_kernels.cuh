// Quadrotor SAA kernels (K1 assemble, K4 rollout, K5 CVaR terms), sm_100a.
//
// What is computed (reference drone/drone_risk.py):
//   rollout            :139-155   x+ = x + dt b(x,u,m) + sqrt(dt) sigma dW
//   constraints        :164-213   x_S - x_final ;  g[o,k] = 1 - (p_k-c_o)^T Q_o (p_k-c_o)
//   per-sample Jacobian:239-280   jacfwd wrt us  (here: closed-form recursion)
//   sample mean / pack :282-374
//
// The three axes decouple (diagonal feedback gain, per-axis drag), so the
// Jacobian of the 6-state rollout is three independent 2x2 chains:
//   p+ = p + dt v ;  v+ = v + dt (u - kp p - kd v - c|v|v)/m + noise
//   d(p+,v+)/d(p,v) = [[1, dt], [a21, a22_k]],  a21 = -kp dt/m,
//   a22_k = 1 - dt (kd + 2c|v_k|)/m,   d v+/du = dt/m.
//
// Thread mapping: one warp owns a tile of 16 samples; lanes 0-15 run the x axis
// of those samples, lanes 16-31 the y axis (the z axis, which only feeds the
// sample-mean rows, is split between the two lanes of a sample by an adjoint
// pass).  For each control step j the lane runs the sensitivity chain
// k = j+1..S in registers and stages the 3*(S-1-j) CSC entries of its sample
// for column (j, axis) in shared memory; the warp then streams the two column
// sub-runs (16 samples x 3*(S-1-j) contiguous doubles each) to global memory
// with consecutive lanes on consecutive addresses.
#pragma once
#include "saa_common.cuh"

// ---- tuning switches (defaults = the measured best; see profiles/ and DESIGN.md) ----
#ifndef SAA_PRELOAD
#define SAA_PRELOAD 1      // load the tile's noise increments before the rollouts
#endif
#ifndef SAA_PREFETCH
#define SAA_PREFETCH 0     // prefetch.global.L2 of the next tile's input lines
#endif
#ifndef SAA_PAIR
#define SAA_PAIR 0         // process chains (J, S-2-J) together
#endif
#ifndef SAA_COPY
#define SAA_COPY 4         // 0: 8-byte loop, 1: 8-deep batches, 2: non-inlined body, 3: TMA bulk stores, 4: 16-byte vector copy
#endif
#ifndef SAA_BPS
#define SAA_BPS 2          // resident blocks per SM the kernel is compiled for
#endif
#ifndef SAA_GX
#define SAA_GX 0           // 1: lanes exchange (p, tangent) once per step; 0: partial sums per (step, obstacle)
#endif
#ifndef SAA_Z_INLINE
#define SAA_Z_INLINE 0     // 1: z-axis mean rows inside the assemble kernel (measured 9 % slower)
#endif

namespace saa {

template <int S> struct DroneRed {
  static constexpr int FIN_P = 0;                 // + a*(S-1) + j   (a<3, j<S-1)
  static constexpr int FIN_V = 3 * (S - 1);       // + a*S + j       (a<3, j<S)
  static constexpr int VAL = FIN_V + 3 * S;       // + r             (r<6)
  static constexpr int N = VAL + 6;
};

template <typename T, int S> struct DroneArgs {
  const T *mass, *dw, *q;   // packed: mass[M]; dw[(k*3+a)*Mpad+s]; q[(o*2+a)*Mpad+s]
  i64 M, Mpad;
  T us[S * 3];
  T dt, noise_c, drag, kp, kd;
  T x0[6], xf[6];
  T oc[3][2];
  T escale;                 // multiplier (x relaxation scale) applied to the Jacobian entries
  T ubscale, ubpad;         // upper bound = ubscale * (-g + grad g . u) - ubpad
  T ztol;                   // Z_i = max g - ztol
  T *Ax;
  // Sub-run of local sample 0 in u column (j, a): Ax + CA(j,a) + M_out*CB(j,a) + first_out*LEN(j)
  // with compile-time CA, CB (DroneChain); the host checks them against Layout::run_start.
  i64 M_out, first_out;
  T *ub;                    // base of the upper-bound vector (nullptr: skip)
  i64 ub_off;               // row of local sample 0's first sample row
  T *Z;                     // per-sample max constraint (nullptr: skip)
  double *partials;         // [gridDim.x][DroneRed<S>::N]
};

// ---- sensitivity chains ----------------------------------------------------------
// Column (J, axis) of a sample holds rows k = J+2..S for each of the 3 obstacles:
// LEN = 3 (S-1-J) contiguous values.  Chains are processed in pairs (J, S-2-J) so
// that every pass has the same amount of work (LEN1 + LEN2 = 3S) and two
// independent dependency chains per thread (ILP).
template <int S, int J> struct DroneChain {
  static constexpr int L = S - 1 - J;      // rows k = J+2..S carry an entry
  static constexpr int LEN = 3 * L;        // per-sample run length in the CSC column
  static constexpr int STRIDE = LEN | 1;   // odd stride: conflict-free 64-bit staging stores
  // CSC position of sample 0's run of column c = 3J + a in a matrix with M samples is
  // CAa + M * CBa: every earlier control step holds 6 final-row + 3 control-row
  // entries and 2 columns of 3(S-1-j') values per sample (layout.cuh, Layout::build)
  static constexpr i64 CA0 = 9 * J + 2, CA1 = CA0 + 3;
  static constexpr i64 CB0 = 6 * J * (S - 1) - 3 * J * (J - 1), CB1 = CB0 + 3 * (S - 1 - J);
};

template <typename T, int S> struct DroneChainState {
  T sp, sv;
};

// one step k -> k+1 of chain J, staging the three entries of row k+1
template <typename T, int S, int J>
__device__ __forceinline__ void drone_chain_step(int k, const T (&P)[S + 1], const T (&A22)[S],
                                                 const T (&q2)[3], const T (&oc)[3], T dt, T a21,
                                                 T *mine, DroneChainState<T, S> &c) {
  using C = DroneChain<S, J>;
  const T nsp = fma(dt, c.sv, c.sp);
  const T nsv = fma(A22[k], c.sv, a21 * c.sp);
  c.sp = nsp; c.sv = nsv;                  // now d(p,v)_{k+1} / du_J
  const int kk = k - J - 1;
#pragma unroll
  for (int o = 0; o < 3; ++o) {
    const T coef = fma(q2[o], P[k + 1], oc[o]);    // escale * d g[o,k+1]/dp = q2 (p - c): oc holds -q2*c
    mine[o * C::L + kk] = coef * c.sp;
  }
}

template <typename T, int S, int WARPS>
struct DroneSmem {
  static constexpr int HALF = 2 * kTileSamples * (3 * S + 2) + 8;   // one staging buffer (>= every Stager<T,LEN>::SIZE)
#if SAA_COPY == 3
  static constexpr int STAGE = 2 * HALF;   // TMA copy-out is double buffered: the engine drains one
                                           // column pair while the warp computes the next
#else
  static constexpr int STAGE = HALF;
#endif
  T stage[WARPS][STAGE];
  double wacc[WARPS][DroneRed<S>::N];
};

// Copy n = ns*LEN staged values (row stride LEN + EXTRA) to a contiguous global run.
template <typename T>
__device__ __noinline__ void copy_run_rt(T *__restrict__ dst, const T *__restrict__ src, int n,
                                         int extra, unsigned magic) {
  const int lane = threadIdx.x & 31;
#pragma unroll 1
  for (int e0 = lane; e0 < n; e0 += 256) {
    T v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int e = min(e0 + u * 32, n - 1);
      v[u] = src[e + (int)__umulhi((unsigned)e, magic) * extra];
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (e0 + u * 32 < n) st_stream(dst + e0 + u * 32, v[u]);
  }
}
template <typename T, int LEN, int EXTRA>
__device__ __forceinline__ void copy_run8(T *__restrict__ dst, const T *__restrict__ src, int n,
                                          int lane) {
  constexpr unsigned MAGIC = (unsigned)((0x100000000ull + LEN - 1) / LEN);
#if SAA_COPY == 2
  copy_run_rt<T>(dst, src, n, EXTRA, MAGIC);
#elif SAA_COPY == 1
#pragma unroll 1
  for (int e0 = lane; e0 < n; e0 += 256) {
    T v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int e = min(e0 + u * 32, n - 1);
      const int i = EXTRA ? (int)__umulhi((unsigned)e, MAGIC) : 0;
      v[u] = src[e + i * EXTRA];
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (e0 + u * 32 < n) st_stream(dst + e0 + u * 32, v[u]);
  }
#else
  copy_run<T, LEN, LEN + EXTRA>(dst, src, n, lane);
#endif
}

// Launch constants the copy-out needs, read from the parameter bank ONCE and pinned in
// registers (an LDC in front of every column store showed up as long-scoreboard stalls).
template <typename T> struct DroneOut {
  T *Ax;
  i64 mout, first;
};

template <typename T, int S, int J>
__device__ __forceinline__ void drone_chain_pairs(const DroneArgs<T, S> &A, const DroneOut<T> &O,
                                                  const T (&P)[S + 1],
                                                  const T (&A22)[S], const T (&q2)[3],
                                                  const T (&oca)[3], T a21, T dtm, T *stage,
                                                  double *wacc, int a, int si, int lane, i64 s0,
                                                  int ns, bool active) {
  using Rd = DroneRed<S>;
  constexpr int J2 = SAA_PAIR ? S - 2 - J : J;            // partner chain (J2 >= J)
  if constexpr (SAA_PAIR ? (J > J2) : (J >= S - 1)) {
    // all chains with sample rows are done; last control step J = S-1 has none:
    // only d v_S/du = dt/m enters the mean rows
    const double rv = sum16((double)(active ? dtm : T(0)));
    if (si == 0) wacc[Rd::FIN_V + a * S + (S - 1)] += rv;
  } else {
    using C1 = DroneChain<S, J>;
    using C2 = DroneChain<S, J2>;
    constexpr bool PAIR = (J2 != J);
    // optimisation barriers: keep per-chain coefficient math from being hoisted
    // (common subexpressions across the unrolled chains would cost ~60 live doubles)
    T q2j[3] = {q2[0], q2[1], q2[2]}, ocj[3] = {-q2[0] * oca[0], -q2[1] * oca[1], -q2[2] * oca[2]};
    opaque(q2j[0]); opaque(q2j[1]); opaque(q2j[2]);
    opaque(ocj[0]); opaque(ocj[1]); opaque(ocj[2]);
#if SAA_COPY >= 3
    static_assert(!SAA_PAIR, "TMA / vector copy-out is implemented for single chains");
    using St = Stager<T, C1::LEN>;
    static_assert(St::SIZE <= DroneSmem<T, S, 1>::HALF, "staging buffer too small");
#if SAA_COPY == 3
    // upper bounds went through buffer 0, chain 0 takes buffer 1, chain 1 buffer 0, ...
    T *const stg = stage + ((J & 1) ? 0 : DroneSmem<T, S, 1>::HALF);
#else
    T *const stg = stage;
#endif
#ifdef SAA_ABL_SAMEADDR
    i64 sbase = (s0 & 0xfff) + O.first, mout = O.mout;   // ablation (timing only): all tiles hit an L2-resident window
#else
    i64 sbase = s0 + O.first, mout = O.mout;
#endif
    opaque(sbase); opaque(mout);   // recompute the column bases per chain (2 IMADs) instead of keeping 2(S-1) of them live
    const i64 g0 = (a ? C1::CA1 + mout * C1::CB1 : C1::CA0 + mout * C1::CB0) + sbase * C1::LEN;
    T *mine1 = St::mine(stg, a, si, g0);
    T *mine2 = mine1;
#if SAA_COPY == 3
    bulk_wait_read1();             // the column pair staged two steps ago has left this buffer
    __syncwarp();
#endif
#else
    T *mine1 = stage + (a * kTileSamples + si) * C1::STRIDE;
    T *mine2 = stage + 2 * kTileSamples * C1::STRIDE + (a * kTileSamples + si) * C2::STRIDE;
#endif
    DroneChainState<T, S> c1{T(0), dtm}, c2{T(0), dtm};   // d(p,v)_{J+1}/du_J = (0, dt/m)
#pragma unroll
    for (int k = J + 1; k < S; ++k) {
#ifdef SAA_ABL_NOCHAIN
      continue;   // ablation (timing only): no chain math, no staging stores
#endif
      drone_chain_step<T, S, J>(k, P, A22, q2j, ocj, A.dt, a21, mine1, c1);
      if (PAIR && k >= J2 + 1) drone_chain_step<T, S, J2>(k, P, A22, q2j, ocj, A.dt, a21, mine2, c2);
    }
    // sample-mean rows: d p_S/du_J, d v_S/du_J summed over the tile
    {
      const double rp = sum16((double)(active ? c1.sp : T(0)));
      const double rv = sum16((double)(active ? c1.sv : T(0)));
      if (si == 0) { wacc[Rd::FIN_P + a * (S - 1) + J] += rp; wacc[Rd::FIN_V + a * S + J] += rv; }
    }
    if (PAIR) {
      const double rp = sum16((double)(active ? c2.sp : T(0)));
      const double rv = sum16((double)(active ? c2.sv : T(0)));
      if (si == 0) { wacc[Rd::FIN_P + a * (S - 1) + J2] += rp; wacc[Rd::FIN_V + a * S + J2] += rv; }
    }
#if SAA_COPY == 3
    fence_async_smem();
    __syncwarp();
    St::flush(O.Ax, stg, a, si, g0, ns);
#elif SAA_COPY == 4
    __syncwarp();
    {
      const i64 g0x = C1::CA0 + mout * C1::CB0 + sbase * C1::LEN;
      const i64 g0y = C1::CA1 + mout * C1::CB1 + sbase * C1::LEN;
      St::copy_vec(O.Ax, stg, 0, g0x, ns, lane);
      St::copy_vec(O.Ax, stg, 1, g0y, ns, lane);
    }
    __syncwarp();
#else
    __syncwarp();
    i64 sbase = s0 + O.first, mout = O.mout;
    opaque(sbase); opaque(mout);   // recompute the column bases per chain (2 IMADs) instead of keeping 2(S-1) of them live
    copy_run8<T, C1::LEN, C1::STRIDE - C1::LEN>(O.Ax + (C1::CA0 + mout * C1::CB0 + sbase * C1::LEN),
                                                stage, ns * C1::LEN, lane);
    copy_run8<T, C1::LEN, C1::STRIDE - C1::LEN>(O.Ax + (C1::CA1 + mout * C1::CB1 + sbase * C1::LEN),
                                                stage + kTileSamples * C1::STRIDE, ns * C1::LEN, lane);
    if (PAIR) {
      const T *st2 = stage + 2 * kTileSamples * C1::STRIDE;
      copy_run8<T, C2::LEN, C2::STRIDE - C2::LEN>(O.Ax + (C2::CA0 + mout * C2::CB0 + sbase * C2::LEN),
                                                  st2, ns * C2::LEN, lane);
      copy_run8<T, C2::LEN, C2::STRIDE - C2::LEN>(O.Ax + (C2::CA1 + mout * C2::CB1 + sbase * C2::LEN),
                                                  st2 + kTileSamples * C2::STRIDE, ns * C2::LEN, lane);
    }
    __syncwarp();
#endif
    drone_chain_pairs<T, S, J + 1>(A, O, P, A22, q2, oca, a21, dtm, stage, wacc, a, si, lane, s0, ns,
                                   active);
  }
}

// ---- K1: linearize + assemble ------------------------------------------------
template <typename T, int S, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, SAA_BPS)
drone_assemble_kernel(const __grid_constant__ DroneArgs<T, S> A) {
  using Rd = DroneRed<S>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  auto &sm = *reinterpret_cast<DroneSmem<T, S, WARPS> *>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int a = lane >> 4, si = lane & 15;
  T *stage = sm.stage[warp];
  double *wacc = sm.wacc[warp];
  for (int r = lane; r < Rd::N; r += 32) wacc[r] = 0.0;
  __syncwarp();

  DroneOut<T> O{A.Ax, A.M_out, A.first_out};
  {
    i64 ax = (i64)O.Ax;
    opaque(ax); opaque(O.mout); opaque(O.first);
    O.Ax = (T *)ax;
  }
  i64 ub_base = (i64)A.ub, ub_off = A.ub_off;
  opaque(ub_base); opaque(ub_off);
  T *const ub_ptr = (T *)ub_base;
  const i64 ntiles = (A.M + kTileSamples - 1) / kTileSamples;
  const i64 tstride = (i64)gridDim.x * WARPS;
#pragma unroll 1
  for (i64 tile = (i64)blockIdx.x * WARPS + warp; tile < ntiles; tile += tstride) {
    const i64 s0 = tile * kTileSamples;
    const int ns = (int)min((i64)kTileSamples, A.M - s0);
    const bool active = si < ns;
    const i64 s = s0 + (active ? si : 0);
    // ---- all inputs of the tile in flight at once (one DRAM round trip) ----------
    const T *dwz_p = A.dw + 2 * A.Mpad + s, *dwa_p = A.dw + a * A.Mpad + s;
    T dwz[S], dwa[S];
#pragma unroll
    for (int k = 0; k < S; ++k) { if (SAA_PRELOAD) dwz[k] = __ldcs(dwz_p + (i64)k * 3 * A.Mpad); }
#pragma unroll
    for (int k = 0; k < S; ++k) { if (SAA_PRELOAD) dwa[k] = __ldcs(dwa_p + (i64)k * 3 * A.Mpad); }
    T q[3];
#pragma unroll
    for (int o = 0; o < 3; ++o) q[o] = __ldcs(A.q + (o * 2 + a) * A.Mpad + s);
    const T mass = __ldcs(A.mass + s);
    // ... and the next tile's lines on their way into L2 (67 lines of 128 B per tile)
    if (SAA_PREFETCH && tile + tstride < ntiles) {
      const i64 sn = s0 + tstride * kTileSamples;
      for (int r = lane; r < 3 * S + 7; r += 32) {
        const T *pf = r < 3 * S ? A.dw + (i64)r * A.Mpad + sn
                    : r < 3 * S + 6 ? A.q + (i64)(r - 3 * S) * A.Mpad + sn : A.mass + sn;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
      }
    }
    const T inv_m = T(1) / mass;
    const T dt = A.dt, dtm = dt * inv_m, a21 = -A.kp * dtm;
    const T nz = A.noise_c * inv_m;
    const T c2 = T(2) * A.drag;

    // ---------------- z axis: feeds only the sample-mean rows -----------------
#if SAA_Z_INLINE    // z-axis mean rows inside the assemble kernel (default: drone_zmean_kernel)
    {
      T a22z[S];
      T p = A.x0[2], v = A.x0[5], tp = T(0), tv = T(0);
#pragma unroll
      for (int k = 0; k < S; ++k) {
        const T absv = fabs(v);
        const T a22 = T(1) - dtm * (A.kd + c2 * absv);
        a22z[k] = a22;
        const T u = A.us[k * 3 + 2];
        const T acc = (u - A.kp * p - A.kd * v - A.drag * absv * v) * inv_m;
        const T ntp = fma(dt, tv, tp);
        const T ntv = fma(a22, tv, fma(a21, tp, dtm * u));
        const T np_ = fma(dt, v, p);
        v = v + dt * acc + nz * (SAA_PRELOAD ? dwz[k] : dwz_p[(i64)k * 3 * A.Mpad]);
        p = np_; tp = ntp; tv = ntv;
      }
      // linearisation offset of the final rows: -(x_S - x_f) + d x_S/du . u  (:271)
      const T valz = (a == 0) ? (-(p - A.xf[2]) + tp) : (-(v - A.xf[5]) + tv);
      const double rz = sum16((double)(active ? valz : T(0)));
      if (si == 0) wacc[Rd::VAL + (a == 0 ? 2 : 5)] += rz;
      // adjoint pass: lane a=0 carries e_p (row p_z), lane a=1 carries e_v (row v_z)
      T lp = (a == 0) ? T(1) : T(0), lv = (a == 0) ? T(0) : T(1);
#pragma unroll
      for (int j = S - 1; j >= 0; --j) {
        const double r = sum16((double)(active ? lv * dtm : T(0)));
        if (si == 0) {
          if (a == 1) wacc[Rd::FIN_V + 2 * S + j] += r;
          else if (j < S - 1) wacc[Rd::FIN_P + 2 * (S - 1) + j] += r;
        }
        const T nlp = fma(a21, lv, lp);
        const T nlv = fma(a22z[j], lv, dt * lp);
        lp = nlp; lv = nlv;
      }
    }
#endif

    // ---------------- own axis (x or y): rollout + constraint values ----------
    T P[S + 1], A22[S];
    T q2[3], oca[3];
#pragma unroll
    for (int o = 0; o < 3; ++o) {
      q2[o] = T(-2) * A.escale * q[o];
      oca[o] = a ? A.oc[o][1] : A.oc[o][0];
    }
    {
      T p = a ? A.x0[1] : A.x0[0], v = a ? A.x0[4] : A.x0[3], tp = T(0), tv = T(0);
      T zmax = -INFINITY;
      P[0] = p;
#if SAA_GX == 1
      T qo[3], oco[3];                                   // the partner axis' obstacle data
#pragma unroll
      for (int o = 0; o < 3; ++o) {
        qo[o] = __shfl_xor_sync(0xffffffffu, q[o], 16);
        oco[o] = a ? A.oc[o][0] : A.oc[o][1];
      }
#endif
#if SAA_COPY >= 3
      using StU = Stager<T, 3 * S>;
      const i64 gu = ub_off + s0 * (3 * S);
      T *ubrow = StU::mine(stage, 0, si, gu);
#if SAA_COPY == 3
      bulk_wait_read1();           // buffer 0 was last used by the second-to-last column pair of the previous tile
      __syncwarp();
#endif
#else
      T *ubrow = stage + si * (3 * S + 1);
#endif
#pragma unroll
      for (int k = 0; k < S; ++k) {
        const T absv = fabs(v);
        const T a22 = T(1) - dtm * (A.kd + c2 * absv);
        A22[k] = a22;
        const T u = a ? A.us[k * 3 + 1] : A.us[k * 3];
        const T acc = (u - A.kp * p - A.kd * v - A.drag * absv * v) * inv_m;
        const T ntp = fma(dt, tv, tp);
        const T ntv = fma(a22, tv, fma(a21, tp, dtm * u));
        const T np_ = fma(dt, v, p);
        v = v + dt * acc + nz * (SAA_PRELOAD ? dwa[k] : dwa_p[(i64)k * 3 * A.Mpad]);
        p = np_; tp = ntp; tv = ntv;
        P[k + 1] = p;
#if SAA_GX == 1
        // one exchange of (p, tangent) per step instead of two per (step, obstacle)
        const T po = __shfl_xor_sync(0xffffffffu, p, 16), tpo = __shfl_xor_sync(0xffffffffu, tp, 16);
#endif
#pragma unroll
        for (int o = 0; o < 3; ++o) {
          const T d = p - oca[o];
          const T w = q[o] * d * d;                       // own-axis part of 1 - g
          const T e = fma(T(-2) * q[o] * d, tp, w);       // own-axis part of 1 - g + grad g . u
#if SAA_GX == 1
          const T d2 = po - oco[o];
          const T w2 = qo[o] * d2 * d2;
          const T e2 = fma(T(-2) * qo[o] * d2, tpo, w2);
          const T wsum = a ? w2 + w : w + w2;             // x part + y part in both lanes (same rounding)
          const T esum = a ? e2 + e : e + e2;
#else
          const T wsum = w + __shfl_xor_sync(0xffffffffu, w, 16);
          const T esum = e + __shfl_xor_sync(0xffffffffu, e, 16);
#endif
          zmax = fmax(zmax, T(1) - wsum);
          if (a == (k & 1))                               // the two lanes of a sample share the stores
            ubrow[o * S + k] = fma(esum - T(1), A.ubscale, -A.ubpad);
        }
      }
      const T valp = -(p - (a ? A.xf[1] : A.xf[0])) + tp, valv = -(v - (a ? A.xf[4] : A.xf[3])) + tv;
      const double rp = sum16((double)(active ? valp : T(0)));
      const double rv = sum16((double)(active ? valv : T(0)));
      if (si == 0) { wacc[Rd::VAL + a] += rp; wacc[Rd::VAL + 3 + a] += rv; }
      if (A.Z != nullptr && a == 0 && active) A.Z[s] = zmax - A.ztol;
#if SAA_COPY == 3
      fence_async_smem();
      __syncwarp();
      if (ub_ptr != nullptr && a == 0) StU::flush(ub_ptr, stage, 0, si, gu, ns);
#elif SAA_COPY == 4
      __syncwarp();
      if (ub_ptr != nullptr) StU::copy_vec(ub_ptr, stage, 0, gu, ns, lane);
      __syncwarp();
#else
      __syncwarp();
      if (ub_ptr != nullptr)
        copy_run8<T, 3 * S, 1>(ub_ptr + ub_off + s0 * (3 * S), stage, ns * 3 * S, lane);
      __syncwarp();
#endif
    }

    // ---------------- sensitivity chains, two CSC column pairs per pass ---------
    drone_chain_pairs<T, S, 0>(A, O, P, A22, q2, oca, a21, dtm, stage, wacc, a, si, lane, s0, ns,
                               active);
  }

#if SAA_COPY == 3
  bulk_wait_all();
#endif
  // ---------------- per-block partial sums (fixed order => deterministic) -----
  __syncthreads();
  for (int r = threadIdx.x; r < Rd::N; r += WARPS * 32) {
    double acc = 0.0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) acc += sm.wacc[w][r];
    A.partials[(i64)blockIdx.x * Rd::N + r] = acc;
  }
}

// ---- z-axis rows of the sample mean (separate, tiny kernel) ---------------------------
// The z axis never enters the planar obstacle rows; it only contributes d x_S^z / d u^z and the
// z linearisation offsets to the sample-mean rows.  One thread per sample: rollout, then one
// adjoint sweep carrying e_p and e_v; the 2S+1 sums stay in registers over the thread's samples
// and are reduced once at the end.  Reads 8 + 8S bytes per sample.
template <typename T, int S, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
drone_zmean_kernel(const __grid_constant__ DroneArgs<T, S> A, double *__restrict__ partials) {
  using Rd = DroneRed<S>;
  __shared__ double red[WARPS][2 * S + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double accp[S - 1], accv[S], valp = 0.0, valv = 0.0;
#pragma unroll
  for (int j = 0; j < S - 1; ++j) accp[j] = 0.0;
#pragma unroll
  for (int j = 0; j < S; ++j) accv[j] = 0.0;
  const T dt = A.dt, c2 = T(2) * A.drag;
  for (i64 s = (i64)blockIdx.x * blockDim.x + threadIdx.x; s < A.M; s += (i64)gridDim.x * blockDim.x) {
    T dwz[S];
#pragma unroll
    for (int k = 0; k < S; ++k) dwz[k] = __ldcs(A.dw + (i64)(k * 3 + 2) * A.Mpad + s);
    const T inv_m = T(1) / __ldcs(A.mass + s);
    const T dtm = dt * inv_m, a21 = -A.kp * dtm, nz = A.noise_c * inv_m;
    T a22z[S];
    T p = A.x0[2], v = A.x0[5], tp = T(0), tv = T(0);
#pragma unroll
    for (int k = 0; k < S; ++k) {
      const T absv = fabs(v);
      const T a22 = T(1) - dtm * (A.kd + c2 * absv);
      a22z[k] = a22;
      const T u = A.us[k * 3 + 2];
      const T acc = (u - A.kp * p - A.kd * v - A.drag * absv * v) * inv_m;
      const T ntp = fma(dt, tv, tp);
      const T ntv = fma(a22, tv, fma(a21, tp, dtm * u));
      const T np_ = fma(dt, v, p);
      v = v + dt * acc + nz * dwz[k];
      p = np_; tp = ntp; tv = ntv;
    }
    valp += (double)(-(p - A.xf[2]) + tp);      // linearisation offsets (:271)
    valv += (double)(-(v - A.xf[5]) + tv);
    T pp = T(1), pv = T(0), vp = T(0), vv = T(1);   // adjoints of p_S (pp, pv) and v_S (vp, vv)
#pragma unroll
    for (int j = S - 1; j >= 0; --j) {
      if (j < S - 1) accp[j] += (double)(pv * dtm);
      accv[j] += (double)(vv * dtm);
      const T npp = fma(a21, pv, pp), npv = fma(a22z[j], pv, dt * pp);
      const T nvp = fma(a21, vv, vp), nvv = fma(a22z[j], vv, dt * vp);
      pp = npp; pv = npv; vp = nvp; vv = nvv;
    }
  }
#pragma unroll
  for (int j = 0; j < S - 1; ++j) { const double r = sum32(accp[j]); if (lane == 0) red[warp][j] = r; }
#pragma unroll
  for (int j = 0; j < S; ++j) { const double r = sum32(accv[j]); if (lane == 0) red[warp][S - 1 + j] = r; }
  { const double r = sum32(valp); if (lane == 0) red[warp][2 * S - 1] = r; }
  { const double r = sum32(valv); if (lane == 0) red[warp][2 * S] = r; }
  __syncthreads();
  // one full row of partials per block: zeros except the z slots
  double *row = partials + (i64)blockIdx.x * Rd::N;
  for (int r = threadIdx.x; r < Rd::N; r += WARPS * 32) {
    int src = -1;
    if (r >= Rd::FIN_P + 2 * (S - 1) && r < Rd::FIN_P + 3 * (S - 1)) src = r - (Rd::FIN_P + 2 * (S - 1));
    else if (r >= Rd::FIN_V + 2 * S && r < Rd::FIN_V + 3 * S) src = S - 1 + r - (Rd::FIN_V + 2 * S);
    else if (r == Rd::VAL + 2) src = 2 * S - 1;
    else if (r == Rd::VAL + 5) src = 2 * S;
    double acc = 0.0;
    if (src >= 0)
      for (int w = 0; w < WARPS; ++w) acc += red[w][src];
    row[r] = acc;
  }
}

// ---- finalize: sum block partials in block order, divide, scatter ------------
// sums_out (optional): the plain sums, for the multi-GPU all-reduce.
// one warp per slot: lanes stride over the blocks, then a fixed-order butterfly
template <typename T>
__global__ void reduce_partials_kernel(const double *__restrict__ partials, int nblocks, int n,
                                       double *__restrict__ sums_out) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= n) return;
  double acc = 0.0;
  for (int b = lane; b < nblocks; b += 32) acc += partials[(i64)b * n + r];
  acc = sum32(acc);
  if (lane == 0) sums_out[r] = acc;
}

template <typename T>
__global__ void scatter_means_kernel(const double *__restrict__ sums, double inv_M, int n_entries,
                                     const i64 *__restrict__ fin_off, int n_fin, T *Ax, T *l,
                                     T *u) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_entries) {
    Ax[fin_off[r]] = (T)(sums[r] * inv_M);
  } else if (r < n_entries + n_fin) {
    const T v = (T)(sums[r] * inv_M);
    l[r - n_entries] = v;          // equality by equal bounds (:272-273)
    u[r - n_entries] = v;
  }
}

// ---- K4 / K5: rollout only, optional trajectory output and CVaR terms --------
template <typename T, int S> struct DroneRollArgs {
  const T *mass, *dw, *q;
  i64 M, Mpad;
  T us[S * 3];
  T dt, noise_c, drag, kp, kd;
  T x0[6];
  T oc[3][2];
  T *Xs;            // (M, S+1, 6) or nullptr
  T *Z;             // (M) or nullptr
  T ztol, t_risk, sat_tol;
  double *partials; // [gridDim.x][3] or nullptr: sum max(Z-t,0), count Z<=sat_tol, max Z
};

// one thread per sample; trajectories are staged per warp so that the
// (S+1)*6 contiguous doubles of each sample leave as full coalesced runs.
template <typename T, int S, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
drone_rollout_kernel(const __grid_constant__ DroneRollArgs<T, S> A) {
  constexpr int ROW = (S + 1) * 6, STRIDE = ROW | 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *stage = reinterpret_cast<T *>(smem_raw) + (threadIdx.x >> 5) * 32 * STRIDE;
  __shared__ double red[WARPS][3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double acc_excess = 0.0, acc_sat = 0.0, acc_max = -INFINITY;
  const i64 ntiles = (A.M + 31) / 32;
  for (i64 tile = (i64)blockIdx.x * WARPS + warp; tile < ntiles; tile += (i64)gridDim.x * WARPS) {
    const i64 s0 = tile * 32;
    const int ns = (int)min((i64)32, A.M - s0);
    const bool active = lane < ns;
    const i64 s = s0 + (active ? lane : 0);
    const T inv_m = T(1) / A.mass[s];
    const T dt = A.dt, nz = A.noise_c * inv_m;
    T p[3], v[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) { p[a] = A.x0[a]; v[a] = A.x0[3 + a]; }
    T qx[3], qy[3];
#pragma unroll
    for (int o = 0; o < 3; ++o) { qx[o] = A.q[(o * 2) * A.Mpad + s]; qy[o] = A.q[(o * 2 + 1) * A.Mpad + s]; }
    T zmax = -INFINITY;
    T *mine = stage + lane * STRIDE;
    if (A.Xs != nullptr) {
#pragma unroll
      for (int a = 0; a < 3; ++a) { mine[a] = p[a]; mine[3 + a] = v[a]; }
    }
    // all noise increments of the sample in flight at once (one DRAM round trip per tile)
    T dw[S * 3];
#pragma unroll
    for (int r = 0; r < S * 3; ++r) dw[r] = __ldcs(A.dw + (i64)r * A.Mpad + s);
#pragma unroll
    for (int k = 0; k < S; ++k) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const T absv = fabs(v[a]);
        const T acc = (A.us[k * 3 + a] - A.kp * p[a] - A.kd * v[a] - A.drag * absv * v[a]) * inv_m;
        const T np_ = fma(dt, v[a], p[a]);
        v[a] = v[a] + dt * acc + nz * dw[k * 3 + a];
        p[a] = np_;
      }
      if (A.Xs != nullptr) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { mine[(k + 1) * 6 + a] = p[a]; mine[(k + 1) * 6 + 3 + a] = v[a]; }
      }
#pragma unroll
      for (int o = 0; o < 3; ++o) {
        const T dx = p[0] - A.oc[o][0], dy = p[1] - A.oc[o][1];
        zmax = fmax(zmax, T(1) - (qx[o] * dx * dx + qy[o] * dy * dy));
      }
    }
    const T Zi = zmax - A.ztol;
    if (A.Z != nullptr && active) A.Z[s] = Zi;
    if (active) {
      acc_excess += (double)fmax(Zi - A.t_risk, T(0));
      acc_sat += (Zi <= A.sat_tol) ? 1.0 : 0.0;
      acc_max = fmax(acc_max, (double)Zi);
    }
    if (A.Xs != nullptr) {
      __syncwarp();
      copy_run<T, ROW, STRIDE>(A.Xs + s0 * ROW, stage, ns * ROW, lane);
      __syncwarp();
    }
  }
  if (A.partials != nullptr) {
    acc_excess = sum32(acc_excess); acc_sat = sum32(acc_sat); acc_max = max32(acc_max);
    if (lane == 0) { red[warp][0] = acc_excess; red[warp][1] = acc_sat; red[warp][2] = acc_max; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double e = 0.0, c = 0.0, m = -INFINITY;
      for (int w = 0; w < WARPS; ++w) { e += red[w][0]; c += red[w][1]; m = fmax(m, red[w][2]); }
      A.partials[(i64)blockIdx.x * 3 + 0] = e;
      A.partials[(i64)blockIdx.x * 3 + 1] = c;
      A.partials[(i64)blockIdx.x * 3 + 2] = m;
    }
  }
}

// out3 = [sum excess, count satisfied, max Z] over blocks, fixed order
__global__ void reduce_cvar_kernel(const double *__restrict__ partials, int nblocks,
                                   double *__restrict__ out3) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double e = 0.0, c = 0.0, m = -INFINITY;
    for (int b = 0; b < nblocks; ++b) {
      e += partials[b * 3]; c += partials[b * 3 + 1]; m = fmax(m, partials[b * 3 + 2]);
    }
    out3[0] = e; out3[1] = c; out3[2] = m;
  }
}

// ---- repack the reference-layout sample set into the kernel's SoA layout -----
template <typename T>
__global__ void drone_pack_kernel(const double *__restrict__ masses, const double *__restrict__ DWs,
                                  const double *__restrict__ obs_Qs, i64 M, i64 Mpad, int S, T *mass,
                                  T *dw, T *q) {
  const i64 s = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= Mpad) return;
  const i64 src = s < M ? s : M - 1;     // pad with a valid sample (never written out)
  mass[s] = (T)masses[src];
  for (int k = 0; k < S; ++k)
    for (int a = 0; a < 3; ++a) dw[(k * 3 + a) * Mpad + s] = (T)DWs[(src * S + k) * 6 + 3 + a];
  for (int o = 0; o < 3; ++o)
    for (int a = 0; a < 2; ++a) q[(o * 2 + a) * Mpad + s] = (T)obs_Qs[((src * 3 + o) * 3 + a) * 3 + a];
}

}  // namespace saa
