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
// of those samples, lanes 16-31 the y axis (the z axis only feeds the sample-mean
// rows: drone_axis_mean_kernel).  For each control step j the lane runs the
// sensitivity chain k = j+1..S in registers and stages the 3*(S-1-j) CSC entries
// of its sample for column (j, axis) in shared memory; the warp then streams the
// two column sub-runs (16 samples x 3*(S-1-j) contiguous doubles each) to global
// memory, consecutive lanes on consecutive 16-byte chunks.
//
// Three modes of the same kernel:
//   FULL    rollout -> chains -> CSC entries                       (single GPU, or a rank's own block)
//   FACTOR  rollout -> chains -> FACTORED record of the block: the sensitivities d p_k/d u_j
//           (190 per sample and axis) and the trajectory (p_1..p_S, -2 escale Q: 23 per sample
//           and axis) instead of their 570 products -- 3.4 KB instead of 9.1 KB per sample,
//           which is what has to cross NVLink in the multi-GPU gather
//   EXPAND  factored record -> CSC entries (run by the rank that owns the matrix); the
//           products are formed by the same instructions as in FULL, so the result is bitwise
//           identical to a single-GPU run
#pragma once
#include "saa_common.cuh"

// ---- tuning switches (defaults = the measured best; history in profiles/README.md) ----
#ifndef SAA_COPY
#define SAA_COPY 4         // 4: 16-byte shared loads -> 16-byte streaming stores (1.93 ms at M = 1e6);
                           // 3: TMA bulk stores (cp.async.bulk with an L2 evict-first hint), double
                           // buffered: correct, 2.6 ms (half the warps fit); single buffered with all 12
                           // warps it ties at 1.97 ms -- the store pattern, not the copy mechanism, is
                           // what bounds the kernel (profiles/README.md)
#endif
#ifndef SAA_BPS
#define SAA_BPS 2          // resident blocks per SM the kernel is compiled for
#endif

namespace saa {

enum { DRONE_FULL = 0, DRONE_FACTOR = 1, DRONE_EXPAND = 2 };

template <int S> struct DroneRed {
  static constexpr int FIN_P = 0;                 // + a*(S-1) + j   (a<3, j<S-1)
  static constexpr int FIN_V = 3 * (S - 1);       // + a*S + j       (a<3, j<S)
  static constexpr int VAL = FIN_V + 3 * S;       // + r             (r<6)
  static constexpr int N = VAL + 6;
};

// rows of the factored trajectory record per (axis): p_1..p_S, then the 3 scaled Q entries
template <int S> struct DroneFac { static constexpr int ROWS = S + 3; };

template <typename T, typename TO, int S> struct DroneArgs {
  const T *mass, *dw, *q;   // packed: mass[M]; dw[(k*3+a)*Mpad+s]; q[(o*2+a)*Mpad+s]
  i64 M, Mpad;
  T us[S * 3];
  T dt, noise_c, drag, kp, kd;
  T x0[6], xf[6];
  T oc[3][2];
  T escale;                 // multiplier (x relaxation scale) applied to the Jacobian entries
  T ubscale, ubpad;         // upper bound = ubscale * (-g + grad g . u) - ubpad
  T ztol;                   // Z_i = max g - ztol
  TO *Ax;                  // TO = storage type of the outputs (double | float); T = arithmetic type
  // Sub-run of local sample 0 in u column (j, a): Ax + CA(j,a) + M_out*CB(j,a) + first_out*LEN(j)
  // with compile-time CA, CB (DroneChain); the host checks them against Layout::run_start.
  i64 M_out, first_out;
  TO *ub;                   // base of the upper-bound vector (nullptr: skip)
  i64 ub_off;               // row of local sample 0's first sample row
  TO *Z;                    // per-sample max constraint (nullptr: skip)
  double *partials;         // [gridDim.x][DroneRed<S>::N]
  // factored record (FACTOR writes, EXPAND reads), laid out for M_out samples:
  //   fsp[M_out * FB(j,a) + sample * (S-1-j) + kk] = d p_{j+2+kk} / d u_{j,a}
  //   fp[(a * (S+3) + r) * M_out + sample]         = p_{r+1} (r < S) | -2 escale Q_o,aa (r = S+o)
  TO *fsp, *fp;
  i64 s_begin;              // EXPAND: first sample (of the M_out geometry) to expand; M = count
};

// ---- column geometry --------------------------------------------------------------------
// Column (J, axis) of a sample holds rows k = J+2..S for each of the 3 obstacles:
// LEN = 3 (S-1-J) contiguous values.
template <int S, int J> struct DroneChain {
  static constexpr int L = S - 1 - J;      // rows k = J+2..S carry an entry
  static constexpr int LEN = 3 * L;        // per-sample run length in the CSC column
  // CSC position of sample 0's run of column c = 3J + a in a matrix with M samples is
  // CAa + M * CBa: every earlier control step holds 6 final-row + 3 control-row
  // entries and 2 columns of 3(S-1-j') values per sample (layout.cuh, Layout::build)
  static constexpr i64 CA0 = 9 * J + 2, CA1 = CA0 + 3;
  static constexpr i64 CB0 = 6 * J * (S - 1) - 3 * J * (J - 1), CB1 = CB0 + 3 * (S - 1 - J);
  // factored record: column (J, a) starts at M * FBa, L values per sample
  static constexpr i64 FB0 = 2 * J * (S - 1) - J * (J - 1), FB1 = FB0 + L;
};

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

// Launch constants the copy-out needs, read from the parameter bank ONCE and pinned in
// registers (an LDC in front of every column store showed up as long-scoreboard stalls).
template <typename T> struct DroneOut {
  T *Ax, *fsp;
  i64 mout, first;
};

// stage -> global for one column pair
template <typename T, int LEN>
__device__ __forceinline__ void drone_flush(T *base, T *stg, i64 g0x, i64 g0y, int a, int si, int ns,
                                            int lane) {
  using St = Stager<T, LEN>;
#if SAA_COPY == 3
  fence_async_smem();
  __syncwarp();
  St::flush(base, stg, a, si, a ? g0y : g0x, ns);
#else
  __syncwarp();
  St::copy_vec(base, stg, 0, g0x, ns, lane);
  St::copy_vec(base, stg, 1, g0y, ns, lane);
  __syncwarp();
#endif
}

#ifndef SAA_LINE_OWN
#define SAA_LINE_OWN 0     // 1: every 128-byte line of a column is written whole by ONE warp (see drone_flush_own)
#endif

// What a warp knows about its neighbours in the block (consecutive tiles = adjacent runs)
template <typename T> struct DroneNbr {
  const T *next;      // staging buffer of warp + 1 (nullptr: no next tile in this block)
  int ns_next;        // its valid samples
  bool has_prev;      // warp - 1 holds the preceding tile
};

// Line-ownership copy-out.  The runs of consecutive tiles are adjacent in the column but start at
// an arbitrary 8-byte phase, so every run shares its first and last 128-byte line with a
// neighbour; a line written in two pieces by two warps costs ~3.4 full-line writes
// (profiles/README.md, tools/wbw3.cu).  Here the warps of a block move through the columns in
// lockstep (two block barriers per column pair): once all have staged, a warp writes its own run
// from the first line boundary on AND completes its last line with the head elements it reads from
// the next warp's staging buffer; only the two ends of the block's 96-sample run stay partial.
template <typename T, int LEN>
__device__ __forceinline__ void drone_flush_own(T *base, const T *stg, const DroneNbr<T> &N, i64 g0x, i64 g0y,
                                                int ns, int lane) {
  using St = Stager<T, LEN>;
  constexpr int VEC = St::VEC, LINE = 128 / (int)sizeof(T);
  __syncthreads();                       // every warp of the block has staged this column pair
  const int n = ns * LEN;
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    const i64 g0 = a ? g0y : g0x;
    const int t0 = (int)(g0 & (LINE - 1)), need0 = (LINE - t0) & (LINE - 1);
    const int e0 = (N.has_prev && t0 != 0 && n >= need0) ? need0 : 0;      // the previous warp completes that line
    const i64 G = g0 + n;
    const int t1 = (int)(G & (LINE - 1)), need1 = LINE - t1;
    const bool own_tail = N.next != nullptr && t1 != 0 && N.ns_next * LEN >= need1 && n - t1 >= e0;
    const int e1 = own_tail ? n - t1 : n;
    St::copy_range(base, stg, a, g0, e0, e1, lane);
    if (own_tail && lane < LINE / VEC) {
      const int off = (int)(g0 & (VEC - 1)), offn = (int)(G & (VEC - 1));
      T tmp[VEC];
#pragma unroll
      for (int q = 0; q < VEC; ++q) {
        const int ge = lane * VEC + q;                                       // position inside the line
        tmp[q] = ge < t1 ? *St::elem(stg, a, n - t1 + ge, off) : *St::elem(N.next, a, ge - t1, offn);
      }
      __stcs(reinterpret_cast<int4 *>(base + (G - t1) + lane * VEC), *reinterpret_cast<const int4 *>(tmp));
    }
  }
  __syncthreads();                       // the staging buffers may be overwritten
}

// barriers of drone_flush_own for a warp of the block that has no tile in this round
template <int S, int J>
__device__ __forceinline__ void drone_idle_barriers() {
  if constexpr (J < S - 1) {
    __syncthreads();
    __syncthreads();
    drone_idle_barriers<S, J + 1>();
  }
}

// ---- sensitivity chains, one CSC column pair (x and y column of control step J) per pass ----
template <typename T, typename TO, int S, int J, int MODE>
__device__ __forceinline__ void drone_chains(const DroneArgs<T, TO, S> &A, const DroneOut<TO> &O,
                                             const T (&P)[S + 1], const T (&A22)[S],
                                             const T (&q2)[3], const T (&oca)[3], T a21, T dtm,
                                             TO *stage, double *wacc, int a, int si, int lane, i64 s0,
                                             int ns, bool active, const DroneNbr<TO> &N) {
  using Rd = DroneRed<S>;
  if constexpr (J >= S - 1) {
    // last control step: no sample rows, only d v_S/du = dt/m enters the mean rows
    if constexpr (MODE != DRONE_EXPAND) {
      const double rv = sum16((double)(active ? dtm : T(0)));
      if (si == 0) wacc[Rd::FIN_V + a * S + (S - 1)] += rv;
    }
  } else {
    using C = DroneChain<S, J>;
    constexpr int SLEN = (MODE == DRONE_FACTOR) ? C::L : C::LEN;   // staged values per sample
    using St = Stager<TO, SLEN>;
    static_assert(St::SIZE <= DroneSmem<TO, S, 1>::HALF, "staging buffer too small");
#if SAA_COPY == 3
    // upper bounds went through buffer 0, chain 0 takes buffer 1, chain 1 buffer 0, ...
    TO *const stg = stage + ((J & 1) ? 0 : DroneSmem<TO, S, 1>::HALF);
#else
    TO *const stg = stage;
#endif
    // optimisation barriers: keep per-chain coefficient math from being hoisted (common
    // subexpressions across the unrolled chains would cost ~60 live doubles), and recompute the
    // column bases per chain (2 IMADs) instead of keeping 2(S-1) of them live
    T q2j[3] = {q2[0], q2[1], q2[2]}, ocj[3] = {-q2[0] * oca[0], -q2[1] * oca[1], -q2[2] * oca[2]};
    opaque(q2j[0]); opaque(q2j[1]); opaque(q2j[2]);
    opaque(ocj[0]); opaque(ocj[1]); opaque(ocj[2]);
    i64 sbase = s0 + O.first, mout = O.mout;
    opaque(sbase); opaque(mout);
    const i64 f0x = mout * C::FB0 + sbase * C::L, f0y = mout * C::FB1 + sbase * C::L;     // factored runs
    const i64 g0x = C::CA0 + mout * C::CB0 + sbase * C::LEN, g0y = C::CA1 + mout * C::CB1 + sbase * C::LEN;
    TO *mine = St::mine(stg, a, si, MODE == DRONE_FACTOR ? (a ? f0y : f0x) : (a ? g0y : g0x));
#if SAA_COPY == 3
    bulk_wait_read1();             // the column pair staged two steps ago has left this buffer
    __syncwarp();
#endif
    T sp = T(0), sv = dtm;         // d(p,v)_{J+1}/du_J = (0, dt/m)
    const TO *fin = O.fsp + (a ? f0y : f0x) + (i64)(active ? si : 0) * C::L;   // EXPAND: this sample's chain
#pragma unroll
    for (int k = J + 1; k < S; ++k) {
      const int kk = k - J - 1;
      if constexpr (MODE == DRONE_EXPAND) {
        sp = (T)fin[kk];           // d p_{k+1}/du_J computed by the rank that owns the sample
      } else {
        const T nsp = fma(A.dt, sv, sp);
        const T nsv = fma(A22[k], sv, a21 * sp);
        sp = nsp; sv = nsv;        // now d(p,v)_{k+1} / du_J
      }
      if constexpr (MODE == DRONE_FACTOR) {
        mine[kk] = (TO)sp;
      } else {
#pragma unroll
        for (int o = 0; o < 3; ++o) {
          const T coef = fma(q2j[o], P[k + 1], ocj[o]);   // escale * d g[o,k+1]/dp = q2 (p - c)
          mine[o * C::L + kk] = (TO)(coef * sp);
        }
      }
    }
    if constexpr (MODE != DRONE_EXPAND) {
      // sample-mean rows: d p_S/du_J, d v_S/du_J summed over the tile
      const double rp = sum16((double)(active ? sp : T(0)));
      const double rv = sum16((double)(active ? sv : T(0)));
      if (si == 0) { wacc[Rd::FIN_P + a * (S - 1) + J] += rp; wacc[Rd::FIN_V + a * S + J] += rv; }
    }
    if constexpr (MODE == DRONE_FACTOR) drone_flush<TO, SLEN>(O.fsp, stg, f0x, f0y, a, si, ns, lane);
#if SAA_LINE_OWN && SAA_COPY != 3
    else drone_flush_own<TO, SLEN>(O.Ax, stg, N, g0x, g0y, ns, lane);
#else
    else drone_flush<TO, SLEN>(O.Ax, stg, g0x, g0y, a, si, ns, lane);
#endif
    drone_chains<T, TO, S, J + 1, MODE>(A, O, P, A22, q2, oca, a21, dtm, stage, wacc, a, si, lane, s0, ns,
                                    active, N);
  }
}

// ---- K1: linearize + assemble ------------------------------------------------
template <typename T, typename TO, int S, int WARPS, int MODE>
__global__ void __launch_bounds__(WARPS * 32, SAA_BPS)
drone_assemble_kernel(const __grid_constant__ DroneArgs<T, TO, S> A) {
  using Rd = DroneRed<S>;
  constexpr int FR = DroneFac<S>::ROWS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  auto &sm = *reinterpret_cast<DroneSmem<TO, S, WARPS> *>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int a = lane >> 4, si = lane & 15;
  TO *stage = sm.stage[warp];
  double *wacc = sm.wacc[warp];
  for (int r = lane; r < Rd::N; r += 32) wacc[r] = 0.0;
  __syncwarp();

  DroneOut<TO> O{A.Ax, A.fsp, A.M_out, A.first_out};
  {
    i64 ax = (i64)O.Ax, fs = (i64)O.fsp;
    opaque(ax); opaque(fs); opaque(O.mout); opaque(O.first);
    O.Ax = (TO *)ax; O.fsp = (TO *)fs;
  }
  if (MODE == DRONE_EXPAND) O.first = A.s_begin;
  i64 ub_base = (i64)A.ub, ub_off = A.ub_off;
  opaque(ub_base); opaque(ub_off);
  TO *const ub_ptr = (TO *)ub_base;
  const i64 ntiles = (A.M + kTileSamples - 1) / kTileSamples;
  const i64 tstride = (i64)gridDim.x * WARPS;
#if SAA_LINE_OWN && SAA_COPY != 3
  constexpr bool kOwn = MODE != DRONE_FACTOR;        // lockstep blocks: every warp runs every round
#else
  constexpr bool kOwn = false;
#endif
#pragma unroll 1
  for (i64 tile0 = (i64)blockIdx.x * WARPS; tile0 < ntiles; tile0 += tstride) {
    const i64 tile = tile0 + warp;
    if (tile >= ntiles) {
      if constexpr (kOwn) { drone_idle_barriers<S, 0>(); continue; }
      else break;
    }
    const i64 s0 = tile * kTileSamples;
    const int ns = (int)min((i64)kTileSamples, A.M - s0);
    const bool active = si < ns;
    DroneNbr<TO> N{nullptr, 0, false};
    if constexpr (kOwn) {
      N.has_prev = warp > 0;
      if (warp + 1 < WARPS && tile + 1 < ntiles) {
        N.next = sm.stage[warp + 1];
        N.ns_next = (int)min((i64)kTileSamples, A.M - (s0 + kTileSamples));
      }
    }
    const i64 s = s0 + (active ? si : 0);
    T P[S + 1], A22[S];
    T q2[3], oca[3];
#pragma unroll
    for (int o = 0; o < 3; ++o) oca[o] = a ? A.oc[o][1] : A.oc[o][0];
    T dtm = T(0), a21 = T(0);

    if constexpr (MODE == DRONE_EXPAND) {
      // the owner of the matrix re-creates the entries of another rank's samples from the record
      const TO *fp = A.fp + (i64)a * FR * O.mout + (O.first + s);
#pragma unroll
      for (int k = 1; k <= S; ++k) P[k] = (T)__ldcs(fp + (i64)(k - 1) * O.mout);
#pragma unroll
      for (int o = 0; o < 3; ++o) q2[o] = (T)__ldcs(fp + (i64)(S + o) * O.mout);
      P[0] = T(0);
#pragma unroll
      for (int k = 0; k < S; ++k) A22[k] = T(0);
    } else {
      // ---- all inputs of the tile in flight at once (one DRAM round trip) ----------
      const T *dwa_p = A.dw + a * A.Mpad + s;
      T dwa[S];
#pragma unroll
      for (int k = 0; k < S; ++k) dwa[k] = __ldcs(dwa_p + (i64)k * 3 * A.Mpad);
      T q[3];
#pragma unroll
      for (int o = 0; o < 3; ++o) q[o] = __ldcs(A.q + (o * 2 + a) * A.Mpad + s);
      const T inv_m = T(1) / __ldcs(A.mass + s);
      const T dt = A.dt;
      dtm = dt * inv_m; a21 = -A.kp * dtm;
      const T nz = A.noise_c * inv_m;
      const T c2 = T(2) * A.drag;
#pragma unroll
      for (int o = 0; o < 3; ++o) q2[o] = T(-2) * A.escale * q[o];

      // ---------------- own axis (x or y): rollout + constraint values ----------
      T p = a ? A.x0[1] : A.x0[0], v = a ? A.x0[4] : A.x0[3], tp = T(0), tv = T(0);
      T zmax = -INFINITY;
      P[0] = p;
      using StU = Stager<TO, 3 * S>;
      const i64 gu = ub_off + s0 * (3 * S);
      TO *ubrow = StU::mine(stage, 0, si, gu);
#if SAA_COPY == 3
      bulk_wait_read1();           // buffer 0 was last used by the second-to-last column pair of the previous tile
      __syncwarp();
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
        v = v + dt * acc + nz * dwa[k];
        p = np_; tp = ntp; tv = ntv;
        P[k + 1] = p;
#pragma unroll
        for (int o = 0; o < 3; ++o) {
          const T d = p - oca[o];
          const T w = q[o] * d * d;                       // own-axis part of 1 - g
          const T e = fma(T(-2) * q[o] * d, tp, w);       // own-axis part of 1 - g + grad g . u
          const T wsum = w + __shfl_xor_sync(0xffffffffu, w, 16);
          const T esum = e + __shfl_xor_sync(0xffffffffu, e, 16);
          zmax = fmax(zmax, T(1) - wsum);
          if (a == (k & 1))                               // the two lanes of a sample share the stores
            ubrow[o * S + k] = (TO)fma(esum - T(1), A.ubscale, -A.ubpad);
        }
      }
      // linearisation offset of the final rows: -(x_S - x_f) + d x_S/du . u  (:271)
      const T valp = -(p - (a ? A.xf[1] : A.xf[0])) + tp, valv = -(v - (a ? A.xf[4] : A.xf[3])) + tv;
      const double rp = sum16((double)(active ? valp : T(0)));
      const double rv = sum16((double)(active ? valv : T(0)));
      if (si == 0) { wacc[Rd::VAL + a] += rp; wacc[Rd::VAL + 3 + a] += rv; }
      if (A.Z != nullptr && a == 0 && active) A.Z[s] = (TO)(zmax - A.ztol);
#if SAA_COPY == 3
      fence_async_smem();
      __syncwarp();
      if (ub_ptr != nullptr && a == 0) StU::flush(ub_ptr, stage, 0, si, gu, ns);
#else
      __syncwarp();
      if (ub_ptr != nullptr) StU::copy_vec(ub_ptr, stage, 0, gu, ns, lane);
      __syncwarp();
#endif
      if constexpr (MODE == DRONE_FACTOR) {
        // trajectory part of the factored record (coalesced: 16 samples per 128-byte line)
        if (active) {
          TO *fp = A.fp + (i64)a * FR * O.mout + (O.first + s);
#pragma unroll
          for (int k = 1; k <= S; ++k) st_stream(fp + (i64)(k - 1) * O.mout, (TO)P[k]);
#pragma unroll
          for (int o = 0; o < 3; ++o) st_stream(fp + (i64)(S + o) * O.mout, (TO)q2[o]);
        }
      }
    }

    // ---------------- sensitivity chains, one CSC column pair per control step -
    drone_chains<T, TO, S, 0, MODE>(A, O, P, A22, q2, oca, a21, dtm, stage, wacc, a, si, lane, s0, ns, active, N);
  }

#if SAA_COPY == 3
  bulk_wait_all();
#endif
  if constexpr (MODE != DRONE_EXPAND) {
    // ---------------- per-block partial sums (fixed order => deterministic) -----
    __syncthreads();
    for (int r = threadIdx.x; r < Rd::N; r += WARPS * 32) {
      double acc = 0.0;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) acc += sm.wacc[w][r];
      A.partials[(i64)blockIdx.x * Rd::N + r] = acc;
    }
  }
}

// ---- sample-mean rows of ONE axis (separate, tiny kernel) ------------------------------
// The z axis never enters the planar obstacle rows; it only contributes d x_S^z / d u^z and the
// z linearisation offsets to the sample-mean rows, so the assemble launch runs this kernel with
// axis = 2.  The means-only pass (saa_linearize_means: expectation rows and Z_i without the
// matrix, for the tail-reduced subproblem) runs it for all three axes.  One thread per sample:
// rollout, then one adjoint sweep carrying e_p and e_v; the 2S+1 sums stay in registers over the
// thread's samples and are reduced once at the end.  Reads 8 + 8S bytes per sample.
template <typename T, typename TO, int S, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
drone_axis_mean_kernel(const __grid_constant__ DroneArgs<T, TO, S> A, int axis, double *__restrict__ partials) {
  using Rd = DroneRed<S>;
  __shared__ double red[WARPS][2 * S + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double accp[S - 1], accv[S], valp = 0.0, valv = 0.0;
#pragma unroll
  for (int j = 0; j < S - 1; ++j) accp[j] = 0.0;
#pragma unroll
  for (int j = 0; j < S; ++j) accv[j] = 0.0;
  const T dt = A.dt, c2 = T(2) * A.drag;
  const T p0 = A.x0[axis], v0 = A.x0[3 + axis], pf = A.xf[axis], vf = A.xf[3 + axis];
  T ua[S];
#pragma unroll
  for (int k = 0; k < S; ++k) ua[k] = A.us[k * 3 + axis];
  for (i64 s = (i64)blockIdx.x * blockDim.x + threadIdx.x; s < A.M; s += (i64)gridDim.x * blockDim.x) {
    T dwz[S];
#pragma unroll
    for (int k = 0; k < S; ++k) dwz[k] = __ldcs(A.dw + (i64)(k * 3 + axis) * A.Mpad + s);
    const T inv_m = T(1) / __ldcs(A.mass + s);
    const T dtm = dt * inv_m, a21 = -A.kp * dtm, nz = A.noise_c * inv_m;
    T a22z[S];
    T p = p0, v = v0, tp = T(0), tv = T(0);
#pragma unroll
    for (int k = 0; k < S; ++k) {
      const T absv = fabs(v);
      const T a22 = T(1) - dtm * (A.kd + c2 * absv);
      a22z[k] = a22;
      const T u = ua[k];
      const T acc = (u - A.kp * p - A.kd * v - A.drag * absv * v) * inv_m;
      const T ntp = fma(dt, tv, tp);
      const T ntv = fma(a22, tv, fma(a21, tp, dtm * u));
      const T np_ = fma(dt, v, p);
      v = v + dt * acc + nz * dwz[k];
      p = np_; tp = ntp; tv = ntv;
    }
    valp += (double)(-(p - pf) + tp);      // linearisation offsets (:271)
    valv += (double)(-(v - vf) + tv);
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
  // one full row of partials per block: zeros except this axis' slots
  double *row = partials + (i64)blockIdx.x * Rd::N;
  for (int r = threadIdx.x; r < Rd::N; r += WARPS * 32) {
    int src = -1;
    if (r >= Rd::FIN_P + axis * (S - 1) && r < Rd::FIN_P + (axis + 1) * (S - 1)) src = r - (Rd::FIN_P + axis * (S - 1));
    else if (r >= Rd::FIN_V + axis * S && r < Rd::FIN_V + (axis + 1) * S) src = S - 1 + r - (Rd::FIN_V + axis * S);
    else if (r == Rd::VAL + axis) src = 2 * S - 1;
    else if (r == Rd::VAL + 3 + axis) src = 2 * S;
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
template <typename T, typename TO, int S> struct DroneRollArgs {
  const T *mass, *dw, *q;
  i64 M, Mpad;
  T us[S * 3];
  T dt, noise_c, drag, kp, kd;
  T x0[6];
  T oc[3][2];
  TO *Xs;           // (M, S+1, 6) or nullptr
  TO *Z;            // (M) or nullptr
  T ztol, t_risk, sat_tol;
  double *partials; // [gridDim.x][3] or nullptr: sum max(Z-t,0), count Z<=sat_tol, max Z
};

// one thread per sample; trajectories are staged per warp so that the
// (S+1)*6 contiguous doubles of each sample leave as full coalesced runs.
template <typename T, typename TO, int S, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
drone_rollout_kernel(const __grid_constant__ DroneRollArgs<T, TO, S> A) {
  constexpr int ROW = (S + 1) * 6, STRIDE = ROW | 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TO *stage = reinterpret_cast<TO *>(smem_raw) + (threadIdx.x >> 5) * 32 * STRIDE;
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
    TO *mine = stage + lane * STRIDE;
    if (A.Xs != nullptr) {
#pragma unroll
      for (int a = 0; a < 3; ++a) { mine[a] = (TO)p[a]; mine[3 + a] = (TO)v[a]; }
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
        for (int a = 0; a < 3; ++a) { mine[(k + 1) * 6 + a] = (TO)p[a]; mine[(k + 1) * 6 + 3 + a] = (TO)v[a]; }
      }
#pragma unroll
      for (int o = 0; o < 3; ++o) {
        const T dx = p[0] - A.oc[o][0], dy = p[1] - A.oc[o][1];
        zmax = fmax(zmax, T(1) - (qx[o] * dx * dx + qy[o] * dy * dy));
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
