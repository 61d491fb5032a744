// Car linearize + assemble with 32-sample tiles (K2, variant selected by SAA_CAR_TILE=32).
//
// Same recursions and the same pass structure as car_assemble_kernel (car_kernels.cuh) -- bitwise
// the same entries -- but LANE = SAMPLE and each lane advances the chains of BOTH controls:
//   * the per-state geometry (n_x, n_y, om) is read once per sample and step instead of once per
//     (sample, control): in the 16-sample kernel the two lanes of a sample issue the same shared
//     loads twice, which is what puts its LSU pipe at 75 % (profiles/ncu_car_r2.txt);
//   * the rollout (rsqrt chain) runs once per sample instead of in both lanes;
//   * geometry and parked noise are private to the lane that wrote them: no intra-warp exchange;
//   * column sub-runs are 32 samples long.
// Price: twice the chain state per lane and 28 KB of shared memory per warp -> 8 warps per SM.
#pragma once
#include "car_kernels.cuh"

namespace saa {

constexpr int kCarTile32 = 32;
#ifndef SAA_CAR32_WARPS
#define SAA_CAR32_WARPS 8
#endif
#ifndef SAA_CAR32_CAP
#define SAA_CAR32_CAP 23
#endif

template <int S, int CAP> struct CarPass32 {
  using P16 = CarPass<S, CAP>;                       // same grouping of the chains into passes
  static constexpr int NPASS = P16::NPASS;
  __host__ __device__ static constexpr int bound(int p) { return P16::bound(p); }
  __host__ __device__ static constexpr int pre(int j0, int j) {
    int s = 0;
    for (int jj = j0; jj < j; ++jj) s += kCarTile32 * ((S - 1 - jj) | 1);
    return s;
  }
  static constexpr int PER_C = kCarTile32 * CAP;
  static constexpr int UBROW = S | 1;
  static constexpr int GEO_ROWS = 3 * (S + 1), DWBASE = GEO_ROWS - 2 * S;
  __host__ __device__ static constexpr int stage_bytes(int szTO) {
    const int a = kCarTile32 * UBROW * szTO, b = 2 * PER_C * szTO;
    return ((a > b ? a : b) + 15) / 16 * 16;
  }
};

template <typename T, typename TO, int S, int WARPS> struct CarSmem32 {
  using Ps = CarPass32<S, SAA_CAR32_CAP>;
  static constexpr int STAGE_BYTES = Ps::stage_bytes((int)sizeof(TO));
  CarEgo<T, S> ego;
  T geo[WARPS][Ps::GEO_ROWS][kCarTile32];
  alignas(16) unsigned char stage[WARPS][STAGE_BYTES];
};

template <typename T, typename TO, int S, int J0, int J, int J1>
__device__ __forceinline__ void car32_copy_cols(const CarArgs<T, TO, S> &A, const TO *stage, i64 sbase, int ns,
                                                int lane) {
  if constexpr (J < J1) {
    using C = CarCol<S, J>;
    using Ps = CarPass32<S, SAA_CAR32_CAP>;
    i64 sb = sbase, mout = A.M_out;
    opaque(sb); opaque(mout);
    car_copy_run_t<TO, C::L, C::STRIDE>(A.Ax + (C::CA0 + mout * C::CB0 + sb * C::L), stage + Ps::pre(J0, J),
                                        ns * C::L, lane);
    car_copy_run_t<TO, C::L, C::STRIDE>(A.Ax + (C::CA1 + mout * C::CB1 + sb * C::L),
                                        stage + Ps::PER_C + Ps::pre(J0, J), ns * C::L, lane);
    car32_copy_cols<T, TO, S, J0, J + 1, J1>(A, stage, sbase, ns, lane);
  }
}

// chains j in [J0, J1) of both controls for this lane's sample
template <typename T, typename TO, int S, int J0, int J1>
__device__ __forceinline__ void car32_chain_pass(const CarArgs<T, TO, S> &A, const CarEgo<T, S> &E,
                                                 const T (*geo)[kCarTile32], TO *stage, int lane, T wsdt) {
  using Ps = CarPass32<S, SAA_CAR32_CAP>;
  constexpr int NJ = J1 - J0;
  const T dt = A.dt;
#pragma nv_diag_suppress 549
  CarChain<T> ch[2][NJ > 0 ? NJ : 1];
  static_for<J0 + 1, S + 1>([&](auto kc) {
    constexpr int k = decltype(kc)::value;
    constexpr int JE = (k - 1 < J1) ? (k - 1) : J1;
    constexpr bool ALIVE = JE > J0;
    T nhx = T(0), nhy = T(0);
    [[maybe_unused]] T dot[2][NJ > 0 ? NJ : 1];
    if constexpr (ALIVE) {
      nhx = geo[3 * k][lane]; nhy = geo[3 * k + 1][lane];
      static_for<J0, JE>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          dot[c][j - J0] = fma(nhx, ch[c][j - J0].rx, nhy * ch[c][j - J0].ry);
          stage[c * Ps::PER_C + Ps::pre(J0, j) + lane * CarCol<S, j>::STRIDE + (k - j - 2)] = (TO)(-dot[c][j - J0]);
        }
      });
    }
    if constexpr (k < S) {
      T tt[2][2];
#pragma unroll
      for (int c = 0; c < 2; ++c) { tt[c][0] = E.tt[c][k][0]; tt[c][1] = E.tt[c][k][1]; }
      if constexpr (ALIVE) {
        const T om_n = geo[3 * k + 2][lane];
        const T onx = om_n * nhx, ony = om_n * nhy;
        static_for<J0, JE>([&](auto jc) {
          constexpr int j = decltype(jc)::value;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            CarChain<T> &h = ch[c][j - J0];
            const T d = dot[c][j - J0];
            const T nwx = fma(onx, d, fma(-om_n, h.rx, fma(-wsdt, h.wy, h.wx)));
            const T nwy = fma(ony, d, fma(-om_n, h.ry, fma(-wsdt, h.wy, h.wy)));
            h.rx = fma(-dt, h.wx, h.rx + tt[c][0]);
            h.ry = fma(-dt, h.wy, h.ry + tt[c][1]);
            h.wx = nwx; h.wy = nwy;
          }
        });
      }
      if constexpr (k - 1 >= J0 && k - 1 < J1) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          ch[c][k - 1 - J0].rx = tt[c][0]; ch[c][k - 1 - J0].ry = tt[c][1];
          ch[c][k - 1 - J0].wx = T(0); ch[c][k - 1 - J0].wy = T(0);
        }
      }
    }
  });
}

template <typename T, typename TO, int S, int P>
__device__ __forceinline__ void car32_passes(const CarArgs<T, TO, S> &A, const CarEgo<T, S> &E,
                                             const T (*geo)[kCarTile32], TO *stage, int lane, i64 s0, int ns, T wsdt) {
  using Ps = CarPass32<S, SAA_CAR32_CAP>;
  if constexpr (P < Ps::NPASS) {
    constexpr int J0 = Ps::bound(P), J1 = Ps::bound(P + 1);
    car32_chain_pass<T, TO, S, J0, J1>(A, E, geo, stage, lane, wsdt);
    __syncwarp();
    car32_copy_cols<T, TO, S, J0, J0, J1>(A, stage, s0 + A.first_out, ns, lane);
    __syncwarp();
    car32_passes<T, TO, S, P + 1>(A, E, geo, stage, lane, s0, ns, wsdt);
  }
}

template <typename T, typename TO, int S, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
car_assemble32_kernel(const __grid_constant__ CarArgs<T, TO, S> A) {
  using Ps = CarPass32<S, SAA_CAR32_CAP>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  auto &sm = *reinterpret_cast<CarSmem32<T, TO, S, WARPS> *>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) car_ego_rollout<T, S>(A.us, A.ego0, A.dt, sm.ego);
  __syncthreads();
  if (blockIdx.x == 0 && A.sums != nullptr) car_final_rows<T, TO, S>(A, sm.ego, threadIdx.x, WARPS * 32);
  if (A.Ax == nullptr) return;
  const CarEgo<T, S> &E = sm.ego;
  T (*geo)[kCarTile32] = sm.geo[warp];
  unsigned char *stage_raw = sm.stage[warp];
  bool bad = false;
  const i64 ntiles = (A.M + kCarTile32 - 1) / kCarTile32;
#pragma unroll 1
  for (i64 tile = (i64)blockIdx.x * WARPS + warp; tile < ntiles; tile += (i64)gridDim.x * WARPS) {
    const i64 s0 = tile * kCarTile32;
    const int ns = (int)min((i64)kCarTile32, A.M - s0);
    const bool active = lane < ns;
    const i64 s = s0 + (active ? lane : 0);
    const T w_s = __ldcs(A.om + s), w_r = __ldcs(A.om + A.Mpad + s);
    // ---- pass A: rollout once per sample; geometry and noise rows are private to this lane ----
    {
      T tmp[2 * S];
#pragma unroll
      for (int r = 0; r < 2 * S; ++r) tmp[r] = __ldcs(A.dw + (i64)r * A.Mpad + s);
#pragma unroll
      for (int r = 0; r < 2 * S; ++r) geo[Ps::DWBASE + r][lane] = tmp[r];
    }
    T qx = __ldcs(A.x0 + s), qy = __ldcs(A.x0 + A.Mpad + s);
    T wx = __ldcs(A.x0 + 2 * A.Mpad + s), wy = __ldcs(A.x0 + 3 * A.Mpad + s);
    const T dt = A.dt, wsdt = w_s * dt, dtwr = dt * w_r;
    TO *ubrow = reinterpret_cast<TO *>(stage_raw) + lane * Ps::UBROW;
    CarChain<T> cu[2] = {{T(0), T(0), T(0), T(0)}, {T(0), T(0), T(0), T(0)}};
    T zmax = -INFINITY;
    static_for<0, S + 1>([&](auto kc) {
      constexpr int k = decltype(kc)::value;
      const T dx = E.p[k][0] - qx, dy = E.p[k][1] - qy;
      const T n2 = fma(dx, dx, dy * dy);
      const T inv_n = rsqrt_t(n2);
      bad |= active && (!(n2 > T(0)) || !(n2 < T(INFINITY)));
      const T nhx = dx * inv_n, nhy = dy * inv_n;
      const T om_n = dtwr * inv_n;
      if constexpr (k >= 1) {
        // same association as the 16-sample kernel: control 0's part + control 1's part
        const T gu = -fma(nhx, cu[0].rx, nhy * cu[0].ry) + -fma(nhx, cu[1].rx, nhy * cu[1].ry);
        const T g = A.d_min - n2 * inv_n;
        zmax = fmax(zmax, g);
        ubrow[k - 1] = (TO)(gu - g);
      }
      // row 3k + q is read back by this lane only; its noise rows (steps < k) have been consumed
      geo[3 * k][lane] = nhx; geo[3 * k + 1][lane] = nhy; geo[3 * k + 2][lane] = om_n;
      if constexpr (k < S) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const T ttx = E.tt[c][k][0], tty = E.tt[c][k][1], uc = E.ucum[c][k];
          const T d = fma(nhx, cu[c].rx, nhy * cu[c].ry);
          const T nwx = fma(om_n * nhx, d, fma(-om_n, cu[c].rx, fma(-wsdt, cu[c].wy, cu[c].wx)));
          const T nwy = fma(om_n * nhy, d, fma(-om_n, cu[c].ry, fma(-wsdt, cu[c].wy, cu[c].wy)));
          cu[c].rx = fma(-dt, cu[c].wx, fma(uc, ttx, cu[c].rx));
          cu[c].ry = fma(-dt, cu[c].wy, fma(uc, tty, cu[c].ry));
          cu[c].wx = nwx; cu[c].wy = nwy;
        }
        const T sp = w_s * (A.v_des - wy);
        const T fx = fma(-w_r, nhx, sp), fy = fma(-w_r, nhy, sp);
        const T nqx = fma(dt, wx, qx), nqy = fma(dt, wy, qy);
        wx = wx + dt * fx + A.noise_c * geo[Ps::DWBASE + 2 * k][lane];
        wy = wy + dt * fy + A.noise_c * geo[Ps::DWBASE + 2 * k + 1][lane];
        qx = nqx; qy = nqy;
      }
    });
    if (A.Z != nullptr && active) A.Z[s] = (TO)(zmax - A.ztol);
    __syncwarp();
    if (A.ub != nullptr)
      car_copy_run_t<TO, S, Ps::UBROW>(A.ub + A.ub_off + s0 * S, reinterpret_cast<const TO *>(stage_raw), ns * S, lane);
    __syncwarp();
    car32_passes<T, TO, S, 0>(A, E, geo, reinterpret_cast<TO *>(stage_raw), lane, s0, ns, wsdt);
  }
  if (A.nonfinite != nullptr) {
    const unsigned m = __ballot_sync(0xffffffffu, bad);
    if (lane == 0 && m) atomicAdd(A.nonfinite, (unsigned long long)__popc(m));
  }
}

}  // namespace saa
