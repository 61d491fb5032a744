// Hopper slip-risk kernels (K3), sm_100a.
//
// Reference hopper/hopper.py: the state/control trajectory is a decision variable
// (direct transcription), so there is no per-sample rollout; the only sample
// dependent quantity is the friction field (:68-81)
//     mu_i(p) = mu_nom + sum_f I_if cos(theta_if p + tau_if)
// evaluated at the end-effector x position of every contact instant (:300-367).
// Its first / second derivatives give the slip-risk rows of jacrev(g) and
// hessian(lambda . g) (:569-580):
//     mu'  = -sum_f I theta   sin(theta p + tau)
//     mu'' = -sum_f I theta^2 cos(theta p + tau)
//
// A block stages the feature rows of a chunk of samples in shared memory with coalesced
// loads (the (M, F) rows of consecutive samples are contiguous in global memory) as records
// (I, theta, tau, -I theta); then one thread per (sample, contact) pair of the chunk loops
// over the features, reading each record with two 16-byte shared-memory broadcasts (a warp
// spans <= 3 samples).  The kernel is bound by FP64 issue (n_features sincos per output
// triple), so the sincos is the 22-instruction branch-free sincos_core (car_kernels.cuh; two
// consecutive features interleave their dependency chains), mu'' is only formed when the
// Hessian sums are requested, and a chunk whose arguments could leave the core's range
// (|theta| max|p| + |tau| >= 1e5) takes the library sincos instead.
#pragma once
#include "saa_common.cuh"
#include "car_kernels.cuh"   // sincos_fast

namespace saa {

constexpr int kHopperMaxContacts = 32;
constexpr int kHopperThreads = 128;
#ifndef SAA_HOPPER_UNROLL
#define SAA_HOPPER_UNROLL 4
#endif
constexpr int kHopperUnroll = SAA_HOPPER_UNROLL;   // features interleaved in the inner loop

// T = arithmetic / feature type, TO = storage type of the outputs (double | float)
template <typename T, typename TO> struct HopperArgs {
  const T *I, *theta, *tau;   // (M, F) row-major
  i64 M;
  int F, n_c;
  int chunk;                  // samples staged per block iteration
  int saa;                    // method == saa: rows carry - t - y_i
  T mu_nom;
  T px[kHopperMaxContacts];   // end-effector x position per contact instant (hopper.py:166-171)
  T fx[kHopperMaxContacts], fz[kHopperMaxContacts];          // contact forces us[t, 2], us[t, 3]
  T gp[3][kHopperMaxContacts];                               // d p / d(x0, x2, x3) = (1, x3 cos x2, sin x2)
  T t_risk, slack;
  const double *y;            // (M) risk variables y_i (device) or nullptr (treated as 0)
  TO *mu, *dmu;               // (M, n_c) friction and its derivative (optional)
  TO *g;                      // (M, n_c) slip-risk rows f_x - mu_i(p) f_z - t - y_i - slack (optional)
  TO *jac;                    // [4][M n_c]: d row / d(x0, x2, x3, f_z) (optional)
  const double *lambda;       // (M, n_c) or nullptr
  double *w;                  // per-block partial sums [gridDim.x][2][n_c] of lambda mu', lambda mu''   (if lambda)
  TO *Z;                      // (M) max_c (f_x - mu_i f_z) (optional; hopper.py:910-925)
  T sat_tol;
  double *cvar;               // per-block [gridDim.x][3]: sum max(Z - t, 0), #{Z <= sat_tol}, max Z (optional)
};

// One staged feature (48 bytes, three 16-byte shared loads; the fast path reads the first two):
// th2 = theta 2/pi and ta2 = tau 2/pi give the argument in quarter turns, nit = -I theta.
template <typename T> struct __align__(16) HopperFeat { T I, nit, th2, ta2, th, ta; };

// sin / cos of (pi/2) y for y = th2 p + ta2 in QUARTER TURNS: n = rint(y) by the 1.5 2^52 shift,
// f = y - n exactly (|f| <= 1/2), then sin(pi/2 f) = f P(f^2), cos(pi/2 f) = 1 + f^2 Q(f^2) with the
// fdlibm minimax kernels rescaled by powers of pi/2 (|error| <= 1.2e-16 on |f| <= 1/2, checked with
// mpmath), quadrant fix-up on the integer pipe.  19 FP64 instructions per feature instead of the 22
// of sincos_core + the argument FMA: the three-step Cody-Waite reduction is replaced by an exact
// subtraction.  The argument carries the rounding of th2 p + ta2, i.e. a few ulp(|x|) like
// numpy's cos(theta p + tau) itself; callers switch to the library routine for |x| >= 1e5.
struct HopperPoly { double s[7], c[7]; };
__constant__ HopperPoly kHopperPoly = {
    {1.5707963267948966192, -0.6459640975062449269, 0.079692626246063343923, -0.0046817541326259334138,
     0.00016044115266738095834, -3.5986495702406919016e-6, 5.6347041138847509518e-8},
    {-1.2337005501361698274, 0.25366950790104761936, -0.020863480763330759822, 0.00091926027439055337847,
     -0.000025202037916917742371, 4.7106415058035018794e-7, -6.3247466788660698911e-9}};

__device__ __forceinline__ void sincos_turn(double y, const HopperPoly &K, double *s, double *c) {
  const double SHIFT = 6755399441055744.0;                    // 1.5 * 2^52
  const double t = y + SHIFT;                                 // integer part of y in the low bits
  const int n = __double2loint(t);
  const double f = y - (t - SHIFT);
  const double z = f * f;
  double ps = fma(z, K.s[6], K.s[5]);
  ps = fma(z, ps, K.s[4]);
  ps = fma(z, ps, K.s[3]);
  ps = fma(z, ps, K.s[2]);
  ps = fma(z, ps, K.s[1]);
  ps = fma(z, ps, K.s[0]);
  const double sn = f * ps;
  double pc = fma(z, K.c[6], K.c[5]);
  pc = fma(z, pc, K.c[4]);
  pc = fma(z, pc, K.c[3]);
  pc = fma(z, pc, K.c[2]);
  pc = fma(z, pc, K.c[1]);
  pc = fma(z, pc, K.c[0]);
  const double cs = fma(z, pc, 1.0);
  // n mod 4: 0 (s, c)  1 (c, -s)  2 (-s, -c)  3 (-c, s)
  const bool swap = n & 1;
  double so = swap ? cs : sn, co = swap ? sn : cs;
  so = __hiloint2double(__double2hiint(so) ^ ((n & 2) << 30), __double2loint(so));
  co = __hiloint2double(__double2hiint(co) ^ (((n + 1) & 2) << 30), __double2loint(co));
  *s = so; *c = co;
}

template <typename T, bool HESS, bool BIG>
__device__ __forceinline__ void hopper_element(const HopperFeat<T> *row, int F, T p, const HopperPoly &K,
                                               T &m0, T &m1, T &m2) {
#pragma unroll kHopperUnroll
  for (int f = 0; f < F; ++f) {
    T sn, cs, I, nit, th;
    if (BIG) {
      const HopperFeat<T> ft = row[f];
      sincos_t(fma(ft.th, p, ft.ta), &sn, &cs);      // library routine (any argument)
      I = ft.I; nit = ft.nit; th = ft.th;
    } else {
      const double4 q = *reinterpret_cast<const double4 *>(&row[f]);     // I, nit, th2, ta2
      sincos_turn(fma(q.z, p, q.w), K, &sn, &cs);    // branch free
      I = q.x; nit = q.y; th = HESS ? row[f].th : T(0);
    }
    m0 = fma(I, cs, m0);
    m1 = fma(nit, sn, m1);
    if (HESS) m2 = fma(nit * th, cs, m2);
  }
}

// Shared memory: [features CH x F] [HESS: 2 x CH x n_c doubles] [CVAR: CH x n_c doubles]
#ifndef SAA_HOPPER_MINBLOCKS
#define SAA_HOPPER_MINBLOCKS 6    // <= 85 registers: 4 features in flight per thread, coefficients in (uniform) registers
#endif
template <typename T, typename TO, bool HESS, bool CVAR>
__global__ void __launch_bounds__(kHopperThreads, SAA_HOPPER_MINBLOCKS)
hopper_friction_kernel(const __grid_constant__ HopperArgs<T, TO> A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int F = A.F, n_c = A.n_c, CH = A.chunk;
  HopperFeat<T> *sF = reinterpret_cast<HopperFeat<T> *>(smem_raw);
  double *sw = reinterpret_cast<double *>(sF + CH * F);    // HESS: [2][CH * n_c] weighted derivatives of the chunk
  double *sc = sw + (HESS ? 2 * CH * n_c : 0);             // CVAR: [CH * n_c] constraint values of the chunk
  __shared__ double red[kHopperThreads / 32][3];
  const i64 nchunks = (A.M + CH - 1) / CH;
  if (HESS) {
    // slot e of the chunk (a fixed (sample slot, contact) pair) is always served by the same
    // thread, so the weighted derivatives accumulate in private shared-memory slots without
    // any synchronisation; they are combined once, in a fixed order, after the last chunk
    for (int e = threadIdx.x; e < 2 * CH * n_c; e += kHopperThreads) sw[e] = 0.0;
  }
  double acc_excess = 0.0, acc_sat = 0.0, acc_max = -INFINITY;
  // polynomial coefficients pinned in registers: left in the constant bank they are re-loaded
  // (LDC) inside the feature loop, in front of the FMAs that consume them
  HopperPoly K = kHopperPoly;
#pragma unroll
  for (int q = 0; q < 7; ++q) { opaque(K.s[q]); opaque(K.c[q]); }
  T pmax = T(0);
  for (int c = 0; c < n_c; ++c) pmax = fmax(pmax, fabs(A.px[c]));
  const i64 N = A.M * n_c;
  for (i64 chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    const i64 i0 = chunk * CH;
    const int ns = (int)min((i64)CH, A.M - i0);
    __syncthreads();                                   // the previous chunk has been consumed
    int big = 0;                                       // an argument may leave the fast sincos's range
    {
      const T *gI = A.I + i0 * F, *gTh = A.theta + i0 * F, *gTa = A.tau + i0 * F;
      for (int e = threadIdx.x; e < ns * F; e += kHopperThreads) {
        HopperFeat<T> ft;
        ft.I = __ldcs(gI + e); ft.th = __ldcs(gTh + e); ft.ta = __ldcs(gTa + e); ft.nit = -ft.I * ft.th;
        ft.th2 = ft.th * T(0.63661977236758134308); ft.ta2 = ft.ta * T(0.63661977236758134308);
        sF[e] = ft;
        big |= sincos_big(fma(fabs(ft.th), pmax, fabs(ft.ta)));
      }
    }
    big = __syncthreads_or(big);
    for (int e = threadIdx.x; e < ns * n_c; e += kHopperThreads) {
      const int il = e / n_c, c = e - il * n_c;
      T m0 = T(0), m1 = T(0), m2 = T(0);
      if (big) hopper_element<T, HESS, true>(sF + il * F, F, A.px[c], K, m0, m1, m2);
      else hopper_element<T, HESS, false>(sF + il * F, F, A.px[c], K, m0, m1, m2);
      const i64 g = i0 * n_c + e;
      const T mu = A.mu_nom + m0;
      const T cons = fma(-mu, A.fz[c], A.fx[c]);       // f_x - mu_i(p) f_z  (hopper.py:318-325)
      if (A.mu != nullptr) { A.mu[g] = (TO)mu; A.dmu[g] = (TO)m1; }
      if (A.g != nullptr) {
        T v = cons - A.slack;
        if (A.saa) v = cons - A.t_risk - (A.y ? (T)A.y[i0 + il] : T(0)) - A.slack;   // :366, same order of operations
        A.g[g] = (TO)v;
      }
      if (A.jac != nullptr) {
        const T d = -A.fz[c] * m1;                     // d row / d p
        st_stream(A.jac + g, (TO)(d * A.gp[0][c]));
        st_stream(A.jac + N + g, (TO)(d * A.gp[1][c]));
        st_stream(A.jac + 2 * N + g, (TO)(d * A.gp[2][c]));
        st_stream(A.jac + 3 * N + g, (TO)(-mu));
      }
      if (HESS) {
        const double lam = A.lambda[g];
        sw[e] += lam * (double)m1;
        sw[CH * n_c + e] += lam * (double)m2;
      }
      if (CVAR) sc[e] = (double)cons;
    }
    if (CVAR) {
      __syncthreads();
      for (int il = threadIdx.x; il < ns; il += kHopperThreads) {
        double z = -INFINITY;
        for (int c = 0; c < n_c; ++c) z = fmax(z, sc[il * n_c + c]);
        if (A.Z != nullptr) A.Z[i0 + il] = (TO)z;
        acc_excess += fmax(z - (double)A.t_risk, 0.0);
        acc_sat += (z <= (double)A.sat_tol) ? 1.0 : 0.0;
        acc_max = fmax(acc_max, z);
      }
    }
  }
  if (HESS) {
    __syncthreads();
    if ((int)threadIdx.x < 2 * n_c) {
      const int which = threadIdx.x / n_c, c = threadIdx.x - which * n_c;
      const double *src = sw + which * CH * n_c + c;
      double hacc = 0.0;
      for (int il = 0; il < CH; ++il) hacc += src[il * n_c];   // sample-slot order (deterministic)
      A.w[(i64)blockIdx.x * 2 * n_c + threadIdx.x] = hacc;
    }
  }
  if (CVAR && A.cvar != nullptr) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    acc_excess = sum32(acc_excess); acc_sat = sum32(acc_sat); acc_max = max32(acc_max);
    if (lane == 0) { red[warp][0] = acc_excess; red[warp][1] = acc_sat; red[warp][2] = acc_max; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double e = 0.0, cc = 0.0, m = -INFINITY;
      for (int w = 0; w < kHopperThreads / 32; ++w) { e += red[w][0]; cc += red[w][1]; m = fmax(m, red[w][2]); }
      A.cvar[(i64)blockIdx.x * 3 + 0] = e;
      A.cvar[(i64)blockIdx.x * 3 + 1] = cc;
      A.cvar[(i64)blockIdx.x * 3 + 2] = m;
    }
  }
}

// The sample-independent rest of the slip-risk block (saa, hopper.py:350-357): row 0 =
// M alpha t + sum_i y_i (deterministic two-level sum), rows 1 + i = -y_i, last row = 0.
template <typename TO>
__global__ void __launch_bounds__(256)
hopper_head_rows_kernel(const double *__restrict__ y, i64 M, TO *__restrict__ rows_y, double *__restrict__ partials) {
  __shared__ double red[8];
  double acc = 0.0;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (i64)gridDim.x * blockDim.x) {
    const double v = y[i];
    rows_y[i] = (TO)(-v);
    acc += v;
  }
  acc = sum32(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    partials[blockIdx.x] = t;
  }
}
template <typename TO>
__global__ void hopper_head_finish_kernel(const double *__restrict__ partials, int nblocks, double M_alpha_t,
                                          TO *__restrict__ row0, TO *__restrict__ row_last) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double t = 0.0;
    for (int b = 0; b < nblocks; ++b) t += partials[b];
    *row0 = (TO)(M_alpha_t + t);
    *row_last = (TO)0;
  }
}

// Hessian of lambda . g restricted to the slip-risk rows: per contact c a symmetric block on
// (x0, x2, x3, f_z)_t, lower triangle in the order (0,0) (1,0) (1,1) (2,0) (2,1) (2,2) (3,0) (3,1) (3,2) (3,3):
//   -f_z (L2 grad p grad p^T + L1 Hess p) on (x0,x2,x3)^2,  -L1 grad p on (f_z, x),  0 on (f_z, f_z)
// with L1 = sum_i lambda_ic mu_i', L2 = sum_i lambda_ic mu_i''; Hess p: (x2,x2) = -x3 sin x2, (x2,x3) = cos x2.
__global__ void hopper_hess_blocks_kernel(const double *__restrict__ sums /* [n_c][2] */, int n_c,
                                          const double *__restrict__ geo /* [6][n_c]: fz, gp0, gp1, gp2, h11, h12 */,
                                          double *__restrict__ out /* [10][n_c] */) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_c) return;
  const double L1 = sums[2 * c], L2 = sums[2 * c + 1];
  const double fz = geo[c], g0 = geo[n_c + c], g1 = geo[2 * n_c + c], g2 = geo[3 * n_c + c];
  const double h11 = geo[4 * n_c + c], h12 = geo[5 * n_c + c];
  out[0 * n_c + c] = -fz * (L2 * g0 * g0);
  out[1 * n_c + c] = -fz * (L2 * g1 * g0);
  out[2 * n_c + c] = -fz * (L2 * g1 * g1 + L1 * h11);
  out[3 * n_c + c] = -fz * (L2 * g2 * g0);
  out[4 * n_c + c] = -fz * (L2 * g2 * g1 + L1 * h12);
  out[5 * n_c + c] = -fz * (L2 * g2 * g2);
  out[6 * n_c + c] = -L1 * g0;
  out[7 * n_c + c] = -L1 * g1;
  out[8 * n_c + c] = -L1 * g2;
  out[9 * n_c + c] = 0.0;
}

// out[2c + which] = sum over blocks of partials[b][which][c]; one warp per output, lanes stride
// over the blocks, fixed-order butterfly (deterministic)
__global__ void hopper_reduce_kernel(const double *__restrict__ partials, int nblocks, int n_c,
                                     double *__restrict__ out) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= 2 * n_c) return;
  double acc = 0.0;
  for (int b = lane; b < nblocks; b += 32) acc += partials[(i64)b * 2 * n_c + r];
  acc = sum32(acc);
  const int which = r / n_c, c = r - which * n_c;
  if (lane == 0) out[2 * c + which] = acc;
}

template <typename T>
__global__ void hopper_pack_kernel(const double *__restrict__ I, const double *__restrict__ th,
                                   const double *__restrict__ ta, i64 n, T *oI, T *oth, T *ota) {
  const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) { oI[e] = (T)I[e]; oth[e] = (T)th[e]; ota[e] = (T)ta[e]; }
}

}  // namespace saa
