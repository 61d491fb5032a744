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

template <typename T> struct HopperArgs {
  const T *I, *theta, *tau;   // (M, F) row-major
  i64 M;
  int F, n_c;
  int chunk;                  // samples staged per block iteration
  T mu_nom;
  T px[kHopperMaxContacts];
  T *mu, *dmu;                // (M, n_c)
  const double *lambda;       // (M, n_c) or nullptr
  double *w;                  // per-block partial sums [gridDim.x][2][n_c] of lambda mu', lambda mu''   (if lambda)
};

// one staged feature: 32 bytes in FP64 (two 16-byte shared loads), 16 in FP32
template <typename T> struct __align__(4 * sizeof(T)) HopperFeat { T I, th, ta, nit; };   // nit = -I theta

template <typename T, bool HESS, bool BIG>
__device__ __forceinline__ void hopper_element(const HopperFeat<T> *row, int F, T p, T &m0, T &m1, T &m2) {
  int f = 0;
  if (!BIG && sizeof(T) == 4) {
    // FP32: two features per iteration in one straight-line block (0.93 -> 0.77 ms); in FP64 ptxas
    // serialises the two evaluations again and the plain loop below is as fast
#pragma unroll 1
    for (; f + 2 <= F; f += 2) {
      const HopperFeat<T> f0 = row[f], f1 = row[f + 1];
      const T x[2] = {fma(f0.th, p, f0.ta), fma(f1.th, p, f1.ta)};
      T sn[2], cs[2];
      sincos_core_n<2>(x, sn, cs);
      m0 = fma(f0.I, cs[0], m0);
      m1 = fma(f0.nit, sn[0], m1);
      if (HESS) m2 = fma(f0.nit * f0.th, cs[0], m2);
      m0 = fma(f1.I, cs[1], m0);
      m1 = fma(f1.nit, sn[1], m1);
      if (HESS) m2 = fma(f1.nit * f1.th, cs[1], m2);
    }
  }
#pragma unroll 2
  for (; f < F; ++f) {
    const HopperFeat<T> ft = row[f];
    const T x = fma(ft.th, p, ft.ta);
    T sn, cs;
    if (BIG) sincos_t(x, &sn, &cs);      // library routine (any argument)
    else sincos_core(x, &sn, &cs);       // branch free
    m0 = fma(ft.I, cs, m0);
    m1 = fma(ft.nit, sn, m1);
    if (HESS) m2 = fma(ft.nit * ft.th, cs, m2);
  }
}

template <typename T, bool HESS>
__global__ void __launch_bounds__(kHopperThreads)
hopper_friction_kernel(const __grid_constant__ HopperArgs<T> A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int F = A.F, n_c = A.n_c, CH = A.chunk;
  HopperFeat<T> *sF = reinterpret_cast<HopperFeat<T> *>(smem_raw);
  double *sw = reinterpret_cast<double *>(sF + CH * F);    // HESS: [2][CH * n_c] weighted derivatives of the chunk
  const i64 nchunks = (A.M + CH - 1) / CH;
  if (HESS) {
    // slot e of the chunk (a fixed (sample slot, contact) pair) is always served by the same
    // thread, so the weighted derivatives accumulate in private shared-memory slots without
    // any synchronisation; they are combined once, in a fixed order, after the last chunk
    for (int e = threadIdx.x; e < 2 * CH * n_c; e += kHopperThreads) sw[e] = 0.0;
  }
  T pmax = T(0);
  for (int c = 0; c < n_c; ++c) pmax = fmax(pmax, fabs(A.px[c]));
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
        sF[e] = ft;
        big |= sincos_big(fma(fabs(ft.th), pmax, fabs(ft.ta)));
      }
    }
    big = __syncthreads_or(big);
    for (int e = threadIdx.x; e < ns * n_c; e += kHopperThreads) {
      const int il = e / n_c, c = e - il * n_c;
      T m0 = T(0), m1 = T(0), m2 = T(0);
      if (big) hopper_element<T, HESS, true>(sF + il * F, F, A.px[c], m0, m1, m2);
      else hopper_element<T, HESS, false>(sF + il * F, F, A.px[c], m0, m1, m2);
      const i64 g = i0 * n_c + e;
      A.mu[g] = A.mu_nom + m0;
      A.dmu[g] = m1;
      if (HESS) {
        const double lam = A.lambda[g];
        sw[e] += lam * (double)m1;
        sw[CH * n_c + e] += lam * (double)m2;
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
