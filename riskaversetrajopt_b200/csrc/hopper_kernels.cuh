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
// One thread per (sample, contact) pair, flattened so that all 32 lanes of a warp
// work (a warp spans <= 3 samples: the feature loads are 1-3 broadcasts).  This
// kernel is FP64-pipe bound (n_features sincos per output triple), not HBM bound.
#pragma once
#include "saa_common.cuh"
#include "car_kernels.cuh"   // sincos_t

namespace saa {

constexpr int kHopperMaxContacts = 32;

template <typename T> struct HopperArgs {
  const T *I, *theta, *tau;   // (M, F) row-major
  i64 M;
  int F, n_c;
  T mu_nom;
  T px[kHopperMaxContacts];
  T *mu, *dmu;                // (M, n_c)
  const double *lambda;       // (M, n_c) or nullptr
  double *w;                  // scratch (2, M, n_c): lambda mu', lambda mu''   (if lambda)
};

template <typename T>
__global__ void __launch_bounds__(256)
hopper_friction_kernel(const __grid_constant__ HopperArgs<T> A) {
  const i64 total = A.M * A.n_c;
  for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
    const i64 i = e / A.n_c;
    const int c = (int)(e - i * A.n_c);
    const T p = A.px[c];
    const T *I = A.I + i * A.F, *th = A.theta + i * A.F, *ta = A.tau + i * A.F;
    T m0 = T(0), m1 = T(0), m2 = T(0);
#pragma unroll 2
    for (int f = 0; f < A.F; ++f) {
      const T t = th[f], in = I[f];
      T sn, cs;
      sincos_t(fma(t, p, ta[f]), &sn, &cs);
      m0 = fma(in, cs, m0);
      const T it = in * t;
      m1 = fma(-it, sn, m1);
      m2 = fma(-it * t, cs, m2);
    }
    A.mu[e] = A.mu_nom + m0;
    A.dmu[e] = m1;
    if (A.lambda != nullptr) {
      const double lam = A.lambda[e];
      A.w[e] = lam * (double)m1;
      A.w[total + e] = lam * (double)m2;
    }
  }
}

// out[2c + which] = sum_i w[which][i][c], fixed summation order (deterministic)
__global__ void hopper_reduce_kernel(const double *__restrict__ w, i64 M, int n_c, double *__restrict__ out) {
  __shared__ double red[256];
  const int c = blockIdx.x % n_c, which = blockIdx.x / n_c;
  const double *src = w + (i64)which * M * n_c + c;
  double acc = 0.0;
  for (i64 i = threadIdx.x; i < M; i += blockDim.x) acc += src[i * n_c];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[2 * c + which] = red[0];
}

template <typename T>
__global__ void hopper_pack_kernel(const double *__restrict__ I, const double *__restrict__ th,
                                   const double *__restrict__ ta, i64 n, T *oI, T *oth, T *ota) {
  const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) { oI[e] = (T)I[e]; oth[e] = (T)th[e]; ota[e] = (T)ta[e]; }
}

}  // namespace saa
