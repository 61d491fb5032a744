// Accuracy of sincos_core (csrc/car_kernels.cuh) against the CUDA library over the argument range
// of the hopper friction field and beyond: nvcc -arch=sm_100a -O3 -std=c++17 tools/sincos_check.cu -o build/sincos_check
#include <cstdio>
#include <cmath>
#include "../riskaversetrajopt_b200/csrc/car_kernels.cuh"

__device__ double ulps(double a, double b, int mant) {
  if (a == b) return 0.0;
  const double u = fabs(b) > 0 ? exp2((double)(ilogb(b) - mant)) : 4.9e-324;
  return fabs(a - b) / u;
}

template <typename T>
__global__ void check(long long n, double range, double *out) {
  double mu = 0.0, ma = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    // low-discrepancy sweep of [-range, range] plus points next to multiples of pi/2
    double xd = range * (2.0 * ((i * 0.6180339887498949) - floor(i * 0.6180339887498949)) - 1.0);
    if ((i & 7) == 0) xd = rint(xd / 1.5707963267948966) * 1.5707963267948966 * (1.0 + 1e-15 * (double)(i % 13));
    const T x = (T)xd;
    double s0, c0;
    sincos((double)x, &s0, &c0);                 // reference: FP64 library on the (rounded) argument
    T s1, c1;
    saa::sincos_fast(x, &s1, &c1);
    const int mant = sizeof(T) == 8 ? 52 : 23;
    mu = fmax(mu, fmax(ulps((double)s1, s0, mant), ulps((double)c1, c0, mant)));
    ma = fmax(ma, fmax(fabs((double)s1 - s0), fabs((double)c1 - c0)));
  }
  atomicMax((unsigned long long *)&out[0], (unsigned long long)__double_as_longlong(mu));
  atomicMax((unsigned long long *)&out[1], (unsigned long long)__double_as_longlong(ma));
}

int main() {
  double *d, h[2];
  cudaMalloc(&d, 16);
  const double ranges[] = {1.0, 10.0, 100.0, 9.9e3, 9.9e4, 1e6, 1e12};
  for (int prec = 0; prec < 2; ++prec)
    for (double r : ranges) {
      cudaMemset(d, 0, 16);
      if (prec == 0) check<double><<<148 * 8, 256>>>(1ll << 26, r, d);
      else check<float><<<148 * 8, 256>>>(1ll << 26, r, d);
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("%s |x| <= %-8g  max error vs FP64 sincos(): %.3f ulp   max abs error: %.3e\n", prec ? "fp32" : "fp64", r, h[0], h[1]);
    }
  return cudaDeviceSynchronize() != cudaSuccess;
}
