// Accuracy of sincos_fast (csrc/car_kernels.cuh) against the CUDA library sincos over the argument
// range of the hopper friction field and beyond: nvcc -arch=sm_100a -O3 tools/sincos_check.cu -o build/sincos_check
#include <cstdio>
#include <cmath>
#include "../riskaversetrajopt_b200/csrc/car_kernels.cuh"

__device__ double ulps(double a, double b) {
  if (a == b) return 0.0;
  const double u = fabs(b) > 0 ? exp2((double)(ilogb(b) - 52)) : 4.9e-324;
  return fabs(a - b) / u;
}

__global__ void check(long long n, double range, double *out) {
  double mu = 0.0, ma = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    // low-discrepancy sweep of [-range, range] plus points next to multiples of pi/2
    double x = range * (2.0 * ((i * 0.6180339887498949) - floor(i * 0.6180339887498949)) - 1.0);
    if ((i & 7) == 0) x = rint(x / 1.5707963267948966) * 1.5707963267948966 * (1.0 + 1e-15 * (double)(i % 13));
    double s0, c0, s1, c1;
    sincos(x, &s0, &c0);
    saa::sincos_fast(x, &s1, &c1);
    mu = fmax(mu, fmax(ulps(s1, s0), ulps(c1, c0)));
    ma = fmax(ma, fmax(fabs(s1 - s0), fabs(c1 - c0)));
  }
  atomicMax((unsigned long long *)&out[0], (unsigned long long)__double_as_longlong(mu));
  atomicMax((unsigned long long *)&out[1], (unsigned long long)__double_as_longlong(ma));
}

int main() {
  double *d, h[2];
  cudaMalloc(&d, 16);
  const double ranges[] = {1.0, 10.0, 100.0, 1e4, 9.9e4, 1e6, 1e12};
  for (double r : ranges) {
    cudaMemset(d, 0, 16);
    check<<<148 * 8, 256>>>(1ll << 26, r, d);
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("|x| <= %-8g  max ulp error vs sincos(): %.3f   max abs error: %.3e\n", r, h[0], h[1]);
  }
  return cudaDeviceSynchronize() != cudaSuccess;
}
