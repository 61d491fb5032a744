// Measured FP64 FMA peak of the device (the denominator of the FP64-pipe rooflines of the car and
// hopper kernels):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/fp64_peak.cu -o build/fp64_peak
// Each thread runs ILP independent dependent-DFMA chains; blocks x threads saturate every SM.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void __launch_bounds__(256) dfma_kernel(double *out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  if (s == 12345.678) out[0] = s;     // never true: keeps the chains alive
}

template <int ILP>
void run(int sms, int warps_per_sm) {
  double *out; cudaMalloc(&out, 8);
  const int iters = 4096;
  const int blocks = sms * (warps_per_sm * 32 / 256);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    dfma_kernel<ILP><<<blocks, 256>>>(out, iters, 0.9999999, 1e-7);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  const double fma = (double)blocks * 256 * iters * 8 * ILP;
  printf("ILP %d, %2d warps/SM: %.3f ms  %.2f T FMA/s = %.2f TFLOP/s  (%.1f FMA/clk/SM at 1.965 GHz)\n", ILP, warps_per_sm, best,
         fma / best / 1e9, 2 * fma / best / 1e9, fma / best / 1e-3 / sms / 1.965e9);
  cudaFree(out);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
  for (int w : {8, 16, 32, 64}) run<4>(p.multiProcessorCount, w);
  for (int w : {8, 16, 32}) run<8>(p.multiProcessorCount, w);
  run<1>(p.multiProcessorCount, 64);
  return 0;
}
