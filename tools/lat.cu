// Dependent-issue latencies that shape the latency-bound kernels (one warp, clock64):
// nvcc -arch=sm_100a -O3 tools/lat.cu -o build/lat
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double *out, long long *cyc, double a, double b) {
  __shared__ double sm[64];
  double x = a + threadIdx.x;
  long long t0, t1;
  const int N = 256;
  // DFMA chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = fma(x, b, a);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = (t1 - t0);
  // DMUL + DFMA alternating (drone chain step)
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { double y = x * a; x = fma(b, x, y); }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[1] = (t1 - t0);
  // rsqrt chain
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) x = rsqrt(x + 2.0);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[2] = (t1 - t0);
  // SHFL.64 + DADD chain
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) x += __shfl_xor_sync(0xffffffffu, x, 1);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[3] = (t1 - t0);
  // STS -> LDS round trip
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) { sm[threadIdx.x] = x; __syncwarp(); x = sm[threadIdx.x ^ 1] + 1.0; __syncwarp(); }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[4] = (t1 - t0);
  // independent DFMAs (throughput, one warp): 8 chains
  double y0 = x, y1 = x + 1, y2 = x + 2, y3 = x + 3, y4 = x + 4, y5 = x + 5, y6 = x + 6, y7 = x + 7;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) {
    y0 = fma(y0, b, a); y1 = fma(y1, b, a); y2 = fma(y2, b, a); y3 = fma(y3, b, a);
    y4 = fma(y4, b, a); y5 = fma(y5, b, a); y6 = fma(y6, b, a); y7 = fma(y7, b, a);
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[5] = (t1 - t0);
  out[threadIdx.x] = x + y0 + y1 + y2 + y3 + y4 + y5 + y6 + y7;
}
int main() {
  double *o; long long *c, h[6];
  cudaMalloc(&o, 32 * 8); cudaMalloc(&c, 6 * 8);
  for (int rep = 0; rep < 2; ++rep) lat<<<1, 32>>>(o, c, 1.0000001, 0.9999999);
  cudaMemcpy(h, c, 48, cudaMemcpyDeviceToHost);
  const char *names[] = {"DFMA dependent", "DMUL+DFMA dependent pair", "rsqrt(double) dependent", "SHFL.64 + DADD dependent",
                         "STS.64 -> syncwarp -> LDS.64 -> DADD -> syncwarp", "DFMA, 8 independent chains (per instruction)"};
  const double div[] = {256, 256, 256, 256, 256, 256 * 8};
  for (int i = 0; i < 6; ++i) printf("%-52s %7.1f cycles\n", names[i], h[i] / div[i]);
  return 0;
}
