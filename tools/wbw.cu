// Write-bandwidth probe: how fast can B200 absorb 9.6 GB of streaming stores when they arrive as
// scattered contiguous chunks (what the CSC column sub-runs look like) instead of one sequential
// stream?   nvcc -arch=sm_100a -O3 tools/wbw.cu -o build/wbw && build/wbw
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef long long i64;

// each warp writes chunk `c` (chunk_bytes contiguous), chunks visited in a strided order so that
// concurrently active warps hit addresses `spread` chunks apart
template <int MODE>   // 0: st.cs, 1: plain
__global__ void write_chunks(double2 *dst, i64 nchunks, int chunk_vec, i64 stride_mul) {
  const int lane = threadIdx.x & 31;
  const i64 warp = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const i64 nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
  for (i64 c = warp; c < nchunks; c += nwarps) {
    const i64 pc = (c * stride_mul) % nchunks;          // permuted chunk index (stride_mul coprime)
    double2 *p = dst + pc * chunk_vec;
    for (int v = lane; v < chunk_vec; v += 32) {
      double2 val = make_double2((double)c, (double)v);
      if (MODE == 0) __stcs(p + v, val); else p[v] = val;
    }
  }
}

int main() {
  const i64 bytes = 9600ll * 1000 * 1000;
  double2 *d;
  cudaMalloc(&d, bytes);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  const int chunk_sizes[] = {384, 1024, 2048, 4096, 8192, 32768, 262144};
  const int warps_per_sm[] = {12, 32, 64};
  printf("%10s %8s %10s %10s %10s\n", "chunk_B", "warps/SM", "order", "st.cs GB/s", "plain GB/s");
  for (int wi = 0; wi < 3; ++wi)
    for (int ci = 0; ci < 7; ++ci)
      for (int ord = 0; ord < 2; ++ord) {
        const int cb = chunk_sizes[ci];
        const i64 nchunks = bytes / cb;
        const i64 mul = ord == 0 ? 1 : 1000003;        // sequential chunk order vs scattered
        const int threads = 128, blocks = 148 * warps_per_sm[wi] * 32 / threads;
        float ms[2];
        for (int mode = 0; mode < 2; ++mode) {
          for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(a);
            if (mode == 0) write_chunks<0><<<blocks, threads>>>(d, nchunks, cb / 16, mul);
            else write_chunks<1><<<blocks, threads>>>(d, nchunks, cb / 16, mul);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            cudaEventElapsedTime(&ms[mode], a, b);
          }
        }
        printf("%10d %8d %10s %10.0f %10.0f\n", cb, warps_per_sm[wi], ord ? "scattered" : "sequential",
               nchunks * (double)cb / ms[0] / 1e6, nchunks * (double)cb / ms[1] / 1e6);
      }
  return 0;
}
