// Write-bandwidth probe 2: effect of chunk misalignment and of staging through shared memory.
// nvcc -arch=sm_100a -O3 tools/wbw2.cu -o build/wbw2
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef long long i64;

// chunks of `chunk_d` doubles placed back to back starting at element `off` (so a chunk start is
// generally not 128-byte aligned); each warp writes whole chunks, 16-byte stores with 8-byte head/tail
template <int VIA_SMEM>
__global__ void write_runs(double *dst, i64 nchunks, int chunk_d, int off, i64 mul) {
  __shared__ double sm[6][1024];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const i64 warp = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const i64 nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
  for (i64 c = warp; c < nchunks; c += nwarps) {
    const i64 pc = (c * mul) % nchunks;
    const i64 g0 = off + pc * chunk_d;
    double *p = dst + g0;
    const int head = (int)(g0 & 1);
    const int nvec = (chunk_d - head) / 2, tail = chunk_d - head - 2 * nvec;
    if (VIA_SMEM) {
      for (int e = lane; e < chunk_d; e += 32) sm[w][(e + head) & 1023] = (double)(c + e);
      __syncwarp();
    }
    for (int v = lane; v < nvec; v += 32) {
      double2 val;
      if (VIA_SMEM) val = *reinterpret_cast<double2 *>(&sm[w][(2 * head + 2 * v) & 1023]);
      else val = make_double2((double)c, (double)v);
      __stcs(reinterpret_cast<double2 *>(p + head) + v, val);
    }
    if (lane == 0 && head) __stcs(p, 1.0);
    if (lane == 0 && tail) __stcs(p + chunk_d - 1, 2.0);
    if (VIA_SMEM) __syncwarp();
  }
}

int main(int argc, char **argv) {
  const i64 MUL = argc > 1 ? atoll(argv[1]) : 1000003;
  const i64 total_d = 1200ll * 1000 * 1000;
  double *d;
  cudaMalloc(&d, (total_d + 64) * 8);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  // chunk sizes in doubles: 16 samples x 3L for L = 1, 2, 4, 8, 10, 19 (even and odd lengths)
  const int chunks[] = {48, 96, 192, 384, 480, 912};
  printf("%8s %6s %6s %12s %12s\n", "chunk_d", "bytes", "off", "direct GB/s", "via smem GB/s");
  for (int ci = 0; ci < 6; ++ci)
    for (int off = 0; off < 4; ++off) {
      const int cd = chunks[ci];
      const i64 nchunks = total_d / cd;
      const int threads = 192, blocks = 148 * 2;
      float ms[2];
      for (int mode = 0; mode < 2; ++mode)
        for (int rep = 0; rep < 3; ++rep) {
          cudaEventRecord(a);
          if (mode == 0) write_runs<0><<<blocks, threads>>>(d, nchunks, cd, off, MUL);
          else write_runs<1><<<blocks, threads>>>(d, nchunks, cd, off, MUL);
          cudaEventRecord(b); cudaEventSynchronize(b);
          cudaEventElapsedTime(&ms[mode], a, b);
        }
      printf("%8d %6d %6d %12.0f %12.0f\n", cd, cd * 8, off, nchunks * cd * 8.0 / ms[0] / 1e6,
             nchunks * cd * 8.0 / ms[1] / 1e6);
    }
  return 0;
}
