// Write-bandwidth probe 3: the drone matrix's real column geometry (38 u columns with sample runs of
// 3(S-1-J) doubles at their true CSC offsets, M = 1e6), written without any arithmetic, to find what
// the store pattern alone can reach:
//   mode 0  a warp owns 16 samples and writes its 38 runs one after the other (the kernel's pattern)
//   mode 1  a block's warps write the block's (16 x WARPS)-sample run of each column together,
//           interleaved 512-byte pieces (block-cooperative copy-out)
//   mode 2  as 0 with 32-sample warp tiles
//   mode 3  line ownership: a warp writes every 128-byte line that STARTS inside its run (the last
//           one extends into the next tile's first sample), so that every line is written once, whole,
//           by one warp, with line-aligned 16-byte vector stores
// nvcc -arch=sm_100a -O3 tools/wbw3.cu -o build/wbw3
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef long long i64;
constexpr int S = 20;

__device__ __forceinline__ i64 col_base(int J, int a, i64 M) {
  const i64 CA = 9 * J + 2 + 3 * a;
  const i64 CB = 6ll * J * (S - 1) - 3ll * J * (J - 1) + (a ? 3 * (S - 1 - J) : 0);
  return CA + M * CB;
}

// write n doubles starting at element g0 (8-byte aligned): 16-byte vectors + scalar head/tail,
// vector index v handled by (v % nthr == tid)
__device__ __forceinline__ void write_run(double *dst, i64 g0, int n, int tid, int nthr, double val) {
  const int head = (int)(g0 & 1);
  const int nvec = (n - head) / 2, tail = n - head - 2 * nvec;
  double2 *d2 = reinterpret_cast<double2 *>(dst + g0 + head);
  for (int v = tid; v < nvec; v += nthr) __stcs(d2 + v, make_double2(val, val));
  if (tid == 0 && head) __stcs(dst + g0, val);
  if (tid == 0 && tail) __stcs(dst + g0 + n - 1, val);
}

// whole lines [ceil128(first byte), ceil128(last byte)) of the run; the matrix-level first / last
// partial lines are written by the first / last tile
__device__ __forceinline__ void write_lines(double *dst, i64 g0, int n, bool first, bool last, int lane, double val) {
  i64 e0 = first ? g0 : ((g0 + 15) & ~15ll), e1 = last ? g0 + n : ((g0 + n + 15) & ~15ll);
  if (first && (e0 & 1)) { if (lane == 0) __stcs(dst + e0, val); ++e0; }
  if (last && (e1 & 1)) { if (lane == 0) __stcs(dst + e1 - 1, val); --e1; }
  double2 *d2 = reinterpret_cast<double2 *>(dst + e0);
  const int nvec = (int)((e1 - e0) >> 1);
  for (int v = lane; v < nvec; v += 32) __stcs(d2 + v, make_double2(val, val));
}

template <int MODE>
__global__ void probe(double *dst, i64 M, int tile) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (MODE == 1) {
    const i64 bt = (i64)tile * nw, nt = (M + bt - 1) / bt;
    for (i64 t = blockIdx.x; t < nt; t += gridDim.x) {
      const i64 s0 = t * bt;
      const int ns = (int)min(bt, M - s0);
      for (int J = 0; J < S - 1; ++J)
        for (int a = 0; a < 2; ++a) {
          const int LEN = 3 * (S - 1 - J);
          write_run(dst, col_base(J, a, M) + s0 * LEN, ns * LEN, threadIdx.x, blockDim.x, (double)J);
        }
    }
  } else {
    const i64 nt = (M + tile - 1) / tile;
    for (i64 t = (i64)blockIdx.x * nw + warp; t < nt; t += (i64)gridDim.x * nw) {
      const i64 s0 = t * tile;
      const int ns = (int)min((i64)tile, M - s0);
      for (int J = 0; J < S - 1; ++J)
        for (int a = 0; a < 2; ++a) {
          const int LEN = 3 * (S - 1 - J);
          if (MODE == 3) write_lines(dst, col_base(J, a, M) + s0 * LEN, ns * LEN, t == 0, t == nt - 1, lane, (double)J);
          else write_run(dst, col_base(J, a, M) + s0 * LEN, ns * LEN, lane, 32, (double)J);
        }
    }
  }
}

// car geometry: 2 x 19 columns with sample runs of S-1-J doubles (csrc/car_kernels.cuh, CarCol)
__global__ void probe_car(double *dst, i64 M, int tile) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const i64 nt = (M + tile - 1) / tile;
  for (i64 t = (i64)blockIdx.x * nw + warp; t < nt; t += (i64)gridDim.x * nw) {
    const i64 s0 = t * tile;
    const int ns = (int)min((i64)tile, M - s0);
    for (int J = 0; J < S - 1; ++J)
      for (int c = 0; c < 2; ++c) {
        const int L = S - 1 - J;
        const i64 base = 8 * J + 3 + 4 * c + M * (2ll * J * (S - 1) - (i64)J * (J - 1) + (c ? L : 0));
        write_run(dst, base + s0 * L, ns * L, lane, 32, (double)J);
      }
  }
}

int main() {
  const i64 M = 1000000;
  const i64 total = 1140 * M + 1024;
  double *d;
  cudaMalloc(&d, total * 8);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  printf("%5s %9s %9s %6s %10s\n", "mode", "warps/blk", "blocks/SM", "tile", "GB/s");
  struct Cfg { int mode, warps, bps, tile; };
  const Cfg cfgs[] = {{0, 6, 2, 16}, {0, 8, 8, 16}, {2, 6, 2, 32}, {1, 6, 2, 16}, {1, 32, 2, 16},
                      {3, 6, 2, 15}, {3, 6, 2, 16}, {3, 6, 2, 31}, {3, 6, 4, 15}, {3, 8, 8, 15}, {3, 6, 1, 15}, {3, 4, 2, 15}};
  for (const Cfg &c : cfgs) {
    float best = 1e9f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(a);
      if (c.mode == 1) probe<1><<<148 * c.bps, c.warps * 32>>>(d, M, c.tile);
      else if (c.mode == 3) probe<3><<<148 * c.bps, c.warps * 32>>>(d, M, c.tile);
      else probe<0><<<148 * c.bps, c.warps * 32>>>(d, M, c.tile);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      if (rep && ms < best) best = ms;
    }
    printf("%5d %9d %9d %6d %10.0f\n", c.mode, c.warps, c.bps, c.tile, 1140.0 * M * 8 / best / 1e6);
  }
  printf("car geometry (380 entries per sample):\n");
  for (int tile : {16, 32, 64})
    for (int bps : {1, 2}) {
      float best = 1e9f;
      for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(a);
        probe_car<<<148 * bps, 12 * 32>>>(d, M, tile);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (rep && ms < best) best = ms;
      }
      printf("  tile %3d  warps/SM %3d  %8.0f GB/s  (%.3f ms)\n", tile, 12 * bps, 380.0 * M * 8 / best / 1e6, best);
    }
  return cudaDeviceSynchronize() != cudaSuccess;
}
