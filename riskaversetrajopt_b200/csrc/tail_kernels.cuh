// Tail selection for the reduced CVaR subproblem, sm_100a.
//
// The Rockafellar-Uryasev reformulation (reference drone/drone_risk.py:327-368) carries one
// auxiliary variable y_i and S*n_obs rows per sample, but at the solution only the samples in the
// upper alpha-tail of Z_i = max_{o,k} g_i[o,k] have y_i > 0.  Feeding the host QP solver only the
// K = ceil((1 + margin) alpha M) samples with the largest Z_i at the current iterate keeps the
// subproblem tractable at M = 1e6 (and cuts the device -> host traffic by M/K).  This file holds
// the selection: an exact, deterministic K-largest radix select over the Z_i (8 bits per pass on an
// order-preserving integer image of the floating-point keys), followed by an order-preserving
// compaction -- the selected sample indices come out ASCENDING, ties at the threshold resolved
// towards the smaller index, so the reduced matrix is the full matrix with rows / columns deleted.
#pragma once
#include "saa_common.cuh"

namespace saa {

struct SelectState {
  unsigned long long prefix;   // key bits decided so far (high digits of the K-th largest key)
  long long k_rem;             // how many keys equal to the final threshold are taken
  long long n_gt;              // keys strictly above the threshold
};

// order-preserving map to unsigned integers (larger value <-> larger key); -0.0 and +0.0 get the
// same key (z + 0 turns -0.0 into +0.0); NaN sorts above +inf
__device__ __forceinline__ unsigned long long select_key(double z) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(z + 0.0);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ unsigned long long select_key(float z) {
  const unsigned b = __float_as_uint(z + 0.0f);
  return (unsigned long long)((b >> 31) ? ~b : (b | 0x80000000u));
}
template <typename T> struct SelectBits { static constexpr int value = 8 * (int)sizeof(T); };

// histogram of the digit at `shift` over the keys whose higher digits equal the prefix
template <typename T>
__global__ void __launch_bounds__(256)
select_hist_kernel(const T *__restrict__ Z, i64 n, int shift, const SelectState *__restrict__ st,
                   unsigned *__restrict__ hist) {
  __shared__ unsigned sh[256];
  sh[threadIdx.x] = 0;
  __syncthreads();
  const unsigned long long prefix = st->prefix;
  const int hi = shift + 8;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
    const unsigned long long k = select_key(Z[i]);
    const bool match = hi >= 64 ? true : ((k >> hi) == (prefix >> hi));
    if (match) atomicAdd(&sh[(k >> shift) & 255], 1u);
  }
  __syncthreads();
  if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}

// pick the digit that holds the K-th largest key, update prefix / k_rem, clear the histogram
__global__ void select_pick_kernel(unsigned *__restrict__ hist, int shift, SelectState *__restrict__ st) {
  if (threadIdx.x == 0) {
    long long k = st->k_rem, above = 0;
    int d = 255;
    for (; d > 0; --d) {
      if (above + (long long)hist[d] >= k) break;
      above += hist[d];
    }
    st->prefix |= (unsigned long long)d << shift;
    st->k_rem = k - above;
    st->n_gt += above;
  }
  __syncthreads();
  for (int b = threadIdx.x; b < 256; b += blockDim.x) hist[b] = 0;
}

constexpr int kSelThreads = 256;
constexpr int kSelItems = 16;     // consecutive keys per thread in the compaction kernels

// per-block counts of keys above / equal to the threshold; a block owns a contiguous chunk
template <typename T>
__global__ void __launch_bounds__(kSelThreads)
select_count_kernel(const T *__restrict__ Z, i64 n, const SelectState *__restrict__ st,
                    i64 *__restrict__ counts) {
  __shared__ i64 red[2][kSelThreads / 32];
  const unsigned long long thr = st->prefix;
  const i64 base = (i64)blockIdx.x * kSelThreads * kSelItems + (i64)threadIdx.x * kSelItems;
  i64 gt = 0, eq = 0;
  for (int q = 0; q < kSelItems; ++q) {
    const i64 i = base + q;
    if (i < n) { const unsigned long long k = select_key(Z[i]); gt += k > thr; eq += k == thr; }
  }
  gt = sum32(gt); eq = sum32(eq);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = gt; red[1][threadIdx.x >> 5] = eq; }
  __syncthreads();
  if (threadIdx.x == 0) {
    i64 a = 0, b = 0;
    for (int w = 0; w < kSelThreads / 32; ++w) { a += red[0][w]; b += red[1][w]; }
    counts[2 * (i64)blockIdx.x] = a; counts[2 * (i64)blockIdx.x + 1] = b;
  }
}

// exclusive prefix sums of the block counts, in place (one block; nblk is a few thousand at most)
__global__ void __launch_bounds__(1024)
select_scan_kernel(i64 *__restrict__ counts, i64 nblk) {
  __shared__ i64 carry[2];
  __shared__ i64 wsum[2][32];
  if (threadIdx.x == 0) { carry[0] = 0; carry[1] = 0; }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (i64 b0 = 0; b0 < nblk; b0 += blockDim.x) {
    const i64 b = b0 + threadIdx.x;
    i64 v[2] = {b < nblk ? counts[2 * b] : 0, b < nblk ? counts[2 * b + 1] : 0};
    i64 inc[2];
    for (int c = 0; c < 2; ++c) {
      i64 x = v[c];
      for (int o = 1; o < 32; o <<= 1) { const i64 y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
      inc[c] = x;
      if (lane == 31) wsum[c][warp] = x;
    }
    __syncthreads();
    if (warp == 0) {
      for (int c = 0; c < 2; ++c) {
        i64 x = lane < (int)(blockDim.x >> 5) ? wsum[c][lane] : 0;
        for (int o = 1; o < 32; o <<= 1) { const i64 y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        wsum[c][lane] = x;                        // inclusive over warps
      }
    }
    __syncthreads();
    for (int c = 0; c < 2; ++c) {
      const i64 before = carry[c] + (warp ? wsum[c][warp - 1] : 0) + inc[c] - v[c];
      if (b < nblk) counts[2 * b + c] = before;
    }
    __syncthreads();
    if (threadIdx.x == 0) { carry[0] += wsum[0][(blockDim.x >> 5) - 1]; carry[1] += wsum[1][(blockDim.x >> 5) - 1]; }
    __syncthreads();
  }
}

// idx_out[#selected before i] = i for every selected key (above the threshold, or equal to it and
// among the first k_rem equal keys in index order)
template <typename T>
__global__ void __launch_bounds__(kSelThreads)
select_write_kernel(const T *__restrict__ Z, i64 n, const SelectState *__restrict__ st,
                    const i64 *__restrict__ offsets, i64 *__restrict__ idx_out) {
  __shared__ i64 wsum[2][kSelThreads / 32];
  const unsigned long long thr = st->prefix;
  const i64 k_rem = st->k_rem;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const i64 base = (i64)blockIdx.x * kSelThreads * kSelItems + (i64)threadIdx.x * kSelItems;
  unsigned gtm = 0, eqm = 0;
  for (int q = 0; q < kSelItems; ++q) {
    const i64 i = base + q;
    if (i < n) {
      const unsigned long long k = select_key(Z[i]);
      gtm |= (unsigned)(k > thr) << q; eqm |= (unsigned)(k == thr) << q;
    }
  }
  i64 cnt[2] = {__popc(gtm), __popc(eqm)}, exc[2];
  for (int c = 0; c < 2; ++c) {
    i64 x = cnt[c];
    for (int o = 1; o < 32; o <<= 1) { const i64 y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    exc[c] = x - cnt[c];
    if (lane == 31) wsum[c][warp] = x;
  }
  __syncthreads();
  i64 gt_before = offsets[2 * (i64)blockIdx.x] + exc[0], eq_before = offsets[2 * (i64)blockIdx.x + 1] + exc[1];
  for (int w = 0; w < warp; ++w) { gt_before += wsum[0][w]; eq_before += wsum[1][w]; }
  for (int q = 0; q < kSelItems; ++q) {
    const bool g = (gtm >> q) & 1, e = (eqm >> q) & 1;
    if (g || (e && eq_before < k_rem)) idx_out[gt_before + min(eq_before, k_rem)] = base + q;
    gt_before += g; eq_before += e;
  }
}

// dst[row * dst_stride + r] = src[row * src_stride + idx[min(r, K - 1)]] for r < n_write:
// packed sample rows of the selected samples (the padding repeats the last one, as the pack kernels do)
template <typename T>
__global__ void __launch_bounds__(256)
gather_rows_kernel(const T *__restrict__ src, i64 src_stride, T *__restrict__ dst, i64 dst_stride, i64 n_write,
                   int nrows, const i64 *__restrict__ idx, i64 K) {
  const i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_write) return;
  const i64 s = idx[r < K ? r : K - 1];
  for (int row = 0; row < nrows; ++row) dst[(i64)row * dst_stride + r] = src[(i64)row * src_stride + s];
}

}  // namespace saa
