// Shared helpers for libsaa_b200 (sm_100a).  See include/saa_b200.h for the ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/saa_b200.h"

namespace saa {

typedef long long i64;

constexpr int kTileSamples = 16;   // samples per warp tile: lanes = 16 samples x 2 "roles"
constexpr int kSMs = 148;          // B200

// ----------------------------------------------------------------------------
// Layout of the assembled matrix for `M` samples (closed form, no scanning).
// Row order / column order: include/saa_b200.h header comment.
// ----------------------------------------------------------------------------
struct Layout {
  int problem = 0, method = 0;
  int S = 0, n_u = 0, n_x = 0, n_fin = 0;
  int blk = 1;          // constraint groups per step (drone: n_obs = 3, car: 1)
  int R = 0;            // sample rows per sample = blk * S
  i64 M = 0;            // samples in the destination matrix
  bool relaxed_pattern = false;   // car scp_iter == 0: rows >= n_x vanish
  int keep_y = 0, keep_s = 0;     // relaxed pattern: surviving "-y_i" rows / sample-0 rows (layout.cuh)
  // rows
  i64 row_cvar = -1, row_y0 = -1, row_s0 = 0, row_slack = -1, row_ctrl0 = 0, n_rows = 0;
  // columns
  int nu = 0;
  i64 n_cols = 0, nnz = 0;
  i64 ycol0 = 0, slackcol = 0, tcol = 0;   // element offsets of the first y column, slack, t
  std::vector<i64> ucol;                   // element offset of each u column (nu + 1)

  // number of final rows hit by u column c, and which
  int fin_rows(int c, int rows[4]) const;
  // sample-run length per sample in u column c (0 if the column carries no sample rows)
  int run_len(int c) const;
  int relaxed_extra(int c) const;   // relaxed pattern: surviving sample-0 entries of u column c
  i64 run_start(int c) const { int r[4]; return ucol[c] + fin_rows(c, r); }
  i64 ycol_len() const;   // entries per y column
  void build(int problem_, int method_, int S_, i64 M_, bool relaxed_pattern_);
  template <typename I> void fill(I *indptr, I *indices) const;
};

// ----------------------------------------------------------------------------
// device helpers
// ----------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T sum16(T v) {   // sum over the 16 lanes that share lane>>4
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}
template <typename T>
__device__ __forceinline__ T sum32(T v) {
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  return sum16(v);
}
template <typename T>
__device__ __forceinline__ T max32(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// compile-time loop: f(std::integral_constant<int, I>) for I in [B, E)
template <int B, int E, typename F>
__device__ __forceinline__ void static_for(F &&f) {
  if constexpr (B < E) {
    f(std::integral_constant<int, B>{});
    static_for<B + 1, E>(f);
  }
}

// Optimisation barrier: the value is unchanged but the compiler must assume it
// was modified.  Used to stop common-subexpression hoisting across the fully
// unrolled sensitivity chains (which would keep ~60 extra doubles live and spill).
__device__ __forceinline__ void opaque(double &v) { asm volatile("" : "+d"(v)); }
__device__ __forceinline__ void opaque(float &v) { asm volatile("" : "+f"(v)); }
__device__ __forceinline__ void opaque(i64 &v) { asm volatile("" : "+l"(v)); }

// streaming (evict-first) store: assembled values are never re-read by this kernel
template <typename T>
__device__ __forceinline__ void st_stream(T *p, T v) {
#ifdef SAA_PLAIN_ST
  *p = v;
#else
  __stcs(p, v);
#endif
}

// Copy a run of n = n_samples * LEN values, staged in shared memory with row
// stride STRIDE (>= LEN, odd => conflict-free 64-bit STS from 16 lanes), to a
// contiguous global run.  Consecutive lanes -> consecutive addresses.
template <typename T, int LEN, int STRIDE>
__device__ __forceinline__ void copy_run(T *__restrict__ dst, const T *__restrict__ src,
                                         int n, int lane) {
#pragma unroll 4
  for (int e = lane; e < n; e += 32) {
    const int i = e / LEN;
    st_stream(dst + e, src[e + i * (STRIDE - LEN)]);
  }
}


// ----------------------------------------------------------------------------
// TMA bulk stores (cp.async.bulk, shared::cta -> global): the copy engine moves a
// staged run to global memory asynchronously; no LDS/STG/index instructions.
// Both addresses must be 16-byte aligned and the size a multiple of 16 bytes.
// ----------------------------------------------------------------------------
__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, unsigned bytes) {
  // L2 evict-first policy (the encoding CUTLASS uses for TMA::CacheHintSm90::EVICT_FIRST): the
  // assembled values are never re-read by this kernel.  Without the hint the copy engine's stores
  // behave like plain st.global and the kernel runs 2.2x slower (5.6 vs 2.6 ms).
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst),
               "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes), "l"(0x12F0000000000000ull)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all of this thread's bulk groups have finished READING shared memory (buffer reusable)
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... all but the most recent group (double-buffered staging)
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// make this thread's generic-proxy shared-memory writes visible to the async proxy (TMA)
__device__ __forceinline__ void fence_async_smem() {
#ifndef SAA_TMA_NOFENCE
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}

// Staging geometry of one CSC column pair (x and y column of a control step):
// 16 sample rows of LEN values per column.
//  * dense (row stride LEN): one bulk store per column run.  Conflict-free 64-bit
//    stores for odd LEN, 2-way for LEN = 2 (mod 4).
//  * row-wise (LEN = 0 mod 4 would be 4..16-way conflicted when dense): padded row
//    stride LEN + VEC, every lane bulk-stores its own row.
// A run that starts at global element g0 is staged at offset (g0 mod VEC) so that
// 16-byte aligned global addresses map to 16-byte aligned shared addresses; the
// (< VEC) head / tail elements go out as plain stores.
template <typename T, int LEN, int ROWS = kTileSamples> struct Stager {
  static constexpr int VEC = 16 / (int)sizeof(T);
#ifdef SAA_TMA_DENSE_ONLY
  static constexpr bool ROWWISE = false;
#else
  static constexpr bool ROWWISE = (LEN % (2 * VEC) == 0);
#endif
  static constexpr int RSTRIDE = LEN + VEC;
  static constexpr int YBASE = ((ROWS * LEN + VEC - 1) / VEC + 1) * VEC;
  static constexpr int SIZE = ROWWISE ? 2 * ROWS * RSTRIDE : 2 * YBASE;

  __device__ __forceinline__ static T *mine(T *stage, int a, int si, i64 g0) {
    const int off = (int)(g0 & (VEC - 1));
    return ROWWISE ? stage + (a * ROWS + si) * RSTRIDE + off
                   : stage + a * YBASE + off + si * LEN;
  }
  // call after fence_async_smem() + __syncwarp(); g0 = global element index (in `base`) of the
  // tile's run of column a; ns = valid sample rows
  __device__ __forceinline__ static void flush(T *base, T *stage, int a, int si, i64 g0, int ns) {
    const int off = (int)(g0 & (VEC - 1));
    const int head = (VEC - off) & (VEC - 1);
    if (ROWWISE) {
      if (si < ns) {
        T *row = stage + (a * ROWS + si) * RSTRIDE + off;
        T *dst = base + g0 + (i64)si * LEN;
        constexpr int dummy = 0; (void)dummy;
        const int bulk = ((LEN - head) / VEC) * VEC, tail = LEN - head - bulk;
        bulk_store(dst + head, row + head, bulk * (unsigned)sizeof(T));
        for (int e = 0; e < head; ++e) st_stream(dst + e, row[e]);
        for (int e = 0; e < tail; ++e) st_stream(dst + head + bulk + e, row[head + bulk + e]);
      }
    } else if (si == 0) {
      const int n = ns * LEN;
      T *run = stage + a * YBASE + off;
      T *dst = base + g0;
      const int h = head < n ? head : n;
      const int bulk = ((n - h) / VEC) * VEC, tail = n - h - bulk;
      if (bulk > 0) bulk_store(dst + h, run + h, bulk * (unsigned)sizeof(T));
      for (int e = 0; e < h; ++e) st_stream(dst + e, run[e]);
      for (int e = 0; e < tail; ++e) st_stream(dst + h + bulk + e, run[h + bulk + e]);
    }
    bulk_commit();
  }

  // ---- line-ownership copy-out (SAA_LINE_OWN) -------------------------------------------------
  // element e of column a's staged run; off = (global start of the run) & (VEC-1)
  __device__ __forceinline__ static const T *elem(const T *stage, int a, int e, int off) {
    if (ROWWISE) {
      constexpr unsigned MAGIC = (unsigned)((0x100000000ull + LEN - 1) / LEN);
      const int row = (int)__umulhi((unsigned)e, MAGIC);
      return stage + (a * ROWS + row) * RSTRIDE + off + (e - row * LEN);
    }
    return stage + a * YBASE + off + e;
  }
  // Copy elements [e0, e1) of column a's run to base + g0 + e with 16-byte stores on 16-byte
  // aligned global addresses (scalars for the < VEC unaligned elements at either end).
  __device__ __forceinline__ static void copy_range(T *base, const T *stage, int a, i64 g0, int e0, int e1,
                                                    int lane) {
    if (e1 <= e0) return;
    const int off = (int)(g0 & (VEC - 1));
    const int h = min(e1 - e0, (int)((VEC - ((g0 + e0) & (VEC - 1))) & (VEC - 1)));
    const int ea = e0 + h;
    const int nvec = (e1 - ea) / VEC, tail = e1 - ea - nvec * VEC;
    T *dst = base + g0;
#pragma unroll 4
    for (int v = lane; v < nvec; v += 32) {
      const int e = ea + v * VEC;
      int4 val;
      if (ROWWISE) {
        constexpr unsigned MAGIC = (unsigned)((0x100000000ull + LEN - 1) / LEN);
        const int row = (int)__umulhi((unsigned)e, MAGIC);
        const int idx = e - row * LEN;
        const T *src = stage + (a * ROWS + row) * RSTRIDE + off + idx;
        if (idx + VEC <= LEN) {
          val = *reinterpret_cast<const int4 *>(src);
        } else {                                   // the vector straddles two (padded) staging rows
          T tmp[VEC];
#pragma unroll
          for (int q = 0; q < VEC; ++q) tmp[q] = *elem(stage, a, e + q, off);
          val = *reinterpret_cast<const int4 *>(tmp);
        }
      } else {
        val = *reinterpret_cast<const int4 *>(stage + a * YBASE + off + e);
      }
      __stcs(reinterpret_cast<int4 *>(dst + e), val);
    }
    if (lane < h) st_stream(dst + e0 + lane, *elem(stage, a, e0 + lane, off));
    if (lane < tail) st_stream(dst + ea + nvec * VEC + lane, *elem(stage, a, ea + nvec * VEC + lane, off));
  }

  // Warp-cooperative copy of column a's staged run with 16-byte shared loads and 16-byte
  // streaming global stores (consecutive lanes -> consecutive 16-byte chunks); the (< VEC)
  // unaligned head / tail elements of a run or row go out as scalar stores.
  __device__ __forceinline__ static void copy_vec(T *base, const T *stage, int a, i64 g0, int ns,
                                                  int lane) {

    const int off = (int)(g0 & (VEC - 1));
    const int head = (VEC - off) & (VEC - 1);
    if (ROWWISE) {
      constexpr int NVMAX = LEN / VEC > 0 ? LEN / VEC : 1;   // vectors per row when off == 0
      const int nv = (LEN - head) / VEC;                     // LEN % VEC == 0: nv = NVMAX or NVMAX - 1
      const int tail = LEN - head - nv * VEC;
      const T *src0 = stage + a * ROWS * RSTRIDE + off + head;
      T *dst0 = base + g0 + head;
      const int total = ns * nv;
      constexpr unsigned MAGIC0 = (unsigned)((0x100000000ull + NVMAX - 1) / NVMAX);
      constexpr unsigned MAGIC1 = NVMAX > 1 ? (unsigned)((0x100000000ull + NVMAX - 2) / (NVMAX - 1)) : 0u;
      const unsigned magic = head ? MAGIC1 : MAGIC0;
#pragma unroll 4
      for (int v = lane; v < total; v += 32) {
        const int row = (int)__umulhi((unsigned)v, magic);
        const int idx = v - row * nv;
        const int4 val = *reinterpret_cast<const int4 *>(src0 + row * RSTRIDE + idx * VEC);
#ifdef SAA_PLAIN_ST
        *reinterpret_cast<int4 *>(dst0 + (i64)row * LEN + idx * VEC) = val;
#else
        __stcs(reinterpret_cast<int4 *>(dst0 + (i64)row * LEN + idx * VEC), val);
#endif
      }
      if (lane < ns) {
        const T *row = stage + (a * ROWS + lane) * RSTRIDE + off;
        T *dst = base + g0 + (i64)lane * LEN;
        for (int e = 0; e < head; ++e) st_stream(dst + e, row[e]);
        for (int e = 0; e < tail; ++e) st_stream(dst + head + nv * VEC + e, row[head + nv * VEC + e]);
      }
    } else {
      const int n = ns * LEN;
      const T *run = stage + a * YBASE + off;
      T *dst = base + g0;
      const int h = head < n ? head : n;
      const int nvec = (n - h) / VEC, tail = n - h - nvec * VEC;
      const int4 *src4 = reinterpret_cast<const int4 *>(run + h);
      int4 *dst4 = reinterpret_cast<int4 *>(dst + h);
#pragma unroll 4
#ifdef SAA_PLAIN_ST
      for (int v = lane; v < nvec; v += 32) dst4[v] = src4[v];
#else
      for (int v = lane; v < nvec; v += 32) __stcs(dst4 + v, src4[v]);
#endif
      if (lane < h) st_stream(dst + lane, run[lane]);
      if (lane < tail) st_stream(dst + h + nvec * VEC + lane, run[h + nvec * VEC + lane]);
    }
  }
};

}  // namespace saa
