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

constexpr int kTileSamples = 16;   // sample rows per warp tile: lanes = 16 samples x 2 "roles"
constexpr int kTileOwn = 15;       // drone: rows 0..14 are the tile's own samples, row 15 the halo (Stager)
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
  // assembled values are never re-read by this kernel; without the hint the same stores ran 2.2x slower
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst),
               "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes), "l"(0x12F0000000000000ull)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all of this thread's bulk groups have finished READING shared memory (buffer reusable)
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// make this thread's generic-proxy shared-memory writes visible to the async proxy (TMA)
__device__ __forceinline__ void fence_async_smem() {
#ifndef SAA_TMA_NOFENCE
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}

// Staging geometry of one CSC column pair (x and y column of a control step):
// kTileSamples sample rows of LEN values per column.
//  * dense (row stride LEN): conflict-free 64-bit stores for odd LEN, 2-way for LEN = 2 (mod 4).
//  * row-wise (LEN = 0 mod 4 would be 4..16-way conflicted when dense): padded row stride LEN + VEC.
// A run that starts at global element g0 is staged at offset (g0 mod VEC) so that 16-byte aligned
// global addresses map to 16-byte aligned shared addresses.
//
// LINE OWNERSHIP.  The runs of neighbouring tiles abut at arbitrary 8-byte offsets, so a tile that
// writes exactly its own elements leaves a partially written 128-byte line at either end of every
// run, completed later by another warp -- measured (tools/wbw3.cu) such partial-line writes cost
// ~3.4 full-line writes each in L2.  Instead the last of the kTileSamples rows is a HALO: a copy
// of the next tile's first sample, computed redundantly by this warp.  A tile then writes every line
// that STARTS inside its own run, whole (the last one extends up to LINE-1 elements into the halo
// row), and nothing before its first line start: every line of the column is written exactly once,
// by one warp, with line-aligned 16-byte vector stores.  This needs LEN >= LINE-1 (one halo row
// covers the overhang); shorter columns, and the first / last tile of a launch, fall back to
// writing exactly their own elements.
template <typename T, int LEN> struct Stager {
  static constexpr int VEC = 16 / (int)sizeof(T);
  static constexpr int LINE = 128 / (int)sizeof(T);
  static constexpr bool OWN_LINES = LEN >= LINE - 1;
#ifdef SAA_TMA_DENSE_ONLY
  static constexpr bool ROWWISE = false;
#else
  static constexpr bool ROWWISE = (LEN % (2 * VEC) == 0);
#endif
  static constexpr int RSTRIDE = LEN + VEC;
  static constexpr int YBASE = ((kTileSamples * LEN + VEC - 1) / VEC + 1) * VEC;
  static constexpr int SIZE = ROWWISE ? 2 * kTileSamples * RSTRIDE : 2 * YBASE;

  __device__ __forceinline__ static T *mine(T *stage, int a, int si, i64 g0) {
    const int off = (int)(g0 & (VEC - 1));
    return ROWWISE ? stage + (a * kTileSamples + si) * RSTRIDE + off
                   : stage + a * YBASE + off + si * LEN;
  }

  // [e0, e1): global elements this tile writes of a run whose own part is [g0, g0 + n_own)
  __device__ __forceinline__ static void span(i64 g0, int n_own, bool first, bool last, i64 &e0, i64 &e1) {
    if (OWN_LINES) {
      e0 = first ? g0 : ((g0 + LINE - 1) & ~(i64)(LINE - 1));
      e1 = last ? g0 + n_own : ((g0 + n_own + LINE - 1) & ~(i64)(LINE - 1));
    } else {
      e0 = g0; e1 = g0 + n_own;
    }
  }

  // staged element i (i = global element - g0) of column a
  __device__ __forceinline__ static const T &elem(const T *stage, int a, int off, int i) {
    if (ROWWISE) {
      const int row = (int)((unsigned)i / (unsigned)LEN), col = i - row * LEN;
      return stage[(a * kTileSamples + row) * RSTRIDE + off + col];
    }
    return stage[a * YBASE + off + i];
  }

  // Warp-cooperative copy of the staged elements [e0 - g0, e1 - g0) of column a to base[e0, e1):
  // 16-byte shared loads -> 16-byte streaming stores, consecutive lanes on consecutive 16-byte
  // chunks (4 whole lines per instruction when e0 is line aligned); the (< VEC) elements in front
  // of / behind the 16-byte aligned part go out as scalar stores.
  __device__ __forceinline__ static void copy_lines(T *base, const T *stage, int a, i64 g0, i64 e0, i64 e1,
                                                    int lane) {
    const int off = (int)(g0 & (VEC - 1));
    const int n = (int)(e1 - e0), i0 = (int)(e0 - g0);
    const int head = (int)((VEC - (e0 & (VEC - 1))) & (VEC - 1));
    const int h = head < n ? head : n;
    const int nvec = (n - h) / VEC, tail = n - h - nvec * VEC;
    T *dst = base + e0;
    if (lane < h) st_stream(dst + lane, elem(stage, a, off, i0 + lane));
    if (lane < tail) st_stream(dst + h + nvec * VEC + lane, elem(stage, a, off, i0 + h + nvec * VEC + lane));
    int4 *dst4 = reinterpret_cast<int4 *>(dst + h);
    if (ROWWISE) {
      const T *src0 = stage + a * kTileSamples * RSTRIDE + off;
#pragma unroll 4
      for (int v = lane; v < nvec; v += 32) {
        const int i = i0 + h + v * VEC;
        const int row = (int)((unsigned)i / (unsigned)LEN), col = i - row * LEN;
        const T *p = src0 + row * RSTRIDE + col;
        int4 val;
        if (col + VEC <= LEN) {
          val = *reinterpret_cast<const int4 *>(p);
        } else {                                   // the vector straddles two sample rows
          T tmp[VEC];
#pragma unroll
          for (int q = 0; q < VEC; ++q) tmp[q] = (col + q < LEN) ? p[q] : p[q + RSTRIDE - LEN];
          val = *reinterpret_cast<const int4 *>(tmp);
        }
#ifdef SAA_PLAIN_ST
        dst4[v] = val;
#else
        __stcs(dst4 + v, val);
#endif
      }
    } else {
      const int4 *src4 = reinterpret_cast<const int4 *>(stage + a * YBASE + off + i0 + h);
#pragma unroll 4
#ifdef SAA_PLAIN_ST
      for (int v = lane; v < nvec; v += 32) dst4[v] = src4[v];
#else
      for (int v = lane; v < nvec; v += 32) __stcs(dst4 + v, src4[v]);
#endif
    }
  }

  // TMA variant for dense staging: ONE lane hands the 16-byte aligned part of the span to the copy
  // engine (cp.async.bulk shared -> global, evict-first) and writes the scalar ends; call after
  // fence_async_smem() + __syncwarp().  The caller commits the bulk group.
  __device__ __forceinline__ static void bulk_lines(T *base, const T *stage, int a, i64 g0, i64 e0, i64 e1) {
    static_assert(!ROWWISE, "bulk stores need the dense layout");
    const int off = (int)(g0 & (VEC - 1));
    const int n = (int)(e1 - e0), i0 = (int)(e0 - g0);
    const int head = (int)((VEC - (e0 & (VEC - 1))) & (VEC - 1));
    const int h = head < n ? head : n;
    const int nvec = (n - h) / VEC, tail = n - h - nvec * VEC;
    const T *run = stage + a * YBASE + off + i0;
    T *dst = base + e0;
    if (nvec > 0) bulk_store(dst + h, run + h, (unsigned)(nvec * VEC * sizeof(T)));
    for (int e = 0; e < h; ++e) st_stream(dst + e, run[e]);
    for (int e = 0; e < tail; ++e) st_stream(dst + h + nvec * VEC + e, run[h + nvec * VEC + e]);
  }
};

}  // namespace saa
