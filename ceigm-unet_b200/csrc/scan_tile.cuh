// Tile staging shared by the forward and backward scan kernels.
//
// A CTA walks L in tiles of LT scan positions. Each tile of every operand is staged ONCE into shared
// memory in SCAN ORDER as fp32 (pitch LTP = LT + 4 floats so that rows start 4 banks apart and every
// 4-float group stays 16-byte aligned), whatever the tensor's dtype, layout or traversal direction.
// That is where cross-scan is folded in: for NATURAL layout the copy reads the (H, W) plane through
// the direction's index map (common.cuh ScanOrder) instead of a materialised permuted tensor.
#pragma once
#include "common.cuh"

namespace ss2d {

constexpr int kThreads = 128;

// dst[r][0..LT) <- row r of `src` at scan positions [l0, l0+len), zero-filled beyond len / rows_valid.
// row_off(r) gives the element offset of row r's plane/sequence start.
// The copy is shared by `nthr` threads; `thr` is this thread's index among them.
template <int LT, int LTP, typename RowOff>
__device__ __forceinline__ void stage_rows(float* __restrict__ dst, const void* __restrict__ src, int dt,
                                           RowOff row_off, int rows_total, int rows_valid, int l0, int len,
                                           const ScanOrder so, int thr = threadIdx.x, int nthr = kThreads) {
  constexpr int G4 = LT / 4;
  const bool vecL = (so.L & 3) == 0;
  for (int i = thr; i < rows_total * G4; i += nthr) {
    const int r = i / G4, c = (i - r * G4) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < rows_valid && c < len) {
      const int64_t ro = row_off(r);
      bool done = false;
      if (so.contiguous() && c + 3 < len) {
        if (!so.reversed()) {
          const int64_t idx = ro + l0 + c;
          if (vec4_ok(src, idx, dt)) { v = load4(src, idx, dt); done = true; }
        } else if (vecL) {
          const int64_t idx = ro + (so.L - 4 - (l0 + c));
          if (vec4_ok(src, idx, dt)) {
            const float4 t = load4(src, idx, dt);
            v = make_float4(t.w, t.z, t.y, t.x);
            done = true;
          }
        }
      }
      if (!done) {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (c + e < len) f4_at(v, e) = load1(src, ro + so.natural(l0 + c + e), dt);
      }
    }
    *reinterpret_cast<float4*>(dst + (LTP == LT ? swz(r, c) : r * LTP + c)) = v;
  }
}

// Writes 4 consecutive scan positions [l, l+4) (clipped to l_end) of one row to a global tensor.
__device__ __forceinline__ void store_scan4(void* __restrict__ dst, int dt, int64_t ro, int l, int l_end, float4 v,
                                            const ScanOrder so, bool accum) {
  if (l >= l_end) return;
  if (!accum && so.contiguous() && l + 3 < l_end) {
    if (!so.reversed()) {
      if (vec4_ok(dst, ro + l, dt)) { store4(dst, ro + l, dt, v); return; }
    } else if ((so.L & 3) == 0) {
      const int64_t idx = ro + (so.L - 4 - l);
      if (vec4_ok(dst, idx, dt)) { store4(dst, idx, dt, make_float4(v.w, v.z, v.y, v.x)); return; }
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    if (l + e < l_end) {
      const int64_t idx = ro + so.natural(l + e);
      float x = f4_at(v, e);
      if (accum) x += load1(dst, idx, dt);
      store1(dst, idx, dt, x);
    }
  }
}

// Reduce-scatter over the R lanes that share a row (lane = row_lane * R + q): on entry every lane holds
// R groups of 4 partial sums, on exit v[0..3] of lane q holds the full sums of group q.
template <int R>
__device__ __forceinline__ void reduce_scatter_groups(float* v, int q) {
#pragma unroll
  for (int s = R / 2; s >= 1; s >>= 1) {
    const bool up = (q & s) != 0;
#pragma unroll
    for (int i = 0; i < s * 4; ++i) {
      const float lo = v[i], hi = v[i + s * 4];
      const float send = up ? lo : hi;
      const float keep = up ? hi : lo;
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
}

}  // namespace ss2d
