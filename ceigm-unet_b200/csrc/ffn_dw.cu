// The depthwise stack of the GroupMamba FFNs on channels-last tensors:
//   PVT2FFN    (/root/reference/gm-unet/model/gm/groupmamba.py:54-83):  fc1 -> DWConv 3x3 (+bias) -> GELU -> fc2
//   custom_ffn (/root/reference/gm-unet/model/gm/custom_mlp.py:338-368): fc1 -> DWConv 3x3 -> GELU ->
//              InceptionDWConv2d_MultiScale (:313-336: x + cat(x_id, dw3x3(x_3), dw5x5(x_5), dw7x7(x_7))) -> fc2
// The reference transposes the (B, L, C) token tensor to NCHW for every convolution and back (DWConv.forward,
// groupmamba.py:446-455; custom_mlp.py:323-336), runs cuDNN / ATen depthwise kernels, a separate GELU, a split, three
// convolutions, a cat and an add. Here the token tensor IS the (B, H, W, C) channels-last image: one stencil kernel with
// a per-channel-segment kernel size and a fused epilogue covers every forward and data-gradient pass
//   epi 0: y = acc         1: y = gelu(acc)         2: y = x + acc (residual)         3: y = aux * gelu'(acc)
// (acc = bias + sum over taps, or acc = x for an identity segment, k = 1; flip = 1 correlates with the flipped kernel =
// transposed convolution), and one reduction kernel gives the weight / bias gradients of a segment. No transposed copy,
// no split / cat, GELU and the residual never touch HBM on their own. HBM-bound: 2 (epi 0-2) or 3 (epi 3) tensor passes
// of s bytes per element. Math in fp32 whatever the I/O type (fp32 / fp16 / bf16); erf-based GELU as nn.GELU() (gelu_phi).
//
// Machine mappings (the first version — runtime dtype, 2 channels x 4 pixels of a row per thread, one image row per CTA row —
// ran at 0.3 TB/s, slower than the reference composition). Common to all kernels: dtype, vector width and "has GELU" are
// template parameters; a thread owns VEC consecutive channels (one 16-byte access in fp32, 8 bytes in 16-bit types; VEC
// drops to 2 / 1 when C or a segment boundary is not a multiple of 4); a CTA = 16 channel lanes (whole 128 / 256-byte row
// pieces per pixel) x 16 pixel lanes; channel block is the fastest grid index.
//   * dwnhwc_walk3_kernel: the single-segment 3 x 3 case (the DWConv of both FFNs in all its passes) — a column walker with
//     the neighbourhood in a register window and the next three rows in flight (see the kernel's comment);
//   * dwnhwc_stencil_kernel: the general tiled kernel (multi-scale stage, narrow vectors): DW_PY vertically adjacent
//     pixels per thread, the K - 1 halo rows shared by the DW_PY outputs; a CTA never straddles a segment — channel blocks
//     are enumerated per segment, the kernel size is CTA-uniform and dispatched to a fully unrolled body; 3 x 3 weights go
//     straight from global memory into registers, 5 x 5 / 7 x 7 weights are staged as [tap][channel] in shared memory (one
//     broadcast LDS.128 per tap and 4 VEC FMAs);
//   * dwnhwc_wgrad_kernel: the weight / bias gradient of one segment — a row walker (see its comment).
#include <type_traits>

#include "common.cuh"
#include "host_util.h"

namespace ss2d {

int scan_path_policy();     // api.cu (test hook)

constexpr int kDwnMaxSeg = 4;
struct DwnSegs {
  int nseg;
  int cbeg[kDwnMaxSeg + 1];        // channels [cbeg[s], cbeg[s + 1]) form segment s
  int k[kDwnMaxSeg];               // 0 (acc = 0), 1 (identity: acc = x), 3, 5 or 7
  const float* w[kDwnMaxSeg];      // (channels of the segment, k, k) fp32, as nn.Conv2d(groups = channels).weight
  const float* b[kDwnMaxSeg];      // (channels of the segment) fp32 or null
};

constexpr int DW_CL = 16;          // channel lanes of a CTA
constexpr int DW_PL = 16;          // pixel lanes of a CTA (image columns in the stencil, flattened pixels in the reduction)
constexpr int DW_PY = 4;           // vertically adjacent outputs per thread (stencil)
constexpr int DW_THREADS = DW_CL * DW_PL;

struct DwnPlan {
  int blk0[kDwnMaxSeg + 1];        // first channel block of each segment (blk0[nseg] = number of blocks)
  int nblk, tiles_w, tiles_h;
};

// ---- VEC consecutive channels as one access ----
template <typename T> __device__ __forceinline__ float2 unpack2(uint32_t raw);
template <> __device__ __forceinline__ float2 unpack2<__half>(uint32_t raw) { return __half22float2(*reinterpret_cast<const __half2*>(&raw)); }
template <> __device__ __forceinline__ float2 unpack2<__nv_bfloat16>(uint32_t raw) {
  return make_float2(__uint_as_float(raw << 16), __uint_as_float(raw & 0xffff0000u));      // bf16 -> fp32 is a shift
}
template <typename T> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

template <typename T, int VEC>
__device__ __forceinline__ void ldvec(const T* __restrict__ p, float (&v)[VEC]) {
  if constexpr (sizeof(T) == 4) {
    if constexpr (VEC == 4) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(p));
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else if constexpr (VEC == 2) {
      const float2 t = __ldg(reinterpret_cast<const float2*>(p));
      v[0] = t.x; v[1] = t.y;
    } else {
      v[0] = __ldg(reinterpret_cast<const float*>(p));
    }
  } else {
    if constexpr (VEC == 4) {
      const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
      const float2 a = unpack2<T>(t.x), b = unpack2<T>(t.y);
      v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    } else if constexpr (VEC == 2) {
      const float2 a = unpack2<T>(__ldg(reinterpret_cast<const uint32_t*>(p)));
      v[0] = a.x; v[1] = a.y;
    } else {
      const uint32_t raw = __ldg(reinterpret_cast<const uint16_t*>(p));
      v[0] = unpack2<T>(raw).x;
    }
  }
}
template <typename T, int VEC>
__device__ __forceinline__ void stvec(T* __restrict__ p, const float (&v)[VEC]) {
  if constexpr (sizeof(T) == 4) {
    if constexpr (VEC == 4) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    else if constexpr (VEC == 2) *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
    else *reinterpret_cast<float*>(p) = v[0];
  } else {
    if constexpr (VEC == 4) *reinterpret_cast<uint2*>(p) = make_uint2(pack2<T>(v[0], v[1]), pack2<T>(v[2], v[3]));
    else if constexpr (VEC == 2) *reinterpret_cast<uint32_t*>(p) = pack2<T>(v[0], v[1]);
    else *reinterpret_cast<uint16_t*>(p) = static_cast<uint16_t>(pack2<T>(v[0], 0.f) & 0xffffu);
  }
}

// VEC channels as loaded (unconverted): what a software-pipelined loop keeps in flight
template <typename T, int VEC>
struct RawVec {
  static constexpr int NW = (sizeof(T) * VEC + 3) / 4;
  uint32_t r[NW];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < NW; ++i) r[i] = 0u;
  }
  __device__ __forceinline__ void load(const T* __restrict__ p) {
    if constexpr (sizeof(T) * VEC == 16) {
      const uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
      r[0] = t.x; r[1] = t.y; r[2] = t.z; r[3] = t.w;
    } else if constexpr (sizeof(T) * VEC == 8) {
      const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
      r[0] = t.x; r[1] = t.y;
    } else if constexpr (sizeof(T) * VEC == 4) {
      r[0] = __ldg(reinterpret_cast<const uint32_t*>(p));
    } else {
      r[0] = __ldg(reinterpret_cast<const uint16_t*>(p));
    }
  }
  __device__ __forceinline__ void get(float (&v)[VEC]) const {
    if constexpr (sizeof(T) == 4) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) v[i] = __uint_as_float(r[i]);
    } else if constexpr (VEC == 1) {
      v[0] = unpack2<T>(r[0]).x;
    } else {
#pragma unroll
      for (int i = 0; i < VEC / 2; ++i) {
        const float2 t = unpack2<T>(r[i]);
        v[2 * i] = t.x; v[2 * i + 1] = t.y;
      }
    }
  }
};

// nn.GELU() = x Phi(x), Phi from erfc(z) = (a1 t + ... + a5 t^5) exp(-z^2), t = 1 / (1 + p z), z = |x| / sqrt(2)
// (Abramowitz & Stegun 7.1.26, |error| <= 1.5e-7 on erf — fp32 rounding level; no cancellation in the negative tail, where
// 1 + erff(z) loses every digit). One MUFU.RCP + one MUFU.EX2 + 10 FMA-pipe instructions; erff() costs about three times that,
// and exp(-z^2) = exp(-x^2 / 2) is also the density term of the derivative.
template <bool GRAD>
__device__ __forceinline__ float gelu_phi(float x, float aux) {
  const float az = fabsf(x) * 0.70710678f;
  const float t = __fdividef(1.f, fmaf(0.3275911f, az, 1.f));
  const float e = ex2f(-az * az * kLog2e);
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(t, poly, 1.421413741f);
  poly = fmaf(t, poly, -0.284496736f);
  poly = fmaf(t, poly, 0.254829592f);
  const float hc = 0.5f * t * poly * e;                       // Phi(-|x|)
  const float phi = x < 0.f ? hc : 1.f - hc;
  if (GRAD) return aux * fmaf(x * 0.3989422804f, e, phi);      // aux * d gelu / dx
  return x * phi;
}

// The same on a channel pair with packed fp32x2 instructions (FFMA2 / FMUL2: two lanes per issue slot); 0.5 is folded into
// the polynomial, 1 / sqrt(2) into the constants.
template <bool GRAD>
__device__ __forceinline__ float2 gelu_phi2(float2 x, float2 aux) {
  const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
  const float2 den = __ffma2_rn(ax, make_float2(0.3275911f * 0.70710678f, 0.3275911f * 0.70710678f), make_float2(1.f, 1.f));
  const float2 t = make_float2(__fdividef(1.f, den.x), __fdividef(1.f, den.y));
  const float2 arg = __fmul2_rn(__fmul2_rn(x, x), make_float2(-0.5f * kLog2e, -0.5f * kLog2e));
  const float2 e = make_float2(ex2f(arg.x), ex2f(arg.y));
  float2 poly = __ffma2_rn(t, make_float2(0.5f * 1.061405429f, 0.5f * 1.061405429f), make_float2(0.5f * -1.453152027f, 0.5f * -1.453152027f));
  poly = __ffma2_rn(t, poly, make_float2(0.5f * 1.421413741f, 0.5f * 1.421413741f));
  poly = __ffma2_rn(t, poly, make_float2(0.5f * -0.284496736f, 0.5f * -0.284496736f));
  poly = __ffma2_rn(t, poly, make_float2(0.5f * 0.254829592f, 0.5f * 0.254829592f));
  const float2 hc = __fmul2_rn(__fmul2_rn(t, poly), e);         // Phi(-|x|)
  const float2 phi = make_float2(x.x < 0.f ? hc.x : 1.f - hc.x, x.y < 0.f ? hc.y : 1.f - hc.y);
  if (GRAD) return __fmul2_rn(aux, __ffma2_rn(__fmul2_rn(x, make_float2(0.3989422804f, 0.3989422804f)), e, phi));
  return __fmul2_rn(x, phi);
}

// acc[py][v] += sum over taps for the DW_PY pixels (h0 + py, w), channels c .. c + VEC - 1.
// xb: this thread's pointer to (row h0 - P, column w - P) of its image (dereferenced only where the image exists); every load
// address is xb + (i * WC + j * C) with compile-time i, j and kernel-uniform WC = W * C, C — one 32-bit uniform offset per
// load, no per-load 64-bit arithmetic. wts: K == 3: the segment's weights in global memory (read straight into registers:
// no shared-memory hop, no barrier between a CTA's start and its loads); K > 3: this thread's column of the CTA's
// [tap][CB] tile in shared memory (flip already applied).
template <typename T, int VEC, int K>
__device__ __forceinline__ void dwn_taps(const T* __restrict__ xb, const float* __restrict__ wts, int flip, int h0, int w, int H,
                                         int W, int WC, int C, float (&acc)[DW_PY][VEC]) {
  constexpr int P = K / 2, CB = DW_CL * VEC;
  constexpr int NR = K == 3 ? 9 : 1;
  float wr[NR][VEC];
  if constexpr (K == 3) {
    float wf[VEC * 9];                                           // [channel][tap], as stored
    if (VEC == 4 && (reinterpret_cast<uintptr_t>(wts) & 15) == 0) {
#pragma unroll
      for (int q = 0; q < (VEC * 9) / 4; ++q) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(wts) + q);
        wf[4 * q] = t.x; wf[4 * q + 1] = t.y; wf[4 * q + 2] = t.z; wf[4 * q + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int q = 0; q < VEC * 9; ++q) wf[q] = __ldg(wts + q);
    }
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int v = 0; v < VEC; ++v) wr[t][v] = flip ? wf[v * 9 + 8 - t] : wf[v * 9 + t];
  }
  bool colok[K];
#pragma unroll
  for (int j = 0; j < K; ++j) colok[j] = w + j - P >= 0 && w + j - P < W;
#pragma unroll
  for (int i = 0; i < DW_PY + K - 1; ++i) {
    const int r = h0 + i - P;
    if (r < 0 || r >= H) continue;                               // CTA-uniform
    float xv[K][VEC];
#pragma unroll
    for (int j = 0; j < K; ++j) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) xv[j][v] = 0.f;
      if (colok[j]) ldvec<T, VEC>(xb + (i * WC + j * C), xv[j]);
    }
#pragma unroll
    for (int py = 0; py < DW_PY; ++py) {
      const int ki = i - py;
      if (ki < 0 || ki >= K) continue;                           // resolved at compile time
#pragma unroll
      for (int j = 0; j < K; ++j) {
        float wv[VEC];
        if constexpr (K == 3) {
#pragma unroll
          for (int v = 0; v < VEC; ++v) wv[v] = wr[ki * 3 + j][v];
        } else if constexpr (VEC == 4) {
          const float4 t = *reinterpret_cast<const float4*>(wts + (ki * K + j) * CB);
          wv[0] = t.x; wv[1] = t.y; wv[2] = t.z; wv[3] = t.w;
        } else {
#pragma unroll
          for (int v = 0; v < VEC; ++v) wv[v] = wts[(ki * K + j) * CB + v];
        }
        // scalar FMAs: packed fp32x2 pairs (as in the column walker) measured SLOWER in this kernel (multi-scale stage 160 -> 185 us)
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[py][v] = fmaf(wv[v], xv[j][v], acc[py][v]);
      }
    }
  }
}

// grid: x = (image, row tile, column tile, channel block), channel block (segment-major) fastest: CTAs that run at the same
// time consume whole pixel rows of C channels, not 128-byte slices 2 C bytes apart. 256 threads = 16 columns x 16 channel lanes.
template <typename T, int VEC, bool GELU>
__global__ void __launch_bounds__(DW_THREADS)
dwnhwc_stencil_kernel(const T* __restrict__ x, const T* __restrict__ aux, T* __restrict__ y, const DwnSegs sg, const DwnPlan pl,
                      int flip, int epi, int H, int W, int C) {
  constexpr int CB = DW_CL * VEC;
  __shared__ __align__(16) float s_w[49 * CB];
  const int chl = threadIdx.x % DW_CL, pxl = threadIdx.x / DW_CL;
  // segment and channel range of this CTA (selects, not a dynamic index into the parameter block)
  const int by = blockIdx.x % pl.nblk;
  int k = sg.k[0], cb = 0, ce = sg.cbeg[1], b0 = 0;
  const float* ws = sg.w[0];
  const float* bs = sg.b[0];
#pragma unroll
  for (int i = 1; i < kDwnMaxSeg; ++i)
    if (i < sg.nseg && by >= pl.blk0[i]) { k = sg.k[i]; cb = sg.cbeg[i]; ce = sg.cbeg[i + 1]; b0 = pl.blk0[i]; ws = sg.w[i]; bs = sg.b[i]; }
  const int cl0 = (by - b0) * CB;                    // first channel of the block inside its segment
  const int nvalid = min(CB, ce - cb - cl0);         // channels of the block that exist
  if (k > 3) {                                       // CTA-uniform: 5 x 5 / 7 x 7 weights go through shared memory
    const int kk = k * k;
    const float* wsrc = ws + (int64_t)cl0 * kk;
    for (int idx = threadIdx.x; idx < nvalid * kk; idx += DW_THREADS) {
      const int cc = idx / kk, t = idx - cc * kk;
      s_w[(flip ? kk - 1 - t : t) * CB + cc] = __ldg(wsrc + idx);
    }
    __syncthreads();
  }
  unsigned bx = blockIdx.x / pl.nblk;
  const int tw = bx % pl.tiles_w; bx /= pl.tiles_w;
  const int th = bx % pl.tiles_h;
  const int b = bx / pl.tiles_h;
  const int w = tw * DW_PL + pxl, h0 = th * DW_PY;
  const int cl = cl0 + chl * VEC;                    // first channel of this thread inside its segment
  if (w >= W || chl * VEC >= nvalid) return;
  const int WC = W * C;
  const int64_t o00 = ((int64_t)(b * H + h0) * W + w) * C + cb + cl;     // element (b, h0, w, c)

  float acc[DW_PY][VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    const float bv = (k > 1 && bs) ? __ldg(bs + cl + v) : 0.f;
#pragma unroll
    for (int py = 0; py < DW_PY; ++py) acc[py][v] = bv;
  }
  if (k == 3) dwn_taps<T, VEC, 3>(x + o00 - (WC + C), ws + (int64_t)cl * 9, flip, h0, w, H, W, WC, C, acc);
  else if (k == 5) dwn_taps<T, VEC, 5>(x + o00 - 2 * (WC + C), s_w + chl * VEC, flip, h0, w, H, W, WC, C, acc);
  else if (k == 7) dwn_taps<T, VEC, 7>(x + o00 - 3 * (WC + C), s_w + chl * VEC, flip, h0, w, H, W, WC, C, acc);

  const T* xo = x + o00;
  const T* ao = aux + o00;
  T* yo = y + o00;
#pragma unroll
  for (int py = 0; py < DW_PY; ++py) {
    if (h0 + py >= H) break;
    float out[VEC];
    if (k == 1) ldvec<T, VEC>(xo + py * WC, acc[py]);       // identity segment: acc = x
    if constexpr (GELU) {
      if (epi == 1) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) out[v] = gelu_phi<false>(acc[py][v], 0.f);
      } else {                                        // epi 3
        float av[VEC];
        ldvec<T, VEC>(ao + py * WC, av);
#pragma unroll
        for (int v = 0; v < VEC; ++v) out[v] = gelu_phi<true>(acc[py][v], av[v]);
      }
    } else {
      if (epi == 2) {
        float xc[VEC];
        ldvec<T, VEC>(xo + py * WC, xc);
#pragma unroll
        for (int v = 0; v < VEC; ++v) out[v] = xc[v] + acc[py][v];
      } else {
#pragma unroll
        for (int v = 0; v < VEC; ++v) out[v] = acc[py][v];
      }
    }
    stvec<T, VEC>(yo + py * WC, out);
  }
}

// ---- the 3 x 3 single-segment case (the DWConv of both FFNs: forward + GELU, GELU' pass, transposed pass) as a COLUMN WALKER ----
// A thread owns 4 channels of one image column and walks `rs` rows down it: per output pixel 3 new loads (the row below:
// columns w - 1, w, w + 1), the 3 x 3 x 4 neighbourhood in a register window (row slot = row mod 3, walk unrolled 3-fold), and
// the loads of the next THREE rows in flight as raw vectors (ring slot = row mod 3) while the current row is multiplied —
// the tiled kernel above issues one batch of 18 loads per short-lived CTA and waits for it (48 % issue utilisation,
// long_scoreboard 5.9 stall cycles per issue; profiles/r2_ncu_ffn_dw.txt). For epi 3 the ring slot of x row r carries aux row r - 1.
// grid.x = (image, row segment, column tile, channel block), channel block fastest.
struct DwnWalk { int nblk, tiles_w, nrs, rs; };

template <typename T>
constexpr int walk_ctas_per_sm() { return sizeof(T) == 2 ? 2 : 1; }

template <typename T, bool GELU>
__global__ void __launch_bounds__(DW_THREADS, walk_ctas_per_sm<T>())
dwnhwc_walk3_kernel(const T* __restrict__ x, const T* __restrict__ aux, T* __restrict__ y, const float* __restrict__ wgt,
                    const float* __restrict__ bias, const DwnWalk wk, int flip, int epi, int H, int W, int C) {
  constexpr int VEC = 4, CB = DW_CL * VEC;
  using RV = RawVec<T, VEC>;
  const int chl = threadIdx.x % DW_CL, pxl = threadIdx.x / DW_CL;
  unsigned bx = blockIdx.x;
  const int by = bx % wk.nblk; bx /= wk.nblk;
  const int tw = bx % wk.tiles_w; bx /= wk.tiles_w;
  const int rsi = bx % wk.nrs;
  const int b = bx / wk.nrs;
  const int c = by * CB + chl * VEC;
  const int w = tw * DW_PL + pxl;
  if (c >= C || w >= W) return;                                  // no barrier in this kernel
  const int r0 = rsi * wk.rs, r1 = min(H, r0 + wk.rs);           // output rows [r0, r1)
  const int WC = W * C;

  float wr[9][VEC], bv[VEC];
  {
    const float* wts = wgt + (int64_t)c * 9;
    float wf[VEC * 9];
    if ((reinterpret_cast<uintptr_t>(wts) & 15) == 0) {
#pragma unroll
      for (int q = 0; q < 9; ++q) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(wts) + q);
        wf[4 * q] = t.x; wf[4 * q + 1] = t.y; wf[4 * q + 2] = t.z; wf[4 * q + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int q = 0; q < VEC * 9; ++q) wf[q] = __ldg(wts + q);
    }
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int v = 0; v < VEC; ++v) wr[t][v] = flip ? wf[v * 9 + 8 - t] : wf[v * 9 + t];
#pragma unroll
    for (int v = 0; v < VEC; ++v) bv[v] = bias ? __ldg(bias + c + v) : 0.f;
  }
  bool colok[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) colok[j] = w + j - 1 >= 0 && w + j - 1 < W;
  const int64_t img = ((int64_t)b * H * W + w) * C + c;          // element (b, 0, w, c)
  const T* xc = x + img - C;                                     // column w - 1 of row 0
  const T* ac = aux + img;
  T* yc = y + img;
  const bool want_aux = GELU && epi == 3;

  RV ring[3][3], ringa[3];
  auto issue = [&](int r, RV (&dst)[3], RV& dsta) {              // x row r (columns w - 1 .. w + 1) and aux row r - 1
    const bool rok = r >= 0 && r < H && r <= r1;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      dst[j].zero();
      if (rok && colok[j]) dst[j].load(xc + (r * WC + j * C));
    }
    dsta.zero();
    if (want_aux && r - 1 >= r0 && r - 1 < r1) dsta.load(ac + (r - 1) * WC);
  };
  float win[3][3][VEC];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int v = 0; v < VEC; ++v) win[i][j][v] = 0.f;

  const int hs = r0 - 2;                                         // the walk starts two rows early: they fill the window
  issue(hs + 1, ring[1], ringa[1]);
  issue(hs + 2, ring[2], ringa[2]);
  issue(hs + 3, ring[0], ringa[0]);
  auto step = [&](int h, auto S) {
    constexpr int s = decltype(S)::value;                        // (h - hs) mod 3
    constexpr int sn = (s + 1) % 3;                              // slot of row h + 1
    float av[VEC];
#pragma unroll
    for (int j = 0; j < 3; ++j) ring[sn][j].get(win[sn][j]);
    ringa[sn].get(av);
    issue(h + 4, ring[sn], ringa[sn]);                           // consumed three steps from now
    if (h >= r0 && h < r1) {                                     // CTA-uniform
      // channel pairs on packed fp32x2 instructions: 18 FFMA2 instead of 36 FFMA
      float2 acc2[VEC / 2];
#pragma unroll
      for (int q = 0; q < VEC / 2; ++q) acc2[q] = make_float2(bv[2 * q], bv[2 * q + 1]);
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
          for (int q = 0; q < VEC / 2; ++q)
            acc2[q] = __ffma2_rn(make_float2(wr[i * 3 + j][2 * q], wr[i * 3 + j][2 * q + 1]),
                                 make_float2(win[(s + i + 2) % 3][j][2 * q], win[(s + i + 2) % 3][j][2 * q + 1]), acc2[q]);
      float out[VEC];
#pragma unroll
      for (int q = 0; q < VEC / 2; ++q) {
        float2 o2;
        if constexpr (GELU) {
          if (epi == 1) o2 = gelu_phi2<false>(acc2[q], make_float2(0.f, 0.f));
          else o2 = gelu_phi2<true>(acc2[q], make_float2(av[2 * q], av[2 * q + 1]));
        } else {
          o2 = epi == 2 ? make_float2(win[s][1][2 * q] + acc2[q].x, win[s][1][2 * q + 1] + acc2[q].y) : acc2[q];
        }
        out[2 * q] = o2.x; out[2 * q + 1] = o2.y;
      }
      stvec<T, VEC>(yc + h * WC, out);
    }
  };
  for (int h = hs; h < r1; h += 3) {
    step(h, std::integral_constant<int, 0>{});
    step(h + 1, std::integral_constant<int, 1>{});
    step(h + 2, std::integral_constant<int, 2>{});
  }
}

template <typename T>
static cudaError_t walk3_launch(const void* x, const void* aux, void* y, const DwnSegs& sg, int flip, int epi, int B, int H, int W,
                                int C, cudaStream_t stream) {
  constexpr int CB = DW_CL * 4;
  DwnWalk wk;
  wk.nblk = (C + CB - 1) / CB;
  wk.tiles_w = (W + DW_PL - 1) / DW_PL;
  // rows per walker against wave quantisation: every segment costs two window-filling steps
  const int64_t slots = (int64_t)sm_count_current_device() * walk_ctas_per_sm<T>();
  const int64_t base = (int64_t)B * wk.tiles_w * wk.nblk;
  int best = 1;
  int64_t best_cost = -1;
  for (int n = 1; n <= 16 && n <= H; ++n) {
    const int rs = (H + n - 1) / n;
    const int64_t cost = ((base * ((H + rs - 1) / rs) + slots - 1) / slots) * (rs + 2);
    if (best_cost < 0 || cost < best_cost) { best = n; best_cost = cost; }
  }
  wk.rs = (H + best - 1) / best;
  wk.nrs = (H + wk.rs - 1) / wk.rs;
  const int64_t ctas = base * wk.nrs;
  if (ctas >= (1ll << 31) || (int64_t)W * C * (H + 4) >= (1ll << 31)) return cudaErrorInvalidValue;
  const T* xs = static_cast<const T*>(x);
  const T* as = static_cast<const T*>(aux);
  T* ys = static_cast<T*>(y);
  if (epi == 1 || epi == 3)
    dwnhwc_walk3_kernel<T, true><<<(unsigned)ctas, DW_THREADS, 0, stream>>>(xs, as, ys, sg.w[0], sg.b[0], wk, flip, epi, H, W, C);
  else
    dwnhwc_walk3_kernel<T, false><<<(unsigned)ctas, DW_THREADS, 0, stream>>>(xs, as, ys, sg.w[0], sg.b[0], wk, flip, epi, H, W, C);
  return cudaGetLastError();
}

template <typename T, int VEC>
static cudaError_t stencil_launch_t(const void* x, const void* aux, void* y, const DwnSegs& sg, int flip, int epi, int B, int H,
                                    int W, int C, cudaStream_t stream) {
  constexpr int CB = DW_CL * VEC;
  DwnPlan pl;
  int nblk = 0;
  for (int s = 0; s < kDwnMaxSeg + 1; ++s) pl.blk0[s] = 0;
  for (int s = 0; s < sg.nseg; ++s) {
    pl.blk0[s] = nblk;
    nblk += (sg.cbeg[s + 1] - sg.cbeg[s] + CB - 1) / CB;
  }
  for (int s = sg.nseg; s < kDwnMaxSeg + 1; ++s) pl.blk0[s] = nblk;
  pl.tiles_w = (W + DW_PL - 1) / DW_PL;
  pl.tiles_h = (H + DW_PY - 1) / DW_PY;
  pl.nblk = nblk;
  const int64_t ctas = (int64_t)B * pl.tiles_w * pl.tiles_h * nblk;
  if (ctas >= (1ll << 31) || (int64_t)W * C * (DW_PY + 8) >= (1ll << 31)) return cudaErrorInvalidValue;
  dim3 grid((unsigned)ctas);
  const T* xs = static_cast<const T*>(x);
  const T* as = static_cast<const T*>(aux);
  T* ys = static_cast<T*>(y);
  if (epi == 1 || epi == 3) dwnhwc_stencil_kernel<T, VEC, true><<<grid, DW_THREADS, 0, stream>>>(xs, as, ys, sg, pl, flip, epi, H, W, C);
  else dwnhwc_stencil_kernel<T, VEC, false><<<grid, DW_THREADS, 0, stream>>>(xs, as, ys, sg, pl, flip, epi, H, W, C);
  return cudaGetLastError();
}

static int dwn_vec(int C, const int* bounds, int nb, size_t esz, const void* p0, const void* p1, const void* p2) {
  int vec = 4;
  auto fits = [&](int v) {
    if (C % v) return false;
    for (int i = 0; i < nb; ++i) if (bounds[i] % v) return false;
    const uintptr_t m = (uintptr_t)v * esz - 1;
    return !((reinterpret_cast<uintptr_t>(p0) & m) || (reinterpret_cast<uintptr_t>(p1) & m) || (reinterpret_cast<uintptr_t>(p2) & m));
  };
  while (vec > 1 && !fits(vec)) vec >>= 1;
  return vec;
}

template <typename T>
static cudaError_t stencil_launch_v(int vec, const void* x, const void* aux, void* y, const DwnSegs& sg, int flip, int epi, int B,
                                    int H, int W, int C, cudaStream_t stream) {
  if (vec == 4) return stencil_launch_t<T, 4>(x, aux, y, sg, flip, epi, B, H, W, C, stream);
  if (vec == 2) return stencil_launch_t<T, 2>(x, aux, y, sg, flip, epi, B, H, W, C, stream);
  return stencil_launch_t<T, 1>(x, aux, y, sg, flip, epi, B, H, W, C, stream);
}

cudaError_t dwnhwc_stencil_launch(const void* x, const void* aux, void* y, const DwnSegs& sg, int flip, int epi, int B, int H,
                                  int W, int C, int dt, cudaStream_t stream) {
  const int vec = dwn_vec(C, sg.cbeg, sg.nseg + 1, dt == SS2D_F32 ? 4 : 2, x, aux, y);
  if (vec == 4 && sg.nseg == 1 && sg.k[0] == 3 && scan_path_policy() != 3) {     // the DWConv of the FFNs: column walker (test hook 3: tiled kernel)
    if (dt == SS2D_F32) return walk3_launch<float>(x, aux, y, sg, flip, epi, B, H, W, C, stream);
    if (dt == SS2D_F16) return walk3_launch<__half>(x, aux, y, sg, flip, epi, B, H, W, C, stream);
    return walk3_launch<__nv_bfloat16>(x, aux, y, sg, flip, epi, B, H, W, C, stream);
  }
  if (dt == SS2D_F32) return stencil_launch_v<float>(vec, x, aux, y, sg, flip, epi, B, H, W, C, stream);
  if (dt == SS2D_F16) return stencil_launch_v<__half>(vec, x, aux, y, sg, flip, epi, B, H, W, C, stream);
  return stencil_launch_v<__nv_bfloat16>(vec, x, aux, y, sg, flip, epi, B, H, W, C, stream);
}

// ---- weight / bias gradient of one segment: dW[c][i][j] = sum_(b,h,w) g[b,h,w,c] x[b,h+i-P,w+j-P,c], db[c] = sum g ----
// CTA = 16 channel lanes (VEC channels each) x 16 pixel lanes; a pixel lane WALKS image rows left to right with the K x K
// neighbourhood of x in a register window: one new column (K loads) and one g per pixel instead of K * K + 1 loads with
// their address and border arithmetic (the first version: 290 instructions per pixel for 40 FMAs). The walk is unrolled
// K-fold so that the window's column slots are compile-time (slot = column mod K), with no branch inside (borders and the
// row's tail are load predicates), so the loads of the next pixels are issued ahead of the current pixel's FMAs.
// The 16 lanes of a CTA walk 16 adjacent image rows in lockstep: the rows above / below are the neighbours' centre rows
// (L1 hits). K * K + 1 sums per channel in registers, folded over the pixel lanes by one shuffle and a shared-memory
// pass; per-slab partials, fixed-order finalize (deterministic).
template <typename T, int K>
constexpr int wgrad_ctas_per_sm() { return (K == 3 && sizeof(T) == 2) ? 2 : 1; }     // register budget: 128 / 255 per thread

template <typename T, int VEC, int K>
__global__ void __launch_bounds__(DW_THREADS, wgrad_ctas_per_sm<T, K>())
dwnhwc_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ g, float* __restrict__ part, int c0, int nc, int rows, int H,
                    int W, int C, int rpl) {
  constexpr int NT = K * K + 1, P = K / 2, CB = DW_CL * VEC, TT = 10;
  const int chl = threadIdx.x % DW_CL, pxl = threadIdx.x / DW_CL;
  const int cl = blockIdx.x * CB + chl * VEC;
  const bool cok = cl < nc;
  const int c = c0 + (cok ? cl : 0);
  float acc[NT][VEC];
#pragma unroll
  for (int t = 0; t < NT; ++t)
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[t][v] = 0.f;
  if (cok) {
    for (int k = 0; k < rpl; ++k) {
      const int row = (blockIdx.y * rpl + k) * DW_PL + pxl;      // flattened (image, h)
      if (row >= rows) break;
      const int h = row % H;
      const T* grow = g + (int64_t)row * W * C + c;
      const T* xrow = x + (int64_t)row * W * C + c;
      bool rv[K];
#pragma unroll
      for (int i = 0; i < K; ++i) rv[i] = h + i - P >= 0 && h + i - P < H;
      float win[K][K][VEC];                                      // [kernel row][column slot], slot = column mod K
#pragma unroll
      for (int i = 0; i < K; ++i)
#pragma unroll
        for (int sl = 0; sl < K; ++sl)
#pragma unroll
          for (int v = 0; v < VEC; ++v) win[i][sl][v] = 0.f;
      // One "body" = K consecutive pixels: per pixel one g and the new window column (K loads), kept RAW (unconverted) so
      // that the next body's loads are in flight while the current body is multiplied. The walk starts one body before
      // the row (g masked), which fills the window's first columns through the same path.
      struct Step { RawVec<T, VEC> g, xn[K]; };
      auto load_step = [&](int w, Step& st) {
        const int cn = w + P;
        st.g.zero();
        if (w >= 0 && w < W) st.g.load(grow + (int64_t)w * C);
#pragma unroll
        for (int i = 0; i < K; ++i) {
          st.xn[i].zero();
          if (rv[i] && cn >= 0 && cn < W) st.xn[i].load(xrow + ((int64_t)(i - P) * W + cn) * C);
        }
      };
      auto compute_step = [&](const Step& st, auto S) {
        constexpr int s = decltype(S)::value;
        float gv[VEC];
        st.g.get(gv);
#pragma unroll
        for (int i = 0; i < K; ++i) st.xn[i].get(win[i][(s + K - 1) % K]);
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[K * K][v] += gv[v];
#pragma unroll
        for (int i = 0; i < K; ++i)
#pragma unroll
          for (int j = 0; j < K; ++j) {
            if constexpr (VEC >= 2) {                            // channel pairs on packed fp32x2 FMAs
#pragma unroll
              for (int q = 0; q < VEC / 2; ++q) {
                const float2 r = __ffma2_rn(make_float2(gv[2 * q], gv[2 * q + 1]),
                                            make_float2(win[i][(s + j) % K][2 * q], win[i][(s + j) % K][2 * q + 1]),
                                            make_float2(acc[i * K + j][2 * q], acc[i * K + j][2 * q + 1]));
                acc[i * K + j][2 * q] = r.x; acc[i * K + j][2 * q + 1] = r.y;
              }
            } else {
              acc[i * K + j][0] = fmaf(gv[0], win[i][(s + j) % K][0], acc[i * K + j][0]);
            }
          }
      };
      if constexpr (K == 3) {                                    // two bodies of three pixels in registers: ping-pong
        Step ba[3], bb[3];
#pragma unroll
        for (int s = 0; s < 3; ++s) load_step(-3 + s, ba[s]);
        for (int w0 = -3; w0 < W; w0 += 6) {
#pragma unroll
          for (int s = 0; s < 3; ++s) load_step(w0 + 3 + s, bb[s]);
          compute_step(ba[0], std::integral_constant<int, 0>{});
          compute_step(ba[1], std::integral_constant<int, 1>{});
          compute_step(ba[2], std::integral_constant<int, 2>{});
#pragma unroll
          for (int s = 0; s < 3; ++s) load_step(w0 + 6 + s, ba[s]);
          compute_step(bb[0], std::integral_constant<int, 0>{});
          compute_step(bb[1], std::integral_constant<int, 1>{});
          compute_step(bb[2], std::integral_constant<int, 2>{});
        }
      } else {                                                   // the wider windows leave registers for one pixel's loads
        for (int w0 = -K; w0 < W; w0 += K) {
          Step st;
          load_step(w0 + 0, st); compute_step(st, std::integral_constant<int, 0>{});
          load_step(w0 + 1, st); compute_step(st, std::integral_constant<int, 1>{});
          load_step(w0 + 2, st); compute_step(st, std::integral_constant<int, 2>{});
          load_step(w0 + 3, st); compute_step(st, std::integral_constant<int, 3>{});
          load_step(w0 + 4, st); compute_step(st, std::integral_constant<int, 4>{});
          if constexpr (K == 7) {
            load_step(w0 + 5, st); compute_step(st, std::integral_constant<int, 5>{});
            load_step(w0 + 6, st); compute_step(st, std::integral_constant<int, 6>{});
          }
        }
      }
    }
  }
  // fold the 16 pixel lanes: lanes l and l ^ 16 of a warp are two pixel lanes of one channel lane
#pragma unroll
  for (int t = 0; t < NT; ++t)
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[t][v] += __shfl_xor_sync(0xffffffffu, acc[t][v], 16);
  __shared__ float s_red[DW_THREADS / 32][TT][CB];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int t0 = 0; t0 < NT; t0 += TT) {
    if (lane < 16) {
#pragma unroll
      for (int tt = 0; tt < TT; ++tt)
        if (t0 + tt < NT) {
#pragma unroll
          for (int v = 0; v < VEC; ++v) s_red[warp][tt][chl * VEC + v] = acc[t0 + tt][v];
        }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < TT * CB; idx += DW_THREADS) {
      const int tt = idx / CB, cc = idx - tt * CB;
      const int ch = blockIdx.x * CB + cc;
      if (t0 + tt < NT && ch < nc) {
        float s = 0.f;
#pragma unroll
        for (int wi = 0; wi < DW_THREADS / 32; ++wi) s += s_red[wi][tt][cc];
        part[((int64_t)blockIdx.y * nc + ch) * NT + t0 + tt] = s;
      }
    }
    __syncthreads();
  }
}

__global__ void dwnhwc_wgrad_finalize_kernel(const float* __restrict__ part, float* __restrict__ dW, float* __restrict__ db, int nc,
                                             int T, int slabs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nc * T) return;
  float v = 0.f;
  for (int s = 0; s < slabs; ++s) v += part[(int64_t)s * nc * T + i];
  const int cl = i / T, t = i - cl * T;
  if (t < T - 1) dW[cl * (T - 1) + t] = v;
  else if (db) db[cl] = v;
}

// Row geometry of the reduction: 16-row groups (one image row per pixel lane), `rpl` groups per CTA, `slabs` CTAs along the
// rows. The workspace query returns the upper bound (one group per CTA); the launch picks rpl against wave quantisation:
// the smallest (waves x rpl) for the CTA slots of the instantiation, ties to the larger rpl (fewer partials).
static int64_t dwn_row_groups(int B, int H) { return ((int64_t)B * H + DW_PL - 1) / DW_PL; }
int dwnhwc_wgrad_slabs(int B, int H, int W, int nc) {
  (void)W; (void)nc;
  const int64_t groups = dwn_row_groups(B, H);
  return (int)(groups < 65535 ? groups : 65535);
}
static int dwn_pick_rpl(int64_t groups, int cblocks, int ctas_per_sm) {
  const int64_t slots = (int64_t)sm_count_current_device() * ctas_per_sm;
  const int rmin = (int)((groups + 65534) / 65535);
  int best = rmin;
  int64_t best_cost = -1;
  for (int r = rmin; r < rmin + 8; ++r) {
    const int64_t ctas = ((groups + r - 1) / r) * cblocks;
    const int64_t cost = ((ctas + slots - 1) / slots) * r;
    if (best_cost < 0 || cost <= best_cost) { best = r; best_cost = cost; }
  }
  return best;
}

template <typename T, int VEC, int K>
static int wgrad_launch_k(const void* x, const void* g, float* part, int c0, int nc, int B, int H, int W, int C, cudaStream_t stream) {
  constexpr int CB = DW_CL * VEC;
  const int cblocks = (nc + CB - 1) / CB;
  const int64_t groups = dwn_row_groups(B, H);
  const int rpl = dwn_pick_rpl(groups, cblocks, wgrad_ctas_per_sm<T, K>());
  const int slabs = (int)((groups + rpl - 1) / rpl);
  dim3 grid(cblocks, slabs);
  dwnhwc_wgrad_kernel<T, VEC, K><<<grid, DW_THREADS, 0, stream>>>(static_cast<const T*>(x), static_cast<const T*>(g), part, c0, nc,
                                                                  B * H, H, W, C, rpl);
  return slabs;
}

template <typename T>
static int wgrad_launch_t(int vec, int K, const void* x, const void* g, float* part, int c0, int nc, int B, int H, int W, int C,
                          cudaStream_t stream) {
  // registers: (K * K + 1) * VEC sums + the K * K * VEC window per thread -> 4 channels for 3 x 3, 2 for 5 x 5 / 7 x 7
  if (K == 3) {
    if (vec == 4) return wgrad_launch_k<T, 4, 3>(x, g, part, c0, nc, B, H, W, C, stream);
    if (vec == 2) return wgrad_launch_k<T, 2, 3>(x, g, part, c0, nc, B, H, W, C, stream);
    return wgrad_launch_k<T, 1, 3>(x, g, part, c0, nc, B, H, W, C, stream);
  }
  if (K == 5) {
    if (vec >= 2) return wgrad_launch_k<T, 2, 5>(x, g, part, c0, nc, B, H, W, C, stream);
    return wgrad_launch_k<T, 1, 5>(x, g, part, c0, nc, B, H, W, C, stream);
  }
  if (vec >= 2) return wgrad_launch_k<T, 2, 7>(x, g, part, c0, nc, B, H, W, C, stream);
  return wgrad_launch_k<T, 1, 7>(x, g, part, c0, nc, B, H, W, C, stream);
}

cudaError_t dwnhwc_wgrad_launch(const void* x, const void* g, int c0, int nc, int K, float* dW, float* db, int B, int H, int W,
                                int C, int dt, float* part, cudaStream_t stream) {
  const int bounds[2] = {c0, c0 + nc};
  const int vec = dwn_vec(C, bounds, 2, dt == SS2D_F32 ? 4 : 2, x, g, nullptr);
  int slabs;
  if (dt == SS2D_F32) slabs = wgrad_launch_t<float>(vec, K, x, g, part, c0, nc, B, H, W, C, stream);
  else if (dt == SS2D_F16) slabs = wgrad_launch_t<__half>(vec, K, x, g, part, c0, nc, B, H, W, C, stream);
  else slabs = wgrad_launch_t<__nv_bfloat16>(vec, K, x, g, part, c0, nc, B, H, W, C, stream);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const int T = K * K + 1;
  dwnhwc_wgrad_finalize_kernel<<<(nc * T + 255) / 256, 256, 0, stream>>>(part, dW, db, nc, T, slabs);
  return cudaGetLastError();
}

}  // namespace ss2d
