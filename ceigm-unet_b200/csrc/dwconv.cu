// Weight (and bias) gradient of SS2D's depthwise 3 x 3 convolution (/root/reference/gm-unet/model/gm/ss2d.py:316-325,
// 512: nn.Conv2d(d_inner, d_inner, groups=d_inner, kernel_size=3, padding=1, bias=True)):
//     dW[c][ky][kx] = sum over (b, h, w) of dY[b,c,h,w] * X[b,c,h+ky-1,w+kx-1]      db[c] = sum dY[b,c,h,w]
// Ten numbers per channel out of a reduction over B*H*W = 75 264 pixels at stage 1 of a 224^2 batch-24 step: another
// tall reduction the library handles poorly (conv_depthwise2d_grad_weight: 64 us per call, 11 % of a graphed
// GroupMambaLayer, plus a separate 25 us reduction for the bias). Here: grid (channel, slab); a thread walks pixels of
// its channel's planes (coalesced along w, the eight neighbours come from L1), keeps the ten sums in registers; warp
// butterflies + one shared-memory fold per CTA, per-slab partials, fixed-order finalize (deterministic).
// HBM-bound: X and dY are read once (4 bytes per pixel-channel each). fp32, NCHW contiguous.
#include "common.cuh"
#include "host_util.h"

namespace ss2d {

constexpr int kDwThreads = 256;

__global__ void __launch_bounds__(kDwThreads)
dwconv3_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ part, int batch, int C,
                     int H, int W, int slabs) {
  const int c = blockIdx.x, slab = blockIdx.y;
  const int HW = H * W;
  const int64_t total = (int64_t)batch * HW;                 // pixels of this channel over the batch
  const int64_t per = ((total + slabs - 1) / slabs + 3) & ~(int64_t)3;      // slabs start on pixel quads
  const int64_t p0 = (int64_t)slab * per < total ? (int64_t)slab * per : total, p1 = p0 + per < total ? p0 + per : total;
  float acc[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) acc[i] = 0.f;
  if ((W & 3) == 0 && (p0 & 3) == 0) {
    // four consecutive pixels of a row per thread: one vector load per input row (+ the two edge neighbours, L1 hits) and
    // one of dy instead of forty scalar loads (the scalar walk below ran at 0.84 TB/s: 137 us at B = 24, C = 192, 56^2)
    for (int64_t p = p0 + 4 * (int64_t)threadIdx.x; p < p1; p += 4 * kDwThreads) {
      const int b = (int)(p / HW), hw = (int)(p - (int64_t)b * HW);
      const int h = hw / W, w = hw - h * W;
      const float* xp = x + ((int64_t)b * C + c) * HW;
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(dy + ((int64_t)b * C + c) * HW + hw));
      const float g[4] = {g4.x, g4.y, g4.z, g4.w};
      const int nq = (int)(p1 - p < 4 ? p1 - p : 4);         // the slab may end inside the quad
      float gm[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) { gm[q] = q < nq ? g[q] : 0.f; acc[9] += gm[q]; }
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int hh = h + ky - 1;
        if (hh < 0 || hh >= H) continue;
        const float* row = xp + hh * W + w;
        const float4 m = __ldg(reinterpret_cast<const float4*>(row));
        const float v[6] = {w > 0 ? __ldg(row - 1) : 0.f, m.x, m.y, m.z, m.w, w + 4 < W ? __ldg(row + 4) : 0.f};
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[ky * 3 + kx] = fmaf(gm[q], v[q + kx], acc[ky * 3 + kx]);
      }
    }
  } else
  for (int64_t p = p0 + threadIdx.x; p < p1; p += kDwThreads) {
    const int b = (int)(p / HW), hw = (int)(p - (int64_t)b * HW);
    const int h = hw / W, w = hw - h * W;
    const float* xp = x + ((int64_t)b * C + c) * HW;
    const float g = __ldg(dy + ((int64_t)b * C + c) * HW + hw);
    acc[9] += g;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int hh = h + ky - 1;
      if (hh < 0 || hh >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ww = w + kx - 1;
        if (ww >= 0 && ww < W) acc[ky * 3 + kx] = fmaf(g, __ldg(xp + hh * W + ww), acc[ky * 3 + kx]);
      }
    }
  }
  __shared__ float s_red[kDwThreads / 32][10];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    float v = acc[i];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_red[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < 10) {
    float v = 0.f;
    for (int k = 0; k < kDwThreads / 32; ++k) v += s_red[k][threadIdx.x];
    part[((int64_t)c * slabs + slab) * 10 + threadIdx.x] = v;
  }
}

__global__ void dwconv3_wgrad_finalize_kernel(const float* __restrict__ part, float* __restrict__ dW, float* __restrict__ db,
                                              int C, int slabs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * 10) return;
  const int c = i / 10, k = i - c * 10;
  float s = 0.f;
  for (int j = 0; j < slabs; ++j) s += part[((int64_t)c * slabs + j) * 10 + k];
  if (k < 9) dW[c * 9 + k] = s;
  else if (db) db[c] = s;
}

int dwconv3_wgrad_slabs(int batch, int C, int H, int W) {
  const int64_t total = (int64_t)batch * H * W;
  int64_t s = (148 * 8 + C - 1) / C;                       // about 8 CTAs per SM in total
  const int64_t most = (total + kDwThreads - 1) / kDwThreads;   // at least one pass of the block per slab
  if (s > most) s = most;
  if (s > 512) s = 512;
  return s < 1 ? 1 : (int)s;
}

cudaError_t dwconv3_wgrad_launch(const float* x, const float* dy, float* dW, float* db, float* workspace, int batch, int C,
                                 int H, int W, cudaStream_t stream) {
  const int slabs = dwconv3_wgrad_slabs(batch, C, H, W);
  dwconv3_wgrad_kernel<<<dim3(C, slabs), kDwThreads, 0, stream>>>(x, dy, workspace, batch, C, H, W, slabs);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  dwconv3_wgrad_finalize_kernel<<<(C * 10 + 255) / 256, 256, 0, stream>>>(workspace, dW, db, C, slabs);
  return cudaGetLastError();
}

// ---- depthwise 3 x 3 convolution fused with its activation (ss2d.py:512-513: conv2d -> SiLU) ------------------------
// MODE 0 (forward):        y  = SiLU(bias + conv3x3(x))
// MODE 1 (backward, 1/2):  y  = dy * SiLU'(bias + conv3x3(x))     the gradient of the pre-activation, recomputed from x
// MODE 2 (backward, 2/2):  y  = conv3x3 of x with the FLIPPED kernel, no bias / activation (the input gradient from MODE 1's
//                               output; zero padding makes the transposed convolution a plain correlation)
// HBM-bound stencils: a thread produces PX consecutive pixels of one row (PX = 4 when W % 4 == 0: one vector load per row
// plus the two edge neighbours, which hit L1), every input element is fetched from HBM once. fp32 math; fp32 / bf16 / fp16
// tensors, NCHW contiguous. Replaces cuDNN / ATen depthwise conv forward + SiLU (2 kernels, one extra round trip of the
// activation tensor) and SiLU backward + conv input gradient.
template <int PX, int MODE>
__global__ void __launch_bounds__(256)
dwconv3_fused_kernel(const void* __restrict__ x, const float* __restrict__ wgt, const float* __restrict__ bias,
                     const void* __restrict__ dy, void* __restrict__ y, int64_t planes, int C, int H, int W, int dt) {
  const int WQ = W / PX;
  const int64_t per_plane = (int64_t)H * WQ, total = planes * per_plane;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t pl = i / per_plane;
    const int r = (int)(i - pl * per_plane);
    const int h = r / WQ, w0 = (r - h * WQ) * PX;
    const int c = (int)(pl % C);
    float k[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) k[j] = __ldg(wgt + c * 9 + (MODE == 2 ? 8 - j : j));
    const float b0 = (MODE != 2 && bias) ? __ldg(bias + c) : 0.f;
    float acc[PX];
#pragma unroll
    for (int q = 0; q < PX; ++q) acc[q] = b0;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int hh = h + ky - 1;
      if (hh < 0 || hh >= H) continue;
      const int64_t ro = (pl * H + hh) * W + w0;
      float v[PX + 2];
      v[0] = w0 > 0 ? load1(x, ro - 1, dt) : 0.f;
      if (PX == 4) {
        const float4 m = load4(x, ro, dt);
        v[1] = m.x; v[2] = m.y; v[3] = m.z; v[4] = m.w;
      } else {
        v[1] = load1(x, ro, dt);
      }
      v[PX + 1] = w0 + PX < W ? load1(x, ro + PX, dt) : 0.f;
#pragma unroll
      for (int q = 0; q < PX; ++q)
        acc[q] = fmaf(k[ky * 3 + 2], v[q + 2], fmaf(k[ky * 3 + 1], v[q + 1], fmaf(k[ky * 3], v[q], acc[q])));
    }
    const int64_t o = (pl * H + h) * W + w0;
    float g[PX];
    if (MODE == 1) {
      if (PX == 4) { const float4 t = load4(dy, o, dt); g[0] = t.x; g[1] = t.y; g[2] = t.z; g[3] = t.w; }
      else g[0] = load1(dy, o, dt);
    }
#pragma unroll
    for (int q = 0; q < PX; ++q) {
      if (MODE == 2) continue;
      const float sgm = __fdividef(1.f, 1.f + ex2f(-acc[q] * kLog2e));
      acc[q] = MODE == 0 ? acc[q] * sgm : g[q] * sgm * (1.f + acc[q] * (1.f - sgm));
    }
    if (PX == 4) store4(y, o, dt, make_float4(acc[0], acc[1], acc[2], acc[3]));
    else store1(y, o, dt, acc[0]);
  }
}

// Tiled variant for H % 4 == 0 and W % 4 == 0: a CTA owns a 32 x 32 pixel tile of one (batch, channel) plane, a thread
// 4 consecutive pixels of a row. On top of the flat kernel it can
//   * (MODE 0) store the result a second time in TRANSPOSED pixel order (the (W, H) image the column-major scan directions
//     2 / 4 traverse as row-major: model/gm/csms6s.py:95-129, 172-206) — both orientations leave in 128-byte row pieces, the
//     transposition goes through a padded shared-memory tile, so no separate permuted-copy pass (and no torch.cat of the two
//     planes: y and yT are the two halves of ONE (B, 2 C, L) buffer addressed by their batch strides) ever runs;
//   * (MODE 1) read a second upstream gradient dyT given in transposed pixel order and add it to dy on the fly — the adjoint
//     of the double store, instead of a transpose-copy + add.
template <int MODE>
__global__ void __launch_bounds__(256)
dwconv3_tiled_kernel(const void* __restrict__ x, const float* __restrict__ wgt, const float* __restrict__ bias,
                     const void* __restrict__ dy, const void* __restrict__ dyT, void* __restrict__ y, void* __restrict__ yT,
                     int C, int H, int W, int tiles_w, int64_t x_bs, int64_t y_bs, int64_t yT_bs, int64_t dy_bs, int64_t dyT_bs,
                     int dt) {
  __shared__ float tile[32][33];
  const int c = blockIdx.y, b = blockIdx.z;
  const int th = blockIdx.x / tiles_w, tw = blockIdx.x - th * tiles_w;
  const int h0 = th * 32, w0 = tw * 32;
  const int ty = threadIdx.x >> 3, tx = threadIdx.x & 7;
  const int h = h0 + ty, w = w0 + 4 * tx;
  const int64_t HW = (int64_t)H * W;
  const bool in = h < H && w < W;                       // whole quads are inside or outside (W % 4 == 0)
  if (MODE == 1 && dyT) {                               // transposed upstream gradient of this tile: rows = w, columns = h
    const int wt = w0 + ty, ht = h0 + 4 * tx;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (wt < W && ht < H) t = load4(dyT, (int64_t)b * dyT_bs + (int64_t)c * HW + (int64_t)wt * H + ht, dt);
    tile[ty][4 * tx] = t.x; tile[ty][4 * tx + 1] = t.y; tile[ty][4 * tx + 2] = t.z; tile[ty][4 * tx + 3] = t.w;
    __syncthreads();
  }
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (in) {
    float k[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) k[j] = __ldg(wgt + c * 9 + j);
    const float b0 = bias ? __ldg(bias + c) : 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q] = b0;
    const int64_t xb = (int64_t)b * x_bs + (int64_t)c * HW;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int hh = h + ky - 1;
      if (hh < 0 || hh >= H) continue;
      const int64_t ro = xb + (int64_t)hh * W + w;
      float v[6];
      v[0] = w > 0 ? load1(x, ro - 1, dt) : 0.f;
      const float4 m = load4(x, ro, dt);
      v[1] = m.x; v[2] = m.y; v[3] = m.z; v[4] = m.w;
      v[5] = w + 4 < W ? load1(x, ro + 4, dt) : 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        acc[q] = fmaf(k[ky * 3 + 2], v[q + 2], fmaf(k[ky * 3 + 1], v[q + 1], fmaf(k[ky * 3], v[q], acc[q])));
    }
    float g[4] = {0.f, 0.f, 0.f, 0.f};
    if (MODE == 1) {
      const float4 t = load4(dy, (int64_t)b * dy_bs + (int64_t)c * HW + (int64_t)h * W + w, dt);
      g[0] = t.x; g[1] = t.y; g[2] = t.z; g[3] = t.w;
      if (dyT) {
#pragma unroll
        for (int q = 0; q < 4; ++q) g[q] += tile[4 * tx + q][ty];
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float sgm = __fdividef(1.f, 1.f + ex2f(-acc[q] * kLog2e));
      acc[q] = MODE == 0 ? acc[q] * sgm : g[q] * sgm * (1.f + acc[q] * (1.f - sgm));
    }
    store4(y, (int64_t)b * y_bs + (int64_t)c * HW + (int64_t)h * W + w, dt, make_float4(acc[0], acc[1], acc[2], acc[3]));
  }
  if (MODE == 0 && yT) {
    tile[ty][4 * tx] = acc[0]; tile[ty][4 * tx + 1] = acc[1]; tile[ty][4 * tx + 2] = acc[2]; tile[ty][4 * tx + 3] = acc[3];
    __syncthreads();
    const int wt = w0 + ty, ht = h0 + 4 * tx;            // this thread now owns column wt, rows ht .. ht + 3
    if (wt < W && ht < H)
      store4(yT, (int64_t)b * yT_bs + (int64_t)c * HW + (int64_t)wt * H + ht, dt,
             make_float4(tile[4 * tx][ty], tile[4 * tx + 1][ty], tile[4 * tx + 2][ty], tile[4 * tx + 3][ty]));
  }
}

cudaError_t dwconv3_tiled_launch(int mode, const void* x, const float* wgt, const float* bias, const void* dy, const void* dyT,
                                 void* y, void* yT, int batch, int C, int H, int W, int64_t x_bs, int64_t y_bs, int64_t yT_bs,
                                 int64_t dy_bs, int64_t dyT_bs, int dt, cudaStream_t stream) {
  const int tiles_w = (W + 31) / 32, tiles_h = (H + 31) / 32;
  dim3 grid(tiles_w * tiles_h, C, batch);
  if (mode == 0)
    dwconv3_tiled_kernel<0><<<grid, 256, 0, stream>>>(x, wgt, bias, dy, dyT, y, yT, C, H, W, tiles_w, x_bs, y_bs, yT_bs, dy_bs, dyT_bs, dt);
  else
    dwconv3_tiled_kernel<1><<<grid, 256, 0, stream>>>(x, wgt, bias, dy, dyT, y, yT, C, H, W, tiles_w, x_bs, y_bs, yT_bs, dy_bs, dyT_bs, dt);
  return cudaGetLastError();
}

cudaError_t dwconv3_fused_launch(int mode, const void* x, const float* wgt, const float* bias, const void* dy, void* y,
                                 int batch, int C, int H, int W, int dt, cudaStream_t stream) {
  const int64_t planes = (int64_t)batch * C;
  const bool v4 = (W & 3) == 0;      // plane rows then start on 16-byte (fp32) / 8-byte (16-bit) boundaries
  const int64_t total = planes * H * (v4 ? W / 4 : W);
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count_current_device() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
#define SS2D_DW(PXV, M) dwconv3_fused_kernel<PXV, M><<<(unsigned)blocks, 256, 0, stream>>>(x, wgt, bias, dy, y, planes, C, H, W, dt)
  if (v4) { if (mode == 0) SS2D_DW(4, 0); else if (mode == 1) SS2D_DW(4, 1); else SS2D_DW(4, 2); }
  else { if (mode == 0) SS2D_DW(1, 0); else if (mode == 1) SS2D_DW(1, 1); else SS2D_DW(1, 2); }
#undef SS2D_DW
  return cudaGetLastError();
}

}  // namespace ss2d
