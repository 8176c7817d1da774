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

namespace ss2d {

constexpr int kDwThreads = 256;

__global__ void __launch_bounds__(kDwThreads)
dwconv3_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ part, int batch, int C,
                     int H, int W, int slabs) {
  const int c = blockIdx.x, slab = blockIdx.y;
  const int HW = H * W;
  const int64_t total = (int64_t)batch * HW;                 // pixels of this channel over the batch
  const int64_t per = (total + slabs - 1) / slabs;
  const int64_t p0 = (int64_t)slab * per, p1 = p0 + per < total ? p0 + per : total;
  float acc[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) acc[i] = 0.f;
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

}  // namespace ss2d
