// Stand-alone cross-scan / cross-merge permutation kernels (the b2 autograd-Function API).
// Semantics: /root/reference/gm-unet/model/gm/csms6s.py:11-206. Exact data movement, bit-exact by construction.
// The fused path (NATURAL layout in scan_fwd.cu / scan_bwd.cu) never calls these.
#include "common.cuh"

namespace ss2d {

struct Dirs { int d[SS2D_MAX_GROUP_DIRS]; };

// scan position of natural pixel (h, w) in direction dir
__device__ __forceinline__ int scan_pos(int dir, int h, int w, int H, int W) {
  const int L = H * W;
  switch (dir) {
    case 1: return h * W + w;
    case 2: return w * H + h;
    case 3: return L - 1 - (h * W + w);
    default: return L - 1 - (w * H + h);
  }
}

// xs[b][k][c][l] = x[b][c][natural_k(l)]
__global__ void cross_scan_kernel(const void* __restrict__ x, void* __restrict__ xs, int batch, int channels, int H,
                                  int W, int K, Dirs dirs, int dtype) {
  const int L = H * W;
  const int64_t total = (int64_t)batch * K * channels * L;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int l = (int)(i % L);
    const int64_t row = i / L;                 // (b * K + k) * channels + c
    const int c = (int)(row % channels);
    const int k = (int)((row / channels) % K);
    const int b = (int)(row / ((int64_t)channels * K));
    ScanOrder so; so.dir = dirs.d[k]; so.H = H; so.W = W; so.L = L;
    const int64_t src = ((int64_t)b * channels + c) * L + so.natural(l);
    if (dtype == SS2D_F32) reinterpret_cast<float*>(xs)[i] = __ldg(reinterpret_cast<const float*>(x) + src);
    else reinterpret_cast<uint16_t*>(xs)[i] = __ldg(reinterpret_cast<const uint16_t*>(x) + src);
  }
}

// y[b][c][p] = sum_k ys[b][k][c][scan_pos_k(p)], K == 4 in the association (k0 + k2) + (k1 + k3)
__global__ void cross_merge_kernel(const void* __restrict__ ys, void* __restrict__ y, int batch, int channels, int H,
                                   int W, int K, Dirs dirs, int dtype) {
  const int L = H * W;
  const int64_t total = (int64_t)batch * channels * L;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int pix = (int)(i % L);
    const int64_t row = i / L;                 // b * channels + c
    const int c = (int)(row % channels);
    const int b = (int)(row / channels);
    const int h = pix / W, w = pix - h * W;
    float v[SS2D_MAX_GROUP_DIRS];
    for (int k = 0; k < K; ++k) {
      const int64_t src = (((int64_t)b * K + k) * channels + c) * L + scan_pos(dirs.d[k], h, w, H, W);
      v[k] = load1(ys, src, dtype);
    }
    float acc;
    if (K == 4) {
      // every partial sum is rounded to the tensor dtype, exactly like the reference's tensor adds
      float a = v[0] + v[2], bb = v[1] + v[3];
      if (dtype == SS2D_F16) { a = __half2float(__float2half_rn(a)); bb = __half2float(__float2half_rn(bb)); }
      if (dtype == SS2D_BF16) { a = __bfloat162float(__float2bfloat16_rn(a)); bb = __bfloat162float(__float2bfloat16_rn(bb)); }
      acc = a + bb;
    } else {
      acc = v[0];
      for (int k = 1; k < K; ++k) acc += v[k];
    }
    store1(y, i, dtype, acc);
  }
}

static int grid_for(int64_t total) {
  int64_t g = (total + 255) / 256;
  const int64_t cap = 148 * 16;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

cudaError_t cross_scan_launch(const void* x, void* xs, int batch, int channels, int H, int W, int K, const int* dirs,
                              int dtype, cudaStream_t stream) {
  Dirs d{};
  for (int k = 0; k < K; ++k) d.d[k] = dirs[k];
  cross_scan_kernel<<<grid_for((int64_t)batch * K * channels * H * W), 256, 0, stream>>>(x, xs, batch, channels, H, W, K,
                                                                                         d, dtype);
  return cudaGetLastError();
}

cudaError_t cross_merge_launch(const void* ys, void* y, int batch, int channels, int H, int W, int K, const int* dirs,
                               int dtype, cudaStream_t stream) {
  Dirs d{};
  for (int k = 0; k < K; ++k) d.d[k] = dirs[k];
  cross_merge_kernel<<<grid_for((int64_t)batch * channels * H * W), 256, 0, stream>>>(ys, y, batch, channels, H, W, K, d,
                                                                                      dtype);
  return cudaGetLastError();
}

}  // namespace ss2d
