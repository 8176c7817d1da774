// Row-wise LayerNorm over C <= 512 channels of channels-last rows, forward and backward: GroupMambaLayer.norm, applied
// twice per layer call with shared weights (/root/reference/gm-unet/model/gm/groupmamba.py:131, 156; nn.LayerNorm).
// At C = 64 ... 448 the library kernel runs at 0.35 TB/s on the (B L, C) rows of a 224^2 batch-24 step (110 us for 19 MB
// in + 19 MB out at stage 1: 8 % of a graphed layer). Here: one warp per row, the row held in registers (<= 16 values per
// lane, coalesced 128-byte steps), two-pass statistics with warp butterflies, no shared memory in the row loop. The
// backward keeps per-lane column sums of dy xn and dy across the rows its warp visits, folds the block's 8 warps through
// shared memory at the end and writes one partial row per CTA (the caller sums the partials: deterministic).
// HBM-bound: forward reads C and writes C (+ 2 statistics) per row, backward reads 2 C and writes C.
#include "common.cuh"

namespace ss2d {

constexpr int kLnThreads = 256;
constexpr int kLnMaxV = 16;            // values per lane -> C <= 512

template <int NV>
__global__ void __launch_bounds__(kLnThreads)
layernorm_fwd_kernel(const void* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                     void* __restrict__ y, float* __restrict__ mean_rstd, int64_t rows, int C, float eps, int dt) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * (kLnThreads / 32) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (kLnThreads / 32);
  float wv[NV], bv[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    wv[i] = (w && c < C) ? __ldg(w + c) : 1.f;
    bv[i] = (b && c < C) ? __ldg(b + c) : 0.f;
  }
  for (int64_t r = warp0; r < rows; r += nwarps) {
    float v[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      v[i] = c < C ? load1(x, r * C + c, dt) : 0.f;
      s += v[i];
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float t = (lane + 32 * i) < C ? v[i] - mean : 0.f;
      q = fmaf(t, t, q);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / C + eps);
    if (lane == 0 && mean_rstd) { mean_rstd[r * 2] = mean; mean_rstd[r * 2 + 1] = rstd; }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      if (c < C) store1(y, r * C + c, dt, fmaf((v[i] - mean) * rstd, wv[i], bv[i]));
    }
  }
}

template <int NV>
__global__ void __launch_bounds__(kLnThreads)
layernorm_bwd_kernel(const void* __restrict__ x, const float* __restrict__ w, const void* __restrict__ dy,
                     const float* __restrict__ mean_rstd, void* __restrict__ dx, float* __restrict__ dw_part,
                     float* __restrict__ db_part, int64_t rows, int C, int dt) {
  __shared__ float s_red[kLnThreads / 32][2][32 * NV];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t warp0 = (int64_t)blockIdx.x * (kLnThreads / 32) + warp;
  const int64_t nwarps = (int64_t)gridDim.x * (kLnThreads / 32);
  float wv[NV], aw[NV], ab[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    wv[i] = (w && c < C) ? __ldg(w + c) : 1.f;
    aw[i] = 0.f; ab[i] = 0.f;
  }
  for (int64_t r = warp0; r < rows; r += nwarps) {
    const float mean = mean_rstd[r * 2], rstd = mean_rstd[r * 2 + 1];
    float xn[NV], g[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      if (c < C) {
        xn[i] = (load1(x, r * C + c, dt) - mean) * rstd;
        const float go = load1(dy, r * C + c, dt);
        aw[i] = fmaf(go, xn[i], aw[i]);
        ab[i] += go;
        g[i] = go * wv[i];
        s1 += g[i];
        s2 = fmaf(g[i], xn[i], s2);
      } else { xn[i] = 0.f; g[i] = 0.f; }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    s1 /= C; s2 /= C;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      if (c < C) store1(dx, r * C + c, dt, rstd * (g[i] - s1 - xn[i] * s2));
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) { s_red[warp][0][lane + 32 * i] = aw[i]; s_red[warp][1][lane + 32 * i] = ab[i]; }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += kLnThreads) {
    float tw = 0.f, tb = 0.f;
#pragma unroll
    for (int k = 0; k < kLnThreads / 32; ++k) { tw += s_red[k][0][c]; tb += s_red[k][1][c]; }
    dw_part[(int64_t)blockIdx.x * C + c] = tw;
    db_part[(int64_t)blockIdx.x * C + c] = tb;
  }
}

int layernorm_bwd_partials(int64_t rows) {
  const int64_t blocks = (rows + kLnThreads / 32 - 1) / (kLnThreads / 32);
  return (int)(blocks < 148 * 4 ? (blocks < 1 ? 1 : blocks) : 148 * 4);
}
int layernorm_max_C() { return 32 * kLnMaxV; }

cudaError_t layernorm_fwd_launch(const void* x, const float* w, const float* b, void* y, float* mean_rstd, int64_t rows,
                                 int C, float eps, int dt, cudaStream_t stream) {
  const int64_t blocks = (rows + kLnThreads / 32 - 1) / (kLnThreads / 32);
  const int grid = (int)(blocks < 148 * 8 ? (blocks < 1 ? 1 : blocks) : 148 * 8);
  const int nv = (C + 31) / 32;
#define SS2D_LN_FWD(NV) layernorm_fwd_kernel<NV><<<grid, kLnThreads, 0, stream>>>(x, w, b, y, mean_rstd, rows, C, eps, dt)
  if (nv <= 2) SS2D_LN_FWD(2);
  else if (nv <= 4) SS2D_LN_FWD(4);
  else if (nv <= 8) SS2D_LN_FWD(8);
  else if (nv <= 12) SS2D_LN_FWD(12);
  else SS2D_LN_FWD(16);
#undef SS2D_LN_FWD
  return cudaGetLastError();
}

cudaError_t layernorm_bwd_launch(const void* x, const float* w, const void* dy, const float* mean_rstd, void* dx,
                                 float* dw_part, float* db_part, int n_partials, int64_t rows, int C, int dt,
                                 cudaStream_t stream) {
  const int nv = (C + 31) / 32;
#define SS2D_LN_BWD(NV) \
  layernorm_bwd_kernel<NV><<<n_partials, kLnThreads, 0, stream>>>(x, w, dy, mean_rstd, dx, dw_part, db_part, rows, C, dt)
  if (nv <= 2) SS2D_LN_BWD(2);
  else if (nv <= 4) SS2D_LN_BWD(4);
  else if (nv <= 8) SS2D_LN_BWD(8);
  else if (nv <= 12) SS2D_LN_BWD(12);
  else SS2D_LN_BWD(16);
#undef SS2D_LN_BWD
  return cudaGetLastError();
}

}  // namespace ss2d
