// The depthwise stack of the GroupMamba FFNs on channels-last tensors:
//   PVT2FFN    (/root/reference/gm-unet/model/gm/groupmamba.py:54-83):  fc1 -> DWConv 3x3 (+bias) -> GELU -> fc2
//   custom_ffn (/root/reference/gm-unet/model/gm/custom_mlp.py:338-368): fc1 -> DWConv 3x3 -> GELU ->
//              InceptionDWConv2d_MultiScale (:313-336: x + cat(x_id, dw3x3(x_3), dw5x5(x_5), dw7x7(x_7))) -> fc2
// The reference transposes the (B, L, C) token tensor to NCHW for every convolution and back (DWConv.forward,
// groupmamba.py:446-455; custom_mlp.py:323-336), runs cuDNN / ATen depthwise kernels, a separate GELU, a split, three
// convolutions, a cat and an add. Here the token tensor IS the (B, H, W, C) channels-last image: one stencil kernel with
// a per-channel-segment kernel size and a fused epilogue covers every forward and data-gradient pass
//   epi 0: y = acc         1: y = gelu(acc)         2: y = x + acc (residual)         3: y = aux * gelu'(acc)
// (acc = bias + sum over taps, or acc = x for an identity segment, k = 1; flip = 1 correlates with the flipped kernel = transposed convolution), and one
// reduction kernel gives the weight / bias gradients of a segment. No transposed copy, no split / cat, GELU and the
// residual never touch HBM on their own. HBM-bound: 2 (epi 0-2) or 3 (epi 3) tensor passes of s bytes per element.
// Math in fp32 whatever the I/O type (fp32 / fp16 / bf16); exact-erf GELU as nn.GELU().
#include "common.cuh"
#include "host_util.h"

namespace ss2d {

constexpr int kDwnMaxSeg = 4;
struct DwnSegs {
  int nseg;
  int cbeg[kDwnMaxSeg + 1];        // channels [cbeg[s], cbeg[s + 1]) form segment s
  int k[kDwnMaxSeg];               // 0 (acc = 0), 1 (identity: acc = x), 3, 5 or 7
  const float* w[kDwnMaxSeg];      // (channels of the segment, k, k) fp32, as nn.Conv2d(groups = channels).weight
  const float* b[kDwnMaxSeg];      // (channels of the segment) fp32 or null
};

template <int VEC>
__device__ __forceinline__ void ldv(const void* base, int64_t idx, int dt, float* v) {
  if (VEC == 1) { v[0] = load1(base, idx, dt); return; }
  if (dt == SS2D_F32) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(reinterpret_cast<const float*>(base) + idx));
    v[0] = t.x; v[1] = t.y;
  } else {
    const uint32_t raw = __ldg(reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint16_t*>(base) + idx));
    const float2 t = dt == SS2D_F16 ? __half22float2(*reinterpret_cast<const __half2*>(&raw))
                                    : __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw));
    v[0] = t.x; v[1] = t.y;
  }
}
template <int VEC>
__device__ __forceinline__ void stv(void* base, int64_t idx, int dt, const float* v) {
  if (VEC == 1) { store1(base, idx, dt, v[0]); return; }
  if (dt == SS2D_F32) {
    *reinterpret_cast<float2*>(reinterpret_cast<float*>(base) + idx) = make_float2(v[0], v[1]);
  } else if (dt == SS2D_F16) {
    *reinterpret_cast<__half2*>(reinterpret_cast<__half*>(base) + idx) = __floats2half2_rn(v[0], v[1]);
  } else {
    *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(base) + idx) = __floats2bfloat162_rn(v[0], v[1]);
  }
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  return 0.5f * (1.f + erff(x * 0.70710678f)) + x * 0.3989422804f * __expf(-0.5f * x * x);
}

constexpr int kDwnPX = 4;      // consecutive pixels of an image row per thread: the K x K weights and the overlapping columns stay in registers

// acc[px][v] += sum over taps for pixels w0 .. w0 + PX - 1 of image row (b, h), channels c .. c + VEC - 1
template <int VEC, int K>
__device__ __forceinline__ void dwn_taps(const void* __restrict__ x, const float* __restrict__ wp, int flip, int b, int h, int w0,
                                         int c, int H, int W, int C, int dt, float (*acc)[VEC]) {
  constexpr int P = K / 2;
  float wr[VEC][K * K];
#pragma unroll
  for (int v = 0; v < VEC; ++v)
#pragma unroll
    for (int t = 0; t < K * K; ++t) wr[v][t] = __ldg(wp + v * K * K + (flip ? K * K - 1 - t : t));
#pragma unroll
  for (int i = 0; i < K; ++i) {
    const int hh = h + i - P;
    if (hh < 0 || hh >= H) continue;
    const int64_t rowoff = ((int64_t)(b * H + hh) * W) * C + c;
#pragma unroll
    for (int jj = 0; jj < kDwnPX + K - 1; ++jj) {
      const int ww = w0 + jj - P;
      float xv[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) xv[v] = 0.f;
      if (ww >= 0 && ww < W) ldv<VEC>(x, rowoff + (int64_t)ww * C, dt, xv);
#pragma unroll
      for (int px = 0; px < kDwnPX; ++px) {
        const int j = jj - px;
        if (j >= 0 && j < K) {
#pragma unroll
          for (int v = 0; v < VEC; ++v) acc[px][v] = fmaf(wr[v][i * K + j], xv[v], acc[px][v]);
        }
      }
    }
  }
}

// grid: x = (pixel quads of a row) x (channel vectors), channel vectors fastest (coalesced); y = image row (b, h)
template <int VEC>
__global__ void __launch_bounds__(256)
dwnhwc_stencil_kernel(const void* __restrict__ x, const void* __restrict__ aux, void* __restrict__ y, const DwnSegs sg, int flip,
                      int epi, int H, int W, int C, int dt) {
  const unsigned CV = C / VEC;
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned wq = t / CV, cv = t - wq * CV;
  const int w0 = (int)wq * kDwnPX;
  if (w0 >= W) return;
  const int h = blockIdx.y % H, b = blockIdx.y / H;
  const int c = (int)cv * VEC;
  // segment of this thread's channels (selects, not a dynamic index into the parameter block)
  int k = sg.k[0], cb = 0;
  const float* ws = sg.w[0];
  const float* bs = sg.b[0];
#pragma unroll
  for (int i = 1; i < kDwnMaxSeg; ++i)
    if (i < sg.nseg && c >= sg.cbeg[i]) { k = sg.k[i]; cb = sg.cbeg[i]; ws = sg.w[i]; bs = sg.b[i]; }
  const int cl = c - cb;
  float acc[kDwnPX][VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    const float bv = (k > 1 && bs) ? __ldg(bs + cl + v) : 0.f;
#pragma unroll
    for (int px = 0; px < kDwnPX; ++px) acc[px][v] = bv;
  }
  const float* wp = ws + (int64_t)cl * k * k;
  if (k == 3) dwn_taps<VEC, 3>(x, wp, flip, b, h, w0, c, H, W, C, dt, acc);
  else if (k == 5) dwn_taps<VEC, 5>(x, wp, flip, b, h, w0, c, H, W, C, dt, acc);
  else if (k == 7) dwn_taps<VEC, 7>(x, wp, flip, b, h, w0, c, H, W, C, dt, acc);
  const int64_t o0 = ((int64_t)blockIdx.y * W + w0) * C + c;
#pragma unroll
  for (int px = 0; px < kDwnPX; ++px) {
    if (w0 + px >= W) break;
    const int64_t o = o0 + (int64_t)px * C;
    float out[VEC];
    if (k == 1) {                  // identity segment: acc = x
      ldv<VEC>(x, o, dt, out);
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[px][v] = out[v];
    }
    if (epi == 0) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) out[v] = acc[px][v];
    } else if (epi == 1) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) out[v] = gelu_erf(acc[px][v]);
    } else if (epi == 2) {
      float xc[VEC];
      ldv<VEC>(x, o, dt, xc);
#pragma unroll
      for (int v = 0; v < VEC; ++v) out[v] = xc[v] + acc[px][v];
    } else {
      float av[VEC];
      ldv<VEC>(aux, o, dt, av);
#pragma unroll
      for (int v = 0; v < VEC; ++v) out[v] = av[v] * gelu_erf_grad(acc[px][v]);
    }
    stv<VEC>(y, o, dt, out);
  }
}

cudaError_t dwnhwc_stencil_launch(const void* x, const void* aux, void* y, const DwnSegs& sg, int flip, int epi, int B, int H,
                                  int W, int C, int dt, cudaStream_t stream) {
  bool even = (C & 1) == 0;
  for (int s = 0; s <= sg.nseg; ++s) even = even && (sg.cbeg[s] & 1) == 0;
  const int vec = even ? 2 : 1;
  const int64_t per_row = (int64_t)((W + kDwnPX - 1) / kDwnPX) * (C / vec);
  const int64_t rows = (int64_t)B * H;
  if (per_row >= (1ll << 31) || rows > 65535 * 32768ll) return cudaErrorInvalidValue;
  // image rows on grid.y (<= 65535): very tall batches are walked in slices of rows
  for (int64_t r0 = 0; r0 < rows; r0 += 65535 - 65535 % H) {
    const int64_t nr = rows - r0 < 65535 - 65535 % H ? rows - r0 : 65535 - 65535 % H;
    dim3 grid((unsigned)((per_row + 255) / 256), (unsigned)nr);
    const size_t eb = (size_t)(dt == SS2D_F32 ? 4 : 2) * (size_t)r0 * W * C;
    const char* xs = static_cast<const char*>(x) + eb;
    const char* as = aux ? static_cast<const char*>(aux) + eb : nullptr;
    char* ys = static_cast<char*>(y) + eb;
    if (vec == 2) dwnhwc_stencil_kernel<2><<<grid, 256, 0, stream>>>(xs, as, ys, sg, flip, epi, H, W, C, dt);
    else dwnhwc_stencil_kernel<1><<<grid, 256, 0, stream>>>(xs, as, ys, sg, flip, epi, H, W, C, dt);
  }
  return cudaGetLastError();
}

// ---- weight / bias gradient of one segment: dW[c][i][j] = sum_(b,h,w) g[b,h,w,c] x[b,h+i-P,w+j-P,c], db[c] = sum g ----
// CTA = 32 consecutive channels x 8 pixel lanes over one slab of pixels; K*K + 1 sums per thread in registers, folded
// over the pixel lanes through shared memory; per-slab partials, fixed-order finalize (deterministic).
template <int K>
__global__ void __launch_bounds__(256)
dwnhwc_wgrad_kernel(const void* __restrict__ x, const void* __restrict__ g, float* __restrict__ part, int c0, int nc, int B, int H,
                    int W, int C, int dt, int slabs) {
  constexpr int T = K * K + 1, P = K / 2;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int cl = blockIdx.x * 32 + tx;
  const bool cok = cl < nc;
  const int c = c0 + (cok ? cl : 0);
  const int64_t total = (int64_t)B * H * W;
  const int64_t per = (total + slabs - 1) / slabs;
  const int64_t p0 = (int64_t)blockIdx.y * per, p1 = p0 + per < total ? p0 + per : total;
  float acc[T];
#pragma unroll
  for (int t = 0; t < T; ++t) acc[t] = 0.f;
  for (int64_t p = p0 + ty; p < p1; p += 8) {
    const unsigned pu = (unsigned)p;                         // B * H * W < 2^31 (checked by the host)
    const unsigned pr = pu / (unsigned)W;
    const int w = (int)(pu - pr * (unsigned)W), h = (int)(pr % (unsigned)H);
    const float gv = load1(g, p * C + c, dt);
    acc[K * K] += gv;
#pragma unroll
    for (int i = 0; i < K; ++i) {
      const int hh = h + i - P;
      if (hh < 0 || hh >= H) continue;
#pragma unroll
      for (int j = 0; j < K; ++j) {
        const int ww = w + j - P;
        if (ww < 0 || ww >= W) continue;
        acc[i * K + j] = fmaf(gv, load1(x, (p + (int64_t)(i - P) * W + (j - P)) * C + c, dt), acc[i * K + j]);
      }
    }
  }
  __shared__ float s_red[8][33];
  float* dst = part + ((int64_t)blockIdx.y * nc + cl) * T;
#pragma unroll
  for (int t = 0; t < T; ++t) {
    s_red[ty][tx] = acc[t];
    __syncthreads();
    if (ty == 0 && cok) {
      float v = 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) v += s_red[r][tx];
      dst[t] = v;
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

int dwnhwc_wgrad_slabs(int B, int H, int W, int nc) {
  const int64_t total = (int64_t)B * H * W;
  const int cblocks = (nc + 31) / 32;
  int64_t slabs = ((int64_t)sm_count_current_device() * 8 + cblocks - 1) / cblocks;
  const int64_t cap = total / 64 > 1 ? total / 64 : 1;
  if (slabs > cap) slabs = cap;
  if (slabs > 65535) slabs = 65535;
  return (int)(slabs < 1 ? 1 : slabs);
}

cudaError_t dwnhwc_wgrad_launch(const void* x, const void* g, int c0, int nc, int K, float* dW, float* db, int B, int H, int W,
                                int C, int dt, float* part, cudaStream_t stream) {
  const int slabs = dwnhwc_wgrad_slabs(B, H, W, nc);
  dim3 grid((nc + 31) / 32, slabs);
  if (K == 3) dwnhwc_wgrad_kernel<3><<<grid, 256, 0, stream>>>(x, g, part, c0, nc, B, H, W, C, dt, slabs);
  else if (K == 5) dwnhwc_wgrad_kernel<5><<<grid, 256, 0, stream>>>(x, g, part, c0, nc, B, H, W, C, dt, slabs);
  else dwnhwc_wgrad_kernel<7><<<grid, 256, 0, stream>>>(x, g, part, c0, nc, B, H, W, C, dt, slabs);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const int T = K * K + 1;
  dwnhwc_wgrad_finalize_kernel<<<(nc * T + 255) / 256, 256, 0, stream>>>(part, dW, db, nc, T, slabs);
  return cudaGetLastError();
}

}  // namespace ss2d
