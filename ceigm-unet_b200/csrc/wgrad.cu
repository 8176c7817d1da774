// Weight gradient of the small projections around the scan (in_proj / out_proj / x_proj / dt_proj of SS2D,
// /root/reference/gm-unet/model/gm/ss2d.py:294-335, 465-477, 504, 518; GroupMambaLayer.proj, groupmamba.py:157):
//     dW[m][n] = sum over (b, r) of dY(b, r, m) * X(b, r, n)
// with M x N tiny (32 x 16 ... 64 x 64) and the reduction length B * L huge (75 264 at stage 1 of a 224^2 batch-24
// step). cuBLAS runs that "tall-skinny" shape as ONE 64 x 64 tile walking the whole reduction (64 - 83 us per call,
// 27 % of a GroupMambaLayer's GPU time); here it is what it is — an HBM-bound reduction: up to 296 CTAs each stream
// a slab of rows through shared memory (4 x 4 register tiles, the 256 threads folded into row groups when M x N is
// small), write an M x N partial, and a second kernel adds the partials in a fixed order (deterministic).
// Operands are addressed by (batch, row, column) element strides, so both the channels-last rows of a Linear and the
// channel-major (B, C, L) operands of the 1 x 1 projections are read in place, coalesced along whichever index is
// contiguous. fp32 / fp16 / bf16 operands, fp32 accumulation and result.
#include "common.cuh"
#include "host_util.h"

namespace ss2d {

constexpr int kWgThreads = 256;
constexpr int kWgRows = 64;           // rows per shared-memory tile
constexpr int kWgMaxTiles = 256;      // (M/4) * (N/4) register tiles must fit one block
constexpr int kWgMaxCtas = 296;

__device__ __forceinline__ int round4i(int x) { return (x + 3) & ~3; }

// Stages rows [g0, g0 + kWgRows) of one operand into dst[r * pitch + c]. Three addressing modes:
//   flat  — fp32, dense rows (row stride == C, C % 4 == 0, one batch, 16-byte aligned): the tile is one contiguous
//           chunk; 128-bit loads, no index arithmetic beyond one division per vector;
//   rows  — column stride 1: threads run along the columns of a row;
//   cols  — row stride 1 (channel-major tensors): threads run along the rows of a column.
__device__ __forceinline__ void stage_tile(float* __restrict__ dst, int pitch, const void* __restrict__ src, int dt, int C,
                                           int CP, int64_t cs, bool flat, const int64_t* __restrict__ rowoff, int64_t g0,
                                           int64_t g_end, int tid) {
  if (flat) {
    const float* base = reinterpret_cast<const float*>(src) + g0 * C;
    const int nvec = kWgRows * C / 4, c4n = C / 4;
    const int64_t valid = (g_end - g0) * C;          // elements of the tile inside this CTA's slab
    for (int v = tid; v < nvec; v += kWgThreads) {
      const int r = v / c4n, c = (v - r * c4n) * 4;
      const float4 x = (int64_t)v * 4 < valid ? __ldg(reinterpret_cast<const float4*>(base) + v) : make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(dst + r * pitch + c) = x;
    }
    return;
  }
  for (int i = tid; i < kWgRows * CP; i += kWgThreads) {
    int r, c;
    if (cs == 1) { r = i / CP; c = i - r * CP; } else { c = i / kWgRows; r = i - c * kWgRows; }
    const int64_t ro = rowoff[r];
    dst[r * pitch + c] = (ro >= 0 && c < C) ? load1(src, ro + (int64_t)c * cs, dt) : 0.f;
  }
}

// Asynchronous variant for fp32 operands: cp.async (LDGSTS) straight into the other shared-memory buffer while the current
// tile is being multiplied; out-of-range elements are zero-filled through the src-size operand.
__device__ __forceinline__ void cp_async_zfill(void* smem_dst, const void* gmem_src, int bytes, bool valid) {
  const int src = valid ? bytes : 0;
  if (bytes == 16)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(src) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(src) : "memory");
}
__device__ __forceinline__ void stage_tile_async(float* __restrict__ dst, int pitch, const float* __restrict__ src, int C, int CP,
                                                 int64_t cs, bool flat, const int64_t* __restrict__ rowoff, int64_t g0,
                                                 int64_t g_end, int tid) {
  if (flat) {
    const float* base = src + g0 * C;
    const int nvec = kWgRows * C / 4, c4n = C / 4;
    const int64_t valid = (g_end - g0) * C;
    for (int v = tid; v < nvec; v += kWgThreads) {
      const int r = v / c4n, c = (v - r * c4n) * 4;
      const bool ok = (int64_t)v * 4 < valid;
      cp_async_zfill(dst + r * pitch + c, ok ? base + (int64_t)v * 4 : src, 16, ok);
    }
    return;
  }
  for (int i = tid; i < kWgRows * CP; i += kWgThreads) {
    int r, c;
    if (cs == 1) { r = i / CP; c = i - r * CP; } else { c = i / kWgRows; r = i - c * kWgRows; }
    const int64_t ro = rowoff[r];
    const bool ok = ro >= 0 && c < C;
    cp_async_zfill(dst + r * pitch + c, ok ? src + ro + (int64_t)c * cs : src, 4, ok);
  }
}

__global__ void __launch_bounds__(kWgThreads)
wgrad_ts_kernel(const void* __restrict__ dY, const void* __restrict__ X, float* __restrict__ part, int64_t total_rows,
                int rows, int M, int N, int64_t y_bs, int64_t y_rs, int64_t y_cs, int64_t x_bs, int64_t x_rs, int64_t x_cs,
                int y_dt, int x_dt, int64_t rows_per_cta) {
  extern __shared__ __align__(16) float s_wg[];
  const int MP = round4i(M), NP = round4i(N);
  const int MQ = MP + 4, NQ = NP + 4;       // row pitches: + 4 floats spreads the transposed staging stores over banks
  float* sY = s_wg;                         // [kWgRows][MQ]
  float* sX = sY + kWgRows * MQ;            // [kWgRows][NQ]
  __shared__ int64_t s_rowY[2][kWgRows], s_rowX[2][kWgRows];
  const int tid = threadIdx.x;
  const int tn_cnt = NP / 4, tpg = (MP / 4) * tn_cnt;          // threads per row group
  const int RG = kWgThreads / tpg;                              // row groups (>= 1)
  const int rg = tid / tpg, tin = tid - rg * tpg;
  const int tm = tin / tn_cnt, tn = tin - tm * tn_cnt;
  const bool worker = rg < RG;
  // dense fp32 rows of a single batch: the tile is one contiguous, 16-byte aligned chunk
  const bool y_flat = y_dt == SS2D_F32 && y_cs == 1 && y_rs == M && (M & 3) == 0 && total_rows == rows &&
                      (reinterpret_cast<uintptr_t>(dY) & 15) == 0;
  const bool x_flat = x_dt == SS2D_F32 && x_cs == 1 && x_rs == N && (N & 3) == 0 && total_rows == rows &&
                      (reinterpret_cast<uintptr_t>(X) & 15) == 0;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int64_t g_begin = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t g_end = g_begin + rows_per_cta < total_rows ? g_begin + rows_per_cta : total_rows;
  auto multiply = [&](const float* tY, const float* tX) {
    if (worker) {
      for (int r = rg; r < kWgRows; r += RG) {
        const float4 y4 = *reinterpret_cast<const float4*>(tY + r * MQ + tm * 4);
        const float4 x4 = *reinterpret_cast<const float4*>(tX + r * NQ + tn * 4);
        const float yv[4] = {y4.x, y4.y, y4.z, y4.w}, xv[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(yv[i], xv[j], acc[i][j]);
      }
    }
  };
  auto row_offsets = [&](int buf, int64_t g0) {        // threads 0 .. kWgRows-1
    const int64_t g = g0 + tid;
    if (g < g_end) {
      const int64_t b = g / rows, r = g - b * rows;
      s_rowY[buf][tid] = b * y_bs + r * y_rs;
      s_rowX[buf][tid] = b * x_bs + r * x_rs;
    } else {
      s_rowY[buf][tid] = -1; s_rowX[buf][tid] = -1;
    }
  };
  if (y_dt == SS2D_F32 && x_dt == SS2D_F32) {
    // double-buffered: tile t + 1 streams into the other buffer (cp.async) while tile t is multiplied; one barrier per tile
    const int tile_floats = kWgRows * (MQ + NQ);
    const float* fY = reinterpret_cast<const float*>(dY);
    const float* fX = reinterpret_cast<const float*>(X);
    if (tid < kWgRows) row_offsets(0, g_begin);
    __syncthreads();
    if (g_begin < g_end) {
      stage_tile_async(sY, MQ, fY, M, MP, y_cs, y_flat, s_rowY[0], g_begin, g_end, tid);
      stage_tile_async(sX, NQ, fX, N, NP, x_cs, x_flat, s_rowX[0], g_begin, g_end, tid);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    int cur = 0;
    for (int64_t g0 = g_begin; g0 < g_end; g0 += kWgRows, cur ^= 1) {
      const bool has_next = g0 + kWgRows < g_end;
      if (has_next && tid < kWgRows) row_offsets(cur ^ 1, g0 + kWgRows);
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();          // tile g0 has landed; everybody is done multiplying the tile that used the other buffer
      if (has_next) {
        float* nY = s_wg + (cur ^ 1) * tile_floats;
        stage_tile_async(nY, MQ, fY, M, MP, y_cs, y_flat, s_rowY[cur ^ 1], g0 + kWgRows, g_end, tid);
        stage_tile_async(nY + kWgRows * MQ, NQ, fX, N, NP, x_cs, x_flat, s_rowX[cur ^ 1], g0 + kWgRows, g_end, tid);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      const float* tY = s_wg + cur * tile_floats;
      multiply(tY, tY + kWgRows * MQ);
    }
  } else {
    for (int64_t g0 = g_begin; g0 < g_end; g0 += kWgRows) {
      __syncthreads();
      if (tid < kWgRows) row_offsets(0, g0);
      __syncthreads();
      stage_tile(sY, MQ, dY, y_dt, M, MP, y_cs, y_flat, s_rowY[0], g0, g_end, tid);
      stage_tile(sX, NQ, X, x_dt, N, NP, x_cs, x_flat, s_rowX[0], g0, g_end, tid);
      __syncthreads();
      multiply(sY, sX);
    }
  }
  // fold the row groups (fixed order), then write this CTA's M x N partial
  __syncthreads();
  float* s_red = s_wg;                      // [RG][tpg][16] <= 256 * 16 floats: the launcher sizes shared memory for it
  if (worker) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s_red[(rg * tpg + tin) * 16 + i * 4 + j] = acc[i][j];
  }
  __syncthreads();
  for (int o = tid; o < tpg * 16; o += kWgThreads) {
    float s = 0.f;
    for (int k = 0; k < RG; ++k) s += s_red[k * tpg * 16 + o];
    const int t = o >> 4, e = o & 15;
    const int m = (t / tn_cnt) * 4 + (e >> 2), n = (t % tn_cnt) * 4 + (e & 3);
    if (m < M && n < N) part[((int64_t)blockIdx.x * M + m) * N + n] = s;
  }
}

// dW[i] = sum over the CTAs' partials, one warp per output, fixed order (lane-strided partial sums, then a butterfly)
__global__ void wgrad_ts_finalize_kernel(const float* __restrict__ part, float* __restrict__ dW, int n_part, int MN) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= MN) return;
  float s = 0.f;
  for (int c = lane; c < n_part; c += 32) s += part[(int64_t)c * MN + i];
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) dW[i] = s;
}

static int wgrad_ctas(int64_t total_rows) {
  int64_t c = (total_rows + kWgRows - 1) / kWgRows;          // at least one tile of rows per CTA
  if (c > kWgMaxCtas) c = kWgMaxCtas;
  return c < 1 ? 1 : (int)c;
}

bool wgrad_ts_supported(int M, int N) {
  return M > 0 && N > 0 && M <= 256 && N <= 256 && (((M + 3) / 4) * ((N + 3) / 4)) <= kWgMaxTiles;
}

size_t wgrad_ts_workspace_floats(int64_t total_rows, int M, int N) { return (size_t)wgrad_ctas(total_rows) * M * N; }

cudaError_t wgrad_ts_launch(const void* dY, const void* X, float* dW, int batch, int rows, int M, int N, int64_t y_bs,
                            int64_t y_rs, int64_t y_cs, int64_t x_bs, int64_t x_rs, int64_t x_cs, int y_dt, int x_dt,
                            float* workspace, cudaStream_t stream) {
  const int64_t total = (int64_t)batch * rows;
  const int ctas = wgrad_ctas(total);
  int64_t rpc = (total + ctas - 1) / ctas;
  rpc = (rpc + kWgRows - 1) / kWgRows * kWgRows;
  const int MP = (M + 3) & ~3, NP = (N + 3) & ~3;
  size_t smem = (size_t)2 * kWgRows * (MP + NP + 8) * 4;      // two tile buffers
  const size_t red = (size_t)kWgThreads * 16 * 4;
  if (smem < red) smem = red;
  if (smem > 48 * 1024) {      // one opt-in to the kernel's largest footprint per device (MP, NP <= 256)
    static PerDeviceOnce once;
    const int max_smem = 2 * kWgRows * (272 + 8) * 4;      // MP * NP <= 4096 and MP, NP <= 256 -> MP + NP <= 272
    cudaError_t e = func_attr_once(once, reinterpret_cast<const void*>(wgrad_ts_kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    if (e != cudaSuccess) return e;
  }
  wgrad_ts_kernel<<<ctas, kWgThreads, smem, stream>>>(dY, X, workspace, total, rows, M, N, y_bs, y_rs, y_cs, x_bs, x_rs,
                                                       x_cs, y_dt, x_dt, rpc);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  wgrad_ts_finalize_kernel<<<(M * N * 32 + 255) / 256, 256, 0, stream>>>(workspace, dW, ctas, M * N);
  return cudaGetLastError();
}

}  // namespace ss2d
