// Shared-memory load probe for B200 (sm_100a): cost of LDS.32/64/128 when the lanes of a warp read
// 1, 2, 4, 8 or 32 distinct addresses (broadcast behaviour). Prints warp-instructions per clock per SM.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048

template <int VEC, int DISTINCT>
__global__ void k_lds(float* out, int seed) {
  __shared__ __align__(16) float buf[8192];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) buf[i] = (float)(i & 15) * 1e-3f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  // lanes are grouped so that (32/DISTINCT) consecutive lanes share an address
  const int grp = lane / (32 / DISTINCT);
  int base = (grp * (VEC + 4 * (VEC == 1 ? 0 : 0))) ;          // consecutive VEC-wide words: conflict-free
  base = grp * VEC + seed;                                       // seed==0 at run time; defeats constant folding
  float acc = 0.f;
  int off = 0;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int idx = (base + off + j * 256) & 8191 & ~(VEC - 1);
      if (VEC == 1) { acc += buf[idx]; }
      else if (VEC == 2) { float2 t = *reinterpret_cast<const float2*>(&buf[idx]); acc += t.x + t.y; }
      else { float4 t = *reinterpret_cast<const float4*>(&buf[idx]); acc += t.x + t.y + t.z + t.w; }
    }
    off = (off + 2048 + 4 * VEC) & 8191;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int VEC, int DISTINCT>
static void run() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int threads = 256, grid = p.multiProcessorCount * 4;
  float* out; cudaMalloc(&out, sizeof(float) * grid * threads);
  for (int w = 0; w < 2; ++w) k_lds<VEC, DISTINCT><<<grid, threads>>>(out, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) k_lds<VEC, DISTINCT><<<grid, threads>>>(out, 0);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  double winstr = 8.0 * ITERS * (double)grid * threads / 32.0;
  printf("{\"probe\":\"lds\",\"vec_words\":%d,\"distinct_addr_per_warp\":%d,\"ms\":%.4f,"
         "\"warp_instr_per_clk_per_sm_at_max_clk\":%.3f,\"cycles_per_warp_instr\":%.2f}\n",
         VEC, DISTINCT, ms, winstr / (ms * 1e-3) / p.multiProcessorCount / (clk_khz * 1e3),
         (ms * 1e-3) * p.multiProcessorCount * (clk_khz * 1e3) / winstr);
  cudaFree(out);
}

int main() {
  run<1, 1>(); run<1, 32>();
  run<2, 1>(); run<2, 2>(); run<2, 32>();
  run<4, 1>(); run<4, 2>(); run<4, 4>(); run<4, 8>(); run<4, 32>();
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
