// Pipe-throughput probe for B200 (sm_100a): MUFU.EX2, FFMA, FFMA2 (fma.rn.f32x2), SHFL, LDS.128.
// Prints lane-ops per clock per SM for each, used in DESIGN.md to bound the N=16 scan (not HBM-bound).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipes pipes.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

#define ITERS 4096

__global__ void k_mufu(float* out, float seed) {
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = seed * (threadIdx.x + j) * 1e-3f;
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[j]));
  }
  float s = 0; for (int j = 0; j < 8; ++j) s += v[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma(float* out, float seed) {
  float v[8]; float a = seed, b = seed * 0.5f;
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = seed * (threadIdx.x + j);
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[j]) : "f"(a), "f"(b));
  }
  float s = 0; for (int j = 0; j < 8; ++j) s += v[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float* out, float seed) {
  unsigned long long v[8]; unsigned long long a, b;
  float2 fa = make_float2(seed, seed * 0.9f), fb = make_float2(seed * 0.5f, seed * 0.4f);
  a = *reinterpret_cast<unsigned long long*>(&fa); b = *reinterpret_cast<unsigned long long*>(&fb);
#pragma unroll
  for (int j = 0; j < 8; ++j) { float2 f = make_float2(seed * (threadIdx.x + j), seed * j); v[j] = *reinterpret_cast<unsigned long long*>(&f); }
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[j]) : "l"(a), "l"(b));
  }
  float s = 0; for (int j = 0; j < 8; ++j) { float2 f = *reinterpret_cast<float2*>(&v[j]); s += f.x + f.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mixed: 1 MUFU : 2 FFMA2 per slot, to see whether the pipes overlap
__global__ void k_mix(float* out, float seed) {
  unsigned long long v[8]; unsigned long long a, b; float m[4];
  float2 fa = make_float2(seed, seed * 0.9f), fb = make_float2(seed * 0.5f, seed * 0.4f);
  a = *reinterpret_cast<unsigned long long*>(&fa); b = *reinterpret_cast<unsigned long long*>(&fb);
#pragma unroll
  for (int j = 0; j < 8; ++j) { float2 f = make_float2(seed * (threadIdx.x + j), seed * j); v[j] = *reinterpret_cast<unsigned long long*>(&f); }
#pragma unroll
  for (int j = 0; j < 4; ++j) m[j] = seed * (threadIdx.x + j) * 1e-3f;
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[j]));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[2 * j]) : "l"(a), "l"(b));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[2 * j + 1]) : "l"(a), "l"(b));
    }
  }
  float s = 0; for (int j = 0; j < 8; ++j) { float2 f = *reinterpret_cast<float2*>(&v[j]); s += f.x + f.y; }
  for (int j = 0; j < 4; ++j) s += m[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_shfl(float* out, float seed) {
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = seed * (threadIdx.x + j);
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __shfl_xor_sync(0xffffffffu, v[j], 1 + (j & 3));
  }
  float s = 0; for (int j = 0; j < 8; ++j) s += v[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_lds128(float* out, float seed) {
  __shared__ float4 buf[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = make_float4(seed, seed, seed, i);
  __syncthreads();
  float4 acc = make_float4(0, 0, 0, 0);
  int idx = threadIdx.x;
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 t = buf[(idx + j * 32) & 1023];
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    idx = (idx + 7) & 1023;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

template <typename K>
static void run(const char* name, K kern, double lane_ops_per_thread, int threads, int ctas_per_sm) {
  int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
  int sms = p.multiProcessorCount;
  int grid = sms * ctas_per_sm;
  float* out; cudaMalloc(&out, sizeof(float) * grid * threads);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 3; ++w) kern<<<grid, threads>>>(out, 1.0001f);
  cudaEventRecord(e0);
  const int reps = 5;
  for (int r = 0; r < reps; ++r) kern<<<grid, threads>>>(out, 1.0001f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
  int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev);
  double ops = lane_ops_per_thread * (double)grid * threads;
  double per_s = ops / (ms * 1e-3);
  printf("{\"probe\":\"%s\",\"threads\":%d,\"ctas_per_sm\":%d,\"ms\":%.4f,\"lane_ops_per_s\":%.4g,"
         "\"lane_ops_per_clk_per_sm_at_max_clk\":%.2f,\"sms\":%d,\"max_clk_mhz\":%d}\n",
         name, threads, ctas_per_sm, ms, per_s, per_s / sms / (clk_khz * 1e3), sms, clk_khz / 1000);
  cudaFree(out);
}

int main() {
  for (int c = 1; c <= 4; c *= 2) {
    run("mufu_ex2", k_mufu, 8.0 * ITERS, 256, c);
    run("ffma", k_ffma, 8.0 * ITERS, 256, c);
    run("ffma2(packed, counts 2 fma per lane-op)", k_ffma2, 16.0 * ITERS, 256, c);
    run("mix(1 ex2 : 2 ffma2) ex2 count", k_mix, 4.0 * ITERS, 256, c);
    run("shfl", k_shfl, 8.0 * ITERS, 256, c);
    run("lds128(bytes/16)", k_lds128, 8.0 * ITERS, 256, c);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
