// Host-side helpers shared by the launchers: per-device one-time kernel attributes and device properties.
// cudaFuncSetAttribute and the SM count are PER DEVICE; a process may drive several GPUs (and the autograd engine calls the
// backward launchers from its own per-device threads), so nothing here is keyed on the process alone.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>

namespace ss2d {

struct PerDeviceOnce { std::atomic<uint64_t> mask{0}; };     // bit d: attribute set on device d (idempotent, so races are benign)

inline cudaError_t func_attr_once(PerDeviceOnce& once, const void* fn, cudaFuncAttribute attr, int value) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const uint64_t bit = 1ull << (dev & 63);
  if (once.mask.load(std::memory_order_acquire) & bit) return cudaSuccess;
  e = cudaFuncSetAttribute(fn, attr, value);
  if (e == cudaSuccess) once.mask.fetch_or(bit, std::memory_order_release);
  return e;
}

inline int sm_count_current_device() {
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  int n = cache[dev & 63].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev & 63].store(n, std::memory_order_relaxed);
  }
  return n;
}

}  // namespace ss2d
