// Internal kernel parameter block (host fills it from ss2d_scan_desc; passed by value to the kernels).
#pragma once
#include <stdint.h>

#include "ss2d_b200.h"

namespace ss2d {

struct ScanParams {
  int batch, dim, L, N, G, dpg;      // dpg = dim / G channels per group
  int io_dtype, out_dtype, softplus;
  int layout, H, W;
  int dirs[SS2D_MAX_GROUP_DIRS];     // 0 in SCAN layout
  int u_mod;                         // 0 or the channel modulo for u (and for dout in the backward)
  int last_il;                       // last_state interleaved (h in odd slots)
  int tma_ok;                        // fp32 operands, 16-byte aligned rows: TMA bulk staging allowed
  int nck;                           // number of SS2D_CHUNK checkpoints per row
  int NP;                            // padded state count of the kernel variant (ckpt row length)
  int A_ld;                          // row stride of A / dA / last_state (total dstate); N is this pass's count
  int accum;                         // 1: out/du/ddelta += (later state passes when dstate > 32), D-skip already applied
  int64_t u_bs, u_ds, dl_bs, dl_ds, out_bs, out_ds;
  int64_t B_bs, B_gs, B_ns, C_bs, C_gs, C_ns;
  const void* u;
  const void* delta;
  const float* A;
  const void* Bm;
  const void* Cm;
  const float* Dv;
  const float* bias;
  void* out;
  float* ckpt;          // [batch][dim][nck][NP]
  float* last_state;    // [batch][dim][N] or null
  // backward only
  const void* dout;
  const float* ckpt_in;
  void* du;
  void* ddelta;
  float* dB;            // fp32 [batch][G][N][L], pre-zeroed, accumulated with red.global when a group spans several CTAs
  float* dC;
  float* part;          // workspace: [batch][dim][N + 2] per-batch partial sums of dA, dD, dbias
};

// Opaque 128-byte TMA tensor-map descriptors (CUtensorMap), filled on the host by cuTensorMapEncodeTiled and passed
// to the kernels as a __grid_constant__ parameter.
struct alignas(64) TMap { unsigned long long v[16]; };
struct TmaMaps { TMap u, dl, B, C, dy; };

// Kernel variant for a state count: NS states per thread, R lanes per row. NS * R >= N.
struct Variant { int NS, R; };
inline Variant pick_variant(int N) {
  if (N <= 1) return {1, 1};
  if (N <= 2) return {2, 1};
  if (N <= 4) return {4, 1};
  if (N <= 8) return {4, 2};
  if (N <= 16) return {4, 4};
  return {4, 8};      // N <= 32; larger N is processed in passes of 32 states by the host
}

}  // namespace ss2d
