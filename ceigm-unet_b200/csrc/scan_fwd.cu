// Selective-scan forward for sm_100a.
//
// Replaces selective_scan_fwd_kernel (/root/reference/gm-unet/kernels/selective_scan/csrc/selective_scan/cus/
// selective_scan_fwd_kernel.cuh:61-172) together with CrossScan*/CrossMerge* (model/gm/csms6s.py:11-206) when the
// NATURAL layout is used. Not a port: the reference runs one CTA per (batch, channel) row and a CUB block scan
// per state; here a CTA owns CH channel rows of one (batch, group), every thread owns NS states of one row
// and walks L sequentially (minimum-op recurrence: 1 ex2 + 4 flops per state and element), the R lanes of a
// row combine their partial y with warp shuffles, and B/C tiles are staged once per CTA and broadcast from
// shared memory to all rows. At d_state = 16 the kernel is bound by MUFU.EX2 (16/clk/SM), not by HBM —
// see DESIGN.md.
#include "scan_params.h"
#include "scan_tile.cuh"

namespace ss2d {

constexpr int kFwdLT = 64;
constexpr int kFwdLTP = kFwdLT + 4;

template <int NS, int R, int RPT>
struct FwdShape {
  static constexpr int RL = 32 / R;          // rows per warp per RPT slot
  static constexpr int RPW = RL * RPT;       // rows per warp
  static constexpr int CH = 4 * RPW;         // rows per CTA
  static constexpr int NP = NS * R;          // padded states
  static constexpr int stage_floats = (2 * CH + 2 * NP) * kFwdLTP;     // delta, u, B, C tiles of one pipeline stage
  static constexpr size_t smem_bytes = (size_t)(2 * stage_floats + CH * kFwdLTP + 2 * CH) * 4 + 16;
};

// Tile pipeline: two shared-memory stages. When the operands are fp32 rows that are contiguous along the scan
// (SCAN layout or direction 1) and 16-byte aligned, warp 0 fetches tile t+1 with TMA bulk copies (UBLKCP, one per
// row, completing on the stage's mbarrier) while all four warps compute tile t. Any other case (16-bit dtypes,
// transposed or reversed traversal, ragged tails) goes through the synchronous, index-mapped stage_rows().
template <int NS, int R, int RPT>
__global__ void __launch_bounds__(kThreads) scan_fwd_kernel(const ScanParams p) {
  using S = FwdShape<NS, R, RPT>;
  constexpr int LT = kFwdLT, LTP = kFwdLTP, CH = S::CH, NP = S::NP, RL = S::RL;
  extern __shared__ __align__(16) float smem[];
  float* s_du = smem + 2 * S::stage_floats;   // delta * u                      [CH][LTP]
  float* s_bias = s_du + CH * LTP;            // [CH]
  float* s_D = s_bias + CH;                   // [CH]
  uint64_t* mbar = reinterpret_cast<uint64_t*>(s_D + CH);   // [2]
  auto st_dl = [&](int s) { return smem + s * S::stage_floats; };              // delta (raw, then activated) [CH][LTP]
  auto st_u = [&](int s) { return smem + s * S::stage_floats + CH * LTP; };    // u                            [CH][LTP]
  auto st_B = [&](int s) { return smem + s * S::stage_floats + 2 * CH * LTP; };            // [NP][LTP]
  auto st_C = [&](int s) { return smem + s * S::stage_floats + (2 * CH + NP) * LTP; };     // [NP][LTP]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q = lane % R, rl = lane / R;
  const int b = blockIdx.z, g = blockIdx.y;
  const int row0 = blockIdx.x * CH;                        // first channel of this CTA inside the group
  const int rows_valid = min(CH, p.dpg - row0);
  const int d0 = g * p.dpg + row0;                         // global channel of row 0
  const int L = p.L;
  ScanOrder so;
  so.dir = p.layout == SS2D_LAYOUT_NATURAL ? p.dirs[g] : 0;
  so.H = p.H; so.W = p.W; so.L = L;
  const bool tma = p.tma_ok && (so.dir == 0 || so.dir == 1);

  for (int r = tid; r < CH; r += kThreads) {
    const bool ok = r < rows_valid;
    s_bias[r] = (ok && p.bias) ? p.bias[d0 + r] : 0.f;
    s_D[r] = (ok && p.Dv && !p.accum) ? p.Dv[d0 + r] : 0.f;
  }
  if (tma) {
    // rows that TMA never writes must not hold garbage that could turn into NaN * 0: padded states and idle rows
    for (int i = tid; i < 2 * S::stage_floats; i += kThreads) smem[i] = 0.f;
    if (tid == 0) { mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1); fence_mbar_init(); }
    fence_proxy_async();
  }

  int rk[RPT];
  float A2[RPT][NS], h[RPT][NS];
#pragma unroll
  for (int k = 0; k < RPT; ++k) {
    rk[k] = warp * S::RPW + k * RL + rl;
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      const int n = j * R + q;
      A2[k][j] = (rk[k] < rows_valid && n < p.N) ? p.A[(int64_t)(d0 + rk[k]) * p.A_ld + n] * kLog2e : 0.f;
      h[k][j] = 0.f;
    }
  }

  const int64_t u_boff = (int64_t)b * p.u_bs, dl_boff = (int64_t)b * p.dl_bs, out_boff = (int64_t)b * p.out_bs;
  auto u_off = [&](int r) { const int d = d0 + r; return u_boff + (int64_t)(p.u_mod > 0 ? d % p.u_mod : d) * p.u_ds; };
  auto dl_off = [&](int r) { return dl_boff + (int64_t)(d0 + r) * p.dl_ds; };
  const int64_t B_base = (int64_t)b * p.B_bs + (int64_t)g * p.B_gs;
  const int64_t C_base = (int64_t)b * p.C_bs + (int64_t)g * p.C_gs;
  auto B_off = [&](int n) { return B_base + (int64_t)n * p.B_ns; };
  auto C_off = [&](int n) { return C_base + (int64_t)n * p.C_ns; };

  // warp 0: one bulk copy per operand row of tile `t` into stage `s`
  auto issue_tile = [&](int t, int s) {
    const int l0 = t * LT;
    const int nrow = 2 * rows_valid + 2 * p.N;
    if (lane == 0) mbar_arrive_expect_tx(&mbar[s], (uint32_t)nrow * LT * 4);
    __syncwarp();
    const float* gu = reinterpret_cast<const float*>(p.u);
    const float* gd = reinterpret_cast<const float*>(p.delta);
    const float* gB = reinterpret_cast<const float*>(p.Bm);
    const float* gC = reinterpret_cast<const float*>(p.Cm);
    for (int i = lane; i < nrow; i += 32) {
      if (i < rows_valid) tma_load_1d(st_u(s) + i * LTP, gu + u_off(i) + l0, LT * 4, &mbar[s]);
      else if (i < 2 * rows_valid) tma_load_1d(st_dl(s) + (i - rows_valid) * LTP, gd + dl_off(i - rows_valid) + l0, LT * 4, &mbar[s]);
      else if (i < 2 * rows_valid + p.N) tma_load_1d(st_B(s) + (i - 2 * rows_valid) * LTP, gB + B_off(i - 2 * rows_valid) + l0, LT * 4, &mbar[s]);
      else tma_load_1d(st_C(s) + (i - 2 * rows_valid - p.N) * LTP, gC + C_off(i - 2 * rows_valid - p.N) + l0, LT * 4, &mbar[s]);
    }
  };

  const int ntiles = (L + LT - 1) / LT;
  auto tile_is_tma = [&](int t) { return tma && (t + 1) * LT <= L; };
  __syncthreads();
  if (warp == 0 && tile_is_tma(0)) issue_tile(0, 0);

  for (int t = 0; t < ntiles; ++t) {
    const int l0 = t * LT, len = min(LT, L - l0), s = t & 1;
    float* s_dl = st_dl(s);
    float* s_u = st_u(s);
    float* s_B = st_B(s);
    float* s_C = st_C(s);
    // prefetch the next tile into the other stage (its previous tile was released by the barrier ending iteration t-1)
    if (warp == 0 && t + 1 < ntiles && tile_is_tma(t + 1)) issue_tile(t + 1, s ^ 1);
    if (tile_is_tma(t)) {
      mbar_wait(&mbar[s], (t >> 1) & 1);
    } else {
      stage_rows<LT, LTP>(s_u, p.u, p.io_dtype, u_off, CH, rows_valid, l0, len, so);
      stage_rows<LT, LTP>(s_dl, p.delta, p.io_dtype, dl_off, CH, rows_valid, l0, len, so);
      stage_rows<LT, LTP>(s_B, p.Bm, p.io_dtype, B_off, NP, p.N, l0, len, so);
      stage_rows<LT, LTP>(s_C, p.Cm, p.io_dtype, C_off, NP, p.N, l0, len, so);
      __syncthreads();
    }
    // activate delta once per element (not once per state lane): delta = softplus(raw + bias); du = delta * u
    for (int i = tid; i < CH * (LT / 4); i += kThreads) {
      const int r = i / (LT / 4), c = (i - r * (LT / 4)) * 4;
      float4 dv = *reinterpret_cast<const float4*>(s_dl + r * LTP + c);
      float4 uv = *reinterpret_cast<const float4*>(s_u + r * LTP + c);
      const float bias = s_bias[r];
      float4 du;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float x = f4_at(dv, e) + bias;
        if (p.softplus) x = softplus20(x);
        const bool live = c + e < len;       // beyond the end of the sequence the state is frozen (a = 1, b = 0)
        f4_at(dv, e) = live ? x : 0.f;
        f4_at(du, e) = live ? x * f4_at(uv, e) : 0.f;
      }
      *reinterpret_cast<float4*>(s_dl + r * LTP + c) = dv;
      *reinterpret_cast<float4*>(s_du + r * LTP + c) = du;
    }
    __syncthreads();

    for (int i4 = 0; i4 < LT / 4; i4 += R) {
      float yacc[RPT][R * 4];
#pragma unroll
      for (int gq = 0; gq < R; ++gq) {
        const int c = (i4 + gq) * 4;
        float4 Bv[NS], Cv[NS];
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          Bv[j] = *reinterpret_cast<const float4*>(s_B + (j * R + q) * LTP + c);
          Cv[j] = *reinterpret_cast<const float4*>(s_C + (j * R + q) * LTP + c);
        }
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
          float4 dv = *reinterpret_cast<const float4*>(s_dl + rk[k] * LTP + c);
          float4 du = *reinterpret_cast<const float4*>(s_du + rk[k] * LTP + c);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float de = f4_at(dv, e), ue = f4_at(du, e);
            float y = 0.f;
#pragma unroll
            for (int j = 0; j < NS; ++j) {
              const float a = ex2f(de * A2[k][j]);
              h[k][j] = fmaf(a, h[k][j], ue * f4_at(Bv[j], e));
              y = fmaf(h[k][j], f4_at(Cv[j], e), y);
            }
            yacc[k][gq * 4 + e] = y;
          }
        }
      }
#pragma unroll
      for (int k = 0; k < RPT; ++k) {
        reduce_scatter_groups<R>(yacc[k], q);
        const int c = (i4 + q) * 4;                 // this lane owns group q of the R groups just finished
        if (p.out != nullptr && rk[k] < rows_valid && c < len) {
          const float4 uv = *reinterpret_cast<const float4*>(s_u + rk[k] * LTP + c);
          const float Dd = s_D[rk[k]];
          const float4 y4 = make_float4(fmaf(Dd, uv.x, yacc[k][0]), fmaf(Dd, uv.y, yacc[k][1]),
                                        fmaf(Dd, uv.z, yacc[k][2]), fmaf(Dd, uv.w, yacc[k][3]));
          store_scan4(p.out, p.out_dtype, out_boff + (int64_t)(d0 + rk[k]) * p.out_ds, l0 + c, l0 + len, y4, so,
                      p.accum != 0);
        }
      }
      // state checkpoint at the end of every SS2D_CHUNK elements (for the backward's recompute)
      if (p.ckpt != nullptr && ((i4 + R) & (SS2D_CHUNK / 4 - 1)) == 0) {
        const int chunk = (l0 + (i4 + R) * 4) / SS2D_CHUNK - 1;
        if (chunk < p.nck) {
#pragma unroll
          for (int k = 0; k < RPT; ++k) {
            if (rk[k] < rows_valid) {
              float* dst = p.ckpt + (((int64_t)b * p.dim + d0 + rk[k]) * p.nck + chunk) * NP + q * NS;
              if (NS == 4) {
                *reinterpret_cast<float4*>(dst) = make_float4(h[k][0], h[k][1 % NS], h[k][2 % NS], h[k][3 % NS]);
              } else {
#pragma unroll
                for (int j = 0; j < NS; ++j) dst[j] = h[k][j];
              }
            }
          }
        }
      }
    }
    // this stage was written through the generic proxy (activated delta); order that before the TMA that refills it
    if (tma) fence_proxy_async();
    __syncthreads();
  }
  if (p.last_state != nullptr) {
#pragma unroll
    for (int k = 0; k < RPT; ++k)
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        const int n = j * R + q;
        if (rk[k] < rows_valid && n < p.N) {
          const int64_t slot = ((int64_t)b * p.dim + d0 + rk[k]) * p.A_ld + n;
          if (p.last_il) { p.last_state[2 * slot] = 0.f; p.last_state[2 * slot + 1] = h[k][j]; }
          else p.last_state[slot] = h[k][j];
        }
      }
  }
}

template <int NS, int R, int RPT>
static cudaError_t launch_fwd(const ScanParams& p, cudaStream_t stream) {
  using S = FwdShape<NS, R, RPT>;
  auto kern = scan_fwd_kernel<NS, R, RPT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::smem_bytes);
  if (e != cudaSuccess) return e;
  dim3 grid((p.dpg + S::CH - 1) / S::CH, p.G, p.batch);
  kern<<<grid, kThreads, S::smem_bytes, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t scan_fwd_dispatch(const ScanParams& p, cudaStream_t stream) {
  const Variant v = pick_variant(p.N);
  if (v.NS == 1) return launch_fwd<1, 1, 1>(p, stream);
  if (v.NS == 2) return launch_fwd<2, 1, 1>(p, stream);
  if (v.R == 1) return launch_fwd<4, 1, 1>(p, stream);
  if (v.R == 2) return launch_fwd<4, 2, 1>(p, stream);
  if (v.R == 4) return launch_fwd<4, 4, 1>(p, stream);
  return launch_fwd<4, 8, 1>(p, stream);
}

}  // namespace ss2d
