// Selective-scan backward for sm_100a.
//
// Replaces selective_scan_bwd_kernel (/root/reference/gm-unet/kernels/selective_scan/csrc/selective_scan/cus/
// selective_scan_bwd_kernel.cuh:66-273) and, in NATURAL layout, the backward of CrossScan*/CrossMerge*
// (model/gm/csms6s.py). Same work decomposition as scan_fwd.cu. Tiles of SS2D_CHUNK scan positions are
// visited last-to-first; inside a tile the states h are RECOMPUTED from the chunk checkpoint written by the
// forward (never stored in HBM per element), kept in shared memory, and consumed by the reverse adjoint
// recurrence  g_l = C_l dy_l + a_(l+1) g_(l+1).  dB/dC are summed over the rows of the CTA in registers /
// shuffles / per-warp slabs before they touch global memory (the reference issues one fp32 atomic per row
// and element); dA, dD, d(delta_bias) go through a per-batch partial buffer and a deterministic second pass.
#include "scan_params.h"
#include "scan_tile.cuh"

namespace ss2d {

constexpr int kBwdLT = SS2D_CHUNK;
constexpr int kBwdLTP = kBwdLT + 4;

template <int NS, int R, int RPT>
struct BwdShape {
  static constexpr int RL = 32 / R;
  static constexpr int RPW = RL * RPT;
  static constexpr int CH = 4 * RPW;
  static constexpr int NP = NS * R;
  static constexpr int CNT = 2 * NS * 4;                 // dB/dC partials per thread and 4-element group
  static constexpr size_t tile_floats = (size_t)(4 * CH + 2 * NP) * kBwdLTP;
  static constexpr size_t slab_floats = (size_t)4 * 2 * NP * kBwdLTP;
  static constexpr size_t h_floats = (size_t)kBwdLT * RPT * kThreads * NS;
  static constexpr size_t smem_bytes = (tile_floats + slab_floats + h_floats + 2 * CH) * 4;
};

// Reduce-scatter of CNT per-lane values over LANES lanes spaced STRIDE apart (lane bits consumed MSB first).
// If CNT >= LANES each lane ends with CNT/LANES totals (slice index = its lane id among LANES); otherwise the
// value index is given by the top log2(CNT) lane bits and the remaining lanes hold replicas.
template <int STRIDE, int CNT, int S>
struct LaneReduceScatter {
  static __device__ __forceinline__ void run(float* v, int lane_id) {
    if constexpr (S >= 1) {
      if constexpr (CNT > 1) {
        constexpr int HALF = CNT / 2;
        const bool up = (lane_id & S) != 0;
#pragma unroll
        for (int i = 0; i < HALF; ++i) {
          const float send = up ? v[i] : v[i + HALF];
          const float keep = up ? v[i + HALF] : v[i];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, S * STRIDE);
        }
        LaneReduceScatter<STRIDE, HALF, S / 2>::run(v, lane_id);
      } else {
        v[0] += __shfl_xor_sync(0xffffffffu, v[0], S * STRIDE);
        LaneReduceScatter<STRIDE, 1, S / 2>::run(v, lane_id);
      }
    }
  }
};
template <int LANES, int STRIDE, int CNT>
__device__ __forceinline__ void lane_reduce_scatter(float* v, int lane_id) {
  LaneReduceScatter<STRIDE, CNT, LANES / 2>::run(v, lane_id);
}

template <int NS>
__device__ __forceinline__ void load_states(const float* __restrict__ src, float* dst) {
  if (NS == 4) {
    const float4 v = *reinterpret_cast<const float4*>(src);
    dst[0] = v.x; dst[1 % NS] = v.y; dst[2 % NS] = v.z; dst[3 % NS] = v.w;
  } else if (NS == 2) {
    const float2 v = *reinterpret_cast<const float2*>(src);
    dst[0] = v.x; dst[1 % NS] = v.y;
  } else {
    dst[0] = src[0];
  }
}

template <int NS, int R, int RPT>
__global__ void __launch_bounds__(kThreads) scan_bwd_kernel(const ScanParams p) {
  using S = BwdShape<NS, R, RPT>;
  constexpr int LT = kBwdLT, LTP = kBwdLTP, CH = S::CH, NP = S::NP, RL = S::RL, CNT = S::CNT;
  extern __shared__ __align__(16) float smem[];
  float* s_dl = smem;                    // delta (activated); reused for the d(delta) tile   [CH][LTP]
  float* s_u = s_dl + CH * LTP;          // u
  float* s_du = s_u + CH * LTP;          // delta * u
  float* s_dy = s_du + CH * LTP;         // dout; reused for the du tile
  float* s_B = s_dy + CH * LTP;          // [NP][LTP]
  float* s_C = s_B + NP * LTP;
  float* s_slab = s_C + NP * LTP;        // [4 warps][2][NP][LTP] per-warp dB / dC sums
  float* s_h = s_slab + S::slab_floats;  // [LT][RPT][128][NS] recomputed states
  float* s_bias = s_h + S::h_floats;     // [CH]
  float* s_D = s_bias + CH;              // [CH]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q = lane % R, rl = lane / R;
  const int b = blockIdx.z, g = blockIdx.y;
  const int row0 = blockIdx.x * CH;
  const int rows_valid = min(CH, p.dpg - row0);
  const int d0 = g * p.dpg + row0;
  const int L = p.L;
  ScanOrder so;
  so.dir = p.layout == SS2D_LAYOUT_NATURAL ? p.dirs[g] : 0;
  so.H = p.H; so.W = p.W; so.L = L;

  for (int r = tid; r < CH; r += kThreads) {
    const bool ok = r < rows_valid;
    s_bias[r] = (ok && p.bias) ? p.bias[d0 + r] : 0.f;
    s_D[r] = (ok && p.Dv && !p.accum) ? p.Dv[d0 + r] : 0.f;
  }

  int rk[RPT];
  float A1[RPT][NS], A2[RPT][NS], carry[RPT][NS], dA[RPT][NS], dDacc[RPT], dbacc[RPT];
#pragma unroll
  for (int k = 0; k < RPT; ++k) {
    rk[k] = warp * S::RPW + k * RL + rl;
    dDacc[k] = 0.f; dbacc[k] = 0.f;
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      const int n = j * R + q;
      A1[k][j] = (rk[k] < rows_valid && n < p.N) ? p.A[(int64_t)(d0 + rk[k]) * p.A_ld + n] : 0.f;
      A2[k][j] = A1[k][j] * kLog2e;
      carry[k][j] = 0.f;     // a_(l+1) * g_(l+1), zero past the end of the sequence
      dA[k][j] = 0.f;
    }
  }

  const int64_t u_boff = (int64_t)b * p.u_bs, dl_boff = (int64_t)b * p.dl_bs, out_boff = (int64_t)b * p.out_bs;
  auto u_off = [&](int r) { const int d = d0 + r; return u_boff + (int64_t)(p.u_mod > 0 ? d % p.u_mod : d) * p.u_ds; };
  // du has u's strides, except when the groups share u (u_mod > 0): then it is a dense (batch, dim, L) tensor
  auto du_off = [&](int r) {
    return p.u_mod > 0 ? ((int64_t)b * p.dim + d0 + r) * (int64_t)L : u_boff + (int64_t)(d0 + r) * p.u_ds;
  };
  auto dl_off = [&](int r) { return dl_boff + (int64_t)(d0 + r) * p.dl_ds; };
  auto dy_off = [&](int r) { const int d = d0 + r; return out_boff + (int64_t)(p.u_mod > 0 ? d % p.u_mod : d) * p.out_ds; };
  const int64_t B_base = (int64_t)b * p.B_bs + (int64_t)g * p.B_gs;
  const int64_t C_base = (int64_t)b * p.C_bs + (int64_t)g * p.C_gs;
  auto B_off = [&](int n) { return B_base + (int64_t)n * p.B_ns; };
  auto C_off = [&](int n) { return C_base + (int64_t)n * p.C_ns; };
  const bool single_cta_group = gridDim.x == 1;
  const int ntiles = (L + LT - 1) / LT;

  for (int t = ntiles - 1; t >= 0; --t) {
    const int l0 = t * LT, len = min(LT, L - l0);
    __syncthreads();
    stage_rows<LT, LTP>(s_u, p.u, p.io_dtype, u_off, CH, rows_valid, l0, len, so);
    stage_rows<LT, LTP>(s_dl, p.delta, p.io_dtype, dl_off, CH, rows_valid, l0, len, so);
    stage_rows<LT, LTP>(s_dy, p.dout, p.out_dtype, dy_off, CH, rows_valid, l0, len, so);
    stage_rows<LT, LTP>(s_B, p.Bm, p.io_dtype, B_off, NP, p.N, l0, len, so);
    stage_rows<LT, LTP>(s_C, p.Cm, p.io_dtype, C_off, NP, p.N, l0, len, so);
    __syncthreads();
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
        if (c + e >= len) x = 0.f;
        f4_at(dv, e) = x;
        f4_at(du, e) = x * f4_at(uv, e);
      }
      *reinterpret_cast<float4*>(s_dl + r * LTP + c) = dv;
      *reinterpret_cast<float4*>(s_du + r * LTP + c) = du;
    }
    __syncthreads();

    // ---- forward recompute of h over the tile, from the checkpoint at the end of the previous chunk ----
    float h0[RPT][NS];
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
#pragma unroll
      for (int j = 0; j < NS; ++j) h0[k][j] = 0.f;
      if (t > 0 && rk[k] < rows_valid) {
        const float* src = p.ckpt_in + (((int64_t)b * p.dim + d0 + rk[k]) * p.nck + (t - 1)) * NP + q * NS;
#pragma unroll
        for (int j = 0; j < NS; ++j) h0[k][j] = __ldg(src + j);
      }
    }
    {
      float h[RPT][NS];
#pragma unroll
      for (int k = 0; k < RPT; ++k)
#pragma unroll
        for (int j = 0; j < NS; ++j) h[k][j] = h0[k][j];
      for (int i4 = 0; i4 < LT / 4; ++i4) {
        const int c = i4 * 4;
        float4 Bv[NS];
#pragma unroll
        for (int j = 0; j < NS; ++j) Bv[j] = *reinterpret_cast<const float4*>(s_B + (j * R + q) * LTP + c);
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
          float4 dv = *reinterpret_cast<const float4*>(s_dl + rk[k] * LTP + c);
          float4 du = *reinterpret_cast<const float4*>(s_du + rk[k] * LTP + c);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float de = f4_at(dv, e), ue = f4_at(du, e);
            float* hs = s_h + ((size_t)((c + e) * RPT + k) * kThreads + tid) * NS;
#pragma unroll
            for (int j = 0; j < NS; ++j) {
              const float a = ex2f(de * A2[k][j]);
              h[k][j] = fmaf(a, h[k][j], ue * f4_at(Bv[j], e));
            }
            if (NS == 4) *reinterpret_cast<float4*>(hs) = make_float4(h[k][0], h[k][1 % NS], h[k][2 % NS], h[k][3 % NS]);
            else if (NS == 2) *reinterpret_cast<float2*>(hs) = make_float2(h[k][0], h[k][1 % NS]);
            else hs[0] = h[k][0];
          }
        }
      }
    }
    // each thread re-reads only what it wrote itself: no barrier needed before the reverse sweep

    // ---- reverse adjoint sweep ----
    for (int i4 = LT / 4 - 1; i4 >= 0; --i4) {
      const int c = i4 * 4;
      float4 Bv[NS], Cv[NS];
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        Bv[j] = *reinterpret_cast<const float4*>(s_B + (j * R + q) * LTP + c);
        Cv[j] = *reinterpret_cast<const float4*>(s_C + (j * R + q) * LTP + c);
      }
      float part[CNT];     // [dB | dC][NS][4], summed over this thread's RPT rows
#pragma unroll
      for (int i = 0; i < CNT; ++i) part[i] = 0.f;
#pragma unroll
      for (int k = 0; k < RPT; ++k) {
        float4 dv = *reinterpret_cast<const float4*>(s_dl + rk[k] * LTP + c);
        float4 du = *reinterpret_cast<const float4*>(s_du + rk[k] * LTP + c);
        float4 dy = *reinterpret_cast<const float4*>(s_dy + rk[k] * LTP + c);
        float sums[8];     // [e][sB | sA]
#pragma unroll
        for (int e = 3; e >= 0; --e) {
          const float de = f4_at(dv, e), ue = f4_at(du, e), dye = f4_at(dy, e);
          float hc[NS], hp[NS];
          load_states<NS>(s_h + ((size_t)((c + e) * RPT + k) * kThreads + tid) * NS, hc);
          if (c + e > 0) {
            load_states<NS>(s_h + ((size_t)((c + e - 1) * RPT + k) * kThreads + tid) * NS, hp);
          } else {
#pragma unroll
            for (int j = 0; j < NS; ++j) hp[j] = h0[k][j];
          }
          float sB = 0.f, sA = 0.f;
#pragma unroll
          for (int j = 0; j < NS; ++j) {
            const float a = ex2f(de * A2[k][j]);
            const float gj = fmaf(f4_at(Cv[j], e), dye, carry[k][j]);
            part[(0 * NS + j) * 4 + e] = fmaf(gj, ue, part[(0 * NS + j) * 4 + e]);
            part[(1 * NS + j) * 4 + e] = fmaf(dye, hc[j], part[(1 * NS + j) * 4 + e]);
            sB = fmaf(gj, f4_at(Bv[j], e), sB);
            const float tj = gj * a;
            carry[k][j] = tj;
            const float w = tj * hp[j];
            sA = fmaf(w, A1[k][j], sA);
            dA[k][j] = fmaf(w, de, dA[k][j]);
          }
          sums[e * 2 + 0] = sB;
          sums[e * 2 + 1] = sA;
        }
        // combine the R state lanes of the row; afterwards lane q owns 4/R elements (or half of one for R = 8)
        lane_reduce_scatter<R, 1, 8>(sums, q);
        constexpr int PER = (8 / R) > 0 ? (8 / R) : 1;
        float other = 0.f;
        if (R == 8) other = __shfl_xor_sync(0xffffffffu, sums[0], 1);
        if (R < 8 || (q & 1) == 0) {
          constexpr int NE = R == 8 ? 1 : PER / 2;     // elements owned
          const int e0 = R == 8 ? (q >> 1) : q * NE;
#pragma unroll
          for (int ee = 0; ee < NE; ++ee) {
            const int e = e0 + ee;
            const float sB = sums[ee * 2 + 0];
            const float sA = R == 8 ? other : sums[ee * 2 + 1];
            const int idx = rk[k] * LTP + c + e;
            const float de = s_dl[idx], uu = s_u[idx], dye = s_dy[idx];
            const float du_out = fmaf(s_D[rk[k]], dye, de * sB);
            float ddl = fmaf(uu, sB, sA);
            if (p.softplus) {   // sigmoid(raw) = 1 - exp(-softplus(raw)); series for small delta avoids cancellation
              const float sig = de < 0.015625f ? de * (1.f - de * (0.5f - de * 0.16666667f)) : 1.f - ex2f(-de * kLog2e);
              ddl *= sig;
            }
            if (c + e >= len) ddl = 0.f;
            dDacc[k] = fmaf(dye, uu, dDacc[k]);
            dbacc[k] += ddl;
            // the row's lanes are done with dy/delta of this group (the shuffles above ordered them): reuse the tiles
            s_dy[idx] = du_out;
            s_dl[idx] = ddl;
          }
        }
      }
      // dB/dC: sum over the RL row lanes of the warp, then park the warp's totals in its slab
      lane_reduce_scatter<RL, R, CNT>(part, rl);
      {
        constexpr int PER = CNT >= RL ? CNT / RL : 1;
        constexpr int REP = CNT >= RL ? 1 : RL / CNT;          // replicas when there are fewer values than lanes
        if ((rl % REP) == 0) {
          const int slice = rl / REP;
          float* slab = s_slab + (size_t)warp * 2 * NP * LTP;
#pragma unroll
          for (int i = 0; i < PER; ++i) {
            const int vi = slice * PER + i;                     // index into [dB | dC][NS][4]
            const int which = vi / (NS * 4), j = (vi / 4) % NS, e = vi & 3;
            slab[(which * NP + j * R + q) * LTP + c + e] = part[i];
          }
        }
      }
    }
    __syncthreads();
    // ---- tile epilogue: du / d(delta) tiles and the CTA's dB / dC sums go to global memory ----
    for (int i = tid; i < CH * (LT / 4); i += kThreads) {
      const int r = i / (LT / 4), c = (i - r * (LT / 4)) * 4;
      if (r < rows_valid && c < len) {
        const float4 a = *reinterpret_cast<const float4*>(s_dy + r * LTP + c);
        const float4 d = *reinterpret_cast<const float4*>(s_dl + r * LTP + c);
        store_scan4(p.du, p.io_dtype, du_off(r), l0 + c, l0 + len, a, so, p.accum != 0);
        store_scan4(p.ddelta, p.io_dtype, dl_off(r), l0 + c, l0 + len, d, so, p.accum != 0);
      }
    }
    for (int i = tid; i < 2 * NP * LT; i += kThreads) {
      const int which = i / (NP * LT), n = (i / LT) % NP, c = i % LT;
      if (n < p.N && c < len) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) v += s_slab[((size_t)(w * 2 + which) * NP + n) * LTP + c];
        float* dst = which == 0 ? p.dB : p.dC;
        const int64_t idx = ((int64_t)(b * p.G + g) * p.A_ld + n) * L + so.natural(l0 + c);
        if (single_cta_group) dst[idx] = v;
        else atomicAdd(dst + idx, v);
      }
    }
  }

  // ---- per-(batch, channel) partials of dA, dD, d(delta_bias) ----
#pragma unroll
  for (int k = 0; k < RPT; ++k) {
    float dd = dDacc[k], db = dbacc[k];
#pragma unroll
    for (int s = R / 2; s >= 1; s >>= 1) {
      dd += __shfl_xor_sync(0xffffffffu, dd, s);
      db += __shfl_xor_sync(0xffffffffu, db, s);
    }
    if (rk[k] < rows_valid) {
      float* dst = p.part + ((int64_t)b * p.dim + d0 + rk[k]) * (p.N + 2);
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        const int n = j * R + q;
        if (n < p.N) dst[n] = dA[k][j];
      }
      if (q == 0) { dst[p.N] = dd; dst[p.N + 1] = db; }
    }
  }
}

// dA[d][n] = sum_b part[b][d][n]; dD[d], dbias[d] likewise. One thread per (d, slot), batch-sequential: deterministic.
__global__ void scan_bwd_finalize_kernel(const float* __restrict__ part, float* __restrict__ dA, float* __restrict__ dD,
                                         float* __restrict__ dbias, int batch, int dim, int N, int A_ld, int accum) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= dim * (N + 2)) return;
  const int d = i / (N + 2), s = i - d * (N + 2);
  float acc = 0.f;
  for (int b = 0; b < batch; ++b) acc += part[((int64_t)b * dim + d) * (N + 2) + s];
  if (s < N) dA[(int64_t)d * A_ld + s] = acc;
  else if (s == N) { if (dD && !accum) dD[d] = acc; }
  else if (dbias) dbias[d] = accum ? dbias[d] + acc : acc;
}

template <int NS, int R, int RPT>
static cudaError_t launch_bwd(const ScanParams& p, cudaStream_t stream) {
  using S = BwdShape<NS, R, RPT>;
  auto kern = scan_bwd_kernel<NS, R, RPT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::smem_bytes);
  if (e != cudaSuccess) return e;
  dim3 grid((p.dpg + S::CH - 1) / S::CH, p.G, p.batch);
  kern<<<grid, kThreads, S::smem_bytes, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t scan_bwd_dispatch(const ScanParams& p, cudaStream_t stream) {
  const Variant v = pick_variant(p.N);
  if (v.NS == 1) return launch_bwd<1, 1, 1>(p, stream);
  if (v.NS == 2) return launch_bwd<2, 1, 1>(p, stream);
  if (v.R == 1) return launch_bwd<4, 1, 1>(p, stream);
  if (v.R == 2) return launch_bwd<4, 2, 1>(p, stream);
  if (v.R == 4) return launch_bwd<4, 4, 1>(p, stream);
  return launch_bwd<4, 8, 1>(p, stream);
}

cudaError_t scan_bwd_finalize(const ScanParams& p, float* dA, float* dD, float* dbias, cudaStream_t stream) {
  const int total = p.dim * (p.N + 2);
  scan_bwd_finalize_kernel<<<(total + 255) / 256, 256, 0, stream>>>(p.part, dA, dD, dbias, p.batch, p.dim, p.N, p.A_ld,
                                                                  p.accum);
  return cudaGetLastError();
}

}  // namespace ss2d
