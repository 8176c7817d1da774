// Selective scan for d_state = 1, fp32, contiguous traversals — the lean kernels the live GM-UNet regime runs on
// (4 single-direction SS2Ds with N = 1 per GroupMambaLayer: D in {16, 32, 87, 112} rows per (batch, direction),
// L in {3136, 784, 196, 49}). Replaces selective_scan_{fwd,bwd}_kernel
// (/root/reference/gm-unet/kernels/selective_scan/csrc/selective_scan/cus/selective_scan_fwd_kernel.cuh:61-172,
// selective_scan_bwd_kernel.cuh:66-273) for that regime; every other dtype / layout keeps scan_par.cu.
//
// Same algorithm as scan_par.cu — the scan is parallel ALONG L: a row is owned by seg x wpr threads, each scanning
// P = 8 consecutive positions in registers; the (decay, state) aggregates are combined by a Kogge-Stone shuffle scan inside
// the warp and one shared-memory hop between the warps of a row — but built for calls that move 2 ... 60 MB:
//   * ONE code path. ncu on scan_par (profiles/r2_ncu_scan_par_small_shapes.txt) showed its 7 700-instruction bodies (generic
//     dtype / direction / tail handling inlined next to the fast path) stalled on instruction fetch (`no_instruction` 19.7
//     cycles per issue at the stage-3 shape): every warp runs the code exactly once. Here eligibility is decided on the
//     host, the geometry (lanes per row, warps per row, rows per CTA) is a kernel ARGUMENT, and a body is ~10x smaller.
//   * the whole row in flight at once: up to 16 warps per row, so a 56^2 row is one pass (scan_par: two sequential chunks).
//   * all loads of a thread (delta, u, B, C [, dout]) are issued before any arithmetic.
//   * backward: dA / dD / d(delta_bias) are added straight into caller-zeroed accumulators (`grads_prezeroed`), so the
//     call is memset + ONE kernel (scan_par: memset + kernel + finalize).
#include "common.cuh"
#include "host_util.h"
#include "scan_params.h"

namespace ss2d {

constexpr int N1_P = 8;             // positions per thread
constexpr int N1_MAX_WARPS = 16;    // warps per CTA

struct N1Geom {
  int seg;       // lanes of one row inside a warp: 4, 8, 16 or 32
  int wpr;       // warps per row (1 when seg < 32)
  int rt;        // rows per CTA
  int lc;        // positions per chunk = seg * wpr * P
  int nchunks;   // ceil(L / lc)
};

template <bool VEC>
__device__ __forceinline__ void n1_load(const float* __restrict__ row, int l0, int L, bool rev, float (&o)[N1_P]) {
  if constexpr (VEC) {
#pragma unroll
    for (int c = 0; c < N1_P; c += 4) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (l0 + c < L) v = __ldg(reinterpret_cast<const float4*>(row + (rev ? L - 4 - (l0 + c) : l0 + c)));
      o[c] = rev ? v.w : v.x; o[c + 1] = rev ? v.z : v.y; o[c + 2] = rev ? v.y : v.z; o[c + 3] = rev ? v.x : v.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < N1_P; ++i) o[i] = l0 + i < L ? __ldg(row + (rev ? L - 1 - (l0 + i) : l0 + i)) : 0.f;
  }
}
template <bool VEC>
__device__ __forceinline__ void n1_store(float* __restrict__ row, int l0, int L, bool rev, const float (&v)[N1_P]) {
  if constexpr (VEC) {
#pragma unroll
    for (int c = 0; c < N1_P; c += 4) {
      if (l0 + c < L) {
        const float4 q = rev ? make_float4(v[c + 3], v[c + 2], v[c + 1], v[c]) : make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
        *reinterpret_cast<float4*>(row + (rev ? L - 4 - (l0 + c) : l0 + c)) = q;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < N1_P; ++i)
      if (l0 + i < L) row[rev ? L - 1 - (l0 + i) : l0 + i] = v[i];
  }
}

// Exclusive carry for every thread of a row + the row total, given the thread's own affine map h -> Pm h + Hm.
// UP: composition of the EARLIER threads (prefix); !UP: of the LATER threads (suffix).
template <bool UP>
__device__ __forceinline__ void n1_row_exclusive(float Pm, float Hm, int lane, int lis, int seg, int wir, int wpr, int slot0,
                                                 float (*s_agg)[2], float& Pex, float& Hex, float& Ptot, float& Htot) {
  for (int off = 1; off < seg; off <<= 1) {
    const float Pp = UP ? __shfl_up_sync(0xffffffffu, Pm, off) : __shfl_down_sync(0xffffffffu, Pm, off);
    const float Hp = UP ? __shfl_up_sync(0xffffffffu, Hm, off) : __shfl_down_sync(0xffffffffu, Hm, off);
    const bool take = UP ? (lis >= off) : (lis + off < seg);
    if (take) { Hm = fmaf(Pm, Hp, Hm); Pm *= Pp; }
  }
  float Pe = UP ? __shfl_up_sync(0xffffffffu, Pm, 1) : __shfl_down_sync(0xffffffffu, Pm, 1);
  float He = UP ? __shfl_up_sync(0xffffffffu, Hm, 1) : __shfl_down_sync(0xffffffffu, Hm, 1);
  if (UP ? (lis == 0) : (lis == seg - 1)) { Pe = 1.f; He = 0.f; }
  if (wpr == 1) {
    const int src = lane - lis + (UP ? seg - 1 : 0);
    Ptot = __shfl_sync(0xffffffffu, Pm, src);
    Htot = __shfl_sync(0xffffffffu, Hm, src);
    Pex = Pe; Hex = He;
  } else {
    if (UP ? (lis == 31) : (lis == 0)) { s_agg[slot0 + wir][0] = Pm; s_agg[slot0 + wir][1] = Hm; }
    __syncthreads();
    float Pq = 1.f, Hq = 0.f;
    Ptot = 1.f; Htot = 0.f;
    for (int i = 0; i < wpr; ++i) {
      const int w = UP ? i : wpr - 1 - i;
      const float Pw = s_agg[slot0 + w][0], Hw = s_agg[slot0 + w][1];
      if (UP ? (w < wir) : (w > wir)) { Hq = fmaf(Pw, Hq, Hw); Pq *= Pw; }
      Htot = fmaf(Pw, Htot, Hw); Ptot *= Pw;
    }
    Hex = fmaf(Pe, Hq, He); Pex = Pe * Pq;
  }
}

struct N1Thread {
  int r, wir, lis, t;     // row inside the CTA, warp inside the row, lane inside the row segment, thread index along the row
};
__device__ __forceinline__ N1Thread n1_thread(const N1Geom gm) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  N1Thread q;
  if (gm.seg == 32) { q.r = warp / gm.wpr; q.wir = warp - q.r * gm.wpr; q.lis = lane; }
  else { q.r = warp * (32 / gm.seg) + lane / gm.seg; q.wir = 0; q.lis = lane % gm.seg; }
  q.t = q.wir * 32 + q.lis;
  return q;
}

// ------------------------------------------------------------------------------------------------- forward
template <bool VEC>
__global__ void __launch_bounds__(N1_MAX_WARPS * 32) scan_n1_fwd_kernel(const ScanParams p, const N1Geom gm) {
  __shared__ float s_agg[2][N1_MAX_WARPS][2];
  const N1Thread q = n1_thread(gm);
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.z, g = blockIdx.y;
  const int row = blockIdx.x * gm.rt + q.r;
  const bool valid = q.r < gm.rt && row < p.dpg;
  const int d = g * p.dpg + (valid ? row : 0);
  const int L = valid ? p.L : 0;                         // invalid rows: every load / store predicated off
  const bool rev = p.layout == SS2D_LAYOUT_NATURAL && p.dirs[g] == 3;

  const float A2 = p.A[d] * kLog2e;
  const float bias = p.bias ? p.bias[d] : 0.f;
  const float Dd = p.Dv ? p.Dv[d] : 0.f;
  const float* fu = static_cast<const float*>(p.u) + (int64_t)b * p.u_bs + (int64_t)(p.u_mod > 0 ? d % p.u_mod : d) * p.u_ds;
  const float* fd = static_cast<const float*>(p.delta) + (int64_t)b * p.dl_bs + (int64_t)d * p.dl_ds;
  const float* fB = static_cast<const float*>(p.Bm) + (int64_t)b * p.B_bs + (int64_t)g * p.B_gs;
  const float* fC = static_cast<const float*>(p.Cm) + (int64_t)b * p.C_bs + (int64_t)g * p.C_gs;
  float* fo = p.out ? static_cast<float*>(p.out) + (int64_t)b * p.out_bs + (int64_t)d * p.out_ds : nullptr;

  float h_carry = 0.f;
  for (int c = 0; c < gm.nchunks; ++c) {
    const int l0 = c * gm.lc + q.t * N1_P;
    float dl[N1_P], uu[N1_P], bb[N1_P], cc[N1_P];
    n1_load<VEC>(fd, l0, L, rev, dl);
    n1_load<VEC>(fu, l0, L, rev, uu);
    n1_load<VEC>(fB, l0, L, rev, bb);
    n1_load<VEC>(fC, l0, L, rev, cc);
    // local scan from h = 0: hl = local state, pc = cumulative decay; beyond the end of the row the state is frozen
    float hl[N1_P], pc[N1_P];
    float h = 0.f, pm = 1.f;
#pragma unroll
    for (int i = 0; i < N1_P; ++i) {
      float x = dl[i] + bias;
      if (p.softplus) x = softplus20(x);
      if (l0 + i >= L) x = 0.f;
      const float a = ex2f(x * A2);
      h = fmaf(a, h, x * uu[i] * bb[i]);
      pm *= a;
      hl[i] = h; pc[i] = pm;
    }
    float Pex, Hex, Ptot, Htot;
    n1_row_exclusive<true>(pm, h, lane, q.lis, gm.seg, q.wir, gm.wpr, q.r * gm.wpr, s_agg[c & 1], Pex, Hex, Ptot, Htot);
    const float h_in = fmaf(Pex, h_carry, Hex);             // state entering this thread's first position
    if (fo) {
      float y[N1_P];
#pragma unroll
      for (int i = 0; i < N1_P; ++i) y[i] = fmaf(cc[i], fmaf(pc[i], h_in, hl[i]), Dd * uu[i]);
      n1_store<VEC>(fo, l0, L, rev, y);
    }
    // chunk checkpoint every SS2D_CHUNK positions: the thread whose last position closes a chunk owns it
    if (p.ckpt != nullptr && valid && ((l0 + N1_P) % SS2D_CHUNK) == 0) {
      const int idx = (l0 + N1_P) / SS2D_CHUNK - 1;
      if (idx < p.nck) p.ckpt[((int64_t)b * p.dim + d) * p.nck + idx] = fmaf(pc[N1_P - 1], h_in, hl[N1_P - 1]);
    }
    h_carry = fmaf(Ptot, h_carry, Htot);
  }
  if (p.last_state != nullptr && valid && q.t == 0) {
    const int64_t slot = (int64_t)b * p.dim + d;
    if (p.last_il) { p.last_state[2 * slot] = 0.f; p.last_state[2 * slot + 1] = h_carry; }
    else p.last_state[slot] = h_carry;
  }
}

// ------------------------------------------------------------------------------------------------- backward
// grads_zeroed != 0: dA / dD / d(delta_bias) are caller-zeroed accumulators (p.part unused); else per-(batch, channel)
// partials go to p.part and scan_bwd_finalize sums them.
template <bool VEC>
__global__ void __launch_bounds__(N1_MAX_WARPS * 32) scan_n1_bwd_kernel(const ScanParams p, const N1Geom gm, float* __restrict__ dA,
                                                                      float* __restrict__ dD, float* __restrict__ dbias,
                                                                      int grads_zeroed) {
  extern __shared__ __align__(16) float s_slab[];          // [2][rt][lc]: per-row dB / dC of the current chunk (rt > 1)
  __shared__ float s_agg[2][N1_MAX_WARPS][2];
  __shared__ float s_red[N1_MAX_WARPS][3];
  const N1Thread q = n1_thread(gm);
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.z, g = blockIdx.y;
  const int row = blockIdx.x * gm.rt + q.r;
  const bool valid = q.r < gm.rt && row < p.dpg;
  const int d = g * p.dpg + (valid ? row : 0);
  const int L = valid ? p.L : 0;
  const bool rev = p.layout == SS2D_LAYOUT_NATURAL && p.dirs[g] == 3;
  const bool single_cta_group = gridDim.x == 1;

  const float A1 = p.A[d];
  const float A2 = A1 * kLog2e;
  const float bias = p.bias ? p.bias[d] : 0.f;
  const float Dd = p.Dv ? p.Dv[d] : 0.f;
  const int u_ch = p.u_mod > 0 ? d % p.u_mod : d;
  const float* fu = static_cast<const float*>(p.u) + (int64_t)b * p.u_bs + (int64_t)u_ch * p.u_ds;
  const float* fd = static_cast<const float*>(p.delta) + (int64_t)b * p.dl_bs + (int64_t)d * p.dl_ds;
  const float* fy = static_cast<const float*>(p.dout) + (int64_t)b * p.out_bs + (int64_t)u_ch * p.out_ds;
  const float* fB = static_cast<const float*>(p.Bm) + (int64_t)b * p.B_bs + (int64_t)g * p.B_gs;
  const float* fC = static_cast<const float*>(p.Cm) + (int64_t)b * p.C_bs + (int64_t)g * p.C_gs;
  float* fdu = static_cast<float*>(p.du) + (p.u_mod > 0 ? ((int64_t)b * p.dim + d) * (int64_t)p.L : (int64_t)b * p.u_bs + (int64_t)d * p.u_ds);
  float* fdd = static_cast<float*>(p.ddelta) + (int64_t)b * p.dl_bs + (int64_t)d * p.dl_ds;
  float* dBrow = p.dB + (int64_t)(b * p.G + g) * p.L;
  float* dCrow = p.dC + (int64_t)(b * p.G + g) * p.L;

  float t_carry = 0.f;          // a_l g_l of the first position of the chunk after this one
  float accA = 0.f, accD = 0.f, accb = 0.f;
  for (int c = gm.nchunks - 1; c >= 0; --c) {
    const int l0 = c * gm.lc + q.t * N1_P;
    float dl[N1_P], uu[N1_P], bb[N1_P], cc[N1_P], dy[N1_P];
    n1_load<VEC>(fd, l0, L, rev, dl);
    n1_load<VEC>(fu, l0, L, rev, uu);
    n1_load<VEC>(fy, l0, L, rev, dy);
    n1_load<VEC>(fB, l0, L, rev, bb);
    n1_load<VEC>(fC, l0, L, rev, cc);
    // state entering the chunk: checkpoint written by the forward at the end of the previous SS2D_CHUNK block
    const float h_chunk = (c > 0 && valid) ? __ldg(p.ckpt_in + ((int64_t)b * p.dim + d) * p.nck + (c * gm.lc) / SS2D_CHUNK - 1) : 0.f;
    // ---- forward recompute: a, local h, decay products ----
    float a[N1_P], hh[N1_P];
    float h = 0.f, pm = 1.f;
#pragma unroll
    for (int i = 0; i < N1_P; ++i) {
      float x = dl[i] + bias;
      if (p.softplus) x = softplus20(x);
      if (l0 + i >= L) x = 0.f;
      dl[i] = x;
      a[i] = ex2f(x * A2);
      h = fmaf(a[i], h, x * uu[i] * bb[i]);
      pm *= a[i];
      hh[i] = h;
      cc[i] *= dy[i];                                        // c_l = C_l dy_l
    }
    float Pex, Hex, Ptot, Htot;
    n1_row_exclusive<true>(pm, h, lane, q.lis, gm.seg, q.wir, gm.wpr, q.r * gm.wpr, s_agg[0], Pex, Hex, Ptot, Htot);
    const float h_in = fmaf(Pex, h_chunk, Hex);              // true state before this thread's first position
    {
      float pc = 1.f;
#pragma unroll
      for (int i = 0; i < N1_P; ++i) { pc *= a[i]; hh[i] = fmaf(pc, h_in, hh[i]); }
    }
    // ---- reverse: t_l = a_l (c_l + t_(l+1)); g_l = c_l + t_(l+1) ----
    float tl[N1_P];
    float tq = 0.f, qm = 1.f;
#pragma unroll
    for (int i = N1_P - 1; i >= 0; --i) { tq = a[i] * (cc[i] + tq); qm *= a[i]; tl[i] = tq; }
    float Qex, Tex, Qtot, Ttot;
    n1_row_exclusive<false>(qm, tq, lane, q.lis, gm.seg, q.wir, gm.wpr, q.r * gm.wpr, s_agg[1], Qex, Tex, Qtot, Ttot);
    const float t_in = fmaf(Qex, t_carry, Tex);              // t of the position right after this thread's last one
    float dub[N1_P], ddl[N1_P], dBv[N1_P], dCv[N1_P];
    {
      float qc = 1.f, t_next = t_in;
#pragma unroll
      for (int i = N1_P - 1; i >= 0; --i) {
        qc *= a[i];
        const float ti = fmaf(qc, t_in, tl[i]);              // true t_i
        const float gi = cc[i] + t_next;                     // true g_i
        t_next = ti;
        const float hprev = i > 0 ? hh[i - 1] : h_in;
        const float sB = gi * bb[i];
        const float w = ti * hprev;
        float dd = fmaf(uu[i], sB, w * A1);
        if (p.softplus) {      // sigmoid(raw) = 1 - exp(-softplus(raw)); series for small delta avoids cancellation
          const float de = dl[i];
          dd *= de < 0.015625f ? de * (1.f - de * (0.5f - de * 0.16666667f)) : 1.f - ex2f(-de * kLog2e);
        }
        if (l0 + i >= L) dd = 0.f;
        dub[i] = fmaf(Dd, dy[i], dl[i] * sB);
        ddl[i] = dd;
        dBv[i] = gi * dl[i] * uu[i];
        dCv[i] = dy[i] * hh[i];
        accA = fmaf(w, dl[i], accA);
        accD = fmaf(dy[i], uu[i], accD);
        accb += dd;
      }
    }
    n1_store<VEC>(fdu, l0, L, rev, dub);
    n1_store<VEC>(fdd, l0, L, rev, ddl);
    t_carry = fmaf(Qtot, t_carry, Ttot);
    // ---- dB / dC: summed over the rows of this CTA, then one store (CTA = whole group) or reduction per 4 positions ----
    if (gm.rt == 1) {
      if (valid) {
#pragma unroll
        for (int which = 0; which < 2; ++which) {
          float* dst = which == 0 ? dBrow : dCrow;
          const float* v = which == 0 ? dBv : dCv;
          if constexpr (VEC) {
#pragma unroll
            for (int k = 0; k < N1_P; k += 4) {
              if (l0 + k < L) {
                float4* q4 = reinterpret_cast<float4*>(dst + (rev ? L - 4 - (l0 + k) : l0 + k));
                const float4 x = rev ? make_float4(v[k + 3], v[k + 2], v[k + 1], v[k]) : make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
                if (single_cta_group) *q4 = x;
                else asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(q4), "f"(x.x), "f"(x.y), "f"(x.z), "f"(x.w) : "memory");
              }
            }
          } else {
#pragma unroll
            for (int k = 0; k < N1_P; ++k) {
              if (l0 + k < L) {
                float* q1 = dst + (rev ? L - 1 - (l0 + k) : l0 + k);
                if (single_cta_group) *q1 = v[k];
                else atomicAdd(q1, v[k]);
              }
            }
          }
        }
      }
    } else {
      float* slabB = s_slab + (size_t)q.r * gm.lc + q.t * N1_P;
      float* slabC = slabB + (size_t)gm.rt * gm.lc;
      if (q.r < gm.rt) {
        *reinterpret_cast<float4*>(slabB) = make_float4(dBv[0], dBv[1], dBv[2], dBv[3]);
        *reinterpret_cast<float4*>(slabB + 4) = make_float4(dBv[4], dBv[5], dBv[6], dBv[7]);
        *reinterpret_cast<float4*>(slabC) = make_float4(dCv[0], dCv[1], dCv[2], dCv[3]);
        *reinterpret_cast<float4*>(slabC + 4) = make_float4(dCv[4], dCv[5], dCv[6], dCv[7]);
      }
      __syncthreads();
      const int L0 = p.L, c4 = gm.lc / 4;
      for (int i = threadIdx.x; i < 2 * c4; i += blockDim.x) {
        const int which = i / c4, pos = (i - which * c4) * 4;
        const int l = c * gm.lc + pos;
        if (l >= L0) continue;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const float* col = s_slab + (size_t)which * gm.rt * gm.lc + pos;
        for (int rr = 0; rr < gm.rt; ++rr) {
          const float4 v = *reinterpret_cast<const float4*>(col + (size_t)rr * gm.lc);
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        float* dst = which == 0 ? dBrow : dCrow;
        if (VEC) {
          float4* q4 = reinterpret_cast<float4*>(dst + (rev ? L0 - 4 - l : l));
          const float4 x = rev ? make_float4(acc.w, acc.z, acc.y, acc.x) : acc;
          if (single_cta_group) *q4 = x;
          else asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(q4), "f"(x.x), "f"(x.y), "f"(x.z), "f"(x.w) : "memory");
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (l + e < L0) {
              float* q1 = dst + (rev ? L0 - 1 - (l + e) : l + e);
              if (single_cta_group) *q1 = f4_at(acc, e);
              else atomicAdd(q1, f4_at(acc, e));
            }
          }
        }
      }
      __syncthreads();
    }
  }
  // ---- dA, dD, d(delta_bias): sum over the threads of the row ----
  for (int off = gm.seg >> 1; off >= 1; off >>= 1) {
    accA += __shfl_xor_sync(0xffffffffu, accA, off);
    accD += __shfl_xor_sync(0xffffffffu, accD, off);
    accb += __shfl_xor_sync(0xffffffffu, accb, off);
  }
  if (gm.wpr > 1) {
    const int warp = threadIdx.x >> 5;
    if (lane == 0) { s_red[warp][0] = accA; s_red[warp][1] = accD; s_red[warp][2] = accb; }
    __syncthreads();
    if (q.t == 0) {
      accA = accD = accb = 0.f;
      for (int w = 0; w < gm.wpr; ++w) { accA += s_red[q.r * gm.wpr + w][0]; accD += s_red[q.r * gm.wpr + w][1]; accb += s_red[q.r * gm.wpr + w][2]; }
    }
  }
  if (q.t == 0 && valid) {
    if (grads_zeroed) {
      atomicAdd(dA + d, accA);
      if (dD) atomicAdd(dD + d, accD);
      if (dbias) atomicAdd(dbias + d, accb);
    } else {
      float* dst = p.part + ((int64_t)b * p.dim + d) * 3;
      dst[0] = accA; dst[1] = accD; dst[2] = accb;
    }
  }
}

// ------------------------------------------------------------------------------------------------- host side
static bool n1_aligned16(const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; }

// fp32 everywhere, d_state 1 in a single pass, contiguous traversal for every group
static bool n1_eligible(const ScanParams& p, bool backward) {
  if (p.N != 1 || p.A_ld != 1 || p.accum || p.io_dtype != SS2D_F32) return false;
  if ((p.out != nullptr || backward) && p.out_dtype != SS2D_F32) return false;
  if (p.u_mod > 0 && p.dim % p.u_mod != 0) return false;
  for (int g = 0; g < p.G; ++g) {
    const int dir = p.layout == SS2D_LAYOUT_NATURAL ? p.dirs[g] : 0;
    if (dir == 2 || dir == 4) return false;
  }
  return true;
}

static bool n1_vec(const ScanParams& p, bool backward) {
  if (p.L & 3) return false;
  const int64_t strides[] = {p.u_bs, p.u_ds, p.dl_bs, p.dl_ds, p.B_bs, p.B_gs, p.C_bs, p.C_gs};
  for (int64_t s : strides) if (s & 3) return false;
  if (!n1_aligned16(p.u) || !n1_aligned16(p.delta) || !n1_aligned16(p.Bm) || !n1_aligned16(p.Cm)) return false;
  if (!backward) return p.out == nullptr || (n1_aligned16(p.out) && !(p.out_bs & 3) && !(p.out_ds & 3));
  return n1_aligned16(p.dout) && !(p.out_bs & 3) && !(p.out_ds & 3) && n1_aligned16(p.du) && n1_aligned16(p.ddelta) &&
         n1_aligned16(p.dB) && n1_aligned16(p.dC);
}

static N1Geom n1_geometry(const ScanParams& p, bool backward) {
  N1Geom gm;
  const int L = p.L;
  if (L <= 32 * N1_P) {                       // one warp (or a fraction of it) per row
    gm.seg = 4;
    while (gm.seg < 32 && gm.seg * N1_P < L) gm.seg <<= 1;
    gm.wpr = 1;
  } else {                                    // whole warps; a row of up to 16 warps x 256 positions is a single pass
    gm.seg = 32;
    const int need = (L + 32 * N1_P - 1) / (32 * N1_P);
    gm.wpr = need <= N1_MAX_WARPS ? need : 8;
  }
  gm.lc = gm.seg * gm.wpr * N1_P;
  gm.nchunks = (L + gm.lc - 1) / gm.lc;
  // rows per CTA: fill about 8 warps, but keep at least ~2 CTAs per SM in the grid when the problem is that small
  const int rows_per_warp = gm.seg == 32 ? 1 : 32 / gm.seg;
  int warps = gm.wpr > 8 ? gm.wpr : 8;
  int rt = gm.seg == 32 ? warps / gm.wpr : warps * rows_per_warp;
  const int sms = sm_count_current_device();
  while (rt > rows_per_warp && rt > 1 && (long)((p.dpg + rt - 1) / rt) * p.G * p.batch < 2L * sms) rt = (rt + 1) / 2;
  if (gm.seg < 32) rt = (rt + rows_per_warp - 1) / rows_per_warp * rows_per_warp;     // whole warps
  if (rt > p.dpg) rt = gm.seg == 32 ? p.dpg : (p.dpg + rows_per_warp - 1) / rows_per_warp * rows_per_warp;
  gm.rt = rt < 1 ? 1 : rt;
  (void)backward;
  return gm;
}

static int n1_threads(const N1Geom& gm) {
  return gm.seg == 32 ? gm.rt * gm.wpr * 32 : gm.rt / (32 / gm.seg) * 32;
}

// Returns true when the lean path took the call (*err holds the launch status).
bool scan_n1_fwd_try(const ScanParams& p, cudaStream_t stream, cudaError_t* err) {
  if (!n1_eligible(p, false)) return false;
  const N1Geom gm = n1_geometry(p, false);
  dim3 grid((p.dpg + gm.rt - 1) / gm.rt, p.G, p.batch);
  if (n1_vec(p, false)) scan_n1_fwd_kernel<true><<<grid, n1_threads(gm), 0, stream>>>(p, gm);
  else scan_n1_fwd_kernel<false><<<grid, n1_threads(gm), 0, stream>>>(p, gm);
  *err = cudaGetLastError();
  return true;
}

bool scan_n1_bwd_try(const ScanParams& p, float* dA, float* dD, float* dbias, int grads_zeroed, cudaStream_t stream,
                     cudaError_t* err) {
  if (!n1_eligible(p, true)) return false;
  const N1Geom gm = n1_geometry(p, true);
  dim3 grid((p.dpg + gm.rt - 1) / gm.rt, p.G, p.batch);
  const size_t smem = gm.rt > 1 ? (size_t)2 * gm.rt * gm.lc * sizeof(float) : 0;
  if (smem > 40 * 1024) return false;
  if (n1_vec(p, true)) scan_n1_bwd_kernel<true><<<grid, n1_threads(gm), smem, stream>>>(p, gm, dA, dD, dbias, grads_zeroed);
  else scan_n1_bwd_kernel<false><<<grid, n1_threads(gm), smem, stream>>>(p, gm, dA, dD, dbias, grads_zeroed);
  *err = cudaGetLastError();
  return true;
}

}  // namespace ss2d
