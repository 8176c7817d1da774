// Selective-scan forward, fast path for the north-star regime (8 < d_state <= 16, fp32, TMA-stageable operands,
// contiguous traversal: SCAN layout or directions 1 / 3). Everything else runs scan_fwd.cu.
//
// Replaces selective_scan_fwd_kernel (/root/reference/gm-unet/kernels/selective_scan/csrc/selective_scan/cus/
// selective_scan_fwd_kernel.cuh:61-172). Design:
//   * a warp owns 16 channel rows = 8 ROW PAIRS; lane = (row pair, state quad): 4 states x 2 rows per lane. All the
//     arithmetic of a row pair is packed f32x2 (FMUL2 / FFMA2, one half per row) — including the recurrence itself,
//     which the position-packed version of scan_fwd.cu has to leave scalar. B/C enter as broadcast scalar operands.
//   * no producer warp and no CTA-wide barrier: every warp streams its own delta / u rows through a private 3-stage
//     TMA ring (lane 0 refills a stage as soon as the warp has finished with it); the B/C tiles are shared by the CTA
//     and refilled by whichever warp releases them last (shared-memory arrival counter).
//   * per tile the warp rewrites its delta / u tiles in place as activated, ROW-PAIR-INTERLEAVED arrays
//     (delta', delta'*u as 64-bit pairs, XOR-swizzled) so that the hot loop loads ready-made register pairs.
//   * the 4 state lanes of a row pair combine their partial y through a 1 KB shared-memory exchange (128-bit reads
//     instead of shuffles); the finished y tile leaves through a TMA store.
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "scan_params.h"
#include "common.cuh"
#include "tma_host.h"

namespace ss2d {

constexpr int F2_STAGES = 3;
constexpr int F2_RPW = 16;                 // rows per warp
constexpr int F2_WSTAGE = 4096;            // delta tile (2 KB) + u tile (2 KB) of one warp
constexpr int F2_WARP_BYTES = F2_STAGES * F2_WSTAGE + 2048 /* u' */ + 2 * 2048 /* out */ + 2 * 1024 /* exchange */;
constexpr int F2_BC_STAGE = 4096;          // B tile (2 KB) + C tile (2 KB), 16 state rows each

struct Fwd2Maps { TMap u, dl, B, C, out; };

template <int NW>
struct Fwd2Shape {
  static constexpr int CH = NW * F2_RPW;
  static constexpr size_t smem_bytes = (size_t)F2_STAGES * F2_BC_STAGE + (size_t)NW * F2_WARP_BYTES + 1024 + 256;
};

__device__ __forceinline__ void tma_store_4d(const void* tmap, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(tmap),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ int atom_add_acqrel_shared(int* addr, int v) {
  int old;
  asm volatile("atom.acq_rel.cta.shared::cta.add.s32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(addr)), "r"(v) : "memory");
  return old;
}

// row-pair-interleaved array P of one warp: 8 row pairs x 32 positions x (row 0, row 1); 16-byte chunk `ch` (two
// positions) of row pair rp lives at rp * 256 + ((ch ^ rp) << 4): the 8 row pairs read at the same position hit 8
// different bank groups
__device__ __forceinline__ int p_off(int rp, int ch) { return rp * 256 + ((ch ^ rp) << 4); }
// TMA SWIZZLE_128B tile of 32-float rows: byte offset of the 16-byte chunk c4 of row r
__device__ __forceinline__ int t_off(int r, int c4) { return r * 128 + (((c4 ^ r) & 7) << 4); }

template <int NW>
__global__ void __launch_bounds__(NW * 32) scan_fwd2_kernel(const ScanParams p, const __grid_constant__ Fwd2Maps maps) {
  extern __shared__ __align__(16) unsigned char smem_raw2[];
  unsigned char* smem = smem_raw2 + ((1024 - (smem_u32(smem_raw2) & 1023)) & 1023);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* s_bc = smem;                                                        // [STAGES][B | C]
  unsigned char* s_w = smem + F2_STAGES * F2_BC_STAGE + warp * F2_WARP_BYTES;       // this warp's private region
  unsigned char* s_up = s_w + F2_STAGES * F2_WSTAGE;                                 // u' (row-pair interleaved)
  unsigned char* s_out = s_up + 2048;                                                // [2] y tiles, TMA layout
  unsigned char* s_ex = s_out + 2 * 2048;                                            // [2] exchange buffers
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + F2_STAGES * F2_BC_STAGE + NW * F2_WARP_BYTES);
  uint64_t* full_w = bars + warp * F2_STAGES;                                        // [NW][STAGES]
  uint64_t* full_bc = bars + NW * F2_STAGES;                                         // [STAGES]
  int* cnt_bc = reinterpret_cast<int*>(full_bc + F2_STAGES);                         // [STAGES]

  const int b = blockIdx.z, g = blockIdx.y;
  const int row0 = blockIdx.x * (NW * F2_RPW) + warp * F2_RPW;      // first row of this warp inside the group
  const int rows_valid = p.dpg - row0;                               // may be <= 0 for an idle warp
  const int d0 = g * p.dpg + row0;
  const int L = p.L;
  const int dir = p.layout == SS2D_LAYOUT_NATURAL ? p.dirs[g] : 0;
  const bool rev = dir == 3;
  const int ntiles = (L + 31) / 32;
  const int ug = p.u_mod > 0 ? g % (p.u_mod / p.dpg) : g;            // group coordinate of u

  if (threadIdx.x == 0) {
    for (int i = 0; i < NW * F2_STAGES; ++i) mbar_init(&bars[i], 1);
    for (int s = 0; s < F2_STAGES; ++s) { mbar_init(&full_bc[s], 1); cnt_bc[s] = 0; }
    fence_mbar_init();
  }
  __syncthreads();

  auto issue_rows = [&](int t) {     // lane 0 of the warp
    const int s = t % F2_STAGES;
    const int m0 = rev ? L - t * 32 - 32 : t * 32;
    mbar_arrive_expect_tx(&full_w[s], F2_WSTAGE);
    tma_load_4d(s_w + s * F2_WSTAGE, &maps.dl, m0, row0, g, b, &full_w[s]);
    tma_load_4d(s_w + s * F2_WSTAGE + 2048, &maps.u, m0, row0, ug, b, &full_w[s]);
  };
  auto issue_bc = [&](int t) {
    const int s = t % F2_STAGES;
    const int m0 = rev ? L - t * 32 - 32 : t * 32;
    mbar_arrive_expect_tx(&full_bc[s], F2_BC_STAGE);
    tma_load_4d(s_bc + s * F2_BC_STAGE, &maps.B, m0, 0, g, b, &full_bc[s]);
    tma_load_4d(s_bc + s * F2_BC_STAGE + 2048, &maps.C, m0, 0, g, b, &full_bc[s]);
  };
  if (lane == 0) {
    if (warp == 0) { tma_prefetch_desc(&maps.B); tma_prefetch_desc(&maps.C); }
    tma_prefetch_desc(&maps.dl); tma_prefetch_desc(&maps.u); tma_prefetch_desc(&maps.out);
    for (int t = 0; t < F2_STAGES && t < ntiles; ++t) {
      if (warp == 0) issue_bc(t);
      issue_rows(t);
    }
  }

  const int q = lane & 3, rl = lane >> 2;
  // main-loop role: row pair rl (rows 2rl, 2rl+1 of the warp), states n = 4 j + q
  float2 A2p[4], h2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = 4 * j + q;
    const bool okn = n < p.N;
    A2p[j].x = (okn && 2 * rl < rows_valid) ? p.A[(int64_t)(d0 + 2 * rl) * p.A_ld + n] * kLog2e : 0.f;
    A2p[j].y = (okn && 2 * rl + 1 < rows_valid) ? p.A[(int64_t)(d0 + 2 * rl + 1) * p.A_ld + n] * kLog2e : 0.f;
    h2[j] = make_float2(0.f, 0.f);
  }
  // exchange-reader / activation role: row pair rl, position e' = q of each group
  float2 Dp, biasp;
  Dp.x = (p.Dv && 2 * rl < rows_valid) ? p.Dv[d0 + 2 * rl] : 0.f;
  Dp.y = (p.Dv && 2 * rl + 1 < rows_valid) ? p.Dv[d0 + 2 * rl + 1] : 0.f;
  biasp.x = (p.bias && 2 * rl < rows_valid) ? p.bias[d0 + 2 * rl] : 0.f;
  biasp.y = (p.bias && 2 * rl + 1 < rows_valid) ? p.bias[d0 + 2 * rl + 1] : 0.f;
  const int ex_w = rl * 128 + q * 8;                     // + ((e ^ (rl & 3)) << 5)
  const int ex_r = rl * 128 + ((q ^ (rl & 3)) << 5);
  const bool softplus = p.softplus != 0;
  const bool want_out = p.out != nullptr;

  auto body = [&](auto REV) {
    constexpr bool REVV = decltype(REV)::value;
    int ob = 0;
    for (int t = 0; t < ntiles; ++t) {
      const int s = t % F2_STAGES;
      const uint32_t par = (t / F2_STAGES) & 1;
      const int len = min(32, L - t * 32);
      unsigned char* s_dl = s_w + s * F2_WSTAGE;          // delta tile -> delta' pairs
      unsigned char* s_du = s_dl + 2048;                  // u tile -> (delta' u) pairs
      mbar_wait(&full_w[s], par);
      // ---- activation + row-pair interleave (in place): this lane handles position quads q and q + 4 of row pair rl
      {
        float4 dv[2][2], uv[2][2];
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          const int c4 = q + 4 * it, tc4 = REVV ? 7 - c4 : c4;
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            float4 a = *reinterpret_cast<const float4*>(s_dl + t_off(2 * rl + r, tc4));
            float4 c = *reinterpret_cast<const float4*>(s_du + t_off(2 * rl + r, tc4));
            if (REVV) { a = make_float4(a.w, a.z, a.y, a.x); c = make_float4(c.w, c.z, c.y, c.x); }
            dv[it][r] = a; uv[it][r] = c;
          }
        }
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          const int c4 = q + 4 * it;
          float dl[2][4], du[2][4];
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float x = f4_at(dv[it][r], e) + (r == 0 ? biasp.x : biasp.y);
              if (softplus) x = softplus20(x);
              const bool live = c4 * 4 + e < len;     // beyond the end of the sequence the state is frozen (a = 1, b = 0)
              dl[r][e] = live ? x : 0.f;
              du[r][e] = live ? x * f4_at(uv[it][r], e) : 0.f;
            }
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {              // chunk 2 c4 + hh: positions 4 c4 + 2 hh, + 1
            const int off = p_off(rl, 2 * c4 + hh);
            *reinterpret_cast<float4*>(s_dl + off) = make_float4(dl[0][2 * hh], dl[1][2 * hh], dl[0][2 * hh + 1], dl[1][2 * hh + 1]);
            *reinterpret_cast<float4*>(s_du + off) = make_float4(du[0][2 * hh], du[1][2 * hh], du[0][2 * hh + 1], du[1][2 * hh + 1]);
            *reinterpret_cast<float4*>(s_up + off) =
                make_float4(f4_at(uv[it][0], 2 * hh), f4_at(uv[it][1], 2 * hh), f4_at(uv[it][0], 2 * hh + 1), f4_at(uv[it][1], 2 * hh + 1));
          }
        }
        __syncwarp();
      }
      mbar_wait(&full_bc[s], par);
      const unsigned char* s_B = s_bc + s * F2_BC_STAGE;
      unsigned char* o_tile = s_out + ob * 2048;

#pragma unroll 2
      for (int gi = 0; gi < 8; ++gi) {
        const int pofs = p_off(rl, 2 * gi);
        const float4 d01 = *reinterpret_cast<const float4*>(s_dl + pofs);
        const float4 d23 = *reinterpret_cast<const float4*>(s_dl + (pofs ^ 16));
        const float4 u01 = *reinterpret_cast<const float4*>(s_du + pofs);
        const float4 u23 = *reinterpret_cast<const float4*>(s_du + (pofs ^ 16));
        const float2 dl2[4] = {make_float2(d01.x, d01.y), make_float2(d01.z, d01.w), make_float2(d23.x, d23.y), make_float2(d23.z, d23.w)};
        const float2 du2[4] = {make_float2(u01.x, u01.y), make_float2(u01.z, u01.w), make_float2(u23.x, u23.y), make_float2(u23.z, u23.w)};
        float2 y2[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) y2[e] = make_float2(0.f, 0.f);
        const int tc4 = REVV ? 7 - gi : gi;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float4 Bq = *reinterpret_cast<const float4*>(s_B + t_off(4 * j + q, tc4));
          float4 Cq = *reinterpret_cast<const float4*>(s_B + 2048 + t_off(4 * j + q, tc4));
          if (REVV) { Bq = make_float4(Bq.w, Bq.z, Bq.y, Bq.x); Cq = make_float4(Cq.w, Cq.z, Cq.y, Cq.x); }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float be = f4_at(Bq, e), ce = f4_at(Cq, e);
            const float2 arg = __fmul2_rn(dl2[e], A2p[j]);
            const float2 a = make_float2(ex2f(arg.x), ex2f(arg.y));
            const float2 bu = __fmul2_rn(du2[e], make_float2(be, be));
            h2[j] = __ffma2_rn(a, h2[j], bu);
            y2[e] = __ffma2_rn(h2[j], make_float2(ce, ce), y2[e]);
          }
        }
        if (want_out) {
          unsigned char* ex = s_ex + (gi & 1) * 1024;
#pragma unroll
          for (int e = 0; e < 4; ++e) *reinterpret_cast<float2*>(ex + ex_w + ((e ^ (rl & 3)) << 5)) = y2[e];
          __syncwarp();
          // reader: position e' = q of this group, row pair rl: sum the 4 state lanes, add the D skip, park y in the out tile
          const float4 p01 = *reinterpret_cast<const float4*>(ex + ex_r);
          const float4 p23 = *reinterpret_cast<const float4*>(ex + ex_r + 16);
          const float2 up = *reinterpret_cast<const float2*>(s_up + p_off(rl, 2 * gi + (q >> 1)) + (q & 1) * 8);
          float2 ys = make_float2((p01.x + p01.z) + (p23.x + p23.z), (p01.y + p01.w) + (p23.y + p23.w));
          ys = __ffma2_rn(Dp, up, ys);
          const int col = REVV ? 31 - (4 * gi + q) : 4 * gi + q;
          *reinterpret_cast<float*>(o_tile + t_off(2 * rl, col >> 2) + (col & 3) * 4) = ys.x;
          *reinterpret_cast<float*>(o_tile + t_off(2 * rl + 1, col >> 2) + (col & 3) * 4) = ys.y;
        }
      }
      // state checkpoint at the end of the tile (= SS2D_CHUNK positions) for the backward's recompute: ckpt[b][d][tile][n]
      if (p.ckpt != nullptr) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          if (2 * rl + r < rows_valid) {
            float* dst = p.ckpt + (((int64_t)b * p.dim + d0 + 2 * rl + r) * p.nck + t) * p.N;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (4 * j + q < p.N) dst[4 * j + q] = r == 0 ? h2[j].x : h2[j].y;
          }
        }
      }
      // hand the y tile to TMA, release the stage: generic-proxy accesses are ordered before the async-proxy ones
      fence_proxy_async();
      __syncwarp();
      const int m0 = REVV ? L - t * 32 - 32 : t * 32;       // memory position of tile column 0
      if (want_out && m0 < 0) {
        // ragged tail of a reversed traversal: the box would start before the row. Plain vector stores instead.
        float* outp = static_cast<float*>(p.out) + (int64_t)b * p.out_bs;
        for (int i = lane; i < F2_RPW * 8; i += 32) {
          const int r = i >> 3, c4 = i & 7, m = m0 + 4 * c4;
          if (r < rows_valid && m >= 0)
            *reinterpret_cast<float4*>(outp + (int64_t)(d0 + r) * p.out_ds + m) = *reinterpret_cast<const float4*>(o_tile + t_off(r, c4));
        }
      }
      if (lane == 0) {
        if (want_out && rows_valid > 0 && m0 >= 0) {
          tma_store_4d(&maps.out, o_tile, m0, row0, g, b);
          tma_store_commit();
          tma_store_wait_read<1>();      // the store issued one tile ago has read its buffer: it may be rewritten
        }
        if (t + F2_STAGES < ntiles) issue_rows(t + F2_STAGES);
        if (atom_add_acqrel_shared(&cnt_bc[s], 1) == NW - 1) {      // last warp to release the B/C stage refills it
          cnt_bc[s] = 0;
          if (t + F2_STAGES < ntiles) issue_bc(t + F2_STAGES);
        }
      }
      __syncwarp();
      ob ^= 1;
    }
  };
  if (rev) body(std::true_type{}); else body(std::false_type{});

  if (p.last_state != nullptr) {
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = 4 * j + q;
        if (2 * rl + r < rows_valid && n < p.N) {
          const int64_t slot = ((int64_t)b * p.dim + d0 + 2 * rl + r) * p.A_ld + n;
          const float hv = r == 0 ? h2[j].x : h2[j].y;
          if (p.last_il) { p.last_state[2 * slot] = 0.f; p.last_state[2 * slot + 1] = hv; }
          else p.last_state[slot] = hv;
        }
      }
  }
  if (lane == 0) tma_store_wait_all();
}

static bool fwd2_maps(const ScanParams& p, Fwd2Maps* m) {
  const long long ugroups = p.u_mod > 0 ? p.u_mod / p.dpg : p.G;
  const long long du[4] = {p.L, p.dpg, ugroups, p.batch}, su[4] = {1, p.u_ds, (long long)p.dpg * p.u_ds, p.u_bs};
  const long long dd[4] = {p.L, p.dpg, p.G, p.batch}, sd[4] = {1, p.dl_ds, (long long)p.dpg * p.dl_ds, p.dl_bs};
  const long long so[4] = {1, p.out_ds, (long long)p.dpg * p.out_ds, p.out_bs};
  const long long d4[4] = {p.L, p.N, p.G, p.batch};
  const long long s4B[4] = {1, p.B_ns, p.B_gs, p.B_bs}, s4C[4] = {1, p.C_ns, p.C_gs, p.C_bs};
  bool ok = make_tmap(&m->u, p.u, 4, du, su, F2_RPW) && make_tmap(&m->dl, p.delta, 4, dd, sd, F2_RPW) &&
            make_tmap(&m->B, p.Bm, 4, d4, s4B, 16) && make_tmap(&m->C, p.Cm, 4, d4, s4C, 16);
  if (ok && p.out) ok = make_tmap(&m->out, p.out, 4, dd, so, F2_RPW);
  else if (ok) m->out = m->dl;
  return ok;
}

template <int NW>
static cudaError_t launch_fwd2(const ScanParams& p, const Fwd2Maps& maps, cudaStream_t stream) {
  using S = Fwd2Shape<NW>;
  auto kern = scan_fwd2_kernel<NW>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::smem_bytes);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  dim3 grid((p.dpg + S::CH - 1) / S::CH, p.G, p.batch);
  kern<<<grid, NW * 32, S::smem_bytes, stream>>>(p, maps);
  return cudaGetLastError();
}

// Returns true when the fast path took the call (*err holds the launch status).
bool scan_fwd2_try(const ScanParams& p, cudaStream_t stream, cudaError_t* err) {
  static const int enabled = getenv("SS2D_FWD_V2") ? atoi(getenv("SS2D_FWD_V2")) : 1;
  static const int nw_forced = getenv("SS2D_FWD2_NW") ? atoi(getenv("SS2D_FWD2_NW")) : 0;
  if (!enabled || !p.tma_ok || p.N <= 8 || p.N > 16 || p.accum || p.io_dtype != SS2D_F32) return false;
  if (p.u_mod > 0 && p.u_mod % p.dpg != 0) return false;
  if (p.out && (p.out_dtype != SS2D_F32 || (p.out_ds & 3) || (p.out_bs & 3) || (reinterpret_cast<uintptr_t>(p.out) & 15))) return false;
  for (int g = 0; g < p.G; ++g) {
    const int dir = p.layout == SS2D_LAYOUT_NATURAL ? p.dirs[g] : 0;
    if (dir == 2 || dir == 4) return false;
  }
  Fwd2Maps maps;
  if (!fwd2_maps(p, &maps)) return false;
  const int nw = nw_forced ? nw_forced : (p.dpg % 64 == 0 || p.dpg > 256 ? 4 : 2);
  *err = nw == 4 ? launch_fwd2<4>(p, maps, stream) : launch_fwd2<2>(p, maps, stream);
  return true;
}

}  // namespace ss2d
