/* ORACLE (test infrastructure only - never linked into or called by the product path).
 *
 * Plain-C restatement of the reference selective scan (forward + hand-derived backward) for
 * fp32 tensors, sequential over L per (batch, channel) row, OpenMP across rows.
 * Follows the semantics of
 *   /root/reference/gm-unet/kernels/selective_scan/test_selective_scan.py:183-234  (forward definition)
 *   /root/reference/gm-unet/kernels/selective_scan/csrc/selective_scan/cus/selective_scan_bwd_kernel.cuh:125-272
 *   (backward equations), re-derived, not transcribed.
 * Accumulation type is `acc_t` (double by default; -DORACLE_ACC_FLOAT for the fp32 CPU baseline).
 * Parity status: PINNED through tests/test_oracle_golden.py against tests/golden/scan_*.npz, which
 * were produced by the reference's own selective_scan_ref + autograd.
 *
 * Layouts (all contiguous): u, delta, out, dout, du, ddelta: (B, Dt, L); A, dA: (Dt, N);
 * Bm, Cm, dB, dC: (B, G, N, L); Dv, bias, dD, dbias: (Dt) or NULL.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#if defined(ORACLE_ACC_FLOAT) && defined(__SSE2__)
#include <xmmintrin.h>
#include <pmmintrin.h>
/* decaying fp32 states underflow into denormals, which x86 handles ~100x slower: flush them (the GPU does too) */
#define ORACLE_FTZ() do { _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON); _MM_SET_DENORMALS_ZERO_MODE(_MM_DENORMALS_ZERO_ON); } while (0)
#else
#define ORACLE_FTZ() do { } while (0)
#endif

#ifdef ORACLE_ACC_FLOAT
typedef float acc_t;
#define EXPF expf
#define LOG1PF log1pf
#else
typedef double acc_t;
#define EXPF exp
#define LOG1PF log1p
#endif

static inline acc_t softplus_thr20(acc_t x) { return x > (acc_t)20 ? x : LOG1PF(EXPF(x)); }

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* forward: out (B,Dt,L) and last_state (B,Dt,N) (may be NULL) */
int oracle_scan_fwd(const float* u, const float* delta, const float* A, const float* Bm, const float* Cm,
                    const float* Dv, const float* bias, int softplus, int nb, int nd, int L, int N, int G,
                    float* out, float* last_state) {
  if (nd % G) return -1;
  const int dpg = nd / G;
#pragma omp parallel
  {
    ORACLE_FTZ();
    acc_t* h = (acc_t*)malloc(sizeof(acc_t) * (size_t)N);
#pragma omp for schedule(static)
    for (long row = 0; row < (long)nb * nd; ++row) {
      const int b = (int)(row / nd), d = (int)(row % nd), g = d / dpg;
      const float* ur = u + row * (long)L;
      const float* dr = delta + row * (long)L;
      const float* Bg = Bm + ((long)b * G + g) * (long)N * L;
      const float* Cg = Cm + ((long)b * G + g) * (long)N * L;
      for (int n = 0; n < N; ++n) h[n] = 0;
      for (int l = 0; l < L; ++l) {
        acc_t raw = (acc_t)dr[l] + (bias ? (acc_t)bias[d] : 0);
        acc_t dt = softplus ? softplus_thr20(raw) : raw;
        acc_t dtu = dt * (acc_t)ur[l];
        acc_t y = Dv ? (acc_t)Dv[d] * (acc_t)ur[l] : 0;
        for (int n = 0; n < N; ++n) {
          acc_t a = EXPF(dt * (acc_t)A[(long)d * N + n]);
          h[n] = a * h[n] + dtu * (acc_t)Bg[(long)n * L + l];
          y += h[n] * (acc_t)Cg[(long)n * L + l];
        }
        out[row * (long)L + l] = (float)y;
      }
      if (last_state) for (int n = 0; n < N; ++n) last_state[row * (long)N + n] = (float)h[n];
    }
    free(h);
  }
  return 0;
}

/* backward: recomputes h in a forward sweep per row, then the reverse adjoint sweep.
 * dB/dC are reduced over the channels of a group deterministically (channel order). */
int oracle_scan_bwd(const float* u, const float* delta, const float* A, const float* Bm, const float* Cm,
                    const float* Dv, const float* bias, const float* dout, int softplus,
                    int nb, int nd, int L, int N, int G,
                    float* du, float* ddelta, float* dA, float* dB, float* dC, float* dD, float* dbias) {
  if (nd % G) return -1;
  const int dpg = nd / G;
  const long NL = (long)N * L;
  /* per-(b,g) units own dB/dC slices; dA/dD/dbias are accumulated per batch then summed over b */
  acc_t* dA_b = (acc_t*)calloc((size_t)nb * nd * N, sizeof(acc_t));
  acc_t* dD_b = (acc_t*)calloc((size_t)nb * nd, sizeof(acc_t));
  acc_t* db_b = (acc_t*)calloc((size_t)nb * nd, sizeof(acc_t));
  if (!dA_b || !dD_b || !db_b) return -2;
#pragma omp parallel
  {
    ORACLE_FTZ();
    acc_t* hs = (acc_t*)malloc(sizeof(acc_t) * (size_t)NL);
    acc_t* gB = (acc_t*)malloc(sizeof(acc_t) * (size_t)NL);
    acc_t* gC = (acc_t*)malloc(sizeof(acc_t) * (size_t)NL);
    acc_t* gv = (acc_t*)malloc(sizeof(acc_t) * (size_t)N * 2);
#pragma omp for schedule(dynamic, 1)
    for (long unit = 0; unit < (long)nb * G; ++unit) {
      const int b = (int)(unit / G), g = (int)(unit % G);
      const float* Bg = Bm + unit * NL;
      const float* Cg = Cm + unit * NL;
      for (long i = 0; i < NL; ++i) { gB[i] = 0; gC[i] = 0; }
      for (int dd = 0; dd < dpg; ++dd) {
        const int d = g * dpg + dd;
        const long row = (long)b * nd + d;
        const float* ur = u + row * (long)L;
        const float* dr = delta + row * (long)L;
        const float* dyr = dout + row * (long)L;
        acc_t* gn = gv; acc_t* an = gv + N;
        /* forward recompute of h */
        for (int n = 0; n < N; ++n) gn[n] = 0;
        for (int l = 0; l < L; ++l) {
          acc_t raw = (acc_t)dr[l] + (bias ? (acc_t)bias[d] : 0);
          acc_t dt = softplus ? softplus_thr20(raw) : raw;
          acc_t dtu = dt * (acc_t)ur[l];
          for (int n = 0; n < N; ++n) {
            acc_t a = EXPF(dt * (acc_t)A[(long)d * N + n]);
            gn[n] = a * gn[n] + dtu * (acc_t)Bg[(long)n * L + l];
            hs[(long)n * L + l] = gn[n];
          }
        }
        /* reverse adjoint sweep */
        for (int n = 0; n < N; ++n) { gn[n] = 0; an[n] = 0; }
        acc_t accD = 0, accb = 0;
        for (int l = L - 1; l >= 0; --l) {
          acc_t raw = (acc_t)dr[l] + (bias ? (acc_t)bias[d] : 0);
          acc_t dt = softplus ? softplus_thr20(raw) : raw;
          acc_t uu = (acc_t)ur[l], dy = (acc_t)dyr[l];
          acc_t dtu = dt * uu, sB = 0, sA = 0;
          for (int n = 0; n < N; ++n) {
            acc_t An = (acc_t)A[(long)d * N + n];
            acc_t a = EXPF(dt * An);
            acc_t gg = (acc_t)Cg[(long)n * L + l] * dy + an[n] * gn[n];
            acc_t hprev = l > 0 ? hs[(long)n * L + l - 1] : 0;
            gC[(long)n * L + l] += dy * hs[(long)n * L + l];
            gB[(long)n * L + l] += gg * dtu;
            sB += gg * (acc_t)Bg[(long)n * L + l];
            acc_t w = gg * a * hprev;
            sA += w * An;
            dA_b[row * (long)N + n] += w * dt;
            gn[n] = gg; an[n] = a;
          }
          acc_t ddt = uu * sB + sA;
          acc_t dd_raw = (softplus && raw <= (acc_t)20) ? ddt / ((acc_t)1 + EXPF(-raw)) : ddt;
          du[row * (long)L + l] = (float)((Dv ? (acc_t)Dv[d] * dy : 0) + dt * sB);
          ddelta[row * (long)L + l] = (float)dd_raw;
          accD += dy * uu; accb += dd_raw;
        }
        dD_b[row] = accD; db_b[row] = accb;
      }
      for (long i = 0; i < NL; ++i) { dB[unit * NL + i] = (float)gB[i]; dC[unit * NL + i] = (float)gC[i]; }
    }
    free(hs); free(gB); free(gC); free(gv);
  }
  for (int d = 0; d < nd; ++d) {
    acc_t sD = 0, sb = 0;
    for (int b = 0; b < nb; ++b) { sD += dD_b[(long)b * nd + d]; sb += db_b[(long)b * nd + d]; }
    if (dD && Dv) dD[d] = (float)sD;
    if (dbias && bias) dbias[d] = (float)sb;
    for (int n = 0; n < N; ++n) {
      acc_t s = 0;
      for (int b = 0; b < nb; ++b) s += dA_b[((long)b * nd + d) * N + n];
      dA[(long)d * N + n] = (float)s;
    }
  }
  free(dA_b); free(dD_b); free(db_b);
  return 0;
}
