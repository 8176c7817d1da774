// Host-side construction of the TMA tensor maps used by the scan kernels (driver entry point resolved at run time,
// so the library links against cudart only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "scan_params.h"

namespace ss2d {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// fp32 tensor with innermost extent dims[0] (stride 4 bytes) and `rank` dimensions; strides in ELEMENTS for dims 1..;
// box = (kTileL, box1, 1, ...) with 128-byte swizzle (or dense rows when swizzle = false). Returns false when the tensor cannot be described.
inline bool make_tmap(TMap* out, const void* base, int rank, const long long* dims, const long long* strides_elems,
                      int box1, bool swizzle = true) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t box[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = (cuuint64_t)dims[i];
    box[i] = 1;
    estr[i] = 1;
    if (i > 0) {
      gstr[i - 1] = (cuuint64_t)strides_elems[i] * 4;
      if (gstr[i - 1] == 0 || (gstr[i - 1] & 15)) return false;
    }
  }
  box[0] = 32;
  box[1] = (cuuint32_t)box1;
  if (box1 > 256) return false;
  static_assert(sizeof(TMap) == sizeof(CUtensorMap), "TMap must mirror CUtensorMap");
  CUresult r = fn(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank,
                  const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// Maps for the operands of one scan call. CH rows of u / delta / dout and NPB rows of B / C per box.
inline bool make_scan_maps(const ScanParams& p, int CH, int NPB, bool with_dy, TmaMaps* m) {
  const long long urows = p.u_mod > 0 ? p.u_mod : p.dim;
  const long long d3u[3] = {p.L, urows, p.batch}, s3u[3] = {1, p.u_ds, p.u_bs};
  const long long d3d[3] = {p.L, p.dim, p.batch}, s3d[3] = {1, p.dl_ds, p.dl_bs};
  const long long d4[4] = {p.L, p.N, p.G, p.batch};
  const long long s4B[4] = {1, p.B_ns, p.B_gs, p.B_bs}, s4C[4] = {1, p.C_ns, p.C_gs, p.C_bs};
  bool ok = make_tmap(&m->u, p.u, 3, d3u, s3u, CH) && make_tmap(&m->dl, p.delta, 3, d3d, s3d, CH) &&
            make_tmap(&m->B, p.Bm, 4, d4, s4B, NPB) && make_tmap(&m->C, p.Cm, 4, d4, s4C, NPB);
  if (ok && with_dy) {
    const long long d3y[3] = {p.L, urows, p.batch}, s3y[3] = {1, p.out_ds, p.out_bs};
    const long long d3yy[3] = {p.L, p.dim, p.batch};
    ok = make_tmap(&m->dy, p.dout, 3, p.u_mod > 0 ? d3y : d3yy, s3y, CH);
  }
  return ok;
}

}  // namespace ss2d
