// Host-side TMA tensor-map helpers shared by the tcgen05 kernels (gemm_tc.cu, attn_tc.cu, xattn_tc.cu).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace ia2p {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda); nullptr if unavailable.  api.cu
EncodeTiledFn tensor_map_encoder();

// bf16 3-D map {cols, tokens, batch} over rows of pitch `ld` elements, box {64, box_rows, 1}, SWIZZLE_128B: the operand /
// output tiles of the attention kernels (one head = 64 columns; rows past `tokens` are zero-filled on load, clipped on store)
int make_map_3d_bf16(CUtensorMap* m, const void* base, int64_t cols, int64_t ld, int64_t tokens, int64_t batch, int box_rows,
                     const char* what);

}  // namespace ia2p
