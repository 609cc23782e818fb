// Host-side TMA tensor-map construction shared by the tcgen05 kernels (defined in gemm.cu).
#pragma once
#include <cuda.h>

namespace ldt {
// bf16 [rows, ld] row-major matrix; box = box_rows x 64 columns (128 bytes), 128-byte swizzle, OOB reads give zero.
int make_tmap_bf16(CUtensorMap* tm, const void* base, int rows, int cols, int ld, int box_rows);
// f32 [rows, ld] row-major matrix; box = box_rows x 32 columns (128 bytes), 128-byte swizzle (kind::tf32 operands).
int make_tmap_f32(CUtensorMap* tm, const void* base, int rows, int cols, int ld, int box_rows);
}  // namespace ldt
