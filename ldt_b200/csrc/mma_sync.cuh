// Warp-level tensor-core helpers (mma.sync m16n8k16, bf16 -> fp32) shared by the attention kernels.
#pragma once
#include "common.cuh"

namespace ldt {

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
#ifdef LDT_QA_NO_HMMA   // A/B builds only: the warp-level MMAs replaced by one FMA each (operands stay live)
  d[0] = fmaf(__uint_as_float(a[0] ^ b0), 1e-30f, d[0]);
  d[1] = fmaf(__uint_as_float(a[1] ^ b1), 1e-30f, d[1]);
  d[2] = fmaf(__uint_as_float(a[2] ^ b0), 1e-30f, d[2]);
  d[3] = fmaf(__uint_as_float(a[3] ^ b1), 1e-30f, d[3]);
  return;
#endif
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t ld_u32(const __nv_bfloat16* p) { return *reinterpret_cast<const uint32_t*>(p); }

}  // namespace ldt
