// Diagnostics exported through the C ABI: a measured FP32-FMA peak for the Chamfer kernel's roofline.
//
// The Chamfer / NN kernels are bound by the FP32 pipe (bit-exact indices forbid the tensor-core form), and
// MEASURED_PEAKS.json carries no FP32 number, so bench.py measures one in the same run: every thread keeps 8 independent
// FMA chains in registers (enough ILP to cover the 4-cycle dependent-issue latency at 8+ resident warps per scheduler),
// scalar `fma.rn.f32` (FFMA, 2 flop per lane-instruction) or packed `fma.rn.f32x2` (FFMA2, 4 flop).  No reference
// counterpart.
#include "common.cuh"
#include "ldt_b200.h"

namespace ldt {

template <bool PACKED>
__global__ void __launch_bounds__(256) fma_peak_kernel(int iters, float seed, float* __restrict__ out) {
  float a = seed + threadIdx.x * 1e-3f, b = 0.999f;
  if constexpr (PACKED) {
    uint64_t acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = pack_f32x2(a + j, a - j);
    const uint64_t m = pack_f32x2(b, b), c = pack_f32x2(1e-3f, 2e-3f);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fma_f32x2(acc[j], m, c);
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float lo, hi;
      unpack_f32x2(acc[j], lo, hi);
      s += lo + hi;
    }
    if (s == 12345.678f) out[0] = s;   // keeps the chains alive; practically never true
  } else {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = a + j;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = __fmaf_rn(acc[j], b, 1e-3f);
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += acc[j];
    if (s == 12345.678f) out[0] = s;
  }
}

}  // namespace ldt

using namespace ldt;

extern "C" int ldt_debug_fma_peak(int iters, int packed, int blocks_per_sm, float* out, long long* flop, void* stream) {
  LDT_REQUIRE(iters > 0 && blocks_per_sm > 0 && out && flop, LDT_ERR_INVALID, "ldt_debug_fma_peak: bad arguments");
  const int grid = num_sms() * blocks_per_sm;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (packed) fma_peak_kernel<true><<<grid, 256, 0, s>>>(iters, 1.0f, out);
  else fma_peak_kernel<false><<<grid, 256, 0, s>>>(iters, 1.0f, out);
  LDT_CUDA_OK(cudaGetLastError());
  *flop = static_cast<long long>(grid) * 256LL * iters * 8LL * (packed ? 4 : 2);
  return LDT_OK;
}
