// Dense contraction core:  C[M,N] = epilogue(A[M,K] . W[N,K]^T + bias)   (bf16 in, fp32 accumulate)
//
// Product path (backend 0): persistent, warp-specialised tcgen05 kernel.
//   warp 0      TMA producer   : cp.async.bulk.tensor 2-D tiles (128B swizzle) into a STAGES-deep ring
//   warp 1      MMA issuer     : one elected thread issues tcgen05.mma (UMMA 128 x BN x 16), accumulators
//                                live in TMEM, double-buffered (2 x BN columns) so the epilogue of tile i
//                                overlaps the mainloop of tile i+1
//   warp 2      TMEM allocator
//   warps 4-11  epilogue       : tcgen05.ld 32x32b -> registers -> fused bias / GELU / gate*acc+residual
// Cross-check path (backend 1): a deliberately naive SIMT kernel sharing the same epilogue code; used by
// tests only.
//
// Replaces every nn.Conv1d(k=1)/nn.Linear of the reference hot path (model/layers.py:120-124,159-161,
// 172,238; model/scorenet/score.py:95; model/Compressor/Network.py:61,153), which the reference runs as
// separate cuDNN/cuBLAS launches followed by separate element-wise kernels.
#include <cuda.h>

#include "common.cuh"
#include "ldt_b200.h"

namespace ldt {

struct EpiParams {
  int M, N;
  const float* bias;
  void* out;
  int ldo;
  const float* resid;
  const float* gate;
  long long gate_stride;
  int rows_per_gate;
};

// One thread finishes 32 consecutive columns [col0, col0+32) of output row `row`.
template <int EPI>
__device__ __forceinline__ void epilogue_row32(const EpiParams& p, int row, int col0, const uint32_t (&acc)[32]) {
  if (row >= p.M || col0 >= p.N) return;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
  const bool full = (col0 + 32 <= p.N);
  if (p.bias != nullptr) {
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
        v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < p.N) v[j] += __ldg(p.bias + col0 + j);
    }
  }
  if constexpr (EPI == LDT_EPI_BIAS_GELU_BF16) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_erf_f(v[j]);
  }
  if constexpr (EPI == LDT_EPI_GATE_RESID_F32) {
    const float* res = p.resid + static_cast<size_t>(row) * p.ldo + col0;
    const float* g = p.gate ? p.gate + static_cast<long long>(row / p.rows_per_gate) * p.gate_stride + col0 : nullptr;
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 r4 = *reinterpret_cast<const float4*>(res + j);
        if (g) {
          const float4 g4 = __ldg(reinterpret_cast<const float4*>(g + j));
          v[j] = r4.x + g4.x * v[j]; v[j + 1] = r4.y + g4.y * v[j + 1];
          v[j + 2] = r4.z + g4.z * v[j + 2]; v[j + 3] = r4.w + g4.w * v[j + 3];
        } else {
          v[j] += r4.x; v[j + 1] += r4.y; v[j + 2] += r4.z; v[j + 3] += r4.w;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < p.N) v[j] = res[j] + (g ? g[j] : 1.0f) * v[j];
    }
  }
  if constexpr (EPI == LDT_EPI_BIAS_F32 || EPI == LDT_EPI_GATE_RESID_F32) {
    float* o = static_cast<float*>(p.out) + static_cast<size_t>(row) * p.ldo + col0;
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < p.N) o[j] = v[j];
    }
  } else {
    __nv_bfloat16* o = static_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(row) * p.ldo + col0;
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v[j], v[j + 1]);
        __nv_bfloat162 h1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]);
        __nv_bfloat162 h3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
        uint4 u;
        u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
        u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(o + j) = u;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < p.N) o[j] = __float2bfloat16_rn(v[j]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// tcgen05 kernel
// ------------------------------------------------------------------------------------------------
constexpr int TC_BM = 128;
constexpr int TC_BK = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int TC_THREADS = 384;
constexpr int TC_EPI_WARP0 = 4;

template <int BN>
struct TcCfg {
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;
  static constexpr int B_BYTES = BN * TC_BK * 2;
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int TMEM_COLS = 2 * BN;  // 512 or 256: powers of two
  static constexpr int SMEM_BYTES = 1024 /*align slack*/ + STAGES * (A_BYTES + B_BYTES) + 256 /*barriers*/;
};

template <int BN, int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const EpiParams p,
               const int K, const int tiles_m, const int tiles_n) {
  using Cfg = TcCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * Cfg::B_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = K / TC_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 8);  // one arrive per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile % tiles_m) * TC_BM;
        const int n0 = (tile / tiles_m) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1u);
          mbar_expect_tx(&full[stage], Cfg::A_BYTES + Cfg::B_BYTES);
          tma_load_2d(sA + stage * Cfg::A_BYTES, &tmA, &full[stage], kb * TC_BK, m0);
          tma_load_2d(sB + stage * Cfg::B_BYTES, &tmW, &full[stage], kb * TC_BK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(TC_BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * Cfg::A_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            umma_bf16_ss(tmem_d, umma_desc_k_sw128(a_addr + k * 32), umma_desc_k_sw128(b_addr + k * 32), idesc,
                         (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[stage]);  // smem slot reusable once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&tfull[acc]);  // accumulator complete
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp >= TC_EPI_WARP0) {
    const int quad = warp & 3;                         // TMEM lane quadrant this warp may access
    const int half = (warp - TC_EPI_WARP0) >> 2;       // which half of the BN columns
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile % tiles_m) * TC_BM;
      const int n0 = (tile / tiles_m) * BN;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const int row = m0 + quad * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < BN / 64; ++c) {
        const int col = half * (BN / 2) + c * 32;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * BN + col);
        uint32_t v[32];
        tmem_ld_32x32(taddr, v);
        tmem_ld_wait();
        epilogue_row32<EPI>(p, row, n0 + col, v);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// Naive SIMT cross-check kernel (tests only): one warp per (row, 32-column chunk).
// ------------------------------------------------------------------------------------------------
template <int EPI>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const __nv_bfloat16* __restrict__ A, int lda,
                                                      const __nv_bfloat16* __restrict__ W, int ldw, int K,
                                                      const EpiParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  const int col0 = blockIdx.y * 32;
  if (row >= p.M) return;
  float accv[32];
  // lane j computes column col0+j, then the row's 32 results are gathered into lane 0's registers
  float acc = 0.f;
  const int col = col0 + lane;
  if (col < p.N) {
    const __nv_bfloat16* a = A + static_cast<size_t>(row) * lda;
    const __nv_bfloat16* w = W + static_cast<size_t>(col) * ldw;
    for (int k = 0; k < K; ++k) acc = fmaf(__bfloat162float(a[k]), __bfloat162float(w[k]), acc);
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) accv[j] = __shfl_sync(0xffffffffu, acc, j);
  if (lane == 0) {
    uint32_t u[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) u[j] = __float_as_uint(accv[j]);
    epilogue_row32<EPI>(p, row, col0, u);
  }
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || ptr == nullptr) {
    set_last_error("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed");
    return nullptr;
  }
  fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  return fn;
}

// bf16 [rows, ld] row-major; tile box = box_rows x 64 columns, 128-byte swizzle, OOB reads give zero.
static int make_tmap_bf16(CUtensorMap* tm, const void* base, int rows, int cols, int ld, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return LDT_ERR_CUDA;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(TC_BK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed: CUresult %d (base=%p rows=%d cols=%d ld=%d box_rows=%d)", (int)r, base,
                   rows, cols, ld, box_rows);
    return LDT_ERR_CUDA;
  }
  return LDT_OK;
}

template <int BN, int EPI>
static int launch_tc(const ldt_gemm_args& a, const EpiParams& p, cudaStream_t s) {
  using Cfg = TcCfg<BN>;
  CUtensorMap tmA, tmW;
  int rc = make_tmap_bf16(&tmA, a.A, a.M, a.K, a.lda, TC_BM);
  if (rc) return rc;
  rc = make_tmap_bf16(&tmW, a.W, a.N, a.K, a.ldw, BN);
  if (rc) return rc;
  static bool attr_done = false;
  if (!attr_done) {
    LDT_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done = true;
  }
  const int tiles_m = (a.M + TC_BM - 1) / TC_BM;
  const int tiles_n = (a.N + BN - 1) / BN;
  const int grid = min(tiles_m * tiles_n, num_sms());
  gemm_tc_kernel<BN, EPI><<<grid, TC_THREADS, Cfg::SMEM_BYTES, s>>>(tmA, tmW, p, a.K, tiles_m, tiles_n);
  LDT_CUDA_OK(cudaGetLastError());
  return LDT_OK;
}

template <int EPI>
static int launch_any(const ldt_gemm_args& a, const EpiParams& p, cudaStream_t s) {
  if (a.backend == 1) {
    dim3 grid((a.M + 7) / 8, (a.N + 31) / 32);
    gemm_simt_kernel<EPI><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(a.A), a.lda,
                                               static_cast<const __nv_bfloat16*>(a.W), a.ldw, a.K, p);
    LDT_CUDA_OK(cudaGetLastError());
    return LDT_OK;
  }
  if (a.N % 256 == 0) return launch_tc<256, EPI>(a, p, s);
  return launch_tc<128, EPI>(a, p, s);
}

}  // namespace ldt

using namespace ldt;

extern "C" int ldt_gemm_bf16(const ldt_gemm_args* args, void* stream) {
  LDT_REQUIRE(args != nullptr, LDT_ERR_INVALID, "ldt_gemm_bf16: null args");
  const ldt_gemm_args& a = *args;
  LDT_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, LDT_ERR_INVALID, "ldt_gemm_bf16: bad shape M=%d N=%d K=%d", a.M, a.N, a.K);
  LDT_REQUIRE(a.K % TC_BK == 0, LDT_ERR_INVALID, "ldt_gemm_bf16: K=%d must be a multiple of %d (pad with zeros)", a.K, TC_BK);
  LDT_REQUIRE(a.N % 8 == 0, LDT_ERR_INVALID, "ldt_gemm_bf16: N=%d must be a multiple of 8", a.N);
  LDT_REQUIRE(a.lda >= a.K && a.ldw >= a.K && a.lda % 8 == 0 && a.ldw % 8 == 0, LDT_ERR_INVALID,
              "ldt_gemm_bf16: lda=%d ldw=%d must be >= K and multiples of 8", a.lda, a.ldw);
  LDT_REQUIRE(a.ldo >= a.N && a.ldo % 8 == 0, LDT_ERR_INVALID, "ldt_gemm_bf16: ldo=%d must be >= N and a multiple of 8", a.ldo);
  LDT_REQUIRE(a.A && a.W && a.out, LDT_ERR_INVALID, "ldt_gemm_bf16: null operand");
  LDT_REQUIRE((reinterpret_cast<uintptr_t>(a.A) | reinterpret_cast<uintptr_t>(a.W) | reinterpret_cast<uintptr_t>(a.out) |
               reinterpret_cast<uintptr_t>(a.bias) | reinterpret_cast<uintptr_t>(a.resid) |
               reinterpret_cast<uintptr_t>(a.gate)) % 16 == 0,
              LDT_ERR_INVALID, "ldt_gemm_bf16: operands must be 16-byte aligned");
  EpiParams p;
  p.M = a.M; p.N = a.N; p.bias = a.bias; p.out = a.out; p.ldo = a.ldo;
  p.resid = a.resid; p.gate = a.gate; p.gate_stride = a.gate_stride;
  p.rows_per_gate = a.rows_per_gate > 0 ? a.rows_per_gate : 1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (a.epilogue) {
    case LDT_EPI_BIAS_F32: return launch_any<LDT_EPI_BIAS_F32>(a, p, s);
    case LDT_EPI_BIAS_BF16: return launch_any<LDT_EPI_BIAS_BF16>(a, p, s);
    case LDT_EPI_BIAS_GELU_BF16: return launch_any<LDT_EPI_BIAS_GELU_BF16>(a, p, s);
    case LDT_EPI_GATE_RESID_F32:
      LDT_REQUIRE(a.resid != nullptr, LDT_ERR_INVALID, "ldt_gemm_bf16: residual epilogue needs resid");
      LDT_REQUIRE(a.gate == nullptr || a.gate_stride % 4 == 0, LDT_ERR_INVALID, "ldt_gemm_bf16: gate_stride must be a multiple of 4");
      return launch_any<LDT_EPI_GATE_RESID_F32>(a, p, s);
    default:
      set_last_error("ldt_gemm_bf16: unknown epilogue %d", a.epilogue);
      return LDT_ERR_INVALID;
  }
}
