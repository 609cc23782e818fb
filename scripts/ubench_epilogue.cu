// Micro-benchmark of the f32 GEMM epilogue's data paths on one SM (and on all SMs at once), phase by phase.
// A 128 x 256 f32 accumulator tile sits in TMEM (contents irrelevant); NW warps move it to global memory the way
// gemm_epilogue.cuh does: tcgen05.ld 32x32b.x32 -> 8 x STS.128 (XOR-swizzled) -> __syncwarp -> 8 x (LDS.128 [+ LDG.128
// residual] -> FFMA2 -> STG.128).  Template flags switch the phases on one by one; clock64 per CTA, averaged.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I ldt_b200/csrc scripts/ubench_epilogue.cu -o /tmp/ubench_epi
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

#include "common.cuh"

using namespace ldt;

constexpr int TILE_M = 128, TILE_N = 256, UNIT_BYTES = 4096;

// PH bit 0: STS   bit 1: LDS   bit 2: STG   bit 3: LDG residual (prefetched one unit ahead)   bit 4: no TMEM load
template <int PH, int NW>
__global__ void __launch_bounds__(NW * 32, 1)
epi_kernel(float* __restrict__ out, const float* __restrict__ resid, int ldo, int tiles, long long* __restrict__ clk) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int quad = warp & 3;
  const int part = warp >> 2, parts = NW / 4;             // column split between the warps of a quadrant
  const uint32_t stg = smem_u32(smem) + warp * UNIT_BYTES;
  const uint32_t st_row = lane * 128u;
  const int rr0 = lane >> 3, cc = lane & 7;
  constexpr int NU = TILE_N / 32;
  float* obase = out + static_cast<size_t>(blockIdx.x) * TILE_M * ldo;
  const float* rbase = resid + static_cast<size_t>(blockIdx.x) * TILE_M * ldo;
  const size_t off0 = static_cast<size_t>(quad * 32 + rr0) * ldo + cc * 4;
  const size_t pitch = static_cast<size_t>(4) * ldo;
  uint32_t sink = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int t = 0; t < tiles; ++t) {
    uint32_t v[32];
    float4 r4[2][8];
    const int u_begin = part * (NU / parts), u_end = u_begin + NU / parts;
    if constexpr (PH & 8) {
#pragma unroll
      for (int i = 0; i < 8; ++i) r4[0][i] = __ldcg(reinterpret_cast<const float4*>(rbase + off0 + u_begin * 32 + i * pitch));
    }
#pragma unroll
    for (int uu = 0; uu < NU / parts; ++uu) {
      const int u = u_begin + uu;
      if constexpr (!(PH & 16)) {
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(u * 32), v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = lane + i + t;
      }
      if constexpr (PH & 8) {
        if (uu + 1 < NU / parts) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            r4[(uu + 1) & 1][i] = __ldcg(reinterpret_cast<const float4*>(rbase + off0 + (u + 1) * 32 + i * pitch));
        }
      }
      if constexpr (PH & 1) {
#pragma unroll
        for (int c = 0; c < 8; ++c)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + st_row + static_cast<uint32_t>((c ^ (lane & 7)) << 4)),
                       "r"(v[4 * c]), "r"(v[4 * c + 1]), "r"(v[4 * c + 2]), "r"(v[4 * c + 3]) : "memory");
        __syncwarp();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) sink ^= v[i];
      }
      if constexpr (PH & 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = rr0 + 4 * i;
          float4 a4;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a4.x), "=f"(a4.y), "=f"(a4.z), "=f"(a4.w)
                       : "r"(stg + static_cast<uint32_t>(rr * 128 + ((cc ^ (rr & 7)) << 4))));
          if constexpr (PH & 8) {
            const float4 r = r4[uu & 1][i];
            a4.x = fmaf(a4.x, 1.5f, r.x); a4.y = fmaf(a4.y, 1.5f, r.y); a4.z = fmaf(a4.z, 1.5f, r.z); a4.w = fmaf(a4.w, 1.5f, r.w);
          }
          if constexpr (PH & 4) *reinterpret_cast<float4*>(obase + off0 + u * 32 + i * pitch) = a4;
          else sink ^= __float_as_uint(a4.x) ^ __float_as_uint(a4.y) ^ __float_as_uint(a4.z) ^ __float_as_uint(a4.w);
        }
        __syncwarp();
      }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
  if (sink == 0x12345678u) out[threadIdx.x] = 1.f;   // keeps the loads alive
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}


// The TMA form: residual box (32 rows x 128 B, 128-byte swizzle) loaded by TMA into one of the warp's two 4 KB buffers,
// each lane reads ITS row of the residual (8 x LDS.128), combines it with its accumulator row, writes the result back in
// place (8 x STS.128) and one lane bulk-stores the buffer.  LSU traffic: 128 KB of LDS + 128 KB of STS per tile; the
// L2 traffic is asynchronous.  MODE 0: wait_group.read 0 after every store, then prefetch unit n+2;  MODE 1: store only
// (no residual).
template <int MODE, int NW>
__global__ void __launch_bounds__(NW * 32, 1)
epi_tma_kernel(const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmR, int tiles,
               long long* __restrict__ clk) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bars[NW * 2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < NW * 2; ++i) mbar_init(&bars[i], 1);
    mbar_fence_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int quad = warp & 3;
  const int part = warp >> 2, parts = NW / 4;
  constexpr int NU = TILE_N / 32, NUW = NU / (NW / 4);
  const uint32_t buf0 = smem_u32(smem) + warp * 2 * UNIT_BYTES;
  const uint32_t bar0 = smem_u32(&bars[warp * 2]);
  const uint32_t my = lane * 128u;
  const int row0 = blockIdx.x * TILE_M + quad * 32;
  const int total = tiles * NUW;
  auto issue_load = [&](int g) {   // lane 0
    if (g >= total) return;
    const int u = part * NUW + g % NUW;
    const uint32_t b = g & 1;
    mbar_expect_tx_u32(bar0 + 8 * b, UNIT_BYTES);
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(buf0 + b * UNIT_BYTES), "l"(&tmR), "r"(u * 32), "r"(row0), "r"(bar0 + 8 * b) : "memory");
  };
  __syncthreads();
  const long long t0 = clock64();
  if (MODE == 0 && lane == 0) { issue_load(0); issue_load(1); }
  for (int g = 0; g < total; ++g) {
    const int u = part * NUW + g % NUW;
    const uint32_t b = g & 1, ph = (g >> 1) & 1;
    uint32_t v[32];
    tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(u * 32), v);
    if (MODE == 0) mbar_wait_u32(bar0 + 8 * b, ph);
    else {
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // store g-2 has read this buffer
      __syncwarp();
    }
    tmem_ld_wait();
    const uint32_t base = buf0 + b * UNIT_BYTES + my;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint32_t a = base + static_cast<uint32_t>((c ^ (lane & 7)) << 4);
      float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
      if (MODE == 0) asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(a));
      float4 o;
      o.x = fmaf(__uint_as_float(v[4 * c]), 1.5f, r.x); o.y = fmaf(__uint_as_float(v[4 * c + 1]), 1.5f, r.y);
      o.z = fmaf(__uint_as_float(v[4 * c + 2]), 1.5f, r.z); o.w = fmaf(__uint_as_float(v[4 * c + 3]), 1.5f, r.w);
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(&tmO, buf0 + b * UNIT_BYTES, u * 32, row0);
      tma_store_commit();
      if (MODE == 0) {
        tma_store_wait_read();
        issue_load(g + 2);
      }
    }
    __syncwarp();
  }
  if (lane == 0) tma_store_wait_read();
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

#include <cuda.h>
typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                            const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static CUtensorMap tmap_f32(void* base, int rows, int cols) {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  CUtensorMap tm;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)cols * 4};
  cuuint32_t box[2] = {32, 32}, estr[2] = {1, 1};
  CUresult r = ((PFN_enc)fp)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  return tm;
}

template <int MODE, int NW>
static void run_tma(const char* name, int ctas, float* out, float* resid, long long* clk, int tiles) {
  const int smem = NW * 2 * UNIT_BYTES + 1024;
  cudaFuncSetAttribute(epi_tma_kernel<MODE, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  CUtensorMap tmO = tmap_f32(out, 148 * TILE_M, TILE_N), tmR = tmap_f32(resid, 148 * TILE_M, TILE_N);
  for (int rep = 0; rep < 2; ++rep) epi_tma_kernel<MODE, NW><<<ctas, NW * 32, smem>>>(tmO, tmR, tiles, clk);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
  long long* h = static_cast<long long*>(malloc(sizeof(long long) * ctas));
  cudaMemcpy(h, clk, sizeof(long long) * ctas, cudaMemcpyDeviceToHost);
  double s = 0;
  for (int i = 0; i < ctas; ++i) s += static_cast<double>(h[i]);
  printf("%-44s NW=%d ctas=%3d: %8.0f clk / tile\n", name, NW, ctas, s / ctas / tiles);
  free(h);
}


// No shared-memory transpose: tcgen05.ld.16x256b.x8 hands lane 4i+j the columns {8k+2j, 8k+2j+1}, k = 0..7, of rows i and
// i+8, so the four lanes of a quad cover one 32-byte sector of a row and an LDG.64 / STG.64 of the warp touches 8 rows x 32
// bytes: full sectors, but eight 128-byte lines per instruction.  MODE bit 0: STG, bit 1: LDG residual.
template <int MODE, int NW>
__global__ void __launch_bounds__(NW * 32, 1)
epi_direct_kernel(float* __restrict__ out, const float* __restrict__ resid, int ldo, int tiles, long long* __restrict__ clk) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int quad = warp & 3, part = warp >> 2;
  constexpr int NU64 = TILE_N / 64, NUW = NU64 / (NW / 4);   // 64-column units per warp
  const int i8 = lane >> 2, j = lane & 3;
  float* obase = out + static_cast<size_t>(blockIdx.x) * TILE_M * ldo;
  const float* rbase = resid + static_cast<size_t>(blockIdx.x) * TILE_M * ldo;
  uint32_t sink = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int t = 0; t < tiles; ++t) {
#pragma unroll 1
    for (int uu = 0; uu < NUW; ++uu) {
      const int col0 = (part * NUW + uu) * 64;
#pragma unroll
      for (int h = 0; h < 2; ++h) {   // lanes 0-15 / 16-31 of the quadrant
        uint32_t v[32];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32 + h * 16) << 16) + static_cast<uint32_t>(col0);
        asm volatile(
            "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr) : "memory");
        float2 r2[16];
        const size_t off = static_cast<size_t>(quad * 32 + h * 16 + i8) * ldo + col0 + 2 * j;
        if constexpr (MODE & 2) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            r2[2 * k] = __ldcg(reinterpret_cast<const float2*>(rbase + off + 8 * k));
            r2[2 * k + 1] = __ldcg(reinterpret_cast<const float2*>(rbase + off + 8 * static_cast<size_t>(ldo) + 8 * k));
          }
        }
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {   // rows i and i + 8
            float2 o = make_float2(__uint_as_float(v[4 * k + 2 * hh]), __uint_as_float(v[4 * k + 2 * hh + 1]));
            if constexpr (MODE & 2) { o.x = fmaf(o.x, 1.5f, r2[2 * k + hh].x); o.y = fmaf(o.y, 1.5f, r2[2 * k + hh].y); }
            if constexpr (MODE & 1) *reinterpret_cast<float2*>(obase + off + hh * 8 * static_cast<size_t>(ldo) + 8 * k) = o;
            else sink ^= __float_as_uint(o.x) ^ __float_as_uint(o.y);
          }
        }
      }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
  if (sink == 0x12345678u) out[threadIdx.x] = 1.f;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

template <int MODE, int NW>
static void run_direct(const char* name, int ctas, float* out, float* resid, long long* clk, int tiles) {
  for (int rep = 0; rep < 2; ++rep) epi_direct_kernel<MODE, NW><<<ctas, NW * 32>>>(out, resid, TILE_N, tiles, clk);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
  long long* h = static_cast<long long*>(malloc(sizeof(long long) * ctas));
  cudaMemcpy(h, clk, sizeof(long long) * ctas, cudaMemcpyDeviceToHost);
  double s = 0;
  for (int i = 0; i < ctas; ++i) s += static_cast<double>(h[i]);
  printf("%-44s NW=%d ctas=%3d: %8.0f clk / tile\n", name, NW, ctas, s / ctas / tiles);
  free(h);
}

template <int PH, int NW>
static void run(const char* name, int ctas, float* out, float* resid, long long* clk, int tiles) {
  const int smem = NW * UNIT_BYTES + 1024;
  cudaFuncSetAttribute(epi_kernel<PH, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  // a second 100 KB of dynamic shared memory would let two CTAs share an SM; one CTA per SM is what the GEMM has
  for (int rep = 0; rep < 2; ++rep) epi_kernel<PH, NW><<<ctas, NW * 32, smem>>>(out, resid, TILE_N, tiles, clk);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
  long long* h = static_cast<long long*>(malloc(sizeof(long long) * ctas));
  cudaMemcpy(h, clk, sizeof(long long) * ctas, cudaMemcpyDeviceToHost);
  double s = 0;
  for (int i = 0; i < ctas; ++i) s += static_cast<double>(h[i]);
  printf("%-44s NW=%d ctas=%3d: %8.0f clk / tile\n", name, NW, ctas, s / ctas / tiles);
  free(h);
}

int main() {
  const int max_ctas = 148, tiles = 16;
  float *out, *resid;
  long long* clk;
  const size_t bytes = static_cast<size_t>(max_ctas) * TILE_M * TILE_N * sizeof(float);
  cudaMalloc(&out, bytes);
  cudaMalloc(&resid, bytes);
  cudaMemset(resid, 0, bytes);
  cudaMalloc(&clk, sizeof(long long) * max_ctas);
  for (int ctas : {1, 148}) {
    run<0, 8>("tcgen05.ld only", ctas, out, resid, clk, tiles);
    run<1, 8>("tcgen05.ld + STS", ctas, out, resid, clk, tiles);
    run<3, 8>("tcgen05.ld + STS + LDS", ctas, out, resid, clk, tiles);
    run<7, 8>("tcgen05.ld + STS + LDS + STG", ctas, out, resid, clk, tiles);
    run<15, 8>("tcgen05.ld + STS + LDS + LDG + STG", ctas, out, resid, clk, tiles);
    run<16 + 15, 8>("(no TMEM) STS + LDS + LDG + STG", ctas, out, resid, clk, tiles);
    run<16 + 7, 8>("(no TMEM) STS + LDS + STG", ctas, out, resid, clk, tiles);
    run<16 + 3, 8>("(no TMEM) STS + LDS", ctas, out, resid, clk, tiles);
    run_tma<0, 8>("TMA: resid load + LDS + STS + bulk store", ctas, out, resid, clk, tiles);
    run_tma<1, 8>("TMA: STS + bulk store (no residual)", ctas, out, resid, clk, tiles);
    run_tma<0, 4>("TMA: resid load + LDS + STS + bulk store", ctas, out, resid, clk, tiles);
    run_direct<0, 8>("direct 16x256b: tcgen05.ld only", ctas, out, resid, clk, tiles);
    run_direct<1, 8>("direct 16x256b: + STG.64", ctas, out, resid, clk, tiles);
    run_direct<3, 8>("direct 16x256b: + LDG.64 + STG.64", ctas, out, resid, clk, tiles);
    run_direct<3, 16>("direct 16x256b: + LDG.64 + STG.64", ctas, out, resid, clk, tiles);
    run<0, 4>("tcgen05.ld only", ctas, out, resid, clk, tiles);
    run<15, 4>("tcgen05.ld + STS + LDS + LDG + STG", ctas, out, resid, clk, tiles);
    run<0, 16>("tcgen05.ld only", ctas, out, resid, clk, tiles);
    run<15, 16>("tcgen05.ld + STS + LDS + LDG + STG", ctas, out, resid, clk, tiles);
  }
  return 0;
}
