// Shared pieces of the tcgen05 contraction kernels (gemm.cu, mlp.cu): tile constants, the CTA-pair pipeline
// configuration, and the fused epilogues (bias / exact-erf GELU / gate*acc + residual) with their staged, coalesced
// store path.  Device-side glue for model/layers.py:120-124,159-161,218-219 of the reference; see gemm.cu.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "ldt_b200.h"

namespace ldt {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int TC_THREADS = 384;
constexpr int TC_EPI_WARP0 = 4;
constexpr int EPI_STG_BYTES = 4096;  // epilogue staging buffer per warp (epilogue_staged)


constexpr int T2_BM = 256;

template <int BN>
struct Tc2Cfg {
  static constexpr int A_BYTES = 128 * TC_BK * 2;
  static constexpr int B_BYTES = (BN / 2) * TC_BK * 2;
#ifdef LDT_T2_STAGES   // A/B builds only
  static constexpr int STAGES = LDT_T2_STAGES;
#else
  static constexpr int STAGES = (BN == 256) ? 6 : 8;
#endif
  static constexpr int ACC_STRIDE = (BN > 128) ? 256 : 128;   // column offset between the two accumulators
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  static constexpr int SMEM_BYTES =
      1024 /*align slack*/ + STAGES * (A_BYTES + B_BYTES) + 256 /*barriers*/ + 8 * EPI_STG_BYTES /*epilogue staging*/;
};


struct EpiParams {
  int M, N;
  const float* bias;
  void* out;
  int ldo;
  const float* resid;
  const float* gate;
  long long gate_stride;
  int rows_per_gate;
  unsigned long long* dbg;  // optional per-CTA stall counters (ldt_debug_set_gemm_counters), else nullptr
  int tma_store;            // bf16 outputs leave through bulk tensor stores (0: per-lane st.global, kept for A/B and tests)
  int dbg_mode;             // experiments only (ldt_debug_set_gemm_mode): 1 skip A loads, 2 skip W loads, 4 skip the epilogue
  int f32_plain;            // LDT_EPI_BIAS_GELU_F32 only: 1 = do not round the output to TF32 (operand_type 2, the 3xTF32 mode)
  int relu;                 // f32 outputs only: 1 = max(., 0) last (LDT_EPI_BIAS_RELU_F32 / LDT_EPI_RESID_RELU_F32)
};

__device__ __forceinline__ float epi_relu(float x, int) { return fmaxf(x, 0.f); }

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
  if constexpr (EPI == LDT_EPI_BIAS_GELU_F32) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      v[j] = gelu_erf_f(v[j]);
      if (!p.f32_plain) v[j] = round_tf32(v[j]);
    }
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
  if constexpr (EPI == LDT_EPI_BIAS_F32 || EPI == LDT_EPI_GATE_RESID_F32 || EPI == LDT_EPI_BIAS_GELU_F32) {
    if constexpr (EPI != LDT_EPI_BIAS_GELU_F32) {
      if (p.relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = epi_relu(v[j], p.relu);
      }
    }
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
// Staged epilogue for the tcgen05 kernels.  tcgen05.ld 32x32b hands every lane ONE output row, so storing straight
// from registers makes each warp-wide store touch 32 different rows (16 B of every 32 B sector): measured, that
// epilogue took 11-15 k cycles per 128x256 tile against an 8 k-cycle mainloop and throttled the tensor pipe.  Here
// each warp transposes a 32-row x 128-byte unit through a private 4 KB shared-memory buffer (16-byte chunks XOR-
// swizzled by row, conflict-free both ways) so that 8 lanes cover one full 128-byte row segment: residual loads and
// output stores are fully coalesced (4 rows x 128 B per instruction).
//   fp32 outputs: unit = 32 columns;  bf16 outputs: unit = 64 columns (bias/GELU applied before the transpose).
// `taddr` is the TMEM address of (lane quadrant base, first column of this warp's slab); `ncols` columns are drained.
// ------------------------------------------------------------------------------------------------

// XM (diagnostics builds only, bf16 outputs; results are wrong): 8 = no TMEM loads, 16 = no shared-memory staging,
// 32 = no global stores, 64 = no GELU arithmetic -- compile-time so that the remaining code is what the product runs.
// tm_out != nullptr (bf16 outputs): the staged 32-row x 64-column unit is written by ONE bulk tensor store issued by lane 0
// (the staging layout IS the 128-byte TMA swizzle; the buffer must be 1024-byte aligned) instead of 8 ld.shared + 8
// st.global per lane: the per-SM store path through the LSU was costing the fc1 epilogue as much as its GELU arithmetic
// (scripts/exp_epi.py).  The caller must run tma_store_wait_all()/..._read() on lane 0 before the buffer or the CTA goes away.
#ifdef LDT_EPI_STAMPS   // variant builds only (scripts/exp_epi_stamps.py): clock64 stamps of the first epilogue warp of CTA 0
#define LDT_STAMP(k)                                                                                          \
  do {                                                                                                        \
    if (p.dbg != nullptr && blockIdx.x == 0 && (threadIdx.x >> 5) == 0 && lane == 0) p.dbg[64 + (k)] = clock64(); \
  } while (0)
#else
#define LDT_STAMP(k) do { } while (0)
#endif

template <int EPI, int NCOLS, int XM = 0, typename WaitFn>
__device__ __forceinline__ void epilogue_staged(const EpiParams& p, uint8_t* stg, int lane, int row_base, int col_base,
                                                uint32_t taddr, WaitFn wait_accumulator, const CUtensorMap* tm_out = nullptr) {
  constexpr int ncols = NCOLS;
  const uint32_t stg_u32 = smem_u32(stg);
  const uint32_t st_row = stg_u32 + static_cast<uint32_t>(lane) * 128u;   // staging row written by this lane
  const int rr0 = lane >> 3, cc = lane & 7;                                // read-back: row rr0 + 4*i, chunk cc
  if constexpr (EPI == LDT_EPI_BIAS_F32 || EPI == LDT_EPI_GATE_RESID_F32 || EPI == LDT_EPI_BIAS_GELU_F32) {
    constexpr int NU = NCOLS / 32;
    // one gate row for the whole 32-row slab (rows_per_gate a multiple of 32, e.g. the 32 latent tokens of a sample)?
    const bool gate_uniform = (p.rows_per_gate & 31) == 0 && (row_base & 31) == 0;
    // FAST (warp-uniform): the whole 32-row x NCOLS slab is inside the matrix and the gate is one row per slab -- every
    // tile of the score net except a ragged last one.  The per-round row / column predicates, the 64-bit address
    // products and the per-row gate branch then disappear from the 8 load-add-store rounds of a unit: the epilogue of a
    // tile is ~1000 warp-instructions of one dependent chain per warp (two such warps per scheduler), i.e. bound by
    // instruction latency, not by TMEM / shared-memory / L2 throughput (scripts/exp_onetile.py: 8 k cycles per tile with the
    // global loads and stores compiled out), so instructions removed from the rounds are time removed from the tail.
    const bool fast = (row_base + 32 <= p.M) && (col_base + NCOLS <= p.N) && (p.gate == nullptr || gate_uniform);
    const size_t off0 = static_cast<size_t>(row_base + rr0) * p.ldo + col_base + cc * 4;   // (first row, first chunk) of this lane
    const size_t pitch = static_cast<size_t>(4) * p.ldo;                                    // one round further = 4 rows
    // Residual rows are software-pipelined one unit ahead (and unit 0 is fetched BEFORE the accumulator is waited
    // for): they do not depend on the MMA, and with <= 1 KB of L1 left beside 225 KB of shared memory every one of
    // them is an L2 round trip.  out may alias resid element for element; a unit's loads precede its stores.
    float4 r4[2][8];
    auto fetch_resid = [&](int u, float4(&r)[8]) {
      if (fast) {
        const float* src = p.resid + off0 + u * 32;
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = __ldcg(reinterpret_cast<const float4*>(src + i * pitch));
        return;
      }
      const int col = col_base + u * 32 + cc * 4;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = row_base + rr0 + 4 * i;
        r[i] = (col < p.N && row < p.M)
                   ? __ldcg(reinterpret_cast<const float4*>(p.resid + static_cast<size_t>(row) * p.ldo + col))
                   : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    if constexpr (EPI == LDT_EPI_GATE_RESID_F32) fetch_resid(0, r4[0]);
    // Gate and bias of EVERY unit are fetched before the accumulator is waited for: they do not depend on the MMA, the
    // inline-asm TMEM / shared-memory steps below are compiler barriers (a load written inside the loop is issued inside
    // the loop), and each one is an L2 round trip.
    float4 g4s[NU], b4s[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      const int col = col_base + u * 32 + cc * 4;
      g4s[u] = make_float4(1.f, 1.f, 1.f, 1.f);
      b4s[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if constexpr (EPI == LDT_EPI_GATE_RESID_F32) {
        if (col < p.N && p.gate != nullptr && gate_uniform && row_base < p.M)   // slabs past M (ragged last tile) have no gate row
          g4s[u] = __ldg(reinterpret_cast<const float4*>(
              p.gate + static_cast<long long>(row_base / p.rows_per_gate) * p.gate_stride + col));
      }
      if (col < p.N && p.bias != nullptr) b4s[u] = __ldg(reinterpret_cast<const float4*>(p.bias + col));
    }
    // one load-add-store round: staged row rr0 + 4*i, 16-byte chunk cc -> bias, gate, residual -> global
    const int relu = p.relu;
    auto finish = [&](float4 a4, const float4& g, const float4& b4, const float4& r) -> float4 {
#ifdef LDT_EPI_SCALAR_F32   // A/B builds only
      a4.x += b4.x; a4.y += b4.y; a4.z += b4.z; a4.w += b4.w;
      if constexpr (EPI == LDT_EPI_GATE_RESID_F32) {
        a4.x = r.x + g.x * a4.x; a4.y = r.y + g.y * a4.y;
        a4.z = r.z + g.z * a4.z; a4.w = r.w + g.w * a4.w;
      }
#else
      // packed fp32 (FADD2 / FFMA2), two columns per instruction; each half rounds like the scalar form
      uint64_t lo = add_f32x2(pack_f32x2(a4.x, a4.y), pack_f32x2(b4.x, b4.y));
      uint64_t hi = add_f32x2(pack_f32x2(a4.z, a4.w), pack_f32x2(b4.z, b4.w));
      if constexpr (EPI == LDT_EPI_GATE_RESID_F32) {
        lo = fma_f32x2(pack_f32x2(g.x, g.y), lo, pack_f32x2(r.x, r.y));
        hi = fma_f32x2(pack_f32x2(g.z, g.w), hi, pack_f32x2(r.z, r.w));
      }
      unpack_f32x2(lo, a4.x, a4.y);
      unpack_f32x2(hi, a4.z, a4.w);
#endif
      if constexpr (EPI == LDT_EPI_BIAS_GELU_F32) {   // TF32 parity mode: exact-erf GELU, rounded to TF32
        a4.x = gelu_erf_f(a4.x); a4.y = gelu_erf_f(a4.y); a4.z = gelu_erf_f(a4.z); a4.w = gelu_erf_f(a4.w);
        if (!p.f32_plain) { a4.x = round_tf32(a4.x); a4.y = round_tf32(a4.y); a4.z = round_tf32(a4.z); a4.w = round_tf32(a4.w); }
      } else if (relu) {   // warp-uniform; the encoder prologue's Conv1d + BatchNorm + ReLU layers (LDT_EPI_*_RELU_F32)
        a4.x = epi_relu(a4.x, relu); a4.y = epi_relu(a4.y, relu); a4.z = epi_relu(a4.z, relu); a4.w = epi_relu(a4.w, relu);
      }
      return a4;
    };
    LDT_STAMP(0);
    wait_accumulator();
    LDT_STAMP(1);
    // The TMEM load of unit u+1 is issued right after unit u's registers have been stored to the staging buffer -- into the
    // SAME registers, which nothing touches until the wait -- so it streams in while unit u's 8 store rounds run.
    uint32_t v[32];
    if (col_base < p.N) {
      tmem_ld_32x32(taddr, v);
      tmem_ld_wait();
    }
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      const int c0 = u * 32;
      const int col = col_base + c0 + cc * 4;
      const bool col_ok = col < p.N;   // N % 8 == 0 and col % 4 == 0: the whole float4 is inside
      const bool next_live = (u + 1 < NU) && (col_base + c0 + 32 < p.N);   // warp-uniform
      if constexpr (EPI == LDT_EPI_GATE_RESID_F32) {
        if (u + 1 < NU) fetch_resid(u + 1, r4[(u + 1) & 1]);
      }
      const float4 g4 = g4s[u], b4 = b4s[u];
      if (col_base + c0 < p.N) {   // warp-uniform: this unit has at least one live column
        LDT_STAMP(2 + 4 * u);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t a = st_row + static_cast<uint32_t>((c ^ (lane & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v[4 * c]), "r"(v[4 * c + 1]),
                       "r"(v[4 * c + 2]), "r"(v[4 * c + 3])
                       : "memory");
        }
        __syncwarp();
        LDT_STAMP(3 + 4 * u);
#ifndef LDT_EPI_NO_TMEM_PIPELINE   // (A/B builds: the load issued after the store rounds instead, i.e. phase after phase)
        if (next_live) tmem_ld_32x32(taddr + static_cast<uint32_t>(c0 + 32), v);   // in flight during the store rounds
#endif
        LDT_STAMP(4 + 4 * u);
        if (fast) {
          float* dst = static_cast<float*>(p.out) + off0 + c0;
          const uint32_t lds0 = stg_u32 + static_cast<uint32_t>(rr0 * 128);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = rr0 + 4 * i;   // (rr & 7) == (rr0 + 4 * (i & 1)) & 7: two swizzle phases per lane
            float4 a4;
            const uint32_t a = lds0 + static_cast<uint32_t>(i * 512 + ((cc ^ (rr & 7)) << 4));
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a4.x), "=f"(a4.y), "=f"(a4.z), "=f"(a4.w) : "r"(a));
            *reinterpret_cast<float4*>(dst + i * pitch) = finish(a4, g4, b4, r4[u & 1][i]);
          }
        } else if (col_ok) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = rr0 + 4 * i;
            const int row = row_base + rr;
            float4 a4;
            const uint32_t a = stg_u32 + static_cast<uint32_t>(rr * 128 + ((cc ^ (rr & 7)) << 4));
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a4.x), "=f"(a4.y), "=f"(a4.z), "=f"(a4.w) : "r"(a));
            if (row < p.M) {
              float4 g = g4;
              if constexpr (EPI == LDT_EPI_GATE_RESID_F32) {
                if (p.gate != nullptr && !gate_uniform)
                  g = __ldg(reinterpret_cast<const float4*>(
                      p.gate + static_cast<long long>(row / p.rows_per_gate) * p.gate_stride + col));
              }
              *reinterpret_cast<float4*>(static_cast<float*>(p.out) + static_cast<size_t>(row) * p.ldo + col) =
                  finish(a4, g, b4, r4[u & 1][i]);
            }
          }
        }
        __syncwarp();
#ifdef LDT_EPI_NO_TMEM_PIPELINE
        if (next_live) tmem_ld_32x32(taddr + static_cast<uint32_t>(c0 + 32), v);
#endif
        if (next_live) tmem_ld_wait();
        LDT_STAMP(5 + 4 * u);
      }
    }
  } else {
    wait_accumulator();
#pragma unroll 1
    for (int c0 = 0; c0 < ncols; c0 += 64) {
      if (col_base + c0 >= p.N) break;
      uint32_t v[64];
      if constexpr (!(XM & 8)) {
        uint32_t(&lo)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[0]);
        uint32_t(&hi)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[32]);
        tmem_ld_32x32(taddr + static_cast<uint32_t>(c0), lo);
        if (c0 + 32 < ncols) tmem_ld_32x32(taddr + static_cast<uint32_t>(c0 + 32), hi);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 64; ++j) v[j] = __float_as_uint(0.01f * static_cast<float>(lane + j + c0));
      }
      const int colb = col_base + c0;
      if (tm_out != nullptr) {   // the previous unit's bulk store must have read the staging buffer
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {   // 8 columns -> one 16-byte chunk of bf16
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[8 * c + j]);
        if (p.bias != nullptr && colb + 8 * c < p.N) {
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + colb + 8 * c));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + colb + 8 * c + 4));
          f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w; f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
        }
        if constexpr (EPI == LDT_EPI_BIAS_GELU_BF16 && !(XM & 64)) {
#ifdef LDT_GELU_SCALAR   // A/B builds only
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = gelu_erf_fast(f[j]);
#else
#pragma unroll
          for (int j = 0; j < 8; j += 2) gelu_erf_fast_x2(f[j], f[j + 1]);
#endif
        }
        const uint32_t a = st_row + static_cast<uint32_t>((c ^ (lane & 7)) << 4);
        if constexpr (!(XM & 16)) {
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pack_bf16x2(f[0], f[1])),
                       "r"(pack_bf16x2(f[2], f[3])), "r"(pack_bf16x2(f[4], f[5])), "r"(pack_bf16x2(f[6], f[7]))
                       : "memory");
        } else {   // keep the arithmetic alive: the packed values are what the stores below write
          v[4 * c] = pack_bf16x2(f[0], f[1]); v[4 * c + 1] = pack_bf16x2(f[2], f[3]);
          v[4 * c + 2] = pack_bf16x2(f[4], f[5]); v[4 * c + 3] = pack_bf16x2(f[6], f[7]);
        }
      }
      if (tm_out != nullptr) {
        if constexpr (!(XM & 32)) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && row_base < p.M) {
            tma_store_2d(tm_out, stg_u32, colb, row_base);
            tma_store_commit();
          }
        }
        continue;
      }
      __syncwarp();
      const int col = colb + cc * 8;
      if (col < p.N) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = rr0 + 4 * i;
          const int row = row_base + rr;
          uint4 u;
          const uint32_t a = stg_u32 + static_cast<uint32_t>(rr * 128 + ((cc ^ (rr & 7)) << 4));
          if constexpr (!(XM & 16))
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(a));
          else
            u = make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          if constexpr (XM & 32) {   // no store: fold the values into something the compiler must keep
            if ((u.x ^ u.y ^ u.z ^ u.w) == 0x12345678u && row < 0) *reinterpret_cast<uint4*>(p.out) = u;
          } else if (row < p.M)
            *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(row) * p.ldo + col) = u;
        }
      }
      __syncwarp();
    }
  }
}

}  // namespace ldt
