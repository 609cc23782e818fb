// Attention over the 32 latent tokens on the 5th-generation tensor cores:  S = Q K^T and O = P V as tcgen05.mma
// (cta_group::1, M = 128 query rows per instruction, accumulators in TMEM), softmax in registers in between.
//
// Replaces ResidualBlock.compute_attention (model/layers.py:183-200) for
//   * the Compressor decoder's cross-attention (model/Compressor/Network.py:80-83): 2048 query points x 32 keys per
//     (sample, head), dh 32 -- one 128-query tile per CTA iteration, keys/values of ONE (sample, head)  [SPT = 1];
//   * the score net's cross-attention to the condition tokens (score.py:148-149) and any 32-query x 32-key call:
//     a 128-row tile stacks the queries of FOUR samples of one head; S is the 128 x 128 product against the four
//     samples' keys, of which each sample reads its own diagonal 32 x 32 block, and P is written block-diagonal
//     (zeros elsewhere, written once) so that one K = 128 product gives every sample its own P V            [SPT = 4].
// Same output layout quirk as attention.cu: [B,H,Nq,dh] stored contiguously (layers.py:197).
//
// Operand layouts (all staged by the CTA's own threads; every row is one 128-byte swizzle row, 16-byte chunk c of row r
// at r*128 + ((c ^ (r & 7)) << 4), i.e. the SWIZZLE_128B pattern TMA would produce):
//   Q [128 rows x dh]    A of S,  K-major           K [keys x dh]   B of S,  K-major (N = keys)
//   P [128 rows x keys]  A of PV, read from TENSOR MEMORY (tcgen05.mma with a TMEM A operand: lane = row, two bf16 per
//                        32-bit column): every thread writes its own lane with tcgen05.st, no shared memory, no proxy fence
//   V [keys x dh]        B of PV, MN-major: rows are the K index (keys), dh contiguous -- exactly how V sits in memory,
//                        so no transpose is needed (instruction-descriptor bit 16 = B is MN-major).
// With tcgen05.ld 32x32b every thread owns ONE query row: the row maximum / sum of the softmax are plain per-thread
// loops over 32 registers (no shuffles), and P leaves as 64 contiguous bytes per thread.
#include <cuda.h>

#include "common.cuh"
#include "ldt_b200.h"

namespace ldt {

constexpr int AT_THREADS = 128;

template <int DH, int SPT>
struct AtCfg {
  static constexpr int NK = 32 * SPT;                    // keys per tile
  static constexpr int Q_BYTES = 128 * 128;
  static constexpr int KV_BYTES = NK * 128;
  // TMEM columns: S (fp32, NK wide) at 0; P (bf16 pairs, NK / 2 wide) and O (fp32, DH wide) behind it
  static constexpr int P_COL = (SPT == 1) ? 32 : 192;
  static constexpr int O_COL = (SPT == 1) ? 64 : 128;
  static constexpr int TMEM_COLS = (SPT == 1) ? 128 : 256;
  static constexpr int SMEM_BYTES = 1024 + Q_BYTES + 2 * KV_BYTES + 64;
};

// Instruction descriptor of common.cuh with the B operand MN-major (bit 16).
__host__ __device__ constexpr uint32_t umma_idesc_bf16_bmn(int m, int n) { return umma_idesc_bf16(m, n) | (1u << 16); }

// UMMA shared-memory descriptor of an MN-major operand tile stored as rows of 128 bytes (row = K index, 64 MN elements
// per row) with the 128-byte swizzle: 8-row groups are 1024 bytes apart (stride byte offset); a single 64-element MN atom,
// so the leading byte offset is never used.  Field layout: cute/arch/mma_sm100_desc.hpp, canonical form
// Swizzle<3,4,3> o ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units (cute/atom/mma_traits_sm100.hpp).
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;                      // leading byte offset (unused: one MN atom)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;              // stride byte offset: 8 K-rows * 128 B
  d |= static_cast<uint64_t>(1) << 46;                      // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;                      // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

template <int DH, int SPT>
__global__ void __launch_bounds__(AT_THREADS)
attention_tc_kernel(int tiles, int B, int H, int Nq, int qtiles, const __nv_bfloat16* __restrict__ q, int ldq,
                    const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v, int ldkv,
                    __nv_bfloat16* __restrict__ o, float scale_log2e) {
  using Cfg = AtCfg<DH, SPT>;
  constexpr int NK = Cfg::NK;
  constexpr int CH = DH / 8;   // 16-byte chunks per operand row
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const uint32_t sQ = smem_u32(smem), sK = sQ + Cfg::Q_BYTES, sV = sK + Cfg::KV_BYTES;
  uint64_t* bar_s = reinterpret_cast<uint64_t*>(smem + Cfg::Q_BYTES + 2 * Cfg::KV_BYTES);
  uint64_t* bar_o = bar_s + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_o + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  pdl_launch_dependents();
  if (tid == 0) {
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    mbar_fence_init();
  }
  if (warp == 1) {   // (a whole warp, and not the one whose lane 0 just diverged to initialise the barriers)
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  uint32_t phase = 0;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    // ---- which rows ----
    int h, row0_q, rows_q_valid, row0_kv, rows_kv_valid;
    size_t out_base;   // element offset of this tile's row 0 in o (SPT = 1), or unused
    int b0;
    if (SPT == 1) {
      const int qt = tile % qtiles, bh = tile / qtiles;
      h = bh % H;
      b0 = bh / H;
      row0_q = b0 * Nq + qt * 128;
      rows_q_valid = min(128, Nq - qt * 128);
      row0_kv = b0 * 32;
      rows_kv_valid = 32;
      out_base = (static_cast<size_t>(bh) * Nq + qt * 128) * DH;
    } else {
      const int bt = tile / H;   // group of four samples
      h = tile % H;
      b0 = bt * 4;
      row0_q = b0 * 32;
      rows_q_valid = min(128, B * 32 - row0_q);
      row0_kv = row0_q;
      rows_kv_valid = rows_q_valid;
      out_base = 0;
    }
    // ---- stage Q (thread = row), K and V (chunk-strided) ----
    {
      const int r = tid;
      const uint4* src = reinterpret_cast<const uint4*>(q + static_cast<size_t>(row0_q + min(r, rows_q_valid - 1)) * ldq + h * DH);
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const uint4 u = src[c];
        st_shared_v4(sQ + r * 128 + ((c ^ (r & 7)) << 4), u.x, u.y, u.z, u.w);
      }
      for (int i = tid; i < NK * CH; i += AT_THREADS) {
        const int kr = i / CH, c = i % CH;
        const bool ok = kr < rows_kv_valid;
        const size_t off = static_cast<size_t>(row0_kv + (ok ? kr : 0)) * ldkv + h * DH + c * 8;
        uint4 uk = *reinterpret_cast<const uint4*>(k + off), uv = *reinterpret_cast<const uint4*>(v + off);
        if (!ok) uk = uv = make_uint4(0u, 0u, 0u, 0u);
        const uint32_t a = kr * 128 + ((c ^ (kr & 7)) << 4);
        st_shared_v4(sK + a, uk.x, uk.y, uk.z, uk.w);
        st_shared_v4(sV + a, uv.x, uv.y, uv.z, uv.w);
      }
    }
    fence_proxy_async_smem();
    __syncthreads();

    // ---- S[128 x NK] = Q K^T ----
    if (tid == 0) {
      tc_fence_after();
      constexpr uint32_t idesc = umma_idesc_bf16(128, NK);
      const uint64_t da = umma_desc_k_sw128(sQ), db = umma_desc_k_sw128(sK);
#pragma unroll
      for (int s = 0; s < DH / 16; ++s) umma_bf16_ss(tmem_base, da + 2 * s, db + 2 * s, idesc, s != 0 ? 1u : 0u);
      umma_commit(bar_s);
    }
    mbar_wait(bar_s, phase);
    tc_fence_after();

    // ---- softmax of this thread's row over its 32 keys; un-normalised P -> bf16 -> shared memory ----
    const int blk = (SPT == 1) ? 0 : warp;   // which 32-key block of the tile belongs to this row's sample
    uint32_t sv[32];
    tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + static_cast<uint32_t>(blk * 32), sv);
    tmem_ld_wait();
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(sv[j]));
    float sum = 0.f;
    uint32_t pw[16];
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      const float p0 = exp2f((__uint_as_float(sv[j]) - m) * scale_log2e);
      const float p1 = exp2f((__uint_as_float(sv[j + 1]) - m) * scale_log2e);
      sum += p0 + p1;
      pw[j >> 1] = pack_bf16x2(p0, p1);
    }
    const float inv_sum = 1.0f / sum;
    {
      // this row's lane of the P operand: its own 32 keys (16 columns of bf16 pairs) at column blk * 16, zeros in the
      // other samples' key ranges (the block-diagonal form; SPT = 1 has a single block)
      const uint32_t prow = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + static_cast<uint32_t>(Cfg::P_COL);
      const uint32_t zero[16] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
#pragma unroll
      for (int j = 0; j < SPT; ++j) {
        if (j == blk) tmem_st_32x16(prow + j * 16, pw);
        else tmem_st_32x16(prow + j * 16, zero);
      }
      tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();

    // ---- O[128 x DH] = P V ----
    if (tid == 0) {
      tc_fence_after();
      constexpr uint32_t idesc = umma_idesc_bf16_bmn(128, DH);
      const uint64_t dv = umma_desc_mn_sw128(sV);
#pragma unroll
      for (int s = 0; s < NK / 16; ++s)
        umma_bf16_ts(tmem_base + Cfg::O_COL, tmem_base + Cfg::P_COL + s * 8, dv + static_cast<uint64_t>(s * (2048 >> 4)), idesc,
                     s != 0 ? 1u : 0u);
      umma_commit(bar_o);
    }
    mbar_wait(bar_o, phase);
    tc_fence_after();
    phase ^= 1u;

    // ---- O / sum -> bf16 -> global, one contiguous row of dh elements per thread ----
    {
      const bool live = tid < rows_q_valid;
      __nv_bfloat16* dst;
      if (SPT == 1) {
        dst = o + out_base + static_cast<size_t>(tid) * DH;
      } else {
        const int b = b0 + warp, n = tid & 31;
        dst = o + ((static_cast<size_t>(b) * H + h) * Nq + n) * DH;
      }
#pragma unroll
      for (int c0 = 0; c0 < DH; c0 += 32) {
        uint32_t ov[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + static_cast<uint32_t>(Cfg::O_COL + c0), ov);
        tmem_ld_wait();
        if (live) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(ov[j]) * inv_sum, __uint_as_float(ov[j + 1]) * inv_sum);
            u.y = pack_bf16x2(__uint_as_float(ov[j + 2]) * inv_sum, __uint_as_float(ov[j + 3]) * inv_sum);
            u.z = pack_bf16x2(__uint_as_float(ov[j + 4]) * inv_sum, __uint_as_float(ov[j + 5]) * inv_sum);
            u.w = pack_bf16x2(__uint_as_float(ov[j + 6]) * inv_sum, __uint_as_float(ov[j + 7]) * inv_sum);
            *reinterpret_cast<uint4*>(dst + c0 + j) = u;
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();   // every thread has read S / O and the MMAs have read Q, K, V, P: the tile buffers are free
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int DH, int SPT>
static int launch_attention_tc(int B, int H, int Nq, const void* q, int ldq, const void* k, const void* v, int ldkv, void* o,
                               cudaStream_t s) {
  using Cfg = AtCfg<DH, SPT>;
  static PerDevice<bool> attr_done;
  if (!attr_done.get()) {
    LDT_CUDA_OK(cudaFuncSetAttribute(attention_tc_kernel<DH, SPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done.get() = true;
  }
  const int qtiles = (Nq + 127) / 128;
  const long long tiles = (SPT == 1) ? static_cast<long long>(B) * H * qtiles : static_cast<long long>((B + 3) / 4) * H;
  LDT_REQUIRE(tiles < (1LL << 31), LDT_ERR_INVALID, "ldt_attention_nk32: too many work units");
  const int per_sm = 512 / Cfg::TMEM_COLS;   // TMEM columns bound the co-resident CTAs
  const int grid = static_cast<int>(tiles < static_cast<long long>(num_sms()) * per_sm ? tiles : static_cast<long long>(num_sms()) * per_sm);
  const float scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(DH));
  LDT_CUDA_OK(launch_pdl(attention_tc_kernel<DH, SPT>, dim3(grid), dim3(AT_THREADS), Cfg::SMEM_BYTES, s, static_cast<int>(tiles), B, H, Nq,
                         qtiles, static_cast<const __nv_bfloat16*>(q), ldq, static_cast<const __nv_bfloat16*>(k),
                         static_cast<const __nv_bfloat16*>(v), ldkv, static_cast<__nv_bfloat16*>(o), scale_log2e));
  return LDT_OK;
}

// Entry used by ldt_attention_nk32 (attention.cu): returns LDT_ERR_UNSUPPORTED for shapes this kernel does not take.
int attention_nk32_tc(int B, int H, int Nq, int dh, const void* q, int ldq, const void* k, const void* v, int ldkv, void* o,
                      cudaStream_t s) {
  if (Nq == 32) {
    if (dh == 64) return launch_attention_tc<64, 4>(B, H, Nq, q, ldq, k, v, ldkv, o, s);
    if (dh == 32) return launch_attention_tc<32, 4>(B, H, Nq, q, ldq, k, v, ldkv, o, s);
  } else if (Nq >= 128) {
    if (dh == 64) return launch_attention_tc<64, 1>(B, H, Nq, q, ldq, k, v, ldkv, o, s);
    if (dh == 32) return launch_attention_tc<32, 1>(B, H, Nq, q, ldq, k, v, ldkv, o, s);
  }
  return LDT_ERR_UNSUPPORTED;
}

}  // namespace ldt
