// Fused self-attention front half of a score-net block:  LN'd activations -> Q/K/V projection -> per-head
// softmax(q k^T / sqrt(dh)) v  in ONE kernel; Q, K, V never reach HBM.
//
// Replaces, for the self-attention blocks, fc_q + fc_kv (model/layers.py:186-189) and compute_attention's
// permute/clone + bmm + mul + softmax + bmm (layers.py:192-197): 2 cuDNN convs + ~7 launches in the reference, and a
// [M,3072] bf16 round trip plus a separate attention launch in this repo's unfused path (gemm.cu + attention.cu).
//
// Shape of the work: M = B*32 token rows, K = hidden (1024), 16 heads x dh 64.  The projection weight is packed head-
// major (rows h*192 + [0,64) = q_h, [64,128) = k_h, [128,192) = v_h; score.py::packed), so one 256-row x 192-column
// GEMM tile holds everything 8 samples need for one head.  32 x 16 = 512 tiles over 74 CTA pairs = 6.92 waves (the
// plain N=3072 GEMM with 256-wide tiles quantises to 5.19 -> 6 waves).
//
// Kernel structure = the CTA-pair tcgen05 GEMM of gemm.cu (TMA producer warp, one MMA thread issuing
// tcgen05.mma.cta_group::2 256x192x16, two TMEM accumulators) with a different epilogue: the 8 epilogue warps form two
// sets; set s drains accumulator s (so every set has two mainloops of time per tile).  Warp (set, quad) owns TMEM lanes
// 32*quad..+31 = the 32 tokens of ONE sample: it pulls Q,K (then V) out of TMEM, adds the bias, rounds to bf16 into a
// private shared-memory tile, and runs the 32x32 attention.  The output is written in the reference's layout quirk:
// [B,H,32,dh] contiguous, which the next layer re-reads as token-major [B*32, H*dh] (layers.py:197).
//
// Attention arithmetic (TC = true, the product path): S = Q K^T and O = P V are tcgen05.mma.cta_group::1 products of the
// CTA's OWN 128 rows (four samples stacked), issued by one thread of the epilogue set and accumulated in the set's own
// accumulator columns once Q | K | V have been drained:
//   Q  -> (+bias) bf16 pairs -> TMEM columns [192,224) of the accumulator slot  (A operand of S, read from tensor memory)
//   K  -> (+bias) bf16 -> shared memory [128 tokens x 64], 128-byte swizzle rows (B operand of S, K-major, N = 128 keys)
//   V  -> (+bias) bf16 -> shared memory [128 tokens x 64]                       (B operand of PV, MN-major)
//   S[128 x 128] -> accumulator columns [0,128): every sample reads its own diagonal 32 x 32 block (one row per thread:
//        the softmax maximum / sum are per-thread loops, no shuffles)
//   P  -> bf16 pairs, block-diagonal (zeros outside the sample's own 32 keys) -> TMEM columns [192,256)  (A of PV)
//   O[128 x 64]  -> accumulator columns [128,192) -> * 1/sum -> bf16 -> staged -> one contiguous 4 KB block per sample
// The accumulator goes back to the mainloop after O has been read (one mainloop of time per tile and set).
// TC = false keeps the warp-level mma.sync m16n8k16 arithmetic of attention.cu (cross-check, ldt_debug_set_attention_backend).
#include <cuda.h>

#include "common.cuh"
#include "ldt_b200.h"
#include "mma_sync.cuh"
#include "tmap.cuh"

namespace ldt {

constexpr int QA_BK = 64;
constexpr int QA_DH = 64;
constexpr int QA_BN = 3 * QA_DH;             // 192: q_h | k_h | v_h
constexpr int QA_THREADS = 384;
constexpr int QA_EPI_WARP0 = 4;
constexpr int QA_A_BYTES = 128 * QA_BK * 2;  // 16 KB: this CTA's 128 rows
constexpr int QA_B_BYTES = (QA_BN / 2) * QA_BK * 2;  // 12 KB: this CTA's half of the head's 192 weight rows
constexpr int QA_STAGES = 5;
constexpr int QA_ACC_STRIDE = 256;
constexpr int QA_TMEM_COLS = 512;
constexpr int QA_QK_LD = 2 * QA_DH + 8;      // bf16 elements per staged Q|K row (272 B: conflict-free fragment loads)
constexpr int QA_V_LD = QA_DH + 8;           // bf16 elements per staged V / O row (144 B)
constexpr int QA_STG_BYTES = 2 * 32 * QA_V_LD * 2;  // 9216 B per warp: Q|K tile (8704 B), later V tile + O tile (4608 B each)
constexpr int QA_STG_ALL = 8 * QA_STG_BYTES;   // 73728 B: the mma.sync path's 8 private tiles; the tcgen05 path uses 2 x (K 16 KB + V 16 KB)
constexpr int QA_SMEM_BYTES = 1024 + QA_STAGES * (QA_A_BYTES + QA_B_BYTES) + QA_STG_ALL + 256;

__host__ __device__ constexpr uint32_t qa_idesc_bmn(int m, int n) { return umma_idesc_bf16(m, n) | (1u << 16); }   // B MN-major
__device__ __forceinline__ uint64_t qa_desc_mn_sw128(uint32_t smem_addr) {   // see attention_tc.cu::umma_desc_mn_sw128
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__device__ __forceinline__ void qa_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

struct QkvAttnParams {
  int M;                 // token rows = B * 32
  int H;                 // heads
  const float* bias;     // [H * 192] head-major packed, or nullptr
  __nv_bfloat16* out;    // [B, H, 32, 64] contiguous
  float scale_log2e;
};

template <bool TC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(QA_THREADS, 1)
qkv_attention_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                     const QkvAttnParams p, const int K, const int tiles_m) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + QA_STAGES * QA_A_BYTES;
  uint8_t* stg_all = sB + QA_STAGES * QA_B_BYTES;   // 1024-byte aligned (UMMA operand tiles of the tcgen05 attention)
  uint64_t* full = reinterpret_cast<uint64_t*>(stg_all + QA_STG_ALL);
  uint64_t* empty = full + QA_STAGES;
  uint64_t* tfull = empty + QA_STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* sbar = tempty + 2;    // [set] S = Q K^T complete
  uint64_t* obar = sbar + 2;      // [set] O = P V complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(obar + 2);

  // Role index = physical warp id rotated by 4: the TMA / MMA / TMEM-alloc warps are PHYSICAL warps 8, 9, 10 and the
  // epilogue warps are physical warps 0-7.  The SM's issue arbiter prefers the highest warp id of a sub-partition, so the
  // single MMA-issuing thread must not sit below ALU-heavy epilogue warps (measured: with the MMA thread in warp 1 the
  // tensor pipe ran at 76 % of its rate under the GELU epilogue).  (role & 3) == (physical & 3): TMEM lane quadrants hold.
  const int warp = (__shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0) + 4) % 12;   // warp-uniform by construction
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int num_tiles = tiles_m * p.H;
  const int num_kb = K / QA_BK;

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < QA_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 8);   // 4 warps of one epilogue set x 2 CTAs
      mbar_init(&sbar[a], 1);
      mbar_init(&obar[a], 1);
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(tmem_slot, QA_TMEM_COLS);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ---- TMA producer: converged warp, one elected lane issues (see gemm.cu "Issue loops") ----
    const uint32_t sA0 = smem_u32(sA), sB0 = smem_u32(sB);
    const uint32_t empty0 = smem_u32(empty), full0 = smem_u32(full);
    const uint32_t full0_leader = mapa_u32(full0, 0);
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int m0 = (tile % tiles_m) * 256 + static_cast<int>(rank) * 128;
      const int n0 = (tile / tiles_m) * QA_BN + static_cast<int>(rank) * (QA_BN / 2);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait_u32(empty0 + stage * 8, phase ^ 1u);
        if (elect_one()) {
          if (rank == 0) mbar_expect_tx_u32(full0 + stage * 8, 2 * (QA_A_BYTES + QA_B_BYTES));
          tma_load_2d_pair_u32(sA0 + stage * QA_A_BYTES, &tmA, full0_leader + stage * 8, kb * QA_BK, m0);
          tma_load_2d_pair_u32(sB0 + stage * QA_B_BYTES, &tmW, full0_leader + stage * 8, kb * QA_BK, n0);
        }
        __syncwarp();
        if (++stage == QA_STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer (leader CTA): converged warp, one elected lane issues ----
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, QA_BN);
      const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty), tfull0 = smem_u32(tfull), tempty0 = smem_u32(tempty);
      const uint64_t descA0 = umma_desc_k_sw128(smem_u32(sA));
      const uint64_t descB0 = umma_desc_k_sw128(smem_u32(sB));
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
        const int acc = it & 1;
        mbar_wait_u32(tempty0 + acc * 8, ((it >> 1) & 1) ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * QA_ACC_STRIDE);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait_u32(full0 + stage * 8, phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t da = descA0 + static_cast<uint64_t>(stage * (QA_A_BYTES >> 4));
            const uint64_t db = descB0 + static_cast<uint64_t>(stage * (QA_B_BYTES >> 4));
            umma_bf16_ss_pair(tmem_d, da, db, idesc, kb != 0 ? 1u : 0u);
#pragma unroll
            for (int k = 1; k < QA_BK / 16; ++k) umma_bf16_ss_pair_acc(tmem_d, da + 2 * k, db + 2 * k, idesc);
            umma_commit_pair_u32(empty0 + stage * 8, 0x3);
          }
          __syncwarp();
          if (++stage == QA_STAGES) { stage = 0; phase ^= 1u; }
        }
        if (elect_one()) umma_commit_pair_u32(tfull0 + acc * 8, 0x3);
        __syncwarp();
      }
    }
  } else if (TC && warp >= QA_EPI_WARP0) {
    // ================= tcgen05 attention epilogue (see the header comment) =================
    const int quad = warp & 3;                       // TMEM lane quadrant = sample within this CTA's 4 samples
    const int set = (warp - QA_EPI_WARP0) >> 2;      // which accumulator this warp drains
    const int r = quad * 32 + lane;                  // this thread's row of the CTA's 128
    const uint32_t sK = smem_u32(stg_all) + static_cast<uint32_t>(set) * 32768u, sV = sK + 16384u;
    const uint32_t acc = tmem_base + static_cast<uint32_t>(set * QA_ACC_STRIDE);   // accumulator slot, column base
    const uint32_t t_row = acc + (static_cast<uint32_t>(quad * 32) << 16);          // ... at this warp's lanes
    const uint32_t st_row = static_cast<uint32_t>(r) * 128u;                        // staged row offset (K / V / O tiles)
    const bool issuer = (quad == 0 && lane == 0);
    int it = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
      if ((it & 1) != set) continue;
      const uint32_t ph = static_cast<uint32_t>(it >> 1) & 1u;
      const int head = tile / tiles_m;
      const int row0 = (tile % tiles_m) * 256 + static_cast<int>(rank) * 128 + quad * 32;   // first token row of the sample
      const bool live = row0 < p.M;                  // whole samples only: M % 32 == 0
      const float* bias = p.bias ? p.bias + head * QA_BN : nullptr;
      mbar_wait(&tfull[set], ph);
      tc_fence_after();

      // ---- drain: Q -> bf16 pairs -> TMEM [192,224);  K, V -> bf16 -> shared-memory operand tiles ----
#pragma unroll 1
      for (int part = 0; part < 3; ++part) {         // 0 = Q, 1 = K, 2 = V: 64 accumulator columns each
        uint32_t v0[32], v1[32];
        tmem_ld_32x32(t_row + static_cast<uint32_t>(part * 64), v0);
        tmem_ld_32x32(t_row + static_cast<uint32_t>(part * 64 + 32), v1);
        tmem_ld_wait();
        uint32_t pk[32];                              // 64 bf16 of this row
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const uint32_t(&v)[32] = hh ? v1 : v0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[8 * j + e]);
            if (bias) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + part * 64 + hh * 32 + 8 * j));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + part * 64 + hh * 32 + 8 * j + 4));
              f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w; f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) pk[hh * 16 + 4 * j + e] = pack_bf16(f[2 * e], f[2 * e + 1]);
          }
        }
        if (part == 0) {
          const uint32_t(&lo)[16] = *reinterpret_cast<const uint32_t(*)[16]>(&pk[0]);
          const uint32_t(&hi)[16] = *reinterpret_cast<const uint32_t(*)[16]>(&pk[16]);
          tmem_st_32x16(t_row + 192u, lo);
          tmem_st_32x16(t_row + 208u, hi);
        } else {
          const uint32_t base = (part == 1 ? sK : sV) + st_row;
#pragma unroll
          for (int c = 0; c < 8; ++c)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + static_cast<uint32_t>((c ^ (r & 7)) << 4)),
                         "r"(pk[4 * c]), "r"(pk[4 * c + 1]), "r"(pk[4 * c + 2]), "r"(pk[4 * c + 3])
                         : "memory");
        }
      }
      tmem_st_wait();
      tc_fence_before();
      fence_proxy_async_smem();
      qa_bar_sync(1 + set, 128);     // all four samples' Q, K, V are in place and out of the accumulator columns
#ifdef LDT_QA_EARLY_RELEASE   // timing experiment only (results are WRONG): hand the accumulator back right after the drain
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[set]), 0));
#endif

      // ---- S[128 x 128] = Q K^T into accumulator columns [0,128) ----
      if (issuer) {
        tc_fence_after();
        constexpr uint32_t idesc = umma_idesc_bf16(128, 128);
        const uint64_t dk = umma_desc_k_sw128(sK);
#pragma unroll
        for (int ks = 0; ks < QA_DH / 16; ++ks) umma_bf16_ts(acc, acc + 192u + ks * 8, dk + 2 * ks, idesc, ks != 0 ? 1u : 0u);
        umma_commit(&sbar[set]);
      }
      mbar_wait(&sbar[set], ph);
      tc_fence_after();

      // ---- softmax of this thread's row over its sample's 32 keys; un-normalised P -> bf16 pairs ----
      uint32_t sv[32];
      tmem_ld_32x32(t_row + static_cast<uint32_t>(quad * 32), sv);
      tmem_ld_wait();
      float m = -INFINITY;
#pragma unroll
      for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(sv[j]));
      float sum = 0.f;
      uint32_t pw[16];
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float p0 = exp2f((__uint_as_float(sv[j]) - m) * p.scale_log2e);
        const float p1 = exp2f((__uint_as_float(sv[j + 1]) - m) * p.scale_log2e);
        sum += p0 + p1;
        pw[j >> 1] = pack_bf16(p0, p1);
      }
      const float inv_sum = 1.0f / sum;
      {
        const uint32_t zero[16] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
#pragma unroll
        for (int j = 0; j < 4; ++j) {   // block-diagonal P: the sample's own keys at columns 16 * quad, zeros elsewhere
          if (j == quad) tmem_st_32x16(t_row + 192u + j * 16, pw);
          else tmem_st_32x16(t_row + 192u + j * 16, zero);
        }
        tmem_st_wait();
      }
      tc_fence_before();
      qa_bar_sync(1 + set, 128);

      // ---- O[128 x 64] = P V into accumulator columns [128,192) ----
      if (issuer) {
        tc_fence_after();
        constexpr uint32_t idesc = qa_idesc_bmn(128, QA_DH);
        const uint64_t dv = qa_desc_mn_sw128(sV);
#pragma unroll
        for (int ks = 0; ks < 128 / 16; ++ks)
          umma_bf16_ts(acc + 128u, acc + 192u + ks * 8, dv + static_cast<uint64_t>(ks * (2048 >> 4)), idesc, ks != 0 ? 1u : 0u);
        umma_commit(&obar[set]);
      }
      mbar_wait(&obar[set], ph);
      tc_fence_after();
      uint32_t o0[32], o1[32];
      tmem_ld_32x32(t_row + 128u, o0);
      tmem_ld_32x32(t_row + 160u, o1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
#ifndef LDT_QA_EARLY_RELEASE
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[set]), 0));   // the accumulator slot is free again
#endif

      // ---- O / sum -> bf16 -> this row of the (now idle) K tile -> one contiguous 4 KB block per (sample, head) ----
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const uint32_t(&v)[32] = hh ? o1 : o0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = hh * 4 + j;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sK + st_row + static_cast<uint32_t>((c ^ (r & 7)) << 4)),
                       "r"(pack_bf16(__uint_as_float(v[8 * j]) * inv_sum, __uint_as_float(v[8 * j + 1]) * inv_sum)),
                       "r"(pack_bf16(__uint_as_float(v[8 * j + 2]) * inv_sum, __uint_as_float(v[8 * j + 3]) * inv_sum)),
                       "r"(pack_bf16(__uint_as_float(v[8 * j + 4]) * inv_sum, __uint_as_float(v[8 * j + 5]) * inv_sum)),
                       "r"(pack_bf16(__uint_as_float(v[8 * j + 6]) * inv_sum, __uint_as_float(v[8 * j + 7]) * inv_sum))
                       : "memory");
        }
      }
      __syncwarp();
      if (live) {
        const int b = row0 >> 5;
        uint4* dst = reinterpret_cast<uint4*>(p.out + (static_cast<size_t>(b) * p.H + head) * (32 * QA_DH));
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int id = i * 32 + lane;   // 16-byte chunk of the sample's [32][64] bf16 tile
          const int rr = quad * 32 + (id >> 3), c = id & 7;
          uint4 u;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
                       : "r"(sK + static_cast<uint32_t>(rr * 128 + ((c ^ (rr & 7)) << 4))));
          dst[id] = u;
        }
      }
      __syncwarp();   // this warp's rows of the K tile are rewritten by its next tile
    }
  } else if (!TC && warp >= QA_EPI_WARP0) {
    const int quad = warp & 3;                       // TMEM lane quadrant = sample within this CTA's 4 samples
    const int set = (warp - QA_EPI_WARP0) >> 2;      // which accumulator this warp drains
    __nv_bfloat16* stg = reinterpret_cast<__nv_bfloat16*>(stg_all + (warp - QA_EPI_WARP0) * QA_STG_BYTES);
    const uint32_t stg_u32 = smem_u32(stg);
    const int g = lane >> 2, t = lane & 3;
    int it = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
      if ((it & 1) != set) continue;
      const int head = tile / tiles_m;
      const int row0 = (tile % tiles_m) * 256 + static_cast<int>(rank) * 128 + quad * 32;   // first token row
      mbar_wait(&tfull[set], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(set * QA_ACC_STRIDE);
      const bool live = row0 < p.M;                  // whole samples only: M % 32 == 0
      const float* bias = p.bias ? p.bias + head * QA_BN : nullptr;

      // ---- Q | K : TMEM -> (+bias) -> bf16 -> staging rows [token = lane][128] ----
#pragma unroll 1
      for (int c = 0; c < 4; c += 2) {
        uint32_t v0[32], v1[32];
        tmem_ld_32x32(taddr + static_cast<uint32_t>(c * 32), v0);
        tmem_ld_32x32(taddr + static_cast<uint32_t>(c * 32 + 32), v1);
        tmem_ld_wait();
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const uint32_t(&v)[32] = hh ? v1 : v0;
          const int col = (c + hh) * 32;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[8 * j + e]);
            if (bias) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + col + 8 * j));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + col + 8 * j + 4));
              f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w; f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
            }
            const uint32_t a = stg_u32 + static_cast<uint32_t>(lane * (QA_QK_LD * 2) + (col + 8 * j) * 2);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pack_bf16(f[0], f[1])),
                         "r"(pack_bf16(f[2], f[3])), "r"(pack_bf16(f[4], f[5])), "r"(pack_bf16(f[6], f[7]))
                         : "memory");
          }
        }
      }
      // ---- V : TMEM -> registers now (so the accumulator can be handed back), staged after S is done ----
      uint32_t vv0[32], vv1[32];
      tmem_ld_32x32(taddr + 128u, vv0);
      tmem_ld_32x32(taddr + 160u, vv1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[set]), 0));

#ifndef LDT_QA_SKIP_ATTN   // A/B builds only (scripts/exp_ab_lib.py): how much of the kernel is the attention arithmetic
      // ---- S = Q K^T (32 x 32), fragments from the staged tile ----
      float s[2][4][4];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
          for (int e = 0; e < 4; ++e) s[mi][ni][e] = 0.f;
#pragma unroll
      for (int ks = 0; ks < QA_DH / 16; ++ks) {
        uint32_t a[2][4];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) {
          const __nv_bfloat16* q0 = stg + (mi * 16 + g) * QA_QK_LD + ks * 16 + 2 * t;
          const __nv_bfloat16* q1 = q0 + 8 * QA_QK_LD;
          a[mi][0] = ld_u32(q0);
          a[mi][1] = ld_u32(q1);
          a[mi][2] = ld_u32(q0 + 8);
          a[mi][3] = ld_u32(q1 + 8);
        }
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
          const __nv_bfloat16* kr = stg + (ni * 8 + g) * QA_QK_LD + QA_DH + ks * 16 + 2 * t;
          const uint32_t b0 = ld_u32(kr), b1 = ld_u32(kr + 8);
          mma_bf16_16816(s[0][ni], a[0], b0, b1);
          mma_bf16_16816(s[1][ni], a[1], b0, b1);
        }
      }
      __syncwarp();   // all lanes are done reading Q|K: the staging tile may be overwritten with V

      // ---- stage V rows [token = lane][64] ----
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const uint32_t(&v)[32] = hh ? vv1 : vv0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[8 * j + e]);
          if (bias) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + 128 + hh * 32 + 8 * j));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + 128 + hh * 32 + 8 * j + 4));
            f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w; f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
          }
          const uint32_t a = stg_u32 + static_cast<uint32_t>(lane * (QA_V_LD * 2) + (hh * 32 + 8 * j) * 2);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pack_bf16(f[0], f[1])),
                       "r"(pack_bf16(f[2], f[3])), "r"(pack_bf16(f[4], f[5])), "r"(pack_bf16(f[6], f[7]))
                       : "memory");
        }
      }

      // ---- softmax over the 32 keys (rows g and g+8 of each 16-row tile), P rounded to bf16 ----
      float inv_sum[2][2];
      uint32_t pfrag[2][2][4];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float m = -INFINITY;
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) m = fmaxf(m, fmaxf(s[mi][ni][2 * hh], s[mi][ni][2 * hh + 1]));
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
          float sum = 0.f;
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) {
            const float p0 = exp2f((s[mi][ni][2 * hh] - m) * p.scale_log2e);
            const float p1 = exp2f((s[mi][ni][2 * hh + 1] - m) * p.scale_log2e);
            s[mi][ni][2 * hh] = p0;
            s[mi][ni][2 * hh + 1] = p1;
            sum += p0 + p1;
          }
          sum += __shfl_xor_sync(0xffffffffu, sum, 1);
          sum += __shfl_xor_sync(0xffffffffu, sum, 2);
          inv_sum[mi][hh] = 1.0f / sum;
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          pfrag[mi][j][0] = pack_bf16(s[mi][2 * j][0], s[mi][2 * j][1]);
          pfrag[mi][j][1] = pack_bf16(s[mi][2 * j][2], s[mi][2 * j][3]);
          pfrag[mi][j][2] = pack_bf16(s[mi][2 * j + 1][0], s[mi][2 * j + 1][1]);
          pfrag[mi][j][3] = pack_bf16(s[mi][2 * j + 1][2], s[mi][2 * j + 1][3]);
        }
      }
      __syncwarp();   // V tile complete

      // ---- O = P V, staged as [token][64] bf16 behind the V tile, then written as one contiguous 4 KB block ----
      __nv_bfloat16* so = stg + 32 * QA_V_LD;
      const int mat = lane >> 3, r8 = lane & 7;
#pragma unroll
      for (int np = 0; np < QA_DH / 16; ++np) {
        float acc[2][2][4];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
          for (int nn = 0; nn < 2; ++nn)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[mi][nn][e] = 0.f;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          uint32_t r0, r1, r2, r3;
          const __nv_bfloat16* addr = stg + (16 * j + (mat & 1) * 8 + r8) * QA_V_LD + (2 * np + (mat >> 1)) * 8;
          asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                       : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                       : "r"(smem_u32(addr)));
#pragma unroll
          for (int mi = 0; mi < 2; ++mi) {
            mma_bf16_16816(acc[mi][0], pfrag[mi][j], r0, r1);
            mma_bf16_16816(acc[mi][1], pfrag[mi][j], r2, r3);
          }
        }
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
          for (int nn = 0; nn < 2; ++nn) {
            const int col = (2 * np + nn) * 8 + 2 * t;
            const int n0 = mi * 16 + g;
            *reinterpret_cast<uint32_t*>(so + n0 * QA_V_LD + col) =
                pack_bf16(acc[mi][nn][0] * inv_sum[mi][0], acc[mi][nn][1] * inv_sum[mi][0]);
            *reinterpret_cast<uint32_t*>(so + (n0 + 8) * QA_V_LD + col) =
                pack_bf16(acc[mi][nn][2] * inv_sum[mi][1], acc[mi][nn][3] * inv_sum[mi][1]);
          }
      }
#else
      __nv_bfloat16* so = stg + 32 * QA_V_LD;
      (void)vv0; (void)vv1; (void)g; (void)t;
#endif
      __syncwarp();
      if (live) {
        const int b = row0 >> 5;
        uint4* dst = reinterpret_cast<uint4*>(p.out + (static_cast<size_t>(b) * p.H + head) * (32 * QA_DH));
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int id = i * 32 + lane;   // 16-byte chunk of the [32][64] bf16 tile
          dst[id] = *reinterpret_cast<const uint4*>(so + (id >> 3) * QA_V_LD + (id & 7) * 8);
        }
      }
      __syncwarp();   // staging is reused by this warp's next tile
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, QA_TMEM_COLS);
  }
}

}  // namespace ldt

using namespace ldt;

extern "C" int ldt_debug_get_attention_backend(void);   // attention.cu

extern "C" int ldt_qkv_attention_bf16(int B, int H, int K, const void* A, int lda, const void* Wp, int ldw,
                                      const float* bias_p, void* out, void* stream) {
  LDT_REQUIRE(B >= 0 && H > 0 && K > 0, LDT_ERR_INVALID, "ldt_qkv_attention_bf16: bad shape B=%d H=%d K=%d", B, H, K);
  if (B == 0) return LDT_OK;
  LDT_REQUIRE(K % QA_BK == 0, LDT_ERR_INVALID, "ldt_qkv_attention_bf16: K=%d must be a multiple of %d", K, QA_BK);
  LDT_REQUIRE(lda >= K && ldw >= K && lda % 8 == 0 && ldw % 8 == 0, LDT_ERR_INVALID,
              "ldt_qkv_attention_bf16: lda=%d ldw=%d must be >= K and multiples of 8", lda, ldw);
  LDT_REQUIRE(A && Wp && out, LDT_ERR_INVALID, "ldt_qkv_attention_bf16: null pointer");
  LDT_REQUIRE((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(Wp) | reinterpret_cast<uintptr_t>(out) |
               reinterpret_cast<uintptr_t>(bias_p)) % 16 == 0,
              LDT_ERR_INVALID, "ldt_qkv_attention_bf16: pointers must be 16-byte aligned");
  const int M = B * 32;
  CUtensorMap tmA, tmW;
  int rc = make_tmap_bf16(&tmA, A, M, K, lda, 128);
  if (rc) return rc;
  rc = make_tmap_bf16(&tmW, Wp, H * QA_BN, K, ldw, QA_BN / 2);
  if (rc) return rc;
  static PerDevice<bool> attr_done;
  if (!attr_done.get()) {
    LDT_CUDA_OK(cudaFuncSetAttribute(qkv_attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, QA_SMEM_BYTES));
    LDT_CUDA_OK(cudaFuncSetAttribute(qkv_attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, QA_SMEM_BYTES));
    attr_done.get() = true;
  }
  QkvAttnParams p;
  p.M = M; p.H = H; p.bias = bias_p; p.out = static_cast<__nv_bfloat16*>(out);
  p.scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(QA_DH));
  const int tiles_m = (M + 255) / 256;
  // only as many CTA pairs as the wave count needs (see launch_tc2 in gemm.cu)
  const int tiles = tiles_m * H, max_pairs = num_sms() / 2;
  const int waves = (tiles + max_pairs - 1) / max_pairs;
  const int pairs = (tiles + waves - 1) / waves;
  if (ldt_debug_get_attention_backend() == 0)
    LDT_CUDA_OK(launch_pdl(qkv_attention_kernel<true>, dim3(2 * pairs), dim3(QA_THREADS), QA_SMEM_BYTES,
                           static_cast<cudaStream_t>(stream), tmA, tmW, p, K, tiles_m));
  else
    LDT_CUDA_OK(launch_pdl(qkv_attention_kernel<false>, dim3(2 * pairs), dim3(QA_THREADS), QA_SMEM_BYTES,
                           static_cast<cudaStream_t>(stream), tmA, tmW, p, K, tiles_m));
  return LDT_OK;
}
