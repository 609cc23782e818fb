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
#include <stdlib.h>

#include "common.cuh"
#include "gemm_epilogue.cuh"
#include "ldt_b200.h"
#include "tmap.cuh"

namespace ldt {

// ------------------------------------------------------------------------------------------------
// tcgen05 kernel
// ------------------------------------------------------------------------------------------------
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

  pdl_launch_dependents();
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
  pdl_wait();   // everything above overlapped the previous kernel's tail; from here on its results are read

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
// tcgen05 kernel, CTA-pair version (cta_group::2): a cluster of two CTAs on one TPC computes a 256 x BN tile.
// Each CTA stages its own 128 rows of A and HALF of the W tile (BN/2 rows), the leader CTA's single MMA thread
// issues tcgen05.mma.cta_group::2 (UMMA 256 x BN x 16) that reads both CTAs' shared memory and writes both CTAs'
// TMEM.  Per CTA this halves the W bytes pulled from L2 per FLOP (128 -> 256 rows of output per W row), which is
// what bounds the single-CTA kernel: 85 FLOP per L2 byte at 128x256 against ~12 TB/s of L2->SM bandwidth.
//   full[s]   (leader's copy used)  TMA bytes of BOTH CTAs land on the leader's barrier
//   empty[s]  (both copies)         tcgen05.commit multicast from the leader: the slot is free in both CTAs
//   tfull[a]  (both copies)         accumulator a complete -> both CTAs' epilogue warps
//   tempty[a] (leader's copy)       16 arrivals: 8 epilogue warps x 2 CTAs (the peer arrives remotely)
// ------------------------------------------------------------------------------------------------
// Issue loops.  The TMA-producer and MMA-issuer warps run CONVERGED (all 32 lanes take the loops, one elected lane
// issues): every address, descriptor and loop counter is then provably warp-uniform and lives in uniform registers,
// so a k-block costs the issuing warp ~20 instructions.  Written as `if (lane == 0) { loops }` the same code compiled
// to ~110 (each tcgen05.mma / TMA wrapped in an ELECT + R2UR waterfall): the issuing warps share their schedulers with
// two ALU-heavy epilogue warps each, and under that contention the long instruction chain per k-block, not the tensor
// pipe, paced the mainloop (measured: MMA thread 10.4 k clk per 256x256x1024 tile with the GELU epilogue running,
// 7.8 k without; scripts/exp_gemm_limits.py).  DBG = per-role stall counters + the load/epilogue-skipping experiments.
// TF32 = true: fp32 operands through the same pipeline (kind::tf32): a stage is still 128 bytes of K per row (32 fp32
// instead of 64 bf16) and four UMMAs of 32 bytes each, so only the tensor maps, the instruction and its descriptor differ.
template <int BN, int EPI, bool DBG, bool TF32 = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                const __grid_constant__ CUtensorMap tmO, const EpiParams p, const int K, const int tiles_m, const int tiles_n) {
  using Cfg = Tc2Cfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint8_t* stg_all = sB + STAGES * Cfg::B_BYTES;   // 8 epilogue warps x EPI_STG_BYTES, 1024-byte aligned (TMA-store swizzle atoms)
  uint64_t* full = reinterpret_cast<uint64_t*>(stg_all + 8 * EPI_STG_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  // Role index = physical warp id rotated by 4: the TMA / MMA / TMEM-alloc warps are PHYSICAL warps 8, 9, 10 and the
  // epilogue warps are physical warps 0-7.  The SM's issue arbiter prefers the highest warp id of a sub-partition, so the
  // single MMA-issuing thread must not sit below ALU-heavy epilogue warps (measured: with the MMA thread in warp 1 the
  // tensor pipe ran at 76 % of its rate under the GELU epilogue).  (role & 3) == (physical & 3): TMEM lane quadrants hold.
  const int warp = (__shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0) + 4) % 12;   // warp-uniform by construction
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();          // 0 = leader (issues the MMAs), 1 = peer
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int num_tiles = tiles_m * tiles_n;
  constexpr int KB_ELEMS = TF32 ? TC_BK / 2 : TC_BK;   // K elements per 128-byte stage row
  const int num_kb = K / KB_ELEMS;

  pdl_launch_dependents();
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
      mbar_init(&tempty[a], 16);
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();   // both CTAs' barriers are initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // everything above overlapped the previous kernel's tail; from here on its results are read

  if (warp == 0) {
    // ---- TMA producer (converged warp, one elected lane issues) ----
    const uint32_t sA0 = smem_u32(sA), sB0 = smem_u32(sB);
    const uint32_t empty0 = smem_u32(empty), full0 = smem_u32(full);
    const uint32_t full0_leader = mapa_u32(full0, 0);   // TMA bytes of both CTAs land on the leader's barrier
    const int dbg_mode = DBG ? p.dbg_mode : 0;
    const uint32_t tx_bytes = 2u * (((dbg_mode & 1) ? 0u : Cfg::A_BYTES) + ((dbg_mode & 2) ? 0u : Cfg::B_BYTES));
    int stage = 0;
    uint32_t phase = 0;
    long long t_begin = 0, t_wait = 0;
    if constexpr (DBG) t_begin = clock64();
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int m0 = (tile % tiles_m) * T2_BM + static_cast<int>(rank) * 128;
      const int n0 = (tile / tiles_m) * BN + static_cast<int>(rank) * (BN / 2);
      for (int kb = 0; kb < num_kb; ++kb) {
        long long w0 = 0;
        if constexpr (DBG) w0 = clock64();
        mbar_wait_u32(empty0 + stage * 8, phase ^ 1u);
        if constexpr (DBG) t_wait += clock64() - w0;
        if (elect_one()) {
          if (rank == 0) mbar_expect_tx_u32(full0 + stage * 8, tx_bytes);
          if (!DBG || !(dbg_mode & 1))
            tma_load_2d_pair_u32(sA0 + stage * Cfg::A_BYTES, &tmA, full0_leader + stage * 8, kb * KB_ELEMS, m0);
          if (!DBG || !(dbg_mode & 2))
            tma_load_2d_pair_u32(sB0 + stage * Cfg::B_BYTES, &tmW, full0_leader + stage * 8, kb * KB_ELEMS, n0);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
    if constexpr (DBG) {
      if (p.dbg && lane == 0) {
        p.dbg[blockIdx.x * 8 + 5] = t_wait;
        p.dbg[blockIdx.x * 8 + 6] = clock64() - t_begin;
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer (leader CTA; converged warp, one elected lane issues) ----
    if (rank == 0) {
      constexpr uint32_t idesc = TF32 ? umma_idesc_tf32(T2_BM, BN) : umma_idesc_bf16(T2_BM, BN);
      const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty), tfull0 = smem_u32(tfull), tempty0 = smem_u32(tempty);
      // descriptors of (stage 0, k-slice 0); the start-address field counts 16-byte units, so later stages / k-slices
      // are plain additions (shared-memory addresses stay below 256 KB: no carry out of the 14-bit field)
      const uint64_t descA0 = umma_desc_k_sw128(smem_u32(sA));
      const uint64_t descB0 = umma_desc_k_sw128(smem_u32(sB));
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      long long t_begin = 0, t_full = 0, t_tempty = 0, ntile = 0;
      if constexpr (DBG) t_begin = clock64();
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        long long w0 = 0;
        if constexpr (DBG) w0 = clock64();
        mbar_wait_u32(tempty0 + acc * 8, acc_phase ^ 1u);
        if constexpr (DBG) { t_tempty += clock64() - w0; ++ntile; }
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * Cfg::ACC_STRIDE);
        for (int kb = 0; kb < num_kb; ++kb) {
          if constexpr (DBG) w0 = clock64();
          mbar_wait_u32(full0 + stage * 8, phase);
          if constexpr (DBG) t_full += clock64() - w0;
          tc_fence_after();
          if (elect_one()) {
            const uint64_t da = descA0 + static_cast<uint64_t>(stage * (Cfg::A_BYTES >> 4));
            const uint64_t db = descB0 + static_cast<uint64_t>(stage * (Cfg::B_BYTES >> 4));
            if constexpr (TF32) {
              umma_tf32_ss_pair(tmem_d, da, db, idesc, kb != 0 ? 1u : 0u);
#pragma unroll
              for (int k = 1; k < TC_BK / 16; ++k) umma_tf32_ss_pair(tmem_d, da + 2 * k, db + 2 * k, idesc, 1u);
            } else {
              umma_bf16_ss_pair(tmem_d, da, db, idesc, kb != 0 ? 1u : 0u);
#pragma unroll
              for (int k = 1; k < TC_BK / 16; ++k) umma_bf16_ss_pair_acc(tmem_d, da + 2 * k, db + 2 * k, idesc);
            }
            umma_commit_pair_u32(empty0 + stage * 8, 0x3);   // frees the stage in both CTAs
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        if (elect_one()) umma_commit_pair_u32(tfull0 + acc * 8, 0x3);   // accumulator complete -> both CTAs' epilogues
        __syncwarp();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
      if constexpr (DBG) {
        if (p.dbg && lane == 0) {
          p.dbg[blockIdx.x * 8 + 0] = clock64() - t_begin;
          p.dbg[blockIdx.x * 8 + 1] = t_full;
          p.dbg[blockIdx.x * 8 + 2] = t_tempty;
          p.dbg[blockIdx.x * 8 + 7] = ntile;
        }
      }
    }
  } else if (warp >= TC_EPI_WARP0) {
    const int quad = warp & 3;
    const int half = (warp - TC_EPI_WARP0) >> 2;
    uint8_t* stg = stg_all + (warp - TC_EPI_WARP0) * EPI_STG_BYTES;
    const uint32_t tempty0_leader = mapa_u32(smem_u32(tempty), 0);
    constexpr bool kBf16Out = (EPI == LDT_EPI_BIAS_BF16 || EPI == LDT_EPI_BIAS_GELU_BF16);
    const CUtensorMap* tm_out = (kBf16Out && p.tma_store) ? &tmO : nullptr;
    int acc = 0;
    uint32_t acc_phase = 0;
    long long t_begin = 0, t_tfull = 0;
    if constexpr (DBG) t_begin = clock64();
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int m0 = (tile % tiles_m) * T2_BM + static_cast<int>(rank) * 128;
      const int n0 = (tile / tiles_m) * BN;
      if (DBG && (p.dbg_mode & 4)) {   // experiment: mainloop only
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
      } else {
        const int col = half * (BN / 2);
        const uint32_t taddr =
            tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * Cfg::ACC_STRIDE + col);
        auto wait_acc = [&]() {
          long long w0 = 0;
          if constexpr (DBG) w0 = clock64();
          mbar_wait(&tfull[acc], acc_phase);
          if constexpr (DBG) t_tfull += clock64() - w0;
          tc_fence_after();
        };
        if constexpr (DBG && EPI == LDT_EPI_BIAS_GELU_BF16) {   // epilogue with parts removed at compile time (dbg_mode >> 3)
          switch ((p.dbg_mode & 255) >> 3) {
            case 1: epilogue_staged<EPI, BN / 2, 8>(p, stg, lane, m0 + quad * 32, n0 + col, taddr, wait_acc, tm_out); break;
            case 2: epilogue_staged<EPI, BN / 2, 16>(p, stg, lane, m0 + quad * 32, n0 + col, taddr, wait_acc, tm_out); break;
            case 4: epilogue_staged<EPI, BN / 2, 32>(p, stg, lane, m0 + quad * 32, n0 + col, taddr, wait_acc, tm_out); break;
            case 8: epilogue_staged<EPI, BN / 2, 64>(p, stg, lane, m0 + quad * 32, n0 + col, taddr, wait_acc, tm_out); break;
            case 6: epilogue_staged<EPI, BN / 2, 48>(p, stg, lane, m0 + quad * 32, n0 + col, taddr, wait_acc, tm_out); break;
            case 7: epilogue_staged<EPI, BN / 2, 56>(p, stg, lane, m0 + quad * 32, n0 + col, taddr, wait_acc, tm_out); break;
            case 14: epilogue_staged<EPI, BN / 2, 112>(p, stg, lane, m0 + quad * 32, n0 + col, taddr, wait_acc, tm_out); break;
            case 12: epilogue_staged<EPI, BN / 2, 96>(p, stg, lane, m0 + quad * 32, n0 + col, taddr, wait_acc, tm_out); break;
            default: epilogue_staged<EPI, BN / 2>(p, stg, lane, m0 + quad * 32, n0 + col, taddr, wait_acc, tm_out); break;
          }
        } else {
          epilogue_staged<EPI, BN / 2>(p, stg, lane, m0 + quad * 32, n0 + col, taddr, wait_acc, tm_out);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty0_leader + acc * 8);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (tm_out != nullptr && lane == 0) tma_store_wait_read();   // the staging buffer outlives the last bulk store's READ; the writes complete with the grid
    if constexpr (DBG) {
      if (p.dbg && warp == TC_EPI_WARP0 && lane == 0) {
        p.dbg[blockIdx.x * 8 + 3] = clock64() - t_begin;
        p.dbg[blockIdx.x * 8 + 4] = t_tfull;
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();   // the peer's TMEM / shared memory stay valid until the leader's last MMA has retired
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// Cluster-of-4 version: two CTA pairs that work on vertically adjacent 256-row tiles of the SAME column tile share the
// weight tile.  Each of the four CTAs fetches one quarter of the W tile (64 rows x 64 k) and TMA-multicasts it to the
// CTA with the same rank-in-pair in the other pair, so a W byte crosses the L2->SM fabric once per 512 output rows
// instead of once per 256: 48 KB instead of 64 KB per pair and k-block.  At the board's power cap the operand traffic
// is a first-order energy term (scripts/exp_sustained.py: the pair kernel with half of its loads removed runs 5-22 %
// faster), so fewer bytes is more clock.
//   full[s]   leader of each pair   own pair's A (2 x 16 KB) + four W quarters landing in the pair (4 x 8 KB)
//   empty[s]  every CTA, count 2    BOTH pairs' MMAs have released the stage (the sibling multicasts into our smem)
//   tfull[a]  both CTAs of a pair   accumulator complete;  tempty[a]: leader of the pair, 16 arrivals
// Grids are whole clusters of 4 (33-34 co-resident on B200's 18/20-SM GPCs: 132-136 of 148 SMs), and an N = 1024 GEMM
// at M = 8192 becomes 64 super-tiles = 1.94 waves instead of 1.73 waves of pair tiles rounded up to 2.
// ------------------------------------------------------------------------------------------------
template <int BN, int EPI>
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(TC_THREADS, 1)
gemm_tc4_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const EpiParams p,
                const int K, const int tiles_m, const int tiles_n) {
  using Cfg = Tc2Cfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int Q_BYTES = Cfg::B_BYTES / 2;   // one quarter of the W tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint8_t* stg_all = sB + STAGES * Cfg::B_BYTES;   // 8 epilogue warps x EPI_STG_BYTES, 1024-byte aligned (TMA-store swizzle atoms)
  uint64_t* full = reinterpret_cast<uint64_t*>(stg_all + 8 * EPI_STG_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = (__shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0) + 4) % 12;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();   // 0..3
  const uint32_t q = crank >> 1;              // pair within the cluster = which of the two stacked 256-row tiles
  const uint32_t h = crank & 1;               // rank within the pair (0 = leader, issues the MMAs)
  const uint32_t lead = crank & ~1u;          // cluster rank of this pair's leader
  const int cluster = blockIdx.x >> 2;
  const int num_clusters = gridDim.x >> 2;
  const int tiles_m2 = (tiles_m + 1) >> 1;
  const int num_super = tiles_m2 * tiles_n;
  const int num_kb = K / TC_BK;

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 2);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 16);
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ---- TMA producer ----
    const uint32_t sA0 = smem_u32(sA), sBq = smem_u32(sB) + q * Q_BYTES;
    const uint32_t empty0 = smem_u32(empty), full0 = smem_u32(full);
    const uint32_t full0_lead = mapa_u32(full0, lead);
    const uint16_t wmask = static_cast<uint16_t>((1u << h) | (1u << (h + 2)));   // same rank-in-pair, both pairs
    int stage = 0;
    uint32_t phase = 0;
    for (int st = cluster; st < num_super; st += num_clusters) {
      const int m0 = ((st % tiles_m2) * 2 + static_cast<int>(q)) * T2_BM + static_cast<int>(h) * 128;
      const int n0 = (st / tiles_m2) * BN + static_cast<int>(h) * (BN / 2) + static_cast<int>(q) * (BN / 4);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait_u32(empty0 + stage * 8, phase ^ 1u);
        if (elect_one()) {
          if (h == 0) mbar_expect_tx_u32(full0 + stage * 8, 2 * (Cfg::A_BYTES + Cfg::B_BYTES));
          tma_load_2d_pair_u32(sA0 + stage * Cfg::A_BYTES, &tmA, full0_lead + stage * 8, kb * TC_BK, m0);
          tma_load_2d_pair_mcast_u32(sBq + stage * Cfg::B_BYTES, &tmW, full0_lead + stage * 8, kb * TC_BK, n0, wmask);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer (leader CTA of each pair) ----
    if (h == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(T2_BM, BN);
      const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty), tfull0 = smem_u32(tfull), tempty0 = smem_u32(tempty);
      const uint64_t descA0 = umma_desc_k_sw128(smem_u32(sA));
      const uint64_t descB0 = umma_desc_k_sw128(smem_u32(sB));
      const uint16_t pair_mask = static_cast<uint16_t>(0x3u << lead);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int st = cluster; st < num_super; st += num_clusters) {
        mbar_wait_u32(tempty0 + acc * 8, acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * Cfg::ACC_STRIDE);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait_u32(full0 + stage * 8, phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t da = descA0 + static_cast<uint64_t>(stage * (Cfg::A_BYTES >> 4));
            const uint64_t db = descB0 + static_cast<uint64_t>(stage * (Cfg::B_BYTES >> 4));
            umma_bf16_ss_pair(tmem_d, da, db, idesc, kb != 0 ? 1u : 0u);
#pragma unroll
            for (int k = 1; k < TC_BK / 16; ++k) umma_bf16_ss_pair_acc(tmem_d, da + 2 * k, db + 2 * k, idesc);
            umma_commit_pair_u32(empty0 + stage * 8, 0xF);   // this pair is done with the stage: tell all four CTAs
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        if (elect_one()) umma_commit_pair_u32(tfull0 + acc * 8, pair_mask);
        __syncwarp();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp >= TC_EPI_WARP0) {
    const int quad = warp & 3;
    const int half = (warp - TC_EPI_WARP0) >> 2;
    uint8_t* stg = stg_all + (warp - TC_EPI_WARP0) * EPI_STG_BYTES;
    const uint32_t tempty0_lead = mapa_u32(smem_u32(tempty), lead);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int st = cluster; st < num_super; st += num_clusters) {
      const int m0 = ((st % tiles_m2) * 2 + static_cast<int>(q)) * T2_BM + static_cast<int>(h) * 128;
      const int col = (st / tiles_m2) * BN + half * (BN / 2);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                             static_cast<uint32_t>(acc * Cfg::ACC_STRIDE + half * (BN / 2));
      epilogue_staged<EPI, BN / 2>(p, stg, lane, m0 + quad * 32, col, taddr, [&]() {
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
      });
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty0_lead + acc * 8);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }
  tc_fence_before();
  cluster_sync_all();   // nobody leaves while a sibling may still multicast into its shared memory / read its TMEM
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
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
int make_tmap_bf16(CUtensorMap* tm, const void* base, int rows, int cols, int ld, int box_rows) {
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

// f32 [rows, ld] row-major; tile box = box_rows x 32 columns (128 bytes), 128-byte swizzle, OOB reads give zero.
int make_tmap_f32(CUtensorMap* tm, const void* base, int rows, int cols, int ld, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return LDT_ERR_CUDA;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * 4};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(TC_BK / 2), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled(f32) failed: CUresult %d (base=%p rows=%d cols=%d ld=%d box_rows=%d)", (int)r, base,
                   rows, cols, ld, box_rows);
    return LDT_ERR_CUDA;
  }
  return LDT_OK;
}

// TF32 parity mode: the CTA-pair kernel with fp32 operands (kind::tf32).
template <int BN, int EPI>
static int launch_tc2_tf32(const ldt_gemm_args& a, const EpiParams& p, cudaStream_t s) {
  using Cfg = Tc2Cfg<BN>;
  CUtensorMap tmA, tmW;
  int rc = make_tmap_f32(&tmA, a.A, a.M, a.K, a.lda, 128);
  if (rc) return rc;
  rc = make_tmap_f32(&tmW, a.W, a.N, a.K, a.ldw, BN / 2);
  if (rc) return rc;
  static PerDevice<bool> attr_done;
  if (!attr_done.get()) {
    LDT_CUDA_OK(cudaFuncSetAttribute(gemm_tc2_kernel<BN, EPI, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done.get() = true;
  }
  const int tiles_m = (a.M + T2_BM - 1) / T2_BM;
  const int tiles_n = (a.N + BN - 1) / BN;
  const int tiles = tiles_m * tiles_n, max_pairs = num_sms() / 2;
  const int waves = (tiles + max_pairs - 1) / max_pairs;
  const int pairs = (tiles + waves - 1) / waves;
  LDT_CUDA_OK(launch_pdl(gemm_tc2_kernel<BN, EPI, false, true>, dim3(2 * pairs), dim3(TC_THREADS), Cfg::SMEM_BYTES, s, tmA, tmW, tmA, p,
                         a.K, tiles_m, tiles_n));
  return LDT_OK;
}

template <int EPI>
static int launch_any_tf32(const ldt_gemm_args& a, const EpiParams& p, cudaStream_t s) {
  if (a.N % 256 == 0) return launch_tc2_tf32<256, EPI>(a, p, s);
  return launch_tc2_tf32<128, EPI>(a, p, s);
}

template <int BN, int EPI>
static int launch_tc(const ldt_gemm_args& a, const EpiParams& p, cudaStream_t s) {
  using Cfg = TcCfg<BN>;
  CUtensorMap tmA, tmW;
  int rc = make_tmap_bf16(&tmA, a.A, a.M, a.K, a.lda, TC_BM);
  if (rc) return rc;
  rc = make_tmap_bf16(&tmW, a.W, a.N, a.K, a.ldw, BN);
  if (rc) return rc;
  static PerDevice<bool> attr_done;
  if (!attr_done.get()) {
    LDT_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done.get() = true;
  }
  const int tiles_m = (a.M + TC_BM - 1) / TC_BM;
  const int tiles_n = (a.N + BN - 1) / BN;
  const int grid = min(tiles_m * tiles_n, num_sms());
  LDT_CUDA_OK(launch_pdl(gemm_tc_kernel<BN, EPI>, dim3(grid), dim3(TC_THREADS), Cfg::SMEM_BYTES, s, tmA, tmW, p, a.K, tiles_m,
                         tiles_n));
  return LDT_OK;
}

template <int BN, int EPI>
static int launch_tc2(const ldt_gemm_args& a, const EpiParams& p, cudaStream_t s) {
  using Cfg = Tc2Cfg<BN>;
  CUtensorMap tmA, tmW;
  int rc = make_tmap_bf16(&tmA, a.A, a.M, a.K, a.lda, 128);
  if (rc) return rc;
  rc = make_tmap_bf16(&tmW, a.W, a.N, a.K, a.ldw, BN / 2);
  if (rc) return rc;
  static PerDevice<bool> attr_done;
  if (!attr_done.get()) {
    LDT_CUDA_OK(cudaFuncSetAttribute(gemm_tc2_kernel<BN, EPI, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    LDT_CUDA_OK(cudaFuncSetAttribute(gemm_tc2_kernel<BN, EPI, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done.get() = true;
  }
  const int tiles_m = (a.M + T2_BM - 1) / T2_BM;
  const int tiles_n = (a.N + BN - 1) / BN;
  // Only as many CTA pairs as the wave count needs: 128 tiles are two waves on 64 pairs exactly as on 74, and at the board's
  // power cap the ten pairs that would idle through the second wave are clock for the others.
  const int tiles = tiles_m * tiles_n, max_pairs = num_sms() / 2;
  const int waves = (tiles + max_pairs - 1) / max_pairs;
  const int pairs = (p.dbg_mode & 512) ? min(tiles, max_pairs) : (tiles + waves - 1) / waves;
  CUtensorMap tmO = tmA;   // only read by the bf16-output epilogues
  if (p.tma_store && (EPI == LDT_EPI_BIAS_BF16 || EPI == LDT_EPI_BIAS_GELU_BF16)) {
    rc = make_tmap_bf16(&tmO, a.out, a.M, a.N, a.ldo, 32);   // store boxes: 32 rows x 64 columns, 128-byte swizzle
    if (rc) return rc;
  }
  if (p.dbg != nullptr || (p.dbg_mode & 255) != 0)   // diagnostics build of the same kernel (counters, skipped loads / epilogue)
    LDT_CUDA_OK(launch_pdl(gemm_tc2_kernel<BN, EPI, true>, dim3(2 * pairs), dim3(TC_THREADS), Cfg::SMEM_BYTES, s, tmA, tmW, tmO, p,
                           a.K, tiles_m, tiles_n));
  else
    LDT_CUDA_OK(launch_pdl(gemm_tc2_kernel<BN, EPI, false>, dim3(2 * pairs), dim3(TC_THREADS), Cfg::SMEM_BYTES, s, tmA, tmW, tmO, p,
                           a.K, tiles_m, tiles_n));
  return LDT_OK;
}

template <int BN, int EPI>
static int launch_tc4(const ldt_gemm_args& a, const EpiParams& p, cudaStream_t s) {
  using Cfg = Tc2Cfg<BN>;
  static PerDevice<int> max_clusters_dev;
  int& max_clusters = max_clusters_dev.get_or(-1);
  if (max_clusters < 0) {
    LDT_CUDA_OK(cudaFuncSetAttribute(gemm_tc4_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((num_sms() / 4) * 4);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 4; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    LDT_CUDA_OK(cudaOccupancyMaxActiveClusters(&n, gemm_tc4_kernel<BN, EPI>, &cfg));
    max_clusters = n;
  }
  LDT_REQUIRE(max_clusters > 0, LDT_ERR_UNSUPPORTED, "ldt_gemm_bf16: no cluster of 4 CTAs fits on this device");
  CUtensorMap tmA, tmW;
  int rc = make_tmap_bf16(&tmA, a.A, a.M, a.K, a.lda, 128);
  if (rc) return rc;
  rc = make_tmap_bf16(&tmW, a.W, a.N, a.K, a.ldw, BN / 4);
  if (rc) return rc;
  const int tiles_m = (a.M + T2_BM - 1) / T2_BM;
  const int tiles_n = (a.N + BN - 1) / BN;
  const int clusters = min(((tiles_m + 1) / 2) * tiles_n, max_clusters);
  LDT_CUDA_OK(launch_pdl(gemm_tc4_kernel<BN, EPI>, dim3(4 * clusters), dim3(TC_THREADS), Cfg::SMEM_BYTES, s, tmA, tmW, p, a.K,
                         tiles_m, tiles_n));
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
  if (a.backend == 4) {   // clusters of two CTA pairs sharing the W tile by TMA multicast
    LDT_REQUIRE(a.N % 256 == 0, LDT_ERR_UNSUPPORTED, "ldt_gemm_bf16: backend 4 needs N %% 256 == 0 (N=%d)", a.N);
    return launch_tc4<256, EPI>(a, p, s);
  }
  // backend 0 picks: CTA pairs (256-row tiles) once there is a full tile of rows, else single-CTA tiles.
  static int pair_min_m = -1;
  if (pair_min_m < 0) {
    // smallest M that goes to the CTA-pair kernel.  One full 256-row tile is enough: at M = 512 (batch 16, BASELINE
    // configs[0]) a token pass takes 1.31 ms on the pair kernel against 2.27 ms on single-CTA tiles, whose epilogue stores
    // are not staged (1.70 vs 2.27 ms at M = 256).  LDT_PAIR_MIN_M overrides it for experiments.
    const char* e = getenv("LDT_PAIR_MIN_M");
    pair_min_m = e ? atoi(e) : 256;
  }
  const bool pair_ok = (a.backend == 3) || (a.backend == 0 && a.M >= pair_min_m);
  if (pair_ok && a.backend != 2) {
    if (a.N % 256 == 0) {
      // 256-wide tiles unless they leave most of the machine idle: at M = 2048 (64 clouds per GPU, the completion
      // workload) an N = 1024 GEMM is 32 tiles for 74 CTA pairs; 128-wide tiles double the tile count at the same
      // mainloop efficiency per column (A is re-read per tile either way)
      const int pairs = num_sms() / 2;
      const int tm = (a.M + T2_BM - 1) / T2_BM;
      const int t256 = tm * (a.N / 256), t128 = tm * (a.N / 128);
      const double u256 = static_cast<double>(t256) / (((t256 + pairs - 1) / pairs) * pairs);
      const double u128 = static_cast<double>(t128) / (((t128 + pairs - 1) / pairs) * pairs);
      if (u128 > u256 + 0.1) return launch_tc2<128, EPI>(a, p, s);
      if ((p.dbg_mode & 1024) && a.N == 1024 && a.K == 1024) return launch_tc2<128, EPI>(a, p, s);   // experiment: fc_o on 256 x 128 tiles
      if ((p.dbg_mode & 2048) && a.N == 1024 && a.K == 4096) return launch_tc2<128, EPI>(a, p, s);   // experiment: fc2 on 256 x 128 tiles
      return launch_tc2<256, EPI>(a, p, s);
    }
    return launch_tc2<128, EPI>(a, p, s);
  }
  if (a.N % 256 == 0) return launch_tc<256, EPI>(a, p, s);
  return launch_tc<128, EPI>(a, p, s);
}

}  // namespace ldt

using namespace ldt;

static unsigned long long* g_gemm_dbg = nullptr;
static int g_gemm_dbg_mode = 0;
extern "C" int ldt_debug_get_gemm_mode() { return g_gemm_dbg_mode; }
extern "C" int ldt_debug_set_gemm_mode(int mode) {
  g_gemm_dbg_mode = mode;
  return LDT_OK;
}
extern "C" int ldt_debug_set_gemm_counters(unsigned long long* dev_buf) {
  g_gemm_dbg = dev_buf;
  return LDT_OK;
}

extern "C" int ldt_gemm_bf16(const ldt_gemm_args* args, void* stream) {
  LDT_REQUIRE(args != nullptr, LDT_ERR_INVALID, "ldt_gemm_bf16: null args");
  const ldt_gemm_args& a = *args;
  LDT_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, LDT_ERR_INVALID, "ldt_gemm_bf16: bad shape M=%d N=%d K=%d", a.M, a.N, a.K);
  const bool tf32 = a.operand_type == 1 || a.operand_type == 2;
  LDT_REQUIRE(a.operand_type >= 0 && a.operand_type <= 2, LDT_ERR_INVALID, "ldt_gemm_bf16: unknown operand_type %d", a.operand_type);
  if (tf32) {
    LDT_REQUIRE(a.K % (TC_BK / 2) == 0, LDT_ERR_INVALID, "ldt_gemm_bf16: K=%d must be a multiple of %d for f32 operands", a.K, TC_BK / 2);
    LDT_REQUIRE(a.backend == 0 || a.backend == 3, LDT_ERR_UNSUPPORTED, "ldt_gemm_bf16: f32 operands run on the CTA-pair kernel only");
    LDT_REQUIRE(a.lda % 4 == 0 && a.ldw % 4 == 0, LDT_ERR_INVALID, "ldt_gemm_bf16: lda=%d ldw=%d must be multiples of 4 for f32 operands", a.lda, a.ldw);
  }
  LDT_REQUIRE(tf32 || a.K % TC_BK == 0, LDT_ERR_INVALID, "ldt_gemm_bf16: K=%d must be a multiple of %d (pad with zeros)", a.K, TC_BK);
  LDT_REQUIRE(a.N % 8 == 0, LDT_ERR_INVALID, "ldt_gemm_bf16: N=%d must be a multiple of 8", a.N);
  LDT_REQUIRE(a.lda >= a.K && a.ldw >= a.K && (tf32 || (a.lda % 8 == 0 && a.ldw % 8 == 0)), LDT_ERR_INVALID,
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
  p.dbg = g_gemm_dbg;
  p.dbg_mode = g_gemm_dbg_mode;
  p.tma_store = (g_gemm_dbg_mode & 256) ? 0 : 1;
  p.relu = 0;
  p.f32_plain = a.operand_type == 2;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int epilogue = a.epilogue;
  if (epilogue == LDT_EPI_BIAS_RELU_F32) {   // the same kernels as the plain f32 epilogues, ReLU as a run-time flag
    epilogue = LDT_EPI_BIAS_F32;
    p.relu = 1;
  } else if (epilogue == LDT_EPI_RESID_RELU_F32) {
    LDT_REQUIRE(a.gate == nullptr, LDT_ERR_INVALID, "ldt_gemm_bf16: LDT_EPI_RESID_RELU_F32 takes no gate");
    epilogue = LDT_EPI_GATE_RESID_F32;
    p.relu = 1;
  }
  if (tf32) {
    switch (epilogue) {
      case LDT_EPI_BIAS_F32: return launch_any_tf32<LDT_EPI_BIAS_F32>(a, p, s);
      case LDT_EPI_BIAS_GELU_F32: return launch_any_tf32<LDT_EPI_BIAS_GELU_F32>(a, p, s);
      case LDT_EPI_GATE_RESID_F32:
        LDT_REQUIRE(a.resid != nullptr, LDT_ERR_INVALID, "ldt_gemm_bf16: residual epilogue needs resid");
        LDT_REQUIRE(a.gate == nullptr || a.gate_stride % 4 == 0, LDT_ERR_INVALID, "ldt_gemm_bf16: gate_stride must be a multiple of 4");
        return launch_any_tf32<LDT_EPI_GATE_RESID_F32>(a, p, s);
      default:
        set_last_error("ldt_gemm_bf16: epilogue %d is not available for f32 operands (0, 3, 4, 5, 6 are)", a.epilogue);
        return LDT_ERR_UNSUPPORTED;
    }
  }
  LDT_REQUIRE(a.epilogue != LDT_EPI_BIAS_GELU_F32, LDT_ERR_UNSUPPORTED, "ldt_gemm_bf16: LDT_EPI_BIAS_GELU_F32 needs operand_type 1");
  switch (epilogue) {
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
