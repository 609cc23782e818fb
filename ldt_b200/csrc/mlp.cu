// Fused MLP half of a transformer block:  out = resid + gate * (GELU(A . W1^T + b1) . W2^T + b2)   in ONE kernel.
//
// Replaces MLP.forward (model/layers.py:110-133: Conv1d(C->4C), exact-erf GELU, Conv1d(4C->C)) plus the gated residual
// add of ResidualBlock.forward (layers.py:219), which gemm.cu runs as two launches (fc1+GELU, fc2+gate+residual).
//
// Why one kernel.  With 256x256 CTA-pair tiles and M = 8192 rows, fc1 is 512 tiles and fc2 is 128 tiles of four times
// the K depth.  As separate persistent launches on 74 CTA pairs that is 7 + 2 waves = 15 tile-units of mainloop (one
// unit = one 256x256x1024 tile), each launch ending in an exposed epilogue and a launch gap, where 13.84 units of
// tensor work exist.  Here every pair runs ONE static work list: its fc1 tiles first (no dependencies), then its fc2
// tiles, and the lists are balanced in k-blocks: pairs that own one fc2 tile fewer take kb2/kb1 more fc1 tiles (14
// units per pair at M = 8192).  fc2 tiles of an m-block (256 rows) may start once the 16 fc1 tiles of that m-block
// have been stored: epilogue warps publish per-m-block counters in global memory (release), the TMA producer of a
// dependent tile acquires them and crosses into the async proxy before loading the hidden activations.  The hidden
// activations still make one bf16 round trip through L2; what disappears is one launch, one exposed epilogue tail and
// the wave quantisation of both GEMMs.
//
// Deadlock freedom: every pair finishes all of its fc1 tiles before it waits for anything, fc1 tiles wait for nothing,
// and the grid is one CTA per SM, all co-resident (same requirement as a cooperative launch; the launcher checks it).
// A watchdog turns a wait longer than ~2^31 cycles into a trap instead of a hang.
#include <cuda.h>

#include "common.cuh"
#include "gemm_epilogue.cuh"
#include "ldt_b200.h"
#include "tmap.cuh"

extern "C" int ldt_debug_get_gemm_mode();

namespace ldt {

constexpr int MLP_BN = 256;

struct MlpSched {
  int P;         // CTA pairs of the grid
  int T1, T2;    // tiles of phase 1 (fc1) / phase 2 (fc2)
  int tn1, tn2;  // column tiles per m-block
  int kb1, kb2;  // k-blocks (64 wide) per tile
  int a, b;      // phase-1 slots of a heavy / light pair
  int r2, q2;    // pairs 0..r2-1 ("heavy") own q2+1 phase-2 tiles, the others ("light") q2
};

// Static schedule.  Phase-1 tiles are numbered m-block-major and dealt out slot by slot (slot s of pair p, all pairs in
// slots < a, only the light pairs in slots a..b-1), so m-blocks complete in ascending order.  Phase-2 tiles are also
// m-block-major and are dealt out in order of their start time: first tiles of the heavy pairs, then rounds of P.
__host__ __device__ inline int mlp_num_items(const MlpSched& S, int p) {
  const bool heavy = p < S.r2;
  return (heavy ? S.a : S.b) + S.q2 + (heavy ? 1 : 0);
}
__host__ __device__ inline bool mlp_item(const MlpSched& S, int p, int it, int& phase, int& mb, int& nt) {
  const bool heavy = p < S.r2;
  const int n1 = heavy ? S.a : S.b;
  if (it < n1) {
    const int idx = (it < S.a) ? it * S.P + p : S.a * S.P + (it - S.a) * (S.P - S.r2) + (p - S.r2);
    if (idx >= S.T1) return false;   // surplus slot
    phase = 1;
    mb = idx / S.tn1;
    nt = idx % S.tn1;
    return true;
  }
  const int j = it - n1;
  const int t2 = heavy ? (j == 0 ? p : S.r2 + (j - 1) * S.P + p) : S.r2 + j * S.P + p;
  phase = 2;
  mb = t2 / S.tn2;
  nt = t2 % S.tn2;
  return true;
}

MlpSched mlp_make_sched(int tiles_m, int tn1, int tn2, int kb1, int kb2, int P) {
  MlpSched S;
  S.P = P;
  S.T1 = tiles_m * tn1;
  S.T2 = tiles_m * tn2;
  S.tn1 = tn1; S.tn2 = tn2; S.kb1 = kb1; S.kb2 = kb2;
  S.q2 = S.T2 / P;
  S.r2 = S.T2 % P;
  if (S.r2 == 0) {
    S.a = S.b = (S.T1 + P - 1) / P;
  } else {
    // d extra phase-1 slots on the light pairs give every pair the same number of k-blocks
    const int d = (kb2 % kb1 == 0) ? kb2 / kb1 : 0;
    const long long num = static_cast<long long>(S.T1) - static_cast<long long>(P - S.r2) * d;
    S.a = num <= 0 ? 0 : static_cast<int>((num + P - 1) / P);
    S.b = S.a + d;
  }
  return S;
}

struct MlpParams {
  EpiParams e1, e2;
  unsigned int* sync;          // [0, tiles_m): stored-fc1-tile counters per m-block; [tiles_m, 2 tiles_m): consumer counters
  unsigned int ready_target;   // tn1 tiles x 16 epilogue warps (8 per CTA of the pair)
  unsigned int done_target;    // tn2 tiles x 2 CTAs
  int tiles_m;
  int nosync;                  // experiments only (ldt_debug_set_gemm_mode bit 3): ignore the counters (WRONG results)
};

// generic-proxy accesses to global memory before / after this fence are ordered against async-proxy (TMA) accesses
// after / before it; compiles to a view fence only (no memory barrier)
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

__device__ __forceinline__ void wait_counter_ge(const unsigned int* ctr, unsigned int target) {
  const long long t0 = clock64();
  for (;;) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    if (v >= target) return;
    __nanosleep(100);
    if (clock64() - t0 > (1ll << 31)) __trap();   // watchdog: a lost dependency must not hang the GPU
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
mlp_tc2_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmW1,
               const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmW2,
               const __grid_constant__ CUtensorMap tmH, const MlpParams p, const MlpSched S) {
  using Cfg = Tc2Cfg<MLP_BN>;
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

  // same role layout as gemm_tc2_kernel (gemm.cu): physical warps 0-7 epilogue, 8 TMA, 9 MMA, 10 TMEM allocator
  const int warp = (__shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0) + 4) % 12;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int nit = mlp_num_items(S, pair);

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmW2);
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
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ---- TMA producer ----
    const uint32_t sA0 = smem_u32(sA), sB0 = smem_u32(sB);
    const uint32_t empty0 = smem_u32(empty), full0 = smem_u32(full);
    const uint32_t full0_leader = mapa_u32(full0, 0);
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < nit; ++it) {
      int ph, mb, nt;
      if (!mlp_item(S, pair, it, ph, mb, nt)) continue;
      const int m0 = mb * T2_BM + static_cast<int>(rank) * 128;
      const int n0 = nt * MLP_BN + static_cast<int>(rank) * (MLP_BN / 2);
      const CUtensorMap* ta = (ph == 1) ? &tmA1 : &tmA2;
      const CUtensorMap* tw = (ph == 1) ? &tmW1 : &tmW2;
      const int num_kb = (ph == 1) ? S.kb1 : S.kb2;
      if (ph == 2) {
        // the hidden activations of this m-block: all fc1 tiles stored (acquire), then generic -> async proxy
        wait_counter_ge(p.sync + mb, p.ready_target);
        fence_proxy_async_global();
        if (elect_one() && !p.nosync) {
          // self-cleaning: the last of the m-block's consumers (every one of them is past its wait, no producer is
          // left) zeroes both counters for the next launch
          const unsigned int old = atomicAdd(p.sync + p.tiles_m + mb, 1u);
          if (old + 1u == p.done_target) {
            p.sync[p.tiles_m + mb] = 0u;
            p.sync[mb] = 0u;
          }
        }
        __syncwarp();
      }
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait_u32(empty0 + stage * 8, phase ^ 1u);
        if (elect_one()) {
          if (rank == 0) mbar_expect_tx_u32(full0 + stage * 8, 2 * (Cfg::A_BYTES + Cfg::B_BYTES));
          tma_load_2d_pair_u32(sA0 + stage * Cfg::A_BYTES, ta, full0_leader + stage * 8, kb * TC_BK, m0);
          tma_load_2d_pair_u32(sB0 + stage * Cfg::B_BYTES, tw, full0_leader + stage * 8, kb * TC_BK, n0);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer (leader CTA) ----
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(T2_BM, MLP_BN);
      const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty), tfull0 = smem_u32(tfull), tempty0 = smem_u32(tempty);
      const uint64_t descA0 = umma_desc_k_sw128(smem_u32(sA));
      const uint64_t descB0 = umma_desc_k_sw128(smem_u32(sB));
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int it = 0; it < nit; ++it) {
        int ph, mb, nt;
        if (!mlp_item(S, pair, it, ph, mb, nt)) continue;
        const int num_kb = (ph == 1) ? S.kb1 : S.kb2;
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
            umma_commit_pair_u32(empty0 + stage * 8, 0x3);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        if (elect_one()) umma_commit_pair_u32(tfull0 + acc * 8, 0x3);
        __syncwarp();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp >= TC_EPI_WARP0) {
    // ---- epilogue warps ----
    const int quad = warp & 3;
    const int half = (warp - TC_EPI_WARP0) >> 2;
    uint8_t* stg = stg_all + (warp - TC_EPI_WARP0) * EPI_STG_BYTES;
    const uint32_t tempty0_leader = mapa_u32(smem_u32(tempty), 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    // Publishing a stored fc1 slab = one release-add by lane 0 (MEMBAR.GPU: waits until the warp's stores have reached
    // L2).  Done right after the stores it would stall the warp for the store drain on every tile; it is therefore
    // deferred to the moment the NEXT fc1 tile's accumulator is ready (the stores drained long ago), and done at once
    // only when the next item is an fc2 tile, which may itself depend on the slab.
    int pending_mb = -1;
    const CUtensorMap* tm_hid = p.e1.tma_store ? &tmH : nullptr;   // hidden activations leave through bulk tensor stores
    auto publish = [&]() {
      if (pending_mb >= 0 && !p.nosync) {
        if (lane == 0) {
          if (tm_hid != nullptr) {
            tma_store_wait_all();          // the slab's bulk stores (async proxy, issued by this lane) are performed ...
            fence_proxy_async_global();    // ... and ordered before the generic-proxy release below
          }
          asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.sync + pending_mb) : "memory");
        }
      }
      pending_mb = -1;
    };
    for (int it = 0; it < nit; ++it) {
      int ph, mb, nt;
      if (!mlp_item(S, pair, it, ph, mb, nt)) continue;
      const int m0 = mb * T2_BM + static_cast<int>(rank) * 128;
      const int col = nt * MLP_BN + half * (MLP_BN / 2);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                             static_cast<uint32_t>(acc * Cfg::ACC_STRIDE + half * (MLP_BN / 2));
      if (ph == 1) {
        epilogue_staged<LDT_EPI_BIAS_GELU_BF16, MLP_BN / 2>(p.e1, stg, lane, m0 + quad * 32, col, taddr, [&]() {
          mbar_wait(&tfull[acc], acc_phase);
          tc_fence_after();
          publish();   // the previous fc1 tile's slab
        }, tm_hid);
      } else {
        publish();     // before blocking on an accumulator that may need this very slab
        epilogue_staged<LDT_EPI_GATE_RESID_F32, MLP_BN / 2>(p.e2, stg, lane, m0 + quad * 32, col, taddr, [&]() {
          mbar_wait(&tfull[acc], acc_phase);
          tc_fence_after();
        });
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty0_leader + acc * 8);
      if (ph == 1) {
        fence_proxy_async_global();   // this lane's stores (generic proxy) before later TMA reads (async proxy)
        __syncwarp();                 // ... and before lane 0's release
        pending_mb = mb;
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    publish();
    if (tm_hid != nullptr && lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace ldt

using namespace ldt;

extern "C" int ldt_mlp_sync_words(int M) { return 2 * ((M + T2_BM - 1) / T2_BM); }

extern "C" int ldt_mlp_bf16(const ldt_mlp_args* args, void* stream) {
  LDT_REQUIRE(args != nullptr, LDT_ERR_INVALID, "ldt_mlp_bf16: null args");
  const ldt_mlp_args& a = *args;
  LDT_REQUIRE(a.M > 0 && a.C > 0 && a.inner > 0, LDT_ERR_INVALID, "ldt_mlp_bf16: bad shape M=%d C=%d inner=%d", a.M, a.C, a.inner);
  LDT_REQUIRE(a.C % MLP_BN == 0 && a.inner % MLP_BN == 0, LDT_ERR_UNSUPPORTED,
              "ldt_mlp_bf16: C=%d and inner=%d must be multiples of %d (use two ldt_gemm_bf16 calls otherwise)", a.C, a.inner, MLP_BN);
  LDT_REQUIRE(a.lda >= a.C && a.ldw1 >= a.C && a.ldh >= a.inner && a.ldw2 >= a.inner && a.ldo >= a.C, LDT_ERR_INVALID,
              "ldt_mlp_bf16: leading dimensions too small");
  LDT_REQUIRE(a.lda % 8 == 0 && a.ldw1 % 8 == 0 && a.ldh % 8 == 0 && a.ldw2 % 8 == 0 && a.ldo % 8 == 0, LDT_ERR_INVALID,
              "ldt_mlp_bf16: leading dimensions must be multiples of 8");
  LDT_REQUIRE(a.A && a.W1 && a.hidden && a.W2 && a.resid && a.out && a.sync, LDT_ERR_INVALID, "ldt_mlp_bf16: null operand");
  LDT_REQUIRE((reinterpret_cast<uintptr_t>(a.A) | reinterpret_cast<uintptr_t>(a.W1) | reinterpret_cast<uintptr_t>(a.hidden) |
               reinterpret_cast<uintptr_t>(a.W2) | reinterpret_cast<uintptr_t>(a.bias1) | reinterpret_cast<uintptr_t>(a.bias2) |
               reinterpret_cast<uintptr_t>(a.resid) | reinterpret_cast<uintptr_t>(a.out) | reinterpret_cast<uintptr_t>(a.gate)) % 16 == 0,
              LDT_ERR_INVALID, "ldt_mlp_bf16: operands must be 16-byte aligned");
  LDT_REQUIRE(a.gate == nullptr || a.gate_stride % 4 == 0, LDT_ERR_INVALID, "ldt_mlp_bf16: gate_stride must be a multiple of 4");
  cudaStream_t s = static_cast<cudaStream_t>(stream);

  using Cfg = Tc2Cfg<MLP_BN>;
  static PerDevice<int> max_pairs_dev;
  int& max_pairs = max_pairs_dev.get_or(-1);
  if (max_pairs < 0) {
    LDT_CUDA_OK(cudaFuncSetAttribute(mlp_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    // every CTA of the grid must be resident at once (tiles wait for tiles of other CTAs): ask the driver how many
    // clusters of two fit, like a cooperative launch would
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(num_sms());
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    LDT_CUDA_OK(cudaOccupancyMaxActiveClusters(&n, mlp_tc2_kernel, &cfg));
    max_pairs = n;
  }
  LDT_REQUIRE(max_pairs > 0, LDT_ERR_UNSUPPORTED, "ldt_mlp_bf16: no CTA pair of this kernel fits on the device");

  const int tiles_m = (a.M + T2_BM - 1) / T2_BM;
  const int tn1 = a.inner / MLP_BN, tn2 = a.C / MLP_BN;
  const int P = min(min(num_sms() / 2, max_pairs), tiles_m * tn1);
  const MlpSched S = mlp_make_sched(tiles_m, tn1, tn2, a.C / TC_BK, a.inner / TC_BK, P);

  CUtensorMap tmA1, tmW1, tmA2, tmW2, tmH;
  int rc = make_tmap_bf16(&tmA1, a.A, a.M, a.C, a.lda, 128);
  if (rc) return rc;
  rc = make_tmap_bf16(&tmW1, a.W1, a.inner, a.C, a.ldw1, MLP_BN / 2);
  if (rc) return rc;
  rc = make_tmap_bf16(&tmA2, a.hidden, a.M, a.inner, a.ldh, 128);
  if (rc) return rc;
  rc = make_tmap_bf16(&tmW2, a.W2, a.C, a.inner, a.ldw2, MLP_BN / 2);
  if (rc) return rc;
  rc = make_tmap_bf16(&tmH, a.hidden, a.M, a.inner, a.ldh, 32);   // store boxes of the fc1 epilogue
  if (rc) return rc;

  MlpParams p;
  p.e1.M = a.M; p.e1.N = a.inner; p.e1.bias = a.bias1; p.e1.out = a.hidden; p.e1.ldo = a.ldh;
  p.e1.resid = nullptr; p.e1.gate = nullptr; p.e1.gate_stride = 0; p.e1.rows_per_gate = 1; p.e1.dbg = nullptr; p.e1.dbg_mode = 0; p.e1.relu = 0; p.e1.f32_plain = 0; p.e1.tma_store = (ldt_debug_get_gemm_mode() & 256) ? 0 : 1;
  p.e2.M = a.M; p.e2.N = a.C; p.e2.bias = a.bias2; p.e2.out = a.out; p.e2.ldo = a.ldo;
  p.e2.resid = a.resid; p.e2.gate = a.gate; p.e2.gate_stride = a.gate_stride;
  p.e2.rows_per_gate = a.rows_per_gate > 0 ? a.rows_per_gate : 1; p.e2.dbg = nullptr; p.e2.dbg_mode = 0; p.e2.relu = 0; p.e2.f32_plain = 0; p.e2.tma_store = 0;
  p.sync = a.sync;
  p.ready_target = static_cast<unsigned int>(tn1 * 16);
  p.done_target = static_cast<unsigned int>(tn2 * 2);
  p.tiles_m = tiles_m;
  p.nosync = (ldt_debug_get_gemm_mode() & 8) ? 1 : 0;
  if (p.nosync) p.ready_target = 0;
  LDT_CUDA_OK(launch_pdl(mlp_tc2_kernel, dim3(2 * P), dim3(TC_THREADS), Cfg::SMEM_BYTES, s, tmA1, tmW1, tmA2, tmW2, tmH, p, S));
  return LDT_OK;
}

// Host-side view of the static schedule (tests): item `it` of pair `p` -> phase (0 = surplus slot), m-block, column tile.
extern "C" int ldt_mlp_schedule_item(int tiles_m, int tn1, int tn2, int kb1, int kb2, int pairs, int p, int it, int* num_items,
                                     int* phase, int* mb, int* nt) {
  LDT_REQUIRE(tiles_m > 0 && tn1 > 0 && tn2 > 0 && kb1 > 0 && kb2 > 0 && pairs > 0 && p >= 0 && p < pairs, LDT_ERR_INVALID,
              "ldt_mlp_schedule_item: bad arguments");
  const MlpSched S = mlp_make_sched(tiles_m, tn1, tn2, kb1, kb2, pairs);
  *num_items = mlp_num_items(S, p);
  *phase = 0; *mb = 0; *nt = 0;
  if (it >= 0 && it < *num_items) {
    int ph, m, n;
    if (mlp_item(S, p, it, ph, m, n)) { *phase = ph; *mb = m; *nt = n; }
  }
  return LDT_OK;
}
