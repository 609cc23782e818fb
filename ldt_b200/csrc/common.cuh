// Shared helpers for the ldt_b200 sm_100a kernels: error plumbing for the C-ABI and thin
// inline-PTX wrappers (mbarrier, TMA, tcgen05/TMEM).  Everything here is device-side glue;
// no reference code corresponds to it.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include <utility>

namespace ldt {

// ---- error plumbing (C-ABI never throws; see include/ldt_b200.h) -------------------------
void set_last_error(const char* fmt, ...);
int  check_cuda(cudaError_t e, const char* what, const char* file, int line);

#define LDT_CUDA_OK(expr)                                                     \
  do {                                                                        \
    int _rc = ::ldt::check_cuda((expr), #expr, __FILE__, __LINE__);           \
    if (_rc != 0) return _rc;                                                 \
  } while (0)

#define LDT_REQUIRE(cond, code, ...)                                          \
  do {                                                                        \
    if (!(cond)) {                                                            \
      ::ldt::set_last_error(__VA_ARGS__);                                     \
      return (code);                                                          \
    }                                                                         \
  } while (0)

enum : int {
  LDT_OK = 0,
  LDT_ERR_INVALID = -1,   // bad argument (shape, null pointer, alignment)
  LDT_ERR_CUDA = -2,      // a CUDA runtime call failed
  LDT_ERR_UNSUPPORTED = -3,
  LDT_ERR_WORKSPACE = -4, // caller-provided workspace too small
};

int  num_sms();           // cudaDevAttrMultiProcessorCount of the CURRENT device (cached per device)
int  current_device();    // cudaGetDevice, clamped to [0, LDT_MAX_DEVICES)

// One-time per-DEVICE set-up (cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies to one device only, occupancy
// and SM-count queries likewise): `static PerDevice<bool> done; if (!done.get()) { ...; done.get() = true; }`.
constexpr int LDT_MAX_DEVICES = 64;
template <typename T>
struct PerDevice {
  T v[LDT_MAX_DEVICES];
  bool set[LDT_MAX_DEVICES];
  T& get() { return v[current_device()]; }
  // value with a sentinel default on first touch of a device
  T& get_or(T init) {
    const int d = current_device();
    if (!set[d]) { v[d] = init; set[d] = true; }
    return v[d];
  }
};
bool pdl_enabled();       // programmatic dependent launch: off by default, on via ldt_set_pdl(1) or env LDT_PDL=1

// ---- programmatic dependent launch (PDL) ----------------------------------------------------------
// Every kernel of the per-step chain starts with pdl_launch_dependents() (the next kernel's CTAs may be scheduled as
// soon as this grid's CTAs have all started and SM resources free up) and executes pdl_wait() before its first access
// to global memory (blocks until the preceding grid has completed and flushed).  Net effect: launch latency and the
// next kernel's prologue (barrier init, TMEM allocation, descriptor prefetch) overlap this kernel's tail.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

// ---- small device helpers -------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// The same on raw shared-memory addresses (uniform-register friendly: the converged issue loops of the CTA-pair kernels
// keep barrier addresses as base + 8 * stage).
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%0], %1;\n\t"
      "@!P bra WAIT_%=;\n\t"
      "}\n"
      ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_u32(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

// ---- TMA (cp.async.bulk.tensor) ------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
// 2-D tiled load: coordinates are (c0 = innermost/contiguous, c1 = row)
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// 2-D tiled TMA STORE of one box from this CTA's shared memory (bulk async-group completion).  Rows / columns outside
// the tensor are clipped by the hardware.
__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all of this thread's bulk groups have finished READING shared memory (the source may be overwritten)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... have completed altogether (the global writes are performed)
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy writes to shared memory made visible to the async proxy (TMA store source)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- tcgen05 / TMEM ---------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate.  One thread issues.
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 columns of 32-bit: thread i of the warp gets lane (base+i), v[j] = column (base+j).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 32 lanes x 16 columns of 32-bit, registers -> TMEM: thread i of the warp writes lane (base+i), column (base+j) = v[j].
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand is read from tensor memory (lane = row, two bf16 K elements per
// 32-bit column), bf16 inputs, fp32 accumulate.  One thread issues.
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---- thread-block clusters / CTA pairs (cta_group::2) ---------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address of this CTA -> shared::cluster address of the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// 2-D tiled TMA load issued by either CTA of a pair; the mbarrier (a shared::cluster address) may live in the peer.
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const void* tmap, uint32_t bar_cluster_addr, int c0,
                                                 int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair_u32(uint32_t smem_dst, const void* tmap, uint32_t bar_cluster_addr, int c0,
                                                     int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
// The same, multicast: the box lands at the same shared-memory offset in every CTA of `cta_mask`, and each destination
// CTA's bytes are signalled on the barrier at `bar`'s offset in the LEADER (even) CTA of that destination's pair (the
// address must have the pair's peer bit clear, e.g. mapa to an even rank).
__device__ __forceinline__ void tma_load_2d_pair_mcast_u32(uint32_t smem_dst, const void* tmap, uint32_t bar_cluster_addr,
                                                           int c0, int c1, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {  // one warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A * B over a CTA pair: M = 256 (128 rows per CTA), B split along N across the pair.
// Issued by one thread of the leader (even) CTA only.
__device__ __forceinline__ void umma_bf16_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                  uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// The same with fp32 operands read as TF32 (kind::tf32: 8 K-elements = 32 bytes per instruction; the low 13 mantissa bits
// of each operand are ignored by the tensor core, so producers round to TF32 first -- round_tf32 below).
__device__ __forceinline__ void umma_tf32_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                  uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// accumulate form without the predicate set-up (every k-slice after the first of a tile)
__device__ __forceinline__ void umma_bf16_ss_pair_acc(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.eq.b32 p, 0, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair_u32(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask)
               : "memory");
}
// Arrive (once all prior MMAs of this thread completed) on the barrier at this smem offset in every CTA of `mask`.
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}

// UMMA shared-memory matrix descriptor for a K-major bf16 tile stored as rows of 128 bytes with the
// 128-byte swizzle TMA writes (8-row groups 1024 B apart).  Field layout: cute/arch/mma_sm100_desc.hpp.
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);   // start address [0,14)
  d |= static_cast<uint64_t>(1) << 16;                      // leading byte offset (unused for SW128 K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;              // stride byte offset: 8 rows * 128 B
  d |= static_cast<uint64_t>(1) << 46;                      // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;                      // layout type SWIZZLE_128B
  return d;
}
// Instruction descriptor: bf16 x bf16 -> fp32, both operands K-major, tile M x N.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n) {
  return (1u << 4)                       // D format  = F32
         | (1u << 7)                     // A format  = BF16
         | (1u << 10)                    // B format  = BF16
         | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

// Instruction descriptor: tf32 x tf32 -> fp32 (fp32 operands in shared memory), both operands K-major, tile M x N.
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int m, int n) {
  return (1u << 4)                       // D format  = F32
         | (2u << 7)                     // A format  = TF32
         | (2u << 10)                    // B format  = TF32
         | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

// ---- math -----------------------------------------------------------------------------------------
// fp32 -> nearest TF32 value (10 explicit mantissa bits), kept in an fp32 container: what cuBLAS / cuDNN feed their TF32
// tensor-core paths (cvt.rna).  Operands of the kind::tf32 contractions are rounded once where they are produced.
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }
// exact-erf GELU (nn.GELU() default, reference tools/utils.py:107-108)
__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
// The same function for GEMM epilogues, branch-free and with ONE special-function op per element.  The epilogue of the
// fc1 GEMM is bound by the MUFU pipe (16 lanes/clk/SM: a warp-wide MUFU occupies its scheduler's share for 8 cycles), so
// the form matters more than the FP32 count:
//     gelu(x) = max(x, 0) - |x| * erfc(|x| / sqrt(2)) / 2,      erfc(|x| / sqrt(2)) / 2 = 2^(q(|x|) - 1),
// q = degree-7 fit of log2(erfc(a / sqrt(2))) on a in [0, 5.6] (weighted for absolute error of erfc); beyond 5.6 erfc
// < 2.2e-8 and the argument is clamped.  fp32 Horner + ex2.approx: |abs error| <= 3.7e-7, relative error <= 3.1e-5
// wherever |gelu| > 1e-3 (the Abramowitz-Stegun 7.1.26 form it replaces: 2.1e-7 / 1.7e-4, with rcp + ex2) -- two orders
// below the bf16 rounding applied to the result.  1 MUFU + 10 FP32-pipe ops.
__device__ __forceinline__ float gelu_erf_fast(float x) {
#ifdef LDT_GELU_AS   // A/B builds only: the Abramowitz-Stegun 7.1.26 form (rcp + ex2) this replaced
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f * 0.70710678118654752440f, fabsf(x), 1.0f)));
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"((x * x) * (-0.5f * 1.4426950408889634f)));
  const float hx0 = 0.5f * x;
  return fmaf(fabsf(hx0), fmaf(-poly, e, 1.0f), hx0);
#else
  const float a = fminf(fabsf(x), 5.6f);
  float q = fmaf(7.732903964e-06f, a, -4.679716980e-05f);
  q = fmaf(q, a, -4.473918779e-04f);
  q = fmaf(q, a, 7.435896264e-03f);
  q = fmaf(q, a, -5.273329248e-02f);
  q = fmaf(q, a, -4.591345845e-01f);
  q = fmaf(q, a, -1.151114419e+00f);
  q = fmaf(q, a, 3.068670393e-07f - 1.0f);   // -1: the factor 1/2 of |x/2| rides in the exponent
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(q));
  return fmaf(-fabsf(x), e, fmaxf(x, 0.0f));
#endif
}
// Two elements at a time with Blackwell's packed fp32 arithmetic (FFMA2: one issue slot for two FMAs): the degree-7 Horner
// chain, which is 7 of the 11 FP32 ops of gelu_erf_fast, issues half as many instructions.  Same arithmetic per element
// (fma.rn.f32x2 rounds each half like fma.rn.f32), so the results are bit-identical to gelu_erf_fast.
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
  return d;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ void gelu_erf_fast_x2(float& x0, float& x1) {
  const uint64_t a = pack_f32x2(fminf(fabsf(x0), 5.6f), fminf(fabsf(x1), 5.6f));
  uint64_t q = fma_f32x2(pack_f32x2(7.732903964e-06f, 7.732903964e-06f), a, pack_f32x2(-4.679716980e-05f, -4.679716980e-05f));
  q = fma_f32x2(q, a, pack_f32x2(-4.473918779e-04f, -4.473918779e-04f));
  q = fma_f32x2(q, a, pack_f32x2(7.435896264e-03f, 7.435896264e-03f));
  q = fma_f32x2(q, a, pack_f32x2(-5.273329248e-02f, -5.273329248e-02f));
  q = fma_f32x2(q, a, pack_f32x2(-4.591345845e-01f, -4.591345845e-01f));
  q = fma_f32x2(q, a, pack_f32x2(-1.151114419e+00f, -1.151114419e+00f));
  q = fma_f32x2(q, a, pack_f32x2(3.068670393e-07f - 1.0f, 3.068670393e-07f - 1.0f));
  float q0, q1, e0, e1;
  unpack_f32x2(q, q0, q1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(q0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(q1));
  x0 = fmaf(-fabsf(x0), e0, fmaxf(x0, 0.0f));
  x1 = fmaf(-fabsf(x1), e1, fmaxf(x1, 0.0f));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

}  // namespace ldt
