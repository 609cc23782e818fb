// Multi-head attention over the 32 latent tokens: softmax(q k^T / sqrt(dh)) v, one warp per
// (batch, head, 32-query tile).  Used by the score net (32 queries x 32 keys, dh 64, 16 heads) and by
// the Compressor decoder (2048 query points x 32 keys, dh 32, 4 heads).
//
// Replaces ResidualBlock.compute_attention (model/layers.py:183-200), which the reference runs as
// permute/clone + bmm + mul + softmax + bmm (9 launches).  The whole score tile (32x32) lives in mma
// accumulator registers, so no online softmax is needed.  Output layout reproduces the reference's
// `(w @ v).reshape(B, N, C)` (layers.py:197): result [B,H,Nq,dh] stored contiguously, which the next
// layer re-reads as token-major [B*Nq, H*dh] without permuting heads back.
//
// Tensor-core path: mma.sync.m16n8k16 (bf16 -> fp32).  At 0.5 % of the block's FLOPs and 32-wide tiles
// this op is bound by moving q/k/v/o (L2-resident), not by tensor throughput; tcgen05's 128-row tiles
// would idle 3/4 of the array on block-diagonal head structure, so the warp-level MMA is the fit here.
#include "common.cuh"
#include "ldt_b200.h"
#include "mma_sync.cuh"

namespace ldt {

constexpr int ATT_WARPS = 4;

template <int DH>
__global__ void __launch_bounds__(ATT_WARPS * 32)
attention_nk32_kernel(int units, int H, int Nq, int qtiles, const __nv_bfloat16* __restrict__ q, int ldq,
                      const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v, int ldkv,
                      __nv_bfloat16* __restrict__ o, float scale_log2e) {
  constexpr int VS = DH + 8;  // padded row (elements): 16-byte aligned rows, conflict-free ldmatrix
  __shared__ __align__(16) __nv_bfloat16 Vs[ATT_WARPS][32 * VS];
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int unit = blockIdx.x * ATT_WARPS + warp;
  if (unit >= units) return;
  const int qt = unit % qtiles;
  const int bh = unit / qtiles;
  const int h = bh % H, b = bh / H;
  const int g = lane >> 2, t = lane & 3;

  // ---- stage V[32, DH] of this (b, h) ----
  __nv_bfloat16* vs = Vs[warp];
  {
    constexpr int CH = DH / 8;  // 16-byte chunks per row
    const __nv_bfloat16* vb = v + static_cast<size_t>(b) * 32 * ldkv + h * DH;
    for (int i = lane; i < 32 * CH; i += 32) {
      const int key = i / CH, c = i % CH;
      *reinterpret_cast<uint4*>(vs + key * VS + c * 8) =
          *reinterpret_cast<const uint4*>(vb + static_cast<size_t>(key) * ldkv + c * 8);
    }
  }
  __syncwarp();

  // ---- S = Q K^T ----
  const int n_base = qt * 32;
  float s[2][4][4];
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[mi][ni][e] = 0.f;
  {
    const __nv_bfloat16* qb = q + static_cast<size_t>(b) * Nq * ldq + h * DH;
    const __nv_bfloat16* qrow[2][2];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
      qrow[mi][0] = qb + static_cast<size_t>(min(n_base + mi * 16 + g, Nq - 1)) * ldq;
      qrow[mi][1] = qb + static_cast<size_t>(min(n_base + mi * 16 + g + 8, Nq - 1)) * ldq;
    }
    const __nv_bfloat16* kb = k + static_cast<size_t>(b) * 32 * ldkv + h * DH;
#pragma unroll
    for (int ks = 0; ks < DH / 16; ++ks) {
      uint32_t a[2][4];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        a[mi][0] = ld_u32(qrow[mi][0] + ks * 16 + 2 * t);
        a[mi][1] = ld_u32(qrow[mi][1] + ks * 16 + 2 * t);
        a[mi][2] = ld_u32(qrow[mi][0] + ks * 16 + 2 * t + 8);
        a[mi][3] = ld_u32(qrow[mi][1] + ks * 16 + 2 * t + 8);
      }
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        const __nv_bfloat16* krow = kb + static_cast<size_t>(ni * 8 + g) * ldkv + ks * 16 + 2 * t;
        const uint32_t b0 = ld_u32(krow), b1 = ld_u32(krow + 8);
        mma_bf16_16816(s[0][ni], a[0], b0, b1);
        mma_bf16_16816(s[1][ni], a[1], b0, b1);
      }
    }
  }

  // ---- softmax over the 32 keys (rows g and g+8 of each 16-row tile) ----
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
        const float p0 = exp2f((s[mi][ni][2 * hh] - m) * scale_log2e);
        const float p1 = exp2f((s[mi][ni][2 * hh + 1] - m) * scale_log2e);
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

  // ---- O = P V ----
  __nv_bfloat16* ob = o + (static_cast<size_t>(bh) * Nq) * DH;
  const int mat = lane >> 3, r8 = lane & 7;
#pragma unroll
  for (int np = 0; np < DH / 16; ++np) {  // pairs of 8-wide output tiles
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
      const __nv_bfloat16* addr = vs + (16 * j + (mat & 1) * 8 + r8) * VS + (2 * np + (mat >> 1)) * 8;
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
        const int n0 = n_base + mi * 16 + g;
        if (n0 < Nq)
          *reinterpret_cast<uint32_t*>(ob + static_cast<size_t>(n0) * DH + col) =
              pack_bf16(acc[mi][nn][0] * inv_sum[mi][0], acc[mi][nn][1] * inv_sum[mi][0]);
        if (n0 + 8 < Nq)
          *reinterpret_cast<uint32_t*>(ob + static_cast<size_t>(n0 + 8) * DH + col) =
              pack_bf16(acc[mi][nn][2] * inv_sum[mi][1], acc[mi][nn][3] * inv_sum[mi][1]);
      }
  }
}

// ------------------------------------------------------------------------------------------------
// Attention of a SHORT query set over a LONG key set (32 latent tokens attending to the 2048 decoded points:
// DecoderBlock.compute_posterior -> ResidualBlock.compute_attention with y = o, model/Compressor/Network.py:62-77,
// model/layers.py:183-200).  One warp per (batch, head, query); keys are streamed in chunks of 32 through shared
// memory with an online softmax in fp32: lane j scores key j of the chunk, lane d owns output channel d.
// Same output layout quirk as above: [B,H,Nq,dh] stored contiguously.  0.07 % of the encoder's FLOPs: SIMT.
// ------------------------------------------------------------------------------------------------
constexpr int ATTL_WARPS = 8;   // queries per CTA

__device__ __forceinline__ float attl_load(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ float attl_load(const float* p) { return *p; }
__device__ __forceinline__ void attl_store(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ void attl_store(float* p, float v) { *p = v; }

// T = __nv_bfloat16 (product path) or float (the fp32 parity mode of Compressor.forward: same kernel, plain fp32 in and out)
template <int DH, typename T>
__global__ void __launch_bounds__(ATTL_WARPS * 32)
attention_longkv_kernel(int H, int Nq, int Nk, const T* __restrict__ q, int ldq,
                        const T* __restrict__ k, const T* __restrict__ v, int ldkv,
                        T* __restrict__ o, float scale_log2e) {
  static_assert(DH == 32 || DH == 64, "lane d owns output channels d, d + 32, ...");
  constexpr int CPL = DH / 32;   // output channels per lane
  __shared__ float Ks[32][DH + 1], Vs[32][DH + 1], Qs[ATTL_WARPS][DH];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qblocks = (Nq + ATTL_WARPS - 1) / ATTL_WARPS;
  const int qb = blockIdx.x % qblocks;
  const int bh = blockIdx.x / qblocks;
  const int h = bh % H, b = bh / H;
  const int qi = qb * ATTL_WARPS + warp;
  const bool live = qi < Nq;
  if (live) {
#pragma unroll
    for (int c = 0; c < CPL; ++c)
      Qs[warp][lane + 32 * c] = attl_load(q + (static_cast<size_t>(b) * Nq + qi) * ldq + h * DH + lane + 32 * c);
  }
  float m = -INFINITY, l = 0.f, acc[CPL];
#pragma unroll
  for (int c = 0; c < CPL; ++c) acc[c] = 0.f;
  for (int k0 = 0; k0 < Nk; k0 += 32) {
    __syncthreads();   // previous chunk fully consumed
    for (int i = threadIdx.x; i < 32 * DH; i += ATTL_WARPS * 32) {
      const int key = i / DH, d = i % DH;
      const bool ok = k0 + key < Nk;
      const size_t off = (static_cast<size_t>(b) * Nk + k0 + key) * ldkv + h * DH + d;
      Ks[key][d] = ok ? attl_load(k + off) : 0.f;
      Vs[key][d] = ok ? attl_load(v + off) : 0.f;
    }
    __syncthreads();
    if (live) {
      float sc = 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) sc = fmaf(Qs[warp][d], Ks[lane][d], sc);
      sc = (k0 + lane < Nk) ? sc * scale_log2e : -INFINITY;
      float cmax = sc;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) cmax = fmaxf(cmax, __shfl_xor_sync(0xffffffffu, cmax, off));
      const float m_new = fmaxf(m, cmax);
      const float corr = exp2f(m - m_new);          // first chunk: exp2(-inf) = 0
      const float pj = exp2f(sc - m_new);
      l = l * corr + warp_sum(pj);
#pragma unroll
      for (int c = 0; c < CPL; ++c) acc[c] *= corr;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float pb = __shfl_sync(0xffffffffu, pj, j);
#pragma unroll
        for (int c = 0; c < CPL; ++c) acc[c] = fmaf(pb, Vs[j][lane + 32 * c], acc[c]);
      }
      m = m_new;
    }
  }
  if (live) {
#pragma unroll
    for (int c = 0; c < CPL; ++c)
      attl_store(o + ((static_cast<size_t>(b) * H + h) * Nq + qi) * DH + lane + 32 * c, acc[c] / l);
  }
}

// ------------------------------------------------------------------------------------------------
// Narrow heads (head dim 8 or 16: the Hybrid trainer's score net is 128 wide with 16 heads,
// experiments/Hybrid_Trainer/airplane/config.yaml:53-54) over the 32 latent tokens: the mma.sync tiles above need a
// contraction of at least 16 per instruction and would run 3/4 empty, so one lane per query row does the whole row in
// registers (32 scores, DH outputs); K and V of the (batch, head) live in shared memory.  Same layouts as above.
// ------------------------------------------------------------------------------------------------
template <int DH>
__global__ void __launch_bounds__(ATT_WARPS * 32)
attention_nk32_narrow_kernel(int units, int H, int Nq, int qtiles, const __nv_bfloat16* __restrict__ q, int ldq,
                             const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v, int ldkv,
                             __nv_bfloat16* __restrict__ o, float scale_log2e) {
  __shared__ float Ks[ATT_WARPS][32][DH + 1], Vs[ATT_WARPS][32][DH + 1];
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int unit = blockIdx.x * ATT_WARPS + warp;
  if (unit >= units) return;
  const int qt = unit % qtiles;
  const int bh = unit / qtiles;
  const int h = bh % H, b = bh / H;
#pragma unroll
  for (int d = 0; d < DH; ++d) {   // lane = key
    Ks[warp][lane][d] = __bfloat162float(k[(static_cast<size_t>(b) * 32 + lane) * ldkv + h * DH + d]);
    Vs[warp][lane][d] = __bfloat162float(v[(static_cast<size_t>(b) * 32 + lane) * ldkv + h * DH + d]);
  }
  __syncwarp();
  const int qi = qt * 32 + lane;
  if (qi >= Nq) return;
  float qr[DH];
#pragma unroll
  for (int d = 0; d < DH; ++d) qr[d] = __bfloat162float(q[(static_cast<size_t>(b) * Nq + qi) * ldq + h * DH + d]);
  float sc[32], mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    float a = 0.f;
#pragma unroll
    for (int d = 0; d < DH; ++d) a = fmaf(qr[d], Ks[warp][j][d], a);
    sc[j] = a * scale_log2e;
    mx = fmaxf(mx, sc[j]);
  }
  float sum = 0.f, out[DH];
#pragma unroll
  for (int d = 0; d < DH; ++d) out[d] = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float e = exp2f(sc[j] - mx);
    sum += e;
    // the wide-head kernel rounds the un-normalised probabilities to bf16 before the PV product; keep the same
    // rounding point for both head widths (the parity tests emulate it)
    const float p = __bfloat162float(__float2bfloat16_rn(e));
#pragma unroll
    for (int d = 0; d < DH; ++d) out[d] = fmaf(p, Vs[warp][j][d], out[d]);
  }
  const float inv = 1.0f / sum;
  __nv_bfloat16* op = o + ((static_cast<size_t>(b) * H + h) * Nq + qi) * DH;
#pragma unroll
  for (int d = 0; d < DH; ++d) op[d] = __float2bfloat16_rn(out[d] * inv);
}

}  // namespace ldt

namespace ldt {
int attention_nk32_tc(int B, int H, int Nq, int dh, const void* q, int ldq, const void* k, const void* v, int ldkv, void* o,
                      cudaStream_t s);   // attention_tc.cu
}

using namespace ldt;

// 0 = tcgen05 kernel wherever it applies (Nq == 32 or Nq >= 128, dh in {32, 64}), 1 = the warp-level mma.sync kernels for
// every shape (kept as the cross-check of the tcgen05 path in tests and for the shapes it does not take)
static int g_attention_backend = 0;
extern "C" int ldt_debug_set_attention_backend(int backend) {
  g_attention_backend = backend;
  return LDT_OK;
}
extern "C" int ldt_debug_get_attention_backend(void) { return g_attention_backend; }

extern "C" int ldt_attention_nk32(int B, int H, int Nq, int dh, const void* q, int ldq, const void* k, const void* v,
                                  int ldkv, void* o, void* stream) {
  LDT_REQUIRE(B >= 0 && H > 0 && Nq > 0, LDT_ERR_INVALID, "ldt_attention_nk32: bad shape B=%d H=%d Nq=%d", B, H, Nq);
  LDT_REQUIRE(dh == 8 || dh == 16 || dh == 32 || dh == 64, LDT_ERR_UNSUPPORTED, "ldt_attention_nk32: head dim %d not in {8,16,32,64}", dh);
  if (B == 0) return LDT_OK;
  LDT_REQUIRE(q && k && v && o, LDT_ERR_INVALID, "ldt_attention_nk32: null pointer");
  LDT_REQUIRE(ldq % 8 == 0 && ldkv % 8 == 0 && ldq >= H * dh && ldkv >= H * dh, LDT_ERR_INVALID,
              "ldt_attention_nk32: ldq=%d ldkv=%d must be multiples of 8 and >= H*dh", ldq, ldkv);
  LDT_REQUIRE((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
               reinterpret_cast<uintptr_t>(o)) % 16 == 0,
              LDT_ERR_INVALID, "ldt_attention_nk32: pointers must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (g_attention_backend == 0 && (dh == 32 || dh == 64) && (Nq == 32 || Nq >= 128)) {
    const int rc = attention_nk32_tc(B, H, Nq, dh, q, ldq, k, v, ldkv, o, s);
    if (rc != LDT_ERR_UNSUPPORTED) return rc;
  }
  const int qtiles = (Nq + 31) / 32;
  const long long units = static_cast<long long>(B) * H * qtiles;
  LDT_REQUIRE(units < (1LL << 31), LDT_ERR_INVALID, "ldt_attention_nk32: too many work units");
  const int grid = static_cast<int>((units + ATT_WARPS - 1) / ATT_WARPS);
  const float scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(dh));
  if (dh == 8 || dh == 16) {
    auto* qq = static_cast<const __nv_bfloat16*>(q);
    auto* kk = static_cast<const __nv_bfloat16*>(k);
    auto* vv = static_cast<const __nv_bfloat16*>(v);
    auto* oo = static_cast<__nv_bfloat16*>(o);
    if (dh == 8)
      attention_nk32_narrow_kernel<8><<<grid, ATT_WARPS * 32, 0, s>>>(static_cast<int>(units), H, Nq, qtiles, qq, ldq, kk, vv, ldkv, oo, scale_log2e);
    else
      attention_nk32_narrow_kernel<16><<<grid, ATT_WARPS * 32, 0, s>>>(static_cast<int>(units), H, Nq, qtiles, qq, ldq, kk, vv, ldkv, oo, scale_log2e);
    LDT_CUDA_OK(cudaGetLastError());
    return LDT_OK;
  }
  if (dh == 64)
    attention_nk32_kernel<64><<<grid, ATT_WARPS * 32, 0, s>>>(static_cast<int>(units), H, Nq, qtiles,
                                                             static_cast<const __nv_bfloat16*>(q), ldq,
                                                             static_cast<const __nv_bfloat16*>(k),
                                                             static_cast<const __nv_bfloat16*>(v), ldkv,
                                                             static_cast<__nv_bfloat16*>(o), scale_log2e);
  else
    attention_nk32_kernel<32><<<grid, ATT_WARPS * 32, 0, s>>>(static_cast<int>(units), H, Nq, qtiles,
                                                             static_cast<const __nv_bfloat16*>(q), ldq,
                                                             static_cast<const __nv_bfloat16*>(k),
                                                             static_cast<const __nv_bfloat16*>(v), ldkv,
                                                             static_cast<__nv_bfloat16*>(o), scale_log2e);
  LDT_CUDA_OK(cudaGetLastError());
  return LDT_OK;
}

extern "C" int ldt_attention_longkv(int B, int H, int Nq, int Nk, int dh, const void* q, int ldq, const void* k, const void* v,
                                    int ldkv, void* o, void* stream) {
  LDT_REQUIRE(B >= 0 && H > 0 && Nq > 0 && Nk > 0, LDT_ERR_INVALID, "ldt_attention_longkv: bad shape B=%d H=%d Nq=%d Nk=%d", B, H,
              Nq, Nk);
  LDT_REQUIRE(dh == 32 || dh == 64, LDT_ERR_UNSUPPORTED, "ldt_attention_longkv: head dim %d not in {32,64}", dh);
  if (B == 0) return LDT_OK;
  LDT_REQUIRE(q && k && v && o, LDT_ERR_INVALID, "ldt_attention_longkv: null pointer");
  LDT_REQUIRE(ldq >= H * dh && ldkv >= H * dh, LDT_ERR_INVALID, "ldt_attention_longkv: ldq=%d ldkv=%d must be >= H*dh", ldq, ldkv);
  const long long blocks = static_cast<long long>(B) * H * ((Nq + ATTL_WARPS - 1) / ATTL_WARPS);
  LDT_REQUIRE(blocks < (1LL << 31), LDT_ERR_INVALID, "ldt_attention_longkv: too many work units");
  const float scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(dh));
  if (dh == 32)
    attention_longkv_kernel<32, __nv_bfloat16><<<static_cast<int>(blocks), ATTL_WARPS * 32, 0, static_cast<cudaStream_t>(stream)>>>(
        H, Nq, Nk, static_cast<const __nv_bfloat16*>(q), ldq, static_cast<const __nv_bfloat16*>(k),
        static_cast<const __nv_bfloat16*>(v), ldkv, static_cast<__nv_bfloat16*>(o), scale_log2e);
  else
    attention_longkv_kernel<64, __nv_bfloat16><<<static_cast<int>(blocks), ATTL_WARPS * 32, 0, static_cast<cudaStream_t>(stream)>>>(
        H, Nq, Nk, static_cast<const __nv_bfloat16*>(q), ldq, static_cast<const __nv_bfloat16*>(k),
        static_cast<const __nv_bfloat16*>(v), ldkv, static_cast<__nv_bfloat16*>(o), scale_log2e);
  LDT_CUDA_OK(cudaGetLastError());
  return LDT_OK;
}

extern "C" int ldt_attention_longkv_f32(int B, int H, int Nq, int Nk, int dh, const float* q, int ldq, const float* k, const float* v,
                                        int ldkv, float* o, void* stream) {
  LDT_REQUIRE(B >= 0 && H > 0 && Nq > 0 && Nk > 0, LDT_ERR_INVALID, "ldt_attention_longkv_f32: bad shape B=%d H=%d Nq=%d Nk=%d", B, H,
              Nq, Nk);
  LDT_REQUIRE(dh == 32 || dh == 64, LDT_ERR_UNSUPPORTED, "ldt_attention_longkv_f32: head dim %d not in {32,64}", dh);
  if (B == 0) return LDT_OK;
  LDT_REQUIRE(q && k && v && o, LDT_ERR_INVALID, "ldt_attention_longkv_f32: null pointer");
  LDT_REQUIRE(ldq >= H * dh && ldkv >= H * dh, LDT_ERR_INVALID, "ldt_attention_longkv_f32: ldq=%d ldkv=%d must be >= H*dh", ldq, ldkv);
  const long long blocks = static_cast<long long>(B) * H * ((Nq + ATTL_WARPS - 1) / ATTL_WARPS);
  LDT_REQUIRE(blocks < (1LL << 31), LDT_ERR_INVALID, "ldt_attention_longkv_f32: too many work units");
  const float scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(dh));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dh == 32) attention_longkv_kernel<32, float><<<static_cast<int>(blocks), ATTL_WARPS * 32, 0, st>>>(H, Nq, Nk, q, ldq, k, v, ldkv, o, scale_log2e);
  else attention_longkv_kernel<64, float><<<static_cast<int>(blocks), ATTL_WARPS * 32, 0, st>>>(H, Nq, Nk, q, ldq, k, v, ldkv, o, scale_log2e);
  LDT_CUDA_OK(cudaGetLastError());
  return LDT_OK;
}
