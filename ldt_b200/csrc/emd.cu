// Approximate earth-mover distance between point sets (the "EMD" half of the generation metrics).
//
// Replaces the forward kernels of evaluation/pytorch_structural_losses/src/approxmatch.cu: approxmatchkernel (:3-182,
// a 9-level soft-assignment auction writing a dense match[b,m,n]) and matchcostkernel (:184-224, sum match * distance),
// which the reference launches as <<<32,512>>> -- one CTA per batch item, 32 CTAs on the whole GPU -- from
// StructuralLosses.match_cost (StructuralLosses/match_cost.py:6-45), called per row of the pairwise matrix by
// _pairwise_EMD_CD_ (evaluation/evaluation_metrics.py:112-162).
//
// Algorithm per cloud pair (n points xyz1, m points xyz2), exactly the reference's:
//   remainL[k] = multiL, remainR[l] = multiR               (integer ratios n/m, m/n as in :5-12)
//   for level in -4^7, -4^6, ..., -4^-1:
//     (1) ratioL[k]  = remainL[k] / (1e-9 + sum_l exp(level*d2(k,l)) * remainR[l])
//     (2) sumr       = remainR[l] * sum_k exp(level*d2) * ratioL[k];  ratioR[l] = min(remainR[l]/(sumr+1e-9), 1) * remainR[l];
//         remainR[l] = max(0, remainR[l] - sumr)
//     (3) w(k,l)     = exp(level*d2) * ratioL[k] * ratioR[l];  match[l,k] += w;  remainL[k] = max(0, remainL[k] - sum_l w)
//   cost = sum_{k,l} match[l,k] * sqrt(d2(k,l))
//
// B200 design: one 1024-thread CTA per cloud pair (grid = pairs, 148 pairs in flight instead of 32), both clouds
// resident in shared memory as float4 (xyz + the per-point factor the current pass needs), every thread owns PPT
// points of each set and keeps their remain/ratio state in registers.  With `match == nullptr` the cost is accumulated
// inside pass (3) and the dense match matrix (16 MB per pair at 2048 x 2048 points) is never materialised; the sum is
// then associated per level instead of per matrix entry, so the cost agrees with the reference to fp32 rounding
// (tests: <= 2e-5 relative against the reference's own kernels), not bit for bit.  exp is __expf, as in the reference.
#include <stdlib.h>

#include "common.cuh"
#include "ldt_b200.h"

namespace ldt {

constexpr int EMD_THREADS = 1024;

struct EmdPairMap {   // which cloud of each set CTA `blockIdx.x` works on
  int nb;             // > 0: pairwise mode, pair = (row_begin + blockIdx.x / nb, blockIdx.x % nb); 0: batched mode (i, i)
  int row_begin;
};

template <int PPT>
__global__ void __launch_bounds__(EMD_THREADS, 1)
approx_match_kernel(int n, int m, const float* __restrict__ xyz1, const float* __restrict__ xyz2, EmdPairMap map,
                    float* __restrict__ match, float* __restrict__ cost, float cost_scale) {
  extern __shared__ float4 emd_smem[];
  float4* s1 = emd_smem;        // [n]  xyz1 + ratioL
  float4* s2 = emd_smem + n;    // [m]  xyz2 + (remainR | ratioR)
  __shared__ double red[EMD_THREADS / 32];

  const int tid = threadIdx.x;
  const int i1 = map.nb > 0 ? map.row_begin + blockIdx.x / map.nb : blockIdx.x;
  const int i2 = map.nb > 0 ? blockIdx.x % map.nb : blockIdx.x;
  const float* p1 = xyz1 + static_cast<size_t>(i1) * n * 3;
  const float* p2 = xyz2 + static_cast<size_t>(i2) * m * 3;
  float* mt = match ? match + static_cast<size_t>(blockIdx.x) * n * m : nullptr;

  float multiL, multiR;
  if (n >= m) { multiL = 1.f; multiR = static_cast<float>(n / m); }
  else        { multiL = static_cast<float>(m / n); multiR = 1.f; }

  float x1[PPT], y1[PPT], z1[PPT], remL[PPT], ratL[PPT];
  float x2[PPT], y2[PPT], z2[PPT], remR[PPT];
  bool ok1[PPT], ok2[PPT];
#pragma unroll
  for (int u = 0; u < PPT; ++u) {
    const int k = tid + u * EMD_THREADS;
    ok1[u] = k < n;
    ok2[u] = k < m;
    x1[u] = ok1[u] ? p1[k * 3 + 0] : 0.f; y1[u] = ok1[u] ? p1[k * 3 + 1] : 0.f; z1[u] = ok1[u] ? p1[k * 3 + 2] : 0.f;
    x2[u] = ok2[u] ? p2[k * 3 + 0] : 0.f; y2[u] = ok2[u] ? p2[k * 3 + 1] : 0.f; z2[u] = ok2[u] ? p2[k * 3 + 2] : 0.f;
    remL[u] = ok1[u] ? multiL : 0.f;   // phantom points (k >= n, l >= m) carry no mass: every w they produce is 0
    remR[u] = ok2[u] ? multiR : 0.f;
    ratL[u] = 0.f;
    if (ok1[u]) s1[k] = make_float4(x1[u], y1[u], z1[u], 0.f);
    if (ok2[u]) s2[k] = make_float4(x2[u], y2[u], z2[u], multiR);
  }
  if (mt != nullptr)
    for (size_t e = tid; e < static_cast<size_t>(n) * m; e += EMD_THREADS) mt[e] = 0.f;
  __syncthreads();

  float my_cost = 0.f;
  for (int j = 7; j > -2; --j) {
    const float level = -powf(4.0f, static_cast<float>(j));
    // ---- (1) left ratios ----
    {
      float suml[PPT];
#pragma unroll
      for (int u = 0; u < PPT; ++u) suml[u] = 1e-9f;
#pragma unroll 4
      for (int l = 0; l < m; ++l) {
        const float4 q = s2[l];
#pragma unroll
        for (int u = 0; u < PPT; ++u) {
          const float dx = q.x - x1[u], dy = q.y - y1[u], dz = q.z - z1[u];
          const float d = level * (dx * dx + dy * dy + dz * dz);
          suml[u] += __expf(d) * q.w;
        }
      }
#pragma unroll
      for (int u = 0; u < PPT; ++u) {
        ratL[u] = remL[u] / suml[u];
        if (ok1[u]) s1[tid + u * EMD_THREADS].w = ratL[u];
      }
    }
    __syncthreads();
    // ---- (2) right consumption ----
    {
      float sumr[PPT];
#pragma unroll
      for (int u = 0; u < PPT; ++u) sumr[u] = 0.f;
#pragma unroll 4
      for (int k = 0; k < n; ++k) {
        const float4 q = s1[k];
#pragma unroll
        for (int u = 0; u < PPT; ++u) {
          const float dx = x2[u] - q.x, dy = y2[u] - q.y, dz = z2[u] - q.z;
          sumr[u] += __expf(level * (dx * dx + dy * dy + dz * dz)) * q.w;
        }
      }
#pragma unroll
      for (int u = 0; u < PPT; ++u) {
        const float sr = sumr[u] * remR[u];
        const float consumption = fminf(remR[u] / (sr + 1e-9f), 1.0f);
        const float ratR = consumption * remR[u];
        remR[u] = fmaxf(0.0f, remR[u] - sr);
        if (ok2[u]) s2[tid + u * EMD_THREADS].w = ratR;
      }
    }
    __syncthreads();
    // ---- (3) matched mass (+ cost) ----
    {
      float suml[PPT];
#pragma unroll
      for (int u = 0; u < PPT; ++u) suml[u] = 0.f;
#pragma unroll 2
      for (int l = 0; l < m; ++l) {
        const float4 q = s2[l];
#pragma unroll
        for (int u = 0; u < PPT; ++u) {
          const float dx = q.x - x1[u], dy = q.y - y1[u], dz = q.z - z1[u];
          const float d2 = dx * dx + dy * dy + dz * dz;
          const float w = __expf(level * d2) * ratL[u] * q.w;
          suml[u] += w;
          if (mt != nullptr) {
            if (ok1[u]) mt[static_cast<size_t>(l) * n + tid + u * EMD_THREADS] += w;
          } else {
            my_cost = fmaf(w, sqrtf(d2), my_cost);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < PPT; ++u) remL[u] = fmaxf(0.0f, remL[u] - suml[u]);
    }
    __syncthreads();
    // next level's pass (1) reads the updated right remainders next to xyz2
#pragma unroll
    for (int u = 0; u < PPT; ++u)
      if (ok2[u]) s2[tid + u * EMD_THREADS].w = remR[u];
    __syncthreads();
  }

  if (cost != nullptr && mt == nullptr) {
    // fp32 per thread (like the reference's per-thread subsum), fp64 across the CTA in a fixed order: deterministic
    double c = static_cast<double>(my_cost);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((tid & 31) == 0) red[tid >> 5] = c;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < EMD_THREADS / 32; ++w) t += red[w];
      cost[blockIdx.x] = static_cast<float>(t) * cost_scale;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// The same kernel for clouds of up to 2048 points (two points of each set per thread) on Blackwell's packed fp32
// arithmetic: the two points a thread owns form one f32x2 operand, the streamed point sits in shared memory with every
// component duplicated, (x,x,y,y) + (z,z,w,w), so a broadcast load is the packed operand.  Per point pair and pass the
// scalar kernel issues ~10 FP32 instructions + one ex2 and is issue-bound; packed it is 4-5 issue slots and the ex2
// unit (16 lanes/clk/SM) becomes the bound.  exp(level * d2) is evaluated as ex2(d2 * (level * log2 e)) with the
// product of constants formed once per level; sqrt in the cost as sqrt.approx -- both inside the 2e-5 agreement the
// tests demand against the reference's own kernels.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t ex2_x2(uint64_t a) {
  float lo, hi;
  unpack_f32x2(a, lo, hi);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(lo) : "f"(lo));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(hi) : "f"(hi));
  return pack_f32x2(lo, hi);
}
__device__ __forceinline__ uint64_t ld_pair_xy(const float4* p, uint64_t& yy) {   // (x,x,y,y) -> xx, yy
  const float4 v = *p;
  yy = pack_f32x2(v.z, v.w);
  return pack_f32x2(v.x, v.y);
}

__global__ void __launch_bounds__(EMD_THREADS, 1)
approx_match_x2_kernel(int n, int m, const float* __restrict__ xyz1, const float* __restrict__ xyz2, EmdPairMap map,
                       float* __restrict__ match, float* __restrict__ cost, float cost_scale) {
  extern __shared__ float4 emd_smem[];
  float4* s1a = emd_smem;            // [n] (x,x,y,y) of xyz1
  float4* s1b = s1a + n;             // [n] (z,z,w,w): w = ratioL
  float4* s2a = s1b + n;             // [m] (x,x,y,y) of xyz2
  float4* s2b = s2a + m;             // [m] (z,z,w,w): w = remainR | ratioR
  __shared__ double red[EMD_THREADS / 32];

  const int tid = threadIdx.x;
  const int i1 = map.nb > 0 ? map.row_begin + blockIdx.x / map.nb : blockIdx.x;
  const int i2 = map.nb > 0 ? blockIdx.x % map.nb : blockIdx.x;
  const float* p1 = xyz1 + static_cast<size_t>(i1) * n * 3;
  const float* p2 = xyz2 + static_cast<size_t>(i2) * m * 3;
  float* mt = match ? match + static_cast<size_t>(blockIdx.x) * n * m : nullptr;

  float multiL, multiR;
  if (n >= m) { multiL = 1.f; multiR = static_cast<float>(n / m); }
  else        { multiL = static_cast<float>(m / n); multiR = 1.f; }

  // this thread's two points of each set: k = tid, tid + 1024
  float c1[2][3], c2[2][3], remLs[2], remRs[2];
  bool ok1[2], ok2[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int k = tid + u * EMD_THREADS;
    ok1[u] = k < n;
    ok2[u] = k < m;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      c1[u][d] = ok1[u] ? p1[k * 3 + d] : 0.f;
      c2[u][d] = ok2[u] ? p2[k * 3 + d] : 0.f;
    }
    remLs[u] = ok1[u] ? multiL : 0.f;   // phantom points carry no mass
    remRs[u] = ok2[u] ? multiR : 0.f;
    if (ok1[u]) { s1a[k] = make_float4(c1[u][0], c1[u][0], c1[u][1], c1[u][1]); s1b[k] = make_float4(c1[u][2], c1[u][2], 0.f, 0.f); }
    if (ok2[u]) { s2a[k] = make_float4(c2[u][0], c2[u][0], c2[u][1], c2[u][1]); s2b[k] = make_float4(c2[u][2], c2[u][2], multiR, multiR); }
  }
  const uint64_t x1 = pack_f32x2(c1[0][0], c1[1][0]), y1 = pack_f32x2(c1[0][1], c1[1][1]), z1 = pack_f32x2(c1[0][2], c1[1][2]);
  const uint64_t x2 = pack_f32x2(c2[0][0], c2[1][0]), y2 = pack_f32x2(c2[0][1], c2[1][1]), z2 = pack_f32x2(c2[0][2], c2[1][2]);
  uint64_t remL = pack_f32x2(remLs[0], remLs[1]);
  float remR[2] = {remRs[0], remRs[1]};
  float ratL[2] = {0.f, 0.f};
  if (mt != nullptr)
    for (size_t e = tid; e < static_cast<size_t>(n) * m; e += EMD_THREADS) mt[e] = 0.f;
  __syncthreads();

  float my_cost = 0.f;
  for (int j = 7; j > -2; --j) {
    const float level = -powf(4.0f, static_cast<float>(j));
    const float l2 = level * 1.4426950408889634f;
    const uint64_t lv = pack_f32x2(l2, l2);
    // ---- (1) left ratios ----
    {
      uint64_t suml = pack_f32x2(1e-9f, 1e-9f);
#pragma unroll 4
      for (int l = 0; l < m; ++l) {
        uint64_t qy, qw;
        const uint64_t qx = ld_pair_xy(s2a + l, qy);
        const uint64_t qz = ld_pair_xy(s2b + l, qw);
        const uint64_t dx = sub2(qx, x1), dy = sub2(qy, y1), dz = sub2(qz, z1);
        const uint64_t d2 = fma_f32x2(dz, dz, fma_f32x2(dy, dy, mul2(dx, dx)));
        suml = fma_f32x2(ex2_x2(mul2(d2, lv)), qw, suml);
      }
      float s0, s1, r0, r1;
      unpack_f32x2(suml, s0, s1);
      unpack_f32x2(remL, r0, r1);
      ratL[0] = r0 / s0;
      ratL[1] = r1 / s1;
#pragma unroll
      for (int u = 0; u < 2; ++u)
        if (ok1[u]) { float4 v = s1b[tid + u * EMD_THREADS]; v.z = ratL[u]; v.w = ratL[u]; s1b[tid + u * EMD_THREADS] = v; }
    }
    __syncthreads();
    // ---- (2) right consumption ----
    {
      uint64_t sumr = pack_f32x2(0.f, 0.f);
#pragma unroll 4
      for (int k = 0; k < n; ++k) {
        uint64_t qy, qw;
        const uint64_t qx = ld_pair_xy(s1a + k, qy);
        const uint64_t qz = ld_pair_xy(s1b + k, qw);
        const uint64_t dx = sub2(x2, qx), dy = sub2(y2, qy), dz = sub2(z2, qz);
        const uint64_t d2 = fma_f32x2(dz, dz, fma_f32x2(dy, dy, mul2(dx, dx)));
        sumr = fma_f32x2(ex2_x2(mul2(d2, lv)), qw, sumr);
      }
      float sr[2];
      unpack_f32x2(sumr, sr[0], sr[1]);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const float s = sr[u] * remR[u];
        const float consumption = fminf(remR[u] / (s + 1e-9f), 1.0f);
        const float ratR = consumption * remR[u];
        remR[u] = fmaxf(0.0f, remR[u] - s);
        if (ok2[u]) { float4 v = s2b[tid + u * EMD_THREADS]; v.z = ratR; v.w = ratR; s2b[tid + u * EMD_THREADS] = v; }
      }
    }
    __syncthreads();
    // ---- (3) matched mass (+ cost) ----
    {
      uint64_t suml = pack_f32x2(0.f, 0.f);
      const uint64_t rl = pack_f32x2(ratL[0], ratL[1]);
#pragma unroll 2
      for (int l = 0; l < m; ++l) {
        uint64_t qy, qw;
        const uint64_t qx = ld_pair_xy(s2a + l, qy);
        const uint64_t qz = ld_pair_xy(s2b + l, qw);
        const uint64_t dx = sub2(qx, x1), dy = sub2(qy, y1), dz = sub2(qz, z1);
        const uint64_t d2 = fma_f32x2(dz, dz, fma_f32x2(dy, dy, mul2(dx, dx)));
        const uint64_t w = mul2(mul2(ex2_x2(mul2(d2, lv)), rl), qw);
        suml = add_f32x2(suml, w);
        float w0, w1, d0, d1;
        unpack_f32x2(w, w0, w1);
        if (mt != nullptr) {
          if (ok1[0]) mt[static_cast<size_t>(l) * n + tid] += w0;
          if (ok1[1]) mt[static_cast<size_t>(l) * n + tid + EMD_THREADS] += w1;
        } else {
          unpack_f32x2(d2, d0, d1);
          float q0, q1;
          asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(q0) : "f"(d0));
          asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(q1) : "f"(d1));
          my_cost = fmaf(w0, q0, my_cost);
          my_cost = fmaf(w1, q1, my_cost);
        }
      }
      float r0, r1, s0, s1;
      unpack_f32x2(remL, r0, r1);
      unpack_f32x2(suml, s0, s1);
      remL = pack_f32x2(fmaxf(0.0f, r0 - s0), fmaxf(0.0f, r1 - s1));
    }
    __syncthreads();
    // next level's pass (1) reads the updated right remainders next to xyz2
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (ok2[u]) { float4 v = s2b[tid + u * EMD_THREADS]; v.z = remR[u]; v.w = remR[u]; s2b[tid + u * EMD_THREADS] = v; }
    __syncthreads();
  }

  if (cost != nullptr && mt == nullptr) {
    double c = static_cast<double>(my_cost);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((tid & 31) == 0) red[tid >> 5] = c;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < EMD_THREADS / 32; ++w) t += red[w];
      cost[blockIdx.x] = static_cast<float>(t) * cost_scale;
    }
  }
}

// cost[i] = sum_{k,l} match[i, l, k] * |xyz1[i,k] - xyz2[i,l]|   (matchcostkernel, approxmatch.cu:184-224)
__global__ void __launch_bounds__(512) match_cost_kernel(int n, int m, const float* __restrict__ xyz1,
                                                       const float* __restrict__ xyz2, const float* __restrict__ match,
                                                       float* __restrict__ out) {
  __shared__ double red[16];
  const int i = blockIdx.x;
  const float* p1 = xyz1 + static_cast<size_t>(i) * n * 3;
  const float* p2 = xyz2 + static_cast<size_t>(i) * m * 3;
  const float* mt = match + static_cast<size_t>(i) * n * m;
  double acc = 0.0;
  for (int l = 0; l < m; ++l) {
    const float x2 = p2[l * 3 + 0], y2 = p2[l * 3 + 1], z2 = p2[l * 3 + 2];
    float sub = 0.f;
    for (int k = threadIdx.x; k < n; k += 512) {   // coalesced along k, the contiguous axis of match
      const float dx = x2 - p1[k * 3 + 0], dy = y2 - p1[k * 3 + 1], dz = z2 - p1[k * 3 + 2];
      sub = fmaf(mt[static_cast<size_t>(l) * n + k], sqrtf(dx * dx + dy * dy + dz * dz), sub);
    }
    acc += static_cast<double>(sub);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 16; ++w) t += red[w];
    out[i] = static_cast<float>(t);
  }
}

static int launch_approx_match(int pairs, int n, int m, const float* xyz1, const float* xyz2, EmdPairMap map, float* match,
                               float* cost, float cost_scale, cudaStream_t s, const char* who) {
  const int mx = n > m ? n : m;
  const int ppt = (mx + EMD_THREADS - 1) / EMD_THREADS;
  LDT_REQUIRE(ppt <= 4, LDT_ERR_UNSUPPORTED, "%s: at most %d points per cloud (got n=%d m=%d)", who, 4 * EMD_THREADS, n, m);
  const size_t smem = static_cast<size_t>(n + m) * sizeof(float4);
#define LDT_EMD_LAUNCH(P)                                                                                             \
  do {                                                                                                                \
    static PerDevice<bool> attr;                                                                                      \
    if (!attr.get()) {                                                                                                \
      LDT_CUDA_OK(cudaFuncSetAttribute(approx_match_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); \
      attr.get() = true;                                                                                              \
    }                                                                                                                 \
    approx_match_kernel<P><<<pairs, EMD_THREADS, smem, s>>>(n, m, xyz1, xyz2, map, match, cost, cost_scale);          \
  } while (0)
  static int scalar_only = -1;
  if (scalar_only < 0) {
    const char* e = getenv("LDT_EMD_SCALAR");   // A/B knob: 1 = the scalar kernel for every size
    scalar_only = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  if (ppt == 2 && !scalar_only) {   // the 1025..2048-point case (ShapeNet clouds): packed-fp32 kernel
    static PerDevice<bool> attr;
    if (!attr.get()) {
      LDT_CUDA_OK(cudaFuncSetAttribute(approx_match_x2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      attr.get() = true;
    }
    approx_match_x2_kernel<<<pairs, EMD_THREADS, 2 * smem, s>>>(n, m, xyz1, xyz2, map, match, cost, cost_scale);
  } else if (ppt <= 1) LDT_EMD_LAUNCH(1);
  else if (ppt == 2) LDT_EMD_LAUNCH(2);
  else LDT_EMD_LAUNCH(4);
#undef LDT_EMD_LAUNCH
  LDT_CUDA_OK(cudaGetLastError());
  return LDT_OK;
}

}  // namespace ldt

using namespace ldt;

extern "C" int ldt_match_cost(int b, int n, int m, const float* xyz1, const float* xyz2, float* cost, void* stream) {
  LDT_REQUIRE(b >= 0 && n > 0 && m > 0, LDT_ERR_INVALID, "ldt_match_cost: bad sizes b=%d n=%d m=%d", b, n, m);
  if (b == 0) return LDT_OK;
  LDT_REQUIRE(xyz1 && xyz2 && cost, LDT_ERR_INVALID, "ldt_match_cost: null pointer");
  EmdPairMap map{0, 0};
  return launch_approx_match(b, n, m, xyz1, xyz2, map, nullptr, cost, 1.0f, static_cast<cudaStream_t>(stream), "ldt_match_cost");
}

extern "C" int ldt_approx_match(int b, int n, int m, const float* xyz1, const float* xyz2, float* match, void* stream) {
  LDT_REQUIRE(b >= 0 && n > 0 && m > 0, LDT_ERR_INVALID, "ldt_approx_match: bad sizes b=%d n=%d m=%d", b, n, m);
  if (b == 0) return LDT_OK;
  LDT_REQUIRE(xyz1 && xyz2 && match, LDT_ERR_INVALID, "ldt_approx_match: null pointer");
  EmdPairMap map{0, 0};
  return launch_approx_match(b, n, m, xyz1, xyz2, map, match, nullptr, 1.0f, static_cast<cudaStream_t>(stream), "ldt_approx_match");
}

extern "C" int ldt_match_cost_from_match(int b, int n, int m, const float* xyz1, const float* xyz2, const float* match,
                                         float* cost, void* stream) {
  LDT_REQUIRE(b >= 0 && n > 0 && m > 0, LDT_ERR_INVALID, "ldt_match_cost_from_match: bad sizes b=%d n=%d m=%d", b, n, m);
  if (b == 0) return LDT_OK;
  LDT_REQUIRE(xyz1 && xyz2 && match && cost, LDT_ERR_INVALID, "ldt_match_cost_from_match: null pointer");
  match_cost_kernel<<<b, 512, 0, static_cast<cudaStream_t>(stream)>>>(n, m, xyz1, xyz2, match, cost);
  LDT_CUDA_OK(cudaGetLastError());
  return LDT_OK;
}

extern "C" int ldt_pairwise_emd(int na, int nb, int p, const float* a, const float* b, int row_begin, int row_end,
                                float* out, void* stream) {
  LDT_REQUIRE(na >= 0 && nb >= 0 && p > 0, LDT_ERR_INVALID, "ldt_pairwise_emd: bad sizes na=%d nb=%d p=%d", na, nb, p);
  LDT_REQUIRE(0 <= row_begin && row_begin <= row_end && row_end <= na, LDT_ERR_INVALID,
              "ldt_pairwise_emd: row range [%d,%d) outside [0,%d)", row_begin, row_end, na);
  const long long pairs = static_cast<long long>(row_end - row_begin) * nb;
  if (pairs == 0) return LDT_OK;
  LDT_REQUIRE(a && b && out, LDT_ERR_INVALID, "ldt_pairwise_emd: null pointer");
  LDT_REQUIRE(pairs < (1LL << 31), LDT_ERR_INVALID, "ldt_pairwise_emd: too many pairs in one call");
  EmdPairMap map{nb, row_begin};
  // emd_approx_cuda divides the match cost by the number of points (evaluation_metrics.py:41-45)
  return launch_approx_match(static_cast<int>(pairs), p, p, a, b, map, nullptr, out, 1.0f / static_cast<float>(p),
                             static_cast<cudaStream_t>(stream), "ldt_pairwise_emd");
}
