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
    static bool attr = false;                                                                                         \
    if (!attr) {                                                                                                      \
      LDT_CUDA_OK(cudaFuncSetAttribute(approx_match_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); \
      attr = true;                                                                                                    \
    }                                                                                                                 \
    approx_match_kernel<P><<<pairs, EMD_THREADS, smem, s>>>(n, m, xyz1, xyz2, map, match, cost, cost_scale);          \
  } while (0)
  if (ppt <= 1) LDT_EMD_LAUNCH(1);
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
