// Nearest-neighbour squared distances between point sets (Chamfer building block) and the fused
// pairwise Chamfer-distance matrix used by 1-NNA / COV / MMD.
//
// Replaces: reference evaluation/pytorch_structural_losses/src/nndistance.cu:2-128 (NmDistanceKernel,
// launched once per direction) and the Python double loop around it in
// evaluation/evaluation_metrics.py:165-198 (_pairwise_CD_).
//
// Bit-exactness contract (checked in tests/test_gpu_kernels.py against oracle/nn_oracle.c and, on the
// GPU box, against the reference kernel itself built into oracle/_ref):
//   d(p,q) = fma(dz,dz, fma(dx,dx, dy*dy))  with d* = q* - p*   -- the contraction nvcc 12.9 emits for
//   the reference source line `x2*x2+y2*y2+z2*z2` (verified in its PTX); the squares make the sign of
//   the subtraction irrelevant, so one evaluation serves both directions.
//   argmin ties: lowest candidate index wins (strict `<` in-tile, strict `>` across tiles).
#include "common.cuh"
#include "ldt_b200.h"

namespace ldt {

__device__ __forceinline__ float sqdist_ref(float dx, float dy, float dz) {
  return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// The same distance for TWO queries against one candidate with Blackwell's packed fp32 arithmetic (FADD2 / FMUL2 / FFMA2:
// one issue slot, two results).  add/mul/fma.rn.f32x2 round each half exactly like their scalar forms, so both halves are
// bit-identical to sqdist_ref.
__device__ __forceinline__ uint64_t sub_f32x2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t mul_f32x2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t sqdist_ref_x2(uint64_t dx, uint64_t dy, uint64_t dz) {
  return fma_f32x2(dz, dz, fma_f32x2(dx, dx, mul_f32x2(dy, dy)));
}

// ------------------------------------------------------------------------------------------------
// Kernel 1: one direction with argmin (drop-in for NmDistanceKernel).  grid = (query tiles, batch).
// Each thread owns QPT query points in registers; candidates stream through shared memory as
// padded float4 so that one broadcast LDS.128 feeds QPT distance evaluations.
// ------------------------------------------------------------------------------------------------
constexpr int NN_THREADS = 128;
constexpr int NN_QPT = 2;
constexpr int NN_TILE = 1024;  // candidates per smem tile (16 KB)

__global__ void __launch_bounds__(NN_THREADS) nn_argmin_kernel(int n, const float* __restrict__ xyz, int m,
                                                             const float* __restrict__ xyz2,
                                                             float* __restrict__ result,
                                                             int* __restrict__ result_i) {
  __shared__ float4 cand_xy[NN_TILE];   // (x, x, y, y): a broadcast load is the packed operand of both queries
  __shared__ float2 cand_z[NN_TILE];    // (z, z)
  const int b = blockIdx.y;
  const float* q = xyz + static_cast<size_t>(b) * n * 3;
  const float* c = xyz2 + static_cast<size_t>(b) * m * 3;
  const int q0 = blockIdx.x * (NN_THREADS * NN_QPT) + threadIdx.x;

  float qx[NN_QPT], qy[NN_QPT], qz[NN_QPT], best[NN_QPT];
  int best_i[NN_QPT];
#pragma unroll
  for (int u = 0; u < NN_QPT; ++u) {
    const int j = q0 + u * NN_THREADS;
    const bool ok = j < n;
    qx[u] = ok ? q[j * 3 + 0] : 0.f;
    qy[u] = ok ? q[j * 3 + 1] : 0.f;
    qz[u] = ok ? q[j * 3 + 2] : 0.f;
    best[u] = 0.f;
    best_i[u] = 0;
  }
  static_assert(NN_QPT == 2, "the two queries of a thread form one packed operand");
  const uint64_t qx2 = pack_f32x2(qx[0], qx[1]), qy2 = pack_f32x2(qy[0], qy[1]), qz2 = pack_f32x2(qz[0], qz[1]);
  for (int k2 = 0; k2 < m; k2 += NN_TILE) {
    const int cnt = min(NN_TILE, m - k2);
    __syncthreads();
    for (int k = threadIdx.x; k < cnt; k += NN_THREADS) {
      const float* p = c + static_cast<size_t>(k2 + k) * 3;
      cand_xy[k] = make_float4(p[0], p[0], p[1], p[1]);
      cand_z[k] = make_float2(p[2], p[2]);
    }
    __syncthreads();
    if (k2 == 0) {  // candidate 0 initialises the running minimum (reference: `k==0 || d<best`)
      const float4 p = cand_xy[0];
      const float pz = cand_z[0].x;
#pragma unroll
      for (int u = 0; u < NN_QPT; ++u) best[u] = sqdist_ref(p.x - qx[u], p.z - qy[u], pz - qz[u]);
    }
#pragma unroll 4
    for (int k = 0; k < cnt; ++k) {
      // both queries against candidate k in one packed evaluation (each half rounds like sqdist_ref: bit-identical)
      const float4 a = cand_xy[k];
      const float2 z = cand_z[k];
      float d0, d1;
      unpack_f32x2(sqdist_ref_x2(sub_f32x2(pack_f32x2(a.x, a.y), qx2), sub_f32x2(pack_f32x2(a.z, a.w), qy2),
                                 sub_f32x2(pack_f32x2(z.x, z.y), qz2)), d0, d1);
      if (d0 < best[0]) {
        best[0] = d0;
        best_i[0] = k2 + k;
      }
      if (d1 < best[1]) {
        best[1] = d1;
        best_i[1] = k2 + k;
      }
    }
  }
#pragma unroll
  for (int u = 0; u < NN_QPT; ++u) {
    const int j = q0 + u * NN_THREADS;
    if (j < n) {
      result[static_cast<size_t>(b) * n + j] = best[u];
      result_i[static_cast<size_t>(b) * n + j] = best_i[u];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Kernel 2: fused pairwise Chamfer matrix.  One CTA per (row cloud i, column cloud j):
//   out[i,j] = mean_p min_q d(a_i[p], b_j[q]) + mean_q min_p d(a_i[p], b_j[q])
// Every point-pair distance is evaluated ONCE and feeds both the per-query running minimum
// (registers) and the per-candidate minimum (warp REDUX.MIN on the float bit pattern -- distances are
// non-negative so unsigned order == float order -- then one shared-memory atomicMin per warp).
// No indices are produced: _pairwise_CD_ discards them (evaluation_metrics.py:190-191).
// The six FP32 operations of a distance run packed, two queries per instruction (candidates sit in shared memory with
// every coordinate duplicated, (x,x,y,y) + (z,z), so a broadcast load IS the packed operand): 3 issue slots per point
// pair for the distance + 1 for the two 3-input minima, where the scalar form needed 6 + 1.
// ------------------------------------------------------------------------------------------------
constexpr int CD_THREADS = 256;
constexpr int CD_QPT = 8;                         // query points per thread
constexpr int CD_MAXP = CD_THREADS * CD_QPT;      // 2048 points per cloud handled in one pass

// row_step / upper: the SYMMETRIC form (A == B, the rr / ss matrices of compute_CD_metrics, evaluation_metrics.py:311-312):
// grid row y is cloud i = row_begin + y * row_step (rows interleaved over ranks balance the triangle), CTAs below the
// diagonal (j < i) exit at once, and the result lands at out[i * ncols_out + j] of the FULL matrix; the caller mirrors.
// Mirroring is exact, not approximate: the two sums below run over rowmin[] and colmin[] in the SAME index order, and
// rowmin of (i, j) == colmin of (j, i) element for element (a minimum does not depend on evaluation order), so the kernel
// returns M[i][j] == M[j][i] bit for bit whichever of the two it is asked for.
__global__ void __launch_bounds__(CD_THREADS) pairwise_cd_kernel(int nb, int pa, int pb,
                                                                const float* __restrict__ A,
                                                                const float* __restrict__ B, int row_begin,
                                                                int ncols_out, int col_begin, int row_step, int upper,
                                                                float* __restrict__ out) {
  const int i = row_begin + blockIdx.y * row_step;
  const int j = col_begin + blockIdx.x;
  if (upper && j < i) return;
  extern __shared__ float4 cd_smem[];
  const int pb32 = (pb + 31) & ~31;                              // candidates padded to whole groups of 32
  float4* cand_xy = cd_smem;                                     // [pb32] (x, x, y, y)
  float2* cand_z = reinterpret_cast<float2*>(cand_xy + pb32);    // [pb32] (z, z)
  unsigned* colmin = reinterpret_cast<unsigned*>(cand_z + pb32); // [pb32]
  float* rowmin = reinterpret_cast<float*>(colmin + pb32);       // [pa]
  __shared__ double red[2][CD_THREADS / 32];

  const float* a = A + static_cast<size_t>(i) * pa * 3;
  const float* b = B + static_cast<size_t>(j) * pb * 3;
  const int lane = threadIdx.x & 31;

  const float inf = __int_as_float(0x7f800000);
  for (int k = threadIdx.x; k < pb32; k += CD_THREADS) {
    // padding candidates sit at +inf: their distance to anything is +inf, so they never win a minimum
    const float cx = (k < pb) ? b[k * 3 + 0] : inf, cy = (k < pb) ? b[k * 3 + 1] : inf, cz = (k < pb) ? b[k * 3 + 2] : inf;
    cand_xy[k] = make_float4(cx, cx, cy, cy);
    cand_z[k] = make_float2(cz, cz);
    colmin[k] = 0x7f800000u;  // +inf
  }
  __syncthreads();

  // queries are processed in passes of CD_MAXP so any pa works; pa == 2048 is a single pass.
  for (int base = 0; base < pa; base += CD_MAXP) {
    float qx[CD_QPT], qy[CD_QPT], qz[CD_QPT], best[CD_QPT];
    bool ok[CD_QPT];
#pragma unroll
    for (int u = 0; u < CD_QPT; ++u) {
      const int p = base + u * CD_THREADS + threadIdx.x;
      ok[u] = p < pa;
      const int pc = ok[u] ? p : 0;   // padded lanes mirror point 0 of this cloud: harmless for minima
      qx[u] = a[pc * 3 + 0];
      qy[u] = a[pc * 3 + 1];
      qz[u] = a[pc * 3 + 2];
      best[u] = inf;
    }
    uint64_t qx2[CD_QPT / 2], qy2[CD_QPT / 2], qz2[CD_QPT / 2];   // queries (u, u+1) packed
#pragma unroll
    for (int u = 0; u < CD_QPT; u += 2) {
      qx2[u / 2] = pack_f32x2(qx[u], qx[u + 1]);
      qy2[u / 2] = pack_f32x2(qy[u], qy[u + 1]);
      qz2[u / 2] = pack_f32x2(qz[u], qz[u + 1]);
    }
    // Candidates go two at a time so that every running minimum is a 3-input FMNMX3 (min of the old value and two
    // new distances): 0.5 min instructions per point pair for the row minima and 0.5 for the column minima, on top of
    // the 6 FP32 operations of the distance itself.  The warp-wide column minimum (REDUX) of candidate k0+j is parked
    // in lane j; one shared-memory atomicMin per 32 candidates publishes them (4 per-candidate bookkeeping
    // instructions instead of a per-candidate atomic).
    for (int k0 = 0; k0 < pb32; k0 += 32) {
      unsigned mycol = 0x7f800000u;
#pragma unroll
      for (int jj = 0; jj < 32; jj += 2) {
        const float4 a0 = cand_xy[k0 + jj], a1 = cand_xy[k0 + jj + 1];
        const float2 b0 = cand_z[k0 + jj], b1 = cand_z[k0 + jj + 1];
        const uint64_t p0x = pack_f32x2(a0.x, a0.y), p0y = pack_f32x2(a0.z, a0.w), p0z = pack_f32x2(b0.x, b0.y);
        const uint64_t p1x = pack_f32x2(a1.x, a1.y), p1y = pack_f32x2(a1.z, a1.w), p1z = pack_f32x2(b1.x, b1.y);
        float cm0 = inf, cm1 = inf;
#pragma unroll
        for (int u = 0; u < CD_QPT; u += 2) {
          float d00, d01, d10, d11;
          unpack_f32x2(sqdist_ref_x2(sub_f32x2(p0x, qx2[u / 2]), sub_f32x2(p0y, qy2[u / 2]), sub_f32x2(p0z, qz2[u / 2])), d00, d01);
          unpack_f32x2(sqdist_ref_x2(sub_f32x2(p1x, qx2[u / 2]), sub_f32x2(p1y, qy2[u / 2]), sub_f32x2(p1z, qz2[u / 2])), d10, d11);
          best[u] = fminf(fminf(best[u], d00), d10);
          best[u + 1] = fminf(fminf(best[u + 1], d01), d11);
          cm0 = fminf(fminf(cm0, d00), d01);
          cm1 = fminf(fminf(cm1, d10), d11);
        }
        const unsigned w0 = __reduce_min_sync(0xffffffffu, __float_as_uint(cm0));
        const unsigned w1 = __reduce_min_sync(0xffffffffu, __float_as_uint(cm1));
        if (lane == jj) mycol = w0;
        if (lane == jj + 1) mycol = w1;
      }
      atomicMin(&colmin[k0 + lane], mycol);
    }
#pragma unroll
    for (int u = 0; u < CD_QPT; ++u)
      if (ok[u]) rowmin[base + u * CD_THREADS + threadIdx.x] = best[u];
  }
  __syncthreads();
  // both means are summed in the same (index-strided, then fixed tree) order: see the symmetry note above
  double row_sum = 0.0, col_sum = 0.0;
  for (int k = threadIdx.x; k < pa; k += CD_THREADS) row_sum += static_cast<double>(rowmin[k]);
  for (int k = threadIdx.x; k < pb; k += CD_THREADS) col_sum += static_cast<double>(__uint_as_float(colmin[k]));

  // block reduction (fixed order => deterministic)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    row_sum += __shfl_xor_sync(0xffffffffu, row_sum, o);
    col_sum += __shfl_xor_sync(0xffffffffu, col_sum, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = row_sum;
    red[1][threadIdx.x >> 5] = col_sum;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double rs = 0.0, cs = 0.0;
    for (int w = 0; w < CD_THREADS / 32; ++w) {
      rs += red[0][w];
      cs += red[1][w];
    }
    // dl.mean(dim=1) + dr.mean(dim=1)  (evaluation_metrics.py:191): two fp32 means, one fp32 add
    const float ml = static_cast<float>(rs / static_cast<double>(pa));
    const float mr = static_cast<float>(cs / static_cast<double>(pb));
    out[static_cast<size_t>(upper ? i : static_cast<int>(blockIdx.y)) * ncols_out + blockIdx.x] = __fadd_rn(ml, mr);
  }
}

}  // namespace ldt

using namespace ldt;

extern "C" int ldt_nn_distance(int b, int n, const float* xyz1, int m, const float* xyz2, float* dist1,
                               int* idx1, float* dist2, int* idx2, void* stream) {
  LDT_REQUIRE(b >= 0 && n >= 0 && m >= 0, LDT_ERR_INVALID, "ldt_nn_distance: negative size (b=%d n=%d m=%d)", b, n, m);
  if (b == 0) return LDT_OK;
  LDT_REQUIRE((n == 0 || (xyz1 && dist1 && idx1)) && (m == 0 || (xyz2 && dist2 && idx2)), LDT_ERR_INVALID,
              "ldt_nn_distance: null pointer");
  // An empty candidate set has no nearest neighbour; the reference leaves its outputs unwritten
  // (nndistance.cu:5 loop body never runs).  We refuse instead of returning garbage.
  LDT_REQUIRE((n > 0) == (m > 0), LDT_ERR_INVALID, "ldt_nn_distance: one point set is empty (n=%d m=%d)", n, m);
  if (n == 0) return LDT_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int per_block = NN_THREADS * NN_QPT;
  // grid.y <= 65535: larger batches go in chunks (the reference loops `for (int i = blockIdx.x; i < b; i += gridDim.x)`,
  // nndistance.cu:5)
  for (int b0 = 0; b0 < b; b0 += 32768) {
    const int nb = min(32768, b - b0);
    const float* x1 = xyz1 + static_cast<size_t>(b0) * n * 3;
    const float* x2 = xyz2 + static_cast<size_t>(b0) * m * 3;
    nn_argmin_kernel<<<dim3((n + per_block - 1) / per_block, nb), NN_THREADS, 0, s>>>(n, x1, m, x2, dist1 + static_cast<size_t>(b0) * n,
                                                                                     idx1 + static_cast<size_t>(b0) * n);
    nn_argmin_kernel<<<dim3((m + per_block - 1) / per_block, nb), NN_THREADS, 0, s>>>(m, x2, n, x1, dist2 + static_cast<size_t>(b0) * m,
                                                                                     idx2 + static_cast<size_t>(b0) * m);
  }
  LDT_CUDA_OK(cudaGetLastError());
  return LDT_OK;
}

static size_t cd_smem_bytes(int pa, int pb) {
  return static_cast<size_t>((pb + 31) & ~31) * (sizeof(float4) + sizeof(float2) + sizeof(unsigned)) +
         static_cast<size_t>(pa) * sizeof(float);
}

static int cd_prepare(const char* who, int pa, int pb, size_t* smem) {
  *smem = cd_smem_bytes(pa, pb);
  LDT_REQUIRE(*smem <= 200 * 1024, LDT_ERR_UNSUPPORTED, "%s: pa=%d pb=%d need %zu B of shared memory", who, pa, pb, *smem);
  static PerDevice<bool> attr_set;
  if (!attr_set.get()) {
    LDT_CUDA_OK(cudaFuncSetAttribute(pairwise_cd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set.get() = true;
  }
  return LDT_OK;
}

extern "C" int ldt_pairwise_cd(int na, int nb, int pa, int pb, const float* a, const float* b, int row_begin,
                               int row_end, float* out, void* stream) {
  LDT_REQUIRE(na >= 0 && nb >= 0 && pa > 0 && pb > 0, LDT_ERR_INVALID, "ldt_pairwise_cd: bad sizes na=%d nb=%d pa=%d pb=%d",
              na, nb, pa, pb);
  LDT_REQUIRE(0 <= row_begin && row_begin <= row_end && row_end <= na, LDT_ERR_INVALID,
              "ldt_pairwise_cd: row range [%d,%d) outside [0,%d)", row_begin, row_end, na);
  const int rows = row_end - row_begin;
  if (rows == 0 || nb == 0) return LDT_OK;
  LDT_REQUIRE(a && b && out, LDT_ERR_INVALID, "ldt_pairwise_cd: null pointer");
  size_t smem = 0;
  int rc = cd_prepare("ldt_pairwise_cd", pa, pb, &smem);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // grid.y is limited to 65535 rows per launch; chunk if a caller ever exceeds it.
  for (int r0 = 0; r0 < rows; r0 += 32768) {
    const int nr = min(32768, rows - r0);
    pairwise_cd_kernel<<<dim3(nb, nr), CD_THREADS, smem, s>>>(nb, pa, pb, a, b, row_begin + r0, nb, 0, 1, 0,
                                                              out + static_cast<size_t>(r0) * nb);
  }
  LDT_CUDA_OK(cudaGetLastError());
  return LDT_OK;
}

extern "C" int ldt_pairwise_cd_upper(int n, int p, const float* a, int row_first, int row_step, float* out, void* stream) {
  LDT_REQUIRE(n >= 0 && p > 0, LDT_ERR_INVALID, "ldt_pairwise_cd_upper: bad sizes n=%d p=%d", n, p);
  LDT_REQUIRE(row_step >= 1 && 0 <= row_first && row_first < row_step, LDT_ERR_INVALID,
              "ldt_pairwise_cd_upper: need 0 <= row_first (%d) < row_step (%d)", row_first, row_step);
  if (n == 0 || row_first >= n) return LDT_OK;
  LDT_REQUIRE(a && out, LDT_ERR_INVALID, "ldt_pairwise_cd_upper: null pointer");
  size_t smem = 0;
  int rc = cd_prepare("ldt_pairwise_cd_upper", p, p, &smem);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int rows = (n - row_first + row_step - 1) / row_step;   // rows row_first, row_first + row_step, ... < n
  for (int r0 = 0; r0 < rows; r0 += 32768) {
    const int nr = min(32768, rows - r0);
    pairwise_cd_kernel<<<dim3(n, nr), CD_THREADS, smem, s>>>(n, p, p, a, a, row_first + r0 * row_step, n, 0, row_step, 1, out);
  }
  LDT_CUDA_OK(cudaGetLastError());
  return LDT_OK;
}
