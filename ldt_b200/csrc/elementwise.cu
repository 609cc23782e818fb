// HBM-bound kernels of the score net / decoder / sampler: casts, LayerNorm+modulation, the time
// embedding MLP, and the fused reverse-SDE predictor update.  All are vectorised (128-bit) and use
// warp-shuffle reductions; none of them is GEMM-shaped.
#include <curand_kernel.h>

#include <algorithm>

#include "common.cuh"
#include "ldt_b200.h"

namespace ldt {

// ------------------------------------------------------------------------------------------------
// f32 [rows, cols] -> bf16 [rows, ld_out], zero-padding the tail columns.  Used for activations
// (x[B*32,120] -> A operand with K padded to 128) and for weight packing.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cast_pad_kernel(long long rows, int cols, const float* __restrict__ in, int ld_in,
                                                     __nv_bfloat16* __restrict__ out, int ld_out) {
  pdl_launch_dependents();
  pdl_wait();
  const int groups = ld_out >> 3;  // 8 outputs (16 bytes) per thread-iteration
  const long long total = rows * groups;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const long long r = i / groups;
    const int c0 = static_cast<int>(i - r * groups) << 3;
    const float* src = in + r * ld_in + c0;
    float v[8];
    if (c0 + 8 <= cols && (ld_in & 3) == 0) {
      const float4 a = *reinterpret_cast<const float4*>(src);
      const float4 b = *reinterpret_cast<const float4*>(src + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (c0 + j < cols) ? src[j] : 0.f;
    }
    __nv_bfloat162 h[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    *reinterpret_cast<uint4*>(out + r * ld_out + c0) = *reinterpret_cast<uint4*>(h);
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm(eps) over C channels, one warp per row, row held in registers (C <= 2048), then either
//   AdaLN:  y = n * (1 + scale[g]) + shift[g]      or      affine:  y = n * weight + bias
// ------------------------------------------------------------------------------------------------
constexpr int LN_MAXV = 16;  // float4 per lane => C <= 2048

// NV = C / 128 float4 per lane, a compile-time constant so the row lives in exactly NV*4 registers: at C = 1024 the
// kernel needs ~56 registers instead of the 96 a runtime-bounded v[16] costs, which more than doubles the resident
// warps per SM; the kernel is latency-bound (one dependent chain load -> 2 reductions -> modulation loads -> store
// per row), so resident warps are what hides it.  4 rows per 128-thread CTA keeps the tail wave short.
template <int NV>
__global__ void __launch_bounds__(128) layernorm_mod_kernel(int rows, const float* __restrict__ x,
                                                          const float* __restrict__ shift,
                                                          const float* __restrict__ scale, long long mod_stride,
                                                          int rows_per_mod, const float* __restrict__ weight,
                                                          const float* __restrict__ bias, float eps,
                                                          __nv_bfloat16* __restrict__ y) {
  constexpr int C = NV * 128;
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * C);
  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = __ldcg(xr + i * 32 + lane);   // streamed once: keep it out of L1
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) / static_cast<float>(C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(C) + eps);
  const float4* m_mul;
  const float4* m_add;
  const bool ada = (scale != nullptr);
  if (ada) {
    const long long g = static_cast<long long>(row / rows_per_mod) * mod_stride;
    m_mul = reinterpret_cast<const float4*>(scale + g);
    m_add = reinterpret_cast<const float4*>(shift + g);
  } else {
    m_mul = reinterpret_cast<const float4*>(weight);
    m_add = reinterpret_cast<const float4*>(bias);
  }
  uint2* yr = reinterpret_cast<uint2*>(y + static_cast<size_t>(row) * C);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int idx = i * 32 + lane;
    float4 mu = make_float4(1.f, 1.f, 1.f, 1.f), ad = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m_mul) mu = __ldg(m_mul + idx);
    if (m_add) ad = __ldg(m_add + idx);
    if (ada) { mu.x += 1.f; mu.y += 1.f; mu.z += 1.f; mu.w += 1.f; }
    const float o0 = (v[i].x - mean) * rstd * mu.x + ad.x;
    const float o1 = (v[i].y - mean) * rstd * mu.y + ad.y;
    const float o2 = (v[i].z - mean) * rstd * mu.z + ad.z;
    const float o3 = (v[i].w - mean) * rstd * mu.w + ad.w;
    __nv_bfloat162 h0 = __floats2bfloat162_rn(o0, o1), h1 = __floats2bfloat162_rn(o2, o3);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0);
    u.y = *reinterpret_cast<uint32_t*>(&h1);
    yr[idx] = u;
  }
}

// ------------------------------------------------------------------------------------------------
// Time embedding.  Stage A: sinusoidal features.  Stage B/C: skinny fp32 linear layers, one warp per
// output feature, 8 rows per pass so each weight row is read once per 8 conditioning rows.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sincos_kernel(int R, int half, const float* __restrict__ t,
                                                   const float* __restrict__ freq, float* __restrict__ emb) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= R * half) return;
  const int r = i / half, k = i - r * half;
  const float a = __fmul_rn(t[r], freq[k]);  // ts * t_emb  (model/layers.py:33)
  emb[static_cast<size_t>(r) * 2 * half + k] = sinf(a);
  emb[static_cast<size_t>(r) * 2 * half + half + k] = cosf(a);
}

// out[r, n] = act_out( sum_k in[r,k] * W[n,k] + b[n] (+ extra[r,n]) );  optionally also bf16(silu(out)).
// ACT: 0 = identity, 1 = SiLU.
template <int ACT>
__global__ void __launch_bounds__(256) skinny_linear_kernel(int R, int Kin, int D, const float* __restrict__ in,
                                                          const float* __restrict__ W, const float* __restrict__ b,
                                                          const float* __restrict__ extra, float* __restrict__ out,
                                                          __nv_bfloat16* __restrict__ silu_out) {
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (n >= D) return;
  const float* w = W + static_cast<size_t>(n) * Kin;
  for (int r0 = blockIdx.y * 8; r0 < R; r0 += gridDim.y * 8) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int k = lane; k < Kin; k += 32) {
      const float wk = __ldg(w + k);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int r = min(r0 + j, R - 1);
        acc[j] = fmaf(in[static_cast<size_t>(r) * Kin + k], wk, acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = warp_sum(acc[j]);
    if (lane < 8 && r0 + lane < R) {
      float v = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j == lane) v = acc[j];
      const int r = r0 + lane;
      v += b[n];
      if (extra) v += extra[static_cast<size_t>(r) * D + n];
      if (ACT == 1) v = silu_f(v);
      out[static_cast<size_t>(r) * D + n] = v;
      if (silu_out) silu_out[static_cast<size_t>(r) * D + n] = __float2bfloat16_rn(silu_f(v));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Reverse-SDE predictor update, fused with score = -params / sqrt(var(t)).
// Thread/element mapping and Philox usage replicate torch's CUDA `normal_` kernel for float tensors
// (grid-stride over blocks of 256 threads, 4 normals per curand_normal4 call, element li handled by
// thread li % T in iteration li / (4T), component (li / T) % 4), so with z == NULL the noise equals
// what `torch.randn_like(x)` would have produced from generator state (seed, offset).
// Arithmetic uses explicitly rounded fp32 ops in the reference's operation order
// (diffusion_continuous.py:141-191) so the update is bit-identical to the PyTorch op sequence given
// identical inputs.
// coef = coef_table + step * 8:
//   [0] sqrt(var(t))
//   ancestral:          [1] beta            [2] sqrt(1-beta)      [3] sqrt(beta)
//   reverse diffusion:  [1] f coeff (-0.5 g2) [2] g2 * (0.5 if pf else 1) [3] dt   [4] g (0 if pf)  [5] sqrt(dt)
//   euler-maruyama:     [1] f coeff          [2] g2 * (0.5 if pf else 1) [3] dt(<0) [4] sqrt(g2)*sqrt(-dt) (0 if pf)
//   ddim:               [1] sqrt(at_next)    [2] sqrt(1-at)        [3] sqrt(at)     [4] sqrt(1-at_next)
//   corrector:          [1] step_size        [2] sqrt(2 * step_size)
// ------------------------------------------------------------------------------------------------
template <int PRED>
__device__ __forceinline__ void sde_update(float x, float prm, float z, const float* __restrict__ c, float& xn,
                                           float& xm) {
  if (PRED == LDT_PRED_ANCESTRAL) {
    const float score = __fdiv_rn(-prm, c[0]);
    xm = __fdiv_rn(__fadd_rn(x, __fmul_rn(c[1], score)), c[2]);
    xn = __fadd_rn(xm, __fmul_rn(c[3], z));
  } else if (PRED == LDT_PRED_REVERSE_DIFFUSION) {
    const float score = __fdiv_rn(-prm, c[0]);
    const float f = __fmul_rn(c[1], x);
    const float dx = __fmul_rn(__fsub_rn(f, __fmul_rn(c[2], score)), c[3]);
    xm = __fsub_rn(x, dx);
    xn = __fadd_rn(xm, __fmul_rn(__fmul_rn(c[4], z), c[5]));   // (g * z) * sqrt(dt), in that order (:149)
  } else if (PRED == LDT_PRED_EULER_MARUYAMA) {
    const float score = __fdiv_rn(-prm, c[0]);
    const float f = __fsub_rn(__fmul_rn(c[1], x), __fmul_rn(c[2], score));
    xm = __fadd_rn(x, __fmul_rn(f, c[3]));
    xn = __fadd_rn(xm, __fmul_rn(c[4], z));
  } else if (PRED == LDT_PRED_DDIM) {  // sigma = 0
    const float a = __fdiv_rn(__fmul_rn(c[1], __fsub_rn(x, __fmul_rn(c[2], prm))), c[3]);
    xm = __fadd_rn(a, __fmul_rn(c[4], prm));
    xn = xm;
  } else {  // corrector step (Langevin / ancestral corrector :193-229): x_mean = x + step*grad, x = x_mean + sqrt(2 step)*z
    const float score = __fdiv_rn(-prm, c[0]);
    xm = __fadd_rn(x, __fmul_rn(c[1], score));
    xn = __fadd_rn(xm, __fmul_rn(c[2], z));
  }
}

// x_next may alias x (the fused loop updates its state in place): every element is read before it is written by the
// same thread and by no other, so neither pointer is declared __restrict__.
// rng_state (device, or NULL): {seed, base offset} of the Philox stream, added to the by-value seed / offset.  A replayed
// CUDA graph reads them at run time, so one capture serves every generator position.
template <int PRED>
__global__ void __launch_bounds__(256) sde_step_kernel(long long numel, const float* x,
                                                     const float* __restrict__ params, const float* __restrict__ z,
                                                     const float* __restrict__ coef_table,
                                                     const int* __restrict__ step_index, unsigned long long seed,
                                                     unsigned long long offset, unsigned long long offset_per_step,
                                                     const unsigned long long* __restrict__ rng_state,
                                                     float* x_next, float* __restrict__ x_mean) {
  pdl_launch_dependents();
  pdl_wait();
  if (rng_state != nullptr) {
    seed += rng_state[0];
    offset += rng_state[1];
  }
  const int step = step_index ? *step_index : 0;
  const float* c = coef_table + static_cast<size_t>(step) * LDT_SDE_COEF_STRIDE;
  const long long T = static_cast<long long>(gridDim.x) * 256;
  const long long tid = blockIdx.x * 256LL + threadIdx.x;
  curandStatePhilox4_32_10_t st;
  if (z == nullptr) curand_init(seed, static_cast<unsigned long long>(tid), offset + offset_per_step * step, &st);
  const long long rounded = ((numel - 1) / (T * 4) + 1) * T * 4;
  for (long long li0 = tid; li0 < rounded; li0 += T * 4) {
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (z == nullptr) r = curand_normal4(&st);
    const float rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const long long li = li0 + T * ii;
      if (li < numel) {
        const float zz = z ? z[li] : rr[ii];
        float xn, xm;
        sde_update<PRED>(x[li], params[li], zz, c, xn, xm);
        x_next[li] = xn;
        if (x_mean) x_mean[li] = xm;
      }
    }
  }
}

__global__ void advance_step_kernel(int* step_index) {
  pdl_launch_dependents();
  pdl_wait();
  *step_index += 1;
}

__global__ void __launch_bounds__(256) select_row_kernel(const float* __restrict__ table, long long row_len,
                                                       const int* __restrict__ step_index, float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const float4* src = reinterpret_cast<const float4*>(table + static_cast<long long>(*step_index) * row_len);
  float4* dst = reinterpret_cast<float4*>(out);
  const long long n4 = row_len >> 2;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += gridDim.x * 256LL) dst[i] = src[i];
}

// c[r, :] = table[step, :] (+ extra[r, :]);  silu_out[r, :] = bf16(SiLU(c[r, :])).
// The per-step conditioning vector of conditional sampling: table holds TimeEmbedding(t_i) for every step i
// (batch-invariant), extra the per-sample image / label embedding (score.py:135: c = t_emb + condition[1]).
// Same operation order as skinny_linear_kernel's `(acc + b) + extra`, so c is bit-identical to the eager forward.
__global__ void __launch_bounds__(256) cond_silu_kernel(int R, int D, const float* __restrict__ table,
                                                      const int* __restrict__ step_index, const float* __restrict__ extra,
                                                      float* __restrict__ c_out, __nv_bfloat16* __restrict__ silu_out) {
  pdl_launch_dependents();
  pdl_wait();
  const float* row = table + static_cast<size_t>(step_index ? *step_index : 0) * D;
  const long long total = static_cast<long long>(R) * D;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const int n = static_cast<int>(i % D);
    float v = row[n];
    if (extra) v += extra[i];
    if (c_out) c_out[i] = v;
    silu_out[i] = __float2bfloat16_rn(silu_f(v));
  }
}

// ---- PNDM (diffusion_continuous.py:260-316) ------------------------------------------------------------------
// transfer(): x_next = x + d * (a * x - b * et) with the three step scalars read from device memory (they are produced
// by the reference's own torch expression on 1-element tensors, so they are bit-identical); explicit roundings in
// the reference's operation order.
__global__ void __launch_bounds__(256) pndm_transfer_kernel(long long numel, const float* __restrict__ x,
                                                          const float* __restrict__ et, const float* __restrict__ coef,
                                                          float* __restrict__ out) {
  const float d = coef[0], a = coef[1], b = coef[2];
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < numel; i += gridDim.x * 256LL)
    out[i] = __fadd_rn(x[i], __fmul_rn(d, __fsub_rn(__fmul_rn(a, x[i]), __fmul_rn(b, et[i]))));
}

// out = scale * (((c0*a0 + c1*a1) + c2*a2) + c3*a3): the Runge-Kutta / Adams-Bashforth noise combinations (:284,300).
__global__ void __launch_bounds__(256) lincomb4_kernel(long long numel, float c0, const float* __restrict__ a0, float c1,
                                                     const float* __restrict__ a1, float c2, const float* __restrict__ a2,
                                                     float c3, const float* __restrict__ a3, float scale,
                                                     float* __restrict__ out) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < numel; i += gridDim.x * 256LL) {
    float v = __fadd_rn(__fmul_rn(c0, a0[i]), __fmul_rn(c1, a1[i]));
    v = __fadd_rn(v, __fmul_rn(c2, a2[i]));
    v = __fadd_rn(v, __fmul_rn(c3, a3[i]));
    out[i] = __fmul_rn(scale, v);
  }
}

// norms[b] = ||x[b, :]||_2 ; then out[0] = mean_b norms[b] (fixed summation order: deterministic).  The two global
// norms of the Langevin corrector (:205-206).
__global__ void __launch_bounds__(256) row_norm_kernel(long long L, const float* __restrict__ x, float* __restrict__ norms) {
  __shared__ float red[8];
  const float* row = x + static_cast<size_t>(blockIdx.x) * L;
  float acc = 0.f;
  for (long long i = threadIdx.x; i < L; i += 256) acc = fmaf(row[i], row[i], acc);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    norms[blockIdx.x] = sqrtf(t);
  }
}
__global__ void mean_kernel(int n, const float* __restrict__ v, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < n; ++i) t += v[i];
    out[0] = t / static_cast<float>(n);
  }
}

}  // namespace ldt

using namespace ldt;

static int cast_pad_impl(const char* who, long long rows, int cols, const float* in, int ld_in, void* out, int ld_out,
                         void* stream) {
  LDT_REQUIRE(rows >= 0 && cols > 0 && ld_in >= cols && ld_out >= cols && ld_out % 8 == 0, LDT_ERR_INVALID,
              "%s: bad shape rows=%lld cols=%d ld_in=%d ld_out=%d (ld_out must be a multiple of 8)", who, rows, cols, ld_in,
              ld_out);
  if (rows == 0) return LDT_OK;
  LDT_REQUIRE(in && out && reinterpret_cast<uintptr_t>(out) % 16 == 0 && reinterpret_cast<uintptr_t>(in) % 16 == 0,
              LDT_ERR_INVALID, "%s: null or misaligned pointer", who);
  const long long total = rows * (ld_out / 8);
  const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, static_cast<long long>(num_sms()) * 16));
  LDT_CUDA_OK(launch_pdl(cast_pad_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), rows, cols, in, ld_in,
                         static_cast<__nv_bfloat16*>(out), ld_out));
  return LDT_OK;
}

extern "C" int ldt_cast_pad_bf16(int rows, int cols, const float* in, int ld_in, void* out, int ld_out, void* stream) {
  return cast_pad_impl("ldt_cast_pad_bf16", rows, cols, in, ld_in, out, ld_out, stream);
}
extern "C" int ldt_pack_weights(int rows, int cols, const float* w, int ld_in, void* out, int ld_out, void* stream) {
  return cast_pad_impl("ldt_pack_weights", rows, cols, w, ld_in, out, ld_out, stream);
}

extern "C" int ldt_layernorm_mod_bf16(int rows, int C, const float* x, const float* shift, const float* scale,
                                      long long mod_stride, int rows_per_mod, const float* weight, const float* bias,
                                      float eps, void* y, void* stream) {
  LDT_REQUIRE(rows >= 0 && C > 0 && C % 128 == 0 && C <= 128 * LN_MAXV, LDT_ERR_INVALID,
              "ldt_layernorm_mod_bf16: C=%d must be a multiple of 128 and <= %d", C, 128 * LN_MAXV);
  if (rows == 0) return LDT_OK;
  LDT_REQUIRE(x && y, LDT_ERR_INVALID, "ldt_layernorm_mod_bf16: null pointer");
  LDT_REQUIRE((shift == nullptr) == (scale == nullptr), LDT_ERR_INVALID, "ldt_layernorm_mod_bf16: shift and scale go together");
  LDT_REQUIRE(!(scale && weight), LDT_ERR_INVALID, "ldt_layernorm_mod_bf16: pass AdaLN (shift,scale) or affine (weight,bias), not both");
  LDT_REQUIRE(mod_stride % 4 == 0, LDT_ERR_INVALID, "ldt_layernorm_mod_bf16: mod_stride must be a multiple of 4");
  if (rows_per_mod <= 0) rows_per_mod = 1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  __nv_bfloat16* yb = static_cast<__nv_bfloat16*>(y);
  const int grid = (rows + 3) / 4;
#define LDT_LN_CASE(NV)                                                                                              \
  case NV:                                                                                                           \
    LDT_CUDA_OK(launch_pdl(layernorm_mod_kernel<NV>, dim3(grid), dim3(128), 0, s, rows, x, shift, scale, mod_stride,   \
                           rows_per_mod, weight, bias, eps, yb));                                                     \
    break
  switch (C / 128) {
    LDT_LN_CASE(1); LDT_LN_CASE(2); LDT_LN_CASE(3); LDT_LN_CASE(4); LDT_LN_CASE(5); LDT_LN_CASE(6); LDT_LN_CASE(7);
    LDT_LN_CASE(8); LDT_LN_CASE(9); LDT_LN_CASE(10); LDT_LN_CASE(11); LDT_LN_CASE(12); LDT_LN_CASE(13);
    LDT_LN_CASE(14); LDT_LN_CASE(15); LDT_LN_CASE(16);
    default: set_last_error("ldt_layernorm_mod_bf16: unsupported C=%d", C); return LDT_ERR_UNSUPPORTED;
  }
#undef LDT_LN_CASE
  LDT_CUDA_OK(cudaGetLastError());
  return LDT_OK;
}

extern "C" int ldt_time_embedding(int R, int half, int D, const float* t, const float* freq, const float* w0,
                                  const float* b0, const float* w1, const float* b1, const float* extra, float* c,
                                  void* silu_c, float* scratch, void* stream) {
  LDT_REQUIRE(R > 0 && half > 0 && D > 0, LDT_ERR_INVALID, "ldt_time_embedding: bad shape R=%d half=%d D=%d", R, half, D);
  LDT_REQUIRE(t && freq && w0 && b0 && w1 && b1 && c && scratch, LDT_ERR_INVALID, "ldt_time_embedding: null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* emb = scratch;                                   // [R, 2*half]
  float* h1 = scratch + static_cast<size_t>(R) * 2 * half;  // [R, D]
  sincos_kernel<<<(R * half + 255) / 256, 256, 0, s>>>(R, half, t, freq, emb);
  const int gy = min((R + 7) / 8, 64);
  skinny_linear_kernel<1><<<dim3((D + 7) / 8, gy), 256, 0, s>>>(R, 2 * half, D, emb, w0, b0, nullptr, h1, nullptr);
  skinny_linear_kernel<0><<<dim3((D + 7) / 8, gy), 256, 0, s>>>(R, D, D, h1, w1, b1, extra, c,
                                                              static_cast<__nv_bfloat16*>(silu_c));
  LDT_CUDA_OK(cudaGetLastError());
  return LDT_OK;
}

extern "C" int ldt_sde_step(int predictor, long long numel, const float* x, const float* params, const float* z,
                            const float* coef_table, const int* step_index, unsigned long long seed,
                            unsigned long long offset, unsigned long long offset_per_step,
                            const unsigned long long* rng_state, int rng_grid, float* x_next, float* x_mean, void* stream) {
  LDT_REQUIRE(numel >= 0, LDT_ERR_INVALID, "ldt_sde_step: negative numel");
  if (numel == 0) return LDT_OK;
  LDT_REQUIRE(x && params && coef_table && x_next, LDT_ERR_INVALID, "ldt_sde_step: null pointer");
  int grid = rng_grid;
  if (grid <= 0) grid = static_cast<int>(std::min<long long>((numel + 255) / 256, static_cast<long long>(num_sms()) * 8));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define LDT_SDE_LAUNCH(P)                                                                                          \
  LDT_CUDA_OK(launch_pdl(sde_step_kernel<P>, dim3(grid), dim3(256), 0, s, numel, x, params, z, coef_table, step_index, \
                         seed, offset, offset_per_step, rng_state, x_next, x_mean))
  switch (predictor) {
    case LDT_PRED_ANCESTRAL: LDT_SDE_LAUNCH(LDT_PRED_ANCESTRAL); break;
    case LDT_PRED_REVERSE_DIFFUSION: LDT_SDE_LAUNCH(LDT_PRED_REVERSE_DIFFUSION); break;
    case LDT_PRED_EULER_MARUYAMA: LDT_SDE_LAUNCH(LDT_PRED_EULER_MARUYAMA); break;
    case LDT_PRED_DDIM: LDT_SDE_LAUNCH(LDT_PRED_DDIM); break;
    case LDT_PRED_CORRECTOR: LDT_SDE_LAUNCH(LDT_PRED_CORRECTOR); break;
    default: set_last_error("ldt_sde_step: unknown predictor %d", predictor); return LDT_ERR_INVALID;
  }
#undef LDT_SDE_LAUNCH
  LDT_CUDA_OK(cudaGetLastError());
  return LDT_OK;
}

extern "C" int ldt_advance_step(int* step_index, void* stream) {
  LDT_REQUIRE(step_index, LDT_ERR_INVALID, "ldt_advance_step: null pointer");
  LDT_CUDA_OK(launch_pdl(advance_step_kernel, dim3(1), dim3(1), 0, static_cast<cudaStream_t>(stream), step_index));
  return LDT_OK;
}

extern "C" int ldt_select_row(const float* table, long long row_len, const int* step_index, float* out, void* stream) {
  LDT_REQUIRE(table && step_index && out && row_len > 0 && row_len % 4 == 0, LDT_ERR_INVALID,
              "ldt_select_row: bad arguments (row_len=%lld must be a multiple of 4)", row_len);
  const int grid = static_cast<int>(std::min<long long>((row_len / 4 + 255) / 256, static_cast<long long>(num_sms()) * 4));
  LDT_CUDA_OK(launch_pdl(select_row_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), table, row_len,
                         step_index, out));
  return LDT_OK;
}

extern "C" int ldt_cond_silu(int R, int D, const float* table, const int* step_index, const float* extra, float* c_out,
                             void* silu_out, void* stream) {
  LDT_REQUIRE(R > 0 && D > 0 && table && silu_out, LDT_ERR_INVALID, "ldt_cond_silu: bad arguments R=%d D=%d", R, D);
  const long long total = static_cast<long long>(R) * D;
  const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, static_cast<long long>(num_sms()) * 8));
  LDT_CUDA_OK(launch_pdl(cond_silu_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), R, D, table,
                         step_index, extra, c_out, static_cast<__nv_bfloat16*>(silu_out)));
  return LDT_OK;
}

extern "C" int ldt_pndm_transfer(long long numel, const float* x, const float* et, const float* coef, float* out, void* stream) {
  LDT_REQUIRE(numel >= 0 && (numel == 0 || (x && et && coef && out)), LDT_ERR_INVALID, "ldt_pndm_transfer: bad arguments");
  if (numel == 0) return LDT_OK;
  const int grid = static_cast<int>(std::min<long long>((numel + 255) / 256, static_cast<long long>(num_sms()) * 8));
  pndm_transfer_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(numel, x, et, coef, out);
  LDT_CUDA_OK(cudaGetLastError());
  return LDT_OK;
}

extern "C" int ldt_lincomb4(long long numel, float c0, const float* a0, float c1, const float* a1, float c2, const float* a2,
                            float c3, const float* a3, float scale, float* out, void* stream) {
  LDT_REQUIRE(numel >= 0 && (numel == 0 || (a0 && a1 && a2 && a3 && out)), LDT_ERR_INVALID, "ldt_lincomb4: bad arguments");
  if (numel == 0) return LDT_OK;
  const int grid = static_cast<int>(std::min<long long>((numel + 255) / 256, static_cast<long long>(num_sms()) * 8));
  lincomb4_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(numel, c0, a0, c1, a1, c2, a2, c3, a3, scale, out);
  LDT_CUDA_OK(cudaGetLastError());
  return LDT_OK;
}

extern "C" int ldt_batch_mean_norm(int rows, long long row_len, const float* x, float* norms, float* out, void* stream) {
  LDT_REQUIRE(rows > 0 && row_len > 0 && x && norms && out, LDT_ERR_INVALID, "ldt_batch_mean_norm: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  row_norm_kernel<<<rows, 256, 0, s>>>(row_len, x, norms);
  mean_kernel<<<1, 32, 0, s>>>(rows, norms, out);
  LDT_CUDA_OK(cudaGetLastError());
  return LDT_OK;
}
