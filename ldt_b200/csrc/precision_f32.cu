// The non-contraction kernels of the TF32 parity mode (Score.precision = "tf32"): the same token path with fp32
// activations end to end and kind::tf32 contractions (gemm.cu, operand_type 1), i.e. the precision of the reference's own
// GPU arithmetic (cuDNN convolutions run TF32 by default, torch.backends.cudnn.allow_tf32 = True).  Every tensor that is an
// operand of a contraction is rounded to the nearest TF32 value where it is produced (cvt.rna, what cuBLAS / cuDNN do on
// their TF32 paths); everything else stays fp32.  Not a performance path: plain SIMT kernels, one warp per row / one
// lane per query.  Mirrors model/layers.py:136-137,163-164 (LayerNorm + modulate) and :183-200 (compute_attention).
#include "common.cuh"
#include "ldt_b200.h"

namespace ldt {

// out[r, 0:ld_out] = tf32_round(in[r, 0:cols]) zero-padded (and optionally SiLU'ed first: the adaLN input, layers.py:172)
template <int ACT>
__global__ void __launch_bounds__(256) round_pad_tf32_kernel(long long rows, int cols, const float* __restrict__ in, int ld_in,
                                                           float* __restrict__ out, int ld_out) {
  const long long total = rows * ld_out;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const long long r = i / ld_out;
    const int c = static_cast<int>(i - r * ld_out);
    float v = (c < cols) ? in[r * ld_in + c] : 0.f;
    if (ACT == 1) v = silu_f(v);
    out[i] = round_tf32(v);
  }
}

// Error-compensated TF32 ("3xTF32"): v = hi + lo with hi = tf32(v), lo = tf32(v - hi) (v - hi is exact in fp32), so that
//   a . w  ~=  a_hi . w_hi + a_hi . w_lo + a_lo . w_hi      (the dropped a_lo . w_lo term is ~2^-22 relative)
// is ONE kind::tf32 contraction over 3 K: activations are laid out [hi | hi | lo], weights [hi | lo | hi], each part
// ld_part columns wide (zero-padded).  fp32-grade results from the tensor cores for the small fp32 layers of the prologues.
template <int ACT>
__global__ void __launch_bounds__(256) split_tf32_kernel(long long rows, int cols, const float* __restrict__ in, int ld_in,
                                                       float* __restrict__ out, int ld_part, int weight_side) {
  const long long total = rows * ld_part;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const long long r = i / ld_part;
    const int c = static_cast<int>(i - r * ld_part);
    float v = (c < cols) ? in[r * ld_in + c] : 0.f;
    if (ACT == 1) v = silu_f(v);
    const float hi = round_tf32(v);
    const float lo = round_tf32(v - hi);
    float* o = out + r * 3 * ld_part + c;
    o[0] = hi;
    o[ld_part] = weight_side ? lo : hi;
    o[2 * ld_part] = weight_side ? hi : lo;
  }
}

// LayerNorm(eps) over C channels (two-pass, as layernorm_mod_kernel), then AdaLN modulate or affine; f32 out, TF32-rounded.
__global__ void __launch_bounds__(128) layernorm_mod_f32_kernel(int rows, int C, const float* __restrict__ x,
                                                              const float* __restrict__ shift, const float* __restrict__ scale,
                                                              long long mod_stride, int rows_per_mod,
                                                              const float* __restrict__ weight, const float* __restrict__ bias,
                                                              float eps, float* __restrict__ y, int round_out) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + static_cast<size_t>(row) * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += xr[c];
  const float mean = warp_sum(s) / static_cast<float>(C);
  float q = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float d = xr[c] - mean;
    q += d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(C) + eps);
  const bool ada = (scale != nullptr);
  const long long g = ada ? static_cast<long long>(row / rows_per_mod) * mod_stride : 0;
  float* yr = y + static_cast<size_t>(row) * C;
  for (int c = lane; c < C; c += 32) {
    float mu = 1.f, ad = 0.f;
    if (ada) {
      mu = 1.f + scale[g + c];
      ad = shift[g + c];
    } else {
      if (weight) mu = weight[c];
      if (bias) ad = bias[c];
    }
    const float o = (xr[c] - mean) * rstd * mu + ad;
    yr[c] = round_out ? round_tf32(o) : o;
  }
}

// softmax(q k^T / sqrt(dh)) v over 32 keys in fp32, one lane per query, one warp per (batch, head, 32-query tile).
// Output in the reference's layout quirk ([B,H,Nq,dh] contiguous, layers.py:197), TF32-rounded (it is fc_o's A operand).
template <int DH>
__global__ void __launch_bounds__(64) attention_nk32_f32_kernel(int units, int H, int Nq, int qtiles, const float* __restrict__ q,
                                                               int ldq, const float* __restrict__ k, const float* __restrict__ v,
                                                               int ldkv, float* __restrict__ o, float scale, int round_out) {
  __shared__ float Ks[2][32][DH + 1], Vs[2][32][DH + 1];   // two warps per CTA: 33 KB at dh 64
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int unit = blockIdx.x * 2 + warp;
  if (unit >= units) return;
  const int qt = unit % qtiles, bh = unit / qtiles;
  const int h = bh % H, b = bh / H;
  for (int i = lane; i < 32 * DH; i += 32) {
    const int key = i / DH, d = i % DH;
    Ks[warp][key][d] = k[(static_cast<size_t>(b) * 32 + key) * ldkv + h * DH + d];
    Vs[warp][key][d] = v[(static_cast<size_t>(b) * 32 + key) * ldkv + h * DH + d];
  }
  __syncwarp();
  const int qi = qt * 32 + lane;
  if (qi >= Nq) return;
  float qr[DH];
#pragma unroll
  for (int d = 0; d < DH; ++d) qr[d] = q[(static_cast<size_t>(b) * Nq + qi) * ldq + h * DH + d];
  float sc[32], mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    float a = 0.f;
#pragma unroll
    for (int d = 0; d < DH; ++d) a = fmaf(qr[d], Ks[warp][j][d], a);
    sc[j] = a * scale;
    mx = fmaxf(mx, sc[j]);
  }
  float sum = 0.f, out[DH];
#pragma unroll
  for (int d = 0; d < DH; ++d) out[d] = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float e = expf(sc[j] - mx);
    sum += e;
#pragma unroll
    for (int d = 0; d < DH; ++d) out[d] = fmaf(e, Vs[warp][j][d], out[d]);
  }
  const float inv = 1.0f / sum;
  float* op = o + ((static_cast<size_t>(b) * H + h) * Nq + qi) * DH;
#pragma unroll
  for (int d = 0; d < DH; ++d) op[d] = round_out ? round_tf32(out[d] * inv) : out[d] * inv;
}

}  // namespace ldt

using namespace ldt;

extern "C" int ldt_round_pad_tf32(long long rows, int cols, const float* in, int ld_in, float* out, int ld_out, int silu,
                                  void* stream) {
  LDT_REQUIRE(rows >= 0 && cols > 0 && ld_in >= cols && ld_out >= cols, LDT_ERR_INVALID,
              "ldt_round_pad_tf32: bad shape rows=%lld cols=%d ld_in=%d ld_out=%d", rows, cols, ld_in, ld_out);
  if (rows == 0) return LDT_OK;
  LDT_REQUIRE(in && out, LDT_ERR_INVALID, "ldt_round_pad_tf32: null pointer");
  const long long total = rows * ld_out;
  const int grid = static_cast<int>(total / 256 + 1 < static_cast<long long>(num_sms()) * 16 ? total / 256 + 1
                                                                                           : static_cast<long long>(num_sms()) * 16);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (silu) round_pad_tf32_kernel<1><<<grid, 256, 0, s>>>(rows, cols, in, ld_in, out, ld_out);
  else round_pad_tf32_kernel<0><<<grid, 256, 0, s>>>(rows, cols, in, ld_in, out, ld_out);
  LDT_CUDA_OK(cudaGetLastError());
  return LDT_OK;
}

extern "C" int ldt_split_tf32(long long rows, int cols, const float* in, int ld_in, float* out, int ld_part, int weight_side,
                              int silu, void* stream) {
  LDT_REQUIRE(rows >= 0 && cols > 0 && ld_in >= cols && ld_part >= cols, LDT_ERR_INVALID,
              "ldt_split_tf32: bad shape rows=%lld cols=%d ld_in=%d ld_part=%d", rows, cols, ld_in, ld_part);
  if (rows == 0) return LDT_OK;
  LDT_REQUIRE(in && out, LDT_ERR_INVALID, "ldt_split_tf32: null pointer");
  const long long total = rows * ld_part;
  const long long want = total / 256 + 1, cap = static_cast<long long>(num_sms()) * 16;
  const int grid = static_cast<int>(want < cap ? want : cap);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (silu) split_tf32_kernel<1><<<grid, 256, 0, st>>>(rows, cols, in, ld_in, out, ld_part, weight_side);
  else split_tf32_kernel<0><<<grid, 256, 0, st>>>(rows, cols, in, ld_in, out, ld_part, weight_side);
  LDT_CUDA_OK(cudaGetLastError());
  return LDT_OK;
}

extern "C" int ldt_layernorm_mod_f32(int rows, int C, const float* x, const float* shift, const float* scale, long long mod_stride,
                                     int rows_per_mod, const float* weight, const float* bias, float eps, float* y, int round_tf32_out,
                                     void* stream) {
  LDT_REQUIRE(rows >= 0 && C > 0, LDT_ERR_INVALID, "ldt_layernorm_mod_f32: bad shape rows=%d C=%d", rows, C);
  if (rows == 0) return LDT_OK;
  LDT_REQUIRE(x && y, LDT_ERR_INVALID, "ldt_layernorm_mod_f32: null pointer");
  LDT_REQUIRE((shift == nullptr) == (scale == nullptr), LDT_ERR_INVALID, "ldt_layernorm_mod_f32: shift and scale go together");
  LDT_REQUIRE(!(scale && weight), LDT_ERR_INVALID, "ldt_layernorm_mod_f32: pass AdaLN (shift,scale) or affine (weight,bias), not both");
  if (rows_per_mod <= 0) rows_per_mod = 1;
  layernorm_mod_f32_kernel<<<(rows + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(rows, C, x, shift, scale, mod_stride,
                                                                                         rows_per_mod, weight, bias, eps, y, round_tf32_out);
  LDT_CUDA_OK(cudaGetLastError());
  return LDT_OK;
}

extern "C" int ldt_attention_nk32_f32(int B, int H, int Nq, int dh, const float* q, int ldq, const float* k, const float* v, int ldkv,
                                      float* o, int round_tf32_out, void* stream) {
  LDT_REQUIRE(B >= 0 && H > 0 && Nq > 0, LDT_ERR_INVALID, "ldt_attention_nk32_f32: bad shape B=%d H=%d Nq=%d", B, H, Nq);
  LDT_REQUIRE(dh == 32 || dh == 64, LDT_ERR_UNSUPPORTED, "ldt_attention_nk32_f32: head dim %d not in {32, 64}", dh);
  if (B == 0) return LDT_OK;
  LDT_REQUIRE(q && k && v && o, LDT_ERR_INVALID, "ldt_attention_nk32_f32: null pointer");
  LDT_REQUIRE(ldq >= H * dh && ldkv >= H * dh, LDT_ERR_INVALID, "ldt_attention_nk32_f32: ldq=%d ldkv=%d must be >= H*dh", ldq, ldkv);
  const int qtiles = (Nq + 31) / 32;
  const long long units = static_cast<long long>(B) * H * qtiles;
  LDT_REQUIRE(units < (1LL << 31), LDT_ERR_INVALID, "ldt_attention_nk32_f32: too many work units");
  const int grid = static_cast<int>((units + 1) / 2);
  const float scale = 1.0f / sqrtf(static_cast<float>(dh));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dh == 64) attention_nk32_f32_kernel<64><<<grid, 64, 0, s>>>(static_cast<int>(units), H, Nq, qtiles, q, ldq, k, v, ldkv, o, scale, round_tf32_out);
  else attention_nk32_f32_kernel<32><<<grid, 64, 0, s>>>(static_cast<int>(units), H, Nq, qtiles, q, ldq, k, v, ldkv, o, scale, round_tf32_out);
  LDT_CUDA_OK(cudaGetLastError());
  return LDT_OK;
}
