// Point-set prologue kernels of the completion path: furthest point sampling and k-nearest-neighbour grouping.
//
// They replace `pointnet2_utils.furthest_point_sample` (an un-vendored dependency of the reference: README.md:22-24,
// called from model/Compressor/layers.py:106 and completion_trainer/Latent_SDE_Trainer.py:182-183) and the
// square_distance + torch.topk pair of `knn_point` (model/Compressor/layers.py:63-98), which materialises a dense
// [B, S, N] distance matrix per call.  Both run once per sample() call, before the reverse-SDE loop.
//
// furthest point sampling: one CTA per cloud, the cloud's points and their running distance-to-set live in REGISTERS
// (PPT points per thread), one block-wide arg-max per selected point (warp shuffles + one shared-memory exchange).
// Semantics follow the pointnet2_ops kernel the reference links against, as far as it is known (the dependency is
// absent, so parity is unpinned): first index is 0, a point's distance-to-set is min-updated with
// fma(dz,dz,fma(dy,dy,dx*dx)), points whose squared norm is <= min_sq_norm are never selected (pointnet2_ops skips
// |p|^2 <= 1e-3; pass a negative value to disable, which is the behaviour of the reference's own in-tree
// model/functional/src/sampling/sampling.cu:86-167), and ties resolve to the LOWEST index.
#include <cfloat>

#include "common.cuh"
#include "ldt_b200.h"

namespace ldt {

constexpr int FPS_THREADS = 512;

struct ArgMax {
  float v;
  int i;
};
__device__ __forceinline__ ArgMax argmax_pick(ArgMax a, ArgMax b) {   // larger value; ties -> lower index
  return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}

template <int PPT>
__global__ void __launch_bounds__(FPS_THREADS) fps_kernel(int n, int m, const float* __restrict__ xyz,
                                                        float min_sq_norm, int* __restrict__ idx_out) {
  __shared__ float s_val[FPS_THREADS / 32];
  __shared__ int s_idx[FPS_THREADS / 32];
  __shared__ float s_sel[2][4];   // coordinates of the selected point, double-buffered across iterations
  const float* pts = xyz + static_cast<size_t>(blockIdx.x) * n * 3;
  int* out = idx_out + static_cast<size_t>(blockIdx.x) * m;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  float px[PPT], py[PPT], pz[PPT], dist[PPT];
  bool live[PPT];
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    const int k = tid + j * FPS_THREADS;
    if (k < n) {
      px[j] = pts[3 * k]; py[j] = pts[3 * k + 1]; pz[j] = pts[3 * k + 2];
      const float mag = fmaf(pz[j], pz[j], fmaf(py[j], py[j], px[j] * px[j]));
      live[j] = mag > min_sq_norm;
    } else {
      px[j] = py[j] = pz[j] = 0.f;
      live[j] = false;
    }
    dist[j] = 1e10f;
  }
  if (tid == 0) {
    out[0] = 0;
    s_sel[0][0] = pts[0]; s_sel[0][1] = pts[1]; s_sel[0][2] = pts[2];
  }
  __syncthreads();
  for (int it = 1; it < m; ++it) {
    const int cur = (it - 1) & 1;
    const float x1 = s_sel[cur][0], y1 = s_sel[cur][1], z1 = s_sel[cur][2];
    ArgMax best{-1.f, 0};
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      if (live[j]) {
        const float dx = px[j] - x1, dy = py[j] - y1, dz = pz[j] - z1;
        const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        const float d2 = fminf(d, dist[j]);
        dist[j] = d2;
        if (d2 > best.v) { best.v = d2; best.i = tid + j * FPS_THREADS; }   // ascending index: strict > keeps the lowest
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ArgMax other{__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.i, o)};
      best = argmax_pick(best, other);
    }
    if (lane == 0) { s_val[warp] = best.v; s_idx[warp] = best.i; }
    __syncthreads();
    if (warp == 0) {
      ArgMax b = (lane < FPS_THREADS / 32) ? ArgMax{s_val[lane], s_idx[lane]} : ArgMax{-2.f, 0x7fffffff};
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        ArgMax other{__shfl_xor_sync(0xffffffffu, b.v, o), __shfl_xor_sync(0xffffffffu, b.i, o)};
        b = argmax_pick(b, other);
      }
      if (lane == 0) {
        out[it] = b.i;
        s_sel[it & 1][0] = pts[3 * b.i]; s_sel[it & 1][1] = pts[3 * b.i + 1]; s_sel[it & 1][2] = pts[3 * b.i + 2];
      }
    }
    __syncthreads();
  }
}

// k nearest points of every centre: one CTA per (cloud, centre); squared distances of the n points to the centre are
// staged in shared memory, then k rounds of block-wide arg-min (ties -> lowest index) pick the neighbours in order of
// increasing distance.  The reference's knn_point returns them in unspecified order (topk sorted=False); every
// consumer (group mean, max-pool over k) is order-invariant.
constexpr int KNN_THREADS = 128;

__global__ void __launch_bounds__(KNN_THREADS) knn_kernel(int n, int s, int k, const float* __restrict__ xyz,
                                                        const float* __restrict__ centers, int* __restrict__ idx_out) {
  extern __shared__ float s_d[];   // n distances
  __shared__ float r_val[KNN_THREADS / 32];
  __shared__ int r_idx[KNN_THREADS / 32];
  __shared__ int s_pick;
  const int cloud = blockIdx.x / s;
  const float* pts = xyz + static_cast<size_t>(cloud) * n * 3;
  const float* c = centers + static_cast<size_t>(blockIdx.x) * 3;
  int* out = idx_out + static_cast<size_t>(blockIdx.x) * k;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float cx = c[0], cy = c[1], cz = c[2];
  for (int i = tid; i < n; i += KNN_THREADS) {
    const float dx = pts[3 * i] - cx, dy = pts[3 * i + 1] - cy, dz = pts[3 * i + 2] - cz;
    s_d[i] = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
  }
  __syncthreads();
  for (int r = 0; r < k; ++r) {
    float bv = FLT_MAX;
    int bi = 0x7fffffff;
    for (int i = tid; i < n; i += KNN_THREADS) {
      const float d = s_d[i];
      if (d < bv) { bv = d; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { r_val[warp] = bv; r_idx[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < KNN_THREADS / 32; ++w)
        if (r_val[w] < bv || (r_val[w] == bv && r_idx[w] < bi)) { bv = r_val[w]; bi = r_idx[w]; }
      out[r] = bi;
      s_pick = bi;
    }
    __syncthreads();
    if (tid == 0) s_d[s_pick] = FLT_MAX;   // taken; FLT_MAX entries lose every later comparison against live points
    __syncthreads();
  }
}

}  // namespace ldt

using namespace ldt;


// ------------------------------------------------------------------------------------------------
// LocalGrouper's normalised group features (reference model/Compressor/layers.py:300-317), written directly as the A
// operand [b*s*k, ld_out] of PreExtraction's first 1x1 convolution: for group (b, s) and neighbour j
//     g      = cat(fea[b, idx[b,s,j], :], xyz[b, idx[b,s,j], :])                         d + 3 channels
//     mean   = cat(fea[b, ci, :], xyz[b, ci, :])  ("anchor", ci = center_idx[b,s])   or   mean_j g  ("center")
//     row    = cat(alpha * ((g - mean) / (std_b + 1e-5)) + beta,  fea[b, ci, :]),  zero-padded
// with std_b the unbiased standard deviation of ALL (g - mean) of sample b (torch.std over reshape(B, -1)).  Two kernels:
// per-group partial sums in double (fixed order, no atomics), then every block re-reduces its sample's s partials.
// ------------------------------------------------------------------------------------------------
constexpr int GRP_THREADS = 128;

__device__ __forceinline__ float group_value(const float* __restrict__ xyz, const float* __restrict__ fea, int d, size_t point, int c) {
  return c < d ? __ldg(fea + point * d + c) : __ldg(xyz + point * 3 + (c - d));
}

template <bool ASSEMBLE>
__global__ void __launch_bounds__(GRP_THREADS)
group_kernel(int n, int s_, int k, int d, const float* __restrict__ xyz, const float* __restrict__ fea,
             const int* __restrict__ center_idx, const int* __restrict__ group_idx, int normalize,
             const float* __restrict__ alpha, const float* __restrict__ beta, double* __restrict__ partial,
             float* __restrict__ out, int ld_out) {
  extern __shared__ int s_idx[];   // the group's k neighbour indices
  __shared__ double red[2][GRP_THREADS];
  __shared__ float s_denom;
  const int s = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const size_t base = static_cast<size_t>(b) * n;
  for (int j = tid; j < k; j += GRP_THREADS) s_idx[j] = group_idx[(static_cast<size_t>(b) * s_ + s) * k + j];
  const size_t ci = base + center_idx[static_cast<size_t>(b) * s_ + s];
  if (ASSEMBLE && tid == 0) {
    float denom = 1.f;
    if (normalize != 0) {
      double S = 0.0, Q = 0.0;
      for (int i = 0; i < s_; ++i) {
        S += partial[(static_cast<size_t>(b) * s_ + i) * 2];
        Q += partial[(static_cast<size_t>(b) * s_ + i) * 2 + 1];
      }
      const double cnt = static_cast<double>(s_) * k * (d + 3);
      const double var = (Q - S * S / cnt) / (cnt - 1.0);
      denom = __fadd_rn(static_cast<float>(sqrt(var > 0.0 ? var : 0.0)), 1e-5f);
    }
    s_denom = denom;
  }
  __syncthreads();
  const int dc = d + 3;
  double sum = 0.0, sq = 0.0;
  for (int c = tid; c < dc; c += GRP_THREADS) {
    float mean = 0.f;
    if (normalize == 2) {
      mean = group_value(xyz, fea, d, ci, c);
    } else if (normalize == 1) {
      float acc = 0.f;
      for (int j = 0; j < k; ++j) acc += group_value(xyz, fea, d, base + s_idx[j], c);
      mean = acc / static_cast<float>(k);
    }
    if constexpr (ASSEMBLE) {
      const float denom = s_denom;
      const float al = normalize != 0 ? __ldg(alpha + c) : 1.f, be = normalize != 0 ? __ldg(beta + c) : 0.f;
      float* o = out + (static_cast<size_t>(b) * s_ + s) * k * ld_out + c;
      for (int j = 0; j < k; ++j) {
        float v = group_value(xyz, fea, d, base + s_idx[j], c);
        if (normalize != 0) v = __fadd_rn(__fmul_rn(al, __fdiv_rn(__fsub_rn(v, mean), denom)), be);
        o[static_cast<size_t>(j) * ld_out] = v;
      }
    } else {
      for (int j = 0; j < k; ++j) {
        const double v = static_cast<double>(__fsub_rn(group_value(xyz, fea, d, base + s_idx[j], c), mean));
        sum += v;
        sq += v * v;
      }
    }
  }
  if constexpr (ASSEMBLE) {
    // the anchor's own feature repeated for every neighbour, then the zero padding
    for (int c = dc + tid; c < ld_out; c += GRP_THREADS) {
      const float v = c < dc + d ? __ldg(fea + ci * d + (c - dc)) : 0.f;
      float* o = out + (static_cast<size_t>(b) * s_ + s) * k * ld_out + c;
      for (int j = 0; j < k; ++j) o[static_cast<size_t>(j) * ld_out] = v;
    }
  } else {
    red[0][tid] = sum;
    red[1][tid] = sq;
    __syncthreads();
    for (int w = GRP_THREADS / 2; w > 0; w >>= 1) {
      if (tid < w) {
        red[0][tid] += red[0][tid + w];
        red[1][tid] += red[1][tid + w];
      }
      __syncthreads();
    }
    if (tid == 0) {
      partial[(static_cast<size_t>(b) * s_ + s) * 2] = red[0][0];
      partial[(static_cast<size_t>(b) * s_ + s) * 2 + 1] = red[1][0];
    }
  }
}

// out[g, c] = max_j x[g*k + j, c]: the max over a group's neighbours that ends PreExtraction (layers.py:189-190) and
// MiniPointnet's max over the centres (Network.py:97).
__global__ void group_max_kernel(int k, int c_, const float* __restrict__ x, int ldx, float* __restrict__ out, int ldo) {
  const size_t g = blockIdx.x;
  for (int c = threadIdx.x; c < c_; c += blockDim.x) {
    const float* px = x + g * k * ldx + c;
    float m = px[0];
    for (int j = 1; j < k; ++j) m = fmaxf(m, px[static_cast<size_t>(j) * ldx]);
    out[g * ldo + c] = m;
  }
}

extern "C" int ldt_furthest_point_sample(int b, int n, int m, const float* xyz, float min_sq_norm, int* idx, void* stream) {
  LDT_REQUIRE(b >= 0 && n > 0 && m > 0 && m <= n, LDT_ERR_INVALID, "ldt_furthest_point_sample: bad shape b=%d n=%d m=%d", b, n, m);
  LDT_REQUIRE(n <= 16 * FPS_THREADS, LDT_ERR_UNSUPPORTED, "ldt_furthest_point_sample: n=%d exceeds %d points per cloud", n,
              16 * FPS_THREADS);
  if (b == 0) return LDT_OK;
  LDT_REQUIRE(xyz && idx, LDT_ERR_INVALID, "ldt_furthest_point_sample: null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int ppt = (n + FPS_THREADS - 1) / FPS_THREADS;
  if (ppt <= 4) fps_kernel<4><<<b, FPS_THREADS, 0, s>>>(n, m, xyz, min_sq_norm, idx);
  else if (ppt <= 8) fps_kernel<8><<<b, FPS_THREADS, 0, s>>>(n, m, xyz, min_sq_norm, idx);
  else fps_kernel<16><<<b, FPS_THREADS, 0, s>>>(n, m, xyz, min_sq_norm, idx);
  LDT_CUDA_OK(cudaGetLastError());
  return LDT_OK;
}

extern "C" int ldt_knn_indices(int b, int n, int s, int k, const float* xyz, const float* centers, int* idx, void* stream) {
  LDT_REQUIRE(b >= 0 && n > 0 && s > 0 && k > 0 && k <= n, LDT_ERR_INVALID, "ldt_knn_indices: bad shape b=%d n=%d s=%d k=%d", b, n,
              s, k);
  LDT_REQUIRE(n <= 12288, LDT_ERR_UNSUPPORTED, "ldt_knn_indices: n=%d exceeds 12288 points per cloud", n);
  if (b == 0) return LDT_OK;
  LDT_REQUIRE(xyz && centers && idx, LDT_ERR_INVALID, "ldt_knn_indices: null pointer");
  knn_kernel<<<b * s, KNN_THREADS, n * sizeof(float), static_cast<cudaStream_t>(stream)>>>(n, s, k, xyz, centers, idx);
  LDT_CUDA_OK(cudaGetLastError());
  return LDT_OK;
}

extern "C" int ldt_group_features(int b, int n, int s, int k, int d, const float* xyz, const float* fea, const int* center_idx,
                                  const int* group_idx, int normalize, const float* alpha, const float* beta, double* partial,
                                  float* out, int ld_out, void* stream) {
  LDT_REQUIRE(b >= 0 && n > 0 && s > 0 && k > 0 && d > 0, LDT_ERR_INVALID, "ldt_group_features: bad shape b=%d n=%d s=%d k=%d d=%d", b, n,
              s, k, d);
  LDT_REQUIRE(normalize >= 0 && normalize <= 2, LDT_ERR_INVALID, "ldt_group_features: normalize=%d (0 none, 1 center, 2 anchor)", normalize);
  LDT_REQUIRE(ld_out >= 2 * d + 3, LDT_ERR_INVALID, "ldt_group_features: ld_out=%d < 2 d + 3 = %d", ld_out, 2 * d + 3);
  LDT_REQUIRE(k <= 8192 && s <= 65535 && b <= 65535, LDT_ERR_UNSUPPORTED, "ldt_group_features: k=%d s=%d b=%d out of range", k, s, b);
  if (b == 0) return LDT_OK;
  LDT_REQUIRE(xyz && fea && center_idx && group_idx && out, LDT_ERR_INVALID, "ldt_group_features: null pointer");
  LDT_REQUIRE(normalize == 0 || (alpha && beta && partial), LDT_ERR_INVALID, "ldt_group_features: normalisation needs alpha, beta and the scratch");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t smem = static_cast<size_t>(k) * sizeof(int);
  if (normalize != 0)
    group_kernel<false><<<dim3(s, b), GRP_THREADS, smem, st>>>(n, s, k, d, xyz, fea, center_idx, group_idx, normalize, alpha, beta, partial,
                                                               out, ld_out);
  group_kernel<true><<<dim3(s, b), GRP_THREADS, smem, st>>>(n, s, k, d, xyz, fea, center_idx, group_idx, normalize, alpha, beta, partial, out,
                                                            ld_out);
  LDT_CUDA_OK(cudaGetLastError());
  return LDT_OK;
}

extern "C" int ldt_group_max(int groups, int k, int c, const float* x, int ldx, float* out, int ldo, void* stream) {
  LDT_REQUIRE(groups >= 0 && k > 0 && c > 0 && ldx >= c && ldo >= c, LDT_ERR_INVALID, "ldt_group_max: bad shape groups=%d k=%d c=%d ldx=%d ldo=%d",
              groups, k, c, ldx, ldo);
  if (groups == 0) return LDT_OK;
  LDT_REQUIRE(x && out, LDT_ERR_INVALID, "ldt_group_max: null pointer");
  group_max_kernel<<<groups, c < 256 ? ((c + 31) / 32) * 32 : 256, 0, static_cast<cudaStream_t>(stream)>>>(k, c, x, ldx, out, ldo);
  LDT_CUDA_OK(cudaGetLastError());
  return LDT_OK;
}
