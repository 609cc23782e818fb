/* CPU oracle for the point-set prologue -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * oracle_fps: furthest point sampling as the reference consumes it through pointnet2_utils.furthest_point_sample
 * (model/Compressor/layers.py:106; completion_trainer/Latent_SDE_Trainer.py:182-183).  The dependency itself
 * (pointnet2_ops, README.md:22-24, no version pin) is NOT under /root/reference, so this restates its published
 * algorithm (erikwijmans/Pointnet2_PyTorch sampling_gpu.cu: start at index 0, temp[k] = min(temp[k], |p_k - p_old|^2),
 * skip points with |p|^2 <= 1e-3, pick the arg-max) -- the skip rule is PARITY UNPINNED against that library; with
 * min_sq_norm < 0 it is the algorithm of the reference's in-tree model/functional/src/sampling/sampling.cu:86-167, and that
 * form IS pinned: the in-tree kernel is compiled into oracle/_ref/libref_fps.so (oracle/Makefile) and the product kernel
 * reproduces its index sequences bit for bit on the GPU (tests/test_gpu_kernels.py).  Exact ties (duplicate
 * points) resolve to the lowest index here; the CUDA originals resolve them by thread layout.
 * Distances use fmaf(dz,dz,fmaf(dy,dy,dx*dx)), nvcc's default contraction of the reference expression.
 *
 * oracle_knn: the k nearest points of each centre ordered by (squared distance, index); knn_point
 * (model/Compressor/layers.py:86-98) returns the same set through square_distance + topk(sorted=False).
 * Compile with -ffp-contract=off.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>

void oracle_fps(int b, int n, int m, const float* xyz, float min_sq_norm, int* idx) {
  float* temp = (float*)malloc(sizeof(float) * (size_t)n);
  for (int c = 0; c < b; ++c) {
    const float* p = xyz + (size_t)c * n * 3;
    int* out = idx + (size_t)c * m;
    for (int k = 0; k < n; ++k) temp[k] = 1e10f;
    int old = 0;
    out[0] = 0;
    for (int j = 1; j < m; ++j) {
      const float x1 = p[3 * old], y1 = p[3 * old + 1], z1 = p[3 * old + 2];
      float best = -1.f;
      int besti = 0;
      for (int k = 0; k < n; ++k) {
        const float x2 = p[3 * k], y2 = p[3 * k + 1], z2 = p[3 * k + 2];
        const float mag = fmaf(z2, z2, fmaf(y2, y2, x2 * x2));
        if (!(mag > min_sq_norm)) continue;
        const float dx = x2 - x1, dy = y2 - y1, dz = z2 - z1;
        const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        const float d2 = d < temp[k] ? d : temp[k];
        temp[k] = d2;
        if (d2 > best) { best = d2; besti = k; }
      }
      old = besti;
      out[j] = old;
    }
  }
  free(temp);
}

void oracle_knn(int b, int n, int s, int k, const float* xyz, const float* centers, int* idx) {
  float* d = (float*)malloc(sizeof(float) * (size_t)n);
  for (int c = 0; c < b; ++c) {
    const float* p = xyz + (size_t)c * n * 3;
    for (int g = 0; g < s; ++g) {
      const float* q = centers + ((size_t)c * s + g) * 3;
      int* out = idx + ((size_t)c * s + g) * k;
      for (int i = 0; i < n; ++i) {
        const float dx = p[3 * i] - q[0], dy = p[3 * i + 1] - q[1], dz = p[3 * i + 2] - q[2];
        d[i] = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
      }
      for (int r = 0; r < k; ++r) {
        float bv = FLT_MAX;
        int bi = 0x7fffffff;
        for (int i = 0; i < n; ++i)
          if (d[i] < bv) { bv = d[i]; bi = i; }
        out[r] = bi;
        d[bi] = FLT_MAX;
      }
    }
  }
  free(d);
}
