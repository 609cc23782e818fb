/* CPU oracle for the approximate-EMD path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference's approxmatchkernel + matchcostkernel
 * (evaluation/pytorch_structural_losses/src/approxmatch.cu:3-182 and :184-224): the 9-level auction
 * (levels -4^7 .. -4^-1), three passes per level, a dense match[m][n] matrix, then cost = sum match * distance.
 * Sequential loops in the reference's per-thread accumulation order; expf() stands in for the GPU's __expf
 * (ex2.approx), so agreement with either GPU implementation is to ~1e-5 relative, not bit for bit.
 * The reference has no golden vectors for this path; the pin is the reference's own CUDA kernels compiled into
 * oracle/_ref/libref_emd.so and run on the GPU box (tests/test_gpu_kernels.py), with this file as the CPU cross-check.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static float d2f(const float* a, const float* b) {
  const float dx = b[0] - a[0], dy = b[1] - a[1], dz = b[2] - a[2];
  return dx * dx + dy * dy + dz * dz;
}

/* xyz1 [b,n,3], xyz2 [b,m,3] -> cost [b]; match_out (optional) [b,m,n] */
void oracle_match_cost(int b, int n, int m, const float* xyz1, const float* xyz2, float* cost, float* match_out) {
  float* match = (float*)malloc(sizeof(float) * (size_t)n * m);
  float* remainL = (float*)malloc(sizeof(float) * n);
  float* remainR = (float*)malloc(sizeof(float) * m);
  float* ratioL = (float*)malloc(sizeof(float) * n);
  float* ratioR = (float*)malloc(sizeof(float) * m);
  float multiL, multiR;
  if (n >= m) { multiL = 1.f; multiR = (float)(n / m); } else { multiL = (float)(m / n); multiR = 1.f; }
  for (int i = 0; i < b; ++i) {
    const float* p1 = xyz1 + (size_t)i * n * 3;
    const float* p2 = xyz2 + (size_t)i * m * 3;
    memset(match, 0, sizeof(float) * (size_t)n * m);
    for (int k = 0; k < n; ++k) remainL[k] = multiL;
    for (int l = 0; l < m; ++l) remainR[l] = multiR;
    for (int j = 7; j > -2; --j) {
      const float level = -powf(4.0f, (float)j);
      for (int k = 0; k < n; ++k) {                                   /* approxmatch.cu:28-61 */
        float suml = 1e-9f;
        for (int l = 0; l < m; ++l) suml += expf(level * d2f(p1 + k * 3, p2 + l * 3)) * remainR[l];
        ratioL[k] = remainL[k] / suml;
      }
      for (int l = 0; l < m; ++l) {                                   /* :78-115 */
        float sumr = 0.f;
        for (int k = 0; k < n; ++k) sumr += expf(level * d2f(p1 + k * 3, p2 + l * 3)) * ratioL[k];
        sumr *= remainR[l];
        const float consumption = fminf(remainR[l] / (sumr + 1e-9f), 1.0f);
        ratioR[l] = consumption * remainR[l];
        remainR[l] = fmaxf(0.0f, remainR[l] - sumr);
      }
      for (int k = 0; k < n; ++k) {                                   /* :133-165 */
        float suml = 0.f;
        for (int l = 0; l < m; ++l) {
          const float w = expf(level * d2f(p1 + k * 3, p2 + l * 3)) * ratioL[k] * ratioR[l];
          match[(size_t)l * n + k] += w;
          suml += w;
        }
        remainL[k] = fmaxf(0.0f, remainL[k] - suml);
      }
    }
    double c = 0.0;                                                   /* :184-224 */
    for (int l = 0; l < m; ++l)
      for (int k = 0; k < n; ++k) c += (double)match[(size_t)l * n + k] * (double)sqrtf(d2f(p1 + k * 3, p2 + l * 3));
    cost[i] = (float)c;
    if (match_out) memcpy(match_out + (size_t)i * n * m, match, sizeof(float) * (size_t)n * m);
  }
  free(match); free(remainL); free(remainR); free(ratioL); free(ratioR);
}
