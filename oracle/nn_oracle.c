/* CPU oracle for the nearest-neighbour / Chamfer path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference's NmDistanceKernel
 * (evaluation/pytorch_structural_losses/src/nndistance.cu:2-124): for each point of set A the minimum over
 * set B of the squared distance and the argmin; strict `<` so the lowest index wins ties (:26,36,116-119).
 * The distance is evaluated with the exact contraction nvcc 12.9 emits for `x2*x2+y2*y2+z2*z2` on
 * sm_100a (checked in the PTX: mul y, fma x, fma z), expressed with fmaf() and -ffp-contract=off.
 * Pinned against ChamferDistancePytorch/unit_test.py:22-33's criterion in tests/test_oracle_golden.py and,
 * on the GPU box, against the reference kernel itself (oracle/_ref/libref_nnd.so).
 *
 * Also a restatement of _pairwise_CD_ (evaluation/evaluation_metrics.py:165-198) used as the CPU baseline
 * in bench.py (POSIX threads over cloud pairs).
 */
#include <math.h>
#include <pthread.h>
#include <stddef.h>

static inline float sqdist(float ax, float ay, float az, float bx, float by, float bz) {
  const float dx = bx - ax, dy = by - ay, dz = bz - az;
  return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/* one direction: for each of n points of xyz (per batch) -> min over m points of xyz2 */
static void nn_one_direction(int b, int n, const float* xyz, int m, const float* xyz2, float* result, int* result_i) {
  for (int i = 0; i < b; ++i) {
    const float* A = xyz + (size_t)i * n * 3;
    const float* B = xyz2 + (size_t)i * m * 3;
    for (int j = 0; j < n; ++j) {
      const float ax = A[j * 3 + 0], ay = A[j * 3 + 1], az = A[j * 3 + 2];
      float best = 0.f;
      int best_i = 0;
      for (int k = 0; k < m; ++k) {
        const float d = sqdist(ax, ay, az, B[k * 3 + 0], B[k * 3 + 1], B[k * 3 + 2]);
        if (k == 0 || d < best) {
          best = d;
          best_i = k;
        }
      }
      result[(size_t)i * n + j] = best;
      result_i[(size_t)i * n + j] = best_i;
    }
  }
}

/* nndistance(), nndistance.cu:125-128: both directions */
void oracle_nn_distance(int b, int n, const float* xyz1, int m, const float* xyz2, float* dist1, int* idx1,
                        float* dist2, int* idx2) {
  nn_one_direction(b, n, xyz1, m, xyz2, dist1, idx1);
  nn_one_direction(b, m, xyz2, n, xyz1, dist2, idx2);
}

/* one entry M[i,j] = mean(dl) + mean(dr); means accumulated in double, rounded to float, then one float
 * add (evaluation_metrics.py:191). */
static float cd_pair(int pa, int pb, const float* A, const float* B) {
  double sl = 0.0, sr = 0.0;
  for (int p = 0; p < pa; ++p) {
    float best = INFINITY;
    for (int q = 0; q < pb; ++q) {
      const float d = sqdist(A[p * 3], A[p * 3 + 1], A[p * 3 + 2], B[q * 3], B[q * 3 + 1], B[q * 3 + 2]);
      best = d < best ? d : best;
    }
    sl += best;
  }
  for (int q = 0; q < pb; ++q) {
    float best = INFINITY;
    for (int p = 0; p < pa; ++p) {
      const float d = sqdist(B[q * 3], B[q * 3 + 1], B[q * 3 + 2], A[p * 3], A[p * 3 + 1], A[p * 3 + 2]);
      best = d < best ? d : best;
    }
    sr += best;
  }
  return (float)(sl / pa) + (float)(sr / pb);
}

typedef struct {
  int nb, pa, pb, row_begin, first, last; /* flat pair range [first,last) over (row, col) */
  const float *a, *b;
  float* out;
} cd_job;

static void* cd_worker(void* arg) {
  const cd_job* j = (const cd_job*)arg;
  for (int f = j->first; f < j->last; ++f) {
    const int r = f / j->nb, c = f % j->nb;
    j->out[f] = cd_pair(j->pa, j->pb, j->a + (size_t)(j->row_begin + r) * j->pa * 3, j->b + (size_t)c * j->pb * 3);
  }
  return NULL;
}

/* rows [row_begin,row_end) of the Chamfer matrix, split over `nthreads` POSIX threads. */
void oracle_pairwise_cd(int na, int nb, int pa, int pb, const float* a, const float* b, int row_begin, int row_end,
                        float* out, int nthreads) {
  (void)na;
  const int total = (row_end - row_begin) * nb;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  if (nthreads > total) nthreads = total > 0 ? total : 1;
  pthread_t th[256];
  cd_job jobs[256];
  for (int t = 0; t < nthreads; ++t) {
    cd_job j = {nb, pa, pb, row_begin, (int)((long long)total * t / nthreads), (int)((long long)total * (t + 1) / nthreads), a, b, out};
    jobs[t] = j;
    pthread_create(&th[t], NULL, cd_worker, &jobs[t]);
  }
  for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
}
