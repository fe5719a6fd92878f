// oracle/_ref harness (TEST INFRASTRUCTURE ONLY): C entry points around the reference's own
// AVX2 polar projections, compiled from the sources where they lie under /root/reference
// (C++/DPGO/src/internal/project_to_SOd.cpp, traits.cpp; see oracle/Makefile).  Only the lane
// packing below is ours; it follows project_to_SO3_d / project_to_SO2_d
// (C++/DPGO/include/DPGO/internal/project_to_SOd.h:12-98) and the tail handling of
// project_to_SO3n / project_to_SO2n (C++/DPGO/include/DPGO/DPGO_utils.h:515-565).
#include <immintrin.h>

namespace DPGO {
namespace internal {
void project_to_SO2(const __m256d &a11, const __m256d &a12, const __m256d &a21, const __m256d &a22, __m256d &u11,
                    __m256d &u21);
void project_to_SO3(const __m256d &a11, const __m256d &a12, const __m256d &a13, const __m256d &a21,
                    const __m256d &a22, const __m256d &a23, const __m256d &a31, const __m256d &a32,
                    const __m256d &a33, __m256d &u11, __m256d &u12, __m256d &u13, __m256d &u21, __m256d &u22,
                    __m256d &u23, __m256d &u31, __m256d &u32, __m256d &u33);
}  // namespace internal
}  // namespace DPGO

static void so3_batch4(const double *A, double *U) {
  // A, U: four consecutive row-major 3 x 3 blocks
  double temp[9][4];
  for (int k = 0; k < 4; ++k)
    for (int e = 0; e < 9; ++e) temp[e][k] = A[9 * k + e];
  __m256d a[9], u[9];
  for (int e = 0; e < 9; ++e) a[e] = _mm256_loadu_pd(temp[e]);
  DPGO::internal::project_to_SO3(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], u[0], u[1], u[2], u[3], u[4],
                                 u[5], u[6], u[7], u[8]);
  for (int e = 0; e < 9; ++e) _mm256_storeu_pd(temp[e], u[e]);
  for (int k = 0; k < 4; ++k)
    for (int e = 0; e < 9; ++e) U[9 * k + e] = temp[e][k];
}

static void so2_batch4(const double *A, double *U) {
  double temp[4][4];
  for (int k = 0; k < 4; ++k)
    for (int e = 0; e < 4; ++e) temp[e][k] = A[4 * k + e];
  __m256d a[4], u[2];
  for (int e = 0; e < 4; ++e) a[e] = _mm256_loadu_pd(temp[e]);
  DPGO::internal::project_to_SO2(a[0], a[1], a[2], a[3], u[0], u[1]);
  double t0[4], t1[4];
  _mm256_storeu_pd(t0, u[0]);
  _mm256_storeu_pd(t1, u[1]);
  for (int k = 0; k < 4; ++k) {
    U[4 * k + 0] = t0[k]; U[4 * k + 1] = -t1[k];
    U[4 * k + 2] = t1[k]; U[4 * k + 3] = t0[k];
  }
}

extern "C" {
// n >= 4 row-major d x d blocks; the last (possibly overlapping) group of four is recomputed
// exactly like the reference's bottomLeftCorner call.  Returns -1 for n < 4 (the reference
// falls back to Eigen::JacobiSVD there, which is not available here).
int ref_project_to_SO3n(const double *A, double *U, long n) {
  if (n < 4) return -1;
  for (long i = 0; i + 4 < n; i += 4) so3_batch4(A + 9 * i, U + 9 * i);
  so3_batch4(A + 9 * (n - 4), U + 9 * (n - 4));
  return 0;
}
int ref_project_to_SO2n(const double *A, double *U, long n) {
  if (n < 4) return -1;
  for (long i = 0; i + 4 < n; i += 4) so2_batch4(A + 4 * i, U + 4 * i);
  so2_batch4(A + 4 * (n - 4), U + 4 * (n - 4));
  return 0;
}
}
