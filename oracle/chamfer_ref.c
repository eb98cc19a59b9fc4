/* Oracle (TEST INFRASTRUCTURE): plain-C restatement of the reference's nearest-neighbour kernel
 * NmDistanceKernel (external/chamfer3D/chamfer3D.cu:12-134) and gradient kernel
 * NmDistanceGradKernel (chamfer3D.cu:155-174).
 *
 * For every point j of cloud A [b,n,3]: squared distance to, and index of, its nearest neighbour in
 * cloud B [b,m,3]; strict `<` while scanning B in ascending index => lowest index wins ties
 * (chamfer3D.cu:33,44,...; running result kept across 512-point tiles with `result > best`, :126-129,
 * which is again "earlier index wins").  The distance expression `x2*x2+y2*y2+z2*z2` (:32) is
 * evaluated as nvcc 12.9 contracts it for sm_100a (checked in SASS):
 *     d = fma(z2, z2, fma(x2, x2, y2*y2)),   x2 = B.x - A.x (target minus query).
 * Built by oracle/build_oracle.py with -ffp-contract=off so that only the explicit fmaf calls fuse.
 */
#include <math.h>
#include <stddef.h>

void chamfer_nn_ref(const float* xyz, const float* xyz2, int b, int n, int m, float* result, int* result_i) {
  for (int i = 0; i < b; ++i) {
    const float* A = xyz + (size_t)i * n * 3;
    const float* B = xyz2 + (size_t)i * m * 3;
    for (int j = 0; j < n; ++j) {
      float x1 = A[j * 3 + 0], y1 = A[j * 3 + 1], z1 = A[j * 3 + 2];
      float best = 0.0f;
      int best_i = 0;
      for (int k = 0; k < m; ++k) {
        float x2 = B[k * 3 + 0] - x1;
        float y2 = B[k * 3 + 1] - y1;
        float z2 = B[k * 3 + 2] - z1;
        float t = y2 * y2;
        float d = fmaf(z2, z2, fmaf(x2, x2, t));
        if (k == 0 || d < best) {
          best = d;
          best_i = k;
        }
      }
      /* m == 0: the reference never writes; its caller zero-filled the outputs (dist_chamfer_3D.py:29-33) */
      result[(size_t)i * n + j] = best;
      result_i[(size_t)i * n + j] = best_i;
    }
  }
}

/* grad_xyz1 += 2 g (p1 - p2); grad_xyz2 -= 2 g (p1 - p2)  (sequential: deterministic order) */
void chamfer_grad_ref(const float* xyz1, const float* xyz2, int b, int n, int m, const float* grad_dist1,
                      const int* idx1, float* grad_xyz1, float* grad_xyz2) {
  for (int i = 0; i < b; ++i)
    for (int j = 0; j < n; ++j) {
      const float* p1 = xyz1 + ((size_t)i * n + j) * 3;
      int j2 = idx1[(size_t)i * n + j];
      const float* p2 = xyz2 + ((size_t)i * m + j2) * 3;
      float g = grad_dist1[(size_t)i * n + j] * 2;
      for (int c = 0; c < 3; ++c) {
        float v = g * (p1[c] - p2[c]);
        grad_xyz1[((size_t)i * n + j) * 3 + c] += v;
        grad_xyz2[((size_t)i * m + j2) * 3 + c] += -v;
      }
    }
}
