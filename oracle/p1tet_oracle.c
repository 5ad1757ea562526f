/* CPU ORACLE -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Scalar C restatement of the arithmetic contract of the
 * headline path: ElementTetP1 Laplace + unit load on an affine tetrahedral
 * mesh.  One element at a time, one IEEE operation per statement, in the order
 * numpy executes the reference's expressions.  Compile with
 * -ffp-contract=off (GCC would otherwise fuse a*b+c under -march=native).
 *
 * Parity status: PINNED -- tests/test_oracle_c.py checks the output bitwise
 * against the golden vectors produced by the real reference
 * (tests/golden/tet_p1_*.npz) and against the numpy oracle.
 *
 * Reference lines (paths relative to /root/reference/skfem/):
 *   geometry   mapping/mapping_affine.py:55-131   (A, detA, invA)
 *   push-fwd   element/element_h1.py:17 + element_tet/element_tet_p1.py:19-45
 *   dx         assembly/basis/cell_basis.py:104-105
 *   laplace    models/poisson.py:7-9, helpers.py:108-110 (einsum 'i...,i...')
 *   qp sum     assembly/form/bilinear_form.py:150-151 (np.sum, n=4: sequential)
 *   unit_load  models/poisson.py:22-24, assembly/form/linear_form.py:41-49
 *   scatter    assembly/form/coo_data.py:102-108 (scipy coo_todense: COO order)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

/* local[j][i][e] for j,i in 0..3 (data[j,i,:] of bilinear_form.py:93-98),
 * p: (3, npts) coordinate-major, t: (4, nel) node-major, W: (nqp) weights,
 * dphi: (4, 3, nqp) reference gradients.  Returns 0. */
int p1tet_laplace_local(const double *p, int64_t npts, const int32_t *t, int64_t nel,
                        const double *dphi, const double *W, int nqp, double *local) {
  for (int64_t e = 0; e < nel; ++e) {
    double A[3][3], inv[3][3];
    const int32_t v0 = t[e];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        const int32_t vj = t[(int64_t)(j + 1) * nel + e];
        A[i][j] = p[(int64_t)i * npts + vj] - p[(int64_t)i * npts + v0];
      }
    /* mapping_affine.py:92-98, left to right */
    const double m0 = A[1][1] * A[2][2] - A[1][2] * A[2][1];
    const double m1 = A[1][0] * A[2][2] - A[1][2] * A[2][0];
    const double m2 = A[1][0] * A[2][1] - A[1][1] * A[2][0];
    const double det = (A[0][0] * m0 - A[0][1] * m1) + A[0][2] * m2;
    /* mapping_affine.py:111-129 */
    inv[0][0] = (-A[1][2] * A[2][1] + A[1][1] * A[2][2]) / det;
    inv[1][0] = (A[1][2] * A[2][0] - A[1][0] * A[2][2]) / det;
    inv[2][0] = (-A[1][1] * A[2][0] + A[1][0] * A[2][1]) / det;
    inv[0][1] = (A[0][2] * A[2][1] - A[0][1] * A[2][2]) / det;
    inv[1][1] = (-A[0][2] * A[2][0] + A[0][0] * A[2][2]) / det;
    inv[2][1] = (A[0][1] * A[2][0] - A[0][0] * A[2][1]) / det;
    inv[0][2] = (-A[0][2] * A[1][1] + A[0][1] * A[1][2]) / det;
    inv[1][2] = (A[0][2] * A[1][0] - A[0][0] * A[1][2]) / det;
    inv[2][2] = (-A[0][1] * A[1][0] + A[0][0] * A[1][1]) / det;
    const double absdet = fabs(det);
    for (int j = 0; j < 4; ++j)
      for (int i = 0; i < 4; ++i) {
        double r = 0.0; /* np.sum over nqp < 8 terms: sequential from 0 */
        for (int q = 0; q < nqp; ++q) {
          double gu[3], gv[3];
          for (int k = 0; k < 3; ++k) { /* einsum('ijkl,il->jkl'): sequential in i */
            double au = inv[0][k] * dphi[(j * 3 + 0) * nqp + q];
            au = au + inv[1][k] * dphi[(j * 3 + 1) * nqp + q];
            au = au + inv[2][k] * dphi[(j * 3 + 2) * nqp + q];
            double av = inv[0][k] * dphi[(i * 3 + 0) * nqp + q];
            av = av + inv[1][k] * dphi[(i * 3 + 1) * nqp + q];
            av = av + inv[2][k] * dphi[(i * 3 + 2) * nqp + q];
            gu[k] = au;
            gv[k] = av;
          }
          double d = gu[0] * gv[0];
          d = d + gu[1] * gv[1];
          d = d + gu[2] * gv[2];
          const double dx = absdet * W[q];
          r = r + d * dx;
        }
        local[((int64_t)(j * 4 + i)) * nel + e] = r;
      }
  }
  return 0;
}

/* unit_load vector: local data (4, nel) then the sequential COO-order scatter. */
int p1tet_unit_load(const double *p, int64_t npts, const int32_t *t, int64_t nel,
                    const double *phi, const double *W, int nqp, int64_t N, double *vec) {
  double *local = (double *)malloc(sizeof(double) * 4 * (size_t)nel);
  if (!local) return 1;
  for (int64_t e = 0; e < nel; ++e) {
    double A[3][3];
    const int32_t v0 = t[e];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        const int32_t vj = t[(int64_t)(j + 1) * nel + e];
        A[i][j] = p[(int64_t)i * npts + vj] - p[(int64_t)i * npts + v0];
      }
    const double m0 = A[1][1] * A[2][2] - A[1][2] * A[2][1];
    const double m1 = A[1][0] * A[2][2] - A[1][2] * A[2][0];
    const double m2 = A[1][0] * A[2][1] - A[1][1] * A[2][0];
    const double absdet = fabs((A[0][0] * m0 - A[0][1] * m1) + A[0][2] * m2);
    for (int i = 0; i < 4; ++i) {
      double r = 0.0;
      for (int q = 0; q < nqp; ++q) r = r + phi[i * nqp + q] * (absdet * W[q]);
      local[(int64_t)i * nel + e] = r;
    }
  }
  for (int64_t k = 0; k < N; ++k) vec[k] = 0.0;
  for (int i = 0; i < 4; ++i) /* coo_todense: COO order = basis-major, element-minor */
    for (int64_t e = 0; e < nel; ++e) vec[t[(int64_t)i * nel + e]] += local[(int64_t)i * nel + e];
  free(local);
  return 0;
}
