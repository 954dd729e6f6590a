// horn_core.h -- rotation of Horn's closed-form absolute orientation (util/horn.py:109-171) as a device function, shared by
// k_horn (rcvvote.cu, the path's a-6) and the ICP refinement (refine.cu, whose per-iteration update is the same problem).
// Quaternion method: eigenvector of the largest eigenvalue of the symmetric 4x4 matrix N built from the cross-covariance
// sums S[r][c] = sum_i a_i[r] * b_i[c] (a = centred source, b = centred target); cyclic Jacobi with the reference's sweep
// order, thresholds and 50-sweep cap (myjacobi, util/horn.py:13-72).  R maps source to target.
#pragma once

namespace rcv {

__device__ __forceinline__ void jac_rot(double (*a)[4], int i, int j, int k, int l, double s, double tau) {
  const double g = a[i][j], h = a[k][l];
  a[i][j] = g - s * (h + g * tau);
  a[k][l] = h + s * (g - h * tau);
}

__device__ inline void horn_rotation_from_S(const double (*S)[3], double (*R)[3]) {
  double A[4][4], V[4][4], d[4], bq[4], zq[4];
  A[0][0] = S[0][0] + S[1][1] + S[2][2]; A[0][1] = S[1][2] - S[2][1]; A[0][2] = S[2][0] - S[0][2]; A[0][3] = S[0][1] - S[1][0];
  A[1][0] = A[0][1]; A[1][1] = S[0][0] - S[1][1] - S[2][2]; A[1][2] = S[0][1] + S[1][0]; A[1][3] = S[2][0] + S[0][2];
  A[2][0] = A[0][2]; A[2][1] = A[1][2]; A[2][2] = -S[0][0] + S[1][1] - S[2][2]; A[2][3] = S[1][2] + S[2][1];
  A[3][0] = A[0][3]; A[3][1] = A[1][3]; A[3][2] = A[2][3]; A[3][3] = -S[0][0] - S[1][1] + S[2][2];
  for (int p = 0; p < 4; ++p) {
    for (int q = 0; q < 4; ++q) V[p][q] = 0.0;
    V[p][p] = 1.0;
    bq[p] = d[p] = A[p][p];
    zq[p] = 0.0;
  }
  for (int sweep = 1; sweep <= 50; ++sweep) {
    double sm = 0.0;
    for (int p = 0; p < 3; ++p)
      for (int q = 0; q < 4; ++q) sm += fabs(A[p][q]);  // util/horn.py:28-30 sums whole rows, diagonal included
    if (sm == 0.0) break;
    const double tresh = sweep < 4 ? 0.2 * sm / 16.0 : 0.0;
    for (int p = 0; p < 3; ++p)
      for (int q = p + 1; q < 4; ++q) {
        const double g = 100.0 * fabs(A[p][q]);
        if (sweep > 4 && fabs(d[p]) + g == fabs(d[p]) && fabs(d[q]) + g == fabs(d[q])) A[p][q] = 0.0;
        else if (fabs(A[p][q]) > tresh) {
          double h = d[q] - d[p], t;
          if (fabs(h) + g == fabs(h)) t = A[p][q] / h;
          else {
            const double theta = 0.5 * h / A[p][q];
            t = 1.0 / (fabs(theta) + sqrt(1.0 + theta * theta));
            if (theta < 0.0) t = -t;
          }
          const double c = 1.0 / sqrt(1 + t * t), s = t * c, tau = s / (1.0 + c);
          h = t * A[p][q];
          zq[p] -= h; zq[q] += h; d[p] -= h; d[q] += h;
          A[p][q] = 0.0;
          for (int j = 0; j < p; ++j) jac_rot(A, j, p, j, q, s, tau);
          for (int j = p + 1; j < q; ++j) jac_rot(A, p, j, j, q, s, tau);
          for (int j = q + 1; j < 4; ++j) jac_rot(A, p, j, q, j, s, tau);
          for (int j = 0; j < 4; ++j) jac_rot(V, j, p, j, q, s, tau);
        }
      }
    for (int p = 0; p < 4; ++p) { bq[p] += zq[p]; d[p] = bq[p]; zq[p] = 0.0; }
  }
  int me = 0;
  for (int p = 1; p < 4; ++p)
    if (d[p] > d[me]) me = p;
  const double q0 = V[0][me], q1 = V[1][me], q2 = V[2][me], q3 = V[3][me];
  R[0][0] = q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3; R[0][1] = 2 * (q1 * q2 - q0 * q3); R[0][2] = 2 * (q1 * q3 + q0 * q2);
  R[1][0] = 2 * (q1 * q2 + q0 * q3); R[1][1] = q0 * q0 + q2 * q2 - q1 * q1 - q3 * q3; R[1][2] = 2 * (q2 * q3 - q0 * q1);
  R[2][0] = 2 * (q1 * q3 - q0 * q2); R[2][1] = 2 * (q2 * q3 + q0 * q1); R[2][2] = q0 * q0 + q3 * q3 - q1 * q1 - q2 * q2;
}

}  // namespace rcv
