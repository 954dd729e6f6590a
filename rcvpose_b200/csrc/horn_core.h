// horn_core.h -- the rotation of Horn's closed-form absolute orientation (what util/horn.py:75-181 computes), written from
// the mathematics of the method rather than from the reference's eigen-solver.  Shared by k_horn (rcvvote.cu, row a-6
// of the path) and the ICP update (refine.cu), which solves the same problem each iteration.
//
// Horn 1987: the unit quaternion q of the rotation that maximises sum_i b_i . (R a_i) is the eigenvector of the largest
// eigenvalue of the symmetric, trace-free 4x4 matrix N built from the cross-covariance S = sum_i a_i b_i^T.  Because
// tr N = 0 its characteristic polynomial has no cubic term,
//        P(l) = l^4 + c2 l^2 + c1 l + c0,     c2 = -2 |S|_F^2,   c1 = -8 det S,   c0 = det N,
// and all four roots are real.  Instead of a general eigen-decomposition (the reference runs up to 50 cyclic Jacobi
// sweeps, util/horn.py:13-72) this solver
//   1. scales S to unit magnitude (eigenvectors are scale-free, the arithmetic stays O(1));
//   2. finds the LARGEST root by Newton's iteration started from Gershgorin's upper bound of the spectrum: for a
//      polynomial with only real roots Newton from above descends monotonically onto the largest one;
//   3. takes the eigenvector as a column of adj(N - l I) (every column of the adjugate of a rank-3 symmetric matrix
//      is a multiple of its null vector), the column of largest norm;
//   4. polishes twice: l <- q^T N q (Rayleigh quotient), q <- adj(N - l I) q (one step of inverse iteration without a
//      division), which brings q to the conditioning limit of the problem -- the same answer as the reference's
//      Jacobi and as SVD-Kabsch to ~1e-15 on well-separated spectra (tests: 1e-12 against the reference's goldens).
// A (numerically) repeated top eigenvalue -- collinear or coincident points, where the optimum is not unique -- makes
// adj vanish; those inputs take a textbook two-sided Jacobi iteration instead (horn_top_eigenvector_jacobi).  S = 0
// returns the identity.
#pragma once

namespace rcv {

__device__ __forceinline__ double horn_det3(double a, double b, double c, double d, double e, double f, double g, double h, double i) {
  return a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
}

// adj(M) for a symmetric 4x4 M (the adjugate of a symmetric matrix is symmetric: 10 cofactors)
__device__ inline void horn_adjugate(const double (*M)[4], double (*A)[4]) {
  for (int i = 0; i < 4; ++i)
    for (int j = i; j < 4; ++j) {
      int r[3], c[3], nr = 0, nc = 0;
      for (int k = 0; k < 4; ++k) { if (k != i) r[nr++] = k; if (k != j) c[nc++] = k; }
      const double m = horn_det3(M[r[0]][c[0]], M[r[0]][c[1]], M[r[0]][c[2]], M[r[1]][c[0]], M[r[1]][c[1]], M[r[1]][c[2]], M[r[2]][c[0]],
                                 M[r[2]][c[1]], M[r[2]][c[2]]);
      A[i][j] = A[j][i] = ((i + j) & 1) ? -m : m;
    }
}

// Fallback for (nearly) repeated top eigenvalues: textbook two-sided Jacobi on the full symmetric matrix (Golub & Van
// Loan, Matrix Computations, "The Jacobi method": for each pivot (p, q) the rotation that annihilates N[p][q] is applied
// as N <- J^T N J, V <- V J; row-cyclic pivots until the off-diagonal mass is gone).  q = column of V with the largest
// diagonal entry.  Unconditionally orthonormal, so the result is a proper rotation whatever the input.
__device__ inline void horn_top_eigenvector_jacobi(const double (*N)[4], double* q) {
  double B[4][4], V[4][4];
  double total = 0.0;
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) { B[r][c] = N[r][c]; V[r][c] = r == c ? 1.0 : 0.0; total += N[r][c] * N[r][c]; }
  for (int sweep = 0; sweep < 40; ++sweep) {
    double off = 0.0;
    for (int r = 0; r < 4; ++r)
      for (int c = r + 1; c < 4; ++c) off += B[r][c] * B[r][c];
    if (off <= 1.0e-34 * total) break;
    for (int p = 0; p < 3; ++p)
      for (int k = p + 1; k < 4; ++k) {
        const double bpk = B[p][k];
        if (bpk == 0.0) continue;
        // tangent of the rotation angle: the smaller root of t^2 + 2 tau t - 1 = 0
        const double tau = (B[k][k] - B[p][p]) / (2.0 * bpk);
        const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
        const double cs = 1.0 / sqrt(1.0 + t * t), sn = t * cs;
        for (int r = 0; r < 4; ++r) {           // columns p and k:  B <- B J
          const double bp = B[r][p], bk = B[r][k];
          B[r][p] = cs * bp - sn * bk; B[r][k] = sn * bp + cs * bk;
          const double vp = V[r][p], vk = V[r][k];
          V[r][p] = cs * vp - sn * vk; V[r][k] = sn * vp + cs * vk;
        }
        for (int c = 0; c < 4; ++c) {           // rows p and k:  B <- J^T B
          const double bp = B[p][c], bk = B[k][c];
          B[p][c] = cs * bp - sn * bk; B[k][c] = sn * bp + cs * bk;
        }
      }
  }
  int best = 0;
  for (int r = 1; r < 4; ++r)
    if (B[r][r] > B[best][best]) best = r;
  for (int r = 0; r < 4; ++r) q[r] = V[r][best];
}

// S[r][c] = sum_i a_i[r] * b_i[c]  (a = centred source, b = centred target);  R maps source to target.
__device__ inline void horn_rotation_from_S(const double (*S_in)[3], double (*R)[3]) {
  double sc = 0.0;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) sc = fmax(sc, fabs(S_in[r][c]));
  if (!(sc > 0.0) || !(sc < 1.0e300)) {   // no correspondence information at all: the identity (what a zero N gives the reference)
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) R[r][c] = r == c ? 1.0 : 0.0;
    return;
  }
  double S[3][3];
  const double inv = 1.0 / sc;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) S[r][c] = S_in[r][c] * inv;
  // Horn's N (symmetric, trace-free)
  double N[4][4];
  N[0][0] = S[0][0] + S[1][1] + S[2][2];
  N[1][1] = S[0][0] - S[1][1] - S[2][2];
  N[2][2] = S[1][1] - S[0][0] - S[2][2];
  N[3][3] = S[2][2] - S[0][0] - S[1][1];
  N[0][1] = N[1][0] = S[1][2] - S[2][1];
  N[0][2] = N[2][0] = S[2][0] - S[0][2];
  N[0][3] = N[3][0] = S[0][1] - S[1][0];
  N[1][2] = N[2][1] = S[0][1] + S[1][0];
  N[1][3] = N[3][1] = S[2][0] + S[0][2];
  N[2][3] = N[3][2] = S[1][2] + S[2][1];
  // characteristic polynomial
  double fro = 0.0;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) fro += S[r][c] * S[r][c];
  const double c2 = -2.0 * fro;
  const double c1 = -8.0 * horn_det3(S[0][0], S[0][1], S[0][2], S[1][0], S[1][1], S[1][2], S[2][0], S[2][1], S[2][2]);
  double c0 = 0.0;
  {
    double A[4][4];
    horn_adjugate(N, A);                       // det N = row 0 of N times column 0 of adj N
    for (int k = 0; k < 4; ++k) c0 += N[0][k] * A[k][0];
  }
  // largest root: Newton from Gershgorin's bound
  double lam = 0.0;
  for (int r = 0; r < 4; ++r) lam = fmax(lam, fabs(N[r][0]) + fabs(N[r][1]) + fabs(N[r][2]) + fabs(N[r][3]));
  for (int it = 0; it < 64; ++it) {
    const double l2 = lam * lam;
    const double P = (l2 + c2) * l2 + c1 * lam + c0;
    const double dP = (4.0 * l2 + 2.0 * c2) * lam + c1;
    if (dP == 0.0) break;
    const double step = P / dP;
    lam -= step;
    if (fabs(step) <= 4.0e-16 * fabs(lam)) break;
  }
  // eigenvector: largest column of adj(N - lam I), then two Rayleigh / inverse-iteration polishing steps.
  // adj(N - l1 I) = (product of the gaps l1 - l_j) v1 v1^T: its entries carry ~1e-16 of absolute rounding noise, so the
  // column is trusted only while that product is large enough for a 1e-12 eigenvector (norm^2 >= 1e-8); closer spectra --
  // (nearly) collinear or coincident points, where the optimum is not unique anyway -- go to the Jacobi fallback.
  double q[4] = {1.0, 0.0, 0.0, 0.0};
  bool have_q = false;
  {
    double M[4][4], A[4][4];
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) M[r][c] = N[r][c] - (r == c ? lam : 0.0);
    horn_adjugate(M, A);
    int best = 0; double bn = -1.0;
    for (int c = 0; c < 4; ++c) {
      const double n2 = A[0][c] * A[0][c] + A[1][c] * A[1][c] + A[2][c] * A[2][c] + A[3][c] * A[3][c];
      if (n2 > bn) { bn = n2; best = c; }
    }
    if (bn >= 1.0e-8) {
      have_q = true;
      const double s = 1.0 / sqrt(bn);
      for (int k = 0; k < 4; ++k) q[k] = A[k][best] * s;
      for (int pol = 0; pol < 2; ++pol) {
        double l = 0.0;
        for (int r = 0; r < 4; ++r) l += q[r] * (N[r][0] * q[0] + N[r][1] * q[1] + N[r][2] * q[2] + N[r][3] * q[3]);
        for (int r = 0; r < 4; ++r)
          for (int c = 0; c < 4; ++c) M[r][c] = N[r][c] - (r == c ? l : 0.0);
        horn_adjugate(M, A);
        double v[4], n2 = 0.0;
        for (int r = 0; r < 4; ++r) { v[r] = A[r][0] * q[0] + A[r][1] * q[1] + A[r][2] * q[2] + A[r][3] * q[3]; n2 += v[r] * v[r]; }
        if (!(n2 >= 1.0e-8)) break;
        const double t = 1.0 / sqrt(n2);
        for (int r = 0; r < 4; ++r) q[r] = v[r] * t;
      }
    }
  }
  if (!have_q) horn_top_eigenvector_jacobi(N, q);
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  R[0][0] = w * w + x * x - y * y - z * z; R[0][1] = 2.0 * (x * y - w * z);         R[0][2] = 2.0 * (x * z + w * y);
  R[1][0] = 2.0 * (x * y + w * z);         R[1][1] = w * w - x * x + y * y - z * z; R[1][2] = 2.0 * (y * z - w * x);
  R[2][0] = 2.0 * (x * z - w * y);         R[2][1] = 2.0 * (y * z + w * x);         R[2][2] = w * w - x * x - y * y + z * z;
}

}  // namespace rcv
