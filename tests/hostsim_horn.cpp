// Host build of rcvpose_b200/csrc/horn_core.h (TEST INFRASTRUCTURE ONLY): the device function compiled for the CPU so that the
// Horn solver can be checked against the reference's goldens in the CPU-only container.  Mirrors k_horn (rcvvote.cu).
#include <cmath>
#define __device__
#define __forceinline__ inline
using std::fmax; using std::fabs; using std::sqrt;
#include "../rcvpose_b200/csrc/horn_core.h"

extern "C" __attribute__((visibility("default")))
void hostsim_horn(const double* P1, const double* P2, int n, double* RT) {
  double C1[3] = {0, 0, 0}, C2[3] = {0, 0, 0};
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < 3; ++j) { C1[j] += P1[3 * i + j]; C2[j] += P2[3 * i + j]; }
  for (int j = 0; j < 3; ++j) { C1[j] /= n; C2[j] /= n; }
  double S[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int i = 0; i < n; ++i) {
    double a[3], b[3];
    for (int j = 0; j < 3; ++j) { a[j] = P1[3 * i + j] - C1[j]; b[j] = P2[3 * i + j] - C2[j]; }
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) S[r][c] += a[r] * b[c];
  }
  double R[3][3];
  rcv::horn_rotation_from_S(S, R);
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) RT[4 * r + c] = R[r][c];
    RT[4 * r + 3] = C2[r] - (R[r][0] * C1[0] + R[r][1] * C1[1] + R[r][2] * C1[2]);
    RT[12 + r] = 0.0;
  }
  RT[15] = 1.0;
}
