// Host/device math of the similarity fit (geometry.cu): symmetric 3x3 eigen-decomposition and Umeyama from raw moments.
// Kept in a header that also compiles as plain C++ so the CPU test suite exercises exactly this code
// (tests/geometry_math_host.cpp) - the kernels around it only stream and reduce.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define F3R_HD __host__ __device__
#else
#define F3R_HD
#endif

namespace f3r {

constexpr int MOM = 17;  // count, sum x (3), sum y (3), sum |x|^2, sum y_i x_j (9, row-major in i)

// eigen-decomposition of a symmetric 3x3 (cyclic Jacobi), eigenvalues descending, eigenvectors in the columns of v
F3R_HD inline void jacobi_eig3(double a[3][3], double v[3][3], double lam[3]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) v[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 32; ++sweep) {
    const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
    const double diag = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
    if (off <= 1e-34 * diag || off == 0.0) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        const double apq = a[p][q];
        if (apq == 0.0) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
        const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(tt * tt + 1.0), s = tt * c;
        a[p][p] -= tt * apq;
        a[q][q] += tt * apq;
        a[p][q] = a[q][p] = 0.0;
        const int r = 3 - p - q;
        const double arp = a[r][p], arq = a[r][q];
        a[r][p] = a[p][r] = c * arp - s * arq;
        a[r][q] = a[q][r] = s * arp + c * arq;
        for (int k = 0; k < 3; ++k) {
          const double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = c * vkp - s * vkq;
          v[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int order[3] = {0, 1, 2};
  double d[3] = {a[0][0], a[1][1], a[2][2]};
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2 - i; ++j)
      if (d[order[j]] < d[order[j + 1]]) {
        const int tmp = order[j];
        order[j] = order[j + 1];
        order[j + 1] = tmp;
      }
  double vs[3][3];
  for (int c = 0; c < 3; ++c) {
    lam[c] = d[order[c]];
    for (int k = 0; k < 3; ++k) vs[k][c] = v[k][order[c]];
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) v[i][j] = vs[i][j];
}

F3R_HD inline void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
F3R_HD inline double norm3(const double* a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

// Umeyama from the moments of one point set: M = sum (y - ym)(x - xm)^T = U S V^T,
// R = u1 v1^T + u2 v2^T + (u1 x u2)(v1 x v2)^T  (= U diag(1, 1, det(U V^T)) V^T), s = (S1 + S2 + sign(det M) S3) / sum |x - xm|^2
F3R_HD inline void umeyama_from_moments(const double* m, float* rts) {
  const double n = m[0];
  const double xm[3] = {m[1] / n, m[2] / n, m[3] / n};
  const double ym[3] = {m[4] / n, m[5] / n, m[6] / n};
  const double var = m[7] - n * (xm[0] * xm[0] + xm[1] * xm[1] + xm[2] * xm[2]);
  double mm[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) mm[i][j] = m[8 + 3 * i + j] - n * ym[i] * xm[j];
  double ata[3][3], v[3][3], lam[3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) ata[i][j] = mm[0][i] * mm[0][j] + mm[1][i] * mm[1][j] + mm[2][i] * mm[2][j];
  jacobi_eig3(ata, v, lam);
  double sig[3];
  for (int i = 0; i < 3; ++i) sig[i] = sqrt(fmax(lam[i], 0.0));
  double v1[3] = {v[0][0], v[1][0], v[2][0]}, v2[3] = {v[0][1], v[1][1], v[2][1]}, v3[3];
  double u1[3], u2[3], u3[3];
  for (int i = 0; i < 3; ++i) u1[i] = mm[i][0] * v1[0] + mm[i][1] * v1[1] + mm[i][2] * v1[2];
  double l1 = norm3(u1);
  if (l1 > 0.0) {
    for (int i = 0; i < 3; ++i) u1[i] /= l1;
  } else {
    u1[0] = 1.0; u1[1] = 0.0; u1[2] = 0.0;
  }
  for (int i = 0; i < 3; ++i) u2[i] = mm[i][0] * v2[0] + mm[i][1] * v2[1] + mm[i][2] * v2[2];
  const double dp = u2[0] * u1[0] + u2[1] * u1[1] + u2[2] * u1[2];
  for (int i = 0; i < 3; ++i) u2[i] -= dp * u1[i];
  double l2 = norm3(u2);
  if (l2 > 1e-12 * fmax(sig[0], 1e-300)) {
    for (int i = 0; i < 3; ++i) u2[i] /= l2;
  } else {  // collinear points: any unit vector orthogonal to u1
    const double e[3] = {fabs(u1[0]) < 0.9 ? 1.0 : 0.0, fabs(u1[0]) < 0.9 ? 0.0 : 1.0, 0.0};
    cross3(u1, e, u2);
    l2 = norm3(u2);
    for (int i = 0; i < 3; ++i) u2[i] /= l2;
  }
  cross3(u1, u2, u3);
  cross3(v1, v2, v3);
  double r[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r[i][j] = u1[i] * v1[j] + u2[i] * v2[j] + u3[i] * v3[j];
  const double det = mm[0][0] * (mm[1][1] * mm[2][2] - mm[1][2] * mm[2][1]) -
                     mm[0][1] * (mm[1][0] * mm[2][2] - mm[1][2] * mm[2][0]) +
                     mm[0][2] * (mm[1][0] * mm[2][1] - mm[1][1] * mm[2][0]);
  const double scale = (sig[0] + sig[1] + (det < 0.0 ? -sig[2] : sig[2])) / var;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) rts[3 * i + j] = static_cast<float>(r[i][j]);
    rts[9 + i] = static_cast<float>(ym[i] - scale * (r[i][0] * xm[0] + r[i][1] * xm[1] + r[i][2] * xm[2]));
  }
  rts[12] = static_cast<float>(scale);
}

}  // namespace f3r
