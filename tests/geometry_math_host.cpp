// Host build of the similarity-fit math the CUDA kernel runs per view (fast3r_b200/csrc/geometry_math.h), so the CPU
// suite checks that exact code against the oracle.  Compiled by tests/test_geometry_math_cpu.py with g++.
#include "geometry_math.h"

extern "C" void f3r_test_umeyama_from_moments(const double* moments, float* rts) { f3r::umeyama_from_moments(moments, rts); }

extern "C" void f3r_test_eig3(const double* sym, double* vec, double* lam) {
  double a[3][3], v[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) a[i][j] = sym[3 * i + j];
  f3r::jacobi_eig3(a, v, lam);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) vec[3 * i + j] = v[i][j];
}
