// Test shim: compiles pysages_b200/csrc/psb_math.cuh for the HOST with g++ so the closed-form CV
// gradients, the Kabsch rotation, the k x k solvers and the bit-exact bin indexer can be unit-tested
// against the oracle on a machine without a GPU.  Test infrastructure only; never loaded by the
// package.
#include "../../pysages_b200/csrc/psb_math.cuh"

using namespace psb;

extern "C" {

double hm_distance(const double* p, double* g) {
  V3 a, b;
  double v = distance_cv(v3(p[0], p[1], p[2]), v3(p[3], p[4], p[5]), &a, &b);
  g[0] = a.x; g[1] = a.y; g[2] = a.z; g[3] = b.x; g[4] = b.y; g[5] = b.z;
  return v;
}

double hm_angle(const double* p, double* g) {
  V3 o[3];
  double v = angle_cv(v3(p[0], p[1], p[2]), v3(p[3], p[4], p[5]), v3(p[6], p[7], p[8]), &o[0], &o[1], &o[2]);
  for (int i = 0; i < 3; ++i) { g[3 * i] = o[i].x; g[3 * i + 1] = o[i].y; g[3 * i + 2] = o[i].z; }
  return v;
}

double hm_dihedral(const double* p, double* g) {
  V3 o[4];
  double v = dihedral_cv(v3(p[0], p[1], p[2]), v3(p[3], p[4], p[5]), v3(p[6], p[7], p[8]),
                         v3(p[9], p[10], p[11]), &o[0], &o[1], &o[2], &o[3]);
  for (int i = 0; i < 4; ++i) { g[3 * i] = o[i].x; g[3 * i + 1] = o[i].y; g[3 * i + 2] = o[i].z; }
  return v;
}

double hm_gyration(int mode, int ja, int jb, const double* g6, double* W9) { return gyration_cv(mode, ja, jb, g6, W9); }
// Cremer-Pople ring puckering exactly as the step kernel evaluates it: Fourier sums over the ring
// (accumulate_nonpoint), ring_cv, then the per-atom gradient (nonpoint_gradient)
double hm_ring(int mode, int n, const double* pos, double* grad) {
  double sums[12] = {0}, E[12];
  for (int t = 0; t < n; ++t) {
    double s, c, S, C;
    ring_coefficients((uint32_t)t, (uint32_t)n, &s, &c, &S, &C);
    for (int a = 0; a < 3; ++a) {
      sums[a] += s * pos[3 * t + a];
      sums[3 + a] += c * pos[3 * t + a];
      sums[6 + a] += C * pos[3 * t + a];
      sums[9 + a] += S * pos[3 * t + a];
    }
  }
  const double v = ring_cv(mode, (uint32_t)n, sums, E);
  for (int t = 0; t < n; ++t) {
    double s, c, S, C;
    ring_coefficients((uint32_t)t, (uint32_t)n, &s, &c, &S, &C);
    for (int a = 0; a < 3; ++a) grad[3 * t + a] = C * E[a] + S * E[3 + a] + s * E[6 + a] + c * E[9 + a];
  }
  return v;
}
void hm_kabsch(const double* C, double* U) { kabsch_rotation(C, U); }
void hm_kabsch_eig(const double* C, double* U) { kabsch_rotation_eig(C, U); }

unsigned hm_grid_index(int kind, double x, double lower, double upper, int shape) {
  return grid_index_1d(kind, x, lower, upper, shape);
}

double hm_mesh_coord(int kind, int k, double lower, double upper, int shape) {
  return mesh_coord(kind, k, lower, upper, shape);
}

int hm_chol_solve(int k, const double* A4, const double* b, double* x) { return chol_solve(k, A4, b, x) ? 1 : 0; }
int hm_spd_solve(int k, const double* A4, const double* b, double* x) { return spd_solve(k, A4, b, x) ? 1 : 0; }
unsigned hm_grid_index_slow(int kind, double x, double lower, double upper, int shape) {
  return grid_index_slow(kind, x, lower, upper, shape);
}

void hm_pinv_solve(int k, const double* A4, const double* b, double* x, double rtol) {
  pinv_sym_solve(k, A4, b, x, rtol);
}

double hm_wrap(double x, double P) { return wrap_period(x, P); }
}
