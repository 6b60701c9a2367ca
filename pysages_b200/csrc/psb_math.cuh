// psb_math.cuh -- scalar fp64 building blocks of the bias step, shared by the CUDA kernels
// (device) and by tests/hostmath (compiled with g++ for the CPU-only unit tests).
// Closed-form CV gradients replace the reference's jax.grad (colvars/core.py:298-299).
#pragma once

#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define PSB_HD __host__ __device__ __forceinline__
// big, rarely executed routines: one out-of-line copy keeps the kernels' instruction footprint small
#define PSB_HD_COLD __host__ __device__ __noinline__ inline
#else
#define PSB_HD inline
#define PSB_HD_COLD inline
#endif

// exact (never FMA-contracted) multiply / add, for arithmetic whose bits must match the reference
#ifdef __CUDA_ARCH__
#define PSB_MUL(a, b) __dmul_rn((a), (b))
#define PSB_ADD(a, b) __dadd_rn((a), (b))
#else
#define PSB_MUL(a, b) ((a) * (b))
#define PSB_ADD(a, b) ((a) + (b))
#endif

namespace psb {

struct V3 {
  double x, y, z;
};

PSB_HD V3 v3(double x, double y, double z) { return V3{x, y, z}; }
PSB_HD V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
PSB_HD V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
PSB_HD V3 operator*(double s, V3 a) { return V3{s * a.x, s * a.y, s * a.z}; }
PSB_HD double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
PSB_HD V3 cross(V3 a, V3 b) {
  return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
PSB_HD double norm(V3 a) { return sqrt(dot(a, a)); }
PSB_HD double comp(V3 a, int c) { return c == 0 ? a.x : (c == 1 ? a.y : a.z); }

// ---- point CVs ------------------------------------------------------------------------
// colvars/coordinates.py:91-109  distance = |r1 - r2|;  d/dr1 = u, d/dr2 = -u
PSB_HD double distance_cv(V3 c1, V3 c2, V3* g1, V3* g2) {
  V3 d = c1 - c2;
  double r = norm(d);
  double inv = 1.0 / r;
  *g1 = inv * d;
  *g2 = (-inv) * d;
  return r;
}

// colvars/angles.py:46-74  theta = atan2(|q x r|, q.r), q = p1-p2, r = p3-p2
PSB_HD double angle_cv(V3 p1, V3 p2, V3 p3, V3* g1, V3* g2, V3* g3) {
  V3 q = p1 - p2, r = p3 - p2;
  V3 n = cross(q, r);
  double nn = norm(n);
  double theta = atan2(nn, dot(q, r));
  // d(theta)/dq = (q x n) / (|q|^2 |n|),  d(theta)/dr = (n x r) / (|r|^2 |n|)
  V3 a = (1.0 / (dot(q, q) * nn)) * cross(q, n);
  V3 b = (1.0 / (dot(r, r) * nn)) * cross(n, r);
  *g1 = a;
  *g3 = b;
  *g2 = (-1.0) * (a + b);
  return theta;
}

// colvars/angles.py:95-128  q = p3-p2, r = (p2-p1) x q, s = q x (p4-p3),
// theta = atan2((r x s).q, (r.s)|q|).  Gradient: Blondel-Karplus form with the reference's sign.
PSB_HD double dihedral_cv(V3 p1, V3 p2, V3 p3, V3 p4, V3* g1, V3* g2, V3* g3, V3* g4) {
  V3 q = p3 - p2;
  V3 r = cross(p2 - p1, q);
  V3 s = cross(q, p4 - p3);
  double theta = atan2(dot(cross(r, s), q), dot(r, s) * norm(q));
  V3 F = p1 - p2, G = p2 - p3, H = p4 - p3;
  V3 A = cross(F, G), B = cross(H, G);
  double A2 = dot(A, A), B2 = dot(B, B), Gn = norm(G);
  double fg = dot(F, G), hg = dot(H, G);
  V3 dA = (Gn / A2) * A;
  V3 dB = (Gn / B2) * B;
  V3 t = (fg / (A2 * Gn)) * A - (hg / (B2 * Gn)) * B;
  *g1 = (-1.0) * dA;
  *g4 = dB;
  *g2 = dA + t;
  *g3 = (-1.0) * dB - t;
  return theta;
}

// ---- colvars/utils.py:22-26 -------------------------------------------------------------
PSB_HD double wrap_period(double x, double P) {
  if (isinf(P)) return x;
  if (fabs(x) > P / 2) return x - (x > 0 ? P : -P);
  return x;
}

// ---- methods/restraints.py:68-73 -------------------------------------------------------
PSB_HD double restraint_force(double lo, double hi, double kl, double kh, double xi) {
  if (xi < lo) return kl * (xi - lo);
  if (xi > hi) return kh * (xi - hi);
  return 0.0;
}

// ---- grids.py:109-158: bit-exact bin index ------------------------------------------------
// jnp.floor_divide on floats (jax/_src/numpy/ufuncs.py _float_divmod): fmod-based, then
// lax.round (half away from zero).  NOT floor(x / h): at exact bin edges they differ.
PSB_HD double jax_floor_divide(double x1, double x2) {
  double mod = fmod(x1, x2);
  double div = (x1 - mod) / x2;
  bool ind = (mod != 0.0) && ((x2 < 0.0) != (mod < 0.0));
  if (ind) div = div - 1.0;
  return round(div);
}

PSB_HD double jax_remainder(double x1, double x2) {
  double t = fmod(x1, x2);
  bool plus = ((t < 0.0) != (x2 < 0.0)) && (t != 0.0);
  return plus ? t + x2 : t;
}

// The same value as jax_floor_divide(a, b) for b > 0 and |a / b| < 2^50, without fmod and with
// one division: floor of the EXACT quotient.  Work on |a| like fmod does: t = floor(fl(|a| / b))
// is off by at most one, and the remainder |a| - t * b of the correct t is exactly representable,
// so one fma gives it exactly (for a wrong t the sign / range test below is still decided
// correctly because rounding is monotonic).  jax's (a - mod) / b differs from the resulting
// integer by < 2^-51 * |q|, so its final round() returns it unchanged.
PSB_HD bool fast_floor_divide(double a, double b, double* out) {
  const double A = fabs(a);
  double t = floor(A / b);
  if (!(t < 1125899906842624.0) || !(b > 0.0)) return false;  // also rejects NaN / inf
  double r = fma(-t, b, A);
  if (r < 0.0) t -= 1.0;
  else if (r >= b) t += 1.0;
  r = fma(-t, b, A);  // exact: the remainder fmod(|a|, b)
  *out = (a >= 0.0) ? t : ((r != 0.0) ? -t - 1.0 : -t);
  return true;
}

// kind: 1 regular, 2 periodic, 3 Chebyshev.  Returns the uint32 index (== shape: out of bounds).
PSB_HD_COLD uint32_t grid_index_slow(int kind, double x, double lower, double upper, int32_t shape) {
  const double size = upper - lower;
  const double n = (double)shape;
  double idx;
  if (kind == 3) {  // grids.py:152-156
    double y = 2.0 * (lower - x) / size + 1.0;
    idx = jax_floor_divide(n * acos(y), 3.141592653589793);
    if (isnan(idx)) idx = n;
  } else {
    const double h = size / n;
    idx = jax_floor_divide(x - lower, h);
    if (kind == 2) {  // grids.py:134-138
      idx = jax_remainder(idx, n);
    } else {  // grids.py:117-121
      if (idx < 0.0 || idx > n) idx = n;
    }
    if (isnan(idx)) idx = n;
  }
  return (uint32_t)idx;
}

// Regular / periodic grids take the division-only path; anything unusual (Chebyshev, huge or
// non-finite quotients, non-positive bin width) goes through the literal restatement above.
PSB_HD uint32_t grid_index_1d(int kind, double x, double lower, double upper, int32_t shape) {
  if (kind == 1 || kind == 2) {
    const double h = (upper - lower) / (double)shape;
    double q;
    if (fast_floor_divide(x - lower, h, &q)) {
      if (kind == 1) return (q < 0.0 || q > (double)shape) ? (uint32_t)shape : (uint32_t)q;  // grids.py:117-121
      if (fabs(q) < 2147483648.0) {  // grids.py:134-138: idx % shape (Python sign convention)
        int r = (int)q % shape;
        if (r < 0) r += shape;
        return (uint32_t)r;
      }
    }
  }
  return grid_index_slow(kind, x, lower, upper, shape);
}

// approxfun/core.py:107-136: centre of bin k along one axis
PSB_HD double mesh_coord(int kind, int32_t k, double lower, double upper, int32_t shape) {
  const double n = (double)shape;
  double u;
  if (kind == 3)
    u = -cos(((double)k + 0.5) * 3.141592653589793 / n);
  else
    u = -1.0 + (2.0 * (double)k + 1.0) / n;
  return PSB_ADD(PSB_MUL((u + 1.0) / 2.0, upper - lower), lower);
}

// ---- small dense linear algebra -------------------------------------------------------------
// Cyclic Jacobi eigen-decomposition of a symmetric n x n matrix (n <= 4), row-major in a[16].
// On return a's diagonal holds the eigenvalues and v (row-major) the eigenvectors in columns.
PSB_HD_COLD void jacobi_eig(int n, double* a, double* v) {
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) v[i * 4 + j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 32; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        if (i != j) off += a[i * 4 + j] * a[i * 4 + j];
        else diag += a[i * 4 + j] * a[i * 4 + j];
      }
    if (off <= 1e-60 * diag || off == 0.0) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        double apq = a[p * 4 + q];
        if (apq == 0.0) continue;
        double theta = (a[q * 4 + q] - a[p * 4 + p]) / (2.0 * apq);
        double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; ++k) {
          double akp = a[k * 4 + p], akq = a[k * 4 + q];
          a[k * 4 + p] = c * akp - s * akq;
          a[k * 4 + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {
          double apk = a[p * 4 + k], aqk = a[q * 4 + k];
          a[p * 4 + k] = c * apk - s * aqk;
          a[q * 4 + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          double vkp = v[k * 4 + p], vkq = v[k * 4 + q];
          v[k * 4 + p] = c * vkp - s * vkq;
          v[k * 4 + q] = s * vkp + c * vkq;
        }
      }
  }
}

// orientation.py:33-73 (Kabsch with the reflection fix): the proper rotation U (row-vector
// convention, P @ U ~ Q) maximising tr(U^T C), C = P^T Q (row-major c[9]).  Computed with Horn's
// quaternion eigenproblem (4x4 symmetric, Jacobi), which yields the same U as
// V diag(1,1,sign(det V det W)) W from the SVD whenever the optimum is unique.
PSB_HD_COLD void kabsch_rotation(const double* c, double* U) {
  const double Sxx = c[0], Sxy = c[1], Sxz = c[2];
  const double Syx = c[3], Syy = c[4], Syz = c[5];
  const double Szx = c[6], Szy = c[7], Szz = c[8];
  double N[16] = {Sxx + Syy + Szz, Syz - Szy,       Szx - Sxz,        Sxy - Syx,
                  Syz - Szy,       Sxx - Syy - Szz, Sxy + Syx,        Szx + Sxz,
                  Szx - Sxz,       Sxy + Syx,       -Sxx + Syy - Szz, Syz + Szy,
                  Sxy - Syx,       Szx + Sxz,       Syz + Szy,        -Sxx - Syy + Szz};
  double scale = 0.0;
  for (int i = 0; i < 16; ++i) scale = fmax(scale, fabs(N[i]));
  if (scale > 0.0)
    for (int i = 0; i < 16; ++i) N[i] /= scale;
  double v[16];
  jacobi_eig(4, N, v);
  int best = 0;
  for (int i = 1; i < 4; ++i)
    if (N[i * 4 + i] > N[best * 4 + best]) best = i;
  double q0 = v[0 * 4 + best], q1 = v[1 * 4 + best], q2 = v[2 * 4 + best], q3 = v[3 * 4 + best];
  double nq = sqrt(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
  q0 /= nq; q1 /= nq; q2 /= nq; q3 /= nq;
  // R (column convention, Q_i ~ R P_i); U = R^T
  double R[9] = {q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3, 2 * (q1 * q2 - q0 * q3), 2 * (q1 * q3 + q0 * q2),
                 2 * (q1 * q2 + q0 * q3), q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3, 2 * (q2 * q3 - q0 * q1),
                 2 * (q1 * q3 - q0 * q2), 2 * (q2 * q3 + q0 * q1), q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) U[i * 3 + j] = R[j * 3 + i];
}

// utils/core.py:128-134 `solve_pos_def(A A^T, A B)`: Cholesky solve of the k x k SPD system
// (row-major, leading dimension 4).  Returns false when the matrix is not positive definite.
PSB_HD_COLD bool chol_solve(int k, const double* A, const double* b, double* x) {
  double L[16];
  for (int i = 0; i < k; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = A[i * 4 + j];
      for (int m = 0; m < j; ++m) s -= L[i * 4 + m] * L[j * 4 + m];
      if (i == j) {
        if (!(s > 0.0)) return false;
        L[i * 4 + i] = sqrt(s);
      } else {
        L[i * 4 + j] = s / L[j * 4 + j];
      }
    }
  double y[4];
  for (int i = 0; i < k; ++i) {
    double s = b[i];
    for (int m = 0; m < i; ++m) s -= L[i * 4 + m] * y[m];
    y[i] = s / L[i * 4 + i];
  }
  for (int i = k - 1; i >= 0; --i) {
    double s = y[i];
    for (int m = i + 1; m < k; ++m) s -= L[m * 4 + i] * x[m];
    x[i] = s / L[i * 4 + i];
  }
  return true;
}

// chol_solve for the sizes that occur in practice, straight-line: k = 1 and k = 2 in registers
// (reciprocal square roots instead of sqrt + divisions); larger systems use the loop version.
PSB_HD bool spd_solve(int k, const double* A, const double* b, double* x) {
  if (k == 1) {
    if (!(A[0] > 0.0)) return false;
    x[0] = b[0] / A[0];
    return true;
  }
  if (k == 2) {
    const double a11 = A[0], a21 = A[4], a22 = A[5];
    if (!(a11 > 0.0)) return false;
    const double i11 = 1.0 / sqrt(a11);
    const double l21 = a21 * i11;
    const double s22 = a22 - l21 * l21;
    if (!(s22 > 0.0)) return false;
    const double i22 = 1.0 / sqrt(s22);
    const double y1 = b[0] * i11;
    const double y2 = (b[1] - l21 * y1) * i22;
    const double x2 = y2 * i22;
    x[1] = x2;
    x[0] = (y1 - l21 * x2) * i11;
    return true;
  }
  return chol_solve(k, A, b, x);
}

// utils/core.py:125-126 `pinv(A.T) @ B` == pinv_sym(A A^T) (A B); cutoff follows
// jnp.linalg.pinv's default rtol = 10 * max(M, N) * eps applied to the singular values of A^T
// (eigenvalues of A A^T are their squares).
PSB_HD_COLD void pinv_sym_solve(int k, const double* A, const double* b, double* x, double rtol_sv) {
  double a[16], v[16];
  for (int i = 0; i < 16; ++i) a[i] = 0.0;
  for (int i = 0; i < k; ++i)
    for (int j = 0; j < k; ++j) a[i * 4 + j] = A[i * 4 + j];
  jacobi_eig(k, a, v);
  double lmax = 0.0;
  for (int i = 0; i < k; ++i) lmax = fmax(lmax, a[i * 4 + i]);
  const double cut = rtol_sv * rtol_sv * lmax;
  for (int i = 0; i < k; ++i) x[i] = 0.0;
  for (int e = 0; e < k; ++e) {
    double lam = a[e * 4 + e];
    if (!(lam > cut)) continue;
    double proj = 0.0;
    for (int i = 0; i < k; ++i) proj += v[i * 4 + e] * b[i];
    proj /= lam;
    for (int i = 0; i < k; ++i) x[i] += v[i * 4 + e] * proj;
  }
}

}  // namespace psb
