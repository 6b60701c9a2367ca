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

// ---- colvars/angles.py:158-205: Cremer-Pople ring puckering (m = 2) ---------------------------------
// Coefficients of ring atom t of N: s = sin(2 pi t / N), c = cos(2 pi t / N), S = sin(4 pi t / N),
// C = cos(4 pi t / N).  All four sum to zero over the ring (N >= 3), so the reference's centring
// (rs - barycenter) drops out of every sum below.
PSB_HD void ring_coefficients(uint32_t t, uint32_t N, double* s, double* c, double* S, double* C) {
  const double x = 2.0 * (double)t / (double)N;
#ifdef __CUDA_ARCH__
  sincospi(x, s, c);
  sincospi(2.0 * x, S, C);
#else
  *s = sin(3.141592653589793 * x); *c = cos(3.141592653589793 * x);
  *S = sin(6.283185307179586 * x); *C = cos(6.283185307179586 * x);
#endif
}
// sums[0..11] = R1 = sum s r, R2 = sum c r, A = sum C r, B = sum S r.  Returns q (mode 0) or phi (mode 1)
// and E[0..11]: the gradient with respect to ring atom t is C_t E[0:3] + S_t E[3:6] + s_t E[6:9] + c_t E[9:12].
// With n = R1 x R2, nh = n / |n|, kappa = sqrt(2 / N): a = kappa A.nh, b = -kappa B.nh (angles.py:196-201),
// d(A.nh)/dr_t = C_t nh + s_t R2 x PA + c_t PA x R1 with PA = (A - nh (nh.A)) / |n| (and the same for B).
PSB_HD double ring_cv(int mode, uint32_t N, const double* sums, double* E) {
  const V3 R1 = v3(sums[0], sums[1], sums[2]), R2 = v3(sums[3], sums[4], sums[5]);
  const V3 A = v3(sums[6], sums[7], sums[8]), B = v3(sums[9], sums[10], sums[11]);
  const V3 n = cross(R1, R2);
  const double nn = norm(n);
  const V3 nh = (1.0 / nn) * n;
  const double kappa = sqrt(2.0 / (double)N);
  const double a = kappa * dot(A, nh), b = -kappa * dot(B, nh);
  const V3 PA = (1.0 / nn) * (A - dot(nh, A) * nh), PB = (1.0 / nn) * (B - dot(nh, B) * nh);
  const V3 Ua = cross(R2, PA), Va = cross(PA, R1), Ub = cross(R2, PB), Vb = cross(PB, R1);
  const double q2 = a * a + b * b, q = sqrt(q2);
  // d xi = wa d(a) + wb d(b);  d(a) = kappa [...A], d(b) = -kappa [...B]
  const double wa = (mode == 0 ? a / q : -b / q2) * kappa, wb = -(mode == 0 ? b / q : a / q2) * kappa;
  const V3 e1 = wa * nh, e2 = wb * nh, e3 = wa * Ua + wb * Ub, e4 = wa * Va + wb * Vb;
  E[0] = e1.x; E[1] = e1.y; E[2] = e1.z;
  E[3] = e2.x; E[4] = e2.y; E[5] = e2.z;
  E[6] = e3.x; E[7] = e3.y; E[8] = e3.z;
  E[9] = e4.x; E[10] = e4.y; E[11] = e4.z;
  return mode == 0 ? q : atan2(b, a);
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

// The same value as jax_floor_divide(a, b) for b > 0 and |a / b| < 2^50, without fmod and without a
// division on the critical path: floor of the EXACT quotient.  Work on |a| like fmod does:
// t = floor(fl(|a| * fl(1 / b))) is off by at most one (two roundings: < 2^-52 |a / b| < 1/4), and the
// remainder |a| - t * b of the correct t is exactly representable, so one fma gives it exactly (for a
// wrong t the sign / range test below is still decided correctly because rounding is monotonic).
// jax's (a - mod) / b differs from the resulting integer by < 2^-51 * |q|, so its final round()
// returns it unchanged.  inv_b must be fl(1 / b).
PSB_HD bool fast_floor_divide(double a, double b, double inv_b, double* out) {
  const double A = fabs(a);
  double t = floor(A * inv_b);
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

// Regular / periodic grids take the fma-corrected path; anything unusual (Chebyshev, huge or
// non-finite quotients, non-positive bin width) goes through the literal restatement above.
// h = (upper - lower) / shape and inv_h = 1 / h are per-axis constants (the step kernel gets them
// from psb_create, computed with the same two IEEE divisions).
PSB_HD uint32_t grid_index_1d(int kind, double x, double lower, double upper, int32_t shape, double h, double inv_h) {
  if (kind == 1 || kind == 2) {
    double q;
    if (fast_floor_divide(x - lower, h, inv_h, &q)) {
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
PSB_HD uint32_t grid_index_1d(int kind, double x, double lower, double upper, int32_t shape) {
  const double h = (upper - lower) / (double)shape;
  return grid_index_1d(kind, x, lower, upper, shape, h, 1.0 / h);
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
    if (off <= 1e-36 * diag || off == 0.0) break;  // |off|_F <= 1e-18 |diag|_F: below fp64 resolution
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

// Gyration-tensor CVs (colvars/shape.py:85-344).  g = (xx, xy, xz, yy, yz, zz) of sum_i r_i r_i^T / n.
// Returns the CV value f(l) for `mode` (psb_gyration_mode; eigenvalues l ascending like
// jnp.linalg.eigvalsh) and W = sum_k (df/dl_k) v_k v_k^T (row-major 3x3): the gradient with respect
// to atom i is (2 / n) W r_i (first-order perturbation of a simple eigenvalue).
PSB_HD_COLD double gyration_cv(int mode, int ja, int jb, const double* g, double* W) {
  double a[16], v[16];
  for (int i = 0; i < 16; ++i) a[i] = 0.0;
  a[0] = g[0]; a[1] = g[1]; a[2] = g[2];
  a[4] = g[1]; a[5] = g[3]; a[6] = g[4];
  a[8] = g[2]; a[9] = g[4]; a[10] = g[5];
  jacobi_eig(3, a, v);
  int o[3] = {0, 1, 2};  // ascending order of the diagonal
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2 - i; ++j)
      if (a[o[j] * 4 + o[j]] > a[o[j + 1] * 4 + o[j + 1]]) { int t = o[j]; o[j] = o[j + 1]; o[j + 1] = t; }
  const double l0 = a[o[0] * 4 + o[0]], l1 = a[o[1] * 4 + o[1]], l2 = a[o[2] * 4 + o[2]];
  const double l[3] = {l0, l1, l2};
  double w[3] = {0.0, 0.0, 0.0}, xi;
  if (mode <= 2) {
    xi = l[mode];
    w[mode] = 1.0;
  } else if (mode == 3) {
    xi = l2 - (l0 + l1) / 2;
    w[0] = -0.5; w[1] = -0.5; w[2] = 1.0;
  } else if (mode == 4) {
    xi = l[jb] - l[ja];
    w[jb] += 1.0; w[ja] -= 1.0;
  } else {
    const double s1 = l0 + l1 + l2, s2 = l0 * l0 + l1 * l1 + l2 * l2;
    xi = (3 * s2 / (s1 * s1) - 1) / 2;
    for (int k = 0; k < 3; ++k) w[k] = 3.0 * l[k] / (s1 * s1) - 3.0 * s2 / (s1 * s1 * s1);
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0.0;
      for (int k = 0; k < 3; ++k) s += w[k] * v[i * 4 + o[k]] * v[j * 4 + o[k]];
      W[i * 3 + j] = s;
    }
  return xi;
}

// orientation.py:33-73 (Kabsch with the reflection fix): the proper rotation U (row-vector
// convention, P @ U ~ Q) maximising tr(U^T C), C = P^T Q (row-major c[9]).  General path: Horn's
// quaternion eigenproblem (4x4 symmetric, Jacobi), which yields the same U as
// V diag(1,1,sign(det V det W)) W from the SVD whenever the optimum is unique.
PSB_HD_COLD void kabsch_rotation_eig(const double* c, double* U) {
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

// Fast path: the orthogonal polar factor X of C (C = X H, H symmetric positive semi-definite,
// X = V W from the SVD) by the scaled Newton iteration X <- (g X + X^-T / g) / 2,
// g = sqrt(|X^-1|_F / |X|_F) (Higham): quadratic convergence, ~6 iterations of straight-line 3x3
// arithmetic in registers instead of several Jacobi sweeps over a 4x4 matrix in local memory (the
// finalizer runs this in ONE thread per CTA, on the critical path of every step).
//   det C > 0: U = X.
//   det C < 0: the reflection fix of orientation.py:66-71 flips the direction of the smallest
//              singular value: U = V diag(1,1,-1) W = X (I - 2 w w^T), w = eigenvector of
//              H = X^T C for its smallest eigenvalue (closed-form symmetric 3x3 eigenvalue,
//              eigenvector from the best-conditioned cross product of rows of H - lambda I).
// Rank-deficient / badly conditioned covariances, a (near-)degenerate smallest singular value or a
// non-converging iteration take the eigen path above.
PSB_HD void kabsch_rotation(const double* c, double* U) {
  double x0 = c[0], x1 = c[1], x2 = c[2], x3 = c[3], x4 = c[4], x5 = c[5], x6 = c[6], x7 = c[7], x8 = c[8];
  const double fro2 = x0 * x0 + x1 * x1 + x2 * x2 + x3 * x3 + x4 * x4 + x5 * x5 + x6 * x6 + x7 * x7 + x8 * x8;
  const double det0 = x0 * (x4 * x8 - x5 * x7) - x1 * (x3 * x8 - x5 * x6) + x2 * (x3 * x7 - x4 * x6);
  bool ok = fabs(det0) > 1e-6 * fro2 * sqrt(fro2) && fro2 > 0.0;  // sigma_min / sigma_max >~ 1e-6
  bool converged = false;
  bool scaled = true;  // Higham's scaling only while X is far from orthogonal: near it g -> 1 and its sqrt(sqrt(a / b))
                       // would be most of the iteration's dependent chain (this runs in one thread, every step)
  for (int it = 0; ok && it < 16; ++it) {
    // adjugate^T / det = X^-T
    const double a0 = x4 * x8 - x5 * x7, a1 = x5 * x6 - x3 * x8, a2 = x3 * x7 - x4 * x6;
    const double a3 = x2 * x7 - x1 * x8, a4 = x0 * x8 - x2 * x6, a5 = x1 * x6 - x0 * x7;
    const double a6 = x1 * x5 - x2 * x4, a7 = x2 * x3 - x0 * x5, a8 = x0 * x4 - x1 * x3;
    const double det = x0 * a0 + x1 * a1 + x2 * a2;
    if (det == 0.0 || (det > 0.0) != (det0 > 0.0)) break;
    const double id = 1.0 / det;
    double p = 0.5, q = 0.5 * id;
    if (scaled) {
      const double nx = x0 * x0 + x1 * x1 + x2 * x2 + x3 * x3 + x4 * x4 + x5 * x5 + x6 * x6 + x7 * x7 + x8 * x8;
      const double na = (a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3 + a4 * a4 + a5 * a5 + a6 * a6 + a7 * a7 + a8 * a8) * id * id;
      const double g = sqrt(sqrt(na / nx));  // sqrt(|X^-1|_F / |X|_F)
      p = 0.5 * g;
      q = 0.5 * id / g;
    }
    const double y0 = p * x0 + q * a0, y1 = p * x1 + q * a1, y2 = p * x2 + q * a2;
    const double y3 = p * x3 + q * a3, y4 = p * x4 + q * a4, y5 = p * x5 + q * a5;
    const double y6 = p * x6 + q * a6, y7 = p * x7 + q * a7, y8 = p * x8 + q * a8;
    const double d2 = (y0 - x0) * (y0 - x0) + (y1 - x1) * (y1 - x1) + (y2 - x2) * (y2 - x2) + (y3 - x3) * (y3 - x3) +
                      (y4 - x4) * (y4 - x4) + (y5 - x5) * (y5 - x5) + (y6 - x6) * (y6 - x6) + (y7 - x7) * (y7 - x7) +
                      (y8 - x8) * (y8 - x8);
    x0 = y0; x1 = y1; x2 = y2; x3 = y3; x4 = y4; x5 = y5; x6 = y6; x7 = y7; x8 = y8;
    if (d2 <= 1e-2) scaled = false;  // |X_k+1 - X_k|_F <= 0.1: the quadratic regime, plain Newton from here
    if (d2 <= 1e-30) {  // |X_k+1 - X_k|_F <= 1e-15 with |X|_F = sqrt(3): converged (quadratically)
      converged = true;
      break;
    }
  }
  if (converged && det0 > 0.0) {
    U[0] = x0; U[1] = x1; U[2] = x2; U[3] = x3; U[4] = x4; U[5] = x5; U[6] = x6; U[7] = x7; U[8] = x8;
    return;
  }
  if (converged) {  // improper polar factor: reflect the direction of the smallest singular value
    // H = X^T C (symmetric)
    const double h00 = x0 * c[0] + x3 * c[3] + x6 * c[6], h01 = x0 * c[1] + x3 * c[4] + x6 * c[7];
    const double h02 = x0 * c[2] + x3 * c[5] + x6 * c[8], h11 = x1 * c[1] + x4 * c[4] + x7 * c[7];
    const double h12 = x1 * c[2] + x4 * c[5] + x7 * c[8], h22 = x2 * c[2] + x5 * c[5] + x8 * c[8];
    const double p1 = h01 * h01 + h02 * h02 + h12 * h12;
    const double qm = (h00 + h11 + h22) / 3.0;
    const double p2 = (h00 - qm) * (h00 - qm) + (h11 - qm) * (h11 - qm) + (h22 - qm) * (h22 - qm) + 2.0 * p1;
    const double pp = sqrt(p2 / 6.0);
    if (pp > 0.0) {
      const double b00 = (h00 - qm) / pp, b11 = (h11 - qm) / pp, b22 = (h22 - qm) / pp;
      const double b01 = h01 / pp, b02 = h02 / pp, b12 = h12 / pp;
      double r = 0.5 * (b00 * (b11 * b22 - b12 * b12) - b01 * (b01 * b22 - b12 * b02) + b02 * (b01 * b12 - b11 * b02));
      r = fmin(1.0, fmax(-1.0, r));
      const double phi = acos(r) / 3.0;
      const double lam = qm + 2.0 * pp * cos(phi + 2.0943951023931953);  // smallest eigenvalue
      // rows of H - lam I; the eigenvector is orthogonal to all of them
      const double r00 = h00 - lam, r11 = h11 - lam, r22 = h22 - lam;
      const double ax = h01 * h12 - h02 * r11, ay = h02 * h01 - r00 * h12, az = r00 * r11 - h01 * h01;   // row0 x row1
      const double bx = h01 * r22 - h02 * h12, by = h02 * h02 - r00 * r22, bz = r00 * h12 - h01 * h02;   // row0 x row2
      const double cx = r11 * r22 - h12 * h12, cy = h12 * h02 - h01 * r22, cz = h01 * h12 - r11 * h02;   // row1 x row2
      const double na2 = ax * ax + ay * ay + az * az, nb2 = bx * bx + by * by + bz * bz, nc2 = cx * cx + cy * cy + cz * cz;
      double wx = ax, wy = ay, wz = az, nw = na2;
      if (nb2 > nw) { wx = bx; wy = by; wz = bz; nw = nb2; }
      if (nc2 > nw) { wx = cx; wy = cy; wz = cz; nw = nc2; }
      const double hs = h00 * h00 + h11 * h11 + h22 * h22 + 2.0 * p1;  // |H|_F^2; cross products scale like |H|^2
      if (nw > 1e-12 * hs * hs) {  // smallest singular value well separated from the middle one
        const double inv = 1.0 / sqrt(nw);
        wx *= inv; wy *= inv; wz *= inv;
        const double t0 = 2.0 * (x0 * wx + x1 * wy + x2 * wz), t1 = 2.0 * (x3 * wx + x4 * wy + x5 * wz),
                     t2 = 2.0 * (x6 * wx + x7 * wy + x8 * wz);  // 2 X w
        U[0] = x0 - t0 * wx; U[1] = x1 - t0 * wy; U[2] = x2 - t0 * wz;
        U[3] = x3 - t1 * wx; U[4] = x4 - t1 * wy; U[5] = x5 - t1 * wz;
        U[6] = x6 - t2 * wx; U[7] = x7 - t2 * wy; U[8] = x8 - t2 * wz;
        return;
      }
    }
  }
  kabsch_rotation_eig(c, U);
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
