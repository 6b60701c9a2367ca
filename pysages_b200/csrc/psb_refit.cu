// psb_refit.cu -- the cadenced Levenberg-Marquardt refit of the ABF family's networks on the device (sm_100a).
//
// FUNN (funn.py:213-249), ANN (ann.py:194-211) and CFF's force network (cff.py:300-307) refit a small tanh
// multilayer perceptron every `train_freq` steps with PySAGES' Levenberg-Marquardt (ml/optimizers.py:188-232) on a
// sum-of-squared-errors loss with float32 residuals and L2 regularisation (ml/objectives.py).  One LM iteration is
//
//   lm_jacobian  J, e = d residuals / d params, residuals   closed-form back-propagation, one warp per sample
//   lm_normal    H = J^T J, g = J^T e                       fp64 syrk, 128 x 128 tiles, 8 x 8 per thread, split over samples
//   lm_solve     dp = solve(H + (r + mu) I, g + r p)        cooperative blocked Cholesky over all SMs (see there)
//   lm_cost      C = (|e(p - dp)|^2 + r |p - dp|^2) / 2     forward pass; the last CTA to finish takes the loop's
//                rho, mu update, accept / reject            decision (float32 mu like the reference)
//
// and the loop's control flow is data dependent (accept / reject, convergence).  All of it lives in device memory:
// the host enqueues the iterations blindly (one CUDA graph of the four kernels, replayed) and every kernel returns
// at once when the `done` flag is set, so a refit never synchronises with the host and is stream-ordered between
// two bias steps.  A rejected step keeps p, so J and H are reused (`fresh == 0` skips the first two kernels).
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "../../include/pysages_b200.h"

namespace psb_refit {

constexpr int MAXW = PSB_MAX_WIDTH;       // units per layer
constexpr int MAXL = PSB_MAX_DENSE;       // dense layers
constexpr int NB = 32;                    // Cholesky block
constexpr int MAX_SPLIT = 32;             // sample splits of the normal-equation kernel (summed in fixed order)
constexpr int TN = 128, TK = 16;          // lm_normal: output tile, rows of J per stage
constexpr int TN64 = 64, TK64 = 32;       // lm_normal64
constexpr int NORMAL64_MAX_P = 384;       // parameter counts up to which the 64 x 64 tiles are used

struct Net {
  int32_t n_dense, nparams;
  int32_t widths[MAXL + 1];
  int32_t w_off[MAXL];  // offset of W_l in the flat vector, b_l follows (ml/utils.py:43-53)
  double lower[PSB_MAX_K], upper[PSB_MAX_K];
};

// Loop state (ml/optimizers.py:188-232 as restated in pysages_b200/ml.py:fit)
struct Control {
  double C_prev;      // cost of the accepted parameters (inf before the first iteration)
  double C_try;
  double denom;       // dp . (mu dp + g)
  double ssq_try;     // |e_try|^2
  float mu;
  int32_t iters, max_iters;
  int32_t fresh;      // J / H must be rebuilt (the parameters changed)
  int32_t done;
  int32_t chol_failed;
  int32_t total_loops;
  float c, mu_min, mu_max, rho_c, rho_min;
  double reg;
  unsigned int bar;         // grid barrier of lm_solve (zeroed by lm_cost's last CTA)
  unsigned int cost_count;  // CTAs of lm_cost that have finished
};

struct Problem {
  Net net;
  long long n;        // samples
  int32_t nout;       // outputs per sample
  int32_t T;          // 32 x 32 blocks per side of the padded normal matrix
  int32_t splits;     // sample splits of lm_normal
  const double* x;    // (n, in)
  const double* y;    // (n, out)
  double* p;          // (P) accepted parameters
  double* p_try;      // (P)
  double* e;          // (n * out) residuals of p
  double* J;          // (n * out, P)
  double* Hpart;      // (splits, P, P)
  double* gpart;      // (splits, P)
  double* H;          // (P, P) J^T J
  double* gsum;       // (P) J^T e
  double* A;          // (32 T, 32 T) working copy of H + shift I, identity in the padding
  double* Lm;         // (32 T, 32 T) Cholesky factor, lower triangle by blocks
  double* Linv;       // (T, 32, 32) inverses of the factor's diagonal blocks
  double* gz;         // (32 T) right-hand side, updated in place by the factorisation
  double* zv;         // (32 T) L z = g
  double* g;          // (P) J^T e + r p
  double* dp;         // (P)
  double* ssq_part;   // per-CTA partial sums of the cost kernel
  int32_t ssq_parts;
  Control* ctl;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// forward pass of one sample by one warp; activations of every layer stay in `act` ((MAXL + 1) x MAXW per warp)
__device__ __forceinline__ void forward(const Net& net, const double* __restrict__ P, const double* __restrict__ xs,
                                        double* act, int lane) {
  const int in0 = net.widths[0];
  if (lane < in0) act[lane] = (xs[lane] - net.lower[lane]) * 2.0 / (net.upper[lane] - net.lower[lane]) - 1.0;  // approxfun/core.py:99-104
  __syncwarp();
  for (int l = 0; l < net.n_dense; ++l) {
    const int in = net.widths[l], out = net.widths[l + 1];
    const double* W = P + net.w_off[l];
    const double* b = W + (size_t)in * out;
    const double* cur = act + (size_t)l * MAXW;
    double* nxt = act + (size_t)(l + 1) * MAXW;
    for (int j = lane; j < out; j += 32) {
      double acc = 0.0;
      for (int i = 0; i < in; ++i) acc = fma(cur[i], W[(size_t)i * out + j], acc);
      acc += b[j];
      nxt[j] = (l + 1 < net.n_dense) ? tanh(acc) : acc;  // ml/models.py:44-82: activation between the dense layers
    }
    __syncwarp();
  }
}

// ---- K1: residuals and Jacobian rows of the accepted parameters -----------------------------------------------
__global__ void __launch_bounds__(256) lm_jacobian(const Problem pr) {
  if (pr.ctl->done || !pr.ctl->fresh) return;
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  double* act = smem + (size_t)warp * ((MAXL + 1) * MAXW + 2 * MAXW);
  double* delta = act + (MAXL + 1) * MAXW;
  double* dnext = delta + MAXW;
  const Net& net = pr.net;
  const int nd = net.n_dense, nout = pr.nout, P = net.nparams;
  for (long long s = (long long)blockIdx.x * nwarps + warp; s < pr.n; s += (long long)gridDim.x * nwarps) {
    forward(net, pr.p, pr.x + s * net.widths[0], act, lane);
    const double* outv = act + (size_t)nd * MAXW;
    for (int o = 0; o < nout; ++o) {
      const long long row = s * nout + o;
      if (lane == 0) pr.e[row] = (double)(float)(outv[o] - pr.y[row]);  // ml/objectives.py: float32 residuals
      double* Jr = pr.J + (size_t)row * P;
      for (int j = lane; j < net.widths[nd]; j += 32) delta[j] = (j == o) ? 1.0 : 0.0;
      __syncwarp();
      for (int l = nd - 1; l >= 0; --l) {
        const int in = net.widths[l], out = net.widths[l + 1];
        const double* a_in = act + (size_t)l * MAXW;
        const double* W = pr.p + net.w_off[l];
        double* JW = Jr + net.w_off[l];
        for (int q = lane; q < in * out; q += 32) JW[q] = a_in[q / out] * delta[q % out];  // d out_o / d W_l[i][j] = a_i delta_j
        for (int j = lane; j < out; j += 32) JW[(size_t)in * out + j] = delta[j];       // d out_o / d b_l[j]
        if (l > 0) {
          for (int i = lane; i < in; i += 32) {
            double acc = 0.0;
            for (int j = 0; j < out; ++j) acc = fma(W[(size_t)i * out + j], delta[j], acc);
            dnext[i] = acc * (1.0 - a_in[i] * a_in[i]);  // through tanh of layer l - 1
          }
          __syncwarp();
          for (int i = lane; i < in; i += 32) delta[i] = dnext[i];
        }
        __syncwarp();
      }
    }
  }
}

// ---- K2: H = J^T J (full, symmetric), g = J^T e; `splits` partial results over disjoint sample ranges -------------
// 128 x 128 output tiles of the upper triangle, 256 threads with 8 x 8 fp64 accumulators each: 64 DFMA per eight
// 16-byte shared-memory loads (a 4 x 4 register tile is bound by shared-memory bandwidth: ncu 62 % short-scoreboard
// stalls, FP64 pipe 28 %).  A thread's rows are contiguous (broadcast loads), its columns are pairs strided by 32 so
// that the 16 lanes of a half-warp read 256 contiguous bytes: no bank conflicts.  The next 16 rows of J travel through
// registers while the current ones are multiplied.
__global__ void __launch_bounds__(256, 1) lm_normal(const Problem pr) {
  if (pr.ctl->done || !pr.ctl->fresh) return;
  __shared__ __align__(16) double sa[TK][TN], sb[TK][TN];
  __shared__ double se[TK];
  const int P = pr.net.nparams;
  const int tiles = (P + TN - 1) / TN;
  const long long M = pr.n * pr.nout;
  const int ntri = tiles * (tiles + 1) / 2;
  const int tri = blockIdx.x % ntri, split = blockIdx.x / ntri;
  int ti = 0, rem = tri;
  while (rem >= tiles - ti) { rem -= tiles - ti; ++ti; }
  const int tj = ti + rem;
  long long per = (M + pr.splits - 1) / pr.splits;
  per = (per + TK - 1) / TK * TK;
  const long long m0 = min(M, split * per), m1 = min(M, m0 + per);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  // loader: 16 rows x 128 columns per matrix and stage = 8 elements per thread and matrix, coalesced along the parameters
  const int lr = tid >> 4, lc = (tid & 15) * 8;
  double ra[8], rb[8], re = 0.0;
  auto fetch = [&](long long m) {
    const long long row = m + lr;
    const bool ok = row < m1;
    const double* Jr = pr.J + (size_t)row * P;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int ca = ti * TN + lc + q, cb = tj * TN + lc + q;
      ra[q] = (ok && ca < P) ? Jr[ca] : 0.0;
      rb[q] = (ok && cb < P) ? Jr[cb] : 0.0;
    }
    if (tid < TK) re = (m + tid < m1) ? pr.e[m + tid] : 0.0;
  };
  double acc[8][8] = {}, gacc = 0.0;
  if (m0 < m1) fetch(m0);
  for (long long m = m0; m < m1; m += TK) {
#pragma unroll
    for (int q = 0; q < 8; ++q) { sa[lr][lc + q] = ra[q]; sb[lr][lc + q] = rb[q]; }
    if (tid < TK) se[tid] = re;
    __syncthreads();
    if (m + TK < m1) fetch(m + TK);
#pragma unroll 1
    for (int r = 0; r < TK; ++r) {
      double a[8], b[8];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const double2 av = *reinterpret_cast<const double2*>(&sa[r][ty * 8 + 2 * q]);
        const double2 bv = *reinterpret_cast<const double2*>(&sb[r][tx * 2 + 32 * q]);
        a[2 * q] = av.x; a[2 * q + 1] = av.y;
        b[2 * q] = bv.x; b[2 * q + 1] = bv.y;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    if (ti == tj && tid < TN) {
#pragma unroll
      for (int r = 0; r < TK; ++r) gacc = fma(sa[r][tid], se[r], gacc);
    }
    __syncthreads();
  }
  double* H = pr.Hpart + (size_t)split * P * P;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int gi = ti * TN + ty * 8 + i, gj = tj * TN + tx * 2 + 32 * (j >> 1) + (j & 1);
      if (gi < P && gj < P) {
        H[(size_t)gi * P + gj] = acc[i][j];
        if (ti != tj) H[(size_t)gj * P + gi] = acc[i][j];
      }
    }
  if (ti == tj && tid < TN && ti * TN + tid < P) pr.gpart[(size_t)split * P + ti * TN + tid] = gacc;
}

// ---- K2 (up to 384 parameters): the same with 64 x 64 tiles (less padding, more CTAs for small matrices) ------------------
// 64 x 64 output tiles of the upper triangle, 256 threads with 4 x 4 accumulators each (16 DFMA per 4 shared-memory
// loads of 16 bytes); the next 32 rows of J travel through registers while the current ones are multiplied.
__global__ void __launch_bounds__(256) lm_normal64(const Problem pr) {
  if (pr.ctl->done || !pr.ctl->fresh) return;
  __shared__ __align__(16) double sa[TK64][TN64], sb[TK64][TN64];
  __shared__ double se[TK64];
  const int P = pr.net.nparams;
  const int tiles = (P + TN64 - 1) / TN64;
  const long long M = pr.n * pr.nout;
  const int ntri = tiles * (tiles + 1) / 2;
  const int tri = blockIdx.x % ntri, split = blockIdx.x / ntri;
  int ti = 0, rem = tri;
  while (rem >= tiles - ti) { rem -= tiles - ti; ++ti; }
  const int tj = ti + rem;
  long long per = (M + pr.splits - 1) / pr.splits;
  per = (per + TK64 - 1) / TK64 * TK64;
  const long long m0 = min(M, split * per), m1 = min(M, m0 + per);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  // loader: 32 rows x 64 columns per matrix and stage = 2 x 4 elements per thread, coalesced along the parameters
  const int lr = tid >> 4, lc = (tid & 15) * 4;
  double ra[2][4], rb[2][4], re = 0.0;
  auto fetch = [&](long long m) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const long long row = m + lr + 16 * h;
      const bool ok = row < m1;
      const double* Jr = pr.J + (size_t)row * P;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int ca = ti * TN64 + lc + q, cb = tj * TN64 + lc + q;
        ra[h][q] = (ok && ca < P) ? Jr[ca] : 0.0;
        rb[h][q] = (ok && cb < P) ? Jr[cb] : 0.0;
      }
    }
    if (tid < TK64) re = (m + tid < m1) ? pr.e[m + tid] : 0.0;
  };
  double acc[4][4] = {}, gacc = 0.0;
  if (m0 < m1) fetch(m0);
  for (long long m = m0; m < m1; m += TK64) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      sa[lr][lc + q] = ra[0][q]; sb[lr][lc + q] = rb[0][q];
      sa[lr + 16][lc + q] = ra[1][q]; sb[lr + 16][lc + q] = rb[1][q];
    }
    if (tid < TK64) se[tid] = re;
    __syncthreads();
    if (m + TK64 < m1) fetch(m + TK64);
#pragma unroll
    for (int r = 0; r < TK64; ++r) {
      const double2 a01 = *reinterpret_cast<const double2*>(&sa[r][ty * 4]), a23 = *reinterpret_cast<const double2*>(&sa[r][ty * 4 + 2]);
      const double2 b01 = *reinterpret_cast<const double2*>(&sb[r][tx * 4]), b23 = *reinterpret_cast<const double2*>(&sb[r][tx * 4 + 2]);
      const double a[4] = {a01.x, a01.y, a23.x, a23.y}, b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    if (ti == tj && tid < TN64) {
#pragma unroll
      for (int r = 0; r < TK64; ++r) gacc = fma(sa[r][tid], se[r], gacc);
    }
    __syncthreads();
  }
  double* H = pr.Hpart + (size_t)split * P * P;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gi = ti * TN64 + ty * 4 + i, gj = tj * TN64 + tx * 4 + j;
      if (gi < P && gj < P) {
        H[(size_t)gi * P + gj] = acc[i][j];
        if (ti != tj) H[(size_t)gj * P + gi] = acc[i][j];
      }
    }
  if (ti == tj && tid < TN64 && ti * TN64 + tid < P) pr.gpart[(size_t)split * P + ti * TN64 + tid] = gacc;
}

// ---- K2b: the partial results added in a fixed order ------------------------------------------------------------------
__global__ void __launch_bounds__(256) lm_reduce(const Problem pr) {
  if (pr.ctl->done || !pr.ctl->fresh) return;
  const int P = pr.net.nparams;
  const size_t PP = (size_t)P * P, q = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (q < PP) {
    double v = 0.0;
    for (int s = 0; s < pr.splits; ++s) v += pr.Hpart[(size_t)s * PP + q];
    pr.H[q] = v;
  }
  if (q < (size_t)P) {
    double v = 0.0;
    for (int s = 0; s < pr.splits; ++s) v += pr.gpart[(size_t)s * P + q];
    pr.gsum[q] = v;
  }
}

// ---- K3: dp = solve(H + (r + mu) I, g), p_try = p - dp, denom = dp . (mu dp + g) ---------------------------------------
// Cooperative launch, one CTA per SM at most.  The matrix (padded to T x T blocks of 32 x 32, identity in the padding)
// stays in global memory (L2 resident: <= 8 MB).  Right-looking blocked Cholesky with ONE grid barrier per block column:
//   every CTA factors the diagonal block A[k,k] = L11 L11^T itself (one warp, rows in registers) and inverts L11;
//   the trailing tiles (I, J), k < J <= I, are dealt round-robin: the CTA forms L_I = A[I,k] L11^-T and L_J = A[J,k] L11^-T
//   (32^3 products against the inverse, no substitutions) and updates A[I,J] -= L_I L_J^T.  The panel column is never
//   overwritten (the factor goes to `Lm`), so no barrier separates the panel from the update;
//   the right-hand side rides along as one more row of the panel: z_k = L11^-1 g_k, g_I -= L_I z_k -- the forward
//   substitution is done when the factorisation is.
// CTA 0 finishes with the backward substitution L^T dp = z by blocks.
struct SolveShared {
  double L11[NB][NB + 1], Li[NB][NB + 1], PA[NB][NB + 1], PB[NB][NB + 1], LI[NB][NB + 1], LJ[NB][NB + 1];
  double zk[NB], gk[NB];
  double part[8][NB];
  double red[8];
  int fail;
};

__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
    } while (v < target);
    __threadfence();
  }
  __syncthreads();
}

// 32 x 32 product against a transposed block: out[r][c] = sum_t X[r][t] Y[c][t], rows r = warp + 8 q, column c = lane
__device__ __forceinline__ void block_abt(const double (*X)[NB + 1], const double (*Y)[NB + 1], double out[4], int lane, int warp) {
#pragma unroll
  for (int q = 0; q < 4; ++q) out[q] = 0.0;
#pragma unroll 8
  for (int t = 0; t < NB; ++t) {
    const double y = Y[lane][t];
#pragma unroll
    for (int q = 0; q < 4; ++q) out[q] = fma(X[warp + 8 * q][t], y, out[q]);
  }
}

__global__ void __launch_bounds__(256) lm_solve(const Problem pr) {
  Control& ctl = *pr.ctl;
  if (ctl.done) return;
  extern __shared__ __align__(16) unsigned char solve_smem[];
  SolveShared& sh = *reinterpret_cast<SolveShared*>(solve_smem);
  const int P = pr.net.nparams, T = pr.T, Pp = T * NB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x;
  double* A = pr.A;
  // ml.py: H.diagonal().add_(float(f32(r) + mu)); Je = J^T e + r p
  const double shift = (double)((float)ctl.reg + ctl.mu);
  const double reg = ctl.reg;
  for (size_t q = (size_t)blockIdx.x * 256 + tid; q < (size_t)Pp * Pp; q += (size_t)G * 256) {
    const int i = (int)(q / Pp), j = (int)(q % Pp);
    double v = (i == j) ? 1.0 : 0.0;
    if (i < P && j < P) v = pr.H[(size_t)i * P + j] + ((i == j) ? shift : 0.0);
    A[q] = v;
  }
  for (int i = blockIdx.x * 256 + tid; i < Pp; i += G * 256) {
    double v = 0.0;
    if (i < P) {
      v = pr.gsum[i] + reg * pr.p[i];
      pr.g[i] = v;
    }
    pr.gz[i] = v;
  }
  if (tid == 0) sh.fail = 0;
  unsigned int epoch = 0;
  grid_barrier(&ctl.bar, (epoch += G));

  for (int k = 0; k < T; ++k) {
    for (int q = tid; q < NB * NB; q += 256) sh.L11[q >> 5][q & 31] = A[(size_t)(k * NB + (q >> 5)) * Pp + k * NB + (q & 31)];
    __syncthreads();
    if (warp == 0) {
      // lane i owns row i of the block; 1 / sqrt(d) by Newton on the hardware estimate keeps square roots and
      // divisions out of the 32-step dependency chain
      double a[NB], dinv[NB];
#pragma unroll
      for (int j = 0; j < NB; ++j) a[j] = sh.L11[lane][j];
      bool fail = false;
#pragma unroll
      for (int c = 0; c < NB; ++c) {
        double d = __shfl_sync(0xffffffffu, a[c], c);
        if (!(d > 0.0) || !(d < 1.79e308)) { fail = true; d = 1.0; }
        double ri = rsqrt(d);
        ri = fma(ri * 0.5, fma(-d * ri, ri, 1.0), ri);  // one more Newton step: full double precision
        dinv[c] = ri;
        a[c] = (lane == c) ? d * ri : a[c] * ri;
#pragma unroll
        for (int j = c + 1; j < NB; ++j) {
          const double ljc = __shfl_sync(0xffffffffu, a[c], j);
          a[j] = fma(-a[c], ljc, a[j]);
        }
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < NB; ++j) sh.L11[lane][j] = (j <= lane) ? a[j] : 0.0;
      if (lane == 0 && fail) sh.fail = 1;
      __syncwarp();
      // lane c owns column c of the inverse: forward substitution against the unit vector
      double x[NB];
#pragma unroll
      for (int i = 0; i < NB; ++i) {
        double acc0 = (i == lane) ? 1.0 : 0.0, acc1 = 0.0;
#pragma unroll
        for (int t = 0; t + 1 < i; t += 2) {
          acc0 = fma(-sh.L11[i][t], x[t], acc0);
          acc1 = fma(-sh.L11[i][t + 1], x[t + 1], acc1);
        }
        if (i & 1) acc0 = fma(-sh.L11[i][i - 1], x[i - 1], acc0);
        x[i] = (acc0 + acc1) * dinv[i];
      }
#pragma unroll
      for (int i = 0; i < NB; ++i) sh.Li[i][lane] = (i >= lane) ? x[i] : 0.0;
    } else if (warp == 1) {
      sh.gk[lane] = pr.gz[k * NB + lane];
    }
    __syncthreads();
    if (sh.fail) break;  // every CTA computed the same block: the decision is grid-uniform
    if (tid < NB) {
      double v = 0.0;
      for (int t = 0; t <= tid; ++t) v = fma(sh.Li[tid][t], sh.gk[t], v);
      sh.zk[tid] = v;
    }
    __syncthreads();
    if (blockIdx.x == 0) {
      for (int q = tid; q < NB * NB; q += 256) {
        pr.Lm[(size_t)(k * NB + (q >> 5)) * Pp + k * NB + (q & 31)] = sh.L11[q >> 5][q & 31];
        pr.Linv[(size_t)k * NB * NB + q] = sh.Li[q >> 5][q & 31];
      }
      if (tid < NB) pr.zv[k * NB + tid] = sh.zk[tid];
    }
    const int rest = T - k - 1;
    const int ntile = rest * (rest + 1) / 2;
    for (int t = blockIdx.x; t < ntile; t += G) {
      // column-major over the trailing lower triangle: the first `rest` tiles are column k + 1
      int J = k + 1, r = t;
      while (r >= T - J) { r -= T - J; ++J; }
      const int I = J + r;
      for (int q = tid; q < NB * NB; q += 256) {
        sh.PA[q >> 5][q & 31] = A[(size_t)(I * NB + (q >> 5)) * Pp + k * NB + (q & 31)];
        if (I != J) sh.PB[q >> 5][q & 31] = A[(size_t)(J * NB + (q >> 5)) * Pp + k * NB + (q & 31)];
      }
      __syncthreads();
      double o[4];
      block_abt(sh.PA, sh.Li, o, lane, warp);
#pragma unroll
      for (int q = 0; q < 4; ++q) sh.LI[warp + 8 * q][lane] = o[q];
      if (I != J) {
        block_abt(sh.PB, sh.Li, o, lane, warp);
#pragma unroll
        for (int q = 0; q < 4; ++q) sh.LJ[warp + 8 * q][lane] = o[q];
      }
      __syncthreads();
      block_abt(sh.LI, (I != J) ? sh.LJ : sh.LI, o, lane, warp);
#pragma unroll
      for (int q = 0; q < 4; ++q) A[(size_t)(I * NB + warp + 8 * q) * Pp + J * NB + lane] -= o[q];
      if (J == k + 1) {
        for (int q = tid; q < NB * NB; q += 256) pr.Lm[(size_t)(I * NB + (q >> 5)) * Pp + k * NB + (q & 31)] = sh.LI[q >> 5][q & 31];
        if (tid < NB) {
          double v = 0.0;
#pragma unroll 8
          for (int c = 0; c < NB; ++c) v = fma(sh.LI[tid][c], sh.zk[c], v);
          pr.gz[I * NB + tid] -= v;
        }
      }
      __syncthreads();
    }
    grid_barrier(&ctl.bar, (epoch += G));
  }
  if (blockIdx.x != 0) return;
  if (sh.fail) {  // not positive definite: the reference's solve yields NaN -> rejected step (ml.py:fit)
    for (int i = tid; i < P; i += 256) { pr.dp[i] = nan(""); pr.p_try[i] = nan(""); }
    if (tid == 0) { ctl.chol_failed = 1; ctl.denom = nan(""); }
    return;
  }
  // backward substitution L^T dp = z by blocks; the solution overwrites zv
  for (int k = T - 1; k >= 0; --k) {
    double acc = 0.0;
    for (int i = (k + 1) * NB + warp; i < Pp; i += 8) acc = fma(pr.Lm[(size_t)i * Pp + k * NB + lane], pr.zv[i], acc);
    sh.part[warp][lane] = acc;
    __syncthreads();
    if (tid < NB) {
      double v = pr.zv[k * NB + tid];
      for (int w = 0; w < 8; ++w) v -= sh.part[w][tid];
      sh.zk[tid] = v;
    }
    __syncthreads();
    if (tid < NB) {
      double v = 0.0;
      const double* Li = pr.Linv + (size_t)k * NB * NB;
      for (int t = tid; t < NB; ++t) v = fma(Li[t * NB + tid], sh.zk[t], v);  // L11^-T
      pr.zv[k * NB + tid] = v;
    }
    __syncthreads();
  }
  double part = 0.0;
  const double mu = (double)ctl.mu;
  for (int i = tid; i < P; i += 256) {
    const double d = pr.zv[i];
    pr.dp[i] = d;
    pr.p_try[i] = pr.p[i] - d;
    part += d * (mu * d + pr.g[i]);
  }
  part = warp_sum(part);
  if (lane == 0) sh.red[warp] = part;
  __syncthreads();
  if (tid == 0) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += sh.red[w];
    ctl.denom = v;
    ctl.chol_failed = 0;
  }
}

// ---- K4: residuals of the candidate and the loop's decision (ml/optimizers.py:188-232; pysages_b200/ml.py:fit) --------
// Per-CTA partial sums of squares; the last CTA to arrive adds them in a fixed order and accepts or rejects the step.
__global__ void __launch_bounds__(256) lm_cost(const Problem pr) {
  Control& ctl = *pr.ctl;
  if (ctl.done) return;
  extern __shared__ double smem[];
  __shared__ double red[8], red2[8];
  __shared__ int s_last, s_bad;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  double* act = smem + (size_t)warp * ((MAXL + 1) * MAXW);
  const Net& net = pr.net;
  double part = 0.0;
  for (long long s = (long long)blockIdx.x * nwarps + warp; s < pr.n; s += (long long)gridDim.x * nwarps) {
    forward(net, pr.p_try, pr.x + s * net.widths[0], act, lane);
    const double* outv = act + (size_t)net.n_dense * MAXW;
    if (lane < pr.nout) {
      const double e = (double)(float)(outv[lane] - pr.y[s * pr.nout + lane]);
      part += e * e;
    }
    __syncwarp();
  }
  part = warp_sum(part);
  if (lane == 0) red[warp] = part;
  __syncthreads();
  if (tid == 0) {
    double v = 0.0;
    for (int w = 0; w < nwarps; ++w) v += red[w];
    pr.ssq_part[blockIdx.x] = v;
    __threadfence();
    s_last = (atomicAdd(&ctl.cost_count, 1u) == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int P = net.nparams;
  double pp = 0.0, ssq = 0.0;
  bool nanp = false;
  for (int i = tid; i < P; i += blockDim.x) {
    const double v = pr.p_try[i];
    pp += v * v;
    nanp = nanp || (v != v);
  }
  for (int b = tid; b < pr.ssq_parts; b += blockDim.x) ssq += __ldcg(pr.ssq_part + b);
  pp = warp_sum(pp);
  ssq = warp_sum(ssq);
  const bool any_nan = __syncthreads_or(nanp ? 1 : 0) != 0;
  if (lane == 0) { red[warp] = pp; red2[warp] = ssq; }
  __syncthreads();
  if (tid == 0) {
    double ptp = 0.0, ssq_all = 0.0;
    for (int w = 0; w < nwarps; ++w) { ptp += red[w]; ssq_all += red2[w]; }
    const double C = (ssq_all + ctl.reg * ptp) / 2.0;
    const double rho = (ctl.C_prev - C) / ctl.denom;
    float mu = ctl.mu;
    if (rho > (double)ctl.rho_c) mu = fmaxf(mu / ctl.c, ctl.mu_min);
    const bool bad = (rho < (double)ctl.rho_min) || any_nan;
    if (bad) mu = fminf(ctl.c * mu, ctl.mu_max);
    const bool improved = bad || (ctl.C_prev > C);
    ctl.mu = mu;
    ctl.C_try = C;
    ctl.ssq_try = ssq_all;
    ctl.fresh = bad ? 0 : 1;
    if (!bad) { ctl.iters += 1; ctl.C_prev = C; }
    ctl.total_loops += 1;
    if (!(improved && ctl.iters < ctl.max_iters && mu < ctl.mu_max)) ctl.done = 1;
    ctl.bar = 0;
    ctl.cost_count = 0;
    s_bad = bad ? 1 : 0;
  }
  __syncthreads();
  if (!s_bad)  // accept: the candidate becomes the current point (its residuals come with the next lm_jacobian)
    for (int i = tid; i < P; i += blockDim.x) pr.p[i] = pr.p_try[i];
}

__global__ void lm_init(Control* ctl, float mu0, float c, float mu_min, float mu_max, float rho_c, float rho_min,
                        double reg, int max_iters) {
  ctl->C_prev = INFINITY;
  ctl->C_try = INFINITY;
  ctl->denom = 1.0;
  ctl->ssq_try = 0.0;
  ctl->mu = mu0;
  ctl->iters = 0;
  ctl->max_iters = max_iters;
  ctl->fresh = 1;
  ctl->done = (max_iters <= 0 || !(mu0 < mu_max)) ? 1 : 0;
  ctl->chol_failed = 0;
  ctl->total_loops = 0;
  ctl->c = c; ctl->mu_min = mu_min; ctl->mu_max = mu_max; ctl->rho_c = rho_c; ctl->rho_min = rho_min;
  ctl->reg = reg;
  ctl->bar = 0;
  ctl->cost_count = 0;
}


// ---- SpectralABF refit (spectral_abf.py:216-219, approxfun/core.py:232-258) ---------------------------------------
// dy = Fsum / max(hist, 1); scale = max over components of the population standard deviation over the grid (1 if 0);
// periodic grids also remove the mean; coefficients = pinv (gradient Vandermonde matrix, built once on the host) . dy.
// series_normalize: ONE CTA (the data is a grid's worth of numbers), two passes for the variance, fixed-order sums.
// series_project: one CTA per coefficient, 256 lanes striding over the (point, component) columns, fixed-order tree.
__global__ void __launch_bounds__(1024) series_normalize(const uint32_t* __restrict__ hist, const double* __restrict__ Fsum,
                                                         long long npoints, int k, int periodic, double* __restrict__ dyn,
                                                         double* __restrict__ out) {
  __shared__ double red[32];
  __shared__ double s_mean[PSB_MAX_K], s_var[PSB_MAX_K];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  auto block_sum = [&](double v) -> double {
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < 32; ++w) t += red[w];
    return t;
  };
  for (int c = 0; c < k; ++c) {
    double a = 0.0;
    for (long long p = tid; p < npoints; p += 1024) a += Fsum[p * k + c] / (double)max(hist[p], 1u);
    const double mean = block_sum(a) / (double)npoints;
    double q = 0.0;
    for (long long p = tid; p < npoints; p += 1024) {
      const double d = Fsum[p * k + c] / (double)max(hist[p], 1u) - mean;
      q += d * d;
    }
    const double var = block_sum(q) / (double)npoints;
    if (tid == 0) { s_mean[c] = mean; s_var[c] = var; }
  }
  __syncthreads();
  double std = 0.0;
  for (int c = 0; c < k; ++c) std = fmax(std, sqrt(s_var[c]));
  if (std == 0.0) std = 1.0;
  if (tid == 0) out[0] = std;
  for (long long i = tid; i < npoints * k; i += 1024) {
    const long long p = i / k;
    const int c = (int)(i % k);
    const double v = Fsum[i] / (double)max(hist[p], 1u);
    dyn[i] = (periodic ? v - s_mean[c] : v) / std;
  }
}

__global__ void __launch_bounds__(256) series_project(const double* __restrict__ pinv, const double* __restrict__ dyn,
                                                      long long ncols, double* __restrict__ out) {
  __shared__ double red[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* row = pinv + (size_t)blockIdx.x * ncols;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  long long j = tid;
  for (; j + 3 * 256 < ncols; j += 4 * 256) {
    a0 = fma(row[j], dyn[j], a0);
    a1 = fma(row[j + 256], dyn[j + 256], a1);
    a2 = fma(row[j + 512], dyn[j + 512], a2);
    a3 = fma(row[j + 768], dyn[j + 768], a3);
  }
  for (; j < ncols; j += 256) a0 = fma(row[j], dyn[j], a0);
  double v = warp_sum((a0 + a1) + (a2 + a3));
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    out[1 + blockIdx.x] = t;
  }
}

}  // namespace psb_refit

using namespace psb_refit;

struct psb_lm_workspace {
  void* block = nullptr;
  size_t bytes = 0;
  cudaGraphExec_t iter_exec = nullptr;
  Problem key{};
  cudaStream_t key_stream = nullptr;
  int device = 0;
};

extern "C" {

psb_status psbi_fail(psb_status code, const char* msg);  // psb_api.cu: sets psb_last_error

psb_status psb_lm_fit(const psb_lm_problem* q, psb_lm_workspace** ws_io, void* cuda_stream) {
  if (!q || !ws_io) return psbi_fail(PSB_ERR_VALUE, "NULL argument");
  if (q->n_dense < 1 || q->n_dense > PSB_MAX_DENSE) return psbi_fail(PSB_ERR_VALUE, "psb_lm_fit: between 1 and 8 dense layers");
  if (q->activation != PSB_ACT_TANH) return psbi_fail(PSB_ERR_UNSUPPORTED, "psb_lm_fit: tanh networks only (sine layers and the gradient / Sobolev losses are fitted by the host layer)");
  if (q->n < 1 || !q->x || !q->y || !q->params) return psbi_fail(PSB_ERR_VALUE, "psb_lm_fit: empty problem");
  Problem pr;
  memset(&pr, 0, sizeof(pr));  // padding bytes too: the struct is the graph cache's key
  Net& net = pr.net;
  net.n_dense = q->n_dense;
  int P = 0;
  for (int l = 0; l <= q->n_dense; ++l) {
    if (q->widths[l] < 1 || q->widths[l] > PSB_MAX_WIDTH) return psbi_fail(PSB_ERR_VALUE, "psb_lm_fit: layer widths must lie in [1, 128]");
    net.widths[l] = q->widths[l];
    if (l < q->n_dense) {
      net.w_off[l] = P;
      P += q->widths[l] * q->widths[l + 1] + q->widths[l + 1];
    }
  }
  if (q->widths[0] > PSB_MAX_K || q->widths[q->n_dense] > 32) return psbi_fail(PSB_ERR_VALUE, "psb_lm_fit: at most 8 inputs and 32 outputs");
  if (P > 1024) return psbi_fail(PSB_ERR_UNSUPPORTED, "psb_lm_fit: at most 1024 network parameters (dense normal equations)");
  net.nparams = P;
  for (int d = 0; d < PSB_MAX_K; ++d) { net.lower[d] = q->lower[d]; net.upper[d] = q->upper[d]; }
  pr.n = q->n;
  pr.nout = q->widths[q->n_dense];
  const size_t M = (size_t)q->n * pr.nout;
  cudaStream_t stream = (cudaStream_t)cuda_stream;
  int sms = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int cost_ctas = 2 * sms;
  const int T = (P + NB - 1) / NB, Pp = T * NB;
  const bool small_tiles = P <= NORMAL64_MAX_P;
  const int tn = small_tiles ? TN64 : TN, tk = small_tiles ? TK64 : TK;
  const int tiles = (P + tn - 1) / tn, ntri = tiles * (tiles + 1) / 2;
  int per_sm = 1;
  if (small_tiles) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lm_normal64, 256, 0);
  else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lm_normal, 256, 0);
  int splits = std::max(1, std::min(MAX_SPLIT, std::max(1, per_sm) * sms / ntri));  // one balanced wave
  splits = (int)std::max<long long>(1, std::min<long long>(splits, ((long long)q->n * pr.nout + 4 * tk - 1) / (4 * tk)));
  pr.T = T;
  pr.splits = splits;
  // workspace layout (doubles unless noted)
  size_t off = 0;
  auto take = [&](size_t count) { size_t o = off; off += (count * 8 + 255) & ~(size_t)255; return o; };
  const size_t o_ptry = take(P), o_e = take(M), o_J = take(M * P), o_H = take((size_t)splits * P * P),
               o_gp = take((size_t)splits * P), o_Hs = take((size_t)P * P), o_gs = take(P), o_A = take((size_t)Pp * Pp), o_L = take((size_t)Pp * Pp),
               o_Li = take((size_t)T * NB * NB), o_gz = take(Pp), o_zv = take(Pp), o_g = take(P), o_dp = take(P),
               o_ssq = take(cost_ctas), o_ctl = take((sizeof(Control) + 7) / 8);
  psb_lm_workspace* ws = *ws_io;
  if (!ws) { ws = new psb_lm_workspace(); memset(&ws->key, 0, sizeof(Problem)); ws->device = dev; *ws_io = ws; }
  if (ws->bytes < off) {
    if (cudaStreamSynchronize(stream) != cudaSuccess) return psbi_fail(PSB_ERR_CUDA, "psb_lm_fit: stream error");
    if (ws->iter_exec) { cudaGraphExecDestroy(ws->iter_exec); ws->iter_exec = nullptr; }
    if (ws->block) cudaFree(ws->block);
    if (cudaMalloc(&ws->block, off) != cudaSuccess) { ws->block = nullptr; ws->bytes = 0; return psbi_fail(PSB_ERR_CUDA, "psb_lm_fit: out of device memory"); }
    ws->bytes = off;
  }
  char* base = reinterpret_cast<char*>(ws->block);
  pr.x = q->x; pr.y = q->y; pr.p = q->params;
  pr.p_try = (double*)(base + o_ptry); pr.e = (double*)(base + o_e);
  pr.J = (double*)(base + o_J); pr.Hpart = (double*)(base + o_H); pr.gpart = (double*)(base + o_gp);
  pr.H = (double*)(base + o_Hs); pr.gsum = (double*)(base + o_gs);
  pr.A = (double*)(base + o_A); pr.Lm = (double*)(base + o_L); pr.Linv = (double*)(base + o_Li);
  pr.gz = (double*)(base + o_gz); pr.zv = (double*)(base + o_zv); pr.g = (double*)(base + o_g); pr.dp = (double*)(base + o_dp);
  pr.ssq_part = (double*)(base + o_ssq); pr.ssq_parts = cost_ctas; pr.ctl = (Control*)(base + o_ctl);

  const size_t smem_j = 8 * ((MAXL + 1) * MAXW + 2 * MAXW) * sizeof(double), smem_c = 8 * ((MAXL + 1) * MAXW) * sizeof(double);
  cudaFuncSetAttribute(lm_jacobian, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_j);
  cudaFuncSetAttribute(lm_cost, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c);
  cudaFuncSetAttribute(lm_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SolveShared));
  const int jac_ctas = (int)std::min<long long>((q->n + 7) / 8, 4LL * sms);
  // the solve is cooperative (grid barriers): one CTA per SM at most, no more than there are trailing tiles
  const int solve_ctas = std::max(1, std::min(sms, std::max(T * (T - 1) / 2, (Pp * Pp + 8191) / 8192)));

  lm_init<<<1, 1, 0, stream>>>(pr.ctl, q->mu_0, q->mu_c, q->mu_min, q->mu_max, q->rho_c, q->rho_min, q->reg, q->max_iters);
  // e of the initial parameters is produced by the first lm_jacobian (fresh = 1)
  auto enqueue_iteration = [&](cudaStream_t s) {
    lm_jacobian<<<jac_ctas, 256, smem_j, s>>>(pr);
    if (small_tiles) lm_normal64<<<ntri * splits, 256, 0, s>>>(pr);
    else lm_normal<<<ntri * splits, 256, 0, s>>>(pr);
    lm_reduce<<<(P * P + 255) / 256, 256, 0, s>>>(pr);
    void* args[] = {(void*)&pr};
    cudaLaunchCooperativeKernel((const void*)lm_solve, dim3(solve_ctas), dim3(256), args, sizeof(SolveShared), s);
    lm_cost<<<cost_ctas, 256, smem_c, s>>>(pr);
  };
  // one iteration captured once per problem shape, replayed: rejected steps do not count, hence the factor
  const bool same = ws->iter_exec && ws->key_stream == stream && memcmp(&ws->key, &pr, sizeof(Problem)) == 0;
  if (!same && stream != nullptr) {
    if (ws->iter_exec) { cudaGraphExecDestroy(ws->iter_exec); ws->iter_exec = nullptr; }
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      enqueue_iteration(stream);
      if (cudaStreamEndCapture(stream, &graph) == cudaSuccess && graph) {
        if (cudaGraphInstantiate(&ws->iter_exec, graph, 0) != cudaSuccess) ws->iter_exec = nullptr;
        cudaGraphDestroy(graph);
      }
    }
    cudaGetLastError();
  }
  memcpy(&ws->key, &pr, sizeof(Problem));
  ws->key_stream = stream;
  // upper bound of loop passes: every accepted step counts towards max_iters; mu grows by c per rejection and the
  // loop ends at mu_max, so at most log_c(mu_max / mu_min) rejections can separate two accepted steps
  const int max_loops = q->max_loops > 0 ? q->max_loops : 2 * q->max_iters + 64;
  for (int it = 0; it < max_loops; ++it) {
    if (ws->iter_exec && stream != nullptr) cudaGraphLaunch(ws->iter_exec, stream);
    else enqueue_iteration(stream);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return psbi_fail(PSB_ERR_CUDA, cudaGetErrorString(e));
  return PSB_OK;
}

psb_status psb_lm_result(psb_lm_workspace* ws, int32_t* iters, double* cost, float* mu, int32_t* loops, void* cuda_stream) {
  if (!ws || !ws->block) return psbi_fail(PSB_ERR_VALUE, "psb_lm_result: no fit has run");
  Control c;
  cudaStream_t stream = (cudaStream_t)cuda_stream;
  if (cudaMemcpyAsync(&c, ws->key.ctl, sizeof(Control), cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
      cudaStreamSynchronize(stream) != cudaSuccess)
    return psbi_fail(PSB_ERR_CUDA, "psb_lm_result: copy failed");
  if (iters) *iters = c.iters;
  if (cost) *cost = c.C_prev;
  if (mu) *mu = c.mu;
  if (loops) *loops = c.total_loops;
  return PSB_OK;
}

psb_status psb_series_fit(const psb_series_fit_desc* d, void* cuda_stream) {
  if (!d || !d->hist || !d->Fsum || !d->pinv || !d->out || !d->work) return psbi_fail(PSB_ERR_VALUE, "psb_series_fit: NULL argument");
  if (d->npoints < 1 || d->k < 1 || d->k > PSB_MAX_K || d->nterms < 1) return psbi_fail(PSB_ERR_VALUE, "psb_series_fit: empty problem or too many components");
  cudaStream_t stream = (cudaStream_t)cuda_stream;
  series_normalize<<<1, 1024, 0, stream>>>(reinterpret_cast<const uint32_t*>(d->hist), d->Fsum, d->npoints, d->k, d->periodic, d->work, d->out);
  series_project<<<d->nterms, 256, 0, stream>>>(d->pinv, d->work, d->npoints * d->k, d->out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return psbi_fail(PSB_ERR_CUDA, cudaGetErrorString(e));
  return PSB_OK;
}

void psb_lm_destroy(psb_lm_workspace* ws) {
  if (!ws) return;
  if (ws->iter_exec) cudaGraphExecDestroy(ws->iter_exec);
  if (ws->block) cudaFree(ws->block);
  delete ws;
}

}  // extern "C"
