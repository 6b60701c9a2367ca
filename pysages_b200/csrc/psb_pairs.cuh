// psb_pairs.cuh -- pairwise coordination-number CV (extension).
//
//   xi = sum_{i in A} sum_{j in B} s(r_ij),  s = (1 - x^n) / (1 - x^m),  x = r / r0,  m = 2n
//      = 1 / (1 + u^h),  u = r^2 / r0^2,  h = n / 2        (n = 6, m = 12 -> 1 / (1 + x^6))
//   dxi/dr_i = sum_j c_ij (r_i - r_j),  dxi/dr_j = -sum_i c_ij (r_i - r_j),
//   c = s'(r) / r = -2 h u^(h-1) s^2 / r0^2                 (no sqrt, one reciprocal per pair)
//
// FP64-pipe bound (positions of both groups are ~0.5 MB and stay in L1/L2).  Each lane keeps R
// rows of A in registers; a 32-column tile of B is loaded coalesced, one column per lane, and
// ROTATED through the warp with shuffles together with its gradient accumulator, so both the row
// and the column gradients accumulate in registers with no per-pair reduction and no atomics in
// the inner loop.  Row/column sums leave the warp once per tile as fp64 RED atomics.
#pragma once

#include "psb_device.cuh"

namespace psb {

constexpr double PAIR_PAD = 1.0e30;  // padded coordinates: s underflows to exactly 0 gradient

// fast fp64 reciprocal for den >= 1: MUFU.RCP64H seed (>= 20 good bits) + one second-order Newton step
// y (1 + e + e^2), e = 1 - d y: relative error e^3 <= 2^-60, one DFMA fewer than two first-order steps
__device__ __forceinline__ double rcp_fast(double d) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double e = fma(-d, y, 1.0);
  const double p = fma(e, e, e);  // e + e^2
  return fma(y, p, y);
}

// Gather + unwrap the positions of one coordination CV into SoA fp64 arrays (padded to 32) and
// clear its gradient workspace.  xa: [3][na_pad], xb: [3][nb_pad].
template <int LAYOUT, typename Real>
__global__ void coord_gather_kernel(const Model* __restrict__ Mdev, const __grid_constant__ Snap S,
                                    int c, double* __restrict__ xa, int na_pad,
                                    double* __restrict__ xb, int nb_pad) {
  using A = Access<LAYOUT, Real>;
  const Model& M = *Mdev;  // the handle's device copy (a by-value parameter would exceed 4 KB)
  const CvDev& cv = M.cv[c];
  const GroupDev& GA = M.grp[cv.g0];
  const GroupDev& GB = M.grp[cv.g0 + 1];
  const int nA = (int)(GA.t1 - GA.t0), nB = (int)(GB.t1 - GB.t0);
  const int total = na_pad + nb_pad;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const bool isA = i < na_pad;
    const int j = isA ? i : i - na_pad;
    const int n = isA ? nA : nB;
    double* dst = isA ? xa : xb;
    const int pad = isA ? na_pad : nb_pad;
    V3 r = v3(isA ? PAIR_PAD * cv.r0 : -PAIR_PAD * cv.r0, 0.0, 0.0);
    if (j < n) {
      const uint32_t t = (isA ? GA.t0 : GB.t0) + j;
      const uint32_t slot = A::slot(S, M, __ldg(M.term_tag + t));
      M.term_slot[t] = slot;
      r = A::position(S, slot);
      double* g = cv.grad + 3 * (size_t)(t - cv.t0);
      g[0] = 0.0; g[1] = 0.0; g[2] = 0.0;
    }
    dst[0 * pad + j] = r.x;
    dst[1 * pad + j] = r.y;
    dst[2 * pad + j] = r.z;
  }
}

template <int R, int H>
__global__ void __launch_bounds__(PAIR_BLOCK, 2)
coord_pairs_kernel(const double* __restrict__ xa, int na_pad, const double* __restrict__ xb,
                   int nb_pad, int nA, int nB, int nrt, int h_rt, double inv_r02,
                   double* __restrict__ gA, double* __restrict__ gB, double* __restrict__ xi_part) {
  __shared__ double s_red[PAIR_BLOCK / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rt = blockIdx.x % nrt, cs = blockIdx.x / nrt;
  const int ncs = gridDim.x / nrt;
  const int rows_per_cta = PAIR_BLOCK * R;
  const int row_base = rt * rows_per_cta + warp * 32 * R + lane;

  double rx[R], ry[R], rz[R], gx[R], gy[R], gz[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int row = row_base + 32 * r;
    const bool ok = row < na_pad;
    rx[r] = ok ? __ldg(xa + 0 * na_pad + row) : PAIR_PAD;
    ry[r] = ok ? __ldg(xa + 1 * na_pad + row) : 0.0;
    rz[r] = ok ? __ldg(xa + 2 * na_pad + row) : 0.0;
    gx[r] = gy[r] = gz[r] = 0.0;
  }
  const int ctiles = nb_pad / 32;
  const int per = (ctiles + ncs - 1) / ncs;
  const int ct0 = min(ctiles, cs * per), ct1 = min(ctiles, ct0 + per);
  const double kc = -2.0 * (double)H * inv_r02;
  const int src = (lane + 1) & 31;
  double xi = 0.0;

  for (int tile = ct0; tile < ct1; ++tile) {
    const int col = tile * 32 + lane;
    double cx = __ldg(xb + 0 * nb_pad + col);
    double cy = __ldg(xb + 1 * nb_pad + col);
    double cz = __ldg(xb + 2 * nb_pad + col);
    double ax = 0.0, ay = 0.0, az = 0.0;
#pragma unroll 2
    for (int s = 0; s < 32; ++s) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const double dx = rx[r] - cx, dy = ry[r] - cy, dz = rz[r] - cz;
        const double u = (dx * dx + dy * dy + dz * dz) * inv_r02;
        double pw = u;  // u^(H-1)
#pragma unroll
        for (int i = 2; i < H; ++i) pw *= u;
        const double sw = rcp_fast(fma(pw, u, 1.0));
        xi += sw;
        const double c = kc * (H > 1 ? pw : 1.0) * sw * sw;
        gx[r] = fma(c, dx, gx[r]); gy[r] = fma(c, dy, gy[r]); gz[r] = fma(c, dz, gz[r]);
        ax = fma(-c, dx, ax); ay = fma(-c, dy, ay); az = fma(-c, dz, az);
      }
      cx = __shfl_sync(0xffffffffu, cx, src); cy = __shfl_sync(0xffffffffu, cy, src);
      cz = __shfl_sync(0xffffffffu, cz, src);
      ax = __shfl_sync(0xffffffffu, ax, src); ay = __shfl_sync(0xffffffffu, ay, src);
      az = __shfl_sync(0xffffffffu, az, src);
    }
    // after 32 rotations lane l again holds column tile*32 + l, now summed over this warp's rows
    if (col < nB) {
      atomicAdd(gB + 3 * (size_t)col + 0, ax);
      atomicAdd(gB + 3 * (size_t)col + 1, ay);
      atomicAdd(gB + 3 * (size_t)col + 2, az);
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int row = row_base + 32 * r;
    if (row < nA && ct1 > ct0) {
      atomicAdd(gA + 3 * (size_t)row + 0, gx[r]);
      atomicAdd(gA + 3 * (size_t)row + 1, gy[r]);
      atomicAdd(gA + 3 * (size_t)row + 2, gz[r]);
    }
  }
  xi = warp_sum(xi);
  if (lane == 0) s_red[warp] = xi;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0;
    for (int w = 0; w < PAIR_BLOCK / 32; ++w) v += s_red[w];
    xi_part[blockIdx.x] = v;
  }
  (void)h_rt;
}

}  // namespace psb
