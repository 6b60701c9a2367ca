// psb_step.cuh -- the fused per-timestep bias kernel (sm_100a).
//
// One persistent (cooperative when gridDim.x > 1) launch per MD step:
//   pass A   gather selected atoms (tag -> slot -> position [+ image unwrap] [+ momentum]),
//            per-group fp64 block reductions                       (colvars/core.py:193 + CV sums)
//   barrier  the LAST CTA to arrive reduces the per-CTA partials deterministically, evaluates
//            the CVs + closed-form Jacobian factors and, when possible, the method update
//   pass B   (RMSD only) residual sum with the optimal rotation    (orientation.py:76-95)
//   pass C   (ABF with position-dependent Jacobians) J J^T and J p (abf.py:210-213)
//   pass H   (Metadynamics, direct) hill sum over TMA-staged hills (metad.py:291, 318-323)
//   pass D   (Metadynamics, grid, deposit steps) grid update       (metad.py:251-256)
//   pass S   scatter: forces[slot].xyz += -sum_c F_c dxi_c/dr      (hoomd.py:219-230, ...)
// The dense (k, 3N) Jacobian and (N, 3) bias of the reference are never materialised
// (unless Model::jac_dense / bias_dense are set for the parity tests).
#pragma once

#include "psb_device.cuh"

namespace psb {

// ---- small device utilities -----------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// TMA 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src, uint32_t bytes,
                                             uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      :: "r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
               :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}"
      :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ long long global_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define PSB_STAMP(slot)                                                              \
  do {                                                                               \
    if (M.dbg && threadIdx.x == 0) M.dbg[(size_t)blockIdx.x * 16 + (slot)] = global_ns(); \
  } while (0)

struct StepShared {
  double red[NWARPS][NACC];
  double tot[MAXG + 1][NACC];
  Fin fin;  // CTA-local copy of the finalizer results
  unsigned long long gen;
  uint64_t mbar[2];
};

// Block-wide sum of acc[0..nacc); result lands in out[0..nacc) (shared or global, stride ostride).
__device__ __forceinline__ void block_reduce_store(StepShared& sh, const double* acc, int nacc,
                                                   double* out, size_t ostride, bool out_global) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int a = 0; a < nacc; ++a) {
    double v = warp_sum(acc[a]);
    if (lane == 0) sh.red[warp][a] = v;
  }
  __syncthreads();
  if ((int)threadIdx.x < nacc) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < NWARPS; ++w) v += sh.red[w][threadIdx.x];
    if (out_global) __stcg(out + (size_t)threadIdx.x * ostride, v);
    else out[(size_t)threadIdx.x * ostride] = v;
  }
  __syncthreads();
}

// Grid barrier with a finalizer: CTA 0 waits until every CTA has arrived, runs `finalize` (all its
// threads) and then releases the others.  Pinning the finalizer to one CTA keeps its (large, rarely
// executed) code resident in that SM's instruction cache from step to step.  Counters are
// monotonic and never reset; gridDim.x is fixed per handle.
template <typename Fn>
__device__ __forceinline__ void grid_barrier(const Model& M, StepShared& sh, int id, Fn&& finalize) {
  if (gridDim.x == 1) {
    __syncthreads();
    finalize();
    __syncthreads();
    return;
  }
  BarrierState* bs = M.bars + id;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long t = atomicAdd(&bs->arrive, 1ULL);
    sh.gen = t / gridDim.x;
  }
  __syncthreads();
  if (blockIdx.x == 0) {
    if (threadIdx.x == 0) {
      const unsigned long long want = (sh.gen + 1) * gridDim.x;
      while (ld_acquire_u64(&bs->arrive) < want) {}
    }
    __syncthreads();
    finalize();
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(&bs->release, 1ULL);
    }
  } else {
    if (threadIdx.x == 0) {
      while (ld_acquire_u64(&bs->release) <= sh.gen) __nanosleep(32);
    }
    __syncthreads();
  }
}

__host__ __device__ __forceinline__ bool is_point_kind(int kind) {
  return kind >= PSB_CV_COMPONENT && kind <= PSB_CV_DIHEDRAL;
}

// accumulators per group in pass `pass` (0 = A, 1 = B)
__device__ __forceinline__ int nacc_for(const Model& M, const Snap& S, int g, int pass) {
  const int kind = M.cv[M.grp[g].cv].kind;
  if (pass == 0) {
    if (is_point_kind(kind)) return (S.need_momenta && M.abf_fast) ? 6 : 3;
    if (kind == PSB_CV_ROG) return 1;
    if (kind == PSB_CV_RMSD) return 15;
    return 0;
  }
  return kind == PSB_CV_RMSD ? 1 : 0;
}

// Deterministic reduction of the per-CTA partials of pass `pass` into sh.tot (run by one CTA).
static __device__ __noinline__ void reduce_partials(const Model& M, const Snap& S, StepShared& sh,
                                                int pass, int vgroup_nacc) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nb = gridDim.x;
  if (nb > 1) {
    const uint32_t chunk = (M.T + nb - 1) / nb;
    const int npairs = (MAXG + 1) * NACC;
    for (int pair = warp; pair < npairs; pair += NWARPS) {
      const int g = pair / NACC, a = pair % NACC;
      int blo, bhi, na;
      if (g < MAXG) {
        if (g >= M.ngroups) continue;
        if (pass > 1) continue;
        na = nacc_for(M, S, g, pass);
        if (M.grp[g].t1 <= M.grp[g].t0) continue;
        blo = (int)(M.grp[g].t0 / chunk);
        bhi = (int)((M.grp[g].t1 - 1) / chunk);
      } else {
        na = vgroup_nacc;
        blo = 0;
        bhi = nb - 1;
      }
      if (a >= na) continue;
      double v = 0.0;
      for (int bb = blo + lane; bb <= bhi; bb += 32) v += __ldcg(M.partials + (size_t)pair * nb + bb);
      v = warp_sum(v);
      if (lane == 0) sh.tot[g][a] = v;
    }
  }
  // coordination CVs: partial sums come from the pair kernel
  if (pass == 0 && M.any_coord) {
    for (int c = warp; c < M.ncv; c += NWARPS) {
      if (M.cv[c].kind != PSB_CV_COORDINATION) continue;
      double v = 0.0;
      for (int i = lane; i < M.cv[c].n_part; i += 32) v += __ldcg(M.cv[c].xi_part + i);
      v = warp_sum(v);
      if (lane == 0) sh.tot[M.cv[c].g0][0] = v;
    }
  }
  __syncthreads();
}

// ---- finalizers (thread 0 of one CTA) ----------------------------------------------------------
__device__ inline V3 tot3(const StepShared& sh, int g, int o) {
  return v3(sh.tot[g][o], sh.tot[g][o + 1], sh.tot[g][o + 2]);
}
__device__ inline void put3(double* dst, V3 v) {
  dst[0] = v.x; dst[1] = v.y; dst[2] = v.z;
}

// CV values and Jacobian factors after pass A.
static __device__ __noinline__ void finalize_cvs_A(const Model& M, StepShared& sh) {
  Fin& f = sh.fin;
  for (int c = 0; c < M.ncv; ++c) {
    const CvDev& cv = M.cv[c];
    const int g0 = cv.g0;
    V3 p[4];
    if (is_point_kind(cv.kind))
      for (int j = 0; j < cv.ngroups; ++j) p[j] = M.grp[g0 + j].inv_n * tot3(sh, g0 + j, 0);
    for (int j = 0; j < cv.ngroups && is_point_kind(cv.kind); ++j)
      for (int q = 0; q < 3; ++q) put3(f.gpt[g0 + j][q], v3(0, 0, 0));
    switch (cv.kind) {
      case PSB_CV_COMPONENT: {  // coordinates.py:28,71
        f.xi[cv.comp0] = comp(p[0], cv.axis);
        f.gpt[g0][0][cv.axis] = 1.0;
      } break;
      case PSB_CV_DISTANCE: {
        V3 a, b;
        f.xi[cv.comp0] = distance_cv(p[0], p[1], &a, &b);
        put3(f.gpt[g0][0], a);
        put3(f.gpt[g0 + 1][0], b);
      } break;
      case PSB_CV_DISPLACEMENT: {  // coordinates.py:148: r2 - r1
        V3 d = p[1] - p[0];
        f.xi[cv.comp0 + 0] = d.x; f.xi[cv.comp0 + 1] = d.y; f.xi[cv.comp0 + 2] = d.z;
        for (int q = 0; q < 3; ++q) {
          f.gpt[g0][q][q] = -1.0;
          f.gpt[g0 + 1][q][q] = 1.0;
        }
      } break;
      case PSB_CV_ANGLE: {
        V3 a, b, d;
        f.xi[cv.comp0] = angle_cv(p[0], p[1], p[2], &a, &b, &d);
        put3(f.gpt[g0][0], a); put3(f.gpt[g0 + 1][0], b); put3(f.gpt[g0 + 2][0], d);
      } break;
      case PSB_CV_DIHEDRAL: {
        V3 a, b, d, e;
        f.xi[cv.comp0] = dihedral_cv(p[0], p[1], p[2], p[3], &a, &b, &d, &e);
        put3(f.gpt[g0][0], a); put3(f.gpt[g0 + 1][0], b);
        put3(f.gpt[g0 + 2][0], d); put3(f.gpt[g0 + 3][0], e);
      } break;
      case PSB_CV_ROG: {  // shape.py:52-57: sum r.r / n ; d/dr_i = 2 r_i / n
        f.xi[cv.comp0] = sh.tot[g0][0] * M.grp[g0].inv_n;
        f.cvx[c][0] = 2.0 * M.grp[g0].inv_n;
      } break;
      case PSB_CV_RMSD: {  // orientation.py:33-73
        const V3 rbar = tot3(sh, g0, 0);  // sum_i w_i r_i
        const V3 sumr = tot3(sh, g0, 3);
        double C[9];
        const double rb[3] = {rbar.x, rbar.y, rbar.z};
        for (int a = 0; a < 3; ++a)
          for (int b = 0; b < 3; ++b) C[a * 3 + b] = sh.tot[g0][6 + a * 3 + b] - rb[a] * cv.sumQ[b];
        double* x = f.cvx[c];
        kabsch_rotation(C, x);  // U -> x[0..8]
        put3(x + 9, rbar);
        // sum_i D_i = sum_i P_i - (sum_i Q_i) U^T
        const double n = 1.0 / M.grp[g0].inv_n;
        const double sp[3] = {sumr.x - n * rbar.x, sumr.y - n * rbar.y, sumr.z - n * rbar.z};
        for (int a = 0; a < 3; ++a)
          x[12 + a] = sp[a] - (cv.sumQ[0] * x[a * 3 + 0] + cv.sumQ[1] * x[a * 3 + 1] + cv.sumQ[2] * x[a * 3 + 2]);
      } break;
      case PSB_CV_COORDINATION: {
        f.xi[cv.comp0] = sh.tot[g0][0];
      } break;
      default: break;
    }
  }
}

// RMSD value after pass B.
static __device__ __noinline__ void finalize_cvs_B(const Model& M, StepShared& sh) {
  Fin& f = sh.fin;
  for (int c = 0; c < M.ncv; ++c) {
    const CvDev& cv = M.cv[c];
    if (cv.kind != PSB_CV_RMSD) continue;
    const double inv_n = M.grp[cv.g0].inv_n;
    const double xi = sqrt(sh.tot[cv.g0][0] * inv_n);  // orientation.py:94
    f.xi[cv.comp0] = xi;
    f.cvx[c][15] = inv_n / xi;  // 1 / (n rmsd)
  }
}

// grid bin of xi; returns true when any index == shape (out of bounds marker, grids.py:120)
static __device__ __noinline__ bool grid_bin(const MethodDev& m, const double* xi, uint32_t* I) {
  bool oob = false;
  for (int d = 0; d < m.grid_ndim; ++d) {
    I[d] = grid_index_1d(m.grid_kind, xi[d], m.g_lower[d], m.g_upper[d], m.g_shape[d]);
    oob |= (I[d] == (uint32_t)m.g_shape[d]);
  }
  return oob;
}
// flat row-major offset with JAX's clamp-on-gather semantics
__device__ inline long long grid_flat_clamped(const MethodDev& m, const uint32_t* I) {
  long long o = 0;
  for (int d = 0; d < m.grid_ndim; ++d) {
    uint32_t i = I[d] < (uint32_t)m.g_shape[d] ? I[d] : (uint32_t)m.g_shape[d] - 1;
    o = o * m.g_shape[d] + i;
  }
  return o;
}

// fg[g] = -sum_j F[comp0+j] dxi_{comp0+j}/dpoint_g / n_g  (the -J^T F of harmonic_bias.py:127,
// metad.py:200, abf.py:223 restricted to one group)
static __device__ __noinline__ void compute_fg(const Model& M, Fin& f) {
  for (int g = 0; g < M.ngroups; ++g) {
    const CvDev& cv = M.cv[M.grp[g].cv];
    double v[3] = {0, 0, 0};
    if (is_point_kind(cv.kind))
      for (int j = 0; j < cv.ncomp; ++j)
        for (int a = 0; a < 3; ++a) v[a] -= f.F[cv.comp0 + j] * f.gpt[g][j][a] * M.grp[g].inv_n;
    f.fg[g][0] = v[0]; f.fg[g][1] = v[1]; f.fg[g][2] = v[2];
  }
}

static __device__ __noinline__ void publish_force(const Model& M, Fin& f, long long ncalls_new) {
  for (int c = 0; c < M.k; ++c) M.st.F[c] = f.F[c];
  *M.st.energy = f.energy;
  *M.st.ncalls = ncalls_new;
  compute_fg(M, f);
}

// abf.py:213-222 given W p inputs
static __device__ __noinline__ void abf_update(const Model& M, const Snap& S, Fin& f, const double* JJt,
                                  const double* Jp) {
  const MethodDev& m = M.m;
  const int k = M.k;
  double Wp[MAXK] = {0, 0, 0, 0};
  if (m.use_pinv) {
    pinv_sym_solve(k, JJt, Jp, Wp, 10.0 * (double)(3 * M.natoms) * 2.220446049250313e-16);
  } else if (!chol_solve(k, JJt, Jp, Wp)) {
    for (int c = 0; c < k; ++c) Wp[c] = nan("");
  }
  uint32_t I[PSB_MAX_GRID_DIMS];
  const bool oob = grid_bin(m, f.xi, I);
  bool in_range = true;
  for (int d = 0; d < k; ++d) in_range &= (I[d] < (uint32_t)m.g_shape[d]);
  const long long flat = grid_flat_clamped(m, I);
  if (in_range) {  // scatter-add is dropped when out of range
    M.st.hist[flat] += 1u;
    for (int c = 0; c < k; ++c) {
      const double dWp_dt = (1.5 * Wp[c] - 2.0 * M.st.Wp[c] + 0.5 * M.st.Wp_[c]) / S.dt;
      M.st.Fsum[flat * k + c] += m.force_unit_factor * dWp_dt + M.st.force[c];
    }
  }
  for (int c = 0; c < k; ++c) {
    double fc;
    if (m.has_restraints && oob) {
      fc = restraint_force(m.r_lower[c], m.r_upper[c], m.r_kl[c], m.r_ku[c], f.xi[c]);
    } else {  // gather clamps
      const unsigned long long h = M.st.hist[flat];
      const double den = (double)((unsigned long long)m.abf_N > h ? (unsigned long long)m.abf_N : h);
      fc = M.st.Fsum[flat * k + c] / den;
    }
    f.F[c] = fc;
    M.st.force[c] = fc;
    M.st.Wp_[c] = M.st.Wp[c];
    M.st.Wp[c] = Wp[c];
  }
  for (int d = 0; d < PSB_MAX_GRID_DIMS; ++d) f.I[d] = d < k ? I[d] : 0;
  f.oob = oob;
  f.energy = 0.0;
}

// Value/gradient of one Gaussian at xi (utils/core.py:113-117, metad.py:318-323)
__device__ __forceinline__ double hill_eval(const MethodDev& m, int k, const double* xi,
                                            const double* centre, double height, double* dV) {
  double d[MAXK], q = 0.0;
#pragma unroll
  for (int c = 0; c < MAXK; ++c)
    if (c < k) {
      d[c] = wrap_period(xi[c] - centre[c], m.periods[c]);
      const double z = d[c] / m.sigma[c];
      q += z * z;
    }
  const double e = height * exp(-q / 2.0);
#pragma unroll
  for (int c = 0; c < MAXK; ++c)
    if (c < k) dV[c] += -e * d[c] / (m.sigma[c] * m.sigma[c]);
  return e;
}

// Everything that can be decided once xi is complete (thread 0 of the finalizing CTA).
// Returns through sh.fin; `stage_done` tells whether F / fg are final.
static __device__ __noinline__ void post_cv(const Model& M, const Snap& S, StepShared& sh) {
  Fin& f = sh.fin;
  const MethodDev& m = M.m;
  const int k = M.k;
  for (int c = 0; c < k; ++c) M.st.xi[c] = f.xi[c];
  f.deposit = 0;
  f.oob = 0;
  f.energy = 0.0;
  for (int c = 0; c < MAXK; ++c) f.F[c] = 0.0;
  if (S.mode == MODE_EVAL) return;
  const long long ncalls = *M.st.ncalls + 1;
  switch (m.kind) {
    case PSB_METHOD_UNBIASED: publish_force(M, f, ncalls); break;
    case PSB_METHOD_HARMONIC: {  // harmonic_bias.py:124-128 (no periodic wrap: quirk kept)
      double e = 0.0;
      for (int a = 0; a < k; ++a) {
        double s = 0.0;
        for (int b = 0; b < k; ++b) s += m.kspring[a * MAXK + b] * (f.xi[b] - m.center[b]);
        f.F[a] = s;
        e += 0.5 * (f.xi[a] - m.center[a]) * s;
      }
      f.energy = e;
      publish_force(M, f, ncalls);
    } break;
    case PSB_METHOD_ABF: {
      if (!M.abf_fast) break;  // needs pass C
      // J J^T and J p from per-group sums: point-CV Jacobians are constant within a group
      double JJt[16], Jp[MAXK];
      for (int i = 0; i < 16; ++i) JJt[i] = 0.0;
      for (int i = 0; i < MAXK; ++i) Jp[i] = 0.0;
      for (int g = 0; g < M.ngroups; ++g) {
        const CvDev& ca = M.cv[M.grp[g].cv];
        const V3 ps = tot3(sh, g, 3);
        for (int ja = 0; ja < ca.ncomp; ++ja) {
          const double* ga = f.gpt[g][ja];
          Jp[ca.comp0 + ja] += (ga[0] * ps.x + ga[1] * ps.y + ga[2] * ps.z) * M.grp[g].inv_n;
          for (int h = 0; h < M.ngroups; ++h) {
            if (M.ov[g][h] == 0) continue;
            const CvDev& cb = M.cv[M.grp[h].cv];
            for (int jb = 0; jb < cb.ncomp; ++jb) {
              const double* gb = f.gpt[h][jb];
              JJt[(ca.comp0 + ja) * 4 + cb.comp0 + jb] += (double)M.ov[g][h] *
                  (ga[0] * gb[0] + ga[1] * gb[1] + ga[2] * gb[2]) * M.grp[g].inv_n * M.grp[h].inv_n;
            }
          }
        }
      }
      abf_update(M, S, f, JJt, Jp);
      publish_force(M, f, ncalls);
    } break;
    case PSB_METHOD_METAD: {
      const bool in_dep = (ncalls > 1) && (ncalls % m.stride == 1);  // metad.py:192-193
      if (m.grid_kind == PSB_GRID_NONE) {
        f.deposit = in_dep ? 1 : 0;  // heights need V: decided after pass H
        break;
      }
      const bool oob = grid_bin(m, f.xi, f.I);
      f.oob = oob;
      const long long flat = grid_flat_clamped(m, f.I);
      if (in_dep && !oob) {  // metad.py:258-271
        double h = m.height;
        if (m.well_tempered) h = m.height * exp(-M.st.gp[flat] / m.deltaT_kB);  // metad.py:223-229
        for (int c = 0; c < k; ++c) f.hill[c] = f.xi[c];
        f.hill[k] = h;
        f.hill[k + 1] = 1.0;
        f.deposit = 1;
        if (S.mode == MODE_STEP) {
          const long long idx = *M.st.idx;
          if (idx < m.ngaussians) {
            M.st.heights[idx] = h;
            for (int c = 0; c < k; ++c) M.st.centers[idx * k + c] = f.xi[c];
          }
          *M.st.idx = idx + 1;
        }
      } else {
        for (int c = 0; c < k; ++c) f.hill[c] = f.xi[c];
        f.hill[k] = 0.0;
        f.hill[k + 1] = 0.0;
        if (S.mode == MODE_STEP) {
          if (oob) {  // metad.py:293-303
            for (int c = 0; c < k; ++c)
              f.F[c] = m.has_restraints
                           ? restraint_force(m.r_lower[c], m.r_upper[c], m.r_kl[c], m.r_ku[c], f.xi[c])
                           : 0.0;
          } else {
            for (int c = 0; c < k; ++c) f.F[c] = M.st.gg[flat * k + c];
            f.energy = M.st.gp ? M.st.gp[flat] : 0.0;
          }
          publish_force(M, f, ncalls);
        }
      }
    } break;
    default: break;
  }
}

// copy the finalizer results to global (by the finalizing CTA) / back to shared (by everyone)
__device__ __forceinline__ void fin_store(const Model& M, const StepShared& sh) {
  const double* src = reinterpret_cast<const double*>(&sh.fin);
  double* dst = reinterpret_cast<double*>(M.fin);
  for (int i = threadIdx.x; i < (int)(sizeof(Fin) / 8); i += BLOCK) __stcg(dst + i, src[i]);
}
__device__ __forceinline__ void fin_load(const Model& M, StepShared& sh) {
  double* dst = reinterpret_cast<double*>(&sh.fin);
  const double* src = reinterpret_cast<const double*>(M.fin);
  for (int i = threadIdx.x; i < (int)(sizeof(Fin) / 8); i += BLOCK) dst[i] = __ldcg(src + i);
  __syncthreads();
}

// per-term gradient of the term's CV components w.r.t. the atom; returns ncomp
template <class A>
__device__ __forceinline__ int term_gradient(const Model& M, const Snap& S, const Fin& f, uint32_t t,
                                             int g, uint32_t slot, V3* gv) {
  const GroupDev& G = M.grp[g];
  const CvDev& cv = M.cv[G.cv];
  switch (cv.kind) {
    case PSB_CV_ROG: {
      gv[0] = f.cvx[G.cv][0] * A::position(S, slot);
      return 1;
    }
    case PSB_CV_RMSD: {
      const double* x = f.cvx[G.cv];
      const V3 r = A::position(S, slot);
      const double P[3] = {r.x - x[9], r.y - x[10], r.z - x[11]};
      const double* q = cv.ref + 3 * (size_t)(t - cv.t0);
      const double q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
      const double w = cv.w ? __ldg(cv.w + (t - cv.t0)) : G.inv_n;
      double D[3];
      for (int a = 0; a < 3; ++a)
        D[a] = P[a] - (q0 * x[a * 3 + 0] + q1 * x[a * 3 + 1] + q2 * x[a * 3 + 2]) - w * x[12 + a];
      gv[0] = v3(x[15] * D[0], x[15] * D[1], x[15] * D[2]);
      return 1;
    }
    case PSB_CV_COORDINATION: {
      const double* gr = cv.grad + 3 * (size_t)(t - cv.t0);
      gv[0] = v3(__ldcg(gr), __ldcg(gr + 1), __ldcg(gr + 2));
      return 1;
    }
    default: {  // point CVs
      for (int j = 0; j < cv.ncomp; ++j)
        gv[j] = v3(f.gpt[g][j][0] * G.inv_n, f.gpt[g][j][1] * G.inv_n, f.gpt[g][j][2] * G.inv_n);
      return cv.ncomp;
    }
  }
}

template <class A>
__device__ __forceinline__ void add_force(const Snap& S, uint32_t slot, V3 fo) {
  A::store_force(S, slot, A::load_force(S, slot), fo);
}

// ---- the kernel -------------------------------------------------------------------------------
template <int LAYOUT, typename Real>
__global__ void __launch_bounds__(BLOCK, 2)
step_kernel(const __grid_constant__ Model M, const __grid_constant__ Snap S) {
  using A = Access<LAYOUT, Real>;
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  __shared__ StepShared sh;

  const int tid = threadIdx.x;
  const int nb = gridDim.x, b = blockIdx.x;
  const int k = M.k;
  const MethodDev& m = M.m;
  const uint32_t chunk = (M.T + nb - 1) / nb;
  const uint32_t c0 = min(M.T, (uint32_t)b * chunk), c1 = min(M.T, c0 + chunk);
  const bool metad_direct = (m.kind == PSB_METHOD_METAD && m.grid_kind == PSB_GRID_NONE);
  const bool skip_cv = (S.mode == MODE_WALKER_FINISH);  // CVs were evaluated by walker_begin

  PSB_STAMP(0);
  // ---- hill staging: kick off the first TMA bulk copies before touching any atom ------------
  long long nh = 0, h0 = 0, h1 = 0;
  const int HC = M.hill_chunk;
  double* stage_c[2];
  double* stage_h[2];
  const bool hills_here = metad_direct && (S.mode == MODE_STEP || S.mode == MODE_WALKER_BEGIN);
  if (hills_here) {
    stage_c[0] = reinterpret_cast<double*>(dyn_smem);
    stage_h[0] = stage_c[0] + (size_t)HC * k;
    stage_c[1] = stage_h[0] + HC;
    stage_h[1] = stage_c[1] + (size_t)HC * k;
    long long idx = __ldcg(M.st.idx);
    nh = idx < m.ngaussians ? idx : m.ngaussians;
    long long per = (nh + nb - 1) / nb;
    per = (per + 1) & ~1LL;  // even: keeps every slice 16-byte aligned
    h0 = min(nh, (long long)b * per);
    h1 = min(nh, h0 + per);
    if (tid == 0) {
      mbar_init(&sh.mbar[0], 1);
      mbar_init(&sh.mbar[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
      for (int s = 0; s < 2; ++s) {
        const long long a = h0 + (long long)s * HC;
        if (a >= h1) break;
        long long cnt = min((long long)HC, h1 - a);
        cnt = (cnt + 1) & ~1LL;  // arrays are padded to an even count with zero heights
        const uint32_t bc = (uint32_t)(cnt * k * 8), bh = (uint32_t)(cnt * 8);
        mbar_expect_tx(&sh.mbar[s], bc + bh);
        tma_bulk_g2s(stage_c[s], M.st.centers + a * k, bc, &sh.mbar[s]);
        tma_bulk_g2s(stage_h[s], M.st.heights + a, bh, &sh.mbar[s]);
      }
    }
  }

  // ---- pass A ---------------------------------------------------------------------------------
  if (!skip_cv) {
    for (int g = 0; g < M.ngroups; ++g) {
      const uint32_t lo = max(c0, M.grp[g].t0), hi = min(c1, M.grp[g].t1);
      if (lo >= hi) continue;
      const CvDev& cv = M.cv[M.grp[g].cv];
      const int na = nacc_for(M, S, g, 0);
      if (na == 0) continue;
      double acc[NACC];
#pragma unroll
      for (int a = 0; a < NACC; ++a) acc[a] = 0.0;
      if (is_point_kind(cv.kind)) {
        const bool with_p = (na == 6);
#pragma unroll 4
        for (uint32_t t = lo + tid; t < hi; t += BLOCK) {
          const uint32_t slot = A::slot(S, M, __ldg(M.term_tag + t));
          M.term_slot[t] = slot;
          const V3 r = A::position(S, slot);
          acc[0] += r.x; acc[1] += r.y; acc[2] += r.z;
          if (with_p) {
            const V3 p = A::momentum(S, slot);
            acc[3] += p.x; acc[4] += p.y; acc[5] += p.z;
          }
        }
      } else if (cv.kind == PSB_CV_ROG) {
#pragma unroll 4
        for (uint32_t t = lo + tid; t < hi; t += BLOCK) {
          const uint32_t slot = A::slot(S, M, __ldg(M.term_tag + t));
          M.term_slot[t] = slot;
          const V3 r = A::position(S, slot);
          acc[0] += r.x * r.x + r.y * r.y + r.z * r.z;
        }
      } else {  // RMSD: weighted centroid, plain sum, raw covariance sum_i r_i (x) Q_i
#pragma unroll 2
        for (uint32_t t = lo + tid; t < hi; t += BLOCK) {
          const uint32_t slot = A::slot(S, M, __ldg(M.term_tag + t));
          M.term_slot[t] = slot;
          const V3 r = A::position(S, slot);
          const double w = cv.w ? __ldg(cv.w + (t - cv.t0)) : M.grp[g].inv_n;
          const double* q = cv.ref + 3 * (size_t)(t - cv.t0);
          const double q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
          acc[0] += w * r.x; acc[1] += w * r.y; acc[2] += w * r.z;
          acc[3] += r.x; acc[4] += r.y; acc[5] += r.z;
          acc[6] += r.x * q0; acc[7] += r.x * q1; acc[8] += r.x * q2;
          acc[9] += r.y * q0; acc[10] += r.y * q1; acc[11] += r.y * q2;
          acc[12] += r.z * q0; acc[13] += r.z * q1; acc[14] += r.z * q2;
        }
      }
      if (nb == 1) block_reduce_store(sh, acc, na, &sh.tot[g][0], 1, false);
      else block_reduce_store(sh, acc, na, M.partials + (size_t)(g * NACC) * nb + b, nb, true);
    }

    PSB_STAMP(1);
    grid_barrier(M, sh, 0, [&] {
      PSB_STAMP(8);
      reduce_partials(M, S, sh, 0, 0);
      PSB_STAMP(9);
      if (tid == 0) {
        finalize_cvs_A(M, sh);
        PSB_STAMP(10);
        if (!M.any_rmsd) post_cv(M, S, sh);
        PSB_STAMP(11);
      }
      __syncthreads();
      fin_store(M, sh);
      PSB_STAMP(12);
    });
    PSB_STAMP(2);

    // ---- pass B: RMSD residual with the optimal rotation ------------------------------------
    if (M.any_rmsd) {
      fin_load(M, sh);
      for (int g = 0; g < M.ngroups; ++g) {
        const CvDev& cv = M.cv[M.grp[g].cv];
        if (cv.kind != PSB_CV_RMSD) continue;
        const uint32_t lo = max(c0, M.grp[g].t0), hi = min(c1, M.grp[g].t1);
        if (lo >= hi) continue;
        const double* x = sh.fin.cvx[M.grp[g].cv];
        double acc[1] = {0.0};
#pragma unroll 2
        for (uint32_t t = lo + tid; t < hi; t += BLOCK) {
          const V3 r = A::position(S, __ldcg(M.term_slot + t));
          const double P[3] = {r.x - x[9], r.y - x[10], r.z - x[11]};
          const double* q = cv.ref + 3 * (size_t)(t - cv.t0);
          double s = 0.0;
#pragma unroll
          for (int bb = 0; bb < 3; ++bb) {
            const double e = P[0] * x[0 + bb] + P[1] * x[3 + bb] + P[2] * x[6 + bb] - __ldg(q + bb);
            s += e * e;
          }
          acc[0] += s;
        }
        if (nb == 1) block_reduce_store(sh, acc, 1, &sh.tot[g][0], 1, false);
        else block_reduce_store(sh, acc, 1, M.partials + (size_t)(g * NACC) * nb + b, nb, true);
      }
      grid_barrier(M, sh, 1, [&] {
        if (nb > 1) fin_load(M, sh);  // the finalizing CTA of barrier 0 may have been another one
        reduce_partials(M, S, sh, 1, 0);
        if (tid == 0) {
          finalize_cvs_B(M, sh);
          post_cv(M, S, sh);
        }
        __syncthreads();
        fin_store(M, sh);
      });
    }
  }
  fin_load(M, sh);
  if (S.mode == MODE_EVAL) return;

  // ---- pass C: J J^T and J p over unique atoms (ABF with position-dependent Jacobians) -----------
  if (m.kind == PSB_METHOD_ABF && !M.abf_fast) {
    const uint32_t uchunk = (M.Su + nb - 1) / nb;
    const uint32_t u0 = min(M.Su, (uint32_t)b * uchunk), u1 = min(M.Su, u0 + uchunk);
    double acc[NACC];
#pragma unroll
    for (int a = 0; a < NACC; ++a) acc[a] = 0.0;
    for (uint32_t u = u0 + tid; u < u1; u += BLOCK) {
      V3 G[MAXK];
#pragma unroll
      for (int c = 0; c < MAXK; ++c) G[c] = v3(0, 0, 0);
      const uint32_t e0 = M.all_unique ? u : __ldg(M.uniq_ptr + u);
      const uint32_t e1 = M.all_unique ? u + 1 : __ldg(M.uniq_ptr + u + 1);
      uint32_t slot = 0;
      for (uint32_t e = e0; e < e1; ++e) {
        const uint32_t t = M.all_unique ? e : __ldg(M.uniq_term + e);
        const int g = __ldg(M.term_group + t);
        slot = __ldcg(M.term_slot + t);
        V3 gv[3];
        const int nc = term_gradient<A>(M, S, sh.fin, t, g, slot, gv);
        const int comp0 = M.cv[M.grp[g].cv].comp0;
        for (int j = 0; j < nc; ++j) {
#pragma unroll
          for (int c = 0; c < MAXK; ++c)
            if (c == comp0 + j) G[c] = G[c] + gv[j];
        }
      }
      const V3 p = A::momentum(S, slot);
      int o = 0;
#pragma unroll
      for (int a = 0; a < MAXK; ++a)
#pragma unroll
        for (int c = a; c < MAXK; ++c) acc[o++] += dot(G[a], G[c]);  // 10 entries
#pragma unroll
      for (int a = 0; a < MAXK; ++a) acc[10 + a] += dot(G[a], p);
    }
    if (nb == 1) block_reduce_store(sh, acc, 14, &sh.tot[VGROUP][0], 1, false);
    else block_reduce_store(sh, acc, 14, M.partials + (size_t)(VGROUP * NACC) * nb + b, nb, true);
    grid_barrier(M, sh, 2, [&] {
      reduce_partials(M, S, sh, 2, 14);
      if (tid == 0) {
        double JJt[16], Jp[MAXK];
        int o = 0;
        for (int a = 0; a < MAXK; ++a)
          for (int c = a; c < MAXK; ++c) {
            JJt[a * 4 + c] = sh.tot[VGROUP][o];
            JJt[c * 4 + a] = sh.tot[VGROUP][o];
            ++o;
          }
        for (int a = 0; a < MAXK; ++a) Jp[a] = sh.tot[VGROUP][10 + a];
        abf_update(M, S, sh.fin, JJt, Jp);
        publish_force(M, sh.fin, *M.st.ncalls + 1);
      }
      __syncthreads();
      fin_store(M, sh);
    });
    fin_load(M, sh);
  }

  // ---- pass H: direct hill sum over the TMA-staged hills ---------------------------------------
  if (hills_here) {
    double acc[NACC];
#pragma unroll
    for (int a = 0; a < NACC; ++a) acc[a] = 0.0;
    double xi[MAXK];
#pragma unroll
    for (int c = 0; c < MAXK; ++c) xi[c] = sh.fin.xi[c];
    uint32_t phase[2] = {0, 0};
    int s = 0;
    for (long long a = h0; a < h1; a += HC, s ^= 1) {
      const int cnt = (int)min((long long)HC, h1 - a);
      mbar_wait(&sh.mbar[s], phase[s]);
      phase[s] ^= 1;
      const double* cc = stage_c[s];
      const double* hh = stage_h[s];
      for (int i = tid; i < cnt; i += BLOCK) acc[0] += hill_eval(m, k, xi, cc + (size_t)i * k, hh[i], acc + 1);
      const long long nxt = a + 2LL * HC;
      if (nxt < h1) {  // refill this stage
        __syncthreads();
        if (tid == 0) {
          long long c2 = min((long long)HC, h1 - nxt);
          c2 = (c2 + 1) & ~1LL;
          const uint32_t bc = (uint32_t)(c2 * k * 8), bh = (uint32_t)(c2 * 8);
          mbar_expect_tx(&sh.mbar[s], bc + bh);
          tma_bulk_g2s(stage_c[s], M.st.centers + nxt * k, bc, &sh.mbar[s]);
          tma_bulk_g2s(stage_h[s], M.st.heights + nxt, bh, &sh.mbar[s]);
        }
      }
    }
    if (nb == 1) block_reduce_store(sh, acc, 1 + k, &sh.tot[VGROUP][0], 1, false);
    else block_reduce_store(sh, acc, 1 + k, M.partials + (size_t)(VGROUP * NACC) * nb + b, nb, true);
    grid_barrier(M, sh, 3, [&] {
      reduce_partials(M, S, sh, 3, 1 + k);
      if (tid == 0) {
        Fin& f = sh.fin;
        const double V = sh.tot[VGROUP][0];
        for (int c = 0; c < k; ++c) f.F[c] = sh.tot[VGROUP][1 + c];
        f.energy = V;
        for (int c = 0; c < k; ++c) f.hill[c] = f.xi[c];
        f.hill[k] = 0.0;
        f.hill[k + 1] = 0.0;
        if (f.deposit) {  // metad.py:219-229, 262-271 (V is the potential BEFORE this hill)
          const double h = m.well_tempered ? m.height * exp(-V / m.deltaT_kB) : m.height;
          f.hill[k] = h;
          f.hill[k + 1] = 1.0;
          if (S.mode == MODE_STEP) {
            const long long idx = *M.st.idx;
            if (idx < m.ngaussians) {
              M.st.heights[idx] = h;
              for (int c = 0; c < k; ++c) M.st.centers[idx * k + c] = f.xi[c];
            }
            *M.st.idx = idx + 1;
            // the new hill is centred on xi: zero gradient, energy grows by h
            f.energy = V + h;
          }
        }
        if (S.mode == MODE_STEP) publish_force(M, f, *M.st.ncalls + 1);
      }
      __syncthreads();
      fin_store(M, sh);
    });
    fin_load(M, sh);
  }

  if (S.mode == MODE_WALKER_BEGIN) {
    if (b == 0 && tid < k + 2) S.record_out[tid] = sh.fin.hill[tid];
    return;
  }

  // ---- pass D: add newly deposited Gaussians to the grids (metad.py:251-256) ---------------------
  const bool grid_metad = (m.kind == PSB_METHOD_METAD && m.grid_kind != PSB_GRID_NONE);
  if ((grid_metad && S.mode == MODE_STEP && sh.fin.deposit) || S.mode == MODE_WALKER_FINISH) {
    const int nrec = (S.mode == MODE_WALKER_FINISH) ? S.nwalkers : 1;
    const int rs = k + 2;
    if (grid_metad) {
      for (long long p = (long long)b * BLOCK + tid; p < m.grid_points; p += (long long)nb * BLOCK) {
        double x[MAXK];
        long long rem = p;
        for (int d = m.grid_ndim - 1; d >= 0; --d) {
          const int i = (int)(rem % m.g_shape[d]);
          rem /= m.g_shape[d];
          x[d] = mesh_coord(m.grid_kind, i, m.g_lower[d], m.g_upper[d], m.g_shape[d]);
        }
        double val = 0.0, dV[MAXK] = {0, 0, 0, 0};
        for (int r = 0; r < nrec; ++r) {
          const double* rec = (S.mode == MODE_WALKER_FINISH) ? S.records + (size_t)r * rs : sh.fin.hill;
          if (rec[k + 1] == 0.0) continue;
          val += hill_eval(m, k, x, rec, rec[k], dV);  // d/dx of the Gaussian centred on rec
        }
        if (M.st.gp) M.st.gp[p] += val;
        for (int c = 0; c < k; ++c) M.st.gg[p * k + c] += dV[c];
      }
    }
    grid_barrier(M, sh, 4, [&] {
      if (tid == 0) {
        Fin& f = sh.fin;
        long long ncalls = *M.st.ncalls + 1;
        if (S.mode == MODE_WALKER_FINISH) {
          // append the walkers' hills in rank order (identical on every rank)
          long long idx = *M.st.idx;
          for (int r = 0; r < nrec; ++r) {
            const double* rec = S.records + (size_t)r * rs;
            if (rec[k + 1] == 0.0) continue;
            if (idx < m.ngaussians) {
              M.st.heights[idx] = rec[k];
              for (int c = 0; c < k; ++c) M.st.centers[idx * k + c] = rec[c];
            }
            ++idx;
            if (!grid_metad) {  // the other walkers' new hills do push on this walker
              double dV[MAXK] = {0, 0, 0, 0};
              f.energy += hill_eval(m, k, f.xi, rec, rec[k], dV);
              for (int c = 0; c < k; ++c) f.F[c] += dV[c];
            }
          }
          *M.st.idx = idx;
        }
        if (grid_metad) {
          if (f.oob) {
            for (int c = 0; c < k; ++c)
              f.F[c] = m.has_restraints
                           ? restraint_force(m.r_lower[c], m.r_upper[c], m.r_kl[c], m.r_ku[c], f.xi[c])
                           : 0.0;
          } else {
            const long long flat = grid_flat_clamped(m, f.I);
            for (int c = 0; c < k; ++c) f.F[c] = M.st.gg[flat * k + c];
            f.energy = M.st.gp ? M.st.gp[flat] : 0.0;
          }
        }
        publish_force(M, f, ncalls);
      }
      __syncthreads();
      fin_store(M, sh);
    });
    fin_load(M, sh);
  }

  if (m.kind == PSB_METHOD_UNBIASED) return;

  PSB_STAMP(3);
  // ---- pass S: in-place force scatter ------------------------------------------------------------
  const Fin& f = sh.fin;
  const bool dense = (M.bias_dense != nullptr);
  if (M.all_unique && !dense) {
    for (int g = 0; g < M.ngroups; ++g) {
      const uint32_t lo = max(c0, M.grp[g].t0), hi = min(c1, M.grp[g].t1);
      if (lo >= hi) continue;
      const CvDev& cv = M.cv[M.grp[g].cv];
      if (is_point_kind(cv.kind)) {
        const V3 fo = v3(f.fg[g][0], f.fg[g][1], f.fg[g][2]);
#pragma unroll 4
        for (uint32_t t = lo + tid; t < hi; t += BLOCK) add_force<A>(S, __ldcg(M.term_slot + t), fo);
      } else {
        const double Fc = -f.F[cv.comp0];
#pragma unroll 2
        for (uint32_t t = lo + tid; t < hi; t += BLOCK) {
          const uint32_t slot = __ldcg(M.term_slot + t);
          V3 gv[3];
          term_gradient<A>(M, S, f, t, g, slot, gv);
          add_force<A>(S, slot, Fc * gv[0]);
        }
      }
    }
  } else {
    const uint32_t uchunk = (M.Su + nb - 1) / nb;
    const uint32_t u0 = min(M.Su, (uint32_t)b * uchunk), u1 = min(M.Su, u0 + uchunk);
    const size_t N3 = 3 * (size_t)M.natoms;
    for (uint32_t u = u0 + tid; u < u1; u += BLOCK) {
      const uint32_t e0 = M.all_unique ? u : __ldg(M.uniq_ptr + u);
      const uint32_t e1 = M.all_unique ? u + 1 : __ldg(M.uniq_ptr + u + 1);
      V3 fo = v3(0, 0, 0);
      uint32_t slot = 0;
      for (uint32_t e = e0; e < e1; ++e) {
        const uint32_t t = M.all_unique ? e : __ldg(M.uniq_term + e);
        const int g = __ldg(M.term_group + t);
        slot = __ldcg(M.term_slot + t);
        V3 gv[3];
        const int nc = term_gradient<A>(M, S, f, t, g, slot, gv);
        const int comp0 = M.cv[M.grp[g].cv].comp0;
        for (int j = 0; j < nc; ++j) {
          fo = fo + (-f.F[comp0 + j]) * gv[j];
          if (dense && M.jac_dense) {
            double* J = M.jac_dense + (size_t)(comp0 + j) * N3 + 3 * (size_t)slot;
            atomicAdd(J + 0, gv[j].x); atomicAdd(J + 1, gv[j].y); atomicAdd(J + 2, gv[j].z);
          }
        }
      }
      add_force<A>(S, slot, fo);
      if (dense) {
        double* B = M.bias_dense + 3 * (size_t)slot;
        B[0] = fo.x; B[1] = fo.y; B[2] = fo.z;
      }
    }
  }
  PSB_STAMP(4);
}

}  // namespace psb
