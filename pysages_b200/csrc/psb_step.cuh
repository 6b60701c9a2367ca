// psb_step.cuh -- the fused per-timestep bias kernel (sm_100a).
//
// One persistent (cooperative when gridDim.x > 1) launch per MD step.  The step is latency-bound
// for every configuration below ~1e6 selected atoms, so the kernel is organised around the length
// of its dependent chain, not around bytes:
//
//   prologue every CTA copies the few method-state scalars it needs into shared memory
//   pass A   gather selected atoms (tag -> slot -> position [+ image unwrap] [+ momentum]);
//            every CTA works on ONE atom group, keeps the slots of its (<= 4 per thread) terms in
//            registers and PREFETCHES their force rows, so nothing is re-read after the barrier;
//            per-group fp64 block reductions -> per-CTA partials   (colvars/core.py:193 + CV sums)
//   barrier  one grid-wide arrive/wait.  EVERY CTA then reduces the partials in the same fixed
//            order and evaluates CVs, closed-form Jacobian factors and the method update
//            redundantly (warp 0, lanes split over CVs / matrix entries / grid axes), so there
//            is no "finalizer CTA -> broadcast -> second barrier" on the critical path.  Only
//            CTA 0 writes method state; bins other CTAs still read (ABF hist/Fsum) are written
//            at the very end, after a non-blocking "readers done" counter.
//   pass B   (RMSD only) residual sum with the optimal rotation    (orientation.py:76-95)
//   pass C   (ABF with position-dependent Jacobians) J J^T and J p (abf.py:210-213)
//   pass H   (Metadynamics, direct) hill sum over TMA-staged hills (metad.py:291, 318-323)
//   pass D   (Metadynamics, grid, deposit steps) grid update       (metad.py:251-256)
//   pass S   scatter: forces[slot].xyz += -sum_c F_c dxi_c/dr      (hoomd.py:219-230, ...)
// The dense (k, 3N) Jacobian and (N, 3) bias of the reference are never materialised
// (unless Model::jac_dense / bias_dense are set for the parity tests).
#pragma once

#include <type_traits>

#include "psb_device.cuh"

namespace psb {


// The step kernel is compiled once per (engine layout, dtype, method, CV class): the finalizer
// code is executed once per launch by one warp, so it runs at instruction-fetch speed; keeping
// every instantiation small (no code for methods / CV kinds it cannot meet) is what keeps that
// path inside the instruction cache.  GENERAL = false: only point CVs (Component ... Dihedral),
// no atom shared between groups, no dense outputs, no walkers.
enum StepMethod { SM_UNBIASED = 0, SM_HARMONIC = 1, SM_METAD_DIRECT = 2, SM_ABF = 3, SM_METAD_GRID = 4, SM_COUNT = 5 };
// SC_POINT: the fast class above; SC_GENERAL: everything; SC_SLOT: the fast class with the slot-ordered
// gather / scatter (most atoms of a very large system selected), kept in its own instantiation so its
// register and code footprint stay out of the latency-critical kernels.
// SC_STREAM: the slot-ordered class with its streams moved by TMA bulk copies through a shared-memory ring, one CTA
// per SM (psb_stream.cuh); HOOMD layout.
// SC_TAGS: the slot-ordered class when the engine also hands over its slot -> tag array (HOOMD `tag`,
// psb_snapshot.tags): the group of a slot comes from a tag -> group table, pass 0 and its grid barrier disappear.
enum StepClass { SC_POINT = 0, SC_GENERAL = 1, SC_SLOT = 2, SC_STREAM = 3, SC_TAGS = 4, SC_COUNT = 5 };
#ifndef PSB_SLOT_CTAS
#define PSB_SLOT_CTAS 2
#endif
constexpr int ROWG_ROWS = 32;  // rows per thread whose group the slot-ordered gather keeps for the scatter
constexpr int SLOT_CTAS_PER_SM = PSB_SLOT_CTAS;  // streaming class: more resident warps, fewer registers

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

}  // namespace psb
#include "psb_stream.cuh"
namespace psb {

__device__ __forceinline__ long long global_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define PSB_CLK(slot)                                                                \
  do {                                                                               \
    if (S.dbg && threadIdx.x == 0) S.dbg[(size_t)sh.vb * 16 + (slot)] = clock64(); \
  } while (0)
#ifdef PSB_PROBE_PASSA  // experiment builds: clock stamps inside pass A instead of the finalizer
#define PSB_CLKA(slot)                                                               \
  do {                                                                               \
    if (S.dbg && threadIdx.x == 0) S.dbg[(size_t)sh.vb * 16 + (slot)] = clock64(); \
  } while (0)
#undef PSB_CLK
#define PSB_CLK(slot) do {} while (0)
#else
#define PSB_CLKA(slot) do {} while (0)
#endif
#ifdef PSB_PROBE_STREAM  // experiment builds: clock stamps inside the streaming passes (warp 0 of each CTA)
#define PSB_STAMPS(slot)                                                             \
  do {                                                                               \
    if (S.dbg && threadIdx.x == 0) S.dbg[(size_t)sh.vb * 16 + (slot)] = clock64(); \
  } while (0)
#undef PSB_CLK
#define PSB_CLK(slot) do {} while (0)
#else
#define PSB_STAMPS(slot) do {} while (0)
#endif
#ifdef PSB_PROBE_FIN  // experiment builds: finer clock stamps inside the ABF finalizer (replaces PSB_CLK 11..14)
#define PSB_CLKF(slot)                                                               \
  do {                                                                               \
    if (S.dbg && threadIdx.x == 0) S.dbg[(size_t)sh.vb * 16 + (slot)] = clock64(); \
  } while (0)
#undef PSB_CLK
#define PSB_CLK(slot)                                                                \
  do {                                                                               \
    if ((slot) != 11 && (slot) != 12 && S.dbg && threadIdx.x == 0) S.dbg[(size_t)sh.vb * 16 + (slot)] = clock64(); \
  } while (0)
#else
#define PSB_CLKF(slot) do {} while (0)
#endif
#define PSB_STAMP(slot)                                                              \
  do {                                                                               \
    if (S.dbg && threadIdx.x == 0) S.dbg[(size_t)sh.vb * 16 + (slot)] = global_ns(); \
  } while (0)

struct StepShared {
  int32_t need_pass_b;  // some RMSD is too close a fit for the one-pass residual (grid-uniform, see finalize_cv_A)
  int32_t vnb, vb;  // the step's own grid: CTAs of this replica and this CTA's index among them (== gridDim.x, blockIdx.x
                    // for psb_step; a sub-range of the launch for psb_batch_step)
  double red[NWARPS][NACC];
  double stage[NACC];  // this CTA's block-reduced pass-A sums, published to `partials` after griddepcontrol.wait
  double stage4[4][6]; // slot-ordered class on the OpenMM layout: the same for its (up to) four groups
  double tot[MAXG + 1][NACC];
  Fin fin;  // this CTA's copy of the finalizer results
  Pre pre;  // method-state scalars as they were when the step began
  unsigned long long epoch[NBARS];  // completed uses of each grid barrier when the step began
  uint32_t bars_used;               // barriers this launch arrived on (thread 0)
  unsigned long long seq;           // launches of this handle before this one (Snap::cta_seq), LL stamp source
  uint64_t mbar[2];
  uint64_t full[NWARPS * STREAM_MAX_STAGES];  // SC_STREAM: one "stage filled" barrier per warp and ring stage
  uint8_t rowg[ROWG_ROWS * BLOCK];            // slot-ordered classes: group + 1 of row r of thread t at [r * BLOCK + t]
  struct NetShared* net;           // ABF instantiations: force-model scratch (nullptr elsewhere)
};
// Force-model scratch of the ABF family (17 KB): only the ABF instantiations allocate it
struct NetShared {
  double nnbuf[2][PSB_MAX_WIDTH];  // force network activations (force_net_eval)
  double nndact[PSB_MAX_DENSE - 1][PSB_MAX_WIDTH];  // activation derivatives per hidden layer (gradient mode)
  ForceNet nnh;                    // force network header and (when they fit) parameters, staged after the
  double nnw[NN_STAGE];            // dependency wait by the two warps that have no rows to reduce
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

// Grid-wide barrier on monotonic counters (never reset; gridDim.x is fixed per handle).
// `epochs[id]` counts the launches that completed barrier `id`; every CTA reads it in the prologue,
// so its target is known up front: the arrival is a fire-and-forget RED and the first poll can
// already succeed (an atomic that returns the old count would cost one more L2 round trip on the
// critical path).  CTA 0 advances the epochs it used at the end of the launch -- by then every
// CTA has long read them.  Replay-safe (CUDA graphs): nothing depends on host-side counters.
__device__ __forceinline__ void red_add_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long grid_arrive(const Model& M, StepShared& sh, int id) {
  __threadfence();
  red_add_u64(&M.bars[id].arrive, 1ULL);
  sh.bars_used |= 1u << id;
  return (sh.epoch[id] + 1ULL) * (unsigned long long)sh.vnb;
}
__device__ __forceinline__ void grid_wait(const Model& M, int id, unsigned long long want) {
  while (ld_acquire_u64(&M.bars[id].arrive) < want) {}
}
__device__ __forceinline__ void grid_sync(const Model& M, StepShared& sh, int id) {
  __syncthreads();
  if (sh.vnb == 1) return;
  if (threadIdx.x == 0) grid_wait(M, id, grid_arrive(M, sh, id));
  __syncthreads();
}
// end of the launch (thread 0 of CTA 0): publish the epochs of the barriers that were used
__device__ __forceinline__ void commit_epochs(const Model& M, const StepShared& sh) {
  for (int id = 0; id < NBARS; ++id)
    if (sh.bars_used & (1u << id)) M.epochs[id] = sh.epoch[id] + 1ULL;
}

// named barrier between warp 0 (finalizer) and warp 1 (its helper): 64 threads
__device__ __forceinline__ void pair_sync(int id) {
  asm volatile("bar.sync %0, 64;" :: "r"(id) : "memory");
}

__host__ __device__ __forceinline__ bool is_point_kind(int kind) {
  return kind >= PSB_CV_COMPONENT && kind <= PSB_CV_DIHEDRAL;
}

// accumulators per group in pass `pass` (0 = A, 1 = B)
__host__ __device__ __forceinline__ int nacc_for_kind(int kind, int pass, bool momenta_sums) {
  if (pass == 0) {
    if (is_point_kind(kind)) return momenta_sums ? 6 : 3;
    if (kind == PSB_CV_ROG) return 1;
    if (kind == PSB_CV_GYRATION) return 6;
    if (kind == PSB_CV_RING) return 12;
    if (kind == PSB_CV_RMSD) return 16;
    return 0;
  }
  return kind == PSB_CV_RMSD ? 1 : 0;
}

// Deterministic reduction of the per-CTA partials into sh.tot, done by EVERY CTA in the same
// order.  pass 0 / 1: the rows of the atom groups (group g lives in CTAs cta_first[g] ..
// cta_first[g+1]-1); pass >= 2: `vrows` rows of the virtual group, spanning all CTAs.
// Eight threads share one row: 4 independent L2 loads each per round, then 3 shuffle steps.
// The row -> (destination, CTA range) lookup does not depend on the barrier, so the caller
// computes it (row_plan) BEFORE waiting and only loads + shuffles remain behind the barrier.
struct RowPlan {
  int dst, blo, bhi;  // dst < 0: no row
};
__device__ __forceinline__ RowPlan row_plan(const Model& M, int pass, int vrows, int r, int nb) {
  RowPlan p{-1, 0, -1};
  const int nrows = pass < 2 ? M.row_first[pass][M.ngroups] : vrows;
  if (nb == 1 || r >= nrows) return p;
  if (pass < 2) {
    int g = 0;
    while (r >= M.row_first[pass][g + 1]) ++g;
    p.dst = g * NACC + (r - M.row_first[pass][g]);
    p.blo = M.cta_lo[g];
    p.bhi = M.cta_hi[g];
  } else {
    p.dst = VGROUP * NACC + r;
    p.blo = 0;
    p.bhi = nb - 1;
  }
  return p;
}
__device__ __forceinline__ void reduce_rows(const Model& M, StepShared& sh, int pass, int vrows, RowPlan first) {
  const int nb = sh.vnb;
  if (nb > 1) {
    const int sub = threadIdx.x & 7, rl = threadIdx.x >> 3;
    const int nrows = pass < 2 ? M.row_first[pass][M.ngroups] : vrows;
    for (int r0 = 0; r0 < nrows; r0 += BLOCK / 8) {
      const RowPlan p = (r0 == 0) ? first : row_plan(M, pass, vrows, r0 + rl, nb);
      double v = 0.0;
      if (p.dst >= 0) {
        const double* src = M.partials + (size_t)p.dst * nb;
        // six loads in flight per lane: a group spread over <= 48 CTAs (the group-aligned split of
        // the headline shape gives 37) costs ONE L2 round trip
        for (int bb = p.blo + sub; bb <= p.bhi; bb += 48) {
          const double a0 = __ldcg(src + bb);
          const double a1 = bb + 8 <= p.bhi ? __ldcg(src + bb + 8) : 0.0;
          const double a2 = bb + 16 <= p.bhi ? __ldcg(src + bb + 16) : 0.0;
          const double a3 = bb + 24 <= p.bhi ? __ldcg(src + bb + 24) : 0.0;
          const double a4 = bb + 32 <= p.bhi ? __ldcg(src + bb + 32) : 0.0;
          const double a5 = bb + 40 <= p.bhi ? __ldcg(src + bb + 40) : 0.0;
          v += ((a0 + a1) + (a2 + a3)) + (a4 + a5);
        }
      }
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      if (p.dst >= 0 && sub == 0) (&sh.tot[0][0])[p.dst] = v;
    }
  }
  __syncthreads();
}

// reduce_rows for the slot-ordered classes, whose groups span ALL CTAs (148 .. 296 partials per row): one warp per
// row, every lane with up to WIDE_U independent loads in flight, so a row costs one L2 round trip instead of
// gridDim / 48.  Fixed summation order (lane-strided partial sums, then the xor tree): bit-identical in every CTA.
constexpr int WIDE_U = 10;
__device__ __forceinline__ void reduce_rows_wide(const Model& M, StepShared& sh) {
  const int nb = sh.vnb;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nrows = M.row_first[0][M.ngroups];
  // WIDE_R rows per warp at a time, every load of the batch issued before the first use: one L2 round trip for up to
  // NWARPS * WIDE_R = 24 rows x 320 CTAs (four groups with momenta) instead of one per row
  constexpr int WIDE_R = 3;
  for (int r0 = warp; r0 < nrows; r0 += NWARPS * WIDE_R) {
    int dst[WIDE_R];
    double v[WIDE_R];
#pragma unroll
    for (int q = 0; q < WIDE_R; ++q) {
      const int r = min(r0 + q * NWARPS, nrows - 1);
      int g = 0;
      while (r >= M.row_first[0][g + 1]) ++g;
      dst[q] = g * NACC + (r - M.row_first[0][g]);
      v[q] = 0.0;
    }
    for (int b0 = lane; b0 < nb; b0 += 32 * WIDE_U) {
      double a[WIDE_R][WIDE_U];
#pragma unroll
      for (int q = 0; q < WIDE_R; ++q)
#pragma unroll
        for (int j = 0; j < WIDE_U; ++j)
          a[q][j] = (r0 + q * NWARPS < nrows && b0 + 32 * j < nb) ? __ldcg(M.partials + (size_t)dst[q] * nb + b0 + 32 * j) : 0.0;
#pragma unroll
      for (int q = 0; q < WIDE_R; ++q)
#pragma unroll
        for (int j = 0; j < WIDE_U; ++j) v[q] += a[q][j];
    }
#pragma unroll
    for (int q = 0; q < WIDE_R; ++q) {
      const double t = warp_sum(v[q]);
      if (lane == 0 && r0 + q * NWARPS < nrows) (&sh.tot[0][0])[dst[q]] = t;
    }
  }
  __syncthreads();
}

// ---- pass-A exchange without a counter barrier ("LL": the flag travels with the data) -----------
// The counter barrier costs three dependent L2 round trips after griddepcontrol.wait: the epoch load
// (the arrival target), arrive -> poll, then the loads of the partials.  Here every partial sum is
// stored as two 8-byte words {32 data bits | 32-bit launch stamp}; an 8-byte store is single-copy
// atomic, so a reader that sees the stamp of this launch in both words has the value -- no fence, no
// counter, and the readers poll the data itself: one round trip after the last CTA has published.
// The stamp is the number of launches this handle has made (Snap::cta_seq: one counter per CTA,
// bumped at kernel entry BEFORE the dependent launch is released, so every CTA of a launch reads the
// same value without any cross-CTA synchronisation) + 1.  Every (row, CTA) entry a launch reads is
// rewritten by that launch, so a stale stamp is always an older one.
// What the barrier also ordered -- CTA 0 overwriting state scalars that other CTAs read in their
// prologue -- is ordered by the readers-done counter instead (end of the kernel).
__device__ __forceinline__ void ll_store(ulonglong2* p, double v, uint32_t stamp) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
  const unsigned long long tag = (unsigned long long)stamp << 32;
  const unsigned long long lo = (bits & 0xFFFFFFFFULL) | tag, hi = (bits >> 32) | tag;
  asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" :: "l"(p), "l"(lo), "l"(hi) : "memory");
}
__device__ __forceinline__ void ll_load(const ulonglong2* p, unsigned long long& lo, unsigned long long& hi) {
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(p) : "memory");
}
// reduce_rows for pass A over the LL buffer; the caller ends it with __syncthreads().
__device__ __forceinline__ void reduce_rows_ll(const Model& M, StepShared& sh, RowPlan first, uint32_t stamp) {
  const int nb = sh.vnb;
  const int sub = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int nrows = M.row_first[0][M.ngroups];
  for (int r0 = 0; r0 < nrows; r0 += BLOCK / 8) {
    const RowPlan p = (r0 == 0) ? first : row_plan(M, 0, 0, r0 + rl, nb);
    double v = 0.0;
    if (p.dst >= 0) {
      const ulonglong2* src = M.ll + (size_t)p.dst * nb;
      for (int bb = p.blo + sub; bb <= p.bhi; bb += 48) {
        unsigned long long lo[6], hi[6];
        bool ok;
        do {
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            lo[j] = hi[j] = (unsigned long long)stamp << 32;  // absent entries read as +0.0
            if (bb + 8 * j <= p.bhi) ll_load(src + bb + 8 * j, lo[j], hi[j]);
          }
          ok = true;
#pragma unroll
          for (int j = 0; j < 6; ++j) ok = ok && (uint32_t)(lo[j] >> 32) == stamp && (uint32_t)(hi[j] >> 32) == stamp;
        } while (!ok);
        double a[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) a[j] = __longlong_as_double((long long)((lo[j] & 0xFFFFFFFFULL) | (hi[j] << 32)));
        v += ((a[0] + a[1]) + (a[2] + a[3])) + (a[4] + a[5]);
      }
    }
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    if (p.dst >= 0 && sub == 0) (&sh.tot[0][0])[p.dst] = v;
  }
}

// Two-level LL exchange of the slot-ordered classes (hundreds of CTAs, up to 24 rows): if every CTA summed every row
// itself, gridDim^2 x rows flagged words would cross the L2 (measured: 1.6 - 2.2 us behind a 1.5 us counter barrier).
// Instead row r has ONE owner CTA (r * gridDim / rows: spread over the SMs); a warp of the owner polls the row's
// gridDim flagged partials, adds them in the fixed order of reduce_rows_wide and publishes the total as a flagged
// word; every CTA then polls the (<= 24) totals.  Two dependent L2 hops, no counter barrier, no fence.
__device__ __forceinline__ void reduce_rows_ll2(const Model& M, StepShared& sh, uint32_t stamp) {
  const int nb = sh.vnb, b = sh.vb;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nrows = M.row_first[0][M.ngroups];
  const unsigned long long want = (unsigned long long)stamp << 32;
  for (int r = warp; r < nrows; r += NWARPS) {
    if ((int)(((long long)r * nb) / nrows) != b) continue;
    int g = 0;
    while (r >= M.row_first[0][g + 1]) ++g;
    const int dst = g * NACC + (r - M.row_first[0][g]);
    const ulonglong2* src = M.ll + (size_t)dst * nb;
    double v = 0.0;
    for (int b0 = lane; b0 < nb; b0 += 32 * WIDE_U) {
      unsigned long long lo[WIDE_U], hi[WIDE_U];
      bool ok;
      do {
#pragma unroll
        for (int j = 0; j < WIDE_U; ++j) {
          lo[j] = hi[j] = want;  // absent entries read as +0.0
          if (b0 + 32 * j < nb) ll_load(src + b0 + 32 * j, lo[j], hi[j]);
        }
        ok = true;
#pragma unroll
        for (int j = 0; j < WIDE_U; ++j) ok = ok && (uint32_t)(lo[j] >> 32) == stamp && (uint32_t)(hi[j] >> 32) == stamp;
      } while (!ok);
#pragma unroll
      for (int j = 0; j < WIDE_U; ++j) v += __longlong_as_double((long long)((lo[j] & 0xFFFFFFFFULL) | (hi[j] << 32)));
    }
    v = warp_sum(v);
    if (lane == 0) ll_store(M.ll_tot + dst, v, stamp);
  }
  if ((int)threadIdx.x < nrows) {
    const int r = threadIdx.x;
    int g = 0;
    while (r >= M.row_first[0][g + 1]) ++g;
    const int dst = g * NACC + (r - M.row_first[0][g]);
    unsigned long long lo, hi;
    do {
      ll_load(M.ll_tot + dst, lo, hi);
    } while ((uint32_t)(lo >> 32) != stamp || (uint32_t)(hi >> 32) != stamp);
    (&sh.tot[0][0])[dst] = __longlong_as_double((long long)((lo & 0xFFFFFFFFULL) | (hi << 32)));
  }
}

// ---- finalizers (warp 0 of every CTA) ----------------------------------------------------------
__device__ __forceinline__ V3 tot3(const StepShared& sh, int g, int o) {
  return v3(sh.tot[g][o], sh.tot[g][o + 1], sh.tot[g][o + 2]);
}
__device__ __forceinline__ void put3(double* dst, V3 v) {
  dst[0] = v.x; dst[1] = v.y; dst[2] = v.z;
}

// CV `c`: value and Jacobian factors after pass A (one lane per CV).  Only the rows
// gpt[g][j < ncomp] of a CV's groups are ever read, and every case writes them completely.
template <bool GENERAL>
static __device__ __forceinline__ void finalize_cv_A(const Model& M, StepShared& sh, int c) {
  Fin& f = sh.fin;
  const CvDev& cv = M.cv[c];
  const int g0 = cv.g0;
  const int kind = cv.kind;
  if (is_point_kind(kind)) {
    // barycentres of the (up to four) groups, in registers
    const V3 p0 = M.grp[g0].inv_n * tot3(sh, g0, 0);
    if (kind == PSB_CV_DISTANCE) {  // coordinates.py:91-109
      const V3 p1 = M.grp[g0 + 1].inv_n * tot3(sh, g0 + 1, 0);
      V3 a, b;
      f.xi[cv.comp0] = distance_cv(p0, p1, &a, &b);
      put3(f.gpt[g0][0], a);
      put3(f.gpt[g0 + 1][0], b);
    } else if (kind == PSB_CV_COMPONENT) {  // coordinates.py:28,71
      f.xi[cv.comp0] = comp(p0, cv.axis);
      put3(f.gpt[g0][0], v3(cv.axis == 0 ? 1.0 : 0.0, cv.axis == 1 ? 1.0 : 0.0, cv.axis == 2 ? 1.0 : 0.0));
    } else if (kind == PSB_CV_DISPLACEMENT) {  // coordinates.py:148: r2 - r1
      const V3 p1 = M.grp[g0 + 1].inv_n * tot3(sh, g0 + 1, 0);
      const V3 d = p1 - p0;
      f.xi[cv.comp0 + 0] = d.x; f.xi[cv.comp0 + 1] = d.y; f.xi[cv.comp0 + 2] = d.z;
      for (int q = 0; q < 3; ++q) {
        const V3 e = v3(q == 0 ? 1.0 : 0.0, q == 1 ? 1.0 : 0.0, q == 2 ? 1.0 : 0.0);
        put3(f.gpt[g0][q], (-1.0) * e);
        put3(f.gpt[g0 + 1][q], e);
      }
    } else {
      const V3 p1 = M.grp[g0 + 1].inv_n * tot3(sh, g0 + 1, 0);
      const V3 p2 = M.grp[g0 + 2].inv_n * tot3(sh, g0 + 2, 0);
      if (kind == PSB_CV_ANGLE) {
        V3 a, b, d;
        f.xi[cv.comp0] = angle_cv(p0, p1, p2, &a, &b, &d);
        put3(f.gpt[g0][0], a); put3(f.gpt[g0 + 1][0], b); put3(f.gpt[g0 + 2][0], d);
      } else {  // PSB_CV_DIHEDRAL
        const V3 p3 = M.grp[g0 + 3].inv_n * tot3(sh, g0 + 3, 0);
        V3 a, b, d, e;
        f.xi[cv.comp0] = dihedral_cv(p0, p1, p2, p3, &a, &b, &d, &e);
        put3(f.gpt[g0][0], a); put3(f.gpt[g0 + 1][0], b);
        put3(f.gpt[g0 + 2][0], d); put3(f.gpt[g0 + 3][0], e);
      }
    }
    return;
  }
  if (!GENERAL) {
    // the slot-ordered class also takes RadiusOfGyration groups (shape.py:52-57: sum r.r / n; d/dr_i = 2 r_i / n)
    if (kind == PSB_CV_ROG) {
      f.xi[cv.comp0] = sh.tot[g0][0] * M.grp[g0].inv_n;
      f.cvx[c][0] = 2.0 * M.grp[g0].inv_n;
    }
    return;
  }
  switch (kind) {
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
      // The residual of orientation.py:76-95 without a second pass over the atoms: with U orthogonal,
      //   sum_i |P_i U - Q_i|^2 = sum |P_i|^2 + sum |Q_i|^2 - 2 sum_ab U_ab C_ab,
      //   sum |P_i|^2 = sum |r_i|^2 - 2 rbar . sum r_i + n |rbar|^2.
      // It is a difference of O(n R^2) terms whose tree-reduced fp64 sums carry a few 1e-16 of relative error
      // each, so it is only taken when it keeps >= 1e-3 of their magnitude (RMSD >~ 5 % of the structure's
      // size): the RMSD then keeps ~1e-13 relative.  That margin is needed downstream, not here: a metadynamics
      // landscape of 1e4 narrow hills turns 1e-11 on xi into 1e-8 on the bias force (measured on cfg4, whose
      // RMSD of 0.2 on a 19-unit globule therefore takes pass B, as does the reference's own ci2 case, 4.9e-4
      // on a 10-unit structure).  Every CTA holds bit-identical sums: the decision is uniform over the grid.
      {
        const double sr2 = sh.tot[g0][15];
        const double rb_sr = rbar.x * sumr.x + rbar.y * sumr.y + rbar.z * sumr.z;
        const double rb2 = rbar.x * rbar.x + rbar.y * rbar.y + rbar.z * rbar.z;
        double uc = 0.0;
        for (int a = 0; a < 9; ++a) uc += x[a] * C[a];
        const double res = ((sr2 - 2.0 * rb_sr) + n * rb2) + cv.gq - 2.0 * uc;
        const double big = sr2 + 2.0 * fabs(rb_sr) + n * rb2 + cv.gq + 2.0 * fabs(uc);
        if (res > 1e-3 * big) {
          const double xi = sqrt(res * M.grp[g0].inv_n);
          f.xi[cv.comp0] = xi;
          x[15] = M.grp[g0].inv_n / xi;  // 1 / (n rmsd)
          x[16] = 0.0;
        } else {
          x[16] = 1.0;
        }
      }
    } break;
    case PSB_CV_ERMSD:  // value and gradient come from ermsd_kernel, like the pair kernel's
    case PSB_CV_COORDINATION: {
      f.xi[cv.comp0] = sh.tot[g0][0];
    } break;
    case PSB_CV_GYRATION: {  // shape.py:85-344
      const double inv_n = M.grp[g0].inv_n;
      double gt[6];
      for (int a = 0; a < 6; ++a) gt[a] = sh.tot[g0][a] * inv_n;
      double* x = f.cvx[c];
      f.xi[cv.comp0] = gyration_cv(cv.axis, cv.nn, cv.mm, gt, x);
      for (int a = 0; a < 9; ++a) x[a] *= 2.0 * inv_n;  // d xi / d r_i = x r_i
    } break;
    case PSB_CV_RING: {  // angles.py:131-281
      f.xi[cv.comp0] = ring_cv(cv.axis, cv.t1 - cv.t0, sh.tot[g0], f.cvx[c]);
    } break;
    default: break;
  }
}

// coordination CVs: the pair kernel left per-CTA partial sums; warp 0 adds them up (fixed order)
__device__ __forceinline__ void reduce_coordination(const Model& M, StepShared& sh, int lane) {
  for (int c = 0; c < M.ncv; ++c) {
    if (M.cv[c].kind != PSB_CV_COORDINATION && M.cv[c].kind != PSB_CV_ERMSD) continue;
    double v = 0.0;
    for (int i = lane; i < M.cv[c].n_part; i += 32) v += __ldcg(M.cv[c].xi_part + i);
    v = warp_sum(v);
    if (lane == 0) sh.tot[M.cv[c].g0][0] = v;
  }
  __syncwarp();
}

// flat row-major offset with JAX's clamp-on-gather semantics
__device__ __forceinline__ long long grid_flat_clamped(const MethodDev& m, const uint32_t* I) {
  long long o = 0;
  for (int d = 0; d < m.grid_ndim; ++d) {
    uint32_t i = I[d] < (uint32_t)m.g_shape[d] ? I[d] : (uint32_t)m.g_shape[d] - 1;
    o = o * m.g_shape[d] + i;
  }
  return o;
}

// grid bin of xi, one lane per axis (grids.py:109-158); lane 0 then derives the out-of-bounds
// flag (any index == shape, grids.py:120) and the clamped flat offset.
__device__ __forceinline__ void grid_bin_warp(const MethodDev& m, Fin& f, int lane) {
  if (lane < m.grid_ndim)
    f.I[lane] = grid_index_1d(m.grid_kind, f.xi[lane], m.g_lower[lane], m.g_upper[lane], m.g_shape[lane],
                              m.g_h[lane], m.g_inv_h[lane]);
  __syncwarp();
  if (lane == 0) {
    bool oob = false;
    for (int d = 0; d < m.grid_ndim; ++d) oob |= (f.I[d] == (uint32_t)m.g_shape[d]);
    f.oob = oob ? 1 : 0;
    f.flat = grid_flat_clamped(m, f.I);
  }
  __syncwarp();
}

// F is final: fg[g] = -sum_j F[comp0+j] dxi_{comp0+j}/dpoint_g / n_g (the -J^T F of
// harmonic_bias.py:127, metad.py:200, abf.py:223 restricted to one group; one lane per group),
// and CTA 0 publishes the generalized force, the bias energy and the call counter.
__device__ __forceinline__ void finish_force(const Model& M, StepShared& sh, int lane, bool cta0) {
  Fin& f = sh.fin;
  __syncwarp();
  if (lane < M.ngroups) {
    const int g = lane;
    const CvDev& cv = M.cv[M.grp[g].cv];
    double v[3] = {0, 0, 0};
    if (is_point_kind(cv.kind)) {
      const double inv_n = M.grp[g].inv_n;
#pragma unroll 1
      for (int j = 0; j < cv.ncomp; ++j) {
        const double Fj = f.F[cv.comp0 + j] * inv_n;
        v[0] -= Fj * f.gpt[g][j][0]; v[1] -= Fj * f.gpt[g][j][1]; v[2] -= Fj * f.gpt[g][j][2];
      }
    }
    f.fg[g][0] = v[0]; f.fg[g][1] = v[1]; f.fg[g][2] = v[2];
  }
  if (cta0) {
    if (lane < M.k) M.st.F[lane] = f.F[lane];
    if (lane == 0) {
      *M.st.energy = f.energy;
      if (!M.ll_reduce) *M.st.ncalls = sh.pre.ncalls + 1;  // LL exchange: stored after the readers-done wait
    }
  }
}

// funn.py:279-284 / cff.py:328-332 predict_force, one warp: F = std * mlp(scale(float32(xi))) + mean with
// scale = approxfun/core.py:99-104 and the dense / activation stack of ml/models.py:44-82.  Lane j
// owns output unit j (+32, ...) of every layer; activations ping-pong through shared memory.
// Returns component `lane` of the force.  Kept out of line: plain ABF never reaches it.
static __device__ __noinline__ double force_net_eval(const ForceNet* nn, const double* P, const MethodDev& m,
                                                     const double* xi, double* buf, double* dact, int lane) {
  const int nd = nn->n_dense;
  double* cur = buf;
  double* nxt = buf + PSB_MAX_WIDTH;
  int in = nn->widths[0];
  if (lane < in) {
    const double x = (double)(float)xi[lane];  // funn.py:283: f32(x), promoted back by the fp64 grid bounds
    cur[lane] = (x - m.g_lower[lane]) * 2.0 / (m.g_upper[lane] - m.g_lower[lane]) - 1.0;
  }
  __syncwarp();
  for (int l = 0; l < nd; ++l) {
    const int out = nn->widths[l + 1];
    const double* W = P + nn->w_off[l];
    const double* b = W + (size_t)in * out;
    const double om = l == 0 ? nn->omega0 : nn->omega;
    for (int j = lane; j < out; j += 32) {
      double acc = 0.0;
#pragma unroll 4
      for (int i = 0; i < in; ++i) acc = fma(cur[i], W[(size_t)i * out + j], acc);
      acc += b[j];
      if (l + 1 < nd) {
        if (nn->activation == PSB_ACT_SIN) {
          double sn, cs;
          sincos(om * acc, &sn, &cs);
          if (dact) dact[l * PSB_MAX_WIDTH + j] = om * cs;
          acc = sn;
        } else {
          acc = tanh(acc);
          if (dact) dact[l * PSB_MAX_WIDTH + j] = 1.0 - acc * acc;
        }
      }
      nxt[j] = acc;
    }
    __syncwarp();
    double* t = cur; cur = nxt; nxt = t;
    in = out;
  }
  if (!nn->gradient) return lane < in ? nn->std[lane] * cur[lane] + nn->mean[lane] : 0.0;
  // ann.py:240-248: gradient of the summed outputs with respect to xi (through `scale`), by back-propagation:
  // delta over the units of a layer, lane i owns unit i (+32, ...)
  for (int j = lane; j < in; j += 32) cur[j] = 1.0;
  __syncwarp();
  for (int l = nd - 1; l >= 0; --l) {
    const int out = nn->widths[l + 1], inw = nn->widths[l];
    const double* W = P + nn->w_off[l];
    for (int i = lane; i < inw; i += 32) {
      double acc = 0.0;
#pragma unroll 4
      for (int j = 0; j < out; ++j) acc = fma(W[(size_t)i * out + j], cur[j], acc);
      if (l > 0) acc *= dact[(l - 1) * PSB_MAX_WIDTH + i];
      nxt[i] = acc;
    }
    __syncwarp();
    double* t = cur; cur = nxt; nxt = t;
  }
  const int k0 = nn->widths[0];
  return lane < k0 ? nn->std[lane] * cur[lane] * 2.0 / (m.g_upper[lane] - m.g_lower[lane]) : 0.0;
}

// spectral_abf.py:245-246 interpolate_force, one warp: gradient of the fitted Fourier / monomial series at
// xi (approxfun/core.py:170-201, 311-336), lanes striding over the K terms; returns component `lane`.
static __device__ __noinline__ double force_series_eval(const ForceSeries* fs, const MethodDev& m, const double* xi,
                                                        double* pw, int lane) {
  const int K = fs->K, dims = fs->dims;
  const int32_t* ns = fs->exponents();
  const double* val = fs->values();
  double x[2] = {0.0, 0.0}, sd[2] = {0.0, 0.0};
  for (int d = 0; d < dims; ++d) {
    const double size = m.g_upper[d] - m.g_lower[d];
    x[d] = (xi[d] - m.g_lower[d]) * 2.0 / size - 1.0;  // approxfun/core.py:99-104
    sd[d] = 2.0 * (fs->kind == PSB_SERIES_FOURIER ? 3.141592653589793 : 1.0) / size;
  }
  if (fs->kind == PSB_SERIES_MONOMIAL) {  // power tables pw[d * 64 + j] = x_d^j, j < 64 (binary exponentiation)
    for (int j = lane; j < 128; j += 32) {
      double base = x[j >> 6], r = 1.0;
      for (int e = j & 63; e > 0; e >>= 1) {
        if (e & 1) r *= base;
        base *= base;
      }
      pw[j] = r;
    }
    __syncwarp();
  }
  double acc[2] = {0.0, 0.0};
  for (int k = lane; k < K; k += 32) {
    const int n0 = ns[k * dims], n1 = dims > 1 ? ns[k * dims + 1] : 0;
    if (fs->kind == PSB_SERIES_FOURIER) {  // imag(exp(-i t)) = -sin t multiplies c_k, real = cos t multiplies c_{K+k}
      double sn, cs;
      sincospi((double)n0 * x[0] + (double)n1 * x[1], &sn, &cs);
      const double w = cs * val[1 + K + k] - sn * val[1 + k];
      acc[0] += sd[0] * (double)n0 * w;
      acc[1] += sd[1] * (double)n1 * w;
    } else {  // x^n and x^max(n - 1, 0) from the tables (0 <= n < 64, checked by psb_set_force_series)
      const double p0 = pw[n0], p0m = pw[max(n0 - 1, 0)], p1 = pw[64 + n1], p1m = pw[64 + max(n1 - 1, 0)];
      const double c = val[1 + k];
      acc[0] += sd[0] * (double)n0 * p0m * p1 * c;
      acc[1] += sd[1] * (double)n1 * p1m * p0 * c;
    }
  }
  const double a0 = warp_sum(acc[0]), a1 = warp_sum(acc[1]);
  return lane < dims ? val[0] * (lane == 0 ? a0 : a1) : 0.0;
}

// Everything of abf.py:213-258 that depends on xi alone: the bin of xi, the gathered hist[I] / Fsum[I]
// (gathers clamp, JAX default), the new count, max(N, count) as a double and the restraint force
// outside the grid.  Point-CV kernels run it on warp 1 while warp 0 does J p / solve.
__device__ __forceinline__ void abf_bin_gather(const Model& M, StepShared& sh, int lane) {
  Fin& f = sh.fin;
  const MethodDev& m = M.m;
  grid_bin_warp(m, f, lane);  // ends with __syncwarp
  if (lane < M.k) {
    const bool oob = f.oob != 0;
    f.Fsum_old[lane] = __ldcg(M.st.Fsum + f.flat * M.k + lane);
    f.rforce[lane] = (m.has_restraints && oob)
                         ? restraint_force(m.r_lower[lane], m.r_upper[lane], m.r_kl[lane], m.r_ku[lane], f.xi[lane])
                         : 0.0;
    if (lane == 0) {
      const uint32_t h = __ldcg(M.st.hist + f.flat) + (oob ? 0u : 1u);  // .at[I].add drops out-of-range updates
      const unsigned long long hh = h, nmin = (unsigned long long)m.abf_N;
      f.hist_new = h;
      f.in_range = oob ? 0 : 1;
      f.den = (double)(nmin > hh ? nmin : hh);
      f.inv_den = 1.0 / f.den;
    }
  }
  __syncwarp();
}
__device__ __forceinline__ void abf_helper(const Model& M, StepShared& sh, int lane) {
  pair_sync(1);  // xi is final
  abf_bin_gather(M, sh, lane);
  pair_sync(2);  // bin + gathered values are in shared memory
}

// abf.py:210-225 once f.Jp (and f.JJt, or f.Wp) are known (warp 0, all lanes).  Every CTA derives the
// new histogram count and force sum of bin I from the values the step began with; CTA 0 stores them
// after the last reader is done (end of the kernel).
__device__ __forceinline__ void abf_finish(const Model& M, const Snap& S, StepShared& sh, int lane, bool cta0,
                                           bool helped, bool have_Wp) {
  Fin& f = sh.fin;
  const MethodDev& m = M.m;
  const int k = M.k;
  __syncwarp();
  if (lane == 0 && !have_Wp) {  // utils/core.py:120-136
    double Wp[MAXK] = {};
    if (m.use_pinv) {
      pinv_sym_solve(k, f.JJt, f.Jp, Wp, 10.0 * (double)(3 * M.natoms) * 2.220446049250313e-16);
    } else if (!spd_solve(k, f.JJt, f.Jp, Wp)) {
      for (int c = 0; c < k; ++c) Wp[c] = nan("");
    }
    for (int c = 0; c < MAXK; ++c) f.Wp[c] = Wp[c];
  }
  PSB_CLK(12);
  if (helped) {
    __syncwarp();
    pair_sync(2);  // warp 1 has binned xi and gathered hist[I], Fsum[I]
  } else {
    abf_bin_gather(M, sh, lane);
  }
  PSB_CLK(13);
  const bool oob = f.oob != 0;
  // Force model, if the handle has one (Model::fm_mode: plain ABF pays one shared-memory load here; kept out of line).
  //   network (funn.py:190,286-296; cff.py:328-332): past predict_after, restraints first;
  //   network gradient, histogram-only variant (ann.py:240-256): inside the grid past predict_after, else zero;
  //   series (spectral_abf.py:186,245-266): past fit_threshold inside the grid; outside restraints or nothing.
  const int fm = M.fm_mode;
  const bool hist_only = (fm & FM_HIST_ONLY) != 0;
  bool use_model = false, zero_force = hist_only;
  double f_model = 0.0;
  if (fm != 0) {
    const long long ncalls = sh.pre.ncalls + 1;
    if (fm & FM_NET) {
      use_model = !((m.has_restraints || hist_only) && oob) && ncalls > sh.net->nnh.predict_after;
      if (use_model)
        f_model = force_net_eval(&sh.net->nnh, M.nn_np > 0 ? sh.net->nnw : M.nn->params(), m, f.xi, &sh.net->nnbuf[0][0],
                                 sh.net->nnh.gradient ? &sh.net->nndact[0][0] : nullptr, lane);
    } else if (fm & FM_SERIES) {
      const ForceSeries* ser = M.fs_staged ? reinterpret_cast<const ForceSeries*>(&sh.net->nnh) : M.fs;
      zero_force = ser->oob_zero && oob && !m.has_restraints;
      use_model = !oob && ncalls > ser->predict_after;
      if (use_model) f_model = force_series_eval(ser, m, f.xi, &sh.net->nnbuf[0][0], lane);
    }
  }
  if (lane < k) {
    const int c = lane;
    const double Wp_now = f.Wp[c], Wp_1 = sh.pre.Wp[c];
    double fs = f.Fsum_old[c];
    if (!oob && !hist_only) {
      // (x / dt and x / den of abf.py:213,244 as products with the reciprocals: one rounding more, ~1e-16 relative;
      // both divisions were dependent ~200-cycle steps of the single-warp chain)
      const double dWp_dt = (1.5 * Wp_now - 2.0 * Wp_1 + 0.5 * sh.pre.Wp_[c]) * S.inv_dt;
      fs += m.force_unit_factor * dWp_dt + sh.pre.force[c];
    }
    const double fc = use_model ? f_model
                      : (zero_force ? 0.0 : ((m.has_restraints && oob) ? f.rforce[c] : fs * f.inv_den));
    f.F[c] = fc;
    f.Fsum_new[c] = fs;
    if (cta0 && !M.ll_reduce) {  // scalars nobody reads after the first barrier (LL exchange: stored at the end)
      M.st.force[c] = fc;
      M.st.Wp_[c] = Wp_1;
      M.st.Wp[c] = Wp_now;
    }
  }
  if (lane == 0) f.energy = 0.0;
  PSB_CLK(14);
  finish_force(M, sh, lane, cta0);
}

// Value/gradient of one Gaussian at xi (utils/core.py:113-117, metad.py:318-323)
template <int KM = MAXK>  // KM: compile-time bound of k (the direct hill sum picks 2 / 4 / MAXK outside its loop)
__device__ __forceinline__ double hill_eval(const MethodDev& m, int k, const double* xi,
                                            const double* centre, double height, double* dV) {
  double d[KM], q = 0.0;
#pragma unroll
  for (int c = 0; c < KM; ++c)
    if (c < k) {
      d[c] = wrap_period(xi[c] - centre[c], m.periods[c]);
      const double z = d[c] / m.sigma[c];
      q += z * z;
    }
  const double e = height * exp(-q / 2.0);
#pragma unroll
  for (int c = 0; c < KM; ++c)
    if (c < k) dV[c] += -e * d[c] / (m.sigma[c] * m.sigma[c]);
  return e;
}

// metad.py:293-312 in grid mode: generalized force from the gradient grid (nearest bin) or, out of
// bounds, from the restraints / zeros.  One lane per component.
__device__ __forceinline__ void metad_grid_force(const Model& M, Fin& f, int lane) {
  const MethodDev& m = M.m;
  const int k = M.k;
  if (lane < k) {
    const int c = lane;
    if (f.oob) {
      f.F[c] = m.has_restraints ? restraint_force(m.r_lower[c], m.r_upper[c], m.r_kl[c], m.r_ku[c], f.xi[c]) : 0.0;
    } else {
      f.F[c] = __ldcg(M.st.gg + f.flat * k + c);
    }
  }
  if (lane == 0) f.energy = (!f.oob && M.st.gp) ? __ldcg(M.st.gp + f.flat) : 0.0;
}

// Direct hill sum over one staged chunk for k > KA components (kept out of line: the common k <= 4 loops then keep
// their registers; metad.py:318-323)
static __device__ __noinline__ void hill_chunk_wide(const MethodDev& m, int k, const double* xi, const double* cc,
                                                    const double* hh, int cnt, int tid, double* acc) {
  for (int i = tid; i < cnt; i += BLOCK) acc[0] += hill_eval<MAXK>(m, k, xi, cc + (size_t)i * k, hh[i], acc + 1);
}

// Everything that can be decided once xi is complete (warp 0, all lanes).
template <int METHOD, bool GENERAL>
static __device__ __forceinline__ void post_cv(const Model& M, const Snap& S, StepShared& sh, int lane, bool cta0) {
  Fin& f = sh.fin;
  const MethodDev& m = M.m;
  const int k = M.k;
  __syncwarp();
  if (cta0 && lane < k) M.st.xi[lane] = f.xi[lane];
  if (lane == 0) {
    f.deposit = 0;
    f.oob = 0;
    f.energy = 0.0;
  }
  if (lane < MAXK) f.F[lane] = 0.0;
  __syncwarp();
  if (S.mode == MODE_EVAL) return;
  const long long ncalls = sh.pre.ncalls + 1;
  switch (METHOD) {
    case SM_UNBIASED: finish_force(M, sh, lane, cta0); break;
    case SM_HARMONIC: {  // harmonic_bias.py:124-128 (no periodic wrap: quirk kept)
      if (lane < k) {
        double s = 0.0;
        for (int b = 0; b < k; ++b) s += m.kspring[lane * MAXK + b] * (f.xi[b] - m.center[b]);
        f.F[lane] = s;
      }
      __syncwarp();
      if (lane == 0) {
        double e = 0.0;
        for (int a = 0; a < k; ++a) e += 0.5 * (f.xi[a] - m.center[a]) * f.F[a];
        f.energy = e;
      }
      finish_force(M, sh, lane, cta0);
    } break;
    case SM_ABF: {
      if (GENERAL && !M.abf_fast) break;  // needs pass C
      if (!GENERAL) {  // point-CV kernels: warp 1 bins xi and gathers the histogram meanwhile
        __syncwarp();
        pair_sync(1);
      }
      PSB_CLKF(11);
      if (M.fm_mode & FM_HIST_ONLY) {  // ann.py:153-167: no momenta, no W p
        if (lane < MAXK) f.Wp[lane] = 0.0;
        abf_finish(M, S, sh, lane, cta0, !GENERAL, true);
        break;
      }
      // J J^T and J p from per-group sums: point-CV Jacobians are constant within a group.
      if (M.jjt_const) {
        // Component / Distance / Displacement CVs without atoms shared between CVs: J J^T is a constant
        // of the index tables (|d distance / d centroid| = 1); its inverse comes from psb_create and
        // solve(J J^T, J p) (abf.py:211, utils/core.py:128-134) is a k x k matrix-vector product.
        if (lane < MAXK) {
          const int a = lane;
          double s = 0.0;
          if (a < k) {
            const CvDev& ca = M.cv[M.comp_cv[a]];
            const int ja = a - ca.comp0;
#pragma unroll 1
            for (int g = ca.g0; g < ca.g0 + ca.ngroups; ++g) {
              const V3 ps = tot3(sh, g, 3);
              const double* ga = f.gpt[g][ja];
              s += (ga[0] * ps.x + ga[1] * ps.y + ga[2] * ps.z) * M.grp[g].inv_n;
            }
          }
          f.Jp[a] = s;
        }
        __syncwarp();
        PSB_CLKF(12);
        if (lane < MAXK) {
          double w = 0.0;
          if (lane < k)
            for (int bq = 0; bq < k; ++bq) w += M.jjt_inv[lane * KA + bq] * f.Jp[bq];
          f.Wp[lane] = w;
        }
        PSB_CLK(11);
        abf_finish(M, S, sh, lane, cta0, !GENERAL, true);
        break;
      }
      // lanes 0..15: entry (a, b) of J J^T; lanes 16..19: entry a of J p.
      static_assert(KA == 4 && 16 + MAXK <= 32, "lanes 0..15: J J^T entries, lanes 16..16+MAXK-1: J p");
      if (lane < 16) {
        const int a = lane >> 2, bq = lane & 3;
        double s = 0.0;
        if (a < k && bq < k) {
          const int cia = M.comp_cv[a], cib = M.comp_cv[bq];
          const CvDev& ca = M.cv[cia];
          const CvDev& cb = M.cv[cib];
          const int ja = a - ca.comp0, jb = bq - cb.comp0;
          if (cia == cib) {  // atoms of one group: (count / n_g^2) * dxi_a/dpoint . dxi_b/dpoint
#pragma unroll 1
            for (int g = ca.g0; g < ca.g0 + ca.ngroups; ++g) {
              const double* ga = f.gpt[g][ja];
              const double* gb = f.gpt[g][jb];
              s += M.grp[g].self_coef * (ga[0] * gb[0] + ga[1] * gb[1] + ga[2] * gb[2]);
            }
          }
          if (M.cross_overlap) {  // atoms shared between different groups
#pragma unroll 1
            for (int g = ca.g0; g < ca.g0 + ca.ngroups; ++g)
#pragma unroll 1
              for (int h = cb.g0; h < cb.g0 + cb.ngroups; ++h) {
                const uint32_t ovgh = (g == h) ? 0u : __ldg(M.ov + g * MAXG + h);
                if (ovgh == 0) continue;
                const double* ga = f.gpt[g][ja];
                const double* gb = f.gpt[h][jb];
                s += (double)ovgh * (ga[0] * gb[0] + ga[1] * gb[1] + ga[2] * gb[2]) * M.grp[g].inv_n *
                     M.grp[h].inv_n;
              }
          }
        }
        f.JJt[lane] = s;
      } else if (lane < 16 + MAXK) {
        const int a = lane - 16;
        double s = 0.0;
        if (a < k) {
          const CvDev& ca = M.cv[M.comp_cv[a]];
          const int ja = a - ca.comp0;
#pragma unroll 1
          for (int g = ca.g0; g < ca.g0 + ca.ngroups; ++g) {
            const V3 ps = tot3(sh, g, 3);
            const double* ga = f.gpt[g][ja];
            s += (ga[0] * ps.x + ga[1] * ps.y + ga[2] * ps.z) * M.grp[g].inv_n;
          }
        }
        f.Jp[a] = s;
      }
      PSB_CLK(11);
      abf_finish(M, S, sh, lane, cta0, !GENERAL, false);
    } break;
    case SM_METAD_DIRECT: {
      const bool in_dep = (ncalls > 1) && (ncalls % m.stride == 1);  // metad.py:192-193
      if (lane == 0) f.deposit = in_dep ? 1 : 0;  // heights need V: decided after pass H
    } break;
    case SM_METAD_GRID: {
      const bool in_dep = (ncalls > 1) && (ncalls % m.stride == 1);  // metad.py:192-193
      grid_bin_warp(m, f, lane);
      if (lane < k) f.hill[lane] = f.xi[lane];
      if (in_dep && !f.oob) {  // metad.py:258-271
        if (lane == 0) {
          double h = m.height;
          if (m.well_tempered) h = m.height * exp(-__ldcg(M.st.gp + f.flat) / m.deltaT_kB);  // metad.py:223-229
          f.hill[k] = h;
          f.hill[k + 1] = 1.0;
          f.deposit = 1;
          if (S.mode == MODE_STEP && cta0) {
            const long long idx = sh.pre.idx;
            if (idx < m.ngaussians) {
              M.st.heights[idx] = h;
              for (int c = 0; c < k; ++c) M.st.centers[idx * k + c] = f.xi[c];
            }
            *M.st.idx = idx + 1;
          }
        }
      } else {
        if (lane == 0) {
          f.hill[k] = 0.0;
          f.hill[k + 1] = 0.0;
        }
        if (S.mode == MODE_STEP) {
          metad_grid_force(M, f, lane);
          finish_force(M, sh, lane, cta0);
        }
      }
    } break;
    default: break;
  }
  __syncwarp();
}

// RMSD value after pass B (one lane per CV).
__device__ __forceinline__ void finalize_cv_B(const Model& M, StepShared& sh, int c) {
  Fin& f = sh.fin;
  const CvDev& cv = M.cv[c];
  if (cv.kind != PSB_CV_RMSD || f.cvx[c][16] == 0.0) return;  // (value already final from pass A)
  const double inv_n = M.grp[cv.g0].inv_n;
  const double xi = sqrt(sh.tot[cv.g0][0] * inv_n);  // orientation.py:94
  f.xi[cv.comp0] = xi;
  f.cvx[c][15] = inv_n / xi;  // 1 / (n rmsd)
}

// the CV stage travels from psb_walker_begin to psb_walker_finish through global memory
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

// ---- per-term arithmetic ------------------------------------------------------------------------
// pass A contribution of one atom of a group of CV `cv` (position r, unwrapped, fp64; momentum p
// only when the ABF fast path sums momenta per group)
// non-point kinds are kept out of line: the general kernels then keep the compact point-CV path
static __device__ __noinline__ void accumulate_nonpoint(const CvDev& cv, double inv_n, uint32_t t, V3 r, double* acc) {
  if (cv.kind == PSB_CV_ROG) {
    acc[0] += r.x * r.x + r.y * r.y + r.z * r.z;
  } else if (cv.kind == PSB_CV_GYRATION) {  // shape.py:118-135: sum_i r_i r_i^T (uncentred, like the reference)
    acc[0] += r.x * r.x; acc[1] += r.x * r.y; acc[2] += r.x * r.z;
    acc[3] += r.y * r.y; acc[4] += r.y * r.z; acc[5] += r.z * r.z;
  } else if (cv.kind == PSB_CV_RING) {  // angles.py:186-199: the four Fourier sums over the ring
    double s, c, S, C;
    ring_coefficients(t - cv.t0, cv.t1 - cv.t0, &s, &c, &S, &C);
    acc[0] += s * r.x; acc[1] += s * r.y; acc[2] += s * r.z;
    acc[3] += c * r.x; acc[4] += c * r.y; acc[5] += c * r.z;
    acc[6] += C * r.x; acc[7] += C * r.y; acc[8] += C * r.z;
    acc[9] += S * r.x; acc[10] += S * r.y; acc[11] += S * r.z;
  } else if (cv.kind == PSB_CV_RMSD) {  // weighted centroid, plain sum, raw covariance sum_i r_i (x) Q_i
    const double w = cv.w ? __ldg(cv.w + (t - cv.t0)) : inv_n;
    const double* q = cv.ref + 3 * (size_t)(t - cv.t0);
    const double q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
    acc[0] += w * r.x; acc[1] += w * r.y; acc[2] += w * r.z;
    acc[3] += r.x; acc[4] += r.y; acc[5] += r.z;
    acc[6] += r.x * q0; acc[7] += r.x * q1; acc[8] += r.x * q2;
    acc[9] += r.y * q0; acc[10] += r.y * q1; acc[11] += r.y * q2;
    acc[12] += r.z * q0; acc[13] += r.z * q1; acc[14] += r.z * q2;
    acc[15] += r.x * r.x + r.y * r.y + r.z * r.z;  // for the residual without a second pass (finalize_cv_A)
  }
}
template <bool GENERAL>
__device__ __forceinline__ void accumulate_term(const CvDev& cv, bool with_p, double inv_n, uint32_t t,
                                                V3 r, V3 p, double* acc) {
  if (!GENERAL || is_point_kind(cv.kind)) {
    acc[0] += r.x; acc[1] += r.y; acc[2] += r.z;
    if (with_p) {
      acc[3] += p.x; acc[4] += p.y; acc[5] += p.z;
    }
  } else {
    accumulate_nonpoint(cv, inv_n, t, r, acc);
  }
}

// pass B contribution: |P_i U - Q_i|^2 (orientation.py:76-95)
__device__ __forceinline__ double rmsd_residual(const CvDev& cv, const double* x, uint32_t t, V3 r) {
  const double P[3] = {r.x - x[9], r.y - x[10], r.z - x[11]};
  const double* q = cv.ref + 3 * (size_t)(t - cv.t0);
  double s = 0.0;
#pragma unroll
  for (int bb = 0; bb < 3; ++bb) {
    const double e = P[0] * x[0 + bb] + P[1] * x[3 + bb] + P[2] * x[6 + bb] - __ldg(q + bb);
    s += e * e;
  }
  return s;
}

// gradient of a non-point CV with respect to the atom of term t (unwrapped position r)
static __device__ __noinline__ V3 nonpoint_gradient(const Fin& f, const CvDev& cv, int c, double inv_n,
                                                    uint32_t t, V3 r) {
  if (cv.kind == PSB_CV_ROG) return f.cvx[c][0] * r;
  if (cv.kind == PSB_CV_GYRATION) {
    const double* x = f.cvx[c];
    return v3(x[0] * r.x + x[1] * r.y + x[2] * r.z, x[3] * r.x + x[4] * r.y + x[5] * r.z,
              x[6] * r.x + x[7] * r.y + x[8] * r.z);
  }
  if (cv.kind == PSB_CV_RING) {
    const double* E = f.cvx[c];
    double s, cc, S, C;
    ring_coefficients(t - cv.t0, cv.t1 - cv.t0, &s, &cc, &S, &C);
    return v3(C * E[0] + S * E[3] + s * E[6] + cc * E[9], C * E[1] + S * E[4] + s * E[7] + cc * E[10],
              C * E[2] + S * E[5] + s * E[8] + cc * E[11]);
  }
  if (cv.kind == PSB_CV_RMSD) {
    const double* x = f.cvx[c];
    const double P[3] = {r.x - x[9], r.y - x[10], r.z - x[11]};
    const double* q = cv.ref + 3 * (size_t)(t - cv.t0);
    const double q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
    const double w = cv.w ? __ldg(cv.w + (t - cv.t0)) : inv_n;
    double D[3];
    for (int a = 0; a < 3; ++a)
      D[a] = P[a] - (q0 * x[a * 3 + 0] + q1 * x[a * 3 + 1] + q2 * x[a * 3 + 2]) - w * x[12 + a];
    return v3(x[15] * D[0], x[15] * D[1], x[15] * D[2]);
  }
  // coordination: the pair kernel left d xi / d r in the gradient workspace
  const double* gr = cv.grad + 3 * (size_t)(t - cv.t0);
  return v3(__ldcg(gr), __ldcg(gr + 1), __ldcg(gr + 2));
}

// per-term gradient of the term's CV components w.r.t. the atom; returns ncomp
// (r: unwrapped position of the atom, needed by the non-point kinds only)
__device__ __forceinline__ int term_gradient_at(const Model& M, const Fin& f, uint32_t t, int g, V3 r, V3* gv) {
  const GroupDev& G = M.grp[g];
  const CvDev& cv = M.cv[G.cv];
  if (is_point_kind(cv.kind)) {
    for (int j = 0; j < cv.ncomp; ++j)
      gv[j] = v3(f.gpt[g][j][0] * G.inv_n, f.gpt[g][j][1] * G.inv_n, f.gpt[g][j][2] * G.inv_n);
    return cv.ncomp;
  }
  gv[0] = nonpoint_gradient(f, cv, G.cv, G.inv_n, t, r);
  return 1;
}
template <class A>
__device__ __forceinline__ int term_gradient(const Model& M, const Snap& S, const Fin& f, uint32_t t,
                                             int g, uint32_t slot, V3* gv) {
  V3 r = v3(0, 0, 0);
  const int kind = M.cv[M.grp[g].cv].kind;
  if (!is_point_kind(kind) && kind != PSB_CV_COORDINATION && kind != PSB_CV_ERMSD) r = A::position(S, slot);
  return term_gradient_at(M, f, t, g, r, gv);
}
// Unique atom u of the selection: its tag, slot and CSR range of terms.  The slot comes straight from
// the tag (not from term_slot), so the chain is {tag, range} -> {slot, terms} -> {force, position}.
template <class A>
__device__ __forceinline__ void unique_atom(const Model& M, const Snap& S, bool all_unique, uint32_t u,
                                            uint32_t& e0, uint32_t& e1, uint32_t& slot) {
  uint32_t tag;
  if (all_unique) {
    e0 = u; e1 = u + 1;
    tag = __ldg(M.term_tag + u);
  } else {
    e0 = __ldg(M.uniq_ptr + u); e1 = __ldg(M.uniq_ptr + u + 1);
    tag = __ldg(M.uniq_tag + u);
  }
  slot = A::slot(S, M, tag);
}

// particle tag of term t: groups given as ascending ranges need no index load at all
__device__ __forceinline__ uint32_t term_tag_of(const Model& M, const GroupDev& G, uint32_t t) {
  return G.contig ? G.tag0 + (t - G.t0) : __ldg(M.term_tag + t);
}

template <class A>
__device__ __forceinline__ void add_force(const Snap& S, uint32_t slot, V3 fo) {
  A::store_force(S, slot, A::load_force(S, slot), fo);
}

// raw velocity row type of the layouts that split the momentum into a load and the arithmetic (int otherwise)
template <typename A, bool RAW> struct RawVelocity { using type = int; };
template <typename A> struct RawVelocity<A, true> { using type = typename A::VR; };

// ---- the kernel -------------------------------------------------------------------------------
// The body of the step: `nb` CTAs (this one is CTA `b` of them) step the simulation described by (Mglobal, S).
// psb_step launches exactly those CTAs (step_kernel); psb_batch_step launches the CTAs of many independent
// simulations at once and hands every CTA its own (model, snapshot, nb, b) (step_kernel_batch).
template <int LAYOUT, typename Real, int METHOD, int CLASS>
__device__ __forceinline__ void step_body(const Model* __restrict__ Mglobal, const Snap& S, const int nb_arg, const int b_arg) {
  constexpr bool GENERAL = (CLASS == SC_GENERAL);
  constexpr bool STREAM = (CLASS == SC_STREAM);  // slot-ordered, streams staged by TMA (HOOMD layout only)
  constexpr bool TAGS = (CLASS == SC_TAGS);      // slot-ordered, slot -> tag array from the engine (HOOMD and LAMMPS layouts)
  constexpr bool SLOT = (CLASS == SC_SLOT) || STREAM || TAGS;
  static_assert(!STREAM || LAYOUT == PSB_LAYOUT_HOOMD, "the TMA-staged streams are written for the HOOMD layout");
  using A = Access<LAYOUT, Real>;
  using FV = typename A::FV;
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  __shared__ StepShared sh;
  __shared__ __align__(16) Model sM;
  if constexpr (METHOD == SM_ABF) {
    __shared__ NetShared nsh;
    if (threadIdx.x == 0) sh.net = &nsh;  // read after the first block barrier only
  }

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (S.dbg && tid == 0) S.dbg[(size_t)b_arg * 16 + 0] = global_ns();  // PSB_STAMP(0) before sh.vb exists
  // The model (index tables, method constants, state pointers: ~3 KB, fixed per handle) lives in
  // device memory and is copied into shared memory by the whole CTA: one L2 round trip.  As a
  // kernel parameter it cost 2 us of launch latency per step (measured: tools/launch_floor.py), and
  // the finalizers, which walk it from a single warp, would take a serialised constant-cache miss
  // per table line.
  {
    static_assert(sizeof(Model) % 8 == 0, "Model is copied in 8-byte words");
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(Mglobal);
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(&sM);
    // Programmatic dependent launch (when the host enables it): the next step's launch, this copy of
    // the immutable model and -- see pass A -- its gather of engine arrays this library never writes
    // overlap the previous step's tail; everything after griddepcontrol.wait sees the previous kernel
    // complete.  Without the launch attribute both instructions are no-ops.
    // LL exchange (Snap::cta_seq set): this CTA's launch count must be bumped before the next launch of
    // the handle may start, so the trigger follows the atomic (one round trip, shared with the copy).
    unsigned long long seq = 0;
    if (S.cta_seq == nullptr) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    else if (tid == 0) seq = atomicAdd(S.cta_seq + b_arg, 1ULL);
    for (int i = tid; i < (int)(sizeof(Model) / 8); i += BLOCK) dst[i] = __ldg(src + i);
    if (tid == 0) {
      sh.seq = seq;
      sh.vnb = nb_arg;
      sh.vb = b_arg;
    }
  }
  __syncthreads();
  if (S.cta_seq != nullptr) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const Model& M = sM;
  PSB_CLKA(8);

  const int nb = nb_arg, b = b_arg;
  const bool cta0 = (b == 0);
  const int k = M.k;
  // ---- SC_STREAM: every WARP owns a contiguous run of slots and a private ring of stages (psb_stream.cuh) ------
  using R4s = typename Real4<Real>::type;
  constexpr uint32_t PB = (uint32_t)sizeof(R4s);  // bytes per position / velocity / force row
  const StreamGeom& SG = M.sg;
  unsigned char* const ring = dyn_smem + (size_t)warp * SG.stages * SG.stage_bytes;  // this warp's stages
  uint64_t* const full = sh.full + warp * STREAM_MAX_STAGES;
  uint32_t st_lo = 0, st_hi = 0;              // this warp's slot range
  int st_n = 0, st_total = 0, st_issued = 0;  // chunks, items (gather [+ scatter] chunks), items issued so far
  unsigned long long grp_nib = 0;             // group + 1 of this lane's first 16 rows, one nibble each (gather -> scatter)
  constexpr int P0U = STREAM ? 16 : 1;        // pass 0: tag -> slot lookups per thread requested in the prologue
  uint32_t p0_slot[P0U];
  // item i: gather chunk i (i < st_n) or scatter chunk i - st_n, placed in stage i % S (tracked incrementally:
  // no divisions on this path); lane 0 arms the stage's barrier and starts the copies
  int st_pstage = 0;  // stage of the next item to be issued
  int st_cstage = 0;  // stage of the next item to be consumed ...
  uint32_t st_cphase = 0;  // ... and the parity its barrier completes with
  auto stream_issue_upto = [&](int limit) {
    if constexpr (STREAM) {
      while (st_issued < limit && st_issued < st_total) {
        const int i = st_issued++;
        const int ps = st_pstage;
        st_pstage = (ps + 1 == SG.stages) ? 0 : ps + 1;
        if (lane != 0) continue;
        unsigned char* stage = ring + (uint32_t)ps * SG.stage_bytes;
        uint64_t* bar = &full[ps];
        if (i < st_n) {
          const uint32_t base = st_lo + (uint32_t)i * (uint32_t)SG.ch;
          const uint32_t cnt = min((uint32_t)SG.ch, st_hi - base), cnt4 = cnt & ~3u;
          const uint32_t bi = (SG.g_img && cnt4) ? cnt4 * 12u : 0u, bv = SG.g_vel ? cnt * PB : 0u;
          mbar_expect_tx(bar, cnt * PB + bi + bv);
          tma_bulk_g2s(stage, reinterpret_cast<const unsigned char*>(S.positions) + (size_t)base * PB, cnt * PB, bar);
          if (bi) tma_bulk_g2s(stage + SG.g_img, reinterpret_cast<const unsigned char*>(S.images) + (size_t)base * 12u, bi, bar);
          if (bv) tma_bulk_g2s(stage + SG.g_vel, reinterpret_cast<const unsigned char*>(S.vel_mass) + (size_t)base * PB, bv, bar);
        } else {
          const uint32_t base = st_lo + (uint32_t)(i - st_n) * (uint32_t)SG.ch;
          const uint32_t cnt = min((uint32_t)SG.ch, st_hi - base), cnt4 = cnt & ~3u;
          const uint32_t bp = SG.s_pos ? cnt * PB : 0u, bi = (SG.s_img && cnt4) ? cnt4 * 12u : 0u;
          mbar_expect_tx(bar, cnt * PB + bp + bi);
          tma_bulk_g2s(stage, reinterpret_cast<const unsigned char*>(S.forces) + (size_t)base * PB, cnt * PB, bar);
          if (bp) tma_bulk_g2s(stage + SG.s_pos, reinterpret_cast<const unsigned char*>(S.positions) + (size_t)base * PB, bp, bar);
          if (bi) tma_bulk_g2s(stage + SG.s_img, reinterpret_cast<const unsigned char*>(S.images) + (size_t)base * 12u, bi, bar);
        }
      }
    }
  };
  // the consumer side: wait for the next item, return its stage, advance
  auto stream_next_stage = [&]() -> unsigned char* {
    unsigned char* stage = ring + (uint32_t)st_cstage * SG.stage_bytes;
    mbar_wait(&full[st_cstage], st_cphase);
    if (++st_cstage == SG.stages) { st_cstage = 0; st_cphase ^= 1u; }
    return stage;
  };
  if constexpr (STREAM) {
    const uint32_t N = (uint32_t)M.natoms;
    const uint32_t wrun = (((N + nb - 1) / nb + NWARPS - 1) / NWARPS + 3u) & ~3u;  // warp runs start on multiples of 4 slots
    const uint32_t schunk = wrun * NWARPS;                                         // (16-byte aligned image rows)
    const uint32_t c_lo = min(N, (uint32_t)b * schunk), c_hi = min(N, c_lo + schunk);
    st_lo = min(c_hi, c_lo + (uint32_t)warp * wrun);
    st_hi = min(c_hi, st_lo + wrun);
    st_n = (int)((st_hi - st_lo + (uint32_t)SG.ch - 1u) / (uint32_t)SG.ch);
    const bool scatters = (METHOD != SM_UNBIASED) && S.mode == MODE_STEP;
    st_total = scatters ? 2 * st_n : st_n;
    if (lane == 0) {
      for (int s2 = 0; s2 < SG.stages; ++s2) mbar_init(&full[s2], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    // Engine arrays this library never writes may be fetched before the dependency wait.  First what the critical
    // path needs first -- the tag -> slot lookups of pass 0, one register batch of P0U per thread -- then the
    // gather chunks, which then stream in behind them while pass 0 scatters its map words.
    {
      const uint32_t chunk0 = (M.T + nb - 1) / nb;
      const uint32_t q0 = min(M.T, (uint32_t)b * chunk0), q1 = min(M.T, q0 + chunk0);
#pragma unroll
      for (int u = 0; u < P0U; ++u) {
        const uint32_t t = min(q0 + tid + (uint32_t)u * BLOCK, max(q1, 1u) - 1u);  // clamped: loads stay unconditional
        int g = 0;
        while (g + 1 < M.ngroups && t >= M.grp[g].t1) ++g;
        p0_slot[u] = A::slot(S, M, term_tag_of(M, M.grp[g], t));
      }
    }
    stream_issue_upto(min(st_n, min(SG.stages, SG.early)));
  }
  const MethodDev& m = M.m;
  constexpr bool metad_direct = (METHOD == SM_METAD_DIRECT);
  constexpr bool grid_metad = (METHOD == SM_METAD_GRID);
  constexpr bool is_abf = (METHOD == SM_ABF);
  constexpr bool WALK = metad_direct || grid_metad;  // walker modes exist for metadynamics only
  const bool any_rmsd = GENERAL && M.any_rmsd;
  const bool any_coord = GENERAL && M.any_coord;
  const bool all_unique = !GENERAL || M.all_unique;
  const bool abf_fast = is_abf && (!GENERAL || M.abf_fast);
  const bool need_term_slot = (GENERAL || WALK) && M.need_term_slot;
  const bool skip_cv = WALK && !SLOT && (S.mode == MODE_WALKER_FINISH);  // CVs were evaluated by walker_begin
  const bool momenta_sums = S.need_momenta && abf_fast;
  constexpr bool slot_major = SLOT;  // the host picks this instantiation only for multi-CTA, non-walker handles
  // slot-ordered class with a slot -> atom array (OpenMM: ids itself, openmm.py:100-107; HOOMD: the optional
  // psb_snapshot.tags): the group of a slot comes from a tag -> group byte table and neither a tag -> slot inverse
  // nor pass 0 is needed
  constexpr bool TAGMAP = SLOT && (LAYOUT == PSB_LAYOUT_OPENMM_CUDA || TAGS);
  const bool will_scatter = (METHOD != SM_UNBIASED) &&
                            (S.mode == MODE_STEP || (WALK && S.mode == MODE_WALKER_FINISH));

  // ---- this CTA's terms ----------------------------------------------------------------------
  // gridDim.x > 1: exactly one group per CTA, an even share of its terms; gridDim.x == 1: all terms.
  // Up to CT terms per thread are "cached": slot, group, position and the prefetched force row stay
  // in registers until the scatter.
  int seg_g = -1;
  uint32_t lo = 0, hi = 0;
  if (!skip_cv) {
    if (slot_major) {
      // no per-CTA group: handled by the slot-ordered passes below
    } else if (nb > 1) {
      for (int g = 0; g < M.ngroups; ++g)
        if (b >= M.cta_lo[g] && b <= M.cta_hi[g]) seg_g = g;
      if (seg_g >= 0) {
        const GroupDev& G = M.grp[seg_g];
        const uint32_t nc = (uint32_t)(M.cta_hi[seg_g] - M.cta_lo[seg_g] + 1);
        const uint32_t per = (G.t1 - G.t0 + nc - 1) / nc;
        lo = min(G.t1, G.t0 + (uint32_t)(b - M.cta_lo[seg_g]) * per);
        hi = min(G.t1, lo + per);
      }
    } else {
      hi = M.T;
    }
  }
  const bool cached = !skip_cv && !slot_major && (nb == 1 || seg_g >= 0) && (hi - lo <= (uint32_t)(CT * BLOCK));
  // pass-A partials through the LL buffer (host: every CTA with a group is `cached`, nothing but the
  // partials crosses CTAs before the finalizer); grid-uniform
  const bool ll = !GENERAL && !STREAM && !WALK && M.ll_reduce && nb > 1;  // slot-ordered classes: two-level (reduce_rows_ll2)
  const uint32_t ll_stamp = (uint32_t)(sh.seq + 1ULL);
  const int g_begin = (skip_cv || slot_major) ? 0 : (nb == 1 ? 0 : max(seg_g, 0));
  const int g_end = (skip_cv || slot_major) ? 0 : (nb == 1 ? M.ngroups : seg_g + 1);  // seg_g < 0: empty
  uint32_t stamp = 0, s_lo = 0, s_hi = 0;  // slot-major mode: launch stamp and this CTA's slot range
  // slot-ordered classes: group + 1 of this thread's rows s_lo + tid + r * BLOCK is kept from the gather for the scatter
  // (sh.rowg; valid while a CTA's slice has at most ROWG_ROWS rows per thread: no second look at the slot words then)
  bool nib_ok = false;
  // slot-ordered class: groups whose CV is RadiusOfGyration (bit g); all others are point groups
  uint32_t rog_bits = 0;
  if (SLOT)
    for (int g = 0; g < M.ngroups && g < 4; ++g)
      if (M.cv[M.grp[g].cv].kind == PSB_CV_ROG) rog_bits |= 1u << g;
  auto slot_word = [&](uint32_t sl) -> uint32_t {  // stamp | group + 1 when slot `sl` holds a selected atom
    if constexpr (TAGMAP) {
      // HOOMD: ParticleData's tag array; LAMMPS: atom->tag, whose ids are 1-based (atom `tag` of a CV is LAMMPS id tag + 1)
      const uint32_t tag = (uint32_t)(__ldg(reinterpret_cast<const int*>(TAGS ? S.tags : S.ids) + sl) -
                                      ((TAGS && LAYOUT == PSB_LAYOUT_LAMMPS) ? 1 : 0));
      const uint32_t g = tag < (uint32_t)M.natoms ? (uint32_t)__ldg(M.tag_group + tag) : 0u;
      return g ? (stamp | g) : 0u;
    } else {
      return __ldcg(M.slot_map + sl);
    }
  };
  uint32_t cslot[CT];
  int cg[CT];
  V3 cr[CT], cp[CT];
  FV cfv[CT];

  // ---- pass A, early part: engine arrays this library never writes (tag -> slot map, positions,
  // images, velocities) may be read while the previous step kernel is still finishing (PDL) ---------
  PSB_CLKA(9);
  if (cached) {
#pragma unroll
    for (int i = 0; i < CT; ++i) {
      const uint32_t t = lo + tid + i * BLOCK;
      cslot[i] = 0;
      cg[i] = -1;
      if (t < hi) {
        cg[i] = (nb == 1) ? (int)__ldg(M.term_group + t) : seg_g;
        cslot[i] = A::slot(S, M, term_tag_of(M, M.grp[cg[i]], t));
      }
    }
    if constexpr (LAYOUT == PSB_LAYOUT_OPENMM_CUDA || LAYOUT == PSB_LAYOUT_PLAIN3) {
      // raw velocity rows first, v / invM after every load has been issued (see Access<OPENMM_CUDA>::momentum_of)
      typename A::VR cv[CT];
#pragma unroll
      for (int i = 0; i < CT; ++i) {
        cp[i] = v3(0, 0, 0);
        if (cg[i] >= 0) {
          cr[i] = A::position(S, cslot[i]);
          if (momenta_sums) cv[i] = A::load_velocity(S, cslot[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < CT; ++i)
        if (cg[i] >= 0 && momenta_sums) cp[i] = A::momentum_of(cv[i]);
    } else {
#pragma unroll
      for (int i = 0; i < CT; ++i) {
        cp[i] = v3(0, 0, 0);
        if (cg[i] >= 0) {
          cr[i] = A::position(S, cslot[i]);
          if (momenta_sums) cp[i] = A::momentum(S, cslot[i]);
        }
      }
    }
  }
  int stage_na = 0;  // multi-CTA cached mode: accumulator rows of this CTA's group waiting in sh.stage
  if (cached) {
    if (nb == 1) {  // bare indices (groups of one atom): the owner writes the "sums", no reduction
#pragma unroll
      for (int i = 0; i < CT; ++i)
        if (cg[i] >= 0 && M.grp[cg[i]].t1 - M.grp[cg[i]].t0 == 1 &&
            (!GENERAL || is_point_kind(M.cv[M.grp[cg[i]].cv].kind))) {
          double* tot = sh.tot[cg[i]];
          tot[0] = cr[i].x; tot[1] = cr[i].y; tot[2] = cr[i].z;
          tot[3] = cp[i].x; tot[4] = cp[i].y; tot[5] = cp[i].z;
        }
    }
    for (int g = g_begin; g < g_end; ++g) {
      const GroupDev& G = M.grp[g];
      const CvDev& cv = M.cv[G.cv];
      if (nb == 1 && G.t1 - G.t0 == 1 && (!GENERAL || is_point_kind(cv.kind))) continue;  // done above
      const int na = nacc_for_kind(cv.kind, 0, momenta_sums);
      if (na == 0) continue;
      const bool with_p = (na == 6);
      double acc[NACC];
#pragma unroll
      for (int a = 0; a < NACC; ++a) acc[a] = 0.0;
#pragma unroll
      for (int i = 0; i < CT; ++i)
        if (cg[i] == g) accumulate_term<GENERAL>(cv, with_p, G.inv_n, lo + tid + i * BLOCK, cr[i], cp[i], acc);
      PSB_CLKA(10);
      if (nb > 1) {  // one group per CTA: the sums wait in shared memory until the previous kernel is done
        block_reduce_store(sh, acc, na, sh.stage, 1, false);
        stage_na = na;
      } else {
        block_reduce_store(sh, acc, na, &sh.tot[g][0], 1, false);
      }
      PSB_CLKA(11);
    }
  }
  // ---- slot-ordered gather (most atoms selected, too many to keep in registers): see pass A below.  The loop
  // only reads engine arrays (and, on the OpenMM layout, the immutable tag -> group table), so that layout runs
  // it BEFORE the dependency wait, like the register-cached gather of the other classes.
  auto slot_gather = [&](double (&acc4)[4][6]) {
    const uint32_t N = (uint32_t)M.natoms;
    uint32_t schunk = (N + nb - 1) / nb;
    schunk = (schunk + 31u) & ~31u;  // warp-aligned slices
    s_lo = min(N, (uint32_t)b * schunk);
    s_hi = min(N, s_lo + schunk);
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
      for (int a = 0; a < 6; ++a) acc4[g][a] = 0.0;
    // SU independent rows per thread in flight: every load is issued before the first use (rows of
    // unselected slots are loaded too -- they exist, and almost every slot is selected here)
#ifndef PSB_SLOT_SUA
#define PSB_SLOT_SUA 8
#endif
    constexpr int SU = is_abf ? 4 : PSB_SLOT_SUA;  // momenta double the registers per row in flight
    // (ABF only: its gather keeps 4 rows per thread in flight and has registers to spare; the 8-row gathers of the
    // other methods sit at the 128-register limit of two CTAs per SM and would spill)
    nib_ok = is_abf && (s_hi - s_lo) <= (uint32_t)ROWG_ROWS * BLOCK;
    constexpr bool RAWV = (LAYOUT == PSB_LAYOUT_OPENMM_CUDA || LAYOUT == PSB_LAYOUT_PLAIN3);  // momenta from raw velocity rows, after the loads
    for (uint32_t base = s_lo + tid; base < s_hi; base += SU * BLOCK) {
      uint32_t e[SU];
      V3 r[SU], p[SU];
      typename RawVelocity<A, RAWV>::type vraw[RAWV ? SU : 1];
#pragma unroll
      for (int u = 0; u < SU; ++u) {
        const uint32_t sl = min(base + u * BLOCK, s_hi - 1);  // clamped: no branch around the loads
        e[u] = slot_word(sl);
        r[u] = A::position(S, sl);
        if constexpr (RAWV) {
          p[u] = v3(0, 0, 0);
          if (momenta_sums) vraw[u] = A::load_velocity(S, sl);
        } else {
          p[u] = momenta_sums ? A::momentum(S, sl) : v3(0, 0, 0);
        }
      }
#pragma unroll
      for (int u = 0; u < SU; ++u) {
        if constexpr (RAWV)
          if (momenta_sums) p[u] = A::momentum_of(vraw[u]);
        const bool hit = (base + u * BLOCK < s_hi) && ((e[u] & ~15u) == stamp);
        const int g = hit ? (int)(e[u] & 15u) - 1 : -1;
        if constexpr (is_abf)
          if (nib_ok) sh.rowg[base - s_lo + u * BLOCK] = (uint8_t)(g + 1);  // [row * BLOCK + tid]
#pragma unroll
        const double rr = r[u].x * r[u].x + r[u].y * r[u].y + r[u].z * r[u].z;
#pragma unroll
        for (int gg = 0; gg < 4; ++gg) {  // branch-free select: lanes of a warp hold different groups
          const double w = (g == gg) ? 1.0 : 0.0;
          acc4[gg][0] = fma(w, (rog_bits >> gg) & 1u ? rr : r[u].x, acc4[gg][0]);  // RoG groups: sum of r.r
          acc4[gg][1] = fma(w, r[u].y, acc4[gg][1]);
          acc4[gg][2] = fma(w, r[u].z, acc4[gg][2]);
          if (momenta_sums) {
            acc4[gg][3] = fma(w, p[u].x, acc4[gg][3]);
            acc4[gg][4] = fma(w, p[u].y, acc4[gg][4]);
            acc4[gg][5] = fma(w, p[u].z, acc4[gg][5]);
          }
        }
      }
    }
  };
  if (slot_major && TAGMAP && !skip_cv) {
    stamp = 16u;  // ids is slot -> atom: group of a slot = tag_group[ids[slot]], no pass 0, no stamp to expire
    double acc4[4][6];
    slot_gather(acc4);
    const int na = momenta_sums ? 6 : 3;
#pragma unroll
    for (int g = 0; g < 4; ++g)
      if (g < M.ngroups) block_reduce_store(sh, acc4[g], na, &sh.stage4[g][0], 1, false);
  }
  // everything below may depend on the previous kernel of the stream: forces, method state, workspaces
  asm volatile("griddepcontrol.wait;" ::: "memory");

  // ---- force network (psb_set_force_net): header + parameters into shared memory -----------------
  if (is_abf && (M.nn != nullptr || (M.fs != nullptr && M.fs_staged)) && tid >= BLOCK - 64) {
    static_assert(sizeof(ForceNet) % 8 == 0 && offsetof(NetShared, nnw) == offsetof(NetShared, nnh) + sizeof(ForceNet),
                  "header and staged parameters are copied as one block");
    const bool net = M.nn != nullptr;  // a handle has a network or a series, never both
    const unsigned long long* src = net ? reinterpret_cast<const unsigned long long*>(M.nn)
                                        : reinterpret_cast<const unsigned long long*>(M.fs);
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(&sh.net->nnh);
    const int words = net ? (int)(sizeof(ForceNet) / 8) + M.nn_np : M.fs_staged;
    for (int i = tid - (BLOCK - 64); i < words; i += 64) dst[i] = __ldcg(src + i);
  }

  // ---- prologue: method-state scalars as they are when the step begins --------------------------
  // The loads are issued here but land in shared memory only after this thread has issued its
  // gather loads (commit_pre below): storing them right away would stall warps 0-2 for one L2 round
  // trip before they even start gathering.
  long long pre_ncalls = 0, pre_idx = 0;
  unsigned long long pre_epoch = 0;
  double pre_abf = 0.0;
  if (tid == 0) {
    pre_ncalls = __ldcg(M.st.ncalls);
    pre_idx = M.st.idx ? __ldcg(M.st.idx) : 0;
  }
  if (tid >= 64 && tid < 64 + NBARS) pre_epoch = __ldcg(M.epochs + (tid - 64));
  if (is_abf && tid >= 32 && tid < 32 + 3 * MAXK) {
    const int j = tid - 32, c = j & (MAXK - 1), which = j / MAXK;
    const double* src = which == 0 ? M.st.Wp : (which == 1 ? M.st.Wp_ : M.st.force);
    pre_abf = __ldcg(src + c);
  }
  auto commit_pre = [&]() {
    if (tid == 0) {
      sh.pre.ncalls = pre_ncalls;
      sh.pre.idx = pre_idx;
      sh.bars_used = 0;
    }
    if (tid >= 64 && tid < 64 + NBARS) sh.epoch[tid - 64] = pre_epoch;
    if (is_abf && tid >= 32 && tid < 32 + 3 * MAXK) {
      const int j = tid - 32, c = j & (MAXK - 1), which = j / MAXK;
      double* dst = which == 0 ? sh.pre.Wp : (which == 1 ? sh.pre.Wp_ : sh.pre.force);
      dst[c] = pre_abf;
    }
  };

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
    if (cached) {
#pragma unroll
      for (int i = 0; i < CT; ++i) {
        if (cg[i] >= 0) {
          if (need_term_slot) M.term_slot[lo + tid + i * BLOCK] = cslot[i];
          if (will_scatter) cfv[i] = A::load_force(S, cslot[i]);
        }
      }
      if (tid < stage_na) {
        if (ll) ll_store(M.ll + (size_t)(seg_g * NACC + tid) * nb + b, sh.stage[tid], ll_stamp);
        else __stcg(M.partials + (size_t)(seg_g * NACC + tid) * nb + b, sh.stage[tid]);
      }
    }
    const bool pre_early = !ll || (SLOT && !TAGMAP);  // pass 0's grid barrier needs the epochs
    if (pre_early) commit_pre();  // LL exchange: the prologue values are not needed before the reduction
    // ---- slot-major gather (most atoms selected, too many to keep in registers) ---------------
    // A random tag -> slot permutation makes the tag-ordered gather pay one L1 tag lookup per
    // 16-byte row (measured: ~6 lookups per atom, 1 lookup/clk/SM -> 20 us per 1e6 atoms, far
    // below HBM speed).  Instead: pass 0 scatters one word per selected atom into slot_map
    // (stamp | group), and the gather and the force update then stream positions, images,
    // velocities and forces in SLOT order, fully coalesced.  At most 4 groups (fast-class only).
    if (slot_major && !TAGMAP) {
      __syncthreads();  // sh.epoch
      // 28-bit launch stamp in 1 .. 2^28 - 1.  Only the slots of currently selected atoms are rewritten each
      // step, so an entry left behind by an atom that has since moved to another slot would look current again
      // once the stamps repeat: the launch that restarts the sequence (the first, then every 2^28 - 1 launches)
      // clears the whole map first -- one extra grid barrier on that launch only.  Model::stamp_period shortens
      // the cycle (tests).
      const unsigned long long period = M.stamp_period ? (unsigned long long)M.stamp_period : 0x0FFFFFFFULL;
      const uint32_t stamp_val = (uint32_t)(sh.epoch[BAR_SLOT] % period) + 1u;
      stamp = stamp_val << 4;
      if (stamp_val == 1u) {
        for (size_t i = (size_t)b * BLOCK + tid; i < (size_t)M.natoms; i += (size_t)nb * BLOCK) M.slot_map[i] = 0u;
        grid_sync(M, sh, BAR_C);  // not used otherwise by this class (its ABF variant has no pass C)
      }
      const uint32_t chunk0 = (M.T + nb - 1) / nb;
      const uint32_t q0 = min(M.T, (uint32_t)b * chunk0), q1 = min(M.T, q0 + chunk0);
      // group of a term and its tag without branches (<= 4 groups): bounds and range starts in registers
      uint32_t gb[3], gt0[4], gtag0[4];
      bool all_contig = true;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool ok = j < M.ngroups;
        if (j < 3) gb[j] = (j < M.ngroups - 1) ? M.grp[j].t1 : 0xFFFFFFFFu;
        gt0[j] = ok ? M.grp[j].t0 : 0u;
        gtag0[j] = ok ? M.grp[j].tag0 : 0u;
        all_contig = all_contig && (!ok || M.grp[j].contig);
      }
      uint32_t p0_done = q0;  // terms below it are handled by the prologue batch
      if constexpr (STREAM) {
        p0_done = min(q1, q0 + (uint32_t)P0U * BLOCK);
#pragma unroll
        for (int u = 0; u < P0U; ++u) {
          const uint32_t t = q0 + tid + (uint32_t)u * BLOCK;
          if (t < p0_done) {
            const int g = (int)(t >= gb[0]) + (int)(t >= gb[1]) + (int)(t >= gb[2]);
            M.slot_map[p0_slot[u]] = stamp | (uint32_t)(g + 1);
          }
        }
      }
      for (uint32_t base = p0_done + tid; base < q1; base += 8 * BLOCK) {
        uint32_t slot[8], word[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint32_t t = min(base + u * BLOCK, q1 - 1);  // clamped: loads stay unconditional
          const int g = (int)(t >= gb[0]) + (int)(t >= gb[1]) + (int)(t >= gb[2]);
          const uint32_t t0g = g == 0 ? gt0[0] : (g == 1 ? gt0[1] : (g == 2 ? gt0[2] : gt0[3]));
          const uint32_t tg0 = g == 0 ? gtag0[0] : (g == 1 ? gtag0[1] : (g == 2 ? gtag0[2] : gtag0[3]));
          const uint32_t tag = all_contig ? tg0 + (t - t0g) : __ldg(M.term_tag + t);
          slot[u] = A::slot(S, M, tag);
          word[u] = stamp | (uint32_t)(g + 1);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (base + u * BLOCK < q1) M.slot_map[slot[u]] = word[u];
      }
      PSB_STAMP(7);
      grid_sync(M, sh, BAR_SLOT);
    }
    if constexpr (STREAM) {
      stream_issue_upto(SG.stages);  // force rows may follow now (the dependency wait is behind us)
      double acc4[4][6];
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int a = 0; a < 6; ++a) acc4[g][a] = 0.0;
      const int S_ = SG.stages;
      const uint32_t CH = (uint32_t)SG.ch;  // 128 slots: four rows per lane
      // slot_map words (written by pass 0; their sectors were written piecewise, a read costs a DRAM-class latency):
      // fetched MG chunks ahead so that one latency covers MG chunks of work
      constexpr int MG = 4;
      uint32_t wnext[MG][STREAM_RPL];
      auto map_words = [&](int c0, uint32_t (&w)[MG][STREAM_RPL]) {
#pragma unroll
        for (int j = 0; j < MG; ++j) {
          const uint32_t base = st_lo + (uint32_t)(c0 + j) * CH;
#pragma unroll
          for (int u = 0; u < STREAM_RPL; ++u) {
            const uint32_t sl = base + lane + 32u * u;
            w[j][u] = (c0 + j < st_n && sl < st_hi) ? __ldcg(M.slot_map + sl) : 0u;
          }
        }
      };
      map_words(0, wnext);
      PSB_STAMPS(8);
      // One instantiation per (group count, RoG present): the loop is issue-bound (conversions run at 16 per clock
      // and SM, fp64 at 64), so it carries no per-row code it does not need.  Point groups sum the fp64-converted
      // stored positions and the image flags separately (integers, exact) and unwrap the SUM, sum (p_i + L n_i) =
      // sum p_i + L sum n_i (hoomd.py:179-181 per atom; equal within a few ulp, far inside the 1e-10 gate); RoG
      // groups need every unwrapped r_i for r_i . r_i and keep the per-atom form.
      auto gather_chunks = [&](auto NGc, auto ROGc) {
        constexpr int NG = decltype(NGc)::value;
        constexpr bool ROG = decltype(ROGc)::value;
        int isum[NG][3];
#pragma unroll
        for (int g = 0; g < NG; ++g) isum[g][0] = isum[g][1] = isum[g][2] = 0;
        for (int c0 = 0; c0 < st_n; c0 += MG) {
          uint32_t wcur[MG][STREAM_RPL];
#pragma unroll
          for (int j = 0; j < MG; ++j)
#pragma unroll
            for (int u = 0; u < STREAM_RPL; ++u) wcur[j][u] = wnext[j][u];
          map_words(c0 + MG, wnext);  // in flight while these MG chunks are processed
#pragma unroll
          for (int j = 0; j < MG; ++j) {
          const int c = c0 + j;
          if (c >= st_n) break;
          const uint32_t base = st_lo + (uint32_t)c * CH;
          const uint32_t cnt = min(CH, st_hi - base), cnt4 = cnt & ~3u;
          uint32_t w[STREAM_RPL];
#pragma unroll
          for (int u = 0; u < STREAM_RPL; ++u) w[u] = wcur[j][u];
          const unsigned char* stage = stream_next_stage();
          if (c < 3) PSB_STAMPS(9 + 2 * c);
          R4s p4[STREAM_RPL], v4[STREAM_RPL];
          int im[STREAM_RPL][3];
#pragma unroll
          for (int u = 0; u < STREAM_RPL; ++u) {  // all shared-memory loads of the rows first
            const uint32_t row = min(lane + 32u * u, cnt - 1u);  // clamped: loads stay unconditional
            p4[u] = reinterpret_cast<const R4s*>(stage)[row];
            if (SG.g_img) {
              if (row < cnt4) {
                const int* q = reinterpret_cast<const int*>(stage + SG.g_img) + 3 * row;
                im[u][0] = q[0]; im[u][1] = q[1]; im[u][2] = q[2];
              } else {  // tail rows of the very last run (bulk copies move multiples of 16 bytes)
                const int* q = reinterpret_cast<const int*>(S.images) + 3 * (size_t)(base + row);
                im[u][0] = __ldg(q); im[u][1] = __ldg(q + 1); im[u][2] = __ldg(q + 2);
              }
            } else {
              im[u][0] = im[u][1] = im[u][2] = 0;
            }
            if (momenta_sums) v4[u] = reinterpret_cast<const R4s*>(stage + SG.g_vel)[row];
          }
#pragma unroll
          for (int u = 0; u < STREAM_RPL; ++u) {
            V3 r = v3((double)p4[u].x, (double)p4[u].y, (double)p4[u].z);
            if (ROG && SG.g_img) {  // hoomd.py:179-181: pos + diag(H) * images, in fp64
              r.x = __dadd_rn(r.x, __dmul_rn(S.L[0], (double)im[u][0]));
              r.y = __dadd_rn(r.y, __dmul_rn(S.L[1], (double)im[u][1]));
              r.z = __dadd_rn(r.z, __dmul_rn(S.L[2], (double)im[u][2]));
            }
            V3 pm = v3(0, 0, 0);
            if (momenta_sums)  // hoomd.py:192-196: (M * V) in the array's own dtype, promoted afterwards
              pm = v3((double)(Real)(v4[u].w * v4[u].x), (double)(Real)(v4[u].w * v4[u].y), (double)(Real)(v4[u].w * v4[u].z));
            const bool hit = (lane + 32u * u < cnt) && ((w[u] & ~15u) == stamp);
            const int g = hit ? (int)(w[u] & 15u) - 1 : -1;
            if (c * STREAM_RPL + u < 16) grp_nib |= (unsigned long long)(g + 1) << (4 * (c * STREAM_RPL + u));  // for the scatter
#pragma unroll
            for (int gg = 0; gg < NG; ++gg) {
              const bool mine = (g == gg);
              const double wgt = mine ? 1.0 : 0.0;
              double a0 = r.x;
              if (ROG && ((rog_bits >> gg) & 1u)) a0 = r.x * r.x + r.y * r.y + r.z * r.z;  // RoG groups: sum of r.r
              acc4[gg][0] = fma(wgt, a0, acc4[gg][0]);
              acc4[gg][1] = fma(wgt, r.y, acc4[gg][1]);
              acc4[gg][2] = fma(wgt, r.z, acc4[gg][2]);
              if (!ROG) {
                const int mi = mine ? 1 : 0;
                isum[gg][0] += mi * im[u][0]; isum[gg][1] += mi * im[u][1]; isum[gg][2] += mi * im[u][2];
              }
              if (momenta_sums) {
                acc4[gg][3] = fma(wgt, pm.x, acc4[gg][3]);
                acc4[gg][4] = fma(wgt, pm.y, acc4[gg][4]);
                acc4[gg][5] = fma(wgt, pm.z, acc4[gg][5]);
              }
            }
          }
          __syncwarp();  // stage drained
          if (c < 3) PSB_STAMPS(10 + 2 * c);
          stream_issue_upto((SG.flags & 1) ? min(c + 1 + S_, st_n) : c + 1 + S_);
          }
        }
        if (!ROG && SG.g_img) {
#pragma unroll
          for (int gg = 0; gg < NG; ++gg) {
            acc4[gg][0] = fma(S.L[0], (double)isum[gg][0], acc4[gg][0]);
            acc4[gg][1] = fma(S.L[1], (double)isum[gg][1], acc4[gg][1]);
            acc4[gg][2] = fma(S.L[2], (double)isum[gg][2], acc4[gg][2]);
          }
        }
      };
      using std::integral_constant;
      if (rog_bits) gather_chunks(integral_constant<int, 4>{}, integral_constant<bool, true>{});
      else if (M.ngroups <= 2) gather_chunks(integral_constant<int, 2>{}, integral_constant<bool, false>{});
      else gather_chunks(integral_constant<int, 4>{}, integral_constant<bool, false>{});
      PSB_STAMPS(15);
      if (SG.flags & 1) stream_issue_upto(st_n + S_);
      const int na = momenta_sums ? 6 : 3;
#pragma unroll
      for (int g = 0; g < 4; ++g)
        if (g < M.ngroups) block_reduce_store(sh, acc4[g], na, M.partials + (size_t)(g * NACC) * nb + b, nb, true);
    }
    if (slot_major && !TAGMAP && !STREAM) {
      double acc4[4][6];
      slot_gather(acc4);
      const int na = momenta_sums ? 6 : 3;
#pragma unroll
      for (int g = 0; g < 4; ++g)
        if (g < M.ngroups) {
          if (ll) block_reduce_store(sh, acc4[g], na, &sh.stage4[g][0], 1, false);
          else block_reduce_store(sh, acc4[g], na, M.partials + (size_t)(g * NACC) * nb + b, nb, true);
        }
    }
    if (slot_major && (TAGMAP || (ll && !STREAM))) {  // TAGMAP: gathered and block-reduced before the dependency wait (see above)
      const int na = momenta_sums ? 6 : 3;
      if (tid < 4 * 6 && tid / 6 < M.ngroups && tid % 6 < na) {
        const size_t at = (size_t)((tid / 6) * NACC + tid % 6) * nb + b;
        if (ll) ll_store(M.ll + at, sh.stage4[tid / 6][tid % 6], ll_stamp);
        else __stcg(M.partials + at, sh.stage4[tid / 6][tid % 6]);
      }
    }

    // groups too large for the register cache (and not slot-ordered): gather loop, slots published
    for (int g = cached ? g_end : g_begin; g < g_end; ++g) {
      const GroupDev& G = M.grp[g];
      const CvDev& cv = M.cv[G.cv];
      const int na = nacc_for_kind(cv.kind, 0, momenta_sums);
      const bool with_p = (na == 6);
      double acc[NACC];
#pragma unroll
      for (int a = 0; a < NACC; ++a) acc[a] = 0.0;
      const uint32_t tlo = max(lo, G.t0), thi = min(hi, G.t1);
#pragma unroll 4
      for (uint32_t t = tlo + tid; t < thi; t += BLOCK) {
        const uint32_t slot = A::slot(S, M, term_tag_of(M, G, t));
        M.term_slot[t] = slot;
        const V3 p = with_p ? A::momentum(S, slot) : v3(0, 0, 0);
        accumulate_term<GENERAL>(cv, with_p, G.inv_n, t, A::position(S, slot), p, acc);
      }
      if (na == 0) continue;
      if (nb > 1) block_reduce_store(sh, acc, na, M.partials + (size_t)(g * NACC) * nb + b, nb, true);
      else block_reduce_store(sh, acc, na, &sh.tot[g][0], 1, false);
    }

    PSB_STAMP(1);
    const RowPlan planA = row_plan(M, 0, 0, tid >> 3, nb);
    if (ll) {
      PSB_STAMP(2);
      PSB_CLK(8);
      if (SLOT) reduce_rows_ll2(M, sh, ll_stamp);
      else reduce_rows_ll(M, sh, planA, ll_stamp);
      if (!pre_early) commit_pre();
      __syncthreads();
    } else {
      grid_sync(M, sh, BAR_A);
      PSB_STAMP(2);
      PSB_CLK(8);
      if (SLOT && nb > 48) reduce_rows_wide(M, sh);
      else reduce_rows(M, sh, 0, 0, planA);
    }
    PSB_CLK(9);
    if (warp == 0) {
      if (any_coord) reduce_coordination(M, sh, lane);
      if (lane < M.ncv) finalize_cv_A<GENERAL>(M, sh, lane);
      PSB_CLK(10);
      if (any_rmsd) {  // does any RMSD need the explicit residual pass?  (uniform over the grid, see finalize_cv_A)
        __syncwarp();
        if (lane == 0) {
          int need = 0;
          for (int c = 0; c < M.ncv; ++c)
            if (M.cv[c].kind == PSB_CV_RMSD && sh.fin.cvx[c][16] != 0.0) need = 1;
          sh.need_pass_b = need;
        }
        __syncwarp();
      }
      if (!any_rmsd || !sh.need_pass_b) post_cv<METHOD, GENERAL>(M, S, sh, lane, cta0);
      PSB_CLK(15);
    } else if (warp == 1 && is_abf && !GENERAL && S.mode != MODE_EVAL) {
#ifdef PSB_PROBE_FIN
      if (S.dbg && lane == 0) S.dbg[(size_t)sh.vb * 16 + 7] = clock64();  // helper warp reaches its barrier
#endif
      abf_helper(M, sh, lane);
#ifdef PSB_PROBE_FIN
      if (S.dbg && lane == 0) S.dbg[(size_t)sh.vb * 16 + 5] = clock64();  // helper warp done (bin + gathers)
#endif
    }
    __syncthreads();
    PSB_STAMP(3);

    // ---- pass B: RMSD residual with the optimal rotation ------------------------------------
    if (any_rmsd && sh.need_pass_b) {
      for (int g = g_begin; g < g_end; ++g) {
        const GroupDev& G = M.grp[g];
        const CvDev& cv = M.cv[G.cv];
        if (cv.kind != PSB_CV_RMSD || sh.fin.cvx[G.cv][16] == 0.0) continue;
        const double* x = sh.fin.cvx[G.cv];
        double acc[1] = {0.0};
        if (cached) {
#pragma unroll
          for (int i = 0; i < CT; ++i)
            if (cg[i] == g) acc[0] += rmsd_residual(cv, x, lo + tid + i * BLOCK, cr[i]);
        } else {
          const uint32_t tlo = max(lo, G.t0), thi = min(hi, G.t1);
#pragma unroll 2
          for (uint32_t t = tlo + tid; t < thi; t += BLOCK)
            acc[0] += rmsd_residual(cv, x, t, A::position(S, __ldcg(M.term_slot + t)));
        }
        if (nb > 1) block_reduce_store(sh, acc, 1, M.partials + (size_t)(g * NACC) * nb + b, nb, true);
        else block_reduce_store(sh, acc, 1, &sh.tot[g][0], 1, false);
      }
      const RowPlan planB = row_plan(M, 1, 0, tid >> 3, nb);
      grid_sync(M, sh, BAR_B);
      reduce_rows(M, sh, 1, 0, planB);
      if (warp == 0) {
        if (lane < M.ncv) finalize_cv_B(M, sh, lane);
        post_cv<METHOD, GENERAL>(M, S, sh, lane, cta0);
      }
      __syncthreads();
    }
  } else {
    commit_pre();
    fin_load(M, sh);
  }
  if (S.mode == MODE_EVAL) {
    if (cta0 && tid == 0) commit_epochs(M, sh);
    return;
  }

  // ---- pass C: J J^T and J p over unique atoms (ABF with position-dependent Jacobians) -----------
  if (is_abf && !abf_fast) {
    const uint32_t uchunk = (M.Su + nb - 1) / nb;
    const uint32_t u0 = min(M.Su, (uint32_t)b * uchunk), u1 = min(M.Su, u0 + uchunk);
    double acc[NACC];
#pragma unroll
    for (int a = 0; a < NACC; ++a) acc[a] = 0.0;
    for (uint32_t u = u0 + tid; u < u1; u += BLOCK) {
      V3 G[KA];
#pragma unroll
      for (int c = 0; c < KA; ++c) G[c] = v3(0, 0, 0);
      uint32_t e0, e1, slot;
      unique_atom<A>(M, S, all_unique, u, e0, e1, slot);
      const V3 r = M.all_point ? v3(0, 0, 0) : A::position(S, slot);
      for (uint32_t e = e0; e < e1; ++e) {
        const uint32_t t = all_unique ? e : __ldg(M.uniq_term + e);
        const int g = all_unique ? (int)__ldg(M.term_group + t) : (int)__ldg(M.uniq_grp + e);
        V3 gv[3];
        const int nc = term_gradient_at(M, sh.fin, t, g, r, gv);
        const int comp0 = M.cv[M.grp[g].cv].comp0;
        for (int j = 0; j < nc; ++j) {
#pragma unroll
          for (int c = 0; c < KA; ++c)
            if (c == comp0 + j) G[c] = G[c] + gv[j];
        }
      }
      const V3 p = A::momentum(S, slot);
      int o = 0;
#pragma unroll
      for (int a = 0; a < KA; ++a)
#pragma unroll
        for (int c = a; c < KA; ++c) acc[o++] += dot(G[a], G[c]);  // 10 entries
#pragma unroll
      for (int a = 0; a < KA; ++a) acc[10 + a] += dot(G[a], p);
    }
    if (nb == 1) block_reduce_store(sh, acc, 14, &sh.tot[VGROUP][0], 1, false);
    else block_reduce_store(sh, acc, 14, M.partials + (size_t)(VGROUP * NACC) * nb + b, nb, true);
    const RowPlan planC = row_plan(M, 2, 14, tid >> 3, nb);
    grid_sync(M, sh, BAR_C);
    reduce_rows(M, sh, 2, 14, planC);
    if (warp == 0) {
      if (lane == 0) {
        int o = 0;
        for (int a = 0; a < KA; ++a)
          for (int c = a; c < KA; ++c) {
            sh.fin.JJt[a * KA + c] = sh.tot[VGROUP][o];
            sh.fin.JJt[c * KA + a] = sh.tot[VGROUP][o];
            ++o;
          }
        for (int a = 0; a < KA; ++a) sh.fin.Jp[a] = sh.tot[VGROUP][10 + a];
      }
      abf_finish(M, S, sh, lane, cta0, false, false);
    }
    __syncthreads();
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
      if (k <= 2) {
        for (int i = tid; i < cnt; i += BLOCK) acc[0] += hill_eval<2>(m, k, xi, cc + (size_t)i * k, hh[i], acc + 1);
      } else if (k <= 4) {
        for (int i = tid; i < cnt; i += BLOCK) acc[0] += hill_eval<4>(m, k, xi, cc + (size_t)i * k, hh[i], acc + 1);
      } else {
        hill_chunk_wide(m, k, xi, cc, hh, cnt, tid, acc);
      }
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
    const RowPlan planH = row_plan(M, 3, 1 + k, tid >> 3, nb);
    grid_sync(M, sh, BAR_H);
    reduce_rows(M, sh, 3, 1 + k, planH);
    if (warp == 0) {
      Fin& f = sh.fin;
      const double V = sh.tot[VGROUP][0];
      if (lane < k) {
        f.F[lane] = sh.tot[VGROUP][1 + lane];
        f.hill[lane] = f.xi[lane];
      }
      __syncwarp();
      if (lane == 0) {
        f.energy = V;
        f.hill[k] = 0.0;
        f.hill[k + 1] = 0.0;
        if (f.deposit) {  // metad.py:219-229, 262-271 (V is the potential BEFORE this hill)
          const double h = m.well_tempered ? m.height * exp(-V / m.deltaT_kB) : m.height;
          f.hill[k] = h;
          f.hill[k + 1] = 1.0;
          if (S.mode == MODE_STEP) {
            if (cta0) {
              const long long idx = sh.pre.idx;
              if (idx < m.ngaussians) {
                M.st.heights[idx] = h;
                for (int c = 0; c < k; ++c) M.st.centers[idx * k + c] = f.xi[c];
              }
              *M.st.idx = idx + 1;
            }
            // the new hill is centred on xi: zero gradient, energy grows by h
            f.energy = V + h;
          }
        }
      }
      if (S.mode == MODE_STEP) finish_force(M, sh, lane, cta0);
    }
    __syncthreads();
  }

  if (WALK && S.mode == MODE_WALKER_BEGIN) {
    if (cta0) {
      if (tid < k + 2) S.record_out[tid] = sh.fin.hill[tid];
      fin_store(M, sh);
      if (tid == 0) commit_epochs(M, sh);
    }
    return;
  }

  // ---- pass D: add newly deposited Gaussians to the grids (metad.py:251-256) ---------------------
  if ((grid_metad && S.mode == MODE_STEP && sh.fin.deposit) || (WALK && S.mode == MODE_WALKER_FINISH)) {
    const bool wfinish = WALK && (S.mode == MODE_WALKER_FINISH);
    const int nrec = wfinish ? S.nwalkers : 1;
    const int rs = k + 2;
    if (grid_metad) {
      // every CTA has read grid_potential[I] (well-tempered height) before any bin is rewritten
      if (S.mode == MODE_STEP && m.well_tempered) grid_sync(M, sh, BAR_PRE_D);
      for (long long p = (long long)b * BLOCK + tid; p < m.grid_points; p += (long long)nb * BLOCK) {
        double x[KA];  // gridded metadynamics: k = grid dimensions <= KA
        long long rem = p;
        for (int d = m.grid_ndim - 1; d >= 0; --d) {
          const int i = (int)(rem % m.g_shape[d]);
          rem /= m.g_shape[d];
          x[d] = mesh_coord(m.grid_kind, i, m.g_lower[d], m.g_upper[d], m.g_shape[d]);
        }
        double val = 0.0, dV[KA] = {};
        for (int r = 0; r < nrec; ++r) {
          const double* rec = wfinish ? S.records + (size_t)r * rs : sh.fin.hill;
          if (rec[k + 1] == 0.0) continue;
          val += hill_eval<KA>(m, k, x, rec, rec[k], dV);  // d/dx of the Gaussian centred on rec
        }
        if (M.st.gp) M.st.gp[p] += val;
        for (int c = 0; c < k; ++c) M.st.gg[p * k + c] += dV[c];
      }
      grid_sync(M, sh, BAR_D);
    }
    if (warp == 0) {
      Fin& f = sh.fin;
      if (wfinish && lane == 0) {
        // append the walkers' hills in rank order (identical on every rank)
        long long idx = sh.pre.idx;
        for (int r = 0; r < nrec; ++r) {
          const double* rec = S.records + (size_t)r * rs;
          if (rec[k + 1] == 0.0) continue;
          if (cta0 && idx < m.ngaussians) {
            M.st.heights[idx] = rec[k];
            for (int c = 0; c < k; ++c) M.st.centers[idx * k + c] = rec[c];
          }
          ++idx;
          if (!grid_metad) {  // the other walkers' new hills do push on this walker
            double dV[MAXK] = {};
            f.energy += hill_eval(m, k, f.xi, rec, rec[k], dV);
            for (int c = 0; c < k; ++c) f.F[c] += dV[c];
          }
        }
        if (cta0) *M.st.idx = idx;
      }
      __syncwarp();
      if (grid_metad) metad_grid_force(M, f, lane);
      finish_force(M, sh, lane, cta0);
    }
    __syncthreads();
  }

  // ABF: this CTA's reads of hist[I] / Fsum[I] have RETURNED (their values were consumed and sit in
  // shared memory, two block barriers ago), so CTA 0 may overwrite the bin once every CTA has said so.
  // Only that write-after-read order is needed: a load that has delivered its value cannot observe a
  // later store, so a relaxed RED with no fence is enough (a release would stall ~0.5 us on a MEMBAR
  // right before the scatter).  Issued here, before the scatter, CTA 0 finds the counter complete when
  // its own scatter is done (no tail wait).
  unsigned long long readers_want = 0;
  const bool abf_commit = (is_abf && nb > 1 && S.mode == MODE_STEP);
  const bool readers = abf_commit || ll;  // LL exchange: also orders CTA 0's state stores after every prologue
  if (readers && tid == 0) {
    red_add_u64(&M.bars[BAR_READERS].arrive, 1ULL);
    sh.bars_used |= 1u << BAR_READERS;
    readers_want = (sh.epoch[BAR_READERS] + 1ULL) * (unsigned long long)nb;
  }
  PSB_STAMP(4);
  // ---- pass S: in-place force scatter ------------------------------------------------------------
  if (will_scatter) {
    const Fin& f = sh.fin;
    const bool dense = GENERAL && (M.bias_dense != nullptr);
    if constexpr (STREAM) {
      // force rows arrive in the ring (most of them were requested before the grid-wide exchange), are updated in
      // shared memory and go back with one bulk store per chunk
      double rs0 = 0.0, rs1 = 0.0, rs2 = 0.0, rs3 = 0.0;
      if (rog_bits) {
        auto scale_of = [&](int g) {
          return ((rog_bits >> g) & 1u) ? -f.F[M.cv[M.grp[g].cv].comp0] * f.cvx[M.grp[g].cv][0] : 0.0;
        };
        rs0 = scale_of(0); rs1 = scale_of(1); rs2 = scale_of(2); rs3 = scale_of(3);
      }
      const int S_ = SG.stages;
      const uint32_t CH = (uint32_t)SG.ch;
      // group of a row: from the nibbles the gather kept (runs of up to 16 rows per lane), else from slot_map again
      const bool nibbles = st_n * STREAM_RPL <= 16;
      constexpr int MG = 4;
      uint32_t wnext[MG][STREAM_RPL];
      auto map_words = [&](int j0, uint32_t (&w)[MG][STREAM_RPL]) {
#pragma unroll
        for (int jj = 0; jj < MG; ++jj) {
          const uint32_t base = st_lo + (uint32_t)(j0 + jj) * CH;
#pragma unroll
          for (int u = 0; u < STREAM_RPL; ++u) {
            const uint32_t sl = base + lane + 32u * u;
            w[jj][u] = (!nibbles && j0 + jj < st_n && sl < st_hi) ? __ldcg(M.slot_map + sl) : 0u;
          }
        }
      };
      map_words(0, wnext);
      for (int j0 = 0; j0 < st_n; j0 += MG) {
        uint32_t wcur[MG][STREAM_RPL];
#pragma unroll
        for (int jj = 0; jj < MG; ++jj)
#pragma unroll
          for (int u = 0; u < STREAM_RPL; ++u) wcur[jj][u] = wnext[jj][u];
        map_words(j0 + MG, wnext);
#pragma unroll
        for (int jj = 0; jj < MG; ++jj) {
        const int j = j0 + jj;
        if (j >= st_n) break;
        const int i = st_n + j;
        const uint32_t base = st_lo + (uint32_t)j * CH;
        const uint32_t cnt = min(CH, st_hi - base), cnt4 = cnt & ~3u;
        unsigned char* stage = stream_next_stage();
#pragma unroll
        for (int u = 0; u < STREAM_RPL; ++u) {
          const uint32_t row = lane + 32u * u;
          int g;
          if (nibbles) g = (int)((grp_nib >> (4 * (j * STREAM_RPL + u))) & 15ULL) - 1;
          else g = ((wcur[jj][u] & ~15u) == stamp) ? (int)(wcur[jj][u] & 15u) - 1 : -1;
          if (row >= cnt || g < 0) continue;
          V3 fo = v3(f.fg[g][0], f.fg[g][1], f.fg[g][2]);
          if (rog_bits && ((rog_bits >> g) & 1u)) {  // -F d<r.r>/dr_i = rs_g * r_i (unwrapped, shape.py:52-57)
            const R4s p4 = reinterpret_cast<const R4s*>(stage + SG.s_pos)[row];
            V3 r = v3((double)p4.x, (double)p4.y, (double)p4.z);
            if (SG.s_img) {
              int ix, iy, iz;
              if (row < cnt4) {
                const int* q = reinterpret_cast<const int*>(stage + SG.s_img) + 3 * row;
                ix = q[0]; iy = q[1]; iz = q[2];
              } else {
                const int* q = reinterpret_cast<const int*>(S.images) + 3 * (size_t)(base + row);
                ix = __ldg(q); iy = __ldg(q + 1); iz = __ldg(q + 2);
              }
              r.x = __dadd_rn(r.x, __dmul_rn(S.L[0], (double)ix));
              r.y = __dadd_rn(r.y, __dmul_rn(S.L[1], (double)iy));
              r.z = __dadd_rn(r.z, __dmul_rn(S.L[2], (double)iz));
            }
            fo = (g == 0 ? rs0 : (g == 1 ? rs1 : (g == 2 ? rs2 : rs3))) * r;
          }
          R4s* fr = reinterpret_cast<R4s*>(stage) + row;  // hoomd.py:229: fp64 sum, rounded once to the force dtype
          R4s q4 = *fr;
          q4.x = (Real)((double)q4.x + fo.x);
          q4.y = (Real)((double)q4.y + fo.y);
          q4.z = (Real)((double)q4.z + fo.z);
          *fr = q4;
        }
        fence_proxy_async();  // generic-proxy writes above -> async-proxy read by the bulk store
        __syncwarp();
        if (lane == 0) {
          tma_bulk_s2g(reinterpret_cast<unsigned char*>(S.forces) + (size_t)base * PB, stage, cnt * PB);
          bulk_commit();
          bulk_wait_read<1>();  // the store of the PREVIOUS chunk has read its stage: that stage may be refilled
        }
        stream_issue_upto(i + S_);  // item (i - 1) + S
        }
      }
      if (lane == 0) bulk_wait_all();
    } else if (slot_major) {
#ifndef PSB_SLOT_SUS
#define PSB_SLOT_SUS 8
#endif
      constexpr int SU = PSB_SLOT_SUS;
      // -F dxi/dr_i = rs_g * r_i for the atoms of RoG group g (scalars, selected without indexing an array)
      double rs0 = 0.0, rs1 = 0.0, rs2 = 0.0, rs3 = 0.0;
      if (rog_bits) {
        auto scale_of = [&](int g) {
          return ((rog_bits >> g) & 1u) ? -f.F[M.cv[M.grp[g].cv].comp0] * f.cvx[M.grp[g].cv][0] : 0.0;
        };
        rs0 = scale_of(0); rs1 = scale_of(1); rs2 = scale_of(2); rs3 = scale_of(3);
      }
      // two copies of the loop: with the groups kept by the gather (sh.rowg) the only loads are the force rows
      auto scatter_rows = [&](auto cached_groups) {
        constexpr bool CG = decltype(cached_groups)::value;
        for (uint32_t base = s_lo + tid; base < s_hi; base += SU * BLOCK) {
          uint32_t e[SU];
          FV fv[SU];
          V3 rp[SU];
#pragma unroll
          for (int u = 0; u < SU; ++u) {
            const uint32_t sl = min(base + u * BLOCK, s_hi - 1);
            if constexpr (CG) e[u] = (uint32_t)sh.rowg[min(base - s_lo + u * BLOCK, (uint32_t)(ROWG_ROWS * BLOCK - 1))];
            else e[u] = slot_word(sl);
            fv[u] = A::load_force(S, sl);
            rp[u] = rog_bits ? A::position(S, sl) : v3(0, 0, 0);
          }
#pragma unroll
          for (int u = 0; u < SU; ++u) {
            const uint32_t sl = base + u * BLOCK;
            const bool hit = CG ? (e[u] != 0u) : ((e[u] & ~15u) == stamp);
            if (sl < s_hi && hit) {
              const int g = (int)(e[u] & 15u) - 1;
              V3 fo = v3(f.fg[g][0], f.fg[g][1], f.fg[g][2]);
              if (rog_bits && ((rog_bits >> g) & 1u)) fo = (g == 0 ? rs0 : (g == 1 ? rs1 : (g == 2 ? rs2 : rs3))) * rp[u];
              A::store_force(S, sl, fv[u], fo);
            }
          }
        }
      };
      if constexpr (is_abf) {
        if (nib_ok) scatter_rows(std::true_type{});
        else scatter_rows(std::false_type{});
      } else {
        scatter_rows(std::false_type{});
      }
    } else if (all_unique && !dense) {
      if (cached) {
#pragma unroll
        for (int i = 0; i < CT; ++i) {
          const int g = cg[i];
          if (g >= 0) {
            const GroupDev& G = M.grp[g];
            const CvDev& cv = M.cv[G.cv];
            V3 fo = v3(f.fg[g][0], f.fg[g][1], f.fg[g][2]);
            if (GENERAL && !is_point_kind(cv.kind))
              fo = (-f.F[cv.comp0]) * nonpoint_gradient(f, cv, G.cv, G.inv_n, lo + tid + i * BLOCK, cr[i]);
            A::store_force(S, cslot[i], cfv[i], fo);
          }
        }
      } else if (!skip_cv) {  // large groups: slots were published in pass A
        for (int g = g_begin; g < g_end; ++g) {
          const GroupDev& G = M.grp[g];
          const CvDev& cv = M.cv[G.cv];
          const uint32_t tlo = max(lo, G.t0), thi = min(hi, G.t1);
          if (!GENERAL || is_point_kind(cv.kind)) {
            const V3 fgv = v3(f.fg[g][0], f.fg[g][1], f.fg[g][2]);
#pragma unroll 4
            for (uint32_t t = tlo + tid; t < thi; t += BLOCK) add_force<A>(S, __ldcg(M.term_slot + t), fgv);
          } else {
            const double Fc = -f.F[cv.comp0];
#pragma unroll 2
            for (uint32_t t = tlo + tid; t < thi; t += BLOCK) {
              const uint32_t slot = __ldcg(M.term_slot + t);
              V3 r = v3(0, 0, 0);
              if (cv.kind != PSB_CV_COORDINATION && cv.kind != PSB_CV_ERMSD) r = A::position(S, slot);
              add_force<A>(S, slot, Fc * nonpoint_gradient(f, cv, G.cv, G.inv_n, t, r));
            }
          }
        }
      } else {
        // walker finish: no pass A ran in this launch, spread all terms over the CTAs
        const uint32_t chunk = (M.T + nb - 1) / nb;
        const uint32_t c0 = min(M.T, (uint32_t)b * chunk), c1 = min(M.T, c0 + chunk);
        for (uint32_t t = c0 + tid; t < c1; t += BLOCK) {
          const int g = __ldg(M.term_group + t);
          const uint32_t slot = __ldcg(M.term_slot + t);
          V3 gv[3];
          const int nc = term_gradient<A>(M, S, f, t, g, slot, gv);
          const int comp0 = M.cv[M.grp[g].cv].comp0;
          V3 fo = v3(0, 0, 0);
          for (int j = 0; j < nc; ++j) fo = fo + (-f.F[comp0 + j]) * gv[j];
          add_force<A>(S, slot, fo);
        }
      }
    } else {
      const uint32_t uchunk = (M.Su + nb - 1) / nb;
      const uint32_t u0 = min(M.Su, (uint32_t)b * uchunk), u1 = min(M.Su, u0 + uchunk);
      const size_t N3 = 3 * (size_t)M.natoms;
      for (uint32_t u = u0 + tid; u < u1; u += BLOCK) {
        uint32_t e0, e1, slot;
        unique_atom<A>(M, S, all_unique, u, e0, e1, slot);
        const FV fv = A::load_force(S, slot);  // in flight while the gradients are assembled
        const V3 r = M.all_point ? v3(0, 0, 0) : A::position(S, slot);
        V3 fo = v3(0, 0, 0);
        for (uint32_t e = e0; e < e1; ++e) {
          const uint32_t t = all_unique ? e : __ldg(M.uniq_term + e);
          const int g = all_unique ? (int)__ldg(M.term_group + t) : (int)__ldg(M.uniq_grp + e);
          V3 gv[3];
          const int nc = term_gradient_at(M, f, t, g, r, gv);
          const int comp0 = M.cv[M.grp[g].cv].comp0;
          for (int j = 0; j < nc; ++j) {
            fo = fo + (-f.F[comp0 + j]) * gv[j];
            if (dense && M.jac_dense) {
              double* J = M.jac_dense + (size_t)(comp0 + j) * N3 + 3 * (size_t)slot;
              atomicAdd(J + 0, gv[j].x); atomicAdd(J + 1, gv[j].y); atomicAdd(J + 2, gv[j].z);
            }
          }
        }
        A::store_force(S, slot, fv, fo);
        if (dense) {
          double* B = M.bias_dense + 3 * (size_t)slot;
          B[0] = fo.x; B[1] = fo.y; B[2] = fo.z;
        }
      }
    }
  }
  PSB_STAMP(5);

  // ---- ABF commit: CTA 0 stores the new count / force sum of bin I (abf.py:216-218) -------------
  if (cta0 && tid == 0) {
    if (readers) grid_wait(M, BAR_READERS, readers_want);
    const Fin& f = sh.fin;
    if (is_abf && S.mode == MODE_STEP) {
      if (f.in_range) {
        M.st.hist[f.flat] = f.hist_new;
        for (int c = 0; c < k; ++c) M.st.Fsum[f.flat * k + c] = f.Fsum_new[c];
        if (M.st.histp) M.st.histp[f.flat] += 1u;  // cff.py:239 (nobody else touches it inside a launch)
      }
      if (ll)  // deferred from abf_finish
        for (int c = 0; c < k; ++c) {
          M.st.force[c] = f.F[c];
          M.st.Wp_[c] = sh.pre.Wp[c];
          M.st.Wp[c] = f.Wp[c];
        }
    }
    if (ll && S.mode != MODE_EVAL) *M.st.ncalls = sh.pre.ncalls + 1;  // deferred from finish_force
    commit_epochs(M, sh);
  }
  PSB_STAMP(6);
}

template <int LAYOUT, typename Real, int METHOD, int CLASS>
__global__ void __launch_bounds__(BLOCK, (CLASS == SC_SLOT || CLASS == SC_TAGS) ? SLOT_CTAS_PER_SM : 2)
step_kernel(const Model* __restrict__ Mglobal, const __grid_constant__ Snap S) {
  step_body<LAYOUT, Real, METHOD, CLASS>(Mglobal, S, (int)gridDim.x, (int)blockIdx.x);
}

// Many independent simulations in ONE cooperative launch (psb_batch_step; umbrella windows,
// umbrella_integration.py:128-203): CTA range [cta0, cta0 + ncta) steps simulation `item`.  Every simulation keeps
// its own model, state, workspaces and grid-barrier counters, so the arithmetic -- and the result, bit for bit --
// is that of its own psb_step launch with ncta CTAs.
struct BatchItem {
  const Model* model;
  Snap snap;
  int32_t cta0, ncta;
};
template <int LAYOUT, typename Real, int METHOD, int CLASS>
__global__ void __launch_bounds__(BLOCK, 2)
step_kernel_batch(const BatchItem* __restrict__ items, const int32_t* __restrict__ cta_item) {
  __shared__ __align__(16) BatchItem it;
  {
    static_assert(sizeof(BatchItem) % 8 == 0, "BatchItem is copied in 8-byte words");
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(items + __ldg(cta_item + blockIdx.x));
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(&it);
    if (threadIdx.x < sizeof(BatchItem) / 8) dst[threadIdx.x] = __ldg(src + threadIdx.x);
  }
  __syncthreads();
  step_body<LAYOUT, Real, METHOD, CLASS>(it.model, it.snap, it.ncta, (int)blockIdx.x - it.cta0);
}

}  // namespace psb
