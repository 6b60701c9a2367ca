// psb_device.cuh -- device-side data model of the fused bias step (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pysages_b200.h"
#include "psb_math.cuh"

namespace psb {

constexpr int MAXK = PSB_MAX_K;
constexpr int KA = PSB_MAX_GRID_DIMS;  // ABF family: k = grid dimensions; J J^T and its solve live in KA x KA storage
constexpr int MAXG = PSB_MAX_GROUPS;
constexpr int MAXCV = PSB_MAX_CVS;
constexpr int NACC = 16;           // fp64 accumulators per group and pass
constexpr int VGROUP = MAXG;       // virtual group slot used by grid-wide method reductions
constexpr int BLOCK = 256;         // threads per CTA of the step kernel
constexpr int NN_STAGE = 1024;     // force-network parameters staged in shared memory per CTA (8 KB)
constexpr int CT = 4;              // terms per thread the step kernel keeps in registers (slot, position, prefetched force)
constexpr int NWARPS = BLOCK / 32;
constexpr int HILL_STAGE_BYTES = 20480;  // per TMA stage (2 stages)
constexpr int PAIR_BLOCK = 256;          // threads per CTA of the pair kernel

struct GroupDev {
  uint32_t t0, t1;  // term range
  int32_t cv;       // owning CV
  int32_t local;    // index of this group among the CV's groups
  double inv_n;     // 1 / group size
  uint32_t tag0;    // first tag
  int32_t contig;   // tags are tag0, tag0 + 1, ... (no index load needed)
  double self_coef; // (atoms of the group, counted with multiplicity^2) / n^2: its share of J J^T
};

struct CvDev {
  int32_t kind, axis, comp0, ncomp, g0, ngroups;
  uint32_t t0, t1;
  const double* ref;  // RMSD: centred reference Q (n,3), device
  const double* w;    // RMSD: weights (n), device, or nullptr (uniform 1/n)
  double sumQ[3];     // RMSD: sum_i Q_i (0 up to rounding for uniform weights)
  double gq;          // RMSD: sum_i |Q_i|^2 of the centred reference
  double r0;          // coordination
  int32_t nn, mm;
  double* grad;       // coordination: per-term gradient workspace (T_cv, 3), device
  double* xi_part;    // coordination: per-CTA partial sums of the pair kernel
  int32_t n_part;
  int32_t pad;
};

struct MethodDev {
  int32_t kind;
  int32_t k;
  double kspring[MAXK * MAXK];
  double center[MAXK];
  double height;
  double sigma[MAXK];
  double periods[MAXK];
  long long stride;
  long long ngaussians;
  int32_t well_tempered;
  int32_t shared_hills;
  double deltaT_kB;  // deltaT * kB
  long long abf_N;
  int32_t use_pinv;
  int32_t has_restraints;
  int32_t hist_only;  // ABF variant 1 (ANN): histogram only, force from the network gradient or zero
  int32_t m_pad;
  double r_lower[MAXK], r_upper[MAXK], r_kl[MAXK], r_ku[MAXK];
  double force_unit_factor;
  int32_t grid_kind, grid_ndim;
  double g_lower[PSB_MAX_GRID_DIMS], g_upper[PSB_MAX_GRID_DIMS];
  double g_h[PSB_MAX_GRID_DIMS], g_inv_h[PSB_MAX_GRID_DIMS];  // bin width (upper - lower) / shape and fl(1 / h)
  int32_t g_shape[PSB_MAX_GRID_DIMS];
  long long grid_points;
};

// Force network of the ABF family (psb_force_net, include/pysages_b200.h): header followed in the
// same allocation by the flat parameter vector.
struct ForceNet {
  int32_t n_dense;
  int32_t widths[PSB_MAX_DENSE + 1];
  int32_t activation;
  int32_t gradient;  // 1: force = std * d(sum of outputs) / d xi (ANN)
  double omega0, omega;  // PSB_ACT_SIN: first layer / other layers
  long long predict_after;
  double mean[MAXK], std[MAXK];
  int32_t w_off[PSB_MAX_DENSE];  // offset of layer l's W in the parameter vector; b follows W
  __host__ __device__ const double* params() const { return reinterpret_cast<const double*>(this + 1); }
};

// Force series of SpectralABF (psb_force_series): header, then the exponents (int32, padded to an even
// count), then [scale, coefficients...] as doubles, all in one allocation.
struct ForceSeries {
  int32_t kind, K, dims, oob_zero;
  long long predict_after;
  int32_t val_off;  // offset of the values, in 8-byte words from the start of the block
  int32_t words;    // size of the whole block in 8-byte words
  __host__ __device__ const int32_t* exponents() const { return reinterpret_cast<const int32_t*>(this + 1); }
  __host__ __device__ const double* values() const { return reinterpret_cast<const double*>(this) + val_off; }
};

enum { FM_NET = 1, FM_SERIES = 2, FM_HIST_ONLY = 4 };

struct StateDev {
  double* xi;        // (k)
  double* F;         // (k) generalized force dV/dxi
  double* energy;    // (1)
  long long* ncalls; // (1)
  // metadynamics
  double* heights;   // (Hpad)
  double* centers;   // (Hpad, k)
  double* sigmas;    // (k)
  double* gp;        // grid_potential (G) or nullptr
  double* gg;        // grid_gradient (G, k) or nullptr
  long long* idx;    // (1)
  // ABF
  uint32_t* hist;    // (G)
  uint32_t* histp;   // (G) CFF: visits since the last training step (cff.py:239), or nullptr
  double* Fsum;      // (G, k)
  double* force;     // (k)
  double* Wp;        // (k)
  double* Wp_;       // (k)
};

// Results of the in-kernel finalizers.  Every CTA computes its own bit-identical copy in shared
// memory (same inputs, same order of operations); the global copy (Model::fin) only carries the
// CV stage from psb_walker_begin to psb_walker_finish.
struct Fin {
  double xi[MAXK];
  double F[MAXK];
  double energy;
  double fg[MAXG][3];       // -sum_c F_c dxi_c/dpoint_g / n_g   (point CVs)
  double gpt[MAXG][3][3];   // dxi_{comp0+j}/dpoint_g
  double cvx[MAXCV][20];    // RMSD: U[0..8], rbar[9..11], sumD[12..14], coef[15], [16] != 0: the residual pass is needed; RoG: coef[0]
  double hill[MAXK + 2];    // candidate / deposited hill: centre[k], height, valid
  double JJt[KA * KA];      // ABF: J J^T (leading dimension KA)
  double Jp[MAXK];          // ABF: J p
  double Wp[MAXK];          // ABF: solve(J J^T, J p)
  double Fsum_new[MAXK];    // ABF: Fsum[I] after this step's scatter-add
  double Fsum_old[MAXK];    // ABF: Fsum[I] as gathered (helper warp)
  double den;               // ABF: max(N, hist[I] after the scatter-add) (helper warp)
  double inv_den;           // ABF: 1 / den, formed off the critical path next to the gathers
  double rforce[MAXK];      // ABF: restraint force when xi is outside the grid (helper warp)
  long long flat;           // clamped flat grid offset of xi
  long long nh;             // number of hills summed this step
  uint32_t I[PSB_MAX_GRID_DIMS];
  uint32_t hist_new;        // ABF: hist[I] after this step's scatter-add
  int32_t in_range;         // ABF: the scatter-add is not dropped
  int32_t oob;
  int32_t deposit;
};

// Method-state scalars every CTA reads before the first grid barrier (CTA 0 overwrites them after it)
struct Pre {
  long long ncalls, idx;
  double Wp[MAXK], Wp_[MAXK], force[MAXK];
};

struct BarrierState {
  unsigned long long arrive;
  unsigned long long pad[15];  // one counter per 128-byte line
};

enum { BAR_A = 0, BAR_B = 1, BAR_C = 2, BAR_H = 3, BAR_D = 4, BAR_READERS = 5, BAR_PRE_D = 6, BAR_SLOT = 7, NBARS = 8 };

// SC_STREAM (psb_step.cuh): shared-memory layout of one ring stage, fixed per handle (psb_create)
constexpr int STREAM_MAX_STAGES = 8;  // per warp
constexpr int STREAM_RPL = 4;         // rows per lane and chunk: a warp chunk is 32 * STREAM_RPL slots
struct StreamGeom {
  int32_t ch;            // slots per warp chunk (32 * STREAM_RPL)
  int32_t stages;        // S, per warp
  uint32_t stage_bytes;  // multiple of 128
  uint32_t g_img, g_vel; // gather chunk = [positions | images | velocities]: byte offsets (0: absent)
  uint32_t s_pos, s_img; // scatter chunk = [forces | positions | images]: byte offsets (RoG groups only; 0: absent)
  int32_t early;         // gather chunks requested before the dependency wait (<= stages)
  int32_t flags;         // experiments: 1 = no scatter chunk is requested before the gather is over
  int32_t pad;
};

struct Model {
  int32_t ncv, k, ngroups, all_unique;
  uint32_t T, Su;
  int32_t any_rmsd, any_coord, abf_fast, all_point;
  int32_t need_term_slot;     // pass A must publish term -> slot (generic scatter, pass C, walkers)
  int32_t hill_chunk;         // hills per TMA stage
  CvDev cv[MAXCV];
  GroupDev grp[MAXG];
  int32_t comp_cv[MAXK];      // CV owning component c
  int32_t cta_lo[MAXG], cta_hi[MAXG];  // gridDim.x > 1: CTAs cta_lo[g] .. cta_hi[g] hold partial sums of group g
  int32_t slot_major;         // gather / scatter in slot order through slot_map (most atoms selected)
  uint32_t stamp_period;      // tests: clear slot_map every `stamp_period` launches as if the 28-bit stamp had wrapped (0: off)
  uint32_t* slot_map;         // (N) slot -> (launch stamp << 4 | group + 1), rewritten every launch
  int32_t row_first[2][MAXG + 1];  // prefix of accumulator rows per group, pass A / pass B
  const uint32_t* ov;         // (MAXG, MAXG) group overlap counts, device memory (ABF with atoms shared between groups)
  const uint32_t* term_tag;   // (T)
  const uint8_t* term_group;  // (T)
  uint32_t* term_slot;        // (T) workspace: slot of each term, written in pass A
  const uint32_t* uniq_ptr;   // (Su+1) CSR over unique atoms -> terms (nullptr if all_unique)
  const uint32_t* uniq_term;  // (T)
  const uint32_t* uniq_tag;   // (Su) particle tag of each unique atom
  const uint8_t* uniq_grp;    // (T) group of uniq_term[e]
  MethodDev m;
  StateDev st;
  Fin* fin;
  double* partials;           // ((MAXG+1) * NACC, gridDim) fp64
  const uint8_t* tag_group;   // slot-ordered class with a slot -> tag array (OpenMM ids, HOOMD tags): tag -> group + 1 (0: not selected), (natoms)
  ulonglong2* ll;             // (MAXG * NACC, gridDim) pass-A partials as two {32 data bits, 32-bit launch stamp} words
  int32_t ll_reduce;          // pass A exchanges its partials through `ll` (no counter barrier): see reduce_rows_ll
  ulonglong2* ll_tot;         // slot-ordered classes: flagged row totals published by the row owners (MAXG * NACC)
  int32_t ll_pad;
  BarrierState* bars;         // (NBARS) monotonic arrival counters
  unsigned long long* epochs; // (NBARS) completed uses of each barrier
  int32_t cross_overlap;      // some atom belongs to two different groups
  int32_t jjt_const;          // ABF: J J^T does not depend on positions; jjt_inv holds its inverse
  double jjt_inv[KA * KA];    // (leading dimension KA)
  const ForceNet* nn;         // ABF family: force network or nullptr (psb_set_force_net)
  int32_t nn_np;              // its parameter count when it fits the shared-memory staging area (NN_STAGE), else 0
  int32_t fs_staged;          // the force series block fits the staging area
  int32_t fm_mode;            // FM_* bits: what abf_finish has to consider besides plain ABF (0 for ABF itself)
  int32_t fm_pad;
  const ForceSeries* fs;      // SpectralABF: force series or nullptr (psb_set_force_series)
  uint32_t* inv_perm;         // OpenMM-CUDA: atom -> slot, rebuilt every step
  StreamGeom sg;              // SC_STREAM ring geometry
  double* bias_dense;         // (N,3) or nullptr
  double* jac_dense;          // (k,3N) or nullptr
  long long natoms;
};

struct Snap {
  const void* positions;
  const void* vel_mass;
  void* forces;
  const void* ids;
  const void* images;
  const void* masses;
  const void* types;
  const void* tags;  // HOOMD: slot -> tag (optional; class SC_TAGS)
  double L[3];
  double dt;
  double inv_dt;  // 1 / dt (host): the finite difference of abf.py:213 multiplies by it instead of dividing
  long long natoms;
  int32_t unwrap;
  int32_t need_momenta;
  int32_t mode;  // 0 full step, 1 evaluate CVs only, 2 walker begin, 3 walker finish
  int32_t nwalkers;
  const double* records;  // walker finish: gathered hill records
  double* record_out;     // walker begin: this walker's record
  long long* dbg;         // optional (PSB_DEBUG_TIMING=1): phase timestamps [2][grid][16], by launch parity
  unsigned long long* cta_seq;  // (gridDim) launches of this handle seen by each CTA; nullptr unless Model::ll_reduce
};

enum { MODE_STEP = 0, MODE_EVAL = 1, MODE_WALKER_BEGIN = 2, MODE_WALKER_FINISH = 3 };

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- engine layouts -------------------------------------------------------------------------
template <int LAYOUT, typename Real>
struct Access;

template <typename Real>
struct Real4;
template <>
struct Real4<float> { using type = float4; };
template <>
struct Real4<double> { using type = double4; };

#ifdef PSB_GATHER_CG
__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldcg(p); }
__device__ __forceinline__ double4 ldg4(const double4* p) {
  const double2 a = __ldcg(reinterpret_cast<const double2*>(p));
  const double2 b = __ldcg(reinterpret_cast<const double2*>(p) + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}
#else
__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }
__device__ __forceinline__ double4 ldg4(const double4* p) {
  const double2 a = __ldg(reinterpret_cast<const double2*>(p));
  const double2 b = __ldg(reinterpret_cast<const double2*>(p) + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}
#endif

// plain (coherent) 16/32-byte accesses for the force array
__device__ __forceinline__ float4 ld4(const float4* p) { return *p; }
__device__ __forceinline__ double4 ld4(const double4* p) {
  const double2 a = *reinterpret_cast<const double2*>(p);
  const double2 b = *(reinterpret_cast<const double2*>(p) + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void st4(float4* p, float4 v) { *p = v; }
__device__ __forceinline__ void st4(double4* p, double4 v) {
  *reinterpret_cast<double2*>(p) = make_double2(v.x, v.y);
  *(reinterpret_cast<double2*>(p) + 1) = make_double2(v.z, v.w);
}

// hoomd.py:148-202
template <typename Real>
struct Access<PSB_LAYOUT_HOOMD, Real> {
  using R4 = typename Real4<Real>::type;
  static __device__ __forceinline__ uint32_t slot(const Snap& s, const Model&, uint32_t tag) {
    return __ldg(reinterpret_cast<const uint32_t*>(s.ids) + tag);
  }
  static __device__ __forceinline__ V3 position(const Snap& s, uint32_t slot) {
    const R4 p = ldg4(reinterpret_cast<const R4*>(s.positions) + slot);
    V3 r = v3((double)p.x, (double)p.y, (double)p.z);
    if (s.unwrap) {  // hoomd.py:179-181: pos + diag(H) * images, in fp64
      const int* im = reinterpret_cast<const int*>(s.images) + 3 * (size_t)slot;
      r.x = __dadd_rn(r.x, __dmul_rn(s.L[0], (double)__ldg(im + 0)));
      r.y = __dadd_rn(r.y, __dmul_rn(s.L[1], (double)__ldg(im + 1)));
      r.z = __dadd_rn(r.z, __dmul_rn(s.L[2], (double)__ldg(im + 2)));
    }
    return r;
  }
  static __device__ __forceinline__ V3 momentum(const Snap& s, uint32_t slot) {
    // hoomd.py:192-196: (M * V) in the array's own dtype, promoted afterwards
    const R4 v = ldg4(reinterpret_cast<const R4*>(s.vel_mass) + slot);
    return v3((double)(Real)(v.w * v.x), (double)(Real)(v.w * v.y), (double)(Real)(v.w * v.z));
  }
  // hoomd.py:229: forces[:, :3] += biases  (fp64 sum, rounded once to the force dtype); split
  // into load / store so callers can batch several independent read-modify-writes
  using FV = R4;
  static __device__ __forceinline__ FV load_force(const Snap& s, uint32_t slot) {
    return ld4(reinterpret_cast<const R4*>(s.forces) + slot);
  }
  static __device__ __forceinline__ void store_force(const Snap& s, uint32_t slot, FV v, V3 b) {
    v.x = (Real)((double)v.x + b.x);
    v.y = (Real)((double)v.y + b.y);
    v.z = (Real)((double)v.z + b.z);
    st4(reinterpret_cast<R4*>(s.forces) + slot, v);
  }
};

// openmm.py:64-89, 100-143 (CUDA platform)
template <typename Real>
struct Access<PSB_LAYOUT_OPENMM_CUDA, Real> {
  using R4 = typename Real4<Real>::type;
  static __device__ __forceinline__ uint32_t slot(const Snap&, const Model& m, uint32_t tag) {
    return __ldcg(m.inv_perm + tag);  // ids.argsort()[tag] (openmm.py:106-107)
  }
  static __device__ __forceinline__ V3 position(const Snap& s, uint32_t slot) {
    const R4 p = ldg4(reinterpret_cast<const R4*>(s.positions) + slot);
    return v3((double)p.x, (double)p.y, (double)p.z);  // never unwrapped (openmm.py:122-123)
  }
  // openmm.py:96-97,125-127: v / invM, or v when invM == 0.  Split into the load and the arithmetic: a streaming
  // loop issues the loads of a whole batch first and divides afterwards (the test on invM and the division's slow
  // path are branches; taken between two loads they serialise the batch -- measured: 48 instead of 11 us per 1e6 rows)
  using VR = R4;
  static __device__ __forceinline__ VR load_velocity(const Snap& s, uint32_t slot) {
    return ldg4(reinterpret_cast<const R4*>(s.vel_mass) + slot);
  }
  static __device__ __forceinline__ V3 momentum_of(VR v) {
    if (v.w == (Real)0) return v3((double)v.x, (double)v.y, (double)v.z);
    return v3((double)(Real)(v.x / v.w), (double)(Real)(v.y / v.w), (double)(Real)(v.z / v.w));
  }
  static __device__ __forceinline__ V3 momentum(const Snap& s, uint32_t slot) { return momentum_of(load_velocity(s, slot)); }
  // openmm.py:141-143,167-169: int64(2**32 * bias.T) added to forces (3, Np)
  struct FV { long long x, y, z; };
  static __device__ __forceinline__ FV load_force(const Snap& s, uint32_t slot) {
    const long long* f = reinterpret_cast<const long long*>(s.forces);
    return FV{f[0 * s.natoms + slot], f[1 * s.natoms + slot], f[2 * s.natoms + slot]};
  }
  static __device__ __forceinline__ void store_force(const Snap& s, uint32_t slot, FV v, V3 b) {
    long long* f = reinterpret_cast<long long*>(s.forces);
    f[0 * s.natoms + slot] = v.x + (long long)(4294967296.0 * b.x);
    f[1 * s.natoms + slot] = v.y + (long long)(4294967296.0 * b.y);
    f[2 * s.natoms + slot] = v.z + (long long)(4294967296.0 * b.z);
  }
};

// (N,3) arrays: LAMMPS (lammps.py:87-113,182-223) and OpenMM CPU-platform style (PLAIN3)
template <typename Real>
struct Plain3 {
  static __device__ __forceinline__ V3 load3(const void* base, uint32_t slot) {
    const Real* p = reinterpret_cast<const Real*>(base) + 3 * (size_t)slot;
    return v3((double)__ldg(p), (double)__ldg(p + 1), (double)__ldg(p + 2));
  }
  struct FV { Real x, y, z; };
  static __device__ __forceinline__ FV load_force(const Snap& s, uint32_t slot) {
    const Real* f = reinterpret_cast<const Real*>(s.forces) + 3 * (size_t)slot;
    return FV{f[0], f[1], f[2]};
  }
  static __device__ __forceinline__ void store_force(const Snap& s, uint32_t slot, FV v, V3 b) {
    Real* f = reinterpret_cast<Real*>(s.forces) + 3 * (size_t)slot;
    f[0] = (Real)((double)v.x + b.x);
    f[1] = (Real)((double)v.y + b.y);
    f[2] = (Real)((double)v.z + b.z);
  }
};

template <typename Real>
struct Access<PSB_LAYOUT_LAMMPS, Real> {
  static __device__ __forceinline__ uint32_t slot(const Snap& s, const Model&, uint32_t tag) {
    return (uint32_t)__ldg(reinterpret_cast<const int*>(s.ids) + tag + 1);  // ids = tags_map[1:]
  }
  static __device__ __forceinline__ V3 position(const Snap& s, uint32_t slot) {
    V3 r = Plain3<Real>::load3(s.positions, slot);
    if (s.unwrap) {  // lammps.py:188-202: (image >> bits & mask) - kImgMax, SMALLBIG packing
      const int im = __ldg(reinterpret_cast<const int*>(s.images) + slot);
      const int ix = (im & 1023) - 512, iy = ((im >> 10) & 1023) - 512, iz = (im >> 20) - 512;
      r.x = __dadd_rn(r.x, __dmul_rn(s.L[0], (double)ix));
      r.y = __dadd_rn(r.y, __dmul_rn(s.L[1], (double)iy));
      r.z = __dadd_rn(r.z, __dmul_rn(s.L[2], (double)iz));
    }
    return r;
  }
  static __device__ __forceinline__ V3 momentum(const Snap& s, uint32_t slot) {
    // lammps.py:213-217: masses[types] * v
    const int ty = __ldg(reinterpret_cast<const int*>(s.types) + slot);
    const double mass = __ldg(reinterpret_cast<const double*>(s.masses) + ty);
    const V3 v = Plain3<Real>::load3(s.vel_mass, slot);
    return v3(mass * v.x, mass * v.y, mass * v.z);
  }
  using FV = typename Plain3<Real>::FV;
  static __device__ __forceinline__ FV load_force(const Snap& s, uint32_t slot) {
    return Plain3<Real>::load_force(s, slot);
  }
  static __device__ __forceinline__ void store_force(const Snap& s, uint32_t slot, FV v, V3 b) {
    Plain3<Real>::store_force(s, slot, v, b);
  }
};

template <typename Real>
struct Access<PSB_LAYOUT_PLAIN3, Real> {
  static __device__ __forceinline__ uint32_t slot(const Snap& s, const Model&, uint32_t tag) {
    return s.ids ? (uint32_t)__ldg(reinterpret_cast<const int*>(s.ids) + tag) : tag;
  }
  static __device__ __forceinline__ V3 position(const Snap& s, uint32_t slot) {
    V3 r = Plain3<Real>::load3(s.positions, slot);
    if (s.unwrap && s.images) {
      const int* im = reinterpret_cast<const int*>(s.images) + 3 * (size_t)slot;
      r.x = __dadd_rn(r.x, __dmul_rn(s.L[0], (double)__ldg(im + 0)));
      r.y = __dadd_rn(r.y, __dmul_rn(s.L[1], (double)__ldg(im + 1)));
      r.z = __dadd_rn(r.z, __dmul_rn(s.L[2], (double)__ldg(im + 2)));
    }
    return r;
  }
  // openmm.py:96-97,125-127 with (velocities, invMass) stored separately; load and arithmetic split as for the
  // OpenMM-CUDA layout (the test on invMass and the division are branches: they must not sit between two loads)
  struct VR { V3 v; Real im; };
  static __device__ __forceinline__ VR load_velocity(const Snap& s, uint32_t slot) {
    return VR{Plain3<Real>::load3(s.vel_mass, slot), __ldg(reinterpret_cast<const Real*>(s.masses) + slot)};
  }
  static __device__ __forceinline__ V3 momentum_of(VR r) {
    if (r.im == (Real)0) return r.v;
    return v3((double)(Real)((Real)r.v.x / r.im), (double)(Real)((Real)r.v.y / r.im),
              (double)(Real)((Real)r.v.z / r.im));
  }
  static __device__ __forceinline__ V3 momentum(const Snap& s, uint32_t slot) { return momentum_of(load_velocity(s, slot)); }
  using FV = typename Plain3<Real>::FV;
  static __device__ __forceinline__ FV load_force(const Snap& s, uint32_t slot) {
    return Plain3<Real>::load_force(s, slot);
  }
  static __device__ __forceinline__ void store_force(const Snap& s, uint32_t slot, FV v, V3 b) {
    Plain3<Real>::store_force(s, slot, v, b);
  }
};

}  // namespace psb
