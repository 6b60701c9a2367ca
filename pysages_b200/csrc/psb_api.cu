// psb_api.cu -- host side of the C ABI declared in include/pysages_b200.h.
// Builds the device tables once (psb_create), then every psb_step is: [OpenMM inverse
// permutation] [coordination gather + pair kernels] one fused cooperative step kernel, all on the
// caller's stream, no host synchronisation.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "psb_launch.cuh"
#include "psb_step.cuh"

using namespace psb;

namespace {

thread_local std::string g_error;

psb_status fail(psb_status code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_error = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return fail(PSB_ERR_CUDA, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(_e),      \
                  __FILE__, __LINE__, #expr);                                               \
  } while (0)

// G vectors of the REFERENCE structure (orientation.py:177-288), the constant half of the eRMSD sum: (n, n, 4)
std::vector<double> ermsd_reference_g(const double* ref, int n, double cutoff, double a, double b) {
  struct F3 { double x, y, z; };
  auto sub = [](F3 p, F3 q) { return F3{p.x - q.x, p.y - q.y, p.z - q.z}; };
  auto dotp = [](F3 p, F3 q) { return p.x * q.x + p.y * q.y + p.z * q.z; };
  auto crossp = [](F3 p, F3 q) { return F3{p.y * q.z - p.z * q.y, p.z * q.x - p.x * q.z, p.x * q.y - p.y * q.x}; };
  auto unit = [&](F3 p) { const double s = 1.0 / std::sqrt(dotp(p, p)); return F3{p.x * s, p.y * s, p.z * s}; };
  std::vector<F3> o(n), xh(n), yh(n), zh(n);
  for (int j = 0; j < n; ++j) {
    const double* r = ref + 9 * (size_t)j;
    const F3 r0{r[0], r[1], r[2]}, r1{r[3], r[4], r[5]}, r2{r[6], r[7], r[8]};
    o[j] = F3{(r0.x + r1.x + r2.x) / 3.0, (r0.y + r1.y + r2.y) / 3.0, (r0.z + r1.z + r2.z) / 3.0};
    xh[j] = unit(sub(r0, o[j]));
    zh[j] = unit(crossp(xh[j], sub(r1, o[j])));
    yh[j] = unit(crossp(zh[j], xh[j]));
  }
  const double gamma = 3.141592653589793 / cutoff;
  std::vector<double> G((size_t)n * n * 4, 0.0);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      if (i == j) continue;  // masked self entry: outside the cutoff
      const F3 p = sub(o[i], o[j]);
      const F3 rt{dotp(p, xh[j]) / a, dotp(p, yh[j]) / a, dotp(p, zh[j]) / b};
      const double rho = std::sqrt(dotp(rt, rt));
      if (!(cutoff - rho > 0.0)) continue;
      const double inv = rho != 0.0 ? 1.0 / rho : 0.0;
      const double s = std::sin(gamma * rho) * inv / gamma;
      double* g = &G[((size_t)i * n + j) * 4];
      g[0] = s * rt.x; g[1] = s * rt.y; g[2] = s * rt.z; g[3] = (1.0 + std::cos(gamma * rho)) / gamma;
    }
  return G;
}

// ERMSDCG (orientation.py:380-560): the ideal C2 / C4 / C6 atoms of every base from its three coarse-grained
// sites: frame of the sites (orientation.py:177-216), then origin + local_coordinates[sequence] . axes
std::vector<double> ermsd_reconstruct(const double* sites, int n, const double* lc36, const double* seq) {
  std::vector<double> out(9 * (size_t)n);
  for (int j = 0; j < n; ++j) {
    const double* r = sites + 9 * (size_t)j;
    double o[3], x[3], c[3], z[3], y[3];
    for (int d = 0; d < 3; ++d) o[d] = (r[d] + r[3 + d] + r[6 + d]) / 3.0;
    double nx = 0.0;
    for (int d = 0; d < 3; ++d) { x[d] = r[d] - o[d]; nx += x[d] * x[d]; }
    for (int d = 0; d < 3; ++d) { x[d] /= std::sqrt(nx); c[d] = r[3 + d] - o[d]; }
    z[0] = x[1] * c[2] - x[2] * c[1]; z[1] = x[2] * c[0] - x[0] * c[2]; z[2] = x[0] * c[1] - x[1] * c[0];
    const double nz = std::sqrt(z[0] * z[0] + z[1] * z[1] + z[2] * z[2]);
    for (int d = 0; d < 3; ++d) z[d] /= nz;
    y[0] = z[1] * x[2] - z[2] * x[1]; y[1] = z[2] * x[0] - z[0] * x[2]; y[2] = z[0] * x[1] - z[1] * x[0];
    const double ny = std::sqrt(y[0] * y[0] + y[1] * y[1] + y[2] * y[2]);
    for (int d = 0; d < 3; ++d) y[d] /= ny;
    const double* lc = lc36 + 9 * (int)seq[j];
    for (int m = 0; m < 3; ++m)
      for (int d = 0; d < 3; ++d)
        out[9 * (size_t)j + 3 * m + d] = o[d] + lc[3 * m] * x[d] + lc[3 * m + 1] * y[d] + lc[3 * m + 2] * z[d];
  }
  return out;
}

struct CoordWork {
  int cv = -1;
  int nA = 0, nB = 0, na_pad = 0, nb_pad = 0, nrt = 0, ncs = 0, R = 1, H = 3;
  double* xa = nullptr;
  double* xb = nullptr;
};

}  // namespace

struct psb_handle {
  Model model;
  Model* d_model = nullptr;  // device copy read by the step kernel (uploaded at the end of psb_create)
  char* d_nn = nullptr;      // ForceNet header + parameters (psb_set_force_net)
  size_t nn_capacity = 0;
  unsigned long long* d_fs = nullptr;  // ForceSeries block (psb_set_force_series)
  size_t fs_capacity = 0;
  unsigned long long* d_cta_seq = nullptr;  // per-CTA launch counters (LL exchange)
  psb_layout_desc layout;
  psb_method_desc method;
  std::vector<void*> allocs;
  std::vector<CoordWork> coord;
  std::vector<int> ermsd;  // CVs evaluated by ermsd_kernel before the step kernel
  int device = 0;
  int num_sms = 0;
  int grid = 1;
  size_t dyn_smem = 0;
  int cooperative = 0;
  int pdl = 1;           // programmatic dependent launch between consecutive step kernels (psb_set_option)
  int use_graph = 1;     // psb_run_cycle may capture / replay a CUDA graph (psb_set_option "graph")
  int step_method = 0;   // StepMethod code of the kernel instantiation
  int step_general = 0;  // StepClass code of the kernel instantiation
  long long launches = 0;
  long long host_ncalls = 0;  // mirror of the device ncalls (deterministic)
  long long algo_bytes = 0;
  long long* dbg = nullptr;  // PSB_DEBUG_TIMING=1: [2][grid][16] phase timestamps
  // psb_run_cycle: one pass over the snapshot ring captured as a CUDA graph
  cudaGraphExec_t cycle_exec = nullptr;
  std::vector<psb_snapshot> cycle_key;
  cudaStream_t cycle_stream = nullptr;
  long long cycle_launches = 0;  // kernel launches inside one graph replay
  int graph_replays = 0;         // 1 when the last psb_run_cycle replayed the graph
  int record_doubles = 0;
  // staging for psb_step_host
  psb_snapshot dev_snap{};
  bool have_staging = false;
  bool tags_class = false;  // the SC_TAGS instantiation is co-resident at this handle's grid (HOOMD, slot-ordered)
  // Slot-ordered handles planned for the tag array run ONE CTA per SM when steps are replayed back to back
  // (psb_run_cycle: the next step's gather overlaps this step's tail) and TWO when an engine steps them one at a
  // time (psb_step: every launch stands alone); grid_pipe == 0: one grid for both.  See replan_grid.
  int grid_pipe = 0, grid_plain = 0;
  std::vector<std::pair<const void*, size_t>> host_ranges;  // engine arrays seen by psb_step_host (and page-locked by it)
  std::vector<void*> host_registered;                        // ... the ones this handle registered itself
  size_t hpad = 0;
  // psb_step_host with S << N: only the rows of the selected atoms travel (see stage_selected_rows)
  std::vector<uint32_t> sel_tags;       // sorted unique tags of every atom a CV touches
  bool compact_ok = false;              // worth it and valid for this handle (decided in psb_create)
  const void* d_ids_compact = nullptr;  // device: tag -> row of the compact arrays (the layout's own `ids` convention)
  char* pin = nullptr;                  // pinned host staging: [positions | images | velocities | types-or-masses | forces] rows
  size_t pin_bytes = 0;
  std::vector<uint32_t> sel_slots;      // this step's engine slots of sel_tags
};

namespace {

template <typename T>
psb_status dev_alloc(psb_handle* h, T** out, size_t count, bool zero = true) {
  void* p = nullptr;
  const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
  CUDA_TRY(cudaMalloc(&p, bytes));
  if (zero) CUDA_TRY(cudaMemset(p, 0, bytes));
  h->allocs.push_back(p);
  *out = reinterpret_cast<T*>(p);
  return PSB_OK;
}

template <typename T>
psb_status dev_upload(psb_handle* h, T** out, const std::vector<T>& v) {
  psb_status st = dev_alloc(h, out, v.size(), false);
  if (st != PSB_OK) return st;
  if (!v.empty()) CUDA_TRY(cudaMemcpy(*out, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return PSB_OK;
}

int expected_groups(int kind) {
  switch (kind) {
    case PSB_CV_COMPONENT: case PSB_CV_ROG: case PSB_CV_RMSD: case PSB_CV_GYRATION: case PSB_CV_RING:
    case PSB_CV_ERMSD: return 1;
    case PSB_CV_DISTANCE: case PSB_CV_DISPLACEMENT: case PSB_CV_COORDINATION: return 2;
    case PSB_CV_ANGLE: return 3;
    case PSB_CV_DIHEDRAL: return 4;
    default: return -1;
  }
}

const StepVTable* pick_vtable(const psb_layout_desc& L) {
  const bool f32 = L.real_bytes == 4;
  switch (L.layout) {
    case PSB_LAYOUT_HOOMD: return f32 ? &vt_hoomd_f32 : &vt_hoomd_f64;
    case PSB_LAYOUT_OPENMM_CUDA: return f32 ? &vt_openmm_f32 : &vt_openmm_f64;
    case PSB_LAYOUT_LAMMPS: return &vt_lammps_f64;
    case PSB_LAYOUT_PLAIN3: return f32 ? &vt_plain3_f32 : &vt_plain3_f64;
    default: return nullptr;
  }
}

bool is_point(int kind) { return kind >= PSB_CV_COMPONENT && kind <= PSB_CV_DIHEDRAL; }

Snap make_snap(const psb_handle* h, const psb_snapshot* s, int mode) {
  Snap S{};
  S.positions = s->positions;
  S.vel_mass = s->vel_mass;
  S.forces = s->forces;
  S.ids = s->ids;
  S.images = s->images;
  S.masses = s->masses;
  S.types = s->types;
  S.tags = s->tags;
  for (int d = 0; d < 3; ++d) S.L[d] = s->box_diag[d];
  S.dt = s->dt;
  S.inv_dt = s->dt != 0.0 ? 1.0 / s->dt : 0.0;
  S.natoms = h->layout.natoms;
  S.unwrap = h->layout.unwrap && h->layout.layout != PSB_LAYOUT_OPENMM_CUDA;
  S.need_momenta = h->layout.need_momenta;
  S.mode = mode;
  S.dbg = h->dbg ? h->dbg + (size_t)(h->launches & 1) * 16 * h->grid : nullptr;
  S.cta_seq = h->model.ll_reduce ? h->d_cta_seq : nullptr;
  return S;
}

psb_status check_snapshot(const psb_handle* h, const psb_snapshot* s, bool need_forces) {
  if (!s) return fail(PSB_ERR_VALUE, "snapshot is NULL");
  if (!s->positions) return fail(PSB_ERR_VALUE, "snapshot.positions is NULL");
  if (need_forces && !s->forces && h->method.kind != PSB_METHOD_UNBIASED)
    return fail(PSB_ERR_VALUE, "snapshot.forces is NULL");
  const int L = h->layout.layout;
  if ((L == PSB_LAYOUT_HOOMD || L == PSB_LAYOUT_OPENMM_CUDA || L == PSB_LAYOUT_LAMMPS) && !s->ids)
    return fail(PSB_ERR_VALUE, "snapshot.ids is NULL");
  if (h->layout.unwrap && (L == PSB_LAYOUT_HOOMD || L == PSB_LAYOUT_LAMMPS) && !s->images)
    return fail(PSB_ERR_VALUE, "snapshot.images is NULL but the method requires box unwrapping");
  if (h->layout.need_momenta) {
    if (!s->vel_mass) return fail(PSB_ERR_VALUE, "snapshot.vel_mass is NULL but momenta are required");
    if ((L == PSB_LAYOUT_LAMMPS && (!s->masses || !s->types)) || (L == PSB_LAYOUT_PLAIN3 && !s->masses))
      return fail(PSB_ERR_VALUE, "snapshot.masses/types are NULL but momenta are required");
    if (!(s->dt > 0.0)) return fail(PSB_ERR_VALUE, "snapshot.dt must be positive");
  }
  return PSB_OK;
}

// Enqueue everything one step needs.  mode: MODE_*
psb_status enqueue(psb_handle* h, const psb_snapshot* s, int mode, cudaStream_t stream,
                   const double* records, int nwalkers, double* record_out) {
  Snap S = make_snap(h, s, mode);
  S.records = records;
  S.nwalkers = nwalkers;
  S.record_out = record_out;
  const Model& M = h->model;
  const bool eval_cvs = (mode != MODE_WALKER_FINISH);
  // A helper kernel of this library in front of the step kernel (inverse permutation, coordination gather + pairs,
  // eRMSD) writes tables the step kernel reads in its EARLY part, before griddepcontrol.wait (tag -> slot through
  // inv_perm in the register-cached gather).  Programmatic dependent launch only makes the previous kernel's
  // writes visible after that wait, so such a step kernel is launched in plain stream order instead.
  bool helper_before = false;

  if (eval_cvs && h->layout.layout == PSB_LAYOUT_OPENMM_CUDA && !M.slot_major) {  // slot-ordered class: no inverse needed
    const long long n = h->layout.natoms;
    const int blocks = (int)std::min<long long>((n + 255) / 256, 4LL * h->num_sms);
    launch_inverse_permutation(reinterpret_cast<const int*>(s->ids), M.inv_perm, n, blocks, stream);
    ++h->launches;
    helper_before = true;
  }
  if (eval_cvs) {
    for (const CoordWork& cw : h->coord) {
      pick_vtable(h->layout)->launch_gather(h->d_model, &S, cw.cv, cw.xa, cw.na_pad, cw.xb, cw.nb_pad, stream);
      const CvDev& cv = M.cv[cw.cv];
      pair_launcher(cw.R, cw.H)(cw.nrt * cw.ncs, stream, cw.xa, cw.na_pad, cw.xb, cw.nb_pad, cw.nA, cw.nB,
                                cw.nrt, 1.0 / (cv.r0 * cv.r0), cv.grad, cv.grad + 3 * (size_t)cw.nA,
                                cv.xi_part);
      h->launches += 2;
      helper_before = true;
    }
  }
  if (eval_cvs)
    for (int c : h->ermsd) {
      pick_vtable(h->layout)->launch_ermsd(h->d_model, &S, c, stream);
      ++h->launches;
      helper_before = true;
    }
  if (M.bias_dense && mode != MODE_EVAL && mode != MODE_WALKER_BEGIN) {
    CUDA_TRY(cudaMemsetAsync(M.bias_dense, 0, sizeof(double) * 3 * (size_t)h->layout.natoms, stream));
    if (M.jac_dense)
      CUDA_TRY(cudaMemsetAsync(M.jac_dense, 0, sizeof(double) * 3 * (size_t)h->layout.natoms * M.k, stream));
  }
  int cls = h->step_general;
  size_t dyn = h->dyn_smem;
  if (cls == SC_STREAM) {  // bulk copies need 16-byte aligned streams; anything else takes the register-staged class
    const uintptr_t bits = (uintptr_t)s->positions | (uintptr_t)s->forces | (S.unwrap ? (uintptr_t)s->images : 0) |
                           (S.need_momenta ? (uintptr_t)s->vel_mass : 0);
    if (bits & 15u) { cls = SC_SLOT; dyn = 0; }
  }
  // the engine's slot -> tag array came along: no tag -> slot inversion pass (HOOMD layout, psb_snapshot.tags)
  if (cls == SC_SLOT && h->tags_class && s->tags) cls = SC_TAGS;
  CUDA_TRY(pick_vtable(h->layout)->launch_step(h->step_method, cls, h->d_model, &S, h->grid, dyn, stream,
                                               h->pdl && !helper_before));
  ++h->launches;
  return PSB_OK;
}

// Switch a dual-grid slot-ordered handle between its two launch shapes.  Everything whose meaning depends on the number
// of CTAs restarts, stream-ordered behind the launches that used the old shape: barrier arrival counters and epochs
// (epoch 0 also makes the next rtag-only launch clear slot_map: its stamps restart), per-CTA launch counts and the
// flagged exchange words (stamp 0 is never current), the model's CTA ranges; a captured cycle graph is dropped.
psb_status replan_grid(psb_handle* h, bool pipelined, cudaStream_t stream) {
  if (!h->grid_pipe) return PSB_OK;
  const int target = pipelined ? h->grid_pipe : h->grid_plain;
  if (h->grid == target) return PSB_OK;
  Model& M = h->model;
  const size_t gmax = (size_t)std::max(h->grid_pipe, h->grid_plain);
  h->grid = target;
  for (int g = 0; g < MAXG; ++g) M.cta_hi[g] = g < M.ngroups ? target - 1 : -1;
  CUDA_TRY(cudaMemsetAsync(M.bars, 0, sizeof(*M.bars) * NBARS, stream));
  CUDA_TRY(cudaMemsetAsync(M.epochs, 0, sizeof(*M.epochs) * 16, stream));
  if (h->d_cta_seq) CUDA_TRY(cudaMemsetAsync(h->d_cta_seq, 0, sizeof(*h->d_cta_seq) * gmax, stream));
  if (M.ll) CUDA_TRY(cudaMemsetAsync(M.ll, 0, sizeof(*M.ll) * (size_t)MAXG * NACC * gmax, stream));
  if (M.ll_tot) CUDA_TRY(cudaMemsetAsync(M.ll_tot, 0, sizeof(*M.ll_tot) * (size_t)MAXG * NACC, stream));
  CUDA_TRY(cudaMemcpyAsync(h->d_model, &h->model, sizeof(Model), cudaMemcpyHostToDevice, stream));
  if (h->cycle_exec) cudaGraphExecDestroy(h->cycle_exec);
  h->cycle_exec = nullptr;
  h->cycle_key.clear();
  return PSB_OK;
}

bool host_next_deposits(const psb_handle* h) {
  if (h->method.kind != PSB_METHOD_METAD) return false;
  const long long n = h->host_ncalls + 1;
  return n > 1 && (n % h->method.stride == 1);
}

}  // namespace

extern "C" {

const char* psb_last_error(void) { return g_error.c_str(); }
psb_status psbi_fail(psb_status code, const char* msg) { return fail(code, "%s", msg); }  // for the other translation units
int32_t psb_abi_version(void) { return 2; }

int32_t psb_wants_tags(psb_handle* h) { return (h && h->tags_class) ? 1 : 0; }

void psb_destroy(psb_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  for (void* p : h->host_registered) cudaHostUnregister(p);
  if (h->pin) cudaFreeHost(h->pin);
  if (h->cycle_exec) cudaGraphExecDestroy(h->cycle_exec);
  for (void* p : h->allocs) cudaFree(p);
  delete h;
}

psb_status psb_create(const psb_cv_desc* cvs, int32_t ncvs, const psb_method_desc* method,
                      const psb_layout_desc* layout, psb_handle** out) {
  if (!cvs || !method || !layout || !out) return fail(PSB_ERR_VALUE, "NULL argument");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(PSB_ERR_CUDA, "no CUDA device available: pysages_b200 has no CPU fallback");
  if (ncvs < 1 || ncvs > PSB_MAX_CVS)
    return fail(PSB_ERR_VALUE, "between 1 and %d collective variables are supported (got %d)", PSB_MAX_CVS, ncvs);
  if (layout->real_bytes != 4 && layout->real_bytes != 8)
    return fail(PSB_ERR_VALUE, "layout.real_bytes must be 4 or 8");
  if (layout->layout == PSB_LAYOUT_LAMMPS && layout->real_bytes != 8)
    return fail(PSB_ERR_VALUE, "the LAMMPS layout is double precision");
  if (!pick_vtable(*layout)) return fail(PSB_ERR_VALUE, "Invalid backend layout %d", layout->layout);
  if (layout->natoms < 1) return fail(PSB_ERR_VALUE, "layout.natoms must be positive");

  psb_handle* h = new psb_handle();
  h->layout = *layout;
  h->method = *method;
  cudaGetDevice(&h->device);
  cudaDeviceProp prop{};
  cudaGetDeviceProperties(&prop, h->device);
  h->num_sms = prop.multiProcessorCount;
  if (prop.major < 10) {
    delete h;
    return fail(PSB_ERR_CUDA, "pysages_b200 kernels are built for sm_100a; device is sm_%d%d", prop.major, prop.minor);
  }
  Model& M = h->model;
  memset(&M, 0, sizeof(M));
  M.ncv = ncvs;
  M.natoms = layout->natoms;

  auto bail = [&](psb_status st) {
    psb_destroy(h);
    return st;
  };

  // ---- terms, groups, components (colvars/core.py:184-207, 277-295) -------------------------
  std::vector<uint32_t> term_tag;
  std::vector<uint8_t> term_group;
  int k = 0, ng = 0;
  M.all_point = 1;
  for (int c = 0; c < ncvs; ++c) {
    const psb_cv_desc& d = cvs[c];
    const int want = expected_groups(d.kind);
    if (want < 0) return bail(fail(PSB_ERR_VALUE, "unknown collective variable kind %d", d.kind));
    if (d.n_groups != want)
      return bail(fail(PSB_ERR_VALUE, "Exactly %d indices or groups must be provided (got %d)", want, d.n_groups));
    if (!d.tags || !d.group_offsets) return bail(fail(PSB_ERR_VALUE, "CV %d has no indices", c));
    if (ng + d.n_groups > PSB_MAX_GROUPS)
      return bail(fail(PSB_ERR_VALUE, "too many atom groups (max %d)", PSB_MAX_GROUPS));
    CvDev& cv = M.cv[c];
    cv.kind = d.kind;
    cv.axis = d.axis;
    if (d.kind == PSB_CV_COMPONENT && (d.axis < 0 || d.axis > 2))
      return bail(fail(PSB_ERR_RUNTIME, "Invalid Cartesian axis %d index choose 0 (X), 1 (Y), 2 (Z)", d.axis));
    cv.comp0 = k;
    cv.ncomp = d.kind == PSB_CV_DISPLACEMENT ? 3 : 1;
    k += cv.ncomp;
    if (k > PSB_MAX_K) return bail(fail(PSB_ERR_VALUE, "at most %d CV components are supported", PSB_MAX_K));
    for (int j = cv.comp0; j < k; ++j) M.comp_cv[j] = c;
    cv.g0 = ng;
    cv.ngroups = d.n_groups;
    cv.t0 = (uint32_t)term_tag.size();
    for (int g = 0; g < d.n_groups; ++g) {
      const uint32_t a = d.group_offsets[g], b = d.group_offsets[g + 1];
      if (b <= a) return bail(fail(PSB_ERR_VALUE, "CV %d: group %d is empty", c, g));
      GroupDev& G = M.grp[ng];
      G.t0 = (uint32_t)term_tag.size();
      for (uint32_t i = a; i < b; ++i) {
        if ((long long)d.tags[i] >= layout->natoms)
          return bail(fail(PSB_ERR_VALUE, "CV %d: particle index %u out of range (natoms = %lld)", c, d.tags[i],
                           (long long)layout->natoms));
        term_tag.push_back(d.tags[i]);
        term_group.push_back((uint8_t)ng);
      }
      G.t1 = (uint32_t)term_tag.size();
      G.tag0 = d.tags[a];
      G.contig = 1;
      for (uint32_t i = a; i < b; ++i)
        if (d.tags[i] != d.tags[a] + (i - a)) G.contig = 0;
      G.cv = c;
      G.local = g;
      G.inv_n = 1.0 / (double)(b - a);
      ++ng;
    }
    cv.t1 = (uint32_t)term_tag.size();
    if (!is_point(d.kind)) M.all_point = 0;
    if (d.kind == PSB_CV_RMSD) M.any_rmsd = 1;
    if (d.kind == PSB_CV_COORDINATION || d.kind == PSB_CV_ERMSD) M.any_coord = 1;  // value / gradient from a pre-kernel
  }
  M.k = k;
  M.ngroups = ng;
  M.T = (uint32_t)term_tag.size();
  {
    h->sel_tags = term_tag;
    std::sort(h->sel_tags.begin(), h->sel_tags.end());
    h->sel_tags.erase(std::unique(h->sel_tags.begin(), h->sel_tags.end()), h->sel_tags.end());
  }

  // ---- method validation -----------------------------------------------------------------------
  MethodDev& m = M.m;
  m.kind = method->kind;
  m.k = k;
  const psb_grid_desc& gd = method->grid;
  const bool gridded = gd.kind != PSB_GRID_NONE;
  if (method->kind < PSB_METHOD_UNBIASED || method->kind > PSB_METHOD_ABF)
    return bail(fail(PSB_ERR_VALUE, "unknown method kind %d", method->kind));
  if (method->kind == PSB_METHOD_ABF && !gridded) return bail(fail(PSB_ERR_VALUE, "ABF requires a grid"));
  if (gridded && (method->kind == PSB_METHOD_ABF || method->kind == PSB_METHOD_METAD)) {
    if (gd.ndim < 1 || gd.ndim > PSB_MAX_GRID_DIMS)
      return bail(fail(PSB_ERR_VALUE, "grids have between 1 and %d dimensions (got %d)", PSB_MAX_GRID_DIMS, gd.ndim));
    if (gd.ndim != k)  // methods/core.py:435-438
      return bail(fail(PSB_ERR_VALUE, "Grid and Collective Variable dimensions must match."));
    m.grid_kind = gd.kind;
    m.grid_ndim = gd.ndim;
    m.grid_points = 1;
    for (int d = 0; d < gd.ndim; ++d) {
      if (gd.shape[d] < 1) return bail(fail(PSB_ERR_VALUE, "grid shape must be positive"));
      m.g_lower[d] = gd.lower[d];
      m.g_upper[d] = gd.upper[d];
      m.g_h[d] = (gd.upper[d] - gd.lower[d]) / (double)gd.shape[d];  // the same two IEEE divisions as grid_index_1d
      m.g_inv_h[d] = 1.0 / m.g_h[d];
      m.g_shape[d] = gd.shape[d];
      m.grid_points *= gd.shape[d];
    }
  }
  memcpy(m.kspring, method->kspring, sizeof(m.kspring));
  memcpy(m.center, method->center, sizeof(m.center));
  m.height = method->height;
  for (int c = 0; c < MAXK; ++c) {
    m.sigma[c] = method->sigma[c];
    m.periods[c] = method->periods[c];
    m.r_lower[c] = method->r_lower[c];
    m.r_upper[c] = method->r_upper[c];
    m.r_kl[c] = method->r_kl[c];
    m.r_ku[c] = method->r_ku[c];
  }
  m.stride = method->stride;
  m.ngaussians = method->ngaussians;
  m.well_tempered = method->well_tempered;
  m.shared_hills = method->shared_hills;
  m.deltaT_kB = method->deltaT * method->kB;
  m.abf_N = method->abf_N;
  m.use_pinv = method->use_pinv;
  m.has_restraints = method->has_restraints;
  if (method->variant != 0 && !(method->kind == PSB_METHOD_ABF && (method->variant == 1 || method->variant == 2)))
    return bail(fail(PSB_ERR_VALUE, "unknown method variant %d", method->variant));
  m.hist_only = method->kind == PSB_METHOD_ABF && method->variant == 1;
  m.force_unit_factor = method->force_unit_factor == 0.0 ? 1.0 : method->force_unit_factor;
  if (method->kind == PSB_METHOD_METAD) {
    if (method->stride < 1) return bail(fail(PSB_ERR_VALUE, "Metadynamics stride must be >= 1"));
    if (method->ngaussians < 1) return bail(fail(PSB_ERR_VALUE, "Metadynamics ngaussians must be >= 1"));
    for (int c = 0; c < k; ++c)
      if (!(method->sigma[c] > 0.0)) return bail(fail(PSB_ERR_VALUE, "Metadynamics sigma must be positive"));
  }
  if (method->kind == PSB_METHOD_ABF && !layout->need_momenta && !m.hist_only)
    return bail(fail(PSB_ERR_VALUE, "ABF needs momenta (layout.need_momenta)"));
  M.abf_fast = (method->kind == PSB_METHOD_ABF && (M.all_point || m.hist_only)) ? 1 : 0;  // histogram only: no J p pass
  M.fm_mode = m.hist_only ? FM_HIST_ONLY : 0;

  // ---- unique atoms (CSR) and group overlaps ---------------------------------------------------
  std::vector<uint32_t> ov((size_t)MAXG * MAXG, 0u);  // group overlap counts
  auto OV = [&](int a, int b2) -> uint32_t& { return ov[(size_t)a * MAXG + b2]; };
  {
    std::vector<std::pair<uint32_t, uint32_t>> order(M.T);
    for (uint32_t t = 0; t < M.T; ++t) order[t] = {term_tag[t], t};
    std::sort(order.begin(), order.end());
    std::vector<uint32_t> uptr, uterm(M.T), utag;
    std::vector<uint8_t> ugrp(M.T);
    uptr.reserve(M.T + 1);
    for (uint32_t i = 0; i < M.T;) {
      uint32_t j = i;
      uptr.push_back(i);
      utag.push_back(order[i].first);
      while (j < M.T && order[j].first == order[i].first) ++j;
      for (uint32_t a = i; a < j; ++a) {
        uterm[a] = order[a].second;
        ugrp[a] = term_group[order[a].second];
        for (uint32_t b2 = i; b2 < j; ++b2)
          OV(term_group[order[a].second], term_group[order[b2].second]) += 1u;
      }
      i = j;
    }
    uptr.push_back(M.T);
    for (int g = 0; g < ng; ++g) {
      M.grp[g].self_coef = (double)OV(g, g) * M.grp[g].inv_n * M.grp[g].inv_n;
      for (int g2 = 0; g2 < ng; ++g2)
        if (g2 != g && OV(g, g2)) M.cross_overlap = 1;
    }
    M.Su = (uint32_t)uptr.size() - 1;
    M.all_unique = (M.Su == M.T) ? 1 : 0;
    psb_status st;
    if (M.cross_overlap) {
      uint32_t* p = nullptr;
      if ((st = dev_upload(h, &p, ov)) != PSB_OK) return bail(st);
      M.ov = p;
    }
    if (!M.all_unique) {
      uint32_t* p = nullptr;
      if ((st = dev_upload(h, &p, uptr)) != PSB_OK) return bail(st);
      M.uniq_ptr = p;
      if ((st = dev_upload(h, &p, uterm)) != PSB_OK) return bail(st);
      M.uniq_term = p;
      if ((st = dev_upload(h, &p, utag)) != PSB_OK) return bail(st);
      M.uniq_tag = p;
      uint8_t* q = nullptr;
      if ((st = dev_upload(h, &q, ugrp)) != PSB_OK) return bail(st);
      M.uniq_grp = q;
    }
    uint32_t* tt = nullptr;
    uint8_t* tg = nullptr;
    if ((st = dev_upload(h, &tt, term_tag)) != PSB_OK) return bail(st);
    if ((st = dev_upload(h, &tg, term_group)) != PSB_OK) return bail(st);
    M.term_tag = tt;
    M.term_group = tg;
    if ((st = dev_alloc(h, &M.term_slot, M.T)) != PSB_OK) return bail(st);
  }

  // ---- per-CV device data ------------------------------------------------------------------------
  for (int c = 0; c < ncvs; ++c) {
    const psb_cv_desc& d = cvs[c];
    CvDev& cv = M.cv[c];
    const uint32_t n = cv.t1 - cv.t0;
    psb_status st;
    if (d.kind == PSB_CV_RMSD) {
      if (!d.reference) return bail(fail(PSB_ERR_VALUE, "RMSD needs reference coordinates"));
      // orientation.py:121: Q = references - weighted_barycenter(references, weights)
      std::vector<double> w(n, 1.0 / (double)n), Q(3 * (size_t)n);
      if (d.weights) std::copy(d.weights, d.weights + n, w.begin());
      double cen[3] = {0, 0, 0};
      for (uint32_t i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) cen[a] += w[i] * d.reference[3 * (size_t)i + a];
      for (int a = 0; a < 3; ++a) cv.sumQ[a] = 0.0;
      for (uint32_t i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) {
          Q[3 * (size_t)i + a] = d.reference[3 * (size_t)i + a] - cen[a];
          cv.sumQ[a] += Q[3 * (size_t)i + a];
          cv.gq += Q[3 * (size_t)i + a] * Q[3 * (size_t)i + a];
        }
      double* p = nullptr;
      if ((st = dev_upload(h, &p, Q)) != PSB_OK) return bail(st);
      cv.ref = p;
      if (d.weights) {
        if ((st = dev_upload(h, &p, w)) != PSB_OK) return bail(st);
        cv.w = p;
      }
    }
    if (d.kind == PSB_CV_ERMSD) {  // orientation.py:129-175
      if (n % 3 != 0)
        return bail(fail(PSB_ERR_VALUE, "Please make sure the indices are in dimension 3 * n_nucleotides! Now it's %u.", n));
      const int nb = (int)n / 3;
      if (nb < 2 || nb > ERMSD_MAX_N)
        return bail(fail(PSB_ERR_VALUE, "CV %d: eRMSD takes between 2 and %d nucleotides (got %d)", c, ERMSD_MAX_N, nb));
      if (!d.reference || !d.weights || !(d.r0 > 0.0) || !(d.weights[0] > 0.0) || !(d.weights[1] > 0.0))
        return bail(fail(PSB_ERR_VALUE, "CV %d: eRMSD needs reference coordinates, a positive cutoff and rescaling lengths", c));
      cv.r0 = d.r0;
      cv.sumQ[0] = d.weights[0];  // a
      cv.sumQ[1] = d.weights[1];  // b
      if (d.axis != 0 && d.axis != 1) return bail(fail(PSB_ERR_VALUE, "CV %d: eRMSD variant must be 0 (all-atom) or 1 (coarse-grained)", c));
      std::vector<double> ref_atoms(d.reference, d.reference + 9 * (size_t)nb);
      if (d.axis == 1) {  // weights = {a, b, local_coordinates (4, 3, 3), sequence (n, as doubles)}
        const double* lc = d.weights + 2;
        const double* seq = d.weights + 38;
        for (int j = 0; j < nb; ++j)
          if (!(seq[j] >= 0.0 && seq[j] <= 3.0 && seq[j] == (double)(int)seq[j]))
            return bail(fail(PSB_ERR_VALUE, "CV %d: sequence entries must be 0 (A), 1 (U), 2 (G) or 3 (C)", c));
        ref_atoms = ermsd_reconstruct(d.reference, nb, lc, seq);
        std::vector<double> w(d.weights + 2, d.weights + 38 + nb);
        double* pw = nullptr;
        if ((st = dev_upload(h, &pw, w)) != PSB_OK) return bail(st);
        cv.w = pw;
      }
      const std::vector<double> gref = ermsd_reference_g(ref_atoms.data(), nb, d.r0, d.weights[0], d.weights[1]);
      double* p = nullptr;
      if ((st = dev_upload(h, &p, gref)) != PSB_OK) return bail(st);
      cv.ref = p;
      if ((st = dev_alloc(h, &cv.grad, 3 * (size_t)n)) != PSB_OK) return bail(st);
      if ((st = dev_alloc(h, &cv.xi_part, 1)) != PSB_OK) return bail(st);
      cv.n_part = 1;
      h->ermsd.push_back(c);
    }
    if (d.kind == PSB_CV_RING) {
      if (d.axis != PSB_RING_AMPLITUDE && d.axis != PSB_RING_PHASE)
        return bail(fail(PSB_ERR_VALUE, "CV %d: ring puckering mode must be amplitude (0) or phase (1)", c));
      if (n < 3) return bail(fail(PSB_ERR_VALUE, "CV %d: a ring needs at least 3 atoms (got %u)", c, n));
    }
    if (d.kind == PSB_CV_GYRATION) {
      if (d.axis < PSB_GYR_MOMENT_0 || d.axis > PSB_GYR_ANISOTROPY)
        return bail(fail(PSB_ERR_VALUE, "CV %d: unknown gyration-tensor function %d", c, d.axis));
      if (d.axis == PSB_GYR_ACYLINDRICITY && (d.nn < 0 || d.nn > 2 || d.mm < 0 || d.mm > 2))
        return bail(fail(PSB_ERR_VALUE, "CV %d: acylindricity eigenvalue indices must be 0, 1 or 2", c));
      cv.nn = d.nn;
      cv.mm = d.mm;
    }
    if (d.kind == PSB_CV_COORDINATION) {
      const int nn = d.nn == 0 ? 6 : d.nn, mm = d.mm == 0 ? 12 : d.mm;
      if (!(d.r0 > 0.0)) return bail(fail(PSB_ERR_VALUE, "CoordinationNumber r0 must be positive"));
      if (mm != 2 * nn || nn % 2 != 0 || !pair_launcher(1, nn / 2))
        return bail(fail(PSB_ERR_UNSUPPORTED, "CoordinationNumber supports even n in {2,4,6,8,12} with m = 2n (got n=%d, m=%d)", nn, mm));
      cv.r0 = d.r0;
      cv.nn = nn;
      cv.mm = mm;
      CoordWork cw;
      cw.cv = c;
      cw.H = nn / 2;
      cw.nA = (int)(M.grp[cv.g0].t1 - M.grp[cv.g0].t0);
      cw.nB = (int)(M.grp[cv.g0 + 1].t1 - M.grp[cv.g0 + 1].t0);
      cw.na_pad = (cw.nA + 31) / 32 * 32;
      cw.nb_pad = (cw.nB + 31) / 32 * 32;
      cw.R = cw.nA > 2048 ? 4 : 1;
      const int rows_per_cta = PAIR_BLOCK * cw.R;
      cw.nrt = (cw.nA + rows_per_cta - 1) / rows_per_cta;
      const int ctiles = cw.nb_pad / 32;
      cw.ncs = std::max(1, std::min(ctiles, (2 * h->num_sms) / cw.nrt));
      if ((st = dev_alloc(h, &cw.xa, 3 * (size_t)cw.na_pad)) != PSB_OK) return bail(st);
      if ((st = dev_alloc(h, &cw.xb, 3 * (size_t)cw.nb_pad)) != PSB_OK) return bail(st);
      if ((st = dev_alloc(h, &cv.grad, 3 * (size_t)n)) != PSB_OK) return bail(st);
      cv.n_part = cw.nrt * cw.ncs;
      if ((st = dev_alloc(h, &cv.xi_part, (size_t)cv.n_part)) != PSB_OK) return bail(st);
      h->coord.push_back(cw);
    }
  }

  // ---- state -------------------------------------------------------------------------------------
  {
    psb_status st;
    StateDev& s = M.st;
    if ((st = dev_alloc(h, &s.xi, MAXK)) != PSB_OK) return bail(st);
    if ((st = dev_alloc(h, &s.F, MAXK)) != PSB_OK) return bail(st);
    if ((st = dev_alloc(h, &s.energy, 1)) != PSB_OK) return bail(st);
    if ((st = dev_alloc(h, &s.ncalls, 1)) != PSB_OK) return bail(st);
    if (method->kind == PSB_METHOD_METAD) {
      h->hpad = ((size_t)method->ngaussians + 2) & ~(size_t)1;
      if ((st = dev_alloc(h, &s.heights, h->hpad)) != PSB_OK) return bail(st);
      if ((st = dev_alloc(h, &s.centers, h->hpad * k)) != PSB_OK) return bail(st);
      std::vector<double> sig(method->sigma, method->sigma + MAXK);
      if ((st = dev_upload(h, &s.sigmas, sig)) != PSB_OK) return bail(st);
      if ((st = dev_alloc(h, &s.idx, 1)) != PSB_OK) return bail(st);
      if (gridded) {
        if (method->well_tempered && (st = dev_alloc(h, &s.gp, (size_t)m.grid_points)) != PSB_OK) return bail(st);
        if ((st = dev_alloc(h, &s.gg, (size_t)m.grid_points * k)) != PSB_OK) return bail(st);
      }
      h->record_doubles = k + 2;
    }
    if (method->kind == PSB_METHOD_ABF) {
      if ((st = dev_alloc(h, &s.hist, (size_t)m.grid_points)) != PSB_OK) return bail(st);
      if ((st = dev_alloc(h, &s.Fsum, (size_t)m.grid_points * k)) != PSB_OK) return bail(st);
      if ((st = dev_alloc(h, &s.force, MAXK)) != PSB_OK) return bail(st);
      if ((st = dev_alloc(h, &s.Wp, MAXK)) != PSB_OK) return bail(st);
      if ((st = dev_alloc(h, &s.Wp_, MAXK)) != PSB_OK) return bail(st);
      if (method->variant == 2 && (st = dev_alloc(h, &s.histp, (size_t)m.grid_points)) != PSB_OK) return bail(st);
    }
    if ((st = dev_alloc(h, &M.fin, 1)) != PSB_OK) return bail(st);
    if ((st = dev_alloc(h, &M.bars, NBARS)) != PSB_OK) return bail(st);
    if ((st = dev_alloc(h, &M.epochs, 16)) != PSB_OK) return bail(st);
    if (layout->layout == PSB_LAYOUT_OPENMM_CUDA &&
        (st = dev_alloc(h, &M.inv_perm, (size_t)layout->natoms)) != PSB_OK)
      return bail(st);
    if (layout->dense_outputs) {
      if ((st = dev_alloc(h, &M.bias_dense, 3 * (size_t)layout->natoms)) != PSB_OK) return bail(st);
      if ((st = dev_alloc(h, &M.jac_dense, 3 * (size_t)layout->natoms * k)) != PSB_OK) return bail(st);
    }
  }

  // ---- launch shape --------------------------------------------------------------------------------
  // gridDim.x == 1: one CTA walks every group (no grid barrier at all).  gridDim.x > 1: every CTA
  // gathers ONE group (CTAs [cta_first[g], cta_first[g+1]) share group g evenly), at most
  // CT * BLOCK terms per CTA when the device has room, so slots and prefetched forces stay in
  // registers across the barrier.
  const bool metad_direct = (method->kind == PSB_METHOD_METAD && !gridded);
  if (metad_direct) {
    M.hill_chunk = (int)((HILL_STAGE_BYTES / ((k + 1) * 8)) & ~1);
    h->dyn_smem = 2 * (size_t)M.hill_chunk * (k + 1) * 8;
  }
  M.need_term_slot = (!M.all_unique || layout->dense_outputs || (method->kind == PSB_METHOD_ABF && !M.abf_fast) ||
                      method->shared_hills || M.any_coord)
                         ? 1
                         : 0;
  switch (method->kind) {
    case PSB_METHOD_UNBIASED: h->step_method = SM_UNBIASED; break;
    case PSB_METHOD_HARMONIC: h->step_method = SM_HARMONIC; break;
    case PSB_METHOD_ABF: h->step_method = SM_ABF; break;
    default: h->step_method = gridded ? SM_METAD_GRID : SM_METAD_DIRECT; break;
  }
  h->step_general = (!M.all_point || !M.all_unique || layout->dense_outputs) ? SC_GENERAL : SC_POINT;
  // the slot-ordered class also takes RadiusOfGyration groups next to point groups (not under ABF: its J J^T pass)
  bool slot_kinds = M.all_unique && !layout->dense_outputs && method->kind != PSB_METHOD_ABF;
  for (int c = 0; c < ncvs; ++c) slot_kinds = slot_kinds && (is_point(cvs[c].kind) || cvs[c].kind == PSB_CV_ROG);
  const int occ = pick_vtable(*layout)->occupancy(h->step_method, h->step_general, h->dyn_smem);
  if (occ < 1)
    return bail(fail(PSB_ERR_CUDA, "step kernel cannot be made resident (%s)", cudaGetErrorString(cudaGetLastError())));
  const long long max_blocks = (long long)std::min(occ, 2) * h->num_sms;
  long long want;
  if ((long long)M.T <= 2LL * BLOCK) {
    want = 1;
  } else if ((long long)M.T <= (long long)CT * BLOCK * h->num_sms) {
    want = std::min<long long>(h->num_sms, ((long long)M.T + BLOCK - 1) / BLOCK);  // >= 1 term per thread
  } else {
    want = max_blocks;
  }
  if (metad_direct) want = std::max(want, (long long)((method->ngaussians + M.hill_chunk - 1) / M.hill_chunk));
  if (method->kind == PSB_METHOD_METAD && gridded)
    want = std::max(want, m.grid_points / (BLOCK * 64));  // deposit steps are rare: 64 bins per thread is fine
  if (const char* env = getenv("PSB_GRID_BLOCKS")) want = atoll(env);
  want = std::max(1LL, std::min(want, max_blocks));
  if (want > 1) want = std::max<long long>(want, ng);
  if (want > max_blocks) return bail(fail(PSB_ERR_CUDA, "device too small for %d atom groups", ng));
  h->grid = (int)want;
  // slot-ordered gather/scatter (psb_step.cuh): worth it when most atoms are selected and the terms do
  // not fit the register cache; PSB_SLOT_MAJOR=1 forces it wherever it is valid (tests), =0 disables it
  {
    const bool eligible = h->grid > 1 && (!h->step_general || slot_kinds) && !method->shared_hills && ng <= 4 &&
                          layout->natoms < 0xFFFFFFF0LL;
    // HOOMD with the engine's slot -> tag array announced (psb_layout_desc.has_tags): the slot-ordered class needs no
    // inversion pass, its gather runs before the dependency wait, and with ONE CTA per SM the next step's CTAs are
    // co-resident, so that gather overlaps this step's exchange, finalizer and scatter (measured, ABF: 2e5 atoms 17.2 ->
    // 10.0 us, 1e6 atoms 28.0 -> 24.2 us against two CTAs per SM).  Worth it as soon as the register cache of one
    // CTA per SM is too small.
    const bool tags_hint = layout->has_tags && (layout->layout == PSB_LAYOUT_HOOMD || layout->layout == PSB_LAYOUT_LAMMPS) &&
                           !(getenv("PSB_NO_TAGS") && atoi(getenv("PSB_NO_TAGS")) != 0);
    bool wanted = (long long)M.T > (long long)CT * BLOCK * (tags_hint ? h->num_sms : h->grid) && 2LL * M.T >= layout->natoms;
    // OpenMM-CUDA: `ids` already IS slot -> atom, so the slot-ordered passes need no tag -> slot inverse at
    // all (the reference's per-step ids.argsort(), openmm.py:106-107, and our inverse-permutation launch
    // disappear); worth it as soon as a fair share of the system is selected
    const bool omm = layout->layout == PSB_LAYOUT_OPENMM_CUDA;
    if (omm && 4LL * M.T >= layout->natoms) wanted = true;
    if (const char* env = getenv("PSB_SLOT_MAJOR")) wanted = atoi(env) != 0;
    M.slot_major = (eligible && wanted) ? 1 : 0;
    if (M.slot_major) {
      const int occ_slot = pick_vtable(*layout)->occupancy(h->step_method, SC_SLOT, h->dyn_smem);
      if (occ_slot < 1) {
        M.slot_major = 0;
      } else if (!getenv("PSB_GRID_BLOCKS")) {
        h->grid = std::min(occ_slot, SLOT_CTAS_PER_SM) * h->num_sms;  // streaming: fill the machine
        // OpenMM layout: one CTA per SM leaves room for the NEXT step's CTAs, whose gather runs before their
        // dependency wait (psb_step.cuh), i.e. under this step's finalizer and scatter (1e6 atoms, ABF: 44.0 -> 39.3 us)
        if (omm) {
          // from ~2e5 atoms both launch shapes are kept, like the HOOMD handles below: single launches want the full
          // machine (plain launches, ABF: 3e5 atoms 22.2 -> 18.8 us, 1e6 atoms 67.5 -> 46.2 us with two CTAs per SM)
          if (layout->natoms > 5LL * BLOCK * h->num_sms && h->grid >= 2 * h->num_sms && !getenv("PSB_DEBUG_TIMING") &&
              !(getenv("PSB_ONE_GRID") && atoi(getenv("PSB_ONE_GRID")) != 0)) {
            h->grid_plain = h->grid;
            h->grid_pipe = h->num_sms;
          } else {
            h->grid = h->num_sms;
          }
        }
        // (not with RadiusOfGyration groups: their scatter streams the positions a second time, the part behind the
        // dependency wait dominates and wants all 16 warps per SM -- measured 29.2 us at two CTAs per SM, 36.6 at one)
        bool rog_groups = false;
        for (int c = 0; c < ncvs; ++c) rog_groups = rog_groups || cvs[c].kind == PSB_CV_ROG;
        if (tags_hint && !rog_groups) {
          if (h->grid >= 2 * h->num_sms && !getenv("PSB_DEBUG_TIMING") && !(getenv("PSB_ONE_GRID") && atoi(getenv("PSB_ONE_GRID")) != 0)) {
            h->grid_plain = h->grid;       // stepping one launch at a time (engine hook): two CTAs per SM (1e6 atoms, ABF: 30.8 us
            h->grid_pipe = h->num_sms;     // against 35.2 with one); replayed back to back: one (21.5 us against 27.9)
          } else {
            h->grid = h->num_sms;
          }
        }
      } else if ((long long)occ_slot * h->num_sms < h->grid) {
        M.slot_major = 0;  // the slot-ordered instantiation would not be co-resident at this grid size
      }
    }
    if (M.slot_major) {
      h->step_general = SC_SLOT;
      // TMA-staged streaming (SC_STREAM, psb_stream.cuh): HOOMD layout, two CTAs per SM, per-warp rings over all the
      // shared memory the instantiation leaves free.  Opt-in (PSB_STREAM=1): measured on B200 it does NOT beat the
      // register-staged slot-ordered class (the passes are issue-bound at 16 warps x fp64, not short
      // of bytes in flight).  Metadynamics by direct summation keeps the slot-ordered class in any case (its hill
      // staging lives in dynamic shared memory too).
      if (layout->layout == PSB_LAYOUT_HOOMD && !metad_direct && getenv("PSB_STREAM") && atoi(getenv("PSB_STREAM")) != 0) {
        const size_t stat = pick_vtable(*layout)->static_smem(h->step_method, SC_STREAM);
        const size_t limit = 227 * 1024;  // per-CTA shared memory on sm_100 (the SM keeps 1 KB per CTA for itself)
        bool any_rog = false;
        for (int c = 0; c < ncvs; ++c) any_rog = any_rog || cvs[c].kind == PSB_CV_ROG;
        const uint32_t P = 4u * (uint32_t)layout->real_bytes, I = layout->unwrap ? 12u : 0u;
        StreamGeom g{};
        g.ch = 32 * STREAM_RPL;  // slots per warp chunk: fine-grained stages, so the whole run of a warp fits its ring
        g.early = STREAM_MAX_STAGES;
        if (const char* env = getenv("PSB_STREAM_EARLY")) g.early = atoi(env);
        if (const char* env = getenv("PSB_STREAM_FLAGS")) g.flags = atoi(env);
        const uint32_t ch = (uint32_t)g.ch;
        uint32_t gbytes = ch * P;
        if (I) { g.g_img = gbytes; gbytes += ch * I; }
        if (layout->need_momenta) { g.g_vel = gbytes; gbytes += ch * P; }
        uint32_t sbytes = ch * P;
        if (any_rog) {
          g.s_pos = sbytes; sbytes += ch * P;
          if (I) { g.s_img = sbytes; sbytes += ch * I; }
        }
        g.stage_bytes = (std::max(gbytes, sbytes) + 127u) & ~127u;
        // two CTAs per SM (16 warps: the passes over the staged rows are issue-bound), each with half of what the
        // SM has left after the static allocations (+ 1 KB per CTA the hardware keeps)
        const size_t per_cta = stat == (size_t)-1 ? 0 : (limit + 1024) / 2 - 1024 - std::min<size_t>(stat, limit / 2);
        const size_t per_warp = per_cta / NWARPS;
        if (per_warp >= 2 * (size_t)g.stage_bytes) {
          g.stages = (int)std::min<size_t>(STREAM_MAX_STAGES, per_warp / g.stage_bytes);
          if (const char* env = getenv("PSB_STREAM_STAGES")) g.stages = std::max(2, std::min(g.stages, atoi(env)));
          const size_t dyn = (size_t)NWARPS * g.stages * g.stage_bytes;
          const int occ_s = pick_vtable(*layout)->occupancy(h->step_method, SC_STREAM, dyn);
          if (occ_s >= 1) {
            h->step_general = SC_STREAM;
            h->dyn_smem = dyn;
            M.sg = g;
            if (!getenv("PSB_GRID_BLOCKS")) h->grid = std::min(occ_s, 2) * h->num_sms;
            else h->grid = std::min(h->grid, std::min(occ_s, 2) * h->num_sms);
          }
        }
      }
      psb_status st = PSB_OK;
      if (omm) {
        std::vector<uint8_t> tag_group((size_t)layout->natoms, 0);
        for (uint32_t t = 0; t < M.T; ++t) tag_group[term_tag[t]] = (uint8_t)(term_group[t] + 1);
        uint8_t* p = nullptr;
        st = dev_upload(h, &p, tag_group);
        M.tag_group = p;
      } else {
        st = dev_alloc(h, &M.slot_map, (size_t)layout->natoms);
        if (st == PSB_OK && (layout->layout == PSB_LAYOUT_HOOMD || layout->layout == PSB_LAYOUT_LAMMPS) && h->step_general == SC_SLOT &&
            !(getenv("PSB_NO_TAGS") && atoi(getenv("PSB_NO_TAGS")) != 0) &&
            (long long)pick_vtable(*layout)->occupancy(h->step_method, SC_TAGS, h->dyn_smem) * h->num_sms >= h->grid) {
          // snapshots that carry the engine's slot -> tag array take the SC_TAGS instantiation (enqueue)
          std::vector<uint8_t> tag_group((size_t)layout->natoms, 0);
          for (uint32_t t = 0; t < M.T; ++t) tag_group[term_tag[t]] = (uint8_t)(term_group[t] + 1);
          uint8_t* p = nullptr;
          st = dev_upload(h, &p, tag_group);
          M.tag_group = p;
          h->tags_class = true;
        }
      }
      if (st != PSB_OK) return bail(st);
    }
  }
  if (h->grid > 1 && M.slot_major) {
    for (int g = 0; g < MAXG; ++g) {
      M.cta_lo[g] = 0;
      M.cta_hi[g] = g < ng ? h->grid - 1 : -1;
    }
    // The slot-ordered classes (hundreds of CTAs) exchange their pass-A sums through flagged words as well, in two
    // levels (reduce_rows_ll2, psb_step.cuh): row owners add the partials, everybody reads the totals.  Same methods
    // as below; not the TMA-staged variant.
    const bool ll_method = method->kind == PSB_METHOD_UNBIASED || method->kind == PSB_METHOD_HARMONIC ||
                           method->kind == PSB_METHOD_ABF;
    if (h->step_general == SC_SLOT && ll_method && !getenv("PSB_NO_LL") &&
        !(getenv("PSB_NO_LL2") && atoi(getenv("PSB_NO_LL2")) != 0)) {
      M.ll_reduce = 1;
      psb_status st = dev_alloc(h, &M.ll, (size_t)MAXG * NACC * h->grid);
      if (st == PSB_OK) st = dev_alloc(h, &M.ll_tot, (size_t)MAXG * NACC);
      if (st == PSB_OK) st = dev_alloc(h, &h->d_cta_seq, (size_t)h->grid);
      if (st != PSB_OK) return bail(st);
    }
  } else if (h->grid > 1) {
    // smallest per-CTA term count whose group-aligned split fits the grid
    auto ctas_needed = [&](long long per) {
      long long n = 0;
      for (int g = 0; g < ng; ++g) n += ((long long)(M.grp[g].t1 - M.grp[g].t0) + per - 1) / per;
      return n;
    };
    long long per = std::max(1LL, ((long long)M.T + h->grid - 1) / h->grid);
    while (ctas_needed(per) > h->grid) per += std::max(1LL, per / 64);
    int first = 0;
    for (int g = 0; g < MAXG; ++g) {
      const int n = g < ng ? (int)(((long long)(M.grp[g].t1 - M.grp[g].t0) + per - 1) / per) : 0;
      M.cta_lo[g] = first;
      M.cta_hi[g] = first + n - 1;
      first += n;
    }
    // Pass-A partials through flagged words instead of a counter barrier (reduce_rows_ll, psb_step.cuh):
    // point-CV class, methods whose finalizer needs nothing else from other CTAs, and every CTA's share
    // small enough for the register cache, so that all of them publish from the staged sums.
    const bool ll_method = method->kind == PSB_METHOD_UNBIASED || method->kind == PSB_METHOD_HARMONIC ||
                           method->kind == PSB_METHOD_ABF;
    if (h->step_general == SC_POINT && ll_method && per <= (long long)CT * BLOCK && !getenv("PSB_NO_LL")) {
      M.ll_reduce = 1;
      psb_status st = dev_alloc(h, &M.ll, (size_t)MAXG * NACC * h->grid);
      if (st == PSB_OK) st = dev_alloc(h, &h->d_cta_seq, (size_t)h->grid);
      if (st != PSB_OK) return bail(st);
    }
  }
  {
    const bool momenta_sums = layout->need_momenta && M.abf_fast;
    for (int pass = 0; pass < 2; ++pass) {
      M.row_first[pass][0] = 0;
      for (int g = 0; g < MAXG; ++g)
        M.row_first[pass][g + 1] =
            M.row_first[pass][g] + (g < ng ? nacc_for_kind(M.cv[M.grp[g].cv].kind, pass, momenta_sums) : 0);
    }
  }
  // ---- ABF: constant J J^T (see post_cv) ---------------------------------------------------------------
  if (method->kind == PSB_METHOD_ABF && M.abf_fast && !method->use_pinv) {
    bool constant = true;
    for (int c = 0; c < ncvs; ++c) {
      const int kd = M.cv[c].kind;
      if (kd != PSB_CV_COMPONENT && kd != PSB_CV_DISTANCE && kd != PSB_CV_DISPLACEMENT) constant = false;
    }
    for (int g = 0; g < ng && constant; ++g)
      for (int g2 = 0; g2 < ng; ++g2)
        if (M.grp[g].cv != M.grp[g2].cv && OV(g, g2)) constant = false;  // atoms shared between CVs
    if (constant) {
      double A[16] = {0};
      for (int c = 0; c < ncvs; ++c) {
        const CvDev& cv = M.cv[c];
        double d = M.grp[cv.g0].self_coef;
        if (cv.ngroups == 2) {  // gradients of the two centroids are opposite unit vectors / +-identity
          const int g1 = cv.g0, g2 = cv.g0 + 1;
          d += M.grp[g2].self_coef - 2.0 * (double)OV(g1, g2) * M.grp[g1].inv_n * M.grp[g2].inv_n;
        }
        for (int j = 0; j < cv.ncomp; ++j) A[(cv.comp0 + j) * 4 + cv.comp0 + j] = d;
      }
      bool ok = true;  // diagonal by construction (no cross-CV overlap): invert entry by entry
      for (int a = 0; a < k; ++a) ok = ok && A[a * 4 + a] > 1e-300;
      if (ok) {
        M.jjt_const = 1;
        for (int a = 0; a < k; ++a) M.jjt_inv[a * 4 + a] = 1.0 / A[a * 4 + a];
      }
    }
    if (getenv("PSB_NO_CONST_JJT")) M.jjt_const = 0;
  }
  // host-resident engines (psb_step_host): when few atoms are selected only their rows travel
  {
    const bool valid = !M.slot_major && !layout->dense_outputs && layout->layout != PSB_LAYOUT_OPENMM_CUDA;
    bool wanted = 8LL * (long long)h->sel_tags.size() <= layout->natoms;
    if (const char* env = getenv("PSB_HOST_COMPACT")) wanted = atoi(env) != 0;
    h->compact_ok = valid && wanted;
  }
  h->cooperative = h->grid > 1;
  if (h->cooperative) {
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device);
    if (!coop) return bail(fail(PSB_ERR_CUDA, "device does not support cooperative launches"));
  }
  {
    psb_status st = dev_alloc(h, &M.partials, (size_t)(MAXG + 1) * NACC * h->grid);
    if (st != PSB_OK) return bail(st);
    if (getenv("PSB_DEBUG_TIMING") && (st = dev_alloc(h, &h->dbg, (size_t)2 * 16 * h->grid)) != PSB_OK) return bail(st);
  }

  // ---- algorithmic bytes per step ----------------------------------------------------
  {
    const long long rb = layout->real_bytes;
    long long pos = 0, frc = 0, vel = 0, img = 0, idb = 8;  // index + tag->slot map
    switch (layout->layout) {
      case PSB_LAYOUT_HOOMD: pos = 4 * rb; frc = 8 * rb; vel = 4 * rb; img = layout->unwrap ? 12 : 0; break;
      case PSB_LAYOUT_OPENMM_CUDA: pos = 4 * rb; frc = 48; vel = 4 * rb; break;
      case PSB_LAYOUT_LAMMPS: pos = 24; frc = 48; vel = 24 + 4 + 8; img = layout->unwrap ? 4 : 0; break;
      default: pos = 3 * rb; frc = 6 * rb; vel = 4 * rb; break;
    }
    long long per_term = idb + pos + img + (layout->need_momenta ? vel : 0);
    long long bytes = (long long)M.T * per_term + (long long)M.Su * frc;
    for (int c = 0; c < ncvs; ++c)
      if (M.cv[c].kind == PSB_CV_RMSD) bytes += (long long)(M.cv[c].t1 - M.cv[c].t0) * (24 + (M.cv[c].w ? 8 : 0));
    if (layout->layout == PSB_LAYOUT_OPENMM_CUDA) bytes += layout->natoms * 8;
    if (metad_direct) bytes += method->ngaussians * (k + 1) * 8;
    if (method->kind == PSB_METHOD_METAD && gridded) bytes += (k + 1) * 8;
    if (method->kind == PSB_METHOD_ABF) bytes += 2 * (4 + 8 * k);
    h->algo_bytes = bytes;
  }
  {
    psb_status st = dev_alloc(h, &h->d_model, 1, false);
    if (st != PSB_OK) return bail(st);
    cudaError_t e = cudaMemcpy(h->d_model, &h->model, sizeof(Model), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return bail(fail(PSB_ERR_CUDA, "CUDA error %s uploading the model", cudaGetErrorString(e)));
  }
  if (getenv("PSB_NO_PDL")) h->pdl = 0;
  *out = h;
  return PSB_OK;
}

psb_status psb_step(psb_handle* h, const psb_snapshot* snap, int64_t timestep, void* cuda_stream) {
  (void)timestep;
  if (!h) return fail(PSB_ERR_VALUE, "NULL handle");
  psb_status st = check_snapshot(h, snap, true);
  if (st != PSB_OK) return st;
  if ((st = replan_grid(h, false, (cudaStream_t)cuda_stream)) != PSB_OK) return st;  // one launch at a time
  st = enqueue(h, snap, MODE_STEP, (cudaStream_t)cuda_stream, nullptr, 0, nullptr);
  if (st == PSB_OK) ++h->host_ncalls;
  return st;
}

psb_status psb_evaluate_cvs(psb_handle* h, const psb_snapshot* snap, void* cuda_stream) {
  if (!h) return fail(PSB_ERR_VALUE, "NULL handle");
  psb_status st = check_snapshot(h, snap, false);
  if (st != PSB_OK) return st;
  return enqueue(h, snap, MODE_EVAL, (cudaStream_t)cuda_stream, nullptr, 0, nullptr);
}

psb_status psb_run_cycle(psb_handle* h, const psb_snapshot* snaps, int32_t nsnaps, int64_t nsteps,
                         void* cuda_stream) {
  if (!h || !snaps || nsnaps < 1) return fail(PSB_ERR_VALUE, "invalid arguments");
  for (int i = 0; i < nsnaps; ++i) {
    psb_status st = check_snapshot(h, snaps + i, true);
    if (st != PSB_OK) return st;
  }
  cudaStream_t stream = (cudaStream_t)cuda_stream;
  int64_t done = 0;
  h->graph_replays = 0;
  // Launch-bound loop: one pass over the snapshot ring is captured once as a CUDA graph and
  // replayed (the kernels keep no host-side counters: barrier epochs and method state live on the
  // device, so replays are exact).  PSB_NO_GRAPH=1 or debug timing fall back to plain launches.
  static const bool no_graph = getenv("PSB_NO_GRAPH") != nullptr;
  // PSB_DEBUG_TIMING=graph keeps the graph: captured nodes alternate between the two stamp halves,
  // which stays consistent across replays when the ring has an even number of snapshots
  static const bool dbg_graph = getenv("PSB_DEBUG_TIMING") && std::string(getenv("PSB_DEBUG_TIMING")) == "graph";
  const bool replay = !no_graph && h->use_graph && (!h->dbg || (dbg_graph && nsnaps % 2 == 0)) && stream != nullptr && nsteps >= (int64_t)nsnaps;
  {
    psb_status st = replan_grid(h, replay, stream);
    if (st != PSB_OK) return st;
  }
  if (replay) {
    const bool same = h->cycle_exec && h->cycle_stream == stream && (int)h->cycle_key.size() == nsnaps &&
                      memcmp(h->cycle_key.data(), snaps, sizeof(psb_snapshot) * nsnaps) == 0;
    if (!same) {
      if (h->cycle_exec) cudaGraphExecDestroy(h->cycle_exec);
      h->cycle_exec = nullptr;
      cudaGraph_t graph = nullptr;
      const long long l0 = h->launches;
      if (cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        psb_status st = PSB_OK;
        for (int i = 0; i < nsnaps && st == PSB_OK; ++i) st = enqueue(h, snaps + i, MODE_STEP, stream, nullptr, 0, nullptr);
        cudaError_t e = cudaStreamEndCapture(stream, &graph);
        if (st == PSB_OK && e == cudaSuccess && graph &&
            cudaGraphInstantiate(&h->cycle_exec, graph, 0) == cudaSuccess) {
          h->cycle_key.assign(snaps, snaps + nsnaps);
          h->cycle_stream = stream;
          h->cycle_launches = h->launches - l0;
        } else {
          h->cycle_exec = nullptr;
        }
        if (graph) cudaGraphDestroy(graph);
      }
      cudaGetLastError();     // a failed capture is not an error: plain launches below
      h->launches = l0;       // captured launches have not run yet
    }
    if (h->cycle_exec) {
      const int64_t reps = nsteps / nsnaps;
      for (int64_t r = 0; r < reps; ++r) CUDA_TRY(cudaGraphLaunch(h->cycle_exec, stream));
      done = reps * nsnaps;
      h->host_ncalls += done;
      h->launches += reps * h->cycle_launches;
      h->graph_replays = 1;
    }
  }
  for (int64_t i = done; i < nsteps; ++i) {
    psb_status st = enqueue(h, snaps + (i % nsnaps), MODE_STEP, stream, nullptr, 0, nullptr);
    if (st != PSB_OK) return st;
    ++h->host_ncalls;
  }
  return PSB_OK;
}

// ---- many independent simulations per launch (umbrella windows) ---------------------------------------------
}  // extern "C"

struct psb_batch {
  std::vector<psb_handle*> hs;
  const StepVTable* vt = nullptr;
  int device = 0;
  int total_ctas = 0;
  int32_t* d_cta_item = nullptr;  // (total_ctas) CTA -> simulation
  // Descriptor arrays live in device memory (as a kernel parameter they would cost launch latency): one per distinct
  // set of snapshot pointers seen, found again by content -- engines re-present the same buffers step after step
  // (OpenMM: persistent; HOOMD: two alternating sets), a benchmark ring its nsnaps sets.
  struct Slot {
    std::vector<BatchItem> host;
    BatchItem* dev = nullptr;
  };
  std::vector<Slot> slots;
  size_t next_victim = 0;
  long long launches = 0;
  cudaGraphExec_t cycle_exec = nullptr;
  std::vector<psb_snapshot> cycle_key;
  cudaStream_t cycle_stream = nullptr;
  int graph_replays = 0;
};

namespace {

constexpr size_t BATCH_MAX_SLOTS = 256;

// descriptor array for `snaps` (n = hs.size() snapshots), uploaded on `stream` if it is not resident yet
psb_status batch_items(psb_batch* B, const psb_snapshot* snaps, cudaStream_t stream, const BatchItem** out) {
  const size_t n = B->hs.size();
  std::vector<BatchItem> items(n);
  int cta0 = 0;
  for (size_t i = 0; i < n; ++i) {
    psb_handle* h = B->hs[i];
    psb_status st = check_snapshot(h, snaps + i, true);
    if (st != PSB_OK) return st;
    memset(&items[i], 0, sizeof(BatchItem));
    items[i].model = h->d_model;
    items[i].snap = make_snap(h, snaps + i, MODE_STEP);
    items[i].snap.dbg = nullptr;
    items[i].cta0 = cta0;
    items[i].ncta = h->grid;
    cta0 += h->grid;
  }
  for (psb_batch::Slot& sl : B->slots)
    if (memcmp(sl.host.data(), items.data(), n * sizeof(BatchItem)) == 0) {
      *out = sl.dev;
      return PSB_OK;
    }
  psb_batch::Slot* sl = nullptr;
  if (B->slots.size() < BATCH_MAX_SLOTS) {
    B->slots.emplace_back();
    sl = &B->slots.back();
    CUDA_TRY(cudaMalloc(&sl->dev, n * sizeof(BatchItem)));
  } else {  // reuse the oldest array: the copy below is ordered after the launches that read it (same stream)
    sl = &B->slots[B->next_victim];
    B->next_victim = (B->next_victim + 1) % BATCH_MAX_SLOTS;
  }
  sl->host = items;
  CUDA_TRY(cudaMemcpyAsync(sl->dev, sl->host.data(), n * sizeof(BatchItem), cudaMemcpyHostToDevice, stream));
  *out = sl->dev;
  return PSB_OK;
}

psb_status batch_enqueue(psb_batch* B, const psb_snapshot* snaps, cudaStream_t stream) {
  const BatchItem* items = nullptr;
  psb_status st = batch_items(B, snaps, stream, &items);
  if (st != PSB_OK) return st;
  psb_handle* h0 = B->hs[0];
  CUDA_TRY(B->vt->launch_batch(h0->step_method, h0->step_general, items, B->d_cta_item, B->total_ctas, h0->dyn_smem,
                               stream, h0->pdl));
  ++B->launches;
  return PSB_OK;
}

}  // namespace

extern "C" {

psb_status psb_batch_create(psb_handle* const* handles, int32_t n, psb_batch** out) {
  if (!handles || !out || n < 1) return fail(PSB_ERR_VALUE, "invalid arguments");
  *out = nullptr;
  psb_handle* h0 = handles[0];
  if (!h0) return fail(PSB_ERR_VALUE, "NULL handle");
  int total = 0;
  for (int i = 0; i < n; ++i) {
    psb_handle* h = handles[i];
    if (!h) return fail(PSB_ERR_VALUE, "NULL handle");
    for (int j = 0; j < i; ++j)
      if (handles[j] == h) return fail(PSB_ERR_VALUE, "simulation %d appears twice in the batch", i);
    // one launch runs one instantiation of the step kernel: same engine layout, dtype, method kind and kernel class
    if (h->layout.layout != h0->layout.layout || h->layout.real_bytes != h0->layout.real_bytes ||
        h->step_method != h0->step_method || h->step_general != h0->step_general || h->dyn_smem != h0->dyn_smem ||
        h->device != h0->device)
      return fail(PSB_ERR_VALUE, "simulation %d differs from simulation 0 in layout, dtype, method or kernel class: "
                                 "a batch shares one kernel instantiation", i);
    if (h->step_general != SC_POINT && h->step_general != SC_GENERAL)
      return fail(PSB_ERR_UNSUPPORTED, "simulation %d uses the slot-ordered class (most atoms of a very large system "
                                       "selected): it fills the device by itself, step it with psb_step", i);
    if (!h->coord.empty() || !h->ermsd.empty() || h->model.bias_dense ||
        (h->layout.layout == PSB_LAYOUT_OPENMM_CUDA && !h->model.slot_major))
      return fail(PSB_ERR_UNSUPPORTED, "simulation %d needs helper launches before its step kernel (pairwise / eRMSD "
                                       "CVs, OpenMM inverse permutation, dense outputs): step it with psb_step", i);
    total += h->grid;
  }
  const StepVTable* vt = pick_vtable(h0->layout);
  const int occ = vt->batch_occupancy(h0->step_method, h0->step_general, h0->dyn_smem);
  if ((long long)occ * h0->num_sms < total)
    return fail(PSB_ERR_VALUE, "the batch needs %d co-resident CTAs, the device holds %d: split it", total, occ * h0->num_sms);
  psb_batch* B = new psb_batch();
  B->hs.assign(handles, handles + n);
  B->vt = vt;
  B->device = h0->device;
  B->total_ctas = total;
  std::vector<int32_t> cta_item((size_t)total);
  int c = 0;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < handles[i]->grid; ++j) cta_item[c++] = i;
  cudaError_t e = cudaMalloc(&B->d_cta_item, sizeof(int32_t) * (size_t)total);
  if (e == cudaSuccess) e = cudaMemcpy(B->d_cta_item, cta_item.data(), sizeof(int32_t) * (size_t)total, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    delete B;
    return fail(PSB_ERR_CUDA, "CUDA error %s in psb_batch_create", cudaGetErrorString(e));
  }
  *out = B;
  return PSB_OK;
}

void psb_batch_destroy(psb_batch* B) {
  if (!B) return;
  cudaSetDevice(B->device);
  if (B->cycle_exec) cudaGraphExecDestroy(B->cycle_exec);
  for (psb_batch::Slot& sl : B->slots) cudaFree(sl.dev);
  cudaFree(B->d_cta_item);
  delete B;
}

psb_status psb_batch_step(psb_batch* B, const psb_snapshot* snaps, int64_t timestep, void* cuda_stream) {
  (void)timestep;
  if (!B || !snaps) return fail(PSB_ERR_VALUE, "NULL argument");
  psb_status st = batch_enqueue(B, snaps, (cudaStream_t)cuda_stream);
  if (st != PSB_OK) return st;
  for (psb_handle* h : B->hs) ++h->host_ncalls;
  return PSB_OK;
}

psb_status psb_batch_run_cycle(psb_batch* B, const psb_snapshot* snaps, int32_t nsnaps, int64_t nsteps, void* cuda_stream) {
  if (!B || !snaps || nsnaps < 1 || nsteps < 0) return fail(PSB_ERR_VALUE, "invalid arguments");
  cudaStream_t stream = (cudaStream_t)cuda_stream;
  const size_t n = B->hs.size();
  int64_t done = 0;
  B->graph_replays = 0;
  static const bool no_graph = getenv("PSB_NO_GRAPH") != nullptr;
  if (!no_graph && stream != nullptr && nsteps >= (int64_t)nsnaps) {
    const bool same = B->cycle_exec && B->cycle_stream == stream && B->cycle_key.size() == n * (size_t)nsnaps &&
                      memcmp(B->cycle_key.data(), snaps, sizeof(psb_snapshot) * n * nsnaps) == 0;
    if (!same) {
      if (B->cycle_exec) cudaGraphExecDestroy(B->cycle_exec);
      B->cycle_exec = nullptr;
      // descriptor uploads must not be part of the captured work: make every ring position resident first
      for (int r = 0; r < nsnaps; ++r) {
        const BatchItem* unused = nullptr;
        psb_status st = batch_items(B, snaps + (size_t)r * n, stream, &unused);
        if (st != PSB_OK) return st;
      }
      CUDA_TRY(cudaStreamSynchronize(stream));
      cudaGraph_t graph = nullptr;
      const long long l0 = B->launches;
      if (cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        psb_status st = PSB_OK;
        for (int r = 0; r < nsnaps && st == PSB_OK; ++r) st = batch_enqueue(B, snaps + (size_t)r * n, stream);
        cudaError_t e = cudaStreamEndCapture(stream, &graph);
        if (st == PSB_OK && e == cudaSuccess && graph && cudaGraphInstantiate(&B->cycle_exec, graph, 0) == cudaSuccess) {
          B->cycle_key.assign(snaps, snaps + n * (size_t)nsnaps);
          B->cycle_stream = stream;
        } else {
          B->cycle_exec = nullptr;
        }
        if (graph) cudaGraphDestroy(graph);
      }
      cudaGetLastError();
      B->launches = l0;
    }
    if (B->cycle_exec) {
      const int64_t reps = nsteps / nsnaps;
      for (int64_t r = 0; r < reps; ++r) CUDA_TRY(cudaGraphLaunch(B->cycle_exec, stream));
      done = reps * nsnaps;
      B->launches += done;
      B->graph_replays = 1;
    }
  }
  for (int64_t i = done; i < nsteps; ++i) {
    psb_status st = batch_enqueue(B, snaps + (size_t)(i % nsnaps) * n, stream);
    if (st != PSB_OK) return st;
  }
  for (psb_handle* h : B->hs) h->host_ncalls += nsteps;
  return PSB_OK;
}

int64_t psb_batch_launch_count(psb_batch* B) { return B ? B->launches : 0; }

psb_status psb_walker_begin(psb_handle* h, const psb_snapshot* snap, int64_t timestep,
                            double* record_dev, void* cuda_stream) {
  (void)timestep;
  if (!h || !record_dev) return fail(PSB_ERR_VALUE, "NULL argument");
  if (h->method.kind != PSB_METHOD_METAD) return fail(PSB_ERR_VALUE, "walkers need Metadynamics");
  psb_status st = check_snapshot(h, snap, true);
  if (st != PSB_OK) return st;
  return enqueue(h, snap, MODE_WALKER_BEGIN, (cudaStream_t)cuda_stream, nullptr, 0, record_dev);
}

psb_status psb_walker_finish(psb_handle* h, const psb_snapshot* snap, const double* records_dev,
                             int32_t nwalkers, void* cuda_stream) {
  if (!h || !records_dev || nwalkers < 1) return fail(PSB_ERR_VALUE, "invalid arguments");
  if (h->method.kind != PSB_METHOD_METAD) return fail(PSB_ERR_VALUE, "walkers need Metadynamics");
  psb_status st = check_snapshot(h, snap, true);
  if (st != PSB_OK) return st;
  st = enqueue(h, snap, MODE_WALKER_FINISH, (cudaStream_t)cuda_stream, records_dev, nwalkers, nullptr);
  if (st == PSB_OK) ++h->host_ncalls;
  return st;
}

int32_t psb_walker_record_doubles(psb_handle* h) { return h ? h->record_doubles : 0; }
int32_t psb_next_step_deposits(psb_handle* h) { return (h && host_next_deposits(h)) ? 1 : 0; }

psb_status psb_state_info(psb_handle* h, const char* field, int64_t* nbytes, int32_t* dtype_code,
                          void** device_ptr) {
  if (!h || !field) return fail(PSB_ERR_VALUE, "NULL argument");
  const Model& M = h->model;
  const StateDev& s = M.st;
  const long long k = M.k, G = M.m.grid_points, H = M.m.ngaussians;
  const std::string f(field);
  void* p = nullptr;
  long long n = 0;
  int code = 0;  // 0 f64, 1 i64, 2 u32
  if (f == "xi") { p = s.xi; n = 8 * k; }
  else if (f == "cv_force") { p = s.F; n = 8 * k; }
  else if (f == "energy") { p = s.energy; n = 8; }
  else if (f == "ncalls") { p = s.ncalls; n = 8; code = 1; }
  else if (f == "heights") { p = s.heights; n = 8 * H; }
  else if (f == "centers") { p = s.centers; n = 8 * H * k; }
  else if (f == "sigmas") { p = s.sigmas; n = 8 * k; }
  else if (f == "grid_potential") { p = s.gp; n = 8 * G; }
  else if (f == "grid_gradient") { p = s.gg; n = 8 * G * k; }
  else if (f == "idx") { p = s.idx; n = 8; code = 1; }
  else if (f == "hist") { p = s.hist; n = 4 * G; code = 2; }
  else if (f == "histp") { p = s.histp; n = 4 * G; code = 2; }
  else if (f == "Fsum") { p = s.Fsum; n = 8 * G * k; }
  else if (f == "force") { p = s.force; n = 8 * k; }
  else if (f == "Wp") { p = s.Wp; n = 8 * k; }
  else if (f == "Wp_") { p = s.Wp_; n = 8 * k; }
  else if (f == "bias") { p = M.bias_dense; n = 8 * 3 * M.natoms; }
  else if (f == "jacobian") { p = M.jac_dense; n = 8 * 3 * M.natoms * k; }
  else if (f == "debug_timing") { p = h->dbg; n = 2 * 8 * 16 * (long long)h->grid; code = 1; }
  else return fail(PSB_ERR_KEY, "unknown state field '%s'", field);
  if (!p) return fail(PSB_ERR_KEY, "state field '%s' does not exist for this method", field);
  if (nbytes) *nbytes = n;
  if (dtype_code) *dtype_code = code;
  if (device_ptr) *device_ptr = p;
  return PSB_OK;
}

psb_status psb_get_state(psb_handle* h, const char* field, void* dst_host, int64_t nbytes, void* cuda_stream) {
  int64_t n = 0;
  void* p = nullptr;
  psb_status st = psb_state_info(h, field, &n, nullptr, &p);
  if (st != PSB_OK) return st;
  if (nbytes != n) return fail(PSB_ERR_VALUE, "state field '%s' has %lld bytes (got %lld)", field, (long long)n, (long long)nbytes);
  CUDA_TRY(cudaMemcpyAsync(dst_host, p, n, cudaMemcpyDeviceToHost, (cudaStream_t)cuda_stream));
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)cuda_stream));
  return PSB_OK;
}

psb_status psb_set_state(psb_handle* h, const char* field, const void* src_host, int64_t nbytes, void* cuda_stream) {
  int64_t n = 0;
  void* p = nullptr;
  psb_status st = psb_state_info(h, field, &n, nullptr, &p);
  if (st != PSB_OK) return st;
  if (nbytes != n) return fail(PSB_ERR_VALUE, "state field '%s' has %lld bytes (got %lld)", field, (long long)n, (long long)nbytes);
  CUDA_TRY(cudaMemcpyAsync(p, src_host, n, cudaMemcpyHostToDevice, (cudaStream_t)cuda_stream));
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)cuda_stream));
  if (std::string(field) == "ncalls") h->host_ncalls = *reinterpret_cast<const long long*>(src_host);
  return PSB_OK;
}

// Engine arrays that live in ordinary (pageable) host memory cross PCIe at a third of the rate of page-locked
// memory (measured on the B200 boxes: 19.6 vs 53.7 GB/s, tools/native/pin_probe.cu) because every copy is staged
// through the driver's bounce buffers.  The arrays of a CPU engine are long-lived, so the first psb_step_host that
// sees one page-locks it in place (cudaHostRegister; undone by psb_destroy).  PSB_NO_HOST_REGISTER=1 disables it.
static void pin_host_range(psb_handle* h, const void* p, size_t bytes) {
  if (!p || !bytes) return;
  for (const auto& r : h->host_ranges)
    if (r.first == p && r.second >= bytes) return;
  h->host_ranges.emplace_back(p, bytes);
  static const bool off = getenv("PSB_NO_HOST_REGISTER") != nullptr;
  if (off) return;
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return; }
  if (at.type != cudaMemoryTypeUnregistered) return;  // already page-locked (or managed / device memory)
  if (cudaHostRegister(const_cast<void*>(p), bytes, cudaHostRegisterDefault) == cudaSuccess)
    h->host_registered.push_back(const_cast<void*>(p));
  else
    cudaGetLastError();  // e.g. part of the range is registered already: copies still work, just slower
}

// host snapshot -> the handle's device staging buffers (allocated on first use)
static psb_status stage_host_snapshot(psb_handle* h, const psb_snapshot* hs, cudaStream_t stream, bool with_forces,
                                      psb_snapshot* out, int64_t* h2d_bytes, size_t* force_bytes) {
  const long long N = h->layout.natoms, rb = h->layout.real_bytes;
  size_t b_pos = 0, b_vel = 0, b_frc = 0, b_ids = 0, b_img = 0, b_mass = 0, b_types = 0, b_tags = 0;
  switch (h->layout.layout) {
    case PSB_LAYOUT_HOOMD: b_pos = b_vel = b_frc = 4 * rb * N; b_ids = 4 * N; b_img = 12 * N; break;
    case PSB_LAYOUT_OPENMM_CUDA: b_pos = b_vel = 4 * rb * N; b_frc = 24 * N; b_ids = 4 * N; break;
    case PSB_LAYOUT_LAMMPS: b_pos = b_vel = b_frc = 24 * N; b_ids = 4 * (N + 1); b_img = 4 * N; b_types = 4 * N; b_mass = 8 * (size_t)std::max(h->layout.ntypes, 1); break;
    default: b_pos = b_vel = b_frc = 3 * rb * N; b_ids = hs->ids ? 4 * N : 0; b_img = hs->images ? 12 * N : 0; b_mass = rb * N; break;
  }
  if (!h->layout.unwrap) b_img = 0;
  if (h->tags_class && hs->tags) b_tags = 4 * N;
  if (!h->layout.need_momenta) b_vel = b_mass = b_types = 0;
  if (!h->have_staging) {
    psb_status st;
    char* p;
    auto grab = [&](size_t bytes, const void** dst) -> psb_status {
      if (!bytes) { *dst = nullptr; return PSB_OK; }
      if ((st = dev_alloc(h, &p, bytes, false)) != PSB_OK) return st;
      *dst = p;
      return PSB_OK;
    };
    const void* f = nullptr;
    if (grab(b_pos, &h->dev_snap.positions) || grab(b_vel, &h->dev_snap.vel_mass) || grab(b_frc, &f) ||
        grab(b_ids, &h->dev_snap.ids) || grab(b_img, &h->dev_snap.images) ||
        grab(b_mass, &h->dev_snap.masses) || grab(b_types, &h->dev_snap.types) ||
        grab(h->tags_class ? 4 * (size_t)N : 0, &h->dev_snap.tags))
      return PSB_ERR_CUDA;
    h->dev_snap.forces = const_cast<void*>(f);
    h->have_staging = true;
  }
  psb_snapshot d = h->dev_snap;
  memcpy(d.box_diag, hs->box_diag, sizeof(d.box_diag));
  d.dt = hs->dt;
  int64_t up = 0;
  auto push = [&](const void* dst, const void* src, size_t bytes) -> cudaError_t {
    if (!bytes || !src) return cudaSuccess;
    pin_host_range(h, src, bytes);
    up += (int64_t)bytes;
    return cudaMemcpyAsync(const_cast<void*>(dst), src, bytes, cudaMemcpyHostToDevice, stream);
  };
  // one queue: a second copy stream did not shorten the step (measured: the link is the bound, 198 vs 203 us)
  CUDA_TRY(push(d.ids, hs->ids, b_ids));
  if (b_tags) CUDA_TRY(push(d.tags, hs->tags, b_tags)); else d.tags = nullptr;
  CUDA_TRY(push(d.positions, hs->positions, b_pos));
  CUDA_TRY(push(d.images, hs->images, b_img));
  CUDA_TRY(push(d.vel_mass, hs->vel_mass, b_vel));
  CUDA_TRY(push(d.masses, hs->masses, b_mass));
  CUDA_TRY(push(d.types, hs->types, b_types));
  if (with_forces) CUDA_TRY(push(d.forces, hs->forces, b_frc));
  *out = d;
  if (h2d_bytes) *h2d_bytes = up;
  if (force_bytes) *force_bytes = b_frc;
  return PSB_OK;
}

// ---- psb_step_host with S << N: selected rows only ----------------------------------------------------------------
// A CPU engine with a handful of CV atoms (a peptide's dihedrals in 1e5 atoms of solvent) should not ship its whole
// arrays over PCIe every step.  The rows of the selected atoms are gathered on the host into ONE pinned block
// [positions | images | velocities | types or inverse masses | forces] (a copy, no arithmetic), go up in ONE transfer
// into a device block of the same shape, the kernels run on that compact snapshot -- `ids` replaced by a static
// tag -> compact row table in the layout's own convention, so no kernel knows the difference -- and the force rows
// come back in one transfer and are written to their slots.
struct CompactGeom {
  size_t pos, img, vel, extra, frc;          // bytes per row (0: array not needed)
  size_t o_pos, o_img, o_vel, o_extra, o_frc, total;  // offsets of the segments for Su rows
};
static CompactGeom compact_geom(const psb_handle* h, const psb_snapshot* hs) {
  CompactGeom g{};
  const size_t rb = (size_t)h->layout.real_bytes, Su = h->sel_tags.size();
  switch (h->layout.layout) {
    case PSB_LAYOUT_HOOMD: g.pos = g.vel = g.frc = 4 * rb; g.img = 12; break;
    case PSB_LAYOUT_LAMMPS: g.pos = g.vel = g.frc = 24; g.img = 4; g.extra = 4; break;           // extra: per-atom types
    default: g.pos = g.vel = g.frc = 3 * rb; g.img = hs->images ? 12 : 0; g.extra = rb; break;   // extra: inverse masses
  }
  if (!h->layout.unwrap) g.img = 0;
  if (!h->layout.need_momenta) g.vel = g.extra = 0;
  auto seg = [&](size_t row, size_t& off, size_t& cursor) { off = cursor; cursor += (row * Su + 255) & ~(size_t)255; };
  size_t c = 0;
  seg(g.pos, g.o_pos, c); seg(g.img, g.o_img, c); seg(g.vel, g.o_vel, c); seg(g.extra, g.o_extra, c);
  const size_t up = c;
  seg(g.frc, g.o_frc, c);
  (void)up;
  g.total = c;
  return g;
}

static psb_status stage_selected_rows(psb_handle* h, const psb_snapshot* hs, cudaStream_t stream, bool with_forces,
                                      psb_snapshot* out, int64_t* h2d_bytes, CompactGeom* geom) {
  const int L = h->layout.layout;
  const size_t Su = h->sel_tags.size(), N = (size_t)h->layout.natoms;
  const CompactGeom g = compact_geom(h, hs);
  if (!h->have_staging) {  // device block + static tag -> compact row table, pinned host block
    char* block = nullptr;
    psb_status st = dev_alloc(h, &block, g.total, false);
    if (st != PSB_OK) return st;
    std::vector<int32_t> table(N + 1, 0);  // LAMMPS indexes the map at tag + 1
    for (size_t u = 0; u < Su; ++u) table[h->sel_tags[u] + (L == PSB_LAYOUT_LAMMPS ? 1 : 0)] = (int32_t)u;
    int32_t* d_table = nullptr;
    if ((st = dev_upload(h, &d_table, table)) != PSB_OK) return st;
    h->d_ids_compact = d_table;
    if (cudaMallocHost(&h->pin, g.total) != cudaSuccess) return fail(PSB_ERR_CUDA, "cannot allocate pinned staging memory");
    h->pin_bytes = g.total;
    h->dev_snap = psb_snapshot{};
    h->dev_snap.positions = block + g.o_pos;
    h->dev_snap.images = g.img ? block + g.o_img : nullptr;
    h->dev_snap.vel_mass = g.vel ? block + g.o_vel : nullptr;
    h->dev_snap.forces = block + g.o_frc;
    h->dev_snap.ids = d_table;
    if (L == PSB_LAYOUT_LAMMPS) {
      h->dev_snap.types = g.extra ? block + g.o_extra : nullptr;
      if (g.extra) {
        char* m = nullptr;
        if ((st = dev_alloc(h, &m, 8 * (size_t)std::max(h->layout.ntypes, 1), false)) != PSB_OK) return st;
        h->dev_snap.masses = m;
      }
    } else if (L == PSB_LAYOUT_PLAIN3) {
      h->dev_snap.masses = g.extra ? block + g.o_extra : nullptr;
    }
    h->sel_slots.resize(Su);
    h->have_staging = true;
  }
  // engine slots of the selected atoms, then the row copies
  const int32_t* ids = reinterpret_cast<const int32_t*>(hs->ids);
  for (size_t u = 0; u < Su; ++u) {
    const uint32_t tag = h->sel_tags[u];
    uint32_t slot = tag;
    if (L == PSB_LAYOUT_LAMMPS) slot = (uint32_t)ids[tag + 1];
    else if (ids) slot = (uint32_t)ids[tag];
    if (slot >= N) return fail(PSB_ERR_VALUE, "snapshot.ids maps atom %u outside the arrays", tag);
    h->sel_slots[u] = slot;
  }
  auto gather = [&](const void* src, size_t row, size_t off) {
    if (!row || !src) return;
    const char* s = reinterpret_cast<const char*>(src);
    char* d = h->pin + off;
    for (size_t u = 0; u < Su; ++u) memcpy(d + u * row, s + (size_t)h->sel_slots[u] * row, row);
  };
  gather(hs->positions, g.pos, g.o_pos);
  gather(hs->images, g.img, g.o_img);
  gather(hs->vel_mass, g.vel, g.o_vel);
  gather(L == PSB_LAYOUT_LAMMPS ? hs->types : hs->masses, g.extra, g.o_extra);
  if (with_forces) gather(hs->forces, g.frc, g.o_frc);
  char* block = reinterpret_cast<char*>(const_cast<void*>(h->dev_snap.positions)) - g.o_pos;
  const size_t up = with_forces ? g.total : g.o_frc;
  CUDA_TRY(cudaMemcpyAsync(block, h->pin, up, cudaMemcpyHostToDevice, stream));
  int64_t moved = (int64_t)up;
  if (L == PSB_LAYOUT_LAMMPS && g.extra && hs->masses) {
    const size_t mb = 8 * (size_t)std::max(h->layout.ntypes, 1);
    CUDA_TRY(cudaMemcpyAsync(const_cast<void*>(h->dev_snap.masses), hs->masses, mb, cudaMemcpyHostToDevice, stream));
    moved += (int64_t)mb;
  }
  psb_snapshot d = h->dev_snap;
  memcpy(d.box_diag, hs->box_diag, sizeof(d.box_diag));
  d.dt = hs->dt;
  d.tags = nullptr;
  *out = d;
  if (h2d_bytes) *h2d_bytes = moved;
  *geom = g;
  return PSB_OK;
}

psb_status psb_step_host(psb_handle* h, const psb_snapshot* hs, int64_t timestep, void* cuda_stream,
                         int64_t* h2d_bytes, int64_t* d2h_bytes) {
  if (!h || !hs) return fail(PSB_ERR_VALUE, "NULL argument");
  cudaStream_t stream = (cudaStream_t)cuda_stream;
  psb_snapshot d;
  if (h->compact_ok) {
    psb_status st = check_snapshot(h, hs, true);
    if (st != PSB_OK) return st;
    CompactGeom g{};
    const bool biased = h->method.kind != PSB_METHOD_UNBIASED;
    if ((st = stage_selected_rows(h, hs, stream, biased, &d, h2d_bytes, &g)) != PSB_OK) return st;
    if ((st = psb_step(h, &d, timestep, cuda_stream)) != PSB_OK) return st;
    const size_t down = biased ? g.frc * h->sel_tags.size() : 0;
    if (down) CUDA_TRY(cudaMemcpyAsync(h->pin + g.o_frc, d.forces, down, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    if (down) {
      char* f = reinterpret_cast<char*>(hs->forces);
      for (size_t u = 0; u < h->sel_tags.size(); ++u) memcpy(f + (size_t)h->sel_slots[u] * g.frc, h->pin + g.o_frc + u * g.frc, g.frc);
    }
    if (d2h_bytes) *d2h_bytes = (int64_t)down;
    return PSB_OK;
  }
  size_t b_frc = 0;
  psb_status st = stage_host_snapshot(h, hs, stream, true, &d, h2d_bytes, &b_frc);
  if (st != PSB_OK) return st;
  st = psb_step(h, &d, timestep, cuda_stream);
  if (st != PSB_OK) return st;
  if (h->method.kind != PSB_METHOD_UNBIASED)
    CUDA_TRY(cudaMemcpyAsync(hs->forces, d.forces, b_frc, cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  if (d2h_bytes) *d2h_bytes = h->method.kind != PSB_METHOD_UNBIASED ? (int64_t)b_frc : 0;
  return PSB_OK;
}

psb_status psb_host_release(psb_handle* h) {
  if (!h) return fail(PSB_ERR_VALUE, "NULL handle");
  for (void* p : h->host_registered) cudaHostUnregister(p);
  cudaGetLastError();
  h->host_registered.clear();
  h->host_ranges.clear();
  return PSB_OK;
}

psb_status psb_evaluate_cvs_host(psb_handle* h, const psb_snapshot* hs, void* cuda_stream) {
  if (!h || !hs) return fail(PSB_ERR_VALUE, "NULL argument");
  cudaStream_t stream = (cudaStream_t)cuda_stream;
  psb_snapshot d;
  psb_status st;
  if (h->compact_ok) {
    if ((st = check_snapshot(h, hs, false)) != PSB_OK) return st;
    CompactGeom g{};
    st = stage_selected_rows(h, hs, stream, false, &d, nullptr, &g);
  } else {
    st = stage_host_snapshot(h, hs, stream, false, &d, nullptr, nullptr);
  }
  if (st != PSB_OK) return st;
  st = psb_evaluate_cvs(h, &d, cuda_stream);
  if (st != PSB_OK) return st;
  CUDA_TRY(cudaStreamSynchronize(stream));
  return PSB_OK;
}

psb_status psb_grid_index(const psb_grid_desc* grid, const double* x_host, int64_t n, uint32_t* out_host) {
  if (!grid || !x_host || !out_host || n < 0) return fail(PSB_ERR_VALUE, "invalid arguments");
  if (grid->kind < PSB_GRID_REGULAR || grid->kind > PSB_GRID_CHEBYSHEV || grid->ndim < 1 || grid->ndim > PSB_MAX_GRID_DIMS)
    return fail(PSB_ERR_VALUE, "invalid grid");
  if (n == 0) return PSB_OK;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(PSB_ERR_CUDA, "no CUDA device available: pysages_b200 has no CPU fallback");
  const size_t cnt = (size_t)n * grid->ndim;
  double* dx = nullptr;
  uint32_t* di = nullptr;
  CUDA_TRY(cudaMalloc(&dx, cnt * 8));
  CUDA_TRY(cudaMalloc(&di, cnt * 4));
  CUDA_TRY(cudaMemcpy(dx, x_host, cnt * 8, cudaMemcpyHostToDevice));
  launch_grid_index(*grid, dx, n, di);
  cudaError_t e = cudaMemcpy(out_host, di, cnt * 4, cudaMemcpyDeviceToHost);
  cudaFree(dx);
  cudaFree(di);
  if (e != cudaSuccess) return fail(PSB_ERR_CUDA, "CUDA error %s in psb_grid_index", cudaGetErrorString(e));
  return PSB_OK;
}

psb_status psb_launch_floor(int32_t grid_blocks, int32_t cooperative, int32_t big_params, int32_t grid_barrier,
                            int64_t nlaunches, void* cuda_stream) {
  if (grid_blocks < 1 || nlaunches < 0) return fail(PSB_ERR_VALUE, "invalid arguments");
  if (grid_barrier && !cooperative) return fail(PSB_ERR_VALUE, "a grid barrier needs a cooperative launch");
  static unsigned long long* counter = nullptr;
  if (!counter) CUDA_TRY(cudaMalloc(&counter, 8));
  CUDA_TRY(launch_floor(grid_blocks, cooperative, big_params, grid_barrier, nlaunches, counter, (cudaStream_t)cuda_stream));
  return PSB_OK;
}

psb_status psb_latency_probe(double* out_host16) {
  if (!out_host16) return fail(PSB_ERR_VALUE, "NULL argument");
  double* d = nullptr;
  CUDA_TRY(cudaMalloc(&d, 16 * sizeof(double)));
  cudaError_t e = launch_latency_probe(d, 0);
  if (e == cudaSuccess) e = cudaMemcpy(out_host16, d, 16 * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return fail(PSB_ERR_CUDA, "CUDA error %s in psb_latency_probe", cudaGetErrorString(e));
  return PSB_OK;
}

psb_status psb_fma_peak(double* out_host5) {
  if (!out_host5) return fail(PSB_ERR_VALUE, "NULL argument");
  int dev = 0, sms = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  cudaError_t e = measure_throughput(out_host5, sms);
  if (e != cudaSuccess) return fail(PSB_ERR_CUDA, "CUDA error %s in psb_fma_peak", cudaGetErrorString(e));
  return PSB_OK;
}

psb_status psb_set_option(psb_handle* h, const char* name, int64_t value) {
  if (!h || !name) return fail(PSB_ERR_VALUE, "NULL argument");
  const std::string n(name);
  if (n == "pdl") {
    h->pdl = value ? 1 : 0;
    if (h->cycle_exec) cudaGraphExecDestroy(h->cycle_exec);  // captured launches carry the old attribute
    h->cycle_exec = nullptr;
    h->cycle_key.clear();
    return PSB_OK;
  }
  if (n == "graph") {  // 0: psb_run_cycle issues one plain launch per step
    h->use_graph = value ? 1 : 0;
    return PSB_OK;
  }
  if (n == "slot_stamp_period") {  // tests: make the slot_map launch stamps repeat after `value` launches
    if (value < 0 || value > 0x0FFFFFFF) return fail(PSB_ERR_VALUE, "slot_stamp_period out of range");
    h->model.stamp_period = (uint32_t)value;
    CUDA_TRY(cudaMemcpy(h->d_model, &h->model, sizeof(Model), cudaMemcpyHostToDevice));
    return PSB_OK;
  }
  return fail(PSB_ERR_KEY, "unknown option '%s'", name);
}

// funn.py:261-296 / cff.py:315-350: the network that replaces the binned average (psb_force_net).
psb_status psb_set_force_net(psb_handle* h, const psb_force_net* net, void* cuda_stream) {
  if (!h || !net) return fail(PSB_ERR_VALUE, "NULL argument");
  Model& M = h->model;
  if (M.m.kind != PSB_METHOD_ABF)
    return fail(PSB_ERR_VALUE, "a force network needs an ABF-family method (funn.py:187-210)");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(cuda_stream);
  const ForceNet* target = nullptr;
  size_t nparams_total = 0;
  if (net->n_dense != 0) {
    if (net->n_dense < 1 || net->n_dense > PSB_MAX_DENSE)
      return fail(PSB_ERR_VALUE, "force network: between 1 and %d dense layers (got %d)", PSB_MAX_DENSE, net->n_dense);
    if (net->widths[0] != M.k || (!net->gradient && net->widths[net->n_dense] != M.k))
      return fail(PSB_ERR_VALUE, "force network: input and output width must equal the %d CV components", M.k);
    if (net->activation != PSB_ACT_TANH && net->activation != PSB_ACT_SIN)
      return fail(PSB_ERR_VALUE, "force network: unknown activation %d", net->activation);
    if (!net->params) return fail(PSB_ERR_VALUE, "force network: NULL parameters");
    if (M.m.grid_kind == PSB_GRID_NONE) return fail(PSB_ERR_VALUE, "force network: the method has no grid");
    ForceNet hn{};
    hn.n_dense = net->n_dense;
    hn.activation = net->activation;
    hn.gradient = net->gradient ? 1 : 0;
    hn.omega0 = net->omega0;
    hn.omega = net->omega;
    hn.predict_after = net->predict_after;
    size_t nparams = 0;
    for (int l = 0; l <= net->n_dense; ++l) {
      if (net->widths[l] < 1 || net->widths[l] > PSB_MAX_WIDTH)
        return fail(PSB_ERR_VALUE, "force network: layer widths must lie in [1, %d] (got %d)", PSB_MAX_WIDTH, net->widths[l]);
      hn.widths[l] = net->widths[l];
      if (l < net->n_dense) {
        hn.w_off[l] = (int32_t)nparams;
        nparams += (size_t)net->widths[l] * net->widths[l + 1] + net->widths[l + 1];
      }
    }
    for (int c = 0; c < MAXK; ++c) { hn.mean[c] = net->mean[c]; hn.std[c] = net->std[c]; }
    nparams_total = nparams;
    const size_t bytes = sizeof(ForceNet) + nparams * sizeof(double);
    if (bytes > h->nn_capacity) {  // first network, or a larger topology than before
      CUDA_TRY(cudaStreamSynchronize(stream));
      char* p = nullptr;
      psb_status st = dev_alloc(h, &p, bytes, false);  // the old block stays alive until psb_destroy
      if (st != PSB_OK) return st;
      h->d_nn = p;
      h->nn_capacity = bytes;
    }
    CUDA_TRY(cudaMemcpyAsync(h->d_nn, &hn, sizeof(ForceNet), cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaMemcpyAsync(h->d_nn + sizeof(ForceNet), net->params, nparams * sizeof(double),
                             net->params_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, stream));
    target = reinterpret_cast<const ForceNet*>(h->d_nn);
  }
  const int32_t staged = (target && nparams_total <= (size_t)NN_STAGE) ? (int32_t)nparams_total : 0;
  if (M.nn != target || M.nn_np != staged || (target && M.fs)) {
    M.nn = target;
    M.nn_np = staged;
    if (target) {  // a handle carries a network or a series
      M.fs = nullptr;
      M.fs_staged = 0;
    }
    M.fm_mode = (M.fm_mode & FM_HIST_ONLY) | (M.nn ? FM_NET : 0) | (M.fs ? FM_SERIES : 0);
    // the step kernel reads the model from device memory; the copy is ordered after every step enqueued so far
    CUDA_TRY(cudaMemcpyAsync(h->d_model, &h->model, sizeof(Model), cudaMemcpyHostToDevice, stream));
  }
  return PSB_OK;
}

// spectral_abf.py:181-268: the fitted series whose gradient replaces the binned average (psb_force_series).
psb_status psb_set_force_series(psb_handle* h, const psb_force_series* ser, void* cuda_stream) {
  if (!h || !ser) return fail(PSB_ERR_VALUE, "NULL argument");
  Model& M = h->model;
  if (M.m.kind != PSB_METHOD_ABF)
    return fail(PSB_ERR_VALUE, "a force series needs an ABF-family method (spectral_abf.py:181-205)");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(cuda_stream);
  const ForceSeries* target = nullptr;
  int32_t staged = 0;
  if (ser->kind != 0) {
    if (ser->kind != PSB_SERIES_FOURIER && ser->kind != PSB_SERIES_MONOMIAL)
      return fail(PSB_ERR_VALUE, "force series: unknown kind %d", ser->kind);
    if (M.m.grid_ndim < 1 || M.m.grid_ndim > 2)
      return fail(PSB_ERR_VALUE, "Only 1D and 2D grids are supported");  // approxfun/core.py:74-75
    if (ser->n_terms < 1 || !ser->exponents || !ser->values)
      return fail(PSB_ERR_VALUE, "force series: exponents and values are required");
    const int dims = M.m.grid_ndim, K = ser->n_terms;
    for (int i = 0; i < K * dims; ++i)
      if (ser->exponents[i] < -63 || ser->exponents[i] > 63 || (ser->kind == PSB_SERIES_MONOMIAL && ser->exponents[i] < 0))
        return fail(PSB_ERR_VALUE, "force series: exponent %d out of range", ser->exponents[i]);
    const size_t exp_words = ((size_t)K * dims + 1) / 2;
    const size_t nval = 1 + (size_t)K * (ser->kind == PSB_SERIES_FOURIER ? 2 : 1);
    const size_t head_words = sizeof(ForceSeries) / 8 + exp_words;
    const size_t words = head_words + nval;
    std::vector<unsigned long long> block(head_words, 0ULL);
    ForceSeries* hs = reinterpret_cast<ForceSeries*>(block.data());
    hs->kind = ser->kind;
    hs->K = K;
    hs->dims = dims;
    hs->oob_zero = ser->oob_zero ? 1 : 0;
    hs->predict_after = ser->predict_after;
    hs->val_off = (int32_t)head_words;
    hs->words = (int32_t)words;
    memcpy(block.data() + sizeof(ForceSeries) / 8, ser->exponents, sizeof(int32_t) * (size_t)K * dims);
    if (words * 8 > h->fs_capacity) {
      CUDA_TRY(cudaStreamSynchronize(stream));
      unsigned long long* p = nullptr;
      psb_status st = dev_alloc(h, &p, words, false);  // the old block stays alive until psb_destroy
      if (st != PSB_OK) return st;
      h->d_fs = p;
      h->fs_capacity = words * 8;
    }
    CUDA_TRY(cudaMemcpyAsync(h->d_fs, block.data(), head_words * 8, cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaMemcpyAsync(h->d_fs + head_words, ser->values, nval * 8,
                             ser->values_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, stream));
    target = reinterpret_cast<const ForceSeries*>(h->d_fs);
    staged = words <= sizeof(ForceNet) / 8 + (size_t)NN_STAGE ? (int32_t)words : 0;
  }
  if (M.fs != target || M.fs_staged != staged || (target && M.nn)) {
    M.fs = target;
    M.fs_staged = staged;
    if (target) {  // a handle carries a network or a series
      M.nn = nullptr;
      M.nn_np = 0;
    }
    M.fm_mode = (M.fm_mode & FM_HIST_ONLY) | (M.nn ? FM_NET : 0) | (M.fs ? FM_SERIES : 0);
    CUDA_TRY(cudaMemcpyAsync(h->d_model, &h->model, sizeof(Model), cudaMemcpyHostToDevice, stream));
  }
  return PSB_OK;
}

int64_t psb_launch_count(psb_handle* h) { return h ? h->launches : 0; }
int32_t psb_cycle_uses_graph(psb_handle* h) { return h ? h->graph_replays : 0; }

psb_status psb_launch_info(psb_handle* h, int32_t* grid_blocks, int32_t* block_threads,
                           int32_t* cooperative, int64_t* algorithmic_bytes_per_step) {
  if (!h) return fail(PSB_ERR_VALUE, "NULL handle");
  if (grid_blocks) *grid_blocks = h->grid;
  if (block_threads) *block_threads = BLOCK;
  if (cooperative) *cooperative = h->cooperative;
  if (algorithmic_bytes_per_step) *algorithmic_bytes_per_step = h->algo_bytes;
  return PSB_OK;
}

}  // extern "C"
