/*
 * pysages_b200.h -- C ABI of the B200-native PySAGES bias step.
 *
 * Drop-in boundary for the per-timestep hot path of PySAGES (reference checkout
 * paths below are relative to the reference repository root):
 *
 *   engine force compute -> [ snapshot read -> CVs + Jacobians -> method bias
 *                             -> in-place force scatter ] -> engine integrator
 *
 * Nothing comparable exists in the reference (it is 100 % Python/JAX); every
 * entry point states which reference interface it replaces.  All pointers in
 * `psb_snapshot` are raw CUDA device pointers taken from the engine's DLPack
 * capsules (DLTensor::data + byte_offset); no torch / DLPack / Python types
 * appear in any signature.  All work is enqueued on the caller's CUDA stream
 * (the engine's stream) and the calls return without synchronising, unless a
 * function says otherwise.  There is no CPU fallback: every function fails with
 * PSB_ERR_CUDA when no sm_100 device is usable.
 *
 * Error convention: functions return a psb_status; the message is available from
 * psb_last_error() (thread-local).  The Python layer maps PSB_ERR_VALUE ->
 * ValueError, PSB_ERR_RUNTIME / PSB_ERR_CUDA -> RuntimeError, PSB_ERR_KEY ->
 * KeyError, matching the reference's exception types.
 */
#ifndef PYSAGES_B200_H
#define PYSAGES_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef PSB_MAX_CVS          /* (overridable at build time for experiments; the shipped library uses these) */
#define PSB_MAX_CVS 8        /* collective variables per method (colvars/core.py:228-274 stacks any number) */
#define PSB_MAX_K 8          /* total CV components (xi has shape (1, k))  */
#define PSB_MAX_GROUPS 32    /* atom groups ("points") over all CVs: eight dihedrals */
#endif
#define PSB_MAX_GRID_DIMS 4  /* gridded methods (ABF family, grid metadynamics): k = grid dimensions <= 4 */

typedef struct psb_handle psb_handle;

typedef enum {
  PSB_OK = 0,
  PSB_ERR_VALUE = 1,   /* ValueError   */
  PSB_ERR_RUNTIME = 2, /* RuntimeError */
  PSB_ERR_KEY = 3,     /* KeyError     */
  PSB_ERR_CUDA = 4,    /* CUDA failure / no device / extension missing */
  PSB_ERR_UNSUPPORTED = 5
} psb_status;

/* ---- collective variables: pysages/colvars/ (all modules) ----------------------------------- */
typedef enum {
  PSB_CV_COMPONENT = 1,    /* colvars/coordinates.py:56-71   */
  PSB_CV_DISTANCE = 2,     /* colvars/coordinates.py:74-109  */
  PSB_CV_DISPLACEMENT = 3, /* colvars/coordinates.py:112-148 (3 components) */
  PSB_CV_ANGLE = 4,        /* colvars/angles.py:24-74        */
  PSB_CV_DIHEDRAL = 5,     /* colvars/angles.py:77-128       */
  PSB_CV_ROG = 6,          /* colvars/shape.py:14-57 (value = <r.r>, reference quirk) */
  PSB_CV_RMSD = 7,         /* colvars/orientation.py:27-126  */
  PSB_CV_COORDINATION = 8, /* extension, absent from the reference */
  PSB_CV_GYRATION = 9,     /* colvars/shape.py:85-344: functions of the eigenvalues of the (uncentred)
                              gyration tensor sum_i r_i r_i^T / n; `axis` selects the function */
  PSB_CV_RING = 10         /* colvars/angles.py:131-281: Cremer-Pople puckering of a monocyclic ring (one group,
                              atoms in ring order, >= 3); `axis` selects amplitude q or phase phi.  The
                              two-component RingPuckeringCoordinates is two CVs over the same atoms */
} psb_cv_kind;
#define PSB_CV_ERMSD 11 /* colvars/orientation.py:129-175,177-378: eRMSD of n nucleotides to a reference; one group of 3n
                           ring atoms (C2, then C6/C4 or C4/C6 per base, the reference's order), `reference` (3n, 3),
                           `r0` = unitless cutoff, `weights` -> {a, b} (rescaling lengths); n <= 256.
                           `axis` = 1: coarse-grained ERMSDCG (orientation.py:380-560): the 3 sites per base are B1/B2/B3,
                           `weights` -> {a, b, local_coordinates (4, 3, 3), sequence (n values 0..3 as doubles)} */

/* psb_cv_desc.axis for PSB_CV_RING */
typedef enum { PSB_RING_AMPLITUDE = 0, PSB_RING_PHASE = 1 } psb_ring_mode;

/* psb_cv_desc.axis for PSB_CV_GYRATION (eigenvalues ascending, l[0] <= l[1] <= l[2]) */
typedef enum {
  PSB_GYR_MOMENT_0 = 0,    /* PrincipalMoment(axis = 0, 1, 2): l[axis] (shape.py:85-176) */
  PSB_GYR_MOMENT_1 = 1,
  PSB_GYR_MOMENT_2 = 2,
  PSB_GYR_ASPHERICITY = 3,  /* l[2] - (l[0] + l[1]) / 2 (shape.py:203-221) */
  PSB_GYR_ACYLINDRICITY = 4, /* l[mm] - l[nn] with the reference's index pairs (shape.py:224-287) */
  PSB_GYR_ANISOTROPY = 5    /* (3 sum l^2 / (sum l)^2 - 1) / 2 (shape.py:290-344) */
} psb_gyration_mode;

/* One CV after the reference's index processing (colvars/core.py:184-207, 277-295):
 * `tags` holds the particle tags of all groups back to back; group g spans
 * tags[group_offsets[g] .. group_offsets[g+1]).  Point CVs take the barycentre of each
 * group as one point (a bare index is a group of one). */
typedef struct {
  int32_t kind;                  /* psb_cv_kind */
  int32_t axis;                  /* Component: 0,1,2 */
  int32_t n_groups;              /* Component/RoG/RMSD/Gyration: 1; Distance/Displacement/Coordination: 2;
                                    Angle: 3; Dihedral: 4 */
  int32_t reserved;
  const uint32_t* tags;          /* host pointer, group_offsets[n_groups] entries */
  const uint32_t* group_offsets; /* host pointer, n_groups + 1 entries */
  const double* reference;       /* RMSD: (n, 3) reference coordinates (host) */
  const double* weights;         /* RMSD: n weights, or NULL for uniform 1/n  */
  double r0;                     /* Coordination: switching radius */
  int32_t nn, mm;                /* Coordination: exponents (default 6, 12); Gyration acylindricity:
                                    eigenvalue indices (l[mm] - l[nn]) */
} psb_cv_desc;

/* ---- grids: pysages/grids.py:30-158 ----------------------------------------------- */
typedef enum { PSB_GRID_NONE = 0, PSB_GRID_REGULAR = 1, PSB_GRID_PERIODIC = 2, PSB_GRID_CHEBYSHEV = 3 } psb_grid_kind;

typedef struct {
  int32_t kind; /* psb_grid_kind */
  int32_t ndim;
  double lower[PSB_MAX_GRID_DIMS];
  double upper[PSB_MAX_GRID_DIMS];
  int32_t shape[PSB_MAX_GRID_DIMS];
} psb_grid_desc;

/* ---- methods: pysages/methods/{harmonic_bias,metad,abf,funn,restraints}.py --------- */
typedef enum {
  PSB_METHOD_UNBIASED = 0, /* CVs only, no force (methods/unbiased.py:64-75) */
  PSB_METHOD_HARMONIC = 1, /* methods/harmonic_bias.py:113-132 */
  PSB_METHOD_METAD = 2,    /* methods/metad.py:156-323 */
  PSB_METHOD_ABF = 3       /* methods/abf.py:142-258 (FUNN's per-step accumulation, funn.py:187-210,
                              is the same update with `force` supplied by psb_set_state) */
} psb_method_kind;

typedef struct {
  int32_t kind; /* psb_method_kind */
  int32_t variant; /* PSB_METHOD_ABF only: 0 = ABF / FUNN / SpectralABF; 1 = histogram only (ANN, methods/ann.py:153-167:
                      no momenta, no force sums; the force is the network gradient or zero); 2 = ABF plus the second
                      histogram `histp` of CFF (methods/cff.py:236-239), reset by the host at every training step */
  /* HarmonicBias: harmonic_bias.py:78-107 canonical (k x k) spring matrix, row-major; center (k) */
  double kspring[PSB_MAX_K * PSB_MAX_K];
  double center[PSB_MAX_K];
  /* Metadynamics: metad.py:140-150 */
  double height;
  double sigma[PSB_MAX_K];
  double periods[PSB_MAX_K]; /* colvars/utils.py:13-19: 2*pi for exact Angle/DihedralAngle, else +inf */
  int64_t stride;
  int64_t ngaussians;
  int32_t well_tempered; /* deltaT is not None */
  int32_t shared_hills;  /* extension: multiple walkers exchange hills (psb_walker_*) */
  double deltaT;
  double kB;
  /* ABF: abf.py:116-118 */
  int64_t abf_N;
  int32_t use_pinv;
  int32_t has_restraints;
  /* CVRestraints after canonicalize (restraints.py:49-61) */
  double r_lower[PSB_MAX_K], r_upper[PSB_MAX_K], r_kl[PSB_MAX_K], r_ku[PSB_MAX_K];
  /* helpers.to_force_units (backends/lammps.py:33,137-139); 1.0 elsewhere */
  double force_unit_factor;
  psb_grid_desc grid; /* kind == PSB_GRID_NONE when the method has no grid */
} psb_method_desc;

/* ---- engine memory layouts: pysages/backends/{hoomd,openmm,lammps}.py -------------- */
typedef enum {
  /* hoomd.py:148-202: positions (N,4) xyz+type, vel_mass (N,4) vxyz+mass, forces (N,4) fxyz+pe,
     ids = rtag (N,) u32 tag->slot, images (N,3) i32 */
  PSB_LAYOUT_HOOMD = 1,
  /* openmm.py:64-89,100-143 on the CUDA platform: positions (Np,4), vel_mass (Np,4) w = 1/m,
     forces (3,Np) int64 fixed point 2^32, ids (Np,) i32 slot->atom (argsort'ed every step) */
  PSB_LAYOUT_OPENMM_CUDA = 2,
  /* lammps.py:87-113,182-223: x,v,f (N,3); ids = tags_map (N+1,) i32 with tag t at [t+1];
     images (N,) packed i32 (10/10/10 bits); masses per type (double) + types (N,) i32 */
  PSB_LAYOUT_LAMMPS = 3,
  /* openmm.py on the CPU/Reference platforms, staged on the device: positions, velocities,
     forces (N,3); masses = inverse masses (N,); ids (N,) i32 or NULL (identity);
     images (N,3) i32 or NULL */
  PSB_LAYOUT_PLAIN3 = 4
} psb_layout_kind;

typedef struct {
  int32_t layout;     /* psb_layout_kind */
  int32_t real_bytes; /* 4 or 8: dtype of positions / velocities / (non fixed-point) forces */
  int64_t natoms;     /* rows of `positions` (padded count for OpenMM-CUDA) */
  int32_t unwrap;     /* method.requires_box_unwrapping (colvars/core.py:50; hoomd.py:177-186) */
  int32_t need_momenta; /* "momenta" in method.snapshot_flags (abf.py:114) */
  int32_t dense_outputs; /* 1: also keep the reference's dense state.bias (N,3) and J (k,3N) */
  int32_t ntypes;        /* LAMMPS: entries of the per-type mass array (ntypes + 1) */
  int32_t has_tags;      /* HOOMD, LAMMPS: the engine will hand over psb_snapshot.tags (slot -> tag) with every snapshot: the
                            slot-ordered class is then planned for it (one CTA per SM, so that the next step's gather --
                            engine arrays only -- overlaps this step's tail; chosen from ~1.5e5 selected atoms upward) */
  int32_t reserved;
} psb_layout_desc;

/* Device pointers of one snapshot (backends/snapshot.py:34-46).  Unused ones may be NULL. */
typedef struct {
  const void* positions;
  const void* vel_mass;
  void* forces; /* modified in place (hoomd.py:219-230, openmm.py:160-170, lammps.py:163-170) */
  const void* ids;
  const void* images;
  const void* masses; /* LAMMPS: per-type masses (double); PLAIN3: inverse masses */
  const void* types;  /* LAMMPS: per-atom types (i32) */
  double box_diag[3]; /* diag(H) -- the reference unwraps with the diagonal only (hoomd.py:180) */
  double dt;
  /* Optional, HOOMD and LAMMPS layouts: the slot -> tag array (HOOMD: ParticleData's `tag`, the inverse of `ids` = rtag,
     hoomd-dlext exports both; LAMMPS: atom->tag, int32 1-based ids, the inverse of `ids` = tags_map).  With it the slot-ordered class streams the engine's arrays in memory order and reads the group of
     a slot from a tag -> group table: no tag -> slot inversion pass, one grid barrier fewer.  NULL: `ids` only. */
  const void* tags;
} psb_snapshot;

/* ---- lifecycle -------------------------------------------------------------------- */

/* Replaces SamplingMethod.build(snapshot, helpers) + initialize() (methods/core.py:108-116,
 * e.g. metad.py:152-185): validates, uploads index tables, allocates device state. */
psb_status psb_create(const psb_cv_desc* cvs, int32_t ncvs, const psb_method_desc* method,
                      const psb_layout_desc* layout, psb_handle** out);
void psb_destroy(psb_handle* h);
const char* psb_last_error(void);

/* Replaces Sampler._update_callback: method_update + bias (hoomd.py:168-171 and the openmm /
 * lammps equivalents): one fused launch sequence on `cuda_stream`, no host sync. */
psb_status psb_step(psb_handle* h, const psb_snapshot* snap, int64_t timestep, void* cuda_stream);

/* Replaces the first CV evaluation done by initialize() (e.g. harmonic_bias.py:119-122):
 * evaluates xi on `snap` without touching forces, state counters or histograms. */
psb_status psb_evaluate_cvs(psb_handle* h, const psb_snapshot* snap, void* cuda_stream);

/* Engine-loop emulation for benchmarks: psb_step on snaps[i % nsnaps] for nsteps steps.  Whenever
 * nsteps >= nsnaps one pass over the ring is captured as a CUDA graph (once per distinct ring) and replayed;
 * the remainder, and shorter calls, are plain launches (PSB_NO_GRAPH=1: always plain launches). */
psb_status psb_run_cycle(psb_handle* h, const psb_snapshot* snaps, int32_t nsnaps, int64_t nsteps,
                         void* cuda_stream);

/* Same step with HOST snapshot buffers (engine running on the CPU, bias on the GPU): copies
 * the arrays the method needs to the device, steps, copies forces back, and synchronises.
 * When the CVs touch few atoms (8 * selected <= natoms) only the rows of those atoms travel: gathered on the host
 * into one pinned block, one transfer up, the force rows back in one transfer and written to their slots (a copy on
 * the host, no arithmetic); the byte counters report what actually moved. */
psb_status psb_step_host(psb_handle* h, const psb_snapshot* host_snap, int64_t timestep,
                         void* cuda_stream, int64_t* h2d_bytes, int64_t* d2h_bytes);
/* psb_step_host page-locks the engine's host arrays in place the first time it sees them (cudaHostRegister: a
 * pageable array crosses PCIe at a third of the rate, 19.6 vs 53.7 GB/s measured) and keeps them registered until
 * psb_destroy.  An engine that frees or reallocates its arrays earlier calls this first.  PSB_NO_HOST_REGISTER=1
 * in the environment disables the registration altogether. */
psb_status psb_host_release(psb_handle* h);
/* psb_evaluate_cvs for HOST snapshot buffers (initial xi of an engine running on the CPU,
 * openmm.py:64-89 with the CPU / Reference platform); synchronises. */
psb_status psb_evaluate_cvs_host(psb_handle* h, const psb_snapshot* host_snap, void* cuda_stream);

/* ---- many independent simulations per launch ------------------------------------------------- */
/* Replaces the replica loop of UmbrellaIntegration's run (methods/umbrella_integration.py:128-203: one
 * `_run_replica` per window, each with its own per-step callback): the windows of one process step in lockstep,
 * ONE cooperative launch per MD step for all of them.  Every simulation keeps its own handle (state, index
 * tables, workspaces, barrier counters) and its own CTA range inside the launch, so its results are bit-identical
 * to psb_step on that handle.  All members of a batch share engine layout, dtype, method kind and kernel class;
 * CVs that need helper kernels (pairwise coordination, eRMSD), dense outputs and the slot-ordered class are
 * refused with PSB_ERR_UNSUPPORTED (step those with psb_step).  `snaps` holds one psb_snapshot per simulation, in
 * the order of `handles`.  Launches are stream-ordered like psb_step; use one stream per batch. */
typedef struct psb_batch psb_batch;
psb_status psb_batch_create(psb_handle* const* handles, int32_t n, psb_batch** out);
void psb_batch_destroy(psb_batch* b);
psb_status psb_batch_step(psb_batch* b, const psb_snapshot* snaps, int64_t timestep, void* cuda_stream);
/* psb_run_cycle for a batch: `snaps` is (nsnaps, n) row-major, step i uses row i % nsnaps; one pass over the ring is
 * captured as a CUDA graph whenever nsteps >= nsnaps. */
psb_status psb_batch_run_cycle(psb_batch* b, const psb_snapshot* snaps, int32_t nsnaps, int64_t nsteps,
                               void* cuda_stream);
int64_t psb_batch_launch_count(psb_batch* b);

/* ---- state access (callbacks, checkpoint/restart: serialization.py:44-93) -------------- */
/* Field names follow the reference State tuples: "xi", "bias", "ncalls", "heights",
 * "centers", "sigmas", "grid_potential", "grid_gradient", "idx", "hist", "Fsum", "force",
 * "Wp", "Wp_"; extensions: "cv_force" (dV/dxi), "energy" (bias energy), "jacobian". */
psb_status psb_state_info(psb_handle* h, const char* field, int64_t* nbytes, int32_t* dtype_code,
                          void** device_ptr);
psb_status psb_get_state(psb_handle* h, const char* field, void* dst_host, int64_t nbytes,
                         void* cuda_stream);
psb_status psb_set_state(psb_handle* h, const char* field, const void* src_host, int64_t nbytes,
                         void* cuda_stream);

/* ---- multiple walkers (extension) ---------------------------------------- */
/* Deposit steps are split so an all-gather can sit between the two halves:
 *   psb_walker_begin : CVs + this walker's candidate hill -> `record` (k + 2 doubles:
 *                      centre[k], height, valid) in device memory;
 *   (caller all-gathers the records over NCCL on the same stream)
 *   psb_walker_finish: appends `nwalkers` records in rank order, then bias + force scatter. */
psb_status psb_walker_begin(psb_handle* h, const psb_snapshot* snap, int64_t timestep,
                            double* record_dev, void* cuda_stream);
psb_status psb_walker_finish(psb_handle* h, const psb_snapshot* snap, const double* records_dev,
                             int32_t nwalkers, void* cuda_stream);
int32_t psb_walker_record_doubles(psb_handle* h);
/* 1 when the NEXT psb_step call is a deposit step (host mirror of metad.py:192-193). */
int32_t psb_next_step_deposits(psb_handle* h);

/* ---- utilities ----------------------------------------------------------------------- */
/* grids.py:109-158 on the device: bin indices of n points x (n, ndim) -> out (n, ndim) u32. */
psb_status psb_grid_index(const psb_grid_desc* grid, const double* x_host, int64_t n,
                          uint32_t* out_host);
/* Kernel launches issued since creation (for bench.py's gpu_launches) and the launch shape. */
/* Options: "pdl" (default 1; PSB_NO_PDL=1 in the environment sets 0): consecutive step kernels of a
 * stream use programmatic dependent launch -- the next step's launch, model copy and its gather of the
 * engine arrays this library never writes (ids, positions, images, velocities) overlap the previous
 * step's finalizer / scatter; forces, method state and workspaces are touched only after the previous
 * kernel has completed.  A kernel that follows an engine kernel is never launched early. */
/* "graph" (default 1): psb_run_cycle may capture and replay a CUDA graph; 0 = one plain launch per step.
 * "slot_stamp_period" (tests): make the launch stamps of the slot-ordered class repeat after `value` launches. */
psb_status psb_set_option(psb_handle* h, const char* name, int64_t value);
/* Continuous force model of the ABF family (row K of SURVEY.md 8a).  FUNN (methods/funn.py:187-210)
 * and CFF (methods/cff.py:223-244) run the ABF accumulation every step and, once
 * `ncalls > predict_after` (funn.py:190: 2 * train_freq, cff.py:226: train_freq), take the
 * generalized force from a multilayer perceptron instead of the binned average
 * (funn.py:261-296 build_force_estimator, cff.py:315-350):
 *     F = std * mlp(scale(float32(xi))) + mean
 * with scale(x) = (x - grid.lower) * 2 / grid.size - 1 (approxfun/core.py:99-104), dense layers
 * h W + b with the activation between them (ml/models.py:44-82) and all arithmetic in fp64 after the
 * float32 rounding of xi (funn.py:283).  Restraints outside the grid take precedence as in ABF.
 * `params` is the flat parameter vector of ml/utils.py:43-53: for every dense layer W (in, out)
 * row-major, then b (out).  The fit itself is cadenced host-side work (every train_freq steps) and
 * hands its result over through this call; the copy is ordered on `cuda_stream`, so steps already
 * enqueued keep the previous network.  Only PSB_METHOD_ABF handles accept it. */
#define PSB_MAX_DENSE 8   /* dense layers, output layer included */
#define PSB_MAX_WIDTH 128 /* units per layer */
typedef enum { PSB_ACT_TANH = 0, PSB_ACT_SIN = 1 } psb_activation;
typedef struct {
  int32_t n_dense;                   /* 0 removes the network (plain ABF again) */
  int32_t widths[PSB_MAX_DENSE + 1]; /* widths[0] = widths[n_dense] = k */
  int32_t activation;                /* psb_activation; PSB_ACT_SIN is sin(omega * x) (ml/models.py:113-121,140-143) */
  int32_t params_on_device;          /* `params` is a device pointer */
  int32_t gradient;                  /* 1: F = std * d(sum of outputs)/d xi (ANN, ann.py:240-248) instead of std * output + mean */
  int32_t reserved;
  double omega0, omega;              /* PSB_ACT_SIN only: frequency of the first / of the other layers */
  int64_t predict_after;
  const double* params;
  double mean[PSB_MAX_K];
  double std[PSB_MAX_K];
} psb_force_net;
psb_status psb_set_force_net(psb_handle* h, const psb_force_net* net, void* cuda_stream);

/* SpectralABF (methods/spectral_abf.py:181-268): ABF accumulation every step; once `ncalls > predict_after`
 * (fit_threshold) the force is the gradient of a series fitted to the binned mean force
 * (approxfun/core.py:311-336 build_grad_evaluator), with x = scale(xi) in [-1, 1]^n (core.py:99-104):
 *   Fourier  (periodic grid):  F_d = scale * sum_k (2 pi / size_d) n_kd [cos(pi n_k.x) c_{K+k} - sin(pi n_k.x) c_k]
 *   monomial (Chebyshev grid): F_d = scale * sum_k (2 / size_d) n_kd x_d^(n_kd - 1) prod_{e != d} x_e^(n_ke) c_k
 * `exponents` are the K integer tuples of collect_exponents (core.py:73-96); `values` = [scale, c_0 ...]
 * (2K coefficients for Fourier: the block multiplying the imaginary parts first, core.py:321-324; K for
 * monomials) and may live in device memory, so a refit never synchronises with the host.  Outside the
 * grid the restraint force applies, or -- `oob_zero` -- no force at all (spectral_abf.py:250-266).
 * Replaces a force network set earlier, and vice versa.  1-D and 2-D grids only, like the reference. */
typedef enum { PSB_SERIES_FOURIER = 1, PSB_SERIES_MONOMIAL = 2 } psb_series_kind;
typedef struct {
  int32_t kind;             /* psb_series_kind; 0 removes the series (plain ABF again) */
  int32_t n_terms;          /* K */
  int32_t oob_zero;
  int32_t values_on_device;
  int64_t predict_after;
  const int32_t* exponents; /* (K, ndim) row-major, host memory */
  const double* values;     /* 1 + 2K (Fourier) or 1 + K (monomial) doubles */
} psb_force_series;
psb_status psb_set_force_series(psb_handle* h, const psb_force_series* series, void* cuda_stream);

/* ---- cadenced refit on the device --------------------------------------------------------------- */
/* Replaces the Levenberg-Marquardt fit FUNN / ANN / CFF's force network run every `train_freq` steps
 * (ml/training.py:42-56 build_fitting_function + ml/optimizers.py:188-232, sum-of-squared-errors loss with float32
 * residuals and L2 regularisation, ml/objectives.py) for tanh multilayer perceptrons with `scale` as input
 * transform (approxfun/core.py:99-104): Jacobian by closed-form back-propagation, J^T J by a tiled fp64 kernel,
 * damped normal equations by a blocked Cholesky in one CTA, candidate cost, gain ratio and damping update -- the
 * whole data-dependent loop runs from device-resident control state, enqueued on `cuda_stream` without any host
 * synchronisation.  `x` (n, widths[0]), `y` (n, widths[n_dense]) and `params` (flat, ml/utils.py:43-53; updated
 * in place) are device pointers.  At most 1024 parameters.  Sine layers and the gradient / Sobolev losses
 * (Sirens, CFF's energy network) stay with the host layer (pysages_b200/ml.py). */
typedef struct {
  int32_t n_dense;
  int32_t widths[PSB_MAX_DENSE + 1];
  int32_t activation; /* PSB_ACT_TANH */
  int32_t max_iters;  /* accepted steps (ml/optimizers.py:121-131) */
  int32_t max_loops;  /* loop passes enqueued; 0: 2 * max_iters + 64 */
  int32_t reserved;
  int64_t n;          /* samples */
  const double* x;
  const double* y;
  double* params;
  double lower[PSB_MAX_K], upper[PSB_MAX_K]; /* grid bounds of the input transform */
  double reg;         /* L2Regularization.coeff */
  float mu_0, mu_c, mu_min, mu_max, rho_c, rho_min; /* LevenbergMarquardtParams (ml/optimizers.py:40-51) */
} psb_lm_problem;
typedef struct psb_lm_workspace psb_lm_workspace;
/* *ws == NULL: a workspace is created; reuse it for later fits of the same shape (the iteration's CUDA graph is kept) */
psb_status psb_lm_fit(const psb_lm_problem* problem, psb_lm_workspace** ws, void* cuda_stream);
/* accepted iterations, final cost, damping and loop passes of the last fit (synchronises `cuda_stream`) */
psb_status psb_lm_result(psb_lm_workspace* ws, int32_t* iters, double* cost, float* mu, int32_t* loops, void* cuda_stream);
void psb_lm_destroy(psb_lm_workspace* ws);

/* SpectralABF refit (spectral_abf.py:216-219 + approxfun/core.py:232-258) in two kernels, stream-ordered:
 * out[0] = scale (largest population standard deviation of the mean-force components over the grid, 1 if zero),
 * out[1 .. nterms] = pinv . ((Fsum / max(hist, 1) [- mean, periodic grids]) / scale).  All pointers are device
 * pointers; `pinv` is the (nterms, npoints * k) pseudo-inverse of the gradient Vandermonde matrix (row-major, built
 * once by the host layer), `work` holds npoints * k doubles. */
typedef struct {
  int64_t npoints;    /* grid points */
  int32_t k;          /* force components per point */
  int32_t nterms;     /* rows of pinv = coefficients */
  int32_t periodic;   /* 1: remove the mean of every component (Fourier basis) */
  int32_t reserved;
  const void* hist;   /* (npoints) uint32 */
  const double* Fsum; /* (npoints, k) */
  const double* pinv; /* (nterms, npoints * k) */
  double* out;        /* (1 + nterms) */
  double* work;       /* (npoints * k) */
} psb_series_fit_desc;
psb_status psb_series_fit(const psb_series_fit_desc* desc, void* cuda_stream);

int64_t psb_launch_count(psb_handle* h);
/* 1 when the last psb_run_cycle replayed a CUDA graph, 0 when it issued plain launches. */
int32_t psb_cycle_uses_graph(psb_handle* h);
psb_status psb_launch_info(psb_handle* h, int32_t* grid_blocks, int32_t* block_threads,
                           int32_t* cooperative, int64_t* algorithmic_bytes_per_step);
/* Launch-latency floor for the bench report: enqueues `nlaunches` kernels of
 * `grid_blocks` x 256 threads that do no work (optionally one grid-wide arrive/wait, cooperative
 * launches only; optionally a parameter block as large as the step kernel's used to be). */
psb_status psb_launch_floor(int32_t grid_blocks, int32_t cooperative, int32_t big_params, int32_t grid_barrier,
                            int64_t nlaunches, void* cuda_stream);
/* Dependent-issue latency (SM cycles per operation) of DFMA, DADD, FFMA, IMAD, LDS, sqrt, 1/x,
 * rsqrt, floor, fmod, atan2, exp in one warp: the floor of the single-warp finalizer code. */
psb_status psb_latency_probe(double* out_host16);
/* Sustained arithmetic throughput of the device, operations per second over all SMs, measured with 8 independent
 * chains per thread and 8 x 256 threads per SM: out[0] DFMA (x 2 = FP64 FLOP/s: the roofline of the pairwise
 * coordination kernel), out[1] FFMA, out[2] float -> double conversions, out[3] int -> double
 * conversions, out[4] DADD.  Synchronous. */
psb_status psb_fma_peak(double* out_host5);
int32_t psb_abi_version(void);
/* 1 when the handle steps faster on snapshots that carry `tags` (slot-ordered class on the HOOMD layout): the adapter
 * then asks the engine for its slot -> tag array each step (hoomd-dlext `tags`), 0 when the field is ignored. */
int32_t psb_wants_tags(psb_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* PYSAGES_B200_H */
