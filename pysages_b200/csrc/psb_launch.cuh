// psb_launch.cuh -- host-callable launch tables; each psb_inst_*.cu instantiates the kernels
// for one engine layout / dtype and exports one table, so translation units build in parallel.
#pragma once

#include <cstdlib>

#include "psb_device.cuh"
#include "psb_ermsd.cuh"
#include "psb_step.cuh"

namespace psb {

struct StepVTable {
  // `method` is a StepMethod code, `general` a StepClass code (psb_step.cuh)
  cudaError_t (*launch_step)(int method, int general, const Model* M, const Snap* S, int grid, size_t dyn_smem,
                             cudaStream_t stream, int pdl);
  int (*occupancy)(int method, int general, size_t dyn_smem);
  size_t (*static_smem)(int method, int general);  // statically allocated shared memory of that instantiation
  // psb_batch_step: the CTAs of many simulations in one cooperative launch (classes SC_POINT / SC_GENERAL)
  cudaError_t (*launch_batch)(int method, int general, const BatchItem* items, const int32_t* cta_item, int grid,
                              size_t dyn_smem, cudaStream_t stream, int pdl);
  int (*batch_occupancy)(int method, int general, size_t dyn_smem);
  void (*launch_gather)(const Model* M, const Snap* S, int cv, double* xa, int na_pad, double* xb,
                        int nb_pad, cudaStream_t stream);
  void (*launch_ermsd)(const Model* M, const Snap* S, int cv, cudaStream_t stream);
};

using PairLaunchFn = void (*)(int grid, cudaStream_t stream, const double* xa, int na_pad,
                              const double* xb, int nb_pad, int nA, int nB, int nrt, double inv_r02,
                              double* gA, double* gB, double* xi_part);
// R in {1, 4}; H = n / 2 in {1, 2, 3, 4, 6}; nullptr when unsupported
PairLaunchFn pair_launcher(int R, int H);

cudaError_t launch_floor(int grid, int cooperative, int big_params, int use_barrier, long long n,
                         unsigned long long* counter, cudaStream_t stream);
cudaError_t launch_latency_probe(double* out_dev, cudaStream_t stream);
cudaError_t measure_throughput(double* out_host5, int num_sms);
void launch_inverse_permutation(const int* ids, uint32_t* inv, long long n, int blocks, cudaStream_t stream);
void launch_grid_index(const psb_grid_desc& g, const double* x, long long n, uint32_t* out);

extern const StepVTable vt_hoomd_f32, vt_hoomd_f64, vt_openmm_f32, vt_openmm_f64, vt_lammps_f64,
    vt_plain3_f32, vt_plain3_f64;

// kernel entry points of one (layout, dtype): [method][general]
// the TMA-staged streaming class exists for the HOOMD layout; the other layouts fall back to the slot-ordered class
template <int L, typename R, int M>
const void* stream_entry() {
  if constexpr (L == PSB_LAYOUT_HOOMD) return (const void*)step_kernel<L, R, M, SC_STREAM>;
  else return (const void*)step_kernel<L, R, M, SC_SLOT>;
}
template <int L, typename R, int M>
const void* tags_entry() {
  if constexpr (L == PSB_LAYOUT_HOOMD || L == PSB_LAYOUT_LAMMPS) return (const void*)step_kernel<L, R, M, SC_TAGS>;
  else return (const void*)step_kernel<L, R, M, SC_SLOT>;
}
#define PSB_STEP_ROW(L, R, M)                                                                     \
  {(const void*)step_kernel<L, R, M, SC_POINT>, (const void*)step_kernel<L, R, M, SC_GENERAL>,    \
   (const void*)step_kernel<L, R, M, SC_SLOT>, stream_entry<L, R, M>(), tags_entry<L, R, M>()}
#define PSB_DEFINE_STEP_VTABLE(NAME, L, R)                                                          \
  static const void* const NAME##_entry[SM_COUNT][SC_COUNT] = {                                            \
      PSB_STEP_ROW(L, R, SM_UNBIASED), PSB_STEP_ROW(L, R, SM_HARMONIC), PSB_STEP_ROW(L, R, SM_METAD_DIRECT), \
      PSB_STEP_ROW(L, R, SM_ABF), PSB_STEP_ROW(L, R, SM_METAD_GRID)};                               \
  static cudaError_t NAME##_launch(int method, int general, const Model* M, const Snap* S, int grid, \
                                   size_t smem, cudaStream_t stream, int pdl) {                     \
    void* args[] = {(void*)&M, (void*)S}; /* M: device pointer, S: by value */                     \
    const void* fn = NAME##_entry[method][general];                                         \
    /* programmatic dependent launch: the next step's launch + model copy overlap this step's tail   \
       (griddepcontrol.* in step_kernel); pdl == 0: plain stream order */                            \
    if (pdl) {                                                                                      \
      cudaLaunchConfig_t cfg = {};                                                                  \
      cfg.gridDim = dim3(grid); cfg.blockDim = dim3(BLOCK); cfg.dynamicSmemBytes = smem; cfg.stream = stream; \
      cudaLaunchAttribute at[2];                                                                    \
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                \
      at[0].val.programmaticStreamSerializationAllowed = 1;                                         \
      at[1].id = cudaLaunchAttributeCooperative;                                                    \
      at[1].val.cooperative = grid > 1 ? 1 : 0;                                                     \
      cfg.attrs = at; cfg.numAttrs = 2;                                                             \
      return cudaLaunchKernelExC(&cfg, fn, args);                                                   \
    }                                                                                               \
    if (grid > 1) return cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(BLOCK), args, smem, stream); \
    return cudaLaunchKernel(fn, dim3(1), dim3(BLOCK), args, smem, stream);                          \
  }                                                                                                 \
  static int NAME##_occupancy(int method, int general, size_t smem) {                               \
    int occ = 0;                                                                                    \
    const void* fn = NAME##_entry[method][general];                                         \
    if (smem > 0 &&                                                                                 \
        cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) \
      return 0;                                                                                     \
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, BLOCK, smem) != cudaSuccess) return 0; \
    return occ;                                                                                     \
  }                                                                                                 \
  static const void* const NAME##_bentry[SM_COUNT][2] = {                                           \
      {(const void*)step_kernel_batch<L, R, SM_UNBIASED, SC_POINT>, (const void*)step_kernel_batch<L, R, SM_UNBIASED, SC_GENERAL>}, \
      {(const void*)step_kernel_batch<L, R, SM_HARMONIC, SC_POINT>, (const void*)step_kernel_batch<L, R, SM_HARMONIC, SC_GENERAL>}, \
      {(const void*)step_kernel_batch<L, R, SM_METAD_DIRECT, SC_POINT>, (const void*)step_kernel_batch<L, R, SM_METAD_DIRECT, SC_GENERAL>}, \
      {(const void*)step_kernel_batch<L, R, SM_ABF, SC_POINT>, (const void*)step_kernel_batch<L, R, SM_ABF, SC_GENERAL>}, \
      {(const void*)step_kernel_batch<L, R, SM_METAD_GRID, SC_POINT>, (const void*)step_kernel_batch<L, R, SM_METAD_GRID, SC_GENERAL>}}; \
  static cudaError_t NAME##_launch_batch(int method, int general, const BatchItem* items, const int32_t* cta_item, \
                                         int grid, size_t smem, cudaStream_t stream, int pdl) {     \
    if (general > SC_GENERAL) return cudaErrorInvalidValue;                                         \
    void* args[] = {(void*)&items, (void*)&cta_item};                                               \
    cudaLaunchConfig_t cfg = {};                                                                    \
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(BLOCK); cfg.dynamicSmemBytes = smem; cfg.stream = stream; \
    cudaLaunchAttribute at[2];                                                                      \
    at[0].id = cudaLaunchAttributeCooperative;                                                      \
    at[0].val.cooperative = 1;                                                                      \
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                  \
    at[1].val.programmaticStreamSerializationAllowed = 1;                                           \
    cfg.attrs = at; cfg.numAttrs = pdl ? 2 : 1;                                                     \
    return cudaLaunchKernelExC(&cfg, NAME##_bentry[method][general], args);                         \
  }                                                                                                 \
  static int NAME##_batch_occupancy(int method, int general, size_t smem) {                         \
    int occ = 0;                                                                                    \
    if (general > SC_GENERAL) return 0;                                                             \
    const void* fn = NAME##_bentry[method][general];                                                \
    if (smem > 0 &&                                                                                 \
        cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) \
      return 0;                                                                                     \
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, BLOCK, smem) != cudaSuccess) return 0; \
    return occ;                                                                                     \
  }                                                                                                 \
  static size_t NAME##_static_smem(int method, int general) {                                       \
    cudaFuncAttributes at{};                                                                        \
    if (cudaFuncGetAttributes(&at, NAME##_entry[method][general]) != cudaSuccess) return (size_t)-1; \
    return at.sharedSizeBytes;                                                                      \
  }                                                                                                 \
  static void NAME##_gather(const Model* M, const Snap* S, int cv, double* xa, int na_pad,          \
                            double* xb, int nb_pad, cudaStream_t stream) {                          \
    const int total = na_pad + nb_pad;                                                              \
    coord_gather_kernel<L, R><<<(total + 255) / 256, 256, 0, stream>>>(M, *S, cv, xa, na_pad, xb,   \
                                                                       nb_pad);                     \
  }                                                                                                 \
  static void NAME##_ermsd(const Model* M, const Snap* S, int cv, cudaStream_t stream) {           \
    ermsd_kernel<L, R><<<1, ERMSD_BLOCK, 0, stream>>>(M, *S, cv);                                   \
  }                                                                                                 \
  const StepVTable NAME = {NAME##_launch, NAME##_occupancy, NAME##_static_smem, NAME##_launch_batch,   \
                           NAME##_batch_occupancy, NAME##_gather, NAME##_ermsd};

}  // namespace psb
