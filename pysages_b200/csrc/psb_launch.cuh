// psb_launch.cuh -- host-callable launch tables; each psb_inst_*.cu instantiates the kernels
// for one engine layout / dtype and exports one table, so translation units build in parallel.
#pragma once

#include "psb_device.cuh"

namespace psb {

struct StepVTable {
  cudaError_t (*launch_step)(const Model* M, const Snap* S, int grid, size_t dyn_smem, cudaStream_t stream);
  int (*occupancy)(size_t dyn_smem);
  void (*launch_gather)(const Model* M, const Snap* S, int cv, double* xa, int na_pad, double* xb,
                        int nb_pad, cudaStream_t stream);
};

using PairLaunchFn = void (*)(int grid, cudaStream_t stream, const double* xa, int na_pad,
                              const double* xb, int nb_pad, int nA, int nB, int nrt, double inv_r02,
                              double* gA, double* gB, double* xi_part);
// R in {1, 4}; H = n / 2 in {1, 2, 3, 4, 6}; nullptr when unsupported
PairLaunchFn pair_launcher(int R, int H);

void launch_inverse_permutation(const int* ids, uint32_t* inv, long long n, int blocks, cudaStream_t stream);
void launch_grid_index(const psb_grid_desc& g, const double* x, long long n, uint32_t* out);

extern const StepVTable vt_hoomd_f32, vt_hoomd_f64, vt_openmm_f32, vt_openmm_f64, vt_lammps_f64,
    vt_plain3_f32, vt_plain3_f64;

#define PSB_DEFINE_STEP_VTABLE(NAME, L, R)                                                          \
  static cudaError_t NAME##_launch(const Model* M, const Snap* S, int grid, size_t smem,            \
                                   cudaStream_t stream) {                                           \
    void* args[] = {(void*)M, (void*)S};                                                            \
    if (grid > 1)                                                                                   \
      return cudaLaunchCooperativeKernel((const void*)step_kernel<L, R>, dim3(grid), dim3(BLOCK),   \
                                         args, smem, stream);                                       \
    return cudaLaunchKernel((const void*)step_kernel<L, R>, dim3(1), dim3(BLOCK), args, smem, stream); \
  }                                                                                                 \
  static int NAME##_occupancy(size_t smem) {                                                        \
    int occ = 0;                                                                                    \
    if (smem > 0 && cudaFuncSetAttribute(step_kernel<L, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         (int)smem) != cudaSuccess)                                 \
      return 0;                                                                                     \
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, step_kernel<L, R>, BLOCK, smem) !=      \
        cudaSuccess)                                                                                \
      return 0;                                                                                     \
    return occ;                                                                                     \
  }                                                                                                 \
  static void NAME##_gather(const Model* M, const Snap* S, int cv, double* xa, int na_pad,          \
                            double* xb, int nb_pad, cudaStream_t stream) {                          \
    const int total = na_pad + nb_pad;                                                              \
    coord_gather_kernel<L, R><<<(total + 255) / 256, 256, 0, stream>>>(*M, *S, cv, xa, na_pad, xb,  \
                                                                       nb_pad);                     \
  }                                                                                                 \
  const StepVTable NAME = {NAME##_launch, NAME##_occupancy, NAME##_gather};

}  // namespace psb
