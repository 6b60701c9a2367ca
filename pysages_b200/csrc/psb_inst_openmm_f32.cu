// fused step kernel + coordination gather instantiated for one engine layout / dtype
#include "psb_launch.cuh"
#include "psb_pairs.cuh"
#include "psb_step.cuh"
namespace psb {
PSB_DEFINE_STEP_VTABLE(vt_openmm_f32, PSB_LAYOUT_OPENMM_CUDA, float)
}  // namespace psb
