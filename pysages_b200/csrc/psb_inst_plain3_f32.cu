// fused step kernel + coordination gather instantiated for one engine layout / dtype
#include "psb_launch.cuh"
#include "psb_pairs.cuh"
#include "psb_step.cuh"
namespace psb {
PSB_DEFINE_STEP_VTABLE(vt_plain3_f32, PSB_LAYOUT_PLAIN3, float)
}  // namespace psb
