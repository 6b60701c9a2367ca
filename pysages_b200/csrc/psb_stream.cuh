// psb_stream.cuh -- bulk-copy (TMA) helpers of the streaming variant of the slot-ordered step (class SC_STREAM, sm_100a).
//
// The slot-ordered class (psb_step.cuh) reads positions / images / velocities and read-modify-writes the force
// rows of (almost) every slot: a pure stream over the engine's arrays.  SC_STREAM moves that stream with 1-D bulk
// copies (cp.async.bulk -> SASS UBLKCP) instead of register loads: two CTAs per SM, every WARP owns a ring of
// shared-memory stages (STREAM_MAX_STAGES x stage_bytes, psb_device.cuh: StreamGeom) and an mbarrier per stage
// (StepShared::full), gather chunks [positions | images | velocities] first, scatter chunks [forces] behind them,
// written back with bulk stores (cp.async.bulk.global.shared::cta) once updated in shared memory.  Stage i completes
// on full[i % S] with parity (i / S) & 1; a stage is re-armed after the warp has consumed it (for scatter stages once
// the bulk store has read it: cp.async.bulk.wait_group.read).  HOOMD layout ((N,4) rows of Real, (N,3) int32 images).
//
// Measured slower than the register path on B200 (30.7 vs 27.4 us at 1e6 rows, DESIGN.md 4.1: the slot-map words and
// the per-pass issue latency bound the step, not the bytes in flight), so the class is opt-in (PSB_STREAM=1).
#pragma once

#include "psb_device.cuh"

namespace psb {

__device__ __forceinline__ void tma_bulk_s2g(void* dst, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               :: "l"(dst), "r"((uint32_t)__cvta_generic_to_shared(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void cp_async_4(void* dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;"
               :: "r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

}  // namespace psb
