// psb_stream.cuh -- TMA-pipelined slot-ordered streaming for the fused step kernel (class SC_STREAM, sm_100a).
//
// The slot-ordered class (psb_step.cuh) reads positions / images / velocities and read-modify-writes the force
// rows of (almost) every slot: a pure stream over the engine's arrays.  Loading it through registers leaves too
// few bytes in flight (two dependent batches of 8 rows per thread: 3.2 TB/s, ncu warps_active 25 %).  Here ONE
// CTA per SM owns a contiguous slice of slots and moves it with 1-D bulk copies (cp.async.bulk -> SASS UBLKCP)
// through a ring of shared-memory stages that fills the SM's whole shared memory (~200 KB in flight per SM):
//
//   item 0 .. n-1   gather chunk c : [positions | images | velocities] of CH slots  -> issued before the
//                                    dependency wait and consumed after pass 0 (their HBM time hides under it)
//   item n .. 2n-1  scatter chunk c: [forces (| positions | images: RoG groups)]    -> issued as gather stages
//                                    drain, i.e. the force rows arrive while the grid-wide exchange and the
//                                    finalizer run; updated in shared memory, written back with bulk stores
//
// Item i lives in stage i % S and completes on mbarrier full[i % S] with parity (i / S) & 1; item i + S is issued
// by thread 0 once item i has been consumed (block barrier; for scatter chunks once the bulk store has read the
// stage: cp.async.bulk.wait_group.read).  HOOMD layout ((N,4) rows of Real, (N,3) int32 images).
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
