// Programmatic dependent launch (griddepcontrol, sm_90+).  The tensor-core kernels are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization: when the kernel in front of them on the stream is one of the
// engine's own (the merge / mix / layout pass that produces an operand, or the previous contraction of the same layer)
// and has called pdl_trigger(), their CTAs may become resident while it drains — barrier init, TMEM allocation and
// tensor-map prefetch (0.9 us) and the launch latency (1.8 us, tools/gemm_trace.py) then overlap its tail.  Nothing in
// global memory is touched before pdl_wait(), which returns once the previous grid has completed and flushed.  Behind
// a kernel that never triggers (every ATen kernel) the launch is an ordinary serialized one.  LYCO_PDL=0 drops the
// launch attribute (the two instructions are then no-ops).
#pragma once

namespace lyco {

__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

}  // namespace lyco
