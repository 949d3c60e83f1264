// Host-side helpers shared by every translation unit of libvmm_sm100.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/vmm.h"

namespace vmm {

// error text of the last failing call on this thread (vmm_last_error)
char* error_buffer();
int set_error(int code, const char* msg);
int set_cuda_error(cudaError_t e, const char* where);
int check_launch(const char* where);
void count_launch();
int num_sms();
bool pdl_enabled();   // VMM_PDL != "0": programmatic dependent launch for the persistent GEMM kernels (see launch_maybe_pdl)

// Programmatic dependent launch (sm_90+).  A kernel launched with the attribute may become resident while the kernel before it in
// the stream is still running; it must execute pdl_wait() before it touches anything that kernel writes.  Everything before the wait
// (barrier initialisation, TMEM allocation, descriptor prefetch) then overlaps the tail of the predecessor.  pdl_trigger() at the
// top of a kernel lets its successor's CTAs be scheduled as soon as all of this kernel's CTAs have started (no-op otherwise).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_maybe_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// temporal attention for frame counts other than 11 (tattn_generic.cu), reached through vmm_tattn_fwd / vmm_tattn_bwd
int tattn_generic_fwd(const void* qkv, const float* ekv, const float* bias, const float* rot, void* out, int fmt, int B, int frames, int HW,
                      float scale, int pre_rotated, cudaStream_t stream);
int tattn_generic_bwd(const void* qkv, const float* ekv, const float* bias, const float* rot, const void* dout, void* dqkv, float* dekv,
                      float* dbias, int fmt, int B, int frames, int HW, float scale, int pre_rotated, cudaStream_t stream);

// cuTensorMapEncodeTiled through cudaGetDriverEntryPoint (no link-time libcuda dependency, so the
// library also loads on a machine without a driver).  128-byte swizzle (or 64), zero fill out of bounds.
int encode_tensor_map(CUtensorMap* map, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* gdim,
                      const uint64_t* gstride_bytes, const uint32_t* box, bool l2_256, int swizzle_bytes = 128);

static inline unsigned int ceil_div(long long a, long long b) { return static_cast<unsigned int>((a + b - 1) / b); }
static inline long long min64(long long a, long long b) { return a < b ? a : b; }

}  // namespace vmm
