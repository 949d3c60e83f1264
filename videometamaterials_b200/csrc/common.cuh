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
