// Warp-level tensor-core helpers (mma.sync m16n8k16 + ldmatrix) for the attention cores.
// The attention matrices here are tiny (11 x 22 x 32 per pixel and head, 32 x 32 contexts): far below the
// 128-row granularity of tcgen05, so they run on the warp-synchronous MMA path with fragments in registers.
#pragma once
#include <stdint.h>

#include "sm100_ptx.cuh"

namespace vmm {

__device__ __forceinline__ void ldsm_x4(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

// transpose of an 8 x 8 matrix of 16-bit elements held one 32-bit register per lane (lane = 4 * row + column pair): the layout of an
// accumulator block packed to 16 bit, and of every 8 x 8 block of an A / B fragment
__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t a) {
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}

// D(16x8, fp32) += A(16x16, row) * B(16x8, col); FMT 0 = fp16, 1 = bf16 operands
template <int FMT>
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, const uint32_t* b) {
  if (FMT == 0) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  } else {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
}

template <int FMT>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  if (FMT == 0) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  } else {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
}
template <int FMT>
__device__ __forceinline__ float2 unpack2(uint32_t v) {
  if (FMT == 0) {
    return __half22float2(*reinterpret_cast<__half2*>(&v));
  } else {
    return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&v));
  }
}

}  // namespace vmm
