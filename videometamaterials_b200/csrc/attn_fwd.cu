// Attention cores (forward): the quadratic spatial attention of the bottleneck (temporal attention: tattn_mma.cu,
// linear attention: lattn_mma.cu).  Projections (to_qkv / to_out) run on vmm_cgemm; these
// kernels take the packed qkv rows [position, 3*heads*32] (q | k | v, each split (head, 32)) and write the
// attention output rows [position, heads*32].  dim_head is fixed at 32 (model.yaml: unet_attn_dim_head).
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace vmm {

constexpr int DH = 32;    // dim_head
constexpr int DHP = 36;   // padded head stride in shared memory (floats): keeps float4 reads of different heads on different banks

__device__ __forceinline__ void ld8f(const uint16_t* p, int fmt, float* v) {
  const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = unpack2_h16(w[j], fmt);
    v[2 * j] = f.x;
    v[2 * j + 1] = f.y;
  }
}
__device__ __forceinline__ void st8f(uint16_t* p, int fmt, const float* v) {
  uint4 q;
  q.x = pack2_h16(v[0], v[1], fmt);
  q.y = pack2_h16(v[2], v[3], fmt);
  q.z = pack2_h16(v[4], v[5], fmt);
  q.w = pack2_h16(v[6], v[7], fmt);
  *reinterpret_cast<uint4*>(p) = q;
}

// ------------------------------------------------------------------------------------------------
// Quadratic spatial attention of the bottleneck, VDDP:687-689: per frame-image, HW queries x (1 cond token +
// HW) keys, no rotary, no bias.  CTA = (head, frame-image); K / V of that head live in shared memory; one
// query per thread with an online softmax.  Also writes the log-sum-exp per query for the backward pass.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sattn_fwd_kernel(const uint16_t* __restrict__ qkv, const float* __restrict__ ekv,
                                                        uint16_t* __restrict__ out, float* __restrict__ lse, int fmt, int HW,
                                                        int heads, int frames, float scale) {
  extern __shared__ float sm[];
  const int NK = HW + 1;
  float* Ks = sm;               // [NK][DH]
  float* Vs = sm + NK * DH;     // [NK][DH]
  const int h = blockIdx.x, bf = blockIdx.y;
  const int HD = heads * DH;
  // key 0 = cond token of this frame: ekv[b][f] == ekv[bf]
  for (int i = threadIdx.x; i < NK * 8; i += blockDim.x) {
    const int r = i >> 3, part = i & 7;
    float v8[8];
    if (r == 0) {
      const float* src = ekv + static_cast<long long>(bf) * 2 * HD + (part < 4 ? 0 : HD) + h * DH + (part & 3) * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) v8[j] = src[j];
    } else {
      const uint16_t* row = qkv + (static_cast<long long>(bf) * HW + (r - 1)) * 3 * HD + (part < 4 ? HD : 2 * HD) + h * DH + (part & 3) * 8;
      ld8f(row, fmt, v8);
    }
    float* dst = (part < 4 ? Ks : Vs) + r * DH + (part & 3) * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) dst[j] = v8[j];
  }
  __syncthreads();
  for (int n = threadIdx.x; n < HW; n += blockDim.x) {
    const uint16_t* row = qkv + (static_cast<long long>(bf) * HW + n) * 3 * HD + h * DH;
    float q[DH];
#pragma unroll
    for (int k = 0; k < 4; ++k) ld8f(row + k * 8, fmt, q + k * 8);
#pragma unroll
    for (int k = 0; k < DH; ++k) q[k] *= scale;
    float m = -1e30f, l = 0.f;
    float o[DH];
#pragma unroll
    for (int k = 0; k < DH; ++k) o[k] = 0.f;
    for (int j = 0; j < NK; ++j) {
      const float* kr = Ks + j * DH;
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < DH; k += 4) {
        const float4 kk = *reinterpret_cast<const float4*>(kr + k);
        s += q[k] * kk.x + q[k + 1] * kk.y + q[k + 2] * kk.z + q[k + 3] * kk.w;
      }
      const float mn = fmaxf(m, s);
      const float cf = __expf(m - mn);
      const float pj = __expf(s - mn);
      l = l * cf + pj;
      const float* vr = Vs + j * DH;
#pragma unroll
      for (int k = 0; k < DH; k += 4) {
        const float4 vv = *reinterpret_cast<const float4*>(vr + k);
        o[k] = o[k] * cf + pj * vv.x;
        o[k + 1] = o[k + 1] * cf + pj * vv.y;
        o[k + 2] = o[k + 2] * cf + pj * vv.z;
        o[k + 3] = o[k + 3] * cf + pj * vv.w;
      }
      m = mn;
    }
    const float inv = 1.f / l;
#pragma unroll
    for (int k = 0; k < DH; ++k) o[k] *= inv;
    uint16_t* orow = out + (static_cast<long long>(bf) * HW + n) * HD + h * DH;
#pragma unroll
    for (int k = 0; k < 4; ++k) st8f(orow + k * 8, fmt, o + k * 8);
    if (lse) lse[(static_cast<long long>(bf) * heads + h) * HW + n] = m + __logf(l);
  }
}

}  // namespace vmm

using namespace vmm;

extern "C" int vmm_sattn_fwd(const void* qkv, const float* ekv, void* out, float* lse, int fmt, int BF, int frames, int HW,
                             int heads, float scale, void* stream_) {
  if (!qkv || !ekv || !out) return set_error(VMM_ERR_ARG, "vmm_sattn_fwd: null pointer");
  const size_t smem = static_cast<size_t>(2) * (HW + 1) * DH * sizeof(float);
  if (smem > 200 * 1024) return set_error(VMM_ERR_UNSUPPORTED, "vmm_sattn_fwd: too many keys for one CTA");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(sattn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return set_cuda_error(e, "vmm_sattn_fwd: attr");
    attr = true;
  }
  int threads = HW < 256 ? (HW + 31) / 32 * 32 : 256;
  sattn_fwd_kernel<<<dim3(heads, BF), threads, smem, stream>>>(static_cast<const uint16_t*>(qkv), ekv, static_cast<uint16_t*>(out), lse,
                                                              fmt, HW, heads, frames, scale);
  count_launch();
  return check_launch("vmm_sattn_fwd");
}
