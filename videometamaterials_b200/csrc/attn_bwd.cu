// Attention cores (backward).  Scores / probabilities are recomputed from the saved qkv rows (nothing of size
// tokens x tokens is kept from the forward).  Gradients with respect to the conditioning keys/values and the
// relative position bias are reduced over pixels in registers / shared memory and added atomically.
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace vmm {

constexpr int BDH = 32;
constexpr int BDHP = 36;

__device__ __forceinline__ void bld8f(const uint16_t* p, int fmt, float* v) {
  const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = unpack2_h16(w[j], fmt);
    v[2 * j] = f.x;
    v[2 * j + 1] = f.y;
  }
}
__device__ __forceinline__ void bst8f(uint16_t* p, int fmt, const float* v) {
  uint4 q;
  q.x = pack2_h16(v[0], v[1], fmt);
  q.y = pack2_h16(v[2], v[3], fmt);
  q.z = pack2_h16(v[4], v[5], fmt);
  q.w = pack2_h16(v[6], v[7], fmt);
  *reinterpret_cast<uint4*>(p) = q;
}

// ------------------------------------------------------------------------------------------------
// quadratic spatial attention backward (bottleneck).  CTA = (head, frame-image).
// phase 1 (thread = query): dq.   phase 2 (thread = key): dk, dv; key 0 is the frame's cond token.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sattn_bwd_kernel(const uint16_t* __restrict__ qkv, const float* __restrict__ ekv,
                                                        const uint16_t* __restrict__ aout, const uint16_t* __restrict__ dout,
                                                        const float* __restrict__ lse, uint16_t* __restrict__ dqkv,
                                                        float* __restrict__ dekv, int fmt, int HW, int heads, float scale) {
  extern __shared__ float sm[];
  const int NK = HW + 1;
  float* Ks = sm;                 // [NK][32]
  float* Vs = Ks + NK * BDH;      // [NK][32]
  float* Qs = Vs + NK * BDH;      // [HW][32]  scaled q
  float* Ds = Qs + HW * BDH;      // [HW][32]  dO
  float* Ls = Ds + HW * BDH;      // [HW] lse
  float* Dd = Ls + HW;            // [HW] rowsum(dO * O)
  const int h = blockIdx.x, bf = blockIdx.y;
  const int HD = heads * BDH;
  for (int i = threadIdx.x; i < NK * 8; i += blockDim.x) {
    const int r = i >> 3, part = i & 7;
    float v8[8];
    if (r == 0) {
      const float* src = ekv + static_cast<long long>(bf) * 2 * HD + (part < 4 ? 0 : HD) + h * BDH + (part & 3) * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) v8[j] = src[j];
    } else {
      const uint16_t* row = qkv + (static_cast<long long>(bf) * HW + (r - 1)) * 3 * HD + (part < 4 ? HD : 2 * HD) + h * BDH + (part & 3) * 8;
      bld8f(row, fmt, v8);
    }
    float* dst = (part < 4 ? Ks : Vs) + r * BDH + (part & 3) * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) dst[j] = v8[j];
  }
  for (int n = threadIdx.x; n < HW; n += blockDim.x) {
    const long long rown = static_cast<long long>(bf) * HW + n;
    float q[BDH], dO[BDH], o[BDH];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      bld8f(qkv + rown * 3 * HD + h * BDH + k * 8, fmt, q + k * 8);
      bld8f(dout + rown * HD + h * BDH + k * 8, fmt, dO + k * 8);
      bld8f(aout + rown * HD + h * BDH + k * 8, fmt, o + k * 8);
    }
    float dd = 0.f;
#pragma unroll
    for (int k = 0; k < BDH; ++k) {
      Qs[n * BDH + k] = q[k] * scale;
      Ds[n * BDH + k] = dO[k];
      dd += dO[k] * o[k];
    }
    Dd[n] = dd;
    Ls[n] = lse[(static_cast<long long>(bf) * heads + h) * HW + n];
  }
  __syncthreads();
  // phase 1: queries
  for (int n = threadIdx.x; n < HW; n += blockDim.x) {
    float q[BDH], dO[BDH], dq[BDH];
#pragma unroll
    for (int k = 0; k < BDH; ++k) {
      q[k] = Qs[n * BDH + k];
      dO[k] = Ds[n * BDH + k];
      dq[k] = 0.f;
    }
    const float l = Ls[n], dd = Dd[n];
    for (int j = 0; j < NK; ++j) {
      const float* kr = Ks + j * BDH;
      const float* vr = Vs + j * BDH;
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int k = 0; k < BDH; k += 4) {
        const float4 kk = *reinterpret_cast<const float4*>(kr + k);
        const float4 vv = *reinterpret_cast<const float4*>(vr + k);
        s += q[k] * kk.x + q[k + 1] * kk.y + q[k + 2] * kk.z + q[k + 3] * kk.w;
        dp += dO[k] * vv.x + dO[k + 1] * vv.y + dO[k + 2] * vv.z + dO[k + 3] * vv.w;
      }
      const float dsv = __expf(s - l) * (dp - dd);
#pragma unroll
      for (int k = 0; k < BDH; k += 4) {
        const float4 kk = *reinterpret_cast<const float4*>(kr + k);
        dq[k] += dsv * kk.x;
        dq[k + 1] += dsv * kk.y;
        dq[k + 2] += dsv * kk.z;
        dq[k + 3] += dsv * kk.w;
      }
    }
#pragma unroll
    for (int k = 0; k < BDH; ++k) dq[k] *= scale;
    uint16_t* orow = dqkv + (static_cast<long long>(bf) * HW + n) * 3 * HD + h * BDH;
#pragma unroll
    for (int k = 0; k < 4; ++k) bst8f(orow + k * 8, fmt, dq + k * 8);
  }
  // phase 2: keys
  for (int j = threadIdx.x; j < NK; j += blockDim.x) {
    float kk[BDH], vv[BDH], dk[BDH], dv[BDH];
#pragma unroll
    for (int k = 0; k < BDH; ++k) {
      kk[k] = Ks[j * BDH + k];
      vv[k] = Vs[j * BDH + k];
      dk[k] = dv[k] = 0.f;
    }
    for (int n = 0; n < HW; ++n) {
      const float* qr = Qs + n * BDH;
      const float* dr = Ds + n * BDH;
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int k = 0; k < BDH; k += 4) {
        const float4 qq = *reinterpret_cast<const float4*>(qr + k);
        const float4 d4 = *reinterpret_cast<const float4*>(dr + k);
        s += qq.x * kk[k] + qq.y * kk[k + 1] + qq.z * kk[k + 2] + qq.w * kk[k + 3];
        dp += d4.x * vv[k] + d4.y * vv[k + 1] + d4.z * vv[k + 2] + d4.w * vv[k + 3];
      }
      const float pr = __expf(s - Ls[n]);
      const float dsv = pr * (dp - Dd[n]);
#pragma unroll
      for (int k = 0; k < BDH; k += 4) {
        const float4 qq = *reinterpret_cast<const float4*>(qr + k);
        const float4 d4 = *reinterpret_cast<const float4*>(dr + k);
        dk[k] += dsv * qq.x;
        dk[k + 1] += dsv * qq.y;
        dk[k + 2] += dsv * qq.z;
        dk[k + 3] += dsv * qq.w;
        dv[k] += pr * d4.x;
        dv[k + 1] += pr * d4.y;
        dv[k + 2] += pr * d4.z;
        dv[k + 3] += pr * d4.w;
      }
    }
    if (j == 0) {
      float* dst = dekv + static_cast<long long>(bf) * 2 * HD + h * BDH;
#pragma unroll
      for (int k = 0; k < BDH; ++k) {
        dst[k] = dk[k];
        dst[HD + k] = dv[k];
      }
    } else {
      uint16_t* orow = dqkv + (static_cast<long long>(bf) * HW + (j - 1)) * 3 * HD + h * BDH;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        bst8f(orow + HD + k * 8, fmt, dk + k * 8);
        bst8f(orow + 2 * HD + k * 8, fmt, dv + k * 8);
      }
    }
  }
}

}  // namespace vmm

using namespace vmm;

extern "C" int vmm_sattn_bwd(const void* qkv, const float* ekv, const void* aout, const void* dout, const float* lse, void* dqkv,
                             float* dekv, int fmt, int BF, int HW, int heads, float scale, void* stream_) {
  if (!qkv || !ekv || !aout || !dout || !lse || !dqkv || !dekv) return set_error(VMM_ERR_ARG, "vmm_sattn_bwd: null pointer");
  const size_t smem = (static_cast<size_t>(2) * (HW + 1) * BDH + static_cast<size_t>(2) * HW * BDH + 2 * HW) * sizeof(float);
  if (smem > 220 * 1024) return set_error(VMM_ERR_UNSUPPORTED, "vmm_sattn_bwd: too many keys for one CTA");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(sattn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return set_cuda_error(e, "vmm_sattn_bwd: attr");
    attr = true;
  }
  int threads = HW + 1 < 256 ? (HW + 1 + 31) / 32 * 32 : 256;
  sattn_bwd_kernel<<<dim3(heads, BF), threads, smem, stream>>>(static_cast<const uint16_t*>(qkv), ekv, static_cast<const uint16_t*>(aout),
                                                              static_cast<const uint16_t*>(dout), lse, static_cast<uint16_t*>(dqkv), dekv, fmt,
                                                              HW, heads, scale);
  count_launch();
  return check_launch("vmm_sattn_bwd");
}
