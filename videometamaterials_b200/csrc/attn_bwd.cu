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
// linear attention backward
//  kernel 1: dctx[bf][h][d][e] = sum_n qs[n,d] dout[n,e]   (qs = softmax_d(q) * scale)        one CTA per (h, bf)
//  kernel 2: per (head, pixel): dq, dk, dv; block 0 also handles the T cond tokens (atomics into dekv)
// ------------------------------------------------------------------------------------------------
constexpr int LB_CHUNK = 128;

__global__ void __launch_bounds__(256) lattn_dctx_kernel(const uint16_t* __restrict__ qkv, const uint16_t* __restrict__ dout,
                                                         float* __restrict__ dctx, int fmt, int HW, int heads, float scale) {
  __shared__ float qs[LB_CHUNK][BDH + 1];
  __shared__ float ds[LB_CHUNK][BDH];
  const int h = blockIdx.x, bf = blockIdx.y;
  const int HD = heads * BDH;
  const int tid = threadIdx.x;
  const int d = tid >> 3, e0 = (tid & 7) * 4;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int n0 = 0; n0 < HW; n0 += LB_CHUNK) {
    const int cnt = min(LB_CHUNK, HW - n0);
    if (tid < LB_CHUNK) {
      float q[BDH];
      if (tid < cnt) {
        const uint16_t* row = qkv + (static_cast<long long>(bf) * HW + n0 + tid) * 3 * HD + h * BDH;
#pragma unroll
        for (int k = 0; k < 4; ++k) bld8f(row + k * 8, fmt, q + k * 8);
        float mx = q[0];
#pragma unroll
        for (int k = 1; k < BDH; ++k) mx = fmaxf(mx, q[k]);
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < BDH; ++k) {
          q[k] = __expf(q[k] - mx);
          sum += q[k];
        }
        const float inv = scale / sum;
#pragma unroll
        for (int k = 0; k < BDH; ++k) qs[tid][k] = q[k] * inv;
      } else {
#pragma unroll
        for (int k = 0; k < BDH; ++k) qs[tid][k] = 0.f;
      }
    } else {
      const int r = tid - LB_CHUNK;
      float v[BDH];
      if (r < cnt) {
        const uint16_t* row = dout + (static_cast<long long>(bf) * HW + n0 + r) * HD + h * BDH;
#pragma unroll
        for (int k = 0; k < 4; ++k) bld8f(row + k * 8, fmt, v + k * 8);
      } else {
#pragma unroll
        for (int k = 0; k < BDH; ++k) v[k] = 0.f;
      }
#pragma unroll
      for (int k = 0; k < BDH; ++k) ds[r][k] = v[k];
    }
    __syncthreads();
    for (int r = 0; r < cnt; ++r) {
      const float w = qs[r][d];
      const float4 vv = *reinterpret_cast<const float4*>(&ds[r][e0]);
      acc[0] += w * vv.x;
      acc[1] += w * vv.y;
      acc[2] += w * vv.z;
      acc[3] += w * vv.w;
    }
    __syncthreads();
  }
  float* c = dctx + ((static_cast<long long>(bf) * heads + h) * BDH + d) * BDH + e0;
  *reinterpret_cast<float4*>(c) = make_float4(acc[0], acc[1], acc[2], acc[3]);
}

__global__ void __launch_bounds__(256) lattn_bwd_apply_kernel(const uint16_t* __restrict__ qkv, const float* __restrict__ ekv, int T,
                                                              const uint16_t* __restrict__ dout, const float* __restrict__ ctx,
                                                              const float* __restrict__ dctx, const float* __restrict__ kstat,
                                                              uint16_t* __restrict__ dqkv, float* __restrict__ dekv, int fmt, int HW,
                                                              int heads, int frames, float scale, float vscale) {
  extern __shared__ float sm[];
  const int HD = heads * BDH;
  float* Cs = sm;                          // [heads][32][32]  ctx
  float* Gs = Cs + heads * BDH * BDH;      // [heads][32][32]  dctx * vscale
  float* Ms = Gs + heads * BDH * BDH;      // [heads][32] max
  float* Zs = Ms + heads * BDH;            // [heads][32] 1/Z
  float* Cc = Zs + heads * BDH;            // [heads][32] c[d] = sum_e dctx[d,e] ctx[d,e]
  const int bf = blockIdx.y;
  const int b = bf / frames;
  for (int i = threadIdx.x; i < heads * BDH * BDH; i += blockDim.x) {
    Cs[i] = ctx[static_cast<long long>(bf) * heads * BDH * BDH + i];
    Gs[i] = dctx[static_cast<long long>(bf) * heads * BDH * BDH + i] * vscale;
  }
  for (int i = threadIdx.x; i < heads * BDH; i += blockDim.x) {
    Ms[i] = kstat[(static_cast<long long>(bf) * heads * BDH + i) * 2];
    Zs[i] = 1.f / kstat[(static_cast<long long>(bf) * heads * BDH + i) * 2 + 1];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < heads * BDH; i += blockDim.x) {
    float a = 0.f;
    for (int e = 0; e < BDH; ++e) a += Gs[i * BDH + e] * Cs[i * BDH + e];
    Cc[i] = a / vscale * 1.f;   // sum_e dctx[d,e] ctx[d,e] (ctx already carries vscale once)
  }
  __syncthreads();
  const int ppb = blockDim.x / heads;
  const int h = threadIdx.x / ppb;
  const int lp = threadIdx.x % ppb;
  const float* ch = Cs + h * BDH * BDH;
  const float* gh = Gs + h * BDH * BDH;
  // ---- cond tokens (first block of each frame-image): m = token j, summed over frames by atomics
  if (blockIdx.x == 0 && lp < T) {
    const int j = lp;
    const float* src = ekv + (static_cast<long long>(b) * T + j) * 2 * HD + h * BDH;
    float w[BDH], dk[BDH], dv[BDH];
#pragma unroll
    for (int d = 0; d < BDH; ++d) w[d] = __expf(src[d] - Ms[h * BDH + d]) * Zs[h * BDH + d];
#pragma unroll
    for (int e = 0; e < BDH; ++e) dv[e] = 0.f;
#pragma unroll
    for (int d = 0; d < BDH; ++d) {
      float dw = 0.f;
#pragma unroll
      for (int e = 0; e < BDH; ++e) {
        dv[e] += w[d] * gh[d * BDH + e];
        dw += gh[d * BDH + e] * src[HD + e];
      }
      dk[d] = w[d] * (dw - Cc[h * BDH + d]);
    }
    float* dst = dekv + (static_cast<long long>(b) * T + j) * 2 * HD + h * BDH;
#pragma unroll
    for (int d = 0; d < BDH; ++d) {
      atomicAdd(dst + d, dk[d]);
      atomicAdd(dst + HD + d, dv[d]);
    }
  }
  const int n = blockIdx.x * ppb + lp;
  if (n >= HW) return;
  const uint16_t* row = qkv + (static_cast<long long>(bf) * HW + n) * 3 * HD + h * BDH;
  const uint16_t* drow = dout + (static_cast<long long>(bf) * HW + n) * HD + h * BDH;
  uint16_t* orow = dqkv + (static_cast<long long>(bf) * HW + n) * 3 * HD + h * BDH;
  float dO[BDH];
#pragma unroll
  for (int k = 0; k < 4; ++k) bld8f(drow + k * 8, fmt, dO + k * 8);
  {  // dq
    float q[BDH];
#pragma unroll
    for (int k = 0; k < 4; ++k) bld8f(row + k * 8, fmt, q + k * 8);
    float mx = q[0];
#pragma unroll
    for (int k = 1; k < BDH; ++k) mx = fmaxf(mx, q[k]);
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < BDH; ++k) {
      q[k] = __expf(q[k] - mx);
      sum += q[k];
    }
    const float inv = 1.f / sum;
    float g[BDH];
    float dot = 0.f;
#pragma unroll
    for (int d = 0; d < BDH; ++d) {
      float a = 0.f;
#pragma unroll
      for (int e = 0; e < BDH; e += 4) {
        const float4 c4 = *reinterpret_cast<const float4*>(ch + d * BDH + e);
        a += c4.x * dO[e] + c4.y * dO[e + 1] + c4.z * dO[e + 2] + c4.w * dO[e + 3];
      }
      q[d] *= inv;           // p[d]
      g[d] = a;              // d qs[d]
      dot += q[d] * a;
    }
#pragma unroll
    for (int d = 0; d < BDH; ++d) g[d] = q[d] * scale * (g[d] - dot);
#pragma unroll
    for (int k = 0; k < 4; ++k) bst8f(orow + k * 8, fmt, g + k * 8);
  }
  {  // dk, dv
    float kk[BDH], vv[BDH], dv[BDH], dk[BDH];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      bld8f(row + HD + k * 8, fmt, kk + k * 8);
      bld8f(row + 2 * HD + k * 8, fmt, vv + k * 8);
    }
#pragma unroll
    for (int e = 0; e < BDH; ++e) dv[e] = 0.f;
#pragma unroll
    for (int d = 0; d < BDH; ++d) {
      const float w = __expf(kk[d] - Ms[h * BDH + d]) * Zs[h * BDH + d];
      float dw = 0.f;
#pragma unroll
      for (int e = 0; e < BDH; e += 4) {
        const float4 g4 = *reinterpret_cast<const float4*>(gh + d * BDH + e);
        dv[e] += w * g4.x;
        dv[e + 1] += w * g4.y;
        dv[e + 2] += w * g4.z;
        dv[e + 3] += w * g4.w;
        dw += g4.x * vv[e] + g4.y * vv[e + 1] + g4.z * vv[e + 2] + g4.w * vv[e + 3];
      }
      dk[d] = w * (dw - Cc[h * BDH + d]);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      bst8f(orow + HD + k * 8, fmt, dk + k * 8);
      bst8f(orow + 2 * HD + k * 8, fmt, dv + k * 8);
    }
  }
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

extern "C" int vmm_lattn_bwd(const void* qkv, const float* ekv, int T, const void* dout, const float* ctx, const float* kstat, float* dctx,
                             void* dqkv, float* dekv, int fmt, int BF, int frames, int HW, int heads, float scale, float vscale,
                             void* stream_) {
  if (!qkv || !ekv || !dout || !ctx || !kstat || !dctx || !dqkv || !dekv) return set_error(VMM_ERR_ARG, "vmm_lattn_bwd: null pointer");
  if (heads < 1 || heads > 8 || (256 % heads) != 0 || T > 256 / heads) return set_error(VMM_ERR_UNSUPPORTED, "vmm_lattn_bwd: heads / T");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  lattn_dctx_kernel<<<dim3(heads, BF), 256, 0, stream>>>(static_cast<const uint16_t*>(qkv), static_cast<const uint16_t*>(dout), dctx, fmt, HW,
                                                         heads, scale);
  count_launch();
  const int ppb = 256 / heads;
  const size_t smem = (static_cast<size_t>(2) * heads * BDH * BDH + 3 * heads * BDH) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(lattn_bwd_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) return set_cuda_error(e, "vmm_lattn_bwd: attr");
    attr = true;
  }
  lattn_bwd_apply_kernel<<<dim3((HW + ppb - 1) / ppb, BF), 256, smem, stream>>>(
      static_cast<const uint16_t*>(qkv), ekv, T, static_cast<const uint16_t*>(dout), ctx, dctx, kstat, static_cast<uint16_t*>(dqkv), dekv,
      fmt, HW, heads, frames, scale, vscale);
  count_launch();
  return check_launch("vmm_lattn_bwd");
}

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
