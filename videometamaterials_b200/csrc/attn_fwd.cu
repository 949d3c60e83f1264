// Attention cores (forward): O(n) linear attention over pixels and the quadratic spatial attention of the
// bottleneck (the temporal attention lives in tattn_mma.cu).  Projections (to_qkv / to_out) run on vmm_cgemm; these
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
// Linear attention, VDDP:331-378 ('self-stacked', all T cond tokens prepended to every frame).
// Kernel 1 (per frame-image bf, head h): context[d][e] = sum_m softmax_m(k[:, d])[m] * v[m, e] / HW over the
// T cond tokens and the HW pixels, with a running (online) max so that one pass over the pixels suffices.
// Also stores the softmax statistics (max, sum) per d for the backward pass.
// ------------------------------------------------------------------------------------------------
constexpr int LA_CHUNK = 128;

__global__ void __launch_bounds__(256) lattn_ctx_kernel(const uint16_t* __restrict__ qkv, const float* __restrict__ ekv, int T,
                                                        float* __restrict__ ctx, float* __restrict__ kstat, int fmt, int HW,
                                                        int heads, int frames) {
  __shared__ float ks[LA_CHUNK][DH + 1];
  __shared__ float vs[LA_CHUNK][DH];
  __shared__ float red[8][DH];
  __shared__ float Mrun[DH], Mnew[DH], corr[DH];
  const int h = blockIdx.x, bf = blockIdx.y;
  const int b = bf / frames;
  const int HD = heads * DH;
  const int tid = threadIdx.x;
  const int d = tid >> 3;            // 0..31   row of the context this thread accumulates
  const int e0 = (tid & 7) * 4;      // 4 consecutive columns
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float z = 0.f;                     // sum of weights for row d (replicated in the 8 threads of a row)
  if (tid < DH) Mrun[tid] = -1e30f;
  __syncthreads();
  const int total = T + HW;          // tokens first, then pixels
  for (int m0 = 0; m0 < total; m0 += LA_CHUNK) {
    const int cnt = min(LA_CHUNK, total - m0);
    // stage k, v of this chunk as fp32
    for (int i = tid; i < LA_CHUNK * 8; i += 256) {
      const int r = i >> 3, part = i & 7;          // part: 0..3 -> k[8*part..], 4..7 -> v
      const int m = m0 + r;
      float v8[8];
      if (r < cnt) {
        if (m < T) {
          const float* src = ekv + (static_cast<long long>(b) * T + m) * 2 * HD + (part < 4 ? 0 : HD) + h * DH + (part & 3) * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) v8[j] = src[j];
        } else {
          const uint16_t* row = qkv + (static_cast<long long>(bf) * HW + (m - T)) * 3 * HD + (part < 4 ? HD : 2 * HD) + h * DH + (part & 3) * 8;
          ld8f(row, fmt, v8);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v8[j] = (part < 4) ? -1e30f : 0.f;
      }
      if (part < 4) {
#pragma unroll
        for (int j = 0; j < 8; ++j) ks[r][part * 8 + j] = v8[j];
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) vs[r][(part - 4) * 8 + j] = v8[j];
      }
    }
    __syncthreads();
    // column max of the chunk: 8 partial maxima per column
    {
      const int col = tid & 31, part = tid >> 5;
      float mx = -1e30f;
      for (int r = part; r < LA_CHUNK; r += 8) mx = fmaxf(mx, ks[r][col]);
      red[part][col] = mx;
    }
    __syncthreads();
    if (tid < DH) {
      float mx = Mrun[tid];
#pragma unroll
      for (int k = 0; k < 8; ++k) mx = fmaxf(mx, red[k][tid]);
      Mnew[tid] = mx;
      corr[tid] = __expf(Mrun[tid] - mx);
      Mrun[tid] = mx;
    }
    __syncthreads();
    // weights in place
    for (int i = tid; i < LA_CHUNK * DH; i += 256) {
      const int r = i >> 5, col = i & 31;
      ks[r][col] = (r < cnt) ? __expf(ks[r][col] - Mnew[col]) : 0.f;
    }
    const float cf = corr[d];
    acc[0] *= cf;
    acc[1] *= cf;
    acc[2] *= cf;
    acc[3] *= cf;
    z *= cf;
    __syncthreads();
    for (int r = 0; r < cnt; ++r) {
      const float w = ks[r][d];
      const float4 vv = *reinterpret_cast<const float4*>(&vs[r][e0]);
      acc[0] += w * vv.x;
      acc[1] += w * vv.y;
      acc[2] += w * vv.z;
      acc[3] += w * vv.w;
      z += w;
    }
    __syncthreads();
  }
  const float nrm = 1.f / (z * static_cast<float>(HW));
  float* c = ctx + ((static_cast<long long>(bf) * heads + h) * DH + d) * DH + e0;
  *reinterpret_cast<float4*>(c) = make_float4(acc[0] * nrm, acc[1] * nrm, acc[2] * nrm, acc[3] * nrm);
  if (kstat && (tid & 7) == 0) {
    kstat[((static_cast<long long>(bf) * heads + h) * DH + d) * 2] = Mrun[d];
    kstat[((static_cast<long long>(bf) * heads + h) * DH + d) * 2 + 1] = z;
  }
}

// Kernel 2: out[n, h*32+e] = sum_d context[d][e] * softmax_d(q[n, h, :])[d] * scale.   thread = (head, pixel)
__global__ void __launch_bounds__(256) lattn_out_kernel(const uint16_t* __restrict__ qkv, const float* __restrict__ ctx,
                                                        uint16_t* __restrict__ out, int fmt, int HW, int heads, float scale) {
  extern __shared__ float cs[];   // [heads][DH][DH]
  const int bf = blockIdx.y;
  const int HD = heads * DH;
  for (int i = threadIdx.x; i < heads * DH * DH; i += blockDim.x) cs[i] = ctx[static_cast<long long>(bf) * heads * DH * DH + i];
  __syncthreads();
  const int ppb = blockDim.x / heads;           // pixels per block (32 with 8 heads)
  const int h = threadIdx.x / ppb;
  const int n = blockIdx.x * ppb + (threadIdx.x % ppb);
  if (n >= HW) return;
  const uint16_t* row = qkv + (static_cast<long long>(bf) * HW + n) * 3 * HD + h * DH;
  float q[DH];
#pragma unroll
  for (int k = 0; k < 4; ++k) ld8f(row + k * 8, fmt, q + k * 8);
  float mx = q[0];
#pragma unroll
  for (int k = 1; k < DH; ++k) mx = fmaxf(mx, q[k]);
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < DH; ++k) {
    q[k] = __expf(q[k] - mx);
    sum += q[k];
  }
  const float inv = scale / sum;
  float o[DH];
#pragma unroll
  for (int k = 0; k < DH; ++k) o[k] = 0.f;
  const float* ch = cs + h * DH * DH;
#pragma unroll
  for (int dd = 0; dd < DH; ++dd) {
    const float w = q[dd] * inv;
#pragma unroll
    for (int e = 0; e < DH; e += 4) {
      const float4 c4 = *reinterpret_cast<const float4*>(ch + dd * DH + e);
      o[e] += w * c4.x;
      o[e + 1] += w * c4.y;
      o[e + 2] += w * c4.z;
      o[e + 3] += w * c4.w;
    }
  }
  uint16_t* orow = out + (static_cast<long long>(bf) * HW + n) * HD + h * DH;
#pragma unroll
  for (int k = 0; k < 4; ++k) st8f(orow + k * 8, fmt, o + k * 8);
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

extern "C" int vmm_lattn_fwd(const void* qkv, const float* ekv, int T, void* out, float* ctx, float* kstat, int fmt, int BF,
                             int frames, int HW, int heads, float scale, void* stream_) {
  if (!qkv || !ekv || !out || !ctx) return set_error(VMM_ERR_ARG, "vmm_lattn_fwd: null pointer");
  if (heads < 1 || heads > 8 || (256 % heads) != 0) return set_error(VMM_ERR_UNSUPPORTED, "vmm_lattn_fwd: heads");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  lattn_ctx_kernel<<<dim3(heads, BF), 256, 0, stream>>>(static_cast<const uint16_t*>(qkv), ekv, T, ctx, kstat, fmt, HW, heads, frames);
  count_launch();
  const int ppb = 256 / heads;
  lattn_out_kernel<<<dim3((HW + ppb - 1) / ppb, BF), 256, heads * DH * DH * sizeof(float), stream>>>(
      static_cast<const uint16_t*>(qkv), ctx, static_cast<uint16_t*>(out), fmt, HW, heads, scale);
  count_launch();
  return check_launch("vmm_lattn_fwd");
}

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
