// GroupNorm (+ time scale/shift + SiLU) and channel LayerNorm, forward and backward.
// All tensors are channels-last [sample, pixel, C] 16-bit; statistics / parameters are fp32 or fp64.
// These kernels are HBM-bound: 16-byte vector accesses, one read (+ one write) of the activation.
#include "common.cuh"
#include "mma_sync.cuh"
#include "sm100_ptx.cuh"

#include <stdlib.h>

namespace vmm {

__device__ __forceinline__ float silu_f(float u) { return u / (1.f + __expf(-u)); }
__device__ __forceinline__ float dsilu_f(float u) {
  const float s = 1.f / (1.f + __expf(-u));
  return s * (1.f + u * (1.f - s));
}

__device__ __forceinline__ void load8(const uint16_t* p, int fmt, float* v) {
  const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = unpack2_h16(w[j], fmt);
    v[2 * j] = f.x;
    v[2 * j + 1] = f.y;
  }
}
__device__ __forceinline__ void store8(uint16_t* p, int fmt, const float* v) {
  uint4 q;
  q.x = pack2_h16(v[0], v[1], fmt);
  q.y = pack2_h16(v[2], v[3], fmt);
  q.z = pack2_h16(v[4], v[5], fmt);
  q.w = pack2_h16(v[6], v[7], fmt);
  *reinterpret_cast<uint4*>(p) = q;
}

// ------------------------------------------------------------------------------------------------
// GroupNorm apply:  y = silu( ((x - mean) * rstd * gamma + beta) * (scale + 1) + shift )     VDDP:279-285
// statistics come from the conv epilogue as fp64 (sum, sum of squares) per (sample, group).
//
// These kernels were ALU / MUFU bound, not HBM bound (per element: runtime-format conversions, 4-7 shared-memory
// coefficient loads, exp + reciprocal).  Now: the 16-bit format is a template parameter, every thread keeps the
// coefficients of its 8 channels in registers (the grid stride is a multiple of C, so they never change), and the sigmoid
// is one MUFU.TANH (sigmoid(u) = 0.5 tanh(0.5 u) + 0.5; relative error 2^-11, below the 2^-9 rounding of the bf16 output;
// the fp16 forward used for sampling keeps exp + division).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <int FMT>
__device__ __forceinline__ float sigmoid_fast(float u) {
  if (FMT == 1) return fmaf(0.5f, tanh_approx(0.5f * u), 0.5f);
  return 1.f / (1.f + __expf(-u));
}
// d silu(u) / du = s (1 + u (1 - s)),  s = sigmoid(u)
template <int FMT>
__device__ __forceinline__ float dsilu_fast(float u) {
  const float s = sigmoid_fast<FMT>(u);
  const float q = fmaf(-u, s, u);      // u (1 - s)
  return fmaf(s, q, s);
}

template <int FMT>
__device__ __forceinline__ void unpack8(const uint4& q, float* v) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = unpack2<FMT>(w[j]);
    v[2 * j] = f.x;
    v[2 * j + 1] = f.y;
  }
}
template <int FMT>
__device__ __forceinline__ uint4 pack8(const float* v) {
  return make_uint4(pack2<FMT>(v[0], v[1]), pack2<FMT>(v[2], v[3]), pack2<FMT>(v[4], v[5]), pack2<FMT>(v[6], v[7]));
}

// Per-thread coefficients of the 8 consecutive channels c0 .. c0 + 7 of sample b, straight from global memory into registers
// (no shared-memory staging, no block barrier: the loads overlap the first activation loads of the thread).
//   forward   u = x * a + d                    (mean / rstd / gamma / beta / scale / shift folded)
// GS8: the channels per group are a multiple of 8, so the 8 channels share one group and mean / rstd are one fp64 evaluation.
template <bool GS8>
struct GnCoef {
  float a[8], d[8], sc[8];
  float mean[GS8 ? 1 : 8], rstd[GS8 ? 1 : 8];
};

__device__ __forceinline__ void gn_group_stats(const double* __restrict__ stats, int b, int g, int groups, double n, float eps, float& mean, float& rstd) {
  const double2 s = *reinterpret_cast<const double2*>(stats + (static_cast<long long>(b) * groups + g) * 2);
  const double m = s.x / n;
  double var = s.y / n - m * m;
  if (var < 0) var = 0;
  mean = static_cast<float>(m);
  rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
}

template <bool GS8>
__device__ __forceinline__ void gn_thread_coefs(GnCoef<GS8>& k, const double* __restrict__ stats, const float* __restrict__ gamma,
                                                const float* __restrict__ beta, const float* __restrict__ scale_shift, int b, int c0,
                                                int C, int groups, long long pix, float eps) {
  const int gs = C / groups;
  const double n = static_cast<double>(pix) * gs;
  float gv[8], bv[8], sh[8];
  {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0 + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c0)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c0 + 4));
    gv[0] = g0.x, gv[1] = g0.y, gv[2] = g0.z, gv[3] = g0.w, gv[4] = g1.x, gv[5] = g1.y, gv[6] = g1.z, gv[7] = g1.w;
    bv[0] = b0.x, bv[1] = b0.y, bv[2] = b0.z, bv[3] = b0.w, bv[4] = b1.x, bv[5] = b1.y, bv[6] = b1.z, bv[7] = b1.w;
  }
  if (scale_shift) {
    const float* sp = scale_shift + static_cast<long long>(b) * 2 * C + c0;
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(sp)), s1 = __ldg(reinterpret_cast<const float4*>(sp + 4));
    const float4 h0 = __ldg(reinterpret_cast<const float4*>(sp + C)), h1 = __ldg(reinterpret_cast<const float4*>(sp + C + 4));
    k.sc[0] = s0.x + 1.f, k.sc[1] = s0.y + 1.f, k.sc[2] = s0.z + 1.f, k.sc[3] = s0.w + 1.f;
    k.sc[4] = s1.x + 1.f, k.sc[5] = s1.y + 1.f, k.sc[6] = s1.z + 1.f, k.sc[7] = s1.w + 1.f;
    sh[0] = h0.x, sh[1] = h0.y, sh[2] = h0.z, sh[3] = h0.w, sh[4] = h1.x, sh[5] = h1.y, sh[6] = h1.z, sh[7] = h1.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) k.sc[j] = 1.f, sh[j] = 0.f;
  }
  if (GS8) {
    gn_group_stats(stats, b, c0 / gs, groups, n, eps, k.mean[0], k.rstd[0]);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) gn_group_stats(stats, b, (c0 + j) / gs, groups, n, eps, k.mean[GS8 ? 0 : j], k.rstd[GS8 ? 0 : j]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float mean = k.mean[GS8 ? 0 : j], rstd = k.rstd[GS8 ? 0 : j];
    float a = rstd * gv[j];
    float d = bv[j] - mean * a;
    k.a[j] = a * k.sc[j];
    k.d[j] = d * k.sc[j] + sh[j];
  }
}

// UNR 16-byte vectors of x (and of res) in flight per thread; MINB resident CTAs per SM (register cap)
template <int FMT, bool GS8, int UNR, int MINB>
__global__ void __launch_bounds__(256, MINB) gn_silu_fwd_kernel(const uint16_t* __restrict__ x, const uint16_t* __restrict__ res,
                                                                uint16_t* __restrict__ y, long long pix, int C, int groups,
                                                                const double* __restrict__ stats, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, const float* __restrict__ scale_shift,
                                                                float eps, int act) {
  pdl_trigger();
  const int b = blockIdx.y;
  const int i00 = blockIdx.x * blockDim.x + threadIdx.x;          // 32-bit vector indices (host: nvec < 2^30)
  const int nvec = static_cast<int>(pix * C / 8);
  if (i00 >= nvec) return;
  const int c0 = static_cast<int>((static_cast<long long>(i00) * 8) % C);   // fixed for this thread: the host keeps (grid stride * 8) % C == 0
  const uint4* xb = reinterpret_cast<const uint4*>(x + static_cast<long long>(b) * pix * C);
  const uint4* rb = res ? reinterpret_cast<const uint4*>(res + static_cast<long long>(b) * pix * C) : nullptr;
  uint4* yb = reinterpret_cast<uint4*>(y + static_cast<long long>(b) * pix * C);
  const int stride = gridDim.x * blockDim.x;
  uint4 xr[UNR], rr[UNR];
#pragma unroll
  for (int u = 0; u < UNR; ++u) {
    const int i = i00 + u * stride;
    if (i < nvec) {
      xr[u] = __ldg(xb + i);
      if (rb) rr[u] = __ldg(rb + i);
    }
  }
  GnCoef<GS8> k;
  gn_thread_coefs<GS8>(k, stats, gamma, beta, scale_shift, b, c0, C, groups, pix, eps);
  for (int i0 = i00;;) {
    uint4 yr[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      float v[8];
      unpack8<FMT>(xr[u], v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float uu = fmaf(v[j], k.a[j], k.d[j]);
        v[j] = act ? uu * sigmoid_fast<FMT>(uu) : uu;
      }
      if (rb) {   // identity skip of a ResnetBlock whose dim == dim_out (VDDP:297,311)
        float r[8];
        unpack8<FMT>(rr[u], r);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += r[j];
      }
      yr[u] = pack8<FMT>(v);
    }
    const int inext = i0 + UNR * stride;
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int i = i0 + u * stride;
      if (i < nvec) yb[i] = yr[u];
      const int in = inext + u * stride;
      if (in < nvec) {
        xr[u] = __ldg(xb + in);
        if (rb) rr[u] = __ldg(rb + in);
      }
    }
    i0 = inext;
    if (i0 >= nvec) break;
  }
}

// Backward pass 1: per (sample, channel)  S1 = sum_pix du,  S2 = sum_pix du * xhat   with du = dy * silu'(u).
// The loop accumulates S1 and T = sum du * x; S2 = rstd * (T - mean * S1) per channel at the end.
// part: [B][C][2] fp32, accumulated atomically (zeroed by the caller together with the arrival counters).
// The LAST CTA of a sample to arrive (counter[b]) turns part[b] into what pass 2 and the parameters need (the former finalize
// kernel):  gm[b][g] = group means of dxhat and dxhat * xhat;  dgamma[c] += (1+sc) S2, dbeta[c] += (1+sc) S1,
// d(scale_shift)[b][c] = gamma S2 + beta S1, [b][C+c] = S1.
template <int FMT, bool GS8, int UNR, int MINB>
__global__ void __launch_bounds__(256, MINB) gn_silu_bwd_reduce_kernel(const uint16_t* __restrict__ x, const uint16_t* __restrict__ dy,
                                                                       long long pix, int C, int groups, const double* __restrict__ stats,
                                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                       const float* __restrict__ scale_shift, float eps, int act,
                                                                       float* __restrict__ part, unsigned* __restrict__ counter,
                                                                       float* __restrict__ gm, float* __restrict__ dgamma,
                                                                       float* __restrict__ dbeta, float* __restrict__ dss) {
  pdl_trigger();
  extern __shared__ float red[];   // [2][C] block partials | [2][groups] group sums of the last CTA
  __shared__ int s_last;
  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  for (int c = tid; c < 2 * C; c += blockDim.x) red[c] = 0.f;
  const int i00 = blockIdx.x * blockDim.x + tid;
  const int nvec = static_cast<int>(pix * C / 8);
  const int c0 = static_cast<int>((static_cast<long long>(i00) * 8) % C);
  const uint4* xb = reinterpret_cast<const uint4*>(x + static_cast<long long>(b) * pix * C);
  const uint4* db = reinterpret_cast<const uint4*>(dy + static_cast<long long>(b) * pix * C);
  const int stride = gridDim.x * blockDim.x;
  uint4 xr[UNR], dr[UNR];
#pragma unroll
  for (int u = 0; u < UNR; ++u) {
    const int i = i00 + u * stride;
    if (i < nvec) {
      xr[u] = __ldg(xb + i);
      dr[u] = __ldg(db + i);
    }
  }
  GnCoef<GS8> k;
  gn_thread_coefs<GS8>(k, stats, gamma, beta, scale_shift, b, c0, C, groups, pix, eps);
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  for (int i0 = i00; i0 < nvec;) {
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int i = i0 + u * stride;
      if (i >= nvec) continue;
      float xv[8], dv[8];
      unpack8<FMT>(xr[u], xv);
      unpack8<FMT>(dr[u], dv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float du = act ? dv[j] * dsilu_fast<FMT>(fmaf(xv[j], k.a[j], k.d[j])) : dv[j];
        s1[j] += du;
        s2[j] = fmaf(du, xv[j], s2[j]);
      }
    }
    i0 += UNR * stride;
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int i = i0 + u * stride;
      if (i < nvec) {
        xr[u] = __ldg(xb + i);
        dr[u] = __ldg(db + i);
      }
    }
  }
  // lanes (l, l + vpr, ...) of a warp hold the same 8 channels: fold them with shuffles, then block partials, then global
  const int vpr = C >> 3;
  bool owner = i00 < nvec;
  if (vpr < 32 && (vpr & (vpr - 1)) == 0) {
    for (int o = 16; o >= vpr; o >>= 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], o);
        s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], o);
      }
    }
    owner = (tid & 31) < vpr;
  }
  __syncthreads();     // red[] zeroed
  if (owner) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&red[c0 + j], s1[j]);
      atomicAdd(&red[C + c0 + j], k.rstd[GS8 ? 0 : j] * (s2[j] - k.mean[GS8 ? 0 : j] * s1[j]));      // sum du * xhat
    }
  }
  __syncthreads();
  for (int c = tid; c < C; c += blockDim.x) {
    atomicAdd(part + (static_cast<long long>(b) * C + c) * 2, red[c]);
    atomicAdd(part + (static_cast<long long>(b) * C + c) * 2 + 1, red[C + c]);
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    s_last = atomicAdd(counter + b, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  float* gsum = red + 2 * C;
  if (tid < 2 * groups) gsum[tid] = 0.f;
  __syncthreads();
  const int gs = C / groups;
  for (int c = tid; c < C; c += blockDim.x) {
    const float sc = scale_shift ? scale_shift[static_cast<long long>(b) * 2 * C + c] + 1.f : 1.f;
    const float S1 = __ldcg(part + (static_cast<long long>(b) * C + c) * 2);
    const float S2 = __ldcg(part + (static_cast<long long>(b) * C + c) * 2 + 1);
    const float kk = gamma[c] * sc;
    atomicAdd(&gsum[c / gs], kk * S1);
    atomicAdd(&gsum[groups + c / gs], kk * S2);
    atomicAdd(dgamma + c, sc * S2);
    atomicAdd(dbeta + c, sc * S1);
    if (dss) {
      dss[static_cast<long long>(b) * 2 * C + c] = gamma[c] * S2 + beta[c] * S1;
      dss[static_cast<long long>(b) * 2 * C + C + c] = S1;
    }
  }
  __syncthreads();
  const float inv_n = 1.f / (static_cast<float>(pix) * gs);
  if (tid < groups) {
    gm[(static_cast<long long>(b) * groups + tid) * 2] = gsum[tid] * inv_n;
    gm[(static_cast<long long>(b) * groups + tid) * 2 + 1] = gsum[groups + tid] * inv_n;
  }
}

// Backward pass 2: dx = rstd * ( gamma*(1+sc)*du - m1_g - xhat * m2_g ) = K du - (x P + Q)  with per-channel
//   K = rstd gamma (1+sc),  P = rstd^2 m2_g,  Q = rstd (m1_g - mean rstd m2_g)
// (Walking the samples in reverse order, to find the rows pass 1 touched last still in L2, measured no difference.)
template <int FMT, bool GS8, int UNR, int MINB>
__global__ void __launch_bounds__(256, MINB) gn_silu_bwd_apply_kernel(const uint16_t* __restrict__ x, const uint16_t* __restrict__ dy,
                                                                      uint16_t* __restrict__ dx, long long pix, int C, int groups,
                                                                      const double* __restrict__ stats, const float* __restrict__ gamma,
                                                                      const float* __restrict__ beta, const float* __restrict__ scale_shift,
                                                                      float eps, int act, const float* __restrict__ gm,
                                                                      float* __restrict__ dx_colsum) {
  pdl_trigger();
  extern __shared__ float cs[];    // [C] block partial of the column sums of dx
  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  if (dx_colsum) {
    for (int c = tid; c < C; c += blockDim.x) cs[c] = 0.f;
  }
  const int i00 = blockIdx.x * blockDim.x + tid;
  const int nvec = static_cast<int>(pix * C / 8);
  const int c0 = static_cast<int>((static_cast<long long>(i00) * 8) % C);
  const uint4* xb = reinterpret_cast<const uint4*>(x + static_cast<long long>(b) * pix * C);
  const uint4* db = reinterpret_cast<const uint4*>(dy + static_cast<long long>(b) * pix * C);
  uint4* ob = reinterpret_cast<uint4*>(dx + static_cast<long long>(b) * pix * C);
  const int stride = gridDim.x * blockDim.x;
  uint4 xr[UNR], dr[UNR];
#pragma unroll
  for (int u = 0; u < UNR; ++u) {
    const int i = i00 + u * stride;
    if (i < nvec) {
      xr[u] = __ldg(xb + i);
      dr[u] = __ldg(db + i);
    }
  }
  GnCoef<GS8> k;
  gn_thread_coefs<GS8>(k, stats, gamma, beta, scale_shift, b, c0, C, groups, pix, eps);
  float cK[8], cP[GS8 ? 1 : 8], cQ[GS8 ? 1 : 8];
  {
    const int gs = C / groups;
    float gv[8];
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0 + 4));
    gv[0] = g0.x, gv[1] = g0.y, gv[2] = g0.z, gv[3] = g0.w, gv[4] = g1.x, gv[5] = g1.y, gv[6] = g1.z, gv[7] = g1.w;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float mean = k.mean[GS8 ? 0 : j], rstd = k.rstd[GS8 ? 0 : j];
      cK[j] = rstd * gv[j] * k.sc[j];
      if (!GS8 || j == 0) {
        const float2 m = *reinterpret_cast<const float2*>(gm + (static_cast<long long>(b) * groups + (c0 + j) / gs) * 2);
        cP[GS8 ? 0 : j] = rstd * rstd * m.y;
        cQ[GS8 ? 0 : j] = rstd * (m.x - mean * rstd * m.y);
      }
    }
  }
  float colacc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int i0 = i00; i0 < nvec;) {
    uint4 orr[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int i = i0 + u * stride;
      if (i >= nvec) continue;
      float xv[8], dv[8], o[8];
      unpack8<FMT>(xr[u], xv);
      unpack8<FMT>(dr[u], dv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float du = act ? dv[j] * dsilu_fast<FMT>(fmaf(xv[j], k.a[j], k.d[j])) : dv[j];
        o[j] = fmaf(cK[j], du, -fmaf(xv[j], cP[GS8 ? 0 : j], cQ[GS8 ? 0 : j]));
        colacc[j] += o[j];
      }
      orr[u] = pack8<FMT>(o);
    }
    const int inext = i0 + UNR * stride;
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int i = i0 + u * stride;
      if (i < nvec) ob[i] = orr[u];
      const int in = inext + u * stride;
      if (in < nvec) {
        xr[u] = __ldg(xb + in);
        dr[u] = __ldg(db + in);
      }
    }
    i0 = inext;
  }
  if (dx_colsum) {
    const int vpr = C >> 3;
    bool owner = i00 < nvec;
    if (vpr < 32 && (vpr & (vpr - 1)) == 0) {
      for (int o = 16; o >= vpr; o >>= 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) colacc[j] += __shfl_xor_sync(0xffffffffu, colacc[j], o);
      }
      owner = (tid & 31) < vpr;
    }
    __syncthreads();
    if (owner) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&cs[c0 + j], colacc[j]);
    }
    __syncthreads();
    for (int c = tid; c < C; c += blockDim.x) atomicAdd(dx_colsum + c, cs[c]);
  }
}

// ------------------------------------------------------------------------------------------------
// channel LayerNorm (gain only, biased variance)  VDDP:245-254, rows = every (b, f, h, w) position
// ------------------------------------------------------------------------------------------------
template <int VPT, int FMT>   // 16-byte vectors per thread (C = 8 * VPT * tpr); 16-bit format
__global__ void __launch_bounds__(256) ln_fwd_kernel(const uint16_t* __restrict__ x, uint16_t* __restrict__ y,
                                                     long long rows, int C, int tpr, const float* __restrict__ gamma, float eps,
                                                     float* __restrict__ mean_rstd) {
  pdl_trigger();
  const int rpb = blockDim.x / tpr;
  const int lr = threadIdx.x / tpr, lc = threadIdx.x % tpr;
  const float inv_c = 1.f / C;
  // the row loop is block-uniform so that the sub-warp shuffles always run with all lanes present
  for (long long rb = static_cast<long long>(blockIdx.x) * rpb; rb < rows; rb += static_cast<long long>(gridDim.x) * rpb) {
    const long long r = rb + lr;
    const bool live = r < rows;
    float v[VPT][8];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      if (live) {
        unpack8<FMT>(__ldg(reinterpret_cast<const uint4*>(x + r * C + (k * tpr + lc) * 8)), v[k]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[k][j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[k][j];
    }
    for (int o = tpr >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * inv_c;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < VPT; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[k][j] - mean;
        q += d * d;
      }
    for (int o = tpr >> 1; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * inv_c + eps);
    if (!live) continue;
    if (mean_rstd && lc == 0) {
      mean_rstd[r * 2] = mean;
      mean_rstd[r * 2 + 1] = rstd;
    }
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const int c0 = (k * tpr + lc) * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) v[k][j] = (v[k][j] - mean) * rstd * __ldg(gamma + c0 + j);
      *reinterpret_cast<uint4*>(y + r * C + c0) = pack8<FMT>(v[k]);
    }
  }
}

// backward: dx = rstd * (g*dy - mean_c(g*dy) - xhat * mean_c(g*dy*xhat)) (+ dres),  dgamma[c] += sum_rows dy*xhat
template <int VPT, int FMT>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const uint16_t* __restrict__ x, const uint16_t* __restrict__ dy,
                                                     const uint16_t* __restrict__ dres, uint16_t* __restrict__ dx,
                                                     long long rows, int C, int tpr, const float* __restrict__ gamma, float eps,
                                                     float* __restrict__ dgamma) {
  pdl_trigger();
  extern __shared__ float sdg[];   // [C] block partial of dgamma
  for (int c = threadIdx.x; c < C; c += blockDim.x) sdg[c] = 0.f;
  __syncthreads();
  const int rpb = blockDim.x / tpr;
  const int lr = threadIdx.x / tpr, lc = threadIdx.x % tpr;
  const float inv_c = 1.f / C;
  float dg[VPT][8];
#pragma unroll
  for (int k = 0; k < VPT; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) dg[k][j] = 0.f;
  for (long long rb = static_cast<long long>(blockIdx.x) * rpb; rb < rows; rb += static_cast<long long>(gridDim.x) * rpb) {
    const long long r = rb + lr;
    const bool live = r < rows;
    float v[VPT][8], d[VPT][8];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      if (live) {
        unpack8<FMT>(__ldg(reinterpret_cast<const uint4*>(x + r * C + (k * tpr + lc) * 8)), v[k]);
        unpack8<FMT>(__ldg(reinterpret_cast<const uint4*>(dy + r * C + (k * tpr + lc) * 8)), d[k]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[k][j] = d[k][j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[k][j];
    }
    for (int o = tpr >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * inv_c;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < VPT; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[k][j] -= mean;
        q += v[k][j] * v[k][j];
      }
    for (int o = tpr >> 1; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * inv_c + eps);
    float a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const int c0 = (k * tpr + lc) * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = v[k][j] * rstd;
        dg[k][j] += d[k][j] * xh;
        const float gd = d[k][j] * __ldg(gamma + c0 + j);
        v[k][j] = xh;
        d[k][j] = gd;
        a1 += gd;
        a2 += gd * xh;
      }
    }
    for (int o = tpr >> 1; o > 0; o >>= 1) {
      a1 += __shfl_xor_sync(0xffffffffu, a1, o);
      a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    a1 *= inv_c;
    a2 *= inv_c;
    if (!live) continue;
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const int c0 = (k * tpr + lc) * 8;
      float o8[8];
      if (dres) unpack8<FMT>(__ldg(reinterpret_cast<const uint4*>(dres + r * C + c0)), o8);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float g = rstd * (d[k][j] - a1 - v[k][j] * a2);
        o8[j] = dres ? o8[j] + g : g;
      }
      *reinterpret_cast<uint4*>(dx + r * C + c0) = pack8<FMT>(o8);
    }
  }
#pragma unroll
  for (int k = 0; k < VPT; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&sdg[(k * tpr + lc) * 8 + j], dg[k][j]);
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(dgamma + c, sdg[c]);
}

static int ln_shape(int C, int& tpr, int& vpt) {
  if (C % 8) return -1;
  const int vec = C / 8;
  tpr = 1;
  while (tpr < 32 && tpr * 2 <= vec && (vec % (tpr * 2)) == 0) tpr *= 2;
  vpt = vec / tpr;
  if (vpt > 4) return -1;
  return 0;
}

}  // namespace vmm

using namespace vmm;

// grid.x for the streaming GroupNorm kernels: `waves` full waves of `per_sm` resident CTAs per SM over the B samples, and
// (gridDim.x * 256 * 8) % C == 0 so that every thread keeps the same 8 channels over its whole grid-stride loop
static int gn_grid_x(long long nvec, int B, int C, int per_sm, int waves) {
  int gx = static_cast<int>(min64((nvec + 255) / 256, (static_cast<long long>(waves) * per_sm * num_sms() + B - 1) / B));
  if (gx < 1) gx = 1;
  int m = 1;
  while ((static_cast<long long>(m) * 2048) % C) ++m;
  return (gx + m - 1) / m * m;
}

static int gn_check(int fmt, long long pix, int C, int groups) {
  if (C % 8 || C % groups || C > 2048) return set_error(VMM_ERR_ARG, "vmm_gn_silu: C must be a multiple of 8 and of groups, at most 2048");
  if (fmt != VMM_FMT_F16 && fmt != VMM_FMT_BF16) return set_error(VMM_ERR_ARG, "vmm_gn_silu: bad fmt");
  if (pix * C / 8 >= (1LL << 30)) return set_error(VMM_ERR_UNSUPPORTED, "vmm_gn_silu: more than 2^33 elements per sample");
  return VMM_OK;
}

extern "C" int vmm_gn_silu_fwd(const void* x, const void* res, void* y, int fmt, int B, long long pix, int C, int groups, const double* stats,
                               const float* gamma, const float* beta, const float* scale_shift, float eps, int act, void* stream) {
  if (!x || !y || !stats || !gamma || !beta) return set_error(VMM_ERR_ARG, "vmm_gn_silu_fwd: null pointer");
  if (int rc = gn_check(fmt, pix, C, groups)) return rc;
  const long long nvec = pix * C / 8;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool gs8 = ((C / groups) % 8) == 0;
  const dim3 grid(gn_grid_x(nvec, B, C, 4, 2), B);
#define GN_FWD(F, G8) gn_silu_fwd_kernel<F, G8, 2, 4><<<grid, 256, 0, st>>>(static_cast<const uint16_t*>(x), static_cast<const uint16_t*>(res), static_cast<uint16_t*>(y), pix, C, groups, stats, gamma, beta, scale_shift, eps, act)
  if (fmt == VMM_FMT_F16) {
    if (gs8) GN_FWD(0, true); else GN_FWD(0, false);
  } else {
    if (gs8) GN_FWD(1, true); else GN_FWD(1, false);
  }
#undef GN_FWD
  count_launch();
  return check_launch("vmm_gn_silu_fwd");
}

// workspace: part [B][C][2] fp32 | group means [B][groups][2] fp32 | arrival counters [B] u32   (all zeroed per call)
extern "C" size_t vmm_gn_silu_bwd_workspace(int B, int C, int groups) {
  return (static_cast<size_t>(B) * C * 2 + static_cast<size_t>(B) + static_cast<size_t>(B) * groups * 2) * sizeof(float);
}

extern "C" int vmm_gn_silu_bwd(const void* x, const void* dy, void* dx, int fmt, int B, long long pix, int C, int groups,
                               const double* stats, const float* gamma, const float* beta, const float* scale_shift, float eps,
                               int act, float* dgamma, float* dbeta, float* dscale_shift, float* dx_colsum, void* workspace,
                               size_t workspace_bytes, void* stream_) {
  if (!x || !dy || !dx || !stats || !gamma || !beta || !dgamma || !dbeta || !workspace)
    return set_error(VMM_ERR_ARG, "vmm_gn_silu_bwd: null pointer");
  if (int rc = gn_check(fmt, pix, C, groups)) return rc;
  if (workspace_bytes < vmm_gn_silu_bwd_workspace(B, C, groups)) return set_error(VMM_ERR_ARG, "vmm_gn_silu_bwd: workspace too small");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  float* part = static_cast<float*>(workspace);
  float* gm = part + static_cast<size_t>(B) * C * 2;
  unsigned* counter = reinterpret_cast<unsigned*>(gm + static_cast<size_t>(B) * groups * 2);
  cudaError_t e = cudaMemsetAsync(part, 0, vmm_gn_silu_bwd_workspace(B, C, groups), stream);
  if (e != cudaSuccess) return set_cuda_error(e, "vmm_gn_silu_bwd: memset");
  const long long nvec = pix * C / 8;
  const uint16_t* xp = static_cast<const uint16_t*>(x);
  const uint16_t* dp = static_cast<const uint16_t*>(dy);
  uint16_t* op = static_cast<uint16_t*>(dx);
  const bool gs8 = ((C / groups) % 8) == 0;
  const dim3 grid(gn_grid_x(nvec, B, C, 2, 2), B);     // measured: 4 vectors in flight x 2 CTAs / SM beats 2 x 3 and 2 x 4 (108 / 123 / 123 us, level 0)
  // The reduce pass ends in 2 C global atomics per CTA onto the B x 2 C partial sums: at the small levels (C = 256 / 512, a few MB per
  // launch) 74 CTAs per sample spent ~35 us mostly in those contended atomics (ncu launch list, round 2: 35 us at 12 x 12 against 22 us at
  // 48 x 48).  Give every thread at least `min_vec` vectors there, i.e. fewer CTAs per sample.
  static const int min_vec = [] { const char* e = getenv("VMM_GN_REDUCE_MIN_VEC"); return e ? atoi(e) : 16; }();
  dim3 grid_r = grid;
  if (min_vec > 0) {
    int m = 1;
    while ((static_cast<long long>(m) * 2048) % C) ++m;
    long long want = nvec / (256LL * min_vec);
    want = want / m * m;
    if (want < m) want = m;
    if (want < static_cast<long long>(grid.x)) grid_r.x = static_cast<unsigned int>(want);
  }
  const size_t sm_r = (static_cast<size_t>(2) * C + 2 * groups) * sizeof(float), sm_a = static_cast<size_t>(C) * sizeof(float);
#define GN_BWD(F, G8)                                                                                                                    \
  do {                                                                                                                                   \
    gn_silu_bwd_reduce_kernel<F, G8, 4, 2><<<grid_r, 256, sm_r, stream>>>(xp, dp, pix, C, groups, stats, gamma, beta, scale_shift, eps, act, \
                                                                        part, counter, gm, dgamma, dbeta, dscale_shift);                 \
    count_launch();                                                                                                                      \
    gn_silu_bwd_apply_kernel<F, G8, 4, 2><<<grid, 256, sm_a, stream>>>(xp, dp, op, pix, C, groups, stats, gamma, beta, scale_shift, eps,  \
                                                                       act, gm, dx_colsum);                                              \
    count_launch();                                                                                                                      \
  } while (0)
  if (fmt == VMM_FMT_F16) {
    if (gs8) GN_BWD(0, true); else GN_BWD(0, false);
  } else {
    if (gs8) GN_BWD(1, true); else GN_BWD(1, false);
  }
#undef GN_BWD
  return check_launch("vmm_gn_silu_bwd");
}

#define LN_DISPATCH(KERNEL, ...)                                  \
  switch (vpt * 2 + (fmt == VMM_FMT_BF16 ? 1 : 0)) {              \
    case 2: KERNEL<1, 0> __VA_ARGS__; break;                      \
    case 3: KERNEL<1, 1> __VA_ARGS__; break;                      \
    case 4: KERNEL<2, 0> __VA_ARGS__; break;                      \
    case 5: KERNEL<2, 1> __VA_ARGS__; break;                      \
    case 8: KERNEL<4, 0> __VA_ARGS__; break;                      \
    case 9: KERNEL<4, 1> __VA_ARGS__; break;                      \
    default: return set_error(VMM_ERR_UNSUPPORTED, "layernorm: unsupported channel count"); \
  }

extern "C" int vmm_ln_fwd(const void* x, void* y, int fmt, long long rows, int C, const float* gamma, float eps, float* mean_rstd,
                          void* stream_) {
  if (!x || !y || !gamma) return set_error(VMM_ERR_ARG, "vmm_ln_fwd: null pointer");
  int tpr, vpt;
  if (ln_shape(C, tpr, vpt)) return set_error(VMM_ERR_UNSUPPORTED, "vmm_ln_fwd: unsupported channel count");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int rpb = 256 / tpr;
  const int grid = static_cast<int>(min64((rows + rpb - 1) / rpb, 8LL * num_sms()));
  LN_DISPATCH(ln_fwd_kernel, <<<grid, 256, 0, stream>>>(static_cast<const uint16_t*>(x), static_cast<uint16_t*>(y), rows, C, tpr,
                                                        gamma, eps, mean_rstd));
  count_launch();
  return check_launch("vmm_ln_fwd");
}

extern "C" int vmm_ln_bwd(const void* x, const void* dy, const void* dres, void* dx, int fmt, long long rows, int C,
                          const float* gamma, float eps, float* dgamma, void* stream_) {
  if (!x || !dy || !dx || !gamma || !dgamma) return set_error(VMM_ERR_ARG, "vmm_ln_bwd: null pointer");
  int tpr, vpt;
  if (ln_shape(C, tpr, vpt)) return set_error(VMM_ERR_UNSUPPORTED, "vmm_ln_bwd: unsupported channel count");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int rpb = 256 / tpr;
  const int grid = static_cast<int>(min64((rows + rpb - 1) / rpb, 4LL * num_sms()));
  LN_DISPATCH(ln_bwd_kernel, <<<grid, 256, C * sizeof(float), stream>>>(static_cast<const uint16_t*>(x), static_cast<const uint16_t*>(dy),
                                                                        static_cast<const uint16_t*>(dres), static_cast<uint16_t*>(dx),
                                                                        rows, C, tpr, gamma, eps, dgamma));
  count_launch();
  return check_launch("vmm_ln_bwd");
}
