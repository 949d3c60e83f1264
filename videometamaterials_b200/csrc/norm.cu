// GroupNorm (+ time scale/shift + SiLU) and channel LayerNorm, forward and backward.
// All tensors are channels-last [sample, pixel, C] 16-bit; statistics / parameters are fp32 or fp64.
// These kernels are HBM-bound: 16-byte vector accesses, one read (+ one write) of the activation.
#include "common.cuh"
#include "mma_sync.cuh"
#include "sm100_ptx.cuh"

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

// per-channel affine of the forward: u = x * a + d   (mean / rstd / gamma / beta / scale / shift folded)
__device__ __forceinline__ void gn_channel_affine(const double* __restrict__ stats, const float* __restrict__ gamma,
                                                  const float* __restrict__ beta, const float* __restrict__ scale_shift, int b, int c,
                                                  int C, int groups, long long pix, float eps, float& a, float& d, float& mean,
                                                  float& rstd, float& sc) {
  const int gs = C / groups;
  const int g = c / gs;
  const double n = static_cast<double>(pix) * gs;
  const double s1 = stats[(static_cast<long long>(b) * groups + g) * 2];
  const double s2 = stats[(static_cast<long long>(b) * groups + g) * 2 + 1];
  const double m = s1 / n;
  double var = s2 / n - m * m;
  if (var < 0) var = 0;
  mean = static_cast<float>(m);
  rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  a = rstd * gamma[c];
  d = beta[c] - mean * a;
  sc = 1.f;
  if (scale_shift) {
    sc = scale_shift[static_cast<long long>(b) * 2 * C + c] + 1.f;
    const float sh = scale_shift[static_cast<long long>(b) * 2 * C + C + c];
    a *= sc;
    d = d * sc + sh;
  }
}

template <int FMT>
__global__ void __launch_bounds__(256) gn_silu_fwd_kernel(const uint16_t* __restrict__ x, const uint16_t* __restrict__ res,
                                                          uint16_t* __restrict__ y, long long pix, int C, int groups,
                                                          const double* __restrict__ stats, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, const float* __restrict__ scale_shift,
                                                          float eps, int act) {
  extern __shared__ float coef[];   // [2][C]: the fp64 statistics -> affine step runs once per channel and block, not per thread
  const int b = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mean, rstd, sc;
    gn_channel_affine(stats, gamma, beta, scale_shift, b, c, C, groups, pix, eps, coef[c], coef[C + c], mean, rstd, sc);
  }
  __syncthreads();
  const long long i00 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int c0 = static_cast<int>((i00 * 8) % C);          // fixed for this thread: the host keeps (grid stride * 8) % C == 0
  float ca[8], cd[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ca[j] = coef[c0 + j];
    cd[j] = coef[C + c0 + j];
  }
  const long long nvec = pix * C / 8;
  const uint4* xb = reinterpret_cast<const uint4*>(x + static_cast<long long>(b) * pix * C);
  const uint4* rb = res ? reinterpret_cast<const uint4*>(res + static_cast<long long>(b) * pix * C) : nullptr;
  uint4* yb = reinterpret_cast<uint4*>(y + static_cast<long long>(b) * pix * C);
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i0 = i00; i0 < nvec; i0 += 4 * stride) {
    uint4 xr[4], rr[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = i0 + u * stride;
      if (i < nvec) {
        xr[u] = __ldg(xb + i);
        if (rb) rr[u] = __ldg(rb + i);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = i0 + u * stride;
      if (i >= nvec) continue;
      float v[8];
      unpack8<FMT>(xr[u], v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float uu = fmaf(v[j], ca[j], cd[j]);
        v[j] = act ? uu * sigmoid_fast<FMT>(uu) : uu;
      }
      if (rb) {   // identity skip of a ResnetBlock whose dim == dim_out (VDDP:297,311)
        float r[8];
        unpack8<FMT>(rr[u], r);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += r[j];
      }
      yb[i] = pack8<FMT>(v);
    }
  }
}

// Backward pass 1: per (sample, channel)  S1 = sum_pix du,  S2 = sum_pix du * xhat   with du = dy * silu'(u).
// The loop accumulates S1 and T = sum du * x; S2 = rstd * (T - mean * S1) per channel at the end.
// part: [B][C][2] fp32, accumulated atomically (zeroed by the caller).
template <int FMT>
__global__ void __launch_bounds__(256) gn_silu_bwd_reduce_kernel(const uint16_t* __restrict__ x, const uint16_t* __restrict__ dy,
                                                                 long long pix, int C, int groups, const double* __restrict__ stats,
                                                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                 const float* __restrict__ scale_shift, float eps, int act,
                                                                 float* __restrict__ part) {
  extern __shared__ float red[];   // [2][C] block partials, then [4][C] per-channel coefficients (a, d, mean, rstd)
  float* coef = red + 2 * C;
  const int b = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float sc;
    red[c] = red[C + c] = 0.f;
    gn_channel_affine(stats, gamma, beta, scale_shift, b, c, C, groups, pix, eps, coef[c], coef[C + c], coef[2 * C + c], coef[3 * C + c], sc);
  }
  __syncthreads();
  const long long i00 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int c0 = static_cast<int>((i00 * 8) % C);
  float ca[8], cd[8], cm[8], cr[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ca[j] = coef[c0 + j];
    cd[j] = coef[C + c0 + j];
    cm[j] = coef[2 * C + c0 + j];
    cr[j] = coef[3 * C + c0 + j];
  }
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  const long long nvec = pix * C / 8;
  const uint4* xb = reinterpret_cast<const uint4*>(x + static_cast<long long>(b) * pix * C);
  const uint4* db = reinterpret_cast<const uint4*>(dy + static_cast<long long>(b) * pix * C);
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i0 = i00; i0 < nvec; i0 += 4 * stride) {
    uint4 xr[4], dr[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = i0 + u * stride;
      if (i < nvec) {
        xr[u] = __ldg(xb + i);
        dr[u] = __ldg(db + i);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = i0 + u * stride;
      if (i >= nvec) continue;
      float xv[8], dv[8];
      unpack8<FMT>(xr[u], xv);
      unpack8<FMT>(dr[u], dv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float du = act ? dv[j] * dsilu_fast<FMT>(fmaf(xv[j], ca[j], cd[j])) : dv[j];
        s1[j] += du;
        s2[j] = fmaf(du, xv[j], s2[j]);
      }
    }
  }
  // lanes (l, l + vpr, ...) of a warp hold the same 8 channels: fold them with shuffles, then block partials, then global
  const int vpr = C >> 3;
  bool owner = true;
  if (vpr < 32 && (vpr & (vpr - 1)) == 0) {
    for (int o = 16; o >= vpr; o >>= 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], o);
        s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], o);
      }
    }
    owner = (threadIdx.x & 31) < vpr;
  }
  if (owner) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&red[c0 + j], s1[j]);
      atomicAdd(&red[C + c0 + j], cr[j] * (s2[j] - cm[j] * s1[j]));      // sum du * xhat
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    atomicAdd(part + (static_cast<long long>(b) * C + c) * 2, red[c]);
    atomicAdd(part + (static_cast<long long>(b) * C + c) * 2 + 1, red[C + c]);
  }
}

// Backward finalize (tiny): from part[B][C][2] produce
//   coefficients for pass 2:  k1[b][c] = rstd*gamma*(1+sc),  m1[b][g], m2[b][g]  (group means of dxhat and dxhat*xhat)
//   parameter grads: dgamma[c] += sum_b (1+sc) S2, dbeta[c] += sum_b (1+sc) S1, d(scale_shift)[b][c] = gamma*S2 + beta*S1, [b][C+c] = S1
__global__ void gn_silu_bwd_finalize_kernel(const float* __restrict__ part, int B, long long pix, int C, int groups,
                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                            const float* __restrict__ scale_shift, float* __restrict__ gm /*[B][groups][2]*/,
                                            float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dss) {
  const int gs = C / groups;
  const float inv_n = 1.f / (static_cast<float>(pix) * gs);
  // group means
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B * groups; i += gridDim.x * blockDim.x) {
    const int b = i / groups, g = i % groups;
    float a1 = 0.f, a2 = 0.f;
    for (int c = g * gs; c < (g + 1) * gs; ++c) {
      const float sc = scale_shift ? scale_shift[static_cast<long long>(b) * 2 * C + c] + 1.f : 1.f;
      const float k = gamma[c] * sc;
      a1 += k * part[(static_cast<long long>(b) * C + c) * 2];
      a2 += k * part[(static_cast<long long>(b) * C + c) * 2 + 1];
    }
    gm[i * 2] = a1 * inv_n;
    gm[i * 2 + 1] = a2 * inv_n;
  }
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
    float dg = 0.f, db = 0.f;
    for (int b = 0; b < B; ++b) {
      const float sc = scale_shift ? scale_shift[static_cast<long long>(b) * 2 * C + c] + 1.f : 1.f;
      const float S1 = part[(static_cast<long long>(b) * C + c) * 2];
      const float S2 = part[(static_cast<long long>(b) * C + c) * 2 + 1];
      dg += sc * S2;
      db += sc * S1;
      if (dss) {
        dss[static_cast<long long>(b) * 2 * C + c] = gamma[c] * S2 + beta[c] * S1;
        dss[static_cast<long long>(b) * 2 * C + C + c] = S1;
      }
    }
    dgamma[c] += dg;
    dbeta[c] += db;
  }
}

// Backward pass 2: dx = rstd * ( gamma*(1+sc)*du - m1_g - xhat * m2_g ) = K du - (x P + Q)  with per-channel
//   K = rstd gamma (1+sc),  P = rstd^2 m2_g,  Q = rstd (m1_g - mean rstd m2_g)
template <int FMT>
__global__ void __launch_bounds__(256) gn_silu_bwd_apply_kernel(const uint16_t* __restrict__ x, const uint16_t* __restrict__ dy,
                                                                uint16_t* __restrict__ dx, long long pix, int C, int groups,
                                                                const double* __restrict__ stats, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, const float* __restrict__ scale_shift,
                                                                float eps, int act, const float* __restrict__ gm,
                                                                float* __restrict__ dx_colsum) {
  extern __shared__ float cs[];    // [C] block partial of the column sums of dx, then [5][C] coefficients (a, d, K, P, Q)
  float* coef = cs + C;
  const int b = blockIdx.y;
  const int gs = C / groups;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mean, rstd, sc;
    cs[c] = 0.f;
    gn_channel_affine(stats, gamma, beta, scale_shift, b, c, C, groups, pix, eps, coef[c], coef[C + c], mean, rstd, sc);
    const int g = c / gs;
    const float m1 = gm[(static_cast<long long>(b) * groups + g) * 2];
    const float m2 = gm[(static_cast<long long>(b) * groups + g) * 2 + 1];
    coef[2 * C + c] = rstd * gamma[c] * sc;
    coef[3 * C + c] = rstd * rstd * m2;
    coef[4 * C + c] = rstd * (m1 - mean * rstd * m2);
  }
  __syncthreads();
  const long long i00 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int c0 = static_cast<int>((i00 * 8) % C);
  float ca[8], cd[8], cK[8], cP[8], cQ[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ca[j] = coef[c0 + j];
    cd[j] = coef[C + c0 + j];
    cK[j] = coef[2 * C + c0 + j];
    cP[j] = coef[3 * C + c0 + j];
    cQ[j] = coef[4 * C + c0 + j];
  }
  const long long nvec = pix * C / 8;
  const uint4* xb = reinterpret_cast<const uint4*>(x + static_cast<long long>(b) * pix * C);
  const uint4* db = reinterpret_cast<const uint4*>(dy + static_cast<long long>(b) * pix * C);
  uint4* ob = reinterpret_cast<uint4*>(dx + static_cast<long long>(b) * pix * C);
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  float colacc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (long long i0 = i00; i0 < nvec; i0 += 4 * stride) {
    uint4 xr[4], dr[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = i0 + u * stride;
      if (i < nvec) {
        xr[u] = __ldg(xb + i);
        dr[u] = __ldg(db + i);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = i0 + u * stride;
      if (i >= nvec) continue;
      float xv[8], dv[8], o[8];
      unpack8<FMT>(xr[u], xv);
      unpack8<FMT>(dr[u], dv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float du = act ? dv[j] * dsilu_fast<FMT>(fmaf(xv[j], ca[j], cd[j])) : dv[j];
        o[j] = fmaf(cK[j], du, -fmaf(xv[j], cP[j], cQ[j]));
        colacc[j] += o[j];
      }
      ob[i] = pack8<FMT>(o);
    }
  }
  if (dx_colsum) {
    const int vpr = C >> 3;
    bool owner = true;
    if (vpr < 32 && (vpr & (vpr - 1)) == 0) {
      for (int o = 16; o >= vpr; o >>= 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) colacc[j] += __shfl_xor_sync(0xffffffffu, colacc[j], o);
      }
      owner = (threadIdx.x & 31) < vpr;
    }
    if (owner) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&cs[c0 + j], colacc[j]);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(dx_colsum + c, cs[c]);
  }
}

// ------------------------------------------------------------------------------------------------
// channel LayerNorm (gain only, biased variance)  VDDP:245-254, rows = every (b, f, h, w) position
// ------------------------------------------------------------------------------------------------
template <int VPT, int FMT>   // 16-byte vectors per thread (C = 8 * VPT * tpr); 16-bit format
__global__ void __launch_bounds__(256) ln_fwd_kernel(const uint16_t* __restrict__ x, uint16_t* __restrict__ y,
                                                     long long rows, int C, int tpr, const float* __restrict__ gamma, float eps,
                                                     float* __restrict__ mean_rstd) {
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

// grid.x for the streaming GroupNorm kernels: enough CTAs to fill the machine, and (gridDim.x * 256 * 8) % C == 0 so that
// every thread keeps the same 8 channels over its whole grid-stride loop
static int gn_grid_x(long long nvec, int B, int C) {
  int gx = static_cast<int>(min64((nvec + 255) / 256, (4LL * num_sms() + B - 1) / B * 2));
  if (gx < 1) gx = 1;
  int m = 1;
  while ((static_cast<long long>(m) * 2048) % C) ++m;
  return (gx + m - 1) / m * m;
}

extern "C" int vmm_gn_silu_fwd(const void* x, const void* res, void* y, int fmt, int B, long long pix, int C, int groups, const double* stats,
                               const float* gamma, const float* beta, const float* scale_shift, float eps, int act, void* stream) {
  if (!x || !y || !stats || !gamma || !beta) return set_error(VMM_ERR_ARG, "vmm_gn_silu_fwd: null pointer");
  if (C % 8 || C % groups || C > 4096) return set_error(VMM_ERR_ARG, "vmm_gn_silu_fwd: C must be a multiple of 8 and of groups");
  if (fmt != VMM_FMT_F16 && fmt != VMM_FMT_BF16) return set_error(VMM_ERR_ARG, "vmm_gn_silu_fwd: bad fmt");
  const long long nvec = pix * C / 8;
  const dim3 grid(gn_grid_x(nvec, B, C), B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (fmt == VMM_FMT_F16)
    gn_silu_fwd_kernel<0><<<grid, 256, 2 * C * sizeof(float), st>>>(static_cast<const uint16_t*>(x), static_cast<const uint16_t*>(res), static_cast<uint16_t*>(y), pix,
                                                C, groups, stats, gamma, beta, scale_shift, eps, act);
  else
    gn_silu_fwd_kernel<1><<<grid, 256, 2 * C * sizeof(float), st>>>(static_cast<const uint16_t*>(x), static_cast<const uint16_t*>(res), static_cast<uint16_t*>(y), pix,
                                                C, groups, stats, gamma, beta, scale_shift, eps, act);
  count_launch();
  return check_launch("vmm_gn_silu_fwd");
}

extern "C" size_t vmm_gn_silu_bwd_workspace(int B, int C, int groups) {
  return (static_cast<size_t>(B) * C * 2 + static_cast<size_t>(B) * groups * 2) * sizeof(float);
}

extern "C" int vmm_gn_silu_bwd(const void* x, const void* dy, void* dx, int fmt, int B, long long pix, int C, int groups,
                               const double* stats, const float* gamma, const float* beta, const float* scale_shift, float eps,
                               int act, float* dgamma, float* dbeta, float* dscale_shift, float* dx_colsum, void* workspace,
                               size_t workspace_bytes, void* stream_) {
  if (!x || !dy || !dx || !stats || !gamma || !beta || !dgamma || !dbeta || !workspace)
    return set_error(VMM_ERR_ARG, "vmm_gn_silu_bwd: null pointer");
  if (C % 8 || C % groups || C > 2048) return set_error(VMM_ERR_ARG, "vmm_gn_silu_bwd: bad C");
  if (fmt != VMM_FMT_F16 && fmt != VMM_FMT_BF16) return set_error(VMM_ERR_ARG, "vmm_gn_silu_bwd: bad fmt");
  if (workspace_bytes < vmm_gn_silu_bwd_workspace(B, C, groups)) return set_error(VMM_ERR_ARG, "vmm_gn_silu_bwd: workspace too small");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  float* part = static_cast<float*>(workspace);
  float* gm = part + static_cast<size_t>(B) * C * 2;
  cudaError_t e = cudaMemsetAsync(part, 0, static_cast<size_t>(B) * C * 2 * sizeof(float), stream);
  if (e != cudaSuccess) return set_cuda_error(e, "vmm_gn_silu_bwd: memset");
  const long long nvec = pix * C / 8;
  const dim3 grid(gn_grid_x(nvec, B, C), B);
  const uint16_t* xp = static_cast<const uint16_t*>(x);
  const uint16_t* dp = static_cast<const uint16_t*>(dy);
  if (fmt == VMM_FMT_F16)
    gn_silu_bwd_reduce_kernel<0><<<grid, 256, 6 * C * sizeof(float), stream>>>(xp, dp, pix, C, groups, stats, gamma, beta, scale_shift, eps, act, part);
  else
    gn_silu_bwd_reduce_kernel<1><<<grid, 256, 6 * C * sizeof(float), stream>>>(xp, dp, pix, C, groups, stats, gamma, beta, scale_shift, eps, act, part);
  count_launch();
  gn_silu_bwd_finalize_kernel<<<4, 256, 0, stream>>>(part, B, pix, C, groups, gamma, beta, scale_shift, gm, dgamma, dbeta, dscale_shift);
  count_launch();
  if (fmt == VMM_FMT_F16)
    gn_silu_bwd_apply_kernel<0><<<grid, 256, 6 * C * sizeof(float), stream>>>(xp, dp, static_cast<uint16_t*>(dx), pix, C, groups, stats, gamma, beta,
                                                                         scale_shift, eps, act, gm, dx_colsum);
  else
    gn_silu_bwd_apply_kernel<1><<<grid, 256, 6 * C * sizeof(float), stream>>>(xp, dp, static_cast<uint16_t*>(dx), pix, C, groups, stats, gamma, beta,
                                                                         scale_shift, eps, act, gm, dx_colsum);
  count_launch();
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
