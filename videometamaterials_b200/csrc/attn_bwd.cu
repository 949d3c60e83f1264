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
// temporal attention backward.  grid = (ctas_per_sample, B); each CTA walks pixel groups of PX pixels.
// phase 1: thread = (pixel, head, query i)  -> dq, and P / dS rows into shared memory
// phase 2: thread = (pixel, head, key j)    -> dk, dv of the frame keys; cond-key gradients kept in registers
// ------------------------------------------------------------------------------------------------
template <int NF>
__global__ void __launch_bounds__(192, 1) tattn_bwd_kernel(const uint16_t* __restrict__ qkv, const float* __restrict__ ekv,
                                                           const float* __restrict__ bias, const float* __restrict__ rot,
                                                           const uint16_t* __restrict__ dout, uint16_t* __restrict__ dqkv,
                                                           float* __restrict__ dekv, float* __restrict__ dbias, int fmt, int HW,
                                                           int heads, float scale, int PX) {
  extern __shared__ float sm[];
  const int HD = heads * BDH;
  const int HS = heads * BDHP;
  const int NK2 = 2 * NF;
  float* Ks = sm;                              // [PX][NF][HS]  rotated keys
  float* Vs = Ks + PX * NF * HS;
  float* Qs = Vs + PX * NF * HS;               // rotated, scaled queries
  float* Ds = Qs + PX * NF * HS;               // dO
  float* Ps = Ds + PX * NF * HS;               // [PX][heads][NF][2NF]
  float* Ss = Ps + PX * heads * NF * NK2;      // dS
  float* EK = Ss + PX * heads * NF * NK2;      // [NF][HS]
  float* EV = EK + NF * HS;
  float* RT = EV + NF * HS;                    // [NF][16][2]
  float* BS = RT + NF * 32;                    // [heads][NF][NF]
  float* GA = BS + heads * NF * NF;            // [64][nth]  cond-key gradients of (head, token) = this thread's (th, ti), summed over pixels
  float* GB = GA + 64 * blockDim.x;            // [NF][nth]  bias gradient row (th, ti, :)
  const int b = blockIdx.y;
  const int tid = threadIdx.x, nth = blockDim.x;
  const bool cond = ekv != nullptr;
  const int NK = cond ? NK2 : NF;

  for (int i = tid; i < NF * 32; i += nth) RT[i] = rot[i];
  for (int i = tid; i < heads * NF * NF; i += nth) BS[i] = bias[i];
  if (cond) {
    for (int i = tid; i < NF * HD; i += nth) {
      const int j = i / HD, c = i % HD;
      EK[j * HS + (c / BDH) * BDHP + (c % BDH)] = ekv[(static_cast<long long>(b) * NF + j) * 2 * HD + c];
      EV[j * HS + (c / BDH) * BDHP + (c % BDH)] = ekv[(static_cast<long long>(b) * NF + j) * 2 * HD + HD + c];
    }
  }
  const int ti = tid % NF;                      // query index (phase 1) / key index (phase 2)
  const int th = (tid / NF) % heads;
  const int tp = tid / (NF * heads);
  const bool tlive = tp < PX;
  for (int k = 0; k < 64; ++k) GA[k * nth + tid] = 0.f;     // thread-private columns: no synchronisation needed
  for (int k = 0; k < NF; ++k) GB[k * nth + tid] = 0.f;
  __syncthreads();

  const int groups = (HW + PX - 1) / PX;
  for (int grp = blockIdx.x; grp < groups; grp += gridDim.x) {
    const int p0 = grp * PX;
    // ---- stage K (rotated) and V
    const int vec_per_row = HD / 8;
    for (int i = tid; i < PX * NF * vec_per_row; i += nth) {
      const int c8 = i % vec_per_row;
      const int f = (i / vec_per_row) % NF;
      const int p = i / (vec_per_row * NF);
      float kv[8], vv[8];
      if (p0 + p < HW) {
        const uint16_t* row = qkv + ((static_cast<long long>(b) * NF + f) * HW + p0 + p) * 3 * HD;
        bld8f(row + HD + c8 * 8, fmt, kv);
        bld8f(row + 2 * HD + c8 * 8, fmt, vv);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) kv[j] = vv[j] = 0.f;
      }
      const int c = c8 * 8;
      const int h = c / BDH, d0 = c % BDH;
      float* kd = Ks + (p * NF + f) * HS + h * BDHP + d0;
      float* vd = Vs + (p * NF + f) * HS + h * BDHP + d0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float cs = RT[(f * 16 + d0 / 2 + j) * 2], sn = RT[(f * 16 + d0 / 2 + j) * 2 + 1];
        kd[2 * j] = kv[2 * j] * cs - kv[2 * j + 1] * sn;
        kd[2 * j + 1] = kv[2 * j + 1] * cs + kv[2 * j] * sn;
        vd[2 * j] = vv[2 * j];
        vd[2 * j + 1] = vv[2 * j + 1];
      }
    }
    __syncthreads();
    const bool live = tlive && (p0 + tp < HW);
    // ---- phase 1: query ti
    if (live) {
      const int i = ti, h = th, p = tp;
      float q[BDH], dO[BDH];
      const long long rowi = (static_cast<long long>(b) * NF + i) * HW + p0 + p;
      {
        const uint16_t* row = qkv + rowi * 3 * HD + h * BDH;
        const uint16_t* drow = dout + rowi * HD + h * BDH;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          bld8f(row + k * 8, fmt, q + k * 8);
          bld8f(drow + k * 8, fmt, dO + k * 8);
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float cs = RT[(i * 16 + k) * 2], sn = RT[(i * 16 + k) * 2 + 1];
          const float a = q[2 * k] * scale, c = q[2 * k + 1] * scale;
          q[2 * k] = a * cs - c * sn;
          q[2 * k + 1] = c * cs + a * sn;
        }
        float* qd = Qs + (p * NF + i) * HS + h * BDHP;
        float* dd = Ds + (p * NF + i) * HS + h * BDHP;
#pragma unroll
        for (int k = 0; k < BDH; ++k) {
          qd[k] = q[k];
          dd[k] = dO[k];
        }
      }
      // scores and dP rows live in shared memory (prow / srow), not in registers: the fully unrolled version
      // spilled ~3 KB per thread
      float* prow = Ps + ((p * heads + h) * NF + i) * NK2;
      float* srow = Ss + ((p * heads + h) * NF + i) * NK2;
      float mx = -1e30f;
#pragma unroll 1
      for (int j = 0; j < NK; ++j) {
        const int jf = cond ? j - NF : j;
        const float* kr = (cond && j < NF) ? (EK + j * HS + h * BDHP) : (Ks + (p * NF + jf) * HS + h * BDHP);
        const float* vr = (cond && j < NF) ? (EV + j * HS + h * BDHP) : (Vs + (p * NF + jf) * HS + h * BDHP);
        float acc = 0.f, accv = 0.f;
#pragma unroll
        for (int k = 0; k < BDH; k += 4) {
          const float4 kk = *reinterpret_cast<const float4*>(kr + k);
          const float4 vv = *reinterpret_cast<const float4*>(vr + k);
          acc += q[k] * kk.x + q[k + 1] * kk.y + q[k + 2] * kk.z + q[k + 3] * kk.w;
          accv += dO[k] * vv.x + dO[k + 1] * vv.y + dO[k + 2] * vv.z + dO[k + 3] * vv.w;
        }
        const int jb = (j < NF) ? j : j - NF;
        acc += BS[(h * NF + i) * NF + jb];
        prow[j] = acc;
        srow[j] = accv;
        mx = fmaxf(mx, acc);
      }
      float sum = 0.f;
#pragma unroll 1
      for (int j = 0; j < NK; ++j) {
        const float e = __expf(prow[j] - mx);
        prow[j] = e;
        sum += e;
      }
      const float inv = 1.f / sum;
      float Dsum = 0.f;
#pragma unroll 1
      for (int j = 0; j < NK; ++j) {
        const float pj = prow[j] * inv;
        prow[j] = pj;
        Dsum += pj * srow[j];
      }
      float dq[BDH];
#pragma unroll
      for (int k = 0; k < BDH; ++k) dq[k] = 0.f;
#pragma unroll 1
      for (int j = 0; j < NK; ++j) {
        const float ds = prow[j] * (srow[j] - Dsum);
        srow[j] = ds;
        const int jb = (j < NF) ? j : j - NF;
        GB[jb * nth + tid] += ds;
        const int jf = cond ? j - NF : j;
        const float* kr = (cond && j < NF) ? (EK + j * HS + h * BDHP) : (Ks + (p * NF + jf) * HS + h * BDHP);
#pragma unroll
        for (int k = 0; k < BDH; k += 4) {
          const float4 kk = *reinterpret_cast<const float4*>(kr + k);
          dq[k] += ds * kk.x;
          dq[k + 1] += ds * kk.y;
          dq[k + 2] += ds * kk.z;
          dq[k + 3] += ds * kk.w;
        }
      }
      // un-rotate and scale: q_rot = R(theta_i) (scale q)  ->  dq = scale R^T dq_rot
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const float cs = RT[(i * 16 + k) * 2], sn = RT[(i * 16 + k) * 2 + 1];
        const float a = dq[2 * k], c = dq[2 * k + 1];
        dq[2 * k] = (a * cs + c * sn) * scale;
        dq[2 * k + 1] = (c * cs - a * sn) * scale;
      }
      uint16_t* orow = dqkv + rowi * 3 * HD + h * BDH;
#pragma unroll
      for (int k = 0; k < 4; ++k) bst8f(orow + k * 8, fmt, dq + k * 8);
    }
    __syncthreads();
    // ---- phase 2: key ti (frame key, and cond key ti when conditioned)
    if (live) {
      const int j = ti, h = th, p = tp;
      float dk[BDH], dv[BDH];
#pragma unroll
      for (int k = 0; k < BDH; ++k) dk[k] = dv[k] = 0.f;
      const int col = cond ? NF + j : j;
#pragma unroll 1
      for (int i = 0; i < NF; ++i) {
        const float ds = Ss[((p * heads + h) * NF + i) * NK2 + col];
        const float pr = Ps[((p * heads + h) * NF + i) * NK2 + col];
        const float* qr = Qs + (p * NF + i) * HS + h * BDHP;
        const float* dr = Ds + (p * NF + i) * HS + h * BDHP;
#pragma unroll
        for (int k = 0; k < BDH; k += 4) {
          const float4 qq = *reinterpret_cast<const float4*>(qr + k);
          const float4 dd = *reinterpret_cast<const float4*>(dr + k);
          dk[k] += ds * qq.x;
          dk[k + 1] += ds * qq.y;
          dk[k + 2] += ds * qq.z;
          dk[k + 3] += ds * qq.w;
          dv[k] += pr * dd.x;
          dv[k + 1] += pr * dd.y;
          dv[k + 2] += pr * dd.z;
          dv[k + 3] += pr * dd.w;
        }
      }
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const float cs = RT[(j * 16 + k) * 2], sn = RT[(j * 16 + k) * 2 + 1];
        const float a = dk[2 * k], c = dk[2 * k + 1];
        dk[2 * k] = a * cs + c * sn;
        dk[2 * k + 1] = c * cs - a * sn;
      }
      const long long rowj = (static_cast<long long>(b) * NF + j) * HW + p0 + p;
      uint16_t* krow = dqkv + rowj * 3 * HD + HD + h * BDH;
      uint16_t* vrow = dqkv + rowj * 3 * HD + 2 * HD + h * BDH;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        bst8f(krow + k * 8, fmt, dk + k * 8);
        bst8f(vrow + k * 8, fmt, dv + k * 8);
      }
      if (cond) {
        float gek[BDH], gev[BDH];
#pragma unroll
        for (int k = 0; k < BDH; ++k) gek[k] = gev[k] = 0.f;
#pragma unroll 1
        for (int i = 0; i < NF; ++i) {
          const float ds = Ss[((p * heads + h) * NF + i) * NK2 + j];
          const float pr = Ps[((p * heads + h) * NF + i) * NK2 + j];
          const float* qr = Qs + (p * NF + i) * HS + h * BDHP;
          const float* dr = Ds + (p * NF + i) * HS + h * BDHP;
#pragma unroll
          for (int k = 0; k < BDH; k += 4) {
            const float4 qq = *reinterpret_cast<const float4*>(qr + k);
            const float4 dd = *reinterpret_cast<const float4*>(dr + k);
            gek[k] += ds * qq.x;
            gek[k + 1] += ds * qq.y;
            gek[k + 2] += ds * qq.z;
            gek[k + 3] += ds * qq.w;
            gev[k] += pr * dd.x;
            gev[k + 1] += pr * dd.y;
            gev[k + 2] += pr * dd.z;
            gev[k + 3] += pr * dd.w;
          }
        }
#pragma unroll
        for (int k = 0; k < BDH; ++k) {
          GA[k * nth + tid] += gek[k];
          GA[(BDH + k) * nth + tid] += gev[k];
        }
      }
    }
    __syncthreads();
  }
  if (tlive) {
    if (cond && dekv) {
      float* ge = dekv + (static_cast<long long>(b) * NF + ti) * 2 * HD + th * BDH;
      for (int k = 0; k < BDH; ++k) {
        atomicAdd(ge + k, GA[k * nth + tid]);
        atomicAdd(ge + HD + k, GA[(BDH + k) * nth + tid]);
      }
    }
    if (dbias) {
      for (int k = 0; k < NF; ++k) atomicAdd(dbias + (th * NF + ti) * NF + k, GB[k * nth + tid]);
    }
  }
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

extern "C" int vmm_tattn_bwd(const void* qkv, const float* ekv, const float* bias, const float* rot, const void* dout, void* dqkv,
                             float* dekv, float* dbias, int fmt, int B, int frames, int HW, int heads, float scale, void* stream_) {
  if (!qkv || !bias || !rot || !dout || !dqkv) return set_error(VMM_ERR_ARG, "vmm_tattn_bwd: null pointer");
  if (frames != 11) return set_error(VMM_ERR_UNSUPPORTED, "vmm_tattn_bwd: only 11 frames");
  if (heads < 1 || heads > 8) return set_error(VMM_ERR_UNSUPPORTED, "vmm_tattn_bwd: heads must be <= 8");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int NF = 11;
  int PX = 192 / (heads * NF);
  if (PX < 1) PX = 1;
  if (PX > 2) PX = 2;
  const int HS = heads * BDHP;
  const int nthreads = PX * heads * NF;
  const size_t smem = (static_cast<size_t>(4) * PX * NF * HS + static_cast<size_t>(2) * PX * heads * NF * 2 * NF + 2 * NF * HS + NF * 32 +
                       heads * NF * NF + static_cast<size_t>(64 + NF) * nthreads) * sizeof(float);
  if (smem > 227 * 1024) return set_error(VMM_ERR_UNSUPPORTED, "vmm_tattn_bwd: shared memory");
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(tattn_bwd_kernel<11>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return set_cuda_error(e, "vmm_tattn_bwd: attr");
    attr = true;
  }
  const int groups = (HW + PX - 1) / PX;
  int cps = (2 * num_sms() + B - 1) / B;   // CTAs per sample
  if (cps > groups) cps = groups;
  tattn_bwd_kernel<11><<<dim3(cps, B), PX * heads * NF, smem, stream>>>(static_cast<const uint16_t*>(qkv), ekv, bias, rot,
                                                                         static_cast<const uint16_t*>(dout), static_cast<uint16_t*>(dqkv),
                                                                         dekv, dbias, fmt, HW, heads, scale, PX);
  count_launch();
  return check_launch("vmm_tattn_bwd");
}

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
