// Temporal attention for frame counts other than 11 (any frames <= 24): plain CUDA-core kernels behind vmm_tattn_fwd / vmm_tattn_bwd.
//
// The reference hard-codes 11 conditioning tokens (VDDP:603), so its shipped configuration only runs at 11 frames; the tensor-core
// kernels of tattn_mma.cu / ftattn.cu are built around that 11 x 22 problem.  BASELINE configs[4] also names 22 frames: there the
// model is built with tokens == frames and these kernels carry the attention core.  Same contract as the 11-frame
// kernels (VDDP:425-535 via VDDP:615): per pixel and head, queries = the `frames` positions of the pixel, keys = [cond tokens
// (ekv, already rotated, fp32) | the pixel's own frames], the relative position bias added to both halves, softmax over all keys.
//
// One CTA walks pixels; its 8 warps are the 8 heads.  Forward: lane = query, online softmax over the keys, key / value rows read
// as shared-memory broadcasts.  Backward: pass A (lane = query) writes the P and dS matrices of the head to shared memory and
// accumulates dq; pass B (lane = key) reads them back transposed for dk / dv (frame keys) and the cond-key / cond-value
// gradients (summed over the CTA's pixels in registers); the position-bias gradient is summed per thread over the dS matrices.
// Arithmetic in fp32 on 16-bit inputs, like the tensor-core kernels' accumulators.  Not a tuned path: ~10x the time per (pixel,
// head) of the 11-frame kernels.
#include "common.cuh"
#include "mma_sync.cuh"

namespace vmm {

constexpr int GMAXF = 24;        // frames (== cond tokens) supported
constexpr int GPITCH = 776;      // qkv row pitch in shared memory (elements)
constexpr int GDPITCH = 264;     // dO row pitch
constexpr int GCPITCH = 520;     // cond row pitch: ek (256) | ev (256) + 8
constexpr int GMP = 33;          // pitch (floats) of a row of the P / dS matrices: [key][query]

template <int FMT>
__device__ __forceinline__ void g_load32(const uint16_t* p, float* v) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 u = *reinterpret_cast<const uint4*>(p + 8 * j);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      const float2 f = unpack2<FMT>(w[x]);
      v[8 * j + 2 * x] = f.x;
      v[8 * j + 2 * x + 1] = f.y;
    }
  }
}

template <int FMT>
__device__ __forceinline__ float g_dot32(const uint16_t* p, const float* q) {
  float s[4] = {0.f, 0.f, 0.f, 0.f};     // four independent chains: with 8 warps per SM a single 32-deep FMA chain is latency-bound
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 u = *reinterpret_cast<const uint4*>(p + 8 * j);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      const float2 f = unpack2<FMT>(w[x]);
      s[x] = fmaf(f.x, q[8 * j + 2 * x], s[x]);
      s[x] = fmaf(f.y, q[8 * j + 2 * x + 1], s[x]);
    }
  }
  return (s[0] + s[1]) + (s[2] + s[3]);
}

template <int FMT>
__device__ __forceinline__ void g_axpy32(const uint16_t* p, float a, float* acc) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 u = *reinterpret_cast<const uint4*>(p + 8 * j);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      const float2 f = unpack2<FMT>(w[x]);
      acc[8 * j + 2 * x] = fmaf(a, f.x, acc[8 * j + 2 * x]);
      acc[8 * j + 2 * x + 1] = fmaf(a, f.y, acc[8 * j + 2 * x + 1]);
    }
  }
}

template <int FMT>
__device__ __forceinline__ void g_store32(uint16_t* p, const float* v) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    *reinterpret_cast<uint4*>(p + 8 * j) = make_uint4(pack2<FMT>(v[8 * j], v[8 * j + 1]), pack2<FMT>(v[8 * j + 2], v[8 * j + 3]),
                                                      pack2<FMT>(v[8 * j + 4], v[8 * j + 5]), pack2<FMT>(v[8 * j + 6], v[8 * j + 7]));
}

// stage the rows of one pixel: q | k | v (and dO) of every frame; optional rotary (+ query scale) in place
template <int FMT>
__device__ __forceinline__ void g_stage(uint16_t* tile, uint16_t* dtile, const uint16_t* __restrict__ qkv, const uint16_t* __restrict__ dout,
                                        const float* RT, long long b, int F, int HW, int px, float scale, int pre_rotated) {
  const int tid = threadIdx.x;
  const int per_row = dout ? 128 : 96;
  for (int i = tid; i < F * per_row; i += 256) {
    const int f = i / per_row, c = i - f * per_row;
    const long long row = (b * F + f) * HW + px;
    if (c < 96) *reinterpret_cast<uint4*>(tile + f * GPITCH + c * 8) = __ldg(reinterpret_cast<const uint4*>(qkv + row * 768) + c);
    else *reinterpret_cast<uint4*>(dtile + f * GDPITCH + (c - 96) * 8) = __ldg(reinterpret_cast<const uint4*>(dout + row * 256) + (c - 96));
  }
  __syncthreads();
  if (!pre_rotated) {
    for (int i = tid; i < F * 256; i += 256) {              // (frame, q | k, head, pair)
      const int f = i >> 8, r = i & 255;
      const int isk = r >> 7, pr = r & 127;                  // pr = head * 16 + pair
      const float cs = RT[(f * 16 + (pr & 15)) * 2], sn = RT[(f * 16 + (pr & 15)) * 2 + 1];
      uint32_t* p = reinterpret_cast<uint32_t*>(tile + f * GPITCH + isk * 256 + pr * 2);
      float2 v = unpack2<FMT>(*p);
      if (!isk) {
        v.x *= scale;
        v.y *= scale;
      }
      *p = pack2<FMT>(v.x * cs - v.y * sn, v.y * cs + v.x * sn);
    }
    __syncthreads();
  }
}

template <int FMT>
__global__ void __launch_bounds__(256) tattn_gen_fwd_kernel(const uint16_t* __restrict__ qkv, const float* __restrict__ ekv,
                                                            const float* __restrict__ bias, const float* __restrict__ rot,
                                                            uint16_t* __restrict__ out, int F, int HW, float scale, int pre_rotated) {
  extern __shared__ __align__(16) uint16_t gsm[];
  uint16_t* tile = gsm;                                   // [F][GPITCH]
  uint16_t* ctile = tile + F * GPITCH;                    // [F][GCPITCH]   cond ek | ev (16-bit)
  float* RT = reinterpret_cast<float*>(ctile + F * GCPITCH);   // [F][16][2]
  float* BS = RT + F * 32;                                // [8][F][F]
  const long long b = blockIdx.y;
  const int tid = threadIdx.x, h = tid >> 5, lane = tid & 31;
  const bool cond = ekv != nullptr;
  for (int i = tid; i < F * 32; i += 256) RT[i] = rot[i];
  for (int i = tid; i < 8 * F * F; i += 256) BS[i] = bias[i];
  if (cond) {
    for (int i = tid; i < F * 512; i += 256) {
      const int j = i >> 9, c = i & 511;
      ctile[j * GCPITCH + c] = f_to_h16(ekv[(b * F + j) * 512 + c], FMT);
    }
  }
  __syncthreads();
  for (int px = blockIdx.x; px < HW; px += gridDim.x) {
    g_stage<FMT>(tile, nullptr, qkv, nullptr, RT, b, F, HW, px, scale, pre_rotated);
    if (lane < F) {
      float q[32], o[32];
      g_load32<FMT>(tile + lane * GPITCH + h * 32, q);
#pragma unroll
      for (int x = 0; x < 32; ++x) o[x] = 0.f;
      float m = -1e30f, l = 0.f;
      const float* brow = BS + (h * F + lane) * F;
      for (int side = cond ? 0 : 1; side < 2; ++side) {
        const uint16_t* kbase = side ? tile + 256 + h * 32 : ctile + h * 32;
        const int kp = side ? GPITCH : GCPITCH;
        for (int j = 0; j < F; ++j) {
          const uint16_t* kr = kbase + j * kp;
          const float s = g_dot32<FMT>(kr, q) + brow[j];
          if (s > m) {                                      // rescale only when the running maximum moves (rare after the first keys)
            const float corr = __expf(m - s);
            l *= corr;
#pragma unroll
            for (int x = 0; x < 32; ++x) o[x] *= corr;
            m = s;
          }
          const float p = __expf(s - m);
          l += p;
          g_axpy32<FMT>(kr + 256, p, o);                   // v row sits 256 columns after the k row in both tiles
        }
      }
      const float inv = 1.f / l;
#pragma unroll
      for (int x = 0; x < 32; ++x) o[x] *= inv;
      g_store32<FMT>(out + ((b * F + lane) * HW + px) * 256 + h * 32, o);
    }
    __syncthreads();       // the next pixel's staging overwrites the tile
  }
}

template <int FMT>
__global__ void __launch_bounds__(256) tattn_gen_bwd_kernel(const uint16_t* __restrict__ qkv, const float* __restrict__ ekv,
                                                            const float* __restrict__ bias, const float* __restrict__ rot,
                                                            const uint16_t* __restrict__ dout, uint16_t* __restrict__ dqkv,
                                                            float* __restrict__ dekv, float* __restrict__ dbias, int F, int HW,
                                                            float scale, int pre_rotated) {
  extern __shared__ __align__(16) uint16_t gsm[];
  uint16_t* tile = gsm;                                   // [F][GPITCH]    q | k | v
  uint16_t* dtile = tile + F * GPITCH;                    // [F][GDPITCH]   dO
  uint16_t* ctile = dtile + F * GDPITCH;                  // [F][GCPITCH]   cond ek | ev
  float* RT = reinterpret_cast<float*>(ctile + F * GCPITCH);   // [F][16][2]
  float* BS = RT + F * 32;                                // [8][F][F]
  float* PM = BS + 8 * F * F;                             // [8 heads][2 F keys][GMP]   P, key-major
  float* DS = PM + 8 * 2 * F * GMP;                       // [8 heads][2 F keys][GMP]   dP, then dS
  const long long b = blockIdx.y;
  const int tid = threadIdx.x, h = tid >> 5, lane = tid & 31;
  const bool cond = ekv != nullptr;
  for (int i = tid; i < F * 32; i += 256) RT[i] = rot[i];
  for (int i = tid; i < 8 * F * F; i += 256) BS[i] = bias[i];
  for (int i = tid; i < 8 * 2 * F * GMP; i += 256) PM[i] = DS[i] = 0.f;      // the cond half stays zero without cond tokens
  if (cond) {
    for (int i = tid; i < F * 512; i += 256) {
      const int j = i >> 9, c = i & 511;
      ctile[j * GCPITCH + c] = f_to_h16(ekv[(b * F + j) * 512 + c], FMT);
    }
  }
  float gEK[32], gEV[32];                                  // lane = cond key: gradient rows summed over this CTA's pixels
#pragma unroll
  for (int x = 0; x < 32; ++x) gEK[x] = gEV[x] = 0.f;
  // position-bias gradient: thread owns elements e = tid + 256 n of the [8][F queries][F keys] table (both key halves add to it)
  constexpr int GBN = (8 * GMAXF * GMAXF + 255) / 256;
  float gb[GBN];
#pragma unroll
  for (int n = 0; n < GBN; ++n) gb[n] = 0.f;
  float* pm = PM + h * 2 * F * GMP;
  float* ds = DS + h * 2 * F * GMP;
  __syncthreads();
  for (int px = blockIdx.x; px < HW; px += gridDim.x) {
    g_stage<FMT>(tile, dtile, qkv, dout, RT, b, F, HW, px, scale, pre_rotated);
    // =========================== pass A: lane = query i
    if (lane < F) {
      float q[32], go[32];
      g_load32<FMT>(tile + lane * GPITCH + h * 32, q);
      g_load32<FMT>(dtile + lane * GDPITCH + h * 32, go);
      const float* brow = BS + (h * F + lane) * F;
      float m = -1e30f;
      for (int side = cond ? 0 : 1; side < 2; ++side) {
        const uint16_t* kbase = side ? tile + 256 + h * 32 : ctile + h * 32;
        const int kp = side ? GPITCH : GCPITCH;
        for (int j = 0; j < F; ++j) {
          const float s = g_dot32<FMT>(kbase + j * kp, q) + brow[j];
          pm[(side * F + j) * GMP + lane] = s;
          m = fmaxf(m, s);
        }
      }
      float l = 0.f;
      for (int side = cond ? 0 : 1; side < 2; ++side)
        for (int j = 0; j < F; ++j) {
          const float p = __expf(pm[(side * F + j) * GMP + lane] - m);
          pm[(side * F + j) * GMP + lane] = p;
          l += p;
        }
      const float inv = 1.f / l;
      float D = 0.f;
      for (int side = cond ? 0 : 1; side < 2; ++side) {
        const uint16_t* vbase = side ? tile + 512 + h * 32 : ctile + 256 + h * 32;
        const int kp = side ? GPITCH : GCPITCH;
        for (int j = 0; j < F; ++j) {
          const float p = pm[(side * F + j) * GMP + lane] * inv;
          const float dp = g_dot32<FMT>(vbase + j * kp, go);
          pm[(side * F + j) * GMP + lane] = p;
          ds[(side * F + j) * GMP + lane] = dp;
          D = fmaf(p, dp, D);
        }
      }
      float dq[32];
#pragma unroll
      for (int x = 0; x < 32; ++x) dq[x] = 0.f;
      for (int side = cond ? 0 : 1; side < 2; ++side) {
        const uint16_t* kbase = side ? tile + 256 + h * 32 : ctile + h * 32;
        const int kp = side ? GPITCH : GCPITCH;
        for (int j = 0; j < F; ++j) {
          const float g = pm[(side * F + j) * GMP + lane] * (ds[(side * F + j) * GMP + lane] - D);
          ds[(side * F + j) * GMP + lane] = g;
          g_axpy32<FMT>(kbase + j * kp, g, dq);
        }
      }
      // dq = scale R^T dQ_rot
      float r[32];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const float cs = RT[(lane * 16 + k) * 2], sn = RT[(lane * 16 + k) * 2 + 1];
        r[2 * k] = (dq[2 * k] * cs + dq[2 * k + 1] * sn) * scale;
        r[2 * k + 1] = (dq[2 * k + 1] * cs - dq[2 * k] * sn) * scale;
      }
      g_store32<FMT>(dqkv + ((b * F + lane) * HW + px) * 768 + h * 32, r);
    }
    __syncwarp();
    // =========================== pass B: lane = key j   (side 0: cond keys -> register sums, side 1: frame keys -> dk, dv rows)
    if (lane < F) {
      for (int side = cond ? 0 : 1; side < 2; ++side) {
        float dk[32], dv[32];
#pragma unroll
        for (int x = 0; x < 32; ++x) dk[x] = dv[x] = 0.f;
        const float* prow = pm + (side * F + lane) * GMP;
        const float* drow = ds + (side * F + lane) * GMP;
        for (int i = 0; i < F; ++i) {
          g_axpy32<FMT>(tile + i * GPITCH + h * 32, drow[i], dk);          // dk_j += dS_ij q_i
          g_axpy32<FMT>(dtile + i * GDPITCH + h * 32, prow[i], dv);        // dv_j += P_ij dO_i
        }
        if (side == 0) {
#pragma unroll
          for (int x = 0; x < 32; ++x) {
            gEK[x] += dk[x];
            gEV[x] += dv[x];
          }
        } else {
          float r[32];
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const float cs = RT[(lane * 16 + k) * 2], sn = RT[(lane * 16 + k) * 2 + 1];
            r[2 * k] = dk[2 * k] * cs + dk[2 * k + 1] * sn;
            r[2 * k + 1] = dk[2 * k + 1] * cs - dk[2 * k] * sn;
          }
          uint16_t* orow = dqkv + ((b * F + lane) * HW + px) * 768 + h * 32;
          g_store32<FMT>(orow + 256, r);
          g_store32<FMT>(orow + 512, dv);
        }
      }
    }
    __syncthreads();
    if (dbias) {
#pragma unroll
      for (int n = 0; n < GBN; ++n) {
        const int e = tid + 256 * n;
        if (e < 8 * F * F) {
          const int hh = e / (F * F), rem = e - hh * F * F, i = rem / F, j = rem - i * F;
          gb[n] += DS[(hh * 2 * F + j) * GMP + i] + DS[(hh * 2 * F + F + j) * GMP + i];
        }
      }
    }
    __syncthreads();       // the next pixel's staging and pass A overwrite the tile and the matrices
  }
  if (cond && dekv && lane < F) {
    float* ge = dekv + (b * F + lane) * 512 + h * 32;
#pragma unroll
    for (int x = 0; x < 32; ++x) {
      atomicAdd(ge + x, gEK[x]);
      atomicAdd(ge + 256 + x, gEV[x]);
    }
  }
  if (dbias) {
#pragma unroll
    for (int n = 0; n < GBN; ++n) {
      const int e = tid + 256 * n;
      if (e < 8 * F * F) atomicAdd(dbias + e, gb[n]);
    }
  }
}

int tattn_generic_fwd(const void* qkv, const float* ekv, const float* bias, const float* rot, void* out, int fmt, int B, int frames, int HW,
                      float scale, int pre_rotated, cudaStream_t stream) {
  if (frames < 1 || frames > GMAXF) return set_error(VMM_ERR_UNSUPPORTED, "vmm_tattn_fwd: frames must be 11 (tensor-core kernels) or at most 24 (generic kernels)");
  const size_t smem = static_cast<size_t>(frames) * (GPITCH + GCPITCH) * 2 + (static_cast<size_t>(frames) * 32 + 8 * frames * frames) * 4;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(tattn_gen_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tattn_gen_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) return set_cuda_error(e, "vmm_tattn_fwd: attr");
    attr = true;
  }
  int gx = (3 * num_sms() + B - 1) / B;
  if (gx > HW) gx = HW;
  const dim3 grid(gx, B);
  if (fmt == VMM_FMT_F16)
    tattn_gen_fwd_kernel<0><<<grid, 256, smem, stream>>>(static_cast<const uint16_t*>(qkv), ekv, bias, rot, static_cast<uint16_t*>(out), frames, HW, scale, pre_rotated);
  else
    tattn_gen_fwd_kernel<1><<<grid, 256, smem, stream>>>(static_cast<const uint16_t*>(qkv), ekv, bias, rot, static_cast<uint16_t*>(out), frames, HW, scale, pre_rotated);
  count_launch();
  return check_launch("vmm_tattn_fwd");
}

int tattn_generic_bwd(const void* qkv, const float* ekv, const float* bias, const float* rot, const void* dout, void* dqkv, float* dekv,
                      float* dbias, int fmt, int B, int frames, int HW, float scale, int pre_rotated, cudaStream_t stream) {
  if (frames < 1 || frames > GMAXF) return set_error(VMM_ERR_UNSUPPORTED, "vmm_tattn_bwd: frames must be 11 (tensor-core kernels) or at most 24 (generic kernels)");
  const size_t smem = static_cast<size_t>(frames) * (GPITCH + GDPITCH + GCPITCH) * 2 +
                      (static_cast<size_t>(frames) * 32 + 8 * frames * frames + 2 * 8 * 2 * frames * GMP) * 4;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(tattn_gen_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tattn_gen_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return set_cuda_error(e, "vmm_tattn_bwd: attr");
    attr = true;
  }
  if (smem > 220 * 1024) return set_error(VMM_ERR_UNSUPPORTED, "vmm_tattn_bwd: shared memory");
  int gx = (num_sms() + B - 1) / B;
  if (gx > HW) gx = HW;
  const dim3 grid(gx, B);
  if (fmt == VMM_FMT_F16)
    tattn_gen_bwd_kernel<0><<<grid, 256, smem, stream>>>(static_cast<const uint16_t*>(qkv), ekv, bias, rot, static_cast<const uint16_t*>(dout),
                                                         static_cast<uint16_t*>(dqkv), dekv, dbias, frames, HW, scale, pre_rotated);
  else
    tattn_gen_bwd_kernel<1><<<grid, 256, smem, stream>>>(static_cast<const uint16_t*>(qkv), ekv, bias, rot, static_cast<const uint16_t*>(dout),
                                                         static_cast<uint16_t*>(dqkv), dekv, dbias, frames, HW, scale, pre_rotated);
  count_launch();
  return check_launch("vmm_tattn_bwd");
}

}  // namespace vmm
