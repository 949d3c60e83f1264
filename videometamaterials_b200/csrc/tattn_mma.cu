// Temporal attention on warp-level tensor cores (mma.sync m16n8k16).  VDDP:425-535 via VDDP:615.
//
// One CTA = 8 warps = the 8 heads of one sample for a group of PXB pixels.  The CTA stages the qkv rows of its
// pixels (11 frames x 768 channels, fully coalesced) in shared memory; warp h then rotates the q / k slices of
// head h in place (rotary, q also scaled) and runs, per pixel,
//     S = Q K^T   (16 x 32 keys: [cond 0..10 | pad][frame 0..10 | pad])      8 MMAs
//     O = P V                                                               8 MMAs
// with the softmax on the accumulator fragments.  The conditioning keys / values of (sample, head) and the
// relative position bias are the same for every pixel and live in registers as ready-made B fragments.
#include "common.cuh"
#include "mma_sync.cuh"

namespace vmm {

constexpr int TNF = 11;          // frames == cond tokens (VDDP:603)
constexpr int TPITCH = 776;      // smem row pitch in elements: 768 + 8 keeps ldmatrix rows on distinct banks
constexpr int TPXB = 2;          // pixels per pipeline stage (two stages in flight)

template <int FMT>
__global__ void __launch_bounds__(256) tattn_fwd_mma_kernel(const uint16_t* __restrict__ qkv, const float* __restrict__ ekv,
                                                            const float* __restrict__ bias, const float* __restrict__ rot,
                                                            uint16_t* __restrict__ out, int HW, int heads, float scale, int pre_rotated) {
  pdl_trigger();
  extern __shared__ __align__(16) uint16_t tsm[];
  uint16_t* tile0 = tsm;                                  // [2 stages][TPXB][TNF][TPITCH]
  uint16_t* zrow = tile0 + 2 * TPXB * TNF * TPITCH;       // one row of zeros (frames >= 11)
  float* RT = reinterpret_cast<float*>(zrow + TPITCH);    // [TNF][16][2]
  const int HD = heads * 32;
  const int b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int h = warp;
  const bool cond = ekv != nullptr;

  for (int i = tid; i < TPITCH; i += 256) zrow[i] = 0;
  for (int i = tid; i < TNF * 32; i += 256) RT[i] = rot[i];

  // ---- per-warp constants: cond key / value fragments, bias (+ key mask)
  uint32_t kc[2][2][2];   // [key tile][k-step][b0b1, b2b3]   B[k = d][n = cond key]
  uint32_t vc[4][2];      // [d tile][b0b1, b2b3]             B[k = cond key][n = d]
  float bs[2][2][2];      // [row half (g, g+8)][key tile][col 2t, 2t+1]  (-inf on padded keys)
  if (h < heads) {
    if (cond) {
      const float* eb = ekv + static_cast<long long>(b) * TNF * 2 * HD;
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int key = 8 * nt + g;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int d = 16 * ks + 8 * hf + 2 * t;
            float a0 = 0.f, a1 = 0.f;
            if (key < TNF) {
              a0 = eb[key * 2 * HD + h * 32 + d];
              a1 = eb[key * 2 * HD + h * 32 + d + 1];
            }
            kc[nt][ks][hf] = pack2<FMT>(a0, a1);
          }
        }
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int d = 8 * nt + g;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int k0 = 2 * t + 8 * hf;
          const float a0 = (k0 < TNF) ? eb[k0 * 2 * HD + HD + h * 32 + d] : 0.f;
          const float a1 = (k0 + 1 < TNF) ? eb[(k0 + 1) * 2 * HD + HD + h * 32 + d] : 0.f;
          vc[nt][hf] = pack2<FMT>(a0, a1);
        }
      }
    }
#pragma unroll
    for (int rh = 0; rh < 2; ++rh) {
      const int i = g + 8 * rh;
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int j = 8 * nt + 2 * t + c;
          bs[rh][nt][c] = (j < TNF) ? ((i < TNF) ? bias[(h * TNF + i) * TNF + j] : 0.f) : -1e30f;
        }
    }
  }
  __syncthreads();

  const uint32_t zrow_s = smem_u32(zrow);
  const int groups = (HW + TPXB - 1) / TPXB;
  // stage = TPXB x 11 rows of 1536 bytes, copied with cp.async (zero fill past the last pixel); the copy of group
  // i+1 overlaps the MMAs of group i
  auto stage_load = [&](int st, int grp) {
    const int p0 = grp * TPXB;
    const uint32_t dst0 = smem_u32(tile0 + st * TPXB * TNF * TPITCH);
    for (int i = tid; i < TPXB * TNF * 96; i += 256) {
      const int c = i % 96;
      const int f = (i / 96) % TNF;
      const int p = i / (96 * TNF);
      const bool ok = p0 + p < HW;
      const uint16_t* src = qkv + ((static_cast<long long>(b) * TNF + f) * HW + (ok ? p0 + p : 0)) * 3 * HD + c * 8;
      cp_async16(dst0 + static_cast<uint32_t>((p * TNF + f) * TPITCH + c * 8) * 2, src, ok ? 16 : 0);
    }
  };
  if (static_cast<int>(blockIdx.x) < groups) stage_load(0, blockIdx.x);
  cp_async_commit();
  int it = 0;
  for (int grp = blockIdx.x; grp < groups; grp += gridDim.x, ++it) {
    const int p0 = grp * TPXB;
    if (grp + static_cast<int>(gridDim.x) < groups) stage_load((it + 1) & 1, grp + gridDim.x);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    uint16_t* tile = tile0 + (it & 1) * TPXB * TNF * TPITCH;
    const uint32_t tile_s = smem_u32(tile);
    if (h < heads) {
      // ---- rotary in place on this head's q (scaled) and k slices (skipped when the to_qkv epilogue already did it)
      static_assert(TPXB == 2, "lane -> (pixel, pair) mapping below assumes two pixels per stage");
#pragma unroll
      for (int f = 0; f < TNF; ++f) {
        if (pre_rotated) break;
        const int k = lane & 15, p = lane >> 4;
        const float2 cssn = *reinterpret_cast<const float2*>(&RT[(f * 16 + k) * 2]);
        const float cs = cssn.x, sn = cssn.y;
        uint32_t* qp = reinterpret_cast<uint32_t*>(tile + (p * TNF + f) * TPITCH + h * 32 + 2 * k);
        uint32_t* kp = reinterpret_cast<uint32_t*>(tile + (p * TNF + f) * TPITCH + HD + h * 32 + 2 * k);
        float2 q = unpack2<FMT>(*qp), kk = unpack2<FMT>(*kp);
        q.x *= scale;
        q.y *= scale;
        *qp = pack2<FMT>(q.x * cs - q.y * sn, q.y * cs + q.x * sn);
        *kp = pack2<FMT>(kk.x * cs - kk.y * sn, kk.y * cs + kk.x * sn);
      }
      __syncwarp();
      for (int p = 0; p < TPXB; ++p) {
        if (p0 + p >= HW) break;
        const uint32_t pbase = tile_s + static_cast<uint32_t>(p * TNF * TPITCH) * 2;
        // lane -> row address helpers
        const int lm = lane >> 3, lr = lane & 7;
        float S[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int c = 0; c < 4; ++c) S[nt][c] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          uint32_t qa[4];
          {   // A tile: rows = frames 0..15, cols = d 16ks..16ks+15
            const int row = lr + 8 * (lm & 1);
            const int col = h * 32 + 16 * ks + 8 * (lm >> 1);
            const uint32_t addr = (row < TNF) ? pbase + static_cast<uint32_t>(row * TPITCH + col) * 2 : zrow_s + static_cast<uint32_t>(col) * 2;
            ldsm_x4(qa, addr);
          }
          if (cond) {
            mma16816<FMT>(S[0], qa, kc[0][ks]);
            mma16816<FMT>(S[1], qa, kc[1][ks]);
          }
          uint32_t kb[4];
          {   // B tiles: keys 0-7 / 8-15, d halves
            const int row = lr + 8 * (lm >> 1);
            const int col = HD + h * 32 + 16 * ks + 8 * (lm & 1);
            const uint32_t addr = (row < TNF) ? pbase + static_cast<uint32_t>(row * TPITCH + col) * 2 : zrow_s + static_cast<uint32_t>(col) * 2;
            ldsm_x4(kb, addr);
          }
          mma16816<FMT>(S[2], qa, kb);
          mma16816<FMT>(S[3], qa, kb + 2);
        }
        // ---- bias, mask, softmax over the 32 key slots (rows g and g+8)
        float mx[2] = {-1e30f, -1e30f}, sum[2] = {0.f, 0.f};
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          if (!cond && nt < 2) continue;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            S[nt][c] += bs[c >> 1][nt & 1][c & 1];
            mx[c >> 1] = fmaxf(mx[c >> 1], S[nt][c]);
          }
        }
#pragma unroll
        for (int rh = 0; rh < 2; ++rh) {
          mx[rh] = fmaxf(mx[rh], __shfl_xor_sync(0xffffffffu, mx[rh], 1));
          mx[rh] = fmaxf(mx[rh], __shfl_xor_sync(0xffffffffu, mx[rh], 2));
        }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float e = (!cond && nt < 2) ? 0.f : __expf(S[nt][c] - mx[c >> 1]);
            S[nt][c] = e;
            sum[c >> 1] += e;
          }
        }
#pragma unroll
        for (int rh = 0; rh < 2; ++rh) {
          sum[rh] += __shfl_xor_sync(0xffffffffu, sum[rh], 1);
          sum[rh] += __shfl_xor_sync(0xffffffffu, sum[rh], 2);
        }
        // ---- O = P V
        float O[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int c = 0; c < 4; ++c) O[nt][c] = 0.f;
        if (cond) {
          uint32_t pa[4] = {pack2<FMT>(S[0][0], S[0][1]), pack2<FMT>(S[0][2], S[0][3]), pack2<FMT>(S[1][0], S[1][1]), pack2<FMT>(S[1][2], S[1][3])};
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) mma16816<FMT>(O[nt], pa, vc[nt]);
        }
        {
          uint32_t pa[4] = {pack2<FMT>(S[2][0], S[2][1]), pack2<FMT>(S[2][2], S[2][3]), pack2<FMT>(S[3][0], S[3][1]), pack2<FMT>(S[3][2], S[3][3])};
#pragma unroll
          for (int dh = 0; dh < 2; ++dh) {
            uint32_t vb[4];
            const int row = lr + 8 * (lm & 1);
            const int col = 2 * HD + h * 32 + 16 * dh + 8 * (lm >> 1);
            const uint32_t addr = (row < TNF) ? pbase + static_cast<uint32_t>(row * TPITCH + col) * 2 : zrow_s + static_cast<uint32_t>(col) * 2;
            ldsm_x4_trans(vb, addr);
            mma16816<FMT>(O[2 * dh], pa, vb);
            mma16816<FMT>(O[2 * dh + 1], pa, vb + 2);
          }
        }
        // ---- normalise and store rows g (< 8 <= 11) and g + 8 (< 11)
        const float inv0 = 1.f / sum[0], inv1 = 1.f / sum[1];
        uint16_t* o0 = out + ((static_cast<long long>(b) * TNF + g) * HW + p0 + p) * HD + h * 32 + 2 * t;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) *reinterpret_cast<uint32_t*>(o0 + 8 * nt) = pack2<FMT>(O[nt][0] * inv0, O[nt][1] * inv0);
        if (g + 8 < TNF) {
          uint16_t* o1 = out + ((static_cast<long long>(b) * TNF + g + 8) * HW + p0 + p) * HD + h * 32 + 2 * t;
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) *reinterpret_cast<uint32_t*>(o1 + 8 * nt) = pack2<FMT>(O[nt][2] * inv1, O[nt][3] * inv1);
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace vmm

using namespace vmm;

extern "C" int vmm_tattn_fwd(const void* qkv, const float* ekv, const float* bias, const float* rot, void* out, int fmt, int B,
                             int frames, int HW, int heads, float scale, int pre_rotated, void* stream_) {
  if (!qkv || !bias || !rot || !out) return set_error(VMM_ERR_ARG, "vmm_tattn_fwd: null pointer");
  if (heads != 8) return set_error(VMM_ERR_UNSUPPORTED, "vmm_tattn_fwd: heads must be 8 (one warp per head, rows of 3*8*32 channels)");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  // the reference's shipped configuration is 11 frames (VDDP:603); other frame counts run on the generic kernels of tattn_generic.cu
  if (frames != TNF) return tattn_generic_fwd(qkv, ekv, bias, rot, out, fmt, B, frames, HW, scale, pre_rotated, stream);
  const size_t smem = (static_cast<size_t>(2) * TPXB * TNF * TPITCH + TPITCH) * sizeof(uint16_t) + TNF * 32 * sizeof(float);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(tattn_fwd_mma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tattn_fwd_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) return set_cuda_error(e, "vmm_tattn_fwd: attr");
    attr = true;
  }
  const int groups = (HW + TPXB - 1) / TPXB;
  int gx = (3 * num_sms() + B - 1) / B;     // ~3 resident CTAs per SM, each walking several pixel groups
  if (gx > groups) gx = groups;
  dim3 grid(gx, B);
  if (fmt == VMM_FMT_F16)
    tattn_fwd_mma_kernel<0><<<grid, 256, smem, stream>>>(static_cast<const uint16_t*>(qkv), ekv, bias, rot, static_cast<uint16_t*>(out), HW, heads, scale, pre_rotated);
  else
    tattn_fwd_mma_kernel<1><<<grid, 256, smem, stream>>>(static_cast<const uint16_t*>(qkv), ekv, bias, rot, static_cast<uint16_t*>(out), HW, heads, scale, pre_rotated);
  count_launch();
  return check_launch("vmm_tattn_fwd");
}

// ================================================================================================
// Backward.  Same CTA shape (8 warps = 8 heads, TPXB pixels per iteration).  Per pixel and head:
//   pass A (rows = queries): S, P, dP = dO V^T, D = rowsum(P dP), dS = P (dP - D), dQ = dS K      -> dq, log-sum-exp / D
//   pass B (rows = keys, once for the 16 cond slots, once for the 16 frame slots):
//            P^T and dS^T = the 16-bit P / dS fragments of pass A transposed in registers (movmatrix, 8 x 8 blocks),
//            dK = dS^T Q, dV = P^T dO                                                            -> dk, dv | cond grads
// Gradients of the cond keys / values and of the position bias are summed over the CTA's pixels in accumulator
// fragments and added atomically once per CTA.  The cond keys / values sit in shared memory as a 16-bit tile so that
// every operand fragment comes from ldmatrix.
// ================================================================================================
namespace vmm {

constexpr int DPITCH = 264;      // dO rows: 256 + 8
constexpr int CPITCH = 520;      // cond rows: ek(256) | ev(256) + 8
constexpr int BPXB = 2;          // pixels per pipeline stage

struct FragAddr {
  uint32_t zrow;
  int lm, lr;
  // A operand (16 x 16, rows = tile rows): matrices (r0-7,c0-7) (r8-15,c0-7) (r0-7,c8-15) (r8-15,c8-15)
  __device__ __forceinline__ uint32_t a(uint32_t base, int pitch, int col) const {
    const int row = lr + 8 * (lm & 1), c = col + 8 * (lm >> 1);
    return (row < TNF) ? base + static_cast<uint32_t>(row * pitch + c) * 2 : zrow + static_cast<uint32_t>(c) * 2;
  }
  // B operand, k contiguous in memory (rows = n index): two n-tiles x (k lo, k hi)
  __device__ __forceinline__ uint32_t b(uint32_t base, int pitch, int col) const {
    const int row = lr + 8 * (lm >> 1), c = col + 8 * (lm & 1);
    return (row < TNF) ? base + static_cast<uint32_t>(row * pitch + c) * 2 : zrow + static_cast<uint32_t>(c) * 2;
  }
  // B operand, n contiguous in memory (rows = k index), loaded with .trans: two n-tiles (cols) x (k lo, k hi)
  __device__ __forceinline__ uint32_t bt(uint32_t base, int pitch, int col) const { return a(base, pitch, col); }
};

template <int FMT>
__global__ void __launch_bounds__(256, 2) tattn_bwd_mma_kernel(const uint16_t* __restrict__ qkv, const float* __restrict__ ekv,
                                                               const float* __restrict__ bias, const float* __restrict__ rot,
                                                               const uint16_t* __restrict__ dout, uint16_t* __restrict__ dqkv,
                                                               float* __restrict__ dekv, float* __restrict__ dbias, int HW, int heads,
                                                               float scale, int pre_rotated) {
  pdl_trigger();
  extern __shared__ __align__(16) uint16_t tsm[];
  uint16_t* tile0 = tsm;                                    // [2][BPXB][TNF][TPITCH]   q | k | v   (q, k rotated in place)
  uint16_t* dtile0 = tile0 + 2 * BPXB * TNF * TPITCH;       // [2][BPXB][TNF][DPITCH]   dO
  uint16_t* ctile = dtile0 + 2 * BPXB * TNF * DPITCH;       // [TNF][CPITCH]            cond ek | ev
  uint16_t* zrow = ctile + TNF * CPITCH;                    // [TPITCH] zeros
  float* RT = reinterpret_cast<float*>(zrow + TPITCH);      // [TNF][16][2]

  constexpr int HD = 256;                                   // heads == 8 (host check)
  const int b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int h = warp;
  const bool cond = ekv != nullptr;

  for (int i = tid; i < TPITCH; i += 256) zrow[i] = 0;
  for (int i = tid; i < TNF * 32; i += 256) RT[i] = rot[i];
  // (cos, sin) of the rows / pairs a lane un-rotates, laid out [nt][row half][lane] so that a warp reads 32 consecutive float2:
  // indexing the [frame][pair] table directly put the 8 rows of a quad column on one bank (8-way conflicts on 32 loads per pixel)
  float2* RL = reinterpret_cast<float2*>(RT + TNF * 32);
  {
    const int l = tid & 31, rh = (tid >> 5) & 1, nt = tid >> 6;
    const int i = min((l >> 2) + 8 * rh, TNF - 1), k = 4 * nt + (l & 3);
    RL[tid] = make_float2(rot[(i * 16 + k) * 2], rot[(i * 16 + k) * 2 + 1]);
  }
  if (cond) {
    for (int i = tid; i < TNF * 2 * HD; i += 256) {
      const int j = i / (2 * HD), c = i % (2 * HD);
      const float v = ekv[(static_cast<long long>(b) * TNF + j) * 2 * HD + c];
      ctile[j * CPITCH + c] = FMT ? __bfloat16_as_ushort(__float2bfloat16_rn(v)) : __half_as_ushort(__float2half_rn(v));
    }
  }
  float bs[2][2][2];
#pragma unroll
  for (int rh = 0; rh < 2; ++rh)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int i = g + 8 * rh, j = 8 * nt + 2 * t + c;                    // pass A: row = query i, col = key j
        bs[rh][nt][c] = (j < TNF) ? ((i < TNF) ? bias[(h * TNF + i) * TNF + j] : 0.f) : -1e30f;
      }
  float gb[2][2][2];
  float gEK[4][4], gEV[4][4];
#pragma unroll
  for (int x = 0; x < 8; ++x) (&gb[0][0][0])[x] = 0.f;
#pragma unroll
  for (int x = 0; x < 16; ++x) (&gEK[0][0])[x] = (&gEV[0][0])[x] = 0.f;
  __syncthreads();

  // ldmatrix row addresses = (per-pixel tile base or the zero row) + lane-constant offset + compile-time column offset:
  //   A pattern (also B with .trans): matrices (r0-7, c0-7) (r8-15, c0-7) (r0-7, c8-15) (r8-15, c8-15)
  //   B pattern, k contiguous (rows = n index): two n-tiles x (k lo, k hi)
  const int lm = lane >> 3, lr = lane & 7;
  const int rowA = lr + 8 * (lm & 1), cA = 8 * (lm >> 1);
  const int rowB = lr + 8 * (lm >> 1), cB = 8 * (lm & 1);
  const bool vA = rowA < TNF, vB = rowB < TNF;
  const uint32_t zA = smem_u32(zrow) + static_cast<uint32_t>(cA + h * 32) * 2, zB = smem_u32(zrow) + static_cast<uint32_t>(cB + h * 32) * 2;
  const uint32_t otA = static_cast<uint32_t>(rowA * TPITCH + cA + h * 32) * 2, otB = static_cast<uint32_t>(rowB * TPITCH + cB + h * 32) * 2;
  const uint32_t odA = static_cast<uint32_t>(rowA * DPITCH + cA + h * 32) * 2;
  const uint32_t ctile_s = smem_u32(ctile);
  const uint32_t cAb = vA ? ctile_s + static_cast<uint32_t>(rowA * CPITCH + cA + h * 32) * 2 : zA;    // cond tile, A pattern
  const uint32_t cBb = vB ? ctile_s + static_cast<uint32_t>(rowB * CPITCH + cB + h * 32) * 2 : zB;    // cond tile, B pattern
  const int groups = (HW + BPXB - 1) / BPXB;
  // Warp-private pipeline: warp h stages, transforms and copies out ONLY the 32 columns of its head in the q | k | v | dO rows
  // (4 x 64 bytes per row), so the main loop needs no block barrier at all: the warps drift apart and hide each other's
  // ldmatrix / mma / shuffle latencies (ncu before: 11 % of the warp samples waiting at the two block barriers of a stage).
  // lane -> (segment q | k | v | dO, 16-byte piece, row parity); 22 rows per stage = 11 copies per lane, each with its own address
  // registers (LDGSTS reads them late: a pointer bumped right after the copy stalls on the write-after-read hazard).
  const int sseg = (lane & 15) >> 2, sv4 = lane & 3, shi = lane >> 4;
  const bool s_q = sseg < 3;                                    // q | k | v segment; else dO
  const uint16_t* s_src = s_q ? qkv + sseg * HD + h * 32 + sv4 * 8 : dout + h * 32 + sv4 * 8;
  const long long s_ld = s_q ? 3 * HD : HD;                     // row length of the source
  const uint32_t s_pitch2 = (s_q ? TPITCH : DPITCH) * 2;        // row pitch of the destination in bytes
  const uint32_t s_dst0 = (s_q ? smem_u32(tile0) + static_cast<uint32_t>(sseg * HD) * 2 : smem_u32(dtile0)) + static_cast<uint32_t>(h * 32 + sv4 * 8) * 2;
  const uint32_t s_stage = (s_q ? BPXB * TNF * TPITCH : BPXB * TNF * DPITCH) * 2;
  auto stage_load = [&](int st, int grp) {
    const int p0 = grp * BPXB;
#pragma unroll
    for (int p = 0; p < BPXB; ++p) {
      const bool ok = p0 + p < HW;
      const int fs = ((p * TNF) & 1) ? 1 - shi : shi;            // alternate the frame parity per pixel: 11 rows per lane and stage
      const uint16_t* src = s_src + ((static_cast<long long>(b) * TNF + fs) * HW + (ok ? p0 + p : 0)) * s_ld;
      const uint32_t dst = s_dst0 + st * s_stage + static_cast<uint32_t>(p * TNF + fs) * s_pitch2;
      const long long sstep = 2LL * HW * s_ld;
#pragma unroll
      for (int k = 0; k < (TNF + 1) / 2; ++k)
        if (fs + 2 * k < TNF) cp_async16(dst + k * 2 * s_pitch2, src + k * sstep, ok ? 16 : 0);
    }
  };
  if (static_cast<int>(blockIdx.x) < groups) stage_load(0, blockIdx.x);
  cp_async_commit();
  int it = 0;
  for (int grp = blockIdx.x; grp < groups; grp += gridDim.x, ++it) {
    const int p0 = grp * BPXB;
    if (grp + static_cast<int>(gridDim.x) < groups) stage_load((it + 1) & 1, grp + gridDim.x);
    cp_async_commit();
    cp_async_wait<1>();
    __syncwarp();          // the lanes of this warp see each other's copies; no other warp touches these columns
    uint16_t* tile = tile0 + (it & 1) * BPXB * TNF * TPITCH;
    const uint32_t tile_s = smem_u32(tile), dtile_s = smem_u32(dtile0 + (it & 1) * BPXB * TNF * DPITCH);
    // rotary in place (q scaled); skipped when the to_qkv epilogue already did it
    static_assert(BPXB == 2, "lane -> (pixel, pair) mapping below assumes two pixels per stage");
#pragma unroll
    for (int f = 0; f < TNF; ++f) {
        if (pre_rotated) break;
      const int k = lane & 15, p = lane >> 4;
      const float2 cssn = *reinterpret_cast<const float2*>(&RT[(f * 16 + k) * 2]);
      const float cs = cssn.x, sn = cssn.y;
      uint32_t* qp = reinterpret_cast<uint32_t*>(tile + (p * TNF + f) * TPITCH + h * 32 + 2 * k);
      uint32_t* kp = reinterpret_cast<uint32_t*>(tile + (p * TNF + f) * TPITCH + HD + h * 32 + 2 * k);
      float2 q = unpack2<FMT>(*qp), kk = unpack2<FMT>(*kp);
      q.x *= scale;
      q.y *= scale;
      *qp = pack2<FMT>(q.x * cs - q.y * sn, q.y * cs + q.x * sn);
      *kp = pack2<FMT>(kk.x * cs - kk.y * sn, kk.y * cs + kk.x * sn);
    }
    __syncwarp();
    for (int p = 0; p < BPXB; ++p) {
      if (p0 + p >= HW) break;
      const uint32_t tp = tile_s + static_cast<uint32_t>(p * TNF * TPITCH) * 2;
      const uint32_t dp_s = dtile_s + static_cast<uint32_t>(p * TNF * DPITCH) * 2;
      const uint32_t tA = vA ? tp + otA : zA, tB = vB ? tp + otB : zB, dA = vA ? dp_s + odA : zA;
      // =========================== pass A: rows = queries
      float S[4][4], dP[4][4];
#pragma unroll
      for (int x = 0; x < 16; ++x) (&S[0][0])[x] = (&dP[0][0])[x] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        uint32_t qa[4], da[4], kb[4], vb[4];
        ldsm_x4(qa, tA + 32 * ks);
        ldsm_x4(da, dA + 32 * ks);
        if (cond) {
          ldsm_x4(kb, cBb + 32 * ks);
          ldsm_x4(vb, cBb + 2 * HD + 32 * ks);
          mma16816<FMT>(S[0], qa, kb);
          mma16816<FMT>(S[1], qa, kb + 2);
          mma16816<FMT>(dP[0], da, vb);
          mma16816<FMT>(dP[1], da, vb + 2);
        }
        ldsm_x4(kb, tB + 2 * HD + 32 * ks);
        ldsm_x4(vb, tB + 4 * HD + 32 * ks);
        mma16816<FMT>(S[2], qa, kb);
        mma16816<FMT>(S[3], qa, kb + 2);
        mma16816<FMT>(dP[2], da, vb);
        mma16816<FMT>(dP[3], da, vb + 2);
      }
      float mx[2] = {-1e30f, -1e30f}, sum[2] = {0.f, 0.f}, Dr[2] = {0.f, 0.f};
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        if (!cond && nt < 2) continue;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          S[nt][c] += bs[c >> 1][nt & 1][c & 1];
          mx[c >> 1] = fmaxf(mx[c >> 1], S[nt][c]);
        }
      }
#pragma unroll
      for (int rh = 0; rh < 2; ++rh) {
        mx[rh] = fmaxf(mx[rh], __shfl_xor_sync(0xffffffffu, mx[rh], 1));
        mx[rh] = fmaxf(mx[rh], __shfl_xor_sync(0xffffffffu, mx[rh], 2));
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float e = (!cond && nt < 2) ? 0.f : __expf(S[nt][c] - mx[c >> 1]);
          S[nt][c] = e;
          sum[c >> 1] += e;
        }
#pragma unroll
      for (int rh = 0; rh < 2; ++rh) {
        sum[rh] += __shfl_xor_sync(0xffffffffu, sum[rh], 1);
        sum[rh] += __shfl_xor_sync(0xffffffffu, sum[rh], 2);
      }
      const float inv[2] = {1.f / sum[0], 1.f / sum[1]};
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          S[nt][c] *= inv[c >> 1];                 // P
          Dr[c >> 1] += S[nt][c] * dP[nt][c];
        }
#pragma unroll
      for (int rh = 0; rh < 2; ++rh) {
        Dr[rh] += __shfl_xor_sync(0xffffffffu, Dr[rh], 1);
        Dr[rh] += __shfl_xor_sync(0xffffffffu, Dr[rh], 2);
      }
      // P as 16-bit fragments [key tile][query rows g | g + 8]: pass B transposes them instead of recomputing S^T
      uint32_t pk[4][2];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        pk[nt][0] = pack2<FMT>(S[nt][0], S[nt][1]);
        pk[nt][1] = pack2<FMT>(S[nt][2], S[nt][3]);
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float ds = S[nt][c] * (dP[nt][c] - Dr[c >> 1]);
          dP[nt][c] = ds;                          // dS
          gb[c >> 1][nt & 1][c & 1] += ds;
        }
      uint32_t sa[2][4];                         // dS as 16-bit A fragments per key tile (also transposed in pass B)
      uint32_t dqp[2][4];                        // dq of rows g | g + 8, written over the q slice once pass B no longer reads it
      {   // dQ_rot = dS K  (k-step 0: cond keys, k-step 1: frame keys)
        float dQ[4][4];
#pragma unroll
        for (int x = 0; x < 16; ++x) (&dQ[0][0])[x] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          sa[ks][0] = pack2<FMT>(dP[2 * ks][0], dP[2 * ks][1]);
          sa[ks][1] = pack2<FMT>(dP[2 * ks][2], dP[2 * ks][3]);
          sa[ks][2] = pack2<FMT>(dP[2 * ks + 1][0], dP[2 * ks + 1][1]);
          sa[ks][3] = pack2<FMT>(dP[2 * ks + 1][2], dP[2 * ks + 1][3]);
        }
#pragma unroll
        for (int dh = 0; dh < 2; ++dh) {
          uint32_t kb[4];
          if (cond) {
            ldsm_x4_trans(kb, cAb + 32 * dh);
            mma16816<FMT>(dQ[2 * dh], sa[0], kb);
            mma16816<FMT>(dQ[2 * dh + 1], sa[0], kb + 2);
          }
          ldsm_x4_trans(kb, tA + 2 * HD + 32 * dh);
          mma16816<FMT>(dQ[2 * dh], sa[1], kb);
          mma16816<FMT>(dQ[2 * dh + 1], sa[1], kb + 2);
        }
        // dq = scale R^T dQ_rot
#pragma unroll
        for (int rh = 0; rh < 2; ++rh) {
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            const float2 rl = RL[(nt * 2 + rh) * 32 + lane];
            const float cs = rl.x, sn = rl.y;
            const float a = dQ[nt][2 * rh], c = dQ[nt][2 * rh + 1];
            dqp[rh][nt] = pack2<FMT>((a * cs + c * sn) * scale, (c * cs - a * sn) * scale);
          }
        }
      }
      __syncwarp();
      // =========================== pass B: rows = keys (mt 0: cond slots, mt 1: frame slots)
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        if (mt == 0 && !cond) continue;
        // rows = keys of tile mt, k = queries: block (keys 8a.., queries 8b..) is the transpose of block (queries 8b.., keys 8a..)
        const uint32_t pa[4] = {movmatrix_trans(pk[2 * mt][0]), movmatrix_trans(pk[2 * mt + 1][0]), movmatrix_trans(pk[2 * mt][1]),
                                movmatrix_trans(pk[2 * mt + 1][1])};
        const uint32_t sat[4] = {movmatrix_trans(sa[mt][0]), movmatrix_trans(sa[mt][2]), movmatrix_trans(sa[mt][1]),
                                 movmatrix_trans(sa[mt][3])};
        float dK[4][4], dV[4][4];
#pragma unroll
        for (int x = 0; x < 16; ++x) (&dK[0][0])[x] = (&dV[0][0])[x] = 0.f;
#pragma unroll
        for (int dh = 0; dh < 2; ++dh) {
          uint32_t qb[4], db[4];
          ldsm_x4_trans(qb, tA + 32 * dh);
          ldsm_x4_trans(db, dA + 32 * dh);
          mma16816<FMT>(dK[2 * dh], sat, qb);
          mma16816<FMT>(dK[2 * dh + 1], sat, qb + 2);
          mma16816<FMT>(dV[2 * dh], pa, db);
          mma16816<FMT>(dV[2 * dh + 1], pa, db + 2);
        }
        if (mt == 0) {
#pragma unroll
          for (int x = 0; x < 16; ++x) {
            (&gEK[0][0])[x] += (&dK[0][0])[x];
            (&gEV[0][0])[x] += (&dV[0][0])[x];
          }
        } else {
          // Gradients go over the staged rows in place: K and V of this pixel are no longer read (dQ was taken in pass A), Q is read
          // by the dK MMAs just above, and the 32 columns of head h in each of the q | k | v segments belong to this warp alone.
          // The CTA then copies whole 1536-byte rows to dqkv with 16-byte stores (4-byte stores straight from the fragments cost
          // eight 16-byte pieces per warp instruction).
          __syncwarp();
#pragma unroll
          for (int rh = 0; rh < 2; ++rh) {
            const int j = g + 8 * rh;
            if (j < TNF) {
              uint16_t* qrow = tile + (p * TNF + j) * TPITCH + h * 32 + 2 * t;
#pragma unroll
              for (int nt = 0; nt < 4; ++nt) {
                const float2 rl = RL[(nt * 2 + rh) * 32 + lane];
                const float cs = rl.x, sn = rl.y;
                const float a = dK[nt][2 * rh], c = dK[nt][2 * rh + 1];
                *reinterpret_cast<uint32_t*>(qrow + 8 * nt) = dqp[rh][nt];
                *reinterpret_cast<uint32_t*>(qrow + HD + 8 * nt) = pack2<FMT>(a * cs + c * sn, c * cs - a * sn);
                *reinterpret_cast<uint32_t*>(qrow + 2 * HD + 8 * nt) = pack2<FMT>(dV[nt][2 * rh], dV[nt][2 * rh + 1]);
              }
            }
          }
        }
      }
      __syncwarp();
    }
    __syncwarp();
    // copy-out of this warp's gradient columns (dq | dk | dv now sit where q | k | v were): 3 x 64 bytes per row
    if (s_q) {
#pragma unroll
      for (int p = 0; p < BPXB; ++p) {
        if (p0 + p >= HW) break;
        const int fs = ((p * TNF) & 1) ? 1 - shi : shi;
        const uint16_t* srow = tile + (p * TNF + fs) * TPITCH + sseg * HD + h * 32 + sv4 * 8;
        uint16_t* drow = dqkv + ((static_cast<long long>(b) * TNF + fs) * HW + p0 + p) * 3 * HD + sseg * HD + h * 32 + sv4 * 8;
        const long long dstep = 2LL * HW * 3 * HD;
        uint4 v[(TNF + 1) / 2];
#pragma unroll
        for (int k = 0; k < (TNF + 1) / 2; ++k)
          if (fs + 2 * k < TNF) v[k] = *reinterpret_cast<const uint4*>(srow + k * 2 * TPITCH);
#pragma unroll
        for (int k = 0; k < (TNF + 1) / 2; ++k)
          if (fs + 2 * k < TNF) *reinterpret_cast<uint4*>(drow + k * dstep) = v[k];
      }
    }
    __syncwarp();          // the next iteration's prefetch refills this buffer
  }
  // ---- flush the per-CTA sums
  if (cond && dekv) {
#pragma unroll
    for (int rh = 0; rh < 2; ++rh) {
      const int j = g + 8 * rh;
      if (j < TNF) {
        float* ge = dekv + (static_cast<long long>(b) * TNF + j) * 2 * HD + h * 32 + 2 * t;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          atomicAdd(ge + 8 * nt, gEK[nt][2 * rh]);
          atomicAdd(ge + 8 * nt + 1, gEK[nt][2 * rh + 1]);
          atomicAdd(ge + HD + 8 * nt, gEV[nt][2 * rh]);
          atomicAdd(ge + HD + 8 * nt + 1, gEV[nt][2 * rh + 1]);
        }
      }
    }
  }
  if (dbias) {
#pragma unroll
    for (int rh = 0; rh < 2; ++rh)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int i = g + 8 * rh, j = 8 * nt + 2 * t + c;
          if (i < TNF && j < TNF) atomicAdd(dbias + (h * TNF + i) * TNF + j, gb[rh][nt][c]);
        }
  }
}

}  // namespace vmm

extern "C" int vmm_tattn_bwd(const void* qkv, const float* ekv, const float* bias, const float* rot, const void* dout, void* dqkv,
                             float* dekv, float* dbias, int fmt, int B, int frames, int HW, int heads, float scale, int pre_rotated,
                             void* stream_) {
  using namespace vmm;
  if (!qkv || !bias || !rot || !dout || !dqkv) return set_error(VMM_ERR_ARG, "vmm_tattn_bwd: null pointer");
  if (heads != 8) return set_error(VMM_ERR_UNSUPPORTED, "vmm_tattn_bwd: heads must be 8");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (frames != TNF) return tattn_generic_bwd(qkv, ekv, bias, rot, dout, dqkv, dekv, dbias, fmt, B, frames, HW, scale, pre_rotated, stream);
  const size_t smem = (static_cast<size_t>(2) * BPXB * TNF * (TPITCH + DPITCH) + TNF * CPITCH + TPITCH) * sizeof(uint16_t) +
                      (TNF * 32 + 512) * sizeof(float);      // rotary table + its per-lane copy (256 float2)
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(tattn_bwd_mma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tattn_bwd_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (e != cudaSuccess) return set_cuda_error(e, "vmm_tattn_bwd: attr");
    attr = true;
  }
  const int groups = (HW + BPXB - 1) / BPXB;
  int gx = (2 * num_sms() + B - 1) / B;
  if (gx > groups) gx = groups;
  dim3 grid(gx, B);
  if (fmt == VMM_FMT_F16)
    tattn_bwd_mma_kernel<0><<<grid, 256, smem, stream>>>(static_cast<const uint16_t*>(qkv), ekv, bias, rot, static_cast<const uint16_t*>(dout),
                                                         static_cast<uint16_t*>(dqkv), dekv, dbias, HW, heads, scale, pre_rotated);
  else
    tattn_bwd_mma_kernel<1><<<grid, 256, smem, stream>>>(static_cast<const uint16_t*>(qkv), ekv, bias, rot, static_cast<const uint16_t*>(dout),
                                                         static_cast<uint16_t*>(dqkv), dekv, dbias, HW, heads, scale, pre_rotated);
  count_launch();
  return check_launch("vmm_tattn_bwd");
}
