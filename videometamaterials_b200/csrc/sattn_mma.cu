// Quadratic spatial attention of the bottleneck (VDDP:687-689 via EinopsToAndFrom 'b f (h w) c') on warp-level tensor
// cores (mma.sync m16n8k16 + ldmatrix), flash-attention style.  Per frame-image bf and head: HW queries x (1 conditioning
// token + HW) keys, d = 32, no rotary, no bias; the token of frame f is key 0 (VDDP:459-462, 473-474).
//
//   forward : CTA = (head, frame-image); q, k, v of the head staged as 16-bit rows in shared memory; each warp takes 16-query
//             tiles and walks the keys in blocks of 64 with an online softmax: S = Q K^T (16 MMAs), P V (16 MMAs) per block.
//             Writes the normalised rows and the log-sum-exp per query.
//   backward: nothing of size queries x keys is kept.  Phase A (warp = 16-query tile): P = exp(S - lse), dP = dO V^T,
//             dS = P (dP - D) with D = rowsum(dO * O) from the saved output, dQ = scale dS K.  Phase B (warp = 16-key tile):
//             the transposed tiles S^T = K Q^T, dP^T = V dO^T are recomputed per 16 queries, dV += P^T dO, dK += scale dS^T Q;
//             key 0 goes to the conditioning-token gradient (this CTA owns its 32 columns: plain stores).
// Replaces the round-1 CUDA-core kernels (one query / key per thread, fp32 K / V in shared memory).
#include "common.cuh"
#include "mma_sync.cuh"

namespace vmm {

constexpr int SP = 40;   // shared-memory row pitch in 16-bit elements (32 + 8): ldmatrix rows land on distinct banks

struct SFrag {
  int lm, lr;
  // A operand 16 x 16: rows r0.., k = col..col+15
  __device__ __forceinline__ uint32_t a(uint32_t base, int r0, int col) const {
    return base + static_cast<uint32_t>((r0 + lr + 8 * (lm & 1)) * SP + col + 8 * (lm >> 1)) * 2;
  }
  // B operand, k contiguous in memory (rows = n index): two n-tiles (r0.., r0+8..) x (k lo, k hi)
  __device__ __forceinline__ uint32_t b(uint32_t base, int r0, int col) const {
    return base + static_cast<uint32_t>((r0 + lr + 8 * (lm >> 1)) * SP + col + 8 * (lm & 1)) * 2;
  }
  // B operand, n contiguous in memory (rows = k index), with ldsm_x4_trans: two n-tiles (cols) x (k lo, k hi)
  __device__ __forceinline__ uint32_t bt(uint32_t base, int r0, int col) const { return a(base, r0, col); }
};

// stage `rows` rows of 32 columns (16-bit) from a [row][ld] global matrix, zero-filling rows >= rows up to rows_pad
template <int FMT>
__device__ __forceinline__ void stage_rows(uint16_t* dst, const uint16_t* src, long long ld, int rows, int rows_pad, int row_shift) {
  // row_shift = 1: destination row r holds source row r - 1 (row 0 is filled by the caller: the conditioning token)
  for (int i = threadIdx.x; i < rows_pad * 4; i += blockDim.x) {
    const int r = i >> 2, c = i & 3;
    const int sr = r - row_shift;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (sr >= 0 && sr < rows) v = __ldg(reinterpret_cast<const uint4*>(src + static_cast<long long>(sr) * ld) + c);
    if (!(row_shift && r == 0)) *reinterpret_cast<uint4*>(dst + r * SP + c * 8) = v;
  }
}

template <int FMT>
__device__ __forceinline__ void stage_token(uint16_t* dst, const float* src) {
  if (threadIdx.x < 16) reinterpret_cast<uint32_t*>(dst)[threadIdx.x] = pack2<FMT>(src[2 * threadIdx.x], src[2 * threadIdx.x + 1]);
}

template <int FMT>
__global__ void __launch_bounds__(256) sattn_fwd_mma_kernel(const uint16_t* __restrict__ qkv, const float* __restrict__ ekv,
                                                            uint16_t* __restrict__ out, float* __restrict__ lse, int HW, int heads,
                                                            float scale) {
  extern __shared__ __align__(16) uint16_t ssm[];
  const int NK = HW + 1;
  const int HWp = (HW + 15) & ~15, NKp = (NK + 63) & ~63;
  uint16_t* Qs = ssm;                // [HWp][SP]
  uint16_t* Ks = Qs + HWp * SP;      // [NKp][SP]   row 0 = the frame's conditioning token
  uint16_t* Vs = Ks + NKp * SP;      // [NKp][SP]
  const int h = blockIdx.x, bf = blockIdx.y;
  const int HD = heads * 32;
  const uint16_t* base = qkv + static_cast<long long>(bf) * HW * 3 * HD + h * 32;
  stage_rows<FMT>(Qs, base, 3 * HD, HW, HWp, 0);
  stage_rows<FMT>(Ks, base + HD, 3 * HD, HW, NKp, 1);
  stage_rows<FMT>(Vs, base + 2 * HD, 3 * HD, HW, NKp, 1);
  stage_token<FMT>(Ks, ekv + static_cast<long long>(bf) * 2 * HD + h * 32);
  if (threadIdx.x >= 32 && threadIdx.x < 48)
    reinterpret_cast<uint32_t*>(Vs)[threadIdx.x - 32] = pack2<FMT>(ekv[static_cast<long long>(bf) * 2 * HD + HD + h * 32 + 2 * (threadIdx.x - 32)],
                                                                   ekv[static_cast<long long>(bf) * 2 * HD + HD + h * 32 + 2 * (threadIdx.x - 32) + 1]);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  SFrag fa;
  fa.lm = lane >> 3;
  fa.lr = lane & 7;
  const uint32_t q_s = smem_u32(Qs), k_s = smem_u32(Ks), v_s = smem_u32(Vs);
  for (int qt = warp; qt < HWp / 16; qt += 8) {
    uint32_t qa[2][4];
    ldsm_x4(qa[0], fa.a(q_s, qt * 16, 0));
    ldsm_x4(qa[1], fa.a(q_s, qt * 16, 16));
    float m[2] = {-1e30f, -1e30f}, l[2] = {0.f, 0.f};
    float O[4][4];
#pragma unroll
    for (int x = 0; x < 16; ++x) (&O[0][0])[x] = 0.f;
    for (int kb = 0; kb < NKp; kb += 64) {
      float S[8][4];
#pragma unroll
      for (int x = 0; x < 32; ++x) (&S[0][0])[x] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t kf[4];
          ldsm_x4(kf, fa.b(k_s, kb + np * 16, 16 * ks));
          mma16816<FMT>(S[2 * np], qa[ks], kf);
          mma16816<FMT>(S[2 * np + 1], qa[ks], kf + 2);
        }
      float mx[2] = {m[0], m[1]};
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int key = kb + nt * 8 + 2 * t + (c & 1);
          S[nt][c] = key < NK ? S[nt][c] * scale : -1e30f;
          mx[c >> 1] = fmaxf(mx[c >> 1], S[nt][c]);
        }
      float corr[2], sum[2] = {0.f, 0.f};
#pragma unroll
      for (int rh = 0; rh < 2; ++rh) {
        mx[rh] = fmaxf(mx[rh], __shfl_xor_sync(0xffffffffu, mx[rh], 1));
        mx[rh] = fmaxf(mx[rh], __shfl_xor_sync(0xffffffffu, mx[rh], 2));
        corr[rh] = __expf(m[rh] - mx[rh]);
        m[rh] = mx[rh];
      }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float e = __expf(S[nt][c] - m[c >> 1]);
          S[nt][c] = e;
          sum[c >> 1] += e;
        }
#pragma unroll
      for (int rh = 0; rh < 2; ++rh) {
        sum[rh] += __shfl_xor_sync(0xffffffffu, sum[rh], 1);
        sum[rh] += __shfl_xor_sync(0xffffffffu, sum[rh], 2);
        l[rh] = l[rh] * corr[rh] + sum[rh];
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        O[nt][0] *= corr[0];
        O[nt][1] *= corr[0];
        O[nt][2] *= corr[1];
        O[nt][3] *= corr[1];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t pa[4] = {pack2<FMT>(S[2 * j][0], S[2 * j][1]), pack2<FMT>(S[2 * j][2], S[2 * j][3]),
                                pack2<FMT>(S[2 * j + 1][0], S[2 * j + 1][1]), pack2<FMT>(S[2 * j + 1][2], S[2 * j + 1][3])};
#pragma unroll
        for (int dh = 0; dh < 2; ++dh) {
          uint32_t vf[4];
          ldsm_x4_trans(vf, fa.bt(v_s, kb + 16 * j, 16 * dh));
          mma16816<FMT>(O[2 * dh], pa, vf);
          mma16816<FMT>(O[2 * dh + 1], pa, vf + 2);
        }
      }
    }
#pragma unroll
    for (int rh = 0; rh < 2; ++rh) {
      const int n = qt * 16 + g + 8 * rh;
      if (n < HW) {
        const float inv = 1.f / l[rh];
        uint16_t* orow = out + (static_cast<long long>(bf) * HW + n) * HD + h * 32 + 2 * t;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) *reinterpret_cast<uint32_t*>(orow + 8 * nt) = pack2<FMT>(O[nt][2 * rh] * inv, O[nt][2 * rh + 1] * inv);
        if (lse && t == 0) lse[(static_cast<long long>(bf) * heads + h) * HW + n] = m[rh] + __logf(l[rh]);
      }
    }
  }
}

template <int FMT>
__global__ void __launch_bounds__(256) sattn_bwd_mma_kernel(const uint16_t* __restrict__ qkv, const float* __restrict__ ekv,
                                                            const uint16_t* __restrict__ aout, const uint16_t* __restrict__ dout,
                                                            const float* __restrict__ lse, uint16_t* __restrict__ dqkv,
                                                            float* __restrict__ dekv, int HW, int heads, float scale) {
  extern __shared__ __align__(16) uint16_t ssm[];
  const int NK = HW + 1;
  const int HWp = (HW + 15) & ~15, NKp = (NK + 63) & ~63;
  uint16_t* Qs = ssm;                // [HWp][SP]
  uint16_t* Ds = Qs + HWp * SP;      // [HWp][SP]  dO
  uint16_t* Ks = Ds + HWp * SP;      // [NKp][SP]
  uint16_t* Vs = Ks + NKp * SP;      // [NKp][SP]
  float* Ls = reinterpret_cast<float*>(Vs + NKp * SP);   // [HWp] log-sum-exp
  float* Dd = Ls + HWp;                                  // [HWp] rowsum(dO * O)
  const int h = blockIdx.x, bf = blockIdx.y;
  const int HD = heads * 32;
  const uint16_t* base = qkv + static_cast<long long>(bf) * HW * 3 * HD + h * 32;
  stage_rows<FMT>(Qs, base, 3 * HD, HW, HWp, 0);
  stage_rows<FMT>(Ds, dout + static_cast<long long>(bf) * HW * HD + h * 32, HD, HW, HWp, 0);
  stage_rows<FMT>(Ks, base + HD, 3 * HD, HW, NKp, 1);
  stage_rows<FMT>(Vs, base + 2 * HD, 3 * HD, HW, NKp, 1);
  stage_token<FMT>(Ks, ekv + static_cast<long long>(bf) * 2 * HD + h * 32);
  if (threadIdx.x >= 32 && threadIdx.x < 48)
    reinterpret_cast<uint32_t*>(Vs)[threadIdx.x - 32] = pack2<FMT>(ekv[static_cast<long long>(bf) * 2 * HD + HD + h * 32 + 2 * (threadIdx.x - 32)],
                                                                   ekv[static_cast<long long>(bf) * 2 * HD + HD + h * 32 + 2 * (threadIdx.x - 32) + 1]);
  for (int n = threadIdx.x; n < HWp; n += blockDim.x) {
    float dd = 0.f, ls = 0.f;
    if (n < HW) {
      const uint4* o4 = reinterpret_cast<const uint4*>(aout + (static_cast<long long>(bf) * HW + n) * HD + h * 32);
      const uint4* d4 = reinterpret_cast<const uint4*>(dout + (static_cast<long long>(bf) * HW + n) * HD + h * 32);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint4 ov = __ldg(o4 + c), dv = __ldg(d4 + c);
        const uint32_t ow[4] = {ov.x, ov.y, ov.z, ov.w}, dw[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 a = unpack2<FMT>(ow[j]), b2 = unpack2<FMT>(dw[j]);
          dd += a.x * b2.x + a.y * b2.y;
        }
      }
      ls = lse[(static_cast<long long>(bf) * heads + h) * HW + n];
    }
    Ls[n] = ls;
    Dd[n] = dd;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  SFrag fa;
  fa.lm = lane >> 3;
  fa.lr = lane & 7;
  const uint32_t q_s = smem_u32(Qs), d_s = smem_u32(Ds), k_s = smem_u32(Ks), v_s = smem_u32(Vs);

  // ---- phase A: dQ per 16-query tile
  for (int qt = warp; qt < HWp / 16; qt += 8) {
    uint32_t qa[2][4], da[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      ldsm_x4(qa[ks], fa.a(q_s, qt * 16, 16 * ks));
      ldsm_x4(da[ks], fa.a(d_s, qt * 16, 16 * ks));
    }
    const float ls[2] = {Ls[qt * 16 + g], Ls[qt * 16 + g + 8]}, dd[2] = {Dd[qt * 16 + g], Dd[qt * 16 + g + 8]};
    float dQ[4][4];
#pragma unroll
    for (int x = 0; x < 16; ++x) (&dQ[0][0])[x] = 0.f;
    for (int kb = 0; kb < NKp; kb += 64) {
      float S[8][4], dP[8][4];
#pragma unroll
      for (int x = 0; x < 32; ++x) (&S[0][0])[x] = (&dP[0][0])[x] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t kf[4], vf[4];
          ldsm_x4(kf, fa.b(k_s, kb + np * 16, 16 * ks));
          ldsm_x4(vf, fa.b(v_s, kb + np * 16, 16 * ks));
          mma16816<FMT>(S[2 * np], qa[ks], kf);
          mma16816<FMT>(S[2 * np + 1], qa[ks], kf + 2);
          mma16816<FMT>(dP[2 * np], da[ks], vf);
          mma16816<FMT>(dP[2 * np + 1], da[ks], vf + 2);
        }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int key = kb + nt * 8 + 2 * t + (c & 1);
          const float pv = key < NK ? __expf(S[nt][c] * scale - ls[c >> 1]) : 0.f;
          S[nt][c] = pv * (dP[nt][c] - dd[c >> 1]);       // dS
        }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t sa[4] = {pack2<FMT>(S[2 * j][0], S[2 * j][1]), pack2<FMT>(S[2 * j][2], S[2 * j][3]),
                                pack2<FMT>(S[2 * j + 1][0], S[2 * j + 1][1]), pack2<FMT>(S[2 * j + 1][2], S[2 * j + 1][3])};
#pragma unroll
        for (int dh = 0; dh < 2; ++dh) {
          uint32_t kf[4];
          ldsm_x4_trans(kf, fa.bt(k_s, kb + 16 * j, 16 * dh));
          mma16816<FMT>(dQ[2 * dh], sa, kf);
          mma16816<FMT>(dQ[2 * dh + 1], sa, kf + 2);
        }
      }
    }
#pragma unroll
    for (int rh = 0; rh < 2; ++rh) {
      const int n = qt * 16 + g + 8 * rh;
      if (n < HW) {
        uint16_t* orow = dqkv + (static_cast<long long>(bf) * HW + n) * 3 * HD + h * 32 + 2 * t;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
          *reinterpret_cast<uint32_t*>(orow + 8 * nt) = pack2<FMT>(dQ[nt][2 * rh] * scale, dQ[nt][2 * rh + 1] * scale);
      }
    }
  }

  // ---- phase B: dK, dV per 16-key tile (rows = keys, columns = queries)
  for (int kt = warp; kt < (NK + 15) / 16; kt += 8) {
    uint32_t ka[2][4], va[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      ldsm_x4(ka[ks], fa.a(k_s, kt * 16, 16 * ks));
      ldsm_x4(va[ks], fa.a(v_s, kt * 16, 16 * ks));
    }
    float dK[4][4], dV[4][4];
#pragma unroll
    for (int x = 0; x < 16; ++x) (&dK[0][0])[x] = (&dV[0][0])[x] = 0.f;
    for (int qb = 0; qb < HWp; qb += 16) {
      float ST[2][4], dPT[2][4];
#pragma unroll
      for (int x = 0; x < 8; ++x) (&ST[0][0])[x] = (&dPT[0][0])[x] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        uint32_t qf[4], df[4];
        ldsm_x4(qf, fa.b(q_s, qb, 16 * ks));
        ldsm_x4(df, fa.b(d_s, qb, 16 * ks));
        mma16816<FMT>(ST[0], ka[ks], qf);
        mma16816<FMT>(ST[1], ka[ks], qf + 2);
        mma16816<FMT>(dPT[0], va[ks], df);
        mma16816<FMT>(dPT[1], va[ks], df + 2);
      }
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int key = kt * 16 + g + 8 * (c >> 1), q = qb + nt * 8 + 2 * t + (c & 1);
          const float pv = (key < NK && q < HW) ? __expf(ST[nt][c] * scale - Ls[q]) : 0.f;
          ST[nt][c] = pv;                                  // P^T
          dPT[nt][c] = pv * (dPT[nt][c] - Dd[q]);          // dS^T
        }
      const uint32_t pa[4] = {pack2<FMT>(ST[0][0], ST[0][1]), pack2<FMT>(ST[0][2], ST[0][3]), pack2<FMT>(ST[1][0], ST[1][1]),
                              pack2<FMT>(ST[1][2], ST[1][3])};
      const uint32_t sa[4] = {pack2<FMT>(dPT[0][0], dPT[0][1]), pack2<FMT>(dPT[0][2], dPT[0][3]), pack2<FMT>(dPT[1][0], dPT[1][1]),
                              pack2<FMT>(dPT[1][2], dPT[1][3])};
#pragma unroll
      for (int dh = 0; dh < 2; ++dh) {
        uint32_t qf[4], df[4];
        ldsm_x4_trans(qf, fa.bt(q_s, qb, 16 * dh));
        ldsm_x4_trans(df, fa.bt(d_s, qb, 16 * dh));
        mma16816<FMT>(dK[2 * dh], sa, qf);
        mma16816<FMT>(dK[2 * dh + 1], sa, qf + 2);
        mma16816<FMT>(dV[2 * dh], pa, df);
        mma16816<FMT>(dV[2 * dh + 1], pa, df + 2);
      }
    }
#pragma unroll
    for (int rh = 0; rh < 2; ++rh) {
      const int key = kt * 16 + g + 8 * rh;
      if (key == 0) {
        float* ge = dekv + static_cast<long long>(bf) * 2 * HD + h * 32 + 2 * t;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          ge[8 * nt] = dK[nt][2 * rh] * scale;
          ge[8 * nt + 1] = dK[nt][2 * rh + 1] * scale;
          ge[HD + 8 * nt] = dV[nt][2 * rh];
          ge[HD + 8 * nt + 1] = dV[nt][2 * rh + 1];
        }
      } else if (key < NK) {
        uint16_t* krow = dqkv + (static_cast<long long>(bf) * HW + key - 1) * 3 * HD + HD + h * 32 + 2 * t;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          *reinterpret_cast<uint32_t*>(krow + 8 * nt) = pack2<FMT>(dK[nt][2 * rh] * scale, dK[nt][2 * rh + 1] * scale);
          *reinterpret_cast<uint32_t*>(krow + HD + 8 * nt) = pack2<FMT>(dV[nt][2 * rh], dV[nt][2 * rh + 1]);
        }
      }
    }
  }
}

}  // namespace vmm

using namespace vmm;

extern "C" int vmm_sattn_fwd(const void* qkv, const float* ekv, void* out, float* lse, int fmt, int BF, int frames, int HW,
                             int heads, float scale, void* stream_) {
  (void)frames;
  if (!qkv || !ekv || !out) return set_error(VMM_ERR_ARG, "vmm_sattn_fwd: null pointer");
  if (fmt != VMM_FMT_F16 && fmt != VMM_FMT_BF16) return set_error(VMM_ERR_ARG, "vmm_sattn_fwd: bad fmt");
  const int HWp = (HW + 15) & ~15, NKp = (HW + 1 + 63) & ~63;
  const size_t smem = static_cast<size_t>(HWp + 2 * NKp) * SP * sizeof(uint16_t);
  if (smem > 200 * 1024) return set_error(VMM_ERR_UNSUPPORTED, "vmm_sattn_fwd: too many keys for one CTA");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(sattn_fwd_mma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sattn_fwd_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return set_cuda_error(e, "vmm_sattn_fwd: attr");
    attr = true;
  }
  if (fmt == VMM_FMT_F16)
    sattn_fwd_mma_kernel<0><<<dim3(heads, BF), 256, smem, stream>>>(static_cast<const uint16_t*>(qkv), ekv, static_cast<uint16_t*>(out), lse, HW, heads, scale);
  else
    sattn_fwd_mma_kernel<1><<<dim3(heads, BF), 256, smem, stream>>>(static_cast<const uint16_t*>(qkv), ekv, static_cast<uint16_t*>(out), lse, HW, heads, scale);
  count_launch();
  return check_launch("vmm_sattn_fwd");
}

extern "C" int vmm_sattn_bwd(const void* qkv, const float* ekv, const void* aout, const void* dout, const float* lse, void* dqkv,
                             float* dekv, int fmt, int BF, int HW, int heads, float scale, void* stream_) {
  if (!qkv || !ekv || !aout || !dout || !lse || !dqkv || !dekv) return set_error(VMM_ERR_ARG, "vmm_sattn_bwd: null pointer");
  if (fmt != VMM_FMT_F16 && fmt != VMM_FMT_BF16) return set_error(VMM_ERR_ARG, "vmm_sattn_bwd: bad fmt");
  const int HWp = (HW + 15) & ~15, NKp = (HW + 1 + 63) & ~63;
  const size_t smem = static_cast<size_t>(2 * HWp + 2 * NKp) * SP * sizeof(uint16_t) + static_cast<size_t>(2) * HWp * sizeof(float);
  if (smem > 200 * 1024) return set_error(VMM_ERR_UNSUPPORTED, "vmm_sattn_bwd: too many keys for one CTA");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(sattn_bwd_mma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sattn_bwd_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return set_cuda_error(e, "vmm_sattn_bwd: attr");
    attr = true;
  }
  if (fmt == VMM_FMT_F16)
    sattn_bwd_mma_kernel<0><<<dim3(heads, BF), 256, smem, stream>>>(static_cast<const uint16_t*>(qkv), ekv, static_cast<const uint16_t*>(aout),
                                                                   static_cast<const uint16_t*>(dout), lse, static_cast<uint16_t*>(dqkv), dekv, HW, heads, scale);
  else
    sattn_bwd_mma_kernel<1><<<dim3(heads, BF), 256, smem, stream>>>(static_cast<const uint16_t*>(qkv), ekv, static_cast<const uint16_t*>(aout),
                                                                   static_cast<const uint16_t*>(dout), lse, static_cast<uint16_t*>(dqkv), dekv, HW, heads, scale);
  count_launch();
  return check_launch("vmm_sattn_bwd");
}
