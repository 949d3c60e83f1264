// vmm_qkv_bwd: the two consumers of an attention block's d(qkv) rows at a 64-channel level in ONE pass over them:
//     dxn[pix, c]  = sum_k dqkv[pix, k] * Wqkv[k, c]          (data gradient of to_qkv, VDDP:437 / 336)
//     dW[k, c]    += sum_pix dqkv[pix, k] * xn[pix, c]         (weight gradient of to_qkv)
// dqkv is 768 x 16 bit per position (1.25 GB at level 0, b = 8).  As two launches (vmm_cgemm + vmm_wgrad) it is streamed from HBM
// twice: 218 + 193 us per block, both at their HBM time (ncu launch list of round 2).  Here a [128 positions][128 columns] tile
// of it is loaded once (TMA, 128-byte swizzle) and read by the tensor core under TWO shared-memory descriptors:
//   * K-major A operand   (rows = positions, K = the 128 columns)      x  Wqkv^T chunk (K-major B, resident)   -> dxn tile  [128 x 64]
//   * MN-major A operand  (M = the 128 columns, K = the 128 positions) x  xn tile (MN-major B, N = 64)         -> dW rows   [128 x 64]
// TMEM: six dW accumulators of 64 columns (768 rows of dW, kept over ALL position tiles of the CTA, reduced into the fp32
// gradient once at the end) + two dxn accumulators (drained by the epilogue warps while the next tile is multiplied).
// Warp roles as in wgrad.cu: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocation, warps 4-7 epilogue.
// Shared memory: Wqkv^T [12 chunks][64 c][64 k] 96 KB resident | 3 stages of [2 chunks][128 positions][64 columns] 96 KB |
// 2 x xn tile 32 KB.
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace vmm {

constexpr int QB_COLS = 768;                 // columns of dqkv: 3 x 8 heads x 32
constexpr int QB_STAGES_PER_TILE = QB_COLS / 128;
constexpr int QB_RING = 3;
constexpr int QB_W_BYTES = 12 * 8192;        // 98304
constexpr int QB_STAGE_BYTES = 2 * 16384;
constexpr int QB_OFF_RING = QB_W_BYTES;
constexpr int QB_OFF_XN = QB_OFF_RING + QB_RING * QB_STAGE_BYTES;
constexpr int QB_OFF_CTL = QB_OFF_XN + 2 * 16384;
constexpr int QB_SMEM = QB_OFF_CTL + 256 + 1024;

struct QbDev {
  CUtensorMap dqmap, xmap, wmap;
  uint16_t* dxn;          // plain form: the data gradient rows; LayerNorm form: dx (gradient of the block input)
  float* dw;
  long long rows;
  int tiles;
  uint32_t idesc_w, idesc_d;
  // LayerNorm form (the PreNorm + Residual in front of to_qkv, VDDP:245-264, 131-137): dx = LN'(x; gamma)(dxn) + dres, dgamma += sum dxn * xhat
  const uint16_t* x;
  const uint16_t* dres;
  const float* gamma;
  float* dgamma;
  float eps;
};

struct __align__(8) QbCtl {
  uint64_t full[QB_RING], empty[QB_RING];
  uint64_t xfull[2], xempty[2];
  uint64_t dfull[2], dempty[2];
  uint64_t wfull, accw_full;
  uint32_t tmem_base;
  uint32_t pad;
};

template <int FMT, bool LN>
__global__ void __launch_bounds__(256, 1) qkv_bwd_kernel(const __grid_constant__ QbDev p) {
  extern __shared__ uint8_t qb_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(qb_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wsm = smem;
  uint8_t* ring = smem + QB_OFF_RING;
  uint8_t* xsm = smem + QB_OFF_XN;
  QbCtl* ctl = reinterpret_cast<QbCtl*>(smem + QB_OFF_CTL);
  __shared__ float s_gamma[64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (LN && threadIdx.x < 64) s_gamma[threadIdx.x] = __ldg(p.gamma + threadIdx.x);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.dqmap);
    tma_prefetch_desc(&p.xmap);
    tma_prefetch_desc(&p.wmap);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < QB_RING; ++s) {
      mbar_init(&ctl->full[s], 1);
      mbar_init(&ctl->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&ctl->xfull[a], 1);
      mbar_init(&ctl->xempty[a], 1);
      mbar_init(&ctl->dfull[a], 1);
      mbar_init(&ctl->dempty[a], 128);
    }
    mbar_init(&ctl->wfull, 1);
    mbar_init(&ctl->accw_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(&ctl->tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
  const bool leader = elect_one();
  // contiguous range of position tiles of this CTA
  const int t0 = static_cast<int>(static_cast<long long>(blockIdx.x) * p.tiles / gridDim.x);
  const int t1 = static_cast<int>(static_cast<long long>(blockIdx.x + 1) * p.tiles / gridDim.x);

  if (warp == 0) {
    if (leader && t0 < t1) {
      mbar_expect_tx(&ctl->wfull, QB_W_BYTES);
      for (int c = 0; c < 12; ++c) tma_load_2d(wsm + c * 8192, &p.wmap, &ctl->wfull, c * 64, 0);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int t = t0; t < t1; ++t, ++it) {
        const int xb = it & 1;
        mbar_wait(&ctl->xempty[xb], ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(&ctl->xfull[xb], 16384);
        tma_load_2d(xsm + xb * 16384, &p.xmap, &ctl->xfull[xb], 0, t * 128);
        for (int j = 0; j < QB_STAGES_PER_TILE; ++j) {
          mbar_wait(&ctl->empty[s], ph ^ 1);
          uint8_t* st = ring + s * QB_STAGE_BYTES;
          mbar_expect_tx(&ctl->full[s], QB_STAGE_BYTES);
          tma_load_2d(st, &p.dqmap, &ctl->full[s], j * 128, t * 128);
          tma_load_2d(st + 16384, &p.dqmap, &ctl->full[s], j * 128 + 64, t * 128);
          if (++s == QB_RING) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (t0 < t1) {
      mbar_wait(&ctl->wfull, 0);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      const uint32_t w_addr = smem_u32(wsm);
      for (int t = t0; t < t1; ++t, ++it) {
        const int xb = it & 1, db = it & 1;
        mbar_wait(&ctl->xfull[xb], (it >> 1) & 1);
        mbar_wait(&ctl->dempty[db], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t x_addr = smem_u32(xsm + xb * 16384);
        const uint32_t d_x = tmem_base + 384 + db * 64;
        for (int j = 0; j < QB_STAGES_PER_TILE; ++j) {
          mbar_wait(&ctl->full[s], ph);
          tc_fence_after();
          if (leader) {
            const uint32_t a_addr = smem_u32(ring + s * QB_STAGE_BYTES);
            // dW rows 128 j .. 128 j + 127: A = dqkv tile as an MN-major operand (M = columns), B = xn tile MN-major (N = channels), K = positions
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const uint64_t adesc = make_smem_desc_sw128(a_addr + k * 2048, 16384, 1024);
              const uint64_t bdesc = make_smem_desc_sw128(x_addr + k * 2048, 16384, 1024);
              umma_f16(tmem_base + j * 64, adesc, bdesc, p.idesc_w, (it > 0 || k > 0) ? 1u : 0u);
            }
            // dxn tile: A = the same bytes as a K-major operand (rows = positions, K = columns), B = Wqkv^T chunks 2 j, 2 j + 1
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
              const uint64_t adesc = make_smem_desc_sw128(a_addr + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024);
              const uint64_t bdesc = make_smem_desc_sw128(w_addr + (2 * j + (kk >> 2)) * 8192 + (kk & 3) * 32, 16, 1024);
              umma_f16(d_x, adesc, bdesc, p.idesc_d, (j > 0 || kk > 0) ? 1u : 0u);
            }
            umma_commit(&ctl->empty[s]);
          }
          __syncwarp();
          if (++s == QB_RING) {
            s = 0;
            ph ^= 1;
          }
        }
        if (leader) {
          umma_commit(&ctl->xempty[xb]);
          umma_commit(&ctl->dfull[db]);
        }
        __syncwarp();
      }
      if (leader) umma_commit(&ctl->accw_full);
      __syncwarp();
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    float dg[LN ? 64 : 1];
    if (LN) {
#pragma unroll
      for (int c = 0; c < 64; ++c) dg[c] = 0.f;
    }
    int it = 0;
    for (int t = t0; t < t1; ++t, ++it) {
      const int db = it & 1;
      const long long r = static_cast<long long>(t) * 128 + row;
      const bool live = r < p.rows;
      uint4 xr[LN ? 8 : 1];
      if (LN) {                       // the row of the block input, requested before the accumulator is waited for
        const uint4* xp = reinterpret_cast<const uint4*>(p.x + (live ? r : 0) * 64);
#pragma unroll
        for (int j = 0; j < 8; ++j) xr[j] = live ? __ldg(xp + j) : make_uint4(0u, 0u, 0u, 0u);
      }
      mbar_wait(&ctl->dfull[db], (it >> 1) & 1);
      tc_fence_after();
      float v[64];
      tmem_ld32f(lane_taddr + 384 + db * 64, v);
      tmem_ld32f(lane_taddr + 384 + db * 64 + 32, v + 32);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&ctl->dempty[db]);          // the accumulator is in registers: the next tile but one may overwrite it
      if (LN) {
        // channel LayerNorm backward of this row (one thread = one position = 64 channels), as vmm_ln_bwd computes it
        const uint32_t* xw = reinterpret_cast<const uint32_t*>(xr);
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float2 f = unpack2_h16(xw[j], FMT);
          sum += f.x + f.y;
        }
        const float mean = sum * (1.f / 64.f);
        float qq = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float2 f = unpack2_h16(xw[j], FMT);
          qq += (f.x - mean) * (f.x - mean) + (f.y - mean) * (f.y - mean);
        }
        const float rstd = rsqrtf(qq * (1.f / 64.f) + p.eps);
        float a1 = 0.f, a2 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float2 f = unpack2_h16(xw[j], FMT);
          const float xh0 = (f.x - mean) * rstd, xh1 = (f.y - mean) * rstd;
          dg[2 * j] += v[2 * j] * xh0;
          dg[2 * j + 1] += v[2 * j + 1] * xh1;
          const float g0 = v[2 * j] * s_gamma[2 * j], g1 = v[2 * j + 1] * s_gamma[2 * j + 1];
          v[2 * j] = g0;
          v[2 * j + 1] = g1;
          a1 += g0 + g1;
          a2 += g0 * xh0 + g1 * xh1;
        }
        a1 *= (1.f / 64.f);
        a2 *= (1.f / 64.f);
        if (live) {
          const uint4* rp = reinterpret_cast<const uint4*>(p.dres + r * 64);
          uint4* dst = reinterpret_cast<uint4*>(p.dxn + r * 64);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 rr = __ldg(rp + j);
            const uint32_t rw[4] = {rr.x, rr.y, rr.z, rr.w};
            uint32_t o[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int c = 8 * j + 2 * u;
              const float2 f = unpack2_h16(xw[4 * j + u], FMT);
              const float2 d2 = unpack2_h16(rw[u], FMT);
              const float o0 = rstd * (v[c] - a1 - (f.x - mean) * rstd * a2) + d2.x;
              const float o1 = rstd * (v[c + 1] - a1 - (f.y - mean) * rstd * a2) + d2.y;
              o[u] = pack2_h16(o0, o1, FMT);
            }
            dst[j] = make_uint4(o[0], o[1], o[2], o[3]);
          }
        }
      } else if (live) {
        uint4* dst = reinterpret_cast<uint4*>(p.dxn + r * 64);     // one 128-byte line per thread
#pragma unroll
        for (int j = 0; j < 8; ++j)
          dst[j] = make_uint4(pack2_h16(v[8 * j], v[8 * j + 1], FMT), pack2_h16(v[8 * j + 2], v[8 * j + 3], FMT),
                              pack2_h16(v[8 * j + 4], v[8 * j + 5], FMT), pack2_h16(v[8 * j + 6], v[8 * j + 7], FMT));
      }
    }
    if (t0 < t1) {
      mbar_wait(&ctl->accw_full, 0);
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < QB_STAGES_PER_TILE; ++j) {
        float* dst = p.dw + (static_cast<long long>(j) * 128 + row) * 64;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          float v[32];
          tmem_ld32f(lane_taddr + j * 64 + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c * 32 + i), "f"(v[i]), "f"(v[i + 1]), "f"(v[i + 2]), "f"(v[i + 3])
                         : "memory");
        }
      }
      if (LN) {
        // every MMA has completed (accw_full): the stage ring is free.  [128 rows][64 + 1] partials -> one global add per channel
        float* red = reinterpret_cast<float*>(ring);
#pragma unroll
        for (int c = 0; c < 64; ++c) red[row * 65 + c] = dg[c];
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (row < 64) {
          float acc = 0.f;
          for (int rr = 0; rr < 128; ++rr) acc += red[rr * 65 + row];
          atomicAdd(p.dgamma + row, acc);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace vmm

using namespace vmm;

static int qkv_bwd_launch(const void* dqkv, const void* xn, const void* wd, void* out, float* dw, long long rows, int fmt, const void* x,
                          const void* dres, const float* gamma, float* dgamma, float eps, bool ln, cudaStream_t stream, const char* who) {
  if (!dqkv || !xn || !wd || !out || !dw || rows < 1) return set_error(VMM_ERR_ARG, "vmm_qkv_bwd: bad arguments");
  if (ln && (!x || !dres || !gamma || !dgamma)) return set_error(VMM_ERR_ARG, "vmm_qkv_ln_bwd: null pointer");
  if (fmt != VMM_FMT_F16 && fmt != VMM_FMT_BF16) return set_error(VMM_ERR_ARG, "vmm_qkv_bwd: bad fmt");
  if (rows >= (1LL << 31)) return set_error(VMM_ERR_UNSUPPORTED, "vmm_qkv_bwd: more than 2^31 rows");
  if ((reinterpret_cast<uintptr_t>(dw) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) || (reinterpret_cast<uintptr_t>(x) & 15) ||
      (reinterpret_cast<uintptr_t>(dres) & 15))
    return set_error(VMM_ERR_ARG, "vmm_qkv_bwd: pointers must be 16-byte aligned");
  QbDev d;
  memset(&d, 0, sizeof(d));
  const CUtensorMapDataType dt = fmt == VMM_FMT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  {
    const uint64_t gdim[2] = {QB_COLS, static_cast<uint64_t>(rows)};
    const uint64_t gstr[1] = {QB_COLS * 2};
    const uint32_t box[2] = {64, 128};
    int rc = encode_tensor_map(&d.dqmap, dt, 2, dqkv, gdim, gstr, box, false);
    if (rc) return rc;
  }
  {
    const uint64_t gdim[2] = {64, static_cast<uint64_t>(rows)};
    const uint64_t gstr[1] = {128};
    const uint32_t box[2] = {64, 128};
    int rc = encode_tensor_map(&d.xmap, dt, 2, xn, gdim, gstr, box, false);
    if (rc) return rc;
  }
  {
    const uint64_t gdim[2] = {QB_COLS, 64};
    const uint64_t gstr[1] = {QB_COLS * 2};
    const uint32_t box[2] = {64, 64};
    int rc = encode_tensor_map(&d.wmap, dt, 2, wd, gdim, gstr, box, true);
    if (rc) return rc;
  }
  d.dxn = static_cast<uint16_t*>(out);
  d.dw = dw;
  d.rows = rows;
  d.tiles = static_cast<int>((rows + 127) / 128);
  d.idesc_w = make_idesc_f16(128, 64, fmt, 1, 1);
  d.idesc_d = make_idesc_f16(128, 64, fmt, 0, 0);
  d.x = static_cast<const uint16_t*>(x);
  d.dres = static_cast<const uint16_t*>(dres);
  d.gamma = gamma;
  d.dgamma = dgamma;
  d.eps = eps;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(qkv_bwd_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, QB_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(qkv_bwd_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, QB_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(qkv_bwd_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, QB_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(qkv_bwd_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, QB_SMEM);
    if (e != cudaSuccess) return set_cuda_error(e, "vmm_qkv_bwd: cudaFuncSetAttribute");
    attr_set = true;
  }
  const int sms = num_sms();
  if (sms <= 0) return set_error(VMM_ERR_CUDA, "vmm_qkv_bwd: no CUDA device");
  const int grid = d.tiles < sms ? d.tiles : sms;
  if (fmt == VMM_FMT_F16) {
    if (ln) qkv_bwd_kernel<0, true><<<grid, 256, QB_SMEM, stream>>>(d);
    else qkv_bwd_kernel<0, false><<<grid, 256, QB_SMEM, stream>>>(d);
  } else {
    if (ln) qkv_bwd_kernel<1, true><<<grid, 256, QB_SMEM, stream>>>(d);
    else qkv_bwd_kernel<1, false><<<grid, 256, QB_SMEM, stream>>>(d);
  }
  count_launch();
  return check_launch(who);
}

extern "C" int vmm_qkv_bwd(const void* dqkv, const void* xn, const void* wd, void* dxn, float* dw, long long rows, int fmt, void* stream_) {
  return qkv_bwd_launch(dqkv, xn, wd, dxn, dw, rows, fmt, nullptr, nullptr, nullptr, nullptr, 0.f, false, static_cast<cudaStream_t>(stream_),
                        "vmm_qkv_bwd");
}

extern "C" int vmm_qkv_ln_bwd(const void* dqkv, const void* xn, const void* wd, const void* x, const void* dres, const float* gamma, float eps,
                              void* dx, float* dw, float* dgamma, long long rows, int fmt, void* stream_) {
  return qkv_bwd_launch(dqkv, xn, wd, dx, dw, rows, fmt, x, dres, gamma, dgamma, eps, true, static_cast<cudaStream_t>(stream_), "vmm_qkv_ln_bwd");
}
