// vmm_flattn_fwd: the whole Residual(PreNorm(SpatialLinearAttention)) block of a 64-channel level WITHOUT materialising qkv
// (inference form: nothing is kept for a backward pass).  VDDP:131-137, 245-264, 313-378.
//
//   kernel 1  flattn_ctx_kernel   x tile (128 pixels x 64 ch) --TMA--> LayerNorm in shared memory -->
//               K^T = Wk xn^T  (tcgen05, M = 128 head-dims of a 4-head group, N = 128 pixels): TMEM lane = head-dim d, so the
//               column-softmax statistics of k (max / sum over pixels, VDDP:368) are per-THREAD reductions over registers;
//               V = xn Wv^T    (tcgen05, M = 128 pixels, N = 128);
//               ctx_tile = w^T V with w = exp(k - m_tile) written as a K-major A tile and V as an MN-major B tile (tcgen05);
//               the 32 x 32 diagonal block of each head is folded into a running (ctx, max, sum) per head-dim with the usual
//               online-softmax rescaling; one partial per CTA.
//   kernel 2  flattn_finalize_kernel  combines the partials of a frame-image with the T conditioning tokens (VDDP:349-353):
//               ctx[d][e] = sum_m w[m][d] v[m][e] / (Z[d] n)  (VDDP:369-373)
//   kernel 3  flattn_out_kernel   x tile --TMA--> LayerNorm --> Q = xn Wq^T (tcgen05, N = 256) --> per-row softmax over each head's
//               32 values x scale (VDDP:367, 370) --> qs ctx (tcgen05 against the block-diagonal ctx of a 4-head group) -->
//               to_out (tcgen05, K = 256, N = 64) + bias + x --> bulk tensor store.
// HBM traffic per block: x three times (twice read, residual from L2) + out once, against ~41 tensor-sized passes unfused.
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace vmm {

constexpr int FL_XS = 0;            // 128 x 128 B   x tile -> normalised rows
constexpr int FL_W1 = 16384;        // kernel 1: Wk [256][64] ; kernel 3: Wq [256][64]                (32 KB)
constexpr int FL_W2 = 49152;        // kernel 1: Wv [256][64] ; kernel 3: Wout [4 k-blocks][64][64]    (32 KB)
constexpr int FL_T1 = 81920;        // kernel 1: w tiles [2 groups][2 k-blocks][128][64] ; kernel 3: qs / ao tile [4 k-blocks][128][64]   (64 KB)
constexpr int FL_T2 = 147456;       // kernel 1: v tiles [2 groups][2 n-chunks][128][64] ; kernel 3: ctx tiles [2 groups][2 k-blocks][128][64] (64 KB)
constexpr int FL_MISC = 212992;     // gamma 256 B | bias 256 B | control
constexpr int FL_SMEM = FL_MISC + 1024 + 1024;

struct FlCtl {
  uint64_t x_full, w_full, m1_full, m2_full, m3_full;
  uint32_t tmem_base;
  uint32_t pad;
};

struct FlDev {
  CUtensorMap xmap, omap, wmap, womap;
  const uint16_t* x;
  const float* gamma;
  const float* bias;       // to_out bias [64]
  float* part;             // [BF][chunks][256][34]: running ctx row (32), max, sum of every head-dim
  const float* ctx;        // [BF][8][32][32] final context (kernel 3)
  int BF, HW, tiles_per_bf, chunks, tiles_per_chunk;
  float eps, scale, hw;
  uint32_t idesc_kk128, idesc_kk256, idesc_kk64, idesc_kmn128;
};

template <int FMT>
__device__ __forceinline__ void fl_layernorm_tile(uint8_t* xs, const float* s_gamma, float eps) {
  // 128 rows x 64 channels, 128-byte swizzle, two threads per row (256 threads)
  const int tid = threadIdx.x;
  const int r = tid >> 1, hf = tid & 1;
  float v[32];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = 4 * hf + j;
    const uint4 u = *reinterpret_cast<const uint4*>(xs + r * 128 + ((c ^ (r & 7)) << 4));
    const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 f = unpack2_h16(w4[q], FMT);
      v[8 * j + 2 * q] = f.x;
      v[8 * j + 2 * q + 1] = f.y;
    }
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) s += v[j];
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  const float mean = s * (1.f / 64.f);
  float q2 = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float d = v[j] - mean;
    q2 += d * d;
  }
  q2 += __shfl_xor_sync(0xffffffffu, q2, 1);
  const float rstd = rsqrtf(q2 * (1.f / 64.f) + eps);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = 4 * hf + j;
    uint32_t w4[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int ch = 8 * c + 2 * q;
      w4[q] = pack2_h16((v[8 * j + 2 * q] - mean) * rstd * s_gamma[ch], (v[8 * j + 2 * q + 1] - mean) * rstd * s_gamma[ch + 1], FMT);
    }
    *reinterpret_cast<uint4*>(xs + r * 128 + ((c ^ (r & 7)) << 4)) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
  }
}

// ------------------------------------------------------------------------------------------------
// kernel 1: partial contexts
// ------------------------------------------------------------------------------------------------
template <int FMT>
__global__ void __launch_bounds__(256, 1) flattn_ctx_kernel(const __grid_constant__ FlDev p) {
  extern __shared__ uint8_t fl_smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(fl_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* xs = sm + FL_XS;
  uint8_t* wk = sm + FL_W1;
  uint8_t* wv = sm + FL_W2;
  uint8_t* wt = sm + FL_T1;
  uint8_t* vt = sm + FL_T2;
  float* s_gamma = reinterpret_cast<float*>(sm + FL_MISC);
  FlCtl* ctl = reinterpret_cast<FlCtl*>(sm + FL_MISC + 512);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bf = blockIdx.y, chunk = blockIdx.x;
  if (tid == 0) {
    tma_prefetch_desc(&p.xmap);
    tma_prefetch_desc(&p.wmap);
    mbar_init(&ctl->x_full, 1);
    mbar_init(&ctl->w_full, 1);
    mbar_init(&ctl->m1_full, 1);
    mbar_init(&ctl->m2_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&ctl->tmem_base, 512);
    tmem_relinquish();
  }
  for (int i = tid; i < 64; i += 256) s_gamma[i] = __ldg(p.gamma + i);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
  const bool lead = elect_one() && warp == 0;
  const int t0 = chunk * p.tiles_per_chunk, t1 = min(t0 + p.tiles_per_chunk, p.tiles_per_bf);
  if (lead && t0 < t1) {
    // k rows 256..511 and v rows 512..767 of the packed to_qkv weight, 128 rows per box
    mbar_expect_tx(&ctl->w_full, 65536);
    for (int g = 0; g < 2; ++g) {
      tma_load_2d(wk + g * 16384, &p.wmap, &ctl->w_full, 0, 256 + g * 128);
      tma_load_2d(wv + g * 16384, &p.wmap, &ctl->w_full, 0, 512 + g * 128);
    }
    mbar_expect_tx(&ctl->x_full, 16384);
    tma_load_4d(xs, &p.xmap, &ctl->x_full, 0, t0 * 128, bf, 0);
  }
  const int g = warp >> 2, q4 = warp & 3;
  const int row = q4 * 32 + lane;                        // TMEM lane: head-dim of the group (K^T, ctx) or pixel (V)
  const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16);
  float ctx_run[32], m_run = -1e30f, z_run = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) ctx_run[j] = 0.f;
  int it = 0;
  for (int t = t0; t < t1; ++t, ++it) {
    mbar_wait(&ctl->x_full, it & 1);
    fl_layernorm_tile<FMT>(xs, s_gamma, p.eps);
    fence_proxy_async_smem();
    __syncthreads();
    if (lead) {
      if (it == 0) mbar_wait(&ctl->w_full, 0);
      tc_fence_after();
      const uint64_t xdesc = make_smem_desc_sw128(smem_u32(xs), 16, 1024);
#pragma unroll
      for (int gg = 0; gg < 2; ++gg) {
        const uint64_t kdesc = make_smem_desc_sw128(smem_u32(wk + gg * 16384), 16, 1024);
        const uint64_t vdesc = make_smem_desc_sw128(smem_u32(wv + gg * 16384), 16, 1024);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          umma_f16(tmem_base + gg * 128, kdesc + static_cast<uint64_t>(kk * 2), xdesc + static_cast<uint64_t>(kk * 2), p.idesc_kk128, kk > 0 ? 1u : 0u);
          umma_f16(tmem_base + 256 + gg * 128, xdesc + static_cast<uint64_t>(kk * 2), vdesc + static_cast<uint64_t>(kk * 2), p.idesc_kk128, kk > 0 ? 1u : 0u);
        }
      }
      umma_commit(&ctl->m1_full);
    }
    __syncwarp();
    mbar_wait(&ctl->m1_full, it & 1);
    tc_fence_after();
    if (lead && t + 1 < t1) {            // the MMAs have read xs: next tile
      mbar_expect_tx(&ctl->x_full, 16384);
      tma_load_4d(xs, &p.xmap, &ctl->x_full, 0, (t + 1) * 128, bf, 0);
    }
    __syncwarp();
    // ---- K^T rows of this thread's head-dim: max over the tile's pixels, then w = exp(k - max).  The four 32-column TMEM loads of a
    // row are issued back to back and waited for once, and the row stays in registers for both passes (one CTA of 256 threads per SM:
    // registers are plentiful; the earlier form read the row twice, one load + wait at a time)
    float m_t = -1e30f, z_t = 0.f;
    {
      float kv[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32f(lane_taddr + static_cast<uint32_t>(g * 128 + c * 32), kv + 32 * c);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 128; ++j) m_t = fmaxf(m_t, kv[j]);
      uint8_t* wrow = wt + g * 32768 + row * 128;
#pragma unroll
      for (int cc = 0; cc < 16; ++cc) {     // 8 pixels per 16-byte chunk: k-block cc >> 3, chunk cc & 7
        float e[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          e[j] = __expf(kv[8 * cc + j] - m_t);
          z_t += e[j];
        }
        const uint4 u = make_uint4(pack2_h16(e[0], e[1], FMT), pack2_h16(e[2], e[3], FMT), pack2_h16(e[4], e[5], FMT), pack2_h16(e[6], e[7], FMT));
        *reinterpret_cast<uint4*>(wrow + (cc >> 3) * 16384 + (((cc & 7) ^ (row & 7)) << 4)) = u;
      }
    }
    // ---- V rows of this thread's pixel: 128 columns of the group as an MN-major B tile [pixel][64 e] x 2
    {
      float vv[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32f(lane_taddr + static_cast<uint32_t>(256 + g * 128 + c * 32), vv + 32 * c);
      tmem_ld_wait();
      uint8_t* vrow = vt + g * 32768 + row * 128;
#pragma unroll
      for (int cc = 0; cc < 16; ++cc) {
        const uint4 u = make_uint4(pack2_h16(vv[8 * cc], vv[8 * cc + 1], FMT), pack2_h16(vv[8 * cc + 2], vv[8 * cc + 3], FMT),
                                   pack2_h16(vv[8 * cc + 4], vv[8 * cc + 5], FMT), pack2_h16(vv[8 * cc + 6], vv[8 * cc + 7], FMT));
        *reinterpret_cast<uint4*>(vrow + (cc >> 3) * 16384 + (((cc & 7) ^ (row & 7)) << 4)) = u;
      }
    }
    tc_fence_before();
    fence_proxy_async_smem();
    __syncthreads();
    if (lead) {
      tc_fence_after();
#pragma unroll
      for (int gg = 0; gg < 2; ++gg) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {       // K = 128 pixels in steps of 16
          const uint64_t adesc = make_smem_desc_sw128(smem_u32(wt + gg * 32768 + (kk >> 2) * 16384) + (kk & 3) * 32, 16, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(smem_u32(vt + gg * 32768) + kk * 2048, 16384, 1024);
          umma_f16(tmem_base + gg * 128, adesc, bdesc, p.idesc_kmn128, kk > 0 ? 1u : 0u);
        }
      }
      umma_commit(&ctl->m2_full);
    }
    __syncwarp();
    mbar_wait(&ctl->m2_full, it & 1);
    tc_fence_after();
    {   // the diagonal block of this thread's head: columns q4 * 32 .. + 31 of the group's 128
      float c32[32];
      tmem_ld32f(lane_taddr + static_cast<uint32_t>(g * 128 + q4 * 32), c32);
      tmem_ld_wait();
      const float m_new = fmaxf(m_run, m_t);
      const float a = __expf(m_run - m_new), b = __expf(m_t - m_new);
#pragma unroll
      for (int j = 0; j < 32; ++j) ctx_run[j] = ctx_run[j] * a + c32[j] * b;
      z_run = z_run * a + z_t * b;
      m_run = m_new;
    }
    tc_fence_before();
    __syncthreads();
  }
  {
    float* dst = p.part + ((static_cast<long long>(bf) * p.chunks + chunk) * 256 + g * 128 + row) * 34;
#pragma unroll
    for (int j = 0; j < 32; ++j) dst[j] = ctx_run[j];
    dst[32] = m_run;
    dst[33] = z_run;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// kernel 2: combine the partials and the conditioning tokens; thread = (frame-image, head-dim)
// ------------------------------------------------------------------------------------------------
__global__ void flattn_finalize_kernel(const float* __restrict__ part, const float* __restrict__ ekv, int T, int chunks, int frames, float inv_hw,
                                       float* __restrict__ ctx, float* __restrict__ kstat, int BF) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= BF * 256) return;
  const int bf = i >> 8, d = i & 255, b = bf / frames;
  float m = -1e30f;
  for (int c = 0; c < chunks; ++c) m = fmaxf(m, part[((static_cast<long long>(bf) * chunks + c) * 256 + d) * 34 + 32]);
  for (int j = 0; j < T; ++j) m = fmaxf(m, ekv[(static_cast<long long>(b) * T + j) * 512 + d]);
  float acc[32], z = 0.f;
#pragma unroll
  for (int e = 0; e < 32; ++e) acc[e] = 0.f;
  for (int c = 0; c < chunks; ++c) {
    const float* pp = part + ((static_cast<long long>(bf) * chunks + c) * 256 + d) * 34;
    const float s = expf(pp[32] - m);
#pragma unroll
    for (int e = 0; e < 32; ++e) acc[e] += pp[e] * s;
    z += pp[33] * s;
  }
  const int h = d >> 5;
  for (int j = 0; j < T; ++j) {
    const float* tk = ekv + (static_cast<long long>(b) * T + j) * 512;
    const float w = expf(tk[d] - m);
#pragma unroll
    for (int e = 0; e < 32; ++e) acc[e] += w * tk[256 + h * 32 + e];
    z += w;
  }
  const float inv = inv_hw / z;
  float* dst = ctx + (static_cast<long long>(bf) * 256 + d) * 32;       // [bf][h][d % 32][e]
#pragma unroll
  for (int e = 0; e < 32; ++e) dst[e] = acc[e] * inv;
  if (kstat) {
    kstat[(static_cast<long long>(bf) * 256 + d) * 2] = m;
    kstat[(static_cast<long long>(bf) * 256 + d) * 2 + 1] = z;
  }
}

// ------------------------------------------------------------------------------------------------
// kernel 3: out = to_out(softmax_d(q) * scale . ctx) + bias + x
// ------------------------------------------------------------------------------------------------
template <int FMT>
__global__ void __launch_bounds__(256, 1) flattn_out_kernel(const __grid_constant__ FlDev p) {
  extern __shared__ uint8_t fl_smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(fl_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* xs = sm + FL_XS;
  uint8_t* wq = sm + FL_W1;
  uint8_t* wo = sm + FL_W2;
  uint8_t* qt = sm + FL_T1;
  uint8_t* ct = sm + FL_T2;
  float* s_gamma = reinterpret_cast<float*>(sm + FL_MISC);
  float* s_bias = s_gamma + 64;
  FlCtl* ctl = reinterpret_cast<FlCtl*>(sm + FL_MISC + 512);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bf = blockIdx.y, chunk = blockIdx.x;
  if (tid == 0) {
    tma_prefetch_desc(&p.xmap);
    tma_prefetch_desc(&p.omap);
    tma_prefetch_desc(&p.wmap);
    tma_prefetch_desc(&p.womap);
    mbar_init(&ctl->x_full, 1);
    mbar_init(&ctl->w_full, 1);
    mbar_init(&ctl->m1_full, 1);
    mbar_init(&ctl->m2_full, 1);
    mbar_init(&ctl->m3_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&ctl->tmem_base, 512);
    tmem_relinquish();
  }
  for (int i = tid; i < 64; i += 256) {
    s_gamma[i] = __ldg(p.gamma + i);
    s_bias[i] = p.bias ? __ldg(p.bias + i) : 0.f;
  }
  // block-diagonal context of each 4-head group as a K-major B tile [e rows][d], 16 bit, scaled by h*w (the context carries the
  // 1/(h w) of VDDP:371 and sits below fp16's normal range; the accumulator is scaled back)
  for (int i = tid; i < 65536 / 16; i += 256) reinterpret_cast<uint4*>(ct)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  for (int i = tid; i < 8 * 32 * 16; i += 256) {            // (head, e, d pair)
    const int dp = i & 15, e = (i >> 4) & 31, h = i >> 9;
    const float* ch = p.ctx + (static_cast<long long>(bf) * 8 + h) * 1024;
    const float c0 = ch[(2 * dp) * 32 + e] * p.hw, c1 = ch[(2 * dp + 1) * 32 + e] * p.hw;
    const int gq = h >> 2, hl = h & 3;
    const int rowi = hl * 32 + e, col = hl * 32 + 2 * dp;    // B[n = e][k = d]
    uint8_t* dst = ct + gq * 32768 + (col >> 6) * 16384 + rowi * 128 + ((((col & 63) >> 3) ^ (rowi & 7)) << 4) + (col & 7) * 2;
    *reinterpret_cast<uint32_t*>(dst) = pack2_h16(c0, c1, FMT);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
  const bool lead = elect_one() && warp == 0;
  const int t0 = chunk * p.tiles_per_chunk, t1 = min(t0 + p.tiles_per_chunk, p.tiles_per_bf);
  if (lead && t0 < t1) {
    mbar_expect_tx(&ctl->w_full, 65536);
    tma_load_2d(wq, &p.wmap, &ctl->w_full, 0, 0);
    tma_load_2d(wq + 16384, &p.wmap, &ctl->w_full, 0, 128);
    for (int kb = 0; kb < 4; ++kb) tma_load_2d(wo + kb * 8192, &p.womap, &ctl->w_full, kb * 64, 0);
    mbar_expect_tx(&ctl->x_full, 16384);
    tma_load_4d(xs, &p.xmap, &ctl->x_full, 0, t0 * 128, bf, 0);
  }
  const int half = warp >> 2, q4 = warp & 3;
  const int row = q4 * 32 + lane;                        // pixel of the tile == TMEM lane
  const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16);
  const float inv_hw = 1.f / p.hw;
  int it = 0;
  for (int t = t0; t < t1; ++t, ++it) {
    mbar_wait(&ctl->x_full, it & 1);
    fl_layernorm_tile<FMT>(xs, s_gamma, p.eps);
    if (lead) bulk_wait_read0();          // the previous tile's bulk store has read its staging rows (in qt)
    fence_proxy_async_smem();
    __syncthreads();
    if (lead) {
      if (it == 0) mbar_wait(&ctl->w_full, 0);
      tc_fence_after();
      const uint64_t xdesc = make_smem_desc_sw128(smem_u32(xs), 16, 1024);
      const uint64_t qdesc = make_smem_desc_sw128(smem_u32(wq), 16, 1024);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        umma_f16(tmem_base, xdesc + static_cast<uint64_t>(kk * 2), qdesc + static_cast<uint64_t>(kk * 2), p.idesc_kk256, kk > 0 ? 1u : 0u);
      umma_commit(&ctl->m1_full);
    }
    __syncwarp();
    // residual columns of this thread's row (32 of 64), requested early
    uint4 xres[4];
    {
      const long long grow = static_cast<long long>(bf) * p.HW + static_cast<long long>(t) * 128 + row;
      const bool ok = t * 128 + row < p.HW;
      const uint4* xr = reinterpret_cast<const uint4*>(p.x + (ok ? grow : 0) * 64 + half * 32);
#pragma unroll
      for (int j = 0; j < 4; ++j) xres[j] = ok ? __ldg(xr + j) : make_uint4(0u, 0u, 0u, 0u);
    }
    mbar_wait(&ctl->m1_full, it & 1);
    tc_fence_after();
    if (lead && t + 1 < t1) {
      mbar_expect_tx(&ctl->x_full, 16384);
      tma_load_4d(xs, &p.xmap, &ctl->x_full, 0, (t + 1) * 128, bf, 0);
    }
    __syncwarp();
    // ---- q rows: softmax over each head's 32 values, x scale -> A tile [pixel][256 d] (4 k-blocks); the four heads of this thread's
    // half are loaded from TMEM back to back and waited for once
    {
      float qv[128];
#pragma unroll
      for (int hh = 0; hh < 4; ++hh) tmem_ld32f(lane_taddr + static_cast<uint32_t>(half * 128 + hh * 32), qv + 32 * hh);
      tmem_ld_wait();
#pragma unroll
      for (int hh = 0; hh < 4; ++hh) {
        float* v = qv + 32 * hh;
        float mx = v[0];
#pragma unroll
        for (int j = 1; j < 32; ++j) mx = fmaxf(mx, v[j]);
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v[j] = __expf(v[j] - mx);
          sum += v[j];
        }
        const float inv = p.scale / sum;
        const int col0 = half * 128 + hh * 32;             // d column
        uint8_t* base = qt + (col0 >> 6) * 16384 + row * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int cc = ((col0 & 63) >> 3) + j;
          const uint4 u = make_uint4(pack2_h16(v[8 * j] * inv, v[8 * j + 1] * inv, FMT), pack2_h16(v[8 * j + 2] * inv, v[8 * j + 3] * inv, FMT),
                                     pack2_h16(v[8 * j + 4] * inv, v[8 * j + 5] * inv, FMT), pack2_h16(v[8 * j + 6] * inv, v[8 * j + 7] * inv, FMT));
          *reinterpret_cast<uint4*>(base + ((cc ^ (row & 7)) << 4)) = u;
        }
      }
    }
    tc_fence_before();
    fence_proxy_async_smem();
    __syncthreads();
    if (lead) {
      tc_fence_after();
#pragma unroll
      for (int gg = 0; gg < 2; ++gg) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t adesc = make_smem_desc_sw128(smem_u32(qt + (gg * 2 + (kk >> 2)) * 16384) + (kk & 3) * 32, 16, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(smem_u32(ct + gg * 32768 + (kk >> 2) * 16384) + (kk & 3) * 32, 16, 1024);
          umma_f16(tmem_base + 256 + gg * 128, adesc, bdesc, p.idesc_kk128, kk > 0 ? 1u : 0u);
        }
      }
      umma_commit(&ctl->m2_full);
    }
    __syncwarp();
    mbar_wait(&ctl->m2_full, it & 1);
    tc_fence_after();
    // ---- attention rows (x 1 / (h w)) over the same tile: A operand of to_out
    {
      float av[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32f(lane_taddr + static_cast<uint32_t>(256 + half * 128 + c * 32), av + 32 * c);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float* v = av + 32 * c;
        const int col0 = half * 128 + c * 32;
        uint8_t* base = qt + (col0 >> 6) * 16384 + row * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int cc = ((col0 & 63) >> 3) + j;
          const uint4 u = make_uint4(pack2_h16(v[8 * j] * inv_hw, v[8 * j + 1] * inv_hw, FMT), pack2_h16(v[8 * j + 2] * inv_hw, v[8 * j + 3] * inv_hw, FMT),
                                     pack2_h16(v[8 * j + 4] * inv_hw, v[8 * j + 5] * inv_hw, FMT), pack2_h16(v[8 * j + 6] * inv_hw, v[8 * j + 7] * inv_hw, FMT));
          *reinterpret_cast<uint4*>(base + ((cc ^ (row & 7)) << 4)) = u;
        }
      }
    }
    tc_fence_before();
    fence_proxy_async_smem();
    __syncthreads();
    if (lead) {
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) {
        const uint64_t adesc = make_smem_desc_sw128(smem_u32(qt + (kk >> 2) * 16384) + (kk & 3) * 32, 16, 1024);
        const uint64_t bdesc = make_smem_desc_sw128(smem_u32(wo + (kk >> 2) * 8192) + (kk & 3) * 32, 16, 1024);
        umma_f16(tmem_base, adesc, bdesc, p.idesc_kk64, kk > 0 ? 1u : 0u);
      }
      umma_commit(&ctl->m3_full);
    }
    __syncwarp();
    mbar_wait(&ctl->m3_full, it & 1);
    tc_fence_after();
    {
      float v[32];
      tmem_ld32f(lane_taddr + static_cast<uint32_t>(half * 32), v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t w4[4] = {xres[j].x, xres[j].y, xres[j].z, xres[j].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 f = unpack2_h16(w4[q], FMT);
          v[8 * j + 2 * q] += f.x + s_bias[half * 32 + 8 * j + 2 * q];
          v[8 * j + 2 * q + 1] += f.y + s_bias[half * 32 + 8 * j + 2 * q + 1];
        }
      }
      // staging for the bulk store: the first 16 KB of qt (the to_out MMAs have finished reading it)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 u = make_uint4(pack2_h16(v[8 * j], v[8 * j + 1], FMT), pack2_h16(v[8 * j + 2], v[8 * j + 3], FMT),
                                   pack2_h16(v[8 * j + 4], v[8 * j + 5], FMT), pack2_h16(v[8 * j + 6], v[8 * j + 7], FMT));
        *reinterpret_cast<uint4*>(qt + row * 128 + (((half * 4 + j) ^ (row & 7)) << 4)) = u;
      }
    }
    tc_fence_before();
    fence_proxy_async_smem();
    __syncthreads();
    if (lead) {
      tma_store_4d(&p.omap, qt, 0, t * 128, bf, 0);
      bulk_commit();
    }
    __syncwarp();
  }
  if (lead) bulk_wait0();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace vmm

using namespace vmm;

extern "C" size_t vmm_flattn_workspace(int BF) { return static_cast<size_t>(BF < 1 ? 1 : BF) * 8 * 256 * 34 * sizeof(float); }

extern "C" int vmm_flattn_fwd(const void* x, void* out, const void* wqkv, const void* wout, const float* gamma, const float* bias_out,
                              const float* ekv, int T, float* ctx, float* kstat, void* workspace, size_t workspace_bytes, int fmt, int BF,
                              int frames, int HW, int C, int heads, float scale, float eps, void* stream_) {
  if (!x || !out || !wqkv || !wout || !gamma || !ekv || !ctx || !workspace) return set_error(VMM_ERR_ARG, "vmm_flattn_fwd: null pointer");
  if (fmt != VMM_FMT_F16 && fmt != VMM_FMT_BF16) return set_error(VMM_ERR_ARG, "vmm_flattn_fwd: bad fmt");
  if (heads != 8 || C != 64) return set_error(VMM_ERR_UNSUPPORTED, "vmm_flattn_fwd: 8 heads of 32 on a 64-channel level only");
  if (HW < 128 || (HW % 128) != 0) return set_error(VMM_ERR_UNSUPPORTED, "vmm_flattn_fwd: the pixel count must be a multiple of 128");
  if (BF < 1 || frames < 1 || T < 0 || T > 64) return set_error(VMM_ERR_ARG, "vmm_flattn_fwd: BF / frames / T");
  if (workspace_bytes < vmm_flattn_workspace(BF)) return set_error(VMM_ERR_ARG, "vmm_flattn_fwd: workspace too small (vmm_flattn_workspace)");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FlDev d;
  memset(&d, 0, sizeof(d));
  const CUtensorMapDataType dt = fmt == VMM_FMT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  {
    const uint64_t gdim[4] = {64, (uint64_t)HW, (uint64_t)BF, 1};
    const uint64_t gstr[3] = {128, (uint64_t)HW * 128, (uint64_t)BF * HW * 128};
    const uint32_t box[4] = {64, 128, 1, 1};
    int rc = encode_tensor_map(&d.xmap, dt, 4, x, gdim, gstr, box, false);
    if (rc) return rc;
    rc = encode_tensor_map(&d.omap, dt, 4, out, gdim, gstr, box, false);
    if (rc) return rc;
  }
  {
    const uint64_t gdim[2] = {64, 768};
    const uint64_t gstr[1] = {128};
    const uint32_t box[2] = {64, 128};
    int rc = encode_tensor_map(&d.wmap, dt, 2, wqkv, gdim, gstr, box, true);
    if (rc) return rc;
  }
  {
    const uint64_t gdim[2] = {256, 64};
    const uint64_t gstr[1] = {512};
    const uint32_t box[2] = {64, 64};
    int rc = encode_tensor_map(&d.womap, dt, 2, wout, gdim, gstr, box, true);
    if (rc) return rc;
  }
  d.x = static_cast<const uint16_t*>(x);
  d.gamma = gamma;
  d.bias = bias_out;
  d.part = static_cast<float*>(workspace);
  d.ctx = ctx;
  d.BF = BF;
  d.HW = HW;
  d.tiles_per_bf = HW / 128;
  // CTAs per frame-image: every CTA walks a contiguous range of the frame-image's tiles.  Pick the split (1..8) that minimises
  // waves x tiles per CTA (352 CTAs of 18 tiles on 148 SMs run as three waves; 440 CTAs of 15 tiles also do)
  int chunks = 1;
  long long best = -1;
  for (int c = 1; c <= 8 && c <= d.tiles_per_bf; ++c) {
    const long long waves = (1LL * BF * c + num_sms() - 1) / num_sms();
    const long long cost = waves * ((d.tiles_per_bf + c - 1) / c) * 16 + c;        // + c: fewer partials on ties
    if (best < 0 || cost < best) {
      best = cost;
      chunks = c;
    }
  }
  d.tiles_per_chunk = (d.tiles_per_bf + chunks - 1) / chunks;
  chunks = (d.tiles_per_bf + d.tiles_per_chunk - 1) / d.tiles_per_chunk;
  d.chunks = chunks;
  d.eps = eps;
  d.scale = scale;
  d.hw = static_cast<float>(HW);
  d.idesc_kk128 = make_idesc_f16(128, 128, fmt, 0, 0);
  d.idesc_kk256 = make_idesc_f16(128, 256, fmt, 0, 0);
  d.idesc_kk64 = make_idesc_f16(128, 64, fmt, 0, 0);
  d.idesc_kmn128 = make_idesc_f16(128, 128, fmt, 0, 1);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(flattn_ctx_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, FL_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(flattn_ctx_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FL_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(flattn_out_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, FL_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(flattn_out_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FL_SMEM);
    if (e != cudaSuccess) return set_cuda_error(e, "vmm_flattn_fwd: attributes");
    attr = true;
  }
  const dim3 grid(chunks, BF);
  if (fmt == VMM_FMT_F16) flattn_ctx_kernel<0><<<grid, 256, FL_SMEM, stream>>>(d);
  else flattn_ctx_kernel<1><<<grid, 256, FL_SMEM, stream>>>(d);
  count_launch();
  flattn_finalize_kernel<<<(BF * 256 + 255) / 256, 256, 0, stream>>>(d.part, ekv, T, chunks, frames, 1.f / static_cast<float>(HW), ctx, kstat, BF);
  count_launch();
  if (fmt == VMM_FMT_F16) flattn_out_kernel<0><<<grid, 256, FL_SMEM, stream>>>(d);
  else flattn_out_kernel<1><<<grid, 256, FL_SMEM, stream>>>(d);
  count_launch();
  return check_launch("vmm_flattn_fwd");
}
