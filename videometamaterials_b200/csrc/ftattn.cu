// vmm_ftattn_fwd: the whole Residual(PreNorm(temporal Attention)) block of a 64-channel level in ONE kernel.
//   VDDP:131-137 (Residual), 245-264 (LayerNorm / PreNorm), 381-394 (EinopsToAndFrom), 396-535 (Attention), 615.
//
//   x tile (11 pixels x 11 frames = 121 rows x 64 ch) --TMA--> smem --LayerNorm in place--> A operand
//   per head pair hg (4 of them):
//       qkv_hg = xn W_hg^T          tcgen05.mma  M=128 N=192 K=64   (weights of the pair streamed by TMA, L2 resident)
//       TMEM -> registers: rotary on q (pre-scaled table) and k -> 16-bit rows in shared memory (never in HBM when sampling)
//       11 x 22 attention per (pixel, head) on mma.sync m16n8k16 with ldmatrix fragments (cond keys / values and the
//       position bias as ready-made register fragments), normalised output -> 16-bit A tile (128-byte swizzle)
//       out += ao_hg Wout_hg^T      tcgen05.mma  M=128 N=64 K=64, accumulated over the four pairs in TMEM
//   out + x (residual) -> 16-bit rows -> one bulk tensor store.
//
// The 768-wide qkv rows and the 256-wide attention output rows exist only on chip; a training forward additionally
// writes them (and the normalised rows) for the backward kernels, which is still less than half of the unfused traffic.
// One CTA per SM (the driver never co-schedules two CTAs of a kernel that uses tcgen05: measured, occupancy 1 at any register /
// shared-memory footprint) made of TWO independent groups of 8 warps.  Each group owns 111 KB of shared memory, 256 TMEM columns,
// its own mbarriers and named barrier, and walks its own tiles: while one group waits for a TMA load or an MMA, the other runs
// its register phase, so nothing inside a group is double buffered.
// Rows of a tile are frame-major (row = frame * 11 + pixel): that is the order a (c, pixel, frame, b) tensor map delivers.
#include "common.cuh"
#include "mma_sync.cuh"
#include "sm100_ptx.cuh"

namespace vmm {

constexpr int FNF = 11;                 // frames == cond tokens (VDDP:603)
constexpr int FPX = 11;                 // pixels per tile
constexpr int FROWS = FNF * FPX;        // 121 rows of the 128-row MMA tile carry data
constexpr int FSEG = 16384;             // one staged segment (q, k or v of the head pair): 121 rows x 128 B, 128-byte swizzle

// shared-memory carve-up (offsets from a 1024-byte aligned base)
constexpr int OFF_WS = 0;               // 192 x 128 B   to_qkv rows of the head pair  (B operand, 128-byte swizzle)
constexpr int OFF_WOS = 24576;          // 64 x 128 B    to_out columns of the head pair (B operand)
constexpr int OFF_XS = 32768;           // 121 x 128 B   x tile -> normalised rows (A operand) -> output rows
constexpr int OFF_MISC = 48256;         // zero row 384 | gamma 256 | control
constexpr int OFF_AOS = 49152;          // 121 x 128 B   attention output of the head pair (A operand)
constexpr int OFF_STG = 65536;          // 3 x (121 x 128 B, padded to 16 KB)   staged q | k | v tiles of the head pair (TMA-storable)
constexpr int FT_GROUP = 114688;                            // bytes per warp group: 65536 + 2 * 16384 + 15488 = 113792, rounded up to a multiple of
                                                            // 1024 so that the swizzled tiles of group 1 stay 1024-byte aligned
constexpr int FT_SMEM = 2 * FT_GROUP + 1024;                // + alignment slack = 230400 <= 227 KB; the rotary tables are read through L1

struct FtattnCtl {
  uint64_t x_full, w_full, wo_full, qkv_full, ao_done;
  uint32_t tmem_base;
  uint32_t pad;
};

struct FtattnDev {
  CUtensorMap xmap, omap, xnmap, wmap, womap, qsmap, asmap;
  const uint32_t* cfrag;   // [B][8 heads][32 lanes][24 words]: cond key / value fragments and bias of every (sample, head), see ftattn_prep_kernel
  const uint16_t* x;
  const float* gamma;
  const float* ekv;
  const float* bias;
  const float* rot;
  int save_qkv, save_ao, save_xn;
  int B, HW, tiles_per_b, total_tiles;
  float eps;
  uint32_t idesc_qkv, idesc_out;
};

// 16-byte chunk `cc` (0..23 = segment * 8 + chunk) of staged row `r`: three 128-byte-swizzled tiles (the layout TMA stores and
// ldmatrix rows on distinct banks both want)
__device__ __forceinline__ uint32_t stg_off(int r, int cc) { return static_cast<uint32_t>((cc >> 3) * FSEG + r * 128 + (((cc ^ r) & 7) << 4)); }

// Per (sample, head, lane) the mma.sync B fragments of the conditioning keys / values and the position bias (+ key mask) in
// the accumulator layout: 24 words, so that the main kernel fetches them with six 16-byte loads per head pair.
template <int FMT>
__global__ void ftattn_prep_kernel(const float* __restrict__ ekv, const float* __restrict__ bias, uint32_t* __restrict__ cfrag, int B) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * 8 * 32) return;
  const int lane = idx & 31, h = (idx >> 5) & 7, b = idx >> 8;
  const int g = lane >> 2, t4 = lane & 3;
  uint32_t w[24];
#pragma unroll
  for (int i = 0; i < 16; ++i) w[i] = 0u;
  if (ekv) {
    const float* eb = ekv + static_cast<long long>(b) * FNF * 512;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const int key = 8 * nt + g;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int d = 16 * ks + 8 * hf + 2 * t4;
          const float a0 = key < FNF ? eb[key * 512 + h * 32 + d] : 0.f;
          const float a1 = key < FNF ? eb[key * 512 + h * 32 + d + 1] : 0.f;
          w[nt * 4 + ks * 2 + hf] = pack2<FMT>(a0, a1);
        }
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int d = 8 * nt + g;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int k0 = 2 * t4 + 8 * hf;
        const float a0 = (k0 < FNF) ? eb[k0 * 512 + 256 + h * 32 + d] : 0.f;
        const float a1 = (k0 + 1 < FNF) ? eb[(k0 + 1) * 512 + 256 + h * 32 + d] : 0.f;
        w[8 + nt * 2 + hf] = pack2<FMT>(a0, a1);
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
        const int j = 8 * nt + 2 * t4 + c;
        w[16 + rh * 4 + nt * 2 + c] = __float_as_uint((j < FNF) ? ((i < FNF) ? bias[(h * FNF + i) * FNF + j] : 0.f) : -1e30f);
      }
  }
  uint4* dst = reinterpret_cast<uint4*>(cfrag + static_cast<long long>(idx) * 24);
#pragma unroll
  for (int i = 0; i < 6; ++i) dst[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
}

template <int FMT>
__global__ void __launch_bounds__(512, 1) ftattn_fwd_kernel(const __grid_constant__ FtattnDev p) {
  pdl_trigger();
  extern __shared__ uint8_t ft_smem_raw[];
  uint8_t* sm0 = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ft_smem_raw) + 1023) & ~uintptr_t(1023));
  const int grp = threadIdx.x >> 8;                       // warp group: an independent worker with its own tiles and buffers
  uint8_t* sm = sm0 + grp * FT_GROUP;
  auto grp_sync = [&]() { asm volatile("bar.sync %0, 256;" ::"r"(1 + grp) : "memory"); };
  uint8_t* ws = sm + OFF_WS;
  uint8_t* wos = sm + OFF_WOS;
  uint8_t* xs = sm + OFF_XS;
  uint8_t* zrow = sm + OFF_MISC;
  float* s_gamma = reinterpret_cast<float*>(sm + OFF_MISC + 384);
  FtattnCtl* ctl = reinterpret_cast<FtattnCtl*>(sm + OFF_MISC + 640);
  uint8_t* aos = sm + OFF_AOS;
  uint8_t* stg = sm + OFF_STG;
  const float* rt = p.rot;                      // [2][11][16][2] fp32, 2.8 KB: L1-resident after the first tile

  const int tid = threadIdx.x & 255, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const bool cond = p.ekv != nullptr;

  if (tid == 0) {
    tma_prefetch_desc(&p.xmap);
    tma_prefetch_desc(&p.omap);
    tma_prefetch_desc(&p.wmap);
    tma_prefetch_desc(&p.womap);
    if (p.save_xn) tma_prefetch_desc(&p.xnmap);
    if (p.save_qkv) tma_prefetch_desc(&p.qsmap);
    if (p.save_ao) tma_prefetch_desc(&p.asmap);
    mbar_init(&ctl->x_full, 1);
    mbar_init(&ctl->w_full, 1);
    mbar_init(&ctl->wo_full, 1);
    mbar_init(&ctl->qkv_full, 1);
    mbar_init(&ctl->ao_done, 1);
    fence_barrier_init();
  }
  if ((threadIdx.x >> 5) == 1) {                          // one warp allocates all 512 columns: 256 per group
    tmem_alloc(&reinterpret_cast<FtattnCtl*>(sm0 + OFF_MISC + 640)->tmem_base, 512);
    tmem_relinquish();
  }
  for (int i = tid; i < 96; i += 256) reinterpret_cast<uint32_t*>(zrow)[i] = 0u;
  for (int i = tid; i < 64; i += 256) s_gamma[i] = __ldg(p.gamma + i);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = reinterpret_cast<FtattnCtl*>(sm0 + OFF_MISC + 640)->tmem_base + static_cast<uint32_t>(grp * 256);
  const bool lead = elect_one() && warp == 0;      // the one thread that issues TMA and tcgen05.mma

  const int t_first = static_cast<int>(blockIdx.x) * 2 + grp, t_stride = static_cast<int>(gridDim.x) * 2;
  const int my_tiles = t_first < p.total_tiles ? (p.total_tiles - t_first + t_stride - 1) / t_stride : 0;
  const uint32_t total_k = static_cast<uint32_t>(my_tiles) * 4u;
  if (lead && my_tiles > 0) {
    mbar_expect_tx(&ctl->w_full, 192 * 128);
    for (int seg = 0; seg < 3; ++seg) tma_load_2d(ws + seg * 8192, &p.wmap, &ctl->w_full, 0, seg * 256);
    mbar_expect_tx(&ctl->wo_full, 64 * 128);
    tma_load_2d(wos, &p.womap, &ctl->wo_full, 0, 0);
  }

  // role constants
  const int q4 = warp & 3, chalf = warp >> 2;         // TMEM lane quarter / column half of the drain and of the final epilogue
  const int row = q4 * 32 + lane;                     // tile row == TMEM lane owned in the drain
  const bool rvalid = row < FROWS;
  const int rf = row / FPX, rp = row - rf * FPX;      // frame, pixel of that row
  const int rfc = rf < FNF ? rf : FNF - 1;            // rows 121..127 carry no data: keep their table index in range
  const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16);
  const int hl = warp & 1, psub = warp >> 1;          // core: head of the pair, pixel subset
  const int lm = lane >> 3, lr = lane & 7;
  const uint32_t stg_s = smem_u32(stg), zrow_s = smem_u32(zrow), aos_s = smem_u32(aos);

  uint32_t k = 0;
  int it = 0;
  for (int t = t_first; t < p.total_tiles; t += t_stride, ++it) {
    const int b = t / p.tiles_per_b;
    const int p0 = (t - b * p.tiles_per_b) * FPX;
    const int npx = min(FPX, p.HW - p0);
    if (lead) {
      mbar_expect_tx(&ctl->x_full, FROWS * 128);
      tma_load_4d(xs, &p.xmap, &ctl->x_full, 0, p0, 0, b);
    }
    __syncwarp();
    mbar_wait(&ctl->x_full, it & 1);

    // ---------------------------------------------------------------- channel LayerNorm in place (two threads per row)
    {
      const int r = tid >> 1, hf = tid & 1;
      float v[32];
      if (r < FROWS) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = 4 * hf + j;
          const uint4 u = *reinterpret_cast<const uint4*>(xs + r * 128 + ((c ^ (r & 7)) << 4));
          const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 f = unpack2<FMT>(w4[q]);
            v[8 * j + 2 * q] = f.x;
            v[8 * j + 2 * q + 1] = f.y;
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
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
      const float rstd = rsqrtf(q2 * (1.f / 64.f) + p.eps);
      if (r < FROWS) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = 4 * hf + j;
          uint32_t w4[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int ch = 8 * c + 2 * q;
            w4[q] = pack2<FMT>((v[8 * j + 2 * q] - mean) * rstd * s_gamma[ch], (v[8 * j + 2 * q + 1] - mean) * rstd * s_gamma[ch + 1]);
          }
          *reinterpret_cast<uint4*>(xs + r * 128 + ((c ^ (r & 7)) << 4)) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
        }
      }
    }
    fence_proxy_async_smem();
    grp_sync();
    if (lead && p.save_xn) {
      tma_store_4d(&p.xnmap, xs, 0, p0, 0, b);
      bulk_commit();
    }

    for (int hg = 0; hg < 4; ++hg, ++k) {
      // ---------------------------------------------------------------- q | k | v of the head pair on tcgen05
      // (pairs 1..3 were issued right after the previous drain, so their MMAs ran under the previous attention core)
      if (hg == 0) {
        if (lead) {
          mbar_wait(&ctl->w_full, k & 1);
          tc_fence_after();
          const uint64_t adesc = make_smem_desc_sw128(smem_u32(xs), 16, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(smem_u32(ws), 16, 1024);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_f16(tmem_base, adesc + static_cast<uint64_t>(kk * 2), bdesc + static_cast<uint64_t>(kk * 2), p.idesc_qkv, kk > 0 ? 1u : 0u);
          umma_commit(&ctl->qkv_full);
        }
        __syncwarp();
      }
      mbar_wait(&ctl->qkv_full, k & 1);
      tc_fence_after();
      if (lead) {
        if (k + 1 < total_k) {                 // the MMAs have read ws: stream the weights of the next pair
          const int hgn = (hg + 1) & 3;
          mbar_expect_tx(&ctl->w_full, 192 * 128);
          for (int seg = 0; seg < 3; ++seg) tma_load_2d(ws + seg * 8192, &p.wmap, &ctl->w_full, 0, seg * 256 + hgn * 64);
        }
        if (p.save_ao) bulk_wait_read0();      // the bulk store of the previous pair's attention rows has read aos (core below rewrites it
      }                                        // after the drain's barrier)
      __syncwarp();
      // ---------------------------------------------------------------- drain: TMEM -> rotary -> staged 16-bit rows
#pragma unroll 1
      for (int s = 0; s < 3; ++s) {
        const int c0 = chalf * 96 + s * 32;              // one (q | k | v, head) slice of 32 columns
        float v[32];
        tmem_ld32f(lane_taddr + static_cast<uint32_t>(c0), v);
        tmem_ld_wait();
        const int seg = c0 >> 6;
        if (seg < 2) {
          const float4* tb = reinterpret_cast<const float4*>(rt + seg * (FNF * 32) + rfc * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 cs = __ldg(tb + j);               // (cos, sin) of pairs 2j and 2j + 1 (q table carries the scale)
            const float a0 = v[4 * j], b0 = v[4 * j + 1], a1 = v[4 * j + 2], b1 = v[4 * j + 3];
            v[4 * j] = a0 * cs.x - b0 * cs.y;
            v[4 * j + 1] = b0 * cs.x + a0 * cs.y;
            v[4 * j + 2] = a1 * cs.z - b1 * cs.w;
            v[4 * j + 3] = b1 * cs.z + a1 * cs.w;
          }
        }
        if (rvalid) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 u = make_uint4(pack2<FMT>(v[8 * j], v[8 * j + 1]), pack2<FMT>(v[8 * j + 2], v[8 * j + 3]),
                                       pack2<FMT>(v[8 * j + 4], v[8 * j + 5]), pack2<FMT>(v[8 * j + 6], v[8 * j + 7]));
            *reinterpret_cast<uint4*>(stg + stg_off(row, (c0 >> 3) + j)) = u;
          }
        }
      }
      tc_fence_before();
      if (p.save_qkv) fence_proxy_async_smem();
      grp_sync();                         // staged rows complete; the qkv accumulator may be overwritten
      if (lead) {
        if (hg < 3) {                          // next pair's projection now: it runs on the tensor pipe under this pair's attention core
          mbar_wait(&ctl->w_full, (k + 1) & 1);
          tc_fence_after();
          const uint64_t adesc = make_smem_desc_sw128(smem_u32(xs), 16, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(smem_u32(ws), 16, 1024);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_f16(tmem_base, adesc + static_cast<uint64_t>(kk * 2), bdesc + static_cast<uint64_t>(kk * 2), p.idesc_qkv, kk > 0 ? 1u : 0u);
          umma_commit(&ctl->qkv_full);
        }
        if (p.save_qkv) {                      // training: q | k | v rows of the pair for the backward kernels (three bulk tensor stores)
          for (int seg = 0; seg < 3; ++seg) tma_store_4d(&p.qsmap, stg + seg * FSEG, seg * 256 + hg * 64, p0, 0, b);
          bulk_commit();
        }
      }
      __syncwarp();

      // ---------------------------------------------------------------- per-warp constants of head h: cond fragments, bias
      uint32_t kc[2][2][2], vc[4][2];
      float bs[2][2][2];
      {
        const uint4* cf = reinterpret_cast<const uint4*>(p.cfrag + ((static_cast<long long>(b) * 8 + hg * 2 + hl) * 32 + lane) * 24);
        const uint4 c0 = __ldg(cf), c1 = __ldg(cf + 1), c2 = __ldg(cf + 2), c3 = __ldg(cf + 3), c4 = __ldg(cf + 4), c5 = __ldg(cf + 5);
        kc[0][0][0] = c0.x; kc[0][0][1] = c0.y; kc[0][1][0] = c0.z; kc[0][1][1] = c0.w;
        kc[1][0][0] = c1.x; kc[1][0][1] = c1.y; kc[1][1][0] = c1.z; kc[1][1][1] = c1.w;
        vc[0][0] = c2.x; vc[0][1] = c2.y; vc[1][0] = c2.z; vc[1][1] = c2.w;
        vc[2][0] = c3.x; vc[2][1] = c3.y; vc[3][0] = c3.z; vc[3][1] = c3.w;
        bs[0][0][0] = __uint_as_float(c4.x); bs[0][0][1] = __uint_as_float(c4.y); bs[0][1][0] = __uint_as_float(c4.z); bs[0][1][1] = __uint_as_float(c4.w);
        bs[1][0][0] = __uint_as_float(c5.x); bs[1][0][1] = __uint_as_float(c5.y); bs[1][1][0] = __uint_as_float(c5.z); bs[1][1][1] = __uint_as_float(c5.w);
      }
      // the to_out MMAs of the previous pair have read aos (and wos): both may be rewritten
      if (k > 0) mbar_wait(&ctl->ao_done, (k - 1) & 1);
      if (lead) {
        if (hg > 0) {                           // (the columns of pair 0 are loaded at the end of the previous tile)
          mbar_expect_tx(&ctl->wo_full, 64 * 128);
          tma_load_2d(wos, &p.womap, &ctl->wo_full, hg * 64, 0);
        }
      }
      __syncwarp();

      // ---------------------------------------------------------------- attention core: warp = (head of the pair, pixel subset)
      for (int pix = psub; pix < npx; pix += 4) {
        auto qaddr = [&](int fr, int col) -> uint32_t {
          return (fr < FNF) ? stg_s + stg_off(fr * FPX + pix, col >> 3) : zrow_s + static_cast<uint32_t>(((col >> 3) & 7) << 4);
        };
        float S[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int c = 0; c < 4; ++c) S[nt][c] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          uint32_t qa[4];
          ldsm_x4(qa, qaddr(lr + 8 * (lm & 1), hl * 32 + 16 * ks + 8 * (lm >> 1)));
          if (cond) {
            mma16816<FMT>(S[0], qa, kc[0][ks]);
            mma16816<FMT>(S[1], qa, kc[1][ks]);
          }
          uint32_t kb[4];
          ldsm_x4(kb, qaddr(lr + 8 * (lm >> 1), 64 + hl * 32 + 16 * ks + 8 * (lm & 1)));
          mma16816<FMT>(S[2], qa, kb);
          mma16816<FMT>(S[3], qa, kb + 2);
        }
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
        float O[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int c = 0; c < 4; ++c) O[nt][c] = 0.f;
        if (cond) {
          const uint32_t pa[4] = {pack2<FMT>(S[0][0], S[0][1]), pack2<FMT>(S[0][2], S[0][3]), pack2<FMT>(S[1][0], S[1][1]),
                                  pack2<FMT>(S[1][2], S[1][3])};
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) mma16816<FMT>(O[nt], pa, vc[nt]);
        }
        {
          const uint32_t pa[4] = {pack2<FMT>(S[2][0], S[2][1]), pack2<FMT>(S[2][2], S[2][3]), pack2<FMT>(S[3][0], S[3][1]),
                                  pack2<FMT>(S[3][2], S[3][3])};
#pragma unroll
          for (int dh = 0; dh < 2; ++dh) {
            uint32_t vb[4];
            ldsm_x4_trans(vb, qaddr(lr + 8 * (lm & 1), 128 + hl * 32 + 16 * dh + 8 * (lm >> 1)));
            mma16816<FMT>(O[2 * dh], pa, vb);
            mma16816<FMT>(O[2 * dh + 1], pa, vb + 2);
          }
        }
        // normalised rows g (< 8) and g + 8 (< 11) into the A tile of the to_out MMA: row-major 128-byte rows, 128-byte swizzle
        const float inv0 = 1.f / sum[0], inv1 = 1.f / sum[1];
        {
          const int r0 = g * FPX + pix;
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
            *reinterpret_cast<uint32_t*>(aos + r0 * 128 + (((hl * 4 + nt) ^ (r0 & 7)) << 4) + 4 * t4) = pack2<FMT>(O[nt][0] * inv0, O[nt][1] * inv0);
          if (g + 8 < FNF) {
            const int r1 = (g + 8) * FPX + pix;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
              *reinterpret_cast<uint32_t*>(aos + r1 * 128 + (((hl * 4 + nt) ^ (r1 & 7)) << 4) + 4 * t4) = pack2<FMT>(O[nt][2] * inv1, O[nt][3] * inv1);
          }
        }
      }
      if (lead && p.save_qkv) bulk_wait_read0();   // the bulk stores of this pair's staged rows have read them (issued a whole core ago)
      fence_proxy_async_smem();
      grp_sync();                         // attention rows of the pair complete; the staged rows may be overwritten
      // ---------------------------------------------------------------- out += ao_pair Wout_pair^T
      if (lead) {
        mbar_wait(&ctl->wo_full, k & 1);
        tc_fence_after();
        const uint64_t adesc = make_smem_desc_sw128(aos_s, 16, 1024);
        const uint64_t bdesc = make_smem_desc_sw128(smem_u32(wos), 16, 1024);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_f16(tmem_base + 192, adesc + static_cast<uint64_t>(kk * 2), bdesc + static_cast<uint64_t>(kk * 2), p.idesc_out,
                   (hg > 0 || kk > 0) ? 1u : 0u);
        umma_commit(&ctl->ao_done);
        if (p.save_ao) {                        // training: the attention rows of the pair (A operand of the to_out weight gradient)
          tma_store_4d(&p.asmap, aos, hg * 64, p0, 0, b);
          bulk_commit();
        }
      }
      __syncwarp();
    }

    // ------------------------------------------------------------------ out + x -> 16-bit rows -> bulk tensor store
    uint4 xres[4];                              // residual columns of this thread's row: requested before the last MMAs are waited for
    {
      const bool rres = rvalid && rp < npx;
      const long long grow = (static_cast<long long>(b) * FNF + rf) * p.HW + p0 + (rres ? rp : 0);
      const uint4* xr = reinterpret_cast<const uint4*>(p.x + (rres ? grow : 0) * 64 + chalf * 32);
#pragma unroll
      for (int j = 0; j < 4; ++j) xres[j] = rres ? __ldg(xr + j) : make_uint4(0u, 0u, 0u, 0u);
    }
    mbar_wait(&ctl->ao_done, (k - 1) & 1);
    tc_fence_after();
    if (lead) bulk_wait_read0();               // every bulk store that reads xs / stg / aos (training) has finished reading
    grp_sync();
    if (lead && k < total_k) {                  // to_out columns of the first pair for the next tile
      mbar_expect_tx(&ctl->wo_full, 64 * 128);
      tma_load_2d(wos, &p.womap, &ctl->wo_full, 0, 0);
    }
    {
      float v[32];
      tmem_ld32f(lane_taddr + 192u + static_cast<uint32_t>(chalf * 32), v);
      tmem_ld_wait();
      if (rvalid) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t w4[4] = {xres[j].x, xres[j].y, xres[j].z, xres[j].w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 f = unpack2<FMT>(w4[q]);
            v[8 * j + 2 * q] += f.x;
            v[8 * j + 2 * q + 1] += f.y;
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 u = make_uint4(pack2<FMT>(v[8 * j], v[8 * j + 1]), pack2<FMT>(v[8 * j + 2], v[8 * j + 3]),
                                     pack2<FMT>(v[8 * j + 4], v[8 * j + 5]), pack2<FMT>(v[8 * j + 6], v[8 * j + 7]));
          *reinterpret_cast<uint4*>(xs + row * 128 + (((chalf * 4 + j) ^ (row & 7)) << 4)) = u;
        }
      }
    }
    tc_fence_before();
    fence_proxy_async_smem();
    grp_sync();
    if (lead) {
      tma_store_4d(&p.omap, xs, 0, p0, 0, b);
      bulk_commit();
      bulk_wait_read0();                       // xs is the destination of the next tile's load
    }
    __syncwarp();
  }

  if (lead) bulk_wait0();
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 1) {
    tc_fence_after();
    tmem_dealloc(reinterpret_cast<FtattnCtl*>(sm0 + OFF_MISC + 640)->tmem_base, 512);
  }
}

}  // namespace vmm

using namespace vmm;

extern "C" size_t vmm_ftattn_workspace(int B) { return static_cast<size_t>(B < 1 ? 1 : B) * 8 * 32 * 24 * sizeof(uint32_t); }

extern "C" int vmm_ftattn_fwd(const void* x, void* out, const void* wqkv, const void* wout, const float* gamma, const float* ekv,
                              const float* bias, const float* rot, void* xn_save, void* qkv_save, void* ao_save, void* workspace,
                              size_t workspace_bytes, int fmt, int B, int frames, int HW, int C, int heads, float eps, void* stream_) {
  if (!x || !out || !wqkv || !wout || !gamma || !bias || !rot || !workspace) return set_error(VMM_ERR_ARG, "vmm_ftattn_fwd: null pointer");
  if (fmt != VMM_FMT_F16 && fmt != VMM_FMT_BF16) return set_error(VMM_ERR_ARG, "vmm_ftattn_fwd: bad fmt");
  if (frames != FNF) return set_error(VMM_ERR_UNSUPPORTED, "vmm_ftattn_fwd: only 11 frames (the reference hard-codes 11 cond tokens, VDDP:603)");
  if (heads != 8 || C != 64) return set_error(VMM_ERR_UNSUPPORTED, "vmm_ftattn_fwd: 8 heads of 32 on a 64-channel level only (wider levels take the unfused kernels)");
  if (B < 1 || HW < 1) return set_error(VMM_ERR_ARG, "vmm_ftattn_fwd: B / HW");
  if (workspace_bytes < vmm_ftattn_workspace(B) || (reinterpret_cast<uintptr_t>(workspace) & 15) != 0)
    return set_error(VMM_ERR_ARG, "vmm_ftattn_fwd: workspace too small (vmm_ftattn_workspace) or not 16-byte aligned");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FtattnDev d;
  memset(&d, 0, sizeof(d));
  const CUtensorMapDataType dt = fmt == VMM_FMT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  {
    // activations (b, f, pixel, c): dims (c, pixel, frame, b); a box of 11 pixels x 11 frames lands frame-major in shared memory
    const uint64_t gdim[4] = {64, (uint64_t)HW, (uint64_t)FNF, (uint64_t)B};
    const uint64_t gstr[3] = {128, (uint64_t)HW * 128, (uint64_t)FNF * HW * 128};
    const uint32_t box[4] = {64, FPX, FNF, 1};
    int rc = encode_tensor_map(&d.xmap, dt, 4, x, gdim, gstr, box, false);
    if (rc) return rc;
    rc = encode_tensor_map(&d.omap, dt, 4, out, gdim, gstr, box, false);
    if (rc) return rc;
    if (xn_save) {
      rc = encode_tensor_map(&d.xnmap, dt, 4, xn_save, gdim, gstr, box, false);
      if (rc) return rc;
    }
    if (qkv_save) {       // rows of 768: a box is the 64 columns of one (q | k | v, head pair) segment
      const uint64_t qdim[4] = {768, (uint64_t)HW, (uint64_t)FNF, (uint64_t)B};
      const uint64_t qstr[3] = {1536, (uint64_t)HW * 1536, (uint64_t)FNF * HW * 1536};
      rc = encode_tensor_map(&d.qsmap, dt, 4, qkv_save, qdim, qstr, box, false);
      if (rc) return rc;
    }
    if (ao_save) {
      const uint64_t adim[4] = {256, (uint64_t)HW, (uint64_t)FNF, (uint64_t)B};
      const uint64_t astr[3] = {512, (uint64_t)HW * 512, (uint64_t)FNF * HW * 512};
      rc = encode_tensor_map(&d.asmap, dt, 4, ao_save, adim, astr, box, false);
      if (rc) return rc;
    }
  }
  {
    const uint64_t gdim[2] = {64, 768};
    const uint64_t gstr[1] = {128};
    const uint32_t box[2] = {64, 64};
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
  d.cfrag = static_cast<const uint32_t*>(workspace);
  d.x = static_cast<const uint16_t*>(x);
  d.gamma = gamma;
  d.ekv = ekv;
  d.bias = bias;
  d.rot = rot;
  d.save_qkv = qkv_save ? 1 : 0;
  d.save_ao = ao_save ? 1 : 0;
  d.save_xn = xn_save ? 1 : 0;
  d.B = B;
  d.HW = HW;
  d.tiles_per_b = (HW + FPX - 1) / FPX;
  const long long total = 1LL * B * d.tiles_per_b;
  if (total > 0x7fffffffLL) return set_error(VMM_ERR_ARG, "vmm_ftattn_fwd: tile count");
  d.total_tiles = static_cast<int>(total);
  d.eps = eps;
  d.idesc_qkv = make_idesc_f16(128, 192, fmt, 0, 0);
  d.idesc_out = make_idesc_f16(128, 64, fmt, 0, 0);

  static int ctas_per_sm = 0;
  if (ctas_per_sm == 0) {
    cudaError_t e = cudaFuncSetAttribute(ftattn_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ftattn_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ftattn_fwd_kernel<0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ftattn_fwd_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return set_cuda_error(e, "vmm_ftattn_fwd: attributes");
    ctas_per_sm = 1;
  }
  {
    const int n = B * 8 * 32;
    if (fmt == VMM_FMT_F16) ftattn_prep_kernel<0><<<(n + 255) / 256, 256, 0, stream>>>(ekv, bias, static_cast<uint32_t*>(workspace), B);
    else ftattn_prep_kernel<1><<<(n + 255) / 256, 256, 0, stream>>>(ekv, bias, static_cast<uint32_t*>(workspace), B);
    count_launch();
  }
  long long grid = 1LL * ctas_per_sm * num_sms();          // persistent: every CTA = two warp groups = two tiles in flight
  if (grid > (total + 1) / 2) grid = (total + 1) / 2;
  if (fmt == VMM_FMT_F16) ftattn_fwd_kernel<0><<<static_cast<unsigned>(grid), 512, FT_SMEM, stream>>>(d);
  else ftattn_fwd_kernel<1><<<static_cast<unsigned>(grid), 512, FT_SMEM, stream>>>(d);
  count_launch();
  return check_launch("vmm_ftattn_fwd");
}

extern "C" int vmm_ftattn_ctas_per_sm(void) {
  int occ = 0;
  cudaFuncSetAttribute(ftattn_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM);
  cudaFuncSetAttribute(ftattn_fwd_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ftattn_fwd_kernel<1>, 512, FT_SMEM) != cudaSuccess) return -1;
  return occ;
}

// diagnostics: out[0..] = numRegs, static smem, max dynamic smem, smem per SM, reserved smem per block, regs per SM,
// occupancy at FT_SMEM / 114944 / 113664 / 106496 / 65536 bytes of dynamic shared memory
extern "C" int vmm_ftattn_diag(int* out) {
  cudaFuncSetAttribute(ftattn_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM);
  cudaFuncSetAttribute(ftattn_fwd_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncAttributes a;
  if (cudaFuncGetAttributes(&a, ftattn_fwd_kernel<1>) != cudaSuccess) return -1;
  int dev = 0, v = 0;
  cudaGetDevice(&dev);
  out[0] = a.numRegs;
  out[1] = static_cast<int>(a.sharedSizeBytes);
  out[2] = a.maxDynamicSharedSizeBytes;
  cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
  out[3] = v;
  cudaDeviceGetAttribute(&v, cudaDevAttrReservedSharedMemoryPerBlock, dev);
  out[4] = v;
  cudaDeviceGetAttribute(&v, cudaDevAttrMaxRegistersPerMultiprocessor, dev);
  out[5] = v;
  const int sizes[5] = {FT_SMEM, 114944, 113664, 106496, 65536};
  for (int i = 0; i < 5; ++i) {
    int occ = -1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ftattn_fwd_kernel<1>, 512, sizes[i]);
    out[6 + i] = occ;
  }
  out[11] = a.localSizeBytes ? static_cast<int>(a.localSizeBytes) : 0;
  return 0;
}
