// vmm_wgrad: weight gradients of every convolution / linear layer as a tcgen05 GEMM whose reduction
// dimension is the pixel axis:
//     dW[n, tap, c] += sum_pix dY[pix, n] * X[pix + d_tap, c]
// Both operands are "MN-major" for the tensor core (channels are contiguous in memory, pixels are K), which
// UMMA supports natively for 16-bit types: TMA drops [128 pixels][64 channels] boxes (128-byte swizzle) and
// the shared-memory descriptors are built with the MN-major flag.  One work item = (tap, 128-wide slice of n,
// <=128-wide slice of c, K range of pixel tiles); partial sums are added atomically into the fp32 gradient in
// the parameter's own (master) layout, so no unpack pass is needed.
//
// Slab mode (3x3 stride-1 taps on a 1 x 16 x 8 pixel tile, the level-0 / level-1 convolutions): the three ky taps of one
// (kx, 64-channel chunk) read pixel tiles that overlap in 14 of their 16 rows, so ONE box of 18 pixel rows is loaded per column
// block and the three 64-channel chunks of the B operand are addressed 1024 bytes (one pixel row = one swizzle atom) apart
// through the descriptor's leading-dimension stride.  B traffic per block drops from 3 to 1.125 tiles: these launches are
// bound by the L2 -> shared-memory fill, not by the MMAs.
#include <stdlib.h>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace vmm {

constexpr int kWgMaxCols = 160;
constexpr int kChunkBytes = 128 * 128;   // [128 pixels][64 ch] 16-bit

struct WgTapDev {
  int16_t a_src, b_src, dy, dx;
  int32_t c;          // valid channels of the B operand for this tap
  long long wofs;     // element offset of this tap inside dW
};

constexpr int kWgMaxChunks = 3;
struct WgColDev {     // one column block of the accumulator = up to 3 chunks of 64 B-operand channels, possibly of different taps
  int16_t n_chunks;
  int16_t a_src;      // all chunks of a block share the A view
  int16_t slab;       // 1: the chunks are the taps dy0, dy0 + 1, dy0 + 2 of one slab load (see above)
  int16_t pad;
  int16_t tap[kWgMaxChunks];
  int16_t c0[kWgMaxChunks];
};

struct WgradDev {
  CUtensorMap amap[VMM_MAX_VIEWS];
  CUtensorMap bmap[VMM_MAX_VIEWS];
  CUtensorMap bslab[VMM_MAX_VIEWS];   // box of th + 2 pixel rows (slab mode)
  uint32_t slab_bytes;
  WgTapDev taps[VMM_MAX_TAPS];
  WgColDev cols[kWgMaxCols];
  int n_cols, m_tiles, ksplit, total_items;
  int N;                 // rows of dW covered (channels of dY)
  int BNc;               // widest column block: 64 * chunks
  int a_chunks;          // 64-channel chunks of A loaded per stage (1 or 2)
  int tf_log, th_log, tw_log;
  int tiles_f, tiles_y, tiles_x, pix_tiles;
  int stages;
  uint32_t stage_bytes, idesc_by_chunks[kWgMaxChunks + 1];
  float* dw;
  long long sM, sC, sC2;
  int cmod, c_valid, k_valid;
};

struct __align__(8) WgSmemCtl {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t tfull[2];
  uint64_t tempty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

__device__ __forceinline__ void wg_decode(const WgradDev& p, int item, int& col, int& mt, int& kt0, int& kt1) {
  // column block fastest: the CTAs that run concurrently work on the same pixel range (same dY tile, neighbouring taps of the
  // same X pixels), so all but the first read of a pixel tile hit L2 instead of HBM
  col = item % p.n_cols;
  const int r = item / p.n_cols;
  mt = r % p.m_tiles;
  const int ks = r / p.m_tiles;
  kt0 = static_cast<int>(static_cast<long long>(ks) * p.pix_tiles / p.ksplit);
  kt1 = static_cast<int>(static_cast<long long>(ks + 1) * p.pix_tiles / p.ksplit);
}

__device__ __forceinline__ void wg_pix(const WgradDev& p, int kt, int& bf0, int& y0, int& x0) {
  const int xt = kt % p.tiles_x;
  int r = kt / p.tiles_x;
  const int yt = r % p.tiles_y;
  const int ft = r / p.tiles_y;
  bf0 = ft << p.tf_log;
  y0 = yt << p.th_log;
  x0 = xt << p.tw_log;
}

__global__ void __launch_bounds__(256, 1) wgrad_kernel(const __grid_constant__ WgradDev p) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  WgSmemCtl* ctl = reinterpret_cast<WgSmemCtl*>(smem + static_cast<size_t>(p.stages) * p.stage_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < VMM_MAX_VIEWS; ++i) {
      tma_prefetch_desc(&p.amap[i]);
      tma_prefetch_desc(&p.bmap[i]);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&ctl->full[s], 1);
      mbar_init(&ctl->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&ctl->tfull[a], 1);
      mbar_init(&ctl->tempty[a], 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(&ctl->tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();        // see cgemm.cu
  const uint32_t tmem_base = ctl->tmem_base;
  const bool leader = elect_one();   // see cgemm.cu: keeps the issue loops in uniform registers

  if (warp == 0) {
    if (leader) {
      int s = 0;
      uint32_t ph = 0;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        int col, mt, kt0, kt1;
        wg_decode(p, item, col, mt, kt0, kt1);
        const WgColDev cd = p.cols[col];
        for (int kt = kt0; kt < kt1; ++kt) {
          int bf0, y0, x0;
          wg_pix(p, kt, bf0, y0, x0);
          mbar_wait(&ctl->empty[s], ph ^ 1);
          uint8_t* st = smem + static_cast<size_t>(s) * p.stage_bytes;
          mbar_expect_tx(&ctl->full[s], cd.slab ? p.a_chunks * kChunkBytes + p.slab_bytes : (p.a_chunks + cd.n_chunks) * kChunkBytes);
          for (int a = 0; a < p.a_chunks; ++a)
            tma_load_4d(st + a * kChunkBytes, &p.amap[cd.a_src], &ctl->full[s], mt * 128 + a * 64, x0, y0, bf0);
          if (cd.slab) {
            const WgTapDev T = p.taps[cd.tap[0]];       // first (lowest dy) tap of the slab
            tma_load_4d(st + p.a_chunks * kChunkBytes, &p.bslab[T.b_src], &ctl->full[s], cd.c0[0], x0 + T.dx, y0 + T.dy, bf0);
          } else {
            for (int b = 0; b < cd.n_chunks; ++b) {
              const WgTapDev T = p.taps[cd.tap[b]];
              tma_load_4d(st + (p.a_chunks + b) * kChunkBytes, &p.bmap[T.b_src], &ctl->full[s], cd.c0[b], x0 + T.dx, y0 + T.dy, bf0);
            }
          }
          if (++s == p.stages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    int s = 0;
    uint32_t ph = 0;
    int it = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++it) {
      int col, mt, kt0, kt1;
      wg_decode(p, item, col, mt, kt0, kt1);
      const int acc = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1;
      mbar_wait(&ctl->tempty[acc], acc_ph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 256;
      const uint32_t idesc = p.idesc_by_chunks[p.cols[col].n_chunks];
      const uint32_t b_lbo = p.cols[col].slab ? 1024u : static_cast<uint32_t>(kChunkBytes);
      uint32_t accumulate = 0;
      for (int kt = kt0; kt < kt1; ++kt) {
        mbar_wait(&ctl->full[s], ph);
        tc_fence_after();
        if (leader) {
          const uint32_t a_addr = smem_u32(smem + static_cast<size_t>(s) * p.stage_bytes);
          const uint32_t b_addr = a_addr + p.a_chunks * kChunkBytes;
#pragma unroll
          for (int k = 0; k < 8; ++k) {   // 8 x 16 pixels
            const uint64_t adesc = make_smem_desc_sw128(a_addr + k * 2048, kChunkBytes, 1024);
            const uint64_t bdesc = make_smem_desc_sw128(b_addr + k * 2048, b_lbo, 1024);
            umma_f16(d_tmem, adesc, bdesc, idesc, accumulate);
            accumulate = 1;
          }
          umma_commit(&ctl->empty[s]);
        }
        __syncwarp();
        if (++s == p.stages) {
          s = 0;
          ph ^= 1;
        }
      }
      if (leader) umma_commit(&ctl->tfull[acc]);
      __syncwarp();
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int it = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++it) {
      int col, mt, kt0, kt1;
      wg_decode(p, item, col, mt, kt0, kt1);
      const WgColDev cd = p.cols[col];
      const int acc = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1;
      mbar_wait(&ctl->tfull[acc], acc_ph);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 256;
      const int m = mt * 128 + row;
      const bool mvalid = (m < p.N) && (kt1 > kt0);
      const bool plain = p.cmod >= (1 << 30);   // no (column -> (lo, hi)) remapping: column cj lives at dst + cj * sC
      for (int ch = 0; ch < cd.n_chunks; ++ch) {
        const WgTapDev T = p.taps[cd.tap[ch]];
        float* dst = p.dw + T.wofs + static_cast<long long>(m) * p.sM;
        for (int c = 0; c < 4; ++c) {
          uint32_t r[16];
          tmem_ld16(t_addr + ch * 64 + c * 16, r);
          tmem_ld_wait();
          const int cj0 = cd.c0[ch] + c * 16;
          if (!mvalid || cj0 >= T.c) continue;
          if (plain) {
            float* d0 = dst + static_cast<long long>(cj0) * p.sC;
            if (cj0 + 16 <= T.c) {
              if (p.sC == 1 && ((reinterpret_cast<uintptr_t>(d0) & 15) == 0)) {
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d0 + j), "f"(__uint_as_float(r[j])),
                               "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j + 2])), "f"(__uint_as_float(r[j + 3]))
                               : "memory");
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) atomicAdd(d0 + j * p.sC, __uint_as_float(r[j]));
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (cj0 + j < T.c) atomicAdd(d0 + j * p.sC, __uint_as_float(r[j]));
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int cj = cj0 + j;
              if (cj < T.c) {
                const int lo = cj % p.cmod, hi = cj / p.cmod;
                if (lo < p.c_valid && hi < p.k_valid) atomicAdd(dst + lo * p.sC + hi * p.sC2, __uint_as_float(r[j]));
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&ctl->tempty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// column sums of a [rows][N] 16-bit matrix into fp32 (bias gradients): out[n] += sum_rows x[row][n]
__global__ void __launch_bounds__(256) colsum_kernel(const uint16_t* __restrict__ x, long long rows, int N, long long ld, int fmt,
                                                     float* __restrict__ out, int rows_per_cta) {
  extern __shared__ float red[];   // [trows][N]
  const int vpr = N / 8;
  const int trows = blockDim.x / vpr;
  const int vc = threadIdx.x % vpr, vr = threadIdx.x / vpr;
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_cta;
  const long long r1 = min(r0 + rows_per_cta, rows);
  if (vr < trows) {
    for (long long r = r0 + vr; r < r1; r += trows) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(x + r * ld + vc * 8));
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack2_h16(w[j], fmt);
        s[2 * j] += f.x;
        s[2 * j + 1] += f.y;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[vr * N + vc * 8 + j] = s[j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    float a = 0.f;
    for (int r = 0; r < trows; ++r) a += red[r * N + i];
    atomicAdd(out + i, a);
  }
}

static int wg_ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return ((1 << l) == v) ? l : -1;
}

}  // namespace vmm

using namespace vmm;

extern "C" int vmm_wgrad(const vmm_wgrad_params* hp, void* stream_) {
  if (!hp) return set_error(VMM_ERR_ARG, "vmm_wgrad: null params");
  const vmm_wgrad_params& h = *hp;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (h.fmt != VMM_FMT_F16 && h.fmt != VMM_FMT_BF16) return set_error(VMM_ERR_ARG, "vmm_wgrad: bad fmt");
  if (h.n_a_views < 1 || h.n_a_views > VMM_MAX_VIEWS || h.n_b_views < 1 || h.n_b_views > VMM_MAX_VIEWS)
    return set_error(VMM_ERR_ARG, "vmm_wgrad: view count");
  if (h.n_taps < 1 || h.n_taps > VMM_MAX_TAPS) return set_error(VMM_ERR_ARG, "vmm_wgrad: n_taps");
  const int tfl = wg_ilog2(h.tf), thl = wg_ilog2(h.th), twl = wg_ilog2(h.tw);
  if (tfl < 0 || thl < 0 || twl < 0 || h.tf * h.th * h.tw != 128) return set_error(VMM_ERR_ARG, "vmm_wgrad: tile");
  if (!h.dw || h.n < 1) return set_error(VMM_ERR_ARG, "vmm_wgrad: dw / n");

  WgradDev d;
  memset(&d, 0, sizeof(d));
  d.N = h.n;
  d.m_tiles = (h.n + 127) / 128;
  d.a_chunks = h.n > 64 ? 2 : 1;
  // blocks of up to 3 chunks (2 when the A operand already needs two 64-channel chunks): more accumulator columns per
  // loaded A tile = fewer re-reads of dY
  const int max_chunks = d.a_chunks == 1 ? 3 : 2;
  int ncols = 0;
  d.BNc = 64;
  for (int t = 0; t < h.n_taps; ++t) {
    const vmm_wgrad_tap& T = h.taps[t];
    if (T.a_src < 0 || T.a_src >= h.n_a_views || T.b_src < 0 || T.b_src >= h.n_b_views || T.c < 1)
      return set_error(VMM_ERR_ARG, "vmm_wgrad: bad tap");
    d.taps[t].a_src = static_cast<int16_t>(T.a_src);
    d.taps[t].b_src = static_cast<int16_t>(T.b_src);
    d.taps[t].dy = static_cast<int16_t>(T.dy);
    d.taps[t].dx = static_cast<int16_t>(T.dx);
    d.taps[t].c = T.c;
    d.taps[t].wofs = T.wofs;
  }
  // slab mode: every tap must belong to a triple (same views, dx, c; dy = d0, d0 + 1, d0 + 2) on the 1 x 16 x 8 tile
  static const bool no_slab = getenv("VMM_WGRAD_NO_SLAB") != nullptr;
  bool slab = !no_slab && h.tf == 1 && h.th == 16 && h.tw == 8 && (h.n_taps % 3) == 0;
  int trip[VMM_MAX_TAPS][3];
  int n_trip = 0;
  if (slab) {
    bool used[VMM_MAX_TAPS] = {false};
    for (int t = 0; t < h.n_taps && slab; ++t) {
      if (used[t]) continue;
      // t must be the lowest dy of its triple: find the partners
      int t1 = -1, t2 = -1, lower = -1;
      for (int u = 0; u < h.n_taps; ++u) {
        if (u == t || used[u]) continue;
        const vmm_wgrad_tap &A = h.taps[t], &Bt = h.taps[u];
        if (A.a_src != Bt.a_src || A.b_src != Bt.b_src || A.dx != Bt.dx || A.c != Bt.c) continue;
        if (Bt.dy == A.dy + 1) t1 = u;
        if (Bt.dy == A.dy + 2) t2 = u;
        if (Bt.dy == A.dy - 1) lower = u;
      }
      if (lower >= 0) continue;            // not the first of its triple: it is picked up from there
      if (t1 < 0 || t2 < 0) {
        slab = false;
        break;
      }
      used[t] = used[t1] = used[t2] = true;
      trip[n_trip][0] = t;
      trip[n_trip][1] = t1;
      trip[n_trip][2] = t2;
      ++n_trip;
    }
    for (int t = 0; t < h.n_taps && slab; ++t)
      if (!used[t]) slab = false;
  }
  if (slab) {
    for (int g = 0; g < n_trip; ++g) {
      const vmm_wgrad_tap& T = h.taps[trip[g][0]];
      for (int c0 = 0; c0 < T.c; c0 += 64) {
        if (ncols >= kWgMaxCols) return set_error(VMM_ERR_UNSUPPORTED, "vmm_wgrad: too many column blocks");
        WgColDev* blk = &d.cols[ncols++];
        blk->n_chunks = 3;
        blk->slab = 1;
        blk->a_src = static_cast<int16_t>(T.a_src);
        for (int j = 0; j < 3; ++j) {
          blk->tap[j] = static_cast<int16_t>(trip[g][j]);
          blk->c0[j] = static_cast<int16_t>(c0);
        }
      }
    }
    d.BNc = 192;
  } else {
    for (int t = 0; t < h.n_taps; ++t) {
      const vmm_wgrad_tap& T = h.taps[t];
      for (int c0 = 0; c0 < T.c; c0 += 64) {
        WgColDev* blk = ncols > 0 ? &d.cols[ncols - 1] : nullptr;
        if (!blk || blk->n_chunks >= max_chunks || blk->a_src != T.a_src) {
          if (ncols >= kWgMaxCols) return set_error(VMM_ERR_UNSUPPORTED, "vmm_wgrad: too many column blocks");
          blk = &d.cols[ncols++];
          blk->n_chunks = 0;
          blk->slab = 0;
          blk->a_src = static_cast<int16_t>(T.a_src);
        }
        blk->tap[blk->n_chunks] = static_cast<int16_t>(t);
        blk->c0[blk->n_chunks] = static_cast<int16_t>(c0);
        ++blk->n_chunks;
        if (64 * blk->n_chunks > d.BNc) d.BNc = 64 * blk->n_chunks;
      }
    }
  }
  d.n_cols = ncols;
  d.tf_log = tfl;
  d.th_log = thl;
  d.tw_log = twl;
  d.tiles_f = (h.bf + h.tf - 1) / h.tf;
  d.tiles_y = (h.oh + h.th - 1) / h.th;
  d.tiles_x = (h.ow + h.tw - 1) / h.tw;
  d.pix_tiles = d.tiles_f * d.tiles_y * d.tiles_x;
  const int base_items = ncols * d.m_tiles;
  // split-K over pixel tiles.  Candidates: 1..4 items per SM; cost model (fitted to the b=8 step, profiles/README.md): the CTAs run
  // `waves` items one after the other, an item costs its pixel tiles plus ~6 tile-times for draining 128 x N fp32 reductions.
  // Few long items win for the deep levels (large dW, few pixels), three per SM for level 0 (3 column blocks x 148 splits).
  static const int forced = getenv("VMM_WGRAD_ITEMS_PER_SM") ? atoi(getenv("VMM_WGRAD_ITEMS_PER_SM")) : 0;
  int ksplit = 1;
  double best_cost = 1e30;
  for (int j = 1; j <= 4; ++j) {
    if (forced && j != forced) continue;
    int ks = (j * num_sms() + base_items - 1) / base_items;
    if (ks > d.pix_tiles) ks = d.pix_tiles;
    if (ks < 1) ks = 1;
    // keep at least 4 pixel tiles per item so the accumulate / drain overhead stays small
    if (d.pix_tiles / ks < 4 && d.pix_tiles >= 4) ks = d.pix_tiles / 4;
    const long long items = 1LL * base_items * ks;
    const long long waves = (items + num_sms() - 1) / num_sms();
    const double cost = static_cast<double>(waves) * ((d.pix_tiles + ks - 1) / ks + 6.0);
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      ksplit = ks;
    }
  }
  d.ksplit = ksplit;
  d.total_items = base_items * ksplit;
  d.slab_bytes = static_cast<uint32_t>((h.th + 2) * h.tw * 128);
  // slab mode: A chunks + one slab, padded to whole 1024-byte atoms; the MMA of the last k-step reads chunk 2 up to 2 atoms past
  // the 16-row tile, i.e. exactly to the end of the 18-row slab
  d.stage_bytes = slab ? (d.a_chunks * kChunkBytes + ((d.slab_bytes + 1023u) & ~1023u)) : (d.a_chunks + (d.BNc >> 6)) * kChunkBytes;
  d.stages = (200 * 1024) / static_cast<int>(d.stage_bytes);
  if (d.stages > 8) d.stages = 8;
  for (int nc = 1; nc <= kWgMaxChunks; ++nc) d.idesc_by_chunks[nc] = make_idesc_f16(128, 64 * nc, h.fmt, 1, 1);
  d.dw = h.dw;
  d.sM = h.s_m;
  d.sC = h.s_c;
  d.sC2 = h.s_c2;
  d.cmod = h.cmod > 0 ? h.cmod : (1 << 30);
  d.c_valid = h.cmod > 0 ? h.c_valid : (1 << 30);
  d.k_valid = h.cmod > 0 ? h.k_valid : 1;

  const CUtensorMapDataType dt = h.fmt == VMM_FMT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  uint32_t box[4] = {64, (uint32_t)h.tw, (uint32_t)h.th, (uint32_t)h.tf};
  for (int i = 0; i < VMM_MAX_VIEWS; ++i) {
    const vmm_view4& va = h.a[i < h.n_a_views ? i : 0];
    const vmm_view4& vb = h.b[i < h.n_b_views ? i : 0];
    if (!va.ptr || !vb.ptr) return set_error(VMM_ERR_ARG, "vmm_wgrad: null view");
    uint64_t gda[4] = {(uint64_t)va.dims[0], (uint64_t)va.dims[1], (uint64_t)va.dims[2], (uint64_t)va.dims[3]};
    uint64_t gsa[3] = {(uint64_t)va.strides[0] * 2, (uint64_t)va.strides[1] * 2, (uint64_t)va.strides[2] * 2};
    int rc = encode_tensor_map(&d.amap[i], dt, 4, va.ptr, gda, gsa, box, false);
    if (rc) return rc;
    uint64_t gdb[4] = {(uint64_t)vb.dims[0], (uint64_t)vb.dims[1], (uint64_t)vb.dims[2], (uint64_t)vb.dims[3]};
    uint64_t gsb[3] = {(uint64_t)vb.strides[0] * 2, (uint64_t)vb.strides[1] * 2, (uint64_t)vb.strides[2] * 2};
    rc = encode_tensor_map(&d.bmap[i], dt, 4, vb.ptr, gdb, gsb, box, false);
    if (rc) return rc;
    if (slab) {
      const uint32_t sbox[4] = {64, (uint32_t)h.tw, (uint32_t)h.th + 2, (uint32_t)h.tf};
      rc = encode_tensor_map(&d.bslab[i], dt, 4, vb.ptr, gdb, gsb, sbox, false);
      if (rc) return rc;
    }
  }
  const size_t smem = static_cast<size_t>(d.stages) * d.stage_bytes + sizeof(WgSmemCtl) + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return set_cuda_error(e, "vmm_wgrad: cudaFuncSetAttribute");
    attr_set = true;
  }
  const int grid = d.total_items < num_sms() ? d.total_items : num_sms();
  {
    cudaError_t le = launch_maybe_pdl(wgrad_kernel, dim3(grid), dim3(256), smem, stream, d);
    if (le != cudaSuccess) return set_cuda_error(le, "vmm_wgrad: launch");
  }
  count_launch();
  return check_launch("vmm_wgrad");
}

extern "C" int vmm_colsum(const void* x, long long rows, int n, long long ld, int fmt, float* out, void* stream_) {
  if (!x || !out || n % 8 || n > 2048) return set_error(VMM_ERR_ARG, "vmm_colsum: bad arguments");
  const int vpr = n / 8;
  int threads = 256;
  if (threads % vpr) threads = (256 / vpr) * vpr;
  if (threads < vpr) threads = vpr;
  if (threads > 1024) return set_error(VMM_ERR_UNSUPPORTED, "vmm_colsum: n too large");
  const int trows = threads / vpr;
  int ctas = 4 * num_sms();
  long long rpc = (rows + ctas - 1) / ctas;
  if (rpc < trows * 4) rpc = trows * 4;
  ctas = static_cast<int>((rows + rpc - 1) / rpc);
  colsum_kernel<<<ctas, threads, static_cast<size_t>(trows) * n * sizeof(float), static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const uint16_t*>(x), rows, n, ld, fmt, out, static_cast<int>(rpc));
  count_launch();
  return check_launch("vmm_colsum");
}
