// vmm_cgemm: persistent, warp-specialised implicit-GEMM convolution for sm_100a.
//
//   warp 0      TMA producer (one ELECTED lane, so that descriptors stay in uniform registers): per K block one 4-D box of A
//               (shifted by the tap) + one 2-D box of W; in halo mode one slab of th + 2 pixel rows per (64-channel chunk, kx)
//               that serves the three ky taps, with the weights resident in shared memory when all of them fit
//   warp 1      MMA issuer (one elected lane): tcgen05.mma kind::f16, M=128, N=BN, K=16 x4 per 64-wide block
//   warp 2      TMEM allocator
//   warps 4..   epilogue, two warps per TMEM lane quarter, each half of the tile's columns: tcgen05.ld -> alpha / bias /
//               residual / rotary / GroupNorm partial sums -> 16-bit rows staged in shared memory (64-byte swizzle) -> one bulk
//               tensor store per warp and 32 columns (TMA clips the tile overhang); fp32 outputs and odd column splits keep
//               per-thread global stores.  8 warps (kernel<FMT, 1>), or 16 in two groups that drain alternate tiles
//               (kernel<FMT, 2>, 64 < BN <= 128).  epilogue_fast is the lean path of the common launch shape.
//
// Two TMEM accumulators (four when 4 x BN <= 512 columns) let the epilogue of tile i overlap the main loop of the next tiles.
// The ring of smem stages is shared across tiles (the producer runs ahead of the MMA warp).
//
// Debug / experiment switches (environment, read once): VMM_NO_HALO, VMM_NO_FAST_EPI, VMM_ONE_EPI_GROUP, VMM_TWO_ACC,
// VMM_TWO_GROUPS_64 (two epilogue groups also at BN = 64; measured in round 2: 87.9 vs 84.1 us for the level-0 3x3, i.e. no gain:
// with GroupNorm statistics and bias removed the same launch takes 80 us, so the epilogue is not what bounds it).
#include "common.cuh"
#include "mma_sync.cuh"
#include "sm100_ptx.cuh"

#include <algorithm>

namespace vmm {

struct TapDev {
  int16_t src, dy, dx, nk;
  int32_t kofs;
};

constexpr int kHaloMaxChunks = 8;

struct CgemmDev {
  CUtensorMap amap[VMM_MAX_VIEWS];
  CUtensorMap bmap;
  CUtensorMap omap[VMM_MAX_PHASES];   // output view of each phase (box = the 32 pixels of one epilogue warp x 32 columns)
  CUtensorMap omap2;                  // columns >= nsplit (single-phase launches only)
  int tstore;                         // epilogue uses bulk tensor stores
  TapDev taps[VMM_MAX_PHASES][VMM_MAX_TAPS];
  int n_taps[VMM_MAX_PHASES];
  int phase_oy[VMM_MAX_PHASES], phase_ox[VMM_MAX_PHASES];
  int n_phases;
  int N, BN, n_ntiles;
  int BF, OH, OW;
  int tf_log, th_log, tw_log;
  int tiles_f, tiles_y, tiles_x;
  int total_tiles;
  int stages;
  uint32_t stage_bytes, tx_bytes, acc_stride, tmem_cols;
  uint32_t idesc;
  void* out;
  long long ldo;
  int out_fp32;
  int OHs, OWs, sy, sx;
  void* out2;
  long long ldo2;
  int nsplit;
  const float* bias;
  const void* res;
  long long ldr;
  const void* res2;
  long long ldr2;
  float alpha;
  double* gn_stats;
  int gn_gs, gn_groups, fps;
  int fmt;
  // halo mode (3x3 stride-1 taps, tile 1 x th x 8): per (64-channel chunk, kx) one A slab of th+2 pixel rows serves the three ky taps
  uint64_t mg_nt, mg_tx, mg_ty, mg_tf, mg_fps;   // ceil(2^32 / d) + exact-floor multipliers (fdiv below)
  int gs_log;                         // log2(gn_gs) when it is a power of two >= 8, else -1
  int fast_epi;                       // the launch qualifies for epilogue_fast (see there)
  int acc_mask, acc_log;              // TMEM accumulator ring: 2 buffers, or 4 when they fit in the 512 columns
  const float* rot;                   // rotary epilogue tables [2][rot_frames][16][2], see vmm.h
  int rot_frames, rot_cols, rot_qcols;
  uint64_t mg_rfr;                    // fdiv multiplier of rot_frames
  int rot_hw;
  int halo;                           // 0 = generic taps
  int h_chunks;                       // 64-channel chunks over all sources
  int b_resident;                     // all weight tiles stay in shared memory for the whole kernel (single n-tile)
  uint32_t slab_bytes, btile_bytes;
  int16_t h_src[kHaloMaxChunks], h_kb[kHaloMaxChunks];
  int32_t h_kofs[9][kHaloMaxChunks];  // K offset of (tap ky*3+kx, chunk) in the packed weights
};

constexpr int kMaxStages = 8;
constexpr int kABytes = 128 * 128;   // 128 rows x 64 16-bit elements

struct __align__(8) CgemmSmemCtl {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tfull[4];
  uint64_t tempty[4];
  uint64_t bfull;
  uint32_t tmem_base;
  uint32_t pad;
};

// floor(t / d) for 0 <= t < 2^32 / d with m = floor(2^32 / d) + 1 (host: magic_of); d == 1 gives m = 2^32 + 1 -> t
__device__ __forceinline__ int fdiv(int t, uint64_t m) { return static_cast<int>((static_cast<uint64_t>(static_cast<uint32_t>(t)) * m) >> 32); }

__device__ __forceinline__ void decode_tile(const CgemmDev& p, int t, int& phase, int& bf0, int& y0, int& x0, int& n0) {
  int r = fdiv(t, p.mg_nt);
  const int nt = t - r * p.n_ntiles;
  int r2 = fdiv(r, p.mg_tx);
  const int xt = r - r2 * p.tiles_x;
  r = fdiv(r2, p.mg_ty);
  const int yt = r2 - r * p.tiles_y;
  phase = fdiv(r, p.mg_tf);
  const int ft = r - phase * p.tiles_f;
  bf0 = ft << p.tf_log;
  y0 = yt << p.th_log;
  x0 = xt << p.tw_log;
  n0 = nt * p.BN;
}

constexpr int kBiasSmem = 1024;

constexpr int kGnGroups = 8;   // GroupNorm groups per n-tile (BN is capped accordingly on the host)

// GroupNorm partial sums live in per-row shared-memory slots s_racc[group][row][sum|sumsq]: the epilogue thread that
// owns (row, column half) adds to them with plain loads/stores (no shuffles, no atomics per tile).  gn_flush reduces the
// 128 rows of each group, separately for the (at most two) samples the rows of the current key belong to, and adds the
// result to the global fp64 statistics.  It is called by all 256 epilogue threads after a named-barrier sync, only when
// the CTA moves to another (first sample, n-tile, frame tile if the tile straddles two samples) key.
constexpr int kEpiThreads = 256;

__device__ __forceinline__ void gn_flush(const CgemmDev& p, float (*racc)[128][2], int ethread, int smp0, int n0, int key_bf0, int bar_id = 1) {
  {
    const int o = ethread >> 3, part = ethread & 7;            // 32 outputs x 8 partial sums
    const int sl = o >> 4, gl = (o >> 1) & (kGnGroups - 1), w = o & 1;
    float val = 0.f;
    if (sl == 0 || key_bf0 >= 0) {
#pragma unroll 4
      for (int i = 0; i < 16; ++i) {
        const int row = part * 16 + ((i + part) & 15);           // skewed: the 8 parts hit different banks
        int rs = 0;
        if (key_bf0 >= 0) rs = fdiv(key_bf0 + (row >> (p.tw_log + p.th_log)), p.mg_fps) - smp0;
        if (rs == sl) val += racc[gl][row][w];
      }
    }
    val += __shfl_xor_sync(0xffffffffu, val, 1);
    val += __shfl_xor_sync(0xffffffffu, val, 2);
    val += __shfl_xor_sync(0xffffffffu, val, 4);
    const int g = (p.gs_log >= 0 ? (n0 >> p.gs_log) : n0 / p.gn_gs) + gl;
    const int nsamp = fdiv(p.BF + p.fps - 1, p.mg_fps);
    if (part == 0 && g < p.gn_groups && smp0 + sl < nsamp && val != 0.f)
      atomicAdd(p.gn_stats + (static_cast<long long>(smp0 + sl) * p.gn_groups + g) * 2 + w, static_cast<double>(val));
  }
  asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
}

__device__ __forceinline__ void gn_thread_add(float (*racc)[128][2], int gl, int row, float a1, float a2) {
  float2* q = reinterpret_cast<float2*>(&racc[gl][row][0]);
  float2 v = *q;
  v.x += a1;
  v.y += a2;
  *q = v;
}

// Lean epilogue for the common launch shape: 16-bit output through bulk tensor stores, every 32-column step full
// (N % 32 == 0), bias in shared memory, GroupNorm groups of 8 / 16 / 32 / 64 columns, 16-byte aligned residual rows.
// Same arithmetic and the same shared-memory / barrier protocol as the generic column loop in cgemm_kernel, with every
// launch-invariant decision hoisted out of the tile and column loops.
// G = 2: two groups of eight epilogue warps take alternate tiles of the CTA (own GroupNorm slots, own named barrier), which
// doubles the drain rate when the per-tile epilogue latency, not the MMA, bounds a narrow tile.
template <int FMT, int G>
__device__ __forceinline__ void epilogue_fast(const CgemmDev& p, CgemmSmemCtl* ctl, const uint32_t tmem_base, float (*s_gn)[128][2],
                                              const float* s_bias, uint8_t* stage, const int warp, const int lane, const bool leader,
                                              const int eg) {
  const int q = warp & 3;
  const int half = ((warp - 4) >> 2) & 1;
  const int row = q * 32 + lane;
  const int ethread = (threadIdx.x - 128) & 255;
  const int bar_id = 1 + eg;
  const bool has_bias = p.bias != nullptr, has_gn = p.gn_stats != nullptr, has_res = p.res != nullptr, split = p.out2 != nullptr;
  const bool scaled = p.alpha != 1.f;
  const bool has_rot = p.rot != nullptr;
  const float alpha = p.alpha;
  const int nsteps = (p.BN + 31) >> 5;
  int hs = (nsteps + 1) >> 1;
  if (has_gn && p.gn_gs > 32 && ((hs * 32) % p.gn_gs) != 0) hs = nsteps;
  const int st_lo = half ? hs : 0;
  const int st_hi = half ? nsteps : hs;
  const int tw_m = (1 << p.tw_log) - 1, th_m = (1 << p.th_log) - 1, tyx_log = p.tw_log + p.th_log;
  const int xl = row & tw_m, yl = (row >> p.tw_log) & th_m, fl = row >> tyx_log;
  const int wxo = (q * 32) & tw_m, wyo = ((q * 32) >> p.tw_log) & th_m, wfo = (q * 32) >> tyx_log;
  const int gs = p.gn_gs, gs_log = p.gs_log, gs_m = p.gn_gs - 1;
  uint8_t* srow = stage + lane * 64;
  const int sw = (lane >> 1) & 3;
  const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
  int it = 0;
  int gn_key_smp = -1, gn_key_n0 = -1, gn_key_bf0 = -1;
  for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
    if (G == 2 && (it & 1) != eg) continue;
    int phase, bf0, y0, x0, n0;
    decode_tile(p, t, phase, bf0, y0, x0, n0);
    const int acc = it & p.acc_mask;
    const uint32_t acc_ph = (it >> p.acc_log) & 1;
    const int bf = bf0 + fl, y = y0 + yl, x = x0 + xl;
    const bool valid = (bf < p.BF) && (y < p.OH) && (x < p.OW);
    const int wx = x0 + wxo, wy = y0 + wyo, wf = bf0 + wfo;
    const float* rot_row = nullptr;      // rotary tables of this row's frame (rows GEMM: x = position index (b, f, pixel))
    if (has_rot) {
      const int q1 = x / p.rot_hw;        // exact division: position * hw exceeds the 2^32 range of fdiv at level 0
      const int fr = q1 - fdiv(q1, p.mg_rfr) * p.rot_frames;
      rot_row = p.rot + fr * 32;
    }
    if (has_gn) {
      const int smp0 = fdiv(bf0, p.mg_fps);
      const int last_bf = min(bf0 + (1 << p.tf_log), p.BF) - 1;
      const int kbf0 = (p.tf_log > 0 && fdiv(last_bf, p.mg_fps) != smp0) ? bf0 : -1;
      if (smp0 != gn_key_smp || n0 != gn_key_n0 || kbf0 != gn_key_bf0) {
        if (gn_key_smp >= 0) {
          asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
          gn_flush(p, s_gn, ethread, gn_key_smp, gn_key_n0, gn_key_bf0, bar_id);
          for (int gl = 0; gl < kGnGroups; ++gl) {
            const int st_g = (gl * gs) >> 5;
            if (st_g >= st_lo && st_g < st_hi) {
              s_gn[gl][row][0] = 0.f;
              s_gn[gl][row][1] = 0.f;
            }
          }
        }
        gn_key_smp = smp0;
        gn_key_n0 = n0;
        gn_key_bf0 = kbf0;
      }
    }
    const uint16_t* rrow1 = nullptr;
    const uint16_t* rrow2 = nullptr;
    uint4 rq[4];
    const int ncol_lo = n0 + st_lo * 32;
    if (has_res && valid) {
      const long long pix =
          (static_cast<long long>(bf) * p.OHs + (y * p.sy + p.phase_oy[phase])) * p.OWs + (x * p.sx + p.phase_ox[phase]);
      rrow1 = reinterpret_cast<const uint16_t*>(p.res) + pix * p.ldr;
      rrow2 = split ? reinterpret_cast<const uint16_t*>(p.res2) + pix * p.ldr2 - p.nsplit : rrow1;
      if (st_lo < st_hi && ncol_lo < p.N) {
        const uint4* rp = reinterpret_cast<const uint4*>(((split && ncol_lo >= p.nsplit) ? rrow2 : rrow1) + ncol_lo);
#pragma unroll
        for (int j = 0; j < 4; ++j) rq[j] = __ldg(rp + j);
      }
    }
    mbar_wait(&ctl->tfull[acc], acc_ph);
    tc_fence_after();
    const uint32_t t_addr = lane_taddr + acc * p.acc_stride;
    float gs1 = 0.f, gs2 = 0.f;
    for (int st = st_lo; st < st_hi; ++st) {
      const int ncol = n0 + st * 32;
      if (ncol >= p.N) break;          // padded columns of the last n-tile (uniform)
      float v[32];
      tmem_ld32f(t_addr + st * 32, v);
      tmem_ld_wait();
      if (has_bias) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[ncol + j]);
          v[j] += b4.x;
          v[j + 1] += b4.y;
          v[j + 2] += b4.z;
          v[j + 3] += b4.w;
        }
      }
      if (scaled) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= alpha;
      }
      if (rrow1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t w4[4] = {rq[j].x, rq[j].y, rq[j].z, rq[j].w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 f = unpack2<FMT>(w4[k]);
            v[j * 8 + 2 * k] += f.x;
            v[j * 8 + 2 * k + 1] += f.y;
          }
        }
        if (st + 1 < st_hi && ncol + 32 < p.N) {
          const uint4* rp = reinterpret_cast<const uint4*>(((split && ncol + 32 >= p.nsplit) ? rrow2 : rrow1) + ncol + 32);
#pragma unroll
          for (int j = 0; j < 4; ++j) rq[j] = __ldg(rp + j);
        }
      }
      if (has_rot && ncol < p.rot_cols) {
        // one head slice per step: pair i of the slice turns by the angle (cos, sin) = table[frame][i]
        const float4* tb = reinterpret_cast<const float4*>(rot_row + (ncol < p.rot_qcols ? 0 : p.rot_frames * 32));
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 cs = __ldg(tb + j);          // (cos, sin) of pairs 2j and 2j + 1
          const float a0 = v[4 * j], b0 = v[4 * j + 1], a1 = v[4 * j + 2], b1 = v[4 * j + 3];
          v[4 * j] = a0 * cs.x - b0 * cs.y;
          v[4 * j + 1] = b0 * cs.x + a0 * cs.y;
          v[4 * j + 2] = a1 * cs.z - b1 * cs.w;
          v[4 * j + 3] = b1 * cs.z + a1 * cs.w;
        }
      }
      uint32_t w[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) w[j] = pack2<FMT>(v[2 * j], v[2 * j + 1]);
      if (leader) bulk_wait_read0();     // the previous bulk store of this warp has read the staging rows
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(srow + ((j ^ sw) << 4)) = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (leader) {
        const bool second = split && ncol >= p.nsplit;
        tma_store_4d(second ? &p.omap2 : &p.omap[phase], stage, second ? ncol - p.nsplit : ncol, wx, wy, wf);
        bulk_commit();
      }
      if (has_gn) {
        // statistics from the fp32 values before the 16-bit rounding, 8-column blocks first
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          float a1 = 0.f, a2 = 0.f;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float xv = v[b * 8 + k];
            a1 += xv;
            a2 = fmaf(xv, xv, a2);
          }
          gs1 += valid ? a1 : 0.f;
          gs2 += valid ? a2 : 0.f;
          const int cend = ncol + 8 * (b + 1);
          if ((cend & gs_m) == 0) {          // group complete (uniform across the warp)
            gn_thread_add(s_gn, (cend - gs - n0) >> gs_log, row, gs1, gs2);
            gs1 = gs2 = 0.f;
          }
        }
      }
    }
    tc_fence_before();
    mbar_arrive(&ctl->tempty[acc]);
  }
  if (has_gn && gn_key_smp >= 0) {
    asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
    gn_flush(p, s_gn, ethread, gn_key_smp, gn_key_n0, gn_key_bf0, bar_id);
  }
  if (leader) bulk_wait0();   // shared memory must outlive the last bulk store
}

template <int FMT, int G>
__global__ void __launch_bounds__(128 + 256 * G, 1) cgemm_kernel(const __grid_constant__ CgemmDev p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) float s_gn[G][kGnGroups][128][2];   // per-row GroupNorm partial sums (per epilogue group), see gn_flush
  __shared__ __align__(16) float s_bias[kBiasSmem];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bres = smem + static_cast<size_t>(p.stages) * p.stage_bytes;      // resident weight tiles (halo mode)
  const size_t bres_bytes = p.b_resident ? static_cast<size_t>(9) * p.h_chunks * p.btile_bytes : 0;
  CgemmSmemCtl* ctl = reinterpret_cast<CgemmSmemCtl*>(bres + bres_bytes);
  // per epilogue warp: 32 rows x 32 columns (16-bit) staged for the bulk tensor store, 64-byte swizzle
  uint8_t (*s_stage)[32 * 64] = reinterpret_cast<uint8_t (*)[32 * 64]>(
      (reinterpret_cast<uintptr_t>(ctl) + sizeof(CgemmSmemCtl) + 1023) & ~uintptr_t(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_trigger();

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < VMM_MAX_VIEWS; ++i) tma_prefetch_desc(&p.amap[i]);
    tma_prefetch_desc(&p.bmap);
    if (p.tstore) {
      for (int i = 0; i < p.n_phases; ++i) tma_prefetch_desc(&p.omap[i]);
      if (p.out2) tma_prefetch_desc(&p.omap2);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&ctl->full[s], 1);
      mbar_init(&ctl->empty[s], 1);
    }
    for (int a = 0; a < 4; ++a) {
      mbar_init(&ctl->tfull[a], 1);
      mbar_init(&ctl->tempty[a], kEpiThreads);
    }
    mbar_init(&ctl->bfull, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(&ctl->tmem_base, p.tmem_cols);
    tmem_relinquish();
  }
  if (warp == 3) {
    float* g = &s_gn[0][0][0][0];
    for (int i = lane; i < G * kGnGroups * 128 * 2; i += 32) g[i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();        // launched with programmatic serialisation (VMM_PDL): everything above overlapped the previous kernel's tail
  const uint32_t tmem_base = ctl->tmem_base;
  // one elected lane per warp issues the TMA / MMA / bulk-store instructions: behind elect.sync the compiler knows that a
  // single thread is active and keeps descriptors and addresses in uniform registers (no per-instruction waterfall loop)
  const bool leader = elect_one();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (leader && p.halo) {
      if (p.b_resident) {
        mbar_expect_tx(&ctl->bfull, static_cast<uint32_t>(bres_bytes));
        for (int tap = 0; tap < 9; ++tap)
          for (int c = 0; c < p.h_chunks; ++c)
            tma_load_2d(bres + static_cast<size_t>(tap * p.h_chunks + c) * p.btile_bytes, &p.bmap, &ctl->bfull, p.h_kofs[tap][c], 0);
      }
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        int phase, bf0, y0, x0, n0;
        decode_tile(p, t, phase, bf0, y0, x0, n0);
        for (int c = 0; c < p.h_chunks; ++c) {
          for (int kx = 0; kx < 3; ++kx) {
            mbar_wait(&ctl->empty[s], ph ^ 1);
            uint8_t* a_s = smem + static_cast<size_t>(s) * p.stage_bytes;
            mbar_expect_tx(&ctl->full[s], p.tx_bytes);
            tma_load_4d(a_s, &p.amap[p.h_src[c]], &ctl->full[s], p.h_kb[c] * 64, x0 + kx - 1, y0 - 1, bf0);
            if (!p.b_resident) {
              for (int ky = 0; ky < 3; ++ky)
                tma_load_2d(a_s + p.slab_bytes + ky * p.btile_bytes, &p.bmap, &ctl->full[s], p.h_kofs[ky * 3 + kx][c], n0);
            }
            if (++s == p.stages) {
              s = 0;
              ph ^= 1;
            }
          }
        }
      }
    } else if (leader) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        int phase, bf0, y0, x0, n0;
        decode_tile(p, t, phase, bf0, y0, x0, n0);
        const int nt = p.n_taps[phase];
        for (int tap = 0; tap < nt; ++tap) {
          const TapDev T = p.taps[phase][tap];
          for (int kb = 0; kb < T.nk; ++kb) {
            mbar_wait(&ctl->empty[s], ph ^ 1);
            uint8_t* a_s = smem + static_cast<size_t>(s) * p.stage_bytes;
            mbar_expect_tx(&ctl->full[s], p.tx_bytes);
            tma_load_4d(a_s, &p.amap[T.src], &ctl->full[s], kb * 64, x0 + T.dx, y0 + T.dy, bf0);
            tma_load_2d(a_s + kABytes, &p.bmap, &ctl->full[s], T.kofs + kb * 64, n0);
            if (++s == p.stages) {
              s = 0;
              ph ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    int s = 0;
    uint32_t ph = 0;
    int it = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
      int phase, bf0, y0, x0, n0;
      decode_tile(p, t, phase, bf0, y0, x0, n0);
      const int acc = it & p.acc_mask;
      const uint32_t acc_ph = (it >> p.acc_log) & 1;
      mbar_wait(&ctl->tempty[acc], acc_ph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * p.acc_stride;
      const int nt = p.halo ? 0 : p.n_taps[phase];
      uint32_t accumulate = 0;
      if (p.halo) {
        if (p.b_resident && it == 0) {
          mbar_wait(&ctl->bfull, 0);
          tc_fence_after();
        }
        for (int c = 0; c < p.h_chunks; ++c) {
          for (int kx = 0; kx < 3; ++kx) {
            mbar_wait(&ctl->full[s], ph);
            tc_fence_after();
            if (leader) {
              const uint32_t a_addr = smem_u32(smem + static_cast<size_t>(s) * p.stage_bytes);
#pragma unroll
              for (int ky = 0; ky < 3; ++ky) {
                // rows of tap ky = slab rows shifted by ky pixel rows (8 pixels x 128 bytes = one swizzle atom each)
                const uint64_t adesc = make_smem_desc_sw128(a_addr + ky * 1024, 16, 1024);
                const uint32_t b_addr = p.b_resident ? smem_u32(bres + static_cast<size_t>((ky * 3 + kx) * p.h_chunks + c) * p.btile_bytes)
                                                     : a_addr + p.slab_bytes + ky * p.btile_bytes;
                const uint64_t bdesc = make_smem_desc_sw128(b_addr, 16, 1024);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  umma_f16(d_tmem, adesc + static_cast<uint64_t>(k * 2), bdesc + static_cast<uint64_t>(k * 2), p.idesc, accumulate);
                  accumulate = 1;
                }
              }
              umma_commit(&ctl->empty[s]);
            }
            __syncwarp();
            if (++s == p.stages) {
              s = 0;
              ph ^= 1;
            }
          }
        }
      }
      for (int tap = 0; tap < nt; ++tap) {
        const int nk = p.taps[phase][tap].nk;
        for (int kb = 0; kb < nk; ++kb) {
          mbar_wait(&ctl->full[s], ph);
          tc_fence_after();
          if (leader) {
            const uint32_t a_addr = smem_u32(smem + static_cast<size_t>(s) * p.stage_bytes);
            const uint64_t adesc = make_smem_desc_sw128(a_addr, 16, 1024);
            const uint64_t bdesc = make_smem_desc_sw128(a_addr + kABytes, 16, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_f16(d_tmem, adesc + static_cast<uint64_t>(k * 2), bdesc + static_cast<uint64_t>(k * 2), p.idesc,
                       accumulate);
              accumulate = 1;
            }
            umma_commit(&ctl->empty[s]);   // frees the stage once these MMAs have read it
          }
          __syncwarp();
          if (++s == p.stages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
      if (leader) umma_commit(&ctl->tfull[acc]);
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3;              // TMEM lane quarter this warp may read
    const int half = (warp - 4) >> 2;    // which half of the tile's columns this warp drains
    const int ew = warp - 4;             // staging buffer
    const int row = q * 32 + lane;       // row of the 128-row tile == TMEM lane
    const int ethread = threadIdx.x - 128;
    // bias lives in shared memory for the whole kernel (global loads in the column loop were the bottleneck)
    const bool bias_smem = p.bias != nullptr && p.N <= kBiasSmem;
    if (bias_smem) {
      for (int i = ethread; i < p.N; i += kEpiThreads * G) s_bias[i] = __ldg(p.bias + i);
      if (G == 1) asm volatile("bar.sync 1, 256;" ::: "memory");
      else asm volatile("bar.sync 3, 512;" ::: "memory");
    }
    if (G == 2) {
      // launched only for fast_epi shapes (host): the generic column loop is not instantiated here
      epilogue_fast<FMT, G>(p, ctl, tmem_base, s_gn[(warp - 4) >> 3], s_bias, &s_stage[ew][0], warp, lane, leader, (warp - 4) >> 3);
    } else if (p.fast_epi) {
      epilogue_fast<FMT, 1>(p, ctl, tmem_base, s_gn[0], s_bias, &s_stage[ew][0], warp, lane, leader, 0);
    } else if (G == 1) {
    // 32-column steps; the two warps of a lane quarter split them in two contiguous halves when the boundary does not
    // cut a GroupNorm group (else the first warp takes all of them)
    const int nsteps = (p.BN + 31) >> 5;
    int hs = (nsteps + 1) >> 1;
    if (p.gn_stats && p.gn_gs > 32 && ((hs * 32) % p.gn_gs) != 0) hs = nsteps;
    const int st_lo = half ? hs : 0;
    const int st_hi = half ? nsteps : hs;
    int it = 0;
    int gn_key_smp = -1, gn_key_n0 = -1, gn_key_bf0 = -1;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
      int phase, bf0, y0, x0, n0;
      decode_tile(p, t, phase, bf0, y0, x0, n0);
      const int acc = it & p.acc_mask;
      const uint32_t acc_ph = (it >> p.acc_log) & 1;

      const int xl = row & ((1 << p.tw_log) - 1);
      const int yl = (row >> p.tw_log) & ((1 << p.th_log) - 1);
      const int fl = row >> (p.tw_log + p.th_log);
      const int bf = bf0 + fl, y = y0 + yl, x = x0 + xl;
      const bool valid = (bf < p.BF) && (y < p.OH) && (x < p.OW);
      // first pixel of this warp's 32 rows (origin of its store box)
      const int wx = x0 + ((q * 32) & ((1 << p.tw_log) - 1));
      const int wy = y0 + (((q * 32) >> p.tw_log) & ((1 << p.th_log) - 1));
      const int wf = bf0 + ((q * 32) >> (p.tw_log + p.th_log));
      const long long pix =
          (static_cast<long long>(bf) * p.OHs + (y * p.sy + p.phase_oy[phase])) * p.OWs + (x * p.sx + p.phase_ox[phase]);

      // GroupNorm bookkeeping.  Partial sums stay in shared memory while consecutive tiles of this CTA belong
      // to the same (first sample, n-tile); they go to global memory (fp64 atomics) only when that key changes,
      // which keeps the number of same-address atomics per launch at O(CTAs x samples) instead of O(tiles).
      if (p.gn_stats) {
        const int smp0 = fdiv(bf0, p.mg_fps);
        const int last_bf = min(bf0 + (1 << p.tf_log), p.BF) - 1;
        const int kbf0 = (p.tf_log > 0 && fdiv(last_bf, p.mg_fps) != smp0) ? bf0 : -1;     // tile straddles two samples
        if (smp0 != gn_key_smp || n0 != gn_key_n0 || kbf0 != gn_key_bf0) {
          if (gn_key_smp >= 0) {
            asm volatile("bar.sync 1, 256;" ::: "memory");
            gn_flush(p, s_gn[0], ethread, gn_key_smp, gn_key_n0, gn_key_bf0);
            // each thread clears the slots it adds to: its row, the groups of its column half
            for (int gl = 0; gl < kGnGroups; ++gl) {
              const int st_g = (gl * p.gn_gs) >> 5;
              if (st_g >= st_lo && st_g < st_hi) {
                s_gn[0][gl][row][0] = 0.f;
                s_gn[0][gl][row][1] = 0.f;
              }
            }
          }
          gn_key_smp = smp0;
          gn_key_n0 = n0;
          gn_key_bf0 = kbf0;
        }
      }

      // residual rows are known before the accumulator is: fetch the first 32 columns while the MMAs still run.
      // With a column split (out2), columns >= nsplit read res2 (same split as the outputs).
      const bool split = p.out2 != nullptr;
      const uint16_t* rrow1 = (p.res && valid) ? reinterpret_cast<const uint16_t*>(p.res) + pix * p.ldr : nullptr;
      const uint16_t* rrow2 = (p.res && valid && split) ? reinterpret_cast<const uint16_t*>(p.res2) + pix * p.ldr2 - p.nsplit : nullptr;
      const bool res_al = ((p.ldr & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.res) & 15) == 0) &&
                          (!split || (((p.ldr2 & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.res2) & 15) == 0)));
      uint4 rq[4];
      auto res_ptr = [&](int col) -> const uint16_t* { return (split && col >= p.nsplit) ? rrow2 + col : rrow1 + col; };
      auto straddles = [&](int col) -> bool { return split && col < p.nsplit && col + 32 > p.nsplit; };
      const int ncol_lo = n0 + st_lo * 32;
      if (rrow1 && res_al && st_lo < st_hi && ncol_lo + 32 <= p.N && !straddles(ncol_lo)) {
        const uint4* rp = reinterpret_cast<const uint4*>(res_ptr(ncol_lo));
#pragma unroll
        for (int j = 0; j < 4; ++j) rq[j] = __ldg(rp + j);
      }

      mbar_wait(&ctl->tfull[acc], acc_ph);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * p.acc_stride;

      float gs1 = 0.f, gs2 = 0.f;   // running sums of the current GroupNorm group
      for (int st = st_lo; st < st_hi; ++st) {
        const int ncol = n0 + st * 32;
        uint32_t r[32];
        tmem_ld16(t_addr + st * 32, r);
        tmem_ld16(t_addr + st * 32 + 16, r + 16);
        tmem_ld_wait();
        if (ncol >= p.N) continue;   // uniform: padded columns of the last n-tile
        const bool full = (ncol + 32 <= p.N);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (p.bias) {
          if (bias_smem && full) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[ncol + j]);
              v[j] += b4.x;
              v[j + 1] += b4.y;
              v[j + 2] += b4.z;
              v[j + 3] += b4.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (ncol + j < p.N) v[j] += bias_smem ? s_bias[ncol + j] : __ldg(p.bias + ncol + j);
          }
        }
        const bool strad = straddles(ncol);
        if (p.alpha != 1.f) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= p.alpha;
        }
        if (rrow1) {
          if (res_al && full && !strad) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t w4[4] = {rq[j].x, rq[j].y, rq[j].z, rq[j].w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float2 f = unpack2<FMT>(w4[k]);
                v[j * 8 + 2 * k] += f.x;
                v[j * 8 + 2 * k + 1] += f.y;
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (ncol + j < p.N) v[j] += h16_to_f(*res_ptr(ncol + j), FMT);
          }
          // prefetch the next 32 columns
          if (res_al && st + 1 < st_hi && ncol + 64 <= p.N && !straddles(ncol + 32)) {
            const uint4* rp = reinterpret_cast<const uint4*>(res_ptr(ncol + 32));
#pragma unroll
            for (int j = 0; j < 4; ++j) rq[j] = __ldg(rp + j);
          }
        }
        // destination (column split for fused concat gradients)
        void* obase = p.out;
        long long ld = p.ldo;
        int ocol = ncol;
        if (split && ncol >= p.nsplit) {
          obase = p.out2;
          ld = p.ldo2;
          ocol = ncol - p.nsplit;
        }
        const bool vec_ok = full && !strad;
        if (p.out_fp32) {
          if (valid) {
            float* op = reinterpret_cast<float*>(obase) + pix * ld + ocol;
            if (vec_ok && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                reinterpret_cast<float4*>(op)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (ncol + j < p.N) {
                  if (split && ncol + j >= p.nsplit)
                    (reinterpret_cast<float*>(p.out2) + pix * p.ldo2)[ncol + j - p.nsplit] = v[j];
                  else
                    (reinterpret_cast<float*>(p.out) + pix * p.ldo)[ncol + j] = v[j];
                }
            }
          }
        } else {
          uint32_t w[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) w[j] = pack2<FMT>(v[2 * j], v[2 * j + 1]);
          if (p.tstore) {
            // the previous bulk store of this warp must have read the staging rows before they are overwritten
            if (leader) bulk_wait_read0();
            __syncwarp();
            uint8_t* srow = &s_stage[ew][lane * 64];
            const int sw = (lane >> 1) & 3;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(srow + ((j ^ sw) << 4)) = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (leader) {
              tma_store_4d((split && ncol >= p.nsplit) ? &p.omap2 : &p.omap[phase], &s_stage[ew][0], ocol, wx, wy, wf);
              bulk_commit();
            }
          } else if (valid) {
            uint16_t* op = reinterpret_cast<uint16_t*>(obase) + pix * ld + ocol;
            if (vec_ok && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
#pragma unroll
              for (int j = 0; j < 4; ++j) reinterpret_cast<uint4*>(op)[j] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (ncol + j < p.N) {
                  const uint16_t hv = static_cast<uint16_t>((w[j >> 1] >> ((j & 1) * 16)) & 0xFFFF);
                  if (split && ncol + j >= p.nsplit)
                    (reinterpret_cast<uint16_t*>(p.out2) + pix * p.ldo2)[ncol + j - p.nsplit] = hv;
                  else
                    (reinterpret_cast<uint16_t*>(p.out) + pix * p.ldo)[ncol + j] = hv;
                }
            }
          }
          if (p.gn_stats) {
            // statistics from the fp32 values before the 16-bit rounding (the rounding error is unbiased and ~2^-9
            // relative per element, far below what mean / variance over >= 10^4 elements can resolve);
            // per 8-column block sums first (static indexing), then group them
            float b1[4], b2[4];
            if (full) {
#pragma unroll
              for (int b = 0; b < 4; ++b) {
                float a1 = 0.f, a2 = 0.f;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                  const float x = v[b * 8 + k];
                  a1 += x;
                  a2 = fmaf(x, x, a2);
                }
                b1[b] = valid ? a1 : 0.f;
                b2[b] = valid ? a2 : 0.f;
              }
            } else {
#pragma unroll
              for (int b = 0; b < 4; ++b) {
                float a1 = 0.f, a2 = 0.f;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                  const float x = v[b * 8 + k];
                  if (ncol + b * 8 + k < p.N) {
                    a1 += x;
                    a2 = fmaf(x, x, a2);
                  }
                }
                b1[b] = valid ? a1 : 0.f;
                b2[b] = valid ? a2 : 0.f;
              }
            }
            const int gs = p.gn_gs;
            if (p.gs_log >= 3) {
              // power-of-two groups of >= 8 columns: no divisions in the column loop
#pragma unroll
              for (int b = 0; b < 4; ++b) {
                gs1 += b1[b];
                gs2 += b2[b];
                const int cend = ncol + 8 * (b + 1);
                if ((cend & (gs - 1)) == 0 && cend - gs < p.N) {   // group complete (uniform across the warp)
                  gn_thread_add(s_gn[0], (cend - gs - n0) >> p.gs_log, row, gs1, gs2);
                  gs1 = gs2 = 0.f;
                }
              }
            } else if (gs >= 8) {
#pragma unroll
              for (int b = 0; b < 4; ++b) {
                gs1 += b1[b];
                gs2 += b2[b];
                const int cend = ncol + 8 * (b + 1);
                if ((cend % gs) == 0 && cend - gs < p.N) {   // group complete (uniform across the warp)
                  gn_thread_add(s_gn[0], (cend - gs - n0) / gs, row, gs1, gs2);
                  gs1 = gs2 = 0.f;
                }
              }
            } else {
              // tiny channel counts (tests): groups of 1, 2 or 4 columns; fully unrolled so that w[] stays in registers
#pragma unroll
              for (int cj = 0; cj < 32; ++cj) {
                if (ncol + cj < p.N && valid) {
                  gs1 += v[cj];
                  gs2 = fmaf(v[cj], v[cj], gs2);
                }
                if (((cj + 1) % gs) == 0) {
                  const int c0g = ncol + cj + 1 - gs;
                  const int gl = (c0g - n0) / gs;
                  if (c0g < p.N && gl < kGnGroups) gn_thread_add(s_gn[0], gl, row, gs1, gs2);
                  gs1 = gs2 = 0.f;
                }
              }
            }
          }
        }
      }
      // accumulator drained: hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      mbar_arrive(&ctl->tempty[acc]);
    }
    if (p.gn_stats && gn_key_smp >= 0) {
      asm volatile("bar.sync 1, 256;" ::: "memory");
      gn_flush(p, s_gn[0], ethread, gn_key_smp, gn_key_n0, gn_key_bf0);
    }
    if (p.tstore && leader) bulk_wait0();   // shared memory must outlive the last bulk store
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
static uint64_t magic_of(int d) { return (uint64_t(1) << 32) / static_cast<uint64_t>(d) + 1; }

static int ilog2_exact(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return ((1 << l) == v) ? l : -1;
}

}  // namespace vmm

using namespace vmm;

extern "C" int vmm_cgemm(const vmm_cgemm_params* hp, void* stream_) {
  if (!hp) return set_error(VMM_ERR_ARG, "vmm_cgemm: null params");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const vmm_cgemm_params& h = *hp;
  if (h.fmt != VMM_FMT_F16 && h.fmt != VMM_FMT_BF16) return set_error(VMM_ERR_ARG, "vmm_cgemm: bad fmt");
  if (h.n_views < 1 || h.n_views > VMM_MAX_VIEWS) return set_error(VMM_ERR_ARG, "vmm_cgemm: n_views");
  if (h.n_phases < 1 || h.n_phases > VMM_MAX_PHASES) return set_error(VMM_ERR_ARG, "vmm_cgemm: n_phases");
  const int tfl = ilog2_exact(h.tf), thl = ilog2_exact(h.th), twl = ilog2_exact(h.tw);
  if (tfl < 0 || thl < 0 || twl < 0 || h.tf * h.th * h.tw != 128)
    return set_error(VMM_ERR_ARG, "vmm_cgemm: tile must be powers of two with tf*th*tw == 128");
  if (h.n < 1 || h.ktot < 64 || (h.ktot % 64) != 0) return set_error(VMM_ERR_ARG, "vmm_cgemm: n / ktot");
  if (!h.out || !h.w) return set_error(VMM_ERR_ARG, "vmm_cgemm: null out / w");
  if (h.out2 && (h.nsplit % 16) != 0) return set_error(VMM_ERR_ARG, "vmm_cgemm: nsplit must be a multiple of 16");
  if (h.gn_stats) {
    if (h.out_fp32) return set_error(VMM_ERR_ARG, "vmm_cgemm: gn_stats needs a 16-bit output");
    if (h.gn_group < 1 || (h.n % h.gn_group) != 0 || h.frames_per_sample < 1)
      return set_error(VMM_ERR_ARG, "vmm_cgemm: gn_group / frames_per_sample");
    if (h.gn_group < 16 && (16 % h.gn_group) != 0) return set_error(VMM_ERR_ARG, "vmm_cgemm: gn_group < 16 must divide 16");
    if (h.gn_group >= 16 && (h.gn_group % 16) != 0) return set_error(VMM_ERR_ARG, "vmm_cgemm: gn_group must be a multiple of 16");
    if (h.tf > h.frames_per_sample) return set_error(VMM_ERR_ARG, "vmm_cgemm: tile spans more than two samples");
  }

  CgemmDev d;
  memset(&d, 0, sizeof(d));
  const int n_pad = (h.n + 15) / 16 * 16;
  int BN = n_pad <= 256 ? n_pad : 256;
  if (n_pad > 256) {
    // prefer an even split for the common 512-wide case, else 256 + remainder
    BN = 256;
  }
  if (h.gn_stats && h.gn_group >= 16 && (BN % h.gn_group) != 0) {
    // keep groups inside one n-tile
    BN = (BN / h.gn_group) * h.gn_group;
    if (BN == 0) return set_error(VMM_ERR_UNSUPPORTED, "vmm_cgemm: gn_group larger than an n-tile");
  }
  if (h.gn_stats && BN / h.gn_group > kGnGroups) BN = kGnGroups * h.gn_group;
  if (h.gn_stats && (BN % 16) != 0) return set_error(VMM_ERR_UNSUPPORTED, "vmm_cgemm: gn_group too small for a 16-wide n-tile");
  d.N = h.n;
  d.BN = BN;
  d.n_ntiles = (n_pad + BN - 1) / BN;
  d.BF = h.bf;
  d.OH = h.oh;
  d.OW = h.ow;
  d.tf_log = tfl;
  d.th_log = thl;
  d.tw_log = twl;
  d.tiles_f = (h.bf + h.tf - 1) / h.tf;
  d.tiles_y = (h.oh + h.th - 1) / h.th;
  d.tiles_x = (h.ow + h.tw - 1) / h.tw;
  d.n_phases = h.n_phases;
  const long long total = 1LL * d.n_phases * d.tiles_f * d.tiles_y * d.tiles_x * d.n_ntiles;
  if (total <= 0 || total > 0x7fffffffLL) return set_error(VMM_ERR_ARG, "vmm_cgemm: tile count");
  d.total_tiles = static_cast<int>(total);
  d.mg_nt = magic_of(d.n_ntiles);
  d.mg_tx = magic_of(d.tiles_x);
  d.mg_ty = magic_of(d.tiles_y);
  d.mg_tf = magic_of(d.tiles_f);
  {
    // fdiv is exact for t * d < 2^32
    const long long dmax = std::max(std::max(d.n_ntiles, d.tiles_x), std::max(d.tiles_y, d.tiles_f));
    if (total * dmax >= (1LL << 32)) return set_error(VMM_ERR_UNSUPPORTED, "vmm_cgemm: tile grid too large for the fast tile decode");
  }
  // Epilogue flavour, decided before the shared-memory budget because the two-group kernel carries more static memory.
  {
    const bool al = (h.ldo % 8) == 0 && (reinterpret_cast<uintptr_t>(h.out) & 15) == 0;
    const bool al2 = !h.out2 || ((h.ldo2 % 8) == 0 && (reinterpret_cast<uintptr_t>(h.out2) & 15) == 0 && (h.nsplit % 32) == 0 &&
                                 h.n_phases == 1 && h.nsplit > 0 && h.nsplit < h.n);
    d.tstore = !h.out_fp32 && al && al2 && ((BN % 32) == 0 || d.n_ntiles == 1);
    static const bool no_fast = getenv("VMM_NO_FAST_EPI") != nullptr;
    const bool res_al = !h.res || (((h.ldr & 7) == 0) && ((reinterpret_cast<uintptr_t>(h.res) & 15) == 0) &&
                                   (!h.out2 || (((h.ldr2 & 7) == 0) && ((reinterpret_cast<uintptr_t>(h.res2) & 15) == 0))));
    const int gsz = h.gn_stats ? h.gn_group : 0;
    const bool gs_ok = !h.gn_stats || (gsz >= 8 && ilog2_exact(gsz) >= 0);
    d.fast_epi = (!no_fast && d.tstore && (h.n % 32) == 0 && (BN % 32) == 0 && (!h.bias || h.n <= kBiasSmem) && gs_ok && res_al) ? 1 : 0;
  }
  if (h.rot) {
    if (!d.fast_epi) return set_error(VMM_ERR_UNSUPPORTED, "vmm_cgemm: the rotary epilogue needs the fast epilogue (16-bit bulk-store output, N % 32 == 0)");
    if (h.rot_frames < 1 || h.rot_hw < 1 || (h.rot_cols % 32) != 0 || (h.rot_qcols % 32) != 0 || h.rot_qcols > h.rot_cols || h.rot_cols > h.n ||
        h.tf != 1 || h.th != 1 || h.bf != 1 || h.oh != 1)
      return set_error(VMM_ERR_ARG, "vmm_cgemm: rotary epilogue: rows GEMM only, column ranges in multiples of 32");
    d.rot = h.rot;
    d.rot_frames = h.rot_frames;
    d.rot_hw = h.rot_hw;
    d.rot_cols = h.rot_cols;
    d.rot_qcols = h.rot_qcols;
    d.mg_rfr = magic_of(h.rot_frames);
  }
  // Narrow tiles (BN <= 128) are bound by the latency of the per-tile epilogue, not by the MMAs: two groups of eight
  // epilogue warps then drain alternate tiles (4 TMEM accumulators).  Costs 24 KB more static shared memory.
  static const bool one_group = getenv("VMM_ONE_EPI_GROUP") != nullptr;
  static const bool groups64 = getenv("VMM_TWO_GROUPS_64") != nullptr;
  const bool two_groups = !one_group && d.fast_epi && (BN > 64 || (groups64 && BN == 64)) && BN <= 128;   // BN = 64 is bound by the operand reads of the MMAs (48 clk each)
  const int smem_budget = (two_groups ? 170 : 194) * 1024;   // + 12 / 20 KB static (bias, GroupNorm slots) + 16 / 32 KB store staging + control block
  // Halo mode: 3x3 stride-1 taps in (ky, kx, source) order on a 1 x th x 8 tile.  Each (64-channel chunk, kx) stage loads ONE
  // slab of th + 2 pixel rows; the three ky taps read it at +0 / +1 / +2 swizzle atoms.  A traffic drops from 9 to 3.4 tiles
  // per chunk, and the weights stay resident when all of them fit beside the ring.
  size_t bres_bytes = 0;
  {
    static const bool no_halo = getenv("VMM_NO_HALO") != nullptr;
    bool halo = !no_halo && h.n_phases == 1 && h.tf == 1 && h.tw == 8 && h.th == 16 && h.n_taps[0] >= 9 && h.n_taps[0] <= VMM_MAX_TAPS && (h.n_taps[0] % 9) == 0;
    const int nsrc = halo ? h.n_taps[0] / 9 : 0;
    int chunks = 0;
    for (int s = 0; halo && s < nsrc; ++s) {
      const vmm_tap& T0 = h.taps[0][s];
      for (int t9 = 0; t9 < 9 && halo; ++t9) {
        const vmm_tap& T = h.taps[0][t9 * nsrc + s];
        if (T.src != T0.src || T.c != T0.c || T.dy != t9 / 3 - 1 || T.dx != t9 % 3 - 1) halo = false;
      }
      for (int kb = 0; halo && kb < (T0.c + 63) / 64; ++kb) {
        if (chunks >= kHaloMaxChunks) {
          halo = false;
          break;
        }
        d.h_src[chunks] = static_cast<int16_t>(T0.src);
        d.h_kb[chunks] = static_cast<int16_t>(kb);
        for (int t9 = 0; t9 < 9; ++t9) d.h_kofs[t9][chunks] = h.taps[0][t9 * nsrc + s].kofs + kb * 64;
        ++chunks;
      }
    }
    if (halo) {
      d.halo = 1;
      d.h_chunks = chunks;
      d.slab_bytes = static_cast<uint32_t>((h.th + 2) * h.tw * 128);
      d.btile_bytes = static_cast<uint32_t>(BN * 128);
      const size_t all_b = static_cast<size_t>(9) * chunks * d.btile_bytes;
      d.b_resident = (d.n_ntiles == 1 && all_b <= 80 * 1024) ? 1 : 0;
      bres_bytes = d.b_resident ? all_b : 0;
      d.stage_bytes = d.slab_bytes + (d.b_resident ? 0 : 3 * d.btile_bytes);
    }
  }
  if (!d.halo) d.stage_bytes = kABytes + BN * 128;
  d.tx_bytes = d.stage_bytes;
  int stages = (smem_budget - static_cast<int>(bres_bytes)) / static_cast<int>(d.stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (const char* e = getenv("VMM_CGEMM_STAGES")) {      // experiment knob: cap the ring depth
    const int cap = atoi(e);
    if (cap >= 2 && cap < stages) stages = cap;
  }
  if (stages < 2) return set_error(VMM_ERR_UNSUPPORTED, "vmm_cgemm: tile too large for shared memory");
  d.stages = stages;
  d.acc_stride = (BN + 31) / 32 * 32;
  {
    static const bool two_acc = getenv("VMM_TWO_ACC") != nullptr;
    const int nacc = (!two_acc && 4 * d.acc_stride <= 512) ? 4 : 2;     // a deeper ring decouples the MMA warp from the epilogue latency
    d.acc_log = nacc == 4 ? 2 : 1;
    d.acc_mask = nacc - 1;
    uint32_t cols = 32;
    while (cols < nacc * d.acc_stride) cols <<= 1;
    d.tmem_cols = cols;
  }
  d.idesc = make_idesc_f16(128, BN, h.fmt, 0, 0);
  d.out = h.out;
  d.ldo = h.ldo;
  d.out_fp32 = h.out_fp32;
  d.OHs = h.ohs;
  d.OWs = h.ows;
  d.sy = h.sy;
  d.sx = h.sx;
  d.out2 = h.out2;
  d.ldo2 = h.ldo2;
  d.nsplit = h.nsplit;
  d.bias = h.bias;
  d.res = h.res;
  d.ldr = h.ldr;
  d.res2 = h.res2;
  d.ldr2 = h.ldr2;
  d.alpha = h.alpha == 0.f ? 1.f : h.alpha;
  if (h.res && h.out2 && !h.res2) return set_error(VMM_ERR_ARG, "vmm_cgemm: res2 is required when both res and out2 are given");
  d.gn_stats = h.gn_stats;
  d.gn_gs = h.gn_group > 0 ? h.gn_group : 1;
  d.gn_groups = h.gn_stats ? h.n / h.gn_group : 1;
  d.fps = h.frames_per_sample > 0 ? h.frames_per_sample : 1;
  d.mg_fps = magic_of(d.fps);
  d.gs_log = (d.gn_gs >= 8 && ilog2_exact(d.gn_gs) >= 0) ? ilog2_exact(d.gn_gs) : -1;
  d.fmt = h.fmt;

  for (int ph = 0; ph < h.n_phases; ++ph) {
    if (h.n_taps[ph] < 1 || h.n_taps[ph] > VMM_MAX_TAPS) return set_error(VMM_ERR_ARG, "vmm_cgemm: n_taps");
    d.n_taps[ph] = h.n_taps[ph];
    d.phase_oy[ph] = h.phase_oy[ph];
    d.phase_ox[ph] = h.phase_ox[ph];
    for (int t = 0; t < h.n_taps[ph]; ++t) {
      const vmm_tap& T = h.taps[ph][t];
      if (T.src < 0 || T.src >= h.n_views || T.c < 1 || (T.kofs % 64) != 0 || T.kofs + (T.c + 63) / 64 * 64 > h.ktot)
        return set_error(VMM_ERR_ARG, "vmm_cgemm: bad tap");
      d.taps[ph][t].src = static_cast<int16_t>(T.src);
      d.taps[ph][t].dy = static_cast<int16_t>(T.dy);
      d.taps[ph][t].dx = static_cast<int16_t>(T.dx);
      d.taps[ph][t].nk = static_cast<int16_t>((T.c + 63) / 64);
      d.taps[ph][t].kofs = T.kofs;
    }
  }

  const CUtensorMapDataType dt = h.fmt == VMM_FMT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  for (int i = 0; i < VMM_MAX_VIEWS; ++i) {
    const vmm_view4& v = h.a[i < h.n_views ? i : 0];
    if (!v.ptr) return set_error(VMM_ERR_ARG, "vmm_cgemm: null view");
    uint64_t gdim[4] = {(uint64_t)v.dims[0], (uint64_t)v.dims[1], (uint64_t)v.dims[2], (uint64_t)v.dims[3]};
    uint64_t gstr[3] = {(uint64_t)v.strides[0] * 2, (uint64_t)v.strides[1] * 2, (uint64_t)v.strides[2] * 2};
    uint32_t box[4] = {64, (uint32_t)h.tw, (uint32_t)(d.halo ? h.th + 2 : h.th), (uint32_t)h.tf};
    int rc = encode_tensor_map(&d.amap[i], dt, 4, v.ptr, gdim, gstr, box, /*l2_256=*/false);
    if (rc) return rc;
  }
  {
    uint64_t gdim[2] = {(uint64_t)h.ktot, (uint64_t)n_pad};
    uint64_t gstr[1] = {(uint64_t)h.ktot * 2};
    uint32_t box[2] = {64, (uint32_t)BN};
    int rc = encode_tensor_map(&d.bmap, dt, 2, h.w, gdim, gstr, box, /*l2_256=*/true);
    if (rc) return rc;
  }

  // Output tensor maps for the bulk-store epilogue: pixel (bf, y, x) of phase ph lives at
  // out + ((bf*OHs + y*sy + oy)*OWs + x*sx + ox) * ldo, i.e. a strided 4-D view (columns, x, y, bf).
  if (d.tstore) {
    // the 32 rows of one epilogue warp form a sub-box of the (tf, th, tw) pixel tile
    const uint32_t bw = h.tw < 32 ? h.tw : 32;
    const uint32_t bh = (32 / bw) < (uint32_t)h.th ? (32 / bw) : (uint32_t)h.th;
    const uint32_t bfr = 32 / (bw * bh);
    const uint32_t box[4] = {32, bw, bh, bfr};
    for (int ph = 0; ph < h.n_phases; ++ph) {
      const long long ncols = h.out2 ? h.nsplit : h.n;
      uint64_t gdim[4] = {(uint64_t)ncols, (uint64_t)h.ow, (uint64_t)h.oh, (uint64_t)h.bf};
      uint64_t gstr[3] = {(uint64_t)h.sx * h.ldo * 2, (uint64_t)h.sy * h.ows * h.ldo * 2, (uint64_t)h.ohs * h.ows * h.ldo * 2};
      uint8_t* base = static_cast<uint8_t*>(h.out) + (1LL * h.phase_oy[ph] * h.ows + h.phase_ox[ph]) * h.ldo * 2;
      int rc = encode_tensor_map(&d.omap[ph], dt, 4, base, gdim, gstr, box, false, 64);
      if (rc) return rc;
    }
    if (h.out2) {
      uint64_t gdim[4] = {(uint64_t)(h.n - h.nsplit), (uint64_t)h.ow, (uint64_t)h.oh, (uint64_t)h.bf};
      uint64_t gstr[3] = {(uint64_t)h.sx * h.ldo2 * 2, (uint64_t)h.sy * h.ows * h.ldo2 * 2, (uint64_t)h.ohs * h.ows * h.ldo2 * 2};
      int rc = encode_tensor_map(&d.omap2, dt, 4, h.out2, gdim, gstr, box, false, 64);
      if (rc) return rc;
    }
  }

  const size_t smem = static_cast<size_t>(d.stages) * d.stage_bytes + bres_bytes + sizeof(CgemmSmemCtl) + 2048 +
                      static_cast<size_t>(two_groups ? 16 : 8) * 2048;   // ring + resident weights + control + store staging
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(cgemm_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 214 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(cgemm_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 214 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(cgemm_kernel<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 206 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(cgemm_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 206 * 1024);
    if (e != cudaSuccess) return set_cuda_error(e, "vmm_cgemm: cudaFuncSetAttribute");
    attr_set = true;
  }
  const int grid = d.total_tiles < num_sms() ? d.total_tiles : num_sms();
  cudaError_t le;
  if (two_groups) {
    if (h.fmt == VMM_FMT_F16) le = launch_maybe_pdl(cgemm_kernel<0, 2>, dim3(grid), dim3(640), smem, stream, d);
    else le = launch_maybe_pdl(cgemm_kernel<1, 2>, dim3(grid), dim3(640), smem, stream, d);
  } else {
    if (h.fmt == VMM_FMT_F16) le = launch_maybe_pdl(cgemm_kernel<0, 1>, dim3(grid), dim3(384), smem, stream, d);
    else le = launch_maybe_pdl(cgemm_kernel<1, 1>, dim3(grid), dim3(384), smem, stream, d);
  }
  if (le != cudaSuccess) return set_cuda_error(le, "vmm_cgemm: launch");
  count_launch();
  return check_launch("vmm_cgemm");
}
