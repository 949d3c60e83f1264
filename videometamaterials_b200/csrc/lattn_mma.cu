// Linear (O(n)) spatial attention on warp-level tensor cores.  VDDP:331-378.
//
// Per frame-image bf and head h (n = h*w pixels, T cond tokens prepended, d = e = 32):
//     w[m, d]  = exp(k[m, d] - M[d]),  Z[d] = sum_m w[m, d]                (softmax over tokens + pixels)
//     ctx[d,e] = sum_m w[m, d] v[m, e] / (Z[d] * n)
//     out[n,e] = sum_d qs[n, d] ctx[d, e],   qs = softmax_d(q[n, :]) * scale
// Every product is a small GEMM whose long axis is the pixel axis, so the CTA (8 warps = 8 heads) stages a strip
// of pixel rows in shared memory, applies the exponentials in place, and each warp runs m16n8k16 MMAs on its
// head's 32 columns; the 32 x 32 matrices (ctx, d ctx) live in registers as ready-made B fragments.
#include "common.cuh"
#include "mma_sync.cuh"

namespace vmm {

constexpr int LROWS = 64;        // pixel rows staged per step
constexpr int LSR = 32;          // rows per step and buffer of the ctx / out / dctx kernels (two buffers, cp.async)
constexpr int LBR = 16;          // rows per step and buffer of lattn_bwd_mma_kernel (two buffers)
constexpr int LP3 = 776;         // pitch of a full qkv row (768 + 8)
constexpr int LP2 = 520;         // pitch of a k|v row pair (512 + 8)
constexpr int LP1 = 264;         // pitch of a 256-wide row (256 + 8)

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

struct LFrag {
  uint32_t zrow;
  int lm, lr;
  // A (16 x 16) from rows = M index, K contiguous: rows r0.., cols c..
  __device__ __forceinline__ uint32_t a(uint32_t base, int pitch, int r0, int col, int nrows) const {
    const int row = r0 + lr + 8 * (lm & 1), c = col + 8 * (lm >> 1);
    return (row < nrows) ? base + static_cast<uint32_t>(row * pitch + c) * 2 : zrow + static_cast<uint32_t>(c & 255) * 2;
  }
  // A (16 x 16) = transpose of memory rows (rows = K index, M contiguous): use with ldsm_x4_trans
  __device__ __forceinline__ uint32_t at(uint32_t base, int pitch, int r0, int col, int nrows) const {
    const int row = r0 + lr + 8 * (lm >> 1), c = col + 8 * (lm & 1);
    return (row < nrows) ? base + static_cast<uint32_t>(row * pitch + c) * 2 : zrow + static_cast<uint32_t>(c & 255) * 2;
  }
  // B (16 x 16 = two n-tiles) from rows = K index, N contiguous: use with ldsm_x4_trans
  __device__ __forceinline__ uint32_t bt(uint32_t base, int pitch, int r0, int col, int nrows) const { return a(base, pitch, r0, col, nrows); }
};

__device__ __forceinline__ uint16_t f2h(float v, int fmt) { return f_to_h16(v, fmt); }

// ------------------------------------------------------------------------------------------------
// column maxima of k over tokens + pixels:  kstat[bf][h][d][0] = M, [1] = 0
// ------------------------------------------------------------------------------------------------
__global__ void lattn_kstat_init_kernel(float* __restrict__ kstat, const float* __restrict__ ekv, int T, int HD, int frames, long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % HD);
    const int b = static_cast<int>(i / HD) / frames;
    float mx = -3.0e38f;
    for (int j = 0; j < T; ++j) mx = fmaxf(mx, ekv[(static_cast<long long>(b) * T + j) * 2 * HD + c]);
    kstat[i * 2] = mx;
    kstat[i * 2 + 1] = 0.f;
  }
}

__global__ void __launch_bounds__(256) lattn_kmax_kernel(const uint16_t* __restrict__ qkv, float* __restrict__ kstat, int fmt, int HW, int HD,
                                                         int rows_per_cta) {
  __shared__ float red[8][256];
  const int bf = blockIdx.y;
  const int vc = threadIdx.x & 31, vr = threadIdx.x >> 5;     // 32 x 16-byte vectors per 256-wide k row, 8 rows in flight
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(r0 + rows_per_cta, HW);
  float mx[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) mx[j] = -3.0e38f;
  for (int r = r0 + vr; r < r1; r += 8) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(qkv + (static_cast<long long>(bf) * HW + r) * 3 * HD + HD) + vc);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack2_h16(w[j], fmt);
      mx[2 * j] = fmaxf(mx[2 * j], f.x);
      mx[2 * j + 1] = fmaxf(mx[2 * j + 1], f.y);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[vr][vc * 8 + j] = mx[j];
  __syncthreads();
  const int c = threadIdx.x;
  if (c < HD) {
    float m = red[0][c];
#pragma unroll
    for (int k = 1; k < 8; ++k) m = fmaxf(m, red[k][c]);
    atomic_max_float(kstat + (static_cast<long long>(bf) * HD + c) * 2, m);
  }
}

// ------------------------------------------------------------------------------------------------
// ctx accumulation:  acc[bf][h][d][e] += sum_m w v,  kstat[..][1] += sum_m w        (fp32 atomics per CTA)
// ------------------------------------------------------------------------------------------------
template <int FMT>
__global__ void __launch_bounds__(256) lattn_ctx_mma_kernel(const uint16_t* __restrict__ qkv, const float* __restrict__ ekv, int T,
                                                            float* __restrict__ acc, float* __restrict__ kstat, int HW, int frames,
                                                            int rows_per_cta) {
  pdl_trigger();
  extern __shared__ __align__(16) uint16_t lsm[];
  constexpr int R = LSR;
  uint16_t* tile0 = lsm;                             // [2][R][LP2]   k -> w | v      (cp.async double buffer, see lattn_bwd_mma_kernel)
  uint16_t* zrow = tile0 + 2 * R * LP2;              // [256] zeros
  float* Ms = reinterpret_cast<float*>(zrow + 256);  // [256] column maxima
  const int HD = 256;
  const int bf = blockIdx.y, b = bf / frames;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3, h = warp;
  const int r_begin = blockIdx.x * rows_per_cta;
  const int r_end = min(r_begin + rows_per_cta, HW);
  const uint32_t tile_s0 = smem_u32(tile0);
  auto prefetch = [&](int s0, int buf) {
    const int cnt = min(R, r_end - s0);
#pragma unroll
    for (int u = 0; u < R * 64 / 256; ++u) {
      const int i = u * 256 + tid;
      const int r = i >> 6, c8 = i & 63;
      const bool ok = r < cnt;
      cp_async16(tile_s0 + static_cast<uint32_t>((buf * R + r) * LP2 + c8 * 8) * 2,
                 qkv + (static_cast<long long>(bf) * HW + s0 + (ok ? r : 0)) * 3 * HD + HD + c8 * 8, ok ? 16 : 0);
    }
    cp_async_commit();
  };
  // the first CTA of each frame-image also owns the T cond tokens (fp32 rows of ekv): staged through buffer 1 before the pixel
  // rows start to arrive in buffer 0
  if (r_begin < r_end) prefetch(r_begin, 0);
  for (int i = tid; i < 256; i += 256) {
    zrow[i] = 0;
    Ms[i] = kstat[(static_cast<long long>(bf) * HD + i) * 2];
  }
  __syncthreads();
  LFrag fa;
  fa.zrow = smem_u32(zrow);
  fa.lm = lane >> 3;
  fa.lr = lane & 7;
  float C[2][5][4];
#pragma unroll
  for (int x = 0; x < 40; ++x) (&C[0][0][0])[x] = 0.f;
  const uint32_t ones = pack2<FMT>(1.f, 1.f);
  const uint32_t bones[2] = {ones, ones};
  auto accumulate = [&](uint32_t tile_s, int cnt) {
    for (int r0 = 0; r0 < cnt; r0 += 16) {
      uint32_t vb[2][4];
      ldsm_x4_trans(vb[0], fa.bt(tile_s, LP2, r0, HD + h * 32, R));
      ldsm_x4_trans(vb[1], fa.bt(tile_s, LP2, r0, HD + h * 32 + 16, R));
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        uint32_t wa[4];
        ldsm_x4_trans(wa, fa.at(tile_s, LP2, r0, h * 32 + 16 * mt, R));
        mma16816<FMT>(C[mt][0], wa, vb[0]);
        mma16816<FMT>(C[mt][1], wa, vb[0] + 2);
        mma16816<FMT>(C[mt][2], wa, vb[1]);
        mma16816<FMT>(C[mt][3], wa, vb[1] + 2);
        mma16816<FMT>(C[mt][4], wa, bones);
      }
    }
  };
  if (blockIdx.x == 0) {
    uint16_t* tk = tile0 + R * LP2;
    for (int j0 = 0; j0 < T; j0 += R) {
      const int cnt = min(R, T - j0);
      for (int i = tid; i < R * 64; i += 256) {
        const int r = i >> 6, c8 = i & 63;
        uint4 o = make_uint4(0, 0, 0, 0);
        if (r < cnt) {
          const float* src = ekv + (static_cast<long long>(b) * T + j0 + r) * 2 * HD + c8 * 8;
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = src[j];
          if (c8 < 32) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __expf(v[j] - Ms[c8 * 8 + j]);
          }
          o = make_uint4(pack2<FMT>(v[0], v[1]), pack2<FMT>(v[2], v[3]), pack2<FMT>(v[4], v[5]), pack2<FMT>(v[6], v[7]));
        }
        *reinterpret_cast<uint4*>(tk + r * LP2 + c8 * 8) = o;
      }
      __syncthreads();
      accumulate(tile_s0 + static_cast<uint32_t>(R * LP2) * 2, cnt);
      __syncthreads();
    }
  }
  // the thread's 8 k channels are the same in every step (tid & 31 picks the 16-byte column): their maxima stay in registers
  float mreg[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) mreg[j] = Ms[(tid & 31) * 8 + j];
  int buf = 0;
  for (int s0 = r_begin; s0 < r_end; s0 += R, buf ^= 1) {
    const int cnt = min(R, r_end - s0);
    uint16_t* tile = tile0 + buf * R * LP2;
    cp_async_wait<0>();
    __syncthreads();       // rows of this step have landed; the MMAs of the previous step are done with the other buffer
    if (s0 + R < r_end) prefetch(s0 + R, buf ^ 1);
    // k -> w = exp(k - M) in place: thread -> (row, 8 channels)
#pragma unroll
    for (int u = 0; u < R * 32 / 256; ++u) {
      const int i = u * 256 + tid;
      const int r = i >> 5, c8 = i & 31;
      if (r < cnt) {
        uint4* kp = reinterpret_cast<uint4*>(tile + r * LP2 + c8 * 8);
        const uint4 v = *kp;
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack2<FMT>(w[j]);
          o[j] = pack2<FMT>(__expf(f.x - mreg[2 * j]), __expf(f.y - mreg[2 * j + 1]));
        }
        *kp = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
    __syncthreads();
    accumulate(tile_s0 + static_cast<uint32_t>(buf * R * LP2) * 2, cnt);
  }
  float* ab = acc + (static_cast<long long>(bf) * 8 + h) * 1024;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int rh = 0; rh < 2; ++rh) {
      const int d = 16 * mt + g + 8 * rh;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        atomicAdd(ab + d * 32 + 8 * nt + 2 * t, C[mt][nt][2 * rh]);
        atomicAdd(ab + d * 32 + 8 * nt + 2 * t + 1, C[mt][nt][2 * rh + 1]);
      }
      if (t == 0) atomicAdd(kstat + (static_cast<long long>(bf) * HD + h * 32 + d) * 2 + 1, C[mt][4][2 * rh]);
    }
}

__global__ void lattn_ctx_finalize_kernel(float* __restrict__ ctx, const float* __restrict__ kstat, long long n, float inv_hw) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x)
    ctx[i] = ctx[i] * inv_hw / kstat[(i >> 5) * 2 + 1];      // i >> 5 == (bf, h, d)
}

// ------------------------------------------------------------------------------------------------
// out = qs ctx   (and, with dout given, dctx += qs^T dout for the backward)
// ------------------------------------------------------------------------------------------------
template <int FMT>
__device__ __forceinline__ void softmax_rows_inplace(uint16_t* tile, int pitch, int col0, int cnt, float scale, int tid, int rows = LROWS) {
  // thread -> (row, head): softmax over the head's 32 values, result * scale written back in place.  16-byte accesses from a
  // per-head rotated start: the 8 lanes of a row sit 64 bytes apart and would otherwise share two bank groups
  for (int i = tid; i < rows * 8; i += 256) {
    const int r = i >> 3, hh = i & 7;
    if (r >= cnt) continue;
    uint4* p4 = reinterpret_cast<uint4*>(tile + r * pitch + col0 + hh * 32);
    const int rotq = hh >> 1;
    float v[32];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 q4 = p4[(j + rotq) & 3];
      const uint32_t w4[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        const float2 f = unpack2<FMT>(w4[x]);
        v[8 * j + 2 * x] = f.x;
        v[8 * j + 2 * x + 1] = f.y;
      }
    }
    float mx = v[0];
#pragma unroll
    for (int j = 1; j < 32; ++j) mx = fmaxf(mx, v[j]);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      v[j] = __expf(v[j] - mx);
      sum += v[j];
    }
    const float inv = scale / sum;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      p4[(j + rotq) & 3] = make_uint4(pack2<FMT>(v[8 * j] * inv, v[8 * j + 1] * inv), pack2<FMT>(v[8 * j + 2] * inv, v[8 * j + 3] * inv),
                                      pack2<FMT>(v[8 * j + 4] * inv, v[8 * j + 5] * inv), pack2<FMT>(v[8 * j + 6] * inv, v[8 * j + 7] * inv));
  }
}

template <int FMT>
__global__ void __launch_bounds__(256) lattn_out_mma_kernel(const uint16_t* __restrict__ qkv, const float* __restrict__ ctx,
                                                            uint16_t* __restrict__ out, int HW, float scale, int rows_per_cta) {
  pdl_trigger();
  extern __shared__ __align__(16) uint16_t lsm[];
  constexpr int R = LSR;
  const float hw = static_cast<float>(HW);
  uint16_t* tile0 = lsm;                     // [2][R][LP1]  q -> qs -> out rows   (cp.async double buffer)
  uint16_t* zrow = tile0 + 2 * R * LP1;
  const int HD = 256;
  const int bf = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3, h = warp;
  const int r_begin = blockIdx.x * rows_per_cta, r_end = min(r_begin + rows_per_cta, HW);
  const uint32_t tile_s0 = smem_u32(tile0);
  auto prefetch = [&](int s0, int buf) {
    const int cnt = min(R, r_end - s0);
#pragma unroll
    for (int u = 0; u < R * 32 / 256; ++u) {
      const int i = u * 256 + tid;
      const int r = i >> 5, c8 = i & 31;
      const bool ok = r < cnt;
      cp_async16(tile_s0 + static_cast<uint32_t>((buf * R + r) * LP1 + c8 * 8) * 2,
                 qkv + (static_cast<long long>(bf) * HW + s0 + (ok ? r : 0)) * 3 * HD + c8 * 8, ok ? 16 : 0);
    }
    cp_async_commit();
  };
  if (r_begin < r_end) prefetch(r_begin, 0);
  for (int i = tid; i < 256; i += 256) zrow[i] = 0;
  // B[k = d][n = e] = ctx[d][e]
  uint32_t cb[2][4][2];
  const float* ch = ctx + (static_cast<long long>(bf) * 8 + h) * 1024;
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int d = 16 * ks + 8 * hf + 2 * t, e = 8 * nt + g;
        // ctx carries the 1/(h*w) of VDDP:371 and is ~1e-4: below fp16's normal range, so the 16-bit fragment holds
        // ctx * (h*w) and the fp32 accumulator is scaled back
        cb[ks][nt][hf] = pack2<FMT>(ch[d * 32 + e] * hw, ch[(d + 1) * 32 + e] * hw);
      }
  const float inv_hw = 1.f / hw;
  LFrag fa;
  fa.zrow = smem_u32(zrow);
  fa.lm = lane >> 3;
  fa.lr = lane & 7;
  int buf = 0;
  for (int s0 = r_begin; s0 < r_end; s0 += R, buf ^= 1) {
    const int cnt = min(R, r_end - s0);
    uint16_t* tile = tile0 + buf * R * LP1;
    const uint32_t tile_s = tile_s0 + static_cast<uint32_t>(buf * R * LP1) * 2;
    cp_async_wait<0>();
    __syncthreads();       // rows landed; the copy-out of the previous step is done with the other buffer
    if (s0 + R < r_end) prefetch(s0 + R, buf ^ 1);
    softmax_rows_inplace<FMT>(tile, LP1, 0, cnt, scale, tid, R);
    __syncthreads();
    for (int r0 = 0; r0 < cnt; r0 += 16) {
      float O[4][4];
#pragma unroll
      for (int x = 0; x < 16; ++x) (&O[0][0])[x] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        uint32_t qa[4];
        ldsm_x4(qa, fa.a(tile_s, LP1, r0, h * 32 + 16 * ks, R));
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma16816<FMT>(O[nt], qa, cb[ks][nt]);
      }
      // in place over this warp's own 32 columns of the 16 rows it has just consumed; whole rows leave with 16-byte stores below
      __syncwarp();
#pragma unroll
      for (int rh = 0; rh < 2; ++rh) {
        const int r = r0 + g + 8 * rh;
        if (r < cnt) {
          uint16_t* orow = tile + r * LP1 + h * 32 + 2 * t;
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
            *reinterpret_cast<uint32_t*>(orow + 8 * nt) = pack2<FMT>(O[nt][2 * rh] * inv_hw, O[nt][2 * rh + 1] * inv_hw);
        }
      }
    }
    __syncthreads();
    for (int i = tid; i < cnt * 32; i += 256) {
      const int r = i >> 5, c8 = i & 31;
      *reinterpret_cast<uint4*>(out + (static_cast<long long>(bf) * HW + s0 + r) * HD + c8 * 8) = *reinterpret_cast<const uint4*>(tile + r * LP1 + c8 * 8);
    }
  }
}

template <int FMT>
__global__ void __launch_bounds__(256) lattn_dctx_mma_kernel(const uint16_t* __restrict__ qkv, const uint16_t* __restrict__ dout,
                                                             float* __restrict__ dctx, int HW, float scale, int rows_per_cta) {
  pdl_trigger();
  extern __shared__ __align__(16) uint16_t lsm[];
  constexpr int R = LSR;
  uint16_t* qt0 = lsm;                       // [2][R][LP1] q -> qs        (cp.async double buffer)
  uint16_t* dt0 = qt0 + 2 * R * LP1;         // [2][R][LP1] dout
  uint16_t* zrow = dt0 + 2 * R * LP1;
  const int HD = 256;
  const int bf = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3, h = warp;
  const int r_begin = blockIdx.x * rows_per_cta, r_end = min(r_begin + rows_per_cta, HW);
  const uint32_t qt_s0 = smem_u32(qt0), dt_s0 = smem_u32(dt0);
  auto prefetch = [&](int s0, int buf) {
    const int cnt = min(R, r_end - s0);
#pragma unroll
    for (int u = 0; u < R * 64 / 256; ++u) {
      const int i = u * 256 + tid;
      const int r = i >> 6, c8 = i & 63;
      const bool ok = r < cnt;
      const long long row = static_cast<long long>(bf) * HW + s0 + (ok ? r : 0);
      if (c8 < 32) cp_async16(qt_s0 + static_cast<uint32_t>((buf * R + r) * LP1 + c8 * 8) * 2, qkv + row * 3 * HD + c8 * 8, ok ? 16 : 0);
      else cp_async16(dt_s0 + static_cast<uint32_t>((buf * R + r) * LP1 + (c8 - 32) * 8) * 2, dout + row * HD + (c8 - 32) * 8, ok ? 16 : 0);
    }
    cp_async_commit();
  };
  if (r_begin < r_end) prefetch(r_begin, 0);
  for (int i = tid; i < 256; i += 256) zrow[i] = 0;
  LFrag fa;
  fa.zrow = smem_u32(zrow);
  fa.lm = lane >> 3;
  fa.lr = lane & 7;
  float C[2][4][4];
#pragma unroll
  for (int x = 0; x < 32; ++x) (&C[0][0][0])[x] = 0.f;
  int buf = 0;
  for (int s0 = r_begin; s0 < r_end; s0 += R, buf ^= 1) {
    const int cnt = min(R, r_end - s0);
    const uint32_t qt_s = qt_s0 + static_cast<uint32_t>(buf * R * LP1) * 2, dt_s = dt_s0 + static_cast<uint32_t>(buf * R * LP1) * 2;
    cp_async_wait<0>();
    __syncthreads();
    if (s0 + R < r_end) prefetch(s0 + R, buf ^ 1);
    softmax_rows_inplace<FMT>(qt0 + buf * R * LP1, LP1, 0, cnt, scale, tid, R);
    __syncthreads();
    for (int r0 = 0; r0 < cnt; r0 += 16) {
      uint32_t db[2][4];
      ldsm_x4_trans(db[0], fa.bt(dt_s, LP1, r0, h * 32, R));
      ldsm_x4_trans(db[1], fa.bt(dt_s, LP1, r0, h * 32 + 16, R));
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        uint32_t qa[4];
        ldsm_x4_trans(qa, fa.at(qt_s, LP1, r0, h * 32 + 16 * mt, R));
        mma16816<FMT>(C[mt][0], qa, db[0]);
        mma16816<FMT>(C[mt][1], qa, db[0] + 2);
        mma16816<FMT>(C[mt][2], qa, db[1]);
        mma16816<FMT>(C[mt][3], qa, db[1] + 2);
      }
    }
  }
  float* ab = dctx + (static_cast<long long>(bf) * 8 + h) * 1024;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int rh = 0; rh < 2; ++rh) {
      const int d = 16 * mt + g + 8 * rh;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        atomicAdd(ab + d * 32 + 8 * nt + 2 * t, C[mt][nt][2 * rh]);
        atomicAdd(ab + d * 32 + 8 * nt + 2 * t + 1, C[mt][nt][2 * rh + 1]);
      }
    }
}

// ------------------------------------------------------------------------------------------------
// backward apply: dq, dk, dv per pixel row from ctx, G = dctx * vscale, the k statistics and c[d] = sum_e dctx ctx
// ------------------------------------------------------------------------------------------------
template <int FMT>
__global__ void __launch_bounds__(256) lattn_bwd_mma_kernel(const uint16_t* __restrict__ qkv, const uint16_t* __restrict__ dout,
                                                            const float* __restrict__ ctx, const float* __restrict__ dctx,
                                                            const float* __restrict__ kstat, uint16_t* __restrict__ dqkv, int HW,
                                                            float scale, float vscale, int rows_per_cta) {
  pdl_trigger();
  extern __shared__ __align__(16) uint16_t lsm[];
  // Two buffers of R = 16 rows: cp.async fills buffer (s + 1) & 1 with the rows of the next step while the CTA turns the rows of
  // step s into p / wn, runs the MMAs and copies the gradients out.  (One buffer of 32 rows and register-staged loads left every
  // phase of a step waiting on HBM: 989 us for the level-0 block, 2.9 TB/s.)
  constexpr int R = LBR;
  uint16_t* tile0 = lsm;                     // [2][R][LP3]  q -> p | k -> wn | v    (then dq | dk | dv in place)
  uint16_t* dt0 = tile0 + 2 * R * LP3;       // [2][R][LP1]  dout
  uint16_t* zrow = dt0 + 2 * R * LP1;        // [256]
  float* Ms = reinterpret_cast<float*>(zrow + 256);   // [256] max
  float* Zi = Ms + 256;                                // [256] 1 / Z
  const int HD = 256;
  const int bf = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3, h = warp;
  const int r_begin = blockIdx.x * rows_per_cta, r_end = min(r_begin + rows_per_cta, HW);
  const uint32_t tile_s0 = smem_u32(tile0), dt_s0 = smem_u32(dt0);
  auto prefetch = [&](int s0, int buf) {
    const int cnt = min(R, r_end - s0);
#pragma unroll
    for (int u = 0; u < R * 128 / 256; ++u) {
      const int i = u * 256 + tid;
      const int r = i >> 7, c8 = i & 127;
      const bool ok = r < cnt;
      const long long row = static_cast<long long>(bf) * HW + s0 + (ok ? r : 0);
      if (c8 < 96) cp_async16(tile_s0 + static_cast<uint32_t>((buf * R + r) * LP3 + c8 * 8) * 2, qkv + row * 3 * HD + c8 * 8, ok ? 16 : 0);
      else cp_async16(dt_s0 + static_cast<uint32_t>((buf * R + r) * LP1 + (c8 - 96) * 8) * 2, dout + row * HD + (c8 - 96) * 8, ok ? 16 : 0);
    }
    cp_async_commit();
  };
  if (r_begin < r_end) prefetch(r_begin, 0);
  for (int i = tid; i < 256; i += 256) {
    zrow[i] = 0;
    // stored as [channel half (0..3 | 4..7)][16-byte column (32)][4]: the thread that owns column c8 reads two consecutive-per-lane
    // float4 of each table (the plain [256] layout put 8 lanes on every bank: 16 conflicting 4-byte loads per vector)
    const int pi = (((i & 7) >> 2) * 32 + (i >> 3)) * 4 + (i & 3);
    Ms[pi] = kstat[(static_cast<long long>(bf) * HD + i) * 2];
    Zi[pi] = 1.f / kstat[(static_cast<long long>(bf) * HD + i) * 2 + 1];
  }
  const float* ch = ctx + (static_cast<long long>(bf) * 8 + h) * 1024;
  const float* gh = dctx + (static_cast<long long>(bf) * 8 + h) * 1024;
  const float ivs = 1.f / vscale;
  uint32_t bc[2][4][2];    // B[k = e][n = d] = ctx[d][e]            (dqs = dout ctx^T)
  uint32_t bg1[2][4][2];   // B[k = d][n = e] = G[d][e]              (dv  = wn G)
  uint32_t bg2[2][4][2];   // B[k = e][n = d] = G[d][e]              (dwn = v G^T)
  float cc[4][2];          // c[d] for d = 8 nt + 2 t + {0, 1}
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int k0 = 16 * ks + 8 * hf + 2 * t, n = 8 * nt + g;
        // 16-bit fragments hold ctx / vscale and dctx (not dctx * vscale): see lattn_out_mma_kernel
        bc[ks][nt][hf] = pack2<FMT>(ch[n * 32 + k0] * ivs, ch[n * 32 + k0 + 1] * ivs);
        bg2[ks][nt][hf] = pack2<FMT>(gh[n * 32 + k0], gh[n * 32 + k0 + 1]);
        bg1[ks][nt][hf] = pack2<FMT>(gh[k0 * 32 + n], gh[(k0 + 1) * 32 + n]);
      }
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int d = 8 * nt + 2 * t + c;
      float a = 0.f;
      for (int e = 0; e < 32; ++e) a += gh[d * 32 + e] * ch[d * 32 + e];
      cc[nt][c] = a;
    }
  LFrag fa;
  fa.zrow = smem_u32(zrow);
  fa.lm = lane >> 3;
  fa.lr = lane & 7;
  int buf = 0;
  for (int s0 = r_begin; s0 < r_end; s0 += R, buf ^= 1) {
    const int cnt = min(R, r_end - s0);
    uint16_t* tile = tile0 + buf * R * LP3;
    const uint32_t tile_s = tile_s0 + static_cast<uint32_t>(buf * R * LP3) * 2, dt_s = dt_s0 + static_cast<uint32_t>(buf * R * LP1) * 2;
    cp_async_wait<0>();
    __syncthreads();       // rows of this step have landed; every thread is done with the other buffer (copy-out of the previous step)
    if (s0 + R < r_end) prefetch(s0 + R, buf ^ 1);
    if (tid < 128) {
      // q -> p = softmax_d(q) (scale applied at the end): thread -> (row, head)
      const int r = tid >> 3, hh = tid & 7;
      if (r < cnt) {
        uint4* p4 = reinterpret_cast<uint4*>(tile + r * LP3 + hh * 32);
        const int rotq = hh >> 1;        // lanes hh = 0..7 of a row sit 64 bytes apart: walking the four 16-byte pieces from a rotated start
        float v[32];                     // keeps the 8 lanes of a shared-memory phase on distinct bank groups
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 q4 = p4[(j + rotq) & 3];
          const uint32_t w4[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            const float2 f = unpack2<FMT>(w4[x]);
            v[8 * j + 2 * x] = f.x;
            v[8 * j + 2 * x + 1] = f.y;
          }
        }
        float mx = v[0];
#pragma unroll
        for (int j = 1; j < 32; ++j) mx = fmaxf(mx, v[j]);
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v[j] = __expf(v[j] - mx);
          sum += v[j];
        }
        const float inv = 1.f / sum;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          p4[(j + rotq) & 3] = make_uint4(pack2<FMT>(v[8 * j] * inv, v[8 * j + 1] * inv), pack2<FMT>(v[8 * j + 2] * inv, v[8 * j + 3] * inv),
                             pack2<FMT>(v[8 * j + 4] * inv, v[8 * j + 5] * inv), pack2<FMT>(v[8 * j + 6] * inv, v[8 * j + 7] * inv));
      }
    } else {
      // k -> wn = exp(k - M) / Z: thread -> (row, 8 channels), 16 rows x 32 vectors over 128 threads
#pragma unroll
      for (int u = 0; u < R * 32 / 128; ++u) {
        const int i = u * 128 + (tid - 128);
        const int r = i >> 5, c8 = i & 31;
        if (r < cnt) {
          uint4* kp = reinterpret_cast<uint4*>(tile + r * LP3 + HD + c8 * 8);
          const uint4 v = *kp;
          const float4 m0 = reinterpret_cast<const float4*>(Ms)[c8], m1 = reinterpret_cast<const float4*>(Ms)[32 + c8];
          const float4 z0 = reinterpret_cast<const float4*>(Zi)[c8], z1 = reinterpret_cast<const float4*>(Zi)[32 + c8];
          const float2 f0 = unpack2<FMT>(v.x), f1 = unpack2<FMT>(v.y), f2 = unpack2<FMT>(v.z), f3 = unpack2<FMT>(v.w);
          *kp = make_uint4(pack2<FMT>(__expf(f0.x - m0.x) * z0.x, __expf(f0.y - m0.y) * z0.y),
                           pack2<FMT>(__expf(f1.x - m0.z) * z0.z, __expf(f1.y - m0.w) * z0.w),
                           pack2<FMT>(__expf(f2.x - m1.x) * z1.x, __expf(f2.y - m1.y) * z1.y),
                           pack2<FMT>(__expf(f3.x - m1.z) * z1.z, __expf(f3.y - m1.w) * z1.w));
        }
      }
    }
    __syncthreads();
    {
      constexpr int r0 = 0;
      float DQ[4][4], DV[4][4], DW[4][4];
#pragma unroll
      for (int x = 0; x < 16; ++x) (&DQ[0][0])[x] = (&DV[0][0])[x] = (&DW[0][0])[x] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        uint32_t da[4], wa[4], va[4];
        ldsm_x4(da, fa.a(dt_s, LP1, r0, h * 32 + 16 * ks, R));
        ldsm_x4(wa, fa.a(tile_s, LP3, r0, HD + h * 32 + 16 * ks, R));
        ldsm_x4(va, fa.a(tile_s, LP3, r0, 2 * HD + h * 32 + 16 * ks, R));
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          mma16816<FMT>(DQ[nt], da, bc[ks][nt]);
          mma16816<FMT>(DV[nt], wa, bg1[ks][nt]);
          mma16816<FMT>(DW[nt], va, bg2[ks][nt]);
        }
      }
#pragma unroll
      for (int rh = 0; rh < 2; ++rh) {
        const int r = r0 + g + 8 * rh;
        const bool live = r < cnt;
        const uint16_t* trow = tile + (live ? r : 0) * LP3 + h * 32 + 2 * t;
        float pv[4][2], wv[4][2];
        float dot = 0.f;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const float2 p2 = unpack2<FMT>(*reinterpret_cast<const uint32_t*>(trow + 8 * nt));
          const float2 w2 = unpack2<FMT>(*reinterpret_cast<const uint32_t*>(trow + HD + 8 * nt));
          pv[nt][0] = p2.x;
          pv[nt][1] = p2.y;
          wv[nt][0] = w2.x;
          wv[nt][1] = w2.y;
          DQ[nt][2 * rh] *= vscale;
          DQ[nt][2 * rh + 1] *= vscale;
          DW[nt][2 * rh] *= vscale;
          DW[nt][2 * rh + 1] *= vscale;
          DV[nt][2 * rh] *= vscale;
          DV[nt][2 * rh + 1] *= vscale;
          dot += p2.x * DQ[nt][2 * rh] + p2.y * DQ[nt][2 * rh + 1];
        }
        dot += __shfl_xor_sync(0xffffffffu, dot, 1);
        dot += __shfl_xor_sync(0xffffffffu, dot, 2);
        if (live) {
          // in place over the staged row: the p / wn values of these very elements were read just above, the MMAs of this 16-row
          // block are done, and the 32 columns of head h in each segment belong to this warp alone.  The CTA then copies whole
          // rows to dqkv with 16-byte stores (4-byte stores from the fragments split into eight pieces per warp instruction).
          uint16_t* orow = tile + r * LP3 + h * 32 + 2 * t;
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            *reinterpret_cast<uint32_t*>(orow + 8 * nt) =
                pack2<FMT>(pv[nt][0] * scale * (DQ[nt][2 * rh] - dot), pv[nt][1] * scale * (DQ[nt][2 * rh + 1] - dot));
            *reinterpret_cast<uint32_t*>(orow + HD + 8 * nt) =
                pack2<FMT>(wv[nt][0] * (DW[nt][2 * rh] - cc[nt][0]), wv[nt][1] * (DW[nt][2 * rh + 1] - cc[nt][1]));
            *reinterpret_cast<uint32_t*>(orow + 2 * HD + 8 * nt) = pack2<FMT>(DV[nt][2 * rh], DV[nt][2 * rh + 1]);
          }
        }
      }
    }
    __syncthreads();
    for (int i = tid; i < cnt * 96; i += 256) {           // coalesced copy-out: 96 x 16 bytes per 768-wide row
      const int r = i / 96, c8 = i - r * 96;
      *reinterpret_cast<uint4*>(dqkv + (static_cast<long long>(bf) * HW + s0 + r) * 3 * HD + c8 * 8) =
          *reinterpret_cast<const uint4*>(tile + r * LP3 + c8 * 8);
    }
  }
}

// cond-token gradients (tiny).  CTA = (frame-image, head): the 32 x 32 ctx / dctx matrices go to shared memory with coalesced
// loads (rows padded to 33 floats: lane = d walks a row without bank conflicts), then one warp per token:
//   lane = d:  dk[d] = w[d] * sum_e dctx[d][e] (vscale v[e] - ctx[d][e]),   lane = e:  dv[e] = vscale * sum_d w[d] dctx[d][e]
// summed over the frames of a sample with atomics.  (Round 1 ran one warp per (frame-image, head, token) straight from global
// memory: 68 us per launch at any size, 0.5 ms per step.)
__global__ void __launch_bounds__(256) lattn_bwd_tokens_kernel(const float* __restrict__ ekv, int T, const float* __restrict__ ctx,
                                                               const float* __restrict__ dctx, const float* __restrict__ kstat,
                                                               float* __restrict__ dekv, int BF, int frames, float vscale) {
  pdl_trigger();
  __shared__ float Cs[32][33], Gs[32][33];
  const int HD = 256;
  const int h = blockIdx.x & 7, bf = blockIdx.x >> 3;
  const int b = bf / frames;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* ch = ctx + (static_cast<long long>(bf) * 8 + h) * 1024;
  const float* gh = dctx + (static_cast<long long>(bf) * 8 + h) * 1024;
  for (int i = threadIdx.x; i < 1024; i += 256) {
    Cs[i >> 5][i & 31] = __ldg(ch + i);
    Gs[i >> 5][i & 31] = __ldg(gh + i);
  }
  const float* ks = kstat + (static_cast<long long>(bf) * HD + h * 32) * 2;
  const float mx = ks[lane * 2], zi = 1.f / ks[lane * 2 + 1];
  __syncthreads();
  for (int j = warp; j < T; j += 8) {
    const float* src = ekv + (static_cast<long long>(b) * T + j) * 2 * HD + h * 32;
    float* dst = dekv + (static_cast<long long>(b) * T + j) * 2 * HD + h * 32;
    const float w = __expf(src[lane] - mx) * zi;
    const float vl = vscale * src[HD + lane];            // lane = e
    float r = 0.f, dv = 0.f;
#pragma unroll
    for (int e = 0; e < 32; ++e) r = fmaf(Gs[lane][e], __shfl_sync(0xffffffffu, vl, e) - Cs[lane][e], r);
#pragma unroll
    for (int d = 0; d < 32; ++d) dv = fmaf(__shfl_sync(0xffffffffu, w, d), Gs[d][lane], dv);
    atomicAdd(dst + lane, w * r);
    atomicAdd(dst + HD + lane, dv * vscale);
  }
}

static int lat_rows_per_cta(int HW, int BF) {
  int chunks = (3 * num_sms() + BF - 1) / BF;
  if (chunks < 1) chunks = 1;
  int rows = (HW + chunks - 1) / chunks;
  rows = (rows + LROWS - 1) / LROWS * LROWS;
  return rows;
}

// Rows per CTA for a kernel with `slots` co-resident CTAs on the device: the kernel's time is (waves of CTAs) x (rows per CTA + a
// prologue worth ~48 rows: fragments, statistics), so pick the chunk count that minimises it instead of a fixed "3 CTAs per SM"
// (level 0, b = 8: 6 chunks x 88 frame-images = 528 CTAs = 1.8 waves of 296 paid as 2; 10 chunks = 2.97 waves, 9 % less time).
static int lat_rows_waves(int HW, int BF, int slots) {
  long long best_cost = -1;
  int best_rows = HW;
  const int max_chunks = (HW + LROWS - 1) / LROWS;
  for (int c = 1; c <= max_chunks && c <= 96; ++c) {
    int rows = (HW + c - 1) / c;
    rows = (rows + LROWS - 1) / LROWS * LROWS;
    const int chunks = (HW + rows - 1) / rows;
    const long long waves = (static_cast<long long>(chunks) * BF + slots - 1) / slots;
    const long long cost = waves * (rows + 48);
    if (best_cost < 0 || cost < best_cost) best_cost = cost, best_rows = rows;
  }
  return best_rows;
}

template <typename K>
static int lat_slots(K kern, size_t smem, int& cache) {
  if (cache == 0) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, 256, smem) != cudaSuccess || n < 1) n = 1;
    cache = n * num_sms();
  }
  return cache;
}

}  // namespace vmm

using namespace vmm;

#define LAT_SET_ATTR(K, BYTES)                                                                         \
  {                                                                                                    \
    cudaError_t e = cudaFuncSetAttribute(K<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, BYTES);    \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(K<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, BYTES); \
    if (e != cudaSuccess) return set_cuda_error(e, "lattn: cudaFuncSetAttribute");                     \
  }

extern "C" int vmm_lattn_fwd(const void* qkv, const float* ekv, int T, void* out, float* ctx, float* kstat, int fmt, int BF, int frames,
                             int HW, int heads, float scale, void* stream_) {
  if (!qkv || !ekv || !out || !ctx || !kstat) return set_error(VMM_ERR_ARG, "vmm_lattn_fwd: null pointer (ctx and kstat are required)");
  if (heads != 8) return set_error(VMM_ERR_UNSUPPORTED, "vmm_lattn_fwd: heads must be 8");
  if (T > 48) return set_error(VMM_ERR_UNSUPPORTED, "vmm_lattn_fwd: too many cond tokens");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const size_t sm_ctx = (static_cast<size_t>(2 * LSR) * LP2 + 256) * 2 + 256 * 4;
  const size_t sm_out = (static_cast<size_t>(2 * LSR) * LP1 + 256) * 2;
  static bool attr = false;
  if (!attr) {
    LAT_SET_ATTR(lattn_ctx_mma_kernel, 100 * 1024);
    LAT_SET_ATTR(lattn_out_mma_kernel, 100 * 1024);
    attr = true;
  }
  const long long nstat = static_cast<long long>(BF) * 256;
  cudaError_t e = cudaMemsetAsync(ctx, 0, static_cast<size_t>(BF) * 8 * 1024 * sizeof(float), stream);
  if (e != cudaSuccess) return set_cuda_error(e, "vmm_lattn_fwd: memset");
  lattn_kstat_init_kernel<<<static_cast<int>((nstat + 255) / 256), 256, 0, stream>>>(kstat, ekv, T, 256, frames, nstat);
  count_launch();
  int rpc = lat_rows_per_cta(HW, BF);
  int chunks = (HW + rpc - 1) / rpc;
  lattn_kmax_kernel<<<dim3(chunks, BF), 256, 0, stream>>>(static_cast<const uint16_t*>(qkv), kstat, fmt, HW, 256, rpc);
  count_launch();
  static int slots_ctx = 0, slots_out = 0;
  rpc = lat_rows_waves(HW, BF, lat_slots(lattn_ctx_mma_kernel<1>, sm_ctx, slots_ctx));
  chunks = (HW + rpc - 1) / rpc;
  if (fmt == VMM_FMT_F16)
    lattn_ctx_mma_kernel<0><<<dim3(chunks, BF), 256, sm_ctx, stream>>>(static_cast<const uint16_t*>(qkv), ekv, T, ctx, kstat, HW, frames, rpc);
  else
    lattn_ctx_mma_kernel<1><<<dim3(chunks, BF), 256, sm_ctx, stream>>>(static_cast<const uint16_t*>(qkv), ekv, T, ctx, kstat, HW, frames, rpc);
  count_launch();
  const long long nctx = static_cast<long long>(BF) * 8 * 1024;
  lattn_ctx_finalize_kernel<<<static_cast<int>((nctx + 255) / 256), 256, 0, stream>>>(ctx, kstat, nctx, 1.f / static_cast<float>(HW));
  count_launch();
  rpc = lat_rows_waves(HW, BF, lat_slots(lattn_out_mma_kernel<1>, sm_out, slots_out));
  chunks = (HW + rpc - 1) / rpc;
  if (fmt == VMM_FMT_F16)
    lattn_out_mma_kernel<0><<<dim3(chunks, BF), 256, sm_out, stream>>>(static_cast<const uint16_t*>(qkv), ctx, static_cast<uint16_t*>(out), HW, scale, rpc);
  else
    lattn_out_mma_kernel<1><<<dim3(chunks, BF), 256, sm_out, stream>>>(static_cast<const uint16_t*>(qkv), ctx, static_cast<uint16_t*>(out), HW, scale, rpc);
  count_launch();
  return check_launch("vmm_lattn_fwd");
}

extern "C" int vmm_lattn_bwd(const void* qkv, const float* ekv, int T, const void* dout, const float* ctx, const float* kstat, float* dctx,
                             void* dqkv, float* dekv, int fmt, int BF, int frames, int HW, int heads, float scale, float vscale,
                             void* stream_) {
  if (!qkv || !ekv || !dout || !ctx || !kstat || !dctx || !dqkv || !dekv) return set_error(VMM_ERR_ARG, "vmm_lattn_bwd: null pointer");
  if (heads != 8) return set_error(VMM_ERR_UNSUPPORTED, "vmm_lattn_bwd: heads must be 8");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const size_t sm_dctx = (static_cast<size_t>(4 * LSR) * LP1 + 256) * 2;
  const size_t sm_bwd = (static_cast<size_t>(2 * LBR) * (LP3 + LP1) + 256) * 2 + 512 * 4;
  static bool attr = false;
  if (!attr) {
    LAT_SET_ATTR(lattn_dctx_mma_kernel, 100 * 1024);
    LAT_SET_ATTR(lattn_bwd_mma_kernel, 100 * 1024);
    attr = true;
  }
  cudaError_t e = cudaMemsetAsync(dctx, 0, static_cast<size_t>(BF) * 8 * 1024 * sizeof(float), stream);
  if (e != cudaSuccess) return set_cuda_error(e, "vmm_lattn_bwd: memset");
  static int slots_dctx = 0, slots_bwd = 0;
  const int rpc_d = lat_rows_waves(HW, BF, lat_slots(lattn_dctx_mma_kernel<1>, sm_dctx, slots_dctx));
  const int rpc_b = lat_rows_waves(HW, BF, lat_slots(lattn_bwd_mma_kernel<1>, sm_bwd, slots_bwd));
  const dim3 grid_d((HW + rpc_d - 1) / rpc_d, BF), grid_b((HW + rpc_b - 1) / rpc_b, BF);
  if (fmt == VMM_FMT_F16) {
    lattn_dctx_mma_kernel<0><<<grid_d, 256, sm_dctx, stream>>>(static_cast<const uint16_t*>(qkv), static_cast<const uint16_t*>(dout), dctx, HW, scale, rpc_d);
    count_launch();
    lattn_bwd_mma_kernel<0><<<grid_b, 256, sm_bwd, stream>>>(static_cast<const uint16_t*>(qkv), static_cast<const uint16_t*>(dout), ctx, dctx, kstat,
                                                             static_cast<uint16_t*>(dqkv), HW, scale, vscale, rpc_b);
  } else {
    lattn_dctx_mma_kernel<1><<<grid_d, 256, sm_dctx, stream>>>(static_cast<const uint16_t*>(qkv), static_cast<const uint16_t*>(dout), dctx, HW, scale, rpc_d);
    count_launch();
    lattn_bwd_mma_kernel<1><<<grid_b, 256, sm_bwd, stream>>>(static_cast<const uint16_t*>(qkv), static_cast<const uint16_t*>(dout), ctx, dctx, kstat,
                                                             static_cast<uint16_t*>(dqkv), HW, scale, vscale, rpc_b);
  }
  count_launch();
  lattn_bwd_tokens_kernel<<<BF * 8, 256, 0, stream>>>(ekv, T, ctx, dctx, kstat, dekv, BF, frames, vscale);
  count_launch();
  return check_launch("vmm_lattn_bwd");
}
