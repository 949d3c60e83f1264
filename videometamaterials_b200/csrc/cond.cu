// vmm_cond_fwd / vmm_cond_bwd: the conditioning / time path of Unet3D.forward (VDDP:139-151, 637-661, 745-788 and the
// per-block consumers VDDP:290-293, 304-306 [ResnetBlock.mlp], 349-353 / 457-474 [to_k / to_v on the conditioning tokens]).
//
//   t     = time_mlp(SinusoidalPosEmb(time))                       (b, td)        td = 4 dim
//   tok   = sign_emb(cond[..., None])                              (b, T, td)     rank one in cond: tok[b, j] = cond[b, j] wse + bse
//   hid   = cond_token_to_hidden(mean_j tok)                       (b, td)        LayerNorm -> Linear -> SiLU -> Linear
//   null  : tok <- null_text_token, hid <- null_text_hidden where the mask is set (classifier-free guidance drop, VDDP:772-784)
//   t     = t + hid ; every ResnetBlock: scale|shift = Linear(SiLU(t))            (b, 2 C_block)
//   every attention block: ek|ev = to_k|to_v(tok)                  (b, T, 2 * 256), keys of the temporal blocks rotated by token index
//   position bias [heads][f][f] from the relative-position table, rotary cos/sin tables [2][f][16][2] (table 0 x 32^-1/2)
//
// ~0.004 GFLOP, but ~150 ATen / cuBLAS launches per step in torch.  Here: ONE forward kernel and TWO backward kernels, fp32 on the
// CUDA cores, reading the fp32 parameter arena and writing the gradient arena in place (no gather / scatter of parameters).
//   forward : every CTA recomputes the small trunk (all samples) in shared memory, then owns a chunk of 256 output columns of
//             either the concatenated block MLPs or one attention block's keys / values.  Because tok is rank one in cond,
//             to_k(tok[b, j]) = cond[b, j] (Wk wse) + (Wk bse), and Wk null_token[j] for dropped samples: 2 + T matrix-vector
//             products per block whatever the batch.
//   backward: kernel 1 mirrors the forward chunks (thread = input index k, loop over the chunk's columns: weight rows are read
//             coalesced, the weight-gradient row is written coalesced, the input gradient accumulates in registers);
//             kernel 2 (one CTA) back-propagates through the trunk once the input gradients of all chunks are complete.
#include <algorithm>

#include "common.cuh"

namespace vmm {

constexpr int CMAXB = 32;      // samples handled by the register-blocked loops
constexpr int CMAXT = 16;      // conditioning tokens (11)
constexpr int CMAXBLK = 24;    // ResnetBlocks with a time MLP (18) / attention blocks (17)

struct CondDev {
  int B, T, dim, td, heads, frames, n_res, n_attn;
  const long long* time;
  const float* cond;
  const unsigned char* null_mask;
  const float* param;       // fp32 parameter arena
  float* grad;              // fp32 gradient arena (backward)
  const float* freqs;       // rotary frequencies [16] (frozen parameter, not in the arena)
  const int* buckets;       // relative-position bucket of (i, j): [frames][frames]
  long long o_w1, o_b1, o_w2, o_b2, o_wse, o_bse, o_lng, o_lnb, o_w3, o_b3, o_w4, o_b4, o_ntok, o_nhid, o_table;
  long long res_w[CMAXBLK], res_b[CMAXBLK];
  int res_c2[CMAXBLK];      // 2 * C_block
  int res_col0[CMAXBLK + 1];   // prefix sums of res_c2: global column index of each block
  long long res_out[CMAXBLK];  // element offset of the block's (b, 2C) output in `out`
  long long att_wk[CMAXBLK], att_wv[CMAXBLK], att_out[CMAXBLK];
  int att_temporal[CMAXBLK];
  long long bias_out, rot_out;
  float* out;               // forward outputs / backward: the gradients w.r.t. them, same layout
  float* ws;                // saved trunk intermediates + backward scratch, see offsets below
  int mlp_chunks, kv_chunks;
};

// workspace layout (floats): per-sample rows of td (or dim) values
__host__ __device__ inline long long ws_e(const CondDev& p) { return 0; }                                           // [B][dim]  sinusoidal embedding
__host__ __device__ inline long long ws_pre1(const CondDev& p) { return ws_e(p) + 1LL * p.B * p.dim; }              // [B][td]   before GELU
__host__ __device__ inline long long ws_hm(const CondDev& p) { return ws_pre1(p) + 1LL * p.B * p.td; }              // [B][td]   mean token
__host__ __device__ inline long long ws_stat(const CondDev& p) { return ws_hm(p) + 1LL * p.B * p.td; }              // [B][2]    LayerNorm mean, rstd
__host__ __device__ inline long long ws_g1(const CondDev& p) { return ws_stat(p) + 2LL * p.B; }                     // [B][td]   before SiLU
__host__ __device__ inline long long ws_tfin(const CondDev& p) { return ws_g1(p) + 1LL * p.B * p.td; }              // [B][td]   t + hid (before the blocks' SiLU)
__host__ __device__ inline long long ws_dst(const CondDev& p) { return ws_tfin(p) + 1LL * p.B * p.td; }             // [B][td]   backward: d SiLU(t)
__host__ __device__ inline long long ws_din(const CondDev& p) { return ws_dst(p) + 1LL * p.B * p.td; }              // [2 + T][td] backward: d wse, d bse, d null_token
__host__ __device__ inline long long ws_total(const CondDev& p) { return ws_din(p) + 1LL * (2 + p.T) * p.td; }

__device__ __forceinline__ float silu_f(float x) { return x / (1.f + expf(-x)); }
__device__ __forceinline__ float dsilu_f(float x) {
  const float s = 1.f / (1.f + expf(-x));
  return s * (1.f + x * (1.f - s));
}
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float dgelu_f(float x) {
  return 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * 0.39894228040143268f * expf(-0.5f * x * x);
}

constexpr int CTK = 32;                 // k-tile of the staged weight rows
constexpr int CTP = 257;                // pitch of a staged k-row (256 outputs + 1: conflict-free transposed stores)

// acc[x] = sum_k W_row(tid)[k] * xs[x][k] for the output row of this thread (tid < nrows <= 256), x < nx <= NX.
// The rows live anywhere in global memory: row r starts at rows_s[r] (shared memory table).  Weight rows are staged through
// shared memory in tiles of 32 k: warp w loads rows 32 w .. 32 w + 31 with lane = k (128-byte coalesced requests, all 32 in
// flight), stores them transposed, and every thread then walks its own row out of shared memory.  (A thread reading its row
// straight from global memory issues 32 sectors per request and serialises on the L2 latency: measured 30 us per 256 x 256 layer.)
template <int NX>
__device__ __forceinline__ void tile_matvec(const float* const* rows_s, int nrows, int K, const float* xs, int nx, float* Ws, float* acc) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#pragma unroll
  for (int x = 0; x < NX; ++x) acc[x] = 0.f;
  for (int k0 = 0; k0 < K; k0 += CTK) {
    float v[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) {
      const int row = warp * 32 + r;
      v[r] = (row < nrows && k0 + lane < K) ? __ldg(rows_s[row] + k0 + lane) : 0.f;
    }
#pragma unroll
    for (int r = 0; r < 32; ++r) Ws[lane * CTP + warp * 32 + r] = v[r];
    __syncthreads();
    const int kn = min(CTK, K - k0);          // K is a multiple of 4 (host check)
    for (int kk = 0; kk < kn; kk += 4) {
      const float w0 = Ws[kk * CTP + tid], w1 = Ws[(kk + 1) * CTP + tid], w2 = Ws[(kk + 2) * CTP + tid], w3 = Ws[(kk + 3) * CTP + tid];
#pragma unroll
      for (int x = 0; x < NX; ++x)
        if (x < nx) {
          const float4 xv = *reinterpret_cast<const float4*>(xs + x * K + k0 + kk);     // same address for every thread: a broadcast
          acc[x] += w0 * xv.x + w1 * xv.y + w2 * xv.z + w3 * xv.w;
        }
    }
    __syncthreads();
  }
}

// rows of a dense [nrows][K] matrix at `W`
__device__ __forceinline__ void dense_rows(const float** rows_s, const float* W, int nrows, int K) {
  for (int r = threadIdx.x; r < 256; r += blockDim.x) rows_s[r] = r < nrows ? W + static_cast<long long>(r) * K : W;
  __syncthreads();
}

// The trunk for all samples; leaves SiLU(t + hid) in sA [B][td].  sB, sC: scratch [B][td].  save: write the intermediates the
// backward needs to the workspace (one CTA does).
template <int NB>
__device__ void cond_trunk(const CondDev& p, float* sA, float* sB, float* sC, float* Ws, const float** rows_s, bool save) {
  const int tid = threadIdx.x, td = p.td, dim = p.dim, B = p.B, T = p.T;
  const float* P = p.param;
  float* ws = p.ws;
  // sinusoidal embedding -> sB [B][dim]
  const int half = dim / 2;
  // (the library is built with --use_fast_math: sinf / cosf / expf would become the reduced-range intrinsics, which are not accurate
  // for arguments up to 255 rad, so the few hundred transcendental evaluations of this path run in double precision; the
  // frequency and the argument are rounded to fp32 where torch rounds them)
  const float rate = static_cast<float>(log(10000.0) / static_cast<double>(half - 1));
  for (int i = tid; i < B * dim; i += blockDim.x) {
    const int b = i / dim, c = i - b * dim;
    const float fr = static_cast<float>(exp(static_cast<double>(-rate * static_cast<float>(c < half ? c : c - half))));
    const float arg = static_cast<float>(p.time[b]) * fr;
    const float v = static_cast<float>(c < half ? sin(static_cast<double>(arg)) : cos(static_cast<double>(arg)));
    sB[i] = v;
    if (save) ws[ws_e(p) + i] = v;
  }
  __syncthreads();
  float acc[NB];
  // h1 = GELU(W1 e + b1) -> sC
  dense_rows(rows_s, P + p.o_w1, td, dim);
  tile_matvec<NB>(rows_s, td, dim, sB, B, Ws, acc);
  if (tid < td) {
    const float bb = __ldg(P + p.o_b1 + tid);
#pragma unroll
    for (int b = 0; b < NB; ++b)
      if (b < B) {
        acc[b] += bb;
        if (save) ws[ws_pre1(p) + b * td + tid] = acc[b];
        sC[b * td + tid] = gelu_f(acc[b]);
      }
  }
  __syncthreads();
  // tt = W2 h1 + b2 -> kept in registers (tacc) until hid is known
  float tacc[NB];
  dense_rows(rows_s, P + p.o_w2, td, td);
  tile_matvec<NB>(rows_s, td, td, sC, B, Ws, tacc);
  if (tid < td) {
    const float bb = __ldg(P + p.o_b2 + tid);
#pragma unroll
    for (int b = 0; b < NB; ++b) tacc[b] += bb;
  }
  __syncthreads();
  // mean token hm[b][k] = mean_j cond[b][j] * wse[k] + bse[k] -> sB ; LayerNorm -> sC
  if (tid < td) {
    const float wse = __ldg(P + p.o_wse + tid), bse = __ldg(P + p.o_bse + tid);
    for (int b = 0; b < B; ++b) {
      float m = 0.f;
      for (int j = 0; j < T; ++j) m += p.cond[b * T + j];
      const float v = (m / static_cast<float>(T)) * wse + bse;
      sB[b * td + tid] = v;
      if (save) ws[ws_hm(p) + b * td + tid] = v;
    }
  }
  __syncthreads();
  {   // one warp per sample: mean / rstd over td
    const int warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
    for (int b = warp; b < B; b += nw) {
      float s = 0.f;
      for (int k = lane; k < td; k += 32) s += sB[b * td + k];
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mean = s / td;
      float q = 0.f;
      for (int k = lane; k < td; k += 32) {
        const float d = sB[b * td + k] - mean;
        q += d * d;
      }
      for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
      const float rstd = rsqrtf(q / td + 1e-5f);
      for (int k = lane; k < td; k += 32)
        sC[b * td + k] = (sB[b * td + k] - mean) * rstd * __ldg(P + p.o_lng + k) + __ldg(P + p.o_lnb + k);
      if (save && lane == 0) {
        ws[ws_stat(p) + 2 * b] = mean;
        ws[ws_stat(p) + 2 * b + 1] = rstd;
      }
    }
  }
  __syncthreads();
  // g1 = W3 ln + b3 -> sB (SiLU applied) ; g2 = W4 silu(g1) + b4
  dense_rows(rows_s, P + p.o_w3, td, td);
  tile_matvec<NB>(rows_s, td, td, sC, B, Ws, acc);
  if (tid < td) {
    const float bb = __ldg(P + p.o_b3 + tid);
#pragma unroll
    for (int b = 0; b < NB; ++b)
      if (b < B) {
        acc[b] += bb;
        if (save) ws[ws_g1(p) + b * td + tid] = acc[b];
        sB[b * td + tid] = silu_f(acc[b]);          // (the mean tokens in sB were last read before the LayerNorm barrier)
      }
  }
  __syncthreads();
  dense_rows(rows_s, P + p.o_w4, td, td);
  tile_matvec<NB>(rows_s, td, td, sB, B, Ws, acc);
  if (tid < td) {
    const float nh = __ldg(P + p.o_nhid + tid), bb = __ldg(P + p.o_b4 + tid);
#pragma unroll
    for (int b = 0; b < NB; ++b)
      if (b < B) {
        const float tf = tacc[b] + (p.null_mask[b] ? nh : acc[b] + bb);
        if (save) ws[ws_tfin(p) + b * td + tid] = tf;
        sA[b * td + tid] = silu_f(tf);
      }
  }
  __syncthreads();
}

template <int NB>
__global__ void __launch_bounds__(256) cond_fwd_kernel(const __grid_constant__ CondDev p) {
  extern __shared__ float csm[];
  const int td = p.td, B = p.B, T = p.T, tid = threadIdx.x;
  float* sA = csm;
  float* sB = sA + B * td;
  float* sC = sB + B * td;
  float* sV = sC + B * td;                  // [2 + T][td]: wse, bse, null tokens (kv chunks)
  float* Ws = sV + (2 + T) * td;            // [32][257] staged weight tile
  const float** rows_s = reinterpret_cast<const float**>(Ws + CTK * CTP);          // [256] row pointers (all sizes before it are even)
  const int chunk = blockIdx.x;
  const float* P = p.param;
  if (chunk < p.mlp_chunks) {
    cond_trunk<NB>(p, sA, sB, sC, Ws, rows_s, chunk == 0);
    // ---- block MLPs: global column g -> (block, column)
    const int g = chunk * 256 + tid;
    const int ncol = min(256, p.res_col0[p.n_res] - chunk * 256);
    int j = 0, c = 0;
    if (tid < ncol) {
      while (g >= p.res_col0[j + 1]) ++j;
      c = g - p.res_col0[j];
    }
    rows_s[tid] = P + p.res_w[j] + static_cast<long long>(c) * td;
    __syncthreads();
    float acc[NB];
    tile_matvec<NB>(rows_s, ncol, td, sA, B, Ws, acc);
    if (tid < ncol) {
      const float bb = __ldg(P + p.res_b[j] + c);
#pragma unroll
      for (int b = 0; b < NB; ++b)
        if (b < B) p.out[p.res_out[j] + static_cast<long long>(b) * p.res_c2[j] + c] = acc[b] + bb;
    }
    if (chunk == 0) {
      // position bias [heads][f][f] and rotary tables [2][f][16][2]
      const int f = p.frames;
      for (int i = tid; i < p.heads * f * f; i += blockDim.x) {
        const int h = i / (f * f), ij = i - h * f * f;
        p.out[p.bias_out + i] = __ldg(P + p.o_table + p.buckets[ij] * p.heads + h);
      }
      for (int i = tid; i < f * 16; i += blockDim.x) {
        const int fr = i / 16, k = i - fr * 16;
        const float ang = static_cast<float>(fr) * __ldg(p.freqs + k);
        const float cs = static_cast<float>(cos(static_cast<double>(ang))), sn = static_cast<float>(sin(static_cast<double>(ang)));
        const float sc = 0.17677669529663687f;        // 32^-1/2: table 0 serves the queries
        p.out[p.rot_out + (fr * 16 + k) * 2] = cs * sc;
        p.out[p.rot_out + (fr * 16 + k) * 2 + 1] = sn * sc;
        p.out[p.rot_out + f * 32 + (fr * 16 + k) * 2] = cs;
        p.out[p.rot_out + f * 32 + (fr * 16 + k) * 2 + 1] = sn;
      }
    }
  } else {
    // ---- keys | values of one attention block: chunk -> (block a, half: 0 = keys, 1 = values), thread = column
    const int kc = chunk - p.mlp_chunks;
    const int a = kc >> 1, half = kc & 1;
    for (int i = tid; i < (2 + T) * td; i += blockDim.x) {
      const int r = i / td, k = i - r * td;
      sV[i] = r == 0 ? __ldg(P + p.o_wse + k) : (r == 1 ? __ldg(P + p.o_bse + k) : __ldg(P + p.o_ntok + (r - 2) * td + k));
    }
    dense_rows(rows_s, P + (half ? p.att_wv[a] : p.att_wk[a]), 256, td);
    float acc[2 + CMAXT];
    tile_matvec<2 + CMAXT>(rows_s, 256, td, sV, 2 + T, Ws, acc);
    if (tid < 256) {
      const bool rotate = half == 0 && p.att_temporal[a];
      const int pair = (tid & 31) >> 1;
      const float fq = __ldg(p.freqs + pair);
      float rc[CMAXT], rs[CMAXT];
#pragma unroll
      for (int j = 0; j < CMAXT; ++j) {
        const float ang = static_cast<float>(j) * fq;
        rc[j] = (rotate && j < T) ? static_cast<float>(cos(static_cast<double>(ang))) : 1.f;
        rs[j] = (rotate && j < T) ? static_cast<float>(sin(static_cast<double>(ang))) : 0.f;
      }
      for (int b = 0; b < B; ++b) {
        const bool nul = p.null_mask[b] != 0;
#pragma unroll
        for (int j = 0; j < CMAXT; ++j)
          if (j < T) {
            float v = nul ? acc[2 + j] : p.cond[b * T + j] * acc[0] + acc[1];
            const float other = __shfl_xor_sync(0xffffffffu, v, 1);      // the partner of the rotary pair (adjacent column)
            if (rotate) v = (tid & 1) ? v * rc[j] + other * rs[j] : v * rc[j] - other * rs[j];
            p.out[p.att_out[a] + (static_cast<long long>(b) * T + j) * 512 + half * 256 + tid] = v;
          }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward, kernel 1: per chunk, thread = input index k; loops over the chunk's columns
// ------------------------------------------------------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(256) cond_bwd_main_kernel(const __grid_constant__ CondDev p) {
  extern __shared__ float csm[];
  const int td = p.td, B = p.B, T = p.T, tid = threadIdx.x;
  const int chunk = blockIdx.x;
  const float* P = p.param;
  float* G = p.grad;
  if (chunk < p.mlp_chunks) {
    float* sD = csm;                 // [B][256] gradients of the chunk's columns
    long long* rowoff_s = reinterpret_cast<long long*>(sD + B * 256);      // [256] arena offset of each column's weight row
    const int g0 = chunk * 256;
    const int ncol = min(256, p.res_col0[p.n_res] - g0);
    // stage d(scale|shift) of the chunk, thread = column
    if (tid < ncol) {
      const int g = g0 + tid;
      int jb = 0;
      while (g >= p.res_col0[jb + 1]) ++jb;
      const int cb = g - p.res_col0[jb];
      float sum = 0.f;
      for (int b = 0; b < B; ++b) {
        const float d = p.out[p.res_out[jb] + static_cast<long long>(b) * p.res_c2[jb] + cb];
        sD[b * 256 + tid] = d;
        sum += d;
      }
      G[p.res_b[jb] + cb] += sum;
      rowoff_s[tid] = p.res_w[jb] + static_cast<long long>(cb) * td;
    }
    __syncthreads();
    if (tid < td) {
      float st[NB], dacc[NB];
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        st[b] = b < B ? silu_f(p.ws[ws_tfin(p) + b * td + tid]) : 0.f;
        dacc[b] = 0.f;
      }
      // eight weight rows (and their gradient rows) in flight per step: the loop is otherwise one L2 round trip per column
      for (int c0 = 0; c0 < ncol; c0 += 8) {
        float w[8], gq[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const bool ok = c0 + u < ncol;
          const long long row = ok ? rowoff_s[c0 + u] + tid : 0;
          w[u] = ok ? __ldg(P + row) : 0.f;
          gq[u] = ok ? G[row] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (c0 + u < ncol) {
            float dw = 0.f;
#pragma unroll
            for (int b = 0; b < NB; ++b)
              if (b < B) {
                const float d = sD[b * 256 + c0 + u];
                dacc[b] += d * w[u];
                dw += d * st[b];
              }
            G[rowoff_s[c0 + u] + tid] = gq[u] + dw;
          }
        }
      }
#pragma unroll
      for (int b = 0; b < NB; ++b)
        if (b < B) atomicAdd(p.ws + ws_dst(p) + b * td + tid, dacc[b]);
    }
    if (chunk == 0) {
      const int f = p.frames;
      for (int i = tid; i < p.heads * f * f; i += blockDim.x) {
        const int h = i / (f * f), ij = i - h * f * f;
        atomicAdd(G + p.o_table + p.buckets[ij] * p.heads + h, p.out[p.bias_out + i]);
      }
    }
  } else {
    const int kc = chunk - p.mlp_chunks;
    const int a = kc >> 1, half = kc & 1;
    float* sR = csm;                 // [2 + T][256]: d(Wk wse), d(Wk bse), d(Wk null_j) per column
    float* sV = sR + (2 + T) * 256;  // [2 + T][td]: wse, bse, null tokens
    for (int i = tid; i < (2 + T) * td; i += blockDim.x) {
      const int r = i / td, k = i - r * td;
      sV[i] = r == 0 ? __ldg(P + p.o_wse + k) : (r == 1 ? __ldg(P + p.o_bse + k) : __ldg(P + p.o_ntok + (r - 2) * td + k));
    }
    {   // thread = column: reduce d ek over (b, j), un-rotating the keys of the temporal blocks
      const bool rotate = half == 0 && p.att_temporal[a];
      const int pair = (tid & 31) >> 1;
      const float fq = __ldg(p.freqs + pair);
      float du = 0.f, dv = 0.f, dn[CMAXT], rc[CMAXT], rs[CMAXT];
#pragma unroll
      for (int j = 0; j < CMAXT; ++j) {
        dn[j] = 0.f;
        const float ang = static_cast<float>(j) * fq;
        rc[j] = (rotate && j < T) ? static_cast<float>(cos(static_cast<double>(ang))) : 1.f;
        rs[j] = (rotate && j < T) ? static_cast<float>(sin(static_cast<double>(ang))) : 0.f;
      }
      for (int b = 0; b < B; ++b) {
        const bool nul = p.null_mask[b] != 0;
#pragma unroll
        for (int j = 0; j < CMAXT; ++j)
          if (j < T) {
            float d = p.out[p.att_out[a] + (static_cast<long long>(b) * T + j) * 512 + half * 256 + tid];
            const float other = __shfl_xor_sync(0xffffffffu, d, 1);
            if (rotate) d = (tid & 1) ? d * rc[j] - other * rs[j] : d * rc[j] + other * rs[j];      // transpose of the rotation
            if (nul) dn[j] += d;
            else {
              du += p.cond[b * T + j] * d;
              dv += d;
            }
          }
      }
      sR[tid] = du;
      sR[256 + tid] = dv;
#pragma unroll
      for (int j = 0; j < CMAXT; ++j)
        if (j < T) sR[(2 + j) * 256 + tid] = dn[j];
    }
    __syncthreads();
    if (tid < td) {
      const long long wbase = half ? p.att_wv[a] : p.att_wk[a];
      float vin[2 + CMAXT], din[2 + CMAXT];
#pragma unroll
      for (int r = 0; r < 2 + CMAXT; ++r) {
        vin[r] = r < 2 + T ? sV[r * td + tid] : 0.f;
        din[r] = 0.f;
      }
      for (int c0 = 0; c0 < 256; c0 += 8) {
        float w[8], gq[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const long long row = wbase + static_cast<long long>(c0 + u) * td + tid;
          w[u] = __ldg(P + row);
          gq[u] = G[row];
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          float dw = 0.f;
#pragma unroll
          for (int r = 0; r < 2 + CMAXT; ++r)
            if (r < 2 + T) {
              const float d = sR[r * 256 + c0 + u];
              din[r] += d * w[u];
              dw += d * vin[r];
            }
          G[wbase + static_cast<long long>(c0 + u) * td + tid] = gq[u] + dw;
        }
      }
#pragma unroll
      for (int r = 0; r < 2 + CMAXT; ++r)
        if (r < 2 + T) atomicAdd(p.ws + ws_din(p) + r * td + tid, din[r]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward, kernel 2 (two CTAs of 256 threads, one per independent branch): through the trunk.  thread = k; every layer loops
// over its outputs o.
// ------------------------------------------------------------------------------------------------
template <int NB>
__device__ __forceinline__ void layer_bwd(const float* __restrict__ P, float* __restrict__ G, long long o_w, long long o_b, int K, int nout, int B,
                                          const float* s_dy /* [B][nout] */, const float* a /* registers: input[b] at index k */, int k,
                                          float* dx /* out: d input[b] at index k */) {
#pragma unroll
  for (int b = 0; b < NB; ++b) dx[b] = 0.f;
  if (k < K) {
    constexpr int U = NB <= 8 ? 16 : 8;         // weight rows (and their gradient rows) in flight
    for (int o0 = 0; o0 < nout; o0 += U) {
      float w[U], gq[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool ok = o0 + u < nout;
        const long long row = o_w + static_cast<long long>(ok ? o0 + u : 0) * K + k;
        w[u] = ok ? __ldg(P + row) : 0.f;
        gq[u] = ok ? G[row] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (o0 + u < nout) {
          float dw = 0.f;
#pragma unroll
          for (int b = 0; b < NB; ++b)
            if (b < B) {
              const float d = s_dy[b * nout + o0 + u];
              dx[b] += d * w[u];
              dw += d * a[b];
            }
          G[o_w + static_cast<long long>(o0 + u) * K + k] = gq[u] + dw;
        }
      }
    }
  }
  if (k < nout) {
    float db = 0.f;
    for (int b = 0; b < B; ++b) db += s_dy[b * nout + k];
    G[o_b + k] += db;
  }
}

template <int NB>
__global__ void __launch_bounds__(256) cond_bwd_tail_kernel(const __grid_constant__ CondDev p) {
  extern __shared__ float csm[];
  const int td = p.td, dim = p.dim, B = p.B, T = p.T, k = threadIdx.x;
  const float* P = p.param;
  float* G = p.grad;
  const float* ws = p.ws;
  float* sD = csm;                 // [B][td] upstream gradient of the current layer
  float* sE = sD + B * td;         // [B][td] scratch
  float a[NB], dx[NB];
  // d(t + hid) = d SiLU(t + hid) * SiLU'
  if (k < td)
    for (int b = 0; b < B; ++b) sD[b * td + k] = ws[ws_dst(p) + b * td + k] * dsilu_f(ws[ws_tfin(p) + b * td + k]);
  __syncthreads();
  if (blockIdx.x == 0) {
    // ---- CTA 0, time MLP: tt = W2 h1 + b2, h1 = GELU(pre1), pre1 = W1 e + b1
#pragma unroll
    for (int b = 0; b < NB; ++b) a[b] = (b < B && k < td) ? gelu_f(ws[ws_pre1(p) + b * td + k]) : 0.f;
    layer_bwd<NB>(P, G, p.o_w2, p.o_b2, td, td, B, sD, a, k, dx);
    if (k < td)
      for (int b = 0; b < B; ++b) sE[b * td + k] = dx[b] * dgelu_f(ws[ws_pre1(p) + b * td + k]);
    __syncthreads();
#pragma unroll
    for (int b = 0; b < NB; ++b) a[b] = (b < B && k < dim) ? ws[ws_e(p) + b * dim + k] : 0.f;
    layer_bwd<NB>(P, G, p.o_w1, p.o_b1, dim, td, B, sE, a, k, dx);
    return;
  }
  // ---- CTA 1, hidden path: dropped samples feed null_text_hidden, the others cond_token_to_hidden
  if (k < td) {
    float dn = 0.f;
    for (int b = 0; b < B; ++b) {
      if (p.null_mask[b]) {
        dn += sD[b * td + k];
        sD[b * td + k] = 0.f;
      }
    }
    G[p.o_nhid + k] += dn;
  }
  __syncthreads();
#pragma unroll
  for (int b = 0; b < NB; ++b) a[b] = (b < B && k < td) ? silu_f(ws[ws_g1(p) + b * td + k]) : 0.f;
  layer_bwd<NB>(P, G, p.o_w4, p.o_b4, td, td, B, sD, a, k, dx);
  if (k < td)
    for (int b = 0; b < B; ++b) sE[b * td + k] = dx[b] * dsilu_f(ws[ws_g1(p) + b * td + k]);
  __syncthreads();
  // LayerNorm output as the input of W3: xhat * gamma + beta
  float xh[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    xh[b] = (b < B && k < td) ? (ws[ws_hm(p) + b * td + k] - ws[ws_stat(p) + 2 * b]) * ws[ws_stat(p) + 2 * b + 1] : 0.f;
    a[b] = (k < td) ? xh[b] * __ldg(P + p.o_lng + k) + __ldg(P + p.o_lnb + k) : 0.f;
  }
  layer_bwd<NB>(P, G, p.o_w3, p.o_b3, td, td, B, sE, a, k, dx);      // dx = d LayerNorm output
  __syncthreads();
  // LayerNorm backward: dgamma, dbeta, d mean token
  float gd[NB];
  if (k < td) {
    float dg = 0.f, dbt = 0.f;
    const float gam = __ldg(P + p.o_lng + k);
#pragma unroll
    for (int b = 0; b < NB; ++b)
      if (b < B) {
        dg += dx[b] * xh[b];
        dbt += dx[b];
        gd[b] = dx[b] * gam;
        sD[b * td + k] = gd[b];
        sE[b * td + k] = gd[b] * xh[b];
      }
    G[p.o_lng + k] += dg;
    G[p.o_lnb + k] += dbt;
  }
  __syncthreads();
  if (k < td) {
    float dwse = ws[ws_din(p) + k], dbse = ws[ws_din(p) + td + k];
    for (int b = 0; b < B; ++b) {
      if (p.null_mask[b]) continue;          // their mean token never reached the output
      float m1 = 0.f, m2 = 0.f;
      for (int i = 0; i < td; ++i) {
        m1 += sD[b * td + i];
        m2 += sE[b * td + i];
      }
      m1 /= td;
      m2 /= td;
      const float dhm = ws[ws_stat(p) + 2 * b + 1] * (gd[b] - m1 - xh[b] * m2);
      float cm = 0.f;
      for (int j = 0; j < T; ++j) cm += p.cond[b * T + j];
      dwse += (cm / static_cast<float>(T)) * dhm;
      dbse += dhm;
    }
    G[p.o_wse + k] += dwse;
    G[p.o_bse + k] += dbse;
    for (int j = 0; j < T; ++j) G[p.o_ntok + j * td + k] += ws[ws_din(p) + (2 + j) * td + k];
  }
}

}  // namespace vmm

using namespace vmm;

static int cond_fill(const vmm_cond_params* h, CondDev& d, const char* who) {
  if (!h) return set_error(VMM_ERR_ARG, "vmm_cond: null params");
  if (h->B < 1 || h->B > CMAXB) return set_error(VMM_ERR_UNSUPPORTED, "vmm_cond: batch above 32 (split the batch)");
  if (h->T < 1 || h->T > CMAXT || h->td < 4 || h->td > 256 || (h->td % 4) != 0 || h->dim < 4 || h->dim > 256 || (h->dim % 4) != 0 || h->heads != 8)
    return set_error(VMM_ERR_UNSUPPORTED, "vmm_cond: needs td, dim <= 256 (multiples of 4), 8 heads, <= 16 tokens");
  if (h->n_res < 1 || h->n_res > CMAXBLK || h->n_attn < 0 || h->n_attn > CMAXBLK) return set_error(VMM_ERR_ARG, "vmm_cond: block counts");
  if (!h->time || !h->cond || !h->null_mask || !h->param || !h->freqs || !h->buckets || !h->out || !h->ws)
    return set_error(VMM_ERR_ARG, who);
  memset(&d, 0, sizeof(d));
  d.B = h->B; d.T = h->T; d.dim = h->dim; d.td = h->td; d.heads = h->heads; d.frames = h->frames; d.n_res = h->n_res; d.n_attn = h->n_attn;
  d.time = reinterpret_cast<const long long*>(h->time); d.cond = h->cond; d.null_mask = h->null_mask; d.param = h->param; d.grad = h->grad; d.freqs = h->freqs; d.buckets = reinterpret_cast<const int*>(h->buckets);
  d.o_w1 = h->o_w1; d.o_b1 = h->o_b1; d.o_w2 = h->o_w2; d.o_b2 = h->o_b2; d.o_wse = h->o_wse; d.o_bse = h->o_bse; d.o_lng = h->o_lng; d.o_lnb = h->o_lnb;
  d.o_w3 = h->o_w3; d.o_b3 = h->o_b3; d.o_w4 = h->o_w4; d.o_b4 = h->o_b4; d.o_ntok = h->o_ntok; d.o_nhid = h->o_nhid; d.o_table = h->o_table;
  int col = 0;
  for (int j = 0; j < h->n_res; ++j) {
    if (h->res_c2[j] < 1) return set_error(VMM_ERR_ARG, "vmm_cond: res_c2");
    d.res_w[j] = h->res_w[j]; d.res_b[j] = h->res_b[j]; d.res_c2[j] = h->res_c2[j]; d.res_out[j] = h->res_out[j];
    d.res_col0[j] = col;
    col += h->res_c2[j];
  }
  d.res_col0[h->n_res] = col;
  for (int a = 0; a < h->n_attn; ++a) {
    d.att_wk[a] = h->att_wk[a]; d.att_wv[a] = h->att_wv[a]; d.att_out[a] = h->att_out[a]; d.att_temporal[a] = h->att_temporal[a];
  }
  d.bias_out = h->bias_out; d.rot_out = h->rot_out; d.out = h->out; d.ws = h->ws;
  d.mlp_chunks = (col + 255) / 256;
  d.kv_chunks = 2 * h->n_attn;
  return VMM_OK;
}

extern "C" size_t vmm_cond_workspace(int B, int T, int dim, int td) {
  CondDev d;
  memset(&d, 0, sizeof(d));
  d.B = B; d.T = T; d.dim = dim; d.td = td;
  return static_cast<size_t>(ws_total(d)) * sizeof(float);
}

#define COND_DISPATCH(KERNEL, GRID, SMEM)                                                       \
  do {                                                                                          \
    if (d.B <= 8) KERNEL<8><<<GRID, 256, SMEM, stream>>>(d);                                    \
    else if (d.B <= 16) KERNEL<16><<<GRID, 256, SMEM, stream>>>(d);                             \
    else KERNEL<32><<<GRID, 256, SMEM, stream>>>(d);                                            \
  } while (0)

extern "C" int vmm_cond_fwd(const vmm_cond_params* hp, void* stream_) {
  CondDev d;
  int rc = cond_fill(hp, d, "vmm_cond_fwd: null pointer");
  if (rc) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const size_t smem = (static_cast<size_t>(3) * d.B * d.td + static_cast<size_t>(2 + d.T) * d.td + CTK * CTP) * sizeof(float) + 256 * sizeof(void*);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(cond_fwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(cond_fwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(cond_fwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return set_cuda_error(e, "vmm_cond_fwd: attr");
    attr = true;
  }
  COND_DISPATCH(cond_fwd_kernel, d.mlp_chunks + d.kv_chunks, smem);
  count_launch();
  return check_launch("vmm_cond_fwd");
}

extern "C" int vmm_cond_bwd(const vmm_cond_params* hp, void* stream_) {
  CondDev d;
  int rc = cond_fill(hp, d, "vmm_cond_bwd: null pointer");
  if (rc) return rc;
  if (!d.grad) return set_error(VMM_ERR_ARG, "vmm_cond_bwd: null gradient arena");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  // scratch accumulators of kernel 1
  cudaError_t e = cudaMemsetAsync(d.ws + ws_dst(d), 0, static_cast<size_t>(ws_total(d) - ws_dst(d)) * sizeof(float), stream);
  if (e != cudaSuccess) return set_cuda_error(e, "vmm_cond_bwd: memset");
  const size_t smem1 = static_cast<size_t>(std::max(d.B * 256 + 512, (2 + d.T) * 256 + (2 + d.T) * d.td)) * sizeof(float);
  const size_t smem2 = static_cast<size_t>(2) * d.B * d.td * sizeof(float);
  static bool attr = false;
  if (!attr) {
    e = cudaFuncSetAttribute(cond_bwd_tail_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e != cudaSuccess) return set_cuda_error(e, "vmm_cond_bwd: attr");
    attr = true;
  }
  COND_DISPATCH(cond_bwd_main_kernel, d.mlp_chunks + d.kv_chunks, smem1);
  count_launch();
  COND_DISPATCH(cond_bwd_tail_kernel, 2, smem2);     // CTA 0: time MLP, CTA 1: hidden path / tokens (independent branches)
  count_launch();
  return check_launch("vmm_cond_bwd");
}
