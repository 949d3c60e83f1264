// Elementwise / selection kernels around the UNet: network-input preparation (q_sample fused), the
// ancestral and DDIM sampling updates with classifier-free guidance, the exact per-sample quantile used by
// dynamic thresholding, the L1 / L2 training loss with its gradient, and a fused Adam (+EMA) step.
// All HBM/L2-bound: vectorised where the layout allows, one launch per step of the chain.
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace vmm {

// ------------------------------------------------------------------------------------------------
// Network input.  x: fp32 (B, C, F, H, W) (reference layout).  Writes the init_conv operand
//   xin[bf][y][w'][8] 16-bit, w' in [0, W+6): pixel w sits at w' = w + 3, channels >= C and the 3-pixel
//   borders are zero, so the (1,7,7) 'zeros'-padded conv (VDDP:626) becomes 7 row-taps of 8 px x 8 ch.
//   value = a[b] * x + c[b] + s[b] * noise        (q_sample VDDP:1036-1042 and normalize_img VDDP:1109)
// ------------------------------------------------------------------------------------------------
__global__ void prep_input_kernel(const float* __restrict__ x, const float* __restrict__ noise, const float* __restrict__ a,
                                  const float* __restrict__ c, const float* __restrict__ s, uint16_t* __restrict__ xin, int fmt,
                                  int B, int C, int F, int H, int W) {
  const long long total = static_cast<long long>(B) * F * H * (W + 6);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int wp = static_cast<int>(i % (W + 6));
    long long r = i / (W + 6);
    const int y = static_cast<int>(r % H);
    r /= H;
    const int f = static_cast<int>(r % F);
    const int b = static_cast<int>(r / F);
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int w = wp - 3;
    if (w >= 0 && w < W) {
      for (int ch = 0; ch < C; ++ch) {
        const long long src = (((static_cast<long long>(b) * C + ch) * F + f) * H + y) * W + w;
        float val = (a ? a[b] : 1.f) * x[src] + (c ? c[b] : 0.f);
        if (noise) val += s[b] * noise[src];
        v[ch] = val;
      }
    }
    uint4 q;
    q.x = pack2_h16(v[0], v[1], fmt);
    q.y = pack2_h16(v[2], v[3], fmt);
    q.z = pack2_h16(v[4], v[5], fmt);
    q.w = pack2_h16(v[6], v[7], fmt);
    reinterpret_cast<uint4*>(xin)[i] = q;
  }
}

// ------------------------------------------------------------------------------------------------
// Loss.  pred: fp32 channels-last [B*F*H*W][C];  target (noise): fp32 (B, C, F, H, W).
//   loss += sum |target - pred| (l1) or (target - pred)^2 (l2)   [caller divides by N]
//   dpred[pix][8] 16-bit (zero padded to 8 channels) = d loss_mean / d pred * grad_scale
// ------------------------------------------------------------------------------------------------
__global__ void loss_kernel(const float* __restrict__ pred, const float* __restrict__ target, float* __restrict__ loss_sum,
                            uint16_t* __restrict__ dpred, int fmt, int B, int C, int F, int H, int W, int l2, float grad_scale) {
  const long long npix = static_cast<long long>(B) * F * H * W;
  const float inv_n = 1.f / (static_cast<float>(npix) * C);
  float local = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < npix;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int w = static_cast<int>(i % W);
    long long r = i / W;
    const int y = static_cast<int>(r % H);
    r /= H;
    const int f = static_cast<int>(r % F);
    const int b = static_cast<int>(r / F);
    float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int ch = 0; ch < C; ++ch) {
      const float t = target[(((static_cast<long long>(b) * C + ch) * F + f) * H + y) * W + w];
      const float d = pred[i * C + ch] - t;
      if (l2) {
        local += d * d;
        g[ch] = 2.f * d * inv_n * grad_scale;
      } else {
        local += fabsf(d);
        g[ch] = (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * inv_n * grad_scale;
      }
    }
    if (dpred) {
      uint4 q;
      q.x = pack2_h16(g[0], g[1], fmt);
      q.y = pack2_h16(g[2], g[3], fmt);
      q.z = pack2_h16(g[4], g[5], fmt);
      q.w = pack2_h16(g[6], g[7], fmt);
      reinterpret_cast<uint4*>(dpred)[i] = q;
    }
  }
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  __shared__ float wsum[32];
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? wsum[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) atomicAdd(loss_sum, v * inv_n);
  }
}

// ------------------------------------------------------------------------------------------------
// Sampling.  State x_t and the outputs are fp32 in the reference layout (B, C, F, H, W).  The network output
// eps is fp32 channels-last [2B or B][F*H*W][C]: rows [0,B) conditional, rows [B,2B) unconditional (the two
// forwards of forward_with_guidance_scale VDDP:723-728 run as one batch).
//   eps = null + (cond - null) * w ;  x0 = sr[b] * x - srm1[b] * eps         VDDP:728, 920-924
// ------------------------------------------------------------------------------------------------
__global__ void cfg_x0_kernel(const float* __restrict__ x, const float* __restrict__ eps_cl, int has_null, float w,
                              const float* __restrict__ sr, const float* __restrict__ srm1, float* __restrict__ x0,
                              float* __restrict__ eps_out, int B, int C, int F, int H, int W) {
  const long long per = static_cast<long long>(C) * F * H * W;
  const long long fhw = static_cast<long long>(F) * H * W;
  const long long total = per * B;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / per);
    const long long r = i % per;
    const int ch = static_cast<int>(r / fhw);
    const long long pos = r % fhw;
    const float ec = eps_cl[(static_cast<long long>(b) * fhw + pos) * C + ch];
    float e = ec;
    if (has_null) {
      const float en = eps_cl[(static_cast<long long>(B + b) * fhw + pos) * C + ch];
      e = en + (ec - en) * w;
    }
    if (eps_out) eps_out[i] = e;
    x0[i] = sr[b] * x[i] - srm1[b] * e;
  }
}

// Exact k-th / (k+1)-th order statistics of |v| per sample by 4-pass 8-bit radix select (one CTA per sample),
// then s = max(lerp(v_k, v_k1, frac), 1) as torch.quantile(..., 'linear') + clamp_(min=1) do.  VDDP:941-947
__global__ void __launch_bounds__(1024) abs_quantile_kernel(const float* __restrict__ v, long long n, long long k, float frac,
                                                            float floor_val, float* __restrict__ s_out) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned int sel_prefix, sel_rank;
  __shared__ unsigned int red[32];
  const float* vb = v + static_cast<long long>(blockIdx.x) * n;
  unsigned int prefix = 0;              // high bits decided so far
  unsigned int rank = static_cast<unsigned int>(k);   // rank of the wanted element among those matching the prefix
  unsigned int count_le_total = 0;      // number of elements <= selected value (for the k+1 query)
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const unsigned int mask_hi = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
    // warp-aggregated histogram: lanes that fall in the same bucket elect one leader to add their count,
    // so the heavily shared exponent buckets of the first pass do not serialise on one shared-memory word
    for (long long base = 0; base < n; base += blockDim.x) {
      const long long i = base + threadIdx.x;
      unsigned int u = 0;
      bool ok = false;
      if (i < n) {
        u = __float_as_uint(fabsf(vb[i]));
        ok = (u & mask_hi) == prefix;
      }
      const unsigned int live = __ballot_sync(0xffffffffu, ok);
      if (ok) {
        const unsigned int bkt = (u >> shift) & 0xFF;
        const unsigned int peers = __match_any_sync(live, bkt);
        if ((threadIdx.x & 31) == static_cast<unsigned int>(__ffs(peers) - 1)) atomicAdd(&hist[bkt], __popc(peers));
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int acc = 0, r = rank;
      int bsel = 255;
      for (int bkt = 0; bkt < 256; ++bkt) {
        if (acc + hist[bkt] > r) {
          bsel = bkt;
          break;
        }
        acc += hist[bkt];
      }
      sel_prefix = prefix | (static_cast<unsigned int>(bsel) << shift);
      sel_rank = r - acc;
    }
    __syncthreads();
    prefix = sel_prefix;
    rank = sel_rank;
    __syncthreads();
  }
  const float vk = __uint_as_float(prefix);
  // count elements <= vk and the smallest element > vk
  unsigned int cnt = 0;
  float nxt = 3.4e38f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const float a = fabsf(vb[i]);
    if (a <= vk) ++cnt;
    else nxt = fminf(nxt, a);
  }
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    nxt = fminf(nxt, __shfl_xor_sync(0xffffffffu, nxt, o));
  }
  __shared__ float redf[32];
  if ((threadIdx.x & 31) == 0) {
    red[threadIdx.x >> 5] = cnt;
    redf[threadIdx.x >> 5] = nxt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i2 = 0; i2 < (blockDim.x >> 5); ++i2) {
      count_le_total += red[i2];
      nxt = i2 == 0 ? redf[0] : fminf(nxt, redf[i2]);
    }
    const float vk1 = (count_le_total > static_cast<unsigned int>(k) + 1u || k + 1 >= n) ? vk : nxt;
    // torch.lerp: w < 0.5 ? a + w (b - a) : b - (b - a)(1 - w)
    const float d = vk1 - vk;
    float q = frac < 0.5f ? vk + frac * d : vk1 - d * (1.f - frac);
    s_out[blockIdx.x] = fmaxf(q, floor_val);
  }
}

// The same selection spread over the whole GPU (one launch per pass; the single CTA per sample above took 494 us for 4 x 304 128
// values, 5 % of a guided sampling step, on 4 of 148 SMs).  hist[b][pass][256] lives in a caller-provided workspace zeroed by the entry
// point.  A CTA of pass p first replays the bucket choices of passes 0 .. p-1 from their finished histograms (p scans of 256 counters),
// then counts its slice of the sample among the values that match the prefix.
__device__ __forceinline__ void quantile_replay(const unsigned int* __restrict__ hist_b, int passes, unsigned int k, unsigned int* s_hist,
                                                unsigned int& prefix, unsigned int& rank) {
  __shared__ unsigned int s_sel[2];
  prefix = 0;
  rank = k;
  for (int q = 0; q < passes; ++q) {
    const int shift = 24 - 8 * q;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = __ldcg(hist_b + q * 256 + i);
    __syncthreads();
    if (threadIdx.x < 32) {
      // lane l owns buckets 8 l .. 8 l + 7: warp scan of the lane sums, then the bucket inside the lane
      unsigned int c[8], sum = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        c[j] = s_hist[threadIdx.x * 8 + j];
        sum += c[j];
      }
      unsigned int incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (threadIdx.x >= static_cast<unsigned int>(o)) incl += t;
      }
      const unsigned int excl = incl - sum;
      const bool mine = excl <= rank && rank < incl;
      const unsigned int any = __ballot_sync(0xffffffffu, mine);
      if (any == 0) {                       // rank beyond the counted values (cannot happen for k < n): last bucket, as the one-CTA kernel
        if (threadIdx.x == 31) {
          s_sel[0] = prefix | (255u << shift);
          s_sel[1] = rank - incl;
        }
      } else if (mine) {
        unsigned int acc = excl;
        int bsel = 7;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (acc + c[j] > rank) { bsel = j; break; }
          acc += c[j];
        }
        s_sel[0] = prefix | (static_cast<unsigned int>(threadIdx.x * 8 + bsel) << shift);
        s_sel[1] = rank - acc;
      }
    }
    __syncthreads();
    prefix = s_sel[0];
    rank = s_sel[1];
    __syncthreads();
  }
}

template <int PASS>
__global__ void __launch_bounds__(512) quantile_hist_kernel(const float* __restrict__ v, long long n, unsigned int k, unsigned int* __restrict__ hist) {
  __shared__ unsigned int s_hist[256];
  const int b = blockIdx.y;
  const float* vb = v + static_cast<long long>(b) * n;
  unsigned int* hist_b = hist + static_cast<size_t>(b) * 4 * 256;
  unsigned int prefix, rank;
  quantile_replay(hist_b, PASS, k, s_hist, prefix, rank);
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  constexpr int shift = 24 - 8 * PASS;
  constexpr unsigned int mask_hi = PASS == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
  const long long per = (n + gridDim.x - 1) / gridDim.x;
  const long long i0 = per * blockIdx.x, i1 = min(i0 + per, n);
  for (long long base = i0; base < i1; base += blockDim.x) {
    const long long i = base + threadIdx.x;
    unsigned int u = 0;
    bool ok = false;
    if (i < i1) {
      u = __float_as_uint(fabsf(vb[i]));
      ok = (u & mask_hi) == prefix;
    }
    const unsigned int live = __ballot_sync(0xffffffffu, ok);
    if (ok) {
      const unsigned int bkt = (u >> shift) & 0xFF;
      const unsigned int peers = __match_any_sync(live, bkt);
      if ((threadIdx.x & 31) == static_cast<unsigned int>(__ffs(peers) - 1)) atomicAdd(&s_hist[bkt], __popc(peers));
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 256; i += blockDim.x)
    if (s_hist[i]) atomicAdd(hist_b + PASS * 256 + i, s_hist[i]);
}

// tail[b] = {count of values <= v_k, bits of the smallest value > v_k, CTAs done, -}; the last CTA of a sample writes s_out[b]
__global__ void __launch_bounds__(512) quantile_final_kernel(const float* __restrict__ v, long long n, unsigned int k, float frac, float floor_val,
                                                             const unsigned int* __restrict__ hist, unsigned int* __restrict__ tail,
                                                             float* __restrict__ s_out) {
  __shared__ unsigned int s_hist[256];
  __shared__ unsigned int red[16];
  __shared__ float redf[16];
  __shared__ int s_last;
  const int b = blockIdx.y;
  const float* vb = v + static_cast<long long>(b) * n;
  unsigned int prefix, rank;
  quantile_replay(hist + static_cast<size_t>(b) * 4 * 256, 4, k, s_hist, prefix, rank);
  const float vk = __uint_as_float(prefix);
  const long long per = (n + gridDim.x - 1) / gridDim.x;
  const long long i0 = per * blockIdx.x, i1 = min(i0 + per, n);
  unsigned int cnt = 0;
  float nxt = 3.4e38f;
  for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    const float a = fabsf(vb[i]);
    if (a <= vk) ++cnt;
    else nxt = fminf(nxt, a);
  }
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    nxt = fminf(nxt, __shfl_xor_sync(0xffffffffu, nxt, o));
  }
  if ((threadIdx.x & 31) == 0) {
    red[threadIdx.x >> 5] = cnt;
    redf[threadIdx.x >> 5] = nxt;
  }
  __syncthreads();
  unsigned int* tb = tail + static_cast<size_t>(b) * 4;
  if (threadIdx.x == 0) {
    for (int w = 1; w < (blockDim.x >> 5); ++w) {
      cnt += red[w];
      nxt = fminf(nxt, redf[w]);
    }
    atomicAdd(tb, cnt);
    atomicMax(tb + 1, ~__float_as_uint(nxt));      // complemented bits: a larger word is a smaller (non-negative) value, and the zeroed workspace means "none yet"
    __threadfence();
    s_last = atomicAdd(tb + 2, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last || threadIdx.x != 0) return;
  __threadfence();
  const unsigned int count_le_total = __ldcg(tb);
  const unsigned int nb = __ldcg(tb + 1);
  const float nx = nb ? __uint_as_float(~nb) : 3.4e38f;
  const float vk1 = (count_le_total > k + 1u || static_cast<long long>(k) + 1 >= n) ? vk : nx;
  const float d = vk1 - vk;
  const float q = frac < 0.5f ? vk + frac * d : vk1 - d * (1.f - frac);     // torch.lerp
  s_out[b] = fmaxf(q, floor_val);
}

//   x0c = clamp(x0, -s, s) / s ; mean = c1[b] x0c + c2[b] x ; out = mean + sig[b] * noise      VDDP:951, 926-933, 963
// sig[b] = (t > 0) * exp(0.5 * posterior_log_variance_clipped[t]) is prepared by the caller from the schedule.
__global__ void posterior_kernel(const float* __restrict__ x0, const float* __restrict__ x, const float* __restrict__ noise,
                                 const float* __restrict__ s, const float* __restrict__ c1, const float* __restrict__ c2,
                                 const float* __restrict__ sig, float* __restrict__ out, long long per, int B) {
  const long long total = per * B;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / per);
    float xc = x0[i];
    if (s) {   // s == NULL: clip_denoised=False
      const float sb = s[b];
      xc = fminf(fmaxf(xc, -sb), sb) / sb;
    }
    out[i] = c1[b] * xc + c2[b] * x[i] + sig[b] * noise[i];
  }
}

//   DDIM (eta = 0): out = x0 * sqrt(alpha_next) + sqrt(1 - alpha_next) * eps            VDDP:1014-1016
__global__ void axpby_kernel(const float* __restrict__ a, const float* __restrict__ b, float ca, float cb, float cc,
                             float* __restrict__ out, long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = ca * a[i] + (b ? cb * b[i] : 0.f) + cc;
}

// ------------------------------------------------------------------------------------------------
// Fused Adam (torch.optim.Adam defaults: no weight decay, no amsgrad) over a flat fp32 parameter arena, with
// the gradient un-scaling (1/world, loss scale) folded in, and the EMA update of VDDP:121-129 on request.
// ------------------------------------------------------------------------------------------------
__global__ void adam_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                float* __restrict__ ema, long long n, float lr, float beta1, float beta2, float eps, float bc1,
                                float bc2_sqrt, float grad_scale, int ema_mode, float ema_beta) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float gi = g[i] * grad_scale;
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    const float pi = p[i] - (lr / bc1) * (mi / denom);
    p[i] = pi;
    if (ema_mode == 1) ema[i] = pi;                                             // step < step_start_ema: copy
    else if (ema_mode == 2) ema[i] = ema[i] * ema_beta + (1.f - ema_beta) * pi;  // EMA update
  }
}

static int grid_for(long long n, int threads) {
  long long g = (n + threads - 1) / threads;
  const long long cap = 8LL * num_sms();
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace vmm

using namespace vmm;

extern "C" int vmm_prep_input(const float* x, const float* noise, const float* a, const float* c, const float* s, void* xin, int fmt,
                              int B, int C, int F, int H, int W, void* stream) {
  if (!x || !xin || C > 8 || (noise && !s)) return set_error(VMM_ERR_ARG, "vmm_prep_input: bad arguments");
  const long long total = static_cast<long long>(B) * F * H * (W + 6);
  prep_input_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, noise, a, c, s, static_cast<uint16_t*>(xin),
                                                                                        fmt, B, C, F, H, W);
  count_launch();
  return check_launch("vmm_prep_input");
}

extern "C" int vmm_loss(const float* pred, const float* target, float* loss_sum, void* dpred, int fmt, int B, int C, int F, int H, int W,
                        int l2, float grad_scale, void* stream) {
  if (!pred || !target || !loss_sum || C > 8) return set_error(VMM_ERR_ARG, "vmm_loss: bad arguments");
  const long long npix = static_cast<long long>(B) * F * H * W;
  loss_kernel<<<grid_for(npix, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(pred, target, loss_sum, static_cast<uint16_t*>(dpred), fmt,
                                                                                 B, C, F, H, W, l2, grad_scale);
  count_launch();
  return check_launch("vmm_loss");
}

extern "C" int vmm_cfg_x0(const float* x, const float* eps_cl, int has_null, float w, const float* sr, const float* srm1, float* x0,
                          float* eps_out, int B, int C, int F, int H, int W, void* stream) {
  if (!x || !eps_cl || !sr || !srm1 || !x0) return set_error(VMM_ERR_ARG, "vmm_cfg_x0: null pointer");
  const long long total = static_cast<long long>(B) * C * F * H * W;
  cfg_x0_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, eps_cl, has_null, w, sr, srm1, x0, eps_out, B, C, F,
                                                                                    H, W);
  count_launch();
  return check_launch("vmm_cfg_x0");
}

extern "C" size_t vmm_abs_quantile_workspace(int B) {
  return B > 0 ? static_cast<size_t>(B) * (4 * 256 + 4) * sizeof(unsigned int) : 0;
}

extern "C" int vmm_abs_quantile(const float* v, int B, long long n, long long k, float frac, float floor_val, float* s_out, void* workspace,
                                size_t workspace_bytes, void* stream) {
  if (!v || !s_out || B < 1 || n < 1 || k < 0 || k >= n) return set_error(VMM_ERR_ARG, "vmm_abs_quantile: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // small samples (or no workspace): one CTA per sample does all four passes; otherwise every pass is spread over the whole GPU
  if (!workspace || n < 65536) {
    abs_quantile_kernel<<<B, 1024, 0, st>>>(v, n, k, frac, floor_val, s_out);
    count_launch();
    return check_launch("vmm_abs_quantile");
  }
  if (n >= (1LL << 32)) return set_error(VMM_ERR_UNSUPPORTED, "vmm_abs_quantile: more than 2^32 elements per sample");
  if (workspace_bytes < vmm_abs_quantile_workspace(B)) return set_error(VMM_ERR_ARG, "vmm_abs_quantile: workspace too small");
  cudaError_t e = cudaMemsetAsync(workspace, 0, vmm_abs_quantile_workspace(B), st);
  if (e != cudaSuccess) return set_cuda_error(e, "vmm_abs_quantile: memset");
  unsigned int* hist = static_cast<unsigned int*>(workspace);
  unsigned int* tail = hist + static_cast<size_t>(B) * 4 * 256;
  // CTAs per sample: the whole GPU over the B samples, at least 2048 elements per CTA
  int gx = (num_sms() + B - 1) / B;
  const long long max_gx = (n + 2047) / 2048;
  if (gx > max_gx) gx = static_cast<int>(max_gx);
  if (gx < 1) gx = 1;
  const dim3 grid(gx, B);
  quantile_hist_kernel<0><<<grid, 512, 0, st>>>(v, n, static_cast<unsigned int>(k), hist);
  quantile_hist_kernel<1><<<grid, 512, 0, st>>>(v, n, static_cast<unsigned int>(k), hist);
  quantile_hist_kernel<2><<<grid, 512, 0, st>>>(v, n, static_cast<unsigned int>(k), hist);
  quantile_hist_kernel<3><<<grid, 512, 0, st>>>(v, n, static_cast<unsigned int>(k), hist);
  quantile_final_kernel<<<grid, 512, 0, st>>>(v, n, static_cast<unsigned int>(k), frac, floor_val, hist, tail, s_out);
  for (int i = 0; i < 5; ++i) count_launch();
  return check_launch("vmm_abs_quantile");
}

extern "C" int vmm_posterior_step(const float* x0, const float* x, const float* noise, const float* s, const float* c1, const float* c2,
                                  const float* sig, float* out, int B, long long per, void* stream) {
  if (!x0 || !x || !noise || !c1 || !c2 || !sig || !out) return set_error(VMM_ERR_ARG, "vmm_posterior_step: null pointer");
  posterior_kernel<<<grid_for(per * B, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x0, x, noise, s, c1, c2, sig, out, per, B);
  count_launch();
  return check_launch("vmm_posterior_step");
}

extern "C" int vmm_axpby(const float* a, const float* b, float ca, float cb, float cc, float* out, long long n, void* stream) {
  if (!a || !out) return set_error(VMM_ERR_ARG, "vmm_axpby: null pointer");
  axpby_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, b, ca, cb, cc, out, n);
  count_launch();
  return check_launch("vmm_axpby");
}

extern "C" int vmm_adam_ema_step(float* p, const float* g, float* m, float* v, float* ema, long long n, float lr, float beta1, float beta2,
                                 float eps, int step, float grad_scale, int ema_mode, float ema_beta, void* stream) {
  if (!p || !g || !m || !v || (ema_mode && !ema) || step < 1) return set_error(VMM_ERR_ARG, "vmm_adam_ema_step: bad arguments");
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), step);
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), step);
  adam_ema_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, g, m, v, ema, n, lr, beta1, beta2, eps,
                                                                                  static_cast<float>(bc1), static_cast<float>(sqrt(bc2)),
                                                                                  grad_scale, ema_mode, ema_beta);
  count_launch();
  return check_launch("vmm_adam_ema_step");
}

// ------------------------------------------------------------------------------------------------
// Weight repack: every 16-bit GEMM operand of the network (forward and data-gradient forms, ~2x the parameter
// count) is a fixed gather of the flat fp32 parameter arena.  dst[i] = idx[i] < 0 ? 0 : cast(src[idx[i]]).
// One launch per optimisation step instead of ~4000 slicing kernels.
// ------------------------------------------------------------------------------------------------
namespace vmm {
__global__ void gather_cast_kernel(const float* __restrict__ src, const int* __restrict__ idx, uint16_t* __restrict__ dst, long long n,
                                   int fmt) {
  const long long nvec = n / 8;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int4 a = __ldg(reinterpret_cast<const int4*>(idx) + 2 * i);
    const int4 b = __ldg(reinterpret_cast<const int4*>(idx) + 2 * i + 1);
    const int id[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = id[j] < 0 ? 0.f : __ldg(src + id[j]);
    uint4 q;
    q.x = pack2_h16(v[0], v[1], fmt);
    q.y = pack2_h16(v[2], v[3], fmt);
    q.z = pack2_h16(v[4], v[5], fmt);
    q.w = pack2_h16(v[6], v[7], fmt);
    reinterpret_cast<uint4*>(dst)[i] = q;
  }
}
}  // namespace vmm

extern "C" int vmm_gather_cast(const float* src, const int* idx, void* dst, long long n, int fmt, void* stream) {
  if (!src || !idx || !dst || (n % 8) != 0) return vmm::set_error(VMM_ERR_ARG, "vmm_gather_cast: bad arguments (n must be a multiple of 8)");
  vmm::gather_cast_kernel<<<vmm::grid_for(n / 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, idx, static_cast<uint16_t*>(dst), n, fmt);
  vmm::count_launch();
  return vmm::check_launch("vmm_gather_cast");
}
