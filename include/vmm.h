/* vmm.h -- C ABI of libvmm_sm100.so, the B200 (sm_100a) kernels behind the VideoMetamaterials hot path.
 *
 * The reference (jhbastek/VideoMetamaterials) is pure Python/PyTorch and has no FFI layer; the
 * boundary a maintainer binds is therefore NEW.  Each entry point names the
 * reference code it replaces as "VDDP:line" =
 * denoising_diffusion_pytorch/video_denoising_diffusion_pytorch.py.
 *
 * Conventions
 *   - plain C: device pointers + sizes, no torch types.  The caller owns every buffer (including
 *     workspaces); the library allocates nothing and never synchronises.
 *   - every call is asynchronous on the `stream` argument (a cudaStream_t passed as void*).
 *   - return value 0 = launched; negative = rejected (vmm_last_error() has the text; per thread).
 *   - activations are channels-last: (b, f, h, w, c) row-major, 16-bit (fmt 0 = fp16, 1 = bf16)
 *     unless stated; parameters and statistics are fp32 / fp64 as stated.
 *   - there is NO CPU implementation behind this ABI: without a CUDA device every compute entry
 *     point fails with VMM_ERR_CUDA.
 */
#ifndef VMM_H_
#define VMM_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VMM_OK 0
#define VMM_ERR_ARG -1
#define VMM_ERR_CUDA -2
#define VMM_ERR_UNSUPPORTED -3

#define VMM_FMT_F16 0
#define VMM_FMT_BF16 1

const char* vmm_last_error(void);
int vmm_abi_version(void);
/* number of kernel launches issued through this library by the calling process (for bench.py) */
uint64_t vmm_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution / linear on tcgen05 tensor cores (TMA-fed, TMEM accumulators).
 *
 *   out[pix, n] = epilogue( sum_{tap} sum_{c} A_tap[pix + (dy,dx), c] * W[n, tap.kofs + c] )
 *
 * replaces: Block.proj  nn.Conv3d(Cin,Cout,(1,3,3))                VDDP:271,278   (9 taps)
 *           ResnetBlock.res_conv / final 1x1x1 conv                VDDP:297,311,708 (1 tap)
 *           Attention.to_qkv / to_out nn.Linear                    VDDP:413,421,437,535
 *           SpatialLinearAttention.to_qkv / to_out 1x1 Conv2d      VDDP:319,325,336,377
 *           Downsample nn.Conv3d (1,4,4)/(1,2,2)                   VDDP:241  (16 taps over 4 parity views)
 *           Upsample nn.ConvTranspose3d (1,4,4)/(1,2,2)            VDDP:155  (4 output phases x 4 taps)
 *           init_conv (1,7,7) on an 8-channel padded input         VDDP:626  (7 taps of 8 px x 8 ch)
 *           and every data-gradient of the above (same kernel, transformed weights).
 *
 * A operand: up to 4 strided 4-D views (c, w, h, bf) of 16-bit activations, read by TMA with
 * out-of-range coordinates zero-filled (that IS the 'zeros' padding mode).  W: packed 16-bit
 * [N_pad, Ktot] K-major.  The GEMM-M space is the output grid (BF, OH, OW), cut into tiles of
 * TF x TH x TW = 128 positions (powers of two).
 * Epilogue (all optional): x alpha; + bias[n]; + residual[pix, n]; per-(sample, group) sum / sum-of-squares
 * for GroupNorm accumulated in fp64 (VDDP:274); store 16-bit or fp32; columns >= nsplit go to out2.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const void* ptr;    /* 16-bit elements */
  int32_t dims[4];    /* {C, W, H, BF}, innermost first */
  int64_t strides[3]; /* element strides of W, H, BF (C is contiguous); multiples of 8 */
} vmm_view4;

typedef struct {
  int32_t src;  /* index into a[] */
  int32_t dy, dx; /* offsets added to the tile origin (y, x) in that view's coordinates */
  int32_t kofs; /* first K column of this tap in W; multiple of 64 */
  int32_t c;    /* channels read from the view for this tap (rounded up to 64 with zero fill) */
} vmm_tap;

#define VMM_MAX_VIEWS 4
#define VMM_MAX_PHASES 4
#define VMM_MAX_TAPS 20

typedef struct {
  int32_t fmt; /* VMM_FMT_* of A, W, residual and 16-bit outputs */
  int32_t n_views;
  vmm_view4 a[VMM_MAX_VIEWS];
  int32_t n_phases; /* 1, or 4 for the transposed conv */
  int32_t n_taps[VMM_MAX_PHASES];
  vmm_tap taps[VMM_MAX_PHASES][VMM_MAX_TAPS];
  int32_t phase_oy[VMM_MAX_PHASES], phase_ox[VMM_MAX_PHASES]; /* output offsets of each phase */
  const void* w; /* [n_pad][ktot] 16-bit, n_pad = N rounded up to 16 */
  int32_t n, ktot;
  int32_t bf, oh, ow; /* GEMM-M grid */
  int32_t tf, th, tw; /* tile, tf*th*tw == 128 */
  /* output pixel index = ((bf * ohs) + y * sy + oy) * ows + x * sx + ox ; element offset = index * ldo + n */
  void* out;
  int64_t ldo;
  int32_t out_fp32;
  int32_t ohs, ows, sy, sx;
  void* out2; /* columns >= nsplit (multiple of 16) are written to out2[:, n - nsplit]; may be NULL */
  int64_t ldo2;
  int32_t nsplit;
  const float* bias;  /* [n] or NULL */
  const void* res;    /* 16-bit residual with the output's pixel indexing, or NULL */
  int64_t ldr;
  const void* res2;   /* residual for the columns >= nsplit (required when res and out2 are both set) */
  int64_t ldr2;
  float alpha;        /* accumulator scale applied before bias / residual; 0 means 1 */
  double* gn_stats;   /* [samples][n / gn_group][2] (sum, sum of squares), accumulated atomically; or NULL */
  int32_t gn_group;   /* channels per group */
  int32_t frames_per_sample;
  /* Rotary epilogue (to_qkv of the temporal attention, VDDP:449,456,496): output columns [0, rot_cols) are rotated in
   * interleaved pairs (2i, 2i+1) of every 32-wide head slice by the angle of the row's frame, f = (row / rot_hw) % rot_frames
   * (rows GEMM: row = position index (b, f, pixel)); columns [0, rot_qcols) use table 0 (queries: cos/sin pre-multiplied by
   * the attention scale), the others table 1 (keys).  rot: fp32 [2][rot_frames][16][2] (cos, sin) or NULL. */
  const float* rot;
  int32_t rot_frames, rot_hw, rot_cols, rot_qcols;
} vmm_cgemm_params;

int vmm_cgemm(const vmm_cgemm_params* p, void* stream);

/* ------------------------------------------------------------------------------------------
 * GroupNorm apply (+ time-embedding scale/shift + SiLU).  replaces Block.norm / scale_shift / act
 * VDDP:279-285.  x, y: [B][pix][C] 16-bit.  stats: fp64 (sum, sumsq) per (sample, group) as
 * accumulated by vmm_cgemm.  scale_shift: [B][2C] fp32 (scale | shift, ResnetBlock.mlp output
 * VDDP:304-306) or NULL.  act: 1 = SiLU, 0 = identity.  res (may be NULL) is added after the
 * activation: the identity skip of ResnetBlock (VDDP:311) when dim == dim_out.
 * The backward also accumulates dgamma/dbeta (+=) and writes d(scale_shift) [B][2C] (may be NULL).
 * dx_colsum (may be NULL): [C] fp32, += sum over samples and pixels of dx: the bias gradient of the convolution that
 * produced x (Block.proj, VDDP:277), folded into the apply pass instead of a separate column-sum launch.
 * ------------------------------------------------------------------------------------------ */
int vmm_gn_silu_fwd(const void* x, const void* res, void* y, int fmt, int B, long long pix, int C, int groups, const double* stats,
                    const float* gamma, const float* beta, const float* scale_shift, float eps, int act, void* stream);
size_t vmm_gn_silu_bwd_workspace(int B, int C, int groups);
int vmm_gn_silu_bwd(const void* x, const void* dy, void* dx, int fmt, int B, long long pix, int C, int groups,
                    const double* stats, const float* gamma, const float* beta, const float* scale_shift, float eps, int act,
                    float* dgamma, float* dbeta, float* dscale_shift, float* dx_colsum, void* workspace, size_t workspace_bytes,
                    void* stream);

/* channel LayerNorm with gain only (VDDP:245-254) over rows of C 16-bit values.
 * backward: dx = LN'(dy) (+ dres if non-NULL, the Residual skip VDDP:137); dgamma accumulated (+=). */
int vmm_ln_fwd(const void* x, void* y, int fmt, long long rows, int C, const float* gamma, float eps, float* mean_rstd,
               void* stream);
int vmm_ln_bwd(const void* x, const void* dy, const void* dres, void* dx, int fmt, long long rows, int C, const float* gamma,
               float eps, float* dgamma, void* stream);

/* ------------------------------------------------------------------------------------------
 * Attention cores.  qkv rows: [position][3 * heads * 32] 16-bit (q | k | v, each (head, 32)); out rows:
 * [position][heads * 32].  ekv: fp32 conditioning keys | values, [B][T][2 * heads * 32].
 *
 * vmm_tattn_*: temporal attention, one sequence per pixel over `frames` tokens, rotary on q and k,
 *   relative position bias [heads][frames][frames] added to the cond half and the frame half (VDDP:425-535,
 *   503-510); ekv (keys already rotated) may be NULL (init_temporal_attn VDDP:743).  rot: [frames][16][2] cos,sin.
 *   positions are ordered (b, f, pixel).
 * vmm_lattn_*: SpatialLinearAttention VDDP:331-378, positions ordered (bf, pixel); T cond tokens prepended
 *   to every frame; ctx [BF][heads][32][32] and kstat [BF][heads][32][2] are outputs kept for the backward.
 * vmm_sattn_*: quadratic spatial attention of the bottleneck VDDP:687-689: one cond token per frame
 *   (ekv row bf), no rotary, no bias; lse [BF][heads][HW] kept for the backward.
 * ------------------------------------------------------------------------------------------ */
/* pre_rotated != 0: the q / k columns of qkv already carry the rotary embedding and q the scale (vmm_cgemm rotary epilogue) */
int vmm_tattn_fwd(const void* qkv, const float* ekv, const float* bias, const float* rot, void* out, int fmt, int B, int frames,
                  int HW, int heads, float scale, int pre_rotated, void* stream);
/* The whole Residual(PreNorm(temporal Attention)) block of a 64-channel level in one kernel (VDDP:131-137, 245-264, 381-535):
 * TMA x tile -> channel LayerNorm -> to_qkv on tcgen05 (TMEM) -> rotary -> 11 x 22 attention per (pixel, head) on mma.sync ->
 * to_out on tcgen05 -> + x -> bulk tensor store.  x, out: [B][frames][HW][64] 16-bit; wqkv [768][64] / wout [64][256]: packed
 * K-major operands of vmm_cgemm; gamma [64]; ekv / bias as vmm_tattn_fwd; rot [2][frames][16][2] (table 0 carries the query scale).
 * xn_save [rows][64], qkv_save [rows][768] (q, k rotated, q scaled), ao_save [rows][256]: optional outputs for the backward
 * kernels (NULL when sampling: then qkv never reaches HBM).  frames == 11, heads == 8, C == 64, else VMM_ERR_UNSUPPORTED. */
size_t vmm_ftattn_workspace(int B);   /* bytes of `workspace`: the per-(sample, head) attention fragments a small pre-kernel prepares */
int vmm_ftattn_fwd(const void* x, void* out, const void* wqkv, const void* wout, const float* gamma, const float* ekv, const float* bias,
                   const float* rot, void* xn_save, void* qkv_save, void* ao_save, void* workspace, size_t workspace_bytes, int fmt,
                   int B, int frames, int HW, int C, int heads, float eps, void* stream);
/* resident CTAs per SM of the fused kernel on the current device (2 expected; diagnostics) */
int vmm_ftattn_ctas_per_sm(void);
/* diagnostics (12 ints): registers, static / max dynamic shared memory, device limits, occupancy at several shared-memory sizes */
int vmm_ftattn_diag(int* out);
/* The whole Residual(PreNorm(SpatialLinearAttention)) block of a 64-channel level without materialising qkv (inference form:
 * nothing is kept for a backward pass), VDDP:131-137, 245-264, 313-378: per 128-pixel tile TMA -> channel LayerNorm ->
 * K^T / V projections on tcgen05 (the column softmax of k becomes a per-thread reduction) -> tile context on tcgen05 -> online
 * combination; a small kernel merges the per-CTA partials with the T conditioning tokens; a third kernel projects q, applies
 * the row softmax, multiplies by the block-diagonal context and by to_out on tcgen05, adds bias and x and stores.
 * x, out: [BF][HW][64] 16-bit; wqkv [768][64], wout [64][256] packed operands; ekv fp32 [B][T][512]; ctx [BF][8][32][32] and
 * kstat [BF][256][2] (may be NULL) are outputs as of vmm_lattn_fwd.  HW % 128 == 0, heads == 8, C == 64. */
size_t vmm_flattn_workspace(int BF);
int vmm_flattn_fwd(const void* x, void* out, const void* wqkv, const void* wout, const float* gamma, const float* bias_out,
                   const float* ekv, int T, float* ctx, float* kstat, void* workspace, size_t workspace_bytes, int fmt, int BF,
                   int frames, int HW, int C, int heads, float scale, float eps, void* stream);
int vmm_lattn_fwd(const void* qkv, const float* ekv, int T, void* out, float* ctx, float* kstat, int fmt, int BF, int frames,
                  int HW, int heads, float scale, void* stream);
int vmm_sattn_fwd(const void* qkv, const float* ekv, void* out, float* lse, int fmt, int BF, int frames, int HW, int heads,
                  float scale, void* stream);

/* Backward of the attention cores: gradients w.r.t. the qkv rows (dqkv, same layout as qkv), the conditioning
 * keys|values (dekv, fp32, ACCUMULATED with atomics except vmm_sattn_bwd which overwrites its rows) and, for the
 * temporal attention, the relative position bias (dbias, accumulated).  vscale = 1 / (h*w) of VDDP:371. */
int vmm_tattn_bwd(const void* qkv, const float* ekv, const float* bias, const float* rot, const void* dout, void* dqkv, float* dekv,
                  float* dbias, int fmt, int B, int frames, int HW, int heads, float scale, int pre_rotated, void* stream);
int vmm_lattn_bwd(const void* qkv, const float* ekv, int T, const void* dout, const float* ctx, const float* kstat, float* dctx,
                  void* dqkv, float* dekv, int fmt, int BF, int frames, int HW, int heads, float scale, float vscale, void* stream);
int vmm_sattn_bwd(const void* qkv, const float* ekv, const void* aout, const void* dout, const float* lse, void* dqkv, float* dekv,
                  int fmt, int BF, int HW, int heads, float scale, void* stream);

/* ------------------------------------------------------------------------------------------
 * Around the network (fp32 tensors in the reference (B, C, F, H, W) layout unless stated).
 * vmm_prep_input : value = a[b]*x + c[b] + s[b]*noise -> 16-bit [BF][H][W+6][8] init_conv operand
 *                  (q_sample VDDP:1036-1042, normalize_img VDDP:1109); the buffer needs 8 trailing zero elements.
 * vmm_loss       : F.l1_loss / F.mse_loss (VDDP:1053-1056) of pred (fp32 channels-last [pix][C]) against
 *                  target; adds the mean to *loss_sum and writes dpred [pix][8] 16-bit (x grad_scale).
 * vmm_cfg_x0     : eps = null + (cond - null) * w (VDDP:728), x0 = sr*x - srm1*eps (VDDP:920-924);
 *                  eps_cl is fp32 channels-last [(2)B][F*H*W][C].
 * vmm_abs_quantile: s[b] = max(quantile(|v_b|), floor) with torch.quantile's linear interpolation between the
 *                  k-th and (k+1)-th order statistics (VDDP:941-947); exact 4-pass 8-bit radix select.  With a workspace of
 *                  vmm_abs_quantile_workspace(B) bytes (zeroed by the call, in stream order) and n >= 65536 every pass is
 *                  one launch over the whole GPU (4 histogram launches + 1 that counts / finds the neighbour and
 *                  writes s); with workspace == NULL or small n one CTA per sample does everything in one launch.
 * vmm_posterior_step: clamp(x0,-s,s)/s (skipped when s == NULL), posterior mean (VDDP:926-933), + sig[b]*noise (VDDP:963).
 * vmm_axpby      : out = ca*a + cb*b + cc  (DDIM update VDDP:1014-1016, unnormalize_img VDDP:1112).
 * vmm_adam_ema_step: torch.optim.Adam (VDDP:1481) over a flat arena + EMA (VDDP:121-129); ema_mode 0 none,
 *                  1 copy (step < step_start_ema, VDDP:1501-1503), 2 update.
 * ------------------------------------------------------------------------------------------ */
int vmm_prep_input(const float* x, const float* noise, const float* a, const float* c, const float* s, void* xin, int fmt, int B,
                   int C, int F, int H, int W, void* stream);
int vmm_loss(const float* pred, const float* target, float* loss_sum, void* dpred, int fmt, int B, int C, int F, int H, int W,
             int l2, float grad_scale, void* stream);
int vmm_cfg_x0(const float* x, const float* eps_cl, int has_null, float w, const float* sr, const float* srm1, float* x0,
               float* eps_out, int B, int C, int F, int H, int W, void* stream);
size_t vmm_abs_quantile_workspace(int B);
int vmm_abs_quantile(const float* v, int B, long long n, long long k, float frac, float floor_val, float* s_out, void* workspace,
                     size_t workspace_bytes, void* stream);
int vmm_posterior_step(const float* x0, const float* x, const float* noise, const float* s, const float* c1, const float* c2,
                       const float* sig, float* out, int B, long long per, void* stream);
int vmm_axpby(const float* a, const float* b, float ca, float cb, float cc, float* out, long long n, void* stream);
/* dst[i] = idx[i] < 0 ? 0 : (16-bit) src[idx[i]]: rebuilds every packed GEMM operand from the fp32 parameter arena. */
int vmm_gather_cast(const float* src, const int* idx, void* dst, long long n, int fmt, void* stream);
int vmm_adam_ema_step(float* p, const float* g, float* m, float* v, float* ema, long long n, float lr, float beta1, float beta2,
                      float eps, int step, float grad_scale, int ema_mode, float ema_beta, void* stream);

/* ------------------------------------------------------------------------------------------
 * Conditioning / time path of Unet3D.forward in one forward kernel and two backward kernels (fp32, CUDA cores):
 *   SinusoidalPosEmb + time_mlp (VDDP:139-151, 637-642, 745), sign_emb tokens (VDDP:653, 753-755), cond_token_to_hidden
 *   (VDDP:656-661, 757-759), classifier-free-guidance null tokens (VDDP:772-784), t + hidden (VDDP:786-788), every
 *   ResnetBlock.mlp (VDDP:290-293, 304-306), every to_k / to_v on the tokens with the rotary embedding of the temporal
 *   blocks' keys (VDDP:349-353, 457-474), the relative position bias (VDDP:70-108, 741) and the rotary cos / sin tables.
 * Parameters are read from / gradients accumulated into flat fp32 arenas at the given element offsets.  `out` (forward) is one
 * flat fp32 buffer: per ResnetBlock (B, 2C) at res_out[j]; per attention block keys|values (B, T, 512) at att_out[a]; the bias
 * [heads][frames][frames] at bias_out; rotary tables [2][frames][16][2] at rot_out (table 0 x dim_head^-1/2).  For
 * vmm_cond_bwd `out` holds the gradients w.r.t. those outputs in the same layout; `ws` (vmm_cond_workspace bytes) carries
 * the saved intermediates from the forward call.  B <= 32, T <= 16, td <= 256, heads == 8.
 * ------------------------------------------------------------------------------------------ */
#define VMM_COND_MAX_BLOCKS 24
typedef struct {
  int32_t B, T, dim, td, heads, frames, n_res, n_attn;
  const int64_t* time;             /* (B) timesteps */
  const float* cond;               /* (B, T) */
  const unsigned char* null_mask;  /* (B) 1 = conditioning dropped */
  const float* param;              /* parameter arena */
  float* grad;                     /* gradient arena (vmm_cond_bwd) */
  const float* freqs;              /* rotary frequencies [16] */
  const int32_t* buckets;          /* relative position bucket of (i, j), [frames][frames] */
  int64_t o_w1, o_b1, o_w2, o_b2;  /* time_mlp.1 / time_mlp.3 */
  int64_t o_wse, o_bse;            /* sign_emb */
  int64_t o_lng, o_lnb, o_w3, o_b3, o_w4, o_b4;   /* cond_token_to_hidden.0 / .1 / .3 */
  int64_t o_ntok, o_nhid, o_table; /* null_text_token, null_text_hidden, time_rel_pos_bias table [32][heads] */
  int64_t res_w[VMM_COND_MAX_BLOCKS], res_b[VMM_COND_MAX_BLOCKS], res_out[VMM_COND_MAX_BLOCKS];
  int32_t res_c2[VMM_COND_MAX_BLOCKS];
  int64_t att_wk[VMM_COND_MAX_BLOCKS], att_wv[VMM_COND_MAX_BLOCKS], att_out[VMM_COND_MAX_BLOCKS];
  int32_t att_temporal[VMM_COND_MAX_BLOCKS];
  int64_t bias_out, rot_out;
  float* out;
  float* ws;
} vmm_cond_params;
size_t vmm_cond_workspace(int B, int T, int dim, int td);
int vmm_cond_fwd(const vmm_cond_params* p, void* stream);
int vmm_cond_bwd(const vmm_cond_params* p, void* stream);

/* ------------------------------------------------------------------------------------------
 * Weight gradients on tcgen05 tensor cores:  dW[n, tap, c] += sum_pix dY[pix, n] * X[pix + d_tap, c].
 * The gradient of every layer vmm_cgemm runs forward (VDDP:271,297,319,325,413,421,241,155,626,708).
 * a[]: views of the output gradient (one, or the 4 parity views for the transposed conv);
 * b[]: views of the layer input (one or two concat sources, or 4 parity views for the strided conv).
 * Partial sums are added atomically into the fp32 master-layout gradient:
 *   dw[tap.wofs + n * s_m + (c % cmod) * s_c + (c / cmod) * s_c2]      (cmod <= 0: plain c * s_c)
 * vmm_colsum: out[n] += sum_rows x[row][n]  (bias gradients).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t a_src, b_src; /* view indices */
  int32_t dy, dx;       /* shift of the b view relative to the a view */
  int32_t c;            /* channels of the b view used by this tap */
  int64_t wofs;         /* element offset of this tap in dw */
} vmm_wgrad_tap;

typedef struct {
  int32_t fmt;
  int32_t n_a_views, n_b_views;
  vmm_view4 a[VMM_MAX_VIEWS];
  vmm_view4 b[VMM_MAX_VIEWS];
  int32_t n_taps;
  vmm_wgrad_tap taps[VMM_MAX_TAPS];
  int32_t n;            /* channels of dY = rows of dW */
  int32_t bf, oh, ow;   /* pixel grid of the a views (the reduction axis) */
  int32_t tf, th, tw;   /* pixel tile, tf*th*tw == 128 */
  float* dw;
  int64_t s_m, s_c, s_c2;
  int32_t cmod, c_valid, k_valid;
} vmm_wgrad_params;

int vmm_wgrad(const vmm_wgrad_params* p, void* stream);
/* Both consumers of an attention block's d(qkv) rows at a 64-channel level in one pass over them (autograd of to_qkv, VDDP:437 / 336):
 *   dxn[row][c] = sum_k dqkv[row][k] * W[k][c]   (16-bit, [rows][64]);      dw[k][c] += sum_row dqkv[row][k] * xn[row][c]   (fp32 [768][64])
 * dqkv: [rows][768] 16-bit; xn: [rows][64] 16-bit (the to_qkv input); wd: W^T packed K-major [64][768] 16-bit (the data-gradient pack
 * vmm_cgemm takes for the same product).  Equivalent to one vmm_cgemm + one vmm_wgrad launch that each stream dqkv from HBM. */
int vmm_qkv_bwd(const void* dqkv, const void* xn, const void* wd, void* dxn, float* dw, long long rows, int fmt, void* stream);
/* The same with the PreNorm + Residual in front of to_qkv folded into the epilogue (VDDP:245-264, 131-137; what vmm_ln_bwd computes from
 * the stored dxn rows): dx[row] = LayerNorm'(x[row]; gamma)(dxn[row]) + dres[row] (16-bit [rows][64]), dgamma[c] += sum_row dxn[row][c] * xhat[row][c].
 * x: the block input, dres: the gradient arriving over the residual connection.  dxn never reaches HBM. */
int vmm_qkv_ln_bwd(const void* dqkv, const void* xn, const void* wd, const void* x, const void* dres, const float* gamma, float eps, void* dx,
                   float* dw, float* dgamma, long long rows, int fmt, void* stream);
int vmm_colsum(const void* x, long long rows, int n, long long ld, int fmt, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Dataset on the device: GIF decode + the Dataset's normalisation.
 *
 * replaces: gif_to_tensor / seek_all_images  (PIL Image.open + seek + convert('L') + ToTensor)   VDDP:1076-1106
 *           Dataset.__getitem__ (unnorm with the sample's range, void pixels -> 0, global normalise) VDDP:1302-1397
 *
 * vmm_gif_scan (HOST, no device work): walks one GIF file held in host memory and fills one vmm_gif_frame per image
 * (offsets are relative to the start of that file).  Returns the number of frames (it keeps counting beyond
 * max_frames without writing), or a negative code: VMM_ERR_ARG for a malformed file, VMM_ERR_UNSUPPORTED for a
 * feature the device compositor does not reproduce (disposal method 3, a transparent FIRST frame, frames that
 * leave the logical screen, more than 2^19 pixels per frame).
 * flags & VMM_GIF_PIL_COMPAT: Pillow (12.2 here, the decoder behind the reference's gif_to_tensor) opens a file whose
 * first frame has the plain grey ramp as its palette in mode 'L' and decodes the FIRST later frame that brings a real
 * palette as raw indices, ignoring that palette (it does not round-trip such files it wrote itself: measured,
 * tests/test_cpu_gif.py); with the flag that frame gets the grey ramp here too, so the planes equal what the reference
 * reads.  Without it every frame is mapped through its own palette (the image the file encodes).
 *
 * vmm_gif_decode (DEVICE): `files` = the bytes of n_files GIF files back to back, file_ofs[i] = start of file i
 * (n_files + 1 entries), frame_begin[i] .. frame_begin[i+1] = its rows in `frames` (n_files + 1 entries,
 * frame_begin[n_files] = n_frames_total; the tables vmm_gif_scan produced, copied to the device).
 * Kernel 1 (one warp per frame) walks the sub-block chain and decodes the LZW stream into `index_ws`
 * (vmm_gif_frame.px_ofs = the frame's offset in it; multiples of 16 get 16-byte stores).  max_frame_px = the largest w*h in
 * `frames` (the caller has the table on the host): up to 12 288 pixels a frame's index stream is built in shared memory and
 * written out once; <= 0 or larger frames decode in global memory.  Kernel 2 (one CTA per
 * file) composites the frames in order on the logical screen (frame rectangle, interlace, transparency, disposal
 * 0 / 1 / 2) and writes out[file][frame][H][W] 8-bit luminance = ITU-R 601-2 luma of the palette colour, which is
 * what PIL's convert('L') returns for every frame of such a file.  Every file has the logical screen H x W and at
 * most `frames_per_file` frames; frames a file does not have are written as zeros.  err[0] is incremented for
 * every frame whose LZW stream was short or invalid (the caller zeroes it).
 *
 * vmm_dataset_items (DEVICE): out[i][ch][f][h][w] fp32 for the n samples in `index`, from the decoded planes
 * u8[sample][plane][f][h][w]: t = u8/255; if the channel has a range: t = t*(smax-smin) + smin (two roundings, the
 * sample's range as fp32), 0 where the topology plane is 0, then (t - gmin) / (gmax - gmin).  Bit-identical to the
 * reference's fp32 tensor arithmetic (every operation correctly rounded, no contraction into FMAs).
 *   ch_plane[c]   plane read by output channel c            ch_has_range[c]  0 = pass-through (the topology itself)
 *   sample_rng    [n_samples][n_ch][2] fp32 = {smin, smax - smin} per sample and channel
 *   global_rng    [n_ch][2] fp32 = {gmin, gmax - gmin}
 *   sample_frames [n_samples] frames the sample's files really hold (NULL: `frames`)
 *   frames_out > the sample's frames: the extra frames are zero (cast_num_frames, VDDP:1114-1124); fewer: truncated.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  uint32_t data_ofs;     /* first data sub-block (the byte after the LZW minimum code size) */
  uint32_t pal_ofs;      /* RGB palette used by this frame (local, else global); 0xffffffff = none (grey ramp) */
  uint32_t px_ofs;       /* offset of this frame's index stream in the decode workspace (filled by the caller) */
  uint16_t x, y, w, h;   /* frame rectangle on the logical screen */
  uint16_t pal_size;     /* number of palette entries */
  uint8_t min_code;      /* LZW minimum code size */
  uint8_t interlace;
  uint8_t disposal;      /* 0 / 1 keep, 2 restore to background (as PIL applies it: sticky when a frame leaves it unspecified) */
  uint8_t has_transp;
  uint8_t transp;        /* transparent colour index */
  uint8_t background;    /* logical screen background colour index */
  uint32_t reserved;     /* sizeof(vmm_gif_frame) == 32 */
} vmm_gif_frame;

typedef struct {
  uint16_t width, height;  /* logical screen */
  int32_t n_frames;
} vmm_gif_info;

#define VMM_GIF_PIL_COMPAT 1 /* reproduce Pillow's handling of a palette that arrives while the image is still in mode 'L' (above) */
int vmm_gif_scan(const uint8_t* file, size_t nbytes, int flags, vmm_gif_info* info, vmm_gif_frame* frames, int max_frames);
int vmm_gif_decode(const uint8_t* files, const uint64_t* file_ofs, const int32_t* frame_begin, const vmm_gif_frame* frames,
                   int n_files, int n_frames_total, int frames_per_file, int H, int W, int max_frame_px, uint8_t* index_ws, uint8_t* out,
                   int32_t* err, void* stream);
int vmm_dataset_items(const uint8_t* u8, const int64_t* index, int n, int n_planes, int topo_plane, int n_ch,
                      const int32_t* ch_plane, const int32_t* ch_has_range, const float* sample_rng, const float* global_rng,
                      const int32_t* sample_frames, int frames, int frames_out, int hw, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VMM_H_ */
