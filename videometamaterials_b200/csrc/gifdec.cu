// Dataset on the device: GIF container scan (host), LZW decode + frame compositing and the
// Dataset's per-item normalisation (device).  Byte / integer work and correctly rounded fp32 arithmetic: results are bit-identical
// to PIL's decode + convert('L') (VDDP:1076-1106) and to Dataset.__getitem__ (VDDP:1302-1397).
//
//   vmm_gif_scan       host: one pass over the container (extensions, image descriptors, sub-block chains) -> frame table
//   gif_lzw_kernel     one warp per frame.  The dictionary is never materialised as strings: the string of the entry that is added
//                      after code i is (string of code i-1) + (first byte of string of code i), and those bytes are CONTIGUOUS in
//                      the output stream, so an entry is (offset, length) into the output written so far (one 32-bit word in shared
//                      memory, 16 KB per warp) and decoding a code is a lane-parallel copy inside the output buffer.  All lanes run
//                      the bit reader redundantly (uniform control flow, broadcast loads).
//   gif_compose_kernel one CTA per file: frames in order on the logical screen (rectangle, interlace, transparency, disposal 0/1/2),
//                      palette colour -> ITU-R 601-2 luma as PIL's convert('L') computes it ((19595 R + 38470 G + 7471 B + 32768) >> 16)
//   dataset_items_kernel  gather + u8/255 -> sample range -> void pixels -> global range, every operation separately rounded
#include <stdlib.h>

#include "common.cuh"

namespace vmm {

static_assert(sizeof(vmm_gif_frame) == 32, "vmm_gif_frame layout (mirrored by ctypes / numpy on the host side)");
static constexpr uint32_t GIF_NO_PALETTE = 0xffffffffu;
static constexpr int GIF_MAX_FRAME_PX = 1 << 19;   // offset field of a dictionary word
static constexpr int GIF_WARPS = 8;                // warps (= frames in flight) per CTA: 8 x 16 KB of dictionary
static constexpr int GIF_SMEM_PX = 12288;          // frames up to this many pixels (96 x 96 = 9216) are decoded into shared memory

// ----------------------------------------------------------------------------------------------------------------------------
// LZW
// ----------------------------------------------------------------------------------------------------------------------------
struct GifByteReader {
  const uint8_t* src;
  uint32_t p, end, rem;
  bool ended;
  __device__ __forceinline__ int next() {
    if (ended) return -1;
    if (rem == 0) {
      if (p >= end) { ended = true; return -1; }
      rem = src[p++];
      if (rem == 0) { ended = true; return -1; }
    }
    if (p >= end) { ended = true; return -1; }
    rem--;
    return src[p++];
  }
};

// SMEM_OUT: the index stream of the frame is built in shared memory (every decoded string is a copy from earlier output: a global-memory
// round trip per code otherwise, ~300 clocks where shared memory takes ~30) and written out once with 16-byte stores.
template <bool SMEM_OUT>
__global__ void __launch_bounds__(GIF_WARPS * 32) gif_lzw_kernel(const uint8_t* __restrict__ files, const uint64_t* __restrict__ file_ofs,
                                                                 const int32_t* __restrict__ frame_begin, const vmm_gif_frame* __restrict__ frames,
                                                                 int n_files, int n_frames_total, uint8_t* ws, int32_t* err) {
  extern __shared__ uint32_t gif_dict_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t* dict = gif_dict_smem + warp * 4096;
  uint8_t* sm_out = reinterpret_cast<uint8_t*>(gif_dict_smem + GIF_WARPS * 4096) + warp * GIF_SMEM_PX;
  const int warps_total = gridDim.x * GIF_WARPS;
  for (int fi = blockIdx.x * GIF_WARPS + warp; fi < n_frames_total; fi += warps_total) {
    // file of this frame: last i with frame_begin[i] <= fi
    int lo = 0, hi = n_files;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (frame_begin[mid] <= fi) lo = mid; else hi = mid;
    }
    const vmm_gif_frame fr = frames[fi];
    GifByteReader rd;
    rd.src = files + file_ofs[lo];
    rd.end = static_cast<uint32_t>(file_ofs[lo + 1] - file_ofs[lo]);
    rd.p = fr.data_ofs;
    rd.rem = 0;
    rd.ended = false;
    uint8_t* out = SMEM_OUT ? sm_out : ws + fr.px_ofs;
    const uint32_t npx = static_cast<uint32_t>(fr.w) * fr.h;
    const uint32_t m = fr.min_code, clear = 1u << m, eoi = clear + 1;
    uint32_t next = clear + 2, size = m + 1, bitbuf = 0, nbits = 0, pos = 0, prev_off = 0, prev_len = 0;
    bool prev_valid = false, bad = false;
    while (pos < npx) {
      while (nbits < size) {
        const int b = rd.next();
        if (b < 0) { bad = true; break; }
        bitbuf |= static_cast<uint32_t>(b) << nbits;
        nbits += 8;
      }
      if (bad) break;
      const uint32_t code = bitbuf & ((1u << size) - 1);
      bitbuf >>= size;
      nbits -= size;
      if (code == clear) { next = clear + 2; size = m + 1; prev_valid = false; continue; }
      if (code == eoi) break;
      uint32_t off = 0, len = 1;
      bool literal = false;
      if (code < clear) {
        literal = true;
      } else if (!prev_valid) {
        bad = true; break;
      } else if (code < next) {
        const uint32_t e = dict[code];
        off = e & (GIF_MAX_FRAME_PX - 1);
        len = e >> 19;
      } else if (code == next) {      // the entry being defined by this very code: previous string + its own first byte
        off = prev_off;
        len = prev_len + 1;
      } else {
        bad = true; break;
      }
      const uint32_t n = min(len, npx - pos);
      if (literal) {
        if (lane == 0) out[pos] = static_cast<uint8_t>(code);
      } else {
        for (uint32_t j = lane; j < n; j += 32) {
          uint32_t s = off + j;
          if (s >= pos) s = s - pos + off;       // only the last byte of a self-referencing entry
          out[pos + j] = out[s];
        }
      }
      if (prev_valid && next < 4096) {
        if (lane == 0) dict[next] = prev_off | ((prev_len + 1) << 19);
        next++;
        if (next == (1u << size) && size < 12) size++;
      }
      __syncwarp();                               // bytes and dictionary word visible to every lane before the next code
      prev_off = pos;
      prev_len = len;
      prev_valid = true;
      pos += n;
    }
    if (pos < npx) {
      for (uint32_t j = pos + lane; j < npx; j += 32) out[j] = 0;
      if (lane == 0) atomicAdd(err, 1);
    }
    __syncwarp();
    if (SMEM_OUT) {
      uint8_t* dst = ws + fr.px_ofs;
      if ((fr.px_ofs & 15u) == 0) {
        const uint32_t n16 = npx >> 4;
        for (uint32_t j = lane; j < n16; j += 32) reinterpret_cast<uint4*>(dst)[j] = reinterpret_cast<const uint4*>(sm_out)[j];
        for (uint32_t j = (n16 << 4) + lane; j < npx; j += 32) dst[j] = sm_out[j];
      } else {
        for (uint32_t j = lane; j < npx; j += 32) dst[j] = sm_out[j];
      }
      __syncwarp();
    }
  }
}

// ----------------------------------------------------------------------------------------------------------------------------
// compositing
// ----------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t gif_luma(const uint8_t* file, uint32_t pal_ofs, uint32_t pal_size, uint32_t idx) {
  if (pal_ofs == GIF_NO_PALETTE || idx >= pal_size) return idx;
  const uint8_t* c = file + pal_ofs + 3 * idx;
  return (19595u * c[0] + 38470u * c[1] + 7471u * c[2] + 0x8000u) >> 16;
}

// stream row of destination row y (relative to the frame) in an interlaced frame of height h
__device__ __forceinline__ uint32_t gif_interlace_src_row(uint32_t y, uint32_t h) {
  const uint32_t n1 = (h + 7) >> 3, n2 = (h + 3) >> 3, n3 = (h + 1) >> 2;
  if ((y & 7) == 0) return y >> 3;
  if ((y & 7) == 4) return n1 + (y >> 3);
  if ((y & 3) == 2) return n1 + n2 + (y >> 2);
  return n1 + n2 + n3 + (y >> 1);
}

__global__ void __launch_bounds__(256) gif_compose_kernel(const uint8_t* __restrict__ files, const uint64_t* __restrict__ file_ofs,
                                                          const int32_t* __restrict__ frame_begin, const vmm_gif_frame* __restrict__ frames,
                                                          int frames_per_file, int H, int W, const uint8_t* __restrict__ ws,
                                                          uint8_t* __restrict__ out) {
  const int f = blockIdx.x;
  const uint8_t* file = files + file_ofs[f];
  const int fb = frame_begin[f];
  const int nf = min(frame_begin[f + 1] - fb, frames_per_file);
  const uint32_t npx = static_cast<uint32_t>(H) * W;
  uint8_t* canvas = out + static_cast<size_t>(f) * frames_per_file * npx;
  for (int k = 0; k < frames_per_file; ++k) {
    uint8_t* dst = canvas + static_cast<size_t>(k) * npx;
    if (k >= nf) {
      for (uint32_t px = threadIdx.x; px < npx; px += blockDim.x) dst[px] = 0;
      continue;
    }
    const vmm_gif_frame fr = frames[fb + k];
    const uint8_t* prev = k > 0 ? dst - npx : nullptr;
    // what the previous frame leaves behind: its rectangle is refilled when its disposal method is 2
    bool refill = false;
    uint32_t rx = 0, ry = 0, rw = 0, rh = 0, fill = 0;
    if (k > 0) {
      const vmm_gif_frame pf = frames[fb + k - 1];
      if (pf.disposal == 2) {
        refill = true;
        rx = pf.x; ry = pf.y; rw = pf.w; rh = pf.h;
        uint32_t color = pf.has_transp ? pf.transp : pf.background;
        if (pf.pal_ofs != GIF_NO_PALETTE && color >= pf.pal_size) color = 0;
        fill = gif_luma(file, pf.pal_ofs, pf.pal_size, color);
      }
    }
    const uint32_t base0 = gif_luma(file, fr.pal_ofs, fr.pal_size, 0);    // untouched screen of the first frame: colour index 0
    const uint8_t* idx_stream = ws + fr.px_ofs;
    for (uint32_t px = threadIdx.x; px < npx; px += blockDim.x) {
      const uint32_t y = px / W, x = px - y * W;
      uint32_t v = k > 0 ? prev[px] : base0;
      if (refill && x - rx < rw && y - ry < rh) v = fill;
      const uint32_t lx = x - fr.x, ly = y - fr.y;       // unsigned: a pixel left of / above the rectangle wraps to a huge value
      if (lx < fr.w && ly < fr.h) {
        const uint32_t srow = fr.interlace ? gif_interlace_src_row(ly, fr.h) : ly;
        const uint32_t idx = idx_stream[srow * fr.w + lx];
        if (!(fr.has_transp && idx == fr.transp)) v = gif_luma(file, fr.pal_ofs, fr.pal_size, idx);
      }
      dst[px] = static_cast<uint8_t>(v);
    }
    __syncthreads();          // frame k complete before frame k + 1 reads it
  }
}

// ----------------------------------------------------------------------------------------------------------------------------
// Dataset.__getitem__ for a batch of sample indices
// ----------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dataset_items_kernel(const uint8_t* __restrict__ u8, const int64_t* __restrict__ index, int n_planes,
                                                            int topo_plane, int n_ch, const int32_t* __restrict__ ch_plane,
                                                            const int32_t* __restrict__ ch_has_range, const float* __restrict__ sample_rng,
                                                            const float* __restrict__ global_rng, const int32_t* __restrict__ sample_frames,
                                                            int frames, int frames_out, int hw, float* __restrict__ out) {
  // grid: (chunks of a plane, n * n_ch)
  const int item = blockIdx.y / n_ch, ch = blockIdx.y - item * n_ch;
  const int64_t s = index[item];
  const int plane = ch_plane[ch];
  const bool ranged = ch_has_range[ch] != 0;
  const float smin = sample_rng[(s * n_ch + ch) * 2 + 0], sspan = sample_rng[(s * n_ch + ch) * 2 + 1];
  const float gmin = global_rng[ch * 2 + 0], gspan = global_rng[ch * 2 + 1];
  const int valid_frames = min(sample_frames ? sample_frames[s] : frames, min(frames, frames_out));
  const size_t plane_elems = static_cast<size_t>(frames) * hw;
  const uint8_t* src = u8 + (static_cast<size_t>(s) * n_planes + plane) * plane_elems;
  const uint8_t* topo = u8 + (static_cast<size_t>(s) * n_planes + topo_plane) * plane_elems;
  float* dst = out + (static_cast<size_t>(item) * n_ch + ch) * frames_out * hw;
  const size_t total = static_cast<size_t>(frames_out) * hw, live = static_cast<size_t>(valid_frames) * hw;
  for (size_t e = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; e < total; e += static_cast<size_t>(gridDim.x) * blockDim.x * 4) {
    float r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float t = 0.f;
      if (e + j < live) {
        t = __fdiv_rn(static_cast<float>(src[e + j]), 255.f);          // ToTensor
        if (ranged) {
          t = __fadd_rn(__fmul_rn(t, sspan), smin);                     // unnorm: arr * (max - min) + min
          if (topo[e + j] == 0) t = 0.f;                                // true zero of the field in the void
          t = __fdiv_rn(__fsub_rn(t, gmin), gspan);                     // normalize: (arr - min) / (max - min)
        }
      }
      r[j] = t;
    }
    if (e + 3 < total) {
      *reinterpret_cast<float4*>(dst + e) = make_float4(r[0], r[1], r[2], r[3]);
    } else {
      for (int j = 0; j < 4 && e + j < total; ++j) dst[e + j] = r[j];
    }
  }
}

}  // namespace vmm

// ----------------------------------------------------------------------------------------------------------------------------
// C ABI
// ----------------------------------------------------------------------------------------------------------------------------
// PIL's GifImagePlugin._is_palette_needed: a palette that is exactly the grey ramp (entry i = (i, i, i)) is dropped
static bool gif_palette_needed(const uint8_t* pal, uint32_t entries) {
  for (uint32_t i = 0; i < entries; ++i)
    if (pal[3 * i] != i || pal[3 * i + 1] != i || pal[3 * i + 2] != i) return true;
  return false;
}

extern "C" int vmm_gif_scan(const uint8_t* f, size_t n, int flags, vmm_gif_info* info, vmm_gif_frame* frames, int max_frames) {
  using namespace vmm;
  if (!f || !info) return set_error(VMM_ERR_ARG, "vmm_gif_scan: null argument");
  if (n < 13 || (memcmp(f, "GIF87a", 6) != 0 && memcmp(f, "GIF89a", 6) != 0)) return set_error(VMM_ERR_ARG, "vmm_gif_scan: not a GIF file");
  if (n >= 0xffffffffull) return set_error(VMM_ERR_ARG, "vmm_gif_scan: file larger than 4 GiB");
  const uint32_t W = f[6] | (f[7] << 8), H = f[8] | (f[9] << 8);
  const uint8_t lflags = f[10], background = f[11];
  size_t p = 13;
  uint32_t gpal_ofs = GIF_NO_PALETTE, gpal_size = 0;
  if (lflags & 0x80) {
    gpal_size = 1u << ((lflags & 7) + 1);
    gpal_ofs = 13;
    p += 3 * static_cast<size_t>(gpal_size);
    if (p > n) return set_error(VMM_ERR_ARG, "vmm_gif_scan: truncated global colour table");
  }
  int count = 0;
  uint8_t sticky_disposal = 0, has_t = 0, t_index = 0;
  // VMM_GIF_PIL_COMPAT: PIL opens a file whose first frame needs no palette in mode 'L' and stays there until a frame brings a palette
  // it needs; THAT frame is decoded into the 'L' canvas as raw indices, its palette ignored (GifImagePlugin.load_prepare only builds a
  // paletted frame image in the RGB modes), after which the image is 'P' / 'RGB' and palettes are honoured.  Reproduced by giving that one
  // frame the grey ramp.
  const bool pil_compat = (flags & VMM_GIF_PIL_COMPAT) != 0;
  bool pil_mode_l = false;
  auto skip_blocks = [&](size_t& q) -> bool {     // false when the chain runs off the file
    while (true) {
      if (q >= n) return false;
      const uint8_t sz = f[q++];
      if (sz == 0) return true;
      q += sz;
    }
  };
  while (p < n) {
    const uint8_t b = f[p++];
    if (b == 0x3B) break;                          // trailer
    if (b == 0x21) {                               // extension
      if (p >= n) break;
      const uint8_t label = f[p++];
      if (label == 0xF9 && p < n && f[p] >= 4 && p + 4 < n) {     // graphic control extension: applies to the next image
        const uint8_t gflags = f[p + 1];
        if (gflags & 1) { has_t = 1; t_index = f[p + 4]; }
        const uint8_t disp = (gflags >> 2) & 7;
        if (disp) sticky_disposal = disp;          // PIL keeps the last specified method when a frame leaves it at 0
      }
      if (!skip_blocks(p)) break;
      continue;
    }
    if (b != 0x2C) continue;                       // PIL skips unknown bytes between blocks
    if (p + 9 > n) return set_error(VMM_ERR_ARG, "vmm_gif_scan: truncated image descriptor");
    vmm_gif_frame fr;
    memset(&fr, 0, sizeof(fr));
    fr.x = f[p] | (f[p + 1] << 8);
    fr.y = f[p + 2] | (f[p + 3] << 8);
    fr.w = f[p + 4] | (f[p + 5] << 8);
    fr.h = f[p + 6] | (f[p + 7] << 8);
    const uint8_t iflags = f[p + 8];
    p += 9;
    fr.interlace = (iflags & 0x40) ? 1 : 0;
    fr.pal_ofs = gpal_ofs;
    fr.pal_size = static_cast<uint16_t>(gpal_size);
    if (iflags & 0x80) {
      const uint32_t lsize = 1u << ((iflags & 7) + 1);
      fr.pal_ofs = static_cast<uint32_t>(p);
      fr.pal_size = static_cast<uint16_t>(lsize);
      p += 3 * static_cast<size_t>(lsize);
    }
    if (p > n) return set_error(VMM_ERR_ARG, "vmm_gif_scan: truncated local colour table");
    if (pil_compat) {
      const bool needed = fr.pal_ofs != GIF_NO_PALETTE && gif_palette_needed(f + fr.pal_ofs, fr.pal_size);
      if (count == 0) {
        pil_mode_l = !needed;
      } else if (pil_mode_l && needed) {
        fr.pal_ofs = GIF_NO_PALETTE;
        fr.pal_size = 0;
        pil_mode_l = false;
      }
    }
    if (p >= n) return set_error(VMM_ERR_ARG, "vmm_gif_scan: truncated image data");
    fr.min_code = f[p++];
    fr.data_ofs = static_cast<uint32_t>(p);
    fr.disposal = sticky_disposal;
    fr.has_transp = has_t;
    fr.transp = t_index;
    fr.background = background;
    has_t = 0;
    const bool chain_ok = skip_blocks(p);
    if (fr.min_code < 1 || fr.min_code > 8) return set_error(VMM_ERR_ARG, "vmm_gif_scan: LZW minimum code size outside 1..8");
    if (fr.w == 0 || fr.h == 0 || static_cast<uint32_t>(fr.x) + fr.w > W || static_cast<uint32_t>(fr.y) + fr.h > H)
      return set_error(VMM_ERR_UNSUPPORTED, "vmm_gif_scan: frame rectangle empty or outside the logical screen");
    if (static_cast<uint32_t>(fr.w) * fr.h > static_cast<uint32_t>(GIF_MAX_FRAME_PX))
      return set_error(VMM_ERR_UNSUPPORTED, "vmm_gif_scan: more than 2^19 pixels in one frame");
    if (fr.disposal >= 3) return set_error(VMM_ERR_UNSUPPORTED, "vmm_gif_scan: disposal method 3 (restore to previous) is not composited on the device");
    if (count == 0 && fr.has_transp) return set_error(VMM_ERR_UNSUPPORTED, "vmm_gif_scan: transparent first frame");
    if (frames && count < max_frames) frames[count] = fr;
    count++;
    if (!chain_ok) break;
  }
  if (count == 0) return set_error(VMM_ERR_ARG, "vmm_gif_scan: no image in the file");
  info->width = static_cast<uint16_t>(W);
  info->height = static_cast<uint16_t>(H);
  info->n_frames = count;
  return count;
}

extern "C" int vmm_gif_decode(const uint8_t* files, const uint64_t* file_ofs, const int32_t* frame_begin, const vmm_gif_frame* frames,
                              int n_files, int n_frames_total, int frames_per_file, int H, int W, int max_frame_px, uint8_t* index_ws, uint8_t* out,
                              int32_t* err, void* stream) {
  using namespace vmm;
  if (!files || !file_ofs || !frame_begin || !frames || !index_ws || !out || !err) return set_error(VMM_ERR_ARG, "vmm_gif_decode: null argument");
  if (n_files <= 0 || n_frames_total <= 0 || frames_per_file <= 0 || H <= 0 || W <= 0) return set_error(VMM_ERR_ARG, "vmm_gif_decode: bad sizes");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static bool attr_set = false;
  const size_t smem_dict = static_cast<size_t>(GIF_WARPS) * 4096 * sizeof(uint32_t);
  const size_t smem_all = smem_dict + static_cast<size_t>(GIF_WARPS) * GIF_SMEM_PX;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gif_lzw_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_dict));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gif_lzw_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_all));
    if (e != cudaSuccess) return set_cuda_error(e, "vmm_gif_decode: cudaFuncSetAttribute");
    attr_set = true;
  }
  static const bool no_smem_out = getenv("VMM_GIF_NO_SMEM_OUT") != nullptr;
  const bool smem_out = !no_smem_out && max_frame_px > 0 && max_frame_px <= GIF_SMEM_PX;     // max_frame_px <= 0: unknown, global-memory form
  const int sms = num_sms();
  if (sms <= 0) return set_error(VMM_ERR_CUDA, "vmm_gif_decode: no CUDA device");
  const unsigned int want = ceil_div(n_frames_total, GIF_WARPS);
  const unsigned int grid1 = want < static_cast<unsigned int>(sms) ? want : static_cast<unsigned int>(sms);     // one CTA (128 KB of dictionaries) per SM
  if (smem_out)
    gif_lzw_kernel<true><<<grid1, GIF_WARPS * 32, smem_all, st>>>(files, file_ofs, frame_begin, frames, n_files, n_frames_total, index_ws, err);
  else
    gif_lzw_kernel<false><<<grid1, GIF_WARPS * 32, smem_dict, st>>>(files, file_ofs, frame_begin, frames, n_files, n_frames_total, index_ws, err);
  count_launch();
  int rc = check_launch("gif_lzw_kernel");
  if (rc) return rc;
  gif_compose_kernel<<<n_files, 256, 0, st>>>(files, file_ofs, frame_begin, frames, frames_per_file, H, W, index_ws, out);
  count_launch();
  return check_launch("gif_compose_kernel");
}

extern "C" int vmm_dataset_items(const uint8_t* u8, const int64_t* index, int n, int n_planes, int topo_plane, int n_ch, const int32_t* ch_plane,
                                 const int32_t* ch_has_range, const float* sample_rng, const float* global_rng, const int32_t* sample_frames,
                                 int frames, int frames_out, int hw, float* out, void* stream) {
  using namespace vmm;
  if (!u8 || !index || !ch_plane || !ch_has_range || !sample_rng || !global_rng || !out) return set_error(VMM_ERR_ARG, "vmm_dataset_items: null argument");
  if (n <= 0 || n_planes <= 0 || n_ch <= 0 || frames <= 0 || frames_out <= 0 || hw <= 0 || topo_plane < 0 || topo_plane >= n_planes)
    return set_error(VMM_ERR_ARG, "vmm_dataset_items: bad sizes");
  if ((static_cast<long long>(frames_out) * hw) % 4 != 0) return set_error(VMM_ERR_ARG, "vmm_dataset_items: frames_out * h * w must be a multiple of 4");
  if (static_cast<long long>(n) * n_ch > 65535) return set_error(VMM_ERR_ARG, "vmm_dataset_items: more than 65535 (item, channel) planes in one call");
  const long long per_plane = static_cast<long long>(frames_out) * hw;
  unsigned int gx = ceil_div(per_plane, 256 * 4 * 4);
  if (gx == 0) gx = 1;
  dataset_items_kernel<<<dim3(gx, static_cast<unsigned int>(n * n_ch)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      u8, index, n_planes, topo_plane, n_ch, ch_plane, ch_has_range, sample_rng, global_rng, sample_frames, frames, frames_out, hw, out);
  count_launch();
  return check_launch("dataset_items_kernel");
}
