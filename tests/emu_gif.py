"""CPU statements of the dataset entry points of the C ABI (include/vmm.h: vmm_gif_decode, vmm_dataset_items)  --  TEST INFRASTRUCTURE.

`decode(blob, table)` restates, in plain Python / numpy, what the two kernels of `vmm_gif_decode` do with the frame table the native
host scanner (`vmm_gif_scan`) produced: LZW with (offset, length) dictionary words into the output stream, then compositing with the
palette luma.  `dataset_items` restates `vmm_dataset_items` with numpy's correctly rounded fp32 operations.  The CPU tests hold these
statements to PIL (what the reference decodes with, VDDP:1076-1106) and to `Dataset.__getitem__` (pinned bit-equal to the
reference's); the `-m gpu` tests hold the kernels to the same two.  Also here: the corpus of GIF variants both suites run over.
Never imported by the product.
"""
import io

import numpy as np

NO_PALETTE = 0xFFFFFFFF


def lzw(blob: bytes, fr) -> np.ndarray:
    """Index stream (w*h,) of one frame: the arithmetic of gif_lzw_kernel."""
    p, rem, ended = int(fr["data_ofs"]), 0, False
    npx = int(fr["w"]) * int(fr["h"])
    m = int(fr["min_code"])
    clear, eoi = 1 << m, (1 << m) + 1
    nxt, size, bitbuf, nbits, pos = clear + 2, m + 1, 0, 0, 0
    prev_valid, prev_off, prev_len = False, 0, 0
    out = np.zeros(npx, dtype=np.uint8)
    dic = {}
    end = len(blob)

    def next_byte():
        nonlocal p, rem, ended
        if ended:
            return -1
        if rem == 0:
            if p >= end:
                ended = True
                return -1
            rem = blob[p]
            p += 1
            if rem == 0:
                ended = True
                return -1
        if p >= end:
            ended = True
            return -1
        rem -= 1
        p += 1
        return blob[p - 1]

    ok = True
    while pos < npx:
        while nbits < size:
            b = next_byte()
            if b < 0:
                ok = False
                break
            bitbuf |= b << nbits
            nbits += 8
        if not ok:
            break
        code = bitbuf & ((1 << size) - 1)
        bitbuf >>= size
        nbits -= size
        if code == clear:
            nxt, size, prev_valid = clear + 2, m + 1, False
            continue
        if code == eoi:
            break
        if code < clear:
            ln = 1
            out[pos] = code
        else:
            if not prev_valid:
                ok = False
                break
            if code < nxt:
                off, ln = dic[code]
            elif code == nxt:
                off, ln = prev_off, prev_len + 1
            else:
                ok = False
                break
            n = min(ln, npx - pos)
            for j in range(n):
                s = off + j
                if s >= pos:
                    s = s - pos + off
                out[pos + j] = out[s]
        if prev_valid and nxt < 4096:
            dic[nxt] = (prev_off, prev_len + 1)
            nxt += 1
            if nxt == (1 << size) and size < 12:
                size += 1
        prev_off, prev_len, prev_valid = pos, ln, True
        pos += min(ln, npx - pos)
    return out, pos >= npx


def _luma(blob, pal_ofs, pal_size, idx):
    if pal_ofs == NO_PALETTE or idx >= pal_size:
        return idx
    r, g, b = blob[pal_ofs + 3 * idx], blob[pal_ofs + 3 * idx + 1], blob[pal_ofs + 3 * idx + 2]
    return (19595 * r + 38470 * g + 7471 * b + 0x8000) >> 16


def _interlace_src_row(y, h):
    n1, n2, n3 = (h + 7) >> 3, (h + 3) >> 3, (h + 1) >> 2
    if y % 8 == 0:
        return y >> 3
    if y % 8 == 4:
        return n1 + (y >> 3)
    if y % 4 == 2:
        return n1 + n2 + (y >> 2)
    return n1 + n2 + n3 + (y >> 1)


def decode(blob: bytes, table, size_wh, frames_per_file=None) -> np.ndarray:
    """(frames, H, W) uint8: the arithmetic of gif_lzw_kernel + gif_compose_kernel for one file."""
    W, H = size_wh
    nf = len(table)
    fpf = frames_per_file if frames_per_file is not None else nf
    out = np.zeros((fpf, H, W), dtype=np.uint8)
    for k in range(min(nf, fpf)):
        fr = table[k]
        pal_ofs, pal_size = int(fr["pal_ofs"]), int(fr["pal_size"])
        lut = np.array([_luma(blob, pal_ofs, pal_size, i) for i in range(256)], dtype=np.uint8)
        if k > 0:
            cur = out[k - 1].copy()
            pf = table[k - 1]
            if int(pf["disposal"]) == 2:
                color = int(pf["transp"]) if pf["has_transp"] else int(pf["background"])
                if int(pf["pal_ofs"]) != NO_PALETTE and color >= int(pf["pal_size"]):
                    color = 0
                cur[pf["y"]:pf["y"] + pf["h"], pf["x"]:pf["x"] + pf["w"]] = _luma(blob, int(pf["pal_ofs"]), int(pf["pal_size"]), color)
        else:
            cur = np.full((H, W), lut[0], dtype=np.uint8)
        idx, _ = lzw(blob, fr)
        w, h = int(fr["w"]), int(fr["h"])
        idx = idx.reshape(h, w)
        if fr["interlace"]:
            idx = idx[[_interlace_src_row(y, h) for y in range(h)]]
        vals = lut[idx]
        region = cur[fr["y"]:fr["y"] + h, fr["x"]:fr["x"] + w]
        if fr["has_transp"]:
            keep = idx == int(fr["transp"])
            vals = np.where(keep, region, vals)
        cur[fr["y"]:fr["y"] + h, fr["x"]:fr["x"] + w] = vals
        out[k] = cur
    return out


def dataset_items(u8, index, topo_plane, ch_plane, ch_has_range, sample_rng, global_rng, sample_frames, frames_out) -> np.ndarray:
    """vmm_dataset_items in numpy float32 (every operation separately rounded).  u8 (n, planes, frames, h, w)."""
    n_ch = len(ch_plane)
    frames, h, w = u8.shape[2:]
    out = np.zeros((len(index), n_ch, frames_out, h, w), dtype=np.float32)
    for i, s in enumerate(index):
        live = min(int(sample_frames[s]), frames, frames_out)
        for c in range(n_ch):
            t = u8[s, ch_plane[c], :live].astype(np.float32) / np.float32(255)
            if ch_has_range[c]:
                t = t * np.float32(sample_rng[s, c, 1]) + np.float32(sample_rng[s, c, 0])
                t = np.where(u8[s, topo_plane, :live] == 0, np.float32(0), t)
                t = (t - np.float32(global_rng[c, 0])) / np.float32(global_rng[c, 1])
            out[i, c, :live] = t
    return out


# ----------------------------------------------------------------------------------------------------------------------------
# corpus: (name, gif bytes); PIL is the checker
# ----------------------------------------------------------------------------------------------------------------------------
def pil_frames(blob: bytes) -> np.ndarray:
    """What the reference's gif_to_tensor(channels=1) reads before ToTensor: every frame convert('L') (VDDP:1076-1106)."""
    from PIL import Image
    img = Image.open(io.BytesIO(blob))
    frames, i = [], 0
    while True:
        try:
            img.seek(i)
        except EOFError:
            break
        frames.append(np.asarray(img.convert('L'), dtype=np.uint8).copy())
        i += 1
    return np.stack(frames)


def quirk_frames(rng, h: int = 48, w: int = 64):
    """Three 'L' frames: the first holds all 256 grey values (Pillow's writer then keeps the plain ramp as the global palette and its
    reader opens the file in mode 'L'), the second misses one value (the writer compacts its LOCAL palette and uses the freed index as
    the transparent colour): the file Pillow's reader does not round-trip."""
    a = rng.integers(0, 256, (3, h, w), dtype=np.uint8)
    a[0].reshape(-1)[:256] = np.arange(256, dtype=np.uint8)
    missing = int(rng.integers(1, 255))
    a[1][a[1] == missing] = missing + 1
    return [a[0], a[1], a[2]]


def corpus(seed: int = 0):
    from PIL import Image
    rng = np.random.default_rng(seed)
    out = []

    def save(name, frames, **kw):
        bio = io.BytesIO()
        frames[0].save(bio, format='GIF', save_all=True, append_images=frames[1:], duration=200, loop=0, **kw)
        out.append((name, bio.getvalue()))

    def grey(h, w, n, lo=0, hi=256):
        return [Image.fromarray(rng.integers(lo, hi, (h, w), dtype=np.uint8), 'L') for _ in range(n)]

    def drifting(h, w, n, mode='L', boxes=True):
        """frames that differ from their predecessor inside a small box only (the writer crops such frames to the box)."""
        base = rng.integers(0, 256, (h, w), dtype=np.uint8)
        fr = []
        for k in range(n):
            if k and boxes:
                y0, x0 = int(rng.integers(0, h - 4)), int(rng.integers(0, w - 4))
                y1, x1 = int(rng.integers(y0 + 1, h)), int(rng.integers(x0 + 1, w))
                base = base.copy()
                base[y0:y1, x0:x1] = rng.integers(0, 256, (y1 - y0, x1 - x0), dtype=np.uint8)
            fr.append(Image.fromarray(base, mode))
        return fr

    save('L_noise_96x96x11', grey(96, 96, 11))                                   # the dataset's own format (interlaced by PIL's default)
    save('L_noise_not_interlaced', grey(96, 96, 3), interlace=False)
    save('L_noise_128x128', grey(128, 128, 3))                                   # above the shared-memory decode limit of the kernel: global-memory form
    save('L_boxes_160x120', drifting(120, 160, 4))
    save('L_boxes', drifting(64, 80, 6))
    save('L_smooth', [Image.fromarray((np.add.outer(np.arange(96), np.arange(96)) * (k + 1) % 256).astype(np.uint8), 'L') for k in range(4)])
    save('L_binary_mask', [Image.fromarray(((rng.random((96, 96)) > 0.3) * 255).astype(np.uint8), 'L') for _ in range(11)])
    save('L_tiny_8x8', grey(8, 8, 3))
    save('L_odd_33x17', grey(17, 33, 4))
    save('L_low_values', grey(32, 32, 3, 0, 4))                                  # long runs: dictionary fills up without a clear code
    save('L_constant', [Image.fromarray(np.full((96, 96), v, dtype=np.uint8), 'L') for v in (0, 255, 7)])
    save('L_duplicate_frames', [Image.fromarray(np.full((32, 32), v, dtype=np.uint8), 'L') for v in (5, 5, 9, 9, 9, 1)])
    save('L_optimize', drifting(48, 48, 5), optimize=True)
    save('L_disposal2', drifting(48, 48, 5), disposal=2)
    save('L_disposal1', drifting(48, 48, 5), disposal=1)
    pal = rng.integers(0, 256, 768, dtype=np.uint8).tobytes()

    def paletted(frames):
        res = []
        for f in frames:
            q = Image.fromarray(np.asarray(f), 'P')
            q.putpalette(pal)
            res.append(q)
        return res

    save('P_colour', paletted(grey(40, 40, 4)))
    save('P_colour_boxes', paletted(drifting(40, 56, 6)))
    save('P_colour_optimize', paletted(drifting(40, 56, 6)), optimize=True)
    save('P_colour_disposal2', paletted(drifting(40, 56, 5)), disposal=2)
    save('P_colour_transparency', paletted(drifting(40, 56, 5)), transparency=3, disposal=2)
    save('P_from_L_convert', [f.convert('L').convert('P') for f in grey(96, 96, 3)])     # video_tensor_to_gif's frames (VDDP:1094-1096)
    save('P_16_colours', paletted(grey(24, 24, 3, 0, 16)))
    save('one_bit', [Image.fromarray((rng.random((40, 40)) > 0.5)).convert('1') for _ in range(3)])
    save('RGB_quantised', [Image.fromarray(rng.integers(0, 256, (32, 32, 3), dtype=np.uint8), 'RGB') for _ in range(3)])
    # small random 'L' frames: not every grey value occurs, so Pillow's writer keeps the full grey ramp for frame 0 (-> mode 'L' on
    # reading) and gives later frames a compacted LOCAL palette + a transparent index: the case Pillow's reader does not round-trip
    # (VMM_GIF_PIL_COMPAT, include/vmm.h)
    for k in range(3):
        save(f'L_pillow_quirk_{k}', [Image.fromarray(a, 'L') for a in quirk_frames(rng)])
    save('L_interlaced', grey(96, 96, 3), interlace=True)
    save('L_interlaced_odd', grey(37, 21, 3), interlace=True)
    # files from the writer below: small code sizes, interlaced partial frames, local palettes, transparency + disposal 2, deferred clear
    for bits in (1, 2, 3, 5, 8):
        n_col = 1 << bits
        gp = rng.integers(0, 256, (n_col, 3), dtype=np.uint8)
        frames = []
        for k in range(4):
            h, w = (30, 44) if k == 0 else (int(rng.integers(1, 30)), int(rng.integers(1, 44)))
            y, x = (0, 0) if k == 0 else (int(rng.integers(0, 31 - h)), int(rng.integers(0, 45 - w)))
            frames.append(dict(x=x, y=y, idx=rng.integers(0, n_col, (h, w), dtype=np.uint8), interlace=bool(k % 2),
                               transp=(int(rng.integers(0, n_col)) if k >= 2 else None), disposal=(0, 1, 2, 2)[k],
                               palette=(rng.integers(0, 256, (n_col, 3), dtype=np.uint8) if k == 3 else None)))
        out.append((f'own_writer_{bits}bit', write_gif(44, 30, gp, frames, background=n_col - 1)))
    # a long low-entropy frame without clear codes once the table is full (deferred clear), and one that clears often
    big = rng.integers(0, 2, (96, 96), dtype=np.uint8) * rng.integers(0, 2, (96, 96), dtype=np.uint8)
    gp = rng.integers(0, 256, (4, 3), dtype=np.uint8)
    out.append(('own_writer_deferred_clear', write_gif(96, 96, gp, [dict(x=0, y=0, idx=np.tile(big, (1, 1)), clear_when_full=False)] * 2)))
    out.append(('own_writer_eager_clear', write_gif(96, 96, gp, [dict(x=0, y=0, idx=big, clear_every=50)] * 2)))
    return out


def _lzw_encode(idx: np.ndarray, min_code: int, clear_when_full: bool = True, clear_every: int = 0) -> bytes:
    """Plain GIF LZW encoder (variable code size, LSB first).  With clear_when_full=False the table is left full (deferred clear)."""
    clear, eoi = 1 << min_code, (1 << min_code) + 1
    table = {(i,): i for i in range(clear)}
    nxt, size = clear + 2, min_code + 1
    bits, nb, data = 0, 0, bytearray()

    def emit(code):
        nonlocal bits, nb
        bits |= code << nb
        nb += size
        while nb >= 8:
            data.append(bits & 255)
            bits >>= 8
            nb -= 8

    emit(clear)
    cur, emitted = (), 0
    for v in idx.ravel().tolist():
        cand = cur + (v,)
        if cand in table:
            cur = cand
            continue
        emit(table[cur])
        emitted += 1
        if nxt < 4096:
            table[cand] = nxt
            nxt += 1
            if nxt - 1 == (1 << size) and size < 12:
                size += 1
        elif clear_when_full:
            emit(clear)
            table = {(i,): i for i in range(clear)}
            nxt, size = clear + 2, min_code + 1
        if clear_every and emitted % clear_every == 0:
            emit(clear)
            table = {(i,): i for i in range(clear)}
            nxt, size = clear + 2, min_code + 1
        cur = (v,)
    if cur:
        emit(table[cur])
    emit(eoi)
    if nb:
        data.append(bits & 255)
    return bytes(data)


def write_gif(W: int, H: int, palette: np.ndarray, frames, background: int = 0) -> bytes:
    """Minimal GIF89a writer for the corpus: global palette of 2^k colours, per frame a rectangle of indices and optionally
    interlace / a transparent index / a disposal method / a local palette."""
    n_col = len(palette)
    k = max(int(np.log2(n_col)), 1)
    b = bytearray(b'GIF89a')
    b += int(W).to_bytes(2, 'little') + int(H).to_bytes(2, 'little') + bytes([0x80 | (k - 1), background, 0])
    b += np.asarray(palette, dtype=np.uint8).tobytes()
    if n_col < 2:
        b += bytes(3)
    for fr in frames:
        idx = np.asarray(fr['idx'], dtype=np.uint8)
        h, w = idx.shape
        transp, disposal = fr.get('transp'), fr.get('disposal', 0)
        b += bytes([0x21, 0xF9, 4, (disposal << 2) | (1 if transp is not None else 0), 20, 0, transp or 0, 0])
        flags = 0x40 if fr.get('interlace') else 0
        lp = fr.get('palette')
        if lp is not None:
            flags |= 0x80 | (max(int(np.log2(len(lp))), 1) - 1)
        b += bytes([0x2C]) + int(fr['x']).to_bytes(2, 'little') + int(fr['y']).to_bytes(2, 'little') + w.to_bytes(2, 'little') + h.to_bytes(2, 'little')
        b += bytes([flags])
        if lp is not None:
            b += np.asarray(lp, dtype=np.uint8).tobytes()
            if len(lp) < 2:
                b += bytes(3)
        if fr.get('interlace'):
            rows = [y for y in range(0, h, 8)] + [y for y in range(4, h, 8)] + [y for y in range(2, h, 4)] + [y for y in range(1, h, 2)]
            idx = idx[rows]
        min_code = max(k, 2)
        data = _lzw_encode(idx, min_code, fr.get('clear_when_full', True), fr.get('clear_every', 0))
        b += bytes([min_code])
        for i in range(0, len(data), 255):
            chunk = data[i:i + 255]
            b += bytes([len(chunk)]) + chunk
        b += b'\x00'
    b += b';'
    return bytes(b)
