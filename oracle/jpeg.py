"""CPU restatement of baseline JPEG decoding as `cv2.imread` performs it — TEST INFRASTRUCTURE ONLY.

The reference opens every scan with `cv2.imread(image_path)` (1_doclayout_bboxes.py:381, 2_edge_box_filter.py:195);
for a `.jpg` that is libjpeg-turbo (bundled in the opencv wheel; third-party, not under /root/reference) with its
defaults: Huffman entropy decoding (ITU-T T.81 Annex F), dequantisation, the "islow" integer inverse DCT
(jidctint.c: 13-bit constants, two passes, PASS1_BITS = 2), level shift + clamp, "fancy" (triangle) chroma
upsampling for subsampled colour files (jdsample.c) and the fixed-point YCbCr -> RGB conversion (jdcolor.c,
16-bit scaled constants).  cv2 then hands back BGR (three equal channels for a greyscale file).

This module restates that published algorithm in numpy / plain Python, sequentially, for small images; it is
pinned by `cv2.imdecode` itself in tests/test_oracle_golden.py (bit-exact on grey and colour files, with and
without restart markers, odd sizes, all common subsamplings).  The CUDA decoder (csrc/pg_jpeg.cu) is checked
against cv2 directly on the GPU box — cv2 is part of the image — and against this restatement where the tests
need intermediate values (coefficients, per-chunk decoder states).
"""
from __future__ import annotations

import numpy as np

ZIGZAG = np.array([
    0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55,
    62, 63], np.int32)


class JpegError(ValueError):
    pass


def parse(data: bytes) -> dict:
    """Marker segments of a baseline (SOF0) / extended-sequential Huffman (SOF1, 8-bit) file with one scan."""
    if data[:2] != b"\xff\xd8":
        raise JpegError("no SOI")
    pos, info = 2, {"qt": {}, "dc": {}, "ac": {}, "dri": 0}
    while True:
        if data[pos] != 0xFF:
            raise JpegError("marker expected")
        while data[pos + 1] == 0xFF:
            pos += 1
        m = data[pos + 1]
        pos += 2
        if m in (0x01,) or 0xD0 <= m <= 0xD7:
            continue
        ln = (data[pos] << 8) | data[pos + 1]
        seg = data[pos + 2: pos + ln]
        if m == 0xDB:
            k = 0
            while k < len(seg):
                pq, tq = seg[k] >> 4, seg[k] & 15
                if pq:
                    q = np.frombuffer(seg[k + 1:k + 129], ">u2").astype(np.int32)
                    k += 129
                else:
                    q = np.frombuffer(seg[k + 1:k + 65], np.uint8).astype(np.int32)
                    k += 65
                nat = np.zeros(64, np.int32)
                nat[ZIGZAG] = q
                info["qt"][tq] = nat
        elif m in (0xC0, 0xC1):
            if seg[0] != 8:
                raise JpegError("precision")
            info["height"], info["width"] = (seg[1] << 8) | seg[2], (seg[3] << 8) | seg[4]
            info["comps"] = [{"id": seg[6 + 3 * i], "h": seg[7 + 3 * i] >> 4, "v": seg[7 + 3 * i] & 15, "tq": seg[8 + 3 * i]}
                             for i in range(seg[5])]
        elif m in (0xC2, 0xC3, 0xC5, 0xC6, 0xC7, 0xC9, 0xCA, 0xCB, 0xCD, 0xCE, 0xCF):
            raise JpegError("not a baseline Huffman file")
        elif m == 0xC4:
            k = 0
            while k < len(seg):
                tc, th = seg[k] >> 4, seg[k] & 15
                counts = list(seg[k + 1:k + 17])
                n = sum(counts)
                info["ac" if tc else "dc"][th] = (counts, list(seg[k + 17:k + 17 + n]))
                k += 17 + n
        elif m == 0xDD:
            info["dri"] = (seg[0] << 8) | seg[1]
        elif m == 0xDA:
            ns = seg[0]
            if ns != len(info["comps"]):
                raise JpegError("multi-scan files are not handled")
            for i in range(ns):
                c = next(c for c in info["comps"] if c["id"] == seg[1 + 2 * i])
                c["td"], c["ta"] = seg[2 + 2 * i] >> 4, seg[2 + 2 * i] & 15
            info["scan_begin"] = pos + ln
            end = data.rfind(b"\xff\xd9")
            info["scan_end"] = end if end >= 0 else len(data)
            return info
        pos += ln


def huff_table(counts, symbols):
    """code -> (length, symbol) as a dict keyed by (length, code) (canonical codes, T.81 Annex C)."""
    table, code, k = {}, 0, 0
    for ln in range(1, 17):
        for _ in range(counts[ln - 1]):
            table[(ln, code)] = symbols[k]
            code += 1
            k += 1
        code <<= 1
    return table


class BitReader:
    def __init__(self, data: bytes):
        self.d, self.pos, self.buf, self.n = data, 0, 0, 0

    def bit(self):
        if self.n == 0:
            b = self.d[self.pos] if self.pos < len(self.d) else 0
            self.pos += 1
            self.buf, self.n = b, 8
        self.n -= 1
        return (self.buf >> self.n) & 1

    def bits(self, k):
        v = 0
        for _ in range(k):
            v = (v << 1) | self.bit()
        return v

    def huff(self, table):
        code = 0
        for ln in range(1, 17):
            code = (code << 1) | self.bit()
            if (ln, code) in table:
                return table[(ln, code)]
        raise JpegError("bad Huffman code")


def extend(v, s):
    return v - (1 << s) + 1 if s and v < (1 << (s - 1)) else v


def unstuff(scan: bytes):
    """Entropy-coded bytes with FF00 -> FF and the RSTn markers cut out; returns (bytes, [interval start offsets])."""
    out, starts, i = bytearray(), [0], 0
    while i < len(scan):
        b = scan[i]
        if b == 0xFF and i + 1 < len(scan):
            nx = scan[i + 1]
            if nx == 0:
                out.append(0xFF)
                i += 2
                continue
            if 0xD0 <= nx <= 0xD7:
                starts.append(len(out))
                i += 2
                continue
            if nx == 0xFF:
                i += 1
                continue
        out.append(b)
        i += 1
    return bytes(out), starts


def decode_coefficients(data: bytes):
    """-> (info, [per component int32 array [blocks_h, blocks_w, 64] in natural order, dequantised])."""
    info = parse(data)
    comps = info["comps"]
    hmax, vmax = max(c["h"] for c in comps), max(c["v"] for c in comps)
    single = len(comps) == 1
    if single:  # a one-component scan is never interleaved: MCU = one block, whatever the sampling factors say
        comps[0]["h"] = comps[0]["v"] = hmax = vmax = 1
    mcus_w = -(-info["width"] // (8 * hmax))
    mcus_h = -(-info["height"] // (8 * vmax))
    for c in comps:
        c["bw"], c["bh"] = mcus_w * c["h"], mcus_h * c["v"]
        c["coef"] = np.zeros((c["bh"], c["bw"], 64), np.int32)
        c["dct"], c["act"] = huff_table(*info["dc"][c["td"]]), huff_table(*info["ac"][c["ta"]])
    stream, starts = unstuff(data[info["scan_begin"]:info["scan_end"]])
    starts.append(len(stream))
    ri = info["dri"] or mcus_w * mcus_h
    mcu = 0
    for seg in range(len(starts) - 1):
        br = BitReader(stream[starts[seg]:starts[seg + 1]])
        pred = [0] * len(comps)
        for _ in range(ri):
            if mcu >= mcus_w * mcus_h:
                break
            my, mx = divmod(mcu, mcus_w)
            for ci, c in enumerate(comps):
                for by in range(c["v"]):
                    for bx in range(c["h"]):
                        blk = c["coef"][my * c["v"] + by, mx * c["h"] + bx]
                        s = br.huff(c["dct"])
                        pred[ci] += extend(br.bits(s), s)
                        blk[0] = pred[ci]
                        k = 1
                        while k < 64:
                            rs = br.huff(c["act"])
                            r, s = rs >> 4, rs & 15
                            if s == 0:
                                if r != 15:
                                    break
                                k += 16
                                continue
                            k += r
                            blk[ZIGZAG[k]] = extend(br.bits(s), s)
                            k += 1
            mcu += 1
    for c in comps:
        c["coef"] *= info["qt"][c["tq"]][None, None, :]
    info.update(hmax=hmax, vmax=vmax, mcus_w=mcus_w, mcus_h=mcus_h)
    return info, [c["coef"] for c in comps]


# ---- jidctint.c ("islow"), vectorised over blocks -------------------------------------------------------
CONST_BITS, PASS1_BITS = 13, 2
F_0_298, F_0_390, F_0_541, F_0_765, F_0_899, F_1_175 = 2446, 3196, 4433, 6270, 7373, 9633
F_1_501, F_1_847, F_1_961, F_2_053, F_2_562, F_3_072 = 12299, 15137, 16069, 16819, 20995, 25172


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def _idct_1d(v, shift):
    """v: [..., 8] int64 along the transformed axis -> [..., 8]."""
    z2, z3 = v[..., 2], v[..., 6]
    z1 = (z2 + z3) * F_0_541
    tmp2 = z1 + z3 * (-F_1_847)
    tmp3 = z1 + z2 * F_0_765
    z2, z3 = v[..., 0], v[..., 4]
    tmp0 = (z2 + z3) << CONST_BITS
    tmp1 = (z2 - z3) << CONST_BITS
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    tmp0, tmp1, tmp2, tmp3 = v[..., 7], v[..., 5], v[..., 3], v[..., 1]
    z1, z2, z3, z4 = tmp0 + tmp3, tmp1 + tmp2, tmp0 + tmp2, tmp1 + tmp3
    z5 = (z3 + z4) * F_1_175
    tmp0, tmp1, tmp2, tmp3 = tmp0 * F_0_298, tmp1 * F_2_053, tmp2 * F_3_072, tmp3 * F_1_501
    z1, z2, z3, z4 = z1 * (-F_0_899), z2 * (-F_2_562), z3 * (-F_1_961), z4 * (-F_0_390)
    z3, z4 = z3 + z5, z4 + z5
    tmp0, tmp1, tmp2, tmp3 = tmp0 + z1 + z3, tmp1 + z2 + z4, tmp2 + z2 + z3, tmp3 + z1 + z4
    out = np.stack([tmp10 + tmp3, tmp11 + tmp2, tmp12 + tmp1, tmp13 + tmp0,
                    tmp13 - tmp0, tmp12 - tmp1, tmp11 - tmp2, tmp10 - tmp3], -1)
    return _descale(out, shift)


def idct_islow(coef: np.ndarray) -> np.ndarray:
    """coef [..., 64] dequantised, natural order -> uint8 samples [..., 8, 8]."""
    b = coef.astype(np.int64).reshape(coef.shape[:-1] + (8, 8))          # [row, col]
    ws = _idct_1d(np.swapaxes(b, -1, -2), CONST_BITS - PASS1_BITS)        # pass 1: along columns -> [col, row']
    ws = np.swapaxes(ws, -1, -2)                                          # [row', col]
    px = _idct_1d(ws, CONST_BITS + PASS1_BITS + 3)                        # pass 2: along rows
    return np.clip(px + 128, 0, 255).astype(np.uint8)


def _plane(coef: np.ndarray) -> np.ndarray:
    bh, bw, _ = coef.shape
    return idct_islow(coef).transpose(0, 2, 1, 3).reshape(bh * 8, bw * 8)


# ---- jdsample.c fancy upsampling ------------------------------------------------------------------------
def _h2v1_fancy(p):
    p = p.astype(np.int32)
    left = np.concatenate([p[:, :1], p[:, :-1]], 1)
    right = np.concatenate([p[:, 1:], p[:, -1:]], 1)
    out = np.empty((p.shape[0], 2 * p.shape[1]), np.int32)
    out[:, 0::2] = (3 * p + left + 1) >> 2
    out[:, 1::2] = (3 * p + right + 2) >> 2
    out[:, 0] = p[:, 0]
    out[:, -1] = p[:, -1]
    return out.astype(np.uint8)


def _h2v2_fancy(p):
    p = p.astype(np.int32)
    up = np.concatenate([p[:1], p[:-1]], 0)
    dn = np.concatenate([p[1:], p[-1:]], 0)
    rows = np.empty((2 * p.shape[0], p.shape[1]), np.int32)
    rows[0::2] = 3 * p + up      # colsum of the upper output row: nearer = this, farther = the row above
    rows[1::2] = 3 * p + dn
    left = np.concatenate([rows[:, :1], rows[:, :-1]], 1)
    right = np.concatenate([rows[:, 1:], rows[:, -1:]], 1)
    out = np.empty((rows.shape[0], 2 * rows.shape[1]), np.int32)
    out[:, 0::2] = (3 * rows + left + 8) >> 4
    out[:, 1::2] = (3 * rows + right + 7) >> 4
    out[:, 0] = (4 * rows[:, 0] + 8) >> 4
    out[:, -1] = (4 * rows[:, -1] + 7) >> 4
    return out.astype(np.uint8)


def _h1v2_fancy(p):
    p = p.astype(np.int32)
    up = np.concatenate([p[:1], p[:-1]], 0)
    dn = np.concatenate([p[1:], p[-1:]], 0)
    out = np.empty((2 * p.shape[0], p.shape[1]), np.int32)
    out[0::2] = (3 * p + up + 1) >> 2
    out[1::2] = (3 * p + dn + 2) >> 2
    return out.astype(np.uint8)


def ycc_to_bgr(y, cb, cr):
    """jdcolor.c: SCALEBITS = 16 tables."""
    y, cb, cr = y.astype(np.int32), cb.astype(np.int32) - 128, cr.astype(np.int32) - 128
    half = 1 << 15
    r = y + ((91881 * cr + half) >> 16)
    b = y + ((116130 * cb + half) >> 16)
    g = y + ((-22554 * cb - 46802 * cr + half) >> 16)
    return np.clip(np.stack([b, g, r], -1), 0, 255).astype(np.uint8)


def decode(data: bytes) -> np.ndarray:
    """What cv2.imdecode(data, cv2.IMREAD_COLOR) returns: uint8 [H, W, 3] BGR."""
    info, coefs = decode_coefficients(data)
    h, w = info["height"], info["width"]
    comps = info["comps"]
    planes = [_plane(c) for c in coefs]
    if len(comps) == 1:
        g = planes[0][:h, :w]
        return np.repeat(g[..., None], 3, -1)
    if len(comps) != 3:
        raise JpegError("component count")
    hmax, vmax = info["hmax"], info["vmax"]
    full = []
    for c, p in zip(comps, planes):
        # the upsampler sees the component's real (downsampled) extent, edge-replicated
        dw, dh = -(-w * c["h"] // hmax), -(-h * c["v"] // vmax)
        p = p[:dh, :dw]
        fx, fy = hmax // c["h"], vmax // c["v"]
        if (fx, fy) == (1, 1):
            q = p
        elif (fx, fy) == (2, 1):
            q = _h2v1_fancy(p)
        elif (fx, fy) == (2, 2):
            q = _h2v2_fancy(p)
        elif (fx, fy) == (1, 2):
            q = _h1v2_fancy(p)
        else:
            raise JpegError("sampling factors")
        full.append(q[:h, :w])
    return ycc_to_bgr(*full)
