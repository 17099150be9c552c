"""Randomised JPEG files shared by the decoder tests (host emulation in test_jpeg_host.py, kernels in
test_gpu_jpeg.py): random sizes (mostly not multiples of 8 or 16), qualities 20-100, restart intervals, greyscale and
every colour subsampling, cv2's and Pillow's encoders (standard and optimised Huffman tables)."""
import io

import cv2
import numpy as np

CV2_SUBSAMPLING = [cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420,
                   cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440]


def scan_like(h, w, seed, noise=4.0):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = 200 + 20 * np.sin(xx / 37.0) + 15 * np.cos(yy / 23.0) + rng.normal(0, noise, (h, w))
    return np.clip(np.where(rng.random((h, w)) < 0.08, 40, img), 0, 255).astype(np.uint8)


def cv2_encode(img, q=95, rst=0, extra=()):
    ok, buf = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_RST_INTERVAL, rst, *extra])
    assert ok
    return buf.tobytes()


def pillow_encode(img, q, optimize, subsampling, restart_blocks=0):
    from PIL import Image, ImageFile
    ImageFile.MAXBLOCK = max(ImageFile.MAXBLOCK, 1 << 24)  # libjpeg's optimising pass writes the whole file in one go
    buf = io.BytesIO()
    kw = {"restart_marker_blocks": restart_blocks} if restart_blocks else {}
    Image.fromarray(img[..., ::-1] if img.ndim == 3 else img).save(buf, format="JPEG", quality=q, optimize=optimize,
                                                                   subsampling=subsampling, **kw)
    return buf.getvalue()


def sweep_files(seed, n, max_side=900):
    """-> list of (file bytes, description)"""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        h, w = int(rng.integers(8, max_side)), int(rng.integers(8, max_side))
        g = scan_like(h, w, int(rng.integers(1 << 30)), float(rng.choice([2.0, 8.0, 30.0])))
        colour = rng.random() < 0.5
        img = np.stack([g, np.roll(g, 3, 1), 255 - np.roll(g, 2, 0)], -1) if colour else g
        q = int(rng.integers(20, 101))
        if rng.random() < 0.3:
            opt, sub = bool(rng.random() < 0.7), int(rng.integers(0, 3))
            rb = int(rng.integers(1, 9)) if rng.random() < 0.4 else 0
            out.append((pillow_encode(img, q, opt, sub, rb), f"pillow {w}x{h} c={colour} q={q} opt={opt} sub={sub} rst={rb}"))
        else:
            sub = int(rng.integers(0, 4))
            rst = int(rng.choice([0, 0, 1, 5, 40]))
            extra = (cv2.IMWRITE_JPEG_SAMPLING_FACTOR, CV2_SUBSAMPLING[sub]) if colour else ()
            out.append((cv2_encode(img, q, rst, extra), f"cv2 {w}x{h} c={colour} q={q} sub={sub} rst={rst}"))
    return out


def mutated_files(seed, n, bases=None):
    """-> n damaged copies of a few small files: bytes flipped in the headers, anywhere, truncations, insertions."""
    rng = np.random.default_rng(seed)
    bases = bases or [f for f, _ in sweep_files(99, 6, max_side=200)]
    out = []
    for it in range(n):
        b = bytearray(bases[it % len(bases)])
        kind = int(rng.integers(0, 4))
        if kind == 0:
            for _ in range(int(rng.integers(1, 4))):
                b[int(rng.integers(0, min(len(b), 700)))] = int(rng.integers(0, 256))
        elif kind == 1:
            b = b[: int(rng.integers(2, len(b)))]
        elif kind == 2:
            for _ in range(int(rng.integers(1, 6))):
                b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        else:
            at = int(rng.integers(0, len(b)))
            b[at:at] = bytes(rng.integers(0, 256, int(rng.integers(1, 40))).tolist())
        out.append(bytes(b))
    return out


def scan_damaged_files(seed, n, bases=None):
    """-> n copies with bytes flipped only inside the entropy-coded data (headers intact: the device decodes them)."""
    rng = np.random.default_rng(seed)
    bases = bases or [f for f, _ in sweep_files(98, 6, max_side=300)]
    out = []
    for it in range(n):
        b = bytearray(bases[it % len(bases)])
        sos = b.rfind(b"\xff\xda")
        lo = sos + 2 + ((b[sos + 2] << 8) | b[sos + 3])
        for _ in range(int(rng.integers(1, 8))):
            b[int(rng.integers(lo, len(b) - 2))] = int(rng.integers(0, 256))
        out.append(bytes(b))
    return out
