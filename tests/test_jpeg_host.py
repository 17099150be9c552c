"""CPU checks of the JPEG path: the numpy restatement (oracle/jpeg.py) against cv2.imdecode — the call the
reference makes through cv2.imread (1_doclayout_bboxes.py:381) — and the decoder's own inline code
(csrc/pg_jpeg.h), evaluated on the host chunk by chunk in the kernels' order, against cv2 as well."""
import ctypes as C

import cv2
import numpy as np
import pytest

from multimodal_embeddings_b200 import synth
from multimodal_embeddings_b200._lib import PageGeomError, check, lib
from oracle import jpeg as oj

SUBSAMPLING = {"444": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, "420": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420,
               "422": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, "440": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440}


def _page(h, w, seed, noise=4.0):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = 200 + 20 * np.sin(xx / 37.0) + 15 * np.cos(yy / 23.0) + rng.normal(0, noise, (h, w))
    return np.clip(np.where(rng.random((h, w)) < 0.08, 40, img), 0, 255).astype(np.uint8)


def _encode(img, q=95, rst=0, extra=()):
    ok, buf = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_RST_INTERVAL, rst, *extra])
    assert ok
    return buf.tobytes()


def host_decode(data: bytes, chunk_bytes: int, max_rounds: int = 1 << 20):
    """-> (grey [H, W] or BGR [H, W, 3], stats) through pg_hostcheck_jpeg_decode."""
    L = lib()
    w, h, st = C.c_int32(), C.c_int32(), (C.c_int64 * 4)()
    b = np.frombuffer(data, np.uint8)
    check(L.pg_hostcheck_jpeg_decode(b.ctypes.data, len(b), chunk_bytes, max_rounds, None, 0, C.byref(w), C.byref(h), st))
    comps = st[3]
    pitch = (comps * w.value + 15) // 16 * 16
    out = np.zeros((h.value, pitch), np.uint8)
    check(L.pg_hostcheck_jpeg_decode(b.ctypes.data, len(b), chunk_bytes, max_rounds, out.ctypes.data, pitch, C.byref(w),
                                     C.byref(h), st))
    img = out[:, :comps * w.value]
    if comps == 3:
        img = img.reshape(h.value, w.value, 3)
    return img, {"rounds": st[0], "replaced_round1": st[1], "chunks": st[2], "restarts": st[3]}


@pytest.mark.parametrize("shape", [(64, 64), (61, 77), (8, 8), (17, 130)])
@pytest.mark.parametrize("q", [95, 30, 100])
def test_numpy_restatement_equals_cv2(shape, q):
    """oracle/jpeg.py: grey and colour (4:4:4, 4:2:0, 4:2:2, 4:4:0), with and without restart markers."""
    h, w = shape
    for rst in (0, 3):
        g = _page(h, w, 1)
        data = _encode(g, q, rst)
        assert np.array_equal(oj.decode(data), cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR))
        c = np.stack([g, np.roll(g, 3, 1), 255 - g], -1)
        for ss in SUBSAMPLING.values():
            data = _encode(c, q, rst, (cv2.IMWRITE_JPEG_SAMPLING_FACTOR, ss))
            assert np.array_equal(oj.decode(data), cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR))


@pytest.mark.parametrize("shape,noise", [((64, 64), 4), ((61, 77), 40), ((8, 8), 4), ((200, 333), 4), ((517, 640), 40)])
def test_chunked_decoder_inline_code_equals_cv2(shape, noise):
    """The three-pass chunked entropy decoder + islow IDCT of csrc/pg_jpeg.h on the host: every chunk size from far
    below a block (a block then spans several chunks) to far above, restart intervals of 1 and 7 MCUs (restart
    boundaries inside chunks, on chunk edges, intervals ending without padding), qualities 30-100."""
    h, w = shape
    for q in (95, 75, 30, 100):
        for rst in (0, 1, 7):
            data = _encode(_page(h, w, 7, noise), q, rst)
            ref = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
            assert np.array_equal(ref[..., 0], ref[..., 1]) and np.array_equal(ref[..., 0], ref[..., 2])
            for chunk in (8, 16, 64, 512, 4096):
                got, st = host_decode(data, chunk)
                assert np.array_equal(got, ref[..., 0]), (shape, q, rst, chunk, st)
                assert st["restarts"] == (0 if rst == 0 else max(0, -(-(-(-h // 8) * -(-w // 8)) // rst) - 1))


@pytest.mark.parametrize("shape", [(64, 64), (61, 77), (8, 8), (17, 130), (100, 33), (203, 317)])
def test_chunked_decoder_colour_files_equal_cv2(shape):
    """Colour files through the same inline code: interleaved MCUs (4:4:4, 4:2:2, 4:2:0, 4:4:0), fancy chroma
    upsampling at odd sizes, YCbCr -> BGR; chunk sizes below and above an MCU, with and without restart markers."""
    h, w = shape
    g = _page(h, w, 9, 12)
    c = np.stack([g, np.roll(g, 5, 1), 255 - np.roll(g, 3, 0)], -1)
    for q in (95, 40):
        for rst in (0, 2):
            for ss in SUBSAMPLING.values():
                data = _encode(c, q, rst, (cv2.IMWRITE_JPEG_SAMPLING_FACTOR, ss))
                ref = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
                for chunk in (16, 256, 4096):
                    got, st = host_decode(data, chunk)
                    assert np.array_equal(got, ref), (shape, q, rst, ss, chunk, st)


def test_chunk_states_converge_in_a_few_rounds_on_a_scan_like_page():
    """Self-synchronisation is what makes the scheme parallel: on a newspaper-like page the speculative pass ends
    most chunks in the true state and the sync rounds fix the rest quickly."""
    g = synth.newspaper_page(2000, 1500, 3)
    for q, chunk, max_rounds in ((95, 512, 4), (75, 512, 3), (95, 2048, 3)):
        data = _encode(g, q)
        got, st = host_decode(data, chunk)
        assert np.array_equal(got, cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_GRAYSCALE))
        assert st["rounds"] <= max_rounds and st["replaced_round1"] < 0.25 * st["chunks"], st


def test_parser_refuses_what_the_device_path_does_not_decode():
    g = _page(40, 40, 2)
    ok, prog = cv2.imencode(".jpg", g, [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    L = lib()
    w, h, st = C.c_int32(), C.c_int32(), (C.c_int64 * 4)()
    for data in (prog.tobytes(), b"\x89PNG\r\n\x1a\n" + bytes(64), _encode(g)[:60]):
        b = np.frombuffer(data, np.uint8)
        rc = L.pg_hostcheck_jpeg_decode(b.ctypes.data, len(b), 512, 8, None, 0, C.byref(w), C.byref(h), st)
        assert rc == 4 and b"unsupported" in L.pg_last_error()  # PG_ERR_UNSUPPORTED
    with pytest.raises(PageGeomError):
        check(4)


def test_host_decode_of_a_scan_equals_cv2_imread_and_exif_turned_files_stay_on_the_host(tmp_path):
    """ops.decode_page: the plane of a greyscale file (PNG, JPEG, BMP) is what cv2.imread returns in each of its three
    channels, colour files come back as cv2.imread gives them; a JPEG whose EXIF orientation asks for a rotation —
    which cv2.imread applies — is refused by the device path's parser and decoded, turned, on the host."""
    from PIL import Image
    from multimodal_embeddings_b200 import ops
    g = _page(120, 200, 4)
    c = np.stack([g, np.roll(g, 3, 1), 255 - g], -1)
    for name, img in (("g.png", g), ("g.jpg", g), ("g.bmp", g), ("c.png", c), ("c.jpg", c), ("g.tif", g)):
        path = str(tmp_path / name)
        assert cv2.imwrite(path, img)
        ref = cv2.imread(path)
        got = ops.decode_page(path)
        if img.ndim == 2:
            assert got.ndim == 2 and all(np.array_equal(got, ref[..., k]) for k in range(3)), name
        else:
            assert np.array_equal(got, ref), name
    assert ops.decode_page(str(tmp_path / "missing.png")) is None
    exif = Image.Exif()
    exif[0x0112] = 6  # rotate 90 degrees clockwise to display
    turned = str(tmp_path / "turned.jpg")
    Image.fromarray(g).save(turned, quality=90, exif=exif)
    plain = str(tmp_path / "plain.jpg")
    Image.fromarray(g).save(plain, quality=90)
    assert cv2.imread(turned).shape[:2] == (200, 120) and cv2.imread(plain).shape[:2] == (120, 200)
    assert ops.jpeg_probe(open(plain, "rb").read()) == (200, 120, 1)
    assert ops.jpeg_probe(open(turned, "rb").read()) is None
    assert b"EXIF orientation" in lib().pg_last_error()
    got = ops.decode_page(turned)
    assert got.shape == (200, 120) and np.array_equal(got, cv2.imread(turned)[..., 0])


def test_files_from_another_encoder_with_optimised_tables(tmp_path):
    """Pillow's encoder (libjpeg via a different front end): optimised Huffman tables (codes up to 16 bits, not the
    standard tables), its own subsampling choices, a comment and an EXIF block without orientation, restart markers
    counted in MCU rows."""
    from PIL import Image
    g = _page(233, 317, 21, 25)
    c = np.stack([g, np.roll(g, 7, 1), 255 - np.roll(g, 5, 0)], -1)
    exif = Image.Exif()
    exif[0x010E] = "a scan"  # ImageDescription: an EXIF block that does not turn the image
    cases = [(g, dict(quality=92, optimize=True)), (g, dict(quality=35, optimize=True, comment=b"hello")),
             (c[..., ::-1], dict(quality=88, optimize=True, subsampling=2)), (c[..., ::-1], dict(quality=97, subsampling=0, exif=exif)),
             (c[..., ::-1], dict(quality=70, optimize=True, subsampling=1, restart_marker_rows=1)),
             (g, dict(quality=100, optimize=True, restart_marker_blocks=3))]
    for i, (img, kw) in enumerate(cases):
        path = str(tmp_path / f"p{i}.jpg")
        Image.fromarray(img).save(path, **kw)
        data = open(path, "rb").read()
        ref = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
        for chunk in (32, 512):
            got, st = host_decode(data, chunk)
            want = ref if got.ndim == 3 else ref[..., 0]
            assert np.array_equal(got, want), (i, kw, chunk, st)


def test_randomised_sweep_through_the_inline_decoder_equals_cv2():
    """The files of the GPU sweep (tests/jpeg_cases.py, same seeds) through the same inline code on the host, chunk by chunk."""
    import jpeg_cases
    for seed, chunk in ((2024, 128), (2025, 1024)):
        for data, what in jpeg_cases.sweep_files(seed, 30):
            ref = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_ANYCOLOR)
            got, _ = host_decode(data, chunk)
            assert got.shape == ref.shape and np.array_equal(got, ref), what


def test_damaged_files_are_refused_or_decoded_but_never_crash():
    """600 damaged copies (tests/jpeg_cases.py: header bytes flipped, truncations, insertions, flips anywhere) through
    the parser and the host evaluation of the decoder: an error code or some pixels, never a crash, never a size the
    file does not declare.  scripts/fuzz_jpeg_asan.sh runs 3000 of them under AddressSanitizer."""
    import jpeg_cases
    L = lib()
    seen = set()
    for data in jpeg_cases.mutated_files(7, 600):
        w, h, st = C.c_int32(), C.c_int32(), (C.c_int64 * 4)()
        b = np.frombuffer(data, np.uint8)
        rc = L.pg_hostcheck_jpeg_decode(b.ctypes.data, len(b), 256, 64, None, 0, C.byref(w), C.byref(h), st)
        seen.add(rc)
        if rc != 0 or w.value * h.value > 4_000_000:
            continue
        assert 1 <= w.value <= 65535 and 1 <= h.value <= 65535 and st[3] in (1, 3)
        pitch = (st[3] * w.value + 15) // 16 * 16
        out = np.zeros((h.value, pitch), np.uint8)
        rc = L.pg_hostcheck_jpeg_decode(b.ctypes.data, len(b), 256, 64, out.ctypes.data, pitch, C.byref(w), C.byref(h), st)
        seen.add(rc)
    assert 0 in seen and seen - {0} and all(r in (0, 1, 4) for r in seen)  # PG_OK / PG_ERR_INVALID / PG_ERR_UNSUPPORTED


def test_huffman_table_with_more_codes_than_its_code_space_is_refused():
    """A DHT whose counts claim three 1-bit codes (same number of symbols as before): the table views index by code,
    so this has to be caught while the table is built."""
    data = bytearray(_encode(_page(64, 64, 3)))
    at = data.find(b"\xff\xc4")
    counts = at + 5  # marker, length (2), class/id (1)
    later = max(range(16), key=lambda i: data[counts + i])
    assert data[counts + later] >= 3 and later > 0
    data[counts + later] -= 3 - data[counts]
    data[counts] = 3
    w, h, st = C.c_int32(), C.c_int32(), (C.c_int64 * 4)()
    b = np.frombuffer(bytes(data), np.uint8)
    rc = lib().pg_hostcheck_jpeg_decode(b.ctypes.data, len(b), 256, 8, None, 0, C.byref(w), C.byref(h), st)
    assert rc == 4 and b"Huffman" in lib().pg_last_error()
