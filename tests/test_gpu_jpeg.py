"""GPU parity of the compressed-input path: JPEG files -> pages in HBM (pg_jpeg_decode) against cv2.imdecode (what
the reference's cv2.imread returns, 1_doclayout_bboxes.py:381), and the one-channel tiler against the
three-channel tiler on the replicated page."""
import cv2
import numpy as np
import pytest
import torch

from multimodal_embeddings_b200 import ops, synth
from multimodal_embeddings_b200._lib import PageGeomError

pytestmark = pytest.mark.gpu


def _page(h, w, seed, noise=4.0):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = 200 + 20 * np.sin(xx / 37.0) + 15 * np.cos(yy / 23.0) + rng.normal(0, noise, (h, w))
    return np.clip(np.where(rng.random((h, w)) < 0.08, 40, img), 0, 255).astype(np.uint8)


def _encode(img, q=95, rst=0, extra=()):
    ok, buf = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_RST_INTERVAL, rst, *extra])
    assert ok
    return buf.tobytes()


def _ref(data):
    bgr = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
    assert np.array_equal(bgr[..., 0], bgr[..., 1]) and np.array_equal(bgr[..., 0], bgr[..., 2])
    return bgr[..., 0]


@pytest.mark.parametrize("chunk", [64, 256, 512, 4096])
def test_decode_batch_of_mixed_files_equals_cv2(chunk):
    """One launch sequence for files of different sizes, qualities and restart intervals (none, every MCU, every
    7 MCUs, one per MCU row), sizes that are not multiples of 8, a single-block image."""
    specs = [(64, 64, 95, 0), (61, 77, 75, 1), (8, 8, 95, 0), (517, 640, 30, 7), (1003, 1501, 95, 0), (1000, 760, 100, 0),
             (333, 200, 90, 25), (1640, 1180, 85, 148), (40, 3000, 95, 0)]
    if chunk < 256:  # blocks of a quality-100 file are longer than such chunks: dozens of sync rounds (its own test below)
        specs = [sp for sp in specs if sp[2] < 100]
    files = [_encode(_page(h, w, 10 + i, 4 if i % 2 else 30), q, rst) for i, (h, w, q, rst) in enumerate(specs)]
    dec = ops.JpegDecoder(chunk_bytes=chunk, sync_rounds=8)
    pages = ops.decode_jpeg_files(files, dec)
    for data, page, (h, w, _, _) in zip(files, pages, specs):
        assert page.shape == (h, ops.row_pitch(w, 1))
        assert np.array_equal(page[:, :w].cpu().numpy(), _ref(data))
    assert dec.status()["status"] == 0


def test_decode_full_size_scan_like_pages_equals_cv2():
    """8000x6000 newspaper-like pages at cv2's default quality and at 75, automatic chunk size."""
    files = [_encode(synth.newspaper_page(8000, 6000, 5), 95), _encode(synth.newspaper_page(8000, 6000, 6), 75),
             _encode(synth.newspaper_page(7934, 5755, 7), 95, rst=992)]
    dec = ops.JpegDecoder()
    pages = ops.decode_jpeg_files(files, dec)
    st = dec.status()
    assert st["status"] == 0 and st["rounds_used"] <= dec.sync_rounds
    for data, page in zip(files, pages):
        ref = _ref(data)
        assert np.array_equal(page[:, :ref.shape[1]].cpu().numpy(), ref)


def test_decode_retries_with_more_rounds_when_the_states_have_not_converged():
    """Quality 95 on a noisy page: blocks of several hundred bits against 256-byte chunks need about a dozen sync
    rounds; with one round configured the status word reports it and check() decodes again with more rounds until the
    fixed point is reached."""
    data = _encode(_page(600, 800, 3, 40), 95)
    dec = ops.JpegDecoder(chunk_bytes=256, sync_rounds=1)
    blob, off = ops.pack_files([data])
    dec.set_files(blob, off)
    pages = dec.decode(blob.cuda())
    torch.cuda.synchronize()
    assert dec.status()["status"] != 0
    st = dec.check()
    assert st["status"] == 0 and dec.sync_rounds > 1
    assert np.array_equal(pages[0][:, :800].cpu().numpy(), _ref(data))


def test_unsupported_files_are_refused_at_parse_time():
    g = _page(64, 64, 1)
    ok, prog = cv2.imencode(".jpg", g, [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    dec = ops.JpegDecoder()
    with pytest.raises(PageGeomError, match="unsupported"):
        dec.set_files(*ops.pack_files([prog.tobytes()]))
    assert ops.jpeg_probe(prog.tobytes()) is None and ops.jpeg_probe(_encode(g)) == (64, 64, 1)


SUBSAMPLING = [cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422,
               cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440]


def test_decode_colour_files_equals_cv2_and_feeds_the_three_channel_tiler():
    """Colour scans (YCbCr, every common subsampling, odd sizes, restart markers) mixed with a greyscale file in one
    batch: BGR pages bit-identical to cv2.imdecode; tiles from the device-decoded pages == tiles from cv2's pages."""
    specs = [(517, 640), (1003, 1501), (333, 201), (64, 64), (1200, 1600), (77, 61), (900, 700), (1001, 777)]
    files, refs = [], []
    for i, (h, w) in enumerate(specs):
        g = synth.newspaper_page(w, h, 40 + i)
        c = np.stack([g, np.roll(g, 5, 1), 255 - np.roll(g, 3, 0)], -1)
        files.append(_encode(c, (95, 75, 40)[i % 3], (0, 7)[i % 2], (cv2.IMWRITE_JPEG_SAMPLING_FACTOR, SUBSAMPLING[i % 4])))
    files.append(_encode(synth.newspaper_page(800, 600, 9), 95))
    dec = ops.JpegDecoder()
    pages = ops.decode_jpeg_files(files, dec)
    assert [sz[2] for sz in dec.sizes] == [3] * len(specs) + [1]
    for data, page, (w, h, c) in zip(files, pages, dec.sizes):
        ref = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
        got = page[:, :c * w].cpu().numpy()
        assert np.array_equal(got, ref.reshape(h, 3 * w) if c == 3 else ref[..., 0])
        refs.append(ref)
    grids = [(1, 1), (2, 2)]
    sizes = [(w, h) for w, h, c in dec.sizes if c == 3]
    b_dev = ops.TileBatch(sizes, grids, 20.0)
    b_dev.bind(pages[:len(specs)])
    b_dev.run()
    b_ref = ops.TileBatch(sizes, grids, 20.0)
    b_ref.bind(ops.upload_pages_pinned(refs[:len(specs)]))
    b_ref.run()
    torch.cuda.synchronize()
    for a, b in zip(b_dev.outs, b_ref.outs):
        assert torch.equal(a, b)


@pytest.mark.parametrize("w,h,grids", [(8000, 6000, [(4, 4)]), (8000, 6000, [(1, 1), (2, 2), (3, 3), (4, 4)]),
                                       (3801, 5601, [(2, 2)]), (1501, 1003, [(1, 1), (2, 3)]), (40000, 700, [(1, 2)])])
def test_one_channel_tiler_equals_three_channel_tiler_on_the_replicated_page(w, h, grids):
    """channels=1 plans: tiles bit-identical to the BGR path fed with the plane replicated three times (what
    cv2.imread returns for a greyscale scan), through the staged kernel and the direct validation kernel, incl. the
    reference's default 30-tile grid set and tiles cut into column chunks."""
    g = synth.newspaper_page(w, h, 11)
    plan3 = ops.TilePlan(w, h, grids, 20.0)
    plan1 = ops.TilePlan(w, h, grids, 20.0, channels=1)
    assert plan1.out_elems == plan3.out_elems and plan1.tiles == plan3.tiles
    assert plan1.algorithmic_bytes == w * h + 2 * plan1.out_elems
    want = plan3.run(ops.upload_pages([np.repeat(g[..., None], 3, -1)], plan3))
    page1 = torch.zeros((1, h, plan1.pitch), dtype=torch.uint8)
    page1[0, :, :w] = torch.from_numpy(g)
    page1 = page1.cuda()
    got = plan1.run(page1)
    got_direct = plan1.run(page1, direct=True)
    torch.cuda.synchronize()
    assert torch.equal(got, want) and torch.equal(got_direct, want)


def test_jpeg_to_tiles_equals_cv2_imread_to_tiles():
    """The whole front of the path on compressed input: files -> device decode -> one-channel tiler batch (pages of
    different sizes in one launch) against cv2.imdecode -> BGR upload -> three-channel tiler batch."""
    sizes = [(1501, 1003), (777, 1001), (2400, 1800)]
    files = [_encode(synth.newspaper_page(w, h, 20 + i), 95) for i, (w, h) in enumerate(sizes)]
    pages = ops.decode_jpeg_files(files)
    grids = [(1, 1), (2, 2)]
    b1 = ops.TileBatch(sizes, grids, 20.0, channels=1)
    b1.bind(pages)
    b1.run()
    bgr = [cv2.imdecode(np.frombuffer(f, np.uint8), cv2.IMREAD_COLOR) for f in files]
    b3 = ops.TileBatch(sizes, grids, 20.0)
    b3.bind(ops.upload_pages_pinned(bgr))
    b3.run()
    torch.cuda.synchronize()
    for a, b in zip(b1.outs, b3.outs):
        assert torch.equal(a, b)


def test_stage1_cli_on_jpeg_scans_equals_the_host_decode_run(tmp_path):
    """1_doclayout_bboxes.py drop-in on a folder of .jpg scans (greyscale, colour 4:2:0, progressive) and a .png: the
    run that decodes baseline JPEGs on the device writes the same JSON files and the same tile pixels as the run with
    `--host_decode` (cv2.imread for every file, the reference's call)."""
    import json
    import os
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    from multimodal_embeddings_b200 import cli
    src = tmp_path / "in"
    src.mkdir()
    g1, g2 = synth.newspaper_page(1501, 1003, 1), synth.newspaper_page(900, 1300, 2)
    (src / "a grey.jpg").write_bytes(_encode(g1, 95))
    (src / "b colour.jpeg").write_bytes(_encode(np.stack([g2, np.roll(g2, 4, 1), 255 - g2], -1), 90,
                                              extra=(cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420)))
    (src / "c progressive.jpg").write_bytes(cv2.imencode(".jpg", g1[:800, :1000], [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])[1].tobytes())
    cv2.imwrite(str(src / "d.png"), g2[:700, :650])
    outs = {}
    for mode, extra in (("device", []), ("host", ["--host_decode"])):
        out = tmp_path / mode
        assert cli.main_stage1(["--input_folder", str(src), "--output_folder", str(out), "--grids", "2x2", "--detector",
                                "stub_detector:StubPlugin", "--write_tiles"] + extra) == 0
        files = {}
        for d, _, fs in os.walk(out):
            for fn in fs:
                path = os.path.join(d, fn)
                rel = os.path.relpath(path, out)
                if fn.endswith(".json"):
                    files[rel] = open(path).read().replace(str(out), "<OUT>")
                elif os.sep + "images" + os.sep in path:
                    files[rel] = cv2.imread(path, cv2.IMREAD_UNCHANGED).tobytes() if not fn.lower().endswith((".jpg", ".jpeg")) else None
        outs[mode] = files
    assert sorted(outs["device"]) == sorted(outs["host"]) and len(outs["device"]) == 4 * (1 + 1 + 4) + 16
    for rel in outs["host"]:
        assert outs["device"][rel] == outs["host"][rel], rel
    std = json.loads(outs["device"]["json/a grey.json"])
    assert std["image_size"] == {"width": 1501, "height": 1003}


def test_decode_randomised_sweep_equals_cv2():
    """Sixty files in two batches (tests/jpeg_cases.py: random sizes 8 ... 900 px, qualities 20-100, restart
    intervals, greyscale and every colour subsampling, cv2's and Pillow's encoders with standard and optimised
    Huffman tables) — every page bit-identical to cv2.imdecode."""
    import jpeg_cases
    for batch, chunk in enumerate((128, 1024)):
        cases = jpeg_cases.sweep_files(2024 + batch, 30)
        files = [f for f, _ in cases]
        dec = ops.JpegDecoder(chunk_bytes=chunk, sync_rounds=8)
        pages = ops.decode_jpeg_files(files, dec)
        for (data, what), page, (w, h, c) in zip(cases, pages, dec.sizes):
            ref = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
            got = page[:, :c * w].cpu().numpy()
            assert np.array_equal(got, ref.reshape(h, 3 * w) if c == 3 else ref[..., 0]), what


def test_decode_of_damaged_entropy_data_is_safe_and_leaves_the_device_usable():
    """Files whose entropy-coded bytes were damaged (headers intact) go through the kernels: whatever pixels come out,
    nothing may trap, hang or write outside its page (canary rows behind every page stay untouched), a batch that
    never converges is reported (PageGeomError after 64 rounds), and a clean batch decodes bit-exactly afterwards."""
    import jpeg_cases
    files = jpeg_cases.scan_damaged_files(11, 48)
    blob, off = ops.pack_files(files)
    dec = ops.JpegDecoder(chunk_bytes=128, sync_rounds=4)
    sizes = dec.set_files(blob, off)
    pages = [torch.full((h + 3, ops.row_pitch(w, c)), 0xA5, dtype=torch.uint8, device="cuda") for w, h, c in sizes]
    dec.decode(blob.cuda(), [p[:h] for p, (w, h, c) in zip(pages, sizes)])
    torch.cuda.synchronize()
    try:
        dec.check()
    except PageGeomError as e:
        assert "converge" in str(e)
    torch.cuda.synchronize()
    for p, (w, h, c) in zip(pages, sizes):
        assert bool((p[h:] == 0xA5).all())
    clean = [f for f, _ in jpeg_cases.sweep_files(98, 6, max_side=300)]
    for data, page in zip(clean, ops.decode_jpeg_files(clean)):
        ref = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_ANYCOLOR)
        hh, ww = ref.shape[:2]
        got = page[:, :(ref.size // hh)].cpu().numpy()
        assert np.array_equal(got, ref.reshape(hh, -1))


def test_files_staged_in_write_combined_pinned_memory_decode_the_same():
    """ops.pack_files(write_combined=True): the blob lives in pg_pinned_alloc memory (the CPU only fills it; the parser
    reads the headers back from it), the asynchronous copy and the decode give cv2's pixels."""
    files = [_encode(_page(300, 400, 1)), _encode(_page(123, 77, 2), q=60, rst=3)]
    blob, off = ops.pack_files(files, write_combined=True)
    assert blob._pg_owner.write_combined  # (torch's is_pinned() does not see allocations of the library's own runtime;
    # the driver does: the copy below is a direct DMA)
    dec = ops.JpegDecoder()
    dec.set_files(blob, off)
    pages = dec.decode(blob.to("cuda", non_blocking=True))
    torch.cuda.synchronize()
    dec.check()
    for data, page, (w, h, c) in zip(files, pages, dec.sizes):
        ref = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_GRAYSCALE)
        assert np.array_equal(page[:, :w].cpu().numpy(), ref)
