"""GPU parity tests: every kernel, called through the C-ABI (libpagegeom.so via ctypes), against
the CPU oracle and the golden vectors produced by the unmodified reference scripts.

Bars (BASELINE.json north_star): kept / merged box sets, indices, medians and column outputs
bit-exact; translated coordinates bit-exact (<= 1e-4 px allowed); resized tile pixels within
+-1 LSB of cv2 (we additionally require > 99.9 % exact, and exact against the cv2 model)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from multimodal_embeddings_b200 import ops, reference_api as api, synth
from oracle import boxes as ob
from oracle import tiler as ot
from oracle.nms_fast import nms_pick_order_c

pytestmark = pytest.mark.gpu

PIXEL_TOL_LSB = 1       # north_star: +-1 LSB of the reference's cv2 resize
COORD_TOL_PX = 1e-4     # north_star: coordinates within 1e-4 px


def _tiles_u8(view_f16: torch.Tensor) -> np.ndarray:
    """fp16 [3,h,w] in [0,1] -> the uint8 value it encodes (fp16(v/255) is injective on 0..255)."""
    table = ot.u8_to_f16_unit(np.arange(256, dtype=np.uint8))
    arr = view_f16.cpu().numpy()
    idx = np.searchsorted(table, arr.ravel())
    idx = np.clip(idx, 0, 255)
    assert np.array_equal(table[idx], arr.ravel()), "tile holds values that are not fp16(v/255)"
    return idx.reshape(arr.shape).astype(np.uint8)


# ============================================================================== K1 tiler
@pytest.mark.parametrize("w,h,grids,imgsz,auto", [
    (1203, 907, [(2, 2)], 256, True),
    (1203, 907, [(1, 1), (2, 2), (3, 3)], 256, False),
    (997, 1501, [(3, 2)], 512, True),
    (333, 217, [(2, 2)], 256, True),        # upscale: y clamps at the borders
    (2800, 2100, [(1, 1)], 1024, True),     # one cfg3-sized tile, scale 2.73
])
def test_tiler_matches_cv2_oracle(w, h, grids, imgsz, auto):
    page = synth.page_pixels(w, h, seed=w * 7 + h)
    plan = ops.TilePlan(w, h, grids, 20.0, imgsz, 32, auto)
    pages = ops.upload_pages([page], plan)
    out = plan.run(pages)
    out_direct = plan.run(pages, direct=True)
    torch.cuda.synchronize()
    assert torch.equal(out, out_direct), "pipeline kernel != direct kernel"
    t = 0
    for rows, cols in grids:
        for cell in ot.split_array_into_grid(page, rows, cols, 20.0):
            info = plan.tiles[t]
            assert (info["x0"], info["y0"], info["x1"], info["y1"]) == cell["slice"]
            ref_u8 = ot.letterbox_tile_cv2(cell["image"], imgsz, 32, auto)
            got = plan.tile_view(out, 0, t)
            assert tuple(got.shape) == ref_u8.shape
            got_u8 = _tiles_u8(got)
            diff = np.abs(got_u8.astype(np.int16) - ref_u8.astype(np.int16))
            assert diff.max() <= PIXEL_TOL_LSB
            assert (diff == 0).mean() > 0.999
            # exact against the cv2 fixed-point model and the fp16 normalisation
            assert np.array_equal(got_u8, ot.letterbox_tile_model(cell["image"], imgsz, 32, auto))
            assert np.array_equal(got.cpu().numpy(), ot.u8_to_f16_unit(ref_u8)) or diff.max() == 1
            t += 1
    assert t == len(plan.tiles)


def test_tiler_batch_of_pages_and_padding_values():
    w, h = 1100, 640
    plan = ops.TilePlan(w, h, [(2, 3)], 20.0, 256, 32, False)  # square letterbox: real pad rows
    imgs = [synth.page_pixels(w, h, seed=s) for s in range(3)]
    pages = ops.upload_pages(imgs, plan)
    out = plan.run(pages)
    torch.cuda.synchronize()
    pad = ot.u8_to_f16_unit(np.array([114], np.uint8))[0]
    for p, img in enumerate(imgs):
        for t, cell in enumerate(ot.split_array_into_grid(img, 2, 3, 20.0)):
            got = plan.tile_view(out, p, t).cpu().numpy()
            ref = ot.u8_to_f16_unit(ot.letterbox_tile_cv2(cell["image"], 256, 32, False))
            assert np.array_equal(got, ref)
            info = plan.tiles[t]
            assert info["pad_t"] > 0 and np.all(got[:, : info["pad_t"], :] == pad)


def test_tiler_reference_default_grids_full_size():
    """The reference's default run (1_doclayout_bboxes.py:718 `--grids 2x2,3x3,4x4` after the full-page pass):
    30 tiles per 8000x6000 page, whose rows need four different shared-memory rings — one launch per ring.
    Pipeline == direct kernel on every tile, three tiles of three grids against cv2."""
    grids = [(1, 1), (2, 2), (3, 3), (4, 4)]
    plan = ops.TilePlan(8000, 6000, grids, 20.0)
    assert len(plan.tiles) == 30
    pages = ops.synth_pages(plan, 2, synth.PAGE_SEED0 + 7)
    out = plan.run(pages)
    out_d = plan.run(pages, direct=True)
    torch.cuda.synchronize()
    assert torch.equal(out, out_d)
    host = pages[0].cpu().numpy()[:, : 3 * 8000].reshape(6000, 8000, 3)
    first = {(1, 1): 0, (2, 2): 1, (3, 3): 5, (4, 4): 14}
    for (rows, cols), t0 in first.items():
        cells = ot.split_array_into_grid(host, rows, cols, 20.0)
        t = t0 + len(cells) - 1  # the last tile of the grid
        ref = ot.u8_to_f16_unit(ot.letterbox_tile_cv2(cells[-1]["image"]))
        assert np.array_equal(plan.tile_view(out, 0, t).cpu().numpy(), ref), (rows, cols)


def test_tiler_full_size_page_cross_check():
    """cfg3 shape (8000x6000, 4x4): pipeline == direct kernel everywhere, and two tiles
    against cv2.  Size-independent property: two independent kernels agree bit-for-bit."""
    plan = ops.TilePlan(8000, 6000, [(4, 4)], 20.0)
    pages = ops.synth_pages(plan, 2, synth.PAGE_SEED0)
    out = plan.run(pages)
    out_d = plan.run(pages, direct=True)
    torch.cuda.synchronize()
    assert torch.equal(out, out_d)
    host = pages[1].cpu().numpy()[:, : 3 * 8000].reshape(6000, 8000, 3)
    assert host.std() > 20  # generator produced structure, not a constant page
    cells = ot.split_array_into_grid(host, 4, 4, 20.0)
    for t in (5, 15):
        ref = ot.u8_to_f16_unit(ot.letterbox_tile_cv2(cells[t]["image"]))
        assert np.array_equal(plan.tile_view(out, 1, t).cpu().numpy(), ref)
    # regenerating the same page index on its own gives the same bytes (sharding independence)
    again = ops.synth_pages(plan, 1, synth.PAGE_SEED0, first_page=1)
    assert torch.equal(again[0], pages[1])


@pytest.mark.parametrize("max_row_bytes", [None, 700, 190])
def test_tiler_random_shapes_and_imgsz_sweep(max_row_bytes, monkeypatch):
    """Random page shapes / grids / overlaps / letterbox sizes: every instantiation of the pipeline kernel
    (1, 2 and 4 pixel-pair iterations, imgsz 320..2048) against the direct kernel everywhere and against
    cv2 on two tiles per case.  Run again with the column-chunk threshold forced down (a plan-time knob), so
    that these small tiles are cut into many chunks: chunk borders, rebased x tables, pad-only chunks."""
    if max_row_bytes is not None:
        monkeypatch.setenv("PG_TILER_MAX_ROW_BYTES", str(max_row_bytes))
    rng = np.random.default_rng(77)
    cases = [(3000, 2200, 1, 1, 0.0, 2048, True), (2100, 1500, 2, 1, 35.0, 1280, False), (801, 613, 1, 3, 20.0, 320, True)]
    for _ in range(9):
        w, h = int(rng.integers(200, 2600)), int(rng.integers(200, 2600))
        rows, cols = int(rng.integers(1, 5)), int(rng.integers(1, 5))
        ov = float(rng.choice([0.0, 10.0, 20.0, 45.0]))
        imgsz = int(rng.choice([320, 640, 1024, 1280]))
        cases.append((w, h, rows, cols, ov, imgsz, bool(rng.integers(0, 2))))
    for (w, h, rows, cols, ov, imgsz, auto) in cases:
        page = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        try:
            plan = ops.TilePlan(w, h, [(rows, cols)], ov, imgsz, 32, auto)
        except Exception as e:  # tiny tiles may letterbox to a zero-sized image: must be reported, not crash
            assert "unsupported" in str(e) or "empty tile" in str(e), e
            continue
        pages = ops.upload_pages([page], plan)
        out = plan.run(pages)
        out_d = plan.run(pages, direct=True)
        torch.cuda.synchronize()
        assert torch.equal(out, out_d), (w, h, rows, cols, ov, imgsz, auto)
        cells = ot.split_array_into_grid(page, rows, cols, ov)
        for t in {0, len(cells) - 1}:
            ref = ot.u8_to_f16_unit(ot.letterbox_tile_cv2(cells[t]["image"], imgsz, 32, auto))
            assert np.array_equal(plan.tile_view(out, 0, t).cpu().numpy(), ref), (w, h, rows, cols, ov, imgsz, auto, t)


@pytest.mark.parametrize("w,h", [(14000, 300), (40000, 64)])
def test_tiler_very_wide_tile_is_cut_into_column_chunks(w, h):
    """A 1x1 grid on a 14000 / 40000 px wide page: 42 / 120 KB per source row.  The plan cuts such a tile into
    chunks of output columns that each stage only the source span they sample (<= 9 KB), so the row length is
    not a limit any more; the result equals cv2 and the (unchunked) direct kernel."""
    page = np.random.default_rng(5).integers(0, 256, (h, w, 3), dtype=np.uint8)
    plan = ops.TilePlan(w, h, [(1, 1)], 20.0, 1024, 32, True)
    pages = ops.upload_pages([page], plan)
    out = plan.run(pages)
    out_d = plan.run(pages, direct=True)
    torch.cuda.synchronize()
    assert torch.equal(out, out_d)
    ref = ot.u8_to_f16_unit(ot.letterbox_tile_cv2(page, 1024, 32, True))
    assert np.array_equal(plan.tile_view(out, 0, 0).cpu().numpy(), ref)


def test_tiler_chunks_made_only_of_pad_columns():
    """A 7000x35000 strip letterboxed into a 1024 square (auto=False): 205 real columns between 409/410 pad
    columns, 21 KB source rows -> sixteen 64-column chunks, most of them nothing but pad (they stage one
    pixel and never sample it)."""
    w, h = 7000, 35000
    plan = ops.TilePlan(w, h, [(1, 1)], 20.0, 1024, 32, False)
    info = plan.tiles[0]
    assert (info["out_w"], info["out_h"], info["new_w"]) == (1024, 1024, 205) and info["pad_l"] >= 384
    pages = ops.synth_pages(plan, 1, synth.PAGE_SEED0 + 3)
    out = plan.run(pages)
    out_d = plan.run(pages, direct=True)
    torch.cuda.synchronize()
    assert torch.equal(out, out_d)
    host = pages[0].cpu().numpy()[:, : 3 * w].reshape(h, w, 3)
    ref = ot.u8_to_f16_unit(ot.letterbox_tile_cv2(host, 1024, 32, False))
    assert np.array_equal(plan.tile_view(out, 0, 0).cpu().numpy(), ref)


def test_tiler_heterogeneous_batch_one_launch():
    """Pages of different sizes (incl. a repeated size and an upscaled one) tiled by ONE launch."""
    sizes = [(1203, 907), (640, 480), (997, 1501), (1203, 907), (333, 217), (2000, 1400)]
    imgs = [synth.page_pixels(w, h, seed=100 + i) for i, (w, h) in enumerate(sizes)]
    batch = ops.TileBatch(sizes, [(1, 1), (2, 2)], 20.0, imgsz=256)
    assert len(batch.plans) == 5
    pages = batch.alloc_pages()
    for t, img, (w, h) in zip(pages, imgs, sizes):
        host = np.zeros((h, t.shape[1]), np.uint8)
        host[:, : 3 * w] = img.reshape(h, 3 * w)
        t.copy_(torch.from_numpy(host))
    batch.bind(pages)
    batch.run()
    batch.run()  # the work counter re-arms itself between launches
    torch.cuda.synchronize()
    for p, img in enumerate(imgs):
        t = 0
        for rows, cols in [(1, 1), (2, 2)]:
            for cell in ot.split_array_into_grid(img, rows, cols, 20.0):
                ref = ot.u8_to_f16_unit(ot.letterbox_tile_cv2(cell["image"], 256))
                assert np.array_equal(batch.tile_view(p, t).cpu().numpy(), ref), (p, t)
                t += 1
    alg = sum(3 * w * h for w, h in sizes) + 2 * sum(batch.plan_of(i).out_elems for i in range(len(sizes)))
    assert batch.algorithmic_bytes == alg


def test_split_image_into_grid_mirror():
    page = synth.page_pixels(900, 700, seed=5)
    cells = api.split_image_into_grid(page, 2, 2, 20.0, imgsz=256)
    ref = ot.split_array_into_grid(page, 2, 2, 20.0)
    assert len(cells) == 4
    for c, r in zip(cells, ref):
        assert c["coordinates"] == r["coordinates"] and (c["row"], c["col"]) == (r["row"], r["col"])
        assert np.array_equal(c["tensor"].cpu().numpy(), ot.u8_to_f16_unit(ot.letterbox_tile_cv2(r["image"], 256)))
    assert api.split_image_into_grid("/nonexistent/page.png", 2, 2, 20.0) == []


# ============================================================================== K2 edge filter
def test_edge_filter_golden_cases():
    for case in load_golden("stage2_filter.json.gz"):
        w, h, thr = case["width"], case["height"], case["threshold"]
        counts = [len(b) for b in case["boxes_original"]]
        boxes = np.concatenate([np.asarray(b, np.float64).reshape(-1, 4) for b in case["boxes_original"]])
        box_cell = np.repeat(np.arange(len(counts), dtype=np.int32), counts)
        cells = np.asarray([ob.cell_tuple(c, w, h) for c in case["cell_coordinates"]], np.float64)
        _, keep, kept_idx, n_kept = ops.edge_filter(boxes, box_cell, cells, [[w, h]], [0, len(boxes)], thr,
                                                    boxes_are_local=False)
        expect = np.concatenate([np.asarray(k, np.int64) + off
                                 for k, off in zip(case["kept"], np.cumsum([0] + counts[:-1]))])
        nk = int(n_kept[0].item())
        assert nk == len(expect)
        assert np.array_equal(kept_idx[:nk].cpu().numpy(), expect)
        assert np.array_equal(np.nonzero(keep.cpu().numpy())[0], expect)


def test_edge_filter_translation_and_multi_page_batch():
    dets = [synth.page_detections(w, h, r, c, 20.0, n, seed)
            for (w, h, r, c, n, seed) in [(8000, 6000, 4, 4, 10000, 1), (3801, 5601, 2, 2, 2000, 2),
                                          (640, 480, 2, 2, 0, 3), (2778, 4187, 3, 3, 1234, 4)]]
    boxes = np.concatenate([d["boxes_local"] for d in dets])
    cell_base = np.cumsum([0] + [len(d["cells"]) for d in dets])
    box_cell = np.concatenate([d["box_cell"] + cell_base[i] for i, d in enumerate(dets)]).astype(np.int32)
    cells = np.concatenate([d["cells"] for d in dets])
    page_wh = [[d["width"], d["height"]] for d in dets]
    page_off = np.cumsum([0] + [len(d["boxes_local"]) for d in dets])
    bp, keep, kept_idx, n_kept = ops.edge_filter(boxes, box_cell, cells, page_wh, page_off, 10)
    bp, kept_idx, n_kept = bp.cpu().numpy(), kept_idx.cpu().numpy(), n_kept.cpu().numpy()
    for i, d in enumerate(dets):
        sl = slice(page_off[i], page_off[i + 1])
        ref_page = np.asarray([ot.translate_boxes([b.tolist()], dict(zip(("x_start", "y_start"), d["cells"][c][:2])))[0]
                               for b, c in zip(d["boxes_local"], d["box_cell"])]).reshape(-1, 4)
        assert np.array_equal(bp[sl], ref_page)   # bit-exact (tolerance would be COORD_TOL_PX)
        assert np.abs(bp[sl] - ref_page).max(initial=0) <= COORD_TOL_PX
        ref_keep = [j for j, (b, c) in enumerate(zip(ref_page, d["box_cell"]))
                    if not ob.touches_internal_edge(b, d["cells"][c], d["width"], d["height"], 10)]
        assert n_kept[i] == len(ref_keep)
        assert np.array_equal(kept_idx[page_off[i]: page_off[i] + n_kept[i]] - page_off[i], ref_keep)
    assert n_kept[2] == 0


def test_filter_grid_info_mirror_and_single_box():
    case = load_golden("stage2_filter.json.gz")[1]
    gi = {"original_image_path": "unused.png", "grid_config": {"rows": case["rows"], "cols": case["cols"],
                                                               "overlap_percentage": case["overlap"]},
          "cells": [{"cell_path": "a", "cell_json_path": "b", "cell_coordinates": cc, "row": 1, "col": 1,
                     "regions": {"boxes": bo, "boxes_original": bo, "classes": [1.0] * len(bo),
                                 "scores": [0.5] * len(bo), "class_names": ["plain_text"] * len(bo)}}
                    for cc, bo in zip(case["cell_coordinates"], case["boxes_original"])]}
    out = api.filter_grid_info(gi, case["threshold"], image_size=(case["width"], case["height"]))
    for cell, bo, kept in zip(out["cells"], case["boxes_original"], case["kept"]):
        assert cell["regions"]["boxes_original"] == [bo[i] for i in kept]
    assert api.filter_grid_info(gi, 10) is None  # image missing -> None like 2_edge_box_filter.py:199-203
    cc = case["cell_coordinates"][0]
    for i in list(range(10)):
        b = case["boxes_original"][0][i]
        assert api.is_box_touching_internal_edge(b, cc, case["width"], case["height"], case["threshold"]) == \
            (i not in case["kept"][0])


# ============================================================================== K3 NMS merge
def test_nms_golden_cases():
    g = load_golden("stage3_nms.npz")
    for name in g["cases"]:
        b, s, c = g[f"{name}_boxes"], g[f"{name}_scores"], g[f"{name}_classes"]
        thr = float(g[f"{name}_thr"][0])
        got = api.nms_keep_indices(b, s, c, thr)
        assert got == g[f"{name}_kept"].tolist(), name
    assert api.apply_non_max_suppression([], [], [], [], 0.5) == ([], [], [], [])


def test_nms_idempotent_on_reference_outputs(f1_pages):
    """The reference's committed stage-3 outputs, all 19 pages in ONE batched launch."""
    boxes = np.concatenate([np.asarray(p["boxes"], np.float64) for p in f1_pages])
    scores = np.concatenate([np.asarray(p["scores"], np.float64) for p in f1_pages])
    classes = np.concatenate([np.asarray(p["classes"], np.float64) for p in f1_pages])
    off = np.cumsum([0] + [len(p["boxes"]) for p in f1_pages])
    kept, n_kept, ws = ops.nms_merge(boxes, scores, classes, off, 0.5)
    kept, n_kept = kept.cpu().numpy(), n_kept.cpu().numpy()
    assert ws.stats()["status"] == 0
    for i in range(len(f1_pages)):
        assert n_kept[i] == off[i + 1] - off[i]
        assert np.array_equal(kept[off[i]: off[i + 1]], np.arange(off[i], off[i + 1]))
    fb, fs, fc, fn = api.apply_non_max_suppression(f1_pages[0]["boxes"], f1_pages[0]["scores"],
                                                   f1_pages[0]["classes"], f1_pages[0]["class_names"], 0.5)
    assert fb == f1_pages[0]["boxes"] and fn == f1_pages[0]["class_names"]


@pytest.mark.parametrize("thr", [0.5, 0.3, 0.0, -1.0])
def test_nms_synthetic_batch_vs_oracle(thr):
    cfgs = [(8000, 6000, 4, 4, 10000, 31), (3801, 5601, 2, 2, 2000, 32), (2000, 2000, 2, 2, 1, 33),
            (2778, 4187, 3, 3, 3333, 34), (640, 480, 1, 1, 0, 35), (4000, 5443, 2, 2, 33, 36)]
    bs, ss, cs, off = [], [], [], [0]
    for (w, h, r, c, n, seed) in cfgs:
        d = synth.page_detections(w, h, r, c, 20.0, n, seed)
        bs.append(d["boxes_local"] + d["cells"][d["box_cell"]][:, [0, 1, 0, 1]] if n else np.zeros((0, 4)))
        ss.append(d["scores"]); cs.append(d["classes"]); off.append(off[-1] + n)
    boxes, scores, classes = np.concatenate(bs), np.concatenate(ss), np.concatenate(cs)
    scores[100:140] = scores[100]  # exact score ties inside page 0
    # thr < 0 makes every same-class pair a suppressor: needs the dense candidate bound
    ws = ops.NmsWorkspace(len(boxes), len(cfgs), pairs_per_block=320 if thr < 0 else 64)
    kept, n_kept, ws = ops.nms_merge(boxes, scores, classes, off, thr, max_boxes_per_page=10000, workspace=ws)
    kept, n_kept = kept.cpu().numpy(), n_kept.cpu().numpy()
    st = ws.stats()
    assert st["status"] == 0 and st["rounds"] >= 1
    for i in range(len(cfgs)):
        sl = slice(off[i], off[i + 1])
        ref = nms_pick_order_c(boxes[sl], scores[sl], classes[sl], thr)
        assert n_kept[i] == len(ref), (i, thr)
        assert np.array_equal(kept[off[i]: off[i] + n_kept[i]] - off[i], ref), (i, thr)


def test_nms_signed_zero_scores_tie_by_position():
    """-0.0 and +0.0 scores compare EQUAL in the reference (`max()` / `index`, 3_combine_grids.py:112), so the
    earlier pooled position wins whatever the sign bit; the emitted pick order must follow the same rule as the
    suppression ranking (advisor finding: the sort key used to order -0.0 after +0.0)."""
    rng = np.random.default_rng(77)
    n = 600
    xy = rng.uniform(0, 900, (n, 2))
    boxes = np.concatenate([xy, xy + rng.uniform(30, 120, (n, 2))], 1)
    classes = rng.integers(0, 2, n).astype(np.float64)
    scores = np.where(rng.random(n) < 0.5, -0.0, 0.0)
    scores[::7] = rng.choice([0.25, -0.25], len(scores[::7]))
    assert np.signbit(scores).any() and (scores == 0).sum() > 100
    for thr in (0.5, 0.1):
        ref = nms_pick_order_c(boxes, scores, classes, thr)
        from oracle import boxes as ob
        assert list(ref) == ob.nms_pick_order(boxes.tolist(), scores.tolist(), classes.tolist(), thr)
        assert api.nms_keep_indices(boxes, scores, classes, thr) == list(ref)


def test_nms_dense_stress_100k_boxes():
    """cfg4: 100k boxes on one page (the reference needs ~10^3 s here; the C oracle ~10 s)."""
    d = synth.page_detections(8000, 6000, 4, 4, 20.0, 100000, 77, dups=6)
    boxes = d["boxes_local"] + d["cells"][d["box_cell"]][:, [0, 1, 0, 1]]
    kept, n_kept, ws = ops.nms_merge(boxes, d["scores"], d["classes"], [0, 100000], 0.5, max_boxes_per_page=100000)
    k = int(n_kept[0].item())
    st = ws.stats()
    assert st["status"] == 0
    ref = nms_pick_order_c(boxes, d["scores"], d["classes"], 0.5)
    assert k == len(ref) and np.array_equal(kept[:k].cpu().numpy(), ref)
    # size-independent property: NMS of the survivors is the identity
    kept2, n2, _ = ops.nms_merge(boxes[ref], d["scores"][ref], d["classes"][ref], [0, len(ref)], 0.5)
    assert int(n2[0].item()) == len(ref) and np.array_equal(kept2[: len(ref)].cpu().numpy(), np.arange(len(ref)))


def _pooled(cfgs, dups=3):
    bs, ss, cs, off = [], [], [], [0]
    for (w, h, r, c, n, seed) in cfgs:
        d = synth.page_detections(w, h, r, c, 20.0, n, seed, dups=dups)
        bs.append(d["boxes_local"] + d["cells"][d["box_cell"]][:, [0, 1, 0, 1]] if n else np.zeros((0, 4)))
        ss.append(d["scores"]); cs.append(d["classes"]); off.append(off[-1] + n)
    return np.concatenate(bs), np.concatenate(ss), np.concatenate(cs), off


@pytest.mark.parametrize("name,cfgs,dups,oracle_pages", [
    # six pages incl. an empty and a one-box page: width 8, every page below 8192 survivors (rank-0 emit path)
    ("mixed", [(8000, 6000, 4, 4, 10000, 31), (3801, 5601, 2, 2, 2000, 32), (2000, 2000, 2, 2, 1, 33),
               (2778, 4187, 3, 3, 3333, 34), (640, 480, 1, 1, 0, 35), (4000, 5443, 2, 2, 33, 36)], 3, [0, 2, 3, 4]),
    # one page whose survivors need two 8192-element chunks
    ("two_chunks", [(8000, 6000, 4, 4, 30000, 41)], 3, [0]),
    # three large pages side by side, 8 CTAs each
    ("three_pages", [(8000, 6000, 4, 4, 40000, 42), (8000, 6000, 4, 4, 50000, 43), (6000, 8000, 3, 3, 35000, 44)], 6, []),
    # > 65536 survivors: 16 chunks over 8 CTAs (two chunks per CTA); a lattice of disjoint boxes plus duplicates
    ("sixteen_chunks", None, 1, []),
    # 30 pages: the cluster narrows to 4 CTAs so that the launch still fits the SMs
    ("width4", [(4000, 3000, 2, 2, 3000 + 37 * i, 50 + i) for i in range(30)], 3, [7]),
])
def test_nms_cluster_kernels_equal_single_cta_kernels(name, cfgs, dups, oracle_pages, monkeypatch):
    """Resolve/emit on one thread-block cluster per page (taken for pages of >= 32768 boxes; forced here through
    PG_NMS_CLUSTER_MIN_BOXES) against the one-CTA-per-page kernels, and against the oracle on some pages."""
    if cfgs is None:
        rng = np.random.default_rng(46)
        gx, gy = np.meshgrid(np.arange(300) * 26.0, np.arange(300) * 20.0)
        lat = np.stack([gx.ravel(), gy.ravel(), gx.ravel() + 18.0, gy.ravel() + 12.0], 1)
        dup = lat[rng.choice(len(lat), 30000, replace=False)] + rng.normal(0, 0.7, (30000, 4))
        boxes = np.concatenate([lat, dup]).astype(np.float32).astype(np.float64)
        boxes = boxes[rng.permutation(len(boxes))]
        scores = rng.uniform(0.1, 0.99, len(boxes)).astype(np.float32).astype(np.float64)
        scores[5000:5600] = scores[5000]
        classes = rng.integers(0, 3, len(boxes)).astype(np.float64)
        off, cfgs = [0, len(boxes)], [None]
    else:
        boxes, scores, classes, off = _pooled(cfgs, dups)
    if name == "mixed":
        scores[100:140] = scores[100]
    mx = int(np.diff(off).max())
    out = {}
    for knob in ("0", "1"):
        monkeypatch.setenv("PG_NMS_CLUSTER_MIN_BOXES", knob)
        ws = ops.NmsWorkspace(len(boxes), len(cfgs), pairs_per_block=96)
        kept, n_kept, ws = ops.nms_merge(boxes, scores, classes, off, 0.5, max_boxes_per_page=mx, workspace=ws)
        assert ws.stats()["status"] == 0
        out[knob] = (kept.cpu().numpy(), n_kept.cpu().numpy(), ws.stats()["rounds"])
    (k0, n0, r0), (k1, n1, r1) = out["0"], out["1"]
    assert np.array_equal(n0, n1) and r0 == r1
    for i in range(len(cfgs)):
        assert np.array_equal(k0[off[i]: off[i] + n0[i]], k1[off[i]: off[i] + n1[i]]), (name, i)
    if name == "sixteen_chunks":
        assert n1[0] > 65536
    if name == "two_chunks":
        assert 8192 < n1[0] <= 16384
    for i in oracle_pages:
        sl = slice(off[i], off[i + 1])
        ref = nms_pick_order_c(boxes[sl], scores[sl], classes[sl], 0.5)
        assert n1[i] == len(ref) and np.array_equal(k1[off[i]: off[i] + n1[i]] - off[i], ref), (name, i)


@pytest.mark.parametrize("cluster", ["0", "1"])
def test_nms_sub_cell_digit_of_the_spatial_sort_changes_nothing(cluster, monkeypatch):
    """The third, finer digit of the spatial sort (crowded pages; forced on every page / disabled here through
    PG_NMS_FINE_MIN) is a pure ordering heuristic: same kept boxes in the same pick order, against the oracle, on the
    one-CTA and on the cluster kernels; on a crowded page it must cut the pair tests."""
    cfgs = [(8000, 6000, 4, 4, 30000, 61), (3801, 5601, 2, 2, 2000, 62), (2000, 2000, 2, 2, 1, 63), (640, 480, 1, 1, 0, 64),
            (2778, 4187, 3, 3, 9000, 65)]
    boxes, scores, classes, off = _pooled(cfgs, 4)
    scores[200:260] = scores[200]
    mx = int(np.diff(off).max())
    monkeypatch.setenv("PG_NMS_CLUSTER_MIN_BOXES", cluster)
    out = {}
    for fine in ("0", "1000000000"):
        monkeypatch.setenv("PG_NMS_FINE_MIN", fine)
        ws = ops.NmsWorkspace(len(boxes), len(cfgs), pairs_per_block=128)
        kept, n_kept, ws = ops.nms_merge(boxes, scores, classes, off, 0.5, max_boxes_per_page=mx, workspace=ws)
        st = ws.stats()
        assert st["status"] == 0
        out[fine] = (kept.cpu().numpy(), n_kept.cpu().numpy(), st)
    (k0, n0, st0), (k1, n1, st1) = out["0"], out["1000000000"]
    assert np.array_equal(n0, n1)
    for i in range(len(cfgs)):
        sl = slice(off[i], off[i + 1])
        assert np.array_equal(k0[off[i]: off[i] + n0[i]], k1[off[i]: off[i] + n1[i]]), i
        ref = nms_pick_order_c(boxes[sl], scores[sl], classes[sl], 0.5)
        assert n0[i] == len(ref) and np.array_equal(k0[off[i]: off[i] + n0[i]] - off[i], ref), i
    assert st0["box_pairs_tested"] < st1["box_pairs_tested"]


def test_nms_cluster_kernels_report_workspace_overflow(monkeypatch):
    monkeypatch.setenv("PG_NMS_CLUSTER_MIN_BOXES", "1")
    n = 3000
    boxes = np.tile(np.array([[10.0, 10.0, 50.0, 60.0]]), (n, 1))
    scores = np.linspace(0.9, 0.1, n)
    classes = (np.arange(n) % 3).astype(np.float64)
    ws = ops.NmsWorkspace(n, 1, pairs_per_block=128)
    kept, n_kept, ws = ops.nms_merge(boxes, scores, classes, [0, n], 0.5, max_boxes_per_page=n, workspace=ws)
    assert ws.stats()["status"] == 0 and kept[: int(n_kept[0].item())].cpu().numpy().tolist() == [0, 1, 2]
    small = ops.NmsWorkspace(n, 1, pairs_per_block=2)
    _, n_kept2, small = ops.nms_merge(boxes, scores, classes, [0, n], 0.5, max_boxes_per_page=n, workspace=small)
    assert small.stats()["status"] == 3 and int(n_kept2[0].item()) == -1


def test_nms_adversarial_identical_boxes_and_workspace_overflow():
    n = 3000
    boxes = np.tile(np.array([[10.0, 10.0, 50.0, 60.0]]), (n, 1))
    scores = np.linspace(0.9, 0.1, n)
    classes = (np.arange(n) % 3).astype(np.float64)
    ws = ops.NmsWorkspace(n, 1, pairs_per_block=128)
    kept, n_kept, ws = ops.nms_merge(boxes, scores, classes, [0, n], 0.5, workspace=ws)
    assert ws.stats()["status"] == 0
    assert kept[: int(n_kept[0].item())].cpu().numpy().tolist() == [0, 1, 2]
    small = ops.NmsWorkspace(n, 1, pairs_per_block=2)   # every block pair is a candidate -> overflow
    _, n_kept2, small = ops.nms_merge(boxes, scores, classes, [0, n], 0.5, workspace=small)
    assert small.stats()["status"] == 3 and int(n_kept2[0].item()) == -1  # reported, never silently wrong


def test_per_tile_nms_matches_torchvision():
    """the per-tile class-agnostic float32 NMS of 1_doclayout_bboxes.py:217-225."""
    g = load_golden("stage1_tile_nms.npz")
    boxes_list = [g[f"{n}_boxes"] for n in g["cases"]]
    scores_list = [g[f"{n}_scores"] for n in g["cases"]]
    for name, b, s in zip(g["cases"], boxes_list, scores_list):   # one tile per call, torchvision's signature
        keep = api.nms(torch.tensor(b), torch.tensor(s), float(g[f"{name}_thr"][0]))
        assert keep.dtype == torch.int64 and keep.tolist() == g[f"{name}_keep"].tolist(), name
    # many tiles in one launch (same threshold), against the float32 restatement and live torchvision
    import torchvision
    rng = np.random.default_rng(8)
    tiles_b, tiles_s = [], []
    for t in range(16):
        n = int(rng.choice([0, 1, 50, 700, 2500]))
        d = synth.page_detections(2800, 2100, 1, 1, 20.0, n, 900 + t, dups=3)
        tiles_b.append(d["boxes_local"].astype(np.float32))
        tiles_s.append(d["scores"].astype(np.float32))
    keeps = api.nms_per_tile(tiles_b, tiles_s, 0.45)
    for b, s, k in zip(tiles_b, tiles_s, keeps):
        ref = torchvision.ops.nms(torch.tensor(b).reshape(-1, 4), torch.tensor(s), 0.45).tolist() if len(s) else []
        assert k.tolist() == ref
        assert k.tolist() == ob.nms_torchvision_f32(b, s, 0.45)


# ============================================================================== K4 / K5
def test_width_median_and_columns_on_reference_outputs(f1_pages, f4):
    boxes = np.concatenate([np.asarray(p["boxes"], np.float64) for p in f1_pages])
    scores = np.concatenate([np.asarray(p["scores"], np.float64) for p in f1_pages])
    names = sum([p["class_names"] for p in f1_pages], [])
    flags = api._flags_from_names(names)
    off = np.cumsum([0] + [len(p["boxes"]) for p in f1_pages])
    wh = [[p["image_size"]["width"], p["image_size"]["height"]] for p in f1_pages]
    med, nb = ops.width_median(boxes, flags, off, wh, 0.2)
    centers, widths, n_cols = ops.column_peaks(boxes, flags, scores, off, wh, med, 0.3)
    med, nb, centers, widths, n_cols = (t.cpu().numpy() for t in (med, nb, centers, widths, n_cols))
    for i, g in enumerate(f4):
        assert med[i] == g["median_width"], g["name"]
        assert nb[i] == g["n_bins"]
        k = n_cols[i]
        assert [float(x) for x in centers[i, :k]] == g["column_centers"], g["name"]
        assert [float(x) for x in widths[i, :k]] == g["column_widths"], g["name"]
    # the single-page mirrors return the reference's shapes
    p, g = f1_pages[12], f4[12]
    m, nbins = api.median_plain_text_width(p["boxes"], p["class_names"], p["image_size"]["width"])
    assert float(m) == g["median_width"] and nbins == g["n_bins"]
    c, w = api.find_column_centers(p["boxes"], p["class_names"], p["scores"], p["image_size"]["width"],
                                   p["image_size"]["height"], m, 0.3)
    assert [float(x) for x in c] == g["column_centers"] and [float(x) for x in w] == g["column_widths"]


def test_width_median_and_columns_synthetic_goldens():
    for g in load_golden("stage45_synth.json"):
        det = synth.page_detections(g["width"], g["height"], g["rows"], g["cols"], 20.0, g["n"], g["seed"])
        b = det["boxes_local"] + det["cells"][det["box_cell"]][:, [0, 1, 0, 1]]
        k = nms_pick_order_c(b, det["scores"], det["classes"], 0.5)
        b, s, c = b[k], det["scores"][k], det["classes"][k]
        names = synth.class_names_of(c)
        m, nbins = api.median_plain_text_width(b.tolist(), names, g["width"], g["min_margin_percent"])
        assert float(m) == g["median_width"] and nbins == g["n_bins"], g
        cc, cw = api.find_column_centers(b.tolist(), names, s.tolist(), g["width"], g["height"], m,
                                         g["min_confidence"])
        assert [float(x) for x in cc] == g["column_centers"], g
        assert [float(x) for x in cw] == g["column_widths"], g


def test_width_median_edge_cases():
    # no plain_text at all -> 0 / 0 bins; single width; even count -> mean of the two middles
    assert api.median_plain_text_width([[0, 0, 10, 10]], ["title"], 1000) == (0, 0)
    assert api.median_plain_text_width([], [], 1000) == (0, 0)
    m, nb = api.median_plain_text_width([[0, 0, 10, 10], [0, 0, 500, 10]], ["plain_text"] * 2, 1000)
    assert (float(m), nb) == (255.0, 2)
    rng = np.random.default_rng(3)
    for trial in range(20):
        n = int(rng.integers(1, 400))
        w = rng.uniform(5, 900, n).astype(np.float32).astype(np.float64)
        boxes = np.stack([np.zeros(n), np.zeros(n), w, np.ones(n)], 1)
        pct = float(rng.choice([0.2, 0.01, 2.0, 0.0, -0.5]))
        names = ["plain_text" if rng.random() < 0.8 else "title" for _ in range(n)]
        ref_m, ref_nb = ob.median_width(boxes.tolist(), names, 4000, pct)
        m, nb = api.median_plain_text_width(boxes.tolist(), names, 4000, pct)
        assert (float(m), nb) == (float(ref_m), ref_nb), (trial, pct)


def test_median_and_columns_random_page_shapes_one_batch():
    """40 pages of very different widths (density resolution 1..9, 500..2000 bins, windows 5..101) and box
    counts, all in ONE launch of K4 and K5, against the oracle (scipy find_peaks, np.convolve)."""
    rng = np.random.default_rng(2024)
    pages = []
    for t in range(40):
        w = int(rng.choice([517, 999, 1000, 1001, 1999, 2000, 2778, 3631, 4029, 6100, 7934, 9001]))
        h = int(rng.integers(400, 6000))
        n = int(rng.choice([0, 1, 7, 60, 400, 1500, 3000]))
        d = synth.page_detections(w, h, 2, 2, 20.0, n, 7000 + t)
        b = d["boxes_local"] + d["cells"][d["box_cell"]][:, [0, 1, 0, 1]] if n else np.zeros((0, 4))
        keep = nms_pick_order_c(b, d["scores"], d["classes"], 0.5) if n else np.zeros(0, np.int64)
        pages.append((w, h, b[keep], d["scores"][keep], d["classes"][keep]))
    boxes = np.concatenate([p[2] for p in pages])
    scores = np.concatenate([p[3] for p in pages])
    classes = np.concatenate([p[4] for p in pages])
    off = np.cumsum([0] + [len(p[2]) for p in pages])
    wh = [[p[0], p[1]] for p in pages]
    flags = ops.class_flags(classes)
    for pct, conf in ((0.2, 0.3), (1.5, 0.6)):
        med, nb = ops.width_median(boxes, flags, off, wh, pct)
        centers, widths, n_cols = ops.column_peaks(boxes, flags, scores, off, wh, med, conf, max_cols=128)
        med, nb, centers, widths, n_cols = (t.cpu().numpy() for t in (med, nb, centers, widths, n_cols))
        n_with_cols = 0
        for i, (w, h, b, s, c) in enumerate(pages):
            names = synth.class_names_of(c)
            ref_m, ref_nb = ob.median_width(b.tolist(), names, w, pct)
            assert (med[i], nb[i]) == (float(ref_m), ref_nb), (i, w, len(b))
            rc, rw = ob.column_centers(b.tolist(), names, s.tolist(), w, h, ref_m, conf) if ref_m > 0 else ([], [])
            k = n_cols[i]
            assert k >= 0
            assert [float(x) for x in centers[i, :k]] == [float(x) for x in rc], (i, w, len(b))
            assert [float(x) for x in widths[i, :k]] == [float(x) for x in rw], (i, w, len(b))
            n_with_cols += k > 0
        assert n_with_cols >= 15


def test_corpus_histograms_match_numpy(f1_pages, f4):
    """K6 inputs: the integer histograms K4/K5 accumulate on request (1-px plain_text widths, per-mille
    column centres) equal numpy's on the reference's 19 pages; accumulating twice doubles them exactly."""
    from multimodal_embeddings_b200._lib import PG_COL_HIST_BINS, PG_WIDTH_HIST_BINS
    boxes = np.concatenate([np.asarray(p["boxes"], np.float64) for p in f1_pages])
    scores = np.concatenate([np.asarray(p["scores"], np.float64) for p in f1_pages])
    names = sum([p["class_names"] for p in f1_pages], [])
    flags = api._flags_from_names(names)
    off = np.cumsum([0] + [len(p["boxes"]) for p in f1_pages])
    wh = [[p["image_size"]["width"], p["image_size"]["height"]] for p in f1_pages]
    wh_hist = torch.zeros(PG_WIDTH_HIST_BINS, dtype=torch.int32, device="cuda")
    col_hist = torch.zeros(PG_COL_HIST_BINS, dtype=torch.int32, device="cuda")
    for _ in range(2):
        med, _ = ops.width_median(boxes, flags, off, wh, 0.2, width_hist=wh_hist)
        ops.column_peaks(boxes, flags, scores, off, wh, med, 0.3, col_hist=col_hist)
    widths = (boxes[:, 2] - boxes[:, 0])[np.asarray(names) == "plain_text"]
    ref_w = np.bincount(np.clip(widths.astype(np.int64), 0, PG_WIDTH_HIST_BINS - 1), minlength=PG_WIDTH_HIST_BINS)
    assert np.array_equal(wh_hist.cpu().numpy(), 2 * ref_w)
    ref_c = np.zeros(PG_COL_HIST_BINS, np.int64)
    for p, g in zip(f1_pages, f4):
        for c in g["column_centers"]:
            ref_c[min(PG_COL_HIST_BINS - 1, int(c) * 1000 // p["image_size"]["width"])] += 1
    assert np.array_equal(col_hist.cpu().numpy(), 2 * ref_c) and ref_c.sum() == sum(len(g["column_centers"]) for g in f4)


def test_column_assignment_matches_specification(f1_pages, f4):
    boxes = np.concatenate([np.asarray(p["boxes"], np.float64) for p in f1_pages])
    off = np.cumsum([0] + [len(p["boxes"]) for p in f1_pages])
    max_cols = 16
    centers = np.zeros((len(f1_pages), max_cols), np.int32)
    n_cols = np.zeros(len(f1_pages), np.int32)
    for i, g in enumerate(f4):
        n_cols[i] = len(g["column_centers"])
        centers[i, : n_cols[i]] = np.asarray(g["column_centers"], np.int32)
    n_cols[3] = 0  # a page without columns -> -1 everywhere
    # select every second box of each page, the rest must stay -1
    sel = np.concatenate([np.arange(off[i], off[i + 1], 2) for i in range(len(f1_pages))]).astype(np.int32)
    n_sel = np.asarray([len(range(off[i], off[i + 1], 2)) for i in range(len(f1_pages))], np.int32)
    sel_buf = np.zeros(len(boxes), np.int32)
    for i in range(len(f1_pages)):
        chunk = np.arange(off[i], off[i + 1], 2)
        sel_buf[off[i]: off[i] + len(chunk)] = chunk
    got = ops.assign_columns(boxes, off, centers, n_cols, sel_idx=sel_buf, n_sel=n_sel).cpu().numpy()
    expect = np.full(len(boxes), -1, np.int64)
    for i, p in enumerate(f1_pages):
        idx = np.arange(off[i], off[i + 1], 2)
        expect[idx] = ob.assign_columns(boxes[idx].tolist(), centers[i, : n_cols[i]].tolist())
    assert np.array_equal(got, expect)
    assert (got >= 0).sum() > 1500 and (got[off[3]: off[4]] == -1).all()


def test_columns_guards():
    assert api.find_column_centers([], [], [], 1000, 1000, 100.0) == ([], [])
    assert api.find_column_centers([[0, 0, 100, 10]], ["figure"], [0.9], 1000, 1000, 100.0) == ([], [])
    assert api.find_column_centers([[0, 0, 100, 10]], ["plain_text"], [0.9], 1000, 1000, 0) == ([], [])


# ============================================================================== chained on device
def test_chained_stages_on_device_match_oracle_chain():
    """filter -> merge -> median -> columns chained through sel_idx/n_sel without leaving the GPU."""
    cfgs = [(8000, 6000, 4, 4, 10000, 41), (3801, 5601, 2, 2, 2000, 42), (2778, 4187, 2, 2, 1500, 43)]
    dets = [synth.page_detections(w, h, r, c, 20.0, n, seed) for (w, h, r, c, n, seed) in cfgs]
    boxes = np.concatenate([d["boxes_local"] for d in dets])
    cell_base = np.cumsum([0] + [len(d["cells"]) for d in dets])
    box_cell = np.concatenate([d["box_cell"] + cell_base[i] for i, d in enumerate(dets)]).astype(np.int32)
    cells = np.concatenate([d["cells"] for d in dets])
    scores = np.concatenate([d["scores"] for d in dets])
    classes = np.concatenate([d["classes"] for d in dets])
    wh = [[d["width"], d["height"]] for d in dets]
    off = np.cumsum([0] + [len(d["boxes_local"]) for d in dets])
    bp, _, kept1, n1 = ops.edge_filter(boxes, box_cell, cells, wh, off, 10)
    kept2, n2, ws = ops.nms_merge(bp, scores, classes, off, 0.5, sel_idx=kept1, n_sel=n1, max_boxes_per_page=10000)
    flags = ops.class_flags(classes)
    med, nb = ops.width_median(bp, flags, off, wh, 0.2, sel_idx=kept2, n_sel=n2)
    centers, widths, n_cols = ops.column_peaks(bp, flags, scores, off, wh, med, 0.3, sel_idx=kept2, n_sel=n2)
    torch.cuda.synchronize()
    assert ws.stats()["status"] == 0
    bp_h, kept2, n2 = bp.cpu().numpy(), kept2.cpu().numpy(), n2.cpu().numpy()
    for i, d in enumerate(dets):
        sl = slice(off[i], off[i + 1])
        page_boxes = bp_h[sl]
        keep1 = [j for j in range(len(page_boxes))
                 if not ob.touches_internal_edge(page_boxes[j], d["cells"][d["box_cell"][j]], d["width"], d["height"], 10)]
        keep1 = np.asarray(keep1)
        order = nms_pick_order_c(page_boxes[keep1], d["scores"][keep1], d["classes"][keep1], 0.5)
        final = keep1[order]
        assert n2[i] == len(final)
        assert np.array_equal(kept2[off[i]: off[i] + n2[i]] - off[i], final)
        fb, fs, fc = page_boxes[final], d["scores"][final], d["classes"][final]
        names = synth.class_names_of(fc)
        ref_m, ref_nb = ob.median_width(fb.tolist(), names, d["width"], 0.2)
        assert float(med[i].item()) == float(ref_m) and int(nb[i].item()) == ref_nb
        rc, rw = ob.column_centers(fb.tolist(), names, fs.tolist(), d["width"], d["height"], ref_m, 0.3)
        k = int(n_cols[i].item())
        assert [float(x) for x in centers[i, :k].cpu().numpy()] == [float(x) for x in rc]
        assert [float(x) for x in widths[i, :k].cpu().numpy()] == [float(x) for x in rw]
