"""Pin the oracle (oracle/) against golden vectors produced by the unmodified
reference scripts (oracle/gen_golden.py) and against the reference's own
committed stage-3 outputs.  CPU only."""
import numpy as np
import pytest

from conftest import load_golden
from multimodal_embeddings_b200 import synth
from oracle import boxes as ob
from oracle import tiler as ot
from oracle.nms_fast import nms_pick_order_c


# ---------------------------------------------------------------- stage 1 geometry
def test_grid_geometry_matches_reference():
    g = load_golden("stage1_geometry.json")
    assert len(g["split"]) == 49
    for case in g["split"]:
        cells = ot.grid_cells(case["width"], case["height"], case["rows"], case["cols"], case["overlap"])
        assert len(cells) == len(case["cells"])
        for mine, ref in zip(cells, case["cells"]):
            assert mine["coordinates"] == ref["coordinates"]
            # int vs float mixing must survive too (JSON schema detail)
            for k, v in ref["coordinates"].items():
                assert type(mine["coordinates"][k]) is type(v), (k, v)
            assert (mine["row"], mine["col"]) == (ref["row"], ref["col"])
            x0, y0, x1, y1 = mine["slice"]
            assert [y1 - y0, x1 - x0] == ref["shape"]
        arr = synth.grid_cells_f64(case["width"], case["height"], case["rows"], case["cols"], case["overlap"])
        ref_arr = np.array([[c["coordinates"][k] for k in ("x_start", "y_start", "x_end", "y_end")]
                            for c in case["cells"]], np.float64)
        assert np.array_equal(arr, ref_arr)


def test_translate_and_parse_match_reference():
    g = load_golden("stage1_geometry.json")
    for t in g["translate"]:
        assert ot.translate_boxes(t["boxes"], t["cell_coordinates"]) == t["out"]
    for p in g["parse_grid_configs"]:
        assert [list(x) for x in ot.parse_grid_configs(p["in"])] == p["out"]


# ---------------------------------------------------------------- tiler pixels
@pytest.mark.parametrize("shape,dst", [((211, 280), (102, 77)), ((180, 240), (128, 96)), ((97, 131), (160, 119)),
                                       ((64, 64), (32, 32)), ((50, 70), (70, 50))])
def test_fixed_point_resize_equals_cv2(shape, dst):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(shape[0] * 1000 + dst[0])
    img = rng.integers(0, 256, (*shape, 3), dtype=np.uint8)
    ref = cv2.resize(img, dst, interpolation=cv2.INTER_LINEAR)
    assert np.array_equal(ot.resize_fixed_point(img, *dst), ref)


def test_letterbox_model_equals_cv2_primitives():
    pytest.importorskip("cv2")
    page = synth.page_pixels(1203, 907, 3)
    for rows, cols in [(1, 1), (2, 2), (3, 2)]:
        for cell in ot.split_array_into_grid(page, rows, cols, 20.0):
            a = ot.letterbox_tile_cv2(cell["image"], imgsz=256, stride=32, auto=True)
            b = ot.letterbox_tile_model(cell["image"], imgsz=256, stride=32, auto=True)
            assert a.shape == b.shape and a.shape[0] == 3
            assert a.shape[1] % 32 == 0 and a.shape[2] % 32 == 0
            assert np.array_equal(a, b)
            sq = ot.letterbox_tile_cv2(cell["image"], imgsz=256, stride=32, auto=False)
            assert sq.shape == (3, 256, 256)
            assert np.array_equal(sq, ot.letterbox_tile_model(cell["image"], imgsz=256, stride=32, auto=False))


def test_letterbox_geometry_cfg3_shapes():
    # cfg3 tile shapes -> 1024x768 / 1024x672 / 1024x896
    assert ot.letterbox_geometry(2800, 2100)["out_w"] == 1024 and ot.letterbox_geometry(2800, 2100)["out_h"] == 768
    g = ot.letterbox_geometry(2800, 1800)
    assert (g["out_w"], g["out_h"], g["new_h"], g["pad_t"]) == (1024, 672, 658, 7)
    assert ot.letterbox_geometry(2400, 2100)["out_h"] == 896
    assert ot.letterbox_geometry(2400, 1800)["out_h"] == 768


def test_u8_to_f16_table():
    v = np.arange(256, dtype=np.uint8)
    t = ot.u8_to_f16_unit(v)
    assert t.dtype == np.float16 and t[0] == 0 and t[255] == 1
    assert t[114] == np.float16(np.float32(114) / np.float32(255))
    assert np.all(np.diff(t.astype(np.float32)) > 0)


# ---------------------------------------------------------------- stage 2
def test_edge_filter_matches_reference():
    cases = load_golden("stage2_filter.json.gz")
    assert len(cases) == 6
    n_drop = 0
    for case in cases:
        cells = [{"cell_coordinates": cc, "boxes_original": bo}
                 for cc, bo in zip(case["cell_coordinates"], case["boxes_original"])]
        kept = ob.filter_cells(cells, case["width"], case["height"], case["threshold"])
        assert kept == case["kept"]
        n_drop += sum(len(b) for b in case["boxes_original"]) - sum(len(k) for k in kept)
    assert n_drop > 1000


# ---------------------------------------------------------------- stage 3
def test_iou_matches_reference():
    g = load_golden("stage3_nms.npz")
    mine = np.array([ob.iou(p[:4].tolist(), p[4:].tolist()) for p in g["iou_pairs"]])
    assert np.array_equal(mine, g["iou_values"])


def test_nms_matches_reference_python_and_c():
    g = load_golden("stage3_nms.npz")
    for name in g["cases"]:
        b, s, c = g[f"{name}_boxes"], g[f"{name}_scores"], g[f"{name}_classes"]
        thr = float(g[f"{name}_thr"][0])
        kept = g[f"{name}_kept"]
        assert np.array_equal(nms_pick_order_c(b, s, c, thr), kept), name
        if len(s) <= 1300:
            assert ob.nms_pick_order(b.tolist(), s.tolist(), c.tolist(), thr) == kept.tolist(), name
    assert ob.nms_pick_order([], [], [], 0.5) == []
    assert len(nms_pick_order_c(np.zeros((0, 4)), np.zeros(0), np.zeros(0))) == 0


def test_nms_idempotent_on_reference_outputs(f1_pages):
    """Known-answer test held by the reference's own data: stage 3 applied to its
    committed outputs is the identity, and no same-class pair exceeds IoU 0.5."""
    assert len(f1_pages) == 19 and sum(len(p["boxes"]) for p in f1_pages) == 4050
    for p in f1_pages:
        n = len(p["boxes"])
        assert ob.nms_pick_order(p["boxes"], p["scores"], p["classes"], 0.5) == list(range(n))
        assert nms_pick_order_c(p["boxes"], p["scores"], p["classes"], 0.5).tolist() == list(range(n))


def test_tile_nms_restatement_matches_torchvision_goldens():
    """Per-tile class-agnostic NMS (1_doclayout_bboxes.py:217-225): the float32 restatement against
    outputs of torchvision.ops.nms itself (goldens), and live when torchvision is importable."""
    g = load_golden("stage1_tile_nms.npz")
    assert len(g["cases"]) == 5
    for name in g["cases"]:
        b, s, thr = g[f"{name}_boxes"], g[f"{name}_scores"], float(g[f"{name}_thr"][0])
        assert ob.nms_torchvision_f32(b, s, thr) == g[f"{name}_keep"].tolist(), name
        assert 0 < len(g[f"{name}_keep"]) < len(s)
    try:
        import torch
        import torchvision
    except Exception:
        return
    rng = np.random.default_rng(3)
    for _ in range(5):
        n = int(rng.integers(1, 300))
        xy = rng.uniform(0, 500, (n, 2)).astype(np.float32)
        wh = rng.uniform(1, 200, (n, 2)).astype(np.float32)
        b = np.concatenate([xy, xy + wh], 1)
        s = rng.uniform(0, 1, n).astype(np.float32)
        ref = torchvision.ops.nms(torch.tensor(b), torch.tensor(s), 0.45).numpy().tolist()
        assert ob.nms_torchvision_f32(b, s, 0.45) == ref


# ---------------------------------------------------------------- stages 4-5
def test_stage45_on_reference_outputs(f1_pages, f4):
    for p, g in zip(f1_pages, f4):
        assert p["name"] == g["name"]
        w, h = p["image_size"]["width"], p["image_size"]["height"]
        med, nb = ob.median_width(p["boxes"], p["class_names"], w, 0.2)
        assert float(med) == g["median_width"] and nb == g["n_bins"]
        for use_scipy in (True, False):
            c, cw = ob.column_centers(p["boxes"], p["class_names"], p["scores"], w, h, med, 0.3, use_scipy=use_scipy)
            assert [float(x) for x in c] == g["column_centers"], (p["name"], use_scipy)
            assert [float(x) for x in cw] == g["column_widths"], (p["name"], use_scipy)


def test_stage45_on_synthetic_kept_sets():
    import hashlib
    for g in load_golden("stage45_synth.json"):
        det = synth.page_detections(g["width"], g["height"], g["rows"], g["cols"], 20.0, g["n"], g["seed"])
        b = det["boxes_local"] + det["cells"][det["box_cell"]][:, [0, 1, 0, 1]]
        k = nms_pick_order_c(b, det["scores"], det["classes"], 0.5)
        b, s, c = b[k], det["scores"][k], det["classes"][k]
        assert hashlib.sha256(b.tobytes() + s.tobytes() + c.tobytes()).hexdigest() == g["input_sha256"], \
            "synthetic generator drifted: regenerate goldens (python -m oracle.gen_golden)"
        names = synth.class_names_of(c)
        med, nb = ob.median_width(b.tolist(), names, g["width"], g["min_margin_percent"])
        assert float(med) == g["median_width"] and nb == g["n_bins"]
        if med > 0:
            cc, cw = ob.column_centers(b.tolist(), names, s.tolist(), g["width"], g["height"], med,
                                       g["min_confidence"], use_scipy=False)
        else:
            cc, cw = [], []
        assert [float(x) for x in cc] == g["column_centers"]
        assert [float(x) for x in cw] == g["column_widths"]


def test_find_peaks_restatement_equals_scipy():
    from scipy.signal import find_peaks
    rng = np.random.default_rng(5)
    for t in range(200):
        n = int(rng.integers(5, 400))
        x = rng.random(n)
        if t % 3 == 0:
            x = np.round(x, 1)  # plateaus and ties
        if t % 4 == 0:
            x = np.convolve(x, np.ones(7) / 7, mode="same")
        h, d, pr = float(rng.uniform(0, 0.6)), float(rng.uniform(1, 20)), float(rng.uniform(0, 0.3))
        ref, _ = find_peaks(x, height=h, distance=d, prominence=pr)
        lm = ob.local_maxima(x)
        assert lm == find_peaks(x)[0].tolist()
        if len({x[p] for p in lm}) != len(lm):
            # equal-height peaks: scipy's distance step uses an unstable argsort, so the
            # winner is implementation-defined; only the tie-free steps are comparable.
            d = 1.0
        ref, _ = find_peaks(x, height=h, distance=d, prominence=pr)
        assert ob.find_peaks_restated(x, h, d, pr) == ref.tolist()
