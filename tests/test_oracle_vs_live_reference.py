"""Differential fuzz of the oracle against the UNMODIFIED reference scripts, imported from /root/reference.

Runs only where the reference is mounted (the build container); on the GPU box — where /root/reference does not
exist — every test here is skipped and the committed goldens (tests/golden/, made by oracle/gen_golden.py from the
same scripts) stand in.  Seeds are fixed; sizes keep the whole file at a few seconds."""
import importlib.util
import logging
import os

import numpy as np
import pytest

from multimodal_embeddings_b200 import synth
from oracle import boxes as ob
from oracle import tiler as ot
from oracle.nms_fast import nms_pick_order_c

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference scripts are not mounted here")


def _ref(stem):
    spec = importlib.util.spec_from_file_location("live_ref_" + stem.split("_")[0], os.path.join(REF, stem + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(autouse=True)
def _quiet():
    logging.disable(logging.CRITICAL)  # the scripts log every call
    yield
    logging.disable(logging.NOTSET)


@pytest.fixture(scope="module")
def r1():
    return _ref("1_doclayout_bboxes")


@pytest.fixture(scope="module")
def r2():
    return _ref("2_edge_box_filter")


@pytest.fixture(scope="module")
def r3():
    return _ref("3_combine_grids")


@pytest.fixture(scope="module")
def r4():
    return _ref("4_extract_median_widths")


@pytest.fixture(scope="module")
def r5():
    return _ref("5_detect_column_centers")


def test_translate_and_grid_parsing(r1):
    rng = np.random.default_rng(1)
    for _ in range(50):
        boxes = rng.uniform(0, 3000, (int(rng.integers(0, 20)), 4)).astype(np.float32).astype(np.float64).tolist()
        cc = {"x_start": float(rng.uniform(0, 5000)), "y_start": int(rng.integers(0, 5000)),
              "x_end": 6000.5, "y_end": 7000}
        assert ot.translate_boxes(boxes, cc) == r1.translate_coordinates_to_original(boxes, cc)
    for s in ["2x2,3x3,4x4", "2x3", " 1x1 , 5x2", "3x", "axb", "", "2x2,,3x3", "0x2", "2X2"]:
        assert ot.parse_grid_configs(s) == r1.parse_grid_configs(s), s


def test_edge_predicate_on_threshold_lattice(r2):
    """Boxes whose sides sit exactly on, one ulp inside and one ulp outside `edge +- threshold` of every cell side."""
    rng = np.random.default_rng(2)
    for _ in range(40):
        w, h = int(rng.integers(300, 9000)), int(rng.integers(300, 9000))
        rows, cols = int(rng.integers(1, 5)), int(rng.integers(1, 5))
        thr = int(rng.choice([0, 1, 10, 25]))
        for gc in ot.grid_cells(w, h, rows, cols, float(rng.choice([0.0, 10.0, 20.0, 33.3]))):
            cell = gc["coordinates"]
            cx0, cy0, cx1, cy1 = cell["x_start"], cell["y_start"], cell["x_end"], cell["y_end"]
            xs = [cx0 + thr, cx1 - thr, cx0, cx1, float(rng.uniform(cx0, cx1))]
            ys = [cy0 + thr, cy1 - thr, cy0, cy1, float(rng.uniform(cy0, cy1))]
            for x in xs:
                for y in ys:
                    for d in (0.0, 1.0, -1.0):
                        xa = float(np.nextafter(x, x + d)) if d else float(x)
                        ya = float(np.nextafter(y, y + d)) if d else float(y)
                        for box in ([xa - 50.0, ya - 40.0, xa, ya], [xa, ya, xa + 50.0, ya + 40.0]):
                            want = r2.is_box_touching_internal_edge(box, cell, w, h, thr)
                            assert ob.touches_internal_edge(box, ob.cell_tuple(cell, w, h), w, h, thr) == want
                            lst = [cx0, cy0, cx1, cy1]
                            assert ob.touches_internal_edge(box, ob.cell_tuple(lst, w, h), w, h, thr) == \
                                r2.is_box_touching_internal_edge(box, lst, w, h, thr)


def test_iou_bitwise(r3):
    rng = np.random.default_rng(3)
    a = rng.uniform(0, 100, (4000, 4))
    b = a + rng.normal(0, 3, a.shape)
    b[::7] = a[::7]                       # identical boxes
    b[1::7, 0] = a[1::7, 2]               # touching sides
    a[2::7, 2] = a[2::7, 0]               # zero-width boxes
    for x, y in zip(a.tolist(), b.tolist()):
        got, want = ob.iou(x, y), r3.calculate_iou(x, y)
        assert got == want and np.float64(got).tobytes() == np.float64(want).tobytes()


@pytest.mark.parametrize("seed", range(6))
def test_nms_pick_order_with_ties_and_duplicates(r3, seed):
    rng = np.random.default_rng(100 + seed)
    d = synth.page_detections(2400, 1800, 2, 2, 20.0, int(rng.integers(1, 500)), 900 + seed, dups=int(rng.integers(1, 5)))
    boxes = (d["boxes_local"] + d["cells"][d["box_cell"]][:, [0, 1, 0, 1]])
    scores, classes = d["scores"].copy(), d["classes"].copy()
    n = len(boxes)
    if n > 10:
        scores[rng.integers(0, n, n // 3)] = scores[0]          # many exact ties
        boxes[rng.integers(0, n, n // 5)] = boxes[1]            # exact duplicates
    thr = float(rng.choice([0.5, 0.3, 0.0, 0.9, -1.0]))
    names = synth.class_names_of(classes)
    fb, fs, fc, fn = r3.apply_non_max_suppression(boxes.tolist(), scores.tolist(), classes.tolist(), names, thr)
    picks = ob.nms_pick_order(boxes.tolist(), scores.tolist(), classes.tolist(), thr)
    assert [boxes[i].tolist() for i in picks] == fb and [scores[i] for i in picks] == fs
    assert [classes[i] for i in picks] == fc and [names[i] for i in picks] == fn
    assert nms_pick_order_c(boxes, scores, classes, thr).tolist() == picks


@pytest.mark.parametrize("seed", range(8))
def test_width_median_and_columns(r4, r5, seed):
    rng = np.random.default_rng(200 + seed)
    w, h = int(rng.integers(600, 9000)), int(rng.integers(600, 9000))
    d = synth.page_detections(w, h, 1, 1, 0.0, int(rng.integers(0, 700)), 300 + seed, dups=1)
    boxes = d["boxes_local"].tolist()
    names = synth.class_names_of(d["classes"])
    scores = d["scores"].tolist()
    pct = float(rng.choice([0.2, 0.05, 1.0, 0.0]))
    widths = [b[2] - b[0] for b, nm in zip(boxes, names) if nm == "plain_text"]
    assert ob.plain_text_widths(boxes, names) == widths
    bins_ref = r4.bin_widths(widths, pct, w)
    bins = ob.bin_widths(widths, pct, w)
    assert list(bins.items()) == list(bins_ref.items())
    med_ref, med = r4.calculate_median_width(bins_ref), ob.median_of_bins(bins)
    assert med == med_ref and type(med) is type(med_ref)
    if med_ref > 0:
        for conf in (0.3, 0.6):
            c_ref, w_ref = r5.find_column_centers(boxes, names, scores, w, h, med_ref, conf)
            c, cw = ob.column_centers(boxes, names, scores, w, h, med, conf)
            assert [float(x) for x in c] == [float(x) for x in c_ref]
            assert [float(x) for x in cw] == [float(x) for x in w_ref]


def test_nms_signed_zero_scores(r3):
    """The reference on scores of both zero signs: max()/index compare values, so -0.0 ties with +0.0 and the
    earlier position wins; both oracles (Python and C) must agree with it."""
    from oracle.nms_fast import nms_pick_order_c
    rng = np.random.default_rng(78)
    n = 300
    xy = rng.uniform(0, 600, (n, 2))
    boxes = np.concatenate([xy, xy + rng.uniform(30, 120, (n, 2))], 1)
    classes = rng.integers(0, 2, n).astype(np.float64)
    scores = np.where(rng.random(n) < 0.5, -0.0, 0.0)
    scores[::5] = 0.5
    _, _, _, picked = r3.apply_non_max_suppression(boxes.tolist(), scores.tolist(), classes.tolist(), list(range(n)), 0.3)
    assert picked == ob.nms_pick_order(boxes.tolist(), scores.tolist(), classes.tolist(), 0.3)
    assert picked == list(nms_pick_order_c(boxes, scores, classes, 0.3))
