"""CPU checks of libpagegeom.so: it loads, exports every symbol include/pagegeom.h declares,
its host-only tile planner reproduces the reference geometry, and the inline arithmetic the
kernels are compiled from (csrc/pg_math.h, evaluated on the host through the pg_hostcheck_*
hooks) agrees bit-for-bit with the oracle.  No kernel is launched here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden
from multimodal_embeddings_b200 import _lib, build, ops, synth
from oracle import boxes as ob
from oracle import tiler as ot


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.lib()


def test_header_symbols_exported(lib):
    text = open(os.path.join(ROOT, "include", "pagegeom.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = set(re.findall(r"\b(pg_[a-z0-9_]+)\s*\(", text))
    assert len(declared) >= 20
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.pg_version() >= 100


def test_invalid_arguments_are_reported(lib):
    handle = C.c_void_p()
    rows = (C.c_int32 * 1)(2)
    cols = (C.c_int32 * 1)(0)
    rc = lib.pg_tile_plan_create(100, 100, rows, cols, 1, 20.0, 1024, 32, 1, 1, C.byref(handle))
    assert rc == 1 and b"grid" in lib.pg_last_error()
    assert lib.pg_edge_filter(None, 0, None, None, None, None, 3, 10.0, None, None, None, None, None) == 1
    assert lib.pg_nms_merge(None, None, None, None, None, None, 0, 0, 0, 0.5, None, None, None, 0, None) == 0


def test_plan_geometry_matches_reference_goldens(lib):
    g = load_golden("stage1_geometry.json")
    for case in g["split"]:
        plan = ops.TilePlan(case["width"], case["height"], [(case["rows"], case["cols"])], case["overlap"])
        assert len(plan.tiles) == len(case["cells"])
        for t, ref in enumerate(case["cells"]):
            cc = plan.cell_coordinates(t)
            assert cc == ref["coordinates"]
            for k, v in ref["coordinates"].items():
                assert type(cc[k]) is type(v)
            info = plan.tiles[t]
            assert (info["row"], info["col"]) == (ref["row"], ref["col"])
            assert [info["y1"] - info["y0"], info["x1"] - info["x0"]] == ref["shape"]


def test_plan_letterbox_and_offsets_match_oracle(lib):
    for (w, h) in [(8000, 6000), (7934, 5755), (3801, 5601), (2778, 4187), (640, 480)]:
        for auto in (True, False):
            plan = ops.TilePlan(w, h, [(1, 1), (2, 2), (3, 3), (4, 4)], 20.0, 1024, 32, auto)
            assert len(plan.tiles) == 30
            off = 0
            for info in plan.tiles:
                g = ot.letterbox_geometry(info["x1"] - info["x0"], info["y1"] - info["y0"], 1024, 32, auto)
                for k in ("new_w", "new_h", "pad_l", "pad_t", "out_w", "out_h"):
                    assert info[k] == g[k], (w, h, auto, k)
                assert info["out_offset"] == off
                off += 3 * info["out_w"] * info["out_h"]
            assert plan.out_elems == off
            assert plan.algorithmic_bytes == 3 * w * h + 2 * off
    # BASELINE.md figures: cfg3 = 8000x6000, 4x4, stride-32 letterbox -> 220.3 MB per page
    plan = ops.TilePlan(8000, 6000, [(4, 4)], 20.0)
    assert abs(plan.algorithmic_bytes / 1e6 - 220.3) < 0.05
    sq = ops.TilePlan(8000, 6000, [(4, 4)], 20.0, auto=False)
    assert abs(sq.algorithmic_bytes / 1e6 - 244.7) < 0.05


def test_hostcheck_iou_bit_exact(lib):
    g = load_golden("stage3_nms.npz")
    for p, ref in zip(g["iou_pairs"], g["iou_values"]):
        a, b = np.ascontiguousarray(p[:4]), np.ascontiguousarray(p[4:])
        assert lib.pg_hostcheck_iou(_lib.ptr(a), _lib.ptr(b)) == ref


def test_hostcheck_iou_gt_equals_divide_then_compare(lib):
    """The merge kernel decides `iou > thr` without a divide except in a 2^-50 band around thr;
    it must agree with the reference expression everywhere, in particular AT the threshold."""
    rng = np.random.default_rng(11)
    cases = []
    g = load_golden("stage3_nms.npz")
    for p in g["iou_pairs"]:
        for thr in (0.5, 0.3, 0.45, 0.0, -1.0, 0.7):
            cases.append((p[:4].copy(), p[4:].copy(), thr))
    # pairs whose IoU is exactly representable and equal to the threshold, then nudged by ulps
    for thr, a, b in [(0.5, [0, 0, 2, 1], [0, 0, 1, 1]), (0.25, [0, 0, 4, 1], [0, 0, 1, 1]),
                      (0.5, [10, 10, 30, 20], [10, 10, 20, 20]), (0.75, [0, 0, 4, 1], [0, 0, 3, 1]),
                      (0.5, [100.5, 7.25, 300.5, 57.25], [100.5, 7.25, 200.5, 57.25])]:
        for _ in range(400):
            aa, bb = np.asarray(a, np.float64), np.asarray(b, np.float64)
            for arr in (aa, bb):
                for k in range(4):
                    steps = int(rng.integers(-3, 4))
                    for _s in range(abs(steps)):
                        arr[k] = np.nextafter(arr[k], np.inf if steps > 0 else -np.inf)
            cases.append((aa, bb, thr))
            cases.append((aa * 3.0, bb * 3.0, thr))
    n_true = n_band = 0
    for a, b, thr in cases:
        a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
        v = ob.iou(a.tolist(), b.tolist())
        ref = v > thr
        n_true += ref
        n_band += (thr > 0 and abs(v - thr) <= 4e-16 * thr)
        assert bool(lib.pg_hostcheck_iou_gt(_lib.ptr(a), _lib.ptr(b), float(thr))) == ref, (a, b, thr, v)
    assert n_true > 100 and n_band > 50  # the band (true-divide path) is really exercised


def test_hostcheck_iou_gt_f32_matches_float32_restatement(lib):
    rng = np.random.default_rng(5)
    n_true = 0
    for _ in range(3000):
        xy = rng.uniform(0, 300, 4).astype(np.float32)
        wh = rng.uniform(0, 150, 4).astype(np.float32)
        a = np.array([xy[0], xy[1], xy[0] + wh[0], xy[1] + wh[1]], np.float32)
        b = np.array([xy[2], xy[3], xy[2] + wh[2], xy[3] + wh[3]], np.float32)
        if rng.random() < 0.3:
            b = (a + rng.normal(0, 3, 4)).astype(np.float32)
        thr = float(rng.choice([0.45, 0.3, 0.7, 0.0]))
        area = lambda q: (q[2] - q[0]) * (q[3] - q[1])  # noqa: E731  float32 arithmetic
        w = max(np.float32(0), min(a[2], b[2]) - max(a[0], b[0]))
        h = max(np.float32(0), min(a[3], b[3]) - max(a[1], b[1]))
        inter = np.float32(w * h)
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / np.float32(np.float32(area(a) + area(b)) - inter)
        ref = bool(np.float64(ovr) > thr)
        n_true += ref
        assert bool(lib.pg_hostcheck_iou_gt_f32(_lib.ptr(a), _lib.ptr(b), thr)) == ref
    assert n_true > 200


def test_hostcheck_edge_touch_matches_reference(lib):
    n = 0
    for case in load_golden("stage2_filter.json.gz"):
        w, h, thr = case["width"], case["height"], case["threshold"]
        for cc, boxes, kept in zip(case["cell_coordinates"], case["boxes_original"], case["kept"]):
            cell = np.asarray(ob.cell_tuple(cc, w, h), np.float64)
            ks = set(kept)
            for i, b in enumerate(boxes[:150]):
                bb = np.asarray(b, np.float64)
                assert bool(lib.pg_hostcheck_edge_touch(_lib.ptr(bb), _lib.ptr(cell), w, h, float(thr))) == (i not in ks)
                n += 1
    assert n > 2000


def test_hostcheck_density_weight(lib):
    rng = np.random.default_rng(0)
    for _ in range(2000):
        left = int(rng.integers(0, 900))
        right = left + int(rng.integers(0, 300))
        center = int(rng.integers(left - 5, right + 6))
        b = int(rng.integers(left, right + 1))
        ref = 1.0 - 0.5 * min(1.0, abs(b - center) / ((right - left) / 2 + 1e-6))
        assert lib.pg_hostcheck_density_weight(b, left, right, center) == ref


def test_hostcheck_density_reciprocal_form_is_exact_on_its_whole_domain(lib):
    # the density kernel replaces n/half by y=1/half, q0=n*y, r=fma(-q0,half,n), fma(r,y,q0);
    # exhaustive over every (right-left, |bin-center|) the kernel can meet (PG_RCP_DOMAIN = 2100)
    assert lib.pg_hostcheck_density_rcp_mismatches(2100, 2100) == 0


@pytest.mark.parametrize("src,dst", [((97, 131), (64, 48)), ((211, 280), (102, 77)), ((40, 50), (100, 80))])
def test_hostcheck_resize_rows_equal_cv2(lib, src, dst):
    cv2 = pytest.importorskip("cv2")
    sh, sw = src
    dw, dh = dst
    img = np.random.default_rng(sh + dw).integers(0, 256, (sh, sw, 3), dtype=np.uint8)
    ref = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)
    out = np.zeros((dw, 3), np.uint8)
    _, _, _, _ = ot.resize_coeffs(sh, dh, False)
    y0, y1, _, _ = ot.resize_coeffs(sh, dh, False)
    for dy in range(dh):
        r0, r1 = np.ascontiguousarray(img[y0[dy]]), np.ascontiguousarray(img[y1[dy]])
        packed = lib.pg_hostcheck_resize_row(_lib.ptr(r0), _lib.ptr(r1), sw, sh, dw, dh, dy, _lib.ptr(out))
        assert (packed & 0xFFFF, packed >> 16) == (int(y0[dy]), int(y1[dy]))
        assert np.array_equal(out, ref[dy])


def test_half2_unit_scale_exhaustive():
    """The tiler's epilogue (pg_tiler.cu vpass_pair_half2) reads the byte v as the fp16 subnormal
    v*2^-24 and computes fma(x, 64608, rn16(x*1185)).  Every product and sum below is exact in
    fp64, so one np.float16 conversion models each fp16 rounding; the result must be
    fp16(fp32(v)/255) — what LetterBox + `.half() / 255` produce — for all 256 byte values."""
    c_hi = float(np.uint16(0x7BE3).view(np.float16))
    c_lo = float(np.uint16(0x64A1).view(np.float16))
    assert (c_hi, c_lo) == (64608.0, 1185.0)
    v = np.arange(256)
    want = (v.astype(np.float32) / np.float32(255.0)).astype(np.float16)
    x = v.astype(np.uint16).view(np.float16).astype(np.float64)
    e = (x * c_lo).astype(np.float16).astype(np.float64)
    got = (x * c_hi + e).astype(np.float16)
    assert np.array_equal(got.view(np.uint16), want.view(np.uint16))
    # and the integer half: ((a>>16)+(b>>16)+2)>>2 == byte 1 of 64*((a+(2<<16))>>16 + (b>>16))
    rng = np.random.default_rng(7)
    b0 = rng.integers(0, 2049, 200000)
    h0, h1 = rng.integers(0, 255 * 2048 + 1, (2, 200000))
    a, b = b0 * (h0 >> 4), (2048 - b0) * (h1 >> 4)
    ref = ((a >> 16) + (b >> 16) + 2) >> 2
    new = ((((a + 0x20000) >> 16) + (b >> 16)) * 64 >> 8) & 0xFF
    assert np.array_equal(ref, new) and (((a + 0x20000) >> 16) + (b >> 16)).max() < 1024 and (a + 0x20000).max() < 2 ** 32


def _format_doubles(lib, x):
    x = np.ascontiguousarray(x, np.float64)
    buf, ln = np.zeros(len(x) * 24, np.uint8), np.zeros(len(x), np.int32)
    total = lib.pg_hostcheck_format_doubles(_lib.ptr(x), len(x), _lib.ptr(buf), _lib.ptr(ln))
    assert total == int(ln.sum()) and ln.max() <= 24
    raw = buf.tobytes()
    return [raw[24 * i: 24 * i + ln[i]].decode("ascii") for i in range(len(x))]


def test_format_double_equals_python_repr(lib):
    """csrc/pg_fmt.h (Ryu shortest digits + CPython's repr layout) against float.__repr__ / json.dumps —
    the number formatting of every record the reference writes (json.dump, e.g. 3_combine_grids.py:441-443)."""
    import json
    special = [0.0, -0.0, 1.0, -1.0, 0.1, 0.5, 1e16, 1e15, 9999999999999998.0, 123456789012345680.0, 1e-4, 1e-5,
               9.999e-5, 1.5e-5, 5e-324, 2.2250738585072014e-308, 2.225073858507201e-308, 1.7976931348623157e308,
               float("inf"), float("-inf"), float("nan"), 1e22, 1e23, 9007199254740993.0, 0.3, 2 / 3, 100.0, 1e21,
               4.35, 0.285, 1.005, 2.675, 1234567890123456.7, 12345678901234567.0, 0.001, 0.00012345]
    assert _format_doubles(lib, special) == [json.dumps(v) for v in special]
    rng = np.random.default_rng(1)
    sets = {
        "bit patterns": rng.integers(0, 2 ** 63, 400_000, dtype=np.int64).view(np.float64),
        "negative bit patterns": -rng.integers(0, 2 ** 63, 100_000, dtype=np.int64).view(np.float64),
        "float32 pixel coordinates": rng.uniform(0, 8000, 300_000).astype(np.float32).astype(np.float64),
        "float32 scores": rng.uniform(0, 1, 200_000).astype(np.float32).astype(np.float64),
        "two-decimal coordinates": np.round(rng.uniform(0, 8000, 200_000), 2),
        "integers": rng.integers(-10 ** 17, 10 ** 17, 100_000).astype(np.float64),
        "powers of ten": np.array([float(f"{m}e{e}") for m in (1, 2, 5, 9, 1.5, 9.5) for e in range(-323, 308)]),
        "powers of two": np.array([2.0 ** e for e in range(-1074, 1024)]),
        "neighbours of powers of two": np.nextafter(np.array([2.0 ** e for e in range(-1000, 1000)]), 0),
    }
    for name, x in sets.items():
        x = x[np.isfinite(x)]
        got = _format_doubles(lib, x)
        bad = [(v, s) for v, s in zip(x.tolist(), got) if s != repr(v)]
        assert not bad, (name, len(bad), bad[:3])


def test_combined_head_tail_brackets_a_json_dump_document():
    import json
    for size, srcs in (({"width": 10, "height": 20}, ["a.json", "b_grid_2x2.json"]), (None, [])):
        head, tail = ops.combined_head_tail('/x/pa"ge é.png', size, 0.5, srcs)
        doc = {"image_path": '/x/pa"ge é.png', "image_size": size, "parameters": {"iou_threshold": 0.5},
               "boxes": [], "classes": [], "scores": [], "class_names": [], "source_jsons": srcs}
        empty = head + b'],\n  "classes": [],\n  "scores": [],\n  "class_names": [],' + tail
        assert empty.decode("ascii") == json.dumps(doc, indent=2)


def test_no_cpu_fallback_without_cuda(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(_lib.PageGeomError):
        ops.edge_filter(np.zeros((1, 4)), [0], np.zeros((1, 4)), [[10, 10]], [0, 1])
    det = synth.page_detections(640, 480, 2, 2, 20.0, 50, 1)
    assert det["boxes_local"].shape == (50, 4)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "multimodal_embeddings_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_parse_json_number_equals_python_float(lib):
    """csrc/pg_fmt.h pg_parse_json_number (Ryu's inverse, <= 17 significant digits) against float(): what
    json.load does for the reference's stage-4/5 readers (4_extract_median_widths.py:103-151)."""
    def parse(strs):
        text = ("\n".join(strs) + "\n").encode()
        lens = np.fromiter((len(s) for s in strs), np.int64, len(strs))
        offs = np.concatenate([[0], np.cumsum(lens + 1)[:-1]]).astype(np.int64)
        buf = np.frombuffer(text, np.uint8).copy()
        out, used = np.zeros(len(strs)), np.zeros(len(strs), np.int32)
        lib.pg_hostcheck_parse_numbers(_lib.ptr(buf), len(buf), _lib.ptr(offs), len(strs), _lib.ptr(out), _lib.ptr(used))
        return out, used, lens
    special = ["0", "-0", "0.0", "-0.0", "1", "0.1", "1e16", "1E-5", "1.5e-05", "5e-324", "2.2250738585072014e-308",
               "1.7976931348623157e+308", "1e309", "-1e309", "1e-400", "2.4703282292062327e-324", "2.4703282292062328e-324",
               "9007199254740993", "0.30000000000000004", "1e22", "1e23", "8.5e-324", "Infinity", "-Infinity",
               "12345678901234567000", "0.00000000000000000000000000000000000001e38", "100"]
    out, used, lens = parse(special)
    ref = np.array([float(s) for s in special])
    assert np.array_equal(out.view(np.uint64), ref.view(np.uint64)) and np.array_equal(used, lens)
    out, used, _ = parse(["NaN", "123456789012345678", "1.2.3", "-", "e5"])
    assert np.isnan(out[0]) and used.tolist() == [3, 0, 0, 0, 0]
    rng = np.random.default_rng(2)
    sets = {"bit patterns": rng.integers(0, 2 ** 63, 300_000, dtype=np.int64).view(np.float64),
            "float32 coordinates": rng.uniform(0, 8000, 200_000).astype(np.float32).astype(np.float64),
            "subnormals": rng.integers(1, 2 ** 52, 100_000, dtype=np.int64).view(np.float64),
            "powers of two": np.array([2.0 ** e for e in range(-1074, 1024)])}
    for name, x in sets.items():  # repr round trip: parse(repr(x)) == x
        x = x[np.isfinite(x)]
        strs = [repr(v) for v in x.tolist()]
        out, used, lens = parse(strs)
        assert np.array_equal(out.view(np.uint64), x.view(np.uint64)) and np.array_equal(used, lens), name
    strs = [f"{int(rng.integers(10 ** (nd - 1), 10 ** nd))}e{int(rng.integers(-345, 310))}"
            for nd in rng.integers(1, 18, 150_000)]  # arbitrary (not shortest) decimals, whole exponent range
    out, used, lens = parse(strs)
    ref = np.array([float(s) for s in strs])
    assert np.array_equal(out.view(np.uint64), ref.view(np.uint64))
    mids = [f"{2 ** 53 + 1 + 2 * k}e-{j}" for k in range(300) for j in (0, 1, 2, 5)]  # ties and near-ties
    out, _, _ = parse(mids)
    assert np.array_equal(out.view(np.uint64), np.array([float(s) for s in mids]).view(np.uint64))
