"""the device-side writer of the stage-3 records must produce, byte for byte, what the
reference's `json.dump(result, f, indent=2)` (3_combine_grids.py:441-443) writes for the same kept boxes."""
import json

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from multimodal_embeddings_b200 import ops  # noqa: E402

pytestmark = pytest.mark.gpu

NAMES = ["title", "plain_text", "abandon", "figure", "figure_caption", "table", 'quo"te\\back', "café ☃", "tab\there"]


def reference_text(image_path, image_size, thr, boxes, classes, scores, names, sources):
    doc = {"image_path": image_path, "image_size": image_size, "parameters": {"iou_threshold": thr},
           "boxes": boxes, "classes": classes, "scores": scores, "class_names": names, "source_jsons": sources}
    return json.dumps(doc, indent=2).encode("ascii")


def make_pages(rng, counts):
    pages = []
    for i, n in enumerate(counts):
        kind = i % 4
        if kind == 0:    # what stage 1 really writes: float32 tensors through .tolist()
            b = rng.uniform(0, 8000, (n, 4)).astype(np.float32).astype(np.float64)
            s = rng.uniform(0.05, 1, n).astype(np.float32).astype(np.float64)
        elif kind == 1:  # arbitrary doubles, both signs, integral values, tiny and huge magnitudes
            b = rng.uniform(-1e4, 1e4, (n, 4))
            b[rng.random((n, 4)) < 0.1] = 0.0
            b[rng.random((n, 4)) < 0.05] = -0.0
            b[::3] = np.round(b[::3])
            b[1::7] *= 1e-7
            b[2::11] *= 1e15
            s = rng.random(n)
        elif kind == 2:  # short decimals
            b = np.round(rng.uniform(0, 5000, (n, 4)), 2)
            s = np.round(rng.random(n), 3)
        else:            # raw bit patterns (finite)
            b = rng.integers(0, 0x7FE0000000000000, (n, 4), dtype=np.int64).view(np.float64)
            b[::2] *= -1.0
            s = rng.integers(0, 0x7FE0000000000000, n, dtype=np.int64).view(np.float64)
        c = rng.integers(0, len(NAMES), n).astype(np.float64)
        pages.append((b, c, s))
    return pages


@pytest.mark.parametrize("with_kept", [False, True])
def test_documents_equal_json_dump(with_kept):
    rng = np.random.default_rng(5 + with_kept)
    counts = [0, 1, 2, 257, 1500, 31, 0, 640]
    pages = make_pages(rng, counts)
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    boxes = np.concatenate([p[0] for p in pages])
    classes = np.concatenate([p[1] for p in pages])
    scores = np.concatenate([p[2] for p in pages])
    name_id = classes.astype(np.int32)
    literals = [json.dumps(n).encode("ascii") for n in NAMES]
    heads, tails, meta = [], [], []
    for i in range(len(counts)):
        path = f"/data/scans/page {i} ü.png"
        size = {"width": 4000 + i, "height": 6000 - i} if i != 3 else None
        srcs = [f"a/{i}.json", f"a/{i}_grid_2x2.json"][: i % 3]
        thr = [0.5, 0.45, 1e-05, 1.0][i % 4]
        h, t = ops.combined_head_tail(path, size, thr, srcs)
        heads.append(h)
        tails.append(t)
        meta.append((path, size, thr, srcs))
    kept = n_kept = None
    sel = [np.arange(c) for c in counts]
    if with_kept:  # an arbitrary ordered subset per page, as the merge's kept_idx would be
        sel = [np.sort(rng.choice(c, size=int(rng.integers(0, c + 1)), replace=False)) if c else np.arange(0) for c in counts]
        sel[4] = rng.permutation(counts[4])[:900]  # pick order need not be ascending
        kept = np.zeros(max(1, off[-1]), np.int32)
        for i, s in enumerate(sel):
            kept[off[i]: off[i] + len(s)] = s + off[i]
        n_kept = np.asarray([len(s) for s in sel], np.int32)
    docs = ops.json_combined(boxes, classes, scores, name_id, off, heads, tails, literals, kept_idx=kept, n_kept=n_kept)
    assert len(docs) == len(counts)
    for i, (path, size, thr, srcs) in enumerate(meta):
        b, c, s = pages[i]
        j = sel[i]
        want = reference_text(path, size, thr, b[j].tolist(), c[j].tolist(), s[j].tolist(),
                              [NAMES[int(x)] for x in c[j]], srcs)
        assert docs[i] == want, (i, docs[i][:200], want[:200])
        assert json.loads(docs[i])["boxes"] == b[j].tolist()  # and it parses back to the same doubles


def test_golden_stage3_tree_is_reproduced():
    """The reference's own stage-3 records (tests/golden/cli_tree.json.gz: what the unmodified
    3_combine_grids.py main() wrote, stored parsed-and-compacted) re-serialised on the device; the expected
    text is json.dumps(record, indent=2), the call at 3:441-443."""
    from conftest import load_golden
    tree = load_golden("cli_tree.json.gz")
    files = {k: v for k, v in tree.items() if k.startswith("3_combined_bboxes/json/") and k.endswith("_combined.json")}
    assert files
    names, docs_in = [], []
    for text in files.values():
        d = json.loads(text)
        docs_in.append((json.dumps(d, indent=2), d))
        for n in d["class_names"]:
            if n not in names:
                names.append(n)
    off = np.concatenate([[0], np.cumsum([len(d["boxes"]) for _, d in docs_in])]).astype(np.int64)
    boxes = np.concatenate([np.asarray(d["boxes"], np.float64).reshape(-1, 4) for _, d in docs_in])
    classes = np.concatenate([np.asarray(d["classes"], np.float64) for _, d in docs_in])
    scores = np.concatenate([np.asarray(d["scores"], np.float64) for _, d in docs_in])
    name_id = np.asarray([names.index(n) for _, d in docs_in for n in d["class_names"]], np.int32)
    ht = [ops.combined_head_tail(d["image_path"], d["image_size"], d["parameters"]["iou_threshold"], d["source_jsons"])
          for _, d in docs_in]
    out = ops.json_combined(boxes, classes, scores, name_id, off, [h for h, _ in ht], [t for _, t in ht],
                            [json.dumps(n).encode("ascii") for n in names])
    for (text, _), got in zip(docs_in, out):
        assert got.decode("ascii") == text


def test_shard_without_any_box():
    """Two pages, no detections at all: four empty arrays per document, exactly as json.dump prints them."""
    meta = [("/a.png", {"width": 10, "height": 20}, 0.5, []), ("/b.png", None, 0.25, ["s.json"])]
    ht = [ops.combined_head_tail(*m) for m in meta]
    docs = ops.json_combined(np.zeros((0, 4)), np.zeros(0), np.zeros(0), np.zeros(0, np.int32), np.zeros(3, np.int64),
                             [h for h, _ in ht], [t for _, t in ht], [b'"plain_text"'])
    for doc, (path, size, thr, srcs) in zip(docs, meta):
        assert doc == reference_text(path, size, thr, [], [], [], [], srcs)


# ------------------------------------------------------------------------------------------------ reader (R1-R3)
def test_device_reader_round_trips_the_writers_documents(tmp_path):
    """records.load_records on files written as json.dump(indent=2) writes them: the number arrays come back
    bit for bit (they were converted on the device), strings and sizes as json.load gives them."""
    from multimodal_embeddings_b200 import records
    rng = np.random.default_rng(11)
    counts = [0, 1, 3, 700, 129, 2050]
    pages = make_pages(rng, counts)
    paths = []
    for i, (b, c, s) in enumerate(pages):
        doc = {"image_path": f'/d/"boxes": [ p{i}.png', "image_size": {"width": 100 + i, "height": 200}, "parameters": {"iou_threshold": 0.5},
               "boxes": b.tolist(), "classes": c.tolist(), "scores": s.tolist(),
               "class_names": [NAMES[int(x)] for x in c], "source_jsons": [f"s{i}.json"]}
        p = tmp_path / f"p{i}_combined.json"
        p.write_text(json.dumps(doc, indent=2))
        paths.append(str(p))
    compact = tmp_path / "compact_combined.json"   # another layout: must still load (through json.loads)
    compact.write_text(json.dumps({"image_path": "c.png", "image_size": None, "parameters": {}, "boxes": [[1, 2, 3, 4]],
                                   "classes": [1.0], "scores": [0.5], "class_names": ["title"], "source_jsons": []}))
    recs = records.load_records(paths + [str(compact)])
    assert list(recs) == paths + [str(compact)]
    for path, (b, c, s) in zip(paths, pages):
        r, want = recs[path], json.load(open(path))
        assert isinstance(r["boxes"], np.ndarray) and list(r) == list(want)
        assert r["boxes"].shape == (len(c), 4)
        assert np.array_equal(r["boxes"].view(np.uint64), b.view(np.uint64).reshape(-1, 4))
        assert np.array_equal(r["classes"], c) and np.array_equal(r["scores"].view(np.uint64), s.view(np.uint64))
        for k in ("image_path", "image_size", "parameters", "class_names", "source_jsons"):
            assert r[k] == want[k]
    assert recs[str(compact)]["boxes"] == [[1, 2, 3, 4]]


def test_device_reader_ranges_counts_and_unconvertible_tokens():
    text = (b"[\n 1.5,\n -2e-3, 3,4 ]  [ ]   [ 0.1234567890123456789, 7.0, NaN, -Infinity, 1e400, 5e-324 ]"
            b"[12345678901234567000, 123456789012345678]")
    a = text.index(b"[ ]")
    b2 = text.index(b"[ 0.12")
    c = text.index(b"[1234")
    ranges = [(0, a), (a, b2), (b2, c), (c, len(text))]
    vals, off, bad = ops.json_parse_numbers(text, ranges)
    v = vals.cpu().numpy()
    assert off.tolist() == [0, 4, 4, 10, 12] and bad.tolist() == [0, 0, 1, 1]
    assert v[:4].tolist() == [1.5, -0.002, 3.0, 4.0]
    assert v[5] == 7.0 and np.isnan(v[6]) and v[7] == -np.inf and v[8] == np.inf and v[9] == 5e-324
    assert v[10] == 1.2345678901234567e19  # 17 digits + dropped zeros are exact; the 18-digit token is flagged


def test_device_reader_dense_text_outgrows_the_first_buffer():
    """Two-byte tokens ("7,") are denser than the wrapper's first capacity guess: the total reported in
    val_off[R] makes it retry with the exact size; order and values must be intact across 2 KB block borders."""
    rng = np.random.default_rng(4)
    digits = rng.integers(0, 10, 30000)
    text = ("[" + ",".join(str(int(d)) for d in digits) + "]").encode()
    split = 12345 * 2 + 1  # a range border in the middle of the list (just after a comma)
    vals, off, bad = ops.json_parse_numbers(text, [(0, split), (split, len(text))])
    assert off.tolist() == [0, 12345, 30000] and bad.tolist() == [0, 0]
    assert np.array_equal(vals.cpu().numpy(), digits.astype(np.float64))


# ------------------------------------------------------------------------------------------------ generic writer
def test_render_documents_equals_json_dumps_on_nested_schemas():
    """ops.render_documents (pg_json_segments): documents shaped like the reference's stage-1/2 grid files
    (cells[].regions with five arrays per cell, arrays at indent 8), like its standard files, and odd ones
    (an array as a list element, empty arrays, a document without any array) against json.dumps(indent=2)."""
    rng = np.random.default_rng(21)
    n = 3000
    local = rng.uniform(0, 3000, (n, 4)).astype(np.float32).astype(np.float64)
    orig = local + rng.integers(0, 5000, (n, 1)).astype(np.float64)
    cls = rng.integers(0, len(NAMES), n).astype(np.float64)
    sc = rng.uniform(0, 1, n).astype(np.float32).astype(np.float64)
    name_id = cls.astype(np.int32)
    kept = rng.permutation(n).astype(np.int32)  # an arbitrary indirection, as kept_idx is
    data = [local, orig, cls, sc, name_id]
    literals = [json.dumps(x).encode("ascii") for x in NAMES]
    A = ops.DeviceArray

    def arrays(start, count):  # placeholders and what they must print
        j = kept[start:start + count]
        ph = {"boxes": A(ops.JSON_KIND_BOX4, 0, start, count), "boxes_original": A(ops.JSON_KIND_BOX4, 1, start, count),
              "classes": A(ops.JSON_KIND_SCALAR, 2, start, count), "scores": A(ops.JSON_KIND_SCALAR, 3, start, count),
              "class_names": A(ops.JSON_KIND_NAME, 4, start, count)}
        real = {"boxes": local[j].tolist(), "boxes_original": orig[j].tolist(), "classes": cls[j].tolist(),
                "scores": sc[j].tolist(), "class_names": [NAMES[i] for i in name_id[j]]}
        return ph, real

    docs, want = [], []
    at = 0
    for f in range(3):  # grid documents
        cells_ph, cells_real = [], []
        for c, cnt in enumerate([0, 1, 130, 517, 2][: 3 + f]):
            ph, real = arrays(at, cnt)
            at += cnt
            meta = {"cell_path": f"/g/é{f}_{c}.png", "cell_json_path": f"/g/{f}_{c}.json",
                    "cell_coordinates": {"x_start": 0, "y_start": 0, "x_end": 2280.6, "y_end": 3360.6}, "row": 1, "col": c + 1}
            cells_ph.append({**meta, "regions": ph})
            cells_real.append({**meta, "regions": real})
        tail = {"grid_config": {"rows": 2, "cols": 2, "overlap_percentage": 20.0}}
        docs.append({"original_image_path": f'/p/"{f}".png', "cells": cells_ph, **tail})
        want.append({"original_image_path": f'/p/"{f}".png', "cells": cells_real, **tail})
    ph, real = arrays(at, 333)  # a standard document (arrays at indent 4), one array used twice
    docs.append({"image_path": "s.png", "image_size": {"width": 1, "height": 2}, **ph, "again": ph["scores"]})
    want.append({"image_path": "s.png", "image_size": {"width": 1, "height": 2}, **real, "again": real["scores"]})
    docs.append({"k": [1, ph["classes"], "x", [ph["boxes"]]], "none": None})  # arrays as list elements, nested
    want.append({"k": [1, real["classes"], "x", [real["boxes"]]], "none": None})
    docs.append({"plain": True, "v": [1.5, "no arrays here"]})
    want.append({"plain": True, "v": [1.5, "no arrays here"]})
    got = ops.render_documents(docs, data, literals, kept_idx=kept)
    assert len(got) == len(docs)
    for g, w in zip(got, want):
        assert g == json.dumps(w, indent=2).encode("ascii")
