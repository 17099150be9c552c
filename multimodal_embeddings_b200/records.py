"""Binary sidecar of the stage-3 record ("... or a binary sidecar").

`<base>_combined.json` stays what the reference writes (3_combine_grids.py:441-443) — it is what the
reference's own stages 4 and 5 read (4:103-151, 5:226-240).  Next to it, on request, stage 3 drops
`<base>_combined.pgrec`: the same record with the numbers as raw little-endian arrays, so that the next
stages of THIS implementation skip a megabyte of decimal text per page.  The sidecar is only trusted
when it is at least as new as the JSON file and names the same box count.

Layout: b"PGREC1\\0\\0" | u64 header_len | header (UTF-8 JSON: image_path, image_size, parameters,
source_jsons, names, n) | pad to 8 | boxes f64[n,4] | classes f64[n] | scores f64[n] | name_id i32[n].
"""
from __future__ import annotations

import json
import os
import struct
from typing import Optional, Sequence

import numpy as np

MAGIC = b"PGREC1\0\0"
SUFFIX = ".pgrec"


def sidecar_path(json_path: str) -> str:
    return os.path.splitext(json_path)[0] + SUFFIX


def write_sidecar(json_path: str, image_path, image_size, parameters: dict, source_jsons: Sequence[str],
                  boxes: np.ndarray, classes: np.ndarray, scores: np.ndarray, names: Sequence[str],
                  name_id: np.ndarray) -> str:
    boxes = np.ascontiguousarray(boxes, "<f8").reshape(-1, 4)
    n = boxes.shape[0]
    classes, scores = np.ascontiguousarray(classes, "<f8"), np.ascontiguousarray(scores, "<f8")
    name_id = np.ascontiguousarray(name_id, "<i4")
    assert classes.shape == (n,) and scores.shape == (n,) and name_id.shape == (n,)
    header = json.dumps({"image_path": image_path, "image_size": image_size, "parameters": parameters,
                         "source_jsons": list(source_jsons), "names": list(names), "n": n}).encode("utf-8")
    pad = (-(len(MAGIC) + 8 + len(header))) % 8
    path = sidecar_path(json_path)
    tmp = f"{path}.tmp.{os.getpid()}"
    with open(tmp, "wb") as f:
        f.write(MAGIC + struct.pack("<Q", len(header)) + header + b"\0" * pad)
        f.write(boxes.tobytes() + classes.tobytes() + scores.tobytes() + name_id.tobytes())
    os.replace(tmp, path)
    return path


def read_sidecar(json_path: str) -> Optional[dict]:
    """The record as `json.load` would return it, except that boxes/classes/scores are numpy arrays
    (`boxes` [n,4]).  None when there is no usable sidecar."""
    path = sidecar_path(json_path)
    try:
        if os.path.getmtime(path) < os.path.getmtime(json_path):
            return None
        with open(path, "rb") as f:
            raw = f.read()
    except OSError:
        return None
    if raw[:8] != MAGIC or len(raw) < 16:
        return None
    (hlen,) = struct.unpack_from("<Q", raw, 8)
    try:
        h = json.loads(raw[16:16 + hlen].decode("utf-8"))
    except ValueError:
        return None
    if not isinstance(h, dict) or any(k not in h for k in ("image_path", "image_size", "parameters", "source_jsons",
                                                           "names", "n")) or not isinstance(h["n"], int):
        return None
    n = int(h["n"])
    at = (16 + hlen + 7) // 8 * 8
    if n < 0 or len(raw) != at + n * (32 + 8 + 8 + 4):
        return None
    boxes = np.frombuffer(raw, "<f8", 4 * n, at).reshape(n, 4)
    classes = np.frombuffer(raw, "<f8", n, at + 32 * n)
    scores = np.frombuffer(raw, "<f8", n, at + 40 * n)
    name_id = np.frombuffer(raw, "<i4", n, at + 48 * n)
    names = h["names"]
    if n and (name_id.min() < 0 or name_id.max() >= len(names)):
        return None
    return {"image_path": h["image_path"], "image_size": h["image_size"], "parameters": h["parameters"],
            "boxes": boxes, "classes": classes, "scores": scores, "class_names": [names[i] for i in name_id.tolist()],
            "source_jsons": h["source_jsons"]}


def load_record(json_path: str) -> dict:
    """Stage-3 record for stages 4 and 5: the sidecar when it is usable, the JSON text otherwise."""
    rec = read_sidecar(json_path)
    if rec is not None:
        return rec
    with open(json_path) as f:
        return json.load(f)


# ------------------------------------------------------------------------------------------------
# Text path: the record as the reference wrote it, numbers converted on the device (pg_json_parse_numbers)
_ARRAY_KEYS = (b'\n  "boxes": [', b'\n  "classes": [', b'\n  "scores": [', b'\n  "class_names": [')


def split_record_text(raw: bytes):
    """Locate the three number arrays of a stage-3 record laid out by `json.dump(indent=2)` with the
    reference's key order (3_combine_grids.py:282-291).  Returns (head dict, tail dict, [(begin, end)] * 3) or
    None when the text is laid out differently (then json.loads is the reader).  JSON strings cannot hold a
    raw newline, so a key found at the start of a line is a key."""
    pos, at = [], 0
    for k in _ARRAY_KEYS:
        i = raw.find(k, at)
        if i < 0:
            return None
        pos.append(i)
        at = i + len(k)
    i_boxes, i_classes, i_scores, i_names = pos
    try:
        head = json.loads(raw[:i_boxes].rstrip().rstrip(b",") + b"}")
        tail = json.loads(b"{" + raw[i_names + 1:])
    except ValueError:
        return None
    if not isinstance(head, dict) or not isinstance(tail, dict) or "class_names" not in tail:
        return None
    ranges = [(i_boxes + len(_ARRAY_KEYS[0]), i_classes), (i_classes + len(_ARRAY_KEYS[1]), i_scores),
              (i_scores + len(_ARRAY_KEYS[2]), i_names)]
    return head, tail, ranges


def load_records(json_paths: Sequence[str]) -> dict:
    """path -> record for a batch of stage-3 files, by the cheapest valid route per file: the binary sidecar;
    else the JSON text with its number arrays converted on the GPU in ONE call for the whole batch; else
    (unusual layout, a token the device does not convert, no CUDA, PG_PYTHON_JSON=1) `json.load`.
    Records from the first two routes carry numpy arrays for boxes [n,4] / classes / scores."""
    out, pending = {}, []
    use_device = os.environ.get("PG_PYTHON_JSON") != "1"
    if use_device:
        try:
            import torch
            use_device = torch.cuda.is_available()
        except ImportError:
            use_device = False
    for path in json_paths:
        rec = read_sidecar(path)
        if rec is not None:
            out[path] = rec
            continue
        with open(path, "rb") as f:
            raw = f.read()
        parts = split_record_text(raw) if use_device else None
        if parts is None:
            out[path] = json.loads(raw)
        else:
            pending.append((path, raw, parts))
    if pending:
        from . import ops
        blob, ranges, base = [], [], 0
        for _, raw, (_, _, rg) in pending:
            blob.append(raw)
            ranges.extend((a + base, b + base) for a, b in rg)
            base += len(raw)
        values, off, bad = ops.json_parse_numbers(b"".join(blob), ranges)
        values = values.cpu().numpy()
        for k, (path, raw, (head, tail, _)) in enumerate(pending):
            nb, nc, ns = (int(off[3 * k + j + 1] - off[3 * k + j]) for j in range(3))
            names = tail["class_names"]
            if bad[3 * k: 3 * k + 3].any() or nb != 4 * nc or nc != ns or nc != len(names):
                out[path] = json.loads(raw)  # ragged or exotic record: let CPython decide what it means
                continue
            rec = dict(head)
            rec["boxes"] = values[off[3 * k]: off[3 * k + 1]].reshape(-1, 4)
            rec["classes"] = values[off[3 * k + 1]: off[3 * k + 2]]
            rec["scores"] = values[off[3 * k + 2]: off[3 * k + 3]]
            rec.update(tail)
            out[path] = rec
    return {path: out[path] for path in json_paths}  # the caller's order


# ------------------------------------------------------------------------------------------------
# Stage-3 INPUTS (the stage-1/2 documents pooled by combine_boxes_for_image, 3_combine_grids.py:222-270),
# same route: arrays located on the host, numbers converted on the device.
_REGIONS_KEY = b'\n      "regions": {'
_REGION_KEYS = (b'\n        "boxes_original": [', b'\n        "classes": [', b'\n        "scores": [',
                b'\n        "class_names": [')
_REGION_LIST_END = b'\n        ]'


def _line_value(raw: bytes, key: bytes):
    """Value of a top-level `"key": <scalar>` line of an indent=2 document (None if absent / not a scalar line)."""
    i = raw.find(b'\n  "' + key + b'": ')
    if i < 0:
        return None
    j = raw.find(b"\n", i + 1)
    try:
        return json.loads(raw[i + len(key) + 7: j if j >= 0 else len(raw)].rstrip().rstrip(b","))
    except ValueError:
        return None


def split_grid_text(raw: bytes):
    """A stage-1/2 grid document (`cells[].regions`, json.dump(indent=2)): the byte ranges of every cell's
    boxes_original / classes / scores arrays in cell order, and the cells' class_names lists.  None unless every
    `regions` object holds the five arrays in the order stage 1 writes them (1_doclayout_bboxes.py:620-640)."""
    if raw.find(b'\n  "cells": [') < 0:
        return None
    starts, at = [], 0
    while True:
        i = raw.find(_REGIONS_KEY, at)
        if i < 0:
            break
        starts.append(i)
        at = i + len(_REGIONS_KEY)
    # (each key is looked up inside its own `regions` object only — `limit` below — so a missing or misplaced
    # array fails the file; counting the keys over the whole text as a cross-check cost more than the parse)
    ranges, names = [], []
    for r, s0 in enumerate(starts):
        limit = starts[r + 1] if r + 1 < len(starts) else len(raw)
        pos, at = [], s0
        for k in _REGION_KEYS:
            i = raw.find(k, at, limit)
            if i < 0:
                return None
            pos.append(i)
            at = i + len(k)
        i_bo, i_cl, i_sc, i_nm = pos
        lst = i_nm + len(_REGION_KEYS[3]) - 1  # the '[' of class_names
        if raw[lst + 1: lst + 2] == b"]":
            end = lst + 2
        else:
            j = raw.find(_REGION_LIST_END, lst, limit)
            if j < 0:
                return None
            end = j + len(_REGION_LIST_END)
        try:
            nm = json.loads(raw[lst:end])
        except ValueError:
            return None
        ranges += [(i_bo + len(_REGION_KEYS[0]), i_cl), (i_cl + len(_REGION_KEYS[1]), i_sc),
                   (i_sc + len(_REGION_KEYS[2]), i_nm)]
        names.append(nm)
    return _line_value(raw, b"original_image_path"), ranges, names


def load_pool_inputs(json_paths: Sequence[str]) -> dict:
    """path -> what combine_boxes_for_image takes from that file, numbers converted on the device in ONE call for
    all files: {"grid": bool, "image_path", "image_size", "boxes" f64 [n,4], "scores", "classes", "class_names"}.
    A path is absent from the result when its text is not laid out as stage 1/2 write it, holds integer literals
    among the numbers (CPython must read those: 5 stays 5), is ragged, or no CUDA device is there — the caller
    reads such files with json.load."""
    try:
        import torch
        if not torch.cuda.is_available() or os.environ.get("PG_PYTHON_JSON") == "1":
            return {}
    except ImportError:
        return {}
    from . import ops
    todo, blob, ranges, base = [], [], [], 0
    for path in json_paths:
        try:
            with open(path, "rb") as f:
                raw = f.read()
        except OSError:
            continue
        if b'\n  "cells": [' in raw:
            parts = split_grid_text(raw)
            if parts is None:
                continue
            image_path, rg, names = parts
            todo.append((path, True, image_path, None, names, len(rg) // 3))
        else:
            parts = split_record_text(raw) if b'"boxes_original"' not in raw else None
            if parts is None:
                continue
            head, tail, rg = parts
            todo.append((path, False, head.get("image_path"), head.get("image_size"), [tail["class_names"]], 1))
        ranges.extend((a + base, b + base) for a, b in rg)
        blob.append(raw)
        base += len(raw)
    if not todo:
        return {}
    values, off, bad = ops.json_parse_numbers(b"".join(blob), ranges, int_literals_to_host=True)
    values = values.cpu().numpy()
    out, r = {}, 0
    for path, grid, image_path, image_size, names, n_regions in todo:
        ok, boxes, classes, scores, flat = True, [], [], [], []
        for k in range(n_regions):
            nb, nc, ns = (int(off[r + j + 1] - off[r + j]) for j in range(3))
            nm = names[k]
            if bad[r: r + 3].any() or nb != 4 * nc or nc != ns or not isinstance(nm, list) or nc != len(nm):
                ok = False
            boxes.append(values[off[r]: off[r + 1]])
            classes.append(values[off[r + 1]: off[r + 2]])
            scores.append(values[off[r + 2]: off[r + 3]])
            flat.extend(nm if isinstance(nm, list) else [])
            r += 3
        if ok:
            cat = (lambda xs: np.concatenate(xs) if xs else np.zeros(0))
            out[path] = {"grid": grid, "image_path": image_path, "image_size": image_size,
                         "boxes": cat(boxes).reshape(-1, 4), "classes": cat(classes), "scores": cat(scores),
                         "class_names": flat}
    return out


# ------------------------------------------------------------------------------------------------
# Stage 2 on grid documents, batched: read -> edge filter -> write, numbers never visiting the host
_REGION_KEYS5 = (b'\n        "boxes": [',) + _REGION_KEYS


def split_grid_text_full(raw: bytes):
    """Like split_grid_text, with the cell-local `boxes` as well: (skeleton dict — the document with the five
    arrays of every `regions` emptied —, [(begin, end)] * 4 per cell in the order boxes, boxes_original,
    classes, scores, class_names lists per cell), or None."""
    if raw.find(b'\n  "cells": [') < 0:
        return None
    starts, at = [], 0
    while True:
        i = raw.find(_REGIONS_KEY, at)
        if i < 0:
            break
        starts.append(i)
        at = i + len(_REGIONS_KEY)
    ranges, names, skel, prev = [], [], [], 0
    empty = b"],".join(_REGION_KEYS5) + b"]"
    for r, s0 in enumerate(starts):
        limit = starts[r + 1] if r + 1 < len(starts) else len(raw)
        pos, at = [], s0
        for k in _REGION_KEYS5:
            i = raw.find(k, at, limit)
            if i < 0:
                return None
            pos.append(i)
            at = i + len(k)
        i_b, i_bo, i_cl, i_sc, i_nm = pos
        lst = i_nm + len(_REGION_KEYS5[4]) - 1
        if raw[lst + 1: lst + 2] == b"]":
            end = lst + 2
        else:
            j = raw.find(_REGION_LIST_END, lst, limit)
            if j < 0:
                return None
            end = j + len(_REGION_LIST_END)
        try:
            nm = json.loads(raw[lst:end])
        except ValueError:
            return None
        ranges += [(i_b + len(_REGION_KEYS5[0]), i_bo), (i_bo + len(_REGION_KEYS5[1]), i_cl),
                   (i_cl + len(_REGION_KEYS5[2]), i_sc), (i_sc + len(_REGION_KEYS5[3]), i_nm)]
        names.append(nm)
        skel += [raw[prev:i_b], empty]
        prev = end
    skel.append(raw[prev:])
    try:
        skeleton = json.loads(b"".join(skel))
    except ValueError:
        return None
    cells = skeleton.get("cells") if isinstance(skeleton, dict) else None
    if not isinstance(cells, list) or len(cells) != len(starts) or any(
            not isinstance(c, dict) or "regions" not in c for c in cells):
        return None
    return skeleton, ranges, names


def filter_grid_files(json_paths: Sequence[str], threshold, page_size_of, cell_tuple, standard_ok=None) -> dict:
    """Stage 2 (filter_grid_info, 2_edge_box_filter.py:148-237, and its json.dump :485-487) for a batch of grid
    documents in three device calls: pg_json_parse_numbers for every array of every cell of every file,
    pg_edge_filter with each cell as its own page, pg_json_segments for the filtered documents.  Returns
    path -> the bytes of the output file for the files it could take; the others (layout, integer literals,
    unknown page size, ragged arrays) are left to the per-file path.  `page_size_of(skeleton)` -> (W, H) or None;
    `cell_tuple(cell_coordinates, W, H)` -> [x_start, y_start, x_end, y_end].
    With `standard_ok(doc_without_arrays) -> bool`, full-page documents (no `cells`, no `cell_coordinates`: stage 2
    copies them, "Non-grid image detected, not filtering any boxes", 2:92-100) take the same route: read and
    re-written on the device, every box kept."""
    try:
        import torch
        if not torch.cuda.is_available() or os.environ.get("PG_PYTHON_JSON") == "1":
            return {}
    except ImportError:
        return {}
    from . import ops
    files, blob, base = [], [], 0
    for path in json_paths:
        try:
            with open(path, "rb") as f:
                raw = f.read()
        except OSError:
            continue
        if b'\n  "cells": [' not in raw:
            if standard_ok is None or b'"boxes_original"' in raw or b'"cell_coordinates"' in raw:
                continue
            parts = split_record_text(raw)
            if parts is None:
                continue
            head, tail, rg = parts
            tail = dict(tail)
            names = tail.pop("class_names")
            try:
                if not standard_ok({**head, **tail}):
                    continue
            except Exception:
                continue
            # one pseudo-cell; its `boxes` array stands in for both box kinds so that the arrays stay aligned
            files.append((path, ("standard", head, tail), [(a + base, b + base) for a, b in (rg[0], rg[0], rg[1], rg[2])],
                          [names], (1, 1), [[0, 0, 1, 1]]))
            blob.append(raw)
            base += len(raw)
            continue
        parts = split_grid_text_full(raw)
        if parts is None:
            continue
        skeleton, rg, names = parts
        if not ("grid_config" in skeleton or "grid_info" in skeleton) or "original_image_path" not in skeleton:
            continue
        try:
            size = page_size_of(skeleton)
            tuples = [cell_tuple(c["cell_coordinates"], size[0], size[1]) for c in skeleton["cells"]] if size else None
            meta = [(c["cell_path"], c["cell_json_path"], c["cell_coordinates"]) for c in skeleton["cells"]]
        except Exception:
            continue
        if not size or any(len(t) != 4 for t in tuples) or not meta and skeleton["cells"]:
            continue
        files.append((path, skeleton, [(a + base, b + base) for a, b in rg], names, size, tuples))
        blob.append(raw)
        base += len(raw)
    if not files:
        return {}
    # ranges kind-major, so that the converted numbers come out as four contiguous arrays over all cells
    ranges = [r for kind in range(4) for f in files for r in f[2][kind::4]]
    n_cells = len(ranges) // 4
    values, off, bad = ops.json_parse_numbers(b"".join(blob), ranges, int_literals_to_host=True)
    cnt = np.diff(off).reshape(4, n_cells)
    ok_cell = (bad.reshape(4, n_cells) == 0).all(0) & (cnt[0] == 4 * cnt[2]) & (cnt[1] == 4 * cnt[2]) & (cnt[3] == cnt[2])
    flat_names, c = [], 0
    for f in files:
        for nm in f[3]:
            if not isinstance(nm, list) or len(nm) != cnt[2][c] or any(not isinstance(x, str) for x in nm):
                ok_cell[c] = False
            flat_names.append(nm if isinstance(nm, list) else [])
            c += 1
    # per-file verdict; a declined file's cells simply stay unused in the arrays
    take, c = [], 0
    for f in files:
        k = len(f[3])
        take.append(bool(ok_cell[c:c + k].all()))
        c += k
    if not all(take):  # a ragged / exotic file would shift the per-kind arrays: redo the batch without it
        good = [f[0] for f, t in zip(files, take) if t]
        return filter_grid_files(good, threshold, page_size_of, cell_tuple) if good else {}
    n = int(cnt[2].sum())
    boxes_local = values[off[0]: off[n_cells]].view(-1, 4)
    boxes_orig = values[off[n_cells]: off[2 * n_cells]].view(-1, 4)
    classes = values[off[2 * n_cells]: off[3 * n_cells]]
    scores = values[off[3 * n_cells]: off[4 * n_cells]]
    cell_off = np.concatenate([[0], np.cumsum(cnt[2])]).astype(np.int64)
    table, name_id = {}, np.zeros(max(n, 1), np.int32)
    for ci, nm in enumerate(flat_names):
        name_id[cell_off[ci]: cell_off[ci + 1]] = [table.setdefault(x, len(table)) for x in nm]
    page_wh = np.asarray([f[4] for f in files for _ in f[3]], np.int32).reshape(-1, 2)
    cells = np.asarray([t for f in files for t in f[5]], np.float64).reshape(-1, 4)
    box_cell = np.repeat(np.arange(n_cells, dtype=np.int32), cnt[2])
    if n:
        _, _, kept, n_kept = ops.edge_filter(boxes_orig, box_cell, cells, page_wh, cell_off, threshold,
                                             boxes_are_local=False, want_boxes_page=False)
        n_kept_h = n_kept.cpu().numpy().copy()
        c = 0
        for f in files:  # full-page documents keep every box, in order
            if isinstance(f[1], tuple):
                a, b = int(cell_off[c]), int(cell_off[c + 1])
                kept[a:b] = torch.arange(a, b, dtype=torch.int32, device=kept.device)
                n_kept_h[c] = b - a
            c += len(f[3])
    else:
        kept, n_kept_h = np.zeros(1, np.int32), np.zeros(n_cells, np.int32)
    A = ops.DeviceArray
    docs, doc_paths, c = [], [], 0
    for f, good in zip(files, take):
        path, skeleton = f[0], f[1]
        if good and isinstance(skeleton, tuple):
            _, head, tail = skeleton
            s, k = int(cell_off[c]), int(n_kept_h[c])
            docs.append({**head, "boxes": A(ops.JSON_KIND_BOX4, 0, s, k), "classes": A(ops.JSON_KIND_SCALAR, 2, s, k),
                         "scores": A(ops.JSON_KIND_SCALAR, 3, s, k), "class_names": A(ops.JSON_KIND_NAME, 4, s, k), **tail})
            doc_paths.append(path)
        elif good:
            filtered = {"original_image_path": skeleton["original_image_path"], "cells": []}
            if "grid_config" in skeleton:
                filtered["grid_config"] = skeleton["grid_config"]
            for j, cell in enumerate(skeleton["cells"]):
                s, k = int(cell_off[c + j]), int(n_kept_h[c + j])
                filtered["cells"].append({
                    "cell_path": cell["cell_path"], "cell_json_path": cell["cell_json_path"],
                    "cell_coordinates": cell["cell_coordinates"], "row": cell.get("row", 0), "col": cell.get("col", 0),
                    "regions": {"boxes": A(ops.JSON_KIND_BOX4, 0, s, k), "boxes_original": A(ops.JSON_KIND_BOX4, 1, s, k),
                                "classes": A(ops.JSON_KIND_SCALAR, 2, s, k), "scores": A(ops.JSON_KIND_SCALAR, 3, s, k),
                                "class_names": A(ops.JSON_KIND_NAME, 4, s, k)}})
            docs.append(filtered)
            doc_paths.append(path)
        c += len(f[3])
    literals = [json.dumps(x).encode("ascii") for x in table] or [b'""']
    texts = ops.render_documents(docs, [boxes_local, boxes_orig, classes, scores, name_id], literals, kept_idx=kept)
    return dict(zip(doc_paths, texts))
