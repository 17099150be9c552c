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
