"""A deterministic stand-in for the DocLayout-YOLO network, shared by the golden generator and the tests.

TEST INFRASTRUCTURE.  The raw (pre-NMS) detections of an image are a pure function of the image's file name
and size, so the UNMODIFIED reference stage-1 driver (oracle/gen_golden_stage1.py wraps `raw_detections` as the
object `YOLODocumentLayoutDetector.model`, 1_doclayout_bboxes.py:178-179,205-215) and this repository's stage-1
command line (`StubPlugin`, the `--detector module:factory` hook of cli.main_stage1) see the same detector.
"""
from __future__ import annotations

import os
import zlib

import numpy as np

from multimodal_embeddings_b200 import synth

PAGES = [  # (file name, W, H, pixel seed)
    ("Daily Argus 1899 page_0001.png", 1180, 1640, 71),
    ("Weekly Post 1912 page_0002.png", 1501, 1003, 72),
    ("Courier 1875 page_0003.bmp", 777, 1001, 73),
]
GRIDS = "2x2,3x3"
OVERLAP = 20.0


def raw_detections(file_name: str, width: int, height: int):
    """boxes f32 [n,4] (xyxy in the image's own pixels), classes f32 [n], scores f32 [n] — what the network's
    `predict(...)[0].boxes.{xyxy,cls,conf}` would hold before the per-tile NMS of 1:217-225."""
    seed = zlib.crc32(os.path.basename(file_name).encode("utf-8")) & 0x7FFFFFFF
    n = max(6, int(width * height / 9000))
    d = synth.page_detections(width, height, 1, 1, 20.0, n, seed, dups=3)
    return (d["boxes_local"].astype(np.float32), d["classes"].astype(np.float32), d["scores"].astype(np.float32))


class StubPlugin:
    """Detector plug-in for cli.main_stage1 (`--detector stub_detector:StubPlugin`): one call per grid of a page,
    raw detections per tile; the command line applies the per-tile NMS and writes the files."""

    def __init__(self, args=None):
        self.args = args

    def detect_page(self, base, width, height, rows, cols, overlap, tiles, tile_names=None, tile_sizes=None, **_):
        out = []
        for name, (w, h) in zip(tile_names, tile_sizes):
            b, c, s = raw_detections(name, w, h)
            out.append({"boxes": b, "classes": c, "scores": s})
        return out
