#!/usr/bin/env python
"""SURVEY 8f rank 1 measurement: per-tile class-agnostic float32 NMS (torchvision.ops.nms semantics,
1_doclayout_bboxes.py:217-225) for a cfg3-shaped batch: 64 pages x 16 tiles, ~625 detections per tile,
one pg_nms_merge_ex launch; torchvision's CPU kernel on the same tiles beside it."""
import json
import os
import sys
import time

import numpy as np
import torch
import torchvision

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_embeddings_b200 import _lib, ops, synth  # noqa: E402

n_pages, per_tile = 64, 625
tiles_b, tiles_s = [], []
for p in range(n_pages):
    d = synth.page_detections(8000, 6000, 4, 4, 20.0, 16 * per_tile, synth.PAGE_SEED0 + p, dups=3)
    for c in range(16):
        m = d["box_cell"] == c
        tiles_b.append(d["boxes_local"][m])
        tiles_s.append(d["scores"][m])
counts = [len(s) for s in tiles_s]
off = np.concatenate([[0], np.cumsum(counts)])
b = torch.from_numpy(np.concatenate(tiles_b)).cuda()
s = torch.from_numpy(np.concatenate(tiles_s)).cuda()
o = torch.from_numpy(off).cuda()
mode = _lib.PG_NMS_CLASS_AGNOSTIC | _lib.PG_NMS_FP32
ws = ops.NmsWorkspace(len(s), len(counts))
kept = torch.empty(len(s), dtype=torch.int32, device="cuda")
nk = torch.zeros(len(counts), dtype=torch.int32, device="cuda")
for _ in range(3):
    ops.nms_merge(b, s, None, o, 0.45, workspace=ws, kept_idx=kept, n_kept=nk, mode=mode, max_boxes_per_page=max(counts))
torch.cuda.synchronize()
assert ws.stats()["status"] == 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 20
e0.record()
for _ in range(steps):
    ops.nms_merge(b, s, None, o, 0.45, workspace=ws, kept_idx=kept, n_kept=nk, mode=mode, max_boxes_per_page=max(counts))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
# torchvision CPU on a sample of tiles (single thread, like the reference's per-tile call)
torch.set_num_threads(1)
kh, nkh = kept.cpu().numpy(), nk.cpu().numpy()
t0 = time.perf_counter()
n_cpu = 128
for t in range(n_cpu):
    ref = torchvision.ops.nms(torch.tensor(tiles_b[t], dtype=torch.float32), torch.tensor(tiles_s[t], dtype=torch.float32), 0.45)
    assert ref.tolist() == (kh[off[t]: off[t] + nkh[t]] - off[t]).tolist(), t
cpu_s = (time.perf_counter() - t0) / n_cpu
print(json.dumps({"workload": f"{n_pages} pages x 16 tiles, {int(np.mean(counts))} boxes/tile, iou 0.45, class-agnostic fp32",
                  "gpu_ms_per_launch": ms, "tiles_per_s": len(counts) / (ms * 1e-3), "pages_per_s": n_pages / (ms * 1e-3),
                  "kept_fraction": float(nkh.sum() / len(s)), "torchvision_cpu_ms_per_tile_1_thread": cpu_s * 1e3,
                  "torchvision_cpu_tiles_per_s_1_thread": 1 / cpu_s, "parity_checked_tiles": n_cpu}))
