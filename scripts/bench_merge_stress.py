#!/usr/bin/env python
"""BASELINE configs[3]: dense-box merge stress — 100k boxes/page through the class-aware NMS on 1 B200.
Stage 3 only (no pixels).  Prints one JSON object: time per page, candidate block pairs, box pairs tested,
pairs/s.  Parity of this size against the oracle is covered by tests/test_gpu_parity.py."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_embeddings_b200 import ops, synth  # noqa: E402

n_pages = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n_boxes = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
dets = [synth.page_detections(8000, 6000, 4, 4, 20.0, n_boxes, synth.PAGE_SEED0 + 900 + p, dups=6) for p in range(n_pages)]
boxes = np.concatenate([d["boxes_local"] + d["cells"][d["box_cell"]][:, [0, 1, 0, 1]] for d in dets])
scores = np.concatenate([d["scores"] for d in dets])
classes = np.concatenate([d["classes"] for d in dets])
off = np.arange(n_pages + 1) * n_boxes
b, s, c = (torch.from_numpy(x).cuda() for x in (boxes, scores, classes))
o = torch.from_numpy(off).cuda()
ws = ops.NmsWorkspace(len(boxes), n_pages, pairs_per_block=96)
kept = torch.empty(len(boxes), dtype=torch.int32, device="cuda")
nk = torch.zeros(n_pages, dtype=torch.int32, device="cuda")
for _ in range(3):
    ops.nms_merge(b, s, c, o, 0.5, workspace=ws, kept_idx=kept, n_kept=nk, max_boxes_per_page=n_boxes)
torch.cuda.synchronize()
st = ws.stats()
assert st["status"] == 0, st
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 10
e0.record()
for _ in range(steps):
    ops.nms_merge(b, s, c, o, 0.5, workspace=ws, kept_idx=kept, n_kept=nk, max_boxes_per_page=n_boxes)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(json.dumps({"workload": f"cfg4: {n_pages} pages x {n_boxes} boxes, stage 3 (pg_nms_merge) only", "ms_per_launch": ms,
                  "ms_per_page": ms / n_pages, "pages_per_s": n_pages / (ms * 1e-3), "kept_per_page": float(nk.float().mean().item()),
                  "candidate_block_pairs": st["candidate_block_pairs"], "box_pairs_tested": st["box_pairs_tested"],
                  "dense_pairs_per_page": n_boxes * (n_boxes - 1) // 2, "resolve_rounds": st["rounds"],
                  "pairs_tested_per_s": st["box_pairs_tested"] / (ms * 1e-3)}))
