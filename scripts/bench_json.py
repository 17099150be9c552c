"""SURVEY 8f rank 2 measurement: the stage-3 record writer on the cfg3 detections (64 pages x 10 k boxes,
kept set from the merge), device time of J1-J4, bytes produced, D2H time, and CPython's json.dumps(indent=2)
on the same records beside it.  Prints one JSON line.

    python scripts/bench_json.py [--pages 64] [--boxes 10000] [--steps 20]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_embeddings_b200 import _lib, ops, synth  # noqa: E402
from multimodal_embeddings_b200._lib import check, lib, ptr  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pages", type=int, default=64)
    ap.add_argument("--boxes", type=int, default=10000)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    w, h, rows, cols = 8000, 6000, 4, 4
    dets = [synth.page_detections(w, h, rows, cols, 20.0, a.boxes, synth.PAGE_SEED0 + i) for i in range(a.pages)]
    off = np.concatenate([[0], np.cumsum([len(d["scores"]) for d in dets])]).astype(np.int64)
    boxes = np.concatenate([d["boxes_local"] + d["cells"][d["box_cell"]][:, [0, 1, 0, 1]] for d in dets])
    scores = np.concatenate([d["scores"] for d in dets])
    classes = np.concatenate([d["classes"] for d in dets])
    kept, n_kept, _ = ops.nms_merge(boxes, scores, classes, off, 0.5, max_boxes_per_page=a.boxes)
    names = sorted(set(synth.class_names_of(classes)))
    by_class = {}
    for c, nm in zip(classes.tolist(), synth.class_names_of(classes)):
        by_class[c] = names.index(nm)
    name_id = np.asarray([by_class[c] for c in classes.tolist()], np.int32)
    ht = [ops.combined_head_tail(f"/corpus/page_{i:06d}.png", {"width": w, "height": h}, 0.5,
                                 [f"/corpus/2/json/page_{i:06d}_grid_4x4.json"]) for i in range(a.pages)]
    literals = [json.dumps(nm).encode("ascii") for nm in names]
    docs = ops.json_combined(boxes, classes, scores, name_id, off, [x for x, _ in ht], [y for _, y in ht], literals,
                             kept_idx=kept, n_kept=n_kept)
    total = sum(len(d) for d in docs)
    # the same launch, timed on the device with resident inputs
    dev = {k: ops._dev(v, t) for k, v, t in (("boxes", boxes, torch.float64), ("classes", classes, torch.float64),
                                           ("scores", scores, torch.float64), ("name_id", name_id, torch.int32),
                                           ("off", off, torch.int64))}
    pieces = [x for x, _ in ht] + [y for _, y in ht] + literals
    offs = np.concatenate([[0], np.cumsum([len(x) for x in pieces])]).astype(np.int64)
    p = a.pages
    text = torch.frombuffer(bytearray(b"".join(pieces) + b"\0"), dtype=torch.uint8).cuda()
    head_off, tail_off, name_off = (torch.from_numpy(offs[s].copy()).cuda() for s in (slice(0, p + 1), slice(p, 2 * p + 1), slice(2 * p, None)))
    n = len(scores)
    ws_bytes = int(lib().pg_json_workspace_bytes(n, p))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    out = torch.empty(total + 1024, dtype=torch.uint8, device="cuda")
    out_off = torch.zeros(p + 1, dtype=torch.int64, device="cuda")
    host = torch.empty(total, dtype=torch.uint8).pin_memory()

    def launch():
        check(lib().pg_json_combined(ptr(dev["boxes"]), ptr(dev["classes"]), ptr(dev["scores"]), ptr(dev["name_id"]),
                                     ptr(kept), ptr(dev["off"]), ptr(n_kept), p, n, a.boxes, ptr(text), ptr(head_off),
                                     ptr(tail_off), ptr(name_off), ptr(out), out.numel(), ptr(out_off), ptr(ws), ws_bytes,
                                     torch.cuda.current_stream().cuda_stream))

    for _ in range(a.warmup):
        launch()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    torch.cuda.synchronize()
    e[0].record()
    for _ in range(a.steps):
        launch()
    e[1].record()
    for _ in range(a.steps):
        host.copy_(out[:total], non_blocking=True)
    e[2].record()
    torch.cuda.synchronize()
    ms_dev, ms_d2h = e[0].elapsed_time(e[1]) / a.steps, e[1].elapsed_time(e[2]) / a.steps
    assert host.numpy().tobytes() == b"".join(docs)
    # CPython on the same records (a sample of pages)
    kh, nk = kept.cpu().numpy(), n_kept.cpu().numpy()
    sample = min(4, p)
    recs = []
    for i in range(sample):
        j = kh[off[i]: off[i] + nk[i]]
        recs.append({"image_path": f"/corpus/page_{i:06d}.png", "image_size": {"width": w, "height": h},
                     "parameters": {"iou_threshold": 0.5}, "boxes": boxes[j].tolist(), "classes": classes[j].tolist(),
                     "scores": scores[j].tolist(), "class_names": synth.class_names_of(classes[j]),
                     "source_jsons": [f"/corpus/2/json/page_{i:06d}_grid_4x4.json"]})
    t0 = time.perf_counter()
    py = [json.dumps(r, indent=2).encode("ascii") for r in recs]
    t_py = (time.perf_counter() - t0) / sample
    assert py == docs[:sample]
    print(json.dumps({"what": "stage-3 record writer (pg_json_combined)", "pages": p, "boxes_in": int(n), "boxes_kept": int(nk.sum()),
                      "json_bytes_per_step": total, "device_ms_per_step": ms_dev, "pages_per_s_device": p / ms_dev * 1e3,
                      "text_gb_per_s_device": total / ms_dev / 1e6, "d2h_ms_per_step": ms_d2h,
                      "cpython_json_dumps_ms_per_page": t_py * 1e3, "cpython_pages_per_s_1core": 1.0 / t_py,
                      "byte_identical_to_cpython": True, "gpu_launches_per_step": 4}))


if __name__ == "__main__":
    main()
