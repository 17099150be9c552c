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
    # ---- reader half: the number arrays of the same documents, text -> f64 on the device (pg_json_parse_numbers)
    from multimodal_embeddings_b200 import records
    ranges, base = [], 0
    for d in docs:
        _, _, rg = records.split_record_text(d)
        ranges.extend((x + base, y + base) for x, y in rg)
        base += len(d)
    rng_np = np.asarray(ranges, np.int64)
    bb = int(lib().pg_json_parse_block_bytes())
    blocks = (rng_np[:, 1] - rng_np[:, 0] + bb - 1) // bb
    blk_off = np.concatenate([[0], np.cumsum(blocks)]).astype(np.int64)
    d_rng, d_blk = torch.from_numpy(rng_np).cuda(), torch.from_numpy(blk_off).cuda()
    tb = int(blk_off[-1])
    r_ws_bytes = int(lib().pg_json_parse_workspace_bytes(tb))
    r_ws = torch.empty(r_ws_bytes, dtype=torch.uint8, device="cuda")
    n_vals = 6 * int(nk.sum())
    vals = torch.empty(n_vals + 16, dtype=torch.float64, device="cuda")
    val_off = torch.zeros(len(ranges) + 1, dtype=torch.int64, device="cuda")
    n_bad = torch.zeros(len(ranges), dtype=torch.int32, device="cuda")

    def read_launch():
        check(lib().pg_json_parse_numbers(ptr(out), ptr(d_rng), len(ranges), ptr(d_blk), tb, ptr(vals), vals.numel(),
                                          ptr(val_off), ptr(n_bad), 0, ptr(r_ws), r_ws_bytes,
                                          torch.cuda.current_stream().cuda_stream))

    for _ in range(a.warmup):
        read_launch()
    torch.cuda.synchronize()
    e[0].record()
    for _ in range(a.steps):
        read_launch()
    e[1].record()
    torch.cuda.synchronize()
    ms_read = e[0].elapsed_time(e[1]) / a.steps
    assert int(val_off[-1].item()) == n_vals and int(n_bad.sum().item()) == 0
    j0 = kh[off[0]: off[0] + nk[0]]
    got0 = vals[: 6 * int(nk[0])].cpu().numpy()
    want0 = np.concatenate([boxes[j0].ravel(), classes[j0], scores[j0]])
    assert np.array_equal(got0.view(np.uint64), want0.view(np.uint64))
    t0 = time.perf_counter()
    for d in docs[:sample]:
        json.loads(d)
    t_load = (time.perf_counter() - t0) / sample
    print(json.dumps({"what": "stage-3 record reader (pg_json_parse_numbers)", "pages": p, "numbers": n_vals,
                      "json_bytes_per_step": total, "device_ms_per_step": ms_read, "pages_per_s_device": p / ms_read * 1e3,
                      "text_gb_per_s_device": total / ms_read / 1e6, "cpython_json_loads_ms_per_page": t_load * 1e3,
                      "bit_identical_to_float": True, "gpu_launches_per_step": 3}))
    print(json.dumps({"what": "stage-3 record writer (pg_json_combined)", "pages": p, "boxes_in": int(n), "boxes_kept": int(nk.sum()),
                      "json_bytes_per_step": total, "device_ms_per_step": ms_dev, "pages_per_s_device": p / ms_dev * 1e3,
                      "text_gb_per_s_device": total / ms_dev / 1e6, "d2h_ms_per_step": ms_d2h,
                      "cpython_json_dumps_ms_per_page": t_py * 1e3, "cpython_pages_per_s_1core": 1.0 / t_py,
                      "byte_identical_to_cpython": True, "gpu_launches_per_step": 4}))


if __name__ == "__main__":
    main()
