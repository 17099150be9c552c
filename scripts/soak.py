"""Soak run: the same inputs through the kernels again and again, every result compared ON THE DEVICE with the first
one (64-bit sums of the raw words of every output buffer) — racecheck is not available on the GPU pool, so repeat
determinism over thousands of launches is the evidence for the mbarrier ring, the cluster kernels and the decoder's
sync rounds.  One JSON line per leg.

    python scripts/soak.py [--seconds 40]
"""
import argparse
import json
import os
import sys
import time

import cv2
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_embeddings_b200 import ops, synth  # noqa: E402
from multimodal_embeddings_b200.pipeline import PagePipeline  # noqa: E402


def word_sum(t):
    b = t.contiguous().view(torch.uint8).flatten()
    n = b.numel() // 8 * 8
    s = b[:n].view(torch.int64).sum()
    return s + b[n:].to(torch.int64).sum()


def soak(name, step, outputs, seconds):
    step()
    torch.cuda.synchronize()
    first = torch.stack([word_sum(t) for t in outputs()])
    bad = torch.zeros((), dtype=torch.int64, device="cuda")
    n, t0 = 0, time.time()
    while time.time() - t0 < seconds:
        for _ in range(20):
            step()
            bad += (torch.stack([word_sum(t) for t in outputs()]) != first).any().to(torch.int64)
            n += 1
        torch.cuda.synchronize()
    print(json.dumps({"leg": name, "launches": n, "seconds": round(time.time() - t0, 1), "results_differing_from_the_first": int(bad.item())}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=40.0)
    a = ap.parse_args()
    # whole path, two streams: tiler (mbarrier ring, bulk copies) under the box stages
    w, h, n_pages = 8000, 6000, 8
    plan = ops.TilePlan(w, h, [(4, 4)], 20.0)
    pages = plan.alloc_pages(n_pages)
    ops.synth_pages(plan, n_pages, synth.PAGE_SEED0, out=pages)
    dets = [synth.page_detections(w, h, 4, 4, 20.0, 10000, synth.PAGE_SEED0 + p) for p in range(n_pages)]
    pipe = PagePipeline(plan, n_pages)
    pipe.set_detections(dets)
    soak("cfg3 step, 8 pages (tiler || box stages)", lambda: pipe.run(pages),
         lambda: [pipe.tiles_out, pipe.kept2, pipe.n_kept2, pipe.median, pipe.centers, pipe.n_cols], a.seconds)
    # reference default grid set: chunked tiles, several grids per page
    plan30 = ops.TilePlan(w, h, [(1, 1), (2, 2), (3, 3), (4, 4)], 20.0)
    out30 = plan30.alloc_out(4)
    soak("tiler, 30 tiles per page, 4 pages", lambda: plan30.run(pages[:4], out=out30), lambda: [out30], a.seconds / 2)
    del out30, pipe
    # cluster kernels: 4 pages of 100k boxes
    rng = np.random.default_rng(1)
    big = [synth.page_detections(w, h, 4, 4, 20.0, 100000, 77 + p) for p in range(4)]
    boxes = np.concatenate([d["boxes_local"] + d["cells"][d["box_cell"]][:, [0, 1, 0, 1]] for d in big])
    scores = np.concatenate([d["scores"] for d in big])
    classes = np.concatenate([d["classes"] for d in big])
    off = [0, 100000, 200000, 300000, 400000]
    ws = ops.NmsWorkspace(len(boxes), 4, pairs_per_block=96)
    bt, st, ct = (torch.from_numpy(x).cuda() for x in (boxes, scores, classes))
    kept = torch.empty(len(boxes), dtype=torch.int32, device="cuda")
    nk = torch.empty(4, dtype=torch.int32, device="cuda")
    ot = torch.tensor(off, dtype=torch.int64, device="cuda")

    def nms():
        ops.nms_merge(bt, st, ct, ot, 0.5, max_boxes_per_page=100000, workspace=ws, kept_idx=kept, n_kept=nk)

    def nms_out():
        k = nk.clamp(min=0)
        mask = torch.arange(len(boxes), device="cuda") < 0
        for p in range(4):
            mask[off[p]: off[p + 1]] = torch.arange(100000, device="cuda") < k[p]
        return [torch.where(mask, kept, torch.zeros_like(kept)), nk]

    soak("stage-3 merge on clusters, 4 pages x 100k boxes", nms, nms_out, a.seconds / 2)
    # decoder: speculative pass + sync rounds + store
    grey = [synth.newspaper_page(4000, 3000, 5 + i) for i in range(8)]
    files = [cv2.imencode(".jpg", g, [cv2.IMWRITE_JPEG_QUALITY, 95])[1].tobytes() for g in grey]
    blob, foff = ops.pack_files(files)
    dec = ops.JpegDecoder()
    dec.set_files(blob, foff)
    dev = blob.cuda()
    outs = dec.alloc_pages()

    def decode():
        dec.set_files(blob, foff)
        dec.decode(dev, outs)

    soak("JPEG decode, 8 pages of 12 Mpixel", decode, lambda: outs, a.seconds / 2)


if __name__ == "__main__":
    main()
