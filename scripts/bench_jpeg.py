"""Decoder timing on the GPU box: greyscale newspaper-like pages (8000x6000) as JPEG -> pages in HBM.
Prints one JSON line per (quality, chunk size): bytes/page, ms/page, pages/s, sync rounds used; then the
one-channel tiler on the decoded pages and the host->device rate of the compressed bytes."""
import json
import sys
import os
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
import numpy as np
import torch

from multimodal_embeddings_b200 import ops, synth


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    once = len(sys.argv) > 2 and sys.argv[2] == "once"  # one decode of the q95 batch (for an ncu launch list)
    reps = 5
    t0 = time.time()
    pages = [synth.newspaper_page(8000, 6000, 100 + i) for i in range(n)]
    print(json.dumps({"generated_pages": n, "seconds": round(time.time() - t0, 1)}), flush=True)
    for q, rst in (((95, 0),) if once else ((95, 0), (75, 0), (95, 1000))):
        files = [cv2.imencode(".jpg", p, [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_RST_INTERVAL, rst])[1].tobytes() for p in pages]
        t0 = time.time()
        ref = cv2.imdecode(np.frombuffer(files[0], np.uint8), cv2.IMREAD_COLOR)
        cv2_s = time.time() - t0
        blob, off = ops.pack_files(files)
        dev = blob.cuda()
        for chunk in ((0,) if once else (0, 128, 256, 512, 1024)):
            dec = ops.JpegDecoder(chunk_bytes=chunk, sync_rounds=4)
            dec.set_files(blob, off)
            outs = dec.alloc_pages()
            dec.decode(dev, outs)
            torch.cuda.synchronize()
            st = dec.check()
            ok = bool(np.array_equal(outs[0][:, :8000].cpu().numpy(), ref[..., 0]))
            if once:
                return
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                dec.set_files(blob, off)
                dec.decode(dev, outs)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / reps
            print(json.dumps({"quality": q, "restart_interval": rst, "chunk_bytes": dec.chunk_bytes, "mb_per_page": round(len(files[0]) / 1e6, 2),
                              "ms_per_batch": round(ms, 3), "pages_per_s": round(n / ms * 1e3, 1), "equal_cv2": ok,
                              "rounds_cfg": dec.sync_rounds, **st, "cv2_imdecode_s_per_page": round(cv2_s, 3)}), flush=True)
        # H2D of the compressed bytes
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            dev.copy_(blob, non_blocking=True)
        b.record()
        torch.cuda.synchronize()
        print(json.dumps({"quality": q, "h2d_gb_per_s": round(blob.numel() * reps / a.elapsed_time(b) / 1e6, 1)}), flush=True)
    # one-channel tiler on decoded pages vs three-channel tiler
    plan1 = ops.TilePlan(8000, 6000, [(4, 4)], 20.0, channels=1)
    plan3 = ops.TilePlan(8000, 6000, [(4, 4)], 20.0)
    p1 = torch.stack([o for o in outs])
    p3 = plan3.alloc_pages(n)
    ops.synth_pages(plan3, n, 1)
    for name, plan, pg in (("grey", plan1, p1), ("bgr", plan3, p3)):
        out = plan.alloc_out(n)
        for _ in range(3):
            plan.run(pg, out=out)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            plan.run(pg, out=out)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        print(json.dumps({"tiler": name, "pages": n, "ms": round(ms, 3), "pages_per_s": round(n / ms * 1e3),
                          "algorithmic_gb_per_s": round(plan.algorithmic_bytes * n / ms / 1e6, 1)}), flush=True)


if __name__ == "__main__":
    main()
