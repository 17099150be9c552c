"""Wall time of the stage-1 command line (1_doclayout_bboxes.py drop-in) on a folder of 8000x6000 greyscale JPEG
scans: baseline JPEGs decoded on the device (default) against `--host_decode` (cv2.imread per file on host threads,
raw BGR upload).  The detector is a trivial plug-in (a few boxes per tile) so that decode + tiling + file writing are
what is timed.  Prints one JSON line.

    python scripts/bench_cli_stage1.py [--pages 16]
"""
import argparse
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2  # noqa: E402
import numpy as np  # noqa: E402

from multimodal_embeddings_b200 import cli, synth  # noqa: E402


class FewBoxes:
    def __init__(self, args=None):
        pass

    def detect_page(self, base, width, height, rows, cols, overlap, tiles, tile_names=None, tile_sizes=None, **_):
        out = []
        for (w, h) in tile_sizes:
            b = np.array([[0.1 * w, 0.1 * h, 0.4 * w, 0.3 * h], [0.5 * w, 0.5 * h, 0.9 * w, 0.8 * h]], np.float32)
            out.append({"boxes": b, "classes": np.array([1.0, 0.0], np.float32), "scores": np.array([0.9, 0.8], np.float32)})
        return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pages", type=int, default=16)
    a = ap.parse_args()
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "in")
        os.makedirs(src)
        nbytes = 0
        for p in range(a.pages):
            ok, buf = cv2.imencode(".jpg", synth.newspaper_page(8000, 6000, 300 + p % 4), [cv2.IMWRITE_JPEG_QUALITY, 95])
            with open(os.path.join(src, f"scan_{p:03d}.jpg"), "wb") as f:
                f.write(buf.tobytes())
            nbytes += len(buf)
        times = {}
        for mode, extra in (("warmup", []), ("device_decode", []), ("host_decode", ["--host_decode"])):
            out = os.path.join(tmp, mode)
            t0 = time.perf_counter()
            assert cli.main_stage1(["--input_folder", src, "--output_folder", out, "--grids", "4x4", "--detector",
                                    "bench_cli_stage1:FewBoxes"] + extra) == 0
            times[mode] = time.perf_counter() - t0
        print(json.dumps({"what": "stage-1 CLI wall time on 8000x6000 greyscale JPEG scans (full page + 4x4 grid, trivial detector)",
                          "pages": a.pages, "jpeg_mb_per_page": nbytes / a.pages / 1e6,
                          "seconds_device_decode": times["device_decode"], "seconds_host_decode": times["host_decode"],
                          "pages_per_s_device_decode": a.pages / times["device_decode"],
                          "pages_per_s_host_decode": a.pages / times["host_decode"],
                          "speedup": times["host_decode"] / times["device_decode"]}))


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    main()
