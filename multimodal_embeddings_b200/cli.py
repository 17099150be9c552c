"""Drop-in command lines for the numbered scripts of calhounpaul/multimodal_embeddings
(stages 1-5 of run.sh:60-70), same argv and same on-disk JSON schemas, with the geometry done by
libpagegeom.so.  Every stage batches all pages it finds into single kernel launches.

    1_doclayout_bboxes.py      -> main_stage1   (tiler + pluggable detector; 1:682-785)
    2_edge_box_filter.py       -> main_stage2   (2:670-766)
    3_combine_grids.py         -> main_stage3   (3:403-458)
    4_extract_median_widths.py -> main_stage4   (4:227-292)
    5_detect_column_centers.py -> main_stage5   (5:541-588)

Not reproduced (out of scope): the DocLayout-YOLO network itself —
stage 1 hands the letterboxed fp16 tiles to a detector object (synthetic or replayed
detections ship here) — and the JPEG visualisations (--viz_alpha is accepted and ignored).
Extra flags, all optional: --no_image_check lets stages 2-5 run on JSON trees whose page images
are absent (page size then comes from the JSON), --detections/--replay_folder/--boxes_per_page
choose stage 1's detection source, --sidecar (stage 3) also writes the binary record.
Under `torchrun --nproc-per-node N` every stage shards its sorted work list by rank (one process per GPU, no
communication); the union of the ranks' output files equals the single-process output.
"""
from __future__ import annotations

import argparse
import glob
import json
import logging
import os
import re
import sys
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .records import load_records, write_sidecar

FMT = "%(asctime)s - %(name)s - %(levelname)s - %(message)s"
IMAGE_EXTENSIONS = (".jpg", ".jpeg", ".png", ".bmp", ".tiff", ".tif", ".webp")


def _logger(name: str) -> logging.Logger:
    lg = logging.getLogger(name)  # same logger names as the reference (1:43, 2:26, 3:28, 4:31, 5:56)
    if not lg.handlers:
        h = logging.StreamHandler()
        h.setFormatter(logging.Formatter(FMT))
        lg.addHandler(h)
        lg.setLevel(logging.INFO)
    return lg


def _my_share(items: list) -> list:
    """Multi-GPU launches (`torchrun --nproc-per-node N <stage>.py ...` sets RANK / WORLD_SIZE / LOCAL_RANK):
    pages are independent (1:749, 2:206, 3:431, 4:263, 5:569), so rank r simply takes the r-th contiguous block of
    the stage's sorted work list and the union of the ranks' output files is the single-process output — no
    communication.  Also binds the process to its GPU."""
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    if world <= 1:
        return items
    try:
        import torch
        if torch.cuda.is_available():
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)) % torch.cuda.device_count())
    except ImportError:
        pass
    from .pipeline import shard_pages
    r = shard_pages(len(items), rank, world)
    return items[r.start:r.stop]


def _dump(obj, path):
    with open(path, "w") as f:
        json.dump(obj, f, indent=2)


def _image_size(path) -> Optional[Tuple[int, int]]:
    try:
        from PIL import Image
        with Image.open(path) as im:  # header only
            return im.width, im.height
    except Exception:
        return None


# =============================================================================================
# stage 1
# =============================================================================================
class SyntheticDetector:
    """Stands in for DocLayout-YOLO (weights unavailable offline): deterministic per-tile detections
    from multimodal_embeddings_b200.synth, keyed by page name.  `--detections synthetic` only."""

    def __init__(self, boxes_per_page: int = 2000, seed: int = 0xB200):
        self.boxes_per_page, self.seed = boxes_per_page, seed

    def detect_page(self, base: str, width: int, height: int, rows: int, cols: int, overlap: float, tiles, **_):
        from . import synth
        n = max(1, self.boxes_per_page)
        seed = (self.seed + sum(ord(c) * (i + 1) for i, c in enumerate(base)) + 97 * rows + cols) & 0x7FFFFFFF
        d = synth.page_detections(width, height, rows, cols, overlap, n, seed)
        out = []
        for ci in range(rows * cols):
            m = d["box_cell"] == ci
            out.append({"boxes": d["boxes_local"][m].tolist(), "classes": d["classes"][m].tolist(),
                        "scores": d["scores"][m].tolist(), "class_names": synth.class_names_of(d["classes"][m])})
        return out


class ReplayDetector:
    """Replays detections from a folder of stage-1 JSONs (``<folder>/json/<base>.json`` and
    ``<folder>/json/<base>_grid_RxC.json``), e.g. a previous run of the reference."""

    def __init__(self, folder: str):
        self.folder = folder

    def detect_page(self, base, width, height, rows, cols, overlap, tiles, **_):
        name = f"{base}.json" if (rows, cols) == (1, 1) else f"{base}_grid_{rows}x{cols}.json"
        with open(os.path.join(self.folder, "json", name)) as f:
            d = json.load(f)
        if "cells" in d:
            return [{k: c["regions"][k] for k in ("boxes", "classes", "scores", "class_names")} for c in d["cells"]]
        return [{k: d[k] for k in ("boxes", "classes", "scores", "class_names")}]


class DocLayoutYoloDetector:
    """The reference's network (1_doclayout_bboxes.py:178-179) fed with this path's letterboxed tiles instead of
    image files.  Needs the third-party `doclayout_yolo` package and a local weights file (`--model_path`);
    neither exists offline, so this class is exercised only where they are installed — and refuses loudly
    otherwise.  The tiles are already what LetterBox -> /255 produces, so the network is called on the tensors
    and its boxes are mapped back to tile pixels by undoing the letterbox (ultralytics' scale_boxes:
    subtract the pad, divide by the gain, clip)."""

    def __init__(self, model_path: str, conf: float, device: Optional[str]):
        if not model_path or not os.path.exists(model_path):
            raise FileNotFoundError(f"Model file not found at {model_path}")  # 1:119-121
        try:
            from doclayout_yolo import YOLOv10
        except ImportError as e:
            raise RuntimeError("--model_path needs the doclayout_yolo package, which is not installed here; "
                               "use --detector module:factory or --detections replay/synthetic") from e
        self.model, self.conf, self.device = YOLOv10(model_path), conf, device or "cuda"

    def detect_page(self, base, width, height, rows, cols, overlap, tiles, tile_infos=None, **_):
        out = []
        for t, info in zip(tiles, tile_infos):
            res = self.model.predict(t.unsqueeze(0).float(), imgsz=max(t.shape[1:]), conf=self.conf, device=self.device)[0]
            b = res.boxes.xyxy.float().cpu().numpy().copy()
            gain = min(info["new_w"] / (info["x1"] - info["x0"]), info["new_h"] / (info["y1"] - info["y0"]))
            b[:, [0, 2]] = np.clip((b[:, [0, 2]] - info["pad_l"]) / gain, 0, info["x1"] - info["x0"])
            b[:, [1, 3]] = np.clip((b[:, [1, 3]] - info["pad_t"]) / gain, 0, info["y1"] - info["y0"])
            out.append({"boxes": b, "classes": res.boxes.cls.cpu().numpy(), "scores": res.boxes.conf.cpu().numpy()})
        return out


def load_detector_plugin(spec: str, args):
    """`--detector module:factory`: factory(args) returns an object with
    detect_page(base, width, height, rows, cols, overlap, tiles, tile_names=, tile_sizes=, tile_infos=) ->
    one dict per tile {boxes [n,4] xyxy in tile pixels, classes [n], scores [n]} BEFORE the per-tile NMS
    (the command line applies 1:217-225 itself); `tiles` are the letterboxed fp16 CHW tensors on the GPU."""
    import importlib
    mod, _, attr = spec.partition(":")
    return getattr(importlib.import_module(mod), attr or "Detector")(args)


def get_image_paths(input_folder):
    """1_doclayout_bboxes.py:345-364."""
    paths = []
    for root, _, files in os.walk(input_folder):
        for file in files:
            if os.path.splitext(file)[1].lower() in IMAGE_EXTENSIONS:
                paths.append(os.path.join(root, file))
    return sorted(paths)


def _take(values, keep):
    if isinstance(values, np.ndarray):
        return values[keep].tolist()
    return [values[i] for i in keep]


def main_stage1(argv: Optional[Sequence[str]] = None) -> int:
    import torch
    from . import ops, synth
    from .reference_api import nms_per_tile, parse_grid_configs, translate_coordinates_to_original
    logger = _logger("DocLayoutAnalyzer")
    p = argparse.ArgumentParser(description="Document Layout Analysis")
    p.add_argument("--input_folder", required=True)
    p.add_argument("--output_folder", required=True)
    p.add_argument("--conf_threshold", type=float, default=0.1)
    p.add_argument("--iou_threshold", type=float, default=0.45)
    p.add_argument("--device", choices=["cpu", "cuda"])
    p.add_argument("--model_path")
    p.add_argument("--skip_errors", action="store_true")
    p.add_argument("--rows", type=int, default=2)
    p.add_argument("--cols", type=int, default=2)
    p.add_argument("--grids", type=str, default="2x2,3x3,4x4")
    p.add_argument("--overlap", type=float, default=20.0)
    p.add_argument("--disable_grid", action="store_true")
    p.add_argument("--detector", help="extension: detector plug-in, module:factory (load_detector_plugin)")
    p.add_argument("--detections", choices=["synthetic", "replay"],
                   help="extension: synthetic detections (benchmarks/tests) or replay of a stage-1 JSON folder")
    p.add_argument("--replay_folder")
    p.add_argument("--boxes_per_page", type=int, default=2000)
    p.add_argument("--imgsz", type=int, default=1024)
    p.add_argument("--write_tiles", action="store_true", help="also write the tile images of 1:568 (grid_RxC/images)")
    p.add_argument("--batch_pages", type=int, default=16, help="pages tiled per launch")
    p.add_argument("--host_decode", action="store_true", help="decode every scan with cv2 on the host, JPEG included")
    args = p.parse_args(argv)
    if args.device == "cpu":
        logger.error("--device cpu: this implementation has no CPU path (libpagegeom.so is CUDA only)")
        return 2
    # The detector.  The reference downloads DocLayout-YOLO weights here (1:125-129); there is no silent
    # substitute: without a plug-in, a weights file or an explicit --detections choice the run is refused.
    try:
        if args.detector:
            detector = load_detector_plugin(args.detector, args)
        elif args.detections == "replay":
            if not args.replay_folder:
                raise RuntimeError("--detections replay needs --replay_folder")
            detector = ReplayDetector(args.replay_folder)
        elif args.detections == "synthetic":
            logger.warning("--detections synthetic: boxes, classes and scores in the output are FABRICATED "
                           "(multimodal_embeddings_b200.synth), not detections of the input images")
            detector = SyntheticDetector(args.boxes_per_page)
        elif args.model_path:
            detector = DocLayoutYoloDetector(args.model_path, args.conf_threshold, args.device)
        else:
            raise RuntimeError("no detector: DocLayout-YOLO weights cannot be downloaded here. Give --model_path "
                               "(with the doclayout_yolo package installed), --detector module:factory, "
                               "--detections replay --replay_folder DIR, or --detections synthetic")
    except Exception as e:
        logger.error(f"Error loading model: {str(e)}")
        return 2
    json_folder = os.path.join(args.output_folder, "json")
    os.makedirs(json_folder, exist_ok=True)
    os.makedirs(os.path.join(args.output_folder, "visualizations"), exist_ok=True)
    grid_configs: List[Tuple[int, int]] = []
    if not args.disable_grid:  # 1:712-725
        if args.grids:
            grid_configs = parse_grid_configs(args.grids) or [(args.rows, args.cols)]
        else:
            grid_configs = [(args.rows, args.cols)]
    image_paths = get_image_paths(args.input_folder)
    if not image_paths:
        logger.error(f"No images found in {args.input_folder}")
        return 0
    image_paths = _my_share(image_paths)
    grids = [(1, 1)] + grid_configs
    params = {"conf_threshold": args.conf_threshold, "iou_threshold": args.iou_threshold}
    processed = errors = 0
    stop = False
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1))
    for b0 in range(0, len(image_paths), max(1, args.batch_pages)):
        if stop:
            break
        chunk = image_paths[b0:b0 + max(1, args.batch_pages)]
        # Baseline JPEG scans cross PCIe as files and are decoded on the device (pg_jpeg_decode: a greyscale file gives
        # one grey plane, tiled by the one-channel plans; a colour file gives BGR like cv2).  Everything else is
        # decoded on host threads like the reference (cv2.imread, 1:381; cv2 releases the GIL) and uploaded as BGR
        # through pinned staging.  Either way ONE tiler launch per kind covers every tile of every grid of every page
        # of the chunk (PgTileBatch).
        jpeg_bytes = {}
        if not args.host_decode:
            for path in chunk:
                if os.path.splitext(path)[1].lower() in (".jpg", ".jpeg"):
                    try:
                        with open(path, "rb") as f:
                            data = f.read()
                        if ops.jpeg_probe(data) is not None:
                            jpeg_bytes[path] = data
                    except OSError:
                        pass
        host_paths = [p_ for p_ in chunk if p_ not in jpeg_bytes]
        decoded = dict(zip(host_paths, pool.map(ops.decode_page, host_paths)))
        for path in host_paths:
            if decoded[path] is None:  # detect_regions -> None -> process_image False (1:240-242, 466-482)
                logger.error(f"Error detecting regions in {os.path.basename(path)}: cannot decode image")
                errors += 1
        live = []  # (path, batch, index in batch, width, height, host page or None)
        try:
            # every page of the chunk as (path, device page [H, pitch], width, height, channels, host pixels or None)
            staged = []
            on_device = [p_ for p_ in chunk if p_ in jpeg_bytes]
            if on_device:
                jdec = ops.JpegDecoder()
                blob, off = ops.pack_files([jpeg_bytes[p_] for p_ in on_device])
                sizes = jdec.set_files(blob, off)
                pages_dev = jdec.decode(blob.to("cuda", non_blocking=True))
                torch.cuda.current_stream().synchronize()
                jdec.check()
                staged += [(p_, pg, sz[0], sz[1], sz[2], None) for p_, pg, sz in zip(on_device, pages_dev, sizes)]
            from_host = [p_ for p_ in host_paths if decoded[p_] is not None]
            if from_host:
                uploaded = ops.upload_pages_pinned([decoded[p_] for p_ in from_host])
                staged += [(p_, pg, decoded[p_].shape[1], decoded[p_].shape[0], 1 if decoded[p_].ndim == 2 else 3, decoded[p_])
                           for p_, pg in zip(from_host, uploaded)]
            for chn in (1, 3):  # grey planes -> one-channel plans, BGR pages -> three-channel plans
                group = [t_ for t_ in staged if t_[4] == chn]
                if not group:
                    continue
                tb = ops.TileBatch([(t_[2], t_[3]) for t_ in group], grids, args.overlap, args.imgsz, channels=chn)
                tb.bind([t_[1] for t_ in group])
                tb.run()
                for i, (p_, pg, w_, h_, _, host_page) in enumerate(group):
                    if args.write_tiles:  # the tile files need the pixels on the host, as cv2.imread gives them
                        if host_page is None:
                            host_page = pg[:, :chn * w_].cpu().numpy().reshape(h_, w_, chn) if chn == 3 else pg[:, :w_].cpu().numpy()
                        if host_page.ndim == 2:
                            host_page = np.repeat(host_page[..., None], 3, -1)
                    live.append((p_, tb, i, w_, h_, host_page))
        except Exception as e:
            errors += len(chunk)
            logger.error(f"Error processing {os.path.basename(chunk[0])}: {str(e)}")
            if not args.skip_errors:
                logger.error("Stopping due to error. Use --skip_errors to continue despite errors.")
                break
            continue
        live.sort(key=lambda t: t[0])  # the reference's order: sorted paths (1:364)
        for image_path, batch, pi, w, h, page in live:
            try:
                plan = batch.plan_of(pi)
                base, ext = os.path.splitext(os.path.basename(image_path))
                t0 = 0
                for rows, cols in grids:
                    n_t = rows * cols
                    infos = plan.tiles[t0:t0 + n_t]
                    full_page = (rows, cols) == (1, 1) and t0 == 0
                    names = [os.path.basename(image_path)] if full_page else \
                        [f"{base}_row{ti['row']}_col{ti['col']}{ext}" for ti in infos]
                    sizes = [(ti["x1"] - ti["x0"], ti["y1"] - ti["y0"]) for ti in infos]
                    views = [batch.tile_view(pi, t0 + k) for k in range(n_t)]
                    dets = detector.detect_page(base, w, h, rows, cols, args.overlap, views, tile_names=names,
                                                tile_sizes=sizes, tile_infos=infos)
                    # per-tile class-agnostic NMS of detect_regions (1:217-225), all tiles of the grid in one launch
                    keeps = nms_per_tile([d["boxes"] for d in dets], [d["scores"] for d in dets], args.iou_threshold)
                    kept = []
                    for d, keep in zip(dets, keeps):
                        keep = keep.tolist()
                        classes = _take(d["classes"], keep)
                        cn = _take(d["class_names"], keep) if "class_names" in d else \
                            [synth.ID_TO_NAMES[int(c)] for c in classes]  # 1:233
                        kept.append({"boxes": _take(d["boxes"], keep), "classes": classes,
                                     "scores": _take(d["scores"], keep), "class_names": cn})
                    if full_page:  # full-page pass, 1:446-482 / schema 1:227-235
                        d = kept[0]
                        _dump({"image_path": image_path, "image_size": {"width": w, "height": h}, "parameters": params,
                               "boxes": d["boxes"], "classes": d["classes"], "scores": d["scores"],
                               "class_names": d["class_names"]}, os.path.join(json_folder, f"{base}.json"))
                    else:  # 1:513-654
                        grid_folder = os.path.join(args.output_folder, f"grid_{rows}x{cols}")
                        for sub in ("images", "json", "visualizations", "visualizations_original_coords"):
                            os.makedirs(os.path.join(grid_folder, sub), exist_ok=True)
                        info = {"original_image_path": image_path,
                                "grid_config": {"rows": rows, "cols": cols, "overlap_percentage": args.overlap}, "cells": []}
                        for k, d in enumerate(kept):
                            ti = infos[k]
                            cc = plan.cell_coordinates(t0 + k)
                            cell_path = os.path.join(grid_folder, "images", names[k])
                            cell_json = os.path.join(grid_folder, "json", names[k].replace(ext, ".json"))
                            if args.write_tiles:  # the slice of 1:424-430, written as 1:568 does
                                import cv2
                                cv2.imwrite(cell_path, page[ti["y0"]:ti["y1"], ti["x0"]:ti["x1"]])
                            orig = translate_coordinates_to_original(d["boxes"], cc)
                            cell_regions = {"image_path": cell_path,
                                            "image_size": {"width": sizes[k][0], "height": sizes[k][1]},
                                            "parameters": params, "boxes": d["boxes"], "classes": d["classes"],
                                            "scores": d["scores"], "class_names": d["class_names"],
                                            "cell_coordinates": cc, "original_image_path": image_path,
                                            "boxes_original": orig,
                                            "grid_info": {"rows": rows, "cols": cols, "row": ti["row"], "col": ti["col"]}}
                            _dump(cell_regions, cell_json)
                            info["cells"].append({"cell_path": cell_path, "cell_json_path": cell_json, "cell_coordinates": cc,
                                                  "row": ti["row"], "col": ti["col"],
                                                  "regions": {"boxes": d["boxes"], "boxes_original": orig,
                                                              "classes": d["classes"], "scores": d["scores"],
                                                              "class_names": d["class_names"]}})
                        if info["cells"]:
                            _dump(info, os.path.join(json_folder, f"{base}_grid_{rows}x{cols}.json"))
                    t0 += n_t
                processed += 1
            except Exception as e:  # 1:778-783
                errors += 1
                logger.error(f"Error processing {os.path.basename(image_path)}: {str(e)}")
                if not args.skip_errors:
                    logger.error("Stopping due to error. Use --skip_errors to continue despite errors.")
                    stop = True
                    break
    pool.shutdown()
    logger.info(f"Processing complete. Successfully processed {processed} images with {errors} errors. "
                f"Results saved to {args.output_folder}")
    return 0


# =============================================================================================
# stage 2
# =============================================================================================
def _grid_page_size(grid_info, no_image_check: bool):
    path = grid_info.get("image_path") if ("image_path" in grid_info and os.path.exists(grid_info["image_path"])) \
        else grid_info.get("original_image_path")
    if path and os.path.exists(path):
        return _image_size(path)
    if no_image_check and grid_info.get("cells"):
        w = max(c["cell_coordinates"]["x_end"] for c in grid_info["cells"])
        h = max(c["cell_coordinates"]["y_end"] for c in grid_info["cells"])
        return int(w), int(h)
    return None


def main_stage2(argv: Optional[Sequence[str]] = None) -> int:
    from . import reference_api as api
    logger = _logger("EdgeBoxFilter")
    p = argparse.ArgumentParser(description="Filter bounding boxes that touch internal grid edges")
    p.add_argument("--input_folder", required=True)
    p.add_argument("--output_folder", required=True)
    p.add_argument("--edge_threshold", type=int, default=10)
    p.add_argument("--viz_alpha", type=float, default=0.3)
    p.add_argument("--skip_errors", action="store_true")
    p.add_argument("--process_grids", action="store_true")
    p.add_argument("--no_image_check", action="store_true")
    args = p.parse_args(argv)
    out_json = os.path.join(args.output_folder, "json")
    os.makedirs(out_json, exist_ok=True)
    os.makedirs(os.path.join(args.output_folder, "visualizations"), exist_ok=True)

    def run_folder(in_folder, out_folder):
        paths = _my_share(sorted(os.path.join(r, f) for r, _, fs in os.walk(in_folder) for f in fs if f.endswith(".json")))
        ok = err = 0
        # Grid documents laid out as stage 1 writes them go through three device calls for the whole folder —
        # numbers text -> f64, edge filter with every cell as its own page, filtered documents -> text
        # (records.filter_grid_files); the results are written in the loop below, in the reference's file order,
        # so that an error in another file stops the run at the same place as in the reference.
        try:
            from .records import filter_grid_files
            def standard_ok(doc):  # the image check of 2:381-412 for full-page documents
                if args.no_image_check:
                    return True
                return any(doc.get(k) and os.path.exists(doc[k]) for k in ("image_path", "original_image_path"))

            fast = filter_grid_files(paths, args.edge_threshold, lambda sk: _grid_page_size(sk, args.no_image_check),
                                     api._cell_tuple, standard_ok)
        except Exception as e:
            logger.error(f"Batch path failed ({e}); processing the files one by one")
            fast = {}
        for path in paths:
            try:
                if path in fast:
                    with open(os.path.join(out_folder, os.path.basename(path)), "wb") as f:
                        f.write(fast[path])
                    ok += 1
                    continue
                with open(path) as f:
                    regions = json.load(f)
                target = os.path.join(out_folder, os.path.basename(path))
                if "cells" in regions and ("grid_config" in regions or "grid_info" in regions):  # 2:374
                    size = _grid_page_size(regions, args.no_image_check)
                    filtered = api.filter_grid_info(regions, args.edge_threshold, image_size=size) if size else None
                    _dump(filtered if filtered else regions, target)  # 2:485-487 / 2:567-573
                    ok += 1
                    continue
                image_path = regions.get("image_path")
                if not args.no_image_check and not (image_path and os.path.exists(image_path)):  # 2:381-412
                    alt = regions.get("original_image_path")
                    if not (alt and os.path.exists(alt)):
                        logger.error(f"Image path not found: {image_path}")
                        err += 1
                        continue
                _dump(api.filter_edge_boxes(regions, args.edge_threshold), target)
                ok += 1
            except Exception as e:
                err += 1
                logger.error(f"Error processing {os.path.basename(path)}: {str(e)}")
                if not args.skip_errors:
                    logger.error("Stopping due to error. Use --skip_errors to continue despite errors.")
                    break
        return ok, err

    json_folder = os.path.join(args.input_folder, "json")
    if os.path.exists(json_folder):
        ok, err = run_folder(json_folder, out_json)
        logger.info(f"Main JSON processing complete. Successfully processed {ok} JSON files with {err} errors")
    else:
        logger.warning(f"Main JSON folder not found: {json_folder}")
    if args.process_grids:  # 2:579-649
        for item in sorted(os.listdir(args.input_folder)):
            src = os.path.join(args.input_folder, item, "json")
            if item.startswith("grid_") and os.path.isdir(src):
                dst = os.path.join(args.output_folder, item, "json")
                os.makedirs(dst, exist_ok=True)
                os.makedirs(os.path.join(args.output_folder, item, "visualizations"), exist_ok=True)
                run_folder(src, dst)
    logger.info(f"All processing complete. Results saved to {args.output_folder}")
    return 0


# =============================================================================================
# stage 3
# =============================================================================================
def find_grid_jsons(input_folder) -> Dict[str, List[str]]:
    """3_combine_grids.py:140-198.  Same grouping rules; globs are sorted here (the reference's
    order is whatever the filesystem returns) with the standard JSON still first (:175)."""
    groups: Dict[str, List[str]] = {}
    json_folder = os.path.join(input_folder, "json")
    if os.path.exists(json_folder):
        for g in sorted(glob.glob(os.path.join(json_folder, "*_grid_*.json"))):
            groups.setdefault(os.path.basename(g).split("_grid_")[0], []).append(g)
        for j in sorted(glob.glob(os.path.join(json_folder, "*.json"))):
            if "_grid_" not in j and "_combined" not in j:
                groups.setdefault(os.path.splitext(os.path.basename(j))[0], []).insert(0, j)
    for sub in sorted(os.listdir(input_folder)):
        sj = os.path.join(input_folder, sub, "json")
        if sub.startswith("grid_") and os.path.isdir(os.path.join(input_folder, sub)) and os.path.exists(sj):
            for g in sorted(glob.glob(os.path.join(sj, "*.json"))):
                fn = os.path.basename(g)
                base = fn.split("_row")[0] if ("_row" in fn and "_col" in fn) else os.path.splitext(fn)[0]
                groups.setdefault(base, []).append(g)
    return groups


def pool_documents(json_paths, logger):
    """Concatenation of combine_boxes_for_image (3_combine_grids.py:222-270)."""
    boxes, scores, classes, names = [], [], [], []
    image_path = image_size = None
    for path in json_paths:
        try:
            with open(path) as f:
                d = json.load(f)
            if "cells" in d:
                if not image_path and "original_image_path" in d:
                    image_path = d["original_image_path"]
                for cell in d["cells"]:
                    if "regions" in cell and "boxes_original" in cell["regions"]:
                        r = cell["regions"]
                        boxes.extend(r["boxes_original"]); scores.extend(r["scores"])
                        classes.extend(r["classes"]); names.extend(r["class_names"])
            elif "boxes" in d:
                if not image_path and "image_path" in d:
                    image_path = d["image_path"]
                if not image_size and "image_size" in d:
                    image_size = d["image_size"]
                boxes.extend(d["boxes_original"] if "boxes_original" in d else d["boxes"])
                scores.extend(d["scores"]); classes.extend(d["classes"]); names.extend(d["class_names"])
        except Exception as e:
            logger.error(f"Error reading {path}: {str(e)}")
    return boxes, scores, classes, names, image_path, image_size


def pool_documents_fast(json_paths, logger, fast: dict):
    """pool_documents for a page whose files were all read by records.load_pool_inputs (numbers converted on
    the device): numpy arrays instead of lists, same order, plus `numbers_are_floats`.  A page with any file the
    device reader declined is pooled by pool_documents (CPython's json.load), whole."""
    if not all(p in fast for p in json_paths):
        b, s, c, n, image_path, image_size = pool_documents(json_paths, logger)
        floats = all(type(v) is float for row in b for v in row) and all(type(v) is float for v in s) and \
            all(type(v) is float for v in c)
        return b, s, c, n, image_path, image_size, floats
    boxes, scores, classes, names = [], [], [], []
    image_path = image_size = None
    for path in json_paths:
        d = fast[path]
        if not image_path and d["image_path"]:
            image_path = d["image_path"]
        if not d["grid"] and not image_size and d["image_size"]:
            image_size = d["image_size"]
        boxes.append(d["boxes"]); scores.append(d["scores"]); classes.append(d["classes"])
        names.extend(d["class_names"])
    return (np.concatenate(boxes).reshape(-1, 4), np.concatenate(scores), np.concatenate(classes), names, image_path,
            image_size, True)


def main_stage3(argv: Optional[Sequence[str]] = None) -> int:
    from . import ops
    logger = _logger("GridBoxCombiner")
    p = argparse.ArgumentParser(description="Combine bounding boxes from different grid patterns")
    p.add_argument("--input_folder", required=True)
    p.add_argument("--output_folder", required=True)
    p.add_argument("--iou_threshold", type=float, default=0.5)
    p.add_argument("--viz_alpha", type=float, default=0.3)
    p.add_argument("--sidecar", action="store_true",
                   help="extension: also write <base>_combined.pgrec (raw arrays) for this implementation's stages 4/5")
    args = p.parse_args(argv)
    out_json = os.path.join(args.output_folder, "json")
    os.makedirs(out_json, exist_ok=True)
    os.makedirs(os.path.join(args.output_folder, "visualizations"), exist_ok=True)
    groups = find_grid_jsons(args.input_folder)
    if not groups:
        logger.error(f"No JSON files found in {args.input_folder}")
        return 0
    items = _my_share(list(groups.items()))
    # the stage-1/2 documents of every page of the shard: numbers converted on the GPU in one call
    # (records.load_pool_inputs); files it declines are read by json.load as in the reference
    try:
        from .records import load_pool_inputs
        fast = load_pool_inputs([p for _, paths in items for p in paths])
    except Exception as e:
        logger.error(f"Batch input reader failed ({e}); reading the files one by one")
        fast = {}
    pooled = []
    for base, paths in items:
        b, s, c, n, image_path, image_size, floats = pool_documents_fast(paths, logger, fast)
        if len(b) == 0:
            logger.warning(f"No boxes found for {base}")
            continue
        pooled.append((base, paths, b, s, c, n, image_path, image_size, floats))
    if pooled:  # one batched merge for every page
        off = np.cumsum([0] + [len(x[2]) for x in pooled])
        boxes = np.concatenate([np.asarray(x[2], np.float64).reshape(-1, 4) for x in pooled])
        scores = np.concatenate([np.asarray(x[3], np.float64) for x in pooled])
        classes = np.concatenate([np.asarray(x[4], np.float64) for x in pooled])
        ws = ops.NmsWorkspace(len(boxes), len(pooled), 64 if args.iou_threshold >= 0 else int(max(np.diff(off)) // 32 + 2))
        # status checked inside (a workspace overflow is retried with the dense bound, anything else raises)
        kept, n_kept, ws = ops.nms_merge(boxes, scores, classes, off, args.iou_threshold, workspace=ws,
                                         max_boxes_per_page=int(max(np.diff(off))), check_status=True)
        kept, n_kept = kept.cpu().numpy(), n_kept.cpu().numpy()
        # Records: laid out on the device (pg_json_combined), byte-identical to json.dump(indent=2).  Inputs that
        # are not what stage 1/2 write (integer literals among the numbers) keep Python's own encoder, which
        # would print them without ".0".
        all_float = all(x[8] for x in pooled)
        if all_float and os.environ.get("PG_PYTHON_JSON") != "1":
            table = {}
            name_id = np.asarray([table.setdefault(nm, len(table)) for x in pooled for nm in x[5]], np.int32)
            ht = [ops.combined_head_tail(x[6], x[7], args.iou_threshold, x[1]) for x in pooled]
            docs = ops.json_combined(boxes, classes, scores, name_id, off, [h for h, _ in ht], [t for _, t in ht],
                                     [json.dumps(nm).encode("ascii") for nm in table], kept_idx=kept, n_kept=n_kept)
            for i, (x, doc) in enumerate(zip(pooled, docs)):
                json_path = os.path.join(out_json, f"{x[0]}_combined.json")
                with open(json_path, "wb") as f:
                    f.write(doc)
                if args.sidecar:  # raw arrays for this implementation's stages 4/5 (records.py)
                    idx = kept[off[i]: off[i] + n_kept[i]]
                    write_sidecar(json_path, x[6], x[7], {"iou_threshold": args.iou_threshold}, x[1], boxes[idx],
                                  classes[idx], scores[idx], list(table), name_id[idx])
            pooled = []
        for i, (base, paths, b, s, c, n, image_path, image_size, _) in enumerate(pooled):
            b, s, c = (v.tolist() if isinstance(v, np.ndarray) else v for v in (b, s, c))
            idx = (kept[off[i]: off[i] + n_kept[i]] - off[i]).tolist()
            _dump({"image_path": image_path, "image_size": image_size,  # key order of 3:282-291
                   "parameters": {"iou_threshold": args.iou_threshold},
                   "boxes": [b[j] for j in idx], "classes": [c[j] for j in idx], "scores": [s[j] for j in idx],
                   "class_names": [n[j] for j in idx], "source_jsons": paths},
                  os.path.join(out_json, f"{base}_combined.json"))
    logger.info(f"Processing complete. Combined results saved to {args.output_folder}")
    return 0


# =============================================================================================
# stage 4
# =============================================================================================
def main_stage4(argv: Optional[Sequence[str]] = None) -> int:
    from . import ops
    from .reference_api import _flags_from_names
    logger = _logger("MedianWidthExtractor")
    p = argparse.ArgumentParser(description="Extract median width of plain_text boxes")
    p.add_argument("--input_folder", required=True)
    p.add_argument("--output_folder", required=True)
    p.add_argument("--min_margin_percent", type=float, default=0.2)
    p.add_argument("--no_image_check", action="store_true")
    args = p.parse_args(argv)
    out_json = os.path.join(args.output_folder, "json")
    os.makedirs(out_json, exist_ok=True)
    os.makedirs(os.path.join(args.output_folder, "visualizations"), exist_ok=True)
    json_folder = args.input_folder  # 4:245-247
    if not os.path.isdir(json_folder):
        json_folder = os.path.join(args.input_folder, "json")
    if not os.path.exists(json_folder):
        logger.error(f"JSON folder not found: {json_folder}")
        return 0
    files = sorted(glob.glob(os.path.join(json_folder, "*.json")))
    if not files:
        logger.error(f"No JSON files found in {json_folder}")
        return 0
    files = _my_share(files)
    # the records of the whole batch: the .pgrec sidecar where stage 3 wrote one, else the JSON text with its
    # number arrays converted on the GPU in one call (records.load_records), else json.load
    try:
        recs = load_records(files)
    except Exception as e:
        logger.error(f"Batch record reader failed ({e}); reading the files one by one")
        recs = {}
    pages = []
    for path in files:
        try:  # 4:103-151
            d = recs[path] if path in recs else load_records([path])[path]
            size = d.get("image_size", {})  # null (grid-only pooling, 3:249-250) raises here like 4:122-124: no file
            names, boxes = d.get("class_names", []), d.get("boxes", [])
            n = min(len(names), len(boxes))
            pages.append((path, d.get("image_path", ""), size.get("width", 0), size.get("height", 0), boxes[:n], names[:n]))
        except Exception as e:
            logger.error(f"Error processing {path}: {str(e)}")
    live = [pg for pg in pages if len(pg[4])]
    med = {}
    if live:
        off = np.cumsum([0] + [len(pg[4]) for pg in live])
        boxes = np.concatenate([np.asarray(pg[4], np.float64).reshape(-1, 4) for pg in live])
        flags = np.concatenate([_flags_from_names(pg[5]) for pg in live])
        m, nb = ops.width_median(boxes, flags, off, [[pg[2], pg[3]] for pg in live], args.min_margin_percent)
        m, nb = m.cpu().numpy(), nb.cpu().numpy()
        med = {pg[0]: (np.float64(m[i]) if nb[i] else 0) for i, pg in enumerate(live)}
    for path, image_path, w, h, _, _ in pages:
        median_width = med.get(path, 0)
        if image_path and (args.no_image_check or os.path.exists(image_path)):  # 4:272
            base = os.path.splitext(os.path.basename(path))[0]
            _dump({"image_path": image_path, "median_width": median_width, "page_width": w, "page_height": h,
                   "width_ratio": median_width / w if w > 0 else 0},
                  os.path.join(out_json, f"{base}_median_width.json"))
    logger.info(f"Processing complete. Individual results saved to {out_json}")
    return 0


# =============================================================================================
# stage 5
# =============================================================================================
def find_matching_median_json(layout_json_path, median_json_folder):
    """5_detect_column_centers.py:480-539 (same fallback ladder)."""
    base = os.path.splitext(os.path.basename(layout_json_path))[0]
    exact = os.path.join(median_json_folder, f"{base}_median_width.json")
    if os.path.exists(exact):
        return exact
    listing = sorted(os.listdir(median_json_folder)) if os.path.isdir(median_json_folder) else []
    if "_grid_" in base:
        prefix = base.split("_grid_")[0]
        cand = os.path.join(median_json_folder, f"{prefix}_median_width.json")
        if os.path.exists(cand):
            return cand
        for fn in listing:
            if fn.endswith("_median_width.json") and fn.startswith(f"{prefix}_"):
                return os.path.join(median_json_folder, fn)
    for part in base.split("_"):
        if part.lower().startswith("page") or (len(part) >= 4 and part.isdigit()):
            for fn in listing:
                if part in fn and fn.endswith("_median_width.json"):
                    return os.path.join(median_json_folder, fn)
    mt = re.search(r"(page[_-]?\d+)", base, re.IGNORECASE)
    if mt:
        for fn in listing:
            if mt.group(1) in fn and fn.endswith("_median_width.json"):
                return os.path.join(median_json_folder, fn)
    med_files = [f for f in listing if f.endswith("_median_width.json")]
    if len(med_files) == 1:
        return os.path.join(median_json_folder, med_files[0])
    return None


def main_stage5(argv: Optional[Sequence[str]] = None) -> int:
    from . import ops
    from .reference_api import _flags_from_names
    logger = _logger("ColumnCenterDetector")
    p = argparse.ArgumentParser(description="Detect column centers in document pages")
    p.add_argument("--input_folder", required=True)
    p.add_argument("--median_folder", required=True)
    p.add_argument("--output_folder", required=True)
    p.add_argument("--min_confidence", type=float, default=0.3)
    p.add_argument("--verbose", action="store_true")
    p.add_argument("--no_image_check", action="store_true")
    args = p.parse_args(argv)
    os.makedirs(args.output_folder, exist_ok=True)
    files = sorted(glob.glob(os.path.join(args.input_folder, "*.json")))
    if not files:
        logger.error(f"No JSON files found in {args.input_folder}")
        return 0
    files = _my_share(files)
    try:
        recs = load_records(files)  # sidecar / device-parsed text / json.load, per file (records.load_records)
    except Exception as e:
        logger.error(f"Batch record reader failed ({e}); reading the files one by one")
        recs = {}
    jobs, failures = [], 0
    for path in files:
        mpath = find_matching_median_json(path, args.median_folder)
        if not mpath:
            logger.warning(f"No matching median width JSON found for {os.path.basename(path)}")
            failures += 1
            continue
        try:  # 5:337-400
            layout = recs[path] if path in recs else load_records([path])[path]
            with open(mpath) as f:
                median_width = json.load(f).get("median_width", 0)
            if median_width <= 0:
                logger.error(f"Invalid median width: {median_width}")
                failures += 1
                continue
            image_path = layout.get("image_path", "")
            size = layout.get("image_size", {})
            if isinstance(size, dict):
                w, h = size.get("width", 0), size.get("height", 0)
            elif isinstance(size, list) and len(size) >= 2:
                w, h = size[0], size[1]
            else:
                w = h = 0
            if not image_path or not (args.no_image_check or os.path.exists(image_path)) or w <= 0 or h <= 0:
                logger.error(f"Invalid image information: {image_path}, {w}x{h}")
                failures += 1
                continue
            boxes, names = layout.get("boxes", []), layout.get("class_names", [])
            scores = layout.get("scores", [1.0] * len(boxes))
            n = min(len(boxes), len(names), len(scores))  # zip() semantics of 5:110
            jobs.append((path, image_path, w, h, median_width, boxes[:n], names[:n], scores[:n]))
        except Exception as e:
            logger.error(f"Error processing {os.path.basename(path)}: {str(e)}")
            failures += 1
    ok = 0
    live = [j for j in jobs if len(j[5])]
    results = {}
    if live:
        off = np.cumsum([0] + [len(j[5]) for j in live])
        boxes = np.concatenate([np.asarray(j[5], np.float64).reshape(-1, 4) for j in live])
        flags = np.concatenate([_flags_from_names(j[6]) for j in live])
        scores = np.concatenate([np.asarray(j[7], np.float64) for j in live])
        centers, widths, n_cols = ops.column_peaks(boxes, flags, scores, off, [[j[2], j[3]] for j in live],
                                                   [float(j[4]) for j in live], args.min_confidence, max_cols=256)
        centers, widths, n_cols = centers.cpu().numpy(), widths.cpu().numpy(), n_cols.cpu().numpy()
        for i, j in enumerate(live):
            if n_cols[i] < 0 or n_cols[i] > 256:
                logger.error(f"Error processing {os.path.basename(j[0])}: page outside kernel limits")
                continue
            results[j[0]] = ([float(x) for x in centers[i, : n_cols[i]]], [float(x) for x in widths[i, : n_cols[i]]])
    for path, image_path, w, h, median_width, *_ in jobs:
        cc, cw = results.get(path, ([], []))
        if not cc:
            logger.warning(f"No column centers found for {os.path.basename(path)}")
            failures += 1
            continue
        out_json = os.path.join(args.output_folder, "json")
        os.makedirs(out_json, exist_ok=True)
        os.makedirs(os.path.join(args.output_folder, "visualizations"), exist_ok=True)
        base = os.path.splitext(os.path.basename(path))[0]
        _dump({"image_path": image_path, "page_width": w, "page_height": h, "median_width": median_width,
               "column_centers": cc, "column_widths": cw, "num_columns": len(cc)},  # 5:426-434
              os.path.join(out_json, f"{base}_columns.json"))
        ok += 1
    logger.info(f"Processing complete. Successfully processed {ok} pages, {failures} failures.")
    return 0


MAINS = {"1": main_stage1, "2": main_stage2, "3": main_stage3, "4": main_stage4, "5": main_stage5}

if __name__ == "__main__":
    if len(sys.argv) < 2 or sys.argv[1] not in MAINS:
        sys.exit("usage: python -m multimodal_embeddings_b200.cli {1,2,3,4,5} [stage args]")
    sys.exit(MAINS[sys.argv[1]](sys.argv[2:]))
