"""Whole hot path for a shard of pages, chained on the device:

    tile+letterbox (K1) -> translate+edge filter (K2) -> NMS merge (K3) -> class flags
    -> plain_text width median (K4) -> column centres (K5)   [-> corpus histograms (K6)]

All buffers are allocated once; a step is a fixed sequence of libpagegeom.so launches on one
stream with no host synchronisation in between (each stage consumes the previous stage's
kept_idx / n_kept directly).  Pages are independent, so multi-GPU runs shard pages across
ranks with no collective; only the optional corpus histograms are all-reduced.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import ops
from ._lib import (PG_COL_HIST_BINS, PG_COL_SPAN_BYTES, PG_ERR_WORKSPACE, PG_WIDTH_HIST_BINS, check, lib, ptr,
                   stream_ptr)

KERNELS_PER_STEP = 12  # tiler, edge filter, 5 NMS kernels, class flags, width median, column prep + density + peaks


def shard_pages(n_pages_total: int, rank: int, world: int) -> range:
    """Contiguous page shard of ``rank`` (pages are independent units)."""
    per = (n_pages_total + world - 1) // world
    lo = min(n_pages_total, rank * per)
    return range(lo, min(n_pages_total, lo + per))


def allreduce_histograms(hist: torch.Tensor, stream=None) -> torch.Tensor:
    """K6 — the only exchange step of the path: corpus-level integer histograms (plain_text widths,
    column centres) summed over ranks in place.  Device histograms go through the C ABI (pg_hist_allreduce: one
    ncclAllReduce over NVLink on a communicator libpagegeom.so owns); host tensors — the world-size-2 `gloo`
    tests of the shard logic — through torch.distributed.  Integer sums are order-independent, so
    1/2/4/8-rank results are bit-identical."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if hist.is_cuda:
            ops.hist_allreduce(hist, stream)
        else:
            dist.all_reduce(hist, op=dist.ReduceOp.SUM)
    return hist


def corpus_median_width(width_hist: torch.Tensor) -> float:
    """Median of the all-reduced 1-px plain_text width histogram (lower-middle / upper-middle mean)."""
    h = width_hist.to(torch.int64).cpu().numpy()
    total = int(h.sum())
    if total == 0:
        return 0.0
    c = np.cumsum(h)
    lo = int(np.searchsorted(c, (total - 1) // 2 + 1))
    hi = int(np.searchsorted(c, total // 2 + 1))
    return (lo + hi) / 2.0


class PagePipeline:
    def __init__(self, plan: ops.TilePlan, n_pages: int, iou_threshold: float = 0.5, edge_threshold: float = 10,
                 min_margin_percent: float = 0.2, min_confidence: float = 0.3, max_cols: int = 64,
                 plain_text_id: float = 1.0, title_id: float = 0.0, corpus_stats: bool = False,
                 overlap: bool = True):
        self.plan, self.n_pages = plan, int(n_pages)
        # The tiler streams pixels (HBM-bound), the box stages are short latency-bound kernels on a few
        # SMs: run them side by side.  The tiler's CTAs retire after a few work items, so the box
        # kernels, queued on a higher-priority stream, pick up SM slots as they free.
        self.overlap = bool(overlap)
        if self.overlap:
            lo, hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)
            self.s_tiler = torch.cuda.Stream(priority=lo)
            self.s_box = torch.cuda.Stream(priority=hi)
            self._ev = [torch.cuda.Event() for _ in range(3)]
        self.iou_threshold, self.edge_threshold = float(iou_threshold), float(edge_threshold)
        self.min_margin_percent, self.min_confidence = float(min_margin_percent), float(min_confidence)
        self.max_cols, self.plain_text_id, self.title_id = int(max_cols), float(plain_text_id), float(title_id)
        # `plan` is a TilePlan (n_pages pages of one size, contiguous) or a TileBatch (pages of mixed sizes,
        # already bound to their buffers): the box stages do not care, page sizes travel in page_wh
        self.batch = plan if isinstance(plan, ops.TileBatch) else None
        self.tiles_out = None if self.batch is not None else plan.alloc_out(self.n_pages)
        self.n_boxes = 0
        self.corpus_stats = corpus_stats
        self.width_hist = self.col_hist = None
        if corpus_stats:
            self.hist = torch.zeros(PG_WIDTH_HIST_BINS + PG_COL_HIST_BINS, dtype=torch.int32, device="cuda")
            self.width_hist = self.hist[:PG_WIDTH_HIST_BINS]
            self.col_hist = self.hist[PG_WIDTH_HIST_BINS:]
        self.gauss = ops.gauss_table()
        self.max_bins = ops.MAX_DENSITY_BINS
        self.records = False

    # ---------------------------------------------------------------- detections
    def set_detections(self, dets: Sequence[dict], stream=None, host_staging: Optional[dict] = None):
        """dets: one dict per page as produced by synth.page_detections (or replayed stage-1
        JSON): cells [C,4], box_cell [n], boxes_local [n,4], scores [n], classes [n]."""
        assert len(dets) == self.n_pages
        counts = [len(d["boxes_local"]) for d in dets]
        off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        cell_base = np.concatenate([[0], np.cumsum([len(d["cells"]) for d in dets])])
        host = {
            "boxes_local": np.concatenate([d["boxes_local"].reshape(-1, 4) for d in dets]).astype(np.float64),
            "box_cell": np.concatenate([d["box_cell"].astype(np.int64) + cell_base[i] for i, d in enumerate(dets)]).astype(np.int32),
            "cells": np.concatenate([d["cells"] for d in dets]).astype(np.float64),
            "scores": np.concatenate([d["scores"] for d in dets]).astype(np.float64),
            "classes": np.concatenate([d["classes"] for d in dets]).astype(np.float64),
            "page_off": off,
            "page_wh": np.asarray([[d["width"], d["height"]] for d in dets], np.int32),
        }
        self._alloc_boxes(int(off[-1]), int(max(counts) if counts else 0))
        self.upload_detections(host, stream)
        return host

    def _alloc_boxes(self, n: int, max_per_page: int):
        if n == self.n_boxes and getattr(self, "max_per_page", -1) == max_per_page:
            return
        self.n_boxes, self.max_per_page = n, max_per_page
        p, dev = self.n_pages, "cuda"
        m = max(n, 1)
        self.boxes_local = torch.empty((m, 4), dtype=torch.float64, device=dev)
        self.boxes_page = torch.empty((m, 4), dtype=torch.float64, device=dev)
        self.box_cell = torch.empty(m, dtype=torch.int32, device=dev)
        self.scores = torch.empty(m, dtype=torch.float64, device=dev)
        self.classes = torch.empty(m, dtype=torch.float64, device=dev)
        self.flags = torch.empty(m, dtype=torch.uint8, device=dev)
        self.page_off = torch.empty(p + 1, dtype=torch.int64, device=dev)
        self.page_wh = torch.empty((p, 2), dtype=torch.int32, device=dev)
        self.cells = None
        self.kept1 = torch.empty(m, dtype=torch.int32, device=dev)
        self.n_kept1 = torch.zeros(p, dtype=torch.int32, device=dev)
        self.kept2 = torch.empty(m, dtype=torch.int32, device=dev)
        self.n_kept2 = torch.zeros(p, dtype=torch.int32, device=dev)
        self.median = torch.zeros(p, dtype=torch.float64, device=dev)
        self.n_bins = torch.zeros(p, dtype=torch.int32, device=dev)
        self.ws_keys = torch.empty(2 * m, dtype=torch.float64, device=dev)
        self.ws_counts = torch.empty(m, dtype=torch.int32, device=dev)
        self.centers = torch.zeros((p, self.max_cols), dtype=torch.int32, device=dev)
        self.col_widths = torch.zeros((p, self.max_cols), dtype=torch.float64, device=dev)
        self.n_cols = torch.zeros(p, dtype=torch.int32, device=dev)
        self.col_ws = torch.empty((p, 2 * self.max_bins), dtype=torch.float64, device=dev)
        self.col_spans = torch.empty(m * PG_COL_SPAN_BYTES, dtype=torch.uint8, device=dev)
        self.col_span_n = torch.zeros(p, dtype=torch.int32, device=dev)
        self.nms_ws = ops.NmsWorkspace(n, p)

    def upload_detections(self, host: Dict[str, np.ndarray], stream=None, pinned: Optional[dict] = None) -> int:
        """Host -> device copy of one step's detections.  Returns bytes copied."""
        nbytes = 0
        if self.cells is None or self.cells.shape[0] != host["cells"].shape[0]:
            self.cells = torch.empty((host["cells"].shape[0], 4), dtype=torch.float64, device="cuda")
        for name, dst in (("boxes_local", self.boxes_local), ("box_cell", self.box_cell), ("cells", self.cells),
                          ("scores", self.scores), ("classes", self.classes), ("page_off", self.page_off),
                          ("page_wh", self.page_wh)):
            src = pinned[name] if pinned is not None else torch.from_numpy(host[name])
            n = src.shape[0]
            if n:
                dst[:n].copy_(src.view(dst[:n].shape), non_blocking=True)
            nbytes += src.numel() * src.element_size()
        return nbytes

    # ---------------------------------------------------------------- one step
    def run(self, pages: torch.Tensor, stream=None, tiler_events=None) -> None:
        """One pass of the hot path over the shard.  pages: cuda uint8 [P, H, pitch]."""
        main = stream if stream is not None else torch.cuda.current_stream()
        if self.overlap and self.n_boxes:
            start, box_done, tiler_done = self._ev
            start.record(main)
            self.s_box.wait_event(start)
            self.s_tiler.wait_event(start)
            if os.environ.get("PG_TILER_FIRST") == "1":  # tuning knob
                self._run_tiler(pages, self.s_tiler, tiler_events)
                tiler_done.record(self.s_tiler)
                self._run_boxes(self.s_box)
                box_done.record(self.s_box)
            else:
                self._run_boxes(self.s_box)       # queued first: its kernels outrank the tiler's CTAs
                box_done.record(self.s_box)
                self._run_tiler(pages, self.s_tiler, tiler_events)
                tiler_done.record(self.s_tiler)
            main.wait_event(box_done)
            main.wait_event(tiler_done)
        else:
            self._run_tiler(pages, main, tiler_events)
            if self.n_boxes:
                self._run_boxes(main)

    def capture(self, pages: torch.Tensor) -> "torch.cuda.CUDAGraph":
        """One step captured into a CUDA graph (both streams: the fork/join through events is recorded as graph
        dependencies).  Replaying it costs one launch instead of 13-17 ctypes calls + launches — what matters
        for short steps (cfg2: 19 pages, 0.4 ms).  Run at least one eager step first: plans upload their
        tables lazily, which is not capturable."""
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.graph(graph, stream=side):
            self.run(pages)
        torch.cuda.current_stream().wait_stream(side)
        return graph

    def _run_tiler(self, pages, stream, tiler_events=None) -> None:
        L, s, p, plan = lib(), stream.cuda_stream, self.n_pages, self.plan
        if tiler_events is not None:
            tiler_events[0].record(stream)
        if self.batch is not None:
            check(L.pg_tile_letterbox_batch(self.batch._h, s))
        else:
            check(L.pg_tile_letterbox(plan._h, ptr(pages), p, pages.shape[2], pages.shape[1] * pages.shape[2],
                                      ptr(self.tiles_out), self.tiles_out.stride(0), s))
        if tiler_events is not None:
            tiler_events[1].record(stream)

    def _run_boxes(self, stream) -> None:
        L, s, p = lib(), stream.cuda_stream, self.n_pages
        check(L.pg_edge_filter(ptr(self.boxes_local), 1, ptr(self.box_cell), ptr(self.cells), ptr(self.page_wh),
                               ptr(self.page_off), p, self.edge_threshold, ptr(self.boxes_page), None,
                               ptr(self.kept1), ptr(self.n_kept1), s))
        check(L.pg_nms_merge(ptr(self.boxes_page), ptr(self.scores), ptr(self.classes), ptr(self.kept1),
                             ptr(self.page_off), ptr(self.n_kept1), p, self.n_boxes, self.max_per_page,
                             self.iou_threshold, ptr(self.kept2), ptr(self.n_kept2), self.nms_ws.ptr,
                             self.nms_ws.nbytes, s))
        check(L.pg_class_flags(ptr(self.classes), self.n_boxes, self.plain_text_id, self.title_id, ptr(self.flags), s))
        check(L.pg_width_median(ptr(self.boxes_page), ptr(self.flags), ptr(self.kept2), ptr(self.page_off),
                                ptr(self.n_kept2), p, self.n_boxes, ptr(self.page_wh), self.min_margin_percent, ptr(self.median),
                                ptr(self.n_bins), ptr(self.ws_keys), ptr(self.ws_counts), ptr(self.width_hist), s))
        check(L.pg_column_peaks(ptr(self.boxes_page), ptr(self.flags), ptr(self.scores), ptr(self.kept2),
                                ptr(self.page_off), ptr(self.n_kept2), p, ptr(self.page_wh), ptr(self.median),
                                ptr(self.gauss.table), ptr(self.gauss.offsets), self.gauss.max_window,
                                self.min_confidence, self.max_cols, ptr(self.centers), ptr(self.col_widths),
                                ptr(self.n_cols), ptr(self.col_ws), self.max_bins, ptr(self.col_spans), ptr(self.col_span_n),
                                ptr(self.col_hist), s))
        if self.records:
            self._run_records(stream)

    # ---------------------------------------------------------------- stage-3 records (J1-J4)
    def enable_records(self, heads: Sequence[bytes], tails: Sequence[bytes], names: Sequence[bytes], name_id,
                       bytes_per_box: int = 260) -> None:
        """Have every step also lay out the `<base>_combined.json` documents of the shard on the device
        (pg_json_combined, 4 more kernels on the box stream).  heads/tails: ops.combined_head_tail per page;
        names: JSON string literals; name_id[box] indexes names.  Call after set_detections."""
        p, n = self.n_pages, self.n_boxes
        assert len(heads) == p and len(tails) == p and n > 0
        pieces = list(heads) + list(tails) + list(names)
        offs = np.concatenate([[0], np.cumsum([len(x) for x in pieces])]).astype(np.int64)
        self.rec_text = torch.frombuffer(bytearray(b"".join(pieces) + b"\0"), dtype=torch.uint8).cuda()
        self.rec_head_off = torch.from_numpy(offs[:p + 1].copy()).cuda()
        self.rec_tail_off = torch.from_numpy(offs[p:2 * p + 1].copy()).cuda()
        self.rec_name_off = torch.from_numpy(offs[2 * p:].copy()).cuda()
        self.rec_name_id = ops._dev(name_id, torch.int32)
        self.rec_ws_bytes = int(lib().pg_json_workspace_bytes(n, p))
        self.rec_ws = torch.empty(self.rec_ws_bytes, dtype=torch.uint8, device="cuda")
        self.rec_capacity = int(offs[2 * p]) + 80 * p + n * bytes_per_box
        self.rec_out = torch.empty(self.rec_capacity, dtype=torch.uint8, device="cuda")
        self.rec_off = torch.zeros(p + 1, dtype=torch.int64, device="cuda")
        self.records = True

    def _run_records(self, stream) -> None:
        check(lib().pg_json_combined(ptr(self.boxes_page), ptr(self.classes), ptr(self.scores), ptr(self.rec_name_id),
                                     ptr(self.kept2), ptr(self.page_off), ptr(self.n_kept2), self.n_pages, self.n_boxes,
                                     self.max_per_page, ptr(self.rec_text), ptr(self.rec_head_off), ptr(self.rec_tail_off),
                                     ptr(self.rec_name_off), ptr(self.rec_out), self.rec_capacity, ptr(self.rec_off),
                                     ptr(self.rec_ws), self.rec_ws_bytes, stream.cuda_stream))

    def records_to_host(self) -> List[bytes]:
        """The documents of the last step (synchronises).  A shard whose text outgrew the buffer is re-run with
        the exact size — the device reports it in rec_off[n_pages]."""
        torch.cuda.synchronize()
        off = self.rec_off.cpu().numpy()
        if int(off[-1]) > self.rec_capacity:
            self.rec_capacity = int(off[-1])
            self.rec_out = torch.empty(self.rec_capacity, dtype=torch.uint8, device="cuda")
            self._run_records(torch.cuda.current_stream())
            torch.cuda.synchronize()
            off = self.rec_off.cpu().numpy()
        host = self.rec_out[:int(off[-1])].cpu().numpy().tobytes()
        return [host[off[i]:off[i + 1]] for i in range(self.n_pages)]

    def allreduce_corpus_stats(self):
        """K6: the one exchange step of the path — integer histograms summed over ranks
        (NCCL over NVLink; order-independent, so 1/2/4/8-GPU results are bit-identical).
        In place, so call it ONCE, when the shard is finished."""
        if self.corpus_stats:
            allreduce_histograms(self.hist, torch.cuda.current_stream())
        return self.hist

    def exchange_corpus_stats_async(self):
        """Running form of K6, callable after every step: a snapshot of the rank's running histograms is
        summed over ranks (pg_hist_allreduce) on a stream of its own, ordered after this step's box kernels only,
        so the exchange runs under the next steps' tiler instead of serialising the ranks.  `self.hist` stays
        rank-local; the corpus-wide totals so far are `self.hist_global` once `finish_exchange()` has been called."""
        if not self.corpus_stats:
            return
        if not hasattr(self, "_xchg_buf"):
            self._xchg_buf = [torch.zeros_like(self.hist) for _ in range(2)]
            self._xchg_i = 0
            self.s_xchg = torch.cuda.Stream()
        k = self._xchg_i & 1
        self._xchg_i += 1
        side = self.s_box if self.overlap else torch.cuda.current_stream()
        self.s_xchg.wait_stream(side)  # the step's box kernels (and, stream-ordered, this buffer's previous exchange)
        with torch.cuda.stream(self.s_xchg):
            self._xchg_buf[k].copy_(self.hist)
            allreduce_histograms(self._xchg_buf[k], self.s_xchg)
        self.hist_global = self._xchg_buf[k]

    def finish_exchange(self, stream=None):
        """Make `stream` (default: current) wait for every outstanding running exchange."""
        if not hasattr(self, "_xchg_buf"):
            return getattr(self, "hist", None)
        s = stream if stream is not None else torch.cuda.current_stream()
        s.wait_stream(self.s_xchg)
        if self.overlap:
            s.wait_stream(self.s_box)
        return self.hist_global

    # ---------------------------------------------------------------- results
    def results_to_host(self, pinned: Optional[dict] = None) -> Dict[str, np.ndarray]:
        """Device -> host read of the step's results (what the stage-3/4/5 JSON writers need)."""
        out = {}
        for name in ("kept2", "n_kept2", "median", "n_bins", "centers", "col_widths", "n_cols"):
            t = getattr(self, name)
            if pinned is not None:
                pinned[name].copy_(t, non_blocking=True)
                out[name] = pinned[name]
            else:
                out[name] = t.cpu()
        return out

    def result_bytes(self) -> int:
        return sum(getattr(self, n).numel() * getattr(self, n).element_size()
                   for n in ("kept2", "n_kept2", "median", "n_bins", "centers", "col_widths", "n_cols"))

    def check_status(self) -> dict:
        """After a step (synchronises): the merge's status word and the column kernel's limits.  A candidate list
        that outgrew the NMS workspace (PG_ERR_WORKSPACE: n_kept = -1, and everything downstream of it in that
        step is void) is handled like the command line does: the workspace is re-sized to the dense bound and the
        box stages of the step are run again before anything is reported."""
        torch.cuda.synchronize()
        st = self.nms_ws.stats()
        if st["status"] == PG_ERR_WORKSPACE:
            self.nms_ws = ops.NmsWorkspace(self.n_boxes, self.n_pages, pairs_per_block=self.max_per_page // 32 + 2)
            self._run_boxes(torch.cuda.current_stream())
            torch.cuda.synchronize()
            st = self.nms_ws.stats()
        if st["status"] != 0:
            raise RuntimeError(f"NMS merge reported an error on device: {st}")
        if bool((self.n_cols < 0).any().item()):
            raise RuntimeError("column kernel: a page exceeded the kernel limits")
        if bool((self.n_cols > self.max_cols).any().item()):
            raise RuntimeError("column kernel: more columns than max_cols")
        return st


class ScanPipeline:
    """The whole path on COMPRESSED input, end to end through host buffers:

        JPEG files (host bytes) + per-tile detections (host arrays)
          -> H2D -> decode on the device (D1-D8) -> one-channel tiler (K1) -> box stages (K2-K5) -> results D2H

    — what `cv2.imread` + the five scripts do per page in the reference (1:381 onwards), with the pixels crossing
    PCIe as the file's bytes.  Pages of one size (one TilePlan, channels=1), `n_pages` per step.  Steps are
    double-buffered: the host->device copy of step i+1 runs on its own stream under the kernels of step i.

        k = sp.submit(blob, file_off, detections)   # asynchronous; blob: pinned uint8 CPU tensor (ops.pack_files)
        res = sp.results(k)                          # waits for that step, checks the status words, host arrays

    The caller keeps `blob` unchanged until the step's copy has run (results(k) is late enough) and reads a
    step's results before submitting step k + depth, which reuses the slot's pinned result buffers."""

    def __init__(self, page_w: int, page_h: int, n_pages: int, grids=((4, 4),), overlap_percentage: float = 20.0,
                 depth: int = 2, **box_kw):
        self.n_pages, self.depth = int(n_pages), int(depth)
        self.plan = ops.TilePlan(page_w, page_h, grids, overlap_percentage, channels=1)
        self.s_copy = torch.cuda.Stream()
        self.s_main = torch.cuda.Stream()
        self.slots = []
        for _ in range(self.depth):
            pipe = PagePipeline(self.plan, self.n_pages, **box_kw)
            self.slots.append({"pipe": pipe, "pages": self.plan.alloc_pages(self.n_pages), "blob": None, "dec": ops.JpegDecoder(),
                               "copied": torch.cuda.Event(), "done": torch.cuda.Event(), "pin_in": None, "pin_out": None,
                               "busy": False})
        self.step = 0
        self.h2d_bytes = self.d2h_bytes = 0
        self.trace = None  # set to [] to record (copy start, copy end, compute start, compute end) events per step

    def timeline_ms(self):
        """With `trace` enabled: per step, the four instants in ms from the first step's copy start."""
        if not self.trace:
            return []
        torch.cuda.synchronize()
        base = self.trace[0][0]
        return [[round(base.elapsed_time(e), 3) for e in evs] for evs in self.trace]

    def submit(self, blob: torch.Tensor, file_off, host_dets: Dict[str, np.ndarray]) -> int:
        k = self.step % self.depth
        sl = self.slots[k]
        pipe, dec = sl["pipe"], sl["dec"]
        sizes = dec.set_files(blob, file_off)  # host: headers of this step's files
        if len(sizes) != self.n_pages or any(sz != (self.plan.page_w, self.plan.page_h, 1) for sz in sizes):
            raise ValueError(f"ScanPipeline was built for {self.n_pages} grey pages of {self.plan.page_w}x{self.plan.page_h}: {sizes[:3]}")
        if sl["blob"] is None or sl["blob"].numel() < blob.numel():
            sl["blob"] = torch.empty(blob.numel() + blob.numel() // 4, dtype=torch.uint8, device="cuda")
        n_boxes = int(host_dets["page_off"][-1])
        counts = np.diff(host_dets["page_off"])
        pipe._alloc_boxes(n_boxes, int(counts.max()) if len(counts) else 0)
        if sl["busy"]:
            sl["copied"].synchronize()  # the slot's previous host->device copies have left its pinned staging
        if sl["pin_in"] is None or any(sl["pin_in"][name].shape != host_dets[name].shape for name in host_dets):
            sl["pin_in"] = {name: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for name, v in host_dets.items()}
        else:
            for name, v in host_dets.items():
                sl["pin_in"][name].numpy()[...] = v
        if sl["pin_out"] is None or sl["pin_out"]["kept2"].shape != pipe.kept2.shape:
            sl["pin_out"] = {n: torch.empty_like(getattr(pipe, n), device="cpu").pin_memory()
                             for n in ("kept2", "n_kept2", "median", "n_bins", "centers", "col_widths", "n_cols")}
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if self.trace is not None else None
        with torch.cuda.stream(self.s_copy):
            if sl["busy"]:
                self.s_copy.wait_event(sl["done"])  # the slot's previous step has consumed its buffers
            if ev:
                ev[0].record(self.s_copy)
            outs = [sl["pages"][i] for i in range(self.n_pages)]
            dec.stage_tables(outs, stream=self.s_copy)  # the decoder's tables travel in front of the files, not behind the next step's
            sl["blob"][:blob.numel()].copy_(blob, non_blocking=True)
            box_bytes = pipe.upload_detections(host_dets, pinned=sl["pin_in"])
            if ev:
                ev[1].record(self.s_copy)
            sl["copied"].record(self.s_copy)
        self.s_main.wait_event(sl["copied"])
        with torch.cuda.stream(self.s_main):
            if ev:
                ev[2].record(self.s_main)
            dec.decode(sl["blob"], outs, stream=self.s_main)
            pipe.run(sl["pages"], stream=self.s_main)
            pipe.results_to_host(pinned=sl["pin_out"])
            if ev:
                ev[3].record(self.s_main)
                self.trace.append(ev)
            sl["done"].record(self.s_main)
        sl["busy"] = True
        self.h2d_bytes = int(blob.numel() + box_bytes)
        self.d2h_bytes = pipe.result_bytes()
        self.step += 1
        return self.step - 1

    def results(self, step: int) -> Dict[str, np.ndarray]:
        sl = self.slots[step % self.depth]
        sl["done"].synchronize()
        st = sl["dec"].status()
        if st["status"] != 0 or sl["pipe"].nms_ws.stats()["status"] != 0:
            # rare: chunk states not converged in the configured rounds / NMS candidates overflowed — redo the step's
            # device part with the larger settings the checks install
            torch.cuda.synchronize()
            with torch.cuda.stream(self.s_main):
                sl["dec"].check()
                sl["pipe"].run(sl["pages"], stream=self.s_main)
                self.s_main.synchronize()
                sl["pipe"].check_status()
                sl["pipe"].results_to_host(pinned=sl["pin_out"])
            self.s_main.synchronize()
        return {k: v.numpy() for k, v in sl["pin_out"].items()}

    def drain(self):
        for sl in self.slots:
            if sl["busy"]:
                sl["done"].synchronize()
