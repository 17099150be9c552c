"""Device-level wrappers: one function per C-ABI entry point, torch tensors in/out.

PyTorch is only the allocator / stream provider here; every computation is a kernel of
libpagegeom.so.  All tensors must live on the current CUDA device.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import PgTileInfo, check, lib, ptr, stream_ptr


def _require_cuda():
    if not torch.cuda.is_available():
        raise _lib.PageGeomError("a CUDA device is required: the page-geometry path has no CPU fallback")


# density bins of 5_detect_column_centers.py:120-121: W // max(1, W // 1000) + 1 <= 2000 for every W
MAX_DENSITY_BINS = 2048


def row_pitch(width: int, channels: int = 3) -> int:
    """Row pitch (bytes) of a uint8 page for the tiler: channels*W rounded up to 16 (bulk-copy alignment)."""
    return (int(channels) * int(width) + 15) // 16 * 16


# --------------------------------------------------------------------------------------------
# K1 tiler
# --------------------------------------------------------------------------------------------
class TilePlan:
    """Tiling + letterbox plan for one page size (pg_tile_plan_*).

    grids: sequence of (rows, cols); (1, 1) is the reference's full-page pass.
    Tiles are enumerated grid-major then row-major, like 1_doclayout_bboxes.py:756,399-400.
    """

    def __init__(self, page_w: int, page_h: int, grids: Sequence[Tuple[int, int]] = ((2, 2),),
                 overlap: float = 20.0, imgsz: int = 1024, stride: int = 32, auto: bool = True,
                 scaleup: bool = True, channels: int = 3):
        self.page_w, self.page_h, self.channels = int(page_w), int(page_h), int(channels)
        self.grids = [(int(r), int(c)) for r, c in grids]
        self.overlap, self.imgsz, self.stride, self.auto = float(overlap), int(imgsz), int(stride), bool(auto)
        rows = (C.c_int32 * len(self.grids))(*[g[0] for g in self.grids])
        cols = (C.c_int32 * len(self.grids))(*[g[1] for g in self.grids])
        handle = C.c_void_p()
        check(lib().pg_tile_plan_create_ex(self.page_w, self.page_h, self.channels, rows, cols, len(self.grids),
                                           self.overlap, self.imgsz, self.stride, int(auto), int(scaleup),
                                           C.byref(handle)))
        self._h = handle
        self.tiles: List[dict] = []
        for t in range(lib().pg_tile_plan_num_tiles(self._h)):
            info = PgTileInfo()
            check(lib().pg_tile_plan_tile(self._h, t, C.byref(info)))
            self.tiles.append({name: getattr(info, name) for name, _ in PgTileInfo._fields_})
        self.out_elems = int(lib().pg_tile_plan_out_elems(self._h))
        self.algorithmic_bytes = int(lib().pg_tile_plan_algorithmic_bytes(self._h))
        self.pitch = row_pitch(self.page_w, self.channels)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                lib().pg_tile_plan_destroy(h)
            except Exception:
                pass
            self._h = None

    # -- cell coordinates with the reference's int/float mixing (1_doclayout_bboxes.py:418-421)
    def cell_coordinates(self, tile: int) -> dict:
        t = self.tiles[tile]

        def lo(v):
            return 0 if v == 0 else v

        def hi(v, bound):
            return bound if v == bound else v

        return {"x_start": lo(t["x_start"]), "y_start": lo(t["y_start"]),
                "x_end": hi(t["x_end"], self.page_w), "y_end": hi(t["y_end"], self.page_h)}

    def alloc_pages(self, n_pages: int) -> torch.Tensor:
        _require_cuda()
        return torch.empty((n_pages, self.page_h, self.pitch), dtype=torch.uint8, device="cuda")

    def alloc_out(self, n_pages: int) -> torch.Tensor:
        _require_cuda()
        return torch.empty((n_pages, self.out_elems), dtype=torch.float16, device="cuda")

    def run(self, pages: torch.Tensor, out: Optional[torch.Tensor] = None, stream=None,
            direct: bool = False) -> torch.Tensor:
        """pages: cuda uint8 [P, H, pitch] (BGR interleaved rows, or one grey plane for channels=1; pitch % 16 == 0).
        Returns fp16 [P, out_elems]; tile_view() slices a tile out of it."""
        _require_cuda()
        assert pages.is_cuda and pages.dtype == torch.uint8 and pages.dim() == 3 and pages.is_contiguous()
        n_pages, h, pitch = pages.shape
        assert h == self.page_h, (h, self.page_h)
        if out is None:
            out = self.alloc_out(n_pages)
        assert out.is_cuda and out.dtype == torch.float16 and out.is_contiguous() and out.shape[0] == n_pages
        fn = lib().pg_tile_letterbox_direct if direct else lib().pg_tile_letterbox
        check(fn(self._h, ptr(pages), n_pages, pitch, h * pitch, ptr(out), out.stride(0), stream_ptr(stream)))
        return out

    def tile_view(self, out: torch.Tensor, page: int, tile: int) -> torch.Tensor:
        t = self.tiles[tile]
        n = 3 * t["out_h"] * t["out_w"]
        return out[page, t["out_offset"]: t["out_offset"] + n].view(3, t["out_h"], t["out_w"])


class TileBatch:
    """Pages of different sizes tiled by one launch (pg_tile_batch_*).

    sizes: [(W, H)] per page; one TilePlan is built per distinct size.  ``pages[i]`` is a cuda uint8
    tensor [H_i, pitch_i]; outputs are allocated here as one fp16 tensor per page."""

    def __init__(self, sizes: Sequence[Tuple[int, int]], grids: Sequence[Tuple[int, int]] = ((2, 2),),
                 overlap: float = 20.0, imgsz: int = 1024, stride: int = 32, auto: bool = True, channels: int = 3):
        _require_cuda()
        self.channels = int(channels)
        self.sizes = [(int(w), int(h)) for w, h in sizes]
        self.plans: List[TilePlan] = []
        index = {}
        self.page_plan = []
        for s in self.sizes:
            if s not in index:
                index[s] = len(self.plans)
                self.plans.append(TilePlan(s[0], s[1], grids, overlap, imgsz, stride, auto, channels=self.channels))
            self.page_plan.append(index[s])
        n = len(self.sizes)
        plan_arr = (C.c_void_p * len(self.plans))(*[p._h for p in self.plans])
        pp = (C.c_int32 * n)(*self.page_plan)
        handle = C.c_void_p()
        check(lib().pg_tile_batch_create(plan_arr, len(self.plans), pp, n, C.byref(handle)))
        self._h = handle
        self.algorithmic_bytes = int(lib().pg_tile_batch_algorithmic_bytes(self._h))
        self.pages: List[torch.Tensor] = []
        self.outs: List[torch.Tensor] = []

    def plan_of(self, page: int) -> TilePlan:
        return self.plans[self.page_plan[page]]

    def alloc_pages(self) -> List[torch.Tensor]:
        return [torch.empty((h, row_pitch(w, self.channels)), dtype=torch.uint8, device="cuda") for (w, h) in self.sizes]

    def bind(self, pages: Sequence[torch.Tensor], stream=None) -> List[torch.Tensor]:
        n = len(self.sizes)
        assert len(pages) == n
        for t, (w, h) in zip(pages, self.sizes):
            assert t.is_cuda and t.dtype == torch.uint8 and t.dim() == 2 and t.shape[0] == h and t.is_contiguous()
        self.pages = list(pages)
        self.outs = [torch.empty(self.plan_of(i).out_elems, dtype=torch.float16, device="cuda") for i in range(n)]
        src = (C.c_void_p * n)(*[t.data_ptr() for t in self.pages])
        pit = (C.c_int64 * n)(*[t.shape[1] for t in self.pages])
        dst = (C.c_void_p * n)(*[t.data_ptr() for t in self.outs])
        check(lib().pg_tile_batch_bind(self._h, src, pit, dst, stream_ptr(stream)))
        return self.outs

    def run(self, stream=None) -> List[torch.Tensor]:
        check(lib().pg_tile_letterbox_batch(self._h, stream_ptr(stream)))
        return self.outs

    def tile_view(self, page: int, tile: int) -> torch.Tensor:
        t = self.plan_of(page).tiles[tile]
        n = 3 * t["out_h"] * t["out_w"]
        return self.outs[page][t["out_offset"]: t["out_offset"] + n].view(3, t["out_h"], t["out_w"])

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                lib().pg_tile_batch_destroy(h)
            except Exception:
                pass
            self._h = None


def upload_pages(images: Sequence[np.ndarray], plan: TilePlan, stream=None) -> torch.Tensor:
    """Host BGR uint8 HxWx3 arrays -> pitched cuda tensor [P, H, pitch] through pinned memory."""
    _require_cuda()
    n = len(images)
    host = torch.zeros((n, plan.page_h, plan.pitch), dtype=torch.uint8).pin_memory()
    hv = host.numpy()
    for i, img in enumerate(images):
        assert img.shape == (plan.page_h, plan.page_w, 3) and img.dtype == np.uint8
        hv[i, :, : 3 * plan.page_w] = img.reshape(plan.page_h, 3 * plan.page_w)
    return host.to("cuda", non_blocking=True)


def decode_page(path) -> Optional[np.ndarray]:
    """Host decode of one scan, as the reference's cv2.imread (1_doclayout_bboxes.py:381): BGR uint8 HxWx3 — or, for
    a greyscale file (where cv2.imread's three channels are equal), the single plane HxW, which crosses PCIe at a
    third of the bytes and is tiled by the one-channel plans.  cv2.IMREAD_ANYCOLOR is IMREAD_COLOR without the
    replication: 8-bit, EXIF orientation applied.  None when the file cannot be decoded (cv2's own convention)."""
    import cv2
    img = cv2.imread(str(path), cv2.IMREAD_ANYCOLOR)
    if img is None or img.dtype != np.uint8 or img.ndim not in (2, 3) or (img.ndim == 3 and img.shape[2] != 3):
        img = cv2.imread(str(path))
    return img


def upload_pages_pinned(images: Sequence[np.ndarray], stream=None) -> List[torch.Tensor]:
    """Host pages of any sizes — BGR HxWx3 or grey HxW, uniformly — -> one pitched cuda tensor [H, pitch] per page
    (TileBatch.bind's input).  All pages go through ONE pinned staging buffer and one asynchronous copy."""
    _require_cuda()
    chans = [1 if img.ndim == 2 else 3 for img in images]
    sizes = [(img.shape[0], row_pitch(img.shape[1], c)) for img, c in zip(images, chans)]
    offs = np.concatenate([[0], np.cumsum([h * p for h, p in sizes])]).astype(np.int64)
    host = torch.empty(int(offs[-1]), dtype=torch.uint8).pin_memory()
    hv = host.numpy()
    for img, c, (h, p), o in zip(images, chans, sizes, offs):
        assert img.dtype == np.uint8 and (img.ndim == 2 or img.shape[2] == 3)
        hv[o:o + h * p].reshape(h, p)[:, : c * img.shape[1]] = img.reshape(h, c * img.shape[1])
    if stream is not None:
        with torch.cuda.stream(stream):
            dev = host.to("cuda", non_blocking=True)
    else:
        dev = host.to("cuda", non_blocking=True)
    return [dev[o:o + h * p].view(h, p) for (h, p), o in zip(sizes, offs)]


# --------------------------------------------------------------------------------------------
# D1-D8 JPEG scans decoded on the device
# --------------------------------------------------------------------------------------------
class JpegDecoder:
    """pg_jpeg_*: baseline JPEG files -> pages in HBM, bit for bit what cv2.imread returns: a greyscale file gives ONE
    grey plane [H, pitch] (cv2's three channels are equal; tile it with channels=1 plans), a colour file gives BGR
    interleaved [H, pitch >= 3W] like cv2.  Usage per batch of files:

        sizes = dec.set_files(blob, file_off)        # host: headers parsed; blob = files back to back
        pages = dec.decode(blob_dev)                 # device, asynchronous; list of uint8 [H, pitch] tensors
        dec.check()                                  # after a synchronisation: raises / retries on non-convergence

    `blob` is a uint8 numpy array or (pinned) CPU tensor; `blob_dev` the same bytes on the GPU."""

    def __init__(self, chunk_bytes: int = 0, sync_rounds: int = 4):
        handle = C.c_void_p()
        check(lib().pg_jpeg_decoder_create(C.byref(handle)))
        self._h = handle
        self.chunk_bytes, self.sync_rounds = int(chunk_bytes), int(sync_rounds)
        self.auto_chunk = chunk_bytes <= 0
        self.ws: Optional[torch.Tensor] = None
        self.sizes: List[Tuple[int, int, int]] = []
        self._last = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                lib().pg_jpeg_decoder_destroy(h)
            except Exception:
                pass
            self._h = None

    def set_files(self, blob, file_off) -> List[Tuple[int, int, int]]:
        """Parses the headers (host).  Returns [(width, height, channels)].  Raises PageGeomError (error 4,
        'unsupported: ...') for anything but baseline / extended-sequential Huffman files (progressive, arithmetic,
        12-bit, CMYK, multi-scan, exotic sampling factors)."""
        arr = blob.numpy() if isinstance(blob, torch.Tensor) else np.asarray(blob)
        assert arr.dtype == np.uint8 and arr.ndim == 1 and arr.flags.c_contiguous
        off = np.ascontiguousarray(np.asarray(file_off, np.int64))
        n = len(off) - 1
        if self.auto_chunk:
            # chunk ~ 10 blocks of the batch's average block length (entropy bytes / 8x8 blocks is not known before
            # the headers are parsed, so the first pass uses the previous size and a re-parse follows only when it differs)
            self.chunk_bytes = self.chunk_bytes or 128
        check(lib().pg_jpeg_decoder_configure(self._h, self.chunk_bytes, self.sync_rounds))
        check(lib().pg_jpeg_decoder_set_files(self._h, arr.ctypes.data, off.ctypes.data, n))
        self.sizes = []
        w, h, c = C.c_int32(), C.c_int32(), C.c_int32()
        for i in range(n):
            check(lib().pg_jpeg_decoder_image_info(self._h, i, C.byref(w), C.byref(h), C.byref(c)))
            self.sizes.append((w.value, h.value, c.value))
        if self.auto_chunk:
            blocks = sum(((w + 7) // 8) * ((h + 7) // 8) for w, h, _ in self.sizes)
            per_block = float(off[-1] - off[0]) / max(blocks, 1)
            want = 128  # measured on B200 (8 pages of 48 Mpixel): shorter chunks = more, shorter walks — 256 bytes beat 1024
            # at quality 95 (2.8 vs 3.2 ms, 23 bytes per block), 128 beat 256 at quality 75 (2.0 vs 2.3 ms, 11 bytes
            # per block); ~10 blocks per chunk keep the share of chunks redone in sync round 1 near 0.1 %
            while want < 10 * per_block and want < 4096:
                want *= 2
            if want != self.chunk_bytes:
                self.chunk_bytes = want
                check(lib().pg_jpeg_decoder_configure(self._h, self.chunk_bytes, self.sync_rounds))
                check(lib().pg_jpeg_decoder_set_files(self._h, arr.ctypes.data, off.ctypes.data, n))
        self._host = (arr, off)
        return self.sizes

    def alloc_pages(self) -> List[torch.Tensor]:
        _require_cuda()
        return [torch.empty((h, row_pitch(w, c)), dtype=torch.uint8, device="cuda") for w, h, c in self.sizes]

    def _call_args(self, outs):
        n = len(self.sizes)
        assert len(outs) == n
        need = int(lib().pg_jpeg_workspace_bytes(self._h))
        if self.ws is None or self.ws.numel() < need + 256:
            self.ws = torch.empty(need + need // 8 + 256, dtype=torch.uint8, device="cuda")
        ws_ptr = self.ws.data_ptr() + ((-self.ws.data_ptr()) % 256)
        ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in outs])
        pitches = (C.c_int64 * n)(*[t.shape[1] for t in outs])
        return ptrs, pitches, ws_ptr, need

    def stage_tables(self, outs: Sequence[torch.Tensor], stream=None) -> None:
        """pg_jpeg_stage_tables: upload the batch's tables on the stream that carries the files (see pagegeom.h); the
        decode() that follows, ordered after it, then issues no host->device copy."""
        _require_cuda()
        ptrs, pitches, ws_ptr, need = self._call_args(outs)
        check(lib().pg_jpeg_stage_tables(self._h, ptrs, pitches, ws_ptr, need, stream_ptr(stream)))

    def decode(self, blob_dev: torch.Tensor, outs: Optional[Sequence[torch.Tensor]] = None, stream=None) -> List[torch.Tensor]:
        _require_cuda()
        assert blob_dev.is_cuda and blob_dev.dtype == torch.uint8 and blob_dev.is_contiguous()
        if outs is None:
            outs = self.alloc_pages()
        ptrs, pitches, ws_ptr, need = self._call_args(outs)
        check(lib().pg_jpeg_decode(self._h, ptr(blob_dev), ptrs, pitches, ws_ptr, need, stream_ptr(stream)))
        self._last = (blob_dev, list(outs), stream)
        return list(outs)

    def status(self) -> dict:
        arr = (C.c_int64 * 4)()
        check(lib().pg_jpeg_decode_status(self._h, arr))
        return {"status": arr[0], "rounds_used": arr[1], "states_replaced_round1": arr[2], "chunks": arr[3]}

    def check(self) -> dict:
        """After the decode's stream has been synchronised.  If the configured sync rounds did not reach the fixed
        point (long blocks against short chunks), the rounds are doubled and the batch decoded again."""
        st = self.status()
        while st["status"] != 0:
            if self.sync_rounds >= 64:
                raise _lib.PageGeomError(f"pg_jpeg_decode: chunk states did not converge: {st}")
            self.sync_rounds = min(64, self.sync_rounds * 2)
            arr, off = self._host
            check(lib().pg_jpeg_decoder_configure(self._h, self.chunk_bytes, self.sync_rounds))
            check(lib().pg_jpeg_decoder_set_files(self._h, arr.ctypes.data, off.ctypes.data, len(off) - 1))
            blob_dev, outs, stream = self._last
            self.decode(blob_dev, outs, stream)
            (stream.synchronize() if stream is not None else torch.cuda.current_stream().synchronize())
            st = self.status()
        return st


def jpeg_probe(data: bytes) -> Optional[Tuple[int, int, int]]:
    """(width, height, channels) when `data` is a JPEG file the device path decodes (baseline / extended sequential
    Huffman, 8-bit, one scan; greyscale, or YCbCr with 4:4:4 / 4:2:2 / 4:2:0 / 4:4:0 sampling), else None.  Host
    only: parses the headers (pg_hostcheck_jpeg_decode with no output buffer)."""
    b = np.frombuffer(data, np.uint8)
    w, h, st = C.c_int32(), C.c_int32(), (C.c_int64 * 4)()
    if lib().pg_hostcheck_jpeg_decode(b.ctypes.data, len(b), 512, 0, None, 0, C.byref(w), C.byref(h), st) != 0:
        return None
    return w.value, h.value, int(st[3])


class PinnedBuffer:
    """pg_pinned_alloc: page-locked host bytes as a uint8 CPU tensor (`.tensor`); write_combined=True for buffers the
    CPU only fills (file bytes on their way to the GPU).  Freed when the object goes away — keep it alive while
    copies from it are in flight."""

    def __init__(self, nbytes: int, write_combined: bool = False):
        p = C.c_void_p()
        check(lib().pg_pinned_alloc(int(nbytes), 1 if write_combined else 0, C.byref(p)))
        self._p, self.nbytes, self.write_combined = p, int(nbytes), bool(write_combined)
        self.array = np.ctypeslib.as_array((C.c_uint8 * int(nbytes)).from_address(p.value))
        self.tensor = torch.from_numpy(self.array)

    def __del__(self):
        p = getattr(self, "_p", None)
        if p:
            try:
                lib().pg_pinned_free(p)
            except Exception:
                pass
            self._p = None


def pack_files(files: Sequence[bytes], pinned: bool = True, write_combined: bool = False):
    """Files back to back, each starting on a 256-byte boundary -> (uint8 CPU tensor, int64 offsets [n+1] of the
    files' FIRST bytes plus the end of the last).  pg_jpeg_decoder_set_files takes file i as
    blob[off[i] .. off[i+1]); the alignment padding at a file's end is ignored by the parser (it lies behind EOI).
    write_combined: the tensor lives in a PinnedBuffer (kept alive through `blob._pg_owner`)."""
    off = [0]
    for f in files:
        off.append(off[-1] + (len(f) + 255) // 256 * 256)
    if write_combined and pinned and torch.cuda.is_available():
        owner = PinnedBuffer(off[-1] + 256, write_combined=True)
        blob = owner.tensor
        blob._pg_owner = owner
        owner.array[off[-1]:] = 0
    else:
        blob = torch.zeros(off[-1] + 256, dtype=torch.uint8)
        if pinned and torch.cuda.is_available():
            blob = blob.pin_memory()
    view = blob.numpy()
    for f, o in zip(files, off):
        view[o:o + len(f)] = np.frombuffer(f, np.uint8)
    return blob, np.asarray(off, np.int64)


def decode_jpeg_files(files: Sequence[bytes], decoder: Optional[JpegDecoder] = None) -> List[torch.Tensor]:
    """Convenience: JPEG files (bytes) -> pages on the GPU (grey plane or BGR, per file; synchronises)."""
    dec = decoder or JpegDecoder()
    blob, off = pack_files(files)
    dec.set_files(blob, off)
    pages = dec.decode(blob.to("cuda", non_blocking=True))
    torch.cuda.current_stream().synchronize()
    dec.check()
    return pages


def synth_pages(plan: TilePlan, n_pages: int, seed0: int, first_page: int = 0, out: Optional[torch.Tensor] = None,
                stream=None) -> torch.Tensor:
    _require_cuda()
    if out is None:
        out = plan.alloc_pages(n_pages)
    check(lib().pg_synth_pages(ptr(out), n_pages, plan.page_w, plan.page_h, plan.pitch, plan.page_h * plan.pitch,
                               seed0, first_page, stream_ptr(stream)))
    return out


# --------------------------------------------------------------------------------------------
# box stages
# --------------------------------------------------------------------------------------------
def _dev(a, dtype) -> torch.Tensor:
    if isinstance(a, torch.Tensor):
        t = a.to(device="cuda", dtype=dtype)
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(a))).to(device="cuda", dtype=dtype)
    return t.contiguous()


def edge_filter(boxes, box_cell, cells, page_wh, page_off, threshold: float = 10, boxes_are_local: bool = True,
                want_boxes_page: bool = True, stream=None):
    """pg_edge_filter.  Returns (boxes_page [N,4] f64 or None, keep [N] u8, kept_idx [N] i32, n_kept [P] i32)."""
    _require_cuda()
    boxes = _dev(boxes, torch.float64).view(-1, 4)
    box_cell = _dev(box_cell, torch.int32)
    cells = _dev(cells, torch.float64).view(-1, 4)
    page_wh = _dev(page_wh, torch.int32).view(-1, 2)
    page_off = _dev(page_off, torch.int64)
    n, p = boxes.shape[0], page_wh.shape[0]
    assert page_off.numel() == p + 1 and box_cell.numel() == n
    boxes_page = torch.empty_like(boxes) if want_boxes_page else None
    keep = torch.empty(n, dtype=torch.uint8, device="cuda")
    kept_idx = torch.empty(max(n, 1), dtype=torch.int32, device="cuda")
    n_kept = torch.zeros(max(p, 1), dtype=torch.int32, device="cuda")
    check(lib().pg_edge_filter(ptr(boxes), int(boxes_are_local), ptr(box_cell), ptr(cells), ptr(page_wh),
                               ptr(page_off), p, float(threshold), ptr(boxes_page), ptr(keep), ptr(kept_idx),
                               ptr(n_kept), stream_ptr(stream)))
    return boxes_page, keep, kept_idx, n_kept[:p]


@dataclass
class NmsWorkspace:
    n_boxes: int
    n_pages: int
    pairs_per_block: int = 64
    buf: torch.Tensor = field(init=False)

    def __post_init__(self):
        nbytes = int(lib().pg_nms_workspace_bytes(self.n_boxes, self.n_pages, self.pairs_per_block))
        self.buf = torch.empty(nbytes + 256, dtype=torch.uint8, device="cuda")
        self.offset = (-self.buf.data_ptr()) % 256
        self.nbytes = nbytes

    @property
    def ptr(self) -> int:
        return self.buf.data_ptr() + self.offset

    def stats(self) -> dict:
        arr = (C.c_int64 * 4)()
        check(lib().pg_nms_stats(self.ptr, arr))
        return {"status": arr[0], "candidate_block_pairs": arr[1], "rounds": arr[2], "box_pairs_tested": arr[3]}


def nms_merge(boxes, scores, classes, page_off, iou_threshold: float = 0.5, sel_idx=None, n_sel=None,
              max_boxes_per_page: int = 0, workspace: Optional[NmsWorkspace] = None, stream=None,
              kept_idx: Optional[torch.Tensor] = None, n_kept: Optional[torch.Tensor] = None, mode: int = 0,
              check_status: bool = False):
    """pg_nms_merge_ex.  Returns (kept_idx [N] i32 global indices in pick order, n_kept [P] i32, workspace).
    mode: 0 = stage-3 semantics (class-aware, fp64); PG_NMS_CLASS_AGNOSTIC | PG_NMS_FP32 = torchvision.ops.nms.
    check_status=True synchronises and reads the status word the kernels leave in the workspace: a candidate list that
    outgrew the workspace (PG_ERR_WORKSPACE, n_kept = -1) is re-run with the dense bound, any other error raises.
    check_status=False (launch-only, for callers that stay asynchronous) leaves that to the caller: `workspace.stats()`."""
    _require_cuda()
    boxes = _dev(boxes, torch.float64).view(-1, 4)
    scores = _dev(scores, torch.float64)
    classes = _dev(classes, torch.float64) if classes is not None else None
    page_off_t = _dev(page_off, torch.int64)
    n, p = boxes.shape[0], page_off_t.numel() - 1
    if sel_idx is not None:
        sel_idx = _dev(sel_idx, torch.int32)
        n_sel = _dev(n_sel, torch.int32)
    if workspace is None or workspace.n_boxes < n or workspace.n_pages < p:
        workspace = NmsWorkspace(n, p)
    if kept_idx is None:
        kept_idx = torch.empty(max(n, 1), dtype=torch.int32, device="cuda")
    if n_kept is None:
        n_kept = torch.zeros(max(p, 1), dtype=torch.int32, device="cuda")

    def launch(ws):
        check(lib().pg_nms_merge_ex(ptr(boxes), ptr(scores), ptr(classes), ptr(sel_idx), ptr(page_off_t), ptr(n_sel), p,
                                         n, int(max_boxes_per_page), float(iou_threshold), int(mode), ptr(kept_idx),
                                    ptr(n_kept), ws.ptr, ws.nbytes, stream_ptr(stream)))

    launch(workspace)
    if check_status:
        (stream.synchronize() if stream is not None else torch.cuda.current_stream().synchronize())
        st = workspace.stats()
        if st["status"] == _lib.PG_ERR_WORKSPACE:
            longest = int(max_boxes_per_page) if max_boxes_per_page > 0 else int((page_off_t[1:] - page_off_t[:-1]).max().item())
            workspace = NmsWorkspace(n, p, pairs_per_block=longest // 32 + 2)
            launch(workspace)
            (stream.synchronize() if stream is not None else torch.cuda.current_stream().synchronize())
            st = workspace.stats()
        if st["status"] != 0:
            raise _lib.PageGeomError(f"pg_nms_merge_ex failed on the device: {st}")
    return kept_idx, n_kept[:p], workspace


def class_flags(classes, plain_text_id: float = 1.0, title_id: float = 0.0, stream=None) -> torch.Tensor:
    _require_cuda()
    classes = _dev(classes, torch.float64)
    flags = torch.empty(classes.numel(), dtype=torch.uint8, device="cuda")
    check(lib().pg_class_flags(ptr(classes), classes.numel(), float(plain_text_id), float(title_id), ptr(flags),
                               stream_ptr(stream)))
    return flags


def width_median(boxes, flags, page_off, page_wh, min_margin_percent: float = 0.2, sel_idx=None, n_sel=None,
                 width_hist: Optional[torch.Tensor] = None, stream=None):
    """pg_width_median.  Returns (median [P] f64, n_bins [P] i32)."""
    _require_cuda()
    boxes = _dev(boxes, torch.float64).view(-1, 4)
    flags = _dev(flags, torch.uint8)
    page_off = _dev(page_off, torch.int64)
    page_wh = _dev(page_wh, torch.int32).view(-1, 2)
    n, p = boxes.shape[0], page_wh.shape[0]
    if sel_idx is not None:
        sel_idx = _dev(sel_idx, torch.int32)
        n_sel = _dev(n_sel, torch.int32)
    median = torch.zeros(max(p, 1), dtype=torch.float64, device="cuda")
    n_bins = torch.zeros(max(p, 1), dtype=torch.int32, device="cuda")
    ws_keys = torch.empty(2 * max(n, 1), dtype=torch.float64, device="cuda")
    ws_counts = torch.empty(max(n, 1), dtype=torch.int32, device="cuda")
    check(lib().pg_width_median(ptr(boxes), ptr(flags), ptr(sel_idx), ptr(page_off), ptr(n_sel), p, n, ptr(page_wh),
                                float(min_margin_percent), ptr(median), ptr(n_bins), ptr(ws_keys), ptr(ws_counts),
                                ptr(width_hist), stream_ptr(stream)))
    return median[:p], n_bins[:p]


class GaussTable:
    """Normalised scipy.signal.windows.gaussian(M, std=M/6) for every odd M <= max_window
    (5_detect_column_centers.py:151-153), built on the host with numpy so np.exp rounding is
    shared with the reference, uploaded once."""

    def __init__(self, max_window: int = 1023):
        _require_cuda()
        self.max_window = int(max_window) | 1
        offs, chunks, off = [], [], 0
        for k in range((self.max_window - 1) // 2 + 1):
            m = 2 * k + 1
            sigma = m / 6.0
            nn = np.arange(0, m) - (m - 1.0) / 2.0
            w = np.exp(-(nn ** 2) / (2 * sigma * sigma))
            w = w / w.sum()
            offs.append(off)
            chunks.append(w)
            off += m
        self.table = torch.from_numpy(np.concatenate(chunks)).to("cuda")
        self.offsets = torch.tensor(offs, dtype=torch.int64, device="cuda")


_GAUSS: Optional[GaussTable] = None


def gauss_table() -> GaussTable:
    global _GAUSS
    if _GAUSS is None:
        _GAUSS = GaussTable()
    return _GAUSS


def column_peaks(boxes, flags, scores, page_off, page_wh, median, min_confidence: float = 0.3, sel_idx=None,
                 n_sel=None, max_cols: int = 64, max_bins: int = 0, col_hist: Optional[torch.Tensor] = None,
                 stream=None, return_ws: bool = False):
    """pg_column_peaks.  Returns (centers [P,max_cols] i32, widths [P,max_cols] f64, n_cols [P] i32)."""
    _require_cuda()
    boxes = _dev(boxes, torch.float64).view(-1, 4)
    flags = _dev(flags, torch.uint8)
    scores = _dev(scores, torch.float64)
    page_off = _dev(page_off, torch.int64)
    page_wh_t = _dev(page_wh, torch.int32).view(-1, 2)
    median = _dev(median, torch.float64)
    p = page_wh_t.shape[0]
    if sel_idx is not None:
        sel_idx = _dev(sel_idx, torch.int32)
        n_sel = _dev(n_sel, torch.int32)
    if max_bins <= 0:
        max_bins = MAX_DENSITY_BINS
    gt = gauss_table()
    centers = torch.zeros((max(p, 1), max_cols), dtype=torch.int32, device="cuda")
    widths = torch.zeros((max(p, 1), max_cols), dtype=torch.float64, device="cuda")
    n_cols = torch.zeros(max(p, 1), dtype=torch.int32, device="cuda")
    ws = torch.empty((max(p, 1), 2 * max_bins), dtype=torch.float64, device="cuda")
    ws_spans = torch.empty(max(boxes.shape[0], 1) * _lib.PG_COL_SPAN_BYTES, dtype=torch.uint8, device="cuda")
    ws_span_counts = torch.zeros(max(p, 1), dtype=torch.int32, device="cuda")
    check(lib().pg_column_peaks(ptr(boxes), ptr(flags), ptr(scores), ptr(sel_idx), ptr(page_off), ptr(n_sel), p,
                                ptr(page_wh_t), ptr(median), ptr(gt.table), ptr(gt.offsets), gt.max_window,
                                float(min_confidence), max_cols, ptr(centers), ptr(widths), ptr(n_cols), ptr(ws),
                                max_bins, ptr(ws_spans), ptr(ws_span_counts), ptr(col_hist), stream_ptr(stream)))
    if return_ws:  # [P, 2, max_bins]: density map and smoothed density (debug / tests)
        return centers[:p], widths[:p], n_cols[:p], ws.view(-1, 2, max_bins)[:p]
    return centers[:p], widths[:p], n_cols[:p]


def assign_columns(boxes, page_off, centers, n_cols, sel_idx=None, n_sel=None, stream=None) -> torch.Tensor:
    """pg_assign_columns.  Returns col_of_box [N] i32: nearest column centre per selected box, -1 otherwise."""
    _require_cuda()
    boxes = _dev(boxes, torch.float64).view(-1, 4)
    page_off = _dev(page_off, torch.int64)
    centers = _dev(centers, torch.int32)
    n_cols = _dev(n_cols, torch.int32)
    n, p = boxes.shape[0], page_off.numel() - 1
    if sel_idx is not None:
        sel_idx = _dev(sel_idx, torch.int32)
        n_sel = _dev(n_sel, torch.int32)
    out = torch.empty(max(n, 1), dtype=torch.int32, device="cuda")
    check(lib().pg_assign_columns(ptr(boxes), ptr(sel_idx), ptr(page_off), ptr(n_sel), p, n, ptr(centers), ptr(n_cols),
                                  centers.shape[1] if centers.dim() == 2 else max(1, centers.numel() // max(p, 1)),
                                  ptr(out), stream_ptr(stream)))
    return out[:n]


# --------------------------------------------------------------------------------------------
# K6 corpus-histogram exchange
# --------------------------------------------------------------------------------------------
class CorpusComm:
    """The NCCL communicator of the K6 exchange, created through the C ABI (pg_comm_*): rank 0 makes the id,
    torch.distributed (whatever backend the process group has) carries its 128 bytes to the other ranks, and
    every rank joins.  One process per GPU; the current CUDA device is the rank's GPU."""

    def __init__(self):
        import torch.distributed as dist
        _require_cuda()
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        ident = (C.c_uint8 * 128)()
        if self.rank == 0:
            check(lib().pg_comm_unique_id(ident))
        box = [bytes(ident)]
        dist.broadcast_object_list(box, src=0)
        ident = (C.c_uint8 * 128).from_buffer_copy(box[0])
        handle = C.c_void_p()
        check(lib().pg_comm_create(ident, self.world, self.rank, C.byref(handle)))
        self._h = handle
        self.nccl = lib().pg_comm_nccl(self._h)

    def close(self):
        if getattr(self, "_h", None):
            lib().pg_comm_destroy(self._h)
            self._h = None


_COMM: Optional[CorpusComm] = None


def corpus_comm() -> CorpusComm:
    global _COMM
    if _COMM is None:
        _COMM = CorpusComm()
    return _COMM


def hist_allreduce(hist: torch.Tensor, stream=None) -> torch.Tensor:
    """pg_hist_allreduce: uint32 bins (int32 storage) summed over the ranks in place, asynchronous on `stream`."""
    _require_cuda()
    assert hist.is_cuda and hist.is_contiguous() and hist.element_size() == 4
    check(lib().pg_hist_allreduce(ptr(hist), hist.numel(), corpus_comm().nccl, stream_ptr(stream)))
    return hist


# --------------------------------------------------------------------------------------------
# J1-J4 stage-3 record writer
# --------------------------------------------------------------------------------------------
def combined_head_tail(image_path, image_size, iou_threshold: float, source_jsons) -> Tuple[bytes, bytes]:
    """The page-invariant text of a `<base>_combined.json` (3_combine_grids.py:282-291) exactly as
    json.dump(indent=2) prints it: everything before the first array, and everything after the last."""
    import json
    first = json.dumps({"image_path": image_path, "image_size": image_size,
                        "parameters": {"iou_threshold": iou_threshold}}, indent=2)
    head = first[:-2] + ',\n  "boxes": ['
    tail = json.dumps({"source_jsons": source_jsons}, indent=2)[1:]
    return head.encode("ascii"), tail.encode("ascii")


def json_combined(boxes, classes, scores, name_id, page_off, heads: Sequence[bytes], tails: Sequence[bytes],
                  names: Sequence[bytes], kept_idx=None, n_kept=None, max_boxes_per_page: Optional[int] = None,
                  stream=None) -> List[bytes]:
    """pg_json_combined.  Returns one bytes object per page: the text json.dump(result, f, indent=2) writes
    for the kept boxes of that page.  heads/tails: per-page text (combined_head_tail); names: JSON string
    literals (with quotes) indexed by name_id[box]."""
    _require_cuda()
    boxes = _dev(boxes, torch.float64).view(-1, 4)
    classes = _dev(classes, torch.float64)
    scores = _dev(scores, torch.float64)
    name_id = _dev(name_id, torch.int32)
    page_off_h = page_off.cpu().numpy() if isinstance(page_off, torch.Tensor) else np.asarray(page_off, np.int64)
    page_off = _dev(page_off, torch.int64)
    n, p = boxes.shape[0], page_off.numel() - 1
    assert len(heads) == p and len(tails) == p
    if kept_idx is not None:
        kept_idx = _dev(kept_idx, torch.int32)
        n_kept = _dev(n_kept, torch.int32)
    if max_boxes_per_page is None:
        max_boxes_per_page = int(np.diff(page_off_h).max()) if p else 0
    pieces = list(heads) + list(tails) + list(names)
    lens = np.fromiter((len(x) for x in pieces), np.int64, len(pieces))
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    text = torch.frombuffer(bytearray(b"".join(pieces) + b"\0"), dtype=torch.uint8).cuda()
    head_off = torch.from_numpy(offs[:p + 1].copy()).cuda()
    tail_off = torch.from_numpy(offs[p:2 * p + 1].copy()).cuda()
    name_off = torch.from_numpy(offs[2 * p:].copy()).cuda()
    ws_bytes = int(lib().pg_json_workspace_bytes(n, p))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    out_off = torch.zeros(p + 1, dtype=torch.int64, device="cuda")
    longest = int(max((len(x) for x in names), default=2))
    capacity = int(lens[:2 * p].sum()) + 80 * p + n * (6 * 20 + 70 + longest)  # typical; retried when short
    while True:
        out = torch.empty(max(capacity, 1), dtype=torch.uint8, device="cuda")
        check(lib().pg_json_combined(ptr(boxes), ptr(classes), ptr(scores), ptr(name_id), ptr(kept_idx), ptr(page_off),
                                     ptr(n_kept), p, n, int(max_boxes_per_page), ptr(text), ptr(head_off), ptr(tail_off),
                                     ptr(name_off), ptr(out), capacity, ptr(out_off), ptr(ws), ws_bytes,
                                     stream_ptr(stream)))
        if stream is not None:
            stream.synchronize()
        off = out_off.cpu().numpy()
        if int(off[-1]) <= capacity:
            break
        capacity = int(off[-1])
    host = out[:int(off[-1])].cpu().numpy().tobytes()
    return [host[off[i]:off[i + 1]] for i in range(p)]


# --------------------------------------------------------------------------------------------
# R1-R3 record reader
# --------------------------------------------------------------------------------------------
def json_parse_numbers(text, ranges, stream=None, int_literals_to_host: bool = False):
    """pg_json_parse_numbers.  text: bytes / uint8 array / cuda uint8 tensor; ranges: [R,2] begin/end byte
    offsets of number-only regions.  Returns (values f64 cuda [total], val_off int64 numpy [R+1],
    n_bad int32 numpy [R]): range r's numbers are values[val_off[r]:val_off[r+1]] in text order."""
    _require_cuda()
    if isinstance(text, (bytes, bytearray, memoryview)):
        text = torch.frombuffer(bytearray(text) + bytearray(1), dtype=torch.uint8)
    text = _dev(text, torch.uint8)
    rng = np.ascontiguousarray(np.asarray(ranges, np.int64).reshape(-1, 2))
    r = rng.shape[0]
    bb = int(lib().pg_json_parse_block_bytes())
    blocks = (np.maximum(rng[:, 1] - rng[:, 0], 0) + bb - 1) // bb
    blk_off = np.concatenate([[0], np.cumsum(blocks)]).astype(np.int64)
    total_blocks = int(blk_off[-1])
    d_rng, d_blk = torch.from_numpy(rng).cuda(), torch.from_numpy(blk_off).cuda()
    ws_bytes = int(lib().pg_json_parse_workspace_bytes(total_blocks))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    val_off = torch.zeros(r + 1, dtype=torch.int64, device="cuda")
    n_bad = torch.zeros(max(r, 1), dtype=torch.int32, device="cuda")
    capacity = int(np.maximum(rng[:, 1] - rng[:, 0], 0).sum()) // 4 + 16  # a number + separator is >= 4 bytes here
    while True:
        values = torch.empty(max(capacity, 1), dtype=torch.float64, device="cuda")
        check(lib().pg_json_parse_numbers(ptr(text), ptr(d_rng), r, ptr(d_blk), total_blocks, ptr(values), capacity,
                                          ptr(val_off), ptr(n_bad),
                                          _lib.PG_JSON_INT_LITERALS_TO_HOST if int_literals_to_host else 0,
                                          ptr(ws), ws_bytes, stream_ptr(stream)))
        if stream is not None:
            stream.synchronize()
        off = val_off.cpu().numpy()
        if int(off[-1]) <= capacity:
            break
        capacity = int(off[-1])
    return values[:int(off[-1])], off, n_bad[:r].cpu().numpy()


# --------------------------------------------------------------------------------------------
# Generic writer: any json.dump(indent=2) document with device-resident arrays (pg_json_segments)
# --------------------------------------------------------------------------------------------
SEGMENT_DTYPE = np.dtype([("head_begin", "<i8"), ("head_end", "<i8"), ("start", "<i8"), ("count", "<i4"),
                          ("kind", "<i4"), ("data_id", "<i4"), ("indent", "<i4")])
JSON_KIND_TEXT, JSON_KIND_BOX4, JSON_KIND_SCALAR, JSON_KIND_NAME = 0, 1, 2, 3


class DeviceArray:
    """Placeholder inside a document handed to render_documents: the JSON array printed at this place is
    elements [start, start + count) (positions into `kept_idx`, or box indices without one) of data[data_id]."""
    __slots__ = ("kind", "data_id", "start", "count")

    def __init__(self, kind: int, data_id: int, start: int, count: int):
        self.kind, self.data_id, self.start, self.count = int(kind), int(data_id), int(start), int(count)


def render_documents(docs: Sequence, data: Sequence, names: Sequence[bytes] = (), kept_idx=None, stream=None) -> List[bytes]:
    """`json.dumps(doc, indent=2)` for every doc, with each DeviceArray placeholder printed on the device from
    `data` (cuda tensors: f64 [N,4] for BOX4, f64 [N] for SCALAR, int32 [N] ids into `names` for NAME; names are
    JSON string literals).  The text around the arrays is encoded by CPython itself (a dump of the document with
    a sentinel string at each placeholder), so key order, escaping and nesting are whatever json.dumps does."""
    import json
    import re
    _require_cuda()
    marks: List[DeviceArray] = []

    def sentinel(o):
        if not isinstance(o, DeviceArray):
            raise TypeError(f"Object of type {type(o).__name__} is not JSON serializable")
        marks.append(o)
        return f"\x00PGARR{len(marks) - 1}\x00"

    pat = re.compile(rb'"\\u0000PGARR(\d+)\\u0000"')
    pieces, segs, doc_first = [], [], []
    for doc in docs:
        text = json.dumps(doc, indent=2, default=sentinel).encode("ascii")
        doc_first.append(len(segs))
        at = 0
        for m in pat.finditer(text):
            head = text[at:m.start()]
            nl = head.rfind(b"\n")
            line = head[nl + 1:]
            key_indent = len(line) - len(line.lstrip(b" "))
            a = marks[int(m.group(1))]
            segs.append((len(pieces), a.start, a.count, a.kind, a.data_id, key_indent + 2))
            pieces.append(head)
            at = m.end()
        segs.append((len(pieces), 0, 0, JSON_KIND_TEXT, 0, 0))
        pieces.append(text[at:])
    doc_first.append(len(segs))
    n_text = len(pieces)
    pieces.extend(names)
    lens = np.fromiter((len(x) for x in pieces), np.int64, len(pieces))
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    seg_arr = np.zeros(len(segs), SEGMENT_DTYPE)
    for i, (pi, start, count, kind, data_id, indent) in enumerate(segs):
        seg_arr[i] = (offs[pi], offs[pi + 1], start, count, kind, data_id, indent)
    elem_off = np.concatenate([[0], np.cumsum(seg_arr["count"].astype(np.int64))]).astype(np.int64)
    n_elems, n_segs = int(elem_off[-1]), len(segs)
    d_text = torch.frombuffer(bytearray(b"".join(pieces) + b"\0"), dtype=torch.uint8).cuda()
    d_segs = torch.from_numpy(seg_arr.view(np.uint8).reshape(-1)).cuda()
    d_elem = torch.from_numpy(elem_off).cuda()
    d_name = torch.from_numpy(offs[n_text:].copy()).cuda() if len(names) else None
    if kept_idx is not None:
        kept_idx = _dev(kept_idx, torch.int32)
    data = [t if isinstance(t, torch.Tensor) and t.is_cuda else _dev(t, torch.int32 if np.asarray(t).dtype.kind in "iu" else torch.float64)
            for t in data]
    ptrs = (C.c_void_p * max(len(data), 1))(*[t.data_ptr() for t in data])
    ws_bytes = int(lib().pg_json_segments_workspace_bytes(n_elems, n_segs))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    seg_off = torch.zeros(n_segs + 1, dtype=torch.int64, device="cuda")
    capacity = int(lens[:n_text].sum()) + 4 * n_segs + 40 * n_elems
    while True:
        out = torch.empty(max(capacity, 1), dtype=torch.uint8, device="cuda")
        check(lib().pg_json_segments(ptr(d_segs), n_segs, ptr(d_elem), n_elems, ptrs, len(data), ptr(kept_idx), ptr(d_text),
                                     ptr(d_name), ptr(out), capacity, ptr(seg_off), ptr(ws), ws_bytes, stream_ptr(stream)))
        if stream is not None:
            stream.synchronize()
        off = seg_off.cpu().numpy()
        if int(off[-1]) <= capacity:
            break
        capacity = int(off[-1])
    host = out[:int(off[-1])].cpu().numpy().tobytes()
    return [host[off[doc_first[i]]:off[doc_first[i + 1]]] for i in range(len(docs))]
