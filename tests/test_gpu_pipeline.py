"""Whole hot path chained on the device (PagePipeline, the object bench.py times) against the oracle,
in both launch orders (two streams / one stream), plus the corpus histograms and their running exchange."""
import json

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from multimodal_embeddings_b200 import ops, synth  # noqa: E402
from multimodal_embeddings_b200._lib import PG_COL_HIST_BINS, PG_WIDTH_HIST_BINS  # noqa: E402
from multimodal_embeddings_b200.pipeline import PagePipeline, corpus_median_width  # noqa: E402
from oracle import boxes as ob  # noqa: E402
from oracle import tiler as ot  # noqa: E402
from oracle.nms_fast import nms_pick_order_c  # noqa: E402

pytestmark = pytest.mark.gpu

W, H, ROWS, COLS, NB, IMGSZ = 1800, 1300, 3, 2, 700, 512


@pytest.fixture(scope="module")
def workload():
    plan = ops.TilePlan(W, H, [(ROWS, COLS)], 20.0, imgsz=IMGSZ)
    imgs = [synth.page_pixels(W, H, seed=s) for s in (11, 12, 13)]
    dets = [synth.page_detections(W, H, ROWS, COLS, 20.0, NB, synth.PAGE_SEED0 + s) for s in (11, 12, 13)]
    return plan, imgs, dets, ops.upload_pages(imgs, plan)


def oracle_page(d):
    bp = d["boxes_local"] + d["cells"][d["box_cell"]][:, [0, 1, 0, 1]]
    keep = np.asarray([j for j in range(len(bp)) if not ob.touches_internal_edge(bp[j], d["cells"][d["box_cell"][j]], W, H, 10)])
    final = keep[nms_pick_order_c(bp[keep], d["scores"][keep], d["classes"][keep], 0.5)]
    names = synth.class_names_of(d["classes"][final])
    med, nb = ob.median_width(bp[final].tolist(), names, W, 0.2)
    cc, cw = ob.column_centers(bp[final].tolist(), names, d["scores"][final].tolist(), W, H, med, 0.3)
    widths = (bp[final][:, 2] - bp[final][:, 0])[np.asarray(names) == "plain_text"]
    return final, float(med), nb, [float(x) for x in cc], [float(x) for x in cw], widths


@pytest.mark.parametrize("overlap", [True, False])
def test_graph_replay_equals_eager_step(workload, overlap):
    """The step captured into a CUDA graph (fork/join over both streams) and replayed gives the eager results."""
    plan, imgs, dets, pages = workload
    pipe = PagePipeline(plan, len(imgs), overlap=overlap)
    pipe.set_detections(dets)
    pipe.run(pages)
    torch.cuda.synchronize()
    want = {k: v.clone() for k, v in pipe.results_to_host().items()}
    tiles = pipe.tiles_out.clone()
    graph = pipe.capture(pages)
    for name in ("kept2", "n_kept2", "median", "centers", "n_cols"):
        getattr(pipe, name).zero_()
    pipe.tiles_out.zero_()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    pipe.check_status()
    got = pipe.results_to_host()
    for k in want:
        if k == "kept2":  # entries past n_kept of a page are scratch
            for p in range(len(imgs)):
                n = int(want["n_kept2"][p])
                assert torch.equal(got[k][p * NB: p * NB + n], want[k][p * NB: p * NB + n])
        else:
            assert torch.equal(got[k], want[k]), k
    assert torch.equal(pipe.tiles_out, tiles)


@pytest.mark.parametrize("overlap", [True, False])
def test_pipeline_matches_oracle_and_accumulates_corpus_histograms(workload, overlap):
    plan, imgs, dets, pages = workload
    pipe = PagePipeline(plan, len(imgs), corpus_stats=True, overlap=overlap)
    pipe.set_detections(dets)
    # stage-3 records laid out on the device in the same step (deliberately small buffer: forces the re-run path)
    all_classes = np.concatenate([d["classes"] for d in dets])
    names = sorted(set(synth.class_names_of(all_classes)))
    name_id = np.asarray([names.index(nm) for nm in synth.class_names_of(all_classes)], np.int32)
    meta = [(f"/scans/p{i}.png", {"width": W, "height": H}, [f"/scans/2/p{i}_grid_{ROWS}x{COLS}.json"]) for i in range(len(imgs))]
    ht = [ops.combined_head_tail(path, size, 0.5, srcs) for path, size, srcs in meta]
    pipe.enable_records([h for h, _ in ht], [t for _, t in ht], [json.dumps(nm).encode("ascii") for nm in names],
                        name_id, bytes_per_box=12 if overlap else 260)
    for _ in range(2):  # the second pass re-uses every buffer and the self-resetting work counters
        pipe.run(pages)
        pipe.exchange_corpus_stats_async()
    total = pipe.finish_exchange()
    torch.cuda.synchronize()
    pipe.check_status()
    res = pipe.results_to_host()
    ref_w = np.zeros(PG_WIDTH_HIST_BINS, np.int64)
    ref_c = np.zeros(PG_COL_HIST_BINS, np.int64)
    for p, (img, d) in enumerate(zip(imgs, dets)):
        for t, cell in enumerate(ot.split_array_into_grid(img, ROWS, COLS, 20.0)):
            want = ot.u8_to_f16_unit(ot.letterbox_tile_cv2(cell["image"], IMGSZ))
            assert np.array_equal(plan.tile_view(pipe.tiles_out, p, t).cpu().numpy(), want), (p, t)
        final, med, nb, cc, cw, widths = oracle_page(d)
        k = int(res["n_kept2"][p])
        assert k == len(final) and np.array_equal(res["kept2"][p * NB: p * NB + k].numpy() - p * NB, final)
        assert float(res["median"][p]) == med and int(res["n_bins"][p]) == nb
        nc = int(res["n_cols"][p])
        assert [float(x) for x in res["centers"][p, :nc]] == cc
        assert [float(x) for x in res["col_widths"][p, :nc]] == cw
        ref_w += np.bincount(np.clip(widths.astype(np.int64), 0, PG_WIDTH_HIST_BINS - 1), minlength=PG_WIDTH_HIST_BINS)
        for c in cc:
            ref_c[min(PG_COL_HIST_BINS - 1, int(c) * 1000 // W)] += 1
    docs = pipe.records_to_host()
    for p, (d, (path, size, srcs)) in enumerate(zip(dets, meta)):
        final = oracle_page(d)[0]
        bp = d["boxes_local"] + d["cells"][d["box_cell"]][:, [0, 1, 0, 1]]
        want = json.dumps({"image_path": path, "image_size": size, "parameters": {"iou_threshold": 0.5},
                           "boxes": bp[final].tolist(), "classes": d["classes"][final].tolist(),
                           "scores": d["scores"][final].tolist(), "class_names": synth.class_names_of(d["classes"][final]),
                           "source_jsons": srcs}, indent=2)
        assert docs[p].decode("ascii") == want
    hist = pipe.hist.cpu().numpy()
    assert np.array_equal(hist[:PG_WIDTH_HIST_BINS], 2 * ref_w) and np.array_equal(hist[PG_WIDTH_HIST_BINS:], 2 * ref_c)
    assert np.array_equal(total.cpu().numpy(), hist)  # one rank: the running exchange is the local total
    if ref_w.sum():
        srt = np.sort(np.repeat(np.arange(PG_WIDTH_HIST_BINS), 2 * ref_w))
        assert corpus_median_width(pipe.width_hist) == (srt[(len(srt) - 1) // 2] + srt[len(srt) // 2]) / 2.0


def test_scan_pipeline_on_jpeg_files_equals_page_pipeline_on_cv2_pages():
    """pipeline.ScanPipeline (files + detections in host memory -> H2D -> device decode -> one-channel tiler -> box
    stages -> D2H, double-buffered) against PagePipeline fed with what cv2.imdecode returns for the same files: same
    tiles, same kept indices, medians and columns — for several steps in flight with different detections per step,
    and with the decoder deliberately configured with too few sync rounds (the step is redone by results())."""
    import cv2
    from multimodal_embeddings_b200.pipeline import ScanPipeline
    w, h, rows, cols, n_pages = 1504, 1000, 2, 2, 3
    rng = np.random.default_rng(5)
    noisy = np.clip(200 + rng.normal(0, 40, (h, w)), 0, 255).astype(np.uint8)  # long blocks: needs many sync rounds
    greys = [synth.newspaper_page(w, h, 31), noisy, synth.newspaper_page(w, h, 33)]
    files = [cv2.imencode(".jpg", g, [cv2.IMWRITE_JPEG_QUALITY, 95])[1].tobytes() for g in greys]
    blob, off = ops.pack_files(files)
    sp = ScanPipeline(w, h, n_pages, [(rows, cols)], 20.0)
    for sl in sp.slots:
        sl["dec"].auto_chunk, sl["dec"].chunk_bytes, sl["dec"].sync_rounds = False, 256, 1
    plan3 = ops.TilePlan(w, h, [(rows, cols)], 20.0)
    bgr = [cv2.imdecode(np.frombuffer(f, np.uint8), cv2.IMREAD_COLOR) for f in files]
    pages3 = ops.upload_pages(bgr, plan3)
    steps = []
    for s_i in range(4):
        dets = [synth.page_detections(w, h, rows, cols, 20.0, 700 + 50 * s_i, 900 + 10 * s_i + p) for p in range(n_pages)]
        probe = PagePipeline(plan3, n_pages)
        host = probe.set_detections(dets)
        probe.run(pages3)
        torch.cuda.synchronize()
        probe.check_status()
        steps.append((host, {k: v.clone() for k, v in probe.results_to_host().items()}, probe.tiles_out.clone()))
    ids = [sp.submit(blob, off, host) for host, _, _ in steps[:2]]
    for s_i in range(4):
        got = sp.results(ids[s_i])
        want = steps[s_i][1]
        for k in ("n_kept2", "median", "n_bins", "n_cols"):
            assert np.array_equal(got[k], want[k].numpy()), (s_i, k)
        offp = steps[s_i][0]["page_off"]
        for p in range(n_pages):
            nk = int(want["n_kept2"][p])
            assert np.array_equal(got["kept2"][offp[p]:offp[p] + nk], want["kept2"][offp[p]:offp[p] + nk].numpy())
            nc = int(want["n_cols"][p])
            assert np.array_equal(got["centers"][p, :nc], want["centers"][p, :nc].numpy())
            assert np.array_equal(got["col_widths"][p, :nc], want["col_widths"][p, :nc].numpy())
        sl = sp.slots[ids[s_i] % sp.depth]
        assert torch.equal(sl["pipe"].tiles_out, steps[s_i][2])  # tiles from the decoded grey planes == tiles from cv2's BGR pages
        if s_i + 2 < 4:
            ids.append(sp.submit(blob, off, steps[s_i + 2][0]))
    assert all(sl["dec"].sync_rounds > 1 for sl in sp.slots)  # the unconverged decodes were noticed and redone
