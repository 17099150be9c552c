"""N>1 host logic on CPU: page sharding and the corpus-histogram all-reduce over gloo, world_size 2.

Pages are independent, so the per-page path has no collective; the only exchange step is the
integer histogram sum (NCCL on GPUs).  Here two gloo ranks build the histograms of their page
shards with numpy and the all-reduced result must equal the single-process histogram bit-for-bit."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multimodal_embeddings_b200 import synth
from multimodal_embeddings_b200._lib import PG_COL_HIST_BINS, PG_WIDTH_HIST_BINS
from multimodal_embeddings_b200.pipeline import allreduce_histograms, corpus_median_width, shard_pages

N_PAGES = 7  # deliberately not divisible by the world size


def _page_hist(page_idx: int) -> np.ndarray:
    d = synth.page_detections(3801, 5601, 2, 2, 20.0, 300, synth.PAGE_SEED0 + page_idx)
    w = d["boxes_local"][:, 2] - d["boxes_local"][:, 0]
    w = w[d["classes"] == 1.0]
    h = np.zeros(PG_WIDTH_HIST_BINS + PG_COL_HIST_BINS, np.int32)
    np.add.at(h, np.clip(w.astype(np.int64), 0, PG_WIDTH_HIST_BINS - 1), 1)
    h[PG_WIDTH_HIST_BINS + (page_idx * 37) % PG_COL_HIST_BINS] += 1
    return h


def _worker(rank: int, world: int, port: int, out_dir: str):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pages = shard_pages(N_PAGES, rank, world)
    h = np.zeros(PG_WIDTH_HIST_BINS + PG_COL_HIST_BINS, np.int32)
    for p in pages:
        h += _page_hist(p)
    t = torch.from_numpy(h)
    allreduce_histograms(t)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), t.numpy())
    np.save(os.path.join(out_dir, f"pages{rank}.npy"), np.asarray(list(pages)))
    dist.destroy_process_group()


def test_shard_pages_partitions_exactly():
    for n in (0, 1, 7, 64, 100000):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                seen.extend(shard_pages(n, r, world))
            assert seen == list(range(n))


def test_histogram_allreduce_world2_gloo(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    expect = sum(_page_hist(p) for p in range(N_PAGES))
    r0 = np.load(tmp_path / "rank0.npy")
    r1 = np.load(tmp_path / "rank1.npy")
    assert np.array_equal(r0, expect) and np.array_equal(r1, expect)
    pages = sorted(np.load(tmp_path / "pages0.npy").tolist() + np.load(tmp_path / "pages1.npy").tolist())
    assert pages == list(range(N_PAGES))
    # corpus statistic derived identically on every rank
    med = corpus_median_width(torch.from_numpy(r0[:PG_WIDTH_HIST_BINS]))
    allw = np.repeat(np.arange(PG_WIDTH_HIST_BINS), expect[:PG_WIDTH_HIST_BINS])
    assert med == float(np.median(allw))
    # world size 1: the all-reduce is a no-op
    t = torch.from_numpy(expect.copy())
    assert np.array_equal(allreduce_histograms(t).numpy(), expect)
