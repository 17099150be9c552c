"""Whole command-line chain against the unmodified reference: stage 1 (`main()` of 1_doclayout_bboxes.py with a
stub network, oracle/gen_golden_stage1.py) -> stage 2 `--process_grids` -> 3 -> 4 -> 5.  Every JSON file this
repository's command lines write must be TEXT-IDENTICAL to the reference's (same files, key order, mixed
int/float cell coordinates, float formatting), and the tile images written behind `--write_tiles` must decode
to the pixels of the reference's tile PNGs (1:424-430, 568).

Two runs of the same comparison: on the CPU with the oracle standing in for the kernels (host logic only,
tests/fake_device.py), and on the GPU through libpagegeom.so (`-m gpu`), where the letterboxed tiles are also
checked against cv2 on the reference's slices.
"""
import hashlib
import os

import numpy as np
import pytest

from conftest import load_golden
from multimodal_embeddings_b200 import cli
import stage1_chain
import stub_detector


def _run_chain(root: str):
    stage1_chain.write_pages(root)
    argv = stage1_chain.chain_argv(root)
    argv[1] = argv[1] + ["--detector", "stub_detector:StubPlugin", "--write_tiles"]
    for stage, main in ((1, cli.main_stage1), (2, cli.main_stage2), (3, cli.main_stage3), (4, cli.main_stage4),
                        (5, cli.main_stage5)):
        assert main(argv[stage]) == 0, f"stage {stage}"
    return stage1_chain.collect(root)


def _compare(got, golden):
    assert sorted(got["files"]) == sorted(golden["files"])
    for name in sorted(golden["files"]):
        assert got["files"][name] == golden["files"][name], f"{name}: text differs from the reference's file"
    assert got["tiles"] == golden["tiles"]  # same tile files, same shapes, same pixels
    # the golden really exercises what it is meant to pin
    per_cell_2 = [k for k in golden["files"] if k.startswith("2_edge_box_filtered" + os.sep + "grid_")]
    assert len(per_cell_2) == 39
    import json
    mixed = any(isinstance(v, int) for k in golden["files"] if "_grid_" in k and k.startswith("1_")
                for c in json.loads(golden["files"][k])["cells"] for v in c["cell_coordinates"].values())
    assert mixed  # int 0 / int W from the clamps of 1:418-421 among the float coordinates


def test_chain_host_logic_against_reference_golden(tmp_path, monkeypatch):
    import fake_device
    fake_device.install(monkeypatch)
    _compare(_run_chain(str(tmp_path)), load_golden("stage1_chain.json.gz"))


@pytest.mark.gpu
def test_chain_on_gpu_against_reference_golden(tmp_path):
    _compare(_run_chain(str(tmp_path)), load_golden("stage1_chain.json.gz"))


@pytest.mark.gpu
def test_letterboxed_tiles_of_the_golden_pages_equal_cv2_on_the_reference_slices(tmp_path):
    """The tensors the detector plug-in receives, against cv2 (resize INTER_LINEAR + 114 border, oracle/tiler.py)
    applied to the slices whose hashes the reference's own tile PNGs pin."""
    import cv2
    import torch
    from multimodal_embeddings_b200 import ops, reference_api
    from oracle import tiler as ot
    golden = load_golden("stage1_chain.json.gz")
    folder = stage1_chain.write_pages(str(tmp_path))
    grids = [(1, 1)] + reference_api.parse_grid_configs(stub_detector.GRIDS)
    pages = [cv2.imread(os.path.join(folder, name)) for name, *_ in stub_detector.PAGES]
    batch = ops.TileBatch([(p.shape[1], p.shape[0]) for p in pages], grids, stub_detector.OVERLAP, 1024)
    batch.bind(ops.upload_pages_pinned(pages))
    batch.run()
    torch.cuda.synchronize()
    checked = 0
    for pi, (name, *_ ) in enumerate(stub_detector.PAGES):
        base, ext = os.path.splitext(name)
        plan = batch.plan_of(pi)
        for t, ti in enumerate(plan.tiles):
            cell = pages[pi][ti["y0"]:ti["y1"], ti["x0"]:ti["x1"]]
            if (ti["grid_rows"], ti["grid_cols"]) != (1, 1):
                rel = os.path.join("1_doclayout_parsed", f"grid_{ti['grid_rows']}x{ti['grid_cols']}", "images",
                                   f"{base}_row{ti['row']}_col{ti['col']}{ext}")
                assert golden["tiles"][rel]["shape"] == list(cell.shape)
                assert golden["tiles"][rel]["sha256"] == hashlib.sha256(np.ascontiguousarray(cell).tobytes()).hexdigest()
            want = ot.letterbox_tile_cv2(cell, 1024)
            got = batch.tile_view(pi, t).cpu().numpy()
            assert got.shape == want.shape
            diff = np.abs(np.rint(got.astype(np.float32) * 255) - want.astype(np.float32))  # want: uint8 from cv2
            assert diff.max() <= 1 and (diff == 0).mean() > 0.999  # +-1 LSB bar of the north star; exact in practice
            checked += 1
    assert checked == 3 * (1 + 4 + 9)
