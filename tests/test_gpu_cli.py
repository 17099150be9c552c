"""Drop-in check at the CLI level: this repo's numbered stage mains run on the same synthetic
stage-1 tree the reference's own main()s were run on (oracle/gen_golden.py -> cli_tree.json.gz);
every JSON file they write must be identical — same files, same key order, same numbers."""
import json

import pytest

from conftest import load_golden
from multimodal_embeddings_b200 import cli
import cli_tree

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("sidecar", [False, True])
def test_cli_stages_2_to_5_match_reference_outputs(tmp_path, sidecar, monkeypatch):
    """sidecar=True: stage 3 also drops the binary `.pgrec` records and stages 4/5 must take their numbers
    from them (json.load of a stage-3 file is made to fail) — the JSON outputs stay identical."""
    golden = load_golden("cli_tree.json.gz")
    root = str(tmp_path)
    cli_tree.build_stage1_tree(root)
    argv = cli_tree.stage_argv(root)
    if sidecar:
        argv[3] = argv[3] + ["--sidecar"]
    for stage, main in ((2, cli.main_stage2), (3, cli.main_stage3)):
        assert main(argv[stage]) == 0
    if sidecar:
        import glob
        import os
        from multimodal_embeddings_b200 import records
        jsons = glob.glob(os.path.join(root, "3_combined_bboxes", "json", "*_combined.json"))
        assert jsons and all(os.path.exists(records.sidecar_path(j)) for j in jsons)
        real_load = json.load

        def guarded(f, *a, **k):
            assert "_combined.json" not in getattr(f, "name", ""), "stage-3 JSON text was parsed despite the sidecar"
            return real_load(f, *a, **k)
        monkeypatch.setattr(json, "load", guarded)
    for stage, main in ((4, cli.main_stage4), (5, cli.main_stage5)):
        assert main(argv[stage]) == 0
    monkeypatch.undo()
    got = cli_tree.collect_outputs(root)
    assert sorted(got) == sorted(golden)
    for name in golden:
        assert json.loads(got[name]) == json.loads(golden[name]), name
        assert got[name] == golden[name], f"{name}: key order / number formatting differs"


def test_cli_stage1_writes_reference_schema(tmp_path):
    import numpy as np
    from PIL import Image
    from multimodal_embeddings_b200 import synth
    from oracle import tiler as ot
    src = tmp_path / "in"
    src.mkdir()
    w, h = 1500, 1100
    Image.fromarray(synth.page_pixels(w, h, 9)[..., ::-1].copy()).save(src / "scan A.png")
    out = tmp_path / "out"
    assert cli.main_stage1(["--input_folder", str(src), "--output_folder", str(out), "--grids", "2x2,3x3",
                            "--detections", "synthetic", "--boxes_per_page", "300", "--imgsz", "512"]) == 0
    std = json.load(open(out / "json" / "scan A.json"))
    assert list(std) == ["image_path", "image_size", "parameters", "boxes", "classes", "scores", "class_names"]
    assert std["image_size"] == {"width": w, "height": h}
    for rows, cols in ((2, 2), (3, 3)):
        gi = json.load(open(out / "json" / f"scan A_grid_{rows}x{cols}.json"))
        assert list(gi) == ["original_image_path", "grid_config", "cells"]
        assert gi["grid_config"] == {"rows": rows, "cols": cols, "overlap_percentage": 20.0}
        ref_cells = ot.grid_cells(w, h, rows, cols, 20.0)
        assert len(gi["cells"]) == rows * cols
        for cell, ref in zip(gi["cells"], ref_cells):
            assert list(cell) == ["cell_path", "cell_json_path", "cell_coordinates", "row", "col", "regions"]
            assert cell["cell_coordinates"] == ref["coordinates"] and (cell["row"], cell["col"]) == (ref["row"], ref["col"])
            assert list(cell["regions"]) == ["boxes", "boxes_original", "classes", "scores", "class_names"]
            assert cell["regions"]["boxes_original"] == ot.translate_boxes(cell["regions"]["boxes"], ref["coordinates"])
            per_cell = json.load(open(cell["cell_json_path"]))
            assert per_cell["grid_info"] == {"rows": rows, "cols": cols, "row": ref["row"], "col": ref["col"]}
            assert per_cell["boxes_original"] == cell["regions"]["boxes_original"]
    # the whole chain runs on what stage 1 wrote
    for stage, main, argv in (
            (2, cli.main_stage2, ["--input_folder", str(out), "--output_folder", str(tmp_path / "s2")]),
            (3, cli.main_stage3, ["--input_folder", str(tmp_path / "s2"), "--output_folder", str(tmp_path / "s3")]),
            (4, cli.main_stage4, ["--input_folder", str(tmp_path / "s3" / "json"), "--output_folder", str(tmp_path / "s4")]),
            (5, cli.main_stage5, ["--input_folder", str(tmp_path / "s3" / "json"), "--median_folder",
                                  str(tmp_path / "s4" / "json"), "--output_folder", str(tmp_path / "s5")])):
        assert main(argv) == 0
    comb = json.load(open(tmp_path / "s3" / "json" / "scan A_combined.json"))
    assert list(comb) == ["image_path", "image_size", "parameters", "boxes", "classes", "scores", "class_names",
                          "source_jsons"]
    assert comb["scores"] == sorted(comb["scores"], reverse=True) and len(comb["boxes"]) > 10
    assert (tmp_path / "s4" / "json" / "scan A_combined_median_width.json").exists()


@pytest.mark.parametrize("world", [2, 3])
def test_cli_rank_sharding_union_equals_single_process(tmp_path, world, monkeypatch):
    """Stages 2-5 run once per rank (RANK / WORLD_SIZE as torchrun sets them; here sequentially on one GPU) into
    the same tree: every rank writes a disjoint subset and the union is the reference's output, file for file."""
    import os
    golden = load_golden("cli_tree.json.gz")
    root = str(tmp_path)
    cli_tree.build_stage1_tree(root)
    argv = cli_tree.stage_argv(root)
    monkeypatch.setenv("WORLD_SIZE", str(world))
    written = []
    for stage, main in ((2, cli.main_stage2), (3, cli.main_stage3), (4, cli.main_stage4), (5, cli.main_stage5)):
        seen = set()
        for rank in range(world):
            monkeypatch.setenv("RANK", str(rank))
            monkeypatch.setenv("LOCAL_RANK", "0")
            assert main(argv[stage]) == 0
            out_dir = argv[stage][argv[stage].index("--output_folder") + 1]
            now = {os.path.join(r, f) for r, _, fs in os.walk(out_dir) for f in fs if f.endswith(".json")}
            written.append((stage, rank, len(now - seen)))
            seen = now
    monkeypatch.undo()
    got = cli_tree.collect_outputs(root)
    assert sorted(got) == sorted(golden)
    for name in golden:
        assert got[name] == golden[name], name
    assert sum(n for s, _, n in written if s == 3) == len([k for k in golden if k.startswith("3_combined_bboxes/")])
    assert any(n > 0 for s, r, n in written if s == 3 and r > 0)  # the later ranks really did part of the work


def test_stage3_inputs_read_on_the_device_equal_json_load(tmp_path):
    """records.load_pool_inputs (arrays located by line-anchored keys, numbers converted on the GPU) against
    pool_documents (json.load) on the stage-2 files of the reference's golden tree; a file holding an integer
    literal among its numbers is left to CPython."""
    import os
    import numpy as np
    from multimodal_embeddings_b200 import records
    root = str(tmp_path)
    cli_tree.build_stage1_tree(root)
    argv = cli_tree.stage_argv(root)
    assert cli.main_stage2(argv[2]) == 0
    groups = cli.find_grid_jsons(os.path.join(root, "2_edge_box_filtered"))
    paths = [p for ps in groups.values() for p in ps]
    fast = records.load_pool_inputs(paths)
    assert sorted(fast) == sorted(paths)
    lg = cli._logger("GridBoxCombiner")
    for base, ps in groups.items():
        b, s, c, n, ip, isz = cli.pool_documents(ps, lg)
        fb, fs, fc, fn, fip, fisz, floats = cli.pool_documents_fast(ps, lg, fast)
        assert floats and isinstance(fb, np.ndarray) and fb.shape == (len(b), 4)
        assert fb.tolist() == b and fs.tolist() == s and fc.tolist() == c and fn == n and (fip, fisz) == (ip, isz)
    # an integer literal among the numbers: that file must not be taken by the device reader
    victim = paths[0]
    doc = json.load(open(victim))
    target = doc["cells"][0]["regions"] if "cells" in doc else doc
    key = "boxes_original" if "cells" in doc else "boxes"
    if target[key]:
        target[key][0][0] = 7
        json.dump(doc, open(victim, "w"), indent=2)
        again = records.load_pool_inputs(paths)
        assert victim not in again and len(again) == len(paths) - 1


def test_stage2_batch_path_equals_per_file_path(tmp_path):
    """records.filter_grid_files (read, edge filter and write on the device for all grid files at once) against
    the per-file path — reference_api.filter_grid_info + json.dumps — on the synthetic stage-1 tree; it must
    actually take every grid file, and decline one that holds an integer literal among its numbers."""
    import glob
    import os
    from multimodal_embeddings_b200 import records
    from multimodal_embeddings_b200 import reference_api as api
    root = str(tmp_path)
    s1 = cli_tree.build_stage1_tree(root)
    paths = sorted(glob.glob(os.path.join(s1, "json", "*.json")))
    grid = [p for p in paths if "_grid_" in os.path.basename(p)]
    size_of = lambda sk: cli._grid_page_size(sk, True)  # noqa: E731
    fast = records.filter_grid_files(paths, 10, size_of, api._cell_tuple)
    assert sorted(fast) == grid and grid
    both = records.filter_grid_files(paths, 10, size_of, api._cell_tuple, standard_ok=lambda doc: True)
    assert sorted(both) == paths and all(both[p] == fast[p] for p in grid)
    for p in paths:
        if p not in grid:  # full-page documents are copied (2:92-100): what json.dump of the loaded file gives
            assert both[p].decode("ascii") == json.dumps(api.filter_edge_boxes(json.load(open(p)), 10), indent=2)
    assert records.filter_grid_files(paths, 10, size_of, api._cell_tuple, standard_ok=lambda doc: False).keys() == fast.keys()
    n_removed = 0
    for p in grid:
        doc = json.load(open(p))
        want = api.filter_grid_info(doc, 10, image_size=size_of(doc))
        assert fast[p].decode("ascii") == json.dumps(want, indent=2)
        n_removed += sum(len(c["regions"]["boxes"]) for c in doc["cells"]) - sum(len(c["regions"]["boxes"]) for c in want["cells"])
    assert n_removed > 0  # the filter really dropped boxes
    victim = next(p for p in grid if any(c["regions"]["scores"] for c in json.load(open(p))["cells"]))
    doc = json.load(open(victim))
    cell = next(c for c in doc["cells"] if c["regions"]["scores"])
    cell["regions"]["boxes"][0][1] = 3
    json.dump(doc, open(victim, "w"), indent=2)
    again = records.filter_grid_files(paths, 10, size_of, api._cell_tuple)
    rest = [p for p in grid if p != victim]
    assert sorted(again) == rest and all(again[p] == fast[p] for p in rest)
