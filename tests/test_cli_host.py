"""CPU checks of the CLI host logic (file discovery, pooling order, median-file matching, argv)."""
import json
import os

import pytest

from multimodal_embeddings_b200 import cli
from oracle import boxes as ob
import cli_tree


def test_discovery_and_pooling_order(tmp_path):
    # NB: like the reference (3:168-170) the "_grid_" / "_combined" tests look at the FULL path
    # string, so this test must not run under a directory whose name contains them.
    root = str(tmp_path)
    s1 = cli_tree.build_stage1_tree(root)
    groups = cli.find_grid_jsons(s1)
    assert sorted(groups) == sorted(p[0] for p in cli_tree.PAGES)
    for base, paths in groups.items():
        assert [os.path.basename(p) for p in paths] == [f"{base}.json", f"{base}_grid_2x2.json"]  # standard first (3:175)
        docs = [json.load(open(p)) for p in paths]
        mine = cli.pool_documents(paths, cli._logger("GridBoxCombiner"))
        ref = ob.pool_boxes(docs)
        assert list(mine) == list(ref)
        assert mine[5] == docs[0]["image_size"]
    # '_combined' files are never re-pooled (3:170)
    open(os.path.join(s1, "json", "x_combined.json"), "w").write("{}")
    assert "x_combined" not in cli.find_grid_jsons(s1)


def test_find_matching_median_json_ladder(tmp_path):
    med = tmp_path / "med"
    med.mkdir()
    for n in ("a_page_0001_combined_median_width.json", "b_page_0002_median_width.json"):
        (med / n).write_text("{}")
    f = cli.find_matching_median_json
    assert f("/x/a_page_0001_combined.json", str(med)).endswith("a_page_0001_combined_median_width.json")
    assert f("/x/b_page_0002_grid_2x2.json", str(med)).endswith("b_page_0002_median_width.json")
    assert f("/x/zzz_0002_other.json", str(med)).endswith("b_page_0002_median_width.json")  # digit-part fallback
    assert f("/x/nomatch.json", str(med)) is None
    (med / "a_page_0001_combined_median_width.json").unlink()
    assert f("/x/nomatch.json", str(med)).endswith("b_page_0002_median_width.json")  # single file fallback


def test_stage_argv_surfaces_match_reference():
    # every flag of the reference CLIs is accepted (run.sh:60-70)
    for main, argv in (
        (cli.main_stage3, ["--input_folder", "/nonexistent_in", "--output_folder", "{out}", "--iou_threshold", "0.4",
                           "--viz_alpha", "0.5"]),
        (cli.main_stage4, ["--input_folder", "/nonexistent_in", "--output_folder", "{out}", "--min_margin_percent", "0.3"]),
        (cli.main_stage5, ["--input_folder", "/nonexistent_in", "--median_folder", "/nonexistent_m", "--output_folder",
                           "{out}", "--min_confidence", "0.25", "--verbose"]),
    ):
        import tempfile
        with tempfile.TemporaryDirectory() as out:
            try:
                rc = main([a.replace("{out}", out) for a in argv])
            except FileNotFoundError:
                rc = 0  # stage 3 lists the (missing) input folder like the reference does
            assert rc == 0
    with pytest.raises(SystemExit):
        cli.main_stage2(["--input_folder", "x"])  # --output_folder is required
    assert cli.get_image_paths("/nonexistent") == []


def test_record_sidecar_round_trip_and_staleness(tmp_path):
    """records.py: the binary sidecar of a stage-3 record gives back exactly what json.load gives for the
    JSON file (numbers bit for bit), and is ignored when stale, truncated or foreign."""
    import json
    import os
    import time

    import numpy as np

    from multimodal_embeddings_b200 import records
    rng = np.random.default_rng(3)
    n = 257
    boxes = rng.uniform(0, 8000, (n, 4)).astype(np.float32).astype(np.float64)
    classes = rng.integers(0, 3, n).astype(np.float64)
    scores = rng.random(n)
    names = ["title", "plain_text", 'fig "x" é']
    name_id = classes.astype(np.int32)
    doc = {"image_path": "/p/a b.png", "image_size": {"width": 8000, "height": 6000}, "parameters": {"iou_threshold": 0.5},
           "boxes": boxes.tolist(), "classes": classes.tolist(), "scores": scores.tolist(),
           "class_names": [names[i] for i in name_id], "source_jsons": ["x.json", "x_grid_2x2.json"]}
    jp = str(tmp_path / "a b_combined.json")
    with open(jp, "w") as f:
        json.dump(doc, f, indent=2)
    assert records.read_sidecar(jp) is None and records.load_record(jp) == doc
    sp = records.write_sidecar(jp, doc["image_path"], doc["image_size"], doc["parameters"], doc["source_jsons"],
                               boxes, classes, scores, names, name_id)
    rec = records.load_record(jp)
    assert isinstance(rec["boxes"], np.ndarray) and list(rec) == list(doc)
    as_lists = {k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in rec.items()}
    assert as_lists == doc
    # stale: the JSON was rewritten after the sidecar
    os.utime(sp, (time.time() - 100, time.time() - 100))
    assert records.read_sidecar(jp) is None and records.load_record(jp) == doc
    os.utime(sp, None)
    assert records.read_sidecar(jp) is not None
    # truncated / foreign bytes are refused, never mis-read
    raw = open(sp, "rb").read()
    for bad in (raw[:-4], b"NOTPGREC" + raw[8:], raw[:17] + b"\xff" + raw[18:], raw[:20] + b"}" + raw[21:]):
        with open(sp, "wb") as f:
            f.write(bad)
        assert records.read_sidecar(jp) is None


def test_split_record_text_locates_the_number_arrays():
    import json

    from multimodal_embeddings_b200 import records
    doc = {"image_path": '/x/\\n  "boxes": [ y.png', "image_size": {"width": 10, "height": 20}, "parameters": {"iou_threshold": 0.5},
           "boxes": [[1.5, 2.0, 3.25, 4.0], [5.0, 6.0, 7.0, 8.5]], "classes": [1.0, 0.0], "scores": [0.9, 0.8],
           "class_names": ["plain_text", 'ti"tle'], "source_jsons": ["a.json"]}
    raw = json.dumps(doc, indent=2).encode()
    head, tail, ranges = records.split_record_text(raw)
    assert head == {k: doc[k] for k in ("image_path", "image_size", "parameters")}
    assert tail == {k: doc[k] for k in ("class_names", "source_jsons")}
    got = [json.loads(b"[" + raw[a:b].rstrip().rstrip(b",")) for a, b in ranges]
    assert got == [doc["boxes"], doc["classes"], doc["scores"]]
    assert records.split_record_text(json.dumps(doc).encode()) is None           # compact layout
    assert records.split_record_text(json.dumps(doc, indent=4).encode()) is None  # another indent
    other = dict(doc)
    other["scores"], other["classes"] = other.pop("classes"), other.pop("scores")
    assert records.split_record_text(json.dumps({k: doc[k] for k in reversed(list(doc))}, indent=2).encode()) is None


def test_stage1_refuses_to_run_without_a_detector(tmp_path, caplog):
    """No silent substitute for the network: without --model_path / --detector / --detections the run is refused
    (exit status 2, nothing written); --model_path without the doclayout_yolo package is refused as loudly."""
    src = tmp_path / "in"
    src.mkdir()
    (src / "a.png").write_bytes(b"not an image")
    out = tmp_path / "out"
    assert cli.main_stage1(["--input_folder", str(src), "--output_folder", str(out)]) == 2
    assert not out.exists()
    weights = tmp_path / "w.pt"
    weights.write_bytes(b"x")
    assert cli.main_stage1(["--input_folder", str(src), "--output_folder", str(out), "--model_path", str(weights)]) == 2
    assert cli.main_stage1(["--input_folder", str(src), "--output_folder", str(out), "--model_path", "/nonexistent.pt"]) == 2
    assert cli.main_stage1(["--input_folder", str(src), "--output_folder", str(out), "--detections", "replay"]) == 2
    assert not out.exists()
