"""CPU oracle for the page-geometry path (tile -> filter -> merge -> columns).

TEST INFRASTRUCTURE ONLY.  Nothing under ``multimodal_embeddings_b200/`` imports this
package; the only callers are ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.

Every function is a from-scratch restatement (numpy / plain Python / C) of the
reference's algorithm and cites the reference ``file:line`` it follows
(paths relative to the upstream repo root).

Pinning status (see tests/test_oracle_golden.py, oracle/gen_golden.py):
  * stages 2-5 and the stage-1 grid geometry are pinned against golden vectors
    produced by importing the *unmodified* reference scripts in the build
    container (tests/golden/*.json, generator committed), and against the
    reference's own committed 3_combined_bboxes/json outputs (NMS idempotence,
    stage 4/5 known answers).
  * the letterbox/resize front half lives in third-party code that is absent
    offline (doclayout-yolo / ultralytics, unpinned in requirements.txt:8,14);
    the oracle restates ultralytics' published LetterBox arithmetic and is
    pinned against the in-container ``cv2.resize(INTER_LINEAR)`` /
    ``cv2.copyMakeBorder`` primitives the reference ends up calling
    ("parity pinned to cv2 4.13, letterbox geometry unpinned").
"""
