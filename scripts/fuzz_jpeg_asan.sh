#!/bin/bash
# 3000 damaged JPEG files through the parser and the host evaluation of the decoder under AddressSanitizer (CPU only).
# Builds _variants/libpagegeom_asan.so (host code instrumented) and preloads libasan into python.
set -e
cd "$(dirname "$0")/.."
python - <<'P'
from multimodal_embeddings_b200 import build
print(build.build_variant("asan", ["-Xcompiler", "-fsanitize=address", "-Xcompiler", "-fno-omit-frame-pointer", "-g"]))
P
ASAN_OPTIONS=detect_leaks=0:protect_shadow_gap=0 LD_PRELOAD=$(gcc -print-file-name=libasan.so) \
PAGEGEOM_LIB=$PWD/multimodal_embeddings_b200/_variants/libpagegeom_asan.so PYTHONPATH=$PWD:$PWD/tests python - <<'P'
import ctypes as C
import numpy as np
import jpeg_cases
from multimodal_embeddings_b200._lib import lib
L = lib()
codes = {}
for data in jpeg_cases.mutated_files(7, 3000):
    w, h, st = C.c_int32(), C.c_int32(), (C.c_int64 * 4)()
    b = np.frombuffer(data, np.uint8)
    rc = L.pg_hostcheck_jpeg_decode(b.ctypes.data, len(b), 256, 64, None, 0, C.byref(w), C.byref(h), st)
    if rc == 0 and w.value * h.value <= 4_000_000:
        pitch = (st[3] * w.value + 15) // 16 * 16
        out = np.zeros((h.value, pitch), np.uint8)
        rc = L.pg_hostcheck_jpeg_decode(b.ctypes.data, len(b), 256, 64, out.ctypes.data, pitch, C.byref(w), C.byref(h), st)
    codes[rc] = codes.get(rc, 0) + 1
print("3000 damaged files under AddressSanitizer, return codes:", codes)
P
# the host-side C-ABI suites (tile plans, IoU / edge predicates, Ryu formatter, number reader, JPEG host evaluation) under the same build
ASAN_OPTIONS=detect_leaks=0:protect_shadow_gap=0 LD_PRELOAD=$(gcc -print-file-name=libasan.so) \
PAGEGEOM_LIB=$PWD/multimodal_embeddings_b200/_variants/libpagegeom_asan.so python -m pytest tests/test_cabi_host.py tests/test_jpeg_host.py -x -q
