#!/usr/bin/env python
"""Drop-in for the reference's 3_combine_grids.py (same argv, same output schema); the work is done by
libpagegeom.so through multimodal_embeddings_b200.cli.main_stage3."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from multimodal_embeddings_b200.cli import main_stage3  # noqa: E402

if __name__ == "__main__":
    sys.exit(main_stage3())
