#!/usr/bin/env python
"""Builds an experimental copy of the library for A/B tuning sweeps:
   python tools/build_variant.py NAME -DBC_MAX_SPT=32 ...   ->  build/variants/librvmc_b200_NAME.so
Run it with RVMC_B200_LIB=<path> (rvmc_b200/_abi.py).  build/ is git-ignored but travels with gpurun."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rvmc_b200 import build as b  # noqa: E402

name, defines = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(ROOT, "build", "variants")
os.makedirs(out_dir, exist_ok=True)
b.OUT = os.path.join(out_dir, "librvmc_b200_%s.so" % name)
b.OBJ_DIR = os.path.join(out_dir, "obj_" + name)
print(b._build(True, False, defines))
