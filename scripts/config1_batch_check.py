"""BASELINE config 1 through bench.config1_block (one stream, two streams, the batched entry point) + an ncu-friendly
single batched call.   python scripts/config1_batch_check.py [N ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manifoldem_python_b200 import _lib   # noqa: E402
import bench                              # noqa: E402

lib = _lib.load()
ctx = _lib.Context(0)
for N in [int(a) for a in sys.argv[1:]] or [256, 128]:
    out = bench.config1_block(ctx, lib, _lib, 0, N)
    print(json.dumps(out))
