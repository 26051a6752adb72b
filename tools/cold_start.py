"""Where the cold start of a fresh process goes: CUDA context, library load, index build, first
and second search call (python tools/cold_start.py)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

t = {}
t0 = time.perf_counter()
import torch  # noqa: E402
t["import torch"] = time.perf_counter() - t0
t0 = time.perf_counter()
torch.cuda.init()
torch.zeros(1, device="cuda")
torch.cuda.synchronize()
t["cuda context (torch.zeros)"] = time.perf_counter() - t0
t0 = time.perf_counter()
from fandom_search_b200 import _native as nt  # noqa: E402
from fandom_search_b200.engine import DeviceIndex  # noqa: E402
nt.load()
t["load libfandom_search.so"] = time.perf_counter() - t0
rng = np.random.default_rng(0)
table = rng.standard_normal((50000, 300), dtype=np.float32)
script = rng.integers(0, 50000, 25000).astype(np.int32)
t0 = time.perf_counter()
idx = DeviceIndex(table, script)
t["index create (50k x 300 table, 25k-token script)"] = time.perf_counter() - t0
tok = rng.integers(0, 50000, 100000).astype(np.int32)
off = np.array([0, 100000], np.int64)
for name in ("first search (100k tokens)", "second search", "third search"):
    t0 = time.perf_counter()
    idx.search_host(tok, off)
    t[name] = time.perf_counter() - t0
big = rng.integers(0, 50000, 2500000).astype(np.int32)
offb = np.array([0, 2500000], np.int64)
for name in ("first 2.5M-token search", "second 2.5M-token search"):
    t0 = time.perf_counter()
    idx.search_host(big, offb)
    t[name] = time.perf_counter() - t0
print(json.dumps(t, indent=1))
