#!/usr/bin/env python
"""Host -> device bandwidth from pinned memory on this box: one contiguous copy (torch) and
the frame-strided copies the library issues (contiguous frames and frames with a gap)."""
import json
import sys
import time
sys.path.insert(0, ".")
import numpy as np
import torch
from mdhelper_b200 import _lib
from mdhelper_b200.universe import pinned_empty

N, F = 500_000, 64
host, keep = pinned_empty((F, N + 1000, 3))
host[:] = 1.0
dev = torch.empty((F, N, 3), dtype=torch.float32, device="cuda")
out = {}
# torch: one contiguous copy of the same number of bytes
h2, k2 = pinned_empty((F, N, 3))
for _ in range(2):
    dev.copy_(torch.from_numpy(h2), non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    dev.copy_(torch.from_numpy(h2), non_blocking=True)
torch.cuda.synchronize()
out["torch_contiguous_gbs"] = 5 * h2.nbytes / (time.perf_counter() - t0) / 1e9
print(json.dumps(out))
