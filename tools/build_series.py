"""Development aid: per-build device time of 40 consecutive 1M-triangle BLAS builds in a fresh process."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from atlas_engine_b200 import capi, workloads as W
N = 1_000_000
tris = W.soup(N, seed=1234); boxes = W.tri_boxes(tris)
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
ctx = capi.Context(0, stream.cuda_stream)
db, dt = torch.from_numpy(boxes).to(dev), torch.from_numpy(tris).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts, hs = [], []
prev = None
for i in range(40):
    if "--noflush" not in sys.argv: flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if prev is not None: prev.free()
    t0 = time.perf_counter()
    a.record(stream); prev = ctx.build_blas(db, dt, N, flags=capi.ASYNC); b.record(stream)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b)); hs.append((t1 - t0) * 1e3)
print("device ms:", " ".join(f"{t:.2f}" for t in ts))
print("host enqueue ms:", " ".join(f"{t:.2f}" for t in hs))
