"""Development aid: time the C2 trace for a grid of scheduling knobs (fresh context per setting)."""
import os, sys, itertools
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from atlas_engine_b200 import capi, workloads as W
N = 1_000_000
tris = W.soup(N, seed=1234); boxes = W.tri_boxes(tris)
lo, hi = boxes[:, :3].min(0), boxes[:, 3:].max(0)
rays = W.random_rays(N, lo, hi, seed=5678)
root = np.concatenate([lo, hi])[None].astype(np.float32)
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
d_rays = torch.from_numpy(rays).to(dev); d_out = torch.empty_like(d_rays)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
grid = [(8, 12, 9, 1), (8, 10, 9, 1), (8, 8, 9, 1), (8, 14, 9, 1), (6, 12, 9, 1), (6, 10, 9, 1), (7, 12, 9, 1), (5, 12, 9, 1), (8, 12, 9, 1), (8, 16, 9, 1)]
for (l, r, b, p) in grid:
    os.environ["ATLAS_RT_TRACE_LEAF_THRESHOLD"] = str(l); os.environ["ATLAS_RT_TRACE_REFILL_THRESHOLD"] = str(r); os.environ["ATLAS_RT_TRACE_BLOCKS_PER_SM"] = str(b); os.environ["ATLAS_RT_TRACE_LONGEST_FIRST"] = str(p)
    ctx = capi.Context(0, stream.cuda_stream)
    blas = ctx.build_blas(boxes, tris); tlas = ctx.build_tlas(root); mesh = ctx.pack_mesh(blas, tris)
    scene = ctx.create_scene([mesh], W.identity_instance(), tlas)
    ts = []
    for i in range(14):
        flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); ctx.trace(scene, d_rays, N, out=d_out, flags=capi.ASYNC); e.record(stream)
        torch.cuda.synchronize(); ts.append(a.elapsed_time(e))
    print(f"leaf={l:2d} refill={r:2d} blocks={b} longest_first={p} ms={np.median(ts[4:]):.4f}", flush=True)
    for o in (scene, mesh, tlas, blas): o.free()
    ctx.close()
