"""Development aid: host-buffer trace time for different chunk splits of the pipelined call (ATLAS_RT_PIPE_SPLIT)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from atlas_engine_b200 import capi, workloads as W
dev = torch.device("cuda", 0)
N = 1_000_000
tris = W.soup(N, seed=1234); boxes = W.tri_boxes(tris)
lo, hi = boxes[:, :3].min(0), boxes[:, 3:].max(0)
rays = W.random_rays(N, lo, hi, seed=5678)
root = np.concatenate([lo, hi])[None].astype(np.float32)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
h_in = torch.from_numpy(rays).pin_memory(); h_out = torch.empty_like(h_in).pin_memory()
ctx = capi.Context(0, stream.cuda_stream)
blas = ctx.build_blas(boxes, tris); tlas = ctx.build_tlas(root); mesh = ctx.pack_mesh(blas, tris)
scene = ctx.create_scene([mesh], W.identity_instance(), tlas)
splits = sys.argv[1:] or ["", "0.33,0.67", "0.2,0.6", "0.15,0.5,0.85", "0.1,0.4,0.7,0.9", "0.1,0.35,0.65,0.9", "0.2,0.5,0.8", "0.25,0.5,0.75", "0.12,0.36,0.62,0.86", "0.08,0.3,0.54,0.78,0.93"]
for sp in splits:
    if sp: os.environ["ATLAS_RT_PIPE_SPLIT"] = sp
    else: os.environ.pop("ATLAS_RT_PIPE_SPLIT", None)
    ts = []
    for i in range(12):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ctx.check(ctx.L.atlas_rt_trace_closest(ctx.h, scene.h, h_in.data_ptr(), N, capi.MASK_ALL, 0.0, capi.INF, h_out.data_ptr(), 0))
        ts.append((time.perf_counter() - t0) * 1e3)
    print(f"split={sp or 'default'} e2e_ms median {np.median(ts[3:]):.3f} min {min(ts[3:]):.3f}", flush=True)
