"""Development aid: host-buffer trace time for combinations of pipeline chunks and compute streams."""
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
for streams in (8,):
    os.environ["ATLAS_RT_PIPE_STREAMS"] = str(streams)
    ctx = capi.Context(0, stream.cuda_stream)
    blas = ctx.build_blas(boxes, tris); tlas = ctx.build_tlas(root); mesh = ctx.pack_mesh(blas, tris)
    scene = ctx.create_scene([mesh], W.identity_instance(), tlas)
    for chunks in (6, 8, 10, 12):
        if chunks == "default": os.environ.pop("ATLAS_RT_PIPE_CHUNKS", None)
        else: os.environ["ATLAS_RT_PIPE_CHUNKS"] = str(chunks)
        ts = []
        for i in range(12):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            ctx.check(ctx.L.atlas_rt_trace_closest(ctx.h, scene.h, h_in.data_ptr(), N, capi.MASK_ALL, 0.0, capi.INF, h_out.data_ptr(), 0))
            ts.append((time.perf_counter() - t0) * 1e3)
        print(f"streams={streams} chunks={chunks} e2e_ms median {np.median(ts[3:]):.3f} min {min(ts[3:]):.3f}", flush=True)
    for o in (scene, mesh, tlas, blas): o.free()
    ctx.close()
