"""C2 device-resident closest-hit trace under different L2 access-policy windows (ATLAS_RT_L2_PERSIST_MB / _HIT_RATIO) and
trace knobs; one process, one context per setting. Usage (GPU box): python tools/l2_sweep.py"""
import os, sys, statistics
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from atlas_engine_b200 import capi, workloads as W
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
tris = W.soup(1_000_000, seed=1234)
boxes = W.tri_boxes(tris)
lo, hi = boxes[:, :3].min(0), boxes[:, 3:].max(0)
rays = W.random_rays(1_000_000, lo, hi, seed=5678)
root = np.concatenate([lo, hi])[None].astype(np.float32)
d_t, d_b = torch.from_numpy(tris).to(dev), torch.from_numpy(boxes).to(dev)
d_r = [torch.from_numpy(rays).to(dev) for _ in range(2)]
d_h = [torch.empty((len(rays), 4), dtype=torch.float32, device=dev) for _ in range(2)]
settings = [{}, {"ATLAS_RT_L2_PERSIST_MB": "64"}, {"ATLAS_RT_L2_PERSIST_MB": "96"}, {}]
if len(sys.argv) > 1 and sys.argv[1] == "full":
    settings = [{}, {"ATLAS_RT_L2_PERSIST_MB": "32"}, {"ATLAS_RT_L2_PERSIST_MB": "64"}, {"ATLAS_RT_L2_PERSIST_MB": "80"},
                {"ATLAS_RT_L2_PERSIST_MB": "64", "ATLAS_RT_L2_HIT_RATIO": "0.5"}, {"ATLAS_RT_TRACE_REFILL_THRESHOLD": "12"},
                {"ATLAS_RT_TRACE_REFILL_THRESHOLD": "8"}, {"ATLAS_RT_TRACE_LEAF_THRESHOLD": "6"}, {"ATLAS_RT_TRACE_LEAF_THRESHOLD": "12"}, {}]
ref = None
for env in settings:
    for k in ("ATLAS_RT_L2_PERSIST_MB", "ATLAS_RT_L2_HIT_RATIO", "ATLAS_RT_TRACE_REFILL_THRESHOLD", "ATLAS_RT_TRACE_LEAF_THRESHOLD"):
        os.environ.pop(k, None)
    os.environ.update(env)
    ctx = capi.Context(0, stream.cuda_stream)
    scenes = []
    for _ in range(2):
        blas = ctx.build_blas(d_b, d_t, len(tris))
        tlas = ctx.build_tlas(root)
        mesh = ctx.pack_mesh(blas, d_t, len(tris))
        scenes.append((ctx.create_scene([mesh], W.identity_instance(), tlas), blas, tlas, mesh))
    for k in range(40):
        ctx.trace(scenes[k & 1][0], d_r[k & 1], len(rays), out=d_h[k & 1], flags=capi.ASYNC | capi.HITS_ONLY)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    K = 40
    for k in range(K):
        ctx.trace(scenes[k & 1][0], d_r[k & 1], len(rays), out=d_h[k & 1], flags=capi.ASYNC | capi.HITS_ONLY)
    b.record(stream)
    torch.cuda.synchronize()
    got = d_h[0].cpu().numpy()
    if ref is None:
        ref = got
    print(f"{a.elapsed_time(b) / K:.4f} ms/step  equal={bool(np.array_equal(got.view(np.uint32), ref.view(np.uint32)))}  {env}", flush=True)
    for sc, blas, tlas, mesh in scenes:
        for o in (sc, mesh, tlas, blas):
            o.free()
    ctx.close()
