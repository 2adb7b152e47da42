"""Quick check of the streaming host-buffer trace (run under `timeout`): results equal the device-resident path, timing."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from atlas_engine_b200 import capi, workloads as W
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ctx = capi.Context(0, stream.cuda_stream)
n_tris = int(os.environ.get("DIAG_TRIS", "1000000"))
tris = W.soup(n_tris, seed=1234)
boxes = W.tri_boxes(tris)
blas = ctx.build_blas(boxes, tris)
lo, hi = boxes[:, :3].min(0), boxes[:, 3:].max(0)
tlas = ctx.build_tlas(np.concatenate([lo, hi])[None].astype(np.float32))
mesh = ctx.pack_mesh(blas, tris)
scene = ctx.create_scene([mesh], W.identity_instance(), tlas)
n = int(os.environ.get("DIAG_RAYS", "1000000"))
rays = W.random_rays(n, lo, hi, seed=5678)
d_r = torch.from_numpy(rays).to(dev)
d_h = torch.empty((n, 4), dtype=torch.float32, device=dev)
ctx.trace(scene, d_r, n, out=d_h, flags=capi.HITS_ONLY)
ref = d_h.cpu().numpy()
h_r = torch.from_numpy(rays).pin_memory()
h_h = torch.empty((n, 4), dtype=torch.float32).pin_memory()
for mode, flags in (("hits_only", capi.HITS_ONLY), ("rays48", 0), ("device_out", capi.HITS_ONLY | capi.DEVICE_OUTPUT)):
    out = h_h if mode == "hits_only" else (torch.empty((n, 12), dtype=torch.float32).pin_memory() if mode == "rays48" else d_h)
    ts = []
    for k in range(8):
        t0 = time.perf_counter()
        ctx.check(ctx.L.atlas_rt_trace_closest(ctx.h, scene.h, h_r.data_ptr(), n, capi.MASK_ALL, 0.0, capi.INF, out.data_ptr(), flags))
        ts.append((time.perf_counter() - t0) * 1e3)
    got = out.cpu().numpy() if mode == "device_out" else out.numpy()
    got = got[:, 8:12] if mode == "rays48" else got
    print(mode, "ms", " ".join(f"{t:.3f}" for t in ts), "equal", bool(np.array_equal(got.view(np.uint32), ref.view(np.uint32))), flush=True)
pg = ctx.trace(scene, rays, flags=capi.HITS_ONLY)      # pageable host memory
print("pageable equal", bool(np.array_equal(pg.view(np.uint32), ref.view(np.uint32))), flush=True)
