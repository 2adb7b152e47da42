"""Development aid: e2e (host buffers) trace time for pipeline chunk counts / rays-per-warp, plus device-resident trace time for small batches."""
import os, sys, time
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
h_in = torch.from_numpy(rays).pin_memory(); h_out = torch.empty_like(h_in).pin_memory()
d_rays = torch.from_numpy(rays).to(dev); d_out = torch.empty_like(d_rays)
for rpw in (0, 1):
    os.environ["ATLAS_RT_TRACE_LONGEST_FIRST"] = str(rpw)
    for chunks in (1, 2, 4):
        os.environ["ATLAS_RT_PIPE_CHUNKS"] = str(chunks)
        ctx = capi.Context(0, stream.cuda_stream)
        blas = ctx.build_blas(boxes, tris); tlas = ctx.build_tlas(root); mesh = ctx.pack_mesh(blas, tris)
        scene = ctx.create_scene([mesh], W.identity_instance(), tlas)
        ts = []
        for i in range(7):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            ctx.check(ctx.L.atlas_rt_trace_closest(ctx.h, scene.h, h_in.data_ptr(), N, capi.MASK_ALL, 0.0, capi.INF, h_out.data_ptr(), 0))
            ts.append((time.perf_counter() - t0) * 1e3)
        small = []
        for n in (250_000, 500_000, 1_000_000):
            tt = []
            for i in range(6):
                a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream); ctx.trace(scene, d_rays, n, out=d_out, flags=capi.ASYNC); e.record(stream); torch.cuda.synchronize(); tt.append(a.elapsed_time(e))
            small.append(round(float(np.median(tt[2:])), 3))
        print(f"longest_first={rpw} chunks={chunks} e2e_ms={np.median(ts[2:]):.3f} device_ms(250k,500k,1M)={small}", flush=True)
        for o in (scene, mesh, tlas, blas): o.free()
        ctx.close()
