"""Development aid: where the time of a host-buffer trace call goes (stage by stage with CUDA events vs the library call)."""
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
d = torch.empty_like(h_in, device=dev)
os.environ["ATLAS_RT_PIPE_CHUNKS"] = sys.argv[1] if len(sys.argv) > 1 else "1"
ctx = capi.Context(0, stream.cuda_stream)
blas = ctx.build_blas(boxes, tris); tlas = ctx.build_tlas(root); mesh = ctx.pack_mesh(blas, tris)
scene = ctx.create_scene([mesh], W.identity_instance(), tlas)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for i in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ev[0].record(stream); d.copy_(h_in, non_blocking=True); ev[1].record(stream)
    ctx.trace(scene, d, N, out=d, flags=capi.ASYNC); ev[2].record(stream)
    h_out.copy_(d, non_blocking=True); ev[3].record(stream)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    if i >= 3: print(f"manual: wall {1e3 * (t1 - t0):.3f} ms  h2d {ev[0].elapsed_time(ev[1]):.3f}  trace {ev[1].elapsed_time(ev[2]):.3f}  d2h {ev[2].elapsed_time(ev[3]):.3f}", flush=True)
for i in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ev[0].record(stream)
    ctx.check(ctx.L.atlas_rt_trace_closest(ctx.h, scene.h, h_in.data_ptr(), N, capi.MASK_ALL, 0.0, capi.INF, h_out.data_ptr(), capi.ASYNC))
    t_enq = time.perf_counter()
    ev[1].record(stream)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    if i >= 3: print(f"library(async)+sync: wall {1e3 * (t1 - t0):.3f} ms  enqueue {1e3 * (t_enq - t0):.3f} ms  device {ev[0].elapsed_time(ev[1]):.3f}", flush=True)
for i in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.check(ctx.L.atlas_rt_trace_closest(ctx.h, scene.h, h_in.data_ptr(), N, capi.MASK_ALL, 0.0, capi.INF, h_out.data_ptr(), 0))
    t1 = time.perf_counter()
    if i >= 3: print(f"library(sync): wall {1e3 * (t1 - t0):.3f} ms", flush=True)
