"""Development aid: pinned-memory PCIe bandwidth of this box (H2D, D2H, both at once) and the host-buffer trace call
(`e2e`) for several pipeline depths, to see how far the end-to-end number is from the copy floor."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from atlas_engine_b200 import capi, workloads as W

dev = torch.device("cuda", 0)
for mb in (12, 24, 48):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device=dev); d2 = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def timed(fn, reps=10):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps): fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps
    t_h2d = timed(lambda: d.copy_(h, non_blocking=True))
    t_d2h = timed(lambda: h2.copy_(d2, non_blocking=True))
    def both():
        with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    t_both = timed(both)
    print(f"{mb} MiB: H2D {n / t_h2d / 1e9:.1f} GB/s  D2H {n / t_d2h / 1e9:.1f} GB/s  both at once {2 * n / t_both / 1e9:.1f} GB/s total ({t_both * 1e3:.2f} ms)", flush=True)

N = 1_000_000
tris = W.soup(N, seed=1234); boxes = W.tri_boxes(tris)
lo, hi = boxes[:, :3].min(0), boxes[:, 3:].max(0)
rays = W.random_rays(N, lo, hi, seed=5678)
root = np.concatenate([lo, hi])[None].astype(np.float32)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
h_in = torch.from_numpy(rays).pin_memory(); h_out = torch.empty_like(h_in).pin_memory()
for chunks in (1, 2, 3, 4, 6, 8):
    os.environ["ATLAS_RT_PIPE_CHUNKS"] = str(chunks)
    ctx = capi.Context(0, stream.cuda_stream)
    blas = ctx.build_blas(boxes, tris); tlas = ctx.build_tlas(root); mesh = ctx.pack_mesh(blas, tris)
    scene = ctx.create_scene([mesh], W.identity_instance(), tlas)
    ts = []
    for i in range(9):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ctx.check(ctx.L.atlas_rt_trace_closest(ctx.h, scene.h, h_in.data_ptr(), N, capi.MASK_ALL, 0.0, capi.INF, h_out.data_ptr(), 0))
        ts.append((time.perf_counter() - t0) * 1e3)
    print(f"chunks={chunks} e2e_ms median {np.median(ts[3:]):.3f} min {min(ts[3:]):.3f}", flush=True)
    for o in (scene, mesh, tlas, blas): o.free()
    ctx.close()
