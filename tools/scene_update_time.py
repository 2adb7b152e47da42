"""Per-frame scene update (what RayTracingWorld::UpdateForSoftwareRayTracing does every frame when instances move): TLAS build
over the instance boxes + scene assembly, on C4 (64 BLASes, 10k instances), host arrays in, synchronous calls, wall clock."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from atlas_engine_b200 import capi, workloads as W
from test_gpu_configs import c4_scene
ctx = capi.Context(0)
meshes, ib, ir = c4_scene()
blas = ctx.build_blas_batch([W.tri_boxes(t) for t in meshes], meshes)
gm = []
for b, t in zip(blas, meshes):
    m = ctx.pack_mesh(b, t); m.pack_shading(t, payload11=ctx.pack_shading_words(t, W.smooth_normals(t))); gm.append(m)
t_tlas, t_scene = [], []
for k in range(12):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); tl = ctx.build_tlas(ib); t1 = time.perf_counter(); sc = ctx.create_scene(gm, ir, tl); t2 = time.perf_counter()
    t_tlas.append((t1 - t0) * 1e3); t_scene.append((t2 - t1) * 1e3)
    sc.free(); tl.free()
print(f"instances {len(ir)} meshes {len(meshes)} triangles {sum(len(t) for t in meshes)}: TLAS build {np.median(t_tlas[2:]):.3f} ms, scene assembly {np.median(t_scene[2:]):.3f} ms (host wall clock, synchronous calls)")
