"""compute-sanitizer target: small builder / traversal / path-tracer cases (SURVEY.md §7 test plan item 4).
Usage (GPU box): compute-sanitizer --tool memcheck python tools/sanitize.py ; ... --tool racecheck ..."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from atlas_engine_b200 import capi, workloads as W

ctx = capi.Context(0)
cases = [W.soup(3000, seed=1), W.soup_with_giants(4000, seed=2), W.coincident(40, 40), W.heightfield(30, 30), W.soup(7, seed=3)]
meshes, blas = [], []
for t in cases:
    b = ctx.build_blas(W.tri_boxes(t), t)
    blas.append(b)
    meshes.append(ctx.pack_mesh(b, t))
mb = [np.concatenate([W.tri_boxes(t)[:, :3].min(0), W.tri_boxes(t)[:, 3:].max(0)]) for t in cases]
ib, ir = W.random_instances(1500, mb, seed=5, extent=(40.0, 10.0, 40.0))
tl = ctx.build_tlas(ib)
sc = ctx.create_scene(meshes, ir, tl)
rays = W.random_rays(20000, ib[:, :3].min(0), ib[:, 3:].max(0), seed=6)
out = ctx.trace(sc, rays, flags=capi.COUNTERS)
sh = rays.copy()
sh[:, 8] = 10.0
occ = ctx.trace(sc, sh, any_hit=True, flags=capi.PER_RAY_TMAX)
print("hits", int((out[:, 9].view(np.int32) >= 0).sum()), "occluded", int((occ[:, 9].view(np.int32) >= 0).sum()), ctx.trace_counters())
# the pipelined host-buffer path (two compute streams, >= 262144 rays) and an in-place device batch
big = W.random_rays(300000, ib[:, :3].min(0), ib[:, 3:].max(0), seed=7)
res = ctx.trace(sc, big)
print("pipelined hits", int((res[:, 9].view(np.int32) >= 0).sum()))
# round 2: batched BLAS builds, the streaming host-buffer trace (persistent launch fed chunk by chunk, per-chunk ordering on a
# second stream), 16-byte hit records, and the path tracer's device-side bounce loop with several lanes of sample passes
bb = ctx.build_blas_batch([W.tri_boxes(t) for t in cases], cases)
print("batch builds", [b.counts() for b in bb])
os.environ["ATLAS_RT_TRACE_STREAMING"] = "1"
sctx = capi.Context(0)
smesh = [sctx.pack_mesh(sctx.build_blas(W.tri_boxes(t), t), t) for t in cases]
ssc = sctx.create_scene(smesh, ir, sctx.build_tlas(ib))
res2 = sctx.trace(ssc, big)
hits2 = sctx.trace(ssc, big, flags=capi.HITS_ONLY)
print("streaming hits", int((res2[:, 9].view(np.int32) >= 0).sum()), "equal", bool(np.array_equal(res2.view(np.uint32), res.view(np.uint32))),
      bool(np.array_equal(hits2.view(np.uint32), res[:, 8:12].view(np.uint32))))
del os.environ["ATLAS_RT_TRACE_STREAMING"]
import torch
for m, t in zip(meshes, cases):
    m.pack_shading(t, payload11=ctx.pack_shading_words(t, W.smooth_normals(t)))
psc = ctx.create_scene(meshes, ir, tl)
psc.set_materials(capi.make_materials(1))
w, h, bounces, frames = 96, 64, 3, 6
cam = W.camera_frame((20.0, 30.0, -20.0), (20.0, 0.0, 20.0), aspect=w / h)
ld = np.array([0.2, 0.8, 0.4]) / np.linalg.norm([0.2, 0.8, 0.4])
prm = capi.pt_params(ld, (3.0, 3.0, 2.5), (0.4, 0.5, 0.8), max_bounces=bounces)
accum = torch.zeros((w * h, 4), dtype=torch.float32, device="cuda")
traced = ctx.pathtrace_bounces(psc, cam, w, h, prm, frames, 0, np.arange(frames * (bounces + 1), dtype=np.float32) * np.float32(0.7) + np.float32(0.5), accum)
print("path tracer", traced, "rays,", float(accum[:, 3].sum().item()), "finished paths")
