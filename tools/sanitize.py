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
