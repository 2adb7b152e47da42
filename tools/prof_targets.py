"""Small, quick targets for ncu (one GPU): python tools/prof_targets.py c2 | terrain | c5 | e2e
  c2       1M-triangle soup, three device-resident closest-hit traces of 1M rays with 16-byte hit records (bench.py's step)
  terrain  8M-triangle terrain, three traces of 4M incoherent rays (bench.py's out_of_l2 figure)
  c5       one path-traced 3840x2160 sample pass, 4 bounces, device-side loop (after one warm-up pass)
  e2e      three host-buffer traces of 1M rays (streaming upload)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from atlas_engine_b200 import capi, workloads as W

what = sys.argv[1] if len(sys.argv) > 1 else "c2"
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ctx = capi.Context(0, stream.cuda_stream)


def single(tris):
    boxes = W.tri_boxes(tris)
    d_t, d_b = torch.from_numpy(tris).to(dev), torch.from_numpy(boxes).to(dev)
    blas = ctx.build_blas(d_b, d_t, len(tris))
    lo, hi = boxes[:, :3].min(0), boxes[:, 3:].max(0)
    tlas = ctx.build_tlas(np.concatenate([lo, hi])[None].astype(np.float32))
    mesh = ctx.pack_mesh(blas, d_t, len(tris))
    return ctx.create_scene([mesh], W.identity_instance(), tlas), lo, hi, (blas, tlas, mesh, d_t, d_b)


if what in ("c2", "e2e"):
    scene, lo, hi, keep = single(W.soup(1_000_000, seed=1234))
    rays = W.random_rays(1_000_000, lo, hi, seed=5678)
    if what == "c2":
        d_r = torch.from_numpy(rays).to(dev)
        d_h = torch.empty((len(rays), 4), dtype=torch.float32, device=dev)
        for _ in range(3):
            ctx.trace(scene, d_r, len(rays), out=d_h, flags=capi.HITS_ONLY)
    else:
        h_r = torch.from_numpy(rays).pin_memory()
        h_h = torch.empty((len(rays), 4), dtype=torch.float32).pin_memory()
        for _ in range(3):
            ctx.check(ctx.L.atlas_rt_trace_closest(ctx.h, scene.h, h_r.data_ptr(), len(rays), capi.MASK_ALL, 0.0, capi.INF, h_h.data_ptr(), capi.HITS_ONLY))
elif what == "terrain":
    scene, lo, hi, keep = single(W.heightfield(2000, 2000))
    hi2 = hi.copy()
    hi2[1] += 40.0
    rays = W.random_rays(4_000_000, lo, hi2, seed=77)
    d_r = torch.from_numpy(rays).to(dev)
    d_h = torch.empty((len(rays), 4), dtype=torch.float32, device=dev)
    for _ in range(3):
        ctx.trace(scene, d_r, len(rays), out=d_h, flags=capi.HITS_ONLY)
elif what == "c5":
    from test_gpu_configs import c4_scene
    meshes, ib, ir = c4_scene()
    blas = ctx.build_blas_batch([W.tri_boxes(t) for t in meshes], meshes)
    gm = []
    for b, t in zip(blas, meshes):
        m = ctx.pack_mesh(b, t)
        m.pack_shading(t, payload11=ctx.pack_shading_words(t, W.smooth_normals(t)))
        gm.append(m)
    tlas = ctx.build_tlas(ib)
    scene = ctx.create_scene(gm, ir, tlas)
    scene.set_materials(capi.make_materials(1))
    w, h, bounces = 3840, 2160, 4
    cam = W.camera_frame((1000.0, 260.0, -300.0), (1000.0, 60.0, 1000.0), aspect=w / h)
    ld = np.array([0.3, 0.9, -0.3]) / np.linalg.norm([0.3, 0.9, -0.3])
    prm = capi.pt_params(ld, (3.0, 3.0, 2.5), (0.4, 0.5, 0.8), max_bounces=bounces)
    accum = torch.zeros((w * h, 4), dtype=torch.float32, device=dev)
    seeds = np.arange(2 * (bounces + 1), dtype=np.float32) * np.float32(0.754878) + np.float32(0.5)
    flags = capi.RAY_BINNING if os.environ.get("ATLAS_BENCH_BINNING") else 0
    print("PROFILE_MARK pathtrace", flush=True)
    traced = ctx.pathtrace_bounces(scene, cam, w, h, prm, 2, 0, seeds, accum, flags=flags)
    print("closest-hit rays in two passes:", traced)
torch.cuda.synchronize()
print("done", what)
