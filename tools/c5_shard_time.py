"""Development aid: time of one 16-spp C5 frame for ONE of `parts` interleaved shards on one GPU (what each rank of an N-GPU run
does, without the gather):  python tools/c5_shard_time.py [parts]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from atlas_engine_b200 import capi, workloads as W
from test_gpu_configs import c4_scene
parts = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
ctx = capi.Context(0, stream.cuda_stream)
meshes, ib, ir = c4_scene()
blas = ctx.build_blas_batch([W.tri_boxes(t) for t in meshes], meshes)
gm = []
for b, t in zip(blas, meshes):
    m = ctx.pack_mesh(b, t); m.pack_shading(t, payload11=ctx.pack_shading_words(t, W.smooth_normals(t))); gm.append(m)
scene = ctx.create_scene(gm, ir, ctx.build_tlas(ib)); scene.set_materials(capi.make_materials(1))
w, h, bounces, spp, block = 3840, 2160, 4, 16, 4096
cam = W.camera_frame((1000.0, 260.0, -300.0), (1000.0, 60.0, 1000.0), aspect=w / h)
ld = np.array([0.3, 0.9, -0.3]) / np.linalg.norm([0.3, 0.9, -0.3])
prm = capi.pt_params(ld, (3.0, 3.0, 2.5), (0.4, 0.5, 0.8), max_bounces=bounces)
seeds = np.arange(spp * (bounces + 1), dtype=np.float32) * np.float32(0.754878) + np.float32(0.5)
npx = ctx.pathtrace_bounces_interleaved(scene, cam, w, h, prm, 1, 0, seeds[:bounces + 1], 0, parts, block)[0]
part = torch.zeros((npx, 4), dtype=torch.float32, device=dev)
ts = []
for k in range(5):
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    t0 = time.perf_counter()
    ctx.pathtrace_bounces_interleaved(scene, cam, w, h, prm, spp, 0, seeds, 0, parts, block, accum_local=part, flags=capi.ASYNC, count_rays=False)
    host_ms = (time.perf_counter() - t0) * 1e3
    e.record(stream); torch.cuda.synchronize(); ts.append(a.elapsed_time(e))
print(f"parts={parts} lanes={os.environ.get('ATLAS_RT_PT_LANES', 'default')} chain={os.environ.get('ATLAS_RT_CHAIN_LAUNCH', 'default')} "
      f"ms/frame {np.median(ts[1:]):.2f}  ms/pass {np.median(ts[1:]) / spp:.3f}  host enqueue of the last frame {host_ms:.2f} ms", flush=True)
