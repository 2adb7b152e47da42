"""C4's 64 BLASes (1k-100k triangles, ~1M in total), device-resident inputs: one atlas_rt_build_blas per mesh back to
back against ONE atlas_rt_build_blas_batch, CUDA-event timed on the context's stream. Usage (GPU box):
python tools/batch_build_bench.py [workers...] > gpurun_out/batch_build.jsonl"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import __graft_entry__ as g

g.build()
from atlas_engine_b200 import capi, workloads as W
from test_gpu_configs import c4_scene

dev = torch.device("cuda", 0)
meshes, ib, ir = c4_scene()
d_tris = [torch.from_numpy(t).to(dev) for t in meshes]
d_boxes = [torch.from_numpy(W.tri_boxes(t)).to(dev) for t in meshes]
counts = [len(t) for t in meshes]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(workers):
    os.environ["ATLAS_RT_BATCH_WORKERS"] = str(workers)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = capi.Context(0, stream.cuda_stream)

    def timed(fn, reps=7):
        ts = []
        for _ in range(reps + 3):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            out = fn()
            b.record(stream)
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
            for o in out:
                o.free()
        return float(np.median(ts[3:]))
    seq = timed(lambda: [ctx.build_blas(b, t, n, flags=capi.ASYNC) for b, t, n in zip(d_boxes, d_tris, counts)])
    bat = timed(lambda: ctx.build_blas_batch(d_boxes, d_tris, counts, flags=capi.ASYNC))
    print(json.dumps(dict(workers=workers, meshes=len(meshes), triangles=int(sum(counts)), sequential_ms=seq, batch_ms=bat,
                          batch_mtris=sum(counts) / bat / 1e3)), flush=True)
    ctx.close()


for w in [int(x) for x in sys.argv[1:]] or [8, 4]:
    run(w)
