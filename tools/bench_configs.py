"""Measure the five BASELINE.json configurations on one GPU (build Mtris/s, closest / any-hit Mrays/s, roofline
fraction from the kernel's own visit counters) and check each against the oracle on a sample.
Usage (GPU box): python tools/bench_configs.py [C1 C3 C4 C5] > gpurun_out/configs.jsonl"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as g

g.build()
from atlas_engine_b200 import capi, workloads as W
from oracle.pyoracle import Oracle, Scene as OScene

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ctx = capi.Context(0, stream.cuda_stream)
orc = Oracle()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
which = set(sys.argv[1:]) or {"C1", "C3", "C4", "C5"}


def timed(fn, reps=5):
    ts = []
    for _ in range(reps + 2):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts[2:]))


def time_build(boxes, tris):
    db, dt = torch.from_numpy(boxes).to(dev), torch.from_numpy(tris).to(dev)
    keep = []

    def run():
        for k in keep:
            k.free()
        keep.clear()
        keep.append(ctx.build_blas(db, dt, len(tris), flags=capi.ASYNC))
    ms = timed(run, 3)
    for k in keep:
        k.free()
    return ms


def trace_stats(scene, rays, any_hit=False, mask=capi.MASK_ALL, per_ray=False):
    d = torch.from_numpy(rays).to(dev)
    o = torch.empty_like(d)
    fl = capi.PER_RAY_TMAX if per_ray else 0
    ms = timed(lambda: ctx.trace(scene, d, len(rays), out=o, any_hit=any_hit, cull_mask=mask, flags=capi.ASYNC | fl))
    ctx.trace(scene, d, len(rays), out=o, any_hit=any_hit, cull_mask=mask, flags=capi.COUNTERS | fl)
    ct = ctx.trace_counters()
    nbytes = 96 * len(rays) + 64 * (ct["tlas_nodes"] + ct["blas_nodes"] + ct["instances"]) + 48 * ct["triangles"]
    out = o.cpu().numpy()
    return dict(ms=ms, mrays=len(rays) / ms / 1e3, bytes_per_ray=nbytes / len(rays), roofline_frac=nbytes / (ms * 1e-3) / 1e9 / PEAK,
                hit_rate=float((out[:, 9].view(np.int32) >= 0).mean()), max_stack=ct["max_stack"]), out


def emit(name, **kw):
    print(json.dumps(dict(config=name, **kw)), flush=True)


def single_mesh_scene(tris):
    boxes = W.tri_boxes(tris)
    blas = ctx.build_blas(boxes, tris)
    root = np.concatenate([boxes[:, :3].min(0), boxes[:, 3:].max(0)])[None].astype(np.float32)
    tlas = ctx.build_tlas(root)
    mesh = ctx.pack_mesh(blas, tris)
    scene = ctx.create_scene([mesh], W.identity_instance(), tlas)
    return scene, blas, mesh, tlas, boxes, root[0]


def oracle_check(scene_parts, tris_list, inst_boxes, inst_records, rays, out, blas_list, sample=100000, **kw):
    """Build equality for every BLAS + TLAS and bit-exact hits on the first `sample` rays."""
    obl = [orc.build_blas(W.tri_boxes(t), t) for t in tris_list]
    ok_build = True
    for b, o in zip(blas_list, obl):
        n, od, e = b.download()
        ok_build &= n.shape == o.nodes.shape and np.array_equal(n, o.nodes) and np.array_equal(od, o.order) and np.array_equal(e, o.end_of_node)
    otl = orc.build_tlas(inst_boxes)
    inst = inst_records[otl.order].copy()
    inst[:, 14] = np.where(otl.end_of_node != 0, -1, np.arange(len(otl.order)) + 1).astype(np.int32).view(np.uint32)
    osc = OScene(otl.gpu_nodes(), inst, [b.gpu_nodes() for b in obl], [W.pack_bvh_triangles(t, b.order, b.end_of_node) for t, b in zip(tris_list, obl)])
    ref, _ = orc.trace(osc, rays[:sample], nthreads=os.cpu_count(), **kw)
    return bool(ok_build), bool(np.array_equal(ref.view(np.uint32), out[:sample].view(np.uint32)))


if "C1" in which:
    tris = W.atrium(128)
    scene, blas, mesh, tlas, boxes, root = single_mesh_scene(tris)
    build_ms = time_build(boxes, tris)
    eye, origin, right, bottom = W.camera_frame((30.0 * 0.05 * 10, 25.0 * 0.05 * 4, 6.0), (600 * 0.05, 3.0, 6.5), fov_deg=47.0)
    rays = ctx.generate_primary_rays(eye, origin, right, bottom, 1920, 1080, 1)
    st, out = trace_stats(scene, rays)
    okb, okt = oracle_check(None, [tris], root[None], W.identity_instance(), rays, out, [blas])
    # the RTAO caller (data/shader/ao/rtao.csh:97-100): short any-hit rays from every primary hit, cosine
    # distributed around the geometric normal, tMax = radius
    hit = out[:, 9].view(np.int32) >= 0
    P = (out[:, 0:3] + out[:, 4:7] * out[:, 8:9])[hit]
    order = blas.download()[1]
    T = tris[order[out[hit, 9].view(np.int32)]].reshape(-1, 3, 3)
    Ng = np.cross(T[:, 0] - T[:, 1], T[:, 0] - T[:, 2])
    Ng /= np.maximum(np.linalg.norm(Ng, axis=1, keepdims=True), 1e-30)
    Ng *= np.where((Ng * out[hit, 4:7]).sum(1, keepdims=True) > 0, -1.0, 1.0)
    rng = np.random.default_rng(11)
    ao = []
    for s4 in range(4):
        u0, u1 = rng.random(len(P)), rng.random(len(P))
        r_, phi = np.sqrt(u0), 2 * np.pi * u1
        up = np.where((np.abs(Ng[:, 2]) < 0.999)[:, None], [0.0, 0.0, 1.0], [1.0, 0.0, 0.0])
        tg = np.cross(up, Ng); tg /= np.linalg.norm(tg, axis=1, keepdims=True)
        bt = np.cross(Ng, tg)
        d_ = tg * (r_ * np.cos(phi))[:, None] + bt * (r_ * np.sin(phi))[:, None] + Ng * np.sqrt(1 - u0)[:, None]
        ao.append(W.pack_rays((P + Ng * 0.01).astype(np.float32), d_.astype(np.float32), t=np.full(len(P), 1.5, np.float32)))
    ao = np.concatenate(ao)
    st_ao, out_ao = trace_stats(scene, ao, any_hit=True, mask=capi.MASK_SHADOW, per_ray=True)
    _, ok_ao = oracle_check(None, [tris], root[None], W.identity_instance(), ao, out_ao, [blas], any_hit=True, per_ray_tmax=True, cull_mask=capi.MASK_SHADOW)
    emit("F1 RTAO-style caller on the C1 scene: 4 short any-hit rays (tMax 1.5) per primary hit", rays=len(ao), any=st_ao,
         hits_equal_oracle_100k=ok_ao)
    emit("C1 atrium stand-in for sponza, 1920x1080 primaries", triangles=len(tris), refs=blas.counts()[1], build_ms=build_ms,
         build_mtris=len(tris) / build_ms / 1e3, stats=blas.stats(), closest=st, build_equals_oracle=okb, hits_equal_oracle_100k=okt)

if "C3" in which:
    tris = W.heightfield(2000, 2000)
    scene, blas, mesh, tlas, boxes, root = single_mesh_scene(tris)
    build_ms = time_build(boxes, tris)
    c = (root[:3] + root[3:]) / 2
    eye = (float(c[0]), float(root[4]) + 60.0, float(c[2]) - 600.0)
    eye_, origin, right, bottom = W.camera_frame(eye, (float(c[0]), float(eye[1]) - 600.0 * np.tan(np.radians(30.0)), float(c[2])), aspect=3840 / 2160)
    rays = ctx.generate_primary_rays(eye_, origin, right, bottom, 3840, 2160, 1)
    st, out = trace_stats(scene, rays)
    hit = out[:, 9].view(np.int32) >= 0
    sun = np.array([0.0, 1.0, 0.33], dtype=np.float64)
    sun /= np.linalg.norm(sun)
    P = out[:, 0:3] + out[:, 4:7] * out[:, 8:9]
    sh = W.pack_rays((P + np.array([0, 0.1, 0], np.float32)).astype(np.float32), np.broadcast_to(sun.astype(np.float32), P.shape).copy(),
                     ids=np.where(hit, np.arange(len(P)), -1), t=np.full(len(P), 1e12, np.float32))
    st2, out2 = trace_stats(scene, sh, any_hit=True, mask=capi.MASK_SHADOW, per_ray=True)
    t0 = time.time()
    okb, okt = oracle_check(None, [tris], root[None], W.identity_instance(), rays, out, [blas], sample=200000)
    emit("C3 8M-triangle terrain, 3840x2160 primaries + shadow any-hit", triangles=len(tris), build_ms=build_ms, build_mtris=len(tris) / build_ms / 1e3,
         closest=st, shadow_any=st2, build_equals_oracle=okb, hits_equal_oracle_200k=okt, oracle_seconds=time.time() - t0)

if "C4" in which or "C5" in which:
    rng = np.random.default_rng(64)
    meshes = []
    for k in range(64):
        n = int(np.exp(rng.uniform(np.log(1000), np.log(100000))))
        if k % 2 == 0:
            seg = max(8, int(np.sqrt(n / 2)))
            meshes.append(W.uv_sphere(seg, max(4, seg // 2), radius=1.0 + 0.1 * k))
        else:
            side = max(4, int(np.sqrt(n / 2)))
            meshes.append(W.heightfield(side, side, spacing=20.0 / side) * np.float32(0.2))
    mb = [np.concatenate([W.tri_boxes(t)[:, :3].min(0), W.tri_boxes(t)[:, 3:].max(0)]) for t in meshes]
    ib, ir = W.random_instances(10000, mb, seed=4242)
    t0 = time.time()
    blas = [ctx.build_blas(W.tri_boxes(t), t) for t in meshes]
    blas_total_ms = (time.time() - t0) * 1e3
    gm = [ctx.pack_mesh(b, t) for b, t in zip(blas, meshes)]
    dib = torch.from_numpy(ib).to(dev)
    keep = []

    def tl():
        for k in keep:
            k.free()
        keep.clear()
        keep.append(ctx.build_tlas(dib, len(ib), flags=capi.ASYNC))
    tlas_ms = timed(tl)
    tlas = ctx.build_tlas(ib)
    scene = ctx.create_scene(gm, ir, tlas)
    lo, hi = ib[:, :3].min(0), ib[:, 3:].max(0)
    if "C4" in which:
        rays = W.random_rays(4_000_000, lo, hi, seed=5678)
        st, out = trace_stats(scene, rays)
        sh = rays.copy()
        sh[:, 8] = 200.0
        st2, _ = trace_stats(scene, sh, any_hit=True, mask=capi.MASK_SHADOW, per_ray=True)
        okb, okt = oracle_check(None, meshes, ib, ir, rays, out, blas + [], sample=100000)
        n, od, e = tlas.download()
        otl = orc.build_tlas(ib)
        emit("C4 TLAS over 10k instances of 64 BLASes (1k-100k tris), 4M random rays", total_triangles=int(sum(len(t) for t in meshes)),
             blas_builds_total_ms_host_timed=blas_total_ms, tlas_build_ms=tlas_ms, closest=st, shadow_any=st2, blas_equal_oracle=okb,
             tlas_equals_oracle=bool(np.array_equal(n, otl.nodes) and np.array_equal(od, otl.order)), hits_equal_oracle_100k=okt)
    if "C5" in which:
        # BASELINE configs[4]: 3840x2160 x 16 spp, 4 bounces, whole loop on the device (atlas_rt_pathtrace_bounces)
        w, h, spp, bounces = 3840, 2160, 16, 4
        cam = W.camera_frame((1000.0, 260.0, -300.0), (1000.0, 60.0, 1000.0), aspect=w / h)
        for m, t in zip(gm, meshes):
            m.pack_shading(t, payload11=ctx.pack_shading_words(t, W.smooth_normals(t)))
        scene5 = ctx.create_scene(gm, ir, tlas)
        scene5.set_materials(capi.make_materials(1))
        ld = np.array([0.3, 0.9, -0.3]) / np.linalg.norm([0.3, 0.9, -0.3])
        prm = capi.pt_params(ld, (3.0, 3.0, 2.5), (0.4, 0.5, 0.8), max_bounces=bounces)
        seeds = np.arange(spp * (bounces + 1), dtype=np.float32) * np.float32(0.754878) + np.float32(0.5)
        for flags, label in ((0, "no binning"), (capi.RAY_BINNING, "octahedral ray binning before bounces >= 1")):
            accum = torch.zeros((w * h, 4), dtype=torch.float32, device=dev)
            ctx.pathtrace_bounces(scene5, cam, w, h, prm, 1, 0, seeds[:bounces + 1], accum, flags=flags)   # warm-up pass
            accum.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record(stream)
            traced = ctx.pathtrace_bounces(scene5, cam, w, h, prm, spp, 0, seeds, accum, flags=flags)
            b.record(stream)
            torch.cuda.synchronize()
            ms = a.elapsed_time(b)
            acc = accum.cpu().numpy()
            emit("C5 path tracer 3840x2160 x 16 spp, 4 bounces (+ shadow rays), device-side loop, " + label, closest_rays_traced=int(traced),
                 total_ms=ms, closest_mrays_per_s=traced / ms / 1e3, paths_finished=float(acc[:, 3].sum()), mean_radiance=acc[:, :3].mean(axis=0).tolist(),
                 note="each closest-hit ray also spawns one shadow any-hit ray when lit; Mrays/s counts closest-hit rays only")
