"""Development aid: run the builder / traversal parity batteries on the GPU and print where the first difference is.
Usage (GPU box): python tools/gpu_diag.py [--big] > gpurun_out/diag.txt"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g

g.build()
from atlas_engine_b200 import capi, workloads as W
from oracle.pyoracle import Oracle, Ref, Scene as OScene
import cases as CS

big = "--big" in sys.argv
orc = Oracle()
ctx = capi.Context(0)


def describe(name, nodes, order, eon, o, st):
    ok = CS.same_tree(nodes, order, eon, o)
    msg = f"{name:18s} nodes {nodes.shape[0]}/{o.nodes.shape[0]} refs {order.shape[0]}/{o.order.shape[0]} ok={ok}"
    if not ok:
        if nodes.shape == o.nodes.shape:
            bad = np.nonzero((nodes != o.nodes).any(axis=1))[0]
            msg += f" badnodes={bad.size} first={bad[:4].tolist()}"
            if bad.size:
                i = bad[0]
                msg += f"\n   gpu {nodes[i].view(np.float32)[:12].tolist()} {nodes[i, 12:].view(np.int32).tolist()}"
                msg += f"\n   cpu {o.nodes[i].view(np.float32)[:12].tolist()} {o.nodes[i, 12:].view(np.int32).tolist()}"
        if order.shape == o.order.shape:
            bad = np.nonzero(order != o.order)[0]
            msg += f" badorder={bad.size} first={bad[:4].tolist()}"
            badf = np.nonzero(eon != o.end_of_node)[0]
            msg += f" badflags={badf.size}"
        msg += f" gpustats={st} cpustats={o.stats}"
    print(msg, flush=True)
    return ok


allok = True
for name, tris in CS.build_cases(big).items():
    boxes = W.tri_boxes(tris)
    try:
        t0 = time.time()
        b = ctx.build_blas(boxes, tris)
        dt = time.time() - t0
        nodes, order, eon = b.download()
        st = b.stats()
        b.free()
    except Exception as e:   # noqa
        print(f"{name:18s} EXCEPTION {e}", flush=True)
        allok = False
        continue
    o = orc.build_blas(boxes, tris)
    allok &= describe(name + f" [{dt*1e3:.1f}ms]", nodes, order, eon, o, st)
for name, boxes in CS.tlas_cases().items():
    try:
        b = ctx.build_tlas(boxes)
        nodes, order, eon = b.download()
        st = b.stats()
        b.free()
    except Exception as e:   # noqa
        print(f"{name:18s} EXCEPTION {e}", flush=True)
        allok = False
        continue
    o = orc.build_tlas(boxes)
    allok &= describe(name, nodes, order, eon, o, st)
print("BUILD ALL OK", allok, flush=True)

# ---- traversal over ORACLE-built trees (isolates the trace kernel from the builder)
def oracle_scene(mesh_tris, inst_boxes, inst_records):
    obl = [orc.build_blas(W.tri_boxes(t), t) for t in mesh_tris]
    otl = orc.build_tlas(inst_boxes)
    inst = inst_records[otl.order].copy()
    inst[:, 14] = np.where(otl.end_of_node != 0, -1, np.arange(len(otl.order)) + 1).astype(np.int32).view(np.uint32)
    osc = OScene(otl.gpu_nodes(), inst, [b.gpu_nodes() for b in obl], [W.pack_bvh_triangles(t, b.order, b.end_of_node) for t, b in zip(mesh_tris, obl)])
    gb = [ctx.upload_bvh(b.nodes, b.order, b.end_of_node) for b in obl]
    gm = [ctx.pack_mesh(b, t) for b, t in zip(gb, mesh_tris)]
    gt = ctx.upload_bvh(otl.nodes, otl.order, otl.end_of_node)
    gs = ctx.create_scene(gm, inst_records, gt)
    return osc, gs


def compare_trace(name, osc, gs, rays, **kw):
    any_hit = kw.get("any_hit", False)
    flags = capi.COUNTERS | (capi.PER_RAY_TMAX if kw.get("per_ray") else 0)
    t0 = time.time()
    out = ctx.trace(gs, rays, any_hit=any_hit, flags=flags, cull_mask=kw.get("mask", capi.MASK_ALL))
    dt = time.time() - t0
    gc = ctx.trace_counters()
    oo, oc = orc.trace(osc, rays, any_hit=any_hit, per_ray_tmax=kw.get("per_ray", False), cull_mask=kw.get("mask", capi.MASK_ALL), nthreads=os.cpu_count())
    same = np.array_equal(out.view(np.uint32), oo.view(np.uint32))
    bad = np.nonzero((out.view(np.uint32) != oo.view(np.uint32)).any(axis=1))[0]
    print(f"trace {name:22s} rays={len(rays)} bitexact={same} bad={bad.size} hitrate={(oo[:,9].view(np.int32)>=0).mean():.3f} counters_equal={all(gc[k]==oc[k] for k in oc)} e2e={dt*1e3:.1f}ms", flush=True)
    if not same:
        i = bad[0]
        print("   gpu", out[i].tolist(), out[i, 9:11].view(np.int32).tolist())
        print("   cpu", oo[i].tolist(), oo[i, 9:11].view(np.int32).tolist())
        print("   counters gpu", gc, "cpu", oc)
    return same


tok = True
meshes = [W.uv_sphere(), W.soup_with_giants(5000, seed=2), W.heightfield(60, 60)]
mb = [np.concatenate([W.tri_boxes(t)[:, :3].min(0), W.tri_boxes(t)[:, 3:].max(0)]) for t in meshes]
for k, t in enumerate(meshes):
    osc, gs = oracle_scene([t], mb[k][None], W.identity_instance())
    lo, hi = mb[k][:3], mb[k][3:]
    pad = (hi - lo) * 0.2
    rays = W.random_rays(100000, lo - pad, hi + pad, seed=10 + k)
    tok &= compare_trace(f"single{k}", osc, gs, rays)
    sh = rays.copy(); sh[:, 8] = 0.3 * np.linalg.norm(hi - lo)
    tok &= compare_trace(f"single{k}_any", osc, gs, sh, any_hit=True, per_ray=True)
ib, ir = W.random_instances(2000, mb, seed=9, extent=(300.0, 60.0, 300.0))
osc, gs = oracle_scene(meshes, ib, ir)
lo, hi = ib[:, :3].min(0), ib[:, 3:].max(0)
rays = W.random_rays(200000, lo, hi, seed=33)
tok &= compare_trace("two_level", osc, gs, rays)
sh = rays.copy(); sh[:, 8] = 80.0
tok &= compare_trace("two_level_any", osc, gs, sh, any_hit=True, per_ray=True)
tok &= compare_trace("two_level_shadowmask", osc, gs, rays, mask=capi.MASK_SHADOW)
rz = rays[:5000].copy(); rz[::2, 4] = 0.0; rz[::3, 5] = 0.0; rz[:, 3] = np.where(np.arange(5000) % 7 == 0, -1, np.arange(5000)).astype(np.int32).view(np.float32)
rz[1::11, 6] = np.nan
tok &= compare_trace("two_level_edge", osc, gs, rz)
print("TRACE ALL OK", tok, flush=True)
