"""GPU edge cases through the C ABI: exotic instance transforms, tMin/tMax windows, the exact-division slow path
(huge coordinates, tiny / zero direction components), stack overflow reporting, -0.0 inputs, random small meshes with
duplicated and degenerate triangles, device-side outputs."""
import numpy as np
import pytest

import cases as CS
from atlas_engine_b200 import capi, workloads as W
from oracle.pyoracle import Scene as OScene

pytestmark = pytest.mark.gpu


def build_pair(ctx, oracle, mesh_tris, inst_boxes, inst_records):
    blas = [ctx.build_blas(W.tri_boxes(t), t) for t in mesh_tris]
    meshes = [ctx.pack_mesh(b, t) for b, t in zip(blas, mesh_tris)]
    tlas = ctx.build_tlas(inst_boxes)
    scene = ctx.create_scene(meshes, inst_records, tlas)
    inst, tnodes = scene.download()
    obl = [oracle.build_blas(W.tri_boxes(t), t) for t in mesh_tris]
    osc = OScene(tnodes, inst, [b.gpu_nodes() for b in obl], [W.pack_bvh_triangles(t, b.order, b.end_of_node) for t, b in zip(mesh_tris, obl)])
    for b, o in zip(blas, obl):
        n, od, e = b.download()
        assert CS.same_tree(n, od, e, o)
    return scene, osc


def test_mirrored_and_anisotropic_instances(ctx, oracle):
    rng = np.random.default_rng(5)
    meshes = [W.uv_sphere(16, 8), W.soup(2000, seed=9, extent=0.2)]
    mb = [np.concatenate([W.tri_boxes(t)[:, :3].min(0), W.tri_boxes(t)[:, 3:].max(0)]) for t in meshes]
    n = 200
    boxes = np.zeros((n, 6), np.float32)
    inst = np.zeros((n, 16), np.uint32)
    for i in range(n):
        A = rng.normal(size=(3, 3)) * rng.uniform(0.3, 3.0, size=(1, 3))      # shear + non-uniform scale
        if i % 2:
            A[:, 0] *= -1.0                                                     # mirrored (negative determinant)
        M = np.eye(4)
        M[:3, :3] = A
        M[:3, 3] = rng.uniform(-30, 30, size=3)
        mesh = i % 2
        boxes[i] = W.transform_box(mb[mesh], M)
        inst[i, :12] = np.linalg.inv(M)[:3].astype(np.float32).reshape(-1).view(np.uint32)
        inst[i, 12] = mesh
        inst[i, 15] = W.MASK_ALL | (W.MASK_SHADOW if i % 3 else 0)
    scene, osc = build_pair(ctx, oracle, meshes, boxes, inst)
    rays = W.random_rays(100000, boxes[:, :3].min(0), boxes[:, 3:].max(0), seed=6)
    for kw in (dict(), dict(t_min=2.0, t_max=25.0), dict(cull_mask=W.MASK_SHADOW)):
        out = ctx.trace(scene, rays, **kw)
        ref, _ = oracle.trace(osc, rays, nthreads=8, **kw)
        assert np.array_equal(out.view(np.uint32), ref.view(np.uint32)), kw
    sh = rays.copy()
    sh[:, 8] = 15.0
    out = ctx.trace(scene, sh, any_hit=True, t_min=1.0, flags=capi.PER_RAY_TMAX)
    ref, _ = oracle.trace(osc, sh, any_hit=True, per_ray_tmax=True, t_min=1.0, nthreads=8)
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))


def test_exact_division_slow_paths(ctx, oracle):
    """Scenes / rays outside the fast reciprocal path's domain must give the same bits via __fdiv_rn."""
    tris = W.soup(3000, seed=3, extent=0.1)
    big = (tris.astype(np.float64) * 1e22).astype(np.float32)                 # coordinates far above 2^60
    for t in (tris, big):
        boxes = W.tri_boxes(t)
        root = np.concatenate([boxes[:, :3].min(0), boxes[:, 3:].max(0)])[None].astype(np.float32)
        scene, osc = build_pair(ctx, oracle, [t], root, W.identity_instance())
        rays = W.random_rays(40000, root[0, :3], root[0, 3:], seed=4)
        rays[::3, 4] = np.float32(1e-30)     # direction component below 2^-64
        rays[1::3, 5] = 0.0                  # exact zero
        rays[2::5, 6] = np.float32(-0.0)
        out = ctx.trace(scene, rays, t_max=3.0e38)
        ref, _ = oracle.trace(osc, rays, t_max=3.0e38, nthreads=8)
        assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))
        if t is tris:   # at 1e21-scale coordinates the triangle test itself overflows (no hits in either implementation)
            assert (ref[:, 9].view(np.int32) >= 0).mean() > 0.02


def test_stack_overflow_is_reported(ctx, oracle):
    """A hand-made 40-deep chain whose every level pushes one entry: undefined behaviour in the shader (32-entry
    stack, unguarded), ATLAS_RT_ERR_STACK here; the oracle counts the same ray as overflowing."""
    depth = 40
    nodes = np.zeros((depth, 14), dtype=np.uint32)
    tris = np.zeros((depth + 1, 9), dtype=np.float32)
    for k in range(depth):
        x0 = np.float32(k)
        # left child: the rest of the chain (or the last leaf), entered first; right child: a leaf further along the ray
        left = np.array([x0 + 0.1, -1, -1, 1000, 1, 1], np.float32)
        right = np.array([x0 + 0.2, -1, -1, 2000 + k, 1, 1], np.float32)
        nodes[k, 0:6] = left.view(np.uint32)
        nodes[k, 6:12] = right.view(np.uint32)
        nodes[k, 12] = np.uint32(k + 1) if k + 1 < depth else np.int32(~depth).view(np.uint32)
        nodes[k, 13] = np.int32(~k).view(np.uint32)
        tris[k] = [1500 + k, -0.5, -0.5, 1500 + k, 0.5, -0.5, 1500 + k, 0.0, 0.5]
    tris[depth] = [900, -0.5, -0.5, 900, 0.5, -0.5, 900, 0.0, 0.5]
    b = ctx.upload_bvh(nodes, np.arange(depth + 1, dtype=np.uint32), np.ones(depth + 1, np.uint8))
    m = ctx.pack_mesh(b, tris)
    tl = ctx.build_tlas(np.array([[0, -1, -1, 3000, 1, 1]], np.float32))
    sc = ctx.create_scene([m], W.identity_instance(), tl)
    ray = W.pack_rays(np.array([[-5.0, 0.01, 0.02]], np.float32), np.array([[1.0, 1e-4, 1e-4]], np.float32))
    osc = OScene(tl.download()[0].view(np.float32).reshape(-1, 14)[:, :0].reshape(0, 16) if False else sc.download()[1], sc.download()[0],
                 [m.download()[0]], [m.download()[1]])
    ref, ct = oracle.trace(osc, ray)
    assert ct["max_stack"] > 32 and ct["rays_stack_gt32"] == 1
    with pytest.raises(capi.AtlasError, match="ERR_STACK"):
        ctx.trace(sc, ray)
    short = W.pack_rays(np.array([[-5.0, 0.01, 0.02]], np.float32), np.array([[1.0, 1e-4, 1e-4]], np.float32))
    ok = ctx.trace(sc, short, t_max=20.0)        # a short ray only sees the first levels: fine
    ref2, ct2 = oracle.trace(osc, short, t_max=20.0)
    assert np.array_equal(ok.view(np.uint32), ref2.view(np.uint32)) and ct2["rays_stack_gt32"] == 0


def test_negative_zero_inputs_are_flagged_and_equal_in_value(ctx, oracle):
    tris = W.flat_grid(20)
    tris[::2, 1] = np.float32(-0.0)
    tris[::2, 4] = np.float32(-0.0)
    boxes = W.tri_boxes(tris)
    boxes[::2, 1] = np.float32(-0.0)      # make sure the sign survives numpy's min/max
    assert np.signbit(boxes[:, 1]).any()
    b = ctx.build_blas(boxes, tris)
    nodes, order, eon = b.download()
    assert b.stats()["neg_zero"] == 1
    o = oracle.build_blas(boxes, tris)
    assert np.array_equal(order, o.order) and np.array_equal(nodes[:, 12:], o.nodes[:, 12:])
    assert np.array_equal(nodes[:, :12].view(np.float32), o.nodes[:, :12].view(np.float32))     # -0 == +0 by value


def test_random_small_meshes(ctx, oracle):
    rng = np.random.default_rng(2024)
    for trial in range(150):
        n = int(rng.integers(2, 300))
        t = W.soup(n, seed=int(rng.integers(1 << 30)), extent=float(rng.choice([0.001, 0.05, 0.5, 2.0])))
        if trial % 3 == 0:                                  # duplicate a block of triangles
            k = max(1, n // 4)
            t[:k] = t[n - k:]
        if trial % 4 == 0:                                  # degenerate: repeated vertices, axis-flat
            t[::5, 3:6] = t[::5, 0:3]
            t[::7, 2] = t[::7, 5] = t[::7, 8] = 0.125
        if trial % 5 == 0:                                  # quantised coordinates -> many ties
            t = (np.round(t * 8) / 8).astype(np.float32)
        boxes = W.tri_boxes(t)
        b = ctx.build_blas(boxes, t)
        nodes, order, eon = b.download()
        neg_zero = b.stats()["neg_zero"]
        b.free()
        assert neg_zero == int(np.signbit(boxes[boxes == 0]).any())
        # rounding produces -0.0 coordinates; there (and only there) node boxes may differ in the sign of a zero
        same = CS.same_tree_up_to_zero_sign if neg_zero else CS.same_tree
        o = oracle.build_blas(boxes, t)
        assert same(nodes, order, eon, o), (trial, n)
        bx = boxes[: max(1, n // 3)]
        tl = ctx.build_tlas(bx)
        tn, to, te = tl.download()
        tl.free()
        ot = oracle.build_tlas(bx)
        assert same(tn, to, te, ot), (trial, "tlas")


def test_device_pointers_and_device_download(ctx):
    import ctypes as C
    import torch
    tris = W.soup(5000, seed=1)
    b = ctx.build_blas(W.tri_boxes(tris), tris)
    n, m = b.counts()
    nodes, order, eon = b.download()
    p_nodes, p_order, p_eon = C.c_void_p(), C.c_void_p(), C.c_void_p()
    ctx.check(ctx.L.atlas_rt_bvh_device_ptrs(b.h, C.byref(p_nodes), C.byref(p_order), C.byref(p_eon)))
    d_nodes = torch.empty((n, 14), dtype=torch.int32, device="cuda")
    d_order = torch.empty(m, dtype=torch.int32, device="cuda")
    ctx.check(ctx.L.atlas_rt_bvh_download(b.h, d_nodes.data_ptr(), d_order.data_ptr(), None, capi.DEVICE_OUTPUT))
    assert np.array_equal(d_nodes.cpu().numpy().view(np.uint32), nodes) and np.array_equal(d_order.cpu().numpy().view(np.uint32), order)
    assert p_nodes.value and p_order.value and p_eon.value


def test_deep_chain_of_big_nodes(ctx, oracle):
    """Clusters at geometrically growing distance: every split peels the farthest cluster(s) off, so the builder runs
    ~50 levels of big nodes (against ~11 for a soup of the same size) with one or two nodes each — the asynchronous level
    loop, its host flags and the chunk tables at their least favourable."""
    tris = W.geometric_clusters()
    boxes = W.tri_boxes(tris)
    b = ctx.build_blas(boxes, tris)
    o = oracle.build_blas(boxes, tris)
    n, od, e = b.download()
    assert CS.same_tree(n, od, e, o)
    assert b.stats()["levels"] >= 40
    b.free()


def test_plain_any_hit_after_culled_instances(ctx, oracle):
    """Plain HitAny keeps the instance-space ray after an instance culled by the mask (bvh.hsh:387-390 restores only after
    a BLAS exit); tests/test_oracle_anyhit_quirk.py pins the oracle against a literal transcription of the shader, here the
    kernel must equal the oracle on the same kind of scene — for MASK_SHADOW shadow rays over instances without the bit."""
    meshes = [W.uv_sphere(10, 6), W.heightfield(6, 6, spacing=0.5)]
    mb = [np.concatenate([W.tri_boxes(t)[:, :3].min(0), W.tri_boxes(t)[:, 3:].max(0)]) for t in meshes]
    ib, ir = W.random_instances(300, mb, seed=5, extent=(40.0, 8.0, 40.0), scale=(0.8, 2.5))
    ir[::2, 15] = W.MASK_ALL
    scene, osc = build_pair(ctx, oracle, meshes, ib, ir)
    rays = W.random_rays(60000, ib[:, :3].min(0) - 1.0, ib[:, 3:].max(0) + 1.0, seed=12)
    for mask in (W.MASK_SHADOW, W.MASK_ALL):
        out = ctx.trace(scene, rays, any_hit=True, cull_mask=mask, t_max=80.0, flags=capi.COUNTERS)
        gc = ctx.trace_counters()
        ref, oc = oracle.trace(osc, rays, any_hit=True, cull_mask=mask, t_max=80.0, nthreads=8)
        assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))
        assert all(gc[k] == oc[k] for k in oc)
    shadow, _ = oracle.trace(osc, rays, any_hit=True, cull_mask=W.MASK_SHADOW, t_max=80.0, nthreads=8)
    closest = ctx.trace(scene, rays, cull_mask=W.MASK_SHADOW, t_max=80.0)      # HitClosest restores unconditionally
    assert ((shadow[:, 9].view(np.int32) >= 0) != (closest[:, 9].view(np.int32) >= 0)).sum() > 0     # the quirk is exercised
