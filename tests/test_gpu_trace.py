"""GPU parity tests for traversal, pack and scene assembly, through the C ABI, against the oracle restatement of
data/shader/raytracer/bvh.hsh. IDs must be bit-exact; t and barycentrics are required within 1e-6 relative by the
north star and are in fact compared bit for bit (the kernels use no FMA contraction)."""
import numpy as np
import pytest

from atlas_engine_b200 import capi, workloads as W
from oracle.pyoracle import Scene as OScene

pytestmark = pytest.mark.gpu
REL_TOL = 1e-6   # north-star tolerance for t / barycentrics


def gpu_and_oracle_scene(ctx, oracle, mesh_tris, inst_boxes, inst_records):
    """Scene built entirely by the CUDA path, and the same scene built by the oracle."""
    blas = [ctx.build_blas(W.tri_boxes(t), t) for t in mesh_tris]
    meshes = [ctx.pack_mesh(b, t) for b, t in zip(blas, mesh_tris)]
    tlas = ctx.build_tlas(inst_boxes)
    scene = ctx.create_scene(meshes, inst_records, tlas)
    obl = [oracle.build_blas(W.tri_boxes(t), t) for t in mesh_tris]
    otl = oracle.build_tlas(inst_boxes)
    inst = inst_records[otl.order].copy()
    inst[:, 14] = np.where(otl.end_of_node != 0, -1, np.arange(len(otl.order)) + 1).astype(np.int32).view(np.uint32)
    osc = OScene(otl.gpu_nodes(), inst, [b.gpu_nodes() for b in obl],
                 [W.pack_bvh_triangles(t, b.order, b.end_of_node) for t, b in zip(mesh_tris, obl)])
    return scene, osc, (blas, meshes, tlas)


def assert_same_hits(out, ref):
    assert np.array_equal(out[:, 9:11].view(np.int32), ref[:, 9:11].view(np.int32))          # hitID, instanceID
    assert np.array_equal(out[:, 0:7].view(np.uint32), ref[:, 0:7].view(np.uint32))          # ray passes through
    for col in (8, 7, 11):                                                                      # t, u, v
        a, b = out[:, col].astype(np.float64), ref[:, col].astype(np.float64)
        assert np.all(np.abs(a - b) <= REL_TOL * np.abs(b))
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))                             # and in fact bitwise


@pytest.fixture(scope="module")
def world(ctx, oracle):
    meshes = [W.uv_sphere(), W.soup_with_giants(5000, seed=2), W.heightfield(60, 60)]
    mb = [np.concatenate([W.tri_boxes(t)[:, :3].min(0), W.tri_boxes(t)[:, 3:].max(0)]) for t in meshes]
    ib, ir = W.random_instances(2000, mb, seed=9, extent=(300.0, 60.0, 300.0))
    ir[::5, 15] = W.MASK_ALL
    scene, osc, keep = gpu_and_oracle_scene(ctx, oracle, meshes, ib, ir)
    return scene, osc, ib, keep


def test_pack_and_scene_layouts(ctx, oracle, world):
    scene, osc, ib, (blas, meshes, tlas) = world
    inst, tnodes = scene.download()
    assert np.array_equal(inst, osc.instances.view(np.uint32))
    assert np.array_equal(tnodes.view(np.uint32), osc.tlas_nodes.view(np.uint32))
    for m, on, ot in zip(meshes, osc.blas_nodes, osc.bvh_tris):
        gn, gt = m.download()
        assert np.array_equal(gn.view(np.uint32), on.view(np.uint32))
        assert np.array_equal(gt.view(np.uint32), ot.view(np.uint32))


def test_two_level_closest(ctx, oracle, world):
    scene, osc, ib, _ = world
    rays = W.random_rays(200000, ib[:, :3].min(0), ib[:, 3:].max(0), seed=33)
    out = ctx.trace(scene, rays, flags=capi.COUNTERS)
    gc = ctx.trace_counters()
    ref, oc = oracle.trace(osc, rays, nthreads=8)
    assert_same_hits(out, ref)
    assert all(gc[k] == oc[k] for k in oc)            # same visit counts => same traversal order
    assert 0.2 < (ref[:, 9].view(np.int32) >= 0).mean() < 0.8
    plain = ctx.trace(scene, rays)                     # the non-counting kernel variant
    assert np.array_equal(plain.view(np.uint32), ref.view(np.uint32))


def test_two_level_any_and_masks(ctx, oracle, world):
    scene, osc, ib, _ = world
    rays = W.random_rays(100000, ib[:, :3].min(0), ib[:, 3:].max(0), seed=34)
    rays[:, 8] = 80.0
    for mask in (W.MASK_ALL, W.MASK_SHADOW):
        out = ctx.trace(scene, rays, any_hit=True, cull_mask=mask, flags=capi.PER_RAY_TMAX)
        ref, _ = oracle.trace(osc, rays, any_hit=True, per_ray_tmax=True, cull_mask=mask, nthreads=8)
        assert_same_hits(out, ref)
    out = ctx.trace(scene, rays, cull_mask=W.MASK_SHADOW)
    ref, _ = oracle.trace(osc, rays, cull_mask=W.MASK_SHADOW, nthreads=8)
    assert_same_hits(out, ref)
    out = ctx.trace(scene, rays, any_hit=True, t_max=25.0)           # global tMax argument
    ref, _ = oracle.trace(osc, rays, any_hit=True, t_max=25.0, nthreads=8)
    assert_same_hits(out, ref)


def test_edge_rays(ctx, oracle, world):
    """Dead rays (ID < 0), NaN directions, exact-zero direction components, empty batch."""
    scene, osc, ib, _ = world
    rays = W.random_rays(5000, ib[:, :3].min(0), ib[:, 3:].max(0), seed=35)
    rays[::2, 4] = 0.0
    rays[::3, 5] = 0.0
    rays[:, 3] = np.where(np.arange(5000) % 7 == 0, -1, np.arange(5000)).astype(np.int32).view(np.float32)
    rays[1::11, 6] = np.nan
    out = ctx.trace(scene, rays)
    ref, _ = oracle.trace(osc, rays, nthreads=4)
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))
    assert ctx.trace(scene, np.zeros((0, 12), np.float32)).shape == (0, 12)


def test_single_instance_quirk_scene(ctx, oracle):
    """TLAS over one instance: the synthetic node with leftPtr == rightPtr == ~0 and the two-entry instance array."""
    tris = W.uv_sphere(24, 12)
    boxes = W.tri_boxes(tris)
    root = np.concatenate([boxes[:, :3].min(0), boxes[:, 3:].max(0)])[None].astype(np.float32)
    scene, osc, _ = gpu_and_oracle_scene(ctx, oracle, [tris], root, W.identity_instance())
    inst, tnodes = scene.download()
    assert inst.shape[0] == 2 and tnodes.shape[0] == 1
    rays = W.random_rays(50000, root[0, :3] - 0.5, root[0, 3:] + 0.5, seed=36)
    out = ctx.trace(scene, rays)
    ref, _ = oracle.trace(osc, rays, nthreads=4)
    assert_same_hits(out, ref)


def test_full_size_c2_properties(ctx, oracle):
    """BASELINE C2 at full size: 1M-triangle soup, 1M rays. Bit-exact against the oracle on a 100k-ray sample, and
    size-independent properties on all rays: (1) an any-hit ray with tMax just above the closest t must report a hit,
    one with tMax just below the closest t of the FIRST hit along the ray must not; (2) re-tracing with tMax = t_closest
    * (1 + eps) finds the same triangle; (3) misses stay misses when the ray is shortened."""
    tris = W.soup(1_000_000, seed=1234)
    boxes = W.tri_boxes(tris)
    root = np.concatenate([boxes[:, :3].min(0), boxes[:, 3:].max(0)])[None].astype(np.float32)
    scene, osc, _ = gpu_and_oracle_scene(ctx, oracle, [tris], root, W.identity_instance())
    rays = W.random_rays(1_000_000, root[0, :3], root[0, 3:], seed=5678)
    out = ctx.trace(scene, rays)
    ref, _ = oracle.trace(osc, rays[:100000], nthreads=8)
    assert np.array_equal(out[:100000].view(np.uint32), ref.view(np.uint32))
    # the host-buffer call is pipelined in chunks over two compute streams: check the last chunk against the oracle too,
    # and the whole batch against the device-pointer call (in place and into a separate buffer)
    tail, _ = oracle.trace(osc, rays[-50000:], nthreads=8)
    assert np.array_equal(out[-50000:].view(np.uint32), tail.view(np.uint32))
    import torch
    d_rays = torch.from_numpy(rays).cuda()
    d_out = torch.empty_like(d_rays)
    ctx.trace(scene, d_rays, len(rays), out=d_out)
    assert np.array_equal(d_out.cpu().numpy().view(np.uint32), out.view(np.uint32))
    ctx.trace(scene, d_rays, len(rays))
    assert np.array_equal(d_rays.cpu().numpy().view(np.uint32), out.view(np.uint32))
    # host rays in, hits left on the device (the multi-GPU bench's call): same pipeline without the download
    ctx.check(ctx.L.atlas_rt_trace_closest(ctx.h, scene.h, rays.ctypes.data, len(rays), capi.MASK_ALL, 0.0, capi.INF, d_out.data_ptr(),
                                           capi.DEVICE_OUTPUT))
    assert np.array_equal(d_out.cpu().numpy().view(np.uint32), out.view(np.uint32))
    hit = out[:, 9].view(np.int32) >= 0
    t = out[:, 8]
    assert 0.1 < hit.mean() < 0.9 and np.all(t[~hit] == np.float32(1e12))
    longer = rays.copy()
    longer[:, 8] = np.where(hit, t * np.float32(1.001), np.float32(0.5))
    a = ctx.trace(scene, longer, any_hit=True, flags=capi.PER_RAY_TMAX)
    assert np.all((a[:, 9].view(np.int32) >= 0)[hit])
    shorter = rays.copy()
    shorter[:, 8] = np.where(hit, t * np.float32(0.999), np.float32(0.5))
    b = ctx.trace(scene, shorter, any_hit=True, flags=capi.PER_RAY_TMAX)
    assert not np.any((b[:, 9].view(np.int32) >= 0)[hit])
    again = ctx.trace(scene, rays, t_max=float(t[hit].max()) * 1.01)
    assert np.array_equal(again[hit, 9].view(np.int32), out[hit, 9].view(np.int32))


def test_primary_rays_match_numpy_recipe(ctx):
    """atlas_rt_generate_primary_rays against the numpy statement of rayGen.csh (IDs exact, directions within 1 ulp
    of the float64-normalised recipe) including the 8x8 tile storage order and a ragged border."""
    import ctypes as C
    eye, origin, right, bottom = W.camera_frame((3.0, 2.0, 1.0), (0.0, 0.5, 0.0))
    cam = (C.c_float * 12)(*eye, *origin, *right, *bottom)
    for (w, h) in ((64, 40), (70, 37)):
        out = np.zeros((w * h, 12), dtype=np.float32)
        ctx.check(ctx.L.atlas_rt_generate_primary_rays(ctx.h, cam, w, h, 1, None, out.ctypes.data, 0))
        ids = out[:, 3].view(np.int32)
        assert np.array_equal(np.sort(ids), np.arange(w * h))
        expect = W.primary_rays(w, h, eye, origin, right, bottom)
        by_id = out[np.argsort(ids)]
        assert np.allclose(by_id[:, 4:7], expect[:, 4:7], rtol=0, atol=2e-7)
        assert np.array_equal(by_id[:, 0:3], expect[:, 0:3])
        if w % 8 == 0 and h % 8 == 0:
            tiled = W.primary_rays(w, h, eye, origin, right, bottom, tile_order=True)
            assert np.array_equal(ids, tiled[:, 3].view(np.int32))
        else:
            full = (w // 8) * (h // 8) * 64
            first = ids[:64]
            assert set((first // w).tolist()) == set(range(8)) and set((first % w).tolist()) == set(range(8))
            assert np.all((ids[full:] % w >= (w // 8) * 8) | (ids[full:] // w >= (h // 8) * 8))


def test_leaf_root_blas_is_traversable(ctx, oracle):
    """A BLAS whose root stayed a leaf has zero nodes (Flatten emits none); the reference shader would read out of
    bounds there. The tree must still equal the oracle's, and the library's synthetic root-leaf node must make rays
    test the leaf's triangles: hits equal a brute-force loop over the triangles."""
    tri = np.array([[0, 0, 0, 0, 0, 0, 0, 0, 0]], dtype=np.float32)     # degenerate: nothing to split
    for tris in (tri, ):
        boxes = W.tri_boxes(tris)
        b = ctx.build_blas(boxes, tris)
        nodes, order, eon = b.download()
        o = oracle.build_blas(boxes, tris)
        assert nodes.shape == o.nodes.shape and np.array_equal(order, o.order) and np.array_equal(eon, o.end_of_node)
        b.free()
    # a hand-made root-leaf BLAS over real triangles (what Flatten produces when the "last resort" fires)
    tris = W.uv_sphere(8, 4)
    n = len(tris)
    b = ctx.upload_bvh(np.zeros((0, 14), np.uint32), np.arange(n, dtype=np.uint32), np.r_[np.zeros(n - 1, np.uint8), np.uint8(1)])
    mesh = ctx.pack_mesh(b, tris)
    tl = ctx.build_tlas(np.array([[-1, -1, -1, 1, 1, 1]], dtype=np.float32))
    sc = ctx.create_scene([mesh], W.identity_instance(), tl)
    rays = W.random_rays(2000, [-2, -2, -2], [2, 2, 2], seed=1)
    out = ctx.trace(sc, rays)
    packed = W.pack_bvh_triangles(tris, np.arange(n), np.r_[np.zeros(n - 1), 1])
    osc = OScene(np.zeros((1, 16), np.float32), W.identity_instance(), [np.zeros((1, 16), np.float32)], [packed])
    bt, btri, _ = oracle.brute_force(osc, rays)
    assert np.array_equal(out[:, 8], bt) and np.array_equal(out[:, 9].view(np.int32), btri)
    assert (btri >= 0).mean() > 0.05


def test_streaming_host_buffer_trace(ctx, oracle):
    """ATLAS_RT_TRACE_STREAMING=1: a host-buffer trace runs as ONE persistent launch fed chunk by chunk (upload stream ->
    per-chunk ordering kernels -> watermark), results going home per finished chunk. Every variant must give the bits of
    the device-pointer call (which test_full_size_c2_properties pins to the oracle); a sample is checked against the oracle
    here as well."""
    import os
    import torch
    tris = W.soup(200_000, seed=77)
    boxes = W.tri_boxes(tris)
    root = np.concatenate([boxes[:, :3].min(0), boxes[:, 3:].max(0)])[None].astype(np.float32)
    n = 400_000
    rays = W.random_rays(n, root[0, :3], root[0, 3:], seed=99)
    rays[::1000, 3] = np.int32(-1).view(np.float32)       # dead IDs ...
    rays[5::1000, 4] = np.float32(np.nan)                 # ... and NaN directions finish at fetch time
    rays[:, 8] = np.float32(0.4)                          # per-ray tMax of the any-hit variant
    old = os.environ.get("ATLAS_RT_TRACE_STREAMING")
    os.environ["ATLAS_RT_TRACE_STREAMING"] = "1"
    try:
        sctx = capi.Context(0)
    finally:
        if old is None:
            del os.environ["ATLAS_RT_TRACE_STREAMING"]
        else:
            os.environ["ATLAS_RT_TRACE_STREAMING"] = old
    try:
        scene, osc, keep = gpu_and_oracle_scene(sctx, oracle, [tris], root, W.identity_instance())
        d_rays = torch.from_numpy(rays).cuda()
        d_out = torch.empty_like(d_rays)
        sctx.trace(scene, d_rays, n, out=d_out)                                   # resident batch: the reference bits
        want = d_out.cpu().numpy()
        ref, _ = oracle.trace(osc, rays[:20000], nthreads=8)
        assert np.array_equal(want[:20000].view(np.uint32), ref.view(np.uint32))
        launches = sctx.launches()
        got = sctx.trace(scene, rays)                                             # pageable host memory, whole rays
        # one persistent launch + 4 ordering launches per chunk + the release kernel (the chunked pipeline would need 4 per chunk)
        assert sctx.launches() - launches <= 4 * 4 + 2
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        h_r = torch.from_numpy(rays).pin_memory()
        h_h = torch.empty((n, 4), dtype=torch.float32).pin_memory()
        for _ in range(3):                                                        # pinned memory, 16-byte hit records, repeated calls
            h_h.zero_()
            sctx.check(sctx.L.atlas_rt_trace_closest(sctx.h, scene.h, h_r.data_ptr(), n, capi.MASK_ALL, 0.0, capi.INF, h_h.data_ptr(), capi.HITS_ONLY))
            assert np.array_equal(h_h.numpy().view(np.uint32), want[:, 8:12].view(np.uint32))
        sctx.trace(scene, d_rays, n, out=d_out, any_hit=True, flags=capi.PER_RAY_TMAX)
        want_any = d_out.cpu().numpy()
        got_any = sctx.trace(scene, rays, any_hit=True, flags=capi.PER_RAY_TMAX)
        assert np.array_equal(got_any.view(np.uint32), want_any.view(np.uint32))
        sctx.check(sctx.L.atlas_rt_trace_closest(sctx.h, scene.h, rays.ctypes.data, n, capi.MASK_ALL, 0.0, capi.INF, d_out.data_ptr(), capi.DEVICE_OUTPUT))
        assert np.array_equal(d_out.cpu().numpy().view(np.uint32), want.view(np.uint32))
    finally:
        sctx.close()


def test_pipelined_host_buffer_calls(ctx, oracle):
    """ATLAS_RT_ASYNC | ATLAS_RT_PIPELINED: successive host-buffer trace calls overlap (two persistent staging sets used in turn,
    completion not ordered into the context stream until atlas_rt_trace_join / atlas_rt_context_synchronize). Five calls in a
    row with different rays and alternating result buffers, closest and any-hit, 16-byte records and whole rays: every buffer
    must hold what the device-pointer call gives; a synchronous call issued while pipelined ones are in flight is correct too."""
    import torch
    tris = W.soup(150_000, seed=31)
    boxes = W.tri_boxes(tris)
    root = np.concatenate([boxes[:, :3].min(0), boxes[:, 3:].max(0)])[None].astype(np.float32)
    scene, osc, keep = gpu_and_oracle_scene(ctx, oracle, [tris], root, W.identity_instance())
    n = 300_000
    batches = [W.random_rays(n, root[0, :3], root[0, 3:], seed=500 + k) for k in range(5)]
    for b in batches:
        b[:, 8] = np.float32(0.5)
    d_out = torch.empty((n, 12), dtype=torch.float32, device="cuda")
    want = []
    for k, b in enumerate(batches):
        ctx.trace(scene, torch.from_numpy(b).cuda(), n, out=d_out, any_hit=bool(k & 1), flags=capi.PER_RAY_TMAX if (k & 1) else 0)
        want.append(d_out.cpu().numpy().copy())
    ref, _ = oracle.trace(osc, batches[0][:20000], nthreads=8)
    assert np.array_equal(want[0][:20000].view(np.uint32), ref.view(np.uint32))
    h_in = [torch.from_numpy(b).pin_memory() for b in batches]
    for hits_only in (True, False):
        width = 4 if hits_only else 12
        h_out = [torch.zeros((n, width), dtype=torch.float32).pin_memory() for _ in batches]
        fl = capi.ASYNC | capi.PIPELINED | (capi.HITS_ONLY if hits_only else 0)
        for k in range(5):
            fn = ctx.L.atlas_rt_trace_any if (k & 1) else ctx.L.atlas_rt_trace_closest
            ctx.check(fn(ctx.h, scene.h, h_in[k].data_ptr(), n, capi.MASK_ALL, 0.0, capi.INF, h_out[k].data_ptr(), fl | (capi.PER_RAY_TMAX if (k & 1) else 0)))
        mid = ctx.trace(scene, batches[2])                       # an ordinary synchronous call in between
        assert np.array_equal(mid.view(np.uint32), want[2].view(np.uint32))
        ctx.synchronize()                                        # joins the pipelined calls
        for k in range(5):
            got = h_out[k].numpy()
            exp = want[k][:, 8:12] if hits_only else want[k]
            assert np.array_equal(got.view(np.uint32), exp.view(np.uint32)), (hits_only, k)
