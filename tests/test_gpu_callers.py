"""GPU parity for the engine's OTHER callers of the traversal: RTAO (ao/rtao.csh:81-100, short any-hit rays),
RT reflections (reflection/rtreflection.csh:104-141, GGX-VNDF directions), RTGI (rtgi/rtgi.csh:110-137, cosine directions)
and DDGI probe rays (ddgi/rayGen.csh:44-83, spherical Fibonacci sets with dead slots). Same scene structures, different ray
distributions; every batch is compared with the oracle bit for bit, plain and opacity-aware (OPACITY_CHECK) variants."""
import numpy as np
import pytest

import cases as CS
from atlas_engine_b200 import capi, workloads as W
from oracle.pyoracle import Scene as OScene

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hall(ctx, oracle):
    """The atrium (single mesh) with shading triangles; G-buffer from 640x360 primary hits."""
    tris = W.atrium(48)
    boxes = W.tri_boxes(tris)
    blas = ctx.build_blas(boxes, tris)
    o = oracle.build_blas(boxes, tris)
    n, od, e = blas.download()
    assert CS.same_tree(n, od, e, o)
    mesh = ctx.pack_mesh(blas, tris)
    nrm = W.smooth_normals(tris)
    mesh.pack_shading(tris, payload11=ctx.pack_shading_words(tris, nrm))
    root = np.concatenate([boxes[:, :3].min(0), boxes[:, 3:].max(0)])[None].astype(np.float32)
    tlas = ctx.build_tlas(root)
    scene = ctx.create_scene([mesh], W.identity_instance(), tlas)
    inst, tnodes = scene.download()
    osc = OScene(tnodes, inst, [o.gpu_nodes()], [W.pack_bvh_triangles(tris, o.order, o.end_of_node)], [mesh.download_shading()])
    cam = W.camera_frame((30.0 * 0.05 * 10, 25.0 * 0.05 * 4, 6.0), (600 * 0.05, 3.0, 6.5), fov_deg=47.0)
    prim = ctx.trace(scene, ctx.generate_primary_rays(*cam, 640, 360, 1))
    P, V, dist, hit = W.gbuffer_from_hits(prim)
    order = o.order[prim[hit, 9].view(np.int32)]
    T = tris[order].reshape(-1, 3, 3).astype(np.float64)
    N = np.cross(T[:, 1] - T[:, 0], T[:, 2] - T[:, 0])
    N /= np.maximum(np.linalg.norm(N, axis=1, keepdims=True), 1e-30)
    N *= np.where((N * V).sum(1, keepdims=True) < 0, -1.0, 1.0)          # face the camera
    return scene, osc, P, N, V, dist, root[0], (blas, mesh, tlas)


def same(ctx, oracle, scene, osc, rays, **kw):
    okw = {k: v for k, v in kw.items() if k != "flags"}
    okw["opacity"] = bool(kw.get("flags", 0) & capi.OPACITY)
    okw["per_ray_tmax"] = bool(kw.get("flags", 0) & capi.PER_RAY_TMAX)
    out = ctx.trace(scene, rays, **kw)
    ref, _ = oracle.trace(osc, rays, nthreads=8, **okw)
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))
    return out


def test_rtao(ctx, oracle, hall):
    scene, osc, P, N, V, dist, root, _ = hall
    u = np.random.default_rng(1).random((len(P), 2))
    rays = W.rtao_rays(P, N, u, radius=1.5)
    out = same(ctx, oracle, scene, osc, rays, any_hit=True, cull_mask=W.MASK_ALL, flags=capi.PER_RAY_TMAX)
    occluded = (out[:, 9].view(np.int32) >= 0).mean()
    assert 0.01 < occluded < 0.6
    tr = same(ctx, oracle, scene, osc, rays, any_hit=True, cull_mask=W.MASK_ALL, flags=capi.PER_RAY_TMAX | capi.OPACITY)   # OPACITY_CHECK variant
    assert np.array_equal(tr[:, 7] == 0.0, out[:, 9].view(np.int32) >= 0)          # all-opaque scene: transparency 0 <=> HitAny hit


def test_rt_reflections(ctx, oracle, hall):
    scene, osc, P, N, V, dist, root, _ = hall
    u = np.random.default_rng(2).random((len(P), 2))
    for rough in (0.0, 0.3, 0.9):
        rays = W.reflection_rays(P, N, V, rough, u, dist)
        out = same(ctx, oracle, scene, osc, rays)
        live = rays[:, 3].view(np.int32) >= 0
        assert live.mean() > 0.5 and (out[live, 9].view(np.int32) >= 0).mean() > 0.9       # a closed hall: reflections hit something (rough lobes lose rays below the surface)
        assert np.all(out[~live, 9].view(np.int32) == -1) and np.all(out[~live, 8] == 0.0)   # rays not cast pass through
        same(ctx, oracle, scene, osc, rays, flags=capi.OPACITY)


def test_rtgi(ctx, oracle, hall):
    scene, osc, P, N, V, dist, root, _ = hall
    u = np.random.default_rng(3).random((len(P), 2))
    rays = W.rtgi_rays(P, N, u, dist, bias=0.1)
    out = same(ctx, oracle, scene, osc, rays)
    assert (out[:, 9].view(np.int32) >= 0).mean() > 0.9
    same(ctx, oracle, scene, osc, rays, flags=capi.OPACITY)


def test_ddgi_probe_rays(ctx, oracle, hall):
    scene, osc, P, N, V, dist, root, _ = hall
    rays = W.ddgi_rays(root[:3], root[3:], probes=(12, 6, 12), rays_per_probe=128)
    assert len(rays) == 12 * 6 * 12 * 128
    out = same(ctx, oracle, scene, osc, rays)
    dead = rays[:, 3].view(np.int32) < 0
    assert 0.05 < dead.mean() < 0.2 and np.all(out[dead, 9].view(np.int32) == -1)
    assert (out[~dead, 9].view(np.int32) >= 0).mean() > 0.4      # the probe grid spans the scene box, which the beams stretch beyond the hall
