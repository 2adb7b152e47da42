"""GPU parity of the opacity-aware traversal variants (HitClosestTransparency / HitAnyTransparency, bvh.hsh:275-357,
443-524) and of the 96-byte GPUTriangle packing, through the C ABI, against the oracle."""
import numpy as np
import pytest

from atlas_engine_b200 import capi, workloads as W
from oracle.pyoracle import Scene as OScene

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def world(ctx, oracle):
    rng = np.random.default_rng(12)
    meshes = [W.uv_sphere(20, 10), W.soup(4000, seed=4, extent=0.15) * np.float32(4.0), W.heightfield(30, 30) * np.float32(0.3)]
    opac, mats, payload = [], [], []
    for t in meshes:
        o = rng.choice(np.array([1.0, 0.5, 0.25, 0.0, -1.0], dtype=np.float32), size=len(t), p=[0.4, 0.25, 0.15, 0.15, 0.05])
        opac.append(o)
        mats.append(rng.integers(0, 7, size=len(t)).astype(np.int32))
        payload.append(rng.integers(0, 2**32, size=(len(t), 11), dtype=np.uint64).astype(np.uint32))
    mb = [np.concatenate([W.tri_boxes(t)[:, :3].min(0), W.tri_boxes(t)[:, 3:].max(0)]) for t in meshes]
    ib, ir = W.random_instances(400, mb, seed=21, extent=(60.0, 15.0, 60.0), scale=(0.8, 2.5))
    blas = [ctx.build_blas(W.tri_boxes(t), t) for t in meshes]
    gm = [ctx.pack_mesh(b, t, material_idx=m, opacity=o) for b, t, m, o in zip(blas, meshes, mats, opac)]
    for g, t, m, o, p in zip(gm, meshes, mats, opac, payload):
        g.pack_shading(t, m, o, p)
    tl = ctx.build_tlas(ib)
    scene = ctx.create_scene(gm, ir, tl)
    inst, tnodes = scene.download()
    obl = [oracle.build_blas(W.tri_boxes(t), t) for t in meshes]
    t96 = [W.pack_shading_triangles(t, b.order, b.end_of_node, m, o, p) for t, b, m, o, p in zip(meshes, obl, mats, opac, payload)]
    t48 = [W.pack_bvh_triangles(t, b.order, b.end_of_node) for t, b in zip(meshes, obl)]
    osc = OScene(tnodes, inst, [b.gpu_nodes() for b in obl], t48, t96)
    return scene, osc, ib, gm, t96


def test_shading_triangle_layout(world):
    scene, osc, ib, gm, t96 = world
    for g, ref in zip(gm, t96):
        assert np.array_equal(g.download_shading().view(np.uint32), ref.view(np.uint32))


def test_closest_with_opacity(ctx, oracle, world):
    scene, osc, ib, gm, t96 = world
    rays = W.random_rays(150000, ib[:, :3].min(0), ib[:, 3:].max(0), seed=8)
    out = ctx.trace(scene, rays, flags=capi.OPACITY | capi.COUNTERS)
    gc = ctx.trace_counters()
    ref, oc = oracle.trace(osc, rays, opacity=True, nthreads=8)
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))
    assert all(gc[k] == oc[k] for k in oc)
    plain = ctx.trace(scene, rays)
    assert not np.array_equal(plain[:, 9].view(np.int32), out[:, 9].view(np.int32))   # fully transparent triangles are skipped
    assert 0.1 < (ref[:, 9].view(np.int32) >= 0).mean() < 0.9


def test_any_with_transparency(ctx, oracle, world):
    scene, osc, ib, gm, t96 = world
    rays = W.random_rays(150000, ib[:, :3].min(0), ib[:, 3:].max(0), seed=9)
    rays[:, 8] = 40.0
    for mask in (W.MASK_ALL, W.MASK_SHADOW):
        out = ctx.trace(scene, rays, any_hit=True, cull_mask=mask, flags=capi.OPACITY | capi.PER_RAY_TMAX)
        ref, _ = oracle.trace(osc, rays, any_hit=True, per_ray_tmax=True, cull_mask=mask, opacity=True, nthreads=8)
        assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))
    tr = ref[:, 7]
    assert (tr == 0).any() and (tr == 1).any() and ((tr > 0) & (tr < 1)).any()


def test_opaque_scene_degenerates_to_plain_variant(ctx):
    tris = W.soup(20000, seed=6, extent=0.05)
    boxes = W.tri_boxes(tris)
    root = np.concatenate([boxes[:, :3].min(0), boxes[:, 3:].max(0)])[None].astype(np.float32)
    b = ctx.build_blas(boxes, tris)
    m = ctx.pack_mesh(b, tris)
    tl = ctx.build_tlas(root)
    with pytest.raises(capi.AtlasError):     # no 96-byte triangles yet
        ctx.trace(ctx.create_scene([m], W.identity_instance(), tl), W.random_rays(64, root[0, :3], root[0, 3:]), flags=capi.OPACITY)
    m.pack_shading(tris)
    sc = ctx.create_scene([m], W.identity_instance(), tl)
    rays = W.random_rays(100000, root[0, :3], root[0, 3:], seed=2)
    assert np.array_equal(ctx.trace(sc, rays).view(np.uint32), ctx.trace(sc, rays, flags=capi.OPACITY).view(np.uint32))
