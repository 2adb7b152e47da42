"""GPU parity at the sizes BASELINE.json names (configs[0], [2], [3]; configs[1] = C2 lives in test_gpu_build.py /
test_gpu_trace.py, configs[4] = C5 in test_gpu_pathtrace.py). Every BLAS / TLAS is compared with the oracle byte for byte
and the traversal bit for bit on a >= 100k-ray sample; the remaining rays are covered by size-independent properties."""
import numpy as np
import pytest

import cases as CS
from atlas_engine_b200 import capi, workloads as W
from oracle.pyoracle import Scene as OScene

pytestmark = pytest.mark.gpu


def single_mesh(ctx, oracle, tris):
    boxes = W.tri_boxes(tris)
    blas = ctx.build_blas(boxes, tris)
    nodes, order, eon = blas.download()
    o = oracle.build_blas(boxes, tris)
    assert CS.same_tree(nodes, order, eon, o)
    root = np.concatenate([boxes[:, :3].min(0), boxes[:, 3:].max(0)])[None].astype(np.float32)
    tlas = ctx.build_tlas(root)
    mesh = ctx.pack_mesh(blas, tris)
    scene = ctx.create_scene([mesh], W.identity_instance(), tlas)
    inst, tnodes = scene.download()
    osc = OScene(tnodes, inst, [o.gpu_nodes()], [W.pack_bvh_triangles(tris, o.order, o.end_of_node)])
    return scene, osc, root[0], (blas, tlas, mesh), o


def sample_equal(oracle, osc, rays, out, idx, **kw):
    ref, _ = oracle.trace(osc, rays[idx], nthreads=16, **kw)
    assert np.array_equal(ref.view(np.uint32), out[idx].view(np.uint32))


def test_c1_atrium_build_and_1080p_primaries(ctx, oracle):
    """configs[0] stand-in (sponza geometry is not in the reference checkout): 354k-triangle atrium whose root takes the
    SBVH spatial split, 1920x1080 primary rays generated on the device in rayGen.csh's tile order."""
    tris = W.atrium(128)
    scene, osc, root, keep, o = single_mesh(ctx, oracle, tris)
    st = keep[0].stats()
    assert st["spatial_chosen"] == 1 and st["duplicates"] == o.stats["duplicates"] > 0
    eye, origin, right, bottom = W.camera_frame((30.0 * 0.05 * 10, 25.0 * 0.05 * 4, 6.0), (600 * 0.05, 3.0, 6.5), fov_deg=47.0)
    rays = ctx.generate_primary_rays(eye, origin, right, bottom, 1920, 1080, 1)
    assert np.array_equal(np.sort(rays[:, 3].view(np.int32)), np.arange(1920 * 1080))
    out = ctx.trace(scene, rays)
    idx = np.r_[0:60000, 1_000_000:1_040_000, len(rays) - 20000:len(rays)]
    sample_equal(oracle, osc, rays, out, idx)
    assert (out[:, 9].view(np.int32) >= 0).mean() > 0.99          # a closed hall: every primary ray hits something
    for obj in (scene,) + keep[1:] + keep[:1]:
        obj.free()


def test_c3_terrain_4k_primaries_and_shadow_rays(ctx, oracle):
    """configs[2]: 8M-triangle heightfield, 3840x2160 primaries, then one shadow (any-hit) ray per hit towards the sun."""
    tris = W.heightfield(2000, 2000)
    scene, osc, root, keep, o = single_mesh(ctx, oracle, tris)
    c = (root[:3] + root[3:]) / 2
    eye = (float(c[0]), float(root[4]) + 60.0, float(c[2]) - 600.0)
    eye_, origin, right, bottom = W.camera_frame(eye, (float(c[0]), float(eye[1]) - 600.0 * np.tan(np.radians(30.0)), float(c[2])), aspect=3840 / 2160)
    rays = ctx.generate_primary_rays(eye_, origin, right, bottom, 3840, 2160, 1)
    out = ctx.trace(scene, rays)
    idx = np.r_[0:50000, 4_000_000:4_050_000, len(rays) - 20000:len(rays)]
    sample_equal(oracle, osc, rays, out, idx)
    hit = out[:, 9].view(np.int32) >= 0
    assert hit.mean() > 0.9
    sun = np.array([0.0, 1.0, 0.33]) / np.linalg.norm([0.0, 1.0, 0.33])
    P = out[:, 0:3] + out[:, 4:7] * out[:, 8:9]
    sh = W.pack_rays((P + np.array([0, 0.1, 0], np.float32)).astype(np.float32), np.broadcast_to(sun.astype(np.float32), P.shape).copy(),
                     ids=np.where(hit, np.arange(len(P)), -1), t=np.full(len(P), 1e12, np.float32))
    occl = ctx.trace(scene, sh, any_hit=True, cull_mask=W.MASK_SHADOW, flags=capi.PER_RAY_TMAX)
    sample_equal(oracle, osc, sh, occl, idx, any_hit=True, per_ray_tmax=True, cull_mask=W.MASK_SHADOW)
    # property over ALL rays: a shadow ray reported occluded must also have a closest hit, and vice versa
    closest = ctx.trace(scene, sh, cull_mask=W.MASK_SHADOW)
    assert np.array_equal(occl[:, 9].view(np.int32) >= 0, closest[:, 9].view(np.int32) >= 0)
    for obj in (scene,) + keep[1:] + keep[:1]:
        obj.free()


def c4_scene():
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
    return meshes, ib, ir


def test_c4_tlas_over_10k_instances(ctx, oracle):
    """configs[3]: 64 BLASes (1k-100k triangles, built as ONE batch), 10k instances, 4M random rays: closest, any-hit and a
    mask that culls a third of the instances."""
    meshes, ib, ir = c4_scene()
    ir[::3, 15] = W.MASK_ALL                       # no shadow bit on every third instance
    blas = ctx.build_blas_batch([W.tri_boxes(t) for t in meshes], meshes)
    obl = [oracle.build_blas(W.tri_boxes(t), t) for t in meshes]
    for b, o in zip(blas, obl):
        n, od, e = b.download()
        assert CS.same_tree(n, od, e, o)
    gm = [ctx.pack_mesh(b, t) for b, t in zip(blas, meshes)]
    tlas = ctx.build_tlas(ib)
    otl = oracle.build_tlas(ib)
    n, od, e = tlas.download()
    assert CS.same_tree(n, od, e, otl)
    scene = ctx.create_scene(gm, ir, tlas)
    inst, tnodes = scene.download()
    osc = OScene(tnodes, inst, [b.gpu_nodes() for b in obl], [W.pack_bvh_triangles(t, b.order, b.end_of_node) for t, b in zip(meshes, obl)])
    rays = W.random_rays(4_000_000, ib[:, :3].min(0), ib[:, 3:].max(0), seed=5678)
    idx = np.r_[0:60000, 2_000_000:2_030_000, len(rays) - 10000:len(rays)]
    out = ctx.trace(scene, rays)
    sample_equal(oracle, osc, rays, out, idx)
    assert 0.2 < (out[:, 9].view(np.int32) >= 0).mean() < 0.6
    sh = rays.copy()
    sh[:, 8] = 200.0
    for mask in (W.MASK_ALL, W.MASK_SHADOW):
        occl = ctx.trace(scene, sh, any_hit=True, cull_mask=mask, flags=capi.PER_RAY_TMAX)
        sample_equal(oracle, osc, sh, occl, idx, any_hit=True, per_ray_tmax=True, cull_mask=mask)
    masked = ctx.trace(scene, rays, cull_mask=W.MASK_SHADOW)
    sample_equal(oracle, osc, rays, masked, idx, cull_mask=W.MASK_SHADOW)
    has_bit = (inst[:, 15] & W.MASK_SHADOW) != 0
    mh = masked[:, 9].view(np.int32) >= 0
    assert np.all(has_bit[masked[mh, 10].view(np.int32)])          # ALL rays: a masked trace never reports a culled instance
    for obj in [scene, tlas] + gm + blas:
        obj.free()
