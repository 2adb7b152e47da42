"""CPU-only: the GLSL-order traversal restatement against the reference's own CPU traversal (BVH::GetIntersection,
volume/BVH.cpp:103-217, via oracle/_ref), a brute-force two-level intersector, and committed golden hits."""
import os

import numpy as np
import pytest

from atlas_engine_b200 import workloads as W
from oracle.pyoracle import Scene

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def single_scene(oracle, tris):
    boxes = W.tri_boxes(tris)
    b = oracle.build_blas(boxes, tris)
    root = np.concatenate([boxes[:, :3].min(0), boxes[:, 3:].max(0)])[None].astype(np.float32)
    t = oracle.build_tlas(root)
    return Scene(t.gpu_nodes(), W.identity_instance(), [b.gpu_nodes()], [W.pack_bvh_triangles(tris, b.order, b.end_of_node)]), b, boxes, root[0]


def rays8(rays, tmax):
    n = len(rays)
    return np.concatenate([rays[:, 0:3], rays[:, 4:7], np.zeros((n, 1), np.float32), np.full((n, 1), tmax, np.float32)], axis=1)


@pytest.mark.parametrize("name", ["sphere", "soup", "terrain", "giants"])
def test_closest_and_any_agree_with_reference_cpu_traversal(oracle, ref, name):
    tris = {"sphere": W.uv_sphere(), "soup": W.soup(30000), "terrain": W.heightfield(120, 120),
            "giants": W.soup_with_giants(8000)}[name]
    sc, b, boxes, root = single_scene(oracle, tris)
    pad = (root[3:] - root[:3]) * 0.2
    rays = W.random_rays(20000, root[:3] - pad, root[3:] + pad, seed=42)
    out, ct = oracle.trace(sc, rays, nthreads=4)
    rb = ref.build_blas(boxes, tris, keep=True)
    tuv, idx = ref.intersect_closest(rb, rays8(rays, 1e12), 4)
    hid = out[:, 9].view(np.int32)
    hit = hid >= 0
    assert 0.05 < hit.mean() < 0.95
    src = np.where(hit, b.order[np.maximum(hid, 0)].astype(np.int64), -1)
    assert np.array_equal(src, idx)                               # same source triangle
    assert np.array_equal(out[hit, 8], tuv[hit, 0])               # same t, bit for bit
    assert np.array_equal(out[hit, 7], tuv[hit, 1]) and np.array_equal(out[hit, 11], tuv[hit, 2])
    assert np.all(out[~hit, 8] == np.float32(1e12))
    assert ct["max_stack"] <= 32 and ct["rays_stack_gt32"] == 0
    tmax = np.float32(0.3 * np.linalg.norm(root[3:] - root[:3]))
    anyref = ref.intersect_any(rb, rays8(rays, tmax), 4)
    ra = rays.copy()
    ra[:, 8] = tmax
    oa, _ = oracle.trace(sc, ra, any_hit=True, per_ray_tmax=True, nthreads=4)
    assert np.array_equal(oa[:, 9].view(np.int32) >= 0, anyref > 0)
    ref.free(rb)


def test_two_level_against_brute_force(oracle):
    meshes = [W.uv_sphere(16, 8), W.soup_with_giants(1500, seed=2), W.heightfield(20, 20)]
    mb = [np.concatenate([W.tri_boxes(t)[:, :3].min(0), W.tri_boxes(t)[:, 3:].max(0)]) for t in meshes]
    ib, ir = W.random_instances(300, mb, seed=9, extent=(150.0, 40.0, 150.0))
    ir[::5, 15] = W.MASK_ALL     # some instances do not cast shadows
    bl = [oracle.build_blas(W.tri_boxes(t), t) for t in meshes]
    tl = oracle.build_tlas(ib)
    inst = ir[tl.order].copy()
    sc = Scene(tl.gpu_nodes(), inst, [b.gpu_nodes() for b in bl], [W.pack_bvh_triangles(t, b.order, b.end_of_node) for t, b in zip(meshes, bl)])
    rays = W.random_rays(3000, ib[:, :3].min(0), ib[:, 3:].max(0), seed=5)
    for mask in (W.MASK_ALL, W.MASK_SHADOW):
        out, ct = oracle.trace(sc, rays, cull_mask=mask, nthreads=4)
        bt, btri, binst = oracle.brute_force(sc, rays, cull_mask=mask, nthreads=4)
        assert np.array_equal(out[:, 8], bt)
        hit = out[:, 9].view(np.int32) >= 0
        assert 0.05 < hit.mean() < 0.95
        assert np.array_equal(out[hit, 10].view(np.int32), binst[hit])
        assert ct["max_stack"] <= 32


def test_batch_wrapper_semantics(oracle):
    """traceClosest.csh:18-35: ID < 0 passes through with hitID = -1, t = 0; NaN direction -> miss with t = tMax."""
    sc, b, boxes, root = single_scene(oracle, W.uv_sphere(12, 6))
    rays = W.random_rays(64, root[:3], root[3:], seed=1)
    rays[::4, 3] = np.int32(-1).view(np.float32)
    rays[1, 4] = np.nan
    out, _ = oracle.trace(sc, rays)
    dead = rays[:, 3].view(np.int32) < 0
    assert np.all(out[dead, 8] == 0.0) and np.all(out[dead, 9].view(np.int32) == -1)
    assert out[1, 8] == np.float32(1e12) and out[1, 9].view(np.int32) == -1
    assert np.array_equal(out[:, 0:3], rays[:, 0:3])


def test_oracle_matches_golden_hits(oracle):
    g = np.load(os.path.join(GOLD, "trace_small.npz"))
    sc, b, boxes, root = single_scene(oracle, W.uv_sphere())
    out, _ = oracle.trace(sc, g["rays"])
    hid = out[:, 9].view(np.int32)
    src = np.where(hid >= 0, b.order[np.maximum(hid, 0)].astype(np.int64), -1)
    assert np.array_equal(src, g["ref_idx"])
    hit = hid >= 0
    assert np.array_equal(out[hit, 8], g["ref_tuv"][hit, 0])
