"""CPU: the shading half of the oracle (oracle/atlas_oracle_shade.cpp — rayGen.csh, rayHit.csh, random.hsh, octahedral bins,
GetOpacity) has no reference-made golden vector (the shaders cannot run here: no Vulkan), so it is cross-checked against
independent numpy statements of the same recipes: hash RNG words, half packing, primary rays, direction bins, bilinear
opacity, and physical invariants of the bounce (misses add the sky, white-furnace energy bound, payload round trip)."""
import numpy as np
import pytest

from atlas_engine_b200 import capi, workloads as W
from oracle.pyoracle import Scene as OScene

f32 = np.float32


def hash1(x):
    x = np.asarray(x, dtype=np.uint32).copy()
    with np.errstate(over="ignore"):
        x += x << np.uint32(10); x ^= x >> np.uint32(6); x += x << np.uint32(3); x ^= x >> np.uint32(11); x += x << np.uint32(15)
    return x


def random2(x, y):
    m = hash1(np.asarray(x, f32).view(np.uint32) ^ hash1(np.asarray(y, f32).view(np.uint32)))
    return ((m & np.uint32(0x007FFFFF)) | np.uint32(0x3F800000)).view(f32) - f32(1.0)


def simple_scene(oracle, materials=None, textures=()):
    tris = W.uv_sphere(16, 8)
    boxes = W.tri_boxes(tris)
    ob = oracle.build_blas(boxes, tris)
    root = np.concatenate([boxes[:, :3].min(0), boxes[:, 3:].max(0)])[None].astype(f32)
    ot = oracle.build_tlas(root)
    words = oracle.pack_shading_words(tris, W.smooth_normals(tris), W.planar_uvs(tris), None)
    t96 = W.pack_shading_triangles(tris, ob.order, ob.end_of_node, payload11=words)
    inst = np.concatenate([W.identity_instance(), W.identity_instance()])
    osc = OScene(ot.gpu_nodes(), inst, [ob.gpu_nodes()], [W.pack_bvh_triangles(tris, ob.order, ob.end_of_node)], [t96])
    osc.set_materials(capi.make_materials(1) if materials is None else materials, textures)
    return osc


def test_jitter_and_rng_words():
    for sc in (0, 1, 7, 123456):
        j = capi.sample_jitter(sc)
        assert j[0] == random2(f32(sc), f32(0.0)) and j[1] == random2(f32(sc), f32(1.0))
    assert hash1(np.uint32(0)) == 0 and hash1(np.uint32(1)) == 307143837      # one-at-a-time hash of 1, worked by hand from random.hsh:5-12
    v = random2(np.arange(1000, dtype=f32), np.full(1000, 3.0, f32))
    assert v.min() >= 0.0 and v.max() < 1.0 and 0.4 < v.mean() < 0.6


def test_raygen_against_numpy_recipe(oracle):
    eye, origin, right, bottom = W.camera_frame((3.0, 2.0, 1.0), (0.0, 0.5, 0.0))
    for (w, h, sc) in ((64, 40, 0), (24, 16, 9)):
        rays = oracle.raygen(eye, origin, right, bottom, w, h, 1, sc)
        jit = capi.sample_jitter(sc)
        expect = W.primary_rays(w, h, eye, origin, right, bottom, jitter=(jit[0], jit[1]), tile_order=True)
        assert np.array_equal(rays[:, 3].view(np.int32), expect[:, 3].view(np.int32))          # IDs and the 8x8 tile storage order
        assert np.allclose(rays[:, 4:7], expect[:, 4:7], rtol=0, atol=3e-7)
        assert np.array_equal(rays[:, 0:3], expect[:, 0:3])
    ragged = oracle.raygen(eye, origin, right, bottom, 13, 9, 2, 1)                                # ragged borders, 2 samples
    ids = ragged[:, 3].view(np.int32)
    assert np.array_equal(np.sort(ids), np.arange(13 * 9 * 2))
    assert np.array_equal(ids[:128:2] // 2 % 13 < 8, np.ones(64, bool)) and np.array_equal(ids[0:128:2] + 1, ids[1:128:2])


def test_direction_bins_against_numpy(oracle):
    rng = np.random.default_rng(3)
    d = rng.normal(size=(20000, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = d.astype(f32)
    bins = oracle.ray_bins(W.pack_rays(np.zeros_like(d), d))
    x, y, z = d[:, 0].astype(np.float64), d[:, 1].astype(np.float64), d[:, 2].astype(np.float64)
    l1 = np.abs(x) + np.abs(y) + np.abs(z)
    x, z = x / l1, z / l1
    lower = y < 0
    ox, oz = x.copy(), z.copy()
    x = np.where(lower, np.where(ox >= 0, 1.0, -1.0) * (1 - np.abs(oz)), x)
    z = np.where(lower, np.where(oz >= 0, 1.0, -1.0) * (1 - np.abs(ox)), z)
    cx, cz = np.clip(0.5 * x + 0.5, 0, 1), np.clip(0.5 * z + 0.5, 0, 1)
    expect = (cz * 8).astype(np.int64) * 8 + (cx * 8).astype(np.int64)
    agree = (bins == expect).mean()
    assert agree > 0.999                       # float32 vs float64 only differs for directions on a bin boundary
    assert bins.max() <= 72 and len(np.unique(bins)) >= 60


def test_opacity_texture_sampling_against_numpy(oracle):
    rng = np.random.default_rng(4)
    tex = (rng.random((9, 13)) * 255).astype(np.uint8)
    mats = capi.make_materials(1)
    mats[0]["opacityTexture"] = 0
    mats[0]["opacity"] = 0.75
    tri = np.zeros(24, f32)
    uv = np.array([[0.1, 0.2], [1.7, -0.4], [0.3, 2.2]], f32)      # outside [0,1): repeat addressing
    halves = uv.astype(np.float16)
    tri[12:15] = (halves[:, 0].view(np.uint16).astype(np.uint32) | (halves[:, 1].view(np.uint16).astype(np.uint32) << 16)).view(f32)
    tri[23] = -1.0
    mat_words = np.ascontiguousarray(mats).view(np.uint32)
    dims = np.array([[13, 9]], np.uint32)
    import ctypes as C
    ptrs = (C.c_void_p * 1)(tex.ctypes.data)
    for (s, t) in ((0.2, 0.3), (0.0, 0.0), (0.5, 0.5), (0.9, 0.05)):
        got = oracle.lib.oracle_get_opacity(tri.ctypes.data_as(C.c_void_p), s, t, mat_words.ctypes.data_as(C.c_void_p), dims.ctypes.data_as(C.c_void_p), ptrs, 1)
        r = 1.0 - s - t
        u, v = (r * halves[0].astype(np.float64) + s * halves[1].astype(np.float64) + t * halves[2].astype(np.float64))
        x, y = u * 13 - 0.5, v * 9 - 0.5
        x0, y0 = int(np.floor(x)), int(np.floor(y))
        wx, wy = x - x0, y - y0
        tx = lambda xx, yy: tex[yy % 9, xx % 13] / 255.0
        expect = ((tx(x0, y0) * (1 - wx) + tx(x0 + 1, y0) * wx) * (1 - wy) + (tx(x0, y0 + 1) * (1 - wx) + tx(x0 + 1, y0 + 1) * wx) * wy) * 0.75
        assert abs(got - expect) < 2e-5


def test_bounce_invariants(oracle):
    mats = capi.make_materials(1)
    mats[0]["baseR"] = mats[0]["baseG"] = mats[0]["baseB"] = 1.0       # white furnace
    osc = simple_scene(oracle, mats)
    cam = W.camera_frame((0.0, 0.5, -4.0), (0.0, 0.0, 0.0), aspect=1.0)
    rays = oracle.raygen(*cam, 48, 48, 1, 0)
    hits, _ = oracle.trace(osc, rays, opacity=True)
    prm = capi.pt_params((0.3, 0.9, -0.3), (3.0, 3.0, 2.5), (0.4, 0.5, 0.8), max_bounces=3)
    sh = oracle.pt_shadow_rays(osc, hits, prm)
    vis, _ = oracle.trace(osc, sh, any_hit=True, cull_mask=W.MASK_SHADOW, t_max=1e12, opacity=True)
    st = oracle.pt_shade(osc, hits, None, vis[:, 7], prm, 2.5, 0)
    hit = hits[:, 9].view(np.int32) >= 0
    assert 0.05 < hit.mean() < 0.5
    # misses: the path ends with min(sky * throughput, 10)
    assert not st["alive"][~hit].any() and np.allclose(st["finished"][~hit], [0.4, 0.5, 0.8])
    # survivors: unit directions leaving the surface, origin near the hit point, hit fields cleared, payload halves decode
    a = st["alive"]
    assert a.sum() > 0.5 * hit.sum()
    assert np.allclose(np.linalg.norm(st["rays"][a, 4:7], axis=1), 1.0, atol=1e-5)
    P = hits[a, 0:3] + hits[a, 4:7] * hits[a, 8:9]
    assert np.allclose(st["rays"][a, 0:3], P, atol=0.11)
    assert np.all(st["rays"][a, 9].view(np.int32) == -1)
    w = st["payload"][a]
    thr = np.stack([(w[:, 1] & 0xffff).astype(np.uint16).view(np.float16), (w[:, 1] >> 16).astype(np.uint16).view(np.float16),
                    (w[:, 2] >> 16).astype(np.uint16).view(np.float16)], 1).astype(np.float64)
    assert np.isfinite(thr).all() and thr.min() >= 0.0 and thr.max() < 20.0
    # Russian roulette bookkeeping: survivors drew below their probability
    assert np.all(st["rr"][a, 0] <= st["rr"][a, 1])
    # the same inputs give the same outputs (pure function of ray.ID and seed); another seed changes the draws
    st2 = oracle.pt_shade(osc, hits, None, vis[:, 7], prm, 2.5, 0)
    assert np.array_equal(st["rays"].view(np.uint32), st2["rays"].view(np.uint32))
    st3 = oracle.pt_shade(osc, hits, None, vis[:, 7], prm, 3.5, 0)
    assert not np.array_equal(st["rr"], st3["rr"])
    # last bounce: nothing survives, every live ray finishes
    last = oracle.pt_shade(osc, hits, st["payload"], vis[:, 7], prm, 2.5, 3)
    assert not last["alive"].any()
