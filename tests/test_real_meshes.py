"""Real geometry from the reference's data directory: data/chromesphere.bin and the three .aemesh
files, copied to tests/golden/meshes/ by tests/golden/make_golden.py together with digests of the trees the UNMODIFIED
reference builder makes from them.

CPU part: the library's .aemesh reader against an independent MessagePack decode, the oracle builder against the
reference-made digests. GPU part (-m gpu): the CUDA builder against both, the device-computed shading words against the
oracle's restatement of MeshData.cpp:176-228, and traversal of the real meshes against the oracle."""
import hashlib
import json
import os

import numpy as np
import pytest

import cases as CS
from atlas_engine_b200 import capi, workloads as W
from oracle.pyoracle import Scene as OScene

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def digest(nodes, order, flags):
    h = hashlib.sha256()
    for a in (nodes, order, flags):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def gold():
    with open(os.path.join(GOLD, "build_hashes.json")) as f:
        return json.load(f)["real"]


def test_aemesh_reader_against_msgpack():
    """atlas_rt_aemesh_* (C++ MessagePack reader + the triangle expansion of MeshData.cpp:89-159) against Python's msgpack
    module and a numpy statement of the same expansion."""
    msgpack = pytest.importorskip("msgpack")
    for name, ntri in (("chromesphere", 1280), ("capsule", 1024), ("metallicwall", 12)):
        m = capi.load_aemesh(CS.aemesh_path(name))
        d = msgpack.unpackb(open(CS.aemesh_path(name), "rb").read(), raw=False)["data"]
        idx = np.frombuffer(d["indices"]["data"], dtype=np.uint32).astype(np.int64)
        pos = np.frombuffer(d["vertices"]["data"], dtype=np.float32).reshape(-1, 3)
        nrm = np.frombuffer(d["normals"]["data"], dtype=np.float32).reshape(-1, 4)
        uv = np.frombuffer(d["texCoords"]["data"], dtype=np.float32).reshape(-1, 2)
        assert m["tris"].shape == (ntri, 9) and m["vertex_count"] == d["vertexCount"] and m["index_count"] == d["indexCount"]
        assert m["materials"] == d["materials"] and m["sub_meshes"] == len(d["subData"])
        assert np.array_equal(m["tris"], pos[idx].reshape(-1, 9))
        assert np.array_equal(m["boxes"], W.tri_boxes(m["tris"]))
        assert np.all(m["material_idx"] == d["subData"][0]["materialIdx"])
        # glm::normalize(vec4): v * (1 / sqrt((x*x + y*y) + (z*z + w*w))), all fp32
        n4 = nrm[idx]
        dd = (n4[:, 0] * n4[:, 0] + n4[:, 1] * n4[:, 1]) + (n4[:, 2] * n4[:, 2] + n4[:, 3] * n4[:, 3])
        inv = np.float32(1.0) / np.sqrt(dd)
        assert np.array_equal(m["normals"], (n4[:, :3] * inv[:, None]).reshape(-1, 9))
        expect_uv = uv[idx].reshape(-1, 6) if len(uv) else np.zeros((ntri, 6), np.float32)
        assert np.array_equal(m["uvs"], expect_uv)
        assert np.all(m["colors"] == 1.0)
    with pytest.raises(capi.AtlasError):
        capi.load_aemesh(os.path.join(GOLD, "build_hashes.json"))     # not MessagePack of the expected schema


def test_oracle_builds_real_meshes_like_the_reference(oracle):
    g = gold()
    small = np.load(os.path.join(GOLD, "build_small.npz"))
    for name, tris in CS.real_mesh_cases().items():
        o = oracle.build_blas(W.tri_boxes(tris), tris)
        assert digest(o.nodes, o.order, o.end_of_node) == g[name], name
        assert np.array_equal(o.order, small["real_" + name + "_order"]), name


def test_shading_words_known_values(oracle):
    """Hand-checked words: axis normals, the 1<<30 w field, half(1.0) = 0x3c00, white = 0xffffffff, and x86's NaN -> 0x80000000
    for the tangent of a triangle without texture coordinates."""
    tri = np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0]], np.float32)
    nrm = np.array([[0, 0, 1, 0, 0, 1, 0, 0, 1]], np.float32)
    uv = np.array([[0, 0, 1, 0, 0, 1]], np.float32)
    w = oracle.pack_shading_words(tri, nrm, uv, None)[0]
    n001 = (511 << 0) | (511 << 10) | (1023 << 20) | (1 << 30)
    assert w[0] == w[1] == w[2] == n001
    assert w[3] == 0 and w[4] == 0x3c00 and w[5] == (0x3c00 << 16)
    assert w[6] == ((1023 << 0) | (511 << 10) | (511 << 20) | (1 << 30))           # tangent = +x
    assert w[7] == ((511 << 0) | (0 << 10) | (511 << 20) | (1 << 30))              # bitangent = -y (handedness -1 * cross(t, n) ... )
    assert w[8] == w[9] == w[10] == 0xffffffff
    w = oracle.pack_shading_words(tri, nrm, None, None)[0]
    assert w[6] == 0xc0000000 and w[7] == 0xc0000000                                 # NaN tangent frame


@pytest.mark.gpu
def test_gpu_builds_real_meshes(ctx, oracle):
    g = gold()
    for name, tris in CS.real_mesh_cases().items():
        boxes = W.tri_boxes(tris)
        b = ctx.build_blas(boxes, tris)
        nodes, order, eon = b.download()
        st = b.stats()
        b.free()
        o = oracle.build_blas(boxes, tris)
        if st["neg_zero"]:     # chromesphere.bin holds -0.0 coordinates: node boxes may differ in the sign of a zero
            assert CS.same_tree_up_to_zero_sign(nodes, order, eon, o), name
        else:
            assert CS.same_tree(nodes, order, eon, o), name
            assert digest(nodes, order, eon) == g[name], name


@pytest.mark.gpu
def test_gpu_shading_words_and_traversal_of_aemesh(ctx, oracle):
    """.aemesh -> triangles -> BLAS -> 48 B + 96 B triangle arrays with device-computed shading words -> closest / any /
    opacity-aware traversal, everything compared with the oracle."""
    for name in ("chromesphere", "capsule", "metallicwall"):
        m = capi.load_aemesh(CS.aemesh_path(name))
        tris, boxes = m["tris"], m["boxes"]
        words = ctx.pack_shading_words(tris, m["normals"], m["uvs"], m["colors"])
        assert np.array_equal(words, oracle.pack_shading_words(tris, m["normals"], m["uvs"], m["colors"])), name
        assert np.array_equal(ctx.pack_shading_words(tris), oracle.pack_shading_words(tris)), name      # all-default attributes
        blas = ctx.build_blas(boxes, tris)
        mesh = ctx.pack_mesh(blas, tris, material_idx=m["material_idx"])
        mesh.pack_shading(tris, material_idx=m["material_idx"], payload11=words)
        o = oracle.build_blas(boxes, tris)
        t96 = W.pack_shading_triangles(tris, o.order, o.end_of_node, material_idx=m["material_idx"], payload11=words)
        assert np.array_equal(mesh.download_shading().view(np.uint32), t96.view(np.uint32)), name
        root = np.concatenate([boxes[:, :3].min(0), boxes[:, 3:].max(0)])[None].astype(np.float32)
        tlas = ctx.build_tlas(root)
        scene = ctx.create_scene([mesh], W.identity_instance(), tlas)
        inst, tnodes = scene.download()
        osc = OScene(tnodes, inst, [o.gpu_nodes()], [W.pack_bvh_triangles(tris, o.order, o.end_of_node, material_idx=int(m["material_idx"][0]))], [t96])
        ext = root[0, 3:] - root[0, :3]
        rays = W.random_rays(40000, root[0, :3] - 0.3 * ext, root[0, 3:] + 0.3 * ext, seed=91)
        for kw, okw in ((dict(), dict()), (dict(any_hit=True), dict(any_hit=True)), (dict(flags=capi.OPACITY), dict(opacity=True))):
            out = ctx.trace(scene, rays, **kw)
            ref, _ = oracle.trace(osc, rays, nthreads=4, **okw)
            assert np.array_equal(out.view(np.uint32), ref.view(np.uint32)), (name, kw)
        assert (out[:, 9].view(np.int32) >= 0).mean() > 0.02, name
        for obj in (scene, tlas, mesh, blas):
            obj.free()
