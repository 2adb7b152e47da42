"""CPU-only: the oracle restatement (oracle/atlas_oracle.cpp) against (a) the unmodified reference compiled into
oracle/_ref and (b) the committed golden fixtures generated from that reference (tests/golden/make_golden.py)."""
import hashlib
import json
import os

import numpy as np
import pytest

import cases as CS
from atlas_engine_b200 import workloads as W

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def digest(tree):
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(tree.nodes).tobytes())
    h.update(np.ascontiguousarray(tree.order).tobytes())
    h.update(np.ascontiguousarray(tree.end_of_node).tobytes())
    return h.hexdigest()


def test_oracle_blas_matches_reference(oracle, ref):
    for name, tris in CS.build_cases().items():
        boxes = W.tri_boxes(tris)
        a = ref.build_blas(boxes, tris, parallel=True)
        b = oracle.build_blas(boxes, tris)
        assert CS.same_tree(a.nodes, a.order, a.end_of_node, b), name


def test_reference_parallel_equals_serial(ref):
    for name in ("giants", "atrium", "heightfield", "coincident"):
        tris = CS.build_cases()[name]
        boxes = W.tri_boxes(tris)
        a = ref.build_blas(boxes, tris, parallel=True)
        b = ref.build_blas(boxes, tris, parallel=False)
        assert CS.same_tree(a.nodes, a.order, a.end_of_node, b), name


def test_oracle_tlas_matches_reference(oracle, ref):
    for name, boxes in CS.tlas_cases().items():
        a = ref.build_tlas(boxes)
        b = oracle.build_tlas(boxes)
        assert CS.same_tree(a.nodes, a.order, a.end_of_node, b), name


def test_tlas_edge_cases(oracle):
    one = oracle.build_tlas(np.array([[0, 0, 0, 1, 2, 3]], dtype=np.float32))
    assert one.nodes.shape[0] == 1 and list(one.order) == [0, 0] and list(one.end_of_node) == [0, 1]
    assert one.nodes[0, 12:].view(np.int32).tolist() == [-1, -1]          # leftPtr = rightPtr = ~0
    two = oracle.build_tlas(np.array([[0, 0, 0, 1, 1, 1], [2, 0, 0, 5, 4, 4]], dtype=np.float32))
    assert two.nodes.shape[0] == 1 and sorted(two.nodes[0, 12:].view(np.int32).tolist()) == [-2, -1]
    assert two.order[0] == 1                                                # larger-area child first


def test_oracle_matches_golden_hashes(oracle):
    """Golden digests were produced by the REFERENCE (make_golden.py); this pins the oracle even where
    /root/reference and oracle/_ref are absent."""
    with open(os.path.join(GOLD, "build_hashes.json")) as f:
        gold = json.load(f)
    blas, tl = CS.build_cases(), CS.tlas_cases()
    assert set(gold["blas"]) == set(blas) and set(gold["tlas"]) == set(tl)
    for name, tris in blas.items():
        assert digest(oracle.build_blas(W.tri_boxes(tris), tris)) == gold["blas"][name], name
    for name, boxes in tl.items():
        assert digest(oracle.build_tlas(boxes)) == gold["tlas"][name], name


def test_oracle_matches_golden_arrays(oracle):
    g = np.load(os.path.join(GOLD, "build_small.npz"))
    for name in ("sphere", "soup33", "identical5", "coincident"):
        tris = CS.build_cases()[name]
        t = oracle.build_blas(W.tri_boxes(tris), tris)
        assert np.array_equal(t.nodes, g[name + "_nodes"]) and np.array_equal(t.order, g[name + "_order"])
        assert np.array_equal(t.end_of_node, g[name + "_flags"])


def test_packed_normal_word_matches_the_reference_function(ref, oracle):
    """Common::Packing::PackSignedVector3x10_1x2 (common/Packing.cpp:24-35) makes the packed normals / tangent / bitangent of
    GPUTriangle (mesh/MeshData.cpp:205-210). The reference's own compiled function (oracle/_ref) against the restatement inside
    oracle/atlas_oracle_shade.cpp: unit vectors, the whole [-1, 1] cube, out-of-range values, zeros, infinities and NaNs (the
    float -> int conversion of the x86 build: 0x80000000)."""
    rng = np.random.default_rng(17)
    v = rng.uniform(-1.0, 1.0, (400_000, 4)).astype(np.float32)
    n = rng.normal(size=(200_000, 3)).astype(np.float32)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    v[:200_000, :3] = n
    v[:200_000, 3] = 0.0
    v[200_000:200_500] = rng.uniform(-3.0, 3.0, (500, 4)).astype(np.float32)
    special = np.array([[1, 1, 1, 1], [-1, -1, -1, -1], [0, 0, 0, 0], [-0.0, 0.0, -0.0, 0.0], [np.nan, 0.5, -0.5, 0.0], [0.25, np.nan, np.nan, np.nan],
                        [np.inf, -np.inf, 1e30, -1e30], [1.0 - 2 ** -24, -1.0 + 2 ** -24, 2 ** -126, -(2 ** -126)]], dtype=np.float32)
    v[-len(special):] = special
    a, b = ref.pack_signed(v), oracle.pack_signed(v)
    assert np.array_equal(a, b)
    assert int(a[-len(special) + 4]) & 0x3ff == 0 and (int(a[-len(special) + 6]) & 0xffffffff) == 0x80000000   # the x86 conversion of NaN / inf
