"""GPU parity tests for the builder, through the C ABI: nodes, primitive order and endOfNode flags must equal the
oracle's (== the reference's) byte for byte."""
import hashlib
import json
import os

import numpy as np
import pytest

import cases as CS
from atlas_engine_b200 import workloads as W

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def digest(nodes, order, flags):
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(nodes).tobytes())
    h.update(np.ascontiguousarray(order).tobytes())
    h.update(np.ascontiguousarray(flags).tobytes())
    return h.hexdigest()


def test_blas_matches_oracle_and_golden(ctx, oracle):
    with open(os.path.join(GOLD, "build_hashes.json")) as f:
        gold = json.load(f)
    for name, tris in CS.build_cases().items():
        boxes = W.tri_boxes(tris)
        b = ctx.build_blas(boxes, tris)
        nodes, order, eon = b.download()
        st = b.stats()
        b.free()
        o = oracle.build_blas(boxes, tris)
        assert CS.same_tree(nodes, order, eon, o), name
        assert digest(nodes, order, eon) == gold["blas"][name], name        # digest produced by the reference itself
        assert st["duplicates"] == o.stats["duplicates"] and st["spatial_chosen"] == o.stats["spatial_chosen"], name
        assert st["median_splits"] == o.stats["median_splits"] and st["sort_fallbacks"] == o.stats["sort_fallbacks"], name


def test_tlas_matches_oracle_and_golden(ctx, oracle):
    with open(os.path.join(GOLD, "build_hashes.json")) as f:
        gold = json.load(f)
    for name, boxes in CS.tlas_cases().items():
        b = ctx.build_tlas(boxes)
        nodes, order, eon = b.download()
        b.free()
        o = oracle.build_tlas(boxes)
        assert CS.same_tree(nodes, order, eon, o), name
        assert digest(nodes, order, eon) == gold["tlas"][name], name


def test_device_resident_input_and_rebuild_determinism(ctx):
    import torch
    tris = W.soup_with_giants(50000, seed=8)
    boxes = W.tri_boxes(tris)
    a = ctx.build_blas(boxes, tris)
    ref_nodes, ref_order, ref_eon = a.download()
    a.free()
    dt, db = torch.from_numpy(tris).cuda(), torch.from_numpy(boxes).cuda()
    for _ in range(3):
        b = ctx.build_blas(db, dt, len(tris))
        nodes, order, eon = b.download()
        b.free()
        assert np.array_equal(nodes, ref_nodes) and np.array_equal(order, ref_order) and np.array_equal(eon, ref_eon)


def check_tree_invariants(nodes, order, eon, boxes):
    """Size-independent properties: every slot is referenced by exactly one leaf pointer, inner pointers form a
    pre-order numbering, every leaf's box in its parent equals... contains the primitive's box, flags all set."""
    n = nodes.shape[0]
    ptr = nodes[:, 12:14].view(np.int32)
    leaf = ptr < 0
    slots = np.sort((~ptr[leaf]).astype(np.int64))
    assert np.array_equal(slots, np.arange(order.shape[0]))
    inner = np.sort(ptr[~leaf].astype(np.int64))
    assert np.array_equal(inner, np.arange(1, n))
    assert np.all(ptr[:, 0][~leaf[:, 0]] == np.arange(n)[~leaf[:, 0]] + 1)      # first child follows its parent
    assert np.all(eon == 1)
    f = nodes[:, :12].view(np.float32).reshape(n, 2, 6)
    for side in (0, 1):
        m = leaf[:, side]
        src = order[(~ptr[m, side]).astype(np.int64)]
        bb = boxes[src]
        assert np.all(f[m, side, :3] <= bb[:, :3]) and np.all(f[m, side, 3:] >= bb[:, 3:])
    area = lambda b: 2 * ((b[:, 3] - b[:, 0]) * (b[:, 4] - b[:, 1]) + (b[:, 4] - b[:, 1]) * (b[:, 5] - b[:, 2]) + (b[:, 5] - b[:, 2]) * (b[:, 3] - b[:, 0]))
    assert np.all(area(f[:, 0]) >= area(f[:, 1]))                                 # larger-area child first


def test_full_size_configs(ctx, oracle):
    """BASELINE sizes. C2 (1M soup) is compared with the oracle directly; the 8M-triangle terrain (C3) through
    structural invariants plus an oracle comparison of a 1/4-size terrain."""
    tris = W.soup(1_000_000, seed=1234)
    boxes = W.tri_boxes(tris)
    b = ctx.build_blas(boxes, tris)
    nodes, order, eon = b.download()
    b.free()
    o = oracle.build_blas(boxes, tris)
    assert CS.same_tree(nodes, order, eon, o)
    check_tree_invariants(nodes, order, eon, boxes)

    tris = W.heightfield(1000, 1000)
    boxes = W.tri_boxes(tris)
    b = ctx.build_blas(boxes, tris)
    nodes, order, eon = b.download()
    b.free()
    o = oracle.build_blas(boxes, tris)
    assert CS.same_tree(nodes, order, eon, o)

    tris = W.heightfield(2000, 2000)
    boxes = W.tri_boxes(tris)
    b = ctx.build_blas(boxes, tris)
    nodes, order, eon = b.download()
    b.free()
    assert nodes.shape[0] == 8_000_000 - 1
    check_tree_invariants(nodes, order, eon, boxes)


def test_bad_arguments(ctx):
    from atlas_engine_b200 import capi
    h = capi.C.c_void_p()
    assert ctx.L.atlas_rt_build_blas(ctx.h, None, None, 5, 0, capi.C.byref(h)) == -1
    assert ctx.L.atlas_rt_build_tlas(ctx.h, None, 5, 0, capi.C.byref(h)) == -1
    empty = ctx.build_blas(np.zeros((0, 6), np.float32), np.zeros((0, 9), np.float32))
    assert empty.counts() == (0, 0)
    empty.free()


def test_wide_level_builder_gives_the_same_trees(oracle):
    """ATLAS_RT_BUILD_WIDE=1 replaces the shared-memory subtree kernel by level kernels over the whole tree (global-memory refs,
    per-size-class node lists; measured slower, kept as an option): every tree of the battery must still carry the reference's
    digest, and the statistics must agree with the oracle."""
    from atlas_engine_b200 import capi
    with open(os.path.join(GOLD, "build_hashes.json")) as f:
        gold = json.load(f)
    old = os.environ.get("ATLAS_RT_BUILD_WIDE")
    os.environ["ATLAS_RT_BUILD_WIDE"] = "1"
    try:
        wctx = capi.Context(0)
    finally:
        if old is None:
            del os.environ["ATLAS_RT_BUILD_WIDE"]
        else:
            os.environ["ATLAS_RT_BUILD_WIDE"] = old
    try:
        for name, tris in CS.build_cases().items():
            boxes = W.tri_boxes(tris)
            b = wctx.build_blas(boxes, tris)
            nodes, order, eon = b.download()
            st = b.stats()
            b.free()
            assert digest(nodes, order, eon) == gold["blas"][name], name
            o = oracle.build_blas(boxes, tris)
            assert st["median_splits"] == o.stats["median_splits"] and st["sort_fallbacks"] == o.stats["sort_fallbacks"], name
        for name, boxes in CS.tlas_cases().items():
            b = wctx.build_tlas(boxes)
            nodes, order, eon = b.download()
            b.free()
            assert digest(nodes, order, eon) == gold["tlas"][name], name
        tris = W.soup(300_000, seed=21)
        a = wctx.build_blas(W.tri_boxes(tris), tris)
        o = oracle.build_blas(W.tri_boxes(tris), tris)
        assert CS.same_tree(*a.download(), o)
        a.free()
    finally:
        wctx.close()
