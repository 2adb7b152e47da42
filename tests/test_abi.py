"""CPU-only: the C-ABI library loads and exports exactly what include/atlas_rt.h declares; host-side helpers."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    from atlas_engine_b200 import capi
    return capi


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "atlas_rt.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(atlas_rt_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(built):
    lib = ctypes.CDLL(built.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/atlas_rt.h but not exported by libatlas_rt.so"


def test_binding_covers_header(built):
    assert sorted(built.SIGNATURES) == declared_symbols()


def test_version_and_no_gpu_behaviour(built):
    L = built.lib()
    assert L.atlas_rt_version() == 1
    import torch
    if not torch.cuda.is_available():
        # no CPU fallback: creating a context without a device must fail loudly
        with pytest.raises(built.AtlasError):
            built.Context(0)


def test_shard_range_matches_python(built):
    from atlas_engine_b200 import sharding
    for count in (0, 1, 63, 64, 65, 1000, 1_000_003, 8_294_400):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for rank in range(world):
                b, e = built.shard_range(count, rank, world, 64)
                assert (b, e) == sharding.shard_bounds(count, rank, world, 64)
                assert b == prev and b <= e
                assert b % 64 == 0 or b == count
                prev = e
            assert prev == count


def test_layout_sizes():
    text = open(os.path.join(ROOT, "atlas_engine_b200", "csrc", "layouts.h")).read()
    for name, size in (("HostAABB", 24), ("HostBVHNode", 56), ("GPUBVHNode", 64), ("GPUBVHTriangle", 48), ("GPUBVHInstance", 64), ("PackedRay", 48)):
        assert f"sizeof({name}) == {size}" in text


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under atlas_engine_b200/ (Python, C++ or CUDA) may import, include,
    load or execute it, and the Python binding must have no CPU fallback (it raises when the library is missing)."""
    import os
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "atlas_engine_b200")
    offenders = []
    for dirpath, _, files in os.walk(root):
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                continue
            text = open(os.path.join(dirpath, f), errors="replace").read()
            if re.search(r"import\s+oracle|from\s+oracle|oracle/|pyoracle|libatlas_oracle|libatlas_ref", text):
                offenders.append(os.path.relpath(os.path.join(dirpath, f), root))
    assert offenders == []
    src = open(os.path.join(root, "capi.py")).read()
    assert "raise" in src and "libatlas_rt.so" in src
