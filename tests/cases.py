"""Shared input batteries for the parity tests (same seeds on CPU and GPU)."""
import os

import numpy as np

from atlas_engine_b200 import workloads as W


def build_cases(big=False):
    """name -> triangles (n, 9). Covers SURVEY.md §8a.1: tiny n, duplicates, zero-area and axis-flat triangles, the
    median/sort fallback with small and large n, the root spatial split, skipped axes."""
    cases = {}
    for n in list(range(1, 34)) + [64, 100, 257]:
        cases[f"soup{n}"] = W.soup(n, seed=n)
    for n in (1, 2, 3, 5, 17, 100):
        t = W.soup(n, seed=100 + n)
        t[:, :] = t[0]
        cases[f"identical{n}"] = t
    cases["soup1000"] = W.soup(1000, seed=1000)
    cases["soup5000"] = W.soup(5000, seed=5000)
    cases["coincident"] = W.coincident(200, 200)
    cases["coincident_big"] = W.coincident(3000, 3000)
    cases["flat_grid"] = W.flat_grid(60)
    cases["heightfield"] = W.heightfield(150, 150)
    cases["sphere"] = W.uv_sphere()
    cases["giants"] = W.soup_with_giants(20000)
    cases["atrium"] = W.atrium(32)
    t = W.soup(500, seed=3)
    t[::3, 3:6] = t[::3, 0:3]
    t[::5, 1] = t[::5, 4] = t[::5, 7] = 0.25
    cases["degenerate"] = t
    cases["wide_soup"] = W.soup(30000, seed=77, extent=0.3)
    if big:
        cases["soup100k"] = W.soup(100000)
        cases["soup1m"] = W.soup(1000000)
        cases["heightfield1m"] = W.heightfield(707, 707)
        cases["atrium_big"] = W.atrium(128)
        cases["giants200k"] = W.soup_with_giants(200000, seed=5)
    return cases


def chromesphere_bin(path=None):
    """data/chromesphere.bin of the reference (copied to tests/golden/meshes by make_golden.py): 3840 float3 positions at
    byte 0 and 3840 u16 indices at byte 122880 (chromesphere.gltf bufferViews 0 / 3) -> 1280 triangles."""
    path = path or os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "meshes", "chromesphere.bin")
    raw = open(path, "rb").read()
    pos = np.frombuffer(raw[:3840 * 12], dtype=np.float32).reshape(-1, 3)
    idx = np.frombuffer(raw[122880:122880 + 3840 * 2], dtype=np.uint16).astype(np.int64)
    return pos[idx].reshape(-1, 9).astype(np.float32)


def aemesh_path(name):
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "meshes", name + ".aemesh")


def real_mesh_cases():
    """Real geometry from the reference's data directory: the glTF sphere's raw buffer and the three .aemesh files read
    through the library's own reader (atlas_rt_aemesh_*)."""
    from atlas_engine_b200 import capi
    cases = {"chromesphere_bin": chromesphere_bin()}
    for n in ("chromesphere", "capsule", "metallicwall"):
        cases["aemesh_" + n] = capi.load_aemesh(aemesh_path(n))["tris"]
    return cases


def tlas_cases():
    cases = {}
    for m in (1, 2, 3, 4, 7, 33, 100, 1025, 10000):
        cases[f"tlas{m}"] = W.tri_boxes(W.soup(m, seed=200 + m, extent=0.2))
    same = np.tile(np.array([[0, 0, 0, 1, 1, 1]], dtype=np.float32), (50, 1))
    cases["tlas_same50"] = same
    return cases


def same_tree(a_nodes, a_order, a_flags, b):
    return (a_nodes.shape == b.nodes.shape and np.array_equal(a_nodes, b.nodes) and np.array_equal(a_order, b.order)
            and np.array_equal(a_flags, b.end_of_node))


def same_tree_up_to_zero_sign(a_nodes, a_order, a_flags, b):
    """Equality that lets node-box floats differ in the sign of a zero only (inputs containing -0.0: the reference's
    own result there depends on primitive order); topology, order and flags must be identical."""
    return (a_nodes.shape == b.nodes.shape and np.array_equal(a_nodes[:, 12:], b.nodes[:, 12:])
            and np.array_equal(a_nodes[:, :12].view(np.float32), b.nodes[:, :12].view(np.float32))
            and np.array_equal(a_order, b.order) and np.array_equal(a_flags, b.end_of_node))
