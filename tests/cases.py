"""Shared input batteries for the parity tests (same seeds on CPU and GPU)."""
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
