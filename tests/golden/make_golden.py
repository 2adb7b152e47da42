"""Generates the golden fixtures from the UNMODIFIED reference (oracle/_ref/libatlas_ref.so, i.e. /root/reference/src/
engine/volume/BVH.cpp compiled in place). Run in the build container only: python tests/golden/make_golden.py"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases as CS
from atlas_engine_b200 import workloads as W
from oracle.pyoracle import Ref


def digest(tree):
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(tree.nodes).tobytes())
    h.update(np.ascontiguousarray(tree.order).tobytes())
    h.update(np.ascontiguousarray(tree.end_of_node).tobytes())
    return h.hexdigest()


ref = Ref()
gold = {"blas": {}, "tlas": {}, "source": "Atlas::Volume::BVH via oracle/_ref (reference compiled unmodified)"}
small = {}
for name, tris in CS.build_cases().items():
    t = ref.build_blas(W.tri_boxes(tris), tris, parallel=True)
    gold["blas"][name] = digest(t)
    if name in ("sphere", "soup33", "identical5", "coincident"):
        small[name + "_nodes"], small[name + "_order"], small[name + "_flags"] = t.nodes, t.order, t.end_of_node
for name, boxes in CS.tlas_cases().items():
    gold["tlas"][name] = digest(ref.build_tlas(boxes))

# ---- real geometry shipped with the reference: copied here so they travel to the GPU box, digests made
# by the reference builder. data/chromesphere.bin: 3840 float3 positions at byte 0, 3840 u16 indices at byte 122880
# (data/chromesphere.gltf bufferViews 0 and 3); data/meshes/*.aemesh: the engine's own mesh files.
import shutil
from atlas_engine_b200 import capi
REFDATA = "/root/reference/data"
os.makedirs(os.path.join(HERE, "meshes"), exist_ok=True)
shutil.copyfile(os.path.join(REFDATA, "chromesphere.bin"), os.path.join(HERE, "meshes", "chromesphere.bin"))
for n in ("chromesphere", "capsule", "metallicwall"):
    shutil.copyfile(os.path.join(REFDATA, "meshes", n + ".aemesh"), os.path.join(HERE, "meshes", n + ".aemesh"))
gold["real"] = {}
for name, tris in CS.real_mesh_cases().items():
    t = ref.build_blas(W.tri_boxes(tris), tris, parallel=True)
    gold["real"][name] = digest(t)
    small["real_" + name + "_order"] = t.order
with open(os.path.join(HERE, "build_hashes.json"), "w") as f:
    json.dump(gold, f, indent=1, sort_keys=True)
np.savez_compressed(os.path.join(HERE, "build_small.npz"), **small)

# reference CPU traversal (BVH::GetIntersection) over the sphere
tris = W.uv_sphere()
boxes = W.tri_boxes(tris)
lo, hi = boxes[:, :3].min(0), boxes[:, 3:].max(0)
rays = W.random_rays(4000, lo - 0.3, hi + 0.3, seed=77)
rb = ref.build_blas(boxes, tris, keep=True)
r8 = np.concatenate([rays[:, 0:3], rays[:, 4:7], np.zeros((len(rays), 1), np.float32), np.full((len(rays), 1), 1e12, np.float32)], axis=1)
tuv, idx = ref.intersect_closest(rb, r8, 1)
np.savez_compressed(os.path.join(HERE, "trace_small.npz"), rays=rays, ref_tuv=tuv, ref_idx=idx.astype(np.int64))
print("golden written:", len(gold["blas"]), "blas,", len(gold["tlas"]), "tlas,", int((idx >= 0).sum()), "hits")
