"""ctypes front-end for the two CPU checkers. TEST INFRASTRUCTURE ONLY.

* ``Ref``    -> oracle/_ref/libatlas_ref.so, the UNMODIFIED reference (Atlas::Volume::BVH, /root/reference/src/engine/
               volume/BVH.cpp) compiled by oracle/Makefile; kind "reference".
* ``Oracle`` -> oracle/libatlas_oracle.so, our restatement (builder + GLSL-order traversal with visit counters);
               kind "port".

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module. The
product package (atlas_engine_b200) never does.
"""
import atexit
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_u64, _u32, _i32, _f32, _vp = C.c_uint64, C.c_uint32, C.c_int32, C.c_float, C.c_void_p


def build(force=False):
    """Compile libatlas_oracle.so always, and _ref/libatlas_ref.so when /root/reference is present."""
    args = ["make", "-C", _HERE, "-s"] + (["-B"] if force else []) + ["all"]
    subprocess.run(args, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def _ptr(a):
    return a.ctypes.data_as(_vp)


def _f32c(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class FlatBVH:
    """Flattened tree in the reference's host layout: nodes (n,14) uint32 words (56 B), order, end_of_node."""

    def __init__(self, nodes_words, order, flags, stats=None):
        self.nodes = nodes_words            # (n, 14) uint32 view of BVHNode
        self.order = order                  # (m,) uint32 source index per flattened slot
        self.end_of_node = flags            # (m,) uint8
        self.stats = stats

    def gpu_nodes(self):
        """GPUBVHNode 64 B layout (RTStructures.h:95-103): 12 floats of boxes, leftPtr, rightPtr, 2 pad ints."""
        out = np.zeros((self.nodes.shape[0], 16), dtype=np.uint32)
        out[:, :14] = self.nodes
        return out.view(np.float32)


class Ref:
    path = os.path.join(_HERE, "_ref", "libatlas_ref.so")

    def __init__(self):
        if not os.path.exists(self.path):
            raise FileNotFoundError(self.path)
        L = self.lib = C.CDLL(self.path)
        L.ref_build_blas.restype = _vp
        L.ref_build_blas.argtypes = [_vp, _vp, _u64, C.c_int]
        L.ref_build_tlas.restype = _vp
        L.ref_build_tlas.argtypes = [_vp, _u64, C.c_int]
        for f in (L.ref_bvh_node_count, L.ref_bvh_ref_count):
            f.restype = _u64
            f.argtypes = [_vp]
        L.ref_bvh_copy_nodes.argtypes = [_vp, _vp]
        L.ref_bvh_copy_order.argtypes = [_vp, _vp, _vp]
        L.ref_bvh_free.argtypes = [_vp]
        L.ref_bvh_intersect_closest.argtypes = [_vp, _vp, _u64, _vp, _vp, C.c_int]
        L.ref_bvh_intersect_any.argtypes = [_vp, _vp, _u64, _vp, C.c_int]
        L.ref_hardware_concurrency.restype = C.c_int
        L.ref_init()
        atexit.register(L.ref_shutdown)

    @classmethod
    def available(cls):
        return os.path.exists(cls.path)

    def cores(self):
        return int(self.lib.ref_hardware_concurrency())

    def _collect(self, h, keep):
        n, m = self.lib.ref_bvh_node_count(h), self.lib.ref_bvh_ref_count(h)
        nodes = np.zeros((n, 14), dtype=np.uint32)
        order = np.zeros(m, dtype=np.uint32)
        flags = np.zeros(m, dtype=np.uint8)
        self.lib.ref_bvh_copy_nodes(h, _ptr(nodes))
        self.lib.ref_bvh_copy_order(h, _ptr(order), _ptr(flags))
        out = FlatBVH(nodes, order, flags)
        if keep:
            out.handle = h
        else:
            self.lib.ref_bvh_free(h)
        return out

    def build_blas(self, aabbs, tris, parallel=True, keep=False):
        aabbs, tris = _f32c(aabbs), _f32c(tris)
        h = self.lib.ref_build_blas(_ptr(aabbs), _ptr(tris), aabbs.shape[0], int(parallel))
        return self._collect(h, keep)

    def build_blas_timed(self, aabbs, tris, parallel=True):
        """Constructor-in to constructor-out wall time in seconds (BVH.cpp:14-56), result discarded."""
        import time
        aabbs, tris = _f32c(aabbs), _f32c(tris)
        t0 = time.perf_counter()
        h = self.lib.ref_build_blas(_ptr(aabbs), _ptr(tris), aabbs.shape[0], int(parallel))
        dt = time.perf_counter() - t0
        self.lib.ref_bvh_free(h)
        return dt

    def build_tlas(self, aabbs, parallel=True):
        aabbs = _f32c(aabbs)
        h = self.lib.ref_build_tlas(_ptr(aabbs), aabbs.shape[0], int(parallel))
        return self._collect(h, False)

    def intersect_closest(self, bvh, rays8, nthreads=1):
        """BVH::GetIntersection per ray; rays8 = (n, 8) origin, direction, tMin, tMax. Returns (tuv, source idx)."""
        rays8 = _f32c(rays8)
        n = rays8.shape[0]
        tuv = np.zeros((n, 3), dtype=np.float32)
        idx = np.zeros(n, dtype=np.int32)
        self.lib.ref_bvh_intersect_closest(bvh.handle, _ptr(rays8), n, _ptr(tuv), _ptr(idx), nthreads)
        return tuv, idx

    def pack_signed(self, vec4s):
        """Atlas::Common::Packing::PackSignedVector3x10_1x2 (the reference's compiled code) over (n, 4) float32 vectors."""
        v = np.ascontiguousarray(vec4s, dtype=np.float32).reshape(-1, 4)
        out = np.zeros(v.shape[0], dtype=np.int32)
        self.lib.ref_pack_signed_3x10_1x2(_ptr(v), C.c_uint64(v.shape[0]), _ptr(out))
        return out

    def intersect_any(self, bvh, rays8, nthreads=1):
        rays8 = _f32c(rays8)
        n = rays8.shape[0]
        hit = np.zeros(n, dtype=np.uint8)
        self.lib.ref_bvh_intersect_any(bvh.handle, _ptr(rays8), n, _ptr(hit), nthreads)
        return hit

    def free(self, bvh):
        self.lib.ref_bvh_free(bvh.handle)
        bvh.handle = None


STAT_NAMES = ("median_splits", "sort_fallbacks", "sort_fallback_max_n", "spatial_tried", "spatial_chosen",
              "axis_skipped", "max_depth", "sum_leaf_depth", "duplicates", "last_resort_leaf")
COUNTER_NAMES = ("tlas_nodes", "instances", "blas_nodes", "triangles", "max_stack", "rays_stack_gt32")


class Scene:
    """Host-side flattened scene in the GPU layouts, as the traversal restatement consumes it."""

    def __init__(self, tlas_nodes, instances, blas_nodes, bvh_tris, triangles96=None):
        self.tlas_nodes = _f32c(tlas_nodes).reshape(-1, 16)
        self.instances = np.ascontiguousarray(instances).view(np.float32).reshape(-1, 16)
        self.blas_nodes = [_f32c(b).reshape(-1, 16) for b in blas_nodes]
        self.bvh_tris = [_f32c(t).reshape(-1, 12) for t in bvh_tris]
        m = len(self.blas_nodes)
        self._node_ptrs = (C.c_void_p * m)(*[b.ctypes.data for b in self.blas_nodes])
        self._tri_ptrs = (C.c_void_p * m)(*[t.ctypes.data for t in self.bvh_tris])
        self.tri_counts = np.array([t.shape[0] for t in self.bvh_tris], dtype=np.uint32)
        self.triangles96 = None if triangles96 is None else [_f32c(t).reshape(-1, 24) for t in triangles96]
        self._tri96_ptrs = None if triangles96 is None else (C.c_void_p * m)(*[t.ctypes.data for t in self.triangles96])
        self.materials = None
        self.textures = []

    def set_materials(self, materials, textures=()):
        """materials: (k, 23) uint32 RaytraceMaterial words; textures: list of (h, w) uint8 arrays (R8 opacity maps)."""
        self.materials = np.ascontiguousarray(materials).view(np.uint32).reshape(-1, 23)
        self.textures = [np.ascontiguousarray(t, dtype=np.uint8) for t in textures]
        self._tex_dims = np.array([[t.shape[1], t.shape[0]] for t in self.textures], dtype=np.uint32).reshape(-1, 2)
        self._tex_ptrs = (C.c_void_p * max(1, len(self.textures)))(*[t.ctypes.data for t in self.textures])

    def material_args(self):
        if self.materials is None:
            return (None, 0, None, None, 0)
        return (_ptr(self.materials), self.materials.shape[0], _ptr(self._tex_dims) if len(self.textures) else None,
                self._tex_ptrs if len(self.textures) else None, len(self.textures))


class Oracle:
    path = os.path.join(_HERE, "libatlas_oracle.so")

    def __init__(self):
        if not os.path.exists(self.path):
            build()
        L = self.lib = C.CDLL(self.path)
        L.oracle_build_blas.restype = _vp
        L.oracle_build_blas.argtypes = [_vp, _vp, _u64]
        L.oracle_build_tlas.restype = _vp
        L.oracle_build_tlas.argtypes = [_vp, _u64]
        for f in (L.oracle_tree_node_count, L.oracle_tree_ref_count):
            f.restype = _u64
            f.argtypes = [_vp]
        L.oracle_tree_copy_nodes.argtypes = [_vp, _vp]
        L.oracle_tree_copy_order.argtypes = [_vp, _vp, _vp]
        L.oracle_tree_stats.argtypes = [_vp, _vp]
        L.oracle_tree_free.argtypes = [_vp]
        L.oracle_trace.argtypes = [_vp, _vp, _vp, _vp, _vp, _u64, _u32, _f32, _f32, C.c_int, C.c_int, _vp, _vp, C.c_int, _vp, C.c_int,
                                   _vp, _u32, _vp, _vp, _u32]
        L.oracle_raygen.argtypes = [_vp, _vp, _vp, _vp, _u32, _u32, _u32, _i32, _vp]
        L.oracle_pt_shadow_rays.argtypes = [_vp, _vp, _vp, _u32, _vp, _vp, _u32, _vp, _u64, _vp, _vp]
        L.oracle_pt_shade.argtypes = [_vp, _vp, _vp, _u32, _vp, _vp, _u32, _vp, _vp, _vp, _u64, _vp, _f32, _u32, _vp, _vp, _vp, _vp, _vp]
        L.oracle_ray_bins.argtypes = [_vp, _u64, _vp]
        L.oracle_get_opacity.restype = _f32
        L.oracle_get_opacity.argtypes = [_vp, _f32, _f32, _vp, _vp, _vp, _u32]
        L.oracle_brute_force.argtypes = [_vp, _u32, _vp, _vp, _vp, _u64, _u32, _f32, _f32, _vp, _vp, _vp, C.c_int]
        L.oracle_pack_shading_words.argtypes = [_vp, _vp, _vp, _vp, _u64, _vp]

    def _collect(self, h):
        n, m = self.lib.oracle_tree_node_count(h), self.lib.oracle_tree_ref_count(h)
        nodes = np.zeros((n, 14), dtype=np.uint32)
        order = np.zeros(m, dtype=np.uint32)
        flags = np.zeros(m, dtype=np.uint8)
        st = np.zeros(10, dtype=np.uint64)
        self.lib.oracle_tree_copy_nodes(h, _ptr(nodes))
        self.lib.oracle_tree_copy_order(h, _ptr(order), _ptr(flags))
        self.lib.oracle_tree_stats(h, _ptr(st))
        self.lib.oracle_tree_free(h)
        return FlatBVH(nodes, order, flags, dict(zip(STAT_NAMES, (int(x) for x in st))))

    def build_blas(self, aabbs, tris):
        aabbs, tris = _f32c(aabbs), _f32c(tris)
        return self._collect(self.lib.oracle_build_blas(_ptr(aabbs), _ptr(tris), aabbs.shape[0]))

    def build_tlas(self, aabbs):
        aabbs = _f32c(aabbs)
        return self._collect(self.lib.oracle_build_tlas(_ptr(aabbs), aabbs.shape[0]))

    def trace(self, scene, rays, cull_mask=1 << 7, t_min=0.0, t_max=1e12, any_hit=False, per_ray_tmax=False,
              nthreads=1, opacity=False):
        """rays (n, 12) float32 PackedRay. Returns (out rays (n, 12), counters dict)."""
        rays = _f32c(rays).reshape(-1, 12)
        out = np.zeros_like(rays)
        ct = np.zeros(6, dtype=np.uint64)
        self.lib.oracle_trace(_ptr(scene.tlas_nodes), _ptr(scene.instances), scene._node_ptrs, scene._tri_ptrs,
                              _ptr(rays), rays.shape[0], cull_mask, t_min, t_max, int(any_hit), int(per_ray_tmax),
                              _ptr(out), _ptr(ct), nthreads, scene._tri96_ptrs if opacity else None, int(opacity),
                              *scene.material_args())
        return out, dict(zip(COUNTER_NAMES, (int(x) for x in ct)))

    # ---------------------------------------------------------------------------------------------- path tracer
    def raygen(self, eye, origin, right, bottom, width, height, samples=1, sample_count=0):
        """pathtracer/rayGen.csh (non-REALTIME): (w*h*samples, 12) PackedRay in the shader's storage order."""
        e, o, r, b = (_f32c(x) for x in (eye, origin, right, bottom))
        out = np.zeros((width * height * samples, 12), dtype=np.float32)
        self.lib.oracle_raygen(_ptr(e), _ptr(o), _ptr(r), _ptr(b), width, height, samples, sample_count, _ptr(out))
        return out

    def pt_shadow_rays(self, scene, rays, params):
        rays = _f32c(rays).reshape(-1, 12)
        out = np.zeros_like(rays)
        self.lib.oracle_pt_shadow_rays(_ptr(scene.instances), scene._tri96_ptrs, *scene.material_args(), _ptr(rays), rays.shape[0],
                                       C.byref(params), _ptr(out))
        return out

    def pt_shade(self, scene, rays, payload_in, visibility, params, seed, bounce):
        """rayHit.csh on traced rays. Returns dict(alive (n,) bool, rays (n,12), payload (n,4) uint32, finished (n,3), rr (n,2))."""
        rays = _f32c(rays).reshape(-1, 12)
        n = rays.shape[0]
        pay = np.zeros((n, 4), dtype=np.uint32) if payload_in is None else np.ascontiguousarray(payload_in).view(np.uint32).reshape(-1, 4)
        vis = _f32c(visibility)
        alive = np.zeros(n, dtype=np.uint8)
        ro = np.zeros((n, 12), dtype=np.float32)
        po = np.zeros((n, 4), dtype=np.uint32)
        fin = np.zeros((n, 3), dtype=np.float32)
        rr = np.zeros((n, 2), dtype=np.float32)
        self.lib.oracle_pt_shade(_ptr(scene.instances), scene._tri96_ptrs, *scene.material_args(), _ptr(rays), _ptr(pay), _ptr(vis), n,
                                 C.byref(params), seed, bounce, _ptr(alive), _ptr(ro), _ptr(po), _ptr(fin), _ptr(rr))
        return dict(alive=alive.astype(bool), rays=ro, payload=po, finished=fin, rr=rr)

    def pack_signed(self, vec4s):
        v = np.ascontiguousarray(vec4s, dtype=np.float32).reshape(-1, 4)
        out = np.zeros(v.shape[0], dtype=np.int32)
        self.lib.oracle_pack_signed_3x10_1x2(_ptr(v), C.c_uint64(v.shape[0]), _ptr(out))
        return out

    def ray_bins(self, rays):
        rays = _f32c(rays).reshape(-1, 12)
        out = np.zeros(rays.shape[0], dtype=np.uint32)
        self.lib.oracle_ray_bins(_ptr(rays), rays.shape[0], _ptr(out))
        return out

    def pack_shading_words(self, tris, normals9=None, uvs6=None, colors12=None):
        """(n, 11) uint32: the packed shading words of MeshData.cpp:176-228 per source triangle."""
        tris = _f32c(tris).reshape(-1, 9)
        arrs = [None if a is None else _f32c(a) for a in (normals9, uvs6, colors12)]
        out = np.zeros((tris.shape[0], 11), dtype=np.uint32)
        self.lib.oracle_pack_shading_words(_ptr(tris), *[None if a is None else _ptr(a) for a in arrs], tris.shape[0], _ptr(out))
        return out

    def brute_force(self, scene, rays, cull_mask=1 << 7, t_min=0.0, t_max=1e12, nthreads=1):
        rays = _f32c(rays).reshape(-1, 12)
        n = rays.shape[0]
        t = np.zeros(n, dtype=np.float32)
        tri = np.zeros(n, dtype=np.int32)
        inst = np.zeros(n, dtype=np.int32)
        self.lib.oracle_brute_force(_ptr(scene.instances), scene.instances.shape[0], scene._tri_ptrs,
                                    _ptr(scene.tri_counts), _ptr(rays), n, cull_mask, t_min, t_max, _ptr(t), _ptr(tri),
                                    _ptr(inst), nthreads)
        return t, tri, inst
