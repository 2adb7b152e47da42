"""ctypes binding of include/atlas_rt.h (libatlas_rt.so, built in-tree by __graft_entry__.build()).

This is plumbing for tests and bench.py; the product is the CUDA library. There is NO fallback: if the shared object
is missing or a call fails, an exception is raised.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libatlas_rt.so")

DEVICE_INPUT, DEVICE_OUTPUT, ASYNC, PER_RAY_TMAX, COUNTERS, OPACITY = 1, 2, 4, 8, 16, 32
RAY_BINNING, ACCUM_TILE_ORDER, HITS_ONLY, PEER_OUTPUT, PIPELINED = 64, 128, 256, 512, 1024
MASK_ALL, MASK_SHADOW = 1 << 7, 1 << 6
INF = 1e12
STATUS = {0: "OK", -1: "ERR_INVALID", -2: "ERR_CUDA", -3: "ERR_OOM", -4: "ERR_UNSUPPORTED", -5: "ERR_STACK"}

_vp, _u64, _u32, _i32, _f32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_float

# name -> (restype, argtypes); every symbol include/atlas_rt.h declares (tests/test_abi.py checks the two agree)
SIGNATURES = {
    "atlas_rt_version": (_i32, []),
    "atlas_rt_context_create": (_i32, [_i32, _vp, C.POINTER(_vp)]),
    "atlas_rt_context_destroy": (None, [_vp]),
    "atlas_rt_context_synchronize": (_i32, [_vp]),
    "atlas_rt_last_error": (C.c_char_p, [_vp]),
    "atlas_rt_kernel_launches": (_u64, [_vp]),
    "atlas_rt_build_blas": (_i32, [_vp, _vp, _vp, _u64, _u32, C.POINTER(_vp)]),
    "atlas_rt_build_blas_batch": (_i32, [_vp, _u32, _vp, _vp, _vp, _u32, _vp]),
    "atlas_rt_build_tlas": (_i32, [_vp, _vp, _u64, _u32, C.POINTER(_vp)]),
    "atlas_rt_bvh_upload": (_i32, [_vp, _vp, _u64, _vp, _vp, _u64, C.POINTER(_vp)]),
    "atlas_rt_bvh_import": (_i32, [_vp, _vp, _u64, _vp, _vp, _u64, _u32, C.POINTER(_vp)]),
    "atlas_rt_bvh_counts": (_i32, [_vp, C.POINTER(_u64), C.POINTER(_u64)]),
    "atlas_rt_bvh_download": (_i32, [_vp, _vp, _vp, _vp, _u32]),
    "atlas_rt_bvh_device_ptrs": (_i32, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    "atlas_rt_bvh_stats": (_i32, [_vp, _vp]),
    "atlas_rt_bvh_free": (None, [_vp]),
    "atlas_rt_pack_mesh": (_i32, [_vp, _vp, _vp, _u64, _vp, _vp, _u32, C.POINTER(_vp)]),
    "atlas_rt_mesh_pack_shading": (_i32, [_vp, _vp, _vp, _u64, _vp, _vp, _vp, _u32]),
    "atlas_rt_mesh_download_shading": (_i32, [_vp, _vp, _u32]),
    "atlas_rt_mesh_counts": (_i32, [_vp, C.POINTER(_u64), C.POINTER(_u64)]),
    "atlas_rt_mesh_download": (_i32, [_vp, _vp, _vp, _u32]),
    "atlas_rt_mesh_free": (None, [_vp]),
    "atlas_rt_pack_shading_words": (_i32, [_vp, _vp, _vp, _vp, _vp, _u64, _vp, _u32]),
    "atlas_rt_aemesh_open": (_i32, [C.c_char_p, C.POINTER(_vp)]),
    "atlas_rt_aemesh_counts": (_i32, [_vp, C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u32)]),
    "atlas_rt_aemesh_material_path": (C.c_char_p, [_vp, _u32]),
    "atlas_rt_aemesh_triangles": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "atlas_rt_aemesh_raw": (_i32, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    "atlas_rt_aemesh_close": (None, [_vp]),
    "atlas_rt_scene_create": (_i32, [_vp, _vp, _u32, _vp, _u64, _vp, _u32, C.POINTER(_vp)]),
    "atlas_rt_scene_download": (_i32, [_vp, _vp, _vp, _u32]),
    "atlas_rt_scene_free": (None, [_vp]),
    "atlas_rt_trace_closest": (_i32, [_vp, _vp, _vp, _u64, _u32, _f32, _f32, _vp, _u32]),
    "atlas_rt_trace_any": (_i32, [_vp, _vp, _vp, _u64, _u32, _f32, _f32, _vp, _u32]),
    "atlas_rt_trace_counters": (_i32, [_vp, _vp]),
    "atlas_rt_generate_primary_rays": (_i32, [_vp, _vp, _u32, _u32, _u32, _vp, _vp, _u32]),
    "atlas_rt_sample_jitter": (None, [_i32, _vp]),
    "atlas_rt_scene_set_materials": (_i32, [_vp, _vp, _vp, _u32, _vp, _u32]),
    "atlas_rt_pathtrace_bounce": (_i32, [_vp, _vp, _vp, _f32, _u32, _vp, _vp, _u64, _vp, _vp, _vp, _u32, _u32, C.POINTER(_u64), _u32]),
    "atlas_rt_pathtrace_bounces": (_i32, [_vp, _vp, _vp, _u32, _u32, _vp, _u32, _i32, _vp, _u64, _u64, _vp, C.POINTER(_u64), _u32]),
    "atlas_rt_pathtrace_bounces_interleaved": (_i32, [_vp, _vp, _vp, _u32, _u32, _vp, _u32, _i32, _vp, _u32, _u32, _u32, _vp, C.POINTER(_u64),
                                                      C.POINTER(_u64), _u32]),
    "atlas_rt_image_from_shards": (_i32, [_vp, _vp, _u32, _u32, _u32, _u32, _vp, _u32]),
    "atlas_rt_bin_rays": (_i32, [_vp, _vp, _vp, _u64, _vp, _vp, _u32]),
    "atlas_rt_comm_unique_id": (_i32, [_vp]),
    "atlas_rt_comm_init": (_i32, [_vp, _vp, _u32, _u32, C.POINTER(_vp)]),
    "atlas_rt_comm_destroy": (None, [_vp]),
    "atlas_rt_comm_info": (_i32, [_vp, C.POINTER(_u32), C.POINTER(_u32)]),
    "atlas_rt_comm_synchronize": (_i32, [_vp]),
    "atlas_rt_bvh_broadcast": (_i32, [_vp, _vp, _u32, C.POINTER(_vp)]),
    "atlas_rt_build_scene_sharded": (_i32, [_vp, _u32, _vp, _vp, _vp, _vp, _vp, _u64, _u32, C.POINTER(_vp)]),
    "atlas_rt_scene_replicate": (_i32, [_vp, _vp, _u32, C.POINTER(_vp)]),
    "atlas_rt_trace_join": (_i32, [_vp]),
    "atlas_rt_trace_sharded": (_i32, [_vp, _vp, _vp, _u64, _u32, _f32, _f32, _vp, _u32, _u32, _i32]),
    "atlas_rt_comm_gather": (_i32, [_vp, _vp, _u64, _vp, _vp, _vp, _u32, _u32]),
    "atlas_rt_comm_peer_hits": (_i32, [_vp, C.POINTER(_vp)]),
    "atlas_rt_shard_range": (_i32, [_u64, _u32, _u32, _u32, C.POINTER(_u64), C.POINTER(_u64)]),
}

_lib = None


def lib():
    """Load libatlas_rt.so (once). Raises if it has not been built — there is no CPU path to fall back to."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(nvcc, sm_100a). atlas_engine_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class AtlasError(RuntimeError):
    pass


class Camera(C.Structure):
    """atlas_rt_camera"""
    _fields_ = [("eye", _f32 * 3), ("origin", _f32 * 3), ("right", _f32 * 3), ("bottom", _f32 * 3)]


class PtParams(C.Structure):
    """atlas_rt_pt_params"""
    _fields_ = [("light_dir", _f32 * 3), ("light_radiance", _f32 * 3), ("light_count", _i32), ("sky_radiance", _f32 * 3),
                ("max_bounces", _u32), ("samples_per_frame", _u32)]


def pt_params(light_dir, light_radiance, sky_radiance, max_bounces, samples_per_frame=1, light_count=1):
    return PtParams((_f32 * 3)(*light_dir), (_f32 * 3)(*light_radiance), light_count, (_f32 * 3)(*sky_radiance), max_bounces, samples_per_frame)


class Texture(C.Structure):
    """atlas_rt_texture"""
    _fields_ = [("width", _u32), ("height", _u32), ("texels", _vp)]


# atlas_rt_material / RaytraceMaterial (data/shader/raytracer/structures.hsh:108-141): 23 words
MATERIAL_DTYPE = np.dtype([("ID", np.int32), ("baseR", np.float32), ("baseG", np.float32), ("baseB", np.float32),
                           ("emissR", np.float32), ("emissG", np.float32), ("emissB", np.float32), ("opacity", np.float32),
                           ("roughness", np.float32), ("metalness", np.float32), ("ao", np.float32), ("reflectance", np.float32),
                           ("normalScale", np.float32), ("invertUVs", np.int32), ("twoSided", np.int32), ("cullBackFaces", np.int32),
                           ("useVertexColors", np.int32), ("baseColorTexture", np.int32), ("opacityTexture", np.int32),
                           ("normalTexture", np.int32), ("roughnessTexture", np.int32), ("metalnessTexture", np.int32), ("aoTexture", np.int32)])


def make_materials(n):
    """n default materials: grey, opaque, rough dielectric, two-sided, no textures."""
    m = np.zeros(n, dtype=MATERIAL_DTYPE)
    m["ID"] = np.arange(n)
    m["baseR"] = m["baseG"] = m["baseB"] = 0.8
    m["opacity"] = m["roughness"] = m["ao"] = 1.0
    m["reflectance"] = 0.5
    m["twoSided"] = 1
    for k in ("baseColorTexture", "opacityTexture", "normalTexture", "roughnessTexture", "metalnessTexture", "aoTexture"):
        m[k] = -1
    return m


def sample_jitter(sample_count):
    out = np.zeros(2, dtype=np.float32)
    lib().atlas_rt_sample_jitter(sample_count, _addr(out))
    return out


def _addr(x):
    """Pointer value of a numpy array (host), an int (device pointer) or None."""
    if x is None:
        return None
    if isinstance(x, (int, np.integer)):
        return int(x)
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"]
        return x.ctypes.data
    if hasattr(x, "data_ptr"):   # torch tensor
        return int(x.data_ptr())
    raise TypeError(type(x))


def _is_device(x):
    return isinstance(x, (int, np.integer)) or (hasattr(x, "is_cuda") and x.is_cuda)


class Context:
    """atlas_rt_context: one CUDA device + stream. stream = raw cudaStream_t value (e.g. torch stream .cuda_stream)."""

    def __init__(self, device=0, stream=None):
        self.L = lib()
        h = _vp()
        rc = self.L.atlas_rt_context_create(device, stream, C.byref(h))
        if rc != 0:
            raise AtlasError(f"atlas_rt_context_create(device={device}) -> {STATUS.get(rc, rc)} (no usable CUDA device?)")
        self.h = h
        self.device = device

    def check(self, rc):
        if rc != 0:
            msg = self.L.atlas_rt_last_error(self.h)
            raise AtlasError(f"{STATUS.get(rc, rc)}: {msg.decode() if msg else ''}")

    def synchronize(self):
        self.check(self.L.atlas_rt_context_synchronize(self.h))

    def launches(self):
        return int(self.L.atlas_rt_kernel_launches(self.h))

    def close(self):
        if self.h:
            self.L.atlas_rt_context_destroy(self.h)
            self.h = None

    # ------------------------------------------------------------------------------------------------ build
    def build_blas(self, aabbs, tris, count=None, flags=0):
        dev = _is_device(aabbs)
        if not dev:
            aabbs = np.ascontiguousarray(aabbs, dtype=np.float32)
            tris = np.ascontiguousarray(tris, dtype=np.float32)
            count = aabbs.shape[0]
        h = _vp()
        self.check(self.L.atlas_rt_build_blas(self.h, _addr(aabbs), _addr(tris), count, flags | (DEVICE_INPUT if dev else 0), C.byref(h)))
        return BVH(self, h)

    def build_blas_batch(self, aabbs_list, tris_list, counts=None, flags=0):
        """atlas_rt_build_blas_batch: lists of (n_i, 6) / (n_i, 9) host arrays, or of CUDA tensors / device pointers with
        `counts` given. Returns one BVH per mesh."""
        m = len(aabbs_list)
        dev = m > 0 and _is_device(aabbs_list[0])
        if not dev:
            aabbs_list = [np.ascontiguousarray(a, dtype=np.float32) for a in aabbs_list]
            tris_list = [np.ascontiguousarray(t, dtype=np.float32) for t in tris_list]
            counts = [a.shape[0] for a in aabbs_list]
        pa = (_vp * m)(*[_addr(a) for a in aabbs_list])
        pt = (_vp * m)(*[_addr(t) for t in tris_list])
        pc = (_u64 * m)(*[int(c) for c in counts])
        out = (_vp * m)()
        self.check(self.L.atlas_rt_build_blas_batch(self.h, m, pa, pt, pc, flags | (DEVICE_INPUT if dev else 0), out))
        return [BVH(self, _vp(out[k])) for k in range(m)]

    def build_tlas(self, aabbs, count=None, flags=0):
        dev = _is_device(aabbs)
        if not dev:
            aabbs = np.ascontiguousarray(aabbs, dtype=np.float32)
            count = aabbs.shape[0]
        h = _vp()
        self.check(self.L.atlas_rt_build_tlas(self.h, _addr(aabbs), count, flags | (DEVICE_INPUT if dev else 0), C.byref(h)))
        return BVH(self, h)

    def upload_bvh(self, nodes56, order, end_of_node):
        nodes56 = np.ascontiguousarray(nodes56, dtype=np.uint32).reshape(-1, 14)
        order = np.ascontiguousarray(order, dtype=np.uint32)
        end_of_node = np.ascontiguousarray(end_of_node, dtype=np.uint8)
        h = _vp()
        self.check(self.L.atlas_rt_bvh_upload(self.h, _addr(nodes56), nodes56.shape[0], _addr(order), _addr(end_of_node), order.shape[0], C.byref(h)))
        return BVH(self, h)

    def import_bvh_device(self, nodes56, order, end_of_node):
        """Wrap a flattened tree whose arrays are CUDA tensors on this device (int32 (n,14), int32 (m,), uint8 (m,))."""
        h = _vp()
        self.check(self.L.atlas_rt_bvh_import(self.h, _addr(nodes56), nodes56.shape[0], _addr(order), _addr(end_of_node), order.shape[0],
                                              DEVICE_INPUT, C.byref(h)))
        return BVH(self, h)

    def pack_mesh(self, blas, tris, count=None, material_idx=None, opacity=None, flags=0):
        dev = _is_device(tris)
        if not dev:
            tris = np.ascontiguousarray(tris, dtype=np.float32)
            count = tris.shape[0]
            if material_idx is not None:
                material_idx = np.ascontiguousarray(material_idx, dtype=np.int32)
            if opacity is not None:
                opacity = np.ascontiguousarray(opacity, dtype=np.float32)
        h = _vp()
        self.check(self.L.atlas_rt_pack_mesh(self.h, blas.h, _addr(tris), count, _addr(material_idx), _addr(opacity),
                                             flags | (DEVICE_INPUT if dev else 0), C.byref(h)))
        return Mesh(self, h, blas)

    def pack_shading_words(self, tris, normals9=None, uvs6=None, colors12=None):
        """(n, 11) uint32 packed shading words (MeshData.cpp:176-228) computed on the device."""
        tris = np.ascontiguousarray(tris, dtype=np.float32).reshape(-1, 9)
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float32) for a in (normals9, uvs6, colors12)]
        out = np.zeros((tris.shape[0], 11), dtype=np.uint32)
        self.check(self.L.atlas_rt_pack_shading_words(self.h, _addr(tris), _addr(arrs[0]), _addr(arrs[1]), _addr(arrs[2]), tris.shape[0],
                                                      _addr(out), 0))
        return out

    def create_scene(self, meshes, instances, tlas, flags=0):
        instances = np.ascontiguousarray(instances).view(np.uint32).reshape(-1, 16)
        arr = (_vp * len(meshes))(*[m.h for m in meshes])
        h = _vp()
        self.check(self.L.atlas_rt_scene_create(self.h, arr, len(meshes), _addr(instances), instances.shape[0], tlas.h, flags, C.byref(h)))
        return Scene(self, h, meshes, tlas)

    # ------------------------------------------------------------------------------------------------ trace
    def trace(self, scene, rays, count=None, out=None, cull_mask=MASK_ALL, t_min=0.0, t_max=INF, any_hit=False, flags=0):
        """rays: (n, 12) float32 numpy array (host) or a device pointer / CUDA tensor with `count` given.
        Returns the output array for host input, or None for device input (results land in `out`)."""
        fn = self.L.atlas_rt_trace_any if any_hit else self.L.atlas_rt_trace_closest
        if _is_device(rays):
            out = rays if out is None else out
            self.check(fn(self.h, scene.h, _addr(rays), count, cull_mask, t_min, t_max, _addr(out), flags | DEVICE_INPUT | DEVICE_OUTPUT))
            return None
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 12)
        res = (np.empty((rays.shape[0], 4), np.float32) if flags & HITS_ONLY else np.empty_like(rays)) if out is None else out
        self.check(fn(self.h, scene.h, _addr(rays), rays.shape[0], cull_mask, t_min, t_max, _addr(res), flags))
        return res

    # ------------------------------------------------------------------------------------------ path tracer
    def generate_primary_rays(self, eye, origin, right, bottom, width, height, samples=1, jitter=None, out=None):
        """rayGen.csh. out: device pointer / CUDA tensor of width*height*samples PackedRay, or None for a host array."""
        cam = Camera((_f32 * 3)(*eye), (_f32 * 3)(*origin), (_f32 * 3)(*right), (_f32 * 3)(*bottom))
        jit = None if jitter is None else np.ascontiguousarray(jitter, dtype=np.float32)
        if out is None:
            res = np.zeros((width * height * samples, 12), dtype=np.float32)
            self.check(self.L.atlas_rt_generate_primary_rays(self.h, C.byref(cam), width, height, samples, _addr(jit), _addr(res), 0))
            return res
        self.check(self.L.atlas_rt_generate_primary_rays(self.h, C.byref(cam), width, height, samples, _addr(jit), _addr(out), DEVICE_OUTPUT))
        return out

    def pathtrace_bounce(self, scene, params, seed, bounce, rays_in, payload_in, count, rays_out, payload_out, accum, width, height, flags=0):
        """One bounce (traceClosest + rayHit.csh) on device buffers; returns the number of surviving rays (compacted into rays_out)."""
        n = _u64()
        self.check(self.L.atlas_rt_pathtrace_bounce(self.h, scene.h, C.byref(params), seed, bounce, _addr(rays_in), _addr(payload_in), count,
                                                    _addr(rays_out), _addr(payload_out), _addr(accum), width, height, C.byref(n),
                                                    flags | DEVICE_INPUT | DEVICE_OUTPUT))
        return int(n.value)

    def pathtrace_bounces(self, scene, camera, width, height, params, frames, first_sample_count, seeds, accum, slot_begin=0, slot_end=0,
                          flags=0, count_rays=True):
        """`frames` sample passes with the whole bounce loop on the device; returns the closest-hit rays traced (or None)."""
        eye, origin, right, bottom = camera
        cam = Camera((_f32 * 3)(*eye), (_f32 * 3)(*origin), (_f32 * 3)(*right), (_f32 * 3)(*bottom))
        seeds = np.ascontiguousarray(seeds, dtype=np.float32)
        assert seeds.size == frames * (params.max_bounces + 1)
        n = _u64()
        self.check(self.L.atlas_rt_pathtrace_bounces(self.h, scene.h, C.byref(cam), width, height, C.byref(params), frames, first_sample_count,
                                                     _addr(seeds), slot_begin, slot_end, _addr(accum), C.byref(n) if count_rays else None, flags))
        return int(n.value) if count_rays else None

    def pathtrace_bounces_interleaved(self, scene, camera, width, height, params, frames, first_sample_count, seeds, part, parts, block_pixels,
                                      accum_local=None, flags=0, count_rays=True):
        """One of `parts` interleaved shards of the frame into a COMPACT accumulation buffer; accum_local=None only returns the
        shard's pixel count. Returns (local_pixels, closest-hit rays traced or None)."""
        eye, origin, right, bottom = camera
        cam = Camera((_f32 * 3)(*eye), (_f32 * 3)(*origin), (_f32 * 3)(*right), (_f32 * 3)(*bottom))
        seeds = np.ascontiguousarray(seeds, dtype=np.float32)
        lp, n = _u64(), _u64()
        self.check(self.L.atlas_rt_pathtrace_bounces_interleaved(self.h, scene.h, C.byref(cam), width, height, C.byref(params), frames, first_sample_count,
                                                                 _addr(seeds), part, parts, block_pixels, _addr(accum_local), C.byref(lp),
                                                                 C.byref(n) if (count_rays and accum_local is not None) else None, flags))
        return int(lp.value), (int(n.value) if (count_rays and accum_local is not None) else None)

    def image_from_shards(self, gathered, width, height, parts, block_pixels, image, flags=0):
        self.check(self.L.atlas_rt_image_from_shards(self.h, _addr(gathered), width, height, parts, block_pixels, _addr(image), flags | DEVICE_INPUT | DEVICE_OUTPUT))

    def bin_rays(self, rays_in, payload_in, count, rays_out, payload_out, flags=0):
        self.check(self.L.atlas_rt_bin_rays(self.h, _addr(rays_in), _addr(payload_in), count, _addr(rays_out), _addr(payload_out),
                                            flags | DEVICE_INPUT | DEVICE_OUTPUT))

    def trace_join(self):
        """Order the context stream after every ASYNC | PIPELINED host-buffer trace call made so far."""
        self.check(self.L.atlas_rt_trace_join(self.h))

    def trace_counters(self):
        out = np.zeros(6, dtype=np.uint64)
        self.check(self.L.atlas_rt_trace_counters(self.h, _addr(out)))
        names = ("tlas_nodes", "instances", "blas_nodes", "triangles", "max_stack", "rays_stack_gt32")
        return dict(zip(names, (int(x) for x in out)))


class BVH:
    def __init__(self, ctx, h):
        self.ctx, self.h = ctx, h

    def counts(self):
        n, m = _u64(), _u64()
        self.ctx.check(self.ctx.L.atlas_rt_bvh_counts(self.h, C.byref(n), C.byref(m)))
        return int(n.value), int(m.value)

    def download(self, nodes=None, order=None, flags=None):
        """(nodes (n,14) uint32 in the 56 B BVHNode layout, order (m,) uint32, end_of_node (m,) uint8). Preallocated
        (e.g. pinned) arrays at least that large may be passed in; the returned arrays are views of their first rows."""
        n, m = self.counts()
        nodes = np.zeros((n, 14), dtype=np.uint32) if nodes is None else nodes[:n]
        order = np.zeros(m, dtype=np.uint32) if order is None else order[:m]
        flags = np.zeros(m, dtype=np.uint8) if flags is None else flags[:m]
        assert nodes.shape == (n, 14) and nodes.dtype == np.uint32 and nodes.flags.c_contiguous
        assert order.shape == (m,) and order.dtype == np.uint32 and flags.shape == (m,) and flags.dtype == np.uint8
        self.ctx.check(self.ctx.L.atlas_rt_bvh_download(self.h, _addr(nodes), _addr(order), _addr(flags), 0))
        return nodes, order, flags

    def download_device(self):
        """(nodes (n,14) int32, order (m,) int32, end_of_node (m,) uint8) as CUDA tensors on the context's device."""
        import torch
        n, m = self.counts()
        dev = torch.device("cuda", self.ctx.device)
        nodes = torch.empty((n, 14), dtype=torch.int32, device=dev)
        order = torch.empty(m, dtype=torch.int32, device=dev)
        flags = torch.empty(m, dtype=torch.uint8, device=dev)
        self.ctx.check(self.ctx.L.atlas_rt_bvh_download(self.h, _addr(nodes) if n else None, _addr(order) if m else None,
                                                        _addr(flags) if m else None, DEVICE_OUTPUT))
        return nodes, order, flags

    def stats(self):
        out = np.zeros(8, dtype=np.uint64)
        self.ctx.check(self.ctx.L.atlas_rt_bvh_stats(self.h, _addr(out)))
        names = ("spatial_tried", "spatial_chosen", "duplicates", "median_splits", "sort_fallbacks", "sort_fallback_max_n",
                 "levels", "neg_zero")
        return dict(zip(names, (int(x) for x in out)))

    def free(self):
        if self.h:
            self.ctx.L.atlas_rt_bvh_free(self.h)
            self.h = None


class Mesh:
    def __init__(self, ctx, h, blas):
        self.ctx, self.h, self.blas = ctx, h, blas

    def download(self):
        n, m = _u64(), _u64()
        self.ctx.check(self.ctx.L.atlas_rt_mesh_counts(self.h, C.byref(n), C.byref(m)))
        nodes = np.zeros((n.value, 16), dtype=np.float32)
        tris = np.zeros((m.value, 12), dtype=np.float32)
        self.ctx.check(self.ctx.L.atlas_rt_mesh_download(self.h, _addr(nodes), _addr(tris), 0))
        return nodes, tris

    def pack_shading(self, tris, material_idx=None, opacity=None, payload11=None):
        """Adds the 96-byte GPUTriangle array the opacity-aware traversal variants read."""
        tris = np.ascontiguousarray(tris, dtype=np.float32)
        m = None if material_idx is None else np.ascontiguousarray(material_idx, dtype=np.int32)
        o = None if opacity is None else np.ascontiguousarray(opacity, dtype=np.float32)
        p = None if payload11 is None else np.ascontiguousarray(payload11, dtype=np.uint32)
        self.ctx.check(self.ctx.L.atlas_rt_mesh_pack_shading(self.ctx.h, self.h, _addr(tris), tris.shape[0], _addr(m), _addr(o), _addr(p), 0))

    def download_shading(self):
        n, m = _u64(), _u64()
        self.ctx.check(self.ctx.L.atlas_rt_mesh_counts(self.h, C.byref(n), C.byref(m)))
        out = np.zeros((m.value, 24), dtype=np.float32)
        self.ctx.check(self.ctx.L.atlas_rt_mesh_download_shading(self.h, _addr(out), 0))
        return out

    def free(self):
        if self.h:
            self.ctx.L.atlas_rt_mesh_free(self.h)
            self.h = None


class Scene:
    def __init__(self, ctx, h, meshes, tlas):
        self.ctx, self.h, self.meshes, self.tlas = ctx, h, list(meshes), tlas

    def set_materials(self, materials, textures=()):
        """materials: array of MATERIAL_DTYPE (or (k, 23) words); textures: list of (h, w) uint8 arrays (R8 opacity maps)."""
        mats = np.ascontiguousarray(materials).view(np.uint32).reshape(-1, 23)
        tex = [np.ascontiguousarray(t, dtype=np.uint8) for t in textures]
        arr = (Texture * max(1, len(tex)))(*[Texture(t.shape[1], t.shape[0], t.ctypes.data) for t in tex])
        self.ctx.check(self.ctx.L.atlas_rt_scene_set_materials(self.ctx.h, self.h, _addr(mats), mats.shape[0], arr if tex else None, len(tex)))

    def download(self, instance_count=None, tlas_node_count=None):
        """Scenes assembled by the library itself (build_scene_sharded / replicate_scene) carry no Python TLAS object: pass the counts."""
        n, m = (tlas_node_count, instance_count) if self.tlas is None else self.tlas.counts()
        inst = np.zeros((m, 16), dtype=np.uint32)
        nodes = np.zeros((n, 16), dtype=np.float32)
        self.ctx.check(self.ctx.L.atlas_rt_scene_download(self.h, _addr(inst), _addr(nodes), 0))
        return inst, nodes

    def free(self):
        if self.h:
            self.ctx.L.atlas_rt_scene_free(self.h)
            self.h = None


def load_aemesh(path):
    """.aemesh -> dict(tris (n,9), boxes (n,6), material_idx (n,), normals (n,9), uvs (n,6), colors (n,12), materials [paths])
    exactly as the first loop of MeshData::BuildBVH expands the file (atlas_rt_aemesh_*; host-only code in the library)."""
    L = lib()
    h = _vp()
    rc = L.atlas_rt_aemesh_open(os.fsencode(path), C.byref(h))
    if rc != 0:
        raise AtlasError(f"atlas_rt_aemesh_open({path}) -> {STATUS.get(rc, rc)}")
    try:
        nv, ni, nt, ns = _u64(), _u64(), _u64(), _u32()
        L.atlas_rt_aemesh_counts(h, C.byref(nv), C.byref(ni), C.byref(nt), C.byref(ns))
        n = int(nt.value)
        out = dict(tris=np.zeros((n, 9), np.float32), boxes=np.zeros((n, 6), np.float32), material_idx=np.zeros(n, np.int32),
                   normals=np.zeros((n, 9), np.float32), uvs=np.zeros((n, 6), np.float32), colors=np.zeros((n, 12), np.float32))
        rc = L.atlas_rt_aemesh_triangles(h, _addr(out["tris"]), _addr(out["boxes"]), _addr(out["material_idx"]), _addr(out["normals"]),
                                         _addr(out["uvs"]), _addr(out["colors"]))
        if rc != 0:
            raise AtlasError(STATUS.get(rc, rc))
        out["vertex_count"], out["index_count"], out["sub_meshes"] = int(nv.value), int(ni.value), int(ns.value)
        out["materials"] = []
        for k in range(64):
            p = L.atlas_rt_aemesh_material_path(h, k)
            if not p:
                break
            out["materials"].append(p.decode())
        return out
    finally:
        L.atlas_rt_aemesh_close(h)


def comm_unique_id():
    """128-byte NCCL unique id (bytes)."""
    buf = (C.c_uint8 * 128)()
    rc = lib().atlas_rt_comm_unique_id(buf)
    if rc != 0:
        raise AtlasError(f"atlas_rt_comm_unique_id -> {STATUS.get(rc, rc)} (is libnccl.so.2 loadable?)")
    return bytes(buf)


class Comm:
    """atlas_rt_comm: NCCL communicator bound to a Context (one rank per GPU)."""

    def __init__(self, ctx, unique_id, rank, world):
        self.ctx, self.rank, self.world = ctx, rank, world
        h = _vp()
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        ctx.check(ctx.L.atlas_rt_comm_init(ctx.h, buf, rank, world, C.byref(h)))
        self.h = h

    def synchronize(self):
        self.ctx.check(self.ctx.L.atlas_rt_comm_synchronize(self.h))

    def close(self):
        if self.h:
            self.ctx.L.atlas_rt_comm_destroy(self.h)
            self.h = None

    def broadcast_bvh(self, bvh, root):
        out = _vp()
        self.ctx.check(self.ctx.L.atlas_rt_bvh_broadcast(self.h, bvh.h if bvh is not None else None, root, C.byref(out)))
        return bvh if self.rank == root else BVH(self.ctx, out)

    def build_scene_sharded(self, mesh_tris, inst_boxes, inst_records):
        from . import workloads as W
        m = len(mesh_tris)
        tris = [np.ascontiguousarray(t, dtype=np.float32) for t in mesh_tris]
        boxes = [W.tri_boxes(t) for t in tris]
        pa = (_vp * m)(*[_addr(a) for a in boxes])
        pt = (_vp * m)(*[_addr(t) for t in tris])
        pc = (_u64 * m)(*[t.shape[0] for t in tris])
        inst = np.ascontiguousarray(inst_records).view(np.uint32).reshape(-1, 16)
        ib = np.ascontiguousarray(inst_boxes, dtype=np.float32)
        h = _vp()
        self.ctx.check(self.ctx.L.atlas_rt_build_scene_sharded(self.h, m, pa, pt, pc, _addr(inst), _addr(ib), ib.shape[0], 0, C.byref(h)))
        return Scene(self.ctx, h, [], None)

    def replicate_scene(self, scene, root):
        h = _vp()
        self.ctx.check(self.ctx.L.atlas_rt_scene_replicate(self.h, scene.h if scene is not None else None, root, C.byref(h)))
        return scene if self.rank == root else Scene(self.ctx, h, [], None)

    def trace_sharded(self, scene, rays, total_count, hits_out=None, root=0, cull_mask=MASK_ALL, t_min=0.0, t_max=INF, any_hit=False, flags=0):
        """rays: this rank's share (host array or CUDA tensor / device pointer); hits_out (root): (total, 4) float32 host array,
        CUDA tensor or None (a host array is made on root)."""
        fl = flags | (DEVICE_INPUT if _is_device(rays) else 0)
        if self.rank == root:
            if hits_out is None and not (flags & PEER_OUTPUT):   # (peer output: the records may stay in the root's window)
                hits_out = np.empty((total_count, 4), dtype=np.float32)
            if _is_device(hits_out):
                fl |= DEVICE_OUTPUT
        if not _is_device(rays):
            rays = np.ascontiguousarray(rays, dtype=np.float32)
        self.ctx.check(self.ctx.L.atlas_rt_trace_sharded(self.h, scene.h, _addr(rays), total_count, cull_mask, t_min, t_max,
                                                         _addr(hits_out) if (self.rank == root and hits_out is not None) else None, root, fl, int(any_hit)))
        return hits_out if self.rank == root else None

    def peer_hits(self):
        """Device address of the most recent PEER_OUTPUT call's records in the root's window (0 elsewhere / before the first call)."""
        p = _vp()
        self.ctx.check(self.ctx.L.atlas_rt_comm_peer_hits(self.h, C.byref(p)))
        return int(p.value or 0)

    def gather(self, send, nbytes, recv, sizes, offsets, root=0, flags=0):
        sz = (_u64 * self.world)(*[int(x) for x in sizes])
        of = (_u64 * self.world)(*[int(x) for x in offsets])
        self.ctx.check(self.ctx.L.atlas_rt_comm_gather(self.h, _addr(send), nbytes, _addr(recv) if self.rank == root else None, sz, of, root, flags))


def shard_range(count, rank, world, align=64):
    b, e = _u64(), _u64()
    rc = lib().atlas_rt_shard_range(count, rank, world, align, C.byref(b), C.byref(e))
    if rc != 0:
        raise AtlasError(STATUS.get(rc, rc))
    return int(b.value), int(e.value)
