"""Synthetic inputs for the BASELINE.json configurations and the parity edge cases.

Everything here is plain numpy on the host and deterministic for a given seed; the same arrays are fed to the CUDA
path, to the oracle and to the reference so parity never depends on how an input was made.

Conventions: triangles are (n, 9) float32 (v0, v1, v2); boxes are (n, 6) float32 (min, max) computed like
MeshData::BuildBVH does (/root/reference/src/engine/mesh/MeshData.cpp:143-146); rays are PackedRay rows of 12 float32
(/root/reference/data/shader/raytracer/structures.hsh:9-13): origin.xyz, bits(ID), direction.xyz, -, hit.xyzw.
"""
import numpy as np

INF = np.float32(1e12)            # common.hsh:8
MASK_ALL = 1 << 7                 # RTStructures.h:9-12
MASK_SHADOW = 1 << 6


def tri_boxes(tris):
    t = np.asarray(tris, dtype=np.float32).reshape(-1, 3, 3)
    return np.concatenate([t.min(axis=1), t.max(axis=1)], axis=1).astype(np.float32)


# --------------------------------------------------------------------------------------------------- geometry
def soup(n, seed=1234, extent=0.01):
    """C2: centre ~ U[0,1)^3, each vertex = centre + extent * U[-0.5,0.5)^3."""
    rng = np.random.default_rng(seed)
    c = rng.random((n, 1, 3), dtype=np.float32)
    off = (rng.random((n, 3, 3), dtype=np.float32) - np.float32(0.5)) * np.float32(extent)
    return (c + off).reshape(n, 9).astype(np.float32)


def soup_with_giants(n, seed=99, frac=0.02):
    """Soup where a fraction of triangles is huge, so the root SBVH spatial split fires and duplicates refs."""
    t = soup(n, seed).reshape(n, 3, 3)
    rng = np.random.default_rng(seed + 1)
    k = max(1, int(n * frac))
    pick = rng.choice(n, k, replace=False)
    c = rng.random((k, 1, 3), dtype=np.float32)
    t[pick] = c + (rng.random((k, 3, 3), dtype=np.float32) - np.float32(0.5)) * np.float32(0.6)
    return t.reshape(n, 9)


def heightfield(nx, nz, spacing=1.0):
    """C3: nx x nz quads = 2*nx*nz triangles, y = 20 sin(.013x) cos(.017z) + 5 sin(.11x + .07z)."""
    xs = (np.arange(nx + 1, dtype=np.float64) * spacing)
    zs = (np.arange(nz + 1, dtype=np.float64) * spacing)
    X, Z = np.meshgrid(xs, zs, indexing="ij")
    Y = 20.0 * np.sin(0.013 * X) * np.cos(0.017 * Z) + 5.0 * np.sin(0.11 * X + 0.07 * Z)
    P = np.stack([X, Y, Z], axis=-1).astype(np.float32)
    p00, p10, p01, p11 = P[:-1, :-1], P[1:, :-1], P[:-1, 1:], P[1:, 1:]
    a = np.stack([p00, p10, p11], axis=2)
    b = np.stack([p00, p11, p01], axis=2)
    return np.stack([a, b], axis=2).reshape(-1, 9).astype(np.float32)


def flat_grid(n):
    """n x n quads in the plane y == 0: one axis is always skipped (extent < 1e-3) and quads share boxes."""
    t = heightfield(n, n).reshape(-1, 3, 3)
    t[:, :, 1] = 0.0
    return t.reshape(-1, 9)


def uv_sphere(segments=32, rings=16, radius=1.0, centre=(0.0, 0.0, 0.0)):
    th = np.linspace(0.0, np.pi, rings + 1)
    ph = np.linspace(0.0, 2.0 * np.pi, segments + 1)
    T, Pn = np.meshgrid(th, ph, indexing="ij")
    P = np.stack([np.sin(T) * np.cos(Pn), np.cos(T), np.sin(T) * np.sin(Pn)], axis=-1) * radius + np.asarray(centre)
    P = P.astype(np.float32)
    tris = []
    for i in range(rings):
        for j in range(segments):
            a, b, c, d = P[i, j], P[i + 1, j], P[i + 1, j + 1], P[i, j + 1]
            if i != 0:
                tris.append(np.concatenate([a, b, d]))
            if i != rings - 1:
                tris.append(np.concatenate([b, c, d]))
    return np.asarray(tris, dtype=np.float32)


def _quad_wall(origin, du, dv, nu, nv):
    o, du, dv = (np.asarray(x, dtype=np.float64) for x in (origin, du, dv))
    U, V = np.meshgrid(np.arange(nu + 1) / nu, np.arange(nv + 1) / nv, indexing="ij")
    P = (o + U[..., None] * du + V[..., None] * dv).astype(np.float32)
    p00, p10, p01, p11 = P[:-1, :-1], P[1:, :-1], P[:-1, 1:], P[1:, 1:]
    a = np.stack([p00, p10, p11], axis=2)
    b = np.stack([p00, p11, p01], axis=2)
    return np.stack([a, b], axis=2).reshape(-1, 9)


def atrium(detail=64, seed=7, scale=0.05, clutter=20000):
    """C1 stand-in for the missing sponza: a closed hall (tessellated floor, ceiling, four walls), two rows of
    12-sided columns, a few long thin triangles that span the hall (so the root spatial split fires) and small
    clutter triangles. detail=64 gives ~75 k triangles, detail=128 ~262 k."""
    parts = []
    L, W, H = 600.0, 250.0, 200.0
    n = detail
    parts.append(_quad_wall((0, 0, 0), (L, 0, 0), (0, 0, W), 2 * n, n))            # floor
    parts.append(_quad_wall((0, H, 0), (L, 0, 0), (0, 0, W), 2 * n, n))            # ceiling
    parts.append(_quad_wall((0, 0, 0), (L, 0, 0), (0, H, 0), 2 * n, n))            # wall z=0
    parts.append(_quad_wall((0, 0, W), (L, 0, 0), (0, H, 0), 2 * n, n))            # wall z=W
    parts.append(_quad_wall((0, 0, 0), (0, 0, W), (0, H, 0), n, n))                # wall x=0
    parts.append(_quad_wall((L, 0, 0), (0, 0, W), (0, H, 0), n, n))                # wall x=L
    ang = np.linspace(0.0, 2.0 * np.pi, 13)
    for row_z in (60.0, W - 60.0):
        for cx in np.linspace(60.0, L - 60.0, 8):
            ring = np.stack([cx + 12.0 * np.cos(ang), np.zeros(13), row_z + 12.0 * np.sin(ang)], axis=-1)
            for k in range(12):
                for seg in range(max(2, n // 8)):
                    y0 = H * 0.8 * seg / max(2, n // 8)
                    y1 = H * 0.8 * (seg + 1) / max(2, n // 8)
                    a, b = ring[k].copy(), ring[k + 1].copy()
                    a0, b0, a1, b1 = a.copy(), b.copy(), a.copy(), b.copy()
                    a0[1] = b0[1] = y0
                    a1[1] = b1[1] = y1
                    parts.append(np.concatenate([a0, b0, b1])[None].astype(np.float32))
                    parts.append(np.concatenate([a0, b1, a1])[None].astype(np.float32))
    rng = np.random.default_rng(seed)
    beams = []
    for _ in range(100):   # long thin triangles (banners/beams)
        p = rng.random(3) * (L, H, W)
        q = p + (rng.random(3) - 0.5) * (L * 0.9, 10.0, W * 0.9)
        r = p + (0.0, 2.0, 0.0)
        beams.append(np.concatenate([p, q, r]))
    parts.append(np.asarray(beams, dtype=np.float32))
    c = rng.random((clutter, 1, 3)) * (L, H, W)
    parts.append((c + (rng.random((clutter, 3, 3)) - 0.5) * 3.0).reshape(clutter, 9).astype(np.float32))
    t = np.concatenate(parts, axis=0).astype(np.float32)
    return (t * np.float32(scale)).astype(np.float32)


def coincident(n_same=200, n_concentric=200):
    """Pathological: n_same identical triangles plus n_concentric concentric ones (forces the std::sort fallback
    of PerformMedianSplit with large n)."""
    base = np.array([0, 0, 0, 1, 0, 0, 0, 1, 0], dtype=np.float32)
    same = np.tile(base, (n_same, 1))
    s = (1.0 + np.arange(n_concentric, dtype=np.float32)[:, None] * np.float32(0.01))
    conc = (base.reshape(3, 3)[None] - np.float32(1 / 3)) * s[:, :, None] + np.float32(1 / 3)
    return np.concatenate([same, conc.reshape(n_concentric, 9)]).astype(np.float32)


# ------------------------------------------------------------------------------------------------------- rays
def pack_rays(origins, dirs, ids=None, t=None):
    n = origins.shape[0]
    r = np.zeros((n, 12), dtype=np.float32)
    r[:, 0:3] = origins
    r[:, 4:7] = dirs
    ids = np.arange(n, dtype=np.int32) if ids is None else np.asarray(ids, dtype=np.int32)
    r[:, 3] = ids.view(np.float32)
    if t is not None:
        r[:, 8] = t
    r[:, 9] = np.full(n, -1, dtype=np.int32).view(np.float32)
    return r


def random_rays(n, lo, hi, seed=5678):
    """C2 rays: origin ~ U over the box, direction = normalize(U[-0.5,0.5)^3), rejecting |d| < 1e-3 or any
    component below 1e-6 in magnitude."""
    rng = np.random.default_rng(seed)
    lo, hi = np.asarray(lo, dtype=np.float32), np.asarray(hi, dtype=np.float32)
    o = lo + rng.random((n, 3), dtype=np.float32) * (hi - lo)
    d = rng.random((n, 3), dtype=np.float32) - np.float32(0.5)
    for _ in range(8):
        nrm = np.linalg.norm(d, axis=1)
        bad = (nrm < 1e-3) | (np.abs(d / np.maximum(nrm, 1e-20)[:, None]).min(axis=1) < 1e-6)
        if not bad.any():
            break
        d[bad] = rng.random((int(bad.sum()), 3), dtype=np.float32) - np.float32(0.5)
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    return pack_rays(o.astype(np.float32), d)


def camera_frame(eye, target, up=(0, 1, 0), fov_deg=47.0, aspect=16 / 9):
    """origin/right/bottom of the near-plane rectangle as PathTracingRenderer.cpp:157-162 derives them from frustum
    corners 4,5,6 (upper-left, upper-right, lower-left of the near plane, CameraComponent.cpp:125-153)."""
    eye, target, up = (np.asarray(x, dtype=np.float64) for x in (eye, target, up))
    f = target - eye
    f /= np.linalg.norm(f)
    r = np.cross(f, up)
    r /= np.linalg.norm(r)
    u = np.cross(r, f)
    near = 0.1
    hh = np.tan(np.radians(fov_deg) / 2.0) * near
    hw = hh * aspect
    c = eye + f * near
    ul, ur, ll = c + u * hh - r * hw, c + u * hh + r * hw, c - u * hh - r * hw
    return (eye.astype(np.float32), ul.astype(np.float32), (ur - ul).astype(np.float32), (ll - ul).astype(np.float32))


def primary_rays(width, height, eye, origin, right, bottom, jitter=(0.5, 0.5), tile_order=False):
    """rayGen.csh:25-91 without the storage permutation unless tile_order: dir = normalize(origin + right*u +
    bottom*v - eye), ID = y*w + x."""
    x, y = np.meshgrid(np.arange(width, dtype=np.float32), np.arange(height, dtype=np.float32), indexing="xy")
    u = (x + np.float32(jitter[0])) / np.float32(width)
    v = (y + np.float32(jitter[1])) / np.float32(height)
    d = origin[None, None] + right[None, None] * u[..., None] + bottom[None, None] * v[..., None] - eye[None, None]
    d = (d / np.linalg.norm(d, axis=-1, keepdims=True)).astype(np.float32).reshape(-1, 3)
    o = np.broadcast_to(eye, d.shape).astype(np.float32)
    ids = (y.astype(np.int32) * width + x.astype(np.int32)).reshape(-1)
    rays = pack_rays(o, d, ids)
    if tile_order and width % 8 == 0 and height % 8 == 0:
        yy, xx = np.divmod(np.arange(width * height), width)
        key = ((yy // 8) * (width // 8) + (xx // 8)) * 64 + (yy % 8) * 8 + (xx % 8)
        rays = rays[np.argsort(key, kind="stable")]
    return rays


# -------------------------------------------------------------------------------------------------- instances
def transform_box(box, M):
    """AABB::Transform (/root/reference/src/engine/volume/AABB.cpp:35-60): transform 8 corners, take min/max.
    Computed in float64 and rounded once; instance world boxes are harness INPUTS."""
    lo, hi = box[:3].astype(np.float64), box[3:].astype(np.float64)
    corners = np.array([[x, y, z, 1.0] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])])
    w = corners @ M.T
    return np.concatenate([w[:, :3].min(axis=0), w[:, :3].max(axis=0)]).astype(np.float32)


def random_instances(n, mesh_boxes, seed=4242, extent=(2000.0, 200.0, 2000.0), scale=(0.5, 4.0)):
    """C4: n instances of len(mesh_boxes) meshes. Returns (world boxes (n,6) f32, GPUBVHInstance records (n,16) u32)
    with inverseMatrix = rows 0-2 of the inverse world matrix (RayTracingWorld.cpp:100), mask = MaskAll|MaskShadow."""
    rng = np.random.default_rng(seed)
    m = len(mesh_boxes)
    boxes = np.zeros((n, 6), dtype=np.float32)
    inst = np.zeros((n, 16), dtype=np.uint32)
    for i in range(n):
        mesh = int(rng.integers(0, m))
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        w, x, y, z = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        s = rng.uniform(*scale)
        M = np.eye(4)
        M[:3, :3] = R * s
        M[:3, 3] = rng.random(3) * np.asarray(extent)
        boxes[i] = transform_box(mesh_boxes[mesh], M)
        inv = np.linalg.inv(M)
        inst[i, :12] = inv[:3, :].astype(np.float32).reshape(-1).view(np.uint32)
        inst[i, 12] = mesh
        inst[i, 13] = 0
        inst[i, 14] = np.uint32(0xFFFFFFFF)
        inst[i, 15] = MASK_ALL | MASK_SHADOW
    return boxes, inst


def identity_instance(mesh=0, mask=MASK_ALL | MASK_SHADOW):
    inst = np.zeros((1, 16), dtype=np.uint32)
    inst[0, :12] = np.eye(4, dtype=np.float32)[:3].reshape(-1).view(np.uint32)
    inst[0, 12] = mesh
    inst[0, 14] = np.uint32(0xFFFFFFFF)
    inst[0, 15] = mask
    return inst


# ------------------------------------------------------------------------------------------------------ packing
def pack_bvh_triangles(tris, order, end_of_node, material_idx=0, opacity=1.0):
    """GPUBVHTriangle rows (12 float32) as MeshData.cpp:242-247 emits them: v0.xyz, endOfNode ? 1 : -1; v1.xyz,
    bits(materialIdx); v2.xyz, opacity. numpy mirror used by tests to cross-check the CUDA pack kernel."""
    t = np.asarray(tris, dtype=np.float32).reshape(-1, 9)[np.asarray(order, dtype=np.int64)]
    out = np.zeros((t.shape[0], 12), dtype=np.float32)
    out[:, 0:3] = t[:, 0:3]
    out[:, 3] = np.where(np.asarray(end_of_node) != 0, np.float32(1.0), np.float32(-1.0))
    out[:, 4:7] = t[:, 3:6]
    out[:, 7] = np.full(t.shape[0], material_idx, dtype=np.int32).view(np.float32)
    out[:, 8:11] = t[:, 6:9]
    out[:, 11] = np.float32(opacity)
    return out


def pack_shading_triangles(tris, order, end_of_node, material_idx=None, opacity=None, payload11=None):
    """GPUTriangle rows (24 float32) as MeshData.cpp:230-239 lays them out; payload11 = the engine's packed shading
    words per SOURCE triangle (n0,n1,n2,uv0,uv1,uv2,t,bt,c0,c1,c2), moved verbatim. numpy mirror of the CUDA pack."""
    order = np.asarray(order, dtype=np.int64)
    t = np.asarray(tris, dtype=np.float32).reshape(-1, 9)[order]
    n = t.shape[0]
    p = np.zeros((n, 11), dtype=np.uint32) if payload11 is None else np.asarray(payload11, dtype=np.uint32)[order]
    mat = np.zeros(n, dtype=np.int32) if material_idx is None else np.asarray(material_idx, dtype=np.int32)[order]
    op = np.ones(n, dtype=np.float32) if opacity is None else np.asarray(opacity, dtype=np.float32)[order]
    out = np.zeros((n, 24), dtype=np.uint32)
    tv = t.view(np.uint32)
    out[:, 0:3], out[:, 3] = tv[:, 0:3], p[:, 0]
    out[:, 4:7], out[:, 7] = tv[:, 3:6], p[:, 1]
    out[:, 8:11], out[:, 11] = tv[:, 6:9], p[:, 2]
    out[:, 12:15], out[:, 15] = p[:, 3:6], mat.view(np.uint32)
    out[:, 16:18] = p[:, 6:8]
    out[:, 18] = np.where(np.asarray(end_of_node) != 0, np.float32(1.0), np.float32(-1.0)).astype(np.float32).view(np.uint32)
    out[:, 20:23], out[:, 23] = p[:, 8:11], op.view(np.uint32)
    return out.view(np.float32)


def geometric_clusters(clusters=150, per_cluster=1500, ratio=1.5, seed=77):
    """Clusters of small triangles at x = ratio^k: every binned split peels only the farthest cluster(s) off, so the tree
    is a long chain of big nodes — dozens of builder levels instead of ~log2(n)."""
    rng = np.random.default_rng(seed)
    out = np.empty((clusters * per_cluster, 3, 3), dtype=np.float32)
    for k in range(clusters):
        x = np.float32(ratio) ** np.float32(k)
        c = rng.random((per_cluster, 1, 3), dtype=np.float32) * np.float32(0.05) * x
        c[:, :, 0] += x
        off = (rng.random((per_cluster, 3, 3), dtype=np.float32) - np.float32(0.5)) * np.float32(0.002) * x
        out[k * per_cluster:(k + 1) * per_cluster] = c + off
    return out.reshape(-1, 9)


def smooth_normals(tris):
    """Per-corner vertex normals (n, 9): area-weighted face normals averaged over corners that share a position, normalised —
    what a modelling tool exports and MeshData keeps in `normals` (inputs to the packed shading words)."""
    t = np.asarray(tris, dtype=np.float32).reshape(-1, 3, 3)
    fn = np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0]).astype(np.float64)
    keys = np.ascontiguousarray(t.reshape(-1, 3)).view([("", np.float32)] * 3).reshape(-1)
    _, inv = np.unique(keys, return_inverse=True)
    acc = np.zeros((inv.max() + 1, 3))
    np.add.at(acc, inv, np.repeat(fn, 3, axis=0))
    nrm = acc[inv]
    ln = np.linalg.norm(nrm, axis=1, keepdims=True)
    nrm = np.where(ln > 1e-20, nrm / np.maximum(ln, 1e-20), np.array([0.0, 1.0, 0.0]))
    return nrm.astype(np.float32).reshape(-1, 9)


def planar_uvs(tris, scale=1.0):
    """(n, 6) texture coordinates: xz of every corner times `scale`."""
    t = np.asarray(tris, dtype=np.float32).reshape(-1, 3, 3)
    return (t[:, :, [0, 2]] * np.float32(scale)).reshape(-1, 6).astype(np.float32)


# ------------------------------------------------------------------------------- other in-tree callers
# Ray distributions of the engine's other users of HitClosest / HitAny, generated from a G-buffer (world positions P, unit
# normals N, unit view vectors V = towards the camera) the way their shaders do. These are workload generators: the same
# arrays go to the CUDA path and to the oracle, so their own arithmetic (float64 here) is not a parity matter.
EPSILON = 0.1   # raytracer/common.hsh:9


def _basis(N):
    up = np.where((np.abs(N[:, 2]) < 0.999)[:, None], np.array([0.0, 0.0, 1.0]), np.array([1.0, 0.0, 0.0]))
    t = np.cross(up, N)
    t /= np.maximum(np.linalg.norm(t, axis=1, keepdims=True), 1e-30)
    return t, np.cross(N, t)


def _cosine_dirs(N, u):
    """ImportanceSampleCosDir / SampleDiffuseBRDF (brdf/importanceSample.hsh:85-104, brdf/brdfSample.hsh:8-32)."""
    N = np.asarray(N, dtype=np.float64)
    r, phi = np.sqrt(u[:, 0]), 2.0 * np.pi * u[:, 1]
    t, b = _basis(N)
    d = t * (r * np.cos(phi))[:, None] + b * (r * np.sin(phi))[:, None] + N * np.sqrt(1.0 - u[:, 0])[:, None]
    return d / np.linalg.norm(d, axis=1, keepdims=True)


def rtao_rays(P, N, u, radius):
    """ao/rtao.csh:81-100: one cosine-distributed ray per pixel, origin = P + dir * EPSILON + N * EPSILON, any hit within
    `radius` (per-ray tMax in hit.x), cull mask INSTANCE_MASK_ALL."""
    d = _cosine_dirs(N, u)
    o = np.asarray(P, np.float64) + d * EPSILON + np.asarray(N, np.float64) * EPSILON
    return pack_rays(o.astype(np.float32), d.astype(np.float32), t=np.full(len(o), radius, np.float32))


def rtgi_rays(P, N, u, view_dist, bias=0.0):
    """rtgi/rtgi.csh:110-137: cosine-distributed ray, origin offset scaled by the view distance, closest hit up to INF."""
    uu = u.copy()
    uu[:, 0] *= (1.0 - bias)
    d = _cosine_dirs(N, uu)
    off = np.maximum(1.0, np.asarray(view_dist, np.float64))[:, None]
    o = np.asarray(P, np.float64) + d * EPSILON * off * 0.01 + np.asarray(N, np.float64) * EPSILON * 0.01 * off
    return pack_rays(o.astype(np.float32), d.astype(np.float32))


def reflection_rays(P, N, V, roughness, u, view_dist, bias=0.0):
    """reflection/rtreflection.csh:104-141: GGX visible-normal importance sample (brdf/importanceSample.hsh:36-83) for
    roughness > 0.01, the mirror direction otherwise; origin offset scaled by the view distance; closest hit up to INF."""
    P, N, V = (np.asarray(x, np.float64) for x in (P, N, V))
    alpha = roughness * roughness
    uu = u.copy()
    uu[:, 1] *= (1.0 - bias)
    t, b = _basis(N)
    b = b / np.maximum(np.linalg.norm(b, axis=1, keepdims=True), 1e-30)
    Vt = np.stack([(V * t).sum(1), (V * b).sum(1), (V * N).sum(1)], axis=1)
    Vh = np.stack([Vt[:, 0] * alpha, Vt[:, 1] * alpha, Vt[:, 2]], axis=1)
    Vh /= np.linalg.norm(Vh, axis=1, keepdims=True)
    phi = 2.0 * np.pi * uu[:, 0]
    z = (1.0 - uu[:, 1]) * (1.0 + Vh[:, 2]) - Vh[:, 2]
    s = np.sqrt(np.clip(1.0 - z * z, 0.0, 1.0))
    H = np.stack([s * np.cos(phi), s * np.sin(phi), z], axis=1) + Vh
    Hs = np.stack([H[:, 0] * alpha, H[:, 1] * alpha, np.maximum(H[:, 2], 0.0)], axis=1)
    Hw = t * Hs[:, 0:1] + b * Hs[:, 1:2] + N * Hs[:, 2:3]
    Hw /= np.maximum(np.linalg.norm(Hw, axis=1, keepdims=True), 1e-30)
    d = 2.0 * (V * Hw).sum(1, keepdims=True) * Hw - V
    if roughness <= 0.01:
        d = 2.0 * (V * N).sum(1, keepdims=True) * N - V
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-30)
    off = np.maximum(1.0, np.asarray(view_dist, np.float64))[:, None] * 0.1
    o = P + d * EPSILON * 0.1 * off + N * EPSILON * off * 0.1
    ids = np.where((d * N).sum(1) >= 0.0, np.arange(len(P)), -1)      # rays below the surface are not cast
    return pack_rays(o.astype(np.float32), d.astype(np.float32), ids=ids)


def ddgi_rays(lo, hi, probes=(12, 6, 12), rays_per_probe=128, inactive_every=7, seed=0):
    """ddgi/rayGen.csh:44-83: every probe of a regular grid casts `rays_per_probe` rays along a spherical Fibonacci set
    rotated by a random rotation; rays a probe does not use (inactive probes cast fewer) carry ID = -1; ID = probe * rays + i."""
    rng = np.random.default_rng(seed)
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    i = np.arange(rays_per_probe) + 0.5
    phi = 2.0 * np.pi * ((i * 0.6180339887498949) % 1.0)
    ct = 1.0 - 2.0 * i / rays_per_probe
    st = np.sqrt(np.clip(1.0 - ct * ct, 0.0, 1.0))
    dirs = (np.stack([st * np.cos(phi), st * np.sin(phi), ct], axis=1) @ R.T)
    gx, gy, gz = (np.linspace(lo[k], hi[k], probes[k] + 2)[1:-1] for k in range(3))
    G = np.stack(np.meshgrid(gx, gy, gz, indexing="ij"), axis=-1).reshape(-1, 3)
    o = np.repeat(G, rays_per_probe, axis=0)
    d = np.tile(dirs, (len(G), 1))
    ids = np.arange(len(o))
    probe = ids // rays_per_probe
    inactive = (probe % inactive_every == 0) & ((ids % rays_per_probe) >= rays_per_probe // 4)   # inactive probes cast a quarter of the rays
    return pack_rays(o.astype(np.float32), d.astype(np.float32), ids=np.where(inactive, -1, ids))


def gbuffer_from_hits(rays_out, tris_by_mesh_slot=None, normals=None):
    """World positions of primary hits and the view vectors towards the camera: P = o + t d, V = -d; rows without a hit are
    dropped. Returns (P, V, view_dist, keep_mask)."""
    hit = rays_out[:, 9].view(np.int32) >= 0
    o, d, t = rays_out[hit, 0:3].astype(np.float64), rays_out[hit, 4:7].astype(np.float64), rays_out[hit, 8].astype(np.float64)
    return o + d * t[:, None], -d, t, hit
