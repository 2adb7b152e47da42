// AtlasRT.h — host-side C++ mirror of the reference interfaces on the ray-tracing acceleration path, implemented on
// top of the C ABI in include/atlas_rt.h. Same class names, member names, argument meaning and error behaviour as the
// reference, so engine code that uses them compiles unchanged when this header replaces the originals:
//
//   Atlas::Volume::AABB, BVHNode, BVHTriangle, BVHBuilder::Ref, BVH     src/engine/volume/AABB.h, BVH.h:14-138
//   Atlas::GPUBVHNode / GPUBVHTriangle / GPUBVHInstance                 src/engine/raytracing/RTStructures.h:23-103
//   Atlas::RayTracing::BuildMeshBVH                                      the BVH half of Mesh::MeshData::BuildBVH, mesh/MeshData.cpp:63-271
//   Atlas::RayTracing::UpdateForSoftwareRayTracing                       raytracing/RayTracingWorld.cpp:267-307
//   Atlas::RayTracing::Tracer                                            the trace-batch boundary of RayTracingHelper::DispatchHitClosest
//
// Inside the engine define ATLAS_RT_USE_GLM before including this header so that vec3 / mat3x4 are the glm types the
// rest of the engine uses (layouts are identical); stand-alone it brings its own PODs.
//
// Threading: like the reference constructors, everything here is synchronous and may be called concurrently from
// several job-system workers; each calling thread lazily gets its own atlas_rt_context (CUDA stream). Long-lived objects
// (MeshBVH, World) may be built on one thread and used or released on another: the library accepts objects from any
// context of the same device and keeps a context alive until the last object created on it has been freed, so a worker
// thread may exit while its meshes live on (the engine builds meshes on job-system workers, MeshData::BuildBVH, and
// assembles the scene elsewhere, RayTracingWorld::UpdateForSoftwareRayTracing).
// Errors: the reference has no error channel on this path (size mismatch => silently empty BVH, BVH.cpp:18-19); the
// same holds here, and a CUDA failure additionally leaves the object empty with the message in Atlas::RayTracing::LastError().
#pragma once

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "../../include/atlas_rt.h"

#ifdef ATLAS_RT_USE_GLM
#include <glm/glm.hpp>
#endif

namespace Atlas {

#ifdef ATLAS_RT_USE_GLM
using glm::vec3;
using glm::vec4;
using glm::mat3x4;
#else
struct vec3 {
    float x = 0.0f, y = 0.0f, z = 0.0f;
    vec3() = default;
    explicit vec3(float s) : x(s), y(s), z(s) {}
    vec3(float x, float y, float z) : x(x), y(y), z(z) {}
    float& operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
};
struct vec4 {
    float x = 0.0f, y = 0.0f, z = 0.0f, w = 0.0f;
    vec4() = default;
    vec4(float x, float y, float z, float w) : x(x), y(y), z(z), w(w) {}
    vec4(const vec3& v, float w) : x(v.x), y(v.y), z(v.z), w(w) {}
};
struct mat3x4 {   // 3 columns of vec4, as glm::mat3x4 / the std430 `mat3x4 inverseMatrix`
    vec4 c[3];
    vec4& operator[](int i) { return c[i]; }
    const vec4& operator[](int i) const { return c[i]; }
};
#endif

enum InstanceCullMasks { MaskAll = 1 << 7, MaskShadow = 1 << 6 };   // RTStructures.h:9-12

struct GPUAABB { vec3 min; vec3 max; };
struct GPUTriangle { vec4 v0; vec4 v1; vec4 v2; vec4 d0; vec4 d1; vec4 d2; };   // RTStructures.h:14-21
struct GPUBVHTriangle { vec4 v0; vec4 v1; vec4 v2; };
struct GPUBVHInstance {
    mat3x4 inverseMatrix;
    int32_t meshOffset = 0;
    int32_t materialOffset = 0;
    int32_t nextInstance = 0;
    uint32_t mask = 0;
};
struct GPUBVHNode {
    GPUAABB leftAABB;
    GPUAABB rightAABB;
    int32_t leftPtr = 0;
    int32_t rightPtr = 0;
    int32_t padding0 = 0;
    int32_t padding1 = 0;
};
struct PackedRay { vec4 origin; vec4 direction; vec4 hit; };   // data/shader/raytracer/structures.hsh:9-13

static_assert(sizeof(GPUTriangle) == 96 && sizeof(GPUBVHTriangle) == 48 && sizeof(GPUBVHInstance) == 64 && sizeof(GPUBVHNode) == 64 && sizeof(PackedRay) == 48, "GPU layouts");

namespace RayTracing {

namespace detail {
inline std::string& ErrorSlot() { static thread_local std::string e; return e; }
struct ThreadContext {
    atlas_rt_context* ctx = nullptr;
    ThreadContext() {
        int device = 0;
        if (const char* e = getenv("ATLAS_RT_DEVICE")) device = atoi(e);
        if (atlas_rt_context_create(device, nullptr, &ctx) != ATLAS_RT_OK) { ctx = nullptr; ErrorSlot() = "atlas_rt_context_create failed (no CUDA device?)"; }
    }
    ~ThreadContext() { if (ctx) atlas_rt_context_destroy(ctx); }
};
inline atlas_rt_context* Context() { static thread_local ThreadContext t; return t.ctx; }
inline bool Check(int rc) {
    if (rc == ATLAS_RT_OK) return true;
    atlas_rt_context* c = Context();
    ErrorSlot() = c ? atlas_rt_last_error(c) : "no context";
    return false;
}
}   // namespace detail

inline const std::string& LastError() { return detail::ErrorSlot(); }

}   // namespace RayTracing

namespace Volume {

class AABB {   // volume/AABB.h:15-105 (data members only; the builder needs nothing else from the host side)
public:
    AABB() = default;
    AABB(vec3 min, vec3 max) : min(min), max(max) {}
    vec3 min = vec3(0.0f);
    vec3 max = vec3(0.0f);
};
static_assert(sizeof(AABB) == 24, "AABB layout");

class BVHNode {   // volume/BVH.h:14-24
public:
    BVHNode() {}
    AABB leftAABB;
    AABB rightAABB;
    int32_t leftPtr = 0;
    int32_t rightPtr = 0;
};
static_assert(sizeof(BVHNode) == 56, "BVHNode layout");

class BVHTriangle {   // volume/BVH.h:26-33
public:
    vec3 v0;
    vec3 v1;
    vec3 v2;
    uint32_t idx;
    bool endOfNode = false;
};

class BVHBuilder {   // only the nested Ref type is part of the public surface (BVH::refs)
public:
    struct Ref {
        uint32_t idx = 0;
        uint32_t nodeIdx = 0;
        bool endOfNode = false;
        AABB aabb;
    };
};

class Ray {   // volume/Ray.h: what BVH::GetIntersection* read
public:
    Ray() = default;
    Ray(vec3 origin, vec3 direction, float tMin = 0.0f, float tMax = 2048.0f) : origin(origin), direction(direction), tMin(tMin), tMax(tMax) {}
    vec3 origin = vec3(0.0f);
    vec3 direction = vec3(0.0f, 1.0f, 0.0f);
    float tMin = 0.0f;
    float tMax = 2048.0f;
};

class BVH {   // volume/BVH.h:114-138
public:
    BVH() = default;

    // BLAS — BVH.cpp:14-56. Results in nodes / data / aabbs; `refs` stays empty like in the reference.
    BVH(const std::vector<AABB>& aabbs, const std::vector<BVHTriangle>& data, bool parallelBuild = true) {
        (void)parallelBuild;   // the GPU build has no serial mode; both reference modes give identical trees anyway
        if (aabbs.size() != data.size()) return;
        atlas_rt_context* ctx = RayTracing::detail::Context();
        if (!ctx) return;
        std::vector<float> tris(data.size() * 9);
        for (size_t i = 0; i < data.size(); i++) {
            const BVHTriangle& t = data[i];
            const float v[9] = {t.v0.x, t.v0.y, t.v0.z, t.v1.x, t.v1.y, t.v1.z, t.v2.x, t.v2.y, t.v2.z};
            std::memcpy(&tris[9 * i], v, sizeof(v));
        }
        atlas_rt_bvh* h = nullptr;
        if (!RayTracing::detail::Check(atlas_rt_build_blas(ctx, reinterpret_cast<const float*>(aabbs.data()), tris.data(), data.size(), 0, &h))) return;
        std::vector<uint32_t> order;
        std::vector<uint8_t> flags;
        if (Fetch(h, order, flags)) {
            this->aabbs.resize(order.size());
            this->data.resize(order.size());
            for (size_t i = 0; i < order.size(); i++) {   // BVH.cpp:47-52
                this->aabbs[i] = aabbs[order[i]];
                this->data[i] = data[order[i]];
                this->data[i].endOfNode = flags[i] != 0;
            }
        }
        atlas_rt_bvh_free(h);
    }

    // TLAS — BVH.cpp:58-101. Results in nodes / refs / aabbs.
    BVH(const std::vector<AABB>& aabbs, bool parallelBuild = true) {
        (void)parallelBuild;
        atlas_rt_context* ctx = RayTracing::detail::Context();
        if (!ctx) return;
        atlas_rt_bvh* h = nullptr;
        if (!RayTracing::detail::Check(atlas_rt_build_tlas(ctx, reinterpret_cast<const float*>(aabbs.data()), aabbs.size(), 0, &h))) return;
        std::vector<uint32_t> order;
        std::vector<uint8_t> flags;
        if (Fetch(h, order, flags)) {
            refs.resize(order.size());
            this->aabbs.resize(order.size());
            for (size_t i = 0; i < order.size(); i++) {
                refs[i].idx = order[i];
                refs[i].endOfNode = flags[i] != 0;
                refs[i].aabb = aabbs[order[i]];
                this->aabbs[i] = aabbs[order[i]];
            }
            FillNodeIdx();
        }
        atlas_rt_bvh_free(h);
    }

    // BVH.cpp:103-174. The closest triangle hit by `ray` within (ray.tMin, ray.tMax): `closest` = data[slot],
    // intersection = (t, u, v) with weights v0: 1-u-v, v1: u, v2: v; intersection.x = ray.tMax when nothing is hit.
    // Runs as a one-ray batch through atlas_rt_trace_closest over a device copy of this tree that is created on first use
    // from the public members (so a tree assigned by hand works too) — there is no CPU traversal in this library. The
    // reference's return value is unreliable (it compares against a tMax it has shrunk itself); here it is "something
    // was hit". `stack` is unused (the traversal stack lives in the kernel).
    bool GetIntersection(std::vector<std::pair<int32_t, float>>& stack, Ray ray, BVHTriangle& closest, vec3& intersection) {
        (void)stack;
        intersection.x = ray.tMax;
        PackedHit hit;
        if (!TraceOne(ray, false, hit) || hit.id < 0 || size_t(hit.id) >= data.size()) return false;
        closest = data[size_t(hit.id)];
        intersection = vec3(hit.t, hit.u, hit.v);
        return true;
    }

    // BVH.cpp:176-217: is anything hit before ray.tMax?
    bool GetIntersectionAny(std::vector<std::pair<int32_t, float>>& stack, Ray ray) {
        (void)stack;
        PackedHit hit;
        return TraceOne(ray, true, hit) && hit.id >= 0;
    }

    std::vector<BVHNode>& GetTree() { return nodes; }

    void Clear() {   // declared in volume/BVH.h:130
        aabbs.clear(); aabbs.shrink_to_fit();
        data.clear(); data.shrink_to_fit();
        refs.clear(); refs.shrink_to_fit();
        nodes.clear(); nodes.shrink_to_fit();
        query.reset();
    }

    std::vector<AABB> aabbs;
    std::vector<BVHTriangle> data;
    std::vector<BVHBuilder::Ref> refs;
    std::vector<BVHNode> nodes;

private:
    struct PackedHit { float t = 0.0f, u = 0.0f, v = 0.0f; int32_t id = -1; };
    // Device copy used by the CPU-query methods; shared between copies of the BVH object, released with the last one.
    struct Query {
        atlas_rt_bvh* blas = nullptr;
        atlas_rt_mesh* mesh = nullptr;
        atlas_rt_bvh* tlas = nullptr;
        atlas_rt_scene* scene = nullptr;
        ~Query() {
            if (scene) atlas_rt_scene_free(scene);
            if (tlas) atlas_rt_bvh_free(tlas);
            if (mesh) atlas_rt_mesh_free(mesh);
            if (blas) atlas_rt_bvh_free(blas);
        }
    };
    std::shared_ptr<Query> query;

    bool EnsureQuery(atlas_rt_context* ctx) {
        if (query) return query->scene != nullptr;
        if (data.empty()) return false;
        query = std::make_shared<Query>();
        std::vector<uint32_t> order(data.size());
        std::vector<uint8_t> flags(data.size());
        std::vector<float> tris(data.size() * 9);
        float box[6] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
        for (size_t i = 0; i < data.size(); i++) {   // slot i of the device tree is data[i] itself
            const BVHTriangle& t = data[i];
            order[i] = uint32_t(i);
            flags[i] = t.endOfNode ? 1 : 0;
            const float v[9] = {t.v0.x, t.v0.y, t.v0.z, t.v1.x, t.v1.y, t.v1.z, t.v2.x, t.v2.y, t.v2.z};
            std::memcpy(&tris[9 * i], v, sizeof(v));
            for (int k = 0; k < 9; k++) { box[k % 3] = v[k] < box[k % 3] ? v[k] : box[k % 3]; box[3 + k % 3] = v[k] > box[3 + k % 3] ? v[k] : box[3 + k % 3]; }
        }
        GPUBVHInstance inst;
        inst.inverseMatrix[0] = vec4(1, 0, 0, 0); inst.inverseMatrix[1] = vec4(0, 1, 0, 0); inst.inverseMatrix[2] = vec4(0, 0, 1, 0);
        inst.mask = MaskAll | MaskShadow;
        using RayTracing::detail::Check;
        const bool ok = Check(atlas_rt_bvh_upload(ctx, nodes.data(), nodes.size(), order.data(), flags.data(), data.size(), &query->blas)) &&
                        Check(atlas_rt_pack_mesh(ctx, query->blas, tris.data(), data.size(), nullptr, nullptr, 0, &query->mesh)) &&
                        Check(atlas_rt_build_tlas(ctx, box, 1, 0, &query->tlas)) &&
                        Check(atlas_rt_scene_create(ctx, &query->mesh, 1, &inst, 1, query->tlas, 0, &query->scene));
        if (!ok) { query = std::make_shared<Query>(); return false; }
        return true;
    }

    bool TraceOne(const Ray& ray, bool any, PackedHit& hit) {
        atlas_rt_context* ctx = RayTracing::detail::Context();
        if (!ctx || !EnsureQuery(ctx)) return false;
        PackedRay in, out;
        in.origin = vec4(ray.origin.x, ray.origin.y, ray.origin.z, 0.0f);   // ID 0 (>= 0: a live ray)
        in.direction = vec4(ray.direction.x, ray.direction.y, ray.direction.z, 0.0f);
        const int32_t none = -1;
        std::memcpy(&in.hit.y, &none, 4);
        const int rc = any ? atlas_rt_trace_any(ctx, query->scene, &in, 1, MaskAll, ray.tMin, ray.tMax, &out, 0)
                           : atlas_rt_trace_closest(ctx, query->scene, &in, 1, MaskAll, ray.tMin, ray.tMax, &out, 0);
        if (!RayTracing::detail::Check(rc)) return false;
        hit.t = out.hit.x; hit.u = out.direction.w; hit.v = out.hit.w;
        std::memcpy(&hit.id, &out.hit.y, 4);
        return true;
    }

    // Ref::nodeIdx as Flatten leaves it (BVH.cpp:413): the index of the node pushed LAST before the leaf was reached in
    // the pre-order walk — the parent for a left leaf, the last node of the left subtree for a right leaf.
    void FillNodeIdx() {
        if (nodes.empty() || refs.empty()) return;
        if (refs.size() == 2 && nodes.size() == 1 && nodes[0].leftPtr == ~0 && nodes[0].rightPtr == ~0) return;   // count == 1 quirk: both stay 0
        std::vector<int32_t> todo;   // pending right children, as pointers
        uint32_t pushed = 0;
        int32_t ptr = 0;
        for (;;) {
            if (ptr >= 0) {
                pushed = uint32_t(ptr) + 1;   // pre-order: node ptr is the pushed-th node
                todo.push_back(nodes[size_t(ptr)].rightPtr);
                ptr = nodes[size_t(ptr)].leftPtr;
                continue;
            }
            for (size_t slot = size_t(~ptr); slot < refs.size(); slot++) {
                refs[slot].nodeIdx = pushed - 1;
                if (refs[slot].endOfNode) break;
            }
            if (todo.empty()) break;
            ptr = todo.back();
            todo.pop_back();
        }
    }

    bool Fetch(atlas_rt_bvh* h, std::vector<uint32_t>& order, std::vector<uint8_t>& flags) {
        uint64_t n = 0, m = 0;
        atlas_rt_bvh_counts(h, &n, &m);
        nodes.resize(n);
        order.resize(m);
        flags.resize(m);
        return RayTracing::detail::Check(atlas_rt_bvh_download(h, nodes.data(), order.data(), flags.data(), 0));
    }
};

}   // namespace Volume

namespace RayTracing {

// The software-RT half of Mesh::MeshData::BuildBVH (mesh/MeshData.cpp:89-269): triangles from an indexed vertex
// buffer, per-triangle boxes, BLAS build and the packed GPUBVHTriangle / GPUBVHNode arrays. The device-resident BLAS
// and mesh are kept (handles) so the scene can be traced without re-uploading; gpuBvhTriangles / gpuBvhNodes are the
// host copies Mesh::BuildBVH would upload (mesh/Mesh.cpp:97-108).
struct MeshBVH {
    std::vector<GPUTriangle> gpuTriangles;        // filled by PackShadingTriangles
    std::vector<GPUBVHTriangle> gpuBvhTriangles;
    std::vector<GPUBVHNode> gpuBvhNodes;
    atlas_rt_bvh* blas = nullptr;
    atlas_rt_mesh* mesh = nullptr;
    MeshBVH() = default;
    MeshBVH(const MeshBVH&) = delete;              // owns device memory: movable, not copyable
    MeshBVH& operator=(const MeshBVH&) = delete;
    MeshBVH(MeshBVH&& o) noexcept { *this = std::move(o); }
    MeshBVH& operator=(MeshBVH&& o) noexcept {
        if (this != &o) {
            Release();
            gpuTriangles = std::move(o.gpuTriangles); gpuBvhTriangles = std::move(o.gpuBvhTriangles); gpuBvhNodes = std::move(o.gpuBvhNodes);
            blas = o.blas; mesh = o.mesh;
            o.blas = nullptr; o.mesh = nullptr;
        }
        return *this;
    }
    ~MeshBVH() { Release(); }
    bool IsBVHBuilt() const { return gpuBvhTriangles.size() > 0; }   // MeshData.cpp:273-277
    // May be called from any thread: the handles keep the context they were created on alive.
    void Release() {
        if (mesh) atlas_rt_mesh_free(mesh);
        if (blas) atlas_rt_bvh_free(blas);
        mesh = nullptr;
        blas = nullptr;
    }
};

inline bool BuildMeshBVH(const std::vector<vec3>& vertices, const std::vector<uint32_t>& indices, int32_t materialIdx, float opacity,
                         MeshBVH& out, bool keepHostCopies = true) {
    atlas_rt_context* ctx = detail::Context();
    if (!ctx) return false;
    const size_t n = indices.size() / 3;
    std::vector<float> tris(n * 9), boxes(n * 6);
    for (size_t k = 0; k < n; k++) {   // MeshData.cpp:102-159
        const vec3 v[3] = {vertices[indices[3 * k]], vertices[indices[3 * k + 1]], vertices[indices[3 * k + 2]]};
        for (int c = 0; c < 3; c++) {
            tris[9 * k + c] = v[0][c]; tris[9 * k + 3 + c] = v[1][c]; tris[9 * k + 6 + c] = v[2][c];
            float lo = v[0][c], hi = v[0][c];   // glm::min(glm::min(v0, v1), v2): (y < x) ? y : x
            lo = (v[1][c] < lo) ? v[1][c] : lo; lo = (v[2][c] < lo) ? v[2][c] : lo;
            hi = (hi < v[1][c]) ? v[1][c] : hi; hi = (hi < v[2][c]) ? v[2][c] : hi;
            boxes[6 * k + c] = lo; boxes[6 * k + 3 + c] = hi;
        }
    }
    out.Release();
    if (!detail::Check(atlas_rt_build_blas(ctx, boxes.data(), tris.data(), n, 0, &out.blas))) return false;
    std::vector<int32_t> mats(n, materialIdx);
    std::vector<float> ops(n, opacity);
    if (!detail::Check(atlas_rt_pack_mesh(ctx, out.blas, tris.data(), n, mats.data(), ops.data(), 0, &out.mesh))) return false;
    if (keepHostCopies) {
        uint64_t nodes = 0, refs = 0;
        atlas_rt_mesh_counts(out.mesh, &nodes, &refs);
        out.gpuBvhNodes.resize(nodes);
        out.gpuBvhTriangles.resize(refs);
        if (!detail::Check(atlas_rt_mesh_download(out.mesh, out.gpuBvhNodes.data(), out.gpuBvhTriangles.data(), 0))) return false;
    }
    return true;
}

// The GPUTriangle half of MeshData::BuildBVH (mesh/MeshData.cpp:176-239): `packed` holds, per SOURCE triangle, the 11
// words the engine already computes there (pn0, pn1, pn2, puv0, puv1, puv2, pt, pbt, pc0, pc1, pc2); they are gathered
// into flattened order next to the vertices, material index, endOfNode flag and opacity. Needed by the opacity-aware
// traversal (ATLAS_RT_OPACITY) and by the engine's hit shaders.
inline bool PackShadingTriangles(MeshBVH& mesh, const std::vector<vec3>& vertices, const std::vector<uint32_t>& indices,
                                 const std::vector<int32_t>& materialIdx, const std::vector<float>& opacity,
                                 const std::vector<uint32_t>& packed, bool keepHostCopy = true) {
    atlas_rt_context* ctx = detail::Context();
    if (!ctx || !mesh.mesh) return false;
    const size_t n = indices.size() / 3;
    if (materialIdx.size() != n || opacity.size() != n || (!packed.empty() && packed.size() != 11 * n)) return false;
    std::vector<float> tris(n * 9);
    for (size_t k = 0; k < n; k++)
        for (int v = 0; v < 3; v++)
            for (int c = 0; c < 3; c++) tris[9 * k + 3 * v + c] = vertices[indices[3 * k + v]][c];
    if (!detail::Check(atlas_rt_mesh_pack_shading(ctx, mesh.mesh, tris.data(), n, materialIdx.data(), opacity.data(),
                                                  packed.empty() ? nullptr : packed.data(), 0))) return false;
    if (keepHostCopy) {
        uint64_t nodes = 0, refs = 0;
        atlas_rt_mesh_counts(mesh.mesh, &nodes, &refs);
        mesh.gpuTriangles.resize(refs);
        return detail::Check(atlas_rt_mesh_download_shading(mesh.mesh, mesh.gpuTriangles.data(), 0));
    }
    return true;
}

// RayTracingWorld::UpdateForSoftwareRayTracing (RayTracingWorld.cpp:267-307): TLAS over the actors' world boxes,
// instances permuted into TLAS order with nextInstance set, TLAS nodes in the GPU layout. `gpuBvhInstances` is replaced
// by the ordered array exactly like the reference does (:303). The device scene is returned for tracing.
struct World {
    std::vector<GPUBVHNode> tlasNodes;
    atlas_rt_bvh* tlas = nullptr;
    atlas_rt_scene* scene = nullptr;
    World() = default;
    World(const World&) = delete;
    World& operator=(const World&) = delete;
    ~World() { Release(); }
    void Release() {
        if (scene) atlas_rt_scene_free(scene);
        if (tlas) atlas_rt_bvh_free(tlas);
        scene = nullptr;
        tlas = nullptr;
    }
};

inline bool UpdateForSoftwareRayTracing(std::vector<GPUBVHInstance>& gpuBvhInstances, const std::vector<Volume::AABB>& actorAABBs,
                                        const std::vector<const MeshBVH*>& meshes, World& world) {
    atlas_rt_context* ctx = detail::Context();
    if (!ctx || gpuBvhInstances.size() != actorAABBs.size() || meshes.empty()) return false;
    world.Release();
    if (!detail::Check(atlas_rt_build_tlas(ctx, reinterpret_cast<const float*>(actorAABBs.data()), actorAABBs.size(), 0, &world.tlas))) return false;
    std::vector<const atlas_rt_mesh*> handles(meshes.size());
    for (size_t i = 0; i < meshes.size(); i++) handles[i] = meshes[i]->mesh;
    if (!detail::Check(atlas_rt_scene_create(ctx, handles.data(), uint32_t(handles.size()), gpuBvhInstances.data(), gpuBvhInstances.size(),
                                             world.tlas, 0, &world.scene))) return false;
    uint64_t nodes = 0, refs = 0;
    atlas_rt_bvh_counts(world.tlas, &nodes, &refs);
    world.tlasNodes.resize(nodes);
    gpuBvhInstances.resize(refs);
    return detail::Check(atlas_rt_scene_download(world.scene, gpuBvhInstances.data(), world.tlasNodes.data(), 0));
}

// The trace-batch boundary: what RayTracingHelper::DispatchHitClosest + traceClosest.csh do to the ray buffer
// (renderer/helper/RayTracingHelper.cpp:346-364): every ray gets its closest hit written back.
class Tracer {
public:
    static bool HitClosest(const World& world, const std::vector<PackedRay>& rays, std::vector<PackedRay>& out,
                           uint32_t cullMask = MaskAll, float tMin = 0.0f, float tMax = ATLAS_RT_INF) {
        out.resize(rays.size());
        return detail::Check(atlas_rt_trace_closest(detail::Context(), world.scene, rays.data(), rays.size(), cullMask, tMin, tMax, out.data(), 0));
    }
    static bool HitAny(const World& world, const std::vector<PackedRay>& rays, std::vector<PackedRay>& out,
                       uint32_t cullMask = MaskShadow, float tMin = 0.0f, float tMax = ATLAS_RT_INF, bool perRayTMax = false) {
        out.resize(rays.size());
        return detail::Check(atlas_rt_trace_any(detail::Context(), world.scene, rays.data(), rays.size(), cullMask, tMin, tMax, out.data(),
                                                perRayTMax ? ATLAS_RT_PER_RAY_TMAX : 0));
    }
};

}   // namespace RayTracing
}   // namespace Atlas
