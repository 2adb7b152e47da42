// C bridge over the UNMODIFIED reference classes (Atlas::Volume::BVH, /root/reference/src/engine/volume/BVH.h:114-138)
// so that tests and bench.py's cpu_baseline can drive the real reference builder / CPU traversal through ctypes.
// TEST INFRASTRUCTURE ONLY — built into oracle/_ref/libatlas_ref.so by oracle/Makefile from the sources where they
// lie under /root/reference; never linked into or called by the product library.
#include "volume/BVH.h"
#include "jobsystem/JobSystem.h"
#include "common/Packing.h"

#include <cstring>
#include <thread>
#include <vector>
#include <atomic>

using namespace Atlas;

static std::atomic<int> g_initialised{0};

extern "C" {

// JobSystem::Init with the engine defaults (jobsystem/JobSystem.h:17-21): hw-1 / hw-3 / hw-4 workers.
int ref_init() {
    if (g_initialised.exchange(1)) return 0;
    JobSystem::Init(JobSystemConfig{});
    return 0;
}

void ref_shutdown() {
    if (!g_initialised.exchange(0)) return;
    JobSystem::Shutdown();
}

int ref_hardware_concurrency() { return int(std::thread::hardware_concurrency()); }

struct RefBVH {
    Volume::BVH bvh;
    bool tlas = false;
};

static void fill_aabbs(std::vector<Volume::AABB>& out, const float* aabbs, uint64_t n) {
    out.resize(n);
    for (uint64_t i = 0; i < n; i++) {
        out[i].min = glm::vec3(aabbs[6 * i + 0], aabbs[6 * i + 1], aabbs[6 * i + 2]);
        out[i].max = glm::vec3(aabbs[6 * i + 3], aabbs[6 * i + 4], aabbs[6 * i + 5]);
    }
}

// BVH(aabbs, data, parallelBuild) — volume/BVH.cpp:14-56. tris = n x 9 floats (v0,v1,v2); idx = position.
void* ref_build_blas(const float* aabbs, const float* tris, uint64_t n, int parallel) {
    ref_init();
    std::vector<Volume::AABB> boxes;
    fill_aabbs(boxes, aabbs, n);
    std::vector<Volume::BVHTriangle> data(n);
    for (uint64_t i = 0; i < n; i++) {
        const float* t = tris + 9 * i;
        data[i].v0 = glm::vec3(t[0], t[1], t[2]);
        data[i].v1 = glm::vec3(t[3], t[4], t[5]);
        data[i].v2 = glm::vec3(t[6], t[7], t[8]);
        data[i].idx = uint32_t(i);
    }
    auto* r = new RefBVH;
    r->bvh = Volume::BVH(boxes, data, parallel != 0);
    return r;
}

// BVH(aabbs, parallelBuild) — volume/BVH.cpp:58-101.
void* ref_build_tlas(const float* aabbs, uint64_t n, int parallel) {
    ref_init();
    std::vector<Volume::AABB> boxes;
    fill_aabbs(boxes, aabbs, n);
    auto* r = new RefBVH;
    r->bvh = Volume::BVH(boxes, parallel != 0);
    r->tlas = true;
    return r;
}

uint64_t ref_bvh_node_count(void* h) { return static_cast<RefBVH*>(h)->bvh.nodes.size(); }

uint64_t ref_bvh_ref_count(void* h) {
    auto* r = static_cast<RefBVH*>(h);
    return r->tlas ? r->bvh.refs.size() : r->bvh.data.size();
}

// nodes as 14 32-bit words each: leftAABB.min, leftAABB.max, rightAABB.min, rightAABB.max, leftPtr, rightPtr.
void ref_bvh_copy_nodes(void* h, void* out) {
    auto* r = static_cast<RefBVH*>(h);
    static_assert(sizeof(Volume::BVHNode) == 56, "BVHNode layout");
    std::memcpy(out, r->bvh.nodes.data(), r->bvh.nodes.size() * sizeof(Volume::BVHNode));
}

// order[i] = source index of the primitive at flattened slot i; flags[i] = endOfNode.
void ref_bvh_copy_order(void* h, uint32_t* order, uint8_t* flags) {
    auto* r = static_cast<RefBVH*>(h);
    if (r->tlas) {
        for (size_t i = 0; i < r->bvh.refs.size(); i++) { order[i] = r->bvh.refs[i].idx; flags[i] = r->bvh.refs[i].endOfNode; }
    } else {
        for (size_t i = 0; i < r->bvh.data.size(); i++) { order[i] = r->bvh.data[i].idx; flags[i] = r->bvh.data[i].endOfNode; }
    }
}

// Ref::nodeIdx of a TLAS as Flatten left it (BVH.cpp:413).
void ref_bvh_copy_node_idx(void* h, uint32_t* out) {
    auto* r = static_cast<RefBVH*>(h);
    for (size_t i = 0; i < r->bvh.refs.size(); i++) out[i] = r->bvh.refs[i].nodeIdx;
}

// aabbs member (unclipped source boxes in flattened order), 6 floats each.
void ref_bvh_copy_aabbs(void* h, float* out) {
    auto* r = static_cast<RefBVH*>(h);
    std::memcpy(out, r->bvh.aabbs.data(), r->bvh.aabbs.size() * sizeof(Volume::AABB));
}

void ref_bvh_free(void* h) { delete static_cast<RefBVH*>(h); }

// BVH::GetIntersection (volume/BVH.cpp:103-174) over a BLAS. rays = n x 8 floats (origin, direction, tMin, tMax).
// out_tuv = n x 3 (intersection.x = t or tMax when nothing was hit), out_slot = slot in data[] of the closest
// triangle or -1. The function's bool return value is unreliable and is ignored.
void ref_bvh_intersect_closest(void* h, const float* rays, uint64_t n, float* out_tuv, int32_t* out_idx, int nthreads) {
    auto* r = static_cast<RefBVH*>(h);
    if (nthreads < 1) nthreads = 1;
    auto work = [&](uint64_t b, uint64_t e) {
        std::vector<std::pair<int32_t, float>> stack(256);
        for (uint64_t i = b; i < e; i++) {
            const float* q = rays + 8 * i;
            Volume::Ray ray(glm::vec3(q[0], q[1], q[2]), glm::vec3(q[3], q[4], q[5]), q[6], q[7]);
            Volume::BVHTriangle closest;
            closest.idx = 0xffffffffu;
            glm::vec3 sol;
            r->bvh.GetIntersection(stack, ray, closest, sol);
            bool hit = sol.x < q[7];
            out_tuv[3 * i + 0] = sol.x; out_tuv[3 * i + 1] = hit ? sol.y : 0.0f; out_tuv[3 * i + 2] = hit ? sol.z : 0.0f;
            out_idx[i] = hit ? int32_t(closest.idx) : -1;
        }
    };
    std::vector<std::thread> pool;
    uint64_t per = (n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; t++) {
        uint64_t b = per * t, e = std::min<uint64_t>(n, b + per);
        if (b < e) pool.emplace_back(work, b, e);
    }
    for (auto& t : pool) t.join();
}

// BVH::GetIntersectionAny (volume/BVH.cpp:176-217).
void ref_bvh_intersect_any(void* h, const float* rays, uint64_t n, uint8_t* out_hit, int nthreads) {
    auto* r = static_cast<RefBVH*>(h);
    if (nthreads < 1) nthreads = 1;
    auto work = [&](uint64_t b, uint64_t e) {
        std::vector<std::pair<int32_t, float>> stack(256);
        for (uint64_t i = b; i < e; i++) {
            const float* q = rays + 8 * i;
            Volume::Ray ray(glm::vec3(q[0], q[1], q[2]), glm::vec3(q[3], q[4], q[5]), q[6], q[7]);
            out_hit[i] = r->bvh.GetIntersectionAny(stack, ray) ? 1 : 0;
        }
    };
    std::vector<std::thread> pool;
    uint64_t per = (n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; t++) {
        uint64_t b = per * t, e = std::min<uint64_t>(n, b + per);
        if (b < e) pool.emplace_back(work, b, e);
    }
    for (auto& t : pool) t.join();
}

// Common::Packing::PackSignedVector3x10_1x2 (common/Packing.cpp:24-35), the reference's own code: n vec4s in, n words out.
// (The packed normals / tangents / bitangents of GPUTriangle, mesh/MeshData.cpp:205-210, are made by exactly this function.)
void ref_pack_signed_3x10_1x2(const float* vec4s, uint64_t n, int32_t* out) {
    for (uint64_t i = 0; i < n; i++)
        out[i] = Common::Packing::PackSignedVector3x10_1x2(glm::vec4(vec4s[4 * i], vec4s[4 * i + 1], vec4s[4 * i + 2], vec4s[4 * i + 3]));
}

}
