// Drives the C++ drop-in classes (atlas_engine_b200/host/AtlasRT.h) the way the engine drives the originals and
// compares with the reference through its C bridge (oracle/_ref/libatlas_ref.so, dlopen'ed — test infrastructure).
// Also builds several meshes from concurrent threads, as the engine's job system does (src/tests/App.cpp:362-370).
#include <dlfcn.h>

#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <random>
#include <thread>
#include <vector>

#include "../../atlas_engine_b200/host/AtlasRT.h"

using namespace Atlas;

typedef void* (*ref_build_blas_t)(const float*, const float*, uint64_t, int);
typedef void* (*ref_build_tlas_t)(const float*, uint64_t, int);
typedef uint64_t (*ref_count_t)(void*);
typedef void (*ref_copy_nodes_t)(void*, void*);
typedef void (*ref_copy_order_t)(void*, uint32_t*, uint8_t*);
typedef void (*ref_free_t)(void*);
typedef void (*ref_copy_node_idx_t)(void*, uint32_t*);
typedef void (*ref_intersect_closest_t)(void*, const float*, uint64_t, float*, int32_t*, int);
typedef void (*ref_intersect_any_t)(void*, const float*, uint64_t, uint8_t*, int);
typedef int (*ref_void_t)();

static std::vector<Volume::BVHTriangle> soup(size_t n, unsigned seed, float extent) {
    std::mt19937 rng(seed);
    std::uniform_real_distribution<float> u(0.0f, 1.0f);
    std::vector<Volume::BVHTriangle> t(n);
    for (size_t i = 0; i < n; i++) {
        const vec3 c(u(rng), u(rng), u(rng));
        vec3* v[3] = {&t[i].v0, &t[i].v1, &t[i].v2};
        for (auto* p : v) *p = vec3(c.x + (u(rng) - 0.5f) * extent, c.y + (u(rng) - 0.5f) * extent, c.z + (u(rng) - 0.5f) * extent);
        t[i].idx = uint32_t(i);
    }
    return t;
}

static std::vector<Volume::AABB> boxes_of(const std::vector<Volume::BVHTriangle>& t) {
    std::vector<Volume::AABB> b(t.size());
    for (size_t i = 0; i < t.size(); i++) {
        for (int c = 0; c < 3; c++) {
            b[i].min[c] = std::min(t[i].v0[c], std::min(t[i].v1[c], t[i].v2[c]));
            b[i].max[c] = std::max(t[i].v0[c], std::max(t[i].v1[c], t[i].v2[c]));
        }
    }
    return b;
}

int main(int argc, char** argv) {
    const char* refPath = argc > 1 ? argv[1] : "oracle/_ref/libatlas_ref.so";
    void* lib = dlopen(refPath, RTLD_NOW);
    if (!lib) { printf("SKIP: cannot load %s\n", refPath); return 77; }
    auto ref_build_blas = (ref_build_blas_t)dlsym(lib, "ref_build_blas");
    auto ref_build_tlas = (ref_build_tlas_t)dlsym(lib, "ref_build_tlas");
    auto ref_nodes = (ref_count_t)dlsym(lib, "ref_bvh_node_count");
    auto ref_refs = (ref_count_t)dlsym(lib, "ref_bvh_ref_count");
    auto ref_copy_nodes = (ref_copy_nodes_t)dlsym(lib, "ref_bvh_copy_nodes");
    auto ref_copy_order = (ref_copy_order_t)dlsym(lib, "ref_bvh_copy_order");
    auto ref_free = (ref_free_t)dlsym(lib, "ref_bvh_free");
    auto ref_shutdown = (ref_void_t)dlsym(lib, "ref_shutdown");
    auto ref_copy_node_idx = (ref_copy_node_idx_t)dlsym(lib, "ref_bvh_copy_node_idx");
    auto ref_intersect_closest = (ref_intersect_closest_t)dlsym(lib, "ref_bvh_intersect_closest");
    auto ref_intersect_any = (ref_intersect_any_t)dlsym(lib, "ref_bvh_intersect_any");
    int failures = 0;

    // ---- BLAS through Volume::BVH(aabbs, data), several meshes concurrently
    const size_t sizes[4] = {1000, 20000, 7, 50000};
    const float extents[4] = {0.05f, 0.01f, 0.5f, 0.3f};
    std::vector<Volume::BVH> built(4);
    std::vector<std::vector<Volume::BVHTriangle>> tris(4);
    std::vector<std::vector<Volume::AABB>> boxes(4);
    for (int k = 0; k < 4; k++) { tris[k] = soup(sizes[k], 100 + k, extents[k]); boxes[k] = boxes_of(tris[k]); }
    std::vector<std::thread> workers;
    for (int k = 0; k < 4; k++) workers.emplace_back([&, k] { built[k] = Volume::BVH(boxes[k], tris[k], true); });
    for (auto& w : workers) w.join();
    for (int k = 0; k < 4; k++) {
        std::vector<float> flat(sizes[k] * 9);
        for (size_t i = 0; i < sizes[k]; i++) {
            const float v[9] = {tris[k][i].v0.x, tris[k][i].v0.y, tris[k][i].v0.z, tris[k][i].v1.x, tris[k][i].v1.y, tris[k][i].v1.z,
                                tris[k][i].v2.x, tris[k][i].v2.y, tris[k][i].v2.z};
            memcpy(&flat[9 * i], v, sizeof(v));
        }
        void* r = ref_build_blas(reinterpret_cast<const float*>(boxes[k].data()), flat.data(), sizes[k], 1);
        std::vector<Volume::BVHNode> rn(ref_nodes(r));
        std::vector<uint32_t> ro(ref_refs(r));
        std::vector<uint8_t> rf(ref_refs(r));
        ref_copy_nodes(r, rn.data());
        ref_copy_order(r, ro.data(), rf.data());
        ref_free(r);
        bool ok = rn.size() == built[k].nodes.size() && ro.size() == built[k].data.size() &&
                  (rn.empty() || memcmp(rn.data(), built[k].nodes.data(), rn.size() * sizeof(Volume::BVHNode)) == 0);
        for (size_t i = 0; ok && i < ro.size(); i++)
            ok = built[k].data[i].idx == ro[i] && built[k].data[i].endOfNode == (rf[i] != 0) &&
                 memcmp(&built[k].aabbs[i], &boxes[k][ro[i]], sizeof(Volume::AABB)) == 0;
        ok = ok && built[k].refs.empty();
        printf("BLAS %zu tris: nodes %zu refs %zu %s\n", sizes[k], built[k].nodes.size(), built[k].data.size(), ok ? "== reference" : "MISMATCH");
        failures += ok ? 0 : 1;
    }
    // size mismatch => silently empty (BVH.cpp:18-19)
    {
        std::vector<Volume::AABB> fewer(boxes[0].begin(), boxes[0].begin() + 10);
        Volume::BVH bad(fewer, tris[0]);
        const bool ok = bad.nodes.empty() && bad.data.empty();
        printf("size mismatch: %s\n", ok ? "empty BVH" : "MISMATCH");
        failures += ok ? 0 : 1;
    }
    // ---- TLAS through Volume::BVH(aabbs)
    for (size_t m : {size_t(1), size_t(2), size_t(3), size_t(300), size_t(5000)}) {
        std::vector<Volume::AABB> ib(boxes[3].begin(), boxes[3].begin() + m);
        Volume::BVH tl(ib);
        void* r = ref_build_tlas(reinterpret_cast<const float*>(ib.data()), m, 1);
        std::vector<Volume::BVHNode> rn(ref_nodes(r));
        std::vector<uint32_t> ro(ref_refs(r));
        std::vector<uint8_t> rf(ref_refs(r));
        ref_copy_nodes(r, rn.data());
        ref_copy_order(r, ro.data(), rf.data());
        ref_free(r);
        bool ok = rn.size() == tl.nodes.size() && ro.size() == tl.refs.size() && memcmp(rn.data(), tl.nodes.data(), rn.size() * sizeof(Volume::BVHNode)) == 0;
        for (size_t i = 0; ok && i < ro.size(); i++) ok = tl.refs[i].idx == ro[i] && tl.refs[i].endOfNode == (rf[i] != 0);
        if (ok && ref_copy_node_idx) {   // Ref::nodeIdx, BVH.cpp:413
            std::vector<uint32_t> ni(ro.size());
            void* r2 = ref_build_tlas(reinterpret_cast<const float*>(ib.data()), m, 1);
            ref_copy_node_idx(r2, ni.data());
            ref_free(r2);
            for (size_t i = 0; ok && i < ni.size(); i++) ok = tl.refs[i].nodeIdx == ni[i];
        }
        printf("TLAS %zu instances: nodes %zu refs %zu %s\n", m, tl.nodes.size(), tl.refs.size(), ok ? "== reference" : "MISMATCH");
        failures += ok ? 0 : 1;
    }
    // ---- MeshData::BuildBVH + UpdateForSoftwareRayTracing + trace, end to end through the mirrors
    {
        std::vector<vec3> verts;
        std::vector<uint32_t> idx;
        const int g = 40;
        for (int z = 0; z <= g; z++) for (int x = 0; x <= g; x++) verts.push_back(vec3(float(x), 2.0f * sinf(0.3f * x) * cosf(0.2f * z), float(z)));
        for (int z = 0; z < g; z++) for (int x = 0; x < g; x++) {
            const uint32_t a = z * (g + 1) + x, b = a + 1, c = a + g + 1, d = c + 1;
            for (uint32_t v : {a, b, d, a, d, c}) idx.push_back(v);
        }
        RayTracing::MeshBVH mesh;
        bool ok = RayTracing::BuildMeshBVH(verts, idx, 3, 1.0f, mesh) && mesh.IsBVHBuilt() && mesh.gpuBvhNodes.size() + 1 == mesh.gpuBvhTriangles.size();
        std::vector<GPUBVHInstance> inst(2);
        std::vector<Volume::AABB> actor(2);
        for (int k = 0; k < 2; k++) {
            const float dx = 100.0f * k;   // instance k is the mesh translated by (dx, 0, 0): inverse translates back
            inst[k].inverseMatrix[0] = vec4(1, 0, 0, -dx); inst[k].inverseMatrix[1] = vec4(0, 1, 0, 0); inst[k].inverseMatrix[2] = vec4(0, 0, 1, 0);
            inst[k].meshOffset = 0; inst[k].mask = MaskAll | MaskShadow;
            actor[k] = Volume::AABB(vec3(dx, -2.0f, 0.0f), vec3(dx + g, 2.0f, float(g)));
        }
        RayTracing::World world;
        ok = ok && RayTracing::UpdateForSoftwareRayTracing(inst, actor, {&mesh}, world) && world.tlasNodes.size() == 1 && inst.size() == 2;
        std::vector<PackedRay> rays(2), out;
        for (int k = 0; k < 2; k++) {
            int id = k;
            rays[k].origin = vec4(100.0f * k + 20.3f, 50.0f, 20.7f, 0.0f);
            memcpy(&rays[k].origin.w, &id, 4);
            rays[k].direction = vec4(0.001f, -1.0f, 0.002f, 0.0f);
        }
        ok = ok && RayTracing::Tracer::HitClosest(world, rays, out);
        int hitInst[2] = {-1, -1};
        for (int k = 0; ok && k < 2; k++) {
            int hitID;
            memcpy(&hitID, &out[k].hit.y, 4);
            memcpy(&hitInst[k], &out[k].hit.z, 4);
            ok = hitID >= 0 && out[k].hit.x > 45.0f && out[k].hit.x < 55.0f;
        }
        ok = ok && hitInst[0] != hitInst[1] && out[0].hit.x == out[1].hit.x;   // same mesh, same local ray => same t
        printf("mesh + world + trace through the mirrors: %s (%s)\n", ok ? "ok" : "MISMATCH", RayTracing::LastError().c_str());
        failures += ok ? 0 : 1;
        world.Release();
        mesh.Release();
    }
    // ---- BVH::GetIntersection / GetIntersectionAny on the drop-in class (volume/BVH.h:124-127) against the reference's
    if (ref_intersect_closest && ref_intersect_any) {
        const int k = 1;   // the 20000-triangle soup
        std::vector<float> flat(sizes[k] * 9);
        for (size_t i = 0; i < sizes[k]; i++) {
            const float v[9] = {tris[k][i].v0.x, tris[k][i].v0.y, tris[k][i].v0.z, tris[k][i].v1.x, tris[k][i].v1.y, tris[k][i].v1.z,
                                tris[k][i].v2.x, tris[k][i].v2.y, tris[k][i].v2.z};
            memcpy(&flat[9 * i], v, sizeof(v));
        }
        void* r = ref_build_blas(reinterpret_cast<const float*>(boxes[k].data()), flat.data(), sizes[k], 1);
        std::mt19937 rng(4711);
        std::uniform_real_distribution<float> u(0.0f, 1.0f);
        const int nq = 400;
        std::vector<float> q(8 * nq), tuv(3 * nq);
        std::vector<int32_t> idx(nq);
        std::vector<uint8_t> anyRef(nq);
        for (int i = 0; i < nq; i++) {
            float d[3] = {u(rng) - 0.5f, u(rng) - 0.5f, u(rng) - 0.5f};
            const float l = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) + 1e-6f;
            const float v[8] = {u(rng), u(rng), u(rng), d[0] / l, d[1] / l, d[2] / l, 0.0f, i % 3 ? 1.0e12f : 0.15f};
            memcpy(&q[8 * i], v, sizeof(v));
        }
        ref_intersect_closest(r, q.data(), nq, tuv.data(), idx.data(), 1);
        ref_intersect_any(r, q.data(), nq, anyRef.data(), 1);
        ref_free(r);
        std::vector<std::pair<int32_t, float>> stack(256);
        bool ok = true;
        int hits = 0;
        Volume::BVH copy = built[k];   // copies share the lazily created device tree
        for (int i = 0; ok && i < nq; i++) {
            Volume::Ray ray(vec3(q[8 * i], q[8 * i + 1], q[8 * i + 2]), vec3(q[8 * i + 3], q[8 * i + 4], q[8 * i + 5]), q[8 * i + 6], q[8 * i + 7]);
            Volume::BVHTriangle closest;
            closest.idx = 0xffffffffu;
            vec3 sol;
            const bool hit = (i & 1 ? copy : built[k]).GetIntersection(stack, ray, closest, sol);
            ok = hit == (idx[i] >= 0) && sol.x == tuv[3 * i] && (!hit || (int32_t(closest.idx) == idx[i] && sol.y == tuv[3 * i + 1] && sol.z == tuv[3 * i + 2]));
            const bool anyHit = built[k].GetIntersectionAny(stack, ray);
            ok = ok && anyHit == (anyRef[i] != 0);
            if (!ok) printf("  ray %d: hit %d idx %d (ref %d) sol %.9g %.9g %.9g ref %.9g %.9g %.9g any %d (ref %d) tMax %g err '%s'\n", i, int(hit), int(closest.idx), idx[i],
                            sol.x, sol.y, sol.z, tuv[3 * i], tuv[3 * i + 1], tuv[3 * i + 2], int(anyHit), int(anyRef[i]), q[8 * i + 7], RayTracing::LastError().c_str());
            hits += hit ? 1 : 0;
        }
        ok = ok && hits > 5;
        printf("BVH::GetIntersection / GetIntersectionAny on %d rays (%d hits): %s\n", nq, hits, ok ? "== reference" : "MISMATCH");
        failures += ok ? 0 : 1;
    }
    // ---- meshes built on job-system-like worker threads that EXIT, scene assembled and traced on this thread, meshes
    // released afterwards (MeshData::BuildBVH on workers, RayTracingWorld::UpdateForSoftwareRayTracing elsewhere)
    {
        const int nm = 6;
        std::vector<RayTracing::MeshBVH> meshes(nm);
        std::vector<std::thread> pool;
        std::vector<int> built_ok(nm, 0);
        for (int k = 0; k < nm; k++) pool.emplace_back([&, k] {
            std::vector<vec3> verts;
            std::vector<uint32_t> idx;
            const int g = 10 + 7 * k;
            for (int z = 0; z <= g; z++) for (int x = 0; x <= g; x++) verts.push_back(vec3(float(x) * 10.0f / g, 0.5f * sinf(0.9f * x + k) * cosf(0.7f * z), float(z) * 10.0f / g));
            for (int z = 0; z < g; z++) for (int x = 0; x < g; x++) {
                const uint32_t a = z * (g + 1) + x, b = a + 1, c = a + g + 1, d = c + 1;
                for (uint32_t v : {a, b, d, a, d, c}) idx.push_back(v);
            }
            RayTracing::MeshBVH local;
            built_ok[k] = RayTracing::BuildMeshBVH(verts, idx, k, 1.0f, local, false) ? 1 : 0;
            meshes[k] = std::move(local);
        });
        for (auto& w : pool) w.join();   // the workers' thread-local contexts are gone now; their meshes must live on
        bool ok = true;
        for (int k = 0; k < nm; k++) ok = ok && built_ok[k] && meshes[k].mesh;
        std::vector<GPUBVHInstance> inst(nm);
        std::vector<Volume::AABB> actor(nm);
        std::vector<const RayTracing::MeshBVH*> ptrs(nm);
        for (int k = 0; k < nm; k++) {
            const float dx = 20.0f * k;
            inst[k].inverseMatrix[0] = vec4(1, 0, 0, -dx); inst[k].inverseMatrix[1] = vec4(0, 1, 0, 0); inst[k].inverseMatrix[2] = vec4(0, 0, 1, 0);
            inst[k].meshOffset = k; inst[k].mask = MaskAll | MaskShadow;
            actor[k] = Volume::AABB(vec3(dx, -0.5f, 0.0f), vec3(dx + 10.0f, 0.5f, 10.0f));
            ptrs[k] = &meshes[k];
        }
        RayTracing::World world;
        ok = ok && RayTracing::UpdateForSoftwareRayTracing(inst, actor, ptrs, world);
        std::vector<PackedRay> rays(nm), out;
        for (int k = 0; k < nm; k++) {
            rays[k].origin = vec4(20.0f * k + 5.1f, 30.0f, 4.9f, 0.0f);
            memcpy(&rays[k].origin.w, &k, 4);
            rays[k].direction = vec4(0.001f, -1.0f, 0.002f, 0.0f);
        }
        ok = ok && RayTracing::Tracer::HitClosest(world, rays, out);
        for (int k = 0; ok && k < nm; k++) {
            int hitID, hitInst;
            memcpy(&hitID, &out[k].hit.y, 4);
            memcpy(&hitInst, &out[k].hit.z, 4);
            ok = hitID >= 0 && hitInst >= 0 && hitInst < nm && inst[hitInst].meshOffset == k && out[k].hit.x > 29.0f && out[k].hit.x < 31.0f;
        }
        std::thread releaser([&] { world.Release(); for (auto& m : meshes) m.Release(); });   // freed from yet another thread
        releaser.join();
        printf("meshes from exited worker threads + scene on the main thread: %s (%s)\n", ok ? "ok" : "MISMATCH", RayTracing::LastError().c_str());
        failures += ok ? 0 : 1;
    }
    if (ref_shutdown) ref_shutdown();
    printf(failures ? "FAILED %d\n" : "ALL OK\n", failures);
    return failures ? 1 : 0;
}
