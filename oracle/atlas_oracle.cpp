// atlas_oracle.cpp — CPU restatement of the Atlas-Engine ray-tracing acceleration path.
//
// TEST INFRASTRUCTURE ONLY. Nothing in the product library (atlas_engine_b200/csrc, include/) may include, link or
// call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load the
// shared object built from it (oracle/Makefile -> oracle/libatlas_oracle.so).
//
// Parity status: PINNED. The builder half is checked bit-for-bit against the unmodified reference compiled into
// oracle/_ref/libatlas_ref.so (tests/test_oracle_vs_ref.py, fixtures in tests/golden/ produced by
// tests/golden/make_golden.py). The traversal half restates GLSL that cannot be executed here (no Vulkan); it is
// pinned indirectly: it must agree with the reference's own CPU traversal BVH::GetIntersection on the hit triangle
// and distance, and with a brute-force two-level intersector (tests/test_oracle_traversal.py).
//
// What is restated (reference file:line, paths relative to /root/reference):
//   builder     src/engine/volume/BVH.cpp:14-101 (constructors), :249-406 (Build x2), :408-441 (Flatten),
//               :444-556 (object split), :558-798 (spatial split), :800-855 (median split)
//   box maths   src/engine/volume/AABB.cpp:88-115
//   traversal   data/shader/raytracer/bvh.hsh:21-37,44-104,172-273,359-441; intersections.hsh:3-58;
//               common.hsh:46-73; traceClosest.csh:12-36
//   third party glm 0.9.8.0 (vcpkg.json:45-48; not vendored): min(x,y)=(y<x)?y:x, max(x,y)=(x<y)?y:x,
//               clamp=min(max(x,lo),hi), mix(x,y,a)=x+a*(y-x); std::sort = this toolchain's libstdc++.
//
// Arithmetic contract: plain IEEE fp32, round-to-nearest, NO fused multiply-add (build with -ffp-contract=off).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

constexpr float kBig = std::numeric_limits<float>::max();

// ---- glm 0.9.8 scalar forms -------------------------------------------------------------------------------------
inline float gmin(float x, float y) { return (y < x) ? y : x; }
inline float gmax(float x, float y) { return (x < y) ? y : x; }
inline float gclamp(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }

struct F3 {
    float v[3];
    float& operator[](int i) { return v[i]; }
    float operator[](int i) const { return v[i]; }
};

// Box with the AABB.cpp operations. "empty" is the builder's InitialAABB (BVH.cpp:863-869).
struct Box {
    F3 lo, hi;
    static Box empty() { return Box{{{kBig, kBig, kBig}}, {{-kBig, -kBig, -kBig}}}; }
    // AABB::Grow(AABB) — AABB.cpp:88-93: max first, accumulated value is the first argument.
    void grow(const Box& o) {
        for (int a = 0; a < 3; a++) hi[a] = gmax(hi[a], o.hi[a]);
        for (int a = 0; a < 3; a++) lo[a] = gmin(lo[a], o.lo[a]);
    }
    // AABB::Grow(vec3) — AABB.cpp:95-100: the point is the FIRST argument here.
    void grow(const F3& p) {
        for (int a = 0; a < 3; a++) hi[a] = gmax(p[a], hi[a]);
        for (int a = 0; a < 3; a++) lo[a] = gmin(p[a], lo[a]);
    }
    // AABB::Intersect — AABB.cpp:102-107.
    void clip(const Box& o) {
        for (int a = 0; a < 3; a++) lo[a] = gmax(lo[a], o.lo[a]);
        for (int a = 0; a < 3; a++) hi[a] = gmin(hi[a], o.hi[a]);
    }
    // AABB::GetSurfaceArea — AABB.cpp:109-115. No clamping: inverted boxes give whatever the formula gives.
    float area() const {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return 2.0f * (dx * dy + dy * dz + dz * dx);
    }
};

struct PrimRef {   // BVHBuilder::Ref (BVH.h:38-43) without the fields Flatten fills in
    uint32_t src;
    Box box;
};

struct SplitPlan {   // BVHBuilder::Split (BVH.h:80-90)
    float cost = kBig;
    int axis = -1;
    uint32_t bin = 0;
    float pos = 0.0f;
    Box left = Box::empty();
    Box right = Box::empty();
};

struct Stats {
    uint64_t medianSplits = 0, sortFallbacks = 0, sortFallbackMaxN = 0, spatialTried = 0, spatialChosen = 0,
             axisSkipped = 0, maxDepth = 0, sumLeafDepth = 0, duplicates = 0, lastResortLeaf = 0;
};

// One builder node (what a BVHBuilder object holds after Build()).
struct BNode {
    Box box;
    uint32_t depth = 0;
    int kid[2] = {-1, -1};
    std::vector<PrimRef> leaf;   // non-empty <=> CreateLeaf was called (BVH.cpp:857-861)
};

struct Builder {
    uint32_t binBudget;            // 256 for a BLAS, 64 for a TLAS (BVH.cpp:34,72)
    const float* tris = nullptr;   // 9 floats per source triangle (BLAS only)
    std::vector<BNode> pool;
    Stats st;

    uint32_t bins_at(uint32_t depth) const { return std::max(binBudget / (depth + 1), 16u); }   // BVH.cpp:447

    // Bin index of a coordinate: uint32_t(clamp((value - start) * inv, 0, bins-1)) — BVH.cpp:477, :545, :590-593.
    static uint32_t bin_of(float value, float start, float inv, uint32_t bins) {
        return uint32_t(gclamp((value - start) * inv, 0.0f, float(bins) - 1.0f));
    }

    // Shared sweep of BVH.cpp:484-522 / :621-661. For the object split count[j] plays both roles
    // (enter == exit == primitiveCount); the spatial split passes separate enter/exit counters.
    static void sweep(int axis, float start, float width, const std::vector<Box>& binBox,
                      const std::vector<uint32_t>& enter, const std::vector<uint32_t>& exit, uint32_t total,
                      SplitPlan& best) {
        const size_t nb = binBox.size();
        std::vector<Box> suffix(nb, Box::empty());
        Box acc = Box::empty();
        for (size_t j = nb - 1; j > 0; j--) {
            acc.grow(binBox[j]);
            suffix[j - 1] = acc;
        }
        Box pre = Box::empty();
        uint32_t nLeft = 0, nRight = total;
        for (size_t j = 1; j < nb; j++) {
            pre.grow(binBox[j - 1]);
            nLeft += enter[j - 1];
            nRight -= exit[j - 1];
            if (!nLeft || !nRight) continue;
            const float cost = pre.area() * float(nLeft) + suffix[j - 1].area() * float(nRight);
            if (cost < best.cost) {
                best.cost = cost;
                best.axis = axis;
                best.bin = uint32_t(j);
                best.pos = start + float(j) * width;
                best.left = pre;
                best.right = suffix[j - 1];
            }
        }
    }

    // FindObjectSplit — BVH.cpp:444-528.
    SplitPlan find_object_split(const BNode& n, const std::vector<PrimRef>& refs) {
        SplitPlan best;
        const uint32_t nb = bins_at(n.depth);
        std::vector<Box> binBox(nb);
        std::vector<uint32_t> count(nb);
        for (int axis = 0; axis < 3; axis++) {
            const float start = n.box.lo[axis], stop = n.box.hi[axis];
            if (std::fabs(stop - start) < 1e-3f) { st.axisSkipped++; continue; }
            std::fill(binBox.begin(), binBox.end(), Box::empty());
            std::fill(count.begin(), count.end(), 0u);
            const float width = (stop - start) / float(nb);
            const float inv = 1.0f / width;
            for (const PrimRef& r : refs) {
                const float centre = 0.5f * (r.box.lo[axis] + r.box.hi[axis]);
                const uint32_t b = bin_of(centre, start, inv, nb);
                count[b]++;
                binBox[b].grow(r.box);
            }
            // nRight in the reference is refs.size() - nLeft (BVH.cpp:501): same as total - sum(exit) here.
            sweep(axis, start, width, binBox, count, count, uint32_t(refs.size()), best);
        }
        return best;
    }

    // SplitReference — BVH.cpp:760-798. tri = 9 floats. Outputs start from the empty box.
    static void split_reference(const float* tri, const Box& current, Box& outL, Box& outR, float plane, int axis) {
        outL = Box::empty();
        outR = Box::empty();
        for (int e = 0; e < 3; e++) {
            F3 a{{tri[3 * e], tri[3 * e + 1], tri[3 * e + 2]}};
            const int e1 = (e + 1) % 3;
            F3 b{{tri[3 * e1], tri[3 * e1 + 1], tri[3 * e1 + 2]}};
            if ((a[axis] < plane && b[axis] > plane) || (a[axis] > plane && b[axis] < plane)) {
                const float off = gclamp((plane - a[axis]) / (b[axis] - a[axis]), 0.0f, 1.0f);
                F3 p;   // glm 0.9.8 mix: x + a * (y - x)
                for (int c = 0; c < 3; c++) p[c] = a[c] + off * (b[c] - a[c]);
                outL.grow(p);
                outR.grow(p);
            }
            if (a[axis] <= plane) outL.grow(a);
            if (a[axis] >= plane) outR.grow(a);
        }
        outL.hi[axis] = plane;
        outR.lo[axis] = plane;
        outL.clip(current);
        outR.clip(current);
    }

    // FindSpatialSplit — BVH.cpp:558-667.
    SplitPlan find_spatial_split(const BNode& n, const std::vector<PrimRef>& refs) {
        SplitPlan best;
        const uint32_t nb = bins_at(n.depth);
        std::vector<Box> binBox(nb);
        std::vector<uint32_t> enter(nb), exit(nb);
        for (int axis = 0; axis < 3; axis++) {
            const float start = n.box.lo[axis], stop = n.box.hi[axis];
            if (std::fabs(stop - start) < 1e-3f) continue;
            std::fill(binBox.begin(), binBox.end(), Box::empty());
            std::fill(enter.begin(), enter.end(), 0u);
            std::fill(exit.begin(), exit.end(), 0u);
            const float width = (stop - start) / float(nb);
            const float inv = 1.0f / width;
            for (const PrimRef& r : refs) {
                const uint32_t b0 = bin_of(r.box.lo[axis], start, inv, nb);
                const uint32_t b1 = bin_of(r.box.hi[axis], start, inv, nb);
                if (b0 == b1) {
                    enter[b0]++;
                    exit[b0]++;
                    binBox[b0].grow(r.box);
                    continue;
                }
                Box rest = r.box;
                for (uint32_t j = b0; j < b1; j++) {
                    Box l, rr;
                    split_reference(tris + 9 * size_t(r.src), rest, l, rr, start + float(j + 1) * width, axis);
                    binBox[j].grow(l);
                    rest = rr;
                }
                binBox[b1].grow(rest);
                enter[b0]++;
                exit[b1]++;
            }
            sweep(axis, start, width, binBox, enter, exit, uint32_t(refs.size()), best);
        }
        return best;
    }

    // PerformObjectSplit — BVH.cpp:530-556 (stable on both sides).
    void do_object_split(const BNode& n, const std::vector<PrimRef>& refs, const SplitPlan& s,
                         std::vector<PrimRef>& L, std::vector<PrimRef>& R) {
        const uint32_t nb = bins_at(n.depth);
        const float start = n.box.lo[s.axis], stop = n.box.hi[s.axis];
        const float width = (stop - start) / float(nb);
        const float inv = 1.0f / width;
        for (const PrimRef& r : refs) {
            const float centre = 0.5f * (r.box.lo[s.axis] + r.box.hi[s.axis]);
            (bin_of(centre, start, inv, nb) < s.bin ? L : R).push_back(r);
        }
    }

    // PerformSpatialSplit — BVH.cpp:669-758. Mutates s.left / s.right.
    void do_spatial_split(const BNode& n, const std::vector<PrimRef>& refs, SplitPlan& s,
                          std::vector<PrimRef>& L, std::vector<PrimRef>& R) {
        const uint32_t nb = bins_at(n.depth);
        const float start = n.box.lo[s.axis], stop = n.box.hi[s.axis];
        const float width = (stop - start) / float(nb);
        const float inv = 1.0f / width;
        for (const PrimRef& r : refs) {
            const uint32_t b0 = bin_of(r.box.lo[s.axis], start, inv, nb);
            const uint32_t b1 = bin_of(r.box.hi[s.axis], start, inv, nb);
            if (b1 < s.bin) { L.push_back(r); s.left.grow(r.box); }
            else if (b0 >= s.bin) { R.push_back(r); s.right.grow(r.box); }
        }
        for (const PrimRef& r : refs) {
            const uint32_t b0 = bin_of(r.box.lo[s.axis], start, inv, nb);
            const uint32_t b1 = bin_of(r.box.hi[s.axis], start, inv, nb);
            if (!(b1 >= s.bin && b0 < s.bin)) continue;
            Box cl, cr;
            split_reference(tris + 9 * size_t(r.src), r.box, cl, cr, s.pos, s.axis);
            Box unsplitL = s.left, unsplitR = s.right, dupL = s.left, dupR = s.right;
            unsplitL.grow(r.box);
            unsplitR.grow(r.box);
            dupL.grow(cl);
            dupR.grow(cr);
            const float nl = float(L.size()), nr = float(R.size());
            const float nl1 = float(L.size() + 1), nr1 = float(R.size() + 1);
            const float sahUnsplitL = unsplitL.area() * nl1 + s.right.area() * nr;
            const float sahUnsplitR = s.left.area() * nl + unsplitR.area() * nr1;
            const float sahDup = dupL.area() * nl1 + dupR.area() * nr1;
            const float best = gmin(sahDup, gmin(sahUnsplitR, sahUnsplitL));
            if (best == sahUnsplitL) { s.left = unsplitL; L.push_back(r); }
            else if (best == sahUnsplitR) { s.right = unsplitR; R.push_back(r); }
            else {
                s.left = dupL;
                s.right = dupR;
                L.push_back(PrimRef{r.src, cl});
                R.push_back(PrimRef{r.src, cr});
                st.duplicates++;
            }
        }
    }

    // PerformMedianSplit — BVH.cpp:800-855. May reorder `refs` (std::sort in the fallback).
    SplitPlan do_median_split(const BNode& n, std::vector<PrimRef>& refs, std::vector<PrimRef>& L, std::vector<PrimRef>& R) {
        st.medianSplits++;
        SplitPlan s;
        const float dim[3] = {n.box.hi[0] - n.box.lo[0], n.box.hi[1] - n.box.lo[1], n.box.hi[2] - n.box.lo[2]};
        int axis = 0;
        if (dim[1] > dim[axis]) axis = 1;
        if (dim[2] > dim[axis]) axis = 2;
        const float cut = n.box.lo[axis] + dim[axis] / 2.0f;
        for (const PrimRef& r : refs) {
            const float centre = (r.box.hi[axis] - r.box.lo[axis]) * 0.5f + r.box.lo[axis];
            if (centre < cut) { s.left.grow(r.box); L.push_back(r); }
            else { s.right.grow(r.box); R.push_back(r); }
        }
        if (L.empty() || R.empty()) {
            st.sortFallbacks++;
            st.sortFallbackMaxN = std::max<uint64_t>(st.sortFallbackMaxN, refs.size());
            s = SplitPlan();
            L.clear();
            R.clear();
            // Keyed on EXTENT along the axis (the reference calls it "center" but computes max - min).
            std::sort(refs.begin(), refs.end(), [axis](const PrimRef& p, const PrimRef& q) {
                return (p.box.hi[axis] - p.box.lo[axis]) < (q.box.hi[axis] - q.box.lo[axis]);
            });
            const uint32_t half = uint32_t(refs.size() / 2);
            for (uint32_t i = 0; i < half; i++) { s.left.grow(refs[i].box); L.push_back(refs[i]); }
            for (uint32_t i = half; i < uint32_t(refs.size()); i++) { s.right.grow(refs[i].box); R.push_back(refs[i]); }
        }
        return s;
    }

    int new_node(const Box& box, uint32_t depth) {
        pool.emplace_back();
        pool.back().box = box;
        pool.back().depth = depth;
        return int(pool.size()) - 1;
    }

    // Second Build overload (BVH.cpp:343-406): every node below a BLAS root, and every TLAS node. Iterative
    // (explicit work list) because tree depth is unbounded; child creation order does not affect the result.
    void build_plain(int rootId, std::vector<PrimRef>&& rootRefs) {
        struct Item { int id; std::vector<PrimRef> refs; };
        std::vector<Item> work;
        work.push_back(Item{rootId, std::move(rootRefs)});
        while (!work.empty()) {
            Item it = std::move(work.back());
            work.pop_back();
            const uint32_t depth = pool[it.id].depth;
            st.maxDepth = std::max<uint64_t>(st.maxDepth, depth);
            if (it.refs.size() == 1) {
                pool[it.id].leaf = it.refs;
                st.sumLeafDepth += depth;
                continue;
            }
            const float nodeCost = float(it.refs.size()) * pool[it.id].box.area();   // BVH.cpp:238
            SplitPlan obj = find_object_split(pool[it.id], it.refs);
            std::vector<PrimRef> L, R;
            SplitPlan used;
            if (obj.axis < 0 || obj.cost >= nodeCost) used = do_median_split(pool[it.id], it.refs, L, R);
            else { do_object_split(pool[it.id], it.refs, obj, L, R); used = obj; }
            it.refs.clear();
            it.refs.shrink_to_fit();
            if (!L.empty()) { int c = new_node(used.left, depth + 1); pool[it.id].kid[0] = c; work.push_back(Item{c, std::move(L)}); }
            if (!R.empty()) { int c = new_node(used.right, depth + 1); pool[it.id].kid[1] = c; work.push_back(Item{c, std::move(R)}); }
        }
    }

    // First Build overload at depth 0 (BVH.cpp:249-341) — the only place it is ever entered (children use the second
    // overload, :314-337), so the depth>0 leaf tests (:253, :284-287) can never fire and are not restated.
    void build_blas_root(int rootId, std::vector<PrimRef>&& refs, float minOverlap) {
        BNode& rootRef = pool[rootId];
        const float nodeCost = float(refs.size()) * rootRef.box.area();   // BVH.cpp:230
        SplitPlan obj = find_object_split(pool[rootId], refs);
        SplitPlan spa;
        {   // depth (0) <= 16 always holds here — BVH.cpp:262-268
            Box overlap = obj.left;
            overlap.clip(obj.right);
            if (overlap.area() >= minOverlap) {
                st.spatialTried++;
                spa = find_spatial_split(pool[rootId], refs);
            }
        }
        std::vector<PrimRef> L, R;
        SplitPlan used;
        if ((obj.axis < 0 || obj.cost >= nodeCost) && (spa.axis < 0 || spa.cost >= nodeCost)) {
            used = do_median_split(pool[rootId], refs, L, R);
        } else {
            if (spa.cost < obj.cost) {
                st.spatialChosen++;
                do_spatial_split(pool[rootId], refs, spa, L, R);
                used = spa;
            }
            if (obj.cost <= spa.cost) {
                L.clear();
                R.clear();
                do_object_split(pool[rootId], refs, obj, L, R);
                used = obj;
            }
        }
        if (L.empty() || R.empty()) {   // "last resort" — BVH.cpp:303-305; no return, children still get built
            pool[rootId].leaf = refs;
            st.lastResortLeaf++;
        }
        if (!L.empty()) { int c = new_node(used.left, 1); pool[rootId].kid[0] = c; build_plain(c, std::move(L)); }
        if (!R.empty()) { int c = new_node(used.right, 1); pool[rootId].kid[1] = c; build_plain(c, std::move(R)); }
    }
};

struct FlatNode {   // Volume::BVHNode, 56 bytes (BVH.h:14-24)
    Box left, right;
    int32_t leftPtr = 0, rightPtr = 0;
};
static_assert(sizeof(FlatNode) == 56, "BVHNode layout");

struct FlatTree {
    std::vector<FlatNode> nodes;
    std::vector<uint32_t> order;
    std::vector<uint8_t> endOfNode;
    Stats st;
};

// Flatten — BVH.cpp:408-441 (DFS pre-order, larger-area child first). Explicit stack instead of recursion.
void flatten(Builder& b, int rootId, FlatTree& out) {
    struct Frame { int id; int stage; size_t nodeIdx; };
    std::vector<Frame> stack;
    stack.push_back(Frame{rootId, 0, 0});
    while (!stack.empty()) {
        Frame& f = stack.back();
        BNode& n = b.pool[f.id];
        if (f.stage == 0) {
            if (!n.leaf.empty()) {
                for (const PrimRef& r : n.leaf) { out.order.push_back(r.src); out.endOfNode.push_back(0); }
                out.endOfNode.back() = 1;
                stack.pop_back();
                continue;
            }
            f.nodeIdx = out.nodes.size();
            out.nodes.emplace_back();
            if (b.pool[n.kid[0]].box.area() < b.pool[n.kid[1]].box.area()) std::swap(n.kid[0], n.kid[1]);
            f.stage = 1;
            const int c = n.kid[0];
            const bool leaf = !b.pool[c].leaf.empty();
            out.nodes[f.nodeIdx].left = b.pool[c].box;
            out.nodes[f.nodeIdx].leftPtr = leaf ? ~int32_t(out.order.size()) : int32_t(out.nodes.size());
            stack.push_back(Frame{c, 0, 0});
        } else if (f.stage == 1) {
            f.stage = 2;
            const int c = n.kid[1];
            const bool leaf = !b.pool[c].leaf.empty();
            const size_t ni = f.nodeIdx;
            out.nodes[ni].right = b.pool[c].box;
            out.nodes[ni].rightPtr = leaf ? ~int32_t(out.order.size()) : int32_t(out.nodes.size());
            stack.push_back(Frame{c, 0, 0});
        } else {
            stack.pop_back();
        }
    }
}

Box union_of(const float* aabbs, uint64_t n) {
    Box all = Box::empty();
    for (uint64_t i = 0; i < n; i++) {
        Box b{{{aabbs[6 * i], aabbs[6 * i + 1], aabbs[6 * i + 2]}}, {{aabbs[6 * i + 3], aabbs[6 * i + 4], aabbs[6 * i + 5]}}};
        all.grow(b);
    }
    return all;
}

std::vector<PrimRef> initial_refs(const float* aabbs, uint64_t n) {
    std::vector<PrimRef> refs(n);
    for (uint64_t i = 0; i < n; i++) {
        refs[i].src = uint32_t(i);
        refs[i].box = Box{{{aabbs[6 * i], aabbs[6 * i + 1], aabbs[6 * i + 2]}}, {{aabbs[6 * i + 3], aabbs[6 * i + 4], aabbs[6 * i + 5]}}};
    }
    return refs;
}

// ================================================= traversal =====================================================
// GLSL min/max: "returns y if y < x, otherwise x" / "returns y if x < y, otherwise x" (NaN case pinned this way).
inline float smin(float x, float y) { return (y < x) ? y : x; }
inline float smax(float x, float y) { return (x < y) ? y : x; }

struct SceneView {
    const float* tlasNodes;          // 16 floats per node (GPUBVHNode, RTStructures.h:95-103)
    const float* instances;          // 16 words per instance (GPUBVHInstance, RTStructures.h:85-93)
    const float* const* blasNodes;   // per mesh: 16 floats per node
    const float* const* bvhTris;     // per mesh: 12 floats per triangle (GPUBVHTriangle, RTStructures.h:23-27)
    const float* const* triangles;   // per mesh: 24 floats per triangle (GPUTriangle, RTStructures.h:14-21); only for the *Transparency variants
    // material / texture tables for GetOpacity (textured opacity); null => textured triangles count as opacity 1
    const uint32_t* materials = nullptr;     // RaytraceMaterial, 23 words each
    uint32_t materialCount = 0;
    const uint32_t* texDims = nullptr;       // width, height per texture
    const uint8_t* const* texels = nullptr;  // R8
    uint32_t textureCount = 0;
};

extern "C" float oracle_get_opacity(const float* tri, float s, float t, const uint32_t* material, const uint32_t* texDims,
                                    const uint8_t* const* texels, uint32_t textureCount);   // atlas_oracle_shade.cpp

struct Counters {
    uint64_t tlasNodes = 0, instances = 0, blasNodes = 0, triangles = 0, maxStack = 0, stackOverflows = 0;
};

struct RayState {
    float o[3], d[3];
    float hitDistance;
    int32_t hitID, hitInstanceID, currentInstanceID;
    float baryU, baryV;
    float transparency;
};

// IntersectAABB with distance — intersections.hsh:19-34 (divides by the direction, no reciprocal).
inline bool slab(const RayState& r, const float* lo, const float* hi, float tmin, float tmax, float& dist) {
    float ts[3], tb[3];
    for (int a = 0; a < 3; a++) {
        const float t0 = (lo[a] - r.o[a]) / r.d[a];
        const float t1 = (hi[a] - r.o[a]) / r.d[a];
        ts[a] = smin(t0, t1);
        tb[a] = smax(t0, t1);
    }
    const float tminf = smax(smax(tmin, ts[0]), smax(ts[1], ts[2]));
    const float tmaxf = smin(smin(tmax, tb[0]), smin(tb[1], tb[2]));
    const bool hit = tminf <= tmaxf;
    dist = hit ? tminf : tmax;
    return hit;
}

// IntersectTriangle — intersections.hsh:36-58.
inline bool tri_test(const RayState& r, const float* v0, const float* v1, const float* v2, float sol[3]) {
    const float e0[3] = {v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2]};
    const float e1[3] = {v2[0] - v0[0], v2[1] - v0[1], v2[2] - v0[2]};
    const float s[3] = {r.o[0] - v0[0], r.o[1] - v0[1], r.o[2] - v0[2]};
    const float p[3] = {s[1] * e0[2] - e0[1] * s[2], s[2] * e0[0] - e0[2] * s[0], s[0] * e0[1] - e0[0] * s[1]};
    const float q[3] = {r.d[1] * e1[2] - e1[1] * r.d[2], r.d[2] * e1[0] - e1[2] * r.d[0], r.d[0] * e1[1] - e1[0] * r.d[1]};
    const float den = (q[0] * e0[0] + q[1] * e0[1]) + q[2] * e0[2];
    sol[0] = ((p[0] * e1[0] + p[1] * e1[1]) + p[2] * e1[2]) / den;
    sol[1] = ((q[0] * s[0] + q[1] * s[1]) + q[2] * s[2]) / den;
    sol[2] = ((p[0] * r.d[0] + p[1] * r.d[1]) + p[2] * r.d[2]) / den;
    return sol[0] >= 0.0f && sol[1] >= 0.0f && sol[2] >= 0.0f && sol[1] + sol[2] <= 1.0f;
}

constexpr uint32_t kStackLimit = 32;                 // STACK_SIZE, bvh.hsh:16
constexpr uint32_t kTlasInvalid = kStackLimit + 2;   // TLAS_INVALID, bvh.hsh:17
constexpr uint32_t kOracleStack = 4096;              // the oracle never overflows; it reports depth > 32 instead

// HitClosest (bvh.hsh:191-273) when ANY == false, HitAny (bvh.hsh:359-441) when ANY == true. With OPACITY the
// *Transparency variants (bvh.hsh:275-357, :443-524) with CheckLeafClosestTransparency (:106-135) / CheckLeafTransparency
// (:137-170): they read the 96-byte triangles[] array and consult tri.opacity. Textured opacity (tri.opacity < 0 ->
// GetOpacity, surface.hsh:147-160) needs the material/texture tables, which are outside this path: it counts as 1.0.
template <bool ANY, bool OPACITY = false>
bool traverse(const SceneView& sc, RayState& ray, uint32_t cullMask, float tMin, float tMax, Counters& ct) {
    if (std::isnan(ray.d[0]) || std::isnan(ray.d[1]) || std::isnan(ray.d[2])) {
        if (!ANY) ray.hitDistance = tMax;
        return false;
    }
    int32_t stack[kOracleStack];
    stack[0] = 0;
    uint32_t sp = 1;
    int32_t nodePtr = 0, meshPtr = 0, materialOffset = 0;
    const float o0[3] = {ray.o[0], ray.o[1], ray.o[2]}, d0[3] = {ray.d[0], ray.d[1], ray.d[2]};
    if (!ANY) ray.hitDistance = tMax;
    uint32_t tlasIndex = kTlasInvalid;
    bool hit = false;
    uint64_t localMax = 1;
    // In the closest-hit loop the slab interval is [tMin, ray.hitDistance]; in the any-hit loop [tMin, tMax].
    if (ANY && OPACITY) ray.transparency = 1.0f;
    while (sp != 0u && !(ANY && !OPACITY && hit) && !(ANY && OPACITY && !(ray.transparency > 0.0f))) {
        const bool inTlas = sp < tlasIndex;
        if (inTlas) {
            // HitClosest and both *Transparency variants restore unconditionally (bvh.hsh:218-220, :302-304, :469-471);
            // plain HitAny only when it has just left a BLAS (:387-390). That difference is observable: CheckInstance
            // transforms the ray BEFORE the mask test, so after an instance culled by the mask plain HitAny walks on
            // through the TLAS with the instance-space ray (and a second culled instance transforms it again) until the
            // next BLAS exit restores it. Reproduced as is — it is what the reference computes.
            if (!ANY || OPACITY || tlasIndex != kTlasInvalid)
                for (int a = 0; a < 3; a++) { ray.o[a] = o0[a]; ray.d[a] = d0[a]; }
            tlasIndex = kTlasInvalid;
        }
        if (inTlas && nodePtr < 0) {
            // CheckInstance — bvh.hsh:172-189.
            const int32_t inst = ~nodePtr;
            const float* I = sc.instances + 16 * size_t(inst);
            ct.instances++;
            float no[3], nd[3];
            for (int c = 0; c < 3; c++) {
                const float* col = I + 4 * c;   // vec4(v, w) * mat3x4 -> dot(vec4, column c)
                no[c] = ((ray.o[0] * col[0] + ray.o[1] * col[1]) + ray.o[2] * col[2]) + 1.0f * col[3];
                nd[c] = ((ray.d[0] * col[0] + ray.d[1] * col[1]) + ray.d[2] * col[2]) + 0.0f * col[3];
            }
            for (int a = 0; a < 3; a++) { ray.o[a] = no[a]; ray.d[a] = nd[a]; }
            ray.currentInstanceID = inst;
            int32_t meshOffset, mask;
            std::memcpy(&meshOffset, I + 12, 4);
            std::memcpy(&mask, I + 15, 4);
            meshPtr = meshOffset;
            std::memcpy(&materialOffset, I + 13, 4);
            nodePtr = 0;
            if ((uint32_t(mask) & cullMask) > 0u) tlasIndex = sp;
            else nodePtr = stack[--sp];
            continue;
        }
        if (!inTlas && nodePtr < 0) {
            // CheckLeafClosest (bvh.hsh:44-72) / CheckLeaf (:74-104) / the two *Transparency leaf loops.
            int32_t triPtr = ~nodePtr;
            bool end = false;
            const float tmaxLeaf = ANY ? tMax : ray.hitDistance;   // captured at call time (by-value parameter)
            float leafTransparency = ray.transparency;              // CheckLeafTransparency's by-value parameter
            while (!end && !(ANY && !OPACITY && hit)) {
                const float* T = OPACITY ? sc.triangles[meshPtr] + 24 * size_t(triPtr) : sc.bvhTris[meshPtr] + 12 * size_t(triPtr);
                end = OPACITY ? (T[18] > 0.0f) : (T[3] > 0.0f);     // d1.z / v0.w
                float sol[3];
                ct.triangles++;
                const bool in = tri_test(ray, T, T + 4, T + 8, sol);
                if (in && sol[0] > tMin && sol[0] < tmaxLeaf) {
                    if (OPACITY) {
                        // tri.opacity < 0.0 ? GetOpacity(tri, sol.yz, materialOffset, 0) : tri.opacity — bvh.hsh:127, :163
                        float triOpacity = T[23];   // d2.w
                        if (triOpacity < 0.0f) {
                            int32_t matIndex;
                            std::memcpy(&matIndex, T + 15, 4);
                            const uint32_t mi = uint32_t(matIndex + materialOffset);
                            triOpacity = (sc.materials && mi < sc.materialCount)
                                             ? oracle_get_opacity(T, sol[1], sol[2], sc.materials + 23 * size_t(mi), sc.texDims, sc.texels, sc.textureCount)
                                             : 1.0f;
                        }
                        if (!ANY) {
                            if (sol[0] < ray.hitDistance && triOpacity > 0.0f) {
                                ray.hitDistance = sol[0]; ray.hitID = triPtr; ray.hitInstanceID = ray.currentInstanceID;
                                ray.baryU = sol[1]; ray.baryV = sol[2];
                            }
                        } else {
                            ray.hitDistance = sol[0]; ray.hitID = triPtr; ray.hitInstanceID = ray.currentInstanceID;
                            leafTransparency *= (1.0f - triOpacity);
                        }
                    } else if (ANY || sol[0] < ray.hitDistance) {
                        ray.hitDistance = sol[0];
                        ray.hitID = triPtr;
                        ray.hitInstanceID = ray.currentInstanceID;
                        ray.baryU = sol[1];
                        ray.baryV = sol[2];
                        if (ANY) hit = true;
                    }
                }
                triPtr++;
            }
            if (ANY && OPACITY) {   // bvh.hsh:489-492: transparency *= CheckLeafTransparency(..., transparency)
                ray.transparency *= leafTransparency;
                if (ray.transparency < 0.000001f) ray.transparency = 0.0f;
            }
            nodePtr = stack[--sp];
            continue;
        }
        // Inner node of the TLAS or of BLAS meshPtr — UnpackNode, bvh.hsh:21-37.
        const float* N = (inTlas ? sc.tlasNodes : sc.blasNodes[meshPtr]) + 16 * size_t(nodePtr);
        if (inTlas) ct.tlasNodes++; else ct.blasNodes++;
        int32_t leftPtr, rightPtr;
        std::memcpy(&leftPtr, N + 12, 4);
        std::memcpy(&rightPtr, N + 13, 4);
        float hitL = 0.0f, hitR = 0.0f;
        const float tfar = ANY ? tMax : ray.hitDistance;
        const bool iL = slab(ray, N + 0, N + 3, tMin, tfar, hitL);
        const bool iR = slab(ray, N + 6, N + 9, tMin, tfar, hitR);
        if (!ANY) {
            nodePtr = hitL <= hitR ? leftPtr : rightPtr;
            if (!iL && !iR) nodePtr = stack[--sp];
            if (iL && iR) stack[sp++] = hitL <= hitR ? rightPtr : leftPtr;
        } else {
            nodePtr = iL ? leftPtr : rightPtr;
            if (!iL && !iR) nodePtr = stack[--sp];
            if (iL && iR) stack[sp++] = rightPtr;
        }
        if (sp > localMax) localMax = sp;
        if (sp >= kOracleStack - 1) break;   // never reached on sane input
    }
    for (int a = 0; a < 3; a++) { ray.o[a] = o0[a]; ray.d[a] = d0[a]; }
    if (localMax > ct.maxStack) ct.maxStack = localMax;
    if (localMax > kStackLimit) ct.stackOverflows++;
    return hit;
}

template <typename F>
void parallel_for(uint64_t n, int nthreads, F&& body) {
    if (nthreads <= 1 || n < 2) { body(0, n, 0); return; }
    std::vector<std::thread> pool;
    const uint64_t per = (n + uint64_t(nthreads) - 1) / uint64_t(nthreads);
    for (int t = 0; t < nthreads; t++) {
        const uint64_t b = per * uint64_t(t), e = std::min<uint64_t>(n, b + per);
        if (b < e) pool.emplace_back([=, &body] { body(b, e, t); });
    }
    for (auto& th : pool) th.join();
}

}   // namespace

// ==================================================== C API ======================================================
extern "C" {

struct OracleTree { FlatTree t; };

// BLAS: BVH(aabbs, data, parallelBuild) — BVH.cpp:14-56. aabbs n x 6, tris n x 9.
void* oracle_build_blas(const float* aabbs, const float* tris, uint64_t n) {
    auto* out = new OracleTree;
    Builder b;
    b.binBudget = 256;
    b.tris = tris;
    b.pool.reserve(2 * n + 2);
    const Box all = union_of(aabbs, n);
    const float minOverlap = all.area() * 10e-6f;
    const int root = b.new_node(all, 0);
    if (n == 0) { out->t.st = b.st; return out; }   // the reference dereferences a null child here; we return empty
    b.build_blas_root(root, initial_refs(aabbs, n), minOverlap);
    flatten(b, root, out->t);
    // Flatten on a leaf root leaves `nodes` empty (BVH.cpp:411-417 with nodes.size()==0).
    out->t.st = b.st;
    return out;
}

// TLAS: BVH(aabbs, parallelBuild) — BVH.cpp:58-101.
void* oracle_build_tlas(const float* aabbs, uint64_t n) {
    auto* out = new OracleTree;
    Builder b;
    b.binBudget = 64;
    b.pool.reserve(2 * n + 2);
    const Box all = union_of(aabbs, n);
    const int root = b.new_node(all, 0);
    if (n == 0) { out->t.st = b.st; return out; }
    std::vector<PrimRef> refs = initial_refs(aabbs, n);
    if (n == 1) {
        // BVH.cpp:346-349 returns before refs.clear(), so the member vector still holds the source ref when
        // Flatten appends the leaf copy (:415): two entries, only the second flagged endOfNode; and :80-88 pushes
        // the synthetic node {leftPtr = rightPtr = ~0, leftAABB = root box, rightAABB = 0}.
        FlatNode node;
        node.left = all;
        node.right = Box{{{0, 0, 0}}, {{0, 0, 0}}};
        node.leftPtr = ~0;
        node.rightPtr = ~0;
        out->t.nodes.push_back(node);
        out->t.order = {0u, 0u};
        out->t.endOfNode = {0, 1};
        out->t.st = b.st;
        return out;
    }
    b.build_plain(root, std::move(refs));
    flatten(b, root, out->t);
    out->t.st = b.st;
    return out;
}

uint64_t oracle_tree_node_count(void* h) { return static_cast<OracleTree*>(h)->t.nodes.size(); }
uint64_t oracle_tree_ref_count(void* h) { return static_cast<OracleTree*>(h)->t.order.size(); }
void oracle_tree_copy_nodes(void* h, void* out) {
    auto& t = static_cast<OracleTree*>(h)->t;
    std::memcpy(out, t.nodes.data(), t.nodes.size() * sizeof(FlatNode));
}
void oracle_tree_copy_order(void* h, uint32_t* order, uint8_t* flags) {
    auto& t = static_cast<OracleTree*>(h)->t;
    std::memcpy(order, t.order.data(), t.order.size() * 4);
    std::memcpy(flags, t.endOfNode.data(), t.endOfNode.size());
}
// stats: medianSplits, sortFallbacks, sortFallbackMaxN, spatialTried, spatialChosen, axisSkipped, maxDepth,
//        sumLeafDepth, duplicates, lastResortLeaf
void oracle_tree_stats(void* h, uint64_t* out10) {
    const Stats& s = static_cast<OracleTree*>(h)->t.st;
    const uint64_t v[10] = {s.medianSplits, s.sortFallbacks, s.sortFallbackMaxN, s.spatialTried, s.spatialChosen,
                            s.axisSkipped, s.maxDepth, s.sumLeafDepth, s.duplicates, s.lastResortLeaf};
    std::memcpy(out10, v, sizeof(v));
}
void oracle_tree_free(void* h) { delete static_cast<OracleTree*>(h); }

// Batch wrappers in the shape of traceClosest.csh:12-36. rays / out: PackedRay, 12 floats each
// (origin.xyz, bits(ID); direction.xyz, [u]; t, bits(hitID), bits(hitInstanceID), [v]). The two lanes the GLSL
// PackRay leaves unwritten (direction.w, hit.w) carry the barycentrics (sol.y, sol.z) of the accepted hit, 0 if none.
// any != 0 selects HitAny with per-ray tMax taken from rays[i].hit.x when perRayTMax != 0.
// counters (6 x u64): tlasNodes, instances, blasNodes, triangles, maxStack, raysWithStack>32.
void oracle_trace(const float* tlasNodes, const float* instances, const float* const* blasNodes,
                  const float* const* bvhTris, const float* rays, uint64_t n, uint32_t cullMask, float tMin,
                  float tMax, int any, int perRayTMax, float* out, uint64_t* counters, int nthreads,
                  const float* const* triangles96, int opacity, const uint32_t* materials, uint32_t materialCount,
                  const uint32_t* texDims, const uint8_t* const* texels, uint32_t textureCount) {
    SceneView sc{tlasNodes, instances, blasNodes, bvhTris, triangles96, materials, materialCount, texDims, texels, textureCount};
    std::vector<Counters> perThread(size_t(std::max(nthreads, 1)));
    parallel_for(n, nthreads, [&](uint64_t b, uint64_t e, int tid) {
        Counters& ct = perThread[size_t(tid)];
        for (uint64_t i = b; i < e; i++) {
            const float* in = rays + 12 * i;
            float* o = out + 12 * i;
            RayState r;
            int32_t id;
            std::memcpy(&id, in + 3, 4);
            for (int a = 0; a < 3; a++) { r.o[a] = in[a]; r.d[a] = in[4 + a]; }
            std::memcpy(&r.hitInstanceID, in + 10, 4);
            r.currentInstanceID = 0;
            r.hitID = -1;
            r.hitDistance = 0.0f;
            r.baryU = r.baryV = 0.0f;
            r.transparency = 1.0f;
            if (id >= 0) {
                if (any) {
                    const float tm = perRayTMax ? in[8] : tMax;
                    r.hitDistance = tm;   // HitAny does not reset hitDistance; report tMax on a miss
                    if (opacity) traverse<true, true>(sc, r, cullMask, tMin, tm, ct);
                    else traverse<true>(sc, r, cullMask, tMin, tm, ct);
                } else {
                    if (opacity) traverse<false, true>(sc, r, cullMask, tMin, tMax, ct);
                    else traverse<false>(sc, r, cullMask, tMin, tMax, ct);
                }
            }
            for (int a = 0; a < 3; a++) { o[a] = r.o[a]; o[4 + a] = r.d[a]; }
            std::memcpy(o + 3, &id, 4);
            o[7] = (any && opacity) ? r.transparency : r.baryU;   // HitAnyTransparency's return value travels in direction.w
            o[8] = r.hitDistance;
            std::memcpy(o + 9, &r.hitID, 4);
            std::memcpy(o + 10, &r.hitInstanceID, 4);
            o[11] = r.baryV;
        }
    });
    if (counters) {
        Counters sum;
        for (const Counters& c : perThread) {
            sum.tlasNodes += c.tlasNodes; sum.instances += c.instances; sum.blasNodes += c.blasNodes;
            sum.triangles += c.triangles; sum.maxStack = std::max(sum.maxStack, c.maxStack);
            sum.stackOverflows += c.stackOverflows;
        }
        const uint64_t v[6] = {sum.tlasNodes, sum.instances, sum.blasNodes, sum.triangles, sum.maxStack, sum.stackOverflows};
        std::memcpy(counters, v, sizeof(v));
    }
}

// Brute force over every instance x triangle in flattened order; used only to sanity-check the traversal above.
// Returns per ray the minimum t (strict <, first in (instance slot, triangle slot) order wins) — IDs may differ
// from traversal order on exact ties, so tests compare t and accept any ID with that t.
void oracle_brute_force(const float* instances, uint32_t instanceCount, const float* const* bvhTris,
                        const uint32_t* triCounts, const float* rays, uint64_t n, uint32_t cullMask, float tMin,
                        float tMax, float* outT, int32_t* outTri, int32_t* outInst, int nthreads) {
    parallel_for(n, nthreads, [&](uint64_t b, uint64_t e, int) {
        for (uint64_t i = b; i < e; i++) {
            const float* in = rays + 12 * i;
            float best = tMax;
            int32_t bt = -1, bi = -1;
            for (uint32_t k = 0; k < instanceCount; k++) {
                const float* I = instances + 16 * size_t(k);
                int32_t meshOffset, mask;
                std::memcpy(&meshOffset, I + 12, 4);
                std::memcpy(&mask, I + 15, 4);
                if (!((uint32_t(mask) & cullMask) > 0u)) continue;
                RayState r;
                for (int c = 0; c < 3; c++) {
                    const float* col = I + 4 * c;
                    r.o[c] = ((in[0] * col[0] + in[1] * col[1]) + in[2] * col[2]) + 1.0f * col[3];
                    r.d[c] = ((in[4] * col[0] + in[5] * col[1]) + in[6] * col[2]) + 0.0f * col[3];
                }
                for (uint32_t t = 0; t < triCounts[meshOffset]; t++) {
                    const float* T = bvhTris[meshOffset] + 12 * size_t(t);
                    float sol[3];
                    if (tri_test(r, T, T + 4, T + 8, sol) && sol[0] > tMin && sol[0] < best) { best = sol[0]; bt = int32_t(t); bi = int32_t(k); }
                }
            }
            outT[i] = best; outTri[i] = bt; outInst[i] = bi;
        }
    });
}

}
