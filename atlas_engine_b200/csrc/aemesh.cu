// aemesh.cu — reader for the engine's .aemesh mesh files and the triangle expansion of MeshData::BuildBVH, i.e. the
// step right before the BLAS build. Host code only (compiled by nvcc with the rest of the library).
//
// Format (reference: src/engine/loader/MeshLoader.cpp:7-37, src/engine/mesh/MeshSerializer.cpp:7-133, MeshSerializer.h:28-70):
// a MessagePack-encoded JSON object (nlohmann json::to_msgpack); mesh["data"] holds indexCount, vertexCount, subData[]
// {indicesOffset, indicesCount, materialIdx, ...} and the components indices / vertices / texCoords / normals / tangents /
// colors, each {"format": ComponentFormat, "data": bin} where the blob is the raw element array of DataComponent<T>:
// uint32_t, vec3, vec2, vec4, vec4, vec4 (mesh/MeshData.h:98-104).
//
// Triangle expansion (src/engine/mesh/MeshData.cpp:89-159): for every sub mesh, for every index triple: the three
// positions, normalize(vec4 normal) truncated to vec3, texture coordinates (zero when the mesh has none), vertex colours
// (one when it has none), the sub mesh's material index, and the box glm::min/max(glm::min/max(v0, v1), v2).
// NOTE the reference indexes triangle k of the WHOLE mesh as indices[3k..3k+2] and ignores sub.indicesOffset; so do we.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include "../../include/atlas_rt.h"

namespace {

// ---- a small MessagePack reader (the subset nlohmann::json emits: nil, bool, ints, float32/64, str, bin, array, map)
struct Value {
    enum Kind { Nil, Bool, Int, Float, Str, Bin, Arr, Map } kind = Nil;
    int64_t i = 0;
    double f = 0.0;
    const uint8_t* p = nullptr;   // Str / Bin payload inside the file buffer
    size_t n = 0;
    std::vector<Value> arr;
    std::vector<std::pair<std::string, Value>> map;

    const Value* get(const char* key) const {
        for (const auto& kv : map)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
    double number() const { return kind == Int ? double(i) : f; }
};

struct Reader {
    const uint8_t* d;
    size_t n, pos = 0;
    bool ok = true;
    int depth = 0;

    uint64_t be(size_t bytes) {
        if (pos + bytes > n) { ok = false; return 0; }
        uint64_t v = 0;
        for (size_t k = 0; k < bytes; k++) v = (v << 8) | d[pos + k];
        pos += bytes;
        return v;
    }
    void blob(Value& v, Value::Kind kind, size_t len) {
        if (pos + len > n) { ok = false; return; }
        v.kind = kind; v.p = d + pos; v.n = len;
        pos += len;
    }
    void array(Value& v, size_t len) {
        v.kind = Value::Arr;
        if (len > n) { ok = false; return; }
        v.arr.resize(len);
        for (size_t k = 0; k < len && ok; k++) parse(v.arr[k]);
    }
    void object(Value& v, size_t len) {
        v.kind = Value::Map;
        if (len > n) { ok = false; return; }
        v.map.resize(len);
        for (size_t k = 0; k < len && ok; k++) {
            Value key;
            parse(key);
            if (key.kind != Value::Str) { ok = false; return; }
            v.map[k].first.assign(reinterpret_cast<const char*>(key.p), key.n);
            parse(v.map[k].second);
        }
    }
    void parse(Value& v) {
        if (!ok || pos >= n || ++depth > 64) { ok = false; return; }
        const uint8_t t = d[pos++];
        if (t <= 0x7f) { v.kind = Value::Int; v.i = t; }
        else if (t >= 0xe0) { v.kind = Value::Int; v.i = int8_t(t); }
        else if (t >= 0x80 && t <= 0x8f) object(v, t & 0x0f);
        else if (t >= 0x90 && t <= 0x9f) array(v, t & 0x0f);
        else if (t >= 0xa0 && t <= 0xbf) blob(v, Value::Str, t & 0x1f);
        else switch (t) {
            case 0xc0: v.kind = Value::Nil; break;
            case 0xc2: v.kind = Value::Bool; v.i = 0; break;
            case 0xc3: v.kind = Value::Bool; v.i = 1; break;
            case 0xc4: blob(v, Value::Bin, size_t(be(1))); break;
            case 0xc5: blob(v, Value::Bin, size_t(be(2))); break;
            case 0xc6: blob(v, Value::Bin, size_t(be(4))); break;
            case 0xca: { const uint32_t b = uint32_t(be(4)); float x; memcpy(&x, &b, 4); v.kind = Value::Float; v.f = x; break; }
            case 0xcb: { const uint64_t b = be(8); double x; memcpy(&x, &b, 8); v.kind = Value::Float; v.f = x; break; }
            case 0xcc: v.kind = Value::Int; v.i = int64_t(be(1)); break;
            case 0xcd: v.kind = Value::Int; v.i = int64_t(be(2)); break;
            case 0xce: v.kind = Value::Int; v.i = int64_t(be(4)); break;
            case 0xcf: v.kind = Value::Int; v.i = int64_t(be(8)); break;
            case 0xd0: v.kind = Value::Int; v.i = int8_t(be(1)); break;
            case 0xd1: v.kind = Value::Int; v.i = int16_t(be(2)); break;
            case 0xd2: v.kind = Value::Int; v.i = int32_t(be(4)); break;
            case 0xd3: v.kind = Value::Int; v.i = int64_t(be(8)); break;
            case 0xd9: blob(v, Value::Str, size_t(be(1))); break;
            case 0xda: blob(v, Value::Str, size_t(be(2))); break;
            case 0xdb: blob(v, Value::Str, size_t(be(4))); break;
            case 0xdc: array(v, size_t(be(2))); break;
            case 0xdd: array(v, size_t(be(4))); break;
            case 0xde: object(v, size_t(be(2))); break;
            case 0xdf: object(v, size_t(be(4))); break;
            default: ok = false;   // ext types: nlohmann never writes them for this schema
        }
        depth--;
    }
};

struct SubMesh { uint32_t indicesOffset = 0, indicesCount = 0; int32_t materialIdx = 0; };

template <typename T>
bool component(const Value& data, const char* name, size_t elemBytes, std::vector<T>& out) {
    const Value* c = data.get(name);
    if (!c || c->kind != Value::Map) return false;
    const Value* blob = c->get("data");
    if (!blob) return false;
    if (blob->kind == Value::Bin) {
        if (blob->n % elemBytes) return false;
        out.resize(blob->n / sizeof(T));
        if (blob->n) memcpy(out.data(), blob->p, blob->n);
        return true;
    }
    if (blob->kind == Value::Arr) {   // "binary": false files store the bytes as a JSON array of numbers
        std::vector<uint8_t> bytes(blob->arr.size());
        for (size_t k = 0; k < bytes.size(); k++) bytes[k] = uint8_t(blob->arr[k].i);
        if (bytes.size() % elemBytes) return false;
        out.resize(bytes.size() / sizeof(T));
        if (!bytes.empty()) memcpy(out.data(), bytes.data(), bytes.size());
        return true;
    }
    return false;
}

}   // namespace

struct atlas_rt_aemesh {
    std::string name;
    std::vector<uint32_t> indices;
    std::vector<float> vertices, texCoords, normals, tangents, colors;   // 3 / 2 / 4 / 4 / 4 floats per vertex
    std::vector<SubMesh> subs;
    std::vector<std::string> materials;
    uint64_t triangleCount = 0;
};

extern "C" {

int atlas_rt_aemesh_open(const char* path, atlas_rt_aemesh** out_mesh) {
    if (!path || !out_mesh) return ATLAS_RT_ERR_INVALID;
    *out_mesh = nullptr;
    FILE* f = fopen(path, "rb");
    if (!f) return ATLAS_RT_ERR_INVALID;
    std::vector<uint8_t> buf;
    uint8_t chunk[1 << 16];
    for (size_t got; (got = fread(chunk, 1, sizeof(chunk), f)) > 0;) buf.insert(buf.end(), chunk, chunk + got);
    fclose(f);
    Reader r{buf.data(), buf.size()};
    Value root;
    r.parse(root);
    if (!r.ok || root.kind != Value::Map) return ATLAS_RT_ERR_UNSUPPORTED;
    const Value* data = root.get("data");
    if (!data || data->kind != Value::Map) return ATLAS_RT_ERR_UNSUPPORTED;
    auto* m = new (std::nothrow) atlas_rt_aemesh;
    if (!m) return ATLAS_RT_ERR_OOM;
    bool ok = component(*data, "indices", 4, m->indices) && component(*data, "vertices", 12, m->vertices) &&
              component(*data, "texCoords", 8, m->texCoords) && component(*data, "normals", 16, m->normals) &&
              component(*data, "tangents", 16, m->tangents) && component(*data, "colors", 16, m->colors);
    if (const Value* nm = data->get("name")) if (nm->kind == Value::Str) m->name.assign(reinterpret_cast<const char*>(nm->p), nm->n);
    if (const Value* mats = data->get("materials"))
        for (const Value& v : mats->arr) if (v.kind == Value::Str) m->materials.emplace_back(reinterpret_cast<const char*>(v.p), v.n);
    const Value* subs = data->get("subData");
    ok = ok && subs && subs->kind == Value::Arr;
    const uint64_t vertexCount = m->vertices.size() / 3;
    if (ok) {
        for (const Value& s : subs->arr) {
            SubMesh sm;
            const Value *a = s.get("indicesOffset"), *b = s.get("indicesCount"), *c = s.get("materialIdx");
            if (!a || !b || !c) { ok = false; break; }
            sm.indicesOffset = uint32_t(a->number());
            sm.indicesCount = uint32_t(b->number());
            sm.materialIdx = int32_t(c->number());
            m->subs.push_back(sm);
            m->triangleCount += sm.indicesCount / 3;
        }
    }
    // what MeshData::BuildBVH dereferences must exist: indices for every triangle, a position and a normal per index
    ok = ok && m->triangleCount * 3 <= m->indices.size() && m->normals.size() / 4 >= vertexCount &&
         (m->texCoords.empty() || m->texCoords.size() / 2 >= vertexCount) && (m->colors.empty() || m->colors.size() / 4 >= vertexCount);
    for (size_t k = 0; ok && k < m->triangleCount * 3; k++) ok = m->indices[k] < vertexCount;
    if (!ok) { delete m; return ATLAS_RT_ERR_UNSUPPORTED; }
    *out_mesh = m;
    return ATLAS_RT_OK;
}

int atlas_rt_aemesh_counts(const atlas_rt_aemesh* m, uint64_t* vertex_count, uint64_t* index_count, uint64_t* triangle_count, uint32_t* sub_mesh_count) {
    if (!m) return ATLAS_RT_ERR_INVALID;
    if (vertex_count) *vertex_count = m->vertices.size() / 3;
    if (index_count) *index_count = m->indices.size();
    if (triangle_count) *triangle_count = m->triangleCount;
    if (sub_mesh_count) *sub_mesh_count = uint32_t(m->subs.size());
    return ATLAS_RT_OK;
}

const char* atlas_rt_aemesh_material_path(const atlas_rt_aemesh* m, uint32_t index) {
    return (m && index < m->materials.size()) ? m->materials[index].c_str() : nullptr;
}

int atlas_rt_aemesh_triangles(const atlas_rt_aemesh* m, float* tris9, float* aabbs6, int32_t* material_idx, float* normals9,
                              float* uvs6, float* colors12) {
    if (!m) return ATLAS_RT_ERR_INVALID;
    const bool hasUV = !m->texCoords.empty(), hasColor = !m->colors.empty();
    uint64_t base = 0;
    for (const SubMesh& sub : m->subs) {
        const uint64_t count = sub.indicesCount / 3;
        for (uint64_t i = 0; i < count; i++) {
            const uint64_t k = i + base;
            const uint32_t idx[3] = {m->indices[3 * k], m->indices[3 * k + 1], m->indices[3 * k + 2]};
            float v[3][3];
            for (int c = 0; c < 3; c++) memcpy(v[c], &m->vertices[3 * size_t(idx[c])], 12);
            if (tris9) memcpy(tris9 + 9 * k, v, 36);
            if (aabbs6) {
                for (int a = 0; a < 3; a++) {   // glm::min(x, y) = (y < x) ? y : x, glm::max(x, y) = (x < y) ? y : x
                    float lo = v[0][a], hi = v[0][a];
                    lo = (v[1][a] < lo) ? v[1][a] : lo; lo = (v[2][a] < lo) ? v[2][a] : lo;
                    hi = (hi < v[1][a]) ? v[1][a] : hi; hi = (hi < v[2][a]) ? v[2][a] : hi;
                    aabbs6[6 * k + a] = lo;
                    aabbs6[6 * k + 3 + a] = hi;
                }
            }
            if (material_idx) material_idx[k] = sub.materialIdx;
            if (normals9) {
                for (int c = 0; c < 3; c++) {   // glm::normalize(vec4) = v * (1 / sqrt(dot)), dot = (x*x + y*y) + (z*z + w*w); then vec3(vec4)
                    const float* n = &m->normals[4 * size_t(idx[c])];
                    const float d = (n[0] * n[0] + n[1] * n[1]) + (n[2] * n[2] + n[3] * n[3]);
                    const float inv = 1.0f / std::sqrt(d);
                    for (int a = 0; a < 3; a++) normals9[9 * k + 3 * c + a] = n[a] * inv;
                }
            }
            if (uvs6) for (int c = 0; c < 3; c++) for (int a = 0; a < 2; a++) uvs6[6 * k + 2 * c + a] = hasUV ? m->texCoords[2 * size_t(idx[c]) + a] : 0.0f;
            if (colors12) for (int c = 0; c < 3; c++) for (int a = 0; a < 4; a++) colors12[12 * k + 4 * c + a] = hasColor ? m->colors[4 * size_t(idx[c]) + a] : 1.0f;
        }
        base += count;
    }
    return ATLAS_RT_OK;
}

int atlas_rt_aemesh_raw(const atlas_rt_aemesh* m, const uint32_t** indices, const float** vertices3, const float** normals4, const float** tex_coords2) {
    if (!m) return ATLAS_RT_ERR_INVALID;
    if (indices) *indices = m->indices.data();
    if (vertices3) *vertices3 = m->vertices.data();
    if (normals4) *normals4 = m->normals.data();
    if (tex_coords2) *tex_coords2 = m->texCoords.empty() ? nullptr : m->texCoords.data();
    return ATLAS_RT_OK;
}

void atlas_rt_aemesh_close(atlas_rt_aemesh* m) { delete m; }

}   // extern "C"
