/*
 * atlas_rt.h — C ABI of the B200-native (CUDA, sm_100a) replacement for Atlas Engine's software ray-tracing
 * acceleration path: BLAS/TLAS build, flatten + pack into the engine's GPU layouts, closest-hit / any-hit traversal,
 * and the path tracer's ray-gen / diffuse-bounce loop.
 *
 * This header is the drop-in boundary. The reference has no FFI for this path (it is plain C++ classes), so every
 * entry point cites the reference interface it stands behind (paths relative to the reference repository root);
 * atlas_engine_b200/host/ holds C++ classes with the reference's own names that forward to these functions, and
 * INTEGRATION.md shows the patch a maintainer applies.
 *
 * Conventions
 *   - Every function returns ATLAS_RT_OK (0) or a negative atlas_rt_status; nothing throws across the boundary.
 *     atlas_rt_last_error(ctx) returns a human-readable message for the most recent failure on that context.
 *   - Plain pointers and sizes only. A pointer argument is HOST memory unless the matching ATLAS_RT_DEVICE_* bit is
 *     set in `flags`, in which case it is a CUDA device pointer valid on the context's device.
 *   - All work for a context is issued on the context's CUDA stream. Calls are synchronous on return (like the
 *     reference's constructors) unless ATLAS_RT_ASYNC is set, in which case the caller synchronises the stream
 *     (atlas_rt_context_synchronize) before touching outputs. Device outputs stay valid until the owning object is freed.
 *   - Contexts are independent: one per host thread / stream makes concurrent builds re-entrant, as the reference
 *     requires (meshes are built concurrently from job-system workers, src/tests/App.cpp:362-370).
 *   - There is no CPU fallback: if no CUDA device is usable the functions fail with ATLAS_RT_ERR_CUDA.
 *
 * Layouts (all little-endian, 4-byte words; asserted in atlas_engine_b200/csrc/layouts.h)
 *   AABB            24 B  min.xyz, max.xyz                                 src/engine/volume/AABB.h:102-103
 *   triangle        36 B  v0.xyz, v1.xyz, v2.xyz                           src/engine/volume/BVH.h:26-33 (vertices only)
 *   BVHNode         56 B  leftAABB, rightAABB, leftPtr, rightPtr           src/engine/volume/BVH.h:14-24
 *   GPUBVHNode      64 B  BVHNode + 2 pad ints                             src/engine/raytracing/RTStructures.h:95-103
 *   GPUBVHTriangle  48 B  v0.xyz,endOfNode?1:-1; v1.xyz,bits(material); v2.xyz,opacity   RTStructures.h:23-27, mesh/MeshData.cpp:242-247
 *   GPUBVHInstance  64 B  mat3x4 inverseMatrix; meshOffset, materialOffset, nextInstance, mask   RTStructures.h:85-93
 *   PackedRay       48 B  origin.xyz,bits(ID); direction.xyz,[u]; t,bits(hitID),bits(hitInstanceID),[v]
 *                         data/shader/raytracer/structures.hsh:9-13, common.hsh:46-73. The two lanes the GLSL PackRay
 *                         leaves unwritten (direction.w, hit.w) carry the barycentrics (sol.y, sol.z) of the accepted hit.
 *   ptr encoding    ptr >= 0: inner-node index; ptr < 0: leaf, ~ptr = first triangle / instance slot.
 */
#ifndef ATLAS_RT_H
#define ATLAS_RT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ATLAS_RT_VERSION 1

typedef enum atlas_rt_status {
    ATLAS_RT_OK = 0,
    ATLAS_RT_ERR_INVALID = -1,      /* null pointer, mismatched sizes, object from another context */
    ATLAS_RT_ERR_CUDA = -2,         /* a CUDA call failed or no device */
    ATLAS_RT_ERR_OOM = -3,          /* device or host allocation failed */
    ATLAS_RT_ERR_UNSUPPORTED = -4,  /* input outside the documented contract (e.g. more than 2^30-1 primitives in one build) */
    ATLAS_RT_ERR_STACK = -5         /* a ray needed more than ATLAS_RT_STACK_SIZE stack entries (UB in the reference) */
} atlas_rt_status;

/* flags */
#define ATLAS_RT_DEVICE_INPUT  (1u << 0)   /* input pointers are device memory */
#define ATLAS_RT_DEVICE_OUTPUT (1u << 1)   /* output pointers are device memory */
#define ATLAS_RT_ASYNC         (1u << 2)   /* do not synchronise the stream before returning */
#define ATLAS_RT_PER_RAY_TMAX  (1u << 3)   /* trace_any: take tMax from each ray's hit.x instead of the argument */
#define ATLAS_RT_OPACITY       (1u << 5)   /* trace_*: the *Transparency variants over the 96-byte triangles (needs atlas_rt_mesh_pack_shading) */
#define ATLAS_RT_COUNTERS      (1u << 4)   /* trace_*: also count visited nodes / triangles (slower; for parity + roofline) */
#define ATLAS_RT_RAY_BINNING   (1u << 6)   /* pathtrace_bounces: order the rays of bounce >= 1 by the reference's 8x8 octahedral direction bins */
#define ATLAS_RT_ACCUM_TILE_ORDER (1u << 7) /* pathtrace_*: index the accumulation buffer in rayGen's tile order instead of y * width + x */
#define ATLAS_RT_PIPELINED     (1u << 10)  /* trace_* with host buffers and ATLAS_RT_ASYNC: successive calls overlap; atlas_rt_trace_join before the results are used */
#define ATLAS_RT_PEER_OUTPUT   (1u << 9)   /* trace_sharded: every rank's traversal kernel stores its hit records straight into the root's memory (NVLink P2P), no gather */
#define ATLAS_RT_HITS_ONLY     (1u << 8)   /* trace_*: rays_out receives count x 16-byte hit records (the ray's `hit` vec4: t, bits(hitID),
                                              bits(hitInstanceID), v) instead of 48-byte rays — the compact stream the multi-GPU gather moves */

/* Instance cull masks — InstanceCullMasks, src/engine/raytracing/RTStructures.h:9-12; common.hsh:14-15. */
#define ATLAS_RT_MASK_ALL    (1u << 7)
#define ATLAS_RT_MASK_SHADOW (1u << 6)

#define ATLAS_RT_STACK_SIZE 32             /* STACK_SIZE, data/shader/raytracer/bvh.hsh:16 */
#define ATLAS_RT_INF 1000000000000.0f      /* INF, data/shader/raytracer/common.hsh:8 */

typedef struct atlas_rt_context atlas_rt_context;
typedef struct atlas_rt_bvh atlas_rt_bvh;
typedef struct atlas_rt_mesh atlas_rt_mesh;
typedef struct atlas_rt_scene atlas_rt_scene;

/* ------------------------------------------------------------------------------------------------ context ---- */
/* device: CUDA ordinal. stream: a cudaStream_t owned by the caller (e.g. torch's current stream), or NULL to let the
 * context create and own a non-blocking stream. Replaces the reference's implicit globals (Graphics::GraphicsDevice::
 * DefaultDevice + the JobSystem pools used by BVH.cpp:36-38). */
int atlas_rt_context_create(int device, void* stream, atlas_rt_context** out_ctx);
void atlas_rt_context_destroy(atlas_rt_context* ctx);
int atlas_rt_context_synchronize(atlas_rt_context* ctx);
const char* atlas_rt_last_error(const atlas_rt_context* ctx);
/* Number of CUDA kernels this context has launched so far (bench.py's gpu_launches). */
uint64_t atlas_rt_kernel_launches(const atlas_rt_context* ctx);
int atlas_rt_version(void);

/* -------------------------------------------------------------------------------------------------- build ---- */
/* BLAS build + flatten. Replaces Atlas::Volume::BVH::BVH(const std::vector<AABB>&, const std::vector<BVHTriangle>&,
 * bool) — src/engine/volume/BVH.cpp:14-56 — including the root SBVH spatial split, 256 bins, one primitive per leaf,
 * median fallback and the larger-area-child-first flatten (BVH.cpp:249-441).
 * aabbs: count x 6 floats. tris: count x 9 floats (v0,v1,v2 of triangle i; its BVHTriangle::idx is i).
 * The result holds nodes, the flattened order (source index per slot, duplicates possible) and endOfNode flags. */
int atlas_rt_build_blas(atlas_rt_context* ctx, const float* aabbs, const float* tris, uint64_t count, uint32_t flags,
                        atlas_rt_bvh** out_bvh);

/* The BLASes of many meshes in one call: aabbs[m] / tris[m] / counts[m] as for atlas_rt_build_blas, out_bvhs[m] receives
 * mesh m's tree. Replaces the engine building its meshes concurrently on job-system workers (src/tests/App.cpp:362-370,
 * src/demo/App.cpp:1092 -> Mesh::MeshData::BuildBVH): up to 8 builds run side by side on the device, so a scene of many
 * small meshes is not serialised behind per-build launch latency. Every tree is identical to the one atlas_rt_build_blas
 * returns for that mesh. */
int atlas_rt_build_blas_batch(atlas_rt_context* ctx, uint32_t mesh_count, const float* const* aabbs, const float* const* tris,
                              const uint64_t* counts, uint32_t flags, atlas_rt_bvh** out_bvhs);

/* TLAS build + flatten. Replaces Atlas::Volume::BVH::BVH(const std::vector<AABB>&, bool) — BVH.cpp:58-101 (64 bins,
 * object/median splits only, the count == 1 special node and its two-entry refs quirk). */
int atlas_rt_build_tlas(atlas_rt_context* ctx, const float* aabbs, uint64_t count, uint32_t flags,
                        atlas_rt_bvh** out_bvh);

/* Wrap an already flattened tree (host BVHNode 56 B layout + order + flags), e.g. one built by the reference, so it
 * can be packed and traversed. Mirrors assigning BVH::nodes / data / refs directly (BVH.h:132-136). */
int atlas_rt_bvh_upload(atlas_rt_context* ctx, const void* nodes56, uint64_t node_count, const uint32_t* order,
                        const uint8_t* end_of_node, uint64_t ref_count, atlas_rt_bvh** out_bvh);

/* Same as atlas_rt_bvh_upload but honouring ATLAS_RT_DEVICE_INPUT: the three arrays may already live on the context's
 * device (e.g. a BLAS another GPU built and sent over NCCL — instance sets are built round-robin across GPUs and
 * exchanged, atlas_engine_b200/sharding.py). */
int atlas_rt_bvh_import(atlas_rt_context* ctx, const void* nodes56, uint64_t node_count, const uint32_t* order,
                        const uint8_t* end_of_node, uint64_t ref_count, uint32_t flags, atlas_rt_bvh** out_bvh);

/* nodes.size() and data.size()/refs.size() of the reference object. */
int atlas_rt_bvh_counts(const atlas_rt_bvh* bvh, uint64_t* node_count, uint64_t* ref_count);

/* Copy out in the reference's host layout: BVH::nodes (56 B each), data[i].idx / refs[i].idx, endOfNode.
 * Any pointer may be NULL to skip it. Honours ATLAS_RT_DEVICE_OUTPUT. */
int atlas_rt_bvh_download(const atlas_rt_bvh* bvh, void* nodes56, uint32_t* order, uint8_t* end_of_node, uint32_t flags);

/* Device pointers owned by the bvh: GPUBVHNode array (64 B each), order (u32), endOfNode (u8). */
int atlas_rt_bvh_device_ptrs(const atlas_rt_bvh* bvh, const void** gpu_nodes64, const uint32_t** order,
                             const uint8_t** end_of_node);

/* Build statistics: out[0] = 1 if the root spatial split was evaluated, [1] = 1 if chosen, [2] = duplicated
 * references, [3] = median splits, [4] = sort fallbacks, [5] = largest sort fallback, [6] = levels processed,
 * [7] = 1 if the input contained -0.0 (node boxes may then differ from the reference in the sign of zero only). */
int atlas_rt_bvh_stats(const atlas_rt_bvh* bvh, uint64_t out[8]);

void atlas_rt_bvh_free(atlas_rt_bvh* bvh);

/* --------------------------------------------------------------------------------------------------- pack ---- */
/* Gather triangles into flattened order and emit GPUBVHTriangle records + keep the GPUBVHNode array. Replaces the
 * packing half of Atlas::Mesh::MeshData::BuildBVH — src/engine/mesh/MeshData.cpp:172-269 (traversal payload only:
 * v0..v2, endOfNode, material index, opacity) — and the uploads of Mesh::BuildBVH, mesh/Mesh.cpp:97-108.
 * tris: the same count x 9 array the BLAS was built from. material_idx / opacity: per-triangle arrays (count entries,
 * source order) or NULL for 0 / 1.0f. */
int atlas_rt_pack_mesh(atlas_rt_context* ctx, const atlas_rt_bvh* blas, const float* tris, uint64_t count,
                       const int32_t* material_idx, const float* opacity, uint32_t flags, atlas_rt_mesh** out_mesh);
/* Adds the 96-byte GPUTriangle array ("triangles[]", RTStructures.h:14-21) that the opacity-aware traversal variants
 * and the hit shaders read — the other half of mesh/MeshData.cpp:172-239. Per flattened slot i (source triangle
 * k = order[i]): v0 = (tris[k].v0, p[0]), v1 = (.., p[1]), v2 = (.., p[2]), d0 = (p[3], p[4], p[5], bits(material_idx[k])),
 * d1 = (p[6], p[7], endOfNode ? 1 : -1, 0), d2 = (p[8], p[9], p[10], opacity[k]) where p = payload11 + 11*k are the
 * engine's already packed shading words (normals 10-10-10-2, uv half2, tangent, bitangent, colours unorm8 — produced by
 * Common::Packing / glm on the engine side exactly as today; they are opaque to this library). payload11 may be NULL
 * (zeros). opacity < 0 marks a textured-opacity triangle (MeshData.cpp:136). */
int atlas_rt_mesh_pack_shading(atlas_rt_context* ctx, atlas_rt_mesh* mesh, const float* tris, uint64_t count,
                               const int32_t* material_idx, const float* opacity, const uint32_t* payload11, uint32_t flags);
/* Copy out the GPUTriangle array (96 B each). */
int atlas_rt_mesh_download_shading(const atlas_rt_mesh* mesh, void* gpu_triangles96, uint32_t flags);
int atlas_rt_mesh_counts(const atlas_rt_mesh* mesh, uint64_t* node_count, uint64_t* triangle_count);
/* Copy out gpuBvhNodes (64 B each) and gpuBvhTriangles (48 B each); either may be NULL. */
int atlas_rt_mesh_download(const atlas_rt_mesh* mesh, void* gpu_nodes64, void* gpu_bvh_triangles48, uint32_t flags);
void atlas_rt_mesh_free(atlas_rt_mesh* mesh);

/* The 11 packed shading words per SOURCE triangle that atlas_rt_mesh_pack_shading takes as `payload11`, computed on the
 * device: pn0, pn1, pn2 (10-10-10-2 signed), puv0, puv1, puv2 (half2), pt, pbt (tangent frame from positions + uvs),
 * pc0, pc1, pc2 (unorm4x8). Replaces the arithmetic of the second loop of MeshData::BuildBVH —
 * src/engine/mesh/MeshData.cpp:176-228 with Common::Packing::PackSignedVector3x10_1x2 (src/engine/common/Packing.cpp:24-35),
 * glm::packHalf2x16 and glm::packUnorm4x8 (glm 0.9.8), including x86's float->int conversion of the NaN tangents that
 * degenerate texture coordinates produce. Inputs are per triangle, expanded as MeshData.cpp:102-133 does: tris count x 9,
 * normals9 count x 9 (already normalised n0 n1 n2; NULL = zero), uvs6 count x 6 (NULL = zero), colors12 count x 12
 * (NULL = one). payload11: count x 11 words. */
int atlas_rt_pack_shading_words(atlas_rt_context* ctx, const float* tris, const float* normals9, const float* uvs6,
                                const float* colors12, uint64_t count, uint32_t* payload11, uint32_t flags);

/* ------------------------------------------------------------------------------------------ mesh ingestion ---- */
/* Reader for the engine's .aemesh files (MessagePack-encoded JSON written by Loader::MeshLoader::SaveMesh —
 * src/engine/loader/MeshLoader.cpp:7-37, src/engine/mesh/MeshSerializer.cpp:62-133, MeshSerializer.h:52-70) and the
 * triangle expansion of the first loop of MeshData::BuildBVH (src/engine/mesh/MeshData.cpp:89-159), whose output feeds
 * atlas_rt_build_blas / atlas_rt_pack_mesh / atlas_rt_pack_shading_words directly. Host-only; needs no context. */
typedef struct atlas_rt_aemesh atlas_rt_aemesh;
int atlas_rt_aemesh_open(const char* path, atlas_rt_aemesh** out_mesh);
int atlas_rt_aemesh_counts(const atlas_rt_aemesh* mesh, uint64_t* vertex_count, uint64_t* index_count,
                           uint64_t* triangle_count, uint32_t* sub_mesh_count);
/* Path of material `index` as stored in the file ("materials/....aematerial"), or NULL. */
const char* atlas_rt_aemesh_material_path(const atlas_rt_aemesh* mesh, uint32_t index);
/* Per triangle k (all sub meshes, in order): tris9 = positions, aabbs6 = glm::min/max box, material_idx = the sub mesh's
 * materialIdx, normals9 = normalize(vec4 normal).xyz per corner, uvs6 (zero without texCoords), colors12 (one without
 * colours). Any pointer may be NULL. */
int atlas_rt_aemesh_triangles(const atlas_rt_aemesh* mesh, float* tris9, float* aabbs6, int32_t* material_idx,
                              float* normals9, float* uvs6, float* colors12);
/* Borrowed pointers to the raw components (valid until close): indices u32, vertices 3 floats, normals 4 floats,
 * texCoords 2 floats (NULL if absent). */
int atlas_rt_aemesh_raw(const atlas_rt_aemesh* mesh, const uint32_t** indices, const float** vertices3,
                        const float** normals4, const float** tex_coords2);
void atlas_rt_aemesh_close(atlas_rt_aemesh* mesh);

/* -------------------------------------------------------------------------------------------------- scene ---- */
/* Assemble the two-level scene. Replaces RayTracingWorld::UpdateForSoftwareRayTracing — src/engine/raytracing/
 * RayTracingWorld.cpp:267-307: instances (64 B GPUBVHInstance each, SOURCE order, meshOffset indexing `meshes`) are
 * permuted by the TLAS order and nextInstance is set to endOfNode ? -1 : slot + 1; TLAS nodes go to the GPUBVHNode
 * layout. The meshes and the tlas must outlive the scene. */
int atlas_rt_scene_create(atlas_rt_context* ctx, const atlas_rt_mesh* const* meshes, uint32_t mesh_count,
                          const void* instances64, uint64_t instance_count, const atlas_rt_bvh* tlas, uint32_t flags,
                          atlas_rt_scene** out_scene);
/* Reordered instance records (tlas ref_count x 64 B) and TLAS nodes (64 B each); either may be NULL. */
int atlas_rt_scene_download(const atlas_rt_scene* scene, void* instances64, void* tlas_nodes64, uint32_t flags);
void atlas_rt_scene_free(atlas_rt_scene* scene);

/* -------------------------------------------------------------------------------------------------- trace ---- */
/* Closest hit for a batch of PackedRay (48 B). Replaces the traceClosest.csh dispatch issued by
 * RayTracingHelper::DispatchHitClosest (src/engine/renderer/helper/RayTracingHelper.cpp:346-364) = HitClosest,
 * data/shader/raytracer/bvh.hsh:191-273: rays with ID < 0 are passed through with hitID = -1, t = 0; otherwise
 * t = tMax and hitID = -1 on a miss. rays_out may alias rays_in (an in-place batch only writes the hit fields).
 * ATLAS_RT_DEVICE_INPUT / ATLAS_RT_DEVICE_OUTPUT are independent: with host rays (pinned memory pays off) batches of
 * 262144 rays or more are uploaded, traced and, for host output, downloaded in overlapping chunks; host rays with
 * device output leave the hits on the device, e.g. for an NCCL gather. */
int atlas_rt_trace_closest(atlas_rt_context* ctx, const atlas_rt_scene* scene, const void* rays_in, uint64_t count,
                           uint32_t cull_mask, float t_min, float t_max, void* rays_out, uint32_t flags);

/* Any hit (shadow rays). HitAny, bvh.hsh:359-441, as called inline by the reference's hit shaders
 * (data/shader/pathtracer/rayHit.csh:327-337). Output hitID >= 0 iff something was hit in (tMin, tMax).
 *
 * With ATLAS_RT_OPACITY both calls run the shader's *Transparency variants over the 96-byte triangles instead
 * (what the path tracer actually dispatches: traceClosest.csh with OPACITY_CHECK, rayHit.csh:331):
 *   closest  HitClosestTransparency, bvh.hsh:275-357: a triangle is only accepted if its opacity is > 0;
 *   any      HitAnyTransparency, bvh.hsh:443-524: walks on until the accumulated transparency reaches 0; the returned
 *            transparency is written to direction.w, and t / hitID / instanceID hold the LAST triangle intersected
 *            (exactly what the shader leaves in the ray), including the shader's `transparency *= leaf(transparency)` form.
 * Triangles with textured opacity (opacity < 0) are resolved through GetOpacity (surface.hsh:147-160) once the scene has
 * its material / texture tables (atlas_rt_scene_set_materials); without them they count as opacity 1.
 * With ATLAS_RT_HITS_ONLY rays_out receives 16-byte hit records instead of rays (see the flag). */
int atlas_rt_trace_any(atlas_rt_context* ctx, const atlas_rt_scene* scene, const void* rays_in, uint64_t count,
                       uint32_t cull_mask, float t_min, float t_max, void* rays_out, uint32_t flags);

/* Traversal work counters of the last trace call on this context that passed ATLAS_RT_COUNTERS, summed over rays:
 * out = tlasNodes, instances entered, blasNodes, triangles tested, max stack depth, rays that overflowed the stack.
 * These are the visit counts SURVEY.md 8(d) turns into algorithmic bytes. */
int atlas_rt_trace_counters(atlas_rt_context* ctx, uint64_t out[6]);

/* Host-buffer trace calls made with ATLAS_RT_ASYNC | ATLAS_RT_PIPELINED (host rays in, host results out) do not order their
 * completion into the context stream: the next such call starts uploading while the previous one's last chunks are still being
 * traced and downloaded - a renderer's double-buffered batches (different host output buffers; the same input buffer may be
 * reused) keep PCIe and the GPU busy across call boundaries. atlas_rt_trace_join orders the context stream after every such
 * call made so far; the results are in the host buffers once the context stream has been synchronised after it
 * (atlas_rt_context_synchronize joins by itself). Stack-overflow reports (ATLAS_RT_ERR_STACK) are not available for such
 * calls. The calls use two persistent device staging sets in turn, so at most two of them are in flight; a third waits. */
int atlas_rt_trace_join(atlas_rt_context* ctx);

/* ------------------------------------------------------------------------------------------- path tracer ---- */
typedef struct atlas_rt_camera {
    float eye[3];      /* globalData.cameraLocation */
    float origin[3];   /* near-plane upper-left corner  (PathTracingRenderer.cpp:157-162) */
    float right[3];    /* upper-right - upper-left */
    float bottom[3];   /* lower-left - upper-left */
} atlas_rt_camera;

/* Primary rays. Replaces pathtracer/rayGen.csh:25-91 as dispatched by PathTracingRenderer::Render
 * (src/engine/renderer/PathTracingRenderer.cpp:146-155): width x height x samples rays, ID = (y*w+x)*samples + s,
 * stored in the shader's 8x8-tile order. jitter = per-sample sub-pixel offset in [0,1)^2 (2 floats per sample; NULL = 0.5;
 * the shader's own value for a frame is atlas_rt_sample_jitter(sampleCount)).
 * rays_out: width*height*samples PackedRay, device memory if ATLAS_RT_DEVICE_OUTPUT. */
int atlas_rt_generate_primary_rays(atlas_rt_context* ctx, const atlas_rt_camera* camera, uint32_t width,
                                   uint32_t height, uint32_t samples, const float* jitter, void* rays_out, uint32_t flags);
/* rayGen.csh:33-34: jitter = (random(vec2(sampleCount, 0)), random(vec2(sampleCount, 1))) from the engine's hash RNG
 * (data/shader/common/random.hsh:5-33). Pure integer arithmetic, computed on the host. */
void atlas_rt_sample_jitter(int32_t sample_count, float jitter_xy[2]);

/* RaytraceMaterial — data/shader/raytracer/structures.hsh:108-141 (23 words, std430 stride 92), as RayTracingWorld fills it. */
typedef struct atlas_rt_material {
    int32_t ID;
    float baseR, baseG, baseB;
    float emissR, emissG, emissB;
    float opacity;
    float roughness, metalness, ao;
    float reflectance;
    float normalScale;
    int32_t invertUVs, twoSided, cullBackFaces, useVertexColors;
    int32_t baseColorTexture, opacityTexture, normalTexture, roughnessTexture, metalnessTexture, aoTexture;
} atlas_rt_material;

/* Single-channel 8-bit texture (an opacity map), row-major, width * height bytes, host memory. */
typedef struct atlas_rt_texture {
    uint32_t width, height;
    const uint8_t* texels;
} atlas_rt_texture;

/* Material and texture tables of a scene (copied to the device). They are what `materials[]` and the bindless opacity
 * textures are to the shaders: with them the opacity-aware traversal resolves triangles with textured opacity
 * (opacity < 0) through GetOpacity — data/shader/raytracer/surface.hsh:147-160, bvh.hsh:127,163: texture coordinates
 * interpolated from the triangle's half2 words, invertUVs, bilinear sample at mip 0 (repeat addressing; fp32 filter
 * weights), times the material opacity — and the path tracer shades with them. A material's index is
 * triangle.materialIndex + instance.materialOffset. Only opacityTexture is consulted; the other texture slots must be
 * negative (material textures are shading inputs outside this path). Without this call textured opacity counts as 1 and the
 * path tracer uses a grey default material. */
int atlas_rt_scene_set_materials(atlas_rt_context* ctx, atlas_rt_scene* scene, const atlas_rt_material* materials,
                                 uint32_t material_count, const atlas_rt_texture* textures, uint32_t texture_count);

typedef struct atlas_rt_pt_params {
    float light_dir[3];        /* direction TO the directional light: -light.N (normalised by the shader, direct.hsh:82) */
    float light_radiance[3];   /* Light.radiance */
    int32_t light_count;       /* PushConstants.lightCount: 0 or 1 */
    float sky_radiance[3];     /* the environment map as a constant colour */
    uint32_t max_bounces;      /* Uniforms.maxBounces */
    uint32_t samples_per_frame;/* Uniforms.samplesPerFrame (1 outside the real-time mode) */
} atlas_rt_pt_params;

/* One iteration of the loop in PathTracingRenderer.cpp:177-192 on device buffers: traceClosest.csh with OPACITY_CHECK
 * (HitClosestTransparency) in place over rays_in, then pathtracer/rayHit.csh:56-337 per ray — environment on a miss,
 * emissive on the first bounce, direct light from one directional light with the shadow ray traced by
 * HitAnyTransparency(INSTANCE_MASK_SHADOW) (only rays that are really cast are traced), the next direction from the
 * diffuse / GGX-VNDF specular / refraction choice with the reference's hash RNG keyed by (ray.ID, seed) in the shader's
 * draw order, Russian roulette, and either accumulation of the finished path or a compacted append of the surviving ray with
 * its half-precision payload (PackRayPayload, raytracer/common.hsh:88-98).
 * Surfaces are read from the 96-byte triangles (interpolated vertex normals, surface.hsh:64-138) and the scene's material
 * table. All buffers are DEVICE memory (flags must carry ATLAS_RT_DEVICE_INPUT | ATLAS_RT_DEVICE_OUTPUT):
 *   rays_in      count PackedRay; updated IN PLACE with the closest hits (the shader's write to the other buffer half)
 *   payload_in   count x 16 B PackedRayPayload; ignored when bounce == 0
 *   rays_out / payload_out   capacity count; receive the surviving rays compacted to the front; must not alias the inputs
 *   accum        width*height x 4 floats (rgb sum, finished-path count) at pixel ray.ID / samples_per_frame, added atomically
 * out_count (host) receives the number of surviving rays; the call synchronises the stream to read it. */
int atlas_rt_pathtrace_bounce(atlas_rt_context* ctx, const atlas_rt_scene* scene, const atlas_rt_pt_params* params,
                              float seed, uint32_t bounce, const void* rays_in, const void* payload_in, uint64_t count,
                              void* rays_out, void* payload_out, float* accum, uint32_t width, uint32_t height,
                              uint64_t* out_count, uint32_t flags);

/* The whole of PathTracingRenderer::Render's ray work for `frames` sample passes (frame f uses sampleCount =
 * first_sample_count + f for the jitter and seeds[f * (max_bounces + 1) + b] as Uniforms.seed of bounce b): ray
 * generation, then max_bounces + 1 bounces, with the ray counts kept ON THE DEVICE — the kernels read them from device
 * memory the way traceDispatch.csh + DispatchIndirect do (renderer/helper/RayTracingHelper.cpp:262-291,363), so the host
 * enqueues everything without waiting once. [slot_begin, slot_end) selects a contiguous range of rayGen's storage slots
 * (slot = tileOrderIndex * samples_per_frame + sample; slot_end == 0 means the whole frame): the unit of sharding an
 * image across GPUs, 64 * samples_per_frame slots per 8x8 tile. accum: device, width*height x 4 floats, indexed by pixel or,
 * with ATLAS_RT_ACCUM_TILE_ORDER, by tile-order index (then a slot range owns a contiguous slice). rays_traced (may be
 * NULL) receives the number of closest-hit rays traced. ATLAS_RT_RAY_BINNING adds the reference's direction binning pass
 * before every bounce after the first.
 * With frames > 1 the passes run on four (ATLAS_RT_PT_LANES: up to eight) lanes side by side: pass f on lane f mod lanes, every lane
 * with buffers and a stream of its own; lane 0 adds into accum, the other lanes into private images that are added to accum in
 * lane order before the call's work ends on the context stream. The image is deterministic and equals the one-lane image up
 * to the order of the float additions per pixel. */
int atlas_rt_pathtrace_bounces(atlas_rt_context* ctx, const atlas_rt_scene* scene, const atlas_rt_camera* camera,
                               uint32_t width, uint32_t height, const atlas_rt_pt_params* params, uint32_t frames,
                               int32_t first_sample_count, const float* seeds, uint64_t slot_begin, uint64_t slot_end,
                               float* accum, uint64_t* rays_traced, uint32_t flags);

/* The same, for one of `parts` INTERLEAVED shards of the frame: the frame's rayGen slots are cut into blocks of
 * block_pixels pixels (a multiple of 64 = whole 8x8 tiles) and shard `part` renders blocks part, part + parts, ... —
 * neighbouring blocks of the image go to different GPUs, so sky and geometry are dealt evenly (a contiguous split of a
 * landscape frame gives one GPU the sky and the other all the bounces). accum_local is COMPACT: local_pixels x 4 floats
 * holding only this shard's pixels, blocks in order; pass accum_local == NULL to query local_pixels. After one gather of
 * the shards' buffers (shard 0 first) atlas_rt_image_from_shards puts every pixel at y * width + x. */
int atlas_rt_pathtrace_bounces_interleaved(atlas_rt_context* ctx, const atlas_rt_scene* scene, const atlas_rt_camera* camera,
                                           uint32_t width, uint32_t height, const atlas_rt_pt_params* params, uint32_t frames,
                                           int32_t first_sample_count, const float* seeds, uint32_t part, uint32_t parts,
                                           uint32_t block_pixels, float* accum_local, uint64_t* local_pixels,
                                           uint64_t* rays_traced, uint32_t flags);
int atlas_rt_image_from_shards(atlas_rt_context* ctx, const float* gathered, uint32_t width, uint32_t height, uint32_t parts,
                               uint32_t block_pixels, float* image, uint32_t flags);

/* Ray binning between bounces — raytracer/tracing.hsh:18-29 (DetermineRayBin: 8x8 octahedral direction bins),
 * binningOffset.csh, binning.csh; call site RayTracingHelper.cpp:304-344 (commented out in the reference). Rays (and their
 * 16-byte payloads, if given) are moved to their bin's segment; inside a bin they keep their order (the shader's order
 * inside a bin is arbitrary). Device buffers only; rays_out must not alias rays_in. */
int atlas_rt_bin_rays(atlas_rt_context* ctx, const void* rays_in, const void* payload_in, uint64_t count, void* rays_out,
                      void* payload_out, uint32_t flags);

/* ---------------------------------------------------------------------------------------------- multi-GPU ---- */
/* One process and one context per GPU of a node; NCCL (bound at run time with dlopen, so single-GPU users never need it)
 * over NVLink / NVSwitch. The reference has no multi-GPU code: these entry points stand where its single-GPU calls stand —
 * scene assembly (RayTracingWorld::UpdateForSoftwareRayTracing, src/engine/raytracing/RayTracingWorld.cpp:267-307) and the
 * trace batch (RayTracingHelper::DispatchHitClosest, src/engine/renderer/helper/RayTracingHelper.cpp:346-364). No collective
 * sits inside the traversal: the scene is replicated, rays are split, ONE gather of 16-byte hit records follows a batch. */

/* Contiguous share of `count` rays for `rank` of `world`, aligned to `align` rays (64 keeps rayGen's 8x8 tiles whole). */
int atlas_rt_shard_range(uint64_t count, uint32_t rank, uint32_t world, uint32_t align, uint64_t* begin, uint64_t* end);

typedef struct atlas_rt_comm atlas_rt_comm;
/* 128-byte NCCL unique id: create it on one rank and hand it to the others by whatever the host application uses (a file, a
 * socket, MPI, torch.distributed). */
int atlas_rt_comm_unique_id(void* id128);
/* Collective over all ranks. The communicator works on the context's device, issues collectives on its own stream and
 * orders them against the context's stream with events. */
int atlas_rt_comm_init(atlas_rt_context* ctx, const void* id128, uint32_t rank, uint32_t world, atlas_rt_comm** out_comm);
void atlas_rt_comm_destroy(atlas_rt_comm* comm);
int atlas_rt_comm_info(const atlas_rt_comm* comm, uint32_t* rank, uint32_t* world);
/* Wait for the context's stream and the communicator's stream (joins every ATLAS_RT_ASYNC sharded call issued so far). */
int atlas_rt_comm_synchronize(atlas_rt_comm* comm);

/* A flattened tree from the rank that built it (`root`, where src is given) to every other rank, where a new object is
 * created. On root *out_bvh == src. */
int atlas_rt_bvh_broadcast(atlas_rt_comm* comm, const atlas_rt_bvh* src, uint32_t root, atlas_rt_bvh** out_bvh);

/* Scene assembly with the BLAS builds of an instanced scene split across the GPUs ("TLAS instance sets are split across
 * GPUs"): every rank passes the SAME host arrays; rank r builds meshes r, r + world, ... as one batch
 * (atlas_rt_build_blas_batch), every tree is broadcast from its owner, the TLAS (instance_aabbs: instance_count x 6) is built
 * on rank 0 and broadcast, and every rank packs its copy and assembles an identical scene, which owns all its parts
 * (atlas_rt_scene_free releases them). The result equals atlas_rt_scene_create over locally built trees bit for bit. */
int atlas_rt_build_scene_sharded(atlas_rt_comm* comm, uint32_t mesh_count, const float* const* aabbs, const float* const* tris,
                                 const uint64_t* counts, const void* instances64, const float* instance_aabbs,
                                 uint64_t instance_count, uint32_t flags, atlas_rt_scene** out_scene);

/* The complete scene of rank `root` (BLAS nodes, 48-byte and 96-byte triangles, TLAS, permuted instances, material and
 * texture tables) on every rank: the "BVH replicated to every GPU" step when only one rank has built it. On root
 * *out_scene == src; elsewhere a new scene that owns its parts. */
int atlas_rt_scene_replicate(atlas_rt_comm* comm, const atlas_rt_scene* src, uint32_t root, atlas_rt_scene** out_scene);

/* The sharded trace batch: every rank passes ITS share of the batch (atlas_rt_shard_range(total_count, rank, world, 64);
 * host or device rays) and its replica of the scene, traces the share into compact 16-byte hit records (ATLAS_RT_HITS_ONLY)
 * and the records are gathered on `root` in global ray order with one group of NCCL sends / receives: hits_out is only
 * used on root (total_count x 16 B; device memory with ATLAS_RT_DEVICE_OUTPUT). any_hit selects HitAny / HitClosest;
 * ATLAS_RT_PER_RAY_TMAX and ATLAS_RT_OPACITY are honoured. With ATLAS_RT_ASYNC the gather of one call overlaps the trace of
 * the next (local records are double buffered); atlas_rt_comm_synchronize joins. */
int atlas_rt_trace_sharded(atlas_rt_comm* comm, const atlas_rt_scene* scene, const void* rays_in, uint64_t total_count,
                           uint32_t cull_mask, float t_min, float t_max, void* hits_out, uint32_t root, uint32_t flags,
                           int any_hit);

/* ATLAS_RT_PEER_OUTPUT on atlas_rt_trace_sharded fuses the gather into the traversal: the root owns a window of device memory
 * (made by the library, opened by every other rank through CUDA IPC - the GPUs must have peer access), and every rank's
 * traversal kernel stores each 16-byte hit record at its global position in that window as the ray finishes (P2P stores over
 * NVLink / NVSwitch). No collective, copy kernel or copy engine touches the records; two counters per rank, also written
 * over NVLink and polled locally (cuStreamWaitValue32), tell the root when a rank has finished a call and the ranks when
 * the root has consumed a slot (two slots: call k + 1 traces while call k is handed on). hits_out on the root may be NULL:
 * the records then stay in the window, where atlas_rt_comm_peer_hits finds the most recent call's (valid after
 * atlas_rt_comm_synchronize and until the second-next call). The first call (and any call with a larger total_count or
 * another root) is collective and synchronises: it creates the window. Results are bit-identical to the NCCL path. */
int atlas_rt_comm_peer_hits(atlas_rt_comm* comm, const void** hits);

/* Variable-size gather of device memory to `root` (e.g. the image slices of a path-tracer frame rendered in
 * ATLAS_RT_ACCUM_TILE_ORDER): every rank sends `bytes`; root receives rank r's sizes[r] bytes at recv + offsets[r]. */
int atlas_rt_comm_gather(atlas_rt_comm* comm, const void* send, uint64_t bytes, void* recv, const uint64_t* sizes,
                         const uint64_t* offsets, uint32_t root, uint32_t flags);

#ifdef __cplusplus
}
#endif
#endif /* ATLAS_RT_H */
