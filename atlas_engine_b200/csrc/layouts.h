// layouts.h — byte layouts shared with the engine, asserted at compile time. Names follow the reference:
//   Volume::AABB / BVHNode          src/engine/volume/AABB.h:102-103, BVH.h:14-24
//   GPUBVHNode / GPUBVHTriangle / GPUBVHInstance   src/engine/raytracing/RTStructures.h:23-27,85-103
//   PackedRay                        data/shader/raytracer/structures.hsh:9-13
#pragma once
#include <stdint.h>

namespace atlas {

struct HostAABB { float min[3]; float max[3]; };
struct HostBVHNode { HostAABB leftAABB; HostAABB rightAABB; int32_t leftPtr; int32_t rightPtr; };
struct GPUBVHNode { HostAABB leftAABB; HostAABB rightAABB; int32_t leftPtr; int32_t rightPtr; int32_t padding0; int32_t padding1; };
struct GPUBVHTriangle { float v0[4]; float v1[4]; float v2[4]; };
struct GPUBVHInstance { float inverseMatrix[12]; int32_t meshOffset; int32_t materialOffset; int32_t nextInstance; uint32_t mask; };
struct PackedRay { float origin[4]; float direction[4]; float hit[4]; };

static_assert(sizeof(HostAABB) == 24, "AABB");
static_assert(sizeof(HostBVHNode) == 56, "BVHNode");
static_assert(sizeof(GPUBVHNode) == 64, "GPUBVHNode");
static_assert(sizeof(GPUBVHTriangle) == 48, "GPUBVHTriangle");
static_assert(sizeof(GPUBVHInstance) == 64, "GPUBVHInstance");
static_assert(sizeof(PackedRay) == 48, "PackedRay");

}   // namespace atlas
