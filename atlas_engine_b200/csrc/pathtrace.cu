// pathtrace.cu — primary-ray generation and the diffuse bounce of the path tracer (rayGen.csh / rayHit.csh).
#include "common.cuh"

namespace atlas {
namespace {

// rayGen.csh:25-91. One thread per (pixel, sample). Storage index: 8x8 pixel tiles are contiguous (64 rays x samples)
// so that a 32-lane warp traces neighbouring pixels; the right and bottom borders that do not fill a tile follow.
__global__ void raygen_kernel(atlas_rt_camera cam, uint32_t width, uint32_t height, uint32_t samples,
                              const float* __restrict__ jitter, float4* __restrict__ out) {
    const uint32_t x = blockIdx.x * 8u + (threadIdx.x & 7u), y = blockIdx.y * 8u + (threadIdx.x >> 3);
    const uint32_t s = blockIdx.z;
    if (x >= width || y >= height) return;
    const float jx = jitter ? jitter[2 * s] : 0.5f, jy = jitter ? jitter[2 * s + 1] : 0.5f;
    const float cu = __fdiv_rn(__fadd_rn(float(x), jx), float(width));
    const float cv = __fdiv_rn(__fadd_rn(float(y), jy), float(height));
    float d[3];
#pragma unroll
    for (int k = 0; k < 3; k++)
        d[k] = __fsub_rn(__fadd_rn(__fadd_rn(cam.origin[k], __fmul_rn(cam.right[k], cu)), __fmul_rn(cam.bottom[k], cv)), cam.eye[k]);
    const float len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
    const int id = int((y * width + x) * samples + s);   // Flatten2D(pixel, resolution) * samples + sample
    // tile-coherent storage order (rayGen.csh:53-80)
    const uint32_t perfX = width / 8u, perfY = height / 8u, overX = width % 8u, overY = height % 8u;
    const uint32_t gx = blockIdx.x, gy = blockIdx.y, local = threadIdx.x;
    uint32_t index;
    if (gx < perfX && gy < perfY) {
        index = local + (gy * perfX + gx) * 64u;
    } else if (gx >= perfX && gy < perfY) {
        const uint32_t off = perfX * perfY * 64u;
        index = y * overX + (x - perfX * 8u) + off;
    } else {
        const uint32_t off = perfX * perfY * 64u + overX * perfY * 8u;
        index = x * overY + (y - perfY * 8u) + off;   // Flatten2D(localID.yx, overlappingPixels.yx)
    }
    const size_t slot = size_t(index) * samples + s;
    out[3 * slot + 0] = make_float4(cam.eye[0], cam.eye[1], cam.eye[2], __int_as_float(id));
    out[3 * slot + 1] = make_float4(__fdiv_rn(d[0], len), __fdiv_rn(d[1], len), __fdiv_rn(d[2], len), 0.0f);
    out[3 * slot + 2] = make_float4(0.0f, __int_as_float(0), 0.0f, 0.0f);
}

}   // namespace
}   // namespace atlas

using namespace atlas;

extern "C" {

int atlas_rt_generate_primary_rays(atlas_rt_context* ctx, const atlas_rt_camera* camera, uint32_t width, uint32_t height,
                                   uint32_t samples, const float* jitter, void* rays_out, uint32_t flags) {
    if (!ctx || !camera || !rays_out || !width || !height || !samples) return fail(ctx, ATLAS_RT_ERR_INVALID, "bad argument");
    ATLAS_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t count = uint64_t(width) * height * samples;
    if (count > 0x7fffffffull) return fail(ctx, ATLAS_RT_ERR_UNSUPPORTED, "more than 2^31-1 primary rays");
    const bool devOut = flags & ATLAS_RT_DEVICE_OUTPUT;
    float4* dOut = static_cast<float4*>(rays_out);
    float4* tmp = nullptr;
    float* dJit = nullptr;
    if (!devOut) { ATLAS_CUDA(ctx, dev_alloc(ctx, &tmp, count * 3)); dOut = tmp; }
    if (jitter) {
        ATLAS_CUDA(ctx, dev_alloc(ctx, &dJit, size_t(samples) * 2));
        ATLAS_CUDA(ctx, cudaMemcpyAsync(dJit, jitter, size_t(samples) * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    const dim3 grid((width + 7) / 8, (height + 7) / 8, samples);
    raygen_kernel<<<grid, 64, 0, ctx->stream>>>(*camera, width, height, samples, dJit, dOut);
    ATLAS_LAUNCH_CHECK(ctx);
    if (!devOut) ATLAS_CUDA(ctx, copy_out(ctx, rays_out, dOut, count * 48, false));
    dev_free(ctx, tmp);
    dev_free(ctx, dJit);
    if (!(flags & ATLAS_RT_ASYNC) || jitter) ATLAS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ATLAS_RT_OK;
}

int atlas_rt_pathtrace_bounce(atlas_rt_context* ctx, const atlas_rt_scene* scene, const atlas_rt_bounce_params* params,
                              const void* rays_in, const void* payload_in, uint64_t count, void* rays_out,
                              void* payload_out, float* accum, uint64_t* out_count, uint32_t flags) {
    (void)scene; (void)params; (void)rays_in; (void)payload_in; (void)count; (void)rays_out; (void)payload_out; (void)accum;
    (void)out_count; (void)flags;
    return fail(ctx, ATLAS_RT_ERR_UNSUPPORTED, "atlas_rt_pathtrace_bounce: not implemented yet");
}

}
