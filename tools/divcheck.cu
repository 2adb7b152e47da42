// Development aid: brute-force check of "exact division by a precomputed correctly-rounded reciprocal" against
// __fdiv_rn, as used by the traversal's slab test (trace.cu). Build: nvcc -arch=sm_100a -fmad=false -O3 divcheck.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t splitmix(uint64_t& s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void check(uint64_t seed, int iters, int mode, unsigned long long* out) {
    uint64_t s = seed + (blockIdx.x * uint64_t(blockDim.x) + threadIdx.x) * 0x1234567ull;
    unsigned long long bad1 = 0, bad2 = 0;
    for (int i = 0; i < iters; i++) {
        const uint64_t r = splitmix(s);
        uint32_t xb = uint32_t(r), db = uint32_t(r >> 32);
        // the admitted domain of trace.cu's fast path: d in [2^-64, 2^30], x == 0 or in [2^-64, 2^61]; mode 1: adversarial
        // mantissas (all ones / near powers of two); mode 2: OUTSIDE the domain — x down to the subnormals, d up to 2^64 — to
        // show that the guards are needed (mismatches expected there)
        uint32_t xe = 127 - 64 + (xb >> 23) % 126, de = 127 - 64 + (db >> 23) % 95;
        if (mode == 2) { xe = (xb >> 23) % 64; de = 127 - 10 + (db >> 23) % 75; }
        uint32_t xm = xb & 0x7fffff, dm = db & 0x7fffff;
        if (mode == 1) { const uint64_t q = splitmix(s); if (q & 1) dm |= 0x7ffff0; if (q & 2) xm |= 0x7fff00; if (q & 4) dm &= 0xf; if (q & 8) xm &= 0xff; }
        float x = __uint_as_float((xb & 0x80000000u) | (xe << 23) | xm);
        if (mode != 2 && (r & 0xff0000u) == 0u) x = (xb & 0x80000000u) ? -0.0f : 0.0f;   // exact zero numerators (origin on a box plane)
        const float d = __uint_as_float((db & 0x80000000u) | (de << 23) | dm);
        const float ref = __fdiv_rn(x, d);
        const float rc = __frcp_rn(d);
        const float q0 = __fmul_rn(x, rc);
        const float e0 = __fmaf_rn(-d, q0, x);
        const float q1 = __fmaf_rn(e0, rc, q0);
        const float e1 = __fmaf_rn(-d, q1, x);
        const float q2 = __fmaf_rn(e1, rc, q1);
        // the sign of a zero quotient is not compared: x = -0 gives +0 here and -0 from the division, and no comparison of
        // the slab test (nor its min / max) can tell the two apart
        bad1 += (__float_as_uint(q1) != __float_as_uint(ref)) && !(q1 == 0.0f && ref == 0.0f);
        bad2 += (__float_as_uint(q2) != __float_as_uint(ref)) && !(q2 == 0.0f && ref == 0.0f);
    }
    atomicAdd(&out[0], bad1);
    atomicAdd(&out[1], bad2);
}

int main() {
    unsigned long long* d;
    cudaMalloc(&d, 16);
    for (int mode = 0; mode < 3; mode++) {
        cudaMemset(d, 0, 16);
        const int blocks = 148 * 16, threads = 256, iters = 1 << 14, rounds = 16;
        for (int r = 0; r < rounds; r++) check<<<blocks, threads>>>(0xABCDEFull * (r + 1) + mode, iters, mode, d);
        cudaDeviceSynchronize();
        unsigned long long h[2];
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        if (mode == 2) printf("(outside the admitted domain: subnormal quotients / residuals) ");
        printf("mode %d: samples %.3e  one-step mismatches %llu  two-step mismatches %llu  (%s)\n", mode,
               double(blocks) * threads * iters * rounds, h[0], h[1], cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
