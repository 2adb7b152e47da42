// build_common.cuh — device helpers shared by the level-synchronous big-node kernels and the shared-memory subtree
// kernel of the builder (build.cu). Everything here restates a piece of src/engine/volume/BVH.cpp with the exact
// operation order of the reference and no FMA contraction; reductions are done on the order-preserving integer image
// of floats (min/max) or on integers (counts), so they are exact whatever the thread interleaving.
#pragma once

#include "common.cuh"

namespace atlas {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kOrdEmptyLo = 0x7f7fffff;    // ord_from_float(+FLT_MAX)
constexpr int kOrdEmptyHi = int(0x80800000);   // ord_from_float(-FLT_MAX) = 0xff7fffff ^ 0x7fffffff

// max(binCount / (depth + 1), 16) — BVH.cpp:447 (256 for a BLAS, 64 for a TLAS).
__host__ __device__ __forceinline__ uint32_t bins_at_depth(uint32_t budget, uint32_t depth) {
    const uint32_t b = budget / (depth + 1u);
    return b > 16u ? b : 16u;
}

// Per-axis binning constants of a node — BVH.cpp:464-472.
struct AxisBins {
    float start, width, inv;
    bool active;   // false when fabsf(stop - start) < 1e-3f (axis skipped)
};

__device__ __forceinline__ AxisBins axis_bins(float start, float stop, uint32_t bins) {
    AxisBins a;
    a.start = start;
    a.active = !(fabsf(__fsub_rn(stop, start)) < 1e-3f);
    a.width = __fdiv_rn(__fsub_rn(stop, start), float(bins));
    a.inv = __fdiv_rn(1.0f, a.width);
    return a;
}

// uint32_t(glm::clamp((value - start) * invBinSize, 0.0f, float(bins) - 1.0f)) — BVH.cpp:477, :545, :590-593.
__device__ __forceinline__ uint32_t bin_of(float value, float start, float inv, uint32_t bins) {
    const float f = gl_clamp(__fmul_rn(__fsub_rn(value, start), inv), 0.0f, __fsub_rn(float(bins), 1.0f));
    return __float2uint_rz(f);
}

// Centre used for object binning: 0.5f * (min + max) — BVH.cpp:475, :542.
__device__ __forceinline__ float bin_centre(float lo, float hi) { return __fmul_rn(0.5f, __fadd_rn(lo, hi)); }
// Centre used by the median split: (max - min) * 0.5f + min — BVH.cpp:814.
__device__ __forceinline__ float median_centre(float lo, float hi) { return __fadd_rn(__fmul_rn(__fsub_rn(hi, lo), 0.5f), lo); }

// Longest axis and cutoff of PerformMedianSplit — BVH.cpp:805-811.
__device__ __forceinline__ void median_plane(const float lo[3], const float hi[3], int& axis, float& cutoff) {
    const float dim[3] = {__fsub_rn(hi[0], lo[0]), __fsub_rn(hi[1], lo[1]), __fsub_rn(hi[2], lo[2])};
    axis = 0;
    axis = dim[1] > dim[axis] ? 1 : axis;
    axis = dim[2] > dim[axis] ? 2 : axis;
    cutoff = __fadd_rn(lo[axis], __fdiv_rn(dim[axis], 2.0f));
}

// ------------------------------------------------------------------------------------------------ bin records
// One bin = 8 ints: ord(lo.xyz), ord(hi.xyz), enter, exit. For the object split enter == exit == primitiveCount.
constexpr int kBinWords = 8;

__device__ __forceinline__ void bin_init(int* b) {
    b[0] = b[1] = b[2] = kOrdEmptyLo;
    b[3] = b[4] = b[5] = kOrdEmptyHi;
    b[6] = 0;
    b[7] = 0;
}

// Box in ordered-int form, used inside warp scans.
struct OBox {
    int lo[3], hi[3];
};
__device__ __forceinline__ OBox obox_empty() {
    OBox b;
    b.lo[0] = b.lo[1] = b.lo[2] = kOrdEmptyLo;
    b.hi[0] = b.hi[1] = b.hi[2] = kOrdEmptyHi;
    return b;
}
__device__ __forceinline__ void obox_grow(OBox& a, const OBox& b) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
        a.lo[k] = min(a.lo[k], b.lo[k]);
        a.hi[k] = max(a.hi[k], b.hi[k]);
    }
}
__device__ __forceinline__ float obox_area(const OBox& b) {
    const float lo[3] = {float_from_ord(b.lo[0]), float_from_ord(b.lo[1]), float_from_ord(b.lo[2])};
    const float hi[3] = {float_from_ord(b.hi[0]), float_from_ord(b.hi[1]), float_from_ord(b.hi[2])};
    return surface_area(lo, hi);
}
__device__ __forceinline__ Box3 obox_to_box(const OBox& b) {
    Box3 r;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        r.lo[k] = float_from_ord(b.lo[k]);
        r.hi[k] = float_from_ord(b.hi[k]);
    }
    return r;
}

// Result of a sweep over one or more axes: the Split of BVH.h:80-90 minus the boxes (recomputed by split_boxes()).
struct BestSplit {
    float cost;
    int axis;
    uint32_t bin;
};
__device__ __forceinline__ BestSplit best_none() { return BestSplit{kFltMax, -1, 0u}; }

// A group of G consecutive lanes of a warp (G = 32: the whole warp, G = 16: a half warp) working on one node. All
// cross-lane operations take the group's own member mask, so the two halves of a warp can build two different nodes
// in the same instruction stream.
template <int G>
struct LaneGroup {
    static constexpr unsigned kBits = (G == 32) ? 0xffffffffu : ((1u << (G & 31)) - 1u);
    unsigned shift, mask, lane;
    __device__ __forceinline__ LaneGroup() {
        const unsigned l = threadIdx.x & 31u;
        shift = l & ~(unsigned(G) - 1u);
        mask = kBits << shift;
        lane = l & (unsigned(G) - 1u);
    }
    __device__ __forceinline__ unsigned ballot(bool p) const { return (__ballot_sync(mask, p) >> shift) & kBits; }
    __device__ __forceinline__ void sync() const { __syncwarp(mask); }
    __device__ __forceinline__ int up(int v, int off) const { return __shfl_up_sync(mask, v, off, G); }
    __device__ __forceinline__ unsigned up(unsigned v, int off) const { return __shfl_up_sync(mask, v, off, G); }
    __device__ __forceinline__ int down(int v, int off) const { return __shfl_down_sync(mask, v, off, G); }
    __device__ __forceinline__ int bcast(int v, int src) const { return __shfl_sync(mask, v, src, G); }
    __device__ __forceinline__ unsigned bcast(unsigned v, int src) const { return __shfl_sync(mask, v, src, G); }
    __device__ __forceinline__ float bxor(float v, int off) const { return __shfl_xor_sync(mask, v, off, G); }
    __device__ __forceinline__ unsigned bxor(unsigned v, int off) const { return __shfl_xor_sync(mask, v, off, G); }
    // Group reductions. The whole-warp group uses REDUX with the constant full mask; a half-warp group uses a shuffle
    // butterfly, because REDUX with a run-time mask compiles to a MATCH.ANY loop and branching to two constant-mask
    // REDUX instructions makes the two halves diverge (measured: slower than the butterfly).
    __device__ __forceinline__ int rmin(int v) const {
        if constexpr (G == 32) {
            return __reduce_min_sync(0xffffffffu, v);
        } else {
#pragma unroll
            for (int off = G / 2; off > 0; off >>= 1) v = min(v, __shfl_xor_sync(mask, v, off, G));
            return v;
        }
    }
    __device__ __forceinline__ int rmax(int v) const {
        if constexpr (G == 32) {
            return __reduce_max_sync(0xffffffffu, v);
        } else {
#pragma unroll
            for (int off = G / 2; off > 0; off >>= 1) v = max(v, __shfl_xor_sync(mask, v, off, G));
            return v;
        }
    }
    __device__ __forceinline__ unsigned radd(unsigned v) const {
        if constexpr (G == 32) {
            return __reduce_add_sync(0xffffffffu, v);
        } else {
#pragma unroll
            for (int off = G / 2; off > 0; off >>= 1) v += __shfl_xor_sync(mask, v, off, G);
            return v;
        }
    }
    __device__ __forceinline__ void reduce(OBox& b) const {
#pragma unroll
        for (int w = 0; w < 3; w++) { b.lo[w] = rmin(b.lo[w]); b.hi[w] = rmax(b.hi[w]); }
    }
};

// Group-cooperative SAH sweep over the bins of ONE axis (BVH.cpp:484-522 object, :621-661 spatial). `bins` = nb
// records of kBinWords ints in shared or global memory, `sfx` = scratch for nb suffix boxes (6 ints each) in the same
// kind of memory, private to this group. All G lanes must call; every lane returns the same updated `best`.
// Candidate j in [1, nb): left = bins[0..j-1], right = bins[j..nb-1], nLeft = sum enter[0..j-1],
// nRight = total - sum exit[0..j-1]; skipped when either is 0; cost = SA(left)*float(nLeft) + SA(right)*float(nRight);
// strictly smaller cost wins, i.e. the lowest (axis, j) among equal costs.
// Returns true when this axis improved `best`. `stride` = ints between two bin records.
template <int G, typename BinPtr>
__device__ inline bool group_sweep_axis(const LaneGroup<G>& g, BinPtr bins, uint32_t nb, int* sfx, uint32_t total, int axis,
                                        BestSplit& best, int stride = kBinWords) {
    const uint32_t lane = g.lane;
    const uint32_t chunks = (nb + G - 1u) / G;
    // ---- backward pass: sfx[k] = union of bins[k..nb-1]
    OBox carry = obox_empty();
    for (int c = int(chunks) - 1; c >= 0; c--) {
        const uint32_t k = uint32_t(c) * G + lane;
        OBox b = obox_empty();
        if (k < nb) {
#pragma unroll
            for (int w = 0; w < 3; w++) { b.lo[w] = bins[k * stride + w]; b.hi[w] = bins[k * stride + 3 + w]; }
        }
#pragma unroll
        for (int off = 1; off < G; off <<= 1) {
            OBox o;
#pragma unroll
            for (int w = 0; w < 3; w++) { o.lo[w] = g.down(b.lo[w], off); o.hi[w] = g.down(b.hi[w], off); }
            if (lane + off < uint32_t(G)) obox_grow(b, o);
        }
        obox_grow(b, carry);
        if (k < nb) {
#pragma unroll
            for (int w = 0; w < 3; w++) { sfx[k * 6 + w] = b.lo[w]; sfx[k * 6 + 3 + w] = b.hi[w]; }
        }
#pragma unroll
        for (int w = 0; w < 3; w++) { carry.lo[w] = g.bcast(b.lo[w], 0); carry.hi[w] = g.bcast(b.hi[w], 0); }
    }
    g.sync();
    // ---- forward pass
    OBox pcarry = obox_empty();
    uint32_t ecarry = 0, xcarry = 0;
    float myCost = kFltMax;
    uint32_t myBin = 0xffffffffu;
    for (uint32_t c = 0; c < chunks; c++) {
        const uint32_t k = c * G + lane;
        OBox b = obox_empty();
        uint32_t en = 0, ex = 0;
        if (k < nb) {
#pragma unroll
            for (int w = 0; w < 3; w++) { b.lo[w] = bins[k * stride + w]; b.hi[w] = bins[k * stride + 3 + w]; }
            en = uint32_t(bins[k * stride + 6]);
            ex = uint32_t(bins[k * stride + 7]);
        }
#pragma unroll
        for (int off = 1; off < G; off <<= 1) {
            OBox o;
#pragma unroll
            for (int w = 0; w < 3; w++) { o.lo[w] = g.up(b.lo[w], off); o.hi[w] = g.up(b.hi[w], off); }
            const uint32_t oe = g.up(en, off), ox = g.up(ex, off);
            if (lane >= uint32_t(off)) { obox_grow(b, o); en += oe; ex += ox; }
        }
        obox_grow(b, pcarry);
        en += ecarry;
        ex += xcarry;
        const uint32_t j = k + 1u;   // split after bin k
        if (j < nb) {
            const uint32_t nLeft = en, nRight = total - ex;
            if (nLeft != 0u && nRight != 0u) {
                OBox r;
#pragma unroll
                for (int w = 0; w < 3; w++) { r.lo[w] = sfx[j * 6 + w]; r.hi[w] = sfx[j * 6 + 3 + w]; }
                const float cost = __fadd_rn(__fmul_rn(obox_area(b), __uint2float_rn(nLeft)),
                                             __fmul_rn(obox_area(r), __uint2float_rn(nRight)));
                if (cost < myCost) { myCost = cost; myBin = j; }   // ascending j within a lane: first wins ties
            }
        }
#pragma unroll
        for (int w = 0; w < 3; w++) { pcarry.lo[w] = g.bcast(b.lo[w], G - 1); pcarry.hi[w] = g.bcast(b.hi[w], G - 1); }
        ecarry = g.bcast(en, G - 1);
        xcarry = g.bcast(ex, G - 1);
    }
    g.sync();
    // ---- group arg-min on (cost, j); NaN costs never satisfy cost < x and so never win (as in the reference)
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) {
        const float oc = g.bxor(myCost, off);
        const uint32_t ob = g.bxor(myBin, off);
        if (oc < myCost || (oc == myCost && ob < myBin)) { myCost = oc; myBin = ob; }
    }
    if (myBin != 0xffffffffu && myCost < best.cost) {
        best.cost = myCost;
        best.axis = axis;
        best.bin = myBin;
        return true;
    }
    return false;
}

// Shared-memory subtree kernel: one bin = 7 ints (ord lo.xyz, ord hi.xyz, primitiveCount). The odd stride keeps the
// lane-per-bin loads of the sweep and the atomics of neighbouring bins in different banks.
constexpr int kSubBinWords = 7;

__device__ __forceinline__ void sub_bin_init(int* b) {
    b[0] = b[1] = b[2] = kOrdEmptyLo;
    b[3] = b[4] = b[5] = kOrdEmptyHi;
    b[6] = 0;
}

// Object-split sweep of ONE axis whose nb bins fit the group (nb <= G): every bin lives in a lane, prefix and suffix
// boxes are scanned with shuffles only, and the boxes / left count of the winning candidate are handed back from the
// lanes that hold them. Same candidates, costs and tie rule as group_sweep_axis.
// REV (whole warp, nb <= 16): lanes 0..15 hold bins 0..15, lanes 16..31 hold the bins in reverse order, so ONE
// width-16 up-scan produces the prefix boxes in the low half and the suffix boxes in the high half.
template <int G, bool REV>
__device__ __forceinline__ bool group_sweep_single(const LaneGroup<G>& g, const int* bins, uint32_t nb, uint32_t total, int axis,
                                                   BestSplit& best, OBox& outL, OBox& outR, uint32_t& outLeft) {
    static_assert(!REV || G == 32, "the reversed layout needs a whole warp");
    const uint32_t lane = g.lane;
    OBox pre = obox_empty();
    OBox right;          // suffix box of candidate j = lane + 1
    OBox suf;            // !REV: suffix box starting at this lane's bin
    uint32_t en = 0;
    if constexpr (REV) {
        const uint32_t m = lane & 15u;
        const uint32_t k = lane < 16u ? m : nb - 1u - m;
        if (m < nb) {
            const int* rec = bins + k * kSubBinWords;
#pragma unroll
            for (int w = 0; w < 3; w++) { pre.lo[w] = rec[w]; pre.hi[w] = rec[3 + w]; }
            en = uint32_t(rec[6]);
        }
#pragma unroll
        for (int off = 1; off < 16; off <<= 1) {
            OBox o;
#pragma unroll
            for (int w = 0; w < 3; w++) { o.lo[w] = __shfl_up_sync(kFullMask, pre.lo[w], off, 16); o.hi[w] = __shfl_up_sync(kFullMask, pre.hi[w], off, 16); }
            const uint32_t oe = __shfl_up_sync(kFullMask, en, off, 16);
            if (m >= uint32_t(off)) { obox_grow(pre, o); en += oe; }
        }
        // suffix over bins [j, nb) sits in lane 16 + (nb - 1 - j)
        const int src = 16 + int((nb - 2u - lane) & 15u);
#pragma unroll
        for (int w = 0; w < 3; w++) { right.lo[w] = __shfl_sync(kFullMask, pre.lo[w], src); right.hi[w] = __shfl_sync(kFullMask, pre.hi[w], src); }
        suf = pre;
    } else {
        if (lane < nb) {
            const int* rec = bins + lane * kSubBinWords;
#pragma unroll
            for (int w = 0; w < 3; w++) { pre.lo[w] = rec[w]; pre.hi[w] = rec[3 + w]; }
            en = uint32_t(rec[6]);
        }
        suf = pre;
#pragma unroll
        for (int off = 1; off < G; off <<= 1) {
            OBox o, p;
#pragma unroll
            for (int w = 0; w < 3; w++) {
                o.lo[w] = g.up(pre.lo[w], off); o.hi[w] = g.up(pre.hi[w], off);
                p.lo[w] = g.down(suf.lo[w], off); p.hi[w] = g.down(suf.hi[w], off);
            }
            const uint32_t oe = g.up(en, off);
            if (lane >= uint32_t(off)) { obox_grow(pre, o); en += oe; }
            if (lane + off < uint32_t(G)) obox_grow(suf, p);
        }
#pragma unroll
        for (int w = 0; w < 3; w++) { right.lo[w] = g.down(suf.lo[w], 1); right.hi[w] = g.down(suf.hi[w], 1); }
    }
    float myCost = kFltMax;
    uint32_t myBin = 0xffffffffu;
    const uint32_t j = lane + 1u;   // split after bin `lane`
    if (j < nb && (!REV || lane < 16u)) {
        const uint32_t nLeft = en, nRight = total - en;
        if (nLeft != 0u && nRight != 0u) {
            const float cost = __fadd_rn(__fmul_rn(obox_area(pre), __uint2float_rn(nLeft)), __fmul_rn(obox_area(right), __uint2float_rn(nRight)));
            if (cost < kFltMax) { myCost = cost; myBin = j; }   // NaN / inf never beat the initial best (BVH.cpp:519)
        }
    }
    if constexpr (G == 32) {
        // costs are sums of non-negative products, so their ordered-int images compare like the floats; one REDUX finds
        // the minimum and the lowest lane holding it is the lowest bin
        const int key = ord_from_float(myCost);
        const int m = __reduce_min_sync(kFullMask, key);
        const unsigned who = __ballot_sync(kFullMask, key == m && myBin != 0xffffffffu);
        if (who == 0u) return false;
        myCost = float_from_ord(m);
        myBin = uint32_t(__ffs(int(who)));   // lane + 1
    } else {
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) {
            const float oc = g.bxor(myCost, off);
            const uint32_t ob = g.bxor(myBin, off);
            if (oc < myCost || (oc == myCost && ob < myBin)) { myCost = oc; myBin = ob; }   // lanes without a candidate hold (FLT_MAX, ~0)
        }
    }
    if (myBin == 0xffffffffu || !(myCost < best.cost)) return false;
    best.cost = myCost;
    best.axis = axis;
    best.bin = myBin;
    const int lsrc = int(myBin) - 1;
    const int rsrc = REV ? 16 + int(nb - 1u - myBin) : int(myBin);
#pragma unroll
    for (int w = 0; w < 3; w++) {
        outL.lo[w] = g.bcast(pre.lo[w], lsrc); outL.hi[w] = g.bcast(pre.hi[w], lsrc);
        outR.lo[w] = g.bcast(suf.lo[w], rsrc); outR.hi[w] = g.bcast(suf.hi[w], rsrc);
    }
    outLeft = g.bcast(en, lsrc);
    return true;
}

// The three axes of a node swept together by one warp, nb <= 16 bins per axis: the reversed layout of
// group_sweep_single<32, true> for each axis, with the scan steps of the three axes issued side by side so that their
// shuffle and min/max latencies overlap. `bins3` = [3][16] records of kSubBinWords ints; `active` = bit a set when axis
// a takes part (BVH.cpp:466). Same candidates, costs and tie rule (lowest axis, then lowest bin) as three calls in a row.
__device__ __forceinline__ void warp_sweep3_rev(const int* bins3, uint32_t nb, uint32_t total, uint32_t active, BestSplit& best,
                                                OBox& outL, OBox& outR, uint32_t& outLeft) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t m = lane & 15u;
    const uint32_t k = lane < 16u ? m : nb - 1u - m;
    OBox pre[3];
    uint32_t en[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        pre[a] = obox_empty();
        en[a] = 0;
        if (m < nb && ((active >> a) & 1u)) {
            const int* rec = bins3 + (a * 16 + k) * kSubBinWords;
#pragma unroll
            for (int w = 0; w < 3; w++) { pre[a].lo[w] = rec[w]; pre[a].hi[w] = rec[3 + w]; }
            en[a] = uint32_t(rec[6]);
        }
    }
#pragma unroll
    for (int off = 1; off < 16; off <<= 1) {
        OBox o[3];
        uint32_t oe[3];
#pragma unroll
        for (int a = 0; a < 3; a++) {
#pragma unroll
            for (int w = 0; w < 3; w++) { o[a].lo[w] = __shfl_up_sync(kFullMask, pre[a].lo[w], off, 16); o[a].hi[w] = __shfl_up_sync(kFullMask, pre[a].hi[w], off, 16); }
            oe[a] = __shfl_up_sync(kFullMask, en[a], off, 16);
        }
        if (m >= uint32_t(off)) {
#pragma unroll
            for (int a = 0; a < 3; a++) { obox_grow(pre[a], o[a]); en[a] += oe[a]; }
        }
    }
    // suffix over bins [j, nb) of an axis sits in lane 16 + (nb - 1 - j)
    const int src = 16 + int((nb - 2u - lane) & 15u);
    const uint32_t j = lane + 1u;
    const bool cand = lane < 16u && j < nb;
    int key[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        OBox right;
#pragma unroll
        for (int w = 0; w < 3; w++) { right.lo[w] = __shfl_sync(kFullMask, pre[a].lo[w], src); right.hi[w] = __shfl_sync(kFullMask, pre[a].hi[w], src); }
        float cost = kFltMax;
        const uint32_t nLeft = en[a], nRight = total - en[a];
        if (cand && nLeft != 0u && nRight != 0u) {
            const float c = __fadd_rn(__fmul_rn(obox_area(pre[a]), __uint2float_rn(nLeft)), __fmul_rn(obox_area(right), __uint2float_rn(nRight)));
            if (c < kFltMax) cost = c;   // NaN / inf never beat the initial best (BVH.cpp:519)
        }
        key[a] = ord_from_float(cost);   // costs are sums of non-negative products: ordered-int images compare like the floats
    }
    const int kEmpty = ord_from_float(kFltMax);
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const int mn = __reduce_min_sync(kFullMask, key[a]);
        if (mn == kEmpty) continue;   // warp-uniform
        const float c = float_from_ord(mn);
        if (!(c < best.cost)) continue;
        const unsigned who = __ballot_sync(kFullMask, key[a] == mn);
        const uint32_t bin = uint32_t(__ffs(int(who)));   // lowest lane holding the minimum = lowest bin; j = lane + 1
        best.cost = c;
        best.axis = a;
        best.bin = bin;
        const int lsrc = int(bin) - 1, rsrc = 16 + int(nb - 1u - bin);
#pragma unroll
        for (int w = 0; w < 3; w++) {
            outL.lo[w] = __shfl_sync(kFullMask, pre[a].lo[w], lsrc); outL.hi[w] = __shfl_sync(kFullMask, pre[a].hi[w], lsrc);
            outR.lo[w] = __shfl_sync(kFullMask, pre[a].lo[w], rsrc); outR.hi[w] = __shfl_sync(kFullMask, pre[a].hi[w], rsrc);
        }
        outLeft = __shfl_sync(kFullMask, en[a], lsrc);
    }
}

__device__ inline void warp_sweep_axis(const int* bins, uint32_t nb, int* sfx, uint32_t total, int axis, BestSplit& best,
                                       int stride = kBinWords) {
    const LaneGroup<32> g;
    group_sweep_axis<32>(g, bins, nb, sfx, total, axis, best, stride);
}

// leftAABB / rightAABB / primitivesLeft of the chosen split, recomputed from the bins of the winning axis.
__device__ inline void warp_split_boxes(const int* bins, uint32_t nb, uint32_t j, OBox& left, OBox& right, uint32_t& nLeft,
                                        uint32_t& nExitLeft, int stride = kBinWords) {
    const uint32_t lane = threadIdx.x & 31u;
    left = obox_empty();
    right = obox_empty();
    uint32_t en = 0, ex = 0;
    for (uint32_t k = lane; k < nb; k += 32u) {
        OBox b;
#pragma unroll
        for (int w = 0; w < 3; w++) { b.lo[w] = bins[k * stride + w]; b.hi[w] = bins[k * stride + 3 + w]; }
        if (k < j) {
            obox_grow(left, b);
            en += uint32_t(bins[k * stride + 6]);
            ex += uint32_t(bins[k * stride + 7]);
        } else {
            obox_grow(right, b);
        }
    }
#pragma unroll
    for (int w = 0; w < 3; w++) {
        left.lo[w] = __reduce_min_sync(kFullMask, left.lo[w]);
        left.hi[w] = __reduce_max_sync(kFullMask, left.hi[w]);
        right.lo[w] = __reduce_min_sync(kFullMask, right.lo[w]);
        right.hi[w] = __reduce_max_sync(kFullMask, right.hi[w]);
    }
    nLeft = __reduce_add_sync(kFullMask, en);
    nExitLeft = __reduce_add_sync(kFullMask, ex);
}

// ------------------------------------------------------------------------------------------- std::sort emulation
// PerformMedianSplit's fallback (BVH.cpp:832-836) calls std::sort with the comparator extent(a) < extent(b). The
// resulting order of equal keys is whatever libstdc++'s introsort produces, so it is restated step for step
// (bits/stl_algo.h: __introsort_loop, __unguarded_partition_pivot, __move_median_to_first, __final_insertion_sort,
// and bits/stl_heap.h for the depth-limit fallback). Single thread; refs are moved as whole (lo4, hi4) pairs.
struct RefArray {
    float4* lo;   // lo.xyz, bits(idx)
    float4* hi;   // hi.xyz, -
    int axis;
    __device__ __forceinline__ float key_of(const float4& l, const float4& h) const {
        const float a = axis == 0 ? l.x : (axis == 1 ? l.y : l.z);
        const float b = axis == 0 ? h.x : (axis == 1 ? h.y : h.z);
        return __fsub_rn(b, a);
    }
    __device__ __forceinline__ float key(int i) const { return key_of(lo[i], hi[i]); }
    __device__ __forceinline__ void move(int dst, int src) { lo[dst] = lo[src]; hi[dst] = hi[src]; }
    __device__ __forceinline__ void swap(int a, int b) {
        const float4 l = lo[a], h = hi[a];
        lo[a] = lo[b]; hi[a] = hi[b];
        lo[b] = l; hi[b] = h;
    }
};

__device__ inline void ss_unguarded_linear_insert(RefArray& r, int last) {
    const float4 vl = r.lo[last], vh = r.hi[last];
    const float vk = r.key_of(vl, vh);
    int next = last - 1;
    while (vk < r.key(next)) {
        r.move(last, next);
        last = next;
        --next;
    }
    r.lo[last] = vl;
    r.hi[last] = vh;
}

__device__ inline void ss_insertion_sort(RefArray& r, int first, int last) {
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        if (r.key(i) < r.key(first)) {
            const float4 vl = r.lo[i], vh = r.hi[i];
            for (int k = i; k > first; --k) r.move(k, k - 1);   // move_backward(first, i, i + 1)
            r.lo[first] = vl;
            r.hi[first] = vh;
        } else {
            ss_unguarded_linear_insert(r, i);
        }
    }
}

__device__ inline void ss_adjust_heap(RefArray& r, int first, int hole, int len, float4 vl, float4 vh) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (r.key(first + child) < r.key(first + child - 1)) child--;
        r.move(first + hole, first + child);
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        r.move(first + hole, first + child - 1);
        hole = child - 1;
    }
    const float vk = r.key_of(vl, vh);   // __push_heap
    int parent = (hole - 1) / 2;
    while (hole > top && r.key(first + parent) < vk) {
        r.move(first + hole, first + parent);
        hole = parent;
        parent = (hole - 1) / 2;
    }
    r.lo[first + hole] = vl;
    r.hi[first + hole] = vh;
}

__device__ inline void ss_heap_sort(RefArray& r, int first, int last) {   // __partial_sort(first, last, last)
    const int len = last - first;
    if (len >= 2) {
        int parent = (len - 2) / 2;
        while (true) {
            ss_adjust_heap(r, first, parent, len, r.lo[first + parent], r.hi[first + parent]);
            if (parent == 0) break;
            parent--;
        }
    }
    int end = last;
    while (end - first > 1) {
        --end;
        const float4 vl = r.lo[end], vh = r.hi[end];
        r.move(end, first);
        ss_adjust_heap(r, first, 0, end - first, vl, vh);
    }
}

__device__ inline void std_sort_refs(RefArray r, int n) {
    if (n <= 1) return;
    constexpr int kThreshold = 16;
    // __introsort_loop with an explicit stack for the recursion on the right part
    int stackFirst[64], stackLast[64], stackDepth[64];
    int sp = 0;
    int depthLimit = 2 * (31 - __clz(n));
    stackFirst[sp] = 0; stackLast[sp] = n; stackDepth[sp] = depthLimit; sp++;
    while (sp > 0) {
        sp--;
        int first = stackFirst[sp], last = stackLast[sp], depth = stackDepth[sp];
        while (last - first > kThreshold) {
            if (depth == 0) {
                ss_heap_sort(r, first, last);
                break;
            }
            --depth;
            // __unguarded_partition_pivot
            const int mid = first + (last - first) / 2;
            {   // __move_median_to_first(result = first, a = first + 1, b = mid, c = last - 1)
                const int a = first + 1, b = mid, c = last - 1;
                const float ka = r.key(a), kb = r.key(b), kc = r.key(c);
                if (ka < kb) {
                    if (kb < kc) r.swap(first, b);
                    else if (ka < kc) r.swap(first, c);
                    else r.swap(first, a);
                } else if (ka < kc) r.swap(first, a);
                else if (kb < kc) r.swap(first, c);
                else r.swap(first, b);
            }
            int lo = first + 1, hi = last;
            const float pivot = r.key(first);   // the pivot element itself never moves during the partition
            while (true) {
                while (r.key(lo) < pivot) ++lo;
                --hi;
                while (pivot < r.key(hi)) --hi;
                if (!(lo < hi)) break;
                r.swap(lo, hi);
                ++lo;
            }
            const int cut = lo;
            if (sp < 64) { stackFirst[sp] = cut; stackLast[sp] = last; stackDepth[sp] = depth; sp++; }
            last = cut;
        }
    }
    // __final_insertion_sort
    if (n > kThreshold) {
        ss_insertion_sort(r, 0, kThreshold);
        for (int i = kThreshold; i != n; ++i) ss_unguarded_linear_insert(r, i);
    } else {
        ss_insertion_sort(r, 0, n);
    }
}

}   // namespace atlas
