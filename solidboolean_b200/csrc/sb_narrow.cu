// Stage 3: narrow phase.  Replaces the candidate loop of SolidBoolean::combine()
// (reference src/solidboolean.cpp:315-320) and intersectTwoFaces (:103-122):
// one thread per candidate pair runs the Guigue-Devillers predicate
// (sb_tritri.cuh) in strict binary64, records (ret, coplanar) in the two low
// bits of the pair key, and compacts accepted pairs (ret && !coplanar) with their
// segment into the hit buffer (warp-aggregated atomics).  Both lists are then
// ordered by (a, b) with the onesweep radix sort.
#include "sb_internal.h"
#include <cstdlib>
#include "sb_radix.cuh"
#include "sb_tritri.cuh"

namespace {

__global__ void __launch_bounds__(128) predicate_kernel(unsigned long long *__restrict__ keys, uint32_t nPairs, unsigned bitsB,
    const double4 *__restrict__ vtxA, const uint32_t *__restrict__ triA,
    const double4 *__restrict__ vtxB, const uint32_t *__restrict__ triB,
    unsigned long long *__restrict__ hitKeys, uint32_t *__restrict__ hitSlot, double2 *__restrict__ hitSeg,
    unsigned int *__restrict__ hitCount, uint8_t *__restrict__ flagsA, uint8_t *__restrict__ flagsB,
    unsigned long long *__restrict__ pathCounts, uint8_t *__restrict__ hitTag)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool hit = false;
    int path = -1;
    unsigned long long ab = 0;
    d3 src = {0, 0, 0}, tgt = {0, 0, 0};
    uint32_t a = 0, b = 0;
    unsigned tag = 0;
    if (i < nPairs) {
        unsigned long long key = keys[i];
        ab = key >> 2;
        a = (uint32_t)(ab >> bitsB);
        b = (uint32_t)(ab & ((1ull << bitsB) - 1));
        uint32_t a0 = __ldg(triA + 3 * (size_t)a), a1 = __ldg(triA + 3 * (size_t)a + 1), a2 = __ldg(triA + 3 * (size_t)a + 2);
        uint32_t b0 = __ldg(triB + 3 * (size_t)b), b1 = __ldg(triB + 3 * (size_t)b + 1), b2 = __ldg(triB + 3 * (size_t)b + 2);
        d3 p1 = load_vertex(vtxA, a0), q1 = load_vertex(vtxA, a1), r1 = load_vertex(vtxA, a2);
        d3 p2 = load_vertex(vtxB, b0), q2 = load_vertex(vtxB, b1), r2 = load_vertex(vtxB, b2);
        int coplanar = 0;
        int ret = tri_tri_intersection(p1, q1, r1, p2, q2, r2, coplanar, src, tgt, path, &tag);
        keys[i] = key | (unsigned long long)((ret ? 1 : 0) | (coplanar ? 2 : 0));
        hit = ret && !coplanar; // intersectTwoFaces, src/solidboolean.cpp:117-121
    }
    if (pathCounts) { // exit histogram (flop accounting): one atomic per warp and exit taken
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            uint32_t pm = __ballot_sync(SB_FULL, path == k);
            if (pm && lane == 0)
                atomicAdd(pathCounts + k, (unsigned long long)__popc(pm));
        }
    }
    uint32_t m = __ballot_sync(SB_FULL, hit);
    if (m == 0)
        return;
    unsigned int base = 0;
    if (lane == __ffs(m) - 1)
        base = atomicAdd(hitCount, (unsigned int)__popc(m));
    base = __shfl_sync(SB_FULL, base, __ffs(m) - 1);
    if (hit) {
        unsigned int slot = base + __popc(m & lanemask_lt());
        hitKeys[slot] = ab;
        hitSlot[slot] = slot;
        double2 *o = hitSeg + 3 * (size_t)slot;
        o[0] = make_double2(src.x, src.y);
        o[1] = make_double2(src.z, tgt.x);
        o[2] = make_double2(tgt.y, tgt.z);
        if (hitTag)
            hitTag[slot] = (uint8_t)tag; // which edges the end points lie on (sb_tritri.cuh tt_segment_tag)
        flagsA[a] = 1;
        flagsB[b] = 1;
    }
}

__global__ void __launch_bounds__(256) gather_hits_kernel(const unsigned long long *__restrict__ sortedKeys,
    const uint32_t *__restrict__ sortedSlot, const double2 *__restrict__ hitSeg, uint32_t nHits, unsigned bitsB,
    uint32_t *__restrict__ outAB, double2 *__restrict__ outSeg, const uint8_t *__restrict__ hitTag, uint8_t *__restrict__ outTag)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nHits)
        return;
    if (outTag)
        outTag[i] = hitTag[sortedSlot[i]];
    unsigned long long ab = sortedKeys[i];
    outAB[2 * (size_t)i] = (uint32_t)(ab >> bitsB);
    outAB[2 * (size_t)i + 1] = (uint32_t)(ab & ((1ull << bitsB) - 1));
    const double2 *src = hitSeg + 3 * (size_t)sortedSlot[i];
    double2 *dst = outSeg + 3 * (size_t)i;
    dst[0] = src[0];
    dst[1] = src[1];
    dst[2] = src[2];
}

__global__ void __launch_bounds__(256) decode_candidates_kernel(const unsigned long long *__restrict__ keys, uint32_t n,
    unsigned bitsB, uint32_t *__restrict__ outAB, uint8_t *__restrict__ outCode)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    unsigned long long key = keys[i];
    unsigned long long ab = key >> 2;
    reinterpret_cast<uint2 *>(outAB)[i] = make_uint2((uint32_t)(ab >> bitsB), (uint32_t)(ab & ((1ull << bitsB) - 1)));
    if (outCode)
        outCode[i] = (uint8_t)(key & 3);
}

__global__ void __launch_bounds__(128) tri_tri_batch_kernel(const double *__restrict__ tris, uint32_t n,
    int32_t *__restrict__ ret, int32_t *__restrict__ coplanar, double *__restrict__ seg)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const double *v = tris + 18 * (size_t)i;
    d3 p1 = {v[0], v[1], v[2]}, q1 = {v[3], v[4], v[5]}, r1 = {v[6], v[7], v[8]};
    d3 p2 = {v[9], v[10], v[11]}, q2 = {v[12], v[13], v[14]}, r2 = {v[15], v[16], v[17]};
    int cop = 0;
    d3 s = {0, 0, 0}, t = {0, 0, 0};
    int r = tri_tri_intersection(p1, q1, r1, p2, q2, r2, cop, s, t);
    ret[i] = r;
    coplanar[i] = cop;
    double *o = seg + 6 * (size_t)i;
    o[0] = s.x; o[1] = s.y; o[2] = s.z;
    o[3] = t.x; o[4] = t.y; o[5] = t.z;
}

} // namespace

cudaError_t sbk_predicate(cudaStream_t s, const MeshDev &A, const MeshDev &B, unsigned long long *keys, uint32_t nPairs,
    unsigned bitsB, unsigned long long *hitKeys, uint32_t *hitSlot, double2 *hitSeg, unsigned int *hitCount,
    uint8_t *flagsA, uint8_t *flagsB, unsigned long long *pathCounts, uint8_t *hitTag, LaunchCounter &lc)
{
    if (nPairs == 0)
        return cudaSuccess;
    predicate_kernel<<<(nPairs + 127) / 128, 128, 0, s>>>(keys, nPairs, bitsB, A.vtx, A.tri, B.vtx, B.tri, hitKeys, hitSlot,
        hitSeg, hitCount, flagsA, flagsB, pathCounts, hitTag);
    lc.kernels += 1;
    return cudaGetLastError();
}

cudaError_t sbk_gather_hits(cudaStream_t s, const unsigned long long *sortedHitKeys, const uint32_t *sortedSlot,
    const double2 *hitSeg, uint32_t nHits, unsigned bitsB, uint32_t *outAB, double2 *outSeg, const uint8_t *hitTag, uint8_t *outTag,
    LaunchCounter &lc)
{
    if (nHits == 0)
        return cudaSuccess;
    gather_hits_kernel<<<(nHits + 255) / 256, 256, 0, s>>>(sortedHitKeys, sortedSlot, hitSeg, nHits, bitsB, outAB, outSeg, hitTag, outTag);
    lc.kernels += 1;
    return cudaGetLastError();
}

cudaError_t sbk_decode_candidates(cudaStream_t s, const unsigned long long *keys, uint32_t n, unsigned bitsB,
    uint32_t *outAB, uint8_t *outCode, LaunchCounter &lc)
{
    if (n == 0)
        return cudaSuccess;
    decode_candidates_kernel<<<(n + 255) / 256, 256, 0, s>>>(keys, n, bitsB, outAB, outCode);
    lc.kernels += 1;
    return cudaGetLastError();
}

cudaError_t sbk_tri_tri_batch(cudaStream_t s, const double *tris18, uint32_t n, int32_t *ret, int32_t *coplanar,
    double *seg6, LaunchCounter &lc)
{
    if (n == 0)
        return cudaSuccess;
    tri_tri_batch_kernel<<<(n + 127) / 128, 128, 0, s>>>(tris18, n, ret, coplanar, seg6);
    lc.kernels += 1;
    return cudaGetLastError();
}

// Issue-rate microbenchmark for the FP64 roofline of the predicate: the
// predicate is built without FMA contraction, so its ceiling is the DADD/DMUL rate.
template <bool FMA>
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters, double seed)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        if (FMA) {
            a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
            a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
        } else {
            a0 = __dadd_rn(__dmul_rn(a0, m), c); a1 = __dadd_rn(__dmul_rn(a1, m), c);
            a2 = __dadd_rn(__dmul_rn(a2, m), c); a3 = __dadd_rn(__dmul_rn(a3, m), c);
            a4 = __dadd_rn(__dmul_rn(a4, m), c); a5 = __dadd_rn(__dmul_rn(a5, m), c);
            a6 = __dadd_rn(__dmul_rn(a6, m), c); a7 = __dadd_rn(__dmul_rn(a7, m), c);
        }
    }
    double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 12345.678)
        out[0] = r; // never true: keeps the chains alive
}

cudaError_t sbk_fp64_peak(cudaStream_t s, int smCount, double *scratch, int iters, bool fma, unsigned long long *flops)
{
    const int blocks = smCount * 8;
    if (fma)
        fp64_peak_kernel<true><<<blocks, 256, 0, s>>>(scratch, iters, 0.5);
    else
        fp64_peak_kernel<false><<<blocks, 256, 0, s>>>(scratch, iters, 0.5);
    *flops = (unsigned long long)blocks * 256ull * 8ull * 2ull * (unsigned long long)iters; // FMA counted as 2
    return cudaGetLastError();
}

// ---- small sorts ---------------------------------------------------------------------------
// A few thousand keys (the hit list of a small or barely touching pair of meshes) do not need the onesweep passes -- a
// histogram and one launch per 8-bit digit, ~10 us each whatever the size: 55-80 us for 1,400-10,000 keys.  One CTA sorts
// (compare key, original position) pairs with a bitonic network in shared memory and gathers keys and values through the
// resulting permutation: the same order as the stable radix sort on bits [beginBit, endBit) gives.
namespace {
constexpr uint32_t SMALL_SORT_MAX = 4096;

__global__ void __launch_bounds__(1024) small_sort_kernel(const unsigned long long *__restrict__ keys, unsigned long long *__restrict__ keysOut,
    const uint32_t *__restrict__ vals, uint32_t *__restrict__ valsOut, uint32_t n, int beginBit, int endBit)
{
    __shared__ unsigned long long sk[SMALL_SORT_MAX];
    __shared__ uint32_t si[SMALL_SORT_MAX];
    uint32_t N = 2;
    while (N < n)
        N <<= 1;
    const int width = endBit - beginBit;
    const unsigned long long mask = width >= 64 ? ~0ull : ((1ull << width) - 1ull);
    for (uint32_t j = threadIdx.x; j < N; j += blockDim.x) {
        sk[j] = j < n ? ((keys[j] >> beginBit) & mask) : ~0ull;
        si[j] = j < n ? j : 0xffffffffu; // (padding sorts behind every real element, whatever its key)
    }
    __syncthreads();
    for (uint32_t k = 2; k <= N; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t t = threadIdx.x; t < N / 2; t += blockDim.x) {
                const uint32_t i = ((t / j) * 2 * j) + (t % j), p = i + j;
                const bool ascending = (i & k) == 0;
                const unsigned long long a = sk[i], b = sk[p];
                const uint32_t ia = si[i], ib = si[p];
                const bool greater = a > b || (a == b && ia > ib);
                if (greater == ascending) {
                    sk[i] = b; sk[p] = a;
                    si[i] = ib; si[p] = ia;
                }
            }
            __syncthreads();
        }
    for (uint32_t j = threadIdx.x; j < n; j += blockDim.x) {
        const uint32_t src = si[j];
        keysOut[j] = keys[src];
        if (vals)
            valsOut[j] = vals[src];
    }
}
} // namespace

cudaError_t sbk_sort_keys(cudaStream_t s, unsigned long long *keys, unsigned long long *keysTmp, uint32_t *vals,
    uint32_t *valsTmp, size_t n, int beginBit, int endBit, uint32_t *radixWs, int smCount,
    unsigned long long **outKeys, uint32_t **outVals, LaunchCounter &lc)
{
    if (n >= 2 && n <= SMALL_SORT_MAX && endBit > beginBit && !getenv("SB_NO_SMALL_SORT")) {
        small_sort_kernel<<<1, 1024, 0, s>>>(keys, keysTmp, vals, valsTmp, (uint32_t)n, beginBit, endBit);
        lc.kernels += 1;
        *outKeys = keysTmp;
        if (outVals)
            *outVals = vals ? valsTmp : nullptr;
        return cudaGetLastError();
    }
    sbradix::Workspace ws;
    ws.mem = radixWs;
    lc.kernels += sbradix::sort<unsigned long long, 8>(s, keys, keysTmp, vals, valsTmp, n, beginBit, endBit, ws, smCount,
        outKeys, outVals);
    return cudaGetLastError();
}
