// Stage 2: broad phase.  Replaces SolidBoolean::searchPotentialIntersectedPairs
// (reference src/solidboolean.cpp:94-101) and the dual descent it calls
// (axisalignedboundingboxtree.h:54-95).  Emits every (a, b) whose exact double
// triangle boxes overlap on closed intervals (axisalignedboundingbox.h:95-105).
//
// One warp per group of 32 Morton-consecutive triangles of A; the group walks
// B's cluster LBVH once (sb_traverse.cuh).  Accepted pairs are compacted with
// ballot + popcount into a per-warp shared staging buffer and flushed to the
// global pair buffer as runs of >= 32 consecutive 8-byte keys with a single
// atomicAdd per flush.
#include "sb_internal.h"
#include "sb_traverse.cuh"

namespace {

constexpr int K = SB_CLUSTER;
constexpr int WARPS_PER_CTA = 8;

struct BroadShared {
    sbtrav::WarpScratch trav;
    unsigned long long out[64];
};

__global__ void __launch_bounds__(WARPS_PER_CTA * 32) broad_phase_kernel(
    const Rec32 *__restrict__ leafA, const double2 *__restrict__ sboxA, uint32_t groupBegin, uint32_t groupEnd,
    const Rec32 *__restrict__ nodesB, const Rec32 *__restrict__ leafB, const double2 *__restrict__ sboxB,
    const int *__restrict__ rootB, const unsigned long long *__restrict__ boundsB, unsigned bitsB,
    unsigned long long *__restrict__ outKeys, unsigned long long capacity, unsigned long long *__restrict__ outCount,
    const double *__restrict__ scentA, double ownLo, double ownHi, int ownMode /* 0 off, 1 [lo, hi), 2 [lo, hi] */)
{
    __shared__ BroadShared sh[WARPS_PER_CTA];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t group = groupBegin + blockIdx.x * WARPS_PER_CTA + warp;
    if (group >= groupEnd)
        return;
    BroadShared &s = sh[warp];
    const uint32_t lt = lanemask_lt();

    const uint32_t j = group * 32 + lane;
    Rec32 me = load_rec(leafA + j);
    BoxF myF = {me.lox, me.loy, me.loz, me.hix, me.hiy, me.hiz};
    if (ownMode) {
        // multi-GPU selection: only the faces whose centroid lies in this rank's slab ask (sb_shard.cu);
        // the others -- and NaN -- hold an empty box like the padding lanes
        const double cz = scentA[3 * (size_t)j + 2];
        if (!(me.ref >= 0 && cz >= ownLo && (ownMode == 2 ? cz <= ownHi : cz < ownHi)))
            myF = empty_boxf();
    }
    // groups whose box misses B's bounding box altogether (most of a mesh, usually) leave
    // before they touch their exact boxes or B's tree
    BoxF all = myF; // padding lanes hold an empty box and never match
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        BoxF o = shfl_xor_box(all, off);
        merge_f(all, o);
    }
    if (boundsB) { // (batch meshes: the float boxes live on the job lattice, the bounds do not)
        const BoxD bb = {dkey_inv(__ldg(boundsB)), dkey_inv(__ldg(boundsB + 1)), dkey_inv(__ldg(boundsB + 2)),
                         dkey_inv(__ldg(boundsB + 3)), dkey_inv(__ldg(boundsB + 4)), dkey_inv(__ldg(boundsB + 5))};
        const BoxF bf = enclose(bb); // conservative: rounded outwards
        if (!overlap_f(all, bf.lox, bf.loy, bf.loz, bf.hix, bf.hiy, bf.hiz))
            return;
    }
    const BoxD myD = load_boxd(sboxA + 3 * (size_t)j);
    const unsigned long long myA = (unsigned long long)(uint32_t)me.ref;
    sbtrav::BvhView bvh = {nodesB, leafB, __ldg(rootB)};
    int ocount = 0;
    auto flush = [&]() {
        unsigned long long base = 0;
        if (lane == 0)
            base = atomicAdd(outCount, (unsigned long long)ocount);
        base = __shfl_sync(SB_FULL, base, 0);
        for (int i = lane; i < ocount; i += 32)
            if (base + i < capacity)
                outKeys[base + i] = s.out[i];
        ocount = 0;
        __syncwarp();
    };

    // surface-area measure of a query box (0 for an empty one); the pad keeps
    // flat boxes from measuring zero
    const float pad = 0.015625f * fmaxf(fmaxf(all.hix - all.lox, all.hiy - all.loy), all.hiz - all.loz);
    auto measure = [pad](const BoxF &b) {
        float ex = b.hix - b.lox, ey = b.hiy - b.loy, ez = b.hiz - b.loz;
        if (ex < 0.0f || ey < 0.0f || ez < 0.0f)
            return 0.0f;
        ex += pad; ey += pad; ez += pad;
        return ex * ey + ey * ez + ez * ex;
    };
    sbtrav::split_and_run(myF, lane, s.trav.segs, measure, [&](const BoxF &G, bool inSeg) {
        sbtrav::group_traverse<K>(bvh, G, s.trav, lane, [&](const Rec32 &r, uint32_t posB) {
            bool pass = inSeg && overlap_f(myF, r.lox, r.loy, r.loz, r.hix, r.hiy, r.hiz);
            if (pass) {
                BoxD bd = load_boxd(sboxB + 3 * (size_t)posB);
                // boxes[a].intersectWith(secondBoxes[b]) -- the reference's leaf test
                pass = overlap_d(myD, bd);
            }
            uint32_t m = __ballot_sync(SB_FULL, pass);
            if (m) {
                if (pass)
                    s.out[ocount + __popc(m & lt)] = ((myA << bitsB) | (unsigned long long)(uint32_t)r.ref) << 2;
                ocount += __popc(m);
                __syncwarp();
                if (ocount >= 32)
                    flush();
            }
        });
    });
    if (ocount > 0)
        flush();
}

} // namespace

cudaError_t sbk_broad_phase(cudaStream_t s, const MeshDev &A, const MeshDev &B, uint32_t groupBegin, uint32_t groupEnd,
    unsigned bitsB, unsigned long long *outKeys, unsigned long long capacity, unsigned long long *outCount,
    int *errFlag, LaunchCounter &lc)
{
    (void)errFlag;
    if (groupEnd <= groupBegin || B.nT == 0)
        return cudaSuccess;
    uint32_t groups = groupEnd - groupBegin;
    uint32_t blocks = (groups + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    broad_phase_kernel<<<blocks, WARPS_PER_CTA * 32, 0, s>>>(A.leaf, A.sbox, groupBegin, groupEnd, B.nodes, B.leaf, B.sbox,
        B.root, B.triJob ? nullptr : B.bounds, bitsB, outKeys, capacity, outCount, A.scent, A.ownLo, A.ownHi,
        A.ownFilter ? (A.ownClosed ? 2 : 1) : 0);
    lc.kernels += 1;
    return cudaGetLastError();
}
