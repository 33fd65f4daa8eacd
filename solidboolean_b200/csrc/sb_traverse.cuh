// Warp-cooperative traversal of a cluster LBVH by a GROUP of 32 queries.
//
// A warp owns 32 spatially coherent queries (Morton-consecutive triangles of
// mesh A, or 32 rays of neighbouring points) and walks the target's LBVH once
// for all of them with the union ("group") box:
//   * a per-warp shared-memory STACK of internal nodes; each step pops up to 32
//     nodes, one per lane, tests both children against the group box, and
//     pushes the survivors with ballot + prefix-popcount compaction;
//   * overlapping leaf CLUSTERS go to a per-warp leaf queue; when it fills (or
//     the stack runs dry) the warp loads 32 leaf records (32/K clusters) with
//     one coalesced 128-bit load per lane pair, parks them in shared memory, and
//     every lane tests ITS OWN query against each of them (broadcast reads).
// This replaces the reference's recursive dual descent
// (AxisAlignedBoudingBoxTree::testNodes, axisalignedboundingboxtree.h:54-88).
// Inner boxes are conservative floats; the visit callback applies the exact
// double test, so the accepted set is the reference's.
#pragma once
#include "sb_common.cuh"

namespace sbtrav {

constexpr int STACK_CAP = 512; // entries per warp
constexpr int STACK_DFS_ROOM = 96; // below this much room pop one node at a time
constexpr int LQ_CAP = 128;    // leaf-cluster queue entries per warp

struct WarpScratch {
    uint32_t stack[STACK_CAP];
    uint32_t lq[LQ_CAP];
    Rec32 stage[32];
    uint16_t segs[32];
};

struct BvhView {
    const Rec32 *nodes;
    const Rec32 *leaf;
    int root; // >=0 internal node, <0 ~cluster (single-cluster mesh)
};

// visit(rec, sortedPos): called by ALL lanes, converged, once per staged leaf.
template <int K, typename Visit>
__device__ __forceinline__ void group_traverse(const BvhView &bvh, const BoxF &G, WarpScratch &ws, int lane, Visit &&visit)
{
    constexpr int PER = 32 / K; // clusters per drain
    const uint32_t lt = lanemask_lt();
    int top = 0, nq = 0;
    if (bvh.root < 0) {
        if (lane == 0)
            ws.lq[0] = (uint32_t)(~bvh.root);
        nq = 1;
    } else {
        if (lane == 0)
            ws.stack[0] = (uint32_t)bvh.root;
        top = 1;
    }
    __syncwarp();
    while (top > 0 || nq > 0) {
        if (top > 0 && nq <= LQ_CAP - 64) {
            // ---- expand up to 32 internal nodes ----
            int k = min(top, 32);
            if (STACK_CAP - top < STACK_DFS_ROOM)
                k = 1;
            bool have = lane < k;
            uint32_t node = have ? ws.stack[top - 1 - lane] : 0u;
            top -= k;
            __syncwarp();
            bool i0 = false, i1 = false, l0 = false, l1 = false;
            int r0 = 0, r1 = 0;
            if (have) {
                Rec32 c0 = load_rec(bvh.nodes + 2 * (size_t)node);
                Rec32 c1 = load_rec(bvh.nodes + 2 * (size_t)node + 1);
                bool o0 = overlap_f(G, c0.lox, c0.loy, c0.loz, c0.hix, c0.hiy, c0.hiz);
                bool o1 = overlap_f(G, c1.lox, c1.loy, c1.loz, c1.hix, c1.hiy, c1.hiz);
                r0 = c0.ref;
                r1 = c1.ref;
                i0 = o0 && r0 >= 0;
                l0 = o0 && r0 < 0;
                i1 = o1 && r1 >= 0;
                l1 = o1 && r1 < 0;
            }
            uint32_t m;
            m = __ballot_sync(SB_FULL, i0);
            if (i0) ws.stack[top + __popc(m & lt)] = (uint32_t)r0;
            top += __popc(m);
            m = __ballot_sync(SB_FULL, i1);
            if (i1) ws.stack[top + __popc(m & lt)] = (uint32_t)r1;
            top += __popc(m);
            m = __ballot_sync(SB_FULL, l0);
            if (l0) ws.lq[nq + __popc(m & lt)] = (uint32_t)(~r0);
            nq += __popc(m);
            m = __ballot_sync(SB_FULL, l1);
            if (l1) ws.lq[nq + __popc(m & lt)] = (uint32_t)(~r1);
            nq += __popc(m);
            __syncwarp();
        } else {
            // ---- drain up to PER clusters = 32 leaf records ----
            int take = min(nq, PER);
            nq -= take;
            int ci = lane / K;
            Rec32 rec;
            rec.lox = rec.loy = rec.loz = FLT_MAX;
            rec.hix = rec.hiy = rec.hiz = -FLT_MAX;
            rec.ref = -1;
            rec.aux = 0;
            if (ci < take) {
                uint32_t pos = ws.lq[nq + ci] * K + (lane % K);
                rec = load_rec(bvh.leaf + pos);
            }
            ws.stage[lane] = rec;
            __syncwarp();
            int count = take * K;
            for (int j = 0; j < count; ++j) {
                Rec32 r = ws.stage[j];
                visit(r, (uint32_t)r.aux);
            }
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------
// Query-side splitting.  32 curve-consecutive queries are usually one compact
// patch, but a space-filling-curve jump can put two distant patches into one
// warp; the union box then covers a large part of the target and that single
// warp would decide the kernel's run time.  Before traversing, the lane range is
// therefore split top-down at the position that minimises
// measure(prefix box) + measure(suffix box) (a surface-area heuristic on the
// QUERY side, evaluated with two warp scans), as long as a split at least
// halves... see SPLIT_RATIO.  run(G, inSeg) is called once per final segment.
constexpr float SPLIT_RATIO = 0.6f;

__device__ __forceinline__ BoxF shfl_up_box(const BoxF &b, int d)
{
    return {__shfl_up_sync(SB_FULL, b.lox, d), __shfl_up_sync(SB_FULL, b.loy, d), __shfl_up_sync(SB_FULL, b.loz, d),
            __shfl_up_sync(SB_FULL, b.hix, d), __shfl_up_sync(SB_FULL, b.hiy, d), __shfl_up_sync(SB_FULL, b.hiz, d)};
}
__device__ __forceinline__ BoxF shfl_down_box(const BoxF &b, int d)
{
    return {__shfl_down_sync(SB_FULL, b.lox, d), __shfl_down_sync(SB_FULL, b.loy, d), __shfl_down_sync(SB_FULL, b.loz, d),
            __shfl_down_sync(SB_FULL, b.hix, d), __shfl_down_sync(SB_FULL, b.hiy, d), __shfl_down_sync(SB_FULL, b.hiz, d)};
}
__device__ __forceinline__ BoxF shfl_box(const BoxF &b, int src)
{
    return {__shfl_sync(SB_FULL, b.lox, src), __shfl_sync(SB_FULL, b.loy, src), __shfl_sync(SB_FULL, b.loz, src),
            __shfl_sync(SB_FULL, b.hix, src), __shfl_sync(SB_FULL, b.hiy, src), __shfl_sync(SB_FULL, b.hiz, src)};
}

// measure(box) must return 0 for an empty box and grow with the expected
// traversal cost of a query box.
template <typename Measure, typename Run>
__device__ __forceinline__ void split_and_run(const BoxF &myF, int lane, uint16_t *segs, Measure &&measure, Run &&run)
{
    int nseg = 1;
    if (lane == 0)
        segs[0] = (uint16_t)(0 | (32 << 8));
    __syncwarp();
    while (nseg > 0) {
        uint16_t se = segs[nseg - 1];
        --nseg;
        __syncwarp();
        const int s = se & 0xff, e = se >> 8;
        const bool inSeg = lane >= s && lane < e;
        BoxF mine = inSeg ? myF : empty_boxf();
        // inclusive prefix / suffix unions over the lanes
        BoxF pre = mine, suf = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            BoxF a = shfl_up_box(pre, d);
            BoxF b = shfl_down_box(suf, d);
            if (lane >= d)
                merge_f(pre, a);
            if (lane + d < 32)
                merge_f(suf, b);
        }
        const BoxF G = shfl_box(pre, 31);
        if (e - s > 1) {
            BoxF nextSuf = shfl_down_box(suf, 1);
            float cost = (lane >= s && lane < e - 1) ? measure(pre) + measure(nextSuf) : 3.0e38f;
            int best = lane;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                float oc = __shfl_xor_sync(SB_FULL, cost, d);
                int ob = __shfl_xor_sync(SB_FULL, best, d);
                if (oc < cost || (oc == cost && ob < best)) {
                    cost = oc;
                    best = ob;
                }
            }
            if (cost < SPLIT_RATIO * measure(G)) {
                if (lane == 0) {
                    segs[nseg] = (uint16_t)((best + 1) | (e << 8));
                    segs[nseg + 1] = (uint16_t)(s | ((best + 1) << 8));
                }
                nseg += 2;
                __syncwarp();
                continue;
            }
        }
        run(G, inSeg);
    }
}

} // namespace sbtrav
