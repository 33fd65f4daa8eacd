// Stage 4: inside/outside classification.  Replaces SolidBoolean::isPointInMesh
// (reference src/solidboolean.cpp:48-92) as driven by decideGroupSide (:482-510).
//
// One warp per group of 32 query points (face centroids in the query mesh's
// Morton order, or caller-supplied points).  For each of the three reference
// axes (g_testAxisList, :31-35) the 32 rays form a thin "beam" whose union box
// walks the target's cluster LBVH once (sb_traverse.cuh); every staged leaf is
// then tested by each lane against ITS ray box -- first the conservative float
// box, then the exact double box (the reference's candidate definition, a ray
// box against triangle boxes through AxisAlignedBoudingBoxTree::test, :55-63) --
// and accepted candidates run the reference arithmetic (sb_raytri.cuh).  Hits
// are de-duplicated per ray by PositionKey in a small per-lane list; parity of
// the distinct count is the per-axis answer, the majority of three the result.
#include "sb_internal.h"
#include "sb_raytri.cuh"
#include "sb_traverse.cuh"

namespace {

constexpr int K = SB_CLUSTER;
constexpr int WARPS_PER_CTA = 8;
constexpr int KEY_LIST = 16; // distinct hit keys kept per ray before the slow path

struct KeyList {
    long long x[KEY_LIST], y[KEY_LIST], z[KEY_LIST];
};

__device__ __forceinline__ BoxD ray_box(const d3 &p, const d3 &e)
{
    // box.update(testPosition); box.update(testEnd)  (src/solidboolean.cpp:55-58)
    BoxD b = {DBL_MAX, DBL_MAX, DBL_MAX, -DBL_MAX, -DBL_MAX, -DBL_MAX};
#define SB_UPD(v)                                 \
    if (v.x > b.hix) b.hix = v.x;                 \
    if (v.x < b.lox) b.lox = v.x;                 \
    if (v.y > b.hiy) b.hiy = v.y;                 \
    if (v.y < b.loy) b.loy = v.y;                 \
    if (v.z > b.hiz) b.hiz = v.z;                 \
    if (v.z < b.loz) b.loz = v.z;
    SB_UPD(p) SB_UPD(e)
#undef SB_UPD
    return b;
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32) classify_kernel(
    const double *__restrict__ pts,              // explicit points, or null
    const Rec32 *__restrict__ qLeaf,             // query mesh leaves (faces mode)
    const double4 *__restrict__ qVtx, const uint32_t *__restrict__ qTri,
    uint32_t begin, uint32_t end,
    const Rec32 *__restrict__ nodes, const Rec32 *__restrict__ leaf, const double2 *__restrict__ sbox,
    const int *__restrict__ root, const double4 *__restrict__ vtx, const uint32_t *__restrict__ tri,
    const double *__restrict__ normal,
    uint8_t *__restrict__ inside, uint8_t *__restrict__ perAxis, unsigned long long *__restrict__ stats,
    uint32_t *__restrict__ overflowList, unsigned int *__restrict__ overflowCount, uint32_t overflowCap)
{
    __shared__ sbtrav::WarpScratch sh[WARPS_PER_CTA];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t group = blockIdx.x * WARPS_PER_CTA + warp;
    const uint32_t first = begin + group * 32;
    if (first >= end)
        return;
    const uint32_t idx = first + lane;

    // ---- this lane's query point ----
    bool active = idx < end;
    uint32_t outIndex = idx;
    d3 p = {0, 0, 0};
    if (pts) {
        if (active)
            p = {pts[3 * (size_t)idx], pts[3 * (size_t)idx + 1], pts[3 * (size_t)idx + 2]};
    } else {
        int t = active ? load_rec(qLeaf + idx).ref : -1;
        active = t >= 0;
        if (active) {
            outIndex = (uint32_t)t;
            d3 a = load_vertex(qVtx, __ldg(qTri + 3 * (size_t)t));
            d3 b = load_vertex(qVtx, __ldg(qTri + 3 * (size_t)t + 1));
            d3 c = load_vertex(qVtx, __ldg(qTri + 3 * (size_t)t + 2));
            // (v0 + v1 + v2) / 3.0  (src/solidboolean.cpp:497-499)
            p = {xdiv(xadd(xadd(a.x, b.x), c.x), 3.0), xdiv(xadd(xadd(a.y, b.y), c.y), 3.0),
                 xdiv(xadd(xadd(a.z, b.z), c.z), 3.0)};
        }
    }

    sbtrav::BvhView bvh = {nodes, leaf, __ldg(root)};
    KeyList keys;
    int insideCount = 0;
    bool overflow = false;
    unsigned int candCount = 0;

    for (int axis = 0; axis < 3; ++axis) {
        const d3 e = ray_end(p, axis);
        const BoxD myD = ray_box(p, e);
        BoxF myF = active ? enclose(myD) : empty_boxf();
        // beam measure: cross-section perpendicular to the ray axis
        BoxF all = myF;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            BoxF o = shfl_xor_box(all, off);
            merge_f(all, o);
        }
        const float ax = all.hix - all.lox, ay = all.hiy - all.loy, az = all.hiz - all.loz;
        const float pad = 0.015625f * (axis == 0 ? fmaxf(ay, az) : axis == 1 ? fmaxf(ax, az) : fmaxf(ax, ay));
        auto measure = [pad, axis](const BoxF &b) {
            float ex = b.hix - b.lox, ey = b.hiy - b.loy, ez = b.hiz - b.loz;
            if (ex < 0.0f || ey < 0.0f || ez < 0.0f)
                return 0.0f;
            float u = axis == 0 ? ey : ex, v = axis == 2 ? ey : ez;
            return (u + pad) * (v + pad);
        };
        int nKeys = 0;
        sbtrav::split_and_run(myF, lane, sh[warp].segs, measure, [&](const BoxF &G, bool inSeg) {
        sbtrav::group_traverse<K>(bvh, G, sh[warp], lane, [&](const Rec32 &r, uint32_t posB) {
            if (!inSeg || !overlap_f(myF, r.lox, r.loy, r.loz, r.hix, r.hiy, r.hiz))
                return;
            BoxD bd = load_boxd(sbox + 3 * (size_t)posB);
            if (!overlap_d(bd, myD)) // meshTree boxes .intersectWith(rayBox)
                return;
            ++candCount;
            const uint32_t f = (uint32_t)r.ref;
            d3 t0 = load_vertex(vtx, __ldg(tri + 3 * (size_t)f));
            d3 t1 = load_vertex(vtx, __ldg(tri + 3 * (size_t)f + 1));
            d3 t2 = load_vertex(vtx, __ldg(tri + 3 * (size_t)f + 2));
            d3 nrm = {__ldg(normal + 3 * (size_t)f), __ldg(normal + 3 * (size_t)f + 1), __ldg(normal + 3 * (size_t)f + 2)};
            d3 hit;
            if (!ray_tri_hit(p, e, t0, t1, t2, nrm, hit))
                return;
            long long kx = position_key(hit.x), ky = position_key(hit.y), kz = position_key(hit.z);
            for (int q = 0; q < nKeys; ++q)
                if (keys.x[q] == kx && keys.y[q] == ky && keys.z[q] == kz)
                    return; // std::set<PositionKey> insert of an existing key
            if (nKeys < KEY_LIST) {
                keys.x[nKeys] = kx;
                keys.y[nKeys] = ky;
                keys.z[nKeys] = kz;
                ++nKeys;
            } else {
                overflow = true;
            }
        });
        });
        bool in = (nKeys & 1) != 0;
        if (active && perAxis)
            perAxis[3 * (size_t)outIndex + axis] = in ? 1 : 0;
        insideCount += in ? 1 : 0;
    }
    if (active) {
        // (float)insideCount / totalCount > 0.5 with totalCount == 3
        inside[outIndex] = insideCount >= 2 ? 1 : 0;
        if (overflow) {
            unsigned int slot = atomicAdd(overflowCount, 1u);
            if (slot < overflowCap)
                overflowList[slot] = pts ? idx : outIndex;
        }
    }
    // work counters (one atomic per warp)
    unsigned int rays = active ? 3u : 0u;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        candCount += __shfl_xor_sync(SB_FULL, candCount, off);
        rays += __shfl_xor_sync(SB_FULL, rays, off);
    }
    if (lane == 0 && stats) {
        atomicAdd(&stats[0], (unsigned long long)rays);
        atomicAdd(&stats[1], (unsigned long long)candCount);
    }
}

// Exact slow path: one thread per (overflowed point, axis), brute force over all
// target triangles, distinct keys kept in global scratch.
__global__ void __launch_bounds__(128) classify_overflow_kernel(
    const double *__restrict__ pts, const double4 *__restrict__ qVtx, const uint32_t *__restrict__ qTri,
    const uint32_t *__restrict__ list, uint32_t nList,
    const double2 *__restrict__ tbox, const double4 *__restrict__ vtx, const uint32_t *__restrict__ tri,
    const double *__restrict__ normal, uint32_t nT,
    long long *__restrict__ scratch, uint32_t keysPerRay, uint8_t *__restrict__ axisOut, int *__restrict__ errFlag)
{
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nList * 3)
        return;
    uint32_t li = g / 3;
    int axis = (int)(g % 3);
    uint32_t id = list[li];
    d3 p;
    if (pts) {
        p = {pts[3 * (size_t)id], pts[3 * (size_t)id + 1], pts[3 * (size_t)id + 2]};
    } else {
        d3 a = load_vertex(qVtx, qTri[3 * (size_t)id]);
        d3 b = load_vertex(qVtx, qTri[3 * (size_t)id + 1]);
        d3 c = load_vertex(qVtx, qTri[3 * (size_t)id + 2]);
        p = {xdiv(xadd(xadd(a.x, b.x), c.x), 3.0), xdiv(xadd(xadd(a.y, b.y), c.y), 3.0),
             xdiv(xadd(xadd(a.z, b.z), c.z), 3.0)};
    }
    const d3 e = ray_end(p, axis);
    const BoxD myD = ray_box(p, e);
    long long *my = scratch + (size_t)g * keysPerRay * 3;
    uint32_t nKeys = 0;
    for (uint32_t f = 0; f < nT; ++f) {
        BoxD bd = load_boxd(tbox + 3 * (size_t)f);
        if (!overlap_d(bd, myD))
            continue;
        d3 t0 = load_vertex(vtx, tri[3 * (size_t)f]);
        d3 t1 = load_vertex(vtx, tri[3 * (size_t)f + 1]);
        d3 t2 = load_vertex(vtx, tri[3 * (size_t)f + 2]);
        d3 nrm = {normal[3 * (size_t)f], normal[3 * (size_t)f + 1], normal[3 * (size_t)f + 2]};
        d3 hit;
        if (!ray_tri_hit(p, e, t0, t1, t2, nrm, hit))
            continue;
        long long kx = position_key(hit.x), ky = position_key(hit.y), kz = position_key(hit.z);
        bool dup = false;
        for (uint32_t q = 0; q < nKeys && !dup; ++q)
            dup = my[3 * q] == kx && my[3 * q + 1] == ky && my[3 * q + 2] == kz;
        if (dup)
            continue;
        if (nKeys >= keysPerRay) {
            *errFlag = 1;
            break;
        }
        my[3 * nKeys] = kx;
        my[3 * nKeys + 1] = ky;
        my[3 * nKeys + 2] = kz;
        ++nKeys;
    }
    axisOut[g] = (uint8_t)(nKeys & 1);
}

__global__ void classify_overflow_finish_kernel(const uint32_t *__restrict__ list, uint32_t nList,
    const uint8_t *__restrict__ axisOut, uint8_t *__restrict__ inside, uint8_t *__restrict__ perAxis)
{
    uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= nList)
        return;
    uint32_t id = list[li];
    int c = axisOut[3 * li] + axisOut[3 * li + 1] + axisOut[3 * li + 2];
    inside[id] = c >= 2 ? 1 : 0;
    if (perAxis)
        for (int k = 0; k < 3; ++k)
            perAxis[3 * (size_t)id + k] = axisOut[3 * li + k];
}

} // namespace

cudaError_t sbk_classify(cudaStream_t s, const MeshDev &target, const ClassifyArgs &a, int *errFlag, LaunchCounter &lc)
{
    (void)errFlag;
    if (a.end <= a.begin)
        return cudaSuccess;
    uint32_t groups = (a.end - a.begin + 31) / 32;
    uint32_t blocks = (groups + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    const MeshDev *q = a.queryMesh;
    classify_kernel<<<blocks, WARPS_PER_CTA * 32, 0, s>>>(a.pts, q ? q->leaf : nullptr, q ? q->vtx : nullptr,
        q ? q->tri : nullptr, a.begin, a.end, target.nodes, target.leaf, target.sbox, target.root, target.vtx, target.tri,
        target.normal, a.inside, a.perAxis, a.stats, a.overflowList, a.overflowCount, a.overflowCap);
    lc.kernels += 1;
    return cudaGetLastError();
}

cudaError_t sbk_classify_overflow(cudaStream_t s, const MeshDev &target, const ClassifyArgs &a, uint32_t nOverflow,
    long long *scratch, uint32_t scratchKeysPerRay, int *errFlag, LaunchCounter &lc)
{
    if (nOverflow == 0)
        return cudaSuccess;
    const MeshDev *q = a.queryMesh;
    uint8_t *axisOut = reinterpret_cast<uint8_t *>(scratch + (size_t)nOverflow * 3 * scratchKeysPerRay * 3);
    uint32_t rays = nOverflow * 3;
    classify_overflow_kernel<<<(rays + 127) / 128, 128, 0, s>>>(a.pts, q ? q->vtx : nullptr, q ? q->tri : nullptr,
        a.overflowList, nOverflow, target.tbox, target.vtx, target.tri, target.normal, target.nT, scratch,
        scratchKeysPerRay, axisOut, errFlag);
    classify_overflow_finish_kernel<<<(nOverflow + 127) / 128, 128, 0, s>>>(a.overflowList, nOverflow, axisOut, a.inside,
        a.perAxis);
    lc.kernels += 2;
    return cudaGetLastError();
}
