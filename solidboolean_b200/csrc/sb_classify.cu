// Stage 4: inside/outside classification.  Replaces SolidBoolean::isPointInMesh
// (reference src/solidboolean.cpp:48-92) as driven by decideGroupSide (:482-510).
//
// One thread per query point (face centroids in the query mesh's Morton order,
// so a warp's 32 rays read neighbouring grid cells; or caller-supplied points).
// For each of the three reference axes (g_testAxisList, :31-35):
//   1. ray box = {p, p + axis} exactly as :53-58;
//   2. candidates: the ray's cell(s) of the target's axis-projected grid
//      (sb_grid.cu) -- 16-byte references, quantised test, then the EXACT double
//      box test that defines the reference's candidate set (ray box against
//      triangle boxes through AxisAlignedBoudingBoxTree::test, :55-63);
//   3. per candidate the reference arithmetic (sb_raytri.cuh), hits de-duplicated
//      by PositionKey (std::set<PositionKey>, :64/:85) in a small per-thread
//      list; odd count = inside for that axis (:89);
// then the majority of the three axes (:508).  A ray that collects more than
// KEY_LIST distinct keys is redone by the exact slow path below.
#include "sb_internal.h"
#include "sb_raytri.cuh"

namespace {

constexpr int KEY_LIST = 16; // distinct hit keys kept per ray before the slow path
constexpr int THREADS = 128;

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

__device__ __forceinline__ uint32_t quant16(double x, double org, double scl)
{
    double t = floor((x - org) * scl);
    t = fmin(fmax(t, 0.0), 65535.0); // NaN -> 0; identical to the grid build
    return (uint32_t)t;
}

__device__ __forceinline__ double comp(const BoxD &b, int d, bool hi)
{
    return d == 0 ? (hi ? b.hix : b.lox) : d == 1 ? (hi ? b.hiy : b.loy) : (hi ? b.hiz : b.loz);
}

struct Target {
    const GridParams *gp;
    const uint32_t *E;
    const uint4 *refs;
    const uint4 *bigRefs;
    uint32_t bigCap;
    uint32_t bigN[3];
    const double2 *tbox;
    const double4 *vtx;
    const uint32_t *tri;
    const double *normal;
};

__global__ void __launch_bounds__(THREADS) classify_kernel(
    const double *__restrict__ pts,              // explicit points, or null
    const Rec32 *__restrict__ qLeaf,             // query mesh leaves (faces mode)
    const double4 *__restrict__ qVtx, const uint32_t *__restrict__ qTri,
    uint32_t begin, uint32_t end, Target T,
    uint8_t *__restrict__ inside, uint8_t *__restrict__ perAxis, unsigned long long *__restrict__ stats,
    uint32_t *__restrict__ overflowList, unsigned int *__restrict__ overflowCount, uint32_t overflowCap)
{
    __shared__ GridParams g;
    if (threadIdx.x == 0)
        g = *T.gp;
    __syncthreads();
    const uint32_t idx = begin + blockIdx.x * THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31;

    // ---- this thread's query point ----
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

    KeyList keys;
    int insideCount = 0;
    bool overflow = false;
    unsigned int candCount = 0;
    const BoxD meshBox = {g.lo[0], g.lo[1], g.lo[2], g.hi[0], g.hi[1], g.hi[2]};

    if (active) {
        for (int axis = 0; axis < 3; ++axis) {
            const d3 e = ray_end(p, axis);
            const BoxD myD = ray_box(p, e);
            int nKeys = 0;

            auto consider = [&](const uint4 &r, uint32_t aU, uint32_t bU, uint32_t aV, uint32_t bV, uint32_t aA) {
                // quantised closed-interval test (over-accepts only)
                if ((r.x & 0xffffu) > bU || (r.x >> 16) < aU || (r.y & 0xffffu) > bV || (r.y >> 16) < aV ||
                    (r.z & 0xffffu) < aA)
                    return;
                const uint32_t f = r.w;
                BoxD bd = load_boxd(T.tbox + 3 * (size_t)f);
                if (!overlap_d(bd, myD)) // the reference's candidate test, exact
                    return;
                ++candCount;
                d3 t0 = load_vertex(T.vtx, __ldg(T.tri + 3 * (size_t)f));
                d3 t1 = load_vertex(T.vtx, __ldg(T.tri + 3 * (size_t)f + 1));
                d3 t2 = load_vertex(T.vtx, __ldg(T.tri + 3 * (size_t)f + 2));
                d3 nrm = {__ldg(T.normal + 3 * (size_t)f), __ldg(T.normal + 3 * (size_t)f + 1),
                          __ldg(T.normal + 3 * (size_t)f + 2)};
                d3 hit;
                if (!ray_tri_hit_filtered(p, e, t0, t1, t2, nrm, hit))
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
            };

            if (overlap_d(meshBox, myD)) { // otherwise no triangle box can overlap the ray box
                const int u = axis == 0 ? 1 : 0, v = axis == 2 ? 1 : 2;
                const uint32_t aU = quant16(comp(myD, u, false), g.org[u], g.scl[u]);
                const uint32_t bU = quant16(comp(myD, u, true), g.org[u], g.scl[u]);
                const uint32_t aV = quant16(comp(myD, v, false), g.org[v], g.scl[v]);
                const uint32_t bV = quant16(comp(myD, v, true), g.org[v], g.scl[v]);
                const uint32_t aA = quant16(comp(myD, axis, false), g.org[axis], g.scl[axis]);
                const uint32_t cu0 = aU >> g.shiftU[axis], cu1 = bU >> g.shiftU[axis];
                const uint32_t cv0 = aV >> g.shiftV[axis], cv1 = bV >> g.shiftV[axis];
                for (uint32_t cv = cv0; cv <= cv1; ++cv)
                    for (uint32_t cu = cu0; cu <= cu1; ++cu) {
                        const uint32_t cell = g.cellBase[axis] + cv * g.nu[axis] + cu;
                        const uint32_t i0 = __ldg(T.E + cell + 1), i1 = __ldg(T.E + cell + 2);
                        for (uint32_t i = i0; i < i1; ++i) {
                            const uint4 r = __ldg(T.refs + i);
                            // a triangle spanning several of the ray's cells is taken in the first one only
                            const uint32_t tcu = max((r.x & 0xffffu) >> g.shiftU[axis], cu0);
                            const uint32_t tcv = max((r.y & 0xffffu) >> g.shiftV[axis], cv0);
                            if (tcu != cu || tcv != cv)
                                continue;
                            consider(r, aU, bU, aV, bV, aA);
                        }
                    }
                for (uint32_t i = 0; i < T.bigN[axis]; ++i)
                    consider(__ldg(T.bigRefs + (size_t)axis * T.bigCap + i), aU, bU, aV, bV, aA);
            }
            const bool in = (nKeys & 1) != 0;
            if (perAxis)
                perAxis[3 * (size_t)outIndex + axis] = in ? 1 : 0;
            insideCount += in ? 1 : 0;
        }
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
    if (lane == 0 && stats && rays) {
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
    uint32_t blocks = (a.end - a.begin + THREADS - 1) / THREADS;
    const MeshDev *q = a.queryMesh;
    Target T;
    T.gp = target.gridParams;
    T.E = target.gridE;
    T.refs = target.gridRefs;
    T.bigRefs = target.gridBigRefs;
    T.bigCap = target.gridBigCap;
    for (int k = 0; k < 3; ++k)
        T.bigN[k] = target.gridBigN[k];
    T.tbox = target.tbox;
    T.vtx = target.vtx;
    T.tri = target.tri;
    T.normal = target.normal;
    classify_kernel<<<blocks, THREADS, 0, s>>>(a.pts, q ? q->leaf : nullptr, q ? q->vtx : nullptr, q ? q->tri : nullptr,
        a.begin, a.end, T, a.inside, a.perAxis, a.stats, a.overflowList, a.overflowCount, a.overflowCap);
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
