// Stage 4: inside/outside classification.  Replaces SolidBoolean::isPointInMesh
// (reference src/solidboolean.cpp:48-92) as driven by decideGroupSide (:482-510).
//
// Three kernels, so that the memory-latency-bound part runs at full occupancy
// and the FP64 part runs on a dense list with every lane busy:
//
//  A  ray_scan_kernel    one thread per RAY (point x axis; a CTA = 256
//     neighbouring points of one axis, so its threads read neighbouring grid
//     cells).  Ray box exactly as :53-58; the ray's cell list of the target's
//     axis-projected grid (sb_grid.cu) is filtered with the quantised 16-byte
//     references -- integer work only, which over-accepts slightly.  A warp scan
//     reserves one contiguous range of the global list per warp (one atomic per
//     warp) and every ray writes its (ray, triangle) entries ray-contiguously.
//  B  ray_hit_kernel     one thread per list entry, dense: first the EXACT
//     double box test that defines the reference's candidate set (ray box against
//     triangle boxes, :55-63), then segment/plane hit and the two edge-normal sign
//     tests (sb_raytri.cuh, bit-exact), PositionKey of the hit.
//  C  ray_finish_kernel  one thread per point: per axis, count the DISTINCT hit
//     keys (std::set<PositionKey>, :64/:85), odd = inside (:89); majority of the
//     three axes ((float)insideCount / totalCount > 0.5, :508).
#include "sb_internal.h"
#include "sb_raytri.cuh"

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int KEEP = 4; // candidates a scan thread keeps in registers before re-scanning

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
    uint32_t bigN0, bigN1, bigN2;
    const double2 *tbox;
    const double4 *vtx;
    const uint32_t *tri;
    const double *normal;
};

struct Query {
    const double *pts;   // explicit points (AoS), or null
    const double *scent; // faces mode: query mesh centroids in Morton order ...
    const Rec32 *leaf;   // ... and its leaves (sorted position -> triangle id)
    uint32_t nT;
    uint32_t begin;      // first point / sorted position
    uint32_t count;      // points in this launch
    // Lazy majority vote: the first pass traces axes 0 and 1 only (naxes = 2); the third
    // ray can change the result only where those two disagree, and is traced in a second
    // pass over just those points (list = their local indices, axis0 = 2, naxes = 1).
    const uint32_t *list;
    int axis0, naxes;
};

__device__ __forceinline__ uint32_t local_point(const Query &q, uint32_t j)
{
    return q.list ? __ldg(q.list + j) : j;
}

// Walk the ray's cell list(s) in a fixed order; visit(triangle id) for each
// triangle whose QUANTISED box overlaps the quantised ray box (a superset of the
// reference's candidates; ray_hit_kernel applies the exact test).
template <typename Visit>
__device__ __forceinline__ void for_each_candidate(const GridParams &g, const Target &T, int axis, const BoxD &myD,
    Visit &&visit)
{
    const BoxD meshBox = {g.lo[0], g.lo[1], g.lo[2], g.hi[0], g.hi[1], g.hi[2]};
    if (!overlap_d(meshBox, myD)) // no triangle box can overlap the ray box
        return;
    const int u = axis == 0 ? 1 : 0, v = axis == 2 ? 1 : 2;
    const uint32_t aU = quant16(comp(myD, u, false), g.org[u], g.scl[u]);
    const uint32_t bU = quant16(comp(myD, u, true), g.org[u], g.scl[u]);
    const uint32_t aV = quant16(comp(myD, v, false), g.org[v], g.scl[v]);
    const uint32_t bV = quant16(comp(myD, v, true), g.org[v], g.scl[v]);
    const uint32_t aA = quant16(comp(myD, axis, false), g.org[axis], g.scl[axis]);
    auto consider = [&](const uint4 &r) {
        // quantised closed-interval test (over-accepts only)
        if ((r.x & 0xffffu) > bU || (r.x >> 16) < aU || (r.y & 0xffffu) > bV || (r.y >> 16) < aV ||
            (r.z & 0xffffu) < aA)
            return;
        visit(r.w);
    };
    const int su = g.shiftU[axis], sv = g.shiftV[axis];
    const uint32_t cu0 = aU >> su, cu1 = bU >> su, cv0 = aV >> sv, cv1 = bV >> sv;
    // a ray box is a point widened by DBL_EPSILON: it almost always sits in ONE cell
    const bool multi = cu0 != cu1 || cv0 != cv1;
    for (uint32_t cv = cv0; cv <= cv1; ++cv)
        for (uint32_t cu = cu0; cu <= cu1; ++cu) {
            const uint32_t cell = g.cellBase[axis] + cv * g.nu[axis] + cu;
            const uint32_t i0 = __ldg(T.E + cell + 1), i1 = __ldg(T.E + cell + 2);
            // four independent 16-byte loads in flight per thread
            auto take = [&](const uint4 &r, uint32_t i) {
                // a triangle spanning several of the ray's cells is taken in the first one only
                if (i < i1 && (!multi || (max((r.x & 0xffffu) >> su, cu0) == cu && max((r.y & 0xffffu) >> sv, cv0) == cv)))
                    consider(r);
            };
            for (uint32_t i = i0; i < i1; i += 4) {
                const uint4 r0 = __ldg(T.refs + i);
                const uint4 r1 = __ldg(T.refs + min(i + 1, i1 - 1));
                const uint4 r2 = __ldg(T.refs + min(i + 2, i1 - 1));
                const uint4 r3 = __ldg(T.refs + min(i + 3, i1 - 1));
                take(r0, i);
                take(r1, i + 1);
                take(r2, i + 2);
                take(r3, i + 3);
            }
        }
    const uint32_t nBig = axis == 0 ? T.bigN0 : axis == 1 ? T.bigN1 : T.bigN2;
    for (uint32_t i = 0; i < nBig; ++i)
        consider(__ldg(T.bigRefs + (size_t)axis * T.bigCap + i));
}

__device__ __forceinline__ bool query_point(const Query &q, uint32_t j, d3 &p, uint32_t &outIndex)
{
    if (j >= q.count)
        return false;
    const uint32_t idx = q.begin + local_point(q, j);
    if (q.pts) {
        p = {q.pts[3 * (size_t)idx], q.pts[3 * (size_t)idx + 1], q.pts[3 * (size_t)idx + 2]};
        outIndex = idx;
        return true;
    }
    // faces mode: the centroids ((v0 + v1) + v2) / 3.0 (src/solidboolean.cpp:497-499) were
    // formed at build time and stored in Morton order; positions >= nT are padding
    if (idx >= q.nT)
        return false;
    outIndex = idx;
    p = {__ldg(q.scent + 3 * (size_t)idx), __ldg(q.scent + 3 * (size_t)idx + 1), __ldg(q.scent + 3 * (size_t)idx + 2)};
    return true;
}

// ---- A ------------------------------------------------------------------------
__global__ void __launch_bounds__(SCAN_THREADS) ray_scan_kernel(Query q, Target T, uint32_t blocksPerAxis,
    uint2 *__restrict__ cand, unsigned long long cap,
    unsigned long long *__restrict__ candCount, uint2 *__restrict__ rayRange)
{
    __shared__ GridParams g;
    if (threadIdx.x == 0)
        g = *T.gp;
    __syncthreads();
    const int slot = blockIdx.x / blocksPerAxis;
    const int axis = q.axis0 + slot;
    const uint32_t j = (blockIdx.x % blocksPerAxis) * SCAN_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31;

    d3 p = {0, 0, 0};
    uint32_t outIndex = 0;
    const bool active = query_point(q, j, p, outIndex);
    const d3 e = ray_end(p, axis);
    const BoxD myD = ray_box(p, e);

    uint32_t n = 0, c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    if (active)
        for_each_candidate(g, T, axis, myD, [&](uint32_t f) {
            if (n == 0) c0 = f;
            else if (n == 1) c1 = f;
            else if (n == 2) c2 = f;
            else if (n == 3) c3 = f;
            ++n;
        });

    // warp-wide exclusive scan of n; one atomic per warp reserves its output range
    // (no block barrier: a warp whose rays hit long cell lists does not hold up the rest)
    uint32_t incl = n;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        uint32_t t = __shfl_up_sync(SB_FULL, incl, off);
        if (lane >= off)
            incl += t;
    }
    const uint32_t total = __shfl_sync(SB_FULL, incl, 31);
    unsigned long long base = 0;
    if (lane == 0 && total)
        base = atomicAdd(candCount, (unsigned long long)total);
    base = __shfl_sync(SB_FULL, base, 0);
    if (!active)
        return;
    const unsigned long long first = base + incl - n;
    const uint32_t ray = (uint32_t)slot * q.count + j;
    rayRange[ray] = make_uint2((uint32_t)min(first, 0xffffffffull), n);
    if (first + n > cap)
        return; // list too small: the host sees candCount > cap and retries
    if (n > 0) cand[first] = make_uint2(ray, c0);
    if (n > 1) cand[first + 1] = make_uint2(ray, c1);
    if (n > 2) cand[first + 2] = make_uint2(ray, c2);
    if (n > 3) cand[first + 3] = make_uint2(ray, c3);
    if (n > KEEP) {
        uint32_t k = 0;
        for_each_candidate(g, T, axis, myD, [&](uint32_t f) {
            if (k >= KEEP)
                cand[first + k] = make_uint2(ray, f);
            ++k;
        });
    }
}

// ---- B ------------------------------------------------------------------------
__global__ void __launch_bounds__(128) ray_hit_kernel(Query q, Target T,
    const uint2 *__restrict__ cand, unsigned long long cap, const unsigned long long *__restrict__ candCount,
    uint32_t nTargetTris, long long *__restrict__ keys /* 3 per entry */, uint8_t *__restrict__ hitFlag,
    unsigned long long *__restrict__ exactCount)
{
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long total = min(__ldg(candCount), cap);
    bool isCand = false, h = false;
    d3 hit = {0, 0, 0};
    if (i < total) {
        const uint2 c = __ldg(cand + i);
        // when the list overflowed (the host will retry) some entries below cap were never written
        if (c.x < (uint32_t)q.naxes * q.count && c.y < nTargetTris) {
            const int axis = q.axis0 + (int)(c.x / q.count);
            const uint32_t j = c.x % q.count;
            const double *src = (q.pts ? q.pts : q.scent) + 3 * (size_t)(q.begin + local_point(q, j));
            const d3 p = {__ldg(src), __ldg(src + 1), __ldg(src + 2)};
            const d3 e = ray_end(p, axis);
            const uint32_t f = c.y;
            // the reference's candidate test: triangle box .intersectWith(ray box), exact doubles
            isCand = overlap_d(load_boxd(T.tbox + 3 * (size_t)f), ray_box(p, e));
            if (isCand) {
                d3 t0 = load_vertex(T.vtx, __ldg(T.tri + 3 * (size_t)f));
                d3 t1 = load_vertex(T.vtx, __ldg(T.tri + 3 * (size_t)f + 1));
                d3 t2 = load_vertex(T.vtx, __ldg(T.tri + 3 * (size_t)f + 2));
                d3 nrm = {__ldg(T.normal + 3 * (size_t)f), __ldg(T.normal + 3 * (size_t)f + 1),
                          __ldg(T.normal + 3 * (size_t)f + 2)};
                h = ray_tri_hit_filtered(p, e, t0, t1, t2, nrm, hit);
            }
        }
        hitFlag[i] = h ? 1 : 0;
        if (h) {
            keys[3 * i] = position_key(hit.x);
            keys[3 * i + 1] = position_key(hit.y);
            keys[3 * i + 2] = position_key(hit.z);
        }
    }
    // exact candidate count (roofline accounting): one atomic per warp
    const uint32_t m = __ballot_sync(SB_FULL, isCand);
    if ((threadIdx.x & 31) == 0 && m)
        atomicAdd(exactCount, (unsigned long long)__popc(m));
}


// ---- C ------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ray_finish_kernel(Query q, const uint2 *__restrict__ rayRange,
    unsigned long long cap, const long long *__restrict__ keys, const uint8_t *__restrict__ hitFlag,
    uint8_t *__restrict__ inside, uint8_t *__restrict__ perAxis, uint32_t *__restrict__ undecided,
    unsigned int *__restrict__ undecidedCount)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= q.count)
        return;
    const uint32_t local = local_point(q, j);
    uint32_t outIndex;
    if (q.pts) {
        outIndex = q.begin + local;
    } else {
        if (q.begin + local >= q.nT)
            return;
        outIndex = (uint32_t)load_rec(q.leaf + q.begin + local).ref;
    }
    int insideCount = 0;
    bool first = false;
    for (int slot = 0; slot < q.naxes; ++slot) {
        const uint2 r = __ldg(rayRange + (size_t)slot * q.count + j);
        uint32_t distinct = 0;
        if (r.y && (unsigned long long)r.x + r.y <= cap) {
            uint32_t hits = 0;
            for (uint32_t a = 0; a < r.y; ++a)
                hits += hitFlag[(size_t)r.x + a];
            distinct = hits;
            if (hits > 1) { // std::set<PositionKey>: equal keys count once
                distinct = 0;
                for (uint32_t a = 0; a < r.y; ++a) {
                    const size_t ia = (size_t)r.x + a;
                    if (!hitFlag[ia])
                        continue;
                    const long long kx = keys[3 * ia], ky = keys[3 * ia + 1], kz = keys[3 * ia + 2];
                    bool dup = false;
                    for (uint32_t b = 0; b < a && !dup; ++b) {
                        const size_t ib = (size_t)r.x + b;
                        dup = hitFlag[ib] && keys[3 * ib] == kx && keys[3 * ib + 1] == ky && keys[3 * ib + 2] == kz;
                    }
                    distinct += dup ? 0u : 1u;
                }
            }
        }
        const bool in = (distinct & 1u) != 0; // odd number of distinct crossings (:89)
        if (perAxis)
            perAxis[3 * (size_t)outIndex + q.axis0 + slot] = in ? 1 : 0;
        insideCount += in ? 1 : 0;
        if (slot == 0)
            first = in;
    }
    if (q.naxes == 3) {
        // (float)insideCount / totalCount > 0.5 with totalCount == 3 (:508)
        inside[outIndex] = insideCount >= 2 ? 1 : 0;
    } else if (q.naxes == 2) {
        if (insideCount != 1) {
            inside[outIndex] = first ? 1 : 0; // two equal votes already are the majority
        } else {
            const unsigned int slot = atomicAdd(undecidedCount, 1u);
            undecided[slot] = local;         // the third ray decides
        }
    } else {
        inside[outIndex] = insideCount ? 1 : 0; // votes 0 and 1 disagreed: the majority is vote 2
    }
}

} // namespace

size_t sbk_classify_scratch_bytes(uint32_t points, unsigned long long cap, int naxes)
{
    size_t b = 0;
    b += ((size_t)cap * 8 + 255) & ~(size_t)255;               // cand
    b += ((size_t)cap * 24 + 255) & ~(size_t)255;              // keys
    b += ((size_t)cap + 255) & ~(size_t)255;                   // hit flags
    b += ((size_t)points * naxes * 8 + 255) & ~(size_t)255;    // rayRange
    b += ((size_t)points * 4 + 255) & ~(size_t)255;            // undecided list
    return b + 256;
}

uint32_t *sbk_classify_undecided_list(void *scratch, uint32_t points, unsigned long long cap, int naxes)
{
    size_t b = 0;
    b += ((size_t)cap * 8 + 255) & ~(size_t)255;
    b += ((size_t)cap * 24 + 255) & ~(size_t)255;
    b += ((size_t)cap + 255) & ~(size_t)255;
    b += ((size_t)points * naxes * 8 + 255) & ~(size_t)255;
    return reinterpret_cast<uint32_t *>(static_cast<char *>(scratch) + b);
}

cudaError_t sbk_classify(cudaStream_t s, const MeshDev &target, const ClassifyArgs &a, const ClassifyPass &pass,
    void *scratch, unsigned long long cap, unsigned long long *candCount, unsigned long long *exactCount,
    unsigned int *undecidedCount, LaunchCounter &lc)
{
    if (a.end <= a.begin)
        return cudaSuccess;
    const MeshDev *qm = a.queryMesh;
    Query q;
    q.pts = a.pts;
    q.scent = qm ? qm->scent : nullptr;
    q.leaf = qm ? qm->leaf : nullptr;
    q.nT = qm ? qm->nT : 0;
    q.begin = a.begin;
    q.count = pass.list ? pass.listCount : a.end - a.begin;
    q.list = pass.list;
    q.axis0 = pass.axis0;
    q.naxes = pass.naxes;
    if (q.count == 0)
        return cudaSuccess;
    Target T;
    T.gp = target.gridParams;
    T.E = target.gridE;
    T.refs = target.gridRefs;
    T.bigRefs = target.gridBigRefs;
    T.bigCap = target.gridBigCap;
    T.bigN0 = target.gridBigN[0];
    T.bigN1 = target.gridBigN[1];
    T.bigN2 = target.gridBigN[2];
    T.tbox = target.tbox;
    T.vtx = target.vtx;
    T.tri = target.tri;
    T.normal = target.normal;

    char *b = static_cast<char *>(scratch);
    auto take = [&](size_t bytes) {
        char *p = b;
        b += (bytes + 255) & ~(size_t)255;
        return p;
    };
    uint2 *cand = reinterpret_cast<uint2 *>(take((size_t)cap * 8));
    long long *keys = reinterpret_cast<long long *>(take((size_t)cap * 24));
    uint8_t *hitFlag = reinterpret_cast<uint8_t *>(take((size_t)cap));
    uint2 *rayRange = reinterpret_cast<uint2 *>(take((size_t)q.count * q.naxes * 8));
    uint32_t *undecided = reinterpret_cast<uint32_t *>(take((size_t)q.count * 4));

    const uint32_t bpa = (q.count + SCAN_THREADS - 1) / SCAN_THREADS;
    ray_scan_kernel<<<q.naxes * bpa, SCAN_THREADS, 0, s>>>(q, T, bpa, cand, cap, candCount, rayRange);
    const unsigned long long hitBlocks = (cap + 127) / 128;
    ray_hit_kernel<<<(unsigned)hitBlocks, 128, 0, s>>>(q, T, cand, cap, candCount, target.nT, keys, hitFlag, exactCount);
    ray_finish_kernel<<<(q.count + 255) / 256, 256, 0, s>>>(q, rayRange, cap, keys, hitFlag, a.inside, a.perAxis, undecided,
        undecidedCount);
    lc.kernels += 3;
    return cudaGetLastError();
}
