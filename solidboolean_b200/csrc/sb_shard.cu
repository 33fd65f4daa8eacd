// Multi-GPU shards.  The reference has no notion of it (one SolidBoolean per call,
// src/solidboolean.cpp:288); what shards naturally is the QUERY side of every stage: candidate
// search and predicate per triangle of A, isPointInMesh (:48-92) per face centroid.
//
// A rank does not build the whole of both meshes for its share of the queries.  The two lazy test
// rays run along x and y (g_testAxisList, :31-35), so a slab  z in [lo, hi)  is closed under them:
// a query point of the slab only ever meets triangles whose box reaches into the slab, and a
// triangle of A whose centroid lies in the slab only overlaps boxes that reach within its own
// height of it.  Rank r therefore takes the faces whose centroid z falls into the r-th of n
// equally populated slabs as ITS queries, selects from both meshes the triangles whose box touches
// the slab widened by the tallest triangle box, and builds its acceleration structures over that
// selection only (an ordinary mesh that shares the parent's vertices and maps its triangle
// numbers back).  The (rare) third ray along z needs the whole target: sb_capi.cu falls back to
// the full meshes for the points whose first two votes disagree.
//
//   tri_z      per triangle of a parent: exact box z range (AxisAlignedBoudingBox::update,
//              src/axisalignedboundingbox.h:31-41), centroid z exactly as the classification forms it
//              ((v0 + v1) + v2) / 3.0 (:497-499), histogram of the centroids, tallest box
//   plan       slab borders = quantiles of the joint centroid histogram (same integers on every rank)
//   select     count / scan / emit: order-preserving compaction of the triangles a rank needs
//   remap      hit pairs from selection numbers back to the parents' triangle ids
#include "sb_internal.h"
#include <algorithm>
#include <math.h>

namespace {

constexpr int ZBINS = 4096;

__device__ __forceinline__ double bound_of(const unsigned long long *b, int k) { return dkey_inv(__ldg(b + k)); }

__global__ void __launch_bounds__(256) tri_z_kernel(const double4 *__restrict__ vtx, const uint32_t *__restrict__ tri, uint32_t nT,
    uint32_t nV, const unsigned long long *__restrict__ boundsA, const unsigned long long *__restrict__ boundsB,
    double *__restrict__ zinfo /* 3 nT: lo, hi, centroid */, uint32_t *__restrict__ hist, unsigned long long *__restrict__ tallest)
{
    __shared__ uint32_t s_hist[ZBINS];
    for (int i = threadIdx.x; i < ZBINS; i += blockDim.x)
        s_hist[i] = 0;
    __syncthreads();
    // joint z range of both meshes (empty meshes leave their seeds: +-DBL_MAX the wrong way round)
    const double za = bound_of(boundsA, 2), zb = bound_of(boundsB, 2), ha = bound_of(boundsA, 5), hb = bound_of(boundsB, 5);
    const double z0 = fmin(za, zb), z1 = fmax(ha, hb);
    const double scale = (z1 > z0 && z1 - z0 < 1.0e300) ? (double)ZBINS / (z1 - z0) : 0.0;
    double tall = 0.0;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < nT; t += gridDim.x * blockDim.x) {
        uint32_t i0 = tri[3 * (size_t)t], i1 = tri[3 * (size_t)t + 1], i2 = tri[3 * (size_t)t + 2];
        if (i0 >= nV || i1 >= nV || i2 >= nV)
            i0 = i1 = i2 = 0; // reported by the build of the selection
        const double a = load_vertex(vtx, i0).z, b = load_vertex(vtx, i1).z, c = load_vertex(vtx, i2).z;
        double lo = DBL_MAX, hi = -DBL_MAX;
        if (a > hi) hi = a;
        if (a < lo) lo = a;
        if (b > hi) hi = b;
        if (b < lo) lo = b;
        if (c > hi) hi = c;
        if (c < lo) lo = c;
        const double cz = xdiv(xadd(xadd(a, b), c), 3.0);
        zinfo[3 * (size_t)t] = lo;
        zinfo[3 * (size_t)t + 1] = hi;
        zinfo[3 * (size_t)t + 2] = cz;
        if (hi - lo > tall)
            tall = hi - lo;
        double bin = floor((cz - z0) * scale);
        bin = fmin(fmax(bin, 0.0), (double)(ZBINS - 1)); // NaN -> 0
        atomicAdd(&s_hist[(int)bin], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ZBINS; i += blockDim.x)
        if (s_hist[i])
            atomicAdd(&hist[i], s_hist[i]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
        tall = fmax(tall, __shfl_xor_sync(SB_FULL, tall, off));
    if ((threadIdx.x & 31) == 0 && tall > 0.0)
        atomicMax(tallest, dkey(tall));
}

// One CTA.  cuts[0] = -inf, cuts[n] = +inf, cuts[k] = upper edge of the bin where the running count
// reaches k / n of the faces; cuts[n + 1] = the margin (tallest triangle box of either mesh, a bit more).
__global__ void __launch_bounds__(1024) plan_kernel(const uint32_t *__restrict__ hist, const unsigned long long *__restrict__ boundsA,
    const unsigned long long *__restrict__ boundsB, const unsigned long long *__restrict__ tallest, int n, double *__restrict__ cuts)
{
    __shared__ uint32_t s_cum[ZBINS];
    __shared__ uint32_t s_part[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t v[4], sum = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        v[k] = hist[4 * tid + k];
        sum += v[k];
    }
    uint32_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(SB_FULL, incl, d);
        if (lane >= d)
            incl += t;
    }
    if (lane == 31)
        s_part[warp] = incl;
    __syncthreads();
    uint32_t base = 0;
    for (int w = 0; w < warp; ++w)
        base += s_part[w];
    uint32_t run = base + incl - sum;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        run += v[k];
        s_cum[4 * tid + k] = run; // inclusive
    }
    __syncthreads();
    const uint32_t total = s_cum[ZBINS - 1];
    const double za = bound_of(boundsA, 2), zb = bound_of(boundsB, 2), ha = bound_of(boundsA, 5), hb = bound_of(boundsB, 5);
    const double z0 = fmin(za, zb), z1 = fmax(ha, hb);
    const double width = (z1 > z0 && z1 - z0 < 1.0e300) ? (z1 - z0) / (double)ZBINS : 0.0;
    if (tid <= n) {
        double c;
        if (tid == 0) {
            c = -INFINITY;
        } else if (tid == n) {
            c = INFINITY;
        } else {
            const unsigned long long want = ((unsigned long long)total * (unsigned)tid + (unsigned)n - 1) / (unsigned)n;
            int lo = 0, hi = ZBINS - 1; // first bin whose inclusive count reaches `want`
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (s_cum[mid] >= want)
                    hi = mid;
                else
                    lo = mid + 1;
            }
            c = z0 + width * (double)(lo + 1);
        }
        cuts[tid] = c;
    }
    if (tid == 0) {
        const double tall = dkey_inv(*tallest);
        // + the DBL_EPSILON by which a ray box reaches beyond its point (src/solidboolean.cpp:31-35, :53)
        cuts[n + 1] = tall * 1.0000001 + 8.0 * 2.2204460492503131e-16 * fmax(1.0, fmax(fabs(z0), fabs(z1)));
    }
}

constexpr int SEL_THREADS = 256;
constexpr int SEL_ITEMS = 8;
constexpr int SEL_TILE = SEL_THREADS * SEL_ITEMS;

__device__ __forceinline__ bool selected(const double *__restrict__ zinfo, uint32_t t, double lo, double hi)
{
    // box z range meets [lo, hi] (closed, like AxisAlignedBoudingBox::intersectWith; NaN never does)
    return zinfo[3 * (size_t)t] <= hi && zinfo[3 * (size_t)t + 1] >= lo;
}

__global__ void __launch_bounds__(SEL_THREADS) select_count_kernel(const double *__restrict__ zinfo, uint32_t nT,
    const double *__restrict__ cuts, int rank, int n, uint32_t *__restrict__ tileCount)
{
    const double m = cuts[n + 1], lo = cuts[rank] - m, hi = cuts[rank + 1] + m;
    uint32_t c = 0;
    const uint32_t base = blockIdx.x * SEL_TILE + threadIdx.x * SEL_ITEMS;
#pragma unroll
    for (int k = 0; k < SEL_ITEMS; ++k)
        if (base + k < nT && selected(zinfo, base + k, lo, hi))
            ++c;
    c = __reduce_add_sync(SB_FULL, c);
    __shared__ uint32_t s[SEL_THREADS / 32];
    if ((threadIdx.x & 31) == 0)
        s[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < SEL_THREADS / 32; ++w)
            t += s[w];
        tileCount[blockIdx.x] = t;
    }
}

// exclusive scan of the tile counts in place (one CTA; tiles <= a few thousand), total -> *total
__global__ void __launch_bounds__(1024) select_scan_kernel(uint32_t *__restrict__ tileCount, uint32_t tiles, uint32_t *__restrict__ total)
{
    __shared__ uint32_t s_part[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0)
        s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t b = 0; b < tiles; b += 1024) {
        const uint32_t i = b + threadIdx.x;
        const uint32_t v = i < tiles ? tileCount[i] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(SB_FULL, incl, d);
            if (lane >= d)
                incl += t;
        }
        if (lane == 31)
            s_part[warp] = incl;
        __syncthreads();
        uint32_t base = s_carry;
        for (int w = 0; w < warp; ++w)
            base += s_part[w];
        if (i < tiles)
            tileCount[i] = base + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023)
            s_carry = base + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0)
        *total = s_carry;
}

__global__ void __launch_bounds__(SEL_THREADS) select_emit_kernel(const double *__restrict__ zinfo, const uint32_t *__restrict__ tri,
    uint32_t nT, const double *__restrict__ cuts, int rank, int n, const uint32_t *__restrict__ tileStart, uint32_t cap,
    uint32_t *__restrict__ outTri, uint32_t *__restrict__ outFace)
{
    const double m = cuts[n + 1], lo = cuts[rank] - m, hi = cuts[rank + 1] + m;
    const uint32_t base = blockIdx.x * SEL_TILE + threadIdx.x * SEL_ITEMS;
    uint32_t mask = 0;
#pragma unroll
    for (int k = 0; k < SEL_ITEMS; ++k)
        if (base + k < nT && selected(zinfo, base + k, lo, hi))
            mask |= 1u << k;
    const uint32_t mine = __popc(mask);
    uint32_t incl = mine;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(SB_FULL, incl, d);
        if (lane >= d)
            incl += t;
    }
    __shared__ uint32_t s[SEL_THREADS / 32];
    if (lane == 31)
        s[warp] = incl;
    __syncthreads();
    uint32_t pos = tileStart[blockIdx.x] + incl - mine;
    for (int w = 0; w < warp; ++w)
        pos += s[w];
#pragma unroll
    for (int k = 0; k < SEL_ITEMS; ++k)
        if ((mask >> k) & 1u) {
            const uint32_t t = base + k;
            if (pos < cap) {
                outTri[3 * (size_t)pos] = tri[3 * (size_t)t];
                outTri[3 * (size_t)pos + 1] = tri[3 * (size_t)t + 1];
                outTri[3 * (size_t)pos + 2] = tri[3 * (size_t)t + 2];
                outFace[pos] = t;
            }
            ++pos;
        }
}

__global__ void __launch_bounds__(256) remap_hits_kernel(uint32_t *__restrict__ ab, uint32_t n, const uint32_t *__restrict__ faceA,
    const uint32_t *__restrict__ faceB)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    ab[2 * (size_t)i] = __ldg(faceA + ab[2 * (size_t)i]);
    ab[2 * (size_t)i + 1] = __ldg(faceB + ab[2 * (size_t)i + 1]);
}

} // namespace

size_t sbk_shard_tiles(uint32_t nT) { return ((size_t)nT + SEL_TILE - 1) / SEL_TILE; }
size_t sbk_shard_hist_words() { return ZBINS; }

cudaError_t sbk_shard_tri_z(cudaStream_t s, const MeshDev &m, const unsigned long long *boundsA, const unsigned long long *boundsB,
    double *zinfo, uint32_t *hist, unsigned long long *tallest, int smCount, LaunchCounter &lc)
{
    if (m.nT == 0)
        return cudaSuccess;
    int blocks = (int)std::min<size_t>(((size_t)m.nT + 255) / 256, (size_t)smCount * 8);
    tri_z_kernel<<<blocks, 256, 0, s>>>(m.vtx, m.tri, m.nT, m.nV, boundsA, boundsB, zinfo, hist, tallest);
    lc.kernels += 1;
    return cudaGetLastError();
}

cudaError_t sbk_shard_plan(cudaStream_t s, const uint32_t *hist, const unsigned long long *boundsA, const unsigned long long *boundsB,
    const unsigned long long *tallest, int n, double *cuts, LaunchCounter &lc)
{
    plan_kernel<<<1, 1024, 0, s>>>(hist, boundsA, boundsB, tallest, n, cuts);
    lc.kernels += 1;
    return cudaGetLastError();
}

cudaError_t sbk_shard_count(cudaStream_t s, const double *zinfo, uint32_t nT, const double *cuts, int rank, int n, uint32_t *tileCount,
    uint32_t *total, LaunchCounter &lc)
{
    const uint32_t tiles = (uint32_t)sbk_shard_tiles(nT);
    if (tiles)
        select_count_kernel<<<tiles, SEL_THREADS, 0, s>>>(zinfo, nT, cuts, rank, n, tileCount);
    select_scan_kernel<<<1, 1024, 0, s>>>(tileCount, tiles, total);
    lc.kernels += tiles ? 2 : 1;
    return cudaGetLastError();
}

cudaError_t sbk_shard_emit(cudaStream_t s, const double *zinfo, const uint32_t *tri, uint32_t nT, const double *cuts, int rank, int n,
    const uint32_t *tileStart, uint32_t cap, uint32_t *outTri, uint32_t *outFace, LaunchCounter &lc)
{
    const uint32_t tiles = (uint32_t)sbk_shard_tiles(nT);
    if (!tiles)
        return cudaSuccess;
    select_emit_kernel<<<tiles, SEL_THREADS, 0, s>>>(zinfo, tri, nT, cuts, rank, n, tileStart, cap, outTri, outFace);
    lc.kernels += 1;
    return cudaGetLastError();
}

cudaError_t sbk_shard_remap_hits(cudaStream_t s, uint32_t *ab, uint32_t n, const uint32_t *faceA, const uint32_t *faceB, LaunchCounter &lc)
{
    if (!n)
        return cudaSuccess;
    remap_hits_kernel<<<(n + 255) / 256, 256, 0, s>>>(ab, n, faceA, faceB);
    lc.kernels += 1;
    return cudaGetLastError();
}
