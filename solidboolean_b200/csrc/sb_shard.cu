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
//   tri_z      per triangle of both parents (one launch): box z range (AxisAlignedBoudingBox::update,
//              src/axisalignedboundingbox.h:31-41), centroid z exactly as the classification forms it
//              ((v0 + v1) + v2) / 3.0 (:497-499), histogram of the centroids, tallest box
//   plan       slab borders = quantiles of the joint centroid histogram (same integers on every rank)
//   select     count / scan / emit: order-preserving compaction of the triangles a rank needs
//   remap      hit pairs from selection numbers back to the parents' triangle ids
#include "sb_internal.h"
#include <algorithm>
#include <math.h>

namespace {

constexpr int ZBINS = 1024; // slab borders are bin edges: 1024 bins place them within 0.1 % of the height

__device__ __forceinline__ double bound_of(const unsigned long long *b, int k) { return dkey_inv(__ldg(b + k)); }

struct TwoMeshes { // the same pass over both parents in one launch (blockIdx below split: mesh 0)
    const float2 *zf[2]; // per vertex: z rounded down / up (written by bounds_pad)
    const uint32_t *tri[2];
    uint32_t nT[2], nV[2];
    const unsigned long long *bounds[2];
    float2 *zr[2];       // per triangle: box z range as floats rounded outwards (selection only)
    uint32_t *tiles[2];  // per selection tile: count, then exclusive start
    uint32_t *outTri[2]; // selection: index triples ...
    uint32_t *outFace[2]; // ... and the parent's triangle id
    uint32_t cap[2];
    uint32_t split;      // first block of mesh 1
};

__global__ void __launch_bounds__(256) tri_z_kernel(const __grid_constant__ TwoMeshes M, uint32_t *__restrict__ hist,
    unsigned long long *__restrict__ tallest)
{
    __shared__ uint32_t s_hist[ZBINS];
    for (int i = threadIdx.x; i < ZBINS; i += blockDim.x)
        s_hist[i] = 0;
    __syncthreads();
    const int k = blockIdx.x >= M.split ? 1 : 0;
    const uint32_t blk = blockIdx.x - (k ? M.split : 0u), nblk = k ? gridDim.x - M.split : M.split;
    const float2 *__restrict__ zf = M.zf[k];
    const uint32_t *__restrict__ tri = M.tri[k];
    const uint32_t nT = M.nT[k], nV = M.nV[k];
    // joint z range of both meshes (empty meshes leave their seeds: +-DBL_MAX the wrong way round)
    const double za = bound_of(M.bounds[0], 2), zb = bound_of(M.bounds[1], 2), ha = bound_of(M.bounds[0], 5), hb = bound_of(M.bounds[1], 5);
    const double z0 = fmin(za, zb), z1 = fmax(ha, hb);
    const double scale = (z1 > z0 && z1 - z0 < 1.0e300) ? (double)ZBINS / (z1 - z0) : 0.0;
    double tall = 0.0;
    for (uint32_t t = blk * blockDim.x + threadIdx.x; t < nT; t += nblk * blockDim.x) {
        uint32_t i0 = tri[3 * (size_t)t], i1 = tri[3 * (size_t)t + 1], i2 = tri[3 * (size_t)t + 2];
        if (i0 >= nV || i1 >= nV || i2 >= nV)
            i0 = i1 = i2 = 0; // reported by the build of the selection
        // float brackets of the three z: the box z range rounded outwards (all the selection needs); the
        // centroid only steers the balance of the slabs, the ownership test reads the exact one (scent)
        const float2 a = __ldg(zf + i0), b = __ldg(zf + i1), c = __ldg(zf + i2);
        float lo = FLT_MAX, hi = -FLT_MAX; // NaN never updates, like AxisAlignedBoudingBox::update
        if (a.y > hi) hi = a.y;
        if (a.x < lo) lo = a.x;
        if (b.y > hi) hi = b.y;
        if (b.x < lo) lo = b.x;
        if (c.y > hi) hi = c.y;
        if (c.x < lo) lo = c.x;
        M.zr[k][t] = make_float2(lo, hi);
        const double ext = (double)hi - (double)lo; // >= the exact height
        if (ext > tall)
            tall = ext;
        const double cz = ((double)a.x + (double)b.x + (double)c.x) * (1.0 / 3.0);
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
__global__ void __launch_bounds__(ZBINS) plan_kernel(const uint32_t *__restrict__ hist, const unsigned long long *__restrict__ boundsA,
    const unsigned long long *__restrict__ boundsB, const unsigned long long *__restrict__ tallest, int n, double *__restrict__ cuts)
{
    __shared__ uint32_t s_cum[ZBINS];
    __shared__ uint32_t s_part[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int PER = 1;
    uint32_t v[PER], sum = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        v[k] = hist[PER * tid + k];
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
    for (int k = 0; k < PER; ++k) {
        run += v[k];
        s_cum[PER * tid + k] = run; // inclusive
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

// the slab of `rank` widened by the margin, as floats rounded outwards (the selection may only be too large)
__device__ __forceinline__ float2 slab_of(const double *__restrict__ cuts, int rank, int n)
{
    const double m = cuts[n + 1];
    return make_float2(__double2float_rd(cuts[rank] - m), __double2float_ru(cuts[rank + 1] + m));
}

// box z range meets [lo, hi] (closed, like AxisAlignedBoudingBox::intersectWith; NaN never does)
__device__ __forceinline__ bool selected(const float2 zr, const float2 slab) { return zr.x <= slab.y && zr.y >= slab.x; }

__global__ void __launch_bounds__(SEL_THREADS) select_count_kernel(const __grid_constant__ TwoMeshes M, const double *__restrict__ cuts,
    int rank, int n)
{
    const int k = blockIdx.x >= M.split ? 1 : 0;
    const uint32_t blk = blockIdx.x - (k ? M.split : 0u);
    const float2 slab = slab_of(cuts, rank, n);
    uint32_t c = 0;
    const uint32_t base = blk * SEL_TILE + threadIdx.x * SEL_ITEMS;
#pragma unroll
    for (int i = 0; i < SEL_ITEMS; ++i)
        if (base + i < M.nT[k] && selected(M.zr[k][base + i], slab))
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
        M.tiles[k][blk] = t;
    }
}

// exclusive scan of the tile counts in place (CTA k: mesh k; a few thousand tiles at most), totals -> total[k];
// with caps given (speculative sizes of a repeated call), *mismatch is raised when a total is not the expected one
__global__ void __launch_bounds__(1024) select_scan_kernel(const __grid_constant__ TwoMeshes M, uint32_t *__restrict__ total,
    int checkCaps, uint32_t *__restrict__ mismatch)
{
    const int k = blockIdx.x;
    uint32_t *__restrict__ tileCount = M.tiles[k];
    const uint32_t tiles = (M.nT[k] + SEL_TILE - 1) / SEL_TILE;
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
    if (threadIdx.x == 0) {
        total[k] = s_carry;
        if (checkCaps && s_carry != M.cap[k])
            atomicOr(mismatch, 1u);
    }
}

__global__ void __launch_bounds__(SEL_THREADS) select_emit_kernel(const __grid_constant__ TwoMeshes M, const double *__restrict__ cuts,
    int rank, int n)
{
    const int k = blockIdx.x >= M.split ? 1 : 0;
    const uint32_t blk = blockIdx.x - (k ? M.split : 0u);
    const float2 slab = slab_of(cuts, rank, n);
    const uint32_t nT = M.nT[k], cap = M.cap[k];
    const uint32_t *__restrict__ tri = M.tri[k];
    const uint32_t base = blk * SEL_TILE + threadIdx.x * SEL_ITEMS;
    uint32_t mask = 0;
#pragma unroll
    for (int i = 0; i < SEL_ITEMS; ++i)
        if (base + i < nT && selected(M.zr[k][base + i], slab))
            mask |= 1u << i;
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
    uint32_t pos = M.tiles[k][blk] + incl - mine;
    for (int w = 0; w < warp; ++w)
        pos += s[w];
#pragma unroll
    for (int i = 0; i < SEL_ITEMS; ++i)
        if ((mask >> i) & 1u) {
            const uint32_t t = base + i;
            if (pos < cap) {
                M.outTri[k][3 * (size_t)pos] = tri[3 * (size_t)t];
                M.outTri[k][3 * (size_t)pos + 1] = tri[3 * (size_t)t + 1];
                M.outTri[k][3 * (size_t)pos + 2] = tri[3 * (size_t)t + 2];
                M.outFace[k][pos] = t;
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

TwoMeshes two(const MeshDev &A, const MeshDev &B, float2 *const zr[2], uint32_t *const tiles[2], float2 *const zf[2] = nullptr)
{
    TwoMeshes M = {};
    const MeshDev *m[2] = {&A, &B};
    for (int k = 0; k < 2; ++k) {
        M.zf[k] = zf ? zf[k] : nullptr;
        M.tri[k] = m[k]->tri;
        M.nT[k] = m[k]->nT;
        M.nV[k] = m[k]->nV;
        M.bounds[k] = m[k]->bounds;
        M.zr[k] = zr[k];
        M.tiles[k] = tiles[k];
    }
    return M;
}

} // namespace

size_t sbk_shard_tiles(uint32_t nT) { return ((size_t)nT + SEL_TILE - 1) / SEL_TILE; }
size_t sbk_shard_hist_words() { return ZBINS; }

// z ranges of every triangle of both parents, joint centroid histogram, tallest box; then the slab borders
cudaError_t sbk_shard_plan(cudaStream_t s, const MeshDev &A, const MeshDev &B, float2 *const zf[2], float2 *const zr[2],
    uint32_t *const tiles[2], uint32_t *hist /* ZBINS words + 2 for the tallest box, cleared here */, int n, double *cuts,
    int smCount, LaunchCounter &lc)
{
    TwoMeshes M = two(A, B, zr, tiles, zf);
    unsigned long long *tallest = reinterpret_cast<unsigned long long *>(hist + ZBINS);
    cudaMemsetAsync(hist, 0, 4 * ZBINS + 8, s);
    const uint32_t ba = (uint32_t)std::min<size_t>(((size_t)A.nT + 255) / 256, (size_t)smCount * 2);
    const uint32_t bb = (uint32_t)std::min<size_t>(((size_t)B.nT + 255) / 256, (size_t)smCount * 2);
    M.split = ba;
    if (ba + bb)
        tri_z_kernel<<<ba + bb, 256, 0, s>>>(M, hist, tallest);
    plan_kernel<<<1, ZBINS, 0, s>>>(hist, A.bounds, B.bounds, tallest, n, cuts);
    lc.kernels += (ba + bb) ? 2 : 1;
    return cudaGetLastError();
}

// sizes of this rank's two selections -> totals[0..1] (tile starts left in `tiles`)
cudaError_t sbk_shard_count(cudaStream_t s, const MeshDev &A, const MeshDev &B, float2 *const zr[2], uint32_t *const tiles[2],
    const double *cuts, int rank, int n, uint32_t *totals, const uint32_t *expect /* 2, or null */, uint32_t *mismatch,
    LaunchCounter &lc)
{
    TwoMeshes M = two(A, B, zr, tiles);
    const uint32_t ta = (uint32_t)sbk_shard_tiles(A.nT), tb = (uint32_t)sbk_shard_tiles(B.nT);
    M.split = ta;
    if (expect) {
        M.cap[0] = expect[0];
        M.cap[1] = expect[1];
    }
    if (ta + tb)
        select_count_kernel<<<ta + tb, SEL_THREADS, 0, s>>>(M, cuts, rank, n);
    select_scan_kernel<<<2, 1024, 0, s>>>(M, totals, expect ? 1 : 0, mismatch);
    lc.kernels += (ta + tb) ? 2 : 1;
    return cudaGetLastError();
}

cudaError_t sbk_shard_emit(cudaStream_t s, const MeshDev &A, const MeshDev &B, float2 *const zr[2], uint32_t *const tiles[2],
    const double *cuts, int rank, int n, const uint32_t cap[2], uint32_t *const outTri[2], uint32_t *const outFace[2],
    LaunchCounter &lc)
{
    TwoMeshes M = two(A, B, zr, tiles);
    const uint32_t ta = (uint32_t)sbk_shard_tiles(A.nT), tb = (uint32_t)sbk_shard_tiles(B.nT);
    M.split = ta;
    for (int k = 0; k < 2; ++k) {
        M.cap[k] = cap[k];
        M.outTri[k] = outTri[k];
        M.outFace[k] = outFace[k];
    }
    if (!(ta + tb))
        return cudaSuccess;
    select_emit_kernel<<<ta + tb, SEL_THREADS, 0, s>>>(M, cuts, rank, n);
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
