// Per-triangle intersection contexts on the device.
//
// Replaces the body of the pair loop of SolidBoolean::combine (reference
// src/solidboolean.cpp:296-339): for every intersecting pair, in the order of the pair list
// (here: ascending (a, b), the order sb_intersect delivers), the two end points of the segment
// are entered into the context of the first mesh's triangle and into that of the second mesh's
// triangle.  A context de-duplicates its points by PositionKey (addIntersectedPoint :305-310:
// the FIRST position seen with a key is kept, numbered in first-seen order) and keeps the
// undirected relation between the two point numbers of a segment unless they coincide
// (:323-328, :332-337).  The reference does this with a std::map and two hash containers per
// triangle; here the hits of one side are a segmented array (contexts = runs of equal triangle
// id) and every step is a rank or a short scan inside the run:
//   cuts_heads     triangle of every hit in side order, run heads
//   (rank)         run index of every hit  -> context list, run starts
//   cuts_keys      PositionKey of the 2 n end points
//   cuts_rep       first end point of the run with the same key (short backward scan)
//   (rank)         global number of every first-seen point -> CSR of points
//   cuts_edges     local point numbers of every segment, first occurrence of its relation
//   (rank)         -> CSR of relations;  cuts_emit writes points and relations, the latter at
//                  their rank in ascending (low, high) order within the context
// Side 1 first orders the hit indices by the second triangle with a stable radix sort (inside
// a run the hits keep their ascending first-triangle order = the order the loop meets them).
#include "sb_internal.h"
#include "sb_radix.cuh"

namespace {

constexpr int RK_THREADS = 256;
constexpr int RK_ITEMS = 8;
constexpr int RK_TILE = RK_THREADS * RK_ITEMS;

// ---- exclusive rank of the set flags of a byte array: count -> scan -> rank ----------------
__global__ void __launch_bounds__(RK_THREADS) rank_count_kernel(const uint8_t *__restrict__ flag, uint32_t n,
    uint32_t *__restrict__ tileCount)
{
    __shared__ uint32_t s_warp[RK_THREADS / 32];
    uint32_t c = 0;
#pragma unroll
    for (int s = 0; s < RK_ITEMS; ++s) {
        const uint32_t i = blockIdx.x * RK_TILE + s * RK_THREADS + threadIdx.x;
        c += (i < n && flag[i]) ? 1u : 0u;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
        c += __shfl_xor_sync(SB_FULL, c, off);
    if ((threadIdx.x & 31) == 0)
        s_warp[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < RK_THREADS / 32; ++w)
            t += s_warp[w];
        tileCount[blockIdx.x] = t;
    }
}

// exclusive scan of the tile counts in place (one CTA), grand total to *total
__global__ void __launch_bounds__(1024) rank_scan_kernel(uint32_t *__restrict__ tileCount, uint32_t tiles,
    uint32_t *__restrict__ total)
{
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0)
        s_carry = 0;
    __syncthreads();
    for (uint32_t begin = 0; begin < tiles; begin += 1024) {
        const uint32_t i = begin + threadIdx.x;
        const uint32_t v = i < tiles ? tileCount[i] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t t = __shfl_up_sync(SB_FULL, incl, off);
            if (lane >= off)
                incl += t;
        }
        if (lane == 31)
            s_warp[warp] = incl;
        __syncthreads();
        uint32_t warpOff = 0, chunk = 0;
        for (int w = 0; w < 32; ++w) {
            if (w < warp)
                warpOff += s_warp[w];
            chunk += s_warp[w];
        }
        const uint32_t carry = s_carry;
        if (i < tiles)
            tileCount[i] = carry + warpOff + incl - v;
        __syncthreads();
        if (threadIdx.x == 0)
            s_carry = carry + chunk;
        __syncthreads();
    }
    if (threadIdx.x == 0)
        *total = s_carry;
}

// rank[i] = number of set flags before i (striped items: coalesced, ranks in index order)
__global__ void __launch_bounds__(RK_THREADS) rank_kernel(const uint8_t *__restrict__ flag, uint32_t n,
    const uint32_t *__restrict__ tileOffset, uint32_t *__restrict__ rank)
{
    constexpr int WARPS = RK_THREADS / 32;
    __shared__ uint32_t s_cnt[RK_ITEMS * WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t before[RK_ITEMS];
#pragma unroll
    for (int s = 0; s < RK_ITEMS; ++s) {
        const uint32_t i = blockIdx.x * RK_TILE + s * RK_THREADS + threadIdx.x;
        const unsigned b = __ballot_sync(SB_FULL, i < n && flag[i]);
        before[s] = __popc(b & lanemask_lt());
        if (lane == 0)
            s_cnt[s * WARPS + warp] = __popc(b);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const uint32_t c0 = s_cnt[2 * lane], c1 = s_cnt[2 * lane + 1];
        uint32_t incl = c0 + c1;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t t = __shfl_up_sync(SB_FULL, incl, off);
            if (lane >= off)
                incl += t;
        }
        s_cnt[2 * lane] = incl - c0 - c1;
        s_cnt[2 * lane + 1] = incl - c1;
    }
    __syncthreads();
    const uint32_t base = tileOffset[blockIdx.x];
#pragma unroll
    for (int s = 0; s < RK_ITEMS; ++s) {
        const uint32_t i = blockIdx.x * RK_TILE + s * RK_THREADS + threadIdx.x;
        if (i < n)
            rank[i] = base + s_cnt[s * WARPS + warp] + before[s];
    }
}
static_assert(RK_ITEMS * (RK_THREADS / 32) == 64, "the 64 segment counts are scanned by one warp, two per lane");

cudaError_t exclusive_rank(cudaStream_t s, const uint8_t *flag, uint32_t n, uint32_t *tileScratch, uint32_t *rank,
    uint32_t *total, LaunchCounter &lc)
{
    const uint32_t tiles = (n + RK_TILE - 1) / RK_TILE;
    rank_count_kernel<<<tiles, RK_THREADS, 0, s>>>(flag, n, tileScratch);
    rank_scan_kernel<<<1, 1024, 0, s>>>(tileScratch, tiles, total);
    rank_kernel<<<tiles, RK_THREADS, 0, s>>>(flag, n, tileScratch, rank);
    lc.kernels += 3;
    return cudaGetLastError();
}

// ---- the stage --------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cuts_iota_kernel(const uint32_t *__restrict__ hitAB, uint32_t n, uint32_t *__restrict__ key,
    uint32_t *__restrict__ val)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) {
        key[j] = __ldg(hitAB + 2 * (size_t)j + 1);
        val[j] = j;
    }
}

// order == null: side order = hit order
__global__ void __launch_bounds__(256) cuts_heads_kernel(const uint32_t *__restrict__ hitAB, const uint32_t *__restrict__ order,
    uint32_t n, int which, uint32_t *__restrict__ tri, uint8_t *__restrict__ head)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n)
        return;
    const uint32_t t = __ldg(hitAB + 2 * (size_t)(order ? __ldg(order + j) : j) + which);
    tri[j] = t;
    head[j] = (j == 0 || __ldg(hitAB + 2 * (size_t)(order ? __ldg(order + j - 1) : j - 1) + which) != t) ? 1 : 0;
}

// the rank counts the heads BEFORE j; the run of j includes its own head
__global__ void __launch_bounds__(256) cuts_fix_kernel(const uint8_t *__restrict__ head, uint32_t *__restrict__ runOf, uint32_t n)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n)
        runOf[j] = runOf[j] + head[j] - 1u;
}

// closing entries, placed on the device (the counts are only known there)
__global__ void cuts_close_runs_kernel(const uint32_t *__restrict__ counts, uint32_t *__restrict__ runStart, uint32_t n)
{
    runStart[counts[0]] = n;
}

__global__ void cuts_close_csr_kernel(const uint32_t *__restrict__ counts, uint32_t *__restrict__ pointStart,
    uint32_t *__restrict__ edgeStart)
{
    pointStart[counts[0]] = counts[1];
    edgeStart[counts[0]] = counts[2];
}

// run starts + context triangles from the run index of every hit
__global__ void __launch_bounds__(256) cuts_runs_kernel(const uint8_t *__restrict__ head, const uint32_t *__restrict__ runOf,
    const uint32_t *__restrict__ tri, uint32_t n, uint32_t *__restrict__ runStart, uint32_t *__restrict__ cutTri)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n && head[j]) {
        runStart[runOf[j]] = j;
        cutTri[runOf[j]] = tri[j];
    }
}

// PositionKey(position) (reference src/positionkey.cpp:32-37): C truncation of x * 100000 to long
__device__ __forceinline__ long long position_key(double v)
{
    const double s = __dmul_rn(v, 100000.0);
    if (!(s > -9223372036854775808.0 && s < 9223372036854775808.0))
        return (long long)0x8000000000000000ull; // x86 cvttsd2si: "integer indefinite" for NaN / out of range
    return __double2ll_rz(s);
}

// end point e = 2 j + s of side position j: its key, and the first end point of the run with that key
__global__ void __launch_bounds__(256) cuts_keys_kernel(const double *__restrict__ seg, const uint32_t *__restrict__ order,
    uint32_t n, long long *__restrict__ keys /* 3 per end point */)
{
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 2 * n)
        return;
    const uint32_t j = e >> 1, h = order ? __ldg(order + j) : j;
    const double *p = seg + 6 * (size_t)h + 3 * (e & 1u);
    keys[3 * (size_t)e] = position_key(__ldg(p));
    keys[3 * (size_t)e + 1] = position_key(__ldg(p + 1));
    keys[3 * (size_t)e + 2] = position_key(__ldg(p + 2));
}

__global__ void __launch_bounds__(256) cuts_rep_kernel(const long long *__restrict__ keys, const uint32_t *__restrict__ runOf,
    const uint32_t *__restrict__ runStart, uint32_t n, uint32_t *__restrict__ rep, uint8_t *__restrict__ first)
{
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 2 * n)
        return;
    const long long k0 = keys[3 * (size_t)e], k1 = keys[3 * (size_t)e + 1], k2 = keys[3 * (size_t)e + 2];
    uint32_t r = e;
    for (uint32_t q = 2 * runStart[runOf[e >> 1]]; q < e; ++q)
        if (keys[3 * (size_t)q] == k0 && keys[3 * (size_t)q + 1] == k1 && keys[3 * (size_t)q + 2] == k2) {
            r = q; // the earliest one: the position the map keeps
            break;
        }
    rep[e] = r;
    first[e] = r == e ? 1 : 0;
}

// per hit: the two point numbers (3 + index in the context), and whether this is the first
// occurrence of their relation in the run (a segment whose ends coincide has none)
__global__ void __launch_bounds__(256) cuts_edges_kernel(const uint32_t *__restrict__ rep, const uint32_t *__restrict__ pointRank,
    const uint32_t *__restrict__ runOf, const uint32_t *__restrict__ runStart, uint32_t n, uint2 *__restrict__ edgeOf,
    uint8_t *__restrict__ firstEdge)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n)
        return;
    const uint32_t js = runStart[runOf[j]];
    const uint32_t base = pointRank[2 * js]; // first point of the context
    auto number = [&](uint32_t e) { return 3u + pointRank[rep[e]] - base; };
    const uint32_t a = number(2 * j), b = number(2 * j + 1);
    const uint2 mine = make_uint2(min(a, b), max(a, b));
    edgeOf[j] = mine;
    bool fresh = a != b;
    for (uint32_t q = js; fresh && q < j; ++q) {
        const uint32_t qa = number(2 * q), qb = number(2 * q + 1);
        if (min(qa, qb) == mine.x && max(qa, qb) == mine.y)
            fresh = false;
    }
    firstEdge[j] = fresh ? 1 : 0;
}

__global__ void __launch_bounds__(256) cuts_emit_kernel(const double *__restrict__ seg, const uint32_t *__restrict__ order,
    const uint8_t *__restrict__ first, const uint32_t *__restrict__ pointRank, const uint8_t *__restrict__ firstEdge,
    const uint32_t *__restrict__ edgeRank, const uint2 *__restrict__ edgeOf, const uint32_t *__restrict__ runOf,
    const uint32_t *__restrict__ runStart /* closed: [runs] = n */, const uint8_t *__restrict__ head, uint32_t n,
    double *__restrict__ points, uint32_t *__restrict__ pointStart, uint32_t *__restrict__ edges, uint32_t *__restrict__ edgeStart)
{
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 2 * n)
        return;
    const uint32_t j = e >> 1;
    if (first[e]) {
        const double *p = seg + 6 * (size_t)(order ? __ldg(order + j) : j) + 3 * (e & 1u);
        double *o = points + 3 * (size_t)pointRank[e];
        o[0] = __ldg(p);
        o[1] = __ldg(p + 1);
        o[2] = __ldg(p + 2);
    }
    if (e & 1u)
        return;
    // one thread per hit from here on
    const uint32_t run = runOf[j], js = runStart[run];
    if (head[j]) {
        pointStart[run] = pointRank[2 * j];
        edgeStart[run] = edgeRank[j];
    }
    if (!firstEdge[j])
        return;
    // place of this relation among the context's relations in ascending (low, high) order
    const uint32_t jend = runStart[run + 1];
    const uint2 mine = edgeOf[j];
    uint32_t smaller = 0;
    for (uint32_t q = js; q < jend; ++q)
        if (firstEdge[q]) {
            const uint2 o = edgeOf[q];
            smaller += (o.x < mine.x || (o.x == mine.x && o.y < mine.y)) ? 1u : 0u;
        }
    const uint32_t at = edgeRank[js] + smaller;
    edges[2 * (size_t)at] = mine.x;
    edges[2 * (size_t)at + 1] = mine.y;
}

} // namespace

// Scratch of sbk_cut_contexts, in 4-byte words, for n hits.
size_t sbk_cut_contexts_scratch(size_t n)
{
    const size_t tiles = (2 * n + RK_TILE - 1) / RK_TILE + 1;
    return tiles                       // rank tiles
           + 4 * n                     // order ping-pong (keys + values, two buffers each)
           + n                         // tri
           + n                         // runOf
           + n + 1                     // runStart (closed)
           + 2 * n + 2 * n             // rep, pointRank
           + n + 2 * n                 // edgeRank, edgeOf
           + 12 * n                    // keys: 3 x int64 per end point
           + (n + 3) / 4 + (2 * n + 3) / 4 + (n + 3) / 4 // head, first, firstEdge (bytes)
           + 64;                       // alignment slack
}

// hitAB / seg: the (a, b)-sorted hits of an intersection (device).  Outputs sized for the worst
// case: cutTri n, pointStart / edgeStart n + 1, points 6 n doubles, edges 2 n; counts[3]
// (device) receives {contexts, points, relations}.
cudaError_t sbk_cut_contexts(cudaStream_t s, const uint32_t *hitAB, const double *seg, uint32_t n, int which, unsigned bitsTri,
    uint32_t *scratch, uint32_t *radixWs, int smCount, uint32_t *cutTri, uint32_t *pointStart, double *points,
    uint32_t *edgeStart, uint32_t *edges, uint32_t *counts, LaunchCounter &lc)
{
    if (n == 0)
        return cudaMemsetAsync(counts, 0, 3 * sizeof(uint32_t), s);
    const size_t tiles = (2 * (size_t)n + RK_TILE - 1) / RK_TILE + 1;
    uint32_t *w = scratch;
    auto take = [&](size_t words) {
        uint32_t *p = w;
        w += (words + 3) & ~(size_t)3; // 16-byte steps
        return p;
    };
    uint32_t *tileScratch = take(tiles);
    uint32_t *ordKey = take(n), *ordKeyTmp = take(n), *ordVal = take(n), *ordValTmp = take(n);
    uint32_t *tri = take(n), *runOf = take(n), *runStart = take((size_t)n + 1);
    uint32_t *rep = take(2 * (size_t)n), *pointRank = take(2 * (size_t)n), *edgeRank = take(n);
    uint2 *edgeOf = reinterpret_cast<uint2 *>(take(2 * (size_t)n));
    long long *keys = reinterpret_cast<long long *>(take(12 * (size_t)n));
    uint8_t *head = reinterpret_cast<uint8_t *>(take((n + 3) / 4));
    uint8_t *first = reinterpret_cast<uint8_t *>(take((2 * (size_t)n + 3) / 4));
    uint8_t *firstEdge = reinterpret_cast<uint8_t *>(take((n + 3) / 4));

    const uint32_t *order = nullptr;
    const uint32_t hb = (n + 255) / 256, eb = (2 * n + 255) / 256;
    if (which == 1) {
        // group by the second triangle; stable, so inside a group the first triangles still ascend
        cuts_iota_kernel<<<hb, 256, 0, s>>>(hitAB, n, ordKey, ordVal);
        lc.kernels += 1;
        sbradix::Workspace ws;
        ws.mem = radixWs;
        uint32_t *sk = nullptr, *sv = nullptr;
        lc.kernels += sbradix::sort<uint32_t, 8>(s, ordKey, ordKeyTmp, ordVal, ordValTmp, n, 0, (int)bitsTri, ws, smCount, &sk, &sv);
        order = sv;
    }
    cuts_heads_kernel<<<hb, 256, 0, s>>>(hitAB, order, n, which, tri, head);
    lc.kernels += 1;
    cudaError_t e = exclusive_rank(s, head, n, tileScratch, runOf, counts + 0, lc);
    if (e != cudaSuccess)
        return e;
    cuts_fix_kernel<<<hb, 256, 0, s>>>(head, runOf, n);
    cuts_close_runs_kernel<<<1, 1, 0, s>>>(counts, runStart, n);
    cuts_runs_kernel<<<hb, 256, 0, s>>>(head, runOf, tri, n, runStart, cutTri);
    cuts_keys_kernel<<<eb, 256, 0, s>>>(seg, order, n, keys);
    cuts_rep_kernel<<<eb, 256, 0, s>>>(keys, runOf, runStart, n, rep, first);
    lc.kernels += 6;
    e = exclusive_rank(s, first, 2 * n, tileScratch, pointRank, counts + 1, lc);
    if (e != cudaSuccess)
        return e;
    cuts_edges_kernel<<<hb, 256, 0, s>>>(rep, pointRank, runOf, runStart, n, edgeOf, firstEdge);
    lc.kernels += 1;
    e = exclusive_rank(s, firstEdge, n, tileScratch, edgeRank, counts + 2, lc);
    if (e != cudaSuccess)
        return e;
    cuts_emit_kernel<<<eb, 256, 0, s>>>(seg, order, first, pointRank, firstEdge, edgeRank, edgeOf, runOf, runStart, head, n,
        points, pointStart, edges, edgeStart);
    cuts_close_csr_kernel<<<1, 1, 0, s>>>(counts, pointStart, edgeStart);
    lc.kernels += 2;
    return cudaGetLastError();
}
