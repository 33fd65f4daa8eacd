// Uncut triangles + half-edge map on the device.
//
// Replaces SolidBoolean::addUnintersectedTriangles (reference src/solidboolean.cpp:250-286):
// every triangle the intersection did not touch is appended to m_newTriangles with its
// vertex ids shifted behind the vertices already taken over (:254, :265-269), and its
// half-edges (0,1), (1,2), (2,0) go into a map keyed by makeHalfEdgeKey(first, second) =
// (first << 32) | second (src/solidboolean.h:75-78) whose value is the new triangle's index
// (:271-282).  In the reference this is one hash insert per half-edge (4-5.7 s at 1M+1M
// triangles); here it is
//   uncut_count -> uncut_scan -> uncut_emit   stream compaction of the uncut faces in their
//                                             original order (new index = rank), keys emitted
//                                             in the reference's insertion order
//   onesweep radix sort                       (key, insertion ordinal), stable
//   halfedge_link                             repeated keys (the reference's "Found repeated
//                                             halfedge"), reference-format keys + owners, and
//                                             for every half-edge the triangle holding the
//                                             opposite one = what buildFaceGroups asks the
//                                             map (:205-224), by binary search
// The sorted (key, owner) arrays ARE the map (lookup = binary search); the adjacency array
// answers buildFaceGroups' lookups without any map.
#include "sb_internal.h"

namespace {

constexpr int HE_THREADS = 256;
constexpr int HE_ITEMS = 8;
constexpr int HE_TILE = HE_THREADS * HE_ITEMS;

// bit i of the result = face base + i is uncut (faces beyond nT count as cut)
__device__ __forceinline__ uint32_t uncut_mask8(const uint8_t *__restrict__ cut, uint32_t base, uint32_t nT)
{
    uint32_t m = 0;
    if (base + HE_ITEMS <= nT) {
        if (!cut)
            return 0xffu;
        const uint2 w = __ldg(reinterpret_cast<const uint2 *>(cut + base)); // base is a multiple of 8
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            m |= ((w.x >> (8 * i)) & 0xffu) ? 0u : (1u << i);
            m |= ((w.y >> (8 * i)) & 0xffu) ? 0u : (1u << (4 + i));
        }
        return m;
    }
    for (int i = 0; i < HE_ITEMS; ++i)
        if (base + i < nT && !(cut && cut[base + i]))
            m |= 1u << i;
    return m;
}

__device__ __forceinline__ uint32_t block_sum(uint32_t v, uint32_t *s_warp /* HE_THREADS / 32 */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
        v += __shfl_xor_sync(SB_FULL, v, off);
    if (lane == 0)
        s_warp[warp] = v;
    __syncthreads();
    uint32_t t = 0;
#pragma unroll
    for (int w = 0; w < HE_THREADS / 32; ++w)
        t += s_warp[w];
    return t;
}

__global__ void __launch_bounds__(HE_THREADS) uncut_count_kernel(const uint8_t *__restrict__ cut, uint32_t nT,
    uint32_t *__restrict__ tileCount)
{
    __shared__ uint32_t s_warp[HE_THREADS / 32];
    const uint32_t base = blockIdx.x * HE_TILE + threadIdx.x * HE_ITEMS;
    const uint32_t t = block_sum(__popc(uncut_mask8(cut, base, nT)), s_warp);
    if (threadIdx.x == 0)
        tileCount[blockIdx.x] = t;
}

// exclusive scan of the tile counts in place (one CTA; at most 2^25 / 2048 = 16384 tiles)
__global__ void __launch_bounds__(1024) uncut_scan_kernel(uint32_t *__restrict__ tileCount, uint32_t tiles,
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
            uint32_t t = __shfl_up_sync(SB_FULL, incl, off);
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

// New triangle `rank` = original face i: the face id, the shifted index triple, and its three
// keys with the reference's insertion ordinal 3 * rank + k as value.  Keys are packed as
// (first << bitsV) | second (same order as the reference's (first << 32) | second, fewer
// radix passes).
__global__ void __launch_bounds__(HE_THREADS) uncut_emit_kernel(const uint8_t *__restrict__ cut,
    const uint32_t *__restrict__ tri, uint32_t nT, const uint32_t *__restrict__ tileOffset, uint32_t vertexOffset,
    unsigned bitsV, uint32_t *__restrict__ face, uint32_t *__restrict__ tri3, unsigned long long *__restrict__ keys,
    uint32_t *__restrict__ ords)
{
    __shared__ uint32_t s_warp[HE_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t base = blockIdx.x * HE_TILE + threadIdx.x * HE_ITEMS;
    const uint32_t mask = uncut_mask8(cut, base, nT);
    const uint32_t mine = __popc(mask);
    uint32_t incl = mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        uint32_t t = __shfl_up_sync(SB_FULL, incl, off);
        if (lane >= off)
            incl += t;
    }
    if (lane == 31)
        s_warp[warp] = incl;
    __syncthreads();
    uint32_t rank = tileOffset[blockIdx.x] + incl - mine;
    for (int w = 0; w < warp; ++w)
        rank += s_warp[w];
#pragma unroll
    for (int i = 0; i < HE_ITEMS; ++i) {
        if (!(mask >> i & 1u))
            continue;
        const uint32_t f = base + i;
        const uint32_t v0 = __ldg(tri + 3 * (size_t)f) + vertexOffset, v1 = __ldg(tri + 3 * (size_t)f + 1) + vertexOffset,
                       v2 = __ldg(tri + 3 * (size_t)f + 2) + vertexOffset;
        face[rank] = f;
        const size_t o = 3 * (size_t)rank;
        tri3[o] = v0;
        tri3[o + 1] = v1;
        tri3[o + 2] = v2;
        keys[o] = ((unsigned long long)v0 << bitsV) | v1;
        keys[o + 1] = ((unsigned long long)v1 << bitsV) | v2;
        keys[o + 2] = ((unsigned long long)v2 << bitsV) | v0;
        ords[o] = (uint32_t)o;
        ords[o + 1] = (uint32_t)o + 1;
        ords[o + 2] = (uint32_t)o + 2;
        ++rank;
    }
}

// One thread per sorted half-edge.
__global__ void __launch_bounds__(256) halfedge_link_kernel(const unsigned long long *__restrict__ keys,
    const uint32_t *__restrict__ ords, uint32_t n, unsigned bitsV, uint32_t triangleOffset,
    unsigned long long *__restrict__ refKeys, uint32_t *__restrict__ owner, int32_t *__restrict__ adj,
    uint32_t *__restrict__ firstRepeat)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n)
        return;
    const unsigned long long key = __ldg(keys + p);
    const uint32_t ord = __ldg(ords + p);
    // the sort is stable and the ordinals ascend in emission order: of two equal keys the
    // later entry is the insertion the reference's map refuses
    if (p > 0 && __ldg(keys + p - 1) == key)
        atomicMin(firstRepeat, ord);
    const unsigned long long lowMask = (1ull << bitsV) - 1ull;
    const unsigned long long from = key >> bitsV, to = key & lowMask;
    refKeys[p] = (from << 32) | to;
    owner[p] = triangleOffset + ord / 3u;
    // the opposite half-edge (to, from): lower bound in the sorted keys
    const unsigned long long want = (to << bitsV) | from;
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(keys + mid) < want)
            lo = mid + 1;
        else
            hi = mid;
    }
    int32_t other = -1;
    if (lo < n && __ldg(keys + lo) == want)
        other = (int32_t)(triangleOffset + __ldg(ords + lo) / 3u);
    adj[ord] = other;
}

// ---- connected components of the uncut triangles -----------------------------------------
// buildFaceGroups' flood (reference src/solidboolean.cpp:229-238) without the queue: lock-free
// union-find over the adjacency array.  A root only ever gets a SMALLER root as parent, so
// every component ends up rooted at its lowest triangle index -- the triangle the reference's
// loop over ascending indices would have opened the group with -- whatever the thread order.
__device__ __forceinline__ uint32_t cc_find(uint32_t *parent, uint32_t x)
{
    uint32_t p = __ldcg(parent + x);
    while (p != x) {
        const uint32_t g = __ldcg(parent + p);
        if (g != p)
            parent[x] = g; // path halving: any ancestor is a valid parent
        x = p;
        p = g;
    }
    return x;
}

__global__ void __launch_bounds__(256) cc_init_kernel(uint32_t *__restrict__ parent, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        parent[i] = i;
}

__global__ void __launch_bounds__(256) cc_hook_kernel(const int32_t *__restrict__ adj, uint32_t n, uint32_t triangleOffset,
    uint32_t *parent)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int32_t o = __ldg(adj + 3 * (size_t)i + k);
        if (o < 0)
            continue;
        const uint32_t j = (uint32_t)o - triangleOffset;
        if (j >= i)
            continue; // every edge once (the relation is symmetric), from its larger end
        uint32_t a = cc_find(parent, i), b = cc_find(parent, j);
        while (a != b) {
            if (a < b) {
                const uint32_t t = a;
                a = b;
                b = t;
            }
            const uint32_t old = atomicCAS(parent + a, a, b); // hang the larger root under the smaller
            if (old == a)
                break;
            a = cc_find(parent, old);
            b = cc_find(parent, b);
        }
    }
}

__global__ void __launch_bounds__(256) cc_label_kernel(uint32_t *parent, uint32_t n, uint32_t triangleOffset,
    uint32_t *__restrict__ label, uint32_t *__restrict__ count)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n;
    uint32_t r = 0;
    if (live) {
        r = cc_find(parent, i);
        label[i] = r + triangleOffset;
    }
    const unsigned roots = __ballot_sync(SB_FULL, live && r == i);
    if ((threadIdx.x & 31) == 0 && roots)
        atomicAdd(count, __popc(roots));
}

} // namespace

// parent: n words of scratch; label: n; *count (device, zeroed by the caller) += components
cudaError_t sbk_uncut_components(cudaStream_t s, const int32_t *adj, uint32_t n, uint32_t triangleOffset, uint32_t *parent,
    uint32_t *label, uint32_t *count, LaunchCounter &lc)
{
    if (n == 0)
        return cudaSuccess;
    const uint32_t blocks = (n + 255) / 256;
    cc_init_kernel<<<blocks, 256, 0, s>>>(parent, n);
    cc_hook_kernel<<<blocks, 256, 0, s>>>(adj, n, triangleOffset, parent);
    cc_label_kernel<<<blocks, 256, 0, s>>>(parent, n, triangleOffset, label, count);
    lc.kernels += 3;
    return cudaGetLastError();
}

uint32_t sbk_uncut_tiles(uint32_t nT) { return (nT + HE_TILE - 1) / HE_TILE; }

// tileScratch: sbk_uncut_tiles(nT) words; *total (device) receives the number of uncut faces
cudaError_t sbk_uncut_count(cudaStream_t s, const uint8_t *cut, uint32_t nT, uint32_t *tileScratch, uint32_t *total,
    LaunchCounter &lc)
{
    const uint32_t tiles = sbk_uncut_tiles(nT);
    if (tiles == 0) {
        return cudaMemsetAsync(total, 0, sizeof(uint32_t), s);
    }
    uncut_count_kernel<<<tiles, HE_THREADS, 0, s>>>(cut, nT, tileScratch);
    uncut_scan_kernel<<<1, 1024, 0, s>>>(tileScratch, tiles, total);
    lc.kernels += 2;
    return cudaGetLastError();
}

cudaError_t sbk_uncut_emit(cudaStream_t s, const uint8_t *cut, const uint32_t *tri, uint32_t nT, const uint32_t *tileScratch,
    uint32_t vertexOffset, unsigned bitsV, uint32_t *face, uint32_t *tri3, unsigned long long *keys, uint32_t *ords,
    LaunchCounter &lc)
{
    const uint32_t tiles = sbk_uncut_tiles(nT);
    if (tiles == 0)
        return cudaSuccess;
    uncut_emit_kernel<<<tiles, HE_THREADS, 0, s>>>(cut, tri, nT, tileScratch, vertexOffset, bitsV, face, tri3, keys, ords);
    lc.kernels += 1;
    return cudaGetLastError();
}

cudaError_t sbk_halfedge_link(cudaStream_t s, const unsigned long long *sortedKeys, const uint32_t *sortedOrds, uint32_t n,
    unsigned bitsV, uint32_t triangleOffset, unsigned long long *refKeys, uint32_t *owner, int32_t *adj,
    uint32_t *firstRepeat, LaunchCounter &lc)
{
    if (n == 0)
        return cudaSuccess;
    halfedge_link_kernel<<<(n + 255) / 256, 256, 0, s>>>(sortedKeys, sortedOrds, n, bitsV, triangleOffset, refKeys, owner, adj,
        firstRepeat);
    lc.kernels += 1;
    return cudaGetLastError();
}
