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
// keys with the reference's insertion ordinal 3 * rank + k as value.  Keys are packed from the
// mesh's own vertex ids as (first << bitsV) | second, bitsV = bits of nV: same order as the
// reference's ((first + offset) << 32) | (second + offset), fewer radix passes.
// Faces are taken in stripes (item s of thread t = face tile + s * 256 + t): neighbouring threads
// read neighbouring triples and, ranks being prefix counts in face order, write neighbouring
// records.  bitsO > 0: the ordinal rides in the low bitsO bits of the key word (one 8-byte
// record per half-edge, a keys-only sort on the bits above it -- the ordinals ascend in emission
// order, so an LSD sort of the upper bits alone leaves equal keys in insertion order); bitsO == 0
// (key and ordinal do not fit 64 bits together): separate ordinal array, (key, value) sort.
__global__ void __launch_bounds__(HE_THREADS) uncut_emit_kernel(const uint8_t *__restrict__ cut,
    const uint32_t *__restrict__ tri, uint32_t nT, uint32_t nV, int *__restrict__ err,
    const uint32_t *__restrict__ tileOffset, uint32_t vertexOffset, unsigned bitsV, unsigned bitsO,
    uint32_t *__restrict__ face, uint32_t *__restrict__ tri3, unsigned long long *__restrict__ keys,
    uint32_t *__restrict__ ords)
{
    constexpr int WARPS = HE_THREADS / 32;
    __shared__ uint32_t s_cnt[HE_ITEMS * WARPS]; // uncut faces per (stripe, warp), then their exclusive scan
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t tile = blockIdx.x * HE_TILE;
    uint32_t before[HE_ITEMS]; // uncut faces of this warp's stripe segment below this lane
    uint32_t keep = 0;
#pragma unroll
    for (int s = 0; s < HE_ITEMS; ++s) {
        const uint32_t f = tile + s * HE_THREADS + threadIdx.x;
        const bool uncut = f < nT && !(cut && __ldg(cut + f));
        const unsigned b = __ballot_sync(SB_FULL, uncut);
        before[s] = __popc(b & lanemask_lt());
        keep |= (uncut ? 1u : 0u) << s;
        if (lane == 0)
            s_cnt[s * WARPS + warp] = __popc(b);
    }
    __syncthreads();
    if (threadIdx.x < 32) { // exclusive scan of the 64 counts (two per lane), in face order
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
    const uint32_t tileRank = tileOffset[blockIdx.x];
#pragma unroll
    for (int s = 0; s < HE_ITEMS; ++s) {
        if (!(keep >> s & 1u))
            continue;
        const uint32_t f = tile + s * HE_THREADS + threadIdx.x;
        const uint32_t rank = tileRank + s_cnt[s * WARPS + warp] + before[s];
        uint32_t v0 = __ldg(tri + 3 * (size_t)f), v1 = __ldg(tri + 3 * (size_t)f + 1), v2 = __ldg(tri + 3 * (size_t)f + 2);
        if (v0 >= nV || v1 >= nV || v2 >= nV) { // the packed keys and the vertex index only hold ids below nV:
            *err = 1;                           // reported as SB_ERR_INVALID by the host
            v0 = v1 = v2 = 0;
        }
        face[rank] = f;
        const size_t o = 3 * (size_t)rank;
        tri3[o] = v0 + vertexOffset;
        tri3[o + 1] = v1 + vertexOffset;
        tri3[o + 2] = v2 + vertexOffset;
        const unsigned long long k0 = ((unsigned long long)v0 << bitsV) | v1, k1 = ((unsigned long long)v1 << bitsV) | v2,
                                 k2 = ((unsigned long long)v2 << bitsV) | v0;
        if (bitsO) {
            keys[o] = (k0 << bitsO) | o;
            keys[o + 1] = (k1 << bitsO) | (o + 1);
            keys[o + 2] = (k2 << bitsO) | (o + 2);
        } else {
            keys[o] = k0;
            keys[o + 1] = k1;
            keys[o + 2] = k2;
            ords[o] = (uint32_t)o;
            ords[o + 1] = (uint32_t)o + 1;
            ords[o + 2] = (uint32_t)o + 2;
        }
    }
}

// First sorted entry of every vertex that starts a half-edge (vstart was filled with 0xff).
__global__ void __launch_bounds__(256) halfedge_vstart_kernel(const unsigned long long *__restrict__ keys, uint32_t n,
    unsigned bitsV, unsigned bitsO, uint32_t *__restrict__ vstart)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n)
        return;
    const uint32_t first = (uint32_t)(__ldg(keys + p) >> (bitsV + bitsO));
    if (p == 0 || (uint32_t)(__ldg(keys + p - 1) >> (bitsV + bitsO)) != first)
        vstart[first] = p;
}

// One thread per sorted half-edge.  ords: the sorted ordinals (bitsO == 0) or null; ordOut
// (bitsO > 0): receives them, unpacked.
__global__ void __launch_bounds__(256) halfedge_link_kernel(const unsigned long long *__restrict__ keys,
    const uint32_t *__restrict__ ords, uint32_t n, unsigned bitsV, unsigned bitsO, uint32_t vertexOffset,
    uint32_t triangleOffset, const uint32_t *__restrict__ vstart, unsigned long long *__restrict__ refKeys,
    uint32_t *__restrict__ owner, int32_t *__restrict__ adj, uint32_t *__restrict__ ordOut,
    uint32_t *__restrict__ firstRepeat)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n)
        return;
    const unsigned long long ordMask = (1ull << bitsO) - 1ull;
    const unsigned long long word = __ldg(keys + p);
    const unsigned long long key = word >> bitsO;
    const uint32_t ord = bitsO ? (uint32_t)(word & ordMask) : __ldg(ords + p);
    if (bitsO)
        ordOut[p] = ord;
    // the sort is stable and the ordinals ascend in emission order: of two equal keys the
    // later entry is the insertion the reference's map refuses
    if (p > 0 && (__ldg(keys + p - 1) >> bitsO) == key)
        atomicMin(firstRepeat, ord);
    const unsigned long long lowMask = (1ull << bitsV) - 1ull;
    const unsigned long long from = key >> bitsV, to = key & lowMask;
    refKeys[p] = ((from + vertexOffset) << 32) | (to + vertexOffset);
    owner[p] = triangleOffset + ord / 3u;
    // the opposite half-edge (to, from) sits among the entries that start at `to`: a handful
    // (the vertex's valence), walked from their first one; a vertex with a huge fan falls back
    // to a binary search of the rest
    const unsigned long long want = (to << bitsV) | from;
    int32_t other = -1;
    uint32_t q = __ldg(vstart + to);
    if (q != 0xffffffffu) {
        int steps = 0;
        while (q < n && (__ldg(keys + q) >> bitsO) < want && ++steps < 16)
            ++q;
        if (steps == 16) {
            uint32_t lo = q, hi = n;
            while (lo < hi) {
                const uint32_t mid = lo + ((hi - lo) >> 1);
                if ((__ldg(keys + mid) >> bitsO) < want)
                    lo = mid + 1;
                else
                    hi = mid;
            }
            q = lo;
        }
        if (q < n) {
            const unsigned long long w = __ldg(keys + q);
            if ((w >> bitsO) == want)
                other = (int32_t)(triangleOffset + (bitsO ? (uint32_t)(w & ordMask) : __ldg(ords + q)) / 3u);
        }
    }
    adj[ord] = other;
}

// ---- connected components of the uncut triangles -----------------------------------------
// buildFaceGroups' flood (reference src/solidboolean.cpp:229-238) without the queue: lock-free
// union-find over the adjacency array, label = lowest triangle index of the component -- the
// triangle the reference's loop over ascending indices opens the group with -- whatever the
// thread order.
//
// The NODES of the union-find are not the new triangles but positions in a spatially coherent
// order: the mesh's Morton order when it has been built (pos2new / new2pos translate; cut faces
// are isolated nodes), else the face order itself.  Face orders as they come have no locality
// to speak of (a subdivision emits all first children, then all second ones; a grid all lower
// triangles, then all upper ones), and without locality every link is an L2 compare-and-swap
// on a tree whose finds are chains of dependent L2 loads.
//   pass 1  cc_tile   union-find inside tiles of CC_TILE consecutive nodes, in shared memory;
//                     every node leaves pointing at its tile root, which also carries the
//                     lowest NEW TRIANGLE index of its tile component
//   pass 2  cc_hook   edges that leave a tile, on the global array, one lane per distinct
//                     (tile root, tile root) pair of a warp, linking by a hashed priority
//                     instead of the position: random linking keeps the trees O(log) deep
//                     (linking by index grows paths: i under i - 1 under i - 2 ...)
//   pass 3  cc_min    the lowest triangle index of every tree (atomicMin at the tree's root)
//   pass 4  cc_label  label = that minimum; roots counted
#ifndef SB_CC_TILE
#define SB_CC_TILE 256 // measured at C3 / C2: 1024 x 256 threads 0.48 / 0.25 ms, 256 x 256 0.40 / 0.16, 64 x 64 0.48 / 0.27
#endif
#ifndef SB_CC_THREADS
#define SB_CC_THREADS 256
#endif
constexpr int CC_TILE = SB_CC_TILE;       // nodes per shared-memory tile (power of two)
constexpr int CC_THREADS = SB_CC_THREADS; // threads per tile
constexpr uint32_t CC_NONE = 0xffffffffu;

// a bijection of the 32-bit ids (murmur3's finaliser) whose order looks random: tile roots are
// mostly multiples of the tile size, and a multiplicative hash keeps long monotone runs along
// arithmetic progressions of them -- exactly the paths the random linking is there to avoid
__device__ __forceinline__ uint32_t cc_prio(uint32_t x)
{
    x ^= x >> 16;
    x *= 0x85ebca6bu;
    x ^= x >> 13;
    x *= 0xc2b2ae35u;
    x ^= x >> 16;
    return x;
}

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

__device__ __forceinline__ uint32_t cc_find_s(volatile uint32_t *sp, uint32_t x)
{
    uint32_t p = sp[x];
    while (p != x) {
        const uint32_t g = sp[p];
        if (g != p)
            sp[x] = g;
        x = p;
        p = g;
    }
    return x;
}

// node order = face order: position p is new triangle p
__global__ void __launch_bounds__(256) cc_identity_kernel(uint32_t n, uint32_t *__restrict__ pos2new, uint32_t *__restrict__ new2pos)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        pos2new[i] = i;
        new2pos[i] = i;
    }
}

// node order = the mesh's Morton order.  faceRank (filled with 0xff): original face -> new triangle.
__global__ void __launch_bounds__(256) cc_rank_kernel(const uint32_t *__restrict__ face, uint32_t nTri, uint32_t *__restrict__ faceRank)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < nTri)
        faceRank[__ldg(face + r)] = r;
}

__global__ void __launch_bounds__(256) cc_order_kernel(const uint32_t *__restrict__ sortedTri, const uint32_t *__restrict__ faceRank,
    uint32_t nT, uint32_t *__restrict__ pos2new, uint32_t *__restrict__ new2pos)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nT)
        return;
    const uint32_t r = __ldg(faceRank + __ldg(sortedTri + p));
    pos2new[p] = r;
    if (r != CC_NONE)
        new2pos[r] = p;
}

__global__ void __launch_bounds__(CC_THREADS) cc_tile_kernel(const int32_t *__restrict__ adj, uint32_t nodes,
    uint32_t triangleOffset, const uint32_t *__restrict__ pos2new, const uint32_t *__restrict__ new2pos,
    uint32_t *__restrict__ parent, uint32_t *__restrict__ tileRoot, uint32_t *__restrict__ minIdx)
{
    __shared__ uint32_t sp[CC_TILE];
    __shared__ uint32_t smin[CC_TILE];
    const uint32_t tileStart = blockIdx.x * CC_TILE;
    for (uint32_t l = threadIdx.x; l < CC_TILE; l += CC_THREADS) {
        sp[l] = l;
        smin[l] = tileStart + l < nodes ? __ldg(pos2new + tileStart + l) : CC_NONE;
    }
    __syncthreads();
    for (uint32_t l = threadIdx.x; l < CC_TILE; l += CC_THREADS) {
        const uint32_t r = smin[l]; // still this node's own triangle: the minima are formed after the next barrier
        if (r == CC_NONE)
            continue;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int32_t o = __ldg(adj + 3 * (size_t)r + k);
            if (o < 0)
                continue;
            const uint32_t q = __ldg(new2pos + ((uint32_t)o - triangleOffset));
            if (q >= tileStart + l || q < tileStart)
                continue; // each edge once, from its later end; other tiles: pass 2
            uint32_t a = cc_find_s(sp, l), b = cc_find_s(sp, q - tileStart);
            while (a != b) {
                if (cc_prio(a) < cc_prio(b)) {
                    const uint32_t t = a;
                    a = b;
                    b = t;
                }
                const uint32_t old = atomicCAS(sp + a, a, b); // random linking here too: shallow trees
                if (old == a)
                    break;
                a = cc_find_s(sp, old);
                b = cc_find_s(sp, b);
            }
        }
    }
    __syncthreads();
    uint32_t root[CC_TILE / CC_THREADS], mine[CC_TILE / CC_THREADS];
#pragma unroll
    for (int s = 0; s < CC_TILE / CC_THREADS; ++s) {
        const uint32_t l = threadIdx.x + s * CC_THREADS;
        root[s] = cc_find_s(sp, l);
        mine[s] = smin[l];
    }
    __syncthreads(); // everybody holds its own triangle: smin may now turn into the per-root minima
#pragma unroll
    for (int s = 0; s < CC_TILE / CC_THREADS; ++s)
        if (mine[s] != CC_NONE && root[s] != threadIdx.x + s * CC_THREADS)
            atomicMin(smin + root[s], mine[s]);
    __syncthreads();
#pragma unroll
    for (int s = 0; s < CC_TILE / CC_THREADS; ++s) {
        const uint32_t l = threadIdx.x + s * CC_THREADS;
        if (tileStart + l < nodes) {
            parent[tileStart + l] = tileStart + root[s];
            tileRoot[tileStart + l] = tileStart + root[s];
            minIdx[tileStart + l] = root[s] == l ? smin[l] : CC_NONE;
        }
    }
}

__global__ void __launch_bounds__(256) cc_hook_kernel(const int32_t *__restrict__ adj, uint32_t nodes, uint32_t triangleOffset,
    const uint32_t *__restrict__ pos2new, const uint32_t *__restrict__ new2pos, const uint32_t *__restrict__ tileRoot,
    uint32_t *parent)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t r = p < nodes ? __ldg(pos2new + p) : CC_NONE;
    const uint32_t tileStart = p & ~(uint32_t)(CC_TILE - 1);
    const uint32_t lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        unsigned long long pair = ~0ull - lane; // no edge: a key nobody shares (bit 63 set; positions are below 2^31)
        if (r != CC_NONE) {
            const int32_t o = __ldg(adj + 3 * (size_t)r + k);
            if (o >= 0) {
                const uint32_t q = __ldg(new2pos + ((uint32_t)o - triangleOffset));
                if (q < tileStart)
                    pair = ((unsigned long long)__ldg(tileRoot + p) << 32) | __ldg(tileRoot + q);
            }
        }
        const unsigned same = __match_any_sync(SB_FULL, pair);
        if ((pair >> 63) || lane != (uint32_t)(__ffs(same) - 1))
            continue;
        uint32_t a = cc_find(parent, (uint32_t)(pair >> 32)), b = cc_find(parent, (uint32_t)pair);
        while (a != b) {
            if (cc_prio(a) < cc_prio(b)) {
                const uint32_t t = a;
                a = b;
                b = t;
            }
            const uint32_t old = atomicCAS(parent + a, a, b); // the root of larger priority value under the other
            if (old == a)
                break;
            a = cc_find(parent, old);
            b = cc_find(parent, b);
        }
    }
}

// every tile root that is not the root of its tree hands its minimum up (its own slot is
// written by nobody else); most only read and leave once a low index has arrived
__global__ void __launch_bounds__(256) cc_min_kernel(uint32_t *parent, const uint32_t *__restrict__ tileRoot, uint32_t nodes,
    uint32_t *minIdx)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nodes || __ldg(tileRoot + p) != p)
        return;
    const uint32_t root = cc_find(parent, p);
    if (root == p)
        return;
    const uint32_t v = __ldcg(minIdx + p);
    if (v < __ldcg(minIdx + root))
        atomicMin(minIdx + root, v);
}

__global__ void __launch_bounds__(256) cc_label_kernel(uint32_t *parent, uint32_t nodes, uint32_t triangleOffset,
    const uint32_t *__restrict__ pos2new, const uint32_t *minIdx, uint32_t *__restrict__ label, uint32_t *__restrict__ count)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t r = p < nodes ? __ldg(pos2new + p) : CC_NONE;
    uint32_t root = CC_NONE;
    if (r != CC_NONE) {
        root = cc_find(parent, p);
        label[r] = __ldcg(minIdx + root) + triangleOffset;
    }
    const unsigned roots = __ballot_sync(SB_FULL, r != CC_NONE && root == p);
    if ((threadIdx.x & 31) == 0 && roots)
        atomicAdd(count, __popc(roots));
}

} // namespace

// Words of scratch sbk_uncut_components needs.
size_t sbk_uncut_components_scratch(uint32_t nTri, uint32_t nT, bool ordered)
{
    const size_t nodes = ordered ? nT : nTri;
    return 4 * nodes + nTri + (ordered ? nT : 0);
}

// sortedTri: the mesh's Morton order (sorted position -> face, nT entries) or null; face: new
// triangle -> face; label: nTri; *count (device, zeroed by the caller) += components
cudaError_t sbk_uncut_components(cudaStream_t s, const int32_t *adj, uint32_t nTri, uint32_t triangleOffset,
    const uint32_t *sortedTri, const uint32_t *face, uint32_t nT, uint32_t *scratch, uint32_t *label, uint32_t *count,
    LaunchCounter &lc)
{
    if (nTri == 0)
        return cudaSuccess;
    const uint32_t nodes = sortedTri ? nT : nTri;
    uint32_t *parent = scratch, *minIdx = parent + nodes, *tileRoot = minIdx + nodes, *pos2new = tileRoot + nodes,
             *new2pos = pos2new + nodes, *faceRank = new2pos + nTri;
    if (sortedTri) {
        cudaMemsetAsync(faceRank, 0xff, sizeof(uint32_t) * (size_t)nT, s);
        cc_rank_kernel<<<(nTri + 255) / 256, 256, 0, s>>>(face, nTri, faceRank);
        cc_order_kernel<<<(nT + 255) / 256, 256, 0, s>>>(sortedTri, faceRank, nT, pos2new, new2pos);
        lc.kernels += 2;
    } else {
        cc_identity_kernel<<<(nTri + 255) / 256, 256, 0, s>>>(nTri, pos2new, new2pos);
        lc.kernels += 1;
    }
    const uint32_t blocks = (nodes + 255) / 256;
    cc_tile_kernel<<<(nodes + CC_TILE - 1) / CC_TILE, CC_THREADS, 0, s>>>(adj, nodes, triangleOffset, pos2new, new2pos, parent,
        tileRoot, minIdx);
    cc_hook_kernel<<<blocks, 256, 0, s>>>(adj, nodes, triangleOffset, pos2new, new2pos, tileRoot, parent);
    cc_min_kernel<<<blocks, 256, 0, s>>>(parent, tileRoot, nodes, minIdx);
    cc_label_kernel<<<blocks, 256, 0, s>>>(parent, nodes, triangleOffset, pos2new, minIdx, label, count);
    lc.kernels += 4;
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

cudaError_t sbk_uncut_emit(cudaStream_t s, const uint8_t *cut, const uint32_t *tri, uint32_t nT, uint32_t nV, int *err,
    const uint32_t *tileScratch, uint32_t vertexOffset, unsigned bitsV, unsigned bitsO, uint32_t *face, uint32_t *tri3,
    unsigned long long *keys, uint32_t *ords, LaunchCounter &lc)
{
    const uint32_t tiles = sbk_uncut_tiles(nT);
    if (tiles == 0)
        return cudaSuccess;
    uncut_emit_kernel<<<tiles, HE_THREADS, 0, s>>>(cut, tri, nT, nV, err, tileScratch, vertexOffset, bitsV, bitsO, face, tri3,
        keys, ords);
    lc.kernels += 1;
    return cudaGetLastError();
}

// vstart: nV words of scratch
cudaError_t sbk_halfedge_link(cudaStream_t s, const unsigned long long *sortedKeys, const uint32_t *sortedOrds, uint32_t n,
    unsigned bitsV, unsigned bitsO, uint32_t nV, uint32_t *vstart, uint32_t vertexOffset, uint32_t triangleOffset,
    unsigned long long *refKeys, uint32_t *owner, int32_t *adj, uint32_t *ordOut, uint32_t *firstRepeat, LaunchCounter &lc)
{
    if (n == 0)
        return cudaSuccess;
    cudaMemsetAsync(vstart, 0xff, sizeof(uint32_t) * (size_t)nV, s);
    halfedge_vstart_kernel<<<(n + 255) / 256, 256, 0, s>>>(sortedKeys, n, bitsV, bitsO, vstart);
    halfedge_link_kernel<<<(n + 255) / 256, 256, 0, s>>>(sortedKeys, sortedOrds, n, bitsV, bitsO, vertexOffset, triangleOffset,
        vstart, refKeys, owner, adj, ordOut, firstRepeat);
    lc.kernels += 2;
    return cudaGetLastError();
}
