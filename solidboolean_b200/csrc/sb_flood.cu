// the flood of SolidBoolean::buildFaceGroups (reference src/solidboolean.cpp:167-239) over ALL
// triangles of one mesh's side of the result -- the uncut triangles (already grouped on the device:
// sb_uncut_components, sb_halfedge.cu) AND the retriangulated pieces the host produced -- as connected
// components on the device:
//   nodes   the components of the uncut triangles (one node each: the rank of their lowest member) and the pieces
//   edges   a piece's half-edge (a, b) joins it to whoever owns the opposite half-edge (b, a): another piece
//           (sorted piece keys, binary search) or an uncut triangle (the uncut half-edge map: sorted keys + owner,
//           the lookup of :216-221), unless (a, b) is an edge of an intersection loop in either direction
//           (the fences of :176-203: a fill never crosses the curve)
//   labels  lock-free union-find, the higher root always hooked under the lower one: the root of a component is
//           its lowest node -- deterministic whatever the thread order.
// Where two loop seeds of the reference reach the same region its queue order splits the region between two
// groups lying on the same side of every loop; a component here is the union of such groups, which keeps the
// triangles each operation selects the same (every group of a component gets the same inside / outside answer).
#include "sb_internal.h"
#include "sb_radix.cuh"
#include <algorithm>

namespace {

__device__ __forceinline__ uint32_t uf_find(uint32_t *parent, uint32_t x)
{
    // path halving; every value read is an ancestor of x (parents only ever move towards lower ids)
    while (true) {
        uint32_t p = __ldcg(parent + x);
        if (p == x)
            return x;
        uint32_t gp = __ldcg(parent + p);
        if (gp != p)
            atomicMin(parent + x, gp);
        x = p;
    }
}

__device__ __forceinline__ void uf_union(uint32_t *parent, uint32_t a, uint32_t b)
{
    while (true) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b)
            return;
        if (a < b) {
            uint32_t t = a; a = b; b = t;
        }
        // a > b: hook root a under b (only succeeds while a still is a root)
        if (atomicCAS(parent + a, a, b) == a)
            return;
    }
}

// first index with keys[i] >= k
__device__ __forceinline__ uint32_t lower_bound64(const unsigned long long *__restrict__ keys, uint32_t n, unsigned long long k)
{
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(keys + mid) < k)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) flood_init_kernel(uint32_t *__restrict__ parent, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        parent[i] = i;
}

// half-edge keys of the pieces (value = 3 * piece + edge) and of the fences (both directions)
__global__ void __launch_bounds__(256) flood_keys_kernel(const uint32_t *__restrict__ pieces, uint32_t nP, const uint32_t *__restrict__ fences,
    uint32_t nF, unsigned long long *__restrict__ pKeys, uint32_t *__restrict__ pVals, unsigned long long *__restrict__ fKeys)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 3 * nP) {
        const uint32_t t = i / 3, k = i % 3;
        const uint32_t a = pieces[3 * (size_t)t + k], b = pieces[3 * (size_t)t + (k + 1) % 3];
        pKeys[i] = ((unsigned long long)a << 32) | b;
        pVals[i] = i;
    }
    if (i < nF) {
        const uint32_t a = fences[2 * (size_t)i], b = fences[2 * (size_t)i + 1];
        fKeys[2 * (size_t)i] = ((unsigned long long)a << 32) | b;
        fKeys[2 * (size_t)i + 1] = ((unsigned long long)b << 32) | a;
    }
}

__global__ void __launch_bounds__(256) flood_link_kernel(const uint32_t *__restrict__ pieces, uint32_t nP,
    const unsigned long long *__restrict__ pKeys, const uint32_t *__restrict__ pVals, const unsigned long long *__restrict__ fKeys,
    uint32_t nF2, const unsigned long long *__restrict__ uKeys, const uint32_t *__restrict__ uOwner, uint32_t nUK,
    const uint32_t *__restrict__ uLabel, uint32_t nU, uint32_t triangleOffset, uint32_t *parent)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * nP)
        return;
    const uint32_t t = i / 3, k = i % 3;
    const uint32_t a = pieces[3 * (size_t)t + k], b = pieces[3 * (size_t)t + (k + 1) % 3];
    const unsigned long long fwd = ((unsigned long long)a << 32) | b, opp = ((unsigned long long)b << 32) | a;
    if (nF2) { // an edge of an intersection loop: the fill stops here (either direction)
        const uint32_t f = lower_bound64(fKeys, nF2, fwd);
        if (f < nF2 && __ldg(fKeys + f) == fwd)
            return;
    }
    const uint32_t me = nU + t;
    const uint32_t q = lower_bound64(pKeys, 3 * nP, opp);
    if (q < 3 * nP && __ldg(pKeys + q) == opp) {
        const uint32_t other = __ldg(pVals + q) / 3;
        if (other != t)
            uf_union(parent, me, nU + other);
        return;
    }
    if (nUK) {
        const uint32_t u = lower_bound64(uKeys, nUK, opp);
        if (u < nUK && __ldg(uKeys + u) == opp) {
            const uint32_t tri = __ldg(uOwner + u) - triangleOffset; // rank of the uncut triangle
            if (tri < nU)
                uf_union(parent, me, __ldg(uLabel + tri) - triangleOffset);
        }
    }
}

// labels: lowest node of the component (uncut triangle i: through its component's node); counts the components
__global__ void __launch_bounds__(256) flood_label_kernel(uint32_t *parent, const uint32_t *__restrict__ uLabel, uint32_t nU, uint32_t nP,
    uint32_t triangleOffset, uint32_t *__restrict__ labelUncut, uint32_t *__restrict__ labelPiece, unsigned int *__restrict__ nGroups)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool isRoot = false;
    if (i < nU) {
        const uint32_t node = __ldg(uLabel + i) - triangleOffset;
        const uint32_t r = uf_find(parent, node);
        labelUncut[i] = r;
        isRoot = r == i; // the lowest member of its uncut component is that component's node
    } else if (i < nU + nP) {
        const uint32_t r = uf_find(parent, i);
        labelPiece[i - nU] = r;
        isRoot = r == i;
    }
    const uint32_t m = __ballot_sync(SB_FULL, isRoot);
    if (m && (threadIdx.x & 31) == 0)
        atomicAdd(nGroups, (unsigned int)__popc(m));
}

} // namespace

size_t sbk_flood_scratch_words(uint32_t nU, uint32_t nP, uint32_t nF)
{
    // parent | pKeys x2 | pVals x2 | fKeys x2 | radix workspace
    const size_t k = 3 * (size_t)nP, f = 2 * (size_t)nF;
    return (size_t)nU + nP + 4 * (k + 1) + 2 * (k + 1) + 4 * (f + 1) + sbradix::Workspace::words(sbradix::tiles_for(std::max<size_t>(k, f) + 1)) + 64;
}

// pieces / fences: device arrays (3 nP / 2 nF vertex ids).  uKeys / uOwner / uLabel: the uncut side (sb_halfedge.cu).
// keyBits: bits of the largest vertex id (both halves of a key are sorted on that many bits).
cudaError_t sbk_flood(cudaStream_t s, const uint32_t *pieces, uint32_t nP, const uint32_t *fences, uint32_t nF,
    const unsigned long long *uKeys, const uint32_t *uOwner, uint32_t nUK, const uint32_t *uLabel, uint32_t nU, uint32_t triangleOffset,
    unsigned keyBits, uint32_t *scratch, int smCount, uint32_t *labelUncut, uint32_t *labelPiece, unsigned int *nGroups, LaunchCounter &lc)
{
    const size_t k = 3 * (size_t)nP, f = 2 * (size_t)nF;
    uint32_t *parent = scratch;
    unsigned long long *pKeys = reinterpret_cast<unsigned long long *>(scratch + (((size_t)nU + nP + 1) & ~(size_t)1));
    unsigned long long *pKeysT = pKeys + (k + 1);
    unsigned long long *fKeys = pKeysT + (k + 1);
    unsigned long long *fKeysT = fKeys + (f + 1);
    uint32_t *pVals = reinterpret_cast<uint32_t *>(fKeysT + (f + 1));
    uint32_t *pValsT = pVals + (k + 1);
    uint32_t *radixWs = pValsT + (k + 2);
    const uint32_t n = nU + nP;
    if (n)
        flood_init_kernel<<<(n + 255) / 256, 256, 0, s>>>(parent, n);
    cudaMemsetAsync(nGroups, 0, sizeof(unsigned int), s);
    lc.kernels += 1;
    unsigned long long *pk = pKeys, *fk = fKeys;
    uint32_t *pv = pVals;
    if (nP) {
        const uint32_t m = (uint32_t)std::max(k, (size_t)nF);
        flood_keys_kernel<<<(m + 255) / 256, 256, 0, s>>>(pieces, nP, fences, nF, pKeys, pVals, fKeys);
        lc.kernels += 1;
        sbradix::Workspace ws;
        ws.mem = radixWs;
        // the low half of a key on bits [0, keyBits), the high half on [32, 32 + keyBits): two sorts of keyBits each
        // would need a stable composite -- simpler: one sort over [0, 32 + keyBits), the zero bits in between cost passes
        // only when keyBits is small; the inputs here are tens of thousands of keys
        lc.kernels += sbradix::sort<unsigned long long, 8>(s, pKeys, pKeysT, pVals, pValsT, k, 0, 32 + (int)keyBits, ws, smCount, &pk, &pv);
        if (nF)
            lc.kernels += sbradix::sort<unsigned long long, 8>(s, fKeys, fKeysT, nullptr, nullptr, f, 0, 32 + (int)keyBits, ws, smCount, &fk, nullptr);
        flood_link_kernel<<<(uint32_t)((k + 255) / 256), 256, 0, s>>>(pieces, nP, pk, pv, fk, (uint32_t)f, uKeys, uOwner, nUK, uLabel, nU,
            triangleOffset, parent);
        lc.kernels += 1;
    }
    if (n) {
        flood_label_kernel<<<(n + 255) / 256, 256, 0, s>>>(parent, uLabel, nU, nP, triangleOffset, labelUncut, labelPiece, nGroups);
        lc.kernels += 1;
    }
    return cudaGetLastError();
}
