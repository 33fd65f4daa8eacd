// Onesweep LSD radix sort (8-bit digits, chained scan with decoupled lookback):
// one histogram pass over the keys for ALL digit positions, then one
// read+write pass per digit.  Used for the Morton sort of triangles (u32 key +
// u32 payload) and for ordering candidate / hit pair keys (u64 key, optional
// u32 payload).  Stable; keys-only or key-value.
#pragma once
#include "sb_common.cuh"

namespace sbradix {

#ifndef SB_RADIX_THREADS
#define SB_RADIX_THREADS 512
#endif
#ifndef SB_RADIX_ITEMS
#define SB_RADIX_ITEMS 8
#endif
// 512 threads x 8 keys: the same 4096-key tile as 256 x 16, but ~40 registers per thread,
// so three CTAs fit an SM and all tiles of a 1M-key sort are resident at once (the
// decoupled look-back never waits for a tile that has not been scheduled yet)
constexpr int THREADS = SB_RADIX_THREADS; // >= 256: thread d < 256 owns digit d in the scan phases
constexpr int WARPS = THREADS / 32;
constexpr int ITEMS = SB_RADIX_ITEMS;
constexpr int TILE = THREADS * ITEMS; // keys per CTA
static_assert(THREADS >= 256 && THREADS % 32 == 0, "one thread per digit");
constexpr int MAX_PASSES = 8;

constexpr uint32_t FLAG_AGG = 1u << 30;
constexpr uint32_t FLAG_PREFIX = 2u << 30;
constexpr uint32_t VALUE_MASK = (1u << 30) - 1;

// Tail of a histogram pass (hist_kernel below, or a producer kernel that counted the digits of the keys
// it wrote -- tri_prepare_kernel of sb_build.cu): the CTA's shared-memory counts are added to the global
// histogram; the last CTA to finish turns each 256-bin histogram into an exclusive prefix (global digit
// offsets).  Called by all threads of the CTA, blockDim.x >= 256.
template <int BITS>
__device__ __forceinline__ void hist_flush_and_scan(uint32_t *s_hist, int passes, uint32_t *__restrict__ hist /* [passes][256] */,
    uint32_t *__restrict__ ticket)
{
    constexpr int RADIX = 1 << BITS;
    __shared__ bool s_last;
    __syncthreads();
    for (int i = threadIdx.x; i < passes * RADIX; i += blockDim.x)
        if (s_hist[i])
            atomicAdd(&hist[i], s_hist[i]);
    if (!ticket)
        return; // one of several launches that count into the same histogram: hist_scan_kernel finishes it
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0)
        s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last)
        return;
    __threadfence();
    // exclusive scan of each pass's bins (thread d < RADIX owns bin d)
    static_assert(RADIX == 256, "one thread per digit");
    __shared__ uint32_t s_scan[RADIX];
    const bool own = threadIdx.x < RADIX;
    for (int p = 0; p < passes; ++p) {
        uint32_t v = own ? __ldcg(&hist[p * RADIX + threadIdx.x]) : 0u;
        if (own)
            s_scan[threadIdx.x] = v;
        __syncthreads();
        for (int off = 1; off < RADIX; off <<= 1) {
            uint32_t t = (own && (int)threadIdx.x >= off) ? s_scan[threadIdx.x - off] : 0;
            __syncthreads();
            if (own)
                s_scan[threadIdx.x] += t;
            __syncthreads();
        }
        if (own)
            hist[p * RADIX + threadIdx.x] = s_scan[threadIdx.x] - v;
        __syncthreads();
    }
}

// the exclusive scan on its own (one CTA of 256 threads), after several launches counted into the same histogram
static __global__ void __launch_bounds__(256) hist_scan_kernel(uint32_t *__restrict__ hist, int passes)
{
    __shared__ uint32_t s_scan[256];
    for (int p = 0; p < passes; ++p) {
        const uint32_t v = hist[p * 256 + threadIdx.x];
        s_scan[threadIdx.x] = v;
        __syncthreads();
        for (int off = 1; off < 256; off <<= 1) {
            const uint32_t t = (int)threadIdx.x >= off ? s_scan[threadIdx.x - off] : 0;
            __syncthreads();
            s_scan[threadIdx.x] += t;
            __syncthreads();
        }
        hist[p * 256 + threadIdx.x] = s_scan[threadIdx.x] - v;
        __syncthreads();
    }
}

// Histogram of every digit position in one pass over the keys.
template <typename KeyT, int BITS>
__global__ void __launch_bounds__(THREADS) hist_kernel(const KeyT *__restrict__ keys, uint32_t n,
    int beginBit, int passes, uint32_t *__restrict__ hist /* [passes][256] */, uint32_t *__restrict__ ticket)
{
    constexpr int RADIX = 1 << BITS;
    constexpr int MAXP = BITS == 8 ? MAX_PASSES : 4;
    __shared__ uint32_t s_hist[MAXP * RADIX];
    for (int i = threadIdx.x; i < passes * RADIX; i += THREADS)
        s_hist[i] = 0;
    __syncthreads();
    for (uint32_t i = blockIdx.x * THREADS + threadIdx.x; i < n; i += gridDim.x * THREADS) {
        KeyT k = keys[i];
        for (int p = 0; p < passes; ++p)
            atomicAdd(&s_hist[p * RADIX + (uint32_t)((k >> (beginBit + BITS * p)) & (RADIX - 1))], 1u);
    }
    hist_flush_and_scan<BITS>(s_hist, passes, hist, ticket);
}

template <typename KeyT, bool HAS_VALUES, int BITS>
__global__ void __launch_bounds__(THREADS, 3) onesweep_kernel(const KeyT *__restrict__ keysIn, KeyT *__restrict__ keysOut,
    const uint32_t *__restrict__ valsIn, uint32_t *__restrict__ valsOut, uint32_t n, int shift,
    const uint32_t *__restrict__ digitBase /* [RADIX] exclusive */, volatile uint32_t *lookback /* [tiles][RADIX] */,
    uint32_t *__restrict__ tileCounter)
{
    constexpr int RADIX = 1 << BITS;
    static_assert(RADIX == 256, "one thread per digit");
    __shared__ uint16_t s_warpHist[WARPS][RADIX]; // counts <= 32 * ITEMS, prefixes <= TILE
    static_assert(TILE <= 65535, "16-bit per-warp digit counters");
    __shared__ uint32_t s_digitStart[RADIX];
    __shared__ uint32_t s_globalOff[RADIX];
    __shared__ uint32_t s_scanTmp[RADIX / 32];
    __shared__ uint32_t s_tile;
    __shared__ __align__(16) unsigned char s_raw[TILE * sizeof(KeyT)];
    KeyT *s_keys = reinterpret_cast<KeyT *>(s_raw);
    uint32_t *s_vals = reinterpret_cast<uint32_t *>(s_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0)
        s_tile = atomicAdd(tileCounter, 1u);
    for (int i = tid; i < WARPS * RADIX / 2; i += THREADS)
        reinterpret_cast<uint32_t *>(&s_warpHist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t tileBase = tile * TILE;
    const uint32_t tileCount = min((uint32_t)TILE, n - tileBase);

    // striped load: item i of lane l sits at warpBase + i*32 + l (memory order =
    // (i, l) lexicographic, which is the order ranks are handed out in -> stable)
    KeyT key[ITEMS];
    uint32_t rank[ITEMS];
    const uint32_t warpBase = tileBase + warp * (32 * ITEMS);
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        uint32_t idx = warpBase + i * 32 + lane;
        key[i] = idx < n ? keysIn[idx] : (KeyT)~(KeyT)0;
    }
    const uint32_t ltMask = lanemask_lt();
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        uint32_t digit = (uint32_t)((key[i] >> shift) & (RADIX - 1));
        uint32_t peers = __match_any_sync(SB_FULL, digit);
        int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (lane == leader) {
            old = s_warpHist[warp][digit];
            s_warpHist[warp][digit] = (uint16_t)(old + __popc(peers));
        }
        old = __shfl_sync(SB_FULL, old, leader);
        rank[i] = old + __popc(peers & ltMask);
        __syncwarp();
    }
    __syncthreads();

    // thread d < RADIX owns digit d: exclusive scan over the warps, tile aggregate, look-back
    {
        const bool own = tid < RADIX;
        uint32_t sum = 0, excl = 0;
        if (own) {
#pragma unroll
            for (int w = 0; w < WARPS; ++w) {
                uint32_t t = s_warpHist[w][tid];
                s_warpHist[w][tid] = (uint16_t)sum;
                sum += t;
            }
            if (tile == 0)
                lookback[tid] = sum | FLAG_PREFIX;
            else
                lookback[tile * RADIX + tid] = sum | FLAG_AGG;
            if (tile != 0) {
                // decoupled look-back, LOOK predecessors per round trip (all tiles of a sort are
                // resident, so the walk back to the nearest published prefix can be long)
                constexpr int LOOK = 8;
                int t = (int)tile - 1;
                bool done = false;
                while (!done) {
                    uint32_t v[LOOK];
#pragma unroll
                    for (int k = 0; k < LOOK; ++k) {
                        v[k] = FLAG_PREFIX + 0u; // before the first tile: an empty prefix
                        if (t - k >= 0)
                            v[k] = lookback[(t - k) * RADIX + tid];
                    }
#pragma unroll
                    for (int k = 0; k < LOOK; ++k) {
                        if (done || v[k] == 0)
                            break; // not published yet: look again from here
                        excl += v[k] & VALUE_MASK;
                        done = (v[k] & FLAG_PREFIX) != 0;
                        --t;
                    }
                }
                lookback[tile * RADIX + tid] = (excl + sum) | FLAG_PREFIX;
            }
        }
        // block-wide exclusive scan of the digit totals (warps 0..7 hold them)
        uint32_t incl = sum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t t = __shfl_up_sync(SB_FULL, incl, off);
            if (lane >= off)
                incl += t;
        }
        if (own && lane == 31)
            s_scanTmp[warp] = incl;
        __syncthreads();
        if (own) {
            uint32_t warpOff = 0;
#pragma unroll
            for (int w = 0; w < RADIX / 32; ++w)
                if (w < warp)
                    warpOff += s_scanTmp[w];
            const uint32_t start = warpOff + incl - sum;
            s_digitStart[tid] = start;
            s_globalOff[tid] = digitBase[tid] + excl - start;
        }
    }
    __syncthreads();

    // local scatter so that the global writes below are runs of consecutive addresses
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        uint32_t digit = (uint32_t)((key[i] >> shift) & (RADIX - 1));
        rank[i] += s_digitStart[digit] + s_warpHist[warp][digit];
        s_keys[rank[i]] = key[i];
    }
    __syncthreads();
    uint32_t gpos[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        uint32_t j = tid + i * THREADS;
        KeyT k = s_keys[j];
        uint32_t digit = (uint32_t)((k >> shift) & (RADIX - 1));
        gpos[i] = s_globalOff[digit] + j;
        if (j < tileCount)
            keysOut[gpos[i]] = k;
    }
    if (HAS_VALUES) {
        __syncthreads();
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            uint32_t idx = warpBase + i * 32 + lane;
            s_vals[rank[i]] = idx < n ? valsIn[idx] : 0u;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            uint32_t j = tid + i * THREADS;
            if (j < tileCount)
                valsOut[gpos[i]] = s_vals[j];
        }
    }
}

struct Workspace {
    uint32_t *mem = nullptr; // [hist: MAX_PASSES*512][ticket+counters: 16][lookback: passes * tiles * RADIX]
    static size_t words(size_t tiles) { return MAX_PASSES * 512 + 16 + MAX_PASSES * tiles * 512; }
};

inline size_t tiles_for(size_t n) { return (n + TILE - 1) / TILE; }

// Sort n keys on bits [beginBit, endBit) with BITS-bit digits.  Results land in
// *outKeys/*outVals (either the primary or the tmp buffers).  Returns the number of
// kernels launched.
//
// histByProducer: the kernel that wrote the keys also counted their digits (hist_flush_and_scan); the caller
// then cleared the workspace with sort_clear BEFORE that kernel ran and passes the hist / ticket pointers
// of sort_hist / sort_ticket to it.
inline int sort_passes(size_t n, int beginBit, int endBit, int bits)
{
    if (n < 2 || endBit <= beginBit)
        return 0;
    const int maxp = bits == 8 ? MAX_PASSES : 4;
    const int passes = (endBit - beginBit + bits - 1) / bits;
    return passes > maxp ? maxp : passes;
}
inline void sort_clear(cudaStream_t stream, const Workspace &ws, size_t n, int passes)
{
    cudaMemsetAsync(ws.mem, 0, sizeof(uint32_t) * (MAX_PASSES * 512 + 16 + (size_t)passes * tiles_for(n) * 256), stream);
}
inline uint32_t *sort_hist(const Workspace &ws) { return ws.mem; }
inline uint32_t *sort_ticket(const Workspace &ws) { return ws.mem + MAX_PASSES * 512; }

template <typename KeyT, int BITS>
int sort(cudaStream_t stream, KeyT *keys, KeyT *keysTmp, uint32_t *vals, uint32_t *valsTmp, size_t n,
    int beginBit, int endBit, const Workspace &ws, int smCount, KeyT **outKeys, uint32_t **outVals, bool histByProducer = false)
{
    constexpr int RADIX = 1 << BITS;
    constexpr int MAXP = BITS == 8 ? MAX_PASSES : 4;
    *outKeys = keys;
    if (outVals)
        *outVals = vals;
    if (n < 2 || endBit <= beginBit)
        return 0;
    int passes = (endBit - beginBit + BITS - 1) / BITS;
    if (passes > MAXP)
        passes = MAXP;
    size_t tiles = tiles_for(n);
    uint32_t *hist = ws.mem;
    uint32_t *ticket = ws.mem + MAX_PASSES * 512;
    uint32_t *counters = ticket + 1;
    uint32_t *lookback = ws.mem + MAX_PASSES * 512 + 16;
    if (!histByProducer) {
        cudaMemsetAsync(ws.mem, 0, sizeof(uint32_t) * (MAX_PASSES * 512 + 16 + (size_t)passes * tiles * RADIX), stream);
        int histBlocks = (int)((n + THREADS * 8 - 1) / (THREADS * 8));
        if (histBlocks > smCount * 4)
            histBlocks = smCount * 4;
        if (histBlocks < 1)
            histBlocks = 1;
        hist_kernel<KeyT, BITS><<<histBlocks, THREADS, 0, stream>>>(keys, (uint32_t)n, beginBit, passes, hist, ticket);
    }
    KeyT *kin = keys, *kout = keysTmp;
    uint32_t *vin = vals, *vout = valsTmp;
    for (int p = 0; p < passes; ++p) {
        if (vals)
            onesweep_kernel<KeyT, true, BITS><<<(unsigned)tiles, THREADS, 0, stream>>>(kin, kout, vin, vout, (uint32_t)n,
                beginBit + BITS * p, hist + p * RADIX, lookback + (size_t)p * tiles * RADIX, counters + p);
        else
            onesweep_kernel<KeyT, false, BITS><<<(unsigned)tiles, THREADS, 0, stream>>>(kin, kout, nullptr, nullptr,
                (uint32_t)n, beginBit + BITS * p, hist + p * RADIX, lookback + (size_t)p * tiles * RADIX, counters + p);
        KeyT *tk = kin; kin = kout; kout = tk;
        uint32_t *tv = vin; vin = vout; vout = tv;
    }
    *outKeys = kin;
    if (outVals)
        *outVals = vin;
    return (histByProducer ? 0 : 1) + passes;
}

} // namespace sbradix
