// Axis-projected uniform grids for the inside/outside rays.
//
// The three test rays of SolidBoolean::isPointInMesh (reference
// src/solidboolean.cpp:31-35, 48-92) are axis aligned, so "all triangles whose
// box overlaps the ray box" is a 2-D point-location problem in the plane
// perpendicular to the ray plus a half-line test along it.  For every mesh and
// every axis we therefore bin the triangles' boxes into a 2-D grid over the two
// perpendicular dimensions (CSR layout: ranges per cell and depth slab -- sb_gridq.cuh -- into one array of
// 8-byte cell-relative references).  A ray then reads the part of ONE cell list that lies at or beyond its own
// depth instead of walking a tree.
//
// Exactness: cells and the 15-bit coordinates inside a reference are produced by
// one monotone quantiser per world axis (sb_gridq.cuh), applied to triangle
// bounds and to ray bounds alike, so the quantised test can only over-accept; the
// exact double test of the reference follows in the classifier.
//
// Build = count (atomics; done by the leaf kernel of sb_build.cu while the box is in
// registers, together with the 15-bit quantised box it stores) -> in-place chained
// scan -> fill (atomics), triangles visited in Morton order so neighbouring threads
// hit neighbouring cells.
#include "sb_internal.h"
#include "sb_gridq.cuh"
#include <algorithm>

#ifndef SB_FILL_AGG
#define SB_FILL_AGG 1 // warp-aggregated slot claims in the fill pass (build 0.562 -> 0.544 ms at C3)
#endif

namespace {


// Cell size = beta x mean triangle-box extent along each world axis, so a
// triangle covers about (1 + 1/beta)^2 cells on every grid whatever the mesh's
// aspect ratio or anisotropy; resolutions are powers of two so that a cell index
// is a shift of the 15-bit coordinate.  maxBits bounds cells per axis (allocation).
__global__ void grid_params_kernel(const unsigned long long *__restrict__ bounds,
    const unsigned long long *__restrict__ extentSum,
    uint32_t nT, int maxBits, float beta, int batch, double latPitch, int naxes, int slabBitsMax, int *err, GridParams *out)
{
    // launched with one warp: the 32 partial extent sums per axis are read side by side
    unsigned long long isumAxis[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        unsigned long long v = extentSum[3 * (threadIdx.x & 31) + d];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
            v += __shfl_xor_sync(SB_FULL, v, off);
        isumAxis[d] = v;
    }
    if (threadIdx.x != 0 || blockIdx.x != 0)
        return;
    GridParams g;
    // batch mesh: 10 job-local bits per axis under 5 lattice bits (sb_gridq.cuh); a cell never spans two jobs
    g.latShift = batch ? SB_BATCH_LAT_SHIFT : 0;
    g.qmax = batch ? 1023.0 : 32767.0;
    const int localBits = batch ? SB_BATCH_LAT_SHIFT : SB_Q_BITS, latBits = batch ? SB_Q_BITS - SB_BATCH_LAT_SHIFT : 0;
    int bitsWanted[3];
    for (int d = 0; d < 3; ++d) {
        double lo = dkey_inv(bounds[d]), hi = dkey_inv(bounds[3 + d]);
        double ext = hi - lo;
        g.org[d] = lo;
        // hi maps to 32767.99..; finite and > 0 extents only
        g.scl[d] = (ext > 0.0 && ext < 1.0e300) ? (g.qmax + 0.999) / ext : 0.0;
        g.lo[d] = lo;
        g.hi[d] = hi;
        // batch: jobs one lattice step apart must not meet, whatever their place inside [-M, M]
        if (batch && !(4.0 * fmax(fabs(lo), fabs(hi)) <= latPitch))
            atomicOr(err, 2);
        const unsigned long long isum = isumAxis[d]; // fixed point: 2^-24 fractions of the mesh extent
        double mean = nT ? (double)isum / 16777216.0 * ext / (double)nT : 0.0;
        double cells = (ext > 0.0 && mean > 0.0) ? ext / ((double)beta * mean) : 1.0;
        int b = (int)floor(log2(fmax(cells, 1.0)) + 0.5);
        bitsWanted[d] = max(0, min(b, localBits)) + latBits;
    }
    uint32_t base = 0;
    for (int a = 0; a < 3; ++a) {
        int u = a == 0 ? 1 : 0, v = a == 2 ? 1 : 2;
        int ku = bitsWanted[u], kv = bitsWanted[v];
        while (ku + kv > maxBits) { // shrink the finer dimension first
            if (ku >= kv && ku > latBits) --ku;
            else if (kv > latBits) --kv;
            else if (ku > latBits) --ku;
            else break;
        }
        g.shiftU[a] = SB_Q_BITS - ku;
        g.shiftV[a] = SB_Q_BITS - kv;
        g.nu[a] = 1u << ku;
        g.cellBase[a] = base;
        // depth slabs (sb_gridq.cuh): about 2 nT / cells of them per cell (a cell holds ~3.6 nT / cells references:
        // a couple per sub-list; E stays within ~2 nT entries per grid), at most 2^slabBitsMax, and inside the
        // allocation bound of 2^maxBits entries per grid; never more than the job-local bits of the coordinate
        int sb = 0;
        while (sb < slabBitsMax && sb < localBits && ku + kv + sb + 1 <= maxBits && ((size_t)3 << (ku + kv + sb)) <= (size_t)4 * nT)
            ++sb;
        g.slabBits[a] = (uint32_t)sb;
        base += 1u << (ku + kv + sb);
        base = (base + 15u) & ~15u; // a scan thread's 16 entries never straddle two grids (different slab counts)
        if (a + 1 == naxes)
            g.usedCells = base;
    }
    g.totalCells = base;
    if (naxes >= 3)
        g.usedCells = base;
    *out = g;
}

// Fill pass: one thread per (triangle, axis) -- three times the parallelism and a third
// of the serial atomic chain of a per-triangle loop.  Triangles are visited in Morton
// order, so neighbouring threads update neighbouring cells.  The quantised boxes were
// formed (and the cells counted) by the leaf kernel of sb_build.cu.
__global__ void __launch_bounds__(256) grid_fill_kernel(const uint4 *__restrict__ qbox, uint32_t nT,
    const GridParams *__restrict__ gp, uint32_t *__restrict__ E, uint2 *__restrict__ refs, uint32_t refCap,
    uint4 *__restrict__ bigRefs, uint32_t *__restrict__ bigCount, uint32_t bigCap)
{
    __shared__ GridParams g;
    if (threadIdx.x < sizeof(GridParams) / 4)
        reinterpret_cast<uint32_t *>(&g)[threadIdx.x] = reinterpret_cast<const uint32_t *>(gp)[threadIdx.x];
    __syncthreads();
    const uint32_t blocksPerAxis = (nT + 255) / 256;
    const int a = blockIdx.x / blocksPerAxis;
    const uint32_t j = (blockIdx.x % blocksPerAxis) * 256 + threadIdx.x;
    if (j >= nT)
        return;
    const uint4 q = __ldg(qbox + j);
    const GridFootprint f = grid_footprint(q, g, a);
    const uint32_t cu0 = f.cu0, cu1 = f.cu1, cv0 = f.cv0, cv1 = f.cv1;
    const uint32_t qlu = f.qu & 0xffffu, qhu = f.qu >> 16, qlv = f.qv & 0xffffu, qhv = f.qv >> 16, qla = f.qa & 0xffffu, qha = f.qa >> 16;
    if ((cu1 - cu0 + 1) * (cv1 - cv0 + 1) > SB_GRID_MAX_CELLS_PER_TRI) {
        uint32_t slot = atomicAdd(&bigCount[3 + a], 1u);
        if (slot < bigCap)
            bigRefs[(size_t)a * bigCap + slot] = grid_ref_pack(qlu, qhu, qlv, qhv, qla, qha, q.w);
        return;
    }
    const uint32_t nu = g.nu[a];
    const int su = g.shiftU[a], sv = g.shiftV[a], sb = (int)g.slabBits[a];
    const uint32_t base = g.cellBase[a] + grid_slab(qha, g, a) + 1; // entry of (cell, slab of the far bound) = base + (cell << sb)
    // the reference of this triangle in cell (cu, cv): its box clipped to the cell, cell-relative
    auto ref_in = [&](uint32_t cu, uint32_t cv) {
        return cell_ref_pack(qlu, qhu, qlv, qhv, qha, cu, cv, su, sv, cu > cu0, cv > cv0, q.w);
    };
    if (cu1 - cu0 <= 1 && cv1 - cv0 <= 1) {
        // common case (footprint at most 2 x 2 cells): all atomics are issued before the
        // first dependent store, so their round trips overlap instead of adding up
        const bool du = cu1 != cu0, dv = cv1 != cv0;
        const uint32_t c00 = cv0 * nu + cu0;
#if SB_FILL_AGG
        // Morton-neighbouring triangles mostly land in the same cells: one atomic per distinct cell
        // of the warp, the lanes that share it take consecutive slots below the returned end
        // (the four atomics of a lane are issued before the first of them is waited for: grouping first,
        // then the leaders' atomics, then the shuffles that hand the results round -- one round trip, not four)
        const unsigned act = __activemask();
        const uint32_t lane = threadIdx.x & 31, below = lanemask_lt();
        const bool want[4] = {true, du, dv, du && dv};
        const uint32_t cell[4] = {base + (c00 << sb), base + ((c00 + 1) << sb), base + ((c00 + nu) << sb), base + ((c00 + nu + 1) << sb)};
        unsigned peers[4];
        uint32_t end[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 4; ++k)
            peers[k] = __match_any_sync(act, want[k] ? cell[k] : 0xffffffffu - lane);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (want[k] && (int)lane == __ffs(peers[k]) - 1)
                end[k] = atomicSub(&E[cell[k]], (uint32_t)__popc(peers[k]));
        uint32_t pos[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            pos[k] = __shfl_sync(act, end[k], __ffs(peers[k]) - 1) - 1u - (uint32_t)__popc(peers[k] & below);
        const uint32_t p00 = pos[0], p10 = pos[1], p01 = pos[2], p11 = pos[3];
#else
        uint32_t p00 = atomicSub(&E[base + (c00 << sb)], 1u) - 1u, p10 = 0, p01 = 0, p11 = 0;
        if (du) p10 = atomicSub(&E[base + ((c00 + 1) << sb)], 1u) - 1u;
        if (dv) p01 = atomicSub(&E[base + ((c00 + nu) << sb)], 1u) - 1u;
        if (du && dv) p11 = atomicSub(&E[base + ((c00 + nu + 1) << sb)], 1u) - 1u;
#endif
        if (p00 < refCap) refs[p00] = ref_in(cu0, cv0);
        if (du && p10 < refCap) refs[p10] = ref_in(cu1, cv0);
        if (dv && p01 < refCap) refs[p01] = ref_in(cu0, cv1);
        if (du && dv && p11 < refCap) refs[p11] = ref_in(cu1, cv1);
        return;
    }
    for (uint32_t cv = cv0; cv <= cv1; ++cv)
        for (uint32_t cu = cu0; cu <= cu1; ++cu) {
            uint32_t pos = atomicSub(&E[base + ((cv * nu + cu) << sb)], 1u) - 1u; // fill each sub-list back to front
            if (pos < refCap)
                refs[pos] = ref_in(cu, cv);
        }
}

// Count pass on its own (the leaf kernel normally does it): used when a mesh built with
// two grids needs the third one after all.
__global__ void __launch_bounds__(256) grid_count_kernel(const uint4 *__restrict__ qbox, uint32_t nT,
    const GridParams *__restrict__ gp, uint32_t *__restrict__ E, uint32_t *__restrict__ bigCount, int naxes)
{
    __shared__ GridParams g;
    if (threadIdx.x < sizeof(GridParams) / 4)
        reinterpret_cast<uint32_t *>(&g)[threadIdx.x] = reinterpret_cast<const uint32_t *>(gp)[threadIdx.x];
    __syncthreads();
    const uint32_t j = blockIdx.x * 256 + threadIdx.x;
    if (j < nT)
        grid_count_tri(__ldg(qbox + j), g, E, bigCount, naxes);
}

// In-place inclusive scan of a u32 array (single pass, chained tiles with
// decoupled look-back -- same protocol as the radix sort's digit offsets).
#ifndef SB_SCAN_THREADS
#define SB_SCAN_THREADS 512
#endif
#ifndef SB_SCAN_ITEMS
#define SB_SCAN_ITEMS 16
#endif
// 8192 cells per tile: the look-back chain of a 2M-cell grid is 256 tiles long (it was 1024 with 2048-cell
// tiles, and the chain -- not the 8 MB of traffic -- set the kernel's 23 us)
constexpr int SCAN_THREADS = SB_SCAN_THREADS;
constexpr int SCAN_ITEMS = SB_SCAN_ITEMS;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
constexpr unsigned long long SCAN_AGG = 1ull << 62, SCAN_PREFIX = 2ull << 62, SCAN_MASK = (1ull << 62) - 1;

template <int K>
__device__ __forceinline__ void pad_cells(uint32_t (&v)[SCAN_ITEMS])
{
    static_assert(SCAN_ITEMS % K == 0, "whole cells per thread");
#pragma unroll
    for (int c = 0; c < SCAN_ITEMS; c += K) {
        uint32_t t = 0;
#pragma unroll
        for (int k = 0; k < K; ++k)
            t += v[c + k];
        v[c] += t & 1u;
    }
}

// zeroes E[0 .. usedCells + 1] (the counts of the grids that are binned, and the closing entry)
__global__ void __launch_bounds__(256) grid_clear_kernel(uint32_t *__restrict__ E, const GridParams *__restrict__ gp)
{
    const uint32_t n4 = (gp->usedCells + 1 + 3) / 4; // uint4 groups from &E[1] (aligned), after E[0]
    if (blockIdx.x == 0 && threadIdx.x == 0)
        E[0] = 0u;
    uint4 *p = reinterpret_cast<uint4 *>(E + 1);
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n4; i += gridDim.x * 256)
        p[i] = make_uint4(0u, 0u, 0u, 0u);
}

__global__ void grid_set_used_kernel(GridParams *gp, int naxes)
{
    gp->usedCells = naxes >= 3 ? gp->totalCells : gp->cellBase[naxes];
}

__global__ void __launch_bounds__(SCAN_THREADS) inclusive_scan_kernel(uint32_t *__restrict__ data,
    const GridParams *__restrict__ gp, volatile unsigned long long *status, uint32_t *__restrict__ tileCounter,
    uint32_t *__restrict__ totalOut)
{
    // data = E + 1 (16-byte aligned): data[e] = number of references of entry e; E[0] stays 0
    const uint32_t n = gp->usedCells;
    if ((unsigned long long)blockIdx.x * SCAN_TILE >= n)
        return; // launched for the allocation bound; only the first ceil(n / TILE) CTAs take a ticket
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_warp[SCAN_THREADS / 32];
    __shared__ unsigned long long s_excl;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0)
        s_tile = atomicAdd(tileCounter, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t base = tile * SCAN_TILE + tid * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t sum = 0;
    static_assert(SCAN_ITEMS == 16, "128-bit loads; whole cells per thread (pad_cells, grid bases are multiples of 16)");
    const bool whole = base + SCAN_ITEMS <= n; // (the array is only allocated up to the cell bound + 2)
    if (whole) {
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; i += 4) {
            const uint4 q = *reinterpret_cast<const uint4 *>(data + base + i);
            v[i] = q.x; v[i + 1] = q.y; v[i + 2] = q.z; v[i + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i)
            v[i] = base + i < n ? data[base + i] : 0u;
    }
    // Every CELL's share (its 1 << slabBits sub-lists together) is rounded up to an even number of
    // references: the cell ends (and the 16-byte pairs the classifier loads) stay aligned.  The unused
    // slot of an odd cell is put in FRONT of its first sub-list, so the sub-lists of a cell stay
    // contiguous up to the cell's (even) end.  A thread's 16 entries lie inside one grid (grid bases are
    // multiples of 16) and hold whole cells (1 << slabBits divides 16).
    {
        const uint32_t sbits = base >= gp->cellBase[2] ? gp->slabBits[2] : base >= gp->cellBase[1] ? gp->slabBits[1] : gp->slabBits[0];
        switch (sbits) {
        case 0: pad_cells<1>(v); break;
        case 1: pad_cells<2>(v); break;
        case 2: pad_cells<4>(v); break;
        case 3: pad_cells<8>(v); break;
        default: pad_cells<16>(v); break;
        }
    }
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        sum += v[i];
        v[i] = sum; // inclusive within the thread
    }
    uint32_t incl = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        uint32_t t = __shfl_up_sync(SB_FULL, incl, off);
        if (lane >= off)
            incl += t;
    }
    if (lane == 31)
        s_warp[warp] = incl;
    __syncthreads();
    uint32_t warpOff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) {
        if (w < warp)
            warpOff += s_warp[w];
        total += s_warp[w];
    }
    if (warp == 0) {
        // decoupled look-back by the whole first warp: 32 predecessors per round trip (lane k reads tile t - k),
        // everything up to the nearest published inclusive prefix is summed in one go
        unsigned long long excl = 0;
        if (tile == 0) {
            if (lane == 0)
                status[0] = (unsigned long long)total | SCAN_PREFIX;
        } else {
            if (lane == 0)
                status[tile] = (unsigned long long)total | SCAN_AGG;
            int t = (int)tile - 1;
            for (;;) {
                const int idx = t - lane;
                const unsigned long long sv = idx >= 0 ? status[idx] : SCAN_PREFIX + 0ull; // before the first tile: an empty prefix
                const unsigned notReady = __ballot_sync(SB_FULL, sv == 0);
                const unsigned isPrefix = __ballot_sync(SB_FULL, (sv & SCAN_PREFIX) != 0);
                const unsigned usable = notReady ? (1u << (__ffs(notReady) - 1)) - 1u : 0xffffffffu; // nearer than the first unpublished tile
                const unsigned pfx = isPrefix & usable;
                const unsigned take = pfx ? (2u << (__ffs(pfx) - 1)) - 1u : usable; // up to and including the nearest prefix
                unsigned long long v = ((take >> lane) & 1u) ? (sv & SCAN_MASK) : 0ull;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1)
                    v += __shfl_xor_sync(SB_FULL, v, d);
                excl += v;
                if (pfx)
                    break;
                t -= __popc(take); // (nothing usable: look again from the same place)
            }
            if (lane == 0)
                status[tile] = (excl + total) | SCAN_PREFIX;
        }
        if (lane == 0)
            s_excl = excl;
    }
    __syncthreads();
    const uint32_t off = (uint32_t)s_excl + warpOff + incl - sum;
    if (whole) {
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; i += 4)
            *reinterpret_cast<uint4 *>(data + base + i) = make_uint4(off + v[i], off + v[i + 1], off + v[i + 2], off + v[i + 3]);
    } else {
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i)
            if (base + i < n)
                data[base + i] = off + v[i];
    }
    if (base < n && n - 1 - base < SCAN_ITEMS) { // this thread holds the last entry's end = number of references
        uint32_t last = 0;
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i)
            if (base + i == n - 1)
                last = off + v[i];
        data[n] = last; // E[usedCells + 1]: end of the last cell after the fill
        *totalOut = last;
    }
}

} // namespace

size_t sbk_grid_entry_bound(uint32_t gridCellBits)
{
    return ((size_t)3 << gridCellBits) + 64; // three grids of at most 2^bits entries, each rounded up to a multiple of 16
}

size_t sbk_grid_scan_status_words(uint32_t gridCellBits)
{
    size_t tiles = (sbk_grid_entry_bound(gridCellBits) + SCAN_TILE - 1) / SCAN_TILE;
    return 2 * tiles + 4; // u64 status per tile + the tile counter
}

static void grid_clear(cudaStream_t s, MeshDev &m, uint32_t *scanScratch, LaunchCounter &lc)
{
    cudaMemsetAsync(m.gridBigCount, 0, sizeof(uint32_t) * 8, s);
    cudaMemsetAsync(scanScratch, 0, sizeof(uint32_t) * sbk_grid_scan_status_words(m.gridCellBits), s);
    // only the entries in use (the allocation bound is ~8x what a typical mesh takes)
    const size_t bound = sbk_grid_entry_bound(m.gridCellBits);
    const uint32_t blocks = (uint32_t)std::min<size_t>((bound / 4 + 255) / 256, 148 * 8);
    grid_clear_kernel<<<blocks, 256, 0, s>>>(m.gridE, m.gridParams);
    lc.kernels += 1;
}

// Phase 0 (before the leaf kernel, which counts): cleared counters + grid parameters.
cudaError_t sbk_grid_prepare(cudaStream_t s, MeshDev &m, uint32_t *scanScratch, float beta, int slabBitsMax, LaunchCounter &lc)
{
    if (m.nT == 0)
        return cudaSuccess;
    grid_params_kernel<<<1, 32, 0, s>>>(m.bounds, m.extentSum, m.nT, (int)m.gridCellBits, beta, m.triJob ? 1 : 0, m.latPitch,
        m.gridAxes, std::max(0, std::min(slabBitsMax, 4)), m.err, m.gridParams);
    lc.kernels += 1;
    grid_clear(s, m, scanScratch, lc);
    return cudaGetLastError();
}

cudaError_t sbk_grid_recount(cudaStream_t s, MeshDev &m, uint32_t *scanScratch, LaunchCounter &lc)
{
    if (m.nT == 0)
        return cudaSuccess;
    grid_set_used_kernel<<<1, 1, 0, s>>>(m.gridParams, m.gridAxes);
    lc.kernels += 1;
    grid_clear(s, m, scanScratch, lc);
    grid_count_kernel<<<(m.nT + 255) / 256, 256, 0, s>>>(m.qbox, m.nT, m.gridParams, m.gridE, m.gridBigCount, m.gridAxes);
    lc.kernels += 1;
    return cudaGetLastError();
}

// Phase 1 (after the leaf kernel): inclusive scan of the counts per entry e = (cell, depth slab).  Afterwards
// E[e + 1] = end of entry e and gridBigCount[6] = total number of references.
cudaError_t sbk_grid_scan(cudaStream_t s, MeshDev &m, uint32_t *scanScratch, LaunchCounter &lc)
{
    if (m.nT == 0)
        return cudaSuccess;
    size_t statusWords = sbk_grid_scan_status_words(m.gridCellBits);
    uint32_t tiles = (uint32_t)((sbk_grid_entry_bound(m.gridCellBits) + SCAN_TILE - 1) / SCAN_TILE);
    unsigned long long *status = reinterpret_cast<unsigned long long *>(scanScratch);
    uint32_t *counter = scanScratch + statusWords - 2;
    inclusive_scan_kernel<<<tiles, SCAN_THREADS, 0, s>>>(m.gridE + 1, m.gridParams, status, counter, m.gridBigCount + 6);
    lc.kernels += 1;
    return cudaGetLastError();
}

// Phase 2 (after the caller sized refs / bigRefs from the counts): every
// E[e + 1] turns from the END of entry e into its START as the sub-list is filled
// back to front; E[usedCells + 1] (written by the scan) closes the last cell.  The sub-lists of a cell are
// contiguous and end on an even slot (the cell's share is rounded up, the unused slot sits in FRONT of its first
// sub-list: pad_cells); what a ray reads is [E[e0 + slab + 1], E[e0 + slabs + 1] & ~1) (grid_ray_range): the share of the next cell may begin
// with an unused slot.
cudaError_t sbk_grid_fill(cudaStream_t s, MeshDev &m, LaunchCounter &lc)
{
    if (m.nT == 0)
        return cudaSuccess;
    grid_fill_kernel<<<m.gridAxes * ((m.nT + 255) / 256), 256, 0, s>>>(m.qbox, m.nT, m.gridParams, m.gridE, m.gridRefs,
        m.gridRefCap, m.gridBigRefs, m.gridBigCount, m.gridBigCap);
    lc.kernels += 1;
    return cudaGetLastError();
}
