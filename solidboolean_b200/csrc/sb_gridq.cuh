// The quantiser shared by the ray-grid build (sb_grid.cu, leaf kernel of sb_build.cu) and
// the classifier (sb_classify.cu), and the grid references built from it: the absolute
// 16-byte form (per-axis big lists) and the 8-byte cell-relative form (cell lists).
//
// One monotone 15-bit quantiser per world axis, q(x) = clamp(floor((x - org) * scl),
// 0, 32767), is applied to triangle bounds and to ray bounds alike.  Monotonicity
// gives lo <= x <= hi  =>  q(lo) <= q(x) <= q(hi): the quantised overlap test can
// only over-accept, and the exact double test of the reference follows.
//
// 15 bits (not 16) so that TWO coordinates fit one 32-bit word with a guard bit
// above each: a reference stores {lo_u, 32767-hi_u}, {lo_v, 32767-hi_v},
// {32767-hi_a, lo_a}; a ray stores the matching upper bounds {hi_u, 32767-lo_u},
// ... with both guard bits set.  (ray - ref) then never borrows across the halves
// and leaves a guard bit set exactly where ref <= ray, so the five comparisons of
// the closed-interval test are 3 subtractions, 2 logic ops and one compare.
#pragma once
#include "sb_common.cuh"
#include "sb_internal.h"

#define SB_Q_MAX 32767u
#define SB_Q_BITS 15
#define SB_Q_GUARD 0x80008000u

__device__ __forceinline__ uint32_t quant15(double x, double org, double scl, double qmax = 32767.0)
{
    // floor, clamp to [0, qmax], NaN -> 0: cvt.rmi.s32.f64 floors, saturates and turns NaN into 0, the
    // clamp is then two integer min/max (the fmin/fmax form costs a dozen NaN-proofing selects)
    const int t = __double2int_rd((x - org) * scl);
    return (uint32_t)min(max(t, 0), (int)qmax);
}

// Batch meshes: position of job j on a 32 x 32 x 32 lattice, chosen so that ANY two of the three
// coordinates identify the job (z = x + y mod 32) -- each ray grid bins two of them, so the jobs
// (at most 1024) never share a cell; the float boxes of the LBVH use all three.
#define SB_BATCH_MAX_JOBS 1024u
#define SB_BATCH_LAT_SHIFT 10
__host__ __device__ __forceinline__ uint32_t lattice3(uint32_t job, int d)
{
    const uint32_t lx = (job >> 5) & 31u, ly = job & 31u;
    return d == 0 ? lx : d == 1 ? ly : ((lx + ly) & 31u);
}

// quantised coordinate along world axis d of a point of job `job` (plain meshes: job 0, latShift 0)
__device__ __forceinline__ uint32_t quant_axis(double x, const GridParams &g, int d, uint32_t job)
{
    const uint32_t q = quant15(x, g.org[d], g.scl[d], g.qmax);
    return g.latShift ? q + (lattice3(job, d) << g.latShift) : q;
}

// reference of a triangle whose quantised box is [qlu,qhu] x [qlv,qhv] across the ray
// axis and [qla,qha] along it
__device__ __forceinline__ uint4 grid_ref_pack(uint32_t qlu, uint32_t qhu, uint32_t qlv, uint32_t qhv, uint32_t qla,
    uint32_t qha, uint32_t id)
{
    return make_uint4(qlu | ((SB_Q_MAX - qhu) << 16), qlv | ((SB_Q_MAX - qhv) << 16), (SB_Q_MAX - qha) | (qla << 16), id);
}


// a ray: [aU,bU] x [aV,bV] across (almost always one point), starting at aA along the axis
struct RayQ {
    uint32_t x, y, z;
};

__device__ __forceinline__ RayQ ray_pack(uint32_t aU, uint32_t bU, uint32_t aV, uint32_t bV, uint32_t aA)
{
    RayQ q;
    q.x = bU | ((SB_Q_MAX - aU) << 16) | SB_Q_GUARD;
    q.y = bV | ((SB_Q_MAX - aV) << 16) | SB_Q_GUARD;
    q.z = (SB_Q_MAX - aA) | (SB_Q_MAX << 16) | SB_Q_GUARD; // lo_a of the triangle is not tested
    return q;
}

// lo_u <= bU && hi_u >= aU && lo_v <= bV && hi_v >= aV && hi_a >= aA
__device__ __forceinline__ bool ray_ref_match(const RayQ &q, const uint4 &r)
{
    return (((q.x - r.x) & (q.y - r.y) & (q.z - r.z)) & SB_Q_GUARD) == SB_Q_GUARD;
}

// ---- 8-byte cell-relative references ------------------------------------------------
// Inside the list of ONE cell only the part of a triangle's box that lies in that cell
// matters, so the reference stores it relative to the cell, 5 bits per bound (the 15-bit
// coordinate minus the cell origin, shifted down to 32 steps per cell: still monotone, still
// only over-accepting), the far depth bound in 12 bits, and the triangle id in the
// remaining 25 bits:
//   x: [lo_u:5 g][31-hi_u:5 g][lo_v:5 g][31-hi_v:5 g][extU][extV][id >> 19 : 6]
//   y: [4095-hi_a:12 g][id & 0x7ffff : 19]
// (g = guard bit, zero in a reference, set in a ray).  extU / extV: the triangle also
// covers the cell before this one along u / v -- what the "first cell only" rule of rays
// that span several cells needs.  Half the bytes of the absolute 16-byte form in the grid
// fill and in every list walk; triangles on the per-axis big lists keep the 16-byte form.
#define SB_CELLREF_ID_BITS 25
#define SB_R_UV_GUARD 0x00820820u // bits 5, 11, 17, 23
#define SB_R_D_GUARD 0x00001000u  // bit 12 of .y
#define SB_R_ALL (SB_R_UV_GUARD | (SB_R_D_GUARD << 12))

// position of a 15-bit coordinate inside its cell, in 1/32 (or finer cells: exact) steps
__device__ __forceinline__ uint32_t cell_rel5(uint32_t q, uint32_t cell, int shift)
{
    const uint32_t r = q - (cell << shift);
    return shift > 5 ? r >> (shift - 5) : r;
}

__device__ __forceinline__ uint2 cell_ref_pack(uint32_t qlu, uint32_t qhu, uint32_t qlv, uint32_t qhv, uint32_t qha,
    uint32_t cu, uint32_t cv, int shiftU, int shiftV, bool extU, bool extV, uint32_t id)
{
    const uint32_t su = cu << shiftU, eu = su + (1u << shiftU) - 1u, sv = cv << shiftV, ev = sv + (1u << shiftV) - 1u;
    const uint32_t lu = cell_rel5(max(qlu, su), cu, shiftU), hu = cell_rel5(min(qhu, eu), cu, shiftU);
    const uint32_t lv = cell_rel5(max(qlv, sv), cv, shiftV), hv = cell_rel5(min(qhv, ev), cv, shiftV);
    uint2 r;
    r.x = lu | ((31u - hu) << 6) | (lv << 12) | ((31u - hv) << 18) | ((extU ? 1u : 0u) << 24) | ((extV ? 1u : 0u) << 25) |
          ((id >> 19) << 26);
    r.y = (4095u - (qha >> 3)) | ((id & 0x7ffffu) << 13);
    return r;
}

__device__ __forceinline__ uint32_t cell_ref_id(const uint2 &r) { return ((r.x >> 26) << 19) | (r.y >> 13); }
__device__ __forceinline__ bool cell_ref_ext_u(const uint2 &r) { return (r.x >> 24) & 1u; }
__device__ __forceinline__ bool cell_ref_ext_v(const uint2 &r) { return (r.x >> 25) & 1u; }

// a ray in cell (cu, cv): [aU,bU] x [aV,bV] across (15-bit, inside the cell), starting at aA along the axis
struct CellRay {
    uint32_t x, y;
};

__device__ __forceinline__ CellRay cell_ray_pack(uint32_t aU, uint32_t bU, uint32_t aV, uint32_t bV, uint32_t aA, uint32_t cu,
    uint32_t cv, int shiftU, int shiftV)
{
    CellRay q;
    q.x = cell_rel5(bU, cu, shiftU) | ((31u - cell_rel5(aU, cu, shiftU)) << 6) | (cell_rel5(bV, cv, shiftV) << 12) |
          ((31u - cell_rel5(aV, cv, shiftV)) << 18) | SB_R_UV_GUARD;
    q.y = (4095u - (aA >> 3)) | SB_R_D_GUARD;
    return q;
}

// lo_u <= bU && hi_u >= aU && lo_v <= bV && hi_v >= aV && hi_a >= aA: no field of (ray - ref)
// borrows from its neighbour, and a guard bit survives exactly where ref <= ray
__device__ __forceinline__ bool cell_ref_match(const CellRay &q, const uint2 &r)
{
    return (((q.x - r.x) & SB_R_UV_GUARD) | (((q.y - r.y) & SB_R_D_GUARD) << 12)) == SB_R_ALL;
}

// ---- binning (shared by the leaf kernel of sb_build.cu, which counts, and sb_grid.cu, which fills) ----
#define SB_GRID_MAX_CELLS_PER_TRI 1024u // larger footprints go to the per-axis "big" list
#ifndef SB_COUNT_AGG
#define SB_COUNT_AGG 1 // warp-aggregated cell counts in the leaf kernel (build 0.548 -> 0.534 ms at C3)
#endif

// quantised triangle box: one word per world axis, lo | hi << 16; .w = triangle id
__device__ __forceinline__ uint4 quantise_box(const BoxD &b, const GridParams &g, uint32_t id, uint32_t job = 0)
{
    return make_uint4(quant_axis(b.lox, g, 0, job) | (quant_axis(b.hix, g, 0, job) << 16),
                      quant_axis(b.loy, g, 1, job) | (quant_axis(b.hiy, g, 1, job) << 16),
                      quant_axis(b.loz, g, 2, job) | (quant_axis(b.hiz, g, 2, job) << 16), id);
}

__device__ __forceinline__ uint32_t qbox_axis(const uint4 &q, int d) { return d == 0 ? q.x : d == 1 ? q.y : q.z; }

// footprint of a quantised box on grid a (rays along axis a; u, v = the other two axes)
struct GridFootprint {
    uint32_t cu0, cu1, cv0, cv1;
    uint32_t qu, qv, qa; // lo | hi << 16 along u, v, a
};

__device__ __forceinline__ GridFootprint grid_footprint(const uint4 &q, const GridParams &g, int a)
{
    GridFootprint f;
    const int u = a == 0 ? 1 : 0, v = a == 2 ? 1 : 2;
    f.qu = qbox_axis(q, u);
    f.qv = qbox_axis(q, v);
    f.qa = qbox_axis(q, a);
    f.cu0 = (f.qu & 0xffffu) >> g.shiftU[a];
    f.cu1 = (f.qu >> 16) >> g.shiftU[a];
    f.cv0 = (f.qv & 0xffffu) >> g.shiftV[a];
    f.cv1 = (f.qv >> 16) >> g.shiftV[a];
    return f;
}

// ---- depth slabs -------------------------------------------------------------------------
// A cell's list is ordered by the depth slab of each triangle's FAR bound along the ray axis
// (slab = the top slabBits[a] bits of the job-local 15- / 10-bit coordinate: monotone, so
// hi_a >= aA implies slab(hi_a) >= slab(aA)).  E holds one entry per (cell, slab), slabs
// ascending inside a cell; a ray that starts at depth aA reads the sub-lists of the slabs
// >= slab(aA) only: [E[e0 + slab + 1], E[e0 + (1 << slabBits) + 1] & ~1), a suffix of the
// cell's list.  References are not duplicated (the far bound picks ONE slab), so the fill
// costs what it did; the 12-bit depth test of the reference itself is unchanged.
__device__ __forceinline__ uint32_t grid_slab(uint32_t qa, const GridParams &g, int a)
{
    const int sb = (int)g.slabBits[a];
    const int sh = (g.latShift ? (int)g.latShift : SB_Q_BITS) - sb;
    return (qa >> sh) & ((1u << sb) - 1u);
}

// first entry (slab 0) of cell (cu, cv) of grid a
__device__ __forceinline__ uint32_t grid_entry0(const GridParams &g, int a, uint32_t cu, uint32_t cv)
{
    return g.cellBase[a] + ((cv * g.nu[a] + cu) << g.slabBits[a]);
}

// the references a ray in cell (cu, cv) of grid a has to look at when it starts at quantised depth aA:
// [i0, i1), i1 even; i0 may be odd (the slot before it then belongs to a nearer slab or is padding)
__device__ __forceinline__ void grid_ray_range(const GridParams &g, const uint32_t *__restrict__ E, int a, uint32_t cu, uint32_t cv,
    uint32_t aA, uint32_t &i0, uint32_t &i1)
{
    const uint32_t e0 = grid_entry0(g, a, cu, cv);
    i0 = __ldg(E + e0 + grid_slab(aA, g, a) + 1);
    i1 = __ldg(E + e0 + (1u << g.slabBits[a]) + 1) & ~1u; // the next cell's share may begin with an unused slot
}

// count pass for one triangle: +1 on every (cell, slab of its far bound) it covers on the first `naxes`
// grids (E[e + 1] = number of references of entry e), or on the axis's big-list counter
__device__ __forceinline__ void grid_count_tri(const uint4 &q, const GridParams &g, uint32_t *__restrict__ E,
    uint32_t *__restrict__ bigCount, int naxes)
{
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        if (a >= naxes)
            break;
        const GridFootprint f = grid_footprint(q, g, a);
        if ((f.cu1 - f.cu0 + 1) * (f.cv1 - f.cv0 + 1) > SB_GRID_MAX_CELLS_PER_TRI) {
            atomicAdd(&bigCount[a], 1u);
            continue;
        }
        const uint32_t nu = g.nu[a];
        const int sb = (int)g.slabBits[a];
        const uint32_t base = g.cellBase[a] + grid_slab(f.qa >> 16, g, a) + 1; // + (cell << sb)
#if SB_COUNT_AGG
        if (f.cu1 - f.cu0 <= 1 && f.cv1 - f.cv0 <= 1) {
            // common case: one atomic per distinct entry of the warp's lanes that took this path
            const unsigned act = __activemask();
            const uint32_t lane = threadIdx.x & 31;
            const bool du = f.cu1 != f.cu0, dv = f.cv1 != f.cv0;
            const uint32_t c00 = f.cv0 * nu + f.cu0;
            auto bump = [&](bool want, uint32_t cell) {
                const uint32_t e = base + (cell << sb);
                const unsigned peers = __match_any_sync(act, want ? e : 0xffffffffu - lane);
                if (want && (int)lane == __ffs(peers) - 1)
                    atomicAdd(&E[e], (uint32_t)__popc(peers));
            };
            bump(true, c00);
            bump(du, c00 + 1);
            bump(dv, c00 + nu);
            bump(du && dv, c00 + nu + 1);
            continue;
        }
#endif
        for (uint32_t cv = f.cv0; cv <= f.cv1; ++cv)
            for (uint32_t cu = f.cu0; cu <= f.cu1; ++cu)
                atomicAdd(&E[base + ((cv * nu + cu) << sb)], 1u);
    }
}
