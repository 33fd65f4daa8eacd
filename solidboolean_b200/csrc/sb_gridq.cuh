// The quantiser shared by the ray-grid build (sb_grid.cu) and the classifier
// (sb_classify.cu), and the 16-byte grid reference built from it.
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

__device__ __forceinline__ uint32_t quant15(double x, double org, double scl)
{
    double t = floor((x - org) * scl);
    t = fmin(fmax(t, 0.0), 32767.0); // NaN -> 0
    return (uint32_t)t;
}

// reference of a triangle whose quantised box is [qlu,qhu] x [qlv,qhv] across the ray
// axis and [qla,qha] along it
__device__ __forceinline__ uint4 grid_ref_pack(uint32_t qlu, uint32_t qhu, uint32_t qlv, uint32_t qhv, uint32_t qla,
    uint32_t qha, uint32_t id)
{
    return make_uint4(qlu | ((SB_Q_MAX - qhu) << 16), qlv | ((SB_Q_MAX - qhv) << 16), (SB_Q_MAX - qha) | (qla << 16), id);
}

__device__ __forceinline__ uint32_t grid_ref_lo_u(const uint4 &r) { return r.x & SB_Q_MAX; }
__device__ __forceinline__ uint32_t grid_ref_lo_v(const uint4 &r) { return r.y & SB_Q_MAX; }

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

// ---- binning (shared by the leaf kernel of sb_build.cu, which counts, and sb_grid.cu, which fills) ----
#define SB_GRID_MAX_CELLS_PER_TRI 1024u // larger footprints go to the per-axis "big" list

// quantised triangle box: one word per world axis, lo | hi << 16; .w = triangle id
__device__ __forceinline__ uint4 quantise_box(const BoxD &b, const GridParams &g, uint32_t id)
{
    return make_uint4(quant15(b.lox, g.org[0], g.scl[0]) | (quant15(b.hix, g.org[0], g.scl[0]) << 16),
                      quant15(b.loy, g.org[1], g.scl[1]) | (quant15(b.hiy, g.org[1], g.scl[1]) << 16),
                      quant15(b.loz, g.org[2], g.scl[2]) | (quant15(b.hiz, g.org[2], g.scl[2]) << 16), id);
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

// count pass for one triangle: +1 on every cell it covers on the first `naxes` grids
// (E[c + 1] = number of references of cell c), or on the axis's big-list counter
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
        const uint32_t base = g.cellBase[a], nu = g.nu[a];
        for (uint32_t cv = f.cv0; cv <= f.cv1; ++cv)
            for (uint32_t cu = f.cu0; cu <= f.cu1; ++cu)
                atomicAdd(&E[base + cv * nu + cu + 1], 1u);
    }
}
