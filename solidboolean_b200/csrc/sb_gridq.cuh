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
