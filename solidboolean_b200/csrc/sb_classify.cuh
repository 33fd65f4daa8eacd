// Device helpers shared by the two classification kernels (sb_classify.cu: the general
// warp-per-32-points kernel; sb_classify2.cu: the balanced single-walk kernel): query /
// target / output descriptors, the ray box and its quantised form, and the evaluation of
// one (ray, triangle) entry exactly as SolidBoolean::isPointInMesh does it
// (reference src/solidboolean.cpp:48-92).
#pragma once
#include "sb_internal.h"
#include "sb_gridq.cuh"
#include "sb_raytri.cuh"

#ifndef SB_CLS_RBOX
#define SB_CLS_RBOX 1   // ray box shortcut for finite points
#endif

namespace {

#define SB_BOX_UPD(b, v)                          \
    if (v.x > b.hix) b.hix = v.x;                 \
    if (v.x < b.lox) b.lox = v.x;                 \
    if (v.y > b.hiy) b.hiy = v.y;                 \
    if (v.y < b.loy) b.loy = v.y;                 \
    if (v.z > b.hiz) b.hiz = v.z;                 \
    if (v.z < b.loz) b.loz = v.z;

__device__ __forceinline__ BoxD ray_box(const d3 &p, const d3 &e)
{
    // box.update(testPosition); box.update(testEnd)  (src/solidboolean.cpp:55-58).
    // For a point with |coordinates| < DBL_MAX the updates leave lo = p and hi = e: the
    // first one stores p in both (-DBL_MAX < p < DBL_MAX), and e = p + (DBL_MAX or
    // DBL_EPSILON) >= p (rounding is monotone; +inf compares greater) only ever raises hi.
#if SB_CLS_RBOX
    if (fabs(p.x) < DBL_MAX && fabs(p.y) < DBL_MAX && fabs(p.z) < DBL_MAX)
        return {p.x, p.y, p.z, e.x, e.y, e.z};
#endif
    BoxD b = {DBL_MAX, DBL_MAX, DBL_MAX, -DBL_MAX, -DBL_MAX, -DBL_MAX};
    SB_BOX_UPD(b, p) SB_BOX_UPD(b, e)
    return b;
}

__device__ __forceinline__ double comp(const BoxD &b, int d, bool hi)
{
    return d == 0 ? (hi ? b.hix : b.lox) : d == 1 ? (hi ? b.hiy : b.loy) : (hi ? b.hiz : b.loz);
}

struct Target {
    const GridParams *gp;
    const uint32_t *E;
    const uint2 *refs;      // cell lists: 8-byte cell-relative references (sb_gridq.cuh)
    const uint4 *bigRefs;   // per-axis big lists: 16-byte absolute references
    uint32_t bigCap;
    uint32_t bigN0, bigN1, bigN2;
    int naxes;                 // ray grids the target has (2: the third is built on demand, see ensure_grid3)
    const double4 *vtx;
    const uint32_t *tri;
    const double4 *nrm4;       // per triangle: unit normal + packed vertex indices (sb_common.cuh)
    // A front end enqueued BEFORE the host has checked that a rebuild's references fitted the lists sized for the old
    // geometry (sb_capi.cu, optimistic enqueue) may walk lists whose tail was never written: reads stay inside the
    // allocation (refPairs) and triangle numbers inside the mesh (nT); such a run's results are thrown away
    uint32_t refPairs;         // readable 16-byte pairs behind refs
    uint32_t nT;
};

struct Query {
    const double *pts;         // explicit points (AoS), or null
    const double *scent;       // faces mode: query mesh centroids in Morton order ...
    const uint32_t *sortedTri; // ... and sorted position -> triangle id
    uint32_t nT;
    uint32_t begin;            // first point / sorted position
    uint32_t count;            // points in this launch
    const uint32_t *list;      // optional: the launch's points as indices relative to `begin`
    const uint16_t *triJob;    // batch meshes (faces mode): job of each query face, or null
    const uint32_t *origFace;  // multi-GPU selection (faces mode): query face -> the parent's triangle id (output index), or null
    int ownMode;               // ... and only the faces with centroid z in [ownLo, ownHi) (2: <= ownHi) are classified
    double ownLo, ownHi;
    // raw faces mode (rawTri != null): the query faces in their ORIGINAL order straight from the uploaded arrays --
    // the query mesh need not be built (a mesh whose bytes have only just arrived, sb_capi.cu front end): point j is
    // face begin + j, j runs from `first` (the launch covers one upload chunk: faces [first, count))
    const uint32_t *rawTri;
    const double *rawXyz;
    uint32_t rawNV;
    uint32_t first;
};

// centroid of face f of a raw query, ((v0 + v1) + v2) / 3.0 (src/solidboolean.cpp:497-499) -- the same operations on
// the same doubles as the leaf kernel of sb_build.cu, which reads the padded copies
__device__ __forceinline__ d3 raw_face_centroid(const Query &q, uint32_t f)
{
    uint32_t i0 = __ldg(q.rawTri + 3 * (size_t)f), i1 = __ldg(q.rawTri + 3 * (size_t)f + 1), i2 = __ldg(q.rawTri + 3 * (size_t)f + 2);
    if (i0 >= q.rawNV || i1 >= q.rawNV || i2 >= q.rawNV) // (the build reports it)
        i0 = i1 = i2 = 0;
    const double *x = q.rawXyz;
    const d3 a = {__ldg(x + 3 * (size_t)i0), __ldg(x + 3 * (size_t)i0 + 1), __ldg(x + 3 * (size_t)i0 + 2)};
    const d3 b = {__ldg(x + 3 * (size_t)i1), __ldg(x + 3 * (size_t)i1 + 1), __ldg(x + 3 * (size_t)i1 + 2)};
    const d3 c = {__ldg(x + 3 * (size_t)i2), __ldg(x + 3 * (size_t)i2 + 1), __ldg(x + 3 * (size_t)i2 + 2)};
    return {xdiv(xadd(xadd(a.x, b.x), c.x), 3.0), xdiv(xadd(xadd(a.y, b.y), c.y), 3.0), xdiv(xadd(xadd(a.z, b.z), c.z), 3.0)};
}

struct Out {
    uint8_t *inside;           // indexed by point index / original triangle id
    uint8_t *perAxis;          // optional, 3 per point: all three rays are traced
    long long *bigKeys;        // scratch of the many-layer rays, 3 x bigCap
    unsigned long long bigCap;
    unsigned long long *bigNeeded;  // entries the many-layer rays asked for (> bigCap: repeat)
    unsigned long long *exactCount; // true candidates (exact box overlap), roofline accounting
    unsigned int *undecidedCount;   // points whose first two votes disagreed
    uint32_t *undecidedList;   // ... listed here (relative to `begin`) when the target has no third grid yet
    int thirdOnly;             // second launch: only the third ray, of the listed points; its vote decides
    uint32_t poolLimit;        // rays with more matches go through big_ray (<= POOL; smaller only in tests)
    unsigned long long *trace; // dev: per CTA {sm id, start ns, end ns, entries evaluated} (SB_CLASSIFY_TRACE), or null
    unsigned int *legacyCount; // sb_classify2.cu: points left to the general kernel ...
    uint32_t *legacyList;      // ... listed here (relative to `begin`)
};

struct RaySetup {
    uint32_t aU, bU, aV, bV, aA; // the ray box, 15-bit quantised: [aU,bU] x [aV,bV] across, from aA along the axis
    uint32_t cu0, cu1, cv0, cv1; // the cells it touches
    bool any;                    // the ray box overlaps the mesh box
};

__device__ __forceinline__ RaySetup ray_setup(const GridParams &g, int axis, const d3 &p, uint32_t job = 0)
{
    RaySetup rs;
    const d3 e = ray_end(p, axis);
    const BoxD myD = ray_box(p, e);
    const BoxD meshBox = {g.lo[0], g.lo[1], g.lo[2], g.hi[0], g.hi[1], g.hi[2]};
    rs.any = overlap_d(meshBox, myD); // otherwise no triangle box can overlap the ray box
    const int u = axis == 0 ? 1 : 0, v = axis == 2 ? 1 : 2;
    rs.aU = quant_axis(comp(myD, u, false), g, u, job);
    rs.bU = quant_axis(comp(myD, u, true), g, u, job);
    rs.aV = quant_axis(comp(myD, v, false), g, v, job);
    rs.bV = quant_axis(comp(myD, v, true), g, v, job);
    rs.aA = quant_axis(comp(myD, axis, false), g, axis, job);
    const int su = g.shiftU[axis], sv = g.shiftV[axis];
    rs.cu0 = rs.aU >> su; rs.cu1 = rs.bU >> su;
    rs.cv0 = rs.aV >> sv; rs.cv1 = rs.bV >> sv;
    return rs;
}

// The same for a point with finite coordinates and a compile-time axis (sb_classify2.cu): the ray
// box is lo = p, hi = p + axis vector (see ray_box), nothing else is formed.
template <int AXIS>
__device__ __forceinline__ RaySetup ray_setup_finite(const GridParams &g, const d3 &p, uint32_t job)
{
    constexpr int u = AXIS == 0 ? 1 : 0, v = AXIS == 2 ? 1 : 2;
    const double pu = u == 0 ? p.x : p.y, pv = v == 1 ? p.y : p.z, pa = AXIS == 0 ? p.x : AXIS == 1 ? p.y : p.z;
    const double eu = xadd(pu, SB_DBL_EPSILON), ev = xadd(pv, SB_DBL_EPSILON), ea = xadd(pa, SB_DBL_MAX);
    RaySetup rs;
    rs.any = (g.lo[u] <= eu) & (g.hi[u] >= pu) & (g.lo[v] <= ev) & (g.hi[v] >= pv) & (g.lo[AXIS] <= ea) & (g.hi[AXIS] >= pa);
    rs.aU = quant_axis(pu, g, u, job);
    rs.bU = quant_axis(eu, g, u, job);
    rs.aV = quant_axis(pv, g, v, job);
    rs.bV = quant_axis(ev, g, v, job);
    rs.aA = quant_axis(pa, g, AXIS, job);
    const int su = g.shiftU[AXIS], sv = g.shiftV[AXIS];
    rs.cu0 = rs.aU >> su; rs.cu1 = rs.bU >> su;
    rs.cv0 = rs.aV >> sv; rs.cv1 = rs.bV >> sv;
    return rs;
}

// the ray as seen from cell (cu, cv): its box clipped to the cell, cell-relative
__device__ __forceinline__ CellRay ray_in_cell(const GridParams &g, int axis, const RaySetup &rs, uint32_t cu, uint32_t cv)
{
    const int su = g.shiftU[axis], sv = g.shiftV[axis];
    const uint32_t u0 = cu << su, u1 = u0 + (1u << su) - 1u, v0 = cv << sv, v1 = v0 + (1u << sv) - 1u;
    return cell_ray_pack(max(rs.aU, u0), min(rs.bU, u1), max(rs.aV, v0), min(rs.bV, v1), rs.aA, cu, cv, su, sv);
}

__device__ __forceinline__ uint32_t big_list_length(const Target &T, int axis)
{
    return axis == 0 ? T.bigN0 : axis == 1 ? T.bigN1 : T.bigN2;
}

// triangle box .intersectWith(ray box) (axisalignedboundingbox.h:95-105) without forming
// the triangle box: AxisAlignedBoudingBox::update (:31-41) leaves lo = min(DBL_MAX,
// {c : c < DBL_MAX}) and hi = max(-DBL_MAX, {c : c > -DBL_MAX}) over the vertex
// coordinates c (NaN never updates), so
//     lo <= X  <=>  DBL_MAX <= X || c0 <= X || c1 <= X || c2 <= X
//     hi >= Y  <=>  -DBL_MAX >= Y || c0 >= Y || c1 >= Y || c2 >= Y
// (a c >= DBL_MAX that satisfies c <= X implies DBL_MAX <= X; mirrored for hi).
// 24 predicate-setting compares chained in PTX (setp.le.or / setp.ge.or), no branches.  Written
// in C the compiler turns every row into a NaN-proofed DSETP.MIN/MAX + select sequence (158
// instructions with the moves around them, a quarter of the evaluation step).
#ifndef SB_CLS_BOXASM
#define SB_CLS_BOXASM 1
#endif
__device__ __forceinline__ bool tri_box_overlaps(const d3 &t0, const d3 &t1, const d3 &t2, const BoxD &rb)
{
#if SB_CLS_BOXASM && defined(__CUDA_ARCH__)
    uint32_t r;
    asm("{\n\t"
        ".reg .pred a, b;\n\t"
        "setp.le.f64 a, %19, %4;\n\t"       // DBL_MAX <= hix
        "setp.le.or.f64 a, %1, %4, a;\n\t"
        "setp.le.or.f64 a, %7, %4, a;\n\t"
        "setp.le.or.f64 a, %13, %4, a;\n\t"
        "setp.ge.f64 b, %20, %5;\n\t"       // -DBL_MAX >= lox
        "setp.ge.or.f64 b, %1, %5, b;\n\t"
        "setp.ge.or.f64 b, %7, %5, b;\n\t"
        "setp.ge.or.f64 b, %13, %5, b;\n\t"
        "and.pred a, a, b;\n\t"
        "setp.le.f64 b, %19, %6;\n\t"       // y
        "setp.le.or.f64 b, %2, %6, b;\n\t"
        "setp.le.or.f64 b, %8, %6, b;\n\t"
        "setp.le.or.f64 b, %14, %6, b;\n\t"
        "and.pred a, a, b;\n\t"
        "setp.ge.f64 b, %20, %10;\n\t"
        "setp.ge.or.f64 b, %2, %10, b;\n\t"
        "setp.ge.or.f64 b, %8, %10, b;\n\t"
        "setp.ge.or.f64 b, %14, %10, b;\n\t"
        "and.pred a, a, b;\n\t"
        "setp.le.f64 b, %19, %11;\n\t"      // z
        "setp.le.or.f64 b, %3, %11, b;\n\t"
        "setp.le.or.f64 b, %9, %11, b;\n\t"
        "setp.le.or.f64 b, %15, %11, b;\n\t"
        "and.pred a, a, b;\n\t"
        "setp.ge.f64 b, %20, %12;\n\t"
        "setp.ge.or.f64 b, %3, %12, b;\n\t"
        "setp.ge.or.f64 b, %9, %12, b;\n\t"
        "setp.ge.or.f64 b, %15, %12, b;\n\t"
        "and.pred a, a, b;\n\t"
        "selp.u32 %0, 1, 0, a;\n\t"
        "}"
        : "=r"(r)
        : "d"(t0.x), "d"(t0.y), "d"(t0.z),           // 1 2 3
          "d"(rb.hix), "d"(rb.lox), "d"(rb.hiy),     // 4 5 6
          "d"(t1.x), "d"(t1.y), "d"(t1.z),           // 7 8 9
          "d"(rb.loy), "d"(rb.hiz), "d"(rb.loz),     // 10 11 12
          "d"(t2.x), "d"(t2.y), "d"(t2.z),           // 13 14 15
          "d"(0.0), "d"(0.0), "d"(0.0),              // 16 17 18 (unused)
          "d"(DBL_MAX), "d"(-DBL_MAX));              // 19 20
    return r != 0;
#else
    return ((DBL_MAX <= rb.hix) | (t0.x <= rb.hix) | (t1.x <= rb.hix) | (t2.x <= rb.hix)) &
           ((-DBL_MAX >= rb.lox) | (t0.x >= rb.lox) | (t1.x >= rb.lox) | (t2.x >= rb.lox)) &
           ((DBL_MAX <= rb.hiy) | (t0.y <= rb.hiy) | (t1.y <= rb.hiy) | (t2.y <= rb.hiy)) &
           ((-DBL_MAX >= rb.loy) | (t0.y >= rb.loy) | (t1.y >= rb.loy) | (t2.y >= rb.loy)) &
           ((DBL_MAX <= rb.hiz) | (t0.z <= rb.hiz) | (t1.z <= rb.hiz) | (t2.z <= rb.hiz)) &
           ((-DBL_MAX >= rb.loz) | (t0.z >= rb.loz) | (t1.z >= rb.loz) | (t2.z >= rb.loz));
#endif
}

// One (ray, triangle) entry: is it a candidate of the reference (triangle box
// .intersectWith(ray box), exact doubles, :55-63), and does the reference insert
// PositionKey(hit) for it (:66-87)?
// FINITE: the caller vouches for |coordinates of p| < DBL_MAX (ray box = {p, e}, see ray_box)
template <bool FINITE = false>
__device__ __forceinline__ bool eval_entry(const Target &T, const d3 &p, int axis, uint32_t f, long long &k0, long long &k1,
    long long &k2, bool &isCand)
{
#if SB_CLS_GUARDS
    if (f >= T.nT) { // (only in a run whose lists overflowed, see Target)
        isCand = false;
        return false;
    }
#endif
    // first round trip: the triangle's record = normal + packed vertex indices (one 256-bit gather) ...
    const double4 rec = ldg256(T.nrm4 + f);
    uint32_t i0, i1, i2;
    const unsigned long long w = (unsigned long long)__double_as_longlong(rec.w);
    if (w != SB_PACKED_IDX_NONE) {
        unpack_tri_idx(w, i0, i1, i2);
    } else { // (rare: corners too far apart in the vertex array, or more than 4 M vertices)
        i0 = __ldg(T.tri + 3 * (size_t)f);
        i1 = __ldg(T.tri + 3 * (size_t)f + 1);
        i2 = __ldg(T.tri + 3 * (size_t)f + 2);
    }
    // ... second: its three vertices (one 256-bit gather each)
    const d3 t0 = load_vertex(T.vtx, i0);
    const d3 t1 = load_vertex(T.vtx, i1);
    const d3 t2 = load_vertex(T.vtx, i2);
    const d3 nrm = {rec.x, rec.y, rec.z};
    const d3 e = ray_end(p, axis);
    isCand = tri_box_overlaps(t0, t1, t2, FINITE ? BoxD{p.x, p.y, p.z, e.x, e.y, e.z} : ray_box(p, e));
    if (!isCand)
        return false;
    d3 hit = {0, 0, 0};
    if (!ray_tri_hit_filtered(p, e, t0, t1, t2, nrm, hit))
        return false;
    k0 = position_key(hit.x);
    k1 = position_key(hit.y);
    k2 = position_key(hit.z);
    return true;
}

} // namespace
