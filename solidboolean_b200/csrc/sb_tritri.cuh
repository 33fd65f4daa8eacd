// Guigue-Devillers triangle/triangle intersection on the device, bit-exact with
// the reference's thirdparty/GuigueDevillers03/tri_tri_intersect.c.
//
// The reference expands the vertex permutations as nested macros (72 copies of
// CONSTRUCT_INTERSECTION).  Here the two canonicalisation steps are computed as
// DATA: sign_case() turns three plane distances into (rotation, swap, coplanar),
// the vertices are selected accordingly, and a single copy of the segment
// construction runs -- every lane of a warp executes the same instructions on
// differently permuted operands, so the kernel stays convergent.
//
// Every arithmetic step is one explicitly rounded binary64 operation in the
// reference's evaluation order (see sb_common.cuh).
#pragma once
#include "sb_fp64.cuh"

struct d2 {
    double x, y;
};

// ORIENT_2D, tri_tri_intersect.c:483
__device__ __forceinline__ double orient2d(const d2 &a, const d2 &b, const d2 &c)
{
    return xsub(xmul(xsub(a.x, c.x), xsub(b.y, c.y)), xmul(xsub(a.y, c.y), xsub(b.x, c.x)));
}

// INTERSECTION_TEST_VERTEX, tri_tri_intersect.c:487-513
__device__ __noinline__ int tt_vertex_2d(d2 P1, d2 Q1, d2 R1, d2 P2, d2 Q2, d2 R2)
{
    if (orient2d(R2, P2, Q1) >= 0.0) {
        if (orient2d(R2, Q2, Q1) <= 0.0) {
            if (orient2d(P1, P2, Q1) > 0.0)
                return orient2d(P1, Q2, Q1) <= 0.0 ? 1 : 0;
            if (orient2d(P1, P2, R1) >= 0.0)
                return orient2d(Q1, R1, P2) >= 0.0 ? 1 : 0;
            return 0;
        }
        if (orient2d(P1, Q2, Q1) <= 0.0) {
            if (orient2d(R2, Q2, R1) <= 0.0)
                return orient2d(Q1, R1, Q2) >= 0.0 ? 1 : 0;
            return 0;
        }
        return 0;
    }
    if (orient2d(R2, P2, R1) >= 0.0) {
        if (orient2d(Q1, R1, R2) >= 0.0)
            return orient2d(P1, P2, R1) >= 0.0 ? 1 : 0;
        if (orient2d(Q1, R1, Q2) >= 0.0)
            return orient2d(R2, R1, Q2) >= 0.0 ? 1 : 0;
        return 0;
    }
    return 0;
}

// INTERSECTION_TEST_EDGE, tri_tri_intersect.c:517-533
__device__ __noinline__ int tt_edge_2d(d2 P1, d2 Q1, d2 R1, d2 P2, d2 R2)
{
    if (orient2d(R2, P2, Q1) >= 0.0) {
        if (orient2d(P1, P2, Q1) >= 0.0)
            return orient2d(P1, Q1, R2) >= 0.0 ? 1 : 0;
        if (orient2d(Q1, R1, P2) >= 0.0)
            return orient2d(R1, P1, P2) >= 0.0 ? 1 : 0;
        return 0;
    }
    if (orient2d(R2, P2, R1) >= 0.0) {
        if (orient2d(P1, P2, R1) >= 0.0) {
            if (orient2d(P1, R1, R2) >= 0.0)
                return 1;
            return orient2d(Q1, R1, R2) >= 0.0 ? 1 : 0;
        }
        return 0;
    }
    return 0;
}

// ccw_tri_tri_intersection_2d, tri_tri_intersect.c:537-555
__device__ __noinline__ int tt_ccw_2d(d2 p1, d2 q1, d2 r1, d2 p2, d2 q2, d2 r2)
{
    if (orient2d(p2, q2, p1) >= 0.0) {
        if (orient2d(q2, r2, p1) >= 0.0) {
            if (orient2d(r2, p2, p1) >= 0.0)
                return 1;
            return tt_edge_2d(p1, q1, r1, p2, r2);
        }
        if (orient2d(r2, p2, p1) >= 0.0)
            return tt_edge_2d(p1, q1, r1, r2, q2);
        return tt_vertex_2d(p1, q1, r1, p2, q2, r2);
    }
    if (orient2d(q2, r2, p1) >= 0.0) {
        if (orient2d(r2, p2, p1) >= 0.0)
            return tt_edge_2d(p1, q1, r1, q2, p2);
        return tt_vertex_2d(p1, q1, r1, q2, r2, p2);
    }
    return tt_vertex_2d(p1, q1, r1, r2, p2, q2);
}

// tri_tri_overlap_test_2d, tri_tri_intersect.c:558-573
__device__ __forceinline__ int tt_overlap_2d(d2 p1, d2 q1, d2 r1, d2 p2, d2 q2, d2 r2)
{
    bool f1 = orient2d(p1, q1, r1) < 0.0;
    bool f2 = orient2d(p2, q2, r2) < 0.0;
    return tt_ccw_2d(p1, f1 ? r1 : q1, f1 ? q1 : r1, p2, f2 ? r2 : q2, f2 ? q2 : r2);
}

// coplanar_tri_tri3d, tri_tri_intersect.c:215-269 (note the p/q swap of the YZ
// and XZ projections).
__device__ __noinline__ int tt_coplanar(const d3 &p1, const d3 &q1, const d3 &r1,
    const d3 &p2, const d3 &q2, const d3 &r2, const d3 &n1)
{
    double nx = n1.x < 0 ? -n1.x : n1.x;
    double ny = n1.y < 0 ? -n1.y : n1.y;
    double nz = n1.z < 0 ? -n1.z : n1.z;
    if (nx > nz && nx >= ny)
        return tt_overlap_2d({q1.z, q1.y}, {p1.z, p1.y}, {r1.z, r1.y}, {q2.z, q2.y}, {p2.z, p2.y}, {r2.z, r2.y});
    if (ny > nz && ny >= nx)
        return tt_overlap_2d({q1.x, q1.z}, {p1.x, p1.z}, {r1.x, r1.z}, {q2.x, q2.z}, {p2.x, p2.z}, {r2.x, r2.z});
    return tt_overlap_2d({p1.x, p1.y}, {q1.x, q1.y}, {r1.x, r1.y}, {p2.x, p2.y}, {q2.x, q2.y}, {r2.x, r2.y});
}

// Decision table shared by both canonicalisation steps
// (tri_tri_intersect.c:443-471 for T1, :360-385 for T2).
//   rot 0: (p,q,r)   rot 1: (r,p,q)   rot 2: (q,r,p)
//   swap: the OTHER triangle's q and r (and their distances) are exchanged.
struct SignCase {
    int rot;
    bool swap;
    bool coplanar;
};

__device__ __forceinline__ SignCase sign_case(double dp, double dq, double dr)
{
    SignCase c = {0, false, false};
    if (dp > 0.0) {
        if (dq > 0.0) { c.rot = 1; c.swap = true; }
        else if (dr > 0.0) { c.rot = 2; c.swap = true; }
        else { c.rot = 0; c.swap = false; }
    } else if (dp < 0.0) {
        if (dq < 0.0) { c.rot = 1; c.swap = false; }
        else if (dr < 0.0) { c.rot = 2; c.swap = false; }
        else { c.rot = 0; c.swap = true; }
    } else {
        if (dq < 0.0) {
            if (dr >= 0.0) { c.rot = 2; c.swap = true; }
            else { c.rot = 0; c.swap = false; }
        } else if (dq > 0.0) {
            if (dr > 0.0) { c.rot = 0; c.swap = true; }
            else { c.rot = 2; c.swap = false; }
        } else {
            if (dr > 0.0) { c.rot = 1; c.swap = false; }
            else if (dr < 0.0) { c.rot = 1; c.swap = true; }
            else c.coplanar = true;
        }
    }
    return c;
}

__device__ __forceinline__ d3 sel3(int rot, const d3 &a, const d3 &b, const d3 &c)
{
    return rot == 0 ? a : (rot == 1 ? b : c);
}

// base - (num.n / den.n) * den   (tri_tri_intersect.c:296-303 and siblings)
__device__ __forceinline__ d3 edge_plane_point(const d3 &base, const d3 &num, const d3 &den, const d3 &n)
{
    double alpha = xdiv(d3dot(num, n), d3dot(den, n));
    d3 s = {xmul(alpha, den.x), xmul(alpha, den.y), xmul(alpha, den.z)};
    return d3sub(base, s);
}

// CONSTRUCT_INTERSECTION, tri_tri_intersect.c:285-356.  N1, N2: normals of the
// caller's UNPERMUTED triangles.
// `which` (0..3) names the branch taken, i.e. which edges the two end points lie on:
//   0: source on T1's edge p1-r1, target on T2's edge p2-r2      1: source on T2's p2-q2, target on T2's p2-r2
//   2: source on T1's p1-r1,      target on T1's p1-q1           3: source on T2's p2-q2, target on T1's p1-q1
// (carried out of the predicate so that retriangulation need not find the edge again with a tolerance)
__device__ __forceinline__ int tt_construct(const d3 &p1, const d3 &q1, const d3 &r1,
    const d3 &p2, const d3 &q2, const d3 &r2, const d3 &N1, const d3 &N2, d3 &source, d3 &target, int &which)
{
    d3 v1 = d3sub(q1, p1);
    d3 v2 = d3sub(r2, p1);
    d3 N = d3cross(v1, v2);
    d3 v = d3sub(p2, p1);
    if (d3dot(v, N) > 0.0) {
        v1 = d3sub(r1, p1);
        N = d3cross(v1, v2);
        if (d3dot(v, N) <= 0.0) {
            v2 = d3sub(q2, p1);
            N = d3cross(v1, v2);
            if (d3dot(v, N) > 0.0) {
                source = edge_plane_point(p1, d3sub(p1, p2), d3sub(p1, r1), N2);
                target = edge_plane_point(p2, d3sub(p2, p1), d3sub(p2, r2), N1);
                which = 0;
                return 1;
            }
            source = edge_plane_point(p2, d3sub(p2, p1), d3sub(p2, q2), N1);
            target = edge_plane_point(p2, d3sub(p2, p1), d3sub(p2, r2), N1);
            which = 1;
            return 1;
        }
        return 0;
    }
    v2 = d3sub(q2, p1);
    N = d3cross(v1, v2);
    if (d3dot(v, N) < 0.0)
        return 0;
    v1 = d3sub(r1, p1);
    N = d3cross(v1, v2);
    if (d3dot(v, N) >= 0.0) {
        source = edge_plane_point(p1, d3sub(p1, p2), d3sub(p1, r1), N2);
        target = edge_plane_point(p1, d3sub(p1, p2), d3sub(p1, q1), N2);
        which = 2;
        return 1;
    }
    source = edge_plane_point(p2, d3sub(p2, p1), d3sub(p2, q2), N1);
    target = edge_plane_point(p1, d3sub(p1, p2), d3sub(p1, q1), N2);
    which = 3;
    return 1;
}

// Edge k of a triangle joins its vertices k and (k + 1) mod 3; the edge between vertices i != j:
__device__ __forceinline__ int tt_edge_of(int i, int j) { return i + j == 1 ? 0 : (i + j == 3 ? 1 : 2); }

// Where the end points of a constructed segment lie, in terms of the caller's UNPERMUTED triangles:
//   bits 0-1: edge of the source point (0..2), bit 2: on T2 (else T1); bits 4-5 / 6: the same for the target; bit 7: set.
#define SB_SEG_TAG_VALID 0x80u
__device__ __forceinline__ unsigned tt_segment_tag(int which, int rot1, bool swap1, int rot2, bool swap2)
{
    // the permutations of tri_tri_intersection below, applied to vertex NUMBERS instead of coordinates
    int i1a = rot1 == 0 ? 0 : (rot1 == 1 ? 2 : 1), i1b = rot1 == 0 ? 1 : (rot1 == 1 ? 0 : 2), i1c = rot1 == 0 ? 2 : (rot1 == 1 ? 1 : 0);
    if (swap2) {
        int t = i1b; i1b = i1c; i1c = t;
    }
    const int jb = swap1 ? 2 : 1, jc = swap1 ? 1 : 2;
    const int i2a = rot2 == 0 ? 0 : (rot2 == 1 ? jc : jb), i2b = rot2 == 0 ? jb : (rot2 == 1 ? 0 : jc), i2c = rot2 == 0 ? jc : (rot2 == 1 ? jb : 0);
    const unsigned e1r = (unsigned)tt_edge_of(i1a, i1c), e1q = (unsigned)tt_edge_of(i1a, i1b);
    const unsigned e2r = (unsigned)tt_edge_of(i2a, i2c) | 4u, e2q = (unsigned)tt_edge_of(i2a, i2b) | 4u;
    const unsigned src = (which == 0 || which == 2) ? e1r : e2q;
    const unsigned tgt = which == 0 ? e2r : (which == 1 ? e2r : e1q);
    return src | (tgt << 4) | SB_SEG_TAG_VALID;
}

// Which exit a pair took (flop accounting of SURVEY 8d: 41 / 82 / 139 / 185 flops)
enum TriTriPath {
    TT_REJECT_PLANE2 = 0, // T1 entirely on one side of plane(T2)   (:421)
    TT_REJECT_PLANE1 = 1, // T2 entirely on one side of plane(T1)   (:438)
    TT_COPLANAR = 2,      // 2-D overlap test
    TT_REJECT_INTERVAL = 3, // intervals on the common line do not overlap
    TT_SEGMENT = 4        // intersection segment constructed
};

// tri_tri_intersection_test_3d, tri_tri_intersect.c:395-472.
// coplanar is only ever set to 1; source/target only written with a segment.
__device__ __forceinline__ int tri_tri_intersection(const d3 &p1, const d3 &q1, const d3 &r1,
    const d3 &p2, const d3 &q2, const d3 &r2, int &coplanar, d3 &source, d3 &target, int &path, unsigned *tag = nullptr)
{
    // signs of T1's vertices against plane(T2)  (:407-421)
    d3 N2 = d3cross(d3sub(p2, r2), d3sub(q2, r2));
    double dp1 = d3dot(d3sub(p1, r2), N2);
    double dq1 = d3dot(d3sub(q1, r2), N2);
    double dr1 = d3dot(d3sub(r1, r2), N2);
    path = TT_REJECT_PLANE2;
    if (xmul(dp1, dq1) > 0.0 && xmul(dp1, dr1) > 0.0)
        return 0;
    // signs of T2's vertices against plane(T1)  (:424-438)
    d3 N1 = d3cross(d3sub(q1, p1), d3sub(r1, p1));
    double dp2 = d3dot(d3sub(p2, r1), N1);
    double dq2 = d3dot(d3sub(q2, r1), N1);
    double dr2 = d3dot(d3sub(r2, r1), N1);
    path = TT_REJECT_PLANE1;
    if (xmul(dp2, dq2) > 0.0 && xmul(dp2, dr2) > 0.0)
        return 0;
    path = TT_COPLANAR;

    // canonical form of T1 (:443-471)
    SignCase c1 = sign_case(dp1, dq1, dr1);
    if (c1.coplanar) {
        coplanar = 1;
        return tt_coplanar(p1, q1, r1, p2, q2, r2, N1);
    }
    d3 a1 = sel3(c1.rot, p1, r1, q1);
    d3 b1 = sel3(c1.rot, q1, p1, r1);
    d3 c1v = sel3(c1.rot, r1, q1, p1);
    d3 b2 = c1.swap ? r2 : q2;
    d3 c2v = c1.swap ? q2 : r2;
    double eq2 = c1.swap ? dr2 : dq2;
    double er2 = c1.swap ? dq2 : dr2;

    // canonical form of T2 (:360-385), on the already permuted operands
    SignCase c2 = sign_case(dp2, eq2, er2);
    if (c2.coplanar) {
        coplanar = 1;
        return tt_coplanar(a1, b1, c1v, p2, b2, c2v, N1);
    }
    d3 a2 = sel3(c2.rot, p2, c2v, b2);
    d3 bb2 = sel3(c2.rot, b2, p2, c2v);
    d3 cc2 = sel3(c2.rot, c2v, b2, p2);
    d3 bb1 = c2.swap ? c1v : b1;
    d3 cc1 = c2.swap ? b1 : c1v;
    int which = 0;
    int r = tt_construct(a1, bb1, cc1, a2, bb2, cc2, N1, N2, source, target, which);
    path = r ? TT_SEGMENT : TT_REJECT_INTERVAL;
    if (tag)
        *tag = r ? tt_segment_tag(which, c1.rot, c1.swap, c2.rot, c2.swap) : 0u;
    return r;
}

__device__ __forceinline__ int tri_tri_intersection(const d3 &p1, const d3 &q1, const d3 &r1,
    const d3 &p2, const d3 &q2, const d3 &r2, int &coplanar, d3 &source, d3 &target)
{
    int path = 0;
    return tri_tri_intersection(p1, q1, r1, p2, q2, r2, coplanar, source, target, path);
}
