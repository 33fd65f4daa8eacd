/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the solidboolean intersection
 * front end.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg may load this library, and only as the checker.  The product (CUDA) path
 * never links or calls it.
 *
 * This is a plain-C restatement of the reference algorithm for the hot path,
 * written against /root/reference (file:line cited per function).  Parity is
 * PINNED: tests/test_oracle_pinning.py checks every function here bit-for-bit
 * against the unmodified reference compiled into oracle/_ref/libsbref.so, on
 * the four bundled test/cases pairs, synthetic meshes and crafted degenerate
 * pairs, and against the committed golden fixtures in tests/golden/.
 *
 * Build: gcc -O2 -ffp-contract=off (see oracle/Makefile).  Every arithmetic
 * step below is one IEEE binary64 operation in the reference's evaluation
 * order; do not re-associate.
 */
#include <float.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------ */
/* small vector helpers: thirdparty/GuigueDevillers03/tri_tri_intersect.c:73-89 */

static inline void v_sub(double *d, const double *a, const double *b)
{
    d[0] = a[0] - b[0];
    d[1] = a[1] - b[1];
    d[2] = a[2] - b[2];
}

static inline void v_cross(double *d, const double *a, const double *b)
{
    d[0] = a[1] * b[2] - a[2] * b[1];
    d[1] = a[2] * b[0] - a[0] * b[2];
    d[2] = a[0] * b[1] - a[1] * b[0];
}

static inline double v_dot(const double *a, const double *b)
{
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}

/* point = base - (num.n / den.n) * den, the SCALAR/SUB tail used four times in
 * CONSTRUCT_INTERSECTION (tri_tri_intersect.c:296-303 and siblings). */
static inline void edge_plane_point(double *out, const double *base,
    const double *num, const double *den, const double *n)
{
    double alpha = v_dot(num, n) / v_dot(den, n);
    double s[3];
    s[0] = alpha * den[0];
    s[1] = alpha * den[1];
    s[2] = alpha * den[2];
    v_sub(out, base, s);
}

/* ------------------------------------------------------------------------ */
/* 2-D overlap test: tri_tri_intersect.c:478-573 */

static inline double orient2d(const double *a, const double *b, const double *c)
{
    return (a[0] - c[0]) * (b[1] - c[1]) - (a[1] - c[1]) * (b[0] - c[0]);
}

/* INTERSECTION_TEST_VERTEX, tri_tri_intersect.c:487-513 */
static int test_vertex_2d(const double *P1, const double *Q1, const double *R1,
    const double *P2, const double *Q2, const double *R2)
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

/* INTERSECTION_TEST_EDGE, tri_tri_intersect.c:517-533 */
static int test_edge_2d(const double *P1, const double *Q1, const double *R1,
    const double *P2, const double *Q2, const double *R2)
{
    (void)Q2;
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

/* ccw_tri_tri_intersection_2d, tri_tri_intersect.c:537-555 */
static int ccw_overlap_2d(const double *p1, const double *q1, const double *r1,
    const double *p2, const double *q2, const double *r2)
{
    if (orient2d(p2, q2, p1) >= 0.0) {
        if (orient2d(q2, r2, p1) >= 0.0) {
            if (orient2d(r2, p2, p1) >= 0.0)
                return 1;
            return test_edge_2d(p1, q1, r1, p2, q2, r2);
        }
        if (orient2d(r2, p2, p1) >= 0.0)
            return test_edge_2d(p1, q1, r1, r2, p2, q2);
        return test_vertex_2d(p1, q1, r1, p2, q2, r2);
    }
    if (orient2d(q2, r2, p1) >= 0.0) {
        if (orient2d(r2, p2, p1) >= 0.0)
            return test_edge_2d(p1, q1, r1, q2, r2, p2);
        return test_vertex_2d(p1, q1, r1, q2, r2, p2);
    }
    return test_vertex_2d(p1, q1, r1, r2, p2, q2);
}

/* tri_tri_overlap_test_2d, tri_tri_intersect.c:558-573 */
static int overlap_2d(const double *p1, const double *q1, const double *r1,
    const double *p2, const double *q2, const double *r2)
{
    if (orient2d(p1, q1, r1) < 0.0) {
        if (orient2d(p2, q2, r2) < 0.0)
            return ccw_overlap_2d(p1, r1, q1, p2, r2, q2);
        return ccw_overlap_2d(p1, r1, q1, p2, q2, r2);
    }
    if (orient2d(p2, q2, r2) < 0.0)
        return ccw_overlap_2d(p1, q1, r1, p2, r2, q2);
    return ccw_overlap_2d(p1, q1, r1, p2, q2, r2);
}

/* coplanar_tri_tri3d, tri_tri_intersect.c:215-269.  Projects on the plane that
 * maximises the area; note the p/q swap on the YZ and XZ projections. */
static int coplanar_3d(const double *p1, const double *q1, const double *r1,
    const double *p2, const double *q2, const double *r2, const double *n1)
{
    double P1[2], Q1[2], R1[2], P2[2], Q2[2], R2[2];
    double nx = n1[0] < 0 ? -n1[0] : n1[0];
    double ny = n1[1] < 0 ? -n1[1] : n1[1];
    double nz = n1[2] < 0 ? -n1[2] : n1[2];
    if (nx > nz && nx >= ny) {
        P1[0] = q1[2]; P1[1] = q1[1];
        Q1[0] = p1[2]; Q1[1] = p1[1];
        R1[0] = r1[2]; R1[1] = r1[1];
        P2[0] = q2[2]; P2[1] = q2[1];
        Q2[0] = p2[2]; Q2[1] = p2[1];
        R2[0] = r2[2]; R2[1] = r2[1];
    } else if (ny > nz && ny >= nx) {
        P1[0] = q1[0]; P1[1] = q1[2];
        Q1[0] = p1[0]; Q1[1] = p1[2];
        R1[0] = r1[0]; R1[1] = r1[2];
        P2[0] = q2[0]; P2[1] = q2[2];
        Q2[0] = p2[0]; Q2[1] = p2[2];
        R2[0] = r2[0]; R2[1] = r2[2];
    } else {
        P1[0] = p1[0]; P1[1] = p1[1];
        Q1[0] = q1[0]; Q1[1] = q1[1];
        R1[0] = r1[0]; R1[1] = r1[1];
        P2[0] = p2[0]; P2[1] = p2[1];
        Q2[0] = q2[0]; Q2[1] = q2[1];
        R2[0] = r2[0]; R2[1] = r2[1];
    }
    return overlap_2d(P1, Q1, R1, P2, Q2, R2);
}

/* ------------------------------------------------------------------------ */
/* CONSTRUCT_INTERSECTION, tri_tri_intersect.c:285-356.  N1/N2 are the plane
 * normals computed from the caller's UNPERMUTED triangles (:410-412, :427-429). */
static int construct_segment(const double *p1, const double *q1, const double *r1,
    const double *p2, const double *q2, const double *r2,
    const double *N1, const double *N2, double *source, double *target)
{
    double v1[3], v2[3], v[3], N[3];
    v_sub(v1, q1, p1);
    v_sub(v2, r2, p1);
    v_cross(N, v1, v2);
    v_sub(v, p2, p1);
    if (v_dot(v, N) > 0.0) {
        v_sub(v1, r1, p1);
        v_cross(N, v1, v2);
        if (v_dot(v, N) <= 0.0) {
            v_sub(v2, q2, p1);
            v_cross(N, v1, v2);
            if (v_dot(v, N) > 0.0) {
                v_sub(v1, p1, p2);
                v_sub(v2, p1, r1);
                edge_plane_point(source, p1, v1, v2, N2);
                v_sub(v1, p2, p1);
                v_sub(v2, p2, r2);
                edge_plane_point(target, p2, v1, v2, N1);
                return 1;
            }
            v_sub(v1, p2, p1);
            v_sub(v2, p2, q2);
            edge_plane_point(source, p2, v1, v2, N1);
            v_sub(v1, p2, p1);
            v_sub(v2, p2, r2);
            edge_plane_point(target, p2, v1, v2, N1);
            return 1;
        }
        return 0;
    }
    v_sub(v2, q2, p1);
    v_cross(N, v1, v2);
    if (v_dot(v, N) < 0.0)
        return 0;
    v_sub(v1, r1, p1);
    v_cross(N, v1, v2);
    if (v_dot(v, N) >= 0.0) {
        v_sub(v1, p1, p2);
        v_sub(v2, p1, r1);
        edge_plane_point(source, p1, v1, v2, N2);
        v_sub(v1, p1, p2);
        v_sub(v2, p1, q1);
        edge_plane_point(target, p1, v1, v2, N2);
        return 1;
    }
    v_sub(v1, p2, p1);
    v_sub(v2, p2, q2);
    edge_plane_point(source, p2, v1, v2, N1);
    v_sub(v1, p1, p2);
    v_sub(v2, p1, q1);
    edge_plane_point(target, p1, v1, v2, N2);
    return 1;
}

/* TRI_TRI_INTER_3D, tri_tri_intersect.c:360-385: canonical permutation of T2
 * from the signs of its vertices against plane(T1). */
static int permute_t2(const double *p1, const double *q1, const double *r1,
    const double *p2, const double *q2, const double *r2,
    double dp2, double dq2, double dr2,
    const double *N1, const double *N2, int *coplanar, double *source, double *target)
{
    if (dp2 > 0.0) {
        if (dq2 > 0.0)
            return construct_segment(p1, r1, q1, r2, p2, q2, N1, N2, source, target);
        if (dr2 > 0.0)
            return construct_segment(p1, r1, q1, q2, r2, p2, N1, N2, source, target);
        return construct_segment(p1, q1, r1, p2, q2, r2, N1, N2, source, target);
    }
    if (dp2 < 0.0) {
        if (dq2 < 0.0)
            return construct_segment(p1, q1, r1, r2, p2, q2, N1, N2, source, target);
        if (dr2 < 0.0)
            return construct_segment(p1, q1, r1, q2, r2, p2, N1, N2, source, target);
        return construct_segment(p1, r1, q1, p2, q2, r2, N1, N2, source, target);
    }
    if (dq2 < 0.0) {
        if (dr2 >= 0.0)
            return construct_segment(p1, r1, q1, q2, r2, p2, N1, N2, source, target);
        return construct_segment(p1, q1, r1, p2, q2, r2, N1, N2, source, target);
    }
    if (dq2 > 0.0) {
        if (dr2 > 0.0)
            return construct_segment(p1, r1, q1, p2, q2, r2, N1, N2, source, target);
        return construct_segment(p1, q1, r1, q2, r2, p2, N1, N2, source, target);
    }
    if (dr2 > 0.0)
        return construct_segment(p1, q1, r1, r2, p2, q2, N1, N2, source, target);
    if (dr2 < 0.0)
        return construct_segment(p1, r1, q1, r2, p2, q2, N1, N2, source, target);
    *coplanar = 1;
    return coplanar_3d(p1, q1, r1, p2, q2, r2, N1);
}

/* tri_tri_intersection_test_3d, tri_tri_intersect.c:395-472.
 * `coplanar` is only ever SET (the caller pre-zeroes it, solidboolean.cpp:107);
 * source/target are written only when a segment is constructed. */
int sbo_tri_tri(const double *p1, const double *q1, const double *r1,
    const double *p2, const double *q2, const double *r2,
    int *coplanar, double *source, double *target)
{
    double v1[3], v2[3], N1[3], N2[3];
    double dp1, dq1, dr1, dp2, dq2, dr2;

    v_sub(v1, p2, r2);
    v_sub(v2, q2, r2);
    v_cross(N2, v1, v2);
    v_sub(v1, p1, r2);
    dp1 = v_dot(v1, N2);
    v_sub(v1, q1, r2);
    dq1 = v_dot(v1, N2);
    v_sub(v1, r1, r2);
    dr1 = v_dot(v1, N2);
    if (dp1 * dq1 > 0.0 && dp1 * dr1 > 0.0)
        return 0;

    v_sub(v1, q1, p1);
    v_sub(v2, r1, p1);
    v_cross(N1, v1, v2);
    v_sub(v1, p2, r1);
    dp2 = v_dot(v1, N1);
    v_sub(v1, q2, r1);
    dq2 = v_dot(v1, N1);
    v_sub(v1, r2, r1);
    dr2 = v_dot(v1, N1);
    if (dp2 * dq2 > 0.0 && dp2 * dr2 > 0.0)
        return 0;

    if (dp1 > 0.0) {
        if (dq1 > 0.0)
            return permute_t2(r1, p1, q1, p2, r2, q2, dp2, dr2, dq2, N1, N2, coplanar, source, target);
        if (dr1 > 0.0)
            return permute_t2(q1, r1, p1, p2, r2, q2, dp2, dr2, dq2, N1, N2, coplanar, source, target);
        return permute_t2(p1, q1, r1, p2, q2, r2, dp2, dq2, dr2, N1, N2, coplanar, source, target);
    }
    if (dp1 < 0.0) {
        if (dq1 < 0.0)
            return permute_t2(r1, p1, q1, p2, q2, r2, dp2, dq2, dr2, N1, N2, coplanar, source, target);
        if (dr1 < 0.0)
            return permute_t2(q1, r1, p1, p2, q2, r2, dp2, dq2, dr2, N1, N2, coplanar, source, target);
        return permute_t2(p1, q1, r1, p2, r2, q2, dp2, dr2, dq2, N1, N2, coplanar, source, target);
    }
    if (dq1 < 0.0) {
        if (dr1 >= 0.0)
            return permute_t2(q1, r1, p1, p2, r2, q2, dp2, dr2, dq2, N1, N2, coplanar, source, target);
        return permute_t2(p1, q1, r1, p2, q2, r2, dp2, dq2, dr2, N1, N2, coplanar, source, target);
    }
    if (dq1 > 0.0) {
        if (dr1 > 0.0)
            return permute_t2(p1, q1, r1, p2, r2, q2, dp2, dr2, dq2, N1, N2, coplanar, source, target);
        return permute_t2(q1, r1, p1, p2, q2, r2, dp2, dq2, dr2, N1, N2, coplanar, source, target);
    }
    if (dr1 > 0.0)
        return permute_t2(r1, p1, q1, p2, q2, r2, dp2, dq2, dr2, N1, N2, coplanar, source, target);
    if (dr1 < 0.0)
        return permute_t2(r1, p1, q1, p2, r2, q2, dp2, dr2, dq2, N1, N2, coplanar, source, target);
    *coplanar = 1;
    return coplanar_3d(p1, q1, r1, p2, q2, r2, N1);
}

/* 18 doubles per pair; seg = 6 doubles per pair, zero when not written. */
void sbo_tri_tri_batch(const double *tris, size_t n, int32_t *ret, int32_t *coplanar, double *seg)
{
    for (size_t i = 0; i < n; ++i) {
        const double *v = tris + 18 * i;
        int cop = 0;
        double s[3] = {0, 0, 0}, t[3] = {0, 0, 0};
        ret[i] = sbo_tri_tri(v, v + 3, v + 6, v + 9, v + 12, v + 15, &cop, s, t);
        coplanar[i] = cop;
        memcpy(seg + 6 * i, s, sizeof(s));
        memcpy(seg + 6 * i + 3, t, sizeof(t));
    }
}

/* ------------------------------------------------------------------------ */
/* SolidMesh::prepare pieces */

/* Vector3::normal, src/vector3.h:155-176; zero vector when |cross| <= DBL_EPSILON
 * (Double::isZero, src/double.h:31-34). */
static void tri_normal(const double *a, const double *b, const double *c, double *out)
{
    double bax = b[0] - a[0], bay = b[1] - a[1], baz = b[2] - a[2];
    double cax = c[0] - a[0], cay = c[1] - a[1], caz = c[2] - a[2];
    double cx = bay * caz - baz * cay;
    double cy = baz * cax - bax * caz;
    double cz = bax * cay - bay * cax;
    double len2 = cx * cx + cy * cy + cz * cz;
    double len = sqrt(len2);
    if (fabs(len) <= DBL_EPSILON) {
        out[0] = out[1] = out[2] = 0.0;
        return;
    }
    out[0] = cx / len;
    out[1] = cy / len;
    out[2] = cz / len;
}

/* src/solidmesh.cpp:47-55 */
void sbo_normals(const double *xyz, const uint32_t *tri, size_t nT, double *out)
{
    for (size_t i = 0; i < nT; ++i)
        tri_normal(xyz + 3 * tri[3 * i], xyz + 3 * tri[3 * i + 1], xyz + 3 * tri[3 * i + 2], out + 3 * i);
}

/* src/solidmesh.cpp:57-62 + AxisAlignedBoudingBox::update,
 * src/axisalignedboundingbox.h:31-41 (strict > / < against +-DBL_MAX seeds).
 * out: lower xyz, upper xyz per triangle. */
void sbo_tri_boxes(const double *xyz, const uint32_t *tri, size_t nT, double *out)
{
    for (size_t i = 0; i < nT; ++i) {
        double *lo = out + 6 * i, *hi = lo + 3;
        for (int k = 0; k < 3; ++k) {
            lo[k] = DBL_MAX;
            hi[k] = -DBL_MAX;
        }
        for (int c = 0; c < 3; ++c) {
            const double *v = xyz + 3 * tri[3 * i + c];
            for (int k = 0; k < 3; ++k) {
                if (v[k] > hi[k]) hi[k] = v[k];
                if (v[k] < lo[k]) lo[k] = v[k];
            }
        }
    }
}

/* decideGroupSide's query point, src/solidboolean.cpp:497-499:
 * ((v0 + v1) + v2) / 3.0 per component. */
void sbo_centroids(const double *xyz, const uint32_t *tri, size_t nT, double *out)
{
    for (size_t i = 0; i < nT; ++i) {
        const double *a = xyz + 3 * tri[3 * i], *b = xyz + 3 * tri[3 * i + 1], *c = xyz + 3 * tri[3 * i + 2];
        for (int k = 0; k < 3; ++k)
            out[3 * i + k] = ((a[k] + b[k]) + c[k]) / 3.0;
    }
}

/* AxisAlignedBoudingBox::intersectWith, src/axisalignedboundingbox.h:95-105:
 * closed intervals on all three axes. */
static inline int box_overlap(const double *a, const double *b)
{
    for (int k = 0; k < 3; ++k) {
        if (a[k] <= b[3 + k] && a[3 + k] >= b[k])
            continue;
        return 0;
    }
    return 1;
}

/* ------------------------------------------------------------------------ */
/* Broad phase.  AxisAlignedBoudingBoxTree::test (axisalignedboundingboxtree.h:
 * 54-95) returns every (a, b) whose TRIANGLE boxes overlap; inner-node boxes
 * are supersets of their leaves, so the set does not depend on the tree.  The
 * oracle therefore restates the set with its own accelerator (a median-split
 * kd hierarchy over B's boxes, leaves <= 8) and the exact leaf test above. */

typedef struct {
    double box[6];
    uint32_t left, right; /* children, or [begin,end) into `order` when leaf */
    int leaf;
} onode;

typedef struct {
    const double *boxes;
    uint32_t *order;
    onode *nodes;
    size_t nNodes, capNodes;
    int axis; /* qsort scratch */
} otree;

static const double *g_sortBoxes;
static int g_sortAxis;

static int cmp_center(const void *pa, const void *pb)
{
    uint32_t a = *(const uint32_t *)pa, b = *(const uint32_t *)pb;
    double ca = g_sortBoxes[6 * a + g_sortAxis] + g_sortBoxes[6 * a + 3 + g_sortAxis];
    double cb = g_sortBoxes[6 * b + g_sortAxis] + g_sortBoxes[6 * b + 3 + g_sortAxis];
    if (ca < cb) return -1;
    if (ca > cb) return 1;
    return a < b ? -1 : (a > b ? 1 : 0);
}

static uint32_t otree_build(otree *t, uint32_t begin, uint32_t end)
{
    uint32_t id = (uint32_t)t->nNodes++;
    onode *n = &t->nodes[id];
    for (int k = 0; k < 3; ++k) {
        n->box[k] = DBL_MAX;
        n->box[3 + k] = -DBL_MAX;
    }
    for (uint32_t i = begin; i < end; ++i) {
        const double *b = t->boxes + 6 * t->order[i];
        for (int k = 0; k < 3; ++k) {
            if (b[k] < n->box[k]) n->box[k] = b[k];
            if (b[3 + k] > n->box[3 + k]) n->box[3 + k] = b[3 + k];
        }
    }
    if (end - begin <= 8) {
        n->leaf = 1;
        n->left = begin;
        n->right = end;
        return id;
    }
    int axis = 0;
    double best = -1.0;
    for (int k = 0; k < 3; ++k) {
        double span = n->box[3 + k] - n->box[k];
        if (span > best) {
            best = span;
            axis = k;
        }
    }
    g_sortBoxes = t->boxes;
    g_sortAxis = axis;
    qsort(t->order + begin, end - begin, sizeof(uint32_t), cmp_center);
    uint32_t mid = begin + (end - begin) / 2;
    uint32_t l = otree_build(t, begin, mid);
    uint32_t r = otree_build(t, mid, end);
    n = &t->nodes[id]; /* nodes array is pre-sized, pointer stable; re-read for clarity */
    n->leaf = 0;
    n->left = l;
    n->right = r;
    return id;
}

static otree *otree_create(const double *boxes, size_t n)
{
    otree *t = (otree *)calloc(1, sizeof(otree));
    t->boxes = boxes;
    t->order = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
    for (size_t i = 0; i < n; ++i)
        t->order[i] = (uint32_t)i;
    t->capNodes = 2 * n + 2;
    t->nodes = (onode *)malloc(sizeof(onode) * t->capNodes);
    t->nNodes = 0;
    if (n)
        otree_build(t, 0, (uint32_t)n);
    return t;
}

static void otree_free(otree *t)
{
    free(t->order);
    free(t->nodes);
    free(t);
}

typedef struct {
    uint32_t *data;
    size_t n, cap;
} u32vec;

static void u32vec_push(u32vec *v, uint32_t x)
{
    if (v->n == v->cap) {
        v->cap = v->cap ? v->cap * 2 : 1024;
        v->data = (uint32_t *)realloc(v->data, v->cap * sizeof(uint32_t));
    }
    v->data[v->n++] = x;
}

/* All B boxes overlapping `q` (exact closed test), appended to out. */
static void otree_query(const otree *t, const double *q, u32vec *out)
{
    if (!t->nNodes)
        return;
    uint32_t stack[128];
    int top = 0;
    stack[top++] = 0;
    while (top) {
        const onode *n = &t->nodes[stack[--top]];
        if (!box_overlap(n->box, q))
            continue;
        if (n->leaf) {
            for (uint32_t i = n->left; i < n->right; ++i) {
                uint32_t b = t->order[i];
                /* argument order of the reference leaf test: boxes[a].intersectWith(secondBoxes[b]);
                 * the predicate is symmetric, kept as written there. */
                if (box_overlap(q, t->boxes + 6 * b))
                    u32vec_push(out, b);
            }
        } else {
            stack[top++] = n->left;
            stack[top++] = n->right;
        }
    }
}

static int cmp_u32(const void *a, const void *b)
{
    uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

/* Candidate pairs sorted by (a, b).  Returns count; *outPairs is malloc'ed
 * (2 u32 per pair), free with sbo_free. */
size_t sbo_candidate_pairs(const double *boxesA, size_t nA, const double *boxesB, size_t nB, uint32_t **outPairs)
{
    otree *t = otree_create(boxesB, nB);
    u32vec pairs = {0, 0, 0};
    u32vec tmp = {0, 0, 0};
    for (size_t a = 0; a < nA; ++a) {
        tmp.n = 0;
        otree_query(t, boxesA + 6 * a, &tmp);
        if (tmp.n > 1)
            qsort(tmp.data, tmp.n, sizeof(uint32_t), cmp_u32);
        for (size_t i = 0; i < tmp.n; ++i) {
            u32vec_push(&pairs, (uint32_t)a);
            u32vec_push(&pairs, tmp.data[i]);
        }
    }
    free(tmp.data);
    otree_free(t);
    *outPairs = pairs.data;
    return pairs.n / 2;
}

void sbo_free(void *p) { free(p); }

/* Narrow phase over a pair list, as the loop at src/solidboolean.cpp:315-320
 * sees it through intersectTwoFaces (:103-122): hit = ret && !coplanar. */
void sbo_predicate_pairs(const double *xyzA, const uint32_t *triA, const double *xyzB, const uint32_t *triB,
    const uint32_t *pairs, size_t n, int8_t *ret, int8_t *coplanar, uint8_t *hit, double *seg)
{
    for (size_t i = 0; i < n; ++i) {
        const uint32_t *fa = triA + 3 * pairs[2 * i];
        const uint32_t *fb = triB + 3 * pairs[2 * i + 1];
        int cop = 0;
        double s[3] = {0, 0, 0}, t[3] = {0, 0, 0};
        int r = sbo_tri_tri(xyzA + 3 * fa[0], xyzA + 3 * fa[1], xyzA + 3 * fa[2],
            xyzB + 3 * fb[0], xyzB + 3 * fb[1], xyzB + 3 * fb[2], &cop, s, t);
        if (ret) ret[i] = (int8_t)r;
        if (coplanar) coplanar[i] = (int8_t)cop;
        if (hit) hit[i] = (r && !cop) ? 1 : 0;
        if (seg) {
            memcpy(seg + 6 * i, s, sizeof(s));
            memcpy(seg + 6 * i + 3, t, sizeof(t));
        }
    }
}

/* ------------------------------------------------------------------------ */
/* Classification: SolidBoolean::isPointInMesh, src/solidboolean.cpp:48-92 */

/* PositionKey(double,double,double), src/positionkey.cpp:32-37: C truncation of
 * x * 100000 to long. */
typedef struct {
    long x, y, z;
} pkey;

static inline long to_key(double v)
{
    return (long)(v * 100000);
}

static inline int pkey_eq(const pkey *a, const pkey *b)
{
    return a->x == b->x && a->y == b->y && a->z == b->z;
}

/* One ray.  axis k: testAxis = DBL_MAX on component k, DBL_EPSILON elsewhere
 * (g_testAxisList, src/solidboolean.cpp:31-35). */
static int point_in_mesh_axis(const otree *t, const double *xyz, const uint32_t *tri, const double *normals,
    const double *p, int axisIndex, u32vec *cand, pkey **keys, size_t *keyCap)
{
    double axis[3] = {DBL_EPSILON, DBL_EPSILON, DBL_EPSILON};
    axis[axisIndex] = DBL_MAX;
    double end[3] = {p[0] + axis[0], p[1] + axis[1], p[2] + axis[2]};
    /* box.update(testPosition); box.update(testEnd)  (:55-58) */
    double rb[6];
    for (int k = 0; k < 3; ++k) {
        rb[k] = DBL_MAX;
        rb[3 + k] = -DBL_MAX;
        if (p[k] > rb[3 + k]) rb[3 + k] = p[k];
        if (p[k] < rb[k]) rb[k] = p[k];
        if (end[k] > rb[3 + k]) rb[3 + k] = end[k];
        if (end[k] < rb[k]) rb[k] = end[k];
    }
    cand->n = 0;
    otree_query(t, rb, cand);
    size_t nKeys = 0;
    for (size_t ci = 0; ci < cand->n; ++ci) {
        uint32_t f = cand->data[ci];
        const double *t0 = xyz + 3 * tri[3 * f], *t1 = xyz + 3 * tri[3 * f + 1], *t2 = xyz + 3 * tri[3 * f + 2];
        const double *nrm = normals + 3 * f;
        /* Vector3::intersectSegmentAndPlane, src/vector3.h:264-280 */
        double u[3] = {end[0] - p[0], end[1] - p[1], end[2] - p[2]};
        double w[3] = {p[0] - t0[0], p[1] - t0[1], p[2] - t0[2]};
        double d = nrm[0] * u[0] + nrm[1] * u[1] + nrm[2] * u[2];
        double n = (-nrm[0]) * w[0] + (-nrm[1]) * w[1] + (-nrm[2]) * w[2];
        if (fabs(d) <= DBL_EPSILON)
            continue;
        double s = n / d;
        if (s < 0 || s > 1 || isnan(s) || isinf(s))
            continue;
        double hit[3] = {p[0] + s * u[0], p[1] + s * u[1], p[2] + s * u[2]};
        /* three edge normals and two sign tests (:78-86) */
        double n0[3], n1[3], n2[3];
        tri_normal(hit, t0, t1, n0);
        tri_normal(hit, t1, t2, n1);
        tri_normal(hit, t2, t0, n2);
        if (n0[0] * n1[0] + n0[1] * n1[1] + n0[2] * n1[2] > 0 &&
            n0[0] * n2[0] + n0[1] * n2[1] + n0[2] * n2[2] > 0) {
            pkey key = {to_key(hit[0]), to_key(hit[1]), to_key(hit[2])};
            int dup = 0;
            for (size_t j = 0; j < nKeys; ++j)
                if (pkey_eq(&(*keys)[j], &key)) {
                    dup = 1;
                    break;
                }
            if (!dup) {
                if (nKeys == *keyCap) {
                    *keyCap = *keyCap ? *keyCap * 2 : 64;
                    *keys = (pkey *)realloc(*keys, *keyCap * sizeof(pkey));
                }
                (*keys)[nKeys++] = key;
            }
        }
    }
    return (int)(nKeys % 2);
}

/* inside[i] = majority of the three axes, exactly decideGroupSide's
 * (float)insideCount / totalCount > 0.5 (src/solidboolean.cpp:508);
 * perAxis (optional) = 3 bytes per point.  candCount (optional) accumulates
 * the total number of ray/triangle candidates (for roofline accounting). */
void sbo_classify(const double *xyz, const uint32_t *tri, size_t nT,
    const double *pts, size_t q, uint8_t *inside, uint8_t *perAxis, uint64_t *candCount)
{
    double *boxes = (double *)malloc(sizeof(double) * 6 * (nT ? nT : 1));
    double *normals = (double *)malloc(sizeof(double) * 3 * (nT ? nT : 1));
    sbo_tri_boxes(xyz, tri, nT, boxes);
    sbo_normals(xyz, tri, nT, normals);
    otree *t = otree_create(boxes, nT);
    uint64_t total = 0;
#pragma omp parallel reduction(+ : total)
    {
        u32vec cand = {0, 0, 0};
        pkey *keys = NULL;
        size_t keyCap = 0;
#pragma omp for schedule(dynamic, 256)
        for (long long i = 0; i < (long long)q; ++i) {
            size_t insideCount = 0, totalCount = 0;
            for (int k = 0; k < 3; ++k) {
                int in = point_in_mesh_axis(t, xyz, tri, normals, pts + 3 * i, k, &cand, &keys, &keyCap);
                total += cand.n;
                if (perAxis) perAxis[3 * i + k] = (uint8_t)in;
                if (in) ++insideCount;
                ++totalCount;
            }
            if (inside) inside[i] = ((float)insideCount / totalCount > 0.5) ? 1 : 0;
        }
        free(cand.data);
        free(keys);
    }
    if (candCount)
        *candCount = total;
    otree_free(t);
    free(boxes);
    free(normals);
}

/* FNV-1a-64 over a byte buffer (SURVEY section 4 cross-check hashes). */
uint64_t sbo_fnv1a64(const void *data, size_t n)
{
    const unsigned char *p = (const unsigned char *)data;
    uint64_t h = 14695981039346656037ull;
    for (size_t i = 0; i < n; ++i) {
        h ^= p[i];
        h *= 1099511628211ull;
    }
    return h;
}

/* ---- uncut triangles + half-edge map -------------------------------
 * SolidBoolean::addUnintersectedTriangles (src/solidboolean.cpp:250-286): every triangle of
 * `mesh` that is not in usedFaces is appended to m_newTriangles with its vertex ids shifted
 * by the number of vertices already in m_newVertices (:254, :265-269), and its three
 * half-edges (0,1), (1,2), (2,0) are inserted into an unordered_map keyed by
 * makeHalfEdgeKey(first, second) = (first << 32) | second (src/solidboolean.h:75-78) with the
 * new triangle's index as value (:271-282).  The FIRST insertion that meets an existing key
 * makes the function return false at once: the offending triangle is already in
 * m_newTriangles, its earlier half-edges are in the map, nothing after it is visited.
 *
 * The map's iteration order is unspecified, so the contents are returned SORTED BY KEY.
 *   cut            nT bytes (non-zero = in usedFaces) or NULL
 *   outFace        original face id of new triangle k            (capacity nT)
 *   outKeys/Owner  half-edge keys ascending + their triangle      (capacity 3 nT)
 * Returns 1 like the reference's `true`, 0 after a repeated half-edge. */
typedef struct {
    uint64_t key;
    uint32_t owner;
} he_entry;

static int cmp_he(const void *pa, const void *pb)
{
    const he_entry *a = (const he_entry *)pa, *b = (const he_entry *)pb;
    return a->key < b->key ? -1 : a->key > b->key;
}

static inline uint64_t he_hash(uint64_t k)
{
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    return k;
}

int sbo_uncut_half_edges(const uint32_t *tri, size_t nT, const uint8_t *cut, uint64_t vertexOffset,
    uint64_t triangleOffset, uint32_t *outFace, size_t *nTriOut, uint64_t *outKeys, uint32_t *outOwner,
    size_t *nKeysOut)
{
    size_t cap = 16;
    while (cap < 6 * nT + 16)
        cap <<= 1;
    uint64_t *slots = (uint64_t *)malloc(cap * sizeof(uint64_t)); /* open addressing; key + 1 so that 0 = empty */
    he_entry *ent = (he_entry *)malloc((3 * nT + 1) * sizeof(he_entry));
    memset(slots, 0, cap * sizeof(uint64_t));
    size_t nTri = 0, nKeys = 0;
    int ok = 1;
    for (size_t i = 0; i < nT && ok; ++i) {
        if (cut && cut[i])
            continue; /* :261-262 */
        uint64_t v[3] = {tri[3 * i] + vertexOffset, tri[3 * i + 1] + vertexOffset, tri[3 * i + 2] + vertexOffset};
        uint64_t index = triangleOffset + nTri; /* :264 newInsertedIndex */
        outFace[nTri++] = (uint32_t)i;          /* :265 push_back */
        for (int k = 0; k < 3; ++k) {
            uint64_t key = (v[k] << 32) | v[(k + 1) % 3];
            size_t h = (size_t)he_hash(key) & (cap - 1);
            while (slots[h] && slots[h] != key + 1)
                h = (h + 1) & (cap - 1);
            if (slots[h]) { /* :271-282 insert(...).second == false */
                ok = 0;
                break;
            }
            slots[h] = key + 1;
            ent[nKeys].key = key;
            ent[nKeys].owner = (uint32_t)index;
            ++nKeys;
        }
    }
    qsort(ent, nKeys, sizeof(he_entry), cmp_he);
    for (size_t j = 0; j < nKeys; ++j) {
        outKeys[j] = ent[j].key;
        outOwner[j] = ent[j].owner;
    }
    *nTriOut = nTri;
    *nKeysOut = nKeys;
    free(slots);
    free(ent);
    return ok;
}

/* What buildFaceGroups asks the half-edge map (src/solidboolean.cpp:205-224): for edge k of a
 * triangle (from, to) the triangle on the other side = halfEdges.find(key(to, from)).  For new
 * triangle j (face outFace[j], vertex ids shifted by vertexOffset) adj[3 j + k] = that owner,
 * or -1 when the map holds no such half-edge.  keys sorted ascending (as returned above). */
void sbo_uncut_adjacency(const uint32_t *tri, const uint32_t *face, size_t nTri, uint64_t vertexOffset,
    const uint64_t *keys, const uint32_t *owner, size_t nKeys, int32_t *adj)
{
    for (size_t j = 0; j < nTri; ++j) {
        const uint32_t *t = tri + 3 * (size_t)face[j];
        for (int k = 0; k < 3; ++k) {
            uint64_t from = t[k] + vertexOffset, to = t[(k + 1) % 3] + vertexOffset;
            uint64_t want = (to << 32) | from;
            size_t lo = 0, hi = nKeys;
            while (lo < hi) {
                size_t mid = (lo + hi) / 2;
                if (keys[mid] < want)
                    lo = mid + 1;
                else
                    hi = mid;
            }
            adj[3 * j + k] = (lo < nKeys && keys[lo] == want) ? (int32_t)owner[lo] : -1;
        }
    }
}

/* ---- face groups of the untouched triangles ------------------------
 * SolidBoolean::buildFaceGroups (src/solidboolean.cpp:167-239) floods the triangles through
 * the half-edge map: from triangle t, edge (i, j) leads to halfEdges.find(key(j, i)) (:216-221)
 * unless that half-edge is fenced by an intersection loop.  With no loops in the way the
 * groups it opens at :229-238 are the connected components of the adjacency relation above,
 * opened in ascending order of their lowest triangle index.  label[j] = lowest new-triangle
 * index of j's component (adj as returned by sbo_uncut_adjacency, indices carry
 * triangleOffset).  Returns the number of components. */
size_t sbo_uncut_components(const int32_t *adj, size_t nTri, uint64_t triangleOffset, uint32_t *label)
{
    size_t *stack = (size_t *)malloc((nTri + 1) * sizeof(size_t));
    uint8_t *seen = (uint8_t *)calloc(nTri + 1, 1);
    size_t comps = 0;
    for (size_t s = 0; s < nTri; ++s) {
        if (seen[s])
            continue;
        ++comps;
        size_t top = 0;
        stack[top++] = s;
        seen[s] = 1;
        while (top) {
            size_t t = stack[--top];
            label[t] = (uint32_t)(triangleOffset + s);
            for (int k = 0; k < 3; ++k) {
                int32_t o = adj[3 * t + k];
                if (o < 0)
                    continue;
                size_t u = (size_t)((uint64_t)o - triangleOffset);
                /* forward only, like the reference's flood; the relation is symmetric unless a
                 * repeated half-edge truncated the map (then the order of the seeds matters) */
                if (u < nTri && !seen[u]) {
                    seen[u] = 1;
                    stack[top++] = u;
                }
            }
        }
    }
    free(stack);
    free(seen);
    return comps;
}

/* ---- per-triangle intersection contexts -----------------------------
 * The body of the pair loop of SolidBoolean::combine (src/solidboolean.cpp:296-339): for every
 * intersecting pair, in the order of the pair list, the segment's two end points are entered
 * into the context of the first mesh's triangle and into the context of the second mesh's
 * triangle.  A context de-duplicates its points by PositionKey (addIntersectedPoint, :305-310:
 * std::map insert, the FIRST position with a key is kept, indices in first-seen order) and keeps
 * the undirected neighbour relation between the two (3 + index) numbers of a segment unless they
 * are the same point (:323-328 / :332-337).
 *
 * hits: nHit pairs in the order the loop sees them (the CUDA path: ascending (a, b)); seg: 6
 * doubles per hit.  which = 0: contexts of the first mesh's triangles, 1: of the second's.
 * Output, contexts in ascending triangle id:
 *   cutTri[c]                          triangle id
 *   pointStart[c] .. pointStart[c+1]   its points in `points` (3 doubles each), first-seen order
 *   edgeStart[c] .. edgeStart[c+1]     its relations in `edges` (2 x uint32, lower number first,
 *                                      ascending): the flattened neighborMap
 * Capacities: cutTri nHit, pointStart/edgeStart nHit + 1, points 6 nHit doubles, edges 2 nHit.
 * Returns the number of contexts. */
typedef struct {
    uint32_t tri, hit;
} ctx_ref;

static int cmp_ctx_ref(const void *pa, const void *pb)
{
    const ctx_ref *a = (const ctx_ref *)pa, *b = (const ctx_ref *)pb;
    if (a->tri != b->tri)
        return a->tri < b->tri ? -1 : 1;
    return a->hit < b->hit ? -1 : a->hit > b->hit; /* the loop's order inside a triangle */
}

static int cmp_edge(const void *pa, const void *pb)
{
    const uint32_t *a = (const uint32_t *)pa, *b = (const uint32_t *)pb;
    if (a[0] != b[0])
        return a[0] < b[0] ? -1 : 1;
    return a[1] < b[1] ? -1 : a[1] > b[1];
}

size_t sbo_cut_contexts(const uint32_t *hits, const double *seg, size_t nHit, int which, uint32_t *cutTri,
    uint32_t *pointStart, double *points, uint32_t *edgeStart, uint32_t *edges)
{
    ctx_ref *order = (ctx_ref *)malloc((nHit + 1) * sizeof(ctx_ref));
    pkey *keys = (pkey *)malloc((2 * nHit + 2) * sizeof(pkey));
    for (size_t h = 0; h < nHit; ++h) {
        order[h].tri = hits[2 * h + which];
        order[h].hit = (uint32_t)h;
    }
    qsort(order, nHit, sizeof(ctx_ref), cmp_ctx_ref);
    size_t nCtx = 0, nPts = 0, nEdges = 0;
    pointStart[0] = 0;
    edgeStart[0] = 0;
    for (size_t i = 0; i < nHit;) {
        size_t j = i;
        const size_t p0 = nPts, e0 = nEdges;
        while (j < nHit && order[j].tri == order[i].tri) {
            uint32_t idx[2];
            for (int s = 0; s < 2; ++s) {
                const double *p = seg + 6 * (size_t)order[j].hit + 3 * s;
                pkey k = {to_key(p[0]), to_key(p[1]), to_key(p[2])};
                size_t q = p0;
                while (q < nPts && !pkey_eq(&keys[q], &k))
                    ++q;
                if (q == nPts) { /* insertResult.second: a new point */
                    keys[nPts] = k;
                    points[3 * nPts] = p[0];
                    points[3 * nPts + 1] = p[1];
                    points[3 * nPts + 2] = p[2];
                    ++nPts;
                }
                idx[s] = 3 + (uint32_t)(q - p0);
            }
            if (idx[0] != idx[1]) {
                uint32_t lo = idx[0] < idx[1] ? idx[0] : idx[1], hi = idx[0] < idx[1] ? idx[1] : idx[0];
                size_t q = e0;
                while (q < nEdges && !(edges[2 * q] == lo && edges[2 * q + 1] == hi))
                    ++q;
                if (q == nEdges) {
                    edges[2 * nEdges] = lo;
                    edges[2 * nEdges + 1] = hi;
                    ++nEdges;
                }
            }
            ++j;
        }
        qsort(edges + 2 * e0, nEdges - e0, 2 * sizeof(uint32_t), cmp_edge);
        cutTri[nCtx++] = order[i].tri;
        pointStart[nCtx] = (uint32_t)nPts;
        edgeStart[nCtx] = (uint32_t)nEdges;
        i = j;
    }
    free(order);
    free(keys);
    return nCtx;
}

/* ---- buildFaceGroups with intersection loops (the part that is still on the host)
 * SolidBoolean::buildFaceGroups (src/solidboolean.cpp:167-239), restated with its queue order:
 * loop k (a closed vertex cycle) fences its half-edges and seeds group 2k through the triangle on
 * half-edge (v_i, v_i+1) and group 2k+1 through the one on (v_i+1, v_i) (:176-196); the queue is
 * drained first-in first-out, a triangle joins the group of the entry that reaches it first, and an
 * edge is crossed through halfEdges.find(key(j, i)) unless key(i, j) is already in the fence map
 * (:201-224); triangles of [remainingStart, remainingStart + remainingCount) no seed reached open
 * further groups in ascending order (:229-238).
 *   tri        3 vertex ids per triangle (all triangles the map can name)
 *   keys/owner the half-edge map, ascending keys ((first << 32) | second)
 *   loopStart  CSR over loopVerts, nLoops + 1 entries
 * group[t] = group index of triangle t, or 0xffffffff if nothing reached it; order[] = the triangles
 * in the order they were assigned (the concatenation of the reference's groups is a stable sort of it
 * by group).  Returns the number of groups. */
static int he_find(const uint64_t *keys, const uint32_t *owner, size_t n, uint64_t want, uint32_t *out)
{
    size_t lo = 0, hi = n;
    while (lo < hi) {
        size_t mid = (lo + hi) / 2;
        if (keys[mid] < want)
            lo = mid + 1;
        else
            hi = mid;
    }
    if (lo < n && keys[lo] == want) {
        *out = owner[lo];
        return 1;
    }
    return 0;
}

typedef struct {
    uint64_t *slot; /* key + 1, 0 = empty */
    size_t cap;
} u64set;

static int u64set_insert(u64set *s, uint64_t key) /* 1 = newly inserted */
{
    size_t h = (size_t)he_hash(key) & (s->cap - 1);
    while (s->slot[h] && s->slot[h] != key + 1)
        h = (h + 1) & (s->cap - 1);
    if (s->slot[h])
        return 0;
    s->slot[h] = key + 1;
    return 1;
}

static int u64set_has(const u64set *s, uint64_t key)
{
    size_t h = (size_t)he_hash(key) & (s->cap - 1);
    while (s->slot[h] && s->slot[h] != key + 1)
        h = (h + 1) & (s->cap - 1);
    return s->slot[h] != 0;
}

size_t sbo_face_groups(const uint32_t *tri, size_t nTri, const uint64_t *keys, const uint32_t *owner, size_t nKeys,
    const uint32_t *loopStart, const uint32_t *loopVerts, size_t nLoops, size_t remainingStart, size_t remainingCount,
    uint32_t *group, uint32_t *order, size_t *nOrdered)
{
    size_t nLoopEdges = loopStart[nLoops];
    u64set fence;
    fence.cap = 16;
    while (fence.cap < 2 * (3 * nTri + 2 * nLoopEdges) + 16)
        fence.cap <<= 1;
    fence.slot = (uint64_t *)calloc(fence.cap, sizeof(uint64_t));
    /* FIFO of (triangle, group) */
    size_t qcap = 4 * nTri + 2 * nLoopEdges + 16, qh = 0, qt = 0; /* <= 3 per processed triangle + the seeds */
    uint32_t *qTri = (uint32_t *)malloc(qcap * sizeof(uint32_t)), *qGrp = (uint32_t *)malloc(qcap * sizeof(uint32_t));
    for (size_t t = 0; t < nTri; ++t)
        group[t] = 0xffffffffu;
    size_t groups = 0, done = 0;
    for (size_t k = 0; k < nLoops; ++k) {
        size_t n = loopStart[k + 1] - loopStart[k];
        const uint32_t *v = loopVerts + loopStart[k];
        for (size_t i = 0; i < n; ++i) {
            size_t j = (i + 1) % n;
            uint64_t fwd = ((uint64_t)v[i] << 32) | v[j], back = ((uint64_t)v[j] << 32) | v[i];
            uint32_t o;
            u64set_insert(&fence, fwd); /* :181 */
            if (he_find(keys, owner, nKeys, fwd, &o)) {
                qTri[qt] = o;
                qGrp[qt++] = (uint32_t)groups;
            }
            u64set_insert(&fence, back); /* :189 */
            if (he_find(keys, owner, nKeys, back, &o)) {
                qTri[qt] = o;
                qGrp[qt++] = (uint32_t)groups + 1;
            }
        }
        groups += 2;
    }
    size_t next = remainingStart;
    for (;;) {
        while (qh < qt) { /* processQueue, :201-226 */
            uint32_t t = qTri[qh], g = qGrp[qh];
            ++qh;
            if (group[t] != 0xffffffffu)
                continue;
            group[t] = g;
            order[done++] = t;
            for (int i = 0; i < 3; ++i) {
                int j = (i + 1) % 3;
                uint64_t e = ((uint64_t)tri[3 * (size_t)t + i] << 32) | tri[3 * (size_t)t + j];
                if (u64set_has(&fence, e))
                    continue;
                u64set_insert(&fence, e);
                uint32_t o;
                if (he_find(keys, owner, nKeys, ((uint64_t)tri[3 * (size_t)t + j] << 32) | tri[3 * (size_t)t + i], &o)) {
                    qTri[qt] = o;
                    qGrp[qt++] = g;
                }
            }
        }
        while (next < remainingStart + remainingCount && group[next] != 0xffffffffu)
            ++next;
        if (next >= remainingStart + remainingCount)
            break;
        qTri[qt] = (uint32_t)next; /* :232-236 */
        qGrp[qt++] = (uint32_t)groups++;
    }
    *nOrdered = done;
    free(fence.slot);
    free(qTri);
    free(qGrp);
    return groups;
}
