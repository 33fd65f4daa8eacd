// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// C-callable harness around the UNMODIFIED reference sources under
// /root/reference (compiled where they lie; nothing is copied).  Built by
// oracle/Makefile into oracle/_ref/libsbref.so (git-ignored, travels to the
// GPU box as a prebuilt binary).  It is used for three things only:
//   * pinning oracle/sb_oracle.c (the portable C restatement),
//   * generating the golden fixtures under tests/golden/,
//   * the "reference" CPU baseline arm of bench.py.
//
// Entry points mirror the reference call sites:
//   SolidMesh::prepare                          src/solidmesh.cpp:42-76
//   SolidBoolean::searchPotentialIntersectedPairs   src/solidboolean.cpp:94-101
//   SolidBoolean::intersectTwoFaces             src/solidboolean.cpp:103-122
//   SolidBoolean::isPointInMesh                 src/solidboolean.cpp:48-92
//   SolidBoolean::combine / fetch*              src/solidboolean.cpp:288-565
//   SolidBoolean::addUnintersectedTriangles     src/solidboolean.cpp:250-286
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <limits>
#include <map>
#include <queue>
#include <set>
#include <sstream>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include <Eigen/Dense>

// Reach the reference's private members (standard headers are already in).
#define private public
#include "solidboolean.h"
#undef private
#include "tri_tri_intersect.h"

#ifdef SBREF_WITH_OBJ
#define TINYOBJLOADER_IMPLEMENTATION
#include "tiny_obj_loader.h"
#endif

namespace {

using Clock = std::chrono::steady_clock;

double msSince(Clock::time_point t0)
{
    return std::chrono::duration<double, std::milli>(Clock::now() - t0).count();
}

struct RefMesh {
    std::vector<Vector3> vertices;
    std::vector<std::vector<size_t>> triangles;
    SolidMesh mesh;
    double prepareMs = 0.0;
};

struct RefOp {
    RefMesh *a = nullptr;
    RefMesh *b = nullptr;
    SolidBoolean *op = nullptr;
    bool combined = false;
    bool combineOk = false;
    std::vector<std::vector<size_t>> fetched[3];
    // ref_op_uncut: the two half-edge maps, flattened and sorted by key
    std::vector<std::pair<uint64_t, size_t>> halfEdges[2];
    size_t uncutEnd[2] = {0, 0}; // m_newTriangles.size() after each addUnintersectedTriangles call
    double uncutMs[2] = {0, 0};  // time inside addUnintersectedTriangles / buildFaceGroups alone
    double groupsMs[2] = {0, 0};
};

// Silence the reference's std::cout chatter (failure messages) on request.
struct CoutSilencer {
    std::streambuf *old;
    std::ostringstream sink;
    CoutSilencer() : old(std::cout.rdbuf(sink.rdbuf())) {}
    ~CoutSilencer() { std::cout.rdbuf(old); }
};

} // namespace

extern "C" {

void *ref_mesh_create(const double *xyz, size_t nV, const uint32_t *tri, size_t nT)
{
    RefMesh *m = new RefMesh;
    m->vertices.resize(nV);
    for (size_t i = 0; i < nV; ++i)
        m->vertices[i] = Vector3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    m->triangles.resize(nT);
    for (size_t i = 0; i < nT; ++i)
        m->triangles[i] = {(size_t)tri[3 * i], (size_t)tri[3 * i + 1], (size_t)tri[3 * i + 2]};
    m->mesh.setVertices(&m->vertices);
    m->mesh.setTriangles(&m->triangles);
    auto t0 = Clock::now();
    m->mesh.prepare();
    m->prepareMs = msSince(t0);
    return m;
}

double ref_mesh_prepare_ms(void *h) { return ((RefMesh *)h)->prepareMs; }

// Re-run prepare() on a fresh SolidMesh over the same arrays; returns ms.
double ref_mesh_time_prepare(void *h)
{
    RefMesh *m = (RefMesh *)h;
    SolidMesh fresh;
    fresh.setVertices(&m->vertices);
    fresh.setTriangles(&m->triangles);
    auto t0 = Clock::now();
    fresh.prepare();
    return msSince(t0);
}

void ref_mesh_normals(void *h, double *out)
{
    RefMesh *m = (RefMesh *)h;
    const auto &n = *m->mesh.triangleNormals();
    for (size_t i = 0; i < n.size(); ++i)
        for (int k = 0; k < 3; ++k)
            out[3 * i + k] = n[i][k];
}

// out: 6 doubles per triangle (lower xyz, upper xyz)
void ref_mesh_boxes(void *h, double *out)
{
    RefMesh *m = (RefMesh *)h;
    const auto &b = *m->mesh.triangleAxisAlignedBoundingBoxes();
    for (size_t i = 0; i < b.size(); ++i)
        for (int k = 0; k < 3; ++k) {
            out[6 * i + k] = b[i].lowerBound()[k];
            out[6 * i + 3 + k] = b[i].upperBound()[k];
        }
}

void ref_mesh_destroy(void *h) { delete (RefMesh *)h; }

void *ref_op_create(void *a, void *b)
{
    RefOp *o = new RefOp;
    o->a = (RefMesh *)a;
    o->b = (RefMesh *)b;
    o->op = new SolidBoolean(&o->a->mesh, &o->b->mesh);
    return o;
}

void ref_op_destroy(void *h)
{
    RefOp *o = (RefOp *)h;
    delete o->op;
    delete o;
}

// Broad phase exactly as combine() runs it.  Returns pair count; *ms = time.
size_t ref_op_search(void *h, double *ms)
{
    RefOp *o = (RefOp *)h;
    o->op->m_potentialIntersectedPairs.clear();
    auto t0 = Clock::now();
    o->op->searchPotentialIntersectedPairs();
    if (ms)
        *ms = msSince(t0);
    return o->op->m_potentialIntersectedPairs.size();
}

// Pairs in the reference's own (DFS) order: out[2*i] = a, out[2*i+1] = b.
void ref_op_pairs(void *h, uint32_t *out)
{
    RefOp *o = (RefOp *)h;
    const auto &p = o->op->m_potentialIntersectedPairs;
    for (size_t i = 0; i < p.size(); ++i) {
        out[2 * i] = (uint32_t)p[i].first;
        out[2 * i + 1] = (uint32_t)p[i].second;
    }
}

// Narrow phase on an arbitrary pair list (n pairs).  For every pair calls the
// third-party predicate with the argument order of intersectTwoFaces and also
// intersectTwoFaces itself.  ret/coplanar: per pair; seg: 6 doubles per pair
// (zero unless the predicate wrote them); hit = intersectTwoFaces result.
// Returns elapsed ms of the intersectTwoFaces loop alone.
double ref_op_predicate(void *h, const uint32_t *pairs, size_t n,
    int8_t *ret, int8_t *coplanar, uint8_t *hit, double *seg)
{
    RefOp *o = (RefOp *)h;
    const auto &va = o->a->vertices;
    const auto &vb = o->b->vertices;
    for (size_t i = 0; i < n; ++i) {
        const auto &fa = o->a->triangles[pairs[2 * i]];
        const auto &fb = o->b->triangles[pairs[2 * i + 1]];
        int cop = 0;
        double s[3] = {0, 0, 0}, t[3] = {0, 0, 0};
        int r = tri_tri_intersection_test_3d(
            (double *)va[fa[0]].constData(), (double *)va[fa[1]].constData(), (double *)va[fa[2]].constData(),
            (double *)vb[fb[0]].constData(), (double *)vb[fb[1]].constData(), (double *)vb[fb[2]].constData(),
            &cop, s, t);
        if (ret) ret[i] = (int8_t)r;
        if (coplanar) coplanar[i] = (int8_t)cop;
        if (seg) {
            for (int k = 0; k < 3; ++k) {
                seg[6 * i + k] = s[k];
                seg[6 * i + 3 + k] = t[k];
            }
        }
    }
    auto t0 = Clock::now();
    for (size_t i = 0; i < n; ++i) {
        std::pair<Vector3, Vector3> edge;
        bool ok = o->op->intersectTwoFaces(pairs[2 * i], pairs[2 * i + 1], edge);
        if (hit) hit[i] = ok ? 1 : 0;
    }
    return msSince(t0);
}

// Raw predicate on 18 doubles per pair (p1 q1 r1 p2 q2 r2), no mesh needed.
void ref_tri_tri_batch(const double *tris, size_t n, int32_t *ret, int32_t *coplanar, double *seg)
{
    for (size_t i = 0; i < n; ++i) {
        double v[18];
        std::memcpy(v, tris + 18 * i, sizeof(v));
        int cop = 0;
        double s[3] = {0, 0, 0}, t[3] = {0, 0, 0};
        int r = tri_tri_intersection_test_3d(v, v + 3, v + 6, v + 9, v + 12, v + 15, &cop, s, t);
        ret[i] = r;
        coplanar[i] = cop;
        for (int k = 0; k < 3; ++k) {
            seg[6 * i + k] = s[k];
            seg[6 * i + 3 + k] = t[k];
        }
    }
}

// isPointInMesh for Q points against mesh `target` (0 = first, 1 = second)
// on all three reference axes.  perAxis: 3 bytes per point; inside: majority
// vote exactly as decideGroupSide does ((float)insideCount/totalCount > 0.5).
// Returns elapsed ms.
double ref_op_classify(void *h, int target, const double *pts, size_t q,
    uint8_t *inside, uint8_t *perAxis)
{
    RefOp *o = (RefOp *)h;
    const SolidMesh *mesh = target == 0 ? &o->a->mesh : &o->b->mesh;
    static const std::vector<Vector3> axes = {
        {std::numeric_limits<double>::max(), std::numeric_limits<double>::epsilon(), std::numeric_limits<double>::epsilon()},
        {std::numeric_limits<double>::epsilon(), std::numeric_limits<double>::max(), std::numeric_limits<double>::epsilon()},
        {std::numeric_limits<double>::epsilon(), std::numeric_limits<double>::epsilon(), std::numeric_limits<double>::max()},
    };
    auto t0 = Clock::now();
    for (size_t i = 0; i < q; ++i) {
        Vector3 p(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
        size_t insideCount = 0, totalCount = 0;
        for (size_t k = 0; k < 3; ++k) {
            bool in = o->op->isPointInMesh(p, mesh, mesh->axisAlignedBoundingBoxTree(), axes[k]);
            if (perAxis) perAxis[3 * i + k] = in ? 1 : 0;
            if (in) ++insideCount;
            ++totalCount;
        }
        if (inside) inside[i] = ((float)insideCount / totalCount > 0.5) ? 1 : 0;
    }
    return msSince(t0);
}

// Face centroids the way decideGroupSide forms its query point:
// (v0 + v1 + v2) / 3.0 with Vector3 operators (solidboolean.cpp:497-499).
void ref_mesh_centroids(void *h, double *out)
{
    RefMesh *m = (RefMesh *)h;
    for (size_t i = 0; i < m->triangles.size(); ++i) {
        const auto &t = m->triangles[i];
        Vector3 c = (m->vertices[t[0]] + m->vertices[t[1]] + m->vertices[t[2]]) / 3.0;
        out[3 * i] = c.x();
        out[3 * i + 1] = c.y();
        out[3 * i + 2] = c.z();
    }
}

// Full combine().  stageMs (7 doubles): search, process, addUnintersected,
// reTriangulate, buildPolygonsFromEdges, buildFaceGroups, decideGroupSide.
int ref_op_combine(void *h, double *stageMs, int quiet)
{
    RefOp *o = (RefOp *)h;
    delete o->op;
    o->op = new SolidBoolean(&o->a->mesh, &o->b->mesh);
    bool ok;
    if (quiet) {
        CoutSilencer s;
        ok = o->op->combine();
    } else {
        ok = o->op->combine();
    }
    o->combined = true;
    o->combineOk = ok;
    if (stageMs) {
        auto d = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        SolidBoolean *s = o->op;
        stageMs[0] = d(s->benchBegin_searchPotentialIntersectedPairs, s->benchEnd_searchPotentialIntersectedPairs);
        stageMs[1] = d(s->benchBegin_processPotentialIntersectedPairs, s->benchEnd_processPotentialIntersectedPairs);
        stageMs[2] = d(s->benchBegin_addUnintersectedTriangles, s->benchEnd_addUnintersectedTriangles);
        stageMs[3] = d(s->benchBegin_reTriangulate, s->benchEnd_reTriangulate);
        stageMs[4] = d(s->benchBegin_buildPolygonsFromEdges, s->benchEnd_buildPolygonsFromEdges);
        stageMs[5] = d(s->benchBegin_buildFaceGroups, s->benchEnd_buildFaceGroups);
        stageMs[6] = d(s->benchBegin_decideGroupSide, s->benchEnd_decideGroupSide);
    }
    for (auto &f : o->fetched)
        f.clear();
    if (ok) {
        o->op->fetchUnion(o->fetched[0]);
        o->op->fetchDiff(o->fetched[1]);
        o->op->fetchIntersect(o->fetched[2]);
    }
    return ok ? 1 : 0;
}

size_t ref_op_result_vertex_count(void *h) { return ((RefOp *)h)->op->resultVertices().size(); }

void ref_op_result_vertices(void *h, double *out)
{
    const auto &v = ((RefOp *)h)->op->resultVertices();
    for (size_t i = 0; i < v.size(); ++i)
        for (int k = 0; k < 3; ++k)
            out[3 * i + k] = v[i][k];
}

// which: 0 union, 1 diff, 2 intersect
size_t ref_op_result_triangle_count(void *h, int which) { return ((RefOp *)h)->fetched[which].size(); }

void ref_op_result_triangles(void *h, int which, uint32_t *out)
{
    const auto &t = ((RefOp *)h)->fetched[which];
    for (size_t i = 0; i < t.size(); ++i)
        for (int k = 0; k < 3; ++k)
            out[3 * i + k] = (uint32_t)t[i][k];
}

size_t ref_op_group_count(void *h, int which)
{
    RefOp *o = (RefOp *)h;
    return which == 0 ? o->op->m_firstTriangleGroups.size() : o->op->m_secondTriangleGroups.size();
}

#ifdef SBREF_WITH_OBJ
// OBJ loading exactly as test/main.cpp:31-71 does it (tinyobj, float coords
// widened to double, polygons triangulated by the loader).
// Two-call protocol: first with null outputs to get the counts.
int ref_load_obj(const char *path, double *xyz, size_t *nV, uint32_t *tri, size_t *nT)
{
    tinyobj::attrib_t attributes;
    std::vector<tinyobj::shape_t> shapes;
    std::vector<tinyobj::material_t> materials;
    std::string warn, err;
    if (!tinyobj::LoadObj(&attributes, &shapes, &materials, &warn, &err, path))
        return 0;
    size_t vc = attributes.vertices.size() / 3;
    size_t tc = 0;
    for (const auto &shape : shapes)
        tc += shape.mesh.indices.size() / 3;
    if (xyz)
        for (size_t i = 0; i < vc * 3; ++i)
            xyz[i] = attributes.vertices[i];
    if (tri) {
        size_t k = 0;
        for (const auto &shape : shapes)
            for (size_t i = 0; i + 2 < shape.mesh.indices.size(); i += 3) {
                tri[k++] = (uint32_t)shape.mesh.indices[i + 0].vertex_index;
                tri[k++] = (uint32_t)shape.mesh.indices[i + 1].vertex_index;
                tri[k++] = (uint32_t)shape.mesh.indices[i + 2].vertex_index;
            }
    }
    *nV = vc;
    *nT = tc;
    return 1;
}
#endif

// addUnintersectedTriangles for both meshes on a FRESH SolidBoolean, in the order combine()
// calls them (src/solidboolean.cpp:411-421): first mesh with vertex offset 0, then the second
// behind it.  cutA / cutB: per-face bytes (non-zero = the face is in usedFaces).
// nTri[0] / nTri[1] = m_newTriangles.size() after the first / second call, nKeys = map sizes.
// Returns bit 0 = first call's result, bit 1 = second call's.
int ref_op_uncut(void *h, const uint8_t *cutA, const uint8_t *cutB, size_t *nTri, size_t *nKeys, int quiet)
{
    RefOp *o = (RefOp *)h;
    SolidBoolean fresh(&o->a->mesh, &o->b->mesh);
    std::unordered_set<size_t> used[2];
    for (size_t i = 0; i < o->a->triangles.size(); ++i)
        if (cutA && cutA[i])
            used[0].insert(i);
    for (size_t i = 0; i < o->b->triangles.size(); ++i)
        if (cutB && cutB[i])
            used[1].insert(i);
    std::unordered_map<uint64_t, size_t> maps[2];
    int result = 0;
    {
        CoutSilencer *s = quiet ? new CoutSilencer : nullptr;
        auto t0 = Clock::now();
        if (fresh.addUnintersectedTriangles(&o->a->mesh, used[0], &maps[0]))
            result |= 1;
        o->uncutMs[0] = msSince(t0);
        nTri[0] = fresh.m_newTriangles.size();
        t0 = Clock::now();
        if (fresh.addUnintersectedTriangles(&o->b->mesh, used[1], &maps[1]))
            result |= 2;
        o->uncutMs[1] = msSince(t0);
        nTri[1] = fresh.m_newTriangles.size();
        delete s;
    }
    o->uncutEnd[0] = nTri[0];
    o->uncutEnd[1] = nTri[1];
    o->fetched[0] = fresh.m_newTriangles; // reuse the triple store for ref_op_result_triangles(h, 0, ..)
    for (int w = 0; w < 2; ++w) {
        o->halfEdges[w].assign(maps[w].begin(), maps[w].end());
        std::sort(o->halfEdges[w].begin(), o->halfEdges[w].end());
        nKeys[w] = o->halfEdges[w].size();
    }
    return result;
}

void ref_op_uncut_fetch(void *h, int which, uint64_t *keys, uint32_t *owner)
{
    RefOp *o = (RefOp *)h;
    for (size_t i = 0; i < o->halfEdges[which].size(); ++i) {
        keys[i] = o->halfEdges[which][i].first;
        owner[i] = (uint32_t)o->halfEdges[which][i].second;
    }
}

// buildFaceGroups' neighbour lookup (src/solidboolean.cpp:205-224) against the map kept by
// ref_op_uncut: the owner of half-edge (from, to), or -1.
void ref_op_uncut_lookup(void *h, int which, const uint64_t *fromTo, size_t n, int32_t *out)
{
    RefOp *o = (RefOp *)h;
    std::unordered_map<uint64_t, size_t> m(o->halfEdges[which].begin(), o->halfEdges[which].end());
    for (size_t i = 0; i < n; ++i) {
        auto it = m.find(SolidBoolean::makeHalfEdgeKey(fromTo[2 * i], fromTo[2 * i + 1]));
        out[i] = it == m.end() ? -1 : (int32_t)it->second;
    }
}

// buildFaceGroups (src/solidboolean.cpp:167-239) over the triangles and the half-edge map kept
// by ref_op_uncut, with NO intersection loops: the flood of :229-238 alone.  label[j] = first
// triangle of the group that new triangle (start + j) landed in; returns the number of groups.
size_t ref_op_uncut_groups(void *h, int which, uint32_t *label)
{
    RefOp *o = (RefOp *)h;
    SolidBoolean fresh(&o->a->mesh, &o->b->mesh);
    std::unordered_map<uint64_t, size_t> m(o->halfEdges[which].begin(), o->halfEdges[which].end());
    std::vector<std::vector<size_t>> none, groups;
    size_t start = which == 0 ? 0 : o->uncutEnd[0];
    size_t count = o->uncutEnd[which] - start;
    auto t0 = Clock::now();
    fresh.buildFaceGroups(none, m, o->fetched[0], start, count, groups);
    o->groupsMs[which] = msSince(t0);
    for (const auto &g : groups)
        for (size_t t : g)
            label[t - start] = (uint32_t)g[0];
    return groups.size();
}

// buildFaceGroups (src/solidboolean.cpp:167-239) on explicit inputs: triangles (3 ids each), a
// half-edge map (keys / owner), intersection loops (CSR of vertex cycles), the range of "remaining"
// triangles.  group[t] = index of the group triangle t landed in (0xffffffff: none); order = the
// groups' members concatenated group by group.  Returns the number of groups.
size_t ref_face_groups(const uint32_t *tri, size_t nTri, const uint64_t *keys, const uint32_t *owner, size_t nKeys,
    const uint32_t *loopStart, const uint32_t *loopVerts, size_t nLoops, size_t remainingStart, size_t remainingCount,
    uint32_t *group, uint32_t *order)
{
    std::vector<Vector3> noVertices;
    std::vector<std::vector<size_t>> noTriangles;
    SolidMesh ma, mb;
    SolidBoolean fresh(&ma, &mb);
    std::vector<std::vector<size_t>> triangles(nTri), loops(nLoops), groups;
    for (size_t t = 0; t < nTri; ++t)
        triangles[t] = {tri[3 * t], tri[3 * t + 1], tri[3 * t + 2]};
    for (size_t k = 0; k < nLoops; ++k)
        loops[k].assign(loopVerts + loopStart[k], loopVerts + loopStart[k + 1]);
    std::unordered_map<uint64_t, size_t> map;
    for (size_t i = 0; i < nKeys; ++i)
        map.insert({keys[i], owner[i]});
    fresh.buildFaceGroups(loops, map, triangles, remainingStart, remainingCount, groups);
    for (size_t t = 0; t < nTri; ++t)
        group[t] = 0xffffffffu;
    size_t at = 0;
    for (size_t g = 0; g < groups.size(); ++g)
        for (size_t t : groups[g]) {
            group[t] = (uint32_t)g;
            order[at++] = (uint32_t)t;
        }
    return groups.size();
}

// milliseconds spent inside the reference's own functions by the last ref_op_uncut /
// ref_op_uncut_groups: what = 0 addUnintersectedTriangles, 1 buildFaceGroups
double ref_op_uncut_ms(void *h, int what, int which)
{
    RefOp *o = (RefOp *)h;
    return what == 0 ? o->uncutMs[which] : o->groupsMs[which];
}

} // extern "C"
