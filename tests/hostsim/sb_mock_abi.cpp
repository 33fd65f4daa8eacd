// TEST ONLY: a CPU stand-in for the subset of the C ABI the C++ host classes call,
// backed by the oracle (oracle/libsboracle.so), so that the HOST logic
// (retriangulation, welding, face groups, fetch*) can be exercised in the CPU test
// suite.  It is linked only into tests/hostsim/libsbhost_mock.so; the product
// library solidboolean_b200/lib/libsolidboolean_host.so links the CUDA library and
// has no CPU path.
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/solidboolean_b200.h"

extern "C" {
void sbo_normals(const double *xyz, const uint32_t *tri, size_t nT, double *out);
void sbo_tri_boxes(const double *xyz, const uint32_t *tri, size_t nT, double *out);
size_t sbo_candidate_pairs(const double *boxesA, size_t nA, const double *boxesB, size_t nB, uint32_t **outPairs);
void sbo_free(void *p);
void sbo_predicate_pairs(const double *xyzA, const uint32_t *triA, const double *xyzB, const uint32_t *triB,
    const uint32_t *pairs, size_t n, int8_t *ret, int8_t *coplanar, uint8_t *hit, double *seg);
void sbo_classify(const double *xyz, const uint32_t *tri, size_t nT, const double *pts, size_t q, uint8_t *inside,
    uint8_t *perAxis, uint64_t *candCount);
int sbo_uncut_half_edges(const uint32_t *tri, size_t nT, const uint8_t *cut, uint64_t vertexOffset, uint64_t triangleOffset,
    uint32_t *outFace, size_t *nTriOut, uint64_t *outKeys, uint32_t *outOwner, size_t *nKeysOut);
void sbo_uncut_adjacency(const uint32_t *tri, const uint32_t *face, size_t nTri, uint64_t vertexOffset,
    const uint64_t *keys, const uint32_t *owner, size_t nKeys, int32_t *adj);
size_t sbo_uncut_components(const int32_t *adj, size_t nTri, uint64_t triangleOffset, uint32_t *label);
size_t sbo_cut_contexts(const uint32_t *hits, const double *seg, size_t nHit, int which, uint32_t *cutTri,
    uint32_t *pointStart, double *points, uint32_t *edgeStart, uint32_t *edges);
}

struct sb_context { int dummy; };
struct sb_mesh {
    std::vector<double> xyz;
    std::vector<uint32_t> tri;
};
struct sb_isect {
    const sb_mesh *A = nullptr, *B = nullptr;
    size_t nCand = 0;
    std::vector<uint32_t> hitAB;
    std::vector<double> hitSeg;
};
struct sb_cuts {
    std::vector<uint32_t> tri, pointStart, edgeStart, edges;
    std::vector<double> points;
};
struct sb_uncut {
    const sb_mesh *mesh = nullptr;
    uint64_t vertexOffset = 0, triangleOffset = 0;
    int ok = 1;
    std::vector<uint32_t> face, tri3, owner;
    std::vector<uint64_t> keys;
};

extern "C" {

const char *sb_last_error(void) { return "mock"; }
int sb_context_create(int, sb_context **out) { *out = new sb_context; return SB_OK; }
void sb_context_destroy(sb_context *c) { delete c; }

int sb_mesh_create(sb_context *, const double *xyz, size_t nV, const uint32_t *tri, size_t nT, sb_mesh **out)
{
    sb_mesh *m = new sb_mesh;
    m->xyz.assign(xyz, xyz + 3 * nV);
    m->tri.assign(tri, tri + 3 * nT);
    *out = m;
    return SB_OK;
}
void sb_mesh_destroy(sb_mesh *m) { delete m; }
int sb_mesh_normals(const sb_mesh *m, double *out) { sbo_normals(m->xyz.data(), m->tri.data(), m->tri.size() / 3, out); return SB_OK; }
int sb_mesh_triangle_boxes(const sb_mesh *m, double *out) { sbo_tri_boxes(m->xyz.data(), m->tri.data(), m->tri.size() / 3, out); return SB_OK; }

int sb_intersect(const sb_mesh *A, const sb_mesh *B, unsigned, sb_isect **out)
{
    size_t nA = A->tri.size() / 3, nB = B->tri.size() / 3;
    std::vector<double> ba(6 * nA + 6), bb(6 * nB + 6);
    sbo_tri_boxes(A->xyz.data(), A->tri.data(), nA, ba.data());
    sbo_tri_boxes(B->xyz.data(), B->tri.data(), nB, bb.data());
    uint32_t *pairs = nullptr;
    size_t n = sbo_candidate_pairs(ba.data(), nA, bb.data(), nB, &pairs);
    std::vector<uint8_t> hit(n + 1);
    std::vector<double> seg(6 * n + 6);
    if (n)
        sbo_predicate_pairs(A->xyz.data(), A->tri.data(), B->xyz.data(), B->tri.data(), pairs, n, nullptr, nullptr,
            hit.data(), seg.data());
    sb_isect *x = new sb_isect;
    x->A = A;
    x->B = B;
    x->nCand = n;
    for (size_t i = 0; i < n; ++i)
        if (hit[i]) {
            x->hitAB.push_back(pairs[2 * i]);
            x->hitAB.push_back(pairs[2 * i + 1]);
            x->hitSeg.insert(x->hitSeg.end(), seg.begin() + 6 * i, seg.begin() + 6 * i + 6);
        }
    sbo_free(pairs);
    *out = x;
    return SB_OK;
}
void sb_isect_destroy(sb_isect *x) { delete x; }
int sb_isect_counts(const sb_isect *x, size_t *nCand, size_t *nHit)
{
    if (nCand) *nCand = x->nCand;
    if (nHit) *nHit = x->hitAB.size() / 2;
    return SB_OK;
}
int sb_isect_hits(const sb_isect *x, uint32_t *ab, double *seg)
{
    if (ab && !x->hitAB.empty()) std::memcpy(ab, x->hitAB.data(), 4 * x->hitAB.size());
    if (seg && !x->hitSeg.empty()) std::memcpy(seg, x->hitSeg.data(), 8 * x->hitSeg.size());
    return SB_OK;
}
int sb_isect_contexts(const sb_isect *x, int which, sb_cuts **out)
{
    size_t n = x->hitAB.size() / 2;
    sb_cuts *k = new sb_cuts;
    k->tri.resize(n + 1);
    k->pointStart.assign(n + 1, 0);
    k->edgeStart.assign(n + 1, 0);
    k->points.resize(6 * n + 1);
    k->edges.resize(2 * n + 1);
    size_t c = n ? sbo_cut_contexts(x->hitAB.data(), x->hitSeg.data(), n, which, k->tri.data(), k->pointStart.data(),
                       k->points.data(), k->edgeStart.data(), k->edges.data())
                 : 0;
    k->tri.resize(c);
    k->pointStart.resize(c + 1);
    k->edgeStart.resize(c + 1);
    *out = k;
    return SB_OK;
}
void sb_cuts_destroy(sb_cuts *k) { delete k; }
int sb_cuts_counts(const sb_cuts *k, size_t *nc, size_t *np, size_t *ne)
{
    if (nc) *nc = k->tri.size();
    if (np) *np = k->pointStart.back();
    if (ne) *ne = k->edgeStart.back();
    return SB_OK;
}
int sb_cuts_fetch(const sb_cuts *k, uint32_t *tri, uint32_t *ps, double *points, uint32_t *es, uint32_t *edges)
{
    if (tri && !k->tri.empty()) std::memcpy(tri, k->tri.data(), 4 * k->tri.size());
    if (ps) std::memcpy(ps, k->pointStart.data(), 4 * k->pointStart.size());
    if (es) std::memcpy(es, k->edgeStart.data(), 4 * k->edgeStart.size());
    if (points && k->pointStart.back()) std::memcpy(points, k->points.data(), 24 * (size_t)k->pointStart.back());
    if (edges && k->edgeStart.back()) std::memcpy(edges, k->edges.data(), 8 * (size_t)k->edgeStart.back());
    return SB_OK;
}
int sb_isect_uncut(const sb_isect *x, int which, size_t vertexOffset, size_t triangleOffset, sb_uncut **out)
{
    const sb_mesh *m = which == 0 ? x->A : x->B;
    size_t nT = m->tri.size() / 3;
    std::vector<uint8_t> cut(nT + 1, 0);
    for (size_t h = 0; h < x->hitAB.size() / 2; ++h)
        cut[x->hitAB[2 * h + which]] = 1;
    sb_uncut *u = new sb_uncut;
    u->mesh = m;
    u->vertexOffset = vertexOffset;
    u->triangleOffset = triangleOffset;
    u->face.resize(nT + 1);
    u->keys.resize(3 * nT + 1);
    u->owner.resize(3 * nT + 1);
    size_t nTri = 0, nKeys = 0;
    u->ok = sbo_uncut_half_edges(m->tri.data(), nT, cut.data(), vertexOffset, triangleOffset, u->face.data(), &nTri,
        u->keys.data(), u->owner.data(), &nKeys);
    u->face.resize(nTri);
    u->keys.resize(nKeys);
    u->owner.resize(nKeys);
    for (size_t j = 0; j < nTri; ++j)
        for (int k = 0; k < 3; ++k)
            u->tri3.push_back(m->tri[3 * (size_t)u->face[j] + k] + (uint32_t)vertexOffset);
    *out = u;
    return SB_OK;
}
void sb_uncut_destroy(sb_uncut *u) { delete u; }
int sb_uncut_counts(const sb_uncut *u, size_t *nTri, size_t *nKeys, int *ok)
{
    if (nTri) *nTri = u->face.size();
    if (nKeys) *nKeys = u->keys.size();
    if (ok) *ok = u->ok;
    return SB_OK;
}
int sb_uncut_triangles(const sb_uncut *u, uint32_t *face, uint32_t *tri3)
{
    if (face && !u->face.empty()) std::memcpy(face, u->face.data(), 4 * u->face.size());
    if (tri3 && !u->tri3.empty()) std::memcpy(tri3, u->tri3.data(), 4 * u->tri3.size());
    return SB_OK;
}
int sb_uncut_half_edges(const sb_uncut *u, uint64_t *keys, uint32_t *owner)
{
    if (keys && !u->keys.empty()) std::memcpy(keys, u->keys.data(), 8 * u->keys.size());
    if (owner && !u->owner.empty()) std::memcpy(owner, u->owner.data(), 4 * u->owner.size());
    return SB_OK;
}
int sb_uncut_adjacency(const sb_uncut *u, int32_t *adj3)
{
    if (!u->face.empty())
        sbo_uncut_adjacency(u->mesh->tri.data(), u->face.data(), u->face.size(), u->vertexOffset, u->keys.data(),
            u->owner.data(), u->keys.size(), adj3);
    return SB_OK;
}
int sb_uncut_components(const sb_uncut *u, uint32_t *label, size_t *n)
{
    std::vector<int32_t> adj(3 * u->face.size() + 3);
    sb_uncut_adjacency(u, adj.data());
    std::vector<uint32_t> lab(u->face.size() + 1);
    size_t c = sbo_uncut_components(adj.data(), u->face.size(), u->triangleOffset, lab.data());
    if (label && !u->face.empty()) std::memcpy(label, lab.data(), 4 * u->face.size());
    if (n) *n = c;
    return SB_OK;
}
// (no edge tags from the oracle: the mirror then attaches polyline ends geometrically)
int sb_isect_hit_edges(const sb_isect *, uint8_t *) { return SB_ERR_INVALID; }
// (the flood over the pieces is a device kernel: the stand-in refuses, the mirror then runs its host flood)
int sb_uncut_face_groups(const sb_uncut *, const uint32_t *, size_t, const uint32_t *, size_t, uint32_t *, uint32_t *, size_t *)
{
    return SB_ERR_INVALID;
}
int sb_classify(const sb_mesh *t, const double *pts, size_t Q, uint8_t *inside, uint8_t *per_axis)
{
    if (Q)
        sbo_classify(t->xyz.data(), t->tri.data(), t->tri.size() / 3, pts, Q, inside, per_axis, nullptr);
    return SB_OK;
}
}
