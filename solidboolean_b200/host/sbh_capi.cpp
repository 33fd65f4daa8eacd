// Flat C entry points over the C++ classes, for scripted tests (ctypes): build two
// meshes, combine, fetch the three results.  Not part of the product ABI.
#include <cstdint>
#include <cstring>
#include <iostream>
#include <sstream>
#include <vector>
#include "solidboolean.h"

namespace
{
struct Job {
    std::vector<Vector3> va, vb;
    std::vector<std::vector<size_t>> ta, tb;
    SolidMesh a, b;
    SolidBoolean *op = nullptr;
    bool ok = false;
    std::vector<std::vector<size_t>> result[3];
    std::string log;
    double stageMs[7] = {0, 0, 0, 0, 0, 0, 0};
    ~Job() { delete op; }
};

void fill(std::vector<Vector3> &v, std::vector<std::vector<size_t>> &t, const double *xyz, size_t nV, const uint32_t *tri, size_t nT)
{
    v.resize(nV);
    for (size_t i = 0; i < nV; ++i)
        v[i] = Vector3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    t.resize(nT);
    for (size_t i = 0; i < nT; ++i)
        t[i] = {(size_t)tri[3 * i], (size_t)tri[3 * i + 1], (size_t)tri[3 * i + 2]};
}
}

extern "C" {

void *sbh_boolean(const double *xyzA, size_t nVA, const uint32_t *triA, size_t nTA,
    const double *xyzB, size_t nVB, const uint32_t *triB, size_t nTB)
{
    Job *j = new Job;
    fill(j->va, j->ta, xyzA, nVA, triA, nTA);
    fill(j->vb, j->tb, xyzB, nVB, triB, nTB);
    std::ostringstream sink;
    std::streambuf *old = std::cout.rdbuf(sink.rdbuf()); // keep the reference-style messages for the caller
    j->a.setVertices(&j->va);
    j->a.setTriangles(&j->ta);
    j->a.prepare();
    j->b.setVertices(&j->vb);
    j->b.setTriangles(&j->tb);
    j->b.prepare();
    j->op = new SolidBoolean(&j->a, &j->b);
    j->ok = j->op->combine();
    if (j->ok) {
        j->op->fetchUnion(j->result[0]);
        j->op->fetchDiff(j->result[1]);
        j->op->fetchIntersect(j->result[2]);
    }
    std::cout.rdbuf(old);
    j->log = sink.str();
    auto ms = [](SolidBoolean::TimePoint a, SolidBoolean::TimePoint b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    SolidBoolean *s = j->op;
    j->stageMs[0] = ms(s->benchBegin_searchPotentialIntersectedPairs, s->benchEnd_searchPotentialIntersectedPairs);
    j->stageMs[1] = ms(s->benchBegin_processPotentialIntersectedPairs, s->benchEnd_processPotentialIntersectedPairs);
    j->stageMs[2] = ms(s->benchBegin_addUnintersectedTriangles, s->benchEnd_addUnintersectedTriangles);
    j->stageMs[3] = ms(s->benchBegin_reTriangulate, s->benchEnd_reTriangulate);
    j->stageMs[4] = ms(s->benchBegin_buildPolygonsFromEdges, s->benchEnd_buildPolygonsFromEdges);
    j->stageMs[5] = ms(s->benchBegin_buildFaceGroups, s->benchEnd_buildFaceGroups);
    j->stageMs[6] = ms(s->benchBegin_decideGroupSide, s->benchEnd_decideGroupSide);
    return j;
}

int sbh_ok(void *h) { return ((Job *)h)->ok ? 1 : 0; }
const char *sbh_log(void *h) { return ((Job *)h)->log.c_str(); }
size_t sbh_candidates(void *h) { return ((Job *)h)->op->candidatePairCount(); }
size_t sbh_hits(void *h) { return ((Job *)h)->op->intersectingPairCount(); }
size_t sbh_vertex_count(void *h) { return ((Job *)h)->op->resultVertices().size(); }
void sbh_vertices(void *h, double *out)
{
    const auto &v = ((Job *)h)->op->resultVertices();
    if (!v.empty())
        std::memcpy(out, v[0].constData(), sizeof(double) * 3 * v.size());
}
size_t sbh_triangle_count(void *h, int which) { return ((Job *)h)->result[which].size(); }
void sbh_triangles(void *h, int which, uint32_t *out)
{
    const auto &t = ((Job *)h)->result[which];
    for (size_t i = 0; i < t.size(); ++i)
        for (int k = 0; k < 3; ++k)
            out[3 * i + k] = (uint32_t)t[i][k];
}
void sbh_stage_ms(void *h, double *out7) { std::memcpy(out7, ((Job *)h)->stageMs, sizeof(double) * 7); }
void sbh_free(void *h) { delete (Job *)h; }
}
