// Flat C entry points over the C++ classes, for scripted tests (ctypes): build two
// meshes, combine, fetch the three results.  Not part of the product ABI.
#include <cstdint>
#include <cstring>
#include <iostream>
#include <mutex>
#include <sstream>
#include <streambuf>
#include <string>
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

// The reference-style messages (std::cout) are kept for the caller.  Several jobs may run at once (the classes allow it:
// tests/test_host_cpp.py::test_concurrent_booleans_from_several_threads_gpu), so std::cout is redirected ONCE while any job
// is active -- into one sink whose writes are locked -- and handed back by the last job to leave.  (A per-job
// rdbuf swap handed a finished job's dead buffer back to std::cout when two jobs overlapped: a crash at process exit.)
class LockedSink : public std::streambuf
{
public:
    size_t size()
    {
        std::lock_guard<std::mutex> l(m_mu);
        return m_text.size();
    }
    std::string since(size_t from)
    {
        std::lock_guard<std::mutex> l(m_mu);
        return from < m_text.size() ? m_text.substr(from) : std::string();
    }
protected:
    int_type overflow(int_type ch) override
    {
        if (ch != traits_type::eof()) {
            std::lock_guard<std::mutex> l(m_mu);
            m_text.push_back((char)ch);
        }
        return ch;
    }
    std::streamsize xsputn(const char *s, std::streamsize n) override
    {
        std::lock_guard<std::mutex> l(m_mu);
        m_text.append(s, (size_t)n);
        return n;
    }
private:
    std::mutex m_mu;
    std::string m_text;
};

std::mutex g_redirectMu;
int g_activeJobs = 0;
LockedSink *g_sink = nullptr;
std::streambuf *g_coutBuf = nullptr;

size_t redirect_enter()
{
    std::lock_guard<std::mutex> l(g_redirectMu);
    if (g_activeJobs++ == 0) {
        g_sink = new LockedSink;
        g_coutBuf = std::cout.rdbuf(g_sink);
    }
    return g_sink->size();
}

std::string redirect_leave(size_t from)
{
    std::lock_guard<std::mutex> l(g_redirectMu);
    std::string text = g_sink->since(from);
    if (--g_activeJobs == 0) {
        std::cout.rdbuf(g_coutBuf);
        delete g_sink;
        g_sink = nullptr;
    }
    return text;
}

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
    const size_t logFrom = redirect_enter(); // keep the reference-style messages for the caller
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
    j->log = redirect_leave(logFrom);
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
