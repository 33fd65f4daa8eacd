// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Observation hook for the UNMODIFIED reference (oracle/_ref/libsbref.so): the per-triangle
// IntersectedContext objects that SolidBoolean::combine() builds (src/solidboolean.cpp:296-339)
// are local variables of that function, but every one of them is handed to
// ReTriangulator::setEdges(points, &neighborMap) (:366-367), a call that crosses translation
// units and therefore goes through the PLT of the reference library.  This library, loaded
// BEFORE libsbref.so (LD_PRELOAD), supplies that symbol: it records the arguments the
// reference's own loop produced and forwards to the real function.  It also wraps
// SolidBoolean::searchPotentialIntersectedPairs() (:94-101, same mechanism) so that a test can
// ask for the pair list in ascending (first, second) order -- the order in which the CUDA path
// delivers its hits -- instead of the reference's tree-traversal order; the loop body that
// consumes the pairs is the reference's own either way.
// Built by oracle/Makefile into oracle/_ref/libsbref_hook.so; used by
// tests/golden/make_golden.py --contexts (fixtures) and tests/test_oracle_pinning.py (live).
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <dlfcn.h>
#include <limits>
#include <map>
#include <set>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include <Eigen/Dense>

#define private public
#include "solidboolean.h"
#include "retriangulator.h"
#undef private

namespace {

struct Captured {
    double origin[3], normal[3], axis[3]; // projection origin (= the triangle's first vertex), plane normal, unit(v1 - v0)
    std::vector<double> points;  // 3 per context point
    std::vector<uint32_t> edges; // 2 per undirected neighbour relation (a < b), indices as the reference numbers them (3 + point)
};
std::vector<Captured> g_captured;
bool g_sorted = false;

void *real(const char *name)
{
    static void *lib = nullptr;
    if (!lib) {
        const char *path = std::getenv("SBREF_PATH");
        lib = dlopen(path ? path : "libsbref.so", RTLD_NOW | RTLD_NOLOAD);
        if (!lib)
            lib = dlopen(path ? path : "libsbref.so", RTLD_NOW);
    }
    void *f = lib ? dlsym(lib, name) : nullptr;
    if (!f) {
        std::fprintf(stderr, "ref_hook: cannot resolve %s in the reference library\n", name);
        std::abort();
    }
    return f;
}

} // namespace

void ReTriangulator::setEdges(const std::vector<Vector3> &points,
    const std::unordered_map<size_t, std::unordered_set<size_t>> *neighborMapFrom3)
{
    Captured c;
    for (int k = 0; k < 3; ++k) {
        c.origin[k] = m_projectOrigin[k];
        c.normal[k] = m_projectNormal[k];
        c.axis[k] = m_projectAxis[k];
    }
    for (const auto &p : points)
        for (int k = 0; k < 3; ++k)
            c.points.push_back(p[k]);
    std::set<std::pair<uint32_t, uint32_t>> e;
    for (const auto &it : *neighborMapFrom3)
        for (size_t other : it.second)
            e.insert({(uint32_t)std::min(it.first, other), (uint32_t)std::max(it.first, other)});
    for (const auto &p : e) {
        c.edges.push_back(p.first);
        c.edges.push_back(p.second);
    }
    g_captured.push_back(std::move(c));
    typedef void (*Fn)(ReTriangulator *, const std::vector<Vector3> &, const std::unordered_map<size_t, std::unordered_set<size_t>> *);
    static Fn fn = (Fn)real("_ZN14ReTriangulator8setEdgesERKSt6vectorI7Vector3SaIS1_EEPKSt13unordered_mapImSt13unordered_setImSt4hashImESt8equal_toImESaImEES9_SB_SaISt4pairIKmSD_EEE");
    fn(this, points, neighborMapFrom3);
}

void SolidBoolean::searchPotentialIntersectedPairs()
{
    typedef void (*Fn)(SolidBoolean *);
    static Fn fn = (Fn)real("_ZN12SolidBoolean31searchPotentialIntersectedPairsEv");
    fn(this);
    if (g_sorted)
        std::sort(m_potentialIntersectedPairs.begin(), m_potentialIntersectedPairs.end());
}

extern "C" {
void hook_reset(int sortedPairs)
{
    g_captured.clear();
    g_sorted = sortedPairs != 0;
}
size_t hook_count(void) { return g_captured.size(); }
void hook_sizes(size_t i, size_t *nPoints, size_t *nEdges)
{
    *nPoints = g_captured[i].points.size() / 3;
    *nEdges = g_captured[i].edges.size() / 2;
}
void hook_get(size_t i, double *origin3, double *normal3, double *axis3, double *points, uint32_t *edges)
{
    const Captured &c = g_captured[i];
    for (int k = 0; k < 3; ++k) {
        origin3[k] = c.origin[k];
        normal3[k] = c.normal[k];
        axis3[k] = c.axis[k];
    }
    std::copy(c.points.begin(), c.points.end(), points);
    std::copy(c.edges.begin(), c.edges.end(), edges);
}
}
