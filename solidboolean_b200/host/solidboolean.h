// SolidBoolean with the reference's public interface (reference
// src/solidboolean.h:33-108).  combine() runs the intersection front end on the
// GPU through the C ABI -- broad phase + Guigue-Devillers predicate
// (sb_intersect) and the inside/outside test of every face group (sb_classify) --
// and keeps retriangulation, vertex welding, face grouping and mesh assembly on
// the host, as the reference does.  Errors follow the reference's convention:
// message on std::cout, combine() returns false, no exceptions.
#ifndef SOLID_BOOLEAN_H
#define SOLID_BOOLEAN_H
#include <array>
#include <chrono>
#include <cstdint>
#include <map>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include "positionkey.h"
#include "solidmesh.h"
#include "vector3.h"

class SolidBoolean
{
public:
    SolidBoolean(const SolidMesh *firstMesh, const SolidMesh *secondMesh);
    ~SolidBoolean();
    bool combine();
    // each call APPENDS index triples into resultVertices()
    void fetchUnion(std::vector<std::vector<size_t>> &resultTriangles);
    void fetchDiff(std::vector<std::vector<size_t>> &resultTriangles);
    void fetchIntersect(std::vector<std::vector<size_t>> &resultTriangles);
    // first mesh's vertices, then the second mesh's, then the welded new points
    const std::vector<Vector3> &resultVertices();

    // stage time points, same names as the reference (src/solidboolean.h:45-58)
    typedef std::chrono::time_point<std::chrono::high_resolution_clock> TimePoint;
    TimePoint benchBegin_searchPotentialIntersectedPairs, benchEnd_searchPotentialIntersectedPairs;
    TimePoint benchBegin_processPotentialIntersectedPairs, benchEnd_processPotentialIntersectedPairs;
    TimePoint benchBegin_addUnintersectedTriangles, benchEnd_addUnintersectedTriangles;
    TimePoint benchBegin_reTriangulate, benchEnd_reTriangulate;
    TimePoint benchBegin_buildPolygonsFromEdges, benchEnd_buildPolygonsFromEdges;
    TimePoint benchBegin_buildFaceGroups, benchEnd_buildFaceGroups;
    TimePoint benchBegin_decideGroupSide, benchEnd_decideGroupSide;

    // front-end results of the last combine() (for tests and diagnostics)
    size_t candidatePairCount() const { return m_candidateCount; }
    size_t intersectingPairCount() const { return m_hitPairs.size() / 2; }

private:
    // per intersected triangle: welded local points + the relations between them, as the GPU builds
    // them (sb_isect_contexts = the reference's IntersectedContext, src/solidboolean.cpp:296-339)
    struct CutTriangle {
        std::vector<Vector3> points;
        std::vector<int> pointEdges; // per point: edge of this triangle it lies on (sb_isect_hit_edges), -1 = interior / unknown
        std::unordered_map<size_t, std::unordered_set<size_t>> neighbors; // indices are 3 + local point
    };
    // half-edge (from << 32 | to) -> triangle.  The uncut triangles' entries arrive from the GPU
    // as one array sorted by key (sb_uncut_half_edges: lookup = binary search); the pieces of
    // retriangulated faces are added on the host behind it.
    class HalfEdgeMap
    {
    public:
        void adopt(std::vector<uint64_t> &&sortedKeys, std::vector<uint32_t> &&owners);
        bool insert(uint64_t key, size_t triangle); // false: the half-edge already exists
        bool find(uint64_t key, size_t &triangle) const;
    private:
        std::vector<uint64_t> m_keys;
        std::vector<uint32_t> m_owners;
        std::unordered_map<uint64_t, size_t> m_added;
    };
    typedef std::unordered_map<size_t, std::unordered_set<size_t>> EdgeGraph;
    // The uncut triangles of one mesh as the GPU hands them over (sb_uncut_*): new triangles
    // [first, first + count), for each its face group among the uncut triangles (label = lowest
    // triangle of the group) and the triangle across each edge (-1: a cut face lies there).
    struct UncutTopology {
        size_t first = 0, count = 0;
        bool grouped = false; // labels / adjacency valid (false after a repeated half-edge)
        std::vector<uint32_t> label;
        std::vector<int32_t> adjacency;
        void *device = nullptr; // sb_uncut kept for the device flood (sb_uncut_face_groups); released by combine()
    };

    static uint64_t halfEdgeKey(size_t from, size_t to) { return ((uint64_t)from << 32) | (uint64_t)to; }
    size_t weldPoint(const Vector3 &p);
    bool appendTriangle(size_t a, size_t b, size_t c, HalfEdgeMap &halfEdges);
    bool copyUncutTriangles(const void *isect, int which, size_t vertexOffset, HalfEdgeMap &halfEdges, UncutTopology &topology);
    bool retriangulateCutTriangles(const std::map<size_t, CutTriangle> &cuts, const SolidMesh *mesh, size_t vertexOffset,
        HalfEdgeMap &halfEdges, EdgeGraph &loopEdges);
    bool traceLoops(const EdgeGraph &edges, std::vector<std::vector<size_t>> &loops);
    bool deviceFaceGroups(const std::vector<std::vector<size_t>> &loops, const UncutTopology &uncut, size_t pieceBegin, size_t pieceEnd,
        std::vector<std::vector<size_t>> &groups);
    void growFaceGroups(const std::vector<std::vector<size_t>> &loops, const HalfEdgeMap &halfEdges, const UncutTopology &uncut,
        size_t firstTriangle, size_t triangleCount, std::vector<std::vector<size_t>> &groups);
    bool classifyGroups(const std::vector<std::vector<size_t>> &groups, const SolidMesh *against, std::vector<bool> &inside);

    const SolidMesh *m_firstMesh = nullptr;
    const SolidMesh *m_secondMesh = nullptr;
    size_t m_candidateCount = 0;
    std::vector<uint32_t> m_hitPairs;  // 2 per intersecting pair, sorted by (first, second)
    std::vector<double> m_hitSegments; // 6 per intersecting pair
    std::vector<Vector3> m_newVertices;
    std::vector<std::array<size_t, 3>> m_newTriangles; // flat triples; fetch* turns them into the reference's vectors
    std::map<PositionKey, size_t> m_weldMap;
    std::vector<std::vector<size_t>> m_firstGroups, m_secondGroups;
    std::vector<bool> m_firstGroupInside, m_secondGroupInside;
};

#endif
