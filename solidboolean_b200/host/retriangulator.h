// Cuts one mesh triangle along the intersection segments that cross it and
// re-triangulates the pieces (the job of the reference's ReTriangulator,
// src/retriangulator.{h,cpp}; same public interface, implementation written for
// this project on top of earclip.h).  Runs on the host, as in the reference.
//
// Points are numbered 0..2 = the triangle's corners, 3.. = the intersection points
// handed to setEdges(); the neighbour map links those points into open polylines
// (ending on the triangle's boundary) and closed loops (strictly inside).
#ifndef RE_TRIANGULATOR_H
#define RE_TRIANGULATOR_H
#include <array>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include "vector3.h"

class ReTriangulator
{
public:
    ReTriangulator(const std::vector<Vector3> &trianglePoints, const Vector3 &normal);
    // pointEdges (optional, one per point): the edge of THIS triangle the point lies on (0..2: edge k joins corners k and
    // k + 1), as the predicate reported it (sb_isect_hit_edges), or -1 = not on the boundary / unknown.  A polyline end with
    // a known edge is attached to it directly; the others fall back to the geometric test.
    void setEdges(const std::vector<Vector3> &points,
        const std::unordered_map<size_t, std::unordered_set<size_t>> *neighborMapFrom3, const std::vector<int> *pointEdges = nullptr);
    bool reTriangulate();
    const std::vector<std::vector<size_t>> &polygons() const { return m_polygons; }
    const std::vector<std::vector<size_t>> &triangles() const { return m_triangles; }
    // regions whose triangulation came out short (see triangulateRegions); 0 on well-formed input
    size_t incompleteRegions() const { return m_incompleteRegions; }

private:
    typedef std::array<double, 2> P2;
    P2 project(const Vector3 &p) const;
    bool collectChains();
    bool splitBoundaryRing();
    void triangulateRegions();
    bool pointInRing(const P2 &p, const std::vector<size_t> &ring) const;

    Vector3 m_origin, m_axisU, m_axisV;
    std::vector<P2> m_points;
    std::vector<std::vector<size_t>> m_adjacency; // per point, sorted neighbours
    std::vector<int> m_pointEdge;                 // per point (corners included: -1), see setEdges
    std::vector<std::vector<size_t>> m_polylines; // open chains, endpoints on the boundary
    std::vector<std::vector<size_t>> m_loops;     // closed chains
    std::vector<std::vector<size_t>> m_polygons;  // boundary ring split by the polylines
    std::vector<std::vector<size_t>> m_triangles;
    size_t m_incompleteRegions = 0;
};

#endif
