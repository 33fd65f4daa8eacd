#include "retriangulator.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <iostream>
#include "earclip.h"

ReTriangulator::ReTriangulator(const std::vector<Vector3> &trianglePoints, const Vector3 &normal)
{
    // 2-D frame in the triangle's plane: u along the first edge, v = n x u, so the
    // triangle is counter-clockwise in (u, v) whenever `normal` is its own normal
    m_origin = trianglePoints[0];
    m_axisU = (trianglePoints[1] - trianglePoints[0]).normalized();
    m_axisV = Vector3::crossProduct(normal, m_axisU);
    for (const Vector3 &p : trianglePoints)
        m_points.push_back(project(p));
}

ReTriangulator::P2 ReTriangulator::project(const Vector3 &p) const
{
    Vector3 d = p - m_origin;
    return P2{Vector3::dotProduct(d, m_axisU), Vector3::dotProduct(d, m_axisV)};
}

void ReTriangulator::setEdges(const std::vector<Vector3> &points,
    const std::unordered_map<size_t, std::unordered_set<size_t>> *neighborMapFrom3, const std::vector<int> *pointEdges)
{
    for (const Vector3 &p : points)
        m_points.push_back(project(p));
    m_pointEdge.assign(m_points.size(), -1);
    if (pointEdges)
        for (size_t i = 0; i < pointEdges->size() && 3 + i < m_points.size(); ++i)
            m_pointEdge[3 + i] = (*pointEdges)[i];
    m_adjacency.assign(m_points.size(), std::vector<size_t>());
    if (neighborMapFrom3)
        for (const auto &it : *neighborMapFrom3) {
            if (it.first >= m_points.size())
                continue;
            for (size_t n : it.second)
                if (n < m_points.size())
                    m_adjacency[it.first].push_back(n);
        }
    for (auto &a : m_adjacency)
        std::sort(a.begin(), a.end()); // deterministic walks whatever the hash order was
}

// Decompose the segment graph into open polylines and closed loops.
bool ReTriangulator::collectChains()
{
    const size_t n = m_points.size();
    std::vector<char> used(n, 0);
    // A point of degree > 2 (PositionKey welding can merge the end points of neighbouring segments) does not
    // stop the reference: lookupPolylinesFromNeighborMap (src/retriangulator.cpp:48-95) starts at the degree-1
    // points, then at whatever is left, and every walk greedily follows the first unvisited neighbour.  Same here
    // (neighbours in ascending order instead of hash order).
    auto walk = [&](size_t start) {
        std::vector<size_t> chain;
        size_t prev = n, cur = start;
        while (true) {
            used[cur] = 1;
            chain.push_back(cur);
            size_t next = n;
            for (size_t nb : m_adjacency[cur])
                if (nb != prev && !used[nb]) {
                    next = nb;
                    break;
                }
            if (next == n)
                break;
            prev = cur;
            cur = next;
        }
        return chain;
    };
    for (size_t i = 3; i < n; ++i) // open chains start at their degree-1 ends
        if (!used[i] && m_adjacency[i].size() == 1) {
            std::vector<size_t> chain = walk(i);
            if (chain.size() >= 2)
                m_polylines.push_back(chain);
        }
    for (size_t i = 3; i < n; ++i) // what is left are cycles
        if (!used[i] && m_adjacency[i].size() >= 2) {
            std::vector<size_t> chain = walk(i);
            bool closed = chain.size() >= 3 &&
                std::find(m_adjacency[chain.back()].begin(), m_adjacency[chain.back()].end(), chain.front()) !=
                    m_adjacency[chain.back()].end();
            if (closed)
                m_loops.push_back(chain);
            else if (chain.size() >= 2)
                m_polylines.push_back(chain);
        }
    return true;
}

// Attach every polyline end to the triangle edge it lies on, order the attached
// points along the boundary, and walk the faces of the resulting planar map.
bool ReTriangulator::splitBoundaryRing()
{
    struct RingEntry {
        size_t point;
        int polyline; // -1: a corner
        bool front;   // this entry is the polyline's first point
        double t;     // position along its edge
    };
    double size2 = 0.0;
    for (int i = 0; i < 3; ++i) {
        const P2 &a = m_points[i], &b = m_points[(i + 1) % 3];
        size2 = std::max(size2, (b[0] - a[0]) * (b[0] - a[0]) + (b[1] - a[1]) * (b[1] - a[1]));
    }
    // A polyline end is either on an edge up to rounding (1e-6 of the triangle's size)
    // or was welded, inside this triangle, to a neighbouring intersection point by the
    // 1e-5 PositionKey grid (the segment between them collapsed): it can then sit up to
    // one key cell away from the edge.  The reference tests exact collinearity with an
    // absolute DBL_EPSILON (src/vector2.h:212-215) and gives up on such triangles.
    const double tolerance2 = std::max(1e-12 * size2, 2e-5 * 2e-5);
    std::vector<RingEntry> onEdge[3];
    auto attach = [&](size_t point, int polyline, bool front) {
        int bestEdge = -1;
        double bestD = 0.0, bestT = 0.0;
        const P2 &p = m_points[point];
        const int known = point < m_pointEdge.size() ? m_pointEdge[point] : -1;
        if (known >= 0 && known < 3) {
            // the predicate constructed this point ON edge `known` (base - alpha * edge vector): no tolerance involved
            const P2 &a = m_points[known], &b = m_points[(known + 1) % 3];
            double ex = b[0] - a[0], ey = b[1] - a[1], len2 = ex * ex + ey * ey;
            if (len2 > 0.0) {
                double t = ((p[0] - a[0]) * ex + (p[1] - a[1]) * ey) / len2;
                onEdge[known].push_back(RingEntry{point, polyline, front, std::min(1.0, std::max(0.0, t))});
                return true;
            }
        }
        for (int i = 0; i < 3; ++i) {
            const P2 &a = m_points[i], &b = m_points[(i + 1) % 3];
            double ex = b[0] - a[0], ey = b[1] - a[1], len2 = ex * ex + ey * ey;
            if (len2 <= 0.0)
                continue;
            double t = ((p[0] - a[0]) * ex + (p[1] - a[1]) * ey) / len2;
            double tc = std::min(1.0, std::max(0.0, t));
            double dx = p[0] - (a[0] + tc * ex), dy = p[1] - (a[1] + tc * ey), d = dx * dx + dy * dy;
            if (bestEdge < 0 || d < bestD) {
                bestEdge = i;
                bestD = d;
                bestT = tc;
            }
        }
        if (bestEdge < 0 || bestD > tolerance2) {
#ifdef RETRI_DEBUG
            fprintf(stderr, "attach fail: point %zu polyline %d (len %zu) front %d dist %g size %g ; npts %zu nlines %zu nloops %zu\n", point, polyline,
                m_polylines[polyline].size(), (int)front, std::sqrt(bestD), std::sqrt(size2), m_points.size(), m_polylines.size(), m_loops.size());
            for (size_t q = 0; q < m_points.size(); ++q) {
                fprintf(stderr, "   p%zu (%.12g, %.12g) adj:", q, m_points[q][0], m_points[q][1]);
                for (size_t nb : m_adjacency[q]) fprintf(stderr, " %zu", nb);
                fprintf(stderr, "\n");
            }
#endif
            return false;
        }
        onEdge[bestEdge].push_back(RingEntry{point, polyline, front, bestT});
        return true;
    };
    for (size_t k = 0; k < m_polylines.size(); ++k)
        if (!attach(m_polylines[k].front(), (int)k, true) || !attach(m_polylines[k].back(), (int)k, false)) {
            std::cout << "Attach point to triangle edge failed" << std::endl;
            return false;
        }
    std::vector<RingEntry> ring;
    for (int i = 0; i < 3; ++i) {
        ring.push_back(RingEntry{(size_t)i, -1, false, 0.0});
        std::stable_sort(onEdge[i].begin(), onEdge[i].end(), [](const RingEntry &a, const RingEntry &b) { return a.t < b.t; });
        ring.insert(ring.end(), onEdge[i].begin(), onEdge[i].end());
    }
    const size_t n = ring.size();
    // ring position of the other end of each polyline
    std::vector<size_t> frontPos(m_polylines.size(), n), backPos(m_polylines.size(), n);
    for (size_t i = 0; i < n; ++i)
        if (ring[i].polyline >= 0)
            (ring[i].front ? frontPos : backPos)[ring[i].polyline] = i;
    std::vector<char> visited(n, 0);
    std::vector<size_t> starts(1, 0);
    while (!starts.empty()) {
        size_t s = starts.back();
        starts.pop_back();
        if (visited[s])
            continue;
        std::vector<size_t> polygon;
        size_t pos = s, steps = 0;
        do {
            visited[pos] = 1;
            const RingEntry &e = ring[pos];
            if (e.polyline < 0) {
                polygon.push_back(e.point);
                pos = (pos + 1) % n;
            } else {
                const std::vector<size_t> &line = m_polylines[e.polyline];
                if (e.front)
                    polygon.insert(polygon.end(), line.begin(), line.end());
                else
                    polygon.insert(polygon.end(), line.rbegin(), line.rend());
                size_t other = e.front ? backPos[e.polyline] : frontPos[e.polyline];
                if (other == n) {
                    std::cerr << "linkTo failed" << std::endl;
                    return false;
                }
                starts.push_back((pos + 1) % n); // the face on the far side of this polyline
                pos = (other + 1) % n;
            }
            if (++steps > 4 * n + 8) {
                std::cout << "ReTriangulator: boundary walk did not close" << std::endl;
                return false;
            }
        } while (pos != s);
        if (polygon.size() >= 3)
            m_polygons.push_back(polygon);
    }
    return true;
}

bool ReTriangulator::pointInRing(const P2 &p, const std::vector<size_t> &ring) const
{
    bool inside = false;
    for (size_t i = 0, j = ring.size() - 1; i < ring.size(); j = i++) {
        const P2 &a = m_points[ring[i]], &b = m_points[ring[j]];
        if (((a[1] > p[1]) != (b[1] > p[1])) && (p[0] < (b[0] - a[0]) * (p[1] - a[1]) / (b[1] - a[1]) + a[0]))
            inside = !inside;
    }
    return inside;
}

void ReTriangulator::triangulateRegions()
{
    // nesting of the closed loops: parent = innermost loop containing it
    const size_t L = m_loops.size();
    std::vector<int> parent(L, -1);
    std::vector<size_t> depth(L, 0);
    for (size_t i = 0; i < L; ++i)
        for (size_t j = 0; j < L; ++j)
            if (i != j && pointInRing(m_points[m_loops[i][0]], m_loops[j]))
                ++depth[i];
    for (size_t i = 0; i < L; ++i) {
        int best = -1;
        for (size_t j = 0; j < L; ++j)
            if (i != j && pointInRing(m_points[m_loops[i][0]], m_loops[j]) && (best < 0 || depth[j] > depth[(size_t)best]))
                best = (int)j;
        parent[i] = best;
    }
    auto emit = [&](const std::vector<size_t> &outer, const std::vector<size_t> &holeLoops) {
        std::vector<std::vector<earclip::Point>> rings;
        std::vector<size_t> local;
        auto push = [&](const std::vector<size_t> &ring) {
            std::vector<earclip::Point> r;
            for (size_t p : ring) {
                r.push_back(m_points[p]);
                local.push_back(p);
            }
            rings.push_back(r);
        };
        push(outer);
        for (size_t h : holeLoops)
            push(m_loops[h]);
        std::vector<size_t> tri = earclip::triangulate(rings);
        // a simple polygon with h holes has n + 2h - 2 triangles; fewer = no ear found, a hole without a visible
        // bridge, or a degenerate rest (earcut.hpp hands back its partial output just as silently: counted here)
        if (tri.size() / 3 != local.size() + 2 * holeLoops.size() - 2)
            ++m_incompleteRegions;
        for (size_t i = 0; i + 2 < tri.size(); i += 3)
            m_triangles.push_back({local[tri[i]], local[tri[i + 1]], local[tri[i + 2]]});
    };
    for (const auto &polygon : m_polygons) {
        std::vector<size_t> holes;
        for (size_t i = 0; i < L; ++i)
            if (parent[i] < 0 && pointInRing(m_points[m_loops[i][0]], polygon))
                holes.push_back(i);
        emit(polygon, holes);
    }
    for (size_t i = 0; i < L; ++i) {
        std::vector<size_t> holes;
        for (size_t j = 0; j < L; ++j)
            if (parent[j] == (int)i)
                holes.push_back(j);
        emit(m_loops[i], holes);
    }
}

bool ReTriangulator::reTriangulate()
{
    if (!collectChains())
        return false;
    if (!splitBoundaryRing()) {
        std::cout << "Build polygons failed" << std::endl;
        return false;
    }
    triangulateRegions();
    return true;
}
