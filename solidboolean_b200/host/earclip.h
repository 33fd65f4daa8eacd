// Ear-clipping triangulation of a simple polygon with holes, written for this
// project (the reference uses mapbox earcut, thirdparty/earcut.hpp, which is not
// vendored here).  Small inputs only: the polygons are the pieces a single mesh
// triangle is cut into by its intersection segments, a handful of vertices each,
// so the O(n^2) textbook algorithm is the right tool.
//
// Input: rings[0] = outer boundary, rings[1..] = holes, as 2-D points.  Vertices
// are numbered consecutively over the rings in the order given (earcut's
// convention).  Orientation of the rings does not matter.  Output: index triples,
// counter-clockwise.  Every input vertex of a non-degenerate ring is used; no
// triangle has another vertex strictly inside one of its edges' interiors unless
// the input is degenerate (so shared boundaries stay free of T-junctions).
#ifndef SB_HOST_EARCLIP_H
#define SB_HOST_EARCLIP_H
#include <array>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <vector>

namespace earclip
{

typedef std::array<double, 2> Point;

inline double cross(const Point &o, const Point &a, const Point &b)
{
    return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0]);
}

inline double ringArea(const std::vector<Point> &p, const std::vector<size_t> &ring)
{
    double a = 0.0;
    for (size_t i = 0, n = ring.size(); i < n; ++i) {
        const Point &u = p[ring[i]], &v = p[ring[(i + 1) % n]];
        a += u[0] * v[1] - v[0] * u[1];
    }
    return 0.5 * a;
}

// proper or touching intersection of segments ab and cd, ignoring shared endpoints
inline bool segmentsCross(const Point &a, const Point &b, const Point &c, const Point &d)
{
    auto same = [](const Point &x, const Point &y) { return x[0] == y[0] && x[1] == y[1]; };
    if (same(a, c) || same(a, d) || same(b, c) || same(b, d))
        return false;
    double d1 = cross(c, d, a), d2 = cross(c, d, b), d3 = cross(a, b, c), d4 = cross(a, b, d);
    if (((d1 > 0 && d2 < 0) || (d1 < 0 && d2 > 0)) && ((d3 > 0 && d4 < 0) || (d3 < 0 && d4 > 0)))
        return true;
    auto onSeg = [](const Point &p, const Point &q, const Point &r) {
        return std::fmin(p[0], q[0]) <= r[0] && r[0] <= std::fmax(p[0], q[0]) && std::fmin(p[1], q[1]) <= r[1] &&
            r[1] <= std::fmax(p[1], q[1]);
    };
    if (d1 == 0 && onSeg(c, d, a)) return true;
    if (d2 == 0 && onSeg(c, d, b)) return true;
    if (d3 == 0 && onSeg(a, b, c)) return true;
    if (d4 == 0 && onSeg(a, b, d)) return true;
    return false;
}

// p inside or on the boundary of CCW triangle abc
inline bool inTriangle(const Point &a, const Point &b, const Point &c, const Point &p)
{
    return cross(a, b, p) >= 0 && cross(b, c, p) >= 0 && cross(c, a, p) >= 0;
}

inline std::vector<size_t> triangulate(const std::vector<std::vector<Point>> &rings)
{
    std::vector<size_t> out;
    if (rings.empty() || rings[0].size() < 3)
        return out;
    std::vector<Point> pts;
    std::vector<std::vector<size_t>> idx(rings.size());
    for (size_t r = 0; r < rings.size(); ++r)
        for (const Point &q : rings[r]) {
            idx[r].push_back(pts.size());
            pts.push_back(q);
        }
    // outer ring counter-clockwise, holes clockwise
    std::vector<size_t> poly = idx[0];
    if (ringArea(pts, poly) < 0)
        poly.assign(idx[0].rbegin(), idx[0].rend());
    std::vector<std::vector<size_t>> holes;
    for (size_t r = 1; r < idx.size(); ++r) {
        if (idx[r].size() < 3)
            continue;
        std::vector<size_t> h = idx[r];
        if (ringArea(pts, h) > 0)
            h.assign(idx[r].rbegin(), idx[r].rend());
        holes.push_back(h);
    }
    // merge holes into the outer ring through bridges, rightmost hole first
    auto maxX = [&](const std::vector<size_t> &h) {
        size_t best = 0;
        for (size_t i = 1; i < h.size(); ++i)
            if (pts[h[i]][0] > pts[h[best]][0])
                best = i;
        return best;
    };
    std::vector<size_t> order(holes.size());
    for (size_t i = 0; i < order.size(); ++i)
        order[i] = i;
    for (size_t i = 0; i < order.size(); ++i)
        for (size_t j = i + 1; j < order.size(); ++j)
            if (pts[holes[order[j]][maxX(holes[order[j]])]][0] > pts[holes[order[i]][maxX(holes[order[i]])]][0])
                std::swap(order[i], order[j]);
    for (size_t oi = 0; oi < order.size(); ++oi) {
        const std::vector<size_t> &h = holes[order[oi]];
        size_t hm = maxX(h);
        const Point &m = pts[h[hm]];
        // visible outer vertex closest to m: the bridge must not cross any edge of
        // the current outer ring or of any hole that is still unmerged
        double bestD = -1.0;
        size_t bestK = poly.size();
        for (size_t k = 0; k < poly.size(); ++k) {
            const Point &q = pts[poly[k]];
            double d = (q[0] - m[0]) * (q[0] - m[0]) + (q[1] - m[1]) * (q[1] - m[1]);
            if (bestD >= 0 && d >= bestD)
                continue;
            bool blocked = false;
            for (size_t e = 0; e < poly.size() && !blocked; ++e)
                blocked = segmentsCross(m, q, pts[poly[e]], pts[poly[(e + 1) % poly.size()]]);
            for (size_t oj = oi; oj < order.size() && !blocked; ++oj) {
                const std::vector<size_t> &g = holes[order[oj]];
                for (size_t e = 0; e < g.size() && !blocked; ++e)
                    blocked = segmentsCross(m, q, pts[g[e]], pts[g[(e + 1) % g.size()]]);
            }
            // the bridge has to leave the outer ring towards its interior
            if (!blocked) {
                const Point &prev = pts[poly[(k + poly.size() - 1) % poly.size()]];
                const Point &next = pts[poly[(k + 1) % poly.size()]];
                bool convex = cross(prev, q, next) > 0;
                bool inside = convex ? (cross(prev, q, m) >= 0 && cross(q, next, m) >= 0)
                                     : !(cross(prev, q, m) < 0 && cross(q, next, m) < 0);
                blocked = !inside;
            }
            if (!blocked) {
                bestD = d;
                bestK = k;
            }
        }
        if (bestK == poly.size())
            continue; // no visible vertex (degenerate input): leave the hole out
        std::vector<size_t> merged;
        merged.reserve(poly.size() + h.size() + 2);
        for (size_t k = 0; k <= bestK; ++k)
            merged.push_back(poly[k]);
        for (size_t k = 0; k <= h.size(); ++k)
            merged.push_back(h[(hm + k) % h.size()]);
        for (size_t k = bestK; k < poly.size(); ++k)
            merged.push_back(poly[k]);
        poly.swap(merged);
    }
#ifdef EARCLIP_DEBUG
    const size_t expectTriangles = poly.size() - 2;
    const std::vector<size_t> poly0 = poly;
#endif
    // clip ears
    auto samePoint = [&](size_t a, size_t b) { return pts[a][0] == pts[b][0] && pts[a][1] == pts[b][1]; };
    size_t guard = 0;
    const size_t guardLimit = 4 * poly.size() + 16;
    while (poly.size() > 3 && guard < guardLimit) {
        size_t n = poly.size();
        bool clipped = false;
        // two passes: strictly convex empty ears first, then (degenerate input) any
        // non-reflex corner, so the loop always terminates
        for (int relaxed = 0; relaxed < 2 && !clipped; ++relaxed) {
            double bestQuality = -1.0;
            size_t bestI = n;
            for (size_t i = 0; i < n; ++i) {
                size_t ia = poly[(i + n - 1) % n], ib = poly[i], ic = poly[(i + 1) % n];
                const Point &a = pts[ia], &b = pts[ib], &c = pts[ic];
                double area2 = cross(a, b, c);
                if (relaxed ? area2 < 0 : area2 <= 0)
                    continue;
                bool empty = true;
                if (!relaxed)
                    for (size_t k = 0; k < n && empty; ++k) {
                        size_t ik = poly[k];
                        if (ik == ia || ik == ib || ik == ic || samePoint(ik, ia) || samePoint(ik, ib) || samePoint(ik, ic))
                            continue;
                        empty = !inTriangle(a, b, c, pts[ik]);
                    }
                if (!empty)
                    continue;
                // prefer well-shaped ears: area over longest-edge squared
                double l1 = (b[0] - a[0]) * (b[0] - a[0]) + (b[1] - a[1]) * (b[1] - a[1]);
                double l2 = (c[0] - b[0]) * (c[0] - b[0]) + (c[1] - b[1]) * (c[1] - b[1]);
                double l3 = (a[0] - c[0]) * (a[0] - c[0]) + (a[1] - c[1]) * (a[1] - c[1]);
                double q = area2 / std::fmax(std::fmax(l1, l2), std::fmax(l3, 1e-300));
                if (q > bestQuality) {
                    bestQuality = q;
                    bestI = i;
                }
            }
            if (bestI < n) {
                size_t ia = poly[(bestI + n - 1) % n], ib = poly[bestI], ic = poly[(bestI + 1) % n];
#ifdef EARCLIP_DEBUG
                if (relaxed) {
                    fprintf(stderr, "earclip relaxed clip at %zu of %zu, area2=%g\n  poly:", bestI, n, cross(pts[ia], pts[ib], pts[ic]));
                    for (size_t k = 0; k < n; ++k) fprintf(stderr, " (%.17g,%.17g)#%zu", pts[poly[k]][0], pts[poly[k]][1], poly[k]);
                    fprintf(stderr, "\n");
                }
#endif
                // a forced zero-area ear is still emitted: dropping its middle vertex
                // silently would leave a T-junction on the neighbouring triangles
                if (cross(pts[ia], pts[ib], pts[ic]) >= 0) {
                    out.push_back(ia);
                    out.push_back(ib);
                    out.push_back(ic);
                }
                poly.erase(poly.begin() + (std::ptrdiff_t)bestI);
                clipped = true;
            }
        }
        if (!clipped)
            break;
        ++guard;
    }
    if (poly.size() == 3 && cross(pts[poly[0]], pts[poly[1]], pts[poly[2]]) >= 0) {
        out.push_back(poly[0]);
        out.push_back(poly[1]);
        out.push_back(poly[2]);
    }
#ifdef EARCLIP_DEBUG
    if (out.size() / 3 != expectTriangles) {
        fprintf(stderr, "earclip: %zu triangles, expected %zu, left %zu (last area2 %g)\n  poly:", out.size() / 3, expectTriangles, poly.size(),
            poly.size() == 3 ? cross(pts[poly[0]], pts[poly[1]], pts[poly[2]]) : -1.0);
        for (size_t k = 0; k < poly0.size(); ++k) fprintf(stderr, " (%.17g,%.17g)#%zu", pts[poly0[k]][0], pts[poly0[k]][1], poly0[k]);
        fprintf(stderr, "\n");
    }
#endif
    return out;
}

} // namespace earclip

#endif
