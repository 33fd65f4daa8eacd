#include "solidboolean.h"
#include "positionkey.h"
#include <algorithm>
#include <cstdlib>
#include <string>
#include <deque>
#include <iostream>
#include "retriangulator.h"
#include "solidboolean_b200.h"

namespace
{
inline SolidBoolean::TimePoint now() { return std::chrono::high_resolution_clock::now(); }
}

SolidBoolean::SolidBoolean(const SolidMesh *firstMesh, const SolidMesh *secondMesh)
    : m_firstMesh(firstMesh), m_secondMesh(secondMesh)
{
}

SolidBoolean::~SolidBoolean() {}

const std::vector<Vector3> &SolidBoolean::resultVertices() { return m_newVertices; }

size_t SolidBoolean::weldPoint(const Vector3 &p)
{
    auto ins = m_weldMap.insert({PositionKey(p), m_newVertices.size()});
    if (ins.second)
        m_newVertices.push_back(p);
    return ins.first->second;
}

void SolidBoolean::HalfEdgeMap::adopt(std::vector<uint64_t> &&sortedKeys, std::vector<uint32_t> &&owners)
{
    m_keys = std::move(sortedKeys);
    m_owners = std::move(owners);
    m_added.clear();
}

bool SolidBoolean::HalfEdgeMap::find(uint64_t key, size_t &triangle) const
{
    auto it = std::lower_bound(m_keys.begin(), m_keys.end(), key);
    if (it != m_keys.end() && *it == key) {
        triangle = m_owners[it - m_keys.begin()];
        return true;
    }
    auto added = m_added.find(key);
    if (added == m_added.end())
        return false;
    triangle = added->second;
    return true;
}

bool SolidBoolean::HalfEdgeMap::insert(uint64_t key, size_t triangle)
{
    size_t existing;
    if (find(key, existing))
        return false;
    m_added.insert({key, triangle});
    return true;
}

bool SolidBoolean::appendTriangle(size_t a, size_t b, size_t c, HalfEdgeMap &halfEdges)
{
    size_t index = m_newTriangles.size();
    m_newTriangles.push_back({{a, b, c}});
    bool ok = true;
    ok &= halfEdges.insert(halfEdgeKey(a, b), index);
    ok &= halfEdges.insert(halfEdgeKey(b, c), index);
    ok &= halfEdges.insert(halfEdgeKey(c, a), index);
    return ok;
}

// Triangles the intersection does not touch are taken over as they are (reference
// addUnintersectedTriangles, src/solidboolean.cpp:250-286).  GPU: sb_isect_uncut compacts
// them, sorts their half-edge keys and reports a repeated half-edge exactly where the
// reference's map insert would have refused it.
bool SolidBoolean::copyUncutTriangles(const void *isect, int which, size_t vertexOffset, HalfEdgeMap &halfEdges,
    UncutTopology &topology)
{
    sb_uncut *uncut = nullptr;
    if (sb_isect_uncut(static_cast<const sb_isect *>(isect), which, vertexOffset, m_newTriangles.size(), &uncut) != SB_OK) {
        std::cout << "addUnintersectedTriangles failed: " << sb_last_error() << std::endl;
        return false;
    }
    size_t triangleCount = 0, halfEdgeCount = 0;
    int ok = 0;
    sb_uncut_counts(uncut, &triangleCount, &halfEdgeCount, &ok);
    std::vector<uint32_t> triples(3 * triangleCount), owners(halfEdgeCount);
    std::vector<uint64_t> keys(halfEdgeCount);
    int rc = sb_uncut_triangles(uncut, nullptr, triples.data());
    if (rc == SB_OK)
        rc = sb_uncut_half_edges(uncut, keys.data(), owners.data());
    topology.first = m_newTriangles.size();
    topology.count = triangleCount;
    topology.grouped = false;
    // SB_HOST_FLOOD=legacy: triangle-by-triangle flood through the half-edge map, as the reference
    // does it (comparison / debugging)
    // SB_HOST_FLOOD=host: the flood over the retriangulated pieces on the host (round 1), the uncut components from the GPU
    static const bool legacyFlood = [] { const char *e = std::getenv("SB_HOST_FLOOD"); return e && std::string(e) == "legacy"; }();
    static const bool hostFlood = [] { const char *e = std::getenv("SB_HOST_FLOOD"); return e && std::string(e) == "host"; }();
    if (rc == SB_OK && ok && !legacyFlood && !hostFlood) {
        // the whole flood runs on the GPU once the pieces exist (deviceFaceGroups): the object stays alive until then
        topology.device = uncut;
        topology.grouped = true;
        uncut = nullptr;
    } else if (rc == SB_OK && ok && triangleCount && !legacyFlood) {
        // face groups + neighbours of the uncut triangles, computed on the GPU
        topology.label.resize(triangleCount);
        topology.adjacency.resize(3 * triangleCount);
        rc = sb_uncut_components(uncut, topology.label.data(), nullptr);
        if (rc == SB_OK)
            rc = sb_uncut_adjacency(uncut, topology.adjacency.data());
        topology.grouped = rc == SB_OK;
    }
    if (uncut)
        sb_uncut_destroy(uncut);
    if (rc != SB_OK) {
        std::cout << "addUnintersectedTriangles failed: " << sb_last_error() << std::endl;
        return false;
    }
    m_newTriangles.reserve(m_newTriangles.size() + triangleCount);
    for (size_t i = 0; i < triangleCount; ++i)
        m_newTriangles.push_back({{triples[3 * i], triples[3 * i + 1], triples[3 * i + 2]}});
    if (!ok && triangleCount) {
        const auto &last = m_newTriangles.back();
        std::cout << "Found repeated halfedge:" << last[0] << "," << last[1] << std::endl;
    }
    halfEdges.adopt(std::move(keys), std::move(owners));
    return ok != 0;
}

// reference reTriangulate lambda, src/solidboolean.cpp:352-407
bool SolidBoolean::retriangulateCutTriangles(const std::map<size_t, CutTriangle> &cuts, const SolidMesh *mesh,
    size_t vertexOffset, HalfEdgeMap &halfEdges, EdgeGraph &loopEdges)
{
    const auto &vertices = *mesh->vertices();
    const auto &triangles = *mesh->triangles();
    const auto &normals = *mesh->triangleNormals();
    for (const auto &it : cuts) {
        const auto &t = triangles[it.first];
        ReTriangulator splitter({vertices[t[0]], vertices[t[1]], vertices[t[2]]}, normals[it.first]);
        splitter.setEdges(it.second.points, &it.second.neighbors, it.second.pointEdges.empty() ? nullptr : &it.second.pointEdges);
        if (!splitter.reTriangulate()) {
            std::cout << "Retriangle failed" << std::endl;
            return false;
        }
        if (splitter.incompleteRegions()) // reported the reference's way (a message, no exception); the result may have a hole there
            std::cout << "Retriangle incomplete: triangle " << it.first << ", " << splitter.incompleteRegions() << " region(s)" << std::endl;
        std::vector<size_t> global = {t[0] + vertexOffset, t[1] + vertexOffset, t[2] + vertexOffset};
        for (const Vector3 &p : it.second.points)
            global.push_back(weldPoint(p));
        for (const auto &piece : splitter.triangles()) {
            size_t a = global[piece[0]], b = global[piece[1]], c = global[piece[2]];
            if (a == b || b == c || c == a)
                continue; // collapsed by welding
            if (!appendTriangle(a, b, c, halfEdges))
                std::cout << "Found repeated halfedge:" << a << "," << b << std::endl;
        }
        for (const auto &nb : it.second.neighbors)
            for (size_t other : nb.second) {
                size_t from = global[nb.first], to = global[other];
                if (from == to)
                    continue;
                loopEdges[from].insert(to);
                loopEdges[to].insert(from);
            }
    }
    return true;
}

// the welded intersection segments form closed curves
// (reference buildPolygonsFromEdges, src/solidboolean.cpp:124-165)
bool SolidBoolean::traceLoops(const EdgeGraph &edges, std::vector<std::vector<size_t>> &loops)
{
    std::vector<size_t> nodes;
    nodes.reserve(edges.size());
    for (const auto &e : edges)
        nodes.push_back(e.first);
    std::sort(nodes.begin(), nodes.end());
    std::unordered_set<size_t> seen;
    for (size_t start : nodes) {
        if (seen.count(start))
            continue;
        std::vector<size_t> loop;
        size_t cur = start;
        while (true) {
            seen.insert(cur);
            loop.push_back(cur);
            auto it = edges.find(cur);
            if (it == edges.end())
                break;
            std::vector<size_t> next(it->second.begin(), it->second.end());
            std::sort(next.begin(), next.end());
            size_t chosen = cur;
            for (size_t n : next)
                if (!seen.count(n)) {
                    chosen = n;
                    break;
                }
            if (chosen == cur)
                break;
            cur = chosen;
        }
        if (loop.size() <= 2) {
            std::cout << "buildPolygonsFromEdges failed, too short" << std::endl;
            return false;
        }
        auto last = edges.find(loop.back());
        if (last == edges.end() || !last->second.count(start)) {
            std::cout << "buildPolygonsFromEdges failed, could not form a ring" << std::endl;
            return false;
        }
        loops.push_back(loop);
    }
    return true;
}

// buildFaceGroups on the GPU (sb_uncut_face_groups): the uncut triangles of this mesh's side (kept on
// the device since addUnintersectedTriangles) and its retriangulated pieces m_newTriangles[pieceBegin, pieceEnd) are
// flooded together, the edges of the intersection loops are the fences.  Groups come back as labels (lowest node of the
// group) and are laid out in ascending label order; within a group the triangles ascend.
bool SolidBoolean::deviceFaceGroups(const std::vector<std::vector<size_t>> &loops, const UncutTopology &uncut, size_t pieceBegin,
    size_t pieceEnd, std::vector<std::vector<size_t>> &groups)
{
    const size_t pieceCount = pieceEnd - pieceBegin;
    std::vector<uint32_t> pieces(3 * pieceCount), fences;
    for (size_t i = 0; i < pieceCount; ++i)
        for (int k = 0; k < 3; ++k)
            pieces[3 * i + k] = (uint32_t)m_newTriangles[pieceBegin + i][k];
    for (const auto &loop : loops)
        for (size_t i = 0; i < loop.size(); ++i) {
            fences.push_back((uint32_t)loop[i]);
            fences.push_back((uint32_t)loop[(i + 1) % loop.size()]);
        }
    std::vector<uint32_t> labelUncut(uncut.count), labelPiece(pieceCount);
    size_t groupCount = 0;
    if (sb_uncut_face_groups(static_cast<const sb_uncut *>(uncut.device), pieces.data(), pieceCount, fences.data(), fences.size() / 2,
            labelUncut.data(), labelPiece.data(), &groupCount) != SB_OK) {
        std::cout << "buildFaceGroups on the device failed: " << sb_last_error() << std::endl;
        return false;
    }
    // label (a node id) -> group index in ascending label order; counting sort of the members
    const size_t nodeCount = uncut.count + pieceCount;
    std::vector<uint32_t> groupOf(nodeCount, 0xffffffffu);
    size_t next = 0;
    for (size_t n = 0; n < nodeCount; ++n) {
        const uint32_t l = n < uncut.count ? labelUncut[n] : labelPiece[n - uncut.count];
        if (l == n)
            groupOf[n] = (uint32_t)next++;
    }
    std::vector<size_t> sizes(next, 0);
    for (size_t n = 0; n < nodeCount; ++n)
        ++sizes[groupOf[n < uncut.count ? labelUncut[n] : labelPiece[n - uncut.count]]];
    groups.assign(next, std::vector<size_t>());
    for (size_t g = 0; g < next; ++g)
        groups[g].reserve(sizes[g]);
    for (size_t n = 0; n < nodeCount; ++n) {
        const uint32_t l = n < uncut.count ? labelUncut[n] : labelPiece[n - uncut.count];
        groups[groupOf[l]].push_back(n < uncut.count ? uncut.first + n : pieceBegin + (n - uncut.count));
    }
    return true;
}

// Flood-fill the triangles of one mesh into groups bounded by the intersection
// loops: loop k seeds group 2k on the side of its forward half-edges and group
// 2k+1 on the other side; triangles no loop reaches form further groups
// (reference buildFaceGroups, src/solidboolean.cpp:167-239).
//
// The uncut triangles arrive pre-grouped from the GPU (sb_uncut_components): no loop runs through
// them, so a fill that reaches one of them takes its whole component at once and carries on
// through the component's open edges (adjacency -1: a retriangulated face lies there).  The host
// flood proper only walks the retriangulated pieces.  Where two seeds reach the same region the
// reference splits it between their groups along the meeting front of its queue; here the first
// seed to arrive takes the component whole -- the two groups lie on the same side of every loop,
// so the triangles each operation keeps are the same.
void SolidBoolean::growFaceGroups(const std::vector<std::vector<size_t>> &loops, const HalfEdgeMap &halfEdges,
    const UncutTopology &uncut, size_t firstTriangle, size_t triangleCount, std::vector<std::vector<size_t>> &groups)
{
    // members of every uncut component, ascending (counting sort by label)
    const bool grouped = uncut.grouped && uncut.count > 0;
    std::vector<uint32_t> memberStart, members;
    std::vector<uint8_t> claimed;
    if (grouped) {
        memberStart.assign(uncut.count + 1, 0);
        for (size_t i = 0; i < uncut.count; ++i)
            ++memberStart[uncut.label[i] - uncut.first + 1];
        for (size_t i = 0; i < uncut.count; ++i)
            memberStart[i + 1] += memberStart[i];
        members.resize(uncut.count);
        std::vector<uint32_t> cursor(memberStart.begin(), memberStart.end() - 1);
        for (size_t i = 0; i < uncut.count; ++i)
            members[cursor[uncut.label[i] - uncut.first]++] = (uint32_t)i;
        claimed.assign(uncut.count, 0);
    }
    auto isGroupedUncut = [&](size_t t) { return grouped && t >= uncut.first && t < uncut.first + uncut.count; };

    std::unordered_map<uint64_t, size_t> fence; // half-edges a fill must not cross again
    std::deque<std::pair<size_t, size_t>> queue;
    size_t groupIndex = 0;
    for (const auto &loop : loops) {
        for (size_t i = 0; i < loop.size(); ++i) {
            size_t a = loop[i], b = loop[(i + 1) % loop.size()];
            for (int side = 0; side < 2; ++side) {
                uint64_t key = side == 0 ? halfEdgeKey(a, b) : halfEdgeKey(b, a);
                fence.insert({key, groupIndex + side});
                size_t owner;
                if (halfEdges.find(key, owner))
                    queue.push_back({owner, groupIndex + side});
            }
        }
        groupIndex += 2;
    }
    groups.assign(groupIndex, std::vector<size_t>());
    std::unordered_set<size_t> visited; // retriangulated pieces (and everything when the GPU groups are missing)
    auto drain = [&]() {
        while (!queue.empty()) {
            auto item = queue.front();
            queue.pop_front();
            if (isGroupedUncut(item.first)) {
                const size_t root = uncut.label[item.first - uncut.first] - uncut.first;
                if (claimed[root])
                    continue;
                claimed[root] = 1;
                auto &group = groups[item.second];
                group.reserve(group.size() + (memberStart[root + 1] - memberStart[root]));
                for (uint32_t m = memberStart[root]; m < memberStart[root + 1]; ++m) {
                    const size_t local = members[m], tri = uncut.first + local;
                    group.push_back(tri);
                    for (size_t i = 0; i < 3; ++i) {
                        if (uncut.adjacency[3 * local + i] >= 0)
                            continue; // the neighbour is in this component
                        const auto &t = m_newTriangles[tri];
                        size_t a = t[i], b = t[(i + 1) % 3];
                        if (!fence.insert({halfEdgeKey(a, b), item.second}).second)
                            continue;
                        size_t opposite;
                        if (halfEdges.find(halfEdgeKey(b, a), opposite))
                            queue.push_back({opposite, item.second});
                    }
                }
                continue;
            }
            if (!visited.insert(item.first).second)
                continue;
            groups[item.second].push_back(item.first);
            const auto &t = m_newTriangles[item.first];
            for (size_t i = 0; i < 3; ++i) {
                size_t a = t[i], b = t[(i + 1) % 3];
                if (!fence.insert({halfEdgeKey(a, b), item.second}).second)
                    continue;
                size_t opposite;
                if (halfEdges.find(halfEdgeKey(b, a), opposite))
                    queue.push_back({opposite, item.second});
            }
        }
    };
    drain();
    for (size_t t = firstTriangle; t < firstTriangle + triangleCount; ++t) {
        if (isGroupedUncut(t)) {
            if (claimed[uncut.label[t - uncut.first] - uncut.first]) // labels are the lowest member: seen first
                continue;
        } else if (visited.count(t))
            continue;
        groups.push_back(std::vector<size_t>());
        queue.push_back({t, groupIndex++});
        drain();
    }
}

// One representative point per group -- the centroid of its first triangle,
// (v0 + v1 + v2) / 3.0 -- classified on the GPU with the reference's three-ray
// majority vote (reference decideGroupSide, src/solidboolean.cpp:482-510).
bool SolidBoolean::classifyGroups(const std::vector<std::vector<size_t>> &groups, const SolidMesh *against,
    std::vector<bool> &inside)
{
    inside.assign(groups.size(), false);
    std::vector<double> points;
    std::vector<size_t> owner;
    for (size_t g = 0; g < groups.size(); ++g) {
        if (groups[g].empty())
            continue;
        const auto &t = m_newTriangles[groups[g][0]];
        Vector3 c = (m_newVertices[t[0]] + m_newVertices[t[1]] + m_newVertices[t[2]]) / 3.0;
        points.push_back(c.x());
        points.push_back(c.y());
        points.push_back(c.z());
        owner.push_back(g);
    }
    if (owner.empty())
        return true;
    std::vector<uint8_t> flags(owner.size(), 0);
    if (!against->deviceMesh() ||
        sb_classify(against->deviceMesh(), points.data(), owner.size(), flags.data(), nullptr) != SB_OK) {
        std::cout << "decideGroupSide failed: " << sb_last_error() << std::endl;
        return false;
    }
    for (size_t k = 0; k < owner.size(); ++k)
        inside[owner[k]] = flags[k] != 0;
    return true;
}

bool SolidBoolean::combine()
{
    m_newVertices.clear();
    m_newTriangles.clear();
    m_weldMap.clear();
    m_firstGroups.clear();
    m_secondGroups.clear();
    m_hitPairs.clear();
    m_hitSegments.clear();
    if (!m_firstMesh || !m_secondMesh || !m_firstMesh->deviceMesh() || !m_secondMesh->deviceMesh()) {
        std::cout << "combine failed: meshes are not prepared" << std::endl;
        return false;
    }

    // ---- GPU: broad phase + tri/tri predicate (replaces searchPotentialIntersectedPairs
    // and the predicate loop, reference src/solidboolean.cpp:292, 315-320) ----
    benchBegin_searchPotentialIntersectedPairs = now();
    sb_isect *isect = nullptr;
    if (sb_intersect(m_firstMesh->deviceMesh(), m_secondMesh->deviceMesh(), SB_ISECT_DEFAULT, &isect) != SB_OK) {
        std::cout << "combine failed: " << sb_last_error() << std::endl;
        return false;
    }
    benchEnd_searchPotentialIntersectedPairs = now();
    benchBegin_processPotentialIntersectedPairs = now();
    size_t hitCount = 0;
    sb_isect_counts(isect, &m_candidateCount, &hitCount);
    m_hitPairs.resize(2 * hitCount);
    m_hitSegments.resize(6 * hitCount);
    int rc = sb_isect_hits(isect, m_hitPairs.data(), m_hitSegments.data());
    if (rc != SB_OK) {
        sb_isect_destroy(isect);
        std::cout << "combine failed: " << sb_last_error() << std::endl;
        return false;
    }
    // which edge every segment end point lies on, straight from the predicate; per cut triangle and welded
    // position: edge 0..2 of THAT triangle, or interior.  (A stand-in ABI without tags leaves the map empty: geometric attach.)
    std::vector<uint8_t> hitTags(hitCount);
    std::map<std::pair<size_t, PositionKey>, int> edgeOfPoint[2];
    if (hitCount && sb_isect_hit_edges(isect, hitTags.data()) == SB_OK) {
        for (size_t h = 0; h < hitCount; ++h)
            for (int end = 0; end < 2; ++end) {
                const unsigned t = hitTags[h] >> (4 * end);
                const int edge = (int)(t & 3u), onSecond = (int)((t >> 2) & 1u);
                const PositionKey key(m_hitSegments[6 * h + 3 * end], m_hitSegments[6 * h + 3 * end + 1], m_hitSegments[6 * h + 3 * end + 2]);
                for (int which = 0; which < 2; ++which) {
                    auto slot = std::make_pair((size_t)m_hitPairs[2 * h + which], key);
                    const int mine = onSecond == which ? edge : -1;
                    auto it = edgeOfPoint[which].find(slot);
                    if (it == edgeOfPoint[which].end())
                        edgeOfPoint[which].insert({slot, mine});
                    else if (it->second < 0)
                        it->second = mine; // a welded twin that does lie on the boundary decides
                }
            }
    }
    // ---- GPU: the per-triangle contexts the reference builds in the body of its pair loop
    // (src/solidboolean.cpp:296-339): welded points in first-seen order + the relations between them
    std::map<size_t, CutTriangle> firstCuts, secondCuts; // ordered: deterministic output
    for (int which = 0; which < 2; ++which) {
        sb_cuts *contexts = nullptr;
        if (sb_isect_contexts(isect, which, &contexts) != SB_OK) {
            sb_isect_destroy(isect);
            std::cout << "combine failed: " << sb_last_error() << std::endl;
            return false;
        }
        size_t contextCount = 0, pointCount = 0, relationCount = 0;
        sb_cuts_counts(contexts, &contextCount, &pointCount, &relationCount);
        std::vector<uint32_t> triangle(contextCount), pointStart(contextCount + 1), relationStart(contextCount + 1),
            relations(2 * relationCount);
        std::vector<double> points(3 * pointCount);
        rc = sb_cuts_fetch(contexts, triangle.data(), pointStart.data(), points.data(), relationStart.data(), relations.data());
        sb_cuts_destroy(contexts);
        if (rc != SB_OK) {
            sb_isect_destroy(isect);
            std::cout << "combine failed: " << sb_last_error() << std::endl;
            return false;
        }
        auto &cuts = which == 0 ? firstCuts : secondCuts;
        for (size_t c = 0; c < contextCount; ++c) {
            CutTriangle &cut = cuts[triangle[c]];
            for (size_t p = pointStart[c]; p < pointStart[c + 1]; ++p) {
                cut.points.push_back(Vector3(points[3 * p], points[3 * p + 1], points[3 * p + 2]));
                if (!edgeOfPoint[which].empty()) {
                    auto it = edgeOfPoint[which].find(std::make_pair((size_t)triangle[c], PositionKey(points[3 * p], points[3 * p + 1], points[3 * p + 2])));
                    cut.pointEdges.push_back(it == edgeOfPoint[which].end() ? -1 : it->second);
                }
            }
            for (size_t r = relationStart[c]; r < relationStart[c + 1]; ++r) {
                cut.neighbors[relations[2 * r]].insert(relations[2 * r + 1]);
                cut.neighbors[relations[2 * r + 1]].insert(relations[2 * r]);
            }
        }
    }
    benchEnd_processPotentialIntersectedPairs = now();

    // ---- host topology, as in the reference ----
    benchBegin_addUnintersectedTriangles = now();
    const size_t firstVertexCount = m_firstMesh->vertices()->size();
    m_newVertices.reserve(firstVertexCount + m_secondMesh->vertices()->size());
    m_newVertices.insert(m_newVertices.end(), m_firstMesh->vertices()->begin(), m_firstMesh->vertices()->end());
    m_newVertices.insert(m_newVertices.end(), m_secondMesh->vertices()->begin(), m_secondMesh->vertices()->end());
    HalfEdgeMap firstHalfEdges, secondHalfEdges;
    UncutTopology firstUncut, secondUncut;
    size_t firstStart = m_newTriangles.size();
    if (!copyUncutTriangles(isect, 0, 0, firstHalfEdges, firstUncut))
        std::cout << "Add first mesh remaining triangles failed" << std::endl;
    size_t firstCount = m_newTriangles.size() - firstStart;
    size_t secondStart = m_newTriangles.size();
    if (!copyUncutTriangles(isect, 1, firstVertexCount, secondHalfEdges, secondUncut))
        std::cout << "Add second mesh remaining triangles failed" << std::endl;
    size_t secondCount = m_newTriangles.size() - secondStart;
    sb_isect_destroy(isect);
    benchEnd_addUnintersectedTriangles = now();

    // (the uncut objects stay on the device for the flood; released on every way out)
    struct UncutRelease {
        UncutTopology &a, &b;
        ~UncutRelease()
        {
            if (a.device) sb_uncut_destroy(static_cast<sb_uncut *>(a.device));
            if (b.device) sb_uncut_destroy(static_cast<sb_uncut *>(b.device));
            a.device = b.device = nullptr;
        }
    } uncutRelease{firstUncut, secondUncut};

    benchBegin_reTriangulate = now();
    EdgeGraph firstLoopEdges, secondLoopEdges;
    const size_t firstPieceBegin = m_newTriangles.size();
    if (!retriangulateCutTriangles(firstCuts, m_firstMesh, 0, firstHalfEdges, firstLoopEdges)) {
        std::cout << "Retriangulate first mesh failed" << std::endl;
        return false;
    }
    const size_t secondPieceBegin = m_newTriangles.size();
    if (!retriangulateCutTriangles(secondCuts, m_secondMesh, firstVertexCount, secondHalfEdges, secondLoopEdges)) {
        std::cout << "Retriangulate second mesh failed" << std::endl;
        return false;
    }
    const size_t secondPieceEnd = m_newTriangles.size();
    benchEnd_reTriangulate = now();

    benchBegin_buildPolygonsFromEdges = now();
    std::vector<std::vector<size_t>> loops;
    if (!traceLoops(firstLoopEdges, loops)) {
        std::cout << "Build polygons from edges failed" << std::endl;
        return false;
    }
    benchEnd_buildPolygonsFromEdges = now();

    benchBegin_buildFaceGroups = now();
    // host flood (the device one was refused or is switched off): the uncut components and their neighbours from the GPU
    // (round 1), the walk over the retriangulated pieces here
    auto hostFlood = [&](UncutTopology &u, const HalfEdgeMap &map, size_t start, size_t count, std::vector<std::vector<size_t>> &groups) {
        if (u.device && u.label.empty() && u.count) {
            u.label.resize(u.count);
            u.adjacency.resize(3 * u.count);
            const sb_uncut *d = static_cast<const sb_uncut *>(u.device);
            u.grouped = sb_uncut_components(d, u.label.data(), nullptr) == SB_OK && sb_uncut_adjacency(d, u.adjacency.data()) == SB_OK;
        } else if (u.device && !u.count) {
            u.grouped = false;
        }
        growFaceGroups(loops, map, u, start, count, groups);
    };
    if (!firstUncut.device || !deviceFaceGroups(loops, firstUncut, firstPieceBegin, secondPieceBegin, m_firstGroups))
        hostFlood(firstUncut, firstHalfEdges, firstStart, firstCount, m_firstGroups);
    if (!secondUncut.device || !deviceFaceGroups(loops, secondUncut, secondPieceBegin, secondPieceEnd, m_secondGroups))
        hostFlood(secondUncut, secondHalfEdges, secondStart, secondCount, m_secondGroups);
    benchEnd_buildFaceGroups = now();

    // ---- GPU: inside/outside of every group ----
    benchBegin_decideGroupSide = now();
    bool ok = classifyGroups(m_firstGroups, m_secondMesh, m_firstGroupInside);
    ok = classifyGroups(m_secondGroups, m_firstMesh, m_secondGroupInside) && ok;
    benchEnd_decideGroupSide = now();
    return ok;
}

// reference src/solidboolean.cpp:512-565
void SolidBoolean::fetchUnion(std::vector<std::vector<size_t>> &resultTriangles)
{
    for (size_t g = 0; g < m_firstGroups.size(); ++g)
        if (!m_firstGroupInside[g])
            for (size_t t : m_firstGroups[g])
                resultTriangles.push_back({m_newTriangles[t][0], m_newTriangles[t][1], m_newTriangles[t][2]});
    for (size_t g = 0; g < m_secondGroups.size(); ++g)
        if (!m_secondGroupInside[g])
            for (size_t t : m_secondGroups[g])
                resultTriangles.push_back({m_newTriangles[t][0], m_newTriangles[t][1], m_newTriangles[t][2]});
}

void SolidBoolean::fetchDiff(std::vector<std::vector<size_t>> &resultTriangles)
{
    for (size_t g = 0; g < m_firstGroups.size(); ++g)
        if (!m_firstGroupInside[g])
            for (size_t t : m_firstGroups[g])
                resultTriangles.push_back({m_newTriangles[t][0], m_newTriangles[t][1], m_newTriangles[t][2]});
    for (size_t g = 0; g < m_secondGroups.size(); ++g)
        if (m_secondGroupInside[g])
            for (size_t t : m_secondGroups[g]) { // the cavity wall faces inward: reverse the winding
                const auto &tri = m_newTriangles[t];
                resultTriangles.push_back({tri[2], tri[1], tri[0]});
            }
}

void SolidBoolean::fetchIntersect(std::vector<std::vector<size_t>> &resultTriangles)
{
    for (size_t g = 0; g < m_firstGroups.size(); ++g)
        if (m_firstGroupInside[g])
            for (size_t t : m_firstGroups[g])
                resultTriangles.push_back({m_newTriangles[t][0], m_newTriangles[t][1], m_newTriangles[t][2]});
    for (size_t g = 0; g < m_secondGroups.size(); ++g)
        if (m_secondGroupInside[g])
            for (size_t t : m_secondGroups[g])
                resultTriangles.push_back({m_newTriangles[t][0], m_newTriangles[t][1], m_newTriangles[t][2]});
}
