// SolidMesh with the reference's public interface (reference src/solidmesh.h:29-74),
// backed by the B200 front end: prepare() uploads the geometry and builds the
// device-side acceleration structures through the C ABI (sb_mesh_create).
// Ownership as in the reference: vertices / triangles are borrowed and must
// outlive the mesh; normals and the device mesh are owned and freed in the dtor.
#ifndef SOLID_MESH_H
#define SOLID_MESH_H
#include <cstddef>
#include <vector>
#include "vector3.h"

struct sb_mesh;
struct sb_context;

// The reference exposes its tree / box types through two accessors; here they are
// opaque handles to the device structures.
typedef sb_mesh AxisAlignedBoudingBoxTree;

class SolidMesh
{
public:
    SolidMesh() = default;
    ~SolidMesh();
    SolidMesh(const SolidMesh &) = delete;
    SolidMesh &operator=(const SolidMesh &) = delete;

    void setVertices(const std::vector<Vector3> *vertices) { m_vertices = vertices; }
    void setTriangles(const std::vector<std::vector<size_t>> *triangles) { m_triangles = triangles; }
    const std::vector<Vector3> *vertices() const { return m_vertices; }
    const std::vector<std::vector<size_t>> *triangles() const { return m_triangles; }
    const std::vector<Vector3> *triangleNormals() const { return m_triangleNormals; }
    const AxisAlignedBoudingBoxTree *axisAlignedBoundingBoxTree() const { return m_deviceMesh; }
    // per-triangle boxes as 6 doubles (lower xyz, upper xyz), fetched on demand
    std::vector<double> triangleAxisAlignedBoundingBoxes() const;

    // normals, boxes and the LBVH / ray grids on the device; a silent no-op
    // without triangles, like the reference (src/solidmesh.cpp:44-45)
    void prepare();

    sb_mesh *deviceMesh() const { return m_deviceMesh; }
    // one context (device + stream) per host thread, created on first use
    static sb_context *sharedContext();

private:
    const std::vector<Vector3> *m_vertices = nullptr;
    const std::vector<std::vector<size_t>> *m_triangles = nullptr;
    std::vector<Vector3> *m_triangleNormals = nullptr;
    sb_mesh *m_deviceMesh = nullptr;
};

#endif
