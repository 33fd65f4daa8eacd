#include "solidmesh.h"
#include <cstdint>
#include <iostream>
#include "solidboolean_b200.h"

// One context per process.  The reference allows concurrent SolidBoolean objects over shared const
// SolidMesh from several threads (it has no mutable global state); here every call of the C ABI locks
// its context for the duration of the call (sb_capi.cu, DeviceGuard), so such callers are serialised on
// the GPU's single stream of work instead of racing on it.  Created once, thread-safely (C++11 static).
sb_context *SolidMesh::sharedContext()
{
    static sb_context *ctx = []() -> sb_context * {
        sb_context *c = nullptr;
        if (sb_context_create(0, &c) != SB_OK) {
            std::cout << "solidboolean_b200: " << sb_last_error() << std::endl;
            c = nullptr;
        }
        return c;
    }();
    return ctx;
}

SolidMesh::~SolidMesh()
{
    delete m_triangleNormals;
    if (m_deviceMesh)
        sb_mesh_destroy(m_deviceMesh);
}

void SolidMesh::prepare()
{
    if (nullptr == m_triangles || nullptr == m_vertices)
        return;
    sb_context *ctx = sharedContext();
    if (!ctx)
        return; // error already reported the reference's way (message on stdout)
    // the reference keeps every triangle in its own heap vector: flatten once
    const size_t count = m_triangles->size();
    std::vector<uint32_t> flat(3 * count);
    for (size_t i = 0; i < count; ++i) {
        const std::vector<size_t> &t = (*m_triangles)[i];
        flat[3 * i] = (uint32_t)t[0];
        flat[3 * i + 1] = (uint32_t)t[1];
        flat[3 * i + 2] = (uint32_t)t[2];
    }
    if (m_deviceMesh) {
        sb_mesh_destroy(m_deviceMesh);
        m_deviceMesh = nullptr;
    }
    const double *xyz = m_vertices->empty() ? nullptr : (*m_vertices)[0].constData();
    if (sb_mesh_create(ctx, xyz, m_vertices->size(), flat.data(), count, &m_deviceMesh) != SB_OK) {
        std::cout << "SolidMesh::prepare failed: " << sb_last_error() << std::endl;
        m_deviceMesh = nullptr;
        return;
    }
    delete m_triangleNormals;
    m_triangleNormals = new std::vector<Vector3>(count);
    if (count && sb_mesh_normals(m_deviceMesh, &(*m_triangleNormals)[0][0]) != SB_OK)
        std::cout << "SolidMesh::prepare failed: " << sb_last_error() << std::endl;
}

std::vector<double> SolidMesh::triangleAxisAlignedBoundingBoxes() const
{
    std::vector<double> boxes;
    if (m_deviceMesh && m_triangles) {
        boxes.resize(6 * m_triangles->size());
        if (!boxes.empty() && sb_mesh_triangle_boxes(m_deviceMesh, boxes.data()) != SB_OK)
            boxes.clear();
    }
    return boxes;
}
