// Quantised position used to weld intersection points (reference
// src/positionkey.cpp:25-54: truncation of x * 100000, lexicographic order).
#ifndef SB_HOST_POSITIONKEY_H
#define SB_HOST_POSITIONKEY_H
#include "vector3.h"

class PositionKey
{
public:
    explicit PositionKey(const Vector3 &v) : PositionKey(v.x(), v.y(), v.z()) {}
    PositionKey(double x, double y, double z)
        : m_x((long)(x * kFactor)), m_y((long)(y * kFactor)), m_z((long)(z * kFactor)) {}
    bool operator<(const PositionKey &o) const
    {
        if (m_x != o.m_x) return m_x < o.m_x;
        if (m_y != o.m_y) return m_y < o.m_y;
        return m_z < o.m_z;
    }
    bool operator==(const PositionKey &o) const { return m_x == o.m_x && m_y == o.m_y && m_z == o.m_z; }

private:
    static constexpr long kFactor = 100000;
    long m_x, m_y, m_z;
};

#endif
