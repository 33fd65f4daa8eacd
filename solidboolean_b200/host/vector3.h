// Vector3 with the reference's public surface (reference src/vector3.h:43-311):
// AoS double[3] so that a std::vector<Vector3> IS the xyz buffer the C ABI takes.
// Written for this project; only the members the SolidMesh / SolidBoolean API and
// its callers (test/main.cpp) use are provided.
#ifndef SB_HOST_VECTOR3_H
#define SB_HOST_VECTOR3_H
#include <cmath>
#include <cstddef>
#include <iostream>
#include <string>
#include <vector>
#include "double.h"

class Vector3
{
public:
    Vector3() : m_data{0.0, 0.0, 0.0} {}
    Vector3(double x, double y, double z) : m_data{x, y, z} {}

    double &operator[](size_t i) { return m_data[i]; }
    const double &operator[](size_t i) const { return m_data[i]; }
    const double &x() const { return m_data[0]; }
    const double &y() const { return m_data[1]; }
    const double &z() const { return m_data[2]; }
    void setX(double v) { m_data[0] = v; }
    void setY(double v) { m_data[1] = v; }
    void setZ(double v) { m_data[2] = v; }
    void setData(double x, double y, double z) { m_data[0] = x; m_data[1] = y; m_data[2] = z; }
    const double *constData() const { return m_data; }

    double lengthSquared() const { return m_data[0] * m_data[0] + m_data[1] * m_data[1] + m_data[2] * m_data[2]; }
    double length() const { return std::sqrt(lengthSquared()); }
    Vector3 normalized() const
    {
        double len = length();
        if (Double::isZero(len))
            return Vector3();
        return Vector3(m_data[0] / len, m_data[1] / len, m_data[2] / len);
    }
    void normalize() { *this = normalized(); }
    bool isZero() const { return Double::isZero(m_data[0]) && Double::isZero(m_data[1]) && Double::isZero(m_data[2]); }

    static Vector3 crossProduct(const Vector3 &a, const Vector3 &b)
    {
        return Vector3(a.y() * b.z() - a.z() * b.y(), a.z() * b.x() - a.x() * b.z(), a.x() * b.y() - a.y() * b.x());
    }
    static double dotProduct(const Vector3 &a, const Vector3 &b) { return a.x() * b.x() + a.y() * b.y() + a.z() * b.z(); }
    // unit normal of triangle (a, b, c); zero vector when degenerate
    static Vector3 normal(const Vector3 &a, const Vector3 &b, const Vector3 &c)
    {
        Vector3 ab(b.x() - a.x(), b.y() - a.y(), b.z() - a.z());
        Vector3 ac(c.x() - a.x(), c.y() - a.y(), c.z() - a.z());
        return crossProduct(ab, ac).normalized();
    }

    Vector3 &operator+=(const Vector3 &o) { m_data[0] += o.x(); m_data[1] += o.y(); m_data[2] += o.z(); return *this; }
    Vector3 &operator-=(const Vector3 &o) { m_data[0] -= o.x(); m_data[1] -= o.y(); m_data[2] -= o.z(); return *this; }
    Vector3 &operator*=(double k) { m_data[0] *= k; m_data[1] *= k; m_data[2] *= k; return *this; }
    Vector3 &operator/=(double k) { m_data[0] /= k; m_data[1] /= k; m_data[2] /= k; return *this; }

private:
    double m_data[3];
};

inline Vector3 operator+(const Vector3 &a, const Vector3 &b) { return Vector3(a.x() + b.x(), a.y() + b.y(), a.z() + b.z()); }
inline Vector3 operator-(const Vector3 &a, const Vector3 &b) { return Vector3(a.x() - b.x(), a.y() - b.y(), a.z() - b.z()); }
inline Vector3 operator-(const Vector3 &v) { return Vector3(-v.x(), -v.y(), -v.z()); }
inline Vector3 operator*(double k, const Vector3 &v) { return Vector3(k * v.x(), k * v.y(), k * v.z()); }
inline Vector3 operator*(const Vector3 &v, double k) { return Vector3(k * v.x(), k * v.y(), k * v.z()); }
inline Vector3 operator/(const Vector3 &v, double k) { return Vector3(v.x() / k, v.y() / k, v.z() / k); }
inline std::string to_string(const Vector3 &v) { return std::to_string(v.x()) + "," + std::to_string(v.y()) + "," + std::to_string(v.z()); }
inline std::ostream &operator<<(std::ostream &os, const Vector3 &v) { return os << v.x() << ',' << v.y() << ',' << v.z(); }

static_assert(sizeof(Vector3) == 3 * sizeof(double), "Vector3 must be a plain double[3]");

#endif
