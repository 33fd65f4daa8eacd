// Per-candidate arithmetic of SolidBoolean::isPointInMesh
// (src/solidboolean.cpp:66-87), bit-exact: segment/plane intersection
// (Vector3::intersectSegmentAndPlane, src/vector3.h:264-280), three edge normals
// (Vector3::normal, src/vector3.h:155-176), two sign tests, PositionKey
// quantisation (src/positionkey.cpp:32-37).
#pragma once
#include "sb_fp64.cuh"

#define SB_DBL_EPSILON 2.2204460492503131e-16
#define SB_DBL_MAX 1.7976931348623157e+308

// n / d, correctly rounded (= __ddiv_rn, bit for bit), for the quotient the classifier forms once per candidate: the ray
// parameter s = n / d of Vector3::intersectSegmentAndPlane (src/vector3.h:264-280), whose denominator is ~1e308 (the test
// axis is DBL_MAX long) and whose result is therefore SUBNORMAL.  The library division leaves its fast path for such operands
// (~100 instructions where the fast path has ~25).  Here: the denominator is scaled by 2^-600 (exact), the quotient q formed
// in the normal range (fast path, correctly rounded), and scaled back by one multiplication, which rounds a second time when
// the result is subnormal.  The two roundings differ from one only if q landed exactly on a midpoint of the subnormal grid
// without the true quotient being there: detected exactly (q - back == half a grid step), decided by the sign of the exact
// residual n - q d' (one FMA).  Everything else falls through to __ddiv_rn.  Checked against the host's division on 1.6e8
// operand pairs incl. 6e7 constructed ties (the harness is described in DESIGN section 4) and by every parity test.
#ifndef SB_RAY_DIV_FAST
#define SB_RAY_DIV_FAST 1
#endif
__device__ __forceinline__ double xdiv_huge_den(double n, double d)
{
#if SB_RAY_DIV_FAST && !defined(SB_HOST_SIM)
    const int en = (__double2hiint(n) >> 20) & 0x7ff, ed = (__double2hiint(d) >> 20) & 0x7ff;
    if (ed >= 1023 + 1000 && ed < 0x7ff && en >= 1023 - 200 && en <= 1023 + 200) {
        const double dS = __dmul_rn(d, 0x1p-600); // exact
        const double q = __ddiv_rn(n, dS);        // normal range
        if (fabs(q) >= 0x1p-470) {
            double s = __dmul_rn(q, 0x1p-600);
            if (fabs(s) < 0x1p-1022) {
                const double back = __dmul_rn(s, 0x1p600); // exact
                const double diff = __dsub_rn(q, back);    // exact
                if (fabs(diff) == 0x1p-475) {              // q on a midpoint of the subnormal grid (step 2^-1074 -> 2^-474 here)
                    const double r = __fma_rn(-q, dS, n);  // sign of n - q d': on which side of q the true quotient lies
                    if (r != 0.0) {
                        const bool above = (r > 0.0) == (dS > 0.0);
                        const double other = __dadd_rn(back, __dmul_rn(2.0, diff));
                        const double lo = back < other ? back : other, hi = back < other ? other : back;
                        s = __dmul_rn(above ? hi : lo, 0x1p-600);
                    }
                }
            }
            return s;
        }
    }
#endif
    return xdiv(n, d);
}

// Vector3::normal; the zero vector when |cross| <= DBL_EPSILON (Double::isZero).
__device__ __forceinline__ d3 tri_normal(const d3 &a, const d3 &b, const d3 &c)
{
    d3 ba = d3sub(b, a);
    d3 ca = d3sub(c, a);
    d3 cr = d3cross(ba, ca);
    double len2 = xadd(xadd(xmul(cr.x, cr.x), xmul(cr.y, cr.y)), xmul(cr.z, cr.z));
    double len = xsqrt(len2);
    if (fabs(len) <= SB_DBL_EPSILON)
        return {0.0, 0.0, 0.0};
    return {xdiv(cr.x, len), xdiv(cr.y, len), xdiv(cr.z, len)};
}

// testEnd = testPosition + testAxis for axis k of g_testAxisList
// (src/solidboolean.cpp:31-35, :53)
__device__ __forceinline__ d3 ray_end(const d3 &p, int axis)
{
    return {xadd(p.x, axis == 0 ? SB_DBL_MAX : SB_DBL_EPSILON),
            xadd(p.y, axis == 1 ? SB_DBL_MAX : SB_DBL_EPSILON),
            xadd(p.z, axis == 2 ? SB_DBL_MAX : SB_DBL_EPSILON)};
}

// true when the reference would insert PositionKey(hit) for this candidate
__device__ __forceinline__ bool ray_tri_hit(const d3 &p, const d3 &end,
    const d3 &t0, const d3 &t1, const d3 &t2, const d3 &nrm, d3 &hit)
{
    d3 u = d3sub(end, p);
    d3 w = d3sub(p, t0);
    double d = d3dot(nrm, u);
    d3 neg = {-nrm.x, -nrm.y, -nrm.z};
    double n = d3dot(neg, w);
    if (fabs(d) <= SB_DBL_EPSILON)
        return false;
    double s = xdiv_huge_den(n, d);
    // s < 0 || s > 1 || isnan(s) || isinf(s)  (src/vector3.h:274)
    if (!(s >= 0.0 && s <= 1.0))
        return false;
    hit = {xadd(p.x, xmul(s, u.x)), xadd(p.y, xmul(s, u.y)), xadd(p.z, xmul(s, u.z))};
    d3 n0 = tri_normal(hit, t0, t1);
    d3 n1 = tri_normal(hit, t1, t2);
    d3 n2 = tri_normal(hit, t2, t0);
    return d3dot(n0, n1) > 0 && d3dot(n0, n2) > 0;
}

// Same predicate, with a floating-point FILTER in front of the three
// normalisations.  Only the SIGNS of n0.n1 and n0.n2 are used by the reference,
// and normalising the edge cross products c_i (one sqrt + three divisions each)
// cannot change the sign of their dot product unless that dot product is within
// rounding error of zero.  So: when every |c_i|^2 is far above the zero-vector
// threshold (DBL_EPSILON^2) and |c0.c1| exceeds 1e-12 * sum|c0_k*c1_k| (the
// reference's own evaluation error is below 2e-15 of that sum), the sign of the
// un-normalised dot IS the sign the reference computes; otherwise fall through
// to the exact sequence.  Bit-identical results, ~9 divisions and 3 square
// roots fewer per candidate.  (tests/test_hostsim.py fuzzes filtered == exact.)
__device__ __forceinline__ bool ray_tri_hit_filtered(const d3 &p, const d3 &end,
    const d3 &t0, const d3 &t1, const d3 &t2, const d3 &nrm, d3 &hit)
{
    d3 u = d3sub(end, p);
    d3 w = d3sub(p, t0);
    double d = d3dot(nrm, u);
    d3 neg = {-nrm.x, -nrm.y, -nrm.z};
    double n = d3dot(neg, w);
    if (fabs(d) <= SB_DBL_EPSILON)
        return false;
    double s = xdiv_huge_den(n, d);
    if (!(s >= 0.0 && s <= 1.0))
        return false;
    hit = {xadd(p.x, xmul(s, u.x)), xadd(p.y, xmul(s, u.y)), xadd(p.z, xmul(s, u.z))};
    // the cross products Vector3::normal(hit, a, b) forms before normalising
    d3 e0 = d3sub(t0, hit), e1 = d3sub(t1, hit), e2 = d3sub(t2, hit);
    d3 c0 = d3cross(e0, e1), c1 = d3cross(e1, e2), c2 = d3cross(e2, e0);
    double l0 = xadd(xadd(xmul(c0.x, c0.x), xmul(c0.y, c0.y)), xmul(c0.z, c0.z));
    double l1 = xadd(xadd(xmul(c1.x, c1.x), xmul(c1.y, c1.y)), xmul(c1.z, c1.z));
    double l2 = xadd(xadd(xmul(c2.x, c2.x), xmul(c2.y, c2.y)), xmul(c2.z, c2.z));
    const double tiny = 1e-28; // >> DBL_EPSILON^2 = 4.9e-32: none of the normals is the zero vector
    if (l0 > tiny && l1 > tiny && l2 > tiny) {
        double s01 = d3dot(c0, c1), s02 = d3dot(c0, c2);
        double m01 = fabs(c0.x * c1.x) + fabs(c0.y * c1.y) + fabs(c0.z * c1.z);
        double m02 = fabs(c0.x * c2.x) + fabs(c0.y * c2.y) + fabs(c0.z * c2.z);
        // sign certain: relative margin 1e-12, and the normalised products stay
        // far from the subnormal range ((m/(|c0||c1|))^2 > 1e-200)
        bool k01 = fabs(s01) > 1e-12 * m01 && m01 * m01 > 1e-200 * (l0 * l1);
        bool k02 = fabs(s02) > 1e-12 * m02 && m02 * m02 > 1e-200 * (l0 * l2);
        if (k01 && !(s01 > 0))
            return false;
        if (k02 && !(s02 > 0))
            return false;
        if (k01 && k02)
            return true;
    }
    d3 n0 = tri_normal(hit, t0, t1);
    d3 n1 = tri_normal(hit, t1, t2);
    d3 n2 = tri_normal(hit, t2, t0);
    return d3dot(n0, n1) > 0 && d3dot(n0, n2) > 0;
}

// (long)(x * 100000) with x86-64 cvttsd2si semantics (NaN / out of range ->
// LLONG_MIN, the "integer indefinite" value).
__device__ __forceinline__ long long position_key(double x)
{
    double v = xmul(x, 100000.0);
#ifdef SB_HOST_SIM
    return (long long)(long)v;
#else
    if (!(v > -9223372036854775808.0 && v < 9223372036854775808.0))
        return (long long)0x8000000000000000ull;
    return __double2ll_rz(v);
#endif
}
